/*
 * opencl_c_on_cpp.h -- just enough of OpenCL C, in C++, to compile the reference's OWN kernel
 * files (resources/kernels/opencl/*.cl, examples/<x>/resources/kernels/<x>.cl) for the host CPU and
 * run them one work-item at a time.  TEST INFRASTRUCTURE ONLY (see oracle/lt_oracle.h): it exists
 * so that the hand restatement in lt_oracle.c can be checked against the reference's kernel TEXT,
 * since no OpenCL implementation exists in this image or on the GPU box.
 *
 * The kernel text itself is not edited by hand and not stored in this repository:
 * oracle/build_ref_cl.sh reads it where it lies under /root/reference, applies one mechanical
 * rewrite -- the OpenCL vector literal `(floatN)(...)` becomes the C++ constructor call
 * `floatN(...)` -- and writes the result into oracle/_ref/cl/ (git-ignored).
 *
 * What this header has to define is what OpenCL C leaves to the implementation:
 *   - the geometric built-ins.  They are the OpenCL specification's own formulae evaluated in
 *     FP32 with one rounding per operation, left to right, never contracted:
 *       dot(a,b)     = ((a.x*b.x + a.y*b.y) + a.z*b.z) + a.w*b.w
 *       cross(a,b)   = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x, 0)
 *       length(a)    = sqrt(dot(a,a));  distance(a,b) = length(a - b);  normalize(a) = a / length(a)
 *       clamp(x,l,h) = fmin(fmax(x,l),h)
 *   - the math built-ins: the host libm (sin/cos/fmod/floor/sqrt/fmax/fabs), float or double by
 *     C++ overload resolution, which here selects exactly what OpenCL C's overloading selects
 *     (a `float` argument -> the float function, an unsuffixed literal or a double operand -> double).
 *   - the work-item functions: a thread-local record filled by the driver (ref_cl_host.cpp).
 * Scalar arithmetic follows the usual arithmetic conversions, identical in OpenCL C (with
 * cl_khr_fp64 enabled, as every one of these kernel files does) and in C++.
 * Mixing a float vector with a double scalar is an error in OpenCL C; the operators below are
 * deleted for double so that such an expression would fail to compile here as well.
 */
#ifndef LT_OPENCL_C_ON_CPP_H
#define LT_OPENCL_C_ON_CPP_H

#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define __kernel
#define __global
#define __constant const
#define __local
#define __private

typedef unsigned int uint;
typedef unsigned short ushort;
typedef unsigned char uchar;

/* ---- work-item functions ---- */
struct clshim_work_item {
  size_t global_id[3], global_size[3], local_id[3], local_size[3], group_id[3], num_groups[3];
};
extern thread_local clshim_work_item clshim_wi;
inline size_t get_global_id(uint d) { return clshim_wi.global_id[d]; }
inline size_t get_global_size(uint d) { return clshim_wi.global_size[d]; }
inline size_t get_local_id(uint d) { return clshim_wi.local_id[d]; }
inline size_t get_local_size(uint d) { return clshim_wi.local_size[d]; }
inline size_t get_group_id(uint d) { return clshim_wi.group_id[d]; }
inline size_t get_num_groups(uint d) { return clshim_wi.num_groups[d]; }

/* ---- vector types (only what the kernel files use: float2/3/4, the .xy swizzle) ---- */
struct float2 {
  float x, y;
  float2() = default;
  template <class A, class B> float2(A a, B b) : x((float)a), y((float)b) {}
  template <class A> explicit float2(A a) : x((float)a), y((float)a) {}
};
struct float3 {
  float x, y, z;
  float3() = default;
  template <class A, class B, class C> float3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
  template <class A> explicit float3(A a) : x((float)a), y((float)a), z((float)a) {}
};
struct clshim_swizzle_xy { /* overlays .x,.y of a float4; reading it yields a float2 */
  float x, y;
  operator float2() const { return float2(x, y); }
};
struct float4 {
  union {
    struct { float x, y, z, w; };
    clshim_swizzle_xy xy;
  };
  float4() = default;
  template <class A, class B, class C, class D> float4(A a, B b, C c, D d) : x((float)a), y((float)b), z((float)c), w((float)d) {}
  template <class D> float4(float3 v, D d) : x(v.x), y(v.y), z(v.z), w((float)d) {}
  template <class A> explicit float4(A a) : x((float)a), y((float)a), z((float)a), w((float)a) {}
};

#define CLSHIM_VEC_OPS(V, ...)                                                                   \
  inline V operator+(V a, V b) { return CLSHIM_ZIP(V, +); }                                      \
  inline V operator-(V a, V b) { return CLSHIM_ZIP(V, -); }                                      \
  inline V operator-(V a) { return CLSHIM_NEG(V); }                                              \
  inline V operator*(V a, V b) { return CLSHIM_ZIP(V, *); }                                      \
  inline V operator/(V a, V b) { return CLSHIM_ZIP(V, /); }                                      \
  inline V operator*(V a, float s) { return a * V(s); }                                          \
  inline V operator*(float s, V a) { return V(s) * a; }                                          \
  inline V operator/(V a, float s) { return a / V(s); }                                          \
  inline V& operator+=(V& a, V b) { a = a + b; return a; }                                       \
  inline V& operator-=(V& a, V b) { a = a - b; return a; }                                       \
  inline V& operator*=(V& a, float s) { a = a * s; return a; }                                   \
  V operator*(V, double) = delete;                                                               \
  V operator*(double, V) = delete;                                                               \
  V operator/(V, double) = delete;

#define CLSHIM_ZIP(V, op) V(a.x op b.x, a.y op b.y)
#define CLSHIM_NEG(V) V(-a.x, -a.y)
CLSHIM_VEC_OPS(float2)
#undef CLSHIM_ZIP
#undef CLSHIM_NEG
#define CLSHIM_ZIP(V, op) V(a.x op b.x, a.y op b.y, a.z op b.z)
#define CLSHIM_NEG(V) V(-a.x, -a.y, -a.z)
CLSHIM_VEC_OPS(float3)
#undef CLSHIM_ZIP
#undef CLSHIM_NEG
#define CLSHIM_ZIP(V, op) V(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w)
#define CLSHIM_NEG(V) V(-a.x, -a.y, -a.z, -a.w)
CLSHIM_VEC_OPS(float4)
#undef CLSHIM_ZIP
#undef CLSHIM_NEG

/* ---- math built-ins: host libm through C++ overload resolution ---- */
using std::cos;
using std::fabs;
using std::floor;
using std::fmax;
using std::fmin;
using std::fmod;
using std::sin;
using std::sqrt;

/* ---- geometric built-ins (OpenCL 3.0 C spec 6.15.5), FP32, one rounding per operation ---- */
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(float4 a, float4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline float3 cross(float3 a, float3 b) {
  return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float4 cross(float4 a, float4 b) {
  return float4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0f);
}
inline float length(float3 a) { return std::sqrt(dot(a, a)); }
inline float length(float4 a) { return std::sqrt(dot(a, a)); }
inline float distance(float3 a, float3 b) { return length(a - b); }
inline float distance(float4 a, float4 b) { return length(a - b); }
inline float3 normalize(float3 a) { return a / length(a); }
inline float4 normalize(float4 a) { return a / length(a); }
inline float clamp(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }

#endif

// Test plug-in with the reference's kernel ABI (resources/kernels/cuda/basic.cu:331-340): echoes one value
// from each of the five input buffers so the test can check that every buffer, the image size and the
// launch shape arrive intact.  Own code; only the struct layouts are the reference's.
struct LinearBVHNode { float boundsMin[3]; float boundsMax[3]; int offset; unsigned short primitiveCount; unsigned char axis; unsigned char pad; };
struct Primitive { float positionA[3], positionB[3], positionC[3], normalA[3], normalB[3], normalC[3]; int materialIndex; };
struct Material { float diffuse[3]; float ior; float dissolve; float emission[3]; };
struct LightContainer { unsigned int count; unsigned int primitives[64]; };
struct Camera { float position[3]; float yaw, pitch, roll; unsigned int frameCount; };

__device__ void echo(LinearBVHNode* nodes, Primitive* prims, Material* mats, LightContainer* lights, Camera* cam,
                     float* out, int idx, int idy, int width, int height, int depth) {
  if (idx >= width || idy >= height) return;
  int id = (idy * width + idx) * depth;
  int p = (idy * width + idx) % 42;
  out[id + 0] = cam->position[2] + (float)idx + (float)cam->frameCount;
  out[id + 1] = mats[prims[p].materialIndex].diffuse[1] + prims[p].positionB[0];
  out[id + 2] = (float)lights->count + nodes[0].boundsMax[1] + (float)idy;
}

extern "C" __global__ void linearKernel(LinearBVHNode* nodes, Primitive* prims, Material* mats, LightContainer* lights,
                                        Camera* cam, float* out, int width, int height, int depth) {
  echo(nodes, prims, mats, lights, cam, out, blockIdx.x * blockDim.x + threadIdx.x,
       blockIdx.y * blockDim.y + threadIdx.y, width, height, depth);
}

extern "C" __global__ void tileKernel(LinearBVHNode* nodes, Primitive* prims, Material* mats, LightContainer* lights,
                                      Camera* cam, float* out, int width, int height, int depth) {
  int block = blockIdx.y * gridDim.x + blockIdx.x;
  echo(nodes, prims, mats, lights, cam, out, (block % gridDim.x) * blockDim.x + threadIdx.x,
       (block / gridDim.x) * blockDim.y + threadIdx.y, width, height, depth);
}

// obj_loader.cpp -- Wavefront OBJ/MTL reader (see include/lens_trace/obj_loader.h).
#include "lens_trace/obj_loader.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <map>

namespace tinyobj {
namespace {

inline bool isSpace(char c) { return c == ' ' || c == '\t'; }
inline bool isDigit(char c) { return c >= '0' && c <= '9'; }
inline bool isEnd(char c) { return c == '\0' || c == '\n' || c == '\r'; }

void skipSpace(const char*& p) {
  while (isSpace(*p)) p++;
}

// Decimal -> double with the same arithmetic as the loader the reference vendors
// (tiny_obj_loader.h:837-960): digits are accumulated in a double, fraction digit k is weighted by
// 10^-k (table for k < 8, pow beyond), a decimal exponent e is applied as ldexp(m * 5^e, e).  This
// is not correctly rounded, which is the point: the floats match the reference's bit for bit.
bool parseDecimal(const char* s, const char* end, double* out) {
  if (s >= end) return false;
  double mant = 0.0;
  int exponent = 0;
  bool neg = false, expNeg = false, leadingDot = false;
  const char* c = s;
  if (*c == '+' || *c == '-') {
    neg = (*c == '-');
    c++;
    if (c != end && *c == '.') leadingDot = true;
  } else if (*c == '.') {
    leadingDot = true;
  } else if (!isDigit(*c)) {
    return false;
  }
  int read = 0;
  if (!leadingDot) {
    while (c != end && isDigit(*c)) {
      mant = mant * 10 + (int)(*c - '0');
      c++;
      read++;
    }
    if (read == 0) return false;
  }
  if (c != end && *c == '.') {
    static const double kPow[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
    c++;
    read = 1;
    while (c != end && isDigit(*c)) {
      mant += (int)(*c - '0') * (read < 8 ? kPow[read] : pow(10.0, -read));
      read++;
      c++;
    }
  }
  if (c != end && (*c == 'e' || *c == 'E')) {
    c++;
    if (c != end && (*c == '+' || *c == '-')) {
      expNeg = (*c == '-');
      c++;
    } else if (c == end || !isDigit(*c)) {
      return false;
    }
    read = 0;
    while (c != end && isDigit(*c)) {
      if (exponent > 214748364) return false;
      exponent = exponent * 10 + (int)(*c - '0');
      c++;
      read++;
    }
    if (read == 0) return false;
    if (expNeg) exponent = -exponent;
  }
  double v = exponent ? ldexp(mant * pow(5.0, exponent), exponent) : mant;
  *out = neg ? -v : v;
  return true;
}

float parseReal(const char*& p, double def = 0.0) {
  skipSpace(p);
  const char* e = p;
  while (!isSpace(*e) && !isEnd(*e)) e++;
  double v = def;
  parseDecimal(p, e, &v);
  p = e;
  return (float)v;
}

int parseInt(const char*& p) {
  skipSpace(p);
  int v = atoi(p);
  while (!isSpace(*p) && !isEnd(*p)) p++;
  return v;
}

std::string parseName(const char*& p) {
  skipSpace(p);
  const char* e = p;
  while (!isEnd(*e)) e++;
  while (e > p && isSpace(e[-1])) e--;
  std::string s(p, e);
  p = e;
  return s;
}

// OBJ indices are 1-based; negative ones count back from the current end.
bool fixIndex(int idx, int n, int* out) {
  if (idx > 0) { *out = idx - 1; return true; }
  if (idx < 0) { *out = n + idx; return true; }
  return false;
}

struct Corner { int v, vt, vn; };

bool parseCorner(const char*& p, int nv, int nvt, int nvn, Corner* c) {
  c->v = c->vt = c->vn = -1;
  if (!fixIndex(atoi(p), nv, &c->v)) return false;
  while (*p != '/' && !isSpace(*p) && !isEnd(*p)) p++;
  if (*p != '/') return true;
  p++;
  if (*p == '/') {  // v//vn
    p++;
    if (!fixIndex(atoi(p), nvn, &c->vn)) return false;
    while (!isSpace(*p) && !isEnd(*p)) p++;
    return true;
  }
  if (!fixIndex(atoi(p), nvt, &c->vt)) return false;  // v/vt[/vn]
  while (*p != '/' && !isSpace(*p) && !isEnd(*p)) p++;
  if (*p != '/') return true;
  p++;
  if (!fixIndex(atoi(p), nvn, &c->vn)) return false;
  while (!isSpace(*p) && !isEnd(*p)) p++;
  return true;
}

// crossing-number point-in-polygon test on a projected triangle
bool insideTriangle2D(const float px[3], const float py[3], float tx, float ty) {
  bool in = false;
  for (int i = 0, j = 2; i < 3; j = i++) {
    if (((py[i] > ty) != (py[j] > ty)) && (tx < (px[j] - px[i]) * (ty - py[i]) / (py[j] - py[i]) + px[i])) in = !in;
  }
  return in;
}

// Ear clipping of a polygon with > 4 corners.  Decision sequence (projection axes from the first
// non-degenerate corner, signed area, candidate ear at a moving cursor, reject reflex corners and
// ears containing another vertex, bounded retries) mirrors tiny_obj_loader.h:1505-1716.
template <class Emit>
void clipEars(const std::vector<Corner>& face, const std::vector<float>& v, Emit emit) {
  size_t n = face.size();
  auto valid = [&](int vi) { return vi >= 0 && (size_t)(3 * vi + 2) < v.size(); };
  size_t ax0 = 1, ax1 = 2;
  for (size_t k = 0; k < n; k++) {
    int a = face[k % n].v, b = face[(k + 1) % n].v, c = face[(k + 2) % n].v;
    if (!valid(a) || !valid(b) || !valid(c)) continue;
    float e0x = v[3 * b] - v[3 * a], e0y = v[3 * b + 1] - v[3 * a + 1], e0z = v[3 * b + 2] - v[3 * a + 2];
    float e1x = v[3 * c] - v[3 * b], e1y = v[3 * c + 1] - v[3 * b + 1], e1z = v[3 * c + 2] - v[3 * b + 2];
    float cx = fabsf(e0y * e1z - e0z * e1y), cy = fabsf(e0z * e1x - e0x * e1z), cz = fabsf(e0x * e1y - e0y * e1x);
    const float eps = 1.1920929e-07f;
    if (cx > eps || cy > eps || cz > eps) {
      if (!(cx > cy && cx > cz)) {
        ax0 = 0;
        if (cz > cx && cz > cy) ax1 = 1;
      }
      break;
    }
  }
  float area = 0;
  for (size_t k = 0; k < n; k++) {
    int a = face[k % n].v, b = face[(k + 1) % n].v;
    if (!valid(a) || !valid(b)) continue;
    area += (v[3 * a + ax0] * v[3 * b + ax1] - v[3 * a + ax1] * v[3 * b + ax0]) * 0.5f;
  }
  std::vector<int> remaining(n);  // positions into `face`
  for (size_t k = 0; k < n; k++) remaining[k] = (int)k;
  size_t cursor = 0, budget = n, previous = n;
  while (remaining.size() > 3 && budget > 0) {
    size_t m = remaining.size();
    if (cursor >= m) cursor -= m;
    if (previous != m) {
      previous = m;
      budget = m;
    } else {
      budget--;
    }
    int pos[3];
    float px[3], py[3];
    for (int k = 0; k < 3; k++) {
      pos[k] = remaining[(cursor + k) % m];
      int vi = face[pos[k]].v;
      px[k] = valid(vi) ? v[3 * vi + ax0] : 0.0f;
      py[k] = valid(vi) ? v[3 * vi + ax1] : 0.0f;
    }
    float cross = (px[1] - px[0]) * (py[2] - py[1]) - (py[1] - py[0]) * (px[2] - px[1]);
    if (cross * area < 0.0f) {  // reflex corner
      cursor++;
      continue;
    }
    bool blocked = false;
    for (size_t o = 3; o < m && !blocked; o++) {
      int vi = face[remaining[(cursor + o) % m]].v;
      if (!valid(vi)) continue;
      blocked = insideTriangle2D(px, py, v[3 * vi + ax0], v[3 * vi + ax1]);
    }
    if (blocked) {
      cursor++;
      continue;
    }
    emit(pos[0], pos[1], pos[2]);
    remaining.erase(remaining.begin() + (cursor + 1) % m);
  }
  if (remaining.size() == 3) emit(remaining[0], remaining[1], remaining[2]);
}

void initMaterial(material_t* m) {
  m->name.clear();
  for (int i = 0; i < 3; i++) m->ambient[i] = m->diffuse[i] = m->specular[i] = m->transmittance[i] = m->emission[i] = 0.f;
  m->illum = 0;
  m->dissolve = 1.f;
  m->shininess = 1.f;
  m->ior = 1.f;
}

bool readLine(FILE* f, std::string* line) {
  line->clear();
  char buf[4096];
  bool any = false;
  while (fgets(buf, sizeof buf, f)) {
    any = true;
    size_t n = strlen(buf);
    line->append(buf, n);
    if (n > 0 && buf[n - 1] == '\n') break;
  }
  while (!line->empty() && (line->back() == '\r' || line->back() == '\n')) line->pop_back();
  return any;
}

void loadMtl(const std::string& path, std::vector<material_t>* materials, std::map<std::string, int>* byName,
             std::string* warn) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    if (warn) *warn += "Material file [ " + path + " ] not found.\n";
    return;
  }
  material_t cur;
  initMaterial(&cur);
  bool have = false, hasD = false;
  std::string line;
  while (readLine(f, &line)) {
    const char* p = line.c_str();
    skipSpace(p);
    if (isEnd(*p) || *p == '#') continue;
    if (!strncmp(p, "newmtl", 6) && isSpace(p[6])) {
      if (have) {
        (*byName)[cur.name] = (int)materials->size();
        materials->push_back(cur);
      }
      initMaterial(&cur);
      hasD = false;
      p += 7;
      cur.name = parseName(p);
      have = true;
    } else if (p[0] == 'K' && p[1] == 'a' && isSpace(p[2])) {
      p += 2; for (int i = 0; i < 3; i++) cur.ambient[i] = parseReal(p);
    } else if (p[0] == 'K' && p[1] == 'd' && isSpace(p[2])) {
      p += 2; for (int i = 0; i < 3; i++) cur.diffuse[i] = parseReal(p);
    } else if (p[0] == 'K' && p[1] == 's' && isSpace(p[2])) {
      p += 2; for (int i = 0; i < 3; i++) cur.specular[i] = parseReal(p);
    } else if (p[0] == 'K' && p[1] == 'e' && isSpace(p[2])) {
      p += 2; for (int i = 0; i < 3; i++) cur.emission[i] = parseReal(p);
    } else if ((p[0] == 'K' && p[1] == 't' && isSpace(p[2])) || (p[0] == 'T' && p[1] == 'f' && isSpace(p[2]))) {
      p += 2; for (int i = 0; i < 3; i++) cur.transmittance[i] = parseReal(p);
    } else if (p[0] == 'N' && p[1] == 'i' && isSpace(p[2])) {
      p += 2; cur.ior = parseReal(p);
    } else if (p[0] == 'N' && p[1] == 's' && isSpace(p[2])) {
      p += 2; cur.shininess = parseReal(p);
    } else if (!strncmp(p, "illum", 5) && isSpace(p[5])) {
      p += 6; cur.illum = parseInt(p);
    } else if (p[0] == 'd' && isSpace(p[1])) {
      p += 1; cur.dissolve = parseReal(p); hasD = true;
    } else if (p[0] == 'T' && p[1] == 'r' && isSpace(p[2])) {
      p += 2;
      float tr = parseReal(p);
      if (!hasD) cur.dissolve = 1.0f - tr;  // `d` wins when both are present
    }
  }
  if (have) {
    (*byName)[cur.name] = (int)materials->size();
    materials->push_back(cur);
  }
  fclose(f);
}

}  // namespace

bool LoadObj(attrib_t* attrib, std::vector<shape_t>* shapes, std::vector<material_t>* materials, std::string* warn,
             std::string* err, const char* filename) {
  attrib->vertices.clear();
  attrib->normals.clear();
  attrib->texcoords.clear();
  shapes->clear();
  materials->clear();
  FILE* f = fopen(filename, "rb");
  if (!f) {
    if (err) *err += std::string("Cannot open file [") + filename + "]\n";
    return false;
  }
  std::vector<char> ioBuffer(1 << 20);
  setvbuf(f, ioBuffer.data(), _IOFBF, ioBuffer.size());
  std::string baseDir;
  {
    std::string fn(filename);
    size_t s = fn.find_last_of('/');
    if (s != std::string::npos) baseDir = fn.substr(0, s + 1);
  }
  std::map<std::string, int> materialByName;
  std::vector<float>& v = attrib->vertices;
  shape_t shape;
  int material = -1;
  auto flushShape = [&]() {
    if (!shape.mesh.indices.empty()) shapes->push_back(shape);
    shape = shape_t();
  };
  std::vector<Corner> face;
  std::string line;
  while (readLine(f, &line)) {
    const char* p = line.c_str();
    skipSpace(p);
    if (isEnd(*p) || *p == '#') continue;
    if (p[0] == 'v' && isSpace(p[1])) {
      p += 2;
      for (int i = 0; i < 3; i++) v.push_back(parseReal(p));
    } else if (p[0] == 'v' && p[1] == 'n' && isSpace(p[2])) {
      p += 3;
      for (int i = 0; i < 3; i++) attrib->normals.push_back(parseReal(p));
    } else if (p[0] == 'v' && p[1] == 't' && isSpace(p[2])) {
      p += 3;
      for (int i = 0; i < 2; i++) attrib->texcoords.push_back(parseReal(p));
    } else if (p[0] == 'f' && isSpace(p[1])) {
      p += 2;
      face.clear();
      skipSpace(p);
      bool ok = true;
      while (!isEnd(*p)) {
        Corner c;
        if (!parseCorner(p, (int)v.size() / 3, (int)attrib->texcoords.size() / 2, (int)attrib->normals.size() / 3, &c)) {
          ok = false;
          break;
        }
        face.push_back(c);
        skipSpace(p);
      }
      if (!ok) {
        if (err) *err += "Failed parse `f' line(e.g. zero value for face index).\n";
        fclose(f);
        return false;
      }
      size_t n = face.size();
      if (n < 3) {
        if (warn) *warn += "Degenerated face found\n.";
        continue;
      }
      auto emit = [&](int a, int b, int c) {
        const int ids[3] = {a, b, c};
        for (int k = 0; k < 3; k++) {
          index_t idx = {face[ids[k]].v, face[ids[k]].vn, face[ids[k]].vt};
          shape.mesh.indices.push_back(idx);
        }
        shape.mesh.num_face_vertices.push_back(3);
        shape.mesh.material_ids.push_back(material);
      };
      if (n == 3) {
        emit(0, 1, 2);
      } else if (n == 4) {
        // split along the shorter diagonal; a tie takes the 1-3 diagonal (tiny_obj_loader.h:1447-1487)
        bool valid = true;
        for (int k = 0; k < 4; k++) valid = valid && face[k].v >= 0 && (size_t)(3 * face[k].v + 2) < v.size();
        if (!valid) {
          if (warn) *warn += "Face with invalid vertex index found.\n";
          continue;
        }
        const float* p0 = &v[3 * face[0].v];
        const float* p1 = &v[3 * face[1].v];
        const float* p2 = &v[3 * face[2].v];
        const float* p3 = &v[3 * face[3].v];
        float e02x = p2[0] - p0[0], e02y = p2[1] - p0[1], e02z = p2[2] - p0[2];
        float e13x = p3[0] - p1[0], e13y = p3[1] - p1[1], e13z = p3[2] - p1[2];
        float sqr02 = e02x * e02x + e02y * e02y + e02z * e02z;
        float sqr13 = e13x * e13x + e13y * e13y + e13z * e13z;
        if (sqr02 < sqr13) {
          emit(0, 1, 2);
          emit(0, 2, 3);
        } else {
          emit(0, 1, 3);
          emit(1, 2, 3);
        }
      } else {
        // Polygons with more than four corners (cornell_box.obj has an 8-corner ceiling ring):
        // ear clipping with the same decisions as the loader the reference vendors
        // (tiny_obj_loader.h:1505-1716), so the triangles -- and primitive numbering -- agree.
        clipEars(face, v, emit);
      }
    } else if (!strncmp(p, "usemtl", 6) && isSpace(p[6])) {
      p += 7;
      std::string name = parseName(p);
      std::map<std::string, int>::iterator it = materialByName.find(name);
      if (it != materialByName.end()) {
        material = it->second;
      } else {
        material = -1;
        if (warn) *warn += "material [ '" + name + "' ] not found in .mtl\n";
      }
    } else if (!strncmp(p, "mtllib", 6) && isSpace(p[6])) {
      p += 7;
      std::string names = parseName(p);
      // first file that loads wins; names are space separated
      size_t pos = 0;
      while (pos < names.size()) {
        size_t e = names.find(' ', pos);
        if (e == std::string::npos) e = names.size();
        std::string one = names.substr(pos, e - pos);
        pos = e + 1;
        if (one.empty()) continue;
        size_t before = materials->size();
        loadMtl(baseDir + one, materials, &materialByName, warn);
        if (materials->size() > before) break;
      }
    } else if ((p[0] == 'g' || p[0] == 'o') && (isSpace(p[1]) || isEnd(p[1]))) {
      flushShape();
      p += 1;
      shape.name = parseName(p);
    }
    // s, l, p, vp, ...: ignored
  }
  flushShape();
  fclose(f);
  return true;
}

}  // namespace tinyobj

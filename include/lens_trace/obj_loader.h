// obj_loader.h -- small Wavefront OBJ/MTL reader (this project's own code).
// The type names live in namespace tinyobj only so that application code written against the
// reference's Model accessors (getAttrib/getShapes/getIndex) keeps compiling; nothing of
// tinyobjloader is included.  Behaviour that matters for parity (triangulation of quads by the
// shorter diagonal, decimal parsing, material defaults) follows tinyobjloader v2.0.0rc as vendored
// by the reference (include/tinyobjloader/tiny_obj_loader.h:837-960,1385-1500).
#pragma once
#include <string>
#include <vector>

namespace tinyobj {

struct index_t {
  int vertex_index;
  int normal_index;
  int texcoord_index;
};

struct attrib_t {
  std::vector<float> vertices;   // xyz
  std::vector<float> normals;    // xyz
  std::vector<float> texcoords;  // uv
};

struct mesh_t {
  std::vector<index_t> indices;
  std::vector<unsigned char> num_face_vertices;
  std::vector<int> material_ids;
};

struct shape_t {
  std::string name;
  mesh_t mesh;
};

struct material_t {
  std::string name;
  float ambient[3];
  float diffuse[3];
  float specular[3];
  float transmittance[3];
  float emission[3];
  float shininess;
  float ior;
  float dissolve;
  int illum;
};

// Reads `filename` (and the .mtl files it names, relative to its directory).  Faces are
// triangulated.  Returns false when the file cannot be opened; problems are appended to warn/err.
bool LoadObj(attrib_t* attrib, std::vector<shape_t>* shapes, std::vector<material_t>* materials, std::string* warn,
             std::string* err, const char* filename);

}  // namespace tinyobj

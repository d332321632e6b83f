// model.cpp -- see include/lens_trace/model.h (reference behaviour: src/model.cpp:5-82).
#include "lens_trace/model.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>

Model::Model(std::string fileName) {
  this->uniqueId = lt::newObjectId();
  this->fileName = Resource::findResource(fileName);
  this->success = tinyobj::LoadObj(&this->attrib, &this->shapes, &this->materials, &this->warning, &this->error,
                                   this->fileName.c_str());
  this->checkError();

  const std::vector<float>& V = this->attrib.vertices;
  const std::vector<float>& N = this->attrib.normals;
  size_t total = 0;
  for (size_t s = 0; s < shapes.size(); s++) total += shapes[s].mesh.num_face_vertices.size();
  primitiveInfoList.reserve(total);

  for (size_t s = 0; s < shapes.size(); s++) {
    const tinyobj::mesh_t& mesh = shapes[s].mesh;
    size_t cursor = 0;
    for (size_t f = 0; f < mesh.num_face_vertices.size(); f++) {
      PrimitiveInfo info;
      memset(&info, 0, sizeof info);
      float* pos[3] = {info.positionA, info.positionB, info.positionC};
      float* nrm[3] = {info.normalA, info.normalB, info.normalC};
      for (int c = 0; c < 3; c++) {
        tinyobj::index_t idx = mesh.indices[cursor + c];
        for (int k = 0; k < 3; k++) {
          pos[c][k] = V[3 * idx.vertex_index + k];
          // the reference indexes normals unchecked (src/model.cpp:27); a face without vn gets zeros here
          nrm[c][k] = (idx.normal_index >= 0 && (size_t)(3 * idx.normal_index + k) < N.size())
                          ? N[3 * idx.normal_index + k]
                          : 0.0f;
        }
      }
      cursor += mesh.num_face_vertices[f];
      info.materialIndex = mesh.material_ids[f];
      for (int k = 0; k < 3; k++) {
        info.boundsMin[k] = std::min(std::min(info.positionA[k], info.positionB[k]), info.positionC[k]);
        info.boundsMax[k] = std::max(std::max(info.positionA[k], info.positionB[k]), info.positionC[k]);
        info.centroid[k] = 0.5f * info.boundsMin[k] + 0.5f * info.boundsMax[k];
      }
      primitiveInfoList.push_back(info);
    }
  }

  materialList.resize(materials.size());
  for (size_t m = 0; m < materials.size(); m++) {
    memcpy(materialList[m].diffuse, materials[m].diffuse, sizeof(float) * 3);
    materialList[m].ior = materials[m].ior;
    materialList[m].dissolve = materials[m].dissolve;
    memcpy(materialList[m].emission, materials[m].emission, sizeof(float) * 3);
  }
}

Model::~Model() { lt::retireObjectId(this->uniqueId); }

std::string Model::getFileName() { return this->fileName; }

bool Model::checkError() {
  if (!this->warning.empty()) printf("%s\n", this->warning.c_str());
  if (!this->error.empty()) printf("%s\n", this->error.c_str());
  return this->success;
}

tinyobj::attrib_t Model::getAttrib() { return this->attrib; }
std::vector<tinyobj::shape_t> Model::getShapes() { return this->shapes; }
std::vector<PrimitiveInfo>* Model::getPrimitiveInfoListP() { return &this->primitiveInfoList; }

uint64_t Model::getMaterialBufferSize() { return sizeof(Material) * this->materialList.size(); }
void* Model::getMaterialBuffer() { return this->materialList.data(); }

float* Model::getVertices() { return this->attrib.vertices.data(); }
uint32_t Model::getVertexCount() { return (uint32_t)this->attrib.vertices.size(); }

tinyobj::index_t Model::getIndex(uint32_t index) {
  for (size_t s = 0; s < this->shapes.size(); s++) {
    size_t n = this->shapes[s].mesh.indices.size();
    if (index < n) return this->shapes[s].mesh.indices[index];
    index -= (uint32_t)n;
  }
  tinyobj::index_t none = {-1, -1, -1};
  return none;
}

uint32_t Model::getIndexCount() {
  uint32_t n = 0;
  for (size_t s = 0; s < this->shapes.size(); s++) n += (uint32_t)this->shapes[s].mesh.indices.size();
  return n;
}

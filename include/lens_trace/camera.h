// camera.h -- camera state + the 28-byte buffer the kernels read
// (reference: include/lens_trace/camera.h:10-41, src/camera.cpp:14-19: pos[3], yaw, pitch, roll, frameCount).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

class Camera {
private:
  struct Packed {  // the device-visible record, kept in sync by every setter
    float position[3];
    float yaw;
    float pitch;
    float roll;
    uint32_t frameCount;
  };
  Packed* packed;

public:
  Camera(float positionX, float positionY, float positionZ, float yaw = 0, float pitch = 0, float roll = 0);
  ~Camera();
  Camera(const Camera&) = delete;
  Camera& operator=(const Camera&) = delete;

  float getPositionX();
  float getPositionY();
  float getPositionZ();
  float getYaw();
  float getPitch();
  float getRoll();
  uint32_t getFrameCount();

  void setPosition(float x, float y, float z);
  void updatePosition(float x, float y, float z);

  void setRotation(float yaw, float pitch, float roll);
  void updateRotation(float yaw, float pitch, float roll);

  void incrementFrameCount();
  void resetFrameCount();
  void setFrameCount(uint32_t frameCount);  // B200 addition (multi-frame render calls)

  void* getCameraBuffer();
  uint64_t getCameraBufferSize();
};

// camera.cpp -- see include/lens_trace/camera.h (reference behaviour: src/camera.cpp).
#include "lens_trace/camera.h"

Camera::Camera(float positionX, float positionY, float positionZ, float yaw, float pitch, float roll) {
  packed = (Packed*)calloc(1, sizeof(Packed));
  packed->position[0] = positionX;
  packed->position[1] = positionY;
  packed->position[2] = positionZ;
  packed->yaw = yaw;
  packed->pitch = pitch;
  packed->roll = roll;
  packed->frameCount = 0;
}

Camera::~Camera() { free(packed); }

float Camera::getPositionX() { return packed->position[0]; }
float Camera::getPositionY() { return packed->position[1]; }
float Camera::getPositionZ() { return packed->position[2]; }
float Camera::getYaw() { return packed->yaw; }
float Camera::getPitch() { return packed->pitch; }
float Camera::getRoll() { return packed->roll; }
uint32_t Camera::getFrameCount() { return packed->frameCount; }

void Camera::setPosition(float x, float y, float z) {
  packed->position[0] = x;
  packed->position[1] = y;
  packed->position[2] = z;
}

void Camera::updatePosition(float x, float y, float z) {
  packed->position[0] += x;
  packed->position[1] += y;
  packed->position[2] += z;
}

void Camera::setRotation(float yaw, float pitch, float roll) {
  packed->yaw = yaw;
  packed->pitch = pitch;
  packed->roll = roll;
}

void Camera::updateRotation(float yaw, float pitch, float roll) {
  packed->yaw += yaw;
  packed->pitch += pitch;
  packed->roll += roll;
}

void Camera::incrementFrameCount() { packed->frameCount += 1; }
void Camera::resetFrameCount() { packed->frameCount = 0; }
void Camera::setFrameCount(uint32_t frameCount) { packed->frameCount = frameCount; }

void* Camera::getCameraBuffer() { return packed; }
uint64_t Camera::getCameraBufferSize() { return sizeof(Packed); }

static_assert(sizeof(float) * 7 == 28, "camera buffer is seven 32-bit words");

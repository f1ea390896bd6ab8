// CameraConfig.h — source-compatible with the reference's src/CameraConfig.h:5-17.
#pragma once

#include <cuda_runtime.h>

struct CameraConfig {
    CameraConfig(float3 lookFrom, float3 front, float vfov = 45.0f, float hfov = 45.0f) : front{front}, lookFrom{lookFrom}, vfov{vfov}, hfov{hfov} {}
    float3 front;
    float3 lookFrom;
    float vfov = 45.0f;
    float hfov = 45.0f;
    float pitch = 0.f;
    float yaw = 0.f;
};

// CameraConfig.h — the camera description the host hands to every DevicePathTracer by reference.
//
// Source-compatible with the reference's type of the same name (src/CameraConfig.h:5-17): same constructor, same public
// members, so its call sites (src/main.cu:40-41, src/RenderManager.h:146-183, the event handlers that turn pitch / yaw
// into a new `front`) compile unchanged.  Differences: pitch and yaw start at zero (the reference leaves them
// uninitialised), and the comparison below exists so the DevicePathTracer shim can tell when the borrowed object changed
// and the device-side camera has to be recomputed (ptcore_set_camera) — the reference re-evaluates the whole camera in
// every thread for every sample instead (src/camera.h:21-36).
#pragma once

#include <cuda_runtime.h>  // float3

struct CameraConfig {
    // Fields of view are in degrees and independent of each other and of the image's aspect ratio (camera.h:24-35).
    CameraConfig(float3 lookFrom, float3 front, float vfov = 45.0f, float hfov = 45.0f) : front{front}, lookFrom{lookFrom}, vfov{vfov}, hfov{hfov} {}

    float3 front;          // viewing direction; the camera looks at lookFrom + front
    float3 lookFrom;       // eye position
    float vfov = 45.0f;    // vertical field of view
    float hfov = 45.0f;    // horizontal field of view
    float pitch = 0.f;     // kept for the interactive front-ends, which derive `front` from them; not read by the renderer
    float yaw = 0.f;

    bool sameViewAs(const CameraConfig &o) const {
        return front.x == o.front.x && front.y == o.front.y && front.z == o.front.z && lookFrom.x == o.lookFrom.x && lookFrom.y == o.lookFrom.y &&
               lookFrom.z == o.lookFrom.z && vfov == o.vfov && hfov == o.hfov;
    }
};

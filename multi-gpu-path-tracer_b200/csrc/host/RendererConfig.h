// RendererConfig.h — the knobs of a render job.  Source-compatible with the reference (src/RendererConfig.h:6-37): every
// field keeps its name, type and default, so RenderManager's setters, the ArgumentLoader and the front-ends' event
// handlers read and write it unchanged.  Added at the end, with defaults that leave the reference's behaviour alone:
// the DYNAMIC scheduling mode with its tile size, and what the headless FileRenderer needs.
#pragma once

#include <cuda_runtime.h>  // dim3, float3

#include <string>

struct Resolution {
    unsigned int width;
    unsigned int height;
};

// How a frame is cut into tasks for the (GPU, stream) workers — src/Scheduling/TaskGenerator.h:58-80.
enum SchedulingAlgorithmType {
    FSFL,    // fixed size, fixed layout: equal cells, never changed
    DSFL,    // dynamic size, fixed layout: cell borders move towards equal render time
    DSDL,    // dynamic size, dynamic layout: time-weighted recursive bisection
    DYNAMIC, // addition: small tiles pulled from one shared counter by every worker
    LPT      // addition: pilot pass -> rays per 8x4 block -> blocks sorted by cost and dealt round-robin to the GPUs, each GPU takes its
             // blocks most-expensive-first in ONE persistent launch; finished blocks go to the master frame with peer stores (NVLink)
};

struct RendererConfig {
    // ---- identification / input ------------------------------------------------------------------------------
    std::string jobId = "0";                  // argv[1]; names the websocket session in the reference
    unsigned int samplesPerPixel = 10;        // spp: one sequential XORWOW stream per pixel, consumed sample by sample
    unsigned int recursionDepth = 3;          // bounces per path (camera::ray_color loop bound)
    std::string modelPath{};                  // argv[2]: .glb / .gltf / .obj (+ .ptscene here)
    // ---- resources ---------------------------------------------------------------------------------------------
    unsigned int gpuNumber = 1;               // devices 0 .. gpuNumber-1 of this box
    unsigned int streamsPerGpu = 1;           // worker threads (one CUDA stream each) per device
    // ---- image and scheduling ----------------------------------------------------------------------------------
    Resolution resolution{400, 400};
    SchedulingAlgorithmType algorithmType = FSFL;
    dim3 threadBlockSize{8, 8};               // kept for the API; the persistent kernel picks its own launch shape
    // ---- camera (CameraConfig is built from these) -------------------------------------------------------------
    float vfov = 45.0f;
    float hfov = 45.0f;
    float3 cameraLookFromVec{0.0f, 0.0f, 0.0f};
    float3 cameraFrontVec{1.0f, 0.0f, 0.0f};
    // ---- task layout -------------------------------------------------------------------------------------------
    unsigned int maxTasksInRow = 2;           // cells per row of the FSFL / DSFL grid (RenderManager::getTaskLayout)
    bool showTasks = true;                    // draw the task borders into the frame (RenderManager::markTasks)
    int kParam = 1;                           // stored by RenderManager::setKParameter; no scheduler reads it (nor in the reference)
    // ---- additions ---------------------------------------------------------------------------------------------
    unsigned int dynamicTileWidth = 256;      // DYNAMIC: tile size
    unsigned int dynamicTileHeight = 128;
    std::string outputPath{};                 // FileRenderer: P6 file written after the last frame (README.md:52-58)
    unsigned int framesToRender = 1;          // FileRenderer: stop after this many frames
};

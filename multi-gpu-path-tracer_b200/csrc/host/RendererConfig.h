// RendererConfig.h — source-compatible with the reference's src/RendererConfig.h:6-37
// (same field names, order and defaults); DYNAMIC is an added scheduling mode.
#pragma once

#include <cuda_runtime.h>

#include <string>

struct Resolution {
    unsigned int width;
    unsigned int height;
};

enum SchedulingAlgorithmType {
    FSFL,    // Fixed size tasks
    DSFL,    // Dynamic tasks with fixed layout
    DSDL,    // Dynamic layout tasks
    DYNAMIC  // addition: small tiles pulled from a shared counter by every GPU worker (work stealing)
};

struct RendererConfig {
    std::string jobId = "0";
    unsigned int samplesPerPixel = 10;
    unsigned int recursionDepth = 3;
    std::string modelPath{};
    unsigned int gpuNumber = 1;
    unsigned int streamsPerGpu = 1;
    Resolution resolution{400, 400};
    SchedulingAlgorithmType algorithmType = FSFL;
    dim3 threadBlockSize{8, 8};
    float vfov = 45.0f;
    float hfov = 45.0f;
    float3 cameraLookFromVec{0.0f, 0.0f, 0.0f};
    float3 cameraFrontVec{1.0f, 0.0f, 0.0f};
    unsigned int maxTasksInRow = 2;
    bool showTasks = true;
    int kParam = 1;
    // --- additions (defaults keep the reference's behaviour) ---
    unsigned int dynamicTileWidth = 256;   // DYNAMIC mode tile size
    unsigned int dynamicTileHeight = 128;
    std::string outputPath{};              // FileRenderer: where out.ppm goes (README.md:52-58)
    unsigned int framesToRender = 1;       // FileRenderer stops after this many frames
};

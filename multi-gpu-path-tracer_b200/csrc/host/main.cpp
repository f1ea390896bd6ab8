// main.cpp — `cuda_project <jobId> <modelPath> [flags]`: the reference's executable (src/main.cu:33-95) on the new core.
// Same flow: ArgumentLoader -> SceneLoader.load -> RenderManager -> { manager.renderFrame(); renderer.renderFrame(); } until
// the renderer says stop; the presenter is the headless FileRenderer (the websocket / GLFW front-ends are out of scope).
// Prints the reference's two timings ("initializing in: N ms", "Path Tracing took: N ms") and one JSON line of statistics.
#include "ArgumentLoader.h"
#include "CameraConfig.h"
#include "FileRenderer.h"
#include "GPUMonitor.h"
#include "HostScene.h"
#include "RenderManager.h"
#include "RendererConfig.h"

#include <chrono>
#include <cstdio>
#include <iostream>
#include <memory>
#include <string>
#include <thread>

int main(int argc, char **argv) {
    RendererConfig config;
    SceneLoader sceneLoader;
    ArgumentLoader argLoader(argc, argv);
    try {
        argLoader.loadArguments(config);
    } catch (const std::exception &e) {
        std::cerr << "argument error: " << e.what() << std::endl;
        return 2;
    }
    // reference src/main.cu:40: the camera the shipped models are authored for
    float3 lookFrom = argLoader.lookFromSet ? config.cameraLookFromVec : make_float3(0, 0, 0.5f);
    float3 front = argLoader.frontSet ? config.cameraFrontVec : make_float3(0, 0, -0.5f);
    CameraConfig cameraConfig(lookFrom, front, config.vfov, config.hfov);

    HostScene hScene;
    try {
        hScene = sceneLoader.load(config.modelPath);
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    std::cout << "Number of triangles: " << hScene.triangles.size() << std::endl;

    auto start_init = std::chrono::high_resolution_clock::now();
    RenderManager manager(config, hScene, cameraConfig, sceneLoader);
    auto stop_init = std::chrono::high_resolution_clock::now();
    std::cout << "initializing in: " << std::chrono::duration_cast<std::chrono::milliseconds>(stop_init - start_init).count() << "ms" << std::endl;

    FileRenderer fileRenderer(config, manager.getFramebuffer());
    Renderer &renderer = fileRenderer;

    // reference src/main.cu:76-77: the monitor runs beside the render loop (here only on request: the headless run is short)
    std::unique_ptr<MonitorThread> monitorObj;
    std::thread monitorThread;
    if (argLoader.monitor) {
        monitorObj.reset(new MonitorThread(renderer));
        monitorThread = std::thread(std::ref(*monitorObj));
    }

    double last_ms = 0;
    while (!renderer.shouldStopRendering()) {
        auto start = std::chrono::high_resolution_clock::now();
        manager.renderFrame();
        auto stop = std::chrono::high_resolution_clock::now();
        last_ms = std::chrono::duration<double, std::milli>(stop - start).count();
        std::cout << "Path Tracing took: " << (long long)last_ms << "ms" << std::endl;
        renderer.renderFrame();
        if (monitorObj) {  // reference src/main.cu:87-88
            manager.updateMetrics(*monitorObj);
            monitorObj->updateFps();
        }
    }
    if (monitorObj) {
        monitorObj->safeTerminate();
        monitorThread.join();
    }
    uint64_t rays = 0, samples = 0;
    for (auto &t : manager.tracers()) {
        PtStats s = t->stats();
        rays += s.rays;
        samples += s.samples;
    }
    const RenderManager::FrameStats &fs = manager.lastFrameStats();
    double frameSamples = (double)config.resolution.width * config.resolution.height * config.samplesPerPixel;
    std::string workerMs = "[", workerTiles = "[";
    for (size_t i = 0; i < fs.worker_ms.size(); i++) {
        char buf[64];
        snprintf(buf, sizeof buf, "%s%.3f", i ? ", " : "", fs.worker_ms[i]);
        workerMs += buf;
        snprintf(buf, sizeof buf, "%s%d", i ? ", " : "", fs.worker_tiles[i]);
        workerTiles += buf;
    }
    workerMs += "]";
    workerTiles += "]";
    printf("CUDA_PROJECT_JSON {\"frame_ms\": %.3f, \"msamples_per_s\": %.3f, \"imbalance\": %.4f, \"gpus\": %u, \"streams_per_gpu\": %u, \"scheduler\": %d, "
           "\"width\": %u, \"height\": %u, \"spp\": %u, \"depth\": %u, \"frames\": %u, \"total_samples\": %llu, \"total_rays\": %llu, "
           "\"worker_ms\": %s, \"worker_tiles\": %s, \"peer_access_failures\": %d}\n",
           fs.frame_ms, frameSamples / (fs.frame_ms / 1e3) / 1e6, fs.imbalance, config.gpuNumber, config.streamsPerGpu, (int)config.algorithmType,
           config.resolution.width, config.resolution.height, config.samplesPerPixel, config.recursionDepth, config.framesToRender,
           (unsigned long long)samples, (unsigned long long)rays, workerMs.c_str(), workerTiles.c_str(), DevicePathTracer::peerAccessFailures().load());
    manager.reset();
    return 0;
}

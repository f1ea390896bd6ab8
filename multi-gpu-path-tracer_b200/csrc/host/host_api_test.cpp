// host_api_test.cpp — exercises the C++ mirror of the reference's host API the way its main loop and its remote-control
// handlers do (src/main.cu:66-89; src/Renderer/RemoteRenderer/RemoteEventHandlers/RenderManagerEventHander.h:13-66):
// deferred setters between frames, scheduler switches, GPU/stream count changes, resolution changes, scene reload.
// Prints one line per check and exits non-zero on the first failure.  Needs a CUDA device (run by tests/test_gpu_host_api.py).
#include "CameraConfig.h"
#include "HostScene.h"
#include "RenderManager.h"
#include "RendererConfig.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static uint64_t checksum(const uint8_t *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
static int fails = 0;
static void check(bool ok, const char *what) {
    printf("%s %s\n", ok ? "ok  " : "FAIL", what);
    if (!ok) fails++;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: host_api_test <scene file>\n"); return 2; }
    SceneLoader loader;
    HostScene scene = loader.load(std::string(argv[1]));
    RendererConfig config;  // reference defaults: 400x400, 10 spp, depth 3, 1 GPU, FSFL, showTasks = true
    check(config.resolution.width == 400 && config.samplesPerPixel == 10 && config.recursionDepth == 3 && config.showTasks && config.algorithmType == FSFL,
          "RendererConfig defaults equal the reference's (src/RendererConfig.h:19-37)");
    config.showTasks = false;
    config.resolution = {192, 108};
    config.samplesPerPixel = 4;
    config.recursionDepth = 6;
    CameraConfig camera(make_float3(0, 0, 0.5f), make_float3(0, 0, -0.5f));
    int available = 0;
    cudaGetDeviceCount(&available);

    RenderManager manager(config, scene, camera, loader);
    manager.renderFrame();
    const size_t bytes = (size_t)192 * 108 * 3;
    const uint64_t base = checksum(manager.getCurrentFrame(), bytes);
    manager.renderFrame();
    check(checksum(manager.getCurrentFrame(), bytes) == base, "every frame re-renders the identical image (RNG state is never advanced, src/DevicePathTracer.h:80)");
    check(manager.getCurrentFrameWidth() == 192 && manager.getCurrentFrameHeight() == 108, "getCurrentFrameWidth/Height");
    const uint64_t yuvBase = checksum(manager.getYUVFrame(), (size_t)192 * 108 * 3 / 2);

    // scheduler switches never change pixels
    for (SchedulingAlgorithmType alg : {DSFL, DSDL, DYNAMIC, FSFL}) {
        manager.setSchedulingAlgorithm(alg);
        for (int f = 0; f < 3; f++) manager.renderFrame();
        check(checksum(manager.getCurrentFrame(), bytes) == base, alg == DSFL ? "DSFL" : alg == DSDL ? "DSDL" : alg == DYNAMIC ? "DYNAMIC" : "FSFL");
    }
    // more streams / more GPUs: same image (the reference renders tasks_[deviceIdx] from every stream thread; fixed here)
    manager.setGpuAndStreamNumber(available >= 2 ? 2 : 1, 3);
    manager.setSchedulingAlgorithm(DSFL);
    for (int f = 0; f < 4; f++) manager.renderFrame();
    check(checksum(manager.getCurrentFrame(), bytes) == base, "setGpuAndStreamNumber + DSFL re-tiling");
    check(checksum(manager.getYUVFrame(), (size_t)192 * 108 * 3 / 2) == yuvBase, "I420 frame follows the RGB frame");
    check((int)manager.tasks().size() == (available >= 2 ? 2 : 1) * 3, "one task per worker");

    // deferred parameter changes take effect at the next renderFrame
    manager.setSamplesPerPixel(2);
    manager.setRecursionDepth(3);
    manager.renderFrame();
    const uint64_t lowSpp = checksum(manager.getCurrentFrame(), bytes);
    check(lowSpp != base, "setSamplesPerPixel / setRecursionDepth change the image");
    manager.setSamplesPerPixel(4);
    manager.setRecursionDepth(6);
    manager.renderFrame();
    check(checksum(manager.getCurrentFrame(), bytes) == base, "... and restoring them restores it");

    manager.setResolution({96, 64});
    manager.setSchedulingAlgorithm(FSFL);
    manager.renderFrame();
    check(manager.getCurrentFrameWidth() == 96 && manager.getCurrentFrameHeight() == 64, "setResolution reallocates the framebuffer");
    const uint64_t small = checksum(manager.getCurrentFrame(), (size_t)96 * 64 * 3);
    manager.setGpuAndStreamNumber(1, 1);
    manager.renderFrame();
    check(checksum(manager.getCurrentFrame(), (size_t)96 * 64 * 3) == small, "same image at the new resolution with 1 GPU / 1 stream");

    // camera is borrowed by reference and snapshotted per task (src/DevicePathTracer.h:210)
    camera.lookFrom = make_float3(20.f, 10.f, 0.5f);
    manager.renderFrame();
    const uint64_t moved = checksum(manager.getCurrentFrame(), (size_t)96 * 64 * 3);
    check(moved != small, "editing the borrowed CameraConfig moves the camera at the next frame");
    camera.lookFrom = make_float3(0, 0, 0.5f);
    manager.renderFrame();
    check(checksum(manager.getCurrentFrame(), (size_t)96 * 64 * 3) == small, "... and back");

    // scene edit + updatePrimitives -> reloadWorld on every tracer
    std::vector<Triangle> saved = scene.triangles;
    scene.triangles.resize(12);  // walls + light only
    manager.updatePrimitives();
    manager.renderFrame();
    check(checksum(manager.getCurrentFrame(), (size_t)96 * 64 * 3) != small, "updatePrimitives reloads the world");
    scene.triangles = saved;
    manager.updatePrimitives();
    manager.renderFrame();
    check(checksum(manager.getCurrentFrame(), (size_t)96 * 64 * 3) == small, "... and reloading the original scene restores the image");

    // showTasks overlay touches only the host copy and only border pixels
    manager.setGpuAndStreamNumber(1, 4);
    manager.setShowTasks(true);
    manager.renderFrame();
    size_t black = 0, differ = 0;
    {
        std::vector<uint8_t> with(manager.getCurrentFrame(), manager.getCurrentFrame() + (size_t)96 * 64 * 3);
        manager.setShowTasks(false);
        manager.renderFrame();
        const uint8_t *without = manager.getCurrentFrame();
        for (size_t i = 0; i < (size_t)96 * 64; i++) {
            bool d = memcmp(&with[3 * i], &without[3 * i], 3) != 0;
            differ += d;
            black += d && with[3 * i] == 0 && with[3 * i + 1] == 0 && with[3 * i + 2] == 0;
        }
    }
    check(differ > 0 && differ == black && differ < (size_t)96 * 64 / 4, "showTasks draws black borders only (src/RenderManager.h:449-507)");

    const RenderManager::FrameStats &fs = manager.lastFrameStats();
    check(fs.frame_ms > 0 && fs.worker_ms.size() == 4 && fs.imbalance >= 1.0, "lastFrameStats: frame time, per-worker times, imbalance = max / mean");
    manager.reset();
    printf("%s (%d failure%s)\n", fails ? "FAILED" : "PASSED", fails, fails == 1 ? "" : "s");
    return fails ? 1 : 0;
}

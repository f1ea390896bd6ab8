// argument_loader_test.cpp — ArgumentLoader without a GPU: the reference's two positionals and defaults
// (src/ArgumentLoader.h:10-13), the added flags, and the error behaviour (exceptions with the offending argument named).
#include "ArgumentLoader.h"

#include <cstdio>
#include <string>
#include <vector>

static int failures = 0;
static void check(bool ok, const char *what) {
    if (!ok) {
        printf("FAIL %s\n", what);
        failures++;
    }
}

static RendererConfig parse(std::vector<std::string> args, ArgumentLoader **keep = nullptr) {
    static std::vector<std::string> store;
    static std::vector<char *> argv;
    store = std::move(args);
    store.insert(store.begin(), "cuda_project");
    argv.clear();
    for (auto &s : store) argv.push_back(s.data());
    static ArgumentLoader *loader = nullptr;
    delete loader;
    loader = new ArgumentLoader((int)argv.size(), argv.data());
    RendererConfig c;
    loader->loadArguments(c);
    if (keep) *keep = loader;
    return c;
}

template <class F>
static bool throws(F f, const char *needle) {
    try {
        f();
    } catch (const std::exception &e) {
        return std::string(e.what()).find(needle) != std::string::npos;
    }
    return false;
}

int main() {
    {
        RendererConfig c = parse({});
        check(c.jobId == "0" && c.modelPath == "models/cornel/cornell_box.gltf", "defaults of the two positionals (reference :10-13)");
        check(c.samplesPerPixel == 10 && c.recursionDepth == 3 && c.resolution.width == 400 && c.resolution.height == 400 && c.algorithmType == FSFL,
              "RendererConfig defaults are the reference's");
    }
    {
        RendererConfig c = parse({"42", "scene.glb"});
        check(c.jobId == "42" && c.modelPath == "scene.glb", "positionals: jobId, modelPath");
    }
    {
        ArgumentLoader *l = nullptr;
        RendererConfig c = parse({"7", "duck.glb", "--width", "1920", "--height", "1080", "--spp", "1024", "--depth", "10", "--gpus", "8", "--streams", "2", "--frames", "3", "--out",
                                  "o.ppm", "--vfov", "30.5", "--hfov", "50", "--show-tasks", "0", "--max-tasks-in-row", "4", "--block", "16x4", "--tile", "128x64", "--lookfrom",
                                  "1,2.5,-3", "--front", "0,0,-1", "--scheduler", "dynamic", "--monitor", "1"},
                                 &l);
        check(c.resolution.width == 1920 && c.resolution.height == 1080 && c.samplesPerPixel == 1024 && c.recursionDepth == 10, "image and sampling flags");
        check(c.gpuNumber == 8 && c.streamsPerGpu == 2 && c.framesToRender == 3 && c.outputPath == "o.ppm", "resources and output flags");
        check(c.vfov == 30.5f && c.hfov == 50.f && !c.showTasks && c.maxTasksInRow == 4, "camera angles, show-tasks, layout");
        check(c.threadBlockSize.x == 16 && c.threadBlockSize.y == 4 && c.dynamicTileWidth == 128 && c.dynamicTileHeight == 64, "AxB values");
        check(c.cameraLookFromVec.x == 1.f && c.cameraLookFromVec.y == 2.5f && c.cameraLookFromVec.z == -3.f && c.cameraFrontVec.z == -1.f, "x,y,z values");
        check(c.algorithmType == DYNAMIC && l && l->monitor && l->lookFromSet && l->frontSet, "scheduler name, monitor, camera-set markers");
    }
    for (auto name : {"fsfl", "dsfl", "dsdl"}) {
        RendererConfig c = parse({"--scheduler", name});
        check((std::string(name) == "fsfl" && c.algorithmType == FSFL) || (std::string(name) == "dsfl" && c.algorithmType == DSFL) ||
                  (std::string(name) == "dsdl" && c.algorithmType == DSDL),
              "legacy scheduler names");
    }
    check(throws([] { parse({"--bogus", "1"}); }, "--bogus"), "unknown flag is an error naming the flag");
    check(throws([] { parse({"--width"}); }, "--width"), "missing value is an error naming the flag");
    check(throws([] { parse({"--scheduler", "magic"}); }, "magic"), "unknown scheduler is an error");
    check(throws([] { parse({"--block", "16"}); }, "AxB"), "malformed AxB is an error");
    printf("%s\n", failures ? "ARGUMENT_LOADER_TEST_FAILED" : "ARGUMENT_LOADER_TEST_OK");
    return failures ? 1 : 0;
}

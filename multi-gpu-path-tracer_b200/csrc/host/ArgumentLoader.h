// ArgumentLoader.h — the reference's two positionals (src/ArgumentLoader.h:10-13: argv[1] = jobId,
// argv[2] = modelPath, same defaults) plus optional flags, because every BASELINE config needs
// parameters the reference can only receive over its websocket:
//   --width N --height N --spp N --depth N --gpus N --streams N --block BXxBY --scheduler fsfl|dsfl|dsdl|dynamic|lpt
//   --tile WxH --out file.ppm --frames N --vfov F --hfov F --lookfrom x,y,z --front x,y,z --show-tasks 0|1
//   --monitor 0|1 (the reference's monitor thread: NVML figures + RENDER_STATS# messages every 500 ms)
#pragma once

#include "RendererConfig.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

class ArgumentLoader {
public:
    ArgumentLoader(int argc, char **argv) : argc(argc), argv(argv) {}

    void loadArguments(RendererConfig &config) {
        int positional = 0;
        config.jobId = "0";
        config.modelPath = "models/cornel/cornell_box.gltf";
        for (int i = 1; i < argc; i++) {
            std::string a = argv[i];
            if (a.rfind("--", 0) != 0) {
                if (positional == 0 && argv[i][0] != '\0') config.jobId = a;
                else if (positional == 1 && argv[i][0] != '\0') config.modelPath = a;
                positional++;
                continue;
            }
            auto next = [&]() -> std::string {
                if (i + 1 >= argc) throw std::runtime_error("missing value for " + a);
                return argv[++i];
            };
            if (a == "--width") config.resolution.width = (unsigned)std::stoul(next());
            else if (a == "--height") config.resolution.height = (unsigned)std::stoul(next());
            else if (a == "--spp") config.samplesPerPixel = (unsigned)std::stoul(next());
            else if (a == "--depth") config.recursionDepth = (unsigned)std::stoul(next());
            else if (a == "--gpus") config.gpuNumber = (unsigned)std::stoul(next());
            else if (a == "--streams") config.streamsPerGpu = (unsigned)std::stoul(next());
            else if (a == "--frames") config.framesToRender = (unsigned)std::stoul(next());
            else if (a == "--out") config.outputPath = next();
            else if (a == "--vfov") config.vfov = std::stof(next());
            else if (a == "--hfov") config.hfov = std::stof(next());
            else if (a == "--show-tasks") config.showTasks = std::stoi(next()) != 0;
            else if (a == "--monitor") monitor = std::stoi(next()) != 0;
            else if (a == "--max-tasks-in-row") config.maxTasksInRow = (unsigned)std::stoul(next());
            else if (a == "--block") { unsigned x = 8, y = 8; parse2(next(), 'x', x, y); config.threadBlockSize = dim3(x, y); }
            else if (a == "--tile") parse2(next(), 'x', config.dynamicTileWidth, config.dynamicTileHeight);
            else if (a == "--lookfrom") { config.cameraLookFromVec = parse3(next()); lookFromSet = true; }
            else if (a == "--front") { config.cameraFrontVec = parse3(next()); frontSet = true; }
            else if (a == "--scheduler") {
                std::string s = next();
                if (s == "fsfl") config.algorithmType = FSFL;
                else if (s == "dsfl") config.algorithmType = DSFL;
                else if (s == "dsdl") config.algorithmType = DSDL;
                else if (s == "dynamic") config.algorithmType = DYNAMIC;
                else if (s == "lpt") config.algorithmType = LPT;
                else throw std::runtime_error("unknown scheduler " + s);
            } else throw std::runtime_error("unknown argument " + a);
        }
    }
    bool lookFromSet = false, frontSet = false, monitor = false;

private:
    static void parse2(const std::string &s, char sep, unsigned &a, unsigned &b) {
        size_t p = s.find(sep);
        if (p == std::string::npos) throw std::runtime_error("expected AxB, got " + s);
        a = (unsigned)std::stoul(s.substr(0, p));
        b = (unsigned)std::stoul(s.substr(p + 1));
    }
    static float3 parse3(const std::string &s) {
        float v[3] = {0, 0, 0};
        size_t pos = 0;
        for (int k = 0; k < 3; k++) {
            size_t c = s.find(',', pos);
            v[k] = std::stof(s.substr(pos, c == std::string::npos ? std::string::npos : c - pos));
            if (c == std::string::npos) break;
            pos = c + 1;
        }
        return make_float3(v[0], v[1], v[2]);
    }
    int argc;
    char **argv;
};

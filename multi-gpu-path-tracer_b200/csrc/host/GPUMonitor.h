// GPUMonitor.h — the reference's monitor (src/Profiling/GPUMonitor.{h,cpp}: GPUMonitor + MonitorThread) on the new host
// side: same public methods, same RENDER_STATS# message ("unit|name|value|" triples per GPU, GPUMonitor.cpp:107-132), same
// 500 ms cadence (GPUMonitor.cpp:139-147), sent through Renderer::send.
//
// Differences, on purpose:
//   * NVML is opened at run time (dlopen of libnvidia-ml.so.1) instead of being linked, so the host binaries start on a
//     box without the driver library; memory / utilisation figures are then reported as 0 and available() says so;
//   * the per-GPU render times feed a fixed array of 8 vectors in the reference (GPUMonitor.h:57) — sized here by the
//     device count — and every accessor takes the mutex the reference lacks (the monitor thread reads while the render
//     loop writes);
//   * free/total bytes come from NVML only: the reference's extra cudaMemGetInfo (GPUMonitor.cpp:38) creates a CUDA
//     context on whichever device is current in the monitor thread.
#pragma once

#include "Renderer.h"

#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

class GPUMonitor {
public:
    struct DeviceInfo {
        void *handle = nullptr;
        char name[96] = {0};
        unsigned long long memTotal = 0, memFree = 0, memUsed = 0;
        unsigned int utilGpu = 0, utilMem = 0;
    };

    GPUMonitor() {
        last_fps_update_ = std::chrono::high_resolution_clock::now();
        lib_ = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!lib_) lib_ = dlopen("libnvidia-ml.so", RTLD_NOW | RTLD_LOCAL);
        if (!lib_) return;
        init_ = (int (*)())sym("nvmlInit_v2");
        shutdown_ = (int (*)())sym("nvmlShutdown");
        count_ = (int (*)(unsigned int *))sym("nvmlDeviceGetCount_v2");
        byIndex_ = (int (*)(unsigned int, void **))sym("nvmlDeviceGetHandleByIndex_v2");
        name_ = (int (*)(void *, char *, unsigned int))sym("nvmlDeviceGetName");
        mem_ = (int (*)(void *, NvmlMemory *))sym("nvmlDeviceGetMemoryInfo");
        util_ = (int (*)(void *, NvmlUtilization *))sym("nvmlDeviceGetUtilizationRates");
        if (!init_ || !count_ || !byIndex_ || init_() != 0) return;
        unsigned int n = 0;
        if (count_(&n) != 0) return;
        device_infos_.resize(n);
        for (unsigned int i = 0; i < n; i++) {
            if (byIndex_(i, &device_infos_[i].handle) != 0) device_infos_[i].handle = nullptr;
            if (device_infos_[i].handle && name_) name_(device_infos_[i].handle, device_infos_[i].name, sizeof device_infos_[i].name);
        }
        timesOfRendering.resize(n);
        ok_ = true;
    }
    ~GPUMonitor() {
        if (ok_ && shutdown_) shutdown_();
        if (lib_) dlclose(lib_);
    }
    GPUMonitor(const GPUMonitor &) = delete;
    GPUMonitor &operator=(const GPUMonitor &) = delete;

    bool available() const { return ok_; }
    unsigned int deviceCount() const { return (unsigned int)device_infos_.size(); }

    void queryStats() {
        std::lock_guard<std::mutex> lock(mu_);
        for (auto &d : device_infos_) {
            if (!d.handle) continue;
            NvmlMemory m{};
            if (mem_ && mem_(d.handle, &m) == 0) {
                d.memTotal = m.total;
                d.memFree = m.free;
                d.memUsed = m.used;
            }
            NvmlUtilization u{};
            if (util_ && util_(d.handle, &u) == 0) {
                d.utilGpu = u.gpu;
                d.utilMem = u.memory;
            }
        }
    }

    void logLatestStats() {
        std::lock_guard<std::mutex> lock(mu_);
        for (size_t i = 0; i < device_infos_.size(); i++) {
            const DeviceInfo &d = device_infos_[i];
            printf("ID: %zu | Name: %s | Mem Total: %llu MB | Mem Free: %llu MB | GPU Util: %u | Mem Util: %u | TOR: %f | Imbalance: %f\n", i, d.name,
                   d.memTotal / 1000000ull, d.memFree / 1000000ull, d.utilGpu, d.utilMem, avg(timesOfRendering[i]), avg(loadImbalances));
        }
        printf("-------------------------------------------------------------------------\n");
        fflush(stdout);
    }

    // "unit|name|value|" triples, the format the reference's front-end parses (GPUMonitor.cpp:118-128); the per-interval
    // accumulators are cleared, as there
    std::string getLatestStats() {
        std::lock_guard<std::mutex> lock(mu_);
        auto now = std::chrono::high_resolution_clock::now();
        auto seconds = std::chrono::duration_cast<std::chrono::seconds>(now - last_fps_update_).count();
        if (seconds > 0) {
            fps_ = (int)(frame_count_ / seconds);
            frame_count_ = 0;
            last_fps_update_ = now;
            average_fps_ = (average_fps_ + fps_) / 2;
        }
        std::ostringstream stats;
        const float imbalance = avg(loadImbalances);
        for (size_t i = 0; i < device_infos_.size(); i++) {
            const DeviceInfo &d = device_infos_[i];
            stats << "FPS|FPS|" << fps_ << "|";
            stats << "FPS|Average FPS|" << average_fps_ << "|";
            stats << "MB|Mem Total GPU " << i << "|" << d.memTotal / 1000000ull << "|";
            stats << "MB|Mem Free GPU " << i << "|" << d.memFree / 1000000ull << "|";
            stats << "%|GPU Util GPU " << i << "|" << d.utilGpu << "|";
            stats << "%|Mem Util GPU " << i << "|" << d.utilMem << "|";
            stats << "ms|TOR " << i << "|" << avg(timesOfRendering[i]) << "|";
            stats << "IM|Imbalance " << i << "|" << imbalance << "|";
            timesOfRendering[i].clear();
        }
        loadImbalances.clear();
        return stats.str();
    }

    void updateFps() {
        std::lock_guard<std::mutex> lock(mu_);
        frame_count_++;
    }
    void updateTimeOfRendering(int gpuIdx, float ms) {
        std::lock_guard<std::mutex> lock(mu_);
        if (gpuIdx < 0) return;
        if ((size_t)gpuIdx >= timesOfRendering.size()) timesOfRendering.resize((size_t)gpuIdx + 1);
        timesOfRendering[(size_t)gpuIdx].push_back(ms);
    }
    float avgTimeOfRendering(int gpuIdx) {
        std::lock_guard<std::mutex> lock(mu_);
        return gpuIdx >= 0 && (size_t)gpuIdx < timesOfRendering.size() ? avg(timesOfRendering[(size_t)gpuIdx]) : 0.f;
    }
    void updateImbalance(float im) {
        std::lock_guard<std::mutex> lock(mu_);
        loadImbalances.push_back(im);
    }
    float avgImbalance() {
        std::lock_guard<std::mutex> lock(mu_);
        return avg(loadImbalances);
    }
    const std::vector<DeviceInfo> &devices() const { return device_infos_; }

private:
    struct NvmlMemory {
        unsigned long long total, free, used;
    };
    struct NvmlUtilization {
        unsigned int gpu, memory;
    };
    void *sym(const char *n) { return lib_ ? dlsym(lib_, n) : nullptr; }
    static float avg(const std::vector<float> &v) {
        if (v.empty()) return 0.f;
        float s = 0.f;
        for (float x : v) s += x;
        return s / (float)v.size();
    }

    void *lib_ = nullptr;
    bool ok_ = false;
    int (*init_)() = nullptr;
    int (*shutdown_)() = nullptr;
    int (*count_)(unsigned int *) = nullptr;
    int (*byIndex_)(unsigned int, void **) = nullptr;
    int (*name_)(void *, char *, unsigned int) = nullptr;
    int (*mem_)(void *, NvmlMemory *) = nullptr;
    int (*util_)(void *, NvmlUtilization *) = nullptr;

    std::mutex mu_;
    std::vector<DeviceInfo> device_infos_;
    int average_fps_ = 0, fps_ = 0;
    long long frame_count_ = 0;
    std::chrono::time_point<std::chrono::high_resolution_clock> last_fps_update_;
    std::vector<std::vector<float>> timesOfRendering;
    std::vector<float> loadImbalances;
};

// Functor for a std::thread, as in the reference's main (src/main.cu:76-93): query, log, send "RENDER_STATS#...", sleep 500 ms.
class MonitorThread {
public:
    explicit MonitorThread(Renderer &renderer, bool logToStdout = true, int periodMs = 500) : renderer(renderer), log_{logToStdout}, periodMs_{periodMs} {}

    void operator()() {
        while (!shouldTerminate) {
            monitor_.queryStats();
            if (log_) monitor_.logLatestStats();
            renderer.send("RENDER_STATS#" + monitor_.getLatestStats());
            for (int waited = 0; waited < periodMs_ && !shouldTerminate; waited += 10) std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
    }
    void safeTerminate() { shouldTerminate = true; }
    void updateFps() { monitor_.updateFps(); }
    void updateTimeOfRendering(int gpuIdx, float ms) { monitor_.updateTimeOfRendering(gpuIdx, ms); }
    void updateImbalance(float im) { monitor_.updateImbalance(im); }
    GPUMonitor &monitor() { return monitor_; }

private:
    std::atomic_bool shouldTerminate{false};
    Renderer &renderer;
    GPUMonitor monitor_;
    bool log_;
    int periodMs_;
};

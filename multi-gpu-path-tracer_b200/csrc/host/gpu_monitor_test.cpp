// gpu_monitor_test.cpp — GPUMonitor / MonitorThread without a GPU: accumulators, the RENDER_STATS# wire format
// (reference src/Profiling/GPUMonitor.cpp:107-132,139-147) and clean shutdown of the monitor thread.  With NVML present
// (GPU box) it additionally expects one block of figures per device.
#include "GPUMonitor.h"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct CaptureRenderer : Renderer {
    std::mutex mu;
    std::vector<std::string> messages;
    void renderFrame() override {}
    bool shouldStopRendering() override { return false; }
    void send(const std::string &data) override {
        std::lock_guard<std::mutex> lock(mu);
        messages.push_back(data);
    }
};

static int failures = 0;
static void check(bool ok, const char *what) {
    printf("%s %s\n", ok ? "ok  " : "FAIL", what);
    if (!ok) failures++;
}

int main() {
    GPUMonitor m;
    printf("nvml available: %d, devices: %u\n", (int)m.available(), m.deviceCount());
    check(m.avgTimeOfRendering(0) == 0.f && m.avgImbalance() == 0.f, "empty accumulators average to 0 (GPUMonitor.cpp:79-81)");
    m.updateTimeOfRendering(0, 10.f);
    m.updateTimeOfRendering(0, 20.f);
    m.updateTimeOfRendering(3, 7.f);
    m.updateImbalance(1.5f);
    m.updateImbalance(2.5f);
    check(m.avgTimeOfRendering(0) == 15.f && m.avgTimeOfRendering(3) == 7.f && m.avgTimeOfRendering(1) == 0.f, "time of rendering is averaged per GPU");
    check(m.avgImbalance() == 2.f, "imbalance is averaged");
    m.queryStats();
    std::string s = m.getLatestStats();
    if (m.available() && m.deviceCount() > 0) {
        check(s.find("FPS|FPS|") == 0, "message starts with the FPS triple");
        check(s.find("MB|Mem Total GPU 0|") != std::string::npos && s.find("%|GPU Util GPU 0|") != std::string::npos, "memory and utilisation triples for GPU 0");
        check(s.find("ms|TOR 0|15|") != std::string::npos && s.find("IM|Imbalance 0|2|") != std::string::npos, "TOR and imbalance of the interval");
        check(m.devices()[0].memTotal > 0 && m.devices()[0].name[0] != 0, "NVML reports a name and a memory size");
    } else {
        check(s.empty(), "no devices: empty statistics (one block per device, as in the reference)");
    }
    check(m.avgImbalance() == 0.f, "getLatestStats clears the interval (GPUMonitor.cpp:129-130)");

    CaptureRenderer r;
    MonitorThread mt(r, /*logToStdout=*/false, /*periodMs=*/20);
    std::thread t(std::ref(mt));
    for (int i = 0; i < 5; i++) {
        mt.updateFps();
        mt.updateTimeOfRendering(0, 1.f + i);
        mt.updateImbalance(1.f);
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
    mt.safeTerminate();
    t.join();
    {
        std::lock_guard<std::mutex> lock(r.mu);
        check(r.messages.size() >= 2, "the monitor thread sends a message per period");
        bool prefix = true;
        for (auto &msg : r.messages) prefix = prefix && msg.rfind("RENDER_STATS#", 0) == 0;
        check(prefix, "every message carries the RENDER_STATS# prefix");
    }
    printf("%s\n", failures ? "GPU_MONITOR_TEST_FAILED" : "GPU_MONITOR_TEST_OK");
    return failures ? 1 : 0;
}

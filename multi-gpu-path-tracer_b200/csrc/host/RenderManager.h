// RenderManager.h — frame orchestration across the GPUs of one box (one process, one worker thread per
// (GPU, stream)), keeping the reference's public surface (src/RenderManager.h:27-659: ctor, setup/reset, the
// deferred setters, renderFrame, markTasks, getFramebuffer/getCurrentFrame/getYUVFrame/getCurrentFrameWidth/
// Height, reloadScene/updatePrimitives) on top of the new core.
//
// What changed underneath:
//   * workers index their own task (the reference's StreamThread renders tasks_[deviceIdx], so with
//     streamsPerGpu > 1 several threads render the same rectangle, src/StreamThread.h:83; fixed here);
//   * FSFL / DSFL / DSDL keep their meaning (fixed cells / fixed layout with borders nudged by <= one thread
//     block per frame towards equal time / time-weighted recursive bisection) for API compatibility;
//   * DYNAMIC (new): the frame is cut into small tiles and every worker pulls the next tile index from one
//     std::atomic counter until the frame is exhausted — work stealing by construction;
//   * no managed framebuffer: tiles are gathered into the frame's master copy on GPU 0 by peer copies (NVLink) on the
//     rendering stream (DevicePathTracer::gatherTile) and downloaded to the pinned host frame once per frame.
#pragma once

#include "CameraConfig.h"
#include "DevicePathTracer.h"
#include "Framebuffer.h"
#include "GPUMonitor.h"
#include "HostScene.h"
#include "RendererConfig.h"
#include "TaskGenerator.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

class RenderManager {
public:
    struct FrameStats {
        double frame_ms = 0;          // wall clock of the last renderFrame (render + gather + download)
        double imbalance = 1;         // max worker time / mean worker time (the reference's metric, :433-447)
        std::vector<double> worker_ms;
        std::vector<int> worker_tiles;
    };

    RenderManager(RendererConfig &config, HostScene &hScene, CameraConfig &cameraConfig, SceneLoader &sceneLoader)
        : hScene_{hScene}, config_{config}, cameraConfig_{cameraConfig}, sceneLoader_{sceneLoader} {
        newConfig_ = config_;
        setup();
    }
    ~RenderManager() { reset(); }

    std::vector<std::vector<int>> getTaskLayout(unsigned int maxTasksInRow) {
        // reference :42-59: rows of at most maxTasksInRow cells
        std::vector<std::vector<int>> layout;
        int total = (int)(config_.gpuNumber * config_.streamsPerGpu), task = 0;
        while (task < total) {
            layout.push_back({});
            for (unsigned r = 0; r < std::max(1u, maxTasksInRow) && task < total; r++) layout.back().push_back(task++);
        }
        return layout;
    }

    void reset() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            shutdown_ = true;
        }
        cvStart_.notify_all();
        for (auto &w : workers_)
            if (w.thread.joinable()) w.thread.join();
        for (auto &w : workers_) {
            cudaSetDevice(w.device);
            for (int k = 0; k < kInFlight; k++) {
                if (w.streams[k]) cudaStreamDestroy(w.streams[k]);
                if (w.events[k]) cudaEventDestroy(w.events[k]);
            }
        }
        workers_.clear();
        devicePathTracers_.clear();
        shutdown_ = false;
    }

    void setup() {
        int available = 0;
        checkCudaErrors(cudaGetDeviceCount(&available));
        if ((int)config_.gpuNumber > available) config_.gpuNumber = (unsigned)std::max(1, available);
        if (config_.gpuNumber == 0) config_.gpuNumber = 1;
        if (config_.streamsPerGpu == 0) config_.streamsPerGpu = 1;
        framebuffer_ = std::make_shared<Framebuffer>(config_.resolution, 0);
        threadCount_ = (int)(config_.gpuNumber * config_.streamsPerGpu);
        for (unsigned i = 0; i < config_.gpuNumber; i++)
            devicePathTracers_.push_back(std::make_shared<DevicePathTracer>((int)i, config_.samplesPerPixel, config_.recursionDepth, config_.threadBlockSize,
                                                                           hScene_, framebuffer_, cameraConfig_));
        workers_ = std::vector<Worker>((size_t)threadCount_);
        for (int w = 0; w < threadCount_; w++) {
            Worker &wk = workers_[(size_t)w];
            wk.index = w;
            wk.device = w / (int)config_.streamsPerGpu;
            wk.seenFrame = frameId_;
            cudaSetDevice(wk.device);
            for (int k = 0; k < kInFlight; k++) {
                checkCudaErrors(cudaStreamCreateWithFlags(&wk.streams[k], cudaStreamNonBlocking));
                checkCudaErrors(cudaEventCreateWithFlags(&wk.events[k], cudaEventDisableTiming));
            }
        }
        regenerateTasks();
        for (auto &wk : workers_) wk.thread = std::thread(&RenderManager::workerMain, this, &wk);
    }

    // ---- deferred parameter changes, applied at the next renderFrame (reference :114-248) ----
    void setKParameter(int val) { newConfig_.kParam = val; shouldUpdatePathTracerParams = true; }
    void setGpuNumber(int gpuNumber) {
        if (config_.algorithmType == SchedulingAlgorithmType::DSDL) {  // reference :188-195: powers of two only
            int p = 1;
            while (p * 2 <= gpuNumber) p *= 2;
            gpuNumber = p;
        }
        newConfig_.gpuNumber = (unsigned)gpuNumber;
        shouldUpdatePathTracerParams = true;
    }
    void setStreamsPerGpu(int streamsPerGpu) { newConfig_.streamsPerGpu = (unsigned)streamsPerGpu; shouldUpdatePathTracerParams = true; }
    void setGpuAndStreamNumber(int gpuNumber, int streamsPerGpu) { setGpuNumber(gpuNumber); setStreamsPerGpu(streamsPerGpu); }
    void setResolution(Resolution res) { newConfig_.resolution = res; shouldUpdatePathTracerParams = true; }
    void setSamplesPerPixel(unsigned int samples) { newConfig_.samplesPerPixel = samples; shouldUpdatePathTracerParams = true; }
    void setRecursionDepth(unsigned int depth) { newConfig_.recursionDepth = depth; shouldUpdatePathTracerParams = true; }
    void setThreadBlockSize(dim3 tbs) { newConfig_.threadBlockSize = tbs; shouldUpdatePathTracerParams = true; }
    void setSchedulingAlgorithm(SchedulingAlgorithmType alg) { newConfig_.algorithmType = alg; shouldUpdatePathTracerParams = true; }
    void setShowTasks(bool val) { newConfig_.showTasks = val; shouldUpdatePathTracerParams = true; }

    void updatePathTracingParamsIfNeeded() {
        if (!shouldUpdatePathTracerParams) return;
        shouldUpdatePathTracerParams = false;
        bool retask = false;
        if (config_.algorithmType != newConfig_.algorithmType) { config_.algorithmType = newConfig_.algorithmType; retask = true; }
        config_.showTasks = newConfig_.showTasks;
        config_.kParam = newConfig_.kParam;
        if (config_.gpuNumber != newConfig_.gpuNumber || config_.streamsPerGpu != newConfig_.streamsPerGpu) {
            reset();
            config_.gpuNumber = newConfig_.gpuNumber;
            config_.streamsPerGpu = newConfig_.streamsPerGpu;
            setup();
            newConfig_.gpuNumber = config_.gpuNumber;
        }
        if (config_.resolution.width != newConfig_.resolution.width || config_.resolution.height != newConfig_.resolution.height) {
            config_.resolution = newConfig_.resolution;
            framebuffer_->setResolution(config_.resolution);
            for (const auto &dpt : devicePathTracers_) dpt->setFramebuffer(framebuffer_);
            retask = true;
        }
        if (config_.samplesPerPixel != newConfig_.samplesPerPixel) {
            config_.samplesPerPixel = newConfig_.samplesPerPixel;
            for (const auto &dpt : devicePathTracers_) dpt->setSamplesPerPixel(config_.samplesPerPixel);
        }
        if (config_.recursionDepth != newConfig_.recursionDepth) {
            config_.recursionDepth = newConfig_.recursionDepth;
            for (const auto &dpt : devicePathTracers_) dpt->setRecursionDepth(config_.recursionDepth);
        }
        if (config_.threadBlockSize.x != newConfig_.threadBlockSize.x || config_.threadBlockSize.y != newConfig_.threadBlockSize.y) {
            config_.threadBlockSize = newConfig_.threadBlockSize;
            for (const auto &dpt : devicePathTracers_) dpt->setThreadBlockSize(config_.threadBlockSize);
        }
        if (retask) regenerateTasks();
    }

    void reloadWorldIfNeeded() {
        if (!shouldReloadWorld) return;
        shouldReloadWorld = false;
        for (const auto &dpt : devicePathTracers_) dpt->reloadWorld();
    }

    void renderFrame() {
        updatePathTracingParamsIfNeeded();
        reloadWorldIfNeeded();
        if (frameCount_ > 0) {
            if (config_.algorithmType == SchedulingAlgorithmType::DSFL) adjustTasksDSFL();
            else if (config_.algorithmType == SchedulingAlgorithmType::DSDL) adjustTasksDSDL();
        }
        auto t0 = std::chrono::high_resolution_clock::now();
        nextTile_.store(0);
        {
            std::lock_guard<std::mutex> lock(mu_);
            pending_ = threadCount_;
            frameId_++;
        }
        cvStart_.notify_all();  // path tracer works here...
        {
            std::unique_lock<std::mutex> lock(mu_);
            cvDone_.wait(lock, [this] { return pending_ == 0; });
        }
        cudaSetDevice(framebuffer_->getMasterDevice());
        framebuffer_->downloadAsync(nullptr);
        checkCudaErrors(cudaStreamSynchronize(nullptr));
        auto t1 = std::chrono::high_resolution_clock::now();
        frameCount_++;

        stats_.frame_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats_.worker_ms.clear();
        stats_.worker_tiles.clear();
        double sum = 0, mx = 0;
        for (auto &w : workers_) {
            stats_.worker_ms.push_back(w.ms);
            stats_.worker_tiles.push_back(w.tiles);
            sum += w.ms;
            mx = std::max(mx, w.ms);
        }
        stats_.imbalance = sum > 0 ? mx / (sum / (double)workers_.size()) : 1.0;
        if (config_.showTasks && config_.algorithmType != SchedulingAlgorithmType::DYNAMIC) markTasks();
    }

    const FrameStats &lastFrameStats() const { return stats_; }

    // reference :433-447: feed the monitor with the last frame's render time per GPU (the slowest of its workers) and the
    // load imbalance (max / mean over the workers)
    void updateMetrics(MonitorThread &monitorThreadObj) {
        std::vector<double> perGpu(config_.gpuNumber, 0.0);
        for (auto &w : workers_)
            if (w.device >= 0 && (size_t)w.device < perGpu.size()) perGpu[(size_t)w.device] = std::max(perGpu[(size_t)w.device], w.ms);
        for (size_t g = 0; g < perGpu.size(); g++) monitorThreadObj.updateTimeOfRendering((int)g, (float)perGpu[g]);
        monitorThreadObj.updateImbalance((float)stats_.imbalance);
    }
    const std::vector<RenderTask> &tasks() const { return renderTasks_; }
    std::vector<std::shared_ptr<DevicePathTracer>> &tracers() { return devicePathTracers_; }

    // black grid over the task borders, on the host copy of the frame (reference :449-507)
    void markTasks() {
        const int W = (int)framebuffer_->getResolution().width, H = (int)framebuffer_->getResolution().height;
        const int boldness = H / 300;
        auto hline = [&](int x0, int x1, int row) {
            for (int x = std::max(0, x0); x < std::min(W, x1); x++)
                for (int r = row; r <= row + boldness && r < H; r++)
                    if (r >= 0) framebuffer_->updatePixel(r * W + x, 0, 0, 0);
        };
        auto vline = [&](int r0, int r1, int col) {
            for (int r = std::max(0, r0); r < std::min(H, r1); r++)
                for (int x = col; x <= col + boldness && x < W; x++)
                    if (x >= 0) framebuffer_->updatePixel(r * W + x, 0, 0, 0);
        };
        for (const RenderTask &t : renderTasks_) {
            if (t.offset_y != 0) hline(t.offset_x, t.offset_x + t.width, t.offset_y);
            hline(t.offset_x, t.offset_x + t.width, t.offset_y + t.height);
            if (t.offset_x != 0) vline(t.offset_y, t.offset_y + t.height, t.offset_x);
            vline(t.offset_y, t.offset_y + t.height, t.offset_x + t.width);
        }
    }

    std::shared_ptr<Framebuffer> &getFramebuffer() { return framebuffer_; }
    uint8_t *getCurrentFrame() { return framebuffer_->getRGBPtr(); }
    uint8_t *getYUVFrame() { return framebuffer_->getYUVPtr(); }
    unsigned int getCurrentFrameWidth() { return framebuffer_->getResolution().width; }
    unsigned int getCurrentFrameHeight() { return framebuffer_->getResolution().height; }

    void reloadScene() {
        std::string objPath = "../files/f" + config_.jobId + ".glb";  // reference :534-539
        hScene_ = sceneLoader_.load(objPath);
        shouldReloadWorld = true;
    }
    void updatePrimitives() { shouldReloadWorld = true; }

private:
    static constexpr int kInFlight = 4;  // DYNAMIC: tile launches a worker keeps in flight (one stream each)
    struct Worker {
        int index = 0, device = 0;
        cudaStream_t streams[kInFlight] = {nullptr, nullptr, nullptr, nullptr};
        cudaEvent_t events[kInFlight] = {nullptr, nullptr, nullptr, nullptr};
        std::thread thread;
        double ms = 0;
        int tiles = 0;
        uint64_t seenFrame = 0;
    };

    void regenerateTasks() {
        const int W = (int)config_.resolution.width, H = (int)config_.resolution.height;
        taskLayout_ = getTaskLayout(config_.maxTasksInRow);
        renderTasks_ = taskGen_.generateEqualTasks(threadCount_, taskLayout_, W, H);
        tiles_ = taskGen_.generateTiles((int)std::max(8u, config_.dynamicTileWidth), (int)std::max(4u, config_.dynamicTileHeight), W, H);
        frameCount_ = 0;
    }

    void workerMain(Worker *w) {
        cudaSetDevice(w->device);
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(mu_);
                cvStart_.wait(lock, [&] { return shutdown_ || frameId_ != w->seenFrame; });
                if (shutdown_) return;
                w->seenFrame = frameId_;
            }
            auto t0 = std::chrono::high_resolution_clock::now();
            DevicePathTracer &dpt = *devicePathTracers_[(size_t)w->device];
            w->tiles = 0;
            if (config_.algorithmType == SchedulingAlgorithmType::DYNAMIC) {
                const int n = (int)tiles_.size();
                for (;;) {
                    // a slot is reusable once its previous tile has finished: claims follow actual progress, so a GPU
                    // whose tiles are cheap simply claims more (work stealing through the shared counter)
                    const int slot = w->tiles % kInFlight;
                    if (w->tiles >= kInFlight) checkCudaErrors(cudaEventSynchronize(w->events[slot]));
                    int i = nextTile_.fetch_add(1);
                    if (i >= n) break;
                    RenderTask t = tiles_[(size_t)i];
                    dpt.renderTaskAsync(t, w->streams[slot]);
                    checkCudaErrors(cudaEventRecord(w->events[slot], w->streams[slot]));
                    w->tiles++;
                }
                for (int k = 0; k < kInFlight; k++) dpt.synchronizeStream(w->streams[k]);
            } else {
                RenderTask &t = renderTasks_[(size_t)w->index];
                dpt.renderTaskAsync(t, w->streams[0]);
                dpt.synchronizeStream(w->streams[0]);
                w->tiles = 1;
            }
            auto t1 = std::chrono::high_resolution_clock::now();
            w->ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
            if (config_.algorithmType != SchedulingAlgorithmType::DYNAMIC) renderTasks_[(size_t)w->index].time = (int)std::lround(w->ms);
            {
                std::lock_guard<std::mutex> lock(mu_);
                if (--pending_ == 0) cvDone_.notify_all();
            }
        }
    }

    // DSFL / DSDL: the arithmetic lives in TaskGenerator (pure functions of the previous frame's task times)
    void adjustTasksDSFL() {
        taskGen_.adjustTasksDSFL(renderTasks_, taskLayout_, (int)config_.resolution.width, (int)config_.resolution.height,
                                 (int)std::max(1u, config_.threadBlockSize.x), (int)std::max(1u, config_.threadBlockSize.y));
    }
    void adjustTasksDSDL() {
        renderTasks_ = taskGen_.bisectTasksDSDL(renderTasks_, threadCount_, (int)config_.resolution.width, (int)config_.resolution.height,
                                                (int)std::max(1u, config_.threadBlockSize.x), (int)std::max(1u, config_.threadBlockSize.y));
    }

    std::vector<std::shared_ptr<DevicePathTracer>> devicePathTracers_{};
    TaskGenerator taskGen_{};
    std::shared_ptr<Framebuffer> framebuffer_;
    HostScene &hScene_;
    std::vector<RenderTask> renderTasks_{};
    std::vector<RenderTask> tiles_{};
    std::vector<Worker> workers_{};
    RendererConfig &config_;
    RendererConfig newConfig_{};
    bool shouldUpdatePathTracerParams = false;
    bool shouldReloadWorld = false;
    CameraConfig &cameraConfig_;
    std::vector<std::vector<int>> taskLayout_;
    int threadCount_ = 0;
    SceneLoader &sceneLoader_;
    uint64_t frameCount_ = 0;

    std::mutex mu_;
    std::condition_variable cvStart_, cvDone_;
    uint64_t frameId_ = 0;
    int pending_ = 0;
    bool shutdown_ = false;
    std::atomic<int> nextTile_{0};
    FrameStats stats_;
};

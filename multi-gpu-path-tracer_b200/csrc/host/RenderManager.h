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
        lptRelease();
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
        if (config_.showTasks && config_.algorithmType != SchedulingAlgorithmType::DYNAMIC && config_.algorithmType != SchedulingAlgorithmType::LPT) markTasks();
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
        const uint32_t nBlocks = (uint32_t)((W + 7) / 8) * (uint32_t)((H + 3) / 4);
        if (lpt_.size() != config_.gpuNumber) lpt_.resize(config_.gpuNumber);
        if (lptHostCostsN_ < nBlocks) {
            if (lptHostCosts_) cudaFreeHost(lptHostCosts_);
            checkCudaErrors(cudaMallocHost((void **)&lptHostCosts_, sizeof(uint32_t) * nBlocks));
            lptHostCostsN_ = nBlocks;
        }
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
            if (config_.algorithmType == SchedulingAlgorithmType::LPT) {
                // one worker per GPU drives the frame; further streams of the same GPU have nothing to add to one persistent launch
                if (w->index % (int)config_.streamsPerGpu == 0) lptFrame(w, dpt);
            } else if (config_.algorithmType == SchedulingAlgorithmType::DYNAMIC) {
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
            if (config_.algorithmType != SchedulingAlgorithmType::DYNAMIC && config_.algorithmType != SchedulingAlgorithmType::LPT)
                renderTasks_[(size_t)w->index].time = (int)std::lround(w->ms);
            {
                std::lock_guard<std::mutex> lock(mu_);
                if (--pending_ == 0) cvDone_.notify_all();
            }
        }
    }

    // ---- LPT: the multi-GPU schedule of the bench (sched.py: render_frame_lpt), in-process ---------------------------------------------
    // A pixel is one sequential chain of spp samples, so a frame ends when its longest chain ends.  Every GPU traces a pilot pass over
    // 1/N of the 8x4 blocks (rays per block), the slices meet in pinned host memory, ONE thread orders the blocks by cost class and deals them
    // round-robin, every GPU renders its list most-expensive-first in one persistent launch and — if it is not the GPU that holds the
    // frame's master copy — pushes its blocks there with peer stores on the same stream.  Replaces the reference's per-frame rectangles
    // (src/RenderManager.h:264-408) for frames that are too short for feedback from the previous frame to help.
    struct LptGpu {
        uint32_t *dCosts = nullptr, *dBlocks = nullptr;
        uint32_t capacity = 0;
        std::vector<uint32_t> blocks;
    };
    void lptEnsure(int g, uint32_t n) {
        LptGpu &s = lpt_[(size_t)g];
        if (s.capacity >= n) return;
        if (s.dCosts) cudaFree(s.dCosts);
        if (s.dBlocks) cudaFree(s.dBlocks);
        checkCudaErrors(cudaMalloc((void **)&s.dCosts, sizeof(uint32_t) * n));
        checkCudaErrors(cudaMalloc((void **)&s.dBlocks, sizeof(uint32_t) * n));
        s.capacity = n;
    }
    void lptRelease() {
        for (size_t g = 0; g < lpt_.size(); g++) {
            cudaSetDevice((int)g);
            if (lpt_[g].dCosts) cudaFree(lpt_[g].dCosts);
            if (lpt_[g].dBlocks) cudaFree(lpt_[g].dBlocks);
        }
        lpt_.clear();
        if (lptHostCosts_) cudaFreeHost(lptHostCosts_);
        lptHostCosts_ = nullptr;
        lptHostCostsN_ = 0;
    }
    // all GPU-driving workers meet here; the last one to arrive runs `fn` before anyone leaves
    template <class F>
    void lptRendezvous(F fn) {
        std::unique_lock<std::mutex> lock(lptMu_);
        const uint64_t gen = lptGen_;
        if (++lptArrived_ == (int)config_.gpuNumber) {
            fn();
            lptArrived_ = 0;
            lptGen_++;
            lptCv_.notify_all();
        } else {
            lptCv_.wait(lock, [&] { return lptGen_ != gen; });
        }
    }
    void lptFrame(Worker *w, DevicePathTracer &dpt) {
        const int g = w->device, N = (int)config_.gpuNumber;
        const uint32_t W = config_.resolution.width, H = config_.resolution.height;
        const uint32_t bw = (W + 7) / 8, bh = (H + 3) / 4, n = bw * bh;
        cudaStream_t st = w->streams[0];
        lptEnsure(g, n);
        LptGpu &s = lpt_[(size_t)g];
        dpt.syncCamera();
        // 1. pilot pass over this GPU's slice of the block grid, slice -> pinned host
        const uint32_t per = (n + (uint32_t)N - 1) / (uint32_t)N, first = std::min(n, (uint32_t)g * per), cnt = std::min(per, n - first);
        checkPtcore(dpt.core(), ptcore_block_costs_range_async(dpt.core(), kLptPilotSpp, s.dCosts, first, cnt, st));
        if (cnt) checkCudaErrors(cudaMemcpyAsync(lptHostCosts_ + first, s.dCosts + first, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, st));
        checkCudaErrors(cudaStreamSynchronize(st));
        // 2. one thread orders the blocks (cost classes, Z-order within a class) and deals them
        lptRendezvous([&] {
            const std::vector<uint32_t> order = TaskGenerator::lptBlockOrder(lptHostCosts_, n, bw, TaskGenerator::lptLevels(N));
            for (int k = 0; k < N; k++) lpt_[(size_t)k].blocks.clear();
            for (uint32_t i = 0; i < n; i++) lpt_[(size_t)(i % (uint32_t)N)].blocks.push_back((order[i] % bw) | ((order[i] / bw) << 16));
        });
        // 3. this GPU's list, most expensive first, one persistent launch; then its pixels travel to the master copy
        const uint32_t mine = (uint32_t)s.blocks.size();
        if (mine) {
            checkCudaErrors(cudaMemcpyAsync(s.dBlocks, s.blocks.data(), sizeof(uint32_t) * mine, cudaMemcpyHostToDevice, st));
            checkPtcore(dpt.core(), ptcore_render_blocks_async(dpt.core(), s.dBlocks, mine, st));
            if (!dpt.rendersIntoMaster())
                checkPtcore(dpt.core(), ptcore_gather_blocks_async(dpt.core(), framebuffer_->getDeviceRGBPtr(), framebuffer_->getDeviceYUVPtr(), s.dBlocks, mine, st));
        }
        dpt.synchronizeStream(st);
        w->tiles = (int)mine;
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
    static constexpr uint32_t kLptPilotSpp = 4;
    std::vector<LptGpu> lpt_;
    uint32_t *lptHostCosts_ = nullptr;
    uint32_t lptHostCostsN_ = 0;
    std::mutex lptMu_;
    std::condition_variable lptCv_;
    int lptArrived_ = 0;
    uint64_t lptGen_ = 0;
};

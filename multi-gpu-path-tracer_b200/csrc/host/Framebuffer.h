// Framebuffer.h — RGB8 + I420 frame shared by every DevicePathTracer and the presenter.
//
// Same public surface as the reference's src/Framebuffer.h:6-89 (getRGBPtr/getYUVPtr/
// getResolution/getPixelCount/setResolution/updatePixel).  Storage differs: the reference
// lets all GPUs' kernels write one cudaMallocManaged buffer (page migration is its only
// inter-GPU data path); here the pointers returned to the host are PINNED host
// memory, each tracer renders into its own device buffer and finished tiles are gathered into
// the frame's device master copy on GPU 0 (peer copies over NVLink) and brought to the host once
// per frame by RenderManager::renderFrame.
#pragma once

#include "RendererConfig.h"
#include "cuda_utils.h"

#include <cstdint>

class Framebuffer {
public:
    explicit Framebuffer(Resolution res, int masterDevice = 0) : masterDevice_{masterDevice} {
        initializePointers(res);
        resolution_ = res;
    }
    Framebuffer(const Framebuffer &) = delete;
    Framebuffer &operator=(const Framebuffer &) = delete;

    void setResolution(Resolution res) {
        if (res.width == resolution_.width && res.height == resolution_.height) return;
        release();
        initializePointers(res);
        resolution_ = res;
    }

    void initializePointers(Resolution res) {
        const size_t totalPixels = (size_t)res.width * res.height;
        checkCudaErrors(cudaMallocHost((void **)&fb_rgb_ptr_, totalPixels * 3));
        checkCudaErrors(cudaMallocHost((void **)&fb_yuv_ptr_, totalPixels + 2 * (totalPixels / 4) + 2));
        int prev = 0;
        cudaGetDevice(&prev);
        checkCudaErrors(cudaSetDevice(masterDevice_));
        checkCudaErrors(cudaMalloc((void **)&dev_rgb_ptr_, totalPixels * 3));
        checkCudaErrors(cudaMalloc((void **)&dev_yuv_ptr_, totalPixels + 2 * (totalPixels / 4) + 2));
        checkCudaErrors(cudaMemset(dev_rgb_ptr_, 0, totalPixels * 3));
        checkCudaErrors(cudaMemset(dev_yuv_ptr_, 0, totalPixels + 2 * (totalPixels / 4) + 2));
        cudaSetDevice(prev);
    }

    Resolution getResolution() { return resolution_; }
    unsigned int getPixelCount() { return resolution_.width * resolution_.height; }
    uint8_t *getRGBPtr() { return fb_rgb_ptr_; }   // host (pinned)
    uint8_t *getYUVPtr() { return fb_yuv_ptr_; }   // host (pinned)
    uint8_t *getDeviceRGBPtr() { return dev_rgb_ptr_; }  // master copy on GPU `masterDevice`
    uint8_t *getDeviceYUVPtr() { return dev_yuv_ptr_; }
    int getMasterDevice() const { return masterDevice_; }

    // device master copy -> pinned host buffers (called once per frame)
    void downloadAsync(cudaStream_t stream) {
        const size_t totalPixels = (size_t)resolution_.width * resolution_.height;
        checkCudaErrors(cudaMemcpyAsync(fb_rgb_ptr_, dev_rgb_ptr_, totalPixels * 3, cudaMemcpyDeviceToHost, stream));
        checkCudaErrors(cudaMemcpyAsync(fb_yuv_ptr_, dev_yuv_ptr_, totalPixels + 2 * (totalPixels / 4), cudaMemcpyDeviceToHost, stream));
    }

    // reference src/Framebuffer.h:57-77 (host-side pixel poke used by the task-grid overlay)
    void updatePixel(int pixel_index, uint8_t r, uint8_t g, uint8_t b) {
        fb_rgb_ptr_[3 * pixel_index] = r;
        fb_rgb_ptr_[3 * pixel_index + 1] = g;
        fb_rgb_ptr_[3 * pixel_index + 2] = b;
        fb_yuv_ptr_[pixel_index] = ((66 * r + 129 * g + 25 * b + 128) >> 8) + 16;
        int blockRow = pixel_index / resolution_.width;
        int blockCol = pixel_index % resolution_.width;
        if (blockRow % 2 == 0 && blockCol % 2 == 0) {
            int totalPixels = resolution_.width * resolution_.height;
            int uvSize = totalPixels / 4;
            int uvIndex = (blockRow / 2) * (resolution_.width / 2) + (blockCol / 2);
            fb_yuv_ptr_[totalPixels + uvIndex] = ((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128;
            fb_yuv_ptr_[totalPixels + uvSize + uvIndex] = ((112 * r - 94 * g - 18 * b + 128) >> 8) + 128;
        }
    }

    ~Framebuffer() { release(); }

private:
    void release() {
        if (fb_rgb_ptr_) cudaFreeHost(fb_rgb_ptr_);
        if (fb_yuv_ptr_) cudaFreeHost(fb_yuv_ptr_);
        if (dev_rgb_ptr_) cudaFree(dev_rgb_ptr_);
        if (dev_yuv_ptr_) cudaFree(dev_yuv_ptr_);
        fb_rgb_ptr_ = fb_yuv_ptr_ = dev_rgb_ptr_ = dev_yuv_ptr_ = nullptr;
    }
    Resolution resolution_{0, 0};
    int masterDevice_ = 0;
    uint8_t *fb_rgb_ptr_ = nullptr;
    uint8_t *fb_yuv_ptr_ = nullptr;
    uint8_t *dev_rgb_ptr_ = nullptr;
    uint8_t *dev_yuv_ptr_ = nullptr;
};

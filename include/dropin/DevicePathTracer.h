// DevicePathTracer.h — THE drop-in: replaces src/DevicePathTracer.h of 3DevApps/multi-gpu-path-tracer, nothing else.
//
// Put this file in the reference's src/ in place of its own DevicePathTracer.h (392 lines: six __global__ kernels, the device
// scene and the class), add include/ (ptcore.h) to the include path and link libptcore.so.  Every other file of the reference
// — RenderManager.h, StreamThread.h, Framebuffer.h, HostScene.h, CameraConfig.h, RendererConfig.h, cuda_utils.h, main.cu —
// stays as it is: this header uses only what those define.  oracle/Makefile (target `dropin`) builds exactly that — the
// reference's unmodified RenderManager + StreamThread over this header — and tests/test_gpu_host_api.py compares its frames
// with the reference's own CUDA renderer byte for byte.
//
//   reference member (src/DevicePathTracer.h)              here
//   ------------------------------------------------------  -------------------------------------------------------------
//   struct RenderTask :19-25                                 same struct (RenderManager / StreamThread / TaskGenerator use it)
//   ctor :169-192 (reloadWorld, reloadCamera, setFramebuffer) ptcore_create + the same three calls in the same order
//   reloadWorld :312-340 (+ loadTextures / loadMaterials /    HostScene -> PtSceneDesc (sticky texture pointers of :269-279
//     loadTrianglesWithTextures :241-310, create_world,         kept) -> ptcore_upload_scene: host SAH BVH, one bulk upload
//     create_lights)
//   reloadCamera :230-239 + CameraConfig by value per launch  ptcore_set_camera, re-sent when the borrowed CameraConfig changed
//     :210
//   setFramebuffer :342-358 (curandState array, render_init)  ptcore_bind_framebuffer on the Framebuffer's own (managed) pointers
//   renderTaskAsync :194-214                                  ptcore_render_tile_async on the caller's stream
//   waitForRenderTask :216-220 / synchronizeStream :222-226   ptcore_wait / ptcore_sync
//   setSamplesPerPixel / setRecursionDepth /                  ptcore_set_params / ptcore_set_thread_block_size
//     setThreadBlockSize :360-370
//   dtor :372-377                                             ptcore_destroy
//   checkCudaErrors: print, cudaDeviceReset, exit(99)         the same on a non-zero return of any ptcore_* call
#pragma once

#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <float.h>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <vector>

#include "helper_math.h"
#include "cuda_utils.h"
#include "RendererConfig.h"
#include "CameraConfig.h"
#include "Framebuffer.h"
#include "HostScene.h"
#include "ptcore.h"

struct RenderTask {
    int width;
    int height;
    int offset_x;
    int offset_y;
    int time = 0;
};

#define checkPtcoreDropin(h, val) check_ptcore_dropin((h), (val), #val, __FILE__, __LINE__)
inline void check_ptcore_dropin(ptcore_t *h, int result, char const *const func, const char *const file, int const line) {
    if (result) {
        std::cerr << "CUDA ERROR = " << static_cast<unsigned int>(result) << " at " << file << ":" << line << " '" << func << "' (" << ptcore_last_error(h) << ") \n";
        cudaDeviceReset();
        exit(99);
    }
}

class DevicePathTracer {
public:
    DevicePathTracer(int device_idx, unsigned int samplesPerPixel, unsigned int recursionDepth, dim3 threadBlockSize, HostScene &hostScene,
                     std::shared_ptr<Framebuffer> framebuffer, CameraConfig &cameraConfig)
        : device_idx_{device_idx}, samplesPerPixel_{samplesPerPixel}, recursionDepth_{recursionDepth}, hostScene_{hostScene},
          threadBlockSize_{threadBlockSize}, framebuffer_{framebuffer}, cameraConfig_{cameraConfig} {
        cudaSetDevice(device_idx_);
        checkPtcoreDropin(nullptr, ptcore_create(device_idx_, &core_));
        checkPtcoreDropin(core_, ptcore_set_params(core_, samplesPerPixel_, recursionDepth_));
        checkPtcoreDropin(core_, ptcore_set_thread_block_size(core_, threadBlockSize_.x, threadBlockSize_.y));
        reloadWorld();
        reloadCamera();
        setFramebuffer(framebuffer_);
    }

    void renderTaskAsync(RenderTask &task, cudaStream_t stream) {
        if (task.width == 0) return;
        cudaSetDevice(device_idx_);
        {
            // the reference hands CameraConfig to the kernel by value at every launch (:210): an edit by another thread shows
            // in the next task.  Here the camera is re-sent when the borrowed object no longer matches what the core has.
            std::lock_guard<std::mutex> lock(mu_);
            if (!sameView(cameraConfig_, sent_)) sendCamera();
            checkPtcoreDropin(core_, ptcore_render_tile_async(core_, task.offset_x, task.offset_y, task.width, task.height, stream));
        }
    }

    void waitForRenderTask() {
        cudaSetDevice(device_idx_);
        checkPtcoreDropin(core_, ptcore_wait(core_));
    }

    void synchronizeStream(cudaStream_t stream) {
        cudaSetDevice(device_idx_);
        checkPtcoreDropin(core_, ptcore_sync(core_, stream));
    }

    // to be called when camera parameters change
    void reloadCamera() {
        std::lock_guard<std::mutex> lock(mu_);
        cudaSetDevice(device_idx_);
        sendCamera();
    }

    // To be called when scene triangles change
    void reloadWorld() {
        cudaSetDevice(device_idx_);
        const size_t n = hostScene_.triangles.size();
        std::vector<float> pos(n * 9), uv(n * 6);
        std::vector<int32_t> mat(n);
        for (size_t i = 0; i < n; i++) {
            const Triangle &t = hostScene_.triangles[i];
            const Vertex *v[3] = {&t.v0, &t.v1, &t.v2};
            for (int k = 0; k < 3; k++) {
                pos[i * 9 + 3 * k] = v[k]->position.x;
                pos[i * 9 + 3 * k + 1] = v[k]->position.y;
                pos[i * 9 + 3 * k + 2] = v[k]->position.z;
                uv[i * 6 + 2 * k] = v[k]->texCoords.x;
                uv[i * 6 + 2 * k + 1] = v[k]->texCoords.y;
            }
            mat[i] = t.materialIdx;
        }
        std::vector<PtMaterial> mats(hostScene_.materials.size());
        int stickyBase = -1, stickyEmis = -1;  // :269-279: a material without a texture inherits the last texture pointer seen
        for (size_t i = 0; i < mats.size(); i++) {
            const HostMaterial &m = hostScene_.materials[i];
            if (m.baseColorTextureIdx.has_value()) stickyBase = m.baseColorTextureIdx.value();
            if (m.emissiveTextureIdx.has_value()) stickyEmis = m.emissiveTextureIdx.value();
            PtMaterial &o = mats[i];
            o.type = PT_MAT_UNIVERSAL;  // the only material class the reference instantiates (:281-289)
            o.base[0] = m.baseColor.x; o.base[1] = m.baseColor.y; o.base[2] = m.baseColor.z;
            o.emis[0] = m.emissiveFactor.x; o.emis[1] = m.emissiveFactor.y; o.emis[2] = m.emissiveFactor.z;
            o.base_tex = stickyBase;
            o.emis_tex = stickyEmis;
            o.fuzz = 0.f;
            o.ior = 1.5f;
        }
        std::vector<PtTexture> tex(hostScene_.textures.size());
        for (size_t i = 0; i < tex.size(); i++) {
            tex[i].width = hostScene_.textures[i].width;
            tex[i].height = hostScene_.textures[i].height;
            tex[i].rgb = hostScene_.textures[i].data.empty() ? nullptr : &hostScene_.textures[i].data[0].x;
        }
        PtSceneDesc d{};
        d.n_tris = (int32_t)n; d.tri_pos = pos.data(); d.tri_uv = uv.data(); d.tri_mat = mat.data();
        d.n_mats = (int32_t)mats.size(); d.mats = mats.data();
        d.n_tex = (int32_t)tex.size(); d.tex = tex.data();
        checkPtcoreDropin(core_, ptcore_upload_scene(core_, &d));
    }

    void setFramebuffer(std::shared_ptr<Framebuffer> framebuffer) {
        framebuffer_ = framebuffer;
        cudaSetDevice(device_idx_);
        const Resolution res = framebuffer_->getResolution();
        // the reference's Framebuffer is cudaMallocManaged (src/Framebuffer.h:27-35) and every GPU's kernel stores into it (:103-119):
        // the same pointers are bound here, so the presenter keeps reading getRGBPtr() / getYUVPtr() exactly as before
        checkPtcoreDropin(core_, ptcore_bind_framebuffer(core_, framebuffer_->getRGBPtr(), framebuffer_->getYUVPtr(), res.width, res.height));
    }

    void setSamplesPerPixel(unsigned int samplesPerPixel) {
        samplesPerPixel_ = samplesPerPixel;
        checkPtcoreDropin(core_, ptcore_set_params(core_, samplesPerPixel_, recursionDepth_));
    }

    void setRecursionDepth(unsigned int recursionDepth) {
        recursionDepth_ = recursionDepth;
        checkPtcoreDropin(core_, ptcore_set_params(core_, samplesPerPixel_, recursionDepth_));
    }

    void setThreadBlockSize(dim3 threadBlockSize) {
        threadBlockSize_ = threadBlockSize;
        checkPtcoreDropin(core_, ptcore_set_thread_block_size(core_, threadBlockSize_.x, threadBlockSize_.y));
    }

    ~DevicePathTracer() {
        cudaSetDevice(device_idx_);
        if (core_) {
            ptcore_wait(core_);
            ptcore_destroy(core_);
        }
    }

private:
    static bool sameView(const CameraConfig &a, const CameraConfig &b) {
        return a.lookFrom.x == b.lookFrom.x && a.lookFrom.y == b.lookFrom.y && a.lookFrom.z == b.lookFrom.z && a.front.x == b.front.x && a.front.y == b.front.y &&
               a.front.z == b.front.z && a.vfov == b.vfov && a.hfov == b.hfov;
    }
    void sendCamera() {
        const CameraConfig snap = cameraConfig_;
        PtCamera c{{snap.lookFrom.x, snap.lookFrom.y, snap.lookFrom.z}, {snap.front.x, snap.front.y, snap.front.z}, snap.vfov, snap.hfov};
        checkPtcoreDropin(core_, ptcore_set_camera(core_, &c));
        sent_ = snap;
    }

    int device_idx_;
    unsigned int samplesPerPixel_;
    unsigned int recursionDepth_;
    HostScene &hostScene_;
    dim3 threadBlockSize_;
    std::shared_ptr<Framebuffer> framebuffer_;
    CameraConfig &cameraConfig_;
    CameraConfig sent_{make_float3(0.f, 0.f, 0.f), make_float3(0.f, 0.f, 0.f), -1.f, -1.f};
    ptcore_t *core_ = nullptr;
    std::mutex mu_;
};

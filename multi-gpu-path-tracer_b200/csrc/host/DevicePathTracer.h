// DevicePathTracer.h — the reference's device boundary class, kept signature for signature
// (reference src/DevicePathTracer.h:167-392) and implemented on the C ABI of include/ptcore.h.
//
//   reference                                         here
//   ------------------------------------------------  ------------------------------------------------
//   ctor: reloadWorld + reloadCamera + setFramebuffer  same order, through ptcore_*
//   reloadWorld: per-triangle device_vector push_back, one host SAH build + one bulk upload
//     create_world<<<1,1>>> (device BVH), lights        (ptcore_upload_scene)
//   reloadCamera: create_camera<<<1,1>>>                ptcore_set_camera (params evaluated once on device)
//   setFramebuffer: cudaMalloc 48 B/pixel curandState   ptcore_bind_framebuffer (RNG state lives in registers)
//     + render_init<<<>>>
//   renderTaskAsync: render<<<grid,block,0,stream>>>    ptcore_render_tile_async (+ tile gather to the frame's
//     into the managed framebuffer                       master copy on GPU 0, peer copies on the same stream)
//   errors: checkCudaErrors -> print, reset, exit(99)   same (checkPtcore)
//
// Kept quirk: loadMaterials' texture pointers are "sticky" across materials (reference :269-279):
// a material without a texture inherits the last one seen.  Reproduced when flattening the scene.
#pragma once

#include "../../../include/ptcore.h"
#include "CameraConfig.h"
#include "RenderTask.h"
#include "Framebuffer.h"
#include "HostScene.h"
#include "RendererConfig.h"
#include "cuda_utils.h"

#include <atomic>
#include <cstring>
#include <iostream>
#include <memory>
#include <mutex>
#include <vector>


class DevicePathTracer {
public:
    DevicePathTracer(int device_idx, unsigned int samplesPerPixel, unsigned int recursionDepth, dim3 threadBlockSize, HostScene &hostScene,
                     std::shared_ptr<Framebuffer> framebuffer, CameraConfig &cameraConfig)
        : device_idx_{device_idx}, threadBlockSize_{threadBlockSize}, hostScene_{hostScene}, samplesPerPixel_{samplesPerPixel},
          recursionDepth_{recursionDepth}, framebuffer_{framebuffer}, cameraConfig_{cameraConfig} {
        cudaSetDevice(device_idx_);
        int rc = ptcore_create(device_idx_, &core_);
        if (rc != 0) check_ptcore(nullptr, rc, "ptcore_create", __FILE__, __LINE__);
        checkPtcore(core_, ptcore_set_params(core_, samplesPerPixel_, recursionDepth_));
        checkPtcore(core_, ptcore_set_thread_block_size(core_, threadBlockSize_.x, threadBlockSize_.y));
        reloadWorld();
        reloadCamera();
        setFramebuffer(framebuffer_);
    }
    DevicePathTracer(const DevicePathTracer &) = delete;
    DevicePathTracer &operator=(const DevicePathTracer &) = delete;

    void renderTaskAsync(RenderTask &task, cudaStream_t stream) {
        if (task.width == 0) return;  // reference :195
        cudaSetDevice(device_idx_);
        {
            // the reference passes CameraConfig by value at every launch (:210), so a camera edited by
            // another thread takes effect at the next task; mirror that with a compare-and-upload
            // the launch snapshots the core's camera (by value into the kernel's parameters): keep the lock until it is issued, so that
            // a worker of another stream cannot re-upload the camera between this check and this launch
            std::lock_guard<std::mutex> lock(mu_);
            if (!sameCamera(cameraConfig_, uploadedCamera_)) uploadCamera();
            checkPtcore(core_, ptcore_render_tile_async(core_, task.offset_x, task.offset_y, task.width, task.height, stream));
        }
        if (!rendersIntoMaster_) gatherTile(task, stream);
    }

    void waitForRenderTask() {
        cudaSetDevice(device_idx_);
        checkPtcore(core_, ptcore_wait(core_));
    }

    void synchronizeStream(cudaStream_t stream) {
        cudaSetDevice(device_idx_);
        checkPtcore(core_, ptcore_sync(core_, stream));
    }

    // to be called when camera parameters change
    void reloadCamera() {
        std::lock_guard<std::mutex> lock(mu_);
        cudaSetDevice(device_idx_);
        uploadCamera();
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
                pos[i * 9 + 3 * k] = v[k]->position.x; pos[i * 9 + 3 * k + 1] = v[k]->position.y; pos[i * 9 + 3 * k + 2] = v[k]->position.z;
                uv[i * 6 + 2 * k] = v[k]->texCoords.x; uv[i * 6 + 2 * k + 1] = v[k]->texCoords.y;
            }
            mat[i] = t.materialIdx;
        }
        std::vector<float> sph(hostScene_.spheres.size() * 4);
        std::vector<int32_t> sphMat(hostScene_.spheres.size());
        for (size_t i = 0; i < hostScene_.spheres.size(); i++) {
            const HostSphere &s = hostScene_.spheres[i];
            sph[i * 4] = s.center.x; sph[i * 4 + 1] = s.center.y; sph[i * 4 + 2] = s.center.z; sph[i * 4 + 3] = s.radius;
            sphMat[i] = s.materialIdx;
        }
        std::vector<PtMaterial> mats(hostScene_.materials.size());
        int stickyBase = -1, stickyEmis = -1;  // reference :269-279
        for (size_t i = 0; i < mats.size(); i++) {
            const HostMaterial &m = hostScene_.materials[i];
            if (m.baseColorTextureIdx.has_value()) stickyBase = m.baseColorTextureIdx.value();
            if (m.emissiveTextureIdx.has_value()) stickyEmis = m.emissiveTextureIdx.value();
            PtMaterial &o = mats[i];
            o.type = (int32_t)m.type;
            o.base[0] = m.baseColor.x; o.base[1] = m.baseColor.y; o.base[2] = m.baseColor.z;
            o.emis[0] = m.emissiveFactor.x; o.emis[1] = m.emissiveFactor.y; o.emis[2] = m.emissiveFactor.z;
            o.base_tex = m.type == UNIVERSAL ? stickyBase : -1;
            o.emis_tex = m.type == UNIVERSAL ? stickyEmis : -1;
            o.fuzz = m.fuzz;
            o.ior = m.ior;
        }
        std::vector<PtTexture> tex(hostScene_.textures.size());
        for (size_t i = 0; i < tex.size(); i++) {
            tex[i].width = hostScene_.textures[i].width;
            tex[i].height = hostScene_.textures[i].height;
            tex[i].rgb = hostScene_.textures[i].data.empty() ? nullptr : &hostScene_.textures[i].data[0].x;
        }
        PtSceneDesc d{};
        d.n_tris = (int32_t)n; d.tri_pos = pos.data(); d.tri_uv = uv.data(); d.tri_mat = mat.data();
        d.n_spheres = (int32_t)sphMat.size(); d.sph = sph.data(); d.sph_mat = sphMat.data();
        d.n_mats = (int32_t)mats.size(); d.mats = mats.data();
        d.n_tex = (int32_t)tex.size(); d.tex = tex.data();
        checkPtcore(core_, ptcore_upload_scene(core_, &d));
    }

    void setFramebuffer(std::shared_ptr<Framebuffer> framebuffer) {
        framebuffer_ = framebuffer;
        cudaSetDevice(device_idx_);
        const Resolution res = framebuffer_->getResolution();
        const size_t px = (size_t)res.width * res.height;
        releasePrivate();
        rendersIntoMaster_ = device_idx_ == framebuffer_->getMasterDevice();
        if (rendersIntoMaster_) {
            checkPtcore(core_, ptcore_bind_framebuffer(core_, framebuffer_->getDeviceRGBPtr(), framebuffer_->getDeviceYUVPtr(), res.width, res.height));
        } else {
            checkCudaErrors(cudaMalloc((void **)&priv_rgb_, px * 3));
            checkCudaErrors(cudaMalloc((void **)&priv_yuv_, px + 2 * (px / 4) + 2));
            // NVLink P2P to the GPU that holds the frame's master copy.  Without it cudaMemcpy2DAsync(cudaMemcpyDefault) still works, but
            // staged through host memory: say so (once per tracer) instead of silently running the slow path.
            cudaError_t e = cudaDeviceEnablePeerAccess(framebuffer_->getMasterDevice(), 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                peerAccessFailures()++;
                std::cerr << "DevicePathTracer: GPU " << device_idx_ << " cannot enable peer access to GPU " << framebuffer_->getMasterDevice() << " (" << cudaGetErrorString(e)
                          << "): tile gather will be staged through host memory" << std::endl;
            }
            if (e != cudaSuccess) cudaGetLastError();
            checkPtcore(core_, ptcore_bind_framebuffer(core_, priv_rgb_, priv_yuv_, res.width, res.height));
        }
    }

    void setSamplesPerPixel(unsigned int samplesPerPixel) {
        samplesPerPixel_ = samplesPerPixel;
        checkPtcore(core_, ptcore_set_params(core_, samplesPerPixel_, recursionDepth_));
    }

    void setRecursionDepth(unsigned int recursionDepth) {
        recursionDepth_ = recursionDepth;
        checkPtcore(core_, ptcore_set_params(core_, samplesPerPixel_, recursionDepth_));
    }

    void setThreadBlockSize(dim3 threadBlockSize) {
        threadBlockSize_ = threadBlockSize;
        checkPtcore(core_, ptcore_set_thread_block_size(core_, threadBlockSize_.x, threadBlockSize_.y));
    }

    // additions
    static std::atomic<int> &peerAccessFailures() {
        static std::atomic<int> n{0};
        return n;
    }
    ptcore_t *core() { return core_; }
    bool rendersIntoMaster() const { return rendersIntoMaster_; }
    // makes sure the core has the camera the borrowed CameraConfig holds now (what renderTaskAsync does per task)
    void syncCamera() {
        std::lock_guard<std::mutex> lock(mu_);
        cudaSetDevice(device_idx_);
        if (!sameCamera(cameraConfig_, uploadedCamera_)) uploadCamera();
    }
    int deviceIndex() const { return device_idx_; }
    PtStats stats() {
        PtStats s{};
        checkPtcore(core_, ptcore_get_stats(core_, &s));
        return s;
    }

    ~DevicePathTracer() {
        cudaSetDevice(device_idx_);
        if (core_) {
            ptcore_wait(core_);
            ptcore_destroy(core_);
        }
        releasePrivate();
    }

private:
    static bool sameCamera(const CameraConfig &a, const CameraConfig &b) { return a.sameViewAs(b); }
    void uploadCamera() {
        CameraConfig snap = cameraConfig_;
        PtCamera c{{snap.lookFrom.x, snap.lookFrom.y, snap.lookFrom.z}, {snap.front.x, snap.front.y, snap.front.z}, snap.vfov, snap.hfov};
        checkPtcore(core_, ptcore_set_camera(core_, &c));
        uploadedCamera_ = snap;
    }
    void releasePrivate() {
        if (priv_rgb_) cudaFree(priv_rgb_);
        if (priv_yuv_) cudaFree(priv_yuv_);
        priv_rgb_ = priv_yuv_ = nullptr;
    }
    // copies exactly the bytes this tile owns (RGB rows, Y rows, the U/V samples of its even-row/even-column
    // pixels, reference :107-119) into the frame's master copy on GPU 0 — the only inter-GPU traffic of the path
    void gatherTile(const RenderTask &task, cudaStream_t stream) {
        const Resolution res = framebuffer_->getResolution();
        const int W = (int)res.width, H = (int)res.height;
        int x0 = std::max(task.offset_x, 0), x1 = std::min(task.offset_x + task.width, W);
        int y0 = std::max(task.offset_y, 0), y1 = std::min(task.offset_y + task.height, H);
        if (x1 <= x0 || y1 <= y0) return;
        const int r0 = H - y1, r1 = H - y0;  // buffer rows [r0, r1): row 0 is the top of the image
        uint8_t *mrgb = framebuffer_->getDeviceRGBPtr(), *myuv = framebuffer_->getDeviceYUVPtr();
        checkCudaErrors(cudaMemcpy2DAsync(mrgb + ((size_t)r0 * W + x0) * 3, (size_t)W * 3, priv_rgb_ + ((size_t)r0 * W + x0) * 3, (size_t)W * 3, (size_t)(x1 - x0) * 3,
                                          (size_t)(r1 - r0), cudaMemcpyDefault, stream));
        checkCudaErrors(cudaMemcpy2DAsync(myuv + (size_t)r0 * W + x0, (size_t)W, priv_yuv_ + (size_t)r0 * W + x0, (size_t)W, (size_t)(x1 - x0), (size_t)(r1 - r0),
                                          cudaMemcpyDefault, stream));
        const int c0 = (x0 + 1) / 2, c1 = (x1 + 1) / 2, k0 = (r0 + 1) / 2, k1 = (r1 + 1) / 2;  // chroma samples with 2c in [x0,x1), 2k in [r0,r1)
        if (c1 > c0 && k1 > k0) {
            const size_t total = (size_t)W * H, uvSize = total / 4, half = (size_t)(W / 2);
            for (size_t plane : {total, total + uvSize})
                checkCudaErrors(cudaMemcpy2DAsync(myuv + plane + (size_t)k0 * half + c0, half, priv_yuv_ + plane + (size_t)k0 * half + c0, half, (size_t)(c1 - c0),
                                                  (size_t)(k1 - k0), cudaMemcpyDefault, stream));
        }
    }

    int device_idx_;
    ptcore_t *core_ = nullptr;
    dim3 threadBlockSize_;
    HostScene &hostScene_;
    unsigned int samplesPerPixel_;
    unsigned int recursionDepth_;
    std::shared_ptr<Framebuffer> framebuffer_;
    CameraConfig &cameraConfig_;
    CameraConfig uploadedCamera_{make_float3(0, 0, 0), make_float3(0, 0, 0), -1.f, -1.f};
    bool rendersIntoMaster_ = true;
    uint8_t *priv_rgb_ = nullptr, *priv_yuv_ = nullptr;
    std::mutex mu_;
};

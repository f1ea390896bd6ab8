// ptcore.cu — C ABI implementation (include/ptcore.h): scene compile + upload, camera,
// framebuffer binding, kernel launches, statistics.  Host code here is the B200-native
// counterpart of DevicePathTracer's methods (reference src/DevicePathTracer.h:167-392).
#include "../../../include/ptcore.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "bvh_builder.h"
#include "pt_kernels.cuh"
#include "pt_pool.cuh"

using namespace ptc;

namespace {

constexpr int kCounterRing = 8192;  // one work counter per launch in flight; a launch reuses a slot only after 8192 later launches were issued on this handle

thread_local std::string g_create_error;

struct SceneBlob {
    std::vector<uint8_t> host;  // staged copy (pinned separately below)
    uint8_t *pinned = nullptr;
    uint8_t *dev = nullptr;
    size_t bytes = 0;
    size_t off_nodes4 = 0;
    size_t off_verts = 0;
    size_t off_nodes = 0, off_nodesq = 0, off_prims = 0, off_shade = 0, off_frames = 0, off_mats = 0, off_texdesc = 0, off_texels = 0, off_lights = 0;
    int32_t n_prims = 0, n_lights = 0, n_nodes = 0;
    bool has_spheres = false, has_rtow = false, has_nodes4 = false;
    QuantGrid grid{};
};

}  // namespace

struct ptcore {
    int device = 0;
    int sm_count = 0;
    std::string error;
    std::mutex mu;

    SceneBlob blob;
    bool have_scene = false;
    DevScene dscene{};

    CamParams cam{};
    bool have_cam = false;
    CamParams *d_cam = nullptr;

    uint32_t spp = 10, depth = 3;  // RendererConfig defaults, src/RendererConfig.h:22-23
    uint32_t bx = 8, by = 8;

    uint8_t *fb_rgb = nullptr, *fb_yuv = nullptr;
    uint32_t fb_w = 0, fb_h = 0;
    uint8_t *own_rgb = nullptr, *own_yuv = nullptr;  // internal framebuffer of ptcore_render_frame_host
    uint32_t own_w = 0, own_h = 0;

    uint32_t *d_work = nullptr;
    std::atomic<uint64_t> launch_seq{0};
    DevCounters *d_counters = nullptr;
    std::atomic<uint64_t> samples{0}, launches{0};

    int kernel = PT_KERNEL_PERSISTENT;
    bool count_tests = false;
    int leaf_max = 4;
    int blocks_per_sm = 0;
    int refill_at = 0;     // 0 = default (20: profiles/r02_knob_sweep.txt, +2.3 % over 24 with shared-memory nodes, +1.6 % without)
    int node_burst = 2;
    int min_blocks = 8;
    int bvh_width = 2;
    int node_format = PT_NODES_AUTO;
    int sah_isect_x100 = 120;
    int lanes_per_warp = 32;
    bool l2_persist_nodes = false, l2_limits_known = false;
    size_t l2_max_persist = 0, l2_max_window = 0, l2_set_aside = 0;
    int rng_mode = PT_RNG_STREAM;
    int rng_chunks = 16;
    bool sticky_textures = true;
    unsigned long long *retire_log = nullptr;
    uint32_t retire_log_warps = 0;
    int smem_nodes = 1;    // wavefront kernel: 1 (default) = one 1024-thread CTA per SM with the quantised nodes in shared memory when they fit, 0 = never
    int smem_nodes_max_bytes = 160 * 1024;
    int grid_ctas = 0;     // wavefront kernels: CTAs of the launch (0 = one per SM / SMs x occupancy)
    int cta_warps = 0;     // shared-memory-node kernel: warps per CTA (0 = auto by the launch's pixel count, smem_cta_warps)
    int pool_slots = 0;    // 0 = auto (pixels per warp of the launch, clamped to 32 .. kPoolSlots)
    int pool_idle_at = 8;
    int pool_period = 2;
    int pool_carveout = 28;
    uint32_t watchdog = 0;
    bool mempool_ready = false;
    bool trace_steps = false;
    uint32_t *ident_blocks = nullptr;
    uint32_t ident_blocks_n = 0, ident_bw = 0;

    PtStats build_stats{};
};

namespace {

int fail(ptcore *h, int code, const std::string &msg) {
    if (h) h->error = msg;
    else g_create_error = msg;
    return code;
}

#define PT_CUDA(h, call)                                                                                        \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) {                                                                                \
            return fail((h), (int)e_, std::string(#call) + ": " + cudaGetErrorName(e_) + " (" + cudaGetErrorString(e_) + ")"); \
        }                                                                                                       \
    } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline float as_float(int32_t i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}

// triangle.h:28 — evaluated on the host by the reference as well (the triangle constructor is host code)
float triangle_area_host(const float *p) {
    float e1[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]};
    float e2[3] = {p[6] - p[0], p[7] - p[1], p[8] - p[2]};
    float cx = e1[1] * e2[2] - e1[2] * e2[1];
    float cy = e1[2] * e2[0] - e1[0] * e2[2];
    float cz = e1[0] * e2[1] - e1[1] * e2[0];
    float d = cx * cx + cy * cy + cz * cz;
    return sqrtf(d) * 0.5f;
}

// Shade class of a primitive (pt_pool.cuh files finished rays under it so that one shade pass runs one branch of the integrator):
// 0 = the path ends on this primitive (an emitter without emissive texture, material.h:62-65 / :210-217), 1 = UniversalMaterial or
// lambertian bounce, 2 = metal, 3 = dielectric.  Scheduling only: shade() decides again from the material itself.
int32_t shade_class(const PtMaterial &m) {
    switch (m.type) {
        case PT_MAT_DIFFUSE_LIGHT: return 0;
        case PT_MAT_METAL: return 2;
        case PT_MAT_DIELECTRIC: return 3;
        case PT_MAT_UNIVERSAL:
            if (m.emis_tex < 0 && (m.emis[0] * 50 > 0.0001f || m.emis[1] * 50 > 0.0001f || m.emis[2] * 50 > 0.0001f)) return 0;
            return 1;
        default: return 1;
    }
}

// bounds of every primitive, padded: the slab test must never cull a hit the primitive test accepts
std::vector<PrimBounds> padded_bounds(const PtSceneDesc *sc) {
    const int64_t n_prims = (int64_t)sc->n_tris + sc->n_spheres;
    std::vector<PrimBounds> pb((size_t)n_prims);
    for (int32_t i = 0; i < sc->n_tris; i++) {
        const float *p = sc->tri_pos + (size_t)i * 9;
        for (int k = 0; k < 3; k++) {
            pb[(size_t)i].lo[k] = std::min(p[k], std::min(p[3 + k], p[6 + k]));
            pb[(size_t)i].hi[k] = std::max(p[k], std::max(p[3 + k], p[6 + k]));
        }
    }
    for (int32_t i = 0; i < sc->n_spheres; i++) {
        const float *s = sc->sph + (size_t)i * 4;
        float r = std::fabs(s[3]);
        for (int k = 0; k < 3; k++) {
            pb[(size_t)sc->n_tris + (size_t)i].lo[k] = s[k] - r;
            pb[(size_t)sc->n_tris + (size_t)i].hi[k] = s[k] + r;
        }
    }
    for (auto &b : pb)
        for (int k = 0; k < 3; k++) {
            float m = std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k]));
            float e = m * 1e-5f + 1e-6f;
            b.lo[k] -= e;
            b.hi[k] += e;
        }
    return pb;
}

// Quantised nodes pay off while their boxes stay close to the float boxes (cornell_duck 1.015: +1 %, the 2 M-triangle mesh
// 1.21: +8 %, 180 K triangles 1.06: +12 %); a scene with far outliers (the sphere field inside its sky sphere: 10.1) makes
// the scene-wide grid too coarse (-45 %).
inline bool use_quantised(const ptcore *h) {
    if (h->node_format == PT_NODES_QUANTISED) return true;
    if (h->node_format == PT_NODES_FULL) return false;
    return h->build_stats.quant_inflation > 0 && h->build_stats.quant_inflation <= 1.3;
}

// The pool kernel keeps one 128-byte record per pixel slot in global memory (pt_pool.cuh).  Launches of one handle may overlap
// (StreamThread with several streams per GPU, tiles in flight), so every launch gets its own records from the device's
// stream-ordered allocator; the pool keeps the memory, so after the first frames this is a pointer bump.
template <bool S, bool R, bool C>
cudaError_t launch_pool(ptcore *h, RenderParams rp, cudaStream_t stream) {
    const uint32_t total = rp.tiles.first_item[rp.tiles.n];
    void (*kernel)(RenderParams) = use_quantised(h) ? pt_pool_kernel<S, R, C, 2> : pt_pool_kernel<S, R, C, 0>;
    // 4 CTAs x (8 warps x 1792 B + 1 KB reserved) = 60 KB: ask for the 64 KB shared-memory configuration (28 % of 228 KB) so that
    // 192 KB stay L1; left to itself the driver picks a larger carve-out and the tree no longer fits L1
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, h->pool_carveout);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kPoolThreads, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    if (h->blocks_per_sm > 0) occ = std::min(occ, std::max(1, h->blocks_per_sm * kBlockThreads / kPoolThreads));
    uint32_t grid = (uint32_t)h->sm_count * (uint32_t)occ;
    const uint32_t needed = (total + kPoolThreads - 1) / kPoolThreads;
    if (grid > needed) grid = needed;
    const uint32_t n_warps = grid * (uint32_t)kPoolWarps;
    int pool = h->pool_slots > 0 ? h->pool_slots : (int)((total + n_warps - 1) / n_warps);
    pool = std::max(32, std::min(pool, kPoolSlots));
    if (!h->mempool_ready) {
        cudaMemPool_t mp;
        if (cudaDeviceGetDefaultMemPool(&mp, h->device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        h->mempool_ready = true;
    }
    void *slots = nullptr;
    e = cudaMallocAsync(&slots, (size_t)n_warps * (size_t)pool * kPoolRecQuads * sizeof(float4), stream);
    if (e != cudaSuccess) return e;
    rp.pool_slots = reinterpret_cast<float4 *>(slots);
    rp.pool_size = pool;
    rp.pool_idle_at = h->pool_idle_at;
    rp.pool_period = h->pool_period;
    rp.watchdog = h->watchdog;
    kernel<<<grid, kPoolThreads, 0, stream>>>(rp);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(slots, stream);
    return e != cudaSuccess ? e : e2;
}

// Warps per CTA (= per SM) of the shared-memory-node kernel.  A pixel is one sequential chain of spp samples and a chain advances 2.3x faster
// when its warp has a scheduler to itself than among 8 warps per scheduler, while the SM's throughput only needs ~5 warps per scheduler: a
// launch with fewer than ~3 pixels per lane (a 1/8 share of a 1080p frame, a 640x360 frame) is bound by its chains, not by throughput, and
// ends sooner on fewer, faster warps — 170 -> 160 ms for the 8-GPU share with 20 instead of 32 (profiles/r02_express_lane_ab.txt).
// Auto: pixels / (SMs x 32 lanes x 2.7), clamped to [12, 32]; full frames on one GPU and launches of under 64 spp keep 32.
static int smem_cta_warps(const ptcore *h, uint32_t pixels, uint32_t spp) {
    if (h->cta_warps) return h->cta_warps;
    if (spp < 64) return 32;  // short chains (the pilot pass): nothing to shorten
    const double w = (double)pixels / ((double)h->sm_count * 32.0 * 2.7);
    return (int)std::min(32.0, std::max(12.0, std::floor(w + 0.5)));
}

template <bool S, bool R, bool C>
cudaError_t launch_variant(ptcore *h, const RenderParams &rp, bool direct, cudaStream_t stream) {
    const uint32_t total = rp.tiles.first_item[rp.tiles.n];
    if (total == 0) return cudaSuccess;
    // the pool kernel walks the two-wide tree; depth 0 (no ray at all, camera.h:52,82) stays with the wavefront kernel
    if (h->kernel == PT_KERNEL_POOL && h->bvh_width != 4 && rp.depth > 0) return launch_pool<S, R, C>(h, rp, stream);
    if (!direct && h->kernel == PT_KERNEL_PERSISTENT && h->smem_nodes && h->bvh_width != 4 && use_quantised(h) &&
        (size_t)h->blob.n_nodes * 32 <= (size_t)h->smem_nodes_max_bytes) {
        const size_t bytes = (size_t)h->blob.n_nodes * 32 + (kWfSmemStack ? (size_t)(kSmemKernelThreads / 32) * kWfStackK * 128 : 0);  // nodes (+ the warps' short stacks)
        cudaError_t e = cudaFuncSetAttribute(pt_wavefront_smem_kernel<S, R, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        // PT_OPT_GRID_CTAS / PT_OPT_CTA_WARPS: an "express" launch of a few small CTAs (one warp per scheduler) for the longest chains, beside a
        // main launch that leaves those SMs free (sched.py: render_frame_lpt, express)
        const uint32_t threads = 32u * (uint32_t)smem_cta_warps(h, total, rp.spp);
        uint32_t grid = h->grid_ctas ? (uint32_t)h->grid_ctas : (uint32_t)h->sm_count;
        const uint32_t per_cta = threads / 32u * (uint32_t)h->lanes_per_warp;
        const uint32_t needed = (total + per_cta - 1) / per_cta;
        if (grid > needed) grid = needed;
        pt_wavefront_smem_kernel<S, R, C><<<grid, threads, bytes, stream>>>(rp, h->blob.n_nodes);
        return cudaGetLastError();
    }
    if (direct) {
        dim3 grid((total + kBlockThreads - 1) / kBlockThreads);
        pt_direct_kernel<S, R, C><<<grid, kBlockThreads, 0, stream>>>(rp);
    } else {
        int occ = 0;
        cudaError_t e = h->kernel == PT_KERNEL_LOCKSTEP ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pt_persistent_kernel<S, R, C>, kBlockThreads, 0)
                        : h->bvh_width == 4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pt_wavefront_kernel<S, R, C, 1>, kBlockThreads, 0)
                        : use_quantised(h) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pt_wavefront_kernel<S, R, C, 2>, kBlockThreads, 0)
                                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pt_wavefront_kernel<S, R, C, 0>, kBlockThreads, 0);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        if (h->blocks_per_sm > 0) occ = std::min(occ, h->blocks_per_sm);
        uint32_t grid = h->grid_ctas ? (uint32_t)h->grid_ctas : (uint32_t)h->sm_count * (uint32_t)occ;
        const uint32_t per_cta = (uint32_t)(kBlockThreads / 32 * (h->kernel == PT_KERNEL_PERSISTENT ? h->lanes_per_warp : 32));
        uint32_t needed = (total + per_cta - 1) / per_cta;
        if (grid > needed) grid = needed;
        if (h->kernel == PT_KERNEL_LOCKSTEP) pt_persistent_kernel<S, R, C><<<grid, kBlockThreads, 0, stream>>>(rp);
        else if (h->bvh_width == 4) pt_wavefront_kernel<S, R, C, 1><<<grid, kBlockThreads, 0, stream>>>(rp);
        else if (use_quantised(h)) pt_wavefront_kernel<S, R, C, 2><<<grid, kBlockThreads, 0, stream>>>(rp);
        else pt_wavefront_kernel<S, R, C, 0><<<grid, kBlockThreads, 0, stream>>>(rp);
    }
    return cudaGetLastError();
}

// PT_RNG_SAMPLE_KEYED launches: the wavefront kernel with KEYED = true (shared-memory nodes when they fit, else quantised or float
// nodes through L1); the test counters are not offered in this mode
template <bool S, bool R>
cudaError_t launch_keyed_variant(ptcore *h, const RenderParams &rp, cudaStream_t stream) {
    const uint32_t total = rp.tiles.first_item[rp.tiles.n] * rp.keyed_my_chunks;
    if (total == 0) return cudaSuccess;
    if (h->smem_nodes && use_quantised(h) && (size_t)h->blob.n_nodes * 32 <= (size_t)h->smem_nodes_max_bytes) {
        const size_t bytes = (size_t)h->blob.n_nodes * 32 + (kWfSmemStack ? (size_t)(kSmemKernelThreads / 32) * kWfStackK * 128 : 0);
        cudaError_t e = cudaFuncSetAttribute(pt_wavefront_smem_kernel<S, R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        uint32_t grid = (uint32_t)h->sm_count;
        const uint32_t needed = (total + kSmemKernelThreads - 1) / kSmemKernelThreads;
        if (grid > needed) grid = needed;
        pt_wavefront_smem_kernel<S, R, false, true><<<grid, kSmemKernelThreads, bytes, stream>>>(rp, h->blob.n_nodes);
        return cudaGetLastError();
    }
    int occ = 0;
    cudaError_t e = use_quantised(h) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pt_wavefront_kernel<S, R, false, 2, true>, kBlockThreads, 0)
                                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pt_wavefront_kernel<S, R, false, 0, true>, kBlockThreads, 0);
    if (e != cudaSuccess) return e;
    uint32_t grid = (uint32_t)h->sm_count * (uint32_t)std::max(1, occ);
    const uint32_t needed = (total + kBlockThreads - 1) / kBlockThreads;
    if (grid > needed) grid = needed;
    if (use_quantised(h)) pt_wavefront_kernel<S, R, false, 2, true><<<grid, kBlockThreads, 0, stream>>>(rp);
    else pt_wavefront_kernel<S, R, false, 0, true><<<grid, kBlockThreads, 0, stream>>>(rp);
    return cudaGetLastError();
}

cudaError_t launch_keyed(ptcore *h, const RenderParams &rp, cudaStream_t stream) {
    const bool S = h->blob.has_spheres, R = h->blob.has_rtow;
    if (!S && !R) return launch_keyed_variant<false, false>(h, rp, stream);
    if (!S && R) return launch_keyed_variant<false, true>(h, rp, stream);
    if (S && !R) return launch_keyed_variant<true, false>(h, rp, stream);
    return launch_keyed_variant<true, true>(h, rp, stream);
}

// PT_OPT_L2_PERSIST_NODES: scenes whose compiled data exceed L2 (the 2 M-triangle mesh: 396 MB) walk the tree at DRAM latency.  The
// quantised node array (36 MB there) is what 13 of a ray's 15 dependent loads touch: mark it persisting in L2 for the launches of this
// stream (access-policy window), so that the triangles, frames and texels that stream through L2 cannot evict it.
void apply_l2_policy(ptcore *h, cudaStream_t stream) {
    if (!h->l2_persist_nodes || !h->have_scene) return;
    if (!h->l2_limits_known) {
        int maxPersist = 0, maxWindow = 0;
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
        cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
        h->l2_max_persist = (size_t)std::max(0, maxPersist);
        h->l2_max_window = (size_t)std::max(0, maxWindow);
        h->l2_limits_known = true;
    }
    const bool quant = use_quantised(h);
    const uint8_t *base = h->blob.dev + (quant ? h->blob.off_nodesq : h->blob.off_nodes);
    size_t bytes = (size_t)h->blob.n_nodes * (quant ? 32 : 64);
    if (h->l2_max_persist == 0 || h->l2_max_window == 0 || bytes == 0) return;
    if (h->l2_set_aside != std::min(bytes, h->l2_max_persist)) {
        h->l2_set_aside = std::min(bytes, h->l2_max_persist);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_set_aside);
    }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.base_ptr = const_cast<uint8_t *>(base);
    attr.accessPolicyWindow.num_bytes = std::min(bytes, h->l2_max_window);
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)h->l2_set_aside / (double)attr.accessPolicyWindow.num_bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(stream ? stream : cudaStreamLegacy, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

cudaError_t launch(ptcore *h, const RenderParams &rp, cudaStream_t stream) {
    apply_l2_policy(h, stream);
    if ((h->kernel == PT_KERNEL_PERSISTENT || h->kernel == PT_KERNEL_POOL) && h->bvh_width == 4 && !h->blob.has_nodes4) return cudaErrorNotSupported;  // upload the scene with PT_OPT_BVH_WIDTH = 4 first
    const bool direct = h->kernel == PT_KERNEL_DIRECT;
    const bool S = h->blob.has_spheres, R = h->blob.has_rtow, C = h->count_tests;
    if (!S && !R && !C) return launch_variant<false, false, false>(h, rp, direct, stream);
    if (!S && !R && C) return launch_variant<false, false, true>(h, rp, direct, stream);
    if (!S && R && !C) return launch_variant<false, true, false>(h, rp, direct, stream);
    if (!S && R && C) return launch_variant<false, true, true>(h, rp, direct, stream);
    if (S && !R && !C) return launch_variant<true, false, false>(h, rp, direct, stream);
    if (S && !R && C) return launch_variant<true, false, true>(h, rp, direct, stream);
    if (S && R && !C) return launch_variant<true, true, false>(h, rp, direct, stream);
    return launch_variant<true, true, true>(h, rp, direct, stream);
}

void fill_dev_scene(ptcore *h) {
    SceneBlob &b = h->blob;
    DevScene &d = h->dscene;
    d.nodes = reinterpret_cast<const float4 *>(b.dev + b.off_nodes);
    d.nodesq = reinterpret_cast<const uint4 *>(b.dev + b.off_nodesq);
    d.nodes4 = reinterpret_cast<const float4 *>(b.dev + b.off_nodes4);
    d.grid_lo = make_float3(b.grid.lo[0], b.grid.lo[1], b.grid.lo[2]);
    d.grid_scale = make_float3(b.grid.scale[0], b.grid.scale[1], b.grid.scale[2]);
    d.pad2[0] = d.pad2[1] = 0.f;
    d.prims = reinterpret_cast<const float4 *>(b.dev + b.off_prims);
    d.primidx = reinterpret_cast<const uint4 *>(b.dev + b.off_prims);
    d.verts = reinterpret_cast<const float4 *>(b.dev + b.off_verts);
    d.shade = reinterpret_cast<const float4 *>(b.dev + b.off_shade);
    d.frames = reinterpret_cast<const float4 *>(b.dev + b.off_frames);
    d.mats = reinterpret_cast<const float4 *>(b.dev + b.off_mats);
    d.texs = reinterpret_cast<const TexDesc *>(b.dev + b.off_texdesc);
    d.texels = reinterpret_cast<const float4 *>(b.dev + b.off_texels);
    d.lights = reinterpret_cast<const float4 *>(b.dev + b.off_lights);
    d.n_lights = b.n_lights;
    d.n_prims = b.n_prims;
    d.light_pick_scale = (b.n_lights - 1) + 0.999999;  // hitable_list.h:24
    d.light_weight = b.n_lights > 0 ? 1.0f / b.n_lights : 0.f;
    d.pad = 0;
}

int render_blocks(ptcore *h, const uint32_t *blocks_dev, uint32_t n_blocks, uint32_t spp, uint32_t *block_cost, cudaStream_t stream) {
    if (!h->have_scene) return fail(h, PT_ERR_NO_SCENE, "no scene uploaded");
    if (!h->fb_rgb) return fail(h, PT_ERR_NO_FRAMEBUFFER, "no framebuffer bound");
    if (!h->have_cam) return fail(h, PT_ERR_INVALID_ARGUMENT, "no camera set");
    if (h->kernel != PT_KERNEL_PERSISTENT && h->kernel != PT_KERNEL_POOL) return fail(h, PT_ERR_UNSUPPORTED, "block lists need a persistent (pool or wavefront) kernel");
    if (n_blocks == 0) return PT_OK;
    if (n_blocks > 0x07ffffffu) return fail(h, PT_ERR_UNSUPPORTED, "too many blocks");
    PT_CUDA(h, cudaSetDevice(h->device));
    RenderParams rp;
    memset(&rp, 0, sizeof rp);
    rp.scene = h->dscene;
    rp.cam = h->cam;
    rp.width = h->fb_w;
    rp.height = h->fb_h;
    rp.spp = spp;
    rp.depth = h->depth;
    rp.refill_at = h->refill_at ? h->refill_at : 20;
    rp.node_burst = h->node_burst;
    rp.lanes_per_warp = h->lanes_per_warp;
    rp.retire_log = h->retire_log;
    rp.retire_log_warps = h->retire_log_warps;
    rp.fb_rgb = h->fb_rgb;
    rp.fb_yuv = h->fb_yuv;
    rp.counters = h->d_counters;
    rp.block_list = blocks_dev;
    rp.n_blocks = n_blocks;
    rp.block_cost = block_cost;
    rp.tiles.n = 1;  // work_total() reads the block list; keep the tile list well-formed for the launch-size computation
    rp.tiles.first_item[0] = 0;
    rp.tiles.first_item[1] = n_blocks * 32u;
    uint64_t seq = h->launch_seq.fetch_add(1);
    rp.work_counter = h->d_work + (seq % kCounterRing);
    PT_CUDA(h, cudaMemsetAsync(rp.work_counter, 0, sizeof(uint32_t), stream));
    PT_CUDA(h, launch(h, rp, stream));
    if (!block_cost) h->samples.fetch_add((uint64_t)n_blocks * 32u * spp);  // upper bound: edge blocks are partly outside the frame
    h->launches.fetch_add(1);
    return PT_OK;
}

int render_tiles(ptcore *h, const PtTile *tiles, int32_t n_tiles, cudaStream_t stream) {
    if (!h->have_scene) return fail(h, PT_ERR_NO_SCENE, "no scene uploaded");
    if (!h->fb_rgb) return fail(h, PT_ERR_NO_FRAMEBUFFER, "no framebuffer bound");
    if (!h->have_cam) return fail(h, PT_ERR_INVALID_ARGUMENT, "no camera set");
    PT_CUDA(h, cudaSetDevice(h->device));
    int32_t done = 0;
    while (done < n_tiles) {
        RenderParams rp;
        memset(&rp, 0, sizeof rp);
        rp.scene = h->dscene;
        rp.cam = h->cam;
        rp.width = h->fb_w;
        rp.height = h->fb_h;
        rp.spp = h->spp;
        rp.depth = h->depth;
        rp.refill_at = h->refill_at ? h->refill_at : 20;
        rp.node_burst = h->node_burst;
        rp.lanes_per_warp = h->lanes_per_warp;
        rp.retire_log = h->retire_log;
        rp.retire_log_warps = h->retire_log_warps;
        rp.fb_rgb = h->fb_rgb;
        rp.fb_yuv = h->fb_yuv;
        rp.counters = h->d_counters;
        rp.block_list = nullptr;
        rp.n_blocks = 0;
        rp.pad2 = 0;
        rp.block_cost = nullptr;
        TileList &tl = rp.tiles;
        tl.n = 0;
        tl.first_item[0] = 0;
        uint64_t pixels = 0;
        const bool direct = h->kernel == PT_KERNEL_DIRECT;
        while (done < n_tiles && tl.n < (direct ? 1 : kMaxInlineTiles)) {
            PtTile t = tiles[done];
            // clip to the framebuffer (the reference over-provisions its grid and relies on the i/j guard, :75)
            int32_t x0 = std::max(t.offset_x, 0), y0 = std::max(t.offset_y, 0);
            int32_t x1 = std::min<int64_t>((int64_t)t.offset_x + t.width, h->fb_w), y1 = std::min<int64_t>((int64_t)t.offset_y + t.height, h->fb_h);
            done++;
            if (t.width <= 0 || t.height <= 0 || x1 <= x0 || y1 <= y0) continue;  // renderTaskAsync :195: width == 0 is a no-op
            uint64_t items = 32ull * (uint64_t)((x1 - x0 + 7) / 8) * (uint64_t)((y1 - y0 + 3) / 4);
            if ((uint64_t)tl.first_item[tl.n] + items > 0xfffffff0ull) { done--; break; }
            tl.ox[tl.n] = x0; tl.oy[tl.n] = y0; tl.w[tl.n] = x1 - x0; tl.h[tl.n] = y1 - y0;
            tl.first_item[tl.n + 1] = tl.first_item[tl.n] + (uint32_t)items;
            pixels += (uint64_t)(x1 - x0) * (uint64_t)(y1 - y0);
            tl.n++;
        }
        if (tl.n == 0) continue;
        uint64_t seq = h->launch_seq.fetch_add(1);
        rp.work_counter = h->d_work + (seq % kCounterRing);
        if (!direct) PT_CUDA(h, cudaMemsetAsync(rp.work_counter, 0, sizeof(uint32_t), stream));
        PT_CUDA(h, launch(h, rp, stream));
        h->samples.fetch_add(pixels * h->spp);
        h->launches.fetch_add(1);
    }
    return PT_OK;
}

}  // namespace

extern "C" {

int ptcore_abi_version(void) { return PTCORE_ABI_VERSION; }

const char *ptcore_last_error(const ptcore_t *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int ptcore_create(int device, ptcore_t **out) {
    if (!out) return fail(nullptr, PT_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(nullptr, (int)e, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e) + " — libptcore needs a CUDA device; there is no CPU fallback");
    if (device < 0 || device >= n) return fail(nullptr, PT_ERR_INVALID_ARGUMENT, "device index out of range");
    ptcore *h = new ptcore();
    h->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_work, sizeof(uint32_t) * kCounterRing);
    if (e == cudaSuccess) e = cudaMemset(h->d_work, 0, sizeof(uint32_t) * kCounterRing);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_counters, sizeof(DevCounters));
    if (e == cudaSuccess) e = cudaMemset(h->d_counters, 0, sizeof(DevCounters));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_cam, sizeof(CamParams));
    if (e != cudaSuccess) {
        int code = fail(nullptr, (int)e, std::string("ptcore_create: ") + cudaGetErrorString(e));
        delete h;
        return code;
    }
    *out = h;
    return PT_OK;
}

int ptcore_destroy(ptcore_t *h) {
    if (!h) return PT_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->d_work);
    cudaFree(h->d_counters);
    cudaFree(h->d_cam);
    cudaFree(h->blob.dev);
    if (h->blob.pinned) cudaFreeHost(h->blob.pinned);
    cudaFree(h->own_rgb);
    cudaFree(h->own_yuv);
    cudaFree(h->ident_blocks);
    delete h;
    return PT_OK;
}

int ptcore_upload_scene(ptcore_t *h, const PtSceneDesc *sc) {
    if (!h || !sc) return fail(h, PT_ERR_INVALID_ARGUMENT, "null argument");
    if (sc->n_tris < 0 || sc->n_spheres < 0 || sc->n_mats <= 0 || sc->n_tex < 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "bad scene counts (at least one material is required)");
    if ((sc->n_tris && (!sc->tri_pos || !sc->tri_mat)) || (sc->n_spheres && (!sc->sph || !sc->sph_mat)) || !sc->mats || (sc->n_tex && !sc->tex))
        return fail(h, PT_ERR_INVALID_ARGUMENT, "scene arrays missing");
    const int64_t n_prims = (int64_t)sc->n_tris + sc->n_spheres;
    if (n_prims >= (1 << 27)) return fail(h, PT_ERR_UNSUPPORTED, "too many primitives for the leaf reference encoding");
    for (int32_t i = 0; i < sc->n_tris; i++)
        if (sc->tri_mat[i] < 0 || sc->tri_mat[i] >= sc->n_mats) return fail(h, PT_ERR_INVALID_ARGUMENT, "triangle material index out of range");
    for (int32_t i = 0; i < sc->n_spheres; i++)
        if (sc->sph_mat[i] < 0 || sc->sph_mat[i] >= sc->n_mats) return fail(h, PT_ERR_INVALID_ARGUMENT, "sphere material index out of range");
    for (int32_t i = 0; i < sc->n_mats; i++) {
        const PtMaterial &m = sc->mats[i];
        if (m.type < 0 || m.type > PT_MAT_UNIVERSAL) return fail(h, PT_ERR_INVALID_ARGUMENT, "unknown material type");
        if (m.base_tex >= sc->n_tex || m.emis_tex >= sc->n_tex) return fail(h, PT_ERR_INVALID_ARGUMENT, "material texture index out of range");
    }
    PT_CUDA(h, cudaSetDevice(h->device));
    auto t0 = std::chrono::high_resolution_clock::now();

    std::vector<PrimBounds> pb = padded_bounds(sc);

    BvhBuildOptions opt;
    opt.leaf_max = h->leaf_max;
    opt.intersect_cost = (float)h->sah_isect_x100 / 100.0f;
    BvhBuildResult bvh = build_bvh(pb, opt);

    // ---- lights: DevicePathTracer.h:302-307 (emissiveFactor channel > 0.0001, scene order) ----
    std::vector<int32_t> lights;
    for (int32_t i = 0; i < sc->n_tris; i++) {
        const PtMaterial &m = sc->mats[sc->tri_mat[i]];
        if (m.type == PT_MAT_UNIVERSAL && (m.emis[0] > 0.0001 || m.emis[1] > 0.0001 || m.emis[2] > 0.0001)) lights.push_back(i);
    }

    // ---- blob layout ----
    SceneBlob nb;
    nb.n_prims = (int32_t)n_prims;
    nb.n_lights = (int32_t)lights.size();
    nb.n_nodes = (int32_t)bvh.nodes.size();
    size_t texels = 0;
    for (int32_t i = 0; i < sc->n_tex; i++) {
        if (sc->tex[i].width < 0 || sc->tex[i].height < 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "negative texture size");
        texels += (size_t)sc->tex[i].width * (size_t)sc->tex[i].height;
    }
    if (texels >= 0xffffffffull) return fail(h, PT_ERR_UNSUPPORTED, "textures too large");
    size_t off = 0;
    nb.off_nodes = off; off = align_up(off + bvh.nodes.size() * sizeof(FlatNode), 256);
    std::vector<QuantNode> nodesq;
    nb.grid = quantise_nodes(bvh.nodes, nodesq);
    nb.off_nodesq = off; off = align_up(off + nodesq.size() * sizeof(QuantNode), 256);
    // the four-wide copy of the tree is only uploaded when it is selected or the scene is small: on large scenes it would
    // just compete with the two-wide nodes for L2
    nb.has_nodes4 = h->bvh_width == 4 || n_prims <= (1 << 18);
    nb.off_nodes4 = off; off = align_up(off + (nb.has_nodes4 ? bvh.nodes4.size() : 1) * sizeof(FlatNode4), 256);
#if PT_INDEXED_PRIMS
    // unique vertices (exact bit patterns) in order of first use along the leaf order, so that neighbouring leaves share lines
    std::vector<uint32_t> prim_index((size_t)n_prims * 4, 0u);
    std::vector<float> verts;
    {
        const size_t cap_hint = (size_t)sc->n_tris * 3 + (size_t)sc->n_spheres + 1;
        size_t cap = 16;
        while (cap < cap_hint * 2) cap <<= 1;
        std::vector<uint32_t> table(cap, 0xffffffffu);
        verts.reserve(((size_t)sc->n_tris / 2 + (size_t)sc->n_spheres + 16) * 4);
        auto intern = [&](const float *p3) -> uint32_t {
            uint32_t b[3];
            memcpy(b, p3, 12);
            uint64_t hsh = (uint64_t)b[0] * 0x9E3779B97F4A7C15ull ^ ((uint64_t)b[1] * 0xC2B2AE3D27D4EB4Full + ((uint64_t)b[2] << 32 | b[2]) * 0x165667B19E3779F9ull);
            hsh ^= hsh >> 29;
            size_t slot = (size_t)hsh & (cap - 1);
            for (;;) {
                const uint32_t id = table[slot];
                if (id == 0xffffffffu) break;
                if (memcmp(&verts[(size_t)id * 4], p3, 12) == 0) return id;
                slot = (slot + 1) & (cap - 1);
            }
            const uint32_t id = (uint32_t)(verts.size() / 4);
            verts.insert(verts.end(), {p3[0], p3[1], p3[2], 0.f});
            table[slot] = id;
            return id;
        };
        for (int64_t k = 0; k < n_prims; k++) {
            const int32_t id = bvh.prim_order[(size_t)k];
            uint32_t *r = &prim_index[(size_t)k * 4];
            if (id < sc->n_tris) {
                const float *p = sc->tri_pos + (size_t)id * 9;
                r[0] = intern(p); r[1] = intern(p + 3); r[2] = intern(p + 6);
                r[3] = 0u | ((uint32_t)shade_class(sc->mats[sc->tri_mat[id]]) << 8);
            } else {
                const int32_t si = id - sc->n_tris;
                const float *sp = sc->sph + (size_t)si * 4;
                const uint32_t v = (uint32_t)(verts.size() / 4);  // spheres are not shared: (centre, radius)
                verts.insert(verts.end(), {sp[0], sp[1], sp[2], sp[3]});
                r[0] = r[1] = r[2] = v;
                r[3] = 1u | ((uint32_t)shade_class(sc->mats[sc->sph_mat[si]]) << 8);
            }
        }
    }
    nb.off_prims = off; off = align_up(off + std::max<size_t>(1, (size_t)n_prims) * 16, 256);
    nb.off_verts = off; off = align_up(off + std::max<size_t>(4, verts.size()) * sizeof(float), 256);
#else
    nb.off_prims = off; off = align_up(off + std::max<size_t>(1, (size_t)n_prims) * 48, 256);
    nb.off_verts = nb.off_prims;
#endif
    nb.off_shade = off; off = align_up(off + std::max<size_t>(1, (size_t)n_prims) * 32, 256);
    nb.off_frames = off; off = align_up(off + std::max<size_t>(1, (size_t)n_prims) * 64, 256);
    nb.off_mats = off; off = align_up(off + (size_t)sc->n_mats * 48, 256);
    nb.off_texdesc = off; off = align_up(off + std::max<size_t>(1, (size_t)sc->n_tex) * sizeof(TexDesc), 256);
    nb.off_texels = off; off = align_up(off + std::max<size_t>(1, texels) * 16, 256);
    nb.off_lights = off; off = align_up(off + std::max<size_t>(1, lights.size()) * 64, 256);
    nb.bytes = off;
    nb.host.assign(nb.bytes, 0);

    memcpy(nb.host.data() + nb.off_nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(FlatNode));
    memcpy(nb.host.data() + nb.off_nodesq, nodesq.data(), nodesq.size() * sizeof(QuantNode));
    if (nb.has_nodes4) {
        memcpy(nb.host.data() + nb.off_nodes4, bvh.nodes4.data(), bvh.nodes4.size() * sizeof(FlatNode4));
        if (bvh.stack4 > (uint32_t)kStackSize - 2) {
            if (h->bvh_width == 4) return fail(h, PT_ERR_UNSUPPORTED, "scene needs a deeper traversal stack than the device provides for four-wide nodes");
            nb.has_nodes4 = false;
        }
    }
#if PT_INDEXED_PRIMS
    memcpy(nb.host.data() + nb.off_prims, prim_index.data(), prim_index.size() * sizeof(uint32_t));
    memcpy(nb.host.data() + nb.off_verts, verts.data(), verts.size() * sizeof(float));
    h->build_stats.n_vertices = (uint32_t)(verts.size() / 4);
#else
    float *prims = reinterpret_cast<float *>(nb.host.data() + nb.off_prims);
    h->build_stats.n_vertices = 0;
#endif
    float *shade = reinterpret_cast<float *>(nb.host.data() + nb.off_shade);
    float *frames = reinterpret_cast<float *>(nb.host.data() + nb.off_frames);
    for (int64_t k = 0; k < n_prims; k++) {
        int32_t id = bvh.prim_order[(size_t)k];
#if !PT_INDEXED_PRIMS
        float *q = prims + k * 12;
#else
        float qdummy[12];
        float *q = qdummy;
#endif
        float *s = shade + k * 8;
        float *f = frames + k * 16;  // the vectors are filled in on the device (pt_frames_kernel)
        if (id < sc->n_tris) {
            const float *p = sc->tri_pos + (size_t)id * 9;
            q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
            q[3] = p[3] - p[0]; q[4] = p[4] - p[1]; q[5] = p[5] - p[2];  // e1 = v1 - v0, triangle.h:67
            q[6] = p[6] - p[0]; q[7] = p[7] - p[1]; q[8] = p[8] - p[2];  // e2 = v2 - v0, triangle.h:68
            q[9] = 0.f; q[10] = as_float(0); q[11] = as_float(shade_class(sc->mats[sc->tri_mat[id]]));
            if (sc->tri_uv) memcpy(s, sc->tri_uv + (size_t)id * 6, 6 * sizeof(float));
            s[6] = as_float(sc->tri_mat[id]);
            f[3] = s[6];
            f[7] = as_float(0);
        } else {
            const int32_t si = id - sc->n_tris;
            const float *sp = sc->sph + (size_t)si * 4;
            q[0] = sp[0]; q[1] = sp[1]; q[2] = sp[2]; q[3] = sp[3];
            q[10] = as_float(1);
            q[11] = as_float(shade_class(sc->mats[sc->sph_mat[si]]));
            s[6] = as_float(sc->sph_mat[si]);
            f[3] = s[6];
            f[7] = as_float(1);
            nb.has_spheres = true;
        }
        s[7] = as_float(id);
    }
    float *mats = reinterpret_cast<float *>(nb.host.data() + nb.off_mats);
    // DevicePathTracer.h:269-279 (loadMaterials): the texture pointers live outside the loop over the materials, so a
    // UniversalMaterial without a texture of its own inherits the last one seen.  Reproduced here, at the one place every front
    // door (C ABI, C++ shims, Python, CLI) goes through; PT_OPT_STICKY_TEXTURES = 0 switches the quirk off.
    int32_t sticky_base = -1, sticky_emis = -1;
    for (int32_t i = 0; i < sc->n_mats; i++) {
        PtMaterial m = sc->mats[i];
        if (h->sticky_textures && m.type == PT_MAT_UNIVERSAL) {
            if (m.base_tex >= 0) sticky_base = m.base_tex;
            if (m.emis_tex >= 0) sticky_emis = m.emis_tex;
            m.base_tex = sticky_base;
            m.emis_tex = sticky_emis;
        }
        float *q = mats + (size_t)i * 12;
        q[0] = as_float(m.type); q[1] = m.base[0]; q[2] = m.base[1]; q[3] = m.base[2];
        q[4] = m.emis[0]; q[5] = m.emis[1]; q[6] = m.emis[2]; q[7] = as_float(m.base_tex < 0 ? -1 : m.base_tex);
        q[8] = as_float(m.emis_tex < 0 ? -1 : m.emis_tex); q[9] = m.fuzz; q[10] = m.ior; q[11] = 0.f;
        if (m.type != PT_MAT_UNIVERSAL) nb.has_rtow = true;
    }
    TexDesc *td = reinterpret_cast<TexDesc *>(nb.host.data() + nb.off_texdesc);
    float *tx = reinterpret_cast<float *>(nb.host.data() + nb.off_texels);
    size_t texel_off = 0;
    for (int32_t i = 0; i < sc->n_tex; i++) {
        td[i].width = sc->tex[i].width;
        td[i].height = sc->tex[i].rgb ? sc->tex[i].height : 0;
        td[i].offset = (uint32_t)texel_off;
        td[i].pad = 0;
        size_t n = (size_t)sc->tex[i].width * (size_t)sc->tex[i].height;
        if (sc->tex[i].rgb) {
            const double color_scale = 1.0 / 255.0;  // Texture.h:45-46: the scale is a double, the product is rounded to float once
            for (size_t k = 0; k < n; k++) {
                float *o = tx + (texel_off + k) * 4;
                o[0] = (float)(color_scale * sc->tex[i].rgb[k * 3 + 0]);
                o[1] = (float)(color_scale * sc->tex[i].rgb[k * 3 + 1]);
                o[2] = (float)(color_scale * sc->tex[i].rgb[k * 3 + 2]);
                o[3] = 0.f;
            }
        }
        texel_off += n;
    }
    float *lt = reinterpret_cast<float *>(nb.host.data() + nb.off_lights);
    for (size_t i = 0; i < lights.size(); i++) {
        const float *p = sc->tri_pos + (size_t)lights[i] * 9;
        float *q = lt + i * 16;
        q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = triangle_area_host(p);
        q[4] = p[3]; q[5] = p[4]; q[6] = p[5]; q[7] = 0.f;
        q[8] = p[6]; q[9] = p[7]; q[10] = p[8]; q[11] = 0.f;
    }
    auto t1 = std::chrono::high_resolution_clock::now();

    // ---- one bulk upload ----
    PT_CUDA(h, cudaDeviceSynchronize());
    uint8_t *dev = nullptr, *pinned = nullptr;
    PT_CUDA(h, cudaMalloc(&dev, nb.bytes));
    cudaError_t e = cudaMallocHost(&pinned, nb.bytes);
    if (e != cudaSuccess) { cudaFree(dev); return fail(h, (int)e, std::string("cudaMallocHost: ") + cudaGetErrorString(e)); }
    memcpy(pinned, nb.host.data(), nb.bytes);
    e = cudaMemcpy(dev, pinned, nb.bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); cudaFreeHost(pinned); return fail(h, (int)e, std::string("scene upload: ") + cudaGetErrorString(e)); }
    cudaFree(h->blob.dev);
    if (h->blob.pinned) cudaFreeHost(h->blob.pinned);
    nb.dev = dev;
    nb.pinned = pinned;
    nb.host.clear();
    nb.host.shrink_to_fit();
    h->blob = std::move(nb);
    fill_dev_scene(h);
    // per-primitive shading frames and light normals, evaluated on the device (rsqrtf is not reproducible on the host) and
    // copied back into the pinned image of the blob so that ptcore_reupload_scene restores them too
    {
        SceneBlob &b = h->blob;
        const int n = std::max(b.n_prims, b.n_lights);
        if (n > 0) {
            pt_frames_kernel<<<(n + 255) / 256, 256>>>(h->dscene, reinterpret_cast<float4 *>(b.dev + b.off_frames), reinterpret_cast<float4 *>(b.dev + b.off_lights));
            PT_CUDA(h, cudaGetLastError());
            PT_CUDA(h, cudaMemcpy(b.pinned + b.off_frames, b.dev + b.off_frames, (size_t)std::max(1, b.n_prims) * 64, cudaMemcpyDeviceToHost));
            PT_CUDA(h, cudaMemcpy(b.pinned + b.off_lights, b.dev + b.off_lights, (size_t)std::max(1, b.n_lights) * 64, cudaMemcpyDeviceToHost));
        }
    }
    h->have_scene = true;

    h->build_stats.bvh_nodes = (uint32_t)bvh.nodes.size();
    h->build_stats.bvh_leaves = bvh.n_leaves;
    h->build_stats.bvh_depth = bvh.depth;
    h->build_stats.bvh4_nodes = (uint32_t)bvh.nodes4.size();
    h->build_stats.bvh4_depth = bvh.depth4;
    validate_quantised(bvh.nodes, nodesq, h->blob.grid, &h->build_stats.quant_inflation);
    h->build_stats.n_lights = (uint32_t)lights.size();
    h->build_stats.bvh_build_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    h->build_stats.sah_cost = bvh.sah_cost;
    h->build_stats.scene_bytes = h->blob.bytes;
    return PT_OK;
}

int ptcore_reupload_scene(ptcore_t *h, void *stream, uint64_t *bytes_copied) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    if (!h->have_scene) return fail(h, PT_ERR_NO_SCENE, "no scene uploaded");
    PT_CUDA(h, cudaSetDevice(h->device));
    PT_CUDA(h, cudaMemcpyAsync(h->blob.dev, h->blob.pinned, h->blob.bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (bytes_copied) *bytes_copied = h->blob.bytes;
    return PT_OK;
}

int ptcore_set_camera(ptcore_t *h, const PtCamera *cam) {
    if (!h || !cam) return fail(h, PT_ERR_INVALID_ARGUMENT, "null argument");
    PT_CUDA(h, cudaSetDevice(h->device));
    pt_camera_kernel<<<1, 1>>>(make_float3(cam->look_from[0], cam->look_from[1], cam->look_from[2]), make_float3(cam->front[0], cam->front[1], cam->front[2]),
                               cam->vfov, cam->hfov, h->d_cam);
    PT_CUDA(h, cudaGetLastError());
    PT_CUDA(h, cudaMemcpy(&h->cam, h->d_cam, sizeof(CamParams), cudaMemcpyDeviceToHost));
    h->have_cam = true;
    return PT_OK;
}

int ptcore_set_params(ptcore_t *h, uint32_t spp, uint32_t depth) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    if (spp == 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "samples_per_pixel must be >= 1");
    h->spp = spp;
    h->depth = depth;
    return PT_OK;
}

int ptcore_set_thread_block_size(ptcore_t *h, uint32_t bx, uint32_t by) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    if (bx == 0 || by == 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "thread block size must be non-zero");
    h->bx = bx;
    h->by = by;
    return PT_OK;
}

int ptcore_set_option(ptcore_t *h, int key, int64_t value) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    switch (key) {
        case PT_OPT_KERNEL:
            if (value != PT_KERNEL_PERSISTENT && value != PT_KERNEL_DIRECT && value != PT_KERNEL_LOCKSTEP && value != PT_KERNEL_POOL) return fail(h, PT_ERR_INVALID_ARGUMENT, "unknown kernel");
            h->kernel = (int)value;
            return PT_OK;
        case PT_OPT_COUNT_TESTS: h->count_tests = value != 0; return PT_OK;
        case PT_OPT_BVH_LEAF_MAX:
            if (value < 1 || value > kMaxLeafPrims) return fail(h, PT_ERR_INVALID_ARGUMENT, "leaf_max must be in [1, 8]");
            h->leaf_max = (int)value;
            return PT_OK;
        case PT_OPT_BLOCKS_PER_SM:
            if (value < 0 || value > 32) return fail(h, PT_ERR_INVALID_ARGUMENT, "blocks_per_sm must be in [0, 32]");
            h->blocks_per_sm = (int)value;
            return PT_OK;
        case PT_OPT_REFILL_AT:
            if (value < 0 || value > 32) return fail(h, PT_ERR_INVALID_ARGUMENT, "refill_at must be in [0, 32]");
            h->refill_at = (int)value;
            return PT_OK;
        case PT_OPT_NODE_BURST:
            if (value < 1 || value > 4) return fail(h, PT_ERR_INVALID_ARGUMENT, "node_burst must be in [1, 4]");
            h->node_burst = (int)value;
            return PT_OK;
        case PT_OPT_MIN_BLOCKS:
            if (value != 8) return fail(h, PT_ERR_INVALID_ARGUMENT, "only the __launch_bounds__(128, 8) build is shipped");
            h->min_blocks = (int)value;
            return PT_OK;
        case PT_OPT_BVH_WIDTH:
            if (value != 2 && value != 4) return fail(h, PT_ERR_INVALID_ARGUMENT, "bvh_width must be 2 or 4");
            h->bvh_width = (int)value;
            return PT_OK;
        case PT_OPT_SAH_INTERSECT_COST:
            if (value < 10 || value > 1000) return fail(h, PT_ERR_INVALID_ARGUMENT, "sah intersect cost must be in [10, 1000] hundredths");
            h->sah_isect_x100 = (int)value;
            return PT_OK;
        case PT_OPT_NODE_FORMAT:
            if (value != PT_NODES_AUTO && value != PT_NODES_FULL && value != PT_NODES_QUANTISED) return fail(h, PT_ERR_INVALID_ARGUMENT, "unknown node format");
            h->node_format = (int)value;
            h->trace_steps = value != PT_NODES_AUTO;
            return PT_OK;
        case PT_OPT_POOL_SLOTS:
            if (value != 0 && (value < 32 || value > kPoolSlots)) return fail(h, PT_ERR_INVALID_ARGUMENT, "pool_slots must be 0 (auto) or in [32, 96]");
            h->pool_slots = (int)value;
            return PT_OK;
        case PT_OPT_POOL_IDLE_AT:
            if (value < 1 || value > 32) return fail(h, PT_ERR_INVALID_ARGUMENT, "pool_idle_at must be in [1, 32]");
            h->pool_idle_at = (int)value;
            return PT_OK;
        case PT_OPT_POOL_PERIOD:
            if (value != 1 && value != 2 && value != 4 && value != 8) return fail(h, PT_ERR_INVALID_ARGUMENT, "pool_period must be 1, 2, 4 or 8");
            h->pool_period = (int)value;
            return PT_OK;
        case PT_OPT_POOL_CARVEOUT:
            if (value < -1 || value > 100) return fail(h, PT_ERR_INVALID_ARGUMENT, "pool_carveout is a percentage (or -1 for the driver's default)");
            h->pool_carveout = (int)value;
            return PT_OK;
        case PT_OPT_SMEM_NODES:
            if (value != 0 && value != 1) return fail(h, PT_ERR_INVALID_ARGUMENT, "smem_nodes must be 0 or 1");
            h->smem_nodes = (int)value;
            return PT_OK;
        case PT_OPT_LANES_PER_WARP:
            if (value < 1 || value > 32) return fail(h, PT_ERR_INVALID_ARGUMENT, "lanes_per_warp must be in [1, 32]");
            h->lanes_per_warp = (int)value;
            return PT_OK;
        case PT_OPT_GRID_CTAS:
            if (value < 0 || value > 65535) return fail(h, PT_ERR_INVALID_ARGUMENT, "grid_ctas must be in [0, 65535]");
            h->grid_ctas = (int)value;
            return PT_OK;
        case PT_OPT_CTA_WARPS:
            if (value < 0 || value > 32) return fail(h, PT_ERR_INVALID_ARGUMENT, "cta_warps must be in [0, 32]");
            h->cta_warps = (int)value;
            return PT_OK;
        case PT_OPT_STICKY_TEXTURES: h->sticky_textures = value != 0; return PT_OK;
        case PT_OPT_L2_PERSIST_NODES: h->l2_persist_nodes = value != 0; return PT_OK;
        case PT_OPT_RNG_MODE:
            if (value != PT_RNG_STREAM && value != PT_RNG_SAMPLE_KEYED) return fail(h, PT_ERR_INVALID_ARGUMENT, "unknown rng mode");
            h->rng_mode = (int)value;
            return PT_OK;
        case PT_OPT_RNG_CHUNKS:
            if (value < 1 || value > 4096) return fail(h, PT_ERR_INVALID_ARGUMENT, "rng_chunks must be in [1, 4096]");
            h->rng_chunks = (int)value;
            return PT_OK;
        case PT_OPT_WATCHDOG:
            if (value < 0 || value > 0xffffffffll) return fail(h, PT_ERR_INVALID_ARGUMENT, "watchdog must fit 32 bits");
            h->watchdog = (uint32_t)value;
            return PT_OK;
        default: return fail(h, PT_ERR_UNSUPPORTED, "unknown option key");
    }
}

int ptcore_bind_framebuffer(ptcore_t *h, uint8_t *rgb, uint8_t *yuv, uint32_t width, uint32_t height) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    if (!rgb || width == 0 || height == 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "framebuffer needs an RGB pointer and non-zero size");
    if ((uint64_t)width * height > 0x7fffffffull / 3) return fail(h, PT_ERR_UNSUPPORTED, "framebuffer too large");
    h->fb_rgb = rgb;
    h->fb_yuv = yuv;
    h->fb_w = width;
    h->fb_h = height;
    return PT_OK;
}

int ptcore_render_tile_async(ptcore_t *h, int32_t ox, int32_t oy, int32_t w, int32_t hgt, void *stream) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    PtTile t{w, hgt, ox, oy};
    return render_tiles(h, &t, 1, (cudaStream_t)stream);
}

int ptcore_render_tiles_async(ptcore_t *h, const PtTile *tiles, int32_t n, void *stream) {
    if (!h || (n > 0 && !tiles) || n < 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "bad tile list");
    return render_tiles(h, tiles, n, (cudaStream_t)stream);
}

int ptcore_render_blocks_async(ptcore_t *h, const uint32_t *blocks_dev, uint32_t n_blocks, void *stream) {
    if (!h || (n_blocks && !blocks_dev)) return fail(h, PT_ERR_INVALID_ARGUMENT, "bad block list");
    return render_blocks(h, blocks_dev, n_blocks, h->spp, nullptr, (cudaStream_t)stream);
}

static int block_costs_range(ptcore_t *h, uint32_t pilot_spp, uint32_t *costs_dev, uint32_t first_block, uint32_t n_range, bool all, void *stream) {
    if (!h || !costs_dev || pilot_spp == 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "bad arguments");
    if (!h->fb_rgb) return fail(h, PT_ERR_NO_FRAMEBUFFER, "no framebuffer bound");
    PT_CUDA(h, cudaSetDevice(h->device));
    const uint32_t bw = (h->fb_w + 7) / 8, bh = (h->fb_h + 3) / 4, n = bw * bh;
    // identity block list of the bound framebuffer (keyed on its shape: two shapes can have the same number of blocks)
    if (h->ident_blocks_n != n || h->ident_bw != bw) {
        std::vector<uint32_t> ident(n);
        for (uint32_t by = 0; by < bh; by++)
            for (uint32_t bx = 0; bx < bw; bx++) ident[by * bw + bx] = bx | (by << 16);
        cudaFree(h->ident_blocks);
        h->ident_blocks = nullptr;
        h->ident_blocks_n = 0;
        PT_CUDA(h, cudaMalloc(&h->ident_blocks, sizeof(uint32_t) * n));
        PT_CUDA(h, cudaMemcpy(h->ident_blocks, ident.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
        h->ident_blocks_n = n;
        h->ident_bw = bw;
    }
    if (all) {
        first_block = 0;
        n_range = n;
    }
    if (first_block > n) first_block = n;
    if (n_range > n - first_block) n_range = n - first_block;
    PT_CUDA(h, cudaMemsetAsync(costs_dev, 0, sizeof(uint32_t) * n, (cudaStream_t)stream));
    return render_blocks(h, h->ident_blocks + first_block, n_range, pilot_spp, costs_dev, (cudaStream_t)stream);
}

int ptcore_block_costs_async(ptcore_t *h, uint32_t pilot_spp, uint32_t *costs_dev, void *stream) {
    return block_costs_range(h, pilot_spp, costs_dev, 0, 0, true, stream);
}

int ptcore_block_costs_range_async(ptcore_t *h, uint32_t pilot_spp, uint32_t *costs_dev, uint32_t first_block, uint32_t n_blocks, void *stream) {
    return block_costs_range(h, pilot_spp, costs_dev, first_block, n_blocks, false, stream);
}

static int keyed_identity_blocks(ptcore_t *h) {
    const uint32_t bw = (h->fb_w + 7) / 8, bh = (h->fb_h + 3) / 4, n = bw * bh;
    if (h->ident_blocks_n == n && h->ident_bw == bw) return PT_OK;
    std::vector<uint32_t> ident(n);
    for (uint32_t by = 0; by < bh; by++)
        for (uint32_t bx = 0; bx < bw; bx++) ident[by * bw + bx] = bx | (by << 16);
    cudaFree(h->ident_blocks);
    h->ident_blocks = nullptr;
    h->ident_blocks_n = 0;
    PT_CUDA(h, cudaMalloc(&h->ident_blocks, sizeof(uint32_t) * n));
    PT_CUDA(h, cudaMemcpy(h->ident_blocks, ident.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    h->ident_blocks_n = n;
    h->ident_bw = bw;
    return PT_OK;
}

int ptcore_render_keyed_async(ptcore_t *h, const uint32_t *blocks_dev, uint32_t n_blocks, float *accum_dev, uint32_t n_chunks, uint32_t first_chunk, uint32_t chunk_step,
                              void *stream) {
    if (!h || !accum_dev || n_chunks == 0 || chunk_step == 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "bad arguments");
    if (!h->have_scene) return fail(h, PT_ERR_NO_SCENE, "no scene uploaded");
    if (!h->fb_rgb) return fail(h, PT_ERR_NO_FRAMEBUFFER, "no framebuffer bound");
    if (!h->have_cam) return fail(h, PT_ERR_INVALID_ARGUMENT, "no camera set");
    if (h->depth == 0) return fail(h, PT_ERR_UNSUPPORTED, "keyed rendering needs recursion depth >= 1");
    PT_CUDA(h, cudaSetDevice(h->device));
    if (!blocks_dev) {
        int rc = keyed_identity_blocks(h);
        if (rc != PT_OK) return rc;
        blocks_dev = h->ident_blocks;
        n_blocks = h->ident_blocks_n;
    }
    if (first_chunk >= n_chunks || n_blocks == 0) return PT_OK;
    const uint32_t my_chunks = (n_chunks - first_chunk + chunk_step - 1) / chunk_step;
    if ((uint64_t)n_blocks * 32ull * my_chunks > 0xfffffff0ull) return fail(h, PT_ERR_UNSUPPORTED, "too many work items");
    RenderParams rp;
    memset(&rp, 0, sizeof rp);
    rp.scene = h->dscene;
    rp.cam = h->cam;
    rp.width = h->fb_w;
    rp.height = h->fb_h;
    rp.spp = h->spp;
    rp.depth = h->depth;
    rp.refill_at = h->refill_at ? h->refill_at : 20;
    rp.node_burst = h->node_burst;
    rp.lanes_per_warp = h->lanes_per_warp;
    rp.retire_log = h->retire_log;
    rp.retire_log_warps = h->retire_log_warps;
    rp.fb_rgb = h->fb_rgb;
    rp.fb_yuv = h->fb_yuv;
    rp.counters = h->d_counters;
    rp.block_list = blocks_dev;
    rp.n_blocks = n_blocks;
    rp.tiles.n = 1;
    rp.tiles.first_item[0] = 0;
    rp.tiles.first_item[1] = n_blocks * 32u;
    rp.keyed_accum = accum_dev;
    rp.keyed_chunk_spp = (h->spp + n_chunks - 1) / n_chunks;
    rp.keyed_first = first_chunk;
    rp.keyed_step = chunk_step;
    rp.keyed_my_chunks = my_chunks;
    uint64_t seq = h->launch_seq.fetch_add(1);
    rp.work_counter = h->d_work + (seq % kCounterRing);
    PT_CUDA(h, cudaMemsetAsync(rp.work_counter, 0, sizeof(uint32_t), (cudaStream_t)stream));
    PT_CUDA(h, launch_keyed(h, rp, (cudaStream_t)stream));
    h->samples.fetch_add((uint64_t)n_blocks * 32u * (uint64_t)std::min<uint64_t>(h->spp, (uint64_t)rp.keyed_chunk_spp * my_chunks));
    h->launches.fetch_add(1);
    return PT_OK;
}

int ptcore_resolve_keyed_async(ptcore_t *h, const float *accum_dev, uint32_t n_chunks, void *stream) {
    if (!h || !accum_dev || n_chunks == 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "bad arguments");
    if (!h->fb_rgb) return fail(h, PT_ERR_NO_FRAMEBUFFER, "no framebuffer bound");
    PT_CUDA(h, cudaSetDevice(h->device));
    RenderParams rp;
    memset(&rp, 0, sizeof rp);
    rp.width = h->fb_w;
    rp.height = h->fb_h;
    rp.spp = h->spp;
    rp.fb_rgb = h->fb_rgb;
    rp.fb_yuv = h->fb_yuv;
    rp.keyed_accum = const_cast<float *>(accum_dev);
    const uint32_t npix = h->fb_w * h->fb_h;
    pt_resolve_keyed_kernel<<<(npix + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rp, n_chunks);
    PT_CUDA(h, cudaGetLastError());
    h->launches.fetch_add(1);
    return PT_OK;
}

int ptcore_gather_blocks_async(ptcore_t *h, uint8_t *dst_rgb, uint8_t *dst_yuv, const uint32_t *blocks_dev, uint32_t n_blocks, void *stream) {
    if (!h || !dst_rgb || (n_blocks && !blocks_dev)) return fail(h, PT_ERR_INVALID_ARGUMENT, "bad arguments");
    if (!h->fb_rgb) return fail(h, PT_ERR_NO_FRAMEBUFFER, "no framebuffer bound");
    if (n_blocks == 0 || dst_rgb == h->fb_rgb) return PT_OK;
    PT_CUDA(h, cudaSetDevice(h->device));
    const uint32_t threads = n_blocks * 32u;
    pt_gather_blocks_kernel<<<(threads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dst_rgb, dst_yuv, h->fb_rgb, h->fb_yuv, blocks_dev, n_blocks, h->fb_w, h->fb_h);
    PT_CUDA(h, cudaGetLastError());
    h->launches.fetch_add(1);
    return PT_OK;
}

int ptcore_set_retire_log(ptcore_t *h, uint64_t *log_dev, uint32_t n_warps) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    h->retire_log = reinterpret_cast<unsigned long long *>(log_dev);
    h->retire_log_warps = log_dev ? n_warps : 0;
    return PT_OK;
}

int ptcore_sync(ptcore_t *h, void *stream) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    PT_CUDA(h, cudaSetDevice(h->device));
    PT_CUDA(h, cudaGetLastError());
    PT_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
    return PT_OK;
}

int ptcore_wait(ptcore_t *h) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    PT_CUDA(h, cudaSetDevice(h->device));
    PT_CUDA(h, cudaGetLastError());
    PT_CUDA(h, cudaDeviceSynchronize());
    return PT_OK;
}

int ptcore_render_frame_host(ptcore_t *h, uint32_t width, uint32_t height, uint8_t *rgb_host, uint8_t *yuv_host) {
    if (!h || !rgb_host || width == 0 || height == 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "bad arguments");
    PT_CUDA(h, cudaSetDevice(h->device));
    const size_t px = (size_t)width * height;
    if (h->own_w != width || h->own_h != height) {
        cudaFree(h->own_rgb);
        cudaFree(h->own_yuv);
        h->own_rgb = h->own_yuv = nullptr;
        h->own_w = h->own_h = 0;
        PT_CUDA(h, cudaMalloc(&h->own_rgb, px * 3));
        PT_CUDA(h, cudaMalloc(&h->own_yuv, px * 3 / 2 + 2));
        h->own_w = width;
        h->own_h = height;
    }
    uint8_t *srgb = h->fb_rgb, *syuv = h->fb_yuv;
    uint32_t sw = h->fb_w, sh = h->fb_h;
    h->fb_rgb = h->own_rgb;
    h->fb_yuv = yuv_host ? h->own_yuv : nullptr;
    h->fb_w = width;
    h->fb_h = height;
    int rc;
    if (h->rng_mode == PT_RNG_SAMPLE_KEYED && h->depth > 0) {
        const uint32_t n_chunks = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)h->rng_chunks, h->spp));
        float *accum = nullptr;
        const size_t bytes = (size_t)n_chunks * px * 3 * sizeof(float);
        cudaError_t e = cudaMalloc(&accum, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(accum, 0, bytes, nullptr);
        rc = e == cudaSuccess ? ptcore_render_keyed_async(h, nullptr, 0, accum, n_chunks, 0, 1, nullptr) : fail(h, (int)e, std::string("keyed accumulation buffer: ") + cudaGetErrorString(e));
        if (rc == PT_OK) rc = ptcore_resolve_keyed_async(h, accum, n_chunks, nullptr);
        cudaStreamSynchronize(nullptr);
        cudaFree(accum);
    } else {
        PtTile t{(int32_t)width, (int32_t)height, 0, 0};
        rc = render_tiles(h, &t, 1, nullptr);
    }
    h->fb_rgb = srgb; h->fb_yuv = syuv; h->fb_w = sw; h->fb_h = sh;
    if (rc != PT_OK) return rc;
    PT_CUDA(h, cudaMemcpyAsync(rgb_host, h->own_rgb, px * 3, cudaMemcpyDeviceToHost, nullptr));
    if (yuv_host) PT_CUDA(h, cudaMemcpyAsync(yuv_host, h->own_yuv, px * 3 / 2, cudaMemcpyDeviceToHost, nullptr));
    PT_CUDA(h, cudaStreamSynchronize(nullptr));
    return PT_OK;
}

int ptcore_get_stats(ptcore_t *h, PtStats *out) {
    if (!h || !out) return fail(h, PT_ERR_INVALID_ARGUMENT, "null argument");
    PT_CUDA(h, cudaSetDevice(h->device));
    PT_CUDA(h, cudaDeviceSynchronize());
    DevCounters c;
    PT_CUDA(h, cudaMemcpy(&c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost));
    *out = h->build_stats;
    out->samples = h->samples.load();
    out->rays = c.rays;
    out->box_tests = c.box_tests;
    out->tri_tests = c.tri_tests;
    out->light_tests = c.light_tests;
    out->launches = h->launches.load();
    return PT_OK;
}

int ptcore_reset_stats(ptcore_t *h) {
    if (!h) return PT_ERR_INVALID_ARGUMENT;
    PT_CUDA(h, cudaSetDevice(h->device));
    PT_CUDA(h, cudaDeviceSynchronize());
    PT_CUDA(h, cudaMemset(h->d_counters, 0, sizeof(DevCounters)));
    h->samples = 0;
    h->launches = 0;
    return PT_OK;
}

int ptcore_debug_trace_pixel(ptcore_t *h, uint32_t width, uint32_t height, int32_t x, int32_t y, float *events, int32_t max_events, int32_t *n_events, float *col) {
    if (!h || !events || !n_events || !col || max_events <= 0) return fail(h, PT_ERR_INVALID_ARGUMENT, "bad arguments");
    if (!h->have_scene) return fail(h, PT_ERR_NO_SCENE, "no scene uploaded");
    if (!h->have_cam) return fail(h, PT_ERR_INVALID_ARGUMENT, "no camera set");
    PT_CUDA(h, cudaSetDevice(h->device));
    float *d_events = nullptr, *d_col = nullptr;
    int *d_n = nullptr;
    PT_CUDA(h, cudaMalloc(&d_events, sizeof(float) * 16 * (size_t)max_events));
    PT_CUDA(h, cudaMalloc(&d_col, sizeof(float) * 3));
    PT_CUDA(h, cudaMalloc(&d_n, sizeof(int)));
    RenderParams rp{};
    rp.scene = h->dscene;
    rp.cam = h->cam;
    rp.width = width; rp.height = height; rp.spp = h->spp; rp.depth = h->depth;
    rp.counters = h->d_counters;
    const bool S = h->blob.has_spheres, R = h->blob.has_rtow;
    // the walk follows the handle's options: closest_hit by default, the wavefront kernel's steps when a node format was chosen
    const int mode = h->trace_steps ? (use_quantised(h) ? 2 : 1) : 0;
#define PT_TRACE(SS, RR)                                                                                         \
    do {                                                                                                         \
        if (mode == 0) pt_trace_kernel<SS, RR, 0><<<1, 1>>>(rp, x, y, d_events, max_events, d_n, d_col);         \
        else if (mode == 1) pt_trace_kernel<SS, RR, 1><<<1, 1>>>(rp, x, y, d_events, max_events, d_n, d_col);    \
        else pt_trace_kernel<SS, RR, 2><<<1, 1>>>(rp, x, y, d_events, max_events, d_n, d_col);                   \
    } while (0)
    if (!S && !R) PT_TRACE(false, false);
    else if (!S && R) PT_TRACE(false, true);
    else if (S && !R) PT_TRACE(true, false);
    else PT_TRACE(true, true);
#undef PT_TRACE
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(n_events, d_n, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(col, d_col, sizeof(float) * 3, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(events, d_events, sizeof(float) * 16 * (size_t)std::min(*n_events, max_events), cudaMemcpyDeviceToHost);
    cudaFree(d_events); cudaFree(d_col); cudaFree(d_n);
    if (e != cudaSuccess) return fail(h, (int)e, std::string("ptcore_debug_trace_pixel: ") + cudaGetErrorString(e));
    return PT_OK;
}

int pt_bvh_selftest(const PtSceneDesc *sc, int32_t leaf_max, PtStats *out, char *msg, size_t msg_len) {
    if (!sc || sc->n_tris < 0 || sc->n_spheres < 0) return PT_ERR_INVALID_ARGUMENT;
    std::vector<PrimBounds> pb = padded_bounds(sc);
    BvhBuildOptions opt;
    opt.leaf_max = leaf_max > 0 ? leaf_max : 4;
    BvhBuildResult bvh = build_bvh(pb, opt);
    const char *why = validate_bvh(bvh, pb);
    // a very deep two-wide tree can collapse into a four-wide one that needs more stack than the device has (three pushes per
    // level): ptcore_upload_scene then simply does not offer the four-wide walk, so it is not an error here either
    const bool wide_ok = bvh.stack4 <= (uint32_t)kStackSize - 2;
    if (why[0] == 0 && wide_ok) why = validate_bvh4(bvh, pb);
    double inflation = 1.0;
    if (why[0] == 0) {
        std::vector<QuantNode> nq;
        QuantGrid grid = quantise_nodes(bvh.nodes, nq);
        why = validate_quantised(bvh.nodes, nq, grid, &inflation);
    }
    if (msg && msg_len) snprintf(msg, msg_len, "%s", why);
    if (out) {
        memset(out, 0, sizeof *out);
        out->bvh_nodes = (uint32_t)bvh.nodes.size();
        out->bvh_leaves = bvh.n_leaves;
        out->bvh_depth = bvh.depth;
        out->bvh4_nodes = wide_ok ? (uint32_t)bvh.nodes4.size() : 0;
        out->bvh4_depth = wide_ok ? bvh.depth4 : 0;
        out->bvh_build_ms = bvh.build_ms;
        out->sah_cost = bvh.sah_cost;
        out->scene_bytes = bvh.nodes.size() * sizeof(FlatNode);
        out->quant_inflation = inflation;
    }
    int max_leaf = 0;
    for (const FlatNode &n : bvh.nodes)
        for (int32_t ref : {n.left, n.right})
            if (ref < 0) max_leaf = std::max(max_leaf, (int)leaf_count(ref));
    if (why[0] == 0 && !pb.empty() && max_leaf > std::max(1, std::min((int)opt.leaf_max, kMaxLeafPrims))) {
        if (msg && msg_len) snprintf(msg, msg_len, "leaf with %d primitives exceeds leaf_max", max_leaf);
        return PT_ERR_SYSTEM;
    }
    return why[0] == 0 ? PT_OK : PT_ERR_SYSTEM;
}

int pt_quant_selftest(const PtSceneDesc *sc, uint32_t n_rays, uint32_t seed, uint64_t counts[3], char *msg, size_t msg_len) {
    if (!sc || !counts || sc->n_tris < 0 || sc->n_spheres < 0) return PT_ERR_INVALID_ARGUMENT;
    std::vector<PrimBounds> pb = padded_bounds(sc);
    BvhBuildOptions opt;
    BvhBuildResult bvh = build_bvh(pb, opt);
    std::vector<QuantNode> nq;
    QuantGrid grid = quantise_nodes(bvh.nodes, nq);
    const char *why = validate_quantised(bvh.nodes, nq, grid, nullptr);
    if (why[0] == 0) why = check_quantised_walk(bvh.nodes, nq, grid, n_rays, seed, 4096, counts);
    if (msg && msg_len) snprintf(msg, msg_len, "%s", why);
    return why[0] == 0 ? PT_OK : PT_ERR_SYSTEM;
}

int pt_walk_selftest(const PtSceneDesc *sc, uint32_t n_rays, uint32_t seed, uint64_t counts[4], char *msg, size_t msg_len) {
    if (!sc || !counts || sc->n_tris <= 0 || sc->n_spheres != 0 || !sc->tri_pos) return PT_ERR_INVALID_ARGUMENT;  // triangle soups only
    std::vector<PrimBounds> pb = padded_bounds(sc);
    BvhBuildOptions opt;
    BvhBuildResult bvh = build_bvh(pb, opt);
    std::vector<QuantNode> nq;
    QuantGrid grid = quantise_nodes(bvh.nodes, nq);
    const char *why = validate_bvh(bvh, pb);
    if (why[0] == 0) why = check_walks(bvh, nq, grid, sc->tri_pos, (size_t)sc->n_tris, n_rays, seed, counts);
    if (msg && msg_len) snprintf(msg, msg_len, "%s", why);
    return why[0] == 0 ? PT_OK : PT_ERR_SYSTEM;
}

int pt_write_ppm(const char *path, const uint8_t *rgb, uint32_t width, uint32_t height) {
    if (!path || !rgb) return PT_ERR_INVALID_ARGUMENT;
    FILE *f = fopen(path, "wb");
    if (!f) return PT_ERR_SYSTEM;
    fprintf(f, "P6\n%u %u\n255\n", width, height);
    size_t n = (size_t)width * height * 3;
    size_t w = fwrite(rgb, 1, n, f);
    fclose(f);
    return w == n ? PT_OK : PT_ERR_SYSTEM;
}

}  // extern "C"

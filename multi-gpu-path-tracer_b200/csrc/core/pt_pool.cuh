// pt_pool.cuh — pt_pool_kernel: the persistent-thread wavefront path tracer with a per-warp POOL of pixels (sm_100a).
//
// What it replaces: the reference's `render` kernel (src/DevicePathTracer.h:73-120, one thread per pixel, the whole
// ray_color of src/camera.h:49-83 inlined under a virtual-dispatch material/pdf hierarchy).  What it improves on:
// pt_wavefront_kernel (pt_kernels.cuh), where a lane owns ONE pixel and therefore idles from the moment its ray finishes
// until 24 of the 32 rays of its warp have finished (ncu: 17.3 of 32 threads per executed instruction).
//
// Here a warp owns up to kPoolSlots pixels (their XORWOW streams must advance sequentially, SURVEY 0.7, so the unit that is
// scheduled is still the pixel), decoupled from its 32 lanes, and runs the four stages of the north-star as
//
//   GENERATE   free slots fetch a pixel (one warp-aggregated atomic on the launch's work counter), finished samples draw
//              the next camera ray (src/DevicePathTracer.h:84-86); every new ray goes into the warp's READY-RAY RING in
//              shared memory (origin, direction, slot);
//   TRAVERSE   lanes pull rays from the ring and walk the tree with the voted uniform steps of pt_device.cuh (node step /
//              leaf step, short stack: the first 8 entries per lane in shared memory); a lane whose ray is finished writes
//              the hit into the pixel's record, files the slot under the SHADE CLASS of what it hit and pulls the next ray
//              — this is the compaction: rays move to lanes, no lane waits for its neighbours;
//   SHADE      runs when a class has 32 hits waiting (or lanes are starving): ONE class per pass, so the 32 lanes execute
//              the same branch of camera::ray_color — terminal hits (miss / emitter: add the contribution, next sample) or
//              bounces of one material kind (UniversalMaterial / lambertian: src/material.h:52-91,110-127; metal :130-144;
//              dielectric :146-179) with the light / cosine mixture sampling of src/pdf.h:57-75.  This is the
//              material-sorted shade stage;
//   COMPACT    is implicit in the ring: a shade pass emits its continuing rays densely.
//
// Pixel records (RNG state, colour sum, throughput, current ray, hit: 7 x 16 bytes, one 128-byte line per pixel) live in
// global memory and are only touched by dense, line-aligned 128-bit accesses that bypass L1 (they are L2-resident: at most
// 58 MB per GPU); shared memory holds what the divergent part of the loop touches: the ray ring, the slot states, the
// traversal stacks.  Results do not depend on any of this scheduling: each pixel's draws happen in the reference's order
// and closest_hit is order-independent (ties go to the lower leaf position, pt_device.cuh).
#pragma once

#include "pt_kernels.cuh"

namespace ptc {

constexpr int kPoolSlots = 96;    // pixel slots per warp (upper bound; RenderParams.pool_size of them are used)
constexpr int kPoolQueue = 32;    // ready rays per warp: one shade pass fills it, the lanes drain it
constexpr int kPoolStackK = 7;    // traversal-stack entries 0..6 of a lane live in shared memory (entry 0 is the sentinel)
#ifndef PT_POOL_MINB
#define PT_POOL_MINB 4  /* resident CTAs per SM the kernel is compiled for: 4 x 256 threads = 64 registers per thread */
#endif
constexpr int kPoolThreads = 256;
constexpr int kPoolWarps = kPoolThreads / 32;
constexpr int kPoolRecQuads = 8;  // 128 bytes per pixel record

enum : uint8_t { kSlotFree = 0, kSlotQueued = 1, kSlotDone = 4 /* + shade class 0..3 */ };

// 1792 bytes per warp: with 32 resident warps per SM the kernel fits the 64 KB shared-memory configuration, which leaves
// 192 KB of L1 for the tree (the 100 / 132 KB configurations cost 25 points of L1 hit rate: profiles/r02_ncu_pool_v0.txt)
struct PoolWarpSmem {
    float q_inv[3][kPoolQueue];   // ready rays, prepared for the slab tests (trav_prepare): reciprocal direction ...
    float q_oinv[3][kPoolQueue];  // ... and constant term; origin and direction stay in the pixel record
    int32_t stack[kPoolStackK][32];
    uint8_t state[kPoolSlots];
    uint8_t q_slot[kPoolQueue];   // slot of each ready ray; doubles as the batch list while a shade pass is being assembled
};

// pixel record, 8 x float4 (one 128-byte line):
//   0: rng.d v0 v1 v2      1: rng.v3 v4, px | py << 16, samples done      2: att.xyz, bounce
//   3: o.xyz, col.x        4: d.xyz, col.y                                5: hit t u v prim
//   6: col.z, rays of this pixel (pilot pass), -, -                       7: unused
template <bool SPHERES, bool RTOW, bool COUNT, int NODES>
__global__ void __launch_bounds__(kPoolThreads, PT_POOL_MINB) pt_pool_kernel(const __grid_constant__ RenderParams p) {
    constexpr bool QUANT = NODES == 2;
    constexpr int NC = RTOW ? 4 : 2;
    __shared__ PoolWarpSmem smem[kPoolWarps];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    PoolWarpSmem &sm = smem[threadIdx.x >> 5];
    const int pool = p.pool_size;
    float4 *const slots = p.pool_slots + ((size_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5)) * (size_t)pool * kPoolRecQuads;
    const uint32_t total_items = work_total(p);
    const int period_mask = p.pool_period - 1;

    for (int s = (int)lane; s < kPoolSlots; s += 32) sm.state[s] = kSlotFree;
    __syncwarp();

    // warp-uniform bookkeeping
    int q_next = 0, q_count = 0, n_live = 0;  // ready rays are q_next .. q_count-1
    int n_done[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) n_done[c] = 0;
    bool exhausted = false;
    uint32_t watchdog = 0;

    // lane state of the TRAVERSE stage
    int my_slot = -1;
    float3 ro = f3(0.f, 0.f, 0.f), rd = f3(0.f, 0.f, 1.f);
    int32_t overflow[kStackSize - kPoolStackK];
    const ShortStack<kPoolStackK> stack{(uint32_t)__cvta_generic_to_shared(&sm.stack[0][lane]), overflow};
    Trav tr;
    trav_start<QUANT>(tr, stack, ro, rd);
    trav_idle(tr);
    uint32_t n_rays = 0, n_box = 0, n_tri = 0, n_light = 0;
    unsigned long long acc_box = 0, acc_tri = 0, acc_light = 0;

    for (;;) {
        // =============================== SHADE + GENERATE: one batch of <= 32 slots ===============================
        // (the ray ring is empty here: q_next == q_count)
        int cls = 0, most = n_done[0];
#pragma unroll
        for (int c = 1; c < NC; c++)
            if (n_done[c] > most) { most = n_done[c]; cls = c; }
        int n_batch_done = 0;
        if (most > 0) {
            const int want = min(most, 32);
            for (int base = 0; base < kPoolSlots; base += 32) {
                const bool is = sm.state[base + (int)lane] == (uint8_t)(kSlotDone + cls);
                const unsigned m = __ballot_sync(kFullMask, is);
                const int r = n_batch_done + __popc(m & lt_mask);
                if (is && r < want) sm.q_slot[r] = (uint8_t)(base + (int)lane);
                n_batch_done = min(want, n_batch_done + __popc(m));
            }
        }
        int n_batch = n_batch_done;
        if (!exhausted) {  // free slots ask for a new pixel
            const int want_new = min(32 - n_batch_done, pool - n_live);
            if (want_new > 0) {
                int got = 0;
                for (int base = 0; base < kPoolSlots; base += 32) {
                    const int s = base + (int)lane;
                    const bool is = s < pool && sm.state[s] == kSlotFree;
                    const unsigned m = __ballot_sync(kFullMask, is);
                    const int r = got + __popc(m & lt_mask);
                    if (is && r < want_new) sm.q_slot[n_batch_done + r] = (uint8_t)s;
                    got = min(want_new, got + __popc(m));
                }
                n_batch += got;
            }
        }
        __syncwarp();
        q_next = q_count = 0;
        if (n_batch > 0) {
            const bool has = (int)lane < n_batch;
            const bool is_done = (int)lane < n_batch_done;
            const int slot = has ? (int)sm.q_slot[lane] : 0;
            __syncwarp();  // q_slot is about to be rewritten as the ray list
            float4 *const rec = slots + (size_t)slot * kPoolRecQuads;
            Rng rng;
            rng_init(rng, 0);
            float3 col = f3(0.f, 0.f, 0.f), att = f3(1.f, 1.f, 1.f), o = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 1.f);
            uint32_t pix = 0, samples_done = 0, bounce = 0, pixel_rays = 0;
            bool want_pixel = has && !is_done, want_cam = false, have_ray = false, freed = false;
            if (is_done) {
                const float4 r0 = __ldcg(rec + 0), r1 = __ldcg(rec + 1), r2 = __ldcg(rec + 2), r3 = __ldcg(rec + 3), r4 = __ldcg(rec + 4), r5 = __ldcg(rec + 5),
                             r6 = __ldcg(rec + 6);
                rng.d = __float_as_uint(r0.x); rng.v0 = __float_as_uint(r0.y); rng.v1 = __float_as_uint(r0.z); rng.v2 = __float_as_uint(r0.w);
                rng.v3 = __float_as_uint(r1.x); rng.v4 = __float_as_uint(r1.y);
                pix = __float_as_uint(r1.z);
                samples_done = __float_as_uint(r1.w);
                att = f3(r2.x, r2.y, r2.z);
                bounce = __float_as_uint(r2.w);
                o = f3(r3.x, r3.y, r3.z);
                d = f3(r4.x, r4.y, r4.z);
                col = f3(r3.w, r4.w, r6.x);
                pixel_rays = __float_as_uint(r6.y);
                Hit h;
                h.t = r5.x; h.u = r5.y; h.v = r5.z; h.prim = __float_as_int(r5.w);
                float3 contrib;
                const bool cont = shade<SPHERES, RTOW, COUNT>(p.scene, h, o, d, att, rng, contrib, n_light);
                bounce++;
                if (cont && bounce < p.depth) {
                    have_ray = true;
                } else {
                    col = col + (cont ? f3(0.0f, 0.0f, 0.0f) : contrib);  // camera.h:82: recursion exhausted -> (0,0,0)
                    samples_done++;
                    if (samples_done == p.spp) {
                        const int px = (int)(pix & 0xffffu), py = (int)(pix >> 16);
                        finish_pixel(p, px, py, ((int)p.height - py - 1) * (int)p.width + px, col, pixel_rays);
                        want_pixel = true;
                    } else {
                        want_cam = true;
                    }
                }
            }
            // new pixels: one atomic per warp on the launch's work counter
            const unsigned m_need = __ballot_sync(kFullMask, want_pixel);
            if (m_need) {
                uint32_t first = total_items;
                if (!exhausted) {
                    const int leader = __ffs((int)m_need) - 1;
                    if ((int)lane == leader) first = atomicAdd(p.work_counter, (uint32_t)__popc(m_need));
                    first = __shfl_sync(kFullMask, first, leader);
                    if (first + (uint32_t)__popc(m_need) >= total_items) exhausted = true;
                }
                bool got_pixel = false;
                if (want_pixel) {
                    const uint32_t item = first + (uint32_t)__popc(m_need & lt_mask);
                    int px, py;
                    if (item < total_items && work_to_pixel(p, item, px, py)) {
                        const int pixel_index = ((int)p.height - py - 1) * (int)p.width + px;
                        pix = (uint32_t)px | ((uint32_t)py << 16);
                        rng_init(rng, (unsigned long long)(long long)(1984 + pixel_index));  // DevicePathTracer.h:54
                        col = f3(0.f, 0.f, 0.f);
                        samples_done = 0;
                        pixel_rays = 0;
                        want_cam = true;
                        got_pixel = true;
                    } else {
                        freed = is_done;  // the slot's pixel is finished and there is none to replace it
                    }
                }
                n_live += __popc(__ballot_sync(kFullMask, got_pixel && !is_done)) - __popc(__ballot_sync(kFullMask, freed));
            }
            if (want_cam) {  // DevicePathTracer.h:84-87
                const int px = (int)(pix & 0xffffu), py = (int)(pix >> 16);
                const float u = float(px + rng_uniform(rng)) / float(p.width);
                const float v = float(py + rng_uniform(rng)) / float(p.height);
                camera_ray(p.cam, u, v, o, d);
                att = f3(1.0f, 1.0f, 1.0f);
                bounce = 0;
                have_ray = true;
            }
            const unsigned m_ray = __ballot_sync(kFullMask, have_ray);
            if (have_ray) {
                n_rays++;
                pixel_rays++;
                __stcg(rec + 0, make_float4(__uint_as_float(rng.d), __uint_as_float(rng.v0), __uint_as_float(rng.v1), __uint_as_float(rng.v2)));
                __stcg(rec + 1, make_float4(__uint_as_float(rng.v3), __uint_as_float(rng.v4), __uint_as_float(pix), __uint_as_float(samples_done)));
                __stcg(rec + 2, make_float4(att.x, att.y, att.z, __uint_as_float(bounce)));
                __stcg(rec + 3, make_float4(o.x, o.y, o.z, col.x));
                __stcg(rec + 4, make_float4(d.x, d.y, d.z, col.y));
                __stcg(rec + 6, make_float4(col.z, __uint_as_float(pixel_rays), 0.f, 0.f));
                float3 inv, oinv;
                if (QUANT) trav_prepare_grid(p.scene, o, d, inv, oinv);
                else trav_prepare(o, d, inv, oinv);
                const int e = __popc(m_ray & lt_mask);
                sm.q_inv[0][e] = inv.x; sm.q_inv[1][e] = inv.y; sm.q_inv[2][e] = inv.z;
                sm.q_oinv[0][e] = oinv.x; sm.q_oinv[1][e] = oinv.y; sm.q_oinv[2][e] = oinv.z;
                sm.q_slot[e] = (uint8_t)slot;
                sm.state[slot] = kSlotQueued;
            } else if (freed) {
                sm.state[slot] = kSlotFree;
            }
            q_count = __popc(m_ray);
#pragma unroll
            for (int c = 0; c < NC; c++)
                if (c == cls) n_done[c] -= n_batch_done;
            __syncwarp();
        }
        if (COUNT) {
            acc_light += n_light;
            n_light = 0;
        }
        if (n_live == 0 && exhausted) break;

        // =============================== TRAVERSE: voted uniform steps, rays pulled from the ring ===============================
        bool housekeeping = true;
        for (int it = 0;; it++) {
            if (p.watchdog && ++watchdog > p.watchdog) {  // debugging aid of the host: a wrong schedule must not hang the GPU
                n_live = 0;
                exhausted = true;
                break;
            }
            if (housekeeping || (it & period_mask) == 0) {
                // ---- lanes whose ray is finished hand the hit to the shade stage ----
                const bool fin = my_slot >= 0 && trav_finished(tr);
                const unsigned m_fin = __ballot_sync(kFullMask, fin);
                if (m_fin) {
                    int c = 0;
                    if (fin) {
                        if (tr.best.prim >= 0) c = prim_shade_class(p.scene, tr.best.prim);  // shade class of the primitive (ptcore_upload_scene)
                        if (!RTOW) c &= 1;
                        __stcg(slots + (size_t)my_slot * kPoolRecQuads + 5, make_float4(tr.best.t, tr.best.u, tr.best.v, __int_as_float(tr.best.prim)));
                        sm.state[my_slot] = (uint8_t)(kSlotDone + c);
                        my_slot = -1;
                    }
#pragma unroll
                    for (int cc = 0; cc < NC; cc++) n_done[cc] += __popc(__ballot_sync(kFullMask, fin && c == cc));
                }
                const unsigned m_idle = __ballot_sync(kFullMask, my_slot < 0);
                if (m_idle) {
                    if (q_next < q_count) {
                        // ---- idle lanes pull the next ready rays ----
                        const int e = q_next + __popc(m_idle & lt_mask);
                        if (my_slot < 0 && e < q_count) {
                            my_slot = (int)sm.q_slot[e];
                            const float4 *rec = slots + (size_t)my_slot * kPoolRecQuads;
                            const float4 r3 = __ldcg(rec + 3), r4 = __ldcg(rec + 4);  // first needed by the first leaf step
                            ro = f3(r3.x, r3.y, r3.z);
                            rd = f3(r4.x, r4.y, r4.z);
                            trav_start<QUANT>(tr, stack, f3(sm.q_inv[0][e], sm.q_inv[1][e], sm.q_inv[2][e]), f3(sm.q_oinv[0][e], sm.q_oinv[1][e], sm.q_oinv[2][e]));
                        }
                        q_next = min(q_count, q_next + __popc(m_idle));
                    } else {
                        // ---- nothing to pull: shade once enough hits are waiting (a warp with few pixels left shades at once, so
                        //      its pixels advance at the speed of their own rays) ----
                        int waiting = 0;
#pragma unroll
                        for (int cc = 0; cc < NC; cc++) waiting += n_done[cc];
                        if (waiting >= min(p.pool_idle_at, max(1, n_live >> 2)) || m_idle == kFullMask) break;
                    }
                }
                housekeeping = false;
            }
            const bool can_node = tr.cur >= 0;
            const bool can_prim = trav_leaf_held(tr);
            const unsigned m_node = __ballot_sync(kFullMask, can_node);
            const unsigned m_prim = __ballot_sync(kFullMask, can_prim);
            if ((m_node | m_prim) == 0) {  // every ray in flight is finished: retire them now
                housekeeping = true;
                continue;
            }
            if (__popc(m_node) >= __popc(m_prim)) {
                if (can_node) trav_node_step<COUNT, QUANT>(p.scene, tr, stack, 0.001f, n_box);
                if (tr.cur >= 0) trav_node_step<COUNT, QUANT>(p.scene, tr, stack, 0.001f, n_box);
            } else if (can_prim) {
                if (SPHERES) trav_prim_step<SPHERES, COUNT>(p.scene, tr, stack, ro, rd, 0.001f, n_tri);
                else trav_prim_step2<COUNT>(p.scene, tr, stack, ro, rd, 0.001f, n_tri);
            }
        }
        if (COUNT) {
            acc_box += n_box; acc_tri += n_tri;
            n_box = n_tri = 0;
        }
        __syncwarp();
    }

    unsigned long long r = n_rays;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(kFullMask, r, o);
    if (lane == 0 && r) atomicAdd(&p.counters->rays, r);
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            acc_box += __shfl_down_sync(kFullMask, acc_box, o);
            acc_tri += __shfl_down_sync(kFullMask, acc_tri, o);
            acc_light += __shfl_down_sync(kFullMask, acc_light, o);
        }
        if (lane == 0) {
            atomicAdd(&p.counters->box_tests, acc_box);
            atomicAdd(&p.counters->tri_tests, acc_tri);
            atomicAdd(&p.counters->light_tests, acc_light);
        }
    }
}

}  // namespace ptc

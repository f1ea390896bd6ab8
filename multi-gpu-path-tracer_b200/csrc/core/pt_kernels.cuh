// pt_kernels.cuh — the render kernels (sm_100a).
//
// pt_persistent_kernel replaces the reference's `render` kernel (src/DevicePathTracer.h:73-120:
// one thread per pixel, spp loop, virtual-dispatch ray_color inlined, 103 registers + 672 B stack).
// Organisation here: a persistent grid (SMs x resident CTAs) whose lanes each own one pixel's
// XORWOW stream and run a per-lane wavefront
//        GENERATE (camera ray / next sample / next pixel)  ->  TRAVERSE  ->  SHADE  -> (COMPACT)
// in warp lock-step; a lane whose pixel has finished its spp refills itself from a global work
// counter with one warp-aggregated atomic (the "compact" stage: idle lanes are replaced, live
// ones keep their registers), so every lane carries a ray into every TRAVERSE stage until the
// launch runs out of pixels.  Exact RNG-stream reproduction pins "one in-flight path per pixel,
// samples sequential" (SURVEY §0.7), which is why the unit of refill is a pixel, not a ray.
//
// pt_direct_kernel is the plain parity slice (one thread per pixel, no refill), kept as the
// A/B baseline for the scheduling choice.
#pragma once

#include "pt_device.cuh"

#include <type_traits>

namespace ptc {

constexpr int kBlockThreads = 128;
constexpr unsigned kFullMask = 0xffffffffu;

// PT_RNG_SAMPLE_KEYED: seed of sample s of pixel p = the counter 1984 + p + s * W * H passed through the splitmix64 finaliser, so that
// the 64 bits from which curand_init(seed, 0, 0) derives the whole XORWOW state (curand_kernel.h:772-797) are well mixed although
// consecutive counters differ in a few low bits only (the CPU checker restates the same function).
__device__ __forceinline__ unsigned long long keyed_seed(uint32_t pixel_index, uint32_t sample, uint32_t npix) {
    unsigned long long z = 1984ull + (unsigned long long)pixel_index + (unsigned long long)sample * (unsigned long long)npix;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// work item -> pixel.  Items enumerate 8x4 pixel blocks of each tile (32 consecutive items = one
// block = one warp's initial fetch), so warps start on spatially coherent rays.
__device__ __forceinline__ bool item_to_pixel(const TileList &tl, uint32_t item, int &x, int &y) {
    int t = 0;
    // tiles are few (<= kMaxInlineTiles): linear scan over the prefix sums
    while (t + 1 < tl.n && item >= tl.first_item[t + 1]) t++;
    uint32_t local = item - tl.first_item[t];
    uint32_t blk = local >> 5, within = local & 31u;
    uint32_t blocks_per_row = ((uint32_t)tl.w[t] + 7u) >> 3;
    uint32_t by = blk / blocks_per_row, bx = blk - by * blocks_per_row;
    int lx = (int)(bx * 8u + (within & 7u));
    int ly = (int)(by * 4u + (within >> 3));
    x = tl.ox[t] + lx;
    y = tl.oy[t] + ly;
    return lx < tl.w[t] && ly < tl.h[t];
}

// work item -> pixel for either kind of work list
__device__ __forceinline__ uint32_t work_total(const RenderParams &p) { return p.block_list ? p.n_blocks * 32u : p.tiles.first_item[p.tiles.n]; }
__device__ __forceinline__ bool work_to_pixel(const RenderParams &p, uint32_t item, int &x, int &y) {
    if (p.block_list) {
        const uint32_t b = __ldg(&p.block_list[item >> 5]), within = item & 31u;
        x = (int)((b & 0xffffu) * 8u + (within & 7u));
        y = (int)((b >> 16) * 4u + (within >> 3));
        return x < (int)p.width && y < (int)p.height;
    }
    return item_to_pixel(p.tiles, item, x, y);
}
// a finished pixel: normally the quantised store; in the pilot pass its ray count goes to the block-cost map instead
__device__ __forceinline__ void finish_pixel(const RenderParams &p, int px, int py, int pixel_index, float3 col, uint32_t pixel_rays) {
    if (p.block_cost) atomicAdd(&p.block_cost[(uint32_t)(py >> 2) * ((p.width + 7u) >> 3) + (uint32_t)(px >> 3)], pixel_rays);
    else store_pixel(p, pixel_index, col);
}

template <bool SPHERES, bool RTOW, bool COUNT>
__global__ void __launch_bounds__(kBlockThreads) pt_persistent_kernel(const __grid_constant__ RenderParams p) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t total_items = p.tiles.first_item[p.tiles.n];

    // per-lane pixel state
    bool retired = false;     // no more work for this lane
    bool have_pixel = false;  // owns a pixel whose samples are not finished
    bool have_path = false;   // carries a live ray
    int px = 0, py = 0, pixel_index = 0;
    uint32_t samples_done = 0, bounce = 0;
    Rng rng;
    rng_init(rng, 0);
    float3 col = f3(0.f, 0.f, 0.f), att = f3(1.f, 1.f, 1.f), ro = f3(0.f, 0.f, 0.f), rd = f3(0.f, 0.f, 1.f);
    uint32_t n_rays = 0, n_box = 0, n_tri = 0, n_light = 0;
    unsigned long long acc_box = 0, acc_tri = 0, acc_light = 0;

    for (;;) {
        // ---------------- GENERATE / COMPACT ----------------
        if (!have_path && have_pixel && samples_done == p.spp) {
            store_pixel(p, pixel_index, col);
            have_pixel = false;
        }
        const bool need = !retired && !have_pixel;
        const unsigned need_mask = __ballot_sync(kFullMask, need);
        if (need_mask) {
            const int leader = __ffs((int)need_mask) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(p.work_counter, (uint32_t)__popc(need_mask));
            base = __shfl_sync(kFullMask, base, leader);
            if (need) {
                const uint32_t item = base + (uint32_t)__popc(need_mask & ((1u << lane) - 1u));
                if (item >= total_items) {
                    retired = true;
                } else if (item_to_pixel(p.tiles, item, px, py)) {
                    // DevicePathTracer.h:79: framebuffer row 0 is the top image row; :54: seed = 1984 + pixel_index
                    pixel_index = ((int)p.height - py - 1) * (int)p.width + px;
                    rng_init(rng, (unsigned long long)(long long)(1984 + pixel_index));
                    col = f3(0.f, 0.f, 0.f);
                    samples_done = 0;
                    have_pixel = true;
                }
            }
        }
        if (__all_sync(kFullMask, retired)) break;

        if (!have_path && have_pixel && samples_done < p.spp) {
            // DevicePathTracer.h:84-87
            float u = float(px + rng_uniform(rng)) / float(p.width);
            float v = float(py + rng_uniform(rng)) / float(p.height);
            camera_ray(p.cam, u, v, ro, rd);
            att = f3(1.0f, 1.0f, 1.0f);
            bounce = 0;
            have_path = true;
            if (p.depth == 0) {  // camera.h:52,82: the loop body never runs
                have_path = false;
                samples_done++;
            }
        }

        // ---------------- TRAVERSE ----------------
        Hit h;
        h.prim = -1;
        h.t = FLT_MAX;
        h.u = h.v = 0.f;
        if (have_path) {
            h = closest_hit<SPHERES, COUNT>(p.scene, ro, rd, 0.001f, n_box, n_tri);
            n_rays++;
        }

        // ---------------- SHADE ----------------
        if (have_path) {
            float3 contrib;
            bool cont = shade<SPHERES, RTOW, COUNT>(p.scene, h, ro, rd, att, rng, contrib, n_light);
            bounce++;
            if (!cont) {
                col = col + contrib;  // DevicePathTracer.h:87
                have_path = false;
                samples_done++;
            } else if (bounce >= p.depth) {
                // camera.h:82: recursion exhausted -> (0,0,0); col += 0 keeps NaN/inf behaviour identical
                col = col + f3(0.0f, 0.0f, 0.0f);
                have_path = false;
                samples_done++;
            }
        }
        if (COUNT) {
            acc_box += n_box; acc_tri += n_tri; acc_light += n_light;
            n_box = n_tri = n_light = 0;
        }
    }

    // counters: one atomic per warp
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

// pt_wavefront_kernel — the default.  Same per-lane pixel ownership and refill as pt_persistent_kernel, but the
// TRAVERSE stage is a warp-synchronous wavefront over uniform steps (pt_device.cuh: trav_node_step /
// trav_prim_step): every iteration the warp votes and executes ONE kind of step for all lanes that can take it,
//       #lanes at an inner node  >=  #lanes holding a leaf primitive   ->  node step      else  primitive step,
// so box tests and Moeller-Trumbore tests each run with most of the warp instead of a handful of lanes
// (ncu on pt_persistent_kernel: 7.0 threads per executed instruction, 3.2 inside the triangle test;
// profiles/r01_ncu_persistent_lockstep.txt).  Lanes whose ray is finished wait; once `refill_at` of them are
// waiting the warp leaves TRAVERSE, shades exactly those lanes (they get their bounce ray or the next camera
// ray) and re-enters with the unfinished lanes resuming where they stopped — warp-level ray compaction without
// moving any state between lanes, which the one-XORWOW-stream-per-pixel contract forbids.
// KEYED (PT_OPT_RNG_MODE = PT_RNG_SAMPLE_KEYED): the XORWOW stream is keyed by (pixel, sample) instead of by pixel — sample s of
// pixel p draws from curand_init(splitmix64(1984 + p + s * W * H), 0, 0) — so the samples of a pixel no longer form one sequential chain.  A
// work item is then one CHUNK of a pixel's samples; its partial colour sum goes to accum[chunk][pixel] and pt_resolve_keyed_kernel
// adds the chunks in order.  Same integrand, same estimator, different random numbers: parity with the reference is statistical
// in this mode (converged RMSE), which is why it is never the default.
// SSTACK: the first kWfStackK entries of every lane's traversal stack in shared memory (ShortStack, pt_device.cuh) instead of local
// memory.  Local memory (stack + spills) is the largest consumer of L1 sectors in the capture of the shipped kernel (100 G of 152 G per
// 1024-spp frame, profiles/r02_ncu_wavefront_1080p_1024spp.txt), so this looked like the next step after the shared-memory nodes —
// measured on B200 it LOSES 12 % (2837 -> 2491 Msamples/s at 128 spp, 2869 -> 2524 at 1024 spp; same pixels): 32 KB more shared memory
// move the carve-out from 100 to 132 KB and the predicated LDS / STS plus the overflow branch cost more than the L1-resident LDL / STL they
// replace.  Compiled out (kWfSmemStack = false); the pool kernel uses the same ShortStack.
constexpr int kWfStackK = 8;
constexpr bool kWfSmemStack = false;
template <bool SPHERES, bool RTOW, bool COUNT, int NODES, bool KEYED = false, bool SSTACK = false>  // NODES: 0 = 64-byte two-child nodes, 1 = four-wide nodes, 2 = 32-byte quantised two-child nodes, 3 = the same in shared memory
__device__ __forceinline__ void wavefront_body(const RenderParams &p, const uint32_t smem_nodes, const uint32_t smem_stack = 0) {
    constexpr bool WIDE = NODES == 1, QUANT = NODES >= 2, SMEM = NODES == 3;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t total_items = KEYED ? work_total(p) * p.keyed_my_chunks : work_total(p);
    uint32_t sample_end = p.spp;  // KEYED: first sample after this lane's chunk
    const int refill_at = p.refill_at;
    const int node_burst = p.node_burst;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p.retire_log && lane == 0 && warp_global < p.retire_log_warps) p.retire_log[2 * warp_global] = global_timer_ns();

    bool retired = (int)lane >= p.lanes_per_warp, have_pixel = false, have_path = false;
    int px = 0, py = 0, pixel_index = 0;
    uint32_t samples_done = 0, bounce = 0;
    Rng rng;
    rng_init(rng, 0);
    float3 col = f3(0.f, 0.f, 0.f), att = f3(1.f, 1.f, 1.f), ro = f3(0.f, 0.f, 0.f), rd = f3(0.f, 0.f, 1.f);
    uint32_t n_rays = 0, n_box = 0, n_tri = 0, n_light = 0, pixel_rays = 0;
    unsigned long long acc_box = 0, acc_tri = 0, acc_light = 0;
    int32_t stack_mem[SSTACK ? kStackSize - kWfStackK : kStackSize];
    typedef typename std::conditional<SSTACK, ShortStack<kWfStackK>, LocalStack>::type StackT;
    const StackT stack = make_stack<StackT>(stack_mem, smem_stack + lane * 4u);
    Trav tr;
    if (QUANT) trav_begin_grid(tr, stack, p.scene, ro, rd);
    else trav_begin(tr, stack, ro, rd);
    trav_idle(tr);

    for (;;) {
        // ---------------- GENERATE / COMPACT ----------------
        if (!have_path && have_pixel && samples_done == (KEYED ? sample_end : p.spp)) {
            if (KEYED) {
                const uint32_t chunk = (sample_end - 1u) / p.keyed_chunk_spp;
                float *a = p.keyed_accum + ((size_t)chunk * (size_t)(p.width * p.height) + (size_t)pixel_index) * 3;
                a[0] = col.x; a[1] = col.y; a[2] = col.z;
            } else {
                finish_pixel(p, px, py, pixel_index, col, pixel_rays);
            }
            have_pixel = false;
        }
        const bool need = !retired && !have_pixel;
        const unsigned need_mask = __ballot_sync(kFullMask, need);
        if (need_mask) {
            const int leader = __ffs((int)need_mask) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(p.work_counter, (uint32_t)__popc(need_mask));
            base = __shfl_sync(kFullMask, base, leader);
            if (need) {
                const uint32_t item = base + (uint32_t)__popc(need_mask & ((1u << lane) - 1u));
                uint32_t pitem = item, chunk = 0;
                if (KEYED) {  // 32 consecutive items = the 32 pixels of one block in ONE chunk; a block's chunks follow each other
                    const uint32_t bi = item >> 5, k = bi % p.keyed_my_chunks;
                    pitem = ((bi / p.keyed_my_chunks) << 5) | (item & 31u);
                    chunk = p.keyed_first + k * p.keyed_step;
                }
                if (item >= total_items) {
                    retired = true;
                } else if (work_to_pixel(p, pitem, px, py)) {
                    pixel_index = ((int)p.height - py - 1) * (int)p.width + px;
                    col = f3(0.f, 0.f, 0.f);
                    if (KEYED) {
                        samples_done = chunk * p.keyed_chunk_spp;
                        sample_end = min(p.spp, samples_done + p.keyed_chunk_spp);
                        have_pixel = samples_done < sample_end;
                    } else {
                        rng_init(rng, (unsigned long long)(long long)(1984 + pixel_index));
                        samples_done = 0;
                        have_pixel = true;
                    }
                    pixel_rays = 0;
                }
            }
        }
        if (__all_sync(kFullMask, retired)) break;

        if (!have_path && have_pixel && samples_done < (KEYED ? sample_end : p.spp)) {
            if (KEYED) rng_init(rng, keyed_seed((uint32_t)pixel_index, samples_done, p.width * p.height));
            float u = float(px + rng_uniform(rng)) / float(p.width);
            float v = float(py + rng_uniform(rng)) / float(p.height);
            camera_ray(p.cam, u, v, ro, rd);
            att = f3(1.0f, 1.0f, 1.0f);
            bounce = 0;
            if (p.depth == 0) {
                samples_done++;  // camera.h:52,82: the loop body never runs
            } else {
                have_path = true;
                if (QUANT) trav_begin_grid(tr, stack, p.scene, ro, rd);
                else trav_begin(tr, stack, ro, rd);
                n_rays++;
                pixel_rays++;
            }
        }

        // ---------------- TRAVERSE (warp-synchronous wavefront over uniform steps) ----------------
        // lanes with a path are either still traversing (one of the two votes below is true) or finished and waiting
        const int n_paths = __popc(__ballot_sync(kFullMask, have_path));
        // near the end of a launch a warp has few lanes left: scale the threshold so that finished lanes never wait for
        // more lanes than exist (otherwise the survivors advance in per-ray lock-step and the tail of the frame doubles)
        const int wait_for = min(refill_at, max(1, (n_paths * 3) >> 2));
        for (;;) {
            const bool can_node = tr.cur >= 0;
            const bool can_prim = trav_leaf_held(tr);
            const unsigned m_node = __ballot_sync(kFullMask, can_node);
            const unsigned m_prim = __ballot_sync(kFullMask, can_prim);
            const int n_active = __popc(m_node | m_prim);
            if (n_active == 0 || n_paths - n_active >= wait_for) break;
            if (__popc(m_node) >= __popc(m_prim)) {
                if (!WIDE && node_burst == 2) {  // the default, without the loop bookkeeping
                    if (can_node) trav_node_step<COUNT, QUANT, StackT, SMEM>(p.scene, tr, stack, 0.001f, n_box, smem_nodes);
                    if (tr.cur >= 0) trav_node_step<COUNT, QUANT, StackT, SMEM>(p.scene, tr, stack, 0.001f, n_box, smem_nodes);
                } else {
                    for (int k = 0; k < node_burst; k++)
                        if (tr.cur >= 0) {
                            if (WIDE) trav_node_step4<COUNT>(p.scene, tr, stack, 0.001f, n_box);
                            else trav_node_step<COUNT, QUANT, StackT, SMEM>(p.scene, tr, stack, 0.001f, n_box, smem_nodes);
                        }
                }
            } else {
                if (can_prim) {
                    if (SPHERES) trav_prim_step<SPHERES, COUNT>(p.scene, tr, stack, ro, rd, 0.001f, n_tri);
                    else trav_prim_step2<COUNT>(p.scene, tr, stack, ro, rd, 0.001f, n_tri);
                }
            }
        }

        // ---------------- SHADE (only lanes whose ray is finished) ----------------
        if (have_path && trav_finished(tr)) {
            float3 contrib;
            bool cont = shade<SPHERES, RTOW, COUNT>(p.scene, tr.best, ro, rd, att, rng, contrib, n_light);
            bounce++;
            if (cont && bounce < p.depth) {
                if (QUANT) trav_begin_grid(tr, stack, p.scene, ro, rd);
                else trav_begin(tr, stack, ro, rd);
                n_rays++;
                pixel_rays++;
            } else {
                col = col + (cont ? f3(0.0f, 0.0f, 0.0f) : contrib);  // camera.h:82 exhausted -> (0,0,0)
                have_path = false;
                samples_done++;
                trav_idle(tr);
            }
        }
        if (COUNT) {
            acc_box += n_box; acc_tri += n_tri; acc_light += n_light;
            n_box = n_tri = n_light = 0;
        }
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
    if (p.retire_log && lane == 0 && warp_global < p.retire_log_warps) p.retire_log[2 * warp_global + 1] = global_timer_ns();
}

template <bool SPHERES, bool RTOW, bool COUNT, int NODES, bool KEYED = false>
__global__ void __launch_bounds__(kBlockThreads, 8) pt_wavefront_kernel(const __grid_constant__ RenderParams p) {
    wavefront_body<SPHERES, RTOW, COUNT, NODES, KEYED>(p, 0u);
}

// The same kernel as ONE 1024-thread CTA per SM whose warps share a copy of the quantised node array in shared memory (scenes
// whose nodes fit: cornell_duck has 2.1 K nodes = 67 KB).  The walk's node fetches — 13 of the 15 dependent loads of an average
// ray — then cost the shared-memory latency; through L1 a warp-wide fetch waits for L2 whenever ANY of its lanes misses, which
// at 86 % hit rate and 18 active lanes is nearly always (profiles/r01_ncu_wavefront_final_1080p_1024spp.txt: 3.0 of the 10.5
// cycles between two issues of a warp are long-scoreboard waits).
constexpr int kSmemKernelThreads = 1024;
template <bool SPHERES, bool RTOW, bool COUNT, bool KEYED = false>
__global__ void __launch_bounds__(kSmemKernelThreads, 1) pt_wavefront_smem_kernel(const __grid_constant__ RenderParams p, const int32_t n_nodes) {
    extern __shared__ uint4 smem_nodesq[];  // [2 * n_nodes] quantised nodes, then [32 warps][kWfStackK][32 lanes] stack words
    for (int i = (int)threadIdx.x; i < n_nodes * 2; i += (int)blockDim.x) smem_nodesq[i] = __ldg(&p.scene.nodesq[i]);
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_nodesq);
    const uint32_t stacks = base + (uint32_t)n_nodes * 32u + (threadIdx.x >> 5) * (uint32_t)(kWfStackK * 128);
    wavefront_body<SPHERES, RTOW, COUNT, 3, KEYED, kWfSmemStack>(p, base, stacks);
}

// PT_RNG_SAMPLE_KEYED: adds the chunk sums of every pixel in chunk order and stores the pixel (quantiser + RGB8 + I420 as always)
__global__ void pt_resolve_keyed_kernel(const __grid_constant__ RenderParams p, const uint32_t n_chunks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, npix = p.width * p.height;
    if (i >= npix) return;
    float3 col = f3(0.f, 0.f, 0.f);
    for (uint32_t c = 0; c < n_chunks; c++) {
        const float *a = p.keyed_accum + ((size_t)c * npix + i) * 3;
        col = col + f3(a[0], a[1], a[2]);
    }
    store_pixel(p, (int)i, col);
}

// Once per scene upload: the shading frame of every triangle (hit_record.normal, triangle.h:103, and the onb that
// cosine_pdf builds from it, onb.h:8-13) and the normal of every light triangle (triangle.h:36), through the same device
// functions a per-hit evaluation would inline, so shade() loads what the reference recomputes at every bounce.
__global__ void pt_frames_kernel(const DevScene sc, float4 *frames, float4 *lights) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < sc.n_prims && __float_as_int(frames[k * 4 + 1].w) == 0) {
        const PrimGeom g = load_prim(sc, k);
        const float3 normal = normalize(cross(g.e1, g.e2));
        const Onb o = make_onb(normal);
        frames[k * 4 + 0] = make_float4(normal.x, normal.y, normal.z, frames[k * 4 + 0].w);
        frames[k * 4 + 1] = make_float4(o.w.x, o.w.y, o.w.z, frames[k * 4 + 1].w);
        frames[k * 4 + 2] = make_float4(o.u.x, o.u.y, o.u.z, o.v.x);
        frames[k * 4 + 3] = make_float4(o.v.y, o.v.z, 0.f, 0.f);
    }
    if (k < sc.n_lights) {
        const float4 l0 = lights[k * 4 + 0], l1 = lights[k * 4 + 1], l2 = lights[k * 4 + 2];
        const float3 n = light_normal(f3(l0.x, l0.y, l0.z), f3(l1.x, l1.y, l1.z), f3(l2.x, l2.y, l2.z));
        lights[k * 4 + 1].w = n.x;
        lights[k * 4 + 2].w = n.y;
        lights[k * 4 + 3] = make_float4(n.z, 0.f, 0.f, 0.f);
    }
}

// One thread per pixel of ONE tile (tiles.n == 1), 8x4 blocks per warp, no refill.
template <bool SPHERES, bool RTOW, bool COUNT>
__global__ void __launch_bounds__(kBlockThreads) pt_direct_kernel(const __grid_constant__ RenderParams p) {
    const uint32_t item = blockIdx.x * (uint32_t)blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t n_rays = 0, n_box = 0, n_tri = 0, n_light = 0;
    int px, py;
    if (item < p.tiles.first_item[p.tiles.n] && item_to_pixel(p.tiles, item, px, py)) {
        const int pixel_index = ((int)p.height - py - 1) * (int)p.width + px;
        Rng rng;
        rng_init(rng, (unsigned long long)(long long)(1984 + pixel_index));
        float3 col = f3(0.f, 0.f, 0.f);
        for (uint32_t s = 0; s < p.spp; s++) {
            float u = float(px + rng_uniform(rng)) / float(p.width);
            float v = float(py + rng_uniform(rng)) / float(p.height);
            float3 ro, rd;
            camera_ray(p.cam, u, v, ro, rd);
            float3 att = f3(1.0f, 1.0f, 1.0f);
            float3 contrib = f3(0.0f, 0.0f, 0.0f);
            for (uint32_t i = 0; i < p.depth; i++) {
                Hit h = closest_hit<SPHERES, COUNT>(p.scene, ro, rd, 0.001f, n_box, n_tri);
                n_rays++;
                if (!shade<SPHERES, RTOW, COUNT>(p.scene, h, ro, rd, att, rng, contrib, n_light)) break;
                contrib = f3(0.0f, 0.0f, 0.0f);
            }
            col = col + contrib;
        }
        store_pixel(p, pixel_index, col);
    }
    unsigned long long r = n_rays, b = n_box, t = n_tri, l = n_light;
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_down_sync(kFullMask, r, o);
        if (COUNT) {
            b += __shfl_down_sync(kFullMask, b, o);
            t += __shfl_down_sync(kFullMask, t, o);
            l += __shfl_down_sync(kFullMask, l, o);
        }
    }
    if (lane == 0) {
        if (r) atomicAdd(&p.counters->rays, r);
        if (COUNT) {
            atomicAdd(&p.counters->box_tests, b);
            atomicAdd(&p.counters->tri_tests, t);
            atomicAdd(&p.counters->light_tests, l);
        }
    }
}

// Debug/parity probe: one thread walks ONE pixel exactly like the render kernels do and records every ray
// (16 floats per event: sample, bounce, original primitive id or -1, t, bary u, bary v, origin xyz, direction xyz,
// throughput xyz before shading, 0).  Used by tests to find the first event where CUDA and the oracle part ways.
// MODE 0 walks the tree with closest_hit (per-lane loop); MODE 1 / 2 with the resumable steps of the wavefront kernel on the
// 64-byte / the quantised nodes.
template <bool SPHERES, bool RTOW, int MODE>
__global__ void pt_trace_kernel(const __grid_constant__ RenderParams p, int px, int py, float *events, int max_events, int *n_events, float *col_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int pixel_index = ((int)p.height - py - 1) * (int)p.width + px;
    Rng rng;
    rng_init(rng, (unsigned long long)(long long)(1984 + pixel_index));
    float3 col = f3(0.f, 0.f, 0.f);
    int n = 0;
    uint32_t nb = 0, nt = 0, nl = 0;
    for (uint32_t s = 0; s < p.spp; s++) {
        float u = float(px + rng_uniform(rng)) / float(p.width);
        float v = float(py + rng_uniform(rng)) / float(p.height);
        float3 ro, rd;
        camera_ray(p.cam, u, v, ro, rd);
        float3 att = f3(1.0f, 1.0f, 1.0f);
        float3 contrib = f3(0.0f, 0.0f, 0.0f);
        for (uint32_t i = 0; i < p.depth; i++) {
            Hit h;
            if (MODE == 0) {
                h = closest_hit<SPHERES, false>(p.scene, ro, rd, 0.001f, nb, nt);
            } else {
                int32_t stack_mem[kStackSize];
                const LocalStack stack{stack_mem};
                Trav tr;
                if (MODE == 2) trav_begin_grid(tr, stack, p.scene, ro, rd);
                else trav_begin(tr, stack, ro, rd);
                while (!trav_finished(tr)) {
                    if (tr.cur >= 0) trav_node_step<false, MODE == 2>(p.scene, tr, stack, 0.001f, nb);
                    else trav_prim_step<SPHERES, false>(p.scene, tr, stack, ro, rd, 0.001f, nt);
                }
                h = tr.best;
            }
            if (n < max_events) {
                float *e = events + (size_t)n * 16;
                e[0] = (float)s; e[1] = (float)i;
                e[2] = h.prim < 0 ? -1.f : (float)__float_as_int(__ldg(&p.scene.shade[h.prim * 2 + 1]).w);
                e[3] = h.t; e[4] = h.u; e[5] = h.v;
                e[6] = ro.x; e[7] = ro.y; e[8] = ro.z; e[9] = rd.x; e[10] = rd.y; e[11] = rd.z;
                e[12] = att.x; e[13] = att.y; e[14] = att.z; e[15] = 0.f;
            }
            n++;
            if (!shade<SPHERES, RTOW, false>(p.scene, h, ro, rd, att, rng, contrib, nl)) break;
            contrib = f3(0.0f, 0.0f, 0.0f);
        }
        col = col + contrib;
    }
    *n_events = n;
    col_out[0] = col.x; col_out[1] = col.y; col_out[2] = col.z;
}

// Framebuffer gather of the multi-GPU path ("only the final framebuffer gather uses NCCL or P2P over NVLink"): copies the
// pixels of the listed 8x4 blocks — RGB, Y, and the U / V samples their even-row / even-column pixels own (DevicePathTracer.h:107-119) —
// from this GPU's private frame into the frame's master copy on another GPU, with peer stores over NVLink.  One thread per pixel.
__global__ void pt_gather_blocks_kernel(uint8_t *__restrict__ dst_rgb, uint8_t *__restrict__ dst_yuv, const uint8_t *__restrict__ src_rgb, const uint8_t *__restrict__ src_yuv,
                                        const uint32_t *__restrict__ blocks, uint32_t n_blocks, uint32_t width, uint32_t height) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if ((i >> 5) >= n_blocks) return;
    const uint32_t b = __ldg(&blocks[i >> 5]), within = i & 31u;
    const uint32_t x = (b & 0xffffu) * 8u + (within & 7u), y = (b >> 16) * 4u + (within >> 3);
    if (x >= width || y >= height) return;
    const uint32_t row = height - 1u - y, pix = row * width + x;  // RenderTask space is bottom-up, buffer row 0 is the top (:77-79)
    dst_rgb[3 * (size_t)pix + 0] = src_rgb[3 * (size_t)pix + 0];
    dst_rgb[3 * (size_t)pix + 1] = src_rgb[3 * (size_t)pix + 1];
    dst_rgb[3 * (size_t)pix + 2] = src_rgb[3 * (size_t)pix + 2];
    if (!dst_yuv || !src_yuv) return;
    dst_yuv[pix] = src_yuv[pix];
    if ((row & 1u) == 0 && (x & 1u) == 0) {
        const uint32_t total = width * height, uvSize = total / 4, uv = (row / 2) * (width / 2) + (x / 2), limit = total + 2 * uvSize;
        if (total + uv < limit) dst_yuv[total + uv] = src_yuv[total + uv];
        if (total + uvSize + uv < limit) dst_yuv[total + uvSize + uv] = src_yuv[total + uvSize + uv];
    }
}

// camera.h:21-36 evaluated with the device's own tanf / rsqrtf so the numbers are the ones the
// reference's kernels compute (there: by every thread for every sample, into a shared object).
__global__ void pt_camera_kernel(float3 lookFrom, float3 front, float vfov, float hfov, CamParams *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float3 vup = f3(0.0f, 1.0f, 0.0f);
    float3 lookAt = lookFrom + front;
    float theta_v = vfov * PT_M_PI / 180;
    float half_height = tanf(theta_v / 2);
    float theta_h = hfov * PT_M_PI / 180;
    float half_width = tanf(theta_h / 2);
    float3 origin = lookFrom;
    float3 w = normalize(lookFrom - lookAt);
    float3 u = normalize(cross(vup, w));
    float3 v = cross(w, u);
    out->origin = origin;
    out->lower_left_corner = origin - half_width * u - half_height * v - w;
    out->horizontal = 2 * half_width * u;
    out->vertical = 2 * half_height * v;
}

}  // namespace ptc

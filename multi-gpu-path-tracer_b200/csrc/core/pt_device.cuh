// pt_device.cuh — device-side building blocks of the path-tracing core (sm_100a).
//
// Every function cites the reference lines whose RESULT it must reproduce
// (paths relative to /root/reference).  The arithmetic keeps the reference's expression
// shapes (operand order, float vs double literals, rsqrtf / sqrtf / sinf / cosf / IEEE
// division) so that nvcc contracts and rounds it the way it does for the reference's own
// kernels; the data layout, traversal, scheduling and RNG-state handling are new.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

// per-hit data (shading frames, texture coordinates, texels) is read once per ray: through L2 only, so that it does not
// push BVH nodes and triangles out of L1
#define PT_LDSHADE(p) __ldcg(p)
namespace ptc {

#ifndef PT_M_PI
#define PT_M_PI 3.14159265358979323846 /* M_PI as the reference sees it (double) */
#endif

// ----------------------------------------------------------------------------------------
// device scene view (pointers into one bulk-uploaded blob, all 128-byte aligned)
// ----------------------------------------------------------------------------------------
struct TexDesc {
    int32_t width, height;
    uint32_t offset;  // first texel, in float4 units
    uint32_t pad;
};

struct DevScene {
    const float4 *nodes;   // 4 per node     (FlatNode, bvh_builder.h)
    const uint4 *nodesq;   // 2 per node (QuantNode, bvh_builder.h): the same tree with 15-bit box planes on a scene-wide grid, 32 B per node
    const float4 *nodes4;  // 8 per wide node (FlatNode4): lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4] refs[4] pad
    const uint4 *primidx;  // PT_INDEXED_PRIMS: 1 per primitive in leaf order, (i0, i1, i2, kind | shade class << 8): indices into `verts`
    const float4 *verts;   //   unique vertex positions (x, y, z, -), or (c.xyz, r) of a sphere
    const float4 *prims;   // otherwise: 3 per primitive in leaf order:
                           //   triangle: (v0.xyz, e1.x) (e1.yz, e2.xy) (e2.z, -, kind=0, shade class)
                           //   sphere  : (c.xyz, r)     (-,-,-,-)      (-, -, kind=1, shade class)
                           //   shade class (pool kernel): 0 = the path ends here (emitter), 1 = UniversalMaterial / lambertian bounce,
                           //   2 = metal, 3 = dielectric
    const float4 *shade;   // 2 per primitive: (uv0, uv1) (uv2, material index, original primitive index)
    const float4 *mats;    // 3 per material : (type, base.rgb) (emis.rgb, base_tex) (emis_tex, fuzz, ior, -)
    const TexDesc *texs;
    const float4 *texels;  // (r,g,b,-) already scaled by 1/255 exactly as Texture.h:45-46 does
    const float4 *frames;  // 4 per primitive, written once per upload by pt_frames_kernel with the very device functions shade() would
                           //   call per hit (hit_record.normal and the onb of cosine_pdf are per-triangle constants):
                           //   (normal.xyz, material index) (w.xyz, kind) (u.xyz, v.x) (v.yz, -, -); spheres: only .w of the first two
    const float4 *lights;  // 4 per light triangle: (v0.xyz, area) (v1.xyz, n.x) (v2.xyz, n.y) (n.z, -, -, -), scene order;
                           //   n = normalize(cross(v1 - v0, v2 - v0)) by pt_frames_kernel (triangle.h:36)
    int32_t n_lights;
    int32_t n_prims;
    double light_pick_scale;  // (n_lights - 1) + 0.999999, hitable_list.h:24
    float light_weight;       // 1.0f / n_lights, hitable_list.h:17
    int32_t pad;
    float3 grid_lo, grid_scale;  // world -> grid of the quantised nodes: g = (x - grid_lo) * grid_scale, 0 <= g <= 32767
    float pad2[2];
};

struct CamParams {  // camera.h:21-36, evaluated once per set_camera by a 1-thread device kernel
    float3 origin, lower_left_corner, horizontal, vertical;
};

constexpr int kMaxInlineTiles = 48;
struct TileList {
    int32_t n;
    int32_t ox[kMaxInlineTiles], oy[kMaxInlineTiles], w[kMaxInlineTiles], h[kMaxInlineTiles];
    uint32_t first_item[kMaxInlineTiles + 1];  // prefix sums of 32 * ceil(w/8) * ceil(h/4)
};

struct DevCounters {
    unsigned long long rays, box_tests, tri_tests, light_tests, samples;
};

struct RenderParams {
    DevScene scene;
    CamParams cam;
    uint32_t width, height, spp, depth;
    int32_t refill_at;   // wavefront kernel: finished lanes per warp that trigger a shade/refill pass (1..32)
    int32_t node_burst;  // wavefront kernel: node steps taken per vote (1..4)
    uint8_t *fb_rgb, *fb_yuv;
    uint32_t *work_counter;
    DevCounters *counters;
    // explicit work list (ptcore_render_blocks_async): 8x4-pixel blocks, block_list[i] = bx | (by << 16); NULL = use `tiles`
    const uint32_t *block_list;
    uint32_t n_blocks;
    uint32_t pad2;
    // pilot pass (ptcore_block_costs_async): rays traced per 8x4 block are accumulated here and no pixel is stored
    uint32_t *block_cost;
    // pool kernel (pt_pool.cuh): pixel records of this launch, 128 bytes each, pool_size per warp of the grid
    float4 *pool_slots;
    int32_t pool_size;     // pixel slots per warp in use (32 .. kPoolSlots)
    int32_t pool_idle_at;  // with an empty ray ring: waiting hits per warp that trigger a shade pass
    uint32_t watchdog;     // 0, or a bound on the traverse iterations of a warp (debugging aid: a wrong schedule must not hang the GPU)
    int32_t pool_period;   // traverse iterations between two rounds of housekeeping (retire finished rays, pull new ones): 1, 2, 4 or 8
    int32_t lanes_per_warp;  // wavefront kernel: lanes of a warp that take pixels (32 = all; fewer = shorter chains per pixel, see sched)
    int32_t pad4;
    // PT_RNG_SAMPLE_KEYED (ptcore_render_keyed_async): partial colour sums, [chunk][pixel][3] floats; this launch renders chunks
    // keyed_first, keyed_first + keyed_step, ... (keyed_my_chunks of them) of keyed_chunk_spp samples each
    float *keyed_accum;
    uint32_t keyed_chunk_spp, keyed_first, keyed_step, keyed_my_chunks;
    unsigned long long *retire_log;  // optional (ptcore_set_retire_log): [2 * global warp index] = globaltimer at warp start, [+1] = at warp exit
    uint32_t retire_log_warps;       // capacity of retire_log in warps
    uint32_t pad5;
    TileList tiles;
};

// ----------------------------------------------------------------------------------------
// float3 helpers with the expression shapes of helper_math.h
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float b) { return f3(a.x * b, a.y * b, a.z * b); }
__device__ __forceinline__ float3 operator*(float b, float3 a) { return f3(b * a.x, b * a.y, b * a.z); }
__device__ __forceinline__ float3 operator/(float3 a, float b) { return f3(a.x / b, a.y / b, a.z / b); }  // helper_math.h:1015-1018
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }    // :1266-1269
__device__ __forceinline__ float3 cross(float3 a, float3 b) {                                             // :1461-1464
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length(float3 v) { return sqrtf(dot(v, v)); }                            // :1309-1312
__device__ __forceinline__ float3 normalize(float3 v) {                                                   // :1327-1331
    float invLen = rsqrtf(dot(v, v));
    return v * invLen;
}

// ----------------------------------------------------------------------------------------
// cuRAND XORWOW, subsequence 0 / offset 0 (curand_kernel.h:772-797, 863-874; curand_uniform.h:69-72).
// The reference keeps a 48-byte curandState per framebuffer pixel in global memory
// (src/DevicePathTracer.h:352) initialised by render_init (:46-55) and never written back (:80);
// the seed only depends on the pixel index, so the state is rebuilt in registers instead.
// ----------------------------------------------------------------------------------------
struct Rng {
    uint32_t d, v0, v1, v2, v3, v4;
};
__device__ __forceinline__ void rng_init(Rng &s, unsigned long long seed) {
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49U;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddU;
    uint32_t t0 = 1099087573U * s0;
    uint32_t t1 = 2591861531U * s1;
    s.d = 6615241U + t1 + t0;
    s.v0 = 123456789U + t0;
    s.v1 = 362436069U ^ t0;
    s.v2 = 521288629U + t1;
    s.v3 = 88675123U ^ t1;
    s.v4 = 5783321U + t0;
}
__device__ __forceinline__ uint32_t rng_next(Rng &s) {
    uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1;
    s.v1 = s.v2;
    s.v2 = s.v3;
    s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    s.d += 362437U;
    return s.v4 + s.d;
}
__device__ __forceinline__ float rng_uniform(Rng &s) {
    // x * 2^-32 + 2^-33: the product is exact, so fused or not gives the same float
    return (float)rng_next(s) * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}

// ----------------------------------------------------------------------------------------
// comparisons against the reference's double literals, folded to float.
// For a float x:  x <  1e-8  (double)  <=>  x <  kDetEpsUp   (smallest float >= 1e-8)
//                 x > -1e-8  (double)  <=>  x > -kDetEpsUp
//                 x > 0.0001 (double)  <=>  x >  0.0001f      (float(0.0001) < 0.0001)
// (exhaustively checked on the host in tests/test_host_logic.py via the oracle's helpers)
// ----------------------------------------------------------------------------------------
#define PT_DET_EPS_UP 1.00000008274037e-08f

// (float)(1.0 / (double)x) == IEEE float division 1.0f / x for every normal x whose
// reciprocal is normal: the double quotient can never sit within 2^-53 of a float rounding
// boundary (x * midpoint is a 49-bit product, so it is either exactly a power of two or at
// least 2^-49 away in relative terms).  triangle.h:84.
__device__ __forceinline__ float rcp_like_double(float x) { return __fdiv_rn(1.0f, x); }

// cosine / M_PI with the reference's double division (material.h:90, pdf.h:21): (float)((double)c / M_PI).
// Evaluated as q0 = c * RN(1/pi); r = fma(-q0, pi, c); q1 = fma(r, RN(1/pi), q0) — three double ops instead of the
// ~50-instruction IEEE division subroutine.  q1 equals the correctly rounded double quotient for every float c in
// [0, 4] (exhaustive check: tests/c/exactness_check.c, run by tests/test_host_logic.py), so the float result is the
// reference's bit for bit.
__device__ __forceinline__ float div_pi(float c) {
    const double kPi = PT_M_PI, kRcpPi = 1.0 / PT_M_PI;
    const double cd = (double)c;
    const double q0 = cd * kRcpPi;
    const double r = fma(-q0, kPi, cd);
    return (float)fma(r, kRcpPi, q0);
}

// ----------------------------------------------------------------------------------------
// primitive tests
// ----------------------------------------------------------------------------------------
struct Hit {
    float t;      // FLT_MAX when nothing was hit
    float u, v;   // barycentrics of the accepted triangle hit
    int32_t prim; // leaf-order primitive position, -1 for a miss
};

// triangle.h:63-113 with e1/e2 precomputed (the same float subtractions the reference redoes per test).
// Branch-free: every lane evaluates the whole test and folds the reference's early-outs into one predicate;
// the NaN behaviour of the original comparisons (u<0||u>1 etc. are false for NaN) is preserved.
// `wins_ties`: whether this triangle replaces a current hit at EXACTLY the same t.  The reference keeps whichever of the two
// its own tree reaches first (strict t < closest, bvh.h:202-205) — an order no other tree can reproduce; it happens where a
// ray meets the shared edge of two triangles (measured: 3 pixels of a 1080p / 128-spp cornell_duck frame).  Here the lower
// leaf-order position wins, which makes the image independent of the walk (kernel, node format, step schedule).
__device__ __forceinline__ bool triangle_test(float3 v0, float3 e1, float3 e2, float3 o, float3 d, float tmin, float tmax,
                                              float &t_out, float &u_out, float &v_out, bool wins_ties = false) {
    float3 pvec = cross(d, e2);
    float det = dot(e1, pvec);
    float inv_det = rcp_like_double(det);
    float3 tvec = o - v0;
    float u = dot(tvec, pvec) * inv_det;
    float3 qvec = cross(tvec, e1);
    float v = dot(d, qvec) * inv_det;
    float t = dot(e2, qvec) * inv_det;
    bool reject = (fabsf(det) < PT_DET_EPS_UP) | (u < 0) | (u > 1) | (v < 0) | (u + v > 1);
    bool ok = !reject & ((t < tmax) | ((t == tmax) & wins_ties)) & (t > tmin);
    t_out = t;
    u_out = u;
    v_out = v;
    return ok;
}

// sphere.h:21-50 (double-precision spots kept: b = 2.0*dot, roots divided by 2.0*a)
__device__ __forceinline__ bool sphere_test(float3 center, float radius, float3 o, float3 d, float tmin, float tmax, float &t_out) {
    float3 oc = o - center;
    float a = dot(d, d);
    float b = (float)(2.0 * dot(oc, d));
    float c = dot(oc, oc) - radius * radius;
    float discriminant = b * b - 4 * a * c;
    float sq = sqrtf(discriminant);
    float t0 = (float)((-b - sq) / (2.0 * a));
    float t1 = (float)((-b + sq) / (2.0 * a));
    bool ok0 = (discriminant > 0) & (t0 < tmax) & (t0 > tmin);
    bool ok1 = (discriminant > 0) & (t1 < tmax) & (t1 > tmin);
    t_out = ok0 ? t0 : t1;
    return ok0 | ok1;
}

// ----------------------------------------------------------------------------------------
// primitive fetch.  Two layouts of the same data (compile-time, PT_INDEXED_PRIMS):
//   direct    (default) 48-byte record (v0, e1, e2, kind, class), edges precomputed on the host: 3 x LDG.128, one load level.
//   indexed   16-byte record (i0, i1, i2, kind | shade class << 8) + a table of unique vertices (float4): the hot set of
//             cornell_duck shrinks from 203 KB to 103 KB, the 2 M-triangle mesh from 96 MB to 48 MB, at the price of one more
//             DEPENDENT load per leaf step.  e1 = v1 - v0 and e2 = v2 - v0 are then the reference's own float subtractions
//             (triangle.h:67-68) done per test (__fsub_rn: never contracted), so the pixels are the same.
//             Measured on B200 (profiles/r02_prim_layout_ab.txt): -4 % on cornell_duck (2684 -> 2571 Msamples/s), -5 % on the
//             2 M-triangle mesh (74.1 -> 77.8 ms): a warp-wide load waits for its slowest lane whatever the hit rate, so the
//             extra load level costs more than the smaller footprint returns.  Kept as an A/B build, not shipped.
// ----------------------------------------------------------------------------------------
#ifndef PT_INDEXED_PRIMS
#define PT_INDEXED_PRIMS 0
#endif
#ifndef PT_NODE_LDG256
#define PT_NODE_LDG256 1
#endif
struct PrimGeom {
    float3 v0, e1, e2;  // sphere: v0 = centre, e1.x = radius
    int32_t kind;       // 0 triangle, 1 sphere
};
__device__ __forceinline__ PrimGeom load_prim(const DevScene &sc, int32_t k) {
    PrimGeom g;
#if PT_INDEXED_PRIMS
    const uint4 r = __ldg(&sc.primidx[k]);
    const float4 p0 = __ldg(&sc.verts[r.x]), p1 = __ldg(&sc.verts[r.y]), p2 = __ldg(&sc.verts[r.z]);
    g.v0 = f3(p0.x, p0.y, p0.z);
    g.kind = (int32_t)(r.w & 0xffu);
    g.e1 = f3(__fsub_rn(p1.x, p0.x), __fsub_rn(p1.y, p0.y), __fsub_rn(p1.z, p0.z));
    g.e2 = f3(__fsub_rn(p2.x, p0.x), __fsub_rn(p2.y, p0.y), __fsub_rn(p2.z, p0.z));
    if (g.kind == 1) g.e1.x = p0.w;
#else
    const float4 q0 = __ldg(&sc.prims[k * 3 + 0]), q1 = __ldg(&sc.prims[k * 3 + 1]), q2 = __ldg(&sc.prims[k * 3 + 2]);
    g.v0 = f3(q0.x, q0.y, q0.z);
    g.e1 = f3(q0.w, q1.x, q1.y);
    g.e2 = f3(q1.z, q1.w, q2.x);
    g.kind = __float_as_int(q2.z);
#endif
    return g;
}
__device__ __forceinline__ int32_t prim_shade_class(const DevScene &sc, int32_t k) {
#if PT_INDEXED_PRIMS
    return (int32_t)(__ldg(&sc.primidx[k].w) >> 8);
#else
    return __float_as_int(__ldg(&sc.prims[k * 3 + 2]).w);
#endif
}

// ----------------------------------------------------------------------------------------
// closest hit: replaces BVH::hit (bvh.h:178-246) + aabb::hit (aabb.h:38-66).
// Same result (closest accepted triangle::hit over all primitives), different walk:
// t-culled slab tests on both children of a 64-byte node, near child first, stack of far
// children in per-thread local memory, leaves of <= 8 primitives fetched as 3 x LDG.128.
// ----------------------------------------------------------------------------------------
constexpr int kStackSize = 64;

// Reciprocal direction for the fma-form slab test t = b * inv + (-o * inv).  A zero (or denormal) direction
// component would give inv = inf and then inf - inf = NaN exactly when the origin lies between the slab planes,
// which min/max would turn into "box missed" — culling a box that holds a valid hit (seen on B200 as 2 wrong
// pixels in the 64x48 camera fixture: primary rays with d.y == 0).  Components are therefore clamped away from
// zero to 2^-80 (sign kept); the slab distances then belong to a ray that is off by < 1e-18 scene units over any
// t the scene allows, far inside the boxes' host-side padding, so the test stays conservative.  Only the BOX
// tests see this: the primitive tests use the unmodified direction, as the reference does.
#ifndef PT_SLAB_RCP_APPROX
#define PT_SLAB_RCP_APPROX 1
#endif
__device__ __forceinline__ float3 slab_inverse(float3 d) {
    const float kTiny = 8.271806125530277e-25f;  // 2^-80
    const float dx = fabsf(d.x) < kTiny ? copysignf(kTiny, d.x) : d.x;
    const float dy = fabsf(d.y) < kTiny ? copysignf(kTiny, d.y) : d.y;
    const float dz = fabsf(d.z) < kTiny ? copysignf(kTiny, d.z) : d.z;
    // box tests only steer the walk (every hit is decided by the primitive tests, which use the exact direction), and the boxes are padded by
    // 1e-5 of their coordinates on the host: the one-ulp error of MUFU.RCP is far inside that, and one instruction replaces the ~10 of an
    // IEEE division, three times per ray (1.7 % of the executed instructions in profiles/r02_ncu_wavefront_source_lines.txt)
#if PT_SLAB_RCP_APPROX
    float ix, iy, iz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ix) : "f"(dx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iy) : "f"(dy));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(dz));
    return f3(ix, iy, iz);
#else
    return f3(1.0f / dx, 1.0f / dy, 1.0f / dz);
#endif
}

template <bool SPHERES, bool COUNT>
__device__ __forceinline__ Hit closest_hit(const DevScene &sc, float3 o, float3 d, float tmin, uint32_t &n_box, uint32_t &n_tri) {
    Hit best;
    best.t = FLT_MAX;
    best.u = best.v = 0.f;
    best.prim = -1;

    const float3 inv = slab_inverse(d);
    const float3 oinv = f3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    const float kSlack = 1.0000004f;  // boxes are padded on the host; this covers the fma rounding of the slab distances

    int32_t stack[kStackSize];
    int sp = 0;
    int32_t node = 0;
    for (;;) {
        if (node >= 0) {
            const float4 bx = __ldg(&sc.nodes[node * 4 + 0]);
            const float4 by = __ldg(&sc.nodes[node * 4 + 1]);
            const float4 bz = __ldg(&sc.nodes[node * 4 + 2]);
            const int4 refs = __ldg(reinterpret_cast<const int4 *>(&sc.nodes[node * 4 + 3]));
            if (COUNT) n_box += 2;
            // left
            float lx0 = fmaf(bx.x, inv.x, oinv.x), lx1 = fmaf(bx.y, inv.x, oinv.x);
            float ly0 = fmaf(by.x, inv.y, oinv.y), ly1 = fmaf(by.y, inv.y, oinv.y);
            float lz0 = fmaf(bz.x, inv.z, oinv.z), lz1 = fmaf(bz.y, inv.z, oinv.z);
            float ln = fmaxf(fmaxf(fminf(lx0, lx1), fminf(ly0, ly1)), fmaxf(fminf(lz0, lz1), tmin));
            float lf = fminf(fminf(fmaxf(lx0, lx1), fmaxf(ly0, ly1)), fminf(fmaxf(lz0, lz1), best.t));
            // right
            float rx0 = fmaf(bx.z, inv.x, oinv.x), rx1 = fmaf(bx.w, inv.x, oinv.x);
            float ry0 = fmaf(by.z, inv.y, oinv.y), ry1 = fmaf(by.w, inv.y, oinv.y);
            float rz0 = fmaf(bz.z, inv.z, oinv.z), rz1 = fmaf(bz.w, inv.z, oinv.z);
            float rn = fmaxf(fmaxf(fminf(rx0, rx1), fminf(ry0, ry1)), fmaxf(fminf(rz0, rz1), tmin));
            float rf = fminf(fminf(fmaxf(rx0, rx1), fmaxf(ry0, ry1)), fminf(fmaxf(rz0, rz1), best.t));
            bool hl = ln <= lf * kSlack;
            bool hr = rn <= rf * kSlack;
            if (hl & hr) {
                bool leftFirst = ln <= rn;
                int32_t nearRef = leftFirst ? refs.x : refs.y;
                int32_t farRef = leftFirst ? refs.y : refs.x;
                stack[sp++] = farRef;
                node = nearRef;
                continue;
            }
            if (hl) { node = refs.x; continue; }
            if (hr) { node = refs.y; continue; }
        } else {
            const int32_t v = ~node;  // (first << 4) | count, bvh_builder.h
            const int32_t first = v >> 4;
            const int32_t count = v & 15;
            for (int32_t i = 0; i < count; i++) {
                const int32_t k = first + i;
                const PrimGeom g = load_prim(sc, k);
                if (COUNT) n_tri += 1;
                if (SPHERES && g.kind == 1) {
                    float t;
                    if (sphere_test(g.v0, g.e1.x, o, d, tmin, best.t, t)) {
                        best.t = t;
                        best.u = best.v = 0.f;
                        best.prim = k;
                    }
                } else {
                    float t, u, w;
                    if (triangle_test(g.v0, g.e1, g.e2, o, d, tmin, best.t, t, u, w, k < best.prim)) {
                        best.t = t;
                        best.u = u;
                        best.v = w;
                        best.prim = k;
                    }
                }
            }
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
    return best;
}

// ----------------------------------------------------------------------------------------
// resumable traversal: the same closest-hit search cut into uniform STEPS so that a warp can run
// "everyone who is at an inner node takes one node step" / "everyone who holds a leaf primitive
// tests one primitive" under warp votes (pt_wavefront_kernel).  A lane may hold one postponed
// leaf while it keeps descending, so node steps and primitive steps batch up across the warp.
// ----------------------------------------------------------------------------------------
constexpr int32_t kTravDone = (int32_t)0x80000000;  // `cur` when nothing is left to visit: negative, but never a valid leaf reference

// Lane state of a resumable traversal.  Encodings are chosen so that the two warp votes are single instructions:
//   can take a node step       <=>  cur >= 0
//   can take a primitive step  <=>  (leaf & 15) != 0
// stack[0] always holds kTravDone, so a pop never needs an emptiness check.
struct Trav {
    int32_t cur;        // >= 0 inner node, < 0 leaf reference (blocked until the held leaf is done), kTravDone = exhausted
    int32_t leaf;       // cursor of the held leaf: (next primitive position << 4) | primitives left; the complement of a leaf
                        // reference is such a cursor, and "+ 15" moves it on by one primitive
    int32_t sp;
    float3 inv, oinv;
    uint3 sel, self;    // quantised nodes: byte-permute selectors of the plane the ray reaches first / last, per axis
    Hit best;
};

// Where the postponed children of a lane live.  LocalStack: a per-thread array (local memory, L1-resident).  ShortStack<K>:
// the first K entries in shared memory — column `lane` of a [K][32] array per warp, so the 32 lanes of a warp hit 32 different
// banks whatever their depths — and only deeper entries in a per-thread overflow array (measured with the host restatement of
// the walk: 0.8 % of the pushes on cornell_duck and 3.5 % on a 180 K-triangle mesh go deeper than 8).
struct LocalStack {
    int32_t *a;
    __device__ __forceinline__ void put(int i, int32_t v) const { a[i] = v; }
    __device__ __forceinline__ int32_t get(int i) const { return a[i]; }
};
template <int K>
struct ShortStack {
    uint32_t col;  // shared-memory address of warp_stack[0][lane]; entry i < K is at col + i * 128
    int32_t *ovf;  // entries K .. kStackSize-1
    // the shared-memory access is ONE predicated instruction (no branch, so lanes at different depths do not diverge); the
    // overflow array is touched by a few per cent of the pushes only
    __device__ __forceinline__ void put(int i, int32_t v) const {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %0, %1;\n\t@p st.shared.b32 [%2], %3;\n\t}" ::"r"(i), "n"(K), "r"(col + (uint32_t)i * 128u), "r"(v) : "memory");
        if (i >= K) ovf[i - K] = v;
    }
    __device__ __forceinline__ int32_t get(int i) const {
        int32_t v;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, %2;\n\tmov.b32 %0, 0;\n\t@p ld.shared.b32 %0, [%3];\n\t}" : "=r"(v) : "r"(i), "n"(K), "r"(col + (uint32_t)i * 128u) : "memory");
        if (i >= K) v = ovf[i - K];
        return v;
    }
};

template <class STACK>
__device__ __forceinline__ STACK make_stack(int32_t *local, uint32_t smem_col);
template <>
__device__ __forceinline__ LocalStack make_stack<LocalStack>(int32_t *local, uint32_t) { return LocalStack{local}; }
template <>
__device__ __forceinline__ ShortStack<8> make_stack<ShortStack<8>>(int32_t *local, uint32_t smem_col) { return ShortStack<8>{smem_col, local}; }

__device__ __forceinline__ bool trav_leaf_held(const Trav &t) { return (t.leaf & 15) != 0; }
template <class STACK>
__device__ __forceinline__ void trav_push(Trav &t, const STACK &stack, int32_t x) { stack.put(t.sp++, x); }
template <class STACK>
__device__ __forceinline__ int32_t trav_pop(Trav &t, const STACK &stack) { return stack.get(--t.sp); }

__device__ __forceinline__ void trav_idle(Trav &t) {
    t.cur = kTravDone;
    t.leaf = 0;
}
// A ray prepared for the slab tests: reciprocal direction and the constant term of t = plane * inv + oinv (float planes), or the
// same in the grid coordinates of the quantised nodes (g = (x - grid_lo) * grid_scale per axis, an affine map that leaves the ray
// parameter t unchanged; a stored plane p (15 bits) becomes the float 2^15 + p by ONE byte permute — exponent byte 0x47, p in
// mantissa bits 8..22 — so the offset 2^15 is folded into the constant term here).
__device__ __forceinline__ void trav_prepare(float3 o, float3 d, float3 &inv, float3 &oinv) {
    inv = slab_inverse(d);
    oinv = f3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
}
__device__ __forceinline__ void trav_prepare_grid(const DevScene &sc, float3 o, float3 d, float3 &inv, float3 &oinv) {
    const float3 og = f3((o.x - sc.grid_lo.x) * sc.grid_scale.x, (o.y - sc.grid_lo.y) * sc.grid_scale.y, (o.z - sc.grid_lo.z) * sc.grid_scale.z);
    inv = slab_inverse(f3(d.x * sc.grid_scale.x, d.y * sc.grid_scale.y, d.z * sc.grid_scale.z));
    oinv = f3(-(32768.0f + og.x) * inv.x, -(32768.0f + og.y) * inv.y, -(32768.0f + og.z) * inv.z);
}
// start the walk of a prepared ray at the root
template <bool QUANT, class STACK>
__device__ __forceinline__ void trav_start(Trav &t, const STACK &stack, float3 inv, float3 oinv) {
    t.cur = 0;
    t.leaf = 0;
    stack.put(0, kTravDone);
    t.sp = 1;
    t.inv = inv;
    t.oinv = oinv;
    t.best.t = FLT_MAX;
    t.best.u = t.best.v = 0.f;
    t.best.prim = -1;
    if (QUANT) {
        // low half of a word = min plane (selector 0x7104), high half = max plane (0x7324); a ray going down an axis meets max first
        t.sel = make_uint3(inv.x < 0.f ? 0x7324u : 0x7104u, inv.y < 0.f ? 0x7324u : 0x7104u, inv.z < 0.f ? 0x7324u : 0x7104u);
        t.self = make_uint3(t.sel.x ^ 0x0220u, t.sel.y ^ 0x0220u, t.sel.z ^ 0x0220u);
    }
}
template <class STACK>
__device__ __forceinline__ void trav_begin(Trav &t, const STACK &stack, float3 o, float3 d) {
    float3 inv, oinv;
    trav_prepare(o, d, inv, oinv);
    trav_start<false>(t, stack, inv, oinv);
}
template <class STACK>
__device__ __forceinline__ void trav_begin_grid(Trav &t, const STACK &stack, const DevScene &sc, float3 o, float3 d) {
    float3 inv, oinv;
    trav_prepare_grid(sc, o, d, inv, oinv);
    trav_start<true>(t, stack, inv, oinv);
}
__device__ __forceinline__ bool trav_finished(const Trav &t) { return t.cur == kTravDone && !trav_leaf_held(t); }
__device__ __forceinline__ void trav_hold_leaf(Trav &t, int32_t ref) { t.leaf = ~ref; }

// prmt.b32 without the selector masking that __byte_perm adds (the selectors here are plain byte indices)
__device__ __forceinline__ float plane_float(uint32_t word, uint32_t selector) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(0x47000000u), "r"(selector));
    return __uint_as_float(r);
}

// one inner-node step (precondition: t.cur >= 0)
// SMEM: the quantised nodes were copied to shared memory by the CTA (pt_wavefront_smem_kernel); `smem_nodes` is their
// shared-space address.  A node fetch then takes the fixed shared-memory latency instead of waiting for whichever lane of
// the warp missed L1.
template <bool COUNT, bool QUANT = false, class STACK = LocalStack, bool SMEM = false>
__device__ __forceinline__ void trav_node_step(const DevScene &sc, Trav &t, const STACK &stack, float tmin, uint32_t &n_box, uint32_t smem_nodes = 0) {
    const float kSlack = 1.0000004f;
    const int32_t node = t.cur;
    float ln, lf, rn, rf;
    int4 refs;
    if (QUANT) {
        // the whole 32-byte node with ONE 256-bit load (sm_100: LDG.E.256): half the load instructions of the walk
        uint4 a, b;
        if (SMEM) {
            const uint32_t addr = smem_nodes + (uint32_t)node * 32u;
            asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(addr));
            asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(addr + 16u));
        } else {
#if PT_NODE_LDG256
        asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
            : "l"(sc.nodesq + (size_t)node * 2));
#else
        a = __ldg(&sc.nodesq[node * 2 + 0]);
        b = __ldg(&sc.nodesq[node * 2 + 1]);
#endif
        }
        if (COUNT) n_box += 2;
        // plane p -> float 2^15 + p by one byte permute; the selector picks the half of the word that the ray reaches first /
        // last on this axis (t.sel, from the sign of the direction), so no per-axis min / max is needed
        const uint32_t nx = t.sel.x, ny = t.sel.y, nz = t.sel.z;
        const uint32_t fx = t.self.x, fy = t.self.y, fz = t.self.z;
#define PT_PLANE(w, s) plane_float((w), (s))
        const float lnx = fmaf(PT_PLANE(a.x, nx), t.inv.x, t.oinv.x), lfx = fmaf(PT_PLANE(a.x, fx), t.inv.x, t.oinv.x);
        const float rnx = fmaf(PT_PLANE(a.y, nx), t.inv.x, t.oinv.x), rfx = fmaf(PT_PLANE(a.y, fx), t.inv.x, t.oinv.x);
        const float lny = fmaf(PT_PLANE(a.z, ny), t.inv.y, t.oinv.y), lfy = fmaf(PT_PLANE(a.z, fy), t.inv.y, t.oinv.y);
        const float rny = fmaf(PT_PLANE(a.w, ny), t.inv.y, t.oinv.y), rfy = fmaf(PT_PLANE(a.w, fy), t.inv.y, t.oinv.y);
        const float lnz = fmaf(PT_PLANE(b.x, nz), t.inv.z, t.oinv.z), lfz = fmaf(PT_PLANE(b.x, fz), t.inv.z, t.oinv.z);
        const float rnz = fmaf(PT_PLANE(b.y, nz), t.inv.z, t.oinv.z), rfz = fmaf(PT_PLANE(b.y, fz), t.inv.z, t.oinv.z);
#undef PT_PLANE
        ln = fmaxf(fmaxf(lnx, lny), fmaxf(lnz, tmin));
        lf = fminf(fminf(lfx, lfy), fminf(lfz, t.best.t));
        rn = fmaxf(fmaxf(rnx, rny), fmaxf(rnz, tmin));
        rf = fminf(fminf(rfx, rfy), fminf(rfz, t.best.t));
        refs = make_int4((int32_t)b.z, (int32_t)b.w, 0, 0);
    } else {
        const float4 bx = __ldg(&sc.nodes[node * 4 + 0]);
        const float4 by = __ldg(&sc.nodes[node * 4 + 1]);
        const float4 bz = __ldg(&sc.nodes[node * 4 + 2]);
        refs = __ldg(reinterpret_cast<const int4 *>(&sc.nodes[node * 4 + 3]));
        if (COUNT) n_box += 2;
        float lx0 = fmaf(bx.x, t.inv.x, t.oinv.x), lx1 = fmaf(bx.y, t.inv.x, t.oinv.x);
        float ly0 = fmaf(by.x, t.inv.y, t.oinv.y), ly1 = fmaf(by.y, t.inv.y, t.oinv.y);
        float lz0 = fmaf(bz.x, t.inv.z, t.oinv.z), lz1 = fmaf(bz.y, t.inv.z, t.oinv.z);
        ln = fmaxf(fmaxf(fminf(lx0, lx1), fminf(ly0, ly1)), fmaxf(fminf(lz0, lz1), tmin));
        lf = fminf(fminf(fmaxf(lx0, lx1), fmaxf(ly0, ly1)), fminf(fmaxf(lz0, lz1), t.best.t));
        float rx0 = fmaf(bx.z, t.inv.x, t.oinv.x), rx1 = fmaf(bx.w, t.inv.x, t.oinv.x);
        float ry0 = fmaf(by.z, t.inv.y, t.oinv.y), ry1 = fmaf(by.w, t.inv.y, t.oinv.y);
        float rz0 = fmaf(bz.z, t.inv.z, t.oinv.z), rz1 = fmaf(bz.w, t.inv.z, t.oinv.z);
        rn = fmaxf(fmaxf(fminf(rx0, rx1), fminf(ry0, ry1)), fmaxf(fminf(rz0, rz1), tmin));
        rf = fminf(fminf(fmaxf(rx0, rx1), fmaxf(ry0, ry1)), fminf(fmaxf(rz0, rz1), t.best.t));
    }
    const bool hl = ln <= lf * kSlack;
    const bool hr = rn <= rf * kSlack;
    // Successor selection with predication instead of nested branches (this tail was 17 % of all executed instructions at
    // 10 active threads per instruction in profiles/r01_ncu_wavefront_final_*.txt):
    const bool both = hl & hr;
    const bool any = hl | hr;
    const bool rightNear = hr & (!hl | (rn < ln));
    const int32_t nearRef = rightNear ? refs.y : refs.x;
    const int32_t farRef = rightNear ? refs.x : refs.y;
    const bool holdNear = any & (nearRef < 0) & !trav_leaf_held(t);  // reached a leaf and the slot is free: hold it
    const bool needPush = both & !holdNear;
    const bool needPop = !any | (holdNear & !both);
    int32_t next = holdNear ? farRef : nearRef;
    if (needPush) stack.put(t.sp, farRef);
    t.sp += needPush ? 1 : 0;
    t.sp -= needPop ? 1 : 0;
    if (needPop) next = stack.get(t.sp);
    if (holdNear) trav_hold_leaf(t, nearRef);
    if (next < 0 && next != kTravDone && !trav_leaf_held(t)) {  // the successor is itself a leaf and the slot is (still) free
        trav_hold_leaf(t, next);
        next = trav_pop(t, stack);
    }
    t.cur = next;
}

// one step on a four-wide node (precondition: t.cur >= 0): four independent slab tests, continue with the nearest
// child that was hit, push the other hit children
template <bool COUNT, class STACK = LocalStack>
__device__ __forceinline__ void trav_node_step4(const DevScene &sc, Trav &t, const STACK &stack, float tmin, uint32_t &n_box) {
    const float kSlack = 1.0000004f;
    const float4 *nd = sc.nodes4 + (size_t)t.cur * 8;
    const float4 lox = __ldg(nd + 0), hix = __ldg(nd + 1), loy = __ldg(nd + 2), hiy = __ldg(nd + 3), loz = __ldg(nd + 4), hiz = __ldg(nd + 5);
    const int4 refs = __ldg(reinterpret_cast<const int4 *>(nd + 6));
    if (COUNT) n_box += 4;
    // an unused slot (all planes +inf, bvh_builder.cpp emptySlot) passes the slab test of a ray with inv > 0 while nothing has been hit yet
    // (far = FLT_MAX * slack = +inf): it is excluded by its reference — an empty leaf in `cur` behind an empty held leaf would never be stepped
    constexpr int32_t kEmptyLeaf = -1;  // leaf_ref(0, 0)
#define PT_SLAB4(c)                                                                                                   \
    const float ax##c = fmaf(lox.c, t.inv.x, t.oinv.x), bx##c = fmaf(hix.c, t.inv.x, t.oinv.x);                       \
    const float ay##c = fmaf(loy.c, t.inv.y, t.oinv.y), by##c = fmaf(hiy.c, t.inv.y, t.oinv.y);                       \
    const float az##c = fmaf(loz.c, t.inv.z, t.oinv.z), bz##c = fmaf(hiz.c, t.inv.z, t.oinv.z);                       \
    const float tn##c = fmaxf(fmaxf(fminf(ax##c, bx##c), fminf(ay##c, by##c)), fmaxf(fminf(az##c, bz##c), tmin));     \
    const float tf##c = fminf(fminf(fmaxf(ax##c, bx##c), fmaxf(ay##c, by##c)), fminf(fmaxf(az##c, bz##c), t.best.t)); \
    const bool h##c = tn##c <= tf##c * kSlack && refs.c != kEmptyLeaf;
    PT_SLAB4(x)
    PT_SLAB4(y)
    PT_SLAB4(z)
    PT_SLAB4(w)
#undef PT_SLAB4
    // entry distances are >= tmin > 0, so their bit patterns order like unsigned integers; the two low mantissa bits
    // carry the slot number (ordering only steers the walk, it never decides a hit)
    const uint32_t k0 = hx ? ((__float_as_uint(tnx) & ~3u) | 0u) : 0xffffffffu;
    const uint32_t k1 = hy ? ((__float_as_uint(tny) & ~3u) | 1u) : 0xffffffffu;
    const uint32_t k2 = hz ? ((__float_as_uint(tnz) & ~3u) | 2u) : 0xffffffffu;
    const uint32_t k3 = hw ? ((__float_as_uint(tnw) & ~3u) | 3u) : 0xffffffffu;
    // sort the four keys (5-comparator network); misses (0xffffffff) end up last
    const uint32_t a0 = min(k0, k1), a1 = max(k0, k1), a2 = min(k2, k3), a3 = max(k2, k3);
    const uint32_t s0 = min(a0, a2), b1 = max(a0, a2), b2 = min(a1, a3), s3 = max(a1, a3);
    const uint32_t s1 = min(b1, b2), s2 = max(b1, b2);
#define PT_REF4(k) (((k) & 2u) ? (((k) & 1u) ? refs.w : refs.z) : (((k) & 1u) ? refs.y : refs.x))
    // far children first, so that the nearer ones are popped first
    if (s3 != 0xffffffffu) trav_push(t, stack, PT_REF4(s3));
    if (s2 != 0xffffffffu) trav_push(t, stack, PT_REF4(s2));
    if (s1 != 0xffffffffu) trav_push(t, stack, PT_REF4(s1));
    int32_t next = PT_REF4(s0);
#undef PT_REF4
    if (s0 == 0xffffffffu) next = trav_pop(t, stack);
    if (next < 0 && next != kTravDone && !trav_leaf_held(t)) {  // a leaf and the slot is free: hold it, continue elsewhere
        trav_hold_leaf(t, next);
        next = trav_pop(t, stack);
    }
    t.cur = next;
}

// one primitive of the held leaf (precondition: trav_leaf_held(t))
template <bool SPHERES, bool COUNT, class STACK = LocalStack>
__device__ __forceinline__ void trav_prim_step(const DevScene &sc, Trav &t, const STACK &stack, float3 o, float3 d, float tmin, uint32_t &n_tri) {
    const int32_t k = t.leaf >> 4;
    const PrimGeom g = load_prim(sc, k);
    if (COUNT) n_tri += 1;
    if (SPHERES && g.kind == 1) {
        float tt;
        if (sphere_test(g.v0, g.e1.x, o, d, tmin, t.best.t, tt)) {
            t.best.t = tt;
            t.best.u = t.best.v = 0.f;
            t.best.prim = k;
        }
    } else {
        float tt, u, w;
        if (triangle_test(g.v0, g.e1, g.e2, o, d, tmin, t.best.t, tt, u, w, k < t.best.prim)) {
            t.best.t = tt;
            t.best.u = u;
            t.best.v = w;
            t.best.prim = k;
        }
    }
    t.leaf += 15;  // position + 1, primitives left - 1
    if (!trav_leaf_held(t) && t.cur < 0 && t.cur != kTravDone) {  // a second leaf was waiting in `cur`
        trav_hold_leaf(t, t.cur);
        t.cur = trav_pop(t, stack);
    }
}

// up to two primitives of the held leaf in one step (triangle-only scenes): the two tests are independent instruction streams
// (a lane's chain of dependent instructions is what bounds the kernel, the FMA pipe is mostly idle), the two acceptances then
// happen in leaf order, exactly as two single steps would
template <bool COUNT, class STACK = LocalStack>
__device__ __forceinline__ void trav_prim_step2(const DevScene &sc, Trav &t, const STACK &stack, float3 o, float3 d, float tmin, uint32_t &n_tri) {
    const int32_t k = t.leaf >> 4;
    const bool two = (t.leaf & 15) >= 2;
    const int32_t k1 = two ? k + 1 : k;
    const PrimGeom ga = load_prim(sc, k), gb = load_prim(sc, k1);
    if (COUNT) n_tri += two ? 2 : 1;
    float ta, ua, va, tb, ub, vb;
    // range test against (tmin, +inf): the comparison with the current closest hit follows, in order
    const bool oka = triangle_test(ga.v0, ga.e1, ga.e2, o, d, tmin, FLT_MAX, ta, ua, va);
    const bool okb = triangle_test(gb.v0, gb.e1, gb.e2, o, d, tmin, FLT_MAX, tb, ub, vb) & two;
    if (oka & ((ta < t.best.t) | ((ta == t.best.t) & (k < t.best.prim)))) {
        t.best.t = ta;
        t.best.u = ua;
        t.best.v = va;
        t.best.prim = k;
    }
    if (okb & ((tb < t.best.t) | ((tb == t.best.t) & (k1 < t.best.prim)))) {
        t.best.t = tb;
        t.best.u = ub;
        t.best.v = vb;
        t.best.prim = k1;
    }
    t.leaf += two ? 30 : 15;
    if (!trav_leaf_held(t) && t.cur < 0 && t.cur != kTravDone) {  // a second leaf was waiting in `cur`
        trav_hold_leaf(t, t.cur);
        t.cur = trav_pop(t, stack);
    }
}

// ----------------------------------------------------------------------------------------
// shading pieces
// ----------------------------------------------------------------------------------------
struct Onb {
    float3 u, v, w;
};
__device__ __forceinline__ Onb make_onb(float3 n) {  // onb.h:8-13
    Onb o;
    o.w = normalize(n);
    float3 a = fabsf(o.w.x) > 0.9f ? f3(0.0f, 1.0f, 0.0f) : f3(1.0f, 0.0f, 0.0f);
    o.v = normalize(cross(o.w, a));
    o.u = cross(o.w, o.v);
    return o;
}
__device__ __forceinline__ float3 onb_local(const Onb &o, float3 a) {  // onb.h:19-21
    return a.x * o.u + a.y * o.v + a.z * o.w;
}

__device__ __forceinline__ float3 random_cosine_direction(Rng &rng) {  // helper_math.h:1519-1528 (2x factor is the reference's)
    float r1 = rng_uniform(rng);
    float r2 = rng_uniform(rng);
    float z = sqrtf(1 - r2);
    float phi = (float)(2 * PT_M_PI * r1);
    float sn, cs;
    sincosf(phi, &sn, &cs);  // one argument reduction; bit-identical to cosf(phi) / sinf(phi) (gate A, tests/test_gpu_parity.py)
    float x = cs * 2 * sqrtf(r2);
    float y = sn * 2 * sqrtf(r2);
    return f3(x, y, z);
}

__device__ __forceinline__ float3 random_in_unit_sphere(Rng &rng) {  // helper_math.h:1504-1518, draw order x,y,z
    float3 p;
    do {
        float a = rng_uniform(rng), b = rng_uniform(rng), c = rng_uniform(rng);
        p = 2.0f * f3(a, b, c) - f3(1.0f, 1.0f, 1.0f);
    } while (dot(p, p) >= 1.0f);
    return p;
}

__device__ __forceinline__ float3 texture_value(const DevScene &sc, int tex, float u, float v) {  // Texture.h:30-70
    const TexDesc td = sc.texs[tex];
    if (td.height <= 0) return f3((float)242 / 255, (float)45 / 255, (float)27 / 255);
    // fmodf(x, 1) == x exactly for 0 <= x < 1 (the usual case); the library call only for the rest
    if (!(u >= 0.0f && u < 1.0f)) u = fmodf(u, 1.0f);
    if (!(v >= 0.0f && v < 1.0f)) v = fmodf(v, 1.0f);
    int i = (int)(u * td.width);
    int j = (int)(v * td.height);
    int x = i < 0 ? 0 : (i < td.width ? i : td.width - 1);
    int y = j < 0 ? 0 : (j < td.height ? j : td.height - 1);
    y = td.height - y;
    if (y >= td.height) y = td.height - 1;  // the reference reads one row past the end here; defined as the last row
    const float4 t = PT_LDSHADE(&sc.texels[td.offset + (uint32_t)y * (uint32_t)td.width + (uint32_t)x]);
    return f3(t.x, t.y, t.z);
}

// triangle::pdf_value (triangle.h:32-40) for one light triangle stored as v0/v1/v2/area
__device__ __forceinline__ float3 light_normal(float3 v0, float3 v1, float3 v2) { return normalize(cross(v1 - v0, v2 - v0)); }
__device__ __forceinline__ float light_pdf_value(const float4 *lt, float3 v0, float3 v1, float3 v2, float nx, float ny, float area, float3 o, float3 dir) {
    float3 e1 = v1 - v0;
    float3 e2 = v2 - v0;
    float t, u, v;
    if (!triangle_test(v0, e1, e2, o, dir, 0.001f, FLT_MAX, t, u, v)) return 0;
    float3 normal = f3(nx, ny, __ldg(&lt[3]).x);
    float distance_squared = t * t * dot(dir, dir);
    float cosine = fabsf(dot(dir, normal) / length(dir));
    return distance_squared / (cosine * area);
}

__device__ __forceinline__ float3 reflect3(float3 i, float3 n) { return i - 2.0f * n * dot(n, i); }  // helper_math.h:1429-1432
__device__ __forceinline__ bool refract3(float3 v, float3 n, float ni_over_nt, float3 &refracted) {  // helper_math.cu:7-17
    float3 uv = normalize(v);
    float dt = dot(uv, n);
    float discriminant = 1.0f - ni_over_nt * ni_over_nt * (1 - dt * dt);
    if (discriminant > 0) {
        refracted = ni_over_nt * (uv - n * dt) - n * sqrtf(discriminant);
        return true;
    }
    return false;
}
__device__ __forceinline__ float schlick(float cosine, float ref_idx) {  // material.h:10-14, pow5 by multiplication
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    float x = 1 - cosine;
    float x2 = x * x, x4 = x2 * x2, x5 = x4 * x;
    return r0 + (1 - r0) * x5;
}

// x / y, bit for bit, for a numerator that is often exactly zero (scattering_pdf == 0 below the surface): the compiler's IEEE
// division sends a zero numerator down its out-of-line slow path (~35 instructions for a handful of lanes); 0 / y is (+-)0
// for every y that is neither 0 nor NaN.
__device__ __forceinline__ float div_often_zero(float x, float y) {
    const bool z = (x == 0.0f) & (fabsf(y) > 0.0f);
    const float q = (z ? 1.0f : x) / y;
    return z ? __int_as_float((__float_as_int(x) ^ __float_as_int(y)) & (int)0x80000000) : q;
}

// camera.h:95-97
__device__ __forceinline__ void camera_ray(const CamParams &c, float u, float v, float3 &o, float3 &d) {
    o = c.origin;
    d = c.lower_left_corner + u * c.horizontal + v * c.vertical - c.origin;
}

// quantiser + RGB8/I420 store: DevicePathTracer.h:98-119
__device__ __forceinline__ void store_pixel(const RenderParams &p, int pixel_index, float3 col) {
    float3 q = 255.99f * col / (float)p.spp * f3(1.f, 1.f, 1.f);
    int r = min(255, __float2int_rz(q.x));
    int g = min(255, __float2int_rz(q.y));
    int b = min(255, __float2int_rz(q.z));
    uint8_t *rgb = p.fb_rgb + 3 * (size_t)pixel_index;
    rgb[0] = (uint8_t)r;
    rgb[1] = (uint8_t)g;
    rgb[2] = (uint8_t)b;
    if (p.fb_yuv) {
        p.fb_yuv[pixel_index] = (uint8_t)(((66 * r + 129 * g + 25 * b + 128) >> 8) + 16);
        int blockRow = pixel_index / (int)p.width;
        int blockCol = pixel_index % (int)p.width;
        if (blockRow % 2 == 0 && blockCol % 2 == 0) {
            int totalPixels = (int)(p.width * p.height);
            int uvSize = totalPixels / 4;
            int uvIndex = (blockRow / 2) * ((int)p.width / 2) + (blockCol / 2);
            // With an odd width or height the reference's index runs past its own W*H + 2*(W*H/4) byte buffer
            // (undefined behaviour there); writes that would leave that buffer are dropped here.
            const int limit = totalPixels + 2 * uvSize;
            if (totalPixels + uvIndex < limit) p.fb_yuv[totalPixels + uvIndex] = (uint8_t)(((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128);
            if (totalPixels + uvSize + uvIndex < limit) p.fb_yuv[totalPixels + uvSize + uvIndex] = (uint8_t)(((112 * r - 94 * g - 18 * b + 128) >> 8) + 128);
        }
    }
}

// ----------------------------------------------------------------------------------------
// shade: everything camera::ray_color (camera.h:49-83) does between two BVH::hit calls.
// Returns true when the path continues (o, d, att updated), false when it ended
// (`contrib` then holds what the sample adds to the pixel).
// ----------------------------------------------------------------------------------------
template <bool SPHERES, bool RTOW, bool COUNT>
__device__ __forceinline__ bool shade(const DevScene &sc, const Hit &h, float3 &o, float3 &d, float3 &att, Rng &rng, float3 &contrib,
                                      uint32_t &n_light) {
    if (h.prim < 0) {
        contrib = f3(0.0f, 0.0f, 0.0f) * att;  // camera.h:79,109: background (0,0,0) * attenuation (NaN/inf propagate as in the reference)
        return false;
    }
    const float4 f0 = PT_LDSHADE(&sc.frames[h.prim * 4 + 0]);
    const float4 f1 = PT_LDSHADE(&sc.frames[h.prim * 4 + 1]);
    const int mat = __float_as_int(f0.w);
    const bool sphere = SPHERES && __float_as_int(f1.w) == 1;

    const float3 p = o + h.t * d;  // ray.h:19
    float3 normal;
    if (sphere) {
        const PrimGeom g = load_prim(sc, h.prim);
        normal = (p - g.v0) / g.e1.x;  // sphere.h:33
    } else {
        normal = f3(f0.x, f0.y, f0.z);  // triangle.h:103 normalize(cross(e1, e2)) — geometric, never flipped towards the ray
    }

    const float4 m0 = __ldg(&sc.mats[mat * 3 + 0]);
    const float4 m1 = __ldg(&sc.mats[mat * 3 + 1]);
    const float4 m2 = __ldg(&sc.mats[mat * 3 + 2]);
    const int type = __float_as_int(m0.x);
    const float3 base = f3(m0.y, m0.z, m0.w);
    const float3 emis = f3(m1.x, m1.y, m1.z);
    const int base_tex = __float_as_int(m1.w);
    const int emis_tex = __float_as_int(m2.x);
    float tu = 0.f, tv = 0.f;
    if (!sphere && (base_tex >= 0 || emis_tex >= 0)) {
        const float4 s0 = PT_LDSHADE(&sc.shade[h.prim * 2 + 0]);
        const float4 s1 = PT_LDSHADE(&sc.shade[h.prim * 2 + 1]);
        tu = (1 - h.u - h.v) * s0.x + h.u * s0.z + h.v * s1.x;  // triangle.h:106-107
        tv = (1 - h.u - h.v) * s0.y + h.u * s0.w + h.v * s1.y;
    }

    if (!RTOW || type == 4 /* PT_MAT_UNIVERSAL */) {
        // UniversalMaterial::emitted / scatter, material.h:52-86
        float3 emitted;
        if (emis_tex >= 0) emitted = texture_value(sc, emis_tex, tu, tv) * emis * 50;
        else emitted = emis * 50;
        if (emitted.x > 0.0001f || emitted.y > 0.0001f || emitted.z > 0.0001f) {
            contrib = att * emitted;  // camera.h:72-75
            return false;
        }
        // material.h:67-69: scatter draws a cosine direction that camera.h:65 then overwrites — only the two draws matter
        (void)rng_next(rng);
        (void)rng_next(rng);
        float3 attenuation = base;
        if (base_tex >= 0) attenuation = attenuation * texture_value(sc, base_tex, tu, tv);  // material.h:71-75

        // camera.h:62-66: mixture_pdf(light list, cosine).generate / .value — pdf.h:57-75
        // the onb of cosine_pdf (pdf.h:16, onb.h:8-13): per-triangle constants from the frames table; built here for spheres
        Onb uvw;
        if (sphere) uvw = make_onb(normal);
        else uvw.w = f3(f1.x, f1.y, f1.z);
        float3 dir;
        const bool have_lights = sc.n_lights > 0;
        const float pick = rng_uniform(rng);
        if (have_lights && pick < 0.5f) {
            // hitable_list::random (hitable_list.h:23-26) -> triangle::random (triangle.h:41-47)
            int index = (int)truncf((float)(rng_uniform(rng) * sc.light_pick_scale));
            const float4 l0 = __ldg(&sc.lights[index * 4 + 0]);
            const float4 l1 = __ldg(&sc.lights[index * 4 + 1]);
            const float4 l2 = __ldg(&sc.lights[index * 4 + 2]);
            float r1 = rng_uniform(rng);
            float r2 = rng_uniform(rng);
            float sqrt_r1 = sqrtf(r1);
            float3 random_point = (1 - sqrt_r1) * f3(l0.x, l0.y, l0.z) + (sqrt_r1 * (1 - r2)) * f3(l1.x, l1.y, l1.z) + (sqrt_r1 * r2) * f3(l2.x, l2.y, l2.z);
            dir = random_point - p;
        } else {
            if (!sphere) {
                const float4 f2 = PT_LDSHADE(&sc.frames[h.prim * 4 + 2]);
                const float4 f3_ = PT_LDSHADE(&sc.frames[h.prim * 4 + 3]);
                uvw.u = f3(f2.x, f2.y, f2.z);
                uvw.v = f3(f2.w, f3_.x, f3_.y);
            }
            dir = onb_local(uvw, random_cosine_direction(rng));  // cosine_pdf::generate, pdf.h:23-25
        }
        // hitable_list::pdf_value, hitable_list.h:16-22
        float light_pdf = 0.0f;
        for (int i = 0; i < sc.n_lights; i++) {
            const float4 l0 = __ldg(&sc.lights[i * 4 + 0]);
            const float4 l1 = __ldg(&sc.lights[i * 4 + 1]);
            const float4 l2 = __ldg(&sc.lights[i * 4 + 2]);
            if (COUNT) n_light += 1;
            light_pdf += sc.light_weight * light_pdf_value(&sc.lights[i * 4], f3(l0.x, l0.y, l0.z), f3(l1.x, l1.y, l1.z), f3(l2.x, l2.y, l2.z), l1.w, l2.w, l0.w, p, dir);
        }
        // cosine_pdf::value, pdf.h:19-22 (w = normalize(normal) once more, as onb's constructor does)
        const float3 ndir = normalize(dir);
        float cosine = dot(ndir, uvw.w);
        float cos_pdf = (cosine <= 0) ? 0 : div_pi(cosine);
        float pdf_value = 0.5f * light_pdf + 0.5f * cos_pdf;  // pdf.h:63-65
        // UniversalMaterial::scattering_pdf, material.h:88-91
        float cosine2 = dot(normal, ndir);
        float scattering_pdf = cosine2 < 0 ? 0 : div_pi(cosine2);
        const float3 num = attenuation * scattering_pdf;
        att = att * f3(div_often_zero(num.x, pdf_value), div_often_zero(num.y, pdf_value), div_often_zero(num.z, pdf_value));  // camera.h:69
        o = p;
        d = dir;
        return true;
    }
    if (RTOW) {
        if (type == 3 /* DIFFUSE_LIGHT */) {  // material.h:210-217 + builder-defined glue
            contrib = att * emis;
            return false;
        }
        float3 attenuation, sdir;
        bool scattered = true;
        if (type == 0 /* LAMBERTIAN */) {  // material.h:113-124
            sdir = normal + random_in_unit_sphere(rng);
            if (fabsf(sdir.x) < 1e-8 && fabsf(sdir.y) < 1e-8 && fabsf(sdir.z) < 1e-8) sdir = normal;
            attenuation = base;
        } else if (type == 1 /* METAL */) {  // material.h:133-140
            float fuzz = m2.y < 1 ? m2.y : 1;
            float3 reflected = reflect3(normalize(d), normal);
            sdir = reflected + fuzz * random_in_unit_sphere(rng);
            attenuation = base;
            scattered = dot(sdir, normal) > 0;
        } else {  // DIELECTRIC, material.h:149-179
            const float ir = m2.z;
            float3 outward_normal;
            float3 reflected = reflect3(d, normal);
            float ni_over_nt;
            attenuation = f3(1.0f, 1.0f, 1.0f);
            float3 refracted = f3(0.f, 0.f, 0.f);
            float reflect_prob;
            float cosine;
            if (dot(d, normal) > 0.0f) {
                outward_normal = -normal;
                ni_over_nt = ir;
                cosine = dot(d, normal) / length(d);
                cosine = sqrtf(1.0f - ir * ir * (1 - cosine * cosine));
            } else {
                outward_normal = normal;
                ni_over_nt = 1.0f / ir;
                cosine = -dot(d, normal) / length(d);
            }
            if (refract3(d, outward_normal, ni_over_nt, refracted)) reflect_prob = schlick(cosine, ir);
            else reflect_prob = 1.0f;
            if (rng_uniform(rng) < reflect_prob) sdir = reflected;
            else sdir = refracted;
        }
        if (!scattered) {
            contrib = f3(0.f, 0.f, 0.f);
            return false;
        }
        att = att * attenuation;
        o = p;
        d = sdir;
        return true;
    }
    contrib = f3(0.f, 0.f, 0.f);
    return false;
}

}  // namespace ptc

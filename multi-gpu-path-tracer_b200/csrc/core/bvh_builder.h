// bvh_builder.h — host-side SAH BVH2 build, flattened to the device layout.
//
// Replaces the reference's single-GPU-thread device build (src/bvh.h:20-176, launched as
// create_world<<<1,1>>>, src/DevicePathTracer.h:134-146).  That build is not reproduced: the
// reference's traversal returns the exact closest hit over all triangles whatever the tree
// looks like (its boxes ignore ray_t, src/aabb.h:38-66), so only the order of exactly-equal
// hits depends on the tree.  What IS different on purpose:
//   * boxes are tight (the reference's all contain the world origin, src/bvh.h:98),
//   * full SAH (sweep for small nodes, 32 bins otherwise) instead of 3 planes per axis,
//   * nodes carry both children's boxes so one 64-byte fetch decides both subtrees,
//   * leaves hold at most `leaf_max` primitives, stored contiguously in leaf order.
#pragma once

#include <cstdint>
#include <vector>

namespace ptc {

// One 64-byte node = four 16-byte quads, fetched with 4 x LDG.128:
//   q[0] = (l.min.x, l.max.x, r.min.x, r.max.x)
//   q[1] = (l.min.y, l.max.y, r.min.y, r.max.y)
//   q[2] = (l.min.z, l.max.z, r.min.z, r.max.z)
//   q[3] = (left ref, right ref, unused, unused) as int32
// child ref >= 0: index of an inner node; ref < 0: leaf, ~ref = (first_prim << 4) | count, 1 <= count <= 8.
// An absent child (single-leaf scenes) has all six planes at +inf and refers to a leaf of ZERO primitives: entering it tests nothing.
struct alignas(64) FlatNode {
    float bx[4];
    float by[4];
    float bz[4];
    int32_t left, right, pad0, pad1;
};
static_assert(sizeof(FlatNode) == 64, "FlatNode must be one 64-byte line");

// The same two-child node in 32 bytes (two 16-byte quads, one sector): box planes as 15-bit integers on a uniform grid over
// the scene's bounds, two planes per word (low half = min, high half = max):
//   q[0] = (l.x, r.x, l.y, r.y)   q[1] = (l.z, r.z, left ref, right ref)
// Planes are rounded outwards and moved one more cell outwards, so a quantised box always contains the float box; the
// kernel tests rays in grid coordinates (trav_begin_grid).  Node i of this array is node i of `nodes`.
struct alignas(32) QuantNode {
    uint32_t lx, rx, ly, ry;
    uint32_t lz, rz;
    int32_t left, right;
};
static_assert(sizeof(QuantNode) == 32, "QuantNode must be 32 bytes");

struct QuantGrid {
    float lo[3];
    float scale[3];  // g = (x - lo) * scale
};

// Wide node: up to four children, boxes as SoA so that one lane tests all four with independent instruction streams.
// 128 bytes = eight 16-byte quads; the kernel fetches the first seven.
//   q[0] = lo.x of children 0..3, q[1] = hi.x, q[2] = lo.y, q[3] = hi.y, q[4] = lo.z, q[5] = hi.z, q[6] = child refs (int32)
// Built by collapsing the SAH BVH2: a node's inner children are opened (largest surface area first) until it has four
// children or only leaves.  Unused slots have all planes at +inf and refer to a leaf of zero primitives.
struct alignas(128) FlatNode4 {
    float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];
    int32_t ref[4];
    int32_t pad[4];
};
static_assert(sizeof(FlatNode4) == 128, "FlatNode4 must be 128 bytes");

constexpr int kLeafCountBits = 4;   // ~ref = (first << 4) | count: the complement IS the kernel's leaf cursor (position << 4 | primitives left)
constexpr int kMaxLeafPrims = 8;
inline int32_t leaf_first(int32_t ref) { return (~ref) >> kLeafCountBits; }
inline int32_t leaf_count(int32_t ref) { return (~ref) & ((1 << kLeafCountBits) - 1); }
constexpr int kMaxTraversalDepth = 48;              // BVH2 depth bound; the device stack (64 entries) covers it for both node widths

struct PrimBounds {
    float lo[3], hi[3];
};

struct BvhBuildOptions {
    int leaf_max = 4;
    float traversal_cost = 1.0f;
    float intersect_cost = 1.2f;
    int bins = 32;
};

struct BvhBuildResult {
    std::vector<FlatNode> nodes;       // nodes[0] is the root
    std::vector<FlatNode4> nodes4;     // the same tree collapsed to four-wide nodes, nodes4[0] is the root
    uint32_t depth4 = 0;               // depth of the collapsed tree
    uint32_t stack4 = 0;               // worst-case traversal stack entries of the collapsed tree (3 per level)
    std::vector<int32_t> prim_order;   // leaf-order position -> input primitive index
    uint32_t n_leaves = 0, depth = 0;
    double sah_cost = 0.0;
    double build_ms = 0.0;
};

// `bounds` must already include whatever conservative padding the caller wants.
BvhBuildResult build_bvh(const std::vector<PrimBounds> &bounds, const BvhBuildOptions &opt);

// Quantises the flat nodes; returns the grid.
QuantGrid quantise_nodes(const std::vector<FlatNode> &nodes, std::vector<QuantNode> &out);

// Checks, with the float arithmetic the kernel uses for ray origins, that every quantised box contains its float box with
// at least half a cell to spare; `inflation` (may be null) receives the mean over the leaves of (surface area of the
// quantised box / surface area of the float box), the price of the coarser planes.  Returns "" if valid.
const char *validate_quantised(const std::vector<FlatNode> &nodes, const std::vector<QuantNode> &q, const QuantGrid &g, double *inflation);

// Host emulation (same float operations, fmaf where the kernel uses fma) of the two slab tests of trav_node_step on `n_rays`
// pseudo-random rays — origins inside and outside the scene, on box planes, directions with zero / tiny / axis-parallel
// components, finite and infinite far limits — against every child box of up to `max_nodes` nodes: a box accepted on float
// planes must be accepted on quantised planes.  counts = {tests, accepted on float planes, accepted on quantised planes}.
const char *check_quantised_walk(const std::vector<FlatNode> &nodes, const std::vector<QuantNode> &q, const QuantGrid &g, uint32_t n_rays, uint32_t seed,
                                 uint32_t max_nodes, uint64_t counts[3]);

// Host restatement of the kernel's resumable walk (trav_begin / trav_begin_grid, trav_node_step with the held leaf and the
// sentinel stack, trav_prim_step2 with the tie rule) on triangle soups: for `n_rays` pseudo-random rays — camera-like,
// bounce-like (origin on a triangle), axis-parallel — the closest hit found by the walk over float planes, by the walk over
// quantised planes and by testing every triangle must be the same triangle at the same t.  `tri_pos` = 9 floats per
// triangle in INPUT order (bvh.prim_order maps leaf order to it).  counts = {rays, rays that hit, node steps float, node
// steps quantised}.  Returns "" if all three agree for every ray.
const char *check_walks(const BvhBuildResult &bvh, const std::vector<QuantNode> &q, const QuantGrid &g, const float *tri_pos, size_t n_tris, uint32_t n_rays,
                        uint32_t seed, uint64_t counts[4]);

// Structural validation used by the tests: every primitive in exactly one leaf, child boxes
// enclose their primitives, refs in range, depth within the device stack. Returns "" if valid.
const char *validate_bvh(const BvhBuildResult &bvh, const std::vector<PrimBounds> &bounds);
const char *validate_bvh4(const BvhBuildResult &bvh, const std::vector<PrimBounds> &bounds);

}  // namespace ptc

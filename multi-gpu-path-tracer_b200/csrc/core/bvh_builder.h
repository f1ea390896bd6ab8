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
// child ref >= 0: index of an inner node; ref < 0: leaf, ~ref = (first_prim << 3) | (count - 1).
// An absent child (single-leaf scenes) has an inverted box (+inf, -inf) that no ray enters.
struct alignas(64) FlatNode {
    float bx[4];
    float by[4];
    float bz[4];
    int32_t left, right, pad0, pad1;
};
static_assert(sizeof(FlatNode) == 64, "FlatNode must be one 64-byte line");

constexpr int kLeafCountBits = 3;
constexpr int kMaxLeafPrims = 1 << kLeafCountBits;  // 8
constexpr int kMaxTraversalDepth = 48;              // device stack (per-thread local memory) holds this many entries

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
    std::vector<int32_t> prim_order;   // leaf-order position -> input primitive index
    uint32_t n_leaves = 0, depth = 0;
    double sah_cost = 0.0;
    double build_ms = 0.0;
};

// `bounds` must already include whatever conservative padding the caller wants.
BvhBuildResult build_bvh(const std::vector<PrimBounds> &bounds, const BvhBuildOptions &opt);

// Structural validation used by the tests: every primitive in exactly one leaf, child boxes
// enclose their primitives, refs in range, depth within the device stack. Returns "" if valid.
const char *validate_bvh(const BvhBuildResult &bvh, const std::vector<PrimBounds> &bounds);

}  // namespace ptc

// bvh_builder.cpp — see bvh_builder.h.
#include "bvh_builder.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <atomic>
#include <numeric>
#include <thread>

namespace ptc {

static inline float inf_f() { return std::numeric_limits<float>::infinity(); }
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
    }
    void grow(const float *l, const float *h) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); }
    }
    void grow(const Box &b) { grow(b.lo, b.hi); }
    float half_area() const {
        float e[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        if (e[0] < 0 || e[1] < 0 || e[2] < 0) return 0.f;
        return e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
    }
};

struct TmpNode {
    Box box;
    int32_t left = -1, right = -1;  // children (TmpNode indices) or -1
    int32_t first = 0, count = 0;   // range in `order`
};

struct Builder {
    const std::vector<PrimBounds> &pb;
    BvhBuildOptions opt;
    std::vector<int32_t> order;
    std::vector<float> cent;  // 3 per prim
    std::vector<TmpNode> tmp;
    uint32_t maxDepth = 0;

    Builder(const std::vector<PrimBounds> &b, const BvhBuildOptions &o) : pb(b), opt(o) {}

    Box rangeBox(int32_t first, int32_t count, Box *centBox) const {
        Box b;
        b.reset();
        if (centBox) centBox->reset();
        for (int32_t i = 0; i < count; i++) {
            int32_t id = order[(size_t)first + (size_t)i];
            b.grow(pb[(size_t)id].lo, pb[(size_t)id].hi);
            if (centBox) {
                const float *c = &cent[(size_t)id * 3];
                centBox->grow(c, c);
            }
        }
        return b;
    }

    // Returns split position in `order` (first < mid < first+count) or -1 for "make a leaf".
    int32_t findSplit(int32_t first, int32_t count, const Box &box, const Box &cbox, bool mustSplit, bool forceMedian) {
        const float parentArea = std::max(box.half_area(), 1e-30f);
        const float leafCost = opt.intersect_cost * (float)count;
        float bestCost = std::numeric_limits<float>::infinity();
        int bestAxis = -1;
        int32_t bestMid = -1;
        float bestPlane = 0.f;
        bool bestIsSweep = false;

        if (!forceMedian) {
            if (count <= 64) {
                // exact sweep over centroid-sorted order, all three axes
                std::vector<int32_t> ids(order.begin() + first, order.begin() + first + count);
                std::vector<float> sweepArea((size_t)count);
                for (int axis = 0; axis < 3; axis++) {
                    if (!(cbox.hi[axis] > cbox.lo[axis])) continue;
                    std::stable_sort(ids.begin(), ids.end(), [&](int32_t a, int32_t b) { return cent[(size_t)a * 3 + axis] < cent[(size_t)b * 3 + axis]; });
                    Box acc;
                    acc.reset();
                    for (int32_t i = count - 1; i > 0; i--) {
                        acc.grow(pb[(size_t)ids[(size_t)i]].lo, pb[(size_t)ids[(size_t)i]].hi);
                        sweepArea[(size_t)i] = acc.half_area();
                    }
                    acc.reset();
                    for (int32_t i = 1; i < count; i++) {
                        acc.grow(pb[(size_t)ids[(size_t)i - 1]].lo, pb[(size_t)ids[(size_t)i - 1]].hi);
                        float cost = opt.traversal_cost + opt.intersect_cost * (acc.half_area() * (float)i + sweepArea[(size_t)i] * (float)(count - i)) / parentArea;
                        if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestMid = i; bestIsSweep = true; }
                    }
                }
            } else {
                const int B = std::max(4, std::min(opt.bins, 64));
                for (int axis = 0; axis < 3; axis++) {
                    float lo = cbox.lo[axis], hi = cbox.hi[axis];
                    if (!(hi > lo)) continue;
                    Box bins[64];
                    int32_t cnt[64];
                    for (int b = 0; b < B; b++) { bins[b].reset(); cnt[b] = 0; }
                    float scale = (float)B / (hi - lo);
                    for (int32_t i = 0; i < count; i++) {
                        int32_t id = order[(size_t)first + (size_t)i];
                        int b = (int)((cent[(size_t)id * 3 + axis] - lo) * scale);
                        b = std::max(0, std::min(B - 1, b));
                        bins[b].grow(pb[(size_t)id].lo, pb[(size_t)id].hi);
                        cnt[b]++;
                    }
                    float rightArea[64];
                    int32_t rightCnt[64];
                    Box acc;
                    acc.reset();
                    int32_t c = 0;
                    for (int b = B - 1; b > 0; b--) {
                        acc.grow(bins[b]);
                        c += cnt[b];
                        rightArea[b] = acc.half_area();
                        rightCnt[b] = c;
                    }
                    acc.reset();
                    c = 0;
                    for (int b = 1; b < B; b++) {
                        acc.grow(bins[b - 1]);
                        c += cnt[b - 1];
                        if (c == 0 || rightCnt[b] == 0) continue;
                        float cost = opt.traversal_cost + opt.intersect_cost * (acc.half_area() * (float)c + rightArea[b] * (float)rightCnt[b]) / parentArea;
                        if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestPlane = lo + (float)b / scale; bestMid = c; bestIsSweep = false; }
                    }
                }
            }
        }

        bool wantLeaf = !mustSplit && (bestAxis < 0 || bestCost >= leafCost);
        if (wantLeaf) return -1;

        if (bestAxis >= 0 && bestIsSweep) {
            int axis = bestAxis;
            std::stable_sort(order.begin() + first, order.begin() + first + count,
                             [&](int32_t a, int32_t b) { return cent[(size_t)a * 3 + axis] < cent[(size_t)b * 3 + axis]; });
            return first + bestMid;
        }
        if (bestAxis >= 0) {
            int axis = bestAxis;
            float lo = cbox.lo[axis], hi = cbox.hi[axis];
            const int B = std::max(4, std::min(opt.bins, 64));
            float scale = (float)B / (hi - lo);
            int planeBin = (int)std::lround((bestPlane - lo) * scale);
            auto mid = std::partition(order.begin() + first, order.begin() + first + count, [&](int32_t id) {
                int b = (int)((cent[(size_t)id * 3 + axis] - lo) * scale);
                b = std::max(0, std::min(B - 1, b));
                return b < planeBin;
            });
            int32_t m = (int32_t)(mid - order.begin());
            if (m > first && m < first + count) return m;
        }
        // median fallback: largest centroid extent, split by count
        int axis = 0;
        for (int k = 1; k < 3; k++)
            if (cbox.hi[k] - cbox.lo[k] > cbox.hi[axis] - cbox.lo[axis]) axis = k;
        int32_t m = first + count / 2;
        std::nth_element(order.begin() + first, order.begin() + (m - 0), order.begin() + first + count,
                         [&](int32_t a, int32_t b) { return cent[(size_t)a * 3 + axis] < cent[(size_t)b * 3 + axis]; });
        return m;
    }

    struct Work { int32_t node; uint32_t depth; };

    // Splits nodes[w.node] if the SAH says so; children are appended to `nodes` and pushed on `stack`.
    void processNode(std::vector<TmpNode> &nodes, Work w, std::vector<Work> &stack, uint32_t &deepest) {
        const int leafMax = std::max(1, std::min(opt.leaf_max, kMaxLeafPrims));
        deepest = std::max(deepest, w.depth);
        int32_t first = nodes[(size_t)w.node].first, count = nodes[(size_t)w.node].count;
        if (count <= 1) return;
        Box cbox;
        Box box = rangeBox(first, count, &cbox);
        bool mustSplit = count > leafMax;
        // keep the tree shallow enough for the device stack: once deep, split by count
        uint32_t remaining = 1;
        while ((1u << remaining) * (uint32_t)leafMax < (uint32_t)count && remaining < 31) remaining++;
        bool forceMedian = w.depth + remaining + 2 >= (uint32_t)kMaxTraversalDepth;
        int32_t mid = findSplit(first, count, box, cbox, mustSplit, forceMedian);
        if (mid < 0) return;
        TmpNode l, r;
        l.first = first; l.count = mid - first;
        r.first = mid; r.count = first + count - mid;
        l.box = rangeBox(l.first, l.count, nullptr);
        r.box = rangeBox(r.first, r.count, nullptr);
        int32_t li = (int32_t)nodes.size();
        nodes.push_back(l);
        int32_t ri = (int32_t)nodes.size();
        nodes.push_back(r);
        nodes[(size_t)w.node].left = li;
        nodes[(size_t)w.node].right = ri;
        stack.push_back({ri, w.depth + 1});
        stack.push_back({li, w.depth + 1});
    }

    void run() {
        const size_t n = pb.size();
        order.resize(n);
        std::iota(order.begin(), order.end(), 0);
        cent.resize(n * 3);
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) cent[i * 3 + k] = 0.5f * (pb[i].lo[k] + pb[i].hi[k]);
        tmp.reserve(n ? 2 * n : 1);
        TmpNode root;
        root.first = 0;
        root.count = (int32_t)n;
        Box cb;
        root.box = rangeBox(0, (int32_t)n, &cb);
        tmp.push_back(root);

        // Phase 1 (serial): split the large nodes; subtrees of at most `grain` primitives are deferred.
        // Phase 2 (threads): deferred subtrees own disjoint ranges of `order`, so they build independently into
        // private node arrays that are appended afterwards.  The tree (and hence every image) does not depend on
        // the thread count: only the numbering of the temporary nodes does, and flattening renumbers them.
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned nthreads = n >= 200000 ? std::min(hw, 32u) : 1u;
        const int32_t grain = nthreads > 1 ? (int32_t)std::max<size_t>(4096, n / (8 * (size_t)nthreads)) : 0;
        std::vector<Work> stack, deferred;
        stack.push_back({0, 1});
        while (!stack.empty()) {
            Work w = stack.back();
            stack.pop_back();
            if (nthreads > 1 && tmp[(size_t)w.node].count <= grain) {
                deferred.push_back(w);
                continue;
            }
            processNode(tmp, w, stack, maxDepth);
        }
        if (deferred.empty()) return;
        std::vector<std::vector<TmpNode>> local(deferred.size());
        std::vector<uint32_t> localDepth(deferred.size(), 0);
        std::atomic<size_t> nextJob{0};
        auto worker = [&]() {
            for (;;) {
                size_t j = nextJob.fetch_add(1);
                if (j >= deferred.size()) return;
                std::vector<TmpNode> &nodes = local[j];
                nodes.push_back(tmp[(size_t)deferred[j].node]);
                std::vector<Work> st;
                st.push_back({0, deferred[j].depth});
                while (!st.empty()) {
                    Work w = st.back();
                    st.pop_back();
                    processNode(nodes, w, st, localDepth[j]);
                }
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nthreads; t++) pool.emplace_back(worker);
        worker();
        for (auto &t : pool) t.join();
        for (size_t j = 0; j < deferred.size(); j++) {
            maxDepth = std::max(maxDepth, localDepth[j]);
            const std::vector<TmpNode> &nodes = local[j];
            const int32_t base = (int32_t)tmp.size() - 1;  // local index k >= 1 becomes base + k
            TmpNode &top = tmp[(size_t)deferred[j].node];
            if (nodes[0].left >= 0) {
                top.left = base + nodes[0].left;
                top.right = base + nodes[0].right;
            }
            for (size_t k = 1; k < nodes.size(); k++) {
                TmpNode c = nodes[k];
                if (c.left >= 0) { c.left += base; c.right += base; }
                tmp.push_back(c);
            }
        }
    }
};

inline int32_t leaf_ref(int32_t first, int32_t count) { return ~((first << kLeafCountBits) | count); }

}  // namespace

BvhBuildResult build_bvh(const std::vector<PrimBounds> &bounds, const BvhBuildOptions &opt) {
    auto t0 = std::chrono::high_resolution_clock::now();
    BvhBuildResult out;
    const float inf = std::numeric_limits<float>::infinity();
    auto emptyChild = [&](FlatNode &n, int side) {
        // all six planes at +inf, and (below) a leaf reference of ZERO primitives: a (+inf, -inf) box is ACCEPTED by a min / max slab test
        //  and even this one is while nothing has been hit yet (far = FLT_MAX * slack = +inf), so what makes an absent child
        // harmless is its empty leaf, not its box
        n.bx[side * 2] = n.by[side * 2] = n.bz[side * 2] = inf;
        n.bx[side * 2 + 1] = n.by[side * 2 + 1] = n.bz[side * 2 + 1] = inf;
    };
    if (bounds.empty()) {
        FlatNode n{};
        emptyChild(n, 0);
        emptyChild(n, 1);
        n.left = n.right = leaf_ref(0, 0);  // a leaf of ZERO primitives: whatever the slab test says, the walk tests nothing (a zero cursor is never 'held')
        out.nodes.push_back(n);
        out.depth = 1;
        FlatNode4 n4{};
        for (int c = 0; c < 4; c++) {
            n4.lox[c] = n4.loy[c] = n4.loz[c] = inf;
            n4.hix[c] = n4.hiy[c] = n4.hiz[c] = inf;
            n4.ref[c] = leaf_ref(0, 0);
        }
        out.nodes4.push_back(n4);
        out.depth4 = 1;
        out.stack4 = 5;
        return out;
    }
    Builder b(bounds, opt);
    b.run();
    out.prim_order = b.order;
    out.depth = b.maxDepth;

    // flatten: inner TmpNodes -> FlatNodes in DFS order; leaves become child refs
    std::vector<int32_t> flatIndex(b.tmp.size(), -1);
    struct Item { int32_t tmp; };
    std::vector<int32_t> dfs;
    if (b.tmp[0].left < 0) {
        // single-leaf scene (possibly more than leaf_max prims if they are all coincident): chain it
        FlatNode n{};
        const TmpNode &t = b.tmp[0];
        int32_t cnt = std::min<int32_t>(t.count, kMaxLeafPrims);
        n.bx[0] = t.box.lo[0]; n.bx[1] = t.box.hi[0];
        n.by[0] = t.box.lo[1]; n.by[1] = t.box.hi[1];
        n.bz[0] = t.box.lo[2]; n.bz[1] = t.box.hi[2];
        emptyChild(n, 1);
        n.left = leaf_ref(t.first, cnt);
        n.right = leaf_ref(0, 0);  // absent child: a leaf of zero primitives
        out.nodes.push_back(n);
        out.n_leaves = 1;
        out.sah_cost = opt.intersect_cost * (double)t.count;
    } else {
        dfs.push_back(0);
        while (!dfs.empty()) {
            int32_t ti = dfs.back();
            dfs.pop_back();
            flatIndex[(size_t)ti] = (int32_t)out.nodes.size();
            out.nodes.emplace_back();
            const TmpNode &t = b.tmp[(size_t)ti];
            if (b.tmp[(size_t)t.right].left >= 0) dfs.push_back(t.right);
            if (b.tmp[(size_t)t.left].left >= 0) dfs.push_back(t.left);
        }
        const double rootArea = std::max((double)b.tmp[0].box.half_area(), 1e-30);
        for (size_t ti = 0; ti < b.tmp.size(); ti++) {
            if (flatIndex[ti] < 0) continue;
            const TmpNode &t = b.tmp[ti];
            FlatNode &n = out.nodes[(size_t)flatIndex[ti]];
            const TmpNode *ch[2] = {&b.tmp[(size_t)t.left], &b.tmp[(size_t)t.right]};
            int32_t ci[2] = {t.left, t.right};
            int32_t refs[2];
            for (int s = 0; s < 2; s++) {
                n.bx[s * 2] = ch[s]->box.lo[0]; n.bx[s * 2 + 1] = ch[s]->box.hi[0];
                n.by[s * 2] = ch[s]->box.lo[1]; n.by[s * 2 + 1] = ch[s]->box.hi[1];
                n.bz[s * 2] = ch[s]->box.lo[2]; n.bz[s * 2 + 1] = ch[s]->box.hi[2];
                if (ch[s]->left >= 0) {
                    refs[s] = flatIndex[(size_t)ci[s]];
                    out.sah_cost += opt.traversal_cost * (double)ch[s]->box.half_area() / rootArea;
                } else {
                    refs[s] = leaf_ref(ch[s]->first, ch[s]->count);
                    out.n_leaves++;
                    out.sah_cost += opt.intersect_cost * (double)ch[s]->count * (double)ch[s]->box.half_area() / rootArea;
                }
            }
            n.left = refs[0];
            n.right = refs[1];
            n.pad0 = n.pad1 = 0;
        }
        out.sah_cost += opt.traversal_cost;
    }
    // ---- collapse to four-wide nodes ----
    {
        auto emptySlot = [&](FlatNode4 &n, int c) {
            n.lox[c] = n.loy[c] = n.loz[c] = inf;
            n.hix[c] = n.hiy[c] = n.hiz[c] = inf;
            n.ref[c] = leaf_ref(0, 0);
        };
        auto setSlot = [&](FlatNode4 &n, int c, const Box &bx, int32_t ref) {
            n.lox[c] = bx.lo[0]; n.hix[c] = bx.hi[0];
            n.loy[c] = bx.lo[1]; n.hiy[c] = bx.hi[1];
            n.loz[c] = bx.lo[2]; n.hiz[c] = bx.hi[2];
            n.ref[c] = ref;
        };
        if (b.tmp[0].left < 0) {
            FlatNode4 n{};
            for (int c = 0; c < 4; c++) emptySlot(n, c);
            const TmpNode &t = b.tmp[0];
            setSlot(n, 0, t.box, leaf_ref(t.first, std::min<int32_t>(t.count, kMaxLeafPrims)));
            out.nodes4.push_back(n);
            out.depth4 = 1;
        } else {
            struct W { int32_t tmp; int32_t flat; uint32_t depth; };
            std::vector<W> st;
            out.nodes4.emplace_back();
            st.push_back({0, 0, 1});
            while (!st.empty()) {
                W w = st.back();
                st.pop_back();
                out.depth4 = std::max(out.depth4, w.depth);
                int32_t kids[4];
                int nk = 2;
                kids[0] = b.tmp[(size_t)w.tmp].left;
                kids[1] = b.tmp[(size_t)w.tmp].right;
                while (nk < 4) {
                    int best = -1;
                    float bestArea = -1.f;
                    for (int k = 0; k < nk; k++) {
                        const TmpNode &c = b.tmp[(size_t)kids[k]];
                        if (c.left >= 0 && c.box.half_area() > bestArea) { bestArea = c.box.half_area(); best = k; }
                    }
                    if (best < 0) break;
                    const TmpNode &c = b.tmp[(size_t)kids[best]];
                    kids[best] = c.left;
                    kids[nk++] = c.right;
                }
                FlatNode4 n{};
                for (int c = 0; c < 4; c++) emptySlot(n, c);
                // children that are inner nodes get consecutive flat indices; visit them in slot order (DFS-ish layout)
                for (int k = nk - 1; k >= 0; k--) {
                    const TmpNode &c = b.tmp[(size_t)kids[k]];
                    if (c.left >= 0) {
                        int32_t fi = (int32_t)out.nodes4.size();
                        out.nodes4.emplace_back();
                        setSlot(n, k, c.box, fi);
                        st.push_back({kids[k], fi, w.depth + 1});
                    } else {
                        setSlot(n, k, c.box, leaf_ref(c.first, c.count));
                    }
                }
                out.nodes4[(size_t)w.flat] = n;
            }
        }
        out.stack4 = 3 * out.depth4 + 2;
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    out.build_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    return out;
}

QuantGrid quantise_nodes(const std::vector<FlatNode> &nodes, std::vector<QuantNode> &out) {
    QuantGrid g{};
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (const FlatNode &n : nodes)
        for (int side = 0; side < 2; side++) {
            const float b[3][2] = {{n.bx[side * 2], n.bx[side * 2 + 1]}, {n.by[side * 2], n.by[side * 2 + 1]}, {n.bz[side * 2], n.bz[side * 2 + 1]}};
            if (!(b[0][0] < inf_f())) continue;  // absent child (all planes at +inf)
            for (int k = 0; k < 3; k++) {
                if (std::isfinite(b[k][0])) lo[k] = std::min(lo[k], (double)b[k][0]);
                if (std::isfinite(b[k][1])) hi[k] = std::max(hi[k], (double)b[k][1]);
            }
        }
    double scale[3];
    for (int k = 0; k < 3; k++) {
        if (!(lo[k] <= hi[k])) lo[k] = hi[k] = 0.0;
        const double ext = hi[k] - lo[k];
        // cells 1 .. 32766 span the scene, cell 0 and 32767 are the outward margin
        float sc = ext > 0 ? (float)(32765.0 / ext) : 1.0f;
        if (!std::isfinite(sc) || sc <= 0.f) sc = 1.0f;
        g.scale[k] = sc;
        g.lo[k] = (float)(lo[k] - 1.0 / (double)sc);
        // the float grid origin may sit a hair above the exact one: planes are computed from the float values below, so
        // the outward rounding stays valid for what the kernel evaluates
        scale[k] = (double)sc;
    }
    auto plane = [&](float x, int k, bool upper) -> uint32_t {
        if (!std::isfinite(x)) return upper ? 32767u : 0u;
        const double c = ((double)x - (double)g.lo[k]) * scale[k];
        double q = upper ? std::ceil(c) + 1.0 : std::floor(c) - 1.0;
        if (q < 0.0) q = 0.0;
        if (q > 32767.0) q = 32767.0;
        return (uint32_t)q;
    };
    out.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); i++) {
        const FlatNode &n = nodes[i];
        QuantNode &q = out[i];
        auto pack = [&](const float *b, int side, int k) -> uint32_t {
            if (!(b[side * 2] < inf_f())) return 0x0000u | (32767u);  // absent child: min 32767, max 0 — near > far for every ray
            return plane(b[side * 2], k, false) | (plane(b[side * 2 + 1], k, true) << 16);
        };
        q.lx = pack(n.bx, 0, 0); q.rx = pack(n.bx, 1, 0);
        q.ly = pack(n.by, 0, 1); q.ry = pack(n.by, 1, 1);
        q.lz = pack(n.bz, 0, 2); q.rz = pack(n.bz, 1, 2);
        q.left = n.left;
        q.right = n.right;
    }
    return g;
}

const char *validate_quantised(const std::vector<FlatNode> &nodes, const std::vector<QuantNode> &q, const QuantGrid &g, double *inflation) {
    if (q.size() != nodes.size()) return "quantised node count differs";
    double areaQ = 0, areaF = 0;
    for (size_t i = 0; i < nodes.size(); i++) {
        const FlatNode &n = nodes[i];
        if (q[i].left != n.left || q[i].right != n.right) return "quantised node references differ";
        const float *fb[3] = {n.bx, n.by, n.bz};
        const uint32_t qb[3][2] = {{q[i].lx, q[i].rx}, {q[i].ly, q[i].ry}, {q[i].lz, q[i].rz}};
        for (int side = 0; side < 2; side++) {
            if (!(n.bx[side * 2] < inf_f())) continue;  // absent child
            double eq[3], ef[3];
            for (int k = 0; k < 3; k++) {
                const float lo = fb[k][side * 2], hi = fb[k][side * 2 + 1];
                const uint32_t w = qb[k][side];
                const float qlo = (float)(w & 0xffffu), qhi = (float)(w >> 16);
                if ((w & 0x8000u) || (w >> 31)) return "quantised plane exceeds 15 bits";
                // the kernel's float evaluation of a point on the float box's planes
                const float glo = (lo - g.lo[k]) * g.scale[k], ghi = (hi - g.lo[k]) * g.scale[k];
                if (std::isfinite(lo) && !(qlo + 0.5f <= glo)) return "quantised lower plane is not below the box";
                if (std::isfinite(hi) && !(qhi - 0.5f >= ghi)) return "quantised upper plane is not above the box";
                eq[k] = (double)qhi - (double)qlo;
                ef[k] = std::isfinite(lo) && std::isfinite(hi) ? ((double)hi - (double)lo) * (double)g.scale[k] : eq[k];
            }
            const int32_t ref = side ? n.right : n.left;
            if (ref < 0) {  // mean over the leaves of (quantised area / float area)
                const double aq = eq[0] * eq[1] + eq[1] * eq[2] + eq[2] * eq[0], af = ef[0] * ef[1] + ef[1] * ef[2] + ef[2] * ef[0];
                if (af > 0) {
                    areaQ += aq / af;
                    areaF += 1.0;
                }
            }
        }
    }
    if (inflation) *inflation = areaF > 0 ? areaQ / areaF : 1.0;
    return "";
}

namespace {
struct Lcg {
    uint64_t s;
    uint32_t next() {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        return (uint32_t)(s >> 32);
    }
    float unit() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }
};
inline float slab_inv1(float d) {  // slab_inverse, pt_device.cuh
    const float kTiny = 8.271806125530277e-25f;
    const float c = std::fabs(d) < kTiny ? std::copysign(kTiny, d) : d;
    return 1.0f / c;
}
}  // namespace

const char *check_quantised_walk(const std::vector<FlatNode> &nodes, const std::vector<QuantNode> &q, const QuantGrid &g, uint32_t n_rays, uint32_t seed,
                                 uint32_t max_nodes, uint64_t counts[3]) {
    counts[0] = counts[1] = counts[2] = 0;
    if (nodes.empty() || q.size() != nodes.size()) return "no nodes";
    const float kSlack = 1.0000004f, tmin = 0.001f;
    Lcg rng{0x9e3779b97f4a7c15ull ^ seed};
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        lo[k] = g.lo[k];
        hi[k] = g.lo[k] + 32767.0f / g.scale[k];
    }
    const size_t stride = std::max<size_t>(1, nodes.size() / std::max<uint32_t>(1, max_nodes));
    for (uint32_t r = 0; r < n_rays; r++) {
        float o[3], d[3];
        const uint32_t kind = rng.next() % 10;
        const size_t pickIdx = rng.next() % nodes.size();
        const FlatNode &pick = nodes[pickIdx];
        for (int k = 0; k < 3; k++) {
            const float ext = hi[k] - lo[k];
            o[k] = lo[k] - 0.25f * ext + 1.5f * ext * rng.unit();
            d[k] = 2.0f * rng.unit() - 1.0f;
        }
        if (kind == 1) o[rng.next() % 3] = pick.bx[rng.next() % 4];                       // origin coordinate taken from a box plane
        if (kind == 2) { o[0] = pick.bx[0]; o[1] = pick.by[1]; o[2] = pick.bz[2]; }          // origin on a box corner
        if (kind == 3) d[rng.next() % 3] = 0.0f;                                            // axis-parallel in one component
        if (kind == 4) { int a = rng.next() % 3; d[(a + 1) % 3] = 0.0f; d[(a + 2) % 3] = -0.0f; }  // axis-parallel ray
        if (kind == 5) d[rng.next() % 3] = 1e-30f * (rng.unit() - 0.5f);                    // tiny component (clamped by slab_inverse)
        if (kind == 6) for (int k = 0; k < 3; k++) d[k] *= 1000.0f;                         // long unnormalised direction (light sampling)
        if (kind >= 7 && pick.bx[0] <= pick.bx[1]) {                                        // aimed at the picked node's left box: inside, on a face, on an edge
            const float *pb3[3] = {pick.bx, pick.by, pick.bz};
            for (int k = 0; k < 3; k++) {
                const uint32_t where = kind == 7 ? 2u : rng.next() % 3;                       // 0 = min plane, 1 = max plane, 2 = inside
                const float target = where == 0 ? pb3[k][0] : where == 1 ? pb3[k][1] : pb3[k][0] + (pb3[k][1] - pb3[k][0]) * rng.unit();
                d[k] = target - o[k];
            }
        }
        const float best = (rng.next() & 1) ? FLT_MAX : 4000.0f * rng.unit();
        // float planes: t = b * inv + (-o * inv)
        float inv[3], oinv[3], ginv[3], goinv[3];
        bool gneg[3];
        for (int k = 0; k < 3; k++) {
            inv[k] = slab_inv1(d[k]);
            oinv[k] = -o[k] * inv[k];
            const float og = (o[k] - g.lo[k]) * g.scale[k];
            ginv[k] = slab_inv1(d[k] * g.scale[k]);
            goinv[k] = -(32768.0f + og) * ginv[k];
            gneg[k] = ginv[k] < 0.f;
        }
        auto test_node = [&](size_t i) -> bool {
            const FlatNode &n = nodes[i];
            const float *fb[3] = {n.bx, n.by, n.bz};
            const uint32_t qb[3][2] = {{q[i].lx, q[i].rx}, {q[i].ly, q[i].ry}, {q[i].lz, q[i].rz}};
            for (int side = 0; side < 2; side++) {
                if (!(n.bx[side * 2] <= n.bx[side * 2 + 1])) continue;
                float tn = tmin, tf = best, qn = tmin, qf = best;
                for (int k = 0; k < 3; k++) {
                    const float t0 = std::fmaf(fb[k][side * 2], inv[k], oinv[k]), t1 = std::fmaf(fb[k][side * 2 + 1], inv[k], oinv[k]);
                    tn = std::fmax(tn, std::fmin(t0, t1));
                    tf = std::fmin(tf, std::fmax(t0, t1));
                    const float plo = 32768.0f + (float)(qb[k][side] & 0xffffu), phi = 32768.0f + (float)(qb[k][side] >> 16);
                    qn = std::fmax(qn, std::fmaf(gneg[k] ? phi : plo, ginv[k], goinv[k]));
                    qf = std::fmin(qf, std::fmaf(gneg[k] ? plo : phi, ginv[k], goinv[k]));
                }
                const bool hf = tn <= tf * kSlack, hq = qn <= qf * kSlack;
                counts[0]++;
                counts[1] += hf;
                counts[2] += hq;
                if (hf && !hq) return false;
            }
            return true;
        };
        bool ok = test_node(pickIdx);  // the node the ray was aimed at, then a strided sample of all nodes
        for (size_t i = (size_t)(rng.next() % stride); ok && i < nodes.size(); i += stride) ok = test_node(i);
        if (!ok) return "a box accepted on float planes was rejected on quantised planes";
    }
    return "";
}

namespace {
struct HostHit {
    float t = FLT_MAX, u = 0.f, v = 0.f;
    int32_t prim = -1;  // leaf-order position
};
// Moeller-Trumbore with the kernel's acceptance rules (strict range, reject |det| < 1e-8, barycentric bounds)
inline bool host_triangle(const float *p, const float o[3], const float d[3], float tmin, float &t, float &u, float &v) {
    const float e1[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]}, e2[3] = {p[6] - p[0], p[7] - p[1], p[8] - p[2]};
    const float pv[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
    const float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
    const float inv = 1.0f / det;
    const float tv[3] = {o[0] - p[0], o[1] - p[1], o[2] - p[2]};
    u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv;
    const float qv[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
    v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv;
    t = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv;
    const bool reject = (std::fabs(det) < 1.00000008274037e-08f) | (u < 0) | (u > 1) | (v < 0) | (u + v > 1);
    return !reject & (t > tmin);
}
inline void host_accept(HostHit &best, float t, float u, float v, int32_t k) {
    if ((t < best.t) | ((t == best.t) & (k < best.prim))) best = HostHit{t, u, v, k};
}
}  // namespace

const char *check_walks(const BvhBuildResult &bvh, const std::vector<QuantNode> &q, const QuantGrid &g, const float *tri_pos, size_t n_tris, uint32_t n_rays,
                        uint32_t seed, uint64_t counts[4]) {
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    if (bvh.nodes.empty() || q.size() != bvh.nodes.size() || bvh.prim_order.size() != n_tris || n_tris == 0) return "nothing to walk";
    const int32_t kDone = (int32_t)0x80000000;
    const float kSlack = 1.0000004f, tmin = 0.001f;
    Lcg rng{0x51ed270b9f1a3c47ull ^ seed};
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        lo[k] = g.lo[k];
        hi[k] = g.lo[k] + 32767.0f / g.scale[k];
    }
    auto tri = [&](int32_t leafPos) { return tri_pos + (size_t)bvh.prim_order[(size_t)leafPos] * 9; };
    for (uint32_t r = 0; r < n_rays; r++) {
        float o[3], d[3];
        const uint32_t kind = rng.next() % 4;
        const float *aim = tri((int32_t)(rng.next() % n_tris));
        float bu = rng.unit(), bv = rng.unit();
        if (bu + bv > 1.f) { bu = 1.f - bu; bv = 1.f - bv; }
        float target[3];
        for (int k = 0; k < 3; k++) target[k] = aim[k] + bu * (aim[3 + k] - aim[k]) + bv * (aim[6 + k] - aim[k]);
        if (kind == 0) {  // camera-like: from outside or inside the scene towards a point on a triangle
            for (int k = 0; k < 3; k++) o[k] = lo[k] - 0.2f * (hi[k] - lo[k]) + 1.4f * (hi[k] - lo[k]) * rng.unit();
            for (int k = 0; k < 3; k++) d[k] = target[k] - o[k];
        } else if (kind == 1) {  // bounce-like: origin exactly on a triangle (self-intersection just above tmin is part of the game)
            const float *from = tri((int32_t)(rng.next() % n_tris));
            float cu = rng.unit(), cv = rng.unit();
            if (cu + cv > 1.f) { cu = 1.f - cu; cv = 1.f - cv; }
            for (int k = 0; k < 3; k++) o[k] = from[k] + cu * (from[3 + k] - from[k]) + cv * (from[6 + k] - from[k]);
            for (int k = 0; k < 3; k++) d[k] = target[k] - o[k];
        } else if (kind == 2) {  // towards a vertex or along an edge: ties between neighbouring triangles
            for (int k = 0; k < 3; k++) o[k] = lo[k] + (hi[k] - lo[k]) * rng.unit();
            const int corner = (int)(rng.next() % 3);
            for (int k = 0; k < 3; k++) d[k] = aim[corner * 3 + k] - o[k];
        } else {  // axis-parallel through a triangle point
            const int a = (int)(rng.next() % 3);
            for (int k = 0; k < 3; k++) { o[k] = target[k]; d[k] = 0.f; }
            o[a] = lo[a] - 1.0f;
            d[a] = 1.0f;
        }
        // 1. every triangle
        HostHit brute;
        for (int32_t k = 0; k < (int32_t)n_tris; k++) {
            float t, u, v;
            if (host_triangle(tri(k), o, d, tmin, t, u, v)) host_accept(brute, t, u, v, k);
        }
        // 2. / 3. the walk on float planes and on quantised planes
        HostHit found[2];
        for (int mode = 0; mode < 2; mode++) {
            float inv[3], oinv[3];
            uint32_t nearLow[3];  // quantised: does the ray reach the min plane first?
            for (int k = 0; k < 3; k++) {
                if (mode == 0) {
                    inv[k] = slab_inv1(d[k]);
                    oinv[k] = -o[k] * inv[k];
                } else {
                    const float og = (o[k] - g.lo[k]) * g.scale[k];
                    inv[k] = slab_inv1(d[k] * g.scale[k]);
                    oinv[k] = -(32768.0f + og) * inv[k];
                }
                nearLow[k] = !(inv[k] < 0.f);
            }
            HostHit best;
            int32_t stack[64];
            int32_t sp = 0, cur = 0, leaf = 0;
            stack[sp++] = kDone;
            auto held = [&]() { return (leaf & 15) != 0; };
            uint64_t steps = 0;
            while (!(cur == kDone && !held())) {
                if (++steps > 100000) return "the walk does not terminate";
                if (cur < 0 && !held()) return "the walk stalls: a leaf waits in cur while no leaf is held (neither vote of the kernel would step this lane)";
                if (cur >= 0) {  // trav_node_step
                    const FlatNode &n = bvh.nodes[(size_t)cur];
                    const QuantNode &qn = q[(size_t)cur];
                    float tn[2], tf[2];
                    for (int side = 0; side < 2; side++) {
                        tn[side] = tmin;
                        tf[side] = best.t;
                        const float *fb[3] = {n.bx, n.by, n.bz};
                        const uint32_t qb[3] = {side ? qn.rx : qn.lx, side ? qn.ry : qn.ly, side ? qn.rz : qn.lz};
                        for (int k = 0; k < 3; k++) {
                            if (mode == 0) {
                                const float t0 = std::fmaf(fb[k][side * 2], inv[k], oinv[k]), t1 = std::fmaf(fb[k][side * 2 + 1], inv[k], oinv[k]);
                                tn[side] = std::fmax(tn[side], std::fmin(t0, t1));
                                tf[side] = std::fmin(tf[side], std::fmax(t0, t1));
                            } else {
                                const float plo = 32768.0f + (float)(qb[k] & 0xffffu), phi = 32768.0f + (float)(qb[k] >> 16);
                                tn[side] = std::fmax(tn[side], std::fmaf(nearLow[k] ? plo : phi, inv[k], oinv[k]));
                                tf[side] = std::fmin(tf[side], std::fmaf(nearLow[k] ? phi : plo, inv[k], oinv[k]));
                            }
                        }
                    }
                    counts[2 + mode]++;
                    const bool hl = tn[0] <= tf[0] * kSlack, hr = tn[1] <= tf[1] * kSlack;
                    const bool both = hl & hr, any = hl | hr;
                    const bool rightNear = hr & (!hl | (tn[1] < tn[0]));
                    const int32_t nearRef = rightNear ? n.right : n.left, farRef = rightNear ? n.left : n.right;
                    const bool holdNear = any & (nearRef < 0) & !held();
                    const bool needPush = both & !holdNear, needPop = !any | (holdNear & !both);
                    int32_t next = holdNear ? farRef : nearRef;
                    if (needPush) {
                        if (sp >= 64) return "the walk overflows the stack";
                        stack[sp++] = farRef;
                    }
                    if (needPop) next = stack[--sp];
                    if (holdNear) leaf = ~nearRef;
                    if (next < 0 && next != kDone && !held()) {
                        leaf = ~next;
                        next = stack[--sp];
                    }
                    cur = next;
                } else {  // trav_prim_step2
                    const int32_t k = leaf >> 4;
                    const bool two = (leaf & 15) >= 2;
                    float t, u, v;
                    if (host_triangle(tri(k), o, d, tmin, t, u, v)) host_accept(best, t, u, v, k);
                    if (two && host_triangle(tri(k + 1), o, d, tmin, t, u, v)) host_accept(best, t, u, v, k + 1);
                    leaf += two ? 30 : 15;
                    if (!held() && cur < 0 && cur != kDone) {
                        leaf = ~cur;
                        cur = stack[--sp];
                    }
                }
                if (sp < 0) return "the walk underflows the stack";
            }
            found[mode] = best;
        }
        // 4. the walk on four-wide nodes (trav_node_step4 + trav_prim_step2), when the collapsed tree fits the device stack
        if (!bvh.nodes4.empty() && bvh.stack4 <= 64) {
            float inv[3], oinv[3];
            for (int k = 0; k < 3; k++) {
                inv[k] = slab_inv1(d[k]);
                oinv[k] = -o[k] * inv[k];
            }
            HostHit best;
            int32_t stack[64];
            int32_t sp = 0, cur = 0, leaf = 0;
            stack[sp++] = kDone;
            auto held = [&]() { return (leaf & 15) != 0; };
            uint64_t steps = 0;
            while (!(cur == kDone && !held())) {
                if (++steps > 100000) return "the four-wide walk does not terminate";
                if (cur < 0 && !held()) return "the four-wide walk stalls: a leaf waits in cur while no leaf is held";
                if (cur >= 0) {
                    const FlatNode4 &n = bvh.nodes4[(size_t)cur];
                    uint32_t key[4];
                    for (int c = 0; c < 4; c++) {
                        const float ax = std::fmaf(n.lox[c], inv[0], oinv[0]), bx = std::fmaf(n.hix[c], inv[0], oinv[0]);
                        const float ay = std::fmaf(n.loy[c], inv[1], oinv[1]), by = std::fmaf(n.hiy[c], inv[1], oinv[1]);
                        const float az = std::fmaf(n.loz[c], inv[2], oinv[2]), bz = std::fmaf(n.hiz[c], inv[2], oinv[2]);
                        const float tn = std::fmax(std::fmax(std::fmin(ax, bx), std::fmin(ay, by)), std::fmax(std::fmin(az, bz), tmin));
                        const float tf = std::fmin(std::fmin(std::fmax(ax, bx), std::fmax(ay, by)), std::fmin(std::fmax(az, bz), best.t));
                        const bool hit = tn <= tf * kSlack && n.ref[c] != leaf_ref(0, 0);  // unused slots are excluded by reference
                        uint32_t bits;
                        std::memcpy(&bits, &tn, 4);
                        key[c] = hit ? ((bits & ~3u) | (uint32_t)c) : 0xffffffffu;
                    }
                    std::sort(key, key + 4);
                    for (int c = 3; c >= 1; c--)
                        if (key[c] != 0xffffffffu) {
                            if (sp >= 64) return "the four-wide walk overflows the stack";
                            stack[sp++] = n.ref[key[c] & 3u];
                        }
                    int32_t next = key[0] != 0xffffffffu ? n.ref[key[0] & 3u] : stack[--sp];
                    if (next < 0 && next != kDone && !held()) {
                        leaf = ~next;
                        next = stack[--sp];
                    }
                    cur = next;
                } else {
                    const int32_t k = leaf >> 4;
                    const bool two = (leaf & 15) >= 2;
                    float t, u, v;
                    if (host_triangle(tri(k), o, d, tmin, t, u, v)) host_accept(best, t, u, v, k);
                    if (two && host_triangle(tri(k + 1), o, d, tmin, t, u, v)) host_accept(best, t, u, v, k + 1);
                    leaf += two ? 30 : 15;
                    if (!held() && cur < 0 && cur != kDone) {
                        leaf = ~cur;
                        cur = stack[--sp];
                    }
                }
                if (sp < 0) return "the four-wide walk underflows the stack";
            }
            if (best.prim != brute.prim || best.t != brute.t || best.u != brute.u || best.v != brute.v)
                return "the walk over four-wide nodes and the test of every triangle disagree";
        }
        counts[0]++;
        counts[1] += brute.prim >= 0;
        for (int mode = 0; mode < 2; mode++)
            if (found[mode].prim != brute.prim || found[mode].t != brute.t || found[mode].u != brute.u || found[mode].v != brute.v)
                return mode == 0 ? "the walk over float planes and the test of every triangle disagree" : "the walk over quantised planes and the test of every triangle disagree";
    }
    return "";
}

const char *validate_bvh(const BvhBuildResult &bvh, const std::vector<PrimBounds> &bounds) {
    const size_t n = bounds.size();
    if (bvh.nodes.empty()) return "no nodes";
    if (n == 0) return "";
    if (bvh.prim_order.size() != n) return "prim_order size mismatch";
    std::vector<uint8_t> seenPrim(n, 0), seenPos(n, 0), seenNode(bvh.nodes.size(), 0);
    for (int32_t id : bvh.prim_order) {
        if (id < 0 || (size_t)id >= n) return "prim_order entry out of range";
        if (seenPrim[(size_t)id]) return "primitive listed twice";
        seenPrim[(size_t)id] = 1;
    }
    struct W { int32_t node; uint32_t depth; float lo[3], hi[3]; };
    std::vector<W> st;
    const float inf = std::numeric_limits<float>::infinity();
    st.push_back({0, 1, {-inf, -inf, -inf}, {inf, inf, inf}});
    uint32_t depth = 0;
    while (!st.empty()) {
        W w = st.back();
        st.pop_back();
        if (w.node < 0 || (size_t)w.node >= bvh.nodes.size()) return "node ref out of range";
        if (seenNode[(size_t)w.node]) return "node reached twice";
        seenNode[(size_t)w.node] = 1;
        depth = std::max(depth, w.depth);
        const FlatNode &nd = bvh.nodes[(size_t)w.node];
        for (int s = 0; s < 2; s++) {
            float lo[3] = {nd.bx[s * 2], nd.by[s * 2], nd.bz[s * 2]}, hi[3] = {nd.bx[s * 2 + 1], nd.by[s * 2 + 1], nd.bz[s * 2 + 1]};
            int32_t ref = s ? nd.right : nd.left;
            if (!(lo[0] < inf_f())) continue;  // absent child
            for (int k = 0; k < 3; k++)
                if (lo[k] < w.lo[k] || hi[k] > w.hi[k]) return "child box not inside parent box";
            if (ref >= 0) {
                W c{ref, w.depth + 1, {lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}};
                st.push_back(c);
            } else {
                int32_t first = leaf_first(ref), count = leaf_count(ref);
                if (first < 0 || (size_t)first + (size_t)count > n) return "leaf range out of bounds";
                for (int32_t i = 0; i < count; i++) {
                    if (seenPos[(size_t)first + (size_t)i]) return "leaf ranges overlap";
                    seenPos[(size_t)first + (size_t)i] = 1;
                    const PrimBounds &p = bounds[(size_t)bvh.prim_order[(size_t)first + (size_t)i]];
                    for (int k = 0; k < 3; k++)
                        if (p.lo[k] < lo[k] || p.hi[k] > hi[k]) return "primitive not inside its leaf box";
                }
            }
        }
    }
    for (size_t i = 0; i < n; i++)
        if (!seenPos[i]) return "primitive not reachable from the root";
    if (depth > (uint32_t)kMaxTraversalDepth) return "tree deeper than the device stack";
    return "";
}

const char *validate_bvh4(const BvhBuildResult &bvh, const std::vector<PrimBounds> &bounds) {
    const size_t n = bounds.size();
    if (bvh.nodes4.empty()) return "no wide nodes";
    if (n == 0) return "";
    std::vector<uint8_t> seenPos(n, 0), seenNode(bvh.nodes4.size(), 0);
    struct W { int32_t node; uint32_t depth; float lo[3], hi[3]; };
    std::vector<W> st;
    const float inf = std::numeric_limits<float>::infinity();
    st.push_back({0, 1, {-inf, -inf, -inf}, {inf, inf, inf}});
    uint32_t depth = 0;
    while (!st.empty()) {
        W w = st.back();
        st.pop_back();
        if (w.node < 0 || (size_t)w.node >= bvh.nodes4.size()) return "wide node ref out of range";
        if (seenNode[(size_t)w.node]) return "wide node reached twice";
        seenNode[(size_t)w.node] = 1;
        depth = std::max(depth, w.depth);
        const FlatNode4 &nd = bvh.nodes4[(size_t)w.node];
        for (int c = 0; c < 4; c++) {
            float lo[3] = {nd.lox[c], nd.loy[c], nd.loz[c]}, hi[3] = {nd.hix[c], nd.hiy[c], nd.hiz[c]};
            if (!(lo[0] < inf_f())) continue;  // unused slot
            for (int k = 0; k < 3; k++)
                if (lo[k] < w.lo[k] || hi[k] > w.hi[k]) return "wide child box not inside parent box";
            int32_t ref = nd.ref[c];
            if (ref >= 0) {
                W cw{ref, w.depth + 1, {lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}};
                st.push_back(cw);
            } else {
                int32_t first = leaf_first(ref), count = leaf_count(ref);
                if (first < 0 || (size_t)first + (size_t)count > n) return "wide leaf range out of bounds";
                for (int32_t i = 0; i < count; i++) {
                    if (seenPos[(size_t)first + (size_t)i]) return "wide leaf ranges overlap";
                    seenPos[(size_t)first + (size_t)i] = 1;
                    const PrimBounds &pb = bounds[(size_t)bvh.prim_order[(size_t)first + (size_t)i]];
                    for (int k = 0; k < 3; k++)
                        if (pb.lo[k] < lo[k] || pb.hi[k] > hi[k]) return "primitive not inside its wide leaf box";
                }
            }
        }
    }
    for (size_t i = 0; i < n; i++)
        if (!seenPos[i]) return "primitive not reachable from the wide root";
    if (depth != bvh.depth4) return "depth4 mismatch";
    if (bvh.stack4 > 62) return "wide tree needs more stack than the device provides";
    return "";
}

}  // namespace ptc

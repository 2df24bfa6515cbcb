// bvh_sah.cpp -- host binned-SAH BVH builder (quality builder for static scenes).
//
// Replaces optixAccelBuild + optixAccelCompact for the scene geometry
// (OR/AudioRenderer.cpp:95-218).  The reference rebuilds its GAS on every emitter or
// receiver move (reload(), OR/AudioRenderer.cpp:466-486); here the scene BVH is built
// once and only the 1020-triangle receiver sub-tree is rebuilt on a move.
// The tree cannot change results (closest hit = min (t, triangle id) over exact
// per-triangle tests), only the number of node visits per segment.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>

#include "../arv2_internal.h"

namespace arv2 {

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
    void grow(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const float p[3]) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    float half_area() const
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0.f && dy >= 0.f && dz >= 0.f)) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Prim { Box b; float c[3]; int32_t id; };

struct TmpNode { Box box[2]; int32_t child[2]; };

// Build parameters.  Defaults are the measured best for the tracer (profiles/r07_bvh_experiments.md);
// the environment variables are tuning aids.
//   leaf_max  largest leaf (<= 8: three count bits in the leaf code)            ARV2_LEAF_MAX
//   ct        cost of one node visit in triangle tests for the SAH leaf test;   ARV2_SAH_CT
//             0 = always split down to leaf_max
//   bins      SAH bins per axis for large nodes                                  ARV2_SAH_BINS
//   sweep     nodes of at most this many triangles use the exact sweep           ARV2_SAH_SWEEP
constexpr int kMaxBins = 64;
struct BuildKnobs { int leaf_max = kMaxLeafTris; float ct = 1.f; int bins = 16; int sweep = 0; };
const BuildKnobs& knobs()
{
    static const BuildKnobs k = [] {
        BuildKnobs b;
        if (const char* e = getenv("ARV2_LEAF_MAX")) { const int x = atoi(e); if (x >= 1 && x <= 8) b.leaf_max = x; }
        if (const char* e = getenv("ARV2_SAH_CT")) { const float x = (float)atof(e); if (x >= 0.f) b.ct = x; }
        if (const char* e = getenv("ARV2_SAH_BINS")) { const int x = atoi(e); if (x >= 2 && x <= kMaxBins) b.bins = x; }
        if (const char* e = getenv("ARV2_SAH_SWEEP")) { const int x = atoi(e); if (x >= 0) b.sweep = x; }
        return b;
    }();
    return k;
}

struct Builder {
    std::vector<Prim> prims;
    std::vector<TmpNode> nodes;
    std::atomic<int32_t> next{0};
    float pad = 0.f;

    static int32_t leaf_code(int64_t first, int count) { return ~(int32_t)((first << kLeafShift) | (count - 1)); }

    // Builds the subtree over prims[first, first+count); returns its child code.
    // `bounds` = exact bounds of these prims (the SAH leaf test needs the parent's area).
    int32_t build(int64_t first, int64_t count, const Box& bounds, int depth, int par_depth)
    {
        const BuildKnobs& K = knobs();
        if (count <= 1) return leaf_code(first, (int)count);
        if (count <= K.leaf_max && K.ct <= 0.f) return leaf_code(first, (int)count);

        Box cb; cb.reset();
        for (int64_t i = first; i < first + count; ++i) cb.grow(prims[i].c);

        float best_cost = INFINITY; int best_axis = -1, best_split = -1;
        int64_t mid = -1;
        if (count <= K.sweep) {
            // small node: exact sweep over the centroid order of every axis
            std::vector<Prim> tmp(prims.begin() + first, prims.begin() + first + count), best_order;
            std::vector<float> right_area((size_t)count);
            for (int a = 0; a < 3; ++a) {
                if (!(cb.hi[a] - cb.lo[a] > 0.f)) continue;
                std::stable_sort(tmp.begin(), tmp.end(), [a](const Prim& x, const Prim& y) { return x.c[a] < y.c[a]; });
                Box acc; acc.reset();
                for (int64_t i = count - 1; i > 0; --i) { acc.grow(tmp[i].b); right_area[i] = acc.half_area(); }
                acc.reset();
                bool better = false;
                for (int64_t i = 0; i < count - 1; ++i) {
                    acc.grow(tmp[i].b);
                    const float cost = acc.half_area() * (float)(i + 1) + right_area[i + 1] * (float)(count - 1 - i);
                    if (cost < best_cost) { best_cost = cost; best_axis = a; mid = first + i + 1; better = true; }
                }
                if (better) best_order = tmp;
            }
            if (best_axis >= 0) {
                if (count <= K.leaf_max && K.ct + best_cost / std::max(bounds.half_area(), 1e-30f) >= (float)count)
                    return leaf_code(first, (int)count);
                std::copy(best_order.begin(), best_order.end(), prims.begin() + first);
            }
        } else {
            const int NB = K.bins;
            for (int a = 0; a < 3; ++a) {
                const float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0.f)) continue;
                Box bb[kMaxBins]; int cnt[kMaxBins];
                for (int k = 0; k < NB; ++k) { bb[k].reset(); cnt[k] = 0; }
                const float scale = NB / ext;
                for (int64_t i = first; i < first + count; ++i) {
                    int k = (int)((prims[i].c[a] - cb.lo[a]) * scale);
                    k = std::min(NB - 1, std::max(0, k));
                    bb[k].grow(prims[i].b); cnt[k]++;
                }
                float right_area[kMaxBins]; int right_cnt[kMaxBins];
                Box acc; acc.reset(); int c = 0;
                for (int k = NB - 1; k > 0; --k) { acc.grow(bb[k]); c += cnt[k]; right_area[k] = acc.half_area(); right_cnt[k] = c; }
                acc.reset(); c = 0;
                for (int k = 0; k < NB - 1; ++k) {
                    acc.grow(bb[k]); c += cnt[k];
                    if (c == 0 || right_cnt[k + 1] == 0) continue;
                    const float cost = acc.half_area() * (float)c + right_area[k + 1] * (float)right_cnt[k + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = k; }
                }
            }
            if (best_axis >= 0) {
                if (count <= K.leaf_max && K.ct + best_cost / std::max(bounds.half_area(), 1e-30f) >= (float)count)
                    return leaf_code(first, (int)count);
                const int a = best_axis;
                const float scale = NB / (cb.hi[a] - cb.lo[a]);
                const float lo = cb.lo[a];
                auto it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](const Prim& p) {
                    int k = (int)((p.c[a] - lo) * scale);
                    k = std::min(NB - 1, std::max(0, k));
                    return k <= best_split;
                });
                mid = it - prims.begin();
            }
        }
        if (best_axis < 0) {
            if (count <= K.leaf_max) return leaf_code(first, (int)count);   // coincident centroids
            mid = first + count / 2;                                       // split by position
        }
        if (mid <= first || mid >= first + count) mid = first + count / 2;

        Box lb, rb; lb.reset(); rb.reset();
        for (int64_t i = first; i < mid; ++i) lb.grow(prims[i].b);
        for (int64_t i = mid; i < first + count; ++i) rb.grow(prims[i].b);

        const int32_t me = next.fetch_add(1);
        TmpNode n;
        n.box[0] = lb; n.box[1] = rb;
        if (depth < par_depth && count > 8192) {
            auto fut = std::async(std::launch::async, [&, mid, first, depth] { return build(first, mid - first, lb, depth + 1, par_depth); });
            n.child[1] = build(mid, first + count - mid, rb, depth + 1, par_depth);
            n.child[0] = fut.get();
        } else {
            n.child[0] = build(first, mid - first, lb, depth + 1, par_depth);
            n.child[1] = build(mid, first + count - mid, rb, depth + 1, par_depth);
        }
        nodes[me] = n;
        return me;
    }
};

void store_box(BvhNode& n, int which, const Box& b, float pad)
{
    Box p = b;
    if (b.lo[0] <= b.hi[0]) for (int a = 0; a < 3; ++a) { p.lo[a] = b.lo[a] - pad; p.hi[a] = b.hi[a] + pad; }
    else for (int a = 0; a < 3; ++a) { p.lo[a] = kEmptyBox; p.hi[a] = kEmptyBox; }     // empty child: never entered
    n.q[which * 4 + 0] = p.lo[0]; n.q[which * 4 + 1] = p.hi[0];
    n.q[which * 4 + 2] = p.lo[1]; n.q[which * 4 + 3] = p.hi[1];
    n.q[8 + which * 2 + 0] = p.lo[2]; n.q[8 + which * 2 + 1] = p.hi[2];
}

} // namespace

void build_bvh_sah(const float* tv, int64_t n, HostBvh* out, int n_threads)
{
    Builder B;
    B.prims.resize(n);
    Box all; all.reset();
    for (int64_t i = 0; i < n; ++i) {
        Prim& p = B.prims[i];
        p.b.reset();
        for (int k = 0; k < 3; ++k) p.b.grow(tv + 9 * i + 3 * k);
        for (int a = 0; a < 3; ++a) p.c[a] = 0.5f * p.b.lo[a] + 0.5f * p.b.hi[a];
        p.id = (int32_t)i;
        all.grow(p.b);
    }
    float ext = 0.f;
    if (n > 0) for (int a = 0; a < 3; ++a) ext = std::max(ext, std::max(std::fabs(all.lo[a]), std::fabs(all.hi[a])));
    const float pad = bvh_pad(ext);
    B.pad = pad;
    B.nodes.resize((size_t)std::max<int64_t>(n, 1));

    int par_depth = 0;
    for (int t = std::max(1, n_threads); t > 1; t >>= 1) par_depth++;

    std::vector<TmpNode> final_nodes;
    int32_t root_code;
    if (n == 0) {
        TmpNode r; r.box[0].reset(); r.box[1].reset(); r.child[0] = r.child[1] = Builder::leaf_code(0, 1);
        B.nodes[0] = r; B.next = 1; root_code = 0;
    } else {
        root_code = n <= kMaxLeafTris ? Builder::leaf_code(0, (int)n) : B.build(0, n, all, 0, par_depth);
        if (root_code < 0) {        // everything fits one leaf: the root must still be an inner node
            TmpNode r; r.box[0] = all; r.box[1].reset();
            r.child[0] = root_code; r.child[1] = Builder::leaf_code(0, 1);
            const int32_t me = B.next.fetch_add(1);
            B.nodes[me] = r; root_code = me;
        }
    }

    // Re-lay out in depth-first pre-order with the root at index 0.
    const int32_t n_nodes = B.next.load();
    out->nodes.assign((size_t)n_nodes, BvhNode{});
    std::vector<int32_t> remap((size_t)n_nodes, -1);
    {
        std::vector<int32_t> stack{root_code};
        int32_t counter = 0;
        while (!stack.empty()) {
            const int32_t c = stack.back(); stack.pop_back();
            remap[c] = counter++;
            const TmpNode& t = B.nodes[c];
            if (t.child[1] >= 0) stack.push_back(t.child[1]);
            if (t.child[0] >= 0) stack.push_back(t.child[0]);
        }
    }
    for (int32_t i = 0; i < n_nodes; ++i) {
        if (remap[i] < 0) continue;
        const TmpNode& t = B.nodes[i];
        BvhNode& d = out->nodes[remap[i]];
        store_box(d, 0, t.box[0], pad);
        store_box(d, 1, t.box[1], pad);
        int32_t c[4] = {t.child[0] >= 0 ? remap[t.child[0]] : t.child[0], t.child[1] >= 0 ? remap[t.child[1]] : t.child[1], 0, 0};
        std::memcpy(&d.q[12], c, sizeof c);
    }
    out->order.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) out->order[i] = B.prims[i].id;
    for (int a = 0; a < 3; ++a) {
        out->lo[a] = n > 0 ? all.lo[a] - pad : kEmptyBox;
        out->hi[a] = n > 0 ? all.hi[a] + pad : kEmptyBox;
    }
}

} // namespace arv2

namespace arv2 {

// Nodes are stored in depth-first pre-order, so every inner child has a larger index
// than its parent: one backward sweep recomputes all boxes.
void refit_bvh(const float* tv, int64_t n, HostBvh* bvh)
{
    const int32_t nn = (int32_t)bvh->nodes.size();
    if (n == 0 || nn == 0) return;
    Box all; all.reset();
    for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) all.grow(tv + 9 * i + 3 * k);
    float ext = 0.f;
    for (int a = 0; a < 3; ++a) ext = std::max(ext, std::max(std::fabs(all.lo[a]), std::fabs(all.hi[a])));
    const float pad = bvh_pad(ext);
    std::vector<Box> nb((size_t)nn);
    for (int32_t i = nn - 1; i >= 0; --i) {
        BvhNode& d = bvh->nodes[i];
        int32_t c[4];
        std::memcpy(c, &d.q[12], sizeof c);
        Box both; both.reset();
        for (int w = 0; w < 2; ++w) {
            Box b; b.reset();
            bool empty = false;
            if (c[w] >= 0) b = nb[c[w]];
            else {
                const int32_t code = ~c[w];
                const int64_t first = code >> kLeafShift;
                const int cnt = (code & 7) + 1;
                empty = d.q[w * 4] == kEmptyBox;
                if (!empty) for (int t = 0; t < cnt; ++t) for (int k = 0; k < 3; ++k) b.grow(tv + 9 * (int64_t)bvh->order[first + t] + 3 * k);
            }
            if (!empty) both.grow(b);
            else b.reset();
            store_box(d, w, b, pad);
        }
        nb[i] = both;
    }
    for (int a = 0; a < 3; ++a) { bvh->lo[a] = all.lo[a] - pad; bvh->hi[a] = all.hi[a] + pad; }
}

} // namespace arv2

// bvh4.cpp -- collapse of the host binary BVH into 4-wide nodes (experiment: library built with -DARV2_WIDE=1 and
// ARV2_BVH4=1 in the environment; measured 21 % SLOWER than the binary tree, profiles/r07_trace_experiments.md
// section 18, so the default build does not use it).
//
// A wide node holds the boxes of up to four grandchildren of a binary node, so a segment makes about half as many
// dependent node visits.  Children keep fixed slots (half 0 = the binary child with the smaller box centre along the
// axis that separates the two children, half 1 = the other; inside a half the same rule), and the node records the
// three axes: the kernel orders the slots front to back from the signs of the ray direction alone, without a
// sorting network.  The tree only prunes; results cannot change.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "../arv2_internal.h"

namespace arv2 {

namespace {

struct Child { float lo[3], hi[3]; int32_t code; bool empty; };

Child child_of(const BvhNode& n, int w)
{
    Child c;
    c.lo[0] = n.q[w * 4 + 0]; c.hi[0] = n.q[w * 4 + 1]; c.lo[1] = n.q[w * 4 + 2]; c.hi[1] = n.q[w * 4 + 3];
    c.lo[2] = n.q[8 + w * 2]; c.hi[2] = n.q[8 + w * 2 + 1];
    int32_t ch[4];
    std::memcpy(ch, &n.q[12], sizeof ch);
    c.code = ch[w];
    c.empty = c.lo[0] == kEmptyBox;
    return c;
}

Child empty_child()
{
    Child c;
    for (int a = 0; a < 3; ++a) { c.lo[a] = kEmptyBox; c.hi[a] = kEmptyBox; }
    c.code = ~0; c.empty = true;
    return c;
}

// axis along which the centres of a and b are farthest apart; *swap: b lies on the smaller side
int split_axis(const Child& a, const Child& b, bool* swap)
{
    *swap = false;
    if (a.empty || b.empty) return 0;
    int best = 0; float bd = -1.f;
    for (int ax = 0; ax < 3; ++ax) {
        const float d = std::fabs((a.lo[ax] + a.hi[ax]) - (b.lo[ax] + b.hi[ax]));
        if (d > bd) { bd = d; best = ax; }
    }
    *swap = (b.lo[best] + b.hi[best]) < (a.lo[best] + a.hi[best]);
    return best;
}

struct Collapser {
    const HostBvh& b;
    std::vector<Bvh4Node>& out;
    int depth_max = 0;

    int32_t build(int32_t bin, int depth)
    {
        depth_max = std::max(depth_max, depth);
        const int32_t me = (int32_t)out.size();
        out.emplace_back();
        Child half[2] = {child_of(b.nodes[bin], 0), child_of(b.nodes[bin], 1)};
        bool sw;
        const int a0 = split_axis(half[0], half[1], &sw);
        if (sw) std::swap(half[0], half[1]);
        Child slot[4];
        int axes[2] = {0, 0};
        for (int h = 0; h < 2; ++h) {
            if (!half[h].empty && half[h].code >= 0) {            // inner binary child: its two children take the half's slots
                Child g0 = child_of(b.nodes[half[h].code], 0), g1 = child_of(b.nodes[half[h].code], 1);
                bool s2;
                axes[h] = split_axis(g0, g1, &s2);
                if (s2) std::swap(g0, g1);
                slot[2 * h] = g0; slot[2 * h + 1] = g1;
            } else {
                slot[2 * h] = half[h]; slot[2 * h + 1] = empty_child();
            }
        }
        Bvh4Node n;
        int32_t codes[4];
        for (int i = 0; i < 4; ++i) {
            n.q[0 + i] = slot[i].lo[0]; n.q[4 + i] = slot[i].hi[0];
            n.q[8 + i] = slot[i].lo[1]; n.q[12 + i] = slot[i].hi[1];
            n.q[16 + i] = slot[i].lo[2]; n.q[20 + i] = slot[i].hi[2];
            codes[i] = slot[i].code;
            if (!slot[i].empty && slot[i].code >= 0) codes[i] = kWideBit | build(slot[i].code, depth + 1);
        }
        std::memcpy(&n.q[24], codes, sizeof codes);
        const int32_t packed = a0 | (axes[0] << 2) | (axes[1] << 4);
        std::memcpy(&n.q[28], &packed, 4);
        n.q[29] = n.q[30] = n.q[31] = 0.f;
        out[(size_t)me] = n;
        return me;
    }
};

} // namespace

// 15-bit planes, rounded outwards and widened by one more cell (the device decodes a plane with a few ulp of error
// relative to the scene extent, far below a cell)
void quantise_bvh2(const HostBvh& b, std::vector<Q16Node>* out, float qk[3], float qinvk[3], float qc[3])
{
    double glo[3], cell[3];
    for (int a = 0; a < 3; ++a) {
        const double ext = std::max((double)b.hi[a] - (double)b.lo[a], 1e-6);
        glo[a] = b.lo[a];
        cell[a] = ext / 32760.0;                       // a few cells of headroom at the top
        qk[a] = (float)(cell[a] / 256.0);
        qinvk[a] = (float)(256.0 / cell[a]);
        qc[a] = (float)(glo[a] - 32768.0 * cell[a]);
    }
    out->assign(b.nodes.size(), Q16Node{});
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        const BvhNode& n = b.nodes[i];
        Q16Node q{};
        int32_t ch[4];
        std::memcpy(ch, &n.q[12], sizeof ch);
        for (int w = 0; w < 2; ++w) {
            const float lo[3] = {n.q[w * 4 + 0], n.q[w * 4 + 2], n.q[8 + w * 2]}, hi[3] = {n.q[w * 4 + 1], n.q[w * 4 + 3], n.q[8 + w * 2 + 1]};
            for (int a = 0; a < 3; ++a) {
                long ql, qh;
                if (lo[0] == kEmptyBox) { ql = qh = 32767; }          // absent child: a point in the far corner of the grid
                else {
                    ql = (long)std::floor(((double)lo[a] - glo[a]) / cell[a]) - 1;
                    qh = (long)std::ceil(((double)hi[a] - glo[a]) / cell[a]) + 1;
                    ql = std::min<long>(32767, std::max<long>(0, ql));
                    qh = std::min<long>(32767, std::max<long>(0, qh));
                }
                q.w[3 * w + a] = (uint32_t)ql | ((uint32_t)qh << 16);
            }
            q.child[w] = ch[w] >= 0 ? (kWideBit | ch[w]) : ch[w];
        }
        (*out)[i] = q;
    }
}

int collapse_bvh4(const HostBvh& bvh2, std::vector<Bvh4Node>* out)
{
    out->clear();
    if (bvh2.nodes.empty()) return 0;
    out->reserve(bvh2.nodes.size() / 2 + 2);
    Collapser c{bvh2, *out};
    c.build(0, 1);
    return c.depth_max;
}

} // namespace arv2

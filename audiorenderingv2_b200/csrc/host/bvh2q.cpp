// bvh2q.cpp -- 16-bit quantisation of the binary BVH into the 32 B node the kernels traverse,
// and the 64 B triangle record with its precomputed normal (layouts: arv2_internal.h).
//
// Why: the r02 profile showed the tracer bound by the L1 data pipe, which spends one
// wavefront per lane and load instruction on divergent gathers (profiles/micro/gather.cu).
// Boxes quantised to a per-tree 65536^3 grid make a two-child node exactly 32 B = ONE
// 256-bit load, and move all dequantisation constants from the node to the ray
// (plane = origin + q * cell; t = q * (cell/dir) + (origin - org)/dir).  A 4-wide node with
// 8-bit boxes was tried first and lost: 3x the ALU per visit for 1.67x fewer visits
// (profiles/r03a_ncu_summary.md).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "../arv2_internal.h"

namespace arv2 {

QuantGrid make_quant_grid(const float lo[3], const float hi[3])
{
    QuantGrid g;
    for (int a = 0; a < 3; ++a) {
        const double ext = std::max((double)hi[a] - (double)lo[a], 1e-6);
        const double cell = ext / 65000.0;                       // ~500 cells of slack for the outward margins
        g.cell[a] = (float)cell;
        g.origin[a] = (float)((double)lo[a] - 8.0 * cell);
    }
    return g;
}

namespace {

// Outward rounding with a 2-cell margin: the kernel evaluates the planes in ray space with
// the 2^23 magic-number conversion, whose cancellation costs up to ~0.3 cell.
void quantize_box(const QuantGrid& g, const float* lo, const float* hi, bool empty, uint32_t qlo[3], uint32_t qhi[3])
{
    for (int a = 0; a < 3; ++a) {
        if (empty) { qlo[a] = 65535; qhi[a] = 0; continue; }     // inverted: can never be entered
        const double l = std::floor(((double)lo[a] - (double)g.origin[a]) / (double)g.cell[a]) - 2.0;
        const double h = std::ceil(((double)hi[a] - (double)g.origin[a]) / (double)g.cell[a]) + 2.0;
        qlo[a] = (uint32_t)std::max(0.0, std::min(65535.0, l));
        qhi[a] = (uint32_t)std::max(0.0, std::min(65535.0, h));
    }
}

} // namespace

void quantize_bvh2(const HostBvh& b2, const QuantGrid& g, int32_t node_offset, int64_t slot_offset, std::vector<QNode>* out)
{
    out->resize(b2.nodes.size());
    for (size_t i = 0; i < b2.nodes.size(); ++i) {
        const BvhNode& n = b2.nodes[i];
        QNode q;
        int32_t ch[4];
        std::memcpy(ch, &n.q[12], sizeof ch);
        for (int w = 0; w < 2; ++w) {
            const float lo[3] = {n.q[w * 4 + 0], n.q[w * 4 + 2], n.q[8 + w * 2 + 0]};
            const float hi[3] = {n.q[w * 4 + 1], n.q[w * 4 + 3], n.q[8 + w * 2 + 1]};
            uint32_t ql[3], qh[3];
            quantize_box(g, lo, hi, lo[0] == kEmptyBox, ql, qh);
            for (int a = 0; a < 3; ++a) q.w[w * 3 + a] = ql[a] | (qh[a] << 16);
            int32_t c = ch[w];
            if (c >= 0) c += node_offset;
            else { const int32_t code = ~c; c = ~(int32_t)((((int64_t)(code >> kLeafShift) + slot_offset) << kLeafShift) | (code & 7)); }
            q.w[6 + w] = (uint32_t)c;
        }
        (*out)[i] = q;
    }
}

int bvh2_depth(const HostBvh& b)
{
    // nodes are in depth-first pre-order: children have larger indices than their parent
    std::vector<int> depth(b.nodes.size(), 1);
    int max_depth = b.nodes.empty() ? 0 : 1;
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        int32_t ch[4];
        std::memcpy(ch, &b.nodes[i].q[12], sizeof ch);
        for (int w = 0; w < 2; ++w)
            if (ch[w] >= 0 && (size_t)ch[w] < depth.size()) { depth[ch[w]] = depth[i] + 1; max_depth = std::max(max_depth, depth[ch[w]]); }
    }
    return max_depth;
}

// Ng = normalize(cross(P2-P1, P3-P1)) exactly as the arithmetic contract spells it
// (fma-form cross and dot, IEEE sqrt and division); built with -ffp-contract=off.
void make_tri_record(const float* v, int32_t id, int32_t material, float* out)
{
    const float e1[3] = {v[3] - v[0], v[4] - v[1], v[5] - v[2]};
    const float e2[3] = {v[6] - v[0], v[7] - v[1], v[8] - v[2]};
    const float nc[3] = {std::fmaf(e1[1], e2[2], -(e1[2] * e2[1])), std::fmaf(e1[2], e2[0], -(e1[0] * e2[2])),
                         std::fmaf(e1[0], e2[1], -(e1[1] * e2[0]))};
    const float dd = std::fmaf(nc[2], nc[2], std::fmaf(nc[1], nc[1], nc[0] * nc[0]));
    const float ninv = 1.0f / std::sqrt(dd);
    const float ng[3] = {nc[0] * ninv, nc[1] * ninv, nc[2] * ninv};
    out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; std::memcpy(&out[3], &id, 4);
    out[4] = v[3]; out[5] = v[4]; out[6] = v[5]; std::memcpy(&out[7], &material, 4);
    out[8] = v[6]; out[9] = v[7]; out[10] = v[8]; out[11] = ng[0];
    out[12] = ng[1]; out[13] = ng[2]; out[14] = 0.f; out[15] = 0.f;
}

} // namespace arv2

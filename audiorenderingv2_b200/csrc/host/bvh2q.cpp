// bvh2q.cpp -- collapse of the binary BVH into the 4-wide, 16-bit-quantised node the kernels
// traverse, and the 64 B triangle record with its precomputed normal (arv2_internal.h).
//
// History of the layout (profiles/): float binary nodes -> L1 wavefront-bound (r02);
// 4-wide nodes with per-node 8-bit grids -> 3x the ALU per visit, issue-bound (r03a);
// binary nodes on one global 16-bit grid -> one load per visit but still latency-bound
// (r03b).  A global grid moves every dequantisation constant to the ray, which makes a
// 4-wide node cheap enough (two PRMT+FFMA per plane) to halve the dependent steps.
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

struct Box { float lo[3], hi[3]; };
struct Child { Box box; int32_t ref; };   // ref: >= 0 binary inner node, < 0 binary leaf code

void child_box(const BvhNode& n, int w, Box* b)
{
    b->lo[0] = n.q[w * 4 + 0]; b->hi[0] = n.q[w * 4 + 1];
    b->lo[1] = n.q[w * 4 + 2]; b->hi[1] = n.q[w * 4 + 3];
    b->lo[2] = n.q[8 + w * 2 + 0]; b->hi[2] = n.q[8 + w * 2 + 1];
}
bool is_empty(const Box& b) { return b.lo[0] == kEmptyBox; }
float area(const Box& b)
{
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

} // namespace

// Outward rounding with a 2-cell margin: the kernel evaluates the planes in ray space with
// the 2^23 magic-number conversion, whose cancellation costs up to ~0.3 cell.
void make_qnode(const QuantGrid& g, const float (*lo)[3], const float (*hi)[3], const int32_t* codes, int n, QNode* out)
{
    for (int i = 0; i < 4; ++i) {
        for (int a = 0; a < 3; ++a) {
            uint32_t ql = 65535, qh = 0;                          // inverted: can never be entered
            if (i < n) {
                const double l = std::floor(((double)lo[i][a] - (double)g.origin[a]) / (double)g.cell[a]) - 2.0;
                const double h = std::ceil(((double)hi[i][a] - (double)g.origin[a]) / (double)g.cell[a]) + 2.0;
                ql = (uint32_t)std::max(0.0, std::min(65535.0, l));
                qh = (uint32_t)std::max(0.0, std::min(65535.0, h));
            }
            out->w[3 * i + a] = ql | (qh << 16);
        }
        out->w[12 + i] = (uint32_t)(i < n ? codes[i] : kEmptyEntry);
    }
}

void collapse_bvh4(const HostBvh& b2, const QuantGrid& g, int32_t node_offset, int64_t slot_offset, std::vector<QNode>* out)
{
    out->clear();
    struct Work { int32_t node4; int32_t node2; };
    std::vector<Work> todo;
    out->emplace_back();
    todo.push_back({0, 0});
    std::vector<Child> ch;
    while (!todo.empty()) {
        const Work w = todo.back(); todo.pop_back();
        // gather up to four children: open the largest binary inner child while there is room
        ch.clear();
        int32_t refs[4];
        std::memcpy(refs, &b2.nodes[w.node2].q[12], sizeof refs);
        for (int k = 0; k < 2; ++k) {
            Child c; child_box(b2.nodes[w.node2], k, &c.box); c.ref = refs[k];
            if (!is_empty(c.box)) ch.push_back(c);
        }
        while (ch.size() < 4) {
            int best = -1; float best_area = -1.f;
            for (size_t i = 0; i < ch.size(); ++i)
                if (ch[i].ref >= 0 && area(ch[i].box) > best_area) { best = (int)i; best_area = area(ch[i].box); }
            if (best < 0) break;
            const int32_t n2 = ch[best].ref;
            std::memcpy(refs, &b2.nodes[n2].q[12], sizeof refs);
            Child a, b; child_box(b2.nodes[n2], 0, &a.box); child_box(b2.nodes[n2], 1, &b.box);
            a.ref = refs[0]; b.ref = refs[1];
            ch.erase(ch.begin() + best);
            if (!is_empty(a.box)) ch.push_back(a);
            if (!is_empty(b.box)) ch.push_back(b);
        }
        float lo[4][3], hi[4][3];
        int32_t codes[4];
        const int n = (int)ch.size();
        for (int i = 0; i < n; ++i) {
            std::memcpy(lo[i], ch[i].box.lo, 12); std::memcpy(hi[i], ch[i].box.hi, 12);
            if (ch[i].ref >= 0) {
                const int32_t idx = (int32_t)out->size();
                out->emplace_back();
                todo.push_back({idx, ch[i].ref});
                codes[i] = idx + node_offset;
            } else {
                const int32_t code = ~ch[i].ref;
                codes[i] = ~(int32_t)((((int64_t)(code >> kLeafShift) + slot_offset) << kLeafShift) | (code & 7));
            }
        }
        make_qnode(g, lo, hi, codes, n, &(*out)[w.node4]);
    }
}

int bvh4_stack_need(const std::vector<QNode>& nodes, int32_t node_offset)
{
    // children are appended after their parent: one forward sweep gives the depths
    std::vector<int> depth(nodes.size(), 1);
    int max_depth = nodes.empty() ? 0 : 1;
    for (size_t i = 0; i < nodes.size(); ++i)
        for (int k = 0; k < 4; ++k) {
            const int32_t e = (int32_t)nodes[i].w[12 + k];
            if (e >= 0) {
                const size_t c = (size_t)(e - node_offset);
                if (c < depth.size()) { depth[c] = depth[i] + 1; max_depth = std::max(max_depth, depth[c]); }
            }
        }
    return 3 * max_depth + 4;
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

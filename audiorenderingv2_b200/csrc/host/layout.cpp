// layout.cpp -- helpers for the device layout: tree depth (stack bound) and the 64 B triangle
// record with its precomputed normal (arv2_internal.h).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "../arv2_internal.h"

namespace arv2 {

int bvh2_depth(const HostBvh& b)
{
    // depth by relaxation over the child links (any node order)
    std::vector<int> depth(b.nodes.size(), 0);
    if (b.nodes.empty()) return 0;
    std::vector<int32_t> todo{0};
    depth[0] = 1;
    int max_depth = 1;
    while (!todo.empty()) {
        const int32_t i = todo.back(); todo.pop_back();
        int32_t ch[4];
        std::memcpy(ch, &b.nodes[i].q[12], sizeof ch);
        for (int w = 0; w < 2; ++w)
            if (ch[w] >= 0 && (size_t)ch[w] < depth.size()) {
                depth[ch[w]] = depth[i] + 1;
                max_depth = std::max(max_depth, depth[ch[w]]);
                todo.push_back(ch[w]);
            }
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

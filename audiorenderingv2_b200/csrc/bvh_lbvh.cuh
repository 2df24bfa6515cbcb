// bvh_lbvh.cuh -- GPU LBVH builder interface (bvh_lbvh.cu).
#pragma once

#include <cuda_runtime.h>

namespace arv2 {

struct LbvhResult { int n_nodes; float lo[3], hi[3]; };

// d_verts: float[n][9], d_mats: int[n] (device).  Writes n_nodes 64 B nodes to d_nodes (node 0 =
// root; inner child indices are stored + node_offset) and n triangles in leaf order to
// d_tris (slots + slot_offset, global ids id_base + input index).  Needs n > 4.
// d_tris (48 B records in leaf order) and d_order (leaf-order slot -> input index) are optional.
cudaError_t build_bvh_lbvh(const float* d_verts, const int* d_mats, int n, int id_base, float4* d_nodes, float4* d_tris,
                           int* d_order, int node_offset, int slot_offset, LbvhResult* out, cudaStream_t stream);

// Stable radix sort of (key, value) pairs on key bits [lo_bit, hi_bit), 6 bits per pass; keys[0] / vals[0]
// hold the input, *result says which ping-pong buffer holds the output.  With scratch == nullptr the per-tile digit
// counts are allocated here and the stream is synchronised before they are freed; with the caller's scratch
// (radix_sort_scratch_bytes(n) bytes) the sort is only enqueued.
size_t radix_sort_scratch_bytes(int n);
cudaError_t radix_sort_pairs(unsigned* keys[2], int* vals[2], int n, int lo_bit, int hi_bit, int* result, cudaStream_t stream,
                             unsigned* scratch = nullptr);

} // namespace arv2

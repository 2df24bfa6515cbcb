// trace.cuh -- launch interface of the sm_100a trace kernels (trace.cu).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace arv2 {

// Everything one render needs, passed by value as the kernel parameter
// (the role of LaunchParams, OR/LaunchParams.h:20-43).
struct TraceParams {
    const float4* nodes;        // 64 B binary nodes: [top][scene tree][receiver tree] (arv2_internal.h)
    const float4* nodes4;       // optional alternative nodes of the scene tree (arv2_internal.h), their codes carry kWideBit:
                                // 128 B 4-wide nodes (-DARV2_WIDE=1) or 32 B quantised binary nodes (-DARV2_QNODES=1)
    float qk[3], qinvk[3], qc[3];   // quantised nodes: plane = qc + (2^23 + 256 q) * qk per axis, qinvk = 1 / qk
    const float4* tris;         // 64 B triangle records, leaf order
    const float* keep;          // [n_mats][bands]  1 - mat_absorption
    const float* scattering;    // [n_mats]
    double* hist;               // [2][bands][ir_len] fp64 accumulation
    unsigned long long* counters; // [0] next ray chunk, [1] segments traced, [2] sweeps (sweep_kernel launches), [7] watchdog
    int* rec_bin; int* rec_ear; float* rec_energy; int* rec_nseg;   // optional per-ray records
    // receiver-independent path cache (optional).  While it is being filled (launch_trace mode 1) it is ray-major with
    // a fixed stride: segment k of ray r at [r*pc_stride + k].  launch_cache_compact then packs it (CSR): the segments
    // of ray r are [pc_off[r], pc_off[r+1]), and adds the quantised path vertices the re-render scans.
    float4* pc_seg;             // 32 B per cached segment: (origin.xyz, t_wall or 1e20 on miss), (dir.xyz, distance before)
    float* pc_energy;           // [segment][bands] energy the segment starts with
    int* pc_nseg;               // [r] segments cached (filling only)
    long long pc_stride;        // records per ray while filling (>= max_bounces)
    const unsigned long long* pc_off;   // [n_rays + 1] CSR offsets
    // path vertices, 8 B each = (x, y, z) on a 16-bit grid over the scene bounds + flags; ray r owns vertices
    // [pc_off[r] + r, pc_off[r+1] + r + 1): the origins of its segments and the end point of the last one.
    // flags bit 0: a segment starts at this vertex; bit 1: that segment leaves the scene (no end point).
    const uint2* pc_vert;
    unsigned* pc_bits;          // one bit per vertex: its segment may enter the receiver's bounding ball
    long long pc_nvert;
    float pc_q0[3], pc_qs[3];   // vertex = pc_q0 + q * pc_qs per axis
    float pc_eps;               // what a true segment may stray from the line through its two quantised vertices
    unsigned long long seed;
    long long ray_begin, n_rays;
    float emitter[3], center[3];
    float recv_radius;          // ball about `center` that contains the placed receiver mesh (padded)
    float energy0, energy_thres, dist_thr, cross_gain, fs;
    unsigned max_bounces;
    int delay, ir_len, mono;
    int root;                   // node the full trace starts at (0 = two-level top node)
    int scene_root, recv_root;  // roots of the two sub-trees (-1: absent)
    int recv_nodes_shared;      // rr_walk_kernel: nodes of the receiver tree (from recv_root on) to keep in shared memory, 0 = none
    int any_scatter;            // 0: skip the diffuse-bounce RNG entirely
    // breadth-first tracer (wave_kernel; optional): per-SM, per-depth queues of path states
    float4* wave_paths;         // [sm][wave_queues][wave_cap][cont_f4(bands)]
    long long wave_cap;         // ring slots per queue = most paths alive per SM
    int wave_queues;            // queue j holds paths of depth (j + 1) * wave_segments
    int wave_segments;          // segments per task
    const int* ray_order;       // optional: the order in which the launch's rays [0, n_rays) are started (direction-sorted)
    int chunk;                  // rays a warp claims per global atomic
    int refill_below;           // lanes of a warp are refilled only while fewer than this many hold a path (32 = always)
};

// Bounce-synchronous tracer (sweep_kernel): the survivors of a sweep are handed over through global memory and re-binned
// by (origin cell, direction cell) for the next one.
struct SweepParams {
    const float4* in;                 // path states the sweep reads (null: the launch's fresh rays, in p.ray_order)
    const int* perm;                  // read order of `in`
    const unsigned long long* n_in;   // paths in `in`
    float4* out;                      // survivors, compacted
    unsigned long long* n_out;        // zeroed: survivors
    unsigned* key; unsigned* rank;    // per survivor: its bin and its arrival rank in the bin
    unsigned* bins;                   // zeroed: survivors per bin
    int segments;                     // segments per sweep
    int cell_bits, dir_bits, dir_major;
    float lo[3], scale[3];            // cell = (origin - lo) * scale per axis
};
struct SweepWork {                    // device workspace of one launch (owned by the context)
    float4* state[2]; unsigned* key; unsigned* rank; int* perm; unsigned* bins; unsigned* tile_sums; unsigned long long* count;
    int first_segments, segments, cell_bits, dir_bits, dir_major;
    float lo[3], scale[3];
};
constexpr int sweep_bins(int cell_bits, int dir_bits) { return (1 << (3 * cell_bits + 2 * dir_bits)) < 4096 ? 4096 : (1 << (3 * cell_bits + 2 * dir_bits)); }
cudaError_t launch_trace_sweeps(const TraceParams& p, const SweepWork& work, int bands, int mode, cudaStream_t stream);

constexpr int kWaveQueues = 64;               // most per-depth queues of wave_kernel
constexpr int kCounters = 24;                 // unsigned long long counters per context
constexpr int kStatInner = 16;                // -DARV2_TRACE_STATS: [16] node visits, [17] warp-level node steps, [18] leaf visits, [19] triangle tests, [20] warp-level leaf steps
__host__ __device__ constexpr int cont_f4(int bands) { return bands == 1 ? 3 : 5; }   // float4 per queued path

// 1 when the kernels were compiled with -DARV2_WIDE=1 (4-wide nodes in TraceParams::nodes4), 2 with -DARV2_QNODES=1
// (quantised binary nodes there), 0 otherwise
int trace_supports_wide_nodes();
// mode 0: full trace (scene + receiver), deposits into hist.
// mode 1: scene-only trace that fills the path cache (no deposits).
cudaError_t launch_trace(const TraceParams& p, int bands, int mode, int sm_count, cudaStream_t stream);
// Pack the freshly filled ray-major cache (src_*; stride p.pc_stride, counts p.pc_nseg) into the CSR arrays
// (p.pc_seg / p.pc_energy as destinations, p.pc_off already scanned) and write the quantised path vertices.
cudaError_t launch_cache_offsets(const int* nseg, long long n_rays, unsigned long long* off, unsigned long long* scratch, cudaStream_t stream);
cudaError_t launch_cache_compact(const TraceParams& p, const float4* src_seg, const float* src_energy, uint2* vert, int bands, cudaStream_t stream);
// Re-deposit from the packed path cache against the current receiver sub-tree:
//   rr_mask_kernel   streams the 8 B path vertices (coalesced, no dependence between rays) and writes one bit per
//                    segment: "may enter the receiver's bounding ball" (conservative);
//   rr_walk_kernel   per ray, in order: the flagged segments are walked through the receiver tree with their exact
//                    32 B record until the first hit (t_recv < t_wall), which deposits -- what a fresh trace does.
cudaError_t launch_rerender(const TraceParams& p, int bands, int sm_count, cudaStream_t stream);
// the r06 persistent per-ray scan of the 32 B records on the same CSR arrays (A/B: ARV2_RR_SERIAL=1)
cudaError_t launch_rerender_serial(const TraceParams& p, int bands, int sm_count, cudaStream_t stream);
// hist (fp64) -> ir_left / ir_right (fp32); mono: L = R = L + R (OR/kernels.cu:519-527).
cudaError_t launch_finalize(const double* hist, int bands, int ir_len, int mono, float* ir_left, float* ir_right,
                            cudaStream_t stream);

// keys[i] = Morton code of the octahedral map of the direction of ray (ray_begin + i), vals[i] = i
cudaError_t launch_direction_keys(unsigned long long seed, long long ray_begin, long long n, unsigned* keys, int* vals,
                                  cudaStream_t stream);

// out = vals (or 0 .. n-1) ordered by the top `bits` bits of keys (counting sort, arrival order inside a bin); keys are
// overwritten with the bins; rank: n entries, bins: 2^bits entries, tile_sums: 2^bits / 4096 entries.
constexpr int kCountingOrderMaxBits = 22;
constexpr size_t kCountingOrderScratch = ((size_t)1 << kCountingOrderMaxBits) + 1024;      // unsigned entries: bins, then tile sums
int counting_order_bits(long long n);
cudaError_t launch_counting_order(unsigned* keys, const int* vals, long long n, int bits, unsigned* rank, unsigned* bins, unsigned* tile_sums, int* out,
                                  cudaStream_t stream);
// The rays of the seeded set [0, n_total) whose direction tile (top tile_bits bits of the direction key) is = rank (mod
// n_ranks): their keys and global ids appended to keys / ids (at most `capacity`), *counter (zeroed) = how many there are.
cudaError_t launch_direction_select(unsigned long long seed, long long n_total, int rank, int n_ranks, int tile_bits, unsigned* keys, int* ids,
                                    unsigned long long* counter, long long capacity, int sm_count, cudaStream_t stream);

} // namespace arv2

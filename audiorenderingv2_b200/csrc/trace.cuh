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
    unsigned long long* counters; // [0] next ray chunk, [1] segments traced, [7] watchdog
    int* rec_bin; int* rec_ear; float* rec_energy; int* rec_nseg;   // optional per-ray records
    // receiver-independent path cache (optional), ray-major: segment k of ray r at [r*stride + k]
    float4* pc_seg;             // 32 B per cached segment: (origin.xyz, t_wall or 1e20 on miss), (dir.xyz, distance before)
    float* pc_energy;           // [r*stride + k][bands]
    int* pc_nseg;               // [r] segments cached
    long long pc_stride;        // records per ray (>= max_bounces)
    unsigned long long seed;
    long long ray_begin, n_rays;
    float emitter[3], center[3];
    float recv_radius;          // ball about `center` that contains the placed receiver mesh (padded)
    float energy0, energy_thres, dist_thr, cross_gain, fs;
    unsigned max_bounces;
    int delay, ir_len, mono;
    int root;                   // node the full trace starts at (0 = two-level top node)
    int scene_root, recv_root;  // roots of the two sub-trees (-1: absent)
    int any_scatter;            // 0: skip the diffuse-bounce RNG entirely
    // breadth-first tracer (wave_kernel; optional): per-SM, per-depth queues of path states
    float4* wave_paths;         // [sm][wave_queues][wave_cap][cont_f4(bands)]
    long long wave_cap;         // ring slots per queue = most paths alive per SM
    int wave_queues;            // queue j holds paths of depth (j + 1) * wave_segments
    int wave_segments;          // segments per task
    // data-parallel re-render (rr_scan / rr_walk / rr_resolve kernels; optional): candidate list and per-ray first hit
    int2* rr_cand;              // [rr_cap] (ray, k) of every cached segment that enters the receiver's bounding ball
    int2* rr_res;               // [rr_cap] (bin, ear) of the candidate's receiver hit, ear 0 = the walk missed
    float* rr_energy;           // [rr_cap][bands] chord-weighted energy of the hit
    int* rr_first;              // [n_rays] smallest k with a receiver hit (pre-set to 0x7f7f7f7f)
    long long rr_cap;           // counters[2] = candidates found (may exceed rr_cap: the caller then falls back)
    const int* ray_order;       // optional: the order in which the launch's rays [0, n_rays) are started (direction-sorted)
    int chunk;                  // rays a warp claims per global atomic
    int refill_below;           // lanes of a warp are refilled only while fewer than this many hold a path (32 = always)
};

constexpr int kWaveQueues = 64;               // most per-depth queues of wave_kernel
constexpr int kCounters = 16;                 // unsigned long long counters per context
__host__ __device__ constexpr int cont_f4(int bands) { return bands == 1 ? 3 : 5; }   // float4 per queued path

// 1 when the kernels were compiled with -DARV2_WIDE=1 (4-wide nodes in TraceParams::nodes4), 2 with -DARV2_QNODES=1
// (quantised binary nodes there), 0 otherwise
int trace_supports_wide_nodes();
// mode 0: full trace (scene + receiver), deposits into hist.
// mode 1: scene-only trace that fills the path cache (no deposits).
cudaError_t launch_trace(const TraceParams& p, int bands, int mode, int sm_count, cudaStream_t stream);
// Re-deposit from the path cache against the current receiver sub-tree: persistent per-ray scan (fallback) ...
cudaError_t launch_rerender(const TraceParams& p, int bands, int sm_count, cudaStream_t stream);
// ... and the data-parallel version (needs p.rr_*; rr_first pre-set to 0x7f bytes, counters zeroed).
cudaError_t launch_rerender_parallel(const TraceParams& p, int bands, int sm_count, cudaStream_t stream);
constexpr int kRrNoHit = 0x7f7f7f7f;
// hist (fp64) -> ir_left / ir_right (fp32); mono: L = R = L + R (OR/kernels.cu:519-527).
cudaError_t launch_finalize(const double* hist, int bands, int ir_len, int mono, float* ir_left, float* ir_right,
                            cudaStream_t stream);

// keys[i] = Morton code of the octahedral map of the direction of ray (ray_begin + i), vals[i] = i
cudaError_t launch_direction_keys(unsigned long long seed, long long ray_begin, long long n, unsigned* keys, int* vals,
                                  cudaStream_t stream);

} // namespace arv2

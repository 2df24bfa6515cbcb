// capi.cu -- C ABI of libarv2.so (include/arv2.h): scene / receiver / config objects and
// the renderer context that stands in for class AudioRenderer
// (OR/AudioRenderer.h:16-152, OR/AudioRenderer.cpp).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

#include <cuda_runtime.h>

#include "arv2_internal.h"
#include "bvh_lbvh.cuh"
#include "comm.cuh"
#include "conv.cuh"
#include "trace.cuh"

namespace arv2 {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

} // namespace arv2

using namespace arv2;

struct arv2_scene { HostScene s; };
struct arv2_receiver { HostReceiver r; };
struct arv2_comm { int device = 0, rank = 0, n_ranks = 1; ncclComm_t comm = nullptr; };
struct arv2_multi { std::vector<arv2_ctx*> ctx; std::vector<arv2_comm*> comm; };

#define CK(expr)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                             \
            return ARV2_ERR_CUDA;                                                                      \
        }                                                                                              \
    } while (0)

#define REQUIRE(cond, msg)                                                                             \
    do { if (!(cond)) { set_error(msg); return ARV2_ERR_INVALID; } } while (0)

struct arv2_ctx {
    int device = 0, sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    arv2_renderer_desc desc{};
    int ir_len = 0, bands = 1;
    long long n_rays_total = 0;

    // host copies
    HostScene scene;
    HostReceiver receiver;
    bool has_receiver = false;
    HostBvh scene_bvh, recv_bvh;     // binary trees (host copies)
    size_t upload_bytes = 0;
    bool recv_bvh_built = false;
    std::vector<float> recv_world;     // [n_recv][3][3]
    int64_t n_scene = 0, n_left = 0, n_right = 0;
    int32_t n_scene_nodes = 0, n_recv_nodes = 0;

    // parameters (LaunchParams)
    float emitter[3] = {0, 0, 0}, center[3] = {0, 0, 0}, yaw = 0.f;
    float base_power = 100.f, energy_thres = 0.f, hrtf = 0.9f;
    unsigned max_bounces = 10;
    int mono = 0;
    unsigned long long seed = 1;
    bool recv_dirty = true, cache_valid = false;
    float recv_radius = 0.f;
    int any_scatter = 0;

    // device
    float4* d_nodes = nullptr; float4* d_tris = nullptr;
    float4* d_nodes4 = nullptr;      // optional 4-wide scene tree (ARV2_BVH4=1, host/bvh4.cpp)
    int32_t scene_root_code = 1;     // node the scene tree is entered at: 1 (binary) or kWideBit | 0
    float qk[3] = {1.f, 1.f, 1.f}, qinvk[3] = {1.f, 1.f, 1.f}, qc[3] = {0.f, 0.f, 0.f};   // grid of the quantised nodes (any grid serves the float nodes)
    float* d_keep = nullptr; float* d_scatter = nullptr;
    double* d_hist = nullptr; float* d_ir_l = nullptr; float* d_ir_r = nullptr;     // d_ir_r = d_ir_l + bands * ir_len (one allocation)
    float* h_ir = nullptr;                        // pinned staging of both ears for arv2_get_ir
    unsigned long long* d_counters = nullptr;
    int* d_rec_bin = nullptr; int* d_rec_ear = nullptr; int* d_rec_nseg = nullptr; float* d_rec_energy = nullptr;
    long long rec_capacity = 0, last_range_rays = 0;
    // receiver-independent path cache, packed (trace.cuh): CSR records + energies, offsets, quantised path vertices, flag bits
    float4* d_pc_seg = nullptr; float* d_pc_energy = nullptr;
    unsigned long long* d_pc_off = nullptr; uint2* d_pc_vert = nullptr; unsigned* d_pc_bits = nullptr;
    long long pc_segs = 0, pc_nvert = 0;
    float pc_q0[3] = {0.f, 0.f, 0.f}, pc_qs[3] = {1.f, 1.f, 1.f}, pc_eps = 0.f;
    bool rr_serial = false;
    // breadth-first tracer: per-depth path queues, grown on demand
    float4* d_wave_paths = nullptr; size_t wave_slots = 0;
    bool wave = true;
    // bounce-synchronous tracer (sweep_kernel) for large launches: path states handed over through global memory
    SweepWork sweep{}; long long sweep_rays = 0; int sweep_nbins = 0;
    long long sweep_min_rays = 3000000;    // launches of at least this many rays use it (<= 0: never); crossover measured at 2 M rays (profiles/r09_trace_sweeps.md)
    // direction-sorted start orders, cached per (seed, ray range); two slots so that alternating ranges do not re-sort
    struct RayOrder { int* d = nullptr; long long begin = -1, n = -1; unsigned long long seed = 0, stamp = 0; } order[2];
    unsigned long long order_clock = 0;
    // workspace of the sort, kept (with the two order buffers, all of order_cap entries) for launches of up to kOrderKeep
    // rays: a render with a new seed then costs the key kernel and eight sort kernels, no allocation and no synchronisation
    unsigned* d_sort_keys[2] = {nullptr, nullptr}; unsigned* d_sort_counts = nullptr;
    long long order_cap = 0;
    // direction-tiled shard of the seeded set (arv2_render_tiles / arv2_render_sharded): the global ids of this rank's rays
    // in direction order, cached per (seed, ray count, rank, ranks)
    struct TileShard { int* d_ids = nullptr; long long n = -1, n_total = -1; unsigned long long seed = 0; int rank = -1, n_ranks = -1; } tiles;
    // workspace of the selection + sort (three id buffers rotate with tiles.d_ids), kept for shards of up to kOrderKeep rays
    struct TileWork { unsigned* keys[2] = {nullptr, nullptr}; int* vals[2] = {nullptr, nullptr}; unsigned* counts = nullptr;
                      unsigned long long* d_count = nullptr; long long cap = 0, ids_cap = 0; } tile_ws;
    int shard_mode = 1;               // arv2_render_sharded: 1 = direction tiles (default), 0 = contiguous slices of ray ids
    bool coherent_order = true;
    // pinned staging for the receiver sub-tree
    float4* h_stage = nullptr; size_t stage_f4 = 0;
    unsigned long long* h_counters = nullptr;
    long long last_segments = 0;
    // file-convolver workspace, grown on demand (the reference mallocs per call)
    struct ConvWork {
        float2* d_tw = nullptr; float* d_x = nullptr; float* d_out = nullptr; float2* d_X = nullptr; float2* d_H = nullptr;
        size_t cap_x = 0, cap_X = 0, cap_H = 0;
    } conv;
};

struct arv2_stream {
    int device = 0, n_src = 0, block = 0, ir_len = 0, P = 0, slot = 0;
    int cluster = 8;                               // CTAs per source of the step kernel: 16 for streams of <= 4 sources (r09)
    int stages = 0;                                // ring depth of the step kernel (deep when a step and its successor find SMs of their own)
    cudaStream_t stream = nullptr;
    float2* d_tw = nullptr; float2* d_fdl = nullptr; float2* d_H[2] = {nullptr, nullptr};
    float2** d_Hptr = nullptr; float2** h_Hptr = nullptr;     // h_Hptr: one pinned entry per (source, swap parity)
    std::vector<int> active;
    float* d_tail = nullptr; float* d_ir = nullptr;
    float* d_in = nullptr; float* d_out = nullptr; float* d_mix = nullptr;     // device side of the host-buffer calls: [kHostBlocks][n_src][block], [..][n_src][2][block], [..][2][block]
    float* d_gain = nullptr;                       // [n_src] mix gains (null = 1)
    // host-buffer entry points: pinned, device-mapped staging the kernels read / write directly (no copy engine, no
    // stream synchronisation per block), and a completion word the host spins on
    float* h_in = nullptr; float* h_out = nullptr; float* h_mix = nullptr; float* h_ir = nullptr;
    unsigned* h_flag = nullptr; unsigned flag_seq = 0; unsigned* d_done = nullptr;
    // ordering between the stream the steps run on (the caller's or our own) and the IR swaps on our own stream
    cudaEvent_t ev_step = nullptr, ev_swap = nullptr, ev_ir_copied = nullptr;
    bool step_pending = false, swap_pending = false, ir_copy_pending = false;
};
constexpr int kHostBlocks = 16;                    // most blocks one host-buffer call carries (an RtAudio callback: 4096 frames = 8)

namespace {

// Binary nodes with their child links moved into the combined index space.
void offset_nodes(const HostBvh& b, int32_t node_offset, int64_t slot_offset, BvhNode* out)
{
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        BvhNode d = b.nodes[i];
        int32_t ch[4];
        std::memcpy(ch, &d.q[12], sizeof ch);
        for (int w = 0; w < 2; ++w) {
            if (ch[w] >= 0) ch[w] += node_offset;
            else { const int32_t code = ~ch[w]; ch[w] = ~(int32_t)((((int64_t)(code >> kLeafShift) + slot_offset) << kLeafShift) | (code & 7)); }
        }
        std::memcpy(&d.q[12], ch, sizeof ch);
        out[i] = d;
    }
}

// two-level top node: child 0 = scene tree (node 1), child 1 = receiver tree
BvhNode make_top_node(const arv2_ctx* c, bool with_receiver)
{
    BvhNode n{};
    for (int i = 0; i < 12; ++i) n.q[i] = kEmptyBox;
    int32_t ch[4] = {~0, ~0, 0, 0};
    if (c->n_scene > 0) {
        n.q[0] = c->scene_bvh.lo[0]; n.q[1] = c->scene_bvh.hi[0]; n.q[2] = c->scene_bvh.lo[1]; n.q[3] = c->scene_bvh.hi[1];
        n.q[8] = c->scene_bvh.lo[2]; n.q[9] = c->scene_bvh.hi[2];
        ch[0] = c->scene_root_code;
    }
    if (with_receiver) {
        n.q[4] = c->recv_bvh.lo[0]; n.q[5] = c->recv_bvh.hi[0]; n.q[6] = c->recv_bvh.lo[1]; n.q[7] = c->recv_bvh.hi[1];
        n.q[10] = c->recv_bvh.lo[2]; n.q[11] = c->recv_bvh.hi[2];
        ch[1] = 1 + c->n_scene_nodes;
    }
    std::memcpy(&n.q[12], ch, sizeof ch);
    return n;
}

int upload_receiver(arv2_ctx* c)
{
    if (!c->has_receiver || !c->recv_dirty) return ARV2_OK;
    const int64_t nl = c->n_left, nr = c->n_right, n = nl + nr;
    c->recv_world.resize((size_t)n * 9);
    place_receiver_half(c->receiver.left, c->center, c->yaw, c->recv_world.data());
    place_receiver_half(c->receiver.right, c->center, c->yaw, c->recv_world.data() + nl * 9);
    {
        double r2 = 0.0;
        for (int64_t i = 0; i < n * 3; ++i) {
            const float* v = c->recv_world.data() + 3 * i;
            const double dx = (double)v[0] - c->center[0], dy = (double)v[1] - c->center[1], dz = (double)v[2] - c->center[2];
            r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
        }
        c->recv_radius = (float)(std::sqrt(r2) * 1.001 + 1e-3);
    }
    if (!c->recv_bvh_built) {
        build_bvh_sah(c->recv_world.data(), n, &c->recv_bvh, 1);
        c->recv_bvh_built = true;
        if ((int32_t)c->recv_bvh.nodes.size() > c->n_recv_nodes) { set_error("receiver BVH larger than reserved"); return ARV2_ERR_STATE; }
    } else {
        refit_bvh(c->recv_world.data(), n, &c->recv_bvh);
    }
    // stage [top node][receiver nodes][receiver triangles]
    const int32_t node_base = 1 + c->n_scene_nodes;
    const int64_t tri_base = c->n_scene;
    const int32_t nn = (int32_t)c->recv_bvh.nodes.size();
    if (nn > c->n_recv_nodes || bvh2_depth(c->recv_bvh) + 3 > kTraversalStack) { set_error("receiver BVH exceeds its reservation"); return ARV2_ERR_STATE; }
    float4* st = c->h_stage;
    const BvhNode top = make_top_node(c, true);
    std::memcpy(st, &top, sizeof top);
    float4* sn = st + 4;
    offset_nodes(c->recv_bvh, node_base, tri_base, (BvhNode*)sn);
    float4* stt = sn + 4 * (size_t)c->n_recv_nodes;
    for (int64_t s = 0; s < n; ++s) {
        const int32_t src = c->recv_bvh.order[s];
        make_tri_record(c->recv_world.data() + 9 * (size_t)src, (int32_t)(tri_base + src), src < nl ? -1 : -2, (float*)(stt + 4 * s));
    }
    CK(cudaMemcpyAsync(c->d_nodes, st, sizeof(BvhNode), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_nodes + 4 * (size_t)node_base, sn, sizeof(BvhNode) * (size_t)nn, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_tris + 4 * (size_t)tri_base, stt, sizeof(float4) * 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    c->upload_bytes = sizeof(BvhNode) * (size_t)(nn + 1) + sizeof(float4) * 4 * (size_t)n;
    c->recv_dirty = false;
    return ARV2_OK;
}

void fill_params(arv2_ctx* c, TraceParams* p, long long ray_begin, long long n_rays)
{
    std::memset(p, 0, sizeof *p);
    p->nodes = c->d_nodes; p->tris = c->d_tris; p->keep = c->d_keep; p->scattering = c->d_scatter;
    p->hist = c->d_hist; p->counters = c->d_counters;
    if (c->desc.record_rays && n_rays <= c->rec_capacity) {
        p->rec_bin = c->d_rec_bin; p->rec_ear = c->d_rec_ear; p->rec_energy = c->d_rec_energy; p->rec_nseg = c->d_rec_nseg;
    }
    p->pc_seg = c->d_pc_seg; p->pc_energy = c->d_pc_energy; p->pc_off = c->d_pc_off; p->pc_vert = c->d_pc_vert; p->pc_bits = c->d_pc_bits;
    p->pc_nvert = c->pc_nvert; p->pc_eps = c->pc_eps;
    for (int a = 0; a < 3; ++a) { p->pc_q0[a] = c->pc_q0[a]; p->pc_qs[a] = c->pc_qs[a]; }
    p->seed = c->seed; p->ray_begin = ray_begin; p->n_rays = n_rays;
    for (int a = 0; a < 3; ++a) { p->emitter[a] = c->emitter[a]; p->center[a] = c->center[a]; }
    p->recv_radius = c->recv_radius;
    // OR/devicePrograms.cu:208  base_power / ((x*y*z) * 4.18879020478), double then narrowed
    p->energy0 = (float)((double)c->base_power / ((double)(int)c->n_rays_total * 4.18879020478));
    p->energy_thres = c->energy_thres;
    // :227-228
    int ir_sec = c->ir_len / c->desc.sample_rate;
    ir_sec = ir_sec < 1 ? 1 : (ir_sec > 999 ? 999 : ir_sec);
    p->dist_thr = (float)(ir_sec * 343 + 1);
    p->cross_gain = 1.0f - c->hrtf;                      // :139 (1 - hrtf_absorption_rate)
    p->fs = (float)c->desc.sample_rate;
    p->max_bounces = c->max_bounces;
    p->delay = (int)((double)c->desc.sample_rate * 0.00044);   // :125
    p->ir_len = c->ir_len; p->mono = c->mono;
    p->root = 0;
    p->scene_root = c->n_scene > 0 ? c->scene_root_code : -1;
    p->nodes4 = c->d_nodes4;
    for (int a = 0; a < 3; ++a) { p->qk[a] = c->qk[a]; p->qinvk[a] = c->qinvk[a]; p->qc[a] = c->qc[a]; }
    p->recv_root = c->has_receiver ? 1 + c->n_scene_nodes : -1;
    // rr_walk_kernel keeps the receiver tree's nodes in shared memory when they fit the default 48 KB with room to spare
    { const size_t nn = c->recv_bvh.nodes.size(); p->recv_nodes_shared = (c->has_receiver && nn > 0 && nn * 64 <= 40 * 1024 && !getenv("ARV2_RR_NO_SHARED")) ? (int)nn : 0; }
    p->any_scatter = c->any_scatter;
    // a warp tops up its free lanes only once at least 9 are free, 16 rays per claim: the rays it
    // starts together are neighbours in the direction order (tuning aids: ARV2_CHUNK, ARV2_REFILL_BELOW)
    p->chunk = 16;
    p->refill_below = c->coherent_order ? 24 : 33;
    if (const char* e = getenv("ARV2_CHUNK")) p->chunk = atoi(e) > 0 ? atoi(e) : p->chunk;
    if (const char* e = getenv("ARV2_REFILL_BELOW")) p->refill_below = atoi(e) > 0 ? atoi(e) : p->refill_below;
    p->ray_order = nullptr;
    for (const auto& o : c->order)
        if (o.d && o.begin == ray_begin && o.n == n_rays && o.seed == c->seed) p->ray_order = o.d;
}

// Coherent ray order: the rays of a launch are started in the order of their emission direction
// (Morton code of the octahedral map, GPU radix sort on its top 24 bits), so the 32 lanes of a
// warp walk the same part of the scene over the first bounces.  Depends only on (seed, range);
// the result per ray and the fp64 histogram do not depend on the order.
constexpr long long kOrderKeep = 32LL << 20;

void free_ray_orders(arv2_ctx* c)
{
    for (auto& o : c->order) { cudaFree(o.d); o = arv2_ctx::RayOrder{}; }
    cudaFree(c->d_sort_keys[0]); cudaFree(c->d_sort_keys[1]); cudaFree(c->d_sort_counts);
    c->d_sort_keys[0] = c->d_sort_keys[1] = nullptr; c->d_sort_counts = nullptr; c->order_cap = 0;
}

int ensure_ray_order(arv2_ctx* c, long long ray_begin, long long n_rays)
{
    if (!c->coherent_order || n_rays <= 0 || n_rays > 0x7fffffffLL) return ARV2_OK;
    arv2_ctx::RayOrder* slot = &c->order[0];
    for (auto& o : c->order) {
        if (o.d && o.begin == ray_begin && o.n == n_rays && o.seed == c->seed) { o.stamp = ++c->order_clock; return ARV2_OK; }
        if (o.stamp < slot->stamp) slot = &o;                     // least recently used
    }
    static const bool radix = getenv("ARV2_ORDER_RADIX") != nullptr;      // A/B: the four-pass stable radix sort of r05-r08
    if (n_rays <= kOrderKeep && !radix) {
        // persistent buffers: grown (and both cached orders dropped) when a launch is larger than any before
        if (n_rays > c->order_cap) {
            CK(cudaStreamSynchronize(c->stream));
            free_ray_orders(c);
            slot = &c->order[0];
            const size_t bytes = (size_t)n_rays * 4;
            cudaError_t e = cudaMalloc(&c->d_sort_keys[0], bytes);
            if (e == cudaSuccess) e = cudaMalloc(&c->d_sort_keys[1], bytes);
            if (e == cudaSuccess) e = cudaMalloc(&c->order[0].d, bytes);
            if (e == cudaSuccess) e = cudaMalloc(&c->order[1].d, bytes);
            if (e == cudaSuccess) e = cudaMalloc(&c->d_sort_counts, kCountingOrderScratch * sizeof(unsigned));      // bins + tile sums
            if (e != cudaSuccess) { free_ray_orders(c); cudaGetLastError(); set_error(std::string("ray order: ") + cudaGetErrorString(e)); return ARV2_ERR_CUDA; }
            c->order_cap = n_rays;
        }
        slot->n = -1;
        // direction keys, then a counting sort on their top bits (keys -> bins in place, ranks in the second key buffer)
        const int bits = counting_order_bits(n_rays);
        cudaError_t e = launch_direction_keys(c->seed, ray_begin, n_rays, c->d_sort_keys[0], nullptr, c->stream);
        if (e == cudaSuccess) e = launch_counting_order(c->d_sort_keys[0], nullptr, n_rays, bits, c->d_sort_keys[1], c->d_sort_counts, c->d_sort_counts + ((size_t)1 << kCountingOrderMaxBits),
                                                        slot->d, c->stream);
        if (e != cudaSuccess) { set_error(std::string("ray order: ") + cudaGetErrorString(e)); return ARV2_ERR_CUDA; }
        slot->begin = ray_begin; slot->n = n_rays; slot->seed = c->seed; slot->stamp = ++c->order_clock;
        return ARV2_OK;
    }
    // very large launches (and the radix A/B): transient workspace (2 x 8 B per ray), only the order itself is kept
    if (c->order_cap > 0) { CK(cudaStreamSynchronize(c->stream)); free_ray_orders(c); slot = &c->order[0]; }
    cudaFree(slot->d); slot->d = nullptr; slot->n = -1;
    unsigned* keys[2] = {nullptr, nullptr};
    int* vals[2] = {nullptr, nullptr};
    auto cleanup = [&]() { cudaFree(keys[0]); cudaFree(keys[1]); cudaFree(vals[0]); cudaFree(vals[1]); };
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaMalloc(&keys[k], (size_t)n_rays * sizeof(unsigned));
        if (e == cudaSuccess) e = cudaMalloc(&vals[k], (size_t)n_rays * sizeof(int));
    }
    int res = 0;
    if (e == cudaSuccess) e = launch_direction_keys(c->seed, ray_begin, n_rays, keys[0], vals[0], c->stream);
    if (e == cudaSuccess && radix) e = radix_sort_pairs(keys, vals, (int)n_rays, 8, 32, &res, c->stream);
    else if (e == cudaSuccess) {
        unsigned* bins = nullptr;
        e = cudaMalloc(&bins, kCountingOrderScratch * sizeof(unsigned));
        if (e == cudaSuccess) e = launch_counting_order(keys[0], nullptr, n_rays, counting_order_bits(n_rays), keys[1], bins, bins + ((size_t)1 << kCountingOrderMaxBits), vals[1], c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(bins);
        res = 1;
    }
    if (e != cudaSuccess) { cleanup(); set_error(std::string("ray order: ") + cudaGetErrorString(e)); return ARV2_ERR_CUDA; }
    slot->d = vals[res]; vals[res] = nullptr;
    cleanup();
    slot->begin = ray_begin; slot->n = n_rays; slot->seed = c->seed; slot->stamp = ++c->order_clock;
    return ARV2_OK;
}

// Direction-tiled shard (trace.cu: direction_select_kernel): this rank's rays of the whole seeded set, sorted by direction
// key.  One pass over all n_total ray ids (Philox + the emission direction: ~0.2 ms per 10 M rays) and a sort of this rank's
// share; cached per (seed, n_total, rank, n_ranks).  Synchronises the stream (the count is read back).
// 2^14 tiles for sets of >= 4 M rays (256+ rays per tile; measured on the 8 shards of 8 M rays in the C2 room: slowest / mean
// shard time 1.052 / 1.024 / 1.026 / 1.007 / 1.011 with 2^8 / 2^10 / 2^12 / 2^14 / 2^16 tiles), fewer for small sets
constexpr int kTileBits = 14;

void free_tiles(arv2_ctx* c)
{
    arv2_ctx::TileWork& w = c->tile_ws;
    cudaFree(w.keys[0]); cudaFree(w.keys[1]); cudaFree(w.vals[0]); cudaFree(w.vals[1]); cudaFree(w.counts); cudaFree(w.d_count);
    w = arv2_ctx::TileWork{};
    cudaFree(c->tiles.d_ids);
    c->tiles = arv2_ctx::TileShard{};
}

int ensure_tiles(arv2_ctx* c, int rank, int n_ranks)
{
    arv2_ctx::TileShard& t = c->tiles;
    arv2_ctx::TileWork& w = c->tile_ws;
    const long long n_total = c->n_rays_total;
    if (t.d_ids && t.seed == c->seed && t.n_total == n_total && t.rank == rank && t.n_ranks == n_ranks) return ARV2_OK;
    CK(cudaStreamSynchronize(c->stream));                   // (the previous shard's ids may still be read by a launch in flight)
    t.n = -1;
    // the octahedral map is not equal-area: a rank's share of the interleaved tiles stays within a few per cent of 1 / R
    const long long cap = n_total / n_ranks + n_total / (4 * n_ranks) + 65536;
    if (cap > w.cap || w.ids_cap < cap) {
        free_tiles(c);
        const size_t bytes = (size_t)cap * 4;
        cudaError_t e = cudaMalloc(&w.d_count, sizeof(unsigned long long));
        for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
            e = cudaMalloc(&w.keys[k], bytes);
            if (e == cudaSuccess) e = cudaMalloc(&w.vals[k], bytes);
        }
        if (e == cudaSuccess) e = cudaMalloc(&t.d_ids, bytes);
        if (e == cudaSuccess) e = cudaMalloc(&w.counts, kCountingOrderScratch * sizeof(unsigned));      // bins + tile sums of the counting sort
        if (e != cudaSuccess) { free_tiles(c); cudaGetLastError(); set_error(std::string("direction tiles: ") + cudaGetErrorString(e)); return ARV2_ERR_CUDA; }
        w.cap = cap; w.ids_cap = cap;
    }
    unsigned long long count = 0;
    int tile_bits = kTileBits;
    while (tile_bits > 6 && (n_total >> tile_bits) < 256) --tile_bits;
    if (const char* env = getenv("ARV2_TILE_BITS")) tile_bits = std::min(20, std::max(4, atoi(env)));      // tuning aid
    cudaError_t e = cudaMemsetAsync(w.d_count, 0, sizeof(unsigned long long), c->stream);
    if (e == cudaSuccess) e = launch_direction_select(c->seed, n_total, rank, n_ranks, tile_bits, w.keys[0], w.vals[0], w.d_count, w.cap, c->sm_count, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&count, w.d_count, sizeof count, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess && (long long)count > w.cap) { set_error("direction tiles: a rank's share exceeds its reservation"); return ARV2_ERR_STATE; }
    // this rank's ids in direction order: counting sort on the top bits of their keys (ranks in the second key buffer)
    if (e == cudaSuccess) e = launch_counting_order(w.keys[0], w.vals[0], (long long)count, counting_order_bits((long long)count), w.keys[1], w.counts,
                                                    w.counts + ((size_t)1 << kCountingOrderMaxBits), t.d_ids, c->stream);
    if (e != cudaSuccess) { set_error(std::string("direction tiles: ") + cudaGetErrorString(e)); return ARV2_ERR_CUDA; }
    t.n = (long long)count; t.n_total = n_total; t.seed = c->seed; t.rank = rank; t.n_ranks = n_ranks;
    if (w.cap > kOrderKeep) {                                // very large shards: only the ids are kept
        int* keep = t.d_ids; t.d_ids = nullptr;
        const arv2_ctx::TileShard meta = t;
        CK(cudaStreamSynchronize(c->stream));
        free_tiles(c);
        c->tiles = meta; c->tiles.d_ids = keep;
    }
    return ARV2_OK;
}

// Per-ray records (desc.record_rays) are indexed by the ray's position in the launch: sized for the largest range
// rendered so far, not for the whole seeded set (a 200k-ray shard of a 100M-ray set needs 200k entries).
int ensure_records(arv2_ctx* c, long long n_rays)
{
    if (!c->desc.record_rays || n_rays <= c->rec_capacity) return ARV2_OK;
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_rec_bin); cudaFree(c->d_rec_ear); cudaFree(c->d_rec_nseg); cudaFree(c->d_rec_energy);
    c->d_rec_bin = nullptr; c->d_rec_ear = nullptr; c->d_rec_nseg = nullptr; c->d_rec_energy = nullptr; c->rec_capacity = 0;
    CK(cudaMalloc(&c->d_rec_bin, (size_t)n_rays * sizeof(int)));
    CK(cudaMalloc(&c->d_rec_ear, (size_t)n_rays * sizeof(int)));
    CK(cudaMalloc(&c->d_rec_nseg, (size_t)n_rays * sizeof(int)));
    CK(cudaMalloc(&c->d_rec_energy, (size_t)n_rays * c->bands * sizeof(float)));
    c->rec_capacity = n_rays;
    return ARV2_OK;
}

// Queues of the breadth-first tracer for a launch of n_rays rays: per SM, wave_queues rings of wave_cap
// path states.  Sets the wave_* fields of p (left null when the queues are off or cannot be allocated:
// the launch then uses the depth-first persistent kernel).
int ensure_wave(arv2_ctx* c, TraceParams* p, long long n_rays)
{
    if (!c->wave || n_rays <= 0 || n_rays > 0x7fffffffLL) return ARV2_OK;
    // 8 segments per task, 2048 paths alive per SM (64 batches for 32 warps): measured best on B200 for 1M..8M rays
    // (profiles/r05_trace_experiments.md section 2: segments 1..8 x cap 1024..8192; r07 section 5: 8 segments are
    // +1.2 % over 4 at 8M rays on the SAH-leaf tree and halve the queue memory)
    int per = 8;
    if (const char* e = getenv("ARV2_WAVE_SEGMENTS")) per = atoi(e) > 0 ? atoi(e) : per;      // tuning aid
    // segments per task so that ceil(max_bounces / per) per-depth queues fit (64-bit: max_bounces may be UINT_MAX)
    const long long mb = c->max_bounces < 1 ? 1LL : (long long)c->max_bounces;
    per = (int)std::max<long long>(per, (mb + kWaveQueues - 1) / kWaveQueues);
    const int nq = (int)((mb + per - 1) / per);
    if (nq < 1 || nq > kWaveQueues || mb > (1LL << 24)) return ARV2_OK;      // absurd depth: depth-first kernel
    long long cap = 2048;
    if (const char* e = getenv("ARV2_WAVE_CAP")) cap = atoll(e) >= 64 ? atoll(e) : cap;       // tuning aid
    const long long share = (2 * n_rays / c->sm_count + 95) / 32 * 32;
    if (share < cap) cap = share;
    const size_t need = (size_t)c->sm_count * (size_t)nq * (size_t)cap;
    if (need > c->wave_slots) {
        cudaFree(c->d_wave_paths);
        c->d_wave_paths = nullptr; c->wave_slots = 0;
        if (cudaMalloc(&c->d_wave_paths, need * cont_f4(c->bands) * sizeof(float4)) != cudaSuccess) {
            cudaGetLastError();
            c->d_wave_paths = nullptr;
            return ARV2_OK;                                   // not enough memory: depth-first kernel
        }
        c->wave_slots = need;
    }
    p->wave_paths = c->d_wave_paths; p->wave_cap = cap; p->wave_queues = nq; p->wave_segments = per;
    return ARV2_OK;
}

void free_sweep(arv2_ctx* c)
{
    cudaFree(c->sweep.state[0]); cudaFree(c->sweep.state[1]); cudaFree(c->sweep.key); cudaFree(c->sweep.rank); cudaFree(c->sweep.perm);
    cudaFree(c->sweep.bins); cudaFree(c->sweep.tile_sums); cudaFree(c->sweep.count);
    c->sweep = SweepWork{}; c->sweep_rays = 0; c->sweep_nbins = 0;
}

// Workspace of the bounce-synchronous tracer for a launch of n_rays rays; returns false (and the launch uses wave_kernel)
// when the launch is too small for it, the depth is absurd or the memory is not there.
bool ensure_sweep(arv2_ctx* c, long long n_rays)
{
    if (c->sweep_min_rays <= 0 || n_rays < c->sweep_min_rays || n_rays > 0x7fffffffLL || c->n_scene <= 0) return false;
    SweepWork& w = c->sweep;
    // 8 segments for the fresh bundles, then 2 per sweep (2 and 3 trace equally fast; 2 keeps 18.5-23 of 32 lanes busy per
    // instruction at every depth, 3 drops to 17.6 in the deep sweeps), bins = 8^3 cells x 32^2 direction cells
    // (profiles/r09_trace_sweeps.md)
    int seg = 2, first = 8, cb = n_rays >= (32LL << 20) ? 4 : 3, db = 5, dm = 0;      // 16^3 cells from 32 M rays on (100 M rays: +2.6 %; 10 M: no difference)
    if (const char* e = getenv("ARV2_SWEEP_SEGMENTS")) seg = atoi(e) > 0 ? atoi(e) : seg;          // tuning aids
    if (const char* e = getenv("ARV2_SWEEP_FIRST")) first = atoi(e) > 0 ? atoi(e) : first;
    if (const char* e = getenv("ARV2_SWEEP_CELL_BITS")) cb = atoi(e) >= 0 ? atoi(e) : cb;
    if (const char* e = getenv("ARV2_SWEEP_DIR_BITS")) db = atoi(e) >= 0 ? atoi(e) : db;
    if (const char* e = getenv("ARV2_SWEEP_DIR_MAJOR")) dm = atoi(e) != 0;
    if (cb > 8 || db > 8 || 3 * cb + 2 * db > 22) { cb = 3; db = 5; }
    if (((long long)c->max_bounces + seg - 1) / seg > 4096) return false;
    const int n_bins = sweep_bins(cb, db);
    if (n_rays > c->sweep_rays || n_bins > c->sweep_nbins) {
        free_sweep(c);
        const size_t st = (size_t)n_rays * cont_f4(c->bands) * sizeof(float4);
        bool ok = cudaMalloc(&w.state[0], st) == cudaSuccess && cudaMalloc(&w.state[1], st) == cudaSuccess &&
                  cudaMalloc(&w.key, (size_t)n_rays * 4) == cudaSuccess && cudaMalloc(&w.rank, (size_t)n_rays * 4) == cudaSuccess &&
                  cudaMalloc(&w.perm, (size_t)n_rays * 4) == cudaSuccess && cudaMalloc(&w.bins, (size_t)n_bins * 4) == cudaSuccess && cudaMalloc(&w.tile_sums, (size_t)(n_bins / 4096) * 4) == cudaSuccess &&
                  cudaMalloc(&w.count, 2 * sizeof(unsigned long long)) == cudaSuccess;
        if (!ok) { cudaGetLastError(); free_sweep(c); return false; }
        c->sweep_rays = n_rays; c->sweep_nbins = n_bins;
    }
    w.segments = seg; w.first_segments = first; w.cell_bits = cb; w.dir_bits = db; w.dir_major = dm;
    for (int a = 0; a < 3; ++a) {
        const float lo = c->scene_bvh.lo[a], hi = c->scene_bvh.hi[a];
        w.lo[a] = lo;
        w.scale[a] = (float)(1 << cb) / std::max(hi - lo, 1e-6f);
    }
    return true;
}

void free_cache(arv2_ctx* c)
{
    cudaFree(c->d_pc_seg); cudaFree(c->d_pc_energy); cudaFree(c->d_pc_off); cudaFree(c->d_pc_vert); cudaFree(c->d_pc_bits);
    c->d_pc_seg = nullptr; c->d_pc_energy = nullptr; c->d_pc_off = nullptr; c->d_pc_vert = nullptr; c->d_pc_bits = nullptr;
    c->pc_segs = 0; c->pc_nvert = 0; c->cache_valid = false;
}

int finish_timed(arv2_ctx* c, double* ms);

// Trace the receiver-independent paths of the whole seeded set (scene tree only, wave_kernel mode 1) into a ray-major
// scratch cache, then pack it: CSR records + energies, and the 8 B path vertices rr_mask_kernel streams.  The scratch
// (max_bounces records per ray) is freed again; what stays is 36 B + 4 B x bands per traced segment.
int build_cache(arv2_ctx* c, double* ms)
{
    const long long n = c->n_rays_total;
    if (c->max_bounces > 65535u || n > 0x7fffffffLL || (double)n * (c->max_bounces + 1.0) >= 4.0e9) {
        set_error("path cache: max_bounces <= 65535, at most 2^31 rays and fewer than 4e9 path vertices"); return ARV2_ERR_INVALID;
    }
    free_cache(c);
    int rc = upload_receiver(c);
    if (rc != ARV2_OK) return rc;
    rc = ensure_ray_order(c, 0, n);
    if (rc != ARV2_OK) return rc;
    const size_t slots = (size_t)n * std::max(1u, c->max_bounces);
    float4* t_seg = nullptr; float* t_energy = nullptr; int* t_nseg = nullptr; unsigned long long* scratch = nullptr;
    auto drop = [&]() { cudaFree(t_seg); cudaFree(t_energy); cudaFree(t_nseg); cudaFree(scratch); };
#define CKB(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(e_)); drop(); free_cache(c); return ARV2_ERR_CUDA; } } while (0)
    CKB(cudaMalloc(&t_seg, slots * 2 * sizeof(float4)));
    CKB(cudaMalloc(&t_energy, slots * sizeof(float) * c->bands));
    CKB(cudaMalloc(&t_nseg, (size_t)n * sizeof(int)));
    CKB(cudaMalloc(&c->d_pc_off, ((size_t)n + 1) * sizeof(unsigned long long)));
    const long long tiles = (n + 2047) / 2048;
    CKB(cudaMalloc(&scratch, ((size_t)tiles + 1) * sizeof(unsigned long long)));
    CKB(cudaMemsetAsync(c->d_counters, 0, kCounters * sizeof(unsigned long long), c->stream));
    TraceParams p;
    fill_params(c, &p, 0, n);
    p.rec_bin = nullptr; p.rec_ear = nullptr; p.rec_energy = nullptr; p.rec_nseg = nullptr;
    p.pc_seg = t_seg; p.pc_energy = t_energy; p.pc_nseg = t_nseg; p.pc_stride = (long long)std::max(1u, c->max_bounces);
    const bool sweeps = ensure_sweep(c, n);
    if (!sweeps) rc = ensure_wave(c, &p, n);
    if (rc != ARV2_OK) { drop(); return rc; }
    CKB(cudaEventRecord(c->ev0, c->stream));
    if (sweeps) CKB(launch_trace_sweeps(p, c->sweep, c->bands, 1, c->stream));
    else CKB(launch_trace(p, c->bands, 1, c->sm_count, c->stream));
    CKB(launch_cache_offsets(t_nseg, n, c->d_pc_off, scratch, c->stream));
    unsigned long long total = 0;
    CKB(cudaMemcpyAsync(&total, c->d_pc_off + n, sizeof total, cudaMemcpyDeviceToHost, c->stream));
    CKB(cudaStreamSynchronize(c->stream));
    c->pc_segs = (long long)total; c->pc_nvert = (long long)total + n;
    // records and energies share the vertices' index space (one unused slot per ray): a flagged vertex is its own record index
    CKB(cudaMalloc(&c->d_pc_seg, (size_t)c->pc_nvert * 2 * sizeof(float4)));
    CKB(cudaMalloc(&c->d_pc_energy, (size_t)c->pc_nvert * sizeof(float) * c->bands));
    // vertices padded to whole 128-vertex tiles plus one (rr_mask_kernel reads tile by tile and one vertex beyond)
    const size_t vert_padded = ((size_t)c->pc_nvert + 127) / 128 * 128 + 128;
    CKB(cudaMalloc(&c->d_pc_vert, vert_padded * sizeof(uint2)));
    CKB(cudaMemsetAsync(c->d_pc_vert + c->pc_nvert, 0, (vert_padded - (size_t)c->pc_nvert) * sizeof(uint2), c->stream));
    CKB(cudaMalloc(&c->d_pc_bits, vert_padded / 32 * sizeof(unsigned)));
    // vertex grid: 16 bits per axis over the scene's bounds and the emitter (every path vertex lies on the scene or at the
    // emitter).  pc_eps = how far a true segment may be from the line through its two quantised vertices: a segment starts
    // 1 mm (x |dir|) off the wall the previous one ended on, each vertex moves by at most half a grid diagonal, float slop.
    float qmax = 0.f;
    for (int a = 0; a < 3; ++a) {
        float lo = c->emitter[a], hi = c->emitter[a];
        if (c->n_scene > 0) { lo = std::min(lo, c->scene_bvh.lo[a]); hi = std::max(hi, c->scene_bvh.hi[a]); }
        const float pad = std::max(0.01f, 1e-3f * (hi - lo));
        c->pc_q0[a] = lo - pad;
        c->pc_qs[a] = (hi - lo + 2.f * pad) / 65535.f;
        qmax = std::max(qmax, c->pc_qs[a]);
    }
    // (rr_mask_kernel folds the grid origin, the centre and 2^23 grid steps into one constant per axis: its rounding moves
    // every vertex by up to one grid step per axis, the same way)
    c->pc_eps = 1.2e-3f + (0.87f + 1.8f) * qmax + 1e-4f + 1e-5f * (65535.f * qmax);
    fill_params(c, &p, 0, n);
    p.pc_stride = (long long)std::max(1u, c->max_bounces);
    CKB(launch_cache_compact(p, t_seg, t_energy, c->d_pc_vert, c->bands, c->stream));
    CKB(cudaEventRecord(c->ev1, c->stream));
    rc = finish_timed(c, ms);
    drop();
#undef CKB
    if (rc != ARV2_OK) { free_cache(c); return rc; }
    c->cache_valid = true;
    return ARV2_OK;
}

int finish_timed(arv2_ctx* c, double* ms)
{
    CK(cudaMemcpyAsync(c->h_counters, c->d_counters, kCounters * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));                    // CUDA_SYNC_CHECK, OR/AudioRenderer.cpp:511
    c->last_segments = (long long)c->h_counters[1];
    if (c->h_counters[7] != 0) { set_error("trace: path-queue watchdog tripped (a warp waited too long for a queued path)"); return ARV2_ERR_CUDA; }
    if (getenv("ARV2_TAILSTAT") && c->h_counters[13]) fprintf(stderr, "tail: CTA busy time min %.3f ms, mean %.3f ms over %llu CTAs\n", (double)c->h_counters[14] * 1e-6,
                                                                (double)c->h_counters[12] / (double)c->h_counters[13] * 1e-6, c->h_counters[13]);
    if (getenv("ARV2_TAILSTAT")) fprintf(stderr, "tail: pool empty -> first warp exit %.3f ms, -> last warp exit %.3f ms\n",
                                         ((double)c->h_counters[6] - (double)c->h_counters[4]) * 1e-6, ((double)c->h_counters[5] - (double)c->h_counters[4]) * 1e-6);
    if (getenv("ARV2_PRINT_STATS")) { for (int i = 0; i < kCounters; ++i) fprintf(stderr, "%llu ", c->h_counters[i]); fprintf(stderr, "\n"); }
    if (ms) { float t = 0.f; CK(cudaEventElapsedTime(&t, c->ev0, c->ev1)); *ms = t; }
    return ARV2_OK;
}

} // namespace

extern "C" {

const char* arv2_last_error(void) { return g_error.c_str(); }
const char* arv2_version(void) { return "arv2-b200 0.2 (sm_100a)"; }

/* ------------------------------------------------------------------ scene -- */
int arv2_scene_load_obj(const char* path, arv2_scene** out)
{
    REQUIRE(path && out, "arv2_scene_load_obj: null argument");
    auto* s = new arv2_scene;
    std::string err;
    const int rc = load_obj(path, &s->s, &err);
    if (rc != ARV2_OK) { set_error(err); delete s; return rc; }
    *out = s;
    return ARV2_OK;
}

int arv2_scene_from_triangles(const float* tv, const int32_t* tm, int64_t n, const char* const* names, int32_t n_meshes,
                              arv2_scene** out)
{
    REQUIRE(out && n >= 0 && n_meshes >= 0 && (n == 0 || (tv && tm)), "arv2_scene_from_triangles: bad argument");
    auto* s = new arv2_scene;
    s->s.tri_verts.assign(tv, tv + 9 * n);
    s->s.tri_mesh.assign(tm, tm + n);
    for (int64_t i = 0; i < n; ++i)
        if (tm[i] < 0 || tm[i] >= n_meshes) { delete s; set_error("tri_mesh out of range"); return ARV2_ERR_INVALID; }
    for (int32_t m = 0; m < n_meshes; ++m) s->s.mesh_material.push_back(names && names[m] ? names[m] : "");
    *out = s;
    return ARV2_OK;
}

int arv2_scene_counts(const arv2_scene* s, int64_t* n_tris, int32_t* n_meshes)
{
    REQUIRE(s, "null scene");
    if (n_tris) *n_tris = s->s.n_tris();
    if (n_meshes) *n_meshes = (int32_t)s->s.mesh_material.size();
    return ARV2_OK;
}

int arv2_scene_get_triangles(const arv2_scene* s, float* tv, int32_t* tm)
{
    REQUIRE(s, "null scene");
    if (tv) std::memcpy(tv, s->s.tri_verts.data(), s->s.tri_verts.size() * sizeof(float));
    if (tm) std::memcpy(tm, s->s.tri_mesh.data(), s->s.tri_mesh.size() * sizeof(int32_t));
    return ARV2_OK;
}

const char* arv2_scene_mesh_material(const arv2_scene* s, int32_t mesh)
{
    if (!s || mesh < 0 || mesh >= (int32_t)s->s.mesh_material.size()) return nullptr;
    return s->s.mesh_material[mesh].c_str();
}

int arv2_scene_bounds(const arv2_scene* s, float* lo, float* hi)
{
    REQUIRE(s && lo && hi, "null argument");
    for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
    const size_t nv = s->s.tri_verts.size() / 3;
    for (size_t i = 0; i < nv; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::fmin(lo[a], s->s.tri_verts[3 * i + a]);
            hi[a] = std::fmax(hi[a], s->s.tri_verts[3 * i + a]);
        }
    return ARV2_OK;
}

int arv2_scene_bvh_stats(const arv2_scene* s, arv2_bvh_stats* out)
{
    REQUIRE(s && out, "null argument");
    const std::vector<float>& tv = s->s.tri_verts;
    const int64_t n = s->s.n_tris();
    HostBvh b;
    build_bvh_sah(tv.data(), n, &b, 8);
    std::memset(out, 0, sizeof *out);
    out->n_tris = n; out->n_nodes = (int64_t)b.nodes.size(); out->depth = bvh2_depth(b);
    auto area = [](const float* lo, const float* hi) {
        const double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    };
    const double root_area = n > 0 ? area(b.lo, b.hi) : 0.0;
    std::vector<int32_t> seen((size_t)n, 0);
    bool ok = b.order.size() == (size_t)n && !b.nodes.empty();
    double cn = n > 0 ? 1.0 : 0.0, ct = 0.0;
    for (size_t i = 0; i < b.nodes.size() && ok; ++i) {
        const BvhNode& nd = b.nodes[i];
        int32_t c[4];
        std::memcpy(c, &nd.q[12], sizeof c);
        for (int w = 0; w < 2; ++w) {
            const float lo[3] = {nd.q[w * 4 + 0], nd.q[w * 4 + 2], nd.q[8 + w * 2]}, hi[3] = {nd.q[w * 4 + 1], nd.q[w * 4 + 3], nd.q[8 + w * 2 + 1]};
            if (lo[0] == kEmptyBox) continue;                         // absent child
            const double a = root_area > 0.0 ? area(lo, hi) / root_area : 0.0;
            if (c[w] >= 0) {
                cn += a;
                if (c[w] <= (int32_t)i || c[w] >= (int32_t)b.nodes.size()) ok = false;
            } else {
                const int32_t code = ~c[w];
                const int64_t first = code >> kLeafShift;
                const int cnt = (code & 7) + 1;
                out->n_leaves++;
                out->max_leaf_tris = std::max(out->max_leaf_tris, cnt);
                ct += a * cnt;
                if (first < 0 || first + cnt > n) { ok = false; break; }
                for (int t = 0; t < cnt; ++t) {
                    const int32_t id = b.order[(size_t)(first + t)];
                    if (id < 0 || id >= n) { ok = false; break; }
                    seen[(size_t)id]++;
                    for (int k = 0; k < 3; ++k)
                        for (int ax = 0; ax < 3; ++ax) {
                            const float v = tv[9 * (size_t)id + 3 * k + ax];
                            if (v < lo[ax] || v > hi[ax]) ok = false;
                        }
                }
            }
        }
    }
    for (int64_t i = 0; i < n && ok; ++i) if (seen[(size_t)i] != 1) ok = false;
    out->valid = ok ? 1 : 0; out->sah_nodes = cn; out->sah_tris = ct;
    return ARV2_OK;
}

void arv2_scene_destroy(arv2_scene* s) { delete s; }

int arv2_receiver_load(const char* left_obj, const char* right_obj, arv2_receiver** out)
{
    REQUIRE(left_obj && right_obj && out, "arv2_receiver_load: null argument");
    HostScene l, r;
    std::string err;
    int rc = load_obj(left_obj, &l, &err);
    if (rc == ARV2_OK) rc = load_obj(right_obj, &r, &err);
    if (rc != ARV2_OK) { set_error("Could not read sphere OBJ model: " + err); return rc; }
    auto* o = new arv2_receiver;
    // HalfSphere keeps shapes[0] only (OR/OptixModel.cpp:199-203); both assets have one mesh
    auto first_mesh = [](const HostScene& s, std::vector<float>* dst) {
        for (int64_t i = 0; i < s.n_tris() && s.tri_mesh[i] == 0; ++i) dst->insert(dst->end(), s.tri_verts.begin() + 9 * i, s.tri_verts.begin() + 9 * i + 9);
    };
    first_mesh(l, &o->r.left);
    first_mesh(r, &o->r.right);
    *out = o;
    return ARV2_OK;
}

int arv2_receiver_from_triangles(const float* left, int64_t nl, const float* right, int64_t nr, arv2_receiver** out)
{
    REQUIRE(out && nl >= 0 && nr >= 0 && (nl == 0 || left) && (nr == 0 || right), "arv2_receiver_from_triangles: bad argument");
    auto* o = new arv2_receiver;
    o->r.left.assign(left, left + 9 * nl);
    o->r.right.assign(right, right + 9 * nr);
    *out = o;
    return ARV2_OK;
}

int arv2_receiver_counts(const arv2_receiver* r, int64_t* nl, int64_t* nr)
{
    REQUIRE(r, "null receiver");
    if (nl) *nl = (int64_t)r->r.left.size() / 9;
    if (nr) *nr = (int64_t)r->r.right.size() / 9;
    return ARV2_OK;
}

int arv2_receiver_place(const arv2_receiver* r, const float cam[3], float rotation_deg, float* left_out, float* right_out)
{
    REQUIRE(r && cam, "null argument");
    if (left_out) place_receiver_half(r->r.left, cam, rotation_deg, left_out);
    if (right_out) place_receiver_half(r->r.right, cam, rotation_deg, right_out);
    return ARV2_OK;
}

void arv2_receiver_destroy(arv2_receiver* r) { delete r; }

float arv2_material_absorption(const char* name, const arv2_material* mats, int32_t n)
{
    return material_absorption(name ? name : "", mats, mats ? n : 0);
}

/* ----------------------------------------------------------------- config -- */
int arv2_config_parse(const char* json, arv2_config* out)
{
    REQUIRE(json && out, "null argument");
    std::string err;
    const int rc = parse_config(json, out, &err);
    if (rc != ARV2_OK) set_error(err);
    return rc;
}

int arv2_config_load(const char* path, arv2_config* out)
{
    REQUIRE(path && out, "null argument");
    std::ifstream in(path);
    if (!in) { set_error(std::string("cannot open ") + path); return ARV2_ERR_IO; }
    std::stringstream ss;
    ss << in.rdbuf();
    return arv2_config_parse(ss.str().c_str(), out);
}

/* --------------------------------------------------------------- renderer -- */
int arv2_create(const arv2_scene* scene, const arv2_receiver* receiver, const arv2_renderer_desc* desc, arv2_ctx** out)
{
    REQUIRE(scene && desc && out, "arv2_create: null argument");
    REQUIRE(desc->bands == 1 || desc->bands == ARV2_MAX_BANDS, "bands must be 1 or 8");
    REQUIRE(desc->sample_rate > 0 && desc->ir_length_in_seconds > 0, "bad ir length / sample rate");
    REQUIRE(desc->rays_x > 0 && desc->rays_y > 0 && desc->rays_z > 0, "bad ray counts");
    const long long n_total = (long long)desc->rays_x * desc->rays_y * desc->rays_z;
    REQUIRE(n_total <= 2147483647LL, "x*y*z must fit the reference's int launch size");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    REQUIRE(desc->device >= 0 && desc->device < ndev, "no such CUDA device");
    CK(cudaSetDevice(desc->device));

    auto* c = new arv2_ctx;
    if (getenv("ARV2_NO_SORT")) c->coherent_order = false;   // tuning aids (A/B)
    if (getenv("ARV2_NO_WAVE")) c->wave = false;
    if (getenv("ARV2_SHARD_CONTIGUOUS")) c->shard_mode = 0;      // A/B: multi-GPU shards are contiguous slices of ray ids
    if (const char* e = getenv("ARV2_SWEEP")) c->sweep_min_rays = atoll(e);      // tuning aid: launches of >= this many rays are traced in sweeps (0 = never)
    if (getenv("ARV2_RR_SERIAL")) c->rr_serial = true;
    c->desc = *desc; c->desc.materials = nullptr; c->desc.n_materials = 0;
    c->device = desc->device;
    c->bands = desc->bands;
    c->ir_len = (int)(desc->ir_length_in_seconds * (unsigned)desc->sample_rate);  // OR/AudioRenderer.cpp:78
    c->n_rays_total = n_total;
    c->scene = scene->s;
    c->n_scene = c->scene.n_tris();
    if (receiver) {
        c->receiver = receiver->r; c->has_receiver = true;
        c->n_left = (int64_t)c->receiver.left.size() / 9; c->n_right = (int64_t)c->receiver.right.size() / 9;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    int rc = ARV2_OK;
    auto fail = [&](int code) { arv2_destroy(c); return code; };
#define CKC(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(e_)); return fail(ARV2_ERR_CUDA); } } while (0)
    CKC(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CKC(cudaEventCreate(&c->ev0));
    CKC(cudaEventCreate(&c->ev1));

    // per-mesh materials: keep = 1 - absorption (OR/devicePrograms.cu:174), scattering
    const int n_mesh = (int)c->scene.mesh_material.size();
    std::vector<float> keep((size_t)std::max(1, n_mesh) * c->bands, 0.5f), scat((size_t)std::max(1, n_mesh), 0.f);
    for (int m = 0; m < n_mesh; ++m) {
        const std::string& name = c->scene.mesh_material[m];
        // getMaterialAbsorption (OR/AudioRenderer.cpp:34-56) turns a mesh named receiver_left / receiver_right, or one
        // with a negative configured absorption, into a receiver (mat_absorption < 0, OR/devicePrograms.cu:91).  Here
        // receivers are their own object (arv2_receiver: refit per move, left out of the path cache), so such a scene
        // is rejected instead of silently tracing those meshes as walls with 1 - a > 1.
        if (material_absorption(name, desc->materials, desc->n_materials) < 0.f) {
            set_error("scene mesh '" + name + "' resolves to a receiver material (name receiver_left/right or absorption < 0): pass receivers through arv2_receiver");
            return fail(ARV2_ERR_INVALID);
        }
        const arv2_material* hit = nullptr;
        for (int i = 0; i < desc->n_materials && !hit; ++i)
            if (desc->materials[i].name && name == desc->materials[i].name) hit = &desc->materials[i];
        for (int b = 0; b < c->bands; ++b) {
            const float a = hit ? hit->mat_absorption[b] : 0.5f;      // default, OR/AudioRenderer.cpp:54-55
            keep[(size_t)m * c->bands + b] = 1.0f - a;
        }
        scat[m] = hit ? hit->scattering : 0.f;
        if (scat[m] > 0.f) c->any_scatter = 1;
    }

    // scene BVH (built once; OR/AudioRenderer.cpp:95-218 rebuilds on every move): binary tree
    // from the host binned-SAH builder or the GPU LBVH builder (K1), laid out as the 64 B
    // nodes the kernels traverse (arv2_internal.h)
    const int64_t n_recv = c->n_left + c->n_right;
    c->n_recv_nodes = (int32_t)std::max<int64_t>(1, n_recv);
    const bool gpu_build = desc->bvh_builder == 1 && c->n_scene > kMaxLeafTris;
    const unsigned hc = std::thread::hardware_concurrency();
    HostBvh& scene2 = c->scene_bvh;
    if (!gpu_build) {
        build_bvh_sah(c->scene.tri_verts.data(), c->n_scene, &scene2, hc ? (int)hc : 1);
    } else {
        float* d_v = nullptr; int* d_m = nullptr; float4* d_n2 = nullptr; int* d_order = nullptr;
        const size_t n = (size_t)c->n_scene;
        cudaError_t e = cudaMalloc(&d_v, n * 9 * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&d_m, n * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&d_n2, n * 4 * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&d_order, n * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(d_v, c->scene.tri_verts.data(), n * 9 * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_m, c->scene.tri_mesh.data(), n * sizeof(int), cudaMemcpyHostToDevice);
        LbvhResult res{};
        if (e == cudaSuccess) e = build_bvh_lbvh(d_v, d_m, (int)n, 0, d_n2, nullptr, d_order, 0, 0, &res, c->stream);
        if (e == cudaSuccess) {
            scene2.nodes.resize((size_t)res.n_nodes);
            scene2.order.resize(n);
            e = cudaMemcpy(scene2.nodes.data(), d_n2, (size_t)res.n_nodes * sizeof(BvhNode), cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(scene2.order.data(), d_order, n * sizeof(int), cudaMemcpyDeviceToHost);
            for (int a = 0; a < 3; ++a) { scene2.lo[a] = res.lo[a]; scene2.hi[a] = res.hi[a]; }
        }
        cudaFree(d_v); cudaFree(d_m); cudaFree(d_n2); cudaFree(d_order);
        if (e != cudaSuccess) { set_error(std::string("build_bvh_lbvh: ") + cudaGetErrorString(e)); return fail(ARV2_ERR_CUDA); }
    }
    c->n_scene_nodes = (int32_t)scene2.nodes.size();
    const size_t total_nodes = 1 + (size_t)c->n_scene_nodes + (size_t)c->n_recv_nodes;
    const size_t total_tris = (size_t)std::max<int64_t>(1, c->n_scene + n_recv);
    CKC(cudaMalloc(&c->d_nodes, total_nodes * sizeof(BvhNode)));
    CKC(cudaMalloc(&c->d_tris, total_tris * 4 * sizeof(float4)));
    CKC(cudaMalloc(&c->d_keep, keep.size() * sizeof(float)));
    CKC(cudaMalloc(&c->d_scatter, scat.size() * sizeof(float)));
    CKC(cudaMemcpy(c->d_keep, keep.data(), keep.size() * sizeof(float), cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(c->d_scatter, scat.data(), scat.size() * sizeof(float), cudaMemcpyHostToDevice));
    {
        if (bvh2_depth(scene2) + 3 > kTraversalStack) { set_error("scene BVH too deep for the traversal stack"); return fail(ARV2_ERR_INVALID); }
        std::vector<BvhNode> nodes(scene2.nodes.size());
        offset_nodes(scene2, 1, 0, nodes.data());
        if (getenv("ARV2_BVH4") && trace_supports_wide_nodes() == 1 && c->n_scene > kMaxLeafTris) {      // experiment (r07 section 18, slower): 4-wide scene tree next to the binary one
            std::vector<Bvh4Node> wide;
            const int depth4 = collapse_bvh4(scene2, &wide);
            if (3 * depth4 + 3 > kTraversalStack) { set_error("wide scene BVH too deep for the traversal stack"); return fail(ARV2_ERR_INVALID); }
            CKC(cudaMalloc(&c->d_nodes4, wide.size() * sizeof(Bvh4Node)));
            CKC(cudaMemcpy(c->d_nodes4, wide.data(), wide.size() * sizeof(Bvh4Node), cudaMemcpyHostToDevice));
            c->scene_root_code = kWideBit | 0;
        }
        if (getenv("ARV2_QNODES") && trace_supports_wide_nodes() == 2 && c->n_scene > kMaxLeafTris) {    // experiment (r07 section 22): 32 B quantised scene nodes
            std::vector<Q16Node> qn;
            quantise_bvh2(scene2, &qn, c->qk, c->qinvk, c->qc);
            CKC(cudaMalloc(&c->d_nodes4, qn.size() * sizeof(Q16Node)));
            CKC(cudaMemcpy(c->d_nodes4, qn.data(), qn.size() * sizeof(Q16Node), cudaMemcpyHostToDevice));
            c->scene_root_code = kWideBit | 0;
        }
        const BvhNode top = make_top_node(c, false);
        CKC(cudaMemcpy(c->d_nodes, &top, sizeof top, cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(c->d_nodes + 4, nodes.data(), nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice));      // scene nodes behind the top node
        std::vector<float> recs((size_t)c->n_scene * 16);
        for (int64_t s = 0; s < c->n_scene; ++s) {
            const int32_t src = scene2.order[s];
            make_tri_record(c->scene.tri_verts.data() + 9 * (size_t)src, src, c->scene.tri_mesh[src], recs.data() + 16 * (size_t)s);
        }
        if (c->n_scene) CKC(cudaMemcpy(c->d_tris, recs.data(), recs.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    c->stage_f4 = 4 + 4 * (size_t)c->n_recv_nodes + 4 * (size_t)std::max<int64_t>(1, n_recv);
    CKC(cudaMallocHost(&c->h_stage, c->stage_f4 * sizeof(float4)));
    CKC(cudaMallocHost(&c->h_counters, kCounters * sizeof(unsigned long long)));

    // IR buffers (OR/AudioRenderer.cpp:81-85) and the fp64 accumulation histogram
    const size_t irn = (size_t)c->bands * c->ir_len;
    CKC(cudaMalloc(&c->d_hist, 2 * irn * sizeof(double)));
    CKC(cudaMalloc(&c->d_ir_l, 2 * irn * sizeof(float)));
    c->d_ir_r = c->d_ir_l + irn;
    CKC(cudaMallocHost(&c->h_ir, 2 * irn * sizeof(float)));
    CKC(cudaMemset(c->d_hist, 0, 2 * irn * sizeof(double)));
    CKC(cudaMemset(c->d_ir_l, 0, 2 * irn * sizeof(float)));
    CKC(cudaMalloc(&c->d_counters, kCounters * sizeof(unsigned long long)));
#undef CKC
    (void)rc;
    *out = c;
    return ARV2_OK;
}

void arv2_destroy(arv2_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_nodes); cudaFree(c->d_nodes4); cudaFree(c->d_tris); cudaFree(c->d_keep); cudaFree(c->d_scatter);
    cudaFree(c->d_hist); cudaFree(c->d_ir_l); cudaFree(c->d_counters);
    if (c->h_ir) cudaFreeHost(c->h_ir);
    cudaFree(c->d_rec_bin); cudaFree(c->d_rec_ear); cudaFree(c->d_rec_nseg); cudaFree(c->d_rec_energy);
    cudaFree(c->d_pc_seg); cudaFree(c->d_pc_energy); cudaFree(c->d_pc_off); cudaFree(c->d_pc_vert); cudaFree(c->d_pc_bits); free_ray_orders(c); free_tiles(c); cudaFree(c->d_wave_paths); free_sweep(c);
    cudaFree(c->conv.d_tw); cudaFree(c->conv.d_x); cudaFree(c->conv.d_out); cudaFree(c->conv.d_X); cudaFree(c->conv.d_H);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int arv2_set_emitter(arv2_ctx* c, float x, float y, float z)
{
    REQUIRE(c, "null ctx");
    c->emitter[0] = x; c->emitter[1] = y; c->emitter[2] = z;
    c->cache_valid = false;
    return ARV2_OK;
}

int arv2_set_receiver(arv2_ctx* c, float x, float y, float z, float yaw_deg)
{
    REQUIRE(c, "null ctx");
    c->center[0] = x; c->center[1] = y; c->center[2] = z; c->yaw = yaw_deg;
    c->recv_dirty = true;
    return ARV2_OK;
}

int arv2_set_thresholds(arv2_ctx* c, float energy, uint32_t max_bounces)
{
    REQUIRE(c, "null ctx");
    c->energy_thres = energy; c->max_bounces = max_bounces; c->cache_valid = false;
    return ARV2_OK;
}

int arv2_set_base_power(arv2_ctx* c, float v) { REQUIRE(c, "null ctx"); c->base_power = v; c->cache_valid = false; return ARV2_OK; }
int arv2_set_hrtf_absorption_rate(arv2_ctx* c, float v) { REQUIRE(c, "null ctx"); c->hrtf = v; return ARV2_OK; }
int arv2_set_mono(arv2_ctx* c, int32_t v) { REQUIRE(c, "null ctx"); c->mono = v ? 1 : 0; return ARV2_OK; }
int arv2_set_seed(arv2_ctx* c, uint64_t s) { REQUIRE(c, "null ctx"); c->seed = s; c->cache_valid = false; return ARV2_OK; }
int arv2_set_coherent_order(arv2_ctx* c, int32_t on) { REQUIRE(c, "null ctx"); c->coherent_order = on != 0; return ARV2_OK; }
int arv2_set_sweep_min_rays(arv2_ctx* c, int64_t n) { REQUIRE(c, "null ctx"); c->sweep_min_rays = n; return ARV2_OK; }
int arv2_set_stream(arv2_ctx* c, void* s) { REQUIRE(c, "null ctx"); c->stream = s ? (cudaStream_t)s : c->own_stream; return ARV2_OK; }

// Everything of a trace up to (not including) a host synchronisation: receiver upload, zeroing, the launch.
// `ids` (optional): the launch traces the n_rays rays ids[0 .. n_rays) of the seeded set (global ray ids, the order they
// are started in) instead of the contiguous range; per-ray records are then indexed by the global id.
static int enqueue_trace(arv2_ctx* c, int64_t ray_begin, int64_t n_rays, int32_t zero_first, const int* ids = nullptr)
{
    REQUIRE(ray_begin >= 0 && n_rays >= 0 && ray_begin + n_rays <= c->n_rays_total, "ray range outside the seeded set");
    CK(cudaSetDevice(c->device));
    int rc = upload_receiver(c);
    if (rc != ARV2_OK) return rc;
    const size_t irn = (size_t)c->bands * c->ir_len;
    if (zero_first) CK(cudaMemsetAsync(c->d_hist, 0, 2 * irn * sizeof(double), c->stream));   // fillZeros, OR/AudioRenderer.cpp:491-492
    CK(cudaMemsetAsync(c->d_counters, 0, kCounters * sizeof(unsigned long long), c->stream));
    if (getenv("ARV2_TAILSTAT")) { CK(cudaMemsetAsync(c->d_counters + 4, 0xFF, 8, c->stream)); CK(cudaMemsetAsync(c->d_counters + 6, 0xFF, 8, c->stream)); CK(cudaMemsetAsync(c->d_counters + 14, 0xFF, 8, c->stream)); }
    if (!ids) rc = ensure_ray_order(c, ray_begin, n_rays);
    if (rc != ARV2_OK) return rc;
    rc = ensure_records(c, ids ? c->n_rays_total : n_rays);
    if (rc != ARV2_OK) return rc;
    TraceParams p;
    fill_params(c, &p, ray_begin, n_rays);
    if (ids) p.ray_order = ids;
    const bool sweeps = ensure_sweep(c, n_rays);
    if (!sweeps) rc = ensure_wave(c, &p, n_rays);
    if (rc != ARV2_OK) return rc;
    CK(cudaEventRecord(c->ev0, c->stream));
    if (n_rays > 0 && sweeps) CK(launch_trace_sweeps(p, c->sweep, c->bands, 0, c->stream));
    else if (n_rays > 0) CK(launch_trace(p, c->bands, 0, c->sm_count, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    c->last_range_rays = ids ? c->n_rays_total : n_rays;
    return ARV2_OK;
}

// One rank's direction tiles of the seeded set, traced into the histogram (no exchange, no finalise).
static int enqueue_tiles(arv2_ctx* c, int32_t rank, int32_t n_ranks, int32_t zero_first)
{
    REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad rank");
    REQUIRE(c->n_rays_total <= 0x7fffffffLL, "direction tiles: at most 2^31 - 1 rays");
    CK(cudaSetDevice(c->device));
    const int rc = ensure_tiles(c, rank, n_ranks);
    if (rc != ARV2_OK) return rc;
    return enqueue_trace(c, 0, c->tiles.n, zero_first, c->tiles.d_ids);
}

int arv2_render_tiles(arv2_ctx* c, int32_t rank, int32_t n_ranks, int32_t zero_first, double* ms)
{
    REQUIRE(c, "null ctx");
    const int rc = enqueue_tiles(c, rank, n_ranks, zero_first);
    return rc != ARV2_OK ? rc : finish_timed(c, ms);
}

int arv2_set_shard_mode(arv2_ctx* c, int32_t mode) { REQUIRE(c && (mode == 0 || mode == 1), "shard mode is 0 (contiguous) or 1 (direction tiles)"); c->shard_mode = mode; return ARV2_OK; }

int arv2_render_range(arv2_ctx* c, int64_t ray_begin, int64_t n_rays, int32_t zero_first, double* ms)
{
    REQUIRE(c, "null ctx");
    const int rc = enqueue_trace(c, ray_begin, n_rays, zero_first);
    return rc != ARV2_OK ? rc : finish_timed(c, ms);
}

int arv2_finalize(arv2_ctx* c)
{
    REQUIRE(c, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(launch_finalize(c->d_hist, c->bands, c->ir_len, c->mono, c->d_ir_l, c->d_ir_r, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ARV2_OK;
}

int arv2_rerender(arv2_ctx* c, double* ms)
{
    REQUIRE(c, "null ctx");
    if (!c->desc.path_cache || !c->cache_valid) { set_error("arv2_rerender: no valid path cache (desc.path_cache + arv2_render first)"); return ARV2_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    int rc = upload_receiver(c);
    if (rc != ARV2_OK) return rc;
    const size_t irn = (size_t)c->bands * c->ir_len;
    rc = ensure_records(c, c->n_rays_total);
    if (rc != ARV2_OK) return rc;
    TraceParams p;
    fill_params(c, &p, 0, c->n_rays_total);
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaMemsetAsync(c->d_hist, 0, 2 * irn * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->d_counters, 0, kCounters * sizeof(unsigned long long), c->stream));
    if (c->rr_serial) CK(launch_rerender_serial(p, c->bands, c->sm_count, c->stream));
    else {
        CK(cudaMemsetAsync(c->d_pc_bits, 0, ((size_t)c->pc_nvert + 127) / 128 * 128 / 32 * sizeof(unsigned) + 128 / 32 * sizeof(unsigned), c->stream));
        CK(launch_rerender(p, c->bands, c->sm_count, c->stream));
    }
    CK(launch_finalize(c->d_hist, c->bands, c->ir_len, c->mono, c->d_ir_l, c->d_ir_r, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    c->last_range_rays = c->n_rays_total;
    return finish_timed(c, ms);
}

int arv2_render(arv2_ctx* c, double* ms)
{
    REQUIRE(c, "null ctx");
    if (!c->desc.path_cache) {
        // trace, finalise, counters: one enqueue, one host synchronisation (OR/AudioRenderer.cpp:489-523)
        const int rc = enqueue_trace(c, 0, c->n_rays_total, 1);
        if (rc != ARV2_OK) return rc;
        CK(launch_finalize(c->d_hist, c->bands, c->ir_len, c->mono, c->d_ir_l, c->d_ir_r, c->stream));
        return finish_timed(c, ms);
    }
    // path-cache mode: trace the receiver-independent paths once, then re-deposit from them
    CK(cudaSetDevice(c->device));
    double ms_build = 0.0;
    if (!c->cache_valid) {
        const int rc = build_cache(c, &ms_build);
        if (rc != ARV2_OK) return rc;
    }
    double ms_scan = 0.0;
    const int rc = arv2_rerender(c, &ms_scan);
    if (ms) *ms = ms_build + ms_scan;
    return rc;
}

/* ------------------------------------------------------ multi-GPU (NCCL) -- */
#define NCK(expr)                                                                                      \
    do {                                                                                               \
        ncclResult_t r_ = (expr);                                                                      \
        if (r_ != ncclSuccess) { set_error(std::string(#expr) + ": " + api->GetErrorString(r_)); return ARV2_ERR_CUDA; } \
    } while (0)

int arv2_comm_unique_id(void* id128)
{
    REQUIRE(id128, "null argument");
    static_assert(sizeof(ncclUniqueId) == ARV2_COMM_ID_BYTES, "ncclUniqueId size");
    std::string err;
    const NcclApi* api = nccl_api(&err);
    if (!api) { set_error(err); return ARV2_ERR_STATE; }
    ncclUniqueId id;
    NCK(api->GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return ARV2_OK;
}

int arv2_comm_create(int32_t device, int32_t rank, int32_t n_ranks, const void* id128, arv2_comm** out)
{
    REQUIRE(out && id128 && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "arv2_comm_create: bad argument");
    std::string err;
    const NcclApi* api = nccl_api(&err);
    if (!api) { set_error(err); return ARV2_ERR_STATE; }
    CK(cudaSetDevice(device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    auto* m = new arv2_comm;
    m->device = device; m->rank = rank; m->n_ranks = n_ranks;
    const ncclResult_t r = api->CommInitRank(&m->comm, n_ranks, id, rank);
    if (r != ncclSuccess) { set_error(std::string("ncclCommInitRank: ") + api->GetErrorString(r)); delete m; return ARV2_ERR_CUDA; }
    // NCCL connects its channels at the first collective (~1 s on 8 GPUs): pay that here, not in the first render
    if (n_ranks > 1) {
        double* d = nullptr;
        if (cudaMalloc(&d, 16 * sizeof(double)) == cudaSuccess) {
            cudaMemset(d, 0, 16 * sizeof(double));
            api->AllReduce(d, d, 16, ncclFloat64, ncclSum, m->comm, (cudaStream_t)0);
            cudaStreamSynchronize((cudaStream_t)0);
            cudaFree(d);
        }
    }
    *out = m;
    return ARV2_OK;
}

void arv2_comm_destroy(arv2_comm* m)
{
    if (!m) return;
    if (m->comm) { if (const NcclApi* api = nccl_api(nullptr)) { cudaSetDevice(m->device); api->CommDestroy(m->comm); } }
    delete m;
}

int arv2_comm_info(const arv2_comm* m, int32_t* rank, int32_t* n_ranks, int32_t* nccl_version)
{
    REQUIRE(m, "null comm");
    if (rank) *rank = m->rank;
    if (n_ranks) *n_ranks = m->n_ranks;
    if (nccl_version) { int v = 0; if (const NcclApi* api = nccl_api(nullptr)) api->GetVersion(&v); *nccl_version = v; }
    return ARV2_OK;
}

// Sum a device float buffer over the ranks onto `root` (in place there): the stereo mix of the sources each rank convolves.
int arv2_comm_reduce_f32(arv2_comm* m, float* d_buf, size_t count, int32_t root, void* cuda_stream)
{
    REQUIRE(m && d_buf && root >= 0 && root < m->n_ranks, "arv2_comm_reduce_f32: bad argument");
    const NcclApi* api = nccl_api(nullptr);
    REQUIRE(api, "NCCL not loaded");
    CK(cudaSetDevice(m->device));
    if (m->n_ranks > 1) NCK(api->Reduce(d_buf, d_buf, count, ncclFloat32, ncclSum, root, m->comm, (cudaStream_t)cuda_stream));
    return ARV2_OK;
}

void arv2_shard_range(int64_t n_rays, int32_t rank, int32_t n_ranks, int64_t* begin, int64_t* count)
{
    const int64_t base = n_rays / n_ranks, rem = n_rays % n_ranks;
    if (begin) *begin = rank * base + std::min<int64_t>(rank, rem);
    if (count) *count = base + (rank < rem ? 1 : 0);
}

// One rank's part of a render over R GPUs: this rank's share of the seeded ray set (its direction tiles, or with
// arv2_set_shard_mode(ctx, 0) its contiguous slice of ray ids), the sum of the fp64
// histograms over NVLink, the fp32 IR -- enqueued back to back on the context's stream, one host synchronisation.
int arv2_render_sharded(arv2_ctx* c, arv2_comm* m, double* ms)
{
    REQUIRE(c && m, "arv2_render_sharded: null argument");
    REQUIRE(m->device == c->device, "communicator and context live on different devices");
    REQUIRE(!c->desc.path_cache, "arv2_render_sharded: not with desc.path_cache");
    const NcclApi* api = nccl_api(nullptr);
    REQUIRE(api, "NCCL not loaded");
    int rc;
    if (c->shard_mode == 1 && m->n_ranks > 1 && c->n_rays_total <= 0x7fffffffLL) {
        rc = enqueue_tiles(c, m->rank, m->n_ranks, 1);
    } else {
        int64_t begin = 0, count = 0;
        arv2_shard_range(c->n_rays_total, m->rank, m->n_ranks, &begin, &count);
        rc = enqueue_trace(c, begin, count, 1);
    }
    if (rc != ARV2_OK) return rc;
    const size_t n = 2 * (size_t)c->bands * c->ir_len;
    if (m->n_ranks > 1) NCK(api->AllReduce(c->d_hist, c->d_hist, n, ncclFloat64, ncclSum, m->comm, c->stream));
    CK(launch_finalize(c->d_hist, c->bands, c->ir_len, c->mono, c->d_ir_l, c->d_ir_r, c->stream));
    return finish_timed(c, ms);
}

// One process, several devices (the reference application is one process, OR/main.cpp:720-777): one context and one
// communicator rank per device, one host thread per device for a render.
int arv2_multi_create(const arv2_scene* scene, const arv2_receiver* receiver, const arv2_renderer_desc* desc, const int32_t* devices,
                      int32_t n_devices, arv2_multi** out)
{
    REQUIRE(scene && desc && devices && out && n_devices >= 1, "arv2_multi_create: bad argument");
    std::string err;
    const NcclApi* api = nccl_api(&err);
    if (!api) { set_error(err); return ARV2_ERR_STATE; }
    auto* mm = new arv2_multi;
    int rc = ARV2_OK;
    for (int i = 0; i < n_devices && rc == ARV2_OK; ++i) {
        arv2_renderer_desc d = *desc;
        d.device = devices[i];
        arv2_ctx* c = nullptr;
        rc = arv2_create(scene, receiver, &d, &c);
        if (rc == ARV2_OK) mm->ctx.push_back(c);
    }
    if (rc == ARV2_OK) {
        std::vector<ncclComm_t> comms((size_t)n_devices);
        std::vector<int> devs(devices, devices + n_devices);
        const ncclResult_t r = api->CommInitAll(comms.data(), n_devices, devs.data());
        if (r != ncclSuccess) { set_error(std::string("ncclCommInitAll: ") + api->GetErrorString(r)); rc = ARV2_ERR_CUDA; }
        else for (int i = 0; i < n_devices; ++i) {
            auto* m = new arv2_comm;
            m->device = devices[i]; m->rank = i; m->n_ranks = n_devices; m->comm = comms[(size_t)i];
            mm->comm.push_back(m);
        }
    }
    if (rc != ARV2_OK) { arv2_multi_destroy(mm); return rc; }
    *out = mm;
    return ARV2_OK;
}

int32_t arv2_multi_size(const arv2_multi* mm) { return mm ? (int32_t)mm->ctx.size() : 0; }
arv2_ctx* arv2_multi_ctx(arv2_multi* mm, int32_t i) { return (mm && i >= 0 && i < (int32_t)mm->ctx.size()) ? mm->ctx[(size_t)i] : nullptr; }

int arv2_multi_render(arv2_multi* mm, double* ms)
{
    REQUIRE(mm && !mm->ctx.empty(), "null multi");
    const size_t n = mm->ctx.size();
    std::vector<int> rcs(n, ARV2_OK);
    std::vector<double> t(n, 0.0);
    std::vector<std::string> errs(n);
    std::vector<std::thread> th;
    for (size_t i = 0; i < n; ++i)
        th.emplace_back([&, i] { rcs[i] = arv2_render_sharded(mm->ctx[i], mm->comm[i], &t[i]); if (rcs[i] != ARV2_OK) errs[i] = arv2_last_error(); });
    for (auto& x : th) x.join();
    double worst = 0.0;
    for (size_t i = 0; i < n; ++i) {
        if (rcs[i] != ARV2_OK) { set_error("device " + std::to_string(mm->ctx[i]->device) + ": " + errs[i]); return rcs[i]; }
        worst = std::max(worst, t[i]);
    }
    if (ms) *ms = worst;
    return ARV2_OK;
}

void arv2_multi_destroy(arv2_multi* mm)
{
    if (!mm) return;
    for (auto* m : mm->comm) arv2_comm_destroy(m);
    for (auto* c : mm->ctx) arv2_destroy(c);
    delete mm;
}

int arv2_ir_length(const arv2_ctx* c, int32_t* ir_length, int32_t* bands)
{
    REQUIRE(c, "null ctx");
    if (ir_length) *ir_length = c->ir_len;
    if (bands) *bands = c->bands;
    return ARV2_OK;
}

int arv2_get_ir(arv2_ctx* c, float* l, float* r)
{
    REQUIRE(c, "null ctx");
    CK(cudaSetDevice(c->device));
    // both ears in one stream-ordered copy into pinned memory, one sync (the reference's getIROnHostMem is two blocking copies)
    const size_t irn = (size_t)c->bands * c->ir_len;
    CK(cudaMemcpyAsync(c->h_ir, c->d_ir_l, 2 * irn * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (l) std::memcpy(l, c->h_ir, irn * sizeof(float));
    if (r) std::memcpy(r, c->h_ir + irn, irn * sizeof(float));
    return ARV2_OK;
}

int arv2_set_ir(arv2_ctx* c, const float* l, const float* r)
{
    REQUIRE(c && l && r, "null argument");
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->ir_len * sizeof(float);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(c->d_ir_l, l, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_ir_r, r, bytes, cudaMemcpyHostToDevice));
    return ARV2_OK;
}

int arv2_ir_device(arv2_ctx* c, float** l, float** r)
{
    REQUIRE(c, "null ctx");
    if (l) *l = c->d_ir_l;
    if (r) *r = c->d_ir_r;
    return ARV2_OK;
}

int arv2_hist_device(arv2_ctx* c, double** h, int64_t* count)
{
    REQUIRE(c, "null ctx");
    if (h) *h = c->d_hist;
    if (count) *count = 2LL * c->bands * c->ir_len;
    return ARV2_OK;
}

int arv2_last_upload_bytes(arv2_ctx* c, int64_t* bytes)
{
    REQUIRE(c && bytes, "null argument");
    *bytes = (int64_t)c->upload_bytes;
    return ARV2_OK;
}

int arv2_path_cache_info(arv2_ctx* c, int64_t* segments, int64_t* bytes)
{
    REQUIRE(c, "null ctx");
    if (!c->desc.path_cache || !c->cache_valid) { set_error("no valid path cache"); return ARV2_ERR_STATE; }
    if (segments) *segments = c->pc_segs;
    if (bytes) *bytes = c->pc_nvert * (int64_t)(2 * sizeof(float4) + sizeof(float) * c->bands + sizeof(uint2)) + (c->n_rays_total + 1) * (int64_t)sizeof(unsigned long long)
                        + (c->pc_nvert + 31) / 32 * (int64_t)sizeof(unsigned);
    return ARV2_OK;
}

int arv2_last_counters(arv2_ctx* c, uint64_t* out, int32_t n)
{
    REQUIRE(c && out && n >= 0, "bad argument");
    for (int i = 0; i < n; ++i) out[i] = i < kCounters ? (uint64_t)c->h_counters[i] : 0u;
    return ARV2_OK;
}

int arv2_last_segments(arv2_ctx* c, int64_t* segs)
{
    REQUIRE(c && segs, "null argument");
    *segs = c->last_segments;
    return ARV2_OK;
}

int arv2_get_records(arv2_ctx* c, int64_t capacity, int32_t* bin, int32_t* ear, float* energy, int32_t* nseg, int64_t* n_rays)
{
    REQUIRE(c && capacity >= 0, "arv2_get_records: bad argument");
    if (!c->desc.record_rays) { set_error("records not enabled (desc.record_rays)"); return ARV2_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (n_rays) *n_rays = c->last_range_rays;
    const size_t n = (size_t)std::min<long long>(capacity, std::min<long long>(c->last_range_rays, c->rec_capacity));
    if (n == 0) return ARV2_OK;
    if (bin) CK(cudaMemcpy(bin, c->d_rec_bin, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (ear) CK(cudaMemcpy(ear, c->d_rec_ear, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (nseg) CK(cudaMemcpy(nseg, c->d_rec_nseg, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (energy) CK(cudaMemcpy(energy, c->d_rec_energy, n * c->bands * sizeof(float), cudaMemcpyDeviceToHost));
    return ARV2_OK;
}

int arv2_write_ir_text(arv2_ctx* c, const char* left_path, const char* right_path)
{
    REQUIRE(c && left_path && right_path, "null argument");
    std::vector<float> l((size_t)c->bands * c->ir_len), r(l.size());
    const int rc = arv2_get_ir(c, l.data(), r.data());
    if (rc != ARV2_OK) return rc;
    std::ofstream fl(left_path), fr(right_path);
    if (!fl.is_open() && !fr.is_open()) { set_error("Error opening the file."); return ARV2_ERR_IO; }   // OR/AudioRenderer.cpp:544-547
    for (int i = 0; i < c->ir_len; ++i) { fl << l[i] << std::endl; fr << r[i] << std::endl; }         // :553-557
    return ARV2_OK;
}

int arv2_write_convolved_text(const char* left_path, const char* right_path, const float* l, const float* r, size_t n)
{
    REQUIRE(left_path && right_path && l && r, "null argument");
    std::ofstream fl(left_path), fr(right_path);
    if (!fl.is_open() && !fr.is_open()) { set_error("Error opening the file."); return ARV2_ERR_IO; }   // OR/AudioRenderer.cpp:722-727
    for (size_t i = 0; i < n; ++i) { fl << l[i] << std::endl; fr << r[i] << std::endl; }               // :733-737
    return ARV2_OK;
}

/* ------------------------------------------------------------ convolution -- */
int arv2_convolve_file(arv2_ctx* c, const float* x, size_t n, float* y_left, float* y_right, int32_t mode,
                       double* conv_ms, double* process_ms)
{
    REQUIRE(c && x && y_left && y_right, "arv2_convolve_file: null argument");
    REQUIRE(mode == ARV2_CONV_LINEAR || mode == ARV2_CONV_REFERENCE, "unknown convolution mode");
    const auto t0 = std::chrono::high_resolution_clock::now();
    CK(cudaSetDevice(c->device));
    if (n == 0) { if (conv_ms) *conv_ms = 0; if (process_ms) *process_ms = 0; return ARV2_OK; }
    const int block = 512, N = 2 * block;
    const int ir_len = c->ir_len;
    const int P = (ir_len + block - 1) / block;
    const long long fs = c->desc.sample_rate;
    long long seg_len, wrap, seg_out; int n_seg, bps; float gain;
    if (mode == ARV2_CONV_LINEAR) {
        seg_len = (long long)n; n_seg = 1; wrap = 0; seg_out = (long long)n; gain = 1.0f;
        bps = (int)((n + block - 1) / block);
    } else {
        // OR/kernels.cu:410  secondsToProcess = samples_len / sample_rate; FFT size == ir_len
        seg_len = fs; n_seg = (int)((long long)n / fs); wrap = ir_len; seg_out = ir_len;
        gain = (float)ir_len / (float)(ir_len / 2);          // OR/AudioRenderer.cpp:709
        bps = (int)((fs + block - 1) / block);
    }
    arv2_ctx::ConvWork& w = c->conv;
    if (!w.d_tw) {
        CK(cudaMalloc(&w.d_tw, N * sizeof(float2)));
        CK(conv_upload_twiddles(w.d_tw, N, c->stream));
    }
    const size_t need_x = n, need_X = (size_t)std::max(1, n_seg * bps) * block, need_H = (size_t)2 * P * block;
    if (w.cap_x < need_x) {
        cudaFree(w.d_x); cudaFree(w.d_out);
        w.d_x = nullptr; w.d_out = nullptr; w.cap_x = 0;
        CK(cudaMalloc(&w.d_x, need_x * sizeof(float)));
        CK(cudaMalloc(&w.d_out, 2 * need_x * sizeof(float)));
        w.cap_x = need_x;
    }
    if (w.cap_X < need_X) { cudaFree(w.d_X); w.d_X = nullptr; w.cap_X = 0; CK(cudaMalloc(&w.d_X, need_X * sizeof(float2))); w.cap_X = need_X; }
    if (w.cap_H < need_H) { cudaFree(w.d_H); w.d_H = nullptr; w.cap_H = 0; CK(cudaMalloc(&w.d_H, need_H * sizeof(float2))); w.cap_H = need_H; }

    CK(cudaMemcpyAsync(w.d_x, x, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));      // OR/AudioRenderer.cpp:673
    CK(cudaMemsetAsync(w.d_out, 0, 2 * n * sizeof(float), c->stream));                        // :682-683
    CK(cudaEventRecord(c->ev0, c->stream));
    // IR spectra of both ears (band 0): items = {ir_left, ir_right}
    CK(conv_ir_spectra(c->d_ir_l, 1, 0, ir_len, block, P, w.d_tw, w.d_H, c->stream));
    CK(conv_ir_spectra(c->d_ir_r, 1, 1, ir_len, block, P, w.d_tw, w.d_H, c->stream));
    if (n_seg > 0) {
        CK(conv_block_spectra(w.d_x, (long long)n, seg_len, n_seg, bps, block, w.d_tw, w.d_X, c->stream));
        ConvFileArgs a{};
        a.X = w.d_X; a.H = w.d_H; a.out_l = w.d_out; a.out_r = w.d_out + n; a.tw = w.d_tw;
        a.n = (long long)n; a.seg_len = seg_len; a.wrap = wrap; a.seg_out = seg_out;
        a.n_seg = n_seg; a.blocks_per_seg = bps; a.block = block; a.P = P; a.gain = gain;
        CK(conv_file(a, c->stream));
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaMemcpyAsync(y_left, w.d_out, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));   // :702-703
    CK(cudaMemcpyAsync(y_right, w.d_out + n, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (conv_ms) { float t = 0.f; CK(cudaEventElapsedTime(&t, c->ev0, c->ev1)); *conv_ms = t; }
    if (process_ms) *process_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    return ARV2_OK;
}

int arv2_stream_open(int32_t device, int32_t n_sources, int32_t block, int32_t ir_length, arv2_stream** out)
{
    REQUIRE(out && n_sources > 0 && ir_length > 0, "arv2_stream_open: bad argument");
    REQUIRE(block >= 64 && block <= 1024 && (block & (block - 1)) == 0, "block must be a power of two in [64,1024]");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
    CK(cudaSetDevice(device));
    auto* s = new arv2_stream;
    s->device = device; s->n_src = n_sources; s->block = block; s->ir_len = ir_length;
    s->P = (ir_length + block - 1) / block;
    {
        cudaDeviceProp prop;
        const int sms = cudaGetDeviceProperties(&prop, device) == cudaSuccess ? prop.multiProcessorCount : 0;
        // experiment (r09, profiles/r09_conv.md): a 16-stage ring when a step and its successor find SMs of their own
        // (2 x sources x 8 CTAs <= SMs).  Measured: 6.42 against 6.53 us per block -- the ring depth is not what bounds a step.
        bool deep = false;
        if (const char* e = getenv("ARV2_CONV_DEEP_RING")) deep = atoi(e) != 0 && (atoi(e) > 1 || 2 * n_sources * 8 <= sms);
        s->stages = conv_ring_stages(block, deep);
        // 16 CTAs per source when there are few sources: every rank then streams at most 12 old partitions before the wait,
        // which is what a programmatic dependent gets through before its predecessor has completed (profiles/r09_conv.md).
        // The summation order over partitions depends on the cluster size, so it is fixed per stream.
        int wide_max = 4;
        if (const char* e = getenv("ARV2_CONV_CLUSTER16")) wide_max = atoi(e);        // A/B: most sources of a stream that runs 16-CTA clusters (0 = never)
        const bool wide = n_sources <= wide_max && !getenv("ARV2_CONV_PERSISTENT");
        s->cluster = (wide && conv_cluster16_ok(n_sources, block)) ? 16 : 8;
    }
    const size_t spec = (size_t)s->P * block;          // float2 per (source) FDL or per ear
    const size_t nin = (size_t)n_sources * block;
#define CKS(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(e_)); arv2_stream_close(s); return ARV2_ERR_CUDA; } } while (0)
    CKS(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CKS(cudaEventCreateWithFlags(&s->ev_step, cudaEventDisableTiming));
    CKS(cudaEventCreateWithFlags(&s->ev_swap, cudaEventDisableTiming));
    CKS(cudaEventCreateWithFlags(&s->ev_ir_copied, cudaEventDisableTiming));
    CKS(cudaMalloc(&s->d_tw, 2 * block * sizeof(float2)));
    CKS(conv_upload_twiddles(s->d_tw, 2 * block, s->stream));
    CKS(cudaMalloc(&s->d_fdl, (size_t)n_sources * spec * sizeof(float2)));
    for (int b = 0; b < 2; ++b) {
        CKS(cudaMalloc(&s->d_H[b], (size_t)n_sources * 2 * spec * sizeof(float2)));
        CKS(cudaMemsetAsync(s->d_H[b], 0, (size_t)n_sources * 2 * spec * sizeof(float2), s->stream));
    }
    CKS(cudaMalloc(&s->d_Hptr, n_sources * sizeof(float2*)));
    CKS(cudaMallocHost(&s->h_Hptr, 2 * n_sources * sizeof(float2*)));
    s->active.assign(n_sources, 0);
    for (int i = 0; i < n_sources; ++i)
        for (int b = 0; b < 2; ++b) s->h_Hptr[2 * i + b] = s->d_H[b] + (size_t)i * 2 * spec;
    for (int i = 0; i < n_sources; ++i)
        CKS(cudaMemcpyAsync(s->d_Hptr + i, s->h_Hptr + 2 * i, sizeof(float2*), cudaMemcpyHostToDevice, s->stream));
    CKS(cudaMalloc(&s->d_tail, 2 * nin * sizeof(float)));
    CKS(cudaMalloc(&s->d_in, (size_t)kHostBlocks * nin * sizeof(float)));
    CKS(cudaMalloc(&s->d_out, (size_t)kHostBlocks * 2 * nin * sizeof(float)));
    CKS(cudaMalloc(&s->d_mix, (size_t)kHostBlocks * 2 * block * sizeof(float)));
    CKS(cudaMalloc(&s->d_done, sizeof(unsigned)));
    CKS(cudaMemsetAsync(s->d_done, 0, sizeof(unsigned), s->stream));
    CKS(cudaMalloc(&s->d_ir, (size_t)2 * ir_length * sizeof(float)));
    CKS(cudaHostAlloc(&s->h_in, (size_t)kHostBlocks * nin * sizeof(float), cudaHostAllocMapped));
    CKS(cudaHostAlloc(&s->h_out, (size_t)kHostBlocks * 2 * nin * sizeof(float), cudaHostAllocMapped));
    CKS(cudaHostAlloc(&s->h_mix, (size_t)kHostBlocks * 2 * block * sizeof(float), cudaHostAllocMapped));
    CKS(cudaHostAlloc(&s->h_flag, sizeof(unsigned), cudaHostAllocMapped));
    *s->h_flag = 0u;
    CKS(cudaMallocHost(&s->h_ir, (size_t)2 * ir_length * sizeof(float)));
#undef CKS
    *out = s;
    return arv2_stream_reset(s);
}

int arv2_stream_reset(arv2_stream* s)
{
    REQUIRE(s, "null stream");
    CK(cudaSetDevice(s->device));
    if (s->step_pending) { CK(cudaStreamWaitEvent(s->stream, s->ev_step, 0)); s->step_pending = false; }
    CK(cudaMemsetAsync(s->d_fdl, 0, (size_t)s->n_src * s->P * s->block * sizeof(float2), s->stream));
    CK(cudaMemsetAsync(s->d_tail, 0, (size_t)s->n_src * 2 * s->block * sizeof(float), s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->slot = s->P - 1;
    return ARV2_OK;
}

// New spectra go into the source's inactive buffer and the pointer table flips, all in order on the convolver's own
// stream; steps may run on another stream (the caller's), so the two are tied with events: the swap waits for the last
// step enqueued anywhere (a step still in flight may be reading the buffer that was active two swaps ago, or the
// pointer table), and the next step waits for the swap.  Nothing blocks the host.
static int stream_swap_ir(arv2_stream* s, int32_t source)
{
    const size_t spec = (size_t)s->P * s->block;
    const int nb = 1 - s->active[source];
    float2* H = s->d_H[nb] + (size_t)source * 2 * spec;
    if (s->step_pending) { CK(cudaStreamWaitEvent(s->stream, s->ev_step, 0)); s->step_pending = false; }
    CK(conv_ir_spectra(s->d_ir, 2, 0, s->ir_len, s->block, s->P, s->d_tw, H, s->stream));
    s->active[source] = nb;
    CK(cudaMemcpyAsync(s->d_Hptr + source, s->h_Hptr + 2 * source + nb, sizeof(float2*), cudaMemcpyHostToDevice, s->stream));
    CK(cudaEventRecord(s->ev_swap, s->stream));
    s->swap_pending = true;
    return ARV2_OK;
}

int arv2_stream_set_ir(arv2_stream* s, int32_t source, const float* l, const float* r)
{
    REQUIRE(s && l && r && source >= 0 && source < s->n_src, "arv2_stream_set_ir: bad argument");
    CK(cudaSetDevice(s->device));
    if (s->ir_copy_pending) { CK(cudaEventSynchronize(s->ev_ir_copied)); s->ir_copy_pending = false; }     // the staging buffer is free again
    std::memcpy(s->h_ir, l, (size_t)s->ir_len * sizeof(float));
    std::memcpy(s->h_ir + s->ir_len, r, (size_t)s->ir_len * sizeof(float));
    if (s->step_pending) { CK(cudaStreamWaitEvent(s->stream, s->ev_step, 0)); s->step_pending = false; }
    CK(cudaMemcpyAsync(s->d_ir, s->h_ir, (size_t)2 * s->ir_len * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    CK(cudaEventRecord(s->ev_ir_copied, s->stream));
    s->ir_copy_pending = true;
    return stream_swap_ir(s, source);
}

int arv2_stream_set_ir_device(arv2_stream* s, int32_t source, const float* dl, const float* dr)
{
    REQUIRE(s && dl && dr && source >= 0 && source < s->n_src, "arv2_stream_set_ir_device: bad argument");
    CK(cudaSetDevice(s->device));
    if (s->step_pending) { CK(cudaStreamWaitEvent(s->stream, s->ev_step, 0)); s->step_pending = false; }
    CK(cudaMemcpyAsync(s->d_ir, dl, (size_t)s->ir_len * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaMemcpyAsync(s->d_ir + s->ir_len, dr, (size_t)s->ir_len * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
    return stream_swap_ir(s, source);
}

int arv2_stream_set_gains(arv2_stream* s, const float* gains)
{
    REQUIRE(s, "null stream");
    CK(cudaSetDevice(s->device));
    if (!gains) { cudaFree(s->d_gain); s->d_gain = nullptr; return ARV2_OK; }
    if (!s->d_gain) CK(cudaMalloc(&s->d_gain, (size_t)s->n_src * sizeof(float)));
    CK(cudaMemcpy(s->d_gain, gains, (size_t)s->n_src * sizeof(float), cudaMemcpyHostToDevice));
    return ARV2_OK;
}

// n_blocks steps on stream `st` (ours or the caller's), ordered after any pending IR swap; leaves the event the next
// swap waits on
static int enqueue_steps(arv2_stream* s, const float* d_in, float* d_out, int32_t n_blocks, cudaStream_t st)
{
    CK(cudaSetDevice(s->device));
    if (s->swap_pending) { if (st != s->stream) CK(cudaStreamWaitEvent(st, s->ev_swap, 0)); s->swap_pending = false; }
    const size_t nin = (size_t)s->n_src * s->block;
    // Experiment (ARV2_CONV_PERSISTENT=1, r08): the blocks of one call in ONE cluster launch, every source's cluster looping
    // over them.  Bit-identical, but slower: 8.6 us per block against 6.5 (2 sources) and 11.0 against 8.1 (16 sources) --
    // with one launch per block and programmatic dependent launch, step k+1's accumulation runs in a second CTA on the
    // same SM while step k's rank 0 is in its FFTs; a looping cluster serialises them (profiles/r08_conv.md).
    const bool persistent = getenv("ARV2_CONV_PERSISTENT") != nullptr;
    // blocks 1.. of a call run their forward FFT before waiting for their predecessor (their input was complete before
    // block 0 passed its wait).  Bit-identical; 6.00 against 6.28 us per block with 2 sources (r09; ARV2_CONV_LATE_FFT=1: A/B)
    const bool late_fft = getenv("ARV2_CONV_LATE_FFT") != nullptr;
    if (n_blocks > 1 && persistent) {
        ConvStreamArgs a{};
        a.in = d_in; a.out = d_out;
        a.fdl = s->d_fdl; a.H = (const float2* const*)s->d_Hptr; a.tail = s->d_tail; a.tw = s->d_tw;
        a.n_src = s->n_src; a.block = s->block; a.P = s->P; a.slot = (s->slot + 1) % s->P; a.n_blocks = n_blocks;
        CK(conv_stream_blocks(a, st));
        s->slot = (s->slot + n_blocks) % s->P;
    } else {
        for (int32_t b = 0; b < n_blocks; ++b) {
            s->slot = (s->slot + 1) % s->P;
            ConvStreamArgs a{};
            a.in = d_in + (size_t)b * nin; a.out = d_out + (size_t)b * 2 * nin;
            a.fdl = s->d_fdl; a.H = (const float2* const*)s->d_Hptr; a.tail = s->d_tail; a.tw = s->d_tw;
            a.n_src = s->n_src; a.block = s->block; a.P = s->P; a.slot = s->slot; a.n_blocks = 1; a.stages = s->stages; a.cluster = s->cluster;
            // the input of blocks 1.. of a call was complete before block 0 passed its wait: their forward FFT may run early
            a.early_input = (b > 0 && !late_fft) ? 1 : 0;
            CK(conv_stream_step(a, st));
        }
    }
    if (st != s->stream && n_blocks > 0) { CK(cudaEventRecord(s->ev_step, st)); s->step_pending = true; }
    return ARV2_OK;
}

int arv2_stream_process_device(arv2_stream* s, const float* d_in, float* d_out, void* cuda_stream)
{
    REQUIRE(s && d_in && d_out, "arv2_stream_process_device: null argument");
    return enqueue_steps(s, d_in, d_out, 1, cuda_stream ? (cudaStream_t)cuda_stream : s->stream);
}

int arv2_stream_process_device_blocks(arv2_stream* s, const float* d_in, float* d_out, int32_t n_blocks, void* cuda_stream)
{
    REQUIRE(s && d_in && d_out && n_blocks >= 0, "arv2_stream_process_device_blocks: bad argument");
    return enqueue_steps(s, d_in, d_out, n_blocks, cuda_stream ? (cudaStream_t)cuda_stream : s->stream);
}

int arv2_stream_mix_device(arv2_stream* s, const float* d_out, float* d_mix, int32_t n_blocks, void* cuda_stream)
{
    REQUIRE(s && d_out && d_mix && n_blocks >= 0, "arv2_stream_mix_device: bad argument");
    CK(cudaSetDevice(s->device));
    CK(conv_mix(d_out, s->n_src, s->block, n_blocks, s->d_gain, d_mix, cuda_stream ? (cudaStream_t)cuda_stream : s->stream));
    return ARV2_OK;
}

// Host buffers in, host buffers out, n_blocks <= kHostBlocks: the steps read the input from mapped pinned memory and the
// result lands in mapped pinned memory; one completion word instead of copies + a stream synchronisation.
static int process_host(arv2_stream* s, const float* in, float* out, float* mix, int32_t n_blocks)
{
    REQUIRE(n_blocks >= 1 && n_blocks <= kHostBlocks, "at most 16 blocks per host-buffer call");
    CK(cudaSetDevice(s->device));
    const size_t nin = (size_t)s->n_src * s->block;
    std::memcpy(s->h_in, in, (size_t)n_blocks * nin * sizeof(float));
    const unsigned seq = ++s->flag_seq;
    static const bool zero_copy = getenv("ARV2_CONV_ZEROCOPY") != nullptr;      // A/B: the step kernels read / write the mapped buffers themselves (slower, r08)
    if (zero_copy) {
        float* step_out = out ? s->h_out : s->d_out;
        const int rc = enqueue_steps(s, s->h_in, step_out, n_blocks, s->stream);
        if (rc != ARV2_OK) return rc;
        if (mix) CK(conv_mix(step_out, s->n_src, s->block, n_blocks, s->d_gain, s->h_mix, s->stream));
        CK(conv_signal(s->h_flag, seq, s->stream));
    } else {
        // staging by one-CTA kernels over the mapped buffers: in -> device, steps (+ mix) on device buffers, results + completion word -> host
        CK(conv_stage(s->h_in, s->d_in, (long long)n_blocks * (long long)nin, nullptr, nullptr, 0, nullptr, 0u, s->d_done, s->stream));
        const int rc = enqueue_steps(s, s->d_in, s->d_out, n_blocks, s->stream);
        if (rc != ARV2_OK) return rc;
        if (mix) CK(conv_mix(s->d_out, s->n_src, s->block, n_blocks, s->d_gain, s->d_mix, s->stream));
        CK(conv_stage(out ? s->d_out : nullptr, s->h_out, out ? (long long)n_blocks * 2 * (long long)nin : 0,
                      mix ? s->d_mix : nullptr, s->h_mix, mix ? (long long)n_blocks * 2 * s->block : 0, s->h_flag, seq, s->d_done, s->stream));
    }
    volatile unsigned* flag = s->h_flag;
    for (unsigned long long spins = 0; *flag != seq; ++spins) {
        if ((spins & 0xfffff) == 0xfffff) {                              // every ~1M polls: is the device still alive?
            const cudaError_t e = cudaStreamQuery(s->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) { set_error(std::string("stream step: ") + cudaGetErrorString(e)); return ARV2_ERR_CUDA; }
        }
    }
    if (out) std::memcpy(out, s->h_out, (size_t)n_blocks * 2 * nin * sizeof(float));
    if (mix) std::memcpy(mix, s->h_mix, (size_t)n_blocks * 2 * s->block * sizeof(float));
    return ARV2_OK;
}

int arv2_stream_process(arv2_stream* s, const float* in, float* out)
{
    REQUIRE(s && in && out, "arv2_stream_process: null argument");
    return process_host(s, in, out, nullptr, 1);
}

int arv2_stream_process_blocks(arv2_stream* s, const float* in, float* out, float* mix, int32_t n_blocks)
{
    REQUIRE(s && in && (out || mix), "arv2_stream_process_blocks: null argument");
    return process_host(s, in, out, mix, n_blocks);
}

void arv2_stream_close(arv2_stream* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->d_tw); cudaFree(s->d_fdl); cudaFree(s->d_H[0]); cudaFree(s->d_H[1]); cudaFree(s->d_Hptr);
    cudaFree(s->d_tail); cudaFree(s->d_in); cudaFree(s->d_out); cudaFree(s->d_mix); cudaFree(s->d_done); cudaFree(s->d_ir); cudaFree(s->d_gain);
    if (s->h_Hptr) cudaFreeHost(s->h_Hptr);
    if (s->h_in) cudaFreeHost(s->h_in);
    if (s->h_out) cudaFreeHost(s->h_out);
    if (s->h_mix) cudaFreeHost(s->h_mix);
    if (s->h_flag) cudaFreeHost(s->h_flag);
    if (s->h_ir) cudaFreeHost(s->h_ir);
    if (s->ev_step) cudaEventDestroy(s->ev_step);
    if (s->ev_swap) cudaEventDestroy(s->ev_swap);
    if (s->ev_ir_copied) cudaEventDestroy(s->ev_ir_copied);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

// audioHandlerWithMic + convoluteLiveInput (OR/main.cpp:99-135, OR/AudioRenderer.cpp:593-661)
int arv2_live_callback(arv2_stream* s, const double* in, size_t n_in, arv2_ring* ring)
{
    REQUIRE(s && in && ring, "arv2_live_callback: null argument");
    REQUIRE(s->n_src == 1, "arv2_live_callback needs a 1-source stream");
    REQUIRE(n_in % (size_t)s->block == 0, "n_in must be a multiple of the stream block");
    const size_t B = (size_t)s->block;
    std::vector<float> x(std::min<size_t>(n_in, (size_t)kHostBlocks * B)), y(2 * x.size());
    std::vector<double> zipped(2 * n_in);
    for (size_t done = 0; done < n_in; done += x.size()) {
        const size_t nb = std::min<size_t>((n_in - done) / B, (size_t)kHostBlocks);     // one call carries up to 16 blocks
        for (size_t i = 0; i < nb * B; ++i) x[i] = (float)in[done + i];
        const int rc = process_host(s, x.data(), y.data(), nullptr, (int32_t)nb);
        if (rc != ARV2_OK) return rc;
        for (size_t b = 0; b < nb; ++b)
            for (size_t i = 0; i < B; ++i) {                     // gain 1/(ir_len/2) of the reference = 2x, then d_zipArrays
                zipped[2 * (done + b * B + i)] = 2.0 * (double)y[b * 2 * B + i];
                zipped[2 * (done + b * B + i) + 1] = 2.0 * (double)y[b * 2 * B + B + i];
            }
    }
    return arv2_ring_add(ring, zipped.data(), zipped.size());
}

/* ------------------------------------------------------------------ audio -- */
int arv2_wav_read(const char* path, float** samples, size_t* n, int32_t* sample_rate, int32_t* channels)
{
    REQUIRE(path && samples && n && sample_rate && channels, "arv2_wav_read: null argument");
    std::string err;
    const int rc = wav_read(path, samples, n, sample_rate, channels, &err);
    if (rc != ARV2_OK) set_error(err);
    return rc;
}

int arv2_wav_write_stereo_normalized(const char* path, const float* l, const float* r, size_t n, int32_t rate)
{
    REQUIRE(path && l && r, "arv2_wav_write_stereo_normalized: null argument");
    std::string err;
    const int rc = wav_write_stereo_normalized(path, l, r, n, rate, &err);
    if (rc != ARV2_OK) set_error(err);
    return rc;
}

void arv2_free(void* p) { std::free(p); }

/* not in include/arv2.h: tuning builds only (-DARV2_CONV_TIMING -DARV2_CONV_TRACE) */
extern "C" int arv2_debug_conv_trace(void* out, size_t bytes) { return arv2::conv_debug_trace(out, bytes) == cudaSuccess ? 0 : -1; }

} // extern "C"

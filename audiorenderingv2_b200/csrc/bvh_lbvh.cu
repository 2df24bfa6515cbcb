// bvh_lbvh.cu -- GPU LBVH builder (K1): Morton codes -> radix sort -> Karras radix tree ->
// bottom-up refit -> collapse of <=4-triangle subtrees into leaves -> emission in the
// binary layout of arv2_internal.h (BvhNode); the host then quantises it like the SAH tree.
// All hand-written sm_100a kernels, no CUB/Thrust.
//
// Replaces optixAccelBuild (OR/AudioRenderer.cpp:95-218) when a fast rebuild matters more
// than tree quality (desc.bvh_builder = 1); the host binned-SAH builder stays the default
// for static scenes.  Either tree yields the same hits (closest = min (t, id) over exact
// triangle tests); only the number of node visits differs.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "arv2_internal.h"
#include "bvh_lbvh.cuh"

namespace arv2 {

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                       // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems; // keys per block
constexpr int kRadixBits = 6;
constexpr int kRadix = 1 << kRadixBits;

__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void bounds_kernel(const float* __restrict__ verts, int n_tris, int* __restrict__ bounds /*[6] ordered ints*/)
{
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < 3LL * n_tris; v += (long long)gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float x = verts[3 * v + a]; lo[a] = fminf(lo[a], x); hi[a] = fmaxf(hi[a], x); }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(bounds + a, f2ord(lo[a])); atomicMax(bounds + 3 + a, f2ord(hi[a])); }
    }
}

__device__ __forceinline__ unsigned expand10(unsigned v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void morton_kernel(const float* __restrict__ verts, int n, const int* __restrict__ bounds, unsigned* __restrict__ keys,
                              int* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(bounds[a]), hi = ord2f(bounds[3 + a]);
        const float* t = verts + 9 * (size_t)i;
        const float mn = fminf(fminf(t[a], t[3 + a]), t[6 + a]), mx = fmaxf(fmaxf(t[a], t[3 + a]), t[6 + a]);
        const float c = 0.5f * mn + 0.5f * mx;
        const float ext = hi - lo;
        float u = ext > 0.f ? (c - lo) / ext : 0.f;
        u = fminf(fmaxf(u * 1024.f, 0.f), 1023.f);
        q[a] = (unsigned)u;
    }
    keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[i] = i;
}

// ---- stable LSD radix sort, 6 bits per pass -------------------------------------------
// Blocked arrangement: thread t owns keys [tile + t*16, tile + (t+1)*16), so a per-thread
// per-digit count followed by a scan across threads gives a stable rank without atomics.
__device__ __forceinline__ void tile_counts(const unsigned* __restrict__ keys, int n, int shift, unsigned short (*cnt)[kSortThreads])
{
    const int t = threadIdx.x;
    for (int d = 0; d < kRadix; ++d) cnt[d][t] = 0;
    const long long base = (long long)blockIdx.x * kSortTile + (long long)t * kSortItems;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k)
        if (base + k < n) cnt[(keys[base + k] >> shift) & (kRadix - 1)][t]++;
    __syncthreads();
}

__global__ void __launch_bounds__(kSortThreads) radix_count_kernel(const unsigned* __restrict__ keys, int n, int shift,
                                                                    unsigned* __restrict__ counts /*[kRadix][gridDim.x]*/)
{
    __shared__ unsigned short cnt[kRadix][kSortThreads];
    tile_counts(keys, n, shift, cnt);
    if (threadIdx.x < kRadix) {
        unsigned s = 0;
        for (int t = 0; t < kSortThreads; ++t) s += cnt[threadIdx.x][t];
        counts[threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
}

// exclusive scan of `m` unsigned values by one block (m up to a few hundred thousand)
__global__ void __launch_bounds__(1024) scan_single_block_kernel(unsigned* __restrict__ data, int m)
{
    __shared__ unsigned part[1024];
    const int t = threadIdx.x;
    const int per = (m + 1023) / 1024;
    const int b = t * per, e = min(m, b + per);
    unsigned s = 0;
    for (int i = b; i < e; ++i) s += data[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned v = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned run = part[t] - s;
    for (int i = b; i < e; ++i) { const unsigned v = data[i]; data[i] = run; run += v; }
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const unsigned* __restrict__ keys, const int* __restrict__ vals,
                                                                      int n, int shift, const unsigned* __restrict__ offsets,
                                                                      unsigned* __restrict__ keys_out, int* __restrict__ vals_out)
{
    __shared__ unsigned short cnt[kRadix][kSortThreads];
    tile_counts(keys, n, shift, cnt);
    if (threadIdx.x < kRadix) {          // exclusive scan across threads, one digit per thread
        unsigned short run = 0;
        for (int t = 0; t < kSortThreads; ++t) { const unsigned short v = cnt[threadIdx.x][t]; cnt[threadIdx.x][t] = run; run += v; }
    }
    __syncthreads();
    const int t = threadIdx.x;
    const long long base = (long long)blockIdx.x * kSortTile + (long long)t * kSortItems;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        if (base + k < n) {
            const unsigned key = keys[base + k];
            const int d = (key >> shift) & (kRadix - 1);
            const unsigned pos = offsets[d * gridDim.x + blockIdx.x] + cnt[d][t]++;
            keys_out[pos] = key;
            vals_out[pos] = vals[base + k];
        }
    }
}

} // namespace

// Stable LSD radix sort of (key, value) pairs on bits [lo_bit, hi_bit) of the keys, kRadixBits
// per pass, with the count / scan / scatter kernels above.  Returns 0 or 1: which of the two
// ping-pong buffers holds the result.
size_t radix_sort_scratch_bytes(int n) { return n <= 0 ? 0 : (size_t)kRadix * (size_t)((n + kSortTile - 1) / kSortTile) * 4; }

cudaError_t radix_sort_pairs(unsigned* keys[2], int* vals[2], int n, int lo_bit, int hi_bit, int* result, cudaStream_t stream, unsigned* scratch)
{
    *result = 0;
    if (n <= 0) return cudaSuccess;
    const int sort_blocks = (n + kSortTile - 1) / kSortTile;
    unsigned* counts = scratch;
    cudaError_t e = cudaSuccess;
    if (!counts) e = cudaMalloc(&counts, (size_t)kRadix * sort_blocks * 4);
    if (e != cudaSuccess) return e;
    int cur = 0;
    for (int shift = lo_bit; shift < hi_bit; shift += kRadixBits) {
        radix_count_kernel<<<sort_blocks, kSortThreads, 0, stream>>>(keys[cur], n, shift, counts);
        scan_single_block_kernel<<<1, 1024, 0, stream>>>(counts, kRadix * sort_blocks);
        radix_scatter_kernel<<<sort_blocks, kSortThreads, 0, stream>>>(keys[cur], vals[cur], n, shift, counts, keys[cur ^ 1], vals[cur ^ 1]);
        cur ^= 1;
    }
    e = cudaGetLastError();
    *result = cur;
    if (scratch) return e;               // the caller's scratch: nothing to free, nothing to wait for
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(counts);
    return e;
}

namespace {

// ---- Karras 2012 radix tree over unique 64-bit keys (morton << 32 | sorted position) ---
__device__ __forceinline__ int delta(const unsigned* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = ((unsigned long long)keys[i] << 32) | (unsigned)i;
    const unsigned long long b = ((unsigned long long)keys[j] << 32) | (unsigned)j;
    return __clzll((long long)(a ^ b));
}

constexpr int kLeafFlag = 0x40000000;     // child reference to sorted leaf position

__global__ void hierarchy_kernel(const unsigned* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                                 int* __restrict__ first, int* __restrict__ last, int* __restrict__ parent /*[2n-1]: internal i, leaf n-1+s*/)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = (lo == gamma) ? (gamma | kLeafFlag) : gamma;
    const int rc = (hi == gamma + 1) ? ((gamma + 1) | kLeafFlag) : gamma + 1;
    left[i] = lc; right[i] = rc; first[i] = lo; last[i] = hi;
    parent[(lc & kLeafFlag) ? (n - 1 + (lc & ~kLeafFlag)) : lc] = i;
    parent[(rc & kLeafFlag) ? (n - 1 + (rc & ~kLeafFlag)) : rc] = i;
    if (i == 0) parent[0] = -1;
}

// boxes: [2n-1][6] (lo xyz, hi xyz); internal i at i, leaf s at n-1+s
__global__ void refit_kernel(const float* __restrict__ verts, const int* __restrict__ vals, int n, const int* __restrict__ left,
                             const int* __restrict__ right, const int* __restrict__ parent, float* __restrict__ boxes,
                             int* __restrict__ visits)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float* t = verts + 9 * (size_t)vals[s];
    float b[6];
#pragma unroll
    for (int a = 0; a < 3; ++a) { b[a] = fminf(fminf(t[a], t[3 + a]), t[6 + a]); b[3 + a] = fmaxf(fmaxf(t[a], t[3 + a]), t[6 + a]); }
    float* dst = boxes + 6 * (size_t)(n - 1 + s);
#pragma unroll
    for (int a = 0; a < 6; ++a) dst[a] = b[a];
    int node = parent[n - 1 + s];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(visits + node, 1) == 0) return;      // the sibling subtree is not ready yet
        const int lc = left[node], rc = right[node];
        const float* bl = boxes + 6 * (size_t)((lc & kLeafFlag) ? (n - 1 + (lc & ~kLeafFlag)) : lc);
        const float* br = boxes + 6 * (size_t)((rc & kLeafFlag) ? (n - 1 + (rc & ~kLeafFlag)) : rc);
        float* bd = boxes + 6 * (size_t)node;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            bd[a] = fminf(__ldcg(bl + a), __ldcg(br + a));
            bd[3 + a] = fmaxf(__ldcg(bl + 3 + a), __ldcg(br + 3 + a));
        }
        node = parent[node];
    }
}

// Collapse of the radix tree into leaves of <= 4 triangles, with the SAH leaf test of the host builder
// (host/bvh_sah.cpp): a subtree of <= 4 triangles stays an inner node only if splitting it as the radix tree does
// is cheaper than testing its triangles, a node visit priced at one triangle test:
//     1 + (A_left * N_left + A_right * N_right) / A  <  N.
// want[i] = node i passes its own test (always, above 4 triangles).
__global__ void emit_want_kernel(const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ first,
                                 const int* __restrict__ last, const float* __restrict__ boxes, int n, unsigned char* __restrict__ want)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int cnt = last[i] - first[i] + 1;
    if (cnt > kMaxLeafTris) { want[i] = 1; return; }
    auto half_area = [&](const float* b) {
        const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
        return dx * dy + dy * dz + dz * dx;
    };
    const int ch[2] = {left[i], right[i]};
    float cost = 0.f;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const int c = ch[w];
        if (c & kLeafFlag) cost += half_area(boxes + 6 * (size_t)(n - 1 + (c & ~kLeafFlag)));
        else cost += half_area(boxes + 6 * (size_t)c) * (float)(last[c] - first[c] + 1);
    }
    const float a = fmaxf(half_area(boxes + 6 * (size_t)i), 1e-30f);
    want[i] = (1.0f + cost / a < (float)cnt) ? 1 : 0;
}

// flags[i] = 1 when internal node i survives the collapse: it passes its test and so does every ancestor inside
// its <= 4-triangle subtree (at most two of them)
__global__ void emit_flag_kernel(const int* __restrict__ first, const int* __restrict__ last, const int* __restrict__ parent,
                                 const unsigned char* __restrict__ want, int n, unsigned* __restrict__ flags, unsigned char* __restrict__ keep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    bool k = want[i] != 0;
    int p = parent[i];
    while (k && p >= 0 && last[p] - first[p] + 1 <= kMaxLeafTris) { k = want[p] != 0; p = parent[p]; }
    keep[i] = k ? 1 : 0;
    flags[i] = k ? 1u : 0u;
}

__global__ void emit_kernel(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ first,
                            const int* __restrict__ last, const float* __restrict__ boxes, const unsigned* __restrict__ new_index,
                            const unsigned char* __restrict__ keep, int node_offset, int slot_offset, float pad, float4* __restrict__ nodes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    if (!keep[i]) return;
    int code[2];
    float bx[2][6];
    const int ch[2] = {left[i], right[i]};
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const int c = ch[w];
        const float* b;
        if (c & kLeafFlag) {
            const int s = c & ~kLeafFlag;
            code[w] = ~(((s + slot_offset) << kLeafShift) | 0);
            b = boxes + 6 * (size_t)(n - 1 + s);
        } else {
            const int cnt = last[c] - first[c] + 1;
            code[w] = keep[c] ? (int)new_index[c] + node_offset : ~(((first[c] + slot_offset) << kLeafShift) | (cnt - 1));
            b = boxes + 6 * (size_t)c;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) { bx[w][a] = b[a] - pad; bx[w][3 + a] = b[3 + a] + pad; }
    }
    float4* d = nodes + 4 * (size_t)new_index[i];
    d[0] = make_float4(bx[0][0], bx[0][3], bx[0][1], bx[0][4]);
    d[1] = make_float4(bx[1][0], bx[1][3], bx[1][1], bx[1][4]);
    d[2] = make_float4(bx[0][2], bx[0][5], bx[1][2], bx[1][5]);
    d[3] = make_float4(__int_as_float(code[0]), __int_as_float(code[1]), 0.f, 0.f);
}

__global__ void tri_gather_kernel(const float* __restrict__ verts, const int* __restrict__ mats, const int* __restrict__ vals, int n,
                                  int id_base, int slot_offset, float4* __restrict__ tris)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int src = vals[s];
    const float* v = verts + 9 * (size_t)src;
    float4* d = tris + 3 * (size_t)(s + slot_offset);
    d[0] = make_float4(v[0], v[1], v[2], __int_as_float(id_base + src));
    d[1] = make_float4(v[3], v[4], v[5], __int_as_float(mats[src]));
    d[2] = make_float4(v[6], v[7], v[8], 0.f);
}

} // namespace

#define LB(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cleanup(); return e_; } } while (0)

cudaError_t build_bvh_lbvh(const float* d_verts, const int* d_mats, int n, int id_base, float4* d_nodes, float4* d_tris,
                           int* d_order, int node_offset, int slot_offset, LbvhResult* out, cudaStream_t stream)
{
    if (n <= kMaxLeafTris) return cudaErrorInvalidValue;    // tiny inputs take the host builder
    unsigned *keys[2] = {nullptr, nullptr}, *counts = nullptr, *flags = nullptr;
    unsigned char *want = nullptr, *keep = nullptr;
    int *vals[2] = {nullptr, nullptr}, *left = nullptr, *right = nullptr, *first = nullptr, *last = nullptr, *parent = nullptr, *visits = nullptr, *bounds = nullptr;
    float* boxes = nullptr;
    auto cleanup = [&]() {
        cudaFree(keys[0]); cudaFree(keys[1]); cudaFree(vals[0]); cudaFree(vals[1]); cudaFree(counts); cudaFree(flags);
        cudaFree(left); cudaFree(right); cudaFree(first); cudaFree(last); cudaFree(parent); cudaFree(visits); cudaFree(bounds); cudaFree(boxes);
        cudaFree(want); cudaFree(keep);
    };
    const int sort_blocks = (n + kSortTile - 1) / kSortTile;
    for (int k = 0; k < 2; ++k) { LB(cudaMalloc(&keys[k], (size_t)n * 4)); LB(cudaMalloc(&vals[k], (size_t)n * 4)); }
    LB(cudaMalloc(&counts, (size_t)kRadix * sort_blocks * 4));
    LB(cudaMalloc(&flags, (size_t)n * 4));
    LB(cudaMalloc(&want, (size_t)n)); LB(cudaMalloc(&keep, (size_t)n));
    LB(cudaMalloc(&left, (size_t)n * 4)); LB(cudaMalloc(&right, (size_t)n * 4));
    LB(cudaMalloc(&first, (size_t)n * 4)); LB(cudaMalloc(&last, (size_t)n * 4));
    LB(cudaMalloc(&parent, (size_t)2 * n * 4)); LB(cudaMalloc(&visits, (size_t)n * 4));
    LB(cudaMalloc(&bounds, 6 * 4)); LB(cudaMalloc(&boxes, (size_t)2 * n * 6 * 4));

    const int T = 256, G = (n + T - 1) / T;
    const int init[6] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000, (int)0x80800000, (int)0x80800000};   // ord(+FLT_MAX), ord(-FLT_MAX)
    LB(cudaMemcpyAsync(bounds, init, sizeof init, cudaMemcpyHostToDevice, stream));
    bounds_kernel<<<std::min(G, 1184), T, 0, stream>>>(d_verts, n, bounds);
    morton_kernel<<<G, T, 0, stream>>>(d_verts, n, bounds, keys[0], vals[0]);
    int cur = 0;
    for (int shift = 0; shift < 30; shift += kRadixBits) {
        radix_count_kernel<<<sort_blocks, kSortThreads, 0, stream>>>(keys[cur], n, shift, counts);
        scan_single_block_kernel<<<1, 1024, 0, stream>>>(counts, kRadix * sort_blocks);
        radix_scatter_kernel<<<sort_blocks, kSortThreads, 0, stream>>>(keys[cur], vals[cur], n, shift, counts, keys[cur ^ 1], vals[cur ^ 1]);
        cur ^= 1;
    }
    hierarchy_kernel<<<G, T, 0, stream>>>(keys[cur], n, left, right, first, last, parent);
    LB(cudaMemsetAsync(visits, 0, (size_t)n * 4, stream));
    refit_kernel<<<G, T, 0, stream>>>(d_verts, vals[cur], n, left, right, parent, boxes, visits);
    emit_want_kernel<<<G, T, 0, stream>>>(left, right, first, last, boxes, n, want);
    emit_flag_kernel<<<G, T, 0, stream>>>(first, last, parent, want, n, flags, keep);
    // exclusive scan of the n-1 flags -> dense node indices (root keeps index 0); the total lands in flags[n-1]
    LB(cudaMemsetAsync(flags + (n - 1), 0, 4, stream));
    scan_single_block_kernel<<<1, 1024, 0, stream>>>(flags, n);
    int h_bounds[6]; unsigned h_nodes = 0;
    LB(cudaMemcpyAsync(h_bounds, bounds, sizeof h_bounds, cudaMemcpyDeviceToHost, stream));
    LB(cudaMemcpyAsync(&h_nodes, flags + (n - 1), 4, cudaMemcpyDeviceToHost, stream));
    LB(cudaStreamSynchronize(stream));
    float ext = 0.f;
    for (int a = 0; a < 3; ++a) {
        const int lo_i = h_bounds[a], hi_i = h_bounds[3 + a];
        int lo_b = lo_i >= 0 ? lo_i : lo_i ^ 0x7fffffff, hi_b = hi_i >= 0 ? hi_i : hi_i ^ 0x7fffffff;
        float lo, hi;
        memcpy(&lo, &lo_b, 4); memcpy(&hi, &hi_b, 4);
        out->lo[a] = lo; out->hi[a] = hi;
        ext = std::max(ext, std::max(std::fabs(lo), std::fabs(hi)));
    }
    const float pad = bvh_pad(ext);
    for (int a = 0; a < 3; ++a) { out->lo[a] -= pad; out->hi[a] += pad; }
    out->n_nodes = (int)h_nodes;
    emit_kernel<<<G, T, 0, stream>>>(n, left, right, first, last, boxes, flags, keep, node_offset, slot_offset, pad, d_nodes);
    if (d_tris) tri_gather_kernel<<<G, T, 0, stream>>>(d_verts, d_mats, vals[cur], n, id_base, slot_offset, d_tris);
    if (d_order) LB(cudaMemcpyAsync(d_order, vals[cur], (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    LB(cudaGetLastError());
    LB(cudaStreamSynchronize(stream));
    cleanup();
    return cudaSuccess;
}

} // namespace arv2

// conv.cu -- sm_100a uniformly partitioned overlap-add (UPOLA) FFT convolver of libarv2.
//
// Replaces the cuFFT convolvers of the reference:
//   convoluteFromAudioBuffer (+load_sample_segment, multiply_samples_segment_and_ir, add)
//                              OR/kernels.cu:382-438,226-255,70-75   -> conv_block_spectra + conv_file
//   convoluteFromLiveInput (+complexCrossMultiplication, normalise, zip)
//                              OR/kernels.cu:327-377,450-487         -> conv_stream_step
// The reference runs one cuFFT of size ir_len per second of audio (file) or per callback
// (live) and re-plans / re-allocates every call.  Here the IR is cut into P partitions
// of `block` samples whose spectra are computed once per re-render (conv_ir_spectra);
// a step is a 2*block-point Stockham FFT of the newest input block, a multiply-
// accumulate of P spectra pairs, and one inverse FFT that yields both ears at once
// (Z = Y_left + i*Y_right).  All of it is one launch of 8-CTA thread-block clusters:
// each CTA accumulates 1/8 of the partitions (a producer lane feeds a shared-memory ring with
// bulk async copies, eight warps consume), the partial sums are reduced through distributed
// shared memory, and the cluster's rank 0 runs the inverse FFT + overlap-add.  Consecutive
// steps of a stream overlap through programmatic dependent launch (stream_step_kernel).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cooperative_groups.h>

#include "conv.cuh"

namespace cg = cooperative_groups;

namespace arv2 {

namespace {

#ifdef ARV2_CONV_TIMING
__device__ long long g_ct_marks[16];                 // clock64 marks of cluster 0, rank 0, thread 0 (debug build only)
#define CT_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_ct_marks[i] = clock64(); } while (0)
// -DARV2_CONV_TRACE (with ARV2_CONV_TIMING): globaltimer stamps of every step without printf: g_ct_trace[slot % 256][cta % 16][64],
// consumer thread 0 in [0, 40) (0 start, 1 ring ready, 2 + i after partition i, 36 mac done, 37 wait done, 38 end), producer
// lane in [40, 64) (40 + i when the copies of partition i are issued); read back with arv2_debug_conv_trace
__device__ unsigned long long g_ct_trace[256][16][96];
#ifdef ARV2_CONV_TRACE
// (stamps go to shared memory and are dumped when the step ends: no global store before the wait)
__shared__ unsigned long long sh_ct_trace[96];
#define CT_TRACE(slot, idx) do { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); sh_ct_trace[idx] = t_; } while (0)
#define CT_TRACE_DUMP(slot) do { __syncthreads(); if (threadIdx.x < 96) g_ct_trace[(slot) & 255][blockIdx.x & 15][threadIdx.x] = sh_ct_trace[threadIdx.x]; } while (0)
#else
#define CT_TRACE(slot, idx) do {} while (0)
#endif
#else
#define CT_MARK(i) do {} while (0)
#define CT_TRACE(slot, idx) do {} while (0)
#endif

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cjmul(float2 a) { return make_float2(-a.y, a.x); }   // i * a

// Stockham autosort FFT on shared memory, radix-4 passes plus one radix-2 pass when
// log2(N) is odd.  x holds the input, y is scratch; returns the buffer with the result.
// inverse = unnormalised conjugate transform.  All threads of the CTA must call it.
__device__ float2* fft_smem(float2* x, float2* y, int N, const float2* __restrict__ tw, bool inverse)
{
    int n = N, s = 1, ls = 0;
    while (n >= 4) {
        const int n1 = n >> 2;
        for (int b = threadIdx.x; b < (N >> 2); b += blockDim.x) {
            const int p = b >> ls, q = b & (s - 1);
            float2 w1 = tw[p * s], w2 = tw[2 * p * s], w3 = tw[3 * p * s];
            if (inverse) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
            const float2 a = x[q + s * p], bb = x[q + s * (p + n1)], c = x[q + s * (p + 2 * n1)], d = x[q + s * (p + 3 * n1)];
            const float2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(bb, d);
            float2 jbmd = cjmul(csub(bb, d));
            if (inverse) { jbmd.x = -jbmd.x; jbmd.y = -jbmd.y; }
            y[q + s * (4 * p + 0)] = cadd(apc, bpd);
            y[q + s * (4 * p + 1)] = cmul(w1, csub(amc, jbmd));
            y[q + s * (4 * p + 2)] = cmul(w2, csub(apc, bpd));
            y[q + s * (4 * p + 3)] = cmul(w3, cadd(amc, jbmd));
        }
        __syncthreads();
        float2* t = x; x = y; y = t;
        n >>= 2; s <<= 2; ls += 2;
    }
    if (n == 2) {
        for (int q = threadIdx.x; q < (N >> 1); q += blockDim.x) {
            const float2 a = x[q], b = x[q + s];
            y[q] = cadd(a, b);
            y[q + s] = csub(a, b);
        }
        __syncthreads();
        float2* t = x; x = y; y = t;
    }
    return x;
}

// Real block -> packed spectrum: FFT of [block samples, block zeros].
// `load(t)` gives sample t (< block).  Result written through `store(k, value)`.
template <class Load, class Store>
__device__ void forward_block(float2* bufa, float2* bufb, int block, const float2* tw, Load load, Store store)
{
    const int N = 2 * block;
    for (int t = threadIdx.x; t < N; t += blockDim.x) bufa[t] = make_float2(t < block ? load(t) : 0.f, 0.f);
    __syncthreads();
    const float2* F = fft_smem(bufa, bufb, N, tw, false);
    for (int k = threadIdx.x; k < block; k += blockDim.x)
        store(k, k == 0 ? make_float2(F[0].x, F[block].x) : F[k]);
    __syncthreads();
}

__global__ void __launch_bounds__(kConvThreads) ir_spectra_kernel(const float* __restrict__ h, int ir_len, int block, int P, int ear0,
                                                                  const float2* __restrict__ tw, float2* __restrict__ H)
{
    extern __shared__ float2 smem[];
    float2* bufa = smem; float2* bufb = smem + 2 * block;
    const int item = blockIdx.x / P, p = blockIdx.x % P;
    const float* src = h + (size_t)item * ir_len;
    float2* dst = H + ((size_t)p * 2 + ear0 + item) * block;      // [P][2 ears][block]: one bulk copy per partition
    forward_block(bufa, bufb, block, tw,
                  [&](int t) { const long long i = (long long)p * block + t; return i < ir_len ? src[i] : 0.f; },
                  [&](int k, float2 v) { dst[k] = v; });
}

__global__ void __launch_bounds__(kConvThreads) block_spectra_kernel(const float* __restrict__ x, long long n, long long seg_len,
                                                                     int blocks_per_seg, int block,
                                                                     const float2* __restrict__ tw, float2* __restrict__ X)
{
    extern __shared__ float2 smem[];
    float2* bufa = smem; float2* bufb = smem + 2 * block;
    const int seg = blockIdx.x / blocks_per_seg, j = blockIdx.x % blocks_per_seg;
    float2* dst = X + (size_t)blockIdx.x * block;
    forward_block(bufa, bufb, block, tw,
                  [&](int t) {
                      const long long off = (long long)j * block + t;
                      const long long i = (long long)seg * seg_len + off;
                      return (off < seg_len && i < n) ? x[i] : 0.f;
                  },
                  [&](int k, float2 v) { dst[k] = v; });
}

// ---- TMA (cp.async.bulk) staging --------------------------------------------------------
// One CTA streams ~300 KB of spectra per step with only 8 warps, far too few loads in flight
// for plain LDGs (r01 profile: 18.5 us per step, long_scoreboard 20 warps/issue).  The rows
// of a partition (X: block*8 B, H_left + H_right: 2*block*8 B, adjacent) are instead pulled into a
// shared-memory ring by bulk async copies that complete on an mbarrier; the consumer threads only
// read shared memory.
// The ring takes what two co-resident CTAs per SM leave (2 x 104 KB): with one CTA per SM the 8-CTA clusters of a
// 16-source step no longer fit the GPCs in one wave (26 us per step instead of 16, r05), and a step could not share
// the SMs with its predecessor.  block = 512: 8 stages x 12 KB in flight per CTA.
#ifndef ARV2_CONV_STAGES
#define ARV2_CONV_STAGES 8
#endif
#ifndef ARV2_CONV_FFTCOST
#define ARV2_CONV_FFTCOST 8
#endif
constexpr int kStages = ARV2_CONV_STAGES;                  // stages (mbarrier pairs) of the ring when two CTAs share an SM
constexpr int kMaxStages = 2 * kStages;                    // ... of the deep ring (one CTA per SM)
__host__ __device__ constexpr int ring_stages(int block)
{
    const int budget = 104 * 1024 - 2 * block * (int)sizeof(float2);       // minus the twiddle table
    const int fit = budget / (3 * block * (int)sizeof(float2));
    return fit < 2 ? 2 : (fit > kStages ? kStages : fit);
}
// The deep ring of a stream with few sources (2 x sources x 8 CTAs <= SMs: a step and its successor find SMs of their
// own, nothing has to share one): twice the bytes in flight per CTA.  A bulk copy from L2 takes ~1.2 us under load, so
// 8 stages of 12 KB stream at 12 KB / 156 ns per CTA whatever the machine could deliver (r09 stamps); 16 stages double it.
__host__ __device__ constexpr int ring_stages_deep(int block)
{
    const int budget = 208 * 1024 - 2 * block * (int)sizeof(float2);
    const int fit = budget / (3 * block * (int)sizeof(float2));
    return fit < 2 ? 2 : (fit > kMaxStages ? kMaxStages : fit);
}
constexpr int kConvStepThreads = kConvThreads + 32;     // stream / file kernels: kConvThreads consumers + one producer warp
constexpr int kFftCostInPartitions = ARV2_CONV_FFTCOST;   // rank 0 also runs the forward FFT and takes this many partitions less per peer (8, 16, 24: same step time, r07)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
#ifdef ARV2_CONV_BULK_CTA
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
#endif
}

// Programmatic dependent launch (sm_90+): step k+1 may become resident while step k is still running; everything
// that depends on step k comes after pdl_wait(), which returns once step k has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// acc += X[k] * H_e[k] for this thread's bins k = tid*BPT + i (contiguous, so that a
// thread's BPT bins are one 8*BPT-byte shared-memory access); bin 0 holds (DC, Nyquist).
__device__ __forceinline__ void mac_bin(int k, float2 xv, float2 hl, float2 hr, float2& accL, float2& accR)
{
    if (k == 0) {   // packed (DC, Nyquist): two real products
        accL.x = fmaf(xv.x, hl.x, accL.x); accL.y = fmaf(xv.y, hl.y, accL.y);
        accR.x = fmaf(xv.x, hr.x, accR.x); accR.y = fmaf(xv.y, hr.y, accR.y);
    } else {
        accL.x = fmaf(xv.x, hl.x, fmaf(-xv.y, hl.y, accL.x)); accL.y = fmaf(xv.x, hl.y, fmaf(xv.y, hl.x, accL.y));
        accR.x = fmaf(xv.x, hr.x, fmaf(-xv.y, hr.y, accR.x)); accR.y = fmaf(xv.x, hr.y, fmaf(xv.y, hr.x, accR.y));
    }
}

template <int BPT>
__device__ __forceinline__ void mac_rows(const float2* X, const float2* HL, const float2* HR, int block, float2 accL[BPT], float2 accR[BPT])
{
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const int k = threadIdx.x * BPT + i;
        if (k < block) mac_bin(k, X[k], HL[k], HR[k], accL[i], accR[i]);
    }
}

// Rows of the partitions one CTA accumulates: partition i pairs the input spectrum at x (block float2) with the
// two IR spectra at h (2*block float2, left then right); x and h advance by fixed steps, x wraps inside the
// delay line.
struct MacRows {
    const float2* x; long long x_step; const float2* x_lo; long long x_wrap;     // after a partition: x += x_step; if (x < x_lo) x += x_wrap
    const float2* h; long long h_step;
};

// Multiply-accumulate `n` partitions through the TMA ring; ring: float2[stages][3][block]; all threads call.
// Warp-specialised: the CTA's last warp is the producer (one lane waits for a stage to be released and issues its
// two bulk copies: 4 KB of input spectrum, 8 KB of IR spectra at block = 512), the first kConvThreads threads consume
// (wait for the stage's bytes, accumulate, release the stage with one arrive per warp).  No CTA-wide barrier inside
// the loop.  The producer's instruction count is what bounds the loop (r07: thread 0 issuing three copies between two
// __syncthreads: 350 ns per partition whatever the ring depth; a producer lane recomputing the three row addresses
// per partition: 85 instructions, 280 ns), hence the stepped pointers and the interleaved IR rows.
template <int BPT>
__device__ __forceinline__ void mac_pipeline(float2* ring, unsigned long long* full, unsigned long long* empty, int n, int block,
                                             MacRows r, float2 accL[BPT], float2 accR[BPT], int stages = 0, int trace_slot = -1)
{
    (void)trace_slot;
    const unsigned row_bytes = (unsigned)block * sizeof(float2);
    const int S = stages > 0 ? stages : ring_stages(block);
    if (threadIdx.x >= kConvThreads) {
        if (threadIdx.x == kConvThreads) {
            int s = 0;
            unsigned parity = 1;                          // parity of the phase that precedes the first release
            float2* dst = ring;
            for (int i = 0; i < n; ++i) {
                if (i >= S) mbar_wait(&empty[s], parity);
                if (trace_slot >= 0 && i < 24) CT_TRACE(trace_slot, 64 + i);
#if defined(ARV2_CONV_TRACE) && defined(ARV2_CONV_TRACE_FINE)
                // where inside the issue of partition i the producer spends its time: [40 + 3j] after expect_tx, + 1 after the
                // copy of the input spectrum, + 2 after the copy of the IR spectra, for i = 10 + j, j < 8
                mbar_expect_tx(&full[s], 3 * row_bytes);
                if (trace_slot >= 0 && i >= 10 && i < 18) CT_TRACE(trace_slot, 40 + 3 * (i - 10));
                bulk_g2s(dst, r.x, row_bytes, &full[s]);
                if (trace_slot >= 0 && i >= 10 && i < 18) CT_TRACE(trace_slot, 41 + 3 * (i - 10));
                bulk_g2s(dst + block, r.h, 2 * row_bytes, &full[s]);
                if (trace_slot >= 0 && i >= 10 && i < 18) CT_TRACE(trace_slot, 42 + 3 * (i - 10));
#else
                mbar_expect_tx(&full[s], 3 * row_bytes);
                bulk_g2s(dst, r.x, row_bytes, &full[s]);
                bulk_g2s(dst + block, r.h, 2 * row_bytes, &full[s]);
                if (trace_slot >= 0 && i < 24) CT_TRACE(trace_slot, 40 + i);
#endif
                r.x += r.x_step; if (r.x < r.x_lo) r.x += r.x_wrap;
                r.h += r.h_step;
                dst += 3 * block;
                if (++s == S) { s = 0; dst = ring; parity ^= 1u; }
            }
        }
        return;
    }
    int s = 0;
    unsigned parity = 0;
    const float2* src = ring;
    for (int i = 0; i < n; ++i) {
        mbar_wait(&full[s], parity);
        mac_rows<BPT>(src, src + block, src + 2 * block, block, accL, accR);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
        if (trace_slot >= 0 && threadIdx.x == 0 && i < 34) CT_TRACE(trace_slot, 2 + i);
        src += 3 * block;
        if (++s == S) { s = 0; src = ring; parity ^= 1u; }
    }
}

// The same pipeline with TWO partitions per stage, for callers whose partitions are consecutive rows (x_step = -block,
// h_step = 2*block: stream_step_kernel): the IR rows of partitions i, i+1 are one 4-row copy, the two input spectra one
// 2-row copy unless the delay line wraps between them -- half as many bulk copies per partition.  A step launched as a
// programmatic dependent stops issuing after its 24th bulk copy until its predecessor has completed (r09 traces), and the
// producer lane's issue rate (~65 ns per copy) bounds the phase besides.  Stage layout (rows of `block` float2):
// [X_{i+1}, X_i, HL_i, HR_i, HL_{i+1}, HR_{i+1}]; the accumulation order (i, then i+1) is that of mac_pipeline, so the
// results are bit-identical.  ring: float2[stages / 2][6][block].
template <int BPT>
__device__ __forceinline__ void mac_pipeline_pairs(float2* ring, unsigned long long* full, unsigned long long* empty, int n, int block,
                                                   MacRows r, float2 accL[BPT], float2 accR[BPT], int stages = 0, int trace_slot = -1)
{
    (void)trace_slot;
    const unsigned row_bytes = (unsigned)block * sizeof(float2);
    const int S1 = stages > 0 ? stages : ring_stages(block);
    const int S = S1 / 2 > 0 ? S1 / 2 : 1;                 // pair stages
    const int pairs = (n + 1) / 2;
    if (threadIdx.x >= kConvThreads) {
        if (threadIdx.x == kConvThreads) {
            int s = 0;
            unsigned parity = 1;
            float2* dst = ring;
            for (int j = 0; j < pairs; ++j) {
                const bool two = 2 * j + 1 < n;
                if (j >= S) mbar_wait(&empty[s], parity);
                const float2* xa = r.x;
                const float2* xb = xa + r.x_step; if (xb < r.x_lo) xb += r.x_wrap;
                mbar_expect_tx(&full[s], (two ? 6u : 3u) * row_bytes);
                if (two) {
                    if (xb + block == xa) bulk_g2s(dst, xb, 2 * row_bytes, &full[s]);
                    else { bulk_g2s(dst, xb, row_bytes, &full[s]); bulk_g2s(dst + block, xa, row_bytes, &full[s]); }
                    bulk_g2s(dst + 2 * block, r.h, 4 * row_bytes, &full[s]);
                    r.x = xb + r.x_step; if (r.x < r.x_lo) r.x += r.x_wrap;
                    r.h += 2 * r.h_step;
                } else {
                    bulk_g2s(dst + block, xa, row_bytes, &full[s]);
                    bulk_g2s(dst + 2 * block, r.h, 2 * row_bytes, &full[s]);
                }
                if (trace_slot >= 0 && j < 24) CT_TRACE(trace_slot, 40 + j);
                dst += 6 * block;
                if (++s == S) { s = 0; dst = ring; parity ^= 1u; }
            }
        }
        return;
    }
    int s = 0;
    unsigned parity = 0;
    const float2* src = ring;
    for (int j = 0; j < pairs; ++j) {
        mbar_wait(&full[s], parity);
        mac_rows<BPT>(src + block, src + 2 * block, src + 3 * block, block, accL, accR);
        if (2 * j + 1 < n) mac_rows<BPT>(src, src + 4 * block, src + 5 * block, block, accL, accR);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
        if (trace_slot >= 0 && threadIdx.x == 0 && j < 34) CT_TRACE(trace_slot, 2 + j);
        src += 6 * block;
        if (++s == S) { s = 0; src = ring; parity ^= 1u; }
    }
}

// Cluster-wide reduction of the per-CTA partial sums, then the inverse transform on rank 0.  Rank r owns the bins
// [r*block/C, (r+1)*block/C) of both ears: it sums the C partial sums through distributed shared memory (all C
// remote loads in flight at once) and writes Z = YL + i*YR, already unpacked to the full Hermitian layout, straight
// into rank 0's FFT buffer.  Rank 0 returns the time-domain buffer (re = left, im = right, unnormalised); other
// ranks return nullptr.
template <int BPT, int C = kConvCluster>
__device__ float2* reduce_and_inverse(cg::cluster_group& cluster, float2* part, float2* bufa, float2* bufb,
                                      int block, const float2* tw, const float2 accL[BPT], const float2 accR[BPT])
{
    const unsigned rank = cluster.block_rank();
    if (threadIdx.x < kConvThreads) {
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int k = threadIdx.x * BPT + i;
            if (k < block) { part[k] = accL[i]; part[block + k] = accR[i]; }
        }
    }
    CT_MARK(4);
    cluster.sync();
    CT_MARK(5);
    {
        const int N = 2 * block;
        const int per = block / C;                       // block is a power of two >= 64, C = 8
        const float2* remote[C];
#pragma unroll
        for (int r = 0; r < C; ++r) remote[r] = cluster.map_shared_rank(part, r);
        float2* z = cluster.map_shared_rank(bufa, 0);
        for (int e = threadIdx.x; e < per; e += blockDim.x) {
            const int k = (int)rank * per + e;
            // (eight remote loads of each ear in flight at a time; the ranks are summed in rank order)
            float2 l = make_float2(0.f, 0.f), rr = make_float2(0.f, 0.f);
#pragma unroll
            for (int r0 = 0; r0 < C; r0 += 8) {
                float2 vl[8], vr[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) { vl[r] = remote[r0 + r][k]; vr[r] = remote[r0 + r][block + k]; }
                if (r0 == 0) { l = vl[0]; rr = vr[0]; }
#pragma unroll
                for (int r = (r0 == 0 ? 1 : 0); r < 8; ++r) { l.x += vl[r].x; l.y += vl[r].y; rr.x += vr[r].x; rr.y += vr[r].y; }
            }
            if (k == 0) {
                z[0] = make_float2(l.x, rr.x);
                z[block] = make_float2(l.y, rr.y);
            } else {
                z[k] = make_float2(l.x - rr.y, l.y + rr.x);
                z[N - k] = make_float2(l.x + rr.y, rr.x - l.y);
            }
        }
    }
    CT_MARK(6);
    cluster.sync();
    CT_MARK(7);
    if (rank != 0) return nullptr;
    float2* yy = fft_smem(bufa, bufb, 2 * block, tw, true);
    CT_MARK(8);
    return yy;
}

#ifdef ARV2_CONV_TIMING
#define CT_DECL unsigned long long ct_[10]; long long cc_[10]; int ct_n = 0;
#define CT_STAMP() do { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); cc_[ct_n] = clock64(); ct_[ct_n++] = t_; } while (0)
#else
#define CT_DECL
#define CT_STAMP() do {} while (0)
#endif

template <int BPT, int CL = kConvCluster>
__global__ void __launch_bounds__(kConvStepThreads) stream_step_kernel(const ConvStreamArgs a)
{
    CT_DECL
    CT_STAMP();
#ifdef ARV2_CONV_TRACE
    if (threadIdx.x < 96) sh_ct_trace[threadIdx.x] = 0ull;
    __syncthreads();
#endif
    if (threadIdx.x == 0) CT_TRACE(a.slot, 0);
    extern __shared__ __align__(128) float2 smem[];
    __shared__ unsigned long long full[kMaxStages], empty[kMaxStages];
    const int block = a.block, N = 2 * block;
    // twiddles, then the TMA ring; the FFT buffers and the partial sums are only needed once the ring has drained
    // and live on top of it (a deeper ring instead of 24 KB of idle buffers)
    float2* stw = smem; float2* ring = stw + N;
    float2* bufa = ring; float2* bufb = bufa + N; float2* part = bufb + N;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    constexpr unsigned C = CL;
    // a 16-CTA cluster per source (streams of few sources): every rank but 0 streams 12 old partitions = 144 KB, what a
    // programmatic dependent gets through before its predecessor has completed (r09 traces); rank 0 takes 6
    constexpr int kFftCost = CL == 16 ? 6 : kFftCostInPartitions;
    const int src = blockIdx.x / C;
    float2* fdl = a.fdl + (size_t)src * a.P * block;
    const float2* H = a.H[src];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConvThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // twiddles (5 passes x 3 gathers per FFT) into shared memory, by the consumer threads while the producer already
    // fills the ring; first used after the __syncthreads that follows the ring
    if (threadIdx.x < kConvThreads)
        for (int t = threadIdx.x; t < N; t += kConvThreads) stw[t] = a.tw[t];

#ifdef ARV2_CONV_EMPTY_STEP
    // TIMING EXPERIMENT ONLY: the launch mechanism alone (8-CTA clusters, programmatic dependent launch, same shared memory)
    { pdl_wait(); pdl_launch_dependents(); return; }
#endif
#ifdef ARV2_CONV_EARLY_TRIGGER
    // TIMING EXPERIMENT ONLY (results are wrong: step k+1 may read partition 2 before step k-1 published it): lets the
    // successor in at once, to see how much of the 6.4 us period is launch latency + pre-wait work (r09)
    pdl_launch_dependents();
#endif
    float2 accL[BPT], accR[BPT];
#pragma unroll
    for (int i = 0; i < BPT; ++i) { accL[i] = make_float2(0.f, 0.f); accR[i] = make_float2(0.f, 0.f); }
    // Consecutive steps overlap (programmatic dependent launch): partitions 2..P-1 pair spectra that step k-1 and
    // older steps wrote into the delay line with the (static) IR spectra, so they are accumulated BEFORE waiting
    // for step k; partition 1 (written by step k), the forward FFT of the newest block (its slot is the one step k
    // read as its oldest partition) and the tail update come after the wait.  A step lets its successor in only
    // after its own wait, so "step k-1 is complete" holds whenever step k+1 runs.
    // Rank 0 runs the forward FFT (~8 partitions' worth of time, measured with clock64) and therefore takes a
    // shorter contiguous range of the old partitions; the rest is split evenly over ranks 1..C-1.
    const int T = max(0, a.P - 2);
    int n0 = (T - kFftCost * ((int)C - 1)) / (int)C;
    n0 = max(0, min(T, n0));
    int first, n;
    if (rank == 0) { first = 2; n = n0; }
    else {
        const int rest = T - n0, per = rest / ((int)C - 1), extra = rest % ((int)C - 1), r1 = (int)rank - 1;
        first = 2 + n0 + r1 * per + min(r1, extra);
        n = per + (r1 < extra ? 1 : 0);
    }
    // IR rows of the partition this rank multiplies after the wait (rank 0: partition 0, rank C-1: partition 1):
    // static data, fetched now so that their latency hides under the ring
    const int pq = rank == 0 ? 0 : 1;
    const bool late = (rank == 0 || (rank == C - 1 && a.P > 1)) && threadIdx.x < kConvThreads;
    float2 hqL[BPT], hqR[BPT];
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const int k = threadIdx.x * BPT + i;
        const bool on = late && k < block;
        hqL[i] = on ? H[(size_t)pq * 2 * block + k] : make_float2(0.f, 0.f);
        hqR[i] = on ? H[((size_t)pq * 2 + 1) * block + k] : make_float2(0.f, 0.f);
    }
    if (rank == 0 && threadIdx.x < kConvThreads) {
        // the newest input block and the tails are read after the wait; a prefetch into L2 (the point of coherence:
        // a line fetched early still sees what the previous step or the producer of `in` writes later) takes part of
        // their round trip off the chain that follows the wait
        const char* pin = (const char*)(a.in + (size_t)src * block);
        const char* ptl = (const char*)(a.tail + (size_t)src * 2 * block);
        const int t128 = threadIdx.x * 128;
        if (t128 < block * 4) prefetch_l2(pin + t128);
        if (t128 < block * 8) prefetch_l2(ptl + t128);
    }
    {
        int s0 = a.slot - first; if (s0 < 0) s0 += a.P;
        MacRows r;
        r.x = fdl + (size_t)s0 * block; r.x_step = -(long long)block; r.x_lo = fdl; r.x_wrap = (long long)a.P * block;
        r.h = H + (size_t)first * 2 * block; r.h_step = 2 * (long long)block;
        if (threadIdx.x == 0) CT_TRACE(a.slot, 1);
#ifdef ARV2_CONV_SINGLE_COPIES
        mac_pipeline<BPT>(ring, full, empty, n, block, r, accL, accR, a.stages, a.slot);
#else
        mac_pipeline_pairs<BPT>(ring, full, empty, n, block, r, accL, accR, a.stages, a.slot);
#endif
        if (threadIdx.x == 0) CT_TRACE(a.slot, 36);
    }
    // newest block: forward FFT (rank 0; every thread of the CTA).  When the caller guarantees that `in` was complete
    // before the PREVIOUS step passed its wait (blocks 1.. of one call: the whole call's input was there before block 0),
    // it runs before this step's wait -- the transform needs nothing from step k, only its result may not be published
    // into the delay line yet (the slot is the one step k reads as its oldest partition) -- which takes the input's
    // round trip and the FFT (2.2 of 5.7 us) off the chain that serialises consecutive steps.
    const float2* F = nullptr;
    auto forward_fft = [&]() {
        const float* in = a.in + (size_t)src * block;
        CT_MARK(0);
        // (read through L2: before the wait nothing has invalidated this SM's L1, which may hold the block of an older call)
        for (int t = threadIdx.x; t < N; t += blockDim.x) bufa[t] = make_float2(t < block ? __ldcg(in + t) : 0.f, 0.f);
        __syncthreads();
        CT_MARK(1);
        F = fft_smem(bufa, bufb, N, stw, false);
        CT_MARK(2);
    };
    if (a.early_input && rank == 0) {
        __syncthreads();                                  // the ring has drained: its memory becomes bufa / bufb / part
        forward_fft();
    }
    CT_STAMP();
    pdl_wait();
#if !defined(ARV2_CONV_EARLY_TRIGGER) && !defined(ARV2_CONV_LATE_TRIGGER)
    pdl_launch_dependents();
#endif
    if (threadIdx.x == 0) CT_TRACE(a.slot, 37);
    CT_STAMP();
    __syncthreads();                                      // the ring has drained: its memory becomes bufa / bufb / part
    float tl[BPT], tr[BPT];                               // overlap-add tails of this thread's output samples (rank 0)
    float* tail = a.tail + (size_t)src * 2 * block;
    if (rank == 0) {
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int t = threadIdx.x * BPT + i;
            const bool on = threadIdx.x < kConvThreads && t < block;
            tl[i] = on ? tail[t] : 0.f; tr[i] = on ? tail[block + t] : 0.f;
        }
        // publish the newest block's spectrum into the frequency-domain delay line, multiply by partition 0
        float2* slot = fdl + (size_t)a.slot * block;
        if (!F) forward_fft();
        if (threadIdx.x < kConvThreads) {
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
                const int k = threadIdx.x * BPT + i;
                if (k < block) {
                    const float2 xv = k == 0 ? make_float2(F[0].x, F[block].x) : F[k];
                    slot[k] = xv;
                    mac_bin(k, xv, hqL[i], hqR[i], accL[i], accR[i]);
                }
            }
        }
        __syncthreads();                                  // F (bufa or bufb) is read above; part and bufa are written next
        CT_MARK(3);
    }
    if (rank == C - 1 && a.P > 1 && threadIdx.x < kConvThreads) {
        // the previous block: its spectrum was published by the step this one waited for
        const int s1 = a.slot >= 1 ? a.slot - 1 : a.slot - 1 + a.P;
        const float2* X1 = fdl + (size_t)s1 * block;
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int k = threadIdx.x * BPT + i;
            if (k < block) mac_bin(k, X1[k], hqL[i], hqR[i], accL[i], accR[i]);
        }
    }
    CT_STAMP();
#ifdef ARV2_CONV_LATE_TRIGGER
    pdl_launch_dependents();        // timing experiment (r09): the successor is let in only when this step is in its reduction
#endif
    float2* y = reduce_and_inverse<BPT, CL>(cluster, part, bufa, bufb, block, stw, accL, accR);
    CT_STAMP();
    if (threadIdx.x == 0) CT_TRACE(a.slot, 38);
#ifdef ARV2_CONV_TRACE
    CT_TRACE_DUMP(a.slot);
#endif
#if defined(ARV2_CONV_TIMING) && !defined(ARV2_CONV_TRACE)
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.slot == 100)
        printf("rank 0 cycles: in-load %lld fwd-fft %lld publish+p0 %lld | part-store %lld sync1 %lld dsmem-reduce %lld sync2 %lld inv-fft %lld\n",
               g_ct_marks[1] - g_ct_marks[0], g_ct_marks[2] - g_ct_marks[1], g_ct_marks[3] - g_ct_marks[2], g_ct_marks[4] - g_ct_marks[3],
               g_ct_marks[5] - g_ct_marks[4], g_ct_marks[6] - g_ct_marks[5], g_ct_marks[7] - g_ct_marks[6], g_ct_marks[8] - g_ct_marks[7]);
    if ((blockIdx.x == 0 || blockIdx.x == 7) && threadIdx.x == 0 && a.slot >= 101 && a.slot <= 104)       // the timeline of consecutive steps
        printf("T slot %d cta %d: start %llu mac-done %llu wait-done %llu fft+p01-done %llu end %llu (ns mod 1e6)\n", a.slot, blockIdx.x,
               ct_[0] % 1000000ull, ct_[1] % 1000000ull, ct_[2] % 1000000ull, ct_[3] % 1000000ull, ct_[4] % 1000000ull);
    if ((blockIdx.x == 0 || blockIdx.x == 7) && threadIdx.x == 0 && a.slot == 100)
        printf("cta %d: mac %llu wait %llu fft+p01 %llu reduce+ifft %llu ns (start %llu); mac %lld cycles, whole %lld cycles / %llu ns\n", blockIdx.x,
               ct_[1] - ct_[0], ct_[2] - ct_[1], ct_[3] - ct_[2], ct_[4] - ct_[3], ct_[0] % 1000000ull, cc_[1] - cc_[0], cc_[4] - cc_[0], ct_[4] - ct_[0]);
#endif
    if (!y || threadIdx.x >= kConvThreads) return;
    const float sc = 1.0f / (float)N;
    float* out = a.out + (size_t)src * 2 * block;
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const int t = threadIdx.x * BPT + i;
        if (t < block) {
            const float2 lo = y[t], hi = y[block + t];
            out[t] = fmaf(lo.x, sc, tl[i]);
            out[block + t] = fmaf(lo.y, sc, tr[i]);
            tail[t] = hi.x * sc;
            tail[block + t] = hi.y * sc;
        }
    }
}

// ---------------------------------------------------------------------------------------
// stream_blocks_kernel (experiment, ARV2_CONV_PERSISTENT=1; measured SLOWER, see below): the n_blocks consecutive steps
// of one call (arv2_stream_process_device_blocks: an RtAudio callback of the reference carries 4096 frames = 8 blocks)
// in ONE cluster launch.  stream_step_kernel is launched once per block, and even with programmatic dependent launch two
// dependent cluster launches are ~2 us apart; here every source's 8-CTA cluster loops over the blocks.  Per block: all
// ranks accumulate their old partitions through the TMA ring; rank 0 transforms the newest input block, publishes its
// spectrum in the delay line and multiplies partition 0, rank C-1 takes partition 1 (published one block earlier);
// cluster-wide DSMEM reduction; rank 0 inverse-transforms and overlap-adds with the tails it keeps in registers.
// Results are bit-identical to n_blocks launches of stream_step_kernel (same partition split, same summation order).
// Measured (r08, profiles/r08_conv.md): 8.6 us per block at 1-4 sources, 11.0 at 16, against 6.5 / 8.1 for one launch
// per block: the period of a looping cluster is rank 0's whole chain (its share of the ring + 5.7 us of input load, two
// 1024-point FFTs and the reduction), while separate launches keep two steps resident per SM and run step k+1's ring
// under step k's FFTs.  Kept as the starting point for splitting the two FFTs over two ranks.
#ifndef ARV2_CONV_PERSIST_FFTCOST
#define ARV2_CONV_PERSIST_FFTCOST 8     // same split as stream_step_kernel: the accumulation order must not change
#endif
template <int BPT>
__global__ void __launch_bounds__(kConvStepThreads) stream_blocks_kernel(const ConvStreamArgs a)
{
    extern __shared__ __align__(128) float2 smem[];
    __shared__ unsigned long long full[kStages], empty[kStages];
    const int block = a.block, N = 2 * block;
    float2* stw = smem; float2* ring = stw + N;
    float2* bufa = ring; float2* bufb = bufa + N; float2* part = bufb + N;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    constexpr unsigned C = kConvCluster;
    const int src = blockIdx.x / C;
    float2* fdl = a.fdl + (size_t)src * a.P * block;
    const float2* H = a.H[src];
    const bool consumer = threadIdx.x < kConvThreads;
    if (consumer)
        for (int t = threadIdx.x; t < N; t += kConvThreads) stw[t] = a.tw[t];

    // the split of the old partitions over the ranks (as in stream_step_kernel)
    const int T = max(0, a.P - 2);
    int n0 = (T - kFftCostInPartitions * ((int)C - 1)) / (int)C;
    n0 = max(0, min(T, n0));
    int first, n;
    if (rank == 0) { first = 2; n = n0; }
    else {
        const int rest = T - n0, per = rest / ((int)C - 1), extra = rest % ((int)C - 1), r1 = (int)rank - 1;
        first = 2 + n0 + r1 * per + min(r1, extra);
        n = per + (r1 < extra ? 1 : 0);
    }
    // static IR rows of the partition this rank multiplies late (rank 0: partition 0, rank C-1: partition 1), and the
    // overlap-add tails of this thread's output samples (rank 0): registers for the whole call
    const int pq = rank == 0 ? 0 : 1;
    const bool late = (rank == 0 || (rank == C - 1 && a.P > 1)) && consumer;
    float2 hqL[BPT], hqR[BPT];
    float tl[BPT], tr[BPT];
    float* tail = a.tail + (size_t)src * 2 * block;
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const int k = threadIdx.x * BPT + i;
        const bool on = late && k < block;
        hqL[i] = on ? H[(size_t)pq * 2 * block + k] : make_float2(0.f, 0.f);
        hqR[i] = on ? H[((size_t)pq * 2 + 1) * block + k] : make_float2(0.f, 0.f);
        const bool ont = rank == 0 && consumer && k < block;
        tl[i] = ont ? tail[k] : 0.f; tr[i] = ont ? tail[block + k] : 0.f;
    }
    const size_t in_step = (size_t)a.n_src * block, out_step = 2 * in_step;
    const float sc = 1.0f / (float)N;

    for (int b = 0; b < a.n_blocks; ++b) {
        const int slot = (a.slot + b) % a.P;
        // the ring's barriers start every block afresh (its memory was the FFT buffers of the previous block)
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int st = 0; st < kStages; ++st) {
                if (b > 0) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[st])) : "memory"); asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&empty[st])) : "memory"); }
                mbar_init(&full[st], 1); mbar_init(&empty[st], kConvThreads / 32);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        float2 accL[BPT], accR[BPT];
#pragma unroll
        for (int i = 0; i < BPT; ++i) { accL[i] = make_float2(0.f, 0.f); accR[i] = make_float2(0.f, 0.f); }
        {
            int s0 = slot - first; if (s0 < 0) s0 += a.P;
            MacRows r;
            r.x = fdl + (size_t)s0 * block; r.x_step = -(long long)block; r.x_lo = fdl; r.x_wrap = (long long)a.P * block;
            r.h = H + (size_t)first * 2 * block; r.h_step = 2 * (long long)block;
            mac_pipeline<BPT>(ring, full, empty, n, block, r, accL, accR);
        }
        __syncthreads();                                  // the ring has drained: its memory becomes bufa / bufb / part
        if (rank == 0) {
            const float* in = a.in + (size_t)b * in_step + (size_t)src * block;
            float2* slot_row = fdl + (size_t)slot * block;
            for (int t = threadIdx.x; t < N; t += blockDim.x) bufa[t] = make_float2(t < block ? in[t] : 0.f, 0.f);
            __syncthreads();
            const float2* F = fft_smem(bufa, bufb, N, stw, false);
            if (consumer) {
#pragma unroll
                for (int i = 0; i < BPT; ++i) {
                    const int k = threadIdx.x * BPT + i;
                    if (k < block) {
                        const float2 xv = k == 0 ? make_float2(F[0].x, F[block].x) : F[k];
                        slot_row[k] = xv;
                        mac_bin(k, xv, hqL[i], hqR[i], accL[i], accR[i]);
                    }
                }
            }
            // the spectrum just stored is read by the other CTAs' bulk copies (async proxy) from the next block on
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();                              // F (bufa or bufb) is read above; part and bufa are written next
        }
        if (rank == C - 1 && a.P > 1 && consumer) {
            // the previous block: published by rank 0 one iteration (or one call) ago, before a cluster barrier
            const int s1 = slot >= 1 ? slot - 1 : slot - 1 + a.P;
            const float2* X1 = fdl + (size_t)s1 * block;
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
                const int k = threadIdx.x * BPT + i;
                if (k < block) mac_bin(k, __ldcg(X1 + k), hqL[i], hqR[i], accL[i], accR[i]);
            }
        }
        float2* y = reduce_and_inverse<BPT>(cluster, part, bufa, bufb, block, stw, accL, accR);
        if (y && consumer) {
            float* out = a.out + (size_t)b * out_step + (size_t)src * 2 * block;
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
                const int t = threadIdx.x * BPT + i;
                if (t < block) {
                    const float2 lo = y[t], hi = y[block + t];
                    out[t] = fmaf(lo.x, sc, tl[i]);
                    out[block + t] = fmaf(lo.y, sc, tr[i]);
                    tl[i] = hi.x * sc; tr[i] = hi.y * sc;
                }
            }
        }
    }
    if (rank == 0 && consumer) {
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int t = threadIdx.x * BPT + i;
            if (t < block) { tail[t] = tl[i]; tail[block + t] = tr[i]; }
        }
    }
}

template <int BPT>
__global__ void __launch_bounds__(kConvStepThreads) file_kernel(const ConvFileArgs a, int out_blocks_per_seg)
{
    extern __shared__ __align__(128) float2 smem[];
    __shared__ unsigned long long full[kStages], empty[kStages];
    const int block = a.block, N = 2 * block;
    float2* stw = smem; float2* ring = stw + N;                    // same layout as stream_step_kernel
    float2* bufa = ring; float2* bufb = bufa + N; float2* part = bufb + N;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const unsigned C = cluster.num_blocks();
    const int cid = blockIdx.x / C;
    const int seg = cid / out_blocks_per_seg, j = cid % out_blocks_per_seg;
    const float2* X = a.X + (size_t)seg * a.blocks_per_seg * block;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConvThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // twiddles (5 passes x 3 gathers per FFT) into shared memory, by the consumer threads while the producer already
    // fills the ring; first used after the __syncthreads that follows the ring
    if (threadIdx.x < kConvThreads)
        for (int t = threadIdx.x; t < N; t += kConvThreads) stw[t] = a.tw[t];

    float2 accL[BPT], accR[BPT];
#pragma unroll
    for (int i = 0; i < BPT; ++i) { accL[i] = make_float2(0.f, 0.f); accR[i] = make_float2(0.f, 0.f); }
    const int p_lo = max(0, j - (a.blocks_per_seg - 1)), p_hi = min(a.P - 1, j);
    const int first = p_lo + (int)rank;
    const int n = first <= p_hi ? (p_hi - first) / (int)C + 1 : 0;
    {
        MacRows r;
        r.x = X + (size_t)(j - first) * block; r.x_step = -(long long)C * block; r.x_lo = nullptr; r.x_wrap = 0;
        r.h = a.H + (size_t)first * 2 * block; r.h_step = 2 * (long long)C * block;
        mac_pipeline<BPT>(ring, full, empty, n, block, r, accL, accR);
    }
    __syncthreads();                                      // the ring has drained: its memory becomes bufa / bufb / part
    float2* y = reduce_and_inverse<BPT>(cluster, part, bufa, bufb, block, stw, accL, accR);
    if (!y) return;
    const float sc = a.gain / (float)N;
    const long long seg_base = (long long)seg * a.seg_len;
    const long long limit = min(a.seg_out, a.n - seg_base);
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        long long i = (long long)j * block + t;
        if (a.wrap > 0 && i >= a.wrap) i -= a.wrap;
        if (i < limit) {
            atomicAdd(a.out_l + seg_base + i, y[t].x * sc);
            atomicAdd(a.out_r + seg_base + i, y[t].y * sc);
        }
    }
}

size_t fft_smem_bytes(int block) { return (size_t)4 * block * sizeof(float2); }
size_t step_smem_bytes(int block, int stages) { return (size_t)(2 + 3 * (stages > 0 ? stages : ring_stages(block))) * block * sizeof(float2); }

template <class K>
cudaError_t launch_cluster(K kernel, unsigned grid, size_t smem, cudaStream_t stream, void** args, bool pdl = false, int cluster = kConvCluster)
{
    // (one attribute call per kernel, device and size class instead of one per launch: a step is launched every few us)
    static thread_local const void* set_for = nullptr; static thread_local size_t set_smem = 0; static thread_local int set_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (set_for != (const void*)kernel || set_smem < smem || set_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        set_for = (const void*)kernel; set_smem = smem; set_dev = dev;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kConvStepThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
    return cudaLaunchKernelExC(&cfg, (const void*)kernel, args);
}

// Stereo mix: the per-source outputs of n_blocks steps summed over the sources in source order (deterministic;
// an in-kernel RED.ADD from the 16 clusters of a step would make the sum depend on their timing).
__global__ void __launch_bounds__(256) mix_kernel(const float* __restrict__ out, int n_src, int per_block /* 2 * block */, long long n,
                                                  const float* __restrict__ gain, float* __restrict__ mix)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [n_blocks][2][block]
    if (i >= n) return;
    const long long b = i / per_block, t = i % per_block;
    const float* src = out + (b * n_src) * per_block + t;
    float acc = 0.f;
    for (int s = 0; s < n_src; ++s) acc = fmaf(gain ? gain[s] : 1.f, src[(long long)s * per_block], acc);
    mix[i] = acc;
}

// Host staging <-> device buffers by the SMs instead of the copy engines (a 32-64 KB cudaMemcpyAsync costs ~6 us of
// latency each way, a one-CTA kernel over mapped pinned memory ~2 us), with the completion word written by the same
// kernel once its stores are out.
__global__ void __launch_bounds__(1024) stage_kernel(const float4* __restrict__ src0, float4* __restrict__ dst0, long long n0,
                                                     const float4* __restrict__ src1, float4* __restrict__ dst1, long long n1,
                                                     volatile unsigned* flag, unsigned value, unsigned* done)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < n0; i += nt) dst0[i] = src0[i];
    for (long long i = tid; i < n1; i += nt) dst1[i] = src1[i];
    if (flag) {
        // the CTA that arrives last publishes the completion word, after every CTA's stores went out at system scope
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned prev = gridDim.x > 1 ? atomicAdd(done, 1u) : 0u;
            if (prev == gridDim.x - 1) {
                if (gridDim.x > 1) *done = 0u;
                __threadfence_system();
                *flag = value;
            }
        }
    }
}

// Stream-ordered completion mark in mapped host memory: the host-buffer entry points spin on it instead of paying a
// cudaStreamSynchronize per 512-sample block.
__global__ void flag_kernel(volatile unsigned* flag, unsigned value) { __threadfence_system(); *flag = value; }

} // namespace

cudaError_t conv_mix(const float* d_out, int n_src, int block, int n_blocks, const float* d_gain, float* d_mix, cudaStream_t stream)
{
    const long long n = (long long)n_blocks * 2 * block;
    if (n <= 0) return cudaSuccess;
    mix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_out, n_src, 2 * block, n, d_gain, d_mix);
    return cudaGetLastError();
}

cudaError_t conv_stage(const float* src0, float* dst0, long long n0, const float* src1, float* dst1, long long n1,
                       unsigned* mapped_flag, unsigned value, unsigned* d_done, cudaStream_t stream)
{
    // one CTA moves ~14 GB/s over PCIe: 16 KB per CTA, at most 16 CTAs
    const long long bytes = (n0 + n1) * 4;
    unsigned grid = (unsigned)((bytes + 16383) / 16384);
    grid = grid < 1u ? 1u : (grid > 16u ? 16u : grid);
    if (!d_done) grid = 1u;
    stage_kernel<<<grid, 1024, 0, stream>>>((const float4*)src0, (float4*)dst0, n0 / 4, (const float4*)src1, (float4*)dst1, n1 / 4, mapped_flag, value, d_done);
    return cudaGetLastError();
}

cudaError_t conv_signal(unsigned* mapped_flag, unsigned value, cudaStream_t stream)
{
    flag_kernel<<<1, 1, 0, stream>>>(mapped_flag, value);
    return cudaGetLastError();
}

cudaError_t conv_upload_twiddles(float2* d_tw, int N, cudaStream_t stream)
{
    std::vector<float2> tw((size_t)N);
    for (int k = 0; k < N; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)N;
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    cudaError_t e = cudaMemcpyAsync(d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);   // tw is a host temporary
}

cudaError_t conv_ir_spectra(const float* d_h, int n_items, int ear0, int ir_len, int block, int P, const float2* d_tw, float2* d_H,
                            cudaStream_t stream)
{
    const size_t smem = fft_smem_bytes(block);
    cudaError_t e = cudaFuncSetAttribute(ir_spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ir_spectra_kernel<<<(unsigned)(n_items * P), kConvThreads, smem, stream>>>(d_h, ir_len, block, P, ear0, d_tw, d_H);
    return cudaGetLastError();
}

cudaError_t conv_block_spectra(const float* d_x, long long n, long long seg_len, int n_seg, int blocks_per_seg, int block,
                               const float2* d_tw, float2* d_X, cudaStream_t stream)
{
    const size_t smem = fft_smem_bytes(block);
    cudaError_t e = cudaFuncSetAttribute(block_spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)(n_seg * blocks_per_seg);
    if (grid == 0) return cudaSuccess;
    block_spectra_kernel<<<grid, kConvThreads, smem, stream>>>(d_x, n, seg_len, blocks_per_seg, block, d_tw, d_X);
    return cudaGetLastError();
}

cudaError_t conv_debug_trace(void* out, size_t bytes)
{
#ifdef ARV2_CONV_TRACE
    return cudaMemcpyFromSymbol(out, g_ct_trace, bytes < sizeof(g_ct_trace) ? bytes : sizeof(g_ct_trace));
#else
    (void)out; (void)bytes;
    return cudaErrorNotSupported;
#endif
}

int conv_ring_stages(int block, bool deep) { return deep ? ring_stages_deep(block) : ring_stages(block); }

template <int BPT>
cudaError_t step16_prepare()
{
    static thread_local int done_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (done_dev == dev) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute((const void*)stream_step_kernel<BPT, 16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) done_dev = dev;
    return e;
}

// Can clusters of 16 CTAs of the step kernel be launched here, n_src of them at a time (twice that with a successor in)?
bool conv_cluster16_ok(int n_src, int block)
{
    const int bpt = (block + kConvThreads - 1) / kConvThreads;
    const void* k = bpt == 1 ? (const void*)stream_step_kernel<1, 16> : bpt == 2 ? (const void*)stream_step_kernel<2, 16> : (const void*)stream_step_kernel<4, 16>;
    const size_t smem = step_smem_bytes(block, 0);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)n_src * 16); cfg.blockDim = dim3(kConvStepThreads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
    return n >= 2 * n_src;
}

cudaError_t conv_stream_step(const ConvStreamArgs& a, cudaStream_t stream)
{
    ConvStreamArgs args = a;
    void* kargs[] = {&args};
    const size_t smem = step_smem_bytes(a.block, a.stages);
    const int bpt = (a.block + kConvThreads - 1) / kConvThreads;
    static const bool pdl = getenv("ARV2_CONV_NO_PDL") == nullptr;     // A/B switch
    if (a.cluster == 16) {
        const unsigned grid16 = (unsigned)(a.n_src * 16);
        cudaError_t e = bpt == 1 ? step16_prepare<1>() : bpt == 2 ? step16_prepare<2>() : step16_prepare<4>();
        if (e != cudaSuccess) return e;
        switch (bpt) {
        case 1: return launch_cluster(stream_step_kernel<1, 16>, grid16, smem, stream, kargs, pdl, 16);
        case 2: return launch_cluster(stream_step_kernel<2, 16>, grid16, smem, stream, kargs, pdl, 16);
        case 4: return launch_cluster(stream_step_kernel<4, 16>, grid16, smem, stream, kargs, pdl, 16);
        default: return cudaErrorInvalidValue;
        }
    }
    const unsigned grid = (unsigned)(a.n_src * kConvCluster);
    switch (bpt) {
    case 1: return launch_cluster(stream_step_kernel<1>, grid, smem, stream, kargs, pdl);
    case 2: return launch_cluster(stream_step_kernel<2>, grid, smem, stream, kargs, pdl);
    case 4: return launch_cluster(stream_step_kernel<4>, grid, smem, stream, kargs, pdl);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t conv_stream_blocks(const ConvStreamArgs& a, cudaStream_t stream)
{
    if (a.n_blocks <= 0) return cudaSuccess;
    ConvStreamArgs args = a;
    void* kargs[] = {&args};
    const unsigned grid = (unsigned)(a.n_src * kConvCluster);
    const size_t smem = step_smem_bytes(a.block, 0);
    const int bpt = (a.block + kConvThreads - 1) / kConvThreads;
    switch (bpt) {
    case 1: return launch_cluster(stream_blocks_kernel<1>, grid, smem, stream, kargs);
    case 2: return launch_cluster(stream_blocks_kernel<2>, grid, smem, stream, kargs);
    case 4: return launch_cluster(stream_blocks_kernel<4>, grid, smem, stream, kargs);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t conv_file(const ConvFileArgs& a, cudaStream_t stream)
{
    ConvFileArgs args = a;
    int out_blocks = a.blocks_per_seg + (a.wrap > 0 ? a.P - 1 : 0);
    void* kargs[] = {&args, &out_blocks};
    const long long clusters = (long long)a.n_seg * out_blocks;
    if (clusters == 0) return cudaSuccess;
    const unsigned grid = (unsigned)(clusters * kConvCluster);
    const size_t smem = step_smem_bytes(a.block, 0);
    const int bpt = (a.block + kConvThreads - 1) / kConvThreads;
    switch (bpt) {
    case 1: return launch_cluster(file_kernel<1>, grid, smem, stream, kargs);
    case 2: return launch_cluster(file_kernel<2>, grid, smem, stream, kargs);
    case 4: return launch_cluster(file_kernel<4>, grid, smem, stream, kargs);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace arv2

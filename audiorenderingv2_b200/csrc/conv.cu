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
// each CTA accumulates 1/8 of the partitions, partial sums are reduce-scattered through
// distributed shared memory, and the cluster's rank 0 runs the inverse FFT + overlap-add.
#include <cmath>
#include <cstdlib>
#include <vector>

#include <cooperative_groups.h>

#include "conv.cuh"

namespace cg = cooperative_groups;

namespace arv2 {

namespace {

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
            float2 w1 = __ldg(tw + p * s), w2 = __ldg(tw + 2 * p * s), w3 = __ldg(tw + 3 * p * s);
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

__global__ void __launch_bounds__(kConvThreads) ir_spectra_kernel(const float* __restrict__ h, int ir_len, int block, int P,
                                                                  const float2* __restrict__ tw, float2* __restrict__ H)
{
    extern __shared__ float2 smem[];
    float2* bufa = smem; float2* bufb = smem + 2 * block;
    const int item = blockIdx.x / P, p = blockIdx.x % P;
    const float* src = h + (size_t)item * ir_len;
    float2* dst = H + ((size_t)item * P + p) * block;
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
// of a partition (X, H_left, H_right: block*8 B each) are instead pulled into a 4-stage
// shared-memory ring by bulk async copies that complete on an mbarrier; the threads only
// read shared memory.
// 4 stages x 12 KB in flight per SM already stream the spectra at the L2 rate (128 SMs x ~60 GB/s); 8 and 12
// stages measured 26 us per step instead of 16.4 (profiles/micro/conv_ab.py)
#ifndef ARV2_CONV_STAGES
#define ARV2_CONV_STAGES 4
#endif
#ifndef ARV2_CONV_FFTCOST
#define ARV2_CONV_FFTCOST 8
#endif
constexpr int kStages = ARV2_CONV_STAGES;
constexpr int kFftCostInPartitions = ARV2_CONV_FFTCOST;   // forward FFT of one block ~ streaming 8 partitions (clock64-measured)

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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Programmatic dependent launch (sm_90+): step k+1 may become resident while step k is still running; everything
// that depends on step k comes after pdl_wait(), which returns once step k has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// acc += X[k] * H_e[k] for this thread's bins k = tid*BPT + i (contiguous, so that a
// thread's BPT bins are one 8*BPT-byte shared-memory access); bin 0 holds (DC, Nyquist).
template <int BPT>
__device__ __forceinline__ void mac_rows(const float2* X, const float2* HL, const float2* HR, int block, float2 accL[BPT], float2 accR[BPT])
{
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const int k = threadIdx.x * BPT + i;
        if (k < block) {
            const float2 xv = X[k], hl = HL[k], hr = HR[k];
            if (k == 0) {   // packed (DC, Nyquist): two real products
                accL[i].x = fmaf(xv.x, hl.x, accL[i].x); accL[i].y = fmaf(xv.y, hl.y, accL[i].y);
                accR[i].x = fmaf(xv.x, hr.x, accR[i].x); accR[i].y = fmaf(xv.y, hr.y, accR[i].y);
            } else {
                accL[i].x = fmaf(xv.x, hl.x, fmaf(-xv.y, hl.y, accL[i].x)); accL[i].y = fmaf(xv.x, hl.y, fmaf(xv.y, hl.x, accL[i].y));
                accR[i].x = fmaf(xv.x, hr.x, fmaf(-xv.y, hr.y, accR[i].x)); accR[i].y = fmaf(xv.x, hr.y, fmaf(xv.y, hr.x, accR[i].y));
            }
        }
    }
}

// Multiply-accumulate `n` partitions through the TMA ring.  rows(i, &X, &HL, &HR) yields the
// global rows of this CTA's i-th partition.  ring: float2[kStages][3][block]; all threads call.
template <int BPT, class Rows>
__device__ __forceinline__ void mac_pipeline(float2* ring, unsigned long long* full, int n, int block, Rows rows,
                                             float2 accL[BPT], float2 accR[BPT])
{
    const unsigned row_bytes = (unsigned)block * sizeof(float2);
    auto issue = [&](int i, int s) {
        const float2 *X, *HL, *HR;
        rows(i, &X, &HL, &HR);
        float2* dst = ring + (size_t)s * 3 * block;
        mbar_expect_tx(&full[s], 3 * row_bytes);
        bulk_g2s(dst, X, row_bytes, &full[s]);
        bulk_g2s(dst + block, HL, row_bytes, &full[s]);
        bulk_g2s(dst + 2 * block, HR, row_bytes, &full[s]);
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < kStages && s < n; ++s) issue(s, s);
    for (int i = 0; i < n; ++i) {
        const int s = i % kStages;
        mbar_wait(&full[s], (unsigned)((i / kStages) & 1));
        const float2* src = ring + (size_t)s * 3 * block;
        mac_rows<BPT>(src, src + block, src + 2 * block, block, accL, accR);
        __syncthreads();                                   // everyone is done with stage s
        if (threadIdx.x == 0 && i + kStages < n) issue(i + kStages, s);
    }
}

// Cluster-wide reduce-scatter of the per-CTA partial sums into rank 0's `yacc`, then
// rank 0 unpacks Z = YL + i*YR, inverse-transforms and returns the time-domain buffer
// (re = left, im = right, unnormalised); other ranks return nullptr.
template <int BPT>
__device__ float2* reduce_and_inverse(cg::cluster_group& cluster, float2* part, float2* yacc, float2* bufa, float2* bufb,
                                      int block, const float2* tw, const float2 accL[BPT], const float2 accR[BPT])
{
    const unsigned rank = cluster.block_rank();
    const unsigned C = cluster.num_blocks();
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const int k = threadIdx.x * BPT + i;
        if (k < block) { part[k] = accL[i]; part[block + k] = accR[i]; }
    }
    cluster.sync();
    {
        // rank r owns elements [r*chunk, (r+1)*chunk) of the 2*block partial sums
        const int total = 2 * block;
        const int chunk = (total + (int)C - 1) / (int)C;
        float2* dst = cluster.map_shared_rank(yacc, 0);
        for (int e = threadIdx.x; e < chunk; e += blockDim.x) {
            const int idx = (int)rank * chunk + e;
            if (idx < total) {
                float2 s = make_float2(0.f, 0.f);
                for (unsigned r = 0; r < C; ++r) {
                    const float2 v = cluster.map_shared_rank(part, r)[idx];
                    s.x += v.x; s.y += v.y;
                }
                dst[idx] = s;
            }
        }
    }
    cluster.sync();
    if (rank != 0) return nullptr;
    const int N = 2 * block;
    for (int k = threadIdx.x; k < block; k += blockDim.x) {
        const float2 l = yacc[k], r = yacc[block + k];
        if (k == 0) {
            bufa[0] = make_float2(l.x, r.x);
            bufa[block] = make_float2(l.y, r.y);
        } else {
            bufa[k] = make_float2(l.x - r.y, l.y + r.x);
            bufa[N - k] = make_float2(l.x + r.y, r.x - l.y);
        }
    }
    __syncthreads();
    return fft_smem(bufa, bufb, N, tw, true);
}

template <int BPT>
__global__ void __launch_bounds__(kConvThreads) stream_step_kernel(const ConvStreamArgs a)
{
    extern __shared__ __align__(128) float2 smem[];
    __shared__ unsigned long long full[kStages];
    const int block = a.block, N = 2 * block;
    float2* bufa = smem; float2* bufb = bufa + N; float2* part = bufb + N; float2* yacc = part + N; float2* ring = yacc + N;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const unsigned C = cluster.num_blocks();
    const int src = blockIdx.x / C;
    float2* fdl = a.fdl + (size_t)src * a.P * block;
    const float2* H = a.H[src];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

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
    int n0 = C > 1 ? (T - kFftCostInPartitions * ((int)C - 1)) / (int)C : T;
    n0 = max(0, min(T, n0));
    int first, n;
    if (rank == 0) { first = 2; n = n0; }
    else {
        const int rest = T - n0, per = rest / ((int)C - 1), extra = rest % ((int)C - 1), r1 = (int)rank - 1;
        first = 2 + n0 + r1 * per + min(r1, extra);
        n = per + (r1 < extra ? 1 : 0);
    }
    mac_pipeline<BPT>(ring, full, n, block,
                      [&](int i, const float2** X, const float2** HL, const float2** HR) {
                          const int p = first + i;
                          int s = a.slot - p; if (s < 0) s += a.P;
                          *X = fdl + (size_t)s * block; *HL = H + (size_t)p * block; *HR = H + ((size_t)a.P + p) * block;
                      },
                      accL, accR);
    pdl_wait();
    pdl_launch_dependents();
    if (rank == 0) {
        // newest block: forward FFT, publish into the frequency-domain delay line
        const float* in = a.in + (size_t)src * block;
        float2* slot = fdl + (size_t)a.slot * block;
        forward_block(bufa, bufb, block, a.tw, [&](int t) { return in[t]; }, [&](int k, float2 v) { slot[k] = v; });
        mac_rows<BPT>(slot, H, H + (size_t)a.P * block, block, accL, accR);
    }
    if (rank == C - 1 && a.P > 1) {
        // the previous block: its spectrum was published by the step this one waited for
        const int s1 = a.slot >= 1 ? a.slot - 1 : a.slot - 1 + a.P;
        mac_rows<BPT>(fdl + (size_t)s1 * block, H + (size_t)block, H + ((size_t)a.P + 1) * block, block, accL, accR);
    }
    float2* y = reduce_and_inverse<BPT>(cluster, part, yacc, bufa, bufb, block, a.tw, accL, accR);
    if (!y) return;
    const float sc = 1.0f / (float)N;
    float* out = a.out + (size_t)src * 2 * block;
    float* tail = a.tail + (size_t)src * 2 * block;
    for (int t = threadIdx.x; t < block; t += blockDim.x) {
        const float2 lo = y[t], hi = y[block + t];
        out[t] = fmaf(lo.x, sc, tail[t]);
        out[block + t] = fmaf(lo.y, sc, tail[block + t]);
        tail[t] = hi.x * sc;
        tail[block + t] = hi.y * sc;
    }
}

template <int BPT>
__global__ void __launch_bounds__(kConvThreads) file_kernel(const ConvFileArgs a, int out_blocks_per_seg)
{
    extern __shared__ __align__(128) float2 smem[];
    __shared__ unsigned long long full[kStages];
    const int block = a.block, N = 2 * block;
    float2* bufa = smem; float2* bufb = bufa + N; float2* part = bufb + N; float2* yacc = part + N; float2* ring = yacc + N;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const unsigned C = cluster.num_blocks();
    const int cid = blockIdx.x / C;
    const int seg = cid / out_blocks_per_seg, j = cid % out_blocks_per_seg;
    const float2* X = a.X + (size_t)seg * a.blocks_per_seg * block;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    float2 accL[BPT], accR[BPT];
#pragma unroll
    for (int i = 0; i < BPT; ++i) { accL[i] = make_float2(0.f, 0.f); accR[i] = make_float2(0.f, 0.f); }
    const int p_lo = max(0, j - (a.blocks_per_seg - 1)), p_hi = min(a.P - 1, j);
    const int first = p_lo + (int)rank;
    const int n = first <= p_hi ? (p_hi - first) / (int)C + 1 : 0;
    mac_pipeline<BPT>(ring, full, n, block,
                      [&](int i, const float2** Xr, const float2** HL, const float2** HR) {
                          const int p = first + i * (int)C;
                          *Xr = X + (size_t)(j - p) * block; *HL = a.H + (size_t)p * block; *HR = a.H + ((size_t)a.P + p) * block;
                      },
                      accL, accR);
    float2* y = reduce_and_inverse<BPT>(cluster, part, yacc, bufa, bufb, block, a.tw, accL, accR);
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
size_t step_smem_bytes(int block) { return (size_t)(8 + 3 * kStages) * block * sizeof(float2); }

template <class K>
cudaError_t launch_cluster(K kernel, unsigned grid, size_t smem, cudaStream_t stream, void** args, bool pdl = false)
{
    cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kConvCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
    return cudaLaunchKernelExC(&cfg, (const void*)kernel, args);
}

} // namespace

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

cudaError_t conv_ir_spectra(const float* d_h, int n_items, int ir_len, int block, int P, const float2* d_tw, float2* d_H,
                            cudaStream_t stream)
{
    const size_t smem = fft_smem_bytes(block);
    cudaError_t e = cudaFuncSetAttribute(ir_spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ir_spectra_kernel<<<(unsigned)(n_items * P), kConvThreads, smem, stream>>>(d_h, ir_len, block, P, d_tw, d_H);
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

cudaError_t conv_stream_step(const ConvStreamArgs& a, cudaStream_t stream)
{
    ConvStreamArgs args = a;
    void* kargs[] = {&args};
    const unsigned grid = (unsigned)(a.n_src * kConvCluster);
    const size_t smem = step_smem_bytes(a.block);
    const int bpt = (a.block + kConvThreads - 1) / kConvThreads;
    static const bool pdl = getenv("ARV2_CONV_NO_PDL") == nullptr;     // A/B switch
    switch (bpt) {
    case 1: return launch_cluster(stream_step_kernel<1>, grid, smem, stream, kargs, pdl);
    case 2: return launch_cluster(stream_step_kernel<2>, grid, smem, stream, kargs, pdl);
    case 4: return launch_cluster(stream_step_kernel<4>, grid, smem, stream, kargs, pdl);
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
    const size_t smem = step_smem_bytes(a.block);
    const int bpt = (a.block + kConvThreads - 1) / kConvThreads;
    switch (bpt) {
    case 1: return launch_cluster(file_kernel<1>, grid, smem, stream, kargs);
    case 2: return launch_cluster(file_kernel<2>, grid, smem, stream, kargs);
    case 4: return launch_cluster(file_kernel<4>, grid, smem, stream, kargs);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace arv2

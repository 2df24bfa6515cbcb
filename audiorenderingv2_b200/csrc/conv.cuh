// conv.cuh -- launch interface of the sm_100a partitioned FFT convolver (conv.cu).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace arv2 {

// Spectra are stored "packed": N = 2*block real samples -> block complex values, bin 0
// holding (DC, Nyquist) in (re, im) -- both are purely real for a real signal -- so a
// spectrum is exactly `block` float2 (4 KB at block = 512) and rows stay 16 B aligned.

constexpr int kConvThreads = 256;
constexpr int kConvCluster = 8;      // CTAs cooperating on one (source, block) via DSMEM

// Twiddle table W_N^k = exp(-2*pi*i*k/N), k < N, computed in fp64 on the host.
cudaError_t conv_upload_twiddles(float2* d_tw, int N, cudaStream_t stream);

// K10: IR -> partition spectra.  h: [n_items][ir_len] (device); H: [P][2 ears][block] (the two ears of a partition
// are adjacent: one bulk copy per partition); item i goes to ear ear0 + i.
cudaError_t conv_ir_spectra(const float* d_h, int n_items, int ear0, int ir_len, int block, int P, const float2* d_tw,
                            float2* d_H, cudaStream_t stream);

// Forward spectra of consecutive zero-padded input blocks (file mode).
// x: [n] samples; block j of segment s covers x[s*seg_len + j*block ...) clipped to the
// segment and to n.  X: [n_seg*blocks_per_seg][block].
cudaError_t conv_block_spectra(const float* d_x, long long n, long long seg_len, int n_seg, int blocks_per_seg,
                               int block, const float2* d_tw, float2* d_X, cudaStream_t stream);

struct ConvStreamArgs {
    const float* in;         // [n_src][block] newest input block per source
    float* out;              // [n_src][2][block]
    float2* fdl;             // [n_src][P][block] ring of input spectra
    const float2* const* H;  // [n_src] -> [P][2][block] active IR spectra of each source
    float* tail;             // [n_src][2][block] overlap-add tails
    const float2* tw;
    int n_src, block, P, slot; // slot = ring position of the newest block (of the first block for conv_stream_blocks)
    int n_blocks;              // conv_stream_blocks: consecutive blocks; in / out advance by n_src*block / n_src*2*block floats per block
    int stages;                // conv_stream_step: stages of the shared-memory ring (0 = conv_ring_stages(block, false))
    int cluster;               // conv_stream_step: CTAs per source, 8 (0 = 8) or 16 (conv_cluster16_ok)
    int early_input;           // conv_stream_step: `in` was complete before the previous step passed its wait (blocks 1.. of one
                               // call): the forward FFT of the newest block runs before this step waits for its predecessor
};
// One streaming step for all sources: forward FFT + FDL write + partitioned spectral MAC
// + DSMEM reduction + stereo inverse FFT + overlap-add, in ONE cluster launch.
cudaError_t conv_stream_step(const ConvStreamArgs& a, cudaStream_t stream);
// ring depth of a step: the default leaves room for two CTAs per SM (a step and its successor share the SMs when a step
// fills the machine), the deep one takes a whole SM (twice the bytes in flight: streams with few sources)
int conv_ring_stages(int block, bool deep);
// clusters of 16 CTAs per source for streams of few sources: launchable here, 2 * n_src of them at a time?
bool conv_cluster16_ok(int n_src, int block);
// builds with -DARV2_CONV_TIMING -DARV2_CONV_TRACE: the stamps of the last steps (conv.cu: g_ct_trace)
cudaError_t conv_debug_trace(void* out, size_t bytes);
// a.n_blocks consecutive steps in ONE cluster launch (every source's cluster loops over the blocks); same results.
cudaError_t conv_stream_blocks(const ConvStreamArgs& a, cudaStream_t stream);

// mix[b][ear][t] = sum over sources s (in order) of gain[s] * out[b][s][ear][t]; gain may be null (= 1).
cudaError_t conv_mix(const float* d_out, int n_src, int block, int n_blocks, const float* d_gain, float* d_mix, cudaStream_t stream);
// Copy (1..16 CTAs) of up to two float arrays (lengths multiples of 4, 16 B aligned; either may be empty) between mapped
// pinned host memory and device memory, then -- if mapped_flag is given -- the completion word.
cudaError_t conv_stage(const float* src0, float* dst0, long long n0, const float* src1, float* dst1, long long n1,
                       unsigned* mapped_flag, unsigned value, unsigned* d_done /* zeroed device word, for > 1 CTA */, cudaStream_t stream);
// *mapped_flag = value once everything enqueued on `stream` before this has completed (mapped, pinned host memory).
cudaError_t conv_signal(unsigned* mapped_flag, unsigned value, cudaStream_t stream);

struct ConvFileArgs {
    const float2* X;         // [n_seg*blocks_per_seg][block]
    const float2* H;         // [P][2][block]
    float* out_l; float* out_r; // [n] pre-zeroed, accumulated with RED.ADD.F32
    const float2* tw;
    long long n;             // output length
    long long seg_len;       // samples between segment starts (== n for one segment)
    long long wrap;          // circular length (reference mode: ir_len); <=0: linear
    long long seg_out;       // samples of a segment's result that are kept (reference: ir_len)
    int n_seg, blocks_per_seg, block, P;
    float gain;
};
cudaError_t conv_file(const ConvFileArgs& a, cudaStream_t stream);

} // namespace arv2

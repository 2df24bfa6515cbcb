/*
 * oracle_conv.cpp -- CPU ORACLE (test infrastructure, not the product).
 * PARITY UNPINNED for the convolvers (cuFFT is closed; see arv2_oracle.h "PIN STATUS").  Restates:
 *   file mode  OR/kernels.cu:382-438 + OR/AudioRenderer.cpp:663-711
 *   live mode  OR/kernels.cu:345-377 + OR/AudioRenderer.cpp:593-661
 * cuFFT (third-party, CUDA 12.1, absent from /root/reference) is an
 * unnormalised DFT, so IFFT(FFT(a).FFT(b)) of size N is N x the circular
 * convolution; the oracle evaluates that circular convolution directly in fp64.
 * Also holds the fp64 direct linear convolution the north star grades against
 * and a scalar fp32 uniformly-partitioned overlap-add port (CPU baseline).
 */
#include "arv2_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>
#include <thread>
#include <vector>

namespace {

template <class F>
void parallel_for(int64_t n, int n_threads, F f)
{
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n));
    if (nt == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([=] { f(n * t / nt, n * (t + 1) / nt); });
    for (auto& t : th) t.join();
}

/* circular convolution of a[0..na) (zero beyond) with h[0..N), size N, fp64 */
void circular(const double* a, int64_t na, const float* h, int64_t N, double* c, int n_threads)
{
    parallel_for(N, n_threads, [=](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
            double acc = 0.0;
            /* j <= i : h[i-j] ; j > i : h[i-j+N] */
            const int64_t j1 = std::min<int64_t>(na, i + 1);
            for (int64_t j = 0; j < j1; ++j) acc += a[j] * (double)h[i - j];
            for (int64_t j = j1; j < na; ++j) acc += a[j] * (double)h[i - j + N];
            c[i] = acc;
        }
    });
}

/* in-place radix-2 DIT complex FFT, float, sign = -1 forward / +1 inverse (unnormalised) */
void fft_r2(std::complex<float>* a, int n, int sign, const std::complex<float>* tw)
{
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                std::complex<float> w = tw[k * step];
                if (sign > 0) w = std::conj(w);
                const std::complex<float> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

} // namespace

extern "C" {

void oracle_direct_conv(const float* x, int64_t n, const float* h, int64_t m, double* y, int32_t n_threads)
{
    const int64_t L = n + m - 1;
    if (n <= 0 || m <= 0) return;
    parallel_for(L, n_threads, [=](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
            const int64_t j0 = std::max<int64_t>(0, i - (n - 1)), j1 = std::min<int64_t>(m - 1, i);
            double acc = 0.0;
            for (int64_t j = j0; j <= j1; ++j) acc += (double)h[j] * (double)x[i - j];
            y[i] = acc;
        }
    });
}

/* outputs [begin, begin+count) of the same fp64 direct convolution (full-size checks: a few blocks of a long stream) */
void oracle_direct_conv_window(const float* x, int64_t n, const float* h, int64_t m, int64_t begin, int64_t count,
                               double* y, int32_t n_threads)
{
    const int64_t L = n + m - 1;
    if (n <= 0 || m <= 0 || begin < 0 || count <= 0 || begin + count > L) return;
    parallel_for(count, n_threads, [=](int64_t b, int64_t e) {
        for (int64_t k = b; k < e; ++k) {
            const int64_t i = begin + k;
            const int64_t j0 = std::max<int64_t>(0, i - (n - 1)), j1 = std::min<int64_t>(m - 1, i);
            double acc = 0.0;
            for (int64_t j = j0; j <= j1; ++j) acc += (double)h[j] * (double)x[i - j];
            y[k] = acc;
        }
    });
}

void oracle_reference_file_conv(const float* x, int64_t n, const float* h, int32_t ir_len, int32_t sample_rate,
                                double* y, int32_t n_threads)
{
    const int64_t N = ir_len, fs = sample_rate;
    for (int64_t i = 0; i < n; ++i) y[i] = 0.0;
    const int64_t seconds = n / fs;                       /* kernels.cu:410 */
    const int64_t seg_samples = std::min<int64_t>(fs, fs * (N / fs)); /* load_sample_segment bound :232-236 */
    std::vector<double> seg(std::max<int64_t>(seg_samples, 1)), c(N);
    for (int64_t s = 0; s < seconds; ++s) {
        for (int64_t i = 0; i < seg_samples; ++i) seg[i] = (double)x[s * fs + i];
        circular(seg.data(), std::min<int64_t>(seg_samples, N), h, N, c.data(), n_threads);
        const int64_t how = (s * fs + N < n) ? N : n - s * fs; /* :425 */
        for (int64_t i = 0; i < how; ++i) y[s * fs + i] += c[i] * (double)N; /* unnormalised cuFFT round trip */
    }
    const double div = (double)(ir_len / 2);              /* AudioRenderer.cpp:709 (integer /2) */
    for (int64_t i = 0; i < n; ++i) y[i] /= div;
}

void oracle_reference_live_conv(const double* x, int64_t n_in, const float* ir_left, const float* ir_right,
                                int32_t ir_len, double* out)
{
    const int64_t N = ir_len;
    const int64_t na = std::min<int64_t>(n_in, N);
    std::vector<double> c(N);
    const double g = (double)N / (double)(ir_len / 2);    /* AudioRenderer.cpp:641 */
    circular(x, na, ir_left, N, c.data(), 1);
    for (int64_t i = 0; i < N; ++i) out[2 * i] = c[i] * g;       /* d_zipArrays kernels.cu:469-479 */
    circular(x, na, ir_right, N, c.data(), 1);
    for (int64_t i = 0; i < N; ++i) out[2 * i + 1] = c[i] * g;
}

double oracle_upola(const float* x, int64_t n_blocks, int32_t block, const float* h_left, const float* h_right,
                    int32_t ir_len, float* out_left, float* out_right)
{
    typedef std::complex<float> cf;
    const int B = block, N = 2 * B, K = B + 1;
    const int P = (ir_len + B - 1) / B;
    std::vector<cf> tw(N / 2);
    for (int k = 0; k < N / 2; ++k) {
        const double a = -2.0 * M_PI * k / N;
        tw[k] = cf((float)cos(a), (float)sin(a));
    }
    std::vector<cf> H[2];
    const float* hs[2] = {h_left, h_right};
    std::vector<cf> buf(N);
    for (int e = 0; e < 2; ++e) {
        H[e].assign((size_t)P * K, cf(0, 0));
        for (int p = 0; p < P; ++p) {
            for (int i = 0; i < N; ++i) {
                const int64_t idx = (int64_t)p * B + i;
                buf[i] = (i < B && idx < ir_len) ? cf(hs[e][idx], 0.f) : cf(0.f, 0.f);
            }
            fft_r2(buf.data(), N, -1, tw.data());
            for (int k = 0; k < K; ++k) H[e][(size_t)p * K + k] = buf[k];
        }
    }
    std::vector<cf> fdl((size_t)P * K, cf(0, 0));
    std::vector<float> tail[2] = {std::vector<float>(B, 0.f), std::vector<float>(B, 0.f)};
    std::vector<cf> acc(K);
    float* outs[2] = {out_left, out_right};
    const auto t0 = std::chrono::steady_clock::now();
    for (int64_t nb = 0; nb < n_blocks; ++nb) {
        const int slot = (int)(nb % P);
        for (int i = 0; i < N; ++i) buf[i] = i < B ? cf(x[nb * B + i], 0.f) : cf(0.f, 0.f);
        fft_r2(buf.data(), N, -1, tw.data());
        for (int k = 0; k < K; ++k) fdl[(size_t)slot * K + k] = buf[k];
        for (int e = 0; e < 2; ++e) {
            for (int k = 0; k < K; ++k) acc[k] = cf(0, 0);
            for (int p = 0; p < P; ++p) {
                const int s = (slot - p + P) % P; /* spectrum of block nb-p (zero until written) */
                const cf* X = &fdl[(size_t)s * K];
                const cf* Hp = &H[e][(size_t)p * K];
                for (int k = 0; k < K; ++k) acc[k] += X[k] * Hp[k];
            }
            for (int k = 0; k < K; ++k) buf[k] = acc[k];
            for (int k = 1; k < B; ++k) buf[N - k] = std::conj(acc[k]);
            fft_r2(buf.data(), N, +1, tw.data());
            const float sc = 1.0f / (float)N;
            for (int i = 0; i < B; ++i) {
                outs[e][nb * B + i] = buf[i].real() * sc + tail[e][i];
                tail[e][i] = buf[B + i].real() * sc;
            }
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"

/*
 * arv2_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain C++ restatement of the hot path of sgrazi/AudioRenderingV2
 * (prebuild/obj_raytracer = "OR/"):
 *   - stochastic sound-ray tracing of a triangle scene into a stereo,
 *     time-binned impulse response     (OR/devicePrograms.cu:62-254)
 *   - convolution of the dry signal with that IR
 *                                      (OR/kernels.cu:345-438, OR/AudioRenderer.cpp:593-750)
 *
 * PIN STATUS.  Pinned against the reference's own sources, compiled where they lie by `make ref` (ref_scene_dump.cpp,
 * ref_device_shim.cpp, ref_stubs/; vectors in tests/golden/ref_*, checks in tests/test_pin_cpu.py):
 *   - loadOBJ + placeReceiver (OR/OptixModel.cpp:75-257): bit for bit;
 *   - __closesthit__radiance (OR/devicePrograms.cu:62-180) on 10^5 hand-made hits: integer outcomes equal, floats to
 *     fp32 rounding;
 *   - the whole per-ray loop (__raygen__renderFrame + programs, :186-254) on BASELINE config 1 (100k rays) and the
 *     closed box (20k rays, 50 bounces): every ray ends in the same ear and bin after the same number of segments.
 * PARITY UNPINNED for the arithmetic that lives inside closed third-party code absent from /root/reference -- OptiX
 * 7.7's ray/triangle test, cuRAND XORWOW seeded by clock64(), cuFFT (CUDA 12.1) -- restated from the published
 * contract (one recipe, used on both sides of the pin); the reference ships no tests or golden vectors for them.
 * The convolvers are anchored on the fp64 direct convolution and on OR/input.txt / OR/output_ir.txt
 * (tests/golden/README.md).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product never links it.
 */
#ifndef ARV2_ORACLE_H
#define ARV2_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-render parameters: the fields of LaunchParams (OR/LaunchParams.h:20-43)
 * plus the new-build extensions (seed, bands, diffuse) that reduce to the
 * reference when bands==1 and every scattering coefficient is 0. */
typedef struct {
    int32_t  size_x, size_y, size_z;  /* launch dims; N = x*y*z rays           */
    float    emitter[3];              /* LaunchParams::emitter_position        */
    float    sphere_center[3];        /* LaunchParams::sphere_center           */
    float    base_power;
    float    energy_thres;
    uint32_t max_bounces;
    float    hrtf_absorption_rate;
    int32_t  sample_rate;
    int32_t  is_mono;
    int32_t  ir_length;               /* samples per ear                        */
    int32_t  bands;                   /* 1 = reference                          */
    uint64_t seed;                    /* replaces clock64() (devicePrograms.cu:217) */
} oracle_params;

/* Flat scene: T triangles in reference mesh order (scene meshes in loadOBJ
 * order, then receiver_left, then receiver_right -- OR/OptixModel.cpp:153-257).
 *   tri_verts : float[T][3][3]  world-space P1,P2,P3
 *   tri_mat   : int32[T]        >=0 wall material index, -1 receiver_left,
 *                               -2 receiver_right (OR/AudioRenderer.cpp:37-45)
 *   absorption: float[M][bands] mat_absorption per band
 *   scattering: float[M]        probability of a Lambert bounce (0 = reference)
 * Outputs (any may be NULL):
 *   hist      : double[2][bands][ir_length] (0 = left, 1 = right), zeroed here
 *   rec_bin   : int32[n_rays]  bin of the receiver hit, -1 if the ray never hit
 *               the receiver (bin >= ir_length is recorded but not deposited)
 *   rec_ear   : int32[n_rays]  0 none, 1 receiver_left, 2 receiver_right
 *   rec_energy: float[n_rays][bands] remaining_factor at the hit
 *   rec_nseg  : int32[n_rays]  optixTrace-equivalent calls made by the ray
 * Rays [ray_begin, ray_begin+n_rays) of the N-ray set are traced.
 * use_bvh: 0 = brute force over all triangles, 1 = oracle's own BVH.
 * Returns the number of traced segments (closest-hit queries), <0 on error. */
int64_t oracle_trace(const oracle_params* p,
                     const float* tri_verts, const int32_t* tri_mat, int64_t n_tris,
                     const float* absorption, const float* scattering, int32_t n_mats,
                     int64_t ray_begin, int64_t n_rays,
                     int32_t use_bvh, int32_t n_threads,
                     double* hist, int32_t* rec_bin, int32_t* rec_ear,
                     float* rec_energy, int32_t* rec_nseg);

/* Same, with the scene (and the oracle's BVH) prepared once: create / trace / destroy. */
void* oracle_scene_create(const float* tri_verts, const int32_t* tri_mat, int64_t n_tris,
                          const float* absorption, const float* scattering, int32_t n_mats,
                          int32_t bands, int32_t use_bvh);
int64_t oracle_trace_scene(void* scene, const oracle_params* p, int64_t ray_begin, int64_t n_rays,
                           int32_t n_threads, double* hist, int32_t* rec_bin, int32_t* rec_ear,
                           float* rec_energy, int32_t* rec_nseg);
void oracle_scene_destroy(void* scene);

/* One closest-hit invocation (single band), flat form shared with oracle/ref_device_shim.cpp (pin test against the
 * reference's own __closesthit__radiance).  prd8 = {remaining_factor, distance, prev_position[3], direction[3]}
 * in/out; *depth in/out; deposits (<= 2): ear 0 = ir_left, 1 = ir_right. */
void oracle_shade_hit(const float* tri9, float mat_absorption, const float* ray_dir, float u, float v,
                      const float* sphere_center, int32_t sample_rate, float hrtf, int32_t is_mono, int32_t ir_length,
                      float* prd8, int32_t* depth, int32_t* n_dep, int32_t* dep_ear, int32_t* dep_idx, float* dep_val);

/* Direction of ray `ray_id` of the seeded set (unit vector, float[3]). */
void oracle_ray_direction(uint64_t seed, uint64_t ray_id, float* dir3);

/* Philox4x32-10 block for (seed, ray, bounce, purpose): uint32[4]. */
void oracle_philox(uint64_t seed, uint64_t ray_id, uint32_t bounce, uint32_t purpose,
                   uint32_t* out4);

/* Closest hit of one ray against the flat scene (brute force).  Returns the
 * triangle id or -1; t,u,v written if hit.  For unit tests of the intersector. */
int64_t oracle_closest_hit(const float* tri_verts, int64_t n_tris,
                           const float* org3, const float* dir3,
                           float* t, float* u, float* v);

/* IR finalisation of AudioRenderer::render (OR/AudioRenderer.cpp:520-523,
 * OR/kernels.cu:519-527): double hist[2][bands][ir_len] -> float L,R per band;
 * mono: L=R=L+R (float add). */
void oracle_finalize_ir(const double* hist, int32_t bands, int32_t ir_length,
                        int32_t is_mono, float* ir_left, float* ir_right);

/* ---- convolution ------------------------------------------------------- */

/* fp64 direct linear convolution: y[n+m-1] = x[n] * h[m]. */
void oracle_direct_conv(const float* x, int64_t n, const float* h, int64_t m,
                        double* y, int32_t n_threads);

/* Outputs [begin, begin+count) of the same convolution (y has count entries). */
void oracle_direct_conv_window(const float* x, int64_t n, const float* h, int64_t m,
                               int64_t begin, int64_t count, double* y, int32_t n_threads);

/* Reference file convolver, restated exactly (OR/kernels.cu:382-438 +
 * OR/AudioRenderer.cpp:702-711): whole seconds only, FFT size == ir_len
 * (circular), overlap-add at second*fs truncated to n, gain 1/(ir_len/2) after
 * cuFFT's unnormalised round trip (net 2x).  fp64 arithmetic. y has n entries. */
void oracle_reference_file_conv(const float* x, int64_t n, const float* h,
                                int32_t ir_len, int32_t sample_rate, double* y,
                                int32_t n_threads);

/* Reference live convolver for one callback (OR/kernels.cu:345-377 +
 * OR/AudioRenderer.cpp:593-661): zero-pad `x[n_in]` to ir_len, circular
 * convolution of size ir_len with each ear, /(ir_len/2), interleave LRLR.
 * out has 2*ir_len doubles. */
void oracle_reference_live_conv(const double* x, int64_t n_in, const float* ir_left,
                                const float* ir_right, int32_t ir_len, double* out);

/* Scalar uniformly-partitioned overlap-add convolver (CPU port used as the
 * cpu_baseline for "conv us per block"): float arithmetic, radix-2 FFT.
 * x: [n_blocks*block] mono input; h_left/h_right: [ir_len]; out: [2][n_blocks*block].
 * Returns seconds spent in the per-block loop (excludes IR partition FFTs). */
double oracle_upola(const float* x, int64_t n_blocks, int32_t block,
                    const float* h_left, const float* h_right, int32_t ir_len,
                    float* out_left, float* out_right);

#ifdef __cplusplus
}
#endif
#endif

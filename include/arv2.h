/*
 * arv2.h -- C ABI of libarv2.so, the B200-native (sm_100a) replacement for the hot
 * path of sgrazi/AudioRenderingV2: stochastic sound-ray tracing of a triangle scene
 * into a stereo, time-binned impulse response (IR), and FFT convolution of a dry
 * signal with that IR.
 *
 * The reference has no FFI layer; its boundary for this path is the C++ class
 * `AudioRenderer` (prebuild/obj_raytracer/AudioRenderer.h:16-152, "OR/" below) plus
 * the scene helpers in OR/OptixModel.{h,cpp} and the config loader OR/Context.cpp.
 * Every entry point below names the reference member it replaces.  A header-only
 * C++ mirror of the class over this ABI is csrc/host/audio_renderer.hpp; the
 * binding a maintainer would add to the reference is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success
 * or a negative arv2_status; no exceptions and no exit() cross the boundary
 * (the reference throws / exits: OR/optix7.h:8-45); arv2_last_error() returns the
 * message of the calling thread's last failure.  Sizes are SAMPLE / element counts,
 * never bytes (the reference mixes both: OR/AudioRenderer.cpp:671-683).
 * There is no CPU fallback: anything that needs the GPU fails with ARV2_ERR_CUDA
 * when no sm_100 device / driver is present.
 */
#ifndef ARV2_H
#define ARV2_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARV2_MAX_BANDS 8

typedef enum {
    ARV2_OK = 0,
    ARV2_ERR_INVALID = -1,  /* bad argument                                        */
    ARV2_ERR_IO = -2,       /* file unreadable / parse error (loadOBJ throws)      */
    ARV2_ERR_CUDA = -3,     /* CUDA runtime / no device (CUDA_CHECK, OR/optix7.h)  */
    ARV2_ERR_STATE = -4,    /* call order (e.g. re-render without a path cache)    */
    ARV2_ERR_NOMEM = -5
} arv2_status;

const char* arv2_last_error(void);
const char* arv2_version(void);

/* ------------------------------------------------------------------ scene -- */
/* struct OptixModel / TriangleMesh (OR/OptixModel.h:9-32): an ordered list of
 * meshes, one per (shape, material) of the OBJ, each with an MTL material name. */
typedef struct arv2_scene arv2_scene;

/* loadOBJ (OR/OptixModel.cpp:75-151).  Same triangulation, mesh split and order as
 * the vendored tinyobjloader; fails like the reference when the file is unreadable
 * or defines no material (OptixModel.cpp:93-100). */
int arv2_scene_load_obj(const char* obj_path, arv2_scene** out);
/* Build a scene from flat triangles: tri_verts float[n][3][3], tri_mesh int32[n]
 * (non-decreasing mesh index), one material name per mesh. */
int arv2_scene_from_triangles(const float* tri_verts, const int32_t* tri_mesh, int64_t n_tris,
                              const char* const* mesh_material_names, int32_t n_meshes,
                              arv2_scene** out);
int arv2_scene_counts(const arv2_scene* s, int64_t* n_tris, int32_t* n_meshes);
int arv2_scene_get_triangles(const arv2_scene* s, float* tri_verts, int32_t* tri_mesh);
const char* arv2_scene_mesh_material(const arv2_scene* s, int32_t mesh);
/* OptixModel::bounds (OR/OptixModel.cpp:143-147): lo[3], hi[3]. */
int arv2_scene_bounds(const arv2_scene* s, float* lo3, float* hi3);
/* Host-only diagnostic of the acceleration structure built for this scene (replaces what optixAccelBuild +
 * optixAccelCompact report, OR/AudioRenderer.cpp:152-207): builds the SAH BVH the renderer would upload and checks
 * it -- every triangle in exactly one leaf, every child box contains its triangles, children stored after their
 * parent.  sah_nodes / sah_tris = expected node visits / triangle tests of a random line (surface-area heuristic). */
typedef struct arv2_bvh_stats {
    int64_t n_tris, n_nodes, n_leaves;
    int32_t max_leaf_tris, depth, valid;
    double sah_nodes, sah_tris;
} arv2_bvh_stats;
int arv2_scene_bvh_stats(const arv2_scene* s, arv2_bvh_stats* out);
void arv2_scene_destroy(arv2_scene* s);

/* class HalfSphere / Sphere (OR/HalfSphere.cpp, OR/Sphere.cpp): the two half-ball
 * template meshes of the receiver. */
typedef struct arv2_receiver arv2_receiver;
int arv2_receiver_load(const char* left_obj, const char* right_obj, arv2_receiver** out);
int arv2_receiver_from_triangles(const float* left, int64_t n_left, const float* right,
                                 int64_t n_right, arv2_receiver** out);
int arv2_receiver_counts(const arv2_receiver* r, int64_t* n_left, int64_t* n_right);
/* placeReceiver / place_receiver_half (OR/OptixModel.cpp:153-257):
 * v' = cam + R_y(-rotation_deg) * v.  Host-only; writes float[n][3][3] per half. */
int arv2_receiver_place(const arv2_receiver* r, const float cam3[3], float rotation_deg,
                        float* left_out, float* right_out);
void arv2_receiver_destroy(arv2_receiver* r);

/* struct Material (OR/LaunchParams.h:14-18) + new-build extensions. */
typedef struct {
    const char* name;
    float mat_absorption[ARV2_MAX_BANDS]; /* [0] is the reference's scalar          */
    float scattering;                     /* probability of a Lambert bounce, 0=ref */
} arv2_material;

/* getMaterialAbsorption (OR/AudioRenderer.cpp:34-56): receiver_left -> -1,
 * receiver_right -> -2, by-name lookup, unknown -> 0.5. */
float arv2_material_absorption(const char* material_name, const arv2_material* materials,
                               int32_t n_materials);

/* ----------------------------------------------------------------- config -- */
/* Context::loadContext (OR/Context.cpp:15-165): config.json schema, the reference's
 * defaults and its rounding quirks (round() of ir_length_in_seconds, width, height,
 * re_render_*_threshold, ray_max_bounces and hrtf_absorption_rate when present). */
#define ARV2_MAX_CONFIG_MATERIALS 64
typedef struct {
    float    initial_volume;
    uint32_t ir_length_in_seconds;
    uint32_t width, height;
    int32_t  write_first_ir_to_file, write_first_output_to_file;
    float    re_render_distance_threshold, re_render_angle_threshold;
    int32_t  mono;
    char     scene_file_path[512];
    char     audio_file_path[512];
    char     materials_file_path[512];   /* parsed, never used (Context.cpp:85-87)  */
    float    initial_receiver_pos[3], initial_emitter_pos[3];
    float    base_power;
    float    rays[3];
    float    ray_energy_threshold;
    uint32_t ray_max_bounces;
    float    hrtf_absorption_rate;
    int32_t  n_materials;
    char     material_names[ARV2_MAX_CONFIG_MATERIALS][64];
    float    material_absorption[ARV2_MAX_CONFIG_MATERIALS];
    /* extensions (absent keys keep the reference behaviour) */
    uint64_t seed;                        /* pathtracer_parameters.seed, default 1   */
    int32_t  bands;                       /* pathtracer_parameters.bands, default 1  */
} arv2_config;
int arv2_config_parse(const char* json_text, arv2_config* out);
int arv2_config_load(const char* json_path, arv2_config* out);

/* --------------------------------------------------------------- renderer -- */
typedef struct arv2_ctx arv2_ctx;

typedef struct {
    uint32_t ir_length_in_seconds;  /* AudioRenderer ctor arg                        */
    int32_t  sample_rate;           /* AudioRenderer ctor arg                        */
    int32_t  rays_x, rays_y, rays_z;/* rays_per_dimension: N = x*y*z                 */
    int32_t  bands;                 /* 1 = reference, or ARV2_MAX_BANDS (8)          */
    int32_t  device;                /* CUDA ordinal (reference hard-wires 0)         */
    int32_t  record_rays;           /* keep per-ray (bin, ear, energy, nseg) records */
    int32_t  path_cache;            /* keep receiver-independent paths for rerender  */
    int32_t  bvh_builder;           /* 0 = host SAH, 1 = GPU LBVH                    */
    const arv2_material* materials; /* ctor arg std::vector<Material>                */
    int32_t  n_materials;
} arv2_renderer_desc;

/* AudioRenderer::AudioRenderer (OR/AudioRenderer.cpp:60-93): uploads the scene,
 * builds the acceleration structure, allocates and zeroes ir_left / ir_right.
 * The scene and receiver are copied; the caller may destroy them afterwards. */
int arv2_create(const arv2_scene* scene, const arv2_receiver* receiver,
                const arv2_renderer_desc* desc, arv2_ctx** out);
void arv2_destroy(arv2_ctx* ctx);

/* setEmitterPosInOptix (OR/AudioRenderer.cpp:752-756). */
int arv2_set_emitter(arv2_ctx* ctx, float x, float y, float z);
/* placeReceiver + setSphereCenterInOptix (OR/OptixModel.cpp:153-160,
 * OR/AudioRenderer.cpp:758-762): only the receiver's two-level sub-BVH is refit;
 * the scene BVH is never rebuilt (the reference rebuilds GAS+pipeline+SBT). */
int arv2_set_receiver(arv2_ctx* ctx, float x, float y, float z, float yaw_deg);
/* setThresholds / setBasePower / set_hrtf_absorption_rate / setMonoOutput
 * (OR/AudioRenderer.cpp:764-803).  Take effect at the next render. */
int arv2_set_thresholds(arv2_ctx* ctx, float energy, uint32_t max_bounces);
int arv2_set_base_power(arv2_ctx* ctx, float base_power);
int arv2_set_hrtf_absorption_rate(arv2_ctx* ctx, float rate);
int arv2_set_mono(arv2_ctx* ctx, int32_t mono);
/* replaces clock64() as the RNG seed (OR/devicePrograms.cu:217). */
int arv2_set_seed(arv2_ctx* ctx, uint64_t seed);
/* Start the rays of a launch in the order of their emission direction (default on).  The
 * ray set, every per-ray result and the IR are unchanged; only the lanes of a warp become
 * neighbours in direction.  The order is computed on the GPU once per (seed, ray range). */
int arv2_set_coherent_order(arv2_ctx* ctx, int32_t on);
/* Launches of at least min_rays rays are traced bounce-synchronously (sweep_kernel: all paths alive advance together, the
 * survivors are re-binned by origin cell and direction between two sweeps); smaller launches run in the per-SM queues of
 * wave_kernel.  Default 3 000 000 (measured crossover on B200: 2 M rays); min_rays <= 0 turns the sweeps off.  Scheduling
 * only: the ray set, every per-ray result and the IR are unchanged. */
int arv2_set_sweep_min_rays(arv2_ctx* ctx, int64_t min_rays);
/* Run everything on this CUDA stream (cudaStream_t as void*; NULL = own stream). */
int arv2_set_stream(arv2_ctx* ctx, void* cuda_stream);

/* AudioRenderer::render (OR/AudioRenderer.cpp:489-568): zero both IRs, trace the
 * full seeded ray set, finalise (mono merge L=R=L+R).  ms (may be NULL) receives
 * the device time of the trace in milliseconds. */
int arv2_render(arv2_ctx* ctx, double* ms);
/* Multi-GPU shard: trace rays [ray_begin, ray_begin+n_rays) of the same seeded set
 * into the fp64 histogram WITHOUT finalising; combine histograms across ranks
 * (arv2_hist_device + all-reduce) and then call arv2_finalize. */
int arv2_render_range(arv2_ctx* ctx, int64_t ray_begin, int64_t n_rays, int32_t zero_first,
                      double* ms);
int arv2_finalize(arv2_ctx* ctx);
/* ---- several GPUs: rays shard, histograms are summed with NCCL over NVLink (the reference
 * hard-wires device 0, OR/AudioRenderer.cpp:252-253).  libarv2 binds libnccl.so.2 at run
 * time: the copy already loaded in the process (a torch rank), else the system's. ---- */
#define ARV2_COMM_ID_BYTES 128
typedef struct arv2_comm arv2_comm;
/* ncclGetUniqueId: one rank makes the 128-byte id and hands it to the others (any transport). */
int arv2_comm_unique_id(void* id128);
/* ncclCommInitRank on `device`; collective over the n_ranks callers (one process or thread per GPU). */
int arv2_comm_create(int32_t device, int32_t rank, int32_t n_ranks, const void* id128, arv2_comm** out);
int arv2_comm_info(const arv2_comm* comm, int32_t* rank, int32_t* n_ranks, int32_t* nccl_version);
void arv2_comm_destroy(arv2_comm* comm);
/* ncclReduce(sum, float32) of `count` device floats onto `root`, in place (the stereo mix of sharded sources). */
int arv2_comm_reduce_f32(arv2_comm* comm, float* d_buf, size_t count, int32_t root, void* cuda_stream);
/* The contiguous slice [begin, begin+count) of an n_rays set that `rank` of n_ranks traces. */
void arv2_shard_range(int64_t n_rays, int32_t rank, int32_t n_ranks, int64_t* begin, int64_t* count);
/* Direction tiles, the default shards of arv2_render_sharded: the octahedral map of the emission directions is cut into
 * up to 16384 tiles and rank r takes the rays of the tiles t = r (mod n_ranks) -- every rank then holds rays as dense in direction
 * as the whole set, which keeps the warps' bundles as coherent as on one GPU (a contiguous slice of ray ids is 1/n_ranks as
 * dense everywhere).  Same seeded set, same per-ray results; the union over the ranks is the whole set.
 * arv2_render_tiles traces one rank's tiles into the fp64 histogram without exchange or finalise (like arv2_render_range;
 * per-ray records are indexed by the global ray id).  arv2_set_shard_mode: 1 = direction tiles (default), 0 = contiguous
 * slices (arv2_shard_range). */
int arv2_render_tiles(arv2_ctx* ctx, int32_t rank, int32_t n_ranks, int32_t zero_first, double* ms);
int arv2_set_shard_mode(arv2_ctx* ctx, int32_t mode);
/* AudioRenderer::render on n_ranks GPUs: trace this rank's share of the seeded ray set, ncclAllReduce
 * the fp64 histogram and finalise on the context's stream, one host synchronisation at the end.
 * Every rank ends up with the full IR; it equals the single-GPU arv2_render to fp32 rounding.
 * Collective: all ranks call it with contexts created from the same scene, parameters and seed. */
int arv2_render_sharded(arv2_ctx* ctx, arv2_comm* comm, double* ms);
/* One process driving several devices (the reference application is one process): one context
 * and one NCCL rank per device (ncclCommInitAll); arv2_multi_render runs arv2_render_sharded on one
 * host thread per device.  Set parameters on every context (arv2_multi_ctx(m, i)); read the IR
 * from any of them. */
typedef struct arv2_multi arv2_multi;
int arv2_multi_create(const arv2_scene* scene, const arv2_receiver* receiver, const arv2_renderer_desc* desc,
                      const int32_t* devices, int32_t n_devices, arv2_multi** out);
int32_t arv2_multi_size(const arv2_multi* m);
arv2_ctx* arv2_multi_ctx(arv2_multi* m, int32_t i);
int arv2_multi_render(arv2_multi* m, double* ms);   /* ms = the slowest device's trace */
void arv2_multi_destroy(arv2_multi* m);

/* Interactive receiver move: re-deposit from the cached receiver-independent
 * paths (requires desc.path_cache and one arv2_render since the last emitter
 * change).  Bit-identical to arv2_render for the same seed. */
int arv2_rerender(arv2_ctx* ctx, double* ms);

/* IR buffer layout of LaunchParams::ir_left / ir_right (OR/LaunchParams.h:41-42):
 * float[bands][ir_length] per ear, index = sample at sample_rate. */
int arv2_ir_length(const arv2_ctx* ctx, int32_t* ir_length, int32_t* bands);
int arv2_get_ir(arv2_ctx* ctx, float* ir_left, float* ir_right);           /* D2H      */
int arv2_set_ir(arv2_ctx* ctx, const float* ir_left, const float* ir_right);/* H2D (band 0) */
int arv2_ir_device(arv2_ctx* ctx, float** d_left, float** d_right);
/* fp64 accumulation histogram double[2][bands][ir_length] (device pointer). */
int arv2_hist_device(arv2_ctx* ctx, double** d_hist, int64_t* count);
/* Host->device bytes of the last receiver placement (top node + receiver tree + triangles). */
int arv2_last_upload_bytes(arv2_ctx* ctx, int64_t* bytes);
/* The receiver-independent path cache arv2_rerender scans (desc.path_cache, after one arv2_render): cached segments
 * of the whole ray set and the device bytes they occupy (32 B record + 4 B x bands energy + 8 B scan vertex each). */
int arv2_path_cache_info(arv2_ctx* ctx, int64_t* segments, int64_t* bytes);
/* Launch counters of the last render / re-render (diagnostics; unsigned 64-bit each): [1] segments traced;
 * in a library built with -DARV2_TRACE_STATS (lib/libarv2_stats.so) also [16] BVH node visits, [17] warp-level
 * node steps, [18] leaf visits, [19] triangle tests, [20] warp-level leaf steps -- what bench.py's "traversal"
 * figures are computed from.  Entries beyond what the library keeps read 0. */
#define ARV2_N_COUNTERS 24
int arv2_last_counters(arv2_ctx* ctx, uint64_t* out, int32_t n);
/* Segments (closest-hit queries) traced by the last render on this context. */
int arv2_last_segments(arv2_ctx* ctx, int64_t* segments);
/* Per-ray records of the last render (desc.record_rays), indexed by the ray's position in the
 * range last traced: *n_rays (may be NULL) = rays in that range; at most `capacity` entries are
 * copied into each non-NULL array (energy is float[capacity][bands]).  Call with capacity 0 to
 * learn the size. */
int arv2_get_records(arv2_ctx* ctx, int64_t capacity, int32_t* bin, int32_t* ear, float* energy,
                     int32_t* nseg, int64_t* n_rays);
/* Text dump of AudioRenderer::render's write_ir_to_file branch
 * (OR/AudioRenderer.cpp:525-567): one value per line, ostream default format. */
int arv2_write_ir_text(arv2_ctx* ctx, const char* left_path, const char* right_path);

/* Text dump of convoluteAudioFile's write_output_to_file branch (OR/AudioRenderer.cpp:720-744;
 * read by utils/main.py:17-28): n values per file, one per line, ostream default format.  The
 * reference names the files output_convolute_left.txt / output_convolute_right.txt. */
int arv2_write_convolved_text(const char* left_path, const char* right_path, const float* left,
                              const float* right, size_t n);

/* ------------------------------------------------------------ convolution -- */
typedef enum {
    ARV2_CONV_LINEAR = 0,    /* true linear convolution, gain 1, first n samples     */
    ARV2_CONV_REFERENCE = 1  /* reference semantics: 1 s segments, circular at ir_len,
                                gain 2, whole seconds only (OR/kernels.cu:382-438,
                                OR/AudioRenderer.cpp:702-711)                        */
} arv2_conv_mode;

/* AudioRenderer::convoluteAudioFile (OR/AudioRenderer.cpp:663-750): convolve the
 * whole dry signal x[n] (host) with the current ir_left / ir_right (band 0);
 * y_left / y_right receive n samples.  conv_ms = device time of the convolution,
 * process_ms = wall time incl. H2D/D2H (the reference's two timers). */
int arv2_convolve_file(arv2_ctx* ctx, const float* x, size_t n, float* y_left, float* y_right,
                       int32_t mode, double* conv_ms, double* process_ms);

/* AudioRenderer::convoluteLiveInput (OR/AudioRenderer.cpp:593-661) re-designed as a
 * uniformly partitioned overlap-add stream: n_sources mono inputs, `block` samples
 * per call, each convolved with its own stereo IR (block must be a power of two in
 * [64,1024]; the FFT size is 2*block). */
typedef struct arv2_stream arv2_stream;
int arv2_stream_open(int32_t device, int32_t n_sources, int32_t block, int32_t ir_length,
                     arv2_stream** out);
/* Load / swap the IR of one source (host float[ir_length] per ear).  Takes effect
 * atomically at the next block boundary (double-buffered partition spectra).  Asynchronous: the
 * spectra are computed and the pointer flips in stream order; steps enqueued before the call (on
 * any stream) finish with the old IR, steps enqueued after it use the new one. */
int arv2_stream_set_ir(arv2_stream* s, int32_t source, const float* ir_left, const float* ir_right);
/* Same, from device pointers (e.g. arv2_ir_device of a renderer on this GPU). */
int arv2_stream_set_ir_device(arv2_stream* s, int32_t source, const float* d_left, const float* d_right);
/* One block: in = float[n_sources][block] (host), out = float[n_sources][2][block]
 * (host, left then right per source). */
int arv2_stream_process(arv2_stream* s, const float* in, float* out);
/* Same with device-resident buffers, enqueued on `cuda_stream` without syncing. */
int arv2_stream_process_device(arv2_stream* s, const float* d_in, float* d_out, void* cuda_stream);
/* n_blocks consecutive blocks in one call (an RtAudio callback of the reference carries 4096 frames = 8 blocks of
 * 512, OR/main.cpp:99-135): d_in = float[n_blocks][n_sources][block], d_out = float[n_blocks][n_sources][2][block],
 * device-resident, enqueued on `cuda_stream` without syncing.  The steps are launched back to back and overlap on
 * the device (programmatic dependent launch; blocks 1.. start their forward FFT before block 0 has finished, so all of
 * d_in must have been produced by work enqueued on `cuda_stream` BEFORE this call, which stream order gives for free);
 * the result is the one n_blocks single-block calls give.  Streams of <= 4 sources run 16 CTAs per source, larger ones 8
 * (the summation order over the IR partitions, hence the last bits of the output, depends on it; fixed per stream). */
int arv2_stream_process_device_blocks(arv2_stream* s, const float* d_in, float* d_out, int32_t n_blocks, void* cuda_stream);
/* Host buffers, up to 16 consecutive blocks per call (one RtAudio callback): in = float[n_blocks][n_sources][block];
 * out (may be NULL) = float[n_blocks][n_sources][2][block]; mix (may be NULL) = float[n_blocks][2][block], the stereo
 * sum of the sources.  The kernels read and write pinned, device-mapped staging directly and the call returns on a
 * completion word: no copy-engine transfers, no stream synchronisation. */
int arv2_stream_process_blocks(arv2_stream* s, const float* in, float* out, float* mix, int32_t n_blocks);
/* The stereo buffer playback consumes (OR/main.cpp:69-97): d_mix[b][ear][t] = sum over this stream's sources, in source
 * order, of gain[s] * d_out[b][s][ear][t] (device buffers; enqueued on `cuda_stream`).  Across GPUs, sum the per-rank
 * mixes onto one rank with arv2_comm_reduce_f32. */
int arv2_stream_mix_device(arv2_stream* s, const float* d_out, float* d_mix, int32_t n_blocks, void* cuda_stream);
/* Per-source gains of the mix (host float[n_sources]; NULL = all 1). */
int arv2_stream_set_gains(arv2_stream* s, const float* gains);
int arv2_stream_reset(arv2_stream* s);
void arv2_stream_close(arv2_stream* s);

/* ------------------------------------------------- playback / re-render policy -- */
/* Camera::calculate_global_angle (OR/Camera.cpp:31-41): degrees(atan2(o.z, o.x)) in
 * [0, 360); the receiver is rotated by -angle about +Y (OR/OptixModel.cpp:178-181). */
float arv2_global_angle(float orientation_x, float orientation_z);

/* Re-render trigger of the GL loop (OR/main.cpp:470-498): a re-render starts when the
 * receiver moved more than distance_threshold since the last render, or turned more than
 * angle_threshold degrees (shortest arc), or more than 1 s (whole seconds, like time())
 * passed since the first movement after the last render -- and none is in flight. */
typedef struct arv2_rerender_policy arv2_rerender_policy;
int arv2_policy_create(float distance_threshold, float angle_threshold_deg, const float start_pos[3],
                       float start_angle_deg, arv2_rerender_policy** out);
/* Returns 1 when a re-render must be started now (and records pos / angle as its origin),
 * 0 otherwise.  now_s = wall clock in seconds; is_rendering = a render is still in flight. */
int arv2_policy_update(arv2_rerender_policy* p, const float pos[3], float angle_deg, double now_s, int32_t is_rendering);
void arv2_policy_destroy(arv2_rerender_policy* p);

/* audioHandler (OR/main.cpp:69-97): fill one RtAudio callback (RTAUDIO_FLOAT64, interleaved
 * LRLR, n_frames frames) from the convolved file buffers.  Reproduces the reference's
 * indexing: position = (int)(stream_time * sample_rate) % n_samples, then for i in
 * [0, 2*n_frames): out[i] = (i even ? left : right)[i + position] * 100 * volume, stopping
 * at output_buffer_len (which the reference sets in BYTES, OR/main.cpp:51,85).
 * Returns the number of doubles written. */
int64_t arv2_playback_fill(double* out, uint32_t n_frames, double stream_time, int32_t sample_rate,
                           const float* left, const float* right, size_t n_samples,
                           size_t output_buffer_len, float volume);

/* CircularBuffer<double> of the live path (OR/CircularBuffer.h): add() overlap-adds at the
 * read position without advancing, get_and_reset() pops n values, zeroing them. */
typedef struct arv2_ring arv2_ring;
int arv2_ring_create(size_t size, arv2_ring** out);
int arv2_ring_add(arv2_ring* r, const double* values, size_t n);
int arv2_ring_get_and_reset(arv2_ring* r, double* out, size_t n);   /* ARV2_ERR_INVALID if n > size */
void arv2_ring_destroy(arv2_ring* r);
/* One mic callback (audioHandlerWithMic + convoluteLiveInput, OR/main.cpp:99-135,
 * OR/AudioRenderer.cpp:593-661) on a 1-source stream convolver: n_in samples (a multiple of
 * the stream's block) are convolved block by block, scaled by the reference's gain 2,
 * interleaved LRLR and overlap-added into the ring. */
int arv2_live_callback(arv2_stream* s, const double* in, size_t n_in, arv2_ring* ring);

/* ------------------------------------------------------------------ audio -- */
/* AudioFile<float>::load as used by Context.cpp:198-213: channel 0 only, int16 ->
 * x/32768, float32 passthrough.  *samples is malloc'd; free with arv2_free. */
int arv2_wav_read(const char* path, float** samples, size_t* n, int32_t* sample_rate, int32_t* channels);
/* export mode (OR/main.cpp:628-718): per-channel min-max normalisation to [-1,1],
 * 16-bit stereo WAV. */
int arv2_wav_write_stereo_normalized(const char* path, const float* left, const float* right,
                                     size_t n, int32_t sample_rate);
void arv2_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* ARV2_H */

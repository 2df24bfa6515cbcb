// oracle/ref_device_shim.cpp -- TEST INFRASTRUCTURE (pins the oracle; never part of the product).
//
// Runs the reference's OWN device programs on the CPU: OR/devicePrograms.cu (__raygen__renderFrame,
// __closesthit__radiance, __miss__radiance) is #included below UNMODIFIED from /root/reference where it lies and
// compiled with g++ -ffp-contract=off (no fast-math) against oracle/ref_stubs/ (host stand-ins for the OptiX device
// API and cuRAND).  What the reference delegates to closed third-party code is supplied by this shim and says so:
//   * optixTrace's closest-hit search (OptiX 7.7 RT cores): a brute-force two-sided Moeller-Trumbore over the flat
//     triangle list, closest = min t, ties to the lower triangle id, 0 < t < 1e20 -- the documented contract the
//     oracle also follows; the barycentrics it reports are relative to (P1; P2, P3) like OptiX's;
//   * curand_uniform (XORWOW seeded by clock64()): the caller passes the two uniforms of every ray.
// Everything else -- energy initialisation, direction from (u1, u2), loop guard, hit point, path length, chord
// weight, bin, inter-aural delay, deposit rule, reflection, absorption, origin offset -- is the reference's code.
//
// Build: oracle/Makefile target `ref` -> oracle/_ref/libref_device.so.  Used by tests/golden/make_golden.py (to write
// tests/golden/ref_shading.npz, ref_render_c1.npz) and, where /root/reference is mounted, live by tests/test_pin_cpu.py.
#include <algorithm>
#include <climits>
#include <cstring>
#include <thread>
#include <vector>

#include "optix_device.h"
#include "curand_kernel.h"

thread_local ref_optix_state_t ref_optix_state;
thread_local const float* ref_uniforms = nullptr;

#include "devicePrograms.cu"      // the reference source, from -I/root/reference/prebuild/obj_raytracer

LaunchParams optixLaunchParams;   // the __constant__ block optixLaunch fills (OR/devicePrograms.cu:20)

namespace {

struct RefMesh { std::vector<glm::vec3> vertex; std::vector<glm::ivec3> index; TriangleMeshSBTData sbt; };

struct RefScene {
    std::vector<RefMesh> meshes;
    std::vector<int> tri_mesh, tri_prim;      // flat triangle -> (mesh, primitive)
    const float* tv = nullptr; int64_t T = 0;
};

thread_local const RefScene* t_scene = nullptr;
// per-thread deposit log of the ray in flight
struct DepLog { int n; int ear[2]; int idx[2]; float val[2]; };
thread_local DepLog t_dep;
thread_local double* t_hist = nullptr;       // [2][ir_length], may be null

struct V { float x, y, z; };
inline V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dotf(V a, V b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline V crossf(V a, V b) { return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))}; }

} // namespace

void ref_atomic_add(float* addr, float v)
{
    const int ir_len = optixLaunchParams.ir_length;
    const bool left = addr >= optixLaunchParams.ir_left && addr < optixLaunchParams.ir_left + ir_len;
    const int idx = (int)(addr - (left ? optixLaunchParams.ir_left : optixLaunchParams.ir_right));
    if (t_dep.n < 2) { t_dep.ear[t_dep.n] = left ? 0 : 1; t_dep.idx[t_dep.n] = idx; t_dep.val[t_dep.n] = v; }
    t_dep.n++;
    if (t_hist && idx >= 0 && idx < ir_len) t_hist[(size_t)(left ? 0 : 1) * ir_len + idx] += (double)v;
}

// Stand-in for the OptiX traversal + program dispatch (see the header comment).
void ref_trace_dispatch(float ox, float oy, float oz, float dx, float dy, float dz, float tmin, float tmax, uint32_t p0, uint32_t p1)
{
    const RefScene& S = *t_scene;
    const V org{ox, oy, oz}, dir{dx, dy, dz};
    float best_t = tmax, best_u = 0.f, best_v = 0.f; int64_t best = -1;
    for (int64_t i = 0; i < S.T; ++i) {
        const float* p = S.tv + 9 * i;
        const V p1v{p[0], p[1], p[2]}, e1 = sub(V{p[3], p[4], p[5]}, p1v), e2 = sub(V{p[6], p[7], p[8]}, p1v);
        const V pvec = crossf(dir, e2);
        float det = dotf(e1, pvec);
        if (det == 0.0f) continue;
        const V tvec = sub(org, p1v);
        float U = dotf(tvec, pvec);
        const V qvec = crossf(tvec, e1);
        float Vv = dotf(dir, qvec), Tt = dotf(e2, qvec);
        if (det < 0.0f) { det = -det; U = -U; Vv = -Vv; Tt = -Tt; }
        if (!(U >= 0.0f && Vv >= 0.0f && U + Vv <= det) || !(Tt > 0.0f)) continue;
        const float t = Tt / det;
        if (!(t > tmin && t < tmax)) continue;
        if (t < best_t) { best_t = t; best_u = U / det; best_v = Vv / det; best = i; }
    }
    ref_optix_state.trace_calls++;
    ref_optix_state.p0 = p0; ref_optix_state.p1 = p1;
    ref_optix_state.ray_dir = make_float3(dx, dy, dz);
    if (best < 0) { ref_optix_state.sbt_data = nullptr; __miss__radiance(); return; }
    ref_optix_state.sbt_data = &S.meshes[S.tri_mesh[best]].sbt;
    ref_optix_state.prim = S.tri_prim[best];
    ref_optix_state.bary = make_float2(best_u, best_v);
    __closesthit__radiance();
}

namespace {

void build_scene(RefScene& S, const float* tri_verts, const int32_t* tri_mat, int64_t T, const float* absorption)
{
    S.tv = tri_verts; S.T = T;
    S.tri_mesh.resize(T); S.tri_prim.resize(T);
    S.meshes.reserve((size_t)T + 1);      // sbt pointers must stay valid
    int cur = INT_MIN;
    for (int64_t i = 0; i < T; ++i) {
        if (S.meshes.empty() || tri_mat[i] != cur) {          // one mesh per run of equal material (loadOBJ order)
            cur = tri_mat[i];
            S.meshes.emplace_back();
            S.meshes.back().sbt.mat_absorption = cur >= 0 ? absorption[cur] : (float)cur;     // -1 / -2: the ears
        }
        RefMesh& m = S.meshes.back();
        const int base = (int)m.vertex.size();
        for (int k = 0; k < 3; ++k) m.vertex.push_back(glm::vec3(tri_verts[9 * i + 3 * k], tri_verts[9 * i + 3 * k + 1], tri_verts[9 * i + 3 * k + 2]));
        S.tri_mesh[i] = (int)S.meshes.size() - 1; S.tri_prim[i] = (int)m.index.size();
        m.index.push_back(glm::ivec3(base, base + 1, base + 2));
    }
    for (RefMesh& m : S.meshes) { m.sbt.vertex = m.vertex.data(); m.sbt.index = m.index.data(); }
}

std::vector<float> g_ir_l, g_ir_r;      // address ranges only: the deposits are logged, never added here

void set_launch(int sx, int sy, int sz, const float* emitter, const float* center, float base_power, float energy_thres,
                unsigned max_bounces, float hrtf, int sample_rate, int is_mono, int ir_length)
{
    g_ir_l.assign((size_t)ir_length, 0.f); g_ir_r.assign((size_t)ir_length, 0.f);
    LaunchParams& L = optixLaunchParams;
    L.size_x = sx; L.size_y = sy; L.size_z = sz;
    L.emitter_position = glm::vec3(emitter[0], emitter[1], emitter[2]);
    L.sphere_center = glm::vec3(center[0], center[1], center[2]);
    L.traversable = 0;
    L.base_power = base_power; L.energy_thres = energy_thres; L.max_bounces = max_bounces;
    L.hrtf_absorption_rate = hrtf; L.sample_rate = sample_rate; L.isMono = is_mono != 0; L.ir_length = ir_length;
    L.ir_left = g_ir_l.data(); L.ir_right = g_ir_r.data();
}

} // namespace

extern "C" {

/* One invocation of the reference's __closesthit__radiance on a hand-made hit.  prd8 = {remaining_factor, distance,
 * prev_position[3], direction[3]} in/out, *depth in/out.  Deposits (<= 2): ear 0 = ir_left, 1 = ir_right. */
void ref_closesthit(const float* tri9, float mat_absorption, const float* ray_dir, float u, float v, const float* sphere_center,
                    int sample_rate, float hrtf, int is_mono, int ir_length, float* prd8, int* depth,
                    int* n_dep, int* dep_ear, int* dep_idx, float* dep_val)
{
    const float zero[3] = {0.f, 0.f, 0.f};
    set_launch(1, 1, 1, zero, sphere_center, 0.f, 0.f, 0u, hrtf, sample_rate, is_mono, ir_length);
    glm::vec3 vertex[3] = {glm::vec3(tri9[0], tri9[1], tri9[2]), glm::vec3(tri9[3], tri9[4], tri9[5]), glm::vec3(tri9[6], tri9[7], tri9[8])};
    glm::ivec3 index(0, 1, 2);
    TriangleMeshSBTData sbt;
    sbt.vertex = vertex; sbt.index = &index; sbt.mat_absorption = mat_absorption;
    PRD prd;
    prd.remaining_factor = prd8[0]; prd.distance = prd8[1];
    prd.sphere_center = glm::vec3(sphere_center[0], sphere_center[1], sphere_center[2]);
    prd.prev_position = glm::vec3(prd8[2], prd8[3], prd8[4]);
    prd.direction = glm::vec3(prd8[5], prd8[6], prd8[7]);
    prd.recursion_depth = *depth;
    uint32_t p0, p1;
    packPointer(&prd, p0, p1);
    ref_optix_state = ref_optix_state_t{};
    ref_optix_state.launch_dims = make_uint3(1, 1, 1);
    ref_optix_state.sbt_data = &sbt; ref_optix_state.prim = 0;
    ref_optix_state.ray_dir = make_float3(ray_dir[0], ray_dir[1], ray_dir[2]);
    ref_optix_state.bary = make_float2(u, v);
    ref_optix_state.p0 = p0; ref_optix_state.p1 = p1;
    t_dep = DepLog{}; t_hist = nullptr;
    __closesthit__radiance();
    prd8[0] = prd.remaining_factor; prd8[1] = prd.distance;
    prd8[2] = prd.prev_position.x; prd8[3] = prd.prev_position.y; prd8[4] = prd.prev_position.z;
    prd8[5] = prd.direction.x; prd8[6] = prd.direction.y; prd8[7] = prd.direction.z;
    *depth = prd.recursion_depth;
    *n_dep = t_dep.n;
    for (int i = 0; i < 2; ++i) { dep_ear[i] = t_dep.ear[i]; dep_idx[i] = t_dep.idx[i]; dep_val[i] = t_dep.val[i]; }
}

/* The reference's __raygen__renderFrame for launch indices [ray_begin, ray_begin + n) of a size_x*size_y*size_z launch
 * over a flat scene (tri_mat >= 0: wall with absorption[tri_mat]; -1 / -2: receiver_left / receiver_right).
 * uniforms = float[n][2]: what curand_uniform returns for theta and phi of each ray.
 * Outputs (any may be NULL): hist double[2][ir_length] = every atomicAdd, summed in fp64; per ray: bin of the primary
 * deposit (-1: nothing deposited), ear of the primary deposit (0 none, 1 ir_left, 2 ir_right), the value deposited
 * there (remaining_factor at the hit), optixTrace calls.
 * Returns the total number of optixTrace calls. */
int64_t ref_render(const float* tri_verts, const int32_t* tri_mat, int64_t n_tris, const float* absorption,
                   int size_x, int size_y, int size_z, const float* emitter, const float* sphere_center, float base_power,
                   float energy_thres, unsigned max_bounces, float hrtf, int sample_rate, int is_mono, int ir_length,
                   const float* uniforms, int64_t ray_begin, int64_t n, int n_threads,
                   double* hist, int32_t* rec_bin, int32_t* rec_ear, float* rec_energy, int32_t* rec_nseg)
{
    RefScene S;
    build_scene(S, tri_verts, tri_mat, n_tris, absorption);
    set_launch(size_x, size_y, size_z, emitter, sphere_center, base_power, energy_thres, max_bounces, hrtf, sample_rate, is_mono, ir_length);
    const int nt = std::max(1, n_threads);
    std::vector<std::vector<double>> priv(nt);
    std::vector<int64_t> calls(nt, 0);
    auto work = [&](int tid) {
        t_scene = &S;
        if (hist) { priv[tid].assign((size_t)2 * ir_length, 0.0); t_hist = priv[tid].data(); } else t_hist = nullptr;
        for (int64_t i = n * tid / nt; i < n * (tid + 1) / nt; ++i) {
            const int64_t id = ray_begin + i;
            ref_optix_state = ref_optix_state_t{};
            ref_optix_state.launch_dims = make_uint3((unsigned)size_x, (unsigned)size_y, (unsigned)size_z);
            ref_optix_state.launch_index = make_uint3((unsigned)(id % size_x), (unsigned)((id / size_x) % size_y), (unsigned)(id / ((int64_t)size_x * size_y)));
            ref_uniforms = uniforms + 2 * i;
            t_dep = DepLog{};
            __raygen__renderFrame();
            calls[tid] += (int64_t)ref_optix_state.trace_calls;
            if (rec_bin) rec_bin[i] = t_dep.n > 0 ? t_dep.idx[0] : -1;
            if (rec_ear) rec_ear[i] = t_dep.n > 0 ? t_dep.ear[0] + 1 : 0;
            if (rec_energy) rec_energy[i] = t_dep.n > 0 ? t_dep.val[0] : 0.f;
            if (rec_nseg) rec_nseg[i] = (int32_t)ref_optix_state.trace_calls;
        }
    };
    if (nt == 1) work(0);
    else { std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(work, t); for (auto& t : th) t.join(); }
    if (hist) {
        std::memset(hist, 0, sizeof(double) * 2 * (size_t)ir_length);
        for (int t = 0; t < nt; ++t) for (size_t i = 0; i < (size_t)2 * ir_length; ++i) hist[i] += priv[t][i];
    }
    int64_t total = 0;
    for (auto c : calls) total += c;
    return total;
}

} // extern "C"

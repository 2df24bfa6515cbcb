// trace.cu -- sm_100a sound-ray tracing kernels of libarv2.
//
// Replaces the OptiX pipeline of the reference (B200 has no RT cores):
//   __raygen__renderFrame     OR/devicePrograms.cu:192-254  -> trace_kernel (path loop)
//   optixTrace / GAS traversal OR/devicePrograms.cu:240-251 -> closest_hit (software BVH)
//   __closesthit__radiance    OR/devicePrograms.cu:62-180   -> shade step in trace_kernel
//   __miss__radiance          OR/devicePrograms.cu:186-190  -> "no hit" branch
//   fillZeros / addIRs        OR/kernels.cu:77-97,519-536   -> cudaMemsetAsync / finalize_kernel
//
// Design: persistent CTAs; every warp keeps 32 paths in flight and refills lanes whose
// path ended from a warp-local chunk of the seeded ray set (ballot + popc compaction),
// so no ray state ever goes through HBM.  Receiver deposits are aggregated across the
// warp (match.any) and accumulated in an fp64 histogram with native RED.F64, which
// makes the result independent of the deposit order to ~1e-16.
#include <climits>

#include "arv2_model.cuh"
#include "trace.cuh"

namespace arv2 {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kChunk = 128;          // rays a warp claims per global atomic
constexpr int kStack = 64;
#ifndef ARV2_THREADS
#define ARV2_THREADS 128
#endif
#ifndef ARV2_MINB
#define ARV2_MINB 8
#endif
constexpr int kThreads = ARV2_THREADS;
constexpr int kSentinel = INT_MIN;
constexpr int kRerenderThreads = 256;
constexpr int kRerenderBatch = 12;   // lanes that must hold a candidate before the receiver walk

struct Hit { float t, u, v; int slot, id; };

// 256-bit read-only global load (sm_100+: LDG.E.ENL2.256.CONSTANT).  A divergent gather
// costs the L1 data pipe one wavefront per load instruction and lane, so a 64 B node is
// fetched with 2 of these instead of 4 x LDG.128 (profiles/micro/gather.cu: 1.34x).
struct __align__(32) F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg256(const float4* p)
{
    F8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
}

__device__ __forceinline__ float safe_rcp(float d)
{
    const float eps = 1e-20f;
    return 1.0f / (fabsf(d) > eps ? d : copysignf(eps, d));
}

// Closest hit = min (t, global triangle id) over all triangles whose exact test
// accepts; the BVH only prunes (boxes are padded, comparison is <=).
__device__ __forceinline__ void closest_hit(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                            int root, F3 org, F3 dir, float tmax, Hit& h)
{
    const float ix = safe_rcp(dir.x), iy = safe_rcp(dir.y), iz = safe_rcp(dir.z);
    const float ox = org.x * ix, oy = org.y * iy, oz = org.z * iz;
    int stack[kStack];
    int sp = 0;
    stack[sp++] = kSentinel;
    int cur = root;
    h.t = tmax; h.u = 0.f; h.v = 0.f; h.slot = -1; h.id = INT_MAX;
#ifdef ARV2_STATS
    int n_inner = 0, n_tri = 0, n_leaf = 0;
#endif

    while (cur != kSentinel) {
        while (cur >= 0) {
#ifdef ARV2_STATS
            n_inner++;
#endif
            const F8 na = ldg256(nodes + cur * 4), nb = ldg256(nodes + cur * 4 + 2);
            const float4 n0 = na.lo, n1 = na.hi, n2 = nb.lo, n3 = nb.hi;
            const float c0lox = fmaf(n0.x, ix, -ox), c0hix = fmaf(n0.y, ix, -ox);
            const float c0loy = fmaf(n0.z, iy, -oy), c0hiy = fmaf(n0.w, iy, -oy);
            const float c0loz = fmaf(n2.x, iz, -oz), c0hiz = fmaf(n2.y, iz, -oz);
            const float c1lox = fmaf(n1.x, ix, -ox), c1hix = fmaf(n1.y, ix, -ox);
            const float c1loy = fmaf(n1.z, iy, -oy), c1hiy = fmaf(n1.w, iy, -oy);
            const float c1loz = fmaf(n2.z, iz, -oz), c1hiz = fmaf(n2.w, iz, -oz);
            const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.f));
            const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), h.t));
            const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.f));
            const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), h.t));
            const bool go0 = c0min <= c0max, go1 = c1min <= c1max;
            const int i0 = __float_as_int(n3.x), i1 = __float_as_int(n3.y);
            if (!go0 && !go1) {
                cur = stack[--sp];
            } else {
                cur = go0 ? i0 : i1;
                if (go0 && go1) {
                    int far = i1;
                    if (c1min < c0min) { cur = i1; far = i0; }
                    stack[sp++] = far;
                }
            }
        }
        if (cur == kSentinel) break;
        // leaf
        const int code = ~cur;
        const int first = code >> 3;
        const int cnt = (code & 7) + 1;
#ifdef ARV2_STATS
        n_leaf++; n_tri += cnt;
#endif
        for (int i = 0; i < cnt; ++i) {
            const int slot = first + i;
            const float4 a = __ldg(tris + slot * 3 + 0);
            const float4 b = __ldg(tris + slot * 3 + 1);
            const float4 c = __ldg(tris + slot * 3 + 2);
            float t, u, v;
            if (tri_test(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), org, dir, &t, &u, &v)) {
                const int id = __float_as_int(a.w);
                if (t < h.t || (t == h.t && id < h.id)) { h.t = t; h.u = u; h.v = v; h.slot = slot; h.id = id; }
            }
        }
        cur = stack[--sp];
    }
#ifdef ARV2_STATS
    h.id = n_inner | (n_leaf << 10) | (n_tri << 20);
#endif
}

// Warp-aggregated deposit into the fp64 histogram (OR/devicePrograms.cu:128-170).
// Called by all 32 lanes; `dep` lanes carry (bin, primary ear, energy[]).
template <int NB>
__device__ __forceinline__ void deposit_warp(const TraceParams& p, bool dep, int bin, int primary, const float* energy)
{
    const unsigned dm = __ballot_sync(FULL, dep);
    if (dm == 0 || !dep) return;
    const int key = bin | (primary << 30);
    const unsigned peers = __match_any_sync(dm, key);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    const int obin = (bin + p.delay < p.ir_len) ? bin + p.delay : bin;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        double s = 0.0, sc = 0.0;
        const float cross = __fmul_rn(energy[b], p.cross_gain);
        unsigned m = peers;
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            s += (double)__shfl_sync(peers, energy[b], src);
            sc += (double)__shfl_sync(peers, cross, src);
        }
        if (lane == leader) {
            atomicAdd(p.hist + ((size_t)(primary * NB + b) * p.ir_len + bin), s);
            if (!p.mono) atomicAdd(p.hist + ((size_t)((1 - primary) * NB + b) * p.ir_len + obin), sc);
        }
    }
}

// Receiver hit: chord weighting and bin (OR/devicePrograms.cu:91-133).
template <int NB>
__device__ __forceinline__ int receiver_hit(const TraceParams& p, F3 pt, F3 dir, float distance, float* energy)
{
    const float dinv = __fdiv_rn(1.0f, __fsqrt_rn(dot3(dir, dir)));
    const F3 nd = f3(__fmul_rn(dir.x, dinv), __fmul_rn(dir.y, dinv), __fmul_rn(dir.z, dinv));
    const F3 oc = sub3(pt, f3(p.center[0], p.center[1], p.center[2]));
    const float a = dot3(nd, nd);
    const float bq = __fmul_rn(2.0f, dot3(oc, nd));
    const float c = __fsub_rn(dot3(oc, oc), 1.0f);
    const float disc = __fmaf_rn(bq, bq, -__fmul_rn(__fmul_rn(4.0f, a), c));
    float wgt = 0.f;
    if (disc > 0.f) {
        const float sq = __fsqrt_rn(disc);
        const float a2 = __fmul_rn(2.0f, a);
        const float t1 = __fdiv_rn(__fsub_rn(-bq, sq), a2);
        const float t2 = __fdiv_rn(__fadd_rn(-bq, sq), a2);
        const F3 i1 = f3(__fmaf_rn(t1, nd.x, pt.x), __fmaf_rn(t1, nd.y, pt.y), __fmaf_rn(t1, nd.z, pt.z));
        const F3 i2 = f3(__fmaf_rn(t2, nd.x, pt.x), __fmaf_rn(t2, nd.y, pt.y), __fmaf_rn(t2, nd.z, pt.z));
        const F3 df = sub3(i1, i2);
        wgt = __fsqrt_rn(dot3(df, df));
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) energy[b] = __fmul_rn(energy[b], wgt);
    const float elapsed = __fdiv_rn(distance, 343.0f);
    return __float2int_rz(roundf(__fmul_rn(elapsed, p.fs)));
}

// Hit point and path length (OR/devicePrograms.cu:79-83).
__device__ __forceinline__ F3 hit_point(F3 p1, F3 p2, F3 p3, float u, float v)
{
    const float w = __fsub_rn(__fsub_rn(1.0f, u), v);
    return f3(__fmaf_rn(v, p3.x, __fmaf_rn(u, p2.x, __fmul_rn(w, p1.x))),
              __fmaf_rn(v, p3.y, __fmaf_rn(u, p2.y, __fmul_rn(w, p1.y))),
              __fmaf_rn(v, p3.z, __fmaf_rn(u, p2.z, __fmul_rn(w, p1.z))));
}

template <int NB, int MODE>
__global__ void __launch_bounds__(kThreads, ARV2_MINB) trace_kernel(const TraceParams p)
{
    const int lane = threadIdx.x & 31;
    long long chunk_next = 0, chunk_end = 0;      // warp-uniform
    bool have = false, exhausted = false;
    long long ray = 0;
    F3 org = f3(0, 0, 0), dir = f3(0, 0, 0);
    float energy[NB];
    float dist = 0.f;
    int depth = 0, nseg = 0;
    unsigned long long segs = 0;

    for (;;) {
        // ---- refill idle lanes from the warp's chunk (ballot/popc compaction)
        unsigned need = __ballot_sync(FULL, !have && !exhausted);
        while (need) {
            if (chunk_next >= chunk_end) {
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(p.counters, (unsigned long long)kChunk);
                b = __shfl_sync(FULL, b, 0);
                chunk_next = (long long)b;
                chunk_end = min((long long)b + kChunk, p.n_rays);
                if (chunk_next >= chunk_end) {
                    if (!have) exhausted = true;
                    break;
                }
            }
            const int avail = (int)min((long long)32, chunk_end - chunk_next);
            const int rank = __popc(need & ((1u << lane) - 1u));
            if (((need >> lane) & 1u) && rank < avail) {
                ray = chunk_next + rank;
                have = true;
                org = f3(p.emitter[0], p.emitter[1], p.emitter[2]);          // :210
                dir = emit_direction(p.seed, (uint64_t)(p.ray_begin + ray)); // :216-224
#pragma unroll
                for (int b = 0; b < NB; ++b) energy[b] = p.energy0;          // :208
                dist = 0.f; depth = 0; nseg = 0;                             // :209,:211
            }
            chunk_next += min(__popc(need), avail);
            need = __ballot_sync(FULL, !have && !exhausted);
        }
        if (!__any_sync(FULL, have)) break;

        bool ended = false, dep = false;
        int bin = -1, ear = 0, primary = 0;
        if (have) {
            float emax = energy[0];
#pragma unroll
            for (int b = 1; b < NB; ++b) emax = fmaxf(emax, energy[b]);
            const bool zero_dir = !(dir.x != 0.f || dir.y != 0.f || dir.z != 0.f);            // :230
            if (zero_dir || !(dist < p.dist_thr && emax > p.energy_thres && (unsigned)depth < p.max_bounces)) {
                ended = true;                                                                   // :233-236
            } else {
                const size_t ci = (size_t)nseg * (size_t)p.pc_stride + (size_t)ray;
                if (MODE == 1) {
                    p.pc_dir_d[ci] = make_float4(dir.x, dir.y, dir.z, dist);
#pragma unroll
                    for (int b = 0; b < NB; ++b) p.pc_energy[ci * NB + b] = energy[b];
                }
                nseg++;
                Hit h;
                closest_hit(p.nodes, p.tris, p.root, org, dir, 1e20f, h);
                if (MODE == 1) p.pc_org_t[ci] = make_float4(org.x, org.y, org.z, h.t);
#ifdef ARV2_STATS
                {
                    // counters[2..]: sum inner, sum warp-max inner * lanes, sum leaf, sum max leaf, sum tri, sum max tri, segments, warp-rounds*32
                    const unsigned am = __activemask();
                    const int ni = h.id & 1023, nl = (h.id >> 10) & 1023, nt = (h.id >> 20) & 1023;
                    const int mi = __reduce_max_sync(am, ni), ml = __reduce_max_sync(am, nl), mt = __reduce_max_sync(am, nt);
                    const int si = __reduce_add_sync(am, ni), sl = __reduce_add_sync(am, nl), st = __reduce_add_sync(am, nt);
                    if ((threadIdx.x & 31) == __ffs(am) - 1) {
                        atomicAdd(p.counters + 2, (unsigned long long)si); atomicAdd(p.counters + 3, (unsigned long long)mi * 32);
                        atomicAdd(p.counters + 4, (unsigned long long)sl); atomicAdd(p.counters + 5, (unsigned long long)ml * 32);
                        atomicAdd(p.counters + 6, (unsigned long long)st); atomicAdd(p.counters + 7, (unsigned long long)mt * 32);
                        atomicAdd(p.counters + 8, (unsigned long long)__popc(am)); atomicAdd(p.counters + 9, 32ull);
                    }
                }
#endif
                if (h.slot < 0) {
                    ended = true;                                                               // miss :186-190
                } else {
                    const float4 a = __ldg(p.tris + h.slot * 3 + 0);
                    const float4 b4 = __ldg(p.tris + h.slot * 3 + 1);
                    const float4 c4 = __ldg(p.tris + h.slot * 3 + 2);
                    const F3 p1 = f3(a.x, a.y, a.z), p2 = f3(b4.x, b4.y, b4.z), p3 = f3(c4.x, c4.y, c4.z);
                    const int mat = __float_as_int(b4.w);
                    const F3 pt = hit_point(p1, p2, p3, h.u, h.v);
                    const F3 dp = sub3(pt, org);
                    dist = __fadd_rn(dist, __fsqrt_rn(dot3(dp, dp)));                           // :83
                    if (mat < 0) {
                        bin = receiver_hit<NB>(p, pt, dir, dist, energy);
                        ear = (mat == -1) ? 1 : 2;
                        primary = (mat == -1) ? 0 : 1;
                        dep = bin >= 0 && bin < p.ir_len;
                        ended = true;                                                           // :147,:169
                    } else {
                        // :75-77  Ng = normalize(cross(P2-P1, P3-P1))
                        const F3 nc = cross3(sub3(p2, p1), sub3(p3, p1));
                        const float ninv = __fdiv_rn(1.0f, __fsqrt_rn(dot3(nc, nc)));
                        const F3 ng = f3(__fmul_rn(nc.x, ninv), __fmul_rn(nc.y, ninv), __fmul_rn(nc.z, ninv));
                        bool diffuse = false;
                        uint32_t r[4];
                        if (p.any_scatter) {
                            const float sc = __ldg(p.scattering + mat);
                            if (sc > 0.f) {
                                philox4x32(p.seed, (uint64_t)(p.ray_begin + ray), (uint32_t)depth, 1u, r);
                                diffuse = __fmul_rn((float)(r[0] >> 8), 0x1p-24f) < sc;
                            }
                        }
                        if (diffuse) {
                            dir = lambert_direction(r, dir, ng);
                        } else {
                            const float k = __fmul_rn(2.0f, dot3(dir, ng));                     // :173
                            dir = f3(__fmaf_rn(-k, ng.x, dir.x), __fmaf_rn(-k, ng.y, dir.y), __fmaf_rn(-k, ng.z, dir.z));
                        }
#pragma unroll
                        for (int b = 0; b < NB; ++b) energy[b] = __fmul_rn(energy[b], __ldg(p.keep + mat * NB + b)); // :174
                        depth++;                                                                // :175
                        org = f3(__fmaf_rn(1e-3f, dir.x, pt.x), __fmaf_rn(1e-3f, dir.y, pt.y), __fmaf_rn(1e-3f, dir.z, pt.z)); // :179
                    }
                }
            }
        }
        if (MODE == 0) deposit_warp<NB>(p, dep, bin, primary, energy);
        if (ended) {
            if (p.rec_bin) p.rec_bin[ray] = bin;
            if (p.rec_ear) p.rec_ear[ray] = ear;
            if (p.rec_nseg) p.rec_nseg[ray] = nseg;
            if (p.rec_energy) {
#pragma unroll
                for (int b = 0; b < NB; ++b) p.rec_energy[ray * NB + b] = ear ? energy[b] : 0.f;
            }
            if (MODE == 1) p.pc_nseg[ray] = nseg;
            segs += (unsigned long long)nseg;
            have = false;
        }
    }
    // one atomic per warp for the segment counter
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(FULL, segs, o);
    if (lane == 0 && segs) atomicAdd(p.counters + 1, segs);
}

// ---------------------------------------------------------------------------------------
// trace2_kernel: the same path tracer as trace_kernel with the lanes of a warp decoupled.
// r01/r02 profiles: per-segment node-visit counts vary so much inside a warp (sum / (32 x max)
// = 49 %) that the inner loop of trace_kernel runs at 12 of 32 lanes.  Here a lane that has
// finished its segment parks (cur == kSentinel) while the others keep traversing; three
// warp-uniform phases are scheduled by ballots:
//   A  shade + deposit + refill + start the next segment -- when >= kTA lanes are parked
//   L  one leaf (<= 4 triangle tests) per lane           -- when >= kTL lanes stand at a leaf
//   I  up to kBurst inner-node steps                      -- otherwise
// so every phase runs with many lanes and nobody waits for the slowest traversal.
#ifndef ARV2_TA
#define ARV2_TA 12
#endif
#ifndef ARV2_TL
#define ARV2_TL 8
#endif
#ifndef ARV2_BURST
#define ARV2_BURST 4
#endif
template <int NB, int MODE>
__global__ void __launch_bounds__(kThreads, ARV2_MINB) trace2_kernel(const TraceParams p)
{
    const int lane = threadIdx.x & 31;
    long long chunk_next = 0, chunk_end = 0;      // warp-uniform
    bool have = false, exhausted = false, pending = false;
    long long ray = 0;
    F3 org = f3(0, 0, 0), dir = f3(0, 0, 0);
    float energy[NB];
    float dist = 0.f;
    int depth = 0, nseg = 0;
    unsigned long long segs = 0;
    // traversal state of the segment in flight
    int stack[kStack];
    int sp = 0, cur = kSentinel;
    float ix = 0.f, iy = 0.f, iz = 0.f, ox = 0.f, oy = 0.f, oz = 0.f;
    Hit h;
    h.t = 1e20f; h.u = 0.f; h.v = 0.f; h.slot = -1; h.id = INT_MAX;
    const float4* __restrict__ nodes = p.nodes;
    const float4* __restrict__ tris = p.tris;

    for (;;) {
        const bool at_inner = cur >= 0;
        const bool at_leaf = cur < 0 && cur != kSentinel;
        const bool parked = cur == kSentinel && !exhausted;
        const unsigned inner_m = __ballot_sync(FULL, at_inner);
        const unsigned leaf_m = __ballot_sync(FULL, at_leaf);
        const unsigned park_m = __ballot_sync(FULL, parked);
        if ((inner_m | leaf_m | park_m) == 0) break;

        if (park_m != 0 && (__popc(park_m) >= ARV2_TA || (inner_m | leaf_m) == 0)) {
            // ======================= phase A: shade, deposit, refill, start next segment
            bool ended = false, dep = false;
            int bin = -1, ear = 0, primary = 0;
            if (parked && have && pending) {
                pending = false;
                if (MODE == 1) p.pc_org_t[(size_t)(nseg - 1) * (size_t)p.pc_stride + (size_t)ray] = make_float4(org.x, org.y, org.z, h.t);
                if (h.slot < 0) {
                    ended = true;                                                               // miss :186-190
                } else {
                    const float4 a = __ldg(tris + h.slot * 3 + 0);
                    const float4 b4 = __ldg(tris + h.slot * 3 + 1);
                    const float4 c4 = __ldg(tris + h.slot * 3 + 2);
                    const F3 p1 = f3(a.x, a.y, a.z), p2 = f3(b4.x, b4.y, b4.z), p3 = f3(c4.x, c4.y, c4.z);
                    const int mat = __float_as_int(b4.w);
                    const F3 pt = hit_point(p1, p2, p3, h.u, h.v);
                    const F3 dp = sub3(pt, org);
                    dist = __fadd_rn(dist, __fsqrt_rn(dot3(dp, dp)));                           // :83
                    if (mat < 0) {
                        bin = receiver_hit<NB>(p, pt, dir, dist, energy);
                        ear = (mat == -1) ? 1 : 2;
                        primary = (mat == -1) ? 0 : 1;
                        dep = bin >= 0 && bin < p.ir_len;
                        ended = true;                                                           // :147,:169
                    } else {
                        const F3 nc = cross3(sub3(p2, p1), sub3(p3, p1));                       // :75-77
                        const float ninv = __fdiv_rn(1.0f, __fsqrt_rn(dot3(nc, nc)));
                        const F3 ng = f3(__fmul_rn(nc.x, ninv), __fmul_rn(nc.y, ninv), __fmul_rn(nc.z, ninv));
                        bool diffuse = false;
                        uint32_t r[4];
                        if (p.any_scatter) {
                            const float sc = __ldg(p.scattering + mat);
                            if (sc > 0.f) {
                                philox4x32(p.seed, (uint64_t)(p.ray_begin + ray), (uint32_t)depth, 1u, r);
                                diffuse = __fmul_rn((float)(r[0] >> 8), 0x1p-24f) < sc;
                            }
                        }
                        if (diffuse) {
                            dir = lambert_direction(r, dir, ng);
                        } else {
                            const float k = __fmul_rn(2.0f, dot3(dir, ng));                     // :173
                            dir = f3(__fmaf_rn(-k, ng.x, dir.x), __fmaf_rn(-k, ng.y, dir.y), __fmaf_rn(-k, ng.z, dir.z));
                        }
#pragma unroll
                        for (int b = 0; b < NB; ++b) energy[b] = __fmul_rn(energy[b], __ldg(p.keep + mat * NB + b)); // :174
                        depth++;                                                                // :175
                        org = f3(__fmaf_rn(1e-3f, dir.x, pt.x), __fmaf_rn(1e-3f, dir.y, pt.y), __fmaf_rn(1e-3f, dir.z, pt.z)); // :179
                    }
                }
            }
            if (MODE == 0) deposit_warp<NB>(p, dep, bin, primary, energy);

            // loop guard for the paths that go on (:233-236)
            if (parked && have && !ended) {
                float emax = energy[0];
#pragma unroll
                for (int b = 1; b < NB; ++b) emax = fmaxf(emax, energy[b]);
                if (!(dist < p.dist_thr && emax > p.energy_thres && (unsigned)depth < p.max_bounces)) ended = true;
            }
            if (parked && have && ended) {
                if (p.rec_bin) p.rec_bin[ray] = bin;
                if (p.rec_ear) p.rec_ear[ray] = ear;
                if (p.rec_nseg) p.rec_nseg[ray] = nseg;
                if (p.rec_energy) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) p.rec_energy[ray * NB + b] = ear ? energy[b] : 0.f;
                }
                if (MODE == 1) p.pc_nseg[ray] = nseg;
                segs += (unsigned long long)nseg;
                have = false;
            }
            // refill parked lanes without a path from the warp's chunk (ballot/popc compaction)
            unsigned need = __ballot_sync(FULL, parked && !have);
            while (need) {
                if (chunk_next >= chunk_end) {
                    unsigned long long b = 0;
                    if (lane == 0) b = atomicAdd(p.counters, (unsigned long long)kChunk);
                    b = __shfl_sync(FULL, b, 0);
                    chunk_next = (long long)b;
                    chunk_end = min((long long)b + kChunk, p.n_rays);
                    if (chunk_next >= chunk_end) {
                        if (parked && !have) exhausted = true;
                        break;
                    }
                }
                const int avail = (int)min((long long)32, chunk_end - chunk_next);
                const int rank = __popc(need & ((1u << lane) - 1u));
                if (((need >> lane) & 1u) && rank < avail) {
                    ray = chunk_next + rank;
                    have = true;
                    org = f3(p.emitter[0], p.emitter[1], p.emitter[2]);          // :210
                    dir = emit_direction(p.seed, (uint64_t)(p.ray_begin + ray)); // :216-224
#pragma unroll
                    for (int b = 0; b < NB; ++b) energy[b] = p.energy0;          // :208
                    dist = 0.f; depth = 0; nseg = 0;                             // :209,:211
                    // a fresh path can only fail the guard through its parameters (or a zero direction, :230)
                    const bool zero_dir = !(dir.x != 0.f || dir.y != 0.f || dir.z != 0.f);
                    if (zero_dir || !(0.f < p.dist_thr && p.energy0 > p.energy_thres && 0u < p.max_bounces)) {
                        if (p.rec_bin) p.rec_bin[ray] = -1;
                        if (p.rec_ear) p.rec_ear[ray] = 0;
                        if (p.rec_nseg) p.rec_nseg[ray] = 0;
                        if (p.rec_energy) {
#pragma unroll
                            for (int b = 0; b < NB; ++b) p.rec_energy[ray * NB + b] = 0.f;
                        }
                        if (MODE == 1) p.pc_nseg[ray] = 0;
                        have = false;                                            // picks another ray in the next round
                    }
                }
                chunk_next += min(__popc(need), avail);
                need = __ballot_sync(FULL, parked && !have && !exhausted);
            }
            // start the next segment
            if (parked && have) {
                if (MODE == 1) {
                    const size_t ci = (size_t)nseg * (size_t)p.pc_stride + (size_t)ray;
                    p.pc_dir_d[ci] = make_float4(dir.x, dir.y, dir.z, dist);
#pragma unroll
                    for (int b = 0; b < NB; ++b) p.pc_energy[ci * NB + b] = energy[b];
                }
                nseg++;
                ix = safe_rcp(dir.x); iy = safe_rcp(dir.y); iz = safe_rcp(dir.z);
                ox = org.x * ix; oy = org.y * iy; oz = org.z * iz;
                stack[0] = kSentinel; sp = 1;
                cur = p.root;
                h.t = 1e20f; h.u = 0.f; h.v = 0.f; h.slot = -1; h.id = INT_MAX;
                pending = true;
            }
        } else if (leaf_m != 0 && (__popc(leaf_m) >= ARV2_TL || inner_m == 0)) {
            // ======================= phase L: one leaf per lane
            if (at_leaf) {
                const int code = ~cur;
                const int first = code >> 3;
                const int cnt = (code & 7) + 1;
                for (int i = 0; i < cnt; ++i) {
                    const int slot = first + i;
                    const float4 a = __ldg(tris + slot * 3 + 0);
                    const float4 b = __ldg(tris + slot * 3 + 1);
                    const float4 c = __ldg(tris + slot * 3 + 2);
                    float t, u, v;
                    if (tri_test(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), org, dir, &t, &u, &v)) {
                        const int id = __float_as_int(a.w);
                        if (t < h.t || (t == h.t && id < h.id)) { h.t = t; h.u = u; h.v = v; h.slot = slot; h.id = id; }
                    }
                }
                cur = stack[--sp];
            }
        } else {
            // ======================= phase I: a burst of inner-node steps
#pragma unroll 1
            for (int k = 0; k < ARV2_BURST; ++k) {
                if (cur >= 0) {
                    const F8 na = ldg256(nodes + cur * 4), nb = ldg256(nodes + cur * 4 + 2);
                    const float4 n0 = na.lo, n1 = na.hi, n2 = nb.lo, n3 = nb.hi;
                    const float c0lox = fmaf(n0.x, ix, -ox), c0hix = fmaf(n0.y, ix, -ox);
                    const float c0loy = fmaf(n0.z, iy, -oy), c0hiy = fmaf(n0.w, iy, -oy);
                    const float c0loz = fmaf(n2.x, iz, -oz), c0hiz = fmaf(n2.y, iz, -oz);
                    const float c1lox = fmaf(n1.x, ix, -ox), c1hix = fmaf(n1.y, ix, -ox);
                    const float c1loy = fmaf(n1.z, iy, -oy), c1hiy = fmaf(n1.w, iy, -oy);
                    const float c1loz = fmaf(n2.z, iz, -oz), c1hiz = fmaf(n2.w, iz, -oz);
                    const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.f));
                    const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), h.t));
                    const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.f));
                    const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), h.t));
                    const bool go0 = c0min <= c0max, go1 = c1min <= c1max;
                    const int i0 = __float_as_int(n3.x), i1 = __float_as_int(n3.y);
                    if (!go0 && !go1) {
                        cur = stack[--sp];
                    } else {
                        cur = go0 ? i0 : i1;
                        if (go0 && go1) {
                            int far = i1;
                            if (c1min < c0min) { cur = i1; far = i0; }
                            stack[sp++] = far;
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(FULL, segs, o);
    if (lane == 0 && segs) atomicAdd(p.counters + 1, segs);
}

// Receiver move: walk each ray's cached receiver-independent segments in order and
// deposit at the first one the receiver intercepts before the wall (t_recv < t_wall;
// ties go to the scene because scene triangle ids are lower).  One thread per ray,
// lanes = consecutive rays, so every load of segment k is a coalesced 512 B row.
// A segment is first tested against the receiver's bounding ball (a few FMAs, exact-
// conservative); lanes whose segment passes park it and keep waiting until enough lanes of
// the warp hold a candidate (or nobody is scanning), then those lanes walk the receiver
// sub-tree together -- the r01 profile had this traversal running at 2.9 of 32 lanes.
template <int NB>
__global__ void __launch_bounds__(kRerenderThreads) rerender_kernel(const TraceParams p)
{
    const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = ray < p.n_rays;
    const int n = valid ? p.pc_nseg[ray] : 0;
    const F3 ctr = f3(p.center[0], p.center[1], p.center[2]);
    const float r2 = p.recv_radius * p.recv_radius;
    int k = 0;
    bool done = n == 0, cand = false;
    float4 ot = make_float4(0, 0, 0, 0), dd = make_float4(0, 0, 0, 0);
    int rbin = -1, rear = 0, rnseg = n;
    float energy[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) energy[b] = 0.f;
    for (;;) {
        // ---- scan: advance to the next segment that enters the bounding ball
        if (!done && !cand) {
#pragma unroll 1
            for (int burst = 0; burst < 16 && !cand && !done; ++burst) {
                const size_t ci = (size_t)k * (size_t)p.pc_stride + (size_t)ray;
                ot = __ldcs(p.pc_org_t + ci);
                dd = __ldcs(p.pc_dir_d + ci);
                const float ocx = ot.x - ctr.x, ocy = ot.y - ctr.y, ocz = ot.z - ctr.z;
                const float b = ocx * dd.x + ocy * dd.y + ocz * dd.z;
                const float c = ocx * ocx + ocy * ocy + ocz * ocz - r2;
                const float d2 = dd.x * dd.x + dd.y * dd.y + dd.z * dd.z;
                const float disc = b * b - d2 * c;
                // enters the ball at t0 = (-b - sqrt(disc))/d2, leaves at t1; needs t1 >= 0 and t0 <= t_wall
                bool pass = false;
                if (disc >= 0.f) {
                    const float sq = sqrtf(disc);
                    pass = (-b + sq) >= 0.f && (-b - sq) <= ot.w * d2;
                }
                if (pass) cand = true;
                else if (++k >= n) done = true;
            }
        }
        const unsigned cm = __ballot_sync(FULL, cand);
        const unsigned sm = __ballot_sync(FULL, !done && !cand);
        if (cm == 0 && sm == 0) break;
        if (cm != 0 && (__popc(cm) >= kRerenderBatch || sm == 0)) {
            bool dep = false;
            int bin = -1, primary = 0;
            if (cand) {
                const F3 org = f3(ot.x, ot.y, ot.z), dir = f3(dd.x, dd.y, dd.z);
                Hit h;
                closest_hit(p.nodes, p.tris, p.recv_root, org, dir, ot.w, h);
                if (h.slot >= 0 && h.t < ot.w) {
                    const float4 a = __ldg(p.tris + h.slot * 3 + 0);
                    const float4 b4 = __ldg(p.tris + h.slot * 3 + 1);
                    const float4 c4 = __ldg(p.tris + h.slot * 3 + 2);
                    const int mat = __float_as_int(b4.w);
                    const F3 pt = hit_point(f3(a.x, a.y, a.z), f3(b4.x, b4.y, b4.z), f3(c4.x, c4.y, c4.z), h.u, h.v);
                    const F3 dp = sub3(pt, org);
                    const float dist = __fadd_rn(dd.w, __fsqrt_rn(dot3(dp, dp)));
                    const size_t ci = (size_t)k * (size_t)p.pc_stride + (size_t)ray;
#pragma unroll
                    for (int b = 0; b < NB; ++b) energy[b] = p.pc_energy[ci * NB + b];
                    bin = receiver_hit<NB>(p, pt, dir, dist, energy);
                    primary = (mat == -1) ? 0 : 1;
                    dep = bin >= 0 && bin < p.ir_len;
                    rbin = bin; rear = (mat == -1) ? 1 : 2; rnseg = k + 1;
                    done = true;
                } else if (++k >= n) {
                    done = true;
                }
                cand = false;
            }
            deposit_warp<NB>(p, dep, bin, primary, energy);
        }
    }
    if (valid) {
        if (p.rec_bin) p.rec_bin[ray] = rbin;
        if (p.rec_ear) p.rec_ear[ray] = rear;
        if (p.rec_nseg) p.rec_nseg[ray] = rnseg;
        if (p.rec_energy) {
#pragma unroll
            for (int b = 0; b < NB; ++b) p.rec_energy[ray * NB + b] = rear ? energy[b] : 0.f;
        }
    }
    unsigned long long segs = valid ? (unsigned long long)rnseg : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(FULL, segs, o);
    if ((threadIdx.x & 31) == 0 && segs) atomicAdd(p.counters + 1, segs);
}

__global__ void finalize_kernel(const double* __restrict__ hist, int n, int mono, float* __restrict__ l, float* __restrict__ r)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = __double2float_rn(hist[i]), b = __double2float_rn(hist[n + i]);
    if (mono) { const float s = __fadd_rn(a, b); l[i] = s; r[i] = s; }
    else { l[i] = a; r[i] = b; }
}

template <int NB, int MODE>
cudaError_t launch_trace_t(const TraceParams& p, int sm_count, cudaStream_t stream)
{
    int per_sm = 0;
#ifdef ARV2_TRACE_V1
    auto kernel = trace_kernel<NB, MODE>;
#else
    auto kernel = trace2_kernel<NB, MODE>;
#endif
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = (p.n_rays + kThreads - 1) / kThreads;
    long long grid = (long long)sm_count * per_sm;
    if (want < grid) grid = want < 1 ? 1 : want;
    kernel<<<(unsigned)grid, kThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

} // namespace

cudaError_t launch_trace(const TraceParams& p, int bands, int mode, int sm_count, cudaStream_t stream)
{
    if (bands == 1) return mode == 0 ? launch_trace_t<1, 0>(p, sm_count, stream) : launch_trace_t<1, 1>(p, sm_count, stream);
    if (bands == 8) return mode == 0 ? launch_trace_t<8, 0>(p, sm_count, stream) : launch_trace_t<8, 1>(p, sm_count, stream);
    return cudaErrorInvalidValue;
}

cudaError_t launch_rerender(const TraceParams& p, int bands, int sm_count, cudaStream_t stream)
{
    (void)sm_count;
    const unsigned grid = (unsigned)((p.n_rays + kRerenderThreads - 1) / kRerenderThreads);
    if (grid == 0) return cudaSuccess;
    if (bands == 1) rerender_kernel<1><<<grid, kRerenderThreads, 0, stream>>>(p);
    else if (bands == 8) rerender_kernel<8><<<grid, kRerenderThreads, 0, stream>>>(p);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_finalize(const double* hist, int bands, int ir_len, int mono, float* ir_left, float* ir_right,
                            cudaStream_t stream)
{
    const int n = bands * ir_len;
    finalize_kernel<<<(n + 255) / 256, 256, 0, stream>>>(hist, n, mono, ir_left, ir_right);
    return cudaGetLastError();
}

} // namespace arv2

// trace.cu -- sm_100a sound-ray tracing kernels of libarv2.
//
// Replaces the OptiX pipeline of the reference (B200 has no RT cores):
//   __raygen__renderFrame     OR/devicePrograms.cu:192-254  -> trace_kernel (path loop)
//   optixTrace / GAS traversal OR/devicePrograms.cu:240-251 -> closest_hit (software BVH)
//   __closesthit__radiance    OR/devicePrograms.cu:62-180   -> shade_segment
//   __miss__radiance          OR/devicePrograms.cu:186-190  -> "no hit" branch
//   fillZeros / addIRs        OR/kernels.cu:77-97,519-536   -> cudaMemsetAsync / finalize_kernel
//
// Design (details at each kernel):
//   wave_kernel      the tracer: persistent, one 32-warp CTA per SM, breadth-first.  A task = 32 paths of one depth
//                    advanced by 8 segments; survivors are re-packed (ballot / popc) into per-SM, per-depth queues
//                    whose counters live in shared memory; rays are started in the order of their emission direction
//                    (direction_keys_kernel + radix sort) so a warp is a coherent bundle for the first bounces.
//                    trace_kernel = depth-first fallback (no queue memory / A-B), trace2_kernel = the decoupled-lane
//                    experiment of r03.
//   sweep_kernel     the tracer of launches of >= 3 M rays: bounce-synchronous, one launch per sweep over ALL paths alive
//                    (8 segments for the fresh bundles, then 2 per sweep); the survivors are compacted into a second state
//                    buffer and re-binned by (cell of the new origin, octahedral cell of the new direction) with a counting
//                    sort whose histogram the sweep takes itself, so a warp is a bundle at EVERY depth (18.0 instead of
//                    15.4 lanes per node step and +12 % on the 1M-triangle scene, +19 % at 100 M rays; profiles/r09).
//   closest_hit      software BVH: binary 64 B nodes with both child boxes in the parent, every record fetched with
//                    256-bit loads (a divergent gather costs the L1 data pipe one wavefront per lane and load
//                    instruction; the whole scene is L2-resident and DRAM idles).  Quantised 32 B nodes and 4-wide
//                    nodes were built and measured slower (profiles/r03, r07 sections 18 and 22).
//   deposits         aggregated across the warp (match.any) and accumulated in an fp64 histogram with native RED.F64:
//                    the result does not depend on the deposit order to ~1e-16.
//   re-render        rr_mask_kernel + rr_walk_kernel over the packed path cache (pc_* kernels build it).
#include <climits>
#include <cstdlib>

#include "arv2_internal.h"
#include "arv2_model.cuh"
#include "trace.cuh"

namespace arv2 {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kChunk = 128;          // rays a re-render warp claims per global atomic
constexpr int kStack = kTraversalStack;
#ifndef ARV2_THREADS
#define ARV2_THREADS 128
#endif
#ifndef ARV2_MINB
#define ARV2_MINB 9                  // 56 registers, 36 warps/SM for one band (no spills); 8 bands: 64 registers, 32 warps/SM
#endif
#ifndef ARV2_MINB8
#define ARV2_MINB8 8
#endif
constexpr int kThreads = ARV2_THREADS;
constexpr int kSentinel = INT_MIN;   // "nothing (left) to traverse"
constexpr int kRerenderThreads = 256;

struct Hit { float t, u, v; int slot, id; };

// 256-bit read-only global load (sm_100+: LDG.E.ENL2.256.CONSTANT).
struct __align__(32) F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg256(const float4* p)
{
    F8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
}

// the same load with an L1 eviction priority: triangle records (read once per leaf visit, two thirds of the scene's
// bytes) must not push the nodes out of L1.  ARV2_TRI_POLICY 0 = default, 1 = L1::no_allocate (kept: +2 %, r07
// section 9), 2 = L1::evict_first; ARV2_NODE_POLICY 1 = L1::evict_last for the nodes (+0.8 %, within noise)
#ifndef ARV2_TRI_POLICY
#define ARV2_TRI_POLICY 1
#endif
#ifndef ARV2_WIDE
#define ARV2_WIDE 0                  // 1: the kernels also walk 4-wide nodes (TraceParams::nodes4, experiment r07 section 18)
#endif
#ifndef ARV2_SHADE_NOALLOC
#define ARV2_SHADE_NOALLOC 1           // the shading re-read of the hit record does not allocate in L1 either (+1 %, r07 section 9)
#endif
#ifndef ARV2_NODE_POLICY
#define ARV2_NODE_POLICY 0           // 1 = L1::evict_last
#endif
__device__ __forceinline__ F8 ldg256_tri(const float4* p)
{
#if ARV2_TRI_POLICY == 0
    return ldg256(p);
#else
    F8 r;
#if ARV2_TRI_POLICY == 1
    asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm("ld.global.nc.L1::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
#endif
}
__device__ __forceinline__ float4 ldg128_tri(const float4* p)
{
#if ARV2_TRI_POLICY == 0
    return __ldg(p);
#else
    float4 r;
#if ARV2_TRI_POLICY == 1
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
#else
    asm("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];"
#endif
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#endif
}
__device__ __forceinline__ F8 ldg256_node(const float4* p)
{
#if ARV2_NODE_POLICY == 0
    return ldg256(p);
#else
    F8 r;
    asm("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
#endif
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

__device__ __forceinline__ void stg256(float4* p, float4 lo, float4 hi)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
}

__device__ __forceinline__ float safe_rcp(float d)
{
    const float eps = 1e-20f;
    return 1.0f / (fabsf(d) > eps ? d : copysignf(eps, d));
}

// Per-segment ray constants of the slab test: t(plane) = plane * (1/dir) - org/dir.
#ifndef ARV2_QNODES
#define ARV2_QNODES 0                // 1: the scene tree is walked through 32 B quantised nodes (TraceParams::nodes4, r07 section 22)
#endif
#if ARV2_QNODES
// Quantised nodes: a plane is stored as a 15-bit integer q on a scene-wide grid; PRMT turns it into the float
// f = 2^23 + 256 q without a conversion, and t(plane) = f * s + b with s = qk / dir and b = qc / dir - org / dir per axis.
// The float nodes of the top level and of the receiver tree recover 1 / dir as s * qinvk.
struct RayGrid {
    float sx, sy, sz, bx, by, bz;
    unsigned sgn;
    __device__ __forceinline__ void setup(const TraceParams& p, F3 org, F3 dir)
    {
        const float ix = safe_rcp(dir.x), iy = safe_rcp(dir.y), iz = safe_rcp(dir.z);
        sx = ix * p.qk[0]; sy = iy * p.qk[1]; sz = iz * p.qk[2];
        bx = fmaf(p.qc[0], ix, -(org.x * ix)); by = fmaf(p.qc[1], iy, -(org.y * iy)); bz = fmaf(p.qc[2], iz, -(org.z * iz));
        sgn = 0u;
    }
};
#else
struct RayGrid {
    float ix, iy, iz, ox, oy, oz;
    unsigned sgn;                    // bit a: the ray travels towards smaller coordinates along axis a (wide nodes)
    __device__ __forceinline__ void setup(const TraceParams&, F3 org, F3 dir)
    {
        ix = safe_rcp(dir.x); iy = safe_rcp(dir.y); iz = safe_rcp(dir.z);
        ox = org.x * ix; oy = org.y * iy; oz = org.z * iz;
        sgn = (ix < 0.f ? 1u : 0u) | (iy < 0.f ? 2u : 0u) | (iz < 0.f ? 4u : 0u);
    }
};
#endif
constexpr int kNone = INT_MIN + 1;   // "this child is not entered" (never a node, a leaf code or the sentinel)

// Traversal of one tree.  Closest hit = min (t, global triangle id) over all triangles
// whose exact test accepts; the tree only prunes (boxes are padded and quantised outwards,
// comparisons are <=), so any tree gives the same answer.  The stack is a separate local
// array on purpose: inside a struct it would drag the scalars into local memory with it.
struct Traversal {
    int sp, cur;
    Hit h;
    // -DARV2_TRACE_STATS (libarv2_stats.so, bench.py's traversal figures): per-lane tallies of the walk
#ifdef ARV2_TRACE_STATS
    unsigned long long st_inner = 0, st_wsteps = 0, st_leaf = 0, st_tri = 0, st_wleaf = 0;
    __device__ __forceinline__ void tally_inner() { ++st_inner; const unsigned m = __activemask(); if ((threadIdx.x & 31) == __ffs(m) - 1) ++st_wsteps; }
    __device__ __forceinline__ void tally_leaf(int cnt) { ++st_leaf; st_tri += cnt; const unsigned m = __activemask(); if ((threadIdx.x & 31) == __ffs(m) - 1) ++st_wleaf; }
    __device__ __forceinline__ void flush_stats(unsigned long long* counters)
    {
        unsigned long long v[5] = {st_inner, st_wsteps, st_leaf, st_tri, st_wleaf};
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
            if ((threadIdx.x & 31) == 0 && v[i]) atomicAdd(counters + kStatInner + i, v[i]);
        }
    }
#else
    __device__ __forceinline__ void tally_inner() {}
    __device__ __forceinline__ void tally_leaf(int) {}
    __device__ __forceinline__ void flush_stats(unsigned long long*) {}
#endif

    __device__ __forceinline__ void reset(float tmax) { h.t = tmax; h.u = 0.f; h.v = 0.f; h.slot = -1; h.id = INT_MAX; cur = kSentinel; sp = 0; }
    __device__ __forceinline__ void enter(int* stack, int root) { stack[0] = kSentinel; sp = 1; cur = root; }
    __device__ __forceinline__ bool at_inner() const { return cur >= 0; }
    __device__ __forceinline__ bool at_leaf() const { return cur < 0 && cur != kSentinel; }
    __device__ __forceinline__ bool finished() const { return cur == kSentinel; }

    // one binary node (64 B = two 256-bit loads, both child boxes in the parent): slab tests,
    // descend into the nearer hit child, push the other
#if ARV2_WIDE
    // one 4-wide node (128 B = four 256-bit loads): four slab tests, the slots ordered front to back from the signs
    // of the ray direction and the node's three split axes (no sorting network), the first entered child next, the
    // others pushed
    __device__ __forceinline__ void step_wide(int* stack, const float4* __restrict__ nodes4, const RayGrid& g)
    {
        const float4* n = nodes4 + (size_t)(cur & (kWideBit - 1)) * 8;
        const F8 X = ldg256_node(n), Y = ldg256_node(n + 2), Z = ldg256_node(n + 4), W = ldg256_node(n + 6);
#define ARV2_SLAB(LOX, HIX, LOY, HIY, LOZ, HIZ, CODE, OUT)                                                                  \
        {                                                                                                                  \
            const float ax = fmaf(LOX, g.ix, -g.ox), bx = fmaf(HIX, g.ix, -g.ox);                                           \
            const float ay = fmaf(LOY, g.iy, -g.oy), by = fmaf(HIY, g.iy, -g.oy);                                           \
            const float az = fmaf(LOZ, g.iz, -g.oz), bz = fmaf(HIZ, g.iz, -g.oz);                                           \
            const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));                       \
            const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), h.t));                       \
            OUT = tmin <= tmax ? __float_as_int(CODE) : kNone;                                                              \
        }
        int c0, c1, c2, c3;
        ARV2_SLAB(X.lo.x, X.hi.x, Y.lo.x, Y.hi.x, Z.lo.x, Z.hi.x, W.lo.x, c0)
        ARV2_SLAB(X.lo.y, X.hi.y, Y.lo.y, Y.hi.y, Z.lo.y, Z.hi.y, W.lo.y, c1)
        ARV2_SLAB(X.lo.z, X.hi.z, Y.lo.z, Y.hi.z, Z.lo.z, Z.hi.z, W.lo.z, c2)
        ARV2_SLAB(X.lo.w, X.hi.w, Y.lo.w, Y.hi.w, Z.lo.w, Z.hi.w, W.lo.w, c3)
#undef ARV2_SLAB
        const int axes = __float_as_int(W.hi.x);
        const bool n0 = (g.sgn >> (axes & 3)) & 1u, nl = (g.sgn >> ((axes >> 2) & 3)) & 1u, nr = (g.sgn >> ((axes >> 4) & 3)) & 1u;
        int t;
        if (nl) { t = c0; c0 = c1; c1 = t; }
        if (nr) { t = c2; c2 = c3; c3 = t; }
        if (n0) { t = c0; c0 = c2; c2 = t; t = c1; c1 = c3; c3 = t; }
        const bool v0 = c0 != kNone, v1 = c1 != kNone, v2 = c2 != kNone, v3 = c3 != kNone;
        if (v3 && (v0 || v1 || v2)) stack[sp++] = c3;
        if (v2 && (v0 || v1)) stack[sp++] = c2;
        if (v1 && v0) stack[sp++] = c1;
        int next = v0 ? c0 : (v1 ? c1 : (v2 ? c2 : c3));
        if (!(v0 || v1 || v2 || v3)) next = stack[--sp];
        cur = next;
    }

#endif

    // the part of a binary node step that follows the twelve plane distances: slab intervals, descend into the nearer
    // entered child, push the other, or pop
    __device__ __forceinline__ void descend(int* stack, float c0lox, float c0hix, float c0loy, float c0hiy, float c0loz, float c0hiz,
                                            float c1lox, float c1hix, float c1loy, float c1hiy, float c1loz, float c1hiz, int i0, int i1)
    {
        const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.f));
        const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), h.t));
        const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.f));
        const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), h.t));
        const bool go0 = c0min <= c0max, go1 = c1min <= c1max;
        // predicated push / pop instead of a divergent if-else: lanes that pop and lanes that descend share the
        // same instructions
        const bool both = go0 && go1;
        const bool swap = both && c1min < c0min;
        int next = (go0 && !swap) ? i0 : i1;
        if (both) stack[sp] = swap ? i0 : i1;
        sp += both ? 1 : 0;
        if (!(go0 || go1)) next = stack[--sp];
        cur = next;
    }

#if ARV2_QNODES
    // one quantised binary node (32 B = one sector, one 256-bit load): six words of two 15-bit planes (lo | hi << 16) for
    // (child, axis) = (0,x) (0,y) (0,z) (1,x) (1,y) (1,z), then the two child codes
    __device__ __forceinline__ void step_q(int* stack, const float4* __restrict__ nodesq, const RayGrid& g)
    {
        const F8 n = ldg256_node(nodesq + (size_t)(cur & (kWideBit - 1)) * 2);
#define ARV2_QLO(w) __uint_as_float(__byte_perm(__float_as_uint(w), 0x4B000000u, 0x7104))
#define ARV2_QHI(w) __uint_as_float(__byte_perm(__float_as_uint(w), 0x4B000000u, 0x7324))
        descend(stack,
                fmaf(ARV2_QLO(n.lo.x), g.sx, g.bx), fmaf(ARV2_QHI(n.lo.x), g.sx, g.bx), fmaf(ARV2_QLO(n.lo.y), g.sy, g.by), fmaf(ARV2_QHI(n.lo.y), g.sy, g.by),
                fmaf(ARV2_QLO(n.lo.z), g.sz, g.bz), fmaf(ARV2_QHI(n.lo.z), g.sz, g.bz),
                fmaf(ARV2_QLO(n.lo.w), g.sx, g.bx), fmaf(ARV2_QHI(n.lo.w), g.sx, g.bx), fmaf(ARV2_QLO(n.hi.x), g.sy, g.by), fmaf(ARV2_QHI(n.hi.x), g.sy, g.by),
                fmaf(ARV2_QLO(n.hi.y), g.sz, g.bz), fmaf(ARV2_QHI(n.hi.y), g.sz, g.bz), __float_as_int(n.hi.z), __float_as_int(n.hi.w));
#undef ARV2_QLO
#undef ARV2_QHI
    }
#endif

    // one binary node (64 B = two 256-bit loads, both child boxes in the parent): slab tests,
    // descend into the nearer hit child, push the other
    __device__ __forceinline__ void step_inner(int* stack, const TraceParams& p, const RayGrid& g, F3 org)
    {
#if ARV2_WIDE
        if (cur & kWideBit) { step_wide(stack, p.nodes4, g); return; }
#endif
#if ARV2_QNODES
        if (cur & kWideBit) { step_q(stack, p.nodes4, g); return; }
        const float ix = g.sx * p.qinvk[0], iy = g.sy * p.qinvk[1], iz = g.sz * p.qinvk[2];
        const float ox = org.x * ix, oy = org.y * iy, oz = org.z * iz;
#else
        (void)org;
        const float ix = g.ix, iy = g.iy, iz = g.iz, ox = g.ox, oy = g.oy, oz = g.oz;
#endif
        tally_inner();
        const float4* __restrict__ nodes = p.nodes;
        const F8 na = ldg256_node(nodes + cur * 4), nb = ldg256_node(nodes + cur * 4 + 2);
        const float4 n0 = na.lo, n1 = na.hi, n2 = nb.lo, n3 = nb.hi;
        descend(stack,
                fmaf(n0.x, ix, -ox), fmaf(n0.y, ix, -ox), fmaf(n0.z, iy, -oy), fmaf(n0.w, iy, -oy), fmaf(n2.x, iz, -oz), fmaf(n2.y, iz, -oz),
                fmaf(n1.x, ix, -ox), fmaf(n1.y, ix, -ox), fmaf(n1.z, iy, -oy), fmaf(n1.w, iy, -oy), fmaf(n2.z, iz, -oz), fmaf(n2.w, iz, -oz),
                __float_as_int(n3.x), __float_as_int(n3.y));
    }

    // the same step on a copy of a (small) tree in shared memory: `nodes` points at the copy of node `first`
    __device__ __forceinline__ void step_inner_shared(int* stack, const float4* nodes, int first, const RayGrid& g)
    {
#if ARV2_QNODES
        (void)stack; (void)nodes; (void)first; (void)g;
#else
        const float ix = g.ix, iy = g.iy, iz = g.iz, ox = g.ox, oy = g.oy, oz = g.oz;
        const float4* n = nodes + (cur - first) * 4;
        const float4 n0 = n[0], n1 = n[1], n2 = n[2], n3 = n[3];
        descend(stack,
                fmaf(n0.x, ix, -ox), fmaf(n0.y, ix, -ox), fmaf(n0.z, iy, -oy), fmaf(n0.w, iy, -oy), fmaf(n2.x, iz, -oz), fmaf(n2.y, iz, -oz),
                fmaf(n1.x, ix, -ox), fmaf(n1.y, ix, -ox), fmaf(n1.z, iy, -oy), fmaf(n1.w, iy, -oy), fmaf(n2.z, iz, -oz), fmaf(n2.w, iz, -oz),
                __float_as_int(n3.x), __float_as_int(n3.y));
#endif
    }

    // one leaf: exact tests of its <= 8 triangles; the test needs the first 48 B of a record (P1, id, P2, material, P3)
    // (CACHED: the records stay in L1 -- the 1020 receiver triangles of a re-render; the scene's are read once per visit)
    template <bool CACHED = false>
    __device__ __forceinline__ void step_leaf(int* stack, const float4* __restrict__ tris, F3 org, F3 dir)
    {
        const int code = ~cur;
        const int first = code >> 3;
        const int cnt = (code & 7) + 1;
        tally_leaf(cnt);
        for (int i = 0; i < cnt; ++i) {
            const int slot = first + i;
            const F8 A = CACHED ? ldg256(tris + slot * 4) : ldg256_tri(tris + slot * 4);
            const float4 B = CACHED ? __ldg(tris + slot * 4 + 2) : ldg128_tri(tris + slot * 4 + 2);
            float t, u, v;
            if (tri_test(f3(A.lo.x, A.lo.y, A.lo.z), f3(A.hi.x, A.hi.y, A.hi.z), f3(B.x, B.y, B.z), org, dir, &t, &u, &v)) {
                const int id = __float_as_int(A.lo.w);
                if (t < h.t || (t == h.t && id < h.id)) { h.t = t; h.u = u; h.v = v; h.slot = slot; h.id = id; }
            }
        }
        cur = stack[--sp];
    }

    __device__ __forceinline__ void walk(int* stack, const TraceParams& p, int root, const RayGrid& g, F3 org, F3 dir)
    {
        const float4* __restrict__ tris = p.tris;
        enter(stack, root);
        while (!finished()) {
            while (at_inner()) step_inner(stack, p, g, org);
            if (finished()) break;
            step_leaf(stack, tris, org, dir);
        }
    }
};

// Does the segment org + t*dir, 0 <= t <= tmax, enter the ball that contains the placed
// receiver mesh?  Exact-conservative (the radius is padded on the host).
__device__ __forceinline__ bool enters_receiver_ball(const TraceParams& p, F3 org, F3 dir, float tmax)
{
    const float ocx = org.x - p.center[0], ocy = org.y - p.center[1], ocz = org.z - p.center[2];
    const float b = ocx * dir.x + ocy * dir.y + ocz * dir.z;
    const float c = ocx * ocx + ocy * ocy + ocz * ocz - p.recv_radius * p.recv_radius;
    const float d2 = dir.x * dir.x + dir.y * dir.y + dir.z * dir.z;
    const float disc = b * b - d2 * c;
    if (!(disc >= 0.f)) return false;
    const float sq = sqrtf(disc);
    return (-b + sq) >= 0.f && (-b - sq) <= tmax * d2;
}

// optixTrace: closest hit below `root` (the two-level top node for a full trace, the scene
// root while filling the path cache, the receiver root for a re-render).
__device__ __forceinline__ void closest_hit(const TraceParams& p, int* stack, Traversal& tr, int root, F3 org, F3 dir, float tmax)
{
    tr.reset(tmax);
    if (root < 0) return;
    RayGrid g;
    g.setup(p, org, dir);
    tr.walk(stack, p, root, g, org, dir);
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

// Hit point (OR/devicePrograms.cu:79-81).
__device__ __forceinline__ F3 hit_point(F3 p1, F3 p2, F3 p3, float u, float v)
{
    const float w = __fsub_rn(__fsub_rn(1.0f, u), v);
    return f3(__fmaf_rn(v, p3.x, __fmaf_rn(u, p2.x, __fmul_rn(w, p1.x))),
              __fmaf_rn(v, p3.y, __fmaf_rn(u, p2.y, __fmul_rn(w, p1.y))),
              __fmaf_rn(v, p3.z, __fmaf_rn(u, p2.z, __fmul_rn(w, p1.z))));
}

// Per-lane path state (struct PRD, OR/PRD.h:5-14, kept in registers).
template <int NB>
struct Path {
    long long ray;
    F3 org, dir;
    float energy[NB];
    float dist;
    int depth, nseg;
};

struct Deposit { bool dep; int bin, ear, primary; };

// __closesthit__radiance / __miss__radiance for a finished segment.  Returns true when the
// path ends (miss or receiver); a wall hit bounces the path in place.
template <int NB, int MODE>
__device__ __forceinline__ bool shade_segment(const TraceParams& p, Path<NB>& s, const Hit& h, Deposit& d)
{
    if (MODE == 1) {
        // path cache: the finished segment as one 32 B record (s.dir, s.dist and s.energy still describe its start)
        const size_t ci = (size_t)s.ray * (size_t)p.pc_stride + (size_t)(s.nseg - 1);
        stg256(p.pc_seg + 2 * ci, make_float4(s.org.x, s.org.y, s.org.z, h.t), make_float4(s.dir.x, s.dir.y, s.dir.z, s.dist));
#pragma unroll
        for (int b = 0; b < NB; ++b) p.pc_energy[ci * NB + b] = s.energy[b];
    }
    if (h.slot < 0) return true;                                                     // miss :186-190
#if ARV2_SHADE_NOALLOC
    const F8 A = ldg256_tri(p.tris + h.slot * 4), B = ldg256_tri(p.tris + h.slot * 4 + 2);
#else
    const F8 A = ldg256(p.tris + h.slot * 4), B = ldg256(p.tris + h.slot * 4 + 2);
#endif
    const F3 p1 = f3(A.lo.x, A.lo.y, A.lo.z), p2 = f3(A.hi.x, A.hi.y, A.hi.z), p3 = f3(B.lo.x, B.lo.y, B.lo.z);
    const int mat = __float_as_int(A.hi.w);
    const F3 pt = hit_point(p1, p2, p3, h.u, h.v);
    const F3 dp = sub3(pt, s.org);
    s.dist = __fadd_rn(s.dist, __fsqrt_rn(dot3(dp, dp)));                            // :83
    if (mat < 0) {
        d.bin = receiver_hit<NB>(p, pt, s.dir, s.dist, s.energy);
        d.ear = (mat == -1) ? 1 : 2;
        d.primary = (mat == -1) ? 0 : 1;
        d.dep = d.bin >= 0 && d.bin < p.ir_len;
        return true;                                                                  // :147,:169
    }
    const F3 ng = f3(B.lo.w, B.hi.x, B.hi.y);                                        // :75-77, precomputed on the host
    bool diffuse = false;
    uint32_t r[4];
    if (p.any_scatter) {
        const float sc = __ldg(p.scattering + mat);
        if (sc > 0.f) {
            philox4x32(p.seed, (uint64_t)(p.ray_begin + s.ray), (uint32_t)s.depth, 1u, r);
            diffuse = __fmul_rn((float)(r[0] >> 8), 0x1p-24f) < sc;
        }
    }
    if (diffuse) {
        s.dir = lambert_direction(r, s.dir, ng);
    } else {
        const float k = __fmul_rn(2.0f, dot3(s.dir, ng));                            // :173
        s.dir = f3(__fmaf_rn(-k, ng.x, s.dir.x), __fmaf_rn(-k, ng.y, s.dir.y), __fmaf_rn(-k, ng.z, s.dir.z));
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) s.energy[b] = __fmul_rn(s.energy[b], __ldg(p.keep + mat * NB + b));    // :174
    s.depth++;                                                                        // :175
    s.org = f3(__fmaf_rn(1e-3f, s.dir.x, pt.x), __fmaf_rn(1e-3f, s.dir.y, pt.y), __fmaf_rn(1e-3f, s.dir.z, pt.z)); // :179
    return false;
}

// loop guard of __raygen__renderFrame (:230, :233-236)
template <int NB>
__device__ __forceinline__ bool path_goes_on(const TraceParams& p, const Path<NB>& s)
{
    float emax = s.energy[0];
#pragma unroll
    for (int b = 1; b < NB; ++b) emax = fmaxf(emax, s.energy[b]);
    const bool zero_dir = !(s.dir.x != 0.f || s.dir.y != 0.f || s.dir.z != 0.f);
    return !zero_dir && s.dist < p.dist_thr && emax > p.energy_thres && (unsigned)s.depth < p.max_bounces;
}

template <int NB, int MODE>
__device__ __forceinline__ void end_path(const TraceParams& p, const Path<NB>& s, const Deposit& d, unsigned long long& segs)
{
    if (p.rec_bin) p.rec_bin[s.ray] = d.bin;
    if (p.rec_ear) p.rec_ear[s.ray] = d.ear;
    if (p.rec_nseg) p.rec_nseg[s.ray] = s.nseg;
    if (p.rec_energy) {
#pragma unroll
        for (int b = 0; b < NB; ++b) p.rec_energy[s.ray * NB + b] = d.ear ? s.energy[b] : 0.f;
    }
    if (MODE == 1) p.pc_nseg[s.ray] = s.nseg;
    segs += (unsigned long long)s.nseg;
}

template <int NB>
__device__ __forceinline__ void new_path(const TraceParams& p, Path<NB>& s, long long ray)
{
    if (p.ray_order) ray = p.ray_order[ray];
    s.ray = ray;
    s.org = f3(p.emitter[0], p.emitter[1], p.emitter[2]);              // :210
    s.dir = emit_direction(p.seed, (uint64_t)(p.ray_begin + ray));     // :216-224
#pragma unroll
    for (int b = 0; b < NB; ++b) s.energy[b] = p.energy0;              // :208
    s.dist = 0.f; s.depth = 0; s.nseg = 0;                             // :209,:211
}

template <int NB, int MODE>
__device__ __forceinline__ void begin_segment(const TraceParams& p, Path<NB>& s)
{
    s.nseg++;
}

// Warp-level refill of the lanes in `want` from the warp's chunk of the ray set
// (ballot / popc compaction).  Lanes that got a ray return true.
template <int NB>
__device__ __forceinline__ bool refill(const TraceParams& p, bool want, long long& chunk_next, long long& chunk_end,
                                       bool& exhausted, Path<NB>& s)
{
    const int lane = threadIdx.x & 31;
    bool got = false;
    unsigned need = __ballot_sync(FULL, want && !exhausted);
    while (need) {
        if (chunk_next >= chunk_end) {
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(p.counters, (unsigned long long)p.chunk);
            b = __shfl_sync(FULL, b, 0);
            chunk_next = (long long)b;
            chunk_end = min((long long)b + p.chunk, p.n_rays);
            if (chunk_next >= chunk_end) {
#ifdef ARV2_TAILSTAT
                if (lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMin(p.counters + 4, t); }
#endif
                if (want && !got) exhausted = true;
                break;
            }
        }
        const int avail = (int)min((long long)32, chunk_end - chunk_next);
        const int rank = __popc(need & ((1u << lane) - 1u));
        if (((need >> lane) & 1u) && rank < avail) {
            new_path<NB>(p, s, chunk_next + rank);
            got = true;
        }
        chunk_next += min(__popc(need), avail);
        need = __ballot_sync(FULL, want && !got && !exhausted);
    }
    return got;
}

// ---- path states in the per-depth queues of wave_kernel ----
// A path at a segment boundary is (ray, org, dir, dist, depth, energy[NB]); nseg == depth there
// (every finished segment either ends the path or bounces it once).
template <int NB>
__device__ __forceinline__ void store_path(float4* o, const Path<NB>& s)
{
    __stcg(o + 0, make_float4(s.org.x, s.org.y, s.org.z, s.dist));
    __stcg(o + 1, make_float4(s.dir.x, s.dir.y, s.dir.z, __int_as_float((int)s.ray)));
    if (NB == 1) {
        __stcg(o + 2, make_float4(s.energy[0], __int_as_float(s.depth), 0.f, 0.f));
    } else {
#pragma unroll
        for (int q = 0; q < NB / 4; ++q) __stcg(o + 2 + q, make_float4(s.energy[4 * q], s.energy[4 * q + 1], s.energy[4 * q + 2], s.energy[4 * q + 3]));
        __stcg(o + 2 + NB / 4, make_float4(__int_as_float(s.depth), 0.f, 0.f, 0.f));
    }
}

template <int NB>
__device__ __forceinline__ void load_path(const float4* i, Path<NB>& s)
{
    const float4 a = __ldcg(i), b = __ldcg(i + 1);
    s.org = f3(a.x, a.y, a.z); s.dist = a.w;
    s.dir = f3(b.x, b.y, b.z); s.ray = (long long)__float_as_int(b.w);
    if (NB == 1) {
        const float4 c = __ldcg(i + 2);
        s.energy[0] = c.x; s.depth = __float_as_int(c.y);
    } else {
#pragma unroll
        for (int q = 0; q < NB / 4; ++q) {
            const float4 c = __ldcg(i + 2 + q);
            s.energy[4 * q] = c.x; s.energy[4 * q + 1] = c.y; s.energy[4 * q + 2] = c.z; s.energy[4 * q + 3] = c.w;
        }
        s.depth = __float_as_int(__ldcg(i + 2 + NB / 4).x);
    }
    s.nseg = s.depth;
}

// One segment of one path (the body of the loop of __raygen__renderFrame, :233-252).
template <int NB, int MODE>
__device__ __forceinline__ bool advance_segment(const TraceParams& p, Path<NB>& s, int* stack, Traversal& tr, Deposit& d)
{
    if (!path_goes_on<NB>(p, s)) return true;                                         // :233-236
    begin_segment<NB, MODE>(p, s);
    // only segments whose line meets the receiver's bounding ball start at the two-level
    // top node; the others go straight into the scene tree (the top node's AABB of
    // the ball would let 3x as many lanes into the receiver tree)
    const bool to_recv = MODE == 0 && p.recv_root >= 0 && enters_receiver_ball(p, s.org, s.dir, 1e20f);
    closest_hit(p, stack, tr, to_recv ? p.root : p.scene_root, s.org, s.dir, 1e20f);
    return shade_segment<NB, MODE>(p, s, tr.h, d);
}

// ---------------------------------------------------------------------------------------
// trace_kernel: while-while traversal, all lanes of a warp advance segment by segment.
template <int NB, int MODE>
__global__ void __launch_bounds__(kThreads, NB == 1 ? ARV2_MINB : ARV2_MINB8) trace_kernel(const TraceParams p)
{
    const int lane = threadIdx.x & 31;
    long long chunk_next = 0, chunk_end = 0;      // warp-uniform
    bool have = false, exhausted = false;
    Path<NB> s;
    s.ray = 0; s.org = f3(0, 0, 0); s.dir = f3(0, 0, 0); s.dist = 0.f; s.depth = 0; s.nseg = 0;
    unsigned long long segs = 0;
    Traversal tr;
    int stack[kStack];

    for (;;) {
        const bool open = __popc(__ballot_sync(FULL, have)) < p.refill_below;
        if (refill<NB>(p, !have && open, chunk_next, chunk_end, exhausted, s)) have = true;
        if (!__any_sync(FULL, have)) break;

        bool ended = false;
        Deposit d; d.dep = false; d.bin = -1; d.ear = 0; d.primary = 0;
        if (have) ended = advance_segment<NB, MODE>(p, s, stack, tr, d);
        if (MODE == 0) deposit_warp<NB>(p, d.dep, d.bin, d.primary, s.energy);
        if (ended) {
            end_path<NB, MODE>(p, s, d, segs);
            have = false;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(FULL, segs, o);
    if (lane == 0 && segs) atomicAdd(p.counters + 1, segs);
    tr.flush_stats(p.counters);
#ifdef ARV2_TAILSTAT
    if (lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMax(p.counters + 5, t); atomicMin(p.counters + 6, t); }
#endif
}

// ---------------------------------------------------------------------------------------
// wave_kernel: breadth-first persistent tracer, one 32-warp CTA per SM.
// A task = up to 32 paths of the same depth advanced by wave_segments segments; the paths that go on
// are appended to the queue of their new depth.  A warp looking for work starts fresh rays while its
// SM has room (wave_cap paths alive) and otherwise takes a batch from the SHALLOWEST non-empty queue.
//   * The paths of a task were started together as neighbours of the direction order and have bounced
//     equally often: a warp stays a coherent bundle for as long as the scene allows, without idle
//     lanes, because the survivors are re-packed at every hand-over.
//   * The youngest paths always run first, so the render has no ramp-down.  In trace_kernel a lane
//     starts its last ray late and the warp then waits up to max_bounces segments for it: the last
//     2 ms of a 7.5 ms render (1M rays, profiles/r05) ran at a few live lanes per warp.
// Every SM owns its queues: positions and counters live in shared memory (reserve `tail`, in-order
// publish `pub`, claim `head` by CAS), only the 48 B path states go through L2.  The only global
// counter is the pool of unstarted rays.  wave_cap bounds the paths alive per SM, which keeps a
// ring slot from being rewritten before it was read.
#ifndef ARV2_WAVE_THREADS
#define ARV2_WAVE_THREADS 1024       // occupancy probe: 512 / 768 / 896 (r07 section 14)
#endif
constexpr int kWaveThreads = ARV2_WAVE_THREADS;
#ifndef ARV2_DRAIN_SPREAD
#define ARV2_DRAIN_SPREAD 1
#endif
#ifndef ARV2_DRAIN_BELOW
#define ARV2_DRAIN_BELOW 2048        // paths alive per SM below which batches shrink (batch = 32 * alive / this)
#endif
#ifndef ARV2_DRAIN_MIN
#define ARV2_DRAIN_MIN 1
#endif

template <int NB, int MODE>
__global__ void __launch_bounds__(kWaveThreads, 1) wave_kernel(const TraceParams p)
{
    __shared__ unsigned q_tail[kWaveQueues], q_pub[kWaveQueues], q_head[kWaveQueues];
    __shared__ int sh_alive, sh_pool_done;
    if (threadIdx.x < kWaveQueues) { q_tail[threadIdx.x] = 0; q_pub[threadIdx.x] = 0; q_head[threadIdx.x] = 0; }
    if (threadIdx.x == 0) { sh_alive = 0; sh_pool_done = 0; }
    __syncthreads();
    volatile unsigned* const v_pub = q_pub;
    volatile unsigned* const v_head = q_head;
    volatile int* const v_alive = &sh_alive;
    volatile int* const v_done = &sh_pool_done;

    const int lane = threadIdx.x & 31;
    const int cap = (int)p.wave_cap;
    const int nq = p.wave_queues;
    float4* const paths = p.wave_paths + (size_t)blockIdx.x * (size_t)nq * (size_t)cap * cont_f4(NB);
    unsigned long long segs = 0;
    Path<NB> s;
    s.ray = 0; s.org = f3(0, 0, 0); s.dir = f3(0, 0, 0); s.dist = 0.f; s.depth = 0; s.nseg = 0;
    Traversal tr;
    int stack[kStack];
    unsigned idle = 0;
#ifdef ARV2_TAILSTAT
    unsigned long long t_begin; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
#endif
#ifdef ARV2_WAVESTAT
    unsigned long long st_tasks = 0, st_rays = 0, st_fail = 0, st_idle = 0;
#define WST(x) x
#else
#define WST(x)
#endif

    for (;;) {
        // ---------------- find work (warp-uniform: src, n, pos)
        int src = -2, n = 0;                     // -1: fresh rays, >= 0: queue index
        unsigned long long pos = 0;
        if (lane == 0 && !*v_done && *v_alive + 32 <= cap) {
            if (atomicAdd(&sh_alive, 32) + 32 <= cap) {
                // (taking the unstarted rays per SM in chunks of consecutive rays of the direction order, so that an SM's warps
                // walk neighbouring bundles, was measured and lost: path lengths depend on the direction and the SMs that drew
                // long-lived chunks finish late -- profiles/r09_trace_sweeps.md)
                const long long b = (long long)atomicAdd(p.counters, 32ull);
                long long take = p.n_rays - b;
                take = take < 0 ? 0 : (take > 32 ? 32 : take);
                if (take < 32) atomicSub(&sh_alive, 32 - (int)take);
                if (take > 0) { src = -1; n = (int)take; pos = (unsigned long long)b; }
                else {
                    *v_done = 1;
#ifdef ARV2_TAILSTAT
                    { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMin(p.counters + 4, t); }
#endif
                }
            } else {
                atomicSub(&sh_alive, 32);
            }
        }
        src = __shfl_sync(FULL, src, 0);
        if (src == -2) {
            // shallowest queue with a full batch; a warp that found nothing last time also takes a partial one.
            // Once no ray is left to start and fewer paths are alive than the SM has lanes, the paths are spread over
            // all warps (a segment takes a warp as long as its slowest lane: 8 paths per warp finish their segments
            // sooner than 32, and the end of a launch is a chain of dependent segments)
            int batch = 32;
#if ARV2_DRAIN_SPREAD
            if (lane == 0 && *v_done) { const int al = *v_alive; if (al < ARV2_DRAIN_BELOW) batch = max(ARV2_DRAIN_MIN, (al * 32 + ARV2_DRAIN_BELOW - 1) / ARV2_DRAIN_BELOW); }
            batch = __shfl_sync(FULL, batch, 0);
#endif
            const int need = (idle == 0 && batch == 32) ? 32 : 1;
            int a0 = 0, a1 = 0;
            if (lane < nq) { const unsigned h = v_head[lane]; a0 = (int)(v_pub[lane] - h); }
            if (lane + 32 < nq) { const unsigned h = v_head[lane + 32]; a1 = (int)(v_pub[lane + 32] - h); }
            unsigned long long cand = (unsigned long long)__ballot_sync(FULL, a0 >= need) | ((unsigned long long)__ballot_sync(FULL, a1 >= need) << 32);
            if (lane == 0) {
                while (cand && src == -2) {
                    const int q = __ffsll((long long)cand) - 1;
                    cand &= cand - 1;
                    for (;;) {
                        const unsigned h = v_head[q];
                        const int av = (int)(v_pub[q] - h);
                        if (av < need) { WST(++st_fail;) break; }
                        const int take = av < batch ? av : batch;
                        if (atomicCAS(&q_head[q], h, h + (unsigned)take) == h) { n = take; pos = h; src = q; break; }
                    }
                }
            }
            src = __shfl_sync(FULL, src, 0);
        }
        n = __shfl_sync(FULL, n, 0);
        pos = __shfl_sync(FULL, pos, 0);

        if (src == -2) {
            // nothing to take: done when no ray is left to start and no path of this SM is alive
            int done = 0;
            if (lane == 0) done = *v_done && *v_alive == 0;
            if (__shfl_sync(FULL, done, 0)) break;
            WST(if (lane == 0) ++st_idle;)
            if (++idle > (1u << 24)) { if (lane == 0) atomicAdd(p.counters + 7, 1ull); break; }      // watchdog
            __nanosleep(100);
            continue;
        }
        idle = 0;
        WST(if (lane == 0) { ++st_tasks; st_rays += n; })

        // ---------------- take the batch
        bool have = lane < n;
        if (src == -1) {
            if (have) new_path<NB>(p, s, (long long)pos + lane);
        } else {
            __threadfence_block();
            if (have) load_path<NB>(paths + ((size_t)src * cap + (unsigned)(pos + lane) % (unsigned)cap) * cont_f4(NB), s);
        }

        // ---------------- advance
        int ended_n = 0;
        for (int k = 0; k < p.wave_segments; ++k) {
            if (!__any_sync(FULL, have)) break;
            bool ended = false;
            Deposit d; d.dep = false; d.bin = -1; d.ear = 0; d.primary = 0;
            if (have) ended = advance_segment<NB, MODE>(p, s, stack, tr, d);
            if (MODE == 0) deposit_warp<NB>(p, d.dep, d.bin, d.primary, s.energy);
            if (ended) { end_path<NB, MODE>(p, s, d, segs); have = false; ++ended_n; }
        }
        if (have && !path_goes_on<NB>(p, s)) {                   // would end at its next loop guard
            Deposit none; none.dep = false; none.bin = -1; none.ear = 0; none.primary = 0;
            end_path<NB, MODE>(p, s, none, segs); have = false; ++ended_n;
        }

        // ---------------- append the paths that go on to the queue of their depth
        const int qn = src + 1;
        const unsigned m = __ballot_sync(FULL, have);
        if (m) {
            if (qn >= nq) { if (lane == 0) atomicAdd(p.counters + 7, 1ull); break; }      // cannot happen: depth < max_bounces
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&q_tail[qn], (unsigned)__popc(m));
            base = __shfl_sync(FULL, base, 0);
            if (have) {
                store_path<NB>(paths + ((size_t)qn * cap + (base + __popc(m & ((1u << lane) - 1u))) % (unsigned)cap) * cont_f4(NB), s);
                __threadfence();
            }
            __syncwarp();
            if (lane == 0) {
                while (v_pub[qn] != base) __nanosleep(20);       // publish in reservation order
                v_pub[qn] = base + (unsigned)__popc(m);
                __threadfence_block();
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ended_n += __shfl_xor_sync(FULL, ended_n, o);
        if (lane == 0 && ended_n) atomicSub(&sh_alive, ended_n);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(FULL, segs, o);
    if (lane == 0 && segs) atomicAdd(p.counters + 1, segs);
    tr.flush_stats(p.counters);
#ifdef ARV2_TAILSTAT
    if (lane == 0) {      // per warp: exit time; per CTA: how long this SM was busy
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(p.counters + 5, t); atomicMin(p.counters + 6, t);
        if (threadIdx.x == 0) { atomicAdd(p.counters + 12, t - t_begin); atomicAdd(p.counters + 13, 1ull); atomicMin(p.counters + 14, t - t_begin); }
    }
#endif
#ifdef ARV2_WAVESTAT
    if (lane == 0) {
        atomicAdd(p.counters + 8, st_tasks); atomicAdd(p.counters + 9, st_rays); atomicAdd(p.counters + 10, st_fail);
        atomicAdd(p.counters + 11, st_idle);
    }
#endif
}

// ---------------------------------------------------------------------------------------
// sweep_kernel: bounce-synchronous tracer for launches of many millions of rays.  Every sweep advances ALL paths alive
// by sweep.segments segments, one lane per path, and hands the survivors over through global memory; between two sweeps
// the survivors are re-binned by (cell of their new origin, octant-map cell of their new direction): a counting sort whose
// histogram and per-path rank are taken by the sweep itself (one atomic per survivor), followed by a scan of the bins and
// a scatter of the read order.  The 32 lanes of a warp then start in the same part of the scene and head the same way at
// EVERY depth, not only over the first bounces of the direction-sorted start order (wave_kernel keeps a warp together,
// but a bundle has fanned out after ~5 bounces).  Results are per ray and the histogram is fp64: nothing depends on the order.
#ifndef ARV2_SWEEP_THREADS
#define ARV2_SWEEP_THREADS 128
#endif
constexpr int kSweepThreads = ARV2_SWEEP_THREADS;
__device__ __forceinline__ unsigned spread16(unsigned v);

__device__ __forceinline__ unsigned spread3_10(unsigned v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ unsigned sweep_bin(const SweepParams& w, F3 org, F3 d)
{
    const float cmax = (float)((1 << w.cell_bits) - 1);
    const unsigned cx = (unsigned)fminf(fmaxf((org.x - w.lo[0]) * w.scale[0], 0.f), cmax);
    const unsigned cy = (unsigned)fminf(fmaxf((org.y - w.lo[1]) * w.scale[1], 0.f), cmax);
    const unsigned cz = (unsigned)fminf(fmaxf((org.z - w.lo[2]) * w.scale[2], 0.f), cmax);
    const unsigned mc = spread3_10(cx) | (spread3_10(cy) << 1) | (spread3_10(cz) << 2);
    const float inv = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + 1e-30f);
    float u = d.x * inv, v = d.y * inv;
    if (d.z < 0.f) { const float uu = (1.f - fabsf(v)) * copysignf(1.f, u), vv = (1.f - fabsf(u)) * copysignf(1.f, v); u = uu; v = vv; }
    const float dmax = (float)((1 << w.dir_bits) - 1), ds = (float)(1 << w.dir_bits);
    const unsigned iu = (unsigned)fminf(fmaxf((u * 0.5f + 0.5f) * ds, 0.f), dmax);
    const unsigned iv = (unsigned)fminf(fmaxf((v * 0.5f + 0.5f) * ds, 0.f), dmax);
    const unsigned md = spread16(iu) | (spread16(iv) << 1);
    return w.dir_major ? ((md << (3 * w.cell_bits)) | mc) : ((mc << (2 * w.dir_bits)) | md);
}

template <int NB, int MODE>
__global__ void __launch_bounds__(kSweepThreads, NB == 1 ? ARV2_MINB : ARV2_MINB8) sweep_kernel(const TraceParams p, const SweepParams w)
{
    __shared__ unsigned sh_cnt[kSweepThreads / 32];
    __shared__ unsigned long long sh_base;
    const long long n_in = w.in ? (long long)*w.n_in : p.n_rays;
    const long long i = (long long)blockIdx.x * kSweepThreads + threadIdx.x;
    if (i == 0) atomicAdd(p.counters + 2, 1ull);             // sweeps of this launch (arv2_last_counters)
    if (i - threadIdx.x >= n_in) return;                     // CTA-uniform
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long segs = 0;
    Path<NB> s;
    s.ray = 0; s.org = f3(0, 0, 0); s.dir = f3(0, 0, 0); s.dist = 0.f; s.depth = 0; s.nseg = 0;
    Traversal tr;
    int stack[kStack];
    bool have = i < n_in;
    if (have) {
        if (!w.in) new_path<NB>(p, s, i);
        else load_path<NB>(w.in + (size_t)(w.perm ? __ldg(w.perm + i) : (int)i) * cont_f4(NB), s);
    }
    for (int k = 0; k < w.segments; ++k) {
        if (!__any_sync(FULL, have)) break;
        bool ended = false;
        Deposit d; d.dep = false; d.bin = -1; d.ear = 0; d.primary = 0;
        if (have) ended = advance_segment<NB, MODE>(p, s, stack, tr, d);
        if (MODE == 0) deposit_warp<NB>(p, d.dep, d.bin, d.primary, s.energy);
        if (ended) { end_path<NB, MODE>(p, s, d, segs); have = false; }
    }
    if (have && !path_goes_on<NB>(p, s)) {                   // would end at its next loop guard
        Deposit none; none.dep = false; none.bin = -1; none.ear = 0; none.primary = 0;
        end_path<NB, MODE>(p, s, none, segs); have = false;
    }
    // ---- survivors: compacted per CTA (one global atomic), binned for the next sweep
    const unsigned m = __ballot_sync(FULL, have);
    if (lane == 0) sh_cnt[warp] = (unsigned)__popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned total = 0;
#pragma unroll
        for (int q = 0; q < kSweepThreads / 32; ++q) total += sh_cnt[q];
        sh_base = total ? atomicAdd(w.n_out, (unsigned long long)total) : 0ull;
    }
    __syncthreads();
    if (have) {
        unsigned long long pos = sh_base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
        for (int q = 0; q < warp; ++q) pos += sh_cnt[q];
        store_path<NB>(w.out + (size_t)pos * cont_f4(NB), s);
        const unsigned bin = sweep_bin(w, s.org, s.dir);
        w.key[pos] = bin;
        w.rank[pos] = atomicAdd(w.bins + bin, 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(FULL, segs, o);
    if (lane == 0 && segs) atomicAdd(p.counters + 1, segs);
    tr.flush_stats(p.counters);
}

// exclusive scan of the bin counts in place, tiles of 4096 bins (n_bins a multiple of 4096): sweep_tile_sums_kernel
// writes one total per tile, sweep_scan_kernel re-reads its tile, adds up the totals of the tiles before it (at most
// 1024 of them: one pass of the CTA) and writes the exclusive prefix
__device__ __forceinline__ unsigned block_exclusive_1024(unsigned v, unsigned* sh, unsigned* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned t = sh[lane];
        unsigned ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(FULL, ti, o); if (lane >= o) ti += u; }
        sh[lane] = ti - t;
        if (lane == 31) sh[32] = ti;
    }
    __syncthreads();
    if (total) *total = sh[32];
    return sh[warp] + inc - v;
}

__global__ void __launch_bounds__(1024) sweep_tile_sums_kernel(const unsigned* __restrict__ bins, unsigned* __restrict__ tile_sums)
{
    __shared__ unsigned sh[33];
    const uint4 v = reinterpret_cast<const uint4*>(bins)[(size_t)blockIdx.x * 1024 + threadIdx.x];
    unsigned total;
    block_exclusive_1024(v.x + v.y + v.z + v.w, sh, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) sweep_scan_kernel(unsigned* __restrict__ bins, const unsigned* __restrict__ tile_sums)
{
    __shared__ unsigned sh[33];
    __shared__ unsigned sh_before;
    unsigned before = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += 1024) before += tile_sums[t];
    unsigned tot;
    block_exclusive_1024(before, sh, &tot);
    if (threadIdx.x == 0) sh_before = tot;
    __syncthreads();
    uint4* const b = reinterpret_cast<uint4*>(bins) + (size_t)blockIdx.x * 1024 + threadIdx.x;
    const uint4 v = *b;
    unsigned run = sh_before + block_exclusive_1024(v.x + v.y + v.z + v.w, sh, nullptr);
    uint4 o;
    o.x = run; run += v.x; o.y = run; run += v.y; o.z = run; run += v.z; o.w = run;
    *b = o;
}

__global__ void __launch_bounds__(256) sweep_scatter_kernel(const unsigned* __restrict__ key, const unsigned* __restrict__ rank,
                                                            const unsigned* __restrict__ bins, const unsigned long long* __restrict__ n, int* __restrict__ perm)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)*n) return;
    perm[__ldg(bins + key[i]) + rank[i]] = (int)i;
}

// ---------------------------------------------------------------------------------------
// trace2_kernel: the same path tracer with the lanes of a warp decoupled.  Per-segment
// node-visit counts vary so much inside a warp (sum / (32 x max) = 49 % on the conference
// scene) that lanes of trace_kernel idle most of the time (9.5 of 32 lanes per instruction,
// r03b profile) and the kernel ends up latency-bound with too few loads in flight.  Here a
// lane that has finished its segment parks while the others keep traversing; three
// warp-uniform phases are scheduled by ballots:
//   A  shade + deposit + refill + start the next segment -- when >= ARV2_TA lanes are parked
//   L  one leaf (<= 4 triangle tests) per lane           -- when >= ARV2_TL lanes stand at a leaf
//   I  up to ARV2_BURST inner-node steps                  -- otherwise
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
__global__ void __launch_bounds__(kThreads, NB == 1 ? ARV2_MINB : ARV2_MINB8) trace2_kernel(const TraceParams p)
{
    const int lane = threadIdx.x & 31;
    long long chunk_next = 0, chunk_end = 0;      // warp-uniform
    bool have = false, exhausted = false, pending = false;
    Path<NB> s;
    s.ray = 0; s.org = f3(0, 0, 0); s.dir = f3(0, 0, 0); s.dist = 0.f; s.depth = 0; s.nseg = 0;
    unsigned long long segs = 0;
    Traversal tr;
    tr.reset(1e20f);
    RayGrid g;
    g.setup(p, f3(0.f, 0.f, 0.f), f3(1.f, 1.f, 1.f));
    int stack[kStack];

    for (;;) {
        const bool parked = tr.finished() && !exhausted;
        const unsigned inner_m = __ballot_sync(FULL, tr.at_inner());
        const unsigned leaf_m = __ballot_sync(FULL, tr.at_leaf());
        const unsigned park_m = __ballot_sync(FULL, parked);
        if ((inner_m | leaf_m | park_m) == 0) break;

        if (park_m != 0 && (__popc(park_m) >= ARV2_TA || (inner_m | leaf_m) == 0)) {
            // ---- phase A
            bool ended = false;
            Deposit d; d.dep = false; d.bin = -1; d.ear = 0; d.primary = 0;
            if (parked && have && pending) {
                pending = false;
                ended = shade_segment<NB, MODE>(p, s, tr.h, d);
            }
            if (MODE == 0) deposit_warp<NB>(p, d.dep, d.bin, d.primary, s.energy);
            if (parked && have && !ended && !path_goes_on<NB>(p, s)) ended = true;
            if (parked && have && ended) { end_path<NB, MODE>(p, s, d, segs); have = false; }
            if (refill<NB>(p, parked && !have, chunk_next, chunk_end, exhausted, s)) {
                have = true;
                if (!path_goes_on<NB>(p, s)) {      // only through its parameters (or a zero direction)
                    Deposit none; none.dep = false; none.bin = -1; none.ear = 0; none.primary = 0;
                    end_path<NB, MODE>(p, s, none, segs);
                    have = false;
                }
            }
            if (parked && have) {
                begin_segment<NB, MODE>(p, s);
                tr.reset(1e20f);
                pending = true;
                const bool to_recv = MODE == 0 && p.recv_root >= 0 && enters_receiver_ball(p, s.org, s.dir, 1e20f);
                const int root = to_recv ? p.root : p.scene_root;
                if (root >= 0) {
                    g.setup(p, s.org, s.dir);
                    tr.enter(stack, root);
                }
            }
        } else if (leaf_m != 0 && (__popc(leaf_m) >= ARV2_TL || inner_m == 0)) {
            // ---- phase L
            if (tr.at_leaf()) tr.step_leaf(stack, p.tris, s.org, s.dir);
        } else {
            // ---- phase I
#pragma unroll 1
            for (int k = 0; k < ARV2_BURST; ++k)
                if (tr.at_inner()) tr.step_inner(stack, p, g, s.org);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(FULL, segs, o);
    if (lane == 0 && segs) atomicAdd(p.counters + 1, segs);
}

// ---------------------------------------------------------------------------------------
// Receiver move.  The cached paths do not depend on the receiver; it only ends them (OR/devicePrograms.cu:147,169), so a
// move re-deposits: per ray, the first cached segment the receiver intercepts before the wall (t_recv < t_wall; ties go
// to the scene because scene triangle ids are lower) deposits, exactly as a fresh trace would.
//
// Two data-parallel passes over the packed (CSR) cache (profiles/r08_rerender.md):
//   rr_mask_kernel  streams the path vertices: the segment between two consecutive 8 B vertices (16-bit grid) against
//                   the receiver's bounding ball, padded by what a true segment may stray from the quantised one --
//                   8 B per segment instead of the 32 B record, coalesced, no dependence between rays (the r06 kernel
//                   scanned each ray serially: a chain of dependent DRAM round trips, 10 of 32 lanes).  The flagged few
//                   per cent are parked per warp and refined 32 at a time with their exact record (first step of the
//                   walk: the root boxes of the receiver tree); what passes sets its bit in the bit array.
//   rr_walk_kernel  one lane per ray: its flagged segments, in order, are walked through the receiver tree with the exact
//                   32 B record until the first hit, which ends the ray (warp-aggregated fp64 deposit); the lane then
//                   takes the next ray.  Walks behind a ray's first hit never happen (the r07 data-parallel experiment
//                   walked every flagged segment).  A first version that packed the offers of 256 rays into full warps
//                   through a shared-memory queue spent its time in the CTA barriers between rounds (r08).
// Conservative flags + exact walks: per-ray results are bit-identical to the serial kernel and to a fresh trace.
constexpr int kRrThreads = 256;
#ifndef ARV2_RRW_MINB
#define ARV2_RRW_MINB 3               // CTAs of rr_walk_kernel per SM
#endif

__device__ __forceinline__ F3 vert_pos(const TraceParams& p, uint2 v)
{
    return f3(fmaf((float)(v.x & 0xffffu), p.pc_qs[0], p.pc_q0[0]), fmaf((float)(v.x >> 16), p.pc_qs[1], p.pc_q0[1]),
              fmaf((float)(v.y & 0xffffu), p.pc_qs[2], p.pc_q0[2]));
}
constexpr unsigned kVertSegment = 1u << 16, kVertEscapes = 2u << 16;      // flags in the top half of uint2::y

// A vertex's coordinate without an integer-to-float conversion (I2F issues at a fraction of the FMA rate): the 16-bit
// grid index is dropped into the mantissa of 2^23, so that fmaf(as_float(0x4B000000 | q), qs, q0 - c - 2^23 qs) is the
// coordinate relative to the ball's centre.
__device__ __forceinline__ float grid_coord(unsigned q16, float qs, float bias) { return fmaf(__uint_as_float(0x4B000000u | q16), qs, bias); }

// The first step of a walk of the receiver tree, on its own: does the segment (org + t dir, 0 <= t <= tmax) enter either
// child box of the tree's root?  Same arithmetic as Traversal::step_inner / descend, so skipping a segment that fails
// it is exactly what the walk would have concluded.
__device__ __forceinline__ bool enters_receiver_root(const TraceParams& p, F3 org, F3 dir, float tmax)
{
    const float ix = safe_rcp(dir.x), iy = safe_rcp(dir.y), iz = safe_rcp(dir.z);
    const float ox = org.x * ix, oy = org.y * iy, oz = org.z * iz;
    const F8 na = ldg256(p.nodes + p.recv_root * 4), nb = ldg256(p.nodes + p.recv_root * 4 + 2);
    const float4 n0 = na.lo, n1 = na.hi, n2 = nb.lo;
    const float c0lox = fmaf(n0.x, ix, -ox), c0hix = fmaf(n0.y, ix, -ox), c0loy = fmaf(n0.z, iy, -oy), c0hiy = fmaf(n0.w, iy, -oy);
    const float c0loz = fmaf(n2.x, iz, -oz), c0hiz = fmaf(n2.y, iz, -oz);
    const float c1lox = fmaf(n1.x, ix, -ox), c1hix = fmaf(n1.y, ix, -ox), c1loy = fmaf(n1.z, iy, -oy), c1hiy = fmaf(n1.w, iy, -oy);
    const float c1loz = fmaf(n2.z, iz, -oz), c1hiz = fmaf(n2.w, iz, -oz);
    const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.f));
    const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), tmax));
    const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.f));
    const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), tmax));
    return c0min <= c0max || c1min <= c1max;
}

// A warp takes 128 consecutive vertices per trip, lane l the four vertices 4l .. 4l+3 (one 256-bit load; the array is
// padded to whole tiles + 1): three of a lane's four segments end on its own vertices, the fourth on the next lane's first.
constexpr int kMaskTile = 128;

__global__ void __launch_bounds__(kRrThreads) rr_mask_kernel(const TraceParams p)
{
    const long long n = p.pc_nvert;
    const float r = p.recv_radius + p.pc_eps, r2 = r * r * 1.0001f;
    const float qx = p.pc_qs[0], qy = p.pc_qs[1], qz = p.pc_qs[2];
    const float bx = p.pc_q0[0] - p.center[0] - 8388608.f * qx, by = p.pc_q0[1] - p.center[1] - 8388608.f * qy, bz = p.pc_q0[2] - p.center[2] - 8388608.f * qz;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    // Second stage, on the flagged segments only (a few per cent): the first step of the walk of the receiver tree --
    // does the exact segment enter a root box before it ends on its wall?  Most flagged segments end on geometry next
    // to the receiver or only cross the rim of the ball; what passes sets its bit in the (zeroed) bit array.  The
    // record of segment i is stored at index i, like its vertex.
    __shared__ unsigned sh_buf[kRrThreads / 32][32 + kMaskTile];
    unsigned* const buf = sh_buf[threadIdx.x >> 5];
    int n_buf = 0;                                 // warp-uniform
    auto refine = [&](unsigned idx) {
        const F8 rec = ldg256_tri(p.pc_seg + 2 * (size_t)idx);
        if (enters_receiver_root(p, f3(rec.lo.x, rec.lo.y, rec.lo.z), f3(rec.hi.x, rec.hi.y, rec.hi.z), rec.lo.w))
            atomicOr(p.pc_bits + (idx >> 5), 1u << (idx & 31));
    };
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp0 * kMaskTile; base < n; base += n_warps * kMaskTile) {
        const uint2* src = p.pc_vert + base + lane * 4;
        uint4 lo, hi;
        asm("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(src));
        // the vertex after this lane's four: the next lane's first, or (lane 31) the first of the next tile
        unsigned nx = __shfl_down_sync(FULL, lo.x, 1), ny = __shfl_down_sync(FULL, lo.y, 1);
        if (lane == 31) { const uint2 w = __ldg(p.pc_vert + base + kMaskTile); nx = w.x; ny = w.y; }
        const unsigned wx[5] = {lo.x, lo.z, hi.x, hi.z, nx}, wy[5] = {lo.y, lo.w, hi.y, hi.w, ny};
        float x[5], y[5], z[5], qq[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            x[j] = grid_coord(wx[j] & 0xffffu, qx, bx); y[j] = grid_coord(wx[j] >> 16, qy, by); z[j] = grid_coord(wy[j] & 0xffffu, qz, bz);
            qq[j] = x[j] * x[j] + y[j] * y[j] + z[j] * z[j];                  // squared distance from the centre
        }
        unsigned nib = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // closest point of the segment (a = vertex j, e = vertex j+1, relative to the centre) to the origin, division-free
            const float dx = x[j + 1] - x[j], dy = y[j + 1] - y[j], dz = z[j + 1] - z[j];
            const float dd = dx * dx + dy * dy + dz * dz;
            const float ad = -(x[j] * dx + y[j] * dy + z[j] * dz);            // (centre - a) . d
            const bool inside = ad <= 0.f ? qq[j] <= r2 : (ad >= dd ? qq[j + 1] <= r2 : (qq[j] - r2) * dd <= ad * ad * 1.0001f);
            const bool cand = (wy[j] & kVertSegment) && (inside || (wy[j] & kVertEscapes));       // no end point: always walked
            nib |= cand ? (1u << j) : 0u;
        }
        // ---- park the flagged segments of this trip in the warp's buffer ...
        if (__any_sync(FULL, nib != 0u)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned c = __ballot_sync(FULL, (nib >> j) & 1u);
                if ((nib >> j) & 1u) buf[n_buf + __popc(c & lt)] = (unsigned)(base + lane * 4 + j);
                n_buf += __popc(c);
            }
            __syncwarp();
        }
        // ... and refine them 32 at a time
        while (n_buf >= 32) { n_buf -= 32; refine(buf[n_buf + lane]); __syncwarp(); }
    }
    if (lane < n_buf) refine(buf[lane]);
}

// first set bit of the bit array in [a, b), or -1
__device__ __forceinline__ long long next_flag(const unsigned* __restrict__ bits, long long a, long long b)
{
    while (a < b) {
        unsigned w = __ldg(bits + (a >> 5)) & (0xffffffffu << (a & 31));
        const long long word_end = (a | 31) + 1;
        if (b < word_end) w &= 0xffffffffu >> (word_end - b);
        if (w) return (a & ~31LL) + (__ffs(w) - 1);
        a = word_end;
    }
    return -1;
}

// One lane = one ray at a time, and the lanes of a warp are decoupled: a walk of the receiver tree takes 5 to 80 node
// visits, so a warp that starts 32 walks together and waits for the longest runs at 6.5 of 32 lanes (r08 capture of that
// version).  Here a lane whose walk has finished parks; once kRrAcquire lanes are parked (or nothing else is left to do)
// the warp resolves them -- a hit deposits and ends the ray, a miss moves on -- and hands each its next offer: the ray's
// next flagged segment that also enters the root boxes of the receiver tree, from the next ray of the warp's chunk
// when the ray is used up (ballot / popc refill).  In between, the warp takes single steps: inner nodes for the lanes
// that stand at one, a leaf for the lanes that stand at one once kRrLeaf of them do.
#ifndef ARV2_RR_ACQUIRE
#define ARV2_RR_ACQUIRE 12
#endif
#ifndef ARV2_RR_LEAF
#define ARV2_RR_LEAF 6
#endif
#ifndef ARV2_RR_BURST
#define ARV2_RR_BURST 4
#endif

template <int NB>
__global__ void __launch_bounds__(kRrThreads, ARV2_RRW_MINB) rr_walk_kernel(const TraceParams p)
{
    // the receiver tree's nodes (a few hundred, 64 B each) in shared memory: a node visit is a dependent load, and most
    // of this kernel's time was spent waiting for those from L1 / L2 (r08 capture: 6.9 warps per issue on long_scoreboard)
    extern __shared__ float4 sh_nodes[];
    const bool in_shared = p.recv_nodes_shared > 0;
    for (int i = threadIdx.x; i < p.recv_nodes_shared * 4; i += blockDim.x) sh_nodes[i] = __ldg(p.nodes + (size_t)p.recv_root * 4 + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    long long chunk_base = 0;                     // warp-uniform: first ray of the warp's chunk of kChunk rays
    unsigned pend[4] = {0u, 0u, 0u, 0u};          // warp-uniform: rays of the chunk that have a flagged segment and no lane yet
    bool pool_empty = false;                      // warp-uniform
    long long ray = -1, vbase = 0;
    unsigned long long off0 = 0;
    int n = 0, cursor = 0, k = -1;
    bool walking = false;                         // a walk of segment k is under way (or has just finished)
    F3 org = f3(0.f, 0.f, 0.f), dir = f3(1.f, 1.f, 1.f);
    float t_wall = 0.f, dist0 = 0.f;
    unsigned long long segs = 0;
    Traversal tr;
    tr.reset(1e20f);
    RayGrid g;
    g.setup(p, org, dir);
    int stack[kStack];
    float energy[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) energy[b] = 0.f;

    for (;;) {
        const unsigned inner_m = __ballot_sync(FULL, walking && tr.at_inner());
        const unsigned leaf_m = __ballot_sync(FULL, walking && tr.at_leaf());
        const unsigned busy = inner_m | leaf_m;
        const int parked = __popc(__ballot_sync(FULL, walking && tr.finished()));      // walks waiting to be resolved
        // (after a service phase every lane walks or, once the pool is empty, is out of rays for good)
        if (busy == 0 || parked >= ARV2_RR_ACQUIRE || (pool_empty && 2 * parked >= __popc(busy) + parked)) {
            // ---- resolve the finished walks
            bool dep = false;
            int bin = -1, primary = 0;
            if (walking && tr.finished()) {
                walking = false;
                const Hit& h = tr.h;
                if (h.slot >= 0 && h.t < t_wall) {
                    const size_t ci = (size_t)(vbase + k);
                    const F8 A = ldg256(p.tris + h.slot * 4), B = ldg256(p.tris + h.slot * 4 + 2);
                    const int mat = __float_as_int(A.hi.w);
                    const F3 pt = hit_point(f3(A.lo.x, A.lo.y, A.lo.z), f3(A.hi.x, A.hi.y, A.hi.z), f3(B.lo.x, B.lo.y, B.lo.z), h.u, h.v);
                    const F3 dp = sub3(pt, org);
                    const float dist = __fadd_rn(dist0, __fsqrt_rn(dot3(dp, dp)));
#pragma unroll
                    for (int b = 0; b < NB; ++b) energy[b] = __ldg(p.pc_energy + ci * NB + b);
                    bin = receiver_hit<NB>(p, pt, dir, dist, energy);
                    primary = (mat == -1) ? 0 : 1;
                    dep = bin >= 0 && bin < p.ir_len;
                    if (p.rec_bin) p.rec_bin[ray] = bin;
                    if (p.rec_ear) p.rec_ear[ray] = (mat == -1) ? 1 : 2;
                    if (p.rec_nseg) p.rec_nseg[ray] = k + 1;
                    if (p.rec_energy) {
#pragma unroll
                        for (int b = 0; b < NB; ++b) p.rec_energy[(size_t)ray * NB + b] = energy[b];
                    }
                    segs += (unsigned long long)(k + 1);
                    ray = -1;                                // the path ends on the receiver (OR/devicePrograms.cu:147,169)
                } else {
                    cursor = k + 1;
                }
            }
            deposit_warp<NB>(p, dep, bin, primary, energy);

            // ---- every parked lane gets its next offer, or the pool runs dry
            bool offer = false;
            for (;;) {
                if (!walking && !offer && ray >= 0) {
                    // the ray's next flagged segment (one that only grazes the bounding ball ends its walk at the root
                    // of the receiver tree: one step of one lane, now that the lanes are decoupled)
                    const long long f = next_flag(p.pc_bits, vbase + cursor, vbase + n);
                    if (f >= 0) {
                        k = (int)(f - vbase); offer = true;
                        const F8 rec = ldg256(p.pc_seg + 2 * (size_t)(vbase + k));
                        org = f3(rec.lo.x, rec.lo.y, rec.lo.z); dir = f3(rec.hi.x, rec.hi.y, rec.hi.z); t_wall = rec.lo.w; dist0 = rec.hi.w;
                    }
                    if (!offer) {                            // every cached segment passed: a miss
                        if (p.rec_bin) p.rec_bin[ray] = -1;
                        if (p.rec_ear) p.rec_ear[ray] = 0;
                        if (p.rec_nseg) p.rec_nseg[ray] = n;
                        if (p.rec_energy) {
#pragma unroll
                            for (int b = 0; b < NB; ++b) p.rec_energy[(size_t)ray * NB + b] = 0.f;
                        }
                        segs += (unsigned long long)n;
                        ray = -1;
                    }
                }
                const unsigned need = __ballot_sync(FULL, ray < 0);
                if (need == 0) break;
                int pending = __popc(pend[0]) + __popc(pend[1]) + __popc(pend[2]) + __popc(pend[3]);
                if (pending == 0) {
                    if (pool_empty) break;
                    // the warp's next 128 rays, looked at by all lanes first: a ray without a flagged segment (in a large
                    // hall: nearly all of them) is closed here and now, the others wait in `pend` for a free lane
                    unsigned long long b = 0;
                    if (lane == 0) b = atomicAdd(p.counters, (unsigned long long)kChunk);
                    chunk_base = (long long)__shfl_sync(FULL, b, 0);
                    if (chunk_base >= p.n_rays) { pool_empty = true; break; }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const long long r = chunk_base + q * 32 + lane;
                        bool open = false;
                        if (r < p.n_rays) {
                            const unsigned long long o0 = __ldg(p.pc_off + r);
                            const int rn = (int)(__ldg(p.pc_off + r + 1) - o0);
                            const long long vb = (long long)o0 + r;
                            open = next_flag(p.pc_bits, vb, vb + rn) >= 0;
                            if (!open) {                     // every cached segment passes: a miss
                                if (p.rec_bin) p.rec_bin[r] = -1;
                                if (p.rec_ear) p.rec_ear[r] = 0;
                                if (p.rec_nseg) p.rec_nseg[r] = rn;
                                if (p.rec_energy) {
#pragma unroll
                                    for (int bb = 0; bb < NB; ++bb) p.rec_energy[(size_t)r * NB + bb] = 0.f;
                                }
                                segs += (unsigned long long)rn;
                            }
                        }
                        pend[q] = __ballot_sync(FULL, open);
                    }
                    continue;
                }
                // hand the first popc(need) pending rays to the lanes in need, in order
                int rank = __popc(need & lt);
                if (ray < 0 && rank < pending) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cq = __popc(pend[q]);
                        if (rank >= 0 && rank < cq) { ray = chunk_base + q * 32 + (int)__fns(pend[q], 0, rank + 1); rank = -1; }
                        else if (rank >= 0) rank -= cq;
                    }
                    off0 = __ldg(p.pc_off + ray);
                    n = (int)(__ldg(p.pc_off + ray + 1) - off0);
                    vbase = (long long)off0 + ray;
                    cursor = 0;
                }
                int take = min(__popc(need), pending);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cq = __popc(pend[q]);
                    const int t = min(take, cq);
                    if (t == cq) pend[q] = 0u;
                    else if (t > 0) pend[q] &= ~((2u << __fns(pend[q], 0, t)) - 1u);
                    take -= t;
                }
            }
            if (offer) {
                walking = true;
                tr.reset(t_wall);
                g.setup(p, org, dir);
                tr.enter(stack, p.recv_root);
            }
            if (!__any_sync(FULL, walking)) break;           // the pool is empty and no ray is open
            continue;
        }
        // ---- single steps
        if (leaf_m != 0 && (__popc(leaf_m) >= ARV2_RR_LEAF || inner_m == 0)) {
            if (walking && tr.at_leaf()) tr.step_leaf<true>(stack, p.tris, org, dir);
        } else {
#pragma unroll 1
            for (int s = 0; s < ARV2_RR_BURST; ++s)
                if (walking && tr.at_inner()) {
                    if (in_shared) tr.step_inner_shared(stack, sh_nodes, p.recv_root, g);
                    else tr.step_inner(stack, p, g, org);
                }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(FULL, segs, o);
    if (lane == 0 && segs) atomicAdd(p.counters + 1, segs);
}

// ---------------------------------------------------------------------------------------
// The r06 kernel on the packed cache (A/B: ARV2_RR_SERIAL=1): every lane scans its ray's 32 B records in order (four
// records = one 128 B line per step, the next two lines prefetched into L2), tests them against the bounding ball,
// parks candidates per warp in shared memory and walks them through the receiver tree 32 at a time; misses resume from
// a warp-local buffer.  0.62 ms on C2 against the two passes above (profiles/r08_rerender.md).
#ifndef ARV2_RR_MINB
#define ARV2_RR_MINB 4
#endif
constexpr int kParkSlots = 64;

template <int NB>
__global__ void __launch_bounds__(kRerenderThreads, ARV2_RR_MINB) rerender_kernel(const TraceParams p)
{
    __shared__ int2 sh_cand[kRerenderThreads / 32][kParkSlots];   // (ray, k | n << 16)
    __shared__ int2 sh_res[kRerenderThreads / 32][kParkSlots];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int2* const cand = sh_cand[threadIdx.x >> 5];
    int2* const res = sh_res[threadIdx.x >> 5];
    int n_cand = 0, n_res = 0;                    // warp-uniform
    long long chunk_next = 0, chunk_end = 0;      // warp-uniform
    bool pool_empty = false;                      // warp-uniform
    bool have = false;
    int ray = 0, k = 0, n = 0;
    size_t base = 0;                              // first record of `ray`
    float energy[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) energy[b] = 0.f;
    unsigned long long segs = 0;
    Traversal tr;
    int stack[kStack];

    // a ray leaves: miss (every cached segment passed) or receiver hit
    auto finish = [&](int r, int bin, int ear, int nseg) {
        if (p.rec_bin) p.rec_bin[r] = bin;
        if (p.rec_ear) p.rec_ear[r] = ear;
        if (p.rec_nseg) p.rec_nseg[r] = nseg;
        if (p.rec_energy) {
#pragma unroll
            for (int b = 0; b < NB; ++b) p.rec_energy[(size_t)r * NB + b] = ear ? energy[b] : 0.f;
        }
        segs += (unsigned long long)nseg;
    };

    for (;;) {
        // ---- refill the lanes without a ray: parked misses first, then the pool
        unsigned need = __ballot_sync(FULL, !have);
        if (need && n_res > 0) {
            const int take = min(__popc(need), n_res);
            const int rank = __popc(need & lt);
            if (!have && rank < take) {
                const int2 e = res[n_res - 1 - rank];
                ray = e.x; k = e.y & 0xffff; n = (int)((unsigned)e.y >> 16); have = true;
                base = (size_t)__ldg(p.pc_off + ray) + (size_t)ray;
            }
            n_res -= take;
            __syncwarp();
            need = __ballot_sync(FULL, !have);
        }
        while (need && !pool_empty && n_cand + n_res <= 32) {
            if (chunk_next >= chunk_end) {
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(p.counters, (unsigned long long)kChunk);
                b = __shfl_sync(FULL, b, 0);
                chunk_next = (long long)b;
                chunk_end = min((long long)b + kChunk, p.n_rays);
                if (chunk_next >= chunk_end) { pool_empty = true; break; }
            }
            const int avail = (int)min((long long)32, chunk_end - chunk_next);
            const int rank = __popc(need & lt);
            if (!have && rank < avail) {
                ray = (int)(chunk_next + rank);
                base = (size_t)__ldg(p.pc_off + ray);
                n = (int)(__ldg(p.pc_off + ray + 1) - base);
                base += (size_t)ray;
                k = 0;
                have = n > 0;
                if (n == 0) finish(ray, -1, 0, 0);
            }
            chunk_next += min(__popc(need), avail);
            need = __ballot_sync(FULL, !have);
        }
        const unsigned hv = __ballot_sync(FULL, have);
        if (hv == 0 && n_cand == 0) {
            if (n_res == 0 && pool_empty) break;
            if (n_res == 0 && !pool_empty) continue;     // (cannot happen: the pool refill ran)
        }

        if (n_cand >= 32 || (hv == 0 && n_cand > 0)) {
            // ---- receiver walk for up to 32 parked candidates
            const int m = min(32, n_cand);
            bool dep = false;
            int bin = -1, primary = 0;
            bool miss = false;
            int c_ray = 0, c_k = 0, c_n = 0;
            if (lane < m) {
                const int2 e = cand[n_cand - m + lane];
                c_ray = e.x; c_k = e.y & 0xffff; c_n = (int)((unsigned)e.y >> 16);
                const size_t ci = (size_t)__ldg(p.pc_off + c_ray) + (size_t)c_ray + (size_t)c_k;
                const F8 rec = ldg256(p.pc_seg + 2 * ci);
                const float4 ot = rec.lo, dd = rec.hi;
                const F3 org = f3(ot.x, ot.y, ot.z), dir = f3(dd.x, dd.y, dd.z);
                closest_hit(p, stack, tr, p.recv_root, org, dir, ot.w);
                const Hit& h = tr.h;
                if (h.slot >= 0 && h.t < ot.w) {
                    const F8 A = ldg256(p.tris + h.slot * 4), B = ldg256(p.tris + h.slot * 4 + 2);
                    const int mat = __float_as_int(A.hi.w);
                    const F3 pt = hit_point(f3(A.lo.x, A.lo.y, A.lo.z), f3(A.hi.x, A.hi.y, A.hi.z), f3(B.lo.x, B.lo.y, B.lo.z), h.u, h.v);
                    const F3 dp = sub3(pt, org);
                    const float dist = __fadd_rn(dd.w, __fsqrt_rn(dot3(dp, dp)));
#pragma unroll
                    for (int b = 0; b < NB; ++b) energy[b] = p.pc_energy[ci * NB + b];
                    bin = receiver_hit<NB>(p, pt, dir, dist, energy);
                    primary = (mat == -1) ? 0 : 1;
                    dep = bin >= 0 && bin < p.ir_len;
                    finish(c_ray, bin, (mat == -1) ? 1 : 2, c_k + 1);
                } else if (c_k + 1 >= c_n) {
                    finish(c_ray, -1, 0, c_n);
                } else {
                    miss = true;
                }
            }
            deposit_warp<NB>(p, dep, bin, primary, energy);
            n_cand -= m;
            const unsigned mm = __ballot_sync(FULL, miss);
            if (miss) res[n_res + __popc(mm & lt)] = make_int2(c_ray, (c_k + 1) | (c_n << 16));
            n_res += __popc(mm);
            __syncwarp();
            continue;
        }

        // ---- scan: advance to the next cached segment that enters the receiver's bounding ball
        bool found = false;
        if (have) {
#pragma unroll 1
            for (int burst = 0; burst < 2 && have && !found; ++burst) {
                const float4* rec = p.pc_seg + 2 * (base + (size_t)k);
                if (k + 4 < n) prefetch_l2(rec + 8);                 // the next lines of this ray: the scan is otherwise a
                if (k + 8 < n) prefetch_l2(rec + 16);                // chain of dependent DRAM round trips
                const F8 r0 = ldg256(rec);
                F8 r1 = r0, r2 = r0, r3 = r0;
                if (k + 1 < n) r1 = ldg256(rec + 2);
                if (k + 2 < n) r2 = ldg256(rec + 4);
                if (k + 3 < n) r3 = ldg256(rec + 6);
#define ARV2_SCAN_STEP(r)                                                                                                          \
                if (have && !found) {                                                                                              \
                    if (enters_receiver_ball(p, f3(r.lo.x, r.lo.y, r.lo.z), f3(r.hi.x, r.hi.y, r.hi.z), r.lo.w)) found = true;     \
                    else if (++k >= n) { finish(ray, -1, 0, n); have = false; }                                                    \
                }
                ARV2_SCAN_STEP(r0) ARV2_SCAN_STEP(r1) ARV2_SCAN_STEP(r2) ARV2_SCAN_STEP(r3)
#undef ARV2_SCAN_STEP
            }
        }
        const unsigned fm = __ballot_sync(FULL, found);
        if (fm) {
            if (found) { cand[n_cand + __popc(fm & lt)] = make_int2(ray, k | (n << 16)); have = false; }
            n_cand += __popc(fm);
            __syncwarp();
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(FULL, segs, o);
    if (lane == 0 && segs) atomicAdd(p.counters + 1, segs);
}

// ---------------------------------------------------------------------------------------
// Packing the freshly traced cache: exclusive scan of the per-ray segment counts (three small kernels), then one warp
// per ray moves its records and energies to their CSR place and writes the quantised path vertices.
constexpr int kScanBlock = 256, kScanPerThread = 8, kScanTile = kScanBlock * kScanPerThread;

__global__ void __launch_bounds__(kScanBlock) pc_tile_sums_kernel(const int* __restrict__ nseg, long long n, unsigned long long* __restrict__ sums)
{
    __shared__ unsigned long long sh[kScanBlock / 32];
    const long long t0 = (long long)blockIdx.x * kScanTile;
    unsigned long long s = 0;
    for (int j = 0; j < kScanPerThread; ++j) { const long long i = t0 + (long long)j * kScanBlock + threadIdx.x; if (i < n) s += (unsigned)nseg[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int w = 0; w < kScanBlock / 32; ++w) t += sh[w]; sums[blockIdx.x] = t; }
}

__global__ void pc_scan_sums_kernel(unsigned long long* sums, long long n_tiles)      // one thread: a few thousand tiles
{
    unsigned long long run = 0;
    for (long long i = 0; i < n_tiles; ++i) { const unsigned long long v = sums[i]; sums[i] = run; run += v; }
    sums[n_tiles] = run;
}

__global__ void __launch_bounds__(kScanBlock) pc_offsets_kernel(const int* __restrict__ nseg, long long n, const unsigned long long* __restrict__ sums,
                                                                unsigned long long* __restrict__ off)
{
    __shared__ unsigned long long sh[kScanBlock];
    const long long i0 = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanPerThread;     // 8 consecutive rays per thread
    unsigned v[kScanPerThread];
    unsigned long long s = 0;
    for (int j = 0; j < kScanPerThread; ++j) { v[j] = i0 + j < n ? (unsigned)nseg[i0 + j] : 0u; s += v[j]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < kScanBlock; d <<= 1) {            // Hillis-Steele inclusive scan of the thread totals
        const unsigned long long add = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    unsigned long long run = sums[blockIdx.x] + sh[threadIdx.x] - s;
    for (int j = 0; j < kScanPerThread; ++j) { if (i0 + j < n) off[i0 + j] = run; run += v[j]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) off[n] = sums[gridDim.x];
}

__device__ __forceinline__ uint2 quantise_vertex(const TraceParams& p, float x, float y, float z, unsigned flags)
{
    const unsigned qx = (unsigned)fminf(fmaxf(rintf((x - p.pc_q0[0]) / p.pc_qs[0]), 0.f), 65535.f);
    const unsigned qy = (unsigned)fminf(fmaxf(rintf((y - p.pc_q0[1]) / p.pc_qs[1]), 0.f), 65535.f);
    const unsigned qz = (unsigned)fminf(fmaxf(rintf((z - p.pc_q0[2]) / p.pc_qs[2]), 0.f), 65535.f);
    return make_uint2(qx | (qy << 16), qz | flags);
}

template <int NB>
__global__ void __launch_bounds__(256) pc_compact_kernel(const TraceParams p, const float4* __restrict__ src_seg, const float* __restrict__ src_energy,
                                                         uint2* __restrict__ vert)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long ray = warp0; ray < p.n_rays; ray += n_warps) {
        const int n = (int)(p.pc_off[ray + 1] - p.pc_off[ray]);
        const unsigned long long o0 = p.pc_off[ray] + (unsigned long long)ray;      // records share the vertices' index space
        const size_t src = (size_t)ray * (size_t)p.pc_stride;
        uint2* const v = vert + o0;
        if (n == 0 && lane == 0) v[0] = quantise_vertex(p, p.emitter[0], p.emitter[1], p.emitter[2], 0u);
        for (int k = lane; k < n; k += 32) {
            const F8 rec = ldg256(src_seg + 2 * (src + k));
            stg256(p.pc_seg + 2 * (o0 + k), rec.lo, rec.hi);
#pragma unroll
            for (int b = 0; b < NB; ++b) p.pc_energy[(o0 + k) * NB + b] = src_energy[(src + k) * NB + b];
            const bool escapes = !(rec.lo.w < 1e20f);
            v[k] = quantise_vertex(p, rec.lo.x, rec.lo.y, rec.lo.z, kVertSegment | (escapes ? kVertEscapes : 0u));
            // the end point of the ray's last segment (every other segment ends where the next one starts, 1 mm off the wall)
            if (k == n - 1) v[n] = escapes ? make_uint2(0u, 0u) : quantise_vertex(p, fmaf(rec.lo.w, rec.hi.x, rec.lo.x), fmaf(rec.lo.w, rec.hi.y, rec.lo.y), fmaf(rec.lo.w, rec.hi.z, rec.lo.z), 0u);
        }
    }
}

__device__ __forceinline__ unsigned spread16(unsigned v)
{
    v = (v | (v << 8)) & 0x00FF00FFu; v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// Morton code of the octahedral map of each ray's emission direction: rays that are neighbours
// in this order leave the emitter in a narrow cone and stay together over the first bounces.
__global__ void direction_keys_kernel(unsigned long long seed, long long ray_begin, long long n, unsigned* __restrict__ keys,
                                      int* __restrict__ vals)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F3 d = emit_direction(seed, (uint64_t)(ray_begin + i));
    const float inv = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + 1e-30f);
    float u = d.x * inv, v = d.y * inv;
    if (d.z < 0.f) { const float uu = (1.f - fabsf(v)) * copysignf(1.f, u), vv = (1.f - fabsf(u)) * copysignf(1.f, v); u = uu; v = vv; }
    const unsigned iu = (unsigned)fminf(fmaxf((u * 0.5f + 0.5f) * 65536.f, 0.f), 65535.f);
    const unsigned iv = (unsigned)fminf(fmaxf((v * 0.5f + 0.5f) * 65536.f, 0.f), 65535.f);
    keys[i] = spread16(iu) | (spread16(iv) << 1);
    if (vals) vals[i] = (int)i;
}

// Direction tiles of the seeded ray set (multi-GPU sharding): the top tile_bits bits of a ray's direction key name one
// of 2^tile_bits cells of the octahedral map; rank r of R traces the rays of the tiles t = r (mod R), so every rank holds
// rays as dense in direction as the whole set (a contiguous slice of ids is 1/R as dense everywhere).  Grid-stride over
// the whole set: keys and ids of this rank's rays are appended (warp-aggregated) for the sort that follows.
__global__ void __launch_bounds__(256) direction_select_kernel(unsigned long long seed, long long n_total, int rank, int n_ranks, int tile_bits,
                                                               unsigned* __restrict__ keys, int* __restrict__ ids, unsigned long long* __restrict__ counter,
                                                               long long capacity)
{
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_round = (n_total + 31) / 32 * 32;          // whole warps take part in every ballot
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool mine = false;
        unsigned key = 0;
        if (i < n_total) {
            const F3 d = emit_direction(seed, (uint64_t)i);
            const float inv = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + 1e-30f);
            float u = d.x * inv, v = d.y * inv;
            if (d.z < 0.f) { const float uu = (1.f - fabsf(v)) * copysignf(1.f, u), vv = (1.f - fabsf(u)) * copysignf(1.f, v); u = uu; v = vv; }
            const unsigned iu = (unsigned)fminf(fmaxf((u * 0.5f + 0.5f) * 65536.f, 0.f), 65535.f);
            const unsigned iv = (unsigned)fminf(fmaxf((v * 0.5f + 0.5f) * 65536.f, 0.f), 65535.f);
            key = spread16(iu) | (spread16(iv) << 1);
            mine = (int)((key >> (32 - tile_bits)) % (unsigned)n_ranks) == rank;
        }
        const unsigned m = __ballot_sync(FULL, mine);
        if (m == 0) continue;
        unsigned long long base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(counter, (unsigned long long)__popc(m));
        base = __shfl_sync(FULL, base, __ffs(m) - 1);
        const long long pos = (long long)base + __popc(m & ((1u << lane) - 1u));
        if (mine && pos < capacity) { keys[pos] = key; ids[pos] = (int)i; }
    }
}

// Counting sort of (key, value) pairs on the top `bits` bits of 32-bit direction keys, in the three steps the sweeps use:
// a pass that takes every key's bin and its arrival rank in the bin (one atomic per key), the scan of the bins
// (sweep_tile_sums_kernel + sweep_scan_kernel), a scatter.  Within a bin the order is the arrival order -- a bin is
// 2^-bits of the sphere of directions, far below what a warp's bundle resolves -- so this replaces the four-pass stable
// radix sort of the ray order (410 us per 1 M rays) at a tenth of its cost.
__global__ void __launch_bounds__(256) order_bins_kernel(unsigned* __restrict__ keys, long long n, int bits, unsigned* __restrict__ rank, unsigned* __restrict__ bins)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const unsigned b = keys[i] >> (32 - bits);
    keys[i] = b;
    rank[i] = atomicAdd(bins + b, 1u);
}

__global__ void __launch_bounds__(256) order_scatter_kernel(const unsigned* __restrict__ key, const unsigned* __restrict__ rank, const unsigned* __restrict__ bins,
                                                            const int* __restrict__ vals, long long n, int* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    out[__ldg(bins + key[i]) + rank[i]] = vals ? vals[i] : (int)i;
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
#ifndef ARV2_TRACE_V2
    if (p.wave_paths) {
        wave_kernel<NB, MODE><<<(unsigned)sm_count, kWaveThreads, 0, stream>>>(p);
        return cudaGetLastError();
    }
    auto kernel = trace_kernel<NB, MODE>;
#else
    auto kernel = trace2_kernel<NB, MODE>;
#endif
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (const char* cap = getenv("ARV2_CTAS_PER_SM")) { const int v = atoi(cap); if (v > 0 && v < per_sm) per_sm = v; }   // tuning aid
    long long want = (p.n_rays + kThreads - 1) / kThreads;
    long long grid = (long long)sm_count * per_sm;
    if (want < grid) grid = want < 1 ? 1 : want;
    kernel<<<(unsigned)grid, kThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

} // namespace

int trace_supports_wide_nodes() { return ARV2_WIDE ? 1 : (ARV2_QNODES ? 2 : 0); }

cudaError_t launch_trace(const TraceParams& p, int bands, int mode, int sm_count, cudaStream_t stream)
{
    if (bands == 1) return mode == 0 ? launch_trace_t<1, 0>(p, sm_count, stream) : launch_trace_t<1, 1>(p, sm_count, stream);
    if (bands == 8) return mode == 0 ? launch_trace_t<8, 0>(p, sm_count, stream) : launch_trace_t<8, 1>(p, sm_count, stream);
    return cudaErrorInvalidValue;
}

template <int NB, int MODE>
cudaError_t launch_sweeps_t(const TraceParams& p, const SweepWork& work, cudaStream_t stream)
{
    const int n_bins = sweep_bins(work.cell_bits, work.dir_bits);
    // the first sweep keeps the direction-sorted bundles of fresh rays together for first_segments segments (a bundle
    // of neighbours is far more coherent than any bin), the others advance by `segments`
    const long long first = work.first_segments < 1 ? 1 : work.first_segments;
    const long long rest = (long long)p.max_bounces - first;
    const long long n_sweeps = 1 + (rest > 0 ? (rest + work.segments - 1) / work.segments : 0);
    const unsigned grid = (unsigned)((p.n_rays + kSweepThreads - 1) / kSweepThreads), grid_sc = (unsigned)((p.n_rays + 255) / 256);
    SweepParams w{};
    w.key = work.key; w.rank = work.rank; w.bins = work.bins;
    w.cell_bits = work.cell_bits; w.dir_bits = work.dir_bits; w.dir_major = work.dir_major;
    for (int a = 0; a < 3; ++a) { w.lo[a] = work.lo[a]; w.scale[a] = work.scale[a]; }
    cudaError_t e = cudaSuccess;
    for (long long k = 0; k < n_sweeps && e == cudaSuccess; ++k) {
        const int cur = (int)(k & 1), nxt = cur ^ 1;
        w.segments = k == 0 ? (int)first : work.segments;
        w.in = k == 0 ? nullptr : work.state[cur];
        w.perm = k == 0 ? nullptr : work.perm;
        w.n_in = work.count + cur;
        w.out = work.state[nxt];
        w.n_out = work.count + nxt;
        if ((e = cudaMemsetAsync(work.bins, 0, (size_t)n_bins * sizeof(unsigned), stream)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(work.count + nxt, 0, sizeof(unsigned long long), stream)) != cudaSuccess) break;
        sweep_kernel<NB, MODE><<<grid, kSweepThreads, 0, stream>>>(p, w);
        if (k + 1 < n_sweeps) {
            sweep_tile_sums_kernel<<<n_bins / 4096, 1024, 0, stream>>>(work.bins, work.tile_sums);
            sweep_scan_kernel<<<n_bins / 4096, 1024, 0, stream>>>(work.bins, work.tile_sums);
            sweep_scatter_kernel<<<grid_sc, 256, 0, stream>>>(work.key, work.rank, work.bins, work.count + nxt, work.perm);
        }
        e = cudaGetLastError();
    }
    return e;
}

cudaError_t launch_trace_sweeps(const TraceParams& p, const SweepWork& work, int bands, int mode, cudaStream_t stream)
{
    if (p.n_rays <= 0) return cudaSuccess;
    if (bands == 1) return mode == 0 ? launch_sweeps_t<1, 0>(p, work, stream) : launch_sweeps_t<1, 1>(p, work, stream);
    if (bands == 8) return mode == 0 ? launch_sweeps_t<8, 0>(p, work, stream) : launch_sweeps_t<8, 1>(p, work, stream);
    return cudaErrorInvalidValue;
}

cudaError_t launch_rerender(const TraceParams& p, int bands, int sm_count, cudaStream_t stream)
{
    if (p.n_rays == 0) return cudaSuccess;
    if (bands != 1 && bands != 8) return cudaErrorInvalidValue;
    if (!p.pc_vert || !p.pc_bits || !p.pc_off) return cudaErrorInvalidValue;
    const long long tiles = (p.pc_nvert + kMaskTile - 1) / kMaskTile;      // one warp per tile and trip
    long long grid = (tiles + kRrThreads / 32 - 1) / (kRrThreads / 32);
    const long long cap = (long long)sm_count * (2048 / kRrThreads) * 2;
    if (grid > cap) grid = cap;
    rr_mask_kernel<<<(unsigned)grid, kRrThreads, 0, stream>>>(p);
    long long wgrid = (p.n_rays + kRrThreads - 1) / kRrThreads;
    long long wcap = (long long)sm_count * ARV2_RRW_MINB;
    if (const char* e = getenv("ARV2_RRW_CTAS")) wcap = (long long)sm_count * (atoi(e) > 0 ? atoi(e) : ARV2_RRW_MINB);      // tuning aid
    if (wgrid > wcap) wgrid = wcap;
    const size_t smem = (size_t)p.recv_nodes_shared * 64;
    if (bands == 1) rr_walk_kernel<1><<<(unsigned)wgrid, kRrThreads, smem, stream>>>(p);
    else rr_walk_kernel<8><<<(unsigned)wgrid, kRrThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_rerender_serial(const TraceParams& p, int bands, int sm_count, cudaStream_t stream)
{
    if (p.n_rays == 0) return cudaSuccess;
    if (bands != 1 && bands != 8) return cudaErrorInvalidValue;
    int per_sm = 0;
    cudaError_t e = bands == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rerender_kernel<1>, kRerenderThreads, 0)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rerender_kernel<8>, kRerenderThreads, 0);
    if (e != cudaSuccess) return e;
    long long grid = (long long)sm_count * (per_sm < 1 ? 1 : per_sm);
    const long long want = (p.n_rays + kRerenderThreads - 1) / kRerenderThreads;
    if (want < grid) grid = want;
    if (bands == 1) rerender_kernel<1><<<(unsigned)grid, kRerenderThreads, 0, stream>>>(p);
    else rerender_kernel<8><<<(unsigned)grid, kRerenderThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_cache_offsets(const int* nseg, long long n_rays, unsigned long long* off, unsigned long long* scratch, cudaStream_t stream)
{
    if (n_rays <= 0) return cudaSuccess;
    const long long tiles = (n_rays + kScanTile - 1) / kScanTile;
    pc_tile_sums_kernel<<<(unsigned)tiles, kScanBlock, 0, stream>>>(nseg, n_rays, scratch);
    pc_scan_sums_kernel<<<1, 1, 0, stream>>>(scratch, tiles);
    pc_offsets_kernel<<<(unsigned)tiles, kScanBlock, 0, stream>>>(nseg, n_rays, scratch, off);
    return cudaGetLastError();
}

cudaError_t launch_cache_compact(const TraceParams& p, const float4* src_seg, const float* src_energy, uint2* vert, int bands, cudaStream_t stream)
{
    if (p.n_rays <= 0) return cudaSuccess;
    long long grid = (p.n_rays * 32 + 255) / 256;
    if (grid > 148LL * 64) grid = 148LL * 64;
    if (bands == 1) pc_compact_kernel<1><<<(unsigned)grid, 256, 0, stream>>>(p, src_seg, src_energy, vert);
    else if (bands == 8) pc_compact_kernel<8><<<(unsigned)grid, 256, 0, stream>>>(p, src_seg, src_energy, vert);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_direction_keys(unsigned long long seed, long long ray_begin, long long n, unsigned* keys, int* vals, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    direction_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(seed, ray_begin, n, keys, vals);
    return cudaGetLastError();
}

cudaError_t launch_direction_select(unsigned long long seed, long long n_total, int rank, int n_ranks, int tile_bits, unsigned* keys, int* ids,
                                    unsigned long long* counter, long long capacity, int sm_count, cudaStream_t stream)
{
    if (n_total <= 0) return cudaSuccess;
    long long grid = (n_total + 255) / 256;
    if (grid > (long long)sm_count * 16) grid = (long long)sm_count * 16;
    direction_select_kernel<<<(unsigned)grid, 256, 0, stream>>>(seed, n_total, rank, n_ranks, tile_bits, keys, ids, counter, capacity);
    return cudaGetLastError();
}

int counting_order_bits(long long n)
{
    // four bins per key up to 1 M keys, at most 2^22 bins (8 keys per bin cost 1.1 % of the trace rate at 1 M rays, one per
    // bin 0.45 %: the order inside a bin is the arrival order), 4096 bins at least (the scan works on tiles of 4096)
    int bits = 12;
    while (bits < kCountingOrderMaxBits && (n >> bits) > 0) ++bits;
    return bits + 2 > kCountingOrderMaxBits ? kCountingOrderMaxBits : bits + 2;
}

cudaError_t launch_counting_order(unsigned* keys, const int* vals, long long n, int bits, unsigned* rank, unsigned* bins, unsigned* tile_sums, int* out,
                                  cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int n_bins = 1 << bits;
    cudaError_t e = cudaMemsetAsync(bins, 0, (size_t)n_bins * sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((n + 255) / 256);
    order_bins_kernel<<<grid, 256, 0, stream>>>(keys, n, bits, rank, bins);
    sweep_tile_sums_kernel<<<n_bins / 4096, 1024, 0, stream>>>(bins, tile_sums);
    sweep_scan_kernel<<<n_bins / 4096, 1024, 0, stream>>>(bins, tile_sums);
    order_scatter_kernel<<<grid, 256, 0, stream>>>(keys, rank, bins, vals, n, out);
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

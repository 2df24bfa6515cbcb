// arv2_model.cuh -- the per-ray arithmetic contract on the device (sm_100a).
//
// Every operation that can move a ray by one ulp is spelled with a round-to-nearest
// intrinsic (__fmul_rn / __fadd_rn / __fmaf_rn / __fdiv_rn / __fsqrt_rn and the fp64
// twins), which nvcc never contracts or approximates, so the kernels produce the same
// bits on every launch geometry and under any -fmad / -use_fast_math setting.
// The formulas restate OR/devicePrograms.cu:62-254 (see DESIGN.md "arithmetic
// contract" for the op order and for what replaces OptiX's triangle test and cuRAND).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace arv2 {

struct F3 { float x, y, z; };

__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 sub3(F3 a, F3 b) { return f3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ float dot3(F3 a, F3 b)
{
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x)));
}
__device__ __forceinline__ F3 cross3(F3 a, F3 b)
{
    return f3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
              __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

// Philox4x32-10, counter = (ray_lo, ray_hi, bounce, purpose), key = (seed_lo, seed_hi).
// Replaces curand_init(clock64(), tid, 0) (OR/devicePrograms.cu:216-217).
__device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t ray, uint32_t bounce, uint32_t purpose, uint32_t r[4])
{
    uint32_t c0 = (uint32_t)ray, c1 = (uint32_t)(ray >> 32), c2 = bounce, c3 = purpose;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

// (cos, sin) of 2*pi*(k+0.5)/2^32: integer quadrant reduction + fp64 Taylor
// polynomials in fma form (bit-identical on any IEEE machine; no libm).
__device__ __forceinline__ void sincos_turn(uint32_t k, double* c_out, double* s_out)
{
    const uint32_t q = k >> 30;
    double f = __dmul_rn(__dadd_rn((double)(k & 0x3FFFFFFFu), 0.5), 0x1p-30);
    const bool swap = f > 0.5;
    if (swap) f = __dsub_rn(1.0, f);
    const double b = __dmul_rn(f, 0x1.921fb54442d18p+0);
    const double b2 = __dmul_rn(b, b);
    double ps = 0x1.952c77030ad4ap-49;
    ps = __fma_rn(ps, b2, -0x1.ae7f3e733b81fp-41);
    ps = __fma_rn(ps, b2, 0x1.6124613a86d09p-33);
    ps = __fma_rn(ps, b2, -0x1.ae64567f544e4p-26);
    ps = __fma_rn(ps, b2, 0x1.71de3a556c734p-19);
    ps = __fma_rn(ps, b2, -0x1.a01a01a01a01ap-13);
    ps = __fma_rn(ps, b2, 0x1.1111111111111p-7);
    ps = __fma_rn(ps, b2, -0x1.5555555555555p-3);
    double s = __fma_rn(__dmul_rn(b, b2), ps, b);
    double pc = 0x1.ae7f3e733b81fp-45;
    pc = __fma_rn(pc, b2, -0x1.93974a8c07c9dp-37);
    pc = __fma_rn(pc, b2, 0x1.1eed8eff8d898p-29);
    pc = __fma_rn(pc, b2, -0x1.27e4fb7789f5cp-22);
    pc = __fma_rn(pc, b2, 0x1.a01a01a01a01ap-16);
    pc = __fma_rn(pc, b2, -0x1.6c16c16c16c17p-10);
    pc = __fma_rn(pc, b2, 0x1.5555555555555p-5);
    pc = __fma_rn(pc, b2, -0x1p-1);
    double c = __fma_rn(pc, b2, 1.0);
    if (swap) { const double t = s; s = c; c = t; }
    if (q == 0) { *c_out = c; *s_out = s; }
    else if (q == 1) { *c_out = -s; *s_out = c; }
    else if (q == 2) { *c_out = -c; *s_out = -s; }
    else { *c_out = s; *s_out = -c; }
}

// OR/devicePrograms.cu:219-224: uniform direction, theta = 2*pi*u1, cos(phi) = 2*u2-1.
__device__ __forceinline__ F3 emit_direction(uint64_t seed, uint64_t ray)
{
    uint32_t r[4];
    philox4x32(seed, ray, 0u, 0u, r);
    double ct, st;
    sincos_turn(r[0], &ct, &st);
    const double u2 = __dmul_rn((double)((r[1] >> 8) + 1u), 0x1p-24);
    const double z = __dsub_rn(__dmul_rn(2.0, u2), 1.0);
    const double sp = __dsqrt_rn(__fma_rn(-z, z, 1.0));
    return f3(__double2float_rn(__dmul_rn(sp, ct)), __double2float_rn(__dmul_rn(sp, st)), __double2float_rn(z));
}

// Lambert bounce (extension; scattering == 0 never reaches this).
__device__ __forceinline__ F3 lambert_direction(const uint32_t r[4], F3 dir, F3 ng)
{
    F3 n = ng;
    if (dot3(dir, ng) > 0.0f) n = f3(-ng.x, -ng.y, -ng.z);
    double cp, sp;
    sincos_turn(r[2], &cp, &sp);
    const double u = __dmul_rn(__dadd_rn((double)(r[1] >> 8), 0.5), 0x1p-24);
    const double sr = __dsqrt_rn(u), cz = __dsqrt_rn(__dsub_rn(1.0, u));
    const float lx = __double2float_rn(__dmul_rn(sr, cp)), ly = __double2float_rn(__dmul_rn(sr, sp));
    const float lz = __double2float_rn(cz);
    const float sg = copysignf(1.0f, n.z);
    const float a = __fdiv_rn(-1.0f, __fadd_rn(sg, n.z));
    const float b = __fmul_rn(__fmul_rn(n.x, n.y), a);
    const F3 t1 = f3(__fmaf_rn(__fmul_rn(sg, n.x), __fmul_rn(n.x, a), 1.0f), __fmul_rn(sg, b), __fmul_rn(-sg, n.x));
    const F3 t2 = f3(b, __fmaf_rn(n.y, __fmul_rn(n.y, a), sg), -n.y);
    return f3(__fmaf_rn(lz, n.x, __fmaf_rn(ly, t2.x, __fmul_rn(lx, t1.x))),
              __fmaf_rn(lz, n.y, __fmaf_rn(ly, t2.y, __fmul_rn(lx, t1.y))),
              __fmaf_rn(lz, n.z, __fmaf_rn(ly, t2.z, __fmul_rn(lx, t1.z))));
}

// Two-sided Moller-Trumbore on (P1, P2-P1, P3-P1), scaled form; the three divisions
// happen only for accepted hits.  Stands in for optixTrace's triangle test
// (OR/devicePrograms.cu:240-251): hit iff 0 < t < 1e20.
__device__ __forceinline__ bool tri_test(F3 p1, F3 p2, F3 p3, F3 org, F3 dir, float* t, float* u, float* v)
{
    const F3 e1 = sub3(p2, p1), e2 = sub3(p3, p1);
    const F3 pvec = cross3(dir, e2);
    float det = dot3(e1, pvec);
    if (det == 0.0f) return false;
    const F3 tvec = sub3(org, p1);
    float U = dot3(tvec, pvec);
    const F3 qvec = cross3(tvec, e1);
    float V = dot3(dir, qvec);
    float T = dot3(e2, qvec);
    if (det < 0.0f) { det = -det; U = -U; V = -V; T = -T; }
    if (!(U >= 0.0f && V >= 0.0f && __fadd_rn(U, V) <= det)) return false;
    if (!(T > 0.0f)) return false;
    const float tt = __fdiv_rn(T, det);
    if (!(tt < 1e20f)) return false;
    *t = tt; *u = __fdiv_rn(U, det); *v = __fdiv_rn(V, det);
    return true;
}

} // namespace arv2

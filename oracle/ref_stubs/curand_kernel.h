/* oracle/ref_stubs/curand_kernel.h -- TEST INFRASTRUCTURE.  Stand-in for cuRAND's device API (OR/devicePrograms.cu:
 * 216-220 seeds XORWOW with clock64(), so no reproducible stream exists to restate): curand_uniform hands out the two
 * uniforms the shim's caller supplies for this ray, in call order (theta first, then phi). */
#pragma once
struct curandState { int tid, calls; };
extern thread_local const float* ref_uniforms;   /* [2] for the ray being generated */
static inline void curand_init(unsigned long long, unsigned long long sequence, unsigned long long, curandState* s) { s->tid = (int)sequence; s->calls = 0; }
static inline float curand_uniform(curandState* s) { return ref_uniforms[s->calls++ & 1]; }

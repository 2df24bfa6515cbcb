/* oracle/ref_stubs/optix_device.h -- TEST INFRASTRUCTURE.  Host-side stand-in for the OptiX 7.7 device API so that
 * oracle/ref_device_shim.cpp can compile OR/devicePrograms.cu (the reference's raygen / closest-hit / miss programs)
 * UNMODIFIED, with g++, from where it lies.  The "device state" one program invocation sees (launch index, SBT record,
 * payload registers, hit attributes) lives in `ref_optix_state`; optixTrace() hands the ray to the shim's own
 * closest-hit search and then runs the reference's closest-hit or miss program on the result -- which is what the
 * OptiX pipeline does (OR/AudioRenderer.cpp:325-411).  Nothing here implements shading arithmetic. */
#pragma once
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

struct ref_optix_state_t {
    uint3 launch_index, launch_dims;
    const void* sbt_data;           /* TriangleMeshSBTData of the hit mesh */
    float3 ray_dir;                 /* optixGetWorldRayDirection          */
    float2 bary;                    /* optixGetTriangleBarycentrics       */
    int prim;                       /* optixGetPrimitiveIndex             */
    uint32_t p0, p1;                /* payload registers                  */
    uint64_t trace_calls;
};
extern thread_local ref_optix_state_t ref_optix_state;

static inline uint3 optixGetLaunchIndex() { return ref_optix_state.launch_index; }
static inline uint3 optixGetLaunchDimensions() { return ref_optix_state.launch_dims; }
static inline unsigned long long optixGetSbtDataPointer() { return (unsigned long long)ref_optix_state.sbt_data; }
static inline float3 optixGetWorldRayDirection() { return ref_optix_state.ray_dir; }
static inline float2 optixGetTriangleBarycentrics() { return ref_optix_state.bary; }
static inline unsigned int optixGetPrimitiveIndex() { return (unsigned)ref_optix_state.prim; }
static inline uint32_t optixGetPayload_0() { return ref_optix_state.p0; }
static inline uint32_t optixGetPayload_1() { return ref_optix_state.p1; }

/* the shim's closest-hit search + program dispatch (oracle/ref_device_shim.cpp) */
void ref_trace_dispatch(float ox, float oy, float oz, float dx, float dy, float dz, float tmin, float tmax, uint32_t p0, uint32_t p1);

template <class V>
static inline void optixTrace(unsigned long long, const V& org, const V& dir, float tmin, float tmax, float, unsigned, unsigned,
                              unsigned, unsigned, unsigned, uint32_t& p0, uint32_t& p1)
{
    ref_trace_dispatch(org.x, org.y, org.z, dir.x, dir.y, dir.z, tmin, tmax, p0, p1);
}

/* CUDA device built-ins the programs call */
void ref_atomic_add(float* addr, float v);      /* records the deposit, then *addr += v */
static inline float atomicAdd(float* addr, float v) { const float old = *addr; ref_atomic_add(addr, v); return old; }
static inline long long clock64() { return 0; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }

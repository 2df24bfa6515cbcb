// Microbenchmark: cost of divergent (one record per lane) gathers from an L2-resident
// table with 32-, 128- and 256-bit loads.  Guides the BVH node layout (how many load
// instructions a node may cost).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 gather.cu
#include <cstdio>
#include <cuda_runtime.h>
struct __align__(32) F8 { float a[8]; };
__device__ __forceinline__ F8 ld256(const void* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a[0]),"=f"(r.a[1]),"=f"(r.a[2]),"=f"(r.a[3]),"=f"(r.a[4]),"=f"(r.a[5]),"=f"(r.a[6]),"=f"(r.a[7]) : "l"(p));
    return r;
}
// MODE 0: 4 x LDG.128 per 64 B record, 1: 2 x LDG.256, 2: 1 x LDG.128 (16 B), 3: 1 x LDG.32, 4: 2 x LDG.128 (32 B), 5: 3 x LDG.128 (48 B),
//      6: 1 x LDG.256 (32 B, one sector)
template <int MODE>
__global__ void gather(const float4* __restrict__ tab, unsigned n_rec, int iters, float* out)
{
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const unsigned r = (s >> 8) % n_rec;
        const float4* p = tab + 4 * (size_t)r;
        if (MODE == 0) { float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3); acc += a.x + b.y + c.z + d.w; }
        if (MODE == 1) { F8 a = ld256(p), b = ld256(p + 2); acc += a.a[0] + a.a[5] + b.a[2] + b.a[7]; }
        if (MODE == 2) { float4 a = __ldg(p); acc += a.x + a.w; }
        if (MODE == 3) { acc += __ldg((const float*)p); }
        if (MODE == 4) { float4 a = __ldg(p), b = __ldg(p + 1); acc += a.x + b.y; }
        if (MODE == 6) { F8 a = ld256(p); acc += a.a[0] + a.a[7]; }
        if (MODE == 5) { float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2); acc += a.x + b.y + c.z; }
        s += __float_as_uint(acc) & 1u;    // dependent chain like a traversal
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, const float4* tab, unsigned n_rec, float* out)
{
    const int iters = 2000, grid = 148 * 4, block = 256;
    gather<MODE><<<grid, block>>>(tab, n_rec, 200, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    gather<MODE><<<grid, block>>>(tab, n_rec, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double recs = (double)grid * block * iters;
    printf("%-22s %8.3f ms  %7.2f G records/s\n", name, ms, recs / ms / 1e6);
}
int main()
{
    for (unsigned mb : {8u, 24u, 96u}) {
        const unsigned n_rec = mb * 1024 * 1024 / 64;
        float4* tab; float* out;
        cudaMalloc(&tab, (size_t)n_rec * 64); cudaMemset(tab, 0, (size_t)n_rec * 64);
        cudaMalloc(&out, 148 * 4 * 256 * 4);
        printf("table %u MB\n", mb);
        run<0>("64B: 4 x LDG.128", tab, n_rec, out);
        run<1>("64B: 2 x LDG.256", tab, n_rec, out);
        run<5>("48B: 3 x LDG.128", tab, n_rec, out);
        run<6>("32B: 1 x LDG.256", tab, n_rec, out);
        run<4>("32B: 2 x LDG.128", tab, n_rec, out);
        run<2>("16B: 1 x LDG.128", tab, n_rec, out);
        run<3>(" 4B: 1 x LDG.32", tab, n_rec, out);
        cudaFree(tab); cudaFree(out);
    }
    return 0;
}

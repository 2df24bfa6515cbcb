/* oracle/ref_stubs/optix.h -- TEST INFRASTRUCTURE.  Host-side stand-in for the OptiX 7.7 SDK header, which is absent
 * here: just enough opaque types for the reference's headers (OR/AudioRenderer.h:89-127, OR/LaunchParams.h:30) to
 * parse when oracle/Makefile compiles reference sources where they lie.  No OptiX behaviour is implemented. */
#pragma once
#include <stddef.h>
#include <stdint.h>
typedef unsigned long long OptixTraversableHandle;
typedef struct OptixDeviceContext_t* OptixDeviceContext;
typedef struct OptixPipeline_t* OptixPipeline;
typedef struct OptixModule_t* OptixModule;
typedef struct OptixProgramGroup_t* OptixProgramGroup;
typedef struct { int unused; } OptixPipelineCompileOptions;
typedef struct { int unused; } OptixPipelineLinkOptions;
typedef struct { int unused; } OptixModuleCompileOptions;
typedef struct { int unused; } OptixShaderBindingTable;
typedef enum { OPTIX_SUCCESS = 0 } OptixResult;
typedef unsigned int OptixVisibilityMask;
enum { OPTIX_RAY_FLAG_NONE = 0, OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1 };
#ifndef CUDA_VERSION
typedef struct CUctx_st* CUcontext;
typedef struct CUstream_st* CUstream;
typedef unsigned long long CUdeviceptr;
#endif

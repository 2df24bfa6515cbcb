// comm.cuh -- NCCL entry points of libarv2.so, bound at run time (dlopen), so that the library has no link-time
// dependency on one particular libnccl: inside a process that already holds NCCL (a torch rank) that copy is used,
// a stand-alone C++ host (arv2_cli) loads the system's libnccl.so.2.
#pragma once

#include <string>

#include <cuda_runtime.h>
#include <nccl.h>

namespace arv2 {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int*);
};

// The bound entry points, or nullptr with *err set (no libnccl.so.2 to be found).  Order: ARV2_NCCL_LIB, a copy the
// process has loaded already, the dynamic linker's libnccl.so.2.
const NcclApi* nccl_api(std::string* err);

} // namespace arv2

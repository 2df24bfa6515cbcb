// comm.cu -- run-time binding of NCCL (comm.cuh).
#include <cstdlib>
#include <mutex>

#include <dlfcn.h>

#include "comm.cuh"

namespace arv2 {

const NcclApi* nccl_api(std::string* err)
{
    static NcclApi api{};
    static bool ok = false;
    static std::string why;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        if (const char* path = std::getenv("ARV2_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // already in the process (e.g. torch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { why = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen(libnccl.so.2) failed"); return; }
        bool all = true;
        auto bind = [&](const char* name) { void* s = dlsym(h, name); if (!s) { all = false; why = std::string("NCCL symbol missing: ") + name; } return s; };
        api.GetUniqueId = (decltype(api.GetUniqueId))bind("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))bind("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))bind("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))bind("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))bind("ncclAllReduce");
        api.Reduce = (decltype(api.Reduce))bind("ncclReduce");
        api.GroupStart = (decltype(api.GroupStart))bind("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))bind("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))bind("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))bind("ncclGetVersion");
        ok = all;
    });
    if (!ok) { if (err) *err = why; return nullptr; }
    return &api;
}

} // namespace arv2

// nccl_dyn.h -- NCCL bound at run time (dlopen "libnccl.so.2"), so that libsjpeg_b200.so neither
// links against NCCL nor needs it for single-GPU use.  In a process that has already loaded an
// NCCL (torch's bundled one) the same library instance is picked up by SONAME.  Only the handful
// of calls the row-stripe exchange needs (engine_stripes.inl); declarations come from <nccl.h>.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

namespace sjb {

struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
};

inline const NcclApi* Nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return;
    bool all = true;
    auto get = [&](const char* name) -> void* {
      void* p = dlsym(h, name);
      if (p == nullptr) all = false;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(get("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(get("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(get("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(get("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(get("ncclAllReduce"));
    api.Send = reinterpret_cast<decltype(api.Send)>(get("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(get("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(get("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(get("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(get("ncclGetErrorString"));
    api.ok = all;
  });
  return api.ok ? &api : nullptr;
}

}  // namespace sjb

// nccl_dyn.h — NCCL bound at run time (dlopen), so the library loads on a single GPU without libnccl and
// shares the NCCL instance a host process (e.g. torch) may already have loaded. Only the five entry points
// the path needs: the camera-side blocks, the partial Schur product and a few scalars are all-reduced.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace apex {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
constexpr int NCCL_DOUBLE = 8;  // ncclFloat64
constexpr int NCCL_SUM = 0;     // ncclSum

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok() const { return GetUniqueId && CommInitRank && AllReduce && CommDestroy; }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  // a copy already mapped into the process (torch's bundled one) wins over the system library
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    void* h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) continue;
    api.handle = h;
    break;
  }
  if (!api.handle)
    for (const char* n : names) {
      void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (!h) continue;
      api.handle = h;
      break;
    }
  if (!api.handle) return api;
  api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
  api.CommDestroy = (int (*)(NcclComm))dlsym(api.handle, "ncclCommDestroy");
  api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
  return api;
}

}  // namespace apex

// nccl_dyn.h -- NCCL bound at run time with dlopen, so libnls_b200.so has no link-time
// dependency on a particular libnccl (the host process usually already has one loaded:
// torch's bundled 2.28.x, or the system 2.27.x; dlopen by soname returns that copy).
// Only the stable core of the NCCL C API is used.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <string>

#define NCCL_DYN_FLOAT64 8   // ncclFloat64
#define NCCL_DYN_SUM 0       // ncclSum
#define NCCL_DYN_ID_BYTES 128

struct NcclDynId { char internal[NCCL_DYN_ID_BYTES]; };

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(NcclDynId *) = nullptr;
  int (*CommInitRank)(void **comm, int nranks, NcclDynId id, int rank) = nullptr;
  int (*CommDestroy)(void *comm) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *comm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t sendcount, int, void *comm, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int peer, void *comm, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int peer, void *comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;

  // returns "" on success, else an error message
  std::string load(const char *path) {
    if (lib) return "";
    const char *cands[] = {path, "libnccl.so.2", "libnccl.so"};
    for (const char *c : cands) {
      if (!c || !*c) continue;
      lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return std::string("cannot dlopen libnccl: ") + dlerror();
#define NLS_SYM(field, name) \
    *(void **)(&field) = dlsym(lib, name); \
    if (!field) return std::string("libnccl lacks symbol ") + name;
    NLS_SYM(GetUniqueId, "ncclGetUniqueId")
    NLS_SYM(CommInitRank, "ncclCommInitRank")
    NLS_SYM(CommDestroy, "ncclCommDestroy")
    NLS_SYM(AllReduce, "ncclAllReduce")
    NLS_SYM(AllGather, "ncclAllGather")
    NLS_SYM(Send, "ncclSend")
    NLS_SYM(Recv, "ncclRecv")
    NLS_SYM(GroupStart, "ncclGroupStart")
    NLS_SYM(GroupEnd, "ncclGroupEnd")
    NLS_SYM(GetErrorString, "ncclGetErrorString")
#undef NLS_SYM
    return "";
  }
};

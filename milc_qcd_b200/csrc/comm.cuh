// comm.cuh -- NCCL binding (dlopen, so single-GPU users carry no NCCL dependency) and the
// communicator state of a one-rank-per-GPU context.
//
// The reference exchanges halos with MPI point-to-point "gathers" re-armed every CG
// iteration (generic/com_mpi.c, generic_ks/d_congrad5_fn_milc.c:264-272) and sums scalars with
// MPI_Allreduce (g_doublesum / g_vecdoublesum).  Here every rank is one B200 of an NVSwitch
// box: halos are ncclSend/ncclRecv groups over NVLink on a dedicated stream, overlapped with
// the interior stencil pass; scalars are ncclAllReduce on the compute stream.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace b200ks {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclChar = 0, ncclFloat64 = 8, ncclDouble = 8 };  // nccl.h ncclDataType_t
enum { ncclSum = 0, ncclMax = 2 };  // nccl.h ncclRedOp_t

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};

inline NcclApi &nccl() {
  static NcclApi api;
  if (api.handle) return api;
  // reuse a libnccl the process already holds (torch ships its own) before loading the system one
  api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api.handle) return api;
#define B200KS_SYM(field, name) *(void **)(&api.field) = dlsym(api.handle, name)
  B200KS_SYM(GetUniqueId, "ncclGetUniqueId");
  B200KS_SYM(CommInitRank, "ncclCommInitRank");
  B200KS_SYM(CommDestroy, "ncclCommDestroy");
  B200KS_SYM(Send, "ncclSend");
  B200KS_SYM(Recv, "ncclRecv");
  B200KS_SYM(AllReduce, "ncclAllReduce");
  B200KS_SYM(GroupStart, "ncclGroupStart");
  B200KS_SYM(GroupEnd, "ncclGroupEnd");
  B200KS_SYM(GetErrorString, "ncclGetErrorString");
#undef B200KS_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.AllReduce &&
           api.GroupStart && api.GroupEnd;
  return api;
}

struct Comm {
  int rank = 0, nranks = 1;
  int grid[4] = {1, 1, 1, 1};
  int coord[4] = {0, 0, 0, 0};
  int nbr[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};  // rank of (d, 0=backward | 1=forward) neighbour
  ncclComm_t halo = nullptr;     // point-to-point halos, comm stream
  ncclComm_t red = nullptr;      // all-reduces, compute stream
  cudaStream_t stream = nullptr; // comm stream
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  void *ghost[2] = {nullptr, nullptr};   // ghost buffers (double-sized), ping-pong by dslash
  void *zsend = nullptr;                 // packed z faces
  int *ext_sites = nullptr;              // boundary-site list for the exterior pass
  int n_ext = 0;
  int flip = 0;
};

}  // namespace b200ks

// comm.cuh -- NCCL binding (dlopen, so single-GPU users carry no NCCL dependency) and the
// communicator state of a one-rank-per-GPU context.
//
// The reference exchanges halos with MPI point-to-point "gathers" re-armed every CG
// iteration (generic/com_mpi.c, generic_ks/d_congrad5_fn_milc.c:264-272) and sums scalars with
// MPI_Allreduce (g_doublesum / g_vecdoublesum).  Here every rank is one B200 of an NVSwitch
// box.  Halos: every rank maps its neighbours' ghost buffers (CUDA IPC) and a push kernel
// stores its boundary slices straight into them over NVLink, then raises an arrival flag in the
// neighbour's memory; the consumer waits on its own flags with a one-warp kernel in front of the
// exterior stencil pass.  No copy engine, no NCCL proxy, no host involvement per exchange.
// (ncclSend/ncclRecv groups remain as the fallback when peer mapping is unavailable, and for
// the one-time link-ghost exchange.)  Scalars are ncclAllReduce on the compute stream.
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
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
  B200KS_SYM(AllGather, "ncclAllGather");
  B200KS_SYM(GroupStart, "ncclGroupStart");
  B200KS_SYM(GroupEnd, "ncclGroupEnd");
  B200KS_SYM(GetErrorString, "ncclGetErrorString");
#undef B200KS_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.AllReduce && api.AllGather &&
           api.GroupStart && api.GroupEnd;
  return api;
}

// Peer-to-peer state.  One cudaMalloc block per rank, mapped by EVERY other rank (CUDA IPC between
// processes, plain peer access inside one process):
//   [ halo arrival flags: 4 x u64 (z-behind, z-ahead, t-behind, t-ahead), padded to 256 B ]
//   [ RedBox: mailboxes of the flag-based all-reduce, up to kP2PHeaderBytes ]
//   [ ghost buffer 0 ][ ghost buffer 1 ]     each 3 colours x gstride sites x sizeof(double2)
// Ghost buffers alternate by exchange sequence number: a neighbour can be at most one
// exchange ahead (its next push needs our halo of the current one), so two buffers suffice.
//
// All-reduce without a collective library (blas.cuh p2p_allreduce_block): every rank stores its
// <= 8 partial values into slot [its rank] of EVERY rank's mailbox over NVLink, then raises
// flag[its rank] there to the reduction's sequence number; it then waits until all of its own
// flags have reached that number and adds the nranks slots in rank order -- the same order on
// every rank, so all ranks hold bit-identical sums.  Two mailbox sets alternate by sequence
// number: completing reduction k needs every rank's delivery of k, which a rank makes only after
// it has finished reading k-1, so a rank is never more than one reduction ahead of another.
constexpr int kMaxRanks = 16;
struct RedBox {
  unsigned long long flag[kMaxRanks];   // flag[q]: number of reductions rank q has delivered here
  unsigned long long count;             // reductions this rank has completed (only its own kernels write it)
  unsigned long long pad_[15];
  double val[2][kMaxRanks][8];
};
constexpr size_t kP2PFlagBytes = 256;
constexpr size_t kP2PHeaderBytes = 4096;
static_assert(kP2PFlagBytes + sizeof(RedBox) <= kP2PHeaderBytes, "RedBox does not fit the block header");

struct P2P {
  bool on = false;
  char *block = nullptr;            // own block
  size_t ghost_bytes = 0;           // bytes of one ghost buffer
  char *peer_all[kMaxRanks] = {};   // every rank's block as mapped here (own rank: block)
  char *peer_block[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [d-2][0 backward | 1 forward] neighbour's block
  void *opened[kMaxRanks] = {};     // cudaIpcOpenMemHandle results to close
  unsigned long long seq = 0;       // exchange sequence number (host mirror)
  unsigned *ticket = nullptr;       // CTA ticket of push_halo_kernel
  const void *fused_ptr = nullptr;  // vector half whose halo its PRODUCER has already pushed as exchange `seq`
  HaloRaise pending = {{nullptr, nullptr, nullptr, nullptr}, 0};   // ... and whose arrival flags the next kernel on the
                                                                   // compute stream has to raise (common.cuh)
  int *err = nullptr;               // device word: nonzero = a halo or reduction wait timed out
};

// kernel argument of the flag-based all-reduce
struct RedComm {
  RedBox *box[kMaxRanks];
  int rank, nranks;
  int *err;
  long long timeout;
};

struct Comm {
  bool active = false;   // some direction is partitioned (nranks > 1, or a forced self-partition)
  P2P p2p;
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
  int n_ext = 0;                         // boundary sites per parity (within 3 slices of a partitioned face)
  int n_int = 0;                         // the rest
  int push_ctas = 148;                   // CTAs of the halo push kernel (one per SM by default)
};

}  // namespace b200ks

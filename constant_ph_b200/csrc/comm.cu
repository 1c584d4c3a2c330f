// comm.cu -- rank group over NCCL (NVLink 5 / NVSwitch).
//
// Replaces the reference's two communication sites: MPI_Allreduce of the site sums
// (fix_constant_pH.cpp:274) -> one ncclAllReduce of [HA, HB, E_vdwl, E_coul, dU/dlambda_s...,
// HB_s-HA_s...]; and LAMMPS ghost communication (comm->reverse_comm at cpp:253 is the only
// call site; with a full neighbour list no reverse fold is needed, only the forward halo
// of {x,y,z,q}) -> grouped ncclSend/ncclRecv between spatial neighbours (halo.cu).
//
// libnccl.so.2 is bound at run time with dlopen so that the process uses the same NCCL
// instance as its host (PyTorch's bundled copy under torchrun, the system copy inside
// LAMMPS) and a single-rank build has no NCCL dependency at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "cph_internal.h"

namespace {

struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

Nccl &nccl() {
  static Nccl n;
  if (n.lib) return n;
  n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!n.lib) return n;
#define BIND(name) *(void **)(&n.name) = dlsym(n.lib, "nccl" #name)
  BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(AllReduce); BIND(Send); BIND(Recv);
  BIND(GroupStart); BIND(GroupEnd); BIND(GetErrorString); BIND(AllGather);
#undef BIND
  n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.Send && n.Recv && n.GroupStart &&
         n.GroupEnd && n.GetErrorString && n.AllGather;
  return n;
}

}  // namespace

#define CPH_NCCL(h, call)                                                                                  \
  do {                                                                                                     \
    ncclResult_t r__ = (call);                                                                             \
    if (r__ != ncclSuccess)                                                                                \
      return cph_fail(h, CPH_ERR_COMM, "%s at %s:%d: %s", #call, __FILE__, __LINE__, nccl().GetErrorString(r__)); \
  } while (0)

extern "C" int cph_comm_unique_id(char *id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!id128) return CPH_ERR_ARG;
  Nccl &n = nccl();
  if (!n.ok) return cph_fail(nullptr, CPH_ERR_COMM, "libnccl.so.2 could not be loaded: %s", dlerror());
  ncclUniqueId id;
  if (n.GetUniqueId(&id) != ncclSuccess) return cph_fail(nullptr, CPH_ERR_COMM, "ncclGetUniqueId failed");
  memcpy(id128, &id, 128);
  return CPH_OK;
}

extern "C" int cph_comm_init_nccl(cph_handle *h, int nranks, int rank, const char *id128) {
  if (nranks < 1 || rank < 0 || rank >= nranks || !id128) return cph_fail(h, CPH_ERR_ARG, "bad rank group arguments");
  if (h->nccl_comm) return cph_fail(h, CPH_ERR_STATE, "rank group already initialised");
  h->nranks = nranks;
  h->rank = rank;
  if (nranks == 1) return CPH_OK;
  Nccl &n = nccl();
  if (!n.ok) return cph_fail(h, CPH_ERR_COMM, "libnccl.so.2 could not be loaded");
  cudaSetDevice(h->device);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  CPH_NCCL(h, n.CommInitRank(&comm, nranks, id, rank));
  h->nccl_comm = (void *)comm;
  return CPH_OK;
}

void cph_comm_destroy(cph_handle *h) {
  if (h->nccl_comm) nccl().CommDestroy((ncclComm_t)h->nccl_comm);
  h->nccl_comm = nullptr;
}

int cph_comm_allreduce(cph_handle *h, double *buf, int n) {
  if (h->nranks == 1) return CPH_OK;
  ProfScope ps(h, 5);
  CPH_NCCL(h, nccl().AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
  return CPH_OK;
}

// host values in, host values out (used for the rebuild decision and overflow flags)
int cph_comm_allreduce_max_u32(cph_handle *h, unsigned int *buf, int n) {
  if (h->nranks == 1) return CPH_OK;
  if (n > 4) return cph_fail(h, CPH_ERR_ARG, "allreduce_max_u32: n > 4");
  unsigned int *d = h->d_flags.p + 12;  // scratch tail of the flags buffer
  CPH_CUDA(h, cudaMemcpyAsync(d, buf, n * sizeof(unsigned int), cudaMemcpyHostToDevice, h->stream));
  CPH_NCCL(h, nccl().AllReduce(d, d, (size_t)n, ncclUint32, ncclMax, (ncclComm_t)h->nccl_comm, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(buf, d, n * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return CPH_OK;
}

int cph_comm_allreduce_max_u32_dev(cph_handle *h, unsigned int *dbuf, int n) {
  if (h->nranks == 1) return CPH_OK;
  CPH_NCCL(h, nccl().AllReduce(dbuf, dbuf, (size_t)n, ncclUint32, ncclMax, (ncclComm_t)h->nccl_comm, h->stream));
  return CPH_OK;
}

// grouped point-to-point exchange used by the halo (halo.cu)
int cph_comm_exchange(cph_handle *h, int npeers, const int *peers, const void *const *sendbuf, const size_t *sendbytes,
                      void *const *recvbuf, const size_t *recvbytes) {
  Nccl &n = nccl();
  CPH_NCCL(h, n.GroupStart());
  for (int p = 0; p < npeers; p++) {
    if (sendbytes[p]) CPH_NCCL(h, n.Send(sendbuf[p], sendbytes[p], ncclUint8, peers[p], (ncclComm_t)h->nccl_comm, h->stream));
    if (recvbytes[p]) CPH_NCCL(h, n.Recv(recvbuf[p], recvbytes[p], ncclUint8, peers[p], (ncclComm_t)h->nccl_comm, h->stream));
  }
  CPH_NCCL(h, n.GroupEnd());
  return CPH_OK;
}

// one int per direction: how many copies each neighbour is about to send (host values in and out)
int cph_comm_exchange_counts(cph_handle *h, const int *active, const int *peer, const int *from, const int *send_count,
                             int *recv_count) {
  Nccl &n = nccl();
  int *d = (int *)(h->d_flags.p + 16);      // 27 ints out, 27 ints in (d_flags holds >= 80 words)
  CPH_CUDA(h, cudaMemcpyAsync(d, send_count, 27 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaMemsetAsync(d + 27, 0, 27 * sizeof(int), h->stream));
  CPH_NCCL(h, n.GroupStart());
  for (int k = 0; k < 27; k++) {
    if (active[k] == 2) CPH_NCCL(h, n.Send(d + k, 1, ncclInt32, peer[k], (ncclComm_t)h->nccl_comm, h->stream));
    if (from[k] >= 0) CPH_NCCL(h, n.Recv(d + 27 + k, 1, ncclInt32, from[k], (ncclComm_t)h->nccl_comm, h->stream));
  }
  CPH_NCCL(h, n.GroupEnd());
  CPH_CUDA(h, cudaMemcpyAsync(recv_count, d + 27, 27 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return CPH_OK;
}

int cph_comm_allgather(cph_handle *h, const void *sendbuf, void *recvbuf, size_t bytes_per_rank) {
  CPH_NCCL(h, nccl().AllGather(sendbuf, recvbuf, bytes_per_rank, ncclUint8, (ncclComm_t)h->nccl_comm, h->stream));
  return CPH_OK;
}

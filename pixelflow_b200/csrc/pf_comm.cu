// pf_comm.cu -- slab-to-slab halo transport over NVLink with NCCL point-to-point calls.
//
// The reference has no communication layer at all (single address space); this is
// the B200-native addition needed by the z-slab decomposition.  NCCL is loaded with
// dlopen only when nranks > 1, so the single-GPU library (and the C++ driver) has no link-time
// dependency on it.  Under torchrun the process already has torch's bundled libnccl.so.2 mapped and
// dlopen returns that one.
#include <dlfcn.h>
#include <nccl.h>

#include "pf_internal.cuh"

namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &api() {
  static NcclApi a;
  if (a.lib) return a;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !a.lib; ++i) a.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!a.lib) throw std::string("multi-GPU requested but libnccl.so.2 could not be loaded: ") + dlerror();
#define SYM(field, name)                                                        \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, name));            \
  if (!a.field) throw std::string("symbol missing in libnccl: ") + name;
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return a;
}

void ok(ncclResult_t r, const char *what) {
  if (r != ncclSuccess) throw std::string(what) + ": " + api().GetErrorString(r);
}
}  // namespace

struct PfComm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  cudaStream_t st = nullptr;
};

int pf_comm_get_unique_id(void *out128, std::string &err) {
  try {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ok(api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out128, &id, sizeof(id));
    return 0;
  } catch (const std::string &e) {
    err = e;
    return 1;
  }
}

PfComm *pf_comm_create(int rank, int nranks, const void *unique_id, cudaStream_t stream) {
  if (!unique_id) throw std::string("nranks > 1 needs pf_config.nccl_unique_id");
  PfComm *c = new PfComm;
  c->rank = rank;
  c->nranks = nranks;
  c->st = stream;
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ok(api().CommInitRank(&c->comm, nranks, id, rank), "ncclCommInitRank");
  return c;
}

void pf_comm_destroy(PfComm *c) {
  if (!c) return;
  if (c->comm) api().CommDestroy(c->comm);
  delete c;
}

// Send order (next first, then prev) and receive order (prev first, then next) are chosen so that
// with 2 ranks -- where prev == next -- the first message from the peer is the one that belongs in
// the low ghost plane.
void pf_comm_exchange(PfComm *c, const double *send_lo, const double *send_hi, double *recv_lo,
                      double *recv_hi, size_t count, int wrap, cudaStream_t on) {
  NcclApi &a = api();
  cudaStream_t st = on ? on : c->st;
  const int P = c->nranks, r = c->rank;
  const bool has_prev = wrap || r > 0, has_next = wrap || r < P - 1;
  const int prev = (r + P - 1) % P, next = (r + 1) % P;
  ok(a.GroupStart(), "ncclGroupStart");
  if (has_next) ok(a.Send(send_hi, count, ncclDouble, next, c->comm, st), "ncclSend");
  if (has_prev) ok(a.Send(send_lo, count, ncclDouble, prev, c->comm, st), "ncclSend");
  if (has_prev) ok(a.Recv(recv_lo, count, ncclDouble, prev, c->comm, st), "ncclRecv");
  if (has_next) ok(a.Recv(recv_hi, count, ncclDouble, next, c->comm, st), "ncclRecv");
  ok(a.GroupEnd(), "ncclGroupEnd");
}

void pf_comm_allreduce_max(PfComm *c, double *dev_value, size_t count) {
  ok(api().AllReduce(dev_value, dev_value, count, ncclDouble, ncclMax, c->comm, c->st), "ncclAllReduce");
}

// pf_comm.cu -- slab-to-slab halo transport over NVLink: NCCL point-to-point calls, and the CUDA-IPC
// mapping of the neighbour slabs that lets the fused SOR kernels store their boundary planes directly
// into the neighbours' ghost planes (pf_peer_open / k_slab_barrier at the end of this file).
//
// The reference has no communication layer at all (single address space); this is
// the B200-native addition needed by the z-slab decomposition.  NCCL is loaded with
// dlopen only when nranks > 1, so the single-GPU library (and the C++ driver) has no link-time
// dependency on it.  Under torchrun the process already has torch's bundled libnccl.so.2 mapped and
// dlopen returns that one.
#include <cuda.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
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
  // NCCL writes its NCCL_DEBUG lines to stdout unless told otherwise; stdout of the driver programs is the
  // reference's log, line for line, so they go to stderr (see also QuietStdout below)
  setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
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

// The bare version line of NCCL_DEBUG=VERSION ignores NCCL_DEBUG_FILE and goes to stdout on the first NCCL call of a
// process.  While that call runs, file descriptor 1 points at stderr.
struct QuietStdout {
  int saved = -1;
  QuietStdout() {
    fflush(stdout);
    saved = dup(1);
    if (saved >= 0) dup2(2, 1);
  }
  ~QuietStdout() {
    if (saved >= 0) {
      fflush(stdout);
      dup2(saved, 1);
      close(saved);
    }
  }
};

int pf_comm_get_unique_id(void *out128, std::string &err) {
  try {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    QuietStdout quiet;
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
  {
    QuietStdout quiet;
    ok(api().CommInitRank(&c->comm, nranks, id, rank), "ncclCommInitRank");
  }
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

void pf_comm_send(PfComm *c, const double *src, size_t count, int to) {
  ok(api().Send(src, count, ncclDouble, to, c->comm, c->st), "ncclSend");
}
void pf_comm_recv(PfComm *c, double *dst, size_t count, int from) {
  ok(api().Recv(dst, count, ncclDouble, from, c->comm, c->st), "ncclRecv");
}

void pf_comm_allreduce_max(PfComm *c, double *dev_value, size_t count) {
  ok(api().AllReduce(dev_value, dev_value, count, ncclDouble, ncclMax, c->comm, c->st), "ncclAllReduce");
}

void pf_comm_allreduce_sum(PfComm *c, double *dev_value, size_t count) {
  ok(api().AllReduce(dev_value, dev_value, count, ncclDouble, ncclSum, c->comm, c->st), "ncclAllReduce");
}

void pf_comm_group_begin(PfComm *) { ok(api().GroupStart(), "ncclGroupStart"); }
void pf_comm_group_end(PfComm *) { ok(api().GroupEnd(), "ncclGroupEnd"); }

// ------------------------------------------------------------------------------------------------------
// Neighbour slabs over CUDA IPC.  Each rank exports ONE allocation (its ping-pong pressure buffers preceded
// by two arrival flags); the handles travel round the ring with the NCCL exchange above, and each rank maps
// its previous and next neighbour's block.  Stores to those addresses go over NVLink.
// ------------------------------------------------------------------------------------------------------
namespace {
struct PeerMsg {             // 80 bytes = 10 doubles
  cudaIpcMemHandle_t handle; // of the underlying allocation (cudaMalloc may sub-allocate small requests)
  unsigned long long offset; // of the block inside it
  unsigned long long ok;
};
static_assert(sizeof(PeerMsg) == 80, "PeerMsg travels as 10 doubles");

typedef CUresult (*GetRangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
GetRangeFn get_range_fn() {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || !p ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<GetRangeFn>(p);
}

__global__ void slab_barrier_kernel(unsigned long long *to_prev, unsigned long long *to_next,
                                    const unsigned long long *from_prev, const unsigned long long *from_next,
                                    unsigned long long *my_seq) {
  if (threadIdx.x != 0) return;
  // the barrier number lives on the device (every rank passes the same barriers in the same order), so the launch
  // has no host-side state and replays from a CUDA graph
  const unsigned long long seq = *my_seq + 1;
  *my_seq = seq;
  // everything this stream did before (the sweep kernel's stores into the neighbours' ghost planes) is
  // complete at this point: kernels on one stream run back to back.  Publish, then wait for both sides.
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(to_prev), "l"(seq) : "memory");
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(to_next), "l"(seq) : "memory");
  unsigned long long t0, t1, a, b;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(from_prev) : "memory");
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(b) : "l"(from_next) : "memory");
    if (a >= seq && b >= seq) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 60000000000ull) __trap();   // 60 s: a neighbour died; fail the launch instead of hanging the GPU
  }
  __threadfence_system();
}
}  // namespace

void k_slab_barrier(unsigned long long *to_prev, unsigned long long *to_next, const unsigned long long *from_prev,
                    const unsigned long long *from_next, unsigned long long *my_seq, cudaStream_t st) {
  slab_barrier_kernel<<<1, 32, 0, st>>>(to_prev, to_next, from_prev, from_next, my_seq);
  pf_count_launch();
}

PfPeer *pf_peer_open(PfComm *c, void *block, std::string &why) {
  const int P = c->nranks, r = c->rank;
  const int prev = (r + P - 1) % P, next = (r + 1) % P;
  PeerMsg msg[3];   // mine, prev's, next's
  memset(msg, 0, sizeof(msg));
  std::string local_why;
  if (GetRangeFn get_range = get_range_fn()) {
    CUdeviceptr base = 0;
    size_t size = 0;
    if (get_range(&base, &size, (CUdeviceptr)block) == CUDA_SUCCESS && base) {
      const cudaError_t e = cudaIpcGetMemHandle(&msg[0].handle, (void *)base);
      if (e == cudaSuccess) {
        msg[0].offset = (unsigned long long)((CUdeviceptr)block - base);
        msg[0].ok = 1;
      } else {
        local_why = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e);
        cudaGetLastError();
      }
    } else {
      local_why = "cuMemGetAddressRange failed";
    }
  } else {
    local_why = "cuMemGetAddressRange is not available";
  }
  // ring exchange of the handles (device staging: NCCL moves device memory)
  double *dev = nullptr;
  PF_CUDA_OK(cudaMalloc(&dev, sizeof(msg) + sizeof(double)));
  PF_CUDA_OK(cudaMemcpyAsync(dev, msg, sizeof(msg), cudaMemcpyHostToDevice, c->st));
  double *d_mine = dev, *d_prev = dev + 10, *d_next = dev + 20, *d_flag = dev + 30;
  pf_comm_exchange(c, d_mine, d_mine, d_prev, d_next, 10, 1);
  PF_CUDA_OK(cudaMemcpyAsync(msg, dev, sizeof(msg), cudaMemcpyDeviceToHost, c->st));
  PF_CUDA_OK(cudaStreamSynchronize(c->st));
  PfPeer *peer = new PfPeer;
  peer->same = prev == next;
  bool good = msg[0].ok && msg[1].ok && msg[2].ok;
  if (good) {
    cudaError_t e = cudaIpcOpenMemHandle(&peer->prev_map, msg[1].handle, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) {
      peer->prev = static_cast<char *>(peer->prev_map) + msg[1].offset;
      if (peer->same) {
        peer->next = peer->prev;
      } else {
        e = cudaIpcOpenMemHandle(&peer->next_map, msg[2].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) peer->next = static_cast<char *>(peer->next_map) + msg[2].offset;
      }
    }
    if (e != cudaSuccess) {
      good = false;
      local_why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
      cudaGetLastError();
    }
  } else if (local_why.empty()) {
    local_why = "a neighbour rank could not export its block";
  }
  // all ranks take the same transport: one failure anywhere sends everybody to NCCL
  const double bad = good ? 0. : 1.;
  PF_CUDA_OK(cudaMemcpyAsync(d_flag, &bad, sizeof(double), cudaMemcpyHostToDevice, c->st));
  pf_comm_allreduce_max(c, d_flag, 1);
  double any_bad = 1.;
  PF_CUDA_OK(cudaMemcpyAsync(&any_bad, d_flag, sizeof(double), cudaMemcpyDeviceToHost, c->st));
  PF_CUDA_OK(cudaStreamSynchronize(c->st));
  cudaFree(dev);
  if (any_bad != 0.) {
    if (peer->prev_map) cudaIpcCloseMemHandle(peer->prev_map);
    if (peer->next_map) cudaIpcCloseMemHandle(peer->next_map);
    delete peer;
    why = local_why.empty() ? "another rank could not map its neighbours" : local_why;
    return nullptr;
  }
  return peer;
}

void pf_peer_close(PfComm *c, PfPeer *p) {
  if (!p) return;
  if (p->prev_map) cudaIpcCloseMemHandle(p->prev_map);
  if (p->next_map) cudaIpcCloseMemHandle(p->next_map);
  delete p;
  // nobody frees its block while a neighbour may still have it mapped
  double *d = nullptr;
  if (cudaMalloc(&d, sizeof(double)) == cudaSuccess) {
    cudaMemsetAsync(d, 0, sizeof(double), c->st);
    try { pf_comm_allreduce_max(c, d, 1); } catch (...) {}
    cudaStreamSynchronize(c->st);
    cudaFree(d);
  }
}

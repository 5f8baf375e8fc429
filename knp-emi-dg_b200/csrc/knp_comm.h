// knp_comm.h - communication layer of the cell-partitioned multi-GPU path.
//
// One context = one process = one GPU = one mesh part.  Every level of every solver
// (DG level 0, the AMG levels) numbers its OWNED unknowns first and the ghost unknowns
// after them, grouped by owning rank; a HaloPlan lists which owned entries go to which
// neighbour.  Two primitives:
//   halo(x)       ghost entries of x <- owners' values    (VecScatter of PETSc's MatMult)
//   allreduce(v)  in-place sum of a few doubles           (MPI_Allreduce of VecDot/VecNorm)
// Transport: NCCL grouped send/recv and allreduce on the context's stream (the functions
// are resolved from the already loaded libnccl.so.2 at run time, so the single-GPU path
// has no NCCL dependency); the host-emulation build delegates to callbacks (gloo in tests).
#pragma once
#include "../../include/knpemi.h"
#include "knp_common.h"
#ifndef KNP_EMU
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace knp {

struct HaloPlan {
  int64_t n_own = 0, n_ghost = 0;
  std::vector<int64_t> send_off, recv_off;   // [nneigh+1], in entries
  std::vector<int32_t> h_send_idx;           // owned entries to pack, neighbour by neighbour
  std::vector<int32_t> ghost_rank, ghost_id; // per ghost entry: owning rank and its index there
  DevBuf<int32_t> send_idx;
  DevBuf<double> sendbuf;
  int64_t nsend() const { return send_off.empty() ? 0 : send_off.back(); }
  void upload(knp_stream_t s) {
    send_idx.upload(h_send_idx, s);
    sendbuf.alloc((size_t)nsend());
  }
};

struct PackKernel {
  const int32_t* idx; const double* x; double* buf;
  KNP_HD void operator()(int64_t k) const { buf[k] = x[idx[k]]; }
};

#ifndef KNP_EMU
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool loaded = false;
  void load() {
    if (loaded) return;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) fail(std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* name) {
      void* p = dlsym(h, name);
      if (!p) fail(std::string("libnccl: missing symbol ") + name);
      return p;
    };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    Send = reinterpret_cast<decltype(Send)>(sym("ncclSend"));
    Recv = reinterpret_cast<decltype(Recv)>(sym("ncclRecv"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    loaded = true;
  }
  void check(ncclResult_t r, const char* what) const {
    if (r != ncclSuccess) fail(std::string("NCCL error in ") + what + ": " + GetErrorString(r));
  }
};
inline NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}
#endif

struct Comm {
  int rank = 0, world = 1;
  std::vector<int32_t> nbr;                  // neighbour ranks, ascending
  knp_exchange_fn xfn = nullptr;
  knp_allreduce_fn rfn = nullptr;
  void* user = nullptr;
  int64_t n_halo = 0, n_allreduce = 0;
#ifndef KNP_EMU
  ncclComm_t nccl = nullptr;
#endif
  bool active() const { return world > 1; }

  void require_transport() const {
#ifdef KNP_EMU
    if (!xfn || !rfn) fail("multi-rank context without a transport: call knp_dist_set_callbacks");
#else
    if (!nccl) fail("multi-rank context without a transport: call knp_dist_init_nccl");
#endif
  }

  // ghost entries of x (x + H.n_own ...) <- the owners' values
  void halo(knp_stream_t s, HaloPlan& H, double* x) {
    if (!active()) return;
    require_transport();
    ++n_halo;
    const int nn = (int)nbr.size();
    if (H.nsend() > 0) {
      PackKernel k{H.send_idx.p, x, H.sendbuf.p};
      parallel_for(s, H.nsend(), k);
    }
#ifdef KNP_EMU
    if (xfn(user, nn, nbr.data(), H.sendbuf.p, H.send_off.data(), x + H.n_own, H.recv_off.data()))
      fail("halo exchange callback failed");
#else
    NcclApi& N = nccl_api();
    N.check(N.GroupStart(), "ncclGroupStart");
    for (int i = 0; i < nn; ++i) {
      const int64_t ns = H.send_off[i + 1] - H.send_off[i], nr = H.recv_off[i + 1] - H.recv_off[i];
      if (ns > 0) N.check(N.Send(H.sendbuf.p + H.send_off[i], (size_t)ns, ncclDouble, nbr[i], nccl, s), "ncclSend");
      if (nr > 0) N.check(N.Recv(x + H.n_own + H.recv_off[i], (size_t)nr, ncclDouble, nbr[i], nccl, s), "ncclRecv");
    }
    N.check(N.GroupEnd(), "ncclGroupEnd");
#endif
  }

  // in-place sum over ranks of n doubles (device memory; host memory in the emulation)
  void allreduce(knp_stream_t s, double* buf, int64_t n) {
    if (!active()) return;
    require_transport();
    ++n_allreduce;
#ifdef KNP_EMU
    (void)s;
    if (rfn(user, buf, n)) fail("allreduce callback failed");
#else
    NcclApi& N = nccl_api();
    N.check(N.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, nccl, s), "ncclAllReduce");
#endif
  }
};

}  // namespace knp

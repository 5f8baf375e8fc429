// knp_comm.h - communication layer of the cell-partitioned multi-GPU path.
//
// One context = one process = one GPU = one mesh part.  Every level of every solver
// (DG level 0, the AMG levels) numbers its OWNED unknowns first and the ghost unknowns
// after them, grouped by owning rank; a HaloPlan lists which owned entries go to which
// neighbour.  Two primitives:
//   halo(x)       ghost entries of x <- owners' values    (VecScatter of PETSc's MatMult)
//   allreduce(v)  in-place sum of a few doubles           (MPI_Allreduce of VecDot/VecNorm)
// Transports:
//   * NCCL grouped send/recv and allreduce on the context's stream (the functions are
//     resolved from the already loaded libnccl.so.2 at run time, so the single-GPU path has
//     no NCCL dependency): bootstrap, setup-time exchanges, large reductions, and the fallback;
//   * peer-memory kernels over NVLink (P2P below): every rank maps its neighbours' "arena"
//     with CUDA IPC; a halo is ONE kernel that packs the owned entries straight into the
//     neighbours' staging buffers, raises a sequence-numbered flag there, waits for the
//     neighbours' flags in its own arena and unpacks - no host involvement, no separate
//     pack / send / recv / unpack launches.  The small allreduce of the Krylov scalars works
//     the same way (every rank writes its contribution to every peer, then sums in rank order:
//     bit-identical results on all ranks).
//   * the host-emulation build delegates to callbacks (gloo in tests).
#pragma once
#include "../../include/knpemi.h"
#include "knp_common.h"
#ifndef KNP_EMU
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace knp {

constexpr int P2P_MAX_NB = 16;      // neighbours per halo plan served by the peer-memory kernel
constexpr int P2P_MAX_WORLD = 16;   // ranks served by the peer-memory allreduce
constexpr int P2P_AR_MAX = 64;      // doubles per peer-memory allreduce (larger ones use NCCL)
constexpr int P2P_MAX_WS = 7;       // independent exchange channels (one per concurrently running solve)

// peer-memory state of one halo plan on one channel (filled by Comm::register_plan)
struct P2PChannel {
  bool ready = false;
  int cap = 0;                                       // vectors per exchange the staging buffers hold
  unsigned long long seq = 0;                        // exchanges done on this channel
  size_t local_flag_off[P2P_MAX_NB] = {0}, local_data_off[P2P_MAX_NB] = {0};    // in my arena
  size_t remote_flag_off[P2P_MAX_NB] = {0}, remote_data_off[P2P_MAX_NB] = {0};  // in the neighbour's arena
  size_t counter_off = 0;
};

struct HaloPlan {
  int64_t n_own = 0, n_ghost = 0;
  std::vector<int32_t> ranks;                // partner ranks of THIS plan; empty = the DG neighbours
  P2PChannel ch[P2P_MAX_WS];
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
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
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

// ---- peer-memory transport ------------------------------------------------------------
typedef unsigned long long p2p_u64;
constexpr long long P2P_SPIN_LIMIT = 40000000000LL;   // clock64 ticks (~20 s): a dead peer must not hang the GPU

struct P2PHaloArgs {
  int nn, nvec;                           // neighbours; vectors exchanged at once (a batch of linear systems)
  double* x[MAX_BATCH]; const int32_t* send_idx; int64_t n_own;
  int64_t send_off[P2P_MAX_NB + 1], recv_off[P2P_MAX_NB + 1];
  double* remote_data[P2P_MAX_NB];        // neighbour's staging buffer for my entries (this parity)
  p2p_u64* remote_flag[P2P_MAX_NB];
  const double* local_data[P2P_MAX_NB];   // my staging buffer for the neighbour's entries
  volatile p2p_u64* local_flag[P2P_MAX_NB];
  p2p_u64 seq;
  unsigned int* counter;
  int* err;
};

// One launch = one halo exchange.  Phase A: all blocks gather the owned entries and store them
// in the neighbours' memory (NVLink stores); the last block to finish raises the flags.
// Phase B: wait for each neighbour's flag, copy its entries into the ghost tail of x.
static __global__ void __launch_bounds__(256) p2p_halo_kernel(const P2PHaloArgs a) {
  if (*(volatile int*)a.err) return;   // an earlier exchange timed out: drain the queue quickly, the host reports it
  const int64_t ns = a.send_off[a.nn];
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < ns; k += (int64_t)gridDim.x * blockDim.x) {
    int i = 0;
    while (k >= a.send_off[i + 1]) ++i;
    const int64_t nsi = a.send_off[i + 1] - a.send_off[i];
    const int32_t src = a.send_idx[k];
    for (int v = 0; v < a.nvec; ++v) a.remote_data[i][v * nsi + (k - a.send_off[i])] = a.x[v][src];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < a.nn) *a.remote_flag[threadIdx.x] = a.seq;
    if (threadIdx.x == 0) *a.counter = 0;
  }
  for (int i = 0; i < a.nn; ++i) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      while (*a.local_flag[i] < a.seq)
        if (clock64() - t0 > P2P_SPIN_LIMIT) { *a.err = 1; break; }
    }
    __syncthreads();
    const int64_t nr = a.recv_off[i + 1] - a.recv_off[i];
    for (int v = 0; v < a.nvec; ++v) {
      double* dst = a.x[v] + a.n_own + a.recv_off[i];
      const double* src = a.local_data[i] + v * nr;
      for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nr; k += (int64_t)gridDim.x * blockDim.x)
        dst[k] = __ldcv(src + k);
    }
  }
}

struct P2PArArgs {
  int world, rank, n;
  double* buf;
  double* peer_data[P2P_MAX_WORLD];   // rank r's slot array [world][P2P_AR_MAX] (this parity)
  p2p_u64* peer_flag[P2P_MAX_WORLD];  // rank r's flag array [world] (this parity)
  const double* my_data; volatile p2p_u64* my_flag;
  p2p_u64 seq;
  int* err;
};

// In-place sum of n <= P2P_AR_MAX doubles over all ranks: every rank stores its contribution
// in slot [rank] of every peer, raises flag [rank] there, waits for all flags in its own
// arena and sums the slots in rank order (same bits on every rank).
static __global__ void __launch_bounds__(128) p2p_allreduce_kernel(const P2PArArgs a) {
  if (*(volatile int*)a.err) return;
  const int t = threadIdx.x;
  if (t < a.n) {
    const double v = a.buf[t];
    for (int r = 0; r < a.world; ++r) a.peer_data[r][a.rank * P2P_AR_MAX + t] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (t < a.world) {
    __threadfence_system();
    a.peer_flag[t][a.rank] = a.seq;
    const long long t0 = clock64();
    while (a.my_flag[t] < a.seq)
      if (clock64() - t0 > P2P_SPIN_LIMIT) { *a.err = 1; break; }
  }
  __syncthreads();
  if (t < a.n) {
    double sum = 0.0;
    for (int r = 0; r < a.world; ++r) sum += __ldcv(a.my_data + r * P2P_AR_MAX + t);
    a.buf[t] = sum;
  }
}

struct P2P {
  bool on = false;
  char* arena = nullptr;
  size_t bytes = 0, bump = 0;
  char* peer[P2P_MAX_WORLD] = {nullptr};
  p2p_u64 ar_seq[P2P_MAX_WS] = {0};
  // fixed header (same offsets on every rank): error word, allreduce flags and slots per channel
  static constexpr size_t ERR_OFF = 0;
  static constexpr size_t AR_FLAG_OFF = 256;                                   // [WS][2][P2P_MAX_WORLD] u64
  static constexpr size_t AR_FLAG_WS = 2 * P2P_MAX_WORLD * sizeof(p2p_u64);
  static constexpr size_t AR_DATA_OFF = AR_FLAG_OFF + P2P_MAX_WS * AR_FLAG_WS; // [WS][2][P2P_MAX_WORLD][P2P_AR_MAX] f64
  static constexpr size_t AR_DATA_WS = 2 * P2P_MAX_WORLD * P2P_AR_MAX * sizeof(double);
  static constexpr size_t HEADER = AR_DATA_OFF + P2P_MAX_WS * AR_DATA_WS;
  size_t alloc(size_t nbytes) {
    const size_t o = (bump + 255) & ~(size_t)255;
    if (o + nbytes > bytes) fail("peer-memory arena exhausted");
    bump = o + (nbytes ? nbytes : 8);
    return o;
  }
};
#endif

struct Comm {
  int rank = 0, world = 1;
  std::vector<int32_t> nbr;                  // neighbour ranks, ascending
  knp_exchange_fn xfn = nullptr;
  knp_allreduce_fn rfn = nullptr;
  void* user = nullptr;
  std::atomic<int64_t> n_halo{0}, n_allreduce{0};
  int batch_cap = 1;                         // vectors per exchange the peer-memory staging is sized for
#ifndef KNP_EMU
  ncclComm_t nccl = nullptr;
  P2P p2p;
  std::atomic<int64_t> n_p2p{0};             // exchanges served by the peer-memory kernels
#endif
  const std::vector<int32_t>& partners(const HaloPlan& H) const { return H.ranks.empty() ? nbr : H.ranks; }
  bool active() const { return world > 1; }

#ifndef KNP_EMU
  void nccl_allreduce(knp_stream_t s, double* buf, int64_t n) {
    NcclApi& N = nccl_api();
    N.check(N.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, nccl, s), "ncclAllReduce");
  }

  // Map every rank's arena (collective; call once after the NCCL communicator exists).  Any
  // failure on any rank leaves the NCCL transport in charge everywhere.
  void setup_p2p(knp_stream_t s) {
    const char* env = getenv("KNP_P2P");
    if (env && env[0] == '0') return;
    if (world > P2P_MAX_WORLD || p2p.on) return;
    P2P& P = p2p;
    int ok = 1;
    P.bytes = (size_t)160 << 20;
    if (cudaMalloc((void**)&P.arena, P.bytes) != cudaSuccess) { cudaGetLastError(); P.arena = nullptr; ok = 0; }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok) {
      KNP_CUDA(cudaMemset(P.arena, 0, P.bytes));
      KNP_CUDA(cudaDeviceSynchronize());
      if (cudaIpcGetMemHandle(&mine, P.arena) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    // all-gather of the 64-byte handles through the NCCL allreduce: one double per byte
    // (exact under summation with zeros), plus one failure counter
    const size_t HB = sizeof(cudaIpcMemHandle_t);
    std::vector<double> hv((size_t)world * HB + 1, 0.0);
    for (size_t b = 0; b < HB; ++b) hv[(size_t)rank * HB + b] = (double)((const unsigned char*)&mine)[b];
    hv.back() = ok ? 0.0 : 1.0;
    DevBuf<double> tmp;
    tmp.upload(hv, s);
    nccl_allreduce(s, tmp.p, (int64_t)hv.size());
    hv = tmp.download(s);
    if (hv.back() == 0.0) {
      for (int r = 0; r < world; ++r) {
        if (r == rank) { P.peer[r] = P.arena; continue; }
        cudaIpcMemHandle_t h;
        for (size_t b = 0; b < HB; ++b) ((unsigned char*)&h)[b] = (unsigned char)hv[(size_t)r * HB + b];
        if (cudaIpcOpenMemHandle((void**)&P.peer[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError(); P.peer[r] = nullptr; ok = 0;
        }
      }
    } else ok = 0;
    double bad = ok ? 0.0 : 1.0;
    h2d(tmp.p, &bad, sizeof bad, s);
    nccl_allreduce(s, tmp.p, 1);
    d2h(&bad, tmp.p, sizeof bad, s);
    if (bad != 0.0) { close_p2p(); return; }
    P.bump = P2P::HEADER;
    P.on = true;
  }

  void close_p2p() {
    P2P& P = p2p;
    for (int r = 0; r < P2P_MAX_WORLD; ++r) {
      if (P.peer[r] && r != rank) cudaIpcCloseMemHandle(P.peer[r]);
      P.peer[r] = nullptr;
    }
    if (P.arena) cudaFree(P.arena);
    P.arena = nullptr; P.on = false;
  }

  // staging buffers and flags of one halo plan on one channel; the partners learn where to
  // write through one small NCCL exchange (collective among the plan's partners; main thread)
  void register_plan(knp_stream_t s, HaloPlan& H, int w, int cap = 0) {
    P2P& P = p2p;
    P2PChannel& C = H.ch[w];
    if (cap < batch_cap) cap = batch_cap;
    if (cap < 1) cap = 1;
    C.cap = cap;
    const std::vector<int32_t>& R = partners(H);
    const int nn = (int)R.size();
    std::vector<double> mine(2 * (size_t)nn + 2, 0.0), theirs(2 * (size_t)nn + 2, 0.0);
    for (int i = 0; i < nn; ++i) {
      const int64_t nr = H.recv_off[i + 1] - H.recv_off[i];
      C.local_flag_off[i] = P.alloc(2 * sizeof(p2p_u64));
      C.local_data_off[i] = P.alloc(2 * (size_t)nr * cap * sizeof(double));
      mine[2 * i] = (double)C.local_flag_off[i]; mine[2 * i + 1] = (double)C.local_data_off[i];
    }
    C.counter_off = P.alloc(sizeof(unsigned int));
    DevBuf<double> sb, rb;
    sb.upload(mine, s); rb.upload(theirs, s);
    NcclApi& N = nccl_api();
    N.check(N.GroupStart(), "ncclGroupStart");
    for (int i = 0; i < nn; ++i) {
      if (R[i] == rank) continue;
      N.check(N.Send(sb.p + 2 * i, 2, ncclDouble, R[i], nccl, s), "ncclSend");
      N.check(N.Recv(rb.p + 2 * i, 2, ncclDouble, R[i], nccl, s), "ncclRecv");
    }
    N.check(N.GroupEnd(), "ncclGroupEnd");
    theirs = rb.download(s);
    for (int i = 0; i < nn; ++i) {
      if (R[i] == rank) { C.remote_flag_off[i] = C.local_flag_off[i]; C.remote_data_off[i] = C.local_data_off[i]; }
      else { C.remote_flag_off[i] = (size_t)theirs[2 * i]; C.remote_data_off[i] = (size_t)theirs[2 * i + 1]; }
    }
    C.seq = 0;
    C.ready = true;
  }
  // make channels 0..nws-1 of a plan usable from worker threads (no NCCL on first use there)
  void prepare_plan(knp_stream_t s, HaloPlan& H, int nws) {
    if (!p2p.on || (int)partners(H).size() > P2P_MAX_NB) return;
    for (int w = 0; w < nws && w < P2P_MAX_WS; ++w)
      if (!H.ch[w].ready) register_plan(s, H, w);
  }
  static bool& in_worker() { static thread_local bool v = false; return v; }

  void check_p2p(knp_stream_t s) {
    if (!p2p.on) return;
    int e = 0;
    d2h(&e, p2p.arena + P2P::ERR_OFF, sizeof e, s);
    if (e) fail("peer-memory exchange timed out (a neighbouring rank stopped responding)");
  }
#endif

  void require_transport() const {
#ifdef KNP_EMU
    if (!xfn || !rfn) fail("multi-rank context without a transport: call knp_dist_set_callbacks");
#else
    if (!nccl) fail("multi-rank context without a transport: call knp_dist_init_nccl");
#endif
  }

  // ghost entries of x (x + H.n_own ...) <- the owners' values
  void halo(knp_stream_t s, HaloPlan& H, double* x, int w = 0) { halo_batch(s, H, 1, &x, w); }

  // the same for `nvec` vectors at once (the solved ions' systems share mesh, partition and halo
  // plan): ONE peer-memory kernel / one flag round trip for all of them
  void halo_batch(knp_stream_t s, HaloPlan& H, int nvec, double* const* xs, int w = 0) {
    if (!active()) return;
    require_transport();
    if (nvec < 1 || nvec > MAX_BATCH) fail("halo_batch: bad vector count");
    ++n_halo;
    const std::vector<int32_t>& nbr = partners(H);   // (shadows the DG neighbour list on purpose)
    const int nn = (int)nbr.size();
#ifndef KNP_EMU
    if (p2p.on && nn <= P2P_MAX_NB && w < P2P_MAX_WS) {
      if (nn == 0) return;
      if (!H.ch[w].ready || H.ch[w].cap < nvec) {
        // (first use, or more vectors than the staging holds: collective among the partners, every
        // rank gets here in the same call)
        if (in_worker()) fail("halo plan was not prepared for concurrent use");
        register_plan(s, H, w, nvec);
      }
      P2P& P = p2p;
      P2PChannel& C = H.ch[w];
      P2PHaloArgs a;
      a.nn = nn; a.nvec = nvec; a.send_idx = H.send_idx.p; a.n_own = H.n_own;
      for (int v = 0; v < nvec; ++v) a.x[v] = xs[v];
      const p2p_u64 seq = ++C.seq;
      const size_t par = (size_t)(seq & 1);
      int64_t most = 1;
      for (int i = 0; i <= nn; ++i) { a.send_off[i] = H.send_off[i]; a.recv_off[i] = H.recv_off[i]; }
      for (int i = 0; i < nn; ++i) {
        const int64_t ns = H.send_off[i + 1] - H.send_off[i], nr = H.recv_off[i + 1] - H.recv_off[i];
        char* pb = P.peer[nbr[i]];
        a.remote_data[i] = reinterpret_cast<double*>(pb + C.remote_data_off[i]) + par * (size_t)ns * C.cap;
        a.remote_flag[i] = reinterpret_cast<p2p_u64*>(pb + C.remote_flag_off[i]) + par;
        a.local_data[i] = reinterpret_cast<const double*>(P.arena + C.local_data_off[i]) + par * (size_t)nr * C.cap;
        a.local_flag[i] = reinterpret_cast<volatile p2p_u64*>(P.arena + C.local_flag_off[i]) + par;
        most = ns > most ? ns : most; most = nr > most ? nr : most;
      }
      a.seq = seq;
      a.counter = reinterpret_cast<unsigned int*>(P.arena + C.counter_off);
      a.err = reinterpret_cast<int*>(P.arena + P2P::ERR_OFF);
      int64_t grid = (most + 255) / 256;   // (one block per 256 entries; every thread moves nvec values)
      if (grid < 1) grid = 1;
      if (grid > 32) grid = 32;
      ++launch_counter(); ++n_p2p;
      p2p_halo_kernel<<<(unsigned)grid, 256, 0, s>>>(a);
      KNP_CUDA(cudaGetLastError());
      return;
    }
#endif
#ifndef KNP_EMU
    if (in_worker()) fail("NCCL exchange requested from a worker thread");
#endif
    // fallback transports: one vector after the other (the send buffer is reused in stream order)
    for (int v = 0; v < nvec; ++v) {
      double* x = xs[v];
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
  }

  // in-place all-gather: rank r's `count` doubles sit at buf + r*count on entry (own segment)
  // and on every rank on return
  void allgather(knp_stream_t s, double* buf, int64_t count) {
    if (!active()) return;
    require_transport();
    ++n_allreduce;
#ifdef KNP_EMU
    (void)s;
    for (int r = 0; r < world; ++r)
      if (r != rank) memset(buf + (int64_t)r * count, 0, (size_t)count * sizeof(double));
    if (rfn(user, buf, (int64_t)world * count)) fail("allgather callback failed");
#else
    NcclApi& N = nccl_api();
    N.check(N.AllGather(buf + (int64_t)rank * count, buf, (size_t)count, ncclDouble, nccl, s), "ncclAllGather");
#endif
  }

  // in-place sum over ranks of n doubles (device memory; host memory in the emulation)
  void allreduce(knp_stream_t s, double* buf, int64_t n, int w = 0) {
    if (!active()) return;
    require_transport();
    ++n_allreduce;
#ifdef KNP_EMU
    (void)s;
    if (rfn(user, buf, n)) fail("allreduce callback failed");
#else
    if (p2p.on && n <= P2P_AR_MAX && w < P2P_MAX_WS) {
      P2P& P = p2p;
      P2PArArgs a;
      a.world = world; a.rank = rank; a.n = (int)n; a.buf = buf;
      const p2p_u64 seq = ++P.ar_seq[w];
      const size_t par = (size_t)(seq & 1);
      const size_t doff = P2P::AR_DATA_OFF + (size_t)w * P2P::AR_DATA_WS, foff = P2P::AR_FLAG_OFF + (size_t)w * P2P::AR_FLAG_WS;
      for (int r = 0; r < world; ++r) {
        a.peer_data[r] = reinterpret_cast<double*>(P.peer[r] + doff) + par * P2P_MAX_WORLD * P2P_AR_MAX;
        a.peer_flag[r] = reinterpret_cast<p2p_u64*>(P.peer[r] + foff) + par * P2P_MAX_WORLD;
      }
      a.my_data = reinterpret_cast<const double*>(P.arena + doff) + par * P2P_MAX_WORLD * P2P_AR_MAX;
      a.my_flag = reinterpret_cast<volatile p2p_u64*>(P.arena + foff) + par * P2P_MAX_WORLD;
      a.seq = seq;
      a.err = reinterpret_cast<int*>(P.arena + P2P::ERR_OFF);
      ++launch_counter(); ++n_p2p;
      p2p_allreduce_kernel<<<1, 128, 0, s>>>(a);
      KNP_CUDA(cudaGetLastError());
      return;
    }
    if (in_worker()) fail("NCCL allreduce requested from a worker thread");
    nccl_allreduce(s, buf, n);
#endif
  }
};

}  // namespace knp

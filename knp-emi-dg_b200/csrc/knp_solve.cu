// knp_solve.cu - Krylov solvers and the AMG preconditioner.
//
//   knp_solve_emi : preconditioned CG on A_emi with the preconditioner built from
//                   B_emi = A_emi + kappa/Lp^2 mass (solver.py:377-395, 425-444, 502-509)
//   knp_solve_knp : left-preconditioned restarted GMRES, one system per solved ion
//                   (the ions are uncoupled in the form, solver.py:550-594, 684-701, 767-771)
// Convergence test as PETSc's default: preconditioned residual norm
// ||M^-1 r|| <= max(rtol * ||M^-1 b||, atol), nonzero initial guess.
#include "../../include/knpemi.h"
#include "knp_ctx.h"
#include <thread>
#include <type_traits>

using namespace knp;

namespace knp {
int set_error(const std::string& s);
BellMat bell_of(knp_ctx* c, int which);
}

#define KNP_TRY try {
#define KNP_CATCH                                             \
  }                                                           \
  catch (const std::exception& e) { return knp::set_error(e.what()); } \
  catch (...) { return knp::set_error("unknown error"); }     \
  return 0;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- per-solve workspace -----------------------------------------------------------------
// Every helper below issues on the stream of the CURRENT workspace and uses its scratch
// buffers: the context's main workspace by default, a per-ion workspace inside the worker
// threads that solve the independent KNP systems concurrently.
static thread_local KrylovWs* tl_ws = nullptr;
static KrylovWs& ws(knp_ctx* c) { return tl_ws ? *tl_ws : c->kr0; }
static knp_stream_t cs(knp_ctx* c) { return ws(c).stream; }

// ---- small helpers ---------------------------------------------------------------------
// Multi-GPU: every kernel below runs over the OWNED rows (the first n_own entries of a
// vector); the operators that read a vector through the matrix first refresh its ghost
// entries from their owners (knp_comm.h).
static void halo0(knp_ctx* c, const double* x) {
  if (c->comm.active()) c->comm.halo(cs(c), c->halo0, const_cast<double*>(x), ws(c).id);
}
template <int ND, typename T>
static void bell_spmv_nd(knp_ctx* c, const BellMatT<T>& A, const double* x, const double* b, double* y, int mode) {
  BellSpmvKernel<ND, T> k{A, x, b, y, mode};
  parallel_for(cs(c), c->n_own, k, 256);
}
template <typename T>
static void bell_spmv(knp_ctx* c, const BellMatT<T>& A, const double* x, const double* b, double* y, int mode) {
  halo0(c, x);
  if (c->nd == 3) bell_spmv_nd<3, T>(c, A, x, b, y, mode); else bell_spmv_nd<4, T>(c, A, x, b, y, mode);
}
template <typename T>
static void block_apply(knp_ctx* c, const T* dinv, const double* r, double* out, double w, int mode) {
  if (c->nd == 3) { BlockDiagApplyKernel<3, T> k{dinv, r, out, w, mode}; parallel_for(cs(c), c->n_own, k); }
  else { BlockDiagApplyKernel<4, T> k{dinv, r, out, w, mode}; parallel_for(cs(c), c->n_own, k); }
}
template <typename T>
static void bell_jacobi(knp_ctx* c, const BellMatT<T>& A, const T* dinv, const double* b,
                        const double* xin, double* xout, double w) {
  halo0(c, xin);
  if (c->nd == 3) { BellJacobiKernel<3, T> k{A, dinv, b, xin, xout, w}; parallel_for(cs(c), c->n_own, k, 192); }
  else { BellJacobiKernel<4, T> k{A, dinv, b, xin, xout, w}; parallel_for(cs(c), c->n_own, k, 256); }
}
// second Chebyshev step: xout = xin + beta (xin - xprev) + w Dinv (b - A xin); xprev may be nullptr (zero)
template <typename T>
static void bell_jacobi_mom(knp_ctx* c, const BellMatT<T>& A, const T* dinv, const double* b, const double* xin,
                            const double* xprev, double* xout, double w, double beta) {
  halo0(c, xin);
  if (c->nd == 3) {
    BellJacobiKernel<3, T, true> k{A, dinv, b, xin, xout, w, nullptr, nullptr, xprev, beta};
    parallel_for(cs(c), c->n_own, k, 192);
  } else {
    BellJacobiKernel<4, T, true> k{A, dinv, b, xin, xout, w, nullptr, nullptr, xprev, beta};
    parallel_for(cs(c), c->n_own, k, 256);
  }
}
// post-smoothing sweep fused with the prolongation: out = x' + w Dinv (b - A x'), x' = xin + P xc
// (xin may be nullptr).  The ghost entries of xin (if any) and of xc must be valid.
static void bell_jacobi_prolong(knp_ctx* c, const BellMat& A, const double* dinv, const double* b,
                                const double* xin, const int32_t* agg, const double* xc, double* xout, double w) {
  if (c->nd == 3) { BellJacobiKernel<3> k{A, dinv, b, xin, xout, w, agg, xc}; parallel_for(cs(c), c->n_own, k, 192); }
  else { BellJacobiKernel<4> k{A, dinv, b, xin, xout, w, agg, xc}; parallel_for(cs(c), c->n_own, k, 256); }
}
static void block_inverse(knp_ctx* c, const double* blocks, double* inv) {
#ifdef KNP_EMU
  if (c->nd == 3) { BlockInverseKernel<3> k{blocks, inv}; parallel_for(cs(c), c->nc_own, k, 128); }
  else { BlockInverseKernel<4> k{blocks, inv}; parallel_for(cs(c), c->nc_own, k, 128); }
#else
  const unsigned grid = (unsigned)((c->nc_own + 127) / 128);
  ++launch_counter();
  if (c->nd == 3) block_inverse_kernel<3><<<grid, 128, 0, cs(c)>>>(c->nc_own, blocks, inv);
  else block_inverse_kernel<4><<<grid, 128, 0, cs(c)>>>(c->nc_own, blocks, inv);
  KNP_CUDA(cudaGetLastError());
#endif
}

// host-visible dot products over all ranks (one sync each)
static void dots_host(knp_ctx* c, int k, const double* V, const double* w, double* out_host) {
  multi_dot_device(cs(c), c->n_own, c->n, k, V, w, ws(c).partial.p, ws(c).scal.p);
  c->comm.allreduce(cs(c), ws(c).scal.p, k, ws(c).id);
  d2h(out_host, ws(c).scal.p, k * sizeof(double), cs(c));
}
// sum over ranks of a few host numbers (setup-time decisions must agree on every rank)
static void global_sum(knp_ctx* c, double* v, int k) {
  if (!c->comm.active()) return;
  double* dev = ws(c).scal.p + 768;
  h2d(dev, v, k * sizeof(double), cs(c));
  c->comm.allreduce(cs(c), dev, k);
  d2h(v, dev, k * sizeof(double), cs(c));
}
static double dot_host(knp_ctx* c, const double* x, const double* y) {
  double v;
  dots_host(c, 1, x, y, &v);
  return v;
}

// ---------------------------------------------------------------------------------
// AMG setup (host plan) -----------------------------------------------------------
// ---------------------------------------------------------------------------------
// level-0 scalar CSR view of a block-ELL matrix with the storage position of every entry
// (slot 0 = diagonal blocks; for the EMI buffer these are B's, see knp_ctx.h).
static HostCsr level0_csr(knp_ctx* c) {
  const int nd = c->nd;
  const int64_t nc = c->nc, bs = c->bs(), ss = c->slot_stride();
  HostCsr A;
  A.n = c->n_own;
  A.ptr.assign(A.n + 1, 0);
  A.col.reserve((size_t)c->nnz_export); A.pos.reserve((size_t)c->nnz_export);
  for (int64_t cell = 0; cell < c->nc_own; ++cell)
    for (int i = 0; i < nd; ++i) {
      const int64_t dbase = cell * bs + i * nd;
      for (int j = 0; j < nd; ++j) { A.col.push_back((int32_t)(cell * nd + j)); A.pos.push_back((int32_t)(dbase + j)); }
      for (int f = 0; f < nd; ++f) {
        const int32_t c2 = c->h_nbr[(size_t)f * nc + cell];
        if (c2 < 0) continue;
        const int64_t obase = (int64_t)(1 + f) * ss + cell * bs + i * nd;
        for (int j = 0; j < nd; ++j) { A.col.push_back(c2 * nd + j); A.pos.push_back((int32_t)(obase + j)); }
      }
      A.ptr[cell * nd + i + 1] = (int32_t)A.col.size();
    }
  return A;
}

// DG dof -> region-wise continuous vertex: union-find over matching dofs of tag-0 facets
// (owned cells only: the gluing stops at partition boundaries, like at membranes, so that
// every coarse unknown lives on one rank)
static int64_t vertex_injection(knp_ctx* c, std::vector<int32_t>& agg) {
  const int nd = c->nd;
  const int64_t nc = c->nc, n = c->n_own;
  std::vector<int32_t> parent(n);
  std::iota(parent.begin(), parent.end(), 0);
  auto find = [&](int32_t x) {
    while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
    return x;
  };
  for (int64_t cell = 0; cell < c->nc_own; ++cell)
    for (int f = 0; f < nd; ++f) {
      const int w = c->h_finfo[(size_t)f * nc + cell];
      if (fi_kind(w) != FK_SIP) continue;
      const int32_t c2 = c->h_nbr[(size_t)f * nc + cell];
      if (c2 < cell || c2 >= c->nc_own) continue;
      for (int a = 0; a < nd; ++a) {
        if (a == f) continue;
        const int32_t x = find((int32_t)(cell * nd + a)), y = find(c2 * nd + fi_perm(w, a));
        if (x != y) parent[x > y ? x : y] = x > y ? y : x;
      }
    }
  agg.assign(n, -1);
  int32_t na = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t r = find((int32_t)i);
    if (agg[r] < 0) agg[r] = na++;
    agg[i] = agg[r];
  }
  return na;
}

// Ghost side of a coarsening step.  `agg_own` maps the owned fine unknowns to owned coarse
// unknowns 0..na_own-1.  Every rank sends the coarse index of its owned fine unknowns through
// the fine level's halo; the distinct (owner, index) pairs received become the ghost unknowns
// of the coarse level (numbered after the owned ones, grouped by owner, ascending index), and
// the image of each fine send list is the coarse send list - the owner and the receiver sort
// by the same key, so the two sides agree without a handshake.  Returns agg for ALL local
// fine unknowns and fills the coarse level's halo plan Hc.
static std::vector<int32_t> extend_aggregates(knp_ctx* c, HaloPlan& Hf, const std::vector<int32_t>& agg_own,
                                              int64_t na_own, HaloPlan& Hc) {
  const int64_t nf_own = Hf.n_own, nf_ghost = Hf.n_ghost;
  const int nn = (int)c->comm.nbr.size();
  std::vector<int32_t> agg((size_t)(nf_own + nf_ghost));
  std::copy(agg_own.begin(), agg_own.begin() + nf_own, agg.begin());
  Hc.n_own = na_own; Hc.n_ghost = 0;
  Hc.send_off.assign(nn + 1, 0); Hc.recv_off.assign(nn + 1, 0);
  Hc.h_send_idx.clear(); Hc.ghost_rank.clear(); Hc.ghost_id.clear();
  if (!c->comm.active()) return agg;
  std::vector<double> ids((size_t)(nf_own + nf_ghost), -1.0);
  for (int64_t i = 0; i < nf_own; ++i) ids[i] = (double)agg_own[i];
  DevBuf<double> tmp;
  tmp.upload(ids, cs(c));
  c->comm.halo(cs(c), Hf, tmp.p);
  ids = tmp.download(cs(c));
  int64_t ng = 0;
  std::vector<int32_t> uniq;
  for (int i = 0; i < nn; ++i) {
    uniq.clear();
    for (int64_t j = Hf.recv_off[i]; j < Hf.recv_off[i + 1]; ++j) uniq.push_back((int32_t)ids[nf_own + j]);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    if (!uniq.empty() && uniq.front() < 0) fail("extend_aggregates: a ghost unknown was not sent by its owner");
    for (int64_t j = Hf.recv_off[i]; j < Hf.recv_off[i + 1]; ++j) {
      const int32_t id = (int32_t)ids[nf_own + j];
      agg[nf_own + j] = (int32_t)(na_own + ng + (std::lower_bound(uniq.begin(), uniq.end(), id) - uniq.begin()));
    }
    for (int32_t id : uniq) { Hc.ghost_rank.push_back(c->comm.nbr[i]); Hc.ghost_id.push_back(id); }
    ng += (int64_t)uniq.size();
    Hc.recv_off[i + 1] = ng;
    uniq.clear();
    for (int64_t k = Hf.send_off[i]; k < Hf.send_off[i + 1]; ++k) uniq.push_back(agg_own[Hf.h_send_idx[k]]);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    Hc.h_send_idx.insert(Hc.h_send_idx.end(), uniq.begin(), uniq.end());
    Hc.send_off[i + 1] = (int64_t)Hc.h_send_idx.size();
  }
  Hc.n_ghost = ng;
  Hc.upload(cs(c));
  return agg;
}

static void upload_level(knp_ctx* c, AmgLevelPlan& L, const GalerkinPlan& G, const HostTransfer& T) {
  knp_stream_t s = cs(c);
  L.n = G.coarse.n; L.nnz = (int64_t)G.coarse.col.size();
  L.nloc = L.n + L.halo.n_ghost;
  L.ptr.upload(G.coarse.ptr, s); L.col.upload(G.coarse.col, s);
  L.gptr.upload(G.gptr, s); L.gidx.upload(G.gidx, s);
  L.g_unit = G.gw.empty();
  if (!L.g_unit) L.gw.upload(G.gw, s);
  L.pptr.upload(T.pptr, s); L.pidx.upload(T.pidx, s);
  L.rptr.upload(T.rptr, s); L.ridx.upload(T.ridx, s);
  L.t_unit = T.unit;
  if (!T.unit) { L.pw.upload(T.pw, s); L.rw.upload(T.rw, s); }
}

static void alloc_values(knp_ctx* c, AmgValues& V) {
  const size_t nl = c->amg.lev.size();
  V.val.resize(nl); V.dinv.resize(nl);
  for (size_t l = 0; l < nl; ++l) { V.val[l].alloc(c->amg.lev[l].nnz); V.dinv[l].alloc(c->amg.lev[l].nloc); }
  V.dense.alloc((size_t)c->amg.m_dense * c->amg.m_dense);
  V.binv.alloc((size_t)c->slot_stride());
  V.solves = 0; V.omega = 0.0; V.age = 0;   // fresh buffers: the next solve must refresh the values
  V.vec.clear(); V.vec.resize(nl);
  for (size_t l = 0; l < nl; ++l) {
    const size_t m = (size_t)c->amg.lev[l].nloc;
    V.vec[l].b.alloc(m); V.vec[l].x.alloc(m); V.vec[l].r.alloc(m); V.vec[l].t.alloc(m);
  }
  V.x0.alloc(c->n); V.r0.alloc(c->n); V.t0.alloc(c->n);
  V.colbuf.alloc(c->amg.m_dense);
  V.dense_b.alloc(c->amg.m_dense); V.dense_x.alloc(c->amg.m_dense);
  if (c->amg.rep_from != (size_t)-1) V.rep_b.alloc((size_t)c->comm.world * c->amg.rep_bstride);
}

static int64_t double_coarsening_rows() {
  const char* e = getenv("KNP_AMG_DOUBLE");
  return e ? atoll(e) : 0;
}

#ifndef KNP_EMU
// CUDA loads kernels lazily (CUDA_MODULE_LOADING=LAZY is the default since 12.2): the first
// launch of a kernel may have to wait until the kernels that are running have finished.  A
// worker thread that launches a not-yet-loaded kernel while the OTHER worker's exchange kernel
// is spinning on a neighbour rank - which may be stuck the same way - is a cross-rank deadlock.
// Every kernel the concurrent solves can launch is therefore loaded up front
// (cudaFuncGetAttributes loads the function).
template <class K>
static void touch_kernel(K kernel) {
  cudaFuncAttributes a;
  if (cudaFuncGetAttributes(&a, (const void*)kernel) != cudaSuccess) cudaGetLastError();   // best effort
}
template <int ND>
static void preload_solver_kernels_nd() {
  touch_kernel(pf_kernel<BellSpmvKernel<ND>>);
  touch_kernel(pf_kernel<BlockDiagApplyKernel<ND>>);
  touch_kernel(pf_kernel<BellJacobiKernel<ND>>);
  touch_kernel(pf_kernel<BellSpmvKernel<ND, float>>);
  touch_kernel(pf_kernel<BlockDiagApplyKernel<ND, float>>);
  touch_kernel(pf_kernel<BellJacobiKernel<ND, float>>);
  touch_kernel(pf_kernel<BellJacobiKernel<ND, double, true>>);
  touch_kernel(pf_kernel<BellJacobiKernel<ND, float, true>>);
  touch_kernel(block_inverse_kernel<ND>);
}
static void preload_solver_kernels(knp_ctx* c) {
  if (c->nd == 3) preload_solver_kernels_nd<3>(); else preload_solver_kernels_nd<4>();
  touch_kernel(pf_kernel<TransferKernel>);
  touch_kernel(pf_kernel<GalerkinKernel>);
  touch_kernel(pf_kernel<CsrL1DiagKernel>);
  touch_kernel(pf_kernel<CsrToDenseKernel>);
  touch_kernel(pf_kernel<CsrSpmvKernel>);
  touch_kernel(pf_kernel<CsrJacobiKernel>);
  touch_kernel(pf_kernel<DiagScaleKernel>);
  touch_kernel(pf_kernel<DenseMatvecKernel>);
  touch_kernel(pf_kernel<ScatterOffsetKernel>);
  touch_kernel(pf_kernel<GatherMapKernel>);
  touch_kernel(pf_kernel<ScaleKernel>);
  touch_kernel(pf_kernel<CombineKernel>);
  touch_kernel(pf_kernel<GsNormalizeKernel>);
  touch_kernel(pf_kernel<AddConstKernel>);
  touch_kernel(pf_kernel<PackKernel>);
  touch_kernel(subwarp_kernel<8, CoarseResidualKernel>);
  touch_kernel(subwarp_kernel<8, TransferRowsKernel>);
  touch_kernel(subwarp_kernel<8, CoarseUpKernel>);
  touch_kernel(coarse_tail_kernel);
  touch_kernel(dense_inverse_smem_kernel);
  touch_kernel(dense_inverse_kernel);
  touch_kernel(multi_dot_final);
  touch_kernel(multi_dot_partial<1>); touch_kernel(multi_dot_partial<2>); touch_kernel(multi_dot_partial<3>);
  touch_kernel(multi_dot_partial<4>); touch_kernel(multi_dot_partial<5>); touch_kernel(multi_dot_partial<6>);
  touch_kernel(multi_dot_partial<7>); touch_kernel(multi_dot_partial<8>);
  touch_kernel(pair_dot_partial<1>); touch_kernel(pair_dot_partial<2>);
  touch_kernel(pair_dot_partial<3>); touch_kernel(pair_dot_partial<4>);
  touch_kernel(p2p_halo_kernel);
  touch_kernel(p2p_allreduce_kernel);
}
#endif

// rows (over all ranks) below which the rest of the hierarchy is replicated; 0 disables
static int64_t replicate_threshold() {
  const char* e = getenv("KNP_AMG_REPLICATE");
  return e ? atoll(e) : 131072;
}

// gather `count` doubles per rank (padded) from every rank to every rank, on the host
static std::vector<double> host_allgather(knp_ctx* c, const std::vector<double>& mine, int64_t count) {
  DevBuf<double> buf;
  buf.alloc((size_t)c->comm.world * count);
  h2d(buf.p + (int64_t)c->comm.rank * count, mine.data(), mine.size() * sizeof(double), cs(c));
  c->comm.allgather(cs(c), buf.p, count);
  return buf.download(cs(c));
}

// Turn the current coarsest distributed level D = lev.back() (pattern + setup values in G.coarse,
// owned rows, local columns) into a GLOBAL matrix known to every rank, appended as the
// replicated level lev[rep_from]; G.coarse becomes that global matrix so that the serial
// coarsening code continues on it unchanged.
static void replicate_level(knp_ctx* c, GalerkinPlan& G) {
  AmgPlan& amg = c->amg;
  Comm& comm = c->comm;
  const int world = comm.world, rank = comm.rank;
  AmgLevelPlan& D = amg.lev.back();
  const HostCsr& A = G.coarse;
  // sizes of every rank's share
  std::vector<double> cnt(2 * (size_t)world, 0.0);
  cnt[2 * rank] = (double)D.n; cnt[2 * rank + 1] = (double)D.nnz;
  global_sum(c, cnt.data(), 2 * world);
  std::vector<int64_t> off(world + 1, 0);
  int64_t bstride = 1, vstride = 1;
  for (int r = 0; r < world; ++r) {
    off[r + 1] = off[r] + (int64_t)cnt[2 * r];
    bstride = std::max<int64_t>(bstride, (int64_t)cnt[2 * r]);
    vstride = std::max<int64_t>(vstride, (int64_t)cnt[2 * r + 1]);
  }
  const int64_t m = off[world];
  if (m >= 2147483647 / 64) fail("replicate_level: level too large");
  // local unknown -> global row
  std::vector<int32_t> xmap((size_t)D.nloc);
  for (int64_t i = 0; i < D.n; ++i) xmap[i] = (int32_t)(off[rank] + i);
  for (int64_t g = 0; g < D.halo.n_ghost; ++g) xmap[D.n + g] = (int32_t)(off[D.halo.ghost_rank[g]] + D.halo.ghost_id[g]);
  // every rank's rows: lengths, global columns, setup values
  std::vector<double> len((size_t)bstride, 0.0), col((size_t)vstride, 0.0), val((size_t)vstride, 0.0);
  for (int64_t i = 0; i < D.n; ++i) len[i] = (double)(A.ptr[i + 1] - A.ptr[i]);
  for (int64_t k = 0; k < D.nnz; ++k) { col[k] = (double)xmap[A.col[k]]; val[k] = A.val[k]; }
  const std::vector<double> all_len = host_allgather(c, len, bstride);
  const std::vector<double> all_col = host_allgather(c, col, vstride);
  const std::vector<double> all_val = host_allgather(c, val, vstride);
  GalerkinPlan R;
  HostCsr& B = R.coarse;
  B.n = m;
  B.ptr.assign(m + 1, 0);
  std::vector<int32_t> bmap((size_t)m);
  R.gptr.push_back(0);
  for (int r = 0; r < world; ++r) {
    int64_t k = 0;
    for (int64_t i = 0; i < (int64_t)cnt[2 * r]; ++i) {
      const int64_t row = off[r] + i;
      bmap[row] = (int32_t)(r * bstride + i);
      const int64_t l = (int64_t)all_len[r * bstride + i];
      for (int64_t e = 0; e < l; ++e, ++k) {
        B.col.push_back((int32_t)all_col[r * vstride + k]);
        B.val.push_back(all_val[r * vstride + k]);
        R.gidx.push_back((int32_t)(r * vstride + k));     // position in the all-gathered value buffer
        R.gptr.push_back((int32_t)R.gidx.size());
      }
      B.ptr[row + 1] = (int32_t)B.col.size();
    }
  }
  B.pos.resize(B.col.size());
  std::iota(B.pos.begin(), B.pos.end(), 0);
  amg.rep_vstride = vstride; amg.rep_bstride = bstride;
  amg.rep_val.alloc((size_t)world * vstride);
  {
    // the right-hand side all-gather as an exchange plan: every rank (self included) receives
    // this rank's padded segment; segment r of the buffer comes from rank r
    HaloPlan& G = amg.rep_plan;
    G = HaloPlan();
    G.n_own = 0; G.n_ghost = (int64_t)world * bstride;
    G.ranks.resize(world);
    G.send_off.assign(world + 1, 0); G.recv_off.assign(world + 1, 0);
    for (int r = 0; r < world; ++r) {
      G.ranks[r] = r;
      for (int64_t k = 0; k < bstride; ++k) G.h_send_idx.push_back((int32_t)(rank * bstride + k));
      G.send_off[r + 1] = (int64_t)(r + 1) * bstride;
      G.recv_off[r + 1] = (int64_t)(r + 1) * bstride;
    }
    G.upload(cs(c));
  }
  amg.rep_bmap.upload(bmap, cs(c));
  amg.rep_xmap.upload(xmap, cs(c));
  amg.rep_from = amg.lev.size();
  amg.lev.emplace_back();
  AmgLevelPlan& T0 = amg.lev.back();
  T0.halo.n_own = m; T0.halo.n_ghost = 0;
  HostTransfer none;   // no transfer operator between a level and its replica (all-gather instead)
  none.pptr.assign(1, 0); none.rptr.assign(1, 0);
  upload_level(c, T0, R, none);
  G = std::move(R);
}

extern "C" int knp_amg_setup(knp_ctx* ctx, double theta, int max_levels, int coarse_size) {
  KNP_TRY
  if (!ctx->emi_assembled) fail("knp_amg_setup: assemble the EMI system first (strength of connection needs values)");
  if (max_levels < 2) max_levels = 2;
  if (coarse_size < 1) coarse_size = 1;
  if (coarse_size > 1024) coarse_size = 1024;
  AmgPlan& amg = ctx->amg;
  amg.lev.clear(); amg.ready = false;
  amg.n0 = ctx->n;
  // fine values of B (A's off-diagonal slots + Bdiag) for the setup-time strength graph
  std::vector<double> fine = ctx->A_emi.download(ctx->stream);
  HostCsr A0 = level0_csr(ctx);
  std::vector<int32_t> agg_own;
  int64_t ncoarse = vertex_injection(ctx, agg_own);
  amg.lev.emplace_back();
  std::vector<int32_t> agg = extend_aggregates(ctx, ctx->halo0, agg_own, ncoarse, amg.lev.back().halo);
  HostTransfer T = transfer_from_aggregates(agg, ncoarse, ctx->n_own);
  GalerkinPlan G = galerkin_plan(A0, T);
  galerkin_numeric_host(G, fine);
  upload_level(ctx, amg.lev.back(), G, T);
  { std::vector<int32_t>().swap(A0.col); std::vector<int32_t>().swap(A0.pos); }
  double gn = (double)G.coarse.n;          // rows of the current coarsest level over all ranks
  global_sum(ctx, &gn, 1);
  amg.rep_from = (size_t)-1;
  bool replicated = false;                 // G.coarse is a global matrix held by every rank
  while ((int)amg.lev.size() + 1 < max_levels && gn > coarse_size) {
    if (ctx->comm.active() && !replicated && gn <= (double)replicate_threshold()) {
      replicate_level(ctx, G);             // appends the global copy of lev.back()
      replicated = true;
      continue;
    }
    std::vector<int32_t> ag2;
    int64_t na = aggregate(G.coarse, theta, ag2);
    // Small levels are latency-bound (each costs three sweeps of a few microseconds per cycle,
    // whatever its size): below `double_rows` rows a level is coarsened TWICE in one step (the
    // aggregates of the aggregates), which halves the number of small levels.  Local levels
    // only (single part, or the replicated tail).
    if ((replicated || !ctx->comm.active()) && G.coarse.n <= double_coarsening_rows() && na > coarse_size) {
      HostTransfer Tm = transfer_from_aggregates(ag2, na);
      std::vector<double> vm = G.coarse.val;
      GalerkinPlan Gm = galerkin_plan(G.coarse, Tm);
      galerkin_numeric_host(Gm, vm);
      std::vector<int32_t> ag3;
      const int64_t nb = aggregate(Gm.coarse, theta, ag3);
      if (nb >= 1 && nb < na * 0.9) {
        for (auto& a : ag2) a = ag3[a];
        na = nb;
      }
    }
    double gna = (double)na;
    if (!replicated) global_sum(ctx, &gna, 1);
    if (gna >= gn * 0.9 || gna < 1) break;  // coarsening stalled
    amg.lev.emplace_back();
    AmgLevelPlan& fineL = amg.lev[amg.lev.size() - 2];
    std::vector<int32_t> ag2_all;
    if (replicated) { ag2_all = ag2; amg.lev.back().halo.n_own = na; amg.lev.back().halo.n_ghost = 0; }
    else ag2_all = extend_aggregates(ctx, fineL.halo, ag2, na, amg.lev.back().halo);
    HostTransfer T2 = transfer_from_aggregates(ag2_all, na, G.coarse.n);
    std::vector<double> vals = G.coarse.val;
    GalerkinPlan G2 = galerkin_plan(G.coarse, T2);
    galerkin_numeric_host(G2, vals);
    upload_level(ctx, amg.lev.back(), G2, T2);
    G = std::move(G2);
    gn = gna;
  }
  // last level: dense inverse of the GLOBAL matrix, replicated on every rank
  if (replicated) {
    amg.m_dense = amg.lev.back().n;
    amg.dense_off = 0;
  } else {
    AmgLevelPlan& L = amg.lev.back();
    const int world = ctx->comm.world, rank = ctx->comm.rank;
    if (world > 256) fail("knp_amg_setup: at most 256 ranks");
    std::vector<double> cnt(world, 0.0);
    cnt[rank] = (double)L.n;
    global_sum(ctx, cnt.data(), world);
    std::vector<int64_t> off(world + 1, 0);
    for (int r = 0; r < world; ++r) off[r + 1] = off[r] + (int64_t)cnt[r];
    amg.m_dense = off[world];
    amg.dense_off = off[rank];
    std::vector<int32_t> map((size_t)L.nloc);
    for (int64_t i = 0; i < L.n; ++i) map[i] = (int32_t)(off[rank] + i);
    for (int64_t g = 0; g < L.halo.n_ghost; ++g) map[L.n + g] = (int32_t)(off[L.halo.ghost_rank[g]] + L.halo.ghost_id[g]);
    amg.dense_map.upload(map, ctx->stream);
  }
  if (amg.m_dense > 4096) fail("knp_amg_setup: coarsest level too large for the dense solve (" +
                                std::to_string(amg.m_dense) + " rows); raise max_levels");
  // the small levels run as one cooperative launch per cycle (coarse_tail_kernel)
  amg.tail_from = (size_t)-1; amg.tail_blocks = 0;
#ifndef KNP_EMU
  {
    const char* e = getenv("KNP_AMG_TAIL");
    if (!(e && e[0] == '0')) {
      const char* tr = getenv("KNP_TAIL_ROWS");
      const int64_t tail_rows = tr ? atoll(tr) : TAIL_MAX_ROWS;
      size_t first = amg.lev.size();
      for (size_t l = amg.lev.size(); l-- > 0;) {
        const bool local = !ctx->comm.active() || l >= amg.rep_from;   // no halos inside the tail
        if (!local || amg.lev[l].n > tail_rows || !amg.lev[l].t_unit) break;
        first = l;
      }
      if (first == amg.rep_from) ++first;   // the hand-over into the replica stays a regular step
      if (first + 2 <= amg.lev.size() && amg.lev.size() - first <= (size_t)TAIL_MAX_LEVELS) {
        amg.tail_from = first;
        amg.tail_blocks = TAIL_CTAS;
      }
    }
  }
#endif
#ifndef KNP_EMU
  if (ctx->comm.active() && ctx->comm.p2p.on) {
    // every exchange channel a concurrently running solve may use is registered now, on the
    // main thread (registration talks NCCL; the worker threads must not)
    const int nws = ctx->P.N;   // main + one per solved ion
    ctx->comm.prepare_plan(ctx->stream, ctx->halo0, nws);
    for (size_t l = 0; l < amg.lev.size() && l < amg.rep_from; ++l) ctx->comm.prepare_plan(ctx->stream, amg.lev[l].halo, nws);
    if (amg.rep_from != (size_t)-1) ctx->comm.prepare_plan(ctx->stream, amg.rep_plan, nws);
  }
#endif
#ifndef KNP_EMU
  preload_solver_kernels(ctx);
#endif
  alloc_values(ctx, ctx->amg_emi);
  for (int k = 0; k < ctx->P.N - 1; ++k) alloc_values(ctx, ctx->amg_knp[k]);
  amg.ready = true;
  KNP_CATCH
}

extern "C" int knp_amg_info(knp_ctx* ctx, int64_t* nlevels, int64_t* rows, int64_t* nnz, int cap) {
  KNP_TRY
  if (!ctx->amg.ready) fail("AMG not set up");
  *nlevels = (int64_t)ctx->amg.lev.size() + 1;
  // owned rows / stored entries of this rank per level
  if (cap > 0) { rows[0] = ctx->n_own; nnz[0] = ctx->nnz_export; }
  for (size_t l = 0; l < ctx->amg.lev.size() && (int)l + 1 < cap; ++l) {
    rows[l + 1] = ctx->amg.lev[l].n; nnz[l + 1] = ctx->amg.lev[l].nnz;
  }
  KNP_CATCH
}

extern "C" int knp_solver_options(knp_ctx* ctx, int pc, int nu_pre, int nu_post, int gamma,
                                  double omega, int gmres_restart, int knp_min_it) {
  KNP_TRY
  if (pc < 0 || pc > 1) fail("pc must be 0 (block-Jacobi) or 1 (AMG)");
  if (nu_pre < 1 || nu_post < 0 || gamma < 1 || gamma > 2) fail("bad cycle parameters");
  ctx->amg_emi.omega = 0.0; ctx->amg_emi.solves = 0;
  for (int k = 0; k < MAX_IONS; ++k) { ctx->amg_knp[k].omega = 0.0; ctx->amg_knp[k].solves = 0; }
  if (gmres_restart < 1 || gmres_restart > 200) fail("bad GMRES restart");
  ctx->opt.pc = pc; ctx->opt.nu_pre = nu_pre; ctx->opt.nu_post = nu_post; ctx->opt.gamma = gamma;
  ctx->opt.omega = omega; ctx->opt.restart = gmres_restart; ctx->opt.knp_min_it = knp_min_it;
  KNP_CATCH
}

// ---------------------------------------------------------------------------------
// numeric refresh of the hierarchy after a re-assembly
// ---------------------------------------------------------------------------------
static CsrMat csr_of(const AmgLevelPlan& L, const AmgValues& V, size_t l) {
  CsrMat M; M.n = L.n; M.ptr = L.ptr.p; M.col = L.col.p; M.val = V.val[l].p;
  return M;
}

// lambda_max(Dinv A) by power iteration (the block-Jacobi smoother converges iff
// omega * lambda_max < 2; the matrices change slowly in time, so the estimate is redone
// only every OMEGA_PERIOD refreshes)
constexpr int OMEGA_PERIOD = 200;
static double estimate_lambda_max(knp_ctx* c, AmgValues& V, const BellMat& A, const double* dinv) {
  const int64_t n = c->n_own;
  double* v = V.x0.p; double* w = V.t0.p; double* u = V.r0.p;
  std::vector<double> h(n);
  for (int64_t i = 0; i < n; ++i) h[i] = 1.0 + 0.5 * sin(1.7 * (double)i) + ((i * 2654435761u) % 1024) / 1024.0;
  h2d(v, h.data(), n * sizeof(double), cs(c));
  double lam = 1.0;
  for (int it = 0; it < 12; ++it) {
    bell_spmv(c, A, v, nullptr, u, 0);
    block_apply(c, dinv, u, w, 1.0, 0);
    const double nv = sqrt(dot_host(c, v, v)), nw = sqrt(dot_host(c, w, w));
    if (!(nv > 0.0) || !(nw > 0.0)) break;
    lam = nw / nv;
    ScaleKernel k{1.0 / nw, w, v};
    parallel_for(cs(c), n, k);
  }
  return lam;
}

static void amg_refresh(knp_ctx* c, AmgValues& V, const BellMat& A0, const double* fine_values, const double* diag_blocks) {
  knp_stream_t s = cs(c);
  block_inverse(c, diag_blocks, V.binv.p);
  if (c->opt.pc_fp32) {
    // single-precision copies for the level-0 sweeps (they stay as they are until the next refresh)
    const int64_t ss = c->slot_stride();
    if (V.a32.n != (size_t)((c->nd + 1) * ss)) { V.a32.alloc((size_t)(c->nd + 1) * ss); V.binv32.alloc(ss); }
    { DemoteKernel k{A0.diag, V.a32.p}; parallel_for(s, ss, k); }
    { DemoteKernel k{A0.off + ss, V.a32.p + ss}; parallel_for(s, (int64_t)c->nd * ss, k); }
    { DemoteKernel k{V.binv.p, V.binv32.p}; parallel_for(s, ss, k); }
  }
  if (c->opt.omega > 0.0) V.omega = c->opt.omega;
  else if (V.omega <= 0.0 || ++V.age >= OMEGA_PERIOD) {
    V.omega = 4.0 / (3.0 * 1.05 * estimate_lambda_max(c, V, A0, V.binv.p));
    V.age = 0;
  }
  const double* fine = fine_values;
  AmgPlan& amg = c->amg;
  for (size_t l = 0; l < amg.lev.size(); ++l) {
    AmgLevelPlan& L = amg.lev[l];
    if (l == amg.rep_from) {
      // the replica's values: every rank's share of the level above, all-gathered
      AmgLevelPlan& D = amg.lev[l - 1];
      d2d(amg.rep_val.p + (int64_t)c->comm.rank * amg.rep_vstride, V.val[l - 1].p, D.nnz * sizeof(double), s);
      c->comm.allgather(s, amg.rep_val.p, amg.rep_vstride);
      fine = amg.rep_val.p;
    }
    GalerkinKernel g{L.gptr.p, L.gidx.p, L.g_unit ? nullptr : L.gw.p, fine, V.val[l].p};
    parallel_for(s, L.nnz, g);
    if (l + 1 != amg.rep_from) {             // the level that is replicated is never smoothed itself
      CsrL1DiagKernel dk{csr_of(L, V, l), V.dinv[l].p};
      parallel_for(s, L.n, dk);
      if (l < amg.rep_from) c->comm.halo(s, L.halo, V.dinv[l].p);   // the fused sweeps read dinv of ghost columns
    }
    fine = V.val[l].p;
  }
  const size_t last = c->amg.lev.size() - 1;
  const int64_t m = c->amg.m_dense;
  dev_zero(V.dense.p, (size_t)m * m * sizeof(double), s);
  const bool dense_distributed = c->comm.active() && last < amg.rep_from;
  CsrToDenseKernel tk{csr_of(c->amg.lev[last], V, last), V.dense.p, m, c->amg.dense_off,
                      dense_distributed ? c->amg.dense_map.p : nullptr};
  parallel_for(s, c->amg.lev[last].n, tk);
  if (dense_distributed) c->comm.allreduce(s, V.dense.p, m * m);   // every rank contributes its rows
  dense_inverse_device(s, (int)m, V.dense.p, V.colbuf.p);
}

// lagged refresh policy (SolverOptions::refresh_period)
static bool refresh_due(knp_ctx* c, AmgValues& V) {
  const int P = c->opt.refresh_period;
  if (P <= 1 || V.solves == 0) return true;
  if (V.solves % P == 0) return true;
  return V.last_iters > V.fresh_iters + V.fresh_iters / 2 + 2;   // the stale hierarchy has become costly
}
static void note_solve(AmgValues& V, bool refreshed, int iters) {
  if (refreshed) { V.fresh_iters = iters; V.solves = 0; }
  V.last_iters = iters;
  V.solves++;
}

// ---------------------------------------------------------------------------------
// cycle
// ---------------------------------------------------------------------------------
static void transfer(knp_ctx* c, int64_t nrows, const int32_t* ptr, const int32_t* idx, const double* w,
                     const double* x, double* y, int add) {
  TransferKernel k{nrows, ptr, idx, w, x, y, add};
  parallel_for(cs(c), nrows, k);
}

// solve level l (>= 1, index into lev = l-1) approximately: L.x <- cycle(L.b)
// On return the ghost entries of L.x are valid when ghost_x is set (the parent's fused
// prolongation+smoothing sweep reads them).
static void coarse_cycle(knp_ctx* c, AmgValues& V, size_t li, bool ghost_x) {
  knp_stream_t s = cs(c);
  AmgLevelPlan& L = c->amg.lev[li];
  LevelVectors& Lv = V.vec[li];
  Comm& comm = c->comm;
  const int wid = ws(c).id;
  const bool dist = comm.active() && li < c->amg.rep_from;   // this level's vectors have ghosts
#ifndef KNP_EMU
  if (li == c->amg.tail_from && c->opt.nu_pre == 1 && c->opt.nu_post == 1 && c->opt.gamma == 1) {
    AmgPlan& amg = c->amg;
    TailArgs a;
    a.nlev = (int)(amg.lev.size() - li);
    for (int k = 0; k < a.nlev; ++k) {
      AmgLevelPlan& T = amg.lev[li + k];
      TailLevel& t = a.L[k];
      t.n = T.n; t.ptr = T.ptr.p; t.col = T.col.p; t.val = V.val[li + k].p; t.dinv = V.dinv[li + k].p;
      t.b = V.vec[li + k].b.p; t.x = V.vec[li + k].x.p; t.r = V.vec[li + k].r.p;
      t.rptr = T.rptr.p; t.ridx = T.ridx.p; t.agg = T.pidx.p;
    }
    a.denseT = V.dense.p;
    ++launch_counter();
    coarse_tail_kernel<<<TAIL_CTAS, TAIL_THREADS, 0, s>>>(a);
    KNP_CUDA(cudaGetLastError());
    return;
  }
#endif
  if (li + 1 == c->amg.rep_from) {
    // hand over to the replicated rest of the hierarchy: all-gather the right-hand side, run
    // the remaining cycle redundantly, pick this rank's owned and ghost unknowns
    AmgPlan& amg = c->amg;
    AmgLevelPlan& T0 = amg.lev[li + 1];
    LevelVectors& T0v = V.vec[li + 1];
    d2d(V.rep_b.p + (int64_t)comm.rank * amg.rep_bstride, Lv.b.p, L.n * sizeof(double), s);
#ifndef KNP_EMU
    if (comm.p2p.on && comm.world <= P2P_MAX_NB) comm.halo(s, amg.rep_plan, V.rep_b.p, wid);
    else
#endif
      comm.allgather(s, V.rep_b.p, amg.rep_bstride);
    { GatherMapKernel k{V.rep_b.p, amg.rep_bmap.p, T0v.b.p}; parallel_for(s, T0.n, k); }
    coarse_cycle(c, V, li + 1, false);
    { GatherMapKernel k{T0v.x.p, amg.rep_xmap.p, Lv.x.p}; parallel_for(s, L.nloc, k); }
    return;
  }
  if (li + 1 == c->amg.lev.size()) {
    if (!dist) {
      DenseMatvecKernel k{L.n, V.dense.p, Lv.b.p, Lv.x.p};
      parallel_for(s, L.n, k, 64);
      return;
    }
    // the global right-hand side is summed over ranks, every rank applies the replicated
    // inverse and picks its owned and ghost unknowns
    AmgPlan& amg = c->amg;
    const int64_t m = amg.m_dense;
    dev_zero(V.dense_b.p, (size_t)m * sizeof(double), s);
    { ScatterOffsetKernel k{Lv.b.p, V.dense_b.p, amg.dense_off}; parallel_for(s, L.n, k); }
    comm.allreduce(s, V.dense_b.p, m, wid);
    { DenseMatvecKernel k{m, V.dense.p, V.dense_b.p, V.dense_x.p}; parallel_for(s, m, k, 64); }
    { GatherMapKernel k{V.dense_x.p, amg.dense_map.p, Lv.x.p}; parallel_for(s, L.nloc, k); }
    return;
  }
  CsrMat A = csr_of(L, V, li);
  AmgLevelPlan& C = c->amg.lev[li + 1];
  LevelVectors& Cv = V.vec[li + 1];
  if (c->opt.nu_pre == 1 && c->opt.nu_post == 1 && c->opt.gamma == 1 && C.t_unit) {
    // fused V(1,1) path: two kernels down (smooth+residual, restrict), one up (prolong+smooth)
    if (dist) comm.halo(s, L.halo, Lv.b.p, wid);
    { CoarseResidualKernel k{A, V.dinv[li].p, Lv.b.p, Lv.x.p, Lv.r.p}; parallel_rows<8>(s, L.n, k); }
    if (L.halo.n_ghost > 0) {   // x = dinv b on the ghost unknowns too (read by the sweep up)
      DiagScaleKernel k{V.dinv[li].p + L.n, Lv.b.p + L.n, Lv.x.p + L.n, 1.0};
      parallel_for(s, L.halo.n_ghost, k);
    }
    { TransferRowsKernel k{C.rptr.p, C.ridx.p, nullptr, Lv.r.p, Cv.b.p, 0}; parallel_rows<8>(s, C.n, k); }
    coarse_cycle(c, V, li + 1, true);
    { CoarseUpKernel k{A, V.dinv[li].p, Lv.b.p, Lv.x.p, C.pidx.p, Cv.x.p, Lv.t.p}; parallel_rows<8>(s, L.n, k); }
    std::swap(Lv.x.p, Lv.t.p);
    if (ghost_x && dist) comm.halo(s, L.halo, Lv.x.p, wid);
    return;
  }
  // pre-smoothing from a zero guess
  { DiagScaleKernel k{V.dinv[li].p, Lv.b.p, Lv.x.p, 1.0}; parallel_for(s, L.n, k); }
  for (int it = 1; it < c->opt.nu_pre; ++it) {
    if (dist) comm.halo(s, L.halo, Lv.x.p, wid);
    CsrJacobiKernel k{A, V.dinv[li].p, Lv.b.p, Lv.x.p, Lv.t.p, 1.0};
    parallel_for(s, L.n, k);
    std::swap(Lv.x.p, Lv.t.p);
  }
  for (int g = 0; g < c->opt.gamma; ++g) {
    if (dist) comm.halo(s, L.halo, Lv.x.p, wid);
    { CsrSpmvKernel k{A, Lv.x.p, Lv.b.p, Lv.r.p, 1}; parallel_for(s, L.n, k); }
    transfer(c, C.n, C.rptr.p, C.ridx.p, C.t_unit ? nullptr : C.rw.p, Lv.r.p, Cv.b.p, 0);
    coarse_cycle(c, V, li + 1, false);
    transfer(c, L.n, C.pptr.p, C.pidx.p, C.t_unit ? nullptr : C.pw.p, Cv.x.p, Lv.x.p, 1);
  }
  for (int it = 0; it < c->opt.nu_post; ++it) {
    if (dist) comm.halo(s, L.halo, Lv.x.p, wid);
    CsrJacobiKernel k{A, V.dinv[li].p, Lv.b.p, Lv.x.p, Lv.t.p, 1.0};
    parallel_for(s, L.n, k);
    std::swap(Lv.x.p, Lv.t.p);
  }
  if (ghost_x && dist) comm.halo(s, L.halo, Lv.x.p, wid);
}

// z = M^-1 r
// `presmooth0` = false drops the pre-smoothing sweep of the DG level (a V(0,1) cycle there:
// the right-hand side is restricted directly, one matrix pass less per application).  The
// resulting operator is not symmetric: fine for GMRES (KNP), not used with CG (EMI).
// one V-cycle; T = the type the level-0 matrix A0 and its inverse diagonal blocks binv are stored in
template <typename T>
static void vcycle(knp_ctx* c, AmgValues& V, const BellMatT<T>& A0, const T* binv, const double* r, double* z,
                   bool presmooth0, int cheby) {
  AmgPlan& amg = c->amg;
  AmgLevelPlan& C = amg.lev[0];
  LevelVectors& Cv = V.vec[0];
  const double w = V.omega;
  double* x = V.x0.p; double* t = V.t0.p;
  if (cheby == 2 && presmooth0) {
    // degree-2 Chebyshev in Dinv A on [lmax / 4, lmax], lmax = 1.05 x the estimate behind V.omega; the
    // same polynomial before and after the coarse correction keeps the cycle symmetric (CG)
    const double lmax = 4.0 / (3.0 * w), lmin = 0.25 * lmax;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin);
    const double sigma = theta / delta, rho0 = 1.0 / sigma, rho1 = 1.0 / (2.0 * sigma - rho0);
    const double w1 = 1.0 / theta, w2 = 2.0 * rho1 / delta, beta = rho1 * rho0;
    block_apply(c, binv, r, t, w1, 0);                                   // x1 = w1 Dinv r         (x0 = 0)
    bell_jacobi_mom(c, A0, binv, r, t, (const double*)nullptr, x, w2, beta);   // x2
    bell_spmv(c, A0, x, r, V.r0.p, 1);
    { TransferRowsKernel k{C.rptr.p, C.ridx.p, C.t_unit ? nullptr : C.rw.p, V.r0.p, Cv.b.p, 0}; parallel_rows<8>(cs(c), C.n, k); }
    coarse_cycle(c, V, 0, false);
    transfer(c, c->n_own, C.pptr.p, C.pidx.p, C.t_unit ? nullptr : C.pw.p, Cv.x.p, x, 1);
    bell_jacobi(c, A0, binv, r, x, t, w1);                               // x1' from x0' = x
    bell_jacobi_mom(c, A0, binv, r, t, x, z, w2, beta);                  // x2' -> z
    return;
  }
  // (measured on B200: the fused prolongation + sweep is ~2 % SLOWER than prolongation and sweep
  // as two launches - 40 extra gathers per row in a kernel that otherwise runs at 0.9 of the
  // HBM roofline - so it is opt-in)
  if constexpr (std::is_same<T, double>::value) {
    if (c->opt.fuse_prolong && c->opt.nu_post == 1 && C.t_unit && c->opt.nu_pre == 1) {
      const double* rr = r;
      if (presmooth0) {
        block_apply(c, binv, r, x, w, 0);
        bell_spmv(c, A0, x, r, V.r0.p, 1);          // refreshes the ghost entries of x as well
        rr = V.r0.p;
      }
      { TransferRowsKernel k{C.rptr.p, C.ridx.p, nullptr, rr, Cv.b.p, 0}; parallel_rows<8>(cs(c), C.n, k); }
      coarse_cycle(c, V, 0, c->comm.active());        // ghost entries of C.x are read by the fused sweep
      bell_jacobi_prolong(c, A0, binv, r, presmooth0 ? x : nullptr, C.pidx.p, Cv.x.p, z, w);
      return;
    }
  }
  const double* rr = r;
  if (presmooth0) {
    block_apply(c, binv, r, x, w, 0);
    for (int it = 1; it < c->opt.nu_pre; ++it) { bell_jacobi(c, A0, binv, r, x, t, w); std::swap(x, t); }
    bell_spmv(c, A0, x, r, V.r0.p, 1);
    rr = V.r0.p;
  }
  { TransferRowsKernel k{C.rptr.p, C.ridx.p, C.t_unit ? nullptr : C.rw.p, rr, Cv.b.p, 0}; parallel_rows<8>(cs(c), C.n, k); }
  coarse_cycle(c, V, 0, false);
  transfer(c, c->n_own, C.pptr.p, C.pidx.p, C.t_unit ? nullptr : C.pw.p, Cv.x.p, x, presmooth0 ? 1 : 0);
  if (c->opt.nu_post == 0) { d2d(z, x, c->n * sizeof(double), cs(c)); return; }
  for (int it = 0; it < c->opt.nu_post; ++it) {
    double* out = (it + 1 == c->opt.nu_post) ? z : t;
    bell_jacobi(c, A0, binv, r, x, out, w);
    if (out != z) std::swap(x, t);
  }
  // keep the plan's buffers in their slots for the next call
  if (x != V.x0.p) std::swap(V.x0.p, V.t0.p);
}

static void precondition(knp_ctx* c, AmgValues& V, const BellMat& A0, const double* bj, const double* r, double* z,
                         bool presmooth0 = true) {
  const int cheby = presmooth0 ? c->opt.cheby : 1;      // the symmetric (EMI) cycle only
  if (c->opt.pc == 0 || !c->amg.ready) {
    block_apply(c, bj, r, z, 1.0, 0);
    return;
  }
  if (c->opt.pc_fp32 && V.a32.n) {
    BellMat32 A32;
    A32.nc = A0.nc; A32.nbr = A0.nbr; A32.off = V.a32.p; A32.diag = V.a32.p;
    vcycle<float>(c, V, A32, V.binv32.p, r, z, presmooth0, cheby);
  } else {
    vcycle<double>(c, V, A0, V.binv.p, r, z, presmooth0, cheby);
  }
}

// ---------------------------------------------------------------------------------
// CG (EMI)
// ---------------------------------------------------------------------------------
// A_emi is singular (pure Neumann: constants, solver.py:465-466).  The preconditioned
// residual is kept orthogonal to the constants, otherwise the round-off component of r
// along them, amplified by the mass-shifted preconditioner, puts a floor under ||M^-1 r||.
static void remove_mean(knp_ctx* c, double* v) {
  const double s = dot_host(c, v, c->kr_ones.p);
  AddConstKernel k{-s / c->n_global, v};
  parallel_for(cs(c), c->n_own, k);
}

static void ensure_krylov(knp_ctx* c) {
  const size_t n = c->n;
  KrylovWs& K = ws(c);
  if (K.scal.n == 0) { K.scal.alloc(1024); K.partial.alloc((size_t)DOT_MAX * RED_BLOCKS); }
  if (K.r.n != 2 * n) { K.r.alloc(2 * n); K.p.alloc(n); K.q.alloc(n); K.w.alloc(n); }
  if (c->kr_ones.n != n) {
    c->kr_ones.alloc(n);
    std::vector<double> one(n, 1.0);
    h2d(c->kr_ones.p, one.data(), n * sizeof(double), cs(c));
  }
  const size_t need = (size_t)(c->opt.restart + 1) * n;
  if (K.V.n < need) K.V.alloc(need);
  if (tl_ws == nullptr) {     // (worker threads find it set by the main thread)
    double ng = (double)c->n_own;
    global_sum(c, &ng, 1);
    c->n_global = ng;
  }
}

extern "C" int knp_solve_emi(knp_ctx* ctx, double rtol, double atol, int maxit, int* niter, double* resid) {
  KNP_TRY
  if (!ctx->emi_assembled) fail("knp_solve_emi: assemble first");
  knp_ctx* c = ctx;
  knp_stream_t s = cs(c);
  stream_sync(s);
  const double t0 = now_s();
  ensure_krylov(c);
  const int64_t n = c->n, no = c->n_own;   // local vector length (stride) / owned rows
  BellMat A = bell_of(c, 0), B = bell_of(c, 1);
  // preconditioner refresh (the reference rebuilds BoomerAMG at every setOperators)
  const bool refreshed = refresh_due(c, c->amg_emi);
  if (refreshed) {
    if (c->opt.pc == 1 && c->amg.ready) amg_refresh(c, c->amg_emi, B, c->A_emi.p, c->Bdiag());
    else { if (c->bj_emi.n != (size_t)c->slot_stride()) c->bj_emi.alloc(c->slot_stride()); block_inverse(c, c->Bdiag(), c->bj_emi.p); }
  }
  double* x = c->phi.p; double* r = ws(c).r.p; double* z = ws(c).r.p + n; double* p = ws(c).p.p; double* q = ws(c).q.p;
  const double* b = c->rhs_emi.p;
  if (c->opt.extrapolate_phi) {
    if (c->phi_old.n != (size_t)n) { c->phi_old.alloc(n); c->phi_old_valid = false; }
    // the first stored field is an initial condition, not a solution: extrapolate only once two
    // consecutive solutions exist (third solve on)
    if (!c->phi_old_valid) c->phi_hist = 0;
    if (c->phi_hist >= 2) { ExtrapolateKernel k{x, c->phi_old.p}; parallel_for(s, n, k); }
    else { d2d(c->phi_old.p, x, n * sizeof(double), s); c->phi_hist++; }
    c->phi_old_valid = true;
  }
  // Two reductions per iteration.  (1) {p.q, 1.q}: with 1.r known, the mean of the updated
  // residual is known before the update and is subtracted in the same pass, so r stays
  // orthogonal to the null space to round-off of ITS OWN size (an unprojected r drifts by
  // eps |b|, which the mass-shifted preconditioner amplifies until it dominates z).
  // (2) {r.z, z.z, 1.z, 1.r}: the mean of z is then a small component and is removed
  // analytically: mu = (1.z)/N,  r.(z-mu) = r.z - mu (1.r),  |z-mu|^2 = z.z - N mu^2.
  const double* ones = c->kr_ones.p;
  const double Ng = c->n_global;
  double sum_r = 0.0;
  auto fused_dots = [&](double& rz_out, double& zz_out, double& mu_out) {
    DotPairs P{{r, z, ones, ones}, {z, z, z, r}};
    pair_dot_device(s, no, 4, P, ws(c).partial.p, ws(c).scal.p);
    c->comm.allreduce(s, ws(c).scal.p, 4);
    double d[4];
    d2h(d, ws(c).scal.p, sizeof d, s);
    mu_out = d[2] / Ng;
    sum_r = d[3];
    rz_out = d[0] - mu_out * d[3];
    zz_out = fmax(d[1] - Ng * mu_out * mu_out, 0.0);
  };
  // reference norm ||M^-1 b||
  precondition(c, c->amg_emi, B, c->bj_emi.p, b, z);
  remove_mean(c, z);
  const double bnorm = sqrt(dot_host(c, z, z));
  const double tol = fmax(rtol * bnorm, atol);
  bell_spmv(c, A, x, b, r, 1);
  remove_mean(c, r);
  precondition(c, c->amg_emi, B, c->bj_emi.p, r, z);
  double rz, zz, mu;
  fused_dots(rz, zz, mu);
  double zn = sqrt(zz);
  int it = 0, best_it = 0;
  double best = zn;
  if (zn > tol) {
    { DirectionKernel k{z, mu, 0.0, p}; parallel_for(s, no, k); }
    for (it = 1; it <= maxit; ++it) {
      bell_spmv(c, A, p, nullptr, q, 0);
      double pq, sum_q;
      {
        DotPairs P{{p, ones, nullptr, nullptr}, {q, q, nullptr, nullptr}};
        pair_dot_device(s, no, 2, P, ws(c).partial.p, ws(c).scal.p);
        c->comm.allreduce(s, ws(c).scal.p, 2);
        double d[2];
        d2h(d, ws(c).scal.p, sizeof d, s);
        pq = d[0]; sum_q = d[1];
      }
      if (!(pq > 0.0)) {
        // at round-off level p can fall into the (constant) null space of A: the iteration
        // has converged as far as fp64 allows
        if (pq == pq && zn <= 1e-8 * bnorm) { --it; break; }
        if (pq == 0.0 || pq != pq) fail("knp_solve_emi: CG breakdown (p.Ap = " + std::to_string(pq) + ")");
        fail("knp_solve_emi: operator or preconditioner is indefinite");
      }
      const double alpha = rz / pq;
      { Axpy2ProjKernel k{alpha, p, q, x, r, (sum_r - alpha * sum_q) / Ng}; parallel_for(s, no, k); }
      precondition(c, c->amg_emi, B, c->bj_emi.p, r, z);
      double rz_new;
      fused_dots(rz_new, zz, mu);
      zn = sqrt(zz);
      if (zn <= tol) break;
      // attainable accuracy: a tolerance below the fp64 floor of this system is treated as
      // reached once the preconditioned residual has stagnated at round-off level
      if (zn < best) { best = zn; best_it = it; }
      else if (it - best_it >= 40 && best <= 1e-10 * bnorm) break;
      const double beta = rz_new / rz;
      rz = rz_new;
      { DirectionKernel k{z, mu, beta, p}; parallel_for(s, no, k); }
    }
    if (it > maxit) fail("knp_solve_emi: CG did not converge in " + std::to_string(maxit) +
                         " iterations (ksp_error_if_not_converged, solver.py:428)");
  }
  note_solve(c->amg_emi, refreshed, it);
  halo0(c, x);   // the assembly of the KNP system reads phi on the ghost cells
  stream_sync(s);
#ifndef KNP_EMU
  c->comm.check_p2p(s);
#endif
  if (niter) *niter = it;
  if (resid) *resid = zn;
  c->timers[T_EMI_SOLVE] += now_s() - t0;
  KNP_CATCH
}

// ---------------------------------------------------------------------------------
// GMRES (KNP), one ion at a time
// ---------------------------------------------------------------------------------
// preconditioner of ion's system after a re-assembly (the reference rebuilds BoomerAMG at
// every setOperators, solver.py:767)
static void knp_refresh(knp_ctx* c, int ion) {
  AmgValues& V = c->amg_knp[ion];
  V.refreshed_now = refresh_due(c, V);
  if (!V.refreshed_now) return;
  BellMat A = bell_of(c, 2 + ion);
  if (c->opt.pc == 1 && c->amg.ready) amg_refresh(c, V, A, c->A_knp[ion].p, c->A_knp[ion].p);
  else { if (c->bj_knp[ion].n != (size_t)c->slot_stride()) c->bj_knp[ion].alloc(c->slot_stride()); block_inverse(c, c->A_knp[ion].p, c->bj_knp[ion].p); }
}

static int gmres_one(knp_ctx* c, int ion, double rtol, double atol, int maxit, double* resid_out, bool refresh = true) {
  knp_stream_t s = cs(c);
  const int64_t n = c->n, no = c->n_own;   // stride of the Krylov basis / owned rows
  const int m = c->opt.restart;
  BellMat A = bell_of(c, 2 + ion);
  AmgValues& Vv = c->amg_knp[ion];
  if (refresh) knp_refresh(c, ion);
  const double* bj = c->bj_knp[ion].p;
  double* x = c->c[ion].p;
  const double* b = c->rhs_knp[ion].p;
  double* V = ws(c).V.p; double* w = ws(c).w.p; double* r = ws(c).r.p;
  double* hdev = ws(c).scal.p + 512;  // device copy of the current Hessenberg column / y
  const bool pre0 = c->opt.knp_presmooth0;
  precondition(c, Vv, A, bj, b, w, pre0);
  const double bnorm = sqrt(dot_host(c, w, w));
  const double tol = fmax(rtol * bnorm, atol);
  std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), y(m), hcol(m + 2);
  int it = 0;
  double res = 0.0;
  while (true) {
    bell_spmv(c, A, x, b, r, 1);
    precondition(c, Vv, A, bj, r, V, pre0);        // V0 = M^-1 (b - A x)
    const double beta = sqrt(dot_host(c, V, V));
    res = beta;
    if ((beta <= tol && it >= c->opt.knp_min_it) || it >= maxit || beta == 0.0) break;
    { ScaleKernel k{1.0 / beta, V, V}; parallel_for(s, no, k); }
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int j = 0;
    bool done = false;
    for (; j < m && it < maxit; ++j) {
      double* vj = V + (int64_t)j * n;
      double* vn = V + (int64_t)(j + 1) * n;
      bell_spmv(c, A, vj, nullptr, r, 0);
      precondition(c, Vv, A, bj, r, vn, pre0);     // w = M^-1 A v_j, built in the next basis slot
      // classical Gram-Schmidt with ONE reduction per step: h = V^T w and |w|^2 in the same
      // pass (w is basis slot j+1), the new norm from Pythagoras, update + normalisation fused;
      // h stays on the device for the update, the host reads it once for the Givens rotations
      multi_dot_device(s, no, n, j + 2, V, vn, ws(c).partial.p, hdev);
      c->comm.allreduce(s, hdev, j + 2, ws(c).id);
      // the Pythagorean norm carries a relative error of about eps |w|^2 / hn^2: keep it below
      // the requested tolerance, otherwise fall back to the explicit norm
      const double cancel_tol = fmax(1e-6, 100.0 * 2.2e-16 / fmax(rtol, 1e-16));
      { GsNormalizeKernel k{n, j + 1, V, hdev, vn, cancel_tol}; parallel_for(s, no, k); }
      d2h(hcol.data(), hdev, (j + 2) * sizeof(double), s);
      double hsum = 0.0;
      for (int i = 0; i <= j; ++i) hsum += hcol[i] * hcol[i];
      double hn2 = hcol[j + 1] - hsum;
      if (!(hn2 >= cancel_tol * hcol[j + 1] && hn2 > 0.0)) {
        // too much cancellation (w almost in span V): explicit norm of the updated vector
        hn2 = dot_host(c, vn, vn);
        if (hn2 > 0.0) { ScaleKernel k{1.0 / sqrt(hn2), vn, vn}; parallel_for(s, no, k); }
      }
      const double hn = hn2 > 0.0 ? sqrt(hn2) : 0.0;
      for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = hcol[i];
      H[(size_t)(j + 1) * m + j] = hn;
      for (int i = 0; i < j; ++i) {                // apply previous rotations
        const double a = H[(size_t)i * m + j], bq = H[(size_t)(i + 1) * m + j];
        H[(size_t)i * m + j] = cs[i] * a + sn[i] * bq;
        H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * bq;
      }
      const double a = H[(size_t)j * m + j], bq = H[(size_t)(j + 1) * m + j];
      const double den = hypot(a, bq);
      cs[j] = den > 0 ? a / den : 1.0; sn[j] = den > 0 ? bq / den : 0.0;
      H[(size_t)j * m + j] = den; H[(size_t)(j + 1) * m + j] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      ++it;
      res = fabs(g[j + 1]);
      if ((res <= tol && it >= c->opt.knp_min_it) || hn == 0.0) { ++j; done = true; break; }
    }
    // y = H^-1 g, x += V y
    const int k = j;
    for (int i = k - 1; i >= 0; --i) {
      double acc = g[i];
      for (int l = i + 1; l < k; ++l) acc -= H[(size_t)i * m + l] * y[l];
      y[i] = acc / H[(size_t)i * m + i];
    }
    if (k > 0) {
      h2d(hdev, y.data(), k * sizeof(double), s);
      CombineKernel ck{n, k, V, hdev, x};
      parallel_for(s, no, ck);
    }
    if (done || it >= maxit) {
      if (!done) { *resid_out = res; return -it; }
      break;
    }
  }
  *resid_out = res;
  note_solve(Vv, Vv.refreshed_now, it);
  return it;
}

extern "C" int knp_solve_knp(knp_ctx* ctx, double rtol, double atol, int maxit, int* niter, double* resid) {
  KNP_TRY
  if (!ctx->knp_assembled) fail("knp_solve_knp: assemble first");
  stream_sync(ctx->stream);
  const double t0 = now_s();
  ensure_krylov(ctx);
  int worst = 0;
  double rmax = 0.0;
  const int nion = ctx->P.N - 1;
  bool concurrent = false;
#ifndef KNP_EMU
  {
    // The ions' systems are independent (solver.py:550-594).  On one GPU they are solved
    // CONCURRENTLY: one host thread and one stream per ion, so the latency-bound parts of one
    // solve (small AMG levels, host round trips of the Krylov scalars) overlap the
    // bandwidth-bound kernels of the other.  (Multi-GPU runs keep them in sequence: the
    // exchange kernels of a plan must be issued in the same order on every rank.)
    // Multi-GPU: every worker talks to its peers on its own exchange channel (peer-memory
    // kernels only - NCCL calls must come from one thread in one order); the preconditioner
    // refresh, which all-gathers matrix values with NCCL, is done up front on the main thread.
    // KNP_CONCURRENT_IONS: unset = concurrent on a single GPU only; 0 = never; 1 = also across
    // ranks.  The multi-rank mode is still OPT-IN: it measured -16 % at N=2, but 4 of 10 runs at
    // N=2 ended in an exchange time-out once the per-solve preconditioner refresh (whose NCCL
    // all-gather had kept the ranks in step) became lagged.  Suspected cause: lazy kernel loading
    // (preload_solver_kernels above); 2 of 2 runs were clean with the preload + eager loading,
    // not yet enough to flip the default.
    const char* e = getenv("KNP_CONCURRENT_IONS");
    const bool dist = ctx->comm.active();
    const bool dist_ok = !dist || (e && e[0] == '1' && ctx->comm.p2p.on && ctx->comm.world <= P2P_MAX_NB &&
                                   nion + 1 <= P2P_MAX_WS && ctx->opt.pc == 1 && ctx->amg.ready &&
                                   (ctx->amg.rep_from != (size_t)-1 || ctx->amg.m_dense <= P2P_AR_MAX));
    concurrent = nion > 1 && dist_ok && !(e && e[0] == '0');
  }
  if (concurrent) {
    // Streams and workspace buffers are created HERE, on the main thread, while this device is
    // idle: cudaMalloc / cudaFree / stream creation may synchronise the whole device, and a
    // device-wide wait issued while a peer-memory exchange kernel of the other worker is spinning
    // on a neighbour rank (whose matching kernel may sit behind ITS allocation) can deadlock
    // across ranks.
    for (int ion = 0; ion < nion; ++ion) {
      KrylovWs& K = ctx->kr_ion[ion];
      K.id = 1 + ion;
      if (!K.own_stream) { KNP_CUDA(cudaStreamCreateWithFlags(&K.stream, cudaStreamNonBlocking)); K.own_stream = true; }
      tl_ws = &K;
      ensure_krylov(ctx);
      tl_ws = nullptr;
    }
    const bool refresh_in_worker = !ctx->comm.active();
    if (!refresh_in_worker)
      for (int ion = 0; ion < nion; ++ion) knp_refresh(ctx, ion);
    cudaEvent_t ready;
    KNP_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    KNP_CUDA(cudaEventRecord(ready, ctx->stream));
    std::vector<std::thread> th;
    std::vector<int> its(nion, 0);
    std::vector<double> ress(nion, 0.0);
    std::vector<std::string> errs(nion);
    for (int ion = 0; ion < nion; ++ion) {
      KrylovWs& K = ctx->kr_ion[ion];
      KNP_CUDA(cudaStreamWaitEvent(K.stream, ready, 0));
      th.emplace_back([ctx, ion, rtol, atol, maxit, refresh_in_worker, &its, &ress, &errs, &K]() {
        try {
          KNP_CUDA(cudaSetDevice(ctx->device));
          tl_ws = &K;
          Comm::in_worker() = true;
          its[ion] = gmres_one(ctx, ion, rtol, atol, maxit, &ress[ion], refresh_in_worker);
          if (its[ion] >= 0) halo0(ctx, ctx->c[ion].p);   // ghost cells of the new concentration
          stream_sync(K.stream);
        } catch (const std::exception& ex) { errs[ion] = ex.what(); }
        catch (...) { errs[ion] = "unknown error"; }
        Comm::in_worker() = false;
        tl_ws = nullptr;
      });
    }
    for (auto& t : th) t.join();
    cudaEventDestroy(ready);
    for (int ion = 0; ion < nion; ++ion) {
      if (!errs[ion].empty()) fail(errs[ion]);
      if (its[ion] < 0) fail("knp_solve_knp: GMRES did not converge for ion " + std::to_string(ion));
      worst = its[ion] > worst ? its[ion] : worst;
      rmax = ress[ion] > rmax ? ress[ion] : rmax;
    }
  }
#endif
  for (int ion = 0; ion < nion && !concurrent; ++ion) {
    double res = 0.0;
    const int it = gmres_one(ctx, ion, rtol, atol, maxit, &res);
    if (it < 0) fail("knp_solve_knp: GMRES did not converge for ion " + std::to_string(ion));
    halo0(ctx, ctx->c[ion].p);   // post-step and the next assembly read c on the ghost cells
    worst = it > worst ? it : worst;
    rmax = res > rmax ? res : rmax;
  }
  stream_sync(ctx->stream);
#ifndef KNP_EMU
  ctx->comm.check_p2p(ctx->stream);
#endif
  if (niter) *niter = worst;
  if (resid) *resid = rmax;
  ctx->timers[T_KNP_SOLVE] += now_s() - t0;
  KNP_CATCH
}

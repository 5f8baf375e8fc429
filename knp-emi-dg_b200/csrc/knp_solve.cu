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

// ---- batches of linear systems --------------------------------------------------------------
// Every operator below acts on a BATCH of nb independent linear systems that share the mesh, the
// sparsity pattern and the AMG plan but have their own matrix values and vectors: nb = 1 for the
// EMI system; nb = N_ions for the KNP solve, where the reference's mixed-space system
// (solver.py:168-169, 684-701) is block diagonal - the ions are uncoupled in the form
// (solver.py:550-594) - and is solved as ONE system: one Krylov space, one residual norm, one
// Hessenberg matrix, and every kernel launched once for all ions (blockIdx.y = ion).
struct Sys {
  AmgValues* V;          // numeric hierarchy and work vectors of this system
  BellMat A;             // level-0 matrix the preconditioner is built from (EMI: B_emi)
  const double* bj;      // inverse diagonal blocks (pc == 0: element block-Jacobi only)
};
struct Batch { int nb; Sys s[MAX_BATCH]; };
struct Vecs { double* p[MAX_BATCH]; };
static Vecs vecs1(const double* x) { Vecs v{}; v.p[0] = const_cast<double*>(x); return v; }
static Vecs offset(const Vecs& v, int nb, int64_t off) {
  Vecs o{};
  for (int b = 0; b < nb; ++b) o.p[b] = v.p[b] ? v.p[b] + off : nullptr;
  return o;
}
template <class Fn>
static void by_nd(knp_ctx* c, Fn&& fn) {
  if (c->nd == 3) fn(std::integral_constant<int, 3>{}); else fn(std::integral_constant<int, 4>{});
}

static KrylovWs& ws(knp_ctx* c) { return c->kr0; }
static knp_stream_t cs(knp_ctx* c) { return c->stream; }

// ---- small helpers ---------------------------------------------------------------------
// Multi-GPU: every kernel below runs over the OWNED rows (the first n_own entries of a
// vector); the operators that read a vector through the matrix first refresh its ghost
// entries from their owners (knp_comm.h) - one exchange for the whole batch.
static void halo_of(knp_ctx* c, HaloPlan& H, int nb, const Vecs& x) {
  if (c->comm.active()) c->comm.halo_batch(cs(c), H, nb, x.p);
}
static void halo0(knp_ctx* c, int nb, const Vecs& x) { halo_of(c, c->halo0, nb, x); }
static void halo0(knp_ctx* c, const double* x) { halo0(c, 1, vecs1(x)); }

// Overlap of the level-0 halo with interior-row work: the exchange kernel is issued on the
// context's second stream once everything queued so far on the main stream (the producer of x) is
// done; the main stream meanwhile runs the rows of the interior cells (no ghost columns) and joins
// the exchange before the rows next to the partition boundary.
static bool overlap_split(knp_ctx* c) {
#ifdef KNP_EMU
  (void)c;
  return false;
#else
  return c->overlap && c->comm.active() && c->nc_int * 4 >= c->nc_own;   // worth it: >= 25 % interior cells
#endif
}
static void halo0_fork(knp_ctx* c, int nb, const Vecs& x) {
#ifndef KNP_EMU
  KNP_CUDA(cudaEventRecord(c->ev_ready, c->stream));
  KNP_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
  c->comm.halo_batch(c->comm_stream, c->halo0, nb, x.p);
  KNP_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
#else
  (void)c; (void)nb; (void)x;
#endif
}
static void halo0_join(knp_ctx* c) {
#ifndef KNP_EMU
  KNP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
#else
  (void)c;
#endif
}
// launch a batch of level-0 row functors that read x through the matrix: with the exchange of x overlapped
template <class F>
static void rows_with_halo(knp_ctx* c, int nb, const Vecs& x, const BatchOf<F>& k, int nd, int block) {
  if (!overlap_split(c)) {
    halo0(c, nb, x);
    parallel_for_batch(cs(c), c->n_own, nb, k, block);
    return;
  }
  const int64_t ni = c->nc_int * nd;
  halo0_fork(c, nb, x);
  parallel_for_batch_range(cs(c), 0, ni, nb, k, block);          // interior cells: no ghost columns
  halo0_join(c);
  parallel_for_batch_range(cs(c), ni, c->n_own, nb, k, block);   // cells next to the partition boundary
}

template <typename T>
static void bell_spmv(knp_ctx* c, int nb, const BellMatT<T>* A, const Vecs& x, const Vecs& b, const Vecs& y, int mode) {
  by_nd(c, [&](auto nd) {
    constexpr int ND = decltype(nd)::value;
    BatchOf<BellSpmvKernel<ND, T>> k{};
    for (int s = 0; s < nb; ++s) k.f[s] = BellSpmvKernel<ND, T>{A[s], x.p[s], b.p[s], y.p[s], mode};
    rows_with_halo(c, nb, x, k, ND, 256);
  });
}
static void bell_spmv(knp_ctx* c, const BellMat& A, const double* x, const double* b, double* y, int mode) {
  bell_spmv<double>(c, 1, &A, vecs1(x), vecs1(b), vecs1(y), mode);
}
template <typename T>
static void block_apply(knp_ctx* c, int nb, const T* const* dinv, const Vecs& r, const Vecs& out, const double* w, int mode) {
  by_nd(c, [&](auto nd) {
    constexpr int ND = decltype(nd)::value;
    BatchOf<BlockDiagApplyKernel<ND, T>> k{};
    for (int s = 0; s < nb; ++s) k.f[s] = BlockDiagApplyKernel<ND, T>{dinv[s], r.p[s], out.p[s], w[s], mode};
    parallel_for_batch(cs(c), c->n_own, nb, k);
  });
}
static void block_apply(knp_ctx* c, const double* dinv, const double* r, double* out, double w, int mode) {
  block_apply<double>(c, 1, &dinv, vecs1(r), vecs1(out), &w, mode);
}
// xout = xin + w Dinv (b - A xin); MOM: second Chebyshev step, + beta (xin - xprev), xprev may hold nullptr (zero)
template <typename T, bool MOM = false>
static void bell_jacobi(knp_ctx* c, int nb, const BellMatT<T>* A, const T* const* dinv, const Vecs& b,
                        const Vecs& xin, const Vecs& xout, const double* w, const Vecs* xprev = nullptr, double beta = 0.0) {
  by_nd(c, [&](auto nd) {
    constexpr int ND = decltype(nd)::value;
    BatchOf<BellJacobiKernel<ND, T, MOM>> k{};
    for (int s = 0; s < nb; ++s) {
      k.f[s] = BellJacobiKernel<ND, T, MOM>{A[s], dinv[s], b.p[s], xin.p[s], xout.p[s], w[s]};
      if (MOM) { k.f[s].xprev = xprev ? xprev->p[s] : nullptr; k.f[s].beta = beta; }
    }
    rows_with_halo(c, nb, xin, k, ND, ND == 3 ? 192 : 256);
  });
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

// host-visible dot products over all ranks and over the systems of a batch (one sync each):
// out[i] = sum_s V_s[i] . w_s
static void dots_host(knp_ctx* c, int k, int nb, const Vecs& V, const Vecs& w, double* out_host) {
  DotBatch B;
  B.nb = nb;
  for (int s = 0; s < nb; ++s) { B.V[s] = V.p[s]; B.w[s] = w.p[s]; }
  multi_dot_batch(cs(c), c->n_own, c->n, k, B, ws(c).partial.p, ws(c).scal.p);
  c->comm.allreduce(cs(c), ws(c).scal.p, k);
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
static double dot_host(knp_ctx* c, int nb, const Vecs& x, const Vecs& y) {
  double v;
  dots_host(c, 1, nb, x, y, &v);
  return v;
}
static double dot_host(knp_ctx* c, const double* x, const double* y) { return dot_host(c, 1, vecs1(x), vecs1(y)); }

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
  // the host-side plan addresses matrix entries with 32-bit offsets: refuse meshes beyond that range
  // (~26 M cells per rank in 3D) instead of overflowing silently; partition the mesh over more GPUs
  if ((int64_t)(nd + 2) * ss > 2147483647LL || c->nnz_export > 2147483647LL)
    fail("knp_amg_setup: " + std::to_string(c->nc) + " cells on one rank exceed the 32-bit entry offsets of the AMG plan; "
         "use more ranks");
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
    // staging buffers sized for one exchange of all solved ions' vectors at once
    ctx->comm.batch_cap = ctx->P.N - 1 > 1 ? ctx->P.N - 1 : 1;
    const int nws = 1;
    ctx->comm.prepare_plan(ctx->stream, ctx->halo0, nws);
    for (size_t l = 0; l < amg.lev.size() && l < amg.rep_from; ++l) ctx->comm.prepare_plan(ctx->stream, amg.lev[l].halo, nws);
    if (amg.rep_from != (size_t)-1) ctx->comm.prepare_plan(ctx->stream, amg.rep_plan, nws);
  }
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
static void transfer(knp_ctx* c, int nb, int64_t nrows, const int32_t* ptr, const int32_t* idx, const double* w,
                     const Vecs& x, const Vecs& y, int add) {
  BatchOf<TransferKernel> k{};
  for (int s = 0; s < nb; ++s) k.f[s] = TransferKernel{nrows, ptr, idx, w, x.p[s], y.p[s], add};
  parallel_for_batch(cs(c), nrows, nb, k);
}
// y (=|+=) P xc for the transfer between a level and the coarser level C
static void prolong(knp_ctx* c, int nb, int64_t nrows, const AmgLevelPlan& C, const Vecs& xc, const Vecs& y, int add) {
  if (!C.t_unit) { transfer(c, nb, nrows, C.pptr.p, C.pidx.p, C.pw.p, xc, y, add); return; }
  BatchOf<ProlongUnitKernel> k{};
  for (int s = 0; s < nb; ++s) k.f[s] = ProlongUnitKernel{C.pidx.p, xc.p[s], y.p[s], add};
  parallel_for_batch(cs(c), nrows, nb, k);
}
static void restrict_rows(knp_ctx* c, int nb, const AmgLevelPlan& C, const Vecs& x, const Vecs& y) {
  BatchOf<TransferRowsKernel> k{};
  for (int s = 0; s < nb; ++s) k.f[s] = TransferRowsKernel{C.rptr.p, C.ridx.p, C.t_unit ? nullptr : C.rw.p, x.p[s], y.p[s], 0};
  parallel_rows_batch<8>(cs(c), C.n, nb, k);
}

// per-system vectors of one level
enum LevelVec { LV_B, LV_X, LV_R, LV_T };
static Vecs level_vecs(const Batch& B, size_t li, LevelVec which) {
  Vecs v{};
  for (int s = 0; s < B.nb; ++s) {
    LevelVectors& L = B.s[s].V->vec[li];
    v.p[s] = which == LV_B ? L.b.p : which == LV_X ? L.x.p : which == LV_R ? L.r.p : L.t.p;
  }
  return v;
}
static void swap_xt(const Batch& B, size_t li) {
  for (int s = 0; s < B.nb; ++s) std::swap(B.s[s].V->vec[li].x.p, B.s[s].V->vec[li].t.p);
}

// solve level l (>= 1, index into lev = l-1) approximately for every system of the batch:
// L.x <- cycle(L.b).  On return the ghost entries of L.x are valid when ghost_x is set.
static void coarse_cycle(knp_ctx* c, const Batch& B, size_t li, bool ghost_x) {
  knp_stream_t s = cs(c);
  const int nb = B.nb;
  AmgLevelPlan& L = c->amg.lev[li];
  Comm& comm = c->comm;
  const bool dist = comm.active() && li < c->amg.rep_from;   // this level's vectors have ghosts
  const Vecs Lb = level_vecs(B, li, LV_B), Lx = level_vecs(B, li, LV_X), Lr = level_vecs(B, li, LV_R), Lt = level_vecs(B, li, LV_T);
#ifndef KNP_EMU
  if (li == c->amg.tail_from && c->opt.nu_pre == 1 && c->opt.nu_post == 1 && c->opt.gamma == 1) {
    AmgPlan& amg = c->amg;
    TailBatch tb;
    for (int q = 0; q < nb; ++q) {
      AmgValues& V = *B.s[q].V;
      TailArgs& a = tb.s[q];
      a.nlev = (int)(amg.lev.size() - li);
      for (int k = 0; k < a.nlev; ++k) {
        AmgLevelPlan& T = amg.lev[li + k];
        TailLevel& t = a.L[k];
        t.n = T.n; t.ptr = T.ptr.p; t.col = T.col.p; t.val = V.val[li + k].p; t.dinv = V.dinv[li + k].p;
        t.b = V.vec[li + k].b.p; t.x = V.vec[li + k].x.p; t.r = V.vec[li + k].r.p;
        t.rptr = T.rptr.p; t.ridx = T.ridx.p; t.agg = T.pidx.p;
      }
      a.denseT = V.dense.p;
    }
    ++launch_counter();
    coarse_tail_kernel<<<TAIL_CTAS * nb, TAIL_THREADS, 0, s>>>(tb);
    KNP_CUDA(cudaGetLastError());
    return;
  }
#endif
  if (li + 1 == c->amg.rep_from) {
    // hand over to the replicated rest of the hierarchy: all-gather the right-hand sides, run
    // the remaining cycle redundantly, pick this rank's owned and ghost unknowns
    AmgPlan& amg = c->amg;
    AmgLevelPlan& T0 = amg.lev[li + 1];
    Vecs rep{};
    for (int q = 0; q < nb; ++q) {
      rep.p[q] = B.s[q].V->rep_b.p;
      d2d(rep.p[q] + (int64_t)comm.rank * amg.rep_bstride, Lb.p[q], L.n * sizeof(double), s);
    }
#ifndef KNP_EMU
    if (comm.p2p.on && comm.world <= P2P_MAX_NB) comm.halo_batch(s, amg.rep_plan, nb, rep.p);
    else
#endif
      for (int q = 0; q < nb; ++q) comm.allgather(s, rep.p[q], amg.rep_bstride);
    const Vecs T0b = level_vecs(B, li + 1, LV_B);
    { BatchOf<GatherMapKernel> k{}; for (int q = 0; q < nb; ++q) k.f[q] = GatherMapKernel{rep.p[q], amg.rep_bmap.p, T0b.p[q]};
      parallel_for_batch(s, T0.n, nb, k); }
    coarse_cycle(c, B, li + 1, false);
    const Vecs T0x = level_vecs(B, li + 1, LV_X);
    { BatchOf<GatherMapKernel> k{}; for (int q = 0; q < nb; ++q) k.f[q] = GatherMapKernel{T0x.p[q], amg.rep_xmap.p, Lx.p[q]};
      parallel_for_batch(s, L.nloc, nb, k); }
    return;
  }
  if (li + 1 == c->amg.lev.size()) {
    if (!dist) {
      BatchOf<DenseMatvecKernel> k{};
      for (int q = 0; q < nb; ++q) k.f[q] = DenseMatvecKernel{L.n, B.s[q].V->dense.p, Lb.p[q], Lx.p[q]};
      parallel_for_batch(s, L.n, nb, k, 64);
      return;
    }
    // the global right-hand side is summed over ranks, every rank applies the replicated
    // inverse and picks its owned and ghost unknowns (small meshes only: one system after the other)
    AmgPlan& amg = c->amg;
    const int64_t m = amg.m_dense;
    for (int q = 0; q < nb; ++q) {
      AmgValues& V = *B.s[q].V;
      dev_zero(V.dense_b.p, (size_t)m * sizeof(double), s);
      { ScatterOffsetKernel k{Lb.p[q], V.dense_b.p, amg.dense_off}; parallel_for(s, L.n, k); }
      comm.allreduce(s, V.dense_b.p, m);
      { DenseMatvecKernel k{m, V.dense.p, V.dense_b.p, V.dense_x.p}; parallel_for(s, m, k, 64); }
      { GatherMapKernel k{V.dense_x.p, amg.dense_map.p, Lx.p[q]}; parallel_for(s, L.nloc, k); }
    }
    return;
  }
  AmgLevelPlan& C = c->amg.lev[li + 1];
  const Vecs Cb = level_vecs(B, li + 1, LV_B);
  auto mat = [&](int q) { return csr_of(L, *B.s[q].V, li); };
  auto dinv = [&](int q) { return B.s[q].V->dinv[li].p; };
  if (c->opt.nu_pre == 1 && c->opt.nu_post == 1 && c->opt.gamma == 1 && C.t_unit) {
    // fused V(1,1) path: two kernels down (smooth+residual, restrict), one up (prolong+smooth)
    if (dist) halo_of(c, L.halo, nb, Lb);
    { BatchOf<CoarseResidualKernel> k{}; for (int q = 0; q < nb; ++q) k.f[q] = CoarseResidualKernel{mat(q), dinv(q), Lb.p[q], Lx.p[q], Lr.p[q]};
      parallel_rows_batch<8>(s, L.n, nb, k); }
    if (L.halo.n_ghost > 0) {   // x = dinv b on the ghost unknowns too (read by the sweep up)
      BatchOf<DiagScaleKernel> k{};
      for (int q = 0; q < nb; ++q) k.f[q] = DiagScaleKernel{dinv(q) + L.n, Lb.p[q] + L.n, Lx.p[q] + L.n, 1.0};
      parallel_for_batch(s, L.halo.n_ghost, nb, k);
    }
    restrict_rows(c, nb, C, Lr, Cb);
    coarse_cycle(c, B, li + 1, true);
    const Vecs Cx = level_vecs(B, li + 1, LV_X);
    { BatchOf<CoarseUpKernel> k{}; for (int q = 0; q < nb; ++q) k.f[q] = CoarseUpKernel{mat(q), dinv(q), Lb.p[q], Lx.p[q], C.pidx.p, Cx.p[q], Lt.p[q]};
      parallel_rows_batch<8>(s, L.n, nb, k); }
    swap_xt(B, li);
    if (ghost_x && dist) halo_of(c, L.halo, nb, level_vecs(B, li, LV_X));
    return;
  }
  // general path (other cycle parameters): sweep by sweep
  auto jacobi = [&]() {
    const Vecs x = level_vecs(B, li, LV_X), t = level_vecs(B, li, LV_T);
    if (dist) halo_of(c, L.halo, nb, x);
    BatchOf<CsrJacobiKernel> k{};
    for (int q = 0; q < nb; ++q) k.f[q] = CsrJacobiKernel{mat(q), dinv(q), Lb.p[q], x.p[q], t.p[q], 1.0};
    parallel_for_batch(s, L.n, nb, k);
    swap_xt(B, li);
  };
  { BatchOf<DiagScaleKernel> k{}; for (int q = 0; q < nb; ++q) k.f[q] = DiagScaleKernel{dinv(q), Lb.p[q], Lx.p[q], 1.0};
    parallel_for_batch(s, L.n, nb, k); }                      // pre-smoothing from a zero guess
  for (int it = 1; it < c->opt.nu_pre; ++it) jacobi();
  for (int g = 0; g < c->opt.gamma; ++g) {
    const Vecs x = level_vecs(B, li, LV_X);
    if (dist) halo_of(c, L.halo, nb, x);
    { BatchOf<CsrSpmvKernel> k{}; for (int q = 0; q < nb; ++q) k.f[q] = CsrSpmvKernel{mat(q), x.p[q], Lb.p[q], Lr.p[q], 1};
      parallel_for_batch(s, L.n, nb, k); }
    transfer(c, nb, C.n, C.rptr.p, C.ridx.p, C.t_unit ? nullptr : C.rw.p, Lr, Cb, 0);
    coarse_cycle(c, B, li + 1, false);
    prolong(c, nb, L.n, C, level_vecs(B, li + 1, LV_X), x, 1);
  }
  for (int it = 0; it < c->opt.nu_post; ++it) jacobi();
  if (ghost_x && dist) halo_of(c, L.halo, nb, level_vecs(B, li, LV_X));
}

// z = M^-1 r for every system of the batch: one V-cycle.
// `presmooth0` = false drops the pre-smoothing sweep of the DG level (a V(0,1) cycle there:
// the right-hand side is restricted directly, one matrix pass less per application).  The
// resulting operator is not symmetric: fine for GMRES (KNP), not used with CG (EMI).
// T = the type the level-0 matrices A0 and their inverse diagonal blocks binv are stored in.
template <typename T>
static void vcycle(knp_ctx* c, const Batch& B, const BellMatT<T>* A0, const T* const* binv, const Vecs& r, const Vecs& z,
                   bool presmooth0, int cheby) {
  AmgPlan& amg = c->amg;
  const int nb = B.nb;
  AmgLevelPlan& C = amg.lev[0];
  const Vecs Cb = level_vecs(B, 0, LV_B);
  double w[MAX_BATCH];
  Vecs x{}, t{}, r0{};
  for (int q = 0; q < nb; ++q) { AmgValues& V = *B.s[q].V; w[q] = V.omega; x.p[q] = V.x0.p; t.p[q] = V.t0.p; r0.p[q] = V.r0.p; }
  if (cheby == 2 && presmooth0) {
    // degree-2 Chebyshev in Dinv A on [lmax / 4, lmax], lmax = 1.05 x the estimate behind V.omega; the
    // same polynomial before and after the coarse correction keeps the cycle symmetric (CG)
    double w1[MAX_BATCH], w2[MAX_BATCH], beta = 0.0;
    for (int q = 0; q < nb; ++q) {
      const double lmax = 4.0 / (3.0 * w[q]), lmin = 0.25 * lmax;
      const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin);
      const double sigma = theta / delta, rho0 = 1.0 / sigma, rho1 = 1.0 / (2.0 * sigma - rho0);
      w1[q] = 1.0 / theta; w2[q] = 2.0 * rho1 / delta; beta = rho1 * rho0;     // (beta depends on the ratio only)
    }
    block_apply<T>(c, nb, binv, r, t, w1, 0);                               // x1 = w1 Dinv r         (x0 = 0)
    bell_jacobi<T, true>(c, nb, A0, binv, r, t, x, w2, nullptr, beta);       // x2
    bell_spmv<T>(c, nb, A0, x, r, r0, 1);
    restrict_rows(c, nb, C, r0, Cb);
    coarse_cycle(c, B, 0, false);
    prolong(c, nb, c->n_own, C, level_vecs(B, 0, LV_X), x, 1);
    bell_jacobi<T>(c, nb, A0, binv, r, x, t, w1);                            // x1' from x0' = x
    bell_jacobi<T, true>(c, nb, A0, binv, r, t, z, w2, &x, beta);            // x2' -> z
    return;
  }
  Vecs rr = r;
  if (presmooth0) {
    block_apply<T>(c, nb, binv, r, x, w, 0);
    for (int it = 1; it < c->opt.nu_pre; ++it) { bell_jacobi<T>(c, nb, A0, binv, r, x, t, w); std::swap(x, t); }
    bell_spmv<T>(c, nb, A0, x, r, r0, 1);
    rr = r0;
  }
  restrict_rows(c, nb, C, rr, Cb);
  coarse_cycle(c, B, 0, false);
  prolong(c, nb, c->n_own, C, level_vecs(B, 0, LV_X), x, presmooth0 ? 1 : 0);
  if (c->opt.nu_post == 0) {
    for (int q = 0; q < nb; ++q) d2d(z.p[q], x.p[q], c->n * sizeof(double), cs(c));
    return;
  }
  for (int it = 0; it < c->opt.nu_post; ++it) {
    const Vecs& out = (it + 1 == c->opt.nu_post) ? z : t;
    bell_jacobi<T>(c, nb, A0, binv, r, x, out, w);
    if (it + 1 != c->opt.nu_post) std::swap(x, t);
  }
}

static void precondition(knp_ctx* c, const Batch& B, const Vecs& r, const Vecs& z, bool presmooth0 = true) {
  const int cheby = presmooth0 ? c->opt.cheby : 1;      // the symmetric (EMI) cycle only
  const int nb = B.nb;
  if (c->opt.pc == 0 || !c->amg.ready) {
    const double* bj[MAX_BATCH]; double one[MAX_BATCH];
    for (int q = 0; q < nb; ++q) { bj[q] = B.s[q].bj; one[q] = 1.0; }
    block_apply<double>(c, nb, bj, r, z, one, 0);
    return;
  }
  bool fp32 = c->opt.pc_fp32;
  for (int q = 0; q < nb; ++q) fp32 = fp32 && B.s[q].V->a32.n;
  if (fp32) {
    BellMat32 A32[MAX_BATCH]; const float* binv[MAX_BATCH];
    for (int q = 0; q < nb; ++q) {
      AmgValues& V = *B.s[q].V;
      A32[q].nc = B.s[q].A.nc; A32[q].nbr = B.s[q].A.nbr; A32[q].off = V.a32.p; A32[q].diag = V.a32.p;
      binv[q] = V.binv32.p;
    }
    vcycle<float>(c, B, A32, binv, r, z, presmooth0, cheby);
  } else {
    BellMat A0[MAX_BATCH]; const double* binv[MAX_BATCH];
    for (int q = 0; q < nb; ++q) { A0[q] = B.s[q].A; binv[q] = B.s[q].V->binv.p; }
    vcycle<double>(c, B, A0, binv, r, z, presmooth0, cheby);
  }
}
static Batch batch1(AmgValues& V, const BellMat& A, const double* bj) {
  Batch B{};
  B.nb = 1; B.s[0] = Sys{&V, A, bj};
  return B;
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

// Krylov vectors: the main workspace serves the EMI solve and system 0 of the KNP batch, the
// further systems of the batch have their own vectors in kr_ion[1..]
static void ensure_vectors(knp_ctx* c, KrylovWs& K, bool basis) {
  const size_t n = c->n;
  if (K.r.n != 2 * n) { K.r.alloc(2 * n); K.p.alloc(n); K.q.alloc(n); K.w.alloc(n); }
  const size_t need = (size_t)(c->opt.restart + 1) * n;
  if (basis && K.V.n < need) K.V.alloc(need);
}
static void ensure_krylov(knp_ctx* c, int nsys = 1) {
  const size_t n = c->n;
  KrylovWs& K = ws(c);
  if (K.scal.n == 0) { K.scal.alloc(1024); K.partial.alloc((size_t)DOT_MAX * RED_BLOCKS); }
  ensure_vectors(c, K, true);
  for (int q = 1; q < nsys; ++q) ensure_vectors(c, c->kr_ion[q], true);
  if (c->kr_ones.n != n) {
    c->kr_ones.alloc(n);
    std::vector<double> one(n, 1.0);
    h2d(c->kr_ones.p, one.data(), n * sizeof(double), cs(c));
  }
  c->comm.batch_cap = c->P.N - 1 > 1 ? c->P.N - 1 : 1;
  double ng = (double)c->n_own;
  global_sum(c, &ng, 1);
  c->n_global = ng;
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
  const Batch PB = batch1(c->amg_emi, B, c->bj_emi.p);          // the preconditioner is built from B_emi
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
  precondition(c, PB, vecs1(b), vecs1(z));
  remove_mean(c, z);
  const double bnorm = sqrt(dot_host(c, z, z));
  const double tol = fmax(rtol * bnorm, atol);
  bell_spmv(c, A, x, b, r, 1);
  remove_mean(c, r);
  precondition(c, PB, vecs1(r), vecs1(z));
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
      precondition(c, PB, vecs1(r), vecs1(z));
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
// preconditioners of the ions' systems after a re-assembly (the reference rebuilds BoomerAMG at
// every setOperators, solver.py:767); the lagged-refresh bookkeeping of the block system is kept
// in the first ion's AmgValues
static bool knp_refresh(knp_ctx* c, int nion) {
  AmgValues& V0 = c->amg_knp[0];
  const bool due = refresh_due(c, V0);
  if (!due) return false;
  for (int ion = 0; ion < nion; ++ion) {
    BellMat A = bell_of(c, 2 + ion);
    if (c->opt.pc == 1 && c->amg.ready) amg_refresh(c, c->amg_knp[ion], A, c->A_knp[ion].p, c->A_knp[ion].p);
    else { if (c->bj_knp[ion].n != (size_t)c->slot_stride()) c->bj_knp[ion].alloc(c->slot_stride()); block_inverse(c, c->A_knp[ion].p, c->bj_knp[ion].p); }
  }
  return true;
}

// Left-preconditioned restarted GMRES on the block-diagonal system of all solved ions (the
// reference's mixed-space system, solver.py:684-701, 767-771): vectors are nb blocks, inner products
// sum over the blocks, convergence is tested on the norm of the whole preconditioned residual.
static int gmres_block(knp_ctx* c, int nb, double rtol, double atol, int maxit, double* resid_out) {
  knp_stream_t s = cs(c);
  const int64_t n = c->n, no = c->n_own;   // stride of the Krylov basis / owned rows
  const int m = c->opt.restart;
  const bool refreshed = knp_refresh(c, nb);
  Batch B{};
  B.nb = nb;
  BellMat A[MAX_BATCH];
  Vecs x{}, b{}, V{}, w{}, r{};
  for (int q = 0; q < nb; ++q) {
    A[q] = bell_of(c, 2 + q);
    B.s[q] = Sys{&c->amg_knp[q], A[q], c->bj_knp[q].p};
    KrylovWs& K = q == 0 ? c->kr0 : c->kr_ion[q];
    x.p[q] = c->c[q].p; b.p[q] = c->rhs_knp[q].p;
    V.p[q] = K.V.p; w.p[q] = K.w.p; r.p[q] = K.r.p;
  }
  const Vecs none{};
  double* hdev = ws(c).scal.p + 512;  // device copy of the current Hessenberg column / y
  const bool pre0 = c->opt.knp_presmooth0;
  precondition(c, B, b, w, pre0);
  const double bnorm = sqrt(dot_host(c, nb, w, w));
  const double tol = fmax(rtol * bnorm, atol);
  std::vector<double> H((size_t)(m + 1) * m), cs_(m), sn(m), g(m + 1), y(m), hcol(m + 2);
  int it = 0;
  double res = 0.0;
  while (true) {
    bell_spmv<double>(c, nb, A, x, b, r, 1);
    precondition(c, B, r, V, pre0);        // V0 = M^-1 (b - A x)
    const double beta = sqrt(dot_host(c, nb, V, V));
    res = beta;
    if ((beta <= tol && it >= c->opt.knp_min_it) || it >= maxit || beta == 0.0) break;
    { BatchOf<ScaleKernel> k{}; for (int q = 0; q < nb; ++q) k.f[q] = ScaleKernel{1.0 / beta, V.p[q], V.p[q]};
      parallel_for_batch(s, no, nb, k); }
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int j = 0;
    bool done = false;
    for (; j < m && it < maxit; ++j) {
      const Vecs vj = offset(V, nb, (int64_t)j * n), vn = offset(V, nb, (int64_t)(j + 1) * n);
      bell_spmv<double>(c, nb, A, vj, none, r, 0);
      precondition(c, B, r, vn, pre0);     // w = M^-1 A v_j, built in the next basis slot
      // classical Gram-Schmidt with ONE reduction per step: h = V^T w and |w|^2 in the same
      // pass (w is basis slot j+1), the new norm from Pythagoras, update + normalisation fused;
      // h stays on the device for the update, the host reads it once for the Givens rotations
      {
        DotBatch D;
        D.nb = nb;
        for (int q = 0; q < nb; ++q) { D.V[q] = V.p[q]; D.w[q] = vn.p[q]; }
        multi_dot_batch(s, no, n, j + 2, D, ws(c).partial.p, hdev);
      }
      c->comm.allreduce(s, hdev, j + 2);
      // the Pythagorean norm carries a relative error of about eps |w|^2 / hn^2: keep it below
      // the requested tolerance, otherwise fall back to the explicit norm
      const double cancel_tol = fmax(1e-6, 100.0 * 2.2e-16 / fmax(rtol, 1e-16));
      { BatchOf<GsNormalizeKernel> k{}; for (int q = 0; q < nb; ++q) k.f[q] = GsNormalizeKernel{n, j + 1, V.p[q], hdev, vn.p[q], cancel_tol};
        parallel_for_batch(s, no, nb, k); }
      d2h(hcol.data(), hdev, (j + 2) * sizeof(double), s);
      double hsum = 0.0;
      for (int i = 0; i <= j; ++i) hsum += hcol[i] * hcol[i];
      double hn2 = hcol[j + 1] - hsum;
      if (!(hn2 >= cancel_tol * hcol[j + 1] && hn2 > 0.0)) {
        // too much cancellation (w almost in span V): explicit norm of the updated vector
        hn2 = dot_host(c, nb, vn, vn);
        if (hn2 > 0.0) {
          BatchOf<ScaleKernel> k{};
          for (int q = 0; q < nb; ++q) k.f[q] = ScaleKernel{1.0 / sqrt(hn2), vn.p[q], vn.p[q]};
          parallel_for_batch(s, no, nb, k);
        }
      }
      const double hn = hn2 > 0.0 ? sqrt(hn2) : 0.0;
      for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = hcol[i];
      H[(size_t)(j + 1) * m + j] = hn;
      for (int i = 0; i < j; ++i) {                // apply previous rotations
        const double a = H[(size_t)i * m + j], bq = H[(size_t)(i + 1) * m + j];
        H[(size_t)i * m + j] = cs_[i] * a + sn[i] * bq;
        H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs_[i] * bq;
      }
      const double a = H[(size_t)j * m + j], bq = H[(size_t)(j + 1) * m + j];
      const double den = hypot(a, bq);
      cs_[j] = den > 0 ? a / den : 1.0; sn[j] = den > 0 ? bq / den : 0.0;
      H[(size_t)j * m + j] = den; H[(size_t)(j + 1) * m + j] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs_[j] * g[j];
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
      BatchOf<CombineKernel> ck{};
      for (int q = 0; q < nb; ++q) ck.f[q] = CombineKernel{n, k, V.p[q], hdev, x.p[q]};
      parallel_for_batch(s, no, nb, ck);
    }
    if (done || it >= maxit) {
      if (!done) { *resid_out = res; return -it; }
      break;
    }
  }
  *resid_out = res;
  note_solve(c->amg_knp[0], refreshed, it);
  return it;
}

extern "C" int knp_solve_knp(knp_ctx* ctx, double rtol, double atol, int maxit, int* niter, double* resid) {
  KNP_TRY
  if (!ctx->knp_assembled) fail("knp_solve_knp: assemble first");
  stream_sync(ctx->stream);
  const double t0 = now_s();
  const int nion = ctx->P.N - 1;
  if (nion > MAX_BATCH) fail("knp_solve_knp: too many ions");
  ensure_krylov(ctx, nion);
  double res = 0.0;
  const int it = gmres_block(ctx, nion, rtol, atol, maxit, &res);
  if (it < 0) fail("knp_solve_knp: GMRES did not converge in " + std::to_string(maxit) +
                   " iterations (ksp_error_if_not_converged, solver.py:428)");
  {
    Vecs x{};
    for (int q = 0; q < nion; ++q) x.p[q] = ctx->c[q].p;
    halo0(ctx, nion, x);   // post-step and the next assembly read c on the ghost cells
  }
  stream_sync(ctx->stream);
#ifndef KNP_EMU
  ctx->comm.check_p2p(ctx->stream);
#endif
  if (niter) *niter = it;
  if (resid) *resid = res;
  ctx->timers[T_KNP_SOLVE] += now_s() - t0;
  KNP_CATCH
}

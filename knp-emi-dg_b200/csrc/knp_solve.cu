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

// ---- small helpers ---------------------------------------------------------------------
template <int ND>
static void bell_spmv(knp_ctx* c, const BellMat& A, const double* x, const double* b, double* y, int mode) {
  BellSpmvKernel<ND> k{A, x, b, y, mode};
  parallel_for(c->stream, c->n, k, 256);
}
static void bell_spmv(knp_ctx* c, const BellMat& A, const double* x, const double* b, double* y, int mode) {
  if (c->nd == 3) bell_spmv<3>(c, A, x, b, y, mode); else bell_spmv<4>(c, A, x, b, y, mode);
}
static void block_apply(knp_ctx* c, const double* dinv, const double* r, double* out, double w, int mode) {
  if (c->nd == 3) { BlockDiagApplyKernel<3> k{dinv, r, out, w, mode}; parallel_for(c->stream, c->n, k); }
  else { BlockDiagApplyKernel<4> k{dinv, r, out, w, mode}; parallel_for(c->stream, c->n, k); }
}
static void bell_jacobi(knp_ctx* c, const BellMat& A, const double* dinv, const double* b,
                        const double* xin, double* xout, double w) {
  if (c->nd == 3) { BellJacobiKernel<3> k{A, dinv, b, xin, xout, w}; parallel_for(c->stream, c->n, k, 192); }
  else { BellJacobiKernel<4> k{A, dinv, b, xin, xout, w}; parallel_for(c->stream, c->n, k, 256); }
}
static void block_inverse(knp_ctx* c, const double* blocks, double* inv) {
  if (c->nd == 3) { BlockInverseKernel<3> k{blocks, inv}; parallel_for(c->stream, c->nc, k, 128); }
  else { BlockInverseKernel<4> k{blocks, inv}; parallel_for(c->stream, c->nc, k, 128); }
}

// host-visible dot products (one sync each)
static void dots_host(knp_ctx* c, int k, const double* V, const double* w, double* out_host) {
  multi_dot_device(c->stream, c->n, k, V, w, c->kr_partial.p, c->kr_scal.p);
  d2h(out_host, c->kr_scal.p, k * sizeof(double), c->stream);
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
  A.n = c->n;
  A.ptr.assign(A.n + 1, 0);
  A.col.reserve((size_t)c->nnz_export); A.pos.reserve((size_t)c->nnz_export);
  for (int64_t cell = 0; cell < nc; ++cell)
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
static int64_t vertex_injection(knp_ctx* c, std::vector<int32_t>& agg) {
  const int nd = c->nd;
  const int64_t nc = c->nc, n = c->n;
  std::vector<int32_t> parent(n);
  std::iota(parent.begin(), parent.end(), 0);
  auto find = [&](int32_t x) {
    while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
    return x;
  };
  for (int64_t cell = 0; cell < nc; ++cell)
    for (int f = 0; f < nd; ++f) {
      const int w = c->h_finfo[(size_t)f * nc + cell];
      if (fi_kind(w) != FK_SIP) continue;
      const int32_t c2 = c->h_nbr[(size_t)f * nc + cell];
      if (c2 < cell) continue;
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

static void upload_level(knp_ctx* c, AmgLevelPlan& L, const GalerkinPlan& G, const HostTransfer& T) {
  knp_stream_t s = c->stream;
  L.n = G.coarse.n; L.nnz = (int64_t)G.coarse.col.size();
  L.ptr.upload(G.coarse.ptr, s); L.col.upload(G.coarse.col, s);
  L.gptr.upload(G.gptr, s); L.gidx.upload(G.gidx, s);
  L.g_unit = G.gw.empty();
  if (!L.g_unit) L.gw.upload(G.gw, s);
  L.pptr.upload(T.pptr, s); L.pidx.upload(T.pidx, s);
  L.rptr.upload(T.rptr, s); L.ridx.upload(T.ridx, s);
  L.t_unit = T.unit;
  if (!T.unit) { L.pw.upload(T.pw, s); L.rw.upload(T.rw, s); }
  L.b.alloc(L.n); L.x.alloc(L.n); L.r.alloc(L.n); L.t.alloc(L.n);
}

static void alloc_values(knp_ctx* c, AmgValues& V) {
  const size_t nl = c->amg.lev.size();
  V.val.resize(nl); V.dinv.resize(nl);
  for (size_t l = 0; l < nl; ++l) { V.val[l].alloc(c->amg.lev[l].nnz); V.dinv[l].alloc(c->amg.lev[l].n); }
  V.dense.alloc((size_t)c->amg.m_dense * c->amg.m_dense);
  V.binv.alloc((size_t)c->slot_stride());
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
  std::vector<int32_t> agg;
  int64_t ncoarse = vertex_injection(ctx, agg);
  HostTransfer T = transfer_from_aggregates(agg, ncoarse);
  GalerkinPlan G = galerkin_plan(A0, T);
  galerkin_numeric_host(G, fine);
  amg.lev.emplace_back();
  upload_level(ctx, amg.lev.back(), G, T);
  { std::vector<int32_t>().swap(A0.col); std::vector<int32_t>().swap(A0.pos); }
  while ((int)amg.lev.size() + 1 < max_levels && G.coarse.n > coarse_size) {
    std::vector<int32_t> ag2;
    const int64_t na = aggregate(G.coarse, theta, ag2);
    if (na >= G.coarse.n * 0.9 || na < 1) break;  // coarsening stalled
    HostTransfer T2 = transfer_from_aggregates(ag2, na);
    std::vector<double> vals = G.coarse.val;
    GalerkinPlan G2 = galerkin_plan(G.coarse, T2);
    galerkin_numeric_host(G2, vals);
    amg.lev.emplace_back();
    upload_level(ctx, amg.lev.back(), G2, T2);
    G = std::move(G2);
  }
  amg.m_dense = amg.lev.back().n;
  if (amg.m_dense > 4096) fail("knp_amg_setup: coarsest level too large for the dense solve (" +
                                std::to_string(amg.m_dense) + " rows); raise max_levels");
  amg.x0.alloc(ctx->n); amg.r0.alloc(ctx->n); amg.t0.alloc(ctx->n);
  amg.colbuf.alloc(amg.m_dense);
  alloc_values(ctx, ctx->amg_emi);
  for (int k = 0; k < ctx->P.N - 1; ++k) alloc_values(ctx, ctx->amg_knp[k]);
  amg.ready = true;
  KNP_CATCH
}

extern "C" int knp_amg_info(knp_ctx* ctx, int64_t* nlevels, int64_t* rows, int64_t* nnz, int cap) {
  KNP_TRY
  if (!ctx->amg.ready) fail("AMG not set up");
  *nlevels = (int64_t)ctx->amg.lev.size() + 1;
  if (cap > 0) { rows[0] = ctx->n; nnz[0] = ctx->nnz_export; }
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
  ctx->amg_emi.omega = 0.0;
  for (int k = 0; k < MAX_IONS; ++k) ctx->amg_knp[k].omega = 0.0;
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
static double estimate_lambda_max(knp_ctx* c, const BellMat& A, const double* dinv) {
  const int64_t n = c->n;
  double* v = c->amg.x0.p; double* w = c->amg.t0.p; double* u = c->amg.r0.p;
  std::vector<double> h(n);
  for (int64_t i = 0; i < n; ++i) h[i] = 1.0 + 0.5 * sin(1.7 * (double)i) + ((i * 2654435761u) % 1024) / 1024.0;
  h2d(v, h.data(), n * sizeof(double), c->stream);
  double lam = 1.0;
  for (int it = 0; it < 12; ++it) {
    bell_spmv(c, A, v, nullptr, u, 0);
    block_apply(c, dinv, u, w, 1.0, 0);
    const double nv = sqrt(dot_host(c, v, v)), nw = sqrt(dot_host(c, w, w));
    if (!(nv > 0.0) || !(nw > 0.0)) break;
    lam = nw / nv;
    ScaleKernel k{1.0 / nw, w, v};
    parallel_for(c->stream, n, k);
  }
  return lam;
}

static void amg_refresh(knp_ctx* c, AmgValues& V, const BellMat& A0, const double* fine_values, const double* diag_blocks) {
  knp_stream_t s = c->stream;
  block_inverse(c, diag_blocks, V.binv.p);
  if (c->opt.omega > 0.0) V.omega = c->opt.omega;
  else if (V.omega <= 0.0 || ++V.age >= OMEGA_PERIOD) {
    V.omega = 4.0 / (3.0 * 1.05 * estimate_lambda_max(c, A0, V.binv.p));
    V.age = 0;
  }
  const double* fine = fine_values;
  for (size_t l = 0; l < c->amg.lev.size(); ++l) {
    AmgLevelPlan& L = c->amg.lev[l];
    GalerkinKernel g{L.gptr.p, L.gidx.p, L.g_unit ? nullptr : L.gw.p, fine, V.val[l].p};
    parallel_for(s, L.nnz, g);
    CsrL1DiagKernel dk{csr_of(L, V, l), V.dinv[l].p};
    parallel_for(s, L.n, dk);
    fine = V.val[l].p;
  }
  const size_t last = c->amg.lev.size() - 1;
  const int64_t m = c->amg.m_dense;
  dev_zero(V.dense.p, (size_t)m * m * sizeof(double), s);
  CsrToDenseKernel tk{csr_of(c->amg.lev[last], V, last), V.dense.p};
  parallel_for(s, m, tk);
  dense_inverse_device(s, (int)m, V.dense.p, c->amg.colbuf.p);
}

// ---------------------------------------------------------------------------------
// cycle
// ---------------------------------------------------------------------------------
static void transfer(knp_ctx* c, int64_t nrows, const int32_t* ptr, const int32_t* idx, const double* w,
                     const double* x, double* y, int add) {
  TransferKernel k{nrows, ptr, idx, w, x, y, add};
  parallel_for(c->stream, nrows, k);
}

// solve level l (>= 1, index into lev = l-1) approximately: L.x <- cycle(L.b)
static void coarse_cycle(knp_ctx* c, AmgValues& V, size_t li) {
  knp_stream_t s = c->stream;
  AmgLevelPlan& L = c->amg.lev[li];
  if (li + 1 == c->amg.lev.size()) {
    DenseMatvecKernel k{L.n, V.dense.p, L.b.p, L.x.p};
    parallel_for(s, L.n, k, 64);
    return;
  }
  CsrMat A = csr_of(L, V, li);
  AmgLevelPlan& C = c->amg.lev[li + 1];
  if (c->opt.nu_pre == 1 && c->opt.nu_post == 1 && c->opt.gamma == 1 && C.t_unit) {
    // fused V(1,1) path: two kernels down (smooth+residual, restrict), one up (prolong+smooth)
    { CoarseResidualKernel k{A, V.dinv[li].p, L.b.p, L.x.p, L.r.p}; parallel_rows<8>(s, L.n, k); }
    { TransferRowsKernel k{C.rptr.p, C.ridx.p, nullptr, L.r.p, C.b.p, 0}; parallel_rows<8>(s, C.n, k); }
    coarse_cycle(c, V, li + 1);
    { CoarseUpKernel k{A, V.dinv[li].p, L.b.p, L.x.p, C.pidx.p, C.x.p, L.t.p}; parallel_rows<8>(s, L.n, k); }
    std::swap(L.x.p, L.t.p);
    return;
  }
  // pre-smoothing from a zero guess
  { DiagScaleKernel k{V.dinv[li].p, L.b.p, L.x.p, 1.0}; parallel_for(s, L.n, k); }
  for (int it = 1; it < c->opt.nu_pre; ++it) {
    CsrJacobiKernel k{A, V.dinv[li].p, L.b.p, L.x.p, L.t.p, 1.0};
    parallel_for(s, L.n, k);
    std::swap(L.x.p, L.t.p);
  }
  for (int g = 0; g < c->opt.gamma; ++g) {
    { CsrSpmvKernel k{A, L.x.p, L.b.p, L.r.p, 1}; parallel_for(s, L.n, k); }
    transfer(c, C.n, C.rptr.p, C.ridx.p, C.t_unit ? nullptr : C.rw.p, L.r.p, C.b.p, 0);
    coarse_cycle(c, V, li + 1);
    transfer(c, L.n, C.pptr.p, C.pidx.p, C.t_unit ? nullptr : C.pw.p, C.x.p, L.x.p, 1);
  }
  for (int it = 0; it < c->opt.nu_post; ++it) {
    CsrJacobiKernel k{A, V.dinv[li].p, L.b.p, L.x.p, L.t.p, 1.0};
    parallel_for(s, L.n, k);
    std::swap(L.x.p, L.t.p);
  }
}

// z = M^-1 r
static void precondition(knp_ctx* c, AmgValues& V, const BellMat& A0, const double* bj, const double* r, double* z) {
  if (c->opt.pc == 0 || !c->amg.ready) {
    block_apply(c, bj, r, z, 1.0, 0);
    return;
  }
  AmgPlan& amg = c->amg;
  AmgLevelPlan& C = amg.lev[0];
  const double w = V.omega;
  double* x = amg.x0.p; double* t = amg.t0.p;
  block_apply(c, V.binv.p, r, x, w, 0);
  for (int it = 1; it < c->opt.nu_pre; ++it) { bell_jacobi(c, A0, V.binv.p, r, x, t, w); std::swap(x, t); }
  bell_spmv(c, A0, x, r, amg.r0.p, 1);
  { TransferRowsKernel k{C.rptr.p, C.ridx.p, C.t_unit ? nullptr : C.rw.p, amg.r0.p, C.b.p, 0}; parallel_rows<8>(c->stream, C.n, k); }
  coarse_cycle(c, V, 0);
  transfer(c, c->n, C.pptr.p, C.pidx.p, C.t_unit ? nullptr : C.pw.p, C.x.p, x, 1);
  if (c->opt.nu_post == 0) { d2d(z, x, c->n * sizeof(double), c->stream); return; }
  for (int it = 0; it < c->opt.nu_post; ++it) {
    double* out = (it + 1 == c->opt.nu_post) ? z : t;
    bell_jacobi(c, A0, V.binv.p, r, x, out, w);
    if (out != z) std::swap(x, t);
  }
  // keep the plan's buffers in their slots for the next call
  if (x != amg.x0.p) std::swap(amg.x0.p, amg.t0.p);
}

// ---------------------------------------------------------------------------------
// CG (EMI)
// ---------------------------------------------------------------------------------
// A_emi is singular (pure Neumann: constants, solver.py:465-466).  The preconditioned
// residual is kept orthogonal to the constants, otherwise the round-off component of r
// along them, amplified by the mass-shifted preconditioner, puts a floor under ||M^-1 r||.
static void remove_mean(knp_ctx* c, double* v) {
  const double s = dot_host(c, v, c->kr_ones.p);
  AddConstKernel k{-s / (double)c->n, v};
  parallel_for(c->stream, c->n, k);
}

static void ensure_krylov(knp_ctx* c) {
  const size_t n = c->n;
  if (c->kr_r.n != 2 * n) { c->kr_r.alloc(2 * n); c->kr_p.alloc(n); c->kr_q.alloc(n); c->kr_w.alloc(n); }
  if (c->kr_ones.n != n) {
    c->kr_ones.alloc(n);
    std::vector<double> one(n, 1.0);
    h2d(c->kr_ones.p, one.data(), n * sizeof(double), c->stream);
  }
  const size_t need = (size_t)(c->opt.restart + 1) * n;
  if (c->kr_V.n < need) c->kr_V.alloc(need);
}

extern "C" int knp_solve_emi(knp_ctx* ctx, double rtol, double atol, int maxit, int* niter, double* resid) {
  KNP_TRY
  if (!ctx->emi_assembled) fail("knp_solve_emi: assemble first");
  knp_ctx* c = ctx;
  knp_stream_t s = c->stream;
  stream_sync(s);
  const double t0 = now_s();
  ensure_krylov(c);
  const int64_t n = c->n;
  BellMat A = bell_of(c, 0), B = bell_of(c, 1);
  // preconditioner refresh (the reference rebuilds BoomerAMG at every setOperators)
  if (c->opt.pc == 1 && c->amg.ready) amg_refresh(c, c->amg_emi, B, c->A_emi.p, c->Bdiag());
  else { if (c->bj_emi.n != (size_t)c->slot_stride()) c->bj_emi.alloc(c->slot_stride()); block_inverse(c, c->Bdiag(), c->bj_emi.p); }
  double* x = c->phi.p; double* r = c->kr_r.p; double* z = c->kr_r.p + n; double* p = c->kr_p.p; double* q = c->kr_q.p;
  const double* b = c->rhs_emi.p;
  // reference norm ||M^-1 b||
  precondition(c, c->amg_emi, B, c->bj_emi.p, b, z);
  remove_mean(c, z);
  const double bnorm = sqrt(dot_host(c, z, z));
  const double tol = fmax(rtol * bnorm, atol);
  bell_spmv(c, A, x, b, r, 1);
  precondition(c, c->amg_emi, B, c->bj_emi.p, r, z);
  remove_mean(c, z);
  double zn = sqrt(dot_host(c, z, z));
  int it = 0, best_it = 0;
  double best = zn;
  if (zn > tol) {
    d2d(p, z, n * sizeof(double), s);
    double rz = dot_host(c, r, z);
    for (it = 1; it <= maxit; ++it) {
      bell_spmv(c, A, p, nullptr, q, 0);
      const double pq = dot_host(c, p, q);
      if (!(pq > 0.0)) {
        // at round-off level p can fall into the (constant) null space of A: the iteration
        // has converged as far as fp64 allows
        if (pq == pq && zn <= 1e-8 * bnorm) { --it; break; }
        if (pq == 0.0 || pq != pq) fail("knp_solve_emi: CG breakdown (p.Ap = " + std::to_string(pq) + ")");
        fail("knp_solve_emi: operator or preconditioner is indefinite");
      }
      const double alpha = rz / pq;
      { Axpy2Kernel k{alpha, p, q, x, r}; parallel_for(s, n, k); }
      precondition(c, c->amg_emi, B, c->bj_emi.p, r, z);
      remove_mean(c, z);
      double d2[2];
      dots_host(c, 2, r, z, d2);     // {r, z} are contiguous: r.z and z.z in one pass
      zn = sqrt(d2[1]);
      if (zn <= tol) break;
      // attainable accuracy: a tolerance below the fp64 floor of this system is treated as
      // reached once the preconditioned residual has stagnated at round-off level
      if (zn < best) { best = zn; best_it = it; }
      else if (it - best_it >= 40 && best <= 1e-10 * bnorm) break;
      const double beta = d2[0] / rz;
      rz = d2[0];
      { AxpbyKernel k{1.0, z, beta, p}; parallel_for(s, n, k); }
    }
    if (it > maxit) fail("knp_solve_emi: CG did not converge in " + std::to_string(maxit) +
                         " iterations (ksp_error_if_not_converged, solver.py:428)");
  }
  stream_sync(s);
  if (niter) *niter = it;
  if (resid) *resid = zn;
  c->timers[T_EMI_SOLVE] += now_s() - t0;
  KNP_CATCH
}

// ---------------------------------------------------------------------------------
// GMRES (KNP), one ion at a time
// ---------------------------------------------------------------------------------
static int gmres_one(knp_ctx* c, int ion, double rtol, double atol, int maxit, double* resid_out) {
  knp_stream_t s = c->stream;
  const int64_t n = c->n;
  const int m = c->opt.restart;
  BellMat A = bell_of(c, 2 + ion);
  AmgValues& Vv = c->amg_knp[ion];
  if (c->opt.pc == 1 && c->amg.ready) amg_refresh(c, Vv, A, c->A_knp[ion].p, c->A_knp[ion].p);
  else { if (c->bj_knp[ion].n != (size_t)c->slot_stride()) c->bj_knp[ion].alloc(c->slot_stride()); block_inverse(c, c->A_knp[ion].p, c->bj_knp[ion].p); }
  const double* bj = c->bj_knp[ion].p;
  double* x = c->c[ion].p;
  const double* b = c->rhs_knp[ion].p;
  double* V = c->kr_V.p; double* w = c->kr_w.p; double* r = c->kr_r.p;
  double* hdev = c->kr_scal.p + 512;  // device copy of the current Hessenberg column / y
  precondition(c, Vv, A, bj, b, w);
  const double bnorm = sqrt(dot_host(c, w, w));
  const double tol = fmax(rtol * bnorm, atol);
  std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), y(m), hcol(m + 2);
  int it = 0;
  double res = 0.0;
  while (true) {
    bell_spmv(c, A, x, b, r, 1);
    precondition(c, Vv, A, bj, r, V);              // V0 = M^-1 (b - A x)
    const double beta = sqrt(dot_host(c, V, V));
    res = beta;
    if ((beta <= tol && it >= c->opt.knp_min_it) || it >= maxit || beta == 0.0) break;
    { ScaleKernel k{1.0 / beta, V, V}; parallel_for(s, n, k); }
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int j = 0;
    bool done = false;
    for (; j < m && it < maxit; ++j) {
      double* vj = V + (int64_t)j * n;
      double* vn = V + (int64_t)(j + 1) * n;
      bell_spmv(c, A, vj, nullptr, r, 0);
      precondition(c, Vv, A, bj, r, w);            // w = M^-1 A v_j
      // classical Gram-Schmidt, one pass: h = V^T w stays on the device for the update,
      // ||w||^2 lands right behind it; one host read per iteration
      multi_dot_device(s, n, j + 1, V, w, c->kr_partial.p, hdev);
      { GsUpdateKernel k{n, j + 1, V, hdev, w}; parallel_for(s, n, k); }
      multi_dot_device(s, n, 1, w, w, c->kr_partial.p, hdev + j + 1);
      d2h(hcol.data(), hdev, (j + 2) * sizeof(double), s);
      const double hn = sqrt(hcol[j + 1]);
      for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = hcol[i];
      H[(size_t)(j + 1) * m + j] = hn;
      if (hn > 0.0) { ScaleKernel k{1.0 / hn, w, vn}; parallel_for(s, n, k); }
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
      parallel_for(s, n, ck);
    }
    if (done || it >= maxit) {
      if (!done) { *resid_out = res; return -it; }
      break;
    }
  }
  *resid_out = res;
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
  for (int ion = 0; ion < ctx->P.N - 1; ++ion) {
    double res = 0.0;
    const int it = gmres_one(ctx, ion, rtol, atol, maxit, &res);
    if (it < 0) fail("knp_solve_knp: GMRES did not converge for ion " + std::to_string(ion));
    worst = it > worst ? it : worst;
    rmax = res > rmax ? res : rmax;
  }
  stream_sync(ctx->stream);
  if (niter) *niter = worst;
  if (resid) *resid = rmax;
  ctx->timers[T_KNP_SOLVE] += now_s() - t0;
  KNP_CATCH
}

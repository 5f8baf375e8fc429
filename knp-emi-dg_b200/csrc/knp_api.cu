// knp_api.cu - C ABI: context, mesh tables, parameters, fields, assembly, post-step,
// membrane ODE step.  The linear solvers live in knp_solve.cu.
#include "../../include/knpemi.h"
#include <algorithm>
#include "knp_ctx.h"
#ifdef KNP_MODELS_HEADER      // a library variant that also carries user-supplied membrane models (build.py)
#include KNP_MODELS_HEADER
#else
#include "generated/models_gen.h"
#endif

using namespace knp;

namespace knp {
thread_local std::string g_last_error;
int set_error(const std::string& s) { g_last_error = s; return 1; }
}  // namespace knp

#define KNP_TRY try {
#define KNP_CATCH                                             \
  }                                                           \
  catch (const std::exception& e) { return knp::set_error(e.what()); } \
  catch (...) { return knp::set_error("unknown error"); }     \
  return 0;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct PhaseTimer {  // wall time of a phase, device work included (sync on both sides)
  knp_ctx* c; int id; double t0;
  PhaseTimer(knp_ctx* ctx, int which) : c(ctx), id(which) { stream_sync(c->stream); t0 = now_s(); }
  ~PhaseTimer() {
    try { stream_sync(c->stream); } catch (...) {}
    c->timers[id] += now_s() - t0;
  }
};

const char* knp_last_error(void) { return g_last_error.c_str(); }
int knp_version(void) { return 100; }
int knp_is_cuda_build(void) {
#ifdef KNP_EMU
  return 0;
#else
  return 1;
#endif
}

int knp_ctx_create(int device, knp_ctx** out) {
  KNP_TRY
  if (!out) fail("knp_ctx_create: out is NULL");
#ifndef KNP_EMU
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    fail("knp_ctx_create: no CUDA device available (libknpemi has no CPU fallback)");
  if (device < 0 || device >= count) fail("knp_ctx_create: bad device index");
  KNP_CUDA(cudaSetDevice(device));
#endif
  std::unique_ptr<knp_ctx> c(new knp_ctx());
  c->device = device;
#ifndef KNP_EMU
  KNP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
#endif
  { const char* e = getenv("KNP_KNP_PRESMOOTH"); c->opt.knp_presmooth0 = (e && e[0] == '1'); }
  { const char* e = getenv("KNP_AMG_FP32"); c->opt.pc_fp32 = !(e && e[0] == '0'); }
  { const char* e = getenv("KNP_AMG_CHEBY"); c->opt.cheby = (e && e[0] == '2') ? 2 : 1; }
  { const char* e = getenv("KNP_EXTRAPOLATE"); c->opt.extrapolate_phi = !(e && e[0] == '0'); }
  { const char* e = getenv("KNP_AMG_REFRESH_PERIOD"); c->opt.refresh_period = e ? std::max(1, atoi(e)) : 8; }
  c->kr0.stream = c->stream;
  c->kr0.scal.alloc(1024);
  c->kr0.partial.alloc((size_t)DOT_MAX * RED_BLOCKS);
  c->ode_stats.alloc(4);
  *out = c.release();
  KNP_CATCH
}

int knp_ctx_destroy(knp_ctx* ctx) {
  KNP_TRY
  if (!ctx) return 0;
#ifndef KNP_EMU
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->comm.close_p2p();
  if (ctx->comm.nccl) { nccl_api().CommDestroy(ctx->comm.nccl); ctx->comm.nccl = nullptr; }
#endif
  knp_stream_t s = ctx->stream;
#ifndef KNP_EMU
  for (auto& k : ctx->kr_ion) if (k.own_stream) { cudaStreamSynchronize(k.stream); cudaStreamDestroy(k.stream); }
  if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); cudaStreamDestroy(ctx->comm_stream); }
  if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
  if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
#endif
  delete ctx;
#ifndef KNP_EMU
  cudaStreamDestroy(s);
#else
  (void)s;
#endif
  KNP_CATCH
}

int knp_sync(knp_ctx* ctx) {
  KNP_TRY
  stream_sync(ctx->stream);
  KNP_CATCH
}

// ---------------------------------------------------------------------------------
// mesh
// ---------------------------------------------------------------------------------
template <int D>
static void build_mesh(knp_ctx* c, int64_t nc, int64_t nv, const double* coords,
                       const int32_t* cells, const int32_t* region, int64_t nf,
                       const int32_t* fcells, const int32_t* ftag, int nmt, const int32_t* mtags) {
  constexpr int ND = D + 1;
  c->d = D; c->nd = ND; c->nc = nc; c->n = nc * ND;
  c->nc_own = nc; c->n_own = c->n; c->n_global = (double)c->n;   // single part until knp_dist_set
  c->comm.world = 1; c->comm.rank = 0; c->comm.nbr.clear();
  c->halo0 = HaloPlan();
  c->halo0.n_own = c->n;
  if (nc * ND * ND * (ND + 2) >= (int64_t)2147483647) fail("mesh too large for 32-bit value positions");
  // geometry: grad lambda_i, |K|, h = max edge (CellDiameter, solver.py:102-103)
  std::vector<double> grad((size_t)nc * ND * D), vol(nc), hh(nc);
  for (int64_t k = 0; k < nc; ++k) {
    double X[ND][D];
    for (int a = 0; a < ND; ++a) {
      const int32_t v = cells[k * ND + a];
      if (v < 0 || v >= nv) fail("cell vertex index out of range");
      for (int x = 0; x < D; ++x) X[a][x] = coords[(int64_t)v * D + x];
    }
    double T[D][D], Ti[D][D];  // T[x][e] = X[e+1][x] - X[0][x]
    for (int x = 0; x < D; ++x)
      for (int e = 0; e < D; ++e) T[x][e] = X[e + 1][x] - X[0][x];
    double det;
    if constexpr (D == 2) {
      det = T[0][0] * T[1][1] - T[0][1] * T[1][0];
      Ti[0][0] = T[1][1] / det; Ti[0][1] = -T[0][1] / det;
      Ti[1][0] = -T[1][0] / det; Ti[1][1] = T[0][0] / det;
    } else {
      const double a = T[0][0], b = T[0][1], cc = T[0][2], dd = T[1][0], e = T[1][1],
                   f = T[1][2], g = T[2][0], hq = T[2][1], i = T[2][2];
      det = a * (e * i - f * hq) - b * (dd * i - f * g) + cc * (dd * hq - e * g);
      Ti[0][0] = (e * i - f * hq) / det; Ti[0][1] = (cc * hq - b * i) / det; Ti[0][2] = (b * f - cc * e) / det;
      Ti[1][0] = (f * g - dd * i) / det; Ti[1][1] = (a * i - cc * g) / det; Ti[1][2] = (cc * dd - a * f) / det;
      Ti[2][0] = (dd * hq - e * g) / det; Ti[2][1] = (b * g - a * hq) / det; Ti[2][2] = (a * e - b * dd) / det;
    }
    if (det == 0.0) fail("degenerate cell");
    vol[k] = fabs(det) / (D == 2 ? 2.0 : 6.0);
    // rows of T^-1 are grad lambda_1..D; stored component major [ND*D][nc] so that
    // consecutive cells (= consecutive lanes) read consecutive addresses
    double g0[D];
    for (int x = 0; x < D; ++x) g0[x] = 0.0;
    for (int e = 0; e < D; ++e)
      for (int x = 0; x < D; ++x) { grad[(size_t)((e + 1) * D + x) * nc + k] = Ti[e][x]; g0[x] -= Ti[e][x]; }
    for (int x = 0; x < D; ++x) grad[(size_t)x * nc + k] = g0[x];
    double hm = 0.0;
    for (int a = 0; a < ND; ++a)
      for (int b = a + 1; b < ND; ++b) {
        double s = 0.0;
        for (int x = 0; x < D; ++x) s += (X[a][x] - X[b][x]) * (X[a][x] - X[b][x]);
        hm = fmax(hm, sqrt(s));
      }
    hh[k] = hm;
  }
  // facet tables
  c->h_nbr.assign((size_t)ND * nc, -1);
  int none_word = FK_NONE;                       // identity column map for facets without terms
  for (int a = 0; a < ND; ++a) none_word |= a << (4 + 2 * a);
  c->h_finfo.assign((size_t)ND * nc, none_word);
  c->h_fmem.assign((size_t)ND * nc, -1);
  c->h_mem_facet.clear(); c->h_mem_ci.clear(); c->h_mem_ce.clear(); c->h_mem_tag.clear();
  c->h_mem_fi.clear();
  c->nsip = 0;
  int64_t nblocks = nc;  // diagonal blocks
  for (int64_t f = 0; f < nf; ++f) {
    const int64_t c0 = fcells[2 * f], c1 = fcells[2 * f + 1];
    if (c0 < 0 || c1 < 0) continue;  // exterior facets carry no terms (zero flux)
    if (c0 >= nc || c1 >= nc) fail("facet cell index out of range");
    const int tag = ftag[f];
    bool is_mem = false;
    for (int t = 0; t < nmt; ++t) is_mem |= (mtags[t] == tag);
    int kind;
    if (is_mem) kind = FK_MEMBRANE;
    else if (tag == 0) kind = FK_SIP;
    else continue;  // tagged facet without a membrane model: no terms
    int p0[ND], p1[ND], f0 = -1, f1 = -1;
    for (int a = 0; a < ND; ++a) { p0[a] = -1; p1[a] = -1; }
    for (int a = 0; a < ND; ++a)
      for (int b = 0; b < ND; ++b)
        if (cells[c0 * ND + a] == cells[c1 * ND + b]) { p0[a] = b; p1[b] = a; }
    int shared = 0;
    for (int a = 0; a < ND; ++a) { if (p0[a] < 0) f0 = a; else ++shared; }
    for (int b = 0; b < ND; ++b) if (p1[b] < 0) f1 = b;
    if (shared != D || f0 < 0 || f1 < 0) fail("facet_cells: the two cells do not share a facet");
    const bool c0_ics = region[c0] >= region[c1];  // n_g: lower -> higher tag (utils.py:80)
    int w0 = kind | (f1 << 2), w1 = kind | (f0 << 2);
    // column map: my vertex a -> neighbour's local index; my opposite vertex -> its opposite vertex
    for (int a = 0; a < ND; ++a) {
      w0 |= ((a != f0) ? p0[a] : f1) << (4 + 2 * a);
      w1 |= ((a != f1) ? p1[a] : f0) << (4 + 2 * a);
    }
    if (kind == FK_MEMBRANE) {
      if (c0_ics) w0 |= 1 << 12; else w1 |= 1 << 12;
      const int32_t m = (int32_t)c->h_mem_facet.size();
      c->h_fmem[(size_t)f0 * nc + c0] = m;
      c->h_fmem[(size_t)f1 * nc + c1] = m;
      c->h_mem_facet.push_back((int32_t)f);
      c->h_mem_ci.push_back((int32_t)(c0_ics ? c0 : c1));
      c->h_mem_ce.push_back((int32_t)(c0_ics ? c1 : c0));
      c->h_mem_fi.push_back(c0_ics ? f0 : f1);
      c->h_mem_tag.push_back(tag);
    } else {
      c->nsip++;
    }
    if (c->h_nbr[(size_t)f0 * nc + c0] >= 0 || c->h_nbr[(size_t)f1 * nc + c1] >= 0)
      fail("two facets claim the same cell side");
    c->h_nbr[(size_t)f0 * nc + c0] = (int32_t)c1; c->h_finfo[(size_t)f0 * nc + c0] = w0;
    c->h_nbr[(size_t)f1 * nc + c1] = (int32_t)c0; c->h_finfo[(size_t)f1 * nc + c1] = w1;
    nblocks += 2;
  }
  c->nm = (int64_t)c->h_mem_facet.size();
  c->nnz_export = nblocks * ND * ND;
  knp_stream_t s = c->stream;
  c->grad.upload(grad, s); c->vol.upload(vol, s); c->h.upload(hh, s);
  c->region.upload(region, nc, s);
  c->nbr.upload(c->h_nbr, s); c->finfo.upload(c->h_finfo, s); c->fmem.upload(c->h_fmem, s);
  c->mem_ci.upload(c->h_mem_ci, s); c->mem_fi.upload(c->h_mem_fi, s);
  c->fgeo.alloc((size_t)FGeo<D>::N * ND * nc);
  { FacetGeomKernel<D> gk{nc, c->grad.p, c->vol.p, c->h.p, c->nbr.p, c->finfo.p, c->fgeo.p}; parallel_for(s, nc, gk, 128); }
  {
    std::vector<int32_t> mc(c->h_mem_ci);
    mc.insert(mc.end(), c->h_mem_ce.begin(), c->h_mem_ce.end());
    std::sort(mc.begin(), mc.end());
    mc.erase(std::unique(mc.begin(), mc.end()), mc.end());
    c->nmc = (int64_t)mc.size();
    c->memcell.upload(mc, s);
  }
  // fields and matrices
  const int64_t n = c->n, nm = c->nm;
  c->phi.alloc(n); c->phiM.alloc(nm); c->rhs_emi.alloc(n);
  c->kappa.alloc(n); c->q.alloc(nc * D); c->gphi.alloc(nc * D);
  c->A_emi.alloc((size_t)(ND + 2) * nc * ND * ND);
  c->trace_tmp.alloc(nm);
  c->emi_assembled = c->knp_assembled = false;
  c->phi_old_valid = false;
  c->amg.ready = false;
  c->membranes.clear();
}

int knp_mesh_set(knp_ctx* ctx, int d, int64_t nc, int64_t nv, const double* coords,
                 const int32_t* cell_verts, const int32_t* cell_region, int64_t nf,
                 const int32_t* facet_cells, const int32_t* facet_tag, int n_mem_tags,
                 const int32_t* mem_tags) {
  KNP_TRY
  if (!ctx) fail("null context");
  if (nc <= 0 || nv <= 0) fail("knp_mesh_set: empty mesh");
  if (d == 2) build_mesh<2>(ctx, nc, nv, coords, cell_verts, cell_region, nf, facet_cells, facet_tag, n_mem_tags, mem_tags);
  else if (d == 3) build_mesh<3>(ctx, nc, nv, coords, cell_verts, cell_region, nf, facet_cells, facet_tag, n_mem_tags, mem_tags);
  else fail("knp_mesh_set: d must be 2 or 3");
  KNP_CATCH
}

int knp_mesh_info(knp_ctx* ctx, int64_t info[8]) {
  KNP_TRY
  info[0] = ctx->d; info[1] = ctx->nc; info[2] = ctx->n; info[3] = ctx->nm;
  info[4] = ctx->nnz_export; info[5] = ctx->nsip; info[6] = ctx->nd + 1; info[7] = 0;
  KNP_CATCH
}

int knp_membrane_table(knp_ctx* ctx, int32_t* facet, int32_t* cell_i, int32_t* cell_e, int32_t* tag) {
  KNP_TRY
  const size_t b = ctx->nm * sizeof(int32_t);
  if (facet) memcpy(facet, ctx->h_mem_facet.data(), b);
  if (cell_i) memcpy(cell_i, ctx->h_mem_ci.data(), b);
  if (cell_e) memcpy(cell_e, ctx->h_mem_ce.data(), b);
  if (tag) memcpy(tag, ctx->h_mem_tag.data(), b);
  KNP_CATCH
}

// ---------------------------------------------------------------------------------
// parameters / fields
// ---------------------------------------------------------------------------------
int knp_params_set(knp_ctx* ctx, double F, double R, double T, double C_M, double C_phi, double dt,
                   double tau_emi, double tau_knp, double Lp, int N, const double* z, int ntags,
                   const double* D, const double* rho, const double* C_sub, int splitting, int mms) {
  KNP_TRY
  if (ctx->nc == 0) fail("knp_params_set: set the mesh first");
  if (N < 2 || N > MAX_IONS) fail("knp_params_set: 2 <= N <= 6 ions supported");
  if (ntags < 1 || ntags > MAX_TAGS) fail("knp_params_set: 1 <= ntags <= 16 cell tags supported");
  Params& P = ctx->P;
  P.F = F; P.R = R; P.T = T; P.psi = F / (R * T); P.C_M = C_M; P.C_phi = C_phi; P.dt = dt;
  P.tau_emi = tau_emi; P.tau_knp = tau_knp; P.inv_Lp2 = 1.0 / (Lp * Lp);
  P.N = N; P.ntags = ntags; P.splitting = splitting; P.mms = mms;
  for (int k = 0; k < N; ++k) {
    P.z[k] = z[k];
    for (int t = 0; t < ntags; ++t) {
      P.D[k][t] = D[k * ntags + t];
      P.Csub[k][t] = (C_sub && k < N - 1) ? C_sub[k * ntags + t] : 0.0;
    }
  }
  for (int t = 0; t < ntags; ++t) P.rho[t] = rho ? rho[t] : 0.0;
  if (mms && !C_sub) fail("knp_params_set: mms mode needs C_sub");
  const int64_t n = ctx->n, nm = ctx->nm;
  for (int k = 0; k < N; ++k) {
    if (ctx->c[k].n != (size_t)n) ctx->c[k].alloc(n);
    if (ctx->Ich[k].n != (size_t)nm || !ctx->Ich[k].p) ctx->Ich[k].alloc(nm);
    if (ctx->E[k].n != (size_t)nm || !ctx->E[k].p) ctx->E[k].alloc(nm);
  }
  for (int k = 0; k < N - 1; ++k) {
    if (ctx->rhs_knp[k].n != (size_t)n) ctx->rhs_knp[k].alloc(n);
    if (ctx->A_knp[k].n != (size_t)(ctx->nd + 1) * ctx->slot_stride())
      ctx->A_knp[k].alloc((size_t)(ctx->nd + 1) * ctx->slot_stride());
  }
  ctx->params_set = true;
  ctx->amg_emi.solves = 0;                       // new coefficients: the next solves refresh their preconditioners
  for (int k = 0; k < MAX_IONS; ++k) ctx->amg_knp[k].solves = 0;
  KNP_CATCH
}

static double* field_ptr(knp_ctx* c, int which, int idx, int64_t& count, bool for_write) {
  const int N = c->P.N;
  auto need = [&](bool ok, const char* what) { if (!ok) fail(std::string("field index out of range: ") + what); };
  if (!c->params_set) fail("set parameters before touching fields");
  switch (which) {
    case KNP_F_C: need(idx >= 0 && idx < N, "C"); count = c->n; return c->c[idx].p;
    case KNP_F_CN:
      need(idx >= 0 && idx < N - 1, "CN"); count = c->n;
      if (for_write && !c->cn_separate[idx]) {
        if (c->cn_own[idx].n != (size_t)c->n) c->cn_own[idx].alloc(c->n);
        c->cn_separate[idx] = true;
      }
      return const_cast<double*>(c->cn(idx));
    case KNP_F_PHI: count = c->n; return c->phi.p;
    case KNP_F_PHIM: count = c->nm; return c->phiM.p;
    case KNP_F_ICH: need(idx >= 0 && idx < N, "ICH"); count = c->nm; return c->Ich[idx].p;
    case KNP_F_NERNST: need(idx >= 0 && idx < N, "NERNST"); count = c->nm; return c->E[idx].p;
    case KNP_F_RHS_EMI: count = c->n; return c->rhs_emi.p;
    case KNP_F_RHS_KNP: need(idx >= 0 && idx < N - 1, "RHS_KNP"); count = c->n; return c->rhs_knp[idx].p;
    case KNP_F_LOAD_EMI:
      count = c->n;
      if (for_write && !c->has_load_emi) { c->load_emi.alloc(c->n); c->has_load_emi = true; }
      if (!c->has_load_emi) fail("no EMI load vector set");
      return c->load_emi.p;
    case KNP_F_LOAD_KNP:
      need(idx >= 0 && idx < N - 1, "LOAD_KNP"); count = c->n;
      if (for_write && !c->has_load_knp[idx]) { c->load_knp[idx].alloc(c->n); c->has_load_knp[idx] = true; }
      if (!c->has_load_knp[idx]) fail("no KNP load vector set");
      return c->load_knp[idx].p;
  }
  fail("unknown field id");
}

int knp_field_set(knp_ctx* ctx, int which, int idx, const double* src, int64_t count) {
  KNP_TRY
  int64_t n = 0;
  double* p = field_ptr(ctx, which, idx, n, true);
  if (count != n) fail("knp_field_set: wrong element count");
  if (which == KNP_F_PHI) ctx->phi_old_valid = false;
  h2d(p, src, n * sizeof(double), ctx->stream);
  KNP_CATCH
}

int knp_field_get(knp_ctx* ctx, int which, int idx, double* dst, int64_t count) {
  KNP_TRY
  int64_t n = 0;
  double* p = field_ptr(ctx, which, idx, n, false);
  if (count != n) fail("knp_field_get: wrong element count");
  d2h(dst, p, n * sizeof(double), ctx->stream);
  KNP_CATCH
}

// ---------------------------------------------------------------------------------
// assembly
// ---------------------------------------------------------------------------------
#ifndef KNP_EMU
static int asm_min_blocks() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("KNP_ASM_MINB"); v = e ? atoi(e) : 4; }
  return v;
}
#endif

template <int D>
static void assemble_emi_t(knp_ctx* c) {
  knp_stream_t s = c->stream;
  EmiPrepassKernel<D> pre;
  pre.P = c->P; pre.nc = c->nc;
  for (int k = 0; k < MAX_IONS; ++k) pre.c[k] = c->c[k].p;
  pre.grad = c->grad.p; pre.region = c->region.p; pre.kappa = c->kappa.p; pre.q = c->q.p;
  parallel_for(s, c->nc, pre, 128);
  EmiArgs<D> k;
  k.P = c->P; k.nc = c->nc; k.nw = c->nc_own;
  k.grad = c->grad.p; k.vol = c->vol.p; k.h = c->h.p; k.fgeo = c->fgeo.p;
  k.nbr = c->nbr.p; k.finfo = c->finfo.p; k.fmem = c->fmem.p;
  k.kappa = c->kappa.p; k.q = c->q.p; k.phiM = c->phiM.p;
  for (int i = 0; i < MAX_IONS; ++i) k.Ich[i] = c->Ich[i].p;
  k.load = c->has_load_emi ? c->load_emi.p : nullptr;
  k.A = c->A_emi.p; k.Adiag = c->Adiag_emi(); k.rhs = c->rhs_emi.p;
#ifdef KNP_EMU
  parallel_for(s, c->nc_own, EmiCellKernel<D>{k}, 128);
#else
  ++launch_counter();
  {
    const unsigned grid = (unsigned)((c->nc_own + ASM_CPB - 1) / ASM_CPB);
    switch (asm_min_blocks()) {   // resident blocks per SM the register budget is compiled for
      case 5: emi_assemble_kernel<D, 5><<<grid, ASM_CPB*(D + 1), 0, s>>>(k); break;
      case 6: emi_assemble_kernel<D, 6><<<grid, ASM_CPB*(D + 1), 0, s>>>(k); break;
      default: emi_assemble_kernel<D, 4><<<grid, ASM_CPB*(D + 1), 0, s>>>(k); break;
    }
  }
  KNP_CUDA(cudaGetLastError());
#endif
}

template <int D>
static void assemble_knp_t(knp_ctx* c) {
  knp_stream_t s = c->stream;
  GradKernel<D> gk;
  gk.nc = c->nc; gk.phi = c->phi.p; gk.grad = c->grad.p; gk.gphi = c->gphi.p;
  parallel_for(s, c->nc, gk, 128);
  {
    KnpArgs<D> k;
    k.P = c->P; k.nc = c->nc; k.nw = c->nc_own; k.nion = c->P.N - 1;
    k.grad = c->grad.p; k.vol = c->vol.p; k.h = c->h.p; k.fgeo = c->fgeo.p; k.region = c->region.p;
    k.nbr = c->nbr.p; k.finfo = c->finfo.p; k.gphi = c->gphi.p;
    for (int i = 0; i < MAX_IONS; ++i) { k.cn[i] = nullptr; k.load[i] = nullptr; k.A[i] = nullptr; k.rhs[i] = nullptr; }
    for (int ion = 0; ion < k.nion; ++ion) {
      k.cn[ion] = c->cn(ion);
      k.load[ion] = c->has_load_knp[ion] ? c->load_knp[ion].p : nullptr;
      k.A[ion] = c->A_knp[ion].p; k.rhs[ion] = c->rhs_knp[ion].p;
    }
#ifdef KNP_EMU
    parallel_for(s, c->nc_own, KnpCellKernel<D>{k}, 128);
#else
    ++launch_counter();
    {
      const unsigned grid = (unsigned)((c->nc_own + ASM_CPB - 1) / ASM_CPB);
      switch (asm_min_blocks()) {
        case 5: knp_assemble_kernel<D, 5><<<grid, ASM_CPB*(D + 1), 0, s>>>(k); break;
        case 6: knp_assemble_kernel<D, 6><<<grid, ASM_CPB*(D + 1), 0, s>>>(k); break;
        default: knp_assemble_kernel<D, 4><<<grid, ASM_CPB*(D + 1), 0, s>>>(k); break;
      }
    }
    KNP_CUDA(cudaGetLastError());
#endif
  }
  if (c->nmc > 0) {
    KnpMembraneRhsKernel<D> m;
    m.P = c->P; m.nc = c->nc; m.memcell = c->memcell.p; m.vol = c->vol.p; m.grad = c->grad.p;
    m.region = c->region.p; m.nbr = c->nbr.p; m.finfo = c->finfo.p; m.fmem = c->fmem.p;
    m.phi = c->phi.p; m.phiM = c->phiM.p;
    for (int i = 0; i < MAX_IONS; ++i) { m.c[i] = c->c[i].p; m.Ich[i] = c->Ich[i].p; m.rhs[i] = c->rhs_knp[i].p; }
    parallel_for(s, c->nmc, m, 64);
  }
}

int knp_assemble_emi(knp_ctx* ctx) {
  KNP_TRY
  if (!ctx->params_set) fail("knp_assemble_emi: parameters not set");
  PhaseTimer t(ctx, T_EMI_ASM);
  if (ctx->d == 2) assemble_emi_t<2>(ctx); else assemble_emi_t<3>(ctx);
  ctx->emi_assembled = true;
  KNP_CATCH
}

int knp_assemble_knp(knp_ctx* ctx) {
  KNP_TRY
  if (!ctx->params_set) fail("knp_assemble_knp: parameters not set");
  PhaseTimer t(ctx, T_KNP_ASM);
  if (ctx->d == 2) assemble_knp_t<2>(ctx); else assemble_knp_t<3>(ctx);
  ctx->knp_assembled = true;
  KNP_CATCH
}

namespace knp {
BellMat bell_of(knp_ctx* c, int which) {
  BellMat M;
  M.nc = c->nc; M.nbr = c->nbr.p;
  if (which == 0) { M.off = c->A_emi.p; M.diag = c->Adiag_emi(); }
  else if (which == 1) { M.off = c->A_emi.p; M.diag = c->Bdiag(); }
  else {
    const int k = which - 2;
    if (k < 0 || k >= c->P.N - 1) fail("matrix id out of range");
    M.off = c->A_knp[k].p; M.diag = c->A_knp[k].p;
  }
  return M;
}
}  // namespace knp

int knp_matrix_export(knp_ctx* ctx, int which, int64_t* rowptr, int32_t* col, double* val) {
  KNP_TRY
  const int nd = ctx->nd;
  const int64_t nc = ctx->nc, bs = ctx->bs(), ss = ctx->slot_stride();
  BellMat M = bell_of(ctx, which);
  std::vector<double> off((size_t)(nd + 1) * ss), diag((size_t)ss);
  d2h(off.data(), M.off, off.size() * sizeof(double), ctx->stream);
  d2h(diag.data(), M.diag, diag.size() * sizeof(double), ctx->stream);
  int64_t k = 0;
  rowptr[0] = 0;
  std::vector<std::pair<int32_t, int>> order;  // (neighbour cell, slot)
  for (int64_t c = 0; c < ctx->nc_own; ++c) {
    order.clear();
    order.push_back({(int32_t)c, 0});
    for (int f = 0; f < nd; ++f) {
      const int32_t c2 = ctx->h_nbr[(size_t)f * nc + c];
      if (c2 >= 0) order.push_back({c2, 1 + f});
    }
    std::sort(order.begin(), order.end());
    for (int i = 0; i < nd; ++i) {
      for (auto& o : order) {
        const double* blk = o.second == 0 ? &diag[c * bs] : &off[(size_t)o.second * ss + c * bs];
        for (int j = 0; j < nd; ++j) { col[k] = o.first * nd + j; val[k] = blk[i * nd + j]; ++k; }
      }
      rowptr[c * nd + i + 1] = k;
    }
  }
  if (k != ctx->nnz_export) fail("internal: export nnz mismatch");
  KNP_CATCH
}

int knp_spmv(knp_ctx* ctx, int which, const double* x, double* y) {
  KNP_TRY
  BellMat M = bell_of(ctx, which);
  DevBuf<double> dx, dy;
  dx.upload(x, ctx->n, ctx->stream); dy.alloc(ctx->n);
  if (ctx->comm.active()) ctx->comm.halo(ctx->stream, ctx->halo0, dx.p);
  if (ctx->d == 2) { BellSpmvKernel<3> k{M, dx.p, nullptr, dy.p, 0}; parallel_for(ctx->stream, ctx->n_own, k); }
  else { BellSpmvKernel<4> k{M, dx.p, nullptr, dy.p, 0}; parallel_for(ctx->stream, ctx->n_own, k); }
  d2h(y, dy.p, ctx->n * sizeof(double), ctx->stream);
  KNP_CATCH
}

// ---------------------------------------------------------------------------------
// post-step
// ---------------------------------------------------------------------------------
template <int D>
static void post_step_t(knp_ctx* c, int what) {
  knp_stream_t s = c->stream;
  if (what & KNP_POST_ELIMINATED) {
  EliminatedIonKernel<D> ek;
  ek.P = c->P; ek.region = c->region.p;
  for (int k = 0; k < MAX_IONS; ++k) ek.c[k] = c->c[k].p;
  ek.celim = c->c[c->P.N - 1].p;
  parallel_for(s, c->n, ek);
  }
  if (c->nm > 0 && (what & (KNP_POST_PHIM | KNP_POST_NERNST))) {
    MembranePostKernel<D> mk;
    mk.P = c->P; mk.nc = c->nc; mk.mem_ci = c->mem_ci.p; mk.mem_fi = c->mem_fi.p;
    mk.nbr = c->nbr.p; mk.finfo = c->finfo.p; mk.phi = c->phi.p;
    for (int k = 0; k < MAX_IONS; ++k) { mk.c[k] = c->c[k].p; mk.E[k] = c->E[k].p; }
    mk.phiM = c->phiM.p;
    mk.do_phim = (what & KNP_POST_PHIM) ? 1 : 0;
    mk.do_nernst = (!c->P.mms && (what & KNP_POST_NERNST)) ? 1 : 0;
    parallel_for(s, c->nm, mk, 128);
  }
}

int knp_post_step(knp_ctx* ctx, int what) {
  KNP_TRY
  if (!ctx->params_set) fail("knp_post_step: parameters not set");
  PhaseTimer t(ctx, T_POST);
  if (ctx->d == 2) post_step_t<2>(ctx, what); else post_step_t<3>(ctx, what);
  // KNP_POST_ALL is the update at the end of a regular time step, which includes
  // c_prev_n.assign(c) (solver.py:809-810): a c_prev_n that was written separately through
  // KNP_F_CN (Picard iteration, tests) is dropped, KNP_F_CN reads c again.
  if ((what & KNP_POST_ALL) == KNP_POST_ALL)
    for (int k = 0; k < MAX_IONS; ++k) ctx->cn_separate[k] = false;
  KNP_CATCH
}

int knp_facet_trace(knp_ctx* ctx, int which, int idx, int side, double* out) {
  KNP_TRY
  int64_t cnt = 0;
  const double* f = field_ptr(ctx, which, idx, cnt, false);
  if (cnt != ctx->n) fail("knp_facet_trace: not a cell field");
  if (ctx->nm == 0) return 0;
  if (ctx->d == 2) {
    FacetTraceKernel<2> k{ctx->nc, ctx->mem_ci.p, ctx->mem_fi.p, ctx->nbr.p, ctx->finfo.p, f, side, ctx->trace_tmp.p};
    parallel_for(ctx->stream, ctx->nm, k);
  } else {
    FacetTraceKernel<3> k{ctx->nc, ctx->mem_ci.p, ctx->mem_fi.p, ctx->nbr.p, ctx->finfo.p, f, side, ctx->trace_tmp.p};
    parallel_for(ctx->stream, ctx->nm, k);
  }
  d2h(out, ctx->trace_tmp.p, ctx->nm * sizeof(double), ctx->stream);
  KNP_CATCH
}

// ---------------------------------------------------------------------------------
// membrane ODEs
// ---------------------------------------------------------------------------------
int knp_model_count(void) { return KNP_NUM_MODELS; }
const char* knp_model_name(int id) { return (id >= 0 && id < KNP_NUM_MODELS) ? knp_model_names[id] : ""; }
int knp_model_dims(int id, int* ns, int* np) {
  KNP_TRY
  if (id < 0 || id >= KNP_NUM_MODELS) fail("unknown model id");
  *ns = knp_model_ns[id]; *np = knp_model_np[id];
  KNP_CATCH
}

static MembraneSet& mset(knp_ctx* c, int handle) {
  if (handle < 0 || handle >= (int)c->membranes.size()) fail("bad membrane handle");
  return *c->membranes[handle];
}

int knp_membrane_register(knp_ctx* ctx, int model_id, int64_t nrows, const int32_t* rows,
                          const double* states, const double* params, int* handle) {
  KNP_TRY
  if (model_id < 0 || model_id >= KNP_NUM_MODELS) fail("unknown model id");
  for (int64_t i = 0; i < nrows; ++i)
    if (rows[i] < 0 || rows[i] >= ctx->nm) fail("membrane row out of range");
  std::unique_ptr<MembraneSet> m(new MembraneSet());
  m->model_id = model_id; m->ns = knp_model_ns[model_id]; m->np = knp_model_np[model_id];
  m->nrows = nrows;
  m->rows.upload(rows, nrows, ctx->stream);
  m->states.upload(states, nrows * m->ns, ctx->stream);
  m->params.upload(params, nrows * m->np, ctx->stream);
  ctx->membranes.push_back(std::move(m));
  *handle = (int)ctx->membranes.size() - 1;
  KNP_CATCH
}

int knp_membrane_states_get(knp_ctx* ctx, int h, double* out) {
  KNP_TRY
  MembraneSet& m = mset(ctx, h);
  d2h(out, m.states.p, m.nrows * m.ns * sizeof(double), ctx->stream);
  KNP_CATCH
}
int knp_membrane_states_set(knp_ctx* ctx, int h, const double* in) {
  KNP_TRY
  MembraneSet& m = mset(ctx, h);
  h2d(m.states.p, in, m.nrows * m.ns * sizeof(double), ctx->stream);
  KNP_CATCH
}
int knp_membrane_params_get(knp_ctx* ctx, int h, double* out) {
  KNP_TRY
  MembraneSet& m = mset(ctx, h);
  d2h(out, m.params.p, m.nrows * m.np * sizeof(double), ctx->stream);
  KNP_CATCH
}
int knp_membrane_params_set(knp_ctx* ctx, int h, const double* in) {
  KNP_TRY
  MembraneSet& m = mset(ctx, h);
  h2d(m.params.p, in, m.nrows * m.np * sizeof(double), ctx->stream);
  KNP_CATCH
}

int knp_membrane_link(knp_ctx* ctx, int h, int col, int kind, int which, int idx, int side) {
  KNP_TRY
  MembraneSet& m = mset(ctx, h);
  if (col < 0 || col >= m.np) fail("link: parameter column out of range");
  int64_t cnt = 0;
  field_ptr(ctx, which, idx, cnt, false);
  if (kind == 0 && cnt != ctx->nm) fail("link kind 0 needs a membrane-row field");
  if (kind == 1 && cnt != ctx->n) fail("link kind 1 needs a cell field");
  for (auto& l : m.links)
    if (l.col == col) { l = LinkSpec{col, kind, which, idx, side}; return 0; }
  if ((int)m.links.size() >= MAX_LINKS) fail("too many links");
  m.links.push_back(LinkSpec{col, kind, which, idx, side});
  KNP_CATCH
}

int knp_membrane_outputs(knp_ctx* ctx, int h, int v_col, int n_ion, const int32_t* ich_cols) {
  KNP_TRY
  MembraneSet& m = mset(ctx, h);
  if (v_col < 0 || v_col >= m.ns) fail("V column out of range");
  if (n_ion < 0 || n_ion > ctx->P.N) fail("too many ion currents");
  m.v_col = v_col; m.n_ion = n_ion;
  for (int k = 0; k < n_ion; ++k) {
    if (ich_cols[k] < 0 || ich_cols[k] >= m.np) fail("I_ch column out of range");
    m.ich_cols[k] = ich_cols[k];
  }
  KNP_CATCH
}

int knp_membrane_stimulus(knp_ctx* ctx, int h, const uint8_t* mask, int ncols, const int32_t* cols,
                          const double* values) {
  KNP_TRY
  MembraneSet& m = mset(ctx, h);
  if (ncols < 0 || ncols > MAX_STIM) fail("at most 4 stimulus parameters");
  m.nstim = ncols;
  for (int i = 0; i < ncols; ++i) {
    if (cols[i] < 0 || cols[i] >= m.np) fail("stimulus column out of range");
    m.stim_cols[i] = cols[i]; m.stim_vals[i] = values[i];
  }
  if (mask && ncols > 0) { m.mask.upload(mask, m.nrows, ctx->stream); m.has_mask = true; }
  else m.has_mask = false;
  KNP_CATCH
}

template <class M, int D>
static void ode_launch(knp_ctx* c, MembraneSet& m, double t0, double dt, double rtol, double atol,
                       int set_v, int64_t* stats_dev) {
  OdeStepKernel<M, D> k;
  k.nc = c->nc; k.rows = m.rows.p; k.states = m.states.p; k.params = m.params.p;
  k.nlinks = (int)m.links.size();
  for (int l = 0; l < k.nlinks; ++l) {
    int64_t cnt = 0;
    k.links[l].col = m.links[l].col; k.links[l].kind = m.links[l].kind; k.links[l].side = m.links[l].side;
    k.links[l].src = field_ptr(c, m.links[l].which, m.links[l].idx, cnt, false);
  }
  for (int l = k.nlinks; l < MAX_LINKS; ++l) k.links[l] = OdeLink{0, 0, 0, nullptr};
  k.set_v = set_v; k.v_col = m.v_col; k.phiM = c->phiM.p;
  k.n_ion = m.n_ion;
  for (int i = 0; i < MAX_IONS; ++i) { k.ich_cols[i] = m.ich_cols[i]; k.Ich[i] = c->Ich[i].p; }
  k.stim_mask = m.has_mask ? m.mask.p : nullptr; k.nstim = m.nstim;
  for (int i = 0; i < MAX_STIM; ++i) { k.stim_cols[i] = m.stim_cols[i]; k.stim_vals[i] = m.stim_vals[i]; }
  k.mem_ci = c->mem_ci.p; k.mem_fi = c->mem_fi.p; k.nbr = c->nbr.p; k.finfo = c->finfo.p;
  k.t0 = t0; k.dt = dt; k.rtol = rtol; k.atol = atol; k.stats = stats_dev;
  parallel_for(c->stream, m.nrows, k, 64);
}

int knp_ode_step(knp_ctx* ctx, int h, double t0, double dt, double rtol, double atol, int set_v,
                 int64_t* stats) {
  KNP_TRY
  MembraneSet& m = mset(ctx, h);
  if (m.v_col < 0) fail("knp_ode_step: call knp_membrane_outputs first");
  PhaseTimer t(ctx, T_ODE);
  dev_zero(ctx->ode_stats.p, 4 * sizeof(int64_t), ctx->stream);
  bool done = false;
#define KNP_MODEL_CASE(ID, TYPE, NAME)                                                          \
  if (m.model_id == ID) {                                                                       \
    if (ctx->d == 2) ode_launch<TYPE, 2>(ctx, m, t0, dt, rtol, atol, set_v, ctx->ode_stats.p);  \
    else ode_launch<TYPE, 3>(ctx, m, t0, dt, rtol, atol, set_v, ctx->ode_stats.p);              \
    done = true;                                                                                \
  }
  KNP_MODEL_LIST(KNP_MODEL_CASE)
#undef KNP_MODEL_CASE
  if (!done) fail("model not compiled in");
  int64_t hs[4];
  d2h(hs, ctx->ode_stats.p, sizeof hs, ctx->stream);
  if (stats) { stats[0] = hs[0]; stats[1] = hs[1]; stats[2] = hs[3]; }
  if (hs[2] != 0)
    fail("knp_ode_step: model '" + std::string(knp_model_names[m.model_id]) + "' (membrane handle " + std::to_string(h) +
         "): the integrator failed on " + std::to_string(hs[2]) + " of " + std::to_string(m.nrows) +
         " facets in [" + std::to_string(t0) + ", " + std::to_string(t0 + dt) + "] (step size underflow, NaN or step limit; " +
         std::to_string(hs[3]) + " facets were on the stiff path)");
  KNP_CATCH
}

// ---------------------------------------------------------------------------------
// multi-GPU
// ---------------------------------------------------------------------------------
int knp_dist_set(knp_ctx* ctx, int rank, int world, int64_t nc_owned, int nneigh, const int32_t* neigh_rank,
                 const int64_t* send_ptr, const int32_t* send_cells, const int64_t* recv_ptr) {
  KNP_TRY
  knp_ctx* c = ctx;
  if (c->nc == 0) fail("knp_dist_set: set the mesh first");
  if (c->params_set) fail("knp_dist_set: call before knp_params_set");
  if (world < 1 || rank < 0 || rank >= world) fail("knp_dist_set: bad rank/world");
  if (nc_owned < 1 || nc_owned > c->nc) fail("knp_dist_set: bad owned-cell count");
  if (nneigh < 0 || (nneigh > 0 && (!neigh_rank || !send_ptr || !recv_ptr))) fail("knp_dist_set: bad neighbour lists");
  const int nd = c->nd;
  const int64_t nghost = c->nc - nc_owned;
  if ((nneigh ? recv_ptr[nneigh] : 0) != nghost) fail("knp_dist_set: recv_ptr does not cover the ghost cells");
  for (int64_t cell = 0; cell < nc_owned; ++cell)
    for (int f = 0; f < nd; ++f)
      if (c->h_nbr[(size_t)f * c->nc + cell] >= c->nc) fail("knp_dist_set: neighbour out of range");
  c->comm.rank = rank; c->comm.world = world;
  c->comm.nbr.assign(neigh_rank, neigh_rank + nneigh);
  for (int i = 0; i < nneigh; ++i) {
    if (neigh_rank[i] < 0 || neigh_rank[i] >= world || neigh_rank[i] == rank) fail("knp_dist_set: bad neighbour rank");
    if (i > 0 && neigh_rank[i] <= neigh_rank[i - 1]) fail("knp_dist_set: neighbour ranks must ascend");
  }
  c->nc_own = nc_owned; c->n_own = nc_owned * nd;
  HaloPlan& H = c->halo0;
  H = HaloPlan();
  H.n_own = c->n_own; H.n_ghost = nghost * nd;
  H.send_off.assign(nneigh + 1, 0); H.recv_off.assign(nneigh + 1, 0);
  for (int i = 0; i < nneigh; ++i) {
    if (send_ptr[i + 1] < send_ptr[i] || recv_ptr[i + 1] < recv_ptr[i]) fail("knp_dist_set: pointers must ascend");
    for (int64_t k = send_ptr[i]; k < send_ptr[i + 1]; ++k) {
      const int32_t cell = send_cells[k];
      if (cell < 0 || cell >= nc_owned) fail("knp_dist_set: send cell is not an owned cell");
      for (int a = 0; a < nd; ++a) H.h_send_idx.push_back(cell * nd + a);
    }
    H.send_off[i + 1] = send_ptr[i + 1] * nd;
    H.recv_off[i + 1] = recv_ptr[i + 1] * nd;
    for (int64_t g = recv_ptr[i]; g < recv_ptr[i + 1]; ++g)
      for (int a = 0; a < nd; ++a) { H.ghost_rank.push_back(neigh_rank[i]); H.ghost_id.push_back(-1); }
  }
  H.upload(c->stream);
  // rows exist for owned cells only
  int64_t nblocks = 0;
  for (int64_t cell = 0; cell < nc_owned; ++cell) {
    ++nblocks;
    for (int f = 0; f < nd; ++f) nblocks += c->h_nbr[(size_t)f * c->nc + cell] >= 0;
  }
  c->nnz_export = nblocks * nd * nd;
  // leading owned cells without a ghost neighbour (their rows do not wait for the halo)
  c->nc_int = 0;
  for (int64_t cell = 0; cell < nc_owned; ++cell) {
    bool interior = true;
    for (int f = 0; f < nd; ++f) interior = interior && c->h_nbr[(size_t)f * c->nc + cell] < nc_owned;
    if (!interior) break;
    c->nc_int = cell + 1;
  }
  {
    std::vector<int32_t> mc;
    for (int32_t v : c->h_mem_ci) if (v < nc_owned) mc.push_back(v);
    for (int32_t v : c->h_mem_ce) if (v < nc_owned) mc.push_back(v);
    std::sort(mc.begin(), mc.end());
    mc.erase(std::unique(mc.begin(), mc.end()), mc.end());
    c->nmc = (int64_t)mc.size();
    c->memcell.upload(mc, c->stream);
  }
  c->amg.ready = false;
  KNP_CATCH
}

int knp_nccl_unique_id(char out[128]) {
  KNP_TRY
#ifdef KNP_EMU
  (void)out;
  fail("knp_nccl_unique_id: the host-emulation build has no NCCL transport");
#else
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  NcclApi& N = nccl_api();
  N.load();
  ncclUniqueId id;
  N.check(N.GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out, &id, 128);
#endif
  KNP_CATCH
}

int knp_dist_init_nccl(knp_ctx* ctx, const char uid[128]) {
  KNP_TRY
#ifdef KNP_EMU
  (void)ctx; (void)uid;
  fail("knp_dist_init_nccl: the host-emulation build has no NCCL transport");
#else
  if (ctx->comm.world < 2) fail("knp_dist_init_nccl: call knp_dist_set first (world >= 2)");
  NcclApi& N = nccl_api();
  N.load();
  KNP_CUDA(cudaSetDevice(ctx->device));
  if (ctx->comm.nccl) { N.CommDestroy(ctx->comm.nccl); ctx->comm.nccl = nullptr; }
  ncclUniqueId id;
  memcpy(&id, uid, 128);
  N.check(N.CommInitRank(&ctx->comm.nccl, ctx->comm.world, id, ctx->comm.rank), "ncclCommInitRank");
  ctx->comm.setup_p2p(ctx->stream);   // peer-memory kernels where CUDA IPC works, NCCL otherwise
  // overlap of the level-0 halo with the interior rows (KNP_OVERLAP=0 disables): a second stream for the
  // exchange kernel, two events to fork from and join the main stream
  {
    const char* e = getenv("KNP_OVERLAP");
    ctx->overlap = ctx->comm.p2p.on && !(e && e[0] == '0');
    if (ctx->overlap && !ctx->comm_stream) {
      // highest priority: the few blocks of the exchange kernel must not queue behind the thousands of
      // blocks of the interior-row kernel that runs beside it
      int prio_lo = 0, prio_hi = 0;
      KNP_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      KNP_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, prio_hi));
      KNP_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
      KNP_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
    }
  }
#endif
  KNP_CATCH
}

int knp_dist_set_callbacks(knp_ctx* ctx, knp_exchange_fn exchange, knp_allreduce_fn allreduce, void* user) {
  KNP_TRY
#ifndef KNP_EMU
  (void)ctx; (void)exchange; (void)allreduce; (void)user;
  fail("knp_dist_set_callbacks: host callbacks are a test transport of the emulation build; "
       "the CUDA build communicates with NCCL (knp_dist_init_nccl)");
#else
  ctx->comm.xfn = exchange; ctx->comm.rfn = allreduce; ctx->comm.user = user;
#endif
  KNP_CATCH
}

int knp_dist_info(knp_ctx* ctx, int64_t info[8]) {
  KNP_TRY
  info[0] = ctx->comm.rank; info[1] = ctx->comm.world; info[2] = ctx->nc_own; info[3] = ctx->nc - ctx->nc_own;
  info[4] = (int64_t)ctx->comm.nbr.size(); info[5] = ctx->comm.n_halo; info[6] = ctx->comm.n_allreduce;
#ifdef KNP_EMU
  info[7] = 0;
#else
  info[7] = ctx->comm.n_p2p;
#endif
  KNP_CATCH
}

int knp_field_halo(knp_ctx* ctx, int which, int idx) {
  KNP_TRY
  int64_t cnt = 0;
  double* p = field_ptr(ctx, which, idx, cnt, false);
  if (cnt != ctx->n) fail("knp_field_halo: not a cell field");
  ctx->comm.halo(ctx->stream, ctx->halo0, p);
  stream_sync(ctx->stream);
  KNP_CATCH
}

int knp_timers_get(knp_ctx* ctx, double* out, int reset) {
  KNP_TRY
  for (int i = 0; i < T_COUNT; ++i) { out[i] = ctx->timers[i]; if (reset) ctx->timers[i] = 0.0; }
  KNP_CATCH
}


// ---------------------------------------------------------------------------------
// measurement hooks
// ---------------------------------------------------------------------------------
long long knp_launch_count(void) { return launch_counter(); }

int knp_timer_start(knp_ctx* ctx) {
  KNP_TRY
#ifdef KNP_EMU
  ctx->host_t0 = now_s();
#else
  if (!ctx->ev0) { KNP_CUDA(cudaEventCreate(&ctx->ev0)); KNP_CUDA(cudaEventCreate(&ctx->ev1)); }
  KNP_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
#endif
  KNP_CATCH
}

int knp_timer_stop(knp_ctx* ctx, double* ms) {
  KNP_TRY
#ifdef KNP_EMU
  *ms = (now_s() - ctx->host_t0) * 1e3;
#else
  if (!ctx->ev0) fail("knp_timer_stop without knp_timer_start");
  KNP_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  KNP_CUDA(cudaEventSynchronize(ctx->ev1));
  float f = 0.f;
  KNP_CUDA(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = f;
#endif
  KNP_CATCH
}

int knp_bench_kernel(knp_ctx* ctx, int kernel, int reps, double* ms, double* bytes) {
  KNP_TRY
  if (!ctx->params_set) fail("knp_bench_kernel: context not set up");
  if (reps < 1) reps = 1;
  const double nc = (double)ctx->nc_own, n = (double)ctx->n_own, nd = ctx->nd, d = ctx->d, bs = (double)ctx->bs();
  const double N = ctx->P.N;
  const double geom = 8.0 * nc * nd * d + 16.0 * nc + 4.0 * nc + 8.0 * nd * nc;  // grad, vol+h, region, nbr+finfo
  DevBuf<double> x, y;
  if (kernel == 0 || kernel == 3) { x.alloc(ctx->n); y.alloc(ctx->n); }
  if (kernel >= 4 && !ctx->comm.active()) { *ms = 0.0; *bytes = 0.0; return 0; }
  stream_sync(ctx->stream);
  double t_ms = 0.0;
  knp_timer_start(ctx);
  for (int r = 0; r < reps; ++r) {
    switch (kernel) {
      case 0: {
        BellMat M = bell_of(ctx, 0);
        if (ctx->d == 2) { BellSpmvKernel<3> k{M, x.p, nullptr, y.p, 0}; parallel_for(ctx->stream, ctx->n_own, k); }
        else { BellSpmvKernel<4> k{M, x.p, nullptr, y.p, 0}; parallel_for(ctx->stream, ctx->n_own, k); }
        break;
      }
      case 1: if (ctx->d == 2) assemble_emi_t<2>(ctx); else assemble_emi_t<3>(ctx); break;
      case 2: if (ctx->d == 2) assemble_knp_t<2>(ctx); else assemble_knp_t<3>(ctx); break;
      case 3: {
        BellMat M = bell_of(ctx, 1);
        if (ctx->d == 2) { BellJacobiKernel<3> k{M, ctx->Adiag_emi(), ctx->rhs_emi.p, x.p, y.p, 0.7}; parallel_for(ctx->stream, ctx->n_own, k, 192); }
        else { BellJacobiKernel<4> k{M, ctx->Adiag_emi(), ctx->rhs_emi.p, x.p, y.p, 0.7}; parallel_for(ctx->stream, ctx->n_own, k, 256); }
        break;
      }
      case 4: ctx->comm.halo(ctx->stream, ctx->halo0, ctx->phi.p); break;          // DG halo exchange
      case 5: ctx->comm.allreduce(ctx->stream, ctx->kr0.scal.p + 900, 4); break;    // Krylov scalars
      case 6:                                                                      // AMG tail all-gather
        if (ctx->amg.ready && ctx->amg.rep_from != (size_t)-1) ctx->comm.allgather(ctx->stream, ctx->amg_emi.rep_b.p, ctx->amg.rep_bstride);
        break;
      default: fail("unknown kernel id");
    }
  }
  if (knp_timer_stop(ctx, &t_ms)) fail(g_last_error);
  *ms = t_ms / reps;
  double b = 0.0;
  const double mat = 8.0 * (double)ctx->nnz_export;
  switch (kernel) {
    case 0: b = mat + 4.0 * nd * nc + 16.0 * n; break;
    case 1: b = /*prepass*/ 8.0 * n * N + 8.0 * nc * nd * d + 4.0 * nc + 8.0 * n + 8.0 * nc * d
              /*cells*/ + geom + 8.0 * n + 8.0 * nc * d + mat + 8.0 * nc * bs + 8.0 * n; break;
    case 2: b = /*grad*/ 8.0 * n + 8.0 * nc * nd * d + 8.0 * nc * d
              + geom + 8.0 * nc * d + (N - 1) * (8.0 * n + mat + 8.0 * n); break;
    case 3: b = mat + 8.0 * nc * bs + 4.0 * nd * nc + 24.0 * n; break;
    case 4: b = 8.0 * (double)(ctx->halo0.nsend() + ctx->halo0.n_ghost); break;
    case 5: b = 32.0; break;
    case 6: b = 8.0 * (double)ctx->amg.rep_bstride * ctx->comm.world; break;
  }
  *bytes = b;
  KNP_CATCH
}

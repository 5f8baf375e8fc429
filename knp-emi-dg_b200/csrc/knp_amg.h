// knp_amg.h - aggregation AMG whose *plan* (aggregates, transfer operators, coarse
// sparsity, Galerkin gather lists) is built once on the host and reused every time
// step; only the numeric Galerkin products are refreshed on the device after each
// re-assembly.
//
// Replaces hypre BoomerAMG as used by the reference (`pc_type hypre`,
// src/knpemidg/solver.py:433-444, 688-701), which redoes its whole setup at every
// `setOperators` (twice per time step, solver.py:505, 767).
//
// Hierarchy for DG-P1:
//   level 0  DG1 (block-ELL)            smoother: damped element block-Jacobi
//   level 1  region-wise continuous P1  = DG dofs glued at mesh vertices across tag-0
//            facets (never across membranes); the prolongation is the natural injection
//   level 2+ plain aggregation on the strength graph of level 1, l1-Jacobi smoothing
//   last     dense inverse
#pragma once
#include <algorithm>
#include <numeric>
#include "knp_common.h"
#include "knp_comm.h"
#include "knp_linalg.h"

namespace knp {

// host CSR used while planning
struct HostCsr {
  int64_t n = 0;
  std::vector<int32_t> ptr, col;
  std::vector<int32_t> pos;   // storage position of each entry in the device value array
  std::vector<double> val;    // values (setup time only)
};

struct HostTransfer {         // P (n_f x n_c) and R = P^T in CSR
  int64_t nf = 0, ncoarse = 0;
  std::vector<int32_t> pptr, pidx; std::vector<double> pw;
  std::vector<int32_t> rptr, ridx; std::vector<double> rw;
  bool unit = true;           // all weights 1 (pure aggregation)
};

// agg[i] = coarse unknown of fine unknown i for ALL local fine unknowns (owned first, then
// ghosts, whose coarse unknowns are ghosts of the coarse level); the restriction has one row
// per OWNED coarse unknown and gathers owned fine unknowns only (aggregates never cross
// a partition boundary).
inline HostTransfer transfer_from_aggregates(const std::vector<int32_t>& agg, int64_t ncoarse,
                                             int64_t nf_own = -1) {
  HostTransfer T;
  T.nf = (int64_t)agg.size(); T.ncoarse = ncoarse; T.unit = true;
  if (nf_own < 0) nf_own = T.nf;
  T.pptr.resize(T.nf + 1); T.pidx.resize(T.nf);
  for (int64_t i = 0; i <= T.nf; ++i) T.pptr[i] = (int32_t)i;
  for (int64_t i = 0; i < T.nf; ++i) T.pidx[i] = agg[i];
  T.rptr.assign(ncoarse + 1, 0);
  for (int64_t i = 0; i < nf_own; ++i) T.rptr[agg[i] + 1]++;
  for (int64_t I = 0; I < ncoarse; ++I) T.rptr[I + 1] += T.rptr[I];
  T.ridx.resize(nf_own);
  std::vector<int32_t> fill(T.rptr.begin(), T.rptr.end() - 1);
  for (int64_t i = 0; i < nf_own; ++i) T.ridx[fill[agg[i]]++] = (int32_t)i;
  return T;
}

struct GalerkinPlan {
  HostCsr coarse;                        // pattern (+ setup-time values)
  std::vector<int32_t> gptr, gidx;       // per coarse nnz: list of fine storage positions
  std::vector<double> gw;                // weights (empty when all 1)
};

// coarse = P^T A P as a gather plan over the fine value array
inline GalerkinPlan galerkin_plan(const HostCsr& A, const HostTransfer& T) {
  GalerkinPlan G;
  const int64_t ncoarse = T.ncoarse;
  G.coarse.n = ncoarse;
  G.coarse.ptr.assign(ncoarse + 1, 0);
  struct Item { int32_t J; int32_t pos; double w; };
  std::vector<Item> items;
  std::vector<int32_t> ccol, gptr, gidx;
  std::vector<double> gw, cval;
  gptr.push_back(0);
  const bool unit = T.unit;
  for (int64_t I = 0; I < ncoarse; ++I) {
    items.clear();
    for (int32_t a = T.rptr[I]; a < T.rptr[I + 1]; ++a) {
      const int32_t i = T.ridx[a];
      const double wi = unit ? 1.0 : T.rw[a];
      for (int32_t e = A.ptr[i]; e < A.ptr[i + 1]; ++e) {
        const int32_t j = A.col[e];
        for (int32_t b = T.pptr[j]; b < T.pptr[j + 1]; ++b)
          items.push_back({T.pidx[b], A.pos[e], wi * (unit ? 1.0 : T.pw[b])});
      }
    }
    std::stable_sort(items.begin(), items.end(),
                     [](const Item& x, const Item& y) { return x.J < y.J; });
    size_t t = 0;
    while (t < items.size()) {
      const int32_t J = items[t].J;
      ccol.push_back(J);
      while (t < items.size() && items[t].J == J) {
        gidx.push_back(items[t].pos);
        if (!unit) gw.push_back(items[t].w);
        ++t;
      }
      gptr.push_back((int32_t)gidx.size());
    }
    G.coarse.ptr[I + 1] = (int32_t)ccol.size();
  }
  G.coarse.col = std::move(ccol);
  G.coarse.pos.resize(G.coarse.col.size());
  std::iota(G.coarse.pos.begin(), G.coarse.pos.end(), 0);
  G.gptr = std::move(gptr); G.gidx = std::move(gidx); G.gw = std::move(gw);
  return G;
}

// numeric Galerkin on the host (setup only: strength of connection needs values)
inline void galerkin_numeric_host(GalerkinPlan& G, const std::vector<double>& fine) {
  const size_t nnz = G.coarse.col.size();
  G.coarse.val.resize(nnz);
  for (size_t k = 0; k < nnz; ++k) {
    double acc = 0.0;
    for (int32_t t = G.gptr[k]; t < G.gptr[k + 1]; ++t)
      acc += (G.gw.empty() ? 1.0 : G.gw[t]) * fine[G.gidx[t]];
    G.coarse.val[k] = acc;
  }
}

// greedy aggregation on the strength graph |a_ij| >= theta sqrt(|a_ii a_jj|)
// (Vanek-Mandel-Brezina three-pass scheme).
inline int64_t aggregate(const HostCsr& A, double theta, std::vector<int32_t>& agg) {
  const int64_t n = A.n;
  std::vector<double> diag(n, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int32_t e = A.ptr[i]; e < A.ptr[i + 1]; ++e)
      if (A.col[e] == i) diag[i] = fabs(A.val[e]);
  std::vector<int32_t> sptr(n + 1, 0), scol;
  scol.reserve(A.col.size());
  for (int64_t i = 0; i < n; ++i) {
    for (int32_t e = A.ptr[i]; e < A.ptr[i + 1]; ++e) {
      const int32_t j = A.col[e];
      if (j == i || j >= n) continue;   // j >= n: ghost unknown, owned by another rank
      if (fabs(A.val[e]) >= theta * sqrt(diag[i] * diag[j]) && A.val[e] != 0.0) scol.push_back(j);
    }
    sptr[i + 1] = (int32_t)scol.size();
  }
  agg.assign(n, -1);
  int32_t na = 0;
  for (int64_t i = 0; i < n; ++i) {       // pass 1: roots with fully free neighbourhoods
    if (agg[i] >= 0 || sptr[i] == sptr[i + 1]) continue;
    bool free_nb = true;
    for (int32_t e = sptr[i]; e < sptr[i + 1]; ++e)
      if (agg[scol[e]] >= 0) { free_nb = false; break; }
    if (!free_nb) continue;
    agg[i] = na;
    for (int32_t e = sptr[i]; e < sptr[i + 1]; ++e) agg[scol[e]] = na;
    ++na;
  }
  std::vector<int32_t> agg2(agg);
  for (int64_t i = 0; i < n; ++i) {       // pass 2: attach to a neighbouring aggregate
    if (agg[i] >= 0) continue;
    for (int32_t e = sptr[i]; e < sptr[i + 1]; ++e)
      if (agg[scol[e]] >= 0) { agg2[i] = agg[scol[e]]; break; }
  }
  agg.swap(agg2);
  for (int64_t i = 0; i < n; ++i) {       // pass 3: leftovers (and isolated rows)
    if (agg[i] >= 0) continue;
    agg[i] = na;
    for (int32_t e = sptr[i]; e < sptr[i + 1]; ++e)
      if (agg[scol[e]] < 0) agg[scol[e]] = na;
    ++na;
  }
  return na;
}

// ---- device side -----------------------------------------------------------------
struct AmgLevelPlan {          // level l >= 1 (CSR)
  int64_t n = 0, nnz = 0;      // owned rows, stored entries
  int64_t nloc = 0;            // owned + ghost unknowns (columns, vector length)
  HaloPlan halo;               // ghost exchange of this level's vectors (multi-GPU)
  DevBuf<int32_t> ptr, col;
  // Galerkin gather from level l-1 values
  DevBuf<int32_t> gptr, gidx; DevBuf<double> gw; bool g_unit = true;
  // transfer between l-1 (fine) and l (coarse)
  DevBuf<int32_t> pptr, pidx, rptr, ridx; DevBuf<double> pw, rw; bool t_unit = true;
};

struct LevelVectors { DevBuf<double> b, x, r, t; };   // work vectors of one level of one system

struct AmgPlan {
  bool ready = false;
  int64_t n0 = 0;
  std::vector<AmgLevelPlan> lev;   // lev[0] = level 1 ...
  // Multi-GPU: levels lev[0 .. rep_from-1] are distributed (owned rows + ghosts, halo per
  // sweep).  Once a level has few enough rows over all ranks, its matrix and right-hand side
  // are all-gathered and the REST of the hierarchy (lev[rep_from] = the global copy of
  // lev[rep_from-1], then serial coarsening down to the dense level) is built and applied
  // redundantly on every rank: one all-gather per cycle instead of two halos per level.
  size_t tail_from = (size_t)-1;   // first level handled by the single-launch coarse_tail_kernel; -1: none
  int tail_blocks = 0;             // its (cooperative) grid
  size_t rep_from = (size_t)-1;    // index of the first replicated level; -1: none
  int64_t rep_vstride = 0, rep_bstride = 0;   // padded per-rank segment: matrix values / rows
  DevBuf<double> rep_val;          // all-gather buffer of the matrix values [world * stride]
  HaloPlan rep_plan;               // the all-gather of the right-hand side as a peer-memory exchange
                                   // with every rank (self included): one kernel, usable from worker threads
  DevBuf<int32_t> rep_bmap;        // global row of the replica -> position in rep_b
  DevBuf<int32_t> rep_xmap;        // local unknown of lev[rep_from-1] -> global row of the replica
  int64_t m_dense = 0;             // rows of the last level (summed over all ranks)
  int64_t dense_off = 0;           // global index of this rank's first last-level row
  DevBuf<int32_t> dense_map;       // last level: local unknown (owned + ghost) -> global index


};

struct AmgValues {                 // numeric part, one per linear system
  std::vector<DevBuf<double>> val;   // per level >= 1
  std::vector<DevBuf<double>> dinv;  // l1-Jacobi diagonals
  DevBuf<double> dense;              // inverse of the last level
  DevBuf<double> binv;               // level-0 inverse diagonal blocks [nc][ND][ND]
  DevBuf<float> a32, binv32;         // single-precision copies of the level-0 matrix ((ND+1) slots) and of
                                     // binv, made at refresh, read by the level-0 sweeps (SolverOptions::pc_fp32)
  // work vectors (per system: independent systems are solved concurrently)
  std::vector<LevelVectors> vec;     // per level >= 1
  DevBuf<double> x0, r0, t0;         // level-0 work vectors
  DevBuf<double> colbuf;             // dense inverse scratch
  DevBuf<double> dense_b, dense_x;   // last level (distributed): global right-hand side / solution
  DevBuf<double> rep_b;              // replicated tail: all-gather buffer of the right-hand side
  double omega = 0.0;                // level-0 damping 4/(3 lambda_max(Dinv A)), estimated
  int age = 0;                       // refreshes since the last estimate
  bool refreshed_now = true;         // did the current solve refresh the values
  long long solves = 0;              // solves since the values were last refreshed (lagged refresh)
  int last_iters = 0, fresh_iters = 0;   // iterations of the last solve / of the solve right after a refresh
};

}  // namespace knp

// knp_ctx.h - the context object behind the C ABI (include/knpemi.h).
#pragma once
#include <chrono>
#include <memory>
#include "knp_common.h"
#include "knp_comm.h"
#include "knp_dg.h"
#include "knp_linalg.h"
#include "knp_amg.h"
#include "knp_ode.h"

namespace knp {

struct LinkSpec { int col, kind, which, idx, side; };

struct MembraneSet {
  int model_id = -1, ns = 0, np = 0;
  int64_t nrows = 0;
  DevBuf<int32_t> rows;
  DevBuf<double> states, params;
  DevBuf<uint8_t> mask;
  bool has_mask = false;
  std::vector<LinkSpec> links;
  int v_col = -1, n_ion = 0;
  int ich_cols[MAX_IONS] = {0};
  int nstim = 0; int stim_cols[MAX_STIM] = {0}; double stim_vals[MAX_STIM] = {0};
};

struct SolverOptions {
  int pc = 1;           // 0 block-Jacobi, 1 AMG
  int nu_pre = 1, nu_post = 1, gamma = 1;
  double omega = 0.0;   // level-0 block-Jacobi damping; <= 0: estimate 4/(3 lambda_max)
  int restart = 30, knp_min_it = 5;
  // V(0,1) on the DG level inside GMRES: 13 % faster time step at equal iteration counts on the
  // bench workload (profiles/); KNP_KNP_PRESMOOTH=1 in the environment restores V(1,1)
  bool knp_presmooth0 = false;
  // The level-0 sweeps of the PRECONDITIONER (block-Jacobi, residual inside the V-cycle) read single-precision
  // copies of the matrix and of the inverse diagonal blocks, made at every refresh; the Krylov operator, all
  // vectors and all accumulation stay fp64, the solves converge to the same tolerances in the same number of
  // iterations (measured on B200, 20.7 M DOFs: -6.7 % step time).  KNP_AMG_FP32=0 keeps fp64 copies.
  bool pc_fp32 = true;
  // KNP_AMG_CHEBY=2: degree-2 Chebyshev (block-Jacobi preconditioned) instead of one damped block-Jacobi
  // sweep before and after the coarse correction of the EMI V-cycle.  Halves the CG iteration count on
  // irregular meshes (two-level experiment) at 2 more level-0 sweeps per cycle; no gain
  // where the extrapolated initial guess already leaves 0-3 iterations (the bench workload)
  int cheby = 1;
  // initial guess of the EMI solve = 2 phi_n - phi_{n-1} instead of phi_n (the reference starts
  // from phi_n, solver.py:431 `ksp_initial_guess_nonzero`); same stopping test, fewer iterations.
  // KNP_EXTRAPOLATE=0 restores the reference's guess
  bool extrapolate_phi = true;
  // Preconditioner refresh: the reference rebuilds BoomerAMG at every solve (setOperators,
  // solver.py:505, 767).  With refresh_period = P > 1 the hierarchy VALUES (Galerkin products,
  // smoother diagonals, block inverses, dense inverse) are recomputed only every P-th solve of a
  // system, or earlier when the Krylov iteration count has grown by more than 50 % (+2) since
  // the last refresh; the Krylov solve itself still runs on the freshly assembled operator to
  // the reference's tolerance.  Default 8 (measured on B200, 81 M DOFs, unchanged iteration
  // counts: period 4 56.2 ms/step, 8 54.6, 16 53.5; round 1 measured -9.5 % for 1 -> 4);
  // KNP_AMG_REFRESH_PERIOD=1 in the environment restores the refresh at every solve.
  int refresh_period = 8;
};

enum { T_EMI_ASM = 0, T_EMI_SOLVE, T_KNP_ASM, T_KNP_SOLVE, T_ODE, T_POST, T_COUNT };

// Mutable state of ONE running Krylov solve (vectors, scalar scratch, the stream it is issued
// on).  The context owns one for the main stream and one per solved ion, so that the
// independent KNP systems can be solved concurrently on separate streams.
struct KrylovWs {
  knp_stream_t stream = 0;
  bool own_stream = false;
  int id = 0;                  // exchange channel of this workspace (knp_comm.h)
  DevBuf<double> r, p, q, w, V, scal, partial;
};

}  // namespace knp

struct knp_ctx {
  int device = 0;
  knp_stream_t stream = 0;
  // mesh
  int d = 0, nd = 0;
  int64_t nc = 0, n = 0, nm = 0, nnz_export = 0, nsip = 0;
  // multi-GPU: nc/n count the LOCAL cells/dofs (owned + ghost, the stride and size of every
  // per-cell array and vector); rows are assembled, solved and reduced over the owned ones
  int64_t nc_own = 0, n_own = 0;
  // owned cells 0 .. nc_int-1 have no ghost neighbour (knpemidg/partition.py numbers them first): their
  // rows of the level-0 operators run WHILE the halo of the input vector is in flight on comm_stream,
  // the remaining owned rows after it has landed (PETSc overlaps the VecScatter of MatMult the same way)
  int64_t nc_int = 0;
  bool overlap = false;
  knp_stream_t comm_stream = 0;
  double n_global = 0.0;        // owned dofs summed over all ranks
  knp::Comm comm;
  knp::HaloPlan halo0;          // level-0 (DG dof) halo
  std::vector<int32_t> h_nbr, h_finfo, h_fmem;                 // [nd][nc]
  std::vector<int32_t> h_mem_facet, h_mem_ci, h_mem_ce, h_mem_tag, h_mem_fi;
  knp::DevBuf<double> grad, vol, h;
  knp::DevBuf<double> fgeo;      // static facet geometry (knp_dg.h: FGeo), filled once by FacetGeomKernel
  knp::DevBuf<int32_t> region, nbr, finfo, fmem, mem_ci, mem_fi, memcell;
  int64_t nmc = 0;   // cells owning at least one membrane facet
  // parameters
  knp::Params P{};
  bool params_set = false;
  // fields
  knp::DevBuf<double> c[knp::MAX_IONS], cn_own[knp::MAX_IONS], phi, phiM;
  bool cn_separate[knp::MAX_IONS] = {false};
  knp::DevBuf<double> Ich[knp::MAX_IONS], E[knp::MAX_IONS];
  knp::DevBuf<double> rhs_emi, rhs_knp[knp::MAX_IONS], load_emi, load_knp[knp::MAX_IONS];
  bool has_load_emi = false, has_load_knp[knp::MAX_IONS] = {false};
  knp::DevBuf<double> kappa, q, gphi;
  knp::DevBuf<double> phi_old;   // potential of the previous time step (initial guess extrapolation)
  bool phi_old_valid = false;
  int phi_hist = 0;
  // matrices: A_emi = (nd+1) slots (slot 0 holds the diagonal blocks of B_emi, so the AMG
  // Galerkin plan addresses EMI and KNP matrices alike) followed by A_emi's own diagonal blocks
  knp::DevBuf<double> A_emi, A_knp[knp::MAX_IONS];
  bool emi_assembled = false, knp_assembled = false;
  // solver
  knp::SolverOptions opt;
  knp::KrylovWs kr0;                               // main-stream workspace
  knp::KrylovWs kr_ion[knp::MAX_IONS];             // Krylov vectors of the 2nd, 3rd ... system of the KNP batch
  knp::DevBuf<double> kr_ones;
  knp::AmgPlan amg;
  knp::AmgValues amg_emi, amg_knp[knp::MAX_IONS];
  knp::DevBuf<double> bj_emi, bj_knp[knp::MAX_IONS];   // block-Jacobi inverses
  // membranes
  std::vector<std::unique_ptr<knp::MembraneSet>> membranes;
  knp::DevBuf<int64_t> ode_stats;
  knp::DevBuf<double> trace_tmp;
  double timers[knp::T_COUNT] = {0};
#ifndef KNP_EMU
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_halo = nullptr;   // main -> comm stream / comm -> main stream
#endif
  double host_t0 = 0.0;

  const double* cn(int k) const { return cn_separate[k] ? cn_own[k].p : c[k].p; }
  int64_t bs() const { return (int64_t)nd * nd; }
  int64_t slot_stride() const { return nc * bs(); }
  double* Bdiag() { return A_emi.p; }
  double* Adiag_emi() { return A_emi.p + (int64_t)(nd + 1) * slot_stride(); }
};

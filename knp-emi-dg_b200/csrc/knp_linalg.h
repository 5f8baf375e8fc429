// knp_linalg.h - SpMV (block-ELL and CSR), smoothers, transfer operators, vector
// kernels and deterministic reductions.
//
// Replaces what the reference gets from PETSc (MatMult on AIJ, VecAXPY/VecDot inside
// KSPSolve, src/knpemidg/solver.py:509, 771) and from hypre BoomerAMG's cycle.
#pragma once
#include "knp_common.h"

namespace knp {

// A scalar matrix in block-ELL form; `diag` may point somewhere else than slot 0 of
// `off` (the EMI preconditioner matrix B shares A's off-diagonal blocks, solver.py:393-395).
// T = double everywhere except the optional single-precision COPY the preconditioner sweeps read
// (SolverOptions::pc_fp32): values are widened on load, all arithmetic stays fp64.
template <typename T>
struct BellMatT {
  int64_t nc = 0;
  const T* off = nullptr;        // (ND+1) slot arrays, slot 0 unused when diag != off
  const T* diag = nullptr;       // [nc][ND][ND]
  const int32_t* nbr = nullptr;  // [ND][nc], -1 = no coupling
};
using BellMat = BellMatT<double>;
using BellMat32 = BellMatT<float>;

struct DemoteKernel {  // out = (float) in
  const double* in; float* out;
  KNP_HD void operator()(int64_t i) const { out[i] = (float)in[i]; }
};

// ND consecutive matrix entries (one row of an ND x ND block) / ND consecutive vector entries (one
// cell) as doubles.  ND = 4 on the device: ONE 16-byte load for a float row, two for a double row or
// a cell of x (rows are 16 / 32 bytes and aligned; cudaMalloc'ed arrays are 256-byte aligned) - the
// fp32 preconditioner sweeps were issue-limited with scalar 4-byte loads (0.82 of the HBM peak).
template <int ND>
KNP_HD void load_row(const double* a, double (&out)[ND]) {
#if defined(__CUDA_ARCH__)
  if (ND == 4) {
    const double2 u = reinterpret_cast<const double2*>(a)[0], v = reinterpret_cast<const double2*>(a)[1];
    out[0] = u.x; out[1] = u.y; out[2] = v.x; out[ND - 1] = v.y;
    return;
  }
#endif
#pragma unroll
  for (int j = 0; j < ND; ++j) out[j] = a[j];
}
template <int ND>
KNP_HD void load_row(const float* a, double (&out)[ND]) {
#if defined(__CUDA_ARCH__)
  if (ND == 4) {
    const float4 u = *reinterpret_cast<const float4*>(a);
    out[0] = (double)u.x; out[1] = (double)u.y; out[2] = (double)u.z; out[ND - 1] = (double)u.w;
    return;
  }
#endif
#pragma unroll
  for (int j = 0; j < ND; ++j) out[j] = (double)a[j];
}

// y = A x (mode 0), y = b - A x (mode 1).  One thread per row; consecutive threads read
// consecutive ND-double row segments of every slot array (fully coalesced), x is
// gathered per neighbour cell (ND contiguous doubles, shared by the ND rows of a cell).
template <int ND, typename T = double>
struct BellSpmvKernel {
  BellMatT<T> A;
  const double* x; const double* b; double* y; int mode;
  KNP_HD void operator()(int64_t row) const {
    const int64_t cell = row / ND;
    const int i = (int)(row - cell * ND);
    const int64_t bs = ND * ND;
    double acc = 0.0;
    double av[ND], xv[ND];
    {
      load_row<ND>(A.diag + cell * bs + i * ND, av);
      load_row<ND>(x + cell * ND, xv);
#pragma unroll
      for (int j = 0; j < ND; ++j) acc += av[j] * xv[j];
    }
#pragma unroll
    for (int f = 0; f < ND; ++f) {
      const int64_t c2 = A.nbr[f * A.nc + cell];
      if (c2 < 0) continue;
      load_row<ND>(A.off + (int64_t)(1 + f) * A.nc * bs + cell * bs + i * ND, av);
      load_row<ND>(x + c2 * ND, xv);
#pragma unroll
      for (int j = 0; j < ND; ++j) acc += av[j] * xv[j];
    }
    y[row] = mode ? b[row] - acc : acc;
  }
};

// z = w * Dinv r (mode 0)   or   x += w * Dinv r (mode 1); Dinv = inverse diagonal blocks
template <int ND, typename T = double>
struct BlockDiagApplyKernel {
  const T* dinv; const double* r; double* out; double w; int mode;
  KNP_HD void operator()(int64_t row) const {
    const int64_t cell = row / ND;
    const int i = (int)(row - cell * ND);
    double av[ND], rv[ND];
    load_row<ND>(dinv + cell * ND * ND + i * ND, av);
    load_row<ND>(r + cell * ND, rv);
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j) acc += av[j] * rv[j];
    if (mode) out[row] += w * acc; else out[row] = w * acc;
  }
};

// fused block-Jacobi sweep on level 0: xout = xin + w Dinv (b - A xin).  One thread per
// row like the SpMV (coalesced slot reads); the ND residuals of a cell are exchanged
// between the ND lanes that own its rows with warp shuffles (ND = 4: groups of 4 lanes
// never straddle a warp; ND = 3: each lane recomputes the cell's other rows, the loads
// are warp broadcasts).
// MOM: second step of a degree-2 Chebyshev smoother, xout = xin + beta (xin - xprev) + w Dinv (b - A xin)
// (xprev == nullptr: zero; xout may alias xprev - a thread touches only its own row of both).
template <int ND, typename T = double, bool MOM = false>
struct BellJacobiKernel {
  BellMatT<T> A; const T* dinv; const double* b; const double* xin; double* xout; double w;
  const double* xprev = nullptr; double beta = 0.0;   // MOM only
  KNP_HD void cell_values(int64_t cell, double (&xv)[ND]) const { load_row<ND>(xin + cell * ND, xv); }
  KNP_HD double row_residual(int64_t cell, int i) const {
    const int64_t bs = ND * ND;
    double acc = b[cell * ND + i];
    double av[ND], xv[ND];
    {
      load_row<ND>(A.diag + cell * bs + i * ND, av);
      cell_values(cell, xv);
#pragma unroll
      for (int j = 0; j < ND; ++j) acc -= av[j] * xv[j];
    }
#pragma unroll
    for (int f = 0; f < ND; ++f) {
      const int64_t c2 = A.nbr[f * A.nc + cell];
      if (c2 < 0) continue;
      load_row<ND>(A.off + (int64_t)(1 + f) * A.nc * bs + cell * bs + i * ND, av);
      cell_values(c2, xv);
#pragma unroll
      for (int j = 0; j < ND; ++j) acc -= av[j] * xv[j];
    }
    return acc;
  }
  KNP_HD void operator()(int64_t row) const {
    const int64_t cell = row / ND;
    const int i = (int)(row - cell * ND);
    double r[ND];
#if defined(__CUDA_ARCH__)
    if (ND == 4) {
      const double mine = row_residual(cell, i);
      const unsigned mask = __activemask();
      const int base = (threadIdx.x & 31) & ~3;
#pragma unroll
      for (int j = 0; j < ND; ++j) r[j] = __shfl_sync(mask, mine, base + j);
    } else
#endif
    {
      for (int j = 0; j < ND; ++j) r[j] = row_residual(cell, j);
    }
    double dv[ND];
    load_row<ND>(dinv + cell * ND * ND + i * ND, dv);
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j) acc += dv[j] * r[j];
    if constexpr (MOM) {
      const double xv = xin[row];
      xout[row] = xv + beta * (xv - (xprev ? xprev[row] : 0.0)) + w * acc;
    } else {
      xout[row] = xin[row] + w * acc;
    }
  }
};

// ---- CSR (coarse levels, transfer operators) -------------------------------------
struct CsrMat {
  int64_t n = 0;
  const int32_t* ptr = nullptr; const int32_t* col = nullptr; const double* val = nullptr;
};

struct CsrSpmvKernel {  // y = A x | y = b - A x
  CsrMat A; const double* x; const double* b; double* y; int mode;
  KNP_HD void operator()(int64_t row) const {
    double acc = 0.0;
    for (int32_t k = A.ptr[row]; k < A.ptr[row + 1]; ++k) acc += A.val[k] * x[A.col[k]];
    y[row] = mode ? b[row] - acc : acc;
  }
};

struct CsrJacobiKernel {  // xout = xin + w dinv (b - A xin)
  CsrMat A; const double* dinv; const double* b; const double* xin; double* xout; double w;
  KNP_HD void operator()(int64_t row) const {
    double acc = b[row];
    for (int32_t k = A.ptr[row]; k < A.ptr[row + 1]; ++k) acc -= A.val[k] * xin[A.col[k]];
    xout[row] = xin[row] + w * dinv[row] * acc;
  }
};

struct DiagScaleKernel {  // out = w dinv b
  const double* dinv; const double* b; double* out; double w;
  KNP_HD void operator()(int64_t row) const { out[row] = w * dinv[row] * b[row]; }
};

// l1-Jacobi diagonal: dinv_i = 1 / sum_j |a_ij|  (always convergent for SPD, no
// eigenvalue estimate needed)
struct CsrL1DiagKernel {
  CsrMat A; double* dinv;
  KNP_HD void operator()(int64_t row) const {
    double s = 0.0;
    for (int32_t k = A.ptr[row]; k < A.ptr[row + 1]; ++k) s += fabs(A.val[k]);
    dinv[row] = s > 0.0 ? 1.0 / s : 0.0;
  }
};

// y (=|+=) T x for a sparse transfer operator in CSR with optional weights (nullptr = 1)
struct TransferKernel {
  int64_t n; const int32_t* ptr; const int32_t* idx; const double* w;
  const double* x; double* y; int add;
  KNP_HD void operator()(int64_t row) const {
    double acc = 0.0;
    if (w) for (int32_t k = ptr[row]; k < ptr[row + 1]; ++k) acc += w[k] * x[idx[k]];
    else   for (int32_t k = ptr[row]; k < ptr[row + 1]; ++k) acc += x[idx[k]];
    if (add) y[row] += acc; else y[row] = acc;
  }
};

// prolongation of a pure aggregation (one unit entry per fine row: (P xc)_i = xc[agg_i]): no row
// pointers to read, y (=|+=) xc[agg]
struct ProlongUnitKernel {
  const int32_t* agg; const double* xc; double* y; int add;
  KNP_HD void operator()(int64_t row) const {
    const double v = xc[agg[row]];
    if (add) y[row] += v; else y[row] = v;
  }
};

// Galerkin refresh with a frozen plan: coarse value k = sum_t w_t * fine[gidx_t]
struct GalerkinKernel {
  const int32_t* gptr; const int32_t* gidx; const double* gw; const double* fine; double* coarse;
  KNP_HD void operator()(int64_t k) const {
    double acc = 0.0;
    if (gw) for (int32_t t = gptr[k]; t < gptr[k + 1]; ++t) acc += gw[t] * fine[gidx[t]];
    else    for (int32_t t = gptr[k]; t < gptr[k + 1]; ++t) acc += fine[gidx[t]];
    coarse[k] = acc;
  }
};

// ---- sub-warp-per-row launcher --------------------------------------------------------
// Coarse AMG levels have 10^2..10^5 rows of ~25 entries: too few rows to hide the latency of
// a serial per-row loop, so LANES lanes share a row (strided partial sums, shuffle tree) and
// lane 0 finishes it.  Functors provide  double partial(row, lane, nlanes)  and
// void finish(row, sum).
#ifndef KNP_EMU
template <int LANES, class F>
__global__ void __launch_bounds__(256) subwarp_kernel(int64_t nrows, const F f) {
  const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = gt / LANES;
  const int lane = (int)(gt % LANES);
  const bool ok = row < nrows;
  double acc = ok ? f.partial(row, lane, LANES) : 0.0;
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LANES);
  if (ok && lane == 0) f.finish(row, acc);
}
#endif
#ifndef KNP_EMU
template <int LANES, class F>
__global__ void __launch_bounds__(256) subwarp_batch_kernel(int64_t nrows, const BatchOf<F> b) {
  const F& f = b.f[blockIdx.y];
  const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = gt / LANES;
  const int lane = (int)(gt % LANES);
  const bool ok = row < nrows;
  double acc = ok ? f.partial(row, lane, LANES) : 0.0;
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LANES);
  if (ok && lane == 0) f.finish(row, acc);
}
#endif
template <int LANES, class F>
inline void parallel_rows(knp_stream_t s, int64_t nrows, const F& f);
template <int LANES, class F>
inline void parallel_rows_batch(knp_stream_t s, int64_t nrows, int nb, const BatchOf<F>& b) {
  if (nrows <= 0 || nb <= 0) return;
#ifdef KNP_EMU
  for (int k = 0; k < nb; ++k) parallel_rows<LANES>(s, nrows, b.f[k]);
#else
  if (nb == 1) { parallel_rows<LANES>(s, nrows, b.f[0]); return; }
  const int64_t threads = nrows * LANES;
  ++launch_counter();
  subwarp_batch_kernel<LANES, F><<<dim3((unsigned)((threads + 255) / 256), (unsigned)nb), 256, 0, s>>>(nrows, b);
  KNP_CUDA(cudaGetLastError());
#endif
}
template <int LANES, class F>
inline void parallel_rows(knp_stream_t s, int64_t nrows, const F& f) {
  if (nrows <= 0) return;
#ifdef KNP_EMU
  (void)s;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (nrows > 2048)
#endif
  for (int64_t r = 0; r < nrows; ++r) f.finish(r, f.partial(r, 0, 1));
#else
  const int64_t threads = nrows * LANES;
  ++launch_counter();
  subwarp_kernel<LANES, F><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(nrows, f);
  KNP_CUDA(cudaGetLastError());
#endif
}

// ---- fused coarse-level sweeps (unit aggregation transfer, V(1,1) cycle) ----------------
// down, pass 1: x = dinv b (pre-smoothing from a zero guess) and r = b - A x in one pass
struct CoarseResidualKernel {
  CsrMat A; const double* dinv; const double* b; double* x; double* r;
  KNP_HD double partial(int64_t i, int lane, int nl) const {
    double acc = 0.0;
    for (int32_t k = A.ptr[i] + lane; k < A.ptr[i + 1]; k += nl) {
      const int32_t j = A.col[k];
      acc += A.val[k] * dinv[j] * b[j];
    }
    return acc;
  }
  KNP_HD void finish(int64_t i, double sum) const {
    const double bi = b[i];
    x[i] = dinv[i] * bi;
    r[i] = bi - sum;
  }
};

// restriction / any CSR transfer with optional weights: y (=|+=) T x
struct TransferRowsKernel {
  const int32_t* ptr; const int32_t* idx; const double* w; const double* x; double* y; int add;
  KNP_HD double partial(int64_t row, int lane, int nl) const {
    double acc = 0.0;
    if (w) for (int32_t k = ptr[row] + lane; k < ptr[row + 1]; k += nl) acc += w[k] * x[idx[k]];
    else   for (int32_t k = ptr[row] + lane; k < ptr[row + 1]; k += nl) acc += x[idx[k]];
    return acc;
  }
  KNP_HD void finish(int64_t row, double sum) const { if (add) y[row] += sum; else y[row] = sum; }
};

// up: x' = x + P xc (prolongation), xout = x' + dinv (b - A x') (post-smoothing) in ONE pass;
// P has exactly one unit entry per row: (P xc)_j = xc[agg_j].
struct CoarseUpKernel {
  CsrMat A; const double* dinv; const double* b; const double* x; const int32_t* agg;
  const double* xc; double* xout;
  KNP_HD double partial(int64_t i, int lane, int nl) const {
    double acc = 0.0;
    for (int32_t k = A.ptr[i] + lane; k < A.ptr[i + 1]; k += nl) {
      const int32_t j = A.col[k];
      acc += A.val[k] * (x[j] + xc[agg[j]]);
    }
    return acc;
  }
  KNP_HD void finish(int64_t i, double sum) const {
    xout[i] = x[i] + xc[agg[i]] + dinv[i] * (b[i] - sum);
  }
};

// dense[m*m] (zeroed before) <- the owned rows of the last level's CSR; `map` (nullptr =
// identity) sends a local unknown to its index in the global last-level numbering, `off` is
// the global index of local row 0.
// ---- the small coarse levels in ONE launch ------------------------------------------------
// Levels with 10^1..10^4 rows cost 5-7 us per launch (launch + drain) and three launches per
// level and cycle; together they were 29 % of a time step (profiles/launches_r01_step.md).
// coarse_tail_kernel runs the whole fused V(1,1) cycle of the smallest levels - sweep down
// (smooth + residual, restrict), dense solve, sweep up (prolong + smooth) - as ONE launch of a
// single thread-block cluster (8 CTAs x 1024 threads on 8 SMs) with the hardware cluster
// barrier (~0.2 us, B300_MICROARCH.md) between the phases.  (A cooperative whole-grid version
// with grid.sync() was measured SLOWER than the launches it replaced: ~5 us per grid barrier.)  The arithmetic (lane assignment,
// reduction order) is exactly that of CoarseResidualKernel / TransferRowsKernel /
// DenseMatvecKernel / CoarseUpKernel above, so the result is bit-identical to the
// launch-per-sweep path (which the host emulation and larger levels keep using).
constexpr int TAIL_MAX_LEVELS = 12;
constexpr int TAIL_MAX_ROWS = 4096;
constexpr int TAIL_CTAS = 8, TAIL_THREADS = 1024;
struct TailLevel {
  int64_t n;
  const int32_t* ptr; const int32_t* col; const double* val; const double* dinv;
  double* b; double* x; double* r;
  const int32_t* rptr; const int32_t* ridx;   // restriction INTO this level from the finer tail level
  const int32_t* agg;                         // finer tail level's unknown -> unknown of this level
};
struct TailArgs {
  int nlev;                // L[nlev-1] is the dense level
  TailLevel L[TAIL_MAX_LEVELS];
  const double* denseT;    // transposed inverse of the last level
};

#ifndef KNP_EMU
}  // namespace knp
#include <cooperative_groups.h>
namespace knp {
struct TailBatch { TailArgs s[MAX_BATCH]; };
static __global__ void __cluster_dims__(TAIL_CTAS, 1, 1) __launch_bounds__(TAIL_THREADS)
coarse_tail_kernel(const TailBatch batch) {
  namespace cg = cooperative_groups;
  cg::cluster_group grid = cg::this_cluster();   // one cluster per linear system of the batch
  constexpr int LANES = 8;
  const TailArgs& a = batch.s[blockIdx.x / TAIL_CTAS];
  const int64_t gt = (blockIdx.x % TAIL_CTAS) * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)TAIL_CTAS * blockDim.x;
  const int lane = (int)(gt % LANES);
  const int64_t nsub = nthreads / LANES;       // rows processed per pass
  const int64_t sub = gt / LANES;
  // sweep down
  for (int k = 0; k + 1 < a.nlev; ++k) {
    const TailLevel& L = a.L[k];
    for (int64_t base = 0; base < L.n; base += nsub) {
      const int64_t i = base + sub;
      const bool ok = i < L.n;
      double acc = 0.0;
      if (ok)
        for (int32_t e = L.ptr[i] + lane; e < L.ptr[i + 1]; e += LANES) {
          const int32_t j = L.col[e];
          acc += L.val[e] * L.dinv[j] * L.b[j];
        }
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LANES);
      if (ok && lane == 0) L.r[i] = L.b[i] - acc;
    }
    grid.sync();
    const TailLevel& C = a.L[k + 1];
    for (int64_t base = 0; base < C.n; base += nsub) {
      const int64_t I = base + sub;
      const bool ok = I < C.n;
      double acc = 0.0;
      if (ok)
        for (int32_t e = C.rptr[I] + lane; e < C.rptr[I + 1]; e += LANES) acc += L.r[C.ridx[e]];
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LANES);
      if (ok && lane == 0) C.b[I] = acc;
    }
    grid.sync();
  }
  // dense level
  {
    const TailLevel& D = a.L[a.nlev - 1];
    const int64_t m = D.n;
    for (int64_t row = gt; row < m; row += nthreads) {
      double acc = 0.0;
      for (int64_t j = 0; j < m; ++j) acc += a.denseT[j * m + row] * D.b[j];
      D.x[row] = acc;
    }
    grid.sync();
  }
  // sweep up
  for (int k = a.nlev - 2; k >= 0; --k) {
    const TailLevel& L = a.L[k];
    const TailLevel& C = a.L[k + 1];
    for (int64_t base = 0; base < L.n; base += nsub) {
      const int64_t i = base + sub;
      const bool ok = i < L.n;
      double acc = 0.0;
      if (ok)
        for (int32_t e = L.ptr[i] + lane; e < L.ptr[i + 1]; e += LANES) {
          const int32_t j = L.col[e];
          acc += L.val[e] * (__dmul_rn(L.dinv[j], L.b[j]) + C.x[C.agg[j]]);
        }
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LANES);
      if (ok && lane == 0) L.x[i] = __dmul_rn(L.dinv[i], L.b[i]) + C.x[C.agg[i]] + L.dinv[i] * (L.b[i] - acc);
    }
    if (k > 0) grid.sync();
  }
}
#endif

struct CsrToDenseKernel {
  CsrMat A; double* dense; int64_t m; int64_t off; const int32_t* map;
  KNP_HD void operator()(int64_t row) const {
    for (int32_t k = A.ptr[row]; k < A.ptr[row + 1]; ++k) {
      const int64_t col = map ? map[A.col[k]] : A.col[k];
      dense[(off + row) * m + col] = A.val[k];
    }
  }
};
struct ScatterOffsetKernel {  // out[off + i] = x[i]
  const double* x; double* out; int64_t off;
  KNP_HD void operator()(int64_t i) const { out[off + i] = x[i]; }
};
struct GatherMapKernel {      // out[i] = x[map[i]]
  const double* x; const int32_t* map; double* out;
  KNP_HD void operator()(int64_t i) const { out[i] = x[map[i]]; }
};

struct DenseMatvecKernel {  // y = M x with M stored TRANSPOSED (MT[j*m + i] = M[i][j])
  int64_t m; const double* MT; const double* x; double* y;
  KNP_HD void operator()(int64_t row) const {
    double acc = 0.0;
    for (int64_t j = 0; j < m; ++j) acc += MT[j * m + row] * x[j];
    y[row] = acc;
  }
};

// ---- vector kernels ---------------------------------------------------------------
struct ExtrapolateKernel {  // x <- 2 x - old, old <- x   (linear extrapolation in time of the initial guess)
  double* x; double* old;
  KNP_HD void operator()(int64_t i) const { const double t = x[i]; x[i] = 2.0 * t - old[i]; old[i] = t; }
};
struct DirectionKernel {  // p = (z - mu) + b p   (CG direction from the mean-free preconditioned residual)
  const double* z; double mu; double b; double* p;
  KNP_HD void operator()(int64_t i) const { p[i] = (b == 0.0) ? (z[i] - mu) : (z[i] - mu) + b * p[i]; }
};
struct Axpy2ProjKernel {  // x += a p ; r = r - a q - shift   (CG update; shift keeps r mean-free)
  double a; const double* p; const double* q; double* x; double* r; double shift;
  KNP_HD void operator()(int64_t i) const { x[i] += a * p[i]; r[i] = (r[i] - a * q[i]) - shift; }
};
struct ScaleKernel {  // y = a x
  double a; const double* x; double* y;
  KNP_HD void operator()(int64_t i) const { y[i] = a * x[i]; }
};
struct AddConstKernel {  // x += a
  double a; double* x;
  KNP_HD void operator()(int64_t i) const { x[i] += a; }
};
// Gram-Schmidt update and normalisation in one pass: v = (v - sum_i h_i V_i) / hn with
// hn^2 = h[k] - sum_i h_i^2, where h[k] = |v|^2 before the update (Pythagoras: one reduction
// per Arnoldi step).  When the subtraction cancels too many digits (hn^2 < tol * |v|^2) the
// vector is left unnormalised and the caller computes the norm explicitly.
struct GsNormalizeKernel {
  int64_t n /*stride of V*/; int k; const double* V; const double* h /*device, k+1 entries*/; double* v; double tol;
  KNP_HD void operator()(int64_t e) const {
    double acc = v[e], s = 0.0;
    for (int i = 0; i < k; ++i) { acc -= h[i] * V[(int64_t)i * n + e]; s += h[i] * h[i]; }
    const double hn2 = h[k] - s;
    v[e] = (hn2 >= tol * h[k] && hn2 > 0.0) ? acc / sqrt(hn2) : acc;
  }
};
// x += sum_i y_i V_i
struct CombineKernel {
  int64_t n; int k; const double* V; const double* yv /*device*/; double* x;
  KNP_HD void operator()(int64_t e) const {
    double acc = x[e];
    for (int i = 0; i < k; ++i) acc += yv[i] * V[(int64_t)i * n + e];
    x[e] = acc;
  }
};

// ---- deterministic reductions -------------------------------------------------------
// out[i] = sum_e V[i*n + e] * w[e], i < k <= DOT_MAX per call.  Two passes with a fixed
// grid: per-block partials, then one block sums them in a fixed order (bit-reproducible
// for a given n; no atomics).
constexpr int DOT_MAX = 8;
constexpr int RED_BLOCKS = 592;   // 4 x 148 SMs
constexpr int RED_THREADS = 256;
// (the sums run over the `nb` data sets of a batch: the ions' systems are ONE block system)
struct DotBatch { int nb; const double* V[MAX_BATCH]; const double* w[MAX_BATCH]; };

#ifndef KNP_EMU
template <int K>
__global__ void __launch_bounds__(RED_THREADS) multi_dot_partial(int64_t n, int64_t stride, const DotBatch B,
                                                                 double* __restrict__ partial) {
  double acc[K];
#pragma unroll
  for (int i = 0; i < K; ++i) acc[i] = 0.0;
  for (int s = 0; s < B.nb; ++s) {
    const double* __restrict__ V = B.V[s];
    const double* __restrict__ w = B.w[s];
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
      const double we = w[e];
#pragma unroll
      for (int i = 0; i < K; ++i) acc[i] += V[(int64_t)i * stride + e] * we;
    }
  }
  __shared__ double sm[K][RED_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double v = 0.0;
    for (int wv = 0; wv < RED_THREADS / 32; ++wv) v += sm[threadIdx.x][wv];
    partial[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = v;
  }
}

static __global__ void __launch_bounds__(RED_THREADS) multi_dot_final(int k, int nblocks,
                                                               const double* __restrict__ partial,
                                                               double* __restrict__ out) {
  __shared__ double sm[RED_THREADS / 32];
  for (int i = 0; i < k; ++i) {
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) v += partial[(int64_t)i * nblocks + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int wv = 0; wv < RED_THREADS / 32; ++wv) s += sm[wv];
      out[i] = s;
    }
    __syncthreads();
  }
}
#endif

// up to 4 independent dot products a_i . b_i in one pass (CG: r.z, z.z, 1.z, 1.r)
struct DotPairs { const double* a[4]; const double* b[4]; };
#ifndef KNP_EMU
template <int K>
__global__ void __launch_bounds__(RED_THREADS) pair_dot_partial(int64_t n, const DotPairs P,
                                                                double* __restrict__ partial) {
  double acc[K];
#pragma unroll
  for (int i = 0; i < K; ++i) acc[i] = 0.0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] += P.a[i][e] * P.b[i][e];
  }
  __shared__ double sm[K][RED_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double v = 0.0;
    for (int wv = 0; wv < RED_THREADS / 32; ++wv) v += sm[threadIdx.x][wv];
    partial[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = v;
  }
}
#endif

// device-side dots over the first n entries of k vectors `stride` apart; `out` (device,
// >= k doubles) receives the results, `partial` is a scratch buffer of DOT_MAX*RED_BLOCKS doubles.
inline void multi_dot_batch(knp_stream_t s, int64_t n, int64_t stride, int k, const DotBatch& B0,
                            double* partial, double* out) {
  for (int base = 0; base < k; base += DOT_MAX) {
    const int kk = (k - base < DOT_MAX) ? k - base : DOT_MAX;
    DotBatch B = B0;
    for (int b = 0; b < B.nb; ++b) B.V[b] = B0.V[b] + (int64_t)base * stride;
#ifdef KNP_EMU
    (void)s; (void)partial;
    for (int i = 0; i < kk; ++i) {
      double acc = 0.0;
      for (int b = 0; b < B.nb; ++b) {
        const double* Vb = B.V[b];
        const double* w = B.w[b];
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : acc) if (n > 2048)
#endif
        for (int64_t e = 0; e < n; ++e) acc += Vb[(int64_t)i * stride + e] * w[e];
      }
      out[base + i] = acc;
    }
#else
    switch (kk) {
#define KNP_CASE(K) case K: multi_dot_partial<K><<<RED_BLOCKS, RED_THREADS, 0, s>>>(n, stride, B, partial); break;
      KNP_CASE(1) KNP_CASE(2) KNP_CASE(3) KNP_CASE(4) KNP_CASE(5) KNP_CASE(6) KNP_CASE(7) KNP_CASE(8)
#undef KNP_CASE
    }
    KNP_CUDA(cudaGetLastError());
    launch_counter() += 2;
    multi_dot_final<<<1, RED_THREADS, 0, s>>>(kk, RED_BLOCKS, partial, out + base);
    KNP_CUDA(cudaGetLastError());
#endif
  }
}
inline void multi_dot_device(knp_stream_t s, int64_t n, int64_t stride, int k, const double* V, const double* w,
                             double* partial, double* out) {
  DotBatch B;
  B.nb = 1; B.V[0] = V; B.w[0] = w;
  multi_dot_batch(s, n, stride, k, B, partial, out);
}

inline void pair_dot_device(knp_stream_t s, int64_t n, int k, const DotPairs& P, double* partial, double* out) {
#ifdef KNP_EMU
  (void)s; (void)partial;
  for (int i = 0; i < k; ++i) {
    double acc = 0.0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : acc) if (n > 2048)
#endif
    for (int64_t e = 0; e < n; ++e) acc += P.a[i][e] * P.b[i][e];
    out[i] = acc;
  }
#else
  switch (k) {
    case 1: pair_dot_partial<1><<<RED_BLOCKS, RED_THREADS, 0, s>>>(n, P, partial); break;
    case 2: pair_dot_partial<2><<<RED_BLOCKS, RED_THREADS, 0, s>>>(n, P, partial); break;
    case 3: pair_dot_partial<3><<<RED_BLOCKS, RED_THREADS, 0, s>>>(n, P, partial); break;
    default: pair_dot_partial<4><<<RED_BLOCKS, RED_THREADS, 0, s>>>(n, P, partial); break;
  }
  KNP_CUDA(cudaGetLastError());
  launch_counter() += 2;
  multi_dot_final<<<1, RED_THREADS, 0, s>>>(k, RED_BLOCKS, partial, out);
  KNP_CUDA(cudaGetLastError());
#endif
}

// inverse of the dense m x m coarsest-level matrix by Gauss-Jordan without pivoting (SPD
// for EMI, diagonally dominant for KNP); the result is stored TRANSPOSED for
// DenseMatvecKernel.  One thread block; the matrix lives in shared memory (odd pitch, no
// bank conflicts on column access) when it fits (m <= 168), else in global memory.
constexpr int DENSE_SMEM_MAX = 168;
#ifndef KNP_EMU
static __global__ void __launch_bounds__(1024) dense_inverse_smem_kernel(int m, double* __restrict__ A) {
  extern __shared__ double sm[];
  const int pitch = m | 1;
  double* a = sm;
  double* colbuf = sm + (size_t)m * pitch;
  for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
    const int i = idx / m, j = idx - i * m;
    a[i * pitch + j] = A[idx];
  }
  __syncthreads();
  for (int p = 0; p < m; ++p) {
    const double ip = 1.0 / a[p * pitch + p];
    for (int i = threadIdx.x; i < m; i += blockDim.x) colbuf[i] = a[i * pitch + p];
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x)
      a[p * pitch + j] = (j == p) ? ip : a[p * pitch + j] * ip;
    __syncthreads();
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
      const int i = idx / m, j = idx - i * m;
      if (i == p) continue;
      const double f = colbuf[i];
      const double prow = a[p * pitch + j];
      a[i * pitch + j] = (j == p) ? -f * prow : a[i * pitch + j] - f * prow;
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
    const int i = idx / m, j = idx - i * m;
    A[(size_t)j * m + i] = a[i * pitch + j];
  }
}

static __global__ void __launch_bounds__(1024) dense_inverse_kernel(int m, double* __restrict__ A,
                                                             double* __restrict__ colbuf) {
  for (int p = 0; p < m; ++p) {
    const double ip = 1.0 / A[(int64_t)p * m + p];
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) colbuf[i] = A[(int64_t)i * m + p];
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x)
      A[(int64_t)p * m + j] = (j == p) ? ip : A[(int64_t)p * m + j] * ip;
    __syncthreads();
    for (int64_t idx = threadIdx.x; idx < (int64_t)m * m; idx += blockDim.x) {
      const int i = (int)(idx / m), j = (int)(idx - (int64_t)i * m);
      if (i == p) continue;
      const double f = colbuf[i];
      const double prow = A[(int64_t)p * m + j];
      A[idx] = (j == p) ? -f * prow : A[idx] - f * prow;
    }
    __syncthreads();
  }
  // transpose in place
  for (int64_t idx = threadIdx.x; idx < (int64_t)m * m; idx += blockDim.x) {
    const int i = (int)(idx / m), j = (int)(idx - (int64_t)i * m);
    if (i < j) { const double t = A[idx]; A[idx] = A[(int64_t)j * m + i]; A[(int64_t)j * m + i] = t; }
  }
}
#endif

inline void dense_inverse_device(knp_stream_t s, int m, double* A, double* colbuf) {
#ifdef KNP_EMU
  (void)s;
  for (int p = 0; p < m; ++p) {
    const double ip = 1.0 / A[(int64_t)p * m + p];
    for (int i = 0; i < m; ++i) colbuf[i] = A[(int64_t)i * m + p];
    for (int j = 0; j < m; ++j) A[(int64_t)p * m + j] = (j == p) ? ip : A[(int64_t)p * m + j] * ip;
    for (int i = 0; i < m; ++i) {
      if (i == p) continue;
      const double f = colbuf[i];
      for (int j = 0; j < m; ++j) {
        const double prow = A[(int64_t)p * m + j];
        A[(int64_t)i * m + j] = (j == p) ? -f * prow : A[(int64_t)i * m + j] - f * prow;
      }
    }
  }
  for (int i = 0; i < m; ++i)
    for (int j = i + 1; j < m; ++j) {
      const double t = A[(int64_t)i * m + j];
      A[(int64_t)i * m + j] = A[(int64_t)j * m + i];
      A[(int64_t)j * m + i] = t;
    }
#else
  ++launch_counter();
  if (m <= DENSE_SMEM_MAX) {
    const size_t bytes = ((size_t)m * (m | 1) + m) * sizeof(double);
    static bool configured = false;
    if (!configured) {
      KNP_CUDA(cudaFuncSetAttribute(dense_inverse_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(((size_t)DENSE_SMEM_MAX * (DENSE_SMEM_MAX | 1) + DENSE_SMEM_MAX) * sizeof(double))));
      configured = true;
    }
    dense_inverse_smem_kernel<<<1, 1024, bytes, s>>>(m, A);
  } else {
    dense_inverse_kernel<<<1, 1024, 0, s>>>(m, A, colbuf);
  }
  KNP_CUDA(cudaGetLastError());
#endif
}

}  // namespace knp

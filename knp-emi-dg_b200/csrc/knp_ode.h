// knp_ode.h - one-thread-per-membrane-facet ODE step with the PDE<->ODE transfer fused in.
//
// Replaces MembraneModel.step_lsoda (src/knpemidg/membrane.py:84-119: a Python loop
// calling numbalsoda's LSODA once per facet with rtol 1e-8, atol 0 and a fresh
// integrator per call) together with the gathers/scatters around it
// (solver.py:1094-1113, membrane.py:122-162).
//
// Integrator: explicit Dormand-Prince 5(4) pair with FSAL, PI-free step control on the
// weighted max norm  |err_i| / (atol + rtol*max(|y_i|,|ynew_i|))  (the norm LSODA uses),
// automatic initial step, last step clipped to land exactly on t0+dt.  The membrane
// models of the reference are non-stiff at the PDE step sizes used (dt = 0.1 ms vs. gate
// time constants >= 0.1 ms), so LSODA itself stays in its Adams mode there.
// Stiff problems (LSODA would switch to BDF, membrane.py:108-112): when the explicit pair has
// spent STIFF_SWITCH steps inside one interval its step is stability-, not accuracy-limited; the
// rest of the interval is integrated by an L-stable linearly implicit Rosenbrock pair (Shampine &
// Reichelt's ode23s: order 2 with a 3rd-order error estimate, finite-difference Jacobian, one
// LU factorisation per step) under the same error norm and tolerances.
//
// Channel currents: the model right-hand side stores I_ch_* into the parameter row as a
// side effect (e.g. examples/idealized-geometries/mm_hh.py:154-159).  After the last
// accepted step the right-hand side is evaluated once more at (t0+dt, y(t0+dt)), so the
// currents handed to the PDEs are I(y(t0+dt)) ("end_state" convention).
#pragma once
#include "knp_common.h"
#include "knp_dg.h"

namespace knp {

constexpr int MAX_LINKS = 12;
constexpr int MAX_STIM = 4;

KNP_HD double knp_pymod(double a, double b) {  // Python % / numpy.mod semantics
  double r = fmod(a, b);
  if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
  return r;
}
KNP_HD double knp_ipow1(double x) { return x; }
KNP_HD double knp_ipow2(double x) { return x * x; }
KNP_HD double knp_ipow3(double x) { return x * x * x; }
KNP_HD double knp_ipow4(double x) { const double s = x * x; return s * s; }

struct OdeLink {
  int col;            // parameter column that receives the value
  int kind;           // 0: membrane-row array src[mrow]; 1: facet mean of side trace of cell field
  int side;           // kind 1: 0 plus/ECS, 1 minus/ICS
  const double* src;
};

constexpr int STIFF_SWITCH = 4000;   // explicit steps inside one interval before the implicit pair takes over

// Linearly implicit Rosenbrock pair (ode23s) from t to t1, in place.  Kept out of line: its
// Jacobian and LU factors live in local memory and must not enlarge the frame of the common path.
template <class M>
#if defined(__CUDACC__) && !defined(KNP_EMU)
__device__ __noinline__
#else
inline
#endif
void ros23s(double t, double t1, double* y, double* p, double rtol, double atol, int& nsteps, int& nfev, int& ok) {
  constexpr int NS = M::NS;
  const double tiny = 1e-300;
  const double d = 1.0 / (2.0 + 1.4142135623730951), e32 = 6.0 + 1.4142135623730951;
  const double span = t1 - t;
  double J[NS][NS], W[NS][NS], F0[NS], F1[NS], F2[NS], T[NS], k1[NS], k2[NS], k3[NS], yt[NS];
  int piv[NS];
  auto solve = [&](double* b) {            // W x = b with the stored LU factors (partial pivoting)
    for (int i = 0; i < NS; ++i) { const double tmp = b[i]; b[i] = b[piv[i]]; b[piv[i]] = tmp;
      for (int j = 0; j < i; ++j) b[i] -= W[i][j] * b[j]; }
    for (int i = NS - 1; i >= 0; --i) { for (int j = i + 1; j < NS; ++j) b[i] -= W[i][j] * b[j]; b[i] /= W[i][i]; }
  };
  double h = fmin(span, fmax(1e-6 * span, 1e-3 * span));
  const int max_steps = 400000;
  while (t < t1) {
    if (nsteps >= max_steps) { ok = 0; return; }
    M::rhs(t, y, F0, p); nfev++;
    // finite-difference Jacobian and time derivative
    for (int j = 0; j < NS; ++j) {
      const double yj = y[j];
      const double dy = 1.4901161193847656e-08 * fmax(fabs(yj), 1e-6);
      y[j] = yj + dy;
      M::rhs(t, y, F1, p); nfev++;
      y[j] = yj;
      for (int i = 0; i < NS; ++i) J[i][j] = (F1[i] - F0[i]) / dy;
    }
    {
      const double dtt = 1.4901161193847656e-08 * fmax(fabs(t), span);
      M::rhs(t + dtt, y, F1, p); nfev++;
      for (int i = 0; i < NS; ++i) T[i] = (F1[i] - F0[i]) / dtt;
    }
    for (;;) {   // retry with a smaller step until accepted
      bool final_step = false;
      if (t + h >= t1 || t1 - (t + h) < 1e-12 * span) { h = t1 - t; final_step = true; }
      for (int i = 0; i < NS; ++i)
        for (int j = 0; j < NS; ++j) W[i][j] = ((i == j) ? 1.0 : 0.0) - h * d * J[i][j];
      bool singular = false;
      for (int c = 0; c < NS; ++c) {          // LU with partial pivoting, row swaps recorded in piv
        int r = c;
        for (int i = c + 1; i < NS; ++i) if (fabs(W[i][c]) > fabs(W[r][c])) r = i;
        piv[c] = r;
        if (r != c) for (int j = 0; j < NS; ++j) { const double tmp = W[c][j]; W[c][j] = W[r][j]; W[r][j] = tmp; }
        if (W[c][c] == 0.0) { singular = true; break; }
        for (int i = c + 1; i < NS; ++i) {
          W[i][c] /= W[c][c];
          for (int j = c + 1; j < NS; ++j) W[i][j] -= W[i][c] * W[c][j];
        }
      }
      double err = 2.0;
      if (!singular) {
        for (int i = 0; i < NS; ++i) k1[i] = F0[i] + h * d * T[i];
        solve(k1);
        for (int i = 0; i < NS; ++i) yt[i] = y[i] + 0.5 * h * k1[i];
        M::rhs(t + 0.5 * h, yt, F1, p); nfev++;
        for (int i = 0; i < NS; ++i) k2[i] = F1[i] - k1[i];
        solve(k2);
        for (int i = 0; i < NS; ++i) { k2[i] += k1[i]; yt[i] = y[i] + h * k2[i]; }
        M::rhs(t + h, yt, F2, p); nfev++;
        for (int i = 0; i < NS; ++i) k3[i] = F2[i] - e32 * (k2[i] - F1[i]) - 2.0 * (k1[i] - F0[i]) + h * d * T[i];
        solve(k3);
        err = 0.0;
        for (int i = 0; i < NS; ++i) {
          const double sc = fmax(atol + rtol * fmax(fabs(y[i]), fabs(yt[i])), tiny);
          err = fmax(err, fabs(h / 6.0 * (k1[i] - 2.0 * k2[i] + k3[i])) / sc);
        }
      }
      nsteps++;
      if (err <= 1.0) {
        t = final_step ? t1 : t + h;
        for (int i = 0; i < NS; ++i) y[i] = yt[i];
        h *= (err < 1e-9) ? 5.0 : fmin(5.0, 0.8 * pow(err, -1.0 / 3.0));
        break;
      }
      if (!(err == err) || h < 1e-14 * span || nsteps >= max_steps) { ok = 0; return; }
      h *= fmax(0.2, 0.8 * pow(err, -1.0 / 3.0));
    }
  }
}

template <class M>
KNP_HD void dopri5(double t0, double t1, double* y, double* p, double rtol, double atol,
                   int& nsteps, int& nfev, int& ok) {
  constexpr int NS = M::NS;
  const double c2 = 1.0 / 5, c3 = 3.0 / 10, c4 = 4.0 / 5, c5 = 8.0 / 9;
  const double a21 = 1.0 / 5;
  const double a31 = 3.0 / 40, a32 = 9.0 / 40;
  const double a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9;
  const double a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729;
  const double a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176,
               a65 = -5103.0 / 18656;
  const double a71 = 35.0 / 384, a73 = 500.0 / 1113, a74 = 125.0 / 192, a75 = -2187.0 / 6784,
               a76 = 11.0 / 84;
  const double e1 = 71.0 / 57600, e3 = -71.0 / 16695, e4 = 71.0 / 1920, e5 = -17253.0 / 339200,
               e6 = 22.0 / 525, e7 = -1.0 / 40;
  const double tiny = 1e-300;
  double k1[NS], k2[NS], k3[NS], k4[NS], k5[NS], k6[NS], k7[NS], yt[NS];
  const double span = t1 - t0;
  double t = t0;
  nsteps = 0; nfev = 0; ok = 1;
  M::rhs(t, y, k1, p); nfev++;
  // initial step (Hairer, Norsett, Wanner II.4)
  double h;
  {
    double d0 = 0.0, d1 = 0.0;
    for (int i = 0; i < NS; ++i) {
      const double sc = fmax(atol + rtol * fabs(y[i]), tiny);
      d0 = fmax(d0, fabs(y[i]) / sc);
      d1 = fmax(d1, fabs(k1[i]) / sc);
    }
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 * span : 0.01 * d0 / d1;
    h0 = fmin(h0, span);
    for (int i = 0; i < NS; ++i) yt[i] = y[i] + h0 * k1[i];
    M::rhs(t + h0, yt, k2, p); nfev++;
    double d2 = 0.0;
    for (int i = 0; i < NS; ++i) {
      const double sc = fmax(atol + rtol * fabs(y[i]), tiny);
      d2 = fmax(d2, fabs(k2[i] - k1[i]) / sc);
    }
    d2 /= h0;
    const double dm = fmax(d1, d2);
    const double h1 = (dm <= 1e-15) ? fmax(1e-6 * span, h0 * 1e-3) : pow(0.01 / dm, 0.2);
    h = fmin(fmin(100.0 * h0, h1), span);
  }
  const int max_steps = 200000;
  bool last_rejected = false;
  while (t < t1) {
    if (nsteps >= max_steps) { ok = 0; break; }
    if (nsteps >= STIFF_SWITCH) {          // stability-limited: hand the rest of the interval to the implicit pair
      ros23s<M>(t, t1, y, p, rtol, atol, nsteps, nfev, ok);
      if (ok) ok = 2;                      // (statistics: this facet needed the stiff path)
      break;
    }
    bool final_step = false;
    if (t + h >= t1 || t1 - (t + h) < 1e-12 * span) { h = t1 - t; final_step = true; }
    for (int i = 0; i < NS; ++i) yt[i] = y[i] + h * a21 * k1[i];
    M::rhs(t + c2 * h, yt, k2, p);
    for (int i = 0; i < NS; ++i) yt[i] = y[i] + h * (a31 * k1[i] + a32 * k2[i]);
    M::rhs(t + c3 * h, yt, k3, p);
    for (int i = 0; i < NS; ++i) yt[i] = y[i] + h * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
    M::rhs(t + c4 * h, yt, k4, p);
    for (int i = 0; i < NS; ++i)
      yt[i] = y[i] + h * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
    M::rhs(t + c5 * h, yt, k5, p);
    for (int i = 0; i < NS; ++i)
      yt[i] = y[i] + h * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
    M::rhs(t + h, yt, k6, p);
    for (int i = 0; i < NS; ++i)
      yt[i] = y[i] + h * (a71 * k1[i] + a73 * k3[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
    M::rhs(t + h, yt, k7, p);
    nfev += 6;
    double err = 0.0;
    for (int i = 0; i < NS; ++i) {
      const double sc = fmax(atol + rtol * fmax(fabs(y[i]), fabs(yt[i])), tiny);
      const double ei = h * (e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * k7[i]);
      err = fmax(err, fabs(ei) / sc);
    }
    nsteps++;
    if (!(err <= 1.0)) {  // reject (also catches NaN)
      if (!(err == err) || h < 1e-14 * span) { ok = 0; break; }
      h *= fmax(0.2, 0.9 * pow(err, -0.2));
      last_rejected = true;
      continue;
    }
    t = final_step ? t1 : t + h;
    for (int i = 0; i < NS; ++i) { y[i] = yt[i]; k1[i] = k7[i]; }
    double fac = (err < 1e-10) ? 5.0 : fmin(5.0, 0.9 * pow(err, -0.2));
    if (last_rejected) fac = fmin(fac, 1.0);
    last_rejected = false;
    h *= fac;
  }
  // currents (and any other side-effect outputs) at the end state
  M::rhs(t1, y, k1, p); nfev++;
}

KNP_HD void stat_max(int64_t* addr, int64_t v) {
#if defined(KNP_EMU) || !defined(__CUDA_ARCH__)
  if (v > *addr) *addr = v;
#else
  atomicMax((long long*)addr, (long long)v);
#endif
}
KNP_HD void stat_add(int64_t* addr, int64_t v) {
#if defined(KNP_EMU) || !defined(__CUDA_ARCH__)
  *addr += v;
#else
  atomicAdd((unsigned long long*)addr, (unsigned long long)v);
#endif
}

template <class M, int D>
struct OdeStepKernel {
  int64_t nc;
  const int32_t* rows;          // membrane row of each ODE point
  double* states; double* params;
  int nlinks; OdeLink links[MAX_LINKS];
  int set_v, v_col;
  double* phiM;
  int n_ion; int ich_cols[MAX_IONS]; double* Ich[MAX_IONS];
  const uint8_t* stim_mask; int nstim; int stim_cols[MAX_STIM]; double stim_vals[MAX_STIM];
  const int32_t* mem_ci; const int32_t* mem_fi; const int32_t* nbr; const int32_t* finfo;
  double t0, dt, rtol, atol;
  int64_t* stats;               // [0] max steps, [1] total rhs evaluations, [2] failures, [3] facets that took the stiff path

  KNP_HD void operator()(int64_t row) const {
    constexpr int NS = M::NS, NP = M::NP;
    const int64_t m = rows[row];
    double y[NS], p[NP];
    for (int i = 0; i < NS; ++i) y[i] = states[row * NS + i];
    for (int i = 0; i < NP; ++i) p[i] = params[row * NP + i];
    // PDE -> ODE (solver.py:1094-1101)
    if (set_v) y[v_col] = phiM[m];
    for (int l = 0; l < nlinks; ++l) {
      double v;
      if (links[l].kind == 0) {
        v = links[l].src[m];
      } else {
        const int64_t ci = mem_ci[m];
        const int fi = mem_fi[m];
        v = facet_mean_trace<D>(links[l].src, links[l].side, ci, fi, finfo[fi * nc + ci],
                                nbr[fi * nc + ci]);
      }
      for (int i = 0; i < NP; ++i)
        if (i == links[l].col) p[i] = v;
    }
    // stimulus overwrite (membrane.py:102-104)
    if (stim_mask && stim_mask[row])
      for (int s = 0; s < nstim; ++s)
        for (int i = 0; i < NP; ++i)
          if (i == stim_cols[s]) p[i] = stim_vals[s];
    int nsteps, nfev, ok;
    dopri5<M>(t0, t0 + dt, y, p, rtol, atol, nsteps, nfev, ok);
    for (int i = 0; i < NS; ++i) states[row * NS + i] = y[i];
    for (int i = 0; i < NP; ++i) params[row * NP + i] = p[i];
    // ODE -> PDE (solver.py:1108-1113)
    phiM[m] = y[v_col];
    for (int k = 0; k < n_ion; ++k)
      for (int i = 0; i < NP; ++i)
        if (i == ich_cols[k]) Ich[k][m] = p[i];
    if (stats) {
      stat_max(stats + 0, nsteps);
      stat_add(stats + 1, nfev);
      if (!ok) stat_add(stats + 2, 1);
      if (ok == 2) stat_add(stats + 3, 1);
    }
  }
};

}  // namespace knp

// knp_dg.h - DG-P1 element arithmetic of the KNP-EMI splitting scheme in closed form.
//
// Restates (does not translate) the UFL forms of the reference:
//   EMI  a, L, B   src/knpemidg/solver.py:289-346, 377-395
//   KNP  a, L      src/knpemidg/solver.py:550-629
//   post-step      src/knpemidg/solver.py:809-842, utils.py:87-124
// The reference integrates these with FFC-generated quadrature; on affine simplices
// with P1 bases every matrix entry and the EMI right-hand side are polynomial
// integrals, evaluated here exactly with
//   int_S prod lambda^alpha = |S| m! alpha! / (|alpha| + m)!      (S an m-simplex).
// Only the membrane right-hand side of KNP (rational in c) and the Nernst
// logarithm use a facet quadrature rule (the default FIAT rules, see oracle/quadrature.py).
//
// Block layout ("block-ELL", slot major): a scalar matrix is (ND+1) slot arrays of
// nc dense ND x ND blocks; slot 0 = diagonal block of the cell, slot 1+f = coupling to
// the neighbour across local facet f (the facet opposite local vertex f).  Block entry
// [i][j] multiplies dof j of the *neighbour* cell.  No row pointer, no column search:
// the owner of a row gathers its ND facets (deterministic, atomic free).
#pragma once
#include "knp_common.h"

namespace knp {

constexpr int MAX_IONS = 6;
constexpr int MAX_TAGS = 16;

struct Params {
  double F, R, T, psi, C_M, C_phi, dt, tau_emi, tau_knp, inv_Lp2;
  int N, ntags, splitting, mms;
  double z[MAX_IONS];
  double D[MAX_IONS][MAX_TAGS];
  double rho[MAX_TAGS];
  double Csub[MAX_IONS][MAX_TAGS];
};

// facet info word: bits 0-1 kind, 2-3 neighbour's local facet, 4.. perm (2 bits per
// local vertex: index of the same mesh vertex in the neighbour), bit 12: this cell is
// the ICS ('minus' of n_g) side of a membrane facet.
enum { FK_SIP = 0, FK_MEMBRANE = 1, FK_NONE = 2 };
KNP_HD int fi_kind(int w) { return w & 3; }
KNP_HD int fi_nfacet(int w) { return (w >> 2) & 3; }
KNP_HD int fi_perm(int w, int a) { return (w >> (4 + 2 * a)) & 3; }
KNP_HD int fi_ics(int w) { return (w >> 12) & 1; }

// ---- facet quadrature rules (barycentric points on the facet, weights sum to 1) ----
// D=2: 3-point Gauss-Legendre (degree 5 >= both estimated degrees 4 and 5).
// D=3: Radon 7-point degree 5 (membrane rhs), Dunavant 6-point degree 4 (Nernst).
template <int D> struct FacetRule5;
template <int D> struct FacetRule4;

template <> struct FacetRule5<2> {
  static constexpr int NQ = 3;
  KNP_HD static void point(int q, double* b, double& w) {
    const double s = 0.7745966692414834;  // sqrt(3/5)
    const double x = q == 0 ? 0.5 * (1.0 - s) : (q == 1 ? 0.5 : 0.5 * (1.0 + s));
    w = q == 1 ? 4.0 / 9.0 : 5.0 / 18.0;
    b[0] = 1.0 - x;
    b[1] = x;
  }
};
template <> struct FacetRule4<2> : FacetRule5<2> {};

template <> struct FacetRule5<3> {
  static constexpr int NQ = 7;
  KNP_HD static void point(int q, double* b, double& w) {
    const double s15 = 3.872983346207417;  // sqrt(15)
    if (q == 0) { b[0] = b[1] = b[2] = 1.0 / 3.0; w = 0.225; return; }
    const bool first = q <= 3;
    const double s = first ? (6.0 - s15) / 21.0 : (6.0 + s15) / 21.0;
    w = first ? (155.0 - s15) / 1200.0 : (155.0 + s15) / 1200.0;
    const int k = (q - 1) % 3;
    b[0] = b[1] = b[2] = s;
    b[k] = 1.0 - 2.0 * s;
  }
};
template <> struct FacetRule4<3> {
  static constexpr int NQ = 6;
  KNP_HD static void point(int q, double* b, double& w) {
    const bool first = q < 3;
    const double s = first ? 0.445948490915965 : 0.091576213509771;
    w = first ? 0.223381589678011 : 0.109951743655322;
    const int k = q % 3;
    b[0] = b[1] = b[2] = s;
    b[k] = 1.0 - 2.0 * s;
  }
};

// multiplicity factor alpha! of lambda_a lambda_b lambda_c
KNP_HD double mult3(int a, int b, int c) {
  if (a == b && b == c) return 6.0;
  if (a == b || b == c || a == c) return 2.0;
  return 1.0;
}

// ---------------------------------------------------------------------------
// pre-pass over cells: nodal conductivity kappa and the diffusive flux vector
//   kappa_m = F psi sum_k z_k^2 D_k c_k,m           (solver.py:306, ALL ions)
//   q       = sum_k F z_k D_k grad c_k               (solver.py:309-310)
// ---------------------------------------------------------------------------
template <int D>
struct EmiPrepassKernel {
  static constexpr int ND = D + 1;
  Params P;
  const double* c[MAX_IONS];
  const double* grad;
  const int32_t* region;
  double* kappa;  // [nc][ND]
  double* q;      // [nc][D]
  KNP_HD void operator()(int64_t cell) const {
    const int r = region[cell];
    double g[ND][D];
    for (int i = 0; i < ND; ++i)
      for (int x = 0; x < D; ++x) g[i][x] = grad[cell * (ND * D) + i * D + x];
    double kap[ND], qq[D];
    for (int m = 0; m < ND; ++m) kap[m] = 0.0;
    for (int x = 0; x < D; ++x) qq[x] = 0.0;
    for (int k = 0; k < P.N; ++k) {
      const double Dk = P.D[k][r], zk = P.z[k];
      const double wk = P.F * zk * zk * Dk * P.psi;
      const double wq = P.F * zk * Dk;
      for (int m = 0; m < ND; ++m) {
        const double cm = c[k][cell * ND + m];
        kap[m] += wk * cm;
        for (int x = 0; x < D; ++x) qq[x] += wq * cm * g[m][x];
      }
    }
    for (int m = 0; m < ND; ++m) kappa[cell * ND + m] = kap[m];
    for (int x = 0; x < D; ++x) q[cell * D + x] = qq[x];
  }
};

template <int D>
struct GradKernel {  // gphi = grad(phi) per cell (solver.py:583, 593)
  static constexpr int ND = D + 1;
  const double* phi;
  const double* grad;
  double* gphi;
  KNP_HD void operator()(int64_t cell) const {
    double out[D];
    for (int x = 0; x < D; ++x) out[x] = 0.0;
    for (int m = 0; m < ND; ++m) {
      const double pm = phi[cell * ND + m];
      for (int x = 0; x < D; ++x) out[x] += pm * grad[cell * (ND * D) + m * D + x];
    }
    for (int x = 0; x < D; ++x) gphi[cell * D + x] = out[x];
  }
};

// ---------------------------------------------------------------------------
// EMI assembly, one cell-row block per index.
// ---------------------------------------------------------------------------
template <int D>
struct EmiCellKernel {
  static constexpr int ND = D + 1;
  Params P;
  int64_t nc;
  const double* grad; const double* vol; const double* h;
  const int32_t* nbr; const int32_t* finfo; const int32_t* fmem;
  const double* kappa; const double* q;
  const double* phiM;            // [nm]
  const double* Ich[MAX_IONS];   // [nm] each (only read when !splitting)
  const double* load;            // extra load vector or nullptr
  double* A;                     // (ND+1) slots; slot 0 receives the diagonal blocks of B
  double* Adiag;                 // [nc][ND][ND] diagonal blocks of A
  double* rhs;                   // [n]

  KNP_HD void operator()(int64_t cell) const {
    constexpr double c_m2 = 1.0 / (D * (D + 1));              // facet mass
    constexpr double c_m3 = (D == 3) ? 1.0 / 60.0 : 1.0 / 24.0;   // facet cubic moment
    constexpr double c_k3 = (D == 3) ? 1.0 / 120.0 : 1.0 / 60.0;  // cell cubic moment
    const int64_t bs = ND * ND;
    double g[ND][D];
    for (int i = 0; i < ND; ++i)
      for (int x = 0; x < D; ++x) g[i][x] = grad[cell * (ND * D) + i * D + x];
    const double K = vol[cell], hK = h[cell];
    double kap[ND], qc[D];
    double kbar = 0.0;
    for (int m = 0; m < ND; ++m) { kap[m] = kappa[cell * ND + m]; kbar += kap[m]; }
    kbar /= ND;
    for (int x = 0; x < D; ++x) qc[x] = q[cell * D + x];

    double dg[ND][ND], bd[ND][ND], r[ND];
    // cell integrals: kappa grad u . grad v (solver.py:325), rhs -q.grad v (:309),
    // and the mass shift of the preconditioner form (:393)
    for (int i = 0; i < ND; ++i) {
      double qi = 0.0;
      for (int x = 0; x < D; ++x) qi += qc[x] * g[i][x];
      r[i] = -K * qi;
      for (int j = 0; j < ND; ++j) {
        double gg = 0.0;
        for (int x = 0; x < D; ++x) gg += g[i][x] * g[j][x];
        dg[i][j] = K * kbar * gg;
        double mk = 0.0;
        for (int m = 0; m < ND; ++m) mk += kap[m] * mult3(m, i, j);
        bd[i][j] = K * c_k3 * mk * P.inv_Lp2;
      }
    }

    for (int f = 0; f < ND; ++f) {
      const int w = finfo[f * nc + cell];
      const int kind = fi_kind(w);
      double O[ND][ND];
      for (int i = 0; i < ND; ++i)
        for (int j = 0; j < ND; ++j) O[i][j] = 0.0;
      if (kind != FK_NONE) {
        const int64_t c2 = nbr[f * nc + cell];
        double gn2 = 0.0;
        for (int x = 0; x < D; ++x) gn2 += g[f][x] * g[f][x];
        const double gnorm = sqrt(gn2);
        const double area = gnorm * D * K;
        double n[D];
        for (int x = 0; x < D; ++x) n[x] = -g[f][x] / gnorm;
        int perm[ND];
        for (int a = 0; a < ND; ++a) perm[a] = fi_perm(w, a);
        if (kind == FK_SIP) {
          double g2[ND][D], kap2[ND];
          for (int i = 0; i < ND; ++i)
            for (int x = 0; x < D; ++x) g2[i][x] = grad[c2 * (ND * D) + i * D + x];
          for (int m = 0; m < ND; ++m) kap2[m] = kappa[c2 * ND + m];
          const double beta = P.tau_emi / (0.5 * (hK + h[c2]));
          double gn_me[ND], gn_nb[ND];
          for (int j = 0; j < ND; ++j) {
            double a1 = 0.0, a2 = 0.0;
            for (int x = 0; x < D; ++x) { a1 += g[j][x] * n[x]; a2 += g2[j][x] * n[x]; }
            gn_me[j] = a1; gn_nb[j] = a2;
          }
          double S_me[ND], S_nb[ND], knb[ND];
          for (int a = 0; a < ND; ++a) knb[a] = (a == f) ? 0.0 : kap2[perm[a]];
          for (int i = 0; i < ND; ++i) {
            double s1 = 0.0, s2 = 0.0;
            if (i != f) {
              for (int a = 0; a < ND; ++a) {
                if (a == f) continue;
                const double m2 = (a == i) ? 2.0 : 1.0;
                s1 += kap[a] * m2; s2 += knb[a] * m2;
              }
            }
            S_me[i] = s1 * c_m2 * area; S_nb[i] = s2 * c_m2 * area;
          }
          for (int i = 0; i < ND; ++i) {
            for (int j = 0; j < ND; ++j) {
              double pen = 0.0;
              if (i != f && j != f) {
                for (int a = 0; a < ND; ++a) {
                  if (a == f) continue;
                  pen += 0.5 * (kap[a] + knb[a]) * mult3(a, i, j);
                }
                pen *= c_m3 * area * beta;
              }
              dg[i][j] += -0.5 * gn_me[j] * S_me[i] - 0.5 * gn_me[i] * S_me[j] + pen;
              if (j != f) O[i][perm[j]] += 0.5 * gn_me[i] * S_me[j] - pen;
            }
            for (int jp = 0; jp < ND; ++jp) O[i][jp] += -0.5 * gn_nb[jp] * S_nb[i];
          }
          // rhs: avg(q).n+ jump(v)  (solver.py:310)
          double fl = 0.0;
          for (int x = 0; x < D; ++x) fl += 0.5 * (qc[x] + q[c2 * D + x]) * n[x];
          fl *= area / D;
          for (int i = 0; i < ND; ++i)
            if (i != f) r[i] += fl;
        } else {  // membrane: C_phi jump(u) jump(v), robin data (solver.py:334-346)
          const double cm = P.C_phi * c_m2 * area;
          for (int i = 0; i < ND; ++i) {
            if (i == f) continue;
            for (int j = 0; j < ND; ++j) {
              if (j == f) continue;
              const double v = cm * ((i == j) ? 2.0 : 1.0);
              dg[i][j] += v;
              O[i][perm[j]] -= v;
            }
          }
          if (!P.mms) {
            const int64_t m = fmem[f * nc + cell];
            double gr = phiM[m];
            if (!P.splitting) {
              double It = 0.0;
              for (int k = 0; k < P.N; ++k) It += Ich[k][m];
              gr -= It / P.C_phi;
            }
            const double st = fi_ics(w) ? 1.0 : -1.0;
            const double v = P.C_phi * gr * st * area / D;
            for (int i = 0; i < ND; ++i)
              if (i != f) r[i] += v;
          }
        }
      }
      double* Of = A + (int64_t)(1 + f) * nc * bs + cell * bs;
      for (int i = 0; i < ND; ++i)
        for (int j = 0; j < ND; ++j) Of[i * ND + j] = O[i][j];
    }
    double* Ad = Adiag + cell * bs;
    double* Bd = A + cell * bs;
    for (int i = 0; i < ND; ++i) {
      for (int j = 0; j < ND; ++j) {
        Ad[i * ND + j] = dg[i][j];
        Bd[i * ND + j] = dg[i][j] + bd[i][j];
      }
      rhs[cell * ND + i] = r[i] + (load ? load[cell * ND + i] : 0.0);
    }
  }
};

// ---------------------------------------------------------------------------
// KNP assembly for one solved ion, one cell-row block per index.
// ---------------------------------------------------------------------------
template <int D>
struct KnpCellKernel {
  static constexpr int ND = D + 1;
  Params P;
  int64_t nc;
  int ion;
  const double* grad; const double* vol; const double* h;
  const int32_t* region; const int32_t* nbr; const int32_t* finfo; const int32_t* fmem;
  const double* gphi;            // [nc][D]
  const double* phi;             // [n]
  const double* c[MAX_IONS];     // c_prev_k of all ions (alpha, solver.py:603)
  const double* cn;              // c_prev_n of this ion
  const double* phiM; const double* Ich[MAX_IONS];
  const double* load;            // extra load vector or nullptr
  double* A;                     // (ND+1) slots
  double* rhs;

  KNP_HD void operator()(int64_t cell) const {
    constexpr double c_m2 = 1.0 / (D * (D + 1));
    constexpr double c_mass = 1.0 / ((D + 1) * (D + 2));
    const int64_t bs = ND * ND;
    const int reg = region[cell];
    const double Dme = P.D[ion][reg], z = P.z[ion];
    const double zpsi = z * P.psi;
    double g[ND][D];
    for (int i = 0; i < ND; ++i)
      for (int x = 0; x < D; ++x) g[i][x] = grad[cell * (ND * D) + i * D + x];
    const double K = vol[cell], hK = h[cell];
    double gp[D];
    for (int x = 0; x < D; ++x) gp[x] = gphi[cell * D + x];

    double dg[ND][ND], r[ND];
    double cnl[ND];
    for (int m = 0; m < ND; ++m) cnl[m] = cn[cell * ND + m];
    for (int i = 0; i < ND; ++i) {
      double dr = 0.0;
      for (int x = 0; x < D; ++x) dr += gp[x] * g[i][x];
      const double drift = zpsi * Dme * dr * K / (D + 1);   // solver.py:593
      double ri = 0.0;
      for (int j = 0; j < ND; ++j) {
        double gg = 0.0;
        for (int x = 0; x < D; ++x) gg += g[i][x] * g[j][x];
        const double mij = K * c_mass * ((i == j) ? 2.0 : 1.0);
        dg[i][j] = mij / P.dt + Dme * K * gg + drift;         // solver.py:586-587
        ri += mij * cnl[j];
      }
      r[i] = ri / P.dt;                                       // solver.py:597
    }

    for (int f = 0; f < ND; ++f) {
      const int w = finfo[f * nc + cell];
      const int kind = fi_kind(w);
      double O[ND][ND];
      for (int i = 0; i < ND; ++i)
        for (int j = 0; j < ND; ++j) O[i][j] = 0.0;
      if (kind != FK_NONE) {
        const int64_t c2 = nbr[f * nc + cell];
        double gn2 = 0.0;
        for (int x = 0; x < D; ++x) gn2 += g[f][x] * g[f][x];
        const double gnorm = sqrt(gn2);
        const double area = gnorm * D * K;
        double n[D];
        for (int x = 0; x < D; ++x) n[x] = -g[f][x] / gnorm;
        int perm[ND];
        for (int a = 0; a < ND; ++a) perm[a] = fi_perm(w, a);
        if (kind == FK_SIP) {
          const int nf = fi_nfacet(w);
          const double Dnb = P.D[ion][region[c2]];
          const double beta = P.tau_knp / (0.5 * (hK + h[c2]));
          double gn_me[ND], gn_nb[ND];
          double un_me = 0.0, un_nb = 0.0;
          for (int x = 0; x < D; ++x) { un_me += gp[x] * n[x]; un_nb -= gphi[c2 * D + x] * n[x]; }
          un_me = fmax(Dme * un_me, 0.0);                      // solver.py:583
          un_nb = fmax(Dnb * un_nb, 0.0);
          for (int j = 0; j < ND; ++j) {
            double a1 = 0.0, a2 = 0.0;
            for (int x = 0; x < D; ++x) {
              a1 += g[j][x] * n[x];
              a2 += grad[c2 * (ND * D) + j * D + x] * n[x];
            }
            gn_me[j] = a1; gn_nb[j] = a2;
          }
          const double af = area / D;       // int_F lambda_a
          const double pm = (beta * Dme - zpsi * un_me) * c_m2 * area;
          const double pn = (-beta * Dnb + zpsi * un_nb) * c_m2 * area;
          for (int i = 0; i < ND; ++i) {
            for (int j = 0; j < ND; ++j) {
              double v = 0.0;
              if (i != f) v += -0.5 * Dme * gn_me[j] * af;
              if (j != f) v += -0.5 * Dme * gn_me[i] * af;
              if (i != f && j != f) {
                const double m2 = (i == j) ? 2.0 : 1.0;
                v += pm * m2;
                O[i][perm[j]] += pn * m2;
              }
              dg[i][j] += v;
            }
            for (int jp = 0; jp < ND; ++jp) {
              double v = 0.0;
              if (i != f) v += -0.5 * Dnb * gn_nb[jp] * af;
              if (jp != nf) v += 0.5 * Dme * gn_me[i] * af;
              O[i][jp] += v;
            }
          }
        } else {
          // membrane right-hand side (solver.py:603-629); matrix gets nothing.
          const int64_t m = fmem[f * nc + cell];
          const double st = fi_ics(w) ? 1.0 : -1.0;
          double pme[ND], pnb[ND];
          for (int a = 0; a < ND; ++a) {
            pme[a] = phi[cell * ND + a];
            pnb[a] = (a == f) ? 0.0 : phi[c2 * ND + perm[a]];
          }
          double cme[MAX_IONS][ND];
          double wk[MAX_IONS];
          if (!P.mms) {
            for (int k = 0; k < P.N; ++k) {
              wk[k] = P.D[k][reg] * P.z[k] * P.z[k];
              for (int a = 0; a < ND; ++a) cme[k][a] = c[k][cell * ND + a];
            }
          }
          double Ik = 0.0, It = 0.0, pM = 0.0;
          if (!P.mms) {
            pM = phiM[m];
            Ik = Ich[ion][m];
            if (P.splitting)
              for (int k = 0; k < P.N; ++k) It += Ich[k][m];
          }
          const double Fz = P.F * z;
          for (int qd = 0; qd < FacetRule5<D>::NQ; ++qd) {
            double b[D], wq;
            FacetRule5<D>::point(qd, b, wq);
            // facet barycentric coordinate of local vertex a (a != f)
            double lam[ND];
            { int t = 0; for (int a = 0; a < ND; ++a) lam[a] = (a == f) ? 0.0 : b[t++]; }
            double dphi = 0.0;
            for (int a = 0; a < ND; ++a) dphi += lam[a] * (pme[a] - pnb[a]);
            dphi *= st;                                       // phi_i - phi_e
            double val;
            if (!P.mms) {
              double num = 0.0, tot = 0.0;
              for (int k = 0; k < P.N; ++k) {
                double ck = 0.0;
                for (int a = 0; a < ND; ++a) ck += lam[a] * cme[k][a];
                const double t = wk[k] * ck;
                tot += t;
                if (k == ion) num = t;
              }
              const double alpha = num / tot;                  // solver.py:603
              const double C = alpha * P.C_M / (Fz * P.dt);    // solver.py:606
              // C*g_robin (solver.py:616-622) minus the jump(phi) coupling (:628-629)
              val = C * pM - Ik / Fz + alpha * It / Fz - C * dphi;
            } else {
              val = -P.Csub[ion][reg] * dphi;
            }
            val *= st * wq * area;
            for (int a = 0; a < ND; ++a) r[a] += val * lam[a];
          }
        }
      }
      double* Of = A + (int64_t)(1 + f) * nc * bs + cell * bs;
      for (int i = 0; i < ND; ++i)
        for (int j = 0; j < ND; ++j) Of[i * ND + j] = O[i][j];
    }
    double* Ad = A + cell * bs;
    for (int i = 0; i < ND; ++i) {
      for (int j = 0; j < ND; ++j) Ad[i * ND + j] = dg[i][j];
      rhs[cell * ND + i] = r[i] + (load ? load[cell * ND + i] : 0.0);
    }
  }
};

// ---------------------------------------------------------------------------
// post-step (solver.py:809-842)
// ---------------------------------------------------------------------------
// eliminated ion: c_N = -(sum_k z_k c_k + rho)/z_N, nodewise (solver.py:831-838; the
// reference L2-projects this DG1 expression onto DG1, which is the identity).
template <int D>
struct EliminatedIonKernel {
  static constexpr int ND = D + 1;
  Params P;
  const int32_t* region;
  const double* c[MAX_IONS];
  double* celim;
  KNP_HD void operator()(int64_t dof) const {
    const int64_t cell = dof / ND;
    double s = P.rho[region[cell]];
    for (int k = 0; k < P.N - 1; ++k) s += P.z[k] * c[k][dof];
    celim[dof] = -s / P.z[P.N - 1];
  }
};

// facet mean of a one-sided P1 trace on membrane row m (utils.py:87-124):
// mean over the facet's D vertices.  side 0 = plus/ECS, 1 = minus/ICS.
template <int D>
KNP_HD double facet_mean_trace(const double* field, int side, int64_t ci, int fi, int w,
                               int64_t ce) {
  constexpr int ND = D + 1;
  double s = 0.0;
  for (int a = 0; a < ND; ++a) {
    if (a == fi) continue;
    s += side ? field[ci * ND + a] : field[ce * ND + fi_perm(w, a)];
  }
  return s / D;
}

template <int D>
struct MembranePostKernel {
  static constexpr int ND = D + 1;
  Params P;
  int64_t nc;
  const int32_t* mem_ci; const int32_t* mem_fi;   // ICS cell and its local facet
  const int32_t* nbr; const int32_t* finfo;
  const double* phi;
  const double* c[MAX_IONS];
  double* phiM;
  double* E[MAX_IONS];
  int do_phim, do_nernst;
  KNP_HD void operator()(int64_t m) const {
    const int64_t ci = mem_ci[m];
    const int fi = mem_fi[m];
    const int w = finfo[fi * nc + ci];
    const int64_t ce = nbr[fi * nc + ci];
    // phi_M = facet mean of phi_i - phi_e (solver.py:813-814)
    if (do_phim)
      phiM[m] = facet_mean_trace<D>(phi, 1, ci, fi, w, ce) - facet_mean_trace<D>(phi, 0, ci, fi, w, ce);
    if (!do_nernst) return;
    // E_k = RT/(F z_k) mean_F ln(c_e/c_i)  (solver.py:299, 823-828, 841-842)
    for (int k = 0; k < P.N; ++k) {
      double acc = 0.0;
      for (int qd = 0; qd < FacetRule4<D>::NQ; ++qd) {
        double b[D], wq;
        FacetRule4<D>::point(qd, b, wq);
        double vi = 0.0, ve = 0.0;
        int t = 0;
        for (int a = 0; a < ND; ++a) {
          if (a == fi) continue;
          vi += b[t] * c[k][ci * ND + a];
          ve += b[t] * c[k][ce * ND + fi_perm(w, a)];
          ++t;
        }
        acc += wq * log(ve / vi);
      }
      E[k][m] = P.R * P.T / (P.F * P.z[k]) * acc;
    }
  }
};

template <int D>
struct FacetTraceKernel {
  int64_t nc;
  const int32_t* mem_ci; const int32_t* mem_fi; const int32_t* nbr; const int32_t* finfo;
  const double* field; int side; double* out;
  KNP_HD void operator()(int64_t m) const {
    const int64_t ci = mem_ci[m];
    const int fi = mem_fi[m];
    out[m] = facet_mean_trace<D>(field, side, ci, fi, finfo[fi * nc + ci], nbr[fi * nc + ci]);
  }
};

// inverse of the ND x ND diagonal blocks (element block-Jacobi), Gauss-Jordan with
// partial pivoting, one block per index.
template <int ND>
struct BlockInverseKernel {
  const double* blocks; double* inv;
  KNP_HD void operator()(int64_t cell) const {
    double a[ND][ND], b[ND][ND];
    for (int i = 0; i < ND; ++i)
      for (int j = 0; j < ND; ++j) { a[i][j] = blocks[cell * ND * ND + i * ND + j]; b[i][j] = (i == j); }
    for (int col = 0; col < ND; ++col) {
      int piv = col;
      double best = fabs(a[col][col]);
      for (int rr = col + 1; rr < ND; ++rr)
        if (fabs(a[rr][col]) > best) { best = fabs(a[rr][col]); piv = rr; }
      if (piv != col)
        for (int j = 0; j < ND; ++j) {
          double t = a[col][j]; a[col][j] = a[piv][j]; a[piv][j] = t;
          t = b[col][j]; b[col][j] = b[piv][j]; b[piv][j] = t;
        }
      const double ip = 1.0 / a[col][col];
      for (int j = 0; j < ND; ++j) { a[col][j] *= ip; b[col][j] *= ip; }
      for (int rr = 0; rr < ND; ++rr) {
        if (rr == col) continue;
        const double fct = a[rr][col];
        for (int j = 0; j < ND; ++j) { a[rr][j] -= fct * a[col][j]; b[rr][j] -= fct * b[col][j]; }
      }
    }
    for (int i = 0; i < ND; ++i)
      for (int j = 0; j < ND; ++j) inv[cell * ND * ND + i * ND + j] = b[i][j];
  }
};

}  // namespace knp

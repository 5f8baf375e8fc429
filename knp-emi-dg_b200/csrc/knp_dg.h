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
    const double t = 1.0 - 2.0 * s;
    b[0] = (k == 0) ? t : s; b[1] = (k == 1) ? t : s; b[2] = (k == 2) ? t : s;
  }
};
template <> struct FacetRule4<3> {
  static constexpr int NQ = 6;
  KNP_HD static void point(int q, double* b, double& w) {
    const bool first = q < 3;
    const double s = first ? 0.445948490915965 : 0.091576213509771;
    w = first ? 0.223381589678011 : 0.109951743655322;
    const int k = q % 3;
    const double t = 1.0 - 2.0 * s;
    b[0] = (k == 0) ? t : s; b[1] = (k == 1) ? t : s; b[2] = (k == 2) ? t : s;
  }
};

// the ND nodal values of one cell of a DG field (dof = ND*cell + i); 16-byte vector loads
// for ND = 4 (cudaMalloc'ed fields are 256-byte aligned, a cell is 32 bytes)
template <int ND>
KNP_HD void load_cell(const double* field, int64_t cell, double (&out)[ND]) {
#if defined(__CUDA_ARCH__)
  if (ND == 4) {
    const double2* p = reinterpret_cast<const double2*>(field + cell * 4);
    const double2 a = p[0], b = p[1];
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[ND - 1] = b.y;
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < ND; ++i) out[i] = field[cell * ND + i];
}

// multiplicity factor alpha! of lambda_a lambda_b lambda_c
KNP_HD double mult3(int a, int b, int c) {
  if (a == b && b == c) return 6.0;
  if (a == b || b == c || a == c) return 2.0;
  return 1.0;
}

// ---------------------------------------------------------------------------
// Static facet geometry, computed once per mesh (knp_mesh_set) and read by both assembly kernels
// at every time step instead of being re-derived from the two cells' gradients (a square root, two
// divisions and 3 (D+1) gathered loads per cell side):
//   fgeo[(k (D+1) + F) nc + cell],  k = 0           |F|
//                                   k = 1           1 / avg(h) = 2 / (h_K + h_K')   (tag-0 facets; else 0)
//                                   k = 2 .. 1+D    outward unit normal of `cell` on its facet F
//                                   k = 2+D ..      grad(lambda^K'_{perm(j)}) . n, j = 0..D: the normal derivatives
//                                                   of the NEIGHBOUR's basis functions in my vertex order
// ---------------------------------------------------------------------------
template <int D> struct FGeo {
  static constexpr int ND = D + 1, N = 2 + D + ND;
  KNP_HD static int64_t at(int k, int F, int64_t nc, int64_t cell) { return ((int64_t)k * ND + F) * nc + cell; }
};

template <int D>
struct FacetGeomKernel {
  static constexpr int ND = D + 1;
  int64_t nc;
  const double* grad; const double* vol; const double* h;
  const int32_t* nbr; const int32_t* finfo;
  double* fgeo;
  KNP_HD void operator()(int64_t cell) const {
    double g[ND][D];
    for (int i = 0; i < ND; ++i)
      for (int x = 0; x < D; ++x) g[i][x] = grad[(i * D + x) * nc + cell];
    const double K = vol[cell], hK = h[cell];
    for (int F = 0; F < ND; ++F) {
      const int w = finfo[F * nc + cell];
      double gn2 = 0.0;
      for (int x = 0; x < D; ++x) gn2 += g[F][x] * g[F][x];
      const double gnorm = sqrt(gn2);
      const double inv_gnorm = 1.0 / gnorm;
      double n[D];
      for (int x = 0; x < D; ++x) n[x] = -g[F][x] * inv_gnorm;
      fgeo[FGeo<D>::at(0, F, nc, cell)] = gnorm * D * K;
      for (int x = 0; x < D; ++x) fgeo[FGeo<D>::at(2 + x, F, nc, cell)] = n[x];
      double inv_havg = 0.0, gn_nb[ND];
      for (int j = 0; j < ND; ++j) gn_nb[j] = 0.0;
      const int64_t c2 = nbr[F * nc + cell];
      if (fi_kind(w) == FK_SIP && c2 >= 0) {
        inv_havg = 1.0 / (0.5 * (hK + h[c2]));
        for (int j = 0; j < ND; ++j) {
          const int64_t pj = fi_perm(w, j);
          double a2 = 0.0;
          for (int x = 0; x < D; ++x) a2 += grad[(pj * D + x) * nc + c2] * n[x];
          gn_nb[j] = a2;
        }
      }
      fgeo[FGeo<D>::at(1, F, nc, cell)] = inv_havg;
      for (int j = 0; j < ND; ++j) fgeo[FGeo<D>::at(2 + D + j, F, nc, cell)] = gn_nb[j];
    }
  }
};

// ---------------------------------------------------------------------------
// pre-pass over cells: nodal conductivity kappa and the diffusive flux vector
//   kappa_m = F psi sum_k z_k^2 D_k c_k,m           (solver.py:306, ALL ions)
//   q       = sum_k F z_k D_k grad c_k               (solver.py:309-310)
// ---------------------------------------------------------------------------
template <int D>
struct EmiPrepassKernel {
  static constexpr int ND = D + 1;
  Params P;
  int64_t nc;
  const double* c[MAX_IONS];
  const double* grad;
  const int32_t* region;
  double* kappa;  // [ND][nc]  (component major: consecutive cells are consecutive addresses)
  double* q;      // [D][nc]
  KNP_HD void operator()(int64_t cell) const {
    const int r = region[cell];
    double g[ND][D];
    for (int i = 0; i < ND; ++i)
      for (int x = 0; x < D; ++x) g[i][x] = grad[(i * D + x) * nc + cell];
    double kap[ND], qq[D];
    for (int m = 0; m < ND; ++m) kap[m] = 0.0;
    for (int x = 0; x < D; ++x) qq[x] = 0.0;
    for (int k = 0; k < P.N; ++k) {
      const double Dk = P.D[k][r], zk = P.z[k];
      const double wk = P.F * zk * zk * Dk * P.psi;
      const double wq = P.F * zk * Dk;
      double cl[ND];
      load_cell<ND>(c[k], cell, cl);
      for (int m = 0; m < ND; ++m) {
        const double cm = cl[m];
        kap[m] += wk * cm;
        for (int x = 0; x < D; ++x) qq[x] += wq * cm * g[m][x];
      }
    }
    for (int m = 0; m < ND; ++m) kappa[m * nc + cell] = kap[m];
    for (int x = 0; x < D; ++x) q[x * nc + cell] = qq[x];
  }
};

template <int D>
struct GradKernel {  // gphi = grad(phi) per cell (solver.py:583, 593)
  static constexpr int ND = D + 1;
  int64_t nc;
  const double* phi;
  const double* grad;
  double* gphi;
  KNP_HD void operator()(int64_t cell) const {
    double out[D];
    for (int x = 0; x < D; ++x) out[x] = 0.0;
    double pl[ND];
    load_cell<ND>(phi, cell, pl);
    for (int m = 0; m < ND; ++m) {
      const double pm = pl[m];
      for (int x = 0; x < D; ++x) out[x] += pm * grad[(m * D + x) * nc + cell];
    }
    for (int x = 0; x < D; ++x) gphi[x * nc + cell] = out[x];
  }
};

// ---------------------------------------------------------------------------
// EMI assembly.  The arithmetic of one cell-row block is split into
//   emi_cell_row : cell integrals for test function i
//   emi_facet    : everything one facet contributes (its off-diagonal block, its
//                  share of the diagonal block and of the right-hand side)
// and is used by two drivers: EmiCellKernel (one index per cell: host emulation and
// small meshes) and emi_assemble_kernel (CUDA: one thread per (cell, facet), blocks staged
// through shared memory so that every global store is a full coalesced line).
// ---------------------------------------------------------------------------
template <int D>
struct EmiArgs {
  static constexpr int ND = D + 1;
  Params P;
  int64_t nc;                    // local cells (owned + ghost): stride of every per-cell array
  int64_t nw;                    // cells whose rows are assembled (the owned ones, numbered first)
  const double* grad; const double* vol; const double* h;
  const double* fgeo;            // static facet geometry (FGeo)
  const int32_t* nbr; const int32_t* finfo; const int32_t* fmem;
  const double* kappa; const double* q;
  const double* phiM;            // [nm]
  const double* Ich[MAX_IONS];   // [nm] each (only read when !splitting)
  const double* load;            // extra load vector or nullptr
  double* A;                     // (ND+1) slots; slot 0 receives the diagonal blocks of B
  double* Adiag;                 // [nc][ND][ND] diagonal blocks of A
  double* rhs;                   // [n]
};

// cell integrals for test function i: kappa grad u . grad v (solver.py:325), rhs
// -q.grad v (:309) and the mass shift of the preconditioner form (:393)
template <int D, int I>
KNP_HD void emi_cell_row(const EmiArgs<D>& a, const double (&g)[D + 1][D], double K,
                         const double (&kap)[D + 1], double kbar, const double (&qc)[D],
                         double* dgrow, double* bdrow, double& ri) {
  constexpr int ND = D + 1;
  constexpr double c_k3 = (D == 3) ? 1.0 / 120.0 : 1.0 / 60.0;  // cell cubic moment
  constexpr int i = I;
  double qi = 0.0;
  #pragma unroll
  for (int x = 0; x < D; ++x) qi += qc[x] * g[i][x];
  ri = -K * qi;
  #pragma unroll
  for (int j = 0; j < ND; ++j) {
    double gg = 0.0;
    #pragma unroll
    for (int x = 0; x < D; ++x) gg += g[i][x] * g[j][x];
    dgrow[j] = K * kbar * gg;
    double mk = 0.0;
    #pragma unroll
    for (int m = 0; m < ND; ++m) mk += kap[m] * mult3(m, i, j);
    bdrow[j] = K * c_k3 * mk * a.P.inv_Lp2;
  }
}

// Where a facet routine puts its block entries.  RegBlocks: plain arrays (one-index-per-cell
// drivers, host emulation).  SmemBlocks: straight into the shared-memory staging rows of the
// CUDA drivers - the entries never live in registers, which is what keeps those kernels'
// register footprint down; the column map is applied in the store address.
template <int ND>
struct RegBlocks {
  double (&O)[ND][ND]; double (&dg)[ND][ND];
  KNP_HD void zeroO() const {
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int j = 0; j < ND; ++j) O[i][j] = 0.0;
  }
  KNP_HD void setO(int i, int j, double v) const { O[i][j] = v; }
  KNP_HD void addD(int i, int j, double v) const { dg[i][j] += v; }
};
// The facet routines touch every entry they write exactly once (all ND x ND entries on a tag-0
// facet, the entries off row / column F on a membrane facet, none otherwise), so the diagonal
// contribution is STORED, not accumulated, and the staging rows need zeroing only for facets that
// are not tag-0 facets (the caller does that): no read-modify-write, no blanket zero fill.
template <int ND>
struct SmemBlocks {   // O, D: rows of ND*ND (+pad) doubles; w: facet info word
  double* O; double* D; int w;
  KNP_HD void zeroO() const {}
  KNP_HD void setO(int i, int j, double v) const { O[i * ND + fi_perm(w, j)] = v; }
  KNP_HD void addD(int i, int j, double v) const { D[i * ND + j] = v; }
};

// facet F of `cell`.  Outputs are in the cell's OWN vertex numbering: O[i][j] couples my
// dof i with the neighbour dof that sits at my vertex j (neighbour local index
// fi_perm(w, j)); column F, which has no counterpart on my side, holds the coupling with
// the neighbour's vertex opposite the facet (fi_perm(w, F) = its local facet index).  The
// writer applies that column map in the store address, so no register array is ever
// indexed dynamically.  dg += share of the diagonal block, r += share of the rhs.
// Returns the facet info word (column map).
template <int D, int F, class Blocks>
KNP_HD int emi_facet(const EmiArgs<D>& a, int64_t cell, const double (&g)[D + 1][D],
                     const double (&kap)[D + 1], const double (&qc)[D], const Blocks& B, double (&r)[D + 1]) {
  constexpr int ND = D + 1;
  constexpr double c_m2 = 1.0 / (D * (D + 1));                  // facet mass
  constexpr double c_m3 = (D == 3) ? 1.0 / 60.0 : 1.0 / 24.0;   // facet cubic moment
  const int64_t nc = a.nc;
  const int w = a.finfo[F * nc + cell];
  const int kind = fi_kind(w);
  B.zeroO();
  if (kind == FK_NONE) return w;
  const int64_t c2 = a.nbr[F * nc + cell];
  const double area = a.fgeo[FGeo<D>::at(0, F, nc, cell)];
  double n[D];
#pragma unroll
  for (int x = 0; x < D; ++x) n[x] = a.fgeo[FGeo<D>::at(2 + x, F, nc, cell)];
  if (kind == FK_SIP) {
    // neighbour data gathered directly in my vertex order
    double gn_me[ND], gn_nb[ND], knb[ND];
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int64_t pj = fi_perm(w, j);
      double a1 = 0.0;
#pragma unroll
      for (int x = 0; x < D; ++x) a1 += g[j][x] * n[x];
      gn_me[j] = a1; gn_nb[j] = a.fgeo[FGeo<D>::at(2 + D + j, F, nc, cell)];
      knb[j] = (j == F) ? 0.0 : a.kappa[pj * nc + c2];
    }
    const double beta = a.P.tau_emi * a.fgeo[FGeo<D>::at(1, F, nc, cell)];
    double S_me[ND], S_nb[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      double s1 = 0.0, s2 = 0.0;
      if (i != F) {
#pragma unroll
        for (int v = 0; v < ND; ++v) {
          if (v == F) continue;
          const double m2 = (v == i) ? 2.0 : 1.0;
          s1 += kap[v] * m2; s2 += knb[v] * m2;
        }
      }
      S_me[i] = s1 * c_m2 * area; S_nb[i] = s2 * c_m2 * area;
    }
    const double pscale = c_m3 * area * beta;
#pragma unroll
    for (int i = 0; i < ND; ++i) {
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        double pen = 0.0;
        if (i != F && j != F) {
#pragma unroll
          for (int v = 0; v < ND; ++v) {
            if (v == F) continue;
            pen += 0.5 * (kap[v] + knb[v]) * mult3(v, i, j);
          }
          pen *= pscale;
        }
        B.addD(i, j, -0.5 * gn_me[j] * S_me[i] - 0.5 * gn_me[i] * S_me[j] + pen);
        B.setO(i, j, -0.5 * gn_nb[j] * S_nb[i] + ((j != F) ? 0.5 * gn_me[i] * S_me[j] - pen : 0.0));
      }
    }
    // rhs: avg(q).n+ jump(v)  (solver.py:310)
    double fl = 0.0;
#pragma unroll
    for (int x = 0; x < D; ++x) fl += 0.5 * (qc[x] + a.q[x * nc + c2]) * n[x];
    fl *= area / D;
#pragma unroll
    for (int i = 0; i < ND; ++i)
      if (i != F) r[i] += fl;
  } else {  // membrane: C_phi jump(u) jump(v), robin data (solver.py:334-346)
    const double cm = a.P.C_phi * c_m2 * area;
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      if (i == F) continue;
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        if (j == F) continue;
        const double v = cm * ((i == j) ? 2.0 : 1.0);
        B.addD(i, j, v);
        B.setO(i, j, -v);
      }
    }
    if (!a.P.mms) {
      const int64_t m = a.fmem[F * nc + cell];
      double gr = a.phiM[m];
      if (!a.P.splitting) {
        double It = 0.0;
        for (int k = 0; k < a.P.N; ++k) It += a.Ich[k][m];
        gr -= It / a.P.C_phi;
      }
      const double st = fi_ics(w) ? 1.0 : -1.0;
      const double v = a.P.C_phi * gr * st * area / D;
#pragma unroll
      for (int i = 0; i < ND; ++i)
        if (i != F) r[i] += v;
    }
  }
  return w;
}

template <int D>
struct EmiCellKernel {
  static constexpr int ND = D + 1;
  EmiArgs<D> a;
  KNP_HD void operator()(int64_t cell) const {
    const int64_t bs = ND * ND;
    double g[ND][D];
    #pragma unroll
    for (int i = 0; i < ND; ++i)
      #pragma unroll
      for (int x = 0; x < D; ++x) g[i][x] = a.grad[(i * D + x) * a.nc + cell];
    const double K = a.vol[cell];
    double kap[ND], qc[D];
    double kbar = 0.0;
    #pragma unroll
    for (int m = 0; m < ND; ++m) { kap[m] = a.kappa[m * a.nc + cell]; kbar += kap[m]; }
    kbar /= ND;
    #pragma unroll
    for (int x = 0; x < D; ++x) qc[x] = a.q[x * a.nc + cell];
    double dg[ND][ND], bd[ND][ND], r[ND];
    emi_cell_row<D, 0>(a, g, K, kap, kbar, qc, dg[0], bd[0], r[0]);
    emi_cell_row<D, 1>(a, g, K, kap, kbar, qc, dg[1], bd[1], r[1]);
    emi_cell_row<D, 2>(a, g, K, kap, kbar, qc, dg[2], bd[2], r[2]);
    if constexpr (D == 3) emi_cell_row<D, D>(a, g, K, kap, kbar, qc, dg[D], bd[D], r[D]);
    #pragma unroll
    for (int f = 0; f < ND; ++f) {
      double O[ND][ND];
      int w;
      if (f == 0) w = emi_facet<D, 0>(a, cell, g, kap, qc, RegBlocks<ND>{O, dg}, r);
      else if (f == 1) w = emi_facet<D, 1>(a, cell, g, kap, qc, RegBlocks<ND>{O, dg}, r);
      else if (f == 2) w = emi_facet<D, 2>(a, cell, g, kap, qc, RegBlocks<ND>{O, dg}, r);
      else w = emi_facet<D, D>(a, cell, g, kap, qc, RegBlocks<ND>{O, dg}, r);
      double* Of = a.A + (int64_t)(1 + f) * a.nc * bs + cell * bs;
      #pragma unroll
      for (int i = 0; i < ND; ++i)
        #pragma unroll
        for (int j = 0; j < ND; ++j) Of[i * ND + fi_perm(w, j)] = O[i][j];
    }
    double* Ad = a.Adiag + cell * bs;
    double* Bd = a.A + cell * bs;
    #pragma unroll
    for (int i = 0; i < ND; ++i) {
      #pragma unroll
      for (int j = 0; j < ND; ++j) {
        Ad[i * ND + j] = dg[i][j];
        Bd[i * ND + j] = dg[i][j] + bd[i][j];
      }
      a.rhs[cell * ND + i] = r[i] + (a.load ? a.load[cell * ND + i] : 0.0);
    }
  }
};

#ifndef KNP_EMU
constexpr int ASM_CPB = 32;  // cells per thread block

// One thread per (cell, facet), warp w of a block = facet w of 32 consecutive cells (the
// facet index is warp uniform, so the templated facet code runs without divergence):
// thread (f, cl) computes facet f's blocks and row f of the cell integrals; the block's
// results are staged in shared memory (pitch 17/10 doubles per ND x ND block: conflict
// free) and written out slot by slot as contiguous runs of ASM_CPB * ND * ND doubles.
template <int D, int MINB>
__global__ void __launch_bounds__(ASM_CPB*(D + 1), MINB) emi_assemble_kernel(const EmiArgs<D> a) {
  constexpr int ND = D + 1, BS = ND * ND, PITCH = BS + 1, NT = ASM_CPB * ND;
  __shared__ double sO[ND][ASM_CPB][PITCH];   // off-diagonal blocks per facet
  __shared__ double sD[ND][ASM_CPB][PITCH];   // diagonal-block contributions per facet thread
  __shared__ double sB[ASM_CPB][PITCH];       // mass shift of B
  __shared__ double sR[ND][ASM_CPB][ND];      // rhs contributions
  const int t = threadIdx.x;
  const int f = t / ASM_CPB, cl = t - f * ASM_CPB;
  const int64_t cell0 = (int64_t)blockIdx.x * ASM_CPB;
  const int64_t cell = cell0 + cl;
  const int64_t nc = a.nc;
  if (cell < a.nw) {
    double g[ND][D];
    #pragma unroll
    for (int i = 0; i < ND; ++i)
      #pragma unroll
      for (int x = 0; x < D; ++x) g[i][x] = a.grad[(i * D + x) * a.nc + cell];
    const double K = a.vol[cell];
    double kap[ND], qc[D];
    double kbar = 0.0;
    #pragma unroll
    for (int m = 0; m < ND; ++m) { kap[m] = a.kappa[m * a.nc + cell]; kbar += kap[m]; }
    kbar /= ND;
    #pragma unroll
    for (int x = 0; x < D; ++x) qc[x] = a.q[x * a.nc + cell];
    double r[ND];
    #pragma unroll
    for (int i = 0; i < ND; ++i) r[i] = 0.0;
    // facet f and row f of the cell integrals (f is warp uniform); the block entries go
    // straight into this thread's staging rows
    double* myO = sO[f][cl];
    double* myD = sD[f][cl];
    const int wf = a.finfo[f * nc + cell];
    if (fi_kind(wf) != FK_SIP) {
      #pragma unroll
      for (int k = 0; k < BS; ++k) { myO[k] = 0.0; myD[k] = 0.0; }
    }
    const SmemBlocks<ND> blk{myO, myD, wf};
    double bdrow[ND], row[ND], ri;
    switch (f) {
      case 0: emi_facet<D, 0>(a, cell, g, kap, qc, blk, r);
              emi_cell_row<D, 0>(a, g, K, kap, kbar, qc, row, bdrow, ri); break;
      case 1: emi_facet<D, 1>(a, cell, g, kap, qc, blk, r);
              emi_cell_row<D, 1>(a, g, K, kap, kbar, qc, row, bdrow, ri); break;
      case 2: emi_facet<D, 2>(a, cell, g, kap, qc, blk, r);
              emi_cell_row<D, 2>(a, g, K, kap, kbar, qc, row, bdrow, ri); break;
      default: emi_facet<D, D>(a, cell, g, kap, qc, blk, r);
              emi_cell_row<D, D>(a, g, K, kap, kbar, qc, row, bdrow, ri); break;
    }
    r[f] += ri;
    #pragma unroll
    for (int j = 0; j < ND; ++j) { myD[f * ND + j] += row[j]; sB[cl][f * ND + j] = bdrow[j]; }
    #pragma unroll
    for (int i = 0; i < ND; ++i) sR[f][cl][i] = r[i];
  }
  __syncthreads();
  const int64_t ncell_blk = (a.nw - cell0 < ASM_CPB) ? (a.nw - cell0) : ASM_CPB;
  if (ncell_blk == ASM_CPB) {
    // full block (all but the last one): every trip count and every shared-memory address is a
    // compile-time function of the thread index - element e = t + k NT of a slot belongs to cell
    // e / BS = t / BS + k NT / BS and entry e % BS = t % BS
    constexpr int STEP = NT / BS;                       // cells advanced per pass (NT is a multiple of BS for ND = 4)
    static_assert(ND != 4 || NT % BS == 0, "store path assumes NT % BS == 0 in 3D");
    if constexpr (NT % BS == 0) {
      const int c0 = t / BS, kk = t % BS;
      #pragma unroll
      for (int s = 0; s < ND; ++s) {
        double* dst = a.A + (int64_t)(1 + s) * nc * BS + cell0 * BS + t;
        #pragma unroll
        for (int k = 0; k < ASM_CPB * BS / NT; ++k) dst[k * NT] = sO[s][c0 + k * STEP][kk];
      }
      double* dA = a.Adiag + cell0 * BS + t;
      double* dB = a.A + cell0 * BS + t;
      #pragma unroll
      for (int k = 0; k < ASM_CPB * BS / NT; ++k) {
        const int c = c0 + k * STEP;
        double v = 0.0;
        #pragma unroll
        for (int s = 0; s < ND; ++s) v += sD[s][c][kk];
        dA[k * NT] = v;
        dB[k * NT] = v + sB[c][kk];
      }
      {   // rhs: ASM_CPB * ND = NT entries, one per thread
        const int c = t / ND, i = t % ND;
        double v = 0.0;
        #pragma unroll
        for (int s = 0; s < ND; ++s) v += sR[s][c][i];
        a.rhs[cell0 * ND + t] = v + (a.load ? a.load[cell0 * ND + t] : 0.0);
      }
      return;
    }
  }
  const int nval = (int)ncell_blk * BS;
  // off-diagonal slots
  #pragma unroll
  for (int s = 0; s < ND; ++s) {
    double* dst = a.A + (int64_t)(1 + s) * nc * BS + cell0 * BS;
    for (int e = t; e < nval; e += NT) dst[e] = sO[s][e / BS][e % BS];
  }
  // diagonal blocks of A and of B
  {
    double* dA = a.Adiag + cell0 * BS;
    double* dB = a.A + cell0 * BS;
    for (int e = t; e < nval; e += NT) {
      const int c = e / BS, k = e % BS;
      double v = 0.0;
      #pragma unroll
      for (int s = 0; s < ND; ++s) v += sD[s][c][k];
      dA[e] = v;
      dB[e] = v + sB[c][k];
    }
  }
  // rhs
  {
    const int nrow = (int)ncell_blk * ND;
    double* dr = a.rhs + cell0 * ND;
    const double* ld = a.load ? a.load + cell0 * ND : nullptr;
    for (int e = t; e < nrow; e += NT) {
      const int c = e / ND, i = e % ND;
      double v = 0.0;
      #pragma unroll
      for (int s = 0; s < ND; ++s) v += sR[s][c][i];
      dr[e] = v + (ld ? ld[e] : 0.0);
    }
  }
}
#endif

// ---------------------------------------------------------------------------
// KNP assembly (same structure as EMI).  The facet work is split into an ion-independent
// part (geometry, normal fluxes of the basis gradients, upwind direction) and a cheap
// per-ion part, so the CUDA kernel assembles ALL solved ions in one launch while loading
// the cell and neighbour data once.
// ---------------------------------------------------------------------------
template <int D>
struct KnpArgs {
  static constexpr int ND = D + 1;
  Params P;
  int64_t nc;                    // local cells (stride)
  int64_t nw;                    // owned cells (rows assembled)
  int nion;                      // number of solved ions (N-1)
  const double* grad; const double* vol; const double* h;
  const double* fgeo;            // static facet geometry (FGeo)
  const int32_t* region; const int32_t* nbr; const int32_t* finfo;
  const double* gphi;            // [D][nc]
  const double* cn[MAX_IONS];    // c_prev_n per solved ion
  const double* load[MAX_IONS];  // extra load vector or nullptr
  double* A[MAX_IONS];           // (ND+1) slots per solved ion
  double* rhs[MAX_IONS];
};

template <int D>
struct KnpFacetGeom {
  int w, kind, regnb;
  double area, beta, dphin_me, dphin_nb;   // grad(phi).n on both sides (own outward normals)
  double gn_me[D + 1], gn_nb[D + 1];
};

// cell integrals for test function I of ion `ion` (solver.py:586-587, 593, 597)
template <int D, int I>
KNP_HD void knp_cell_row(const Params& P, int ion, const double (&g)[D + 1][D], double K, double Dme,
                         const double (&gp)[D], const double (&cnl)[D + 1], double* dgrow, double& ri) {
  constexpr int ND = D + 1;
  constexpr double c_mass = 1.0 / ((D + 1) * (D + 2));
  constexpr int i = I;
  const double zpsi = P.z[ion] * P.psi;
  double dr = 0.0;
#pragma unroll
  for (int x = 0; x < D; ++x) dr += gp[x] * g[i][x];
  const double drift = zpsi * Dme * dr * K / (D + 1);
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    double gg = 0.0;
#pragma unroll
    for (int x = 0; x < D; ++x) gg += g[i][x] * g[j][x];
    const double mij = K * c_mass * ((i == j) ? 2.0 : 1.0);
    dgrow[j] = mij / P.dt + Dme * K * gg + drift;
    acc += mij * cnl[j];
  }
  ri = acc / P.dt;
}

// ion-independent part of facet F (neighbour data gathered in my vertex order, see emi_facet)
template <int D, int F>
KNP_HD void knp_facet_geom(const KnpArgs<D>& a, int64_t cell, const double (&g)[D + 1][D],
                           const double (&gp)[D], KnpFacetGeom<D>& G) {
  constexpr int ND = D + 1;
  const int64_t nc = a.nc;
  G.w = a.finfo[F * nc + cell];
  G.kind = fi_kind(G.w);
  if (G.kind != FK_SIP) return;   // membrane / untagged facets: nothing in the KNP matrix
  const int64_t c2 = a.nbr[F * nc + cell];
  G.area = a.fgeo[FGeo<D>::at(0, F, nc, cell)];
  double n[D];
#pragma unroll
  for (int x = 0; x < D; ++x) n[x] = a.fgeo[FGeo<D>::at(2 + x, F, nc, cell)];
  G.regnb = a.region[c2];
  G.beta = a.P.tau_knp * a.fgeo[FGeo<D>::at(1, F, nc, cell)];
  double d1 = 0.0, d2 = 0.0;
#pragma unroll
  for (int x = 0; x < D; ++x) { d1 += gp[x] * n[x]; d2 -= a.gphi[x * nc + c2] * n[x]; }
  G.dphin_me = d1; G.dphin_nb = d2;
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    double a1 = 0.0;
#pragma unroll
    for (int x = 0; x < D; ++x) a1 += g[j][x] * n[x];
    G.gn_me[j] = a1; G.gn_nb[j] = a.fgeo[FGeo<D>::at(2 + D + j, F, nc, cell)];
  }
}

// per-ion part of facet F: O (own vertex order, column F = neighbour's opposite vertex),
// dg += share of the diagonal block (solver.py:583-594)
template <int D, int F, class Blocks>
KNP_HD void knp_facet_ion(const Params& P, int ion, const KnpFacetGeom<D>& G, double Dme, const Blocks& B) {
  constexpr int ND = D + 1;
  constexpr double c_m2 = 1.0 / (D * (D + 1));
  B.zeroO();
  if (G.kind != FK_SIP) return;
  const double zpsi = P.z[ion] * P.psi;
  const double Dnb = P.D[ion][G.regnb];
  const double un_me = fmax(Dme * G.dphin_me, 0.0);     // upwinding, solver.py:583
  const double un_nb = fmax(Dnb * G.dphin_nb, 0.0);
  const double af = G.area / D;                          // int_F lambda_a
  const double pm = (G.beta * Dme - zpsi * un_me) * c_m2 * G.area;
  const double pn = (-G.beta * Dnb + zpsi * un_nb) * c_m2 * G.area;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      double v = 0.0, o = 0.0;
      if (i != F) { v += -0.5 * Dme * G.gn_me[j] * af; o += -0.5 * Dnb * G.gn_nb[j] * af; }
      if (j != F) { v += -0.5 * Dme * G.gn_me[i] * af; o += 0.5 * Dme * G.gn_me[i] * af; }
      if (i != F && j != F) {
        const double m2 = (i == j) ? 2.0 : 1.0;
        v += pm * m2;
        o += pn * m2;
      }
      B.addD(i, j, v);
      B.setO(i, j, o);
    }
  }
}

// membrane right-hand side of KNP (solver.py:603-629) for ALL solved ions, one index per
// cell that owns at least one membrane facet (facets visited in local order: deterministic,
// atomic free).  Runs after the cell/facet kernel and adds to its right-hand sides.
template <int D>
struct KnpMembraneRhsKernel {
  static constexpr int ND = D + 1;
  Params P;
  int64_t nc;
  const int32_t* memcell;        // cells with >= 1 membrane facet
  const double* vol; const double* grad;
  const int32_t* region; const int32_t* nbr; const int32_t* finfo; const int32_t* fmem;
  const double* phi;
  const double* c[MAX_IONS];
  const double* phiM; const double* Ich[MAX_IONS];
  double* rhs[MAX_IONS];
  KNP_HD void operator()(int64_t idx) const {
    const int64_t cell = memcell[idx];
    const int reg = region[cell];
    const double K = vol[cell];
    double r[MAX_IONS][ND];
    for (int k = 0; k < MAX_IONS; ++k)
      for (int v = 0; v < ND; ++v) r[k][v] = 0.0;
    double pme[ND], cme[MAX_IONS][ND], wk[MAX_IONS];
    for (int v = 0; v < ND; ++v) pme[v] = phi[cell * ND + v];
    if (!P.mms)
      for (int k = 0; k < P.N; ++k) {
        wk[k] = P.D[k][reg] * P.z[k] * P.z[k];
        for (int v = 0; v < ND; ++v) cme[k][v] = c[k][cell * ND + v];
      }
    for (int f = 0; f < ND; ++f) {
      const int w = finfo[f * nc + cell];
      if (fi_kind(w) != FK_MEMBRANE) continue;
      const int64_t c2 = nbr[f * nc + cell];
      const int64_t m = fmem[f * nc + cell];
      double gn2 = 0.0;
      for (int x = 0; x < D; ++x) { const double gx = grad[(f * D + x) * nc + cell]; gn2 += gx * gx; }
      const double area = sqrt(gn2) * D * K;
      const double st = fi_ics(w) ? 1.0 : -1.0;
      double pnb[ND];
      for (int v = 0; v < ND; ++v) pnb[v] = (v == f) ? 0.0 : phi[c2 * ND + fi_perm(w, v)];
      double It = 0.0, pM = 0.0;
      if (!P.mms) {
        pM = phiM[m];
        if (P.splitting)
          for (int k = 0; k < P.N; ++k) It += Ich[k][m];
      }
      for (int qd = 0; qd < FacetRule5<D>::NQ; ++qd) {
        double b[D], wq;
        FacetRule5<D>::point(qd, b, wq);
        double lam[ND];   // facet barycentric coordinate of local vertex v (v != f)
        for (int v = 0; v < ND; ++v) lam[v] = (v == f) ? 0.0 : b[v < f ? v : v - 1];
        double dphi = 0.0;
        for (int v = 0; v < ND; ++v) dphi += lam[v] * (pme[v] - pnb[v]);
        dphi *= st;                                       // phi_i - phi_e
        double tk[MAX_IONS], tot = 0.0;
        if (!P.mms)
          for (int k = 0; k < P.N; ++k) {
            double ck = 0.0;
            for (int v = 0; v < ND; ++v) ck += lam[v] * cme[k][v];
            tk[k] = wk[k] * ck;
            tot += tk[k];
          }
        for (int ion = 0; ion < P.N - 1; ++ion) {
          double val;
          if (!P.mms) {
            const double Fz = P.F * P.z[ion];
            const double alpha = tk[ion] / tot;               // solver.py:603
            const double C = alpha * P.C_M / (Fz * P.dt);     // solver.py:606
            // C*g_robin (solver.py:616-622) minus the jump(phi) coupling (:628-629)
            val = C * pM - Ich[ion][m] / Fz + alpha * It / Fz - C * dphi;
          } else {
            val = -P.Csub[ion][reg] * dphi;
          }
          val *= st * wq * area;
          for (int v = 0; v < ND; ++v) r[ion][v] += val * lam[v];
        }
      }
    }
    for (int ion = 0; ion < P.N - 1; ++ion)
      for (int v = 0; v < ND; ++v) rhs[ion][cell * ND + v] += r[ion][v];
  }
};

template <int D>
struct KnpCellKernel {   // one index per cell, all solved ions (host emulation / reference driver)
  static constexpr int ND = D + 1;
  KnpArgs<D> a;
  KNP_HD void operator()(int64_t cell) const {
    const int64_t bs = ND * ND;
    const int reg = a.region[cell];
    double g[ND][D];
    for (int i = 0; i < ND; ++i)
      for (int x = 0; x < D; ++x) g[i][x] = a.grad[(i * D + x) * a.nc + cell];
    const double K = a.vol[cell];
    double gp[D];
    for (int x = 0; x < D; ++x) gp[x] = a.gphi[x * a.nc + cell];
    KnpFacetGeom<D> G[ND];
    knp_facet_geom<D, 0>(a, cell, g, gp, G[0]);
    knp_facet_geom<D, 1>(a, cell, g, gp, G[1]);
    knp_facet_geom<D, 2>(a, cell, g, gp, G[2]);
    if constexpr (D == 3) knp_facet_geom<D, D>(a, cell, g, gp, G[D]);
    for (int ion = 0; ion < a.nion; ++ion) {
      const double Dme = a.P.D[ion][reg];
      double cnl[ND];
      load_cell<ND>(a.cn[ion], cell, cnl);
      double dg[ND][ND], r[ND];
      knp_cell_row<D, 0>(a.P, ion, g, K, Dme, gp, cnl, dg[0], r[0]);
      knp_cell_row<D, 1>(a.P, ion, g, K, Dme, gp, cnl, dg[1], r[1]);
      knp_cell_row<D, 2>(a.P, ion, g, K, Dme, gp, cnl, dg[2], r[2]);
      if constexpr (D == 3) knp_cell_row<D, D>(a.P, ion, g, K, Dme, gp, cnl, dg[D], r[D]);
      for (int f = 0; f < ND; ++f) {
        double O[ND][ND];
        if (f == 0) knp_facet_ion<D, 0>(a.P, ion, G[0], Dme, RegBlocks<ND>{O, dg});
        else if (f == 1) knp_facet_ion<D, 1>(a.P, ion, G[1], Dme, RegBlocks<ND>{O, dg});
        else if (f == 2) knp_facet_ion<D, 2>(a.P, ion, G[2], Dme, RegBlocks<ND>{O, dg});
        else knp_facet_ion<D, D>(a.P, ion, G[D], Dme, RegBlocks<ND>{O, dg});
        double* Of = a.A[ion] + (int64_t)(1 + f) * a.nc * bs + cell * bs;
        for (int i = 0; i < ND; ++i)
          for (int j = 0; j < ND; ++j) Of[i * ND + fi_perm(G[f].w, j)] = O[i][j];
      }
      double* Ad = a.A[ion] + cell * bs;
      for (int i = 0; i < ND; ++i) {
        for (int j = 0; j < ND; ++j) Ad[i * ND + j] = dg[i][j];
        a.rhs[ion][cell * ND + i] = r[i] + (a.load[ion] ? a.load[ion][cell * ND + i] : 0.0);
      }
    }
  }
};

#ifndef KNP_EMU
template <int D, int MINB>
__global__ void __launch_bounds__(ASM_CPB*(D + 1), MINB) knp_assemble_kernel(const KnpArgs<D> a) {
  constexpr int ND = D + 1, BS = ND * ND, PITCH = BS + 1, NT = ASM_CPB * ND;
  __shared__ double sO[ND][ASM_CPB][PITCH];
  __shared__ double sD[ND][ASM_CPB][PITCH];
  __shared__ double sR[ASM_CPB][ND];
  const int t = threadIdx.x;
  const int f = t / ASM_CPB, cl = t - f * ASM_CPB;
  const int64_t cell0 = (int64_t)blockIdx.x * ASM_CPB;
  const int64_t cell = cell0 + cl;
  const int64_t nc = a.nc;
  const bool active = cell < a.nw;
  double g[ND][D], gp[D];
  double K = 0.0;
  int reg = 0;
  KnpFacetGeom<D> G;
  G.w = 0; G.kind = FK_NONE;
  if (active) {
    reg = a.region[cell];
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int x = 0; x < D; ++x) g[i][x] = a.grad[(i * D + x) * nc + cell];
    K = a.vol[cell];
#pragma unroll
    for (int x = 0; x < D; ++x) gp[x] = a.gphi[x * nc + cell];
    switch (f) {   // warp uniform
      case 0: knp_facet_geom<D, 0>(a, cell, g, gp, G); break;
      case 1: knp_facet_geom<D, 1>(a, cell, g, gp, G); break;
      case 2: knp_facet_geom<D, 2>(a, cell, g, gp, G); break;
      default: knp_facet_geom<D, D>(a, cell, g, gp, G); break;
    }
  }
  const int64_t ncell_blk = (a.nw - cell0 < ASM_CPB) ? (a.nw - cell0) : ASM_CPB;
  const int nval = (int)ncell_blk * BS;
  const int nrow = (int)ncell_blk * ND;
  for (int ion = 0; ion < a.nion; ++ion) {
    if (active) {
      const double Dme = a.P.D[ion][reg];
      double cnl[ND];
      load_cell<ND>(a.cn[ion], cell, cnl);
      double row[ND], ri;
      double* myO = sO[f][cl];
      double* myD = sD[f][cl];
      if (G.kind != FK_SIP) {
#pragma unroll
        for (int k = 0; k < BS; ++k) { myO[k] = 0.0; myD[k] = 0.0; }
      }
      const SmemBlocks<ND> blk{myO, myD, G.w};
      switch (f) {
        case 0: knp_facet_ion<D, 0>(a.P, ion, G, Dme, blk);
                knp_cell_row<D, 0>(a.P, ion, g, K, Dme, gp, cnl, row, ri); break;
        case 1: knp_facet_ion<D, 1>(a.P, ion, G, Dme, blk);
                knp_cell_row<D, 1>(a.P, ion, g, K, Dme, gp, cnl, row, ri); break;
        case 2: knp_facet_ion<D, 2>(a.P, ion, G, Dme, blk);
                knp_cell_row<D, 2>(a.P, ion, g, K, Dme, gp, cnl, row, ri); break;
        default: knp_facet_ion<D, D>(a.P, ion, G, Dme, blk);
                knp_cell_row<D, D>(a.P, ion, g, K, Dme, gp, cnl, row, ri); break;
      }
#pragma unroll
      for (int j = 0; j < ND; ++j) myD[f * ND + j] += row[j];
      sR[cl][f] = ri;   // the facets add nothing to the KNP rhs (membrane part: KnpMembraneRhsKernel)
    }
    __syncthreads();
    double* Aion = a.A[ion];
    bool stored = false;
    if constexpr (NT % BS == 0) {
      if (ncell_blk == ASM_CPB) {        // full block: compile-time trip counts and staging addresses (see emi_assemble_kernel)
        constexpr int STEP = NT / BS;
        const int c0 = t / BS, kk = t % BS;
#pragma unroll
        for (int s = 0; s < ND; ++s) {
          double* dst = Aion + (int64_t)(1 + s) * nc * BS + cell0 * BS + t;
#pragma unroll
          for (int k = 0; k < ASM_CPB * BS / NT; ++k) dst[k * NT] = sO[s][c0 + k * STEP][kk];
        }
        double* dA = Aion + cell0 * BS + t;
#pragma unroll
        for (int k = 0; k < ASM_CPB * BS / NT; ++k) {
          const int c = c0 + k * STEP;
          double v = 0.0;
#pragma unroll
          for (int s = 0; s < ND; ++s) v += sD[s][c][kk];
          dA[k * NT] = v;
        }
        a.rhs[ion][cell0 * ND + t] = sR[t / ND][t % ND] + (a.load[ion] ? a.load[ion][cell0 * ND + t] : 0.0);
        stored = true;
      }
    }
    if (!stored) {
    for (int s = 0; s < ND; ++s) {
      double* dst = Aion + (int64_t)(1 + s) * nc * BS + cell0 * BS;
      for (int e = t; e < nval; e += NT) dst[e] = sO[s][e / BS][e % BS];
    }
    {
      double* dA = Aion + cell0 * BS;
      for (int e = t; e < nval; e += NT) {
        const int c = e / BS, k = e % BS;
        double v = 0.0;
#pragma unroll
        for (int s = 0; s < ND; ++s) v += sD[s][c][k];
        dA[e] = v;
      }
    }
    {
      double* dr = a.rhs[ion] + cell0 * ND;
      const double* ld = a.load[ion] ? a.load[ion] + cell0 * ND : nullptr;
      for (int e = t; e < nrow; e += NT) dr[e] = sR[e / ND][e % ND] + (ld ? ld[e] : 0.0);
    }
    }
    __syncthreads();
  }
}
#endif

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
KNP_HD void invert_block(const double* in, double* out) {   // in/out: ND*ND doubles, row major
  double a[ND][ND], b[ND][ND];
  for (int i = 0; i < ND; ++i)
    for (int j = 0; j < ND; ++j) { a[i][j] = in[i * ND + j]; b[i][j] = (i == j); }
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
    for (int j = 0; j < ND; ++j) out[i * ND + j] = b[i][j];
}

template <int ND>
struct BlockInverseKernel {
  const double* blocks; double* inv;
  KNP_HD void operator()(int64_t cell) const { invert_block<ND>(blocks + cell * ND * ND, inv + cell * ND * ND); }
};

#ifndef KNP_EMU
// CUDA driver of the same arithmetic: 128 cells per block, the blocks are read and written as
// contiguous runs through shared memory (pitch ND*ND+1), one thread inverts one block there.
// (The one-index-per-cell functor reads 128-byte strided: 81 us for 420k cells; this: coalesced.)
template <int ND>
__global__ void __launch_bounds__(128) block_inverse_kernel(int64_t nc, const double* __restrict__ blocks,
                                                            double* __restrict__ inv) {
  constexpr int BS = ND * ND, PITCH = BS + 1, CPB = 128;
  __shared__ double sm[CPB][PITCH];
  const int t = threadIdx.x;
  const int64_t cell0 = (int64_t)blockIdx.x * CPB;
  const int ncell = (int)((nc - cell0 < CPB) ? (nc - cell0) : CPB);
  const int nval = ncell * BS;
  const double* src = blocks + cell0 * BS;
  for (int e = t; e < nval; e += CPB) sm[e / BS][e % BS] = src[e];
  __syncthreads();
  if (t < ncell) {
    double in[BS], out[BS];
#pragma unroll
    for (int k = 0; k < BS; ++k) in[k] = sm[t][k];
    invert_block<ND>(in, out);
#pragma unroll
    for (int k = 0; k < BS; ++k) sm[t][k] = out[k];
  }
  __syncthreads();
  double* dst = inv + cell0 * BS;
  for (int e = t; e < nval; e += CPB) dst[e] = sm[e / BS][e % BS];
}
#endif

}  // namespace knp

// qgd_controls.cuh -- kernel K1: control-function / carrier evaluation with all time derivatives.
//
// Every control family on the hot path (GRAPE, BSpline2, FortranBSpline, and CarrierControl of
// those) is LINEAR in its coefficients, so one pcof-independent table
//     table[n][pq][r][theta] = d( p_k(theta)^(r)(t_n) / r! ) / d theta        (pq = 0: p, 1: q)
// gives both what the time stepper consumes,
//     cvals[b][n][pq][r][k] = sum_{theta in slice k} table[n][pq][r][theta] * pcof[b][theta]
//     (= fill_p_mat!/fill_q_mat!, reference src/Controls/Control.jl:99-149), and what the gradient
// needs (eval_grad_{p,q}_derivative!, e.g. src/Controls/FortranBSpline.jl:149-189), for every
// control vector of a batch.  The table is built once per (problem, nsteps, order).
#pragma once
#include "qgd_common.h"

#define QGD_CONTROL_GRAPE 1
#define QGD_CONTROL_BSPLINE2 2
#define QGD_CONTROL_FORTRAN_BSPLINE 3
#define QGD_CONTROL_HOST_TABLE 4
#define QGD_FBS_MAXORDER 20

namespace qgd {

// ---- pppack bsplvb / bsplvd (reference src/Fortran/bsplvb.f:73-90, bsplvd.f:44-110), stateless:
// the Fortran SAVE variables j/deltal/deltar live in the caller's frame.
struct BsplvbState {
  int j;
  double deltal[QGD_FBS_MAXORDER + 1], deltar[QGD_FBS_MAXORDER + 1];
};

__device__ inline void bsplvb(const double* t, int jhigh, int index, double x, int left, double* biatx, BsplvbState& s) {
  if (index == 1) {
    s.j = 1;
    biatx[0] = 1.0;
    if (s.j >= jhigh) return;
  }
  do {
    int jp1 = s.j + 1;
    s.deltar[s.j] = t[left + s.j - 1] - x;
    s.deltal[s.j] = x - t[left + 1 - s.j - 1];
    double saved = 0.0;
    for (int i = 1; i <= s.j; ++i) {
      double term = biatx[i - 1] / (s.deltar[i] + s.deltal[jp1 - i]);
      biatx[i - 1] = saved + s.deltar[i] * term;
      saved = s.deltal[jp1 - i] * term;
    }
    biatx[jp1 - 1] = saved;
    s.j = jp1;
  } while (s.j < jhigh);
}

// dbiatx: column-major, leading dimension k, columns 1..mhigh written.
__device__ inline void bsplvd(const double* t, int k, double x, int left, double* a, double* dbiatx, int nderiv) {
  BsplvbState s;
  s.j = 1;
  int mhigh = max(min(nderiv, k), 1);
  int kp1 = k + 1;
  bsplvb(t, kp1 - mhigh, 1, x, left, dbiatx, s);
  if (mhigh == 1) return;
#define QGD_DB(i, m_) dbiatx[((i)-1) + k * ((m_)-1)]
#define QGD_AA(i, j_) a[((i)-1) + k * ((j_)-1)]
  int ideriv = mhigh;
  for (int m = 2; m <= mhigh; ++m) {
    int jp1mid = 1;
    for (int j = ideriv; j <= k; ++j) { QGD_DB(j, ideriv) = QGD_DB(jp1mid, 1); jp1mid++; }
    ideriv--;
    bsplvb(t, kp1 - ideriv, 2, x, left, dbiatx, s);
  }
  int jlow = 1;
  for (int i = 1; i <= k; ++i) {
    for (int j = jlow; j <= k; ++j) QGD_AA(j, i) = 0.0;
    jlow = i;
    QGD_AA(i, i) = 1.0;
  }
  for (int m = 2; m <= mhigh; ++m) {
    int kp1mm = kp1 - m;
    double fkp1mm = (double)kp1mm;
    int il = left;
    int i = k;
    for (int ld = 1; ld <= kp1mm; ++ld) {
      double factor = fkp1mm / (t[il + kp1mm - 1] - t[il - 1]);
      for (int j = 1; j <= i; ++j) QGD_AA(i, j) = (QGD_AA(i, j) - QGD_AA(i - 1, j)) * factor;
      il--; i--;
    }
    for (i = 1; i <= k; ++i) {
      double sum = 0.0;
      jlow = max(i, m);
      for (int j = jlow; j <= k; ++j) sum = QGD_AA(j, i) * QGD_DB(j, m) + sum;
      QGD_DB(i, m) = sum;
    }
  }
#undef QGD_DB
#undef QGD_AA
}

__device__ inline double ipow_d(double x, int p) {
  double r = 1.0, b = x;
  while (p > 0) { if (p & 1) r *= b; p >>= 1; if (p) b *= b; }
  return r;
}

// CarrierControl.jl:48-66 / 77-95: 4-cycle of d^k/dt^k of cos/sin carriers.  which: 0 = p, 1 = q.
__device__ inline void carrier_vals(double w, double t, int k, int which, double& v1, double& v2) {
  double wk = ipow_d(w, k), s = sin(w * t), c = cos(w * t);
  int r = k & 3;
  if (which == 0) {
    if (r == 0) { v1 = c * wk; v2 = -s * wk; }
    else if (r == 1) { v1 = -s * wk; v2 = -c * wk; }
    else if (r == 2) { v1 = -c * wk; v2 = s * wk; }
    else { v1 = s * wk; v2 = c * wk; }
  } else {
    if (r == 0) { v1 = s * wk; v2 = c * wk; }
    else if (r == 1) { v1 = c * wk; v2 = -s * wk; }
    else if (r == 2) { v1 = -s * wk; v2 = -c * wk; }
    else { v1 = -c * wk; v2 = s * wk; }
  }
}

// Base-control basis: the (at most `nsup`) basis functions that do not vanish at t, their first
// base-local index `i0` (0-based, within one half of the base slice) and their t-derivatives
// bas[i + QGD_FBS_MAXORDER * s] = B_{i0+i}^{(s)}(t), s = 0..nd-1 (un-scaled).
__device__ inline void base_basis(const QgdDevControl& c, double t, int nd, int& i0, int& nsup, double* bas,
                                  double* work_a, double* work_db) {
  for (int e = 0; e < QGD_FBS_MAXORDER * (QGD_MAX_M + 1); ++e) bas[e] = 0.0;
  if (c.type == QGD_CONTROL_GRAPE) {  // grape_control.jl:28-99 (region index; r >= 1 vanishes)
    double width = c.tf / (double)c.n_amp;
    int region = min((int)floor(t / width) + 1, c.n_amp);
    if (region < 1) region = 1;
    i0 = region - 1;
    nsup = 1;
    bas[0] = 1.0;
  } else if (c.type == QGD_CONTROL_BSPLINE2) {  // bspline_control.jl:139-270
    double width = 3.0 * c.dtknot;
    int k = max(3, (int)ceil(t / c.dtknot + 2.0));
    k = min(k, c.D1);
    i0 = k - 3;
    nsup = 3;
    // support order: i0 -> coefficient k-2, i0+1 -> k-1, i0+2 -> k (1-based names of the reference)
    double tc_k = c.dtknot * ((double)k - 1.5), tc_k1 = c.dtknot * ((double)(k - 1) - 1.5), tc_k2 = c.dtknot * ((double)(k - 2) - 1.5);
    double tau0 = (t - tc_k) / width, tau1 = (t - tc_k1) / width, tau2 = (t - tc_k2) / width;
    bas[2] = 9.0 / 8.0 + 4.5 * tau0 + 4.5 * (tau0 * tau0);
    bas[1] = 0.75 - 9.0 * (tau1 * tau1);
    bas[0] = 9.0 / 8.0 - 4.5 * tau2 + 4.5 * (tau2 * tau2);
    if (nd > 1) {
      bas[2 + QGD_FBS_MAXORDER] = (4.5 + 9.0 * tau0) / width;
      bas[1 + QGD_FBS_MAXORDER] = (-18.0 * tau1) / width;
      bas[0 + QGD_FBS_MAXORDER] = (-4.5 + 9.0 * tau2) / width;
    }
    if (nd > 2) {
      bas[2 + 2 * QGD_FBS_MAXORDER] = 9.0 / (width * width);
      bas[1 + 2 * QGD_FBS_MAXORDER] = -18.0 / (width * width);
      bas[0 + 2 * QGD_FBS_MAXORDER] = 9.0 / (width * width);
    }
  } else {  // FortranBSpline.jl:71-107, 267-278
    double x = t / c.tf;
    int k = c.order;
    int left = (int)floor(x * (double)(c.N_distinct - 1) + (double)k);
    left = min(left, c.N_knots - k);
    int off = (int)floor(x * (double)(c.N_distinct - 1) + 1.0);
    off = min(off, c.N_distinct - 1);
    i0 = off - 1;
    nsup = k;
    for (int e = 0; e < k * (QGD_MAX_M + 1); ++e) work_db[e] = 0.0;
    bsplvd(c.knots, k, x, left, work_a, work_db, nd);
    int mhigh = max(min(nd, k), 1);
    for (int s = 0; s < mhigh; ++s) {
      double sc = ipow_d(c.tf, s);
      for (int i = 0; i < k; ++i) bas[i + QGD_FBS_MAXORDER * s] = work_db[i + k * s] / sc;
    }
  }
}

// One thread per (time level n, control k).
__global__ void k_control_table(const QgdDevControl* ctrls, int Nc, int P, int m, int ntimes, const double* times,
                                double t0, double dt, double* table) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ntimes * Nc) return;
  int n = idx / Nc, k = idx % Nc;
  const QgdDevControl c = ctrls[k];
  double t = times ? times[n] : t0 + (double)n * dt;
  const int nd = m + 1;
  double bas[QGD_FBS_MAXORDER * (QGD_MAX_M + 1)];
  double work_a[QGD_FBS_MAXORDER * QGD_FBS_MAXORDER];
  double work_db[QGD_FBS_MAXORDER * (QGD_MAX_M + 1)];
  int i0, nsup;
  base_basis(c, t, nd, i0, nsup, bas, work_a, work_db);
  // zero this control's slice for every (pq, r)
  for (int pq = 0; pq < 2; ++pq)
    for (int r = 0; r < nd; ++r) {
      double* row = table + (((size_t)n * 2 + pq) * nd + r) * P + c.offset;
      for (int e = 0; e < c.ncoeff; ++e) row[e] = 0.0;
    }
  const int half = c.base_ncoeff / 2;
  double fact = 1.0;
  for (int r = 0; r < nd; ++r) {
    if (r > 0) fact *= (double)r;
    double* rowp = table + (((size_t)n * 2 + 0) * nd + r) * P + c.offset;
    double* rowq = table + (((size_t)n * 2 + 1) * nd + r) * P + c.offset;
    if (c.n_carriers == 0) {
      for (int i = 0; i < nsup; ++i) {
        double bv = bas[i + QGD_FBS_MAXORDER * r] / fact;
        rowp[i0 + i] = bv;          // p depends on the first half only
        rowq[half + i0 + i] = bv;   // q on the second half
      }
    } else {
      for (int f = 0; f < c.n_carriers; ++f) {
        double w = c.freqs[f];
        double* gp = rowp + f * c.base_ncoeff;
        double* gq = rowq + f * c.base_ncoeff;
        double binom = 1.0;  // C(r, kk)
        for (int kk = 0; kk <= r; ++kk) {
          if (kk > 0) binom = binom * (double)(r - kk + 1) / (double)kk;
          double p1, p2, q1, q2;
          carrier_vals(w, t, kk, 0, p1, p2);
          carrier_vals(w, t, kk, 1, q1, q2);
          for (int i = 0; i < nsup; ++i) {
            double bv = bas[i + QGD_FBS_MAXORDER * (r - kk)];
            gp[i0 + i] += (p1 * binom) * bv;          // d p^(r) / d (base p-coefficient)
            gp[half + i0 + i] += (p2 * binom) * bv;   // d p^(r) / d (base q-coefficient)
            gq[i0 + i] += (q1 * binom) * bv;
            gq[half + i0 + i] += (q2 * binom) * bv;
          }
        }
        for (int i = 0; i < nsup; ++i) {
          gp[i0 + i] /= fact; gp[half + i0 + i] /= fact;
          gq[i0 + i] /= fact; gq[half + i0 + i] /= fact;
        }
      }
    }
  }
}

// cvals[b][n][pq][r][k] = <table[n][pq][r][slice k], pcof[b][slice k]>.  One thread per output.
__global__ void k_control_values(const QgdDevControl* ctrls, int Nc, int P, int m, int ntimes, const double* table,
                                 const double* pcof, int B, double* cvals) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nd = m + 1;
  size_t total = (size_t)B * ntimes * 2 * nd * Nc;
  if (idx >= total) return;
  int k = (int)(idx % Nc);
  size_t rest = idx / Nc;
  int r = (int)(rest % nd); rest /= nd;
  int pq = (int)(rest % 2); rest /= 2;
  int n = (int)(rest % ntimes);
  int b = (int)(rest / ntimes);
  const int off = ctrls[k].offset, nco = ctrls[k].ncoeff;
  const double* row = table + (((size_t)n * 2 + pq) * nd + r) * P + off;
  const double* pc = pcof + (size_t)b * P + off;
  double s = 0.0;
  for (int e = 0; e < nco; ++e) s = fma(row[e], pc[e], s);
  cvals[idx] = s;
}

}  // namespace qgd

// qgd_common.h -- structures shared by the host layer and the sm_100a kernels.
#pragma once
#include <stdint.h>

#define QGD_MAX_OPS 9  // drift + up to 8 control operators
#define QGD_MAX_M 10   // order <= 20 (the reference's coefficient() uses factorial(2m): Int64 limit)
#ifndef QGD_WARPS_PER_CTA
#define QGD_WARPS_PER_CTA 8  // resident warps (= columns in flight) per SM of the sweep kernels
#endif

// Byte offsets (16-byte aligned) into the operator blob that every sweep CTA stages into shared
// memory with one TMA bulk copy.  Operator k (0 = drift K_s/S_s, 1..Nc = control K_c/S_c) is stored
// as a row-ELL over the UNION pattern of its symmetric and antisymmetric parts, so one gather of
// (u[col], v[col]) feeds both: col[s][N] (int32), valK[s][N], valS[s][N] (Float64), s < L[k].
// Padding entries have col = own row and zero values.
struct QgdOpLayout {
  int n_ops;
  int L[QGD_MAX_OPS];
  int off_col[QGD_MAX_OPS];
  int off_vk[QGD_MAX_OPS];
  int off_vs[QGD_MAX_OPS];
  int LW;        // guard projector W [2N,2N] as row-ELL: col[LW][2N] (int32), val[LW][2N]
  int off_wcol;
  int off_wval;
  int off_pre[2];  // DiagonalHamiltonianPreconditioner (forward, adjoint): d[2N], up[N], lo[N]
  int bytes;
};

// One control (base + optional carriers), device copy.
struct QgdDevControl {
  int type;
  int n_carriers;
  int base_ncoeff;
  int ncoeff;
  int offset;  // start of this control's slice in pcof
  int n_amp, D1, degree, n_basis;
  int order, N_knots, N_distinct;  // FortranBSpline
  double tf;
  double dtknot;  // BSpline2
  const double* freqs;  // [n_carriers] (device)
  const double* knots;  // [N_knots] (device), FortranBSpline
};

struct QgdDevProb {
  int N, N2, Nc, nic, Ness;
  int col0, ncol;  // owned initial-condition columns [col0, col0+ncol)
  int m, nsteps, P;
  int precond;     // QGD_PRECOND_*
  int ops_in_smem;
  double dt, tf, abstol, reltol;
  double a_rhs[QGD_MAX_M + 1];  // c_j dt^j        (build_RHS!)
  double a_lhs[QGD_MAX_M + 1];  // c_j (-dt)^j     (build_LHS!)
  double a_tay[QGD_MAX_M + 1];  // dt^j / j!       (taylor_expand!, on top of the 1/j! already in w_j)
  QgdOpLayout lay;
  const unsigned char* blob;  // operator blob in global memory
  const double* minv[2];      // LU preconditioner: explicit inverse [2N,2N] col-major (forward, adjoint)
  const double* u0;           // [N, nic]
  const double* v0;
  const double* table;        // control basis table [nsteps+1][2][m+1][P], Taylor-scaled
};

// qgd_warp.cuh -- warp-level building blocks of the sweep kernels (one warp = one initial-condition
// column of one control vector).
//
//   * fwd_derivs   : fused Taylor-derivative recursion w_{j+1} = (1/(j+1)) sum_i A_{j-i} w_i over ALL
//                    orders (reference compute_derivatives!, src/hermite.jl:56-101, with
//                    apply_hamiltonian!, :556-588) plus the Hermite combinations build_RHS!/build_LHS!/
//                    taylor_expand! (:394-457) accumulated on the fly.
//   * adj_sweep    : the transpose  (sum_j alpha_j W_j(t))^T x  as ONE reverse sweep through the same
//                    triangular recursion (m(m+1)/2 operator applications instead of the reference's
//                    2^d - 1 per order, src/hermite.jl:225-305), optionally accumulating the gradient
//                    inner products <dH/dtheta w_i, what_j> (reference recursive_magic!,
//                    src/eval_grad_discrete_adjoint.jl:656-726, compute_inner_prod_S!/K! :764-800).
//   * precond_apply: DiagonalHamiltonianPreconditioner / LU (src/preconditioners.jl:44-126).
//   * gmres_warp   : restarted, left-preconditioned GMRES with modified Gram-Schmidt whose dot
//                    products are warp-shuffle reductions, the null-vector residual recurrence and the
//                    Givens least-squares solve of IterativeSolvers.jl -- same
//                    iteration counts as the reference algorithm.
//
// State vectors are distributed over the 32 lanes: lane l owns the level rows r = l + 32 e (e < EL)
// and for each of them BOTH the u and the v component, so the 2x2 preconditioner blocks and the
// real-split structure are lane-local.  Rows >= N hold zeros.
#pragma once
#include "qgd_common.h"

namespace qgd {

constexpr unsigned FULL_MASK = 0xffffffffu;

#define QGD_PRECOND_IDENTITY 0
#define QGD_PRECOND_LU 1
#define QGD_PRECOND_DIAGONAL 2

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

template <int EL>
struct Vec {
  double u[EL];
  double v[EL];
};

template <int EL>
__device__ __forceinline__ void vzero(Vec<EL>& a) {
#pragma unroll
  for (int e = 0; e < EL; ++e) { a.u[e] = 0.0; a.v[e] = 0.0; }
}
template <int EL>
__device__ __forceinline__ void vaxpy(Vec<EL>& y, double a, const Vec<EL>& x) {  // y += a x
#pragma unroll
  for (int e = 0; e < EL; ++e) { y.u[e] = fma(a, x.u[e], y.u[e]); y.v[e] = fma(a, x.v[e], y.v[e]); }
}
template <int EL>
__device__ __forceinline__ void vscale(Vec<EL>& y, double a) {
#pragma unroll
  for (int e = 0; e < EL; ++e) { y.u[e] *= a; y.v[e] *= a; }
}
// Sum of one double per lane, result in every lane, by two FP64 tensor-core DMMA.8x8x4 with a ones operand (measured 75
// vs 175 cycles for the 5-stage shuffle butterfly, profiles/r01_microbench.txt): B operand (4x8, lane l <-> B[l%4][l/4]) = p,
// A = ones: D[i][j] = sum of the 4 lanes of group j, lane l receives groups 2(l%4), 2(l%4)+1; their sum as A operand
// (8x4, lane l <-> A[l/4][l%4]) times ones gives the total in every lane.
__device__ __forceinline__ double warp_sum_dmma(double p) {
  double c0, c1, t0, t1;
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(c0), "=d"(c1) : "d"(1.0), "d"(p), "d"(0.0), "d"(0.0));
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(t0), "=d"(t1) : "d"(c0 + c1), "d"(1.0), "d"(0.0), "d"(0.0));
  (void)t1;
  return t0;
}
#ifndef QGD_GENERIC_DMMA_RED
#define QGD_GENERIC_DMMA_RED 1  // 1: the dot products of the generic sweeps reduce with warp_sum_dmma instead of the shuffle butterfly
#endif

template <int EL>
__device__ __forceinline__ double vdot(const Vec<EL>& a, const Vec<EL>& b) {  // warp-level batched dot
  double s = 0.0;
#pragma unroll
  for (int e = 0; e < EL; ++e) s = fma(a.u[e], b.u[e], s);
#pragma unroll
  for (int e = 0; e < EL; ++e) s = fma(a.v[e], b.v[e], s);
#if QGD_GENERIC_DMMA_RED
  return warp_sum_dmma(s);
#else
  return warp_sum(s);
#endif
}

// Per-warp context: everything lane-invariant.
struct WarpCtx {
  const QgdDevProb* d;
  const unsigned char* ops;  // operator blob (shared memory when it fits, else global)
  int lane;
  double* wv;     // smem [max(m,2)][2N]  vectors other lanes gather from
  double* cv;     // smem [2][m+1][Nc]    control Taylor coefficients of the current time level
  double* hcol;   // smem [2N+1]          current Hessenberg column
  double* nullv;  // smem [2N+1]          left null vector (residual recurrence)
  double* yv;     // smem [2N+1]          least-squares rhs / solution
  double* Vg;     // global [(restart+1)][2N]   Krylov basis of this warp
  double* Hg;     // global packed Hessenberg: column j (0-based) rows 0..j+1 at j(j+3)/2 + i
};

template <int EL>
__device__ __forceinline__ void vload(Vec<EL>& a, const double* p, int N, int lane) {  // p: [2N] contiguous
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    int r = lane + 32 * e;
    bool ok = r < N;
    a.u[e] = ok ? p[r] : 0.0;
    a.v[e] = ok ? p[N + r] : 0.0;
  }
}
template <int EL>
__device__ __forceinline__ void vstore(const Vec<EL>& a, double* p, int N, int lane) {
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    int r = lane + 32 * e;
    if (r < N) { p[r] = a.u[e]; p[N + r] = a.v[e]; }
  }
}
// Streaming (evict-first) variants for the state history: it is written once by the forward sweep and read once by
// the adjoint sweep, and must not push the L2-resident Krylov / Hessenberg workspaces out of the cache.
#ifndef QGD_HIST_CS
#define QGD_HIST_CS 1
#endif
template <int EL>
__device__ __forceinline__ void vload_cs(Vec<EL>& a, const double* p, int N, int lane) {
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    int r = lane + 32 * e;
    bool ok = r < N;
#if QGD_HIST_CS
    a.u[e] = ok ? __ldcs(p + r) : 0.0;
    a.v[e] = ok ? __ldcs(p + N + r) : 0.0;
#else
    a.u[e] = ok ? p[r] : 0.0;
    a.v[e] = ok ? p[N + r] : 0.0;
#endif
  }
}
template <int EL>
__device__ __forceinline__ void vstore_cs(const Vec<EL>& a, double* p, int N, int lane) {
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    int r = lane + 32 * e;
#if QGD_HIST_CS
    if (r < N) { __stcs(p + r, a.u[e]); __stcs(p + N + r, a.v[e]); }
#else
    if (r < N) { p[r] = a.u[e]; p[N + r] = a.v[e]; }
#endif
  }
}

// (K_k x_v, K_k x_u, S_k x_u, S_k x_v)[row r] for operator k from a vector in shared memory.
__device__ __forceinline__ void op_zsums(const WarpCtx& c, int k, int r, const double* x, double& zKu, double& zKv,
                                         double& zSu, double& zSv) {
  const QgdOpLayout& L = c.d->lay;
  const int N = c.d->N;
  const int* col = reinterpret_cast<const int*>(c.ops + L.off_col[k]);
  const double* vk = reinterpret_cast<const double*>(c.ops + L.off_vk[k]);
  const double* vs = reinterpret_cast<const double*>(c.ops + L.off_vs[k]);
  zKu = zKv = zSu = zSv = 0.0;
  const int Lk = L.L[k];
  for (int s = 0; s < Lk; ++s) {
    const int cc = col[s * N + r];
    const double a = vk[s * N + r], b = vs[s * N + r];
    const double xu = x[cc], xv = x[N + cc];
    zKu = fma(a, xu, zKu); zKv = fma(a, xv, zKv);
    zSu = fma(b, xu, zSu); zSv = fma(b, xv, zSv);
  }
}

// Control Taylor coefficients of the current time level: p_k^(d)/d! and q_k^(d)/d!
__device__ __forceinline__ double cvK(const WarpCtx& c, int d, int k) { return c.cv[(0 * (c.d->m + 1) + d) * c.d->Nc + k]; }
__device__ __forceinline__ double cvS(const WarpCtx& c, int d, int k) { return c.cv[(1 * (c.d->m + 1) + d) * c.d->Nc + k]; }

__device__ __forceinline__ void load_cv(const WarpCtx& c, const double* src /*global [2][m+1][Nc]*/) {
  const int n = 2 * (c.d->m + 1) * c.d->Nc;
  __syncwarp();
  for (int i = c.lane; i < n; i += 32) c.cv[i] = src[i];
  __syncwarp();
}

// Forcing of the Taylor recursion at one time level (compute_derivatives!(...; forcing_matrix), reference
// src/hermite.jl:91-95): either an explicit array (eval_forward!(...; forcing), src/forward_evolution.jl:118-129) or
// the forcing of eval_grad_forced (src/eval_grad_forced.jl:82-131) formed on the fly,
//   f_j = sum_{i<=j} d/dtheta [p^(j-i)/(j-i)!] [K w_i]-part + d/dtheta [q^(j-i)/(j-i)!] [S w_i]-part,
// from the Taylor columns w_i of the unforced solution (hb, global memory) and the control basis table.
struct ForcingSrc {
  const double* arr;  // [2N][m] explicit forcing of this time level, or null
  const double* hb;   // [2N][1+m] unforced Taylor columns of this time level, or null
  const double* tp;   // tp[r * P] = d/dtheta p^(r)/r! at this time level; tq likewise
  const double* tq;
  int P, op;          // op: blob index of the control operator theta belongs to
};

// out = sum_j alpha_j w_j with w_0 = x and w_{j+1} = (1/(j+1)) (sum_{i<=j} A_{j-i} w_i + f_j).
// STEP: also guess = sum_j a_tay_j w_j and (if hist != nullptr) hist[:, j] = w_j (history slot).
template <int EL, bool STEP>
__device__ void fwd_derivs(const WarpCtx& c, const Vec<EL>& x, const double* alpha, Vec<EL>& out, Vec<EL>* guess,
                           double* hist, const ForcingSrc* F = nullptr) {
  const QgdDevProb& d = *c.d;
  const int N = d.N, N2 = d.N2, m = d.m, nops = d.lay.n_ops, lane = c.lane;
  __syncwarp();
  vstore(x, c.wv, N, lane);
  out = x;
  vscale(out, alpha[0]);
  if (STEP) {
    *guess = x;  // a_tay[0] = 1
    if (hist) vstore(x, hist, N, lane);
  }
  __syncwarp();
  for (int j = 0; j < m; ++j) {
    Vec<EL> acc;
    const double inv = 1.0 / (double)(j + 1);
#pragma unroll
    for (int e = 0; e < EL; ++e) {
      const int r = lane + 32 * e;
      double au = 0.0, av = 0.0;
      if (r < N) {
        for (int i = 0; i <= j; ++i) {
          const int dd = j - i;
          const double* xi = c.wv + i * N2;
          double zKu, zKv, zSu, zSv;
          if (dd == 0) {  // drift only enters the zeroth Taylor coefficient of A(t)
            op_zsums(c, 0, r, xi, zKu, zKv, zSu, zSv);
            au += zSu + zKv;
            av += zSv - zKu;
          }
          for (int k = 1; k < nops; ++k) {
            op_zsums(c, k, r, xi, zKu, zKv, zSu, zSv);
            const double ck = cvK(c, dd, k - 1), cs = cvS(c, dd, k - 1);
            au = fma(cs, zSu, fma(ck, zKv, au));
            av = fma(cs, zSv, fma(-ck, zKu, av));
          }
        }
        if (F) {
          double fu = 0.0, fv = 0.0;
          if (F->arr) { fu = F->arr[r + (size_t)N2 * j]; fv = F->arr[N + r + (size_t)N2 * j]; }
          if (F->hb) {
            for (int i = j; i >= 0; --i) {
              const int dd = j - i;
              double zKu, zKv, zSu, zSv;
              op_zsums(c, F->op, r, F->hb + (size_t)i * N2, zKu, zKv, zSu, zSv);
              const double pv = F->tp[(size_t)dd * F->P], qv = F->tq[(size_t)dd * F->P];
              fu = fma(pv, zKv, fma(qv, zSu, fu));
              fv = fma(-pv, zKu, fma(qv, zSv, fv));
            }
          }
          au += fu; av += fv;
        }
      }
      acc.u[e] = au * inv;
      acc.v[e] = av * inv;
    }
    vaxpy(out, alpha[j + 1], acc);
    if (STEP) {
      vaxpy(*guess, d.a_tay[j + 1], acc);
      if (hist) vstore(acc, hist + (size_t)(j + 1) * N2, N, lane);
    }
    if (j + 1 < m) {
      vstore(acc, c.wv + (j + 1) * N2, N, lane);
      __syncwarp();
    }
  }
}

// Reverse sweep: what_j = alpha_j x; for j = m-1..0: what_{j-d} -= (1/(j+1)) A_d what_{j+1}, d = 0..j.
// out = what_0 = (sum_j alpha_j W_j)^T x.
// GRAD: with w_i (i < m) the forward Taylor columns at the same time level (hist: global, [2N] per
// column, own rows only) accumulate, per lane,
//   gK[r][k] += IPK_k(w_i, what_{j+1})/(j+1),  gS[r][k] += IPS_k(w_i, what_{j+1})/(j+1),  r = j - i
// (SURVEY A.6).  The caller reduces over lanes and contracts with the control basis table.
template <int EL, bool GRAD>
__device__ void adj_sweep(const WarpCtx& c, const Vec<EL>& x, const double* alpha, Vec<EL>& out, const double* hist,
                          double* gK /*[QGD_MAX_M][QGD_MAX_OPS-1] per lane*/, double* gS) {
  const QgdDevProb& d = *c.d;
  const int N = d.N, N2 = d.N2, m = d.m, nops = d.lay.n_ops, lane = c.lane;
  Vec<EL> what[QGD_MAX_M + 1];
  for (int j = 0; j <= m; ++j) { what[j] = x; vscale(what[j], alpha[j]); }
  for (int j = m - 1; j >= 0; --j) {
    double* zb = c.wv + (j & 1) * N2;  // double buffered gather source
    vstore(what[j + 1], zb, N, lane);
    __syncwarp();
    const double inv = 1.0 / (double)(j + 1);
#pragma unroll
    for (int e = 0; e < EL; ++e) {
      const int r = lane + 32 * e;
      if (r < N) {
        double zKu, zKv, zSu, zSv;
        op_zsums(c, 0, r, zb, zKu, zKv, zSu, zSv);
        what[j].u[e] -= inv * (zSu + zKv);  // d = 0 drift part
        what[j].v[e] -= inv * (zSv - zKu);
        for (int k = 1; k < nops; ++k) {
          op_zsums(c, k, r, zb, zKu, zKv, zSu, zSv);
          for (int dd = 0; dd <= j; ++dd) {
            const double ck = cvK(c, dd, k - 1) * inv, cs = cvS(c, dd, k - 1) * inv;
            what[j - dd].u[e] -= fma(cs, zSu, ck * zKv);
            what[j - dd].v[e] -= fma(cs, zSv, -ck * zKu);
          }
          if (GRAD) {
            for (int i = 0; i <= j; ++i) {
              const double wu = hist[(size_t)i * N2 + r], wvv = hist[(size_t)i * N2 + N + r];
              const int rr = j - i;
              gK[rr * (QGD_MAX_OPS - 1) + (k - 1)] += inv * (wvv * zKu - wu * zKv);
              gS[rr * (QGD_MAX_OPS - 1) + (k - 1)] -= inv * (wu * zSu + wvv * zSv);
            }
          }
        }
      }
    }
  }
  out = what[0];
}

// Left preconditioner  x <- Pl^{-1} x.   dir: 0 forward sweep, 1 adjoint sweep, < 0 none.
template <int EL>
__device__ void precond_apply(const WarpCtx& c, Vec<EL>& x, int dir) {
  const QgdDevProb& d = *c.d;
  if (dir < 0 || d.precond == QGD_PRECOND_IDENTITY) return;
  const int N = d.N, N2 = d.N2, lane = c.lane;
  if (d.precond == QGD_PRECOND_DIAGONAL) {
    // preconditioners.jl:108-126; ratio = lo/d and den = d[N+i] - up*ratio are precomputed on the host
    const double* pd = reinterpret_cast<const double*>(c.ops + d.lay.off_pre[dir]);
    const double* dg = pd; const double* up = pd + N2; const double* ratio = up + N; const double* den = ratio + N;
#pragma unroll
    for (int e = 0; e < EL; ++e) {
      const int r = lane + 32 * e;
      if (r < N) {
        double xv = x.v[e] - x.u[e] * ratio[r];
        xv = xv / den[r];
        double xu = x.u[e] - up[r] * xv;
        xu = xu / dg[r];
        x.u[e] = xu; x.v[e] = xv;
      }
    }
  } else {  // LU: the factorisation is applied as the explicit inverse (one dense mat-vec from L2)
    const double* Mi = d.minv[dir];
    __syncwarp();
    vstore(x, c.wv, N, lane);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < EL; ++e) {
      const int r = lane + 32 * e;
      double su = 0.0, sv = 0.0;
      if (r < N) {
        for (int cc = 0; cc < N2; ++cc) {
          const double xc = c.wv[cc];
          su = fma(Mi[r + (size_t)N2 * cc], xc, su);
          sv = fma(Mi[N + r + (size_t)N2 * cc], xc, sv);
        }
      }
      x.u[e] = su; x.v[e] = sv;
    }
    __syncwarp();
  }
}

// LinearAlgebra.givensAlgorithm for reals (sign convention of LAPACK dlartg)
__device__ __forceinline__ void givens(double f, double g, double& cs, double& sn) {
  if (g == 0.0) { cs = 1.0; sn = 0.0; }
  else if (f == 0.0) { cs = 0.0; sn = 1.0; }
  else {
    double r = sqrt(f * f + g * g);
    cs = f / r; sn = g / r;
    if (fabs(f) > fabs(g) && cs < 0.0) { cs = -cs; sn = -sn; }
  }
}

__device__ __forceinline__ int hoff(int j) { return (j * (j + 3)) >> 1; }

// Least squares min || H[0:width+1, 0:width] y - beta e1 || by Givens rotations, lanes parallel over
// the trailing columns (IterativeSolvers solve_least_squares! + FastHessenberg ldiv!).  y -> c.yv.
__device__ inline void solve_least_squares(const WarpCtx& c, int width, double beta) {
  const int lane = c.lane;
  double* H = c.Hg;
  double* y = c.yv;
  for (int i = lane; i <= width; i += 32) y[i] = (i == 0) ? beta : 0.0;
  __syncwarp();
  for (int i = 0; i < width; ++i) {
    const double hii = H[hoff(i) + i], hi1 = H[hoff(i) + i + 1];
    double cs, sn;
    givens(hii, hi1, cs, sn);
    __syncwarp();
    if (lane == 0) {
      H[hoff(i) + i] = cs * hii + sn * hi1;
      const double yi = y[i], yi1 = y[i + 1];
      y[i] = cs * yi + sn * yi1;
      y[i + 1] = -sn * yi + cs * yi1;
    }
    for (int j = i + 1 + lane; j < width; j += 32) {
      const double a = H[hoff(j) + i], b = H[hoff(j) + i + 1];
      H[hoff(j) + i] = cs * a + sn * b;
      H[hoff(j) + i + 1] = -sn * a + cs * b;
    }
    __syncwarp();
  }
  for (int j = width - 1; j >= 0; --j) {  // trsv('U','N'), column oriented
    const double yj = y[j] / H[hoff(j) + j];
    __syncwarp();
    if (lane == 0) y[j] = yj;
    for (int i = lane; i < j; i += 32) y[i] -= yj * H[hoff(j) + i];
    __syncwarp();
  }
}

#ifndef QGD_GENERIC_PREFETCH
#define QGD_GENERIC_PREFETCH 1  // Krylov basis vectors of the generic sweeps prefetched two ahead (0: one L2 round trip per vector)
#endif
// GMRES (IterativeSolvers.jl gmres_iterable! / iterate).  OP::apply(ctx, in, out).
// reltol < 0: fixed tolerance `abstol` (the time-stepping solves: update_gmres_iterable! never
// refreshes the tolerance); else tol = max(reltol*beta0, abstol) (gmres! driver).
// Returns the number of iterations (loop bodies).
template <int EL, class OP>
__device__ int gmres_warp(const WarpCtx& c, const OP& op, Vec<EL>& x, const Vec<EL>& b, double abstol, double reltol,
                          int restart, int maxiter, int pdir) {
  const int N = c.d->N, N2 = c.d->N2, lane = c.lane;
  Vec<EL> v, w;
  op.apply(c, x, w);
#pragma unroll
  for (int e = 0; e < EL; ++e) { v.u[e] = b.u[e] - w.u[e]; v.v[e] = b.v[e] - w.v[e]; }
  precond_apply(c, v, pdir);
  double beta = sqrt(vdot(v, v));
  vscale(v, 1.0 / beta);
  vstore(v, c.Vg, N, lane);
  const double tol = (reltol < 0.0) ? abstol : fmax(reltol * beta, abstol);
  double cur = beta, res_beta = beta, accum = 1.0;
  __syncwarp();
  if (lane == 0) c.nullv[0] = 1.0;
  __syncwarp();
  int k = 1, it = 0;
  while (it < maxiter && cur > tol) {
    op.apply(c, v, w);  // expand!
    precond_apply(c, w, pdir);
    double dsum = 0.0;
#if !QGD_GENERIC_PREFETCH
    for (int i = 0; i < k; ++i) {  // modified Gram-Schmidt, one warp-shuffle dot per basis vector
      Vec<EL> vi;
      vload(vi, c.Vg + (size_t)i * N2, N, lane);
      const double h = vdot(vi, w);
      if (lane == 0) c.hcol[i] = h;
      vaxpy(w, -h, vi);
      dsum += c.nullv[i] * h;
    }
#else
    {  // modified Gram-Schmidt, one warp-shuffle dot per basis vector (strict: same operations in the same order as the
       // reference).  The basis lives in L2: vectors i+1 and i+2 are in flight while vector i is being reduced, so the
       // chain per vector is the reduction, not an L2 round trip (round 2: the generic sweeps were bound by exactly that).
      Vec<EL> vi, vn, vnn;
      vload(vi, c.Vg, N, lane);
      vload(vn, c.Vg + (size_t)min(1, k - 1) * N2, N, lane);
      for (int i = 0; i < k; ++i) {
        vload(vnn, c.Vg + (size_t)min(i + 2, k - 1) * N2, N, lane);
        const double h = vdot(vi, w);
        if (lane == 0) c.hcol[i] = h;
        vaxpy(w, -h, vi);
        dsum += c.nullv[i] * h;
        vi = vn; vn = vnn;
      }
    }
#endif
    const double nrm = sqrt(vdot(w, w));
    vscale(w, 1.0 / nrm);
    vstore(w, c.Vg + (size_t)k * N2, N, lane);
    if (lane == 0) c.hcol[k] = nrm;
    const double nv = -(dsum / nrm);  // update_residual!
    if (lane == 0) c.nullv[k] = nv;
    accum += nv * nv;
    cur = res_beta / sqrt(accum);
    __syncwarp();
    {
      double* Hc = c.Hg + hoff(k - 1);
      for (int i = lane; i <= k; i += 32) Hc[i] = c.hcol[i];
    }
    k += 1;
    v = w;
    if (k == restart + 1 || cur <= tol) {
      const int width = k - 1;
      __syncwarp();
      solve_least_squares(c, width, beta);
      {  // update_solution!: x += V[:, 1:k-1] y, two basis vectors in flight
        Vec<EL> vj, vn, vnn;
        vload(vj, c.Vg, N, lane);
        vload(vn, c.Vg + (size_t)min(1, width - 1) * N2, N, lane);
        for (int j = 0; j < width; ++j) {
          vload(vnn, c.Vg + (size_t)min(j + 2, width - 1) * N2, N, lane);
          vaxpy(x, c.yv[j], vj);
          vj = vn; vn = vnn;
        }
      }
      k = 1;
      if (cur > tol) {  // restart (residual.current keeps its value, as in the package)
        op.apply(c, x, w);
#pragma unroll
        for (int e = 0; e < EL; ++e) { v.u[e] = b.u[e] - w.u[e]; v.v[e] = b.v[e] - w.v[e]; }
        precond_apply(c, v, pdir);
        beta = sqrt(vdot(v, v));
        vscale(v, 1.0 / beta);
        vstore(v, c.Vg, N, lane);
        accum = 1.0;
        res_beta = beta;
        __syncwarp();
        if (lane == 0) c.nullv[0] = 1.0;
      }
    }
    __syncwarp();
    it += 1;
  }
  return it;
}

// The two step operators as GMRES operands.
template <int EL>
struct FwdOp {  // LHSHolder (src/forward_evolution.jl:583-592)
  __device__ void apply(const WarpCtx& c, const Vec<EL>& in, Vec<EL>& out) const {
    fwd_derivs<EL, false>(c, in, c.d->a_lhs, out, nullptr, nullptr);
  }
};
template <int EL>
struct AdjOp {  // LHSHolderAdjoint (:624-633), via the reverse sweep
  __device__ void apply(const WarpCtx& c, const Vec<EL>& in, Vec<EL>& out) const {
    adj_sweep<EL, false>(c, in, c.d->a_lhs, out, nullptr, nullptr, nullptr);
  }
};

// Guard projector rows: (W x)[own rows] from a vector in shared memory.
template <int EL>
__device__ void guard_apply(const WarpCtx& c, const double* xs, Vec<EL>& out) {
  const QgdDevProb& d = *c.d;
  const int N = d.N, N2 = d.N2, lane = c.lane, LW = d.lay.LW;
  const int* col = reinterpret_cast<const int*>(c.ops + d.lay.off_wcol);
  const double* val = reinterpret_cast<const double*>(c.ops + d.lay.off_wval);
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    const int r = lane + 32 * e;
    double su = 0.0, sv = 0.0;
    if (r < N) {
      for (int s = 0; s < LW; ++s) {
        su = fma(val[s * N2 + r], xs[col[s * N2 + r]], su);
        sv = fma(val[s * N2 + N + r], xs[col[s * N2 + N + r]], sv);
      }
    }
    out.u[e] = su; out.v[e] = sv;
  }
}

}  // namespace qgd

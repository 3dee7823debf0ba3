// qgd_dense.cu -- FP64 tensor-core (DMMA.8x8x4) form of the Taylor-derivative recursion for DENSE Hamiltonians
// (compute_derivatives!, reference src/hermite.jl:56-101, with apply_hamiltonian!, :556-588): where the system and
// control operators are dense and many state columns are advanced together, the recursion
//     W_{j+1} = (1/(j+1)) sum_{i<=j} A_{j-i} W_i,   A_d = [S_d K_d; -K_d S_d],
//     K_d = [d = 0] K_s + sum_k p_k^(d)/d! K_k,   S_d = [d = 0] S_s + sum_k q_k^(d)/d! S_k
// is a real dense contraction  [2N x 2N] x [2N x columns]  (BASELINE north_star).
//
//   k_dense_combine   the per-order operators K_d, S_d of one time level, row-major [d][2][N][N] (one pass over the
//                     Nc + 1 dense operators; they are then read ONCE per Taylor pair instead of once per operator)
//   k_derivs_dense    one CTA = 8 state columns, N / 32 warps; warp w owns the level rows [32 w, 32 w + 32) of BOTH
//                     the u and the v block, so every A fragment (8 x 4 of S_d and of K_d, streamed from L2 with
//                     32-byte-sector loads) feeds two DMMA each; the Taylor columns W_i of the 8 state columns live in
//                     shared memory, transposed and padded ([column][2N + 4]) so that the B-fragment loads are
//                     bank-conflict free; accumulators (4 row tiles x (u, v) x 2) stay in registers across the pair sum.
//
//   k_forward_dense / k_terminal_dense / k_backward_dense   the sweeps of a gradient evaluation around that contraction
//                     (eval_forward!, compute_terminal_condition, eval_adjoint! + accumulate_gradient!): 8 columns per
//                     CTA in lockstep, per-level pre-combined operators (k_dense_combine_levels), GMRES vector work per
//                     warp (one warp per column), gradient inner products as contractions with the raw control operators.
//
// The generic sweep kernels reach 0.4 TFLOP/s on the C4 shape (profiles/r01_c4_generic.json), these 17 TFLOP/s;
// qgd_compute_derivatives, qgd_eval_forward and qgd_discrete_adjoint* route dense problems with N a multiple of 32 here.
#include "qgd_host.h"
#include "qgd_kernels.cuh"

#include <cstdlib>

namespace qgd {

__device__ __forceinline__ void dmma884_acc(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void k_dense_combine(const double* __restrict__ ops /*[Nc+1][2][N][N] row-major, operator 0 = drift*/, int N, int Nc, int m,
                                const double* __restrict__ cv /*[2][m+1][Nc]*/, double* __restrict__ comb /*[m][2][N][N]*/) {
  const size_t nn = (size_t)N * N;
  const size_t total = (size_t)m * 2 * nn;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t e = idx % nn;
    const int ks = (int)((idx / nn) % 2), d = (int)(idx / (2 * nn));
    double s = d == 0 ? ops[(size_t)ks * nn + e] : 0.0;
    for (int k = 0; k < Nc; ++k) s = fma(cv[((size_t)ks * (m + 1) + d) * Nc + k], ops[((size_t)(k + 1) * 2 + ks) * nn + e], s);
    comb[idx] = s;
  }
}

// One Taylor pair: (aU, aV) += A_d B for the 32 level rows of this warp, B = a [2N x 8] block in shared memory
// ([column][S] layout), A_d = [S_d K_d; -K_d S_d] streamed from L2.
__device__ __forceinline__ void dense_pair(const double* __restrict__ comb, int d, int N, int r0, int lane, const double* B, int S,
                                           double (&aU)[4][2], double (&aV)[4][2]) {
  // k mapping of a step of 8 columns: lane ak supplies columns k0 + 2 ak (first DMMA) and k0 + 2 ak + 1 (second DMMA) of A
  // and the same rows of B, so both operands come in with 16-byte loads (the sum over k does not care about the order)
  const int ar = lane >> 2, ak = lane & 3;
  const size_t nn = (size_t)N * N;
  const double* Kd = comb + ((size_t)d * 2 + 0) * nn + (size_t)(r0 + ar) * N + 2 * ak;
  const double* Sd = comb + ((size_t)d * 2 + 1) * nn + (size_t)(r0 + ar) * N + 2 * ak;
  const double* Bu = B + (size_t)ar * S + 2 * ak;
  const double* Bv = Bu + N;
#pragma unroll 2
  for (int k0 = 0; k0 < N; k0 += 8) {
    const double2 bu = *reinterpret_cast<const double2*>(Bu + k0), bv = *reinterpret_cast<const double2*>(Bv + k0);
    const double nbx = -bu.x, nby = -bu.y;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double2 aS = __ldg(reinterpret_cast<const double2*>(Sd + (size_t)8 * t * N + k0));
      const double2 aK = __ldg(reinterpret_cast<const double2*>(Kd + (size_t)8 * t * N + k0));
      dmma884_acc(aU[t][0], aU[t][1], aS.x, bu.x);
      dmma884_acc(aU[t][0], aU[t][1], aK.x, bv.x);
      dmma884_acc(aV[t][0], aV[t][1], aS.x, bv.x);
      dmma884_acc(aV[t][0], aV[t][1], aK.x, nbx);
      dmma884_acc(aU[t][0], aU[t][1], aS.y, bu.y);
      dmma884_acc(aU[t][0], aU[t][1], aK.y, bv.y);
      dmma884_acc(aV[t][0], aV[t][1], aS.y, bv.y);
      dmma884_acc(aV[t][0], aV[t][1], aK.y, nby);
    }
  }
}

// The Taylor recursion of one time level on the 8 state columns of a CTA: Wt[0] (the [column][S] tile of W_0) is
// filled and visible (barrier done by the caller); fills Wt[1..M]; ends with a CTA barrier.  Warp w owns the level
// rows [32 w, 32 w + 32) of the u and the v block; blockDim.x = N.
template <int M>
__device__ __forceinline__ void dense_recursion(const double* __restrict__ comb, int N, double* Wt, int S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = 32 * warp;                 // level rows of this warp
  const int ar = lane >> 2, ak = lane & 3;  // A fragment: row ar, k index ak;  B fragment: k index ak, column ar
#pragma unroll 1
  for (int j = 0; j < M; ++j) {
    double aU[4][2], aV[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) { aU[t][0] = aU[t][1] = aV[t][0] = aV[t][1] = 0.0; }
#pragma unroll 1
    for (int i = 0; i <= j; ++i) {
      dense_pair(comb, j - i, N, r0, lane, Wt + (size_t)i * 8 * S, S, aU, aV);  // W_i is the B operand
    }
    // D fragment: row ar, columns 2 ak, 2 ak + 1
    const double inv = 1.0 / (double)(j + 1);
    double* Wn = Wt + (size_t)(j + 1) * 8 * S;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int row = r0 + 8 * t + ar;
      Wn[(size_t)(2 * ak) * S + row] = aU[t][0] * inv;
      Wn[(size_t)(2 * ak + 1) * S + row] = aU[t][1] * inv;
      Wn[(size_t)(2 * ak) * S + N + row] = aV[t][0] * inv;
      Wn[(size_t)(2 * ak + 1) * S + N + row] = aV[t][1] * inv;
    }
    __syncthreads();
  }
}

// uv: [2N][1+M][ncols], column 0 of every state column given; columns 1..M are written.
template <int M>
__global__ void __launch_bounds__(256, 1) k_derivs_dense(const double* __restrict__ comb, int N, double* uv, int ncols) {
  extern __shared__ __align__(16) double Wt[];  // [(M+1)][8][S]
  const int N2 = 2 * N, S = N2 + 4;
  const int c0 = blockIdx.x * 8;
  // Taylor column 0 of the 8 state columns (zeros beyond ncols)
  for (int idx = threadIdx.x; idx < 8 * N2; idx += blockDim.x) {
    const int col = idx / N2, row = idx % N2;
    Wt[(size_t)col * S + row] = (c0 + col < ncols) ? uv[(size_t)row + (size_t)N2 * (M + 1) * (c0 + col)] : 0.0;
  }
  __syncthreads();
  dense_recursion<M>(comb, N, Wt, S);
  for (int idx = threadIdx.x; idx < M * 8 * N2; idx += blockDim.x) {
    const int row = idx % N2, col = (idx / N2) % 8, jj = 1 + idx / (8 * N2);
    if (c0 + col < ncols) uv[(size_t)row + (size_t)N2 * (jj + (size_t)(M + 1) * (c0 + col))] = Wt[((size_t)jj * 8 + col) * S + row];
  }
}

// Adjoint columns Lambda_jt = W_jt(t)^T x, jt = 1..M (compute_adjoint_derivatives!, reference src/hermite.jl:284-305), each
// by the O(m^2) reverse sweep  what_j = [j = jt] x;  for j = jt-1..0: what_{j-d} -= (1/(j+1)) A_d what_{j+1}, d = 0..j
// (K symmetric, S antisymmetric: A_d^T = -A_d), as the same tensor-core contraction.
template <int M>
__global__ void __launch_bounds__(256, 1) k_derivs_dense_adj(const double* __restrict__ comb, int N, double* uv, int ncols) {
  extern __shared__ __align__(16) double Wt[];  // what[(M+1)][8][S]
  const int N2 = 2 * N, S = N2 + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 8;
  const int r0 = 32 * warp, ar = lane >> 2, ak = lane & 3;
#pragma unroll 1
  for (int jt = 1; jt <= M; ++jt) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < (jt + 1) * 8 * N2; idx += blockDim.x) {
      const int row = idx % N2, col = (idx / N2) % 8, jj = idx / (8 * N2);
      Wt[((size_t)jj * 8 + col) * S + row] =
          (jj == jt && c0 + col < ncols) ? uv[(size_t)row + (size_t)N2 * (M + 1) * (c0 + col)] : 0.0;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = jt - 1; j >= 0; --j) {
      const double inv = 1.0 / (double)(j + 1);
      const double* Y = Wt + (size_t)(j + 1) * 8 * S;
#pragma unroll 1
      for (int dd = 0; dd <= j; ++dd) {
        double aU[4][2], aV[4][2];
#pragma unroll
        for (int t = 0; t < 4; ++t) { aU[t][0] = aU[t][1] = aV[t][0] = aV[t][1] = 0.0; }
        dense_pair(comb, dd, N, r0, lane, Y, S, aU, aV);
        double* Wn = Wt + (size_t)(j - dd) * 8 * S;  // own rows only: no other warp touches them in this stage
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int row = r0 + 8 * t + ar;
          Wn[(size_t)(2 * ak) * S + row] -= aU[t][0] * inv;
          Wn[(size_t)(2 * ak + 1) * S + row] -= aU[t][1] * inv;
          Wn[(size_t)(2 * ak) * S + N + row] -= aV[t][0] * inv;
          Wn[(size_t)(2 * ak + 1) * S + N + row] -= aV[t][1] * inv;
        }
      }
      __syncthreads();  // what_j is complete: it is the B operand of the next stage
    }
    for (int idx = threadIdx.x; idx < 8 * N2; idx += blockDim.x) {
      const int row = idx % N2, col = idx / N2;
      if (c0 + col < ncols) uv[(size_t)row + (size_t)N2 * (jt + (size_t)(M + 1) * (c0 + col))] = Wt[(size_t)col * S + row];
    }
  }
}


// ---- the dense forward sweep ----------------------------------------------------------------------------------
// eval_forward! (reference src/forward_evolution.jl:88-245) for dense Hamiltonians: one CTA marches 8 initial-condition
// columns of one control vector through all time steps in lockstep.  Every operator application -- the explicit
// Taylor columns at t_n and each GMRES matvec with LHS(t_{n+1}) (LHSHolder, :583-592) -- is the CTA-wide tensor-core
// contraction [2N x 2N] x [2N x 8] above with the pre-combined operators of that time level (k_dense_combine_levels);
// the per-column vector work of GMRES (IterativeSolvers gmres_iterable!, SURVEY App. B: modified Gram-Schmidt with
// warp-shuffle dot products, null-vector residual recurrence, Givens least squares) stays per warp, one warp per
// column, exactly as in gmres_warp (qgd_warp.cuh).  The columns of a CTA are a small state machine each (initial /
// restart residual, Arnoldi step, done): a column that has converged idles through the remaining contractions of the
// step.  Krylov basis, Hessenberg matrix, x and b live in global memory (L2); the work vector w lives in the W_1 tile.
struct DenseSweepArgs {
  const double* comb;  // [B][nsteps+1][M][2][N][N]
  double* xs;          // [grid][8][2N]  state / GMRES iterate
  double* bs;          // [grid][8][2N]  right-hand side
  double* xp;          // [grid][8][2N]  adjoint sweep: lambda_{n+1} kept beside lambda_n
  double* aux;         // [grid][8][3][2N+2]  Hessenberg column, null vector, least-squares rhs
};

__global__ void k_dense_combine_levels(const double* __restrict__ ops, int N, int Nc, int m, const double* __restrict__ cvals /*[levels][2][m+1][Nc]*/,
                                       double* __restrict__ comb /*[levels][m][2][N][N]*/) {
  const size_t nn = (size_t)N * N;
  const size_t total = (size_t)m * 2 * nn;
  const double* cv = cvals + (size_t)blockIdx.y * 2 * (m + 1) * Nc;
  double* out = comb + (size_t)blockIdx.y * total;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t e = idx % nn;
    const int ks = (int)((idx / nn) % 2), d = (int)(idx / (2 * nn));
    double s = d == 0 ? ops[(size_t)ks * nn + e] : 0.0;
    for (int k = 0; k < Nc; ++k) s = fma(cv[((size_t)ks * (m + 1) + d) * Nc + k], ops[((size_t)(k + 1) * 2 + ks) * nn + e], s);
    out[idx] = s;
  }
}

enum { DCOL_DONE = 0, DCOL_RESID0 = 1, DCOL_RESID = 2, DCOL_ARNOLDI = 3 };

__device__ __forceinline__ double dense_dot(const double* a, const double* b, int N2, int lane) {
  double s = 0.0;
  for (int r = lane; r < N2; r += 32) s = fma(a[r], b[r], s);
  return warp_sum(s);
}

// Left preconditioner on a vector in shared memory: DiagonalHamiltonianPreconditioner (src/preconditioners.jl:108-126) or
// LUPreconditioner (:44-76) applied as the explicit inverse, one mat-vec from L2 per column in the order of precond_apply
// (qgd_warp.cuh).  Lane l owns the elements l + 32 e.
__device__ __forceinline__ void dense_precond(const QgdDevProb& d, double* w, int lane, int dir) {
  const int N = d.N, N2 = d.N2;
  if (d.precond == QGD_PRECOND_DIAGONAL) {
    const double* pd = reinterpret_cast<const double*>(d.blob + d.lay.off_pre[dir]);
    const double* dg = pd; const double* up = pd + N2; const double* ratio = up + N; const double* den = ratio + N;
    for (int r = lane; r < N; r += 32) {
      double xv = w[N + r] - w[r] * ratio[r];
      xv = xv / den[r];
      double xu = w[r] - up[r] * xv;
      xu = xu / dg[r];
      w[r] = xu; w[N + r] = xv;
    }
    __syncwarp();
  } else if (d.precond == QGD_PRECOND_LU) {
    const double* Mi = d.minv[dir] + lane;  // [2N][2N] column-major
    double acc[16];                         // 2N <= 512
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.0;
    __syncwarp();
#pragma unroll 2
    for (int c = 0; c < N2; ++c) {
      const double xc = w[c];
      const double* mc = Mi + (size_t)N2 * c;
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (32 * e < N2) acc[e] = fma(__ldg(mc + 32 * e), xc, acc[e]);
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if (32 * e < N2) w[lane + 32 * e] = acc[e];
    __syncwarp();
  }
}

// GMRES bookkeeping of the 8 columns of a CTA (shared memory; lane 0 of the owning warp writes)
struct DenseColState {
  int state[8], k[8], it[8];
  double beta[8], cur[8], resb[8], acc[8], tol[8];
};

// Hand the next operand v of column `col` to the contraction: forward sweep -- the W_0 tile; adjoint sweep -- the
// tiles what_j = c_j (-dt)^j v of the reverse recursion (LHSHolderAdjoint, src/forward_evolution.jl:624-633).
template <int M, bool ADJ>
__device__ __forceinline__ void dense_publish(const QgdDevProb& d, double* Wt, int S, int col, int r, double v) {
  if (!ADJ) {
    Wt[(size_t)col * S + r] = v;
  } else {
#pragma unroll
    for (int j = 0; j <= M; ++j) Wt[((size_t)j * 8 + col) * S + r] = d.a_lhs[j] * v;
  }
}

// One GMRES event of one column after a CTA-wide operator application whose result is in `ws` (shared memory): the
// residual of the initial guess / of a restart, or one Arnoldi step with modified Gram-Schmidt, the null-vector
// residual recurrence and, at convergence or restart, the Givens least-squares solve and the solution update
// (gmres_warp, qgd_warp.cuh).  One warp; every vector element r is owned by lane r % 32 throughout.
template <int M, bool ADJ>
__device__ __forceinline__ void dense_column_step(const QgdDevProb& d, const SweepArgs& a, const DenseSweepArgs& ds, DenseColState& cs,
                                                  double* Wt, int S, int col, double* ws, int lane, int restart, double reltol, int pdir) {
  const int N2 = d.N2, maxiter = N2;
  double tol = cs.tol[col];
  const int st = cs.state[col];
  const size_t ws_slot = (size_t)blockIdx.x * 8 + col;
  double* X = ds.xs + ws_slot * N2;
  const double* Bv = ds.bs + ws_slot * N2;
  double* Vg = a.Vws + ws_slot * a.v_stride;
  double* hcol = ds.aux + ws_slot * 3 * (N2 + 2);
  double* nullv = hcol + (N2 + 2);
  WarpCtx c;
  c.d = &d; c.lane = lane; c.Hg = a.Hws + ws_slot * a.h_stride; c.yv = nullv + (N2 + 2);
  int k = cs.k[col], it = cs.it[col], nst = DCOL_ARNOLDI;
  double beta = cs.beta[col], cur = cs.cur[col], resb = cs.resb[col], acc = cs.acc[col];
  if (st != DCOL_ARNOLDI) {  // v_1 = Pl^-1 (b - A x) / beta
    for (int r = lane; r < N2; r += 32) ws[r] = Bv[r] - ws[r];
    __syncwarp();
    if (pdir >= 0) dense_precond(d, ws, lane, pdir);
    beta = sqrt(dense_dot(ws, ws, N2, lane));
    const double inv = 1.0 / beta;
    for (int r = lane; r < N2; r += 32) { const double v = ws[r] * inv; Vg[r] = v; dense_publish<M, ADJ>(d, Wt, S, col, r, v); }
    if (st == DCOL_RESID0) {  // a restart keeps residual.current and the tolerance, as the package does
      cur = beta;
      tol = (reltol < 0.0) ? d.abstol : fmax(reltol * beta, d.abstol);  // reltol < 0: the time-stepping solves
    }
    resb = beta; acc = 1.0; k = 1;
    if (lane == 0) nullv[0] = 1.0;
    if (st == DCOL_RESID0 && !(cur > tol)) nst = DCOL_DONE;
  } else {  // expand!, orthogonalize_and_normalize!, update_residual!
    if (pdir >= 0) dense_precond(d, ws, lane, pdir);
    double dsum = 0.0;
    for (int i = 0; i < k; ++i) {
      const double* vi = Vg + (size_t)i * N2;
      const double hh = dense_dot(vi, ws, N2, lane);
      if (lane == 0) hcol[i] = hh;
      for (int r = lane; r < N2; r += 32) ws[r] = fma(-hh, vi[r], ws[r]);
      dsum += nullv[i] * hh;
    }
    const double nrm = sqrt(dense_dot(ws, ws, N2, lane));
    const double inv = 1.0 / nrm;
    for (int r = lane; r < N2; r += 32) { const double v = ws[r] * inv; Vg[(size_t)k * N2 + r] = v; dense_publish<M, ADJ>(d, Wt, S, col, r, v); }
    const double nv = -(dsum / nrm);
    if (lane == 0) { hcol[k] = nrm; nullv[k] = nv; }
    acc += nv * nv;
    cur = resb / sqrt(acc);
    __syncwarp();
    {
      double* Hc = c.Hg + hoff(k - 1);
      for (int i = lane; i <= k; i += 32) Hc[i] = hcol[i];
    }
    k += 1; it += 1;
    if (k == restart + 1 || !(cur > tol)) {
      const int width = k - 1;
      __syncwarp();
      solve_least_squares(c, width, beta);
      for (int j = 0; j < width; ++j) {  // update_solution!: x += V[:, 1:k-1] y
        const double yj = c.yv[j];
        const double* vj = Vg + (size_t)j * N2;
        for (int r = lane; r < N2; r += 32) X[r] = fma(yj, vj[r], X[r]);
      }
      k = 1;
      if (cur > tol && it < maxiter) {
        nst = DCOL_RESID;
        for (int r = lane; r < N2; r += 32) dense_publish<M, ADJ>(d, Wt, S, col, r, X[r]);
      } else {
        nst = DCOL_DONE;
      }
    } else if (it >= maxiter) {
      nst = DCOL_DONE;
    }
  }
  __syncwarp();
  if (lane == 0) {
    cs.state[col] = nst; cs.k[col] = k; cs.it[col] = it;
    cs.beta[col] = beta; cs.cur[col] = cur; cs.resb[col] = resb; cs.acc[col] = acc; cs.tol[col] = tol;
  }
}

template <int M>
__global__ void __launch_bounds__(256, 1) k_forward_dense(const __grid_constant__ QgdDevProb d, const __grid_constant__ SweepArgs a,
                                                           const __grid_constant__ DenseSweepArgs ds) {
  extern __shared__ __align__(16) double Wt[];  // [(M+1)][8][S]
  __shared__ DenseColState cs;
  const int N = d.N, N2 = d.N2, S = N2 + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int groups = (d.ncol + 7) / 8;
  const int items = a.B * groups;
  const size_t nn = (size_t)N * N, lvl = (size_t)M * 2 * nn;
  const size_t slot_sz = (size_t)N2 * (M + 1);
#pragma unroll 1
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / groups, c0 = (item % groups) * 8;
    const double* comb_b = ds.comb + (size_t)b * (d.nsteps + 1) * lvl;
    for (int col = warp; col < 8; col += nwarps) {
      const int cl = c0 + col;
      double* X = ds.xs + ((size_t)blockIdx.x * 8 + col) * N2;
      for (int r = lane; r < N; r += 32) {
        X[r] = cl < d.ncol ? d.u0[r + (size_t)N * (d.col0 + cl)] : 0.0;
        X[N + r] = cl < d.ncol ? d.v0[r + (size_t)N * (d.col0 + cl)] : 0.0;
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int n = 0; n <= d.nsteps; ++n) {
      // ---- explicit side: Taylor columns of x at t_n, history slot, right-hand side and initial guess
      for (int col = warp; col < 8; col += nwarps) {
        const double* X = ds.xs + ((size_t)blockIdx.x * 8 + col) * N2;
        for (int r = lane; r < N2; r += 32) Wt[(size_t)col * S + r] = X[r];
      }
      __syncthreads();
      dense_recursion<M>(comb_b + (size_t)n * lvl, N, Wt, S);
      for (int col = warp; col < 8; col += nwarps) {
        const int cl = c0 + col;
        const bool valid = cl < d.ncol;
        double* X = ds.xs + ((size_t)blockIdx.x * 8 + col) * N2;
        double* Bv = ds.bs + ((size_t)blockIdx.x * 8 + col) * N2;
        double* slot = nullptr;
        if (valid && a.history && n % a.save_every == 0)
          slot = a.history + slot_sz * ((size_t)(n / a.save_every) + (size_t)a.nslots * ((size_t)cl + (size_t)d.ncol * b));
        for (int r = lane; r < N2; r += 32) {
          const double w0 = Wt[(size_t)col * S + r];
          double rhs = w0 * d.a_rhs[0], guess = w0;
          if (slot) __stcs(slot + r, w0);
#pragma unroll
          for (int j = 1; j <= M; ++j) {
            const double wj = Wt[((size_t)j * 8 + col) * S + r];
            rhs = fma(d.a_rhs[j], wj, rhs);
            guess = fma(d.a_tay[j], wj, guess);
            if (slot) __stcs(slot + (size_t)j * N2 + r, wj);
          }
          if (n < d.nsteps) {
            Bv[r] = rhs; X[r] = guess;
            Wt[(size_t)col * S + r] = guess;  // x0 = Taylor expansion: the first contraction forms the initial residual
          } else if (valid) {
            a.final_state[(size_t)r + (size_t)N2 * ((size_t)cl + (size_t)d.ncol * b)] = w0;
          }
        }
        if (lane == 0) { cs.state[col] = valid ? DCOL_RESID0 : DCOL_DONE; cs.it[col] = 0; cs.k[col] = 1; }
      }
      __syncthreads();
      if (n == d.nsteps) break;
      // ---- implicit side: LHS(t_{n+1}) x = rhs by GMRES, the 8 columns in lockstep
      const double* comb1 = comb_b + (size_t)(n + 1) * lvl;
#pragma unroll 1
      while (true) {
        dense_recursion<M>(comb1, N, Wt, S);
        for (int col = warp; col < 8; col += nwarps) {
          if (cs.state[col] == DCOL_DONE) continue;
          const double* xin = Wt + (size_t)col * S;
          double* ws = Wt + (size_t)(8 + col) * S;   // work vector w (the W_1 tile of this column)
          for (int r = lane; r < N2; r += 32) {      // out = sum_j c_j (-dt)^j W_j  (build_LHS!, src/hermite.jl:435-457)
            double o = xin[r] * d.a_lhs[0];
#pragma unroll
            for (int j = 1; j <= M; ++j) o = fma(d.a_lhs[j], Wt[((size_t)j * 8 + col) * S + r], o);
            ws[r] = o;
          }
          __syncwarp();
          dense_column_step<M, false>(d, a, ds, cs, Wt, S, col, ws, lane, N2, -1.0, 0);
        }
        __syncthreads();
        int any = 0;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) any |= cs.state[cc];
        if (!any) break;
      }
      if (a.iters)
        for (int col = warp; col < 8; col += nwarps)
          if (lane == 0 && c0 + col < d.ncol)
            a.iters[(size_t)n + (size_t)d.nsteps * ((size_t)(c0 + col) + (size_t)d.ncol * b)] = cs.it[col];
      // the barrier after the next explicit-side fill orders these reads before the next writes of cs.it
    }
  }
}

// ---- the dense adjoint sweep ----------------------------------------------------------------------------------
// Raw products of one control operator with the tile Y: zKu = K Y_u, zKv = K Y_v, zSu = S Y_u, zSv = S Y_v for the 32
// level rows of this warp (compute_inner_prod_S!/K!, src/eval_grad_discrete_adjoint.jl:764-800, as contractions).
__device__ __forceinline__ void dense_raw4(const double* __restrict__ Kk, const double* __restrict__ Sk, int N, int r0, int lane,
                                           const double* B, int S, double (&zKu)[4][2], double (&zKv)[4][2], double (&zSu)[4][2],
                                           double (&zSv)[4][2]) {
  const int ar = lane >> 2, ak = lane & 3;
  const double* Kp = Kk + (size_t)(r0 + ar) * N + 2 * ak;  // k mapping and 16-byte loads as in dense_pair
  const double* Sp = Sk + (size_t)(r0 + ar) * N + 2 * ak;
  const double* Bu = B + (size_t)ar * S + 2 * ak;
  const double* Bv = Bu + N;
#pragma unroll 1
  for (int k0 = 0; k0 < N; k0 += 8) {
    const double2 bu = *reinterpret_cast<const double2*>(Bu + k0), bv = *reinterpret_cast<const double2*>(Bv + k0);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double2 aS = __ldg(reinterpret_cast<const double2*>(Sp + (size_t)8 * t * N + k0));
      const double2 aK = __ldg(reinterpret_cast<const double2*>(Kp + (size_t)8 * t * N + k0));
      dmma884_acc(zSu[t][0], zSu[t][1], aS.x, bu.x);
      dmma884_acc(zKv[t][0], zKv[t][1], aK.x, bv.x);
      dmma884_acc(zSv[t][0], zSv[t][1], aS.x, bv.x);
      dmma884_acc(zKu[t][0], zKu[t][1], aK.x, bu.x);
      dmma884_acc(zSu[t][0], zSu[t][1], aS.y, bu.y);
      dmma884_acc(zKv[t][0], zKv[t][1], aK.y, bv.y);
      dmma884_acc(zSv[t][0], zSv[t][1], aS.y, bv.y);
      dmma884_acc(zKu[t][0], zKu[t][1], aK.y, bu.y);
    }
  }
}

// Reverse sweep (sum_j alpha_j W_j(t))^T x on the 8 columns of a CTA: the tiles what_j = alpha_j x are filled and
// visible; for j = M-1..0: what_{j-d} -= (1/(j+1)) A_d what_{j+1}, d = 0..j (A_d^T = -A_d); the result is the what_0
// tile; ends with a CTA barrier.  GRAD: also the gradient inner products of this time level (recursive_magic!,
// src/eval_grad_discrete_adjoint.jl:656-726, as in adj_sweep of qgd_warp.cuh),
//   gK[j-i][k] += <[0 K_k; -K_k 0] w_i, what_{j+1}> / (j+1),  gS[j-i][k] += <[S_k 0; 0 S_k] w_i, what_{j+1}> / (j+1),
// per column, with w_i the forward Taylor columns of the level (history); every warp adds the part of its level rows to
// its own slot gp[2][M][Nc][8] (fixed summation order; the owner warp of a column adds the slots up).
template <int M, bool GRAD>
__device__ __forceinline__ void dense_reverse(const double* __restrict__ comb, int N, double* Wt, int S, const double* __restrict__ ops,
                                              int Nc, const double* hist0, size_t hist_col_stride, int ncols_valid, double* gp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = 32 * warp, ar = lane >> 2, ak = lane & 3, N2 = 2 * N;
  const size_t nn = (size_t)N * N;
#pragma unroll 1
  for (int j = M - 1; j >= 0; --j) {
    const double inv = 1.0 / (double)(j + 1);
    const double* Y = Wt + (size_t)(j + 1) * 8 * S;
#pragma unroll 1
    for (int dd = 0; dd <= j; ++dd) {
      double aU[4][2], aV[4][2];
#pragma unroll
      for (int t = 0; t < 4; ++t) { aU[t][0] = aU[t][1] = aV[t][0] = aV[t][1] = 0.0; }
      dense_pair(comb, dd, N, r0, lane, Y, S, aU, aV);
      double* Wn = Wt + (size_t)(j - dd) * 8 * S;  // own rows only: no other warp touches them in this stage
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int row = r0 + 8 * t + ar;
        Wn[(size_t)(2 * ak) * S + row] -= aU[t][0] * inv;
        Wn[(size_t)(2 * ak + 1) * S + row] -= aU[t][1] * inv;
        Wn[(size_t)(2 * ak) * S + N + row] -= aV[t][0] * inv;
        Wn[(size_t)(2 * ak + 1) * S + N + row] -= aV[t][1] * inv;
      }
    }
    if (GRAD) {
#pragma unroll 1
      for (int k = 0; k < Nc; ++k) {
        double zKu[4][2], zKv[4][2], zSu[4][2], zSv[4][2];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          zKu[t][0] = zKu[t][1] = zKv[t][0] = zKv[t][1] = 0.0;
          zSu[t][0] = zSu[t][1] = zSv[t][0] = zSv[t][1] = 0.0;
        }
        dense_raw4(ops + ((size_t)(k + 1) * 2 + 0) * nn, ops + ((size_t)(k + 1) * 2 + 1) * nn, N, r0, lane, Y, S, zKu, zKv, zSu, zSv);
#pragma unroll 1
        for (int i = 0; i <= j; ++i) {
          double pK[2] = {0.0, 0.0}, pS[2] = {0.0, 0.0};
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int col = 2 * ak + cc;
            if (col < ncols_valid) {
              const double* wi = hist0 + hist_col_stride * col + (size_t)i * N2;
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int row = r0 + 8 * t + ar;
                const double wu = __ldg(wi + row), wv = __ldg(wi + N + row);
                pK[cc] += wv * zKu[t][cc] - wu * zKv[t][cc];
                pS[cc] += wu * zSu[t][cc] + wv * zSv[t][cc];
              }
            }
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              pK[cc] += __shfl_xor_sync(FULL_MASK, pK[cc], o);
              pS[cc] += __shfl_xor_sync(FULL_MASK, pS[cc], o);
            }
          }
          if (ar == 0) {
            const int rr = j - i;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              gp[(((size_t)0 * M + rr) * Nc + k) * 8 + 2 * ak + cc] += inv * pK[cc];
              gp[(((size_t)1 * M + rr) * Nc + k) * 8 + 2 * ak + cc] -= inv * pS[cc];
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

// eval_adjoint! + accumulate_gradient! (src/forward_evolution.jl:352-483, src/eval_grad_discrete_adjoint.jl:582-726)
// for dense Hamiltonians; the structure of k_backward (qgd_kernels.cuh) with 8 columns per CTA in lockstep.
template <int M>
__global__ void __launch_bounds__(256, 1) k_backward_dense(const __grid_constant__ QgdDevProb d, const __grid_constant__ SweepArgs a,
                                                            const __grid_constant__ DenseSweepArgs ds, const double* __restrict__ ops,
                                                            const QgdDevControl* __restrict__ ctrls) {
  extern __shared__ __align__(16) double Wt[];  // [(M+1)][8][S] | gpart [nwarps][2][M][Nc][8] | gred [2][M][Nc][8]
  __shared__ DenseColState cs;
  const int N = d.N, N2 = d.N2, S = N2 + 4, Nc = d.Nc, P = d.P, Nt = d.nsteps + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int G = 2 * M * Nc * 8;
  double* gpart = Wt + (size_t)(M + 1) * 8 * S;
  double* gred = gpart + (size_t)nwarps * G;
  double* gp = gpart + (size_t)warp * G;
  const int groups = (d.ncol + 7) / 8;
  const int items = a.B * groups;
  const size_t nn = (size_t)N * N, lvl = (size_t)M * 2 * nn;
  const size_t slot_sz = (size_t)N2 * (M + 1), tab_stride = (size_t)2 * (M + 1) * P;
  const size_t hist_col_stride = slot_sz * Nt;

  // what_j = s1 al1_j x1 (+ s2 al2_j x2) for the owned columns; zero the gradient slot of this warp
  auto fill_tiles = [&](const double* al1, double s1, const double* base1, const double* al2, double s2, const double* base2) {
    for (int col = warp; col < 8; col += nwarps) {
      const double* X1 = base1 + ((size_t)blockIdx.x * 8 + col) * N2;
      const double* X2 = base2 ? base2 + ((size_t)blockIdx.x * 8 + col) * N2 : nullptr;
      for (int r = lane; r < N2; r += 32) {
        const double x1 = X1[r], x2 = X2 ? X2[r] : 0.0;
#pragma unroll
        for (int j = 0; j <= M; ++j) {
          double t = (s1 * al1[j]) * x1;
          if (X2) t = fma(s2 * al2[j], x2, t);
          Wt[((size_t)j * 8 + col) * S + r] = t;
        }
      }
    }
    for (int i = lane; i < G; i += 32) gp[i] = 0.0;
  };
  // owner warps: add the per-warp slots up (fixed order)
  auto reduce_g = [&]() {
    for (int col = warp; col < 8; col += nwarps)
      for (int i = lane; i < 2 * M * Nc; i += 32) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += gpart[(size_t)w * G + (size_t)i * 8 + col];
        gred[(size_t)i * 8 + col] = s;
      }
  };
  // grad[theta] -= sum_r d/dtheta p^(r)/r! gK[r][k(theta)] + d/dtheta q^(r)/r! gS[r][k(theta)]   (accumulate_grad)
  auto accumulate = [&](int b, int c0, const double* table_n) {
    for (int col = warp; col < 8; col += nwarps) {
      const int cl = c0 + col;
      if (cl >= d.ncol) continue;
      double* gacc = a.gradcol + (size_t)P * ((size_t)cl + (size_t)d.ncol * b);
      for (int k = 0; k < Nc; ++k) {
        const int off = ctrls[k].offset, nco = ctrls[k].ncoeff;
        for (int t = lane; t < nco; t += 32) {
          double s = 0.0;
          for (int r = 0; r < M; ++r) {
            s = fma(table_n[((size_t)0 * (M + 1) + r) * P + off + t], gred[(((size_t)0 * M + r) * Nc + k) * 8 + col], s);
            s = fma(table_n[((size_t)1 * (M + 1) + r) * P + off + t], gred[(((size_t)1 * M + r) * Nc + k) * 8 + col], s);
          }
          gacc[off + t] -= s;
        }
      }
    }
  };

#pragma unroll 1
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / groups, c0 = (item % groups) * 8;
    const double* comb_b = ds.comb + (size_t)b * Nt * lvl;
    const double* hist_b = a.history + hist_col_stride * ((size_t)c0 + (size_t)d.ncol * b);  // column c0, level 0
    const int ncv = d.ncol - c0;
    for (int col = warp; col < 8; col += nwarps) {
      const int cl = c0 + col;
      const bool valid = cl < d.ncol;
      double* X = ds.xs + ((size_t)blockIdx.x * 8 + col) * N2;
      const double* term = a.terminal + (size_t)N2 * ((size_t)(d.col0 + cl) + (size_t)d.nic * b);
      double* lam0 = (valid && a.lambda0) ? a.lambda0 + (size_t)N2 * Nt * ((size_t)cl + (size_t)d.ncol * b) : nullptr;
      for (int r = lane; r < N2; r += 32) {
        const double x = valid ? term[r] : 0.0;
        X[r] = x;
        if (lam0) lam0[(size_t)N2 * d.nsteps + r] = x;
      }
      if (valid) {
        double* gacc = a.gradcol + (size_t)P * ((size_t)cl + (size_t)d.ncol * b);
        for (int t = lane; t < P; t += 32) gacc[t] = 0.0;
      }
    }
    __syncwarp();
    // The gradient is linear in the tiles of the reverse sweep, so the two sweeps of a time level L -- the implicit side of
    // step L-1 (alpha_j = -c_j (-dt)^j, lambda_L) and the explicit side of step L (alpha_j = c_j dt^j, lambda_{L+1}) -- run
    // as ONE sweep over what_j = c_j dt^j lambda_{L+1} - c_j (-dt)^j lambda_L once lambda_L is known: per step one plain
    // sweep (the right-hand side R(t_n)^T lambda_{n+1}) and one gradient sweep instead of two gradient sweeps.
    auto grad_sweep = [&](int level) {
      __syncthreads();
      dense_reverse<M, true>(comb_b + (size_t)level * lvl, N, Wt, S, ops, Nc, hist_b + slot_sz * level, hist_col_stride, ncv, gp);
      reduce_g();
      __syncthreads();
      accumulate(b, c0, d.table + (size_t)level * tab_stride);
    };
    fill_tiles(d.a_lhs, -1.0, ds.xs, nullptr, 0.0, nullptr);  // level N: implicit side only
    grad_sweep(d.nsteps);
#pragma unroll 1
    for (int n = d.nsteps - 1; n >= 0; --n) {
      if (n == 0) {  // level 0: explicit side only, no solve
        fill_tiles(d.a_rhs, 1.0, ds.xs, nullptr, 0.0, nullptr);
        grad_sweep(0);
        __syncthreads();
        break;
      }
      // ---- explicit side: what_0 = R(t_n)^T lambda_{n+1}
      fill_tiles(d.a_rhs, 1.0, ds.xs, nullptr, 0.0, nullptr);
      __syncthreads();
      dense_reverse<M, false>(comb_b + (size_t)n * lvl, N, Wt, S, nullptr, 0, nullptr, 0, 0, nullptr);
      // ---- lambda_n: LHS(t_n)^T lambda_n = R(t_n)^T lambda_{n+1} + f_n, x0 = lambda_{n+1} (forward_evolution.jl:450)
      for (int col = warp; col < 8; col += nwarps) {
        const int cl = c0 + col;
        const bool valid = cl < d.ncol;
        double* Bv = ds.bs + ((size_t)blockIdx.x * 8 + col) * N2;
        const double* X = ds.xs + ((size_t)blockIdx.x * 8 + col) * N2;
        const double* wn = hist_b + hist_col_stride * col + slot_sz * n;  // w_n (Taylor column 0)
        const int* wcol = reinterpret_cast<const int*>(d.blob + d.lay.off_wcol);
        const double* wval = reinterpret_cast<const double*>(d.blob + d.lay.off_wval);
        const double fsc = -2.0 * d.dt / d.tf;  // guard forcing f_n = -(2 dt/tf) W w_n (interior point: weight 1)
        for (int r = lane; r < N2; r += 32) {
          double g = 0.0;
          if (valid)
            for (int s = 0; s < d.lay.LW; ++s) g = fma(wval[(size_t)s * N2 + r], __ldg(wn + wcol[(size_t)s * N2 + r]), g);
          Bv[r] = fma(fsc, g, Wt[(size_t)col * S + r]);
        }
        double* Xp = ds.xp + ((size_t)blockIdx.x * 8 + col) * N2;
        for (int r = lane; r < N2; r += 32) { const double x = X[r]; Xp[r] = x; dense_publish<M, true>(d, Wt, S, col, r, x); }
        if (lane == 0) { cs.state[col] = valid ? DCOL_RESID0 : DCOL_DONE; cs.it[col] = 0; cs.k[col] = 1; }
      }
      __syncthreads();
      const double* combn = comb_b + (size_t)n * lvl;
#pragma unroll 1
      while (true) {
        dense_reverse<M, false>(combn, N, Wt, S, nullptr, 0, nullptr, 0, 0, nullptr);
        for (int col = warp; col < 8; col += nwarps) {
          if (cs.state[col] == DCOL_DONE) continue;
          dense_column_step<M, true>(d, a, ds, cs, Wt, S, col, Wt + (size_t)col * S, lane, N2, -1.0, 1);
        }
        __syncthreads();
        int any = 0;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) any |= cs.state[cc];
        if (!any) break;
      }
      for (int col = warp; col < 8; col += nwarps) {
        const int cl = c0 + col;
        if (cl >= d.ncol) continue;
        const double* X = ds.xs + ((size_t)blockIdx.x * 8 + col) * N2;
        if (a.lambda0) {
          double* lam0 = a.lambda0 + (size_t)N2 * ((size_t)n + (size_t)Nt * ((size_t)cl + (size_t)d.ncol * b));
          for (int r = lane; r < N2; r += 32) lam0[r] = X[r];
        }
        if (a.iters && lane == 0) a.iters[(size_t)n + (size_t)d.nsteps * ((size_t)cl + (size_t)d.ncol * b)] = cs.it[col];
      }
      // ---- gradient of time level n: explicit side with lambda_{n+1}, implicit side with lambda_n, one sweep
      fill_tiles(d.a_rhs, 1.0, ds.xp, d.a_lhs, -1.0, ds.xs);
      grad_sweep(n);
    }
  }
}


// Infidelity and terminal condition (infidelity_real, src/infidelity.jl:7-18; compute_terminal_condition,
// src/eval_grad_discrete_adjoint.jl:1-67) for dense Hamiltonians: LHS(tf)^T lambda_N = (2/N_ess^2)(<psi_N,R> R +
// <psi_N,T> T) + f_N, un-preconditioned GMRES with restart 20 and tol = max(reltol beta_0, abstol) as the reference's
// gmres! call, the operator application being the CTA-wide tensor-core reverse recursion.
//   sequential (default): one CTA per control vector solves the columns one after the other and, like the reference,
//     leaves the solution of column i-1 in the buffer as the initial guess of column i (:60-64) -- same iteration counts
//     and the same lambda_N as the reference; one of the 8 contraction columns is used.
//   parallel (QGD_DENSE_TERMINAL_PARALLEL): 8 columns per CTA in lockstep, every column from a zero guess -- lambda_N
//     then agrees with the reference to the GMRES tolerance only.
template <int M>
__global__ void __launch_bounds__(256, 1) k_terminal_dense(const __grid_constant__ QgdDevProb d, const __grid_constant__ SweepArgs a,
                                                            const __grid_constant__ DenseSweepArgs ds, const int sequential) {
  extern __shared__ __align__(16) double Wt[];  // [(M+1)][8][S]
  __shared__ DenseColState cs;
  __shared__ double s_red[2][8];
  const int N = d.N, N2 = d.N2, S = N2 + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int groups = sequential ? 1 : (d.nic + 7) / 8;
  const int items = a.B * groups;
  const size_t nn = (size_t)N * N, lvl = (size_t)M * 2 * nn;
  const int restart = N2 < 20 ? N2 : 20;
  const int nsub = sequential ? d.nic : 1, nact = sequential ? 1 : 8;
#pragma unroll 1
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / groups;
    const double* psi = a.final_all + (size_t)N2 * d.nic * b;
    // <psi_N, R> and <psi_N, T>, T = [R_v; -R_u], over ALL columns
    double dR = 0.0, dT = 0.0;
    for (int col = 0; col < d.nic; ++col) {
      const double* p = psi + (size_t)N2 * col;
      const double* R = a.target + (size_t)N2 * col;
      for (int r = threadIdx.x; r < N; r += blockDim.x) {
        const double pu = p[r], pv = p[N + r], Ru = R[r], Rv = R[N + r];
        dR += pu * Ru + pv * Rv;
        dT += pu * Rv - pv * Ru;
      }
    }
    dR = warp_sum(dR);
    dT = warp_sum(dT);
    __syncthreads();  // previous item: everybody is done with s_red and cs
    if (lane == 0) { s_red[0][warp] = dR; s_red[1][warp] = dT; }
    __syncthreads();
    dR = 0.0; dT = 0.0;
    for (int w = 0; w < nwarps; ++w) { dR += s_red[0][w]; dT += s_red[1][w]; }
    const double ness2 = (double)d.Ness * (double)d.Ness;
    if (item % groups == 0 && threadIdx.x == 0) a.infidelity[b] = 1.0 - (dR * dR + dT * dT) / ness2;
    const double sc = 2.0 / ness2;
    const double fsc = -2.0 * d.dt / d.tf * 0.5;  // forcing[:, end, :]: trapezoid weight 1/2
#pragma unroll 1
   for (int sub = 0; sub < nsub; ++sub) {
    const int c0 = sequential ? sub : (item % groups) * 8;
    for (int col = warp; col < 8; col += nwarps) {
      const int cg = c0 + col;
      const bool valid = col < nact && cg < d.nic;
      double* X = ds.xs + ((size_t)blockIdx.x * 8 + col) * N2;
      double* Bv = ds.bs + ((size_t)blockIdx.x * 8 + col) * N2;
      const double* p = psi + (size_t)N2 * cg;
      const double* R = a.target + (size_t)N2 * cg;
      const int* wcol = reinterpret_cast<const int*>(d.blob + d.lay.off_wcol);
      const double* wval = reinterpret_cast<const double*>(d.blob + d.lay.off_wval);
      for (int r = lane; r < N; r += 32) {
        double gu = 0.0, gv = 0.0, Ru = 0.0, Rv = 0.0;
        if (valid) {
          for (int s = 0; s < d.lay.LW; ++s) {
            gu = fma(wval[(size_t)s * N2 + r], p[wcol[(size_t)s * N2 + r]], gu);
            gv = fma(wval[(size_t)s * N2 + N + r], p[wcol[(size_t)s * N2 + N + r]], gv);
          }
          Ru = R[r]; Rv = R[N + r];
        }
        Bv[r] = (dR * Ru + dT * Rv) * sc + fsc * gu;
        Bv[N + r] = (dR * Rv + dT * (-Ru)) * sc + fsc * gv;
        if (!sequential || sub == 0) { X[r] = 0.0; X[N + r] = 0.0; }
      }
      for (int r = lane; r < N2; r += 32) dense_publish<M, true>(d, Wt, S, col, r, valid ? X[r] : 0.0);
      if (lane == 0) { cs.state[col] = valid ? DCOL_RESID0 : DCOL_DONE; cs.it[col] = 0; cs.k[col] = 1; }
    }
    __syncthreads();
    const double* combN = ds.comb + ((size_t)b * (d.nsteps + 1) + d.nsteps) * lvl;  // controls at t = tf
#pragma unroll 1
    while (true) {
      dense_reverse<M, false>(combN, N, Wt, S, nullptr, 0, nullptr, 0, 0, nullptr);
      for (int col = warp; col < 8; col += nwarps) {
        if (cs.state[col] == DCOL_DONE) continue;
        dense_column_step<M, true>(d, a, ds, cs, Wt, S, col, Wt + (size_t)col * S, lane, restart, d.reltol, -1);
      }
      __syncthreads();
      int any = 0;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) any |= cs.state[cc];
      if (!any) break;
    }
    for (int col = warp; col < 8; col += nwarps) {
      const int cg = c0 + col;
      if (col >= nact || cg >= d.nic) continue;
      const double* X = ds.xs + ((size_t)blockIdx.x * 8 + col) * N2;
      double* out = a.terminal_out + (size_t)N2 * ((size_t)cg + (size_t)d.nic * b);
      for (int r = lane; r < N2; r += 32) out[r] = X[r];
      if (a.iters_term && lane == 0) a.iters_term[(size_t)cg + (size_t)d.nic * b] = cs.it[col];
    }
    __syncthreads();  // cs and the tiles are rewritten by the next column
   }
  }
}

}  // namespace qgd

namespace {
template <int M>
void launch_dense_t(qgd_handle* h, const double* comb, double* uv, int ncols, int adjoint) {
  const int N = h->N, S = 2 * N + 4;
  const size_t smem = (size_t)(M + 1) * 8 * S * 8;
  if (adjoint) {
    CUDA_CHECK(cudaFuncSetAttribute(qgd::k_derivs_dense_adj<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qgd::k_derivs_dense_adj<M><<<(ncols + 7) / 8, 32 * (N / 32), smem, h->stream>>>(comb, N, uv, ncols);
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(qgd::k_derivs_dense<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qgd::k_derivs_dense<M><<<(ncols + 7) / 8, 32 * (N / 32), smem, h->stream>>>(comb, N, uv, ncols);
  }
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}
}  // namespace

// Dense DMMA path of compute_derivatives! applicable?  (dense-classified problem, N a multiple of 32 up to 256, the 8
// Taylor-column tiles fit the shared memory)
bool dense_derivs_applicable(const qgd_handle* h, int m) {
  if (h->opt[QGD_OPT_DISABLE_DENSE_DMMA]) return false;
  if (h->fast_ok || h->dense_ops.empty() || h->N % 32 != 0 || h->N > 256 || m < 1 || m > 6) return false;
  return (size_t)(m + 1) * 8 * (2 * h->N + 4) * 8 <= h->prop.sharedMemPerBlockOptin;
}

// d_uv [2N][1+m][ncols] and d_cv [2][m+1][Nc] on the device; adjoint != 0: the columns Lambda_j = W_j^T x.
void launch_derivs_dense(qgd_handle* h, int m, double* d_uv, int ncols, const double* d_cv, int adjoint) {
  const int N = h->N;
  const size_t nn = (size_t)N * N;
  if (h->d_dense.cap == 0) {  // dense row-major copies of the operators, uploaded on first use
    h->d_dense.reserve(h->dense_ops.size() * 8);
    CUDA_CHECK(cudaMemcpyAsync(h->d_dense.p, h->dense_ops.data(), h->dense_ops.size() * 8, cudaMemcpyHostToDevice, h->stream));
  }
  h->d_comb.reserve((size_t)m * 2 * nn * 8);
  const size_t total = (size_t)m * 2 * nn;
  qgd::k_dense_combine<<<(unsigned)std::min<size_t>((total + 255) / 256, 4096), 256, 0, h->stream>>>(
      h->d_dense.as<double>(), N, h->Nc, m, d_cv, h->d_comb.as<double>());
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
  const double* comb = h->d_comb.as<double>();
  switch (m) {
    case 1: launch_dense_t<1>(h, comb, d_uv, ncols, adjoint); break;
    case 2: launch_dense_t<2>(h, comb, d_uv, ncols, adjoint); break;
    case 3: launch_dense_t<3>(h, comb, d_uv, ncols, adjoint); break;
    case 4: launch_dense_t<4>(h, comb, d_uv, ncols, adjoint); break;
    case 5: launch_dense_t<5>(h, comb, d_uv, ncols, adjoint); break;
    default: launch_dense_t<6>(h, comb, d_uv, ncols, adjoint); break;
  }
}

namespace {
template <int M>
void launch_forward_dense_t(qgd_handle* h, const QgdDevProb& d, qgd::SweepArgs a, qgd::DenseSweepArgs ds, int grid, size_t smem) {
  CUDA_CHECK(cudaFuncSetAttribute(qgd::k_forward_dense<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qgd::k_forward_dense<M><<<grid, d.N, smem, h->stream>>>(d, a, ds);
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}
template <int M>
void launch_terminal_dense_t(qgd_handle* h, const QgdDevProb& d, qgd::SweepArgs a, qgd::DenseSweepArgs ds, int grid, size_t smem, int seq) {
  CUDA_CHECK(cudaFuncSetAttribute(qgd::k_terminal_dense<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qgd::k_terminal_dense<M><<<grid, d.N, smem, h->stream>>>(d, a, ds, seq);
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}
template <int M>
void launch_backward_dense_t(qgd_handle* h, const QgdDevProb& d, qgd::SweepArgs a, qgd::DenseSweepArgs ds, int grid, size_t smem) {
  CUDA_CHECK(cudaFuncSetAttribute(qgd::k_backward_dense<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qgd::k_backward_dense<M><<<grid, d.N, smem, h->stream>>>(d, a, ds, h->d_dense.as<double>(), h->d_ctrls.as<QgdDevControl>());
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}

// Shared set-up of the two dense sweeps: applicability, the per-level combined operators of the whole time grid,
// Krylov / state workspaces for `grid` CTAs of 8 columns.  extra_smem: bytes beyond the Taylor tiles.
bool prepare_dense_sweep(qgd_handle* h, const QgdDevProb& d, qgd::SweepArgs& a, qgd::DenseSweepArgs& ds, int& grid, size_t& smem,
                         size_t extra_smem, int ncols) {
  const int m = d.m, N = h->N, N2 = h->N2;
  if (h->opt[QGD_OPT_DISABLE_DENSE_SWEEP]) return false;
  if (!dense_derivs_applicable(h, m) || h->Nc < 1) return false;
  smem = (size_t)(m + 1) * 8 * (N2 + 4) * 8 + extra_smem;
  if (smem + 1024 > h->prop.sharedMemPerBlockOptin) return false;
  if ((size_t)a.B * (h->nsteps + 1) > 65535) return false;  // grid.y of the operator combination
  const size_t nn = (size_t)N * N, levels = (size_t)a.B * (h->nsteps + 1);
  const size_t comb_bytes = levels * m * 2 * nn * 8;
  size_t free_b = 0, total_b = 0;
  CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
  if (comb_bytes > h->d_comb.cap && comb_bytes > (free_b / 10) * 7) return false;
  if (h->d_dense.cap == 0) {
    h->d_dense.reserve(h->dense_ops.size() * 8);
    CUDA_CHECK(cudaMemcpyAsync(h->d_dense.p, h->dense_ops.data(), h->dense_ops.size() * 8, cudaMemcpyHostToDevice, h->stream));
  }
  h->d_comb.reserve(comb_bytes);
  {
    const size_t total = (size_t)m * 2 * nn;
    dim3 cgrid((unsigned)std::min<size_t>((total + 255) / 256, 1024), (unsigned)levels);
    qgd::k_dense_combine_levels<<<cgrid, 256, 0, h->stream>>>(h->d_dense.as<double>(), N, h->Nc, m, a.cvals, h->d_comb.as<double>());
    CUDA_CHECK(cudaGetLastError());
    h->stats.kernel_launches++;
  }
  const int groups = (ncols + 7) / 8;
  grid = std::max(1, std::min(a.B * groups, h->prop.multiProcessorCount));
  const size_t slots = (size_t)grid * 8;
  a.v_stride = (size_t)(N2 + 1) * N2;
  a.h_stride = (size_t)N2 * (N2 + 3) / 2 + 2;
  h->d_V.reserve(slots * a.v_stride * 8);
  h->d_H.reserve(slots * a.h_stride * 8);
  a.Vws = h->d_V.as<double>();
  a.Hws = h->d_H.as<double>();
  h->d_dense_ws.reserve(slots * (3 * (size_t)N2 + 3 * ((size_t)N2 + 2)) * 8);
  ds.comb = h->d_comb.as<double>();
  ds.xs = h->d_dense_ws.as<double>();
  ds.bs = ds.xs + slots * N2;
  ds.xp = ds.bs + slots * N2;
  ds.aux = ds.xp + slots * N2;
  return true;
}
}  // namespace

// Forward sweep of a dense problem on the tensor-core contraction (k_forward_dense).  false: not applicable (sparse or
// register-operator problem, level count not a multiple of 32, the per-level operators of the whole
// time grid do not fit the device memory) -- the caller then uses the generic kernels.
bool try_forward_dense(qgd_handle* h, const QgdDevProb& d, const qgd::SweepArgs& a_in) {
  qgd::SweepArgs a = a_in;
  qgd::DenseSweepArgs ds{};
  int grid = 0;
  size_t smem = 0;
  if (!prepare_dense_sweep(h, d, a, ds, grid, smem, 0, h->ncol)) return false;
  switch (d.m) {
    case 1: launch_forward_dense_t<1>(h, d, a, ds, grid, smem); break;
    case 2: launch_forward_dense_t<2>(h, d, a, ds, grid, smem); break;
    case 3: launch_forward_dense_t<3>(h, d, a, ds, grid, smem); break;
    case 4: launch_forward_dense_t<4>(h, d, a, ds, grid, smem); break;
    case 5: launch_forward_dense_t<5>(h, d, a, ds, grid, smem); break;
    default: launch_forward_dense_t<6>(h, d, a, ds, grid, smem); break;
  }
  h->stats.fast_path_launches++;
  return true;
}

// Adjoint sweep + gradient accumulation of a dense problem (k_backward_dense); needs the full history (saveEveryNsteps 1)
// on the device.  false: not applicable (as above, or the gradient slots do not fit beside the Taylor tiles).
bool try_backward_dense(qgd_handle* h, const QgdDevProb& d, const qgd::SweepArgs& a_in) {
  qgd::SweepArgs a = a_in;
  qgd::DenseSweepArgs ds{};
  int grid = 0;
  size_t smem = 0;
  const size_t G = (size_t)2 * d.m * h->Nc * 8;
  if (!prepare_dense_sweep(h, d, a, ds, grid, smem, (size_t)(h->N / 32 + 1) * G * 8, h->ncol)) return false;
  switch (d.m) {
    case 1: launch_backward_dense_t<1>(h, d, a, ds, grid, smem); break;
    case 2: launch_backward_dense_t<2>(h, d, a, ds, grid, smem); break;
    case 3: launch_backward_dense_t<3>(h, d, a, ds, grid, smem); break;
    case 4: launch_backward_dense_t<4>(h, d, a, ds, grid, smem); break;
    case 5: launch_backward_dense_t<5>(h, d, a, ds, grid, smem); break;
    default: launch_backward_dense_t<6>(h, d, a, ds, grid, smem); break;
  }
  h->stats.fast_path_launches++;
  return true;
}

// Infidelity + terminal condition of a dense problem (k_terminal_dense): all nic columns, 8 per CTA.
bool try_terminal_dense(qgd_handle* h, const QgdDevProb& d, const qgd::SweepArgs& a_in) {
  if (h->opt[QGD_OPT_DENSE_TERMINAL] == 2) return false;  // one warp per control vector (k_terminal)
  const int seq = h->opt[QGD_OPT_DENSE_TERMINAL] == 1 ? 0 : 1;
  qgd::SweepArgs a = a_in;
  qgd::DenseSweepArgs ds{};
  int grid = 0;
  size_t smem = 0;
  if (!prepare_dense_sweep(h, d, a, ds, grid, smem, 0, seq ? 1 : h->nic)) return false;
  switch (d.m) {
    case 1: launch_terminal_dense_t<1>(h, d, a, ds, grid, smem, seq); break;
    case 2: launch_terminal_dense_t<2>(h, d, a, ds, grid, smem, seq); break;
    case 3: launch_terminal_dense_t<3>(h, d, a, ds, grid, smem, seq); break;
    case 4: launch_terminal_dense_t<4>(h, d, a, ds, grid, smem, seq); break;
    case 5: launch_terminal_dense_t<5>(h, d, a, ds, grid, smem, seq); break;
    default: launch_terminal_dense_t<6>(h, d, a, ds, grid, smem, seq); break;
  }
  return true;
}


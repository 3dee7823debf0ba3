// qgd_fast.cuh -- register-resident-operator sweep kernels (sm_100a): the hot path for the problems the
// reference is built around (DispersiveProblem-style: diagonal drift, control operators with at most two
// entries per row such as a +- a^dagger, diagonal guard projector, Identity or DiagonalHamiltonian
// preconditioner, N_tot_levels <= 256).  Everything else runs on the generic kernels of qgd_kernels.cuh.
//
// One warp = one initial-condition column of one control vector, marching all time steps on-device (N <= 64; for
// 64 < N <= 256 a row-split group of two or four warps, FastCtx RS below).
//   * Hamiltonian blocks live in REGISTERS (values + gather columns of the <= 2 entries per row and
//     operator), the state is exchanged between lanes through a 16-byte (u,v)-interleaved shared-memory
//     buffer (one LDS.128 per gathered entry).
//   * Taylor recursion in scatter form: the sparse products K_k w_i, S_k w_i are formed ONCE per Taylor
//     column (m sparse sweeps per operator evaluation instead of m(m+1)/2) and immediately scattered
//     into all later columns with the control Taylor coefficients (compute_derivatives!, reference
//     src/hermite.jl:56-101; apply_hamiltonian!, :556-588).
//   * GMRES (IterativeSolvers.jl semantics): Krylov basis in shared memory (first KS
//     vectors) with an L2-resident tail that is prefetched one vector ahead; modified Gram-Schmidt dot
//     products reduced across the warp either by the 5-stage shuffle butterfly or by two FP64 tensor-core
//     DMMA.8x8x4 with a ones operand (measured 66 vs 175 cycles, profiles/r01_microbench.txt); Givens
//     rotations applied progressively to each new Hessenberg column (same rotations in the same order as
//     solve_least_squares!, so same numbers), triangular solve from the packed R factor.
#pragma once
#include "qgd_kernels.cuh"

namespace qgd {

#ifndef QGD_FAST_RED
#define QGD_FAST_RED 1  // 0: shuffle butterfly, 1: DMMA ones-matrix all-reduce
#endif
// Gram-Schmidt block width B.  1: strict modified Gram-Schmidt, one projection at a time, Givens rotations applied
// progressively (gmres_fast_strict).  4 or 8: the Krylov basis is swept in blocks of B vectors; the B coefficients of
// a block are taken from the same vector (classical inside a block, modified across blocks), so their warp
// reductions fuse into one transposing reduction, and the Givens QR of the Hessenberg matrix is done once per solve
// (as IterativeSolvers' solve_least_squares! does).  The two orthogonalisations differ by h_i <v_i, v_j> = O(eps kappa)
// inside a block -- the size of the rounding differences between two dot-product orders.
#ifndef QGD_MGS_BLOCK
#define QGD_MGS_BLOCK 8
#endif
// Per-kernel code-shape switches of the blocked orthogonalisation, chosen by measurement:
// bit 0 = the three basis tiers share one copy of the block arithmetic (smaller code: the adjoint kernel's hot
// loop then fits the 32 KB instruction cache), bit 1 = block reduction through shared memory instead of shuffles.
// Two-pass super-blocks (QGD_GS_SUPER = S, a multiple of 8; 0 = the one-pass blocks above, the default).  The S
// coefficients of a super-block come from the same vector; pass 1 streams the basis once for the dot products, pass 2
// streams it again for the updates: one dependent reduction chain per S vectors instead of one per 8.
// tools/gs_block_experiment.py measures on the CPU oracle that widths from 8 to the whole basis leave every GMRES
// iteration count of the C2 problem unchanged (tolerances 1e-12 .. 1e-15), so the numerics would allow it -- but
// MEASURED ON B200 (round 2, profiles/r02_gs_superblock.txt) it is SLOWER at every width: 324 vs 388 evals/s at batch
// 592 and 411 vs 337 ms for a single evaluation (S = 32, 64, 128 alike).  A warp issues in order, so the load ->
// dots -> transposition chain of each group of 8 is exposed twice (once per pass) unless the groups are software-
// pipelined, and that needs the partials of a whole super-block (64 registers) or one transposition buffer per group
// (8 KB of shared memory per warp), neither of which exists beside the register-resident operators and the Krylov
// tiers.  Kept compiled out as the record of the experiment.
#ifndef QGD_GS_SUPER
#define QGD_GS_SUPER 0
#endif
#ifndef QGD_FWD_VARIANT
#define QGD_FWD_VARIANT 2
#endif
#ifndef QGD_BWD_VARIANT
#define QGD_BWD_VARIANT 3
#endif
// Lean least squares (qr_solve_lean) and compact per-warp shared memory: the rotated right-hand side `g` lives in the
// gather buffer (idle during the least-squares solve), the subdiagonal of H is read from the packed matrix, the
// adjoint sweep's gradient accumulator lives in L2 -- which frees 3.7 KB per warp and lets 24 instead of 16 Krylov
// vectors stay in shared memory.
#ifndef QGD_QR_LEAN
#define QGD_QR_LEAN 1
#endif
#define QGD_COMPACT_SMEM (QGD_QR_LEAN && QGD_MGS_BLOCK > 1)
#ifndef QGD_BWD_MERGE_SIDES
#define QGD_BWD_MERGE_SIDES 0  // gradient sweeps of an adjoint step: 0 two inlined copies, 1 one rolled loop, 2 one real function
#endif

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
      : "=d"(d0), "=d"(d1)
      : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}

// Sum of one double per lane, result in every lane.
__device__ __forceinline__ double warp_allsum(double p) {
#if QGD_FAST_RED == 1
  // B operand (4x8, lane l <-> B[l%4][l/4]) = p, A = ones: D[i][j] = sum of the 4 lanes of group j; lane l
  // receives groups 2(l%4), 2(l%4)+1.  Their sum as A operand (8x4, lane l <-> A[l/4][l%4]) times ones
  // gives the total in every lane.
  double c0, c1, t0, t1;
  dmma884(c0, c1, 1.0, p);
  dmma884(t0, t1, c0 + c1, 1.0);
  return t0;
#else
  return warp_sum(p);
#endif
}

// 1: one FMA chain per local dot product (2 EL instructions) instead of two chains and an add (2 EL + 1).  MEASURED ON B200 (round 2,
// A/B through tools/build_variant.sh on the final library): adjoint sweep 812 -> 794 ms at batch 592, forward sweep unchanged
// (701 ms), 389.7 -> 394.5 evals/s; every GPU parity test unchanged (same GMRES iteration counts).  On.
#ifndef QGD_DOT_CHAIN
#define QGD_DOT_CHAIN 1
#endif
template <int EL>
__device__ __forceinline__ double vdot_local(const Vec<EL>& a, const Vec<EL>& b) {
#if QGD_DOT_CHAIN
  double s = a.u[0] * b.u[0];
  s = fma(a.v[0], b.v[0], s);
#pragma unroll
  for (int e = 1; e < EL; ++e) { s = fma(a.u[e], b.u[e], s); s = fma(a.v[e], b.v[e], s); }
  return s;
#else
  double s0 = a.u[0] * b.u[0], s1 = a.v[0] * b.v[0];
#pragma unroll
  for (int e = 1; e < EL; ++e) { s0 = fma(a.u[e], b.u[e], s0); s1 = fma(a.v[e], b.v[e], s1); }
  return s0 + s1;
#endif
}

// ---- per-lane operator registers ---------------------------------------------------------------------
template <int EL, int NC>
struct RegOps {
  double kd[EL];            // drift: diagonal of K_s (an antisymmetric S_s has a zero diagonal)
  int col[EL][NC][2];       // gather column of entry s of row (lane + 32 e) of control operator k
  double vk[EL][NC][2];     // K_c value
  double vs[EL][NC][2];     // S_c value
  double pr_ratio[EL], pr_up[EL], pr_rden[EL], pr_rdg[EL];  // DiagonalHamiltonianPreconditioner (reciprocal pivots)
  double wu[EL], wv[EL];    // guard projector diagonal (u rows, v rows)
};

template <int EL, int NC>
__device__ __forceinline__ void load_regops(RegOps<EL, NC>& R, const QgdDevProb& d, int lane, int dir) {
  const QgdOpLayout& L = d.lay;
  const int N = d.N;
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    const int r = lane + 32 * e;
    const bool ok = r < N;
    R.kd[e] = (ok && L.L[0] > 0) ? reinterpret_cast<const double*>(d.blob + L.off_vk[0])[r] : 0.0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int* col = reinterpret_cast<const int*>(d.blob + L.off_col[k + 1]);
      const double* vk = reinterpret_cast<const double*>(d.blob + L.off_vk[k + 1]);
      const double* vs = reinterpret_cast<const double*>(d.blob + L.off_vs[k + 1]);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const bool have = ok && s < L.L[k + 1];
        R.col[e][k][s] = have ? col[s * N + r] : 0;
        R.vk[e][k][s] = have ? vk[s * N + r] : 0.0;
        R.vs[e][k][s] = have ? vs[s * N + r] : 0.0;
      }
    }
    if (d.precond == QGD_PRECOND_DIAGONAL && dir >= 0) {
      const double* pd = reinterpret_cast<const double*>(d.blob + L.off_pre[dir]);
      R.pr_rdg[e] = ok ? 1.0 / pd[r] : 1.0;
      R.pr_up[e] = ok ? pd[d.N2 + r] : 0.0;
      R.pr_ratio[e] = ok ? pd[d.N2 + N + r] : 0.0;
      R.pr_rden[e] = ok ? 1.0 / pd[d.N2 + 2 * N + r] : 1.0;
    } else {
      R.pr_rdg[e] = 1.0; R.pr_up[e] = 0.0; R.pr_ratio[e] = 0.0; R.pr_rden[e] = 1.0;
    }
    if (L.LW > 0 && ok) {
      const double* wv = reinterpret_cast<const double*>(d.blob + L.off_wval);
      R.wu[e] = wv[r]; R.wv[e] = wv[N + r];
    } else {
      R.wu[e] = 0.0; R.wv[e] = 0.0;
    }
  }
}

// ---- per-warp context ----------------------------------------------------------------------------------
// RS (row split, 1 / 2 / 4): the level rows of a column are dealt over RS warps of a CTA -- warp `slice` owns the rows
// [32 EL slice, 32 EL (slice + 1)) with the SAME per-lane registers as a one-warp column (operators, state, its slice of
// every Krylov vector in its own TMEM quarter / shared memory / L2 tail).  The warps of a group exchange the state for the
// operator gathers through ONE shared gather buffer (gx) and complete every row-space reduction through gred, each
// behind a named barrier of the group; all scalar work of GMRES (Hessenberg column, residual recurrence, least squares)
// is replicated per warp from bitwise identical sums, so the warps of a group take every branch together.
template <int EL, int RS = 1>
struct FastCtx {
  int lane, N, N2;
  int vl;         // row index of this lane's first element: lane + 32 EL slice  (row-space loads, stores, gathers)
  int slice, bar; // RS > 1: position in the group, named barrier of the group
  double2* gx;    // RS > 1: smem [2][32 EL RS] gather buffer of the group
  double* gred;   // RS > 1: smem [2][RS][8] partial sums of the group (double buffered: one barrier per reduction)
  unsigned* gtick;  // RS > 1: smem ticket word of the group
  double2* stage;   // RS > 1: smem [8][32 EL] staging block of the L2 tier (the last 8 slots of the shared-memory tier), or null
  mutable unsigned par;  // RS > 1: parity of the next reduction
  int KT, KS;     // Krylov vectors [0,KT) live in TMEM, [KT,KT+KS) in shared memory, the rest in L2
  uint32_t tm;    // TMEM address of this warp's region (its 32 lanes, its column range)
  // xs: (u,v) gather buffers of the operator application (double buffered, 2 x 32 EL double2); the same storage is
  // the 8 x 32 transposition buffer of block_allsum and the cp.async ring of qr_solve_fast (>= 256 doubles)
  static constexpr int kRS = RS;
  static constexpr int kRing0 = 2 * 2 * 32 * EL < 256 ? 256 : 2 * 2 * 32 * EL;
  static constexpr int kRingDoubles = RS == 1 ? kRing0 : (kRing0 > 64 * EL * RS + 2 ? kRing0 : 64 * EL * RS + 2);  // RS > 1: g [2N + 2]
  double2* xs;    // smem [kRingDoubles / 2]
  double2* cv;    // smem [M+1][NC]    (p_k^(d)/d!, q_k^(d)/d!) of the current time level
  double* nullv;  // smem [N2+2]       left null vector of the residual recurrence
  double2* rot;   // smem [N2+2+8]     rot[0] = (1,0); rot[i+1] = Givens (cs, sn) of rotation i   (strict path)
  double* hcol;   // smem [N2+10]      same storage as rot (blocked path): current Hessenberg column
  double* sub;    // smem [N2+2]       same storage as rot (blocked path): subdiagonal H[j+1][j]
  double* g;      // smem [N2+2]       rotated right-hand side, then the least-squares solution
  double2* Vs;    // smem [KS][32*EL]  Krylov basis, shared-memory tier
  double2* Vg;    // global            Krylov basis, tail (vector i >= KT+KS at (i-KT-KS)*32*EL)
  double* Rg;     // global            packed upper-triangular R: column j at j(j+1)/2
  double* team;   // smem: shared state of the latency team (TeamView), null in the one-warp-per-column kernels
};
template <int EL, int M, int NC, bool STRICT = false, int RS = 1>
__host__ __device__ constexpr int fast_fixed_doubles(int N2) {
#if QGD_COMPACT_SMEM
  // strict kernels (progressive Givens): xs (+ gKs, gSs aliased) + cv + nullv + rot + g
  if (STRICT) return FastCtx<EL, RS>::kRingDoubles + 2 * (M + 1) * NC + (N2 + 2) + 2 * (N2 + 2 + 8) + (N2 + 2);
  // xs (+ g, gKs, gSs aliased) + cv + nullv + hcol
  return FastCtx<EL, RS>::kRingDoubles + 2 * (M + 1) * NC + (N2 + 2) + (N2 + 2 + 8);
#else
  // xs + cv + nullv + rot + g
  return FastCtx<EL, RS>::kRingDoubles + 2 * (M + 1) * NC + (N2 + 2) + 2 * (N2 + 2 + 8) + (N2 + 2);
#endif
}

template <int EL>
__device__ __forceinline__ void xs_store(double2* xs, const Vec<EL>& a, int lane) {
#pragma unroll
  for (int e = 0; e < EL; ++e) xs[lane + 32 * e] = make_double2(a.u[e], a.v[e]);
}
template <int EL>
__device__ __forceinline__ void v2_load(Vec<EL>& a, const double2* p, int lane) {
#pragma unroll
  for (int e = 0; e < EL; ++e) { const double2 t = p[lane + 32 * e]; a.u[e] = t.x; a.v[e] = t.y; }
}
template <int EL>
__device__ __forceinline__ void v2_load_cg(Vec<EL>& a, const double2* p, int lane) {
#pragma unroll
  for (int e = 0; e < EL; ++e) { const double2 t = __ldcg(p + lane + 32 * e); a.u[e] = t.x; a.v[e] = t.y; }
}
template <int EL>
__device__ __forceinline__ void v2_store(double2* p, const Vec<EL>& a, int lane) {
#pragma unroll
  for (int e = 0; e < EL; ++e) p[lane + 32 * e] = make_double2(a.u[e], a.v[e]);
}

// ---- row-split groups (RS > 1) ---------------------------------------------------------------------------
template <int EL, int RS>
__device__ __forceinline__ void gsync(const FastCtx<EL, RS>& c) {
  if constexpr (RS == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(c.bar), "n"(32 * RS) : "memory");
}
// gather buffer `which` (0 / 1) of the operator application: per warp, or the group's
template <int EL, int RS>
__device__ __forceinline__ double2* gather_buf(const FastCtx<EL, RS>& c, int which) {
  if constexpr (RS == 1) return c.xs + which * 32 * EL;
  else return c.gx + which * 32 * EL * RS;
}
// Sum of one double per lane over the ROWS of the column: over the warp, and for RS > 1 over the warps of the group -- every
// warp adds the RS warp totals in slice order, so all of them hold the same bits.
template <int EL, int RS>
__device__ __forceinline__ double row_allsum(const FastCtx<EL, RS>& c, double p) {
  const double t = warp_allsum(p);
  if constexpr (RS == 1) return t;
  else {
    double* buf = c.gred + (c.par & 1u) * (RS * 8);
    c.par ^= 1u;
    if (c.lane == 0) buf[c.slice * 8] = t;
    gsync(c);
    double r = buf[0];
#pragma unroll
    for (int sl = 1; sl < RS; ++sl) r += buf[sl * 8];
    return r;
  }
}
// The same for the 8 coefficients of a Gram-Schmidt block: warp totals in slot[0..8) (shared memory, this warp's Hessenberg
// column) on entry, group totals there and in p on return.
template <int EL, int RS>
__device__ __forceinline__ void row_combine8(const FastCtx<EL, RS>& c, double (&p)[8], double* slot) {
  if constexpr (RS > 1) {
    double* buf = c.gred + (c.par & 1u) * (RS * 8);
    c.par ^= 1u;
    if (c.lane < 8) buf[c.slice * 8 + c.lane] = slot[c.lane];
    gsync(c);
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      double2 t = reinterpret_cast<const double2*>(buf)[q >> 1];
#pragma unroll
      for (int sl = 1; sl < RS; ++sl) {
        const double2 u = reinterpret_cast<const double2*>(buf + sl * 8)[q >> 1];
        t.x += u.x; t.y += u.y;
      }
      p[q] = t.x; p[q + 1] = t.y;
      if (c.lane == (q >> 1)) reinterpret_cast<double2*>(slot)[q >> 1] = t;
    }
    __syncwarp();
  }
}

// block_allsum8_smem for a row-split group in one go: the warp's totals leave the tensor-core reduction straight into the
// group's slot (no round trip through the warp's own Hessenberg column first), one barrier, every warp sums the RS slices.
#ifndef QGD_RS_FUSED_RED
#define QGD_RS_FUSED_RED 1
#endif
template <int EL, int RS>
__device__ __forceinline__ void block_allsum8_group(const FastCtx<EL, RS>& c, double (&p)[8], double* T, double* slot) {
  const int lane = c.lane;
#pragma unroll
  for (int q = 0; q < 8; ++q) T[32 * q + (lane ^ (4 * (q & 3)))] = p[q];
  __syncwarp();
  const int qv = lane >> 2, s = lane & 3;
  const double* row = T + 32 * qv + s;
  const int sw = qv & 3;
  double a[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) a[t] = row[4 * (t ^ sw)];
  const double r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  double c0, c1;
  dmma884(c0, c1, 1.0, r);
  double* buf = c.gred + (c.par & 1u) * (RS * 8);
  c.par ^= 1u;
  if (lane < 4) reinterpret_cast<double2*>(buf + c.slice * 8)[lane] = make_double2(c0, c1);  // totals of values 2 lane, 2 lane + 1
  gsync(c);
#pragma unroll
  for (int q = 0; q < 8; q += 2) {
    double2 t = reinterpret_cast<const double2*>(buf)[q >> 1];
#pragma unroll
    for (int sl = 1; sl < RS; ++sl) {
      const double2 u = reinterpret_cast<const double2*>(buf + sl * 8)[q >> 1];
      t.x += u.x; t.y += u.y;
    }
    p[q] = t.x; p[q + 1] = t.y;
    if (lane == (q >> 1)) reinterpret_cast<double2*>(slot)[q >> 1] = t;
  }
  __syncwarp();
}

// z-sums of one vector (in the gather buffer): K_k x and S_k x restricted to this lane's rows.
template <int EL, int NC>
struct ZS { double Ku[EL][NC], Kv[EL][NC], Su[EL][NC], Sv[EL][NC]; };

template <int EL, int NC>
__device__ __forceinline__ void zsums(const RegOps<EL, NC>& R, const double2* xs, ZS<EL, NC>& z) {
#pragma unroll
  for (int e = 0; e < EL; ++e)
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const double2 g0 = xs[R.col[e][k][0]], g1 = xs[R.col[e][k][1]];
      z.Ku[e][k] = fma(R.vk[e][k][1], g1.x, R.vk[e][k][0] * g0.x);
      z.Kv[e][k] = fma(R.vk[e][k][1], g1.y, R.vk[e][k][0] * g0.y);
      z.Su[e][k] = fma(R.vs[e][k][1], g1.x, R.vs[e][k][0] * g0.x);
      z.Sv[e][k] = fma(R.vs[e][k][1], g1.y, R.vs[e][k][0] * g0.y);
    }
}

// out = sum_j alpha_j w_j, w_0 = x, w_{j+1} = (1/(j+1)) sum_{i<=j} A_{j-i} w_i  (scatter form).
// STEP: also guess = sum_j a_tay_j w_j and, if hist != nullptr, hist[:, j] = w_j.
// FORCE: the recursion carries a forcing, w_{j+1} = (1/(j+1)) (sum_{i<=j} A_{j-i} w_i + f_j)
// (compute_derivatives!(...; forcing_matrix), reference src/hermite.jl:91-95).  f_j is either read from an explicit array
// (eval_forward!(...; forcing), src/forward_evolution.jl:118-129) or formed on the fly for eval_grad_forced
// (src/eval_grad_forced.jl:82-131) from the Taylor columns w_i of the unforced solution and the control basis table,
//   f_j = sum_{i<=j} d/dtheta [p^(j-i)/(j-i)!] (K w_i)-part + d/dtheta [q^(j-i)/(j-i)!] (S w_i)-part
// with K, S the control operator the parameter theta belongs to: the same scatter as the recursion itself, with the
// basis-table entries of theta in the place of the control values and the history columns in the place of the w_i.
struct FastForcing {
  const double* arr;  // [2N][M] explicit forcing of this time level, or null
  const double* hb;   // [2N][1+M] unforced Taylor columns of this time level (global), or null
  const double* tp;   // tp[dd * P] = d/dtheta p^(dd)/dd! at this time level; tq likewise
  const double* tq;
  int P, kop;         // kop: index (0-based) of the control operator of theta
};

template <int EL, int M, int NC, bool STEP, bool FORCE = false, int RS>
__device__ __forceinline__ void fwd_fast(const FastCtx<EL, RS>& c, const RegOps<EL, NC>& R, const Vec<EL>& x, const double* alpha,
                                         Vec<EL>& out, const double* a_tay, Vec<EL>* guess, double* hist,
                                         const FastForcing* F = nullptr) {
  const int lane = c.vl, N = c.N, N2 = c.N2;  // row space: lane + 32 EL slice
  Vec<EL> acc[M + 1];
#pragma unroll
  for (int j = 1; j <= M; ++j) vzero(acc[j]);
  if constexpr (FORCE) {
    if (F->arr) {
#pragma unroll
      for (int j = 0; j < M; ++j) vload_cg(acc[j + 1], F->arr + (size_t)j * N2, N, lane);
    }
    if (F->hb) {
      gsync(c);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        Vec<EL> wi;
        vload_cg(wi, F->hb + (size_t)i * N2, N, lane);
        double2* xb = gather_buf(c, i & 1);
        xs_store<EL>(xb, wi, lane);
        gsync(c);
        ZS<EL, NC> z;
        zsums<EL, NC>(R, xb, z);
        double Ku[EL], Kv[EL], Su[EL], Sv[EL];
#pragma unroll
        for (int e = 0; e < EL; ++e) {
          Ku[e] = z.Ku[e][0]; Kv[e] = z.Kv[e][0]; Su[e] = z.Su[e][0]; Sv[e] = z.Sv[e][0];
#pragma unroll
          for (int k = 1; k < NC; ++k)
            if (k == F->kop) { Ku[e] = z.Ku[e][k]; Kv[e] = z.Kv[e][k]; Su[e] = z.Su[e][k]; Sv[e] = z.Sv[e][k]; }
        }
#pragma unroll
        for (int j = i; j < M; ++j) {
          const double pv = F->tp[(size_t)(j - i) * F->P], qv = F->tq[(size_t)(j - i) * F->P];
#pragma unroll
          for (int e = 0; e < EL; ++e) {
            acc[j + 1].u[e] = fma(pv, Kv[e], fma(qv, Su[e], acc[j + 1].u[e]));
            acc[j + 1].v[e] = fma(-pv, Ku[e], fma(qv, Sv[e], acc[j + 1].v[e]));
          }
        }
      }
      gsync(c);
    }
  }
  Vec<EL> w = x;
  out = x;
  vscale(out, alpha[0]);
  if (STEP) {
    *guess = x;
    if (hist) vstore_cs(x, hist, N, lane);
  }
  gsync(c);
#pragma unroll
  for (int i = 0; i < M; ++i) {
    if (i > 0) {
      const double inv = 1.0 / (double)i;
      w = acc[i];
      vscale(w, inv);
      vaxpy(out, alpha[i], w);
      if (STEP) {
        vaxpy(*guess, a_tay[i], w);
        if (hist) vstore_cs(w, hist + (size_t)i * N2, N, lane);
      }
    }
    double2* xb = gather_buf(c, i & 1);
    xs_store<EL>(xb, w, lane);
    gsync(c);
    ZS<EL, NC> z;
    zsums<EL, NC>(R, xb, z);
#pragma unroll
    for (int j = i; j < M; ++j) {
      const int dd = j - i;
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        double au = acc[j + 1].u[e], av = acc[j + 1].v[e];
        if (dd == 0) { au = fma(R.kd[e], w.v[e], au); av = fma(-R.kd[e], w.u[e], av); }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const double2 cc = c.cv[dd * NC + k];
          au = fma(cc.y, z.Su[e][k], fma(cc.x, z.Kv[e][k], au));
          av = fma(cc.y, z.Sv[e][k], fma(-cc.x, z.Ku[e][k], av));
        }
        acc[j + 1].u[e] = au; acc[j + 1].v[e] = av;
      }
    }
  }
  {
    const double inv = 1.0 / (double)M;
    w = acc[M];
    vscale(w, inv);
    vaxpy(out, alpha[M], w);
    if (STEP) {
      vaxpy(*guess, a_tay[M], w);
      if (hist) vstore_cs(w, hist + (size_t)M * N2, N, lane);
    }
  }
}

// Reverse sweep: what_j = alpha_j x; for j = M-1..0: what_{j-d} -= (1/(j+1)) A_d what_{j+1}, d = 0..j.
// out = what_0 = (sum_j alpha_j W_j)^T x.  GRAD: accumulate per lane
//   gK[r][k] += IPK_k(w_i, what_{j+1})/(j+1), gS[r][k] += IPS_k(w_i, what_{j+1})/(j+1), r = j - i,
// with w_i the forward Taylor columns of the same time level (hist, global).  SURVEY A.6.
// LAST: this is the last use of the history level (evict-first load); otherwise the level is read once more by the
// next adjoint step and is loaded with the default L2 policy so that it is still resident then.
template <int EL, int M, int NC, bool GRAD, int RS>
__device__ __forceinline__ void adj_fast(const FastCtx<EL, RS>& c, const RegOps<EL, NC>& R, const Vec<EL>& x, const double* alpha,
                                         Vec<EL>& out, const double* hist, double (&gK)[M][NC], double (&gS)[M][NC],
                                         bool last_use = true) {
  const int lane = c.vl, N = c.N, N2 = c.N2;  // row space: lane + 32 EL slice
  Vec<EL> what[M + 1];
#pragma unroll
  for (int j = 0; j <= M; ++j) { what[j] = x; vscale(what[j], alpha[j]); }
  Vec<EL> wh[M];
  if (GRAD) {
    if (last_use) {
#pragma unroll
      for (int i = 0; i < M; ++i) vload_cs(wh[i], hist + (size_t)i * N2, N, lane);
    } else {
#pragma unroll
      for (int i = 0; i < M; ++i) vload_cg(wh[i], hist + (size_t)i * N2, N, lane);
    }
  }
  gsync(c);
#pragma unroll
  for (int j = M - 1; j >= 0; --j) {
    double2* xb = gather_buf(c, j & 1);
    xs_store<EL>(xb, what[j + 1], lane);
    gsync(c);
    ZS<EL, NC> z;
    zsums<EL, NC>(R, xb, z);
    const double inv = 1.0 / (double)(j + 1);
#pragma unroll
    for (int e = 0; e < EL; ++e) {
      what[j].u[e] -= inv * (R.kd[e] * what[j + 1].v[e]);
      what[j].v[e] -= inv * (-R.kd[e] * what[j + 1].u[e]);
    }
#pragma unroll
    for (int dd = 0; dd <= j; ++dd)
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        double au = 0.0, av = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const double2 cc = c.cv[dd * NC + k];
          au = fma(cc.y, z.Su[e][k], fma(cc.x, z.Kv[e][k], au));
          av = fma(cc.y, z.Sv[e][k], fma(-cc.x, z.Ku[e][k], av));
        }
        what[j - dd].u[e] = fma(-inv, au, what[j - dd].u[e]);
        what[j - dd].v[e] = fma(-inv, av, what[j - dd].v[e]);
      }
    if (GRAD) {
#pragma unroll
      for (int i = 0; i <= j; ++i) {
        const int rr = j - i;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          double sK = 0.0, sS = 0.0;
#pragma unroll
          for (int e = 0; e < EL; ++e) {
            sK = fma(wh[i].v[e], z.Ku[e][k], fma(-wh[i].u[e], z.Kv[e][k], sK));
            sS = fma(wh[i].u[e], z.Su[e][k], fma(wh[i].v[e], z.Sv[e][k], sS));
          }
          gK[rr][k] = fma(inv, sK, gK[rr][k]);
          gS[rr][k] = fma(-inv, sS, gS[rr][k]);
        }
      }
    }
  }
  out = what[0];
}

template <int EL, int NC>
__device__ __forceinline__ void precond_fast(const RegOps<EL, NC>& R, Vec<EL>& x) {  // preconditioners.jl:108-126
  // the two pivot divisions of the reference are multiplications by reciprocals formed once per sweep (<= 1 ulp apart)
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    const double xv = (x.v[e] - x.u[e] * R.pr_ratio[e]) * R.pr_rden[e];
    const double xu = (x.u[e] - R.pr_up[e] * xv) * R.pr_rdg[e];
    x.u[e] = xu; x.v[e] = xv;
  }
}

// ---- Tensor memory as a scratchpad ---------------------------------------------------------------------
// The FP64 path never touches the 5th-generation tensor cores, so their 256 KB of TMEM per SM is idle: each
// warp parks Krylov vectors in the 32 TMEM lanes it may address (lane quarter = warp % 4) with tcgen05.st and
// reads them back with tcgen05.ld (32x32b shape: lane l <-> TMEM lane, 4*EL consecutive 32-bit columns = its
// 2*EL doubles of one vector).  Measured on B200 (tools/tmem_test.cu): 23 cycles load-to-use, > 380 B/clk/SM.
__device__ __forceinline__ uint32_t tmem_alloc_cols(uint32_t* slot_smem, uint32_t ncols) {  // whole CTA calls; returns the base address
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot_smem;
}
__device__ __forceinline__ void tmem_free_cols(uint32_t base, uint32_t ncols) {  // whole CTA calls
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void u2(double v, uint32_t& lo, uint32_t& hi) { lo = (uint32_t)__double2loint(v); hi = (uint32_t)__double2hiint(v); }
__device__ __forceinline__ double d2(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

template <int EL>
__device__ __forceinline__ void tmem_store(uint32_t taddr, const Vec<EL>& a) {
  static_assert(EL == 1 || EL == 2, "TMEM tier is built for EL <= 2");
  uint32_t r[4 * EL];
#pragma unroll
  for (int e = 0; e < EL; ++e) { u2(a.u[e], r[4 * e], r[4 * e + 1]); u2(a.v[e], r[4 * e + 2], r[4 * e + 3]); }
  if constexpr (EL == 2)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
  else
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// issue only; the registers are valid after tmem_wait_ld()
template <int EL>
__device__ __forceinline__ void tmem_load_issue(uint32_t taddr, uint32_t (&r)[4 * EL]) {
  if constexpr (EL == 2)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  else
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int EL>
__device__ __forceinline__ void tmem_unpack(const uint32_t (&r)[4 * EL], Vec<EL>& a) {
#pragma unroll
  for (int e = 0; e < EL; ++e) { a.u[e] = d2(r[4 * e], r[4 * e + 1]); a.v[e] = d2(r[4 * e + 2], r[4 * e + 3]); }
}
template <int EL>
__device__ __forceinline__ void tmem_load(uint32_t taddr, Vec<EL>& a) {
  uint32_t r[4 * EL];
  tmem_load_issue<EL>(taddr, r);
  tmem_wait_ld();
  tmem_unpack<EL>(r, a);
}

// ---- Krylov basis access: TMEM tier, shared-memory tier, L2 tail ---------------------------------------
template <int EL, int RS>
__device__ __forceinline__ void basis_load(const FastCtx<EL, RS>& c, int i, Vec<EL>& a) {
  if (i < c.KT) tmem_load<EL>(c.tm + 4 * EL * i, a);
  else if (i < c.KT + c.KS) v2_load<EL>(a, c.Vs + (size_t)(i - c.KT) * 32 * EL, c.lane);
  else v2_load_cg<EL>(a, c.Vg + (size_t)(i - c.KT - c.KS) * 32 * EL, c.lane);
}
template <int EL, int RS>
__device__ __forceinline__ void basis_store(const FastCtx<EL, RS>& c, int i, const Vec<EL>& a) {
  if (i < c.KT) tmem_store<EL>(c.tm + 4 * EL * i, a);
  else if (i < c.KT + c.KS) v2_store<EL>(c.Vs + (size_t)(i - c.KT) * 32 * EL, a, c.lane);
  else v2_store<EL>(c.Vg + (size_t)(i - c.KT - c.KS) * 32 * EL, a, c.lane);
}

__device__ __forceinline__ int roff(int j) { return (j * (j + 1)) >> 1; }

// Load from the packed Hessenberg / R workspace of the warp.  The workspace is L2-resident global memory at full batch
// (cache-global: it is written by other lanes of the warp) and shared memory when few columns are in flight
// (FastCfg::h_smem): a generic-address load serves both, the cache operator is a hint that shared memory ignores.
__device__ __forceinline__ double2 hld2(const double* p) {  // 16-byte aligned pair
  double2 v;
  asm volatile("ld.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
// Offset of column j (rows 0..j+1) of the packed Hessenberg matrix.  QGD_QR_ROWS4 = 1: every column starts on a 32-byte
// boundary (its j + 2 entries padded to a multiple of 4), so that a lane walks down ITS column four rows -- one 32-byte
// sector -- per request in the end-of-solve QR (qr_rotation_phase) instead of one row per request, with two groups in
// flight.  MEASURED ON B200 (round 2, profiles/r02_kernel_variants.txt): 353 vs 385 evals/s at batch 592 (forward sweep
// 760 vs 701 ms) and no change for a single evaluation -- a quarter of the scattered requests, but 24 more live registers
// and a longer per-solve code path beside a GMRES loop that has to stay in the instruction cache.  Off by default.
#ifndef QGD_QR_ROWS4
#define QGD_QR_ROWS4 0
#endif
__device__ __host__ __forceinline__ int hpk(int j) {
#if QGD_QR_ROWS4
  // 4 * sum_{c<j} ceil((c+2)/4) = 4 * (S(j+4) - 1), S(n) = sum_{t<=n} floor(t/4) = 2q(q-1) + q(r+1), n = 4q + r
  const int n = j + 4, q = n >> 2, r = n & 3;
  return 4 * (2 * q * (q - 1) + q * (r + 1) - 1);
#else
  return (j * (j + 3)) >> 1;
#endif
}
__device__ __forceinline__ double hld(const double* p) {
  double v;
  asm volatile("ld.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// LinearAlgebra.givensAlgorithm for reals with one reciprocal square root instead of sqrt + two divisions
__device__ __forceinline__ void givens_fast(double f, double g, double& cs, double& sn) {
  if (g == 0.0) { cs = 1.0; sn = 0.0; }
  else if (f == 0.0) { cs = 0.0; sn = 1.0; }
  else {
    const double rr = rsqrt(fma(f, f, g * g));
    cs = f * rr; sn = g * rr;
    if (fabs(f) > fabs(g) && cs < 0.0) { cs = -cs; sn = -sn; }
  }
}

// Solve R y = g (R upper triangular, packed columns in L2 with RECIPROCAL diagonal), y overwrites c.g (shared).
// Column j-1 is fetched while column j is being eliminated.
template <int EL, int RS>
__device__ __forceinline__ void trsv_fast(const FastCtx<EL, RS>& c, int width) {
  const int lane = c.lane;
  constexpr int CH = 4;  // rows handled per lane: width <= restart <= 128
  double cur[CH], nxt[CH], dcur, dnxt = 0.0;
  {
    const double* col = c.Rg + roff(width - 1);
#pragma unroll
    for (int q = 0; q < CH; ++q) { const int i = lane + 32 * q; cur[q] = i < width - 1 ? hld(col + i) : 0.0; }
    dcur = hld(col + width - 1);
  }
  __syncwarp();
  for (int j = width - 1; j >= 0; --j) {
    if (j > 0) {
      const double* col = c.Rg + roff(j - 1);
#pragma unroll
      for (int q = 0; q < CH; ++q) { const int i = lane + 32 * q; nxt[q] = i < j - 1 ? hld(col + i) : 0.0; }
      dnxt = hld(col + j - 1);
    }
    const double yj = c.g[j] * dcur;  // dcur = 1 / R[j][j]
    __syncwarp();
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int i = lane + 32 * q;
      if (i < j) c.g[i] = fma(-yj, cur[q], c.g[i]);
      else if (i == j) c.g[i] = yj;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < CH; ++q) cur[q] = nxt[q];
    dcur = dnxt;
  }
}

// One modified Gram-Schmidt step against basis vector i (already in registers) fused with the progressive
// Givens update of the Hessenberg column: branch free; nul = nullv[i], rt = rot[i] (rot[0] = identity).
template <int EL, int RS>
__device__ __forceinline__ void mgs_step(const FastCtx<EL, RS>& c, int i, const Vec<EL>& vi, Vec<EL>& w, double& dsum, double& hprev, double* rcol,
                                         double nul, double2 rt) {
  const double h = row_allsum(c, vdot_local<EL>(vi, w));
  vaxpy(w, -h, vi);
  dsum = fma(nul, h, dsum);
  const double r = rt.x * hprev + rt.y * h;  // row i-1 of R, final after rotation i-1
  hprev = -rt.y * hprev + rt.x * h;
  if (i > 0) rcol[i - 1] = r;                 // every lane, same address, same value
}

// GMRES for the time-stepping solves (fixed absolute tolerance, restart = maxiter = 2N).
// OP: apply(in, out) = A in; the left preconditioner is applied here.  Returns the number of iterations.
template <int EL, int NC, class OP, int RS>
__device__ int gmres_fast_strict(const FastCtx<EL, RS>& c, const RegOps<EL, NC>& R, const OP& op, Vec<EL>& x, const Vec<EL>& b, double tol,
                                 int restart, int maxiter, double reltol = -1.0) {
  const int lane = c.lane;
  Vec<EL> v, w;
  op.apply(x, w);
#pragma unroll
  for (int e = 0; e < EL; ++e) { v.u[e] = b.u[e] - w.u[e]; v.v[e] = b.v[e] - w.v[e]; }
  precond_fast<EL, NC>(R, v);
  double beta2 = row_allsum(c, vdot_local<EL>(v, v));
  double rbeta = rsqrt(beta2), beta = beta2 * rbeta;
  // reltol >= 0 (terminal-condition solves, gmres! driver): tol = max(reltol * ||initial residual||, tol) as IterativeSolvers
  // sets it; the time-stepping solves pass -1 (fixed absolute tolerance)
  if (reltol >= 0.0) tol = fmax(reltol * beta, tol);
  vscale(v, rbeta);
  basis_store<EL>(c, 0, v);
  double cur = beta, res_beta = beta, accum = 1.0, gcur = beta;
  __syncwarp();
  if (lane == 0) { c.nullv[0] = 1.0; c.rot[0] = make_double2(1.0, 0.0); }
  __syncwarp();
  int k = 1, it = 0;
  while (it < maxiter && cur > tol) {
    op.apply(v, w);  // expand!
    precond_fast<EL, NC>(R, w);
    // modified Gram-Schmidt against V[:, 0..k-1], tier by tier; the next vector and its scalars are fetched while
    // the current one is being reduced
    double dsum = 0.0, hprev = 0.0;
    double* rcol = c.Rg + roff(k - 1);
    const int nT = min(k, c.KT), nS = min(k, c.KT + c.KS);
    int i = 0;
    if (nT > 0) {  // ---- TMEM tier
      uint32_t rn[4 * EL];
      Vec<EL> vi;
      tmem_load<EL>(c.tm, vi);
      double nul = c.nullv[0];
      double2 rt = c.rot[0];
      for (; i < nT; ++i) {
        const int inext = min(i + 1, c.KT - 1);  // always a valid slot: no predicate in the loop body
        tmem_load_issue<EL>(c.tm + 4 * EL * inext, rn);
        const double nul_n = c.nullv[i + 1];
        const double2 rt_n = c.rot[i + 1];
        mgs_step<EL>(c, i, vi, w, dsum, hprev, rcol, nul, rt);
        tmem_wait_ld();
        tmem_unpack<EL>(rn, vi);
        nul = nul_n; rt = rt_n;
      }
    }
    if (i < nS) {  // ---- shared-memory tier
      Vec<EL> vi, vn;
      v2_load<EL>(vi, c.Vs + (size_t)(i - c.KT) * 32 * EL, lane);
      double nul = c.nullv[i];
      double2 rt = c.rot[i];
      for (; i < nS; ++i) {
        const int inext = min(i + 1 - c.KT, c.KS - 1);
        v2_load<EL>(vn, c.Vs + (size_t)inext * 32 * EL, lane);
        const double nul_n = c.nullv[i + 1];
        const double2 rt_n = c.rot[i + 1];
        mgs_step<EL>(c, i, vi, w, dsum, hprev, rcol, nul, rt);
        vi = vn; nul = nul_n; rt = rt_n;
      }
    }
    if (i < k) {  // ---- L2 tail (two vectors in flight)
      Vec<EL> vi, vn, vnn;
      const double2* base = c.Vg - (size_t)(c.KT + c.KS) * 32 * EL;
      v2_load_cg<EL>(vi, base + (size_t)i * 32 * EL, lane);
      v2_load_cg<EL>(vn, base + (size_t)min(i + 1, k - 1) * 32 * EL, lane);
      double nul = c.nullv[i];
      double2 rt = c.rot[i];
      for (; i < k; ++i) {
        v2_load_cg<EL>(vnn, base + (size_t)min(i + 2, k - 1) * 32 * EL, lane);
        const double nul_n = c.nullv[i + 1];
        const double2 rt_n = c.rot[i + 1];
        mgs_step<EL>(c, i, vi, w, dsum, hprev, rcol, nul, rt);
        vi = vn; vn = vnn; nul = nul_n; rt = rt_n;
      }
    }
    const double nrm2 = row_allsum(c, vdot_local<EL>(w, w));
    // 1/||w|| and ||w|| from one reciprocal square root; an exactly vanishing w (the Krylov space is exhausted: happy
    // breakdown) gives H[k+1][k] = 0 as in the reference (its estimate then drops to zero and the solve ends), not 0 * inf
    const double rnrm = rsqrt(nrm2), nrm = nrm2 > 0.0 ? nrm2 * rnrm : 0.0;
    vscale(w, rnrm);
    basis_store<EL>(c, k, w);
    {  // new rotation (k-1) from (hprev, nrm); update the rotated right-hand side
      double cs, sn;
      givens_fast(hprev, nrm, cs, sn);
      const double rkk = cs * hprev + sn * nrm;
      rcol[k - 1] = 1.0 / rkk;  // the triangular solve only ever divides by the diagonal: keep its reciprocal
      if (lane == 0) {
        c.rot[k] = make_double2(cs, sn);
        c.g[k - 1] = cs * gcur;
      }
      gcur = -sn * gcur;
    }
    const double nv = -(dsum * rnrm);  // update_residual!
    if (lane == 0) c.nullv[k] = nv;
    accum = fma(nv, nv, accum);
    cur = res_beta * rsqrt(accum);
    k += 1;
    v = w;
    __syncwarp();
    if (k == restart + 1 || cur <= tol) {
      const int width = k - 1;
      trsv_fast<EL>(c, width);
      Vec<EL> vi;
      basis_load<EL>(c, 0, vi);
      for (int j = 0; j < width; ++j) {  // update_solution!: x += V[:, 0..width-1] y
        Vec<EL> vn;
        if (j + 1 < width) basis_load<EL>(c, j + 1, vn);
        vaxpy(x, c.g[j], vi);
        if (j + 1 < width) vi = vn;
      }
      k = 1;
      if (cur > tol) {  // restart (residual.current keeps its value, as in the package)
        op.apply(x, w);
#pragma unroll
        for (int e = 0; e < EL; ++e) { v.u[e] = b.u[e] - w.u[e]; v.v[e] = b.v[e] - w.v[e]; }
        precond_fast<EL, NC>(R, v);
        beta2 = row_allsum(c, vdot_local<EL>(v, v));
        rbeta = rsqrt(beta2); beta = beta2 * rbeta;
        vscale(v, rbeta);
        basis_store<EL>(c, 0, v);
        accum = 1.0; res_beta = beta; gcur = beta;
        __syncwarp();
        if (lane == 0) c.nullv[0] = 1.0;
      }
      __syncwarp();
    }
    it += 1;
  }
  return it;
}

// ========================================================================================================
// Blocked orthogonalisation (QGD_MGS_BLOCK = 4 or 8)
// ========================================================================================================

// Sum over the warp of BLK values per lane at once; every lane receives all BLK totals (also left in slot[0..BLK)).
// log2(BLK) exchange-and-halve shuffle stages leave one value per lane (value index = lane bits 4..), one DMMA with a
// ones operand sums the remaining lane bits, lanes 0..3 publish the totals through shared memory.
// Measured on B200 (tools/microbench.cu): DMMA.8x8x4 costs 4 SM-cycles, a 64-bit SHFL 2, so the per-value cost drops
// from 8.5 (2 DMMA + DADD each) to about 3 SM-cycles.
template <int BLK>
__device__ __forceinline__ void block_allsum(double (&p)[BLK], double* slot, int lane) {
  static_assert(BLK == 4 || BLK == 8, "block width");
  const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
  double a[BLK / 2];
#pragma unroll
  for (int q = 0; q < BLK / 2; ++q) {
    const double keep = up16 ? p[q + BLK / 2] : p[q], send = up16 ? p[q] : p[q + BLK / 2];
    a[q] = keep + __shfl_xor_sync(FULL_MASK, send, 16);
  }
  double b[BLK / 4];
#pragma unroll
  for (int q = 0; q < BLK / 4; ++q) {
    const double keep = up8 ? a[q + BLK / 4] : a[q], send = up8 ? a[q] : a[q + BLK / 4];
    b[q] = keep + __shfl_xor_sync(FULL_MASK, send, 8);
  }
  double c0, c1;
  if constexpr (BLK == 8) {
    const bool up4 = (lane & 4) != 0;
    const double keep = up4 ? b[1] : b[0], send = up4 ? b[0] : b[1];
    const double r = keep + __shfl_xor_sync(FULL_MASK, send, 4);
    dmma884(c0, c1, 1.0, r);  // lane l: totals of values 2(l%4), 2(l%4)+1
    if (lane < 4) reinterpret_cast<double2*>(slot)[lane] = make_double2(c0, c1);
  } else {
    dmma884(c0, c1, 1.0, b[0]);  // lane l: the two halves of the total of value l%4
    if (lane < 4) slot[lane] = c0 + c1;
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < BLK; q += 2) { const double2 t = reinterpret_cast<const double2*>(slot)[q >> 1]; p[q] = t.x; p[q + 1] = t.y; }
}

// Same result through shared memory (variant bit 1): the 8 x 32 partials are transposed through the
// (idle) gather buffer instead of three select-and-shuffle stages -- 8 STS.64 + 8 LDS.64 + 7 DADD instead of
// 28 selects + 14 SHFL + 7 DADD per block, and two shared-memory round trips shorter.  Value q of lane l sits at
// T[32 q + (l ^ 4 (q & 3))]: stores are a permutation of a row, and the reader of value qv = lane / 4, s = lane % 4
// (which sums the partials of lanes 4 t + s, t = 0..7) hits 16 distinct 8-byte banks per half warp.
__device__ __forceinline__ void block_allsum8_smem(double (&p)[8], double* T, double* slot, int lane) {
#pragma unroll
  for (int q = 0; q < 8; ++q) T[32 * q + (lane ^ (4 * (q & 3)))] = p[q];
  __syncwarp();
  const int qv = lane >> 2, s = lane & 3;
  const double* row = T + 32 * qv + s;
  const int sw = qv & 3;
  double a[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) a[t] = row[4 * (t ^ sw)];
  const double r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  double c0, c1;
  dmma884(c0, c1, 1.0, r);  // lane l: totals of values 2(l%4), 2(l%4)+1
  if (lane < 4) reinterpret_cast<double2*>(slot)[lane] = make_double2(c0, c1);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; q += 2) { const double2 t = reinterpret_cast<const double2*>(slot)[q >> 1]; p[q] = t.x; p[q + 1] = t.y; }
}

// Same result with tensor-core reductions only (variant bit 2): per value one DMMA with a ones A operand sums the
// 4-lane groups (lane l receives the sums of groups 2(l%4), 2(l%4)+1), their sum fed back as the A operand of a second
// DMMA against ones sums the remaining four -- 16 independent, pipelined DMMA and 8 DADD per block, no shared-memory
// round trip and no warp synchronisation; every lane ends up with all 8 totals.  The coefficients still have to reach
// c.hcol (the Hessenberg column): lane 0 stores them.
__device__ __forceinline__ void block_allsum8_dmma(double (&p)[8], double* slot, int lane) {
  double t[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { double c0, c1; dmma884(c0, c1, 1.0, p[q]); t[q] = c0 + c1; }
#pragma unroll
  for (int q = 0; q < 8; ++q) { double d0, d1; dmma884(d0, d1, t[q], 1.0); p[q] = d0; }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 8; q += 2) reinterpret_cast<double2*>(slot)[q >> 1] = make_double2(p[q], p[q + 1]);
  }
}

// tcgen05.ld of N consecutive 32-bit columns of this warp's 32 TMEM lanes (issue only)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                 "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                 "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]) {
  static_assert(N == 16 || N == 32 || N == 64, "TMEM block load width");
  if constexpr (N == 16) tmem_ld16(taddr, r);
  else if constexpr (N == 32) tmem_ld32(taddr, r);
  else {
    uint32_t (&lo)[32] = reinterpret_cast<uint32_t (&)[32]>(r[0]);
    uint32_t (&hi)[32] = reinterpret_cast<uint32_t (&)[32]>(r[32]);
    tmem_ld32(taddr, lo);
    tmem_ld32(taddr + 32, hi);
  }
}

// Basis vectors i0 .. i0+BLK-1 of one tier (0: TMEM, 1: shared memory, 2: L2).  Tier boundaries are multiples of
// BLK (plan_fast), so a block never straddles two tiers; slots past the newest vector hold stale but addressable data.
template <int EL, int BLK, int TIER, int RS>
__device__ __forceinline__ void gs_load_block(const FastCtx<EL, RS>& c, int i0, Vec<EL> (&vb)[BLK]) {
  if constexpr (TIER == 0) {
    uint32_t r[4 * EL * BLK];
    tmem_ld<4 * EL * BLK>(c.tm + 4 * EL * i0, r);
    tmem_wait_ld();
#pragma unroll
    for (int q = 0; q < BLK; ++q)
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        vb[q].u[e] = d2(r[4 * EL * q + 4 * e], r[4 * EL * q + 4 * e + 1]);
        vb[q].v[e] = d2(r[4 * EL * q + 4 * e + 2], r[4 * EL * q + 4 * e + 3]);
      }
  } else if constexpr (TIER == 1) {
    const double2* p = c.Vs + (size_t)(i0 - c.KT) * 32 * EL;
#pragma unroll
    for (int q = 0; q < BLK; ++q) v2_load<EL>(vb[q], p + (size_t)q * 32 * EL, c.lane);
  } else {
    const double2* p = c.Vg + (size_t)(i0 - c.KT - c.KS) * 32 * EL;
#pragma unroll
    for (int q = 0; q < BLK; ++q) v2_load_cg<EL>(vb[q], p + (size_t)q * 32 * EL, c.lane);
  }
}

// One block: h_q = <v_{i0+q}, w> for the whole block from the same w, then w -= sum_{q < nb} h_q v_{i0+q}.
// The coefficients are left in c.hcol[i0 .. i0+BLK).
template <int EL, int BLK, int VARIANT, int RS>
__device__ __forceinline__ void gs_block(const FastCtx<EL, RS>& c, int i0, int nb, const Vec<EL> (&vb)[BLK], Vec<EL>& w) {
  double h[BLK];
#pragma unroll
  for (int q = 0; q < BLK; ++q) h[q] = vdot_local<EL>(vb[q], w);
  if constexpr (RS > 1 && QGD_RS_FUSED_RED && BLK == 8 && (VARIANT & 6) == 2) {
    block_allsum8_group(c, h, reinterpret_cast<double*>(c.xs), c.hcol + i0);
  } else {
    if constexpr (BLK == 8 && (VARIANT & 4) != 0) block_allsum8_dmma(h, c.hcol + i0, c.lane);
    else if constexpr (BLK == 8 && (VARIANT & 2) != 0) block_allsum8_smem(h, reinterpret_cast<double*>(c.xs), c.hcol + i0, c.lane);
    else block_allsum<BLK>(h, c.hcol + i0, c.lane);
    if constexpr (RS > 1) {
      static_assert(BLK == 8 && (VARIANT & 4) == 0, "row-split groups: blocks of 8, coefficients published through shared memory");
      row_combine8(c, h, c.hcol + i0);
    }
  }
#pragma unroll
  for (int q = 0; q < BLK; ++q)
    if (q < nb) vaxpy(w, -h[q], vb[q]);
}

// Row-split groups run Krylov spaces of 150-250 vectors of which 40-48 per warp are on chip: the L2 tier is most of the basis
// and a warp has no registers left to prefetch it.  The block after the current one is therefore copied L2 -> shared memory
// by cp.async (LDGSTS) into a staging block carved from the shared-memory tier (QGD_RS_STAGE_L2), the first L2 block of an
// orthogonalisation already while the TMEM and shared-memory tiers are being swept.  Every lane copies and later reads only
// its own 16-byte slots, so cp.async.wait_group is the only synchronisation.
// MEASURED ON B200 (round 2, profiles/r02_row_split_groups.txt): correct (all parity tests green) but SLOWER -- N = 125, 74 control
// vectors x 60 steps: forward 109 -> 125 ms, adjoint 166 -> 171 ms; one evaluation 158 -> 165 ms.  The staging block costs 8 of the
// 16 resident shared-memory vectors, the staged path needs the one-copy loop shape that costs the forward kernel its three
// specialised tier loops, and the extra shared-memory round trip of every L2 block is not cheaper than the exposed L2 latency it
// replaces.  Off by default; kept as the record of the experiment.
#ifndef QGD_RS_STAGE_L2
#define QGD_RS_STAGE_L2 0
#endif
__device__ __forceinline__ void cp16(double2* dst_smem, const double2* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
template <int EL, int RS>
__device__ __forceinline__ void stage_issue(const FastCtx<EL, RS>& c, int i0) {  // L2-tier block i0 .. i0+7 -> staging
  const double2* p = c.Vg + (size_t)(i0 - c.KT - c.KS) * 32 * EL + c.lane;
  double2* dst = c.stage + c.lane;
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int e = 0; e < EL; ++e) cp16(dst + q * 32 * EL + 32 * e, p + q * 32 * EL + 32 * e);
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int EL, int RS>
__device__ __forceinline__ void stage_take(const FastCtx<EL, RS>& c, Vec<EL> (&vb)[8]) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int q = 0; q < 8; ++q) v2_load<EL>(vb[q], c.stage + (size_t)q * 32 * EL, c.lane);
}

// w is orthogonalised against V[:, 0..k-1]; c.hcol[0..k-1] receives the coefficients.
template <int EL, int BLK, int VARIANT, int RS>
__device__ __forceinline__ void gs_orthogonalize(const FastCtx<EL, RS>& c, int k, Vec<EL>& w) {
  __syncwarp();  // the gather buffer (read by the operator application) becomes the transposition buffer
  if constexpr (RS > 1 && QGD_RS_STAGE_L2 && BLK == 8) {
    if (c.stage != nullptr) {
      const int bT = c.KT, bS = c.KT + c.KS;
      if (k > bS) stage_issue(c, bS);
      for (int i0 = 0; i0 < k; i0 += 8) {
        Vec<EL> vb[8];
        if (i0 < bT) gs_load_block<EL, 8, 0>(c, i0, vb);
        else if (i0 < bS) gs_load_block<EL, 8, 1>(c, i0, vb);
        else {
          stage_take(c, vb);
          if (i0 + 8 < k) stage_issue(c, i0 + 8);
        }
        gs_block<EL, 8, VARIANT>(c, i0, k - i0, vb, w);
      }
      return;
    }
  }
  if constexpr ((VARIANT & 1) != 0) {
  // one copy of the block arithmetic for the three tiers: the hot loop has to fit the 32 KB instruction cache
  const int bT = c.KT, bS = c.KT + c.KS;
  for (int i0 = 0; i0 < k; i0 += BLK) {
    Vec<EL> vb[BLK];
    if (i0 < bT) gs_load_block<EL, BLK, 0>(c, i0, vb);
    else if (i0 < bS) gs_load_block<EL, BLK, 1>(c, i0, vb);
    else gs_load_block<EL, BLK, 2>(c, i0, vb);
    gs_block<EL, BLK, VARIANT>(c, i0, k - i0, vb, w);
  }
  } else {
  const int eT = min(k, c.KT), eS = min(k, c.KT + c.KS);
  int i0 = 0;
  for (; i0 < eT; i0 += BLK) { Vec<EL> vb[BLK]; gs_load_block<EL, BLK, 0>(c, i0, vb); gs_block<EL, BLK, VARIANT>(c, i0, k - i0, vb, w); }
  for (; i0 < eS; i0 += BLK) { Vec<EL> vb[BLK]; gs_load_block<EL, BLK, 1>(c, i0, vb); gs_block<EL, BLK, VARIANT>(c, i0, k - i0, vb, w); }
  for (; i0 < k; i0 += BLK) { Vec<EL> vb[BLK]; gs_load_block<EL, BLK, 2>(c, i0, vb); gs_block<EL, BLK, VARIANT>(c, i0, k - i0, vb, w); }
  }
}

// ---- two-pass super-block orthogonalisation (QGD_GS_SUPER) ------------------------------------------------
// This lane's share of the transposed partial sums: value q = lane / 4, quarter s = lane % 4 (sum over the lanes
// 4 t + s).  Same swizzled layout as block_allsum8_smem.
__device__ __forceinline__ double transpose_partials8(const double (&p)[8], double* T, int lane) {
#pragma unroll
  for (int q = 0; q < 8; ++q) T[32 * q + (lane ^ (4 * (q & 3)))] = p[q];
  __syncwarp();
  const int qv = lane >> 2, s = lane & 3;
  const double* row = T + 32 * qv + s;
  const int sw = qv & 3;
  double a[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) a[t] = row[4 * (t ^ sw)];
  return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

template <int EL, int RS>
__device__ __forceinline__ void gs_load_group(const FastCtx<EL, RS>& c, int i0, Vec<EL> (&vb)[8]) {
  if (i0 < c.KT) gs_load_block<EL, 8, 0>(c, i0, vb);
  else if (i0 < c.KT + c.KS) gs_load_block<EL, 8, 1>(c, i0, vb);
  else gs_load_block<EL, 8, 2>(c, i0, vb);
}

template <int EL, int SUPER, int RS>
__device__ __forceinline__ void gs_orthogonalize_super(const FastCtx<EL, RS>& c, int k, Vec<EL>& w) {
  static_assert(SUPER % 8 == 0 && SUPER >= 8, "super-block width");
  double* T = reinterpret_cast<double*>(c.xs);
  const int lane = c.lane;
#pragma unroll 1
  for (int s0 = 0; s0 < k; s0 += SUPER) {
    const int s1 = min(k, s0 + SUPER);
    // pass 1: h_i = <v_i, w> for the whole super-block from the same w
#pragma unroll 1
    for (int i0 = s0; i0 < s1; i0 += 8) {
      Vec<EL> vb[8];
      gs_load_group<EL>(c, i0, vb);
      double h[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) h[q] = vdot_local<EL>(vb[q], w);
      __syncwarp();  // the previous group's partials have been read
      const double r = transpose_partials8(h, T, lane);
      double c0, c1;
      dmma884(c0, c1, 1.0, r);  // lane l: totals of values 2(l%4), 2(l%4)+1
      if (lane < 4) reinterpret_cast<double2*>(c.hcol + i0)[lane] = make_double2(c0, c1);
    }
    __syncwarp();
    // pass 2: w -= sum_i h_i v_i
#pragma unroll 1
    for (int i0 = s0; i0 < s1; i0 += 8) {
      Vec<EL> vb[8];
      gs_load_group<EL>(c, i0, vb);
      double h[8];
#pragma unroll
      for (int q = 0; q < 8; q += 2) { const double2 t = reinterpret_cast<const double2*>(c.hcol + i0)[q >> 1]; h[q] = t.x; h[q + 1] = t.y; }
      const int nb = k - i0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < nb) vaxpy(w, -h[q], vb[q]);
    }
  }
}

// ========================================================================================================
// Latency team (round 2): ONE column on FOUR warps of an otherwise empty SM
// ========================================================================================================
// When no more columns are in flight than the GPU has SMs (a single gradient evaluation, what optimize_gate asks for:
// src/ipopt_optimal_control.jl:257,304) a warp is alone on its SM and its GMRES iteration is one long dependent chain,
// 45 % of it the orthogonalisation (profiles/r02_ncu_fwd_b1.txt).  The team spreads that part: the Krylov basis is dealt
// out in blocks of 8 vectors to the four warps of the CTA -- block b lives in the tensor memory of warp b % 4 (each warp
// owns a 32-lane quarter: 64 vectors each, the whole basis on chip) -- and a Gram-Schmidt step takes the coefficients of
// FOUR blocks (one per warp) from the same vector: every warp loads its block, forms its 8 dot products and reductions and
// its part of the update  sum_q h_q v_q, the parts are exchanged through shared memory and everybody applies all four.
// Classical Gram-Schmidt inside a super-block of 32, modified across super-blocks (tools/gs_block_experiment.py: block
// widths up to the whole basis leave every iteration count of the C2 problem unchanged).  Warp 0 (the main warp) runs
// everything else of the sweep unchanged; warps 1-3 wait at a named barrier for its commands.
#define QGD_TEAM_WARPS 4
enum { TEAM_CMD_GS = 1, TEAM_CMD_UPD = 2, TEAM_CMD_EXIT = 3 };

template <int EL>
struct TeamView {
  int* cmd;      // [0] command, [1] k (GS) or width (UPD), [2] index of a basis vector waiting in VN for its owner (-1: none)
  double2* W;    // the vector being orthogonalised
  double2* VN;   // a new basis vector on its way to the tensor memory of its owner
  double2* C;    // [2][4] update parts, double buffered over the super-blocks
  double* T;     // [3] transposition buffers of the helper warps (the main warp uses its gather buffer)
};
template <int EL>
__host__ __device__ constexpr int team_doubles() { return 4 + 10 * (2 * 32 * EL) + 3 * 256; }
template <int EL>
__device__ __forceinline__ TeamView<EL> team_view(double* base) {
  constexpr int VD = 2 * 32 * EL;
  TeamView<EL> t;
  t.cmd = reinterpret_cast<int*>(base);
  t.W = reinterpret_cast<double2*>(base + 4);
  t.VN = reinterpret_cast<double2*>(base + 4 + VD);
  t.C = reinterpret_cast<double2*>(base + 4 + 2 * VD);
  t.T = base + 4 + 10 * VD;
  return t;
}
__device__ __forceinline__ void team_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ int team_owner(int i) { return (i >> 3) & 3; }
__device__ __forceinline__ int team_slot(int i) { return ((i >> 5) << 3) + (i & 7); }

// The block of warp `wid` in super-block `sb`: coefficients into hcol, update part into Cdst.
template <int EL>
__device__ __forceinline__ void team_block(uint32_t tm_own, int wid, int sb, int k, const Vec<EL>& w, double* T, double* hcol,
                                           double2* Cdst, int lane) {
  const int i0 = 8 * (4 * sb + wid);
  Vec<EL> corr;
  vzero(corr);
  if (i0 < k) {
    Vec<EL> vb[8];
    uint32_t r[4 * EL * 8];
    tmem_ld<4 * EL * 8>(tm_own + 4 * EL * (8 * sb), r);
    tmem_wait_ld();
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        vb[q].u[e] = d2(r[4 * EL * q + 4 * e], r[4 * EL * q + 4 * e + 1]);
        vb[q].v[e] = d2(r[4 * EL * q + 4 * e + 2], r[4 * EL * q + 4 * e + 3]);
      }
    double h[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) h[q] = vdot_local<EL>(vb[q], w);
    __syncwarp();
    block_allsum8_smem(h, T, hcol + i0, lane);
    const int nb = k - i0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < nb) vaxpy(corr, h[q], vb[q]);
  }
  v2_store<EL>(Cdst, corr, lane);
}

template <int EL>
__device__ __forceinline__ void team_apply_parts(const double2* Cpar, Vec<EL>& w, int lane) {
#pragma unroll
  for (int wid = 0; wid < QGD_TEAM_WARPS; ++wid) {
    Vec<EL> cw;
    v2_load<EL>(cw, Cpar + (size_t)wid * 32 * EL, lane);
    vaxpy(w, -1.0, cw);
  }
}

// main warp: orthogonalise w against V[:, 0..k-1]; coefficients in c.hcol[0..k)
template <int EL, int RS>
__device__ __forceinline__ void gs_team_main(const FastCtx<EL, RS>& c, int k, Vec<EL>& w, int pending) {
  const TeamView<EL> tv = team_view<EL>(c.team);
  const int lane = c.lane;
  v2_store<EL>(tv.W, w, lane);
  if (lane == 0) { tv.cmd[0] = TEAM_CMD_GS; tv.cmd[1] = k; tv.cmd[2] = pending; }
  team_bar();
  const int nsb = (((k + 7) >> 3) + 3) >> 2;
  for (int sb = 0; sb < nsb; ++sb) {
    double2* Cpar = tv.C + (size_t)(sb & 1) * QGD_TEAM_WARPS * 32 * EL;
    team_block<EL>(c.tm, 0, sb, k, w, reinterpret_cast<double*>(c.xs), c.hcol, Cpar, lane);
    team_bar();
    team_apply_parts<EL>(Cpar, w, lane);
  }
}

// main warp: x += V[:, 0..width-1] y with y in c.g (shared memory)
template <int EL, int RS>
__device__ __forceinline__ void update_team_main(const FastCtx<EL, RS>& c, int width, Vec<EL>& x, int pending) {
  const TeamView<EL> tv = team_view<EL>(c.team);
  const int lane = c.lane;
  if (lane == 0) { tv.cmd[0] = TEAM_CMD_UPD; tv.cmd[1] = width; tv.cmd[2] = pending; }
  team_bar();
  for (int j0 = 0; j0 < width; j0 += 32)   // blocks j0/8 + 0 of each super-block belong to warp 0
    for (int q = 0; q < 8 && j0 + q < width; ++q) {
      Vec<EL> vj;
      tmem_load<EL>(c.tm + 4 * EL * team_slot(j0 + q), vj);
      vaxpy(x, c.g[j0 + q], vj);
    }
  team_bar();
#pragma unroll
  for (int wid = 1; wid < QGD_TEAM_WARPS; ++wid) {
    Vec<EL> cw;
    v2_load<EL>(cw, tv.C + (size_t)wid * 32 * EL, lane);
    vaxpy(x, 1.0, cw);
  }
}

template <int EL, int RS>
__device__ __forceinline__ void basis_store_team(const FastCtx<EL, RS>& c, int i, const Vec<EL>& a, int& pending) {
  if (team_owner(i) == 0) { tmem_store<EL>(c.tm + 4 * EL * team_slot(i), a); }
  else { v2_store<EL>(team_view<EL>(c.team).VN, a, c.lane); pending = i; }
}

// helper warps 1..3: serve the main warp's commands until it says EXIT
template <int EL>
__device__ void team_helper_loop(double* team_base, double* hcol, const double* g, uint32_t tm_own, int wid, int lane) {
  const TeamView<EL> tv = team_view<EL>(team_base);
  double* T = tv.T + (size_t)(wid - 1) * 256;
  for (;;) {
    team_bar();
    const int cmd = tv.cmd[0], arg = tv.cmd[1], pend = tv.cmd[2];
    if (cmd == TEAM_CMD_EXIT) break;
    if (pend >= 0 && team_owner(pend) == wid) {
      Vec<EL> vn;
      v2_load<EL>(vn, tv.VN, lane);
      tmem_store<EL>(tm_own + 4 * EL * team_slot(pend), vn);
    }
    if (cmd == TEAM_CMD_GS) {
      const int k = arg;
      Vec<EL> w;
      v2_load<EL>(w, tv.W, lane);
      const int nsb = (((k + 7) >> 3) + 3) >> 2;
      for (int sb = 0; sb < nsb; ++sb) {
        double2* Cpar = tv.C + (size_t)(sb & 1) * QGD_TEAM_WARPS * 32 * EL;
        team_block<EL>(tm_own, wid, sb, k, w, T, hcol, Cpar + (size_t)wid * 32 * EL, lane);
        team_bar();
        if (sb + 1 < nsb) team_apply_parts<EL>(Cpar, w, lane);
      }
    } else {  // TEAM_CMD_UPD
      const int width = arg;
      Vec<EL> part;
      vzero(part);
      for (int j0 = 8 * wid; j0 < width; j0 += 32)
        for (int q = 0; q < 8 && j0 + q < width; ++q) {
          Vec<EL> vj;
          tmem_load<EL>(tm_own + 4 * EL * team_slot(j0 + q), vj);
          vaxpy(part, g[j0 + q], vj);
        }
      v2_store<EL>(tv.C + (size_t)wid * 32 * EL, part, lane);
      team_bar();
    }
  }
}

// ---- asynchronous 8-byte copies L2 -> shared memory (cp.async, SASS LDGSTS) -----------------------------
__device__ __forceinline__ void cp8(double* dst_smem, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// Least squares  min || H y - beta e_1 ||  for the (width+1) x width Hessenberg matrix packed in c.Rg (L2), as
// solve_least_squares! does: Givens QR, then back substitution; y is left in c.g (shared memory).
// Lanes own the columns j = lane + 32 s; rotation i is formed from the pivot held by the owner of column i and
// applied by every lane to its columns j >= i.  An L2 round trip is several rotations long, so the rows of H (and,
// in the back substitution, the columns of R) are streamed into a small shared-memory ring by cp.async a few steps
// ahead; the ring is indexed dynamically, so the loops stay rolled and the code small (the kernel has to fit the
// instruction cache).  Scratch: c.xs (ring), c.hcol (1 / R[j][j]).
template <class CTX>
__device__ __forceinline__ void qr_solve_fast(const CTX& c, int width, double beta) {
  const int lane = c.lane;
  constexpr int CH = 4;                   // width <= restart <= 128
  const int nsl = (width + 31) >> 5;
  // ring in c.xs: 4 steps x 2 column slots or 2 steps x 4 slots (256 doubles); 2 x 2 when xs only has 128 (N <= 32,
  // where width <= 64)
  const bool deep = nsl <= 2 && CTX::kRingDoubles >= 256;
  const int D = deep ? 4 : 2, dmask = D - 1, RS = nsl <= 2 ? 2 : 4;
  double* ring = reinterpret_cast<double*>(c.xs) + lane;
  double* dinv = c.hcol;
  __syncwarp();
  double* colp[CH];  // start of the owned columns of H
  double top[CH];
#pragma unroll
  for (int s = 0; s < CH; ++s) {
    const int j = lane + 32 * s;
    colp[s] = c.Rg + hpk(j);
    top[s] = j < width ? hld(colp[s]) : 0.0;
  }
  double gcur = beta;
  for (int i = -D; i < width; ++i) {  // the first D trips only fill the ring (rows 1..D)
    if (i >= 0) {
      if (deep) cp_wait<3>(); else cp_wait<1>();  // row i+1 has landed
      const int so = i >> 5;
      double a_own = top[0];
#pragma unroll
      for (int s = 1; s < CH; ++s) a_own = (so == s) ? top[s] : a_own;
      const double a = __shfl_sync(FULL_MASK, a_own, i & 31);
      const double b = c.sub[i];
      double cs, sn;
      givens_fast(a, b, cs, sn);
      const double rii = cs * a + sn * b;
      const double* rrow = ring + (((i + 1) & dmask) * RS) * 32;
#pragma unroll
      for (int s = 0; s < CH; ++s) {
        const int j = lane + 32 * s;
        if (s < nsl && j < width && j >= i) {
          const double lo = rrow[s * 32];
          const double r = cs * top[s] + sn * lo;
          top[s] = -sn * top[s] + cs * lo;
          if (j > i) colp[s][i] = r;
        }
      }
      if (lane == 0) { dinv[i] = 1.0 / rii; c.g[i] = cs * gcur; }
      gcur = -sn * gcur;
    }
    const int r = i + 1 + D;  // row to fetch now
    double* wrow = ring + ((r & dmask) * RS) * 32;
#pragma unroll
    for (int s = 0; s < CH; ++s) {
      const int j = lane + 32 * s;
      if (s < nsl && j < width && r <= j + 1) cp8(wrow + s * 32, colp[s] + r);
    }
    cp_commit();
  }
  cp_wait<0>();
  __syncwarp();
  // back substitution R y = g: right-hand side rows i = lane + 32 q in registers, column j of R through the ring
  double gi[CH];
#pragma unroll
  for (int q = 0; q < CH; ++q) { const int i = lane + 32 * q; gi[q] = i < width ? c.g[i] : 0.0; }
  for (int j = width - 1 + D; j >= 0; --j) {  // the first D trips only fill the ring (columns width-1 .. width-D)
    if (j < width) {
      if (deep) cp_wait<3>(); else cp_wait<1>();  // column j has landed
      const int so = j >> 5;
      double g_own = gi[0];
#pragma unroll
      for (int q = 1; q < CH; ++q) g_own = (so == q) ? gi[q] : g_own;
      const double yj = __shfl_sync(FULL_MASK, g_own, j & 31) * dinv[j];
      const double* rcol = ring + ((j & dmask) * RS) * 32;
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int i = lane + 32 * q;
        if (q < nsl) {
          if (i < j) gi[q] = fma(-yj, rcol[q * 32], gi[q]);
          else if (i == j) gi[q] = yj;
        }
      }
    }
    const int jn = j - D;  // column to fetch now
    if (jn >= 0) {
      double* wcol = ring + ((jn & dmask) * RS) * 32;
      const double* src = c.Rg + hpk(jn);
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int i = lane + 32 * q;
        if (q < nsl && i < jn) cp8(wcol + q * 32, src + i);
      }
    }
    cp_commit();
  }
  cp_wait<0>();
#pragma unroll
  for (int q = 0; q < CH; ++q) { const int i = lane + 32 * q; if (i < width) c.g[i] = gi[q]; }
  __syncwarp();
}

// ---- the same least-squares solve with a short dependent chain (QGD_QR_LEAN, the default) -----------------
// Profile of the ring version (profiles/r01_ncu_source_v8_fwd.txt): ~200 instructions and ~900 cycles per rotation,
// ~500 per back-substitution step -- 18 % of a sweep.  Here the rotation loop is split into phases of 32 rotations so
// that the pivot owner's register is indexed statically, the rows of H two and three rotations ahead are prefetched
// straight into registers, the special cases of givensAlgorithm fold into the sign of one reciprocal square root
// (1 / r_ii IS that reciprocal square root: no division), and lane 0's stores are predicated, not branched.
template <int S0, int CH>
__device__ __forceinline__ void qr_rotation_phase(double (&top)[CH], double* const (&colp)[CH], double* dinv,
                                                  double* g, double& gcur, int width, int nsl, int lane) {
  const int i0 = 32 * S0;
  if (i0 >= width) return;
  const int iend = min(width, i0 + 32);
#if !QGD_QR_ROWS4
  double loA[CH], loB[CH];
  auto fetch = [&](double (&lo)[CH], int r) {  // row r of the owned columns j = lane + 32 s >= r - 1
#pragma unroll
    for (int s = S0; s < CH; ++s) {
      const int j = lane + 32 * s;
      lo[s] = (s < nsl && j < width && r <= j + 1) ? hld(colp[s] + r) : 0.0;
    }
  };
#endif
  auto rotate = [&](int i, const double (&lo)[CH]) {
    const double a = __shfl_sync(FULL_MASK, top[S0], i & 31);
    const double b = __shfl_sync(FULL_MASK, lo[S0], i & 31);  // H[i+1][i], the row of the pivot owner's own column
    double rr = rsqrt(fma(a, a, b * b));
    // LinearAlgebra.givensAlgorithm: r = +-sqrt(a^2 + b^2), negative only when |a| > |b| and a < 0
    rr = (a < 0.0 && fabs(a) > fabs(b)) ? -rr : rr;
    const double cs = a * rr, sn = b * rr;
#pragma unroll
    for (int s = S0; s < CH; ++s) {
      const int j = lane + 32 * s;
      if (s < nsl && j < width && j >= i) {
        const double r = cs * top[s] + sn * lo[s];
        top[s] = -sn * top[s] + cs * lo[s];
        if (j > i) colp[s][i] = r;
      }
    }
    if (lane == 0) { dinv[i] = rr; g[i] = cs * gcur; }  // r_ii = cs a + sn b = 1 / rr
    gcur = -sn * gcur;
  };
#if QGD_QR_ROWS4
  // Rows in groups of four: the owned columns start 32-byte aligned, a lane fetches rows 4q..4q+3 of each of them with two
  // 16-byte loads (ONE sector) and the next two groups are in flight while a group is being rotated -- a quarter of the
  // scattered requests of the row-at-a-time version and 4 to 8 rotations of L2 latency hidden instead of 2.
  double A[CH][4], B[CH][4];
  auto fetch4 = [&](double (&buf)[CH][4], int gq) {  // rows 4 gq .. 4 gq + 3 of the owned columns j >= 4 gq - 1
#pragma unroll
    for (int s = S0; s < CH; ++s) {
      const int j = lane + 32 * s;
      if (s < nsl && j < width && 4 * gq <= j + 1) {
        const double2 lo2 = hld2(colp[s] + 4 * gq), hi2 = hld2(colp[s] + 4 * gq + 2);
        buf[s][0] = lo2.x; buf[s][1] = lo2.y; buf[s][2] = hi2.x; buf[s][3] = hi2.y;
      } else {
        buf[s][0] = buf[s][1] = buf[s][2] = buf[s][3] = 0.0;
      }
    }
  };
  auto rotate_group = [&](const double (&buf)[CH][4], int gq) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = 4 * gq + e - 1;  // rotation i uses row i + 1
      if (i >= i0 && i < iend) {
        double lo[CH];
#pragma unroll
        for (int s = 0; s < CH; ++s) lo[s] = buf[s][e];
        rotate(i, lo);
      }
    }
  };
  int gq = (i0 + 1) >> 2;
  fetch4(A, gq);
  fetch4(B, gq + 1);
  for (; 4 * gq <= iend; gq += 2) {
    rotate_group(A, gq);
    fetch4(A, gq + 2);
    rotate_group(B, gq + 1);
    fetch4(B, gq + 3);
  }
#else
  // two rows ahead, hand-unrolled.  Measured: prefetching 1 / 2 / 4 rows ahead gives 382 / 388 / 343
  // evals/s, a generic depth-D loop at D = 2 gives 377, L1-cached loads 382 -- the kernels are that sensitive to the size
  // and layout of the per-solve code.
  fetch(loA, i0 + 1);
  fetch(loB, i0 + 2);
  for (int i = i0; i < iend; i += 2) {
    rotate(i, loA);
    fetch(loA, i + 3);
    if (i + 1 < iend) {
      rotate(i + 1, loB);
      fetch(loB, i + 4);
    }
  }
#endif
}

// back substitution R y = g for the rows [32 S0, 32 S0 + 32) of the pivot: right-hand side rows i = lane + 32 q in
// registers, column j of R (contiguous in L2) prefetched two steps ahead
template <int S0, int CH>
__device__ __forceinline__ void qr_backsub_phase(double (&gi)[CH], const double* Rg, const double* dinv, int width, int lane) {
  const int j_lo = 32 * S0;
  if (j_lo >= width) return;
  const int j_hi = min(width, j_lo + 32) - 1;
  double cA[CH], cB[CH];
  auto fetch = [&](double (&col)[CH], int j) {  // rows i < j of column j
    const double* src = Rg + hpk(max(j, 0));
#pragma unroll
    for (int q = 0; q <= S0; ++q) {
      const int i = lane + 32 * q;
      col[q] = (j >= 0 && i < j) ? hld(src + i) : 0.0;
    }
  };
  auto step = [&](int j, const double (&col)[CH]) {
    const double yj = __shfl_sync(FULL_MASK, gi[S0], j & 31) * dinv[j];
#pragma unroll
    for (int q = 0; q <= S0; ++q) {
      const int i = lane + 32 * q;
      gi[q] = i < j ? fma(-yj, col[q], gi[q]) : (i == j ? yj : gi[q]);
    }
  };
  fetch(cA, j_hi);
  fetch(cB, j_hi - 1);
  for (int j = j_hi; j >= j_lo; j -= 2) {
    step(j, cA);
    fetch(cA, j - 2);
    if (j - 1 >= j_lo) {
      step(j - 1, cB);
      fetch(cB, j - 3);
    }
  }
}

// the phases in sequence (compile-time loops: CH = 4 for one-warp columns, 4 RS for row-split groups whose restart is 2N <= 128 RS)
template <int S0, int CH>
struct QrRotations {
  static __device__ __forceinline__ void run(double (&top)[CH], double* const (&colp)[CH], double* dinv, double* g, double& gcur, int width,
                                             int nsl, int lane) {
    qr_rotation_phase<S0, CH>(top, colp, dinv, g, gcur, width, nsl, lane);
    if constexpr (S0 + 1 < CH) QrRotations<S0 + 1, CH>::run(top, colp, dinv, g, gcur, width, nsl, lane);
  }
};
template <int S0, int CH>
struct QrBacksubs {
  static __device__ __forceinline__ void run(double (&gi)[CH], const double* Rg, const double* dinv, int width, int lane) {
    qr_backsub_phase<S0, CH>(gi, Rg, dinv, width, lane);
    if constexpr (S0 > 0) QrBacksubs<S0 - 1, CH>::run(gi, Rg, dinv, width, lane);
  }
};

template <class CTX>
__device__ __forceinline__ void qr_solve_lean(const CTX& c, int width, double beta) {
  const int lane = c.lane;
  constexpr int CH = 4 * CTX::kRS;  // width <= restart <= 128 RS
  const int nsl = (width + 31) >> 5;
  double* dinv = c.hcol;
  __syncwarp();
  double* colp[CH];
  double top[CH];
#pragma unroll
  for (int s = 0; s < CH; ++s) {
    const int j = lane + 32 * s;
    colp[s] = c.Rg + hpk(j);
    top[s] = j < width ? hld(colp[s]) : 0.0;
  }
  double gcur = beta;
  QrRotations<0, CH>::run(top, colp, dinv, c.g, gcur, width, nsl, lane);
  __threadfence_block();
  __syncwarp();  // R (L2), dinv and g (shared memory) of all lanes are visible
  double gi[CH];
#pragma unroll
  for (int q = 0; q < CH; ++q) { const int i = lane + 32 * q; gi[q] = i < width ? c.g[i] : 0.0; }
  QrBacksubs<CH - 1, CH>::run(gi, c.Rg, dinv, width, lane);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < CH; ++q) { const int i = lane + 32 * q; if (i < width) c.g[i] = gi[q]; }
  __syncwarp();
}

// GMRES with the blocked orthogonalisation; same interface and iteration semantics as gmres_fast_strict.
template <int EL, int NC, int VARIANT, bool TEAM, class OP, int RS>
__device__ int gmres_fast_blocked(const FastCtx<EL, RS>& c, const RegOps<EL, NC>& R, const OP& op, Vec<EL>& x, const Vec<EL>& b, double tol,
                                  int restart, int maxiter) {
  int pending = -1;  // TEAM: index of a basis vector handed to its owner warp with the next command
  constexpr int BLK = QGD_MGS_BLOCK > 1 ? QGD_MGS_BLOCK : 4;
  constexpr int CHK = 2 * EL * RS;  // k <= restart <= 2N <= 64 EL RS: CHK chunks of 32 rows cover a Hessenberg column
  const int lane = c.lane;
  // residual estimate beta / sqrt(accum) against tol, tested as beta^2 <= tol^2 accum (no square root on the
  // critical path of an iteration; the two tests differ only when the estimate is within an ulp of tol)
  double tol2 = tol * tol;
  double res_beta = 0.0, res_beta2 = 0.0, accum = 1.0;
  bool conv = false, first = true, start = true;
  int k = 1, it = 0;
  Vec<EL> v = x, w;
  // ONE inlined copy of the operator application serves the initial residual, the Arnoldi steps and the restarts
  // (three copies cost 20 KB of instruction cache in a kernel whose hot loop has to live in 32 KB)
#pragma unroll 1
  for (;;) {
    op.apply(v, w);
    if (start) {  // residual of the current iterate x (init!, init_residual!)
#pragma unroll
      for (int e = 0; e < EL; ++e) { v.u[e] = b.u[e] - w.u[e]; v.v[e] = b.v[e] - w.v[e]; }
      precond_fast<EL, NC>(R, v);
      const double beta2 = row_allsum(c, vdot_local<EL>(v, v));
      const double rbeta = rsqrt(beta2);
      vscale(v, rbeta);
      if constexpr (TEAM) basis_store_team<EL>(c, 0, v, pending);
      else basis_store<EL>(c, 0, v);
      accum = 1.0; res_beta = beta2 * rbeta; res_beta2 = beta2;
      if (first) conv = !(beta2 > tol2);  // at a restart residual.current keeps its value, as in the package
      first = false; start = false;
      k = 1;
      __syncwarp();
      if (lane == 0) c.nullv[0] = 1.0;
      __syncwarp();
      if (conv || it >= maxiter) break;
      continue;
    }
    precond_fast<EL, NC>(R, w);  // expand!
    if constexpr (TEAM) { __syncwarp(); gs_team_main<EL>(c, k, w, pending); pending = -1; __syncwarp(); }
    else if constexpr (QGD_GS_SUPER > 0 && BLK == 8) { __syncwarp(); gs_orthogonalize_super<EL, QGD_GS_SUPER>(c, k, w); }
    else gs_orthogonalize<EL, BLK, VARIANT>(c, k, w);
    if constexpr ((VARIANT & 4) != 0) __syncwarp();  // lane 0's coefficient stores (block_allsum8_dmma) before c.hcol is read
    // ||w||^2 and the null-vector recurrence <nullvec[0..k), H[0..k, k-1]> (update_residual!)
    double dpart = 0.0, hreg[CHK];
#pragma unroll
    for (int s = 0; s < CHK; ++s) {
      const int i = lane + 32 * s;
      const bool ok = i < k;
      hreg[s] = ok ? c.hcol[i] : 0.0;
      const double nl = ok ? c.nullv[i] : 0.0;
      dpart = fma(nl, hreg[s], dpart);
    }
    const double nrm2 = row_allsum(c, vdot_local<EL>(w, w));
    const double dsum = warp_allsum(dpart);
    // an exactly vanishing w (Krylov space exhausted, happy breakdown) gives H[k+1][k] = 0 as in the reference -- whose
    // residual estimate then drops to zero and ends the solve -- not 0 * inf
    const double rnrm = rsqrt(nrm2), nrm = nrm2 > 0.0 ? nrm2 * rnrm : 0.0;
    vscale(w, rnrm);
    if constexpr (TEAM) basis_store_team<EL>(c, k, w, pending);
    else basis_store<EL>(c, k, w);
    const double nv = -(dsum * rnrm);
    {  // Hessenberg column k-1 (rows 0..k) to the packed matrix in L2
      double* hc = c.Rg + hpk(k - 1);
#pragma unroll
      for (int s = 0; s < CHK; ++s) {
        const int i = lane + 32 * s;
        if (i < k) hc[i] = hreg[s];
      }
#if QGD_COMPACT_SMEM
      if (lane == 0) { hc[k] = nrm; c.nullv[k] = nv; }
#else
      if (lane == 0) { hc[k] = nrm; c.sub[k - 1] = nrm; c.nullv[k] = nv; }
#endif
    }
    accum = fma(nv, nv, accum);
    conv = !(res_beta2 > tol2 * accum);
    k += 1;
    it += 1;
    v = w;
    __syncwarp();
    const bool done = conv || it >= maxiter;
    if (k == restart + 1 || done) {  // x only at the end of the iterations and at a restart
      const int width = k - 1;
      if constexpr (QGD_QR_LEAN) qr_solve_lean(c, width, res_beta);
      else qr_solve_fast(c, width, res_beta);
      if constexpr (TEAM) {
        update_team_main<EL>(c, width, x, pending);
        pending = -1;
      } else {
        Vec<EL> vi;
        basis_load<EL>(c, 0, vi);
        for (int j = 0; j < width; ++j) {  // update_solution!: x += V[:, 0..width-1] y
          Vec<EL> vn;
          if (j + 1 < width) basis_load<EL>(c, j + 1, vn);
          vaxpy(x, c.g[j], vi);
          if (j + 1 < width) vi = vn;
        }
      }
      __syncwarp();
      if (done) break;
      v = x; start = true;  // restart from the residual of the updated iterate
    }
  }
  return it;
}

template <int EL, int NC, int VARIANT, bool STRICT, bool TEAM, class OP, int RS>
__device__ __forceinline__ int gmres_fast(const FastCtx<EL, RS>& c, const RegOps<EL, NC>& R, const OP& op, Vec<EL>& x, const Vec<EL>& b,
                                          double tol, int restart, int maxiter) {
  if constexpr (QGD_MGS_BLOCK > 1 && !STRICT) return gmres_fast_blocked<EL, NC, VARIANT, TEAM, OP>(c, R, op, x, b, tol, restart, maxiter);
  else return gmres_fast_strict<EL, NC, OP>(c, R, op, x, b, tol, restart, maxiter);
}

template <int EL, int M, int NC, int RS = 1>
struct FwdOpFast {  // LHSHolder (src/forward_evolution.jl:583-592)
  const FastCtx<EL, RS>& c; const RegOps<EL, NC>& R; const double* a_lhs;
  __device__ __forceinline__ void apply(const Vec<EL>& in, Vec<EL>& out) const {
    fwd_fast<EL, M, NC, false>(c, R, in, a_lhs, out, nullptr, nullptr, nullptr);
  }
};
template <int EL, int M, int NC, int RS = 1>
struct AdjOpFast {  // LHSHolderAdjoint (:624-633) through the reverse sweep
  const FastCtx<EL, RS>& c; const RegOps<EL, NC>& R; const double* a_lhs;
  __device__ __forceinline__ void apply(const Vec<EL>& in, Vec<EL>& out) const {
    double dK[M][NC], dS[M][NC];
    adj_fast<EL, M, NC, false>(c, R, in, a_lhs, out, nullptr, dK, dS);
  }
};

__device__ __forceinline__ size_t next_item(unsigned int* counter, int lane) {
  unsigned int v = 0;
  if (lane == 0) v = atomicAdd(counter, 1u);
  return (size_t)__shfl_sync(FULL_MASK, v, 0);
}

// ---- time-sliced work queue -----------------------------------------------------------------------------
// A column's time loop is strictly sequential, but nothing except one state vector (and, in the adjoint sweep, the
// gradient partial) is carried from one time step to the next.  The sweeps therefore hand out TICKETS of
// `seg_steps` time steps: ticket t = (segment t / items, item t % items).  A warp that draws a ticket waits until
// the previous segment of that item has been published (it was drawn earlier, so it is finished or running on a
// resident warp: no deadlock), continues the item for seg_steps steps and publishes it again.  Columns thereby
// migrate between warps and the idle tail at the end of a sweep shrinks from one whole column to one segment.
template <int EL>
__device__ __forceinline__ void vload_cg(Vec<EL>& a, const double* p, int N, int lane) {
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    const int r = lane + 32 * e;
    const bool ok = r < N;
    a.u[e] = ok ? __ldcg(p + r) : 0.0;
    a.v[e] = ok ? __ldcg(p + N + r) : 0.0;
  }
}
__device__ __forceinline__ void wait_segment(const int* progress, int seg, int lane, int* err) {
  if (lane == 0) {
    const volatile int* f = progress;
    unsigned spins = 0;
    // The ticket order rules out a deadlock (the previous segment was drawn earlier, so it is finished or running on
    // a resident warp); the bound (about half a minute) only keeps a logic error or a wedged device from hanging the
    // process: the sweep then continues with an unpublished state and raises the error word, which every host entry
    // point checks after the sweeps (QGD_ESTATE) -- wrong numbers never leave the library silently.
    while (*f < seg && ++spins < (1u << 27)) __nanosleep(256);
    if (*f < seg) atomicExch(err, 1);
  }
  __syncwarp();
  __threadfence();
}
__device__ __forceinline__ void publish_segment(int* progress, int seg, int lane) {
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicExch(progress, seg + 1);
}

// the same for a row-split group: the leader draws / publishes, the group's barrier orders the slices' stores before it
template <int EL, int RS>
__device__ __forceinline__ size_t next_item_rs(const FastCtx<EL, RS>& c, unsigned int* counter) {
  if constexpr (RS == 1) return next_item(counter, c.lane);
  else {
    if (c.slice == 0 && c.lane == 0) *c.gtick = atomicAdd(counter, 1u);
    gsync(c);
    const unsigned int v = *reinterpret_cast<volatile unsigned int*>(c.gtick);
    gsync(c);  // everyone has read the word before the leader can draw again
    return (size_t)v;
  }
}
template <int EL, int RS>
__device__ __forceinline__ void publish_segment_rs(const FastCtx<EL, RS>& c, int* progress, int seg) {
  if constexpr (RS == 1) publish_segment(progress, seg, c.lane);
  else {
    __threadfence();
    gsync(c);
    if (c.slice == 0 && c.lane == 0) atomicExch(progress, seg + 1);
  }
}

// control Taylor coefficients of a time level: global [2][M+1][NC] -> shared (p, q) pairs [M+1][NC]
template <int EL, int M, int NC, int RS>
__device__ __forceinline__ void load_cv_fast(const FastCtx<EL, RS>& c, const double* src) {
  __syncwarp();
  for (int i = c.lane; i < (M + 1) * NC; i += 32) c.cv[i] = make_double2(src[i], src[(M + 1) * NC + i]);
  __syncwarp();
}

// shared-memory carve-up: [16 bytes: TMEM address slot][warp regions]; each region = fixed part + KS basis
// vectors (+ extra doubles)
// doubles of shared memory per row-split group: gather buffers [2][32 EL RS] double2, partial sums [2][RS][8], ticket word
template <int EL, int RS>
__host__ __device__ constexpr int group_doubles() { return RS == 1 ? 0 : 2 * 2 * 32 * EL * RS + 2 * RS * 8 + 2; }

template <int EL, int M, int NC, bool STRICT = false, bool TEAM = false, int RS = 1>
__device__ __forceinline__ FastCtx<EL, RS> make_fast_ctx(const QgdDevProb& d, const SweepArgs& a, unsigned char* smem, double** extra,
                                                         uint32_t tmem_base) {
  static_assert(RS == 1 || !TEAM, "a row-split group is not a latency team");
  FastCtx<EL, RS> c;
  c.lane = threadIdx.x & 31;
  c.slice = RS == 1 ? 0 : (int)((threadIdx.x >> 5) % RS);
  c.bar = 1 + (int)((threadIdx.x >> 5) / RS);
  c.vl = c.lane + 32 * EL * c.slice;
  c.par = 0u;
  c.gx = nullptr; c.gred = nullptr; c.gtick = nullptr; c.stage = nullptr;
  if constexpr (RS > 1) {  // group regions behind the warp regions
    double* gbase = reinterpret_cast<double*>(smem + 16) + (size_t)(blockDim.x >> 5) * a.warp_smem_doubles +
                    (size_t)((threadIdx.x >> 5) / RS) * group_doubles<EL, RS>();
    c.gx = reinterpret_cast<double2*>(gbase);
    c.gred = gbase + 2 * 2 * 32 * EL * RS;
    c.gtick = reinterpret_cast<unsigned*>(c.gred + 2 * RS * 8);
  }
  c.N = d.N; c.N2 = d.N2; c.KS = a.ks;
  // TEAM: the four warps of the CTA serve ONE column: they share the (single) per-warp region; each owns a TMEM lane quarter
  const int warp = TEAM ? 0 : (int)(threadIdx.x >> 5);
  // TMEM: warp w may address lanes [32 (w % 4), +32); warps w and w + 4 split the 512 columns
  const int groups = ((blockDim.x >> 5) + 3) >> 2;
  const int cols = a.tmem_cols / groups;
  c.KT = a.kt;
  c.tm = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(cols * (warp >> 2));
  if constexpr (TEAM) c.tm = tmem_base + ((uint32_t)(32 * (threadIdx.x >> 5)) << 16);
  double* w = reinterpret_cast<double*>(smem + 16) + (size_t)warp * a.warp_smem_doubles;
  c.xs = reinterpret_cast<double2*>(w); w += FastCtx<EL, RS>::kRingDoubles;
  c.cv = reinterpret_cast<double2*>(w); w += 2 * (M + 1) * NC;
#if QGD_COMPACT_SMEM
  static_assert(FastCtx<EL, RS>::kRingDoubles >= 64 * EL * RS + 2, "g (2N + 2 doubles) aliases the gather buffer");
  if constexpr (STRICT) {
    c.rot = reinterpret_cast<double2*>(w); c.hcol = w; c.sub = w + d.N2 + 10; w += 2 * (d.N2 + 2 + 8);
    c.nullv = w; w += d.N2 + 2;
    c.g = w; w += d.N2 + 2;
  } else {
    c.rot = nullptr; c.hcol = w; c.sub = nullptr; w += d.N2 + 2 + 8;
    c.nullv = w; w += d.N2 + 2;
    c.g = reinterpret_cast<double*>(c.xs);
  }
#else
  c.rot = reinterpret_cast<double2*>(w); c.hcol = w; c.sub = w + d.N2 + 10; w += 2 * (d.N2 + 2 + 8);
  c.nullv = w; w += d.N2 + 2;
  c.g = w; w += d.N2 + 2;
#endif
  c.Vs = reinterpret_cast<double2*>(w); w += (size_t)a.ks * 2 * 32 * EL;
  if constexpr (RS > 1) {
    if (a.stage_l2) { c.stage = reinterpret_cast<double2*>(w); w += (size_t)8 * 2 * 32 * EL; }
  }
  double* hs = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(w) + 31) & ~(uintptr_t)31);  // 32-byte aligned columns (hpk)
  w += a.h_smem_doubles;
  c.team = nullptr;
  if constexpr (TEAM) { c.team = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(w) + 15) & ~(uintptr_t)15); w = c.team + team_doubles<EL>(); }
  *extra = w;
  const size_t slot = TEAM ? (size_t)blockIdx.x : (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
  c.Vg = reinterpret_cast<double2*>(a.Vws + slot * a.v_stride);
  c.Rg = a.h_smem_doubles ? hs : a.Hws + slot * a.h_stride;
  return c;
}

// ------------------------------------------------------------------------------------------------------
// FORCED: forced forward solves -- eval_forward!(...; forcing) with an explicit forcing array (a.forcing_in), or the P x ncol
// forced solves of eval_grad_forced (a.base_history: item b is control parameter b, zero initial state, control vector 0,
// the guard-penalty derivative accumulated on the way, no history written).
// TEAM: the latency team above -- a CTA of four warps per column; warps 1-3 only serve the orthogonalisation.
template <int EL, int M, int NC, bool STRICT, bool FORCED = false, bool TEAM = false, int RS = 1>
__global__ void __launch_bounds__(32 * QGD_WARPS_PER_CTA, 1) k_forward_fast(const __grid_constant__ QgdDevProb d, const __grid_constant__ SweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* extra;
  const uint32_t tmem_base = a.tmem_cols ? tmem_alloc_cols(reinterpret_cast<uint32_t*>(smem), (uint32_t)a.tmem_cols) : 0u;
  const FastCtx<EL, RS> c = make_fast_ctx<EL, M, NC, STRICT, TEAM, RS>(d, a, smem, &extra, tmem_base);
  if constexpr (TEAM) {
    if ((threadIdx.x >> 5) > 0) {
      team_helper_loop<EL>(c.team, c.hcol, c.g, c.tm, (int)(threadIdx.x >> 5), c.lane);
      tmem_free_cols(tmem_base, (uint32_t)a.tmem_cols);
      return;
    }
  }
  const int lane = c.lane, vl = c.vl;  // vl: row space (lane + 32 EL slice)
  const int N = d.N, N2 = d.N2;
  RegOps<EL, NC> R;
  load_regops<EL, NC>(R, d, vl, 0);
  double a_rhs[M + 1], a_lhs[M + 1], a_tay[M + 1];
#pragma unroll
  for (int j = 0; j <= M; ++j) { a_rhs[j] = d.a_rhs[j]; a_lhs[j] = d.a_lhs[j]; a_tay[j] = d.a_tay[j]; }
  const size_t items = (size_t)a.B * d.ncol;
  const size_t cv_stride = (size_t)2 * (M + 1) * NC;
  const size_t slot_sz = (size_t)N2 * (M + 1);
  const FwdOpFast<EL, M, NC, RS> op{c, R, a_lhs};
  // columns need different numbers of GMRES iterations: warps draw (segment, control vector, column) tickets
  const int S = a.seg_steps, nseg = (d.nsteps + S - 1) / S;
  const size_t tickets = items * (size_t)nseg;
  for (size_t ticket = next_item_rs(c, a.work_counter); ticket < tickets; ticket = next_item_rs(c, a.work_counter)) {
    const int seg = (int)(ticket / items);
    const size_t item = ticket % items;
    const int b = (int)(item / d.ncol), cl = (int)(item % d.ncol), col = d.col0 + cl;
    const bool gradf = FORCED && a.base_history != nullptr;
    const double* cvb = a.cvals + (gradf ? (size_t)0 : (size_t)b * (d.nsteps + 1) * cv_stride);
    double* hist = a.history ? a.history + slot_sz * a.nslots * ((size_t)cl + (size_t)d.ncol * b) : nullptr;
    double* carry = a.final_state + (size_t)N2 * ((size_t)cl + (size_t)d.ncol * b);  // state between segments; w_N at the end
    const int n0 = seg * S, n1 = min(n0 + S, d.nsteps);
    const bool last = seg == nseg - 1;
    // forcing of time level n (FORCED kernels only)
    const double* fin = (FORCED && a.forcing_in) ? a.forcing_in + (size_t)N2 * M * (d.nsteps + 1) * ((size_t)cl + (size_t)d.ncol * b) : nullptr;
    const double* hb = gradf ? a.base_history + slot_sz * (d.nsteps + 1) * (size_t)cl : nullptr;
    const size_t tab_stride = (size_t)2 * (M + 1) * d.P;
    auto forcing_at = [&](int n) {
      FastForcing F;
      F.arr = fin ? fin + (size_t)N2 * M * n : nullptr;
      F.hb = hb ? hb + slot_sz * n : nullptr;
      F.tp = gradf ? d.table + tab_stride * n + b : nullptr;
      F.tq = gradf ? F.tp + (size_t)(M + 1) * d.P : nullptr;
      F.P = d.P; F.kop = gradf ? a.theta_op[b] - 1 : 0;  // theta_op holds blob indices: operator 0 is the drift
      return F;
    };
    // eval_grad_forced: d/dtheta of the guard penalty, dt/tf sum_n wt_n (<dpsi_n, W psi_n> + <psi_n, W dpsi_n>) with the
    // diagonal W of this kernel family = 2 wt_n sum_r W_rr dpsi_r psi_r (src/eval_grad_forced.jl:150-172); per-lane partial
    double gpen = 0.0;
    auto guard_cross = [&](const Vec<EL>& dx, int n) {
      Vec<EL> w0;
      vload_cg(w0, hb + slot_sz * n, N, vl);
      const double wt = (n == 0 || n == d.nsteps) ? 1.0 : 2.0;
#pragma unroll
      for (int e = 0; e < EL; ++e) gpen = fma(wt * R.wu[e] * dx.u[e], w0.u[e], fma(wt * R.wv[e] * dx.v[e], w0.v[e], gpen));
    };
    Vec<EL> x;
    if (seg == 0) {
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        const int r = vl + 32 * e;
        x.u[e] = (r < N && !gradf) ? d.u0[r + (size_t)N * col] : 0.0;
        x.v[e] = (r < N && !gradf) ? d.v0[r + (size_t)N * col] : 0.0;
      }
    } else {
      wait_segment(a.progress + item, seg, lane, a.err);
      vload_cg(x, carry, N, vl);
    }
    load_cv_fast<EL, M, NC>(c, cvb + (size_t)n0 * cv_stride);
    const int nend = last ? n1 : n1 - 1;  // the last segment also forms the Taylor columns at the final time (forward_evolution.jl:232-242)
    for (int n = n0; n <= nend; ++n) {
      Vec<EL> rhs, guess;
      double* slot = (hist && n % a.save_every == 0) ? hist + slot_sz * (n / a.save_every) : nullptr;
      if constexpr (FORCED) {
        if (gradf) guard_cross(x, n);
        if (n == d.nsteps) {  // Taylor columns at the final time, WITHOUT forcing as in the reference (forward_evolution.jl:232-242)
          fwd_fast<EL, M, NC, true>(c, R, x, a_rhs, rhs, a_tay, &guess, slot);
          break;
        }
        const FastForcing F = forcing_at(n);
        fwd_fast<EL, M, NC, true, true>(c, R, x, a_rhs, rhs, a_tay, &guess, slot, &F);  // explicit part at t_n
      } else {
        fwd_fast<EL, M, NC, true>(c, R, x, a_rhs, rhs, a_tay, &guess, slot);      // explicit part at t_n
        if (n == d.nsteps) break;
      }
      load_cv_fast<EL, M, NC>(c, cvb + (size_t)(n + 1) * cv_stride);              // implicit part uses t_{n+1}
      if constexpr (FORCED) {  // the forcing of t_{n+1} is explicit: its implicit-side combination moves to the right-hand side
        const FastForcing F1 = forcing_at(n + 1);                                 // (forward_evolution.jl:196-206)
        Vec<EL> zero, fh;
        vzero(zero);
        fwd_fast<EL, M, NC, false, true>(c, R, zero, a_lhs, fh, nullptr, nullptr, nullptr, &F1);
        vaxpy(rhs, -1.0, fh);
      }
      x = guess;
      const int it = gmres_fast<EL, NC, QGD_FWD_VARIANT, STRICT, TEAM>(c, R, op, x, rhs, d.abstol, N2, N2);
      if (a.iters && lane == 0) a.iters[(size_t)n + (size_t)d.nsteps * ((size_t)cl + (size_t)d.ncol * b)] = it;
    }
    vstore(x, carry, N, vl);
    if constexpr (FORCED) {
      if (gradf) {  // guard-penalty derivative partial of this segment; the sum is carried in guardcol across segments
        const double g = row_allsum(c, gpen) * (d.dt / d.tf);
        double* gc = a.guardcol + ((size_t)cl + (size_t)d.ncol * b);
        if (lane == 0 && c.slice == 0) *gc = (seg == 0 ? 0.0 : __ldcg(gc)) + g;
      }
    }
    publish_segment_rs(c, a.progress + item, seg);
  }
  if constexpr (TEAM) {  // release the helper warps
    if (lane == 0) team_view<EL>(c.team).cmd[0] = TEAM_CMD_EXIT;
    team_bar();
  }
  if (a.tmem_cols) tmem_free_cols(tmem_base, (uint32_t)a.tmem_cols);
}

// One side of an adjoint step as a real (not inlined) function: out = (sum_j sign alpha_j W_j(t))^T lam together with
// the reduced inner products g^K, g^S [M][NC] of that time level (left in gKs / gSs, shared memory).  It runs twice per
// time step; inlining it twice costs 21 KB of instruction cache in a kernel whose GMRES loop has to stay resident, and
// rolling the two calls into a loop spills registers into that loop.
template <int EL, int M, int NC, int RS>
__device__ __noinline__ void grad_side_fast(const FastCtx<EL, RS>& c, const RegOps<EL, NC>& R, const Vec<EL>& lam, const double* alpha_src,
                                            double sign, Vec<EL>& out, const double* hist, bool last_use, double* gKs, double* gSs) {
  double gK[M][NC], gS[M][NC], alpha[M + 1];
#pragma unroll
  for (int j = 0; j <= M; ++j) alpha[j] = sign * alpha_src[j];
#pragma unroll
  for (int r = 0; r < M; ++r)
#pragma unroll
    for (int k = 0; k < NC; ++k) { gK[r][k] = 0.0; gS[r][k] = 0.0; }
  adj_fast<EL, M, NC, true>(c, R, lam, alpha, out, hist, gK, gS, last_use);
#pragma unroll
  for (int r = 0; r < M; ++r)
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const double sK = warp_allsum(gK[r][k]), sS = warp_allsum(gS[r][k]);
      if (c.lane == 0) { gKs[r * NC + k] = sK; gSs[r * NC + k] = sS; }
    }
  __syncwarp();
}

// grad_acc[theta] -= sum_r table_p[r][theta] gK[r][k(theta)] + table_q[r][theta] gS[r][k(theta)]
template <int M, int NC>
__device__ __forceinline__ void accumulate_grad_fast(int t0, int tstep, int P, const QgdDevControl* ctrls, const double* table_n,
                                                     const double* gKs, const double* gSs, double* gacc) {
  const int nd = M + 1;
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int off = ctrls[k].offset, nco = ctrls[k].ncoeff;
    for (int t = t0; t < nco; t += tstep) {  // a row-split group deals the parameters over its warps
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < M; ++r) {
        s = fma(table_n[((size_t)0 * nd + r) * P + off + t], gKs[r * NC + k], s);
        s = fma(table_n[((size_t)1 * nd + r) * P + off + t], gSs[r * NC + k], s);
      }
#if QGD_COMPACT_SMEM
      // the accumulator is the L2-resident gradient partial of the column (it migrates between SMs with the tickets).
      // One lane of one warp owns an entry at any time, so the reduction instruction is deterministic; unlike a
      // load-subtract-store it does not wait for the L2 round trip.
      atomicAdd(gacc + off + t, -s);
#else
      gacc[off + t] -= s;
#endif
    }
  }
}

template <int EL, int M, int NC, bool STRICT, bool TEAM = false, int RS = 1>
__global__ void __launch_bounds__(32 * QGD_WARPS_PER_CTA, 1) k_backward_fast(const __grid_constant__ QgdDevProb d, const __grid_constant__ SweepArgs a,
                                                                               const QgdDevControl* __restrict__ ctrls) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* extra;
  const uint32_t tmem_base = a.tmem_cols ? tmem_alloc_cols(reinterpret_cast<uint32_t*>(smem), (uint32_t)a.tmem_cols) : 0u;
  const FastCtx<EL, RS> c = make_fast_ctx<EL, M, NC, STRICT, TEAM, RS>(d, a, smem, &extra, tmem_base);
  if constexpr (TEAM) {
    if ((threadIdx.x >> 5) > 0) {
      team_helper_loop<EL>(c.team, c.hcol, c.g, c.tm, (int)(threadIdx.x >> 5), c.lane);
      tmem_free_cols(tmem_base, (uint32_t)a.tmem_cols);
      return;
    }
  }
  const int lane = c.lane, vl = c.vl;  // vl: row space (lane + 32 EL slice)
  const int N = d.N, N2 = d.N2, Nt = d.nsteps + 1, P = d.P;
  RegOps<EL, NC> R0;
  load_regops<EL, NC>(R0, d, vl, 1);
  const RegOps<EL, NC> R = R0;  // const object: the non-inlined callee cannot legally change it
  double a_rhs[M + 1], a_lhs[M + 1], a_imp[M + 1];
#pragma unroll
  for (int j = 0; j <= M; ++j) { a_rhs[j] = d.a_rhs[j]; a_lhs[j] = d.a_lhs[j]; a_imp[j] = -d.a_lhs[j]; }
  const size_t items = (size_t)a.B * d.ncol;
  const size_t cv_stride = (size_t)2 * (M + 1) * NC;
  const size_t slot_sz = (size_t)N2 * (M + 1);
  const size_t tab_stride = (size_t)2 * (M + 1) * P;
#if QGD_COMPACT_SMEM
  // reduced g^K, g^S [M][NC] each: in the gather buffer (idle between the gradient sweep and the contraction with the
  // control basis table); the gradient partial itself accumulates in L2 (gcol), touched once per time step
  static_assert(FastCtx<EL, RS>::kRingDoubles >= 2 * M * NC, "gKs / gSs alias the gather buffer");
  double* gKs = reinterpret_cast<double*>(c.xs);
  double* gSs = gKs + M * NC;
  (void)extra;
#else
  double* gacc = extra;          // [P]
  double* gKs = gacc + P;        // [M][NC] reduced g^K
  double* gSs = gKs + M * NC;    // [M][NC] reduced g^S
#endif
  const AdjOpFast<EL, M, NC, RS> op{c, R, a_lhs};
  const double fsc = -2.0 * d.dt / d.tf;
  const int S = a.seg_steps, nseg = (d.nsteps + S - 1) / S;
  const size_t tickets = items * (size_t)nseg;
  for (size_t ticket = next_item_rs(c, a.work_counter); ticket < tickets; ticket = next_item_rs(c, a.work_counter)) {
    const int seg = (int)(ticket / items);
    const size_t item = ticket % items;
    const int b = (int)(item / d.ncol), cl = (int)(item % d.ncol), col = d.col0 + cl;
    const double* cvb = a.cvals + (size_t)b * Nt * cv_stride;
    const double* hist = a.history + slot_sz * Nt * ((size_t)cl + (size_t)d.ncol * b);
    double* lam0 = a.lambda0 ? a.lambda0 + (size_t)N2 * Nt * ((size_t)cl + (size_t)d.ncol * b) : nullptr;
    double* gcol = a.gradcol + (size_t)P * ((size_t)cl + (size_t)d.ncol * b);  // gradient partial, also carried between segments
    double* carry = a.carry + (size_t)N2 * ((size_t)cl + (size_t)d.ncol * b);   // lambda between segments
    const int n_hi = d.nsteps - 1 - seg * S, n_lo = max(n_hi - S + 1, 0);
    Vec<EL> lam;
#if QGD_COMPACT_SMEM
    double* gacc = gcol;
#endif
    if (seg == 0) {
      for (int t = lane; t < P; t += 32) gacc[t] = 0.0;
      vload(lam, a.terminal + (size_t)N2 * ((size_t)col + (size_t)d.nic * b), N, vl);
      if (lam0) vstore(lam, lam0 + (size_t)N2 * d.nsteps, N, vl);
    } else {
      wait_segment(a.progress + item, seg, lane, a.err);
#if !QGD_COMPACT_SMEM
      for (int t = lane; t < P; t += 32) gacc[t] = __ldcg(gcol + t);
#endif
      vload_cg(lam, carry, N, vl);
    }
    load_cv_fast<EL, M, NC>(c, cvb + (size_t)(n_hi + 1) * cv_stride);
    for (int n = n_hi; n >= n_lo; --n) {
      Vec<EL> w0, rhs;
#if QGD_BWD_MERGE_SIDES == 2
      // implicit side: time level n+1 (its control values are the ones currently loaded), coefficients -a_lhs
      grad_side_fast<EL, M, NC>(c, R, lam, d.a_lhs, -1.0, w0, hist + slot_sz * (n + 1), true, gKs, gSs);
      accumulate_grad_fast<M, NC>(lane + 32 * c.slice, 32 * RS, P, ctrls, d.table + (size_t)(n + 1) * tab_stride, gKs, gSs, gacc);
      __syncwarp();
      // explicit side: time level n, coefficients a_rhs; rhs = R(t_n)^T lambda_{n+1}
      load_cv_fast<EL, M, NC>(c, cvb + (size_t)n * cv_stride);
      grad_side_fast<EL, M, NC>(c, R, lam, d.a_rhs, 1.0, rhs, hist + slot_sz * n, false, gKs, gSs);
      accumulate_grad_fast<M, NC>(lane + 32 * c.slice, 32 * RS, P, ctrls, d.table + (size_t)n * tab_stride, gKs, gSs, gacc);
      __syncwarp();
#elif !QGD_BWD_MERGE_SIDES
      {
      double gK[M][NC], gS[M][NC];
      // ---- implicit side: time level n+1 (its control values are the ones currently loaded)
#pragma unroll
      for (int r = 0; r < M; ++r)
#pragma unroll
        for (int k = 0; k < NC; ++k) { gK[r][k] = 0.0; gS[r][k] = 0.0; }
      adj_fast<EL, M, NC, true>(c, R, lam, a_imp, w0, hist + slot_sz * (n + 1), gK, gS);
#pragma unroll
      for (int r = 0; r < M; ++r)
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const double sK = row_allsum(c, gK[r][k]), sS = row_allsum(c, gS[r][k]);
          if (lane == 0) { gKs[r * NC + k] = sK; gSs[r * NC + k] = sS; }
        }
      __syncwarp();
      accumulate_grad_fast<M, NC>(lane + 32 * c.slice, 32 * RS, P, ctrls, d.table + (size_t)(n + 1) * tab_stride, gKs, gSs, gacc);
      __syncwarp();
      // ---- explicit side: time level n
      load_cv_fast<EL, M, NC>(c, cvb + (size_t)n * cv_stride);
#pragma unroll
      for (int r = 0; r < M; ++r)
#pragma unroll
        for (int k = 0; k < NC; ++k) { gK[r][k] = 0.0; gS[r][k] = 0.0; }
      adj_fast<EL, M, NC, true>(c, R, lam, a_rhs, rhs, hist + slot_sz * n, gK, gS, false);  // rhs = R(t_n)^T lambda_{n+1}
#pragma unroll
      for (int r = 0; r < M; ++r)
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const double sK = row_allsum(c, gK[r][k]), sS = row_allsum(c, gS[r][k]);
          if (lane == 0) { gKs[r * NC + k] = sK; gSs[r * NC + k] = sS; }
        }
      __syncwarp();
      accumulate_grad_fast<M, NC>(lane + 32 * c.slice, 32 * RS, P, ctrls, d.table + (size_t)n * tab_stride, gKs, gSs, gacc);
      __syncwarp();
      }
#else
      // side 0: implicit side, time level n+1 (its control values are the ones currently loaded), -LHS coefficients;
      // side 1: explicit side, time level n, RHS coefficients; rhs = R(t_n)^T lambda_{n+1}.  One rolled loop: a second
      // inlined copy of the gradient sweep would cost 21 KB of instruction cache.
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        double gK[M][NC], gS[M][NC], alpha[M + 1];
#pragma unroll
        for (int j = 0; j <= M; ++j) alpha[j] = side ? a_rhs[j] : a_imp[j];
#pragma unroll
        for (int r = 0; r < M; ++r)
#pragma unroll
          for (int k = 0; k < NC; ++k) { gK[r][k] = 0.0; gS[r][k] = 0.0; }
        const int lvl = n + 1 - side;
        if (side) load_cv_fast<EL, M, NC>(c, cvb + (size_t)n * cv_stride);
        adj_fast<EL, M, NC, true>(c, R, lam, alpha, rhs, hist + slot_sz * lvl, gK, gS);
#pragma unroll
        for (int r = 0; r < M; ++r)
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            const double sK = row_allsum(c, gK[r][k]), sS = row_allsum(c, gS[r][k]);
            if (lane == 0) { gKs[r * NC + k] = sK; gSs[r * NC + k] = sS; }
          }
        __syncwarp();
        accumulate_grad_fast<M, NC>(lane + 32 * c.slice, 32 * RS, P, ctrls, d.table + (size_t)lvl * tab_stride, gKs, gSs, gacc);
        __syncwarp();
      }
#endif
      if (n >= 1) {
        // guard forcing f_n = -(2 dt/tf) W w_n (interior point: trapezoid weight 1), W diagonal
        vload_cg(w0, hist + slot_sz * n, N, vl);
#pragma unroll
        for (int e = 0; e < EL; ++e) {
          rhs.u[e] = fma(fsc, R.wu[e] * w0.u[e], rhs.u[e]);
          rhs.v[e] = fma(fsc, R.wv[e] * w0.v[e], rhs.v[e]);
        }
        // x0 = lambda_{n+1} (forward_evolution.jl:450)
        const int it = gmres_fast<EL, NC, QGD_BWD_VARIANT, STRICT, TEAM>(c, R, op, lam, rhs, d.abstol, N2, N2);
        if (lam0) vstore(lam, lam0 + (size_t)N2 * n, N, vl);
        if (a.iters && lane == 0) a.iters[(size_t)n + (size_t)d.nsteps * ((size_t)cl + (size_t)d.ncol * b)] = it;
      }
    }
    __syncwarp();
#if !QGD_COMPACT_SMEM
    for (int t = lane; t < P; t += 32) gcol[t] = gacc[t];
#endif
    if (seg < nseg - 1) vstore(lam, carry, N, vl);
    publish_segment_rs(c, a.progress + item, seg);
  }
  if constexpr (TEAM) {  // release the helper warps
    if (lane == 0) team_view<EL>(c.team).cmd[0] = TEAM_CMD_EXIT;
    team_bar();
  }
  if (a.tmem_cols) tmem_free_cols(tmem_base, (uint32_t)a.tmem_cols);
}

// Infidelity + terminal condition (src/infidelity.jl:7-18, src/eval_grad_discrete_adjoint.jl:1-67) on the register operators:
// one warp per control vector; <psi_N, R>, <psi_N, T> over ALL columns, then per column the un-preconditioned GMRES(20) solve
//   LHS(tf)^T lambda_N = (2 / N_ess^2) (<psi,R> R + <psi,T> T) - (dt / tf) W psi_N
// with the solution of column i-1 as the initial guess of column i, as the reference carries it.  (The generic k_terminal does
// the same through the row-ELL operators; at one evaluation it took 9 % of the call.)  The solve uses the STRICT Gram-Schmidt:
// at 64 levels and tolerances of 1e-12 and below it runs into its 2N-iteration cap un-converged (in the reference too), and
// the iterate it stops at then moves by 1e-5 relative under the blocked orthogonalisation's different rounding (measured,
// DESIGN 9.2) -- which would shift every adjoint iteration count downstream.  It is one short solve per column.
template <int EL, int M, int NC, int RS = 1>
__global__ void __launch_bounds__(32 * QGD_WARPS_PER_CTA, 1) k_terminal_fast(const __grid_constant__ QgdDevProb d, const __grid_constant__ SweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* extra;
  const uint32_t tmem_base = a.tmem_cols ? tmem_alloc_cols(reinterpret_cast<uint32_t*>(smem), (uint32_t)a.tmem_cols) : 0u;
  const FastCtx<EL, RS> c = make_fast_ctx<EL, M, NC, true, false, RS>(d, a, smem, &extra, tmem_base);
  const int gpc = (int)(blockDim.x >> 5) / RS, grp = (int)(threadIdx.x >> 5) / RS, lane = c.lane, vl = c.vl;  // groups per CTA (RS = 1: warps)
  const int N = d.N, N2 = d.N2;
  RegOps<EL, NC> R;
  load_regops<EL, NC>(R, d, vl, -1);  // no preconditioner in the terminal solves
  double a_lhs[M + 1];
#pragma unroll
  for (int j = 0; j <= M; ++j) a_lhs[j] = d.a_lhs[j];
  const AdjOpFast<EL, M, NC, RS> op{c, R, a_lhs};
  const size_t cv_stride = (size_t)2 * (M + 1) * NC;
  const double ness2 = (double)d.Ness * (double)d.Ness, sc = 2.0 / ness2;
  const double fsc = -2.0 * d.dt / d.tf * 0.5;  // forcing[:, end, :]: trapezoid weight 1/2
  const int restart = N2 < 20 ? N2 : 20;
  for (int b = blockIdx.x * gpc + grp; b < a.B; b += gridDim.x * gpc) {
    const double* psi = a.final_all + (size_t)N2 * d.nic * b;
    double dR = 0.0, dT = 0.0;
    int tc0 = 0, tc1 = d.nic;
    if (a.dots_in) {
      dR = a.dots_in[2 * b]; dT = a.dots_in[2 * b + 1];
      tc0 = a.term_col0; tc1 = a.term_col0 + a.term_ncol;
    } else {
      for (int col = 0; col < d.nic; ++col) {
        Vec<EL> p, Rt;
        vload(p, psi + (size_t)N2 * col, N, vl);
        vload(Rt, a.target + (size_t)N2 * col, N, vl);
#pragma unroll
        for (int e = 0; e < EL; ++e) {
          dR += p.u[e] * Rt.u[e] + p.v[e] * Rt.v[e];
          dT += p.u[e] * Rt.v[e] - p.v[e] * Rt.u[e];  // T = [R_v; -R_u]
        }
      }
      if constexpr (RS == 1) { dR = warp_sum(dR); dT = warp_sum(dT); }
      else { dR = row_allsum(c, dR); dT = row_allsum(c, dT); }
    }
    if (lane == 0) a.infidelity[b] = 1.0 - (dR * dR + dT * dT) / ness2;
    load_cv_fast<EL, M, NC>(c, a.cvals + ((size_t)b * (d.nsteps + 1) + d.nsteps) * cv_stride);  // controls at t = tf
    Vec<EL> x;
    vzero(x);
    for (int col = tc0; col < tc1; ++col) {
      Vec<EL> p, Rt, rhs;
      vload(p, psi + (size_t)N2 * col, N, vl);
      vload(Rt, a.target + (size_t)N2 * col, N, vl);
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        rhs.u[e] = (dR * Rt.u[e] + dT * Rt.v[e]) * sc + fsc * (R.wu[e] * p.u[e]);
        rhs.v[e] = (dR * Rt.v[e] + dT * (-Rt.u[e])) * sc + fsc * (R.wv[e] * p.v[e]);
      }
      const int it = gmres_fast_strict<EL, NC>(c, R, op, x, rhs, d.abstol, restart, N2, d.reltol);
      vstore(x, a.terminal_out + (size_t)N2 * ((size_t)col + (size_t)d.nic * b), N, vl);
      if (a.iters_term && lane == 0) a.iters_term[(size_t)col + (size_t)d.nic * b] = it;
    }
  }
  if (a.tmem_cols) tmem_free_cols(tmem_base, (uint32_t)a.tmem_cols);
}

// K2 on its own (tests): uv [2N][1+m][ncols]; forward Taylor columns or Lambda_j = W_j^T x.
template <int EL, int M, int NC>
__global__ void __launch_bounds__(32 * QGD_WARPS_PER_CTA, 1) k_derivs_fast(const __grid_constant__ QgdDevProb d, const __grid_constant__ SweepArgs a, double* uv,
                                                                             int ncols, const double* cv, int adjoint) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* extra;
  const FastCtx<EL> c = make_fast_ctx<EL, M, NC>(d, a, smem, &extra, 0u);
  const int wpc = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = c.lane;
  const int N = d.N, N2 = d.N2;
  RegOps<EL, NC> R;
  load_regops<EL, NC>(R, d, lane, -1);
  double a_rhs[M + 1], a_tay[M + 1];
#pragma unroll
  for (int j = 0; j <= M; ++j) { a_rhs[j] = d.a_rhs[j]; a_tay[j] = d.a_tay[j]; }
  load_cv_fast<EL, M, NC>(c, cv);
  for (int col = blockIdx.x * wpc + warp; col < ncols; col += gridDim.x * wpc) {
    double* slot = uv + (size_t)N2 * (M + 1) * col;
    Vec<EL> x, out, guess;
    vload(x, slot, N, lane);
    if (!adjoint) {
      fwd_fast<EL, M, NC, true>(c, R, x, a_rhs, out, a_tay, &guess, slot);
    } else {
#pragma unroll
      for (int j = 1; j <= M; ++j) {  // Lambda_j = W_j^T x: reverse sweep with alpha = e_j
        double alpha[M + 1];
#pragma unroll
        for (int i = 0; i <= M; ++i) alpha[i] = (i == j) ? 1.0 : 0.0;
        double dK[M][NC], dS[M][NC];
        adj_fast<EL, M, NC, false>(c, R, x, alpha, out, nullptr, dK, dS);
        vstore(out, slot + (size_t)j * N2, N, lane);
      }
    }
  }
}

}  // namespace qgd

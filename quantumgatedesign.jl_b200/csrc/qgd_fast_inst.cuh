// qgd_fast_inst.cuh -- launch planning + launchers of the register-operator kernels (qgd_fast.cuh) for one
// Taylor depth M = order/2.  Included by qgd_fast_m*.cu with QGD_FAST_M defined so that the heavily unrolled
// kernels compile in parallel translation units.
#pragma once
#include "qgd_host.h"
#include <cstdlib>

#include "qgd_fast.cuh"

namespace {
using namespace qgd;

struct FastCfg { int grid, threads, wpc, ks, kt, tmem_cols, warp_doubles, h_smem, stage_l2; size_t smem; };

// One CTA per SM, all of its shared memory split between the warps; whatever is left after the fixed
// per-warp arrays holds the resident part of the Krylov basis.
template <class K>
FastCfg plan_fast(qgd_handle* h, K kernel, int fixed_doubles, int el, size_t items, int extra_doubles, int restart, int rs = 1,
                  int group_dbl = 0) {
  const size_t max_smem = h->prop.sharedMemPerBlockOptin;
  const int sms = h->prop.multiProcessorCount;
  FastCfg L{};
  // rs > 1 (row-split groups): an item takes rs warps of a CTA
  L.wpc = rs * (int)std::min<size_t>(QGD_WARPS_PER_CTA / rs, std::max<size_t>(1, (items + sms - 1) / sms));
  if (h->opt[QGD_OPT_LATENCY_WARPS] > 0 && rs == 1) L.wpc = (int)h->opt[QGD_OPT_LATENCY_WARPS];
  const int ngroups = rs > 1 ? L.wpc / rs : 0;
  const int vec = 2 * 32 * el;
  const int base = (fixed_doubles + extra_doubles + 1) & ~1;
  // tensor-memory tier: warps w and w + 4 of a CTA share a lane quarter, so each gets 512 / groups columns;
  // a vector takes 4 * el columns.  Allocation must be a power of two >= 32 columns.
  const int groups = (L.wpc + 3) / 4;
  L.kt = restart >= 2 && !h->opt[QGD_OPT_DISABLE_TMEM] ? std::min(restart + 1, (512 / groups) / (4 * el)) : 0;
  L.tmem_cols = 0;
  if (L.kt > 0) {
    int need = groups * L.kt * 4 * el, pow2 = 32;
    while (pow2 < need) pow2 *= 2;
    L.tmem_cols = pow2;
    L.kt = std::min(restart + 1, (pow2 / groups) / (4 * el));
  }
  constexpr int blk = QGD_MGS_BLOCK > 1 ? QGD_MGS_BLOCK : 1;  // a Gram-Schmidt block never straddles two tiers
  L.kt = (L.kt / blk) * blk;
  const long per_warp = (long)(((max_smem - 16) / 8 - (size_t)ngroups * group_dbl) / L.wpc) & ~1L;
  // Few columns in flight (one or two warps per SM -- a single gradient evaluation, what optimize_gate asks for): the
  // packed Hessenberg matrix of the warp fits into shared memory beside the whole Krylov basis, so the end-of-solve
  // least squares reads shared memory instead of waiting an L2 round trip per rotation.
  L.h_smem = 0;
  {
    const long hd = ((long)hpk(restart) + 8 + 3) & ~3L;  // packed Hessenberg matrix + alignment slack
    const long want_vec = std::max(restart + 1 - L.kt, 0);
    if (L.wpc <= 2 && restart >= 2 && per_warp - base - hd >= std::min<long>(want_vec, 16) * vec) L.h_smem = (int)hd;
  }
  long ks = (per_warp - base - L.h_smem) / vec;
  ks = std::min<long>(ks, std::max(restart + 1 - L.kt, 0));
  if (ks < 0) throw QgdError(QGD_EUNSUPPORTED, "fast path: shared memory too small for the per-warp state");
  L.ks = (int)(ks / blk) * blk;
  ks = L.ks;
  // row-split groups: when part of the basis has to live in L2, the last block of the shared-memory tier becomes the staging
  // block its L2 tier is prefetched into (qgd_fast.cuh, QGD_RS_STAGE_L2)
  L.stage_l2 = 0;
  if (rs > 1 && QGD_RS_STAGE_L2 && blk == 8 && L.ks >= 16 && restart + 1 > L.kt + L.ks) { L.stage_l2 = 1; L.ks -= 8; }
  L.warp_doubles = base + (int)ks * vec + L.h_smem;
  L.threads = 32 * L.wpc;
  L.smem = 16 + ((size_t)L.wpc * L.warp_doubles + (size_t)ngroups * group_dbl) * 8;
  CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  const size_t per_cta = (size_t)(L.wpc / rs);
  const size_t ctas = (items + per_cta - 1) / per_cta;
  L.grid = (int)std::max<size_t>(1, std::min<size_t>(ctas, (size_t)sms));
  return L;
}

// Latency team (qgd_fast.cuh): one CTA of four warps per column, the whole Krylov basis in the four TMEM lane quarters, the
// packed Hessenberg matrix and the team's exchange buffers in shared memory.
template <class K>
FastCfg plan_fast_team(qgd_handle* h, K kernel, int fixed_doubles, int el, size_t items, int restart) {
  const size_t max_smem = h->prop.sharedMemPerBlockOptin;
  FastCfg L{};
  L.wpc = QGD_TEAM_WARPS;
  L.threads = 32 * QGD_TEAM_WARPS;
  L.kt = 64; L.ks = 0; L.tmem_cols = 512;
  if (restart + 1 > QGD_TEAM_WARPS * 64) throw QgdError(QGD_EUNSUPPORTED, "latency team: Krylov basis larger than the tensor memory of four warps");
  L.h_smem = (int)(((long)hpk(restart) + 8 + 3) & ~3L);
  const int team = el == 1 ? team_doubles<1>() : team_doubles<2>();
  L.warp_doubles = ((fixed_doubles + 1) & ~1) + L.h_smem + team + 8;
  L.smem = 16 + (size_t)L.warp_doubles * 8;
  if (L.smem > max_smem) throw QgdError(QGD_EUNSUPPORTED, "latency team: shared memory too small");
  CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  L.grid = (int)std::max<size_t>(1, std::min<size_t>(items, (size_t)h->prop.multiProcessorCount));
  return L;
}

// The Krylov tail and the packed Hessenberg matrices of all resident warps live in ONE allocation (d_V) so that a
// single L2 access-policy window can cover them (launch_sweep below).
void ensure_krylov_fast(qgd_handle* h, const FastCfg& L, int el, int restart, SweepArgs& a) {
  const size_t warps = (size_t)L.grid * L.wpc;
  a.v_stride = (size_t)(std::max(restart + 1 - L.ks - L.kt, 1) + QGD_MGS_BLOCK) * 2 * 32 * el;
  a.h_stride = ((size_t)std::max(hpk(restart), restart * (restart + 3) / 2) + 8 + 3) & ~(size_t)3;  // 32-byte aligned slots
  const size_t v_bytes = (warps * a.v_stride * 8 + 255) & ~(size_t)255;
  h->d_V.reserve(v_bytes + warps * a.h_stride * 8);
  a.Vws = h->d_V.as<double>();
  a.Hws = reinterpret_cast<double*>(h->d_V.as<unsigned char>() + v_bytes);
  ensure_counter(h);
  CUDA_CHECK(cudaMemsetAsync(h->d_counter.p, 0, 4, h->stream));  // ticket counter; the error word behind it is sticky
  a.work_counter = h->d_counter.as<unsigned int>();
  a.err = h->d_counter.as<int>() + 1;
  // time-sliced tickets: about 24 segments per column, published through d_progress
  const size_t items = (size_t)a.B * h->ncol;
  const int nsteps = (int)h->nsteps;
  a.seg_steps = h->opt[QGD_OPT_SEG_STEPS] > 0 ? (int)h->opt[QGD_OPT_SEG_STEPS] : std::max(1, (nsteps + 23) / 24);
  h->d_progress.reserve(std::max<size_t>(items, 1) * 4);
  CUDA_CHECK(cudaMemsetAsync(h->d_progress.p, 0, std::max<size_t>(items, 1) * 4, h->stream));
  a.progress = h->d_progress.as<int>();
  h->d_carry.reserve(std::max<size_t>(items, 1) * (size_t)h->N2 * 8);
  a.carry = h->d_carry.as<double>();
  a.ks = L.ks; a.kt = L.kt; a.tmem_cols = L.tmem_cols; a.h_smem_doubles = L.h_smem; a.stage_l2 = L.stage_l2;
  a.warp_smem_doubles = L.warp_doubles;
}

// Launch a sweep kernel, optionally (QGD_L2_PERSIST=1) with a persisting L2 access-policy window over the Krylov /
// Hessenberg workspace.  Measured on B200 (profiles/r01_l2_policy.txt): the window TRIPLES the DRAM write traffic of a
// sweep (7.8 GB instead of 2.3 GB at batch 148 x 110 steps) at equal run time, so it is off by default; what does help
// is accessing the state history with evict-first loads and stores (-18 % writes, -50 % reads).
template <class... KArgs, class... Args>
void launch_sweep(qgd_handle* h, void (*kernel)(KArgs...), const FastCfg& L, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)L.grid); cfg.blockDim = dim3((unsigned)L.threads);
  cfg.dynamicSmemBytes = L.smem; cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  int n = 0;

  const size_t carve = (size_t)std::max(h->prop.persistingL2CacheMaxSize, 0);
  const size_t max_win = (size_t)std::max(h->prop.accessPolicyMaxWindowSize, 0);
  if (h->opt[QGD_OPT_L2_PERSIST] == 1 && carve > 0 && max_win > 0 && h->d_V.cap > 0) {
    if (!h->l2_carved) {
      CUDA_CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
      h->l2_carved = true;
    }
    const size_t bytes = std::min(h->d_V.cap, max_win);
    at[0].id = cudaLaunchAttributeAccessPolicyWindow;
    at[0].val.accessPolicyWindow.base_ptr = h->d_V.p;
    at[0].val.accessPolicyWindow.num_bytes = bytes;
    at[0].val.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)bytes);
    at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    n = 1;
  }
  cfg.attrs = at; cfg.numAttrs = (unsigned)n;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, args...));
  h->stats.kernel_launches++;
}

template <int EL, int M, int NC, bool STRICT, bool FORCED = false>
void launch_forward_fast_t(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  FastCfg L = plan_fast(h, k_forward_fast<EL, M, NC, STRICT, FORCED>, fast_fixed_doubles<EL, M, NC, STRICT>(d.N2), EL, (size_t)a.B * d.ncol, 0, d.N2);
  ensure_krylov_fast(h, L, EL, d.N2, a);
  launch_sweep(h, k_forward_fast<EL, M, NC, STRICT, FORCED>, L, d, a);
}
template <int EL, int M, int NC, bool STRICT>
void launch_backward_fast_t(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  const int extra = QGD_COMPACT_SMEM ? 0 : d.P + 2 * M * NC;
  FastCfg L = plan_fast(h, k_backward_fast<EL, M, NC, STRICT>, fast_fixed_doubles<EL, M, NC, STRICT>(d.N2), EL, (size_t)a.B * d.ncol, extra, d.N2);
  ensure_krylov_fast(h, L, EL, d.N2, a);
  launch_sweep(h, k_backward_fast<EL, M, NC, STRICT>, L, d, a, (const QgdDevControl*)h->d_ctrls.as<QgdDevControl>());
}
// Row-split groups (qgd_fast.cuh, FastCtx RS): sparse problems with 64 < N <= 128 (two warps per column) and 128 < N <= 256 (four)
template <int M, int NC, int RS>
void launch_forward_fast_rs_t(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  auto kernel = k_forward_fast<2, M, NC, false, false, false, RS>;
  FastCfg L = plan_fast(h, kernel, fast_fixed_doubles<2, M, NC, false, RS>(d.N2), 2, (size_t)a.B * d.ncol, 0, d.N2, RS, group_doubles<2, RS>());
  ensure_krylov_fast(h, L, 2, d.N2, a);
  launch_sweep(h, kernel, L, d, a);
}
template <int M, int NC, int RS>
void launch_forward_forced_fast_rs_t(qgd_handle* h, QgdDevProb d, SweepArgs a) {  // forced solves on the groups
  auto kernel = k_forward_fast<2, M, NC, false, true, false, RS>;
  FastCfg L = plan_fast(h, kernel, fast_fixed_doubles<2, M, NC, false, RS>(d.N2), 2, (size_t)a.B * d.ncol, 0, d.N2, RS, group_doubles<2, RS>());
  ensure_krylov_fast(h, L, 2, d.N2, a);
  launch_sweep(h, kernel, L, d, a);
}
template <int M, int NC, int RS>
void launch_backward_fast_rs_t(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  auto kernel = k_backward_fast<2, M, NC, false, false, RS>;
  FastCfg L = plan_fast(h, kernel, fast_fixed_doubles<2, M, NC, false, RS>(d.N2), 2, (size_t)a.B * d.ncol, 0, d.N2, RS, group_doubles<2, RS>());
  ensure_krylov_fast(h, L, 2, d.N2, a);
  launch_sweep(h, kernel, L, d, a, (const QgdDevControl*)h->d_ctrls.as<QgdDevControl>());
}
template <int M, int NC, int RS>
void launch_terminal_fast_rs_t(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  const int restart = std::min(20, d.N2);
  auto kernel = k_terminal_fast<2, M, NC, RS>;
  FastCfg L = plan_fast(h, kernel, fast_fixed_doubles<2, M, NC, true, RS>(d.N2), 2, (size_t)a.B, 0, restart, RS, group_doubles<2, RS>());
  ensure_krylov_fast(h, L, 2, restart, a);
  launch_sweep(h, kernel, L, d, a);
}
template <int EL, int M, int NC>
void launch_terminal_fast_t(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  const int restart = std::min(20, d.N2);
  FastCfg L = plan_fast(h, k_terminal_fast<EL, M, NC>, fast_fixed_doubles<EL, M, NC, true>(d.N2), EL, (size_t)a.B, 0, restart);
  ensure_krylov_fast(h, L, EL, restart, a);
  launch_sweep(h, k_terminal_fast<EL, M, NC>, L, d, a);
}
template <int EL, int M, int NC>
void launch_derivs_fast_t(qgd_handle* h, QgdDevProb d, SweepArgs a, double* uv, int ncols, const double* cv, int adjoint) {
  FastCfg L = plan_fast(h, k_derivs_fast<EL, M, NC>, fast_fixed_doubles<EL, M, NC>(d.N2), EL, (size_t)ncols, 0, 1);
  ensure_krylov_fast(h, L, EL, 1, a);
  k_derivs_fast<EL, M, NC><<<L.grid, L.threads, L.smem, h->stream>>>(d, a, uv, ncols, cv, adjoint);
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}

// (EL, NC) shapes built for every M: N <= 32 with 1..4 control operators (four two-level qubits: N = 16), N <= 64 with 2..4
// (four qubits with guard levels on some: (3,3,2,2) = 36 ... (3,3,3,2) = 54 levels).
#define QGD_FAST_SHAPES(X, M) X(1, M, 1) X(1, M, 2) X(1, M, 3) X(1, M, 4) X(2, M, 2) X(2, M, 3) X(2, M, 4)

}  // namespace

#define QGD_FAST_CASE_FWD(EL, M, NC) if (el == EL && nc == NC) { launch_forward_fast_t<EL, M, NC, false>(h, d, a); return true; }
#define QGD_FAST_CASE_BWD(EL, M, NC) if (el == EL && nc == NC) { launch_backward_fast_t<EL, M, NC, false>(h, d, a); return true; }
#define QGD_FAST_CASE_FWD_S(EL, M, NC) if (el == EL && nc == NC) { launch_forward_fast_t<EL, M, NC, true>(h, d, a); return true; }
#define QGD_FAST_CASE_FWD_F(EL, M, NC) if (el == EL && nc == NC) { launch_forward_fast_t<EL, M, NC, false, true>(h, d, a); return true; }
template <int EL, int M, int NC>
void launch_forward_fast_team_t(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  FastCfg L = plan_fast_team(h, k_forward_fast<EL, M, NC, false, false, true>, fast_fixed_doubles<EL, M, NC, false>(d.N2), EL, (size_t)a.B * d.ncol, d.N2);
  ensure_krylov_fast(h, L, EL, d.N2, a);
  launch_sweep(h, k_forward_fast<EL, M, NC, false, false, true>, L, d, a);
}
template <int EL, int M, int NC>
void launch_backward_fast_team_t(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  FastCfg L = plan_fast_team(h, k_backward_fast<EL, M, NC, false, true>, fast_fixed_doubles<EL, M, NC, false>(d.N2), EL, (size_t)a.B * d.ncol, d.N2);
  ensure_krylov_fast(h, L, EL, d.N2, a);
  launch_sweep(h, k_backward_fast<EL, M, NC, false, true>, L, d, a, (const QgdDevControl*)h->d_ctrls.as<QgdDevControl>());
}
#define QGD_FAST_CASE_FWD_T(EL, M, NC) if (el == EL && nc == NC) { launch_forward_fast_team_t<EL, M, NC>(h, d, a); return true; }
#define QGD_FAST_CASE_BWD_T(EL, M, NC) if (el == EL && nc == NC) { launch_backward_fast_team_t<EL, M, NC>(h, d, a); return true; }
#define QGD_FAST_CASE_BWD_S(EL, M, NC) if (el == EL && nc == NC) { launch_backward_fast_t<EL, M, NC, true>(h, d, a); return true; }
#define QGD_FAST_CASE_TRM(EL, M, NC) if (el == EL && nc == NC) { launch_terminal_fast_t<EL, M, NC>(h, d, a); return true; }
#define QGD_FAST_CASE_DER(EL, M, NC) if (el == EL && nc == NC) { launch_derivs_fast_t<EL, M, NC>(h, d, a, uv, ncols, cv, adjoint); return true; }

#define QGD_DEFINE_FAST_LAUNCHERS(M)                                                                                        \
  bool launch_forward_fast_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc) {                              \
    QGD_FAST_SHAPES(QGD_FAST_CASE_FWD, M) return false;                                                                     \
  }                                                                                                                         \
  bool launch_backward_fast_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc) {                             \
    QGD_FAST_SHAPES(QGD_FAST_CASE_BWD, M) return false;                                                                     \
  }                                                                                                                         \
  bool launch_derivs_fast_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc, double* uv, int ncols,          \
                               const double* cv, int adjoint) {                                                             \
    QGD_FAST_SHAPES(QGD_FAST_CASE_DER, M) return false;                                                                     \
  }                                                                                                                         \
  bool launch_terminal_fast_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc) {                             \
    QGD_FAST_SHAPES(QGD_FAST_CASE_TRM, M) return false;                                                                     \
  }

// The same sweeps with strict modified Gram-Schmidt (QGD_OPT_STRICT_MGS), in translation units of their own.
#define QGD_DEFINE_FAST_LAUNCHERS_STRICT(M)                                                                                 \
  bool launch_forward_fast_strict_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc) {                       \
    QGD_FAST_SHAPES(QGD_FAST_CASE_FWD_S, M) return false;                                                                   \
  }                                                                                                                         \
  bool launch_backward_fast_strict_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc) {                      \
    QGD_FAST_SHAPES(QGD_FAST_CASE_BWD_S, M) return false;                                                                   \
  }

// The forced forward solves (eval_forward! with forcing, eval_grad_forced) on the register-operator sweeps, in translation
// units of their own.
#define QGD_DEFINE_FAST_LAUNCHERS_FORCED(M)                                                                                 \
  bool launch_forward_fast_forced_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc) {                       \
    QGD_FAST_SHAPES(QGD_FAST_CASE_FWD_F, M) return false;                                                                   \
  }

// Row-split groups in translation units of their own: (RS, NC) in {2, 4} x {2, 3, 4}.
#define QGD_FAST_CASE_RS(WHICH, M, NC, RS) if (rs == RS && nc == NC) { launch_##WHICH##_fast_rs_t<M, NC, RS>(h, d, a); return true; }
#define QGD_DEFINE_FAST_LAUNCHERS_RS(M)                                                                                     \
  bool launch_forward_fast_rs_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int rs, int nc) {                           \
    QGD_FAST_CASE_RS(forward, M, 2, 2) QGD_FAST_CASE_RS(forward, M, 3, 2) QGD_FAST_CASE_RS(forward, M, 2, 4)                  \
    QGD_FAST_CASE_RS(forward, M, 3, 4) QGD_FAST_CASE_RS(forward, M, 4, 2) QGD_FAST_CASE_RS(forward, M, 4, 4) return false;                                                                        \
  }                                                                                                                         \
  bool launch_backward_fast_rs_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int rs, int nc) {                          \
    QGD_FAST_CASE_RS(backward, M, 2, 2) QGD_FAST_CASE_RS(backward, M, 3, 2) QGD_FAST_CASE_RS(backward, M, 2, 4)               \
    QGD_FAST_CASE_RS(backward, M, 3, 4) QGD_FAST_CASE_RS(backward, M, 4, 2) QGD_FAST_CASE_RS(backward, M, 4, 4) return false; \
  }                                                                                                                         \
  bool launch_terminal_fast_rs_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int rs, int nc) {                          \
    QGD_FAST_CASE_RS(terminal, M, 2, 2) QGD_FAST_CASE_RS(terminal, M, 3, 2) QGD_FAST_CASE_RS(terminal, M, 2, 4)               \
    QGD_FAST_CASE_RS(terminal, M, 3, 4) QGD_FAST_CASE_RS(terminal, M, 4, 2) QGD_FAST_CASE_RS(terminal, M, 4, 4) return false; \
  }

// The forced forward solves on the row-split groups, again in translation units of their own (qgd_fast_rf_m<m>.o).
#define QGD_DEFINE_FAST_LAUNCHERS_RS_FORCED(M)                                                                              \
  bool launch_forward_fast_forced_rs_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int rs, int nc) {                    \
    QGD_FAST_CASE_RS(forward_forced, M, 2, 2) QGD_FAST_CASE_RS(forward_forced, M, 3, 2) QGD_FAST_CASE_RS(forward_forced, M, 4, 2) \
    QGD_FAST_CASE_RS(forward_forced, M, 2, 4) QGD_FAST_CASE_RS(forward_forced, M, 3, 4) QGD_FAST_CASE_RS(forward_forced, M, 4, 4) \
    return false;                                                                                                           \
  }

// The latency team (four warps per column) in translation units of its own.
#define QGD_DEFINE_FAST_LAUNCHERS_TEAM(M)                                                                                   \
  bool launch_forward_fast_team_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc) {                         \
    QGD_FAST_SHAPES(QGD_FAST_CASE_FWD_T, M) return false;                                                                   \
  }                                                                                                                         \
  bool launch_backward_fast_team_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc) {                        \
    QGD_FAST_SHAPES(QGD_FAST_CASE_BWD_T, M) return false;                                                                   \
  }

// qgd_inst.cuh -- launch planning + the per-EL kernel launchers.  Included by qgd_inst_el*.cu with QGD_EL defined
// so that the (slow to compile) sweep kernels build in parallel translation units.
#pragma once
#include "qgd_host.h"
#include "qgd_kernels.cuh"

namespace {
using namespace qgd;

struct LaunchCfg { int grid, threads, wpc; size_t smem; int warp_doubles; bool ops_in_smem; };

// Shared-memory carve-up and grid size for a sweep-type kernel.
template <class K>
LaunchCfg plan_launch(qgd_handle* h, K kernel, int m, size_t items, int extra_doubles) {
  const int nv = std::max(m, 2);
  const int wd = nv * h->N2 + 2 * (m + 1) * h->Nc + 3 * (h->N2 + 1) + extra_doubles;
  const int warp_doubles = (wd + 1) & ~1;
  const size_t max_smem = h->prop.sharedMemPerBlockOptin;
  LaunchCfg L{};
  L.warp_doubles = warp_doubles;
  int wpc = (int)std::min<size_t>(QGD_WARPS_PER_CTA, std::max<size_t>(1, (items + h->prop.multiProcessorCount - 1) / h->prop.multiProcessorCount));
  L.ops_in_smem = (16 + (size_t)h->lay.bytes + (size_t)warp_doubles * 8) <= max_smem;
  const size_t fixed = 16 + (L.ops_in_smem ? (size_t)h->lay.bytes : 0);
  while (wpc > 1 && fixed + (size_t)wpc * warp_doubles * 8 > max_smem) --wpc;
  if (fixed + (size_t)wpc * warp_doubles * 8 > max_smem)
    throw QgdError(QGD_EUNSUPPORTED, "problem too large for the shared-memory resident warp state");
  L.wpc = wpc; L.threads = 32 * wpc; L.smem = fixed + (size_t)wpc * warp_doubles * 8;
  CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  int per_sm = 0;
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, L.threads, L.smem));
  if (per_sm < 1) throw QgdError(QGD_ECUDA, "kernel cannot be resident with this configuration");
  const size_t ctas = (items + wpc - 1) / wpc;
  L.grid = (int)std::max<size_t>(1, std::min<size_t>(ctas, (size_t)per_sm * h->prop.multiProcessorCount));
  return L;
}

void ensure_krylov(qgd_handle* h, const LaunchCfg& L, int restart, SweepArgs& a) {
  const size_t warps = (size_t)L.grid * L.wpc;
  a.v_stride = (size_t)(restart + 1) * h->N2;
  a.h_stride = (size_t)restart * (restart + 3) / 2 + 2;
  h->d_V.reserve(warps * a.v_stride * 8);
  h->d_H.reserve(warps * a.h_stride * 8);
  a.Vws = h->d_V.as<double>();
  a.Hws = h->d_H.as<double>();
}

template <int EL>
void launch_forward(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  LaunchCfg L = plan_launch(h, k_forward<EL>, d.m, (size_t)a.B * d.ncol, 0);
  d.ops_in_smem = L.ops_in_smem; a.warp_smem_doubles = L.warp_doubles;
  ensure_krylov(h, L, d.N2, a);
  k_forward<EL><<<L.grid, L.threads, L.smem, h->stream>>>(d, a);
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}
template <int EL>
void launch_guard(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  LaunchCfg L = plan_launch(h, k_guard<EL>, d.m, (size_t)a.B * d.ncol, 0);
  d.ops_in_smem = L.ops_in_smem; a.warp_smem_doubles = L.warp_doubles;
  ensure_krylov(h, L, 1, a);
  k_guard<EL><<<L.grid, L.threads, L.smem, h->stream>>>(d, a);
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}
template <int EL>
void launch_terminal(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  LaunchCfg L = plan_launch(h, k_terminal<EL>, d.m, (size_t)a.B, 0);
  d.ops_in_smem = L.ops_in_smem; a.warp_smem_doubles = L.warp_doubles;
  ensure_krylov(h, L, std::min(20, d.N2), a);
  k_terminal<EL><<<L.grid, L.threads, L.smem, h->stream>>>(d, a);
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}
template <int EL>
void launch_backward(qgd_handle* h, QgdDevProb d, SweepArgs a) {
  const int extra = d.P + 2 * QGD_MAX_M * (QGD_MAX_OPS - 1);
  LaunchCfg L = plan_launch(h, k_backward<EL>, d.m, (size_t)a.B * d.ncol, extra);
  d.ops_in_smem = L.ops_in_smem; a.warp_smem_doubles = L.warp_doubles;
  ensure_krylov(h, L, d.N2, a);
  k_backward<EL><<<L.grid, L.threads, L.smem, h->stream>>>(d, a, h->d_ctrls.as<QgdDevControl>());
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}
template <int EL>
void launch_lambda_columns(qgd_handle* h, QgdDevProb d, SweepArgs a, double* lam_hist) {
  LaunchCfg L = plan_launch(h, k_lambda_columns<EL>, d.m, (size_t)a.B * d.ncol * (d.nsteps + 1), 0);
  d.ops_in_smem = L.ops_in_smem; a.warp_smem_doubles = L.warp_doubles;
  ensure_krylov(h, L, 1, a);
  k_lambda_columns<EL><<<L.grid, L.threads, L.smem, h->stream>>>(d, a, lam_hist);
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}
template <int EL>
void launch_derivs(qgd_handle* h, QgdDevProb d, SweepArgs a, double* uv, int ncols, const double* cv, int adjoint) {
  LaunchCfg L = plan_launch(h, k_derivs<EL>, d.m, (size_t)ncols, 0);
  d.ops_in_smem = L.ops_in_smem; a.warp_smem_doubles = L.warp_doubles;
  ensure_krylov(h, L, 1, a);
  k_derivs<EL><<<L.grid, L.threads, L.smem, h->stream>>>(d, a, uv, ncols, cv, adjoint);
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}


}  // namespace

#define QGD_DEFINE_LAUNCHERS(EL)                                                                                                  \
  void launch_forward_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a) { launch_forward<EL>(h, d, a); }                        \
  void launch_guard_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a) { launch_guard<EL>(h, d, a); }                            \
  void launch_terminal_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a) { launch_terminal<EL>(h, d, a); }                      \
  void launch_backward_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a) { launch_backward<EL>(h, d, a); }                      \
  void launch_lambda_columns_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, double* lam_hist) {                              \
    launch_lambda_columns<EL>(h, d, a, lam_hist);                                                                                 \
  }                                                                                                                               \
  void launch_derivs_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, double* uv, int ncols, const double* cv, int adjoint) { \
    launch_derivs<EL>(h, d, a, uv, ncols, cv, adjoint);                                                                           \
  }

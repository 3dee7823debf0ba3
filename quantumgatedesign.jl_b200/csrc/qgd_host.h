// qgd_host.h -- host-side types shared by the API translation unit and the per-EL kernel instantiation units.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qgd_b200.h"
#include "qgd_common.h"

namespace qgd {
struct SweepArgs;
}

struct QgdError : std::runtime_error {
  int code;
  QgdError(int c, const std::string& s) : std::runtime_error(s), code(c) {}
};

#define CUDA_CHECK(expr)                                                                                   \
  do {                                                                                                     \
    cudaError_t e_ = (expr);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      throw QgdError(e_ == cudaErrorMemoryAllocation ? QGD_ENOMEM : QGD_ECUDA,                              \
                     std::string(#expr) + ": " + cudaGetErrorString(e_));                                  \
  } while (0)

typedef std::vector<double> dvec;

struct DevBuf {  // grow-only device buffer
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    CUDA_CHECK(cudaMalloc(&p, bytes));
    cap = bytes;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct qgd_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop;
  // problem
  int N = 0, N2 = 0, Nc = 0, nic = 0, Ness = 0, P = 0, precond = 0;
  int64_t nsteps = 0;
  double tf = 0, abstol = 0, reltol = 0;
  int col0 = 0, ncol = 0;
  dvec Ks, Ss;  // dense drift (preconditioner setup)
  std::vector<QgdDevControl> ctrls;
  std::vector<dvec> ctrl_freqs, ctrl_knots;
  // operator blob
  QgdOpLayout lay;
  std::vector<unsigned char> blob;
  int pre_key_nsteps = -1, pre_key_order = -1;
  DevBuf d_blob, d_minv[2], d_u0, d_v0, d_ctrls, d_aux;
  // control table cache
  int tab_key_nsteps = -1, tab_key_m = -1;
  DevBuf d_table;
  // batch buffers
  DevBuf d_pcof, d_cvals, d_history, d_final, d_final_all, d_terminal, d_lambda0, d_lamhist, d_gradcol, d_grad, d_guardcol,
      d_guard, d_infid, d_iters_f, d_iters_a, d_iters_t, d_target, d_forcing, d_V, d_H, d_scratch, d_counter, d_progress, d_carry, d_theta_op;
  // state of the device-resident history
  int hist_B = 0, hist_order = 0;
  int64_t hist_nsteps = 0, hist_save = 0;
  bool hist_valid = false;
  int phase_B = 0, phase_order = 0;  // two-phase API
  int pend_B = 0, pend_order = 0, pend_iters = 0;  // qgd_eval_forward_async waiting for its _collect
  int64_t pend_save = 0;
  // register-operator fast path (qgd_fast.cuh): structure test done once at creation
  bool fast_ok = false;
  int fast_el = 0;
  int fast_rs = 1;  // warps per column: 1 (N <= 64), 2 (N <= 128), 4 (N <= 256) -- row-split groups of qgd_fast.cuh
  // dense tensor-core path (qgd_dense.cu): row-major dense copies of the operators [(Nc+1)][2][N][N] (operator 0 = drift)
  std::vector<double> dense_ops;
  DevBuf d_dense, d_comb, d_dense_ws;
  bool host_controls = false;     // some control is QGD_CONTROL_HOST_TABLE: only the qgd_*_tables entry points work
  bool tables_from_host = false;  // inside a qgd_*_tables call: d_cvals / d_table hold the caller's tables
  bool l2_carved = false;  // cudaLimitPersistingL2CacheSize set for the workspace window (qgd_fast_inst.cuh)
  int64_t opt[16] = {0};   // qgd_set_option values, indexed by QGD_OPT_*
  // multi-GPU (qgd_multi.cu): NCCL communicator over the handles that share one evaluation by columns
  void* comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  DevBuf d_pack, d_dots;   // [grad P*B | guard B] all-reduced at the end of an evaluation; [2][B] terminal inner products
  int* d_err = nullptr;    // device error word of the sweeps (lives behind the ticket counter in d_counter)
  std::vector<double> hist_pcof;  // the control vectors the resident history was computed for (history_precomputed check)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  qgd_stats_t stats{};
};

// [0] ticket counter of the sweep being launched, [1] sticky device error word (see wait_segment)
inline void ensure_counter(qgd_handle* h) {
  if (h->d_counter.cap) return;
  h->d_counter.reserve(64);
  CUDA_CHECK(cudaMemsetAsync(h->d_counter.p, 0, 64, h->stream));
}

// NCCL, bound at run time (qgd_multi.cu)
namespace qgd_nccl {
void set_library(const char* path);
int version();
void unique_id(unsigned char out[128]);
void* init_rank(int nranks, int rank, const unsigned char id[128]);
void init_all(void** comms, int n, const int* devices);
void destroy(void* comm);
void allreduce_sum(void* comm, double* buf, size_t n, cudaStream_t stream);
void group_start();
void group_end();
}  // namespace qgd_nccl

// Kernel launchers, one set per EL = levels per lane (instantiated in qgd_inst_el*.cu).
#define QGD_DECLARE_LAUNCHERS(EL)                                                                                   \
  void launch_forward_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a);                                          \
  void launch_guard_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a);                                            \
  void launch_terminal_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a);                                         \
  void launch_backward_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a);                                         \
  void launch_lambda_columns_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, double* lam_hist);                 \
  void launch_derivs_##EL(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, double* uv, int ncols, const double* cv, int adjoint);

QGD_DECLARE_LAUNCHERS(1)
QGD_DECLARE_LAUNCHERS(2)
QGD_DECLARE_LAUNCHERS(4)
QGD_DECLARE_LAUNCHERS(8)

// Register-operator kernels (qgd_fast.cuh), one translation unit per Taylor depth M = order/2.  Each launcher
// returns false when the (levels-per-lane, operator count) shape was not built; the caller then uses the
// generic kernels.
#define QGD_DECLARE_FAST_LAUNCHERS(M)                                                                              \
  bool launch_forward_fast_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc);                     \
  bool launch_backward_fast_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc);                    \
  bool launch_derivs_fast_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc, double* uv, int ncols, \
                               const double* cv, int adjoint);                                                     \
  bool launch_terminal_fast_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc);                    \
  bool launch_forward_fast_strict_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc);              \
  bool launch_backward_fast_strict_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc);             \
  bool launch_forward_fast_forced_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc);              \
  bool launch_forward_fast_team_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc);                \
  bool launch_backward_fast_team_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int el, int nc);               \
  bool launch_forward_fast_rs_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int rs, int nc);                  \
  bool launch_backward_fast_rs_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int rs, int nc);                 \
  bool launch_terminal_fast_rs_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int rs, int nc);                 \
  bool launch_forward_fast_forced_rs_m##M(qgd_handle* h, QgdDevProb d, qgd::SweepArgs a, int rs, int nc);
QGD_DECLARE_FAST_LAUNCHERS(1)
QGD_DECLARE_FAST_LAUNCHERS(2)
QGD_DECLARE_FAST_LAUNCHERS(3)
QGD_DECLARE_FAST_LAUNCHERS(4)
QGD_DECLARE_FAST_LAUNCHERS(5)
QGD_DECLARE_FAST_LAUNCHERS(6)
#define QGD_FAST_MAX_M 6

// FP64 tensor-core Taylor recursion for dense operators (qgd_dense.cu)
bool dense_derivs_applicable(const qgd_handle* h, int m);
void launch_derivs_dense(qgd_handle* h, int m, double* d_uv, int ncols, const double* d_cv, int adjoint);
bool try_forward_dense(qgd_handle* h, const QgdDevProb& d, const qgd::SweepArgs& a);
bool try_backward_dense(qgd_handle* h, const QgdDevProb& d, const qgd::SweepArgs& a);
bool try_terminal_dense(qgd_handle* h, const QgdDevProb& d, const qgd::SweepArgs& a);

// qgd_b200.cu -- host layer + C ABI (include/qgd_b200.h) of the B200-native gradient hot path.
// Standalone CUDA-runtime code: no torch, no CPU fallback.  Every entry point fails with QGD_ECUDA
// when no sm_100 device is available.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qgd_b200.h"
#include "qgd_host.h"
#include "qgd_controls.cuh"
#include "qgd_kernels.cuh"

namespace {

thread_local std::string g_err;

// ---- small dense helpers (host-side setup only) --------------------------------------------------------
dvec mat_to_dense(const qgd_matrix_t& a) {
  dvec d((size_t)a.nrows * a.ncols, 0.0);
  if (a.kind == QGD_MAT_CSC) {
    for (int64_t j = 0; j < a.ncols; ++j)
      for (int64_t p = a.colptr[j] - 1; p < a.colptr[j + 1] - 1; ++p) d[(size_t)(a.rowval[p] - 1) + a.nrows * j] += a.nzval[p];
  } else if (a.kind == QGD_MAT_DENSE) {
    std::copy(a.dense, a.dense + (size_t)a.nrows * a.ncols, d.begin());
  } else {
    throw QgdError(QGD_EINVAL, "unknown matrix kind");
  }
  return d;
}
dvec matmul(const dvec& A, const dvec& B, int n) {
  dvec C((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int k = 0; k < n; ++k) {
      double b = B[k + (size_t)n * j];
      if (b == 0.0) continue;
      for (int i = 0; i < n; ++i) C[i + (size_t)n * j] += A[i + (size_t)n * k] * b;
    }
  return C;
}
dvec matpow(const dvec& A, int p, int n) {  // Julia's A^p (Base.power_by_squaring)
  auto tz = [](int v) { int c = 0; while (!(v & 1)) { v >>= 1; ++c; } return c; };
  dvec x = A;
  int t = tz(p) + 1; p >>= t;
  while ((t -= 1) > 0) x = matmul(x, x, n);
  dvec y = x;
  while (p > 0) {
    t = tz(p) + 1; p >>= t;
    while ((t -= 1) >= 0) x = matmul(x, x, n);
    y = matmul(y, x, n);
  }
  return y;
}
double factorial_d(int n) { double f = 1; for (int i = 2; i <= n; ++i) f *= i; return f; }
double ipow(double x, int p) { double r = 1, b = x; while (p > 0) { if (p & 1) r *= b; p >>= 1; if (p) b *= b; } return r; }
double coefficient(int j, int p, int q) {  // src/hermite.jl:389-391
  return factorial_d(p) * factorial_d(p + q - j) / (factorial_d(p + q) * factorial_d(p - j));
}

int align16(int x) { return (x + 15) & ~15; }

}  // namespace

namespace qgd {

// grad[b][theta] = sum over owned columns (fixed order); guard[b] likewise.
__global__ void k_finalize(int P, int ncol, int B, const double* gradcol, const double* guardcol, double* grad, double* guard) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (size_t)P * B) {
    const int b = (int)(idx / P), t = (int)(idx % P);
    double s = 0.0;
    for (int cl = 0; cl < ncol; ++cl) s += gradcol[(size_t)t + (size_t)P * ((size_t)cl + (size_t)ncol * b)];
    grad[idx] = s;
  }
  if (guard && idx < (size_t)B) {
    double s = 0.0;
    for (int cl = 0; cl < ncol; ++cl) s += guardcol[(size_t)cl + (size_t)ncol * idx];
    guard[idx] = s;
  }
}

// Column sharding with the scalar exchange: this rank's part of dot(final_state, R), dot(final_state, T)
// (src/eval_grad_discrete_adjoint.jl:27-28), one warp per control vector; final_all holds the owned columns in place.
__global__ void k_partial_dots(int N, int nic, int col0, int ncol, int B, const double* final_all, const double* target, double* dots) {
  const int b = blockIdx.x, lane = threadIdx.x;
  if (b >= B) return;
  const int N2 = 2 * N;
  double dR = 0.0, dT = 0.0;
  for (int col = col0; col < col0 + ncol; ++col) {
    const double* p = final_all + (size_t)N2 * ((size_t)col + (size_t)nic * b);
    const double* R = target + (size_t)N2 * col;
    for (int r = lane; r < N; r += 32) {
      dR += p[r] * R[r] + p[N + r] * R[N + r];
      dT += p[r] * R[N + r] - p[N + r] * R[r];
    }
  }
  for (int o = 16; o > 0; o >>= 1) { dR += __shfl_xor_sync(0xffffffffu, dR, o); dT += __shfl_xor_sync(0xffffffffu, dT, o); }
  if (lane == 0) { dots[2 * b] = dR; dots[2 * b + 1] = dT; }
}

// FP64 FMA throughput micro-benchmark (roofline denominator MEASURED_PEAKS.json does not carry).
__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3, a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
  const double m = 0.999999999, c = 1e-12;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// FP64 tensor-core (DMMA.8x8x4) throughput micro-benchmark: 8 independent accumulator chains per warp.
__global__ void k_dmma_peak(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * (threadIdx.x + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace qgd

namespace {

using namespace qgd;

int pick_el(int N) {
  int el = 1;
  while (32 * el < N) el *= 2;
  if (el > 8) throw QgdError(QGD_EUNSUPPORTED, "N_tot_levels > 256 is not supported by the warp-per-column kernels (see DESIGN.md)");
  return el;
}

void validate_and_load(qgd_handle* h, const qgd_problem_t* p) {
  if (!p) throw QgdError(QGD_EINVAL, "null problem");
  const int64_t N = p->N_tot_levels;
  if (N < 1) throw QgdError(QGD_EINVAL, "N_tot_levels must be positive");
  if (p->N_operators < 0 || p->N_operators > QGD_MAX_OPS - 1) throw QgdError(QGD_EUNSUPPORTED, "at most 8 control operators are supported");
  if (p->N_initial_conditions < 1) throw QgdError(QGD_EINVAL, "need at least one initial condition column");
  if (p->N_ess_levels > N) throw QgdError(QGD_EINVAL, "Number of essential levels cannot be greater than the total number of levels.");
  if (p->nsteps < 1) throw QgdError(QGD_EINVAL, "nsteps must be >= 1");
  if (!(p->tf > 0)) throw QgdError(QGD_EINVAL, "tf must be positive");
  auto check_sq = [&](const qgd_matrix_t& a, int64_t n, const char* what) {
    if (a.nrows != n || a.ncols != n) throw QgdError(QGD_EINVAL, std::string("Size of ") + what + " does not match the size of the system Hamiltonian.");
  };
  check_sq(p->system_sym, N, "real part of the system Hamiltonian");
  check_sq(p->system_asym, N, "imaginary part of the system Hamiltonian");
  check_sq(p->guard_subspace_projector, 2 * N, "guard subspace projector (should be twice the complex system size)");
  h->N = (int)N; h->N2 = 2 * (int)N; h->Nc = (int)p->N_operators; h->nic = (int)p->N_initial_conditions; h->Ness = (int)p->N_ess_levels;
  h->nsteps = p->nsteps; h->tf = p->tf; h->abstol = p->gmres_abstol; h->reltol = p->gmres_reltol; h->precond = p->preconditioner;
  if (h->precond < 0 || h->precond > 2) throw QgdError(QGD_EINVAL, "preconditioner_type is not an AbstractQGDPreconditioner.");
  h->col0 = 0; h->ncol = h->nic;
  pick_el(h->N);

  // dense copies + symmetry checks (src/SchrodingerProb.jl:73-91)
  const int n = h->N;
  h->Ks = mat_to_dense(p->system_sym);
  h->Ss = mat_to_dense(p->system_asym);
  std::vector<dvec> Kc(h->Nc), Sc(h->Nc);
  auto is_sym = [&](const dvec& A, double sign) {
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) if (A[i + (size_t)n * j] != sign * A[j + (size_t)n * i]) return false;
    return true;
  };
  if (!is_sym(h->Ks, 1.0)) throw QgdError(QGD_EINVAL, "Real part of system Hamiltonian is not symmetric.");
  if (!is_sym(h->Ss, -1.0)) throw QgdError(QGD_EINVAL, "Imaginary part of system Hamiltonian is not anti-symmetric.");
  for (int k = 0; k < h->Nc; ++k) {
    check_sq(p->sym_operators[k], N, "a symmetric operator");
    check_sq(p->asym_operators[k], N, "an anti-symmetric operator");
    Kc[k] = mat_to_dense(p->sym_operators[k]);
    Sc[k] = mat_to_dense(p->asym_operators[k]);
    if (!is_sym(Kc[k], 1.0)) throw QgdError(QGD_EINVAL, "Symmetric operator " + std::to_string(k + 1) + " is not symmetric.");
    if (!is_sym(Sc[k], -1.0)) throw QgdError(QGD_EINVAL, "Anti-symmetric operator " + std::to_string(k + 1) + " is not anti-symmetric.");
  }
  dvec W = mat_to_dense(p->guard_subspace_projector);

  // ---- controls
  h->ctrls.resize(h->Nc); h->ctrl_freqs.resize(h->Nc); h->ctrl_knots.resize(h->Nc);
  int off = 0;
  for (int k = 0; k < h->Nc; ++k) {
    const qgd_control_t& c = p->controls[k];
    QgdDevControl dc{};
    dc.type = c.type; dc.tf = c.tf; dc.n_carriers = (int)std::max<int64_t>(c.n_carriers, 0);
    dc.n_amp = (int)c.n_amplitudes; dc.D1 = (int)c.D1; dc.degree = (int)c.degree; dc.n_basis = (int)c.n_basis;
    if (c.type == QGD_CONTROL_GRAPE) {
      if (c.n_amplitudes < 1) throw QgdError(QGD_EINVAL, "GRAPEControl: N_amplitudes must be >= 1");
      dc.base_ncoeff = 2 * dc.n_amp;
    } else if (c.type == QGD_CONTROL_BSPLINE2) {
      if (c.D1 < 3) throw QgdError(QGD_EINVAL, "Number of coefficients per spline (D1) must be >= 3.");
      dc.base_ncoeff = 2 * dc.D1;
      dc.dtknot = c.tf / (double)(c.D1 - 2);
    } else if (c.type == QGD_CONTROL_FORTRAN_BSPLINE) {
      dc.order = dc.degree + 1;
      dc.N_knots = dc.n_basis + dc.order;
      dc.N_distinct = dc.N_knots - 2 * (dc.order - 1);
      if (dc.N_distinct < 2) throw QgdError(QGD_EINVAL, "FortranBSplineControl: too few basis functions for this degree.");
      if (dc.order > QGD_FBS_MAXORDER) throw QgdError(QGD_EINVAL, "FortranBSplineControl: pppack supports order <= 20 (bsplvb.f jmax).");
      dc.base_ncoeff = 2 * dc.n_basis;
      dvec& kn = h->ctrl_knots[k];
      for (int i = 0; i < dc.order - 1; ++i) kn.push_back(0.0);
      for (int i = 0; i < dc.N_distinct; ++i) kn.push_back((double)i / (double)(dc.N_distinct - 1));
      for (int i = 0; i < dc.order - 1; ++i) kn.push_back(1.0);
    } else if (c.type == QGD_CONTROL_HOST_TABLE) {
      if (c.n_amplitudes < 0 || c.n_carriers > 0) throw QgdError(QGD_EINVAL, "host-table control: n_amplitudes = N_coeff >= 0, no carriers");
      dc.base_ncoeff = dc.n_amp;
      h->host_controls = true;
    } else {
      throw QgdError(QGD_EUNSUPPORTED, "control type is not evaluated on the device (GRAPE, BSpline2, FortranBSpline, Carrier are); "
                                       "declare it QGD_CONTROL_HOST_TABLE and use the qgd_*_tables entry points");
    }
    if (dc.n_carriers > 0) h->ctrl_freqs[k].assign(c.carrier_freqs, c.carrier_freqs + dc.n_carriers);
    dc.ncoeff = dc.n_carriers > 0 ? dc.base_ncoeff * dc.n_carriers : dc.base_ncoeff;
    dc.offset = off;
    off += dc.ncoeff;
    h->ctrls[k] = dc;
  }
  h->P = off;

  // ---- operator blob: per-operator row-ELL over the union pattern of (K, S)
  QgdOpLayout& L = h->lay;
  std::memset(&L, 0, sizeof(L));
  L.n_ops = h->Nc + 1;
  std::vector<std::vector<std::vector<int>>> cols(L.n_ops);  // [op][row] -> columns
  for (int k = 0; k < L.n_ops; ++k) {
    const dvec& K = (k == 0) ? h->Ks : Kc[k - 1];
    const dvec& S = (k == 0) ? h->Ss : Sc[k - 1];
    cols[k].resize(n);
    int Lk = 0;
    for (int r = 0; r < n; ++r) {
      for (int c = 0; c < n; ++c)
        if (K[r + (size_t)n * c] != 0.0 || S[r + (size_t)n * c] != 0.0) cols[k][r].push_back(c);
      Lk = std::max(Lk, (int)cols[k][r].size());
    }
    L.L[k] = Lk;
  }
  int pos = 0;
  for (int k = 0; k < L.n_ops; ++k) {
    L.off_col[k] = pos; pos = align16(pos + L.L[k] * n * 4);
    L.off_vk[k] = pos; pos = align16(pos + L.L[k] * n * 8);
    L.off_vs[k] = pos; pos = align16(pos + L.L[k] * n * 8);
  }
  const int n2 = h->N2;
  std::vector<std::vector<int>> wcols(n2);
  int LW = 0;
  for (int r = 0; r < n2; ++r) {
    for (int c = 0; c < n2; ++c) if (W[r + (size_t)n2 * c] != 0.0) wcols[r].push_back(c);
    LW = std::max(LW, (int)wcols[r].size());
  }
  L.LW = LW;
  L.off_wcol = pos; pos = align16(pos + LW * n2 * 4);
  L.off_wval = pos; pos = align16(pos + LW * n2 * 8);
  for (int dir = 0; dir < 2; ++dir) { L.off_pre[dir] = pos; pos = align16(pos + (n2 + 3 * n) * 8); }
  L.bytes = align16(pos);
  h->blob.assign(L.bytes, 0);
  for (int k = 0; k < L.n_ops; ++k) {
    const dvec& K = (k == 0) ? h->Ks : Kc[k - 1];
    const dvec& S = (k == 0) ? h->Ss : Sc[k - 1];
    int* col = reinterpret_cast<int*>(h->blob.data() + L.off_col[k]);
    double* vk = reinterpret_cast<double*>(h->blob.data() + L.off_vk[k]);
    double* vs = reinterpret_cast<double*>(h->blob.data() + L.off_vs[k]);
    for (int r = 0; r < n; ++r)
      for (int s = 0; s < L.L[k]; ++s) {
        if (s < (int)cols[k][r].size()) {
          int c = cols[k][r][s];
          col[s * n + r] = c; vk[s * n + r] = K[r + (size_t)n * c]; vs[s * n + r] = S[r + (size_t)n * c];
        } else {
          col[s * n + r] = r; vk[s * n + r] = 0.0; vs[s * n + r] = 0.0;
        }
      }
  }
  {
    int* col = reinterpret_cast<int*>(h->blob.data() + L.off_wcol);
    double* val = reinterpret_cast<double*>(h->blob.data() + L.off_wval);
    for (int r = 0; r < n2; ++r)
      for (int s = 0; s < LW; ++s) {
        if (s < (int)wcols[r].size()) { col[s * n2 + r] = wcols[r][s]; val[s * n2 + r] = W[r + (size_t)n2 * wcols[r][s]]; }
        else { col[s * n2 + r] = r; val[s * n2 + r] = 0.0; }
      }
  }

  // ---- does the problem have the structure of the register-operator fast path?  diagonal drift, <= 2 entries per
  // row and control operator, diagonal guard projector; N <= 64 on one warp per column, N <= 256 on row-split groups of
  // two or four warps (qgd_fast.cuh)
  {
    bool ok = n <= 256 && h->Nc >= 1 && L.L[0] <= 1 && LW <= 1;
    for (int k = 1; k < L.n_ops && ok; ++k) ok = L.L[k] <= 2;
    if (ok && L.L[0] == 1) {
      const int* col = reinterpret_cast<const int*>(h->blob.data() + L.off_col[0]);
      const double* vs = reinterpret_cast<const double*>(h->blob.data() + L.off_vs[0]);
      for (int r = 0; r < n && ok; ++r) ok = col[r] == r && vs[r] == 0.0;
    }
    if (ok && LW == 1) {
      const int* col = reinterpret_cast<const int*>(h->blob.data() + L.off_wcol);
      for (int r = 0; r < n2 && ok; ++r) ok = col[r] == r;
    }
    h->fast_ok = ok;
    h->fast_el = n <= 32 ? 1 : 2;
    h->fast_rs = n <= 64 ? 1 : (n <= 128 ? 2 : 4);
  }
  // ---- dense problems whose level count is a multiple of 32: row-major dense copies for the tensor-core path
  // (qgd_dense.cu); "dense" = more than a quarter of the entries of some operator present
  if (!h->fast_ok && n % 32 == 0 && n <= 256) {
    bool dense = false;
    for (int k = 0; k < L.n_ops; ++k) dense = dense || 4 * L.L[k] > n;
    if (dense) {
      const size_t nn = (size_t)n * n;
      h->dense_ops.assign((size_t)L.n_ops * 2 * nn, 0.0);
      for (int k = 0; k < L.n_ops; ++k) {
        const dvec& K = (k == 0) ? h->Ks : Kc[k - 1];
        const dvec& S = (k == 0) ? h->Ss : Sc[k - 1];
        for (int r = 0; r < n; ++r)
          for (int c = 0; c < n; ++c) {  // column-major (r, c) -> row-major
            h->dense_ops[((size_t)k * 2 + 0) * nn + (size_t)r * n + c] = K[r + (size_t)n * c];
            h->dense_ops[((size_t)k * 2 + 1) * nn + (size_t)r * n + c] = S[r + (size_t)n * c];
          }
      }
    }
  }

  // ---- upload the static parts
  h->d_u0.reserve((size_t)n * h->nic * 8); h->d_v0.reserve((size_t)n * h->nic * 8);
  CUDA_CHECK(cudaMemcpy(h->d_u0.p, p->u0, (size_t)n * h->nic * 8, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(h->d_v0.p, p->v0, (size_t)n * h->nic * 8, cudaMemcpyHostToDevice));
  // controls: frequencies and knots packed in one aux buffer
  size_t aux = 0;
  for (int k = 0; k < h->Nc; ++k) aux += h->ctrl_freqs[k].size() + h->ctrl_knots[k].size();
  h->d_aux.reserve(std::max<size_t>(aux, 1) * 8);
  size_t apos = 0;
  for (int k = 0; k < h->Nc; ++k) {
    double* base = h->d_aux.as<double>();
    if (!h->ctrl_freqs[k].empty()) {
      CUDA_CHECK(cudaMemcpy(base + apos, h->ctrl_freqs[k].data(), h->ctrl_freqs[k].size() * 8, cudaMemcpyHostToDevice));
      h->ctrls[k].freqs = base + apos; apos += h->ctrl_freqs[k].size();
    }
    if (!h->ctrl_knots[k].empty()) {
      CUDA_CHECK(cudaMemcpy(base + apos, h->ctrl_knots[k].data(), h->ctrl_knots[k].size() * 8, cudaMemcpyHostToDevice));
      h->ctrls[k].knots = base + apos; apos += h->ctrl_knots[k].size();
    }
  }
  h->d_ctrls.reserve(std::max<size_t>(h->Nc, 1) * sizeof(QgdDevControl));
  if (h->Nc) CUDA_CHECK(cudaMemcpy(h->d_ctrls.p, h->ctrls.data(), h->Nc * sizeof(QgdDevControl), cudaMemcpyHostToDevice));
}

// Preconditioner factors for (nsteps, order): form_LHS_no_control (src/forward_evolution.jl:772-802;
// NOTE as written: I + sum_j (-dt)^j c_j A^j, without the 1/j!) and the two preconditioner types.
void build_preconditioner(qgd_handle* h, int order) {
  if (h->pre_key_nsteps == (int)h->nsteps && h->pre_key_order == order) return;
  const int n = h->N, n2 = h->N2, m = order / 2;
  const double dt = h->tf / (double)h->nsteps;
  if (h->precond != QGD_PRECOND_IDENTITY) {
    for (int dir = 0; dir < 2; ++dir) {
      dvec A((size_t)n2 * n2, 0.0);
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
          A[i + (size_t)n2 * j] = h->Ss[i + (size_t)n * j];
          A[i + (size_t)n2 * (j + n)] = h->Ks[i + (size_t)n * j];
          A[(i + n) + (size_t)n2 * j] = -h->Ks[i + (size_t)n * j];
          A[(i + n) + (size_t)n2 * (j + n)] = h->Ss[i + (size_t)n * j];
        }
      if (dir == 1) {
        dvec At((size_t)n2 * n2);
        for (int j = 0; j < n2; ++j) for (int i = 0; i < n2; ++i) At[i + (size_t)n2 * j] = A[j + (size_t)n2 * i];
        A.swap(At);
      }
      dvec Lm((size_t)n2 * n2, 0.0);
      for (int i = 0; i < n2; ++i) Lm[i + (size_t)n2 * i] = 1.0;
      for (int j = 1; j <= m; ++j) {
        const double coeff = ipow(-dt, j) * coefficient(j, m, m);
        dvec Aj = matpow(A, j, n2);
        for (size_t e = 0; e < Lm.size(); ++e) Lm[e] += coeff * Aj[e];
      }
      if (h->precond == QGD_PRECOND_DIAGONAL) {  // preconditioners.jl:84-126
        double* pd = reinterpret_cast<double*>(h->blob.data() + h->lay.off_pre[dir]);
        double* dg = pd; double* up = pd + n2; double* ratio = up + n; double* den = ratio + n;
        for (int i = 0; i < n2; ++i) {
          dg[i] = Lm[i + (size_t)n2 * i];
          if (dg[i] == 0.0) throw QgdError(QGD_EINVAL, "DiagonalHamiltonianPreconditioner: zero diagonal entry in the LHS");
        }
        for (int i = 0; i < n; ++i) {
          up[i] = Lm[i + (size_t)n2 * (i + n)];
          const double lo = Lm[(i + n) + (size_t)n2 * i];
          ratio[i] = lo / dg[i];
          den[i] = dg[n + i] - up[i] * ratio[i];
        }
      } else {  // LU: explicit inverse by Gauss-Jordan with partial pivoting
        dvec M = Lm, Inv((size_t)n2 * n2, 0.0);
        for (int i = 0; i < n2; ++i) Inv[i + (size_t)n2 * i] = 1.0;
        for (int k = 0; k < n2; ++k) {
          int pv = k; double mx = std::fabs(M[k + (size_t)n2 * k]);
          for (int i = k + 1; i < n2; ++i) if (std::fabs(M[i + (size_t)n2 * k]) > mx) { mx = std::fabs(M[i + (size_t)n2 * k]); pv = i; }
          if (mx == 0.0) throw QgdError(QGD_EINVAL, "LUPreconditioner: singular LHS");
          if (pv != k) for (int j = 0; j < n2; ++j) { std::swap(M[k + (size_t)n2 * j], M[pv + (size_t)n2 * j]); std::swap(Inv[k + (size_t)n2 * j], Inv[pv + (size_t)n2 * j]); }
          const double dinv = 1.0 / M[k + (size_t)n2 * k];
          for (int j = 0; j < n2; ++j) { M[k + (size_t)n2 * j] *= dinv; Inv[k + (size_t)n2 * j] *= dinv; }
          for (int i = 0; i < n2; ++i) {
            if (i == k) continue;
            const double f = M[i + (size_t)n2 * k];
            if (f == 0.0) continue;
            for (int j = 0; j < n2; ++j) { M[i + (size_t)n2 * j] -= f * M[k + (size_t)n2 * j]; Inv[i + (size_t)n2 * j] -= f * Inv[k + (size_t)n2 * j]; }
          }
        }
        h->d_minv[dir].reserve(Inv.size() * 8);
        CUDA_CHECK(cudaMemcpy(h->d_minv[dir].p, Inv.data(), Inv.size() * 8, cudaMemcpyHostToDevice));
      }
    }
  }
  h->d_blob.reserve(h->blob.size());
  CUDA_CHECK(cudaMemcpy(h->d_blob.p, h->blob.data(), h->blob.size(), cudaMemcpyHostToDevice));
  h->pre_key_nsteps = (int)h->nsteps; h->pre_key_order = order;
  h->hist_valid = false;
}

QgdDevProb make_devprob(qgd_handle* h, int order) {
  QgdDevProb d{};
  const int m = order / 2;
  d.N = h->N; d.N2 = h->N2; d.Nc = h->Nc; d.nic = h->nic; d.Ness = h->Ness;
  d.col0 = h->col0; d.ncol = h->ncol; d.m = m; d.nsteps = (int)h->nsteps; d.P = h->P; d.precond = h->precond;
  d.dt = h->tf / (double)h->nsteps; d.tf = h->tf; d.abstol = h->abstol; d.reltol = h->reltol;
  for (int j = 0; j <= m; ++j) {
    const double cj = coefficient(j, m, m);
    d.a_rhs[j] = ipow(d.dt, j) * cj;
    d.a_lhs[j] = ipow(-d.dt, j) * cj;
    d.a_tay[j] = ipow(d.dt, j) / factorial_d(j);
  }
  d.lay = h->lay;
  d.blob = h->d_blob.as<unsigned char>();
  d.minv[0] = h->d_minv[0].as<double>(); d.minv[1] = h->d_minv[1].as<double>();
  d.u0 = h->d_u0.as<double>(); d.v0 = h->d_v0.as<double>();
  d.table = h->d_table.as<double>();
  return d;
}

void check_order(int order) {
  if (order < 2 || (order & 1) || order / 2 > QGD_MAX_M)
    throw QgdError(QGD_EINVAL, "order must be even and between 2 and " + std::to_string(2 * QGD_MAX_M));
}

// control basis table for t_n = n dt, n = 0..nsteps (cached per (nsteps, m))
void ensure_table(qgd_handle* h, int m) {
  if (h->tables_from_host) return;  // qgd_*_tables: the caller's tables are already in d_table / d_cvals
  if (h->host_controls)
    throw QgdError(QGD_EUNSUPPORTED, "this problem has host-evaluated controls (QGD_CONTROL_HOST_TABLE): use the qgd_*_tables entry points");
  if (h->tab_key_nsteps == (int)h->nsteps && h->tab_key_m == m) return;
  const int Nt = (int)h->nsteps + 1;
  const size_t sz = (size_t)Nt * 2 * (m + 1) * std::max(h->P, 1);
  h->d_table.reserve(sz * 8);
  if (h->Nc > 0) {
    const int threads = 128, total = Nt * h->Nc;
    k_control_table<<<(total + threads - 1) / threads, threads, 0, h->stream>>>(
        h->d_ctrls.as<QgdDevControl>(), h->Nc, h->P, m, Nt, nullptr, 0.0, h->tf / (double)h->nsteps, h->d_table.as<double>());
    CUDA_CHECK(cudaGetLastError());
    h->stats.kernel_launches++;
  }
  h->tab_key_nsteps = (int)h->nsteps; h->tab_key_m = m;
}

void compute_cvals(qgd_handle* h, int m, int B, const double* d_pcof) {
  if (h->tables_from_host) return;
  const int Nt = (int)h->nsteps + 1;
  const size_t total = (size_t)B * Nt * 2 * (m + 1) * h->Nc;
  h->d_cvals.reserve(std::max<size_t>(total, 1) * 8);
  if (total == 0) return;
  const int threads = 256;
  k_control_values<<<(unsigned)((total + threads - 1) / threads), threads, 0, h->stream>>>(
      h->d_ctrls.as<QgdDevControl>(), h->Nc, h->P, m, Nt, h->d_table.as<double>(), d_pcof, B, h->d_cvals.as<double>());
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
}

#define QGD_DISPATCH_EL(el, NAME, ...)            \
  switch (el) {                                   \
    case 1: NAME##_1(__VA_ARGS__); break;         \
    case 2: NAME##_2(__VA_ARGS__); break;         \
    case 4: NAME##_4(__VA_ARGS__); break;         \
    default: NAME##_8(__VA_ARGS__); break;        \
  }

void h2d(qgd_handle* h, void* dst, const void* src, size_t bytes) {
  CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  h->stats.h2d_bytes += (int64_t)bytes;
}
void d2h(qgd_handle* h, void* dst, const void* src, size_t bytes) {
  CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  h->stats.d2h_bytes += (int64_t)bytes;
}
void iters_out(qgd_handle* h, int64_t* dst, DevBuf& src, size_t n) {  // device int32 -> host int64
  if (!dst) return;
  std::vector<int> tmp(n);
  d2h(h, tmp.data(), src.p, n * 4);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < n; ++i) dst[i] = tmp[i];
}

void reset_stats(qgd_handle* h) { h->stats = qgd_stats_t{}; }

// The sweeps raise a sticky device error word instead of returning wrong numbers silently (wait_segment in
// qgd_fast.cuh).  Called by every synchronising entry point after its stream synchronisation.
void check_device_error(qgd_handle* h) {
  if (!h->d_counter.cap) return;
  int err = 0;
  CUDA_CHECK(cudaMemcpyAsync(&err, h->d_counter.as<int>() + 1, 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (err != 0) {
    CUDA_CHECK(cudaMemsetAsync(h->d_counter.as<int>() + 1, 0, 4, h->stream));
    h->hist_valid = false;
    throw QgdError(QGD_ESTATE, "a sweep kernel gave up waiting for a time segment that was never published (device error word " +
                                   std::to_string(err) + "): the results of this call are not valid");
  }
}

// the control vectors the resident history belongs to (history_precomputed is only honoured for the same ones)
void remember_hist_pcof(qgd_handle* h, const double* pcof, int B) {
  if (pcof) h->hist_pcof.assign(pcof, pcof + (size_t)h->P * B);
  else h->hist_pcof.clear();
}

// ---- register-operator fast path dispatch (false: shape not built / not applicable -> generic kernels)
bool fast_applicable(const qgd_handle* h, int m) {
  return h->fast_ok && !h->opt[QGD_OPT_DISABLE_FAST] && h->precond != QGD_PRECOND_LU && m >= 1 && m <= QGD_FAST_MAX_M;
}
#define QGD_FAST_SWITCH(m, NAME, ...)                   \
  bool done_ = false;                                   \
  switch (m) {                                          \
    case 1: done_ = NAME##_m1(__VA_ARGS__); break;      \
    case 2: done_ = NAME##_m2(__VA_ARGS__); break;      \
    case 3: done_ = NAME##_m3(__VA_ARGS__); break;      \
    case 4: done_ = NAME##_m4(__VA_ARGS__); break;      \
    case 5: done_ = NAME##_m5(__VA_ARGS__); break;      \
    case 6: done_ = NAME##_m6(__VA_ARGS__); break;      \
    default: break;                                     \
  }                                                     \
  if (done_) h->stats.fast_path_launches++;             \
  return done_;
bool try_forward_fast_default(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  QGD_FAST_SWITCH(d.m, launch_forward_fast, h, d, a, h->fast_el, h->Nc)
}
bool try_forward_fast_strict(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  QGD_FAST_SWITCH(d.m, launch_forward_fast_strict, h, d, a, h->fast_el, h->Nc)
}
bool try_backward_fast_default(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  QGD_FAST_SWITCH(d.m, launch_backward_fast, h, d, a, h->fast_el, h->Nc)
}
bool try_backward_fast_strict(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  QGD_FAST_SWITCH(d.m, launch_backward_fast_strict, h, d, a, h->fast_el, h->Nc)
}
bool try_forward_fast_rs(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  QGD_FAST_SWITCH(d.m, launch_forward_fast_rs, h, d, a, h->fast_rs, h->Nc)
}
bool try_backward_fast_rs(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  QGD_FAST_SWITCH(d.m, launch_backward_fast_rs, h, d, a, h->fast_rs, h->Nc)
}
bool try_forward_fast_forced(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {  // forced solves: eval_forward!(...; forcing), eval_grad_forced
  if (!fast_applicable(h, d.m)) return false;
  if (h->fast_rs > 1) { QGD_FAST_SWITCH(d.m, launch_forward_fast_forced_rs, h, d, a, h->fast_rs, h->Nc) }
  QGD_FAST_SWITCH(d.m, launch_forward_fast_forced, h, d, a, h->fast_el, h->Nc)
}
bool try_forward_fast_team(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  QGD_FAST_SWITCH(d.m, launch_forward_fast_team, h, d, a, h->fast_el, h->Nc)
}
bool try_backward_fast_team(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  QGD_FAST_SWITCH(d.m, launch_backward_fast_team, h, d, a, h->fast_el, h->Nc)
}
// The latency team (four warps per column) pays when no more columns are in flight than the GPU has SMs
bool team_wanted(const qgd_handle* h, const SweepArgs& a) {
  if (h->opt[QGD_OPT_STRICT_MGS] || h->opt[QGD_OPT_LATENCY_TEAM] == 2) return false;
  if (h->opt[QGD_OPT_LATENCY_TEAM] == 1) return true;
  return (size_t)a.B * h->ncol <= (size_t)h->prop.multiProcessorCount;
}
// QGD_OPT_STRICT_MGS selects the strict modified Gram-Schmidt instantiation of the same sweeps
bool try_forward_fast(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  if (!fast_applicable(h, d.m)) return false;
  // 64 < N <= 256: row-split groups, default orthogonalisation only (the strict option keeps the generic kernels)
  if (h->fast_rs > 1) return !h->opt[QGD_OPT_STRICT_MGS] && try_forward_fast_rs(h, d, a);
  if (team_wanted(h, a) && try_forward_fast_team(h, d, a)) return true;
  return h->opt[QGD_OPT_STRICT_MGS] ? try_forward_fast_strict(h, d, a) : try_forward_fast_default(h, d, a);
}
bool try_backward_fast(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  if (!fast_applicable(h, d.m)) return false;
  if (h->fast_rs > 1) return !h->opt[QGD_OPT_STRICT_MGS] && try_backward_fast_rs(h, d, a);
  if (team_wanted(h, a) && try_backward_fast_team(h, d, a)) return true;
  return h->opt[QGD_OPT_STRICT_MGS] ? try_backward_fast_strict(h, d, a) : try_backward_fast_default(h, d, a);
}
// terminal condition on the register operators (default orthogonalisation only: the strict option keeps the generic kernel,
// whose Gram-Schmidt is the reference's one projection at a time)
bool try_terminal_fast(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a) {
  if (!fast_applicable(h, d.m) || h->opt[QGD_OPT_STRICT_MGS]) return false;
  const int64_t sweeps = h->stats.fast_path_launches;  // that counter is for the two sweep kernels
  const bool done = h->fast_rs > 1 ? [&]() -> bool { QGD_FAST_SWITCH(d.m, launch_terminal_fast_rs, h, d, a, h->fast_rs, h->Nc) }()
                                   : [&]() -> bool { QGD_FAST_SWITCH(d.m, launch_terminal_fast, h, d, a, h->fast_el, h->Nc) }();
  h->stats.fast_path_launches = sweeps;
  return done;
}
bool try_derivs_fast(qgd_handle* h, const QgdDevProb& d, const SweepArgs& a, double* uv, int ncols, const double* cv, int adjoint) {
  if (!fast_applicable(h, d.m) || h->fast_rs > 1) return false;
  QGD_FAST_SWITCH(d.m, launch_derivs_fast, h, d, a, h->fast_el, h->Nc, uv, ncols, cv, adjoint)
}

// forward sweep on device for B control vectors already in d_pcof; d_forcing: explicit forcing array
// [2N][m][nsteps+1][ncol][B] on the device (generic kernels), or null
void run_forward(qgd_handle* h, const double* d_pcof, int B, int order, int64_t save_every, bool want_iters,
                 const double* d_forcing = nullptr) {
  check_order(order);
  if (save_every < 1) throw QgdError(QGD_EINVAL, "saveEveryNsteps must be >= 1");
  const int m = order / 2, el = pick_el(h->N);
  build_preconditioner(h, order);
  ensure_table(h, m);
  compute_cvals(h, m, B, d_pcof);
  QgdDevProb d = make_devprob(h, order);
  SweepArgs a{};
  a.B = B; a.save_every = (int)save_every; a.nslots = 1 + (int)(h->nsteps / save_every);
  a.cvals = h->d_cvals.as<double>();
  const size_t hist_sz = (size_t)h->N2 * (m + 1) * a.nslots * h->ncol * B;
  h->d_history.reserve(hist_sz * 8);
  h->d_final.reserve((size_t)h->N2 * h->ncol * B * 8);
  if (h->nsteps % save_every != 0 || save_every != 1)  // slots that are never written must read as zero
    CUDA_CHECK(cudaMemsetAsync(h->d_history.p, 0, hist_sz * 8, h->stream));
  a.history = h->d_history.as<double>();
  a.final_state = h->d_final.as<double>();
  if (want_iters) { h->d_iters_f.reserve((size_t)h->nsteps * h->ncol * B * 4); a.iters = h->d_iters_f.as<int>(); }
  a.forcing_in = d_forcing;
  CUDA_CHECK(cudaEventRecord(h->ev[0], h->stream));
  const bool done = d_forcing ? try_forward_fast_forced(h, d, a) : (try_forward_fast(h, d, a) || try_forward_dense(h, d, a));
  if (!done) { QGD_DISPATCH_EL(el, launch_forward, h, d, a); }
  CUDA_CHECK(cudaEventRecord(h->ev[1], h->stream));
  h->hist_B = B; h->hist_order = order; h->hist_nsteps = h->nsteps; h->hist_save = save_every; h->hist_valid = d_forcing == nullptr;
}

// eval_grad_forced on the device: the unforced history of ONE control vector is resident (run_forward, save_every 1);
// P forced forward solves with zero initial state -- one per control parameter, batched as P x ncol items -- form
// their forcing on the fly from that history and the control basis table.  Leaves d psi_N / d theta in d_scratch
// [2N][ncol][P] and the guard-penalty derivative partials in d_guardcol [ncol][P].
void run_forced_gradient_solves(qgd_handle* h, int order) {
  const int m = order / 2, el = pick_el(h->N), P = h->P;
  QgdDevProb d = make_devprob(h, order);
  SweepArgs a{};
  a.B = P; a.save_every = 1; a.nslots = 1 + (int)h->nsteps;
  a.cvals = h->d_cvals.as<double>();
  a.history = nullptr;
  h->d_scratch.reserve((size_t)h->N2 * h->ncol * P * 8);
  a.final_state = h->d_scratch.as<double>();
  h->d_guardcol.reserve((size_t)h->ncol * P * 8);
  a.guardcol = h->d_guardcol.as<double>();
  a.base_history = h->d_history.as<double>();
  std::vector<int> top((size_t)P);
  for (int k = 0; k < h->Nc; ++k)
    for (int t = 0; t < h->ctrls[k].ncoeff; ++t) top[(size_t)h->ctrls[k].offset + t] = k + 1;  // blob operator 0 is the drift
  h->d_theta_op.reserve((size_t)P * 4);
  CUDA_CHECK(cudaMemcpyAsync(h->d_theta_op.p, top.data(), (size_t)P * 4, cudaMemcpyHostToDevice, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));  // `top` is a local
  a.theta_op = h->d_theta_op.as<int>();
  (void)m;
  if (!try_forward_fast_forced(h, d, a)) { QGD_DISPATCH_EL(el, launch_forward, h, d, a); }
}

// guard partials + (optionally) forcing array from the device-resident history
void run_guard(qgd_handle* h, int B, int order, bool want_forcing) {
  const int m = order / 2, el = pick_el(h->N);
  QgdDevProb d = make_devprob(h, order);
  SweepArgs a{};
  a.B = B; a.save_every = 1; a.nslots = (int)h->nsteps + 1;
  a.history = h->d_history.as<double>();
  h->d_guardcol.reserve((size_t)h->ncol * B * 8);
  a.guardcol = h->d_guardcol.as<double>();
  if (want_forcing) { h->d_forcing.reserve((size_t)h->N2 * (h->nsteps + 1) * h->ncol * B * 8); a.forcing_out = h->d_forcing.as<double>(); }
  (void)m;
  QGD_DISPATCH_EL(el, launch_guard, h, d, a);
}

// infidelity + terminal condition from d_final_all, then backward sweep and the column reduction
void run_adjoint(qgd_handle* h, int B, int order, const double* d_target, bool want_iters, bool want_lambda0,
                 double* d_grad_out, double* d_infid_out, double* d_guard_out) {
  const int m = order / 2, el = pick_el(h->N);
  QgdDevProb d = make_devprob(h, order);
  SweepArgs a{};
  a.B = B; a.save_every = 1; a.nslots = (int)h->nsteps + 1;
  a.cvals = h->d_cvals.as<double>();
  a.history = h->d_history.as<double>();
  a.final_all = h->d_final_all.as<double>();
  a.target = d_target;
  h->d_terminal.reserve((size_t)h->N2 * h->nic * B * 8);
  a.terminal_out = h->d_terminal.as<double>();
  a.terminal = h->d_terminal.as<double>();
  a.infidelity = d_infid_out;
  if (want_iters) {
    h->d_iters_t.reserve((size_t)h->nic * B * 4);
    h->d_iters_a.reserve((size_t)h->nsteps * h->ncol * B * 4);
    CUDA_CHECK(cudaMemsetAsync(h->d_iters_a.p, 0, (size_t)h->nsteps * h->ncol * B * 4, h->stream));
    a.iters_term = h->d_iters_t.as<int>();
  }
  const bool scalar_exchange = h->comm && h->opt[QGD_OPT_TERMINAL_EXCHANGE] == 1;
  if (scalar_exchange) { a.dots_in = h->d_dots.as<double>(); a.term_col0 = h->col0; a.term_ncol = h->ncol; }
  if (!try_terminal_fast(h, d, a)) {
    if (scalar_exchange || !try_terminal_dense(h, d, a)) { QGD_DISPATCH_EL(el, launch_terminal, h, d, a); }
  }
  a.iters = want_iters ? h->d_iters_a.as<int>() : nullptr;
  if (want_lambda0) {
    const size_t sz = (size_t)h->N2 * (h->nsteps + 1) * h->ncol * B * 8;
    h->d_lambda0.reserve(sz);
    CUDA_CHECK(cudaMemsetAsync(h->d_lambda0.p, 0, sz, h->stream));
    a.lambda0 = h->d_lambda0.as<double>();
  }
  h->d_gradcol.reserve((size_t)std::max(h->P, 1) * h->ncol * B * 8);
  a.gradcol = h->d_gradcol.as<double>();
  CUDA_CHECK(cudaEventRecord(h->ev[2], h->stream));
  if (!(try_backward_fast(h, d, a) || try_backward_dense(h, d, a))) { QGD_DISPATCH_EL(el, launch_backward, h, d, a); }
  CUDA_CHECK(cudaEventRecord(h->ev[3], h->stream));
  const size_t tot = std::max<size_t>((size_t)h->P * B, (size_t)B);
  k_finalize<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->P, h->ncol, B, h->d_gradcol.as<double>(),
                                                                  h->d_guardcol.as<double>(), d_grad_out, d_guard_out);
  CUDA_CHECK(cudaGetLastError());
  h->stats.kernel_launches++;
  (void)m;
}

void finish_timing(qgd_handle* h, bool fwd, bool bwd) {
  float ms = 0;
  if (fwd) { cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->stats.last_forward_ms = ms; }
  if (bwd) { cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); h->stats.last_backward_ms = ms; }
  if (fwd && bwd) { cudaEventElapsedTime(&ms, h->ev[0], h->ev[3]); h->stats.last_total_ms = ms; }
}

int guarded(const std::function<void()>& f) {
  try { f(); return QGD_OK; }
  catch (const QgdError& e) { g_err = e.what(); return e.code; }
  catch (const std::exception& e) { g_err = e.what(); return QGD_EINVAL; }
}
void require(bool ok, const char* msg) { if (!ok) throw QgdError(QGD_EINVAL, msg); }
}  // namespace

// =========================================================================================================
extern "C" {

const char* qgd_last_error(void) { return g_err.c_str(); }

int64_t qgd_control_n_coeff(const qgd_control_t* c) {
  int64_t base = 0;
  if (c->type == QGD_CONTROL_GRAPE) base = 2 * c->n_amplitudes;
  else if (c->type == QGD_CONTROL_BSPLINE2) base = 2 * c->D1;
  else if (c->type == QGD_CONTROL_FORTRAN_BSPLINE) base = 2 * c->n_basis;
  else if (c->type == QGD_CONTROL_HOST_TABLE) base = c->n_amplitudes;
  return c->n_carriers > 0 ? base * c->n_carriers : base;
}
int64_t qgd_problem_n_coeff(const qgd_problem_t* p) {
  int64_t s = 0;
  for (int64_t k = 0; k < p->N_operators; ++k) s += qgd_control_n_coeff(&p->controls[k]);
  return s;
}

int qgd_create(const qgd_problem_t* prob, int device, qgd_handle_t** out) {
  qgd_handle* h = nullptr;
  int rc = guarded([&]() {
    require(out != nullptr, "null output handle");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) throw QgdError(QGD_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0) CUDA_CHECK(cudaGetDevice(&device));
    CUDA_CHECK(cudaSetDevice(device));
    h = new qgd_handle();
    h->device = device;
    CUDA_CHECK(cudaGetDeviceProperties(&h->prop, device));
    if (h->prop.major < 10) throw QgdError(QGD_ECUDA, "device is not sm_100 class (Blackwell B200 required)");
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& ev : h->ev) CUDA_CHECK(cudaEventCreate(&ev));
    validate_and_load(h, prob);
    *out = h;
  });
  if (rc != QGD_OK && h) { qgd_destroy(h); }
  return rc;
}

int qgd_destroy(qgd_handle_t* h) {
  if (!h) return QGD_OK;
  cudaSetDevice(h->device);
  DevBuf* bufs[] = {&h->d_blob, &h->d_minv[0], &h->d_minv[1], &h->d_u0, &h->d_v0, &h->d_ctrls, &h->d_aux, &h->d_table, &h->d_pcof,
                    &h->d_cvals, &h->d_history, &h->d_final, &h->d_final_all, &h->d_terminal, &h->d_lambda0, &h->d_lamhist,
                    &h->d_gradcol, &h->d_grad, &h->d_guardcol, &h->d_guard, &h->d_infid, &h->d_iters_f, &h->d_iters_a,
                    &h->d_iters_t, &h->d_target, &h->d_forcing, &h->d_V, &h->d_H, &h->d_scratch, &h->d_counter, &h->d_progress, &h->d_carry,
                    &h->d_theta_op, &h->d_dense, &h->d_comb, &h->d_dense_ws, &h->d_pack, &h->d_dots};
  if (h->comm) { cudaStreamSynchronize(h->stream); qgd_nccl::destroy(h->comm); h->comm = nullptr; }
  for (DevBuf* b : bufs) b->release();
  if (h->l2_carved) cudaCtxResetPersistingL2Cache();  // hand the persisting L2 lines of the workspace window back
  for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return QGD_OK;
}

// Setting a knob to the value it already has is a no-op: a binding may re-sync the three mutable fields of the
// reference's SchrodingerProb before every call without invalidating the resident history.
int qgd_set_nsteps(qgd_handle_t* h, int64_t nsteps) {
  return guarded([&]() {
    require(h && nsteps >= 1, "nsteps must be >= 1");
    if (h->nsteps == nsteps) return;
    h->nsteps = nsteps; h->hist_valid = false;
  });
}
int qgd_set_gmres_tolerances(qgd_handle_t* h, double abstol, double reltol) {
  return guarded([&]() {
    require(h != nullptr, "null handle");
    if (h->abstol == abstol && h->reltol == reltol) return;
    h->abstol = abstol; h->reltol = reltol; h->hist_valid = false;
  });
}
int qgd_set_option(qgd_handle_t* h, int32_t key, int64_t value) {
  return guarded([&]() {
    require(h != nullptr, "null handle");
    switch (key) {
      case QGD_OPT_STRICT_MGS: case QGD_OPT_DISABLE_FAST: case QGD_OPT_DISABLE_DENSE_SWEEP: case QGD_OPT_DISABLE_DENSE_DMMA:
      case QGD_OPT_DISABLE_TMEM: case QGD_OPT_L2_PERSIST:
        require(value == 0 || value == 1, "option value must be 0 or 1"); break;
      case QGD_OPT_DENSE_TERMINAL: require(value >= 0 && value <= 2, "QGD_OPT_DENSE_TERMINAL must be 0, 1 or 2"); break;
      case QGD_OPT_SEG_STEPS: require(value >= 0 && value <= (1 << 30), "QGD_OPT_SEG_STEPS must be >= 0"); break;
      case QGD_OPT_LATENCY_WARPS: require(value >= 0 && value <= QGD_WARPS_PER_CTA, "QGD_OPT_LATENCY_WARPS must be 0..8"); break;
      case QGD_OPT_TERMINAL_EXCHANGE: require(value == 0 || value == 1, "QGD_OPT_TERMINAL_EXCHANGE must be 0 or 1"); break;
      case QGD_OPT_LATENCY_TEAM: require(value >= 0 && value <= 2, "QGD_OPT_LATENCY_TEAM must be 0, 1 or 2"); break;
      default: throw QgdError(QGD_EINVAL, "unknown option key " + std::to_string(key));
    }
    if (h->opt[key] != value) { h->opt[key] = value; h->hist_valid = false; }
  });
}
int qgd_get_option(qgd_handle_t* h, int32_t key, int64_t* value) {
  return guarded([&]() {
    require(h && value && key >= 1 && key <= QGD_OPT_MAX, "bad arguments");
    *value = h->opt[key];
  });
}
int qgd_set_column_shard(qgd_handle_t* h, int64_t col_begin, int64_t col_count) {
  return guarded([&]() {
    require(h && col_begin >= 0 && col_count >= 1 && col_begin + col_count <= h->nic, "column shard out of range");
    h->col0 = (int)col_begin; h->ncol = (int)col_count; h->hist_valid = false;
  });
}

int qgd_eval_forward(qgd_handle_t* h, const double* pcof, int64_t n_batch, int32_t order, int64_t save_every, double* history,
                     double* final_state, int64_t* gmres_iters) {
  return guarded([&]() {
    require(h && pcof && n_batch >= 1, "bad arguments");
    CUDA_CHECK(cudaSetDevice(h->device));
    reset_stats(h);
    const int B = (int)n_batch, m = order / 2;
    h->d_pcof.reserve((size_t)std::max(h->P, 1) * B * 8);
    h2d(h, h->d_pcof.p, pcof, (size_t)h->P * B * 8);
    run_forward(h, h->d_pcof.as<double>(), B, order, save_every, gmres_iters != nullptr);
    remember_hist_pcof(h, pcof, B);
    const int nslots = 1 + (int)(h->nsteps / save_every);
    if (history) d2h(h, history, h->d_history.p, (size_t)h->N2 * (m + 1) * nslots * h->ncol * B * 8);
    if (final_state) d2h(h, final_state, h->d_final.p, (size_t)h->N2 * h->ncol * B * 8);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    iters_out(h, gmres_iters, h->d_iters_f, (size_t)h->nsteps * h->ncol * B);
    finish_timing(h, true, false);
  });
}

int qgd_eval_forward_async(qgd_handle_t* h, const double* pcof, int64_t n_batch, int32_t order, int64_t save_every, int32_t want_iters) {
  return guarded([&]() {
    require(h && pcof && n_batch >= 1, "bad arguments");
    CUDA_CHECK(cudaSetDevice(h->device));
    reset_stats(h);
    const int B = (int)n_batch;
    h->pend_B = 0;
    h->d_pcof.reserve((size_t)std::max(h->P, 1) * B * 8);
    h2d(h, h->d_pcof.p, pcof, (size_t)h->P * B * 8);  // pageable source: staged before the call returns
    run_forward(h, h->d_pcof.as<double>(), B, order, save_every, want_iters != 0);
    remember_hist_pcof(h, pcof, B);
    h->pend_B = B; h->pend_order = order; h->pend_save = save_every; h->pend_iters = want_iters;
  });
}
int qgd_eval_forward_collect(qgd_handle_t* h, double* history, double* final_state, int64_t* gmres_iters) {
  return guarded([&]() {
    require(h != nullptr, "null handle");
    if (h->pend_B < 1) throw QgdError(QGD_ESTATE, "qgd_eval_forward_collect without a pending qgd_eval_forward_async");
    if (gmres_iters && !h->pend_iters) throw QgdError(QGD_EINVAL, "iteration counts were not requested from qgd_eval_forward_async");
    CUDA_CHECK(cudaSetDevice(h->device));
    const int B = h->pend_B, m = h->pend_order / 2;
    const int nslots = 1 + (int)(h->nsteps / h->pend_save);
    h->pend_B = 0;
    if (history) d2h(h, history, h->d_history.p, (size_t)h->N2 * (m + 1) * nslots * h->ncol * B * 8);
    if (final_state) d2h(h, final_state, h->d_final.p, (size_t)h->N2 * h->ncol * B * 8);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    iters_out(h, gmres_iters, h->d_iters_f, (size_t)h->nsteps * h->ncol * B);
    finish_timing(h, true, false);
  });
}

int qgd_eval_forward_forced(qgd_handle_t* h, const double* pcof, int64_t n_batch, int32_t order, int64_t save_every,
                            const double* forcing, double* history, double* final_state, int64_t* gmres_iters) {
  return guarded([&]() {
    require(h && pcof && forcing && n_batch >= 1, "bad arguments");
    CUDA_CHECK(cudaSetDevice(h->device));
    reset_stats(h);
    const int B = (int)n_batch, m = order / 2;
    check_order(order);
    h->d_pcof.reserve((size_t)std::max(h->P, 1) * B * 8);
    h2d(h, h->d_pcof.p, pcof, (size_t)h->P * B * 8);
    const size_t fbytes = (size_t)h->N2 * m * (h->nsteps + 1) * h->ncol * B * 8;
    h->d_forcing.reserve(fbytes);
    h2d(h, h->d_forcing.p, forcing, fbytes);
    run_forward(h, h->d_pcof.as<double>(), B, order, save_every, gmres_iters != nullptr, h->d_forcing.as<double>());
    const int nslots = 1 + (int)(h->nsteps / save_every);
    if (history) d2h(h, history, h->d_history.p, (size_t)h->N2 * (m + 1) * nslots * h->ncol * B * 8);
    if (final_state) d2h(h, final_state, h->d_final.p, (size_t)h->N2 * h->ncol * B * 8);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    iters_out(h, gmres_iters, h->d_iters_f, (size_t)h->nsteps * h->ncol * B);
    finish_timing(h, true, false);
  });
}

int qgd_eval_grad_forced(qgd_handle_t* h, const double* pcof, const double* target, int32_t order, double* grad) {
  return guarded([&]() {
    require(h && pcof && target && grad, "bad arguments");
    if (h->ncol != h->nic) throw QgdError(QGD_EUNSUPPORTED, "qgd_eval_grad_forced needs all columns on one handle");
    CUDA_CHECK(cudaSetDevice(h->device));
    reset_stats(h);
    const int P = h->P, N = h->N, N2 = h->N2, nic = h->nic;
    h->d_pcof.reserve((size_t)std::max(P, 1) * 8);
    h2d(h, h->d_pcof.p, pcof, (size_t)P * 8);
    run_forward(h, h->d_pcof.as<double>(), 1, order, 1, false);  // history, d_final = psi_N (eval_grad_forced.jl:54-56)
    if (P == 0) { CUDA_CHECK(cudaStreamSynchronize(h->stream)); return; }
    run_forced_gradient_solves(h, order);
    dvec fs((size_t)N2 * nic), dfs((size_t)N2 * nic * P), gc((size_t)nic * P);
    d2h(h, fs.data(), h->d_final.p, fs.size() * 8);
    d2h(h, dfs.data(), h->d_scratch.p, dfs.size() * 8);
    d2h(h, gc.data(), h->d_guardcol.p, gc.size() * 8);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    // d infidelity / d theta = -(2 / N_ess^2) (<psi,R><dpsi,R> + <psi,T><dpsi,T>), T = [R_im; -R_re]  (:134-147)
    auto dots = [&](const double* psi, double& dR, double& dT) {
      dR = 0.0; dT = 0.0;
      for (int c = 0; c < nic; ++c)
        for (int r = 0; r < N; ++r) {
          const double pu = psi[r + (size_t)N2 * c], pv = psi[N + r + (size_t)N2 * c];
          const double Ru = target[r + (size_t)N2 * c], Rv = target[N + r + (size_t)N2 * c];
          dR += pu * Ru + pv * Rv;
          dT += pu * Rv - pv * Ru;
        }
    };
    double dRf, dTf;
    dots(fs.data(), dRf, dTf);
    for (int t = 0; t < P; ++t) {
      double dRp, dTp;
      dots(dfs.data() + (size_t)N2 * nic * t, dRp, dTp);
      double g = -(2.0 / ((double)h->Ness * h->Ness)) * (dRf * dRp + dTf * dTp);
      for (int c = 0; c < nic; ++c) g += gc[(size_t)c + (size_t)nic * t];
      grad[t] = g;
    }
    h->hist_valid = true;
    finish_timing(h, true, false);
  });
}

// ---- host-evaluated controls: the caller's tables replace k_control_table / k_control_values --------------------
namespace {
struct HostTables {  // uploads the tables and routes ensure_table / compute_cvals around the device control kernels
  qgd_handle* h;
  HostTables(qgd_handle* h_, int B, int m, const double* cvals, const double* table) : h(h_) {
    const size_t Nt = (size_t)h->nsteps + 1;
    const size_t cbytes = (size_t)B * Nt * 2 * (m + 1) * std::max(h->Nc, 1) * 8;
    h->d_cvals.reserve(cbytes);
    if (h->Nc > 0) h2d(h, h->d_cvals.p, cvals, (size_t)B * Nt * 2 * (m + 1) * h->Nc * 8);
    const size_t tbytes = Nt * 2 * (m + 1) * (size_t)std::max(h->P, 1) * 8;
    h->d_table.reserve(tbytes);
    if (table && h->P > 0) h2d(h, h->d_table.p, table, Nt * 2 * (m + 1) * (size_t)h->P * 8);
    h->tab_key_nsteps = -1; h->tab_key_m = -1;  // the cached device-evaluated table is gone
    h->tables_from_host = true;
  }
  ~HostTables() { h->tables_from_host = false; h->hist_valid = false; }
};
}  // namespace

int qgd_eval_forward_tables(qgd_handle_t* h, int64_t n_batch, int32_t order, int64_t save_every, const double* cvals,
                            double* history, double* final_state, int64_t* gmres_iters) {
  return guarded([&]() {
    require(h && cvals && n_batch >= 1, "bad arguments");
    CUDA_CHECK(cudaSetDevice(h->device));
    reset_stats(h);
    check_order(order);
    const int B = (int)n_batch, m = order / 2;
    HostTables guard(h, B, m, cvals, nullptr);
    run_forward(h, nullptr, B, order, save_every, gmres_iters != nullptr);
    const int nslots = 1 + (int)(h->nsteps / save_every);
    if (history) d2h(h, history, h->d_history.p, (size_t)h->N2 * (m + 1) * nslots * h->ncol * B * 8);
    if (final_state) d2h(h, final_state, h->d_final.p, (size_t)h->N2 * h->ncol * B * 8);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    iters_out(h, gmres_iters, h->d_iters_f, (size_t)h->nsteps * h->ncol * B);
    finish_timing(h, true, false);
  });
}

int qgd_discrete_adjoint_tables(qgd_handle_t* h, int64_t n_batch, int32_t order, const double* cvals, const double* table,
                                const double* target, double* grad, double* infidelity, double* guard_penalty) {
  return guarded([&]() {
    require(h && cvals && table && target && n_batch >= 1, "bad arguments");
    if (h->ncol != h->nic) throw QgdError(QGD_ESTATE, "column-sharded handle: the table entry points need all columns");
    CUDA_CHECK(cudaSetDevice(h->device));
    reset_stats(h);
    check_order(order);
    const int B = (int)n_batch, m = order / 2;
    HostTables guard(h, B, m, cvals, table);
    h->d_target.reserve((size_t)h->N2 * h->nic * 8);
    h2d(h, h->d_target.p, target, (size_t)h->N2 * h->nic * 8);
    run_forward(h, nullptr, B, order, 1, false);
    run_guard(h, B, order, false);
    h->d_final_all.reserve((size_t)h->N2 * h->nic * B * 8);
    CUDA_CHECK(cudaMemcpyAsync(h->d_final_all.p, h->d_final.p, (size_t)h->N2 * h->nic * B * 8, cudaMemcpyDeviceToDevice, h->stream));
    h->d_infid.reserve((size_t)B * 8); h->d_guard.reserve((size_t)B * 8);
    h->d_grad.reserve((size_t)std::max(h->P, 1) * B * 8);
    run_adjoint(h, B, order, h->d_target.as<double>(), false, false, h->d_grad.as<double>(), h->d_infid.as<double>(),
                h->d_guard.as<double>());
    if (grad) d2h(h, grad, h->d_grad.p, (size_t)h->P * B * 8);
    if (infidelity) d2h(h, infidelity, h->d_infid.p, (size_t)B * 8);
    if (guard_penalty) d2h(h, guard_penalty, h->d_guard.p, (size_t)B * 8);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    finish_timing(h, true, true);
  });
}

int qgd_adjoint_phase1(qgd_handle_t* h, const double* pcof, int64_t n_batch, int32_t order, double* final_state_local,
                       double* guard_local) {
  return guarded([&]() {
    require(h && pcof && n_batch >= 1, "bad arguments");
    if (h->comm) throw QgdError(QGD_ESTATE, "this handle has a communicator: qgd_discrete_adjoint does both phases and the exchanges on the device");
    CUDA_CHECK(cudaSetDevice(h->device));
    reset_stats(h);
    const int B = (int)n_batch;
    h->d_pcof.reserve((size_t)std::max(h->P, 1) * B * 8);
    h2d(h, h->d_pcof.p, pcof, (size_t)h->P * B * 8);
    run_forward(h, h->d_pcof.as<double>(), B, order, 1, false);
    remember_hist_pcof(h, pcof, B);  // the history stays resident: a following qgd_discrete_adjoint(history_precomputed) may reuse it
    run_guard(h, B, order, false);
    h->d_guard.reserve((size_t)B * 8);
    h->d_grad.reserve((size_t)std::max(h->P, 1) * B * 8);
    h->d_gradcol.reserve((size_t)std::max(h->P, 1) * h->ncol * B * 8);
    CUDA_CHECK(cudaMemsetAsync(h->d_gradcol.p, 0, (size_t)std::max(h->P, 1) * h->ncol * B * 8, h->stream));
    k_finalize<<<(unsigned)(((size_t)std::max(h->P, 1) * B + 255) / 256), 256, 0, h->stream>>>(
        h->P, h->ncol, B, h->d_gradcol.as<double>(), h->d_guardcol.as<double>(), h->d_grad.as<double>(), h->d_guard.as<double>());
    CUDA_CHECK(cudaGetLastError());
    h->stats.kernel_launches++;
    if (final_state_local) d2h(h, final_state_local, h->d_final.p, (size_t)h->N2 * h->ncol * B * 8);
    if (guard_local) d2h(h, guard_local, h->d_guard.p, (size_t)B * 8);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    h->phase_B = B; h->phase_order = order;
    finish_timing(h, true, false);
  });
}

int qgd_adjoint_phase2(qgd_handle_t* h, const double* target, const double* final_state_all, double* grad_local, double* infidelity) {
  return guarded([&]() {
    require(h && target && final_state_all, "bad arguments");
    if (h->comm) throw QgdError(QGD_ESTATE, "this handle has a communicator: qgd_discrete_adjoint does both phases and the exchanges on the device");
    if (!h->hist_valid || h->phase_B < 1) throw QgdError(QGD_ESTATE, "qgd_adjoint_phase2 without a preceding qgd_adjoint_phase1");
    CUDA_CHECK(cudaSetDevice(h->device));
    const int B = h->phase_B, order = h->phase_order;
    h->d_target.reserve((size_t)h->N2 * h->nic * 8);
    h2d(h, h->d_target.p, target, (size_t)h->N2 * h->nic * 8);
    h->d_final_all.reserve((size_t)h->N2 * h->nic * B * 8);
    h2d(h, h->d_final_all.p, final_state_all, (size_t)h->N2 * h->nic * B * 8);
    h->d_infid.reserve((size_t)B * 8);
    h->d_grad.reserve((size_t)std::max(h->P, 1) * B * 8);
    run_adjoint(h, B, order, h->d_target.as<double>(), false, false, h->d_grad.as<double>(), h->d_infid.as<double>(), nullptr);
    if (grad_local) d2h(h, grad_local, h->d_grad.p, (size_t)h->P * B * 8);
    if (infidelity) d2h(h, infidelity, h->d_infid.p, (size_t)B * 8);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    finish_timing(h, false, true);
  });
}

// ---- one gradient evaluation in stages, so that several handles (column shards of ONE evaluation on several GPUs)
// can be driven in lock step with the two NCCL exchanges grouped between the stages ---------------------------------
namespace {
struct AdjCall {
  const double* d_pcof; int B; int order; const double* d_target;
  bool precomputed, want_iters, want_lambda0, want_forcing;
  double *grad_out, *infid_out, *guard_out;  // device; with a communicator: the packed buffer, copied out after the all-reduce
};

// forward sweep (unless the history is resident), guard partials, own final states placed into the all-columns array
void adj_stage1(qgd_handle* h, const AdjCall& c) {
  const int B = c.B, m = c.order / 2;
  if (c.precomputed) {
    build_preconditioner(h, c.order);
    ensure_table(h, m);
    compute_cvals(h, m, B, c.d_pcof);
    CUDA_CHECK(cudaEventRecord(h->ev[0], h->stream));
    CUDA_CHECK(cudaEventRecord(h->ev[1], h->stream));
  } else {
    run_forward(h, c.d_pcof, B, c.order, 1, c.want_iters);
  }
  run_guard(h, B, c.order, c.want_forcing);
  const size_t all_bytes = (size_t)h->N2 * h->nic * B * 8;
  h->d_final_all.reserve(all_bytes);
  if (!h->comm) {  // all columns are local: the final states are the terminal kernel's input as they are
    CUDA_CHECK(cudaMemcpyAsync(h->d_final_all.p, h->d_final.p, all_bytes, cudaMemcpyDeviceToDevice, h->stream));
    return;
  }
  // column shard: [2N][ncol][B] -> the owned columns of [2N][nic][B], zero elsewhere (x + 0 = x: the all-reduce that
  // follows IS the all-gather, for uneven shards too, and is bitwise deterministic)
  CUDA_CHECK(cudaMemsetAsync(h->d_final_all.p, 0, all_bytes, h->stream));
  CUDA_CHECK(cudaMemcpy2DAsync(h->d_final_all.as<double>() + (size_t)h->N2 * h->col0, (size_t)h->N2 * h->nic * 8, h->d_final.p,
                               (size_t)h->N2 * h->ncol * 8, (size_t)h->N2 * h->ncol * 8, (size_t)B, cudaMemcpyDeviceToDevice, h->stream));
  if (h->opt[QGD_OPT_TERMINAL_EXCHANGE] == 1) {
    h->d_dots.reserve((size_t)2 * B * 8);
    k_partial_dots<<<B, 32, 0, h->stream>>>(h->N, h->nic, h->col0, h->ncol, B, h->d_final_all.as<double>(), c.d_target, h->d_dots.as<double>());
    CUDA_CHECK(cudaGetLastError());
    h->stats.kernel_launches++;
  }
}
// exchange 1 (between the sweeps): the final states of all columns -- or only the two inner products per control vector
void adj_exchange1(qgd_handle* h, const AdjCall& c) {
  if (!h->comm) return;
  if (h->opt[QGD_OPT_TERMINAL_EXCHANGE] == 1) qgd_nccl::allreduce_sum(h->comm, h->d_dots.as<double>(), (size_t)2 * c.B, h->stream);
  else qgd_nccl::allreduce_sum(h->comm, h->d_final_all.as<double>(), (size_t)h->N2 * h->nic * c.B, h->stream);
  h->stats.collectives++;
}
void adj_stage2(qgd_handle* h, const AdjCall& c) {
  double *g = c.grad_out, *gu = c.guard_out;
  if (h->comm) {  // [grad P*B | guard B] contiguous: ONE all-reduce at the end of the evaluation
    h->d_pack.reserve(((size_t)std::max(h->P, 1) * c.B + c.B) * 8);
    g = h->d_pack.as<double>(); gu = g + (size_t)h->P * c.B;
  }
  run_adjoint(h, c.B, c.order, c.d_target, c.want_iters, c.want_lambda0, g, c.infid_out, gu);
}
void adj_exchange2(qgd_handle* h, const AdjCall& c) {
  if (!h->comm) return;
  qgd_nccl::allreduce_sum(h->comm, h->d_pack.as<double>(), (size_t)h->P * c.B + c.B, h->stream);
  h->stats.collectives++;
}
void adj_stage3(qgd_handle* h, const AdjCall& c) {
  if (!h->comm) return;
  const double* g = h->d_pack.as<double>();
  if (c.grad_out && h->P > 0) CUDA_CHECK(cudaMemcpyAsync(c.grad_out, g, (size_t)h->P * c.B * 8, cudaMemcpyDeviceToDevice, h->stream));
  if (c.guard_out) CUDA_CHECK(cudaMemcpyAsync(c.guard_out, g + (size_t)h->P * c.B, (size_t)c.B * 8, cudaMemcpyDeviceToDevice, h->stream));
}
void adj_all_stages(qgd_handle* h, const AdjCall& c) {
  adj_stage1(h, c); adj_exchange1(h, c); adj_stage2(h, c); adj_exchange2(h, c); adj_stage3(h, c);
}
void require_all_columns_or_comm(const qgd_handle* h) {
  if (h->ncol != h->nic && !h->comm)
    throw QgdError(QGD_ESTATE, "column-sharded handle without a communicator: attach one (qgd_comm_init_rank) or use qgd_adjoint_phase1/phase2");
}
}  // namespace

int qgd_discrete_adjoint(qgd_handle_t* h, const double* pcof, int64_t n_batch, const double* target, int32_t order,
                         int32_t history_precomputed, double* grad, double* infidelity, double* guard_penalty, double* history,
                         double* lambda_history, double* adjoint_forcing, int64_t* iters_fwd, int64_t* iters_adj, int64_t* iters_term) {
  return guarded([&]() {
    require(h && pcof && target && n_batch >= 1, "bad arguments");
    require_all_columns_or_comm(h);
    CUDA_CHECK(cudaSetDevice(h->device));
    reset_stats(h);
    check_order(order);
    const int B = (int)n_batch, m = order / 2;
    const size_t Nt = (size_t)h->nsteps + 1;
    h->d_pcof.reserve((size_t)std::max(h->P, 1) * B * 8);
    h2d(h, h->d_pcof.p, pcof, (size_t)h->P * B * 8);
    h->d_target.reserve((size_t)h->N2 * h->nic * 8);
    h2d(h, h->d_target.p, target, (size_t)h->N2 * h->nic * 8);
    const bool want_iters = iters_fwd || iters_adj || iters_term;
    if (history_precomputed) {
      if (!(h->hist_valid && h->hist_B == B && h->hist_order == order && h->hist_nsteps == h->nsteps && h->hist_save == 1))
        throw QgdError(QGD_ESTATE, "history_precomputed: no matching history is resident on the device (call qgd_eval_forward with "
                                   "saveEveryNsteps=1, the same batch and order first)");
      // The reference combines whatever history array the caller passes with the new pcof (eval_grad_discrete_adjoint.jl:
      // 118-140); here the history stays on the device, so it must be the one of these very control vectors.
      if (h->hist_pcof.size() != (size_t)h->P * B || !std::equal(h->hist_pcof.begin(), h->hist_pcof.end(), pcof))
        throw QgdError(QGD_ESTATE, "history_precomputed: the resident history was computed for different control vectors");
    }
    h->d_infid.reserve((size_t)B * 8); h->d_guard.reserve((size_t)B * 8);
    h->d_grad.reserve((size_t)std::max(h->P, 1) * B * 8);
    AdjCall c{h->d_pcof.as<double>(), B, order, h->d_target.as<double>(), history_precomputed != 0, want_iters, lambda_history != nullptr,
              adjoint_forcing != nullptr, h->d_grad.as<double>(), h->d_infid.as<double>(), h->d_guard.as<double>()};
    adj_all_stages(h, c);
    if (!history_precomputed) remember_hist_pcof(h, pcof, B);
    if (lambda_history) {
      const int el = pick_el(h->N);
      const size_t sz = (size_t)h->N2 * (m + 1) * Nt * h->ncol * B * 8;
      h->d_lamhist.reserve(sz);
      QgdDevProb d = make_devprob(h, order);
      SweepArgs a{};
      a.B = B; a.cvals = h->d_cvals.as<double>(); a.lambda0 = h->d_lambda0.as<double>();
      QGD_DISPATCH_EL(el, launch_lambda_columns, h, d, a, h->d_lamhist.as<double>());
      d2h(h, lambda_history, h->d_lamhist.p, sz);
    }
    if (grad) d2h(h, grad, h->d_grad.p, (size_t)h->P * B * 8);
    if (infidelity) d2h(h, infidelity, h->d_infid.p, (size_t)B * 8);
    if (guard_penalty) d2h(h, guard_penalty, h->d_guard.p, (size_t)B * 8);
    if (history) d2h(h, history, h->d_history.p, (size_t)h->N2 * (m + 1) * Nt * h->ncol * B * 8);
    if (adjoint_forcing) d2h(h, adjoint_forcing, h->d_forcing.p, (size_t)h->N2 * Nt * h->ncol * B * 8);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
    if (!history_precomputed) iters_out(h, iters_fwd, h->d_iters_f, (size_t)h->nsteps * h->ncol * B);
    iters_out(h, iters_adj, h->d_iters_a, (size_t)h->nsteps * h->ncol * B);
    iters_out(h, iters_term, h->d_iters_t, (size_t)h->nic * B);
    finish_timing(h, true, true);
  });
}

int qgd_discrete_adjoint_device(qgd_handle_t* h, const double* d_pcof, int64_t n_batch, const double* d_target, int32_t order,
                                double* d_grad, double* d_infidelity, double* d_guard_penalty, void* stream) {
  return guarded([&]() {
    require(h && d_pcof && d_target && d_grad && d_infidelity && d_guard_penalty && n_batch >= 1, "bad arguments");
    require_all_columns_or_comm(h);
    CUDA_CHECK(cudaSetDevice(h->device));
    reset_stats(h);
    check_order(order);
    cudaStream_t own = h->stream;
    if (stream) h->stream = reinterpret_cast<cudaStream_t>(stream);
    try {
      AdjCall c{d_pcof, (int)n_batch, order, d_target, false, false, false, false, d_grad, d_infidelity, d_guard_penalty};
      adj_all_stages(h, c);
      remember_hist_pcof(h, nullptr, (int)n_batch);  // the control vectors never passed through the host: not reusable by history_precomputed
    } catch (...) { h->stream = own; throw; }
    h->stream = own;
  });
}

// Synchronise the handle's stream (or `stream`) and report the sticky device error word of the sweeps: the way to check an
// asynchronous qgd_discrete_adjoint_device call.
int qgd_synchronize(qgd_handle_t* h, void* stream) {
  return guarded([&]() {
    require(h != nullptr, "null handle");
    CUDA_CHECK(cudaSetDevice(h->device));
    if (stream) CUDA_CHECK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    check_device_error(h);
  });
}

// ---- multi-GPU, one process per GPU: attach an NCCL communicator to the handle and shard the columns ----------------
int qgd_comm_set_nccl_library(const char* path) {
  return guarded([&]() { qgd_nccl::set_library(path); });
}
int qgd_comm_get_unique_id(unsigned char* id128) {
  return guarded([&]() { require(id128 != nullptr, "bad arguments"); qgd_nccl::unique_id(id128); });
}
int qgd_comm_init_rank(qgd_handle_t* h, int32_t n_ranks, int32_t rank, const unsigned char* id128) {
  return guarded([&]() {
    require(h && id128 && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad arguments");
    if (n_ranks > h->nic) throw QgdError(QGD_EINVAL, "more ranks than initial-condition columns: shard control vectors over the extra GPUs instead");
    if (h->comm) throw QgdError(QGD_ESTATE, "the handle already has a communicator");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->comm = qgd_nccl::init_rank(n_ranks, rank, id128);
    h->comm_rank = rank; h->comm_size = n_ranks;
    const int c0 = (int)((int64_t)rank * h->nic / n_ranks), c1 = (int)((int64_t)(rank + 1) * h->nic / n_ranks);
    h->col0 = c0; h->ncol = c1 - c0; h->hist_valid = false;
  });
}
int qgd_comm_finalize(qgd_handle_t* h) {
  return guarded([&]() {
    require(h != nullptr, "null handle");
    if (!h->comm) return;
    CUDA_CHECK(cudaSetDevice(h->device));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    qgd_nccl::destroy(h->comm);
    h->comm = nullptr; h->comm_rank = 0; h->comm_size = 1;
    h->col0 = 0; h->ncol = h->nic; h->hist_valid = false;
  });
}

// ---- multi-GPU, one process driving n GPUs -------------------------------------------------------------------------------
struct qgd_multi {
  std::vector<qgd_handle*> hs;
  std::vector<int> devices;
};

int qgd_init_multi_gpu(const qgd_problem_t* prob, int32_t n_gpus, const int32_t* devices, qgd_multi_t** out) {
  qgd_multi* mg = nullptr;
  int rc = guarded([&]() {
    require(prob && out && n_gpus >= 1, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw QgdError(QGD_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    if (n_gpus > ndev) throw QgdError(QGD_EINVAL, "qgd_init_multi_gpu: " + std::to_string(n_gpus) + " GPUs requested, " + std::to_string(ndev) + " visible");
    mg = new qgd_multi();
    for (int i = 0; i < n_gpus; ++i) mg->devices.push_back(devices ? devices[i] : i);
    for (int i = 0; i < n_gpus; ++i) {
      qgd_handle_t* h = nullptr;
      const int rc1 = qgd_create(prob, mg->devices[i], &h);
      if (rc1 != QGD_OK) throw QgdError(rc1, g_err);
      mg->hs.push_back(h);
    }
    if (n_gpus > 1) {
      std::vector<void*> comms((size_t)n_gpus, nullptr);
      qgd_nccl::init_all(comms.data(), n_gpus, mg->devices.data());
      for (int i = 0; i < n_gpus; ++i) { mg->hs[i]->comm = comms[i]; mg->hs[i]->comm_rank = i; mg->hs[i]->comm_size = n_gpus; }
    }
    *out = mg;
  });
  if (rc != QGD_OK && mg) { std::string keep = g_err; qgd_multi_destroy(mg); g_err = keep; }
  return rc;
}
int qgd_multi_n_gpus(qgd_multi_t* mg) { return mg ? (int)mg->hs.size() : 0; }
int qgd_multi_handle(qgd_multi_t* mg, int32_t i, qgd_handle_t** out) {
  return guarded([&]() { require(mg && out && i >= 0 && i < (int)mg->hs.size(), "bad arguments"); *out = mg->hs[i]; });
}
int qgd_multi_set_nsteps(qgd_multi_t* mg, int64_t nsteps) {
  if (!mg) return QGD_EINVAL;
  for (qgd_handle* h : mg->hs) { const int rc = qgd_set_nsteps(h, nsteps); if (rc) return rc; }
  return QGD_OK;
}
int qgd_multi_set_gmres_tolerances(qgd_multi_t* mg, double abstol, double reltol) {
  if (!mg) return QGD_EINVAL;
  for (qgd_handle* h : mg->hs) { const int rc = qgd_set_gmres_tolerances(h, abstol, reltol); if (rc) return rc; }
  return QGD_OK;
}
int qgd_multi_destroy(qgd_multi_t* mg) {
  if (!mg) return QGD_OK;
  for (qgd_handle* h : mg->hs) qgd_destroy(h);
  delete mg;
  return QGD_OK;
}

int qgd_multi_discrete_adjoint(qgd_multi_t* mg, const double* pcof, int64_t n_batch, const double* target, int32_t order,
                               int32_t shard, double* grad, double* infidelity, double* guard_penalty) {
  return guarded([&]() {
    require(mg && pcof && target && n_batch >= 1, "bad arguments");
    require(shard == QGD_SHARD_COLUMNS || shard == QGD_SHARD_CONTROL_VECTORS, "shard must be QGD_SHARD_COLUMNS or QGD_SHARD_CONTROL_VECTORS");
    check_order(order);
    const int G = (int)mg->hs.size();
    const int Btot = (int)n_batch;
    std::vector<AdjCall> calls((size_t)G);
    std::vector<int> b0((size_t)G, 0), nb((size_t)G, Btot);
    std::vector<char> active((size_t)G, 1);
    for (int i = 0; i < G; ++i) {
      qgd_handle* h = mg->hs[i];
      const int P = h->P;
      if (shard == QGD_SHARD_COLUMNS) {
        if (G > h->nic) throw QgdError(QGD_EINVAL, "more GPUs than initial-condition columns: use QGD_SHARD_CONTROL_VECTORS");
        const int c0 = (int)((int64_t)i * h->nic / G), c1 = (int)((int64_t)(i + 1) * h->nic / G);
        if (h->col0 != c0 || h->ncol != c1 - c0) { h->col0 = c0; h->ncol = c1 - c0; h->hist_valid = false; }
      } else {
        if (h->col0 != 0 || h->ncol != h->nic) { h->col0 = 0; h->ncol = h->nic; h->hist_valid = false; }
        b0[i] = (int)((int64_t)i * Btot / G); nb[i] = (int)((int64_t)(i + 1) * Btot / G) - b0[i];
        active[i] = nb[i] > 0;
      }
      if (!active[i]) continue;
      CUDA_CHECK(cudaSetDevice(h->device));
      reset_stats(h);
      const int B = nb[i];
      h->d_pcof.reserve((size_t)std::max(P, 1) * B * 8);
      h2d(h, h->d_pcof.p, pcof + (size_t)P * b0[i], (size_t)P * B * 8);
      h->d_target.reserve((size_t)h->N2 * h->nic * 8);
      h2d(h, h->d_target.p, target, (size_t)h->N2 * h->nic * 8);
      h->d_infid.reserve((size_t)B * 8); h->d_guard.reserve((size_t)B * 8);
      h->d_grad.reserve((size_t)std::max(P, 1) * B * 8);
      calls[i] = AdjCall{h->d_pcof.as<double>(), B, order, h->d_target.as<double>(), false, false, false, false,
                         h->d_grad.as<double>(), h->d_infid.as<double>(), h->d_guard.as<double>()};
    }
    // control-vector sharding has no collective: detach the communicators for the duration of the call
    std::vector<void*> comms((size_t)G, nullptr);
    if (shard == QGD_SHARD_CONTROL_VECTORS) for (int i = 0; i < G; ++i) { comms[i] = mg->hs[i]->comm; mg->hs[i]->comm = nullptr; }
    auto restore = [&]() { if (shard == QGD_SHARD_CONTROL_VECTORS) for (int i = 0; i < G; ++i) mg->hs[i]->comm = comms[i]; };
    try {
      const bool coll = shard == QGD_SHARD_COLUMNS && G > 1;
      auto each = [&](void (*f)(qgd_handle*, const AdjCall&), bool grouped) {
        if (grouped) qgd_nccl::group_start();
        for (int i = 0; i < G; ++i) {
          if (!active[i]) continue;
          CUDA_CHECK(cudaSetDevice(mg->hs[i]->device));
          f(mg->hs[i], calls[i]);
        }
        if (grouped) qgd_nccl::group_end();
      };
      each(adj_stage1, false);
      each(adj_exchange1, coll);
      each(adj_stage2, false);
      each(adj_exchange2, coll);
      each(adj_stage3, false);
      for (int i = 0; i < G; ++i) {
        if (!active[i]) continue;
        qgd_handle* h = mg->hs[i];
        CUDA_CHECK(cudaSetDevice(h->device));
        const int B = nb[i];
        const bool writer = shard == QGD_SHARD_CONTROL_VECTORS || i == 0;  // column sharding: every GPU holds the complete result
        if (writer) {
          if (grad) d2h(h, grad + (size_t)h->P * b0[i], h->d_grad.p, (size_t)h->P * B * 8);
          if (infidelity) d2h(h, infidelity + b0[i], h->d_infid.p, (size_t)B * 8);
          if (guard_penalty) d2h(h, guard_penalty + b0[i], h->d_guard.p, (size_t)B * 8);
        }
        remember_hist_pcof(h, pcof + (size_t)h->P * b0[i], B);
      }
      for (int i = 0; i < G; ++i) {
        if (!active[i]) continue;
        qgd_handle* h = mg->hs[i];
        CUDA_CHECK(cudaSetDevice(h->device));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        check_device_error(h);
        finish_timing(h, true, true);
      }
    } catch (...) { restore(); throw; }
    restore();
  });
}

int qgd_infidelity_real(qgd_handle_t* h, const double* final_state, const double* target, int64_t n_batch, double* infidelity) {
  return guarded([&]() {
    require(h && final_state && target && infidelity && n_batch >= 1, "bad arguments");
    // Host arithmetic on 2 * 2N * nic numbers the caller already holds in host memory, as in the reference
    // (infidelity_real is a pair of `dot`s on the final state, src/infidelity.jl:7-18); the gradient path forms the
    // same two inner products on the device in k_terminal.
    const int B = (int)n_batch;
    const int N = h->N, N2 = h->N2;
    for (int b = 0; b < B; ++b) {  // O(2N*nic) work; objective-only calls are host side in the reference too
      double dR = 0, dT = 0;
      for (int c = 0; c < h->nic; ++c)
        for (int r = 0; r < N; ++r) {
          const double pu = final_state[r + (size_t)N2 * (c + (size_t)h->nic * b)], pv = final_state[N + r + (size_t)N2 * (c + (size_t)h->nic * b)];
          const double Ru = target[r + (size_t)N2 * c], Rv = target[N + r + (size_t)N2 * c];
          dR += pu * Ru + pv * Rv;
          dT += pu * Rv - pv * Ru;
        }
      infidelity[b] = 1.0 - (dR * dR + dT * dT) / ((double)h->Ness * (double)h->Ness);
    }
  });
}

int qgd_eval_controls(qgd_handle_t* h, const double* pcof, const double* times, int64_t ntimes, int32_t nderiv, double* p_out,
                      double* q_out, double* gp_out, double* gq_out) {
  return guarded([&]() {
    require(h && times && ntimes >= 1 && nderiv >= 1 && nderiv <= QGD_MAX_M + 1, "bad arguments");
    CUDA_CHECK(cudaSetDevice(h->device));
    const int m = nderiv - 1, nt = (int)ntimes, P = std::max(h->P, 1);
    DevBuf d_times, d_tab, d_pc, d_cv;
    try {
      d_times.reserve((size_t)nt * 8);
      CUDA_CHECK(cudaMemcpyAsync(d_times.p, times, (size_t)nt * 8, cudaMemcpyHostToDevice, h->stream));
      const size_t tsz = (size_t)nt * 2 * nderiv * P;
      d_tab.reserve(tsz * 8);
      k_control_table<<<(nt * h->Nc + 127) / 128, 128, 0, h->stream>>>(h->d_ctrls.as<QgdDevControl>(), h->Nc, h->P, m, nt,
                                                                      d_times.as<double>(), 0.0, 0.0, d_tab.as<double>());
      CUDA_CHECK(cudaGetLastError());
      std::vector<double> tab(tsz);
      if (gp_out || gq_out) {
        CUDA_CHECK(cudaMemcpyAsync(tab.data(), d_tab.p, tsz * 8, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        // table is Taylor scaled; the reference's eval_grad_* are un-scaled: multiply r! back
        for (int n = 0; n < nt; ++n)
          for (int r = 0; r < nderiv; ++r) {
            const double f = factorial_d(r);
            for (int t = 0; t < h->P; ++t) {
              if (gp_out) gp_out[t + (size_t)h->P * (r + (size_t)nderiv * n)] = tab[(((size_t)n * 2 + 0) * nderiv + r) * P + t] * f;
              if (gq_out) gq_out[t + (size_t)h->P * (r + (size_t)nderiv * n)] = tab[(((size_t)n * 2 + 1) * nderiv + r) * P + t] * f;
            }
          }
      }
      if (p_out || q_out) {
        require(pcof != nullptr, "pcof required for control values");
        d_pc.reserve((size_t)P * 8);
        CUDA_CHECK(cudaMemcpyAsync(d_pc.p, pcof, (size_t)h->P * 8, cudaMemcpyHostToDevice, h->stream));
        const size_t total = (size_t)nt * 2 * nderiv * h->Nc;
        d_cv.reserve(std::max<size_t>(total, 1) * 8);
        k_control_values<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(h->d_ctrls.as<QgdDevControl>(), h->Nc, h->P, m, nt,
                                                                                d_tab.as<double>(), d_pc.as<double>(), 1, d_cv.as<double>());
        CUDA_CHECK(cudaGetLastError());
        std::vector<double> cv(total);
        CUDA_CHECK(cudaMemcpyAsync(cv.data(), d_cv.p, total * 8, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        for (int n = 0; n < nt; ++n)
          for (int r = 0; r < nderiv; ++r)
            for (int k = 0; k < h->Nc; ++k) {
              if (p_out) p_out[r + (size_t)nderiv * (k + (size_t)h->Nc * n)] = cv[(((size_t)n * 2 + 0) * nderiv + r) * h->Nc + k];
              if (q_out) q_out[r + (size_t)nderiv * (k + (size_t)h->Nc * n)] = cv[(((size_t)n * 2 + 1) * nderiv + r) * h->Nc + k];
            }
      }
    } catch (...) { d_times.release(); d_tab.release(); d_pc.release(); d_cv.release(); throw; }
    d_times.release(); d_tab.release(); d_pc.release(); d_cv.release();
  });
}

int qgd_compute_derivatives(qgd_handle_t* h, double* uv, int64_t ncols_in, int32_t order, const double* cvals_re,
                            const double* cvals_im, int32_t adjoint) {
  return guarded([&]() {
    require(h && uv && cvals_re && cvals_im && ncols_in >= 1, "bad arguments");
    check_order(order);
    CUDA_CHECK(cudaSetDevice(h->device));
    const int m = order / 2, el = pick_el(h->N);
    build_preconditioner(h, order);
    const size_t sz = (size_t)h->N2 * (m + 1) * ncols_in * 8;
    h->d_scratch.reserve(sz + (size_t)2 * (m + 1) * std::max(h->Nc, 1) * 8);
    double* d_uv = h->d_scratch.as<double>();
    double* d_cv = d_uv + (size_t)h->N2 * (m + 1) * ncols_in;
    CUDA_CHECK(cudaMemcpyAsync(d_uv, uv, sz, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(d_cv, cvals_re, (size_t)(m + 1) * h->Nc * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(d_cv + (size_t)(m + 1) * h->Nc, cvals_im, (size_t)(m + 1) * h->Nc * 8, cudaMemcpyHostToDevice, h->stream));
    QgdDevProb d = make_devprob(h, order);
    // the host passes [1+m, Nc] column-major = device layout [k][r]; device wants [r][k]: transpose on the fly
    std::vector<double> cv((size_t)2 * (m + 1) * std::max(h->Nc, 1));
    for (int r = 0; r <= m; ++r)
      for (int k = 0; k < h->Nc; ++k) {
        cv[((size_t)0 * (m + 1) + r) * h->Nc + k] = cvals_re[r + (size_t)(m + 1) * k];
        cv[((size_t)1 * (m + 1) + r) * h->Nc + k] = cvals_im[r + (size_t)(m + 1) * k];
      }
    CUDA_CHECK(cudaMemcpyAsync(d_cv, cv.data(), cv.size() * 8, cudaMemcpyHostToDevice, h->stream));
    SweepArgs a{};
    reset_stats(h);
    CUDA_CHECK(cudaEventRecord(h->ev[0], h->stream));
    if (dense_derivs_applicable(h, m)) {  // dense operators: FP64 tensor-core contraction (qgd_dense.cu)
      launch_derivs_dense(h, m, d_uv, (int)ncols_in, d_cv, adjoint);
      h->stats.fast_path_launches = -1;  // marks the tensor-core path in the stats of this call
    } else if (!try_derivs_fast(h, d, a, d_uv, (int)ncols_in, d_cv, adjoint)) {
      QGD_DISPATCH_EL(el, launch_derivs, h, d, a, d_uv, (int)ncols_in, d_cv, adjoint);
    }
    CUDA_CHECK(cudaEventRecord(h->ev[1], h->stream));
    CUDA_CHECK(cudaMemcpyAsync(uv, d_uv, sz, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    finish_timing(h, true, false);  // last_forward_ms = device time of the derivative kernel(s)
  });
}

int qgd_get_stats(qgd_handle_t* h, qgd_stats_t* out) {
  return guarded([&]() { require(h && out, "bad arguments"); *out = h->stats; });
}

int qgd_measure_fp64_peak(int device, double* tflops) {
  return guarded([&]() {
    require(tflops != nullptr, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw QgdError(QGD_ECUDA, "no CUDA device available");
    if (device >= 0) CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    const int threads = 512, blocks = prop.multiProcessorCount * 4, iters = 1 << 15;
    double* d_out = nullptr;
    CUDA_CHECK(cudaMalloc(&d_out, (size_t)threads * blocks * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      qgd::k_fp64_peak<<<blocks, threads>>>(d_out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fl = 2.0 * 8.0 * (double)iters * threads * blocks;
      if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
  });
}

// Measured FP64 tensor-core throughput (DMMA.8x8x4, 512 flop per warp instruction) in TFLOP/s: the roofline denominator of
// the dense sweeps (qgd_dense.cu).
int qgd_measure_dmma_peak(int device, double* tflops) {
  return guarded([&]() {
    require(tflops != nullptr, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw QgdError(QGD_ECUDA, "no CUDA device available");
    if (device >= 0) CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    const int threads = 256, blocks = prop.multiProcessorCount * 4, iters = 1 << 13;
    double* d_out = nullptr;
    CUDA_CHECK(cudaMalloc(&d_out, (size_t)threads * blocks * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      qgd::k_dmma_peak<<<blocks, threads>>>(d_out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fl = 512.0 * 8.0 * (double)iters * (threads / 32) * blocks;
      if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
  });
}

}  // extern "C"

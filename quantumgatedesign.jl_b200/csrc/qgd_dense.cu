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
// This is the building block of the dense (C4) sweep: the generic sweep kernels reach 0.4 TFLOP/s on that shape
// (profiles/r01_c4_generic.json); qgd_compute_derivatives routes dense problems with N a multiple of 32 here.
#include "qgd_host.h"

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
  const int ar = lane >> 2, ak = lane & 3;
  const size_t nn = (size_t)N * N;
  const double* Kd = comb + ((size_t)d * 2 + 0) * nn + (size_t)(r0 + ar) * N + ak;
  const double* Sd = comb + ((size_t)d * 2 + 1) * nn + (size_t)(r0 + ar) * N + ak;
  const double* Bu = B + (size_t)ar * S + ak;
  const double* Bv = Bu + N;
#pragma unroll 4
  for (int k0 = 0; k0 < N; k0 += 4) {
    const double bu = Bu[k0], bv = Bv[k0], nbu = -bu;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double aS = __ldg(Sd + (size_t)8 * t * N + k0), aK = __ldg(Kd + (size_t)8 * t * N + k0);
      dmma884_acc(aU[t][0], aU[t][1], aS, bu);
      dmma884_acc(aU[t][0], aU[t][1], aK, bv);
      dmma884_acc(aV[t][0], aV[t][1], aS, bv);
      dmma884_acc(aV[t][0], aV[t][1], aK, nbu);
    }
  }
}

// uv: [2N][1+M][ncols], column 0 of every state column given; columns 1..M are written.
template <int M>
__global__ void __launch_bounds__(256, 1) k_derivs_dense(const double* __restrict__ comb, int N, double* uv, int ncols) {
  extern __shared__ __align__(16) double Wt[];  // [(M+1)][8][S]
  const int N2 = 2 * N, S = N2 + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 8;
  // Taylor column 0 of the 8 state columns (zeros beyond ncols)
  for (int idx = threadIdx.x; idx < 8 * N2; idx += blockDim.x) {
    const int col = idx / N2, row = idx % N2;
    Wt[(size_t)col * S + row] = (c0 + col < ncols) ? uv[(size_t)row + (size_t)N2 * (M + 1) * (c0 + col)] : 0.0;
  }
  __syncthreads();
  const int r0 = 32 * warp;           // level rows of this warp
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
  if (getenv("QGD_DISABLE_DENSE_DMMA")) return false;
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

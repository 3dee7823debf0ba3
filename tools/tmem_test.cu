// tmem_test.cu -- can TMEM (256 KB/SM, otherwise idle in an FP64 kernel) serve as a per-warp scratchpad for Krylov
// basis vectors?  Each of 8 warps of a CTA owns 32 TMEM lanes x 256 columns (32 KB): tcgen05.st / tcgen05.ld with the
// 32x32b shape move 8 x 32-bit per lane (one 128-double basis vector as 4 doubles per lane) per instruction.
// Verifies round-trip integrity and measures the load latency / throughput.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const double (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
               "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3]))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, double (&v)[4]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

__global__ void __launch_bounds__(256, 1) k_tmem(double* out, long long* cyc, int* errors, int iters) {
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tbase_s;
  // this warp's region: lanes 32*(warp%4).., columns 256*(warp/4)..+255  -> 32 slots of 8 columns
  const uint32_t my = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(256 * (warp >> 2));
  int bad = 0;
  for (int s = 0; s < 32; ++s) {
    double v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = 1000.0 * blockIdx.x + 100.0 * warp + s + 0.001 * lane + 0.0001 * e;
    tmem_st8(my + 8 * s, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int s = 31; s >= 0; --s) {
    double v[4];
    tmem_ld8(my + 8 * s, v);
#pragma unroll
    for (int e = 0; e < 4; ++e) bad += (v[e] != 1000.0 * blockIdx.x + 100.0 * warp + s + 0.001 * lane + 0.0001 * e);
  }
  if (bad) atomicAdd(errors, bad);
  // latency: dependent chain slot -> next slot index derived from loaded data
  double acc = 0.0;
  int s = lane & 0;  // uniform
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double v[4];
    tmem_ld8(my + 8 * s, v);
    acc += v[0];
    s = (s + 1 + (__double2loint(v[1]) & 0)) & 31;  // data dependent (always +1)
  }
  long long t1 = clock64();
  __syncthreads();
  // throughput: all 8 warps, independent loads
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) {
    double v[4], w[4];
    tmem_ld8(my + 8 * (i & 31), v);
    tmem_ld8(my + 8 * ((i + 7) & 31), w);
    acc += v[0] + w[3];
  }
  long long t3 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

int main() {
  double* d; long long* c; int* e;
  cudaMalloc(&d, 148 * 256 * 8); cudaMalloc(&c, 16); cudaMalloc(&e, 4); cudaMemset(e, 0, 4);
  const int iters = 4096;
  k_tmem<<<148, 256>>>(d, c, e, iters);
  cudaError_t err = cudaDeviceSynchronize();
  long long h[2]; int he;
  cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost); cudaMemcpy(&he, e, 4, cudaMemcpyDeviceToHost);
  printf("status: %s, round-trip mismatches: %d\n", cudaGetErrorString(err), he);
  printf("tcgen05.ld 32x32b.x8 (+wait::ld) dependent latency: %.1f cycles (8 warps running)\n", (double)h[0] / iters);
  printf("tcgen05.ld 32x32b.x8 throughput, 8 warps x 2 loads/iter: %.1f cycles/iter/warp -> %.1f B/clk/SM\n", (double)h[1] / iters,
         8.0 * 2 * 1024 / ((double)h[1] / iters));
  return (err != cudaSuccess || he != 0);
}

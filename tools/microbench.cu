// microbench.cu -- latency / throughput of the FP64-path primitives the sweep kernels are built from
// (DFMA, 64-bit SHFL butterfly, DMMA.8x8x4 ones-matrix all-reduce, LDS.128), measured with clock64() on
// one warp (latency: dependent chain) and with many warps (throughput).  B200 has no public numbers
// for these; DESIGN.md quotes the output (profiles/r01_microbench.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__device__ __forceinline__ double allreduce_shfl(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double allreduce_dmma(double p) {
  double c0, c1;
  dmma(c0, c1, 1.0, p, 0.0, 0.0);      // column sums of the 4x8 B operand
  const double e = c0 + c1;
  double t0, t1;
  dmma(t0, t1, e, 1.0, 0.0, 0.0);      // row sums of the 8x4 A operand: the total, in every lane
  return t0;
}

template <int MODE>
__global__ void k_lat(double* out, long long* cyc, int iters) {
  double x = 1.0 + threadIdx.x * 1e-3, y = 0.5;
  __shared__ double sm[1024];
  sm[threadIdx.x] = x; sm[threadIdx.x + 32] = y;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { x = fma(x, 0.999999, y); }                    // DFMA dependent chain
    if (MODE == 1) { x = allreduce_shfl(x) * (1.0 / 32.0); }        // 5-stage shuffle butterfly (+DMUL)
    if (MODE == 2) { x = allreduce_dmma(x) * (1.0 / 32.0); }        // 2 DMMA + DADD (+DMUL)
    if (MODE == 3) { double a, b; dmma(a, b, x, 1.0, 0.0, 0.0); x = a * 0.25; }   // one DMMA (+DMUL)
    if (MODE == 4) { int idx = ((int)x) & 31; x = sm[idx] + 1e-9; sm[idx + 64] = x; }  // LDS dependent (+cvt, DADD)
    if (MODE == 5) { x = __shfl_xor_sync(0xffffffffu, x, 1); }     // one 64-bit shuffle
    if (MODE == 6) { x = x + y; }                                  // DADD chain
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// throughput: ILP independent chains per thread, many warps
template <int MODE>
__global__ void k_tput(double* out, int iters) {
  double x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = 1.0 + threadIdx.x * 1e-3 + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) x[j] = fma(x[j], 0.999999, 0.5);
      if (MODE == 1) x[j] = __shfl_xor_sync(0xffffffffu, x[j], 1 + (j & 15));
      if (MODE == 2) { double a, b; dmma(a, b, x[j], 1.0, 0.0, 0.0); x[j] = a; }
      if (MODE == 3) { double a, b; dmma(a, b, x[j], 1.0, 0.0, 0.0); x[j] = fma(a, 0.25, b); }  // DMMA + DFMA mixed
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* d; long long* c;
  cudaMalloc(&d, 148 * 4 * 1024 * 8); cudaMalloc(&c, 8);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const double ghz = p.clockRate * 1e-6;
  printf("device %s, %d SMs, clockRate %.0f MHz\n", p.name, p.multiProcessorCount, p.clockRate * 1e-3);
  const int it = 4096;
  const char* names[] = {"DFMA dependent", "shfl-butterfly allreduce(32) + DMUL", "DMMA allreduce(32) (2 DMMA+DADD) + DMUL",
                         "single DMMA + DMUL", "LDS dependent (+F2I, DADD, STS)", "64-bit SHFL", "DADD dependent"};
#define LAT(M) { k_lat<M><<<1, 32>>>(d, c, it); k_lat<M><<<1, 32>>>(d, c, it); long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
                 printf("latency  %-44s %7.1f cycles/iter\n", names[M], (double)h / it); }
  LAT(0) LAT(1) LAT(2) LAT(3) LAT(4) LAT(5) LAT(6)
  const char* tn[] = {"DFMA", "64-bit SHFL", "DMMA.8x8x4", "DMMA+DFMA pairs"};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define TP(M) { const int blocks = p.multiProcessorCount * 2, thr = 512, iters = 8192; \
    k_tput<M><<<blocks, thr>>>(d, iters); cudaEventRecord(e0); k_tput<M><<<blocks, thr>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); \
    float ms; cudaEventElapsedTime(&ms, e0, e1); double ops = 8.0 * iters * (double)blocks * thr / 32; \
    printf("throughput %-18s %8.3f warp-instr/ns chip  = %6.3f warp-instr/clk/SM (at %.3f GHz nominal)\n", tn[M], ops / (ms * 1e6), ops / (ms * 1e6) / p.multiProcessorCount / ghz, ghz); }
  TP(0) TP(1) TP(2) TP(3)
  // DFMA issue rate as a function of warps per SM sub-partition (8 independent chains per thread)
  for (int wps = 1; wps <= 8; wps *= 2) {
    const int blocks = p.multiProcessorCount, thr = 128 * wps, iters = 16384;
    k_tput<0><<<blocks, thr>>>(d, iters); cudaEventRecord(e0); k_tput<0><<<blocks, thr>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); double ops = 8.0 * iters * (double)blocks * thr / 32;
    printf("DFMA, %d warp(s) per sub-partition: %6.3f warp-instr/clk/SM\n", wps, ops / (ms * 1e6) / p.multiProcessorCount / ghz);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}

// Microbenchmark: sustained FP64 issue rate on this GPU (DFMA / DADD / DMUL, many independent chains).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
                   x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b); }
    if (OP == 1) { x0 = __dadd_rn(x0, a); x1 = __dadd_rn(x1, a); x2 = __dadd_rn(x2, a); x3 = __dadd_rn(x3, a);
                   x4 = __dadd_rn(x4, a); x5 = __dadd_rn(x5, a); x6 = __dadd_rn(x6, a); x7 = __dadd_rn(x7, a); }
    if (OP == 2) { x0 = __dmul_rn(x0, a); x1 = __dmul_rn(x1, a); x2 = __dmul_rn(x2, a); x3 = __dmul_rn(x3, a);
                   x4 = __dmul_rn(x4, a); x5 = __dmul_rn(x5, a); x6 = __dmul_rn(x6, a); x7 = __dmul_rn(x7, a); }
    if (OP == 3) { float f0 = (float)x0; f0 = __fmaf_rn(f0, (float)a, (float)b); x0 = f0; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <int OP>
void run(const char* name, double* d) {
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * 8;
  printf("%s: %.3f ms, %.2f Tops/s (x2 flops for FMA), %.2f warp-instr/clk/SM at 1.93 GHz\n", name, ms, ops / ms / 1e9,
         ops / 32 / (ms * 1e-3) / 148 / 1.93e9);
}
int main() {
  double* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(double));
  run<0>("DFMA", d); run<1>("DADD", d); run<2>("DMUL", d);
  return 0;
}

// MUFU throughput probe: independent chains of tanh.approx / ex2.approx / rcp.approx per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters) {
  float a[8];
  for (int j = 0; j < 8; ++j) a[j] = 0.001f * (threadIdx.x + j) + 0.1f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[j]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
      if (OP == 3) a[j] = fmaf(a[j], 1.0001f, 0.5f);
    }
  }
  float s = 0;
  for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name) {
  float* d;
  cudaMalloc(&d, 148 * 4 * 1024 * 4);
  const int iters = 4096;
  k<OP><<<148 * 2, 1024>>>(d, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<148 * 2, 1024>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * 2 * 1024 * 8.0 * iters;
  printf("%-8s %8.3f ms  %7.1f Gop/s  = %5.2f ops/clk/SM at 1.9 GHz\n", name, ms, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.9);
  cudaFree(d);
}
int main() {
  run<0>("tanh"); run<1>("ex2"); run<2>("rcp"); run<3>("ffma");
  return 0;
}

// Micro-benchmark (B200): issue rate of the special-function ops a fused SiLU could use, per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rates mufu_rates.cu && ./mufu_rates
// Each thread runs a long dependent-free stream of one op on 8 independent registers; 148 x 4 blocks of 256 threads.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + 0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
      if (OP == 4) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, float* out) {
  const int iters = 4096, blocks = 148 * 4, threads = 256;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<blocks, threads>>>(out, 64, 0.5f);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads * iters * 8 * (OP == 4 ? 2 : 1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-22s %8.3f ms  %7.1f Gop/s  %6.2f lanes/clk/SM at the %d MHz nominal clock\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  run<0>("tanh.approx.f32", out); run<1>("ex2.approx.f32", out); run<2>("rcp.approx.f32", out); run<3>("fma.rn.f32", out);
  run<4>("tanh.approx.f16x2", out);
  return 0;
}

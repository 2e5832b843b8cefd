// Micro-benchmark: issue rate of ex2.approx (f32 / f16x2), cvt.f16x2 and FFMA per SM sub-partition on sm_100a.
#include <cstdio>
#include <cuda_fp16.h>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = 0.001f * (threadIdx.x + i);
  uint32_t h[16];
  for (int i = 0; i < 16; ++i) h[i] = 0x3c003c00u + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
      if (MODE == 3) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[i]), "f"(x[(i + 1) & 15])); }
      if (MODE == 4) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[(i + 8) & 15])); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const char* names[] = {"ex2.f32", "ex2.f16x2", "ffma", "cvt.f16x2", "ex2.bf16x2", "ex2.f32+ffma"};
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int mode = 0; mode < 6; ++mode) {
      const int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        switch (mode) {
          case 0: k<0><<<1, warps * 32>>>(out, cyc, iters); break;
          case 1: k<1><<<1, warps * 32>>>(out, cyc, iters); break;
          case 2: k<2><<<1, warps * 32>>>(out, cyc, iters); break;
          case 3: k<3><<<1, warps * 32>>>(out, cyc, iters); break;
          case 4: k<4><<<1, warps * 32>>>(out, cyc, iters); break;
          case 5: k<5><<<1, warps * 32>>>(out, cyc, iters); break;
        }
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double per = (double)c / (iters * 16.0 * (warps / 4));   // cycles per warp-instruction per SMSP
      printf("warps %2d %-14s %8lld cycles  -> %.2f cycles per warp-instr per SMSP\n", warps, names[mode], c, per);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

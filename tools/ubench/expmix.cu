// Micro-benchmark: the softmax inner loop's instruction mix per pair of scores on sm_100a:
//   2 x FFMA (x = s*c - m), F2FP.F16.F32.PACK_AB, ex2.approx.f16x2 (= 2 MUFU.EX2.F16), PRMT -> packed f16x2
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float c, float m) {
  float s[32];
  for (int i = 0; i < 32; ++i) s[i] = 0.001f * (threadIdx.x + i);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float x0 = s[2 * j], x1 = s[2 * j + 1];
      if (MODE >= 1) { x0 = fmaf(x0, c, -m); x1 = fmaf(x1, c, -m); }
      uint32_t in, o;
      if (MODE == 4) {          // fp32 ex2 then pack
        float e0, e1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
        __half2 h = __floats2half2_rn(e0, e1);
        o = *reinterpret_cast<uint32_t*>(&h);
      } else {
        __half2 h = __floats2half2_rn(x0, x1);
        in = *reinterpret_cast<uint32_t*>(&h);
        if (MODE == 3) o = in;  // no MUFU at all: FFMA + F2FP only
        else asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(o) : "r"(in));
      }
      pk[j] = o;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc ^= pk[j]; s[2 * j] += __uint_as_float(pk[j] & 1); }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc) + s[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const char* names[] = {"cvt + ex2.f16x2", "ffma + cvt + ex2.f16x2", "(unused)", "ffma + cvt only", "ffma + 2 ex2.f32 + cvt"};
  for (int warps = 4; warps <= 16; warps *= 2)
    for (int mode = 0; mode < 5; ++mode) {
      if (mode == 2) continue;
      const int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters, 0.2f, 1.0f);
        if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, iters, 0.2f, 1.0f);
        if (mode == 3) k<3><<<1, warps * 32>>>(out, cyc, iters, 0.2f, 1.0f);
        if (mode == 4) k<4><<<1, warps * 32>>>(out, cyc, iters, 0.2f, 1.0f);
      }
      long long cc; cudaMemcpy(&cc, cyc, 8, cudaMemcpyDeviceToHost);
      printf("warps/SMSP %d %-26s %8lld cycles -> %.1f cycles per 32-score chunk per warp (%.2f per score per SMSP)\n", warps / 4,
             names[mode], cc, (double)cc / iters, (double)cc / iters / 32 / 1.0 * 1.0 / (warps / 4));
    }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

// Micro-benchmark: instruction ORDER of the softmax inner loop on sm_100a.  Per 32-score chunk and thread:
//   mode 0: 32 MUFU.EX2 only (independent inputs, results xor-consumed)      -> raw MUFU issue rate
//   mode 1: compiler order  (FFMA.., then MUFU, MUFU, F2FP triples)
//   mode 2: forced order    32 FFMA, 32 MUFU back to back, then 16 F2FP
//   mode 3: forced order, two chunks software-pipelined (MUFU of chunk b between the F2FP of chunk a)
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
__device__ __forceinline__ float ex2v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2n(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packv(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float c, float m) {
  float s[64];
  for (int i = 0; i < 64; ++i) s[i] = -0.01f * ((threadIdx.x + i) & 63);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) e[j] = ex2v(s[j]);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= __float_as_uint(e[j]);
    } else if (MODE == 1) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __half2 h = __floats2half2_rn(ex2n(fmaf(s[2 * j], c, -m)), ex2n(fmaf(s[2 * j + 1], c, -m)));
        pk[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= pk[j];
    } else if (MODE == 2) {
      float e[32];
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; ++j) e[j] = fmaf(s[j], c, -m);
#pragma unroll
      for (int j = 0; j < 32; ++j) e[j] = ex2v(e[j]);
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = packv(e[2 * j], e[2 * j + 1]);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= pk[j];
    } else {
      float e[64];
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 64; ++j) e[j] = fmaf(s[j], c, -m);
#pragma unroll
      for (int j = 0; j < 32; ++j) e[j] = ex2v(e[j]);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        e[32 + 2 * j] = ex2v(e[32 + 2 * j]);
        e[33 + 2 * j] = ex2v(e[33 + 2 * j]);
        pk[j] = packv(e[2 * j], e[2 * j + 1]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[16 + j] = packv(e[32 + 2 * j], e[33 + 2 * j]);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= pk[j];
    }
    s[it & 63] += __uint_as_float(acc & 1);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc) + s[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const char* names[] = {"32 MUFU only", "compiler order", "forced: FFMA, MUFU, F2FP", "forced, 2 chunks pipelined"};
  const int scores[] = {32, 32, 32, 64};
  for (int warps = 4; warps <= 16; warps *= 2)
    for (int mode = 0; mode < 4; ++mode) {
      const int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters, 0.2f, 1.0f);
        if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, iters, 0.2f, 1.0f);
        if (mode == 2) k<2><<<1, warps * 32>>>(out, cyc, iters, 0.2f, 1.0f);
        if (mode == 3) k<3><<<1, warps * 32>>>(out, cyc, iters, 0.2f, 1.0f);
      }
      long long cc; cudaMemcpy(&cc, cyc, 8, cudaMemcpyDeviceToHost);
      printf("warps/SMSP %d %-28s %8lld cycles -> %.1f cycles per 32 scores per warp (%.2f per score per SMSP)\n", warps / 4,
             names[mode], cc, (double)cc / iters * 32 / scores[mode], (double)cc / iters / scores[mode] / (warps / 4));
    }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

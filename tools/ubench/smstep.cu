// Micro-benchmark: the softmax step of tc_layer.cuh in isolation (no MMA, no TMA): per iteration and warp
//   mbarrier try_wait (already complete) -> tcgen05.ld 64 columns + wait -> votes -> row maximum (32 FMNMX3) -> lazy-max vote
//   -> 64 x (FFMA, MUFU.EX2, 1/2 F2FP) -> 2 x tcgen05.st + wait::st + fence -> __syncwarp + elected mbarrier arrive.
// KO bits remove one component each, WARPS = compute warps per CTA (4 = one per sub-partition, 8 = two, 16 = four).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include "../../edge_diffusion_tts_b200/csrc/umma.cuh"
using namespace edtts::tc;
enum { KO_WAIT = 1, KO_LD = 2, KO_VOTE = 4, KO_MAX = 8, KO_GROW = 16, KO_EXP = 32, KO_ST = 64, KO_ARRIVE = 128 };
template <int KO>
__global__ void k(long long* cyc, float* out, int iters, float c) {
  __shared__ uint64_t bar_done, bar_arr;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_arr, (1 << 20) - 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) mbar_arrive(&bar_done);                    // phase 0 of bar_done complete: waits on parity 0 succeed at once
  __syncthreads();
  const uint32_t tS = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64 % 512;
  float m_run = -INFINITY, acc = 0.f;
  uint32_t cur[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) cur[j] = __float_as_uint(0.01f * ((lane + j) & 31));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (!(KO & KO_WAIT)) mbar_wait(&bar_done, 0);
    tc_fence_after();
    if (!(KO & KO_LD)) {
      tmem_ld32_nw(tS, cur);
      tmem_ld32_nw(tS + 32, cur + 32);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_tie16(cur + 16 * q);
#pragma unroll
      for (int j = 0; j < 64; ++j) cur[j] = (cur[j] & 0x3fffffffu) | 0x3c000000u;   // keep the garbage finite and small
    }
    bool full0 = true, full1 = true;
    if (!(KO & KO_VOTE)) {
      const int lo = it & 1 ? 0 : -1, hi = 63 + (lane >> 5);
      full0 = __any_sync(0xffffffffu, lo <= 31) && __all_sync(0xffffffffu, lo <= 0 && hi >= 31);
      full1 = __any_sync(0xffffffffu, hi >= 32) && __all_sync(0xffffffffu, lo <= 32 && hi >= 63);
    }
    float bmax = m_run;
    if (!(KO & KO_MAX)) {
      float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        b0 = fmaxf(fmaxf(b0, __uint_as_float(cur[j])), __uint_as_float(cur[j + 1]));
        b1 = fmaxf(fmaxf(b1, __uint_as_float(cur[j + 2])), __uint_as_float(cur[j + 3]));
        b2 = fmaxf(fmaxf(b2, __uint_as_float(cur[j + 4])), __uint_as_float(cur[j + 5]));
        b3 = fmaxf(fmaxf(b3, __uint_as_float(cur[j + 6])), __uint_as_float(cur[j + 7]));
      }
      bmax = fmaxf(fmaxf(b0, b1), fmaxf(b2, b3)) * c;
    }
    if (!(KO & KO_GROW)) {
      const bool grow = bmax > m_run + 6.0f;
      if (__any_sync(0xffffffffu, grow)) m_run = grow ? bmax : m_run;
    }
    const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
    uint32_t pk[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x0 = fmaf(__uint_as_float(cur[2 * j]), c, -m_use), x1 = fmaf(__uint_as_float(cur[2 * j + 1]), c, -m_use);
      __half2 h;
      if (KO & KO_EXP) h = __floats2half2_rn(x0 * 0.001f, x1 * 0.001f);
      else h = __floats2half2_rn(ex2_approx(x0), ex2_approx(x1));
      pk[j] = *reinterpret_cast<uint32_t*>(&h);
      if (!full0 || !full1) pk[j] &= 0xffff;
    }
    if (!(KO & KO_ST)) {
      tmem_st16u(tS, pk);
      tmem_st16u(tS + 16, pk + 16);
      tmem_st_wait();
      tc_fence_before();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += __uint_as_float(pk[j]);
    }
    if (!(KO & KO_ARRIVE)) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_arr);
    }
  }
  const long long t1 = clock64();
  out[tid] = acc + m_run + __uint_as_float(cur[5]);
  if (tid == 0) cyc[0] = t1 - t0;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}
template <int KO>
void run(const char* name, long long* cyc, float* out) {
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int iters = 1000;
    for (int rep = 0; rep < 2; ++rep) k<KO><<<1, warps * 32>>>(cyc, out, iters, 0.2f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SMSP %d: %7.1f cycles per step\n", name, warps / 4, (double)c / iters);
  }
}
int main() {
  long long* cyc; float* out; cudaMalloc(&cyc, 64); cudaMalloc(&out, 4096);
  run<0>("full step", cyc, out);
  run<KO_EXP>("no MUFU", cyc, out);
  run<KO_MAX>("no maximum", cyc, out);
  run<KO_VOTE | KO_GROW>("no votes", cyc, out);
  run<KO_LD>("no tcgen05.ld", cyc, out);
  run<KO_ST>("no tcgen05.st / wait::st", cyc, out);
  run<KO_WAIT>("no mbarrier wait", cyc, out);
  run<KO_ARRIVE>("no syncwarp + arrive", cyc, out);
  run<KO_WAIT | KO_LD | KO_VOTE | KO_MAX | KO_GROW | KO_ST | KO_ARRIVE>("exp section only", cyc, out);
  run<KO_EXP | KO_MAX>("skeleton (no MUFU, no maximum)", cyc, out);
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

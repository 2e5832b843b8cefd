// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, cta_group::1) as a function of N, for the A operand in
// shared memory (SS) or tensor memory (TS), issued back to back by one thread; one commit + mbarrier wait at the end.
#include <cstdio>
#include <cstdint>
#include "../../edge_diffusion_tts_b200/csrc/umma.cuh"
using namespace edtts::tc;
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// CONV: the whole warp runs the loop converged and one elected lane issues (operands stay in uniform registers);
// otherwise a single thread inside a divergent branch issues (operands pass through R2UR for every instruction).
template <bool TS, bool CONV>
__global__ void k(long long* cyc, int N, int nmma, int ndist) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid * 16; i < 96 * 1024; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (CONV ? (__shfl_sync(0xffffffffu, warp, 0) == 0) : (tid == 0)) {
    const uint32_t idesc = make_idesc(128, (uint32_t)N);
    const uint64_t da = make_desc(smem_u32(smem), 2048, 128);
    const uint64_t db = make_desc(smem_u32(smem) + 48 * 1024, (uint32_t)N * 16, 128);
    long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const uint32_t d = tm + (uint32_t)((i & (ndist - 1)) * 256);   // ndist = 1: same accumulator, 2: alternate
      if (!CONV || elect_one()) {
        if (TS) umma_f16_ts(d, tm + 480, db, idesc, i >= ndist);
        else umma_bf16(d, da, db, idesc, i >= ndist);
      }
    }
    long long t1 = clock64();
    if (!CONV || elect_one()) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
int main() {
  long long* cyc; cudaMalloc(&cyc, 64);
  cudaFuncSetAttribute(k<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int Ns[] = {16, 32, 48, 64, 96, 128, 160, 256};
  for (int conv = 0; conv < 2; ++conv)
  for (int ts = 0; ts < 2; ++ts)
    for (int nd = 1; nd <= 2; ++nd)
      for (int N : Ns) {
        if (nd == 2 && N > 192) continue;
        const int nmma = 512;
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
          if (conv) { if (ts) k<true, true><<<1, 128, 100 * 1024>>>(cyc, N, nmma, nd); else k<false, true><<<1, 128, 100 * 1024>>>(cyc, N, nmma, nd); }
          else { if (ts) k<true, false><<<1, 128, 100 * 1024>>>(cyc, N, nmma, nd); else k<false, false><<<1, 128, 100 * 1024>>>(cyc, N, nmma, nd); }
        }
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        if (conv == 0 && (N != 64 && N != 256)) continue;
        printf("%s %s accumulators %d N %3d: issue %.1f cycles/MMA, complete %.1f cycles/MMA (floor 128*N/256 = %d)\n", conv ? "converged" : "divergent", ts ? "TS" : "SS", nd, N,
               (double)h[0] / nmma, (double)h[1] / nmma, N / 2);
      }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

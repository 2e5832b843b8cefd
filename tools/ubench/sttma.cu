// Micro-benchmark: per-SM store throughput, LSU (st.global.v4 per thread) against the TMA engine (cp.async.bulk shared -> global).
// Per "tile" a CTA stores `chunks` pieces of 2 KB (128 rows x 16 B), chunk-major layout as in the fused kernel's tail.
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(uint4* out, long long R, int ntiles, int mode, int chunks, int piece) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  for (int i = tid * 16; i < chunks * 2048; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(tid, i, 0, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long row0 = (long long)t * 128;
    if (mode == 0) {
      const int row = tid & 127, half = tid >> 7;
      for (int g = 0; g < chunks / 2; ++g) {
        const int c = half * (chunks / 2) + g;
        out[(long long)c * R + row0 + row] = *reinterpret_cast<uint4*>(smem + c * 2048 + row * 16);
      }
    } else {
      // `piece` bytes per bulk copy; issued by `mode` threads (1 = one elected thread, 32 = a warp's lanes share the chunks)
      const int per = 2048 / piece;
      if (tid < mode) {
        for (int j = tid; j < chunks * per; j += mode) {
          const int c = j / per, o = (j % per) * piece;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint8_t*>(out + (long long)c * R + row0) + o),
                       "r"(smem_u32(smem + c * 2048 + o)), "r"(piece) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncthreads();
    }
  }
  if (mode != 0 && tid < mode) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
  const int ntiles = 1792 * 4;
  const long long R = (long long)ntiles * 128;
  uint4* out;
  cudaMalloc(&out, (size_t)R * 40 * 16 + (1 << 20));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 2048);
  for (int grid : {148, 16})
    for (int chunks : {20, 40})
      for (int cfg = 0; cfg < 4; ++cfg) {
        const int mode = cfg == 0 ? 0 : cfg == 1 ? 1 : 32, piece = cfg == 3 ? 1024 : 2048;
        float best = 1e9;
        const int nt = grid == 148 ? ntiles : ntiles / 8;
        for (int rep = 0; rep < 5; ++rep) {
          cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
          cudaEventRecord(a);
          k<<<grid, 256, chunks * 2048>>>(out, R, nt, mode, chunks, piece);
          cudaEventRecord(b); cudaEventSynchronize(b);
          float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        const double bytes = (double)nt * 128 * chunks * 16;
        printf("grid %3d chunks %d %-28s: %.3f ms, %.1f GB/s, %.1f B/clk/SM, %.0f cycles per tile\n", grid, chunks,
               cfg == 0 ? "st.global.v4" : cfg == 1 ? "bulk 2 KB, 1 thread" : cfg == 2 ? "bulk 2 KB, 32 threads" : "bulk 1 KB, 32 threads",
               best, bytes / best / 1e6, bytes / (best * 1e-3) / grid / 1.965e9, best * 1e-3 * 1.965e9 * grid / nt);
      }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

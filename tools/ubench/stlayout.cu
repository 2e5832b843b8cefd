// Micro-benchmark: store throughput of the fused kernel's tail pattern.  Per "tile" a CTA of 256 threads stores 128 rows x 20 chunks
// of 16 bytes (thread = (row, half): 10 chunks each).  Layout A (chunk-major, as shipped): chunk c of row r at ((c * R + r) * 16);
// layout B (tile-major): tile t at t * 40960 + (c * 128 + r) * 16.  One CTA per SM, tiles dealt round-robin.
#include <cstdio>
#include <cstdint>
__global__ void k(uint4* out, long long R, int ntiles, int layout, int chunks, long long* cyc) {
  const int tid = threadIdx.x, row = tid & 127, half = tid >> 7;
  const long long t0 = clock64();
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long row0 = (long long)t * 128;
    for (int g = 0; g < chunks / 2; ++g) {
      const int c = half * (chunks / 2) + g;
      const long long idx = layout == 0 ? (long long)c * R + row0 + row : (long long)t * (128 * chunks) + c * 128 + row;
      out[idx] = make_uint4(tid, t, g, 1);
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  const int ntiles = 1792 * 4;
  const long long R = (long long)ntiles * 128;
  uint4* out; long long* cyc;
  cudaMalloc(&out, (size_t)R * 40 * 16 + (1 << 20)); cudaMalloc(&cyc, 148 * 8);
  for (int grid : {148, 16})
  for (int chunks : {20, 40})
    for (int layout = 0; layout < 2; ++layout) {
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        k<<<grid, 256>>>(out, R, grid == 148 ? ntiles : ntiles / 8, layout, chunks, cyc);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      const int nt = grid == 148 ? ntiles : ntiles / 8;
      const double bytes = (double)nt * 128 * chunks * 16;
      printf("grid %3d chunks %d layout %s: %.3f ms, %.1f GB/s, %.1f B/clk/SM (1.965 GHz), %.0f cycles per tile\n", grid, chunks, layout ? "tile-major " : "chunk-major",
             best, bytes / best / 1e6, bytes / (best * 1e-3) / grid / 1.965e9, best * 1e-3 * 1.965e9 * grid / nt);
    }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

// Tensor-core routes of SemanticEncoder.proj (models/encoder.py:41-46) and of the VectorQuantizer nearest-codeword search
// (models/vq.py:75-83, 148-159): fp32-grade tf32 x 3 GEMMs (tf32x3.cuh) with the operator's epilogue fused.
//
//   t3_linear_kernel   out[rows, N] = epi(A[rows, K] W^T + b), one CTA per 128 rows, N <= 256.  The A rows stream in 32
//                      columns at a time with 16-byte asynchronous copies (double-buffered), are split hi / lo into the
//                      operand image, the weight chunk arrives by one bulk copy (TMA engine), accumulator in tensor memory.
//                      Epilogues: bias; bias + exact GELU; bias + GELU + LayerNorm over the N outputs (Linear -> GELU ->
//                      LayerNorm of encoder.py:42-44 in one kernel; the second Linear is another launch of the same kernel).
//   t3_vq_kernel       the distance product z E^T for 128 rows x all K <= 512 codewords (two 256-column accumulators = all
//                      512 tensor-memory columns), then per row d = (||z||^2 - 2 z.e) + ||e||^2 in the reference's operation
//                      order, best and second-best kept while the accumulator is read out, and -- exactly as in the CUDA-core
//                      kernel (vq.cu) -- an fp64 re-rank of the two candidates when they are closer than the rounding bound:
//                      the result is the exact argmin with torch.argmin's first-minimum tie-break.
#include "tf32x3.cuh"
#include <stdlib.h>

namespace edtts {
namespace t3 {

constexpr int THREADS = 256;
constexpr int XS_LD = KC + 4;                         // padded row of the fp32 staging tile (16-byte aligned, conflict-free float4 reads)

__global__ void pack_w_tf32_kernel(const float* __restrict__ W, float* __restrict__ img, int k, int n, int ldw, int nchunk) {
  const int total = nchunk * 8 * n * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 3, row = (i >> 2) % n, s = ((i >> 2) / n) & 7, c = (i >> 2) / (n * 8);
    const int kk = KC * c + 4 * s + j;
    const float w = kk < k ? W[(int64_t)row * ldw + kk] : 0.f;
    const float hi = tf32_rna(w);
    const int64_t base = (int64_t)c * (2 * 8 * n * 4);
    img[base + ((int64_t)s * n + row) * 4 + j] = hi;
    img[base + 8 * n * 4 + ((int64_t)s * n + row) * 4 + j] = w - hi;
  }
}
int pack_w_tf32(const float* W, float* img, int k, int n, int ldw, cudaStream_t st) {
  const int nchunk = (k + KC - 1) / KC;
  const int total = nchunk * 8 * n * 4;
  LaunchScope ls(KC_TC_MISC, st);
  pack_w_tf32_kernel<<<(total + 255) / 256, 256, 0, st>>>(W, img, k, n, ldw, nchunk);
  return check_launch("pack_w_tf32");
}

// ---- shared pieces of the two kernels ------------------------------------------------------------------------------
// stage columns [32 c, 32 c + 32) of the tile's 128 rows into xs (fp32, row stride XS_LD): 16-byte asynchronous copies;
// rows beyond `rows` and columns beyond K are zero-filled with plain stores
__device__ __forceinline__ void stage_rows(const float* __restrict__ A, int64_t row0, int64_t rows, int K, int lda, int c, float* xs) {
  for (int i = threadIdx.x; i < TM * (KC / 4); i += THREADS) {
    const int r = i >> 3, p = i & 7;
    const int col = KC * c + 4 * p;
    float* dst = xs + r * XS_LD + 4 * p;
    if (row0 + r < rows && col + 3 < K) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(A + (row0 + r) * lda + col) : "memory");
    } else {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < rows) {
        const float* src = A + (row0 + r) * lda;
        v.x = col < K ? src[col] : 0.f;
        v.y = col + 1 < K ? src[col + 1] : 0.f;
        v.z = col + 2 < K ? src[col + 2] : 0.f;
        v.w = col + 3 < K ? src[col + 3] : 0.f;
      }
      *reinterpret_cast<float4*>(dst) = v;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
// thread (r = tid & 127, h = tid >> 7): 16 of the chunk's 32 values of row r -> operand image; returns their sum of squares
__device__ __forceinline__ float split_rows(const float* xs, uint8_t* sAh, uint8_t* sAl) {
  const int r = threadIdx.x & (TM - 1), h = threadIdx.x >> 7;
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v4 = *reinterpret_cast<const float4*>(xs + r * XS_LD + 16 * h + 4 * q);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    ss = fmaf(v[0], v[0], ss); ss = fmaf(v[1], v[1], ss); ss = fmaf(v[2], v[2], ss); ss = fmaf(v[3], v[3], ss);
    split_store(sAh, sAl, 4 * h + q, r, v);
  }
  return ss;
}

// ---- linear -----------------------------------------------------------------------------------------------------------
struct LinArgs {
  const float* A;        // [rows][lda]
  const float* wimg;     // weight image (pack_w_tf32), N rows
  const float* bias;     // [N]
  const float* ln_w;     // EPI_T3_GELU_LN: LayerNorm weight / bias [N]
  const float* ln_b;
  float* out;            // [rows][ldo]
  int64_t rows;
  int K, lda, N, ldo, epi, nchunk;
  float ln_eps;
};

__global__ void __launch_bounds__(THREADS, 2) t3_linear_kernel(const LinArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sAh = smem;
  uint8_t* sAl = smem + A_HALF;
  uint8_t* sW = smem + 2 * A_HALF;
  const int w_half = 8 * a.N * 16;
  float* xs0 = reinterpret_cast<float*>(sW + 2 * w_half);
  float* sred = xs0 + 2 * TM * XS_LD;                    // [2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sred + 2 * TM);
  uint64_t* bar_w = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * TM;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  uint32_t ph_w = 0, ph_m = 0;

  stage_rows(a.A, row0, a.rows, a.K, a.lda, 0, xs0);
  for (int c = 0; c < a.nchunk; ++c) {
    const float* xs = xs0 + (c & 1) * TM * XS_LD;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (c + 1 < a.nchunk) stage_rows(a.A, row0, a.rows, a.K, a.lda, c + 1, xs0 + ((c + 1) & 1) * TM * XS_LD);
    if (c > 0) {
      mbar_wait(bar_mma, ph_m);
      ph_m ^= 1;
      tc_fence_after();
    }
    if (tid == 0) {
      mbar_expect_tx(bar_w, 2 * w_half);
      bulk_g2s(sW, a.wimg + (int64_t)c * (2 * w_half / 4), 2 * w_half, bar_w);
    }
    split_rows(xs, sAh, sAl);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_w, ph_w);
      tc_fence_after();
      const int nks = min(KC, a.K - KC * c + 7) / 8;
      issue_chunk(tmem, smem_u32(sAh), smem_u32(sAl), smem_u32(sW), smem_u32(sW) + w_half, nks, a.N, c > 0);
      umma_commit(bar_mma);
    }
    ph_w ^= 1;
  }
  mbar_wait(bar_mma, ph_m);
  tc_fence_after();

  // ---- epilogue: thread = (row r, half of the N columns); N / 2 <= 128 values in registers ----------------------------
  {
    const int lq = warp & 3, half = warp >> 2;
    const int r = lq * 32 + lane;
    const int ncol = a.N / 2;                            // a multiple of 8
    const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16) + half * ncol;
    const bool valid = row0 + r < a.rows;
    float* orow = a.out + (row0 + r) * a.ldo + half * ncol;
    if (a.epi != EPI_T3_GELU_LN) {
      for (int c8 = 0; c8 < ncol; c8 += 8) {
        float v[8];
        tmem_ld8(trow + c8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] += a.bias[half * ncol + c8 + j];
          if (a.epi == EPI_T3_GELU) v[j] = gelu_erf(v[j]);
        }
        if (valid) {
          *reinterpret_cast<float4*>(orow + c8) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(orow + c8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    } else {
      // Linear -> GELU -> LayerNorm (two-pass statistics over the row's N values, the halves meet in shared memory)
      float v[64];                                       // N <= 128
      float s1 = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < 64; c8 += 8) {
        if (c8 < ncol) {
          tmem_ld8(trow + c8, v + c8);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[c8 + j] = gelu_erf(v[c8 + j] + a.bias[half * ncol + c8 + j]);
            s1 += v[c8 + j];
          }
        }
      }
      sred[half * TM + r] = s1;
      __syncthreads();
      const float mean = (sred[r] + sred[TM + r]) / (float)a.N;
      float s2 = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < 64; ++c8)
        if (c8 < ncol) s2 = fmaf(v[c8] - mean, v[c8] - mean, s2);
      __syncthreads();
      sred[half * TM + r] = s2;
      __syncthreads();
      const float rstd = rsqrtf((sred[r] + sred[TM + r]) / (float)a.N + a.ln_eps);
#pragma unroll
      for (int c4 = 0; c4 < 64; c4 += 4) {
        if (c4 < ncol && valid) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = (v[c4 + j] - mean) * rstd * a.ln_w[half * ncol + c4 + j] + a.ln_b[half * ncol + c4 + j];
          *reinterpret_cast<float4*>(orow + c4) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

static int linear_smem(int N) { return 2 * A_HALF + 2 * 8 * N * 16 + 2 * TM * XS_LD * 4 + 2 * TM * 4 + 64; }

int launch_t3_linear(const float* A, int64_t rows, int K, int lda, const float* wimg, const float* bias, int N, float* out, int ldo,
                     int epi, const float* ln_w, const float* ln_b, float ln_eps, cudaStream_t st) {
  EDTTS_REQUIRE(N % 16 == 0 && N <= 256 && K % 4 == 0 && lda % 4 == 0 && ldo % 4 == 0, EDTTS_ENOTSUP, "t3_linear: K=%d N=%d", K, N);
  EDTTS_REQUIRE(epi != EPI_T3_GELU_LN || N <= 128, EDTTS_ENOTSUP, "t3_linear: LayerNorm epilogue needs N <= 128");
  LinArgs a;
  a.A = A; a.rows = rows; a.K = K; a.lda = lda; a.wimg = wimg; a.bias = bias; a.N = N; a.out = out; a.ldo = ldo; a.epi = epi;
  a.ln_w = ln_w; a.ln_b = ln_b; a.ln_eps = ln_eps;
  a.nchunk = (K + KC - 1) / KC;
  static PerDeviceOnce configured;
  if (configured.need()) {
    if (cudaFuncSetAttribute(t3_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, linear_smem(256)) != cudaSuccess)
      return check_launch("t3_linear smem attribute");
    configured.set();
  }
  LaunchScope ls(KC_TC_GEMM, st);
  t3_linear_kernel<<<(unsigned)((rows + TM - 1) / TM), THREADS, linear_smem(N), st>>>(a);
  return check_launch("t3_linear");
}

// ---- VQ argmin ----------------------------------------------------------------------------------------------------------
struct Cand {
  float v;
  int i;
};
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) { return a.v < b.v || (a.v == b.v && a.i < b.i); }
__device__ __forceinline__ void insert(Cand& best, Cand& sec, const Cand c) {
  if (better(c, best)) {
    sec = best;
    best = c;
  } else if (better(c, sec)) {
    sec = c;
  }
}

struct VqArgs {
  const float* z;        // [rows][D]
  const float* E;        // [K][D] (fp64 re-rank reads it)
  const float* wimg;     // codebook image, rows padded to 512 (zeros)
  const float* ee;       // [K] ||e||^2
  int64_t* idx;
  int64_t rows;
  int D, K, nchunk;
};

__global__ void __launch_bounds__(THREADS, 1) t3_vq_kernel(const VqArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NT = 256;                                // codes per accumulator
  constexpr int W_HALF = 8 * NT * 16;                    // bytes of the hi (or lo) image of one 256-code tile of a chunk
  uint8_t* sAh = smem;
  uint8_t* sAl = smem + A_HALF;
  uint8_t* sW = smem + 2 * A_HALF;                       // [2 code tiles][hi | lo]
  float* xs0 = reinterpret_cast<float*>(sW + 4 * W_HALF);
  float* szz = xs0 + 2 * TM * XS_LD;                     // [2][128] partial ||z||^2
  Cand* scand = reinterpret_cast<Cand*>(szz + 2 * TM);   // [128][2] best / second of the upper code tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(scand + 2 * TM);
  uint64_t* bar_w = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * TM;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  uint32_t ph_w = 0, ph_m = 0;
  float zz = 0.f;

  stage_rows(a.z, row0, a.rows, a.D, a.D, 0, xs0);
  for (int c = 0; c < a.nchunk; ++c) {
    const float* xs = xs0 + (c & 1) * TM * XS_LD;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (c + 1 < a.nchunk) stage_rows(a.z, row0, a.rows, a.D, a.D, c + 1, xs0 + ((c + 1) & 1) * TM * XS_LD);
    if (c > 0) {
      mbar_wait(bar_mma, ph_m);
      ph_m ^= 1;
      tc_fence_after();
    }
    if (tid == 0) {                                      // both 256-code tiles of this chunk: [tile][hi | lo] in the image too
      mbar_expect_tx(bar_w, 4 * W_HALF);
      bulk_g2s(sW, a.wimg + (int64_t)c * (4 * W_HALF / 4), 4 * W_HALF, bar_w);
    }
    zz += split_rows(xs, sAh, sAl);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_w, ph_w);
      tc_fence_after();
      const int nks = min(KC, a.D - KC * c) / 8;
      for (int nt = 0; nt < 2; ++nt)
        issue_chunk(tmem + nt * NT, smem_u32(sAh), smem_u32(sAl), smem_u32(sW) + nt * 2 * W_HALF, smem_u32(sW) + nt * 2 * W_HALF + W_HALF,
                    nks, NT, c > 0);
      umma_commit(bar_mma);
    }
    ph_w ^= 1;
  }
  szz[(tid >> 7) * TM + (tid & (TM - 1))] = zz;
  mbar_wait(bar_mma, ph_m);
  tc_fence_after();
  __syncthreads();

  // ---- read-out: thread = (row r, code tile nt); best / second-best of 256 codes ------------------------------------
  const int lq = warp & 3, nt = warp >> 2;
  const int r = lq * 32 + lane;
  const int64_t row = row0 + r;
  const float zzr = szz[r] + szz[TM + r];
  Cand best = {INFINITY, 0x7fffffff}, sec = {INFINITY, 0x7fffffff};
  const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16) + nt * NT;
  for (int c16 = 0; c16 < NT; c16 += 16) {
    const int code0 = nt * NT + c16;
    if (code0 >= a.K) break;                             // warp-uniform
    float acc[16];
    tmem_ld16(trow + c16, acc);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int code = code0 + j;
      if (code < a.K) {
        // vq.py:75-79: (||z||^2 - (2 z) @ E^T) + ||e||^2 ; the factor 2 is exact
        const float d = __fadd_rn(__fsub_rn(zzr, 2.0f * acc[j]), __ldg(a.ee + code));
        insert(best, sec, Cand{d, code});
      }
    }
  }
  if (nt == 1) {
    scand[2 * r] = best;
    scand[2 * r + 1] = sec;
  }
  tc_fence_before();
  __syncthreads();
  if (nt == 0 && row < a.rows) {
    insert(best, sec, scand[2 * r]);
    insert(best, sec, scand[2 * r + 1]);
    int winner = best.i;
    const float tol = 4e-3f + 1e-5f * fabsf(best.v);     // the fp32 distance (magnitude ~ 2 D) carries ~1e-5 of rounding noise
    if (sec.i != 0x7fffffff && (sec.v - best.v) <= tol) {
      double d1 = 0.0, d2 = 0.0;
      for (int d = 0; d < a.D; ++d) {
        const double zv = (double)a.z[row * a.D + d];
        const double e1 = zv - (double)a.E[(int64_t)best.i * a.D + d];
        const double e2 = zv - (double)a.E[(int64_t)sec.i * a.D + d];
        d1 = fma(e1, e1, d1);
        d2 = fma(e2, e2, d2);
      }
      if (d2 < d1 || (d2 == d1 && sec.i < best.i)) winner = sec.i;
    }
    a.idx[row] = (int64_t)winner;
  }
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// codebook image with the two 256-code tiles of a chunk adjacent: [chunk][tile][hi | lo][slab][256][4]; codes >= K are zero
__global__ void pack_codebook_kernel(const float* __restrict__ E, float* __restrict__ img, int D, int K, int nchunk) {
  const int total = nchunk * 2 * 8 * 256 * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 3, n = (i >> 2) & 255, s = (i >> 10) & 7, t = (i >> 13) & 1, c = i >> 14;
    const int code = 256 * t + n, kk = KC * c + 4 * s + j;
    const float w = (code < K && kk < D) ? E[(int64_t)code * D + kk] : 0.f;
    const float hi = tf32_rna(w);
    const int64_t base = ((int64_t)c * 2 + t) * (2 * 8 * 256 * 4);
    img[base + ((int64_t)s * 256 + n) * 4 + j] = hi;
    img[base + 8 * 256 * 4 + ((int64_t)s * 256 + n) * 4 + j] = w - hi;
  }
}

static int vq_smem() { return 2 * A_HALF + 4 * 8 * 256 * 16 + 2 * TM * XS_LD * 4 + 2 * TM * 4 + 2 * TM * 8 + 64; }

bool t3_vq_ok(int D, int K) {
  static const bool off = getenv("EDTTS_VQ_SIMT") != nullptr;     // development: force the CUDA-core route
  return !off && D % 8 == 0 && D >= 8 && D <= 1024 && K >= 1 && K <= 512;
}
int64_t t3_vq_image_bytes(int D) { return (int64_t)((D + KC - 1) / KC) * 4 * 8 * 256 * 16; }

int pack_codebook(const float* E, float* wimg, int D, int K, cudaStream_t st) {
  const int nchunk = (D + KC - 1) / KC;
  LaunchScope ls(KC_VQ, st);
  const int total = nchunk * 2 * 8 * 256 * 4;
  pack_codebook_kernel<<<(total + 255) / 256, 256, 0, st>>>(E, wimg, D, K, nchunk);
  return check_launch("pack_codebook");
}

int launch_t3_vq_packed(const float* z, const float* E, const float* ee, const float* wimg, int64_t* idx, int64_t rows, int D, int K,
                        cudaStream_t st) {
  VqArgs a;
  a.z = z; a.E = E; a.wimg = wimg; a.ee = ee; a.idx = idx; a.rows = rows; a.D = D; a.K = K;
  a.nchunk = (D + KC - 1) / KC;
  static PerDeviceOnce configured;
  if (configured.need()) {
    if (cudaFuncSetAttribute(t3_vq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vq_smem()) != cudaSuccess)
      return check_launch("t3_vq smem attribute");
    configured.set();
  }
  LaunchScope ls(KC_VQ, st);
  t3_vq_kernel<<<(unsigned)((rows + TM - 1) / TM), THREADS, vq_smem(), st>>>(a);
  return check_launch("t3_vq");
}

int launch_t3_vq(const float* z, const float* E, const float* ee, float* wimg, int64_t* idx, int64_t rows, int D, int K, cudaStream_t st) {
  if (int rc = pack_codebook(E, wimg, D, K, st)) return rc;
  return launch_t3_vq_packed(z, E, ee, wimg, idx, rows, D, K, st);
}

}  // namespace t3
}  // namespace edtts

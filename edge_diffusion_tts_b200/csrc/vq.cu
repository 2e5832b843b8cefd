// VectorQuantizer nearest-codeword search (models/vq.py:75-83, :148-159).
//
// Fused distance GEMM + argmin: a block owns 128 rows of z, walks the codebook in
// 64-code tiles with an fp32 FFMA register-tiled product (K = dim sliced by 16),
// forms d = (||z||^2 - 2 z.e) + ||e||^2 in the reference's operation order and
// keeps, per row, the best and second-best (distance, index) pair; the 16 lanes
// that share a row merge their pairs with warp shuffles (lowest index wins a tie,
// i.e. torch.argmin's first minimum).  The [rows, K] distance matrix of the
// reference (393 MB at BASELINE config 5) is never written.
//
// Bit-exactness: fp32 distances of magnitude ~170 carry ~1e-5 of rounding noise,
// so when best and second-best are closer than a conservative bound the two
// candidates are re-ranked with fp64 distances.  The result is the exact argmin;
// the reference's fp32 argmin equals it on every seed tested
// (tests/test_gpu_parity.py::test_vq_indices_bit_exact_vs_oracle_fp32_and_fp64 checks against both the fp32
// oracle and its fp64 restatement; tests/test_gpu_pinned_configs.py::test_raw_features_to_vq_indices the chain from raw features).
#include "common.cuh"
#include "tf32x3.cuh"

namespace edtts {

constexpr int VQ_BM = 128, VQ_BK = 16, VQ_TN = 4, VQ_BN = 64, VQ_THREADS = 256;
constexpr int VQ_ASLD = VQ_BM + 4, VQ_WSLD = VQ_BN + 4;

__global__ void vq_code_norms_kernel(const float* __restrict__ E, float* __restrict__ ee, int K, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= K) return;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = E[(int64_t)warp * D + d];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) ee[warp] = s;
}

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

__global__ void __launch_bounds__(VQ_THREADS) vq_argmin_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                               const float* __restrict__ ee,
                                                               int64_t* __restrict__ idx_out, int64_t rows, int D,
                                                               int K) {
  __shared__ __align__(16) float As[VQ_BK * VQ_ASLD];
  __shared__ __align__(16) float Ws[VQ_BK * VQ_WSLD];
  __shared__ float s_zz[VQ_BM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * VQ_BM;

  {  // ||z||^2 per row: a warp per row
    const int warp = tid >> 5;
    for (int r = warp; r < VQ_BM; r += VQ_THREADS / 32) {
      const int64_t row = row0 + r;
      float s = 0.f;
      if (row < rows)
        for (int d = lane; d < D; d += 32) {
          const float v = z[row * D + d];
          s = fmaf(v, v, s);
        }
      s = warp_sum(s);
      if (lane == 0) s_zz[r] = s;
    }
  }
  __syncthreads();

  Cand best[8], sec[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    best[i] = {INFINITY, 0x7fffffff};
    sec[i] = {INFINITY, 0x7fffffff};
  }

  for (int n0 = 0; n0 < K; n0 += VQ_BN) {
    float acc[8][VQ_TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < VQ_TN; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += VQ_BK) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int id = tid + it * VQ_THREADS;
        const int r = id >> 2, kq = (id & 3) * 4;
        const int64_t row = row0 + r;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) a = *reinterpret_cast<const float4*>(z + row * D + k0 + kq);
        As[(kq + 0) * VQ_ASLD + r] = a.x;
        As[(kq + 1) * VQ_ASLD + r] = a.y;
        As[(kq + 2) * VQ_ASLD + r] = a.z;
        As[(kq + 3) * VQ_ASLD + r] = a.w;
      }
      {
        const int n = tid >> 2, kq = (tid & 3) * 4;   // 64 codes x 4 float4 = 256 threads
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + n < K) w = *reinterpret_cast<const float4*>(E + (int64_t)(n0 + n) * D + k0 + kq);
        Ws[(kq + 0) * VQ_WSLD + n] = w.x;
        Ws[(kq + 1) * VQ_WSLD + n] = w.y;
        Ws[(kq + 2) * VQ_WSLD + n] = w.z;
        Ws[(kq + 3) * VQ_WSLD + n] = w.w;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < VQ_BK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk * VQ_ASLD + ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk * VQ_ASLD + ty * 8 + 4]);
        const float4 bb = *reinterpret_cast<const float4*>(&Ws[kk * VQ_WSLD + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < VQ_TN; ++j) {
      const int code = n0 + tx * 4 + j;
      if (code < K) {
        const float e2 = ee[code];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // vq.py:75-79: (||z||^2 - (2 z) @ E^T) + ||e||^2 ; the factor 2 is exact
          const float d = __fadd_rn(__fsub_rn(s_zz[ty * 8 + i], 2.0f * acc[i][j]), e2);
          insert(best[i], sec[i], Cand{d, code});
        }
      }
    }
  }

  // merge the 16 lanes that share a row, then fp64 re-rank of near ties
  const unsigned gmask = 0xffffu << (lane & 16);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    Cand b = best[i], s = sec[i];
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      Cand ob, os;
      ob.v = __shfl_xor_sync(gmask, b.v, o);
      ob.i = __shfl_xor_sync(gmask, b.i, o);
      os.v = __shfl_xor_sync(gmask, s.v, o);
      os.i = __shfl_xor_sync(gmask, s.i, o);
      if (better(ob, b)) {
        s = better(b, os) ? b : os;
        b = ob;
      } else {
        s = better(ob, s) ? ob : s;
      }
    }
    const int64_t row = row0 + ty * 8 + i;
    int winner = b.i;
    const float tol = 4e-3f + 1e-5f * fabsf(b.v);
    if (row < rows && s.i != 0x7fffffff && (s.v - b.v) <= tol) {   // group-uniform condition
      double d1 = 0.0, d2 = 0.0;
      for (int d = tx; d < D; d += 16) {
        const double zv = (double)z[row * D + d];
        const double e1 = zv - (double)E[(int64_t)b.i * D + d];
        const double e2 = zv - (double)E[(int64_t)s.i * D + d];
        d1 = fma(e1, e1, d1);
        d2 = fma(e2, e2, d2);
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        d1 += __shfl_xor_sync(gmask, d1, o);
        d2 += __shfl_xor_sync(gmask, d2, o);
      }
      if (d2 < d1 || (d2 == d1 && s.i < b.i)) winner = s.i;
    }
    if (tx == 0 && row < rows) idx_out[row] = (int64_t)winner;
  }
}

__global__ void vq_gather_ste_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                     const int64_t* __restrict__ idx, float* __restrict__ zq, int64_t rows, int D,
                                     int K) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  const int64_t r = i / D;
  const int d = (int)(i % D);
  int64_t c = idx[r];
  c = c < 0 ? 0 : (c >= K ? K - 1 : c);
  const float zv = z[i];
  zq[i] = __fadd_rn(zv, __fsub_rn(E[c * D + d], zv));   // vq.py:98
}

__global__ void vq_bincount_kernel(const int64_t* __restrict__ idx, int32_t* __restrict__ counts, int64_t rows, int K) {
  extern __shared__ int32_t hist[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx[i];
    if (c >= 0 && c < K) atomicAdd(&hist[c], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    if (hist[k]) atomicAdd(&counts[k], hist[k]);
}

}  // namespace edtts

using namespace edtts;

extern "C" int64_t edtts_vq_workspace_bytes(int32_t codebook_size, int32_t dim) {
  return align_up((int64_t)codebook_size * 4, 256) + (t3::t3_vq_ok(dim, codebook_size) ? t3::t3_vq_image_bytes(dim) : 0);
}

extern "C" int edtts_vq_argmin(const float* z, const float* codebook, int64_t* idx_out, int64_t rows, int32_t dim,
                               int32_t codebook_size, void* workspace, void* stream) {
  EDTTS_REQUIRE(z && codebook && idx_out && workspace, EDTTS_EINVAL, "vq_argmin: null argument");
  EDTTS_REQUIRE(dim > 0 && dim % VQ_BK == 0 && codebook_size > 0, EDTTS_EINVAL,
                "vq_argmin: dim=%d must be a positive multiple of 16, codebook_size=%d", dim, codebook_size);
  if (rows == 0) return EDTTS_OK;
  cudaStream_t st = as_stream(stream);
  float* ee = reinterpret_cast<float*>(workspace);
  int rc;
  {
    LaunchScope ls(KC_VQ, st);
    vq_code_norms_kernel<<<(codebook_size * 32 + 255) / 256, 256, 0, st>>>(codebook, ee, codebook_size, dim);
    rc = check_launch("vq_code_norms");
  }
  if (rc) return rc;
  if (t3::t3_vq_ok(dim, codebook_size))       // distance product on the tensor cores (tf32 x 3), same candidates / re-rank
    return t3::launch_t3_vq(z, codebook, ee, reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align_up((int64_t)codebook_size * 4, 256)),
                            idx_out, rows, dim, codebook_size, st);
  LaunchScope ls(KC_VQ, st);
  vq_argmin_kernel<<<(unsigned)((rows + VQ_BM - 1) / VQ_BM), VQ_THREADS, 0, st>>>(z, codebook, ee, idx_out, rows, dim,
                                                                                 codebook_size);
  return check_launch("vq_argmin");
}

// Code norms + the tensor-core codebook image, packed once per codebook: packed_out holds edtts_vq_workspace_bytes(K, D) bytes.
extern "C" int edtts_vq_pack(const float* codebook, int32_t dim, int32_t codebook_size, void* packed_out, void* stream) {
  EDTTS_REQUIRE(codebook && packed_out && dim > 0 && dim % VQ_BK == 0 && codebook_size > 0, EDTTS_EINVAL, "vq_pack: bad argument");
  cudaStream_t st = as_stream(stream);
  float* ee = reinterpret_cast<float*>(packed_out);
  {
    LaunchScope ls(KC_VQ, st);
    vq_code_norms_kernel<<<(codebook_size * 32 + 255) / 256, 256, 0, st>>>(codebook, ee, codebook_size, dim);
    if (int rc = check_launch("vq_code_norms")) return rc;
  }
  if (!t3::t3_vq_ok(dim, codebook_size)) return EDTTS_OK;
  return t3::pack_codebook(codebook, reinterpret_cast<float*>(reinterpret_cast<char*>(packed_out) + align_up((int64_t)codebook_size * 4, 256)), dim,
                           codebook_size, st);
}
// edtts_vq_argmin with the norms / image packed earlier by edtts_vq_pack (one kernel launch)
extern "C" int edtts_vq_argmin_packed(const float* z, const float* codebook, const void* packed, int64_t* idx_out, int64_t rows,
                                      int32_t dim, int32_t codebook_size, void* stream) {
  EDTTS_REQUIRE(z && codebook && packed && idx_out, EDTTS_EINVAL, "vq_argmin_packed: null argument");
  EDTTS_REQUIRE(dim > 0 && dim % VQ_BK == 0 && codebook_size > 0, EDTTS_EINVAL, "vq_argmin_packed: dim=%d codebook_size=%d", dim, codebook_size);
  if (rows == 0) return EDTTS_OK;
  cudaStream_t st = as_stream(stream);
  const float* ee = reinterpret_cast<const float*>(packed);
  if (t3::t3_vq_ok(dim, codebook_size))
    return t3::launch_t3_vq_packed(z, codebook, ee, reinterpret_cast<const float*>(reinterpret_cast<const char*>(packed) + align_up((int64_t)codebook_size * 4, 256)),
                                   idx_out, rows, dim, codebook_size, st);
  LaunchScope ls(KC_VQ, st);
  vq_argmin_kernel<<<(unsigned)((rows + VQ_BM - 1) / VQ_BM), VQ_THREADS, 0, st>>>(z, codebook, ee, idx_out, rows, dim, codebook_size);
  return check_launch("vq_argmin");
}

extern "C" int edtts_vq_gather_ste(const float* z, const float* codebook, const int64_t* idx, float* zq_out,
                                   int64_t rows, int32_t dim, int32_t codebook_size, void* stream) {
  EDTTS_REQUIRE(z && codebook && idx && zq_out, EDTTS_EINVAL, "vq_gather_ste: null argument");
  if (rows == 0) return EDTTS_OK;
  const int64_t n = rows * dim;
  LaunchScope ls(KC_VQ, as_stream(stream));
  vq_gather_ste_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(z, codebook, idx, zq_out, rows, dim,
                                                                                  codebook_size);
  return check_launch("vq_gather_ste");
}

extern "C" int edtts_vq_bincount(const int64_t* idx, int32_t* counts_out, int64_t rows, int32_t codebook_size,
                                 void* stream) {
  EDTTS_REQUIRE(idx && counts_out && codebook_size > 0 && codebook_size <= 8192, EDTTS_EINVAL,
                "vq_bincount: bad argument");
  cudaStream_t st = as_stream(stream);
  if (cudaMemsetAsync(counts_out, 0, (size_t)codebook_size * 4, st) != cudaSuccess) return check_launch("memset");
  if (rows == 0) return EDTTS_OK;
  const int blocks = (int)((rows + 256 * 8 - 1) / (256 * 8));
  LaunchScope ls(KC_VQ, st);
  vq_bincount_kernel<<<blocks < 1 ? 1 : (blocks > 1184 ? 1184 : blocks), 256, (size_t)codebook_size * 4, st>>>(
      idx, counts_out, rows, codebook_size);
  return check_launch("vq_bincount");
}

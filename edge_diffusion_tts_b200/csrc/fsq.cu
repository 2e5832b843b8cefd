// Finite scalar quantisation (models/fsq.py:18-132): bound (tanh) -> per-dimension rounding to `levels` values ->
// straight-through value -> flat index, one pass over z.  Pure streaming kernel: reads z (rows x dim fp32) once, writes
// z_q and one int64 index per row; a thread handles one row (dim <= 8), grid sized in multiples of the SM count.
#include "common.cuh"

namespace edtts {

constexpr int FSQ_MAX_DIM = 8;

struct FsqParams {
  int dim;
  int levels[FSQ_MAX_DIM];
  long long basis[FSQ_MAX_DIM];
};

// mode 0: forward (z -> z_q, idx); mode 1: codes_to_indices (z is already quantised codes in [-1, 1])
template <int MODE>
__global__ void __launch_bounds__(256) fsq_kernel(const float* __restrict__ z, FsqParams p, float* __restrict__ zq,
                                                  int64_t* __restrict__ idx, int64_t rows) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    long long flat = 0;
#pragma unroll
    for (int d = 0; d < FSQ_MAX_DIM; ++d) {
      if (d >= p.dim) break;
      const float lv = (float)p.levels[d];
      const float half = __fdiv_rn(__fsub_rn(lv, 1.0f), 2.0f);          // (levels - 1) / 2          fsq.py:70
      float q;                                                          // the value codes_to_indices sees
      if (MODE == 0) {
        const float zb = tanhf(z[r * p.dim + d]);                       // bound()                  fsq.py:59-61
        float s = rintf(__fmul_rn(__fadd_rn(zb, 1.0f), half));          // round((z + 1) * half)    fsq.py:73-74
        s = fminf(fmaxf(s, 0.0f), __fsub_rn(lv, 1.0f));                 // clamp to [0, L - 1]      fsq.py:77-79
        const float out = __fsub_rn(__fdiv_rn(s, half), 1.0f);          // back to [-1, 1]          fsq.py:82
        q = __fadd_rn(zb, __fsub_rn(out, zb));                          // straight-through value   fsq.py:104
        zq[r * p.dim + d] = q;
      } else {
        q = z[r * p.dim + d];
      }
      const long long code = (long long)rintf(__fmul_rn(__fadd_rn(q, 1.0f), half));   // fsq.py:115
      flat += code * p.basis[d];                                        // first dimension fastest  fsq.py:118
    }
    if (idx) idx[r] = flat;
  }
}

// indices_to_codes (fsq.py:121-132): decodes with the LAST dimension fastest -- not the inverse of the basis above unless
// all levels are equal; mirrored as the reference has it.
__global__ void __launch_bounds__(256) fsq_decode_kernel(const int64_t* __restrict__ idx, FsqParams p,
                                                         float* __restrict__ codes, int64_t rows) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    long long v = idx[r];
    for (int d = p.dim - 1; d >= 0; --d) {
      const long long L = p.levels[d];
      long long m = v % L;                                              // torch remainder / floor division semantics
      if (m < 0) m += L;
      v = (v - m) / L;
      const float half = __fdiv_rn(__fsub_rn((float)L, 1.0f), 2.0f);
      codes[r * p.dim + d] = __fsub_rn(__fdiv_rn((float)m, half), 1.0f);
    }
  }
}

// FSQEncoder (models/fsq.py:135-222) in one pass: z [rows, D] -> z_low = proj_down(z) (D -> dim <= 8) -> FSQ -> z_q =
// proj_up(z_q_low) (dim -> D), plus the flat index; or, with idx_in, the decode path indices_to_codes -> proj_up.
// A warp per row: lanes read the row as coalesced float4 pieces (512 B per row at D = 128), the `dim` dot products are
// reduced with shuffles (fixed order: deterministic), every lane quantises all `dim` values redundantly and then writes its
// own float4 pieces of z_q.  Both weight matrices (2 x dim x D floats, 8 KB at D = 128) sit in shared memory.  HBM-bound:
// 2 x rows x D x 4 bytes.
__global__ void __launch_bounds__(256) fsq_encoder_kernel(const float* __restrict__ z, const int64_t* __restrict__ idx_in,
                                                          const float* __restrict__ wd, const float* __restrict__ bd,
                                                          const float* __restrict__ wu, const float* __restrict__ bu, FsqParams p,
                                                          int D, float* __restrict__ zq, int64_t* __restrict__ idx_out,
                                                          int64_t rows) {
  extern __shared__ float sm[];
  float* swd = sm;                       // [dim][D]
  float* swu = sm + p.dim * D;           // [dim][D]: transposed copy of nn.Linear(dim, D).weight (conflict-free lane reads)
  float* sbu = swu + p.dim * D;          // [D]
  for (int i = threadIdx.x; i < p.dim * D; i += blockDim.x) {
    swd[i] = wd ? wd[i] : 0.f;
    swu[(i % p.dim) * D + i / p.dim] = wu[i];
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) sbu[i] = bu[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int D4 = D >> 2;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    float q[FSQ_MAX_DIM];
    long long flat = 0;
    if (idx_in == nullptr) {
      float acc[FSQ_MAX_DIM];
#pragma unroll
      for (int d = 0; d < FSQ_MAX_DIM; ++d) acc[d] = 0.f;
      for (int c = lane; c < D4; c += 32) {
        const float4 v = *reinterpret_cast<const float4*>(z + r * D + 4 * c);
#pragma unroll
        for (int d = 0; d < FSQ_MAX_DIM; ++d) {
          if (d >= p.dim) break;
          const float4 w = *reinterpret_cast<const float4*>(swd + d * D + 4 * c);
          acc[d] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[d]))));
        }
      }
      // lane d quantises dimension d (tanh, rounding, straight-through value); the 8 results travel back by shuffles
      float mine = 0.f;
#pragma unroll
      for (int d = 0; d < FSQ_MAX_DIM; ++d) {
        if (d >= p.dim) break;
        const float sd = warp_sum(acc[d]);
        mine = lane == d ? sd : mine;
      }
      float q_mine = 0.f;
      int code_mine = 0;
      if (lane < p.dim) {
        const float lv = (float)p.levels[lane];
        const float half = __fdiv_rn(__fsub_rn(lv, 1.0f), 2.0f);
        const float zb = tanhf(mine + bd[lane]);
        float s = rintf(__fmul_rn(__fadd_rn(zb, 1.0f), half));
        s = fminf(fmaxf(s, 0.0f), __fsub_rn(lv, 1.0f));
        const float out = __fsub_rn(__fdiv_rn(s, half), 1.0f);
        q_mine = __fadd_rn(zb, __fsub_rn(out, zb));                       // straight-through value, fsq.py:104
        code_mine = (int)rintf(__fmul_rn(__fadd_rn(q_mine, 1.0f), half));
      }
#pragma unroll
      for (int d = 0; d < FSQ_MAX_DIM; ++d) {
        if (d >= p.dim) break;
        q[d] = __shfl_sync(0xffffffffu, q_mine, d);
        flat += (long long)__shfl_sync(0xffffffffu, code_mine, d) * p.basis[d];
      }
      if (idx_out && lane == 0) idx_out[r] = flat;
    } else {
      long long v = idx_in[r];
      for (int d = p.dim - 1; d >= 0; --d) {                              // indices_to_codes, fsq.py:121-132
        const long long L = p.levels[d];
        long long m = v % L;
        if (m < 0) m += L;
        v = (v - m) / L;
        const float half = __fdiv_rn(__fsub_rn((float)L, 1.0f), 2.0f);
        q[d] = __fsub_rn(__fdiv_rn((float)m, half), 1.0f);
      }
    }
    if (zq) {
      for (int c = lane; c < D4; c += 32) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = sbu[4 * c + j];
#pragma unroll
          for (int d = 0; d < FSQ_MAX_DIM; ++d) {
            if (d >= p.dim) break;
            a = fmaf(q[d], swu[d * D + 4 * c + j], a);
          }
          o[j] = a;
        }
        *reinterpret_cast<float4*>(zq + r * D + 4 * c) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

static int fsq_params(const int32_t* levels, int32_t dim, FsqParams& p) {
  EDTTS_REQUIRE(levels && dim >= 1 && dim <= FSQ_MAX_DIM, EDTTS_EINVAL, "fsq: dim=%d (1..%d)", dim, FSQ_MAX_DIM);
  p.dim = dim;
  long long b = 1;
  for (int d = 0; d < dim; ++d) {
    EDTTS_REQUIRE(levels[d] >= 2, EDTTS_EINVAL, "fsq: levels[%d]=%d", d, levels[d]);
    p.levels[d] = levels[d];
    p.basis[d] = b;                                                     // cumprod([1] + levels[:-1])  fsq.py:44-49
    b *= levels[d];
  }
  return EDTTS_OK;
}
static inline unsigned fsq_grid(int64_t rows) {
  const int64_t blocks = (rows + 255) / 256, cap = stream_grid_cap(8);
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace edtts

using namespace edtts;

extern "C" int edtts_fsq_forward(const float* z, const int32_t* levels_host, int32_t dim, int32_t codes_only, float* zq_out,
                                 int64_t* idx_out, int64_t rows, void* stream) {
  FsqParams p;
  if (int rc = fsq_params(levels_host, dim, p)) return rc;
  if (rows == 0) return EDTTS_OK;
  EDTTS_REQUIRE(z && rows > 0 && (codes_only ? idx_out != nullptr : zq_out != nullptr), EDTTS_EINVAL, "fsq_forward: null argument");
  LaunchScope ls(KC_VQ, as_stream(stream));
  if (codes_only) fsq_kernel<1><<<fsq_grid(rows), 256, 0, as_stream(stream)>>>(z, p, nullptr, idx_out, rows);
  else fsq_kernel<0><<<fsq_grid(rows), 256, 0, as_stream(stream)>>>(z, p, zq_out, idx_out, rows);
  return check_launch("fsq_forward");
}

extern "C" int edtts_fsq_decode(const int64_t* idx, const int32_t* levels_host, int32_t dim, float* codes_out, int64_t rows,
                                void* stream) {
  FsqParams p;
  if (int rc = fsq_params(levels_host, dim, p)) return rc;
  if (rows == 0) return EDTTS_OK;
  EDTTS_REQUIRE(idx && codes_out && rows > 0, EDTTS_EINVAL, "fsq_decode: null argument");
  LaunchScope ls(KC_VQ, as_stream(stream));
  fsq_decode_kernel<<<fsq_grid(rows), 256, 0, as_stream(stream)>>>(idx, p, codes_out, rows);
  return check_launch("fsq_decode");
}

extern "C" int edtts_fsq_encoder(const float* z, const int64_t* idx_in, const float* w_down, const float* b_down,
                                 const float* w_up, const float* b_up, const int32_t* levels_host, int32_t dim, int32_t in_dim,
                                 float* zq_out, int64_t* idx_out, int64_t rows, void* stream) {
  FsqParams p;
  if (int rc = fsq_params(levels_host, dim, p)) return rc;
  if (rows == 0) return EDTTS_OK;
  EDTTS_REQUIRE(rows > 0 && w_up && b_up && (idx_in != nullptr || (z && w_down && b_down)) && (zq_out || idx_out), EDTTS_EINVAL,
                "fsq_encoder: null argument");
  EDTTS_REQUIRE(in_dim >= 4 && in_dim % 4 == 0 && in_dim <= 1024, EDTTS_ENOTSUP, "fsq_encoder: in_dim=%d (multiple of 4, <= 1024)",
                in_dim);
  const int smem = (2 * dim * in_dim + in_dim) * (int)sizeof(float);
  static PerDeviceOnce configured;
  if (smem > 48 * 1024 && configured.need()) {
    if (cudaFuncSetAttribute(fsq_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * FSQ_MAX_DIM * 1024 * 4 + 4096) !=
        cudaSuccess)
      return check_launch("fsq_encoder smem attribute");
    configured.set();
  }
  const int64_t blocks = (rows + 7) / 8, cap = stream_grid_cap(4);
  LaunchScope ls(KC_VQ, as_stream(stream));
  fsq_encoder_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, smem, as_stream(stream)>>>(
      z, idx_in, w_down, b_down, w_up, b_up, p, in_dim, zq_out, idx_out, rows);
  return check_launch("fsq_encoder");
}

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
  const int64_t blocks = (rows + 255) / 256, cap = 148 * 8;
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

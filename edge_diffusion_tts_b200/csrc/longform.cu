// Long-form pipeline around the sampling loop:
//   * normalize_mel / denormalize_mel            (reference edge_diffusion_tts/utils/audio.py:10-19)
//   * cross-fade stitch of generated chunks      (reference inference_pipeline.py:359-375)
//   * weight normalisation, trim, 5x3 smoothing  (reference inference_pipeline.py:377-393)
// All of it is streaming fp32 work: HBM-bound, one read and one write per element, coalesced on the
// contiguous axis of each side (the stitch transposes [T, mel] -> [mel, frames] through a padded
// shared-memory tile so both the read and the read-modify-write are 128-byte row segments).
#include "common.cuh"

namespace edtts {

// ---- per-(utterance, mel bin) mean and unbiased standard deviation over the frames --------------------
// mel [B, T, M] (M contiguous).  Block = 32 bins x 8 frame lanes; every warp reads 32 consecutive bins of a frame
// (one 128-byte segment).  Two passes (mean, then centred squares) with fp64 accumulators: the result is the
// correctly rounded statistic up to the final fp32 rounding, whatever summation order torch used for its own.
constexpr int ST_BINS = 32, ST_LANES = 8;

__global__ void __launch_bounds__(ST_BINS* ST_LANES) mel_stats_kernel(const float* __restrict__ mel, float* __restrict__ mean_out,
                                                                     float* __restrict__ std_out, int T, int Mb) {
  __shared__ double red[ST_LANES][ST_BINS];
  __shared__ double mean_s[ST_BINS];
  const int b = blockIdx.y, m = blockIdx.x * ST_BINS + threadIdx.x, ly = threadIdx.y;
  const float* src = mel + (int64_t)b * T * Mb;
  const bool live = m < Mb;
  double s = 0.0;
  if (live)
    for (int t = ly; t < T; t += ST_LANES) s += (double)__ldg(src + (int64_t)t * Mb + m);
  red[ly][threadIdx.x] = s;
  __syncthreads();
  if (ly == 0) {
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < ST_LANES; ++i) tot += red[i][threadIdx.x];
    mean_s[threadIdx.x] = tot / (double)T;
  }
  __syncthreads();
  const double mu = mean_s[threadIdx.x];
  double q = 0.0;
  if (live)
    for (int t = ly; t < T; t += ST_LANES) {
      const double d = (double)__ldg(src + (int64_t)t * Mb + m) - mu;
      q += d * d;
    }
  __syncthreads();
  red[ly][threadIdx.x] = q;
  __syncthreads();
  if (ly == 0 && live) {
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < ST_LANES; ++i) tot += red[i][threadIdx.x];
    // torch.std default: Bessel's correction; T == 1 gives 0/0 = NaN exactly as torch does, and clamp_min keeps NaN
    const float sd = (float)sqrt(tot / (double)(T - 1));
    mean_out[(int64_t)b * Mb + m] = (float)mu;
    std_out[(int64_t)b * Mb + m] = sd < 1e-5f ? 1e-5f : sd;      // NaN < x is false -> NaN propagates (clamp_min)
  }
}

// (mel - mean) / std and mel_n * std + mean: the reference's operation order, no FMA contraction -> bit-exact.
template <bool kDenorm>
__global__ void __launch_bounds__(256) mel_affine_kernel(const float* __restrict__ in, const float* __restrict__ mean,
                                                         const float* __restrict__ sd, float* __restrict__ out, int64_t total,
                                                         int64_t per_b, int Mb) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per_b;
    const int m = (int)(i % Mb);
    const float mu = __ldg(mean + b * Mb + m), s = __ldg(sd + b * Mb + m);
    const float v = __ldcs(in + i);
    __stcs(out + i, kDenorm ? __fadd_rn(__fmul_rn(v, s), mu) : __fdiv_rn(__fsub_rn(v, mu), s));
  }
}

// ---- cross-fade stitch: final_mel[b, m, start + t] += exp(x[b, t, m] * std[b, m] + mean[b, m]) * window[t] ----
// One 32 x 32 (frames x bins) tile per block; the block with blockIdx.y == 0 also accumulates the window weights.
__global__ void __launch_bounds__(256) stitch_add_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                         const float* __restrict__ sd, const float* __restrict__ window,
                                                         float* __restrict__ final_mel, float* __restrict__ final_w, int T,
                                                         int Mb, int64_t F, int64_t start) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 8 rows per sweep
  const float* xb = x + (int64_t)b * T * Mb;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, m = m0 + tx;
    float v = 0.f;
    if (t < T && m < Mb) {
      const float den = __fadd_rn(__fmul_rn(__ldcs(xb + (int64_t)t * Mb + m), __ldg(sd + (int64_t)b * Mb + m)),
                                  __ldg(mean + (int64_t)b * Mb + m));
      v = __fmul_rn(expf(den), __ldg(window + t));
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  float* fm = final_mel + (int64_t)b * Mb * F;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, t = t0 + tx;
    if (m < Mb && t < T && start + t < F) {
      float* p = fm + (int64_t)m * F + start + t;
      *p = __fadd_rn(*p, tile[tx][r]);
    }
  }
  if (blockIdx.y == 0 && b == 0 && threadIdx.x < 32) {
    const int t = t0 + threadIdx.x;
    if (t < T && start + t < F) final_w[start + t] = __fadd_rn(final_w[start + t], __ldg(window + t));
  }
}

// ---- final_mel / clamp(final_w, 1e-5), trimmed to `total` frames, optionally followed by the kh x kw average
// pooling (stride 1, zero padding, count_include_pad: always divided by kh * kw) of inference_pipeline.py:385-393.
// The pooled sum visits the window rows top to bottom and each row left to right, as ATen's CPU kernel does.
__global__ void __launch_bounds__(256) stitch_finalize_kernel(const float* __restrict__ final_mel, const float* __restrict__ final_w,
                                                              float* __restrict__ mel_out, float* __restrict__ smooth_out, int Mb,
                                                              int64_t F, int64_t total, int kh, int kw, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i % total;
    const int m = (int)((i / total) % Mb);
    const int64_t b = i / (total * Mb);
    const float* fm = final_mel + b * Mb * F;
    if (mel_out) mel_out[i] = __fdiv_rn(__ldg(fm + (int64_t)m * F + t), fmaxf(__ldg(final_w + t), 1e-5f));
    if (smooth_out) {
      float acc = 0.f;
      for (int dm = -(kh / 2); dm < kh - kh / 2; ++dm) {
        const int mm = m + dm;
        if (mm < 0 || mm >= Mb) continue;
        for (int dt = -(kw / 2); dt < kw - kw / 2; ++dt) {
          const int64_t tt = t + dt;
          if (tt < 0 || tt >= total) continue;
          acc = __fadd_rn(acc, __fdiv_rn(__ldg(fm + (int64_t)mm * F + tt), fmaxf(__ldg(final_w + tt), 1e-5f)));
        }
      }
      smooth_out[i] = __fdiv_rn(acc, (float)(kh * kw));
    }
  }
}

// ---- InverseMelScale: spec[b, f, t] = relu(sum_m P[f, m] * mel[b, m, t]),  P = pinv(fb^T)  [n_stft, n_mels] ----
// (torchaudio solves min ||fb^T X - mel|| per frame with lstsq and clamps at 0; for the full-rank filter bank that is the
// pseudo-inverse applied to every frame: one skinny GEMM, K = n_mels.)  64 x 64 output tile per block, 4 x 4 per thread,
// both operands staged in shared memory with coalesced row reads; HBM-bound on the [n_stft, T] write.
constexpr int IM_TF = 64, IM_TT = 64, IM_KC = 16;
__global__ void __launch_bounds__(256) inverse_mel_kernel(const float* __restrict__ P, const float* __restrict__ mel,
                                                          float* __restrict__ out, int n_stft, int n_mels, int64_t T) {
  __shared__ float sP[IM_KC][IM_TF + 1];     // [m][f]
  __shared__ __align__(16) float sM[IM_KC][IM_TT];         // [m][t]
  const int b = blockIdx.z, f0 = blockIdx.y * IM_TF;
  const int64_t t0 = (int64_t)blockIdx.x * IM_TT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // thread: frames t0 + 4 tx .. +3, bins f0 + 4 ty .. +3
  const float* melb = mel + (int64_t)b * n_mels * T;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < n_mels; k0 += IM_KC) {
    for (int i = threadIdx.x; i < IM_KC * IM_TF; i += 256) {   // P[f0 + f][k0 + m]: m fastest in memory
      const int f = i / IM_KC, m = i % IM_KC;
      sP[m][f] = (f0 + f < n_stft && k0 + m < n_mels) ? __ldg(P + (int64_t)(f0 + f) * n_mels + k0 + m) : 0.f;
    }
    for (int i = threadIdx.x; i < IM_KC * IM_TT; i += 256) {
      const int m = i / IM_TT, t = i % IM_TT;
      sM[m][t] = (k0 + m < n_mels && t0 + t < T) ? __ldcs(melb + (int64_t)(k0 + m) * T + t0 + t) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < IM_KC; ++m) {
      float pf[4], mt[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pf[i] = sP[m][4 * ty + i];
      const float4 mv = *reinterpret_cast<const float4*>(&sM[m][4 * tx]);
      mt[0] = mv.x; mt[1] = mv.y; mt[2] = mv.z; mt[3] = mv.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pf[i], mt[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* ob = out + (int64_t)b * n_stft * T;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = f0 + 4 * ty + i;
    if (f >= n_stft) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t t = t0 + 4 * tx + j;
      if (t < T) __stcs(ob + (int64_t)f * T + t, fmaxf(acc[i][j], 0.f));
    }
  }
}

static inline unsigned stream_grid(int64_t work_items) {
  const int64_t blocks = (work_items + 255) / 256;
  const int64_t cap = stream_grid_cap(8);   // 8 resident 256-thread CTAs per SM
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace edtts

using namespace edtts;

extern "C" int edtts_normalize_mel(const float* mel, float* mel_n_out, float* mean_out, float* std_out, int32_t B, int32_t T,
                                   int32_t n_mels, void* stream) {
  EDTTS_REQUIRE(mel && mean_out && std_out && B > 0 && T > 0 && n_mels > 0 && B <= 65535, EDTTS_EINVAL,
                "normalize_mel: B=%d T=%d n_mels=%d", B, T, n_mels);
  {
    LaunchScope ls(KC_MEL, as_stream(stream));
    mel_stats_kernel<<<dim3((n_mels + ST_BINS - 1) / ST_BINS, B), dim3(ST_BINS, ST_LANES), 0, as_stream(stream)>>>(
        mel, mean_out, std_out, T, n_mels);
  }
  if (mel_n_out) {
    const int64_t per_b = (int64_t)T * n_mels, total = per_b * B;
    LaunchScope ls(KC_MEL, as_stream(stream));
    mel_affine_kernel<false><<<stream_grid(total), 256, 0, as_stream(stream)>>>(mel, mean_out, std_out, mel_n_out, total, per_b,
                                                                               n_mels);
  }
  return check_launch("normalize_mel");
}

extern "C" int edtts_denormalize_mel(const float* mel_n, const float* mean, const float* std_, float* mel_out, int32_t B,
                                     int32_t T, int32_t n_mels, void* stream) {
  EDTTS_REQUIRE(mel_n && mean && std_ && mel_out && B > 0 && T > 0 && n_mels > 0, EDTTS_EINVAL,
                "denormalize_mel: B=%d T=%d n_mels=%d", B, T, n_mels);
  const int64_t per_b = (int64_t)T * n_mels, total = per_b * B;
  LaunchScope ls(KC_MEL, as_stream(stream));
  mel_affine_kernel<true><<<stream_grid(total), 256, 0, as_stream(stream)>>>(mel_n, mean, std_, mel_out, total, per_b, n_mels);
  return check_launch("denormalize_mel");
}

extern "C" int edtts_stitch_add(float* final_mel, float* final_weights, const float* x_chunk, const float* mean,
                                const float* std_, const float* window, int32_t B, int32_t T, int32_t n_mels,
                                int64_t total_frames, int64_t start_frame, void* stream) {
  EDTTS_REQUIRE(final_mel && final_weights && x_chunk && mean && std_ && window && B > 0 && B <= 65535 && T > 0 &&
                    n_mels > 0 && total_frames > 0 && start_frame >= 0,
                EDTTS_EINVAL, "stitch_add: B=%d T=%d n_mels=%d frames=%lld start=%lld", B, T, n_mels,
                (long long)total_frames, (long long)start_frame);
  // torch slicing clips at the end of the buffer and the += would then fail to broadcast: refuse it as the reference does
  EDTTS_REQUIRE(start_frame + T <= total_frames, EDTTS_EINVAL, "stitch_add: chunk [%lld, %lld) exceeds the %lld-frame buffer",
                (long long)start_frame, (long long)(start_frame + T), (long long)total_frames);
  LaunchScope ls(KC_MEL, as_stream(stream));
  stitch_add_kernel<<<dim3((T + 31) / 32, (n_mels + 31) / 32, B), 256, 0, as_stream(stream)>>>(
      x_chunk, mean, std_, window, final_mel, final_weights, T, n_mels, total_frames, start_frame);
  return check_launch("stitch_add");
}

extern "C" int edtts_stitch_finalize(const float* final_mel, const float* final_weights, float* mel_out, float* smooth_out,
                                     int32_t B, int32_t n_mels, int64_t buffer_frames, int64_t total_frames, int32_t kernel_h,
                                     int32_t kernel_w, void* stream) {
  EDTTS_REQUIRE(final_mel && final_weights && (mel_out || smooth_out) && B > 0 && n_mels > 0 && total_frames > 0 &&
                    total_frames <= buffer_frames && kernel_h >= 1 && kernel_w >= 1 && (kernel_h & 1) && (kernel_w & 1),
                EDTTS_EINVAL, "stitch_finalize: B=%d n_mels=%d frames=%lld/%lld kernel=%dx%d (odd sizes only)", B, n_mels,
                (long long)total_frames, (long long)buffer_frames, kernel_h, kernel_w);
  const int64_t n = (int64_t)B * n_mels * total_frames;
  LaunchScope ls(KC_MEL, as_stream(stream));
  stitch_finalize_kernel<<<stream_grid(n), 256, 0, as_stream(stream)>>>(final_mel, final_weights, mel_out, smooth_out, n_mels,
                                                                        buffer_frames, total_frames, kernel_h, kernel_w, n);
  return check_launch("stitch_finalize");
}

extern "C" int edtts_inverse_mel(const float* pinv_fb, const float* mel, float* spec_out, int32_t B, int32_t n_stft, int32_t n_mels,
                                 int64_t T, void* stream) {
  EDTTS_REQUIRE(pinv_fb && mel && spec_out && B > 0 && B <= 65535 && n_stft > 0 && n_mels > 0 && T > 0, EDTTS_EINVAL,
                "inverse_mel: B=%d n_stft=%d n_mels=%d T=%lld", B, n_stft, n_mels, (long long)T);
  LaunchScope ls(KC_MEL, as_stream(stream));
  inverse_mel_kernel<<<dim3((unsigned)((T + IM_TT - 1) / IM_TT), (n_stft + IM_TF - 1) / IM_TF, B), 256, 0, as_stream(stream)>>>(
      pinv_fb, mel, spec_out, n_stft, n_mels, T);
  return check_launch("inverse_mel");
}

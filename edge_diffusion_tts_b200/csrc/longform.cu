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

static inline unsigned stream_grid(int64_t work_items) {
  const int64_t blocks = (work_items + 255) / 256;
  const int64_t cap = 148 * 8;   // 8 resident 256-thread CTAs per SM
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

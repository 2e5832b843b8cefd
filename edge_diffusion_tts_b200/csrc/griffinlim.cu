// Griffin-Lim phase reconstruction as the reference applies it after the sampling path (generate_sample.py:135-141,
// inference_pipeline.py:89,398: torchaudio.transforms.GriffinLim(n_fft = 1024, n_iter = 32 / 100, win_length = 1024,
// hop_length = 160, power = 2) = torchaudio.functional.griffinlim, a third-party algorithm):
//
//   mag = spec^(1 / power);  angles = initial phases (complex);  tprev = 0
//   repeat n_iter times:  inverse = istft(mag * angles);  rebuilt = stft(inverse, center, reflect);
//                         a = rebuilt - momentum / (1 + momentum) * tprev;  angles = a / (|a| + 1e-16);  tprev = rebuilt
//   waveform = istft(mag * angles)
//
// Two kernels per iteration, one block per (utterance, frame), each with ONE n_fft-point complex FFT in shared memory
// (radix-2, twiddles from a shared table built with sincospif); the time signal between istft and stft is never written:
//   gl_synth_kernel   frame f: mag * angles (one-sided, n_fft / 2 + 1 bins) -> Hermitian spectrum -> inverse FFT -> real part *
//                     window / n_fft -> windowed frame xw[f][n_fft]
//   gl_analyse_kernel frame f: its n_fft samples of the overlap-added signal are GATHERED from the <= ceil(n_fft / hop)
//                     windowed frames that cover each sample, divided by the window envelope (sum of w^2 over the same
//                     frames, exactly torch.istft's normalisation) with the centre trim and the reflect padding of torch.stft
//                     folded into the index; * window -> FFT -> bins 0 .. n_fft / 2 -> the phase update above, in place
//   gl_wave_kernel    the same gather once more for the output samples.
// xw is (frames x n_fft x 4) bytes per utterance (3.3 MB at 800 frames: L2-resident between the two kernels).
#include "common.cuh"

namespace edtts {
namespace gl {

constexpr int THREADS = 256;

// in-place radix-2 decimation-in-time FFT of n = 1 << lg points in shared memory; sign = -1: forward, +1: inverse (unscaled).
// tw[j] = exp(-2 pi i j / n), j < n / 2.  The input must already be in bit-reversed order.
__device__ __forceinline__ void fft_inplace(float2* s, const float2* tw, int lg, float sign) {
  const int n = 1 << lg;
  for (int st = 0; st < lg; ++st) {
    const int half = 1 << st;
    for (int i = threadIdx.x; i < n / 2; i += THREADS) {
      const int j = i & (half - 1);
      const int a = ((i >> st) << (st + 1)) + j, b = a + half;
      float2 w = tw[j << (lg - 1 - st)];
      w.y *= -sign;                                       // table holds the forward twiddles (negative angle)
      const float2 u = s[a], v = s[b];
      const float2 t = make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
      s[a] = make_float2(u.x + t.x, u.y + t.y);
      s[b] = make_float2(u.x - t.x, u.y - t.y);
    }
    __syncthreads();
  }
}
__device__ __forceinline__ int bitrev(int x, int lg) { return (int)(__brev((unsigned)x) >> (32 - lg)); }
__device__ __forceinline__ void build_twiddles(float2* tw, int n) {
  for (int j = threadIdx.x; j < n / 2; j += THREADS) {
    float sn, cs;
    sincospif(-2.0f * (float)j / (float)n, &sn, &cs);
    tw[j] = make_float2(cs, sn);
  }
}

struct Args {
  const float* spec;     // [B][n_freq][F]  (power spectrogram, torchaudio layout)
  float2* angles;        // [B][F][n_freq]  current phases (in: the initial ones)
  float2* tprev;         // [B][F][n_freq]
  const float* window;   // [n_fft]  (win_length window zero-padded to n_fft, centred, as torch.stft does)
  float* xw;             // [B][F][n_fft]  windowed synthesis frames
  float* wave;           // [B][L]
  int F, n_fft, lg, hop, n_freq;
  int64_t L;             // hop * (F - 1)
  float inv_power, mom;  // 1 / power, momentum / (1 + momentum)
  int first;             // 1: tprev is still the scalar 0
};

__global__ void __launch_bounds__(THREADS) gl_synth_kernel(const Args a) {
  extern __shared__ float2 sm[];
  float2* s = sm;
  float2* tw = sm + a.n_fft;
  const int f = blockIdx.x, b = blockIdx.y, n = a.n_fft;
  build_twiddles(tw, n);
  const float2* ang = a.angles + ((int64_t)b * a.F + f) * a.n_freq;
  const float* sp = a.spec + (int64_t)b * a.n_freq * a.F + f;
  for (int k = threadIdx.x; k < n; k += THREADS) {
    const int kk = k <= n / 2 ? k : n - k;                // Hermitian extension of the one-sided spectrum
    const float p = sp[(int64_t)kk * a.F];
    const float mag = a.inv_power == 0.5f ? sqrtf(p) : powf(p, a.inv_power);
    float2 z = ang[kk];
    z = make_float2(mag * z.x, mag * z.y);
    if (k > n / 2) z.y = -z.y;
    if (k == 0 || k == n / 2) z.y = 0.f;                  // irfft ignores the imaginary part of the DC / Nyquist bins
    s[bitrev(k, a.lg)] = z;
  }
  __syncthreads();
  fft_inplace(s, tw, a.lg, +1.0f);
  float* out = a.xw + ((int64_t)b * a.F + f) * n;
  const float sc = 1.0f / (float)n;
  for (int i = threadIdx.x; i < n; i += THREADS) out[i] = s[i].x * sc * a.window[i];
}

// sample u of the UNTRIMMED overlap-added signal (length hop (F - 1) + n_fft) divided by the window envelope
__device__ __forceinline__ float ola_sample(const Args& a, const float* xwb, int64_t u) {
  const int n = a.n_fft;
  int64_t g1 = u / a.hop;
  if (g1 > a.F - 1) g1 = a.F - 1;
  int64_t g0 = (u - n + a.hop) / a.hop;                   // ceil((u - n + 1) / hop)
  if (u - n + 1 <= 0) g0 = 0;
  float acc = 0.f, env = 0.f;
  for (int64_t g = g0; g <= g1; ++g) {                    // ascending frame order: torch's fold accumulates the same way
    const int i = (int)(u - g * a.hop);
    const float w = a.window[i];
    acc += xwb[g * n + i];
    env = fmaf(w, w, env);
  }
  return acc / env;
}

__global__ void __launch_bounds__(THREADS) gl_analyse_kernel(const Args a) {
  extern __shared__ float2 sm[];
  float2* s = sm;
  float2* tw = sm + a.n_fft;
  const int f = blockIdx.x, b = blockIdx.y, n = a.n_fft;
  build_twiddles(tw, n);
  const float* xwb = a.xw + (int64_t)b * a.F * n;
  for (int i = threadIdx.x; i < n; i += THREADS) {
    int64_t r = (int64_t)f * a.hop - n / 2 + i;           // index into the trimmed signal [0, L), reflect-padded by n / 2
    if (r < 0) r = -r;
    if (r >= a.L) r = 2 * (a.L - 1) - r;
    const float y = ola_sample(a, xwb, r + n / 2);
    s[bitrev(i, a.lg)] = make_float2(y * a.window[i], 0.f);
  }
  __syncthreads();
  fft_inplace(s, tw, a.lg, -1.0f);
  float2* ang = a.angles + ((int64_t)b * a.F + f) * a.n_freq;
  float2* tp = a.tprev + ((int64_t)b * a.F + f) * a.n_freq;
  for (int k = threadIdx.x; k < a.n_freq; k += THREADS) {
    const float2 rb = s[k];
    float2 v = rb;
    if (!a.first && a.mom != 0.f) {
      const float2 p = tp[k];
      v = make_float2(rb.x - p.x * a.mom, rb.y - p.y * a.mom);
    }
    const float d = hypotf(v.x, v.y) + 1e-16f;
    ang[k] = make_float2(v.x / d, v.y / d);
    tp[k] = rb;
  }
}

__global__ void __launch_bounds__(THREADS) gl_wave_kernel(const Args a, int64_t out_len) {
  const int b = blockIdx.y;
  const float* xwb = a.xw + (int64_t)b * a.F * a.n_fft;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < out_len; i += (int64_t)gridDim.x * THREADS)
    a.wave[(int64_t)b * out_len + i] = i < a.L ? ola_sample(a, xwb, i + a.n_fft / 2) : 0.f;
}

}  // namespace gl
}  // namespace edtts

using namespace edtts;

extern "C" int64_t edtts_griffinlim_workspace_bytes(int32_t B, int32_t frames, int32_t n_fft) {
  const int64_t n_freq = n_fft / 2 + 1;
  return align_up((int64_t)B * frames * n_freq * 8, 256) * 2 + align_up((int64_t)B * frames * n_fft * 4, 256);
}

extern "C" int edtts_griffinlim(const float* spec, const float* angles_init, const float* window, float* wave_out, void* workspace,
                                int64_t workspace_bytes, int32_t B, int32_t frames, int32_t n_fft, int32_t hop, int32_t n_iter,
                                float power, float momentum, int64_t out_len, void* stream) {
  EDTTS_REQUIRE(spec && angles_init && window && wave_out && workspace && B > 0 && B <= 65535 && frames > 1 && hop > 0 && n_iter >= 0,
                EDTTS_EINVAL, "griffinlim: B=%d frames=%d hop=%d n_iter=%d", B, frames, hop, n_iter);
  int lg = 0;
  while ((1 << lg) < n_fft) ++lg;
  EDTTS_REQUIRE((1 << lg) == n_fft && n_fft >= 64 && n_fft <= 4096, EDTTS_ENOTSUP, "griffinlim: n_fft=%d (power of two, 64..4096)",
                n_fft);
  EDTTS_REQUIRE(hop <= n_fft && power > 0.f && momentum >= 0.f && momentum < 1.f, EDTTS_EINVAL, "griffinlim: hop=%d power=%g momentum=%g",
                hop, power, momentum);
  EDTTS_REQUIRE(workspace_bytes >= edtts_griffinlim_workspace_bytes(B, frames, n_fft), EDTTS_ENOSPC, "griffinlim: workspace");
  EDTTS_REQUIRE((int64_t)hop * (frames - 1) > n_fft / 2, EDTTS_EINVAL, "griffinlim: %d frames are too few for reflect padding", frames);
  cudaStream_t st = as_stream(stream);
  const int64_t n_freq = n_fft / 2 + 1;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int64_t abytes = align_up((int64_t)B * frames * n_freq * 8, 256);
  gl::Args a;
  a.spec = spec;
  a.angles = reinterpret_cast<float2*>(ws);
  a.tprev = reinterpret_cast<float2*>(ws + abytes);
  a.xw = reinterpret_cast<float*>(ws + 2 * abytes);
  a.window = window;
  a.wave = wave_out;
  a.F = frames; a.n_fft = n_fft; a.lg = lg; a.hop = hop; a.n_freq = (int)n_freq;
  a.L = (int64_t)hop * (frames - 1);
  a.inv_power = 1.0f / power;
  a.mom = momentum / (1.0f + momentum);
  if (cudaMemcpyAsync(a.angles, angles_init, (size_t)B * frames * n_freq * 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return check_launch("griffinlim copy");
  const int smem = (n_fft + n_fft / 2) * 8;
  static PerDeviceOnce configured;
  if (smem > 48 * 1024 && configured.need()) {
    if (cudaFuncSetAttribute(gl::gl_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 12) != cudaSuccess ||
        cudaFuncSetAttribute(gl::gl_analyse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 12) != cudaSuccess)
      return check_launch("griffinlim smem attribute");
    configured.set();
  }
  const dim3 grid(frames, B);
  for (int it = 0; it < n_iter; ++it) {
    a.first = it == 0;
    {
      LaunchScope ls(KC_MEL, st);
      gl::gl_synth_kernel<<<grid, gl::THREADS, smem, st>>>(a);
    }
    LaunchScope ls(KC_MEL, st);
    gl::gl_analyse_kernel<<<grid, gl::THREADS, smem, st>>>(a);
  }
  {
    LaunchScope ls(KC_MEL, st);
    gl::gl_synth_kernel<<<grid, gl::THREADS, smem, st>>>(a);
  }
  if (out_len <= 0) out_len = a.L;
  const int64_t blocks = (out_len + gl::THREADS - 1) / gl::THREADS;
  LaunchScope ls(KC_MEL, st);
  gl::gl_wave_kernel<<<dim3((unsigned)(blocks > 2048 ? 2048 : blocks), B), gl::THREADS, 0, st>>>(a, out_len);
  return check_launch("griffinlim");
}

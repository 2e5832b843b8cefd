/*
 * edtts.h -- C ABI of libedtts.so: the B200 (sm_100a) implementation of the
 * edge-diffusion-tts few-step sampling path.
 *
 * The reference (Krabbens/edge-diffusion-tts) has no FFI layer: its boundary for
 * this path is the Python class API.  Each entry point
 * below replaces the body of one reference method; the host-side mirror in
 * edge_diffusion_tts_b200/ binds them with ctypes (see INTEGRATION.md for the
 * stub a reference maintainer would add).  Reference citations are relative to
 * /root/reference/edge_diffusion_tts.
 *
 * Contract for every call:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless
 *     the name says host; the CALLER owns every buffer and the workspace;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *     allocation, no device synchronisation, no host read-back => every call is
 *     legal inside CUDA-graph capture;
 *   - returns 0 on success, a negative EDTTS_E* code otherwise; the message is
 *     available from edtts_last_error() (thread-local); C++ exceptions never
 *     cross the ABI;
 *   - all floating-point tensors are contiguous fp32, indices are int64
 *     (torch.long), exactly as the reference passes them.
 *
 * The kernels are specialised for the reference's default CFG (config.py:97-111):
 * n_mels 80, hidden 160, 4 layers, 4 heads (head_dim 40), ffn hidden 320,
 * kv_lora_rank 80, attention window 64, semantic_dim 128, codebook 512.
 */
#ifndef EDTTS_H
#define EDTTS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDTTS_N_MELS 80
#define EDTTS_HIDDEN 160
#define EDTTS_LAYERS 4
#define EDTTS_HEADS 4
#define EDTTS_HEAD_DIM 40
#define EDTTS_KV_RANK 80
#define EDTTS_FFN_HIDDEN 320
#define EDTTS_WINDOW 64
#define EDTTS_SEMANTIC_DIM 128
#define EDTTS_STEP_EMB_ROWS 16

#define EDTTS_OK 0
#define EDTTS_EINVAL (-1)    /* bad argument (null pointer, unsupported size)      */
#define EDTTS_ENOSPC (-2)    /* workspace too small                                */
#define EDTTS_ECUDA (-3)     /* a CUDA runtime call / launch failed                */
#define EDTTS_ENOTSUP (-4)   /* configuration outside the specialised kernel set   */

/* arithmetic of the decoder contractions */
#define EDTTS_PREC_FP32 0    /* CUDA-core FFMA, fp32 in / fp32 accumulate (parity path, 1e-4) */
#define EDTTS_PREC_BF16 1    /* tcgen05 tensor cores, bf16 in / fp32 accumulate (1e-2 rel-L2)  */
#define EDTTS_PREC_TF32X3 2  /* tcgen05 tensor cores, every operand split tf32 hi + lo, three MMAs per product,
                              * fp32 accumulate: fp32-grade (1e-4) results at tensor-core speed (decoder step only;
                              * edtts_context_prepare computes the once-per-utterance K | V with the FP32 kernels) */

/* epilogue fused into the last kernel of a decoder step */
#define EDTTS_STEP_EPS 0     /* write eps only           (decoder.py:109)          */
#define EDTTS_STEP_DDIM 1    /* eps -> x_prev, x0        (schedule.py:179-202)     */
#define EDTTS_STEP_DDPM 2    /* eps -> x_prev with noise (schedule.py:221-238)     */
#define EDTTS_STEP_DPM 3     /* v / x0 -> x0 (clamped), x_prev: DPM-Solver++ (schedule.py:326-438, 479-481) */

/* One DiffusionTransformerBlock (layers/transformer.py:71-160).  Names follow the
 * reference state-dict keys `layers.<i>.<...>`. */
typedef struct edtts_layer_weights {
  const float* norm1_norm_w;   /* norm1.norm.weight            [160]       */
  const float* norm1_proj_w;   /* norm1.proj.weight            [320,160]   */
  const float* norm1_proj_b;   /* norm1.proj.bias              [320]       */
  const float* attn_qkv_w;     /* attn.qkv.weight              [480,160]   */
  const float* attn_proj_w;    /* attn.proj.weight             [160,160]   */
  const float* attn_proj_b;    /* attn.proj.bias               [160]       */
  const float* norm2_w;        /* norm2.weight                 [160]       */
  const float* q_proj_w;       /* cross_attn.q_proj.weight     [160,160]   */
  const float* kv_down_w;      /* cross_attn.kv_down_proj.weight [80,160]  */
  const float* kv_norm_w;      /* cross_attn.kv_norm.weight    [80]        */
  const float* kv_up_w;        /* cross_attn.kv_up_proj.weight [320,80]    */
  const float* cross_out_w;    /* cross_attn.out_proj.weight   [160,160]   */
  const float* norm3_norm_w;   /* norm3.norm.weight            [160]       */
  const float* norm3_proj_w;   /* norm3.proj.weight            [320,160]   */
  const float* norm3_proj_b;   /* norm3.proj.bias              [320]       */
  const float* ffn0_w;         /* ffn.net.0.weight             [640,160]   */
  const float* ffn0_b;         /* ffn.net.0.bias               [640]       */
  const float* ffn3_w;         /* ffn.net.3.weight             [160,320]   */
  const float* ffn3_b;         /* ffn.net.3.bias               [160]       */
} edtts_layer_weights;

/* EdgeDiffusionDecoder parameters (models/decoder.py:17-64), device pointers. */
typedef struct edtts_decoder_weights {
  const float* token_emb;      /* token_emb.weight   [codebook,160]                 */
  const float* sem_proj_w;     /* sem_proj.weight    [160,128]                      */
  const float* sem_proj_b;     /* sem_proj.bias      [160]                          */
  const float* time1_w;        /* time_emb.1.weight  [160,160]                      */
  const float* time1_b;        /* time_emb.1.bias    [160]                          */
  const float* time3_w;        /* time_emb.3.weight  [160,160]                      */
  const float* time3_b;        /* time_emb.3.bias    [160]                          */
  const float* step_emb;       /* step_emb.weight    [16,160]                       */
  const float* in_proj_w;      /* in_proj.weight     [160,80]                       */
  const float* in_proj_b;      /* in_proj.bias       [160]                          */
  const float* pos_pe;         /* pos_emb.pe         [pos_rows,160]  (host may extend rows, F7) */
  const float* ctx_pe;         /* context_pos_emb.pe [ctx_rows,160]                 */
  const float* time_freqs;     /* exp(arange(80) * -ln(1e4)/79), embeddings.py:38-41 [80] */
  const float* final_norm_w;   /* final_norm.weight  [160]                          */
  const float* final_norm_b;   /* final_norm.bias    [160]                          */
  const float* out_proj_w;     /* out_proj.weight    [80,160]                       */
  const float* out_proj_b;     /* out_proj.bias      [80]                           */
  edtts_layer_weights layers[EDTTS_LAYERS];
  int32_t codebook_size;       /* rows of token_emb                                  */
  int32_t pos_rows;
  int32_t ctx_rows;
  int32_t reserved;
  const void* packed_bf16;     /* image written by edtts_pack_weights_bf16 (or NULL)  */
  const float* pos_pe_cm;      /* optional: pos_emb.pe chunk-major [40][pos_rows][4] (pe_cm[c][t][j] = pe[t][4c+j]); the fused bf16
                                * kernel then reads the table with coalesced loads.  NULL: it reads pos_pe row-major.         */
} edtts_decoder_weights;

/* --- library ------------------------------------------------------------- */
int edtts_version(void);
const char* edtts_last_error(void);
/* 1 if the device `stream` belongs to can run the kernels (sm_100), else 0. */
int edtts_device_supported(void);

/* --- measurement helpers (bench.py; never on the sampling path) ------------- */
/* Kernels are grouped in classes; every launch made through this library is counted. */
int edtts_kernel_classes(void);
const char* edtts_kernel_class_name(int cls);
int edtts_launch_counts(uint64_t* counts_out, int n);
/* When enabled, every launch outside stream capture is bracketed by CUDA events on its
 * stream; edtts_prof_collect synchronises, returns per-class elapsed ms / launches, clears. */
int edtts_prof_enable(int on);
int edtts_prof_collect(double* ms_out, uint64_t* n_out, int n);

/* --- VectorQuantizer (models/vq.py) --------------------------------------- */
/* vq.py:75-82 / :153-159: idx[r] = argmin_k ||z_r - E_k||^2, first minimum.
 * fp32 FFMA distances; near-ties are re-ranked in fp64 so the result equals the
 * exact argmin.  workspace: edtts_vq_workspace_bytes(K, D) bytes. */
int edtts_vq_argmin(const float* z, const float* codebook, int64_t* idx_out, int64_t rows, int32_t dim,
                    int32_t codebook_size, void* workspace, void* stream);
int64_t edtts_vq_workspace_bytes(int32_t codebook_size, int32_t dim);
/* The same search with the per-codebook work (code norms, tensor-core codebook image) done once: edtts_vq_pack fills
 * packed_out (edtts_vq_workspace_bytes(K, D) bytes); edtts_vq_argmin_packed is then ONE kernel launch per call. */
int edtts_vq_pack(const float* codebook, int32_t dim, int32_t codebook_size, void* packed_out, void* stream);
int edtts_vq_argmin_packed(const float* z, const float* codebook, const void* packed, int64_t* idx_out, int64_t rows, int32_t dim,
                           int32_t codebook_size, void* stream);
/* vq.py:83,98: z_q = z + (E[idx] - z)  (straight-through rounding included). */
int edtts_vq_gather_ste(const float* z, const float* codebook, const int64_t* idx, float* zq_out, int64_t rows,
                        int32_t dim, int32_t codebook_size, void* stream);
/* vq.py:102: counts = bincount(idx, minlength=K) as int32 (zeroed by the call). */
int edtts_vq_bincount(const int64_t* idx, int32_t* counts_out, int64_t rows, int32_t codebook_size, void* stream);

/* --- SemanticEncoder.proj (models/encoder.py:41-46) ------------------------ */
/* z = Linear(128,128)(LayerNorm(GELU(Linear(768,128)(h)))); workspace: edtts_encoder_proj_workspace_bytes(rows, in_dim)
 * (the intermediate rows and, for the tensor-core route, the tf32 hi | lo weight images packed per call). */
int64_t edtts_encoder_proj_workspace_bytes(int64_t rows, int32_t in_dim);
/* Weight images packed once per weight set (edtts_encoder_proj_image_bytes(in_dim) bytes), then two launches per call;
 * in_dim % 4 == 0; workspace: rows * 128 floats. */
int64_t edtts_encoder_proj_image_bytes(int32_t in_dim);
int edtts_encoder_proj_pack(const float* w0, const float* w3, int32_t in_dim, void* images_out, void* stream);
int edtts_encoder_proj_packed(const float* h, const void* images, const float* b0, const float* ln_w, const float* ln_b,
                              const float* b3, float* z_out, float* workspace, int64_t rows, int32_t in_dim, void* stream);
int edtts_encoder_proj(const float* h, const float* w0, const float* b0, const float* ln_w, const float* ln_b,
                       const float* w3, const float* b3, float* z_out, float* workspace, int64_t rows,
                       int32_t in_dim, void* stream);

/* --- EdgeDiffusionDecoder (models/decoder.py:66-109) ----------------------- */
/* decoder.py:77-80 + transformer.py:64-66 for all 8 AdaLayerNorms:
 *   cond[B,160]; mod[B, 2*LAYERS, 320] = (scale | shift) of norm1, norm3 per layer.
 * t int64[B]; step_idx int64[B] or NULL. */
int edtts_cond_prepare(const edtts_decoder_weights* w, const int64_t* t, const int64_t* step_idx, float* cond_out,
                       float* mod_out, int32_t B, void* stream);
/* decoder.py:83-93 + mla.py:144-153 for all layers (step-invariant):
 *   kv_out[LAYERS][B*S][320] = (k | v), head-major inside each half.
 * Exactly one of sem_idx (int64[B,S]) / sem_features (fp32[B,S,128]) is non-NULL.
 * kv_out holds edtts_context_kv_bytes(B, S, precision) bytes: for EDTTS_PREC_TF32X3 the rows are followed by the tf32 hi | lo
 * operand images of every layer's k | v, which the cross-attention of every edtts_decoder_step of that precision streams in;
 * the same buffer goes to edtts_decoder_step as `kv`. */
int edtts_context_prepare(const edtts_decoder_weights* w, const int64_t* sem_idx, const float* sem_features,
                          float* kv_out, void* workspace, int64_t workspace_bytes, int32_t B, int32_t S,
                          int32_t precision, void* stream);
int64_t edtts_context_workspace_bytes(int32_t B, int32_t S);
int64_t edtts_context_kv_bytes(int32_t B, int32_t S, int32_t precision);

/* Coefficients of the fused update, gathered on the device from the schedule
 * tables (so the call stays graph-capturable): t, t_prev int64[B]. */
typedef struct edtts_step_args {
  int32_t mode;                 /* EDTTS_STEP_*                                          */
  int32_t write_x_prev;         /* DDIM: 0 skips x_prev on the last step     */
  const int64_t* t;             /* [B]                                                   */
  const int64_t* t_prev;        /* [B] (DDIM)                                            */
  const float* alpha_bar;       /* schedule.alpha_bar          [T]                       */
  const float* alphas;          /* schedule.alphas             [T] (DDPM)                */
  const float* betas;           /* schedule.betas              [T] (DDPM)                */
  const float* posterior_var;   /* schedule.posterior_variance [T] (DDPM)                */
  const float* noise;           /* [B,T,80] N(0,1) draws (DDPM)                          */
  float* eps_out;               /* [B,T,80] or NULL                                      */
  float* x_prev_out;            /* [B,T,80] or NULL                                      */
  float* x0_out;                /* [B,T,80] or NULL (DDIM, DPM)                          */
  /* EDTTS_STEP_DPM (same meaning as the arguments of edtts_dpm_step) */
  const float* dpm_coef;        /* [B][8] per-row coefficients                           */
  const float* dpm_hist1;       /* [B,T,80] x0 history (order >= 2) or NULL              */
  const float* dpm_hist2;       /* [B,T,80] x0 history (order 3) or NULL                 */
  int32_t dpm_order;            /* 1, 2 or 3: the update rule used for this step         */
  int32_t dpm_predict_x0;       /* 0: the decoder output is v; 1: it is x0               */
} edtts_step_args;

/* One decoder evaluation (decoder.py:96-109) with the update rule fused into the
 * last kernel.  x_t [B,T,80]; mod from edtts_cond_prepare; kv from
 * edtts_context_prepare (S context tokens per utterance). */
int edtts_decoder_step(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv,
                       const edtts_step_args* args, void* workspace, int64_t workspace_bytes, int32_t B,
                       int32_t T, int32_t S, int32_t precision, void* stream);
int64_t edtts_decoder_workspace_bytes(int32_t B, int32_t T, int32_t S, int32_t precision);

/* bf16 tensor-core path: repack the fp32 parameters into the shared-memory tile
 * images the tcgen05 kernels bulk-copy (K-major core-matrix layout, bf16). */
int64_t edtts_packed_bf16_bytes(void);
int edtts_pack_weights_bf16(const edtts_decoder_weights* w, void* packed_out, void* stream);

/* --- DiffusionSchedule update rules, standalone (schedule.py) -------------- */
/* get_ddim_step (schedule.py:157-202): per-row t / t_prev gathered from alpha_bar.
 * n = elements per batch row (T*80).  noise may be NULL when eta == 0. */
int edtts_ddim_step(const float* x_t, const float* eps, const float* noise, const float* alpha_bar,
                    const int64_t* t, const int64_t* t_prev, float eta, float* x_prev_out, float* x0_out,
                    int32_t B, int64_t n, void* stream);
/* ddpm_step (schedule.py:204-238). */
int edtts_ddpm_step(const float* x_t, const float* eps, const float* noise, const float* alphas,
                    const float* alpha_bar, const float* betas, const float* posterior_var, const int64_t* t,
                    float* x_prev_out, int32_t B, int64_t n, void* stream);

/* DPMSolverPP first / second / third_order_update fused with model_to_x0 and the +-3 clamp (schedule.py:326-438,
 * 479-481).  coef is [B][8] fp32 per batch row: {sqrt_alpha_bar[t], sqrt_one_minus_alpha_bar[t], c0 = sigma_prev /
 * sigma_t, c1 = alpha_prev (1 - e^-h), c2 = alpha_prev ((1 - e^-h) / h + 1), 1 / r, c3 = alpha_prev ((1 - e^-h) / h^2 +
 * 0.5 / h + 0.5), unused}; the caller builds them with the reference's own tensor expressions so that the update is
 * bit-identical.  order_used 1: no history; 2: hist1 = x0_history[-1]; 3: hist1 = x0_history[-2] (older), hist2 =
 * x0_history[-1] (the reference's list order, schedule.py:510).  predict_x0 = 0: model_out is v (x0 = sa x - sb v,
 * clamped to +-3); 1: model_out is x0 (clamped); 2: model_out is an x0 that is used as given (the stand-alone
 * *_order_update methods).  Writes x_prev_out and, if non-NULL, x0_out; n = elements per batch row. */
int edtts_dpm_step(const float* x_t, const float* model_out, const float* hist1, const float* hist2, const float* coef,
                   int32_t order_used, int32_t predict_x0, float* x_prev_out, float* x0_out, int32_t B, int64_t n,
                   void* stream);

/* --- in-painting refine loop of the long-form pipeline (inference_pipeline.py:145-196) --------------------------------- */
/* One v-prediction DDIM step (inference_pipeline.py:181-192): v = v_uncond ? v_uncond + cfg_scale (v_cond - v_uncond) :
 * v_cond; x0 = clamp(sa x - sb v, +-3); eps = sb x + sa v; x_next = an x0 + bn eps.  coef is [B][4] fp32 per batch row:
 * {sa = sqrt_alpha_bar[t], sb = sqrt_one_minus_alpha_bar[t], an = sqrt(alpha_bar[t_next]), bn = sqrt(1 - alpha_bar[t_next])},
 * built by the caller with the reference's tensor expressions.  x_next_out may alias x_t; n = elements per batch row. */
int edtts_vddim_step(const float* x_t, const float* v_cond, const float* v_uncond, float cfg_scale, const float* coef,
                     float* x_next_out, float* x0_out, int32_t B, int64_t n, void* stream);
/* In-painting injection (inference_pipeline.py:170-174): x[b, :L, :] = sa known[b] + sb noise[b] for the first L of the T
 * frames (q_sample of the known frames at the current timestep); coef as above (sa, sb are read). */
int edtts_inpaint_inject(float* x, const float* known, const float* noise, const float* coef, int32_t B, int32_t T,
                         int32_t L, int32_t D, void* stream);

/* --- mel statistics and the cross-fade stitch of the long-form pipeline --------------------------- */
/* normalize_mel (edge_diffusion_tts/utils/audio.py:10-14): mel [B,T,n_mels] -> mean / std over the frames per (utterance,
 * mel bin) ([B,n_mels], i.e. the reference's [B,1,n_mels]; std unbiased, clamped to >= 1e-5) and, if mel_n_out is
 * non-NULL, (mel - mean) / std.  The statistics are accumulated in fp64 (parity with torch: <= 1e-6 relative); the
 * element-wise part is bit-exact given the statistics. */
int edtts_normalize_mel(const float* mel, float* mel_n_out, float* mean_out, float* std_out, int32_t B, int32_t T,
                        int32_t n_mels, void* stream);
/* denormalize_mel (utils/audio.py:17-19): mel_n * std + mean (two roundings, as torch) -- bit-exact. */
int edtts_denormalize_mel(const float* mel_n, const float* mean, const float* std_, float* mel_out, int32_t B, int32_t T,
                          int32_t n_mels, void* stream);
/* One chunk of the overlap-add (inference_pipeline.py:359-375), fused: final_mel[b, m, start + t] +=
 * exp(x_chunk[b, t, m] * std[b, m] + mean[b, m]) * window[t]; final_weights[start + t] += window[t].
 * final_mel [B, n_mels, total_frames], final_weights [total_frames], x_chunk [B, T, n_mels] (the refined, normalised
 * chunk), mean / std [B, n_mels] (per-chunk statistics), window [T].  start + T must fit the buffer. */
int edtts_stitch_add(float* final_mel, float* final_weights, const float* x_chunk, const float* mean, const float* std_,
                     const float* window, int32_t B, int32_t T, int32_t n_mels, int64_t total_frames, int64_t start_frame,
                     void* stream);
/* inference_pipeline.py:377-393: mel_out [B, n_mels, total_frames] = final_mel[..., :total_frames] /
 * clamp(final_weights, 1e-5); smooth_out = avg_pool2d(mel_out, (kernel_h, kernel_w), stride 1, padding k/2) (zero
 * padded, always divided by kernel_h * kernel_w).  Either output may be NULL; buffer_frames = row stride of final_mel. */
int edtts_stitch_finalize(const float* final_mel, const float* final_weights, float* mel_out, float* smooth_out, int32_t B,
                          int32_t n_mels, int64_t buffer_frames, int64_t total_frames, int32_t kernel_h, int32_t kernel_w,
                          void* stream);

/* InverseMelScale as the reference applies it (generate_sample.py:125-141, inference_pipeline.py:88,395; torchaudio
 * transforms.InverseMelScale = relu(lstsq(fb^T, mel)), i.e. the pseudo-inverse of the mel filter bank applied to every
 * frame): spec_out [B, n_stft, T] = relu(pinv_fb [n_stft, n_mels] @ mel [B, n_mels, T]).  pinv_fb is built once by the
 * caller (host, fp64 pinv of the filter bank). */
int edtts_inverse_mel(const float* pinv_fb, const float* mel, float* spec_out, int32_t B, int32_t n_stft, int32_t n_mels,
                      int64_t T, void* stream);

/* Griffin-Lim (torchaudio.transforms.GriffinLim at generate_sample.py:135-141, inference_pipeline.py:89,398; the algorithm is
 * torchaudio.functional.griffinlim: n_iter rounds of istft -> stft (centred, reflect padding) -> momentum phase update, then a
 * last istft).  spec [B, n_fft/2+1, frames] power spectrogram; angles_init [B, frames, n_fft/2+1][2] interleaved complex
 * initial phases (NOT normalised: torch.rand / ones, as the reference draws them); window [n_fft] (the win_length window
 * zero-padded, centred); wave_out [B, out_len], out_len = hop (frames - 1) (0 selects it; a longer out_len is zero-filled).
 * n_fft: a power of two in [64, 4096]. */
int edtts_griffinlim(const float* spec, const float* angles_init, const float* window, float* wave_out, void* workspace,
                     int64_t workspace_bytes, int32_t B, int32_t frames, int32_t n_fft, int32_t hop, int32_t n_iter, float power,
                     float momentum, int64_t out_len, void* stream);
int64_t edtts_griffinlim_workspace_bytes(int32_t B, int32_t frames, int32_t n_fft);

/* --- FSQ (models/fsq.py:18-132), the reference's alternative quantiser ------- */
/* forward (fsq.py:84-108): z [rows, dim] -> z_q = tanh(z) + (quantise(tanh(z)) - tanh(z)) and the flat index per row
 * (basis = cumprod([1] + levels[:-1]), first dimension fastest).  levels is a HOST array of dim (<= 8) ints.
 * codes_only != 0: codes_to_indices (fsq.py:110-119) of already quantised codes; zq_out unused. */
int edtts_fsq_forward(const float* z, const int32_t* levels_host, int32_t dim, int32_t codes_only, float* zq_out,
                      int64_t* idx_out, int64_t rows, void* stream);
/* indices_to_codes (fsq.py:121-132): last dimension fastest, as the reference decodes. */
int edtts_fsq_decode(const int64_t* idx, const int32_t* levels_host, int32_t dim, float* codes_out, int64_t rows,
                     void* stream);
/* FSQEncoder (models/fsq.py:135-222), the quantiser SemanticEncoder builds when cfg.use_fsq (models/encoder.py:49-50), fused:
 * forward (fsq.py:161-198, idx_in == NULL): z [rows, in_dim] -> proj_down (w_down [dim, in_dim], b_down [dim]) -> FSQ ->
 * proj_up (w_up [in_dim, dim], b_up [in_dim]) -> zq_out [rows, in_dim] (may be NULL: encode, fsq.py:212-216) and
 * idx_out [rows] (may be NULL); decode (fsq.py:218-221, idx_in != NULL): indices_to_codes -> proj_up -> zq_out. */
int edtts_fsq_encoder(const float* z, const int64_t* idx_in, const float* w_down, const float* b_down, const float* w_up,
                      const float* b_up, const int32_t* levels_host, int32_t dim, int32_t in_dim, float* zq_out,
                      int64_t* idx_out, int64_t rows, void* stream);

/* --- DepthwiseSeparableConv (layers/conv.py:10-64), operator level ---------- */
/* x [B,C_in,T] -> y [B,C_out,T_out], T_out = (T + 2*(k/2) - k)/stride + 1:
 * depthwise k taps (no bias) -> pointwise 1x1 (+bias) -> GroupNorm(min(8,C_out)) -> GELU.
 * workspace: edtts_dsconv_workspace_bytes(B, C_in, C_out, T_out). */
int edtts_dsconv_forward(const float* x, const float* dw_w, const float* pw_w, const float* pw_b,
                         const float* gn_w, const float* gn_b, float* y_out, void* workspace,
                         int64_t workspace_bytes, int32_t B, int32_t c_in, int32_t c_out, int32_t T,
                         int32_t kernel_size, int32_t stride, void* stream);
int64_t edtts_dsconv_workspace_bytes(int32_t B, int32_t c_in, int32_t c_out, int32_t t_out);

/* --- unit-test hooks for single kernels (parity tests call them through the ABI) */
/* y[rows,N] = x[rows,K] @ w[N,K]^T (+bias) in the given precision. */
int edtts_test_linear(const float* x, const float* w, const float* bias, float* y, int64_t rows, int32_t K,
                      int32_t N, int32_t precision, void* stream);
/* o[B,Tq,160] = attention(q,k,v) per head; window < 0 => full attention.
 * q rows have stride q_stride floats, k/v rows kv_stride floats. */
int edtts_test_attention(const float* q, int32_t q_stride, const float* k, const float* v, int32_t kv_stride,
                         float* o, int32_t B, int32_t Tq, int32_t Tk, int32_t window, int32_t precision,
                         void* stream);

/* y[rows,N] = epi(pro(x)[rows,K] @ w^T + bias) with the fused prologues / epilogues of the fp32 paths, on the CUDA cores
 * (use_tc = 0) or as a tf32 x 3 tensor-core GEMM (use_tc = 1; workspace: edtts_test_gemm_workspace_bytes).
 * pro: 0 none, 1 RMSNorm(norm_w), 2 AdaRMSNorm(norm_w; mod [rows / rows_per_batch][2K] = scale | shift), 3 LayerNorm(norm_w,
 * norm_b); epi: 0 store, 1 GELU, 2 + resid [rows,N], 3 + pe [(row % pe_period), N], 4 SwiGLU (w has 2N rows, bias 2N). */
int edtts_test_gemm(const float* x, const float* w, const float* bias, float* y, int64_t rows, int32_t K, int32_t N,
                    int32_t pro, int32_t epi, const float* norm_w, const float* norm_b, float norm_eps, const float* mod,
                    int32_t rows_per_batch, const float* resid, const float* pe, int32_t pe_period, int32_t use_tc,
                    void* workspace, int64_t workspace_bytes, void* stream);
int64_t edtts_test_gemm_workspace_bytes(int64_t rows, int32_t K, int32_t N, int32_t epi);

/* Residual stream h [B*T,160] of the bf16 path after in_proj and `n_layers` transformer blocks, the last block
 * optionally stopped early: stop_phase 1 = after x + attn(norm1(x)) (transformer.py:146), 2 = after the
 * cross-attention residual (:151), 0 = whole block (:158).  fused = 1 runs each block as the single fused
 * kernel, 0 as separate GEMM / attention launches.  mod, kv, workspace as for edtts_decoder_step (bf16). */
int edtts_test_hidden(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv,
                      float* h_out, void* workspace, int64_t workspace_bytes, int32_t B, int32_t T, int32_t S,
                      int32_t n_layers, int32_t stop_phase, int32_t fused, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EDTTS_H */

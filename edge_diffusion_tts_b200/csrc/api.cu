// Library-level entry points, error plumbing, encoder projection and the
// single-kernel test hooks of the C ABI.
#include <stdarg.h>
#include <string.h>
#include <vector>

#define EDTTS_DECL_ONLY
#include "common.cuh"
#include "gemm_simt.cuh"
#include "attention_simt.cuh"
#include "tc_path.cuh"
#include "tf32x3.cuh"
#include "t3_decoder.cuh"
#include <stdlib.h>

namespace edtts {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return EDTTS_ECUDA;
  }
  return EDTTS_OK;
}

int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int sm_count() {
  static int cached[64] = {};
  const int dev = current_device();
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  if (dev < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 1;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

// ---- launch counter + per-class event timing ---------------------------------
struct ProfSlot {
  cudaEvent_t a, b;
  int cls;
};
static unsigned long long g_launches[KC_COUNT] = {0};
static bool g_prof_on = false;
static std::vector<ProfSlot>* g_slots = nullptr;

LaunchScope::LaunchScope(int cls, cudaStream_t stream) : cls_(cls), stream_(stream), slot_(nullptr) {
  ++g_launches[cls];
  if (!g_prof_on) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
  if (!g_slots) g_slots = new std::vector<ProfSlot>();
  ProfSlot ps;
  ps.cls = cls;
  if (cudaEventCreate(&ps.a) != cudaSuccess || cudaEventCreate(&ps.b) != cudaSuccess) return;
  cudaEventRecord(ps.a, stream);
  g_slots->push_back(ps);
  slot_ = reinterpret_cast<void*>(g_slots->size());   // index + 1
}

LaunchScope::~LaunchScope() {
  if (slot_) cudaEventRecord((*g_slots)[reinterpret_cast<size_t>(slot_) - 1].b, stream_);
}

}  // namespace edtts

using namespace edtts;

extern "C" int edtts_kernel_classes(void) { return KC_COUNT; }

extern "C" const char* edtts_kernel_class_name(int cls) {
  static const char* names[KC_COUNT] = {"gemm_simt_fp32", "attn_window_simt_fp32", "attn_cross_simt_fp32", "cond",
                                        "embed_ctx", "vq", "schedule", "dsconv", "tc_gemm_bf16",
                                        "tc_attn_window_bf16", "tc_attn_cross_bf16", "tc_misc", "tc_layer_bf16", "mel_longform",
                                        "t3_gemm_tf32x3", "t3_attn_window_tf32x3", "t3_attn_cross_tf32x3"};
  return (cls >= 0 && cls < KC_COUNT) ? names[cls] : "?";
}

extern "C" int edtts_launch_counts(uint64_t* counts_out, int n) {
  EDTTS_REQUIRE(counts_out && n >= KC_COUNT, EDTTS_EINVAL, "launch_counts: need room for %d classes", KC_COUNT);
  for (int i = 0; i < KC_COUNT; ++i) counts_out[i] = g_launches[i];
  return EDTTS_OK;
}

extern "C" int edtts_prof_enable(int on) {
  g_prof_on = on != 0;
  return EDTTS_OK;
}

// Synchronises the device (host-side measurement helper, never called on the sampling path),
// accumulates per-class elapsed ms and launch counts of the recorded launches, then clears them.
extern "C" int edtts_prof_collect(double* ms_out, uint64_t* n_out, int n) {
  EDTTS_REQUIRE(ms_out && n_out && n >= KC_COUNT, EDTTS_EINVAL, "prof_collect: need room for %d classes", KC_COUNT);
  for (int i = 0; i < KC_COUNT; ++i) {
    ms_out[i] = 0.0;
    n_out[i] = 0;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return check_launch("prof_collect sync");
  if (!g_slots) return EDTTS_OK;
  for (auto& ps : *g_slots) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ps.a, ps.b) == cudaSuccess) {
      ms_out[ps.cls] += ms;
      n_out[ps.cls] += 1;
    }
    cudaEventDestroy(ps.a);
    cudaEventDestroy(ps.b);
  }
  g_slots->clear();
  return EDTTS_OK;
}

extern "C" int edtts_version(void) { return 100; }
extern "C" const char* edtts_last_error(void) { return g_err; }

extern "C" int edtts_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

// SemanticEncoder.proj (models/encoder.py:41-46): Linear(768,128) -> GELU -> LayerNorm(128) -> Linear(128,128).
// Tensor-core route (in_dim % 4 == 0): two launches of the tf32 x 3 GEMM kernel -- Linear + GELU + LayerNorm fused in the
// first one's epilogue, the second Linear on its output; the weight images are packed per call into the workspace.
// CUDA-core route otherwise (the LayerNorm is the prologue of the second FFMA GEMM).
static bool proj_tc(int in_dim) {
  static const bool off = getenv("EDTTS_PROJ_SIMT") != nullptr;       // development: force the CUDA-core route
  return !off && in_dim % 4 == 0;
}
extern "C" int64_t edtts_encoder_proj_workspace_bytes(int64_t rows, int32_t in_dim) {
  const int D = EDTTS_SEMANTIC_DIM;
  return align_up(rows * D * 4, 256) + align_up(t3::w_image_bytes(in_dim, D), 256) + align_up(t3::w_image_bytes(D, D), 256);
}
extern "C" int edtts_encoder_proj(const float* h, const float* w0, const float* b0, const float* ln_w,
                                  const float* ln_b, const float* w3, const float* b3, float* z_out, float* workspace,
                                  int64_t rows, int32_t in_dim, void* stream) {
  EDTTS_REQUIRE(h && w0 && b0 && ln_w && ln_b && w3 && b3 && z_out && workspace && rows > 0, EDTTS_EINVAL,
                "encoder_proj: null argument");
  const int D = EDTTS_SEMANTIC_DIM;
  if (proj_tc(in_dim)) {
    cudaStream_t st = as_stream(stream);
    char* ws = reinterpret_cast<char*>(workspace);
    float* y = reinterpret_cast<float*>(ws);
    float* img0 = reinterpret_cast<float*>(ws + align_up(rows * D * 4, 256));
    float* img3 = reinterpret_cast<float*>(ws + align_up(rows * D * 4, 256) + align_up(t3::w_image_bytes(in_dim, D), 256));
    int rc = t3::pack_w_tf32(w0, img0, in_dim, D, in_dim, st);
    if (!rc) rc = t3::pack_w_tf32(w3, img3, D, D, D, st);
    if (!rc) rc = t3::launch_t3_linear(h, rows, in_dim, in_dim, img0, b0, D, y, D, t3::EPI_T3_GELU_LN, ln_w, ln_b, 1e-5f, st);
    if (!rc) rc = t3::launch_t3_linear(y, rows, D, D, img3, b3, D, z_out, D, t3::EPI_T3_NONE, nullptr, nullptr, 0.f, st);
    return rc;
  }
  GemmArgs a;
  a.A = h; a.rows = rows; a.K = in_dim; a.lda = in_dim; a.W = w0; a.N = D; a.bias = b0; a.out = workspace; a.ldo = D;
  a.epi = EPI_GELU;
  int rc = launch_gemm_simt(a, as_stream(stream));
  if (rc) return rc;
  GemmArgs b;
  b.A = workspace; b.rows = rows; b.K = D; b.lda = D; b.W = w3; b.N = D; b.bias = b3; b.out = z_out; b.ldo = D;
  b.pro = PRO_LN; b.norm_w = ln_w; b.norm_b = ln_b; b.norm_eps = 1e-5f;
  return launch_gemm_simt(b, as_stream(stream));
}

// The two weight images of the tensor-core route (tf32 hi | lo, core-matrix layout), packed once per weight set:
// images_out holds edtts_encoder_proj_image_bytes(in_dim) bytes.
extern "C" int64_t edtts_encoder_proj_image_bytes(int32_t in_dim) {
  const int D = EDTTS_SEMANTIC_DIM;
  return align_up(t3::w_image_bytes(in_dim, D), 256) + align_up(t3::w_image_bytes(D, D), 256);
}
extern "C" int edtts_encoder_proj_pack(const float* w0, const float* w3, int32_t in_dim, void* images_out, void* stream) {
  EDTTS_REQUIRE(w0 && w3 && images_out && in_dim > 0, EDTTS_EINVAL, "encoder_proj_pack: null argument");
  const int D = EDTTS_SEMANTIC_DIM;
  char* img = reinterpret_cast<char*>(images_out);
  int rc = t3::pack_w_tf32(w0, reinterpret_cast<float*>(img), in_dim, D, in_dim, as_stream(stream));
  if (!rc) rc = t3::pack_w_tf32(w3, reinterpret_cast<float*>(img + align_up(t3::w_image_bytes(in_dim, D), 256)), D, D, D, as_stream(stream));
  return rc;
}
// edtts_encoder_proj with images packed earlier by edtts_encoder_proj_pack (tensor-core route only: in_dim % 4 == 0);
// workspace: rows * 128 floats.
extern "C" int edtts_encoder_proj_packed(const float* h, const void* images, const float* b0, const float* ln_w, const float* ln_b,
                                         const float* b3, float* z_out, float* workspace, int64_t rows, int32_t in_dim, void* stream) {
  EDTTS_REQUIRE(h && images && b0 && ln_w && ln_b && b3 && z_out && workspace && rows > 0, EDTTS_EINVAL, "encoder_proj_packed: null argument");
  EDTTS_REQUIRE(in_dim % 4 == 0, EDTTS_ENOTSUP, "encoder_proj_packed: in_dim=%d (multiple of 4)", in_dim);
  const int D = EDTTS_SEMANTIC_DIM;
  const char* img = reinterpret_cast<const char*>(images);
  const float* img0 = reinterpret_cast<const float*>(img);
  const float* img3 = reinterpret_cast<const float*>(img + align_up(t3::w_image_bytes(in_dim, D), 256));
  int rc = t3::launch_t3_linear(h, rows, in_dim, in_dim, img0, b0, D, workspace, D, t3::EPI_T3_GELU_LN, ln_w, ln_b, 1e-5f, as_stream(stream));
  if (!rc) rc = t3::launch_t3_linear(workspace, rows, D, D, img3, b3, D, z_out, D, t3::EPI_T3_NONE, nullptr, nullptr, 0.f, as_stream(stream));
  return rc;
}

extern "C" int edtts_test_linear(const float* x, const float* w, const float* bias, float* y, int64_t rows, int32_t K,
                                 int32_t N, int32_t precision, void* stream) {
  EDTTS_REQUIRE(x && w && y && rows > 0, EDTTS_EINVAL, "test_linear: null argument");
  if (precision >= EDTTS_PREC_BF16)   // 1: fp32-A prologue path, 2: chunk-major A, 3: chunk-major out (see tc_path.cu)
    return tc_test_linear(x, w, bias, y, rows, K, N, precision, as_stream(stream));
  GemmArgs g;
  g.A = x; g.rows = rows; g.K = K; g.lda = K; g.W = w; g.N = N; g.bias = bias; g.out = y; g.ldo = N;
  return launch_gemm_simt(g, as_stream(stream));
}

extern "C" int edtts_test_attention(const float* q, int32_t q_stride, const float* k, const float* v,
                                    int32_t kv_stride, float* o, int32_t B, int32_t Tq, int32_t Tk, int32_t window,
                                    int32_t precision, void* stream) {
  EDTTS_REQUIRE(q && k && v && o && B > 0 && Tq > 0 && Tk > 0, EDTTS_EINVAL, "test_attention: null argument");
  if (precision == EDTTS_PREC_BF16)
    return tc_test_attention(q, q_stride, k, v, kv_stride, o, B, Tq, Tk, window, as_stream(stream));
  AttnArgs a{q, q_stride, k, v, kv_stride, o, H, Tq, Tk, window, 1.0f / sqrtf((float)HD)};
  if (precision == EDTTS_PREC_TF32X3) {   // full-context attention: with the K | V operand images the decoder step uses (this unit-test
    void* img = nullptr;                  // hook owns the scratch: allocate, run, synchronise, free)
    if (window < 0 && cudaMalloc(&img, (size_t)t3::t3_kvimg_bytes(B, Tk)) != cudaSuccess) return check_launch("test_attention scratch");
    int rc = t3::launch_t3_attn(a, B, img, true, as_stream(stream));
    if (img) {
      cudaStreamSynchronize(as_stream(stream));
      cudaFree(img);
    }
    return rc;
  }
  return launch_attn_simt(a, B, as_stream(stream));
}

extern "C" int64_t edtts_test_gemm_workspace_bytes(int64_t rows, int32_t K, int32_t N, int32_t epi) {
  const bool swi = epi == EPI_SWIGLU;
  const int NB = (N % 160 == 0 || swi) ? 160 : 80;
  return align_up(t3::t3_gemm_image_floats(K, N, NB, swi) * 4, 256) + rows * 8;   // weight images | row statistics
}

extern "C" int edtts_test_gemm(const float* x, const float* w, const float* bias, float* y, int64_t rows, int32_t K, int32_t N,
                               int32_t pro, int32_t epi, const float* norm_w, const float* norm_b, float norm_eps, const float* mod,
                               int32_t rows_per_batch, const float* resid, const float* pe, int32_t pe_period, int32_t use_tc,
                               void* workspace, int64_t workspace_bytes, void* stream) {
  EDTTS_REQUIRE(x && w && y && rows > 0, EDTTS_EINVAL, "test_gemm: null argument");
  EDTTS_REQUIRE(pro >= PRO_NONE && pro <= PRO_LN && epi >= EPI_STORE && epi <= EPI_SWIGLU, EDTTS_EINVAL, "test_gemm: pro=%d epi=%d", pro, epi);
  GemmArgs g;
  g.A = x; g.rows = rows; g.K = K; g.lda = K; g.W = w; g.N = N; g.bias = bias; g.out = y; g.ldo = N; g.pro = pro; g.epi = epi;
  g.norm_w = norm_w; g.norm_b = norm_b; g.norm_eps = norm_eps; g.mod = mod; g.mod_stride = 2 * K;
  g.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : 1; g.resid = resid; g.pe = pe; g.pe_period = pe_period > 0 ? pe_period : 1;
  if (!use_tc) return launch_gemm_simt(g, as_stream(stream));
  const bool swi = epi == EPI_SWIGLU;
  EDTTS_REQUIRE(N % 80 == 0, EDTTS_ENOTSUP, "test_gemm: N=%d (multiple of 80)", N);
  const int NB = (N % 160 == 0 || swi) ? 160 : 80;
  EDTTS_REQUIRE(workspace && workspace_bytes >= edtts_test_gemm_workspace_bytes(rows, K, N, epi), EDTTS_ENOSPC, "test_gemm: workspace");
  if (epi == EPI_RESID) {   // the tensor-core GEMM accumulates the residual in place (resid == out), as the decoder step uses it
    EDTTS_REQUIRE(resid, EDTTS_EINVAL, "test_gemm: resid is null");
    if (cudaMemcpyAsync(y, resid, (size_t)rows * N * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)) != cudaSuccess)
      return check_launch("test_gemm residual copy");
    g.resid = y;
  }
  int rc = t3::pack_w_blocks(w, reinterpret_cast<float*>(workspace), K, N, NB, swi, as_stream(stream));
  if (rc) return rc;
  float* stats = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align_up(t3::t3_gemm_image_floats(K, N, NB, swi) * 4, 256));
  return t3::launch_t3_gemm(g, reinterpret_cast<const float*>(workspace), t3::t3_gemm_block_stride(K, NB), NB, stats, as_stream(stream));
}

extern "C" int edtts_test_hidden(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv,
                                 float* h_out, void* workspace, int64_t workspace_bytes, int32_t B, int32_t T, int32_t S,
                                 int32_t n_layers, int32_t stop_phase, int32_t fused, void* stream) {
  EDTTS_REQUIRE(w && x_t && mod && kv && h_out && workspace && B > 0 && T > 0 && S > 0, EDTTS_EINVAL,
                "test_hidden: null argument");
  EDTTS_REQUIRE(workspace_bytes >= edtts_decoder_workspace_bytes(B, T, S, EDTTS_PREC_BF16), EDTTS_ENOSPC,
                "test_hidden: workspace too small");
  return tc_test_hidden(w, x_t, mod, kv, h_out, workspace, B, T, S, n_layers, stop_phase, fused, as_stream(stream));
}

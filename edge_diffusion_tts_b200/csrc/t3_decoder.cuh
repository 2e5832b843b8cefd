// fp32-grade decoder step on the tensor cores (EDTTS_PREC_TF32X3): every Linear of the path as a tf32 x 3 GEMM with the
// normalisation prologues / epilogues of gemm_simt.cuh, both attentions as a tf32 x 3 streaming kernel (t3_decoder.cu).
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"        // GemmArgs; a translation unit other than decoder.cu defines EDTTS_DECL_ONLY first
#include "attention_simt.cuh"   // AttnArgs

namespace edtts {
namespace t3 {

// y = epi(pro(A) W^T + bias) with the semantics of launch_gemm_simt (g.W is only read by the packer): the weight matrix comes
// as operand images of NB-row blocks (pack_w_blocks), block j at wimg + j * img_stride floats.  EPI_SWIGLU: block j holds the x
// rows 80 j .. 80 j + 79 followed by their gate rows (NB = 160) and produces output columns 80 j .. 80 j + 79.
// stats: rows * 2 floats of scratch (normalisation prologues: the row statistics are computed by a small launch in front)
int launch_t3_gemm(const GemmArgs& g, const float* wimg, int64_t img_stride, int NB, float* stats, cudaStream_t st);
// weight images of W [N or 2N][K] for launch_t3_gemm into img (t3_gemm_image_floats(K, N, swiglu) floats), one launch
int pack_w_blocks(const float* W, float* img, int K, int N, int NB, bool swiglu, cudaStream_t st);
int64_t t3_gemm_image_floats(int K, int N, int NB, bool swiglu);
int64_t t3_gemm_block_stride(int K, int NB);

// attention with the semantics of launch_attn_simt; kvimg (full-context attention only): t3_kvimg_bytes(B, Tk) bytes of scratch for
// the K | V operand images the issuer then streams in by bulk copies, or null
int64_t t3_kvimg_bytes(int B, int Tk);
// build = true: the images are built from a.k / a.v in front of the attention; false: kvimg already holds them (launch_t3_kvimg)
int launch_t3_attn(const AttnArgs& a, int B, void* kvimg, bool build, cudaStream_t st);
int launch_t3_kvimg(const float* k, const float* v, int kv_stride, int Tk, int B, void* kvimg, cudaStream_t st);
// layout of the kv buffer of EDTTS_PREC_TF32X3: [NL][B S][320] fp32 rows (k | v), then per layer the operand images
int64_t t3_kv_rows_bytes(int B, int S);              // offset of the first image
int64_t t3_kv_total_bytes(int B, int S);

// context K | V of all layers (mla.py:144-153): kv_down -> kv_norm -> kv_up as tf32 x 3 GEMMs; ctx [rows,160] in, craw [rows,80]
// scratch, kv_out [4][rows][320]; scratch: t3_context_scratch_bytes(rows) bytes (weight images + row statistics)
int64_t t3_context_scratch_bytes(int64_t rows);
int t3_context_kv(const edtts_decoder_weights* w, const float* ctx, float* craw, float* kv_out, void* scratch, int64_t rows, int B, int S,
                  cudaStream_t st);

// one decoder evaluation (decoder.cu: decoder_step_fp32 with the kernels above); workspace as sized below
int64_t t3_decoder_workspace_bytes(int B, int T, int S);
int t3_decoder_step(const edtts_decoder_weights* w, const float* x_t, const float* mod, const float* kv, const edtts_step_args* args,
                    void* workspace, int B, int T, int S, cudaStream_t st);

}  // namespace t3
}  // namespace edtts

// fp32-grade GEMMs on the tensor cores: every operand is split into a tf32 head and a tf32 tail (x = hi + lo, 22 mantissa
// bits) and a k-step issues three kind::tf32 MMAs (hi hi, lo hi, hi lo) into an fp32 accumulator in tensor memory.  The
// dropped lo lo term and the tail's own truncation are below 2^-20 of |x||w| per product, i.e. the result is as good as an
// fp32 FFMA chain -- which is what the operators using it (DepthwiseSeparableConv's pointwise product, SemanticEncoder.proj,
// the VQ distance product) are tested against (max-abs 2e-5 / bit-exact indices).
//
// Operand images (shared memory, the UMMA canonical K-major no-swizzle layout with 32-bit elements): slab s holds elements
// k = 4 s .. 4 s + 3 of every row as one 16-byte unit, [slab][row][4]; a k-step of 8 elements is two slabs.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace edtts {
namespace t3 {
using namespace tc;

constexpr int TM = 128;                              // rows per tile (MMA M)
constexpr int KC = 32;                               // contraction elements per chunk
constexpr int A_HALF = (KC / 4) * TM * 16;           // bytes of the A operand's hi (or lo) half of one chunk: 8 slabs x 128 rows x 16 B

// Instruction descriptor, kind::tf32: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), both K-major, N >> 3, M >> 4.
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M_, uint32_t N_) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N_ >> 3) << 17) | ((M_ >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// 32 lanes x 4 consecutive columns, NOT waited for
__device__ __forceinline__ void tmem_ld4_nw(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}

// 4 values of one row -> the 16-byte units of slab `slab` in the hi and the lo image
__device__ __forceinline__ void split_store(uint8_t* sAh, uint8_t* sAl, int slab, int r, const float* v) {
  float4 hi, lo;
  hi.x = tf32_rna(v[0]); hi.y = tf32_rna(v[1]); hi.z = tf32_rna(v[2]); hi.w = tf32_rna(v[3]);
  lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
  *reinterpret_cast<float4*>(sAh + slab * (TM * 16) + r * 16) = hi;
  *reinterpret_cast<float4*>(sAl + slab * (TM * 16) + r * 16) = lo;
}

// the same with the tail rounded to tf32 as well (cvt.rna): the MMA TRUNCATES the low 13 bits of a 32-bit operand, which on the
// exact tail x - hi is an error of up to 2^-21 |x|, always toward zero; rounded to nearest it is 2^-22 |x| and unbiased
__device__ __forceinline__ void split_store_r(uint8_t* sAh, uint8_t* sAl, int slab, int rows, int r, const float* v) {
  float4 hi, lo;
  hi.x = tf32_rna(v[0]); hi.y = tf32_rna(v[1]); hi.z = tf32_rna(v[2]); hi.w = tf32_rna(v[3]);
  lo.x = tf32_rna(v[0] - hi.x); lo.y = tf32_rna(v[1] - hi.y); lo.z = tf32_rna(v[2] - hi.z); lo.w = tf32_rna(v[3] - hi.w);
  *reinterpret_cast<float4*>(sAh + slab * (rows * 16) + r * 16) = hi;
  *reinterpret_cast<float4*>(sAl + slab * (rows * 16) + r * 16) = lo;
}

// the three MMAs of every 8-element k-step of one chunk: D[128 x n] (+)= A[128 x 8 nks] W[n x 8 nks]^T
__device__ __forceinline__ void issue_chunk(uint32_t d_tmem, uint32_t ah, uint32_t al, uint32_t wh, uint32_t wl, int nks, int n,
                                            bool accumulate) {
  const uint32_t idesc = idesc_tf32(TM, (uint32_t)n);
  for (int ks = 0; ks < nks; ++ks) {
    const uint32_t ao = ks * 2 * (TM * 16), wo = ks * 2 * (n * 16);
    const uint64_t dah = make_desc(ah + ao, TM * 16, 128), dal = make_desc(al + ao, TM * 16, 128);
    const uint64_t dwh = make_desc(wh + wo, n * 16, 128), dwl = make_desc(wl + wo, n * 16, 128);
    umma_tf32(d_tmem, dah, dwh, idesc, accumulate || ks > 0);
    umma_tf32(d_tmem, dal, dwh, idesc, true);
    umma_tf32(d_tmem, dah, dwl, idesc, true);
  }
}

// Weight image of an nn.Linear-style matrix W [n][k] (row stride ldw): per 32-element chunk c: [hi | lo][slab s < 8][row < n][4]
// = element (row, 32 c + 4 s + j); elements beyond k are zero.  Bytes per chunk: 2 * 8 * n * 16.
__global__ void pack_w_tf32_kernel(const float* __restrict__ W, float* __restrict__ img, int k, int n, int ldw, int nchunk);
int pack_w_tf32(const float* W, float* img, int k, int n, int ldw, cudaStream_t st);          // launches the kernel above
enum { EPI_T3_NONE = 0, EPI_T3_GELU = 1, EPI_T3_GELU_LN = 2 };
// out[rows, N] = epi(A[rows, K] W^T + bias); N % 16 == 0, N <= 256 (LayerNorm epilogue: N <= 128); K, lda, ldo multiples of 4
int launch_t3_linear(const float* A, int64_t rows, int K, int lda, const float* wimg, const float* bias, int N, float* out, int ldo,
                     int epi, const float* ln_w, const float* ln_b, float ln_eps, cudaStream_t st);
// VQ nearest codeword on the tensor cores (D % 8 == 0, K <= 512); wimg: t3_vq_image_bytes(D) bytes of scratch for the codebook image
bool t3_vq_ok(int D, int K);
int64_t t3_vq_image_bytes(int D);
int launch_t3_vq(const float* z, const float* E, const float* ee, float* wimg, int64_t* idx, int64_t rows, int D, int K, cudaStream_t st);
int pack_codebook(const float* E, float* wimg, int D, int K, cudaStream_t st);
int launch_t3_vq_packed(const float* z, const float* E, const float* ee, const float* wimg, int64_t* idx, int64_t rows, int D, int K,
                        cudaStream_t st);
static inline int64_t w_image_bytes(int k, int n) { return (int64_t)((k + KC - 1) / KC) * 2 * 8 * n * 16; }

}  // namespace t3
}  // namespace edtts

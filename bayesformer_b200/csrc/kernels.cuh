// Internal launch wrappers (transformer_layers.cu) used by the host driver (transformer.cu).
#pragma once
#include "common.cuh"

namespace bfb {

// Per-row (sum, sum of squares) accumulated by several GEMM epilogue threads.  Fixed-point 64-bit integers, not
// floats: integer atomic adds commute, so the statistics — and every score downstream — are bit-identical whatever
// order the column tiles finish in (chunking, sharding, run to run).  Sum in units of 2^-24, squares in units of 2^-16:
// exact to ~1e-7 relative at LayerNorm-input magnitudes, no overflow up to 512 x 65504^2.
struct RowStats { long long sum, sq; };
constexpr float kRowStatSumScale = 16777216.0f, kRowStatSqScale = 65536.0f;
__device__ __forceinline__ void row_stats_add(RowStats* s, float sum, float sq) {
  atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum), (unsigned long long)__float2ll_rn(sum * kRowStatSumScale));
  atomicAdd(reinterpret_cast<unsigned long long*>(&s->sq), (unsigned long long)__float2ll_rn(sq * kRowStatSqScale));
}
__device__ __forceinline__ float2 row_stats_load(const RowStats* s) {
  const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(s));
  return make_float2(__ll2float_rn(v.x) * (1.0f / kRowStatSumScale), __ll2float_rn(v.y) * (1.0f / kRowStatSqScale));
}

int launch_init_tokens(const int64_t* prompts, int64_t n_items_total, int64_t item0, int64_t n_seq,
                       int n_draws, int prompt_len, int stride, int vocab, int32_t* tokens,
                       cudaStream_t st);
int launch_embed(const int32_t* tokens, int tok_stride, const float* emb, const float* pe, int64_t n_seq,
                 int len, int d, const MaskSource& ms, int site, float* x, cudaStream_t st, void* x16 = nullptr,
                 uint32_t fmt = 0);
int launch_sgemm(const float* A, const float* W, const float* bias, float* C, int64_t M, int N, int K,
                 int64_t a_row_mul, int64_t a_row_off, int relu, cudaStream_t st);
// Q, K, V projections with Head.bayes input dropout (transformer.py:50-53): projection c of head h multiplies its own
// masked copy of x, site = site_base + 3 * h + c
int launch_qkv_head_in(const float* x, const float* w_qkv, float* qkv, int64_t n_seq, int len, int d, int n_heads,
                       const MaskSource& ms, int site_base, cudaStream_t st);
int launch_attention(const float* qkv, float* out, int64_t n_seq, int len, int d, int n_heads,
                     const MaskSource& ms, int site, cudaStream_t st);
int launch_resid_ln(const float* x, const float* y, float* out, int64_t n_rows, int len, int d,
                    const float* gamma, const float* beta, float eps, const MaskSource& ms,
                    int branch_site, int out_site, cudaStream_t st, void* out16 = nullptr, uint32_t fmt = 0,
                    const void* y16 = nullptr);
int launch_ln16_apply(const void* v16, const RowStats* stats, void* out16, int64_t n_rows, int len, int d, const float* gamma,
                      const float* beta, float eps, const MaskSource& ms, int out_site, uint32_t fmt, cudaStream_t st);
int launch_score_step(const float* logits, int64_t n_seq, int vocab, int scheme, int mode,
                      float temperature, const MaskSource& ms, const int32_t* forced, int32_t* tokens,
                      int tok_stride, int write_pos, float* seq_scores, float* logp_out, cudaStream_t st,
                      int pitch = 0);
int launch_logsoftmax_seqfirst(const float* logits, int64_t n_seq, int len, int vocab, float* out,
                               cudaStream_t st, int pitch = 0);
int launch_to16(const float* src, void* dst, int64_t n, uint32_t fmt, cudaStream_t st);
// LayerNorm folded into the product that consumes it: w16g = W * diag(gamma) (16-bit), col_s[n] = sum_k w16g[n, k],
// col_c[n] = sum_k W[n, k] beta[k] + b[n]; W (L, N, K), b (L, N), gamma / beta (L, K)
int launch_fold_ln_pack(const float* W, const float* b, const float* gamma, const float* beta, int L, int N, int K,
                        uint32_t fmt, void* w16g, float* col_s, float* col_c, cudaStream_t st);
int launch_reduce_draws(const float* seq_scores, int64_t n_items_chunk, int n_draws, float* dst,
                        cudaStream_t st);
int launch_step_mean(const float* step_scores, int64_t n_items, int n_steps, float* pool_scores,
                     cudaStream_t st);


// fused MC-forward kernel (mc_fused.cu)
bool fused_supported(const bfb200_transformer& m);
size_t fused_image_bytes(const bfb200_transformer& m);
int fused_pack(const bfb200_transformer& m, void* image, size_t bytes, cudaStream_t st);
int launch_mc_forward_fused(const bfb200_transformer& m, const int32_t* tokens, int32_t* tokens_out, int tok_stride,
                            int64_t n_seq, int len, int scheme, int mode, float temperature, const MaskSource& ms,
                            const int32_t* forced, float* seq_scores, float* logp_out, cudaStream_t st);


// TMA-fed tcgen05 GEMM (gemm_tc.cu): C = A W^T (+ bias)(ReLU), 16-bit operands (fmt: umma::kFmtF16 / kFmtBF16)
int launch_gemm_tc(const void* A, int64_t a_pitch, const void* W, const float* bias, void* out16, float* out32,
                   int64_t ldc, int64_t M, int N, int K, int relu, uint32_t fmt, cudaStream_t st);
// Fused epilogues of the same kernel:
//   resid16 / stats_out : out16 = acc (+ bias) + R16, and (sum, sum of squares) of every fp32 result row accumulated
//                         atomically into stats_out (M RowStats, zeroed by the caller) — residual add + LayerNorm statistics;
//   stats_in / col_s / col_c : out = rstd_r * (acc - mean_r * col_s[n]) + col_c[n] (ReLU) — a LayerNorm folded into the
//                         product that consumes it ((mean, rstd) of row r from stats_in over stats_width values).
struct GemmFusion {
  const void* resid16 = nullptr;
  int64_t resid_pitch = 0;
  RowStats* stats_out = nullptr;
  const RowStats* stats_in = nullptr;
  const float* col_s = nullptr;
  const float* col_c = nullptr;
  int stats_width = 0;
  float stats_eps = 0.f;
};
int launch_gemm_tc_ex(const void* A, int64_t a_pitch, const void* W, const float* bias, void* out16, float* out32,
                      int64_t ldc, int64_t M, int N, int K, int relu, uint32_t fmt, const GemmFusion& fu, cudaStream_t st);


// tcgen05 causal attention for heads of width 64 (attention_tc.cu): qkv16 (R, 3d) -> out16 (R, d)
int launch_attention_tc(const void* qkv16, void* out16, int64_t n_seq, int len, int d, int n_heads, uint32_t fmt,
                        const MaskSource& ms, int site, cudaStream_t st);

}  // namespace bfb

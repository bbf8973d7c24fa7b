// Host-side driver of the transformer path: workspace carving, one forward pass over a chunk
// of sequences, the decode/scoring loop of *_uncertainty, and the C ABI around them.
// Reference control flow: src/libs/active_learning.py:38-155 and :209; src/model/transformer.py:349-375.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace bfb {

struct Workspace {
  int32_t* tokens;  // (S, max_len)
  float* x;         // (R, d) block input / output
  float* qkv;       // (R, 3d)
  float* att;       // (R, d) concatenated heads
  float* y;         // (R, d) branch output (W_out / FFN)
  float* x1;        // (R, d) LayerNorm1 output
  float* h;         // (R, F)
  float* logits;    // (S or R, V)
  float* seq_scores;  // (S)
  // layered tensor-core path: 16-bit operand copies and a second residual buffer
  uint16_t* x16;    // (R, d)
  uint16_t* qkv16;  // (R, 3d)
  uint16_t* att16;  // (R, d) attention output, then LayerNorm1 output
  uint16_t* h16;    // (R, F)
  float* x2;        // (R, d) next block input
  RowStats* stats;  // (R) per-row (sum, sum of squares) of a LayerNorm's input, fixed point (kernels.cuh)
};

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// bytes for n_seq sequences of at most max_len tokens; full_logits: logits for every row
// ---- which path serves a model ---------------------------------------------------------------
//   fp32            : layered CUDA-core kernels, any shape
//   f16, fused class: fused MC-forward kernel (mc_fused.cu)
//   f16 / bf16 else : layered tensor-core path — TMA-fed tcgen05 GEMMs (gemm_tc.cu), tcgen05 attention
//                     (attention_tc.cu), fp32 residual stream / LayerNorm; heads of width 64, len <= 256
static bool use_fused(const bfb200_transformer& m) { return m.compute == BFB200_COMPUTE_F16 && fused_supported(m); }
// the per-head input-dropout site (Head.bayes) is served by the fp32 kernels only
static bool tc_shape_ok(const bfb200_transformer& m) {
  return !(m.site_flags & BFB200_SITE_HEAD_IN) && m.n_layers >= 1 && m.d_model == m.n_heads * 64 && (m.d_ff & 7) == 0 && m.d_ff >= 8;
}
static bool use_tc(const bfb200_transformer& m) {
  return (m.compute == BFB200_COMPUTE_F16 && !fused_supported(m)) || m.compute == BFB200_COMPUTE_BF16;
}
static uint32_t tc_fmt(const bfb200_transformer& m) {
  return m.compute == BFB200_COMPUTE_BF16 ? umma::kFmtBF16 : umma::kFmtF16;
}
static int logits_pitch(const bfb200_transformer& m) { return use_tc(m) ? (m.ntoken + 3) & ~3 : m.ntoken; }

// 16-bit weight image of the layered tensor-core path: [W_qkv | W_out | FFN W1 | FFN W2 | head | W1 * gamma1 |
// col_s (fp32) | col_c (fp32)], offsets in 16-bit elements.  The last three serve LayerNorm1 folded into the first
// FFN product (forward_chunk_tc): W1g = W1 * diag(gamma1), col_s[n] = sum_k W1g[n, k] (of the ROUNDED 16-bit values
// the tensor core multiplies), col_c = W1 beta1 + b1.
struct TcImage {
  size_t qkv, out, ff1, ff2, head, ff1g, col_s, col_c, elems;
};
static TcImage tc_image(const bfb200_transformer& m) {
  const size_t L = m.n_layers, d = m.d_model, F = m.d_ff, V = m.ntoken;
  TcImage t;
  t.qkv = 0;
  t.out = t.qkv + L * 3 * d * d;
  t.ff1 = t.out + L * d * d;
  t.ff2 = t.ff1 + L * F * d;
  t.head = t.ff2 + L * d * F;
  t.ff1g = (t.head + V * d + 7) & ~(size_t)7;
  t.col_s = t.ff1g + L * F * d;
  t.col_c = t.col_s + 2 * L * F;
  t.elems = t.col_c + 2 * L * F;
  return t;
}

static size_t workspace_bytes(const bfb200_transformer& m, int64_t n_seq, int max_len, bool full_logits) {
  if (use_fused(m) && !full_logits)  // fused path: activations never leave the chip
    return 256 + align256((size_t)n_seq * max_len * sizeof(int32_t)) + align256((size_t)n_seq * sizeof(float));
  const size_t R = (size_t)n_seq * max_len;
  const size_t d = m.d_model, F = m.d_ff, V = m.ntoken;
  if (use_tc(m)) {
    const size_t Vp = (size_t)logits_pitch(m);
    return 256 + align256((size_t)n_seq * max_len * sizeof(int32_t)) + 3 * align256(R * d * 4) +
           2 * align256(R * d * 2) + align256(R * 3 * d * 2) + align256(R * F * 2) + align256(R * sizeof(RowStats)) +
           align256((full_logits ? R : (size_t)n_seq) * Vp * 4) + align256((size_t)n_seq * 4);
  }
  size_t b = 256;
  b += align256((size_t)n_seq * max_len * sizeof(int32_t));
  b += 4 * align256(R * d * sizeof(float));
  b += align256(R * 3 * d * sizeof(float));
  b += align256(R * F * sizeof(float));
  b += align256((full_logits ? R : (size_t)n_seq) * V * sizeof(float));
  b += align256((size_t)n_seq * sizeof(float));
  return b;
}

static Workspace carve(void* base, const bfb200_transformer& m, int64_t n_seq, int max_len, bool full_logits) {
  const size_t R = (size_t)n_seq * max_len;
  const size_t d = m.d_model, F = m.d_ff, V = m.ntoken;
  char* p = (char*)(((uintptr_t)base + 255) & ~(uintptr_t)255);
  Workspace w;
  memset(&w, 0, sizeof(w));
  auto take = [&](size_t bytes) { char* q = p; p += align256(bytes); return q; };
  w.tokens = (int32_t*)take((size_t)n_seq * max_len * sizeof(int32_t));
  if (use_fused(m) && !full_logits) {
    w.seq_scores = (float*)take((size_t)n_seq * sizeof(float));
    w.x = w.att = w.y = w.x1 = w.qkv = w.h = w.logits = nullptr;
    return w;
  }
  if (use_tc(m)) {
    const size_t Vp = (size_t)logits_pitch(m);
    w.x = (float*)take(R * d * 4);
    w.x2 = (float*)take(R * d * 4);
    w.y = (float*)take(R * d * 4);
    w.x16 = (uint16_t*)take(R * d * 2);
    w.att16 = (uint16_t*)take(R * d * 2);
    w.qkv16 = (uint16_t*)take(R * 3 * d * 2);
    w.h16 = (uint16_t*)take(R * F * 2);
    w.stats = (RowStats*)take(R * sizeof(RowStats));
    w.logits = (float*)take((full_logits ? R : (size_t)n_seq) * Vp * 4);
    w.seq_scores = (float*)take((size_t)n_seq * 4);
    w.att = w.x1 = w.qkv = w.h = nullptr;
    return w;
  }
  w.x = (float*)take(R * d * sizeof(float));
  w.att = (float*)take(R * d * sizeof(float));
  w.y = (float*)take(R * d * sizeof(float));
  w.x1 = (float*)take(R * d * sizeof(float));
  w.qkv = (float*)take(R * 3 * d * sizeof(float));
  w.h = (float*)take(R * F * sizeof(float));
  w.logits = (float*)take((full_logits ? R : (size_t)n_seq) * V * sizeof(float));
  w.seq_scores = (float*)take((size_t)n_seq * sizeof(float));
  return w;
}

static int validate_model(const bfb200_transformer* m) {
  BFB_REQUIRE(m != nullptr, BFB200_ERR_INVALID_ARG, "model is NULL");
  BFB_REQUIRE(m->ntoken >= 1 && m->d_model >= 1 && m->n_heads >= 1 && m->d_ff >= 1 && m->n_layers >= 0,
              BFB200_ERR_INVALID_ARG, "bad model dimensions");
  BFB_REQUIRE(m->d_model % m->n_heads == 0, BFB200_ERR_INVALID_ARG, "d_model must be divisible by n_heads");
  BFB_REQUIRE((m->d_model & 3) == 0 && (m->d_ff & 3) == 0, BFB200_ERR_UNSUPPORTED,
              "d_model (%d) and dim_feedforward (%d) must be multiples of 4", m->d_model, m->d_ff);
  BFB_REQUIRE(m->compute == BFB200_COMPUTE_FP32 || m->compute == BFB200_COMPUTE_F16 || m->compute == BFB200_COMPUTE_BF16,
              BFB200_ERR_INVALID_ARG, "unknown compute type %d", m->compute);
  if (m->compute != BFB200_COMPUTE_FP32) {
    BFB_REQUIRE(!(m->site_flags & BFB200_SITE_HEAD_IN), BFB200_ERR_UNSUPPORTED,
                "the Head.bayes input-dropout site runs on the fp32 path only (compute = BFB200_COMPUTE_FP32)");
    BFB_REQUIRE(use_fused(*m) || tc_shape_ok(*m), BFB200_ERR_UNSUPPORTED,
                "16-bit tensor-core compute serves (a) f16: d_model 64, 8 heads, F <= 64, V <= 128 (fused kernel) and "
                "(b) f16 / bf16: heads of width 64, F %% 8 == 0 (layered tensor-core path); got d=%d H=%d F=%d V=%d L=%d",
                m->d_model, m->n_heads, m->d_ff, m->ntoken, m->n_layers);
    BFB_DEV(m->fused_image);
  }
  if (m->site_flags != 0)
    BFB_REQUIRE(m->p_drop >= 0.f && m->p_drop < 1.f, BFB200_ERR_INVALID_ARG, "p_drop %f outside [0, 1)", m->p_drop);
  BFB_DEV(m->embedding); BFB_DEV(m->pe); BFB_DEV(m->head_w); BFB_DEV(m->head_b);
  if (m->n_layers > 0) {
    BFB_DEV(m->w_qkv); BFB_DEV(m->w_out); BFB_DEV(m->ln1_w); BFB_DEV(m->ln1_b); BFB_DEV(m->ff1_w);
    BFB_DEV(m->ff1_b); BFB_DEV(m->ff2_w); BFB_DEV(m->ff2_b); BFB_DEV(m->ln2_w); BFB_DEV(m->ln2_b);
  }
  return BFB200_OK;
}

static MaskSource make_mask_source(const bfb200_transformer& m, const bfb200_masks* masks) {
  MaskSource ms;
  memset(&ms, 0, sizeof(ms));
  const float p = m.site_flags != 0 ? m.p_drop : 0.f;
  ms.bits = masks != nullptr ? masks->bits : nullptr;
  const uint64_t seed = masks != nullptr ? masks->seed : 0ull;
  ms.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  ms.threshold = keep_threshold(p);
  ms.scale = 1.0f / (1.0f - p);
  ms.n_sites = 1 + BFB200_SITES_PER_LAYER * m.n_layers + ((m.site_flags & BFB200_SITE_HEAD_IN) ? 3 * m.n_heads * m.n_layers : 0);
  ms.max_len = masks != nullptr ? masks->max_len : 0;
  ms.words = (m.d_model + 31) / 32;
  ms.n_draws = 1;
  return ms;
}

// One forward pass over S sequences of `len` tokens (rows r = s * len + t); leaves the last
// block's output in ws.x.
static int forward_chunk(const bfb200_transformer& m, Workspace& ws, int64_t S, int len,
                         int tok_stride, const MaskSource& ms, cudaStream_t st) {
  const int d = m.d_model, F = m.d_ff;
  const int64_t R = S * len;
  const uint32_t fl = m.site_flags;
  BFB_REQUIRE(len <= m.pe_len, BFB200_ERR_INVALID_ARG, "sequence length %d exceeds the positional table (%d)",
              len, m.pe_len);
  BFB_CALL(launch_embed(ws.tokens, tok_stride, m.embedding, m.pe, S, len, d, ms,
                        (fl & BFB200_SITE_EMBED) ? 0 : -1, ws.x, st));
  for (int l = 0; l < m.n_layers; ++l) {
    const int sb = 1 + BFB200_SITES_PER_LAYER * l;
    const float* wqkv = m.w_qkv + (size_t)l * 3 * d * d;
    const float* wout = m.w_out + (size_t)l * d * d;
    if (fl & BFB200_SITE_HEAD_IN)   // Head.bayes: three masked copies of x per head (transformer.py:50-53)
      BFB_CALL(launch_qkv_head_in(ws.x, wqkv, ws.qkv, S, len, d, m.n_heads, ms,
                                  1 + BFB200_SITES_PER_LAYER * m.n_layers + 3 * m.n_heads * l, st));
    else
      BFB_CALL(launch_sgemm(ws.x, wqkv, nullptr, ws.qkv, R, 3 * d, d, 1, 0, 0, st));
    BFB_CALL(launch_attention(ws.qkv, ws.att, S, len, d, m.n_heads, ms,
                              (fl & BFB200_SITE_HEAD_OUT) ? sb + 4 : -1, st));
    BFB_CALL(launch_sgemm(ws.att, wout, nullptr, ws.y, R, d, d, 1, 0, 0, st));
    BFB_CALL(launch_resid_ln(ws.x, ws.y, ws.x1, R, len, d, m.ln1_w + (size_t)l * d, m.ln1_b + (size_t)l * d,
                             m.ln_eps, ms, (fl & BFB200_SITE_ATTN_RESID) ? sb + 1 : -1,
                             (fl & BFB200_SITE_FFN_IN) ? sb + 2 : -1, st));
    BFB_CALL(launch_sgemm(ws.x1, m.ff1_w + (size_t)l * F * d, m.ff1_b + (size_t)l * F, ws.h, R, F, d, 1, 0, 1, st));
    BFB_CALL(launch_sgemm(ws.h, m.ff2_w + (size_t)l * d * F, m.ff2_b + (size_t)l * d, ws.y, R, d, F, 1, 0, 0, st));
    // both residuals start from the BLOCK INPUT x (transformer.py:231-233)
    BFB_CALL(launch_resid_ln(ws.x, ws.y, ws.att, R, len, d, m.ln2_w + (size_t)l * d, m.ln2_b + (size_t)l * d,
                             m.ln_eps, ms, (fl & BFB200_SITE_FFN_RESID) ? sb + 3 : -1,
                             (fl & BFB200_SITE_LAYER_OUT) ? sb + 0 : -1, st));
    float* next = ws.att;  // the block output becomes the next block's input; `att` is free again
    ws.att = ws.x;
    ws.x = next;
  }
  return BFB200_OK;
}

// One forward pass of the layered tensor-core path; leaves the last block's output in ws.x16 (and, outside the fp16
// residual-stream mode, in ws.x as fp32).
static int forward_chunk_tc(const bfb200_transformer& m, Workspace& ws, int64_t S, int len, int tok_stride,
                            const MaskSource& ms, cudaStream_t st) {
  const int d = m.d_model, F = m.d_ff;
  const int64_t R = S * len;
  const uint32_t fl = m.site_flags, fmt = tc_fmt(m);
  const TcImage img = tc_image(m);
  const uint16_t* w16 = (const uint16_t*)m.fused_image;
  BFB_REQUIRE(len <= m.pe_len, BFB200_ERR_INVALID_ARG, "sequence length %d exceeds the positional table (%d)", len, m.pe_len);
  BFB_REQUIRE(len <= 256, BFB200_ERR_UNSUPPORTED, "the tensor-core attention kernel takes sequences of at most 256 tokens (got %d)", len);
  // LayerNorm1 is folded into the products around it whenever no dropout site touches its input or output (the fused
  // epilogues live in the two-SM GEMM kernel, which serves products with more than 128 output columns).
  // BFB200_TC_FP16_RESIDUAL=1 additionally carries the residual stream itself as fp16 (FFN2's epilogue adds it and
  // accumulates LayerNorm2's statistics, ln16_apply replaces the LayerNorm2 kernel, the fp32 copy of x is never
  // written).  Measured on C4: 1.5 % faster, worst probability error 0.39 instead of 0.17 of the 1e-2 budget — not
  // worth the accuracy, so the fp32 stream is the default.
  const bool fold = d > 128 && !(fl & (BFB200_SITE_ATTN_RESID | BFB200_SITE_FFN_IN));
  const char* lite_env = getenv("BFB200_TC_FP16_RESIDUAL");
  const bool lite = fold && fmt == umma::kFmtF16 && !(fl & BFB200_SITE_FFN_RESID) && lite_env != nullptr && lite_env[0] == '1';
  BFB_CALL(launch_embed(ws.tokens, tok_stride, m.embedding, m.pe, S, len, d, ms, (fl & BFB200_SITE_EMBED) ? 0 : -1,
                        lite ? nullptr : ws.x, st, ws.x16, fmt));
  for (int l = 0; l < m.n_layers; ++l) {
    const int sb = 1 + BFB200_SITES_PER_LAYER * l;
    BFB_CALL(launch_gemm_tc(ws.x16, d, w16 + img.qkv + (size_t)l * 3 * d * d, nullptr, ws.qkv16, nullptr, 3 * d, R, 3 * d, d, 0,
                            fmt, st));
    BFB_CALL(launch_attention_tc(ws.qkv16, ws.att16, S, len, d, m.n_heads, fmt, ms,
                                 (fl & BFB200_SITE_HEAD_OUT) ? sb + 4 : -1, st));
    // the branch outputs (W_out, FFN2) travel as 16-bit: they are added to the fp32 residual stream right away
    uint16_t* y16 = reinterpret_cast<uint16_t*>(ws.y);
    if (fold) {
      // No dropout site touches LayerNorm1's input or output (the reference default, transformer.py:226-228 latent):
      // LayerNorm1 only feeds the first FFN product, so it is folded into the two products around it instead of
      // running as a kernel.  W_out's epilogue adds the block input (its 16-bit copy), writes v = x + attn W_out^T
      // as the next A operand and accumulates each row's sum / sum of squares; the FFN1 product runs on v with
      // W1 * gamma1 and its epilogue applies rstd (acc - mean col_s) + col_c, ReLU.
      const cudaError_t e = cudaMemsetAsync(ws.stats, 0, (size_t)R * sizeof(RowStats), st);
      BFB_REQUIRE(e == cudaSuccess, BFB200_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e));
      GemmFusion add;
      add.resid16 = ws.x16; add.resid_pitch = d; add.stats_out = ws.stats;
      BFB_CALL(launch_gemm_tc_ex(ws.att16, d, w16 + img.out + (size_t)l * d * d, nullptr, y16, nullptr, d, R, d, d, 0, fmt, add, st));
      GemmFusion ln;
      ln.stats_in = ws.stats; ln.stats_width = d; ln.stats_eps = m.ln_eps;
      ln.col_s = reinterpret_cast<const float*>(w16 + img.col_s) + (size_t)l * F;
      ln.col_c = reinterpret_cast<const float*>(w16 + img.col_c) + (size_t)l * F;
      BFB_CALL(launch_gemm_tc_ex(y16, d, w16 + img.ff1g + (size_t)l * F * d, nullptr, ws.h16, nullptr, F, R, F, d, 1, fmt, ln, st));
    } else {
      BFB_CALL(launch_gemm_tc(ws.att16, d, w16 + img.out + (size_t)l * d * d, nullptr, y16, nullptr, d, R, d, d, 0, fmt, st));
      // LayerNorm1 output only feeds the FFN: 16-bit copy only (att16 is free again)
      BFB_CALL(launch_resid_ln(ws.x, ws.y, nullptr, R, len, d, m.ln1_w + (size_t)l * d, m.ln1_b + (size_t)l * d, m.ln_eps, ms,
                               (fl & BFB200_SITE_ATTN_RESID) ? sb + 1 : -1, (fl & BFB200_SITE_FFN_IN) ? sb + 2 : -1, st,
                               ws.att16, fmt, y16));
      BFB_CALL(launch_gemm_tc(ws.att16, d, w16 + img.ff1 + (size_t)l * F * d, m.ff1_b + (size_t)l * F, ws.h16, nullptr, F, R, F, d,
                              1, fmt, st));
    }
    if (lite) {
      // v2 = x + FFN(LN1(v1)) (both residuals start from the BLOCK INPUT x, transformer.py:231-233) and its row
      // statistics come out of the FFN2 product; LayerNorm2 + layer-out dropout is a 16-bit read-modify-write
      const cudaError_t e = cudaMemsetAsync(ws.stats, 0, (size_t)R * sizeof(RowStats), st);
      BFB_REQUIRE(e == cudaSuccess, BFB200_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e));
      GemmFusion add;
      add.resid16 = ws.x16; add.resid_pitch = d; add.stats_out = ws.stats;
      BFB_CALL(launch_gemm_tc_ex(ws.h16, F, w16 + img.ff2 + (size_t)l * d * F, m.ff2_b + (size_t)l * d, y16, nullptr, d, R, d, F, 0,
                                 fmt, add, st));
      BFB_CALL(launch_ln16_apply(y16, ws.stats, ws.x16, R, len, d, m.ln2_w + (size_t)l * d, m.ln2_b + (size_t)l * d, m.ln_eps, ms,
                                 (fl & BFB200_SITE_LAYER_OUT) ? sb + 0 : -1, fmt, st));
      continue;
    }
    BFB_CALL(launch_gemm_tc(ws.h16, F, w16 + img.ff2 + (size_t)l * d * F, m.ff2_b + (size_t)l * d, y16, nullptr, d, R, d, F, 0,
                            fmt, st));
    // both residuals start from the BLOCK INPUT x (transformer.py:231-233)
    BFB_CALL(launch_resid_ln(ws.x, ws.y, ws.x2, R, len, d, m.ln2_w + (size_t)l * d, m.ln2_b + (size_t)l * d, m.ln_eps, ms,
                             (fl & BFB200_SITE_FFN_RESID) ? sb + 3 : -1, (fl & BFB200_SITE_LAYER_OUT) ? sb + 0 : -1, st,
                             ws.x16, fmt, y16));
    float* t = ws.x; ws.x = ws.x2; ws.x2 = t;
  }
  return BFB200_OK;
}

}  // namespace bfb

using namespace bfb;

extern "C" int bfb200_compute_path(const bfb200_transformer* model) {
  if (model == nullptr) return -1;
  if (model->compute == BFB200_COMPUTE_FP32) return 0;
  if (use_fused(*model)) return 1;
  if (use_tc(*model) && tc_shape_ok(*model)) return 2;
  return -1;
}

extern "C" size_t bfb200_fused_image_bytes(const bfb200_transformer* model) {
  if (model == nullptr) return 0;
  if (use_fused(*model)) return fused_image_bytes(*model);
  if (use_tc(*model) && tc_shape_ok(*model)) return (tc_image(*model).elems * 2 + 255) & ~(size_t)255;
  return 0;
}

extern "C" int bfb200_fused_pack(const bfb200_transformer* model, void* image, size_t image_bytes, void* stream) {
  BFB_REQUIRE(model != nullptr, BFB200_ERR_INVALID_ARG, "model is NULL");
  BFB_DEV(image);
  BFB_DEV(model->embedding); BFB_DEV(model->pe); BFB_DEV(model->head_w); BFB_DEV(model->head_b);
  BFB_DEV(model->w_qkv); BFB_DEV(model->w_out); BFB_DEV(model->ln1_w); BFB_DEV(model->ln1_b); BFB_DEV(model->ff1_w);
  BFB_DEV(model->ff1_b); BFB_DEV(model->ff2_w); BFB_DEV(model->ff2_b); BFB_DEV(model->ln2_w); BFB_DEV(model->ln2_b);
  cudaStream_t st = (cudaStream_t)stream;
  const bfb200_transformer& m = *model;
  if (use_fused(m)) return fused_pack(m, image, image_bytes, st);
  BFB_REQUIRE(use_tc(m) && tc_shape_ok(m), BFB200_ERR_UNSUPPORTED, "no 16-bit weight image for this model / compute type");
  const TcImage img = tc_image(m);
  BFB_REQUIRE(image_bytes >= img.elems * 2, BFB200_ERR_WORKSPACE, "weight image buffer too small: %zu < %zu", image_bytes,
              img.elems * 2);
  const size_t L = m.n_layers, d = m.d_model, F = m.d_ff, V = m.ntoken;
  uint16_t* w16 = (uint16_t*)image;
  const uint32_t fmt = tc_fmt(m);
  BFB_CALL(launch_to16(m.w_qkv, w16 + img.qkv, (int64_t)(L * 3 * d * d), fmt, st));
  BFB_CALL(launch_to16(m.w_out, w16 + img.out, (int64_t)(L * d * d), fmt, st));
  BFB_CALL(launch_to16(m.ff1_w, w16 + img.ff1, (int64_t)(L * F * d), fmt, st));
  BFB_CALL(launch_to16(m.ff2_w, w16 + img.ff2, (int64_t)(L * d * F), fmt, st));
  BFB_CALL(launch_to16(m.head_w, w16 + img.head, (int64_t)(V * d), fmt, st));
  BFB_CALL(launch_fold_ln_pack(m.ff1_w, m.ff1_b, m.ln1_w, m.ln1_b, (int)L, (int)F, (int)d, fmt, w16 + img.ff1g,
                               reinterpret_cast<float*>(w16 + img.col_s), reinterpret_cast<float*>(w16 + img.col_c), st));
  return BFB200_OK;
}

extern "C" size_t bfb200_transformer_workspace_bytes(const bfb200_transformer* model, int64_t n_sequences,
                                                     int32_t max_len) {
  if (model == nullptr || n_sequences < 0 || max_len < 0) return 0;
  return workspace_bytes(*model, n_sequences, max_len, true);
}

extern "C" int bfb200_transformer_forward(const bfb200_transformer* model, const int64_t* src, int32_t len,
                                          int64_t batch, const bfb200_masks* masks, uint32_t draw,
                                          int64_t index_base, float* logp_out, void* workspace,
                                          size_t workspace_bytes_, void* stream) {
  BFB_CALL(validate_model(model));
  BFB_REQUIRE(!use_fused(*model), BFB200_ERR_UNSUPPORTED,
              "bfb200_transformer_forward (all-position log-probs) is served by the fp32 and the layered tensor-core paths; "
              "pack this model with compute=fp32");
  BFB_REQUIRE(len >= 0 && batch >= 0, BFB200_ERR_INVALID_ARG, "bad shape (%d, %lld)", len, (long long)batch);
  if (len == 0 || batch == 0) return BFB200_OK;
  BFB_DEV(src);
  BFB_DEV(logp_out);
  BFB_DEV(workspace);
  if (masks != nullptr) BFB_DEV_OPT(masks->bits);
  const bfb200_transformer& m = *model;
  BFB_REQUIRE(workspace_bytes_ >= workspace_bytes(m, batch, len, true), BFB200_ERR_WORKSPACE,
              "workspace too small: %zu < %zu", workspace_bytes_, workspace_bytes(m, batch, len, true));
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws = carve(workspace, m, batch, len, true);
  MaskSource ms = make_mask_source(m, masks);
  ms.n_seq = batch;
  ms.seq_base = 0;
  ms.step = 0;
  ms.n_draws = 1;
  ms.draw_base = draw;
  ms.item_base = index_base;
  if (ms.bits != nullptr)
    BFB_REQUIRE(ms.max_len >= len, BFB200_ERR_INVALID_ARG, "masks.max_len %d < len %d", ms.max_len, len);
  BFB_CALL(launch_init_tokens(src, batch, 0, batch, 1, len, len, m.ntoken, ws.tokens, st));
  if (use_tc(m)) {
    BFB_CALL(forward_chunk_tc(m, ws, batch, len, len, ms, st));
    BFB_CALL(launch_gemm_tc(ws.x16, m.d_model, (const uint16_t*)m.fused_image + tc_image(m).head, m.head_b, nullptr, ws.logits,
                            logits_pitch(m), batch * len, m.ntoken, m.d_model, 0, tc_fmt(m), st));
  } else {
    BFB_CALL(forward_chunk(m, ws, batch, len, len, ms, st));
    BFB_CALL(launch_sgemm(ws.x, m.head_w, m.head_b, ws.logits, batch * len, m.ntoken, m.d_model, 1, 0, 0, st));
  }
  BFB_CALL(launch_logsoftmax_seqfirst(ws.logits, batch, len, m.ntoken, logp_out, st, logits_pitch(m)));
  return BFB200_OK;
}

static int validate_mc(const bfb200_transformer* model, const bfb200_mc_args* a) {
  BFB_REQUIRE(a != nullptr, BFB200_ERR_INVALID_ARG, "args is NULL");
  BFB_REQUIRE(a->n_items >= 0 && a->prompt_len >= 1 && a->new_tokens >= 0 && a->n_draws >= 1,
              BFB200_ERR_INVALID_ARG, "bad shape (B=%lld, P=%d, N=%d, T=%d)", (long long)a->n_items,
              a->prompt_len, a->new_tokens, a->n_draws);
  BFB_REQUIRE(a->scheme >= BFB200_SCHEME_GREEDY && a->scheme <= BFB200_SCHEME_DROPOUT, BFB200_ERR_INVALID_ARG,
              "Unknown compute scheme %d", a->scheme);
  BFB_REQUIRE(a->mode >= BFB200_MODE_MAX && a->mode <= BFB200_MODE_ENTROPY, BFB200_ERR_INVALID_ARG,
              "Unknown mode %d. Available modes are 'max', 'margin' and 'entropy'.", a->mode);
  if (a->scheme == BFB200_SCHEME_SAMPLING)
    BFB_REQUIRE(a->temperature > 0.f, BFB200_ERR_INVALID_ARG, "temperature must be > 0");
  BFB_REQUIRE(a->prompt_len + a->new_tokens - 1 <= model->pe_len || a->new_tokens == 0, BFB200_ERR_INVALID_ARG,
              "P + N - 1 = %d exceeds the positional table (%d)", a->prompt_len + a->new_tokens - 1, model->pe_len);
  BFB_REQUIRE(a->prompt_len + a->new_tokens < 65536, BFB200_ERR_UNSUPPORTED, "sequence too long for the RNG counter");
  return BFB200_OK;
}

static int64_t max_chunk_items(const bfb200_transformer& m, const bfb200_mc_args& a, size_t bytes) {
  const int max_len = a.prompt_len + a.new_tokens;
  const size_t per_item = workspace_bytes(m, a.n_draws, max_len, false) - 256;
  const size_t fixed = 256 + 9 * 256;
  if (bytes <= fixed) return 0;
  int64_t n = (int64_t)((bytes - fixed) / per_item);
  // one item's buffers were aligned individually in per_item; the estimate is conservative
  return n;
}

extern "C" size_t bfb200_mc_workspace_bytes(const bfb200_transformer* model, const bfb200_mc_args* args,
                                            int64_t items_per_chunk) {
  if (model == nullptr || args == nullptr || items_per_chunk < 1) return 0;
  const int max_len = args->prompt_len + args->new_tokens;
  // sized so that max_chunk_items(bytes) >= items_per_chunk
  const size_t per_item = workspace_bytes(*model, args->n_draws, max_len, false) - 256;
  return 256 + 9 * 256 + per_item * (size_t)items_per_chunk;
}

extern "C" int bfb200_mc_score_transformer(const bfb200_transformer* model, const bfb200_mc_args* args,
                                           void* workspace, size_t workspace_bytes_, void* stream) {
  BFB_CALL(validate_model(model));
  BFB_CALL(validate_mc(model, args));
  const bfb200_transformer& m = *model;
  const bfb200_mc_args& a = *args;
  if (a.n_items == 0 || a.new_tokens == 0) return BFB200_OK;
  BFB_DEV(a.prompts);
  BFB_DEV(a.step_scores);
  BFB_DEV(workspace);
  BFB_DEV_OPT(a.pool_scores); BFB_DEV_OPT(a.logp_out); BFB_DEV_OPT(a.tokens_out);
  BFB_DEV_OPT(a.forced_tokens); BFB_DEV_OPT(a.masks.bits);
  cudaStream_t st = (cudaStream_t)stream;
  const int T = a.n_draws, P = a.prompt_len, N = a.new_tokens;
  const int max_len = P + N;  // room for the last appended token (only returned, never embedded)
  int64_t chunk = max_chunk_items(m, a, workspace_bytes_);
  BFB_REQUIRE(chunk >= 1, BFB200_ERR_WORKSPACE, "workspace too small for one pool item: %zu < %zu",
              workspace_bytes_, bfb200_mc_workspace_bytes(model, args, 1));
  if (chunk > a.n_items) chunk = a.n_items;
  if (a.masks.bits != nullptr)
    BFB_REQUIRE(a.masks.max_len >= P + N - 1, BFB200_ERR_INVALID_ARG, "masks.max_len %d < P + N - 1 = %d",
                a.masks.max_len, P + N - 1);
  const int64_t S_total = a.n_items * T;
  const bool fused = use_fused(m);
  if (fused)
    BFB_REQUIRE(P + N - 1 <= 128, BFB200_ERR_UNSUPPORTED,
                "compute=f16 (fused kernel) takes sequences of at most 128 tokens; P + N - 1 = %d", P + N - 1);
  if (use_tc(m))
    BFB_REQUIRE(P + N - 1 <= 256, BFB200_ERR_UNSUPPORTED,
                "the layered tensor-core path takes sequences of at most 256 tokens; P + N - 1 = %d", P + N - 1);

  for (int64_t item0 = 0; item0 < a.n_items; item0 += chunk) {
    const int64_t items = (a.n_items - item0 < chunk) ? a.n_items - item0 : chunk;
    const int64_t S = items * T;
    Workspace ws = carve(workspace, m, S, max_len, false);
    BFB_CALL(launch_init_tokens(a.prompts, a.n_items, item0, S, T, P, max_len, m.ntoken, ws.tokens, st));
    for (int step = 0; step < N; ++step) {
      const int len = P + step;
      MaskSource ms = make_mask_source(m, &a.masks);
      ms.n_seq = S_total;
      ms.seq_base = item0 * T;
      ms.step = (uint32_t)step;
      ms.n_draws = (uint32_t)T;
      ms.draw_base = 0;
      ms.item_base = a.index_base + item0;
      const int32_t* forced = a.forced_tokens != nullptr ? a.forced_tokens + (int64_t)step * S_total + item0 * T : nullptr;
      float* logp = a.logp_out != nullptr ? a.logp_out + ((int64_t)step * S_total + item0 * T) * m.ntoken : nullptr;
      if (fused) {
        BFB_CALL(launch_mc_forward_fused(m, ws.tokens, ws.tokens, max_len, S, len, a.scheme, a.mode, a.temperature,
                                         ms, forced, ws.seq_scores, logp, st));
      } else if (use_tc(m)) {
        BFB_CALL(forward_chunk_tc(m, ws, S, len, max_len, ms, st));
        // vocabulary head on the last position only: A rows are len * d elements apart, starting at position len - 1
        BFB_CALL(launch_gemm_tc(ws.x16 + (size_t)(len - 1) * m.d_model, (int64_t)len * m.d_model,
                                (const uint16_t*)m.fused_image + tc_image(m).head, m.head_b, nullptr, ws.logits,
                                logits_pitch(m), S, m.ntoken, m.d_model, 0, tc_fmt(m), st));
        BFB_CALL(launch_score_step(ws.logits, S, m.ntoken, a.scheme, a.mode, a.temperature, ms, forced, ws.tokens,
                                   max_len, len, ws.seq_scores, logp, st, logits_pitch(m)));
      } else {
        BFB_CALL(forward_chunk(m, ws, S, len, max_len, ms, st));
        // vocabulary head on the last position only (active_learning.py:62, :103, :146)
        BFB_CALL(launch_sgemm(ws.x, m.head_w, m.head_b, ws.logits, S, m.ntoken, m.d_model, len, len - 1, 0, st));
        BFB_CALL(launch_score_step(ws.logits, S, m.ntoken, a.scheme, a.mode, a.temperature, ms, forced, ws.tokens,
                                   max_len, len, ws.seq_scores, logp, st));
      }
      BFB_CALL(launch_reduce_draws(ws.seq_scores, items, T, a.step_scores + (int64_t)step * a.n_items + item0, st));
    }
    if (a.tokens_out != nullptr) {
      cudaError_t e = cudaMemcpyAsync(a.tokens_out + item0 * T * max_len, ws.tokens,
                                      (size_t)S * max_len * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
      BFB_REQUIRE(e == cudaSuccess, BFB200_ERR_CUDA, "token copy failed: %s", cudaGetErrorString(e));
    }
  }
  if (a.pool_scores != nullptr) BFB_CALL(launch_step_mean(a.step_scores, a.n_items, N, a.pool_scores, st));
  return BFB200_OK;
}

extern "C" int bfb200_step_mean(const float* step_scores, int32_t n_steps, int64_t n_items, float* pool_scores,
                                void* stream) {
  BFB_REQUIRE(n_steps >= 1 && n_items >= 0, BFB200_ERR_INVALID_ARG, "bad shape (N=%d, B=%lld)", n_steps,
              (long long)n_items);
  if (n_items == 0) return BFB200_OK;
  BFB_DEV(step_scores);
  BFB_DEV(pool_scores);
  return launch_step_mean(step_scores, n_items, n_steps, pool_scores, (cudaStream_t)stream);
}

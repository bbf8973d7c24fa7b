// Fused MC-forward kernel for the notebook-default model class (d_model = 64, 8 heads of 8, small F, V <= 128):
// ONE kernel runs the whole forward pass of a decode step — embedding + MC dropout + positional encoding, every
// TransformerBlock (QKV / W_out / FFN contractions on tcgen05 tensor cores, fp32 TMEM accumulators, causal attention,
// both residual + LayerNorm pairs, the always-on dropout sites), the last-position vocabulary head, log-softmax, the
// per-draw uncertainty score and the next token — with every activation resident on chip.
//
// Layout.  A CTA holds the fp16 weight image (pre-swizzled by fused_pack_kernel, fetched with 1-D bulk TMA copies) in
// shared memory and runs two independent 256-thread groups.  A group owns a tile of floor(128 / len) whole sequences
// = up to 128 token rows, two threads per row (32 features each):
//   * row phases    — read the fp32 accumulator row from TMEM (tcgen05.ld 32x32b), do bias / ReLU / residual /
//                     LayerNorm / dropout in registers and leave the next product's A operand IN TENSOR MEMORY
//                     (tcgen05.st; the "TS" form of tcgen05.mma reads it from there) as fp16 hi + lo parts;
//   * pair phase    — causal softmax(QK^T / sqrt(dh)) V in fp32 on CUDA cores, FOUR HEADS AT A TIME: the fp32 Q / K / V
//                     values of half the heads are copied from the accumulator into the three 16 KB tiles, work item =
//                     (query, sequence, head), one streamed pass over the keys; then the other half;
//   * score phase   — 8 or 16 threads per sequence: log-softmax over the vocabulary, probabilities, max /
//                     margin / entropy, greedy arg-max or inverse-CDF sample.
//
// Precision (the 16-bit class with margin; tools/emulate_fused_rounding.py).  The tensor pipe is
// idle most of the time here, so operands whose fp16 rounding matters are given as TWO fp16 numbers, hi + lo, and
// multiplied twice into the same fp32 accumulator:
//   - the residual stream x lives in TMEM as fp16 hi + lo (64 columns, 2^-22 relative): it is at once the residual and
//     the A operand of the QKV, FFN1 and vocabulary-head products;
//   - the FFN hidden row and the attention output likewise; the Q / K rows of W_qkv and W2 have a lo image too
//     (W_qkv's is streamed per layer from L2 into the dead V tile by a bulk copy);
//   - Q, K and V never leave fp32: attention works on the accumulator's values.
// What stays single fp16: W_v, W_out, W1, the vocabulary head's weights.
//
// With the reference's default dropout sites W_out and FFN1 are ONE product (`merged`): FFN1(LN1(v)), v = x + attn W_out^T,
// needs v only through its row statistics and through v (W1 g1)^T = x (W1 g1)^T + attn ((W1 g1) W_out)^T.
//
// HBM traffic per sequence and step: len token ids in, one score and one token out.
//
// Reference arithmetic: src/model/transformer.py:349-375 (forward), :211-234 (block), :37-65 (head),
// :151-170 (FFN); src/libs/active_learning.py:62-68 / :103-112 / :146-152 (per-step scoring).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace bfb {

namespace {

constexpr int FD = 64;          // d_model
constexpr int FH = 8;           // heads
constexpr int FDH = 8;          // head width
constexpr int FROWS = 128;      // token rows per group tile (UMMA M)
constexpr int FGROUPS = 2;      // independent groups per CTA, one tile of 128 token rows each
constexpr int FGROUP_THREADS = 256;   // two threads per token row (32 features each)
constexpr int FTHREADS = FGROUPS * FGROUP_THREADS;
constexpr int FMAXLEN = 128;    // longest sequence the fused path takes (one sequence per 128-row tile)
constexpr int FMAXV = 128;
constexpr uint32_t TILE_BYTES = FROWS * 128;  // one fp16 operand tile (128 rows x 64 halves)
// shared-memory tiles of a group: Q, K, V hold the fp32 values of four heads at a time during attention (V afterwards:
// landing zone of the next QKV product's W_lo rows; Q afterwards: attention output of heads 4..7 as [hi | lo], then the
// streamed vocabulary head); L: attention output of heads 0..3 as [hi | lo]
constexpr int T_Q = 0, T_K = 1, T_L = 2, T_V = 3, FTILES = 4;
constexpr uint32_t WLO_QK_BYTES = 2 * FD * 128, WLO_V_BYTES = FD * 128;   // lo image of one layer's W_qkv: Q and K rows, V rows
constexpr uint32_t WLO_BYTES = WLO_QK_BYTES + WLO_V_BYTES;
constexpr uint32_t HEAD_BYTES = FMAXV * 128;  // vocabulary head: streamed per tile into the dead Q tile
// TMEM columns of a group's 256-column window
constexpr uint32_t C_XHI = 0, C_XLO = 32, C_ACC = 64;
constexpr size_t kFusedSmemMax = 227 * 1024;
constexpr uint32_t kBarrierBytes = 64;
#ifndef BFB_KEY_UNROLL
#define BFB_KEY_UNROLL 1   // key loop of the pair phase (measured: rolled 39.3 ms per pass, x2 40.1, x4 39.4 — instruction fetch)
#endif
constexpr int kKeyUnroll = BFB_KEY_UNROLL;
constexpr uint32_t kTilesBytes = (uint32_t)(FGROUPS * FTILES) * TILE_BYTES;

struct FusedLayout {
  uint32_t fp;            // F rounded up to 16
  uint32_t merged;        // W_out + FFN1 as one product (no dropout site between attention and FFN1)
  uint32_t om_rows;       // rows of the [W_out ; (W1 g1) W_out] slot: 64 (+ fp when merged)
  uint32_t layer_bytes;   // per layer: W_qkv | [W_out ; M] | W1
  uint32_t off_om, off_w1;  // inside a layer (qkv at 0)
  uint32_t off_w2;        // absolute: tiles of 64 rows x 128 bytes; a layer's K = fp columns sit at K offset (l % w2_per_tile) * fp,
  uint32_t w2_per_tile;   //   their lo parts 32 columns further (fp <= 32; wider FFNs have no W2 lo part)
  uint32_t w2_lo;
  uint32_t off_params;    // absolute: fp32 vectors, params_per_layer floats per layer
  uint32_t params_per_layer;
  uint32_t smem_image_bytes;   // what the kernel keeps in shared memory
  uint32_t off_head;      // absolute (global memory only from here on): vocabulary head, HEAD_BYTES
  uint32_t off_headb;     // absolute: FMAXV fp32 head biases, -inf beyond the vocabulary (padded classes vanish from every softmax)
  uint32_t off_wlo;       // absolute: n_layers x WLO_BYTES
  uint32_t total_bytes;
};

// parameter vectors of a layer (floats from the layer's base): merged flow b2, ln2_w, ln2_b, col_s, col_c;
// otherwise b2, ln2_w, ln2_b, ln1_w, ln1_b, b1
constexpr uint32_t P_B2 = 0, P_LN2W = 64, P_LN2B = 128, P_COLS = 192, P_LN1W = 192, P_LN1B = 256, P_B1 = 320;

__host__ __device__ inline FusedLayout fused_layout(int L, int F, uint32_t site_flags) {
  FusedLayout lay;
  lay.fp = (uint32_t)((F + 15) / 16 * 16);
  lay.merged = (site_flags & (BFB200_SITE_ATTN_RESID | BFB200_SITE_FFN_IN)) ? 0u : 1u;
  lay.om_rows = FD + (lay.merged ? lay.fp : 0u);
  lay.off_om = 3 * FD * 128;
  lay.off_w1 = lay.off_om + lay.om_rows * 128;
  lay.layer_bytes = lay.off_w1 + lay.fp * 128;
  lay.off_w2 = (uint32_t)L * lay.layer_bytes;
  lay.w2_lo = lay.fp <= 32u ? 1u : 0u;
  lay.w2_per_tile = lay.w2_lo ? 32u / lay.fp : 1u;
  const uint32_t w2_tiles = ((uint32_t)L + lay.w2_per_tile - 1) / lay.w2_per_tile;
  lay.off_params = lay.off_w2 + w2_tiles * (FD * 128);
  lay.params_per_layer = lay.merged ? P_COLS + 2 * lay.fp : P_B1 + lay.fp;
  lay.smem_image_bytes = (lay.off_params + (uint32_t)L * lay.params_per_layer * 4 + 15u) & ~15u;
  lay.off_head = (lay.smem_image_bytes + 1023u) & ~1023u;
  lay.off_headb = lay.off_head + HEAD_BYTES;
  lay.off_wlo = lay.off_headb + 1024u;
  lay.total_bytes = lay.off_wlo + (uint32_t)L * WLO_BYTES;
  return lay;
}

inline size_t fused_smem_bytes(const FusedLayout& lay) {   // tiles + image + barriers, before the alignment pad
  return (size_t)kTilesBytes + lay.smem_image_bytes + kBarrierBytes;
}

// ------------------------------------------------------------------------------------------
// weight image: fp16 matrices in the swizzled K-major layout + fp32 vectors (+ the global-only W_lo part)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sat16(float w) { return fminf(fmaxf(w, -65504.f), 65504.f); }
__device__ __forceinline__ float round16(float w) { return __half2float(__float2half_rn(sat16(w))); }

__global__ void fused_pack_kernel(bfb200_transformer m, FusedLayout lay, uint8_t* __restrict__ image) {
  const int L = m.n_layers, F = m.d_ff, V = m.ntoken;
  const bool merged = lay.merged != 0;
  // element counts per layer: W_qkv, W_out, M (merged), W1, W2, W_qkv lo
  const int n_qkv = 3 * FD * FD, n_out = FD * FD, n_m = merged ? F * FD : 0, n_w1 = F * FD, n_w2 = FD * F, n_lo = 3 * FD * FD;
  const int per_layer = n_qkv + n_out + n_m + n_w1 + n_w2 + n_lo;
  const int n_mat = L * per_layer + V * FD;
  const int n_par = L * (int)lay.params_per_layer;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < FMAXV; i += gridDim.x * blockDim.x)
    reinterpret_cast<float*>(image + lay.off_headb)[i] = i < V ? m.head_b[i] : -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_mat + n_par; i += gridDim.x * blockDim.x) {
    if (i < n_mat) {
      float w;
      uint32_t base, n, k;
      if (i < L * per_layer) {
        const int l = i / per_layer;
        int j = i - l * per_layer;
        base = (uint32_t)l * lay.layer_bytes;
        if (j < n_qkv) {                             // W_qkv (3d, d)
          n = j / FD; k = j % FD; w = m.w_qkv[(size_t)l * n_qkv + j];
        } else if ((j -= n_qkv) < n_out) {           // W_out (d, d)
          n = j / FD; k = j % FD; w = m.w_out[(size_t)l * n_out + j]; base += lay.off_om;
        } else if ((j -= n_out) < n_m) {             // M = fp16(W1 g1) W_out (F, d): the attention rows' share of the
          n = j / FD; k = j % FD;                    // merged product, right behind W_out (rows 64 .. 64 + F)
          w = 0.f;
          for (int q = 0; q < FD; ++q)
            w = fmaf(round16(m.ff1_w[((size_t)l * F + n) * FD + q] * m.ln1_w[l * FD + q]), m.w_out[((size_t)l * FD + q) * FD + k], w);
          n += FD; base += lay.off_om;
        } else if ((j -= n_m) < n_w1) {              // FFN W1 (F, d), times gamma1 when LayerNorm1 is folded into it
          n = j / FD; k = j % FD; w = m.ff1_w[(size_t)l * n_w1 + j]; base += lay.off_w1;
          if (merged) w *= m.ln1_w[l * FD + k];
        } else if ((j -= n_w1) < n_w2) {             // FFN W2 (d, F): K = F, layer l at K offset (l % w2_per_tile) * fp,
          n = j / F; k = j % F + (l % lay.w2_per_tile) * lay.fp; w = m.ff2_w[(size_t)l * n_w2 + j];   // lo part 32 further
          base = lay.off_w2 + (uint32_t)(l / lay.w2_per_tile) * (FD * 128);
          if (lay.w2_lo) {
            const uint32_t off_lo = base + umma::sw128_offset(n, (k + 32u) >> 3) + (k & 7u) * 2u;
            *reinterpret_cast<__half*>(image + off_lo) = __float2half_rn(sat16(w - round16(w)));
          }
        } else {                                     // lo parts of W_qkv (global memory only): Q and K rows, then the V rows
          j -= n_w2;                                 // as a tile of their own (they land in different shared-memory tiles)
          n = j / FD; k = j % FD;
          const float full = m.w_qkv[(size_t)l * n_qkv + j];
          w = full - round16(full);
          base = lay.off_wlo + (uint32_t)l * WLO_BYTES;
          if (n >= 2u * FD) { n -= 2u * FD; base += WLO_QK_BYTES; }
        }
      } else {                                       // vocabulary head (V, d)
        const int j = i - L * per_layer;
        n = j / FD; k = j % FD; w = m.head_w[j]; base = lay.off_head;
      }
      const uint32_t off = base + umma::sw128_offset(n, k >> 3) + (k & 7u) * 2u;
      *reinterpret_cast<__half*>(image + off) = __float2half_rn(sat16(w));
    } else {
      const int j = i - n_mat;
      const int l = j / (int)lay.params_per_layer, q = j % (int)lay.params_per_layer;
      float v = 0.f;
      if (q < (int)P_LN2W) v = m.ff2_b[l * FD + q];
      else if (q < (int)P_LN2B) v = m.ln2_w[l * FD + q - P_LN2W];
      else if (q < (int)P_COLS) v = m.ln2_b[l * FD + q - P_LN2B];
      else if (merged) {
        const int c = (q - (int)P_COLS) % (int)lay.fp;
        if (c < F) {
          if (q < (int)(P_COLS + lay.fp)) {   // col_s[n] = sum_k fp16(W1[n, k] g1[k]) — of the rounded values the tensor core multiplies
            for (int k = 0; k < FD; ++k) v += round16(m.ff1_w[((size_t)l * F + c) * FD + k] * m.ln1_w[l * FD + k]);
          } else {                            // col_c[n] = b1[n] + sum_k W1[n, k] beta1[k]
            v = m.ff1_b[l * F + c];
            for (int k = 0; k < FD; ++k) v = fmaf(m.ff1_w[((size_t)l * F + c) * FD + k], m.ln1_b[l * FD + k], v);
          }
        }
      } else if (q < (int)P_LN1B) v = m.ln1_w[l * FD + q - P_LN1W];
      else if (q < (int)P_B1) v = m.ln1_b[l * FD + q - P_LN1B];
      else v = (q - (int)P_B1) < F ? m.ff1_b[l * F + q - P_B1] : 0.f;
      reinterpret_cast<float*>(image + lay.off_params)[j] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------
struct FusedArgs {
  const uint8_t* image;
  const float* embedding;      // (V, 64) fp32
  const float* pe;             // (pe_len, 64) fp32
  FusedLayout lay;
  int n_layers, d_ff, vocab;
  uint32_t site_flags;
  float ln_eps;
  uint32_t smem_bytes;         // dynamic shared memory of the launch
  uint32_t opaque_zero;        // 0, unknown to the compiler: ties work meant to overlap a product to the wait that follows it
  // this launch
  int len;                     // tokens per sequence in this step
  int spt;                     // sequences per tile = 128 / len
  int tps;                     // score-phase threads per sequence (8 or 16)
  int64_t n_seq;               // sequences in the chunk
  int64_t n_tiles;
  const int32_t* tokens;       // (n_seq, tok_stride)
  int32_t* tokens_out;         // same buffer (token appended at column len) or nullptr
  int tok_stride;
  int scheme, mode;
  float temperature;
  const int32_t* forced;       // (n_seq) or nullptr
  float* seq_scores;           // (n_seq)
  float* logp_out;             // (n_seq, V) or nullptr
  MaskSource ms;
};

// per-row addressing of the mask source, hoisted out of the feature loops
struct RowRng {
  uint32_t item, draw;
  uint32_t seq_local;  // sequence inside this launch (bits mode adds seq_base)
};

__device__ __forceinline__ RowRng make_row_rng(const MaskSource& m, uint32_t s) {
  RowRng r;
  const uint32_t q = s / m.n_draws;
  r.item = (uint32_t)((uint64_t)m.item_base + q);
  r.draw = m.draw_base + (s - q * m.n_draws);
  r.seq_local = s;
  return r;
}

// 32 keep bits of one half row from the Philox stream: features [8 * c3_call, 8 * c3_call + 32) of the (row, site)
// encoded in (c2, c3): four calls of eight 16-bit decisions each.  Deliberately NOT inlined: the dropout
// sites of the generic instantiation share this one copy of the rounds (the kernel is instruction-fetch sensitive).
__device__ __noinline__ uint32_t philox_keep32(uint32_t key_x, uint32_t key_y, uint32_t threshold, uint32_t item,
                                               uint32_t draw, uint32_t c2, uint32_t c3) {
  const uint2 key = make_uint2(key_x, key_y);
  const uint4 r0 = philox4x32_10(make_uint4(item, draw, c2, c3 + 0), key);
  const uint4 r1 = philox4x32_10(make_uint4(item, draw, c2, c3 + 1), key);
  const uint4 r2 = philox4x32_10(make_uint4(item, draw, c2, c3 + 2), key);
  const uint4 r3 = philox4x32_10(make_uint4(item, draw, c2, c3 + 3), key);
  return keep8_of(r0, threshold) | (keep8_of(r1, threshold) << 8) | (keep8_of(r2, threshold) << 16) |
         (keep8_of(r3, threshold) << 24);
}

using umma::fh_ll;
using umma::fh_hh;
using umma::fh_lh;
using umma::fh_hl;

__device__ __forceinline__ float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// reductions over an aligned group of `n` (power of two <= 32) lanes
__device__ __forceinline__ float sub_max(float v, int n) {
  for (int o = n >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sub_sum(float v, int n) {
  for (int o = n >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// score phase: TPS threads per sequence over the staged last-position logits.
//   log-probs = log_softmax(logits)                                   transformer.py:375
//   probs     = softmax(log-probs [/ temperature])                    active_learning.py:67, :104-105, :147
//   score     = max p | p(1) - p(2) | -sum p log(p + 1e-10)           active_learning.py:25-31
//   token     = argmax log-probs | inverse-CDF sample | forced        active_learning.py:63, :106, :148
// softmax(log-probs) is exp(log-probs) up to rounding, so the un-tempered probabilities reuse the exponentials
// of the log-softmax, and log(p + 1e-10) is taken as log p (the terms where they differ weigh < 1e-8).
// ------------------------------------------------------------------------------------------
template <int TPS, bool NO_SAMPLING>
__device__ __forceinline__ void score_phase(const FusedArgs& a, const float* __restrict__ stage, int stage_stride,
                                            int gt, int nseq, uint32_t s0, int len) {
  constexpr int NPT = FMAXV / TPS;
  const int V = a.vocab;
  const int sl = gt / TPS, sub = gt % TPS;
  const bool active = sl < nseq;
  const float* row = stage + (active ? sl : 0) * stage_stride;
  const uint32_t sq = s0 + (active ? (uint32_t)sl : 0u);
  float lg[NPT], pr[NPT];
  float mx = -INFINITY;
  int best_i = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    const int c = sub + j * TPS;
    lg[j] = row[c];   // classes >= V hold -inf (head bias padding)
    if (lg[j] > mx) { mx = lg[j]; best_i = c; }
  }
#pragma unroll
  for (int o = TPS >> 1; o > 0; o >>= 1) {   // arg-max, lowest class wins ties
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (om > mx || (om == mx && oi < best_i)) { mx = om; best_i = oi; }
  }
  float z = 0.f;
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    pr[j] = __expf(lg[j] - mx);
    z += pr[j];
  }
  z = sub_sum(z, TPS);
  const float lse = __logf(z);
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    const int c = sub + j * TPS;
    lg[j] = (lg[j] - mx) - lse;   // log-probabilities (class >= V: -inf)
    if (a.logp_out != nullptr && active && c < V) a.logp_out[(size_t)sq * V + c] = lg[j];
  }
  float ent = 0.f, p1, p2 = -INFINITY;
  if (NO_SAMPLING || a.scheme != BFB200_SCHEME_SAMPLING) {
    // un-tempered distribution: p_j = e_j / z with e_j = exp(logit_j - max) already at hand, so only the requested
    // score is formed (warp-uniform branches): max p = 1 / z; margin needs the runner-up; entropy sum e_j log p_j
    const float inv_z = __fdividef(1.0f, z);
    p1 = inv_z;
    if (a.mode == BFB200_MODE_ENTROPY) {
#pragma unroll
      for (int j = 0; j < NPT; ++j) ent = fmaf(pr[j], fmaxf(lg[j], -1e30f), ent);   // 0 * -inf guard (padded classes)
      ent = -sub_sum(ent, TPS) * inv_z;
    } else if (a.mode == BFB200_MODE_MARGIN) {
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        const int c = sub + j * TPS;
        if (c != best_i) p2 = fmaxf(p2, pr[j]);
      }
      p2 = sub_max(p2, TPS) * inv_z;
    }
  } else {
    // tempered distribution: softmax(log-probs / temperature); its maximum sits at the arg-max class
    const float inv_t = 1.0f / a.temperature;
    float pz = 0.f;
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      lg[j] = (lg[j] + lse) * inv_t;   // (logit - max) / T  <= 0
      pr[j] = __expf(lg[j]);
      pz += pr[j];
    }
    pz = sub_sum(pz, TPS);
    const float inv_pz = __fdividef(1.0f, pz), log_pz = __logf(pz);
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      pr[j] *= inv_pz;
      ent = fmaf(pr[j], fmaxf(lg[j] - log_pz, -1e30f), ent);
    }
    ent = -sub_sum(ent, TPS);
    p1 = -INFINITY;
    int p1_i = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int c = sub + j * TPS;
      if (pr[j] > p1) { p1 = pr[j]; p1_i = c; }
    }
#pragma unroll
    for (int o = TPS >> 1; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, p1, o);
      const int oi = __shfl_xor_sync(0xffffffffu, p1_i, o);
      if (ob > p1 || (ob == p1 && oi < p1_i)) { p1 = ob; p1_i = oi; }
    }
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int c = sub + j * TPS;
      if (c != p1_i) p2 = fmaxf(p2, pr[j]);
    }
    p2 = sub_max(p2, TPS);
  }

  int token = best_i;
  if (!NO_SAMPLING && a.scheme == BFB200_SCHEME_SAMPLING) {
    // inverse CDF in class order from one Philox word (the library's documented sampling rule)
    const RowRng srr = make_row_rng(a.ms, sq);
    const uint4 rnd = philox4x32_10(
        make_uint4(srr.item, srr.draw, a.ms.step << 16, BFB200_SITE_SAMPLING_TOKEN << 16), a.ms.key);
    const float u = ((float)(rnd.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float running = 0.f;
    int first = V - 1;
#pragma unroll 1
    for (int j = 0; j < NPT; ++j) {
      float c = 0.f;
#pragma unroll
      for (int q = 0; q < NPT; ++q) c = q == j ? pr[q] : c;   // register select keeps pr[] out of local memory
#pragma unroll
      for (int o = 1; o < TPS; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, c, o, TPS);
        if (sub >= o) c += up;
      }
      c += running;
      const int cls = sub + j * TPS;
      if (c > u && cls < V && cls < first) first = cls;
      running = __shfl_sync(0xffffffffu, c, TPS - 1, TPS);
    }
#pragma unroll
    for (int o = TPS >> 1; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    token = first;
  }
  if (active && sub == 0) {
    a.seq_scores[sq] = a.mode == BFB200_MODE_MAX ? p1 : (a.mode == BFB200_MODE_MARGIN ? p1 - p2 : ent);
    if (a.forced != nullptr) token = a.forced[sq];
    if (a.tokens_out != nullptr && len < a.tok_stride) a.tokens_out[(size_t)sq * a.tok_stride + len] = token;
  }
}
// ------------------------------------------------------------------------------------------
// half-row helpers: a token row's 64 features are split between two threads (features [32h, 32h + 32))
// ------------------------------------------------------------------------------------------
constexpr int HW = FD / 2;   // features per thread

__device__ __forceinline__ void drop_half(float (&v)[HW], uint32_t k, float scale) {
#pragma unroll
  for (uint32_t j = 0; j < HW; ++j) v[j] = ((k >> j) & 1u) ? v[j] * scale : 0.f;
}

// 32 keep bits (features [32h, 32h + 32)) of one (row, site)
__device__ __forceinline__ uint32_t half_keep_mask(const MaskSource& m, const RowRng& rr, uint32_t pos, uint32_t site,
                                                   uint32_t h) {
  if (m.bits != nullptr) {
    const int64_t row = (((int64_t)m.step * m.n_sites + site) * m.n_seq + (m.seq_base + rr.seq_local)) * m.max_len + pos;
    return m.bits[row * m.words + h];
  }
  return philox_keep32(m.key.x, m.key.y, m.threshold, rr.item, rr.draw, (m.step << 16) | pos, (site << 16) | (4u * h));
}

// FAST == 1 (two dropout sites): the four Philox calls of a half row decide straight on the 16-bit halves of their
// output words, eight features at a time (same decisions as philox_keep32 + drop_half: feature j of a call <- half j & 1
// of word j >> 1, kept iff >= threshold).  In two steps, so that the random words can be drawn while the group waits for
// a product (they depend on the row's identity only) and applied after it: philox_half_raw during the wait,
// drop_half_raw after; drop_half_direct does both at once.
__device__ __forceinline__ void philox_half_raw(uint4 (&r)[4], const MaskSource& m, const RowRng& rr, uint32_t pos,
                                                uint32_t site, uint32_t h) {
#pragma unroll
  for (uint32_t c = 0; c < 4; ++c) {
    r[c] = philox4x32_10(make_uint4(rr.item, rr.draw, (m.step << 16) | pos, (site << 16) | (4u * h + c)), m.key);
  }
}

// A word that depends on every draw; the caller ANDs it with a run-time zero into the wait's operand so that ptxas
// cannot schedule the draws after the wait (it otherwise hoists the first try_wait above independent arithmetic).
__device__ __forceinline__ uint32_t fold_raw(const uint4 (&r)[4]) {
  uint32_t f = 0u;
#pragma unroll
  for (uint32_t c = 0; c < 4; ++c) f ^= r[c].x ^ r[c].y ^ r[c].z ^ r[c].w;
  return f;
}

__device__ __forceinline__ void drop_half_raw(float (&v)[HW], const uint4 (&r)[4], const MaskSource& m) {
  const uint32_t thr_hi = m.threshold << 16;
#pragma unroll
  for (uint32_t c = 0; c < 4; ++c) {
    const uint32_t w[4] = {r[c].x, r[c].y, r[c].z, r[c].w};
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
      v[8 * c + 2 * i] = (w[i] << 16) >= thr_hi ? v[8 * c + 2 * i] * m.scale : 0.f;
      v[8 * c + 2 * i + 1] = w[i] >= thr_hi ? v[8 * c + 2 * i + 1] * m.scale : 0.f;
    }
  }
}

__device__ __forceinline__ void drop_half_direct(float (&v)[HW], const MaskSource& m, const RowRng& rr, uint32_t pos,
                                                 uint32_t site, uint32_t h) {
  uint4 r[4];
  philox_half_raw(r, m, rr, pos, site, h);
  drop_half_raw(v, r, m);
}

__device__ __forceinline__ void tmem_load_half(uint32_t taddr, float (&v)[HW]) {
  uint32_t r[32];
  umma::tmem_ld32(taddr, r);
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < HW; ++i) v[i] = __uint_as_float(r[i]);
}

// v += x for this thread's 32 features of the residual stream (fp16 hi + lo words in TMEM: 16 + 16 columns)
__device__ __forceinline__ void add_residual_half(uint32_t t_hi, uint32_t t_lo, float (&v)[HW]) {
  uint32_t hi[16], lo[16];
  umma::tmem_ld16(t_hi, hi);
  umma::tmem_ld16(t_lo, lo);
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[2 * i] = umma::add_hilo_l(hi[i], lo[i], v[2 * i]);
    v[2 * i + 1] = umma::add_hilo_h(hi[i], lo[i], v[2 * i + 1]);
  }
}

// this thread's 32 features as the fp16 hi + lo words of a TS-form A operand (and of the residual stream)
__device__ __forceinline__ void store_hilo_half(uint32_t t_hi, uint32_t t_lo, const float (&v)[HW]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) umma::split_f16x2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  umma::tmem_st16(t_hi, hi);
  umma::tmem_st16(t_lo, lo);
}

// LayerNorm over a 64-feature row held as two 32-feature halves in two threads (biased variance, eps inside the
// sqrt: nn.LayerNorm).  Each half computes its mean and centred sum of squares exactly (ln_partial), the pair
// exchanges them through `red` across a group barrier and combines with the parallel-variance formula
// (n_a = n_b = 32):  mean = (m_a + m_b) / 2,  M2 = M2_a + M2_b + 16 (m_a - m_b)^2      (ln_finish)
__device__ __forceinline__ float2 ln_partial(const float (&v)[HW]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < HW; i += 4) { s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3]; }
  const float m_own = ((s0 + s1) + (s2 + s3)) * (1.0f / HW);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int i = 0; i < HW; i += 2) {
    const float c0 = v[i] - m_own, c1 = v[i + 1] - m_own;
    q0 = fmaf(c0, c0, q0);
    q1 = fmaf(c1, c1, q1);
  }
  return make_float2(m_own, q0 + q1);
}
__device__ __forceinline__ void ln_combine(float2 own, float2 other, float eps, float& mean, float& rstd) {
  const float dm = own.x - other.x;
  mean = 0.5f * (own.x + other.x);
  rstd = rsqrtf((own.y + other.y + 16.0f * dm * dm) * (1.0f / FD) + eps);
}
__device__ __forceinline__ void ln_finish(float (&v)[HW], float2 own, float2 other, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float eps) {
  float mean, rstd;
  ln_combine(own, other, eps, mean, rstd);
  const float shift = -mean * rstd;
#pragma unroll
  for (int i = 0; i < HW; ++i) v[i] = fmaf(fmaf(v[i], rstd, shift), gamma[i], beta[i]);
}

#ifdef BFB_FUSED_TIMING
// debug build only (make EXTRA=-DBFB_FUSED_TIMING): cycles one observer thread per group spends in the product
// round trips (barrier -> MMA issue -> completion wait), and in total
__device__ unsigned long long bfb_fused_dbg[8];
__device__ unsigned long long bfb_fused_phase[16];   // cycles per phase, summed over the observer lanes
#define BFB_PHASE(i) do { const long long _n = clock64(); ph[i] += _n - ph_last; ph_last = _n; } while (0)
#else
#define BFB_PHASE(i) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------
// The MMAs of one product, issued by ONE thread of group G.  Every operand is built from kernel-uniform values (the
// shared-memory window base, kernel parameters, the compile-time group index; the TMEM allocation is the whole
// 512 columns, so its base is column 0) and therefore lives in uniform registers.
// ------------------------------------------------------------------------------------------
enum { PR_QKV = 0, PR_MERGED, PR_WOUT, PR_FFN1, PR_FFN2, PR_HEAD };
template <int K> struct KindC { static constexpr int value = K; };

struct IssueCtx {
  uint32_t tiles_u32;   // shared-memory address of group 0's first tile
  uint32_t img_u32;     // shared-memory address of the weight image
  uint32_t fp;
  uint32_t merged;
  uint32_t w2_per_tile;
  uint32_t w2_lo;
  uint32_t attn_lo;     // the pair phase left the attention output's lo parts in the L tile
  uint32_t off_om, off_w1, off_w2, layer_bytes;
};

template <int G, int kind>
__device__ __forceinline__ void fused_issue(const IssueCtx c, int layer, uint64_t* group_bar) {
  const uint32_t tile0 = c.tiles_u32 + (uint32_t)(G * FTILES) * TILE_BYTES;
  const uint32_t gb = (uint32_t)G * 256u;       // TMEM window of the group
  const uint32_t wl = c.img_u32 + (uint32_t)layer * c.layer_bytes;
  const uint32_t fp = c.fp;
  if (kind == PR_QKV) {
    // [Q | K | V] = x W_qkv^T with x = hi + lo from TMEM; the Q and K rows of W_qkv additionally as hi + lo (their lo
    // rows were streamed into the V tile; the x_lo W_lo term is dropped)
    const uint64_t dw = umma::smem_desc_sw128(wl), dl = umma::smem_desc_sw128(tile0 + T_V * TILE_BYTES);
    const uint32_t id192 = umma::instr_desc_f16(FROWS, 3 * FD), id128 = umma::instr_desc_f16(FROWS, 2 * FD);
    for (int k = 0; k < 4; ++k) umma::mma_f16_ts(gb + C_ACC, gb + C_XHI + 8u * k, dw + 2 * k, id192, k > 0);
    for (int k = 0; k < 4; ++k) umma::mma_f16_ts(gb + C_ACC, gb + C_XLO + 8u * k, dw + 2 * k, id192, 1u);
    for (int k = 0; k < 4; ++k) umma::mma_f16_ts(gb + C_ACC, gb + C_XHI + 8u * k, dl + 2 * k, id128, 1u);
  } else if (kind == PR_MERGED) {
    // columns [0, 64) = attn W_out^T ; [64, 64 + fp) = attn ((W1 g1) W_out)^T + x (W1 g1)^T
    // The attention output sits in two tiles: features 0..31 (heads 0..3) in the L tile, features 32..63 in the Q tile,
    // each row as [fp16 hi of its 32 features | their lo parts]: K steps 0, 1 read the L tile, 2, 3 the Q tile; the lo
    // operand of a step starts 64 bytes (descriptor + 4) behind its hi operand.
    const uint64_t daL = umma::smem_desc_sw128(tile0 + T_L * TILE_BYTES), daQ = umma::smem_desc_sw128(tile0 + T_Q * TILE_BYTES);
    const uint64_t dom = umma::smem_desc_sw128(wl + c.off_om);
    const uint64_t dw1 = umma::smem_desc_sw128(wl + c.off_w1);
    const uint32_t id_om = umma::instr_desc_f16(FROWS, FD + fp), id_fp = umma::instr_desc_f16(FROWS, fp);
    for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(gb + C_ACC, (k < 2 ? daL : daQ) + 2 * (k & 1), dom + 2 * k, id_om, k > 0);
    for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(gb + C_ACC, (k < 2 ? daL : daQ) + 4 + 2 * (k & 1), dom + 2 * k, id_om, 1u);
    for (int k = 0; k < 4; ++k) umma::mma_f16_ts(gb + C_ACC + FD, gb + C_XHI + 8u * k, dw1 + 2 * k, id_fp, 1u);
    for (int k = 0; k < 4; ++k) umma::mma_f16_ts(gb + C_ACC + FD, gb + C_XLO + 8u * k, dw1 + 2 * k, id_fp, 1u);
  } else if (kind == PR_WOUT) {
    const uint64_t daL = umma::smem_desc_sw128(tile0 + T_L * TILE_BYTES), daQ = umma::smem_desc_sw128(tile0 + T_Q * TILE_BYTES);
    const uint64_t dom = umma::smem_desc_sw128(wl + c.off_om);
    const uint32_t id64 = umma::instr_desc_f16(FROWS, FD);
    for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(gb + C_ACC, (k < 2 ? daL : daQ) + 2 * (k & 1), dom + 2 * k, id64, k > 0);
    for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(gb + C_ACC, (k < 2 ? daL : daQ) + 4 + 2 * (k & 1), dom + 2 * k, id64, 1u);
  } else if (kind == PR_FFN1) {
    // unmerged flow: LayerNorm1 output (hi at columns [64, 96), lo at [96, 128) of the accumulator window) times W1
    const uint64_t dw1 = umma::smem_desc_sw128(wl + c.off_w1);
    const uint32_t id_fp = umma::instr_desc_f16(FROWS, fp);
    for (int k = 0; k < 4; ++k) umma::mma_f16_ts(gb + C_ACC + 128u, gb + C_ACC + 64u + 8u * k, dw1 + 2 * k, id_fp, k > 0);
    for (int k = 0; k < 4; ++k) umma::mma_f16_ts(gb + C_ACC + 128u, gb + C_ACC + 96u + 8u * k, dw1 + 2 * k, id_fp, 1u);
  } else if (kind == PR_FFN2) {
    // FFN hidden row (hi + lo, 32 columns each, at accumulator-window column 128 in the merged flow — the FFN1
    // accumulator still occupies [64, 64 + fp) — and at 64 otherwise) times W2 -> columns [0, 64)
    const uint32_t w2 = c.img_u32 + c.off_w2 + ((uint32_t)layer / c.w2_per_tile) * (FD * 128u) +
                        ((uint32_t)layer % c.w2_per_tile) * fp * 2u;
    const uint64_t dw2 = umma::smem_desc_sw128(w2);
    const uint32_t id64 = umma::instr_desc_f16(FROWS, FD);
    const int ks = (int)(fp >> 4);
    const uint32_t hcol = gb + C_ACC + (c.merged ? 128u : 64u);
    for (int k = 0; k < ks; ++k) umma::mma_f16_ts(gb + C_ACC, hcol + 8u * k, dw2 + 2 * k, id64, k > 0);
    for (int k = 0; k < ks; ++k) umma::mma_f16_ts(gb + C_ACC, hcol + 32u + 8u * k, dw2 + 2 * k, id64, 1u);
    if (c.w2_lo)   // W2's lo columns sit 32 K-columns (64 bytes) behind the hi ones
      for (int k = 0; k < ks; ++k) umma::mma_f16_ts(gb + C_ACC, hcol + 8u * k, dw2 + 4 + 2 * k, id64, 1u);
  } else {
    // vocabulary head on x = hi + lo: 128 classes; its weights were streamed into the (dead) Q tile
    const uint64_t dh = umma::smem_desc_sw128(tile0 + T_Q * TILE_BYTES);
    const uint32_t id128 = umma::instr_desc_f16(FROWS, FMAXV);
    for (int k = 0; k < 4; ++k) umma::mma_f16_ts(gb + C_ACC, gb + C_XHI + 8u * k, dh + 2 * k, id128, k > 0);
    for (int k = 0; k < 4; ++k) umma::mma_f16_ts(gb + C_ACC, gb + C_XLO + 8u * k, dh + 2 * k, id128, 1u);
  }
  umma::mma_commit(group_bar);
}

// FAST != 0 = the configurations every acquisition round runs (the reference's default dropout sites E + layer-out of a
// BayesFormer, or none for the standard Transformer; Philox masks, greedy decoding, no log-prob / forced-token side
// channels): the latent-site, mask-injection, sampling and side-channel branches are compiled out instead of tested
// at run time.
template <int FAST>   // 0: everything at run time; 1: sites E + layer-out (BayesFormer); 2: no site (Transformer);
                      // 3: no site, temperature sampling (Transformer, sampling_uncertainty)
__global__ void __launch_bounds__(FTHREADS, 1) mc_forward_kernel(const FusedArgs a_in) {
  FusedArgs a = a_in;
  if (FAST) {
    a.site_flags = FAST == 1 ? (BFB200_SITE_EMBED | BFB200_SITE_LAYER_OUT) : 0u;
    a.scheme = FAST == 3 ? BFB200_SCHEME_SAMPLING : a.scheme;
    a.ms.bits = nullptr;
    a.logp_out = nullptr;
    a.forced = nullptr;
  }
  // Everything lives in dynamic shared memory (no static allocation in front of it, so its base is 1024-byte aligned
  // in practice; the pad is computed as an offset so that the pointers stay in the shared address space):
  //   [FGROUPS x FTILES tiles of 16 KB][FGROUPS V-remainder tiles of 8 KB][weight image][mbarriers, TMEM slot]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const FusedLayout lay = a.lay;
  if (pad + kTilesBytes + lay.smem_image_bytes + kBarrierBytes > a.smem_bytes) __trap();
  uint8_t* img = smem + kTilesBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(img + lay.smem_image_bytes);
  uint64_t* bar_image = bars;          // weight image landed
  uint64_t* bar_group = bars + 1;      // [FGROUPS] product of the group complete
  uint64_t* bar_wlo = bars + 3;        // [FGROUPS] W_qkv lo rows of the next QKV product landed in the V tiles
  uint64_t* bar_head = bars + 5;       // [FGROUPS] vocabulary head landed in the Q tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x;
  // group and feature half are warp-uniform; taking them from votes lets the compiler know (uniform datapath for the
  // address arithmetic and the branches they steer)
  const int g = __any_sync(0xffffffffu, tid >= FGROUP_THREADS) ? 1 : 0, gt = tid % FGROUP_THREADS;   // group, thread in group
  const int h = __any_sync(0xffffffffu, (gt >> 7) != 0) ? 1 : 0, row = gt & 127;                     // feature half, token row
  const int gw = row >> 5;                                         // TMEM lane quarter (= CTA warp % 4)
  uint8_t* TQ = smem + (uint32_t)(g * FTILES + T_Q) * TILE_BYTES;
  uint8_t* TK = smem + (uint32_t)(g * FTILES + T_K) * TILE_BYTES;
  uint8_t* TL = smem + (uint32_t)(g * FTILES + T_L) * TILE_BYTES;
  uint8_t* TV = smem + (uint32_t)(g * FTILES + T_V) * TILE_BYTES;
  const float* params = reinterpret_cast<const float*>(img + lay.off_params);
  float2* red = reinterpret_cast<float2*>(TK);   // LayerNorm exchange: K is dead whenever it is in use

  if (tid < 32) {
    umma::tmem_alloc(tmem_slot, 512);
    umma::tmem_relinquish();
  }
  if (tid == 0) {
    for (int i = 0; i < 7; ++i) umma::mbar_init(&bars[i], 1);
    umma::fence_mbar_init();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  if (tid == 0) {
    umma::mbar_expect_tx(bar_image, lay.smem_image_bytes);
    for (uint32_t off = 0; off < lay.smem_image_bytes; off += 16384u) {
      const uint32_t n = lay.smem_image_bytes - off < 16384u ? lay.smem_image_bytes - off : 16384u;
      umma::bulk_g2s(img + off, a.image + off, n, bar_image);
    }
  }
  const int64_t tile_first = (int64_t)blockIdx.x * FGROUPS + g, tile_step = (int64_t)gridDim.x * FGROUPS;
  const uint8_t* wlo_src = a.image + lay.off_wlo;
  auto stream_wlo = [&](int layer) {   // one thread: W_qkv lo rows of `layer` -> V tile (Q, K rows) and V-remainder tile (V rows)
    // (the lo rows of W_v are packed in the image too but not used: with V kept in fp32 the W_v remainder moves the
    //  worst probability error by < 0.1 of the tolerance — tools/fused_accuracy.py — and costs 1.9 % of the pass)
    umma::mbar_expect_tx(&bar_wlo[g], WLO_QK_BYTES);
    umma::bulk_g2s(TV, wlo_src + (uint32_t)layer * WLO_BYTES, WLO_QK_BYTES, &bar_wlo[g]);
  };
  if (gt == 0 && tile_first < a.n_tiles) stream_wlo(0);   // for the first QKV product
  umma::mbar_wait(bar_image, 0);

  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(gw * 32) << 16;
  const uint32_t tG = tmem_base + (uint32_t)g * 256u + lane_base;
  const uint32_t tXh = tG + C_XHI + 16u * (uint32_t)h, tXl = tG + C_XLO + 16u * (uint32_t)h;   // residual stream (this half)
  const uint32_t tA = tG + C_ACC;                                                              // accumulators, 192 columns
  if (tmem_base != 0u) __trap();   // all 512 columns are ours: the MMA issue path relies on base column 0
  uint64_t* bar = &bar_group[g];
  uint32_t phase = 0, wlo_phase = 0, head_phase = 0;
  const int bar_id = 1 + g;
  // LayerNorm exchanges only couple the two threads of a row (warps w and w + 4 of the group): a 64-thread named barrier
  // per warp pair instead of the group barrier (ids 3..10)
  const int pair_bar_id = 3 + 4 * g + gw;

  const int len = a.len, spt = a.spt;
  // rows are position-major inside a tile (row = pos * spt + sequence): the rows of the LAST position, the only
  // ones the vocabulary head consumes, are then contiguous, so that in the last layer whole warps can skip the
  // work whose results nothing reads (see `live` below)
  const int pos = row / spt, seq_l = row - pos * spt;
  const int last_lo = (len - 1) * spt, last_hi = len * spt - 1;              // rows of the last position
  // does this warp hold a row of the last position?  A vote rather than range arithmetic on the warp's rows: the compiler
  // then KNOWS the flag is warp-uniform, and the branches it steers (which contain MMA issue code) stay convergent
  const bool warp_has_last = __any_sync(0xffffffffu, row >= last_lo && row <= last_hi);
  const float sqrt_d = sqrtf((float)FD);
  const float qk_scale_log2e = 1.4426950408889634f / sqrtf((float)FDH);   // scores / sqrt(dh), in base-2 exponent units
  const uint32_t fl = a.site_flags;
  const bool merged = lay.merged != 0;   // the same rule fused_pack_kernel laid the image out by
  const MaskSource& ms = a.ms;

  IssueCtx ic;
  ic.tiles_u32 = umma::smem_u32(smem);
  ic.img_u32 = umma::smem_u32(img);
  ic.fp = lay.fp; ic.merged = lay.merged; ic.w2_per_tile = lay.w2_per_tile; ic.w2_lo = lay.w2_lo;
  ic.attn_lo = 1u;
  ic.off_om = lay.off_om; ic.off_w1 = lay.off_w1; ic.off_w2 = lay.off_w2;
  ic.layer_bytes = lay.layer_bytes;

#ifdef BFB_FUSED_TIMING
  long long dbg[3] = {0, 0, 0};   // observer lanes: round-trip cycles, round trips
  const long long t_begin = clock64();
  long long ph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, ph_last = t_begin;
#endif
  // One product: operands are in TMEM (tcgen05.st of this group's threads) or shared memory -> accumulators in TMEM.
  // next_wlo >= 0: the V tiles are dead from here on, so the issuing thread starts the bulk copies of that layer's
  // W_qkv lo rows into them (consumed by the next PR_QKV product, which waits for them).  stage_head: likewise the Q
  // tile is dead and receives the vocabulary head's weights (consumed by PR_HEAD).
  // `overlap` runs between the issue and the wait: work that does not depend on the product (dropout draws).
  auto product_ov = [&](auto kind_c, int layer, int next_wlo, bool stage_head, auto&& overlap) {
    constexpr int kind = decltype(kind_c)::value;
#ifdef BFB_FUSED_TIMING
    const long long _t0 = clock64();
#endif
    if (kind == PR_MERGED || kind == PR_WOUT) umma::fence_async_smem();
    __syncwarp();   // the pair phase leaves the lanes of a warp at different trip counts; tcgen05.wait is warp-aligned
    umma::tmem_st_wait();
    umma::tc_fence_before_sync();
    umma::group_sync(bar_id, FGROUP_THREADS);
    if (((tid >> 5) & 7) == 0) {   // warp 0 of the group (warp-uniform), then one lane: keeps the operand arithmetic on the uniform datapath
     if ((tid & 31) == 0) {
      umma::tc_fence_after_sync();
      if (next_wlo >= 0) stream_wlo(next_wlo);
      if (stage_head) {
        umma::mbar_expect_tx(&bar_head[g], HEAD_BYTES);
        umma::bulk_g2s(TQ, a.image + lay.off_head, HEAD_BYTES, &bar_head[g]);
      }
      if (kind == PR_QKV) umma::mbar_wait(&bar_wlo[g], wlo_phase);
      if (kind == PR_HEAD) umma::mbar_wait(&bar_head[g], head_phase);
      if (g == 0) fused_issue<0, kind>(ic, layer, &bar_group[0]);
      else        fused_issue<1, kind>(ic, layer, &bar_group[1]);
     }
    }
    if (kind == PR_QKV) wlo_phase ^= 1u;
    if (kind == PR_HEAD) head_phase ^= 1u;
    umma::mbar_wait(bar, phase ^ (overlap() & a.opaque_zero));
    phase ^= 1u;
    umma::tc_fence_after_sync();
#ifdef BFB_FUSED_TIMING
    if ((gt & 31) == 5) { dbg[0] += clock64() - _t0; dbg[1] += 1; }
#endif
  };
  auto product = [&](auto kind_c, int layer, int next_wlo, bool stage_head = false) {
    product_ov(kind_c, layer, next_wlo, stage_head, [] { return 0u; });
  };

  for (int64_t tile = tile_first; tile < a.n_tiles; tile += tile_step) {
    const uint32_t s0 = (uint32_t)tile * (uint32_t)spt;   // n_seq < 2^31 per launch (checked on the host)
    const int nseq = (int)(((uint32_t)a.n_seq - s0) < (uint32_t)spt ? ((uint32_t)a.n_seq - s0) : (uint32_t)spt);
    const bool row_valid = pos < len && seq_l < nseq;
    const uint32_t s = s0 + (row_valid ? (uint32_t)seq_l : 0u);
    const RowRng rr = make_row_rng(ms, s);
    const bool more_tiles = tile + tile_step < a.n_tiles;

    // ---- embedding + dropout E + sqrt(d) scale + positional encoding (transformer.py:355-362) ----
    {
      const uint32_t k_embed = (!FAST && (fl & BFB200_SITE_EMBED)) ? half_keep_mask(ms, rr, (uint32_t)pos, 0u, (uint32_t)h) : 0u;
      float x[HW];
      if (row_valid) {
        const int tok = a.tokens[(size_t)s * a.tok_stride + pos];
        const float4* e = reinterpret_cast<const float4*>(a.embedding + (size_t)tok * FD + h * HW);
        const float4* pe = reinterpret_cast<const float4*>(a.pe + (size_t)pos * FD + h * HW);
        float pv[HW];
#pragma unroll
        for (int i = 0; i < HW / 4; ++i) {
          const float4 v = __ldg(e + i), q = __ldg(pe + i);
          x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
          pv[4 * i] = q.x; pv[4 * i + 1] = q.y; pv[4 * i + 2] = q.z; pv[4 * i + 3] = q.w;
        }
        if (FAST == 1) drop_half_direct(x, ms, rr, (uint32_t)pos, 0u, (uint32_t)h);
        else if (fl & BFB200_SITE_EMBED) drop_half(x, k_embed, ms.scale);
#pragma unroll
        for (int i = 0; i < HW; ++i) x[i] = fmaf(x[i], sqrt_d, pv[i]);
      } else {
#pragma unroll
        for (int i = 0; i < HW; ++i) x[i] = 0.f;
      }
      store_hilo_half(tXh, tXl, x);
    }

    BFB_PHASE(0);
    for (int l = 0; l < a.n_layers; ++l) {
      const float* pl = params + (uint32_t)l * lay.params_per_layer;
      const uint32_t sb = 1u + BFB200_SITES_PER_LAYER * (uint32_t)l;
      // Last layer: only the last position's output is read (vocabulary head, active_learning.py:146), so its
      // queries, W_out / LayerNorm / FFN rows and layer-out masks for the other positions are dead work; K and V of
      // every position are still needed.  `live` is warp-uniform.
      const bool prune = l + 1 == a.n_layers;
      const bool live = !prune || warp_has_last;
      const int next_wlo = !prune ? l + 1 : (more_tiles ? 0 : -1);

      // ---- Q, K, V = x W^T for all heads (transformer.py:54-57): one N = 192 product ----
      product(KindC<PR_QKV>{}, l, -1);
      BFB_PHASE(1);
      // ---- causal attention (transformer.py:58-65), in fp32, FOUR HEADS AT A TIME ----
      // The accumulator columns [0,64) = Q, [64,128) = K, [128,192) = V stay in tensor memory; per pass hp the fp32
      // values of heads [4 hp, 4 hp + 4) are copied into the Q / K / V tiles (128 rows x 4 heads x 8 floats = 16 KB each:
      // the same three tiles hold either half) and the pairs of those heads are worked off with plain fp32 FMAs — no
      // operand rounding, no remainders to carry, a third of the instructions of the split-fp16 form.  (row, head hh)
      // occupies the two 16-byte chunks 2 hh, 2 hh + 1 of the row in the usual 128-byte swizzle: the stores of eight
      // consecutive rows and the loads of a quarter-warp (two consecutive rows x four heads) are conflict-free.
      // The output of pass 0 goes straight into the L tile as the merged product's A operand (bytes [0,64) of a row:
      // fp16 hi of features 0..31, bytes [64,128): their lo parts); pass 1's output belongs in the Q tile, which other
      // items still read, so it waits in registers for the barrier.
      uint4 keep_hi[2], keep_lo[2];
      uint32_t keep_off[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll 1
      for (uint32_t hp = 0; hp < 2; ++hp) {
        BFB_PHASE(3);
        if (hp == 1) umma::group_sync(bar_id, FGROUP_THREADS);   // pass 0 has read its last K / V row
        BFB_PHASE(4);
        // this thread copies heads 2 h, 2 h + 1 of the pass (16 accumulator columns per matrix)
        {
          uint32_t rq[16], rk[16], rv[16];   // the three loads travel together: one tensor-memory latency per pass
          const uint32_t c0 = tA + 32u * hp + 16u * (uint32_t)h;
          umma::tmem_ld16(c0, rq);
          umma::tmem_ld16(c0 + 64u, rk);
          umma::tmem_ld16(c0 + 128u, rv);
          umma::tmem_ld_wait();
#pragma unroll
          for (uint32_t c = 0; c < 4; ++c) {
            const uint32_t off = umma::sw128_offset(row, 4u * (uint32_t)h + c);
            *reinterpret_cast<uint4*>(TQ + off) = make_uint4(rq[4 * c], rq[4 * c + 1], rq[4 * c + 2], rq[4 * c + 3]);
            *reinterpret_cast<uint4*>(TK + off) = make_uint4(rk[4 * c], rk[4 * c + 1], rk[4 * c + 2], rk[4 * c + 3]);
            *reinterpret_cast<uint4*>(TV + off) = make_uint4(rv[4 * c], rv[4 * c + 1], rv[4 * c + 2], rv[4 * c + 3]);
          }
        }
        umma::tc_fence_before_sync();
        BFB_PHASE(2);
        umma::group_sync(bar_id, FGROUP_THREADS);
        BFB_PHASE(4);
        // Work item = (query position, sequence, head of the pass), query-major (the lanes of a warp walk the same
        // number of keys), heaviest queries first; at most 128 x 4 items, two per thread: item gt and, from the other
        // end of the list, item 511 - gt, so that every thread visits about the same number of keys.  ONE pass over the
        // keys: the softmax is taken relative to the first key's score (it is shift-invariant); in the rare case that a
        // score exceeds the reference by more than 14 octaves the accumulators are rescaled and the reference moves.
        const int npairs = nseq * 4;
        const int nq = prune ? 1 : len;                      // last layer: only the last query is live
        const int nitems = npairs * nq;
        const float inv_npairs = __fdividef(1.0f, (float)npairs);
#pragma unroll 1
        for (int u = 0; u < 2; ++u) {
          const int it = u == 0 ? gt : 2 * FGROUP_THREADS - 1 - gt;
          uint32_t off_out = 0xFFFFFFFFu;
          uint4 w = make_uint4(0u, 0u, 0u, 0u), wl = w;
          if (it < nitems) {
            const int b = (int)(((float)it + 0.5f) * inv_npairs);  // query block = it / npairs (exact: it < 512, npairs <= 128)
            const int p = it - b * npairs;
            const int t = len - 1 - b;
            const int sl = p >> 2, hd = p & 3;
            const uint32_t qrow = (uint32_t)(sl + t * spt);
            const uint32_t qoff = umma::sw128_offset(qrow, 2u * (uint32_t)hd);
            const float4 q0 = *reinterpret_cast<const float4*>(TQ + qoff);
            const float4 q1 = *reinterpret_cast<const float4*>(TQ + (qoff ^ 16u));
            float ref = 0.f, z = 0.f;
            float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            uint32_t krow = (uint32_t)sl;
#pragma unroll kKeyUnroll
            for (int j = 0; j <= t; ++j, krow += (uint32_t)spt) {
              const uint32_t koff = umma::sw128_offset(krow, 2u * (uint32_t)hd);
              const float4 k0 = *reinterpret_cast<const float4*>(TK + koff);
              const float4 k1 = *reinterpret_cast<const float4*>(TK + (koff ^ 16u));
              // two explicit FMA chains (intrinsics are never re-contracted: every instantiation of the kernel rounds the
              // score the same way, which the FAST-vs-generic bit-equality test relies on)
              const float sa = __fmaf_rn(q0.w, k0.w, __fmaf_rn(q0.z, k0.z, __fmaf_rn(q0.y, k0.y, __fmul_rn(q0.x, k0.x))));
              const float sb2 = __fmaf_rn(q1.w, k1.w, __fmaf_rn(q1.z, k1.z, __fmaf_rn(q1.y, k1.y, __fmul_rn(q1.x, k1.x))));
              const float sj = fminf(fmaxf(__fadd_rn(sa, sb2), -3.0e38f), 3.0e38f);
              ref = j == 0 ? sj : ref;
              float arg = (sj - ref) * qk_scale_log2e;
              if (arg > 14.0f) {   // rare
                const float f = ex2_approx(-arg);
                z *= f;
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] *= f;
                ref = sj;
                arg = 0.f;
              }
              const float e = ex2_approx(arg);
              z += e;
              const float4 v0 = *reinterpret_cast<const float4*>(TV + koff);
              const float4 v1 = *reinterpret_cast<const float4*>(TV + (koff ^ 16u));
              o[0] = fmaf(e, v0.x, o[0]); o[1] = fmaf(e, v0.y, o[1]); o[2] = fmaf(e, v0.z, o[2]); o[3] = fmaf(e, v0.w, o[3]);
              o[4] = fmaf(e, v1.x, o[4]); o[5] = fmaf(e, v1.y, o[5]); o[6] = fmaf(e, v1.z, o[6]); o[7] = fmaf(e, v1.w, o[7]);
            }
            const float inv_z = __fdividef(1.0f, z);
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] *= inv_z;
            if (fl & BFB200_SITE_HEAD_OUT) {   // transformer.py:116-122
              const RowRng prr = make_row_rng(ms, s0 + (uint32_t)sl);
              const uint32_t k = (half_keep_mask(ms, prr, (uint32_t)t, sb + 4u, hp) >> (8 * hd)) & 0xFFu;
#pragma unroll
              for (int c = 0; c < 8; ++c) o[c] = ((k >> c) & 1u) ? o[c] * ms.scale : 0.f;
            }
            umma::split_f16x2(o[0], o[1], w.x, wl.x); umma::split_f16x2(o[2], o[3], w.y, wl.y);
            umma::split_f16x2(o[4], o[5], w.z, wl.z); umma::split_f16x2(o[6], o[7], w.w, wl.w);
            off_out = umma::sw128_offset(qrow, (uint32_t)hd);   // hi chunk; the lo chunk is four chunks (64 bytes) further
            if (hp == 0) {
              *reinterpret_cast<uint4*>(TL + off_out) = w;
              *reinterpret_cast<uint4*>(TL + (off_out ^ 64u)) = wl;
            }
          }
          if (hp == 1) {
#pragma unroll
            for (int k = 0; k < 2; ++k)   // (register select: keeps the arrays out of local memory)
              if (u == k) { keep_hi[k] = w; keep_lo[k] = wl; keep_off[k] = off_out; }
          }
        }
      }
      BFB_PHASE(3);
      umma::group_sync(bar_id, FGROUP_THREADS);   // every Q / K / V row of pass 1 has been read
      BFB_PHASE(4);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (keep_off[k] != 0xFFFFFFFFu) {
          *reinterpret_cast<uint4*>(TQ + keep_off[k]) = keep_hi[k];
          *reinterpret_cast<uint4*>(TQ + (keep_off[k] ^ 64u)) = keep_lo[k];
        }

      if (!live) {
        // dead rows of the last layer: take part in the barriers of the products (thread 0 of the group still issues
        // the MMAs inside product()), compute nothing; the LayerNorm exchanges are per warp pair and both warps of a
        // pair are dead together
        if (merged) {
          product(KindC<PR_MERGED>{}, l, next_wlo);
        } else {
          product(KindC<PR_WOUT>{}, l, next_wlo);
          product(KindC<PR_FFN1>{}, l, -1);
        }
        product(KindC<PR_FFN2>{}, l, -1, prune);
      } else {
      // (no `continue` above: a divergent loop-continue would make the layer counter look thread-dependent to the
      // compiler, and every MMA operand derived from it would leave the uniform datapath)

      const uint32_t t_h = tA + (merged ? 128u : 64u);   // FFN hidden row as a TS operand: hi, lo at +32
      if (merged) {
        // ---- W_out and FFN1 in one product; residual from the block input; LayerNorm1 folded (transformer.py:125, :231-232) ----
        product(KindC<PR_MERGED>{}, l, next_wlo);
        BFB_PHASE(5);
        {
          float v[HW];
          tmem_load_half(tA + h * HW, v);
          add_residual_half(tXh, tXl, v);
          red[h * FROWS + row] = ln_partial(v);
        }
        umma::group_sync(pair_bar_id, 64);   // the two half-row threads of a row sit in warps w and w + 4
        {
          // W1 LN1(v) + b1 = rstd (v (W1 g1)^T - mean col_s) + col_c, ReLU (transformer.py:151-155); each half takes fp / 2 columns
          float ln_mean, ln_rstd;
          ln_combine(red[h * FROWS + row], red[(h ^ 1) * FROWS + row], a.ln_eps, ln_mean, ln_rstd);
          const float* col_s = pl + P_COLS;
          const float* col_c = pl + P_COLS + lay.fp;
          const uint32_t half_cols = lay.fp >> 1;   // multiple of 8
          for (uint32_t c0 = half_cols * h; c0 < half_cols * (h + 1); c0 += 8) {
            uint32_t r[8];
            umma::tmem_ld8(tA + FD + c0, r);
            umma::tmem_ld_wait();
            float hc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              hc[i] = fmaxf(fmaf(ln_rstd, fmaf(-ln_mean, col_s[c0 + i], __uint_as_float(r[i])), col_c[c0 + i]), 0.f);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) umma::split_f16x2(hc[2 * i], hc[2 * i + 1], hi[i], lo[i]);
            umma::tmem_st4(t_h + (c0 >> 1), hi);
            umma::tmem_st4(t_h + 32u + (c0 >> 1), lo);
          }
        }
      } else {
        // ---- W_out, residual from the block input, LayerNorm1 (transformer.py:125, :226, :231), then FFN1 ----
        product(KindC<PR_WOUT>{}, l, next_wlo);
        {
          const uint32_t k_attn = (fl & BFB200_SITE_ATTN_RESID) ? half_keep_mask(ms, rr, (uint32_t)pos, sb + 1u, (uint32_t)h) : 0u;
          const uint32_t k_ffn_in = (fl & BFB200_SITE_FFN_IN) ? half_keep_mask(ms, rr, (uint32_t)pos, sb + 2u, (uint32_t)h) : 0u;
          float v[HW];
          tmem_load_half(tA + h * HW, v);
          if (fl & BFB200_SITE_ATTN_RESID) drop_half(v, k_attn, ms.scale);
          add_residual_half(tXh, tXl, v);
          const float2 own = ln_partial(v);
          red[h * FROWS + row] = own;
          umma::group_sync(pair_bar_id, 64);   // the two half-row threads of a row sit in warps w and w + 4
          ln_finish(v, own, red[(h ^ 1) * FROWS + row], pl + P_LN1W + h * HW, pl + P_LN1B + h * HW, a.ln_eps);
          if (fl & BFB200_SITE_FFN_IN) drop_half(v, k_ffn_in, ms.scale);
          store_hilo_half(tA + 64u + 16u * (uint32_t)h, tA + 96u + 16u * (uint32_t)h, v);
        }
        product(KindC<PR_FFN1>{}, l, -1);
        {
          const float* b1 = pl + P_B1;
          const uint32_t half_cols = lay.fp >> 1;
          for (uint32_t c0 = half_cols * h; c0 < half_cols * (h + 1); c0 += 8) {
            uint32_t r[8];
            umma::tmem_ld8(tA + 128u + c0, r);
            umma::tmem_ld_wait();
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              umma::split_f16x2(fmaxf(__uint_as_float(r[2 * i]) + b1[c0 + 2 * i], 0.f),
                                fmaxf(__uint_as_float(r[2 * i + 1]) + b1[c0 + 2 * i + 1], 0.f), hi[i], lo[i]);
            umma::tmem_st4(t_h + (c0 >> 1), hi);
            umma::tmem_st4(t_h + 32u + (c0 >> 1), lo);
          }
        }
      }

      // ---- FFN: Linear(F -> d) + bias, residual from the block input, LayerNorm2, layer-out dropout ----
      BFB_PHASE(6);
      uint4 rnd_out[4];   // FAST == 1: the layer-out site's random words, drawn while the FFN2 product runs
      product_ov(KindC<PR_FFN2>{}, l, -1, prune,   // last layer: the Q tile is dead now and receives the vocabulary head
                 [&] {
                   if (FAST != 1) return 0u;
                   philox_half_raw(rnd_out, ms, rr, (uint32_t)pos, sb + 0u, (uint32_t)h);
                   return fold_raw(rnd_out);
                 });
      BFB_PHASE(7);
      {
        const uint32_t k_resid = (fl & BFB200_SITE_FFN_RESID) ? half_keep_mask(ms, rr, (uint32_t)pos, sb + 3u, (uint32_t)h) : 0u;
        const uint32_t k_out = (!FAST && (fl & BFB200_SITE_LAYER_OUT)) ? half_keep_mask(ms, rr, (uint32_t)pos, sb + 0u, (uint32_t)h) : 0u;
        float v[HW];
        tmem_load_half(tA + h * HW, v);
        const float* b2 = pl + P_B2 + h * HW;
#pragma unroll
        for (int i = 0; i < HW; ++i) v[i] += b2[i];
        if (fl & BFB200_SITE_FFN_RESID) drop_half(v, k_resid, ms.scale);
        add_residual_half(tXh, tXl, v);
        const float2 own = ln_partial(v);
        red[h * FROWS + row] = own;
        umma::group_sync(pair_bar_id, 64);   // the two half-row threads of a row sit in warps w and w + 4
        ln_finish(v, own, red[(h ^ 1) * FROWS + row], pl + P_LN2W + h * HW, pl + P_LN2B + h * HW, a.ln_eps);
        if (FAST == 1) drop_half_raw(v, rnd_out, ms);   // transformer.py:370
        else if (fl & BFB200_SITE_LAYER_OUT) drop_half(v, k_out, ms.scale);
        store_hilo_half(tXh, tXl, v);   // next block's input / the vocabulary head's operand
      }
      }   // live
      BFB_PHASE(8);
    }

    // ---- vocabulary head on every row (only the last position is consumed, active_learning.py:146) ----
    product(KindC<PR_HEAD>{}, 0, -1);
    BFB_PHASE(9);
    const int stage_stride = FMAXV + a.tps;
    float* stage = reinterpret_cast<float*>(TK);  // K + L tiles: 32 KB (dead between the last W_out product and the next QKV product)
    {
      const bool is_last = row_valid && pos == len - 1;
      const float4* hb = reinterpret_cast<const float4*>(a.image + lay.off_headb);   // classes >= V carry -inf
      for (uint32_t c0 = 64u * h; c0 < 64u * h + 64u; c0 += 16) {
        uint32_t r[16];
        umma::tmem_ld16(tA + c0, r);
        umma::tmem_ld_wait();
        if (is_last) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b = __ldg(hb + (c0 >> 2) + i);
            float* dstp = stage + seq_l * stage_stride + c0 + 4 * i;
            dstp[0] = __uint_as_float(r[4 * i]) + b.x; dstp[1] = __uint_as_float(r[4 * i + 1]) + b.y;
            dstp[2] = __uint_as_float(r[4 * i + 2]) + b.z; dstp[3] = __uint_as_float(r[4 * i + 3]) + b.w;
          }
        }
      }
    }
    umma::tc_fence_before_sync();
    umma::group_sync(bar_id, FGROUP_THREADS);

    BFB_PHASE(10);
    // ---- log-softmax, probabilities, score, next token (transformer.py:375; active_learning.py:146-152) ----
    if (a.tps == 16) score_phase<16, FAST == 1 || FAST == 2>(a, stage, stage_stride, gt, nseq, s0, len);
    else             score_phase<8, FAST == 1 || FAST == 2>(a, stage, stage_stride, gt, nseq, s0, len);
    // the staging rows are rewritten only after the next tile's first product: barriers in between order it
    BFB_PHASE(11);
  }

#ifdef BFB_FUSED_TIMING
  if ((gt & 31) == 5) {
    atomicAdd(&bfb_fused_dbg[0], (unsigned long long)dbg[0]);
    atomicAdd(&bfb_fused_dbg[1], (unsigned long long)dbg[1]);
    atomicAdd(&bfb_fused_dbg[3], (unsigned long long)(clock64() - t_begin));
    atomicAdd(&bfb_fused_dbg[4], 1ull);
    for (int i = 0; i < 12; ++i) atomicAdd(&bfb_fused_phase[i], (unsigned long long)ph[i]);
  }
#endif
  umma::tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem_base, 512);
}

}  // namespace

#ifdef BFB_FUSED_TIMING
extern "C" int bfb200_debug_fused_phases(unsigned long long* out12, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out12, bfb_fused_phase, sizeof(unsigned long long) * 12);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(bfb_fused_phase, z, sizeof(z));
  }
  return 0;
}
extern "C" int bfb200_debug_fused_timing(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, bfb_fused_dbg, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(bfb_fused_dbg, z, sizeof(z));
  }
  return 0;
}
#endif

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
bool fused_supported(const bfb200_transformer& m) {
  if (m.d_model != FD || m.n_heads != FH || m.n_layers < 1) return false;
  if (m.d_ff < 1 || m.d_ff > 64 || m.ntoken < 2 || m.ntoken > FMAXV) return false;
  if (m.site_flags & BFB200_SITE_HEAD_IN) return false;   // Head.bayes input dropout: fp32 path only
  // tiles + weight image + the worst-case alignment pad must fit the 227 KB a CTA can have
  const FusedLayout lay = fused_layout(m.n_layers, m.d_ff, m.site_flags);
  return fused_smem_bytes(lay) + 1024 <= kFusedSmemMax;
}

size_t fused_image_bytes(const bfb200_transformer& m) {
  return fused_supported(m) ? fused_layout(m.n_layers, m.d_ff, m.site_flags).total_bytes : 0;
}

int fused_pack(const bfb200_transformer& m, void* image, size_t bytes, cudaStream_t st) {
  BFB_REQUIRE(fused_supported(m), BFB200_ERR_UNSUPPORTED, "model shape is outside the fused kernel's class");
  const FusedLayout lay = fused_layout(m.n_layers, m.d_ff, m.site_flags);
  BFB_REQUIRE(bytes >= lay.total_bytes, BFB200_ERR_WORKSPACE, "fused image buffer too small: %zu < %u", bytes,
              lay.total_bytes);
  cudaError_t e = cudaMemsetAsync(image, 0, lay.total_bytes, st);
  BFB_REQUIRE(e == cudaSuccess, BFB200_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e));
  fused_pack_kernel<<<64, 256, 0, st>>>(m, lay, (uint8_t*)image);
  BFB_LAUNCH_CHECK("fused_pack_kernel", st);
  return BFB200_OK;
}

int launch_mc_forward_fused(const bfb200_transformer& m, const int32_t* tokens, int32_t* tokens_out, int tok_stride,
                            int64_t n_seq, int len, int scheme, int mode, float temperature, const MaskSource& ms,
                            const int32_t* forced, float* seq_scores, float* logp_out, cudaStream_t st) {
  BFB_REQUIRE(m.fused_image != nullptr, BFB200_ERR_INVALID_ARG, "compute=f16 needs model.fused_image (bfb200_fused_pack)");
  BFB_REQUIRE(len >= 1 && len <= FMAXLEN, BFB200_ERR_UNSUPPORTED, "fused kernel takes sequences of 1..%d tokens (got %d)",
              FMAXLEN, len);
  BFB_REQUIRE(len <= m.pe_len, BFB200_ERR_INVALID_ARG, "sequence length %d exceeds the positional table (%d)", len, m.pe_len);
  BFB_REQUIRE(n_seq >= 0 && n_seq < ((int64_t)1 << 31), BFB200_ERR_UNSUPPORTED,
              "fused kernel takes fewer than 2^31 sequences per launch (got %lld): use a smaller workspace chunk", (long long)n_seq);
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.image = (const uint8_t*)m.fused_image;
  a.embedding = m.embedding;
  a.pe = m.pe;
  a.lay = fused_layout(m.n_layers, m.d_ff, m.site_flags);
  a.n_layers = m.n_layers; a.d_ff = m.d_ff; a.vocab = m.ntoken;
  a.site_flags = m.site_flags;
  a.ln_eps = m.ln_eps;
  a.len = len;
  a.spt = FROWS / len < 32 ? FROWS / len : 32;   // the score phase gives every sequence >= 8 threads
  a.tps = (FGROUP_THREADS / a.spt) >= 16 ? 16 : 8;
  a.n_seq = n_seq;
  a.n_tiles = (n_seq + a.spt - 1) / a.spt;
  a.tokens = tokens; a.tokens_out = tokens_out; a.tok_stride = tok_stride;
  a.scheme = scheme; a.mode = mode; a.temperature = temperature;
  a.forced = forced; a.seq_scores = seq_scores; a.logp_out = logp_out;
  a.ms = ms;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  size_t smem = fused_smem_bytes(a.lay) + 1024;
  if (smem > kFusedSmemMax) smem = kFusedSmemMax;
  a.smem_bytes = (uint32_t)smem;
  a.opaque_zero = 0u;
  int64_t grid = (a.n_tiles + FGROUPS - 1) / FGROUPS;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  const bool plain = a.ms.bits == nullptr && a.logp_out == nullptr && a.forced == nullptr;
  const bool sampling = scheme == BFB200_SCHEME_SAMPLING;
  const int fast = !plain ? 0
                   : (!sampling && a.site_flags == (uint32_t)(BFB200_SITE_EMBED | BFB200_SITE_LAYER_OUT)) ? 1
                   : (a.site_flags == 0u) ? (sampling ? 3 : 2) : 0;
#define FUSED_LAUNCH(FV)                                                                                \
  do {                                                                                                  \
    cudaFuncSetAttribute(mc_forward_kernel<FV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    mc_forward_kernel<FV><<<(int)grid, FTHREADS, smem, st>>>(a);                                        \
  } while (0)
  if (fast == 1) FUSED_LAUNCH(1);
  else if (fast == 2) FUSED_LAUNCH(2);
  else if (fast == 3) FUSED_LAUNCH(3);
  else FUSED_LAUNCH(0);
#undef FUSED_LAUNCH
  BFB_LAUNCH_CHECK("mc_forward_kernel", st);
  return BFB200_OK;
}

}  // namespace bfb

// Fused MC-forward kernel for the notebook-default model class (d_model = 64, 8 heads of 8, F <= 64,
// V <= 128, sequences of at most 40 tokens): ONE kernel runs the whole forward pass of a decode step —
// embedding + MC dropout + positional encoding, every TransformerBlock (QKV / W_out / FFN contractions on
// tcgen05 tensor cores with fp16 operands and fp32 TMEM accumulators, causal attention, both
// residual + LayerNorm pairs, the always-on dropout sites), the last-position vocabulary head, log-softmax,
// the per-draw uncertainty score and the next token — with every activation resident on chip.
//
// Layout.  A CTA holds the complete fp16 weight image (pre-swizzled by fused_pack_kernel, fetched with 1-D
// bulk TMA copies) in shared memory and runs two independent 128-thread groups.  A group owns a tile of
// floor(128 / len) whole sequences = up to 128 token rows:
//   * row phases    — thread = token row: reads its fp32 accumulator row from TMEM (tcgen05.ld 32x32b), does
//                     bias / ReLU / residual / LayerNorm / dropout in registers, writes the next operand row as
//                     fp16 into the 128-byte-swizzled K-major tile the next tcgen05.mma consumes;
//   * pair phase    — thread = (sequence, head) (or (query, sequence, head) beyond 16 tokens): causal softmax(QK^T / sqrt(dh)) V from the
//                     fp16 Q/K/V tiles (a head's 8 values are exactly one 16-byte swizzle chunk);
//   * score phase   — 4 or 8 threads per sequence: log-softmax over the vocabulary, probabilities, max /
//                     margin / entropy, greedy arg-max or inverse-CDF sample.
// The residual stream x stays in fp32 in 64 TMEM columns of the group.  HBM traffic per sequence and step:
// len token ids in, one score and one token out.
//
// Reference arithmetic: src/model/transformer.py:349-375 (forward), :211-234 (block), :37-65 (head),
// :151-170 (FFN); src/libs/active_learning.py:62-68 / :103-112 / :146-152 (per-step scoring).
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
constexpr int FMAXLEN = 40;     // longest sequence the fused path takes (rows of the positional table in the image)
constexpr int FMAXV = 128;
constexpr uint32_t TILE_BYTES = FROWS * 128;  // one fp16 operand tile (128 rows x 64 halves)
constexpr int PARAMS_PER_LAYER = 8 * 64;      // ln1_w, ln1_b, b1, b2, ln2_w, ln2_b, col_s, col_c (floats)

struct FusedLayout {
  uint32_t fp;            // F rounded up to 16
  uint32_t vp;            // V rounded up to 16
  uint32_t layer_bytes;   // matrices of one layer
  uint32_t off_out, off_w1, off_w2;  // inside a layer (qkv at 0)
  uint32_t off_head;      // absolute
  uint32_t off_params;    // absolute
  uint32_t image_bytes;   // multiple of 1024
};

__host__ __device__ inline FusedLayout fused_layout(int L, int F, int V) {
  FusedLayout lay;
  lay.fp = (uint32_t)((F + 15) / 16 * 16);
  lay.vp = FMAXV;   // the head product always spans 128 classes; classes >= V carry zero weights and bias -inf
  lay.off_out = 3 * FD * 128;
  lay.off_w1 = lay.off_out + FD * 128;
  lay.off_w2 = lay.off_w1 + lay.fp * 128;
  lay.layer_bytes = lay.off_w2 + FD * 128;
  lay.off_head = (uint32_t)L * lay.layer_bytes;
  lay.off_params = lay.off_head + lay.vp * 128;
  const uint32_t n_param_floats = (uint32_t)L * PARAMS_PER_LAYER + FMAXV + FMAXLEN * FD;
  lay.image_bytes = (lay.off_params + n_param_floats * 4 + 1023u) & ~1023u;
  return lay;
}

constexpr size_t kFusedStaticSlack = 4608;  // static __shared__ of the kernel (LayerNorm exchange 4 KB, barriers)

inline size_t fused_smem_bytes(const FusedLayout& lay) {
  return 1024 + (size_t)lay.image_bytes + (size_t)FGROUPS * 3 * TILE_BYTES;
}

// ------------------------------------------------------------------------------------------
// weight image: fp16 matrices in the swizzled K-major layout + fp32 vectors
// ------------------------------------------------------------------------------------------
__global__ void fused_pack_kernel(bfb200_transformer m, FusedLayout lay, uint8_t* __restrict__ image) {
  const int L = m.n_layers, F = m.d_ff, V = m.ntoken;
  const int per_layer = 3 * FD * FD + FD * FD + F * FD + FD * F;
  // LayerNorm1 is folded into the FFN1 product unless a dropout site sits between them (mc_forward_kernel): the W1
  // slot then holds W1 * diag(gamma1)
  const bool fold_ln1 = !(m.site_flags & BFB200_SITE_FFN_IN);
  const int n_mat = L * per_layer + V * FD;
  const int n_par = L * PARAMS_PER_LAYER + FMAXV + FMAXLEN * FD;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_mat + n_par; i += gridDim.x * blockDim.x) {
    if (i < n_mat) {
      float w;
      uint32_t base, n, k;
      if (i < L * per_layer) {
        const int l = i / per_layer;
        int j = i - l * per_layer;
        base = (uint32_t)l * lay.layer_bytes;
        if (j < 3 * FD * FD) {                       // W_qkv (3d, d)
          n = j / FD; k = j % FD; w = m.w_qkv[(size_t)l * 3 * FD * FD + j];
        } else if ((j -= 3 * FD * FD) < FD * FD) {   // W_out (d, d)
          n = j / FD; k = j % FD; w = m.w_out[(size_t)l * FD * FD + j]; base += lay.off_out;
        } else if ((j -= FD * FD) < F * FD) {        // FFN W1 (F, d)
          n = j / FD; k = j % FD; w = m.ff1_w[(size_t)l * F * FD + j]; base += lay.off_w1;
          if (fold_ln1) w *= m.ln1_w[l * FD + k];
        } else {                                     // FFN W2 (d, F): K = F
          j -= F * FD;
          n = j / F; k = j % F; w = m.ff2_w[(size_t)l * FD * F + j]; base += lay.off_w2;
        }
      } else {                                       // vocabulary head (V, d)
        const int j = i - L * per_layer;
        n = j / FD; k = j % FD; w = m.head_w[j]; base = lay.off_head;
      }
      const uint32_t off = base + umma::sw128_offset(n, k >> 3) + (k & 7u) * 2u;
      *reinterpret_cast<__half*>(image + off) = __float2half_rn(fminf(fmaxf(w, -65504.f), 65504.f));
    } else {
      const int j = i - n_mat;
      float v = 0.f;
      if (j < L * PARAMS_PER_LAYER) {
        const int l = j / PARAMS_PER_LAYER, q = j % PARAMS_PER_LAYER, which = q >> 6, c = q & 63;
        switch (which) {
          case 0: v = m.ln1_w[l * FD + c]; break;
          case 1: v = m.ln1_b[l * FD + c]; break;
          case 2: v = c < F ? m.ff1_b[l * F + c] : 0.f; break;
          case 3: v = m.ff2_b[l * FD + c]; break;
          case 4: v = m.ln2_w[l * FD + c]; break;
          case 5: v = m.ln2_b[l * FD + c]; break;
          case 6:    // col_s[n] = sum_k fp16(W1[n, k] gamma1[k]) — of the rounded values the tensor core multiplies
            if (c < F)
              for (int k = 0; k < FD; ++k)
                v += __half2float(__float2half_rn(fminf(fmaxf(m.ff1_w[((size_t)l * F + c) * FD + k] * m.ln1_w[l * FD + k], -65504.f), 65504.f)));
            break;
          default:   // col_c[n] = b1[n] + sum_k W1[n, k] beta1[k]
            if (c < F) {
              v = m.ff1_b[l * F + c];
              for (int k = 0; k < FD; ++k) v = fmaf(m.ff1_w[((size_t)l * F + c) * FD + k], m.ln1_b[l * FD + k], v);
            }
            break;
        }
      } else if (j < L * PARAMS_PER_LAYER + FMAXV) {
        const int c = j - L * PARAMS_PER_LAYER;
        v = c < V ? m.head_b[c] : -INFINITY;   // padded classes vanish from every softmax
      } else {
        const int q = j - L * PARAMS_PER_LAYER - FMAXV;  // pe[pos][c], pos < FMAXLEN
        const int pos = q >> 6;
        v = pos < m.pe_len ? m.pe[(size_t)pos * FD + (q & 63)] : 0.f;
      }
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
  FusedLayout lay;
  int n_layers, d_ff, vocab;
  uint32_t site_flags;
  float ln_eps;
  // this launch
  int len;                     // tokens per sequence in this step
  int spt;                     // sequences per tile = 128 / len
  int tps;                     // score-phase threads per sequence (4 or 8)
  uint32_t qsplit;             // pair phase with two threads per (sequence, head): bit t = which of them takes query t
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
// encoded in (c2, c3): four calls of eight 16-bit decisions each.  Deliberately NOT inlined: the seven dropout
// sites of the kernel share this one copy of the rounds (the kernel is instruction-fetch sensitive); arguments are
// scalars so nothing spills to local memory at the call; the four counters advance together so the rounds of
// independent streams interleave.
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

__device__ __forceinline__ float dot8_f16(const uint4& a, const uint4& b) {
  float d = fh_ll(a.x, b.x, 0.f);
  d = fh_hh(a.x, b.x, d);
  d = fh_ll(a.y, b.y, d);
  d = fh_hh(a.y, b.y, d);
  d = fh_ll(a.z, b.z, d);
  d = fh_hh(a.z, b.z, d);
  d = fh_ll(a.w, b.w, d);
  return fh_hh(a.w, b.w, d);
}
// o[0..8) += p * v[0..8), p = half `HI` of pp
template <bool HI>
__device__ __forceinline__ void axpy8_f16(uint32_t pp, const uint4& v, float (&o)[8]) {
  if (HI) {
    o[0] = fh_hl(pp, v.x, o[0]); o[1] = fh_hh(pp, v.x, o[1]); o[2] = fh_hl(pp, v.y, o[2]); o[3] = fh_hh(pp, v.y, o[3]);
    o[4] = fh_hl(pp, v.z, o[4]); o[5] = fh_hh(pp, v.z, o[5]); o[6] = fh_hl(pp, v.w, o[6]); o[7] = fh_hh(pp, v.w, o[7]);
  } else {
    o[0] = fh_ll(pp, v.x, o[0]); o[1] = fh_lh(pp, v.x, o[1]); o[2] = fh_ll(pp, v.y, o[2]); o[3] = fh_lh(pp, v.y, o[3]);
    o[4] = fh_ll(pp, v.z, o[4]); o[5] = fh_lh(pp, v.z, o[5]); o[6] = fh_ll(pp, v.w, o[6]); o[7] = fh_lh(pp, v.w, o[7]);
  }
}
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

// FAST instantiations (two dropout sites): the four Philox calls of a half row are inlined and decide straight on the
// 16-bit halves of their output words, eight features at a time — no keep-bit packing / unpacking in between
// (same decisions as philox_keep32 + drop_half: feature j of a call <- half j & 1 of word j >> 1, kept iff >= threshold).
__device__ __forceinline__ void drop_half_direct(float (&v)[HW], const MaskSource& m, const RowRng& rr, uint32_t pos,
                                                 uint32_t site, uint32_t h) {
  const uint32_t thr_hi = m.threshold << 16;
#pragma unroll
  for (uint32_t c = 0; c < 4; ++c) {
    const uint4 r = philox4x32_10(make_uint4(rr.item, rr.draw, (m.step << 16) | pos, (site << 16) | (4u * h + c)), m.key);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
      v[8 * c + 2 * i] = (w[i] << 16) >= thr_hi ? v[8 * c + 2 * i] * m.scale : 0.f;
      v[8 * c + 2 * i + 1] = w[i] >= thr_hi ? v[8 * c + 2 * i + 1] * m.scale : 0.f;
    }
  }
}

__device__ __forceinline__ void store_half_f16(uint8_t* tile, uint32_t row, uint32_t h, const float (&v)[HW]) {
#pragma unroll
  for (uint32_t c = 0; c < 4; ++c) {
    uint4 w;
    w.x = umma::pack_f16x2(v[8 * c + 0], v[8 * c + 1]);
    w.y = umma::pack_f16x2(v[8 * c + 2], v[8 * c + 3]);
    w.z = umma::pack_f16x2(v[8 * c + 4], v[8 * c + 5]);
    w.w = umma::pack_f16x2(v[8 * c + 6], v[8 * c + 7]);
    *reinterpret_cast<uint4*>(tile + umma::sw128_offset(row, 4 * h + c)) = w;
  }
}

__device__ __forceinline__ void tmem_load_half(uint32_t taddr, float (&v)[HW]) {
  uint32_t r[32];
  umma::tmem_ld32(taddr, r);
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < HW; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_store_half(uint32_t taddr, const float (&v)[HW]) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < HW; ++i) r[i] = __float_as_uint(v[i]);
  umma::tmem_st32(taddr, r);
  umma::tmem_st_wait();
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
__device__ __forceinline__ void ln_finish(float (&v)[HW], float2 own, float2 other, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float eps) {
  const float dm = own.x - other.x;
  const float mean = 0.5f * (own.x + other.x);
  const float var = (own.y + other.y + 16.0f * dm * dm) * (1.0f / FD);
  const float rstd = rsqrtf(var + eps);
  const float shift = -mean * rstd;
#pragma unroll
  for (int i = 0; i < HW; ++i) v[i] = fmaf(fmaf(v[i], rstd, shift), gamma[i], beta[i]);
}

#ifdef BFB_FUSED_TIMING
// debug build only (make EXTRA=-DBFB_FUSED_TIMING): cycles one observer thread per group spends in the product
// round trips (barrier -> MMA issue -> completion wait), in the other group barriers, and in total
__device__ unsigned long long bfb_fused_dbg[8];
#define BFB_T0() const long long _t0 = clock64()
#define BFB_T1(slot) do { if ((gt & 31) == 5) dbg[slot] += clock64() - _t0; } while (0)
#else
#define BFB_T0() do { } while (0)
#define BFB_T1(slot) do { } while (0)
#endif

// The MMAs of one product, issued by ONE thread of group G.  Every operand is built from kernel-uniform values (the
// shared-memory window base, kernel parameters, the compile-time group index; the TMEM allocation is the whole
// 512 columns, so its base is column 0) and therefore lives in uniform registers: with per-thread operands the
// compiler wraps each UTCHMMA in an ELECT / R2UR.BROADCAST / branch "waterfall" loop, ~140 cycles per instruction
// on the critical path of the whole group.
template <int G>
__device__ __forceinline__ void fused_issue_product(uint32_t tiles_u32, int a_sel, uint32_t b_addr, int N, int k_steps,
                                                    uint64_t* group_bars) {
  const uint32_t a_addr = tiles_u32 + (uint32_t)(G * 3 + a_sel) * TILE_BYTES;
  const uint64_t da = umma::smem_desc_sw128(a_addr), db = umma::smem_desc_sw128(b_addr);
  const uint32_t idesc = umma::instr_desc_f16(FROWS, N);
  const uint32_t d_tmem = (uint32_t)G * 256u + 64u;
  for (int k = 0; k < k_steps; ++k) umma::mma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
  umma::mma_commit(&group_bars[G]);
}

// FAST != 0 = the configurations every acquisition round runs (the reference's default dropout sites E + layer-out of a
// BayesFormer, or none for the standard Transformer; Philox masks, greedy decoding, no log-prob / forced-token side
// channels): the latent-site, mask-injection, sampling and side-channel branches are compiled out instead of tested
// at run time.
template <int MAXLEN, int FAST>   // 0: everything at run time; 1: sites E + layer-out (BayesFormer); 2: no site (Transformer);
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
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment of the swizzled tiles, computed as an offset so the pointers stay in the shared
  // address space (LDS / STS instead of generic loads)
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_image;
  __shared__ uint64_t bar_group[FGROUPS];
  __shared__ uint32_t tmem_slot;
  __shared__ float2 ln_red[FGROUPS][2 * FROWS];

  const int tid = threadIdx.x;
  const int g = tid / FGROUP_THREADS, gt = tid % FGROUP_THREADS;   // group, thread in group
  const int h = gt >> 7, row = gt & 127;                           // feature half, token row
  const int gw = row >> 5;                                         // TMEM lane quarter (= CTA warp % 4)
  const FusedLayout lay = a.lay;
  uint8_t* img = smem;
  uint8_t* B0 = smem + lay.image_bytes + (uint32_t)g * 3u * TILE_BYTES;
  uint8_t* B1 = B0 + TILE_BYTES;
  uint8_t* B2 = B1 + TILE_BYTES;
  const float* params = reinterpret_cast<const float*>(img + lay.off_params);
  float2* red = ln_red[g];

  if (tid < 32) {
    umma::tmem_alloc(&tmem_slot, 512);
    umma::tmem_relinquish();
  }
  if (tid == 0) {
    umma::mbar_init(&bar_image, 1);
    umma::mbar_init(&bar_group[0], 1);
    umma::mbar_init(&bar_group[1], 1);
    umma::fence_mbar_init();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  if (tid == 0) {
    umma::mbar_expect_tx(&bar_image, lay.image_bytes);
    for (uint32_t off = 0; off < lay.image_bytes; off += 16384u) {
      const uint32_t n = lay.image_bytes - off < 16384u ? lay.image_bytes - off : 16384u;
      umma::bulk_g2s(img + off, a.image + off, n, &bar_image);
    }
  }
  umma::mbar_wait(&bar_image, 0);

  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_base = (uint32_t)(gw * 32) << 16;
  const uint32_t tX = tmem_base + (uint32_t)g * 256u + lane_base + (uint32_t)h * HW;  // residual stream (this half)
  const uint32_t tW = tmem_base + (uint32_t)g * 256u + lane_base + 64u;               // accumulators, 192 columns
  if (tmem_base != 0u) __trap();   // all 512 columns are ours: the MMA issue path relies on base column 0
  uint64_t* bar = &bar_group[g];
  uint32_t phase = 0;
  const int bar_id = 1 + g;

  const int len = a.len, spt = a.spt;
  // rows are position-major inside a tile (row = pos * spt + sequence): the rows of the LAST position, the only
  // ones the vocabulary head consumes, are then contiguous, so that in the last layer whole warps can skip the
  // work whose results nothing reads (see `live` below)
  const int pos = row / spt, seq_l = row - pos * spt;
  const int last_lo = (len - 1) * spt, last_hi = len * spt - 1;              // rows of the last position
  const bool warp_has_last = (row & ~31) <= last_hi && (row | 31) >= last_lo;   // warp-uniform
  const float sqrt_d = sqrtf((float)FD);
  const float qk_scale_log2e = 1.4426950408889634f / sqrtf((float)FDH);   // scores / sqrt(dh), in base-2 exponent units
  const uint32_t fl = a.site_flags;
  const bool fold_ln1 = !(fl & BFB200_SITE_FFN_IN);   // the same rule fused_pack_kernel packed the W1 slot by
  const MaskSource& ms = a.ms;

  const uint32_t img_u32 = umma::smem_u32(img);

  // one GEMM phase: operand rows are in shared memory -> accumulators in TMEM columns [tW, tW + N)
#ifdef BFB_FUSED_TIMING
  long long dbg[4] = {0, 0, 0, 0};   // observer lanes: round-trip cycles, round trips, cycles after the barrier
  long long dbi[3] = {0, 0, 0};      // issuing thread: barrier exit -> commit issued, -> completion seen, fence
  const long long t_begin = clock64();
#endif
  const uint32_t tiles_u32 = umma::smem_u32(smem) + lay.image_bytes;   // kernel-uniform
  auto gemm = [&](int a_sel, uint32_t b_addr, int N, int k_steps) {
    BFB_T0();
    umma::fence_async_smem();
    umma::tc_fence_before_sync();
    umma::group_sync(bar_id, FGROUP_THREADS);
#ifdef BFB_FUSED_TIMING
    const long long _t_issue = clock64();
#endif
    if (gt == 0) {
      umma::tc_fence_after_sync();
#ifdef BFB_FUSED_TIMING
      const long long _t1 = clock64();
#endif
      if (g == 0) fused_issue_product<0>(tiles_u32, a_sel, b_addr, N, k_steps, bar_group);
      else        fused_issue_product<1>(tiles_u32, a_sel, b_addr, N, k_steps, bar_group);
#ifdef BFB_FUSED_TIMING
      const long long _t3 = clock64();
      dbi[0] += _t3 - _t_issue;
      dbi[2] += _t1 - _t_issue;
#endif
    }
    umma::mbar_wait(bar, phase);
#ifdef BFB_FUSED_TIMING
    if (gt == 0) dbi[1] += clock64() - _t_issue;
#endif
    phase ^= 1u;
    umma::tc_fence_after_sync();
    BFB_T1(0);
#ifdef BFB_FUSED_TIMING
    if ((gt & 31) == 5) { dbg[1] += 1; dbg[2] += clock64() - _t_issue; }
#endif
  };

  for (int64_t tile = (int64_t)blockIdx.x * FGROUPS + g; tile < a.n_tiles; tile += (int64_t)gridDim.x * FGROUPS) {
    const uint32_t s0 = (uint32_t)tile * (uint32_t)spt;   // n_seq < 2^31 per launch (checked on the host)
    const int nseq = (int)(((uint32_t)a.n_seq - s0) < (uint32_t)spt ? ((uint32_t)a.n_seq - s0) : (uint32_t)spt);
    const bool row_valid = pos < len && seq_l < nseq;
    const uint32_t s = s0 + (row_valid ? (uint32_t)seq_l : 0u);
    const RowRng rr = make_row_rng(ms, s);

    // ---- embedding + dropout E + sqrt(d) scale + positional encoding (transformer.py:355-362) ----
    {
      const uint32_t k_embed = (!FAST && (fl & BFB200_SITE_EMBED)) ? half_keep_mask(ms, rr, (uint32_t)pos, 0u, (uint32_t)h) : 0u;
      float x[HW];
      if (row_valid) {
        const int tok = a.tokens[(size_t)s * a.tok_stride + pos];
        const float4* e = reinterpret_cast<const float4*>(a.embedding + (size_t)tok * FD + h * HW);
#pragma unroll
        for (int i = 0; i < HW / 4; ++i) {
          const float4 v = __ldg(e + i);
          x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
        if (FAST == 1) drop_half_direct(x, ms, rr, (uint32_t)pos, 0u, (uint32_t)h);
        else if (fl & BFB200_SITE_EMBED) drop_half(x, k_embed, ms.scale);
        const float* pe = params + a.n_layers * PARAMS_PER_LAYER + FMAXV + pos * FD + h * HW;
#pragma unroll
        for (int i = 0; i < HW; ++i) x[i] = x[i] * sqrt_d + pe[i];
      } else {
#pragma unroll
        for (int i = 0; i < HW; ++i) x[i] = 0.f;
      }
      tmem_store_half(tX, x);
      store_half_f16(B0, row, h, x);
    }

    for (int l = 0; l < a.n_layers; ++l) {
      const uint32_t wl = img_u32 + (uint32_t)l * lay.layer_bytes;
      const float* pl = params + l * PARAMS_PER_LAYER;
      const uint32_t sb = 1u + BFB200_SITES_PER_LAYER * (uint32_t)l;
      // Last layer: only the last position's output is read (vocabulary head, active_learning.py:146), so its
      // queries, W_out / LayerNorm / FFN rows and layer-out masks for the other positions are dead work; K and V of
      // every position are still needed.  `live` is warp-uniform.
      const bool prune = l + 1 == a.n_layers;
      const bool live = !prune || warp_has_last;

      // ---- Q, K, V = x W^T for all heads (transformer.py:54-57): one N = 192 product ----
      gemm(0, wl, 3 * FD, 4);
      // accumulator columns [0,64) = Q, [64,128) = K, [128,192) = V -> fp16 operand tiles B0, B1, B2
      // (x in B0 is dead once the product has completed); each half copies 96 columns, 16 per trip
#pragma unroll 1
      for (uint32_t c0 = 96u * h; c0 < 96u * h + 96u; c0 += 16) {
        uint32_t r[16];
        umma::tmem_ld16(tW + c0, r);
        umma::tmem_ld_wait();
        uint8_t* dst = B0 + (c0 >> 6) * TILE_BYTES;   // B0, B1, B2 are contiguous
        const uint32_t ch = (c0 & 63u) >> 3;
        uint4 w0, w1;
        w0.x = umma::pack_f16x2(__uint_as_float(r[0]), __uint_as_float(r[1]));
        w0.y = umma::pack_f16x2(__uint_as_float(r[2]), __uint_as_float(r[3]));
        w0.z = umma::pack_f16x2(__uint_as_float(r[4]), __uint_as_float(r[5]));
        w0.w = umma::pack_f16x2(__uint_as_float(r[6]), __uint_as_float(r[7]));
        w1.x = umma::pack_f16x2(__uint_as_float(r[8]), __uint_as_float(r[9]));
        w1.y = umma::pack_f16x2(__uint_as_float(r[10]), __uint_as_float(r[11]));
        w1.z = umma::pack_f16x2(__uint_as_float(r[12]), __uint_as_float(r[13]));
        w1.w = umma::pack_f16x2(__uint_as_float(r[14]), __uint_as_float(r[15]));
        *reinterpret_cast<uint4*>(dst + umma::sw128_offset(row, ch)) = w0;
        *reinterpret_cast<uint4*>(dst + umma::sw128_offset(row, ch + 1)) = w1;
      }
      umma::tc_fence_before_sync();
      umma::group_sync(bar_id, FGROUP_THREADS);

      // ---- causal attention (transformer.py:58-65) ----
      if constexpr (MAXLEN > 16) {
        // Long sequences (the char-level configuration: 32..36 tokens): work item = (query position, sequence, head),
        // query-major so that the lanes of a warp walk (nearly) the same number of keys.  K and V rows are streamed
        // from the fp16 tiles, two passes over the keys (row maximum, then weights and the PV sum): no per-key state
        // survives in registers, whatever the length.
        const int npairs = nseq * FH;
        const int nitems = prune ? npairs : npairs * len;
        for (int it = gt; it < nitems; it += FGROUP_THREADS) {
          const int t = prune ? len - 1 : it / npairs;
          const int p = prune ? it : it - t * npairs;
          const int sl = p >> 3, hd = p & 7;
          const uint32_t row0 = (uint32_t)sl;
          uint8_t* qptr = B0 + umma::sw128_offset(row0 + t * spt, hd);
          const uint4 qw = *reinterpret_cast<const uint4*>(qptr);
          float mx = -INFINITY;
          for (int j = 0; j <= t; ++j)
            mx = fmaxf(mx, dot8_f16(qw, *reinterpret_cast<const uint4*>(B1 + umma::sw128_offset(row0 + j * spt, hd))));
          const float nmx = -mx * qk_scale_log2e;
          float z = 0.f;
          float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j <= t; ++j) {
            const float sj = dot8_f16(qw, *reinterpret_cast<const uint4*>(B1 + umma::sw128_offset(row0 + j * spt, hd)));
            const float e = ex2_approx(fmaf(sj, qk_scale_log2e, nmx));
            z += e;
            const uint32_t ph = umma::pack_f16x2(e, 0.f);                                  // e = hi + lo to 2^-22
            const uint32_t pl = umma::pack_f16x2(e - umma::unpack_f16x2(ph).x, 0.f);
            const uint4 vw = *reinterpret_cast<const uint4*>(B2 + umma::sw128_offset(row0 + j * spt, hd));
            axpy8_f16<false>(ph, vw, o);
            axpy8_f16<false>(pl, vw, o);
          }
          const float inv_z = __fdividef(1.0f, z);
#pragma unroll
          for (int c = 0; c < 8; ++c) o[c] *= inv_z;
          if (fl & BFB200_SITE_HEAD_OUT) {   // transformer.py:116-122
            const RowRng prr = make_row_rng(ms, s0 + (uint32_t)sl);
            const uint32_t k = (half_keep_mask(ms, prr, (uint32_t)t, sb + 4u, (uint32_t)(hd >> 2)) >> (8 * (hd & 3))) & 0xFFu;
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = ((k >> c) & 1u) ? o[c] * ms.scale : 0.f;
          }
          uint4 w;
          w.x = umma::pack_f16x2(o[0], o[1]); w.y = umma::pack_f16x2(o[2], o[3]);
          w.z = umma::pack_f16x2(o[4], o[5]); w.w = umma::pack_f16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(qptr) = w;   // in place: other items read only K / V rows, never this Q chunk
        }
      } else {
        // short sequences: work item = (sequence, head[, query parity]), K and V held in registers
        const int npairs = nseq * FH;
        // few pairs: two threads share one, queries split so that both walk about the same number of keys (query t
        // costs t + 1 key visits; a.qsplit is the balanced assignment); in the last layer only the last query is live
        const int nsplit = (!prune && npairs * 2 <= FGROUP_THREADS) ? 2 : 1;
        for (int it = gt; it < npairs * nsplit; it += FGROUP_THREADS) {
          const int p = nsplit == 2 ? it >> 1 : it;
          const int qpar = nsplit == 2 ? (it & 1) : -1;
          const int sl = p >> 3, hd = p & 7;
          const uint32_t row0 = (uint32_t)sl;   // row of (sequence sl, position j) = j * spt + sl
          // K and V stay packed (4 registers per key / value row); the dot products read the fp16 halves in place
          uint4 kp[MAXLEN], vp[MAXLEN];
#pragma unroll
          for (int j = 0; j < MAXLEN; ++j) {
            if (j < len) {
              kp[j] = *reinterpret_cast<const uint4*>(B1 + umma::sw128_offset(row0 + j * spt, hd));
              vp[j] = *reinterpret_cast<const uint4*>(B2 + umma::sw128_offset(row0 + j * spt, hd));
            }
          }
          const RowRng prr = make_row_rng(ms, s0 + (uint32_t)sl);
#pragma unroll
          for (int t = 0; t < MAXLEN; ++t) {
            if (t < len && (qpar < 0 || (int)((a.qsplit >> t) & 1u) == qpar) && (!prune || t == len - 1)) {
              uint8_t* qptr = B0 + umma::sw128_offset(row0 + t * spt, hd);
              const uint4 qw = *reinterpret_cast<const uint4*>(qptr);
              float sc[MAXLEN];
              float mx = -INFINITY;
#pragma unroll
              for (int j = 0; j <= t; ++j) {
                sc[j] = dot8_f16(qw, kp[j]);
                mx = fmaxf(mx, sc[j]);
              }
              // softmax weights e_j = exp((s_j - max) / sqrt(dh)) in fp32, handed to the fp16-operand PV product as
              // a (high, low) pair of halves: e_j = hi_j + lo_j to 2^-22, so the product keeps fp32 weights
              const float nmx = -mx * qk_scale_log2e;
              uint32_t ph[(MAXLEN + 1) / 2], pl[(MAXLEN + 1) / 2];
              float z = 0.f;
#pragma unroll
              for (int j = 0; j <= t; j += 2) {
                const float e0 = ex2_approx(fmaf(sc[j], qk_scale_log2e, nmx));
                const float e1 = j + 1 <= t ? ex2_approx(fmaf(sc[j + 1], qk_scale_log2e, nmx)) : 0.f;
                z += e0 + e1;
                ph[j >> 1] = umma::pack_f16x2(e0, e1);
                const float2 hi = umma::unpack_f16x2(ph[j >> 1]);
                pl[j >> 1] = umma::pack_f16x2(e0 - hi.x, e1 - hi.y);
              }
              const float inv_z = __fdividef(1.0f, z);
              float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int j = 0; j <= t; ++j) {
                if (j & 1) { axpy8_f16<true>(ph[j >> 1], vp[j], o); axpy8_f16<true>(pl[j >> 1], vp[j], o); }
                else       { axpy8_f16<false>(ph[j >> 1], vp[j], o); axpy8_f16<false>(pl[j >> 1], vp[j], o); }
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) o[c] *= inv_z;
              if (fl & BFB200_SITE_HEAD_OUT) {   // transformer.py:116-122
                const uint32_t k = (half_keep_mask(ms, prr, (uint32_t)t, sb + 4u, (uint32_t)(hd >> 2)) >> (8 * (hd & 3))) & 0xFFu;
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = ((k >> c) & 1u) ? o[c] * ms.scale : 0.f;
              }
              uint4 w;
              w.x = umma::pack_f16x2(o[0], o[1]); w.y = umma::pack_f16x2(o[2], o[3]);
              w.z = umma::pack_f16x2(o[4], o[5]); w.w = umma::pack_f16x2(o[6], o[7]);
              *reinterpret_cast<uint4*>(qptr) = w;
            }
          }
        }
      }

      if (!live) {
        // dead rows of the last layer: take part in the barriers of the three products and the two LayerNorm
        // exchanges (thread 0 of the group still issues the MMAs inside gemm()), compute nothing
        gemm(0, wl + lay.off_out, FD, 4);
        if (!fold_ln1) umma::group_sync(bar_id, FGROUP_THREADS);
        gemm(0, wl + lay.off_w1, (int)lay.fp, 4);
        gemm(1, wl + lay.off_w2, FD, (int)(lay.fp >> 4));
        umma::group_sync(bar_id, FGROUP_THREADS);
        continue;
      }

      // ---- W_out, residual from the block input, LayerNorm1 (transformer.py:125, :226, :231) ----
      gemm(0, wl + lay.off_out, FD, 4);
      {
        const uint32_t k_attn = (fl & BFB200_SITE_ATTN_RESID) ? half_keep_mask(ms, rr, (uint32_t)pos, sb + 1u, (uint32_t)h) : 0u;
        const uint32_t k_ffn_in = (fl & BFB200_SITE_FFN_IN) ? half_keep_mask(ms, rr, (uint32_t)pos, sb + 2u, (uint32_t)h) : 0u;
        float v[HW];
        tmem_load_half(tW + h * HW, v);
        if (fl & BFB200_SITE_ATTN_RESID) drop_half(v, k_attn, ms.scale);
        {
          float x[HW];
          tmem_load_half(tX, x);
#pragma unroll
          for (int i = 0; i < HW; ++i) v[i] += x[i];
        }
        const float2 own = ln_partial(v);
        red[h * FROWS + row] = own;
        if (!fold_ln1) {
          umma::group_sync(bar_id, FGROUP_THREADS);
          ln_finish(v, own, red[(h ^ 1) * FROWS + row], pl + h * HW, pl + 64 + h * HW, a.ln_eps);
          drop_half(v, k_ffn_in, ms.scale);
        }
        // Otherwise LayerNorm1 only feeds the FFN1 product with no dropout site in between, so it is folded into that
        // product (W1 * gamma1 in the image): the UN-normalised row is the operand, the statistics wait in `red` for
        // the FFN1 epilogue, and the exchange barrier merges with the product's own
        store_half_f16(B0, row, h, v);
      }

      // ---- FFN: Linear(d -> F) + bias + ReLU (transformer.py:151-155); each half takes fp / 2 columns ----
      gemm(0, wl + lay.off_w1, (int)lay.fp, 4);
      {
        const float* b1 = pl + 128;
        const float* col_s = pl + 384;
        const float* col_c = pl + 448;
        float ln_mean = 0.f, ln_rstd = 1.f;
        if (fold_ln1) {   // W1 LN(v) + b1 = rstd (v (W1 gamma)^T - mean col_s) + col_c
          const float2 own = red[h * FROWS + row], other = red[(h ^ 1) * FROWS + row];
          const float dm = own.x - other.x;
          ln_mean = 0.5f * (own.x + other.x);
          ln_rstd = rsqrtf((own.y + other.y + 16.0f * dm * dm) * (1.0f / FD) + a.ln_eps);
        }
        const uint32_t half_cols = lay.fp >> 1;   // multiple of 8
        for (uint32_t c0 = half_cols * h; c0 < half_cols * (h + 1); c0 += 8) {
          uint32_t r[8];
          umma::tmem_ld8(tW + c0, r);
          umma::tmem_ld_wait();
          float hc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            hc[i] = fold_ln1 ? fmaxf(fmaf(ln_rstd, fmaf(-ln_mean, col_s[c0 + i], __uint_as_float(r[i])), col_c[c0 + i]), 0.f)
                             : fmaxf(__uint_as_float(r[i]) + b1[c0 + i], 0.f);
          uint4 w;
          w.x = umma::pack_f16x2(hc[0], hc[1]); w.y = umma::pack_f16x2(hc[2], hc[3]);
          w.z = umma::pack_f16x2(hc[4], hc[5]); w.w = umma::pack_f16x2(hc[6], hc[7]);
          *reinterpret_cast<uint4*>(B1 + umma::sw128_offset(row, c0 >> 3)) = w;
        }
      }

      // ---- FFN: Linear(F -> d) + bias, residual from the block input, LayerNorm2, layer-out dropout ----
      gemm(1, wl + lay.off_w2, FD, (int)(lay.fp >> 4));
      {
        const uint32_t k_resid = (fl & BFB200_SITE_FFN_RESID) ? half_keep_mask(ms, rr, (uint32_t)pos, sb + 3u, (uint32_t)h) : 0u;
        const uint32_t k_out = (!FAST && (fl & BFB200_SITE_LAYER_OUT)) ? half_keep_mask(ms, rr, (uint32_t)pos, sb + 0u, (uint32_t)h) : 0u;
        float v[HW];
        tmem_load_half(tW + h * HW, v);
        const float* b2 = pl + 192 + h * HW;
#pragma unroll
        for (int i = 0; i < HW; ++i) v[i] += b2[i];
        if (fl & BFB200_SITE_FFN_RESID) drop_half(v, k_resid, ms.scale);
        {
          float x[HW];
          tmem_load_half(tX, x);
#pragma unroll
          for (int i = 0; i < HW; ++i) v[i] += x[i];
        }
        const float2 own = ln_partial(v);
        red[h * FROWS + row] = own;
        umma::group_sync(bar_id, FGROUP_THREADS);
        ln_finish(v, own, red[(h ^ 1) * FROWS + row], pl + 256 + h * HW, pl + 320 + h * HW, a.ln_eps);
        if (FAST == 1) drop_half_direct(v, ms, rr, (uint32_t)pos, sb + 0u, (uint32_t)h);   // transformer.py:370
        else if (fl & BFB200_SITE_LAYER_OUT) drop_half(v, k_out, ms.scale);
        if (!prune) tmem_store_half(tX, v);
        store_half_f16(B0, row, h, v);
      }
    }

    // ---- vocabulary head on every row (only the last position is consumed, active_learning.py:146) ----
    gemm(0, img_u32 + lay.off_head, FMAXV, 4);
    const int stage_stride = FMAXV + a.tps;
    float* stage = reinterpret_cast<float*>(B1);  // B1 + B2: 32 KB (K / V / FFN hidden are dead)
    {
      const float* hb = params + a.n_layers * PARAMS_PER_LAYER;
      const bool is_last = row_valid && pos == len - 1;
      for (uint32_t c0 = 64u * h; c0 < 64u * h + 64u; c0 += 16) {
        uint32_t r[16];
        umma::tmem_ld16(tW + c0, r);
        umma::tmem_ld_wait();
        if (is_last) {
#pragma unroll
          for (int i = 0; i < 16; ++i) stage[seq_l * stage_stride + c0 + i] = __uint_as_float(r[i]) + hb[c0 + i];
        }
      }
    }
    umma::tc_fence_before_sync();
    umma::group_sync(bar_id, FGROUP_THREADS);

    // ---- log-softmax, probabilities, score, next token (transformer.py:375; active_learning.py:146-152) ----
    if (a.tps == 16) score_phase<16, FAST == 1 || FAST == 2>(a, stage, stage_stride, gt, nseq, s0, len);
    else             score_phase<8, FAST == 1 || FAST == 2>(a, stage, stage_stride, gt, nseq, s0, len);
    // the staging rows are rewritten only after the next tile's first product: barriers in between order it
  }

#ifdef BFB_FUSED_TIMING
  if ((gt & 31) == 5) {
    atomicAdd(&bfb_fused_dbg[0], (unsigned long long)dbg[0]);
    atomicAdd(&bfb_fused_dbg[1], (unsigned long long)dbg[1]);
    atomicAdd(&bfb_fused_dbg[2], (unsigned long long)dbg[2]);
    atomicAdd(&bfb_fused_dbg[3], (unsigned long long)(clock64() - t_begin));
    atomicAdd(&bfb_fused_dbg[4], 1ull);
  }
  if (gt == 0) {
    atomicAdd(&bfb_fused_dbg[5], (unsigned long long)dbi[0]);
    atomicAdd(&bfb_fused_dbg[6], (unsigned long long)dbi[1]);
    atomicAdd(&bfb_fused_dbg[7], (unsigned long long)dbi[2]);
  }
#endif
  umma::tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem_base, 512);
}

}  // namespace

#ifdef BFB_FUSED_TIMING
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
  const FusedLayout lay = fused_layout(m.n_layers, m.d_ff, m.ntoken);
  return fused_smem_bytes(lay) + kFusedStaticSlack <= 227 * 1024;
}

size_t fused_image_bytes(const bfb200_transformer& m) {
  return fused_supported(m) ? fused_layout(m.n_layers, m.d_ff, m.ntoken).image_bytes : 0;
}

int fused_pack(const bfb200_transformer& m, void* image, size_t bytes, cudaStream_t st) {
  BFB_REQUIRE(fused_supported(m), BFB200_ERR_UNSUPPORTED, "model shape is outside the fused kernel's class");
  const FusedLayout lay = fused_layout(m.n_layers, m.d_ff, m.ntoken);
  BFB_REQUIRE(bytes >= lay.image_bytes, BFB200_ERR_WORKSPACE, "fused image buffer too small: %zu < %u", bytes,
              lay.image_bytes);
  cudaError_t e = cudaMemsetAsync(image, 0, lay.image_bytes, st);
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
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.image = (const uint8_t*)m.fused_image;
  a.embedding = m.embedding;
  a.lay = fused_layout(m.n_layers, m.d_ff, m.ntoken);
  a.n_layers = m.n_layers; a.d_ff = m.d_ff; a.vocab = m.ntoken;
  a.site_flags = m.site_flags;
  a.ln_eps = m.ln_eps;
  a.len = len;
  a.spt = FROWS / len < 32 ? FROWS / len : 32;   // the score phase gives every sequence >= 8 threads
  a.tps = (FGROUP_THREADS / a.spt) >= 16 ? 16 : 8;
  {   // balanced two-way split of the queries 0..len-1 with weights t + 1 (exhaustive, cached per length)
    static uint32_t cache[17] = {0};
    static bool have[17] = {false};
    const int n = len < 16 ? len : 16;
    if (!have[n]) {
      uint32_t best = 0xAAAAAAAAu;
      int best_max = 1 << 30;
      for (uint32_t m = 0; m < (1u << n); ++m) {
        int s1 = 0, s0 = 0;
        for (int t = 0; t < n; ++t) ((m >> t) & 1u ? s1 : s0) += t + 1;
        const int mx = s1 > s0 ? s1 : s0;
        if (mx < best_max) { best_max = mx; best = m; }
      }
      cache[n] = best;
      have[n] = true;
    }
    a.qsplit = cache[n];
  }
  a.n_seq = n_seq;
  a.n_tiles = (n_seq + a.spt - 1) / a.spt;
  a.tokens = tokens; a.tokens_out = tokens_out; a.tok_stride = tok_stride;
  a.scheme = scheme; a.mode = mode; a.temperature = temperature;
  a.forced = forced; a.seq_scores = seq_scores; a.logp_out = logp_out;
  a.ms = ms;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const size_t smem = fused_smem_bytes(a.lay);
  int64_t grid = (a.n_tiles + FGROUPS - 1) / FGROUPS;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  const bool plain = a.ms.bits == nullptr && a.logp_out == nullptr && a.forced == nullptr;
  const bool sampling = scheme == BFB200_SCHEME_SAMPLING;
  const int fast = !plain ? 0
                   : (!sampling && a.site_flags == (uint32_t)(BFB200_SITE_EMBED | BFB200_SITE_LAYER_OUT)) ? 1
                   : (a.site_flags == 0u) ? (sampling ? 3 : 2) : 0;
#define FUSED_LAUNCH_F(ML, FV)                                                                              \
  do {                                                                                                      \
    cudaFuncSetAttribute(mc_forward_kernel<ML, FV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    mc_forward_kernel<ML, FV><<<(int)grid, FTHREADS, smem, st>>>(a);                                        \
  } while (0)
#define FUSED_LAUNCH(ML)                  \
  do {                                    \
    if (fast == 1) FUSED_LAUNCH_F(ML, 1); \
    else if (fast == 2) FUSED_LAUNCH_F(ML, 2); \
    else if (fast == 3) FUSED_LAUNCH_F(ML, 3); \
    else FUSED_LAUNCH_F(ML, 0);           \
  } while (0)
  // the pair phase keeps a sequence's K and V rows in registers: one instantiation per length up to 8 (the decode steps
  // of the reference's addition task run at lengths 4..8), so that the arrays are exactly as long as the sequences
  if (len <= 4) FUSED_LAUNCH(4);
  else if (len == 5) FUSED_LAUNCH(5);
  else if (len == 6) FUSED_LAUNCH(6);
  else if (len == 7) FUSED_LAUNCH(7);
  else if (len == 8) FUSED_LAUNCH(8);
  else if (len <= 16) FUSED_LAUNCH(16);
  else FUSED_LAUNCH(FMAXLEN);
#undef FUSED_LAUNCH
#undef FUSED_LAUNCH_F
  BFB_LAUNCH_CHECK("mc_forward_kernel", st);
  return BFB200_OK;
}

}  // namespace bfb

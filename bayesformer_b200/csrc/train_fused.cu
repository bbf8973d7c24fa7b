// Fused fine-tune step of the notebook-default model class: forward, cross-entropy loss, backward and the AdamW
// update of ONE batch (reference loop: src/libs/utils.py:68-78 — model(train batch), CrossEntropyLoss on the
// log-softmax outputs of the answer positions, loss.backward(), AdamW.step()).
//
// Why: the model has 50 k parameters and a batch is 20 sequences x 9 tokens = 180 token rows; the reference step is
// ~600 PyTorch kernels of a few microseconds each (launch-bound: 9.4 ms eager, 1.9 ms as a CUDA graph).  Here the whole
// forward + backward of ONE SEQUENCE runs in ONE CTA (1024 threads) with every activation of that sequence in shared
// memory (~130 KB at 16 tokens, 2 layers): attention never crosses sequences, so the only cross-sequence data are the
// weights (read-only during the step, shared through L2) and the parameter gradients.  Each CTA writes its sequence's
// gradient to its own slab; a second kernel sums the slabs in a fixed order (deterministic, no float atomics across
// CTAs), applies AdamW and scatters the parameters back.  Four launches per step (gather, transpose, sequences,
// AdamW), ~50 CTA barriers instead of the round-1 design's 39 cluster barriers over L2-resident activations.
// fp32 CUDA-core arithmetic throughout (21 M multiply-adds per step: nothing for tensor cores to win here).
//
// Arithmetic follows src/model/transformer.py in TRAINING mode: nn.Dropout(p) after the positional encoding (:272), on
// the attention branch (:231) and on the FFN branch (:233), plus the always-on bayes sites (embedding :356, layer output
// :370); both residuals start from the block input (:231-233); LayerNorm eps 1e-5, biased variance.  Dropout masks come
// from the library's Philox stream (key = seed, counter = (row, step id, site, feature / 4)) and are re-drawn, not
// stored, in the backward pass.  The loss is CrossEntropyLoss applied to log-probabilities (utils.py:72-76): since
// log_softmax is idempotent this is the ordinary negative log-likelihood, averaged over ALL answer positions (PAD
// targets included — the reference sets no ignore_index).  AdamW: torch.optim.AdamW defaults (betas 0.9 / 0.999, eps
// 1e-8, weight decay 0.01), decoupled decay, bias correction.
//
// Parameters stay where torch keeps them: a segment table (flat offset, device pointer, length) maps the module's
// parameter tensors — 24 per-head projection matrices per layer among them — to one flat layout; the first kernel
// gathers them, the last scatters the updated values back.
#include "common.cuh"
#include "kernels.cuh"

namespace bfb {

namespace {

constexpr int TD = 64, TH = 8, TDH = 8;
constexpr int TMAXT = 16;   // longest training sequence (prompt + answer + EOS): per-query score rows live in registers
constexpr int TTHREADS = 1024;
constexpr int TSMEM_MAX = 227 * 1024;
constexpr int TSTAGE = 3 * TD * TD;   // floats of the weight stage: the largest matrix a phase parks (W_qkv)

struct TrainLayout {   // offsets in floats inside the flat parameter layout
  int V, F, L;
  int emb, layer0, per_layer, head_w, head_b, total;
  // inside a layer
  int wqkv, wout, ln1w, ln1b, w1, b1, w2, b2, ln2w, ln2b;
};

__host__ __device__ inline TrainLayout train_layout(int V, int F, int L) {
  TrainLayout t;
  t.V = V; t.F = F; t.L = L;
  t.emb = 0;
  t.layer0 = V * TD;
  t.wqkv = 0;
  t.wout = t.wqkv + 3 * TD * TD;
  t.ln1w = t.wout + TD * TD;
  t.ln1b = t.ln1w + TD;
  t.w1 = t.ln1b + TD;
  t.b1 = t.w1 + F * TD;
  t.w2 = t.b1 + F;
  t.b2 = t.w2 + TD * F;
  t.ln2w = t.b2 + TD;
  t.ln2b = t.ln2w + TD;
  t.per_layer = t.ln2b + TD;
  t.head_w = t.layer0 + L * t.per_layer;
  t.head_b = t.head_w + V * TD;
  t.total = t.head_b + V;
  return t;
}


struct SeqScratch {   // offsets in floats inside a CTA's shared memory: the activations of ONE sequence (T rows)
  int x;        // (L + 1) x T x 64 : block inputs, x[L] = last block's output
  int qkv;      // L x T x 192
  int prob;     // L x H x T x T
  int att;      // L x T x 64   concatenated head outputs
  int v1;       // L x T x 64   x + dropout(attn W_out^T): LayerNorm1 input
  int h1;       // L x T x 64   LayerNorm1 output
  int u;        // L x T x F    FFN pre-activation
  int y;        // L x T x 64   x + dropout(FFN): LayerNorm2 input
  int stat;     // L x T x 4    mean1, rstd1, mean2, rstd2
  int logit;    // (T - P) x V  logits -> d logits
  int dx;       // T x 64       gradient w.r.t. the current block output
  int dxin;     // T x 64       gradient w.r.t. the current block input (accumulated)
  int dtmp;     // T x 64       d(attn branch) / d(ffn branch)
  int dqkv;     // T x 192
  int du;       // T x F
  int datt;     // T x 64
  int dscore;   // H x T x T    d(scores)
  int total;
};

__host__ __device__ inline SeqScratch seq_scratch(int T, int P, int V, int F, int L) {
  SeqScratch s;
  int o = 0;
  auto take = [&](int n) { const int at = o; o += (n + 3) & ~3; return at; };   // float4-aligned regions
  s.x = take((L + 1) * T * TD);
  s.qkv = take(L * T * 3 * TD);
  s.prob = take(L * TH * T * T);
  s.att = take(L * T * TD);
  s.v1 = take(L * T * TD);
  s.h1 = take(L * T * TD);
  s.u = take(L * T * F);
  s.y = take(L * T * TD);
  s.stat = take(L * T * 4);
  s.logit = take((T - P) * V);
  s.dx = take(T * TD);
  s.dxin = take(T * TD);
  s.dtmp = take(T * TD);
  s.dqkv = take(T * 3 * TD);
  s.du = take(T * F);
  s.datt = take(T * TD);
  s.dscore = take(TH * T * T);
  s.total = o;
  return s;
}

struct TrainGlobal {   // offsets in floats inside the global scratch buffer
  // transposed copies for the forward products: per layer W_qkv^T (64 x 192), W_out^T (64 x 64), W1^T (64 x F); then
  // head^T (64 x Vp), Vp = V rounded up to 4 (a thread owns 4 adjacent outputs and reads them as one float4 per k)
  int wt, wt_layer, wt_out, wt_w1, wt_head, vp;
  int slab;     // B x total : one gradient slab per sequence
  int loss;     // B         : per-sequence loss sums
  int total;
};

__host__ __device__ inline TrainGlobal train_global(int B, int V, int F, int L, int n_params) {
  TrainGlobal g;
  g.vp = (V + 3) & ~3;
  g.wt = 0;
  g.wt_out = TD * 3 * TD;
  g.wt_w1 = g.wt_out + TD * TD;
  g.wt_layer = (g.wt_w1 + TD * F + 3) & ~3;
  g.wt_head = L * g.wt_layer;
  int o = g.wt_head + TD * g.vp;
  o = (o + 3) & ~3;
  g.slab = o; o += B * ((n_params + 3) & ~3);
  g.loss = o; o += B;
  g.total = (o + 3) & ~3;
  return g;
}

struct TrainSeg { int offset, count; float* ptr; };   // flat [offset, offset + count) <-> ptr[0 .. count)

struct TrainArgs {
  float* flat;          // parameters (gathered), total floats
  float* grad;          // gradients of the step (slabs summed), total floats
  float* adam_m;        // AdamW first / second moments, total floats each (persist between steps)
  float* adam_v;
  const TrainSeg* segs;
  int n_segs;
  const int64_t* tokens;   // (T, B) seq-first: prompts followed by answers (utils.py:61-66)
  const float* pe;         // (pe_len, 64)
  float* scratch;
  float* loss_out;         // += mean loss of this batch
  TrainLayout lay;
  SeqScratch sc;
  TrainGlobal gl;
  int slab_stride;
  int B, T, P;
  float lr, beta1, beta2, eps, weight_decay, bias1, bias2;   // bias_k = 1 - beta_k^step
  float p_train, p_bayes;   // nn.Dropout rate (train mode) and bayes rate (0: site off)
  uint32_t thr_train, thr_bayes;   // keep iff a Philox word >= threshold (p * 2^32)
  float ln_eps;
  uint2 key;
  uint32_t step_id;
};

// dropout sites of the training forward (counter word 2)
enum { TS_EMBED = 0, TS_PE = 1, TS_ATTN = 2 /* + 4 l */, TS_FFN = 3, TS_LAYER_OUT = 4 };

// four keep decisions for features [4 g4, 4 g4 + 4) of (row, site): one Philox call, one 32-bit word per decision
__device__ __forceinline__ uint32_t keep4_train(const TrainArgs& a, uint32_t row, uint32_t site, uint32_t g4, uint32_t thr) {
  const uint4 r = philox4x32_10(make_uint4(row, a.step_id, site, g4), a.key);
  return (r.x >= thr ? 1u : 0u) | (r.y >= thr ? 2u : 0u) | (r.z >= thr ? 4u : 0u) | (r.w >= thr ? 8u : 0u);
}
__device__ __forceinline__ float drop1(const TrainArgs& a, uint32_t row, uint32_t site, uint32_t f, uint32_t thr, float p, float v) {
  if (p <= 0.f) return v;
  const uint32_t k = keep4_train(a, row, site, f >> 2, thr);
  return ((k >> (f & 3u)) & 1u) ? v * (1.0f / (1.0f - p)) : 0.f;
}
// the same for four adjacent features (one Philox call)
__device__ __forceinline__ float4 drop4(const TrainArgs& a, uint32_t row, uint32_t site, uint32_t g4, uint32_t thr, float p, float4 v) {
  if (p <= 0.f) return v;
  const uint32_t k = keep4_train(a, row, site, g4, thr);
  const float s = 1.0f / (1.0f - p);
  return make_float4((k & 1u) ? v.x * s : 0.f, (k & 2u) ? v.y * s : 0.f, (k & 4u) ? v.z * s : 0.f, (k & 8u) ? v.w * s : 0.f);
}
__device__ __forceinline__ void fma4(float4& acc, float d, const float4& w) {
  acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// "Quad" phases: out[r][4 f4 ..] = sum_c in[r][c] * Wm[c][4 f4 ..] with thread = (row r, feature group f4, quarter
// `part` of the c range).  Lane bits 0-2 are the low bits of f4 and bits 3-4 the part, so the eight lanes of a
// quarter-warp — the unit a 128-bit shared-memory load is served in — read 128 contiguous bytes of one Wm row
// (with the part in the low lane bits they hit the same eight banks four times: measured 7.5 us instead of 2).
__device__ __forceinline__ int quad_f4(int tid) { return (tid & 7) | ((tid >> 2) & 8); }
__device__ __forceinline__ int quad_part(int tid) { return (tid >> 3) & 3; }
__device__ __forceinline__ int quad_row(int tid) { return tid >> 6; }
__device__ __forceinline__ float4 quad_sum(float4 v) {   // sum over the 4 lanes that differ in the part bits
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, o); v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, o);
    v.z += __shfl_xor_sync(0xFFFFFFFFu, v.z, o); v.w += __shfl_xor_sync(0xFFFFFFFFu, v.w, o);
  }
  return v;
}

#ifdef BFB_TRAIN_TIMING
__device__ unsigned long long bfb_train_dbg[128];
__device__ int bfb_train_dbg_n;
#endif
__device__ __forceinline__ void cta_sync() {
  __syncthreads();
#ifdef BFB_TRAIN_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int k = bfb_train_dbg_n;
    if (k < 128) { bfb_train_dbg[k] = t; bfb_train_dbg_n = k + 1; }
  }
#endif
}

// flat <- the module's parameter tensors; block (segment, chunk)
__global__ void __launch_bounds__(256) train_gather_kernel(const TrainArgs a) {
  const TrainSeg sg = a.segs[blockIdx.x];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < sg.count; i += gridDim.y * blockDim.x) a.flat[sg.offset + i] = sg.ptr[i];
}

// transposed weight copies for the forward products (reads strided, writes coalesced; 30 k elements)
__global__ void __launch_bounds__(256) train_transpose_kernel(const TrainArgs a) {
  const TrainLayout& lay = a.lay;
  const TrainGlobal& gl = a.gl;
  const float* W = a.flat;
  float* WT = a.scratch + gl.wt;
  const int F = lay.F, V = lay.V;
  const int per = TD * 3 * TD + TD * TD + TD * F;
  const int n = lay.L * per + TD * gl.vp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < lay.L * per) {
      const int l = i / per;
      int e = i % per;
      const float* Wl = W + lay.layer0 + l * lay.per_layer;
      float* T0 = WT + l * gl.wt_layer;
      if (e < TD * 3 * TD) { const int k = e / (3 * TD), c = e % (3 * TD); T0[e] = Wl[lay.wqkv + c * TD + k]; continue; }
      e -= TD * 3 * TD;
      if (e < TD * TD) { const int k = e / TD, nn = e % TD; T0[gl.wt_out + e] = Wl[lay.wout + nn * TD + k]; continue; }
      e -= TD * TD;
      { const int k = e / F, m = e % F; T0[gl.wt_w1 + e] = Wl[lay.w1 + m * TD + k]; }
    } else {
      const int e = i - lay.L * per, k = e / gl.vp, v = e % gl.vp;
      WT[gl.wt_head + e] = v < V ? W[lay.head_w + v * TD + k] : 0.f;
    }
  }
}

// forward + backward of sequence blockIdx.x; its parameter gradient goes to slab blockIdx.x
__global__ void __launch_bounds__(TTHREADS, 1) train_seq_kernel(const TrainArgs a) {
  extern __shared__ __align__(16) float S[];
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARPS = TTHREADS / 32;
  const TrainLayout& lay = a.lay;
  const SeqScratch& sc = a.sc;
  const TrainGlobal& gl = a.gl;
  const int b = (int)blockIdx.x;
  const int B = a.B, T = a.T, P = a.P, V = lay.V, F = lay.F, L = lay.L;
  const int NPL = T - P;                       // predicted positions of this sequence: rows t in [P - 1, T - 2] (utils.py:72-76)
  const float inv_np = 1.0f / (float)(B * NPL);
  const uint32_t row0 = (uint32_t)(b * T);     // Philox row id of token t: row0 + t
  const float* W = a.flat;
  const float* WT = a.scratch + gl.wt;
  float* G = a.scratch + gl.slab + (size_t)b * a.slab_stride;
  const float inv_sqrt_dh = rsqrtf((float)TDH);
  const float sqrt_d = sqrtf((float)TD);
  // Weight stage (48 KB behind the activations): a phase that multiplies by a weight matrix first prefetches it into
  // registers with coalesced loads (one L2 round trip for the whole CTA), does its weight-free work, then parks it here,
  // so that the per-thread reduction loops read shared memory instead of paying the L2 latency once per unrolled batch.
  float* WS = S + sc.total + 4;
  auto prefetch = [&](float4 (&pf)[3], const float* src, int n4) {
#pragma unroll
    for (int q = 0; q < 3; ++q) pf[q] = tid + q * TTHREADS < n4 ? ld4(src + 4 * (tid + q * TTHREADS)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto park = [&](const float4 (&pf)[3], int n4) {
#pragma unroll
    for (int q = 0; q < 3; ++q) if (tid + q * TTHREADS < n4) st4(WS + 4 * (tid + q * TTHREADS), pf[q]);
    __syncthreads();
  };
  float4 pf[3];
#define FOR_T(i, n) for (int i = tid; i < (n); i += TTHREADS)

  // ---- forward ----------------------------------------------------------------------------------------------
  __shared__ int s_tok[TMAXT];
  __shared__ float s_lossv[TMAXT];   // per predicted position, summed in order
  if (tid < T) s_tok[tid] = (int)a.tokens[(int64_t)tid * B + b];
  FOR_T(i, V * TD / 4) st4(G + lay.emb + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));   // embedding rows are accumulated at the end
  // embedding, bayes dropout E, sqrt(d) scale, positional encoding, nn.Dropout (transformer.py:355-362, :269-272)
  FOR_T(i, T * (TD / 4)) {
    const int t = i / (TD / 4), g4 = i % (TD / 4);
    const int tok = (int)a.tokens[(int64_t)t * B + b];
    float4 v = ld4(W + lay.emb + tok * TD + 4 * g4);
    const float4 pe = ld4(a.pe + t * TD + 4 * g4);
    v = drop4(a, row0 + t, TS_EMBED, (uint32_t)g4, a.thr_bayes, a.p_bayes, v);
    v = make_float4(fmaf(v.x, sqrt_d, pe.x), fmaf(v.y, sqrt_d, pe.y), fmaf(v.z, sqrt_d, pe.z), fmaf(v.w, sqrt_d, pe.w));
    st4(S + sc.x + t * TD + 4 * g4, drop4(a, row0 + t, TS_PE, (uint32_t)g4, a.thr_train, a.p_train, v));
  }
  cta_sync();

  for (int l = 0; l < L; ++l) {
    const float* Wl = W + lay.layer0 + l * lay.per_layer;
    const float* X = S + sc.x + l * T * TD;
    float* QKV = S + sc.qkv + l * T * 3 * TD;
    float* PR = S + sc.prob + l * TH * T * T;
    float* ATT = S + sc.att + l * T * TD;
    float* V1 = S + sc.v1 + l * T * TD;
    float* H1 = S + sc.h1 + l * T * TD;
    float* U = S + sc.u + l * T * F;
    float* Y = S + sc.y + l * T * TD;
    float* ST = S + sc.stat + l * T * 4;
    float* XN = S + sc.x + (l + 1) * T * TD;
    const uint32_t site0 = 4u * (uint32_t)l;
    const float* WTl = WT + l * gl.wt_layer;
    // Q, K, V of all heads (bias-free, transformer.py:54-57): thread = (row, four adjacent outputs)
    prefetch(pf, WTl, 3 * TD * TD / 4);
    park(pf, 3 * TD * TD / 4);
    FOR_T(i, T * (3 * TD / 4)) {
      const int r = i / (3 * TD / 4), c4 = i % (3 * TD / 4);
      const float* xr = X + r * TD;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16
      for (int k = 0; k < TD; ++k) fma4(acc, xr[k], ld4(WS + k * 3 * TD + 4 * c4));
      st4(QKV + r * 3 * TD + 4 * c4, acc);
    }
    cta_sync();
    // causal softmax(QK^T / sqrt(dh)) V, thread = (head, query) (transformer.py:58-65); scores in registers
    prefetch(pf, WTl + gl.wt_out, TD * TD / 4);   // W_out^T for the next phase travels meanwhile
    FOR_T(i, TH * T) {
      const int h = i / T, t = i % T;
      const float* qp = QKV + t * 3 * TD + h * TDH;
      const float4 q0 = ld4(qp), q1 = ld4(qp + 4);
      float s[TMAXT];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) {
        s[j] = -INFINITY;
        if (j <= t) {
          const float* kp = QKV + j * 3 * TD + TD + h * TDH;
          const float4 k0 = ld4(kp), k1 = ld4(kp + 4);
          const float d = ((q0.x * k0.x + q0.y * k0.y) + (q0.z * k0.z + q0.w * k0.w)) +
                          ((q1.x * k1.x + q1.y * k1.y) + (q1.z * k1.z + q1.w * k1.w));
          s[j] = d * inv_sqrt_dh;
          mx = fmaxf(mx, s[j]);
        }
      }
      float z = 0.f;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) { s[j] = j <= t ? __expf(s[j] - mx) : 0.f; z += s[j]; }
      const float inv_z = 1.0f / z;
      float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
      float* pr = PR + (h * T + t) * T;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) {
        if (j < T) {
          const float p = s[j] * inv_z;
          pr[j] = p;
          if (j <= t) {
            const float* vp = QKV + j * 3 * TD + 2 * TD + h * TDH;
            fma4(o0, p, ld4(vp)); fma4(o1, p, ld4(vp + 4));
          }
        }
      }
      st4(ATT + t * TD + h * TDH, o0);
      st4(ATT + t * TD + h * TDH + 4, o1);
    }
    park(pf, TD * TD / 4);
    cta_sync();
    // W_out, nn.Dropout, residual from the block input (transformer.py:125, :231)
    FOR_T(i, T * (TD / 4)) {
      const int r = i / (TD / 4), n4 = i % (TD / 4);
      const float* ar = ATT + r * TD;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16
      for (int k = 0; k < TD; ++k) fma4(acc, ar[k], ld4(WS + k * TD + 4 * n4));
      acc = drop4(a, row0 + r, site0 + TS_ATTN, (uint32_t)n4, a.thr_train, a.p_train, acc);
      const float4 xv = ld4(X + r * TD + 4 * n4);
      st4(V1 + r * TD + 4 * n4, make_float4(xv.x + acc.x, xv.y + acc.y, xv.z + acc.z, xv.w + acc.w));
    }
    cta_sync();
    // LayerNorm1: warp per row, two features per lane
    for (int r = warp; r < T; r += NWARPS) {
      const float v0 = V1[r * TD + lane], v1 = V1[r * TD + 32 + lane];
      const float mean = warp_sum(v0 + v1) * (1.0f / TD);
      const float c0 = v0 - mean, c1 = v1 - mean;
      const float rstd = rsqrtf(warp_sum(c0 * c0 + c1 * c1) * (1.0f / TD) + a.ln_eps);
      if (lane == 0) { ST[r * 4 + 0] = mean; ST[r * 4 + 1] = rstd; }
      H1[r * TD + lane] = c0 * rstd * Wl[lay.ln1w + lane] + Wl[lay.ln1b + lane];
      H1[r * TD + 32 + lane] = c1 * rstd * Wl[lay.ln1w + 32 + lane] + Wl[lay.ln1b + 32 + lane];
    }
    cta_sync();
    // FFN: Linear(d -> F) (pre-activation kept for the backward pass), transformer.py:151-155
    FOR_T(i, T * F) {
      const int r = i / F, m = i % F;
      const float* hr = H1 + r * TD;
      float acc = Wl[lay.b1 + m];
#pragma unroll 16
      for (int k = 0; k < TD; ++k) acc = fmaf(hr[k], WTl[gl.wt_w1 + k * F + m], acc);
      U[i] = acc;
    }
    cta_sync();
    // Linear(F -> d) on ReLU, nn.Dropout, residual from the block input (transformer.py:233)
    FOR_T(i, TD * F) WS[(i % F) * TD + i / F] = Wl[lay.w2 + i];   // W2^T (F x 64) into the stage: coalesced read
    __syncthreads();
    FOR_T(i, T * (TD / 4)) {   // thread = (row, four adjacent outputs)
      const int r = i / (TD / 4), n4 = i % (TD / 4);
      float4 acc = ld4(Wl + lay.b2 + 4 * n4);
      for (int m = 0; m < F; ++m) fma4(acc, fmaxf(U[r * F + m], 0.f), ld4(WS + m * TD + 4 * n4));
      acc = drop4(a, row0 + r, site0 + TS_FFN, (uint32_t)n4, a.thr_train, a.p_train, acc);
      const float4 xv = ld4(X + r * TD + 4 * n4);
      st4(Y + r * TD + 4 * n4, make_float4(xv.x + acc.x, xv.y + acc.y, xv.z + acc.z, xv.w + acc.w));
    }
    cta_sync();
    // LayerNorm2 + bayes layer-output dropout (transformer.py:233, :370)
    for (int r = warp; r < T; r += NWARPS) {
      const float v0 = Y[r * TD + lane], v1 = Y[r * TD + 32 + lane];
      const float mean = warp_sum(v0 + v1) * (1.0f / TD);
      const float c0 = v0 - mean, c1 = v1 - mean;
      const float rstd = rsqrtf(warp_sum(c0 * c0 + c1 * c1) * (1.0f / TD) + a.ln_eps);
      if (lane == 0) { ST[r * 4 + 2] = mean; ST[r * 4 + 3] = rstd; }
      const float z0 = c0 * rstd * Wl[lay.ln2w + lane] + Wl[lay.ln2b + lane];
      const float z1 = c1 * rstd * Wl[lay.ln2w + 32 + lane] + Wl[lay.ln2b + 32 + lane];
      XN[r * TD + lane] = drop1(a, row0 + r, site0 + TS_LAYER_OUT, (uint32_t)lane, a.thr_bayes, a.p_bayes, z0);
      XN[r * TD + 32 + lane] = drop1(a, row0 + r, site0 + TS_LAYER_OUT, (uint32_t)(32 + lane), a.thr_bayes, a.p_bayes, z1);
    }
    cta_sync();
  }

  // vocabulary head on the predicting positions (transformer.py:374), local pred row q = t - (P - 1)
  const float* XL = S + sc.x + L * T * TD;
  float* LG = S + sc.logit;
  prefetch(pf, WT + gl.wt_head, TD * gl.vp / 4);
  park(pf, TD * gl.vp / 4);
  FOR_T(i, NPL * (gl.vp / 4)) {
    const int q = i / (gl.vp / 4), v4 = i % (gl.vp / 4);
    const float* xr = XL + (P - 1 + q) * TD;
    const float* HT = WS;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16
    for (int k = 0; k < TD; ++k) fma4(acc, xr[k], ld4(HT + k * gl.vp + 4 * v4));
    const float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (4 * v4 + e < V) LG[q * V + 4 * v4 + e] = o[e] + W[lay.head_b + 4 * v4 + e];
  }
  cta_sync();
  // log-softmax, NLL of the next token, d logits = (softmax - onehot) / NP   (utils.py:72-76; transformer.py:375)
  for (int q = warp; q < NPL; q += NWARPS) {   // warp per predicted position
    float* lg = LG + q * V;
    const int target = (int)a.tokens[(int64_t)(P + q) * B + b];
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, lg[v]);
    mx = warp_max(mx);
    float z = 0.f;
    for (int v = lane; v < V; v += 32) z += __expf(lg[v] - mx);
    const float lse = mx + __logf(warp_sum(z));
    if (lane == 0) s_lossv[q] = lse - lg[target];
    __syncwarp();
    for (int v = lane; v < V; v += 32) lg[v] = (__expf(lg[v] - lse) - (v == target ? 1.f : 0.f)) * inv_np;
  }
  cta_sync();
  if (tid == 0) {
    float ls = 0.f;
    for (int q = 0; q < NPL; ++q) ls += s_lossv[q];
    a.scratch[gl.loss + b] = ls * inv_np;
  }

  // ---- backward ---------------------------------------------------------------------------------------------
  float* DX = S + sc.dx;
  float* DXIN = S + sc.dxin;
  float* DT = S + sc.dtmp;
  float* DQKV = S + sc.dqkv;
  float* DU = S + sc.du;
  float* DATT = S + sc.datt;
  // head: dW, db, dx_L
  prefetch(pf, W + lay.head_w, V * TD / 4);
  FOR_T(i, V * (TD / 4)) {
    const int v = i / (TD / 4), f4 = i % (TD / 4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < NPL; ++q) fma4(acc, LG[q * V + v], ld4(XL + (P - 1 + q) * TD + 4 * f4));
    st4(G + lay.head_w + v * TD + 4 * f4, acc);
  }
  FOR_T(v, V) {
    float acc = 0.f;
    for (int q = 0; q < NPL; ++q) acc += LG[q * V + v];
    G[lay.head_b + v] = acc;
  }
  park(pf, V * TD / 4);
  {   // dx_L[t] = d logits[t] W_head: thread = (row, four features, quarter of the vocabulary)
    const int part = quad_part(tid), f4 = quad_f4(tid), t = quad_row(tid);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < T && t >= P - 1 && t <= T - 2) {
      const float* dl = LG + (t - (P - 1)) * V;
#pragma unroll 8
      for (int v = part; v < V; v += 4) fma4(acc, dl[v], ld4(WS + v * TD + 4 * f4));
    }
    acc = quad_sum(acc);
    if (part == 0 && t < T) st4(DX + t * TD + 4 * f4, acc);
  }
  cta_sync();

  for (int l = L - 1; l >= 0; --l) {
    const float* Wl = W + lay.layer0 + l * lay.per_layer;
    float* Gl = G + lay.layer0 + l * lay.per_layer;
    const float* X = S + sc.x + l * T * TD;
    const float* QKV = S + sc.qkv + l * T * 3 * TD;
    const float* PR = S + sc.prob + l * TH * T * T;
    const float* ATT = S + sc.att + l * T * TD;
    const float* V1 = S + sc.v1 + l * T * TD;
    const float* H1 = S + sc.h1 + l * T * TD;
    const float* U = S + sc.u + l * T * F;
    const float* Y = S + sc.y + l * T * TD;
    const float* ST = S + sc.stat + l * T * 4;
    const uint32_t site0 = 4u * (uint32_t)l;
    // bayes layer-output dropout backward: DX <- DX * mask (in place)
    FOR_T(i, T * (TD / 4)) {
      const int r = i / (TD / 4), g4 = i % (TD / 4);
      st4(DX + 4 * i, drop4(a, row0 + r, site0 + TS_LAYER_OUT, (uint32_t)g4, a.thr_bayes, a.p_bayes, ld4(DX + 4 * i)));
    }
    cta_sync();
    // LayerNorm2: d gamma2, d beta2 (warps T ..), input gradient (warp per row): dy -> DXIN (residual to the block
    // input) and, through the FFN dropout mask, the FFN branch gradient -> DT
    if (warp >= T && warp < T + 2) {
      const int f = (warp - T) * 32 + lane;
      float dg = 0.f, db = 0.f;
      for (int r = 0; r < T; ++r) {
        const float yh = (Y[r * TD + f] - ST[r * 4 + 2]) * ST[r * 4 + 3];
        dg = fmaf(DX[r * TD + f], yh, dg);
        db += DX[r * TD + f];
      }
      Gl[lay.ln2w + f] = dg; Gl[lay.ln2b + f] = db;
    }
    if (warp < T) {
      const int r = warp;
      const float mean = ST[r * 4 + 2], rstd = ST[r * 4 + 3];
      const float dz0 = DX[r * TD + lane] * Wl[lay.ln2w + lane], dz1 = DX[r * TD + 32 + lane] * Wl[lay.ln2w + 32 + lane];
      const float yh0 = (Y[r * TD + lane] - mean) * rstd, yh1 = (Y[r * TD + 32 + lane] - mean) * rstd;
      const float s1 = warp_sum(dz0 + dz1) * (1.0f / TD), s2 = warp_sum(dz0 * yh0 + dz1 * yh1) * (1.0f / TD);
      const float dy0 = rstd * (dz0 - s1 - yh0 * s2), dy1 = rstd * (dz1 - s1 - yh1 * s2);
      DXIN[r * TD + lane] = dy0;
      DXIN[r * TD + 32 + lane] = dy1;
      DT[r * TD + lane] = drop1(a, row0 + r, site0 + TS_FFN, (uint32_t)lane, a.thr_train, a.p_train, dy0);
      DT[r * TD + 32 + lane] = drop1(a, row0 + r, site0 + TS_FFN, (uint32_t)(32 + lane), a.thr_train, a.p_train, dy1);
    }
    cta_sync();
    // FFN second layer: dW2, db2, d(pre-activation)
    FOR_T(i, TD * F) {
      const int n = i / F, m = i % F;
      float acc = 0.f;
      for (int r = 0; r < T; ++r) acc = fmaf(DT[r * TD + n], fmaxf(U[r * F + m], 0.f), acc);
      Gl[lay.w2 + i] = acc;
    }
    FOR_T(n, TD) {
      float acc = 0.f;
      for (int r = 0; r < T; ++r) acc += DT[r * TD + n];
      Gl[lay.b2 + n] = acc;
    }
    FOR_T(i, T * F) {
      const int r = i / F, m = i % F;
      float acc = 0.f;
#pragma unroll 8
      for (int n = 0; n < TD; ++n) acc = fmaf(DT[r * TD + n], Wl[lay.w2 + n * F + m], acc);
      DU[i] = U[i] > 0.f ? acc : 0.f;
    }
    cta_sync();
    // FFN first layer: dW1, db1, d(LayerNorm1 output) -> DT
    FOR_T(i, F * (TD / 4)) {
      const int m = i / (TD / 4), f4 = i % (TD / 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < T; ++r) fma4(acc, DU[r * F + m], ld4(H1 + r * TD + 4 * f4));
      st4(Gl + lay.w1 + m * TD + 4 * f4, acc);
    }
    FOR_T(m, F) {
      float acc = 0.f;
      for (int r = 0; r < T; ++r) acc += DU[r * F + m];
      Gl[lay.b1 + m] = acc;
    }
    FOR_T(i, T * (TD / 4)) {
      const int r = i / (TD / 4), f4 = i % (TD / 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int m = 0; m < F; ++m) fma4(acc, DU[r * F + m], ld4(Wl + lay.w1 + m * TD + 4 * f4));
      st4(DT + r * TD + 4 * f4, acc);
    }
    cta_sync();
    // LayerNorm1: parameter gradients, input gradient -> DXIN += dv1, attention-branch gradient (through its dropout) -> DX
    if (warp >= T && warp < T + 2) {
      const int f = (warp - T) * 32 + lane;
      float dg = 0.f, db = 0.f;
      for (int r = 0; r < T; ++r) {
        const float vh = (V1[r * TD + f] - ST[r * 4 + 0]) * ST[r * 4 + 1];
        dg = fmaf(DT[r * TD + f], vh, dg);
        db += DT[r * TD + f];
      }
      Gl[lay.ln1w + f] = dg; Gl[lay.ln1b + f] = db;
    }
    float keep_dv0 = 0.f, keep_dv1 = 0.f;
    if (warp < T) {
      const int r = warp;
      const float mean = ST[r * 4 + 0], rstd = ST[r * 4 + 1];
      const float dz0 = DT[r * TD + lane] * Wl[lay.ln1w + lane], dz1 = DT[r * TD + 32 + lane] * Wl[lay.ln1w + 32 + lane];
      const float vh0 = (V1[r * TD + lane] - mean) * rstd, vh1 = (V1[r * TD + 32 + lane] - mean) * rstd;
      const float s1 = warp_sum(dz0 + dz1) * (1.0f / TD), s2 = warp_sum(dz0 * vh0 + dz1 * vh1) * (1.0f / TD);
      keep_dv0 = rstd * (dz0 - s1 - vh0 * s2); keep_dv1 = rstd * (dz1 - s1 - vh1 * s2);
      DXIN[r * TD + lane] += keep_dv0;
      DXIN[r * TD + 32 + lane] += keep_dv1;
    }
    if (warp < T) {   // (the d gamma1 warps read DT and V1 only: DX may be overwritten in this phase)
      const int r = warp;
      DX[r * TD + lane] = drop1(a, row0 + r, site0 + TS_ATTN, (uint32_t)lane, a.thr_train, a.p_train, keep_dv0);
      DX[r * TD + 32 + lane] = drop1(a, row0 + r, site0 + TS_ATTN, (uint32_t)(32 + lane), a.thr_train, a.p_train, keep_dv1);
    }
    cta_sync();
    // W_out: dW, d(concatenated head outputs) -> DATT
    prefetch(pf, Wl + lay.wout, TD * TD / 4);
    FOR_T(i, TD * (TD / 4)) {
      const int n = i / (TD / 4), k4 = i % (TD / 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < T; ++r) fma4(acc, DX[r * TD + n], ld4(ATT + r * TD + 4 * k4));
      st4(Gl + lay.wout + n * TD + 4 * k4, acc);
    }
    park(pf, TD * TD / 4);
    {   // thread = (row, four inputs, quarter of the outputs)
      const int part = quad_part(tid), k4 = quad_f4(tid), r = quad_row(tid);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < T) {
#pragma unroll 16
        for (int n = part; n < TD; n += 4) fma4(acc, DX[r * TD + n], ld4(WS + n * TD + 4 * k4));
      }
      acc = quad_sum(acc);
      if (part == 0 && r < T) st4(DATT + r * TD + 4 * k4, acc);
    }
    cta_sync();
    // attention backward in two phases, every accumulation in registers.
    // (1) thread = (head, query): dP_j = dO . v_j, dS_j = P_j (dP_j - sum_j' dP_j' P_j') / sqrt(dh), and
    //     dQ = sum_j dS_j k_j
    float* DS = S + sc.dscore;
    FOR_T(i, TH * T) {
      const int h = i / T, t = i % T;
      const float* pr = PR + (h * T + t) * T;
      const float4 dO0 = ld4(DATT + t * TD + h * TDH), dO1 = ld4(DATT + t * TD + h * TDH + 4);
      float dp[TMAXT];
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) {
        dp[j] = 0.f;
        if (j <= t) {
          const float* v = QKV + j * 3 * TD + 2 * TD + h * TDH;
          const float4 v0 = ld4(v), v1 = ld4(v + 4);
          const float acc = ((dO0.x * v0.x + dO0.y * v0.y) + (dO0.z * v0.z + dO0.w * v0.w)) +
                            ((dO1.x * v1.x + dO1.y * v1.y) + (dO1.z * v1.z + dO1.w * v1.w));
          dp[j] = acc;
          dot = fmaf(acc, pr[j], dot);
        }
      }
      float4 dq0 = make_float4(0.f, 0.f, 0.f, 0.f), dq1 = dq0;
      float* ds = DS + (h * T + t) * T;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) {
        if (j < T) {
          float d = 0.f;
          if (j <= t) {
            d = pr[j] * (dp[j] - dot) * inv_sqrt_dh;
            const float* k = QKV + j * 3 * TD + TD + h * TDH;
            fma4(dq0, d, ld4(k)); fma4(dq1, d, ld4(k + 4));
          }
          ds[j] = d;
        }
      }
      st4(DQKV + t * 3 * TD + h * TDH, dq0);
      st4(DQKV + t * 3 * TD + h * TDH + 4, dq1);
    }
    cta_sync();
    // (2) thread = (head, key): dK = sum_t dS_tj q_t, dV = sum_t P_tj dO_t over the queries t >= j
    FOR_T(i, TH * T) {
      const int h = i / T, j = i % T;
      float4 dk0 = make_float4(0.f, 0.f, 0.f, 0.f), dk1 = dk0, dv0 = dk0, dv1 = dk0;
      for (int t = j; t < T; ++t) {
        const float d = DS[(h * T + t) * T + j], p = PR[(h * T + t) * T + j];
        const float* q = QKV + t * 3 * TD + h * TDH;
        const float* dOp = DATT + t * TD + h * TDH;
        fma4(dk0, d, ld4(q)); fma4(dk1, d, ld4(q + 4));
        fma4(dv0, p, ld4(dOp)); fma4(dv1, p, ld4(dOp + 4));
      }
      st4(DQKV + j * 3 * TD + TD + h * TDH, dk0); st4(DQKV + j * 3 * TD + TD + h * TDH + 4, dk1);
      st4(DQKV + j * 3 * TD + 2 * TD + h * TDH, dv0); st4(DQKV + j * 3 * TD + 2 * TD + h * TDH + 4, dv1);
    }
    cta_sync();
    // Q/K/V projections: dW, and the block-input gradient DX <- DXIN + dQKV W
    prefetch(pf, Wl + lay.wqkv, 3 * TD * TD / 4);
    FOR_T(i, 3 * TD * (TD / 4)) {
      const int c = i / (TD / 4), f4 = i % (TD / 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < T; ++r) fma4(acc, DQKV[r * 3 * TD + c], ld4(X + r * TD + 4 * f4));
      st4(Gl + lay.wqkv + c * TD + 4 * f4, acc);
    }
    park(pf, 3 * TD * TD / 4);
    {   // thread = (row, four features, quarter of the 192 projection outputs)
      const int part = quad_part(tid), f4 = quad_f4(tid), r = quad_row(tid);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < T) {
#pragma unroll 12
        for (int c = part; c < 3 * TD; c += 4) fma4(acc, DQKV[r * 3 * TD + c], ld4(WS + c * TD + 4 * f4));
      }
      acc = quad_sum(acc);
      if (part == 0 && r < T) {
        const float4 in = ld4(DXIN + r * TD + 4 * f4);
        st4(DX + r * TD + 4 * f4, make_float4(in.x + acc.x, in.y + acc.y, in.z + acc.z, in.w + acc.w));
      }
    }
    cta_sync();
  }
  // embedding: through nn.Dropout, the sqrt(d) scale and the bayes mask into the rows of the table.  A token may occur
  // several times in a sequence: the thread of its FIRST occurrence sums all of them in position order (deterministic,
  // no atomics); rows of tokens the sequence does not contain stay zero.
  FOR_T(i, T * TD) {
    const int t = i / TD, f = i % TD;
    float g = drop1(a, row0 + t, TS_PE, (uint32_t)f, a.thr_train, a.p_train, DX[i]) * sqrt_d;
    DX[i] = drop1(a, row0 + t, TS_EMBED, (uint32_t)f, a.thr_bayes, a.p_bayes, g);
  }
  __syncthreads();
  FOR_T(i, T * TD) {
    const int t = i / TD, f = i % TD;
    const int tok = s_tok[t];
    bool first = true;
    for (int u = 0; u < t; ++u) first = first && s_tok[u] != tok;
    if (first) {
      float acc = 0.f;
      for (int u = t; u < T; ++u)
        if (s_tok[u] == tok) acc += DX[u * TD + f];
      G[lay.emb + tok * TD + f] = acc;
    }
  }
#undef FOR_T
}

// gradient = sum of the sequence slabs (in order); AdamW (torch.optim.AdamW: decoupled weight decay, bias-corrected
// moments); scatter back to the module's tensors.  block (segment, chunk)
__global__ void __launch_bounds__(256) train_adam_kernel(const TrainArgs a) {
  const TrainSeg sg = a.segs[blockIdx.x];
  const float* slab = a.scratch + a.gl.slab;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < sg.count; i += gridDim.y * blockDim.x) {
    const int k = sg.offset + i;
    float g = 0.f;
#pragma unroll 8
    for (int b = 0; b < a.B; ++b) g += __ldcs(slab + (size_t)b * a.slab_stride + k);
    a.grad[k] = g;
    const float m = a.beta1 * a.adam_m[k] + (1.0f - a.beta1) * g;
    const float v = a.beta2 * a.adam_v[k] + (1.0f - a.beta2) * g * g;
    a.adam_m[k] = m;
    a.adam_v[k] = v;
    float p = a.flat[k] * (1.0f - a.lr * a.weight_decay);
    const float denom = sqrtf(v) / sqrtf(a.bias2) + a.eps;
    p -= (a.lr / a.bias1) * (m / denom);
    sg.ptr[i] = p;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    float loss = 0.f;
    for (int b = 0; b < a.B; ++b) loss += a.scratch[a.gl.loss + b];
    *a.loss_out += loss;
  }
}

}  // namespace

}  // namespace bfb

using namespace bfb;

#ifdef BFB_TRAIN_TIMING
extern "C" int bfb200_debug_train_timing(unsigned long long* out128) {   // debug build only: globaltimer at every barrier of CTA 0 in the last step
  cudaDeviceSynchronize();
  int n = 0;
  cudaMemcpyFromSymbol(&n, bfb_train_dbg_n, sizeof(int));
  cudaMemcpyFromSymbol(out128, bfb_train_dbg, sizeof(unsigned long long) * 128);
  const int zero = 0;
  cudaMemcpyToSymbol(bfb_train_dbg_n, &zero, sizeof(int));
  return n;
}
#endif

extern "C" size_t bfb200_train_state_floats(int32_t ntoken, int32_t d_ff, int32_t n_layers) {
  if (ntoken < 1 || d_ff < 1 || n_layers < 1) return 0;
  return (size_t)train_layout(ntoken, d_ff, n_layers).total;
}

extern "C" size_t bfb200_train_scratch_floats(int32_t batch, int32_t seq_len, int32_t prompt_len, int32_t ntoken,
                                              int32_t d_ff, int32_t n_layers) {
  if (batch < 1 || seq_len < 2 || prompt_len < 1 || prompt_len >= seq_len || seq_len > TMAXT) return 0;
  if (ntoken < 2 || d_ff < 1 || n_layers < 1) return 0;
  if ((d_ff & 3) != 0 || d_ff * TD > TSTAGE) return 0;   // float4 alignment of the flat layout; W2^T is parked in the stage
  // 0 also when one sequence's activations do not fit a CTA's shared memory (the caller trains on another path)
  if ((size_t)(seq_scratch(seq_len, prompt_len, ntoken, d_ff, n_layers).total + 4 + TSTAGE) * sizeof(float) > (size_t)TSMEM_MAX) return 0;
  if (((ntoken + 3) & ~3) * TD > TSTAGE) return 0;   // the vocabulary head is parked in the same stage
  return (size_t)train_global(batch, ntoken, d_ff, n_layers, train_layout(ntoken, d_ff, n_layers).total).total;
}

extern "C" int bfb200_train_flat_offset(int32_t ntoken, int32_t d_ff, int32_t n_layers, int32_t layer, int32_t which) {
  const TrainLayout t = train_layout(ntoken, d_ff, n_layers);
  const int base = t.layer0 + layer * t.per_layer;
  switch (which) {
    case 0: return t.emb;
    case 1: return base + t.wqkv;
    case 2: return base + t.wout;
    case 3: return base + t.ln1w;
    case 4: return base + t.ln1b;
    case 5: return base + t.w1;
    case 6: return base + t.b1;
    case 7: return base + t.w2;
    case 8: return base + t.b2;
    case 9: return base + t.ln2w;
    case 10: return base + t.ln2b;
    case 11: return t.head_w;
    case 12: return t.head_b;
    default: return -1;
  }
}

extern "C" int bfb200_train_step(const bfb200_train_args* ta, void* stream) {
  BFB_REQUIRE(ta != nullptr, BFB200_ERR_INVALID_ARG, "args is NULL");
  BFB_REQUIRE(ta->d_model == TD && ta->n_heads == TH, BFB200_ERR_UNSUPPORTED,
              "fused train step serves d_model 64 with 8 heads (got d=%d H=%d)", ta->d_model, ta->n_heads);
  BFB_REQUIRE(ta->ntoken >= 2 && ta->d_ff >= 1 && ta->n_layers >= 1, BFB200_ERR_INVALID_ARG, "bad model dimensions");
  BFB_REQUIRE((ta->d_ff & 3) == 0 && ta->d_ff * TD <= TSTAGE, BFB200_ERR_UNSUPPORTED,
              "fused train step takes d_ff a multiple of 4 up to %d (got %d)", TSTAGE / TD, ta->d_ff);
  BFB_REQUIRE(ta->batch >= 1 && ta->prompt_len >= 1 && ta->seq_len > ta->prompt_len, BFB200_ERR_INVALID_ARG,
              "bad batch shape (B=%d, T=%d, P=%d)", ta->batch, ta->seq_len, ta->prompt_len);
  BFB_REQUIRE(ta->seq_len <= ta->pe_len, BFB200_ERR_INVALID_ARG, "sequence longer than the positional table");
  BFB_REQUIRE(ta->seq_len <= TMAXT, BFB200_ERR_UNSUPPORTED, "fused train step takes sequences of at most %d tokens (got %d)",
              TMAXT, ta->seq_len);
  BFB_REQUIRE(ta->n_segments >= 1 && ta->n_segments <= 65535, BFB200_ERR_INVALID_ARG, "bad segment count %d", ta->n_segments);
  BFB_REQUIRE(ta->step >= 1, BFB200_ERR_INVALID_ARG, "step counts from 1");
  BFB_REQUIRE(ta->p_train >= 0.f && ta->p_train < 1.f && ta->p_bayes >= 0.f && ta->p_bayes < 1.f, BFB200_ERR_INVALID_ARG,
              "dropout rates outside [0, 1)");
  BFB_DEV(ta->flat); BFB_DEV(ta->grad); BFB_DEV(ta->adam_m); BFB_DEV(ta->adam_v); BFB_DEV(ta->segments);
  BFB_DEV(ta->tokens); BFB_DEV(ta->pe); BFB_DEV(ta->scratch); BFB_DEV(ta->loss_out);
  TrainArgs a;
  memset(&a, 0, sizeof(a));
  a.flat = ta->flat; a.grad = ta->grad; a.adam_m = ta->adam_m; a.adam_v = ta->adam_v;
  a.segs = reinterpret_cast<const TrainSeg*>(ta->segments);
  a.n_segs = ta->n_segments;
  a.tokens = ta->tokens; a.pe = ta->pe; a.scratch = ta->scratch; a.loss_out = ta->loss_out;
  a.lay = train_layout(ta->ntoken, ta->d_ff, ta->n_layers);
  a.sc = seq_scratch(ta->seq_len, ta->prompt_len, ta->ntoken, ta->d_ff, ta->n_layers);
  a.gl = train_global(ta->batch, ta->ntoken, ta->d_ff, ta->n_layers, a.lay.total);
  a.slab_stride = (a.lay.total + 3) & ~3;
  const size_t smem = (size_t)(a.sc.total + 4 + TSTAGE) * sizeof(float);
  BFB_REQUIRE(smem <= (size_t)TSMEM_MAX && ((ta->ntoken + 3) & ~3) * TD <= TSTAGE, BFB200_ERR_UNSUPPORTED,
              "one sequence's activations (%zu bytes) do not fit a CTA's shared memory, or the vocabulary (%d) exceeds %d",
              smem, ta->ntoken, TSTAGE / TD);
  a.B = ta->batch; a.T = ta->seq_len; a.P = ta->prompt_len;
  a.lr = ta->lr; a.beta1 = ta->beta1; a.beta2 = ta->beta2; a.eps = ta->eps; a.weight_decay = ta->weight_decay;
  a.bias1 = 1.0f - powf(ta->beta1, (float)ta->step);
  a.bias2 = 1.0f - powf(ta->beta2, (float)ta->step);
  a.p_train = ta->p_train; a.p_bayes = ta->p_bayes;
  a.thr_train = (uint32_t)fmin(4294967295.0, floor((double)ta->p_train * 4294967296.0));
  a.thr_bayes = (uint32_t)fmin(4294967295.0, floor((double)ta->p_bayes * 4294967296.0));
  a.ln_eps = ta->ln_eps;
  a.key = make_uint2((uint32_t)ta->seed, (uint32_t)(ta->seed >> 32));
  a.step_id = (uint32_t)ta->step;
  cudaStream_t st = (cudaStream_t)stream;
  const cudaError_t ea = cudaFuncSetAttribute(train_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  BFB_REQUIRE(ea == cudaSuccess, BFB200_ERR_CUDA, "cudaFuncSetAttribute(train_seq_kernel): %s", cudaGetErrorString(ea));
  train_gather_kernel<<<dim3(a.n_segs, 8), 256, 0, st>>>(a);
  BFB_LAUNCH_CHECK("train_gather_kernel", st);
  train_transpose_kernel<<<64, 256, 0, st>>>(a);
  BFB_LAUNCH_CHECK("train_transpose_kernel", st);
  train_seq_kernel<<<a.B, TTHREADS, smem, st>>>(a);
  BFB_LAUNCH_CHECK("train_seq_kernel", st);
  train_adam_kernel<<<dim3(a.n_segs, 32), 256, 0, st>>>(a);   // 8,192 threads per segment: one element each for all but huge vocabularies
  BFB_LAUNCH_CHECK("train_adam_kernel", st);
  return BFB200_OK;
}

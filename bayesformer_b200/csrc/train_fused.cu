// Fused fine-tune step of the notebook-default model class: forward, cross-entropy loss, backward and the AdamW
// update of ONE batch in ONE kernel (reference loop: src/libs/utils.py:68-78 — model(train batch), CrossEntropyLoss on
// the log-softmax outputs of the answer positions, loss.backward(), AdamW.step()).
//
// Why: the model has 50 k parameters and a batch is 20 sequences x 9 tokens = 180 token rows; the reference step is
// ~600 PyTorch kernels of a few microseconds each (launch-bound: 9.4 ms eager, 1.9 ms as a CUDA graph).  Here one
// cluster of 8 CTAs x 1024 threads walks the ~45 phases of the step with a hardware cluster barrier in between; every
// activation lives in an L2-resident scratch buffer (~1 MB), every phase is a cluster-stride loop over its output
// elements.  fp32 CUDA-core arithmetic throughout (21 M multiply-adds per step: nothing for tensor cores to win here).
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
// parameter tensors — 24 per-head projection matrices per layer among them — to one flat layout; the kernel gathers
// them at the start and scatters the updated values back at the end.
#include "common.cuh"
#include "kernels.cuh"

namespace bfb {

namespace {

constexpr int TD = 64, TH = 8, TDH = 8;
constexpr int TMAXT = 16;   // longest training sequence (prompt + answer + EOS): per-query score rows live in registers
constexpr int TCLUSTER = 8, TTHREADS = 1024;

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

struct TrainScratch {   // offsets in floats inside the scratch buffer (R = B * T rows)
  int x;        // (L + 1) x R x 64 : block inputs, x[L] = last block's output
  int qkv;      // L x R x 192
  int prob;     // L x B x H x T x T
  int att;      // L x R x 64   concatenated head outputs
  int v1;       // L x R x 64   x + dropout(attn W_out^T): LayerNorm1 input
  int h1;       // L x R x 64   LayerNorm1 output
  int u;        // L x R x F    FFN pre-activation
  int y;        // L x R x 64   x + dropout(FFN): LayerNorm2 input
  int stat;     // L x R x 4    mean1, rstd1, mean2, rstd2
  int logit;    // NP x V       logits -> d logits
  int dx;       // R x 64       gradient w.r.t. the current block output
  int dxin;     // R x 64       gradient w.r.t. the current block input (accumulated)
  int dtmp;     // R x 64       d(attn branch) / d(ffn branch)
  int dqkv;     // R x 192
  int du;       // R x F
  int datt;     // R x 64
  int dscore;   // B x H x T x T   d(scores)
  int wt;       // transposed copies for the forward products: per layer W_qkv^T (64 x 192), W_out^T (64 x 64),
                // W1^T (64 x F); then head^T (64 x Vp), Vp = V rounded up to 4 (a thread owns 4 adjacent outputs and
                // reads them as one coalesced float4 per k)
  int wt_layer, wt_out, wt_w1, wt_head, vp;
  int total;
};

__host__ __device__ inline TrainScratch train_scratch(int B, int T, int P, int V, int F, int L) {
  const int R = B * T, NP = B * (T - P);
  TrainScratch s;
  int o = 0;
  s.x = o; o += (L + 1) * R * TD;
  s.qkv = o; o += L * R * 3 * TD;
  s.prob = o; o += L * B * TH * T * T;
  s.att = o; o += L * R * TD;
  s.v1 = o; o += L * R * TD;
  s.h1 = o; o += L * R * TD;
  s.u = o; o += L * R * F;
  s.y = o; o += L * R * TD;
  s.stat = o; o += L * R * 4;
  s.logit = o; o += NP * V;
  s.dx = o; o += R * TD;
  s.dxin = o; o += R * TD;
  s.dtmp = o; o += R * TD;
  s.dqkv = o; o += R * 3 * TD;
  s.du = o; o += R * F;
  s.datt = o; o += R * TD;
  s.dscore = o; o += B * TH * T * T;
  o = (o + 3) & ~3;
  s.vp = (V + 3) & ~3;
  s.wt = o;
  s.wt_out = TD * 3 * TD;
  s.wt_w1 = s.wt_out + TD * TD;
  s.wt_layer = (s.wt_w1 + TD * F + 3) & ~3;
  s.wt_head = L * s.wt_layer;
  o += s.wt_head + TD * s.vp;
  s.total = (o + 3) & ~3;
  return s;
}

struct TrainSeg { int offset, count; float* ptr; };   // flat [offset, offset + count) <-> ptr[0 .. count)

struct TrainArgs {
  float* flat;          // parameters (gathered), total floats
  float* grad;          // gradients, total floats (zeroed by the kernel)
  float* adam_m;        // AdamW first / second moments, total floats each (persist between steps)
  float* adam_v;
  const TrainSeg* segs;
  int n_segs;
  const int64_t* tokens;   // (T, B) seq-first: prompts followed by answers (utils.py:61-66)
  const float* pe;         // (pe_len, 64)
  float* scratch;
  float* loss_out;         // += mean loss of this batch
  TrainLayout lay;
  TrainScratch sc;
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

#ifdef BFB_TRAIN_TIMING
__device__ unsigned long long bfb_train_dbg[128];
__device__ int bfb_train_dbg_n;
#endif
__device__ __forceinline__ void cluster_sync_all() {   // release / acquire at cluster scope: orders the global-memory traffic
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#ifdef BFB_TRAIN_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int k = bfb_train_dbg_n;
    if (k < 128) { bfb_train_dbg[k] = t; bfb_train_dbg_n = k + 1; }
  }
#endif
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__global__ void __launch_bounds__(TTHREADS, 1) train_step_kernel(const TrainArgs a) {
  const int gsize = TCLUSTER * TTHREADS;
  const int gtid = (int)cluster_ctarank() * TTHREADS + (int)threadIdx.x;
  const TrainLayout& lay = a.lay;
  const TrainScratch& sc = a.sc;
  const int B = a.B, T = a.T, P = a.P, R = B * T, V = lay.V, F = lay.F, L = lay.L;
  const int NP = B * (T - P);        // predicted positions: rows t in [P - 1, T - 2] (utils.py:72-76)
  float* S = a.scratch;
  float* W = a.flat;
  float* G = a.grad;
  const float inv_sqrt_dh = rsqrtf((float)TDH);
  const float sqrt_d = sqrtf((float)TD);
  // row r = b * T + t
#define FOR_ALL(i, n) for (int i = gtid; i < (n); i += gsize)

  // ---- gather the parameters, zero the gradients ------------------------------------------------------------
  for (int sgi = 0; sgi < a.n_segs; ++sgi) {
    const TrainSeg sg = a.segs[sgi];
    FOR_ALL(i, sg.count) W[sg.offset + i] = sg.ptr[i];
  }
  FOR_ALL(i, lay.total) G[i] = 0.f;
  cluster_sync_all();
  const int gwarp = gtid >> 5, lane = gtid & 31, nwarps = gsize >> 5;
  {   // transposed weight copies for the forward products (reads strided, writes coalesced; 30 k elements)
    float* WT = S + sc.wt;
    for (int l = 0; l < L; ++l) {
      const float* Wl = W + lay.layer0 + l * lay.per_layer;
      float* T0 = WT + l * sc.wt_layer;
      FOR_ALL(i, TD * 3 * TD) { const int k = i / (3 * TD), c = i % (3 * TD); T0[i] = Wl[lay.wqkv + c * TD + k]; }
      FOR_ALL(i, TD * TD) { const int k = i / TD, n = i % TD; T0[sc.wt_out + i] = Wl[lay.wout + n * TD + k]; }
      FOR_ALL(i, TD * F) { const int k = i / F, m = i % F; T0[sc.wt_w1 + i] = Wl[lay.w1 + m * TD + k]; }
    }
    FOR_ALL(i, TD * sc.vp) { const int k = i / sc.vp, v = i % sc.vp; WT[sc.wt_head + i] = v < V ? W[lay.head_w + v * TD + k] : 0.f; }
  }
  cluster_sync_all();

  // ---- forward ----------------------------------------------------------------------------------------------
  // embedding, bayes dropout E, sqrt(d) scale, positional encoding, nn.Dropout (transformer.py:355-362, :269-272)
  FOR_ALL(i, R * (TD / 4)) {
    const int r = i / (TD / 4), g4 = i % (TD / 4), b = r / T, t = r % T;
    const int tok = (int)a.tokens[(int64_t)t * B + b];
    const float4 e = *reinterpret_cast<const float4*>(W + lay.emb + tok * TD + 4 * g4);
    const float4 pe = *reinterpret_cast<const float4*>(a.pe + t * TD + 4 * g4);
    float v[4] = {e.x, e.y, e.z, e.w};
    const float pv[4] = {pe.x, pe.y, pe.z, pe.w};
    const uint32_t kE = a.p_bayes > 0.f ? keep4_train(a, (uint32_t)r, TS_EMBED, (uint32_t)g4, a.thr_bayes) : 15u;
    const uint32_t kP = a.p_train > 0.f ? keep4_train(a, (uint32_t)r, TS_PE, (uint32_t)g4, a.thr_train) : 15u;
    const float sE = a.p_bayes > 0.f ? 1.0f / (1.0f - a.p_bayes) : 1.0f, sP = a.p_train > 0.f ? 1.0f / (1.0f - a.p_train) : 1.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float x = ((kE >> q) & 1u) ? v[q] * sE : 0.f;
      x = x * sqrt_d + pv[q];
      v[q] = ((kP >> q) & 1u) ? x * sP : 0.f;
    }
    *reinterpret_cast<float4*>(S + sc.x + r * TD + 4 * g4) = make_float4(v[0], v[1], v[2], v[3]);
  }
  cluster_sync_all();

  for (int l = 0; l < L; ++l) {
    const float* Wl = W + lay.layer0 + l * lay.per_layer;
    const float* X = S + sc.x + l * R * TD;
    float* QKV = S + sc.qkv + l * R * 3 * TD;
    float* PR = S + sc.prob + l * B * TH * T * T;
    float* ATT = S + sc.att + l * R * TD;
    float* V1 = S + sc.v1 + l * R * TD;
    float* H1 = S + sc.h1 + l * R * TD;
    float* U = S + sc.u + l * R * F;
    float* Y = S + sc.y + l * R * TD;
    float* ST = S + sc.stat + l * R * 4;
    float* XN = S + sc.x + (l + 1) * R * TD;
    const uint32_t site0 = 4u * (uint32_t)l;
    // Q, K, V of all heads (bias-free, transformer.py:54-57)
    const float* WTl = S + sc.wt + l * sc.wt_layer;
    FOR_ALL(i, R * (3 * TD / 4)) {   // thread = (row, four adjacent outputs)
      const int r = i / (3 * TD / 4), c4 = i % (3 * TD / 4);
      const float* xr = X + r * TD;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16
      for (int k = 0; k < TD; ++k) {
        const float xv = xr[k];
        const float4 w = *reinterpret_cast<const float4*>(WTl + k * 3 * TD + 4 * c4);
        acc.x = fmaf(xv, w.x, acc.x); acc.y = fmaf(xv, w.y, acc.y); acc.z = fmaf(xv, w.z, acc.z); acc.w = fmaf(xv, w.w, acc.w);
      }
      *reinterpret_cast<float4*>(QKV + r * 3 * TD + 4 * c4) = acc;
    }
    cluster_sync_all();
    // causal softmax(QK^T / sqrt(dh)) V, thread = (sequence, head, query) (transformer.py:58-65); scores in registers
    FOR_ALL(i, B * TH * T) {
      const int b = i / (TH * T), h = (i / T) % TH, t = i % T;
      const float* qp = QKV + (b * T + t) * 3 * TD + h * TDH;
      const float4 q0 = *reinterpret_cast<const float4*>(qp), q1 = *reinterpret_cast<const float4*>(qp + 4);
      float sc[TMAXT];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) {
        sc[j] = -INFINITY;
        if (j <= t) {
          const float* kp = QKV + (b * T + j) * 3 * TD + TD + h * TDH;
          const float4 k0 = *reinterpret_cast<const float4*>(kp), k1 = *reinterpret_cast<const float4*>(kp + 4);
          const float s = ((q0.x * k0.x + q0.y * k0.y) + (q0.z * k0.z + q0.w * k0.w)) +
                          ((q1.x * k1.x + q1.y * k1.y) + (q1.z * k1.z + q1.w * k1.w));
          sc[j] = s * inv_sqrt_dh;
          mx = fmaxf(mx, sc[j]);
        }
      }
      float z = 0.f;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) { sc[j] = j <= t ? __expf(sc[j] - mx) : 0.f; z += sc[j]; }
      const float inv_z = 1.0f / z;
      float o[TDH] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float* pr = PR + ((b * TH + h) * T + t) * T;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) {
        if (j < T) {
          const float p = sc[j] * inv_z;
          pr[j] = p;
          if (j <= t) {
            const float* vp = QKV + (b * T + j) * 3 * TD + 2 * TD + h * TDH;
            const float4 v0 = *reinterpret_cast<const float4*>(vp), v1 = *reinterpret_cast<const float4*>(vp + 4);
            o[0] = fmaf(p, v0.x, o[0]); o[1] = fmaf(p, v0.y, o[1]); o[2] = fmaf(p, v0.z, o[2]); o[3] = fmaf(p, v0.w, o[3]);
            o[4] = fmaf(p, v1.x, o[4]); o[5] = fmaf(p, v1.y, o[5]); o[6] = fmaf(p, v1.z, o[6]); o[7] = fmaf(p, v1.w, o[7]);
          }
        }
      }
      float* dst = ATT + (b * T + t) * TD + h * TDH;
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    cluster_sync_all();
    // W_out, nn.Dropout, residual from the block input (transformer.py:125, :231)
    FOR_ALL(i, R * (TD / 4)) {
      const int r = i / (TD / 4), n4 = i % (TD / 4);
      const float* ar = ATT + r * TD;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 16
      for (int k = 0; k < TD; ++k) {
        const float av = ar[k];
        const float4 w = *reinterpret_cast<const float4*>(WTl + sc.wt_out + k * TD + 4 * n4);
        acc[0] = fmaf(av, w.x, acc[0]); acc[1] = fmaf(av, w.y, acc[1]); acc[2] = fmaf(av, w.z, acc[2]); acc[3] = fmaf(av, w.w, acc[3]);
      }
      const uint32_t kk = a.p_train > 0.f ? keep4_train(a, (uint32_t)r, site0 + TS_ATTN, (uint32_t)n4, a.thr_train) : 15u;
      const float sd = a.p_train > 0.f ? 1.0f / (1.0f - a.p_train) : 1.0f;
      const float4 xv = *reinterpret_cast<const float4*>(X + r * TD + 4 * n4);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      float o4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) o4[q] = xs[q] + (((kk >> q) & 1u) ? acc[q] * sd : 0.f);
      *reinterpret_cast<float4*>(V1 + r * TD + 4 * n4) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
    cluster_sync_all();
    // LayerNorm1 (thread per row)
    for (int r = gwarp; r < R; r += nwarps) {   // warp per row, two features per lane
      const float v0 = V1[r * TD + lane], v1 = V1[r * TD + 32 + lane];
      const float mean = warp_sum(v0 + v1) * (1.0f / TD);
      const float c0 = v0 - mean, c1 = v1 - mean;
      const float rstd = rsqrtf(warp_sum(c0 * c0 + c1 * c1) * (1.0f / TD) + a.ln_eps);
      if (lane == 0) { ST[r * 4 + 0] = mean; ST[r * 4 + 1] = rstd; }
      H1[r * TD + lane] = c0 * rstd * Wl[lay.ln1w + lane] + Wl[lay.ln1b + lane];
      H1[r * TD + 32 + lane] = c1 * rstd * Wl[lay.ln1w + 32 + lane] + Wl[lay.ln1b + 32 + lane];
    }
    cluster_sync_all();
    // FFN: Linear(d -> F) (pre-activation kept for the backward pass), transformer.py:151-155
    FOR_ALL(i, R * F) {
      const int r = i / F, m = i % F;
      const float* hr = H1 + r * TD;
      float acc = Wl[lay.b1 + m];
#pragma unroll 16
      for (int k = 0; k < TD; ++k) acc = fmaf(hr[k], WTl[sc.wt_w1 + k * F + m], acc);
      U[i] = acc;
    }
    cluster_sync_all();
    // Linear(F -> d) on ReLU, nn.Dropout, residual from the block input (transformer.py:233)
    FOR_ALL(i, R * TD) {
      const int r = i / TD, n = i % TD;
      float acc = Wl[lay.b2 + n];
      for (int m = 0; m < F; ++m) acc = fmaf(fmaxf(U[r * F + m], 0.f), Wl[lay.w2 + n * F + m], acc);
      Y[i] = X[i] + drop1(a, (uint32_t)r, site0 + TS_FFN, (uint32_t)n, a.thr_train, a.p_train, acc);
    }
    cluster_sync_all();
    // LayerNorm2 + bayes layer-output dropout (transformer.py:233, :370)
    for (int r = gwarp; r < R; r += nwarps) {
      const float v0 = Y[r * TD + lane], v1 = Y[r * TD + 32 + lane];
      const float mean = warp_sum(v0 + v1) * (1.0f / TD);
      const float c0 = v0 - mean, c1 = v1 - mean;
      const float rstd = rsqrtf(warp_sum(c0 * c0 + c1 * c1) * (1.0f / TD) + a.ln_eps);
      if (lane == 0) { ST[r * 4 + 2] = mean; ST[r * 4 + 3] = rstd; }
      const float z0 = c0 * rstd * Wl[lay.ln2w + lane] + Wl[lay.ln2b + lane];
      const float z1 = c1 * rstd * Wl[lay.ln2w + 32 + lane] + Wl[lay.ln2b + 32 + lane];
      XN[r * TD + lane] = drop1(a, (uint32_t)r, site0 + TS_LAYER_OUT, (uint32_t)lane, a.thr_bayes, a.p_bayes, z0);
      XN[r * TD + 32 + lane] = drop1(a, (uint32_t)r, site0 + TS_LAYER_OUT, (uint32_t)(32 + lane), a.thr_bayes, a.p_bayes, z1);
    }
    cluster_sync_all();
  }

  // vocabulary head on the predicting positions (transformer.py:374), pred row q = b * (T - P) + (t - (P - 1))
  const float* XL = S + sc.x + L * R * TD;
  float* LG = S + sc.logit;
  FOR_ALL(i, NP * (sc.vp / 4)) {
    const int q = i / (sc.vp / 4), v4 = i % (sc.vp / 4), b = q / (T - P), t = P - 1 + q % (T - P);
    const float* xr = XL + (b * T + t) * TD;
    const float* HT = S + sc.wt + sc.wt_head;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 16
    for (int k = 0; k < TD; ++k) {
      const float xv = xr[k];
      const float4 w = *reinterpret_cast<const float4*>(HT + k * sc.vp + 4 * v4);
      acc[0] = fmaf(xv, w.x, acc[0]); acc[1] = fmaf(xv, w.y, acc[1]); acc[2] = fmaf(xv, w.z, acc[2]); acc[3] = fmaf(xv, w.w, acc[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (4 * v4 + e < V) LG[q * V + 4 * v4 + e] = acc[e] + W[lay.head_b + 4 * v4 + e];
  }
  cluster_sync_all();
  // log-softmax, NLL of the next token, d logits = (softmax - onehot) / NP   (utils.py:72-76; transformer.py:375)
  for (int q = gwarp; q < NP; q += nwarps) {   // warp per predicted position
    float* lg = LG + q * V;
    const int b = q / (T - P), t = P - 1 + q % (T - P);
    const int target = (int)a.tokens[(int64_t)(t + 1) * B + b];
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, lg[v]);
    mx = warp_max(mx);
    float z = 0.f;
    for (int v = lane; v < V; v += 32) z += __expf(lg[v] - mx);
    const float lse = mx + __logf(warp_sum(z));
    if (lane == 0) atomicAdd(a.loss_out, (lse - lg[target]) * (1.0f / NP));
    __syncwarp();
    for (int v = lane; v < V; v += 32) lg[v] = (__expf(lg[v] - lse) - (v == target ? 1.f : 0.f)) * (1.0f / NP);
  }
  cluster_sync_all();

  // ---- backward ---------------------------------------------------------------------------------------------
  float* DX = S + sc.dx;
  float* DXIN = S + sc.dxin;
  float* DT = S + sc.dtmp;
  float* DQKV = S + sc.dqkv;
  float* DU = S + sc.du;
  float* DATT = S + sc.datt;
  // head: dW, db, dx_L
  FOR_ALL(i, V * (TD / 4)) {
    const int v = i / (TD / 4), f4 = i % (TD / 4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int q = 0; q < NP; ++q) {
      const int b = q / (T - P), t = P - 1 + q % (T - P);
      const float d = LG[q * V + v];
      const float4 x = *reinterpret_cast<const float4*>(XL + (b * T + t) * TD + 4 * f4);
      acc.x = fmaf(d, x.x, acc.x); acc.y = fmaf(d, x.y, acc.y); acc.z = fmaf(d, x.z, acc.z); acc.w = fmaf(d, x.w, acc.w);
    }
    *reinterpret_cast<float4*>(G + lay.head_w + v * TD + 4 * f4) = acc;
  }
  FOR_ALL(v, V) {
    float acc = 0.f;
#pragma unroll 8
    for (int q = 0; q < NP; ++q) acc += LG[q * V + v];
    G[lay.head_b + v] = acc;
  }
  FOR_ALL(i, R * (TD / 4)) {
    const int r = i / (TD / 4), f4 = i % (TD / 4), b = r / T, t = r % T;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= P - 1 && t <= T - 2) {
      const float* dl = LG + (b * (T - P) + t - (P - 1)) * V;
#pragma unroll 4
      for (int v = 0; v < V; ++v) {
        const float d = dl[v];
        const float4 w = *reinterpret_cast<const float4*>(W + lay.head_w + v * TD + 4 * f4);
        acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(DX + r * TD + 4 * f4) = acc;
  }
  cluster_sync_all();

  for (int l = L - 1; l >= 0; --l) {
    const float* Wl = W + lay.layer0 + l * lay.per_layer;
    float* Gl = G + lay.layer0 + l * lay.per_layer;
    const float* X = S + sc.x + l * R * TD;
    const float* QKV = S + sc.qkv + l * R * 3 * TD;
    const float* PR = S + sc.prob + l * B * TH * T * T;
    const float* ATT = S + sc.att + l * R * TD;
    const float* V1 = S + sc.v1 + l * R * TD;
    const float* H1 = S + sc.h1 + l * R * TD;
    const float* U = S + sc.u + l * R * F;
    const float* Y = S + sc.y + l * R * TD;
    const float* ST = S + sc.stat + l * R * 4;
    const uint32_t site0 = 4u * (uint32_t)l;
    // bayes layer-output dropout backward: DX <- DX * mask (in place), then LayerNorm2 parameter gradients
    FOR_ALL(i, R * TD) {
      const int r = i / TD, f = i % TD;
      DX[i] = drop1(a, (uint32_t)r, site0 + TS_LAYER_OUT, (uint32_t)f, a.thr_bayes, a.p_bayes, DX[i]);
    }
    cluster_sync_all();
    FOR_ALL(f, TD) {   // d gamma2, d beta2
      float dg = 0.f, db = 0.f;
#pragma unroll 8
      for (int r = 0; r < R; ++r) {
        const float yh = (Y[r * TD + f] - ST[r * 4 + 2]) * ST[r * 4 + 3];
        dg = fmaf(DX[r * TD + f], yh, dg);
        db += DX[r * TD + f];
      }
      Gl[lay.ln2w + f] = dg; Gl[lay.ln2b + f] = db;
    }
    // LayerNorm2 input gradient (thread per row): dy -> DXIN (residual to the block input) and, through the FFN
    // dropout mask, the FFN branch gradient -> DT
    for (int r = gwarp; r < R; r += nwarps) {   // warp per row
      const float mean = ST[r * 4 + 2], rstd = ST[r * 4 + 3];
      const float dz0 = DX[r * TD + lane] * Wl[lay.ln2w + lane], dz1 = DX[r * TD + 32 + lane] * Wl[lay.ln2w + 32 + lane];
      const float yh0 = (Y[r * TD + lane] - mean) * rstd, yh1 = (Y[r * TD + 32 + lane] - mean) * rstd;
      const float s1 = warp_sum(dz0 + dz1) * (1.0f / TD), s2 = warp_sum(dz0 * yh0 + dz1 * yh1) * (1.0f / TD);
      const float dy0 = rstd * (dz0 - s1 - yh0 * s2), dy1 = rstd * (dz1 - s1 - yh1 * s2);
      DXIN[r * TD + lane] = dy0;
      DXIN[r * TD + 32 + lane] = dy1;
      DT[r * TD + lane] = drop1(a, (uint32_t)r, site0 + TS_FFN, (uint32_t)lane, a.thr_train, a.p_train, dy0);
      DT[r * TD + 32 + lane] = drop1(a, (uint32_t)r, site0 + TS_FFN, (uint32_t)(32 + lane), a.thr_train, a.p_train, dy1);
    }
    cluster_sync_all();
    // FFN second layer: dW2, db2, d(pre-activation)
    FOR_ALL(i, TD * F) {
      const int n = i / F, m = i % F;
      float acc = 0.f;
#pragma unroll 8
      for (int r = 0; r < R; ++r) acc = fmaf(DT[r * TD + n], fmaxf(U[r * F + m], 0.f), acc);
      Gl[lay.w2 + i] = acc;
    }
    FOR_ALL(n, TD) {
      float acc = 0.f;
#pragma unroll 8
      for (int r = 0; r < R; ++r) acc += DT[r * TD + n];
      Gl[lay.b2 + n] = acc;
    }
    FOR_ALL(i, R * F) {
      const int r = i / F, m = i % F;
      float acc = 0.f;
      for (int n = 0; n < TD; ++n) acc = fmaf(DT[r * TD + n], Wl[lay.w2 + n * F + m], acc);
      DU[i] = U[i] > 0.f ? acc : 0.f;
    }
    cluster_sync_all();
    // FFN first layer: dW1, db1, d(LayerNorm1 output) -> DT
    FOR_ALL(i, F * (TD / 4)) {
      const int m = i / (TD / 4), f4 = i % (TD / 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int r = 0; r < R; ++r) {
        const float d = DU[r * F + m];
        const float4 x = *reinterpret_cast<const float4*>(H1 + r * TD + 4 * f4);
        acc.x = fmaf(d, x.x, acc.x); acc.y = fmaf(d, x.y, acc.y); acc.z = fmaf(d, x.z, acc.z); acc.w = fmaf(d, x.w, acc.w);
      }
      *reinterpret_cast<float4*>(Gl + lay.w1 + m * TD + 4 * f4) = acc;
    }
    FOR_ALL(m, F) {
      float acc = 0.f;
#pragma unroll 8
      for (int r = 0; r < R; ++r) acc += DU[r * F + m];
      Gl[lay.b1 + m] = acc;
    }
    FOR_ALL(i, R * TD) {
      const int r = i / TD, f = i % TD;
      float acc = 0.f;
      for (int m = 0; m < F; ++m) acc = fmaf(DU[r * F + m], Wl[lay.w1 + m * TD + f], acc);
      DT[i] = acc;
    }
    cluster_sync_all();
    // LayerNorm1: parameter gradients, input gradient -> DXIN += dv1, attention-branch gradient (through its dropout) -> DX
    FOR_ALL(f, TD) {
      float dg = 0.f, db = 0.f;
#pragma unroll 8
      for (int r = 0; r < R; ++r) {
        const float vh = (V1[r * TD + f] - ST[r * 4 + 0]) * ST[r * 4 + 1];
        dg = fmaf(DT[r * TD + f], vh, dg);
        db += DT[r * TD + f];
      }
      Gl[lay.ln1w + f] = dg; Gl[lay.ln1b + f] = db;
    }
    for (int r = gwarp; r < R; r += nwarps) {
      const float mean = ST[r * 4 + 0], rstd = ST[r * 4 + 1];
      const float dz0 = DT[r * TD + lane] * Wl[lay.ln1w + lane], dz1 = DT[r * TD + 32 + lane] * Wl[lay.ln1w + 32 + lane];
      const float vh0 = (V1[r * TD + lane] - mean) * rstd, vh1 = (V1[r * TD + 32 + lane] - mean) * rstd;
      const float s1 = warp_sum(dz0 + dz1) * (1.0f / TD), s2 = warp_sum(dz0 * vh0 + dz1 * vh1) * (1.0f / TD);
      const float dv0 = rstd * (dz0 - s1 - vh0 * s2), dv1 = rstd * (dz1 - s1 - vh1 * s2);
      DXIN[r * TD + lane] += dv0;
      DXIN[r * TD + 32 + lane] += dv1;
      DX[r * TD + lane] = drop1(a, (uint32_t)r, site0 + TS_ATTN, (uint32_t)lane, a.thr_train, a.p_train, dv0);
      DX[r * TD + 32 + lane] = drop1(a, (uint32_t)r, site0 + TS_ATTN, (uint32_t)(32 + lane), a.thr_train, a.p_train, dv1);
    }
    cluster_sync_all();
    // W_out: dW, d(concatenated head outputs) -> DATT
    FOR_ALL(i, TD * (TD / 4)) {
      const int n = i / (TD / 4), k4 = i % (TD / 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int r = 0; r < R; ++r) {
        const float d = DX[r * TD + n];
        const float4 x = *reinterpret_cast<const float4*>(ATT + r * TD + 4 * k4);
        acc.x = fmaf(d, x.x, acc.x); acc.y = fmaf(d, x.y, acc.y); acc.z = fmaf(d, x.z, acc.z); acc.w = fmaf(d, x.w, acc.w);
      }
      *reinterpret_cast<float4*>(Gl + lay.wout + n * TD + 4 * k4) = acc;
    }
    FOR_ALL(i, R * (TD / 4)) {
      const int r = i / (TD / 4), k4 = i % (TD / 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int n = 0; n < TD; ++n) {
        const float d = DX[r * TD + n];
        const float4 w = *reinterpret_cast<const float4*>(Wl + lay.wout + n * TD + 4 * k4);
        acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
      }
      *reinterpret_cast<float4*>(DATT + r * TD + 4 * k4) = acc;
    }
    cluster_sync_all();
    // attention backward in two phases, every accumulation in registers.
    // (1) thread = (sequence, head, query): dP_j = dO . v_j, dS_j = P_j (dP_j - sum_j' dP_j' P_j') / sqrt(dh), and
    //     dQ = sum_j dS_j k_j
    float* DS = S + sc.dscore;
    FOR_ALL(i, B * TH * T) {
      const int b = i / (TH * T), h = (i / T) % TH, t = i % T;
      const float* pr = PR + ((b * TH + h) * T + t) * T;
      const float* dOp = DATT + (b * T + t) * TD + h * TDH;
      float dO[TDH];
#pragma unroll
      for (int e = 0; e < TDH; ++e) dO[e] = dOp[e];
      float dp[TMAXT];
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) {
        dp[j] = 0.f;
        if (j <= t) {
          const float* v = QKV + (b * T + j) * 3 * TD + 2 * TD + h * TDH;
          float acc = 0.f;
#pragma unroll
          for (int e = 0; e < TDH; ++e) acc = fmaf(dO[e], v[e], acc);
          dp[j] = acc;
          dot = fmaf(acc, pr[j], dot);
        }
      }
      float dq[TDH] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float* ds = DS + ((b * TH + h) * T + t) * T;
#pragma unroll
      for (int j = 0; j < TMAXT; ++j) {
        if (j < T) {
          float d = 0.f;
          if (j <= t) {
            d = pr[j] * (dp[j] - dot) * inv_sqrt_dh;
            const float* k = QKV + (b * T + j) * 3 * TD + TD + h * TDH;
#pragma unroll
            for (int e = 0; e < TDH; ++e) dq[e] = fmaf(d, k[e], dq[e]);
          }
          ds[j] = d;
        }
      }
      float* dqo = DQKV + (b * T + t) * 3 * TD + h * TDH;
#pragma unroll
      for (int e = 0; e < TDH; ++e) dqo[e] = dq[e];
    }
    cluster_sync_all();
    // (2) thread = (sequence, head, key): dK = sum_t dS_tj q_t, dV = sum_t P_tj dO_t over the queries t >= j
    FOR_ALL(i, B * TH * T) {
      const int b = i / (TH * T), h = (i / T) % TH, j = i % T;
      float dk[TDH] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dv[TDH] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int t = j; t < T; ++t) {
        const float d = DS[((b * TH + h) * T + t) * T + j], p = PR[((b * TH + h) * T + t) * T + j];
        const float* q = QKV + (b * T + t) * 3 * TD + h * TDH;
        const float* dOp = DATT + (b * T + t) * TD + h * TDH;
#pragma unroll
        for (int e = 0; e < TDH; ++e) { dk[e] = fmaf(d, q[e], dk[e]); dv[e] = fmaf(p, dOp[e], dv[e]); }
      }
      float* dko = DQKV + (b * T + j) * 3 * TD + TD + h * TDH;
      float* dvo = DQKV + (b * T + j) * 3 * TD + 2 * TD + h * TDH;
#pragma unroll
      for (int e = 0; e < TDH; ++e) { dko[e] = dk[e]; dvo[e] = dv[e]; }
    }
    cluster_sync_all();
    // Q/K/V projections: dW, and the block-input gradient DX <- DXIN + dQKV W
    FOR_ALL(i, 3 * TD * (TD / 4)) {
      const int c = i / (TD / 4), f4 = i % (TD / 4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int r = 0; r < R; ++r) {
        const float d = DQKV[r * 3 * TD + c];
        const float4 x = *reinterpret_cast<const float4*>(X + r * TD + 4 * f4);
        acc.x = fmaf(d, x.x, acc.x); acc.y = fmaf(d, x.y, acc.y); acc.z = fmaf(d, x.z, acc.z); acc.w = fmaf(d, x.w, acc.w);
      }
      *reinterpret_cast<float4*>(Gl + lay.wqkv + c * TD + 4 * f4) = acc;
    }
    FOR_ALL(i, R * (TD / 4)) {
      const int r = i / (TD / 4), f4 = i % (TD / 4);
      float4 acc = *reinterpret_cast<const float4*>(DXIN + r * TD + 4 * f4);
#pragma unroll 8
      for (int c = 0; c < 3 * TD; ++c) {
        const float d = DQKV[r * 3 * TD + c];
        const float4 w = *reinterpret_cast<const float4*>(Wl + lay.wqkv + c * TD + 4 * f4);
        acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
      }
      *reinterpret_cast<float4*>(DX + r * TD + 4 * f4) = acc;
    }
    cluster_sync_all();
  }
  // embedding: through nn.Dropout, the sqrt(d) scale and the bayes mask into the rows of the table
  FOR_ALL(i, R * TD) {
    const int r = i / TD, f = i % TD, b = r / T, t = r % T;
    float g = drop1(a, (uint32_t)r, TS_PE, (uint32_t)f, a.thr_train, a.p_train, DX[i]) * sqrt_d;
    g = drop1(a, (uint32_t)r, TS_EMBED, (uint32_t)f, a.thr_bayes, a.p_bayes, g);
    const int tok = (int)a.tokens[(int64_t)t * B + b];
    if (g != 0.f) atomicAdd(G + lay.emb + tok * TD + f, g);
  }
  cluster_sync_all();

  // ---- AdamW (torch.optim.AdamW: decoupled weight decay, bias-corrected moments) and scatter back --------------
  for (int sgi = 0; sgi < a.n_segs; ++sgi) {
    const TrainSeg sg = a.segs[sgi];
    FOR_ALL(i, sg.count) {
      const int k = sg.offset + i;
      const float g = G[k];
      const float m = a.beta1 * a.adam_m[k] + (1.0f - a.beta1) * g;
      const float v = a.beta2 * a.adam_v[k] + (1.0f - a.beta2) * g * g;
      a.adam_m[k] = m;
      a.adam_v[k] = v;
      float p = W[k] * (1.0f - a.lr * a.weight_decay);
      const float denom = sqrtf(v) / sqrtf(a.bias2) + a.eps;
      p -= (a.lr / a.bias1) * (m / denom);
      sg.ptr[i] = p;
    }
  }
#undef FOR_ALL
}

}  // namespace

}  // namespace bfb

using namespace bfb;

#ifdef BFB_TRAIN_TIMING
extern "C" int bfb200_debug_train_timing(unsigned long long* out128) {   // debug build only: globaltimer at every barrier of the last step
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
  if (batch < 1 || seq_len < 2 || prompt_len < 1 || prompt_len >= seq_len) return 0;
  return (size_t)train_scratch(batch, seq_len, prompt_len, ntoken, d_ff, n_layers).total;
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
  BFB_REQUIRE(ta->batch >= 1 && ta->prompt_len >= 1 && ta->seq_len > ta->prompt_len, BFB200_ERR_INVALID_ARG,
              "bad batch shape (B=%d, T=%d, P=%d)", ta->batch, ta->seq_len, ta->prompt_len);
  BFB_REQUIRE(ta->seq_len <= ta->pe_len, BFB200_ERR_INVALID_ARG, "sequence longer than the positional table");
  BFB_REQUIRE(ta->seq_len <= TMAXT, BFB200_ERR_UNSUPPORTED, "fused train step takes sequences of at most %d tokens (got %d)",
              TMAXT, ta->seq_len);
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
  a.sc = train_scratch(ta->batch, ta->seq_len, ta->prompt_len, ta->ntoken, ta->d_ff, ta->n_layers);
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
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(TCLUSTER, 1, 1);
  cfg.blockDim = dim3(TTHREADS, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TCLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, train_step_kernel, a);
  BFB_REQUIRE(e == cudaSuccess, BFB200_ERR_CUDA, "launch of train_step_kernel failed: %s", cudaGetErrorString(e));
  BFB_LAUNCH_CHECK("train_step_kernel", (cudaStream_t)stream);
  return BFB200_OK;
}

// Layer-by-layer kernels of the MC-batched transformer forward (fp32 path).
//
// Activations are row-major (R, d) with row r = s * len + t: `s` is the sequence inside
// the chunk (item-major, draw-minor: s = item * n_draws + draw), `t` the position, so one
// sequence's rows are contiguous (attention) and one item's draws are adjacent (acquisition).
// Reference arithmetic: src/model/transformer.py (cited per kernel).
#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace bfb {

static int sm_count() {   // of the CURRENT device, asked per call (a process may drive several devices)
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

static int stride_grid(int64_t blocks_needed, int blocks_per_sm) {
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  if (blocks_needed < 1) blocks_needed = 1;
  return (int)(blocks_needed < cap ? blocks_needed : cap);
}

// ---------------------------------------------------------------------------
// prompts (P, B_total) int64 seq-first -> tokens (S, stride) int32, replicated over draws
// ---------------------------------------------------------------------------
__global__ void init_tokens_kernel(const int64_t* __restrict__ prompts, int64_t n_items_total,
                                   int64_t item0, int64_t n_seq, int n_draws, int prompt_len,
                                   int stride, int vocab, int32_t* __restrict__ tokens) {
  const int64_t total = n_seq * prompt_len;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / prompt_len;
    const int t = (int)(i - s * prompt_len);
    int64_t tok = prompts[(int64_t)t * n_items_total + item0 + s / n_draws];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);  // nn.Embedding would raise; stay in bounds
    tokens[s * stride + t] = (int32_t)tok;
  }
}

// ---------------------------------------------------------------------------
// x = dropout_E(embedding[tok]) * sqrt(d) + pe[t]      (transformer.py:355-362, :269-272)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_kernel(const int32_t* __restrict__ tokens, int tok_stride,
                                                    const float* __restrict__ emb,
                                                    const float* __restrict__ pe, int64_t n_seq, int len,
                                                    int d, float sqrt_d, MaskSource ms, int site,
                                                    float* __restrict__ x, uint16_t* __restrict__ x16,
                                                    uint32_t fmt) {
  const int d4 = d >> 2;
  const int64_t total = n_seq * len * d4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d4;
    const int f4 = (int)(i - r * d4);
    const int64_t s = r / len;
    const int t = (int)(r - s * len);
    const int tok = tokens[s * tok_stride + t];
    float4 e = __ldg(reinterpret_cast<const float4*>(emb + (int64_t)tok * d) + f4);
    if (site >= 0) {
      const uint32_t keep = keep4(ms, s, (uint32_t)t, (uint32_t)site, (uint32_t)f4);
      e.x = (keep & 1u) ? e.x * ms.scale : 0.f;
      e.y = (keep & 2u) ? e.y * ms.scale : 0.f;
      e.z = (keep & 4u) ? e.z * ms.scale : 0.f;
      e.w = (keep & 8u) ? e.w * ms.scale : 0.f;
    }
    const float4 p = __ldg(reinterpret_cast<const float4*>(pe + (int64_t)t * d) + f4);
    float4 o;
    o.x = e.x * sqrt_d + p.x;
    o.y = e.y * sqrt_d + p.y;
    o.z = e.z * sqrt_d + p.z;
    o.w = e.w * sqrt_d + p.w;
    reinterpret_cast<float4*>(x)[i] = o;
    if (x16 != nullptr) {   // 16-bit copy: the operand of the next tensor-core product
      uint2 w;
      w.x = fmt == umma::kFmtF16 ? umma::pack_f16x2(o.x, o.y) : umma::pack_bf16x2(o.x, o.y);
      w.y = fmt == umma::kFmtF16 ? umma::pack_f16x2(o.z, o.w) : umma::pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(x16)[i] = w;
    }
  }
}

// ---------------------------------------------------------------------------
// C[M,N] = A[M,K] * W[N,K]^T (+ bias) (ReLU)     nn.Linear (transformer.py:33-35,100,152-154,336)
// fp32 CUDA-core GEMM: 128x64 block tile, 8x4 register tile, K staged 16 at a time
// through shared memory (stored k-major so the inner product reads are conflict-free
// float4 loads).  A row m lives at A + (m * a_row_mul + a_row_off) * K, which lets the
// vocabulary head read only each sequence's last position.
// ---------------------------------------------------------------------------
constexpr int GB_M = 128, GB_N = 64, GB_K = 16, GT_M = 8, GT_N = 4;

__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ A,
                                                       const float* __restrict__ W,
                                                       const float* __restrict__ bias,
                                                       float* __restrict__ C, int64_t M, int N, int K,
                                                       int64_t a_row_mul, int64_t a_row_off, int relu) {
  __shared__ __align__(16) float As[GB_K][GB_M + 4];
  __shared__ __align__(16) float Ws[GB_K][GB_N + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * GB_M;
  const int n0 = blockIdx.y * GB_N;

  float acc[GT_M][GT_N];
#pragma unroll
  for (int i = 0; i < GT_M; ++i)
#pragma unroll
    for (int j = 0; j < GT_N; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += GB_K) {
    // A tile: 128 rows x 16 k = 512 float4, two per thread
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * 256;
      const int row = idx >> 2, kq = (idx & 3) << 2;
      const int64_t m = m0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && k0 + kq < K)
        v = __ldg(reinterpret_cast<const float4*>(A + (m * a_row_mul + a_row_off) * K + k0 + kq));
      As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
    }
    // W tile: 64 rows x 16 k = 256 float4, one per thread
    {
      const int row = tid >> 2, kq = (tid & 3) << 2;
      const int n = n0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N && k0 + kq < K) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * K + k0 + kq));
      Ws[kq + 0][row] = v.x; Ws[kq + 1][row] = v.y; Ws[kq + 2][row] = v.z; Ws[kq + 3][row] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GB_K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * GT_M]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * GT_M + 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * GT_N]);
      const float a[GT_M] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[GT_N] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < GT_M; ++i)
#pragma unroll
        for (int j = 0; j < GT_N; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < GT_M; ++i) {
    const int64_t m = m0 + ty * GT_M + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < GT_N; ++j) {
      const int n = n0 + tx * GT_N + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += __ldg(bias + n);
      if (relu) v = fmaxf(v, 0.f);
      C[m * N + n] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// causal multi-head attention (transformer.py:58-65, :123-124)
//   scores = Q K^T / sqrt(dh) + causal(-inf); softmax; @ V; heads concatenated
// one block per (sequence, head); K/V of that head staged in shared memory; each warp owns
// query rows t = warp, warp + 4, ...; keys are spread over lanes.
// ---------------------------------------------------------------------------
template <int MAXJ>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv,
                                                        float* __restrict__ out, int len, int d, int dh,
                                                        float sqrt_dh, MaskSource ms, int site) {
  extern __shared__ float sm[];
  const int h = blockIdx.y;
  const int64_t s = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldk = dh + 1;
  float* Ks = sm;                    // len * (dh + 1)
  float* Vs = Ks + len * ldk;        // len * dh
  float* Qs = Vs + len * dh;         // 4 * dh
  float* Ps = Qs + 4 * dh;           // 4 * len
  const float* base = qkv + s * len * 3 * (int64_t)d + h * dh;
  for (int i = threadIdx.x; i < len * dh; i += blockDim.x) {
    const int t = i / dh, c = i - t * dh;
    Ks[t * ldk + c] = __ldg(base + (int64_t)t * 3 * d + d + c);
    Vs[t * dh + c] = __ldg(base + (int64_t)t * 3 * d + 2 * d + c);
  }
  __syncthreads();
  float* q = Qs + warp * dh;
  float* p = Ps + warp * len;
  for (int t = warp; t < len; t += 4) {
    for (int c = lane; c < dh; c += 32) q[c] = __ldg(base + (int64_t)t * 3 * d + c);
    __syncwarp();
    float sc[MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      float a = -INFINITY;
      if (j <= t) {
        a = 0.f;
        for (int c = 0; c < dh; ++c) a = fmaf(q[c], Ks[j * ldk + c], a);
        a = a / sqrt_dh;
      }
      sc[jj] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float z = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      const float e = (j <= t) ? expf(sc[jj] - mx) : 0.f;
      sc[jj] = e;
      z += e;
    }
    z = warp_sum(z);
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j <= t) p[j] = sc[jj] / z;
    }
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float a = 0.f;
      for (int j = 0; j <= t; ++j) a = fmaf(p[j], Vs[j * dh + c], a);
      if (site >= 0) {
        const int f = h * dh + c;
        const uint32_t keep = keep4(ms, s, (uint32_t)t, (uint32_t)site, (uint32_t)(f >> 2));
        a = ((keep >> (f & 3)) & 1u) ? a * ms.scale : 0.f;
      }
      out[(s * len + t) * (int64_t)d + h * dh + c] = a;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// causal attention for short sequences and narrow heads (len <= 8, dh == 8: the notebook model on the fp32 path).
// thread = (sequence, head): K and V of the pair live in registers, queries are walked in order; consecutive
// threads read consecutive heads of one row = 256 contiguous bytes.  Same arithmetic as attention_kernel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attention_small_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                              int64_t n_seq, int len, int d, int n_heads, float sqrt_dh,
                                                              MaskSource ms, int site) {
  const int64_t pair = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n_seq * n_heads) return;
  const int64_t s = pair / n_heads;
  const int h = (int)(pair - s * n_heads);
  const float* base = qkv + s * len * 3 * (int64_t)d + h * 8;
  float k[8][8], v[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < len) {
      const float4* kr = reinterpret_cast<const float4*>(base + (int64_t)j * 3 * d + d);
      const float4* vr = reinterpret_cast<const float4*>(base + (int64_t)j * 3 * d + 2 * d);
      const float4 k0 = __ldg(kr), k1 = __ldg(kr + 1), v0 = __ldg(vr), v1 = __ldg(vr + 1);
      k[j][0] = k0.x; k[j][1] = k0.y; k[j][2] = k0.z; k[j][3] = k0.w; k[j][4] = k1.x; k[j][5] = k1.y; k[j][6] = k1.z; k[j][7] = k1.w;
      v[j][0] = v0.x; v[j][1] = v0.y; v[j][2] = v0.z; v[j][3] = v0.w; v[j][4] = v1.x; v[j][5] = v1.y; v[j][6] = v1.z; v[j][7] = v1.w;
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    if (t < len) {
      const float4* qr = reinterpret_cast<const float4*>(base + (int64_t)t * 3 * d);
      const float4 q0 = __ldg(qr), q1 = __ldg(qr + 1);
      const float q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      float sc[8];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j <= t; ++j) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) a = fmaf(q[c], k[j][c], a);
        a = a / sqrt_dh;
        sc[j] = a;
        mx = fmaxf(mx, a);
      }
      float z = 0.f;
#pragma unroll
      for (int j = 0; j <= t; ++j) {
        sc[j] = expf(sc[j] - mx);
        z += sc[j];
      }
      float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j <= t; ++j) {
        const float pj = sc[j] / z;
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = fmaf(pj, v[j][c], o[c]);
      }
      if (site >= 0) {
#pragma unroll
        for (int f4 = 0; f4 < 2; ++f4) {
          const uint32_t keep = keep4(ms, s, (uint32_t)t, (uint32_t)site, (uint32_t)(h * 2 + f4));
#pragma unroll
          for (int c = 0; c < 4; ++c) o[4 * f4 + c] = ((keep >> c) & 1u) ? o[4 * f4 + c] * ms.scale : 0.f;
        }
      }
      float4* dst = reinterpret_cast<float4*>(out + (s * len + t) * (int64_t)d + h * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// ---------------------------------------------------------------------------
// out = dropout_out( LayerNorm(x + dropout_branch(y)) )     (transformer.py:226-233, :370)
// one warp per row, 128-bit loads, statistics by warp shuffle; biased variance, eps inside
// the square root (nn.LayerNorm).
// ---------------------------------------------------------------------------
template <int NPL4>
__global__ void __launch_bounds__(256, 4) resid_ln_kernel(const float* __restrict__ x,
                                                       const float* __restrict__ y,
                                                       float* __restrict__ out, int64_t n_rows, int len,
                                                       int d, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps,
                                                       MaskSource ms, int branch_site, int out_site,
                                                       uint16_t* __restrict__ out16, uint32_t fmt,
                                                       const uint16_t* __restrict__ y16) {
  const int lane = threadIdx.x & 31;
  const int d4 = d >> 2;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows;
       r += warps_total) {
    const int64_t s = r / len;
    const uint32_t t = (uint32_t)(r - s * len);
    const float4* xr = reinterpret_cast<const float4*>(x + r * d);
    const float4* yr = reinterpret_cast<const float4*>(y + r * d);
    const uint2* yr16 = reinterpret_cast<const uint2*>(y16 + r * d);
    float4 v[NPL4];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NPL4; ++j) {
      const int f4 = lane + 32 * j;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f4 < d4) {
        const float4 a = __ldg(xr + f4);
        float4 b;
        if (y16 != nullptr) {   // branch delivered by the tensor-core GEMM as 16-bit
          const uint2 w = __ldg(yr16 + f4);
          if (fmt == umma::kFmtF16) {
            const float2 lo = umma::unpack_f16x2(w.x), hi = umma::unpack_f16x2(w.y);
            b = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            b = make_float4(umma::bf16_lo(w.x), umma::bf16_hi(w.x), umma::bf16_lo(w.y), umma::bf16_hi(w.y));
          }
        } else {
          b = __ldg(yr + f4);
        }
        if (branch_site >= 0) {
          const uint32_t keep = keep4(ms, s, t, (uint32_t)branch_site, (uint32_t)f4);
          b.x = (keep & 1u) ? b.x * ms.scale : 0.f;
          b.y = (keep & 2u) ? b.y * ms.scale : 0.f;
          b.z = (keep & 4u) ? b.z * ms.scale : 0.f;
          b.w = (keep & 8u) ? b.w * ms.scale : 0.f;
        }
        v[j] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    }
    const float mean = warp_sum(sum) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NPL4; ++j) {
      if (lane + 32 * j < d4) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
        sq += (a * a + b * b) + (c * c + e * e);
      }
    }
    const float rstd = 1.f / sqrtf(warp_sum(sq) / (float)d + eps);
#pragma unroll
    for (int j = 0; j < NPL4; ++j) {
      const int f4 = lane + 32 * j;
      if (f4 < d4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + f4);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + f4);
        float4 o;
        o.x = (v[j].x - mean) * rstd * g.x + bb.x;
        o.y = (v[j].y - mean) * rstd * g.y + bb.y;
        o.z = (v[j].z - mean) * rstd * g.z + bb.z;
        o.w = (v[j].w - mean) * rstd * g.w + bb.w;
        if (out_site >= 0) {
          const uint32_t keep = keep4(ms, s, t, (uint32_t)out_site, (uint32_t)f4);
          o.x = (keep & 1u) ? o.x * ms.scale : 0.f;
          o.y = (keep & 2u) ? o.y * ms.scale : 0.f;
          o.z = (keep & 4u) ? o.z * ms.scale : 0.f;
          o.w = (keep & 8u) ? o.w * ms.scale : 0.f;
        }
        if (out != nullptr) reinterpret_cast<float4*>(out + r * d)[f4] = o;
        if (out16 != nullptr) {
          uint2 w;
          w.x = fmt == umma::kFmtF16 ? umma::pack_f16x2(o.x, o.y) : umma::pack_bf16x2(o.x, o.y);
          w.y = fmt == umma::kFmtF16 ? umma::pack_f16x2(o.z, o.w) : umma::pack_bf16x2(o.z, o.w);
          reinterpret_cast<uint2*>(out16 + r * d)[f4] = w;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Tensor-core path variants of the two elementwise kernels (16-bit branch in, fp32 and / or 16-bit out).  A thread
// owns EIGHT contiguous features, which is exactly one Philox call of the mask stream (common.cuh: one call decides
// eight features) — the float4-per-thread kernels above evaluate every call twice.  Per thread and group: 2 x 128-bit
// loads of the fp32 residual, one 128-bit load of the 16-bit branch, one 128-bit store of the 16-bit output.
// ---------------------------------------------------------------------------
// Mask addressing of one token row, computed once per row with 32-bit arithmetic (a chunk never holds 2^31 rows)
struct RowKey {
  uint32_t item, draw, seq, pos;
};
__device__ __forceinline__ RowKey make_row_key(const MaskSource& m, uint32_t row, uint32_t len) {
  RowKey k;
  k.seq = row / len;
  k.pos = row - k.seq * len;
  const uint32_t q = k.seq / m.n_draws;
  k.item = (uint32_t)((uint64_t)m.item_base + q);
  k.draw = m.draw_base + (k.seq - q * m.n_draws);
  return k;
}

// v[0..8) (features 8 g .. 8 g + 7 of the row) through the dropout site: kept values are scaled by 1 / (1 - p).
// Philox mode decides straight on the 16-bit halves of the four output words (common.cuh: feature j <- half j & 1 of
// word j >> 1, kept iff half >= threshold); mask-injection mode reads the packed bits.
__device__ __forceinline__ void drop8_site(float (&v)[8], const MaskSource& m, const RowKey& k, uint32_t site, uint32_t g) {
  if (m.bits != nullptr) {
    const int64_t row = (((int64_t)m.step * m.n_sites + site) * m.n_seq + (m.seq_base + k.seq)) * m.max_len + k.pos;
    const uint32_t keep = (m.bits[row * m.words + (g >> 2)] >> ((g & 3u) * 8u)) & 0xFFu;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ((keep >> i) & 1u) ? v[i] * m.scale : 0.f;
    return;
  }
  const uint4 r = philox4x32_10(make_uint4(k.item, k.draw, (m.step << 16) | k.pos, (site << 16) | g), m.key);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  const uint32_t thr_hi = m.threshold << 16;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = (w[i] << 16) >= thr_hi ? v[2 * i] * m.scale : 0.f;          // low half  >= threshold
    v[2 * i + 1] = w[i] >= thr_hi ? v[2 * i + 1] * m.scale : 0.f;          // high half >= threshold
  }
}

__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8_f32(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void load8_16(const uint16_t* p, uint32_t fmt, float (&v)[8]) {
  const uint4 w = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (fmt == umma::kFmtF16) {
      const float2 f = umma::unpack_f16x2(ww[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    } else {
      v[2 * i] = umma::bf16_lo(ww[i]); v[2 * i + 1] = umma::bf16_hi(ww[i]);
    }
  }
}
__device__ __forceinline__ void store8_16(uint16_t* p, uint32_t fmt, const float (&v)[8]) {
  uint4 w;
  if (fmt == umma::kFmtF16) {
    w.x = umma::pack_f16x2(v[0], v[1]); w.y = umma::pack_f16x2(v[2], v[3]);
    w.z = umma::pack_f16x2(v[4], v[5]); w.w = umma::pack_f16x2(v[6], v[7]);
  } else {
    w.x = umma::pack_bf16x2(v[0], v[1]); w.y = umma::pack_bf16x2(v[2], v[3]);
    w.z = umma::pack_bf16x2(v[4], v[5]); w.w = umma::pack_bf16x2(v[6], v[7]);
  }
  *reinterpret_cast<uint4*>(p) = w;
}

// x = dropout_E(embedding[tok]) * sqrt(d) + pe[t], fp32 and 16-bit copies      (transformer.py:355-362, :269-272)
template <bool F16>
__global__ void __launch_bounds__(256) embed16_kernel(const int32_t* __restrict__ tokens, int tok_stride,
                                                      const float* __restrict__ emb, const float* __restrict__ pe,
                                                      int64_t n_seq, int len, int d, float sqrt_d, MaskSource ms, int site,
                                                      float* __restrict__ x, uint16_t* __restrict__ x16) {
  constexpr uint32_t fmt = F16 ? umma::kFmtF16 : umma::kFmtBF16;
  const int d8 = d >> 3;
  const int64_t total = n_seq * len * d8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(i / d8);
    const int g = (int)(i - (int64_t)r * d8);
    const RowKey rk = make_row_key(ms, r, (uint32_t)len);
    const int tok = tokens[(int64_t)rk.seq * tok_stride + rk.pos];
    float e[8], p[8];
    load8_f32(emb + (int64_t)tok * d + 8 * g, e);
    load8_f32(pe + (int64_t)rk.pos * d + 8 * g, p);
    if (site >= 0) drop8_site(e, ms, rk, (uint32_t)site, (uint32_t)g);
#pragma unroll
    for (int c = 0; c < 8; ++c) e[c] = e[c] * sqrt_d + p[c];
    if (x != nullptr) store8_f32(x + i * 8, e);
    store8_16(x16 + i * 8, fmt, e);
  }
}

// out = dropout_out( LayerNorm(x + dropout_branch(y16)) ), one warp per row      (transformer.py:226-233, :370)
template <int NJ>
__global__ void __launch_bounds__(256, 4) resid_ln16_kernel(const float* __restrict__ x, const uint16_t* __restrict__ y16,
                                                         float* __restrict__ out, uint16_t* __restrict__ out16,
                                                         int64_t n_rows, int len, int d, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, MaskSource ms,
                                                         int branch_site, int out_site, uint32_t fmt) {
  const int lane = threadIdx.x & 31;
  const int d8 = d >> 3;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps_total) {
    const RowKey rk = make_row_key(ms, (uint32_t)r, (uint32_t)len);
    float v[NJ][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int g = lane + 32 * j;
#pragma unroll
      for (int c = 0; c < 8; ++c) v[j][c] = 0.f;
      if (g < d8) {
        float b[8];
        load8_f32(x + r * d + 8 * g, v[j]);
        load8_16(y16 + r * d + 8 * g, fmt, b);
        if (branch_site >= 0) drop8_site(b, ms, rk, (uint32_t)branch_site, (uint32_t)g);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[j][c] += b[c];
        sum += ((v[j][0] + v[j][1]) + (v[j][2] + v[j][3])) + ((v[j][4] + v[j][5]) + (v[j][6] + v[j][7]));
      }
    }
    const float mean = warp_sum(sum) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (lane + 32 * j < d8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float a = v[j][c] - mean;
          sq = fmaf(a, a, sq);
        }
      }
    }
    const float rstd = 1.f / sqrtf(warp_sum(sq) / (float)d + eps);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int g = lane + 32 * j;
      if (g < d8) {
        float gm[8], bt[8], o[8];
        load8_f32(gamma + 8 * g, gm);
        load8_f32(beta + 8 * g, bt);
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = (v[j][c] - mean) * rstd * gm[c] + bt[c];
        if (out_site >= 0) drop8_site(o, ms, rk, (uint32_t)out_site, (uint32_t)g);
        if (out != nullptr) store8_f32(out + r * d + 8 * g, o);
        if (out16 != nullptr) store8_16(out16 + r * d + 8 * g, fmt, o);
      }
    }
  }
}

// out16 = dropout_out( LayerNorm(v16) ) with the row statistics already known: the GEMM that produced v (its epilogue
// added the residual) also accumulated each row's sum and sum of squares (launch_gemm_tc_ex), so this kernel is one
// 16-bit read and one 16-bit write per element (transformer.py:233, :370).  One warp per row.
template <int NJ, bool F16>
__global__ void __launch_bounds__(256) ln16_apply_kernel(const uint16_t* __restrict__ v16, const RowStats* __restrict__ stats,
                                                         uint16_t* __restrict__ out16, int64_t n_rows, int len, int d,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, MaskSource ms, int out_site) {
  constexpr uint32_t fmt = F16 ? umma::kFmtF16 : umma::kFmtBF16;   // compile-time: no per-element format branches
  const int lane = threadIdx.x & 31;
  const int d8 = d >> 3;
  const float inv_d = 1.0f / (float)d;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps_total) {
    const RowKey rk = make_row_key(ms, (uint32_t)r, (uint32_t)len);
    const float2 st = row_stats_load(stats + r);
    const float mean = st.x * inv_d;
    const float rstd = rsqrtf(fmaxf(st.y * inv_d - mean * mean, 0.f) + eps);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int g = lane + 32 * j;
      if (g < d8) {
        float v[8], gm[8], bt[8];
        load8_16(v16 + r * d + 8 * g, fmt, v);
        load8_f32(gamma + 8 * g, gm);
        load8_f32(beta + 8 * g, bt);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = (v[c] - mean) * rstd * gm[c] + bt[c];
        if (out_site >= 0) drop8_site(v, ms, rk, (uint32_t)out_site, (uint32_t)g);
        store8_16(out16 + r * d + 8 * g, fmt, v);
      }
    }
  }
}

int launch_ln16_apply(const void* v16, const RowStats* stats, void* out16, int64_t n_rows, int len, int d, const float* gamma,
                      const float* beta, float eps, const MaskSource& ms, int out_site, uint32_t fmt, cudaStream_t st) {
  BFB_REQUIRE((d & 7) == 0 && d <= 2048, BFB200_ERR_UNSUPPORTED, "ln16_apply needs d_model %% 8 == 0 and <= 2048 (got %d)", d);
  BFB_REQUIRE(n_rows < 0x7FFFFFFFLL, BFB200_ERR_UNSUPPORTED, "more than 2^31 token rows in one chunk");
  const int grid = stride_grid((n_rows + 7) / 8, 8);
#define LAUNCH_LNA(NJ)                                                                                                         \
  do {                                                                                                                         \
    if (fmt == umma::kFmtF16)                                                                                                  \
      ln16_apply_kernel<NJ, true><<<grid, 256, 0, st>>>((const uint16_t*)v16, stats, (uint16_t*)out16, n_rows, len, d, gamma, beta, eps, ms, out_site); \
    else                                                                                                                       \
      ln16_apply_kernel<NJ, false><<<grid, 256, 0, st>>>((const uint16_t*)v16, stats, (uint16_t*)out16, n_rows, len, d, gamma, beta, eps, ms, out_site); \
  } while (0)
  if (d <= 256) LAUNCH_LNA(1);
  else if (d <= 512) LAUNCH_LNA(2);
  else if (d <= 1024) LAUNCH_LNA(4);
  else LAUNCH_LNA(8);
#undef LAUNCH_LNA
  BFB_LAUNCH_CHECK("ln16_apply_kernel", st);
  return BFB200_OK;
}

// ---------------------------------------------------------------------------
// last-position logits -> log-softmax -> probabilities -> per-draw score + next token
// (transformer.py:375; active_learning.py:62-68, :103-112, :146-152)
// one warp per sequence.
// ---------------------------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(256) score_step_kernel(
    const float* __restrict__ logits, int64_t n_seq, int vocab, int scheme, int mode, float temperature,
    MaskSource ms, const int32_t* __restrict__ forced, int32_t* __restrict__ tokens, int tok_stride,
    int write_pos, float* __restrict__ seq_scores, float* __restrict__ logp_out, int pitch) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n_seq;
       s += warps_total) {
    float l[NPL];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int c = lane + 32 * j;
      l[j] = c < vocab ? __ldg(logits + s * pitch + c) : -INFINITY;
      mx = fmaxf(mx, l[j]);
    }
    mx = warp_max(mx);
    float z = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) z += (lane + 32 * j) < vocab ? expf(l[j] - mx) : 0.f;
    const float lse = logf(warp_sum(z));
    // log-probs + greedy arg-max (first index wins ties)
    float best = -INFINITY;
    int best_i = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int c = lane + 32 * j;
      if (c < vocab) {
        l[j] = (l[j] - mx) - lse;
        if (logp_out != nullptr) logp_out[s * vocab + c] = l[j];
        if (l[j] > best) { best = l[j]; best_i = c; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    // probs = softmax(log-probs [/ temperature])
    float p[NPL];
    float pm = -INFINITY;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      p[j] = (scheme == BFB200_SCHEME_SAMPLING) ? l[j] / temperature : l[j];
      if ((lane + 32 * j) < vocab) pm = fmaxf(pm, p[j]);
    }
    pm = warp_max(pm);
    float pz = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      p[j] = (lane + 32 * j) < vocab ? expf(p[j] - pm) : 0.f;
      pz += p[j];
    }
    pz = warp_sum(pz);
#pragma unroll
    for (int j = 0; j < NPL; ++j) p[j] = p[j] / pz;

    int token = best_i;
    if (scheme == BFB200_SCHEME_SAMPLING) {
      // inverse-CDF draw in class order from one Philox word
      const uint64_t item = (uint64_t)(ms.item_base + s / ms.n_draws);
      const uint32_t draw = ms.draw_base + (uint32_t)(s % ms.n_draws);
      const uint4 r = philox4x32_10(
          make_uint4((uint32_t)item, draw, ms.step << 16, BFB200_SITE_SAMPLING_TOKEN << 16), ms.key);
      const float u = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
      float running = 0.f;
      token = vocab - 1;
      bool found = false;
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        float c = p[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float up = __shfl_up_sync(0xffffffffu, c, o);
          if (lane >= o) c += up;
        }
        c += running;
        const unsigned hit = __ballot_sync(0xffffffffu, c > u && (lane + 32 * j) < vocab);
        if (!found && hit != 0u) { token = 32 * j + (__ffs(hit) - 1); found = true; }
        running = __shfl_sync(0xffffffffu, c, 31);
      }
    }
    float s0, s1, s2;
    warp_scores<NPL>(p, vocab, lane, s0, s1, s2);
    if (lane == 0) {
      seq_scores[s] = mode == BFB200_MODE_MAX ? s0 : (mode == BFB200_MODE_MARGIN ? s1 : s2);
      if (forced != nullptr) token = forced[s];
      if (tokens != nullptr && write_pos < tok_stride) tokens[s * tok_stride + write_pos] = token;
    }
  }
}

// logits (R, V) rows r = b * len + t -> log-probs written seq-first (len, B, V)
template <int NPL>
__global__ void __launch_bounds__(256) logsoftmax_seqfirst_kernel(const float* __restrict__ logits,
                                                                  int64_t n_seq, int len, int vocab,
                                                                  float* __restrict__ out, int pitch) {
  const int lane = threadIdx.x & 31;
  const int64_t n_rows = n_seq * len;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows;
       r += warps_total) {
    const int64_t b = r / len;
    const int64_t t = r - b * len;
    float l[NPL];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int c = lane + 32 * j;
      l[j] = c < vocab ? __ldg(logits + r * pitch + c) : -INFINITY;
      mx = fmaxf(mx, l[j]);
    }
    mx = warp_max(mx);
    float z = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) z += (lane + 32 * j) < vocab ? expf(l[j] - mx) : 0.f;
    const float lse = logf(warp_sum(z));
    float* dst = out + (t * n_seq + b) * vocab;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int c = lane + 32 * j;
      if (c < vocab) dst[c] = (l[j] - mx) - lse;
    }
  }
}

// step_scores[step, item] = (sum over draws in order) / T       (active_learning.py:154-155)
__global__ void reduce_draws_kernel(const float* __restrict__ seq_scores, int64_t n_items_chunk,
                                    int n_draws, float* __restrict__ step_scores_row) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items_chunk) return;
  float acc = 0.f;
  for (int t = 0; t < n_draws; ++t) acc += seq_scores[i * n_draws + t];
  step_scores_row[i] = acc / (float)n_draws;
}

// pool_scores[item] = mean over steps                            (active_learning.py:209)
__global__ void step_mean_kernel(const float* __restrict__ step_scores, int64_t n_items, int n_steps,
                                 float* __restrict__ pool_scores) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  float acc = 0.f;
  for (int t = 0; t < n_steps; ++t) acc += step_scores[(int64_t)t * n_items + i];
  pool_scores[i] = acc / (float)n_steps;
}

// fp32 -> 16-bit copy (weight image of the layered tensor-core path)
__global__ void to16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n, uint32_t fmt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t w = fmt == umma::kFmtF16 ? umma::pack_f16x2(src[i], 0.f) : umma::pack_bf16x2(src[i], 0.f);
    dst[i] = (uint16_t)(w & 0xFFFFu);
  }
}

// LayerNorm folded into the product that consumes it:  W LN(v) + b = rstd (v - mean) (W g)^T + (W beta + b).  The
// mean term costs nothing when the ROWS of W g are centred: with W'[n, :] = (W g)[n, :] - avg_k (W g)[n, k],
// v W'[n, :]^T = v (W g)[n, :]^T - (sum_k v_k) avg = (v - mean) (W g)[n, :]^T exactly, so the product's epilogue is
// rstd * acc + col_c — one FMA per element, like a bias.  Rounding W' to 16 bits leaves a row sum of ~5e-3 rms(W')
// instead of 0 (returned in col_s for inspection): against the product's own scale sqrt(K) rms(W') that is 2e-4 of
// mean / std of the row — below the operands' rounding noise, and the epilogue ignores it.
// One warp per output row n of layer l.
__global__ void __launch_bounds__(256) fold_ln_pack_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int64_t rows_total, int N, int K, uint32_t fmt,
                                                           uint16_t* __restrict__ w16g, float* __restrict__ col_s,
                                                           float* __restrict__ col_c) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows_total) return;
  const int64_t l = r / N;
  const float* w = W + r * K;
  const float* g = gamma + l * K;
  const float* be = beta + l * K;
  float tot = 0.f;
  for (int k = lane; k < K; k += 32) tot += w[k] * g[k];
  const float avg = warp_sum(tot) / (float)K;
  float s = 0.f, c = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wg = w[k] * g[k] - avg;
    uint16_t h;
    float back;
    if (fmt == umma::kFmtF16) {
      const __half hh = __float2half_rn(fminf(fmaxf(wg, -65504.f), 65504.f));
      h = *reinterpret_cast<const uint16_t*>(&hh);
      back = __half2float(hh);
    } else {
      const __nv_bfloat16 bb = __float2bfloat16_rn(wg);
      h = *reinterpret_cast<const uint16_t*>(&bb);
      back = __bfloat162float(bb);
    }
    w16g[r * K + k] = h;
    s += back;
    c = fmaf(w[k], be[k], c);
  }
  s = warp_sum(s);
  c = warp_sum(c);
  if (lane == 0) {
    col_s[r] = s;
    col_c[r] = c + b[r];
  }
}

int launch_fold_ln_pack(const float* W, const float* b, const float* gamma, const float* beta, int L, int N, int K,
                        uint32_t fmt, void* w16g, float* col_s, float* col_c, cudaStream_t st) {
  const int64_t rows = (int64_t)L * N;
  if (rows == 0) return BFB200_OK;
  fold_ln_pack_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(W, b, gamma, beta, rows, N, K, fmt, (uint16_t*)w16g, col_s,
                                                                 col_c);
  BFB_LAUNCH_CHECK("fold_ln_pack_kernel", st);
  return BFB200_OK;
}

int launch_to16(const float* src, void* dst, int64_t n, uint32_t fmt, cudaStream_t st) {
  if (n == 0) return BFB200_OK;
  to16_kernel<<<stride_grid((n + 255) / 256, 8), 256, 0, st>>>(src, (uint16_t*)dst, n, fmt);
  BFB_LAUNCH_CHECK("to16_kernel", st);
  return BFB200_OK;
}

// ---------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------
int launch_init_tokens(const int64_t* prompts, int64_t n_items_total, int64_t item0, int64_t n_seq,
                       int n_draws, int prompt_len, int stride, int vocab, int32_t* tokens,
                       cudaStream_t st) {
  const int grid = stride_grid((n_seq * prompt_len + 255) / 256, 8);
  init_tokens_kernel<<<grid, 256, 0, st>>>(prompts, n_items_total, item0, n_seq, n_draws, prompt_len,
                                           stride, vocab, tokens);
  BFB_LAUNCH_CHECK("init_tokens_kernel", st);
  return BFB200_OK;
}

int launch_embed(const int32_t* tokens, int tok_stride, const float* emb, const float* pe, int64_t n_seq,
                 int len, int d, const MaskSource& ms, int site, float* x, cudaStream_t st, void* x16, uint32_t fmt) {
  if (x16 != nullptr && (d & 7) == 0) {
    BFB_REQUIRE(n_seq * len < 0x7FFFFFFFLL, BFB200_ERR_UNSUPPORTED, "more than 2^31 token rows in one chunk");
    const int64_t total8 = n_seq * len * (d >> 3);
    if (fmt == umma::kFmtF16)
      embed16_kernel<true><<<stride_grid((total8 + 255) / 256, 8), 256, 0, st>>>(tokens, tok_stride, emb, pe, n_seq, len, d,
                                                                                sqrtf((float)d), ms, site, x, (uint16_t*)x16);
    else
      embed16_kernel<false><<<stride_grid((total8 + 255) / 256, 8), 256, 0, st>>>(tokens, tok_stride, emb, pe, n_seq, len, d,
                                                                                 sqrtf((float)d), ms, site, x, (uint16_t*)x16);
    BFB_LAUNCH_CHECK("embed16_kernel", st);
    return BFB200_OK;
  }
  const int64_t total = n_seq * len * (d >> 2);
  const int grid = stride_grid((total + 255) / 256, 8);
  embed_kernel<<<grid, 256, 0, st>>>(tokens, tok_stride, emb, pe, n_seq, len, d, sqrtf((float)d), ms, site, x,
                                     (uint16_t*)x16, fmt);
  BFB_LAUNCH_CHECK("embed_kernel", st);
  return BFB200_OK;
}

int launch_sgemm(const float* A, const float* W, const float* bias, float* C, int64_t M, int N, int K,
                 int64_t a_row_mul, int64_t a_row_off, int relu, cudaStream_t st) {
  if (M == 0) return BFB200_OK;
  BFB_REQUIRE((K & 3) == 0, BFB200_ERR_UNSUPPORTED, "GEMM K=%d must be a multiple of 4", K);
  dim3 grid((unsigned)((M + GB_M - 1) / GB_M), (unsigned)((N + GB_N - 1) / GB_N));
  sgemm_tn_kernel<<<grid, 256, 0, st>>>(A, W, bias, C, M, N, K, a_row_mul, a_row_off, relu);
  BFB_LAUNCH_CHECK("sgemm_tn_kernel", st);
  return BFB200_OK;
}

int launch_attention(const float* qkv, float* out, int64_t n_seq, int len, int d, int n_heads,
                     const MaskSource& ms, int site, cudaStream_t st) {
  const int dh = d / n_heads;
  if (dh == 8 && len <= 8) {
    const int64_t pairs = n_seq * n_heads;
    attention_small_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(qkv, out, n_seq, len, d, n_heads,
                                                                            sqrtf((float)dh), ms, site);
    BFB_LAUNCH_CHECK("attention_small_kernel", st);
    return BFB200_OK;
  }
  const size_t smem = ((size_t)len * (2 * dh + 1) + 4 * (size_t)dh + 4 * (size_t)len) * sizeof(float);
  BFB_REQUIRE(len <= 1024, BFB200_ERR_UNSUPPORTED, "sequence length %d > 1024", len);
  BFB_REQUIRE(smem <= 227 * 1024, BFB200_ERR_UNSUPPORTED,
              "attention tile (len=%d, dh=%d) does not fit shared memory", len, dh);
  BFB_REQUIRE(n_heads <= 65535, BFB200_ERR_UNSUPPORTED, "n_heads %d > 65535", n_heads);
  dim3 grid((unsigned)n_seq, (unsigned)n_heads);
#define LAUNCH_ATT(MAXJ)                                                                              \
  do {                                                                                                \
    if (smem > 48 * 1024)                                                                             \
      cudaFuncSetAttribute(attention_kernel<MAXJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    attention_kernel<MAXJ><<<grid, 128, smem, st>>>(qkv, out, len, d, dh, sqrtf((float)dh), ms, site); \
  } while (0)
  if (len <= 32) LAUNCH_ATT(1);
  else if (len <= 64) LAUNCH_ATT(2);
  else if (len <= 128) LAUNCH_ATT(4);
  else if (len <= 256) LAUNCH_ATT(8);
  else LAUNCH_ATT(32);
#undef LAUNCH_ATT
  BFB_LAUNCH_CHECK("attention_kernel", st);
  return BFB200_OK;
}

int launch_resid_ln(const float* x, const float* y, float* out, int64_t n_rows, int len, int d,
                    const float* gamma, const float* beta, float eps, const MaskSource& ms,
                    int branch_site, int out_site, cudaStream_t st, void* out16, uint32_t fmt, const void* y16) {
  BFB_REQUIRE(d <= 4096, BFB200_ERR_UNSUPPORTED, "d_model %d > 4096", d);
  const int grid = stride_grid((n_rows + 7) / 8, 8);
  if (y16 != nullptr && (d & 7) == 0 && d <= 2048) {   // tensor-core path: eight features per thread
    BFB_REQUIRE(n_rows < 0x7FFFFFFFLL, BFB200_ERR_UNSUPPORTED, "more than 2^31 token rows in one chunk");
#define LAUNCH_LN16(NJ) resid_ln16_kernel<NJ><<<grid, 256, 0, st>>>(x, (const uint16_t*)y16, out, (uint16_t*)out16, n_rows, len, d, gamma, beta, eps, ms, branch_site, out_site, fmt)
    if (d <= 256) LAUNCH_LN16(1);
    else if (d <= 512) LAUNCH_LN16(2);
    else if (d <= 1024) LAUNCH_LN16(4);
    else LAUNCH_LN16(8);
#undef LAUNCH_LN16
    BFB_LAUNCH_CHECK(out_site >= 0 || branch_site >= 0 ? "resid_ln16_kernel[dropout]" : "resid_ln16_kernel", st);
    return BFB200_OK;
  }
#define LAUNCH_LN(NPL4) resid_ln_kernel<NPL4><<<grid, 256, 0, st>>>(x, y, out, n_rows, len, d, gamma, beta, eps, ms, branch_site, out_site, (uint16_t*)out16, fmt, (const uint16_t*)y16)
  const int d4 = d >> 2;
  if (d4 <= 32) LAUNCH_LN(1);
  else if (d4 <= 64) LAUNCH_LN(2);
  else if (d4 <= 128) LAUNCH_LN(4);
  else if (d4 <= 256) LAUNCH_LN(8);
  else LAUNCH_LN(32);
#undef LAUNCH_LN
  BFB_LAUNCH_CHECK("resid_ln_kernel", st);
  return BFB200_OK;
}

int launch_score_step(const float* logits, int64_t n_seq, int vocab, int scheme, int mode,
                      float temperature, const MaskSource& ms, const int32_t* forced, int32_t* tokens,
                      int tok_stride, int write_pos, float* seq_scores, float* logp_out, cudaStream_t st, int pitch) {
  if (pitch <= 0) pitch = vocab;
  BFB_REQUIRE(vocab <= 1024, BFB200_ERR_UNSUPPORTED, "ntoken %d > 1024", vocab);
  const int grid = stride_grid((n_seq + 7) / 8, 8);
#define LAUNCH_SC(NPL) score_step_kernel<NPL><<<grid, 256, 0, st>>>(logits, n_seq, vocab, scheme, mode, temperature, ms, forced, tokens, tok_stride, write_pos, seq_scores, logp_out, pitch)
  if (vocab <= 32) LAUNCH_SC(1);
  else if (vocab <= 64) LAUNCH_SC(2);
  else if (vocab <= 128) LAUNCH_SC(4);
  else if (vocab <= 256) LAUNCH_SC(8);
  else LAUNCH_SC(32);
#undef LAUNCH_SC
  BFB_LAUNCH_CHECK("score_step_kernel", st);
  return BFB200_OK;
}

int launch_logsoftmax_seqfirst(const float* logits, int64_t n_seq, int len, int vocab, float* out,
                               cudaStream_t st, int pitch) {
  if (pitch <= 0) pitch = vocab;
  BFB_REQUIRE(vocab <= 1024, BFB200_ERR_UNSUPPORTED, "ntoken %d > 1024", vocab);
  const int grid = stride_grid((n_seq * len + 7) / 8, 8);
#define LAUNCH_LS(NPL) logsoftmax_seqfirst_kernel<NPL><<<grid, 256, 0, st>>>(logits, n_seq, len, vocab, out, pitch)
  if (vocab <= 32) LAUNCH_LS(1);
  else if (vocab <= 64) LAUNCH_LS(2);
  else if (vocab <= 128) LAUNCH_LS(4);
  else if (vocab <= 256) LAUNCH_LS(8);
  else LAUNCH_LS(32);
#undef LAUNCH_LS
  BFB_LAUNCH_CHECK("logsoftmax_seqfirst_kernel", st);
  return BFB200_OK;
}

int launch_reduce_draws(const float* seq_scores, int64_t n_items_chunk, int n_draws, float* dst,
                        cudaStream_t st) {
  reduce_draws_kernel<<<(unsigned)((n_items_chunk + 255) / 256), 256, 0, st>>>(seq_scores, n_items_chunk,
                                                                                n_draws, dst);
  BFB_LAUNCH_CHECK("reduce_draws_kernel", st);
  return BFB200_OK;
}

int launch_step_mean(const float* step_scores, int64_t n_items, int n_steps, float* pool_scores,
                     cudaStream_t st) {
  step_mean_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, st>>>(step_scores, n_items, n_steps,
                                                                       pool_scores);
  BFB_LAUNCH_CHECK("step_mean_kernel", st);
  return BFB200_OK;
}

// ---------------------------------------------------------------------------
// Head.bayes (transformer.py:50-53, a latent site: Head(..., bayes=True) built by hand): every head draws THREE
// independent (l, B, d) masks, one per projection, on the block input:
//   Q_h = W_q,h (x * m_hq / (1 - p)),  K_h = W_k,h (x * m_hk / (1 - p)),  V_h = W_v,h (x * m_hv / (1 - p))
// so the packed QKV product does not apply (3 H differently masked copies of x).  One block per token row: the row
// and the 3 H keep masks (d bits each) are staged in shared memory, then thread = output element.  Correctness path
// of a site nothing constructs by default; not tuned.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qkv_head_in_kernel(const float* __restrict__ x, const float* __restrict__ w_qkv,
                                                          float* __restrict__ qkv, int64_t n_rows, int len, int d,
                                                          int n_heads, MaskSource ms, int site_base) {
  extern __shared__ float hin_smem[];
  float* xs = hin_smem;                                         // d
  uint32_t* keep = reinterpret_cast<uint32_t*>(hin_smem + d);   // 3 H x ceil(d / 32) words
  const int words = (d + 31) >> 5, dh = d / n_heads, n_sites = 3 * n_heads;
  for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const int64_t s = r / len;
    const uint32_t t = (uint32_t)(r - s * len);
    for (int f = threadIdx.x; f < d; f += blockDim.x) xs[f] = x[r * d + f];
    for (int i = threadIdx.x; i < n_sites * words; i += blockDim.x) {
      const int site = i / words, w = i - site * words;
      uint32_t bits = 0u;
      for (int q = 0; q < 8 && (w * 8 + q) * 4 < d; ++q)
        bits |= keep4(ms, s, t, (uint32_t)(site_base + site), (uint32_t)(w * 8 + q)) << (4 * q);
      keep[i] = bits;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 3 * d; o += blockDim.x) {
      const int c = o / d, within = o - c * d, h = within / dh;   // output column o = c * d + h * dh + j
      const uint32_t* kb = keep + (3 * h + c) * words;
      const float* wrow = w_qkv + (int64_t)o * d;
      float acc = 0.f;
      for (int f = 0; f < d; ++f) {
        const float xv = (kb[f >> 5] >> (f & 31)) & 1u ? xs[f] * ms.scale : 0.f;
        acc = fmaf(xv, __ldg(wrow + f), acc);
      }
      qkv[r * 3 * d + o] = acc;
    }
    __syncthreads();
  }
}

int launch_qkv_head_in(const float* x, const float* w_qkv, float* qkv, int64_t n_seq, int len, int d, int n_heads,
                       const MaskSource& ms, int site_base, cudaStream_t st) {
  const int64_t rows = n_seq * len;
  if (rows == 0) return BFB200_OK;
  const size_t smem = (size_t)d * 4 + (size_t)3 * n_heads * ((d + 31) / 32) * 4;
  BFB_REQUIRE(smem <= 48 * 1024, BFB200_ERR_UNSUPPORTED, "Head.bayes site: d_model %d x %d heads exceeds the staging buffer", d, n_heads);
  qkv_head_in_kernel<<<stride_grid(rows, 8), 256, smem, st>>>(x, w_qkv, qkv, rows, len, d, n_heads, ms, site_base);
  BFB_LAUNCH_CHECK("qkv_head_in_kernel", st);
  return BFB200_OK;
}

}  // namespace bfb

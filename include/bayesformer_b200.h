/*
 * bayesformer_b200.h — C ABI of libbayesformer_b200.so
 *
 * B200 (sm_100a) implementation of ONE hot path of Suzie14/Bayesformer:
 * Monte-Carlo-dropout scoring of an unlabeled pool + uncertainty acquisition +
 * top-k.  The reference has no FFI of its own (pure Python/PyTorch); each entry
 * point below names the reference function (file:line under the reference tree)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every pointer marked [dev] must be CUDA device memory of the current device;
 *    the library has NO CPU path: host pointers are rejected with
 *    BFB200_ERR_NOT_DEVICE_PTR.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    All work is enqueued asynchronously on it; nothing synchronises.
 *  - functions return a BFB200_* status; bfb200_last_error() gives the message
 *    of the last failure on the calling thread.
 *  - no allocation happens inside: callers pass workspaces sized by the
 *    *_workspace_bytes queries (torch owns all memory).
 */
#ifndef BAYESFORMER_B200_H
#define BAYESFORMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFB200_ABI_VERSION 4

enum {
  BFB200_OK = 0,
  BFB200_ERR_INVALID_ARG = 1,
  BFB200_ERR_NOT_DEVICE_PTR = 2,
  BFB200_ERR_WORKSPACE = 3,
  BFB200_ERR_CUDA = 4,
  BFB200_ERR_UNSUPPORTED = 5
};

/* uncertainty score, src/libs/active_learning.py:25-31 */
enum { BFB200_MODE_MAX = 0, BFB200_MODE_MARGIN = 1, BFB200_MODE_ENTROPY = 2 };

/* stochastic-prediction scheme, src/libs/active_learning.py:38,73,118 */
enum { BFB200_SCHEME_GREEDY = 0, BFB200_SCHEME_SAMPLING = 1, BFB200_SCHEME_DROPOUT = 2 };

/* arithmetic of the contractions */
enum {
  BFB200_COMPUTE_FP32 = 0, /* fp32 CUDA-core FMA everywhere (1e-3 parity class) */
  BFB200_COMPUTE_F16 = 1   /* 16-bit operands on tcgen05 tensor cores (kind::f16), fp32 accumulate in TMEM, fp32
                              softmax / LayerNorm statistics: the north star's 1e-2 parity
                              class.  Operands are IEEE fp16 (saturating): bf16's 8-bit mantissa moves the
                              trained model's mid-range probabilities by up to 5 %, fp16 keeps them within
                              1 % (measured).  Served by the
                              fused MC-forward kernel: d_model 64, 8 heads, F <= 64, V <= 128, P + N - 1 <= 128
                              (the reference notebooks' model; operands split into hi + lo parts, residual
                              stream as an fp16 pair).  Other models with heads of width 64 and
                              sequences of at most 256 tokens (the scaled configuration) run the layered
                              tensor-core path: TMA-fed tcgen05 GEMMs + tcgen05 attention, fp32 residual stream */,
  BFB200_COMPUTE_BF16 = 2  /* layered tensor-core path with bf16 operands (heads of width 64, len <= 256) */
};

/* always-on dropout sites of src/model/transformer.py (bit set = site live) */
enum {
  BFB200_SITE_EMBED = 1u << 0,      /* :355-358, mask (l,B,d) before the sqrt(d) scale   */
  BFB200_SITE_LAYER_OUT = 1u << 1,  /* :368-370, mask (l,B,d) after every block          */
  BFB200_SITE_ATTN_RESID = 1u << 2, /* :226, latent: TransformerBlock.bayes              */
  BFB200_SITE_FFN_IN = 1u << 3,     /* :167-168, latent: FeedForward.bayes_dropout       */
  BFB200_SITE_FFN_RESID = 1u << 4,  /* :228, latent: TransformerBlock.bayes              */
  BFB200_SITE_HEAD_OUT = 1u << 5,   /* :116-122, latent: MultiHeadAttention.bayes        */
  BFB200_SITE_HEAD_IN = 1u << 6     /* :50-53, latent: Head.bayes — three independent (l,B,d) masks per head on the
                                       inputs of its Q, K and V projections (fp32 path only)    */
};

/* site ids used in the Philox counter and in the packed-mask layout:
 *   0                : embedding
 *   1 + 5*i + 0      : layer i output        (LAYER_OUT)
 *   1 + 5*i + 1      : layer i attn residual (ATTN_RESID)
 *   1 + 5*i + 2      : layer i FFN input     (FFN_IN)
 *   1 + 5*i + 3      : layer i FFN residual  (FFN_RESID)
 *   1 + 5*i + 4      : layer i head outputs  (HEAD_OUT, all heads concatenated)
 *   1 + 5*L + 3*(i*H + h) + c : layer i, head h, input of projection c (0 Q, 1 K, 2 V)   (HEAD_IN; these sites
 *                      exist — in the counter scheme and in the packed-mask layout — only when the flag is set:
 *                      n_sites = 1 + 5*L + 3*H*L then)
 */
#define BFB200_SITES_PER_LAYER 5
#define BFB200_SITE_SAMPLING_TOKEN 0xFFFFu /* multinomial draw of the sampling scheme */
#define BFB200_SITE_MLP_HIDDEN 0x100u      /* MLP hidden-layer dropout                 */
#define BFB200_SITE_VMLP_EPS 0x101u        /* VariationalMLP weight noise              */

/*
 * Weights of src/model/transformer.py:MyTransformer (:290-375), fp32, row-major
 * exactly as nn.Linear / nn.Embedding / nn.LayerNorm store them; the per-head
 * W_q/W_k/W_v (:33-35) are concatenated: rows [0,d) = Q of heads 0..H-1,
 * [d,2d) = K, [2d,3d) = V.
 */
typedef struct bfb200_transformer {
  int32_t ntoken;   /* V */
  int32_t d_model;  /* d */
  int32_t n_heads;  /* H, d % H == 0 */
  int32_t d_ff;     /* F */
  int32_t n_layers; /* L */
  int32_t pe_len;   /* rows of `pe` (5000 in the reference, :238) */
  float p_drop;     /* bayes_dropout; ignored when site_flags == 0 */
  uint32_t site_flags;
  const float* embedding; /* [dev] (V, d) */
  const float* pe;        /* [dev] (pe_len, d) */
  const float* w_qkv;     /* [dev] (L, 3d, d) */
  const float* w_out;     /* [dev] (L, d, d) */
  const float* ln1_w;     /* [dev] (L, d) */
  const float* ln1_b;     /* [dev] (L, d) */
  const float* ff1_w;     /* [dev] (L, F, d) */
  const float* ff1_b;     /* [dev] (L, F) */
  const float* ff2_w;     /* [dev] (L, d, F) */
  const float* ff2_b;     /* [dev] (L, d) */
  const float* ln2_w;     /* [dev] (L, d) */
  const float* ln2_b;     /* [dev] (L, d) */
  const float* head_w;    /* [dev] (V, d) */
  const float* head_b;    /* [dev] (V) */
  float ln_eps;           /* 1e-5 (nn.LayerNorm default) */
  int32_t compute;        /* BFB200_COMPUTE_* */
  const void* fused_image; /* [dev] fp16 weight image written by bfb200_fused_pack, or NULL; needed by
                              compute = BFB200_COMPUTE_F16 (see bfb200_fused_image_bytes) */
} bfb200_transformer;

/*
 * Source of dropout masks.
 *  - bits == NULL: counter-based Philox4x32-10, key = (seed_lo, seed_hi), counter =
 *      (global item index, draw, (step << 16) | position, (site << 16) | (feature >> 3));
 *      one call decides 8 features: feature f uses the 16-bit half (f & 1) of output word (f >> 1) & 3
 *      and is KEPT iff half >= min(65535, floor(p * 65536)) (p resolved to 1.5e-5).  Results do not
 *      depend on batching, chunking or the GPU a pool item lands on.
 *  - bits != NULL (mask-injection mode used by the parity tests): bit-packed keep
 *      masks, little-endian bit order inside uint32 words, laid out
 *      [step][site][sequence][position < max_len][ceil(d/32) words]
 *      with sequence = item * n_draws + draw, `n_sites` = 1 + 5 * L.
 */
typedef struct bfb200_masks {
  uint64_t seed;
  const uint32_t* bits; /* [dev] or NULL */
  int32_t max_len;      /* position stride of `bits` */
} bfb200_masks;

/*
 * One pool-scoring call: the loops of greedy_uncertainty / sampling_uncertainty /
 * dropout_uncertainty (src/libs/active_learning.py:38-155) with the n_draws MC
 * draws folded into the batch, followed by the step mean of select_samples (:209).
 */
typedef struct bfb200_mc_args {
  const int64_t* prompts; /* [dev] (P, B) seq-first, as preprocessing.get_batch returns */
  int64_t n_items;        /* B */
  int32_t prompt_len;     /* P */
  int32_t new_tokens;     /* N = answers_length + 1 */
  int32_t n_draws;        /* T_mc (1 for greedy) */
  int32_t scheme;         /* BFB200_SCHEME_* */
  int32_t mode;           /* BFB200_MODE_* */
  float temperature;      /* sampling scheme only (:104), 0.8 in the reference */
  int64_t index_base;     /* global pool index of item 0 (RNG counter; sharding) */
  bfb200_masks masks;
  const int32_t* forced_tokens; /* [dev] (N, B * n_draws) or NULL: teacher-forced appended tokens */
  float* step_scores;           /* [dev] (N, B): mean over draws of the per-draw score (:155) */
  float* pool_scores;           /* [dev] (B) or NULL: mean over steps (:209) */
  float* logp_out;              /* [dev] (N, B * n_draws, V) or NULL: last-position log-probs */
  int32_t* tokens_out;          /* [dev] (B * n_draws, P + N) or NULL: full trajectories */
} bfb200_mc_args;

int bfb200_abi_version(void);
const char* bfb200_last_error(void);

/* compute_uncertainty, src/libs/active_learning.py:10-35. probs (B, C) -> out (B). */
int bfb200_compute_uncertainty(const float* probs, int64_t n_rows, int32_t n_classes,
                               int32_t mode, float* out, void* stream);

/*
 * Acquisition reduction over MC draws. x is (T, B, C) probabilities (is_logp = 0)
 * or log-probabilities (is_logp = 1).  Any output may be NULL.
 *   mean_p          (B, C) : predictive mean (src/libs/visualization.py:113)
 *   score_of_mean   (3, B) : max / margin / entropy of mean_p
 *   mean_of_scores  (3, B) : mean over draws of the per-draw score
 *                            (what dropout_uncertainty accumulates, active_learning.py:152-155)
 */
int bfb200_acquire(const float* x, int32_t is_logp, int32_t n_draws, int64_t n_items,
                   int32_t n_classes, float* mean_p, float* score_of_mean,
                   float* mean_of_scores, void* stream);

/*
 * torch.topk(scores, k) of select_samples (active_learning.py:210), largest first,
 * ties broken towards the LOWEST index.  Output: k packed 64-bit keys, descending:
 *   key = (orderable(score) << 32) | (0xFFFFFFFF - (index_base + i))
 * so that keys from several GPUs merge by plain integer comparison.
 */
size_t bfb200_topk_workspace_bytes(int64_t n, int32_t k);
int bfb200_topk(const float* scores, int64_t n, int32_t k, int64_t index_base,
                uint64_t* keys_out, void* workspace, size_t workspace_bytes, void* stream);
/* top-k of already packed keys (the all-gathered per-GPU candidates). */
int bfb200_merge_topk(const uint64_t* keys, int64_t n, int32_t k, uint64_t* keys_out,
                      void* workspace, size_t workspace_bytes, void* stream);
/* keys -> (scores, int64 indices) */
int bfb200_unpack_keys(const uint64_t* keys, int32_t k, float* scores, int64_t* indices,
                       void* stream);

/*
 * MC-dropout MLP classifier, src/model/bayesian_classification.py:184-198 driven by
 * the loop of src/libs/visualization.py:105-113: n_draws x sigmoid(W2 drop(relu(W1 x + b1)) + b2).
 *   x (B, in_dim); w1 (hidden, in_dim); b1 (hidden); w2 (1, hidden); b2 (1)
 *   mask_bits: NULL (Philox) or [draw][item][ceil(hidden/32)] keep bits
 *   probs_out (T, B) or NULL; mean_p (B) or NULL;
 *   score_of_mean / mean_of_scores (3, B) over the 2-class distribution [1 - p, p], or NULL
 */
int bfb200_mlp_mc_score(const float* x, int64_t n_items, int32_t in_dim, int32_t hidden,
                        const float* w1, const float* b1, const float* w2, const float* b2,
                        float p_drop, int32_t n_draws, uint64_t seed, int64_t index_base,
                        const uint32_t* mask_bits, float* probs_out, float* mean_p,
                        float* score_of_mean, float* mean_of_scores, void* stream);

/*
 * Bayes-by-backprop MLP, src/model/bayesian_classification.py:118-130 with
 * LinearVariational.sampling (:31-44): per draw W = mu + eps * log1p(exp(rho)), one
 * weight draw shared by the whole pool.  eps (T, in_dim*hidden + hidden*1) fp32 [dev]
 * holds the N(0,1) noise for (hidden_layer.w, output_layer.w) of every draw.
 *   w1_mu/w1_rho (in_dim, hidden); b1 (hidden); w2_mu/w2_rho (hidden, 1); b2 (1)
 */
int bfb200_vmlp_mc_score(const float* x, int64_t n_items, int32_t in_dim, int32_t hidden,
                         const float* w1_mu, const float* w1_rho, const float* b1,
                         const float* w2_mu, const float* w2_rho, const float* b2,
                         const float* eps, int32_t n_draws, float* probs_out, float* mean_p,
                         float* score_of_mean, float* mean_of_scores, void* stream);

/*
 * MyTransformer.forward under eval()/no_grad (src/model/transformer.py:349-375):
 * src (len, B) int64 seq-first -> log-probs (len, B, V).  `draw` selects the Philox
 * stream of this call (masks->bits uses step 0, n_draws 1).
 */
size_t bfb200_transformer_workspace_bytes(const bfb200_transformer* model, int64_t n_sequences,
                                          int32_t max_len);
int bfb200_transformer_forward(const bfb200_transformer* model, const int64_t* src,
                               int32_t len, int64_t batch, const bfb200_masks* masks,
                               uint32_t draw, int64_t index_base, float* logp_out,
                               void* workspace, size_t workspace_bytes, void* stream);

/*
 * The MC scoring loop (see bfb200_mc_args).  The pool is processed in chunks that
 * fit `workspace_bytes`; bfb200_mc_workspace_bytes(items_per_chunk) sizes it and any
 * workspace >= the size for 1 item is accepted.
 */
size_t bfb200_mc_workspace_bytes(const bfb200_transformer* model, const bfb200_mc_args* args,
                                 int64_t items_per_chunk);
int bfb200_mc_score_transformer(const bfb200_transformer* model, const bfb200_mc_args* args,
                                void* workspace, size_t workspace_bytes, void* stream);

/*
 * 16-bit weight image of the tensor-core paths (model->compute = F16 / BF16), written into model->fused_image's
 * buffer.  Fused class: every nn.Linear weight of src/model/transformer.py as fp16 in the 128-byte-swizzled K-major
 * shared-memory layout tcgen05.mma reads (W_out stacked with the merged matrix (W1 gamma1) W_out, W1 scaled by
 * gamma1, lo parts of W_qkv and W2), followed by the fp32 LayerNorm / bias / folding vectors and a global-only tail
 * (vocabulary head, padded head bias, W_qkv lo rows) that the kernel streams per use; DESIGN.md section 5.  Layered tensor-core path: plain row-major 16-bit copies [W_qkv | W_out | FFN W1 | FFN W2 | head] that the
 * GEMM kernel reads through TMA.  _bytes returns 0 when no tensor-core path serves the model's shape.  Re-pack
 * after every weight update.
 */
size_t bfb200_fused_image_bytes(const bfb200_transformer* model);
/* which implementation model->compute selects for this shape: 0 layered fp32 kernels, 1 fused MC-forward kernel,
 * 2 layered tensor-core path, -1 none (BFB200_ERR_UNSUPPORTED at call time) */
int bfb200_compute_path(const bfb200_transformer* model);
int bfb200_fused_pack(const bfb200_transformer* model, void* image, size_t image_bytes, void* stream);

/*
 * MC-batched tensor-core linear layer, nn.Linear of src/model/transformer.py (:33-35, :100, :152-154, :336) over the
 * token rows of a chunk: C[M, N] = A[M, K] W[N, K]^T (+ bias) (ReLU).  A and W are 16-bit [dev] matrices (fmt 0 =
 * fp16, 1 = bf16), K contiguous, A rows `a_pitch` elements apart (a pitch of len * K with A offset to the last
 * position evaluates the vocabulary head on last positions only).  out16 (same format) and/or out32 (fp32) receive
 * the result with a row pitch of `ldc` >= N elements (ldc * element size a multiple of 16 bytes: results leave the
 * chip through 2-D TMA stores); either may be NULL.  TMA-fed tcgen05 kernel, fp32 accumulation in TMEM.
 */
int bfb200_linear_tc(const void* A, int64_t a_pitch, const void* W, const float* bias, void* out16, float* out32,
                     int64_t ldc, int64_t M, int32_t N, int32_t K, int32_t relu, int32_t fmt, void* stream);

/*
 * Causal multi-head attention of src/model/transformer.py:58-65, :123-124 on tcgen05 tensor cores, for heads of
 * width 64 and sequences of at most 256 tokens: qkv (n_seq * len, 3 * d_model) 16-bit [dev] with columns
 * [Q heads | K heads | V heads] -> out (n_seq * len, d_model) 16-bit [dev] (heads concatenated).  fmt 0 = fp16, 1 = bf16.
 */
int bfb200_attention_tc(const void* qkv, void* out, int64_t n_seq, int32_t len, int32_t d_model, int32_t n_heads,
                        int32_t fmt, void* stream);

/*
 * One fine-tune step of the reference's training loop (src/libs/utils.py:68-78: model(batch) in train mode,
 * CrossEntropyLoss on the log-softmax outputs of the answer positions, backward, AdamW.step()) for the notebook
 * model class (d_model 64, 8 heads) as ONE kernel launch — forward, loss, backward and the parameter update.
 * The parameters stay in the module's own tensors: `segments` maps them to the flat layout the kernel works in
 * (bfb200_train_flat_offset gives the layout; per-head W_q/W_k/W_v are rows [c*64 + h*8, +8) of the layer's W_qkv
 * block, c = 0 Q, 1 K, 2 V).  adam_m / adam_v persist between steps (zero before the first), `step` counts from 1.
 * Dropout masks (nn.Dropout at rate p_train on the positional-encoding output and on both branch outputs,
 * transformer.py:231-233, :272; the always-on bayes sites at rate p_bayes, :356, :370) are drawn from Philox with key
 * `seed`; both rates 0 give the deterministic step the parity test compares with eager autograd.
 */
typedef struct bfb200_train_segment {
  int32_t offset; /* first float of the segment in the flat layout */
  int32_t count;  /* floats */
  float* ptr;     /* [dev] the parameter tensor's storage (contiguous) */
} bfb200_train_segment;

typedef struct bfb200_train_args {
  int32_t ntoken, d_model, n_heads, d_ff, n_layers, pe_len;
  int32_t batch, seq_len, prompt_len; /* tokens (seq_len, batch): prompt_len prompt rows, then the answer rows */
  int32_t n_segments;
  int32_t step;                        /* AdamW step number of this call (1, 2, ...) */
  float lr, beta1, beta2, eps, weight_decay;
  float p_train, p_bayes, ln_eps;
  uint64_t seed;
  const bfb200_train_segment* segments; /* [dev] n_segments entries covering the flat layout exactly once */
  const int64_t* tokens;                /* [dev] (seq_len, batch) */
  const float* pe;                      /* [dev] (pe_len, d_model) */
  float* flat;                          /* [dev] bfb200_train_state_floats() floats: parameter scratch */
  float* grad;                          /* [dev] same size: gradients of this step (output, for inspection) */
  float* adam_m;                        /* [dev] same size, persistent */
  float* adam_v;                        /* [dev] same size, persistent */
  float* scratch;                       /* [dev] bfb200_train_scratch_floats() floats */
  float* loss_out;                      /* [dev] 1 float: += mean loss of the batch */
} bfb200_train_args;

size_t bfb200_train_state_floats(int32_t ntoken, int32_t d_ff, int32_t n_layers);
size_t bfb200_train_scratch_floats(int32_t batch, int32_t seq_len, int32_t prompt_len, int32_t ntoken, int32_t d_ff,
                                   int32_t n_layers);
/* flat offset of a parameter: which = 0 embedding, 1 W_qkv, 2 W_out, 3 norm1.weight, 4 norm1.bias, 5 ff.net.0.weight,
 * 6 ff.net.0.bias, 7 ff.net.2.weight, 8 ff.net.2.bias, 9 norm2.weight, 10 norm2.bias (1..10: of `layer`),
 * 11 linear_out.weight, 12 linear_out.bias */
int bfb200_train_flat_offset(int32_t ntoken, int32_t d_ff, int32_t n_layers, int32_t layer, int32_t which);
int bfb200_train_step(const bfb200_train_args* args, void* stream);

/* torch.mean(uncertainties, dim=0) of select_samples (src/libs/active_learning.py:209):
 * step_scores (n_steps, n_items) -> pool_scores (n_items), steps summed in order then divided. */
int bfb200_step_mean(const float* step_scores, int32_t n_steps, int64_t n_items, float* pool_scores,
                     void* stream);

/* number of kernel launches issued by this library on the calling thread since
 * the last bfb200_reset_launch_count() (bench.py's gpu_launches). */
int64_t bfb200_launch_count(void);
void bfb200_reset_launch_count(void);

/*
 * Per-kernel device timing for bench.py's roofline leg (no reference counterpart).  Between
 * _begin and _end every kernel launch of the calling thread is followed by a CUDA event on its stream;
 * _end synchronises and writes {"kernel_name": {"launches": n, "ms": total}, ...} into json_out.
 */
int bfb200_profile_begin(void* stream);
int bfb200_profile_end(char* json_out, size_t json_bytes);

#ifdef __cplusplus
}
#endif
#endif /* BAYESFORMER_B200_H */

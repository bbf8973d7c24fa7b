"""Host side of the native path: packs module weights for the C ABI, owns (torch-allocated)
workspaces and wraps every entry point of include/bayesformer_b200.h in a tensor-level call.

PyTorch is used here for device memory, streams and (in sharding.py) torch.distributed only;
all arithmetic of the hot path happens inside libbayesformer_b200.so.
"""

from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import torch

from bayesformer_b200 import _lib

# Every stochastic native call gets its own Philox stream, the way each F.dropout call of the reference advances the
# global generator.  The stream id is taken from — and advances — the philox offset of torch's default CUDA generator,
# so `torch.manual_seed(s)` restarts the sequence exactly as it does for the reference (same seed => same masks, same
# acquired indices), and ranks that ran different numbers of unrelated native calls still need the explicit `seed=`
# argument to agree.  (Fallback without that generator API: a process-global counter.)
_call_counter = 0


def _next_call_id(device=None) -> int:
    global _call_counter
    try:
        idx = torch.cuda.current_device() if device is None or device.index is None else device.index
        gen = torch.cuda.default_generators[idx]
        off = int(gen.get_offset())
        gen.set_offset(off + 4)          # one Philox block, like a small torch random op
        return off // 4 + 1
    except (AttributeError, IndexError, RuntimeError):
        _call_counter += 1
        return _call_counter


def default_seed() -> int:
    """Philox key of the next stochastic call: torch's CUDA seed (torch.manual_seed drives it)."""
    return int(torch.cuda.initial_seed()) & 0xFFFFFFFFFFFFFFFF


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: bayesformer_b200 has no CPU path")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(dtype=torch.float32).contiguous()


# ---------------------------------------------------------------------------
# acquisition / top-k
# ---------------------------------------------------------------------------


def compute_uncertainty(probs: torch.Tensor, mode: str = "max") -> torch.Tensor:
    """Native `compute_uncertainty` (reference src/libs/active_learning.py:10-35) for (B, C) CUDA tensors."""
    if mode not in _lib.MODE:
        raise ValueError(f"Unknown mode {mode}. Available modes are 'max', 'margin' and 'entropy'.")
    _require_cuda(probs, "probs")
    if probs.dim() != 2:
        raise IndexError("compute_uncertainty expects a 2-D (batch, classes) tensor")
    p = _f32c(probs)
    out = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
    lib = _lib.load()
    with torch.cuda.device(p.device):
        _lib.check(lib.bfb200_compute_uncertainty(p.data_ptr(), p.shape[0], p.shape[1], _lib.MODE[mode],
                                                  out.data_ptr(), _stream_ptr(p.device)))
    return out


@dataclass
class Acquisition:
    mean_p: torch.Tensor | None          # (B, C)
    score_of_mean: torch.Tensor | None   # (3, B): max, margin, entropy of the predictive mean
    mean_of_scores: torch.Tensor  # (3, B): mean over draws of the per-draw scores


def acquire(x: torch.Tensor, is_logp: bool = False, want_mean_p: bool = True) -> Acquisition:
    """Reduce MC outputs x (T, B, C) to predictive mean + all three scores in one pass.  With
    `want_mean_p=False` only the mean-of-scores (what dropout_uncertainty accumulates) is produced and the
    predictive-mean outputs are None."""
    _require_cuda(x, "x")
    if x.dim() != 3:
        raise ValueError("acquire expects (draws, items, classes)")
    x = _f32c(x)
    T, B, Cn = x.shape
    mean_p = torch.empty(B, Cn, dtype=torch.float32, device=x.device) if want_mean_p else None
    som = torch.empty(3, B, dtype=torch.float32, device=x.device) if want_mean_p else None
    mos = torch.empty(3, B, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.bfb200_acquire(x.data_ptr(), int(is_logp), T, B, Cn, _ptr(mean_p), _ptr(som),
                                      mos.data_ptr(), _stream_ptr(x.device)))
    return Acquisition(mean_p, som, mos)


def topk_keys(scores: torch.Tensor, k: int, index_base: int = 0) -> torch.Tensor:
    """k packed (score, index) keys of the largest scores, descending, lowest index first on ties."""
    _require_cuda(scores, "scores")
    s = _f32c(scores).reshape(-1)
    n = s.numel()
    if k > n:
        raise RuntimeError("selected index k out of range")
    lib = _lib.load()
    keys = torch.empty(k, dtype=torch.int64, device=s.device)
    ws_bytes = lib.bfb200_topk_workspace_bytes(n, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(lib.bfb200_topk(s.data_ptr(), n, k, index_base, keys.data_ptr(), ws.data_ptr(), ws_bytes,
                                   _stream_ptr(s.device)))
    return keys


def merge_topk_keys(keys: torch.Tensor, k: int) -> torch.Tensor:
    """Final k of already packed candidate keys (e.g. the all-gathered per-GPU top-k)."""
    _require_cuda(keys, "keys")
    kk = keys.contiguous().reshape(-1)
    n = kk.numel()
    if k > n:
        raise RuntimeError("selected index k out of range")
    lib = _lib.load()
    out = torch.empty(k, dtype=torch.int64, device=kk.device)
    ws_bytes = lib.bfb200_topk_workspace_bytes(n, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=kk.device)
    with torch.cuda.device(kk.device):
        _lib.check(lib.bfb200_merge_topk(kk.data_ptr(), n, k, out.data_ptr(), ws.data_ptr(), ws_bytes,
                                         _stream_ptr(kk.device)))
    return out


def unpack_keys(keys: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """keys -> (scores fp32, indices int64)"""
    _require_cuda(keys, "keys")
    k = keys.numel()
    scores = torch.empty(k, dtype=torch.float32, device=keys.device)
    idx = torch.empty(k, dtype=torch.int64, device=keys.device)
    lib = _lib.load()
    with torch.cuda.device(keys.device):
        _lib.check(lib.bfb200_unpack_keys(keys.data_ptr(), k, scores.data_ptr(), idx.data_ptr(),
                                          _stream_ptr(keys.device)))
    return scores, idx


def topk(scores: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for `torch.topk(scores, k)` on a 1-D CUDA tensor (reference active_learning.py:210)."""
    return unpack_keys(topk_keys(scores, k))


def linear_tc(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, relu: bool = False,
              out_dtype: torch.dtype | None = None, a_pitch: int | None = None, rows: int | None = None) -> torch.Tensor:
    """`a @ w.T (+ bias) (ReLU)` on the tcgen05 tensor cores (bfb200_linear_tc).  a (M, K) and w (N, K) are fp16 or
    bf16 CUDA tensors of the same dtype; the result is `out_dtype` (the operand dtype by default, or torch.float32).
    `a_pitch` / `rows` address M = rows rows of K elements that are `a_pitch` elements apart starting at a's first
    element (used to evaluate the vocabulary head on last positions only)."""
    _require_cuda(a, "a")
    if a.dtype not in (torch.float16, torch.bfloat16) or w.dtype != a.dtype:
        raise TypeError("linear_tc takes fp16 or bf16 operands of one dtype")
    w = w.contiguous()
    N, K = w.shape
    if a_pitch is None:
        a = a.contiguous()
        M, a_pitch = a.shape[0], K
    else:
        M = int(rows)
    out_dtype = out_dtype or a.dtype
    align = 4 if out_dtype == torch.float32 else 8          # output rows leave through TMA: pitch * size % 16 == 0
    ldc = (N + align - 1) // align * align
    out = torch.empty(M, ldc, dtype=out_dtype, device=a.device)
    fmt = 0 if a.dtype == torch.float16 else 1
    b = None if bias is None else _f32c(bias).to(a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.bfb200_linear_tc(a.data_ptr(), a_pitch, w.data_ptr(), _ptr(b),
                                        out.data_ptr() if out_dtype != torch.float32 else None,
                                        out.data_ptr() if out_dtype == torch.float32 else None,
                                        ldc, M, N, K, int(relu), fmt, _stream_ptr(a.device)))
    return out if ldc == N else out[:, :N]


def attention_tc(qkv: torch.Tensor, n_seq: int, length: int, n_heads: int) -> torch.Tensor:
    """Causal multi-head attention on the tcgen05 tensor cores (bfb200_attention_tc): qkv (n_seq * length, 3 d) fp16 or
    bf16 with columns [Q | K | V] -> (n_seq * length, d).  Heads of width 64, length <= 256."""
    _require_cuda(qkv, "qkv")
    if qkv.dtype not in (torch.float16, torch.bfloat16):
        raise TypeError("attention_tc takes fp16 or bf16")
    qkv = qkv.contiguous()
    d = qkv.shape[1] // 3
    out = torch.empty(qkv.shape[0], d, dtype=qkv.dtype, device=qkv.device)
    lib = _lib.load()
    with torch.cuda.device(qkv.device):
        _lib.check(lib.bfb200_attention_tc(qkv.data_ptr(), out.data_ptr(), n_seq, length, d, n_heads,
                                           0 if qkv.dtype == torch.float16 else 1, _stream_ptr(qkv.device)))
    return out


def step_mean(step_scores: torch.Tensor) -> torch.Tensor:
    """`torch.mean(uncertainties, dim=0)` of select_samples (reference active_learning.py:209): (N, B) -> (B,)."""
    _require_cuda(step_scores, "step_scores")
    s = _f32c(step_scores)
    if s.dim() != 2:
        raise ValueError("step_scores must be (new_tokens, B)")
    out = torch.empty(s.shape[1], dtype=torch.float32, device=s.device)
    lib = _lib.load()
    with torch.cuda.device(s.device):
        _lib.check(lib.bfb200_step_mean(s.data_ptr(), s.shape[0], s.shape[1], out.data_ptr(), _stream_ptr(s.device)))
    return out


# ---------------------------------------------------------------------------
# toy classifiers (src/model/bayesian_classification.py)
# ---------------------------------------------------------------------------


@dataclass
class ClassifierScores:
    probs: torch.Tensor | None    # (T, B) per-draw sigmoid outputs
    mean_p: torch.Tensor          # (B,)
    score_of_mean: torch.Tensor   # (3, B)
    mean_of_scores: torch.Tensor  # (3, B)


def mlp_mc_score(x, w1, b1, w2, b2, p_drop: float, n_draws: int, seed: int | None = None, index_base: int = 0,
                 mask_bits: torch.Tensor | None = None, return_probs: bool = True) -> ClassifierScores:
    """n_draws stochastic passes of MLP.forward (bayesian_classification.py:184-198) + sigmoid + reductions."""
    _require_cuda(x, "x")
    x = _f32c(x)
    w1, b1, w2, b2 = (_f32c(t).to(x.device) for t in (w1, b1, w2, b2))
    B, in_dim = x.shape
    hidden = w1.shape[0]
    if w2.numel() != hidden:
        raise NotImplementedError("native MLP scoring supports output_dimension == 1")
    dev = x.device
    probs = torch.empty(n_draws, B, dtype=torch.float32, device=dev) if return_probs else None
    mean_p = torch.empty(B, dtype=torch.float32, device=dev)
    som = torch.empty(3, B, dtype=torch.float32, device=dev)
    mos = torch.empty(3, B, dtype=torch.float32, device=dev)
    if seed is None:
        seed = default_seed() ^ (_next_call_id() << 32)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.bfb200_mlp_mc_score(
            x.data_ptr(), B, in_dim, hidden, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
            float(p_drop), n_draws, seed & 0xFFFFFFFFFFFFFFFF, index_base, _ptr(mask_bits), _ptr(probs),
            mean_p.data_ptr(), som.data_ptr(), mos.data_ptr(), _stream_ptr(dev)))
    return ClassifierScores(probs, mean_p, som, mos)


def vmlp_mc_score(x, w1_mu, w1_rho, b1, w2_mu, w2_rho, b2, eps: torch.Tensor,
                  return_probs: bool = True) -> ClassifierScores:
    """VariationalMLP draws (bayesian_classification.py:118-130); eps (T, in*hidden + hidden) N(0,1) noise."""
    _require_cuda(x, "x")
    x = _f32c(x)
    dev = x.device
    w1_mu, w1_rho, b1, w2_mu, w2_rho, b2, eps = (_f32c(t).to(dev) for t in (w1_mu, w1_rho, b1, w2_mu, w2_rho, b2, eps))
    B, in_dim = x.shape
    hidden = w1_mu.shape[1]
    if w2_mu.numel() != hidden:
        raise NotImplementedError("native VariationalMLP scoring supports output_dimension == 1")
    n_draws = eps.shape[0]
    if eps.shape[1] != in_dim * hidden + hidden:
        raise ValueError("eps must be (T, in_dim * hidden + hidden)")
    probs = torch.empty(n_draws, B, dtype=torch.float32, device=dev) if return_probs else None
    mean_p = torch.empty(B, dtype=torch.float32, device=dev)
    som = torch.empty(3, B, dtype=torch.float32, device=dev)
    mos = torch.empty(3, B, dtype=torch.float32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.bfb200_vmlp_mc_score(
            x.data_ptr(), B, in_dim, hidden, w1_mu.data_ptr(), w1_rho.data_ptr(), b1.data_ptr(), w2_mu.data_ptr(),
            w2_rho.data_ptr(), b2.data_ptr(), eps.data_ptr(), n_draws, _ptr(probs), mean_p.data_ptr(),
            som.data_ptr(), mos.data_ptr(), _stream_ptr(dev)))
    return ClassifierScores(probs, mean_p, som, mos)


# ---------------------------------------------------------------------------
# transformer
# ---------------------------------------------------------------------------


class PackedTransformer:
    """Device-resident, C-ABI-shaped copy of a MyTransformer's weights.

    Per-head W_q/W_k/W_v (reference transformer.py:33-35) are concatenated into one (3d, d)
    matrix per layer.  The pack is cached on the module and rebuilt when any parameter's
    `_version` or storage changes (weights change between active-learning rounds).
    """

    def __init__(self, module, device: torch.device, compute: str = "fp32"):
        self.device = device
        L = len(module.layers)
        d = module.d_model
        first = module.layers[0] if L else None
        H = first.self_attn.n_heads if L else 1
        F = first.ff.net[0].out_features if L else 4

        def dev(t):
            # an OWNED copy: a no-copy view of a live parameter would follow in-place updates (optimizer steps, graph
            # replays) while the stacked / 16-bit copies below do not, leaving the pack half fresh
            return t.detach().to(device=device, dtype=torch.float32, copy=True).contiguous()

        def stack(fn):
            return torch.stack([dev(fn(layer)) for layer in module.layers]).contiguous() if L else None

        self.embedding = dev(module.embedding.weight)
        self.pe = dev(module.pos_enc.pe[:, 0, :])
        self.w_qkv = stack(lambda l: torch.cat(
            [torch.cat([getattr(h, name).weight for h in l.self_attn.heads], 0) for name in ("W_q", "W_k", "W_v")], 0))
        self.w_out = stack(lambda l: l.self_attn.W_out.weight)
        self.ln1_w = stack(lambda l: l.norm1.weight)
        self.ln1_b = stack(lambda l: l.norm1.bias)
        self.ff1_w = stack(lambda l: l.ff.net[0].weight)
        self.ff1_b = stack(lambda l: l.ff.net[0].bias)
        self.ff2_w = stack(lambda l: l.ff.net[2].weight)
        self.ff2_b = stack(lambda l: l.ff.net[2].bias)
        self.ln2_w = stack(lambda l: l.norm2.weight)
        self.ln2_b = stack(lambda l: l.norm2.bias)
        self.head_w = dev(module.linear_out.weight)
        self.head_b = dev(module.linear_out.bias)

        flags, p = site_flags_of(module)
        eps = first.norm1.eps if L else 1e-5
        self.desc = _lib.Transformer(
            ntoken=module.ntoken, d_model=d, n_heads=H, d_ff=F, n_layers=L, pe_len=self.pe.shape[0],
            p_drop=float(p), site_flags=flags,
            embedding=self.embedding.data_ptr(), pe=self.pe.data_ptr(),
            w_qkv=_ptr(self.w_qkv), w_out=_ptr(self.w_out), ln1_w=_ptr(self.ln1_w), ln1_b=_ptr(self.ln1_b),
            ff1_w=_ptr(self.ff1_w), ff1_b=_ptr(self.ff1_b), ff2_w=_ptr(self.ff2_w), ff2_b=_ptr(self.ff2_b),
            ln2_w=_ptr(self.ln2_w), ln2_b=_ptr(self.ln2_b), head_w=self.head_w.data_ptr(),
            head_b=self.head_b.data_ptr(), ln_eps=float(eps), compute=_lib.COMPUTE[compute], fused_image=None)
        self.ntoken = module.ntoken
        self.d_model = d
        self.n_layers = L
        self.compute = compute
        self._workspace: torch.Tensor | None = None
        self.fused_image: torch.Tensor | None = None
        self.path = 0   # 0 layered fp32, 1 fused kernel, 2 layered tensor-core path
        if compute in ("fp16", "bf16"):
            lib = _lib.load()
            self.path = lib.bfb200_compute_path(C.byref(self.desc))
            nbytes = lib.bfb200_fused_image_bytes(C.byref(self.desc))
            if self.path < 0 or nbytes == 0:
                raise NotImplementedError(
                    f"compute={compute!r} needs the fused kernel's class (fp16: d_model 64, 8 heads, dim_feedforward <= 64, "
                    "ntoken <= 128) or heads of width 64 with dim_feedforward % 8 == 0 (layered tensor-core path); "
                    f"got d={d} H={H} F={F} V={module.ntoken} L={L}")
            self.fused_image = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self.desc.fused_image = self.fused_image.data_ptr()
            with torch.cuda.device(device):
                _lib.check(lib.bfb200_fused_pack(C.byref(self.desc), self.fused_image.data_ptr(), nbytes,
                                                 _stream_ptr(device)))

    def path_for(self, compute: str) -> int:
        """Implementation the library would pick for this model under `compute` (see bfb200_compute_path)."""
        probe = _lib.Transformer.from_buffer_copy(self.desc)
        probe.compute = _lib.COMPUTE[compute]
        return _lib.load().bfb200_compute_path(C.byref(probe))

    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._workspace


def site_flags_of(module) -> tuple[int, float]:
    """Which always-on dropout sites of reference transformer.py are live for this module tree.

    Under the public MyTransformer constructor only the embedding and layer-output sites can be
    live (reference :324-335 never forwards `bayes` to the blocks); the block-level sites are
    reachable by setting the sub-modules' attributes by hand and are honoured when every layer
    agrees and a single rate is used.
    """
    flags = 0
    rates = set()
    if getattr(module, "bayes", False):
        flags |= _lib.SITE_EMBED | _lib.SITE_LAYER_OUT
        rates.add(module.bayes_dropout)
    per_layer = []
    for layer in module.layers:
        lf = 0
        if getattr(layer, "bayes", False):
            lf |= _lib.SITE_ATTN_RESID | _lib.SITE_FFN_RESID
            rates.add(layer.bayes_dropout)
        if layer.ff.bayes_dropout is not None:
            lf |= _lib.SITE_FFN_IN
            rates.add(layer.ff.bayes_dropout)
        if getattr(layer.self_attn, "bayes", False):
            lf |= _lib.SITE_HEAD_OUT
            rates.add(layer.self_attn.bayes_dropout)
        head_flags = [bool(getattr(h, "bayes", False)) for h in layer.self_attn.heads]
        if any(head_flags):   # Head.bayes (reference transformer.py:50-53): three input masks per head, fp32 path only
            if not all(head_flags):
                raise NotImplementedError("native path needs Head.bayes set on all heads of a layer or on none")
            lf |= _lib.SITE_HEAD_IN
            rates.update(h.bayes_dropout for h in layer.self_attn.heads)
        per_layer.append(lf)
    if len(set(per_layer)) > 1:
        raise NotImplementedError("native path needs the same latent dropout sites in every layer")
    if per_layer:
        flags |= per_layer[0]
    if not flags:
        return 0, 0.0
    if len(rates) != 1:
        raise NotImplementedError(f"native path needs one dropout rate for all sites, got {sorted(map(str, rates))}")
    p = rates.pop()
    if p is None or not (0.0 <= float(p) < 1.0):
        raise ValueError(f"dropout probability has to be in [0, 1), but got {p}")
    return flags, float(p)


# module -> (signature, pack).  Kept outside the module so copy.deepcopy / pickling of the
# nn.Module (active_learning.ipynb cell 24) never sees ctypes pointers.
_PACKS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _param_signature(module, device) -> tuple:
    # `_bfb_weights_generation` is bumped by code that updates the weights without going through the dispatcher (a CUDA
    # graph replay of the optimizer step does not touch `_version`): see mark_weights_changed
    return ((str(device), getattr(module, "_bfb_weights_generation", 0)) +
            tuple((p.data_ptr(), p._version) for p in module.parameters()) + site_flags_of(module))


def mark_weights_changed(module) -> None:
    """Invalidate the packed / 16-bit weight images of `module` (weights were updated behind autograd's back, e.g. by
    replaying a captured optimizer step)."""
    module._bfb_weights_generation = getattr(module, "_bfb_weights_generation", 0) + 1


def packed(module, device: torch.device, compute: str = "fp32") -> PackedTransformer:
    """Cached PackedTransformer of `module` for `device` and `compute` (rebuilt after any in-place weight
    update or change of the dropout configuration)."""
    sig = _param_signature(module, device)
    per_module = _PACKS.get(module)
    if per_module is None:
        per_module = {}
        _PACKS[module] = per_module
    entry = per_module.get(compute)
    if entry is None or entry[0] != sig:
        entry = (sig, PackedTransformer(module, device, compute))
        per_module[compute] = entry
    return entry[1]


def default_compute(module, scheme: str) -> str:
    """Arithmetic class of a scoring call that does not name one.  `module.native_compute` ("fp32" | "fp16" | "bf16" |
    "auto") wins when set.  Otherwise the stochastic schemes (MC dropout, temperature sampling) take 'auto' — a
    tensor-core path where the model is in its class: its error (1e-2 class, measured 5e-4 rms on log-probs) is far
    below their sampling noise — while the DETERMINISTIC greedy scheme (greedy_uncertainty, generate, evaluate) takes
    the exact fp32 CUDA path: with no MC noise to absorb it, a near-tie arg-max could otherwise differ from the
    reference's fp32 decision.  Set `model.native_compute = "fp16"` to opt the greedy calls into the fused kernel."""
    explicit = getattr(module, "native_compute", None)
    if explicit is not None:
        return explicit
    return "fp32" if scheme == "greedy" else "auto"


FUSED_MAX_LEN = 128   # longest sequence (P + N - 1) the fused fp16 kernel takes (csrc/mc_fused.cu FMAXLEN)
TC_MAX_LEN = 256     # ... and the layered tensor-core path


def scoring_pack(module, device: torch.device, prompt_len: int, new_tokens: int, compute: str = "auto") -> PackedTransformer:
    """Pack used by the MC-scoring loop.  compute='auto' picks a tensor-core path whenever the model and sequence
    lengths are in its class — the fused fp16 kernel for the reference notebooks' configuration, the layered
    TMA/tcgen05 path for models with heads of width 64 — and the exact fp32 CUDA path otherwise; 'fp32' / 'fp16' /
    'bf16' force one (raising NotImplementedError outside the class)."""
    if compute not in ("auto", "fp32", "fp16", "bf16"):
        raise ValueError(f"unknown compute {compute!r}")
    if compute == "auto":
        base = packed(module, device, "fp32")
        path, longest = base.path_for("fp16"), prompt_len + new_tokens - 1
        if (path == 1 and longest <= FUSED_MAX_LEN) or (path == 2 and longest <= TC_MAX_LEN):
            return packed(module, device, "fp16")
        return base
    return packed(module, device, compute)


def transformer_forward(pack: PackedTransformer, src: torch.Tensor, seed: int | None = None, draw: int | None = None,
                        index_base: int = 0, mask_bits: torch.Tensor | None = None) -> torch.Tensor:
    """MyTransformer.forward under eval/no_grad: src (len, B) int64 CUDA -> log-probs (len, B, V)."""
    _require_cuda(src, "src")
    if src.dim() != 2:
        raise ValueError("src must be (T, B)")
    src = src.to(torch.int64).contiguous()
    length, batch = src.shape
    dev = src.device
    out = torch.empty(length, batch, pack.ntoken, dtype=torch.float32, device=dev)
    if length == 0 or batch == 0:
        return out
    lib = _lib.load()
    nbytes = lib.bfb200_transformer_workspace_bytes(C.byref(pack.desc), batch, length)
    ws = pack.workspace(nbytes)
    masks = _lib.Masks(seed=(default_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF,
                       bits=_ptr(mask_bits), max_len=length)
    if draw is None:
        draw = _next_call_id() & 0xFFFFFFFF
    with torch.cuda.device(dev):
        _lib.check(lib.bfb200_transformer_forward(C.byref(pack.desc), src.data_ptr(), length, batch, C.byref(masks),
                                                  draw, index_base, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  _stream_ptr(dev)))
    return out


class profile:
    """Context manager around bfb200_profile_begin/_end: per-kernel launch counts and device milliseconds
    of everything the library launched on the current stream inside the block (bench.py's roofline leg)."""

    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self.kernels: dict = {}

    def __enter__(self):
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().bfb200_profile_begin(_stream_ptr(self.device)))
        return self

    def __exit__(self, *exc):
        import json

        buf = C.create_string_buffer(1 << 16)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().bfb200_profile_end(buf, len(buf)))
        self.kernels = json.loads(buf.value.decode())
        return False


@dataclass
class McResult:
    step_scores: torch.Tensor        # (N, B)
    pool_scores: torch.Tensor        # (B,)
    logp: torch.Tensor | None        # (N, B*T, V)
    tokens: torch.Tensor | None      # (B*T, P+N)


def mc_score_transformer(pack: PackedTransformer, prompts: torch.Tensor, new_tokens: int, n_draws: int,
                         scheme: str, mode: str, temperature: float = 0.8, seed: int | None = None,
                         index_base: int = 0, mask_bits: torch.Tensor | None = None, mask_max_len: int = 0,
                         forced_tokens: torch.Tensor | None = None, return_logp: bool = False,
                         return_tokens: bool = False, workspace_bytes: int | None = None) -> McResult:
    """The loops of greedy/sampling/dropout_uncertainty (reference active_learning.py:38-155) with the
    draws folded into the batch; prompts (P, B) int64 CUDA, seq-first as `get_batch` returns."""
    if mode not in _lib.MODE:
        raise ValueError(f"Unknown mode {mode}. Available modes are 'max', 'margin' and 'entropy'.")
    if scheme not in _lib.SCHEME:
        raise ValueError(
            f"Unknown compute mode: {scheme}. For Transformer, use 'greedy' or 'sampling'. For BayesFormer, use 'dropout'.")
    _require_cuda(prompts, "prompts")
    prompts = prompts.to(torch.int64).contiguous()
    P, B = prompts.shape
    dev = prompts.device
    N, T, V = int(new_tokens), int(n_draws), pack.ntoken
    step_scores = torch.zeros(N, B, dtype=torch.float32, device=dev)
    pool_scores = torch.zeros(B, dtype=torch.float32, device=dev)
    logp = torch.empty(N, B * T, V, dtype=torch.float32, device=dev) if return_logp else None
    tokens = torch.empty(B * T, P + N, dtype=torch.int32, device=dev) if return_tokens else None
    if forced_tokens is not None:
        forced_tokens = forced_tokens.to(device=dev, dtype=torch.int32).contiguous()
        if tuple(forced_tokens.shape) != (N, B * T):
            raise ValueError("forced_tokens must be (new_tokens, B * n_draws)")
    if seed is None:
        seed = default_seed() ^ (_next_call_id() << 32)
    args = _lib.McArgs(
        prompts=prompts.data_ptr(), n_items=B, prompt_len=P, new_tokens=N, n_draws=T, scheme=_lib.SCHEME[scheme],
        mode=_lib.MODE[mode], temperature=float(temperature), index_base=index_base,
        masks=_lib.Masks(seed=seed & 0xFFFFFFFFFFFFFFFF, bits=_ptr(mask_bits), max_len=mask_max_len),
        forced_tokens=_ptr(forced_tokens), step_scores=step_scores.data_ptr(), pool_scores=pool_scores.data_ptr(),
        logp_out=_ptr(logp), tokens_out=_ptr(tokens))
    if B == 0 or N == 0:
        return McResult(step_scores, pool_scores, logp, tokens)
    lib = _lib.load()
    if workspace_bytes is None:
        workspace_bytes = default_mc_workspace_bytes(pack, args, B)
    ws = pack.workspace(workspace_bytes)
    with torch.cuda.device(dev):
        _lib.check(lib.bfb200_mc_score_transformer(C.byref(pack.desc), C.byref(args), ws.data_ptr(), workspace_bytes,
                                                   _stream_ptr(dev)))
    return McResult(step_scores, pool_scores, logp, tokens)


# Activations of one chunk are sized to stay well inside the 126 MB L2 where possible and never
# above a few GB of the 180 GB HBM; the C side loops over chunks.
_MAX_WORKSPACE_BYTES = 8 << 30


def default_mc_workspace_bytes(pack: PackedTransformer, args, n_items: int) -> int:
    lib = _lib.load()
    full = lib.bfb200_mc_workspace_bytes(C.byref(pack.desc), C.byref(args), max(1, n_items))
    if full <= _MAX_WORKSPACE_BYTES:
        return full
    one = lib.bfb200_mc_workspace_bytes(C.byref(pack.desc), C.byref(args), 1)
    return max(one, _MAX_WORKSPACE_BYTES)


# ---------------------------------------------------------------------------
# fused fine-tune step (reference src/libs/utils.py:68-78)
# ---------------------------------------------------------------------------


class FusedTrainer:
    """AdamW fine-tuning of a MyTransformer of the notebook class (d_model 64, 8 heads, default dropout sites) through
    `bfb200_train_step`: forward, loss, backward and the update of one batch per kernel launch.  Holds the optimizer
    state (flat first / second moments, step count) and the segment table that maps the module's parameter tensors to
    the kernel's flat layout; the parameters themselves stay in the module (updated in place, like torch.optim.AdamW).
    """

    BETAS, EPS, WEIGHT_DECAY = (0.9, 0.999), 1e-8, 0.01   # torch.optim.AdamW defaults, as the reference uses them
    MAX_SEQ = 16                                          # csrc/train_fused.cu TMAXT

    @staticmethod
    def supports(module) -> bool:
        try:
            flags, _p = site_flags_of(module)
        except (NotImplementedError, ValueError):
            return False
        if flags & ~(_lib.SITE_EMBED | _lib.SITE_LAYER_OUT):
            return False   # latent block-level sites: the eager / graphed path trains those
        heads = module.layers[0].self_attn.heads if len(module.layers) else []
        if not (len(module.layers) >= 1 and module.embedding.weight.shape[1] == 64 and len(heads) == 8
                and module.embedding.weight.is_cuda and module.embedding.weight.dtype == torch.float32):
            return False
        # one sequence's activations must fit a CTA's shared memory (0 = they do not)
        V, L, F = module.embedding.weight.shape[0], len(module.layers), module.layers[0].ff.net[0].weight.shape[0]
        return int(_lib.load().bfb200_train_scratch_floats(1, FusedTrainer.MAX_SEQ, 1, V, F, L)) > 0

    def __init__(self, module, lr: float):
        lib = _lib.load()
        self.module, self.lr = module, float(lr)
        dev = module.embedding.weight.device
        self.device = dev
        V = module.embedding.weight.shape[0]
        L = len(module.layers)
        F = module.layers[0].ff.net[0].weight.shape[0]
        self.V, self.F, self.L = V, F, L
        self.n_state = int(lib.bfb200_train_state_floats(V, F, L))
        off = lambda layer, which: int(lib.bfb200_train_flat_offset(V, F, L, layer, which))
        segs = [(off(0, 0), module.embedding.weight)]
        for i, layer in enumerate(module.layers):
            base = off(i, 1)
            for c, name in enumerate(("W_q", "W_k", "W_v")):
                for h, head in enumerate(layer.self_attn.heads):
                    segs.append((base + (c * 64 + h * 8) * 64, getattr(head, name).weight))
            segs += [(off(i, 2), layer.self_attn.W_out.weight), (off(i, 3), layer.norm1.weight), (off(i, 4), layer.norm1.bias),
                     (off(i, 5), layer.ff.net[0].weight), (off(i, 6), layer.ff.net[0].bias),
                     (off(i, 7), layer.ff.net[2].weight), (off(i, 8), layer.ff.net[2].bias),
                     (off(i, 9), layer.norm2.weight), (off(i, 10), layer.norm2.bias)]
        segs += [(off(0, 11), module.linear_out.weight), (off(0, 12), module.linear_out.bias)]
        covered = sum(p.numel() for _, p in segs)
        if covered != self.n_state or covered != sum(p.numel() for p in module.parameters()):
            raise RuntimeError("segment table does not cover the module's parameters")
        for _, p in segs:
            if not p.is_contiguous() or p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError("fused training needs contiguous fp32 parameters on one CUDA device")
        self._params = [p for _, p in segs]
        arr = (_lib.TrainSegment * len(segs))(*[_lib.TrainSegment(o, p.numel(), p.data_ptr()) for o, p in segs])
        self.segments = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.n_segments = len(segs)
        self.flat = torch.empty(self.n_state, dtype=torch.float32, device=dev)
        self.grad = torch.empty(self.n_state, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(self.n_state, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(self.n_state, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        self._scratch = None
        self.pe = module.pos_enc.pe.detach().reshape(-1, 64).to(dev, torch.float32).contiguous()

    def valid_for(self, module) -> bool:
        """The segment table holds raw pointers: still pointing at this module's (unmoved) parameters?"""
        return module is self.module and all(p.data_ptr() == int(s.ptr) for p, s in zip(self._params, self._segment_view()))

    def _segment_view(self):
        raw = bytes(self.segments.cpu().numpy().tobytes())
        return (_lib.TrainSegment * self.n_segments).from_buffer_copy(raw)

    def step(self, tokens: torch.Tensor, prompt_len: int, p_train: float, p_bayes: float, seed: int | None = None) -> None:
        """One AdamW step on the batch `tokens` (T, B) int64 = prompts followed by answers; the batch's mean loss is
        ADDED to `self.loss` (read it once per epoch)."""
        lib = _lib.load()
        T, B = tokens.shape
        tokens = tokens.to(self.device, torch.int64).contiguous()
        need = int(lib.bfb200_train_scratch_floats(B, T, prompt_len, self.V, self.F, self.L))
        if need == 0:
            raise ValueError(f"fused train step does not take this batch shape (T={T}, B={B}, prompt {prompt_len})")
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.float32, device=self.device)
        self.step_count += 1
        if seed is None:
            seed = default_seed() ^ (_next_call_id(self.device) << 32)
        b1, b2 = self.BETAS
        args = _lib.TrainArgs(
            ntoken=self.V, d_model=64, n_heads=8, d_ff=self.F, n_layers=self.L, pe_len=self.pe.shape[0],
            batch=B, seq_len=T, prompt_len=prompt_len, n_segments=self.n_segments, step=self.step_count,
            lr=self.lr, beta1=b1, beta2=b2, eps=self.EPS, weight_decay=self.WEIGHT_DECAY,
            p_train=float(p_train), p_bayes=float(p_bayes), ln_eps=1e-5, seed=seed & 0xFFFFFFFFFFFFFFFF,
            segments=self.segments.data_ptr(), tokens=tokens.data_ptr(), pe=self.pe.data_ptr(),
            flat=self.flat.data_ptr(), grad=self.grad.data_ptr(), adam_m=self.adam_m.data_ptr(),
            adam_v=self.adam_v.data_ptr(), scratch=self._scratch.data_ptr(), loss_out=self.loss.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(lib.bfb200_train_step(C.byref(args), _stream_ptr(self.device)))
        mark_weights_changed(self.module)   # in-place updates behind autograd's back: the scoring packs must follow

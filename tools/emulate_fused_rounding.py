"""CPU emulation of the fused kernel's fp16 rounding points (no GPU): which operands need hi + lo parts to give
the 1e-2 class real margin?  Runs the trained C2 fixture model in float64 with explicit fp16 roundings at the
places mc_fused.cu rounds, against the same arithmetic without any rounding, on the prompt sets of
tools/fused_accuracy.py (same dropout masks on both sides, the exact path's tokens forced).
usage: python tools/emulate_fused_rounding.py [variant ...]     (variants: see VARIANTS)"""
import math, os, sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from oracle import restatement as R

torch.set_num_threads(os.cpu_count())


def r16(t):
    return t.float().half().double()


def hilo(t):   # fp16 hi + fp16 lo
    hi = t.float().half()
    lo = (t.float() - hi.float()).half()
    return hi.double() + lo.double()


def forward(sd, src, keeps, p, cfg):
    """src (l, S); keeps: list of (l, S, d) bool keep masks [E, LO_0, ..]; cfg: dict operand -> rounding fn"""
    D = R.model_dims(sd)
    d, Hh, dh, L = D["d"], D["H"], D["dh"], D["L"]
    n = src.shape[0]
    rd = lambda name, t: cfg.get(name, lambda u: u)(t)
    scale = 1.0 / (1.0 - p)
    x = sd["embedding.weight"].double()[src]
    x = x * keeps[0].double() * scale * math.sqrt(d) + sd["pos_enc.pe"][:n].double()
    causal = torch.triu(torch.full((n, n), float("-inf"), dtype=torch.float64), diagonal=1)
    for i in range(L):
        pre = f"layers.{i}."
        wq = torch.cat([sd[f"{pre}self_attn.heads.{h}.W_q.weight"] for h in range(Hh)]).double()
        wk = torch.cat([sd[f"{pre}self_attn.heads.{h}.W_k.weight"] for h in range(Hh)]).double()
        wv = torch.cat([sd[f"{pre}self_attn.heads.{h}.W_v.weight"] for h in range(Hh)]).double()
        xa = rd("x_qkv", x)
        q = rd("q", xa @ rd("w_qk", wq).t()); k = rd("k", xa @ rd("w_qk", wk).t()); v = rd("v", xa @ rd("w_v", wv).t())
        S = x.shape[1]
        qh = q.view(n, S, Hh, dh).permute(1, 2, 0, 3); kh = k.view(n, S, Hh, dh).permute(1, 2, 0, 3)
        vh = v.view(n, S, Hh, dh).permute(1, 2, 0, 3)
        sc = qh @ kh.transpose(-1, -2) / math.sqrt(dh) + causal
        o = (torch.softmax(sc, -1) @ vh).permute(2, 0, 1, 3).reshape(n, S, d)
        attn = rd("attn_out", o) @ rd("w_out", sd[f"{pre}self_attn.W_out.weight"].double()).t()
        v1 = x + attn
        mean = v1.mean(-1, keepdim=True); var = v1.var(-1, unbiased=False, keepdim=True)
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        g1 = sd[f"{pre}norm1.weight"].double(); b1n = sd[f"{pre}norm1.bias"].double()
        W1 = sd[f"{pre}ff.net.0.weight"].double(); bb1 = sd[f"{pre}ff.net.0.bias"].double()
        if cfg.get("fold_ln1", True):
            w1g = rd("w1", W1 * g1)
            col_s = w1g.sum(-1); col_c = W1 @ b1n + bb1
            acc = rd("v1", v1) @ w1g.t()
            hdn = torch.relu(rstd * (acc - mean * col_s) + col_c)
        else:
            h1 = (v1 - mean) * rstd * g1 + b1n
            hdn = torch.relu(rd("v1", h1) @ rd("w1", W1).t() + bb1)
        ff = rd("h", hdn) @ rd("w2", sd[f"{pre}ff.net.2.weight"].double()).t() + sd[f"{pre}ff.net.2.bias"].double()
        y = x + ff
        x = torch.nn.functional.layer_norm(y, (d,), sd[f"{pre}norm2.weight"].double(), sd[f"{pre}norm2.bias"].double(), 1e-5)
        x = x * keeps[1 + i].double() * scale
    logits = rd("x_head", x[-1]) @ rd("w_head", sd["linear_out.weight"].double()).t() + sd["linear_out.bias"].double()
    return torch.log_softmax(logits, -1)


_orig_forward = forward


ALL16 = ["x_qkv", "w_qk", "w_v", "q", "k", "v", "attn_out", "w_out", "v1", "w1", "h", "w2", "x_head", "w_head"]


def make_cfg(hilo_names=(), exact_names=(), fold=True):
    cfg = {n: r16 for n in ALL16}
    for n in hilo_names:
        cfg[n] = hilo
    for n in exact_names:
        cfg.pop(n)
    cfg["fold_ln1"] = fold
    return cfg


VARIANTS = {
    "shipped": make_cfg(),
    "qk_hilo": make_cfg(["q", "k"]),
    "q_hilo": make_cfg(["q"]),
    "qk+x_qkv": make_cfg(["q", "k", "x_qkv"]),
    "qk+x_qkv+w_qk": make_cfg(["q", "k", "x_qkv", "w_qk"]),
    "x_qkv+w_qk": make_cfg(["x_qkv", "w_qk"]),
    "all_A_hilo": make_cfg(["q", "k", "x_qkv", "attn_out", "v1", "h", "x_head"]),
    "all_A+qkW": make_cfg(["q", "k", "x_qkv", "attn_out", "v1", "h", "x_head", "w_qk"]),
    "v1_hilo": make_cfg(["v1"]),
    "nofold": make_cfg(fold=False),
    "qk+x+wqk+v1": make_cfg(["q", "k", "x_qkv", "w_qk", "v1"]),
    "qk+x+wqk+v1+head": make_cfg(["q", "k", "x_qkv", "w_qk", "v1", "x_head", "w_head"]),
    "all_hilo": make_cfg(ALL16),
}


def run(sd, prompts, N, T, p, seed, variants):
    g = torch.Generator().manual_seed(seed)
    P, B = prompts.shape
    S = B * T
    seq = prompts.repeat_interleave(T, dim=1)
    L = R.model_dims(sd)["L"]
    worst = {v: 0.0 for v in variants}
    maxd = {v: 0.0 for v in variants}
    sq = {v: 0.0 for v in variants}
    cnt = 0
    for step in range(N):
        n = seq.shape[0]
        keeps = [torch.rand(n, S, 64, generator=g) >= p for _ in range(1 + L)]
        ex = forward(sd, seq, keeps, p, {"fold_ln1": False})
        qx = ex.exp()
        keep = ex > -20
        for v in variants:
            a = forward(sd, seq, keeps, p, VARIANTS[v])
            ratio = (a.exp() - qx).abs() / (1e-2 * qx + 1e-4)
            worst[v] = max(worst[v], ratio.max().item())
            dd = (a - ex)[keep]
            maxd[v] = max(maxd[v], dd.abs().max().item())
            sq[v] += (dd * dd).sum().item()
        cnt += keep.sum().item()
        seq = torch.cat((seq, ex.argmax(-1).view(1, -1)), 0)
    return {v: (math.sqrt(sq[v] / cnt), maxd[v], worst[v]) for v in variants}


if __name__ == "__main__":
    z = H.load("c2_transformer.npz")
    sd = H.state_dict_from(z)
    p = float(z["p"])
    variants = sys.argv[1:] or list(VARIANTS)
    nr = int(os.environ.get("NPROMPTS", "2000"))
    sets = (("fixture prompts", torch.from_numpy(z["prompts"]), int(z["N"])),
            ("random tokens P=4", torch.randint(0, 114, (4, nr), generator=torch.Generator().manual_seed(1)), 5),
            ("random tokens P=12", torch.randint(0, 114, (12, nr), generator=torch.Generator().manual_seed(2)), 5))
    for name, prompts, N in sets:
        for seed in (11, 12):
            res = run(sd, prompts, N, 8, p, seed, variants)
            for v in variants:
                rms, mx, w = res[v]
                print(f"{name:20s} seed {seed} {v:22s}: rms dlogp {rms:.3e}  max |dlogp| {mx:.3e}  worst ratio {w:.3f}", flush=True)

"""CPU emulation (float64 + explicit roundings) of the rounding points of the round-2 fused kernel design, to rank
optional refinements.  See tools/emulate_fused_rounding.py for the method.  usage: python tools/emulate_fused_v2.py"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emulate_fused_rounding as E
import torch
from tests import helpers as H
from oracle import restatement as R

r16, hilo = E.r16, E.hilo


def hilo8(t):
    hi = t.float().half()
    lo = (t.float() - hi.float()).to(torch.float8_e5m2).float()
    return hi.double() + lo.double()


def forward_v2(sd, src, keeps, p, attn_lo=False, w2_lo=False, mlo=False, w1_lo=False, v_lo=False, wv_lo=False, head_lo=False, wout_lo=False, qk8=True, wqk_lo=True, xlo_qk=True, exact_qkv=False, h_lo=True, xlo_v=True):
    D = R.model_dims(sd); d, Hh, dh, L = D["d"], D["H"], D["dh"], D["L"]
    n = src.shape[0]; scale = 1.0 / (1.0 - p)
    x = sd["embedding.weight"].double()[src]
    x = x * keeps[0].double() * scale * math.sqrt(d) + sd["pos_enc.pe"][:n].double()
    causal = torch.triu(torch.full((n, n), float("-inf"), dtype=torch.float64), diagonal=1)
    for i in range(L):
        pre = f"layers.{i}."
        x = hilo(x)
        wq = torch.cat([sd[f"{pre}self_attn.heads.{h}.W_q.weight"] for h in range(Hh)]).double()
        wk = torch.cat([sd[f"{pre}self_attn.heads.{h}.W_k.weight"] for h in range(Hh)]).double()
        wv = torch.cat([sd[f"{pre}self_attn.heads.{h}.W_v.weight"] for h in range(Hh)]).double()
        fq = (lambda t: t) if exact_qkv else (hilo8 if qk8 else r16)
        fw = hilo if wqk_lo else r16
        xq = x if xlo_qk else r16(x)
        q = fq(xq @ fw(wq).t()); k = fq(xq @ fw(wk).t())
        v = ((lambda t: t) if exact_qkv else (hilo if v_lo else r16))((x if xlo_v else r16(x)) @ (hilo if wv_lo else r16)(wv).t())
        S = x.shape[1]
        qh = q.view(n, S, Hh, dh).permute(1, 2, 0, 3); kh = k.view(n, S, Hh, dh).permute(1, 2, 0, 3)
        vh = v.view(n, S, Hh, dh).permute(1, 2, 0, 3)
        sc = qh @ kh.transpose(-1, -2) / math.sqrt(dh) + causal
        o = (torch.softmax(sc, -1) @ vh).permute(2, 0, 1, 3).reshape(n, S, d)
        o = hilo(o) if attn_lo else r16(o)
        Wout = sd[f"{pre}self_attn.W_out.weight"].double()
        v1 = x + o @ (hilo if wout_lo else r16)(Wout).t()
        mean = v1.mean(-1, keepdim=True); var = v1.var(-1, unbiased=False, keepdim=True)
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        g1 = sd[f"{pre}norm1.weight"].double(); b1n = sd[f"{pre}norm1.bias"].double()
        W1 = sd[f"{pre}ff.net.0.weight"].double(); bb1 = sd[f"{pre}ff.net.0.bias"].double()
        w1g16 = r16(W1 * g1)
        w1g = hilo(W1 * g1) if w1_lo else w1g16
        col_s = w1g.sum(-1); col_c = W1 @ b1n + bb1
        M = w1g @ Wout
        M16 = hilo(M) if mlo else r16(M)
        acc = x @ w1g.t() + o @ M16.t()
        hdn = torch.relu(rstd * (acc - mean * col_s) + col_c)
        W2 = sd[f"{pre}ff.net.2.weight"].double()
        ff = (hilo if h_lo else r16)(hdn) @ (hilo if w2_lo else r16)(W2).t() + sd[f"{pre}ff.net.2.bias"].double()
        y = x + ff
        x = torch.nn.functional.layer_norm(y, (d,), sd[f"{pre}norm2.weight"].double(), sd[f"{pre}norm2.bias"].double(), 1e-5)
        x = x * keeps[1 + i].double() * scale
    logits = hilo(x[-1]) @ (hilo if head_lo else r16)(sd["linear_out.weight"].double()).t() + sd["linear_out.bias"].double()
    return torch.log_softmax(logits, -1)


if __name__ == "__main__":
    z = H.load("c2_transformer.npz"); sd = H.state_dict_from(z); p = float(z["p"])
    S = dict(attn_lo=True, w2_lo=True, v_lo=True, wv_lo=True)
    E.VARIANTS = {"shipped": S, "no_qk8": dict(S, qk8=False), "no_wqk_lo": dict(S, wqk_lo=False),
                  "no_qk8_wqk": dict(S, qk8=False, wqk_lo=False), "no_qk8_wqk_xlo": dict(S, qk8=False, wqk_lo=False, xlo_qk=False),
                  "no_w2lo": dict(S, w2_lo=False), "no_xlo_qk": dict(S, xlo_qk=False),
                  "new": dict(S, exact_qkv=True), "new_no_wqk": dict(S, exact_qkv=True, wqk_lo=False), "new_no_wv": dict(S, exact_qkv=True, wv_lo=False),
                  "new_no_wqkv": dict(S, exact_qkv=True, wqk_lo=False, wv_lo=False), "new_no_w2": dict(S, exact_qkv=True, w2_lo=False),
                  "new_no_hlo": dict(S, exact_qkv=True, h_lo=False), "new_no_attnlo": dict(S, exact_qkv=True, attn_lo=False),
                  "new_no_wqkv_w2": dict(S, exact_qkv=True, wqk_lo=False, wv_lo=False, w2_lo=False), "shipped_v2": dict(), "attn_lo": dict(attn_lo=True), "attn+w2": dict(attn_lo=True, w2_lo=True),
                  "attn+w2+m+w1": dict(attn_lo=True, w2_lo=True, mlo=True, w1_lo=True),
                  "attn+w2+m+w1+v+wv": dict(attn_lo=True, w2_lo=True, mlo=True, w1_lo=True, v_lo=True, wv_lo=True),
                  "all": dict(attn_lo=True, w2_lo=True, mlo=True, w1_lo=True, v_lo=True, wv_lo=True, head_lo=True, wout_lo=True)}
    if len(sys.argv) > 1:
        E.VARIANTS = {k: v for k, v in E.VARIANTS.items() if k in sys.argv[1:]}
    E.forward = lambda sd, src, keeps, p, cfg: (forward_v2(sd, src, keeps, p, **cfg) if "fold_ln1" not in cfg else E.__dict__["_orig_forward"](sd, src, keeps, p, cfg))
    names = list(E.VARIANTS)
    fx = torch.from_numpy(z["prompts"])
    for name, prompts in (("fixture", fx), ("P=4", torch.randint(0, 114, (4, 2000), generator=torch.Generator().manual_seed(1))),
                          ("P=12", torch.randint(0, 114, (12, 2000), generator=torch.Generator().manual_seed(2)))):
        for seed in (11, 12, 13):
            res = E.run(sd, prompts, 5, 8, p, seed, names)
            for v in names:
                rms, mx, w = res[v]
                print(f"{name} seed {seed} {v:20s}: rms {rms:.3e} max {mx:.3e} worst ratio {w:.3f}", flush=True)

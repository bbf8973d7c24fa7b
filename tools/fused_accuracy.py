"""Error statistics of the fused fp16-operand kernel against the exact fp32 CUDA path (same Philox masks, the exact
path's tokens forced): RMS / max of the log-prob difference and the worst ratio |dp| / (1e-2 p + 1e-4) over the
trained C2 fixture model, for the fixture's prompts and for uniformly random token prompts.
usage: [BAYESFORMER_B200_LIB=...] python tools/fused_accuracy.py"""
import os, sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from bayesformer_b200 import _lib, engine

_lib.load()
dev = torch.device("cuda:0")
z = H.load("c2_transformer.npz")
sd = H.state_dict_from(z)
model = H.build_model(sd, float(z["p"]), True).to(dev)


def stats(prompts, N, T, seed):
    exact = engine.mc_score_transformer(model.native_pack(dev, "fp32"), prompts, N, T, "dropout", "entropy", seed=seed,
                                        return_logp=True, return_tokens=True)
    forced = exact.tokens[:, prompts.shape[0]:].t().contiguous()
    fused = engine.mc_score_transformer(model.native_pack(dev, "fp16"), prompts, N, T, "dropout", "entropy", seed=seed,
                                        forced_tokens=forced, return_logp=True)
    a, b = fused.logp.double().cpu().numpy(), exact.logp.double().cpu().numpy()
    keep = b > -20
    d = (a - b)[keep]
    p, q = np.exp(a), np.exp(b)
    ratio = np.abs(p - q) / (1e-2 * q + 1e-4)
    if os.environ.get("ROWS"):   # per-row view: is the tail a few bad rows (amplification) or spread out?
        dd = np.where(keep, a - b, 0.0)                      # (steps, sequences, classes)
        row_rms = np.sqrt((dd * dd).mean(axis=2))
        qs = np.quantile(row_rms, [0.5, 0.9, 0.99, 0.999, 1.0])
        print("   per-row rms dlogp quantiles 50/90/99/99.9/100 %:", " ".join(f"{x:.2e}" for x in qs))
        flat = np.argsort(row_rms, axis=None)[::-1][:6]
        for f in flat:
            st, sq = np.unravel_index(f, row_rms.shape)
            top = np.sort(q[st, sq])[::-1][:3]
            ent = -(q[st, sq] * np.log(q[st, sq] + 1e-30)).sum()
            print(f"   worst row: step {st} seq {sq} (item {sq // T}, draw {sq % T}) row rms {row_rms[st, sq]:.2e} "
                  f"max |dlogp| {np.abs(dd[st, sq]).max():.2e} top-3 p {top.round(3)} entropy {ent:.2f}")
    return np.sqrt(np.mean(d * d)), np.abs(d).max(), ratio.max(), np.sqrt(np.mean(ratio * ratio))


print("lib:", os.environ.get("BAYESFORMER_B200_LIB", "default"))
fixture = torch.from_numpy(z["prompts"]).to(dev)
for name, prompts, N in (("fixture prompts", fixture, int(z["N"])),
                         ("random tokens P=4", torch.randint(0, 114, (4, 2000), generator=torch.Generator().manual_seed(1)).to(dev), 5),
                         ("random tokens P=12", torch.randint(0, 114, (12, 2000), generator=torch.Generator().manual_seed(2)).to(dev), 5)):
    for seed in (11, 12):
        rms, mx, worst, rr = stats(prompts, N, 8, seed)
        print(f"{name:20s} seed {seed}: rms dlogp {rms:.3e}  max |dlogp| {mx:.3e}  worst |dp|/(1e-2 p + 1e-4) {worst:.3f}  rms ratio {rr:.4f}")

"""Debug aid: with the library built by `make EXTRA=-DBFB_FUSED_TIMING`, prints the share of a fused-kernel thread's
time spent in product round trips (barrier -> MMA issue -> completion wait)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesformer_b200 import _lib, engine
from bayesformer_b200.model.transformer import MyTransformer
dev = torch.device("cuda:0")
lib = _lib.load()
m = MyTransformer(114, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).eval().to(dev)
pk = engine.scoring_pack(m, dev, 4, 5, "auto")
pr = torch.randint(0, 114, (4, 100000), device=dev)
out = (C.c_ulonglong * 8)()
engine.mc_score_transformer(pk, pr, 5, 20, "dropout", "margin", seed=1)
lib.bfb200_debug_fused_timing(out, 1)
engine.mc_score_transformer(pk, pr, 5, 20, "dropout", "margin", seed=1)
lib.bfb200_debug_fused_timing(out, 1)
rt, n, post, tot, obs, iss, done, fence = (out[i] for i in range(8))
calls = max(1, n / 8)   # eight observer lanes (one per warp) see every round trip of their group
print(f"{obs} observer lanes; {n / obs:.0f} product round trips each; {rt / n:.0f} cycles per round trip "
      f"({post / n:.0f} of them after the group barrier); {rt / tot:.1%} of a thread's time")
print(f"issuing thread: barrier exit -> tcgen05.commit issued {iss / calls:.0f} cycles (tcgen05.fence {fence / calls:.0f}), "
      f"barrier exit -> completion seen {done / calls:.0f} cycles")
ph = (C.c_ulonglong * 12)()
lib.bfb200_debug_fused_phases(ph, 1)
engine.mc_score_transformer(pk, pr, 5, 20, "dropout", "margin", seed=1)
lib.bfb200_debug_fused_phases(ph, 1)
names = ["embedding", "QKV product", "copy Q/K/V to shared", "attention pairs", "attention barriers + output write", "merged product",
         "LN1 stats + FFN hidden", "FFN2 product (+ layer-out draws)", "LN2 + dropout (incl. dead warps' products)", "head product",
         "head epilogue + barrier", "score phase"]
tot_ph = sum(ph[i] for i in range(12))
for i in range(12):
    print(f"  {names[i]:45s} {100.0 * ph[i] / tot_ph:5.1f} %")

"""Multi-tile-per-cluster run of the layered tensor-core path (debug aid): compares fp16 with the fp32 path on the same seed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesformer_b200 import engine
from bayesformer_b200.model.transformer import MyTransformer
dev = torch.device("cuda:0")
torch.manual_seed(0)
d, H, F, L, ln = [int(v) for v in (sys.argv[1:6] or (256, 4, 512, 2, 64))]
B, T = int(sys.argv[6]) if len(sys.argv) > 6 else 64, 10
m = MyTransformer(114, d, H, F, L, bayes_dropout=0.3, bayes=True).eval().to(dev)
pr = torch.randint(0, 114, (ln, B), device=dev)
ex = engine.mc_score_transformer(m.native_pack(dev, "fp32"), pr, 1, T, "dropout", "entropy", seed=5, return_logp=True)
torch.cuda.synchronize(); print("fp32 ok")
got = engine.mc_score_transformer(m.native_pack(dev, "fp16"), pr, 1, T, "dropout", "entropy", seed=5, return_logp=True)
torch.cuda.synchronize(); print("fp16 ok")
e = (got.logp.exp() - ex.logp.exp()).abs().max().item()
rel = ((got.logp.exp() - ex.logp.exp()).abs() / (ex.logp.exp() * 1e-2 + 1e-4)).max().item()
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    engine.mc_score_transformer(m.native_pack(dev, "fp16"), pr, 1, T, "dropout", "entropy", seed=5)
torch.cuda.synchronize()
print("max |dp|", e, " worst error as a fraction of the 1e-2 tolerance", rel, " ms per pass", (time.perf_counter() - t0) / 3 * 1e3)

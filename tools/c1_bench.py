"""SURVEY C1 shape (char-level model: V=67, d=64, 8 heads, F=8, 2 layers; P=32 prompt chars, N=5 decode steps, T=20,
entropy) on synthetic token windows: pool samples/s of the path `auto` picks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesformer_b200 import engine
from bayesformer_b200.model.transformer import MyTransformer
dev = torch.device("cuda:0")
torch.manual_seed(42)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
m = MyTransformer(67, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).eval().to(dev)
pr = torch.randint(0, 65, (32, B), device=dev)
pk = engine.scoring_pack(m, dev, 32, 5, "auto")
print("compute", pk.compute, "path", pk.path)
for _ in range(2):
    engine.mc_score_transformer(pk, pr, 5, 20, "dropout", "entropy", seed=42)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(3):
    engine.mc_score_transformer(pk, pr, 5, 20, "dropout", "entropy", seed=42)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"C1 pool {B}: {ms:.2f} ms per pass, {B / ms * 1e3:.0f} pool samples/s, {B / ms * 1e3 * 268.1e6 / 1e12:.1f} TFLOP/s algorithmic")

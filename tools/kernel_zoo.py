"""Launch every kernel family of the library once or twice at representative shapes (for `ncu -k regex:...`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bayesformer_b200 import engine
from bayesformer_b200.model.transformer import MyTransformer

dev = torch.device("cuda:0")
torch.manual_seed(0)
# scaled configuration, one chunk of 160 sequences x 256 tokens (layered tensor-core path)
big = MyTransformer(114, 512, 8, 2048, 2, bayes_dropout=0.3, bayes=True).eval().to(dev)
prompts = torch.randint(0, 114, (256, 16), device=dev)
pack = engine.scoring_pack(big, dev, 256, 1, "auto")
for _ in range(2):
    engine.mc_score_transformer(pack, prompts, 1, 10, "dropout", "entropy", seed=1)
# acquisition kernels
x = torch.log_softmax(torch.randn(20, 200_000, 114, device=dev), -1)
for _ in range(2):
    engine.acquire(x, is_logp=True, want_mean_p=False)
    engine.acquire(x, is_logp=True, want_mean_p=True)
    engine.compute_uncertainty(torch.softmax(x[0], -1), "margin")
    engine.topk_keys(torch.randn(1_000_000, device=dev), 200)
torch.cuda.synchronize()
print("zoo ok")

import sys, os; sys.path.insert(0, os.getcwd())
import torch
from bayesformer_b200 import engine
dev = torch.device("cuda:0")
for n in (100000, 1000000, 10000000):
    s = torch.rand(n, device=dev)
    for _ in range(5): engine.topk_keys(s, 200)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): engine.topk_keys(s, 200)
    b.record(); torch.cuda.synchronize()
    print(n, "topk_keys us per call:", a.elapsed_time(b) * 1000 / 50)

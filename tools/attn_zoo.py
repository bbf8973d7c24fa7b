"""Launch the tensor-core attention kernel at the scaled configuration's shape (for `ncu -k regex:attention_tc`) and time it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bayesformer_b200 import engine

dev = torch.device("cuda:0")
torch.manual_seed(0)
n_seq, length, heads = 1600, 256, 8
qkv = (torch.randn(n_seq * length, 3 * heads * 64, device=dev) * 0.7).half()
for _ in range(2):
    out = engine.attention_tc(qkv, n_seq, length, heads)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    out = engine.attention_tc(qkv, n_seq, length, heads)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
flops = n_seq * 2 * 256 * 257 * 512
print(f"attention_tc {n_seq}x{length}x{heads}: {ms:.3f} ms per layer, {flops / ms / 1e9:.1f} TFLOP/s causal-algorithmic")

# with a -DBFB_ATTN_TIMING build: average cycles of a 256-key CTA up to each milestone
import ctypes as C
from bayesformer_b200 import _lib
lib = _lib.load()
if hasattr(lib, "bfb200_debug_attn_timing"):
    out = (C.c_ulonglong * 8)()
    lib.bfb200_debug_attn_timing(out, 1)
    engine.attention_tc(qkv, n_seq, length, heads)
    lib.bfb200_debug_attn_timing(out, 1)
    n = max(1, out[6])
    names = ["TMEM + barriers ready", "Q/K landed, S = QK^T done", "row maxima", "weights written", "O = PV done", "output stored"]
    print("256-key CTA, average cycles since CTA start:", ", ".join(f"{nm} {out[i] / n:.0f}" for i, nm in enumerate(names)))

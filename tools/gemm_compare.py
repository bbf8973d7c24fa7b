"""Times torch.matmul (cuBLAS) on the scaled configuration's GEMM shapes next to bfb200_linear_tc — a library
comparator for the numbers in bench.py's aux_kernels (not a product path).  usage: python tools/gemm_compare.py"""
import torch
from bayesformer_b200 import engine

dev = torch.device("cuda:0")


def timed(fn, reps=20):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


R = 65536
for name, n, k in (("qkv", 1536, 512), ("w_out", 512, 512), ("ffn1", 2048, 512), ("ffn2", 512, 2048)):
    a = (torch.randn(R, k, device=dev) * 0.5).half()
    w = (torch.randn(n, k, device=dev) * 0.05).half()
    wt = w.t().contiguous()
    ms_lib = timed(lambda: torch.matmul(a, wt))
    ms_lib2 = timed(lambda: torch.nn.functional.linear(a, w))
    ms_own = timed(lambda: engine.linear_tc(a, w, None, False))
    fl = 2.0 * R * n * k / 1e9
    print(f"{name:6s} {R}x{n}x{k}: cuBLAS matmul {fl/ms_lib:7.1f} TF  F.linear {fl/ms_lib2:7.1f} TF  bfb200_linear_tc {fl/ms_own:7.1f} TF")

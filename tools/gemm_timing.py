"""Debug aid: with the library built by `make EXTRA=-DBFB_GEMM_TIMING`, prints where the roles of the two-SM GEMM
kernel wait (shares of each role's lifetime) at the scaled configuration's shapes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesformer_b200 import _lib, engine
dev = torch.device("cuda:0")
lib = _lib.load()
out = (C.c_ulonglong * 16)()
R = 65536 * 4
for name, n, k in (("qkv", 1536, 512), ("w_out", 512, 512), ("ffn1", 2048, 512), ("ffn2", 512, 2048)):
    a = (torch.randn(R, k, device=dev) * 0.5).half()
    w = (torch.randn(n, k, device=dev) * 0.05).half()
    for _ in range(2):
        engine.linear_tc(a, w, None, False)
    lib.bfb200_debug_gemm_timing(out, 1)
    engine.linear_tc(a, w, None, False)
    lib.bfb200_debug_gemm_timing(out, 1)
    prod_wait, prod_tot, mma_acc, mma_full, mma_tot, epi_wait, epi_busy, epi_tiles, epi_tot = (out[i] for i in range(9))
    print(f"{name:6s}: producer waits for a free slot {prod_wait / prod_tot:.1%} | MMA thread waits: accumulator free "
          f"{mma_acc / mma_tot:.1%}, operands {mma_full / mma_tot:.1%} | epilogue warp: waits for the accumulator "
          f"{epi_wait / epi_tot:.1%}, busy {epi_busy / epi_tot:.1%} ({epi_busy / max(1, epi_tiles):.0f} cycles per tile)")

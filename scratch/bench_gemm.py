"""Stand-alone timings of the gene-sized GEMMs of a KGWAS layer (CUDA events, median of 9)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kgwas_b200 import _lib

def timeit(fn, n=9):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]

dev = torch.device("cuda:0")
M = 20371
shapes = [("NT Z G->S", _lib.KGB_NT, M, 768, 128), ("NT Z G->G", _lib.KGB_NT, M, 640, 128), ("NT root", _lib.KGB_NT, M, 128, 128),
          ("NN dA", _lib.KGB_NN, M, 768, 128), ("NT af", _lib.KGB_NT, M, 128, 768), ("NN dx xf", _lib.KGB_NN, M, 128, 768),
          ("TN dW xf", _lib.KGB_TN, 768, 128, M), ("TN dW af", _lib.KGB_TN, 128, 768, M), ("NT GO", _lib.KGB_NT, 4563, 256, 128),
          ("NT SNP/8 root", _lib.KGB_NT, 98032, 128, 128), ("NT SNP root", _lib.KGB_NT, 784256, 128, 128),
          ("NN SNP dx", _lib.KGB_NN, 784256, 128, 128), ("TN SNP dWr", _lib.KGB_TN, 128, 128, 784256)]
if "--big" in sys.argv:
    shapes = shapes[-4:]
for name, lay, m, n, k in shapes:
    if lay == _lib.KGB_NT:
        a, b = torch.randn(m, k, device=dev), torch.randn(n, k, device=dev)
    elif lay == _lib.KGB_NN:
        a, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev)
    else:
        a, b = torch.randn(k, m, device=dev), torch.randn(k, n, device=dev)
    c = torch.empty(m, n, device=dev)
    t = timeit(lambda: _lib.gemm(lay, a, b, c, m, n, k))
    by = 4 * (a.numel() + b.numel() + c.numel())
    print(f"{name:14s} m={m} n={n} k={k}: {t:7.1f} us  ({by / t / 1e3:.0f} GB/s of operand bytes, {2.0 * m * n * k / t / 1e6:.1f} TFLOP/s fp32-equivalent)")

import sys, torch
sys.path.insert(0, '.')
from kgwas_b200 import _lib
d = torch.device('cuda')
a1 = torch.randn(20371, 768, device=d); b1 = torch.randn(128, 768, device=d); c1 = torch.zeros(20371, 128, device=d)
a2 = torch.randn(784256, 128, device=d); b2 = torch.randn(128, 128, device=d); c2 = torch.zeros(784256, 128, device=d)
cases = [("NT 20371x128x768", _lib.KGB_NT, a1, b1, c1, 20371, 128, 768, 0.0),
         ("NT 784256x128x128 b0", _lib.KGB_NT, a2, b2, c2, 784256, 128, 128, 0.0),
         ("NT 784256x128x128 b1", _lib.KGB_NT, a2, b2, c2, 784256, 128, 128, 1.0),
         ("NN 784256x128x128 b0", _lib.KGB_NN, a2, b2, c2, 784256, 128, 128, 0.0),
         ("TN 128x128x784256", _lib.KGB_TN, a2, c2, torch.zeros(128, 128, device=d), 128, 128, 784256, 0.0)]
for name, lay, a, b, c, m, n, k, beta in cases:
    for _ in range(3): _lib.gemm(lay, a, b, c, m, n, k, beta=beta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): _lib.gemm(lay, a, b, c, m, n, k, beta=beta)
    e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / 10 * 1e3, 1), 'us')

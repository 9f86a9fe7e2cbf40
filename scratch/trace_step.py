"""Timeline of one bench step: every launch that goes through the layer scheduler, per stream, with start / end
relative to the step start (CUDA events; the two extra event records per launch perturb timing slightly)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import kgwas_b200  # noqa: E402
from kgwas_b200 import make_synth_kg, ops  # noqa: E402


def main():
    h, L = 128, 2
    dev = torch.device("cuda:0")
    data = make_synth_kg(1.0, 42, hidden=h)
    n_snp = data["SNP"].x.size(0)
    torch.manual_seed(0)
    model = kgwas_b200.HeteroGNN(data, h, 1, L, "SAGE", "sum", h, h, h, 1).to(dev)
    g = data.to(dev)
    ei = g.edge_index_dict
    x = {k: v.clone().requires_grad_() for k, v in g.x_dict.items()}
    y, w = torch.randn(n_snp, device=dev), torch.rand(n_snp, device=dev, dtype=torch.float64)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=5e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        for v in x.values():
            v.grad = None
        pred = model.forward_from_hidden(x, ei, n_snp).reshape(-1)
        loss = torch.mean(w * (pred - y) ** 2)
        loss.backward()
        opt.step()

    import time
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    c0 = time.perf_counter()
    for _ in range(5):
        step()
    c1 = time.perf_counter()
    torch.cuda.synchronize()
    c2 = time.perf_counter()
    print(f"CPU issue time per step {1e3 * (c1 - c0) / 5:.3f} ms; wall per step incl. drain {1e3 * (c2 - c0) / 5:.3f} ms")
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.TRACE = []
    t0.record()
    step()
    t1.record()
    torch.cuda.synchronize()
    tr, ops.TRACE = ops.TRACE, None
    print(f"step {t0.elapsed_time(t1):.3f} ms, {len(tr)} scheduled launches")
    busy = {}
    for label, st, a, b in tr:
        s, e = t0.elapsed_time(a), t0.elapsed_time(b)
        busy[st] = busy.get(st, 0.0) + e - s
        print(f"{st:5s} {s:8.3f} {e:8.3f} {1e3 * (e - s):8.1f} us  {label}")
    print("sum of durations per stream (ms):", busy)


if __name__ == "__main__":
    main()

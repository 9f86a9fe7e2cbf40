"""The reference's REAL training regime (kgwas/kgwas.py:99-151): NeighborLoader([-1]*L, batch_size=512) mini-batches,
full model (input MLPs on the raw fast-mode widths + L conv layers + head), loss.backward(), Adam -- on kgwas-synth-v1
with the graph resident on the GPU (data_to_cuda=True).  Prints s/it of training and it/s of inference next to the only
numbers the reference publishes (demo/kgwas_101.ipynb: 1.53 s/it training, 3.11 it/s inference, unnamed GPU).

    python scratch/bench_minibatch.py [GAT|SAGE] [iters]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import kgwas_b200  # noqa: E402
from kgwas_b200 import make_synth_kg  # noqa: E402
from kgwas_b200.loader import NeighborLoader  # noqa: E402


def main():
    backbone = sys.argv[1] if len(sys.argv) > 1 else "GAT"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    h, L, bs = 128, 2, 512
    data = make_synth_kg(scale=scale, seed=42, hidden=None)              # raw widths: SNP 20, Gene 5120, GO 128
    n_snp = data["SNP"].num_nodes
    rng = np.random.default_rng(0)
    labelled = rng.permutation(n_snp)[: int(n_snp * 0.69)]               # 542 758 of 784 256 SNPs carry sumstats
    y = torch.full((n_snp,), -1.0)
    y[torch.from_numpy(labelled)] = torch.rand(len(labelled)) * 4.0
    data["SNP"].y = y
    data["SNP"].n_id = torch.arange(n_snp)
    w_table = (0.5 + torch.rand(n_snp, dtype=torch.float64)).to(dev)
    t0 = time.perf_counter()
    gdata = data.to(dev)
    loader = NeighborLoader(gdata, num_neighbors=[-1] * L, input_nodes=("SNP", labelled), batch_size=bs, drop_last=True)
    torch.cuda.synchronize()
    t_loader = time.perf_counter() - t0
    torch.manual_seed(0)
    model = kgwas_b200.HeteroGNN(data, h, 1, L, backbone, "sum", 20, 5120, 128, 1).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=5e-4)     # before the first forward (kgwas.py:116)
    times = {"sample": [], "step": [], "total": []}
    sizes = []
    it = iter(loader)
    for i in range(iters + 2):
        torch.cuda.synchronize()
        ta = time.perf_counter()
        batch = next(it)
        torch.cuda.synchronize()
        tb = time.perf_counter()
        opt.zero_grad()
        b = batch["SNP"].batch_size
        pred = model(batch.x_dict, batch.edge_index_dict, b).reshape(-1)
        yb = batch["SNP"].y[:b]
        wb = w_table[batch["SNP"].n_id[:b]]
        loss = torch.mean(wb * (pred - yb) ** 2)
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        tc = time.perf_counter()
        if i >= 2:
            times["sample"].append(tb - ta)
            times["step"].append(tc - tb)
            times["total"].append(tc - ta)
            sizes.append((sum(int(v.size(0)) for v in batch.x_dict.values()),
                          sum(int(v.size(1)) for v in batch.edge_index_dict.values())))
    # inference: forward only (utils.evaluate_minibatch_clean)
    model.eval()
    inf = []
    with torch.no_grad():
        for i in range(iters):
            torch.cuda.synchronize()
            ta = time.perf_counter()
            batch = next(it)
            b = batch["SNP"].batch_size
            model(batch.x_dict, batch.edge_index_dict, b)
            torch.cuda.synchronize()
            inf.append(time.perf_counter() - ta)
    med = lambda v: float(np.median(v))
    out = {"regime": f"KGWAS.train mini-batches: batch_size {bs} seed SNPs, full {L}-hop neighbourhoods, {backbone} h={h}, "
                     f"kgwas-synth-v1 scale {scale}, raw fast-mode widths through the input MLPs, graph resident on the GPU",
           "train_s_per_it": med(times["total"]), "train_sample_s": med(times["sample"]), "train_fwd_bwd_adam_s": med(times["step"]),
           "inference_it_per_s": 1.0 / med(inf), "batch_nodes_median": int(np.median([s[0] for s in sizes])),
           "batch_edges_median": int(np.median([s[1] for s in sizes])), "loader_setup_s": t_loader, "iters": iters,
           "reference_published": {"train_s_per_it": 1.53, "inference_it_per_s": 3.11,
                                   "source": "demo/kgwas_101.ipynb cell 6 (GAT h=128, batch 512, fast-mode KG, unnamed GPU, "
                                             "CUDA_LAUNCH_BLOCKING=1, host-side sampling); context, not a like-for-like run"},
           "plan_builds": kgwas_b200.plan.plan_builds}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""GPU: SNP-sharded execution reproduces the single-GPU logits and parameter gradients.  With 2 visible GPUs
(gpurun --gpus 2) two ranks run over NCCL; with one GPU the same code path (owned root-row ranges, un-fused ReLU after
the cross-rank sum, gradient all-reduce) runs as a 1-rank NCCL group."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret, h, backbone="SAGE"):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import kgwas_b200
        from kgwas_b200 import dist as kd, make_synth_kg
        data = make_synth_kg(scale=0.004, seed=11, hidden=h)
        n_snp = data["SNP"].num_nodes
        torch.manual_seed(0)
        model = kgwas_b200.HeteroGNN(data, h, 1, 2, backbone, "sum", h, h, h, 1, no_relu=True).to(dev)
        g = torch.Generator().manual_seed(5)
        y, w = torch.randn(n_snp, generator=g).to(dev), torch.rand(n_snp, generator=g, dtype=torch.float64).to(dev)
        # single-GPU reference on this rank
        full = data.to(dev)
        pred_full = model(full.x_dict, full.edge_index_dict, n_snp).reshape(-1)
        (torch.sum(w * (pred_full - y) ** 2) / n_snp).backward()
        ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()
                     if not isinstance(p, torch.nn.parameter.UninitializedParameter) and p.grad is not None}
        model.zero_grad(set_to_none=True)
        kgwas_b200.plan.clear_plan_cache()
        # sharded
        local, shard, (lo, hi) = kd.shard_graph(data, rank, world)
        kd.attach(model, shard)
        loc = local.to(dev)
        pred = model(loc.x_dict, loc.edge_index_dict, hi - lo).reshape(-1)
        (torch.sum(w[lo:hi] * (pred - y[lo:hi]) ** 2) / n_snp).backward()
        params = [p for p in model.parameters() if not isinstance(p, torch.nn.parameter.UninitializedParameter)]
        kd.all_reduce_gradients(params)
        torch.cuda.synchronize()
        err = ((pred - pred_full[lo:hi]).abs().max() / pred_full.abs().max()).item()
        gerr = 0.0
        scale = max(v.abs().max().item() for v in ref_grads.values())
        for k, p in model.named_parameters():
            if isinstance(p, torch.nn.parameter.UninitializedParameter):
                continue
            if k in ref_grads:
                assert p.grad is not None, k
                gerr = max(gerr, (p.grad - ref_grads[k]).abs().max().item() / scale)
        ret[rank] = (err, gerr)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(200)
@pytest.mark.parametrize("h", [64, 128])
def test_sharded_matches_single(cuda, h):
    import torch.multiprocessing as mp
    world = 2 if torch.cuda.device_count() >= 2 else 1
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret, h)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=150)
    for p in procs:
        if p.is_alive():
            p.terminate()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in range(world):
        err, gerr = ret[r]
        assert err < 1e-4 and gerr < 1e-4, (r, err, gerr)

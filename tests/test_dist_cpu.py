"""CPU, world_size 2, gloo: host-side logic of the SNP-sharded multi-GPU path -- graph partition, the
autograd all-reduce, global in-degrees, gradient all-reduce (the CUDA arithmetic is covered by the
2-GPU parity test in test_dist_gpu.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kgwas_b200 import make_synth_kg
        from kgwas_b200 import dist as kd
        data = make_synth_kg(scale=0.002, seed=3, hidden=8)
        local, shard, (lo, hi) = kd.shard_graph(data, rank, world)
        # 1. partition: every edge lives on exactly one rank; SNP indices are relabelled; shared rows split by dst
        ok = True
        for et in data.edge_types:
            n_local = torch.tensor([local[et].edge_index.size(1)])
            dist.all_reduce(n_local)
            ok &= int(n_local) == data[et].edge_index.size(1)
            ei = local[et].edge_index
            if et[0] == "SNP":
                ok &= bool(((ei[0] >= 0) & (ei[0] < hi - lo)).all())
            elif et[2] == "SNP":
                ok &= bool(((ei[1] >= 0) & (ei[1] < hi - lo)).all())
            else:
                d_lo, d_hi = shard.root_range[et[2]]
                ok &= bool(((ei[1] >= d_lo) & (ei[1] < d_hi)).all())
        ok &= torch.equal(local["SNP"].x, data["SNP"].x[lo:hi]) and torch.equal(local["Gene"].x, data["Gene"].x)
        # 2. autograd all-reduce: y = x_0 + x_1 on both ranks, dL/dx_r = sum_r dL/dy_r
        x = torch.full((3,), float(rank + 1), requires_grad=True)
        yv = kd.all_reduce_sum(x)
        (yv * (rank + 1)).sum().backward()
        ok &= torch.equal(yv.detach(), torch.full((3,), 3.0)) and torch.equal(x.grad, torch.full((3,), 3.0))
        # 3. global in-degrees for mean weights: local bincounts all-reduced == unsharded bincount
        et = ("SNP", "TSS", "Gene")
        with shard.building_plan():
            from kgwas_b200 import plan
            deg = plan.DEG_REDUCE(torch.bincount(local[et].edge_index[1], minlength=data["Gene"].num_nodes), "Gene")
            own = torch.arange(3 + rank)
            ok &= plan.DEG_REDUCE(own, "SNP") is own          # owned (sharded) rows: no collective, sizes may differ
        ok &= torch.equal(deg, torch.bincount(data[et].edge_index[1], minlength=data["Gene"].num_nodes))
        ok &= plan.DEG_REDUCE is None
        # 4. flat gradient all-reduce
        p = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 2)), torch.nn.Parameter(torch.zeros(1))]
        p[0].grad = torch.full((4,), float(rank)); p[1].grad = torch.full((2, 2), 2.0 * (rank + 1))
        kd.all_reduce_gradients(p)
        ok &= torch.equal(p[0].grad, torch.full((4,), 1.0)) and torch.equal(p[1].grad, torch.full((2, 2), 6.0)) and p[2].grad is None
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_host_logic_gloo_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(0) is True and ret.get(1) is True


def test_split_range_covers():
    from kgwas_b200.dist import split_range
    for n in (0, 1, 7, 784256):
        for world in (1, 2, 3, 8):
            spans = [split_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


class _Patch:
    """monkeypatch-like setter for the spawned workers (no pytest fixtures there)."""

    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _parity_worker(rank, world, port, ret, backbone, h):
    """Sharded forward + backward on `world` gloo ranks == the single-process result, with the CUDA kernels replaced by
    the CPU stand-ins (tests/_cpu_kernels.py): partition, owned root rows, cross-rank sums (and, for GAT, the cross-rank
    softmax), un-fused ReLU, gradient all-reduce -- everything on the host side of the multi-GPU path."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _cpu_kernels
        import kgwas_b200
        from kgwas_b200 import dist as kd, make_synth_kg
        _cpu_kernels.install(_Patch())
        data = make_synth_kg(scale=0.003, seed=11, hidden=h)
        n_snp = data["SNP"].num_nodes
        torch.manual_seed(0)
        model = kgwas_b200.HeteroGNN(data, h, 1, 2, backbone, "sum", h, h, h, 1, no_relu=True)
        g = torch.Generator().manual_seed(5)
        y, w = torch.randn(n_snp, generator=g), torch.rand(n_snp, generator=g, dtype=torch.float64)
        pred_full = model(data.x_dict, data.edge_index_dict, n_snp).reshape(-1)
        (torch.sum(w * (pred_full - y) ** 2) / n_snp).backward()
        lazy = torch.nn.parameter.UninitializedParameter
        ref = {k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters() if not isinstance(p, lazy)}
        model.zero_grad(set_to_none=True)
        kgwas_b200.plan.clear_plan_cache()
        local, shard, (lo, hi) = kd.shard_graph(data, rank, world)
        kd.attach(model, shard)
        pred = model(local.x_dict, local.edge_index_dict, hi - lo).reshape(-1)
        (torch.sum(w[lo:hi] * (pred - y[lo:hi]) ** 2) / n_snp).backward()
        kd.all_reduce_gradients([p for p in model.parameters() if not isinstance(p, lazy)])
        err = ((pred - pred_full[lo:hi]).abs().max() / pred_full.abs().max()).item()
        scale = max(v.abs().max().item() for v in ref.values() if v is not None)
        gerr, none_ok = 0.0, True
        for k, p in model.named_parameters():
            if isinstance(p, lazy):
                continue
            if ref[k] is None:
                none_ok &= p.grad is None or float(p.grad.abs().max()) == 0.0
            else:
                none_ok &= p.grad is not None
                gerr = max(gerr, (p.grad - ref[k]).abs().max().item() / scale)
        ret[rank] = (err, gerr, bool(none_ok))
    finally:
        dist.destroy_process_group()


def _run_parity(backbone, h):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_parity_worker, args=(r, 2, port, ret, backbone, h)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in range(2):
        err, gerr, none_ok = ret[r]
        assert err < 1e-4 and gerr < 2e-4 and none_ok, (backbone, r, err, gerr, none_ok)


def test_sharded_sage_equals_single_process_gloo_world2():
    _run_parity("SAGE", 32)


def test_sharded_gat_equals_single_process_gloo_world2():
    """incl. the softmax over SNP -> Gene groups whose in-edges are spread over the ranks"""
    _run_parity("GAT", 32)

"""Multi-GPU execution: the SNP-node axis is sharded across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink / NVSwitch).  Rank r owns a contiguous slice of
the SNP rows with their features, labels and every edge that touches one of them; Gene / GO node features and
all parameters are replicated.  Edges and root terms of the shared node types are split by destination row,
so EVERY per-rank quantity of a shared type is a partial sum and one rule makes the result exact:

  forward   partial pre-activation rows of Gene / GO types --all-reduce(sum)--> ReLU          (per layer)
  backward  gradients w.r.t. the replicated Gene / GO inputs --all-reduce(sum)               (per layer)
  step      parameter gradients --all-reduce(sum); the loss is normalised by the GLOBAL number of seeds

Mean-aggregation weights of shared destinations use global in-degrees (all-reduced once at plan time).
Rows of SNP type are owned, never communicated.  ~22 MB per exchange at h=128: latency-bound on NVLink 5.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys
import time
from typing import Dict, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import plan as _plan
from .graph import HeteroData

SHARDED_TYPE = "SNP"


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x (identical on every rank).  Every rank's copy of y feeds that rank's own part of
    the loss, so dL/dx_r = sum over ranks of dL/dy_r: the backward is the same all-reduce."""

    @staticmethod
    def forward(ctx, x):
        y = x.contiguous().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g


def all_reduce_sum(x: torch.Tensor) -> torch.Tensor:
    return _AllReduceSum.apply(x)


def split_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n rows for one rank."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardContext:
    def __init__(self, rank: int, world: int, root_range: Dict[str, Tuple[int, int]]):
        self.rank, self.world, self.root_range = rank, world, root_range
        self.sharded_type = SHARDED_TYPE

    @contextlib.contextmanager
    def building_plan(self):
        def reduce_deg(deg, dst_type):
            if dst_type == SHARDED_TYPE:       # owned rows: every in-edge of a local SNP is local
                return deg
            deg = deg.clone()
            dist.all_reduce(deg, op=dist.ReduceOp.SUM)
            return deg
        prev, _plan.DEG_REDUCE = _plan.DEG_REDUCE, reduce_deg
        try:
            yield
        finally:
            _plan.DEG_REDUCE = prev

    def combine(self, out: Dict[str, torch.Tensor], relu: bool) -> Dict[str, torch.Tensor]:
        """One all-reduce per layer and direction: the partial rows of all shared node types travel as one flat
        buffer (~22 MB at h=128), then the ReLU runs on the summed rows."""
        shared = [t for t in out if t in self.root_range]
        res = dict(out)
        if shared:
            flat = all_reduce_sum(torch.cat([out[t].reshape(-1) for t in shared]))
            if relu:
                flat = flat.relu()
            off = 0
            for t in shared:
                n = out[t].numel()
                res[t] = flat[off:off + n].view_as(out[t])
                off += n
        return res


def shard_graph(data: HeteroData, rank: int, world: int, shard_snp: bool = True):
    """Rank-local view of a KG: SNP rows [lo, hi) (relabelled from 0) + all Gene / GO rows; an edge is kept by the
    rank that owns its SNP endpoint, or -- if it has none -- its destination row.  ``shard_snp=False``: ``data``
    already holds only this rank's SNP block (weak scaling), so only the shared part is split.
    Returns (local data, ShardContext, (snp_lo, snp_hi))."""
    n = {t: data[t].num_nodes for t in data.node_types}
    lo, hi = split_range(n[SHARDED_TYPE], rank, world) if shard_snp else (0, n[SHARDED_TYPE])
    root_range = {t: split_range(n[t], rank, world) for t in data.node_types if t != SHARDED_TYPE}
    local = HeteroData()
    for t in data.node_types:
        for key, val in data[t].items():
            if torch.is_tensor(val) and val.dim() >= 1 and val.size(0) == n[t]:
                local[t][key] = val[lo:hi].clone() if t == SHARDED_TYPE else val
            else:
                local[t][key] = val
    for et in data.edge_types:
        s, _, d = et
        ei = data[et].edge_index
        if s == SHARDED_TYPE:
            keep = (ei[0] >= lo) & (ei[0] < hi)
            sub = ei[:, keep].clone()
            sub[0] -= lo
        elif d == SHARDED_TYPE:
            keep = (ei[1] >= lo) & (ei[1] < hi)
            sub = ei[:, keep].clone()
            sub[1] -= lo
        else:
            d_lo, d_hi = root_range[d]
            keep = (ei[1] >= d_lo) & (ei[1] < d_hi)
            sub = ei[:, keep].clone()
        local[et].edge_index = sub
    return local, ShardContext(rank, world, root_range), (lo, hi)


def attach(model, shard: ShardContext):
    """Make every HeteroConv of a HeteroGNN run SNP-sharded."""
    for conv in model.convs:
        conv.shard = shard
    return model


def all_reduce_gradients(params):
    """Sum the partial parameter gradients of all ranks (one flat all-reduce)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    views = [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)]
    torch._foreach_copy_(grads, views)          # one fused multi-tensor copy instead of ~200 small launches


# -------------------------------------------------------------------------------------------------
# bench.py entry for N > 1 (torchrun, one rank per GPU)
# -------------------------------------------------------------------------------------------------


def bench_sharded(args, rank: int, world: int, dev, helpers):
    """Weak scaling: every rank owns its own block of 784 256 SNPs (and their ~16 M typed SNP<->Gene edges)
    against one shared gene / GO graph whose edges are split by destination row: the global KG has world x 784 256
    variants -- the "millions of variants" regime of north_star.  value = edges of ALL ranks x layers / max-rank time."""
    import kgwas_b200
    from . import _lib
    from .graph import make_synth_kg
    h, L = args.hidden, args.layers
    full = make_synth_kg(scale=args.scale, seed=42, hidden=h, snp_block=rank)
    local, shard, _ = shard_graph(full, rank, world, shard_snp=False)
    sizes = {et: int(ei.size(1)) for et, ei in local.edge_index_dict.items()}
    nodes = {t: int(x.size(0)) for t, x in local.x_dict.items()}
    n_snp = nodes["SNP"]
    g = torch.Generator().manual_seed(43 + rank)
    y = (torch.rand(n_snp, generator=g) * 4.0).to(dev)
    w = 0.5 + torch.rand(n_snp, generator=g, dtype=torch.float64)
    w = (w / w.mean()).to(dev)
    torch.manual_seed(0)
    model = kgwas_b200.HeteroGNN(local, h, 1, L, args.backbone, "sum", h, h, h, 1).to(dev)
    attach(model, shard)
    gdata = local.to(dev)
    ei = gdata.edge_index_dict
    x_dev = {k: v.clone().requires_grad_() for k, v in gdata.x_dict.items()}
    x_host = {k: v.pin_memory() for k, v in local.x_dict.items()}
    n_global = n_snp * world
    opt = None

    def step(x):
        nonlocal opt
        if opt is not None:
            opt.zero_grad(set_to_none=True)
        for v in x.values():
            v.grad = None
        pred = model.forward_from_hidden(x, ei, n_snp).reshape(-1)
        loss = torch.sum(w * (pred - y) ** 2) / n_global          # global mean: sum of the ranks' losses
        loss.backward()
        params = [p for p in model.parameters() if not isinstance(p, torch.nn.parameter.UninitializedParameter)]
        all_reduce_gradients(params)
        if opt is None:
            opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=5e-4, fused=True, capturable=True)
        opt.step()
        return pred, loss

    for _ in range(max(args.warmup, 3)):
        step(x_dev)

    # Same as the single-GPU arm: the step (collectives included -- NCCL operations are legal graph nodes) is captured
    # once and replayed; every rank must agree, so a capture failure anywhere sends all ranks back to eager steps.
    graphed, graph_note = None, "off (--no-cuda-graph)"
    if not getattr(args, "no_cuda_graph", False):
        ok = torch.ones(1, device=dev)
        try:
            from .graphed import GraphedStep
            graphed = GraphedStep(step, x_dev, warmup=3)
            graph_note = "whole step (fwd + bwd + all-reduces + Adam) captured once per rank, replayed per step"
        except Exception as e:                               # noqa: BLE001
            graphed, graph_note = None, f"capture failed, eager steps: {type(e).__name__}: {e}"[:300]
            ok.zero_()
            torch.cuda.synchronize()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0 and graphed is not None:
            graphed, graph_note = None, "capture failed on another rank, eager steps"
    run_step = (lambda x: graphed(x)) if graphed is not None else step
    for _ in range(3):
        run_step(x_dev)

    def timed(fn, steps):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                    # device time, max over ranks
        return float(t.item())

    clocks = helpers["ClockSampler"](dev.index)
    clocks.start()
    k0 = _lib.kernel_launch_count()
    ms = timed(lambda: run_step(x_dev), args.steps)
    launches = _lib.kernel_launch_count() - k0
    if graphed is not None:
        launches = graphed.kernels_per_replay * args.steps
    clk = clocks.stop()

    out_host = torch.empty(n_snp, dtype=torch.float32).pin_memory()

    def e2e_step():
        if graphed is not None:                              # H2D straight into the graph's static inputs, then replay
            with torch.no_grad():
                for k, v in x_host.items():
                    x_dev[k].copy_(v, non_blocking=True)
            pred, loss = graphed()
        else:
            x = {k: v.to(dev, non_blocking=True).requires_grad_() for k, v in x_host.items()}
            pred, loss = step(x)
        out_host.copy_(pred.detach(), non_blocking=True)
        return loss.item()

    e2e_ms = None
    if not args.no_e2e:
        e2e_step()
        e2e_ms = timed(e2e_step, args.steps)

    edges_local = torch.tensor([sum(sizes.values())], dtype=torch.float64, device=dev)
    dist.all_reduce(edges_local)
    edges_layer = float(edges_local.item())
    edges_step = L * edges_layer
    if rank == 0:
        cfg = helpers["workload_config"](args, world)
        cfg["workload"] += f"; weak scaling: {world} SNP blocks of {n_snp} variants, shared gene/GO graph split by destination"
        line = {"metric": "kg_edges_aggregated_per_s_fwd_bwd", "value": edges_step / (ms * 1e-3), "unit": "edges/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "edges_per_step": edges_step, "edges_per_layer_all_ranks": edges_layer,
                "rank0_num_nodes": nodes, "collectives_per_step": "per layer: all-reduce(sum) of the shared node "
                "types' partial rows forward and of their input gradients backward; one flat all-reduce of the "
                "parameter gradients", "clocks": clk, "gpu_launches": launches,
                "gpu_launches_per_step": launches / args.steps, "cuda_graph": graph_note}
        if e2e_ms is not None:
            line["e2e"] = {"value": edges_step / (e2e_ms * 1e-3), "unit": "edges/s", "ms_per_step": e2e_ms,
                           "h2d_bytes_per_step": world * sum(v.numel() * 4 for v in x_host.values()),
                           "d2h_bytes_per_step": world * (n_snp * 4 + 8)}
        print(json.dumps(line), flush=True)
    if graphed is not None:
        import threading
        # belt and braces for the teardown problem described below: whatever blocks after the result is out, the
        # process leaves with status 0 half a minute later
        t = threading.Timer(30.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
    dist.barrier()
    torch.cuda.synchronize()
    if graphed is not None:
        # Measured on 2 x B200 (round 1): after the result line is out, destroy_process_group() blocks for as long as a
        # CUDA graph that captured this communicator's collectives is alive, and the graph cannot be released in a way
        # that is ordered with NCCL's own teardown.  Every rank has passed the barrier and drained its device: leave
        # through a hard exit (exit status 0) instead of tearing the communicator down.
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    dist.destroy_process_group()

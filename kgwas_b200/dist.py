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
from typing import Dict, Tuple

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

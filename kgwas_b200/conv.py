"""Drop-in ``SAGEConv`` / ``GATConv`` / ``HeteroConv`` modules backed by libkgwas_b200.

Constructor arguments, ``forward`` signatures, parameter names and state-dict keys follow the
reference: PyG ``SAGEConv((-1,-1), h)`` and ``HeteroConv(convs, aggr)`` as instantiated at
kgwas/model.py:38,47, and the in-tree ``GATConv`` fork kgwas/conv.py:36-232.  The arithmetic runs
in the CUDA kernels only -- a CPU tensor raises (no fallback).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from ._lib import KGB_NN, KGB_NT, KGB_TN
from .ops import HeteroSageLayerFn, _SageLayerCtx
from .plan import get_plan

EdgeType = Tuple[str, str, str]
MERGE_XF = __import__("os").environ.get("KGB_MERGE_XF", "1") == "1"   # one multi-source gather-reduce per destination type


def glorot_(t: Tensor) -> Tensor:
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-a, a)


class Linear(nn.Module):
    """PyG-style ``Linear`` (weight ``[out, in]``; ``in_channels=-1`` is lazy and materialised in
    place on first use so that an optimiser created earlier -- kgwas/kgwas.py:116 -- keeps the
    same Parameter object).  Used as a parameter holder; the product y = x W^T runs in kgb_gemm."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels, self.weight_initializer = in_channels, out_channels, weight_initializer
        if in_channels > 0:
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        else:
            self.weight = nn.parameter.UninitializedParameter()
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    @property
    def is_lazy(self):
        return isinstance(self.weight, nn.parameter.UninitializedParameter)

    def reset_parameters(self):
        if self.is_lazy:
            return
        if self.weight_initializer == "glorot":
            glorot_(self.weight)
        else:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.weight.size(1)) if self.weight.size(1) > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)

    def materialize(self, in_channels: int):
        if self.is_lazy:
            self.in_channels = in_channels
            self.weight.materialize((self.out_channels, in_channels))
            self.reset_parameters()

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        if self.is_lazy:                      # same as PyG: a lazy weight is stored as-is
            destination[prefix + "weight"] = self.weight
            if self.bias is not None:
                destination[prefix + "bias"] = self.bias if keep_vars else self.bias.detach()
        else:
            super()._save_to_state_dict(destination, prefix, keep_vars)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        key = prefix + "weight"
        w = state_dict.get(key)
        if w is not None and isinstance(w, nn.parameter.UninitializedParameter):
            # saved while still lazy (e.g. GATConv.lin_dst of a same-type relation is never used,
            # kgwas/conv.py:136-138): nothing to copy
            filtered = {k: v for k, v in state_dict.items() if k != key}
            super()._load_from_state_dict(filtered, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                          error_msgs)
            if key in missing_keys:
                missing_keys.remove(key)
            return
        if w is not None and self.is_lazy:
            self.materialize(w.size(1))
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def forward(self, x: Tensor) -> Tensor:
        self.materialize(x.size(-1))
        return _LinearFn.apply(x, self.weight, self.bias)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    @_lib.on_device_of
    def forward(ctx, x, w, b):
        x = x.contiguous()
        out = torch.empty((x.size(0), w.size(0)), dtype=torch.float32, device=x.device)
        _lib.gemm(KGB_NT, x, w, out, x.size(0), w.size(0), w.size(1), bias=b)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return out

    @staticmethod
    @_lib.on_device_of
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.gemm(KGB_NN, g, w, dx, x.size(0), w.size(1), w.size(0))
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            _lib.gemm(KGB_TN, g, x, dw, w.size(0), w.size(1), x.size(0))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(w.size(0), dtype=torch.float32, device=g.device)
            _lib.wcolsum(g, w.size(0), db)
        return dx, dw, db


# -------------------------------------------------------------------------------------------------
# SAGEConv
# -------------------------------------------------------------------------------------------------


class SAGEConv(nn.Module):
    """``SAGEConv(in_channels, out_channels)`` with PyG defaults (aggr='mean', root_weight=True,
    bias=True): ``out = lin_l(mean_{j in N(i)} x_j) + lin_r(x_i)``.  Parameters: ``lin_l.weight``,
    ``lin_l.bias``, ``lin_r.weight``."""

    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int, aggr: str = "mean",
                 normalize: bool = False, root_weight: bool = True, project: bool = False, bias: bool = True):
        super().__init__()
        if aggr != "mean" or normalize or project or not root_weight or not bias:
            raise NotImplementedError("kgwas_b200.SAGEConv implements the configuration KGWAS uses "
                                      "(aggr='mean', root_weight=True, bias=True, normalize=False, project=False)")
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = Linear(in_channels[0], out_channels, bias=True)
        self.lin_r = Linear(in_channels[1], out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def materialize(self, in_src: int, in_dst: int):
        self.lin_l.materialize(in_src)
        self.lin_r.materialize(in_dst)

    def forward(self, x: Union[Tensor, Tuple[Tensor, Tensor]], edge_index: Tensor) -> Tensor:
        same = isinstance(x, Tensor)
        x_src, x_dst = (x, x) if same else x
        self.materialize(x_src.size(-1), x_dst.size(-1))
        et = ("src", "to", "src") if same else ("src", "to", "dst")
        x_dict = {"src": x_src} if same else {"src": x_src, "dst": x_dst}
        out = _hetero_sage({et: self}, x_dict, {et: edge_index}, "sum", False)
        return out[et[2]]

    def __repr__(self):
        return f"SAGEConv({self.in_channels}, {self.out_channels}, aggr=mean)"


def _hetero_sage(convs: Dict[EdgeType, SAGEConv], x_dict, edge_index_dict, aggr: str, relu: bool, shard=None,
                 head=None):
    """Fused multi-relation SAGE layer (all widths equal).  ``shard``: a dist.ShardContext for SNP-sharded runs.
    ``head`` = (node type, weight [1,h]): also return ``out[node type] . weight^T`` ([N,1], computed in the epilogue
    of the layer's last kernel) under the key ``('head', node type)``."""
    node_types = list(x_dict.keys())
    num_nodes = {t: int(x.size(0)) for t, x in x_dict.items()}
    if shard is not None:
        with shard.building_plan():
            plan = get_plan(edge_index_dict, num_nodes, frozenset(convs.keys()), merge_xf=MERGE_XF)
    else:
        plan = get_plan(edge_index_dict, num_nodes, frozenset(convs.keys()), merge_xf=MERGE_XF)
    if not plan.rel_order:
        return {}
    h = convs[plan.rel_order[0]].out_channels
    for t in node_types:
        if x_dict[t].size(-1) != h:
            raise NotImplementedError(f"fused hetero-SAGE needs equal widths: x['{t}'] has {x_dict[t].size(-1)}, "
                                      f"hidden is {h} (KGWAS projects every node type to hidden first, model.py:56-60)")
        if x_dict[t].dtype != torch.float32:
            raise TypeError("kgwas_b200 computes in fp32 (the reference's dtype)")
    rel_scale = {}
    for T in plan.dst_types:
        a, b = plan.rel_range[T]
        rel_scale[T] = 1.0 if aggr == "sum" else 1.0 / (b - a)
    root_range = shard.root_range if shard is not None else None
    if head is not None and not (relu and head[0] in plan.dst_types and head[0] not in (root_range or {})):
        head = None
    meta = _SageLayerCtx(plan, node_types, h, relu, rel_scale, root_range, head[0] if head is not None else None)
    cs = [convs[et] for et in plan.rel_order]
    for c, et in zip(cs, plan.rel_order):
        c.materialize(h, h)
    outs = HeteroSageLayerFn.apply(meta, *[x_dict[t] for t in node_types], *[c.lin_l.weight for c in cs],
                                   *[c.lin_l.bias for c in cs], *[c.lin_r.weight for c in cs],
                                   *([head[1]] if head is not None else []))
    out = dict(zip(plan.dst_types, outs))
    # (sharded runs: the layer Function itself sums the partial rows of the shared node types across ranks and applies
    #  their ReLU -- an all-reduce per shared type on a side stream, overlapping the SNP-row kernels; see ops.py)
    if head is not None:
        out[("head", head[0])] = outs[-1]
    return out


# -------------------------------------------------------------------------------------------------
# HeteroConv
# -------------------------------------------------------------------------------------------------


def _key(edge_type: EdgeType) -> str:
    return "__".join(edge_type)


class HeteroConv(nn.Module):
    """``HeteroConv({edge_type: conv}, aggr)`` (PyG; kgwas/model.py:47).  ``forward(x_dict,
    edge_index_dict, **kwargs_dict)`` returns ``{dst_type: tensor}``; node types that are never a
    destination disappear.  Sub-module keys are ``'__'.join(edge_type)``
    (PyG <= 2.3); ``load_state_dict`` also accepts the PyG >= 2.4 spelling ``<a___b___c>``."""

    def __init__(self, convs: Dict[EdgeType, nn.Module], aggr: Optional[str] = "sum"):
        super().__init__()
        if aggr not in ("sum", "mean", "min", "max", None):
            raise ValueError(f"unknown aggr {aggr!r}")
        self.convs = nn.ModuleDict({_key(k): v for k, v in convs.items()})
        self._edge_types: List[EdgeType] = list(convs.keys())
        self.aggr = aggr
        self.shard = None                 # dist.ShardContext when the SNP axis is sharded over several GPUs
        self._register_load_state_dict_pre_hook(self._rename_pyg24_keys)

    def _rename_pyg24_keys(self, state_dict, prefix, *args):
        for k in list(state_dict.keys()):
            if k.startswith(prefix + "convs.<"):
                rest = k[len(prefix + "convs.<"):]
                name, _, tail = rest.partition(">")
                new = prefix + "convs." + name.replace("___", "__").replace("#", ".") + tail
                state_dict[new] = state_dict.pop(k)

    def conv_dict(self) -> Dict[EdgeType, nn.Module]:
        return {et: self.convs[_key(et)] for et in self._edge_types}

    def forward(self, x_dict, edge_index_dict, *, _fuse_relu: bool = False, _head=None, **kwargs_dict):
        """``_fuse_relu`` / ``_head`` are engine-internal (used by HeteroGNN): apply the ReLU of kgwas/model.py:75 in
        the layer's last kernel; ``_head=(node_type, weight[1,h])`` additionally asks for ``relu(out[node_type]) .
        weight^T`` under the key ``('head', node_type)`` when the fused SAGE path can provide it (absent otherwise)."""
        convs = {et: c for et, c in self.conv_dict().items() if et in edge_index_dict}
        kinds = {type(c) for c in convs.values()}
        widths = {int(x.size(-1)) for x in x_dict.values()}
        fusable = self.aggr in ("sum", "mean") and len(widths) == 1 and not kwargs_dict
        if fusable and kinds == {SAGEConv}:
            return _hetero_sage(convs, x_dict, edge_index_dict, self.aggr, _fuse_relu, self.shard, _head)
        from .gat import GATConv, hetero_gat   # noqa: local import keeps module load light
        if self.aggr in ("sum", "mean") and len(widths) == 1 and kinds == {GATConv}:
            return hetero_gat(convs, x_dict, edge_index_dict, self.aggr, _fuse_relu, kwargs_dict, self.shard)
        if self.shard is not None:
            raise NotImplementedError("SNP-sharded multi-GPU execution needs the fused hetero-SAGE / hetero-GAT layer "
                                      "(sum or mean aggregation, one conv kind, equal widths)")
        return self._per_relation(convs, x_dict, edge_index_dict, _fuse_relu, kwargs_dict)

    def _per_relation(self, convs, x_dict, edge_index_dict, relu, kwargs_dict):
        """Generic path (mixed conv kinds, min / max / None aggregation): one call per relation,
        then PyG's ``group``.  Same kernels, one relation at a time."""
        out_dict: Dict[str, List] = {}
        for et, ei in edge_index_dict.items():
            if et not in convs:
                continue
            src, _, dst = et
            kw = {k[:-5]: v[et] for k, v in kwargs_dict.items() if et in v}
            x = x_dict[src] if src == dst else (x_dict[src], x_dict[dst])
            out_dict.setdefault(dst, []).append(convs[et](x, ei, **kw))
        res = {}
        for k, xs in out_dict.items():
            res[k] = _group(xs, self.aggr)
            if relu and torch.is_tensor(res[k]):
                res[k] = res[k].relu()
        return res

    def __repr__(self):
        return f"HeteroConv(num_relations={len(self.convs)}, aggr={self.aggr})"


def _group(xs: List, aggr: Optional[str]):
    """PyG ``group`` incl. the tuple-aware patch documented at kgwas/utils.py:53-71."""
    if len(xs) == 0:
        return None
    if aggr is None:
        return torch.stack(xs, dim=1)
    if len(xs) == 1:
        return xs[0]
    if isinstance(xs[0], tuple):
        out = torch.stack([i[0] for i in xs], dim=0)
        out = getattr(torch, aggr)(out, dim=0)
        out = out[0] if isinstance(out, tuple) else out
        return (out, [i[1] for i in xs])
    out = torch.stack(xs, dim=0)
    out = getattr(torch, aggr)(out, dim=0)
    return out[0] if isinstance(out, tuple) else out

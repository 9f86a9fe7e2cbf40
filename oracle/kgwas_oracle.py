"""CPU oracle for the KGWAS knowledge-graph convolution path.  TEST INFRASTRUCTURE ONLY.

This file is a pure-PyTorch restatement of the operator sequence the reference executes
for ``HeteroGNN`` (variant / gene / GO hetero-GNN) -- it is the *checker* for the CUDA path
in ``kgwas_b200`` and the timed CPU baseline of ``bench.py``.  Nothing under ``kgwas_b200/``
imports it; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs do.

PARITY STATUS: **partially pinned**.
  * The in-tree half of the reference (``kgwas/model.py``, ``kgwas/conv.py``) is executed
    *verbatim* from /root/reference by ``oracle/gen_golden_from_reference.py`` on top of a
    minimal stand-in for the PyG base classes (``oracle/pyg_standin``) and this oracle must
    reproduce those outputs (``tests/golden/ref_*.pt``: h = 32 models, and h = 128 / 256, L = 2 / 3
    models on a 3 000-SNP KG with hub genes -- ``ref_mid_*.pt``; see tests/test_oracle_golden.py).
  * The third-party half (torch_geometric's ``SAGEConv`` / ``HeteroConv`` / ``MessagePassing``
    / ``softmax`` / ``scatter``) is NOT in /root/reference, is un-pinned there
    (requirements.txt:6, environment.yml:8-10) and cannot be installed here, so those
    semantics are restated from PyG 2.1-2.6 behaviour and anchored
    only by hand-computed micro-cases (tests/test_oracle_microcases.py).  The reference
    ships no tests or golden vectors of its own: "parity unpinned" for that half.

Every function cites the reference file:line (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

EdgeType = Tuple[str, str, str]

# ----------------------------------------------------------------------------------------
# scatter / softmax primitives (PyG torch_geometric.utils.scatter / softmax; Appendix A.2/A.3)
# ----------------------------------------------------------------------------------------


def scatter_sum(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
    """``scatter(src, index, dim=0, dim_size, reduce='sum')`` -- what PyG >= 2.3 lowers to
    without torch_scatter: zeros + index_add_ (used via kgwas/conv.py:182 propagate)."""
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def scatter_mean(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
    """``scatter(..., reduce='mean')``: sum / clamp(count, min=1) (SAGEConv aggr='mean',
    instantiated at kgwas/model.py:38).  Isolated destinations yield 0."""
    total = scatter_sum(src, index, dim_size)
    count = src.new_zeros(dim_size).index_add_(0, index, src.new_ones(index.numel()))
    count = count.clamp(min=1)
    return total / count.view((-1,) + (1,) * (src.dim() - 1))


def pyg_softmax(src: Tensor, index: Tensor, num_nodes: int) -> Tensor:
    """``torch_geometric.utils.softmax(src, index, num_nodes=N)`` (called at
    kgwas/conv.py:223): exp(src - max_per_group) / (sum_per_group + 1e-16), max detached."""
    if src.numel() == 0:
        return src.clone()
    shape = (num_nodes,) + tuple(src.shape[1:])
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    src_max = src.new_full(shape, float("-inf")).scatter_reduce_(
        0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = scatter_sum(out, index, num_nodes) + 1e-16
    return out / out_sum.index_select(0, index)


# ----------------------------------------------------------------------------------------
# parameter initialisers (torch_geometric.nn.inits / dense.linear; Appendix A.3, A.4)
# ----------------------------------------------------------------------------------------


def glorot_(t: Tensor) -> Tensor:
    """PyG ``glorot``: U(-a, a), a = sqrt(6 / (size(-2) + size(-1))) (kgwas/conv.py:117-119)."""
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-a, a)


class PygLinear(nn.Module):
    """``torch_geometric.nn.Linear`` (kgwas/model.py:50, kgwas/conv.py:81-89): y = x W^T + b,
    weight [out, in]; ``in_channels = -1`` is lazy and materialised in place at the first
    forward (same Parameter object, so an optimiser built earlier keeps working --
    kgwas/kgwas.py:116 builds Adam before the first forward)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer = weight_initializer
        if in_channels > 0:
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        else:
            self.weight = nn.parameter.UninitializedParameter()
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        if isinstance(self.weight, nn.parameter.UninitializedParameter):   # PyG stores a lazy weight as-is
            destination[prefix + "weight"] = self.weight
            if self.bias is not None:
                destination[prefix + "bias"] = self.bias.detach()
        else:
            super()._save_to_state_dict(destination, prefix, keep_vars)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing, unexpected, errors):
        w = state_dict.get(prefix + "weight")
        lazy_in = isinstance(w, nn.parameter.UninitializedParameter)
        if w is not None and not lazy_in and isinstance(self.weight, nn.parameter.UninitializedParameter):
            self.in_channels = w.size(1)
            self.weight.materialize((self.out_channels, self.in_channels))
        if lazy_in:
            state_dict = {k: v for k, v in state_dict.items() if k != prefix + "weight"}
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing, unexpected, errors)
        if lazy_in and prefix + "weight" in missing:
            missing.remove(prefix + "weight")

    def reset_parameters(self):
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            return
        if self.weight_initializer == "glorot":
            glorot_(self.weight)
        else:  # kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.weight.size(1)) if self.weight.size(1) > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            self.in_channels = x.size(-1)
            self.weight.materialize((self.out_channels, self.in_channels))
            self.reset_parameters()
        return F.linear(x, self.weight, self.bias)


# ----------------------------------------------------------------------------------------
# SAGEConv (PyG; instantiated kgwas/model.py:38) -- Appendix A.2
# ----------------------------------------------------------------------------------------


class SAGEConv(nn.Module):
    """``SAGEConv((-1,-1), h)`` defaults: aggr='mean', root_weight=True, bias=True,
    normalize=False, project=False.  out = lin_l(mean_j x_j) + lin_r(x_dst)."""

    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = PygLinear(in_channels[0], out_channels, bias=True)
        self.lin_r = PygLinear(in_channels[1], out_channels, bias=False)

    def forward(self, x: Union[Tensor, Tuple[Tensor, Tensor]], edge_index: Tensor) -> Tensor:
        if isinstance(x, Tensor):
            x = (x, x)
        x_src, x_dst = x
        msg = x_src.index_select(0, edge_index[0])                 # gather  [E, h]
        agg = scatter_mean(msg, edge_index[1], x_dst.size(0))      # reduce  [N_dst, h]
        return self.lin_l(agg) + self.lin_r(x_dst)


# ----------------------------------------------------------------------------------------
# GATConv -- restatement of the in-tree fork kgwas/conv.py:36-232, as configured by
# kgwas/model.py:40-42 (add_self_loops=False, edge_dim=None, concat=True, dropout=0)
# ----------------------------------------------------------------------------------------


class GATConv(nn.Module):
    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int,
                 heads: int = 1, concat: bool = True, negative_slope: float = 0.2,
                 dropout: float = 0.0, add_self_loops: bool = True, bias: bool = True,
                 sigmoid_gat: bool = False, temperature: float = 1.0):
        super().__init__()
        assert not add_self_loops, "oracle covers the configuration KGWAS uses (model.py:42)"
        assert concat and dropout == 0.0
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope, self.sigmoid_gat, self.temperature = negative_slope, sigmoid_gat, temperature
        if isinstance(in_channels, int):                                   # conv.py:81-84
            self.lin_src = PygLinear(in_channels, heads * out_channels, False, "glorot")
            self.lin_dst = self.lin_src
        else:                                                              # conv.py:85-89
            self.lin_src = PygLinear(in_channels[0], heads * out_channels, False, "glorot")
            self.lin_dst = PygLinear(in_channels[1], heads * out_channels, False, "glorot")
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))   # conv.py:92-93
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels)) if bias else None
        glorot_(self.att_src); glorot_(self.att_dst)                       # conv.py:117-120
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index: Tensor, return_attention_weights=None,
                return_raw_attention_weights=None):
        H, C = self.heads, self.out_channels
        raw = bool(return_raw_attention_weights)                           # conv.py:127-130
        if isinstance(x, Tensor):                                          # conv.py:136-138
            x_src = x_dst = self.lin_src(x).view(-1, H, C)
        else:                                                              # conv.py:139-144
            x_src, x_dst = x
            x_src = self.lin_src(x_src).view(-1, H, C)
            if x_dst is not None:
                x_dst = self.lin_dst(x_dst).view(-1, H, C)
        alpha_src = (x_src * self.att_src).sum(dim=-1)                     # conv.py:150
        alpha_dst = None if x_dst is None else (x_dst * self.att_dst).sum(-1)  # conv.py:151
        src, dst = edge_index[0], edge_index[1]
        n_dst = x_dst.size(0) if x_dst is not None else x_src.size(0)
        # edge_update (conv.py:200-225)
        alpha = alpha_src.index_select(0, src)
        if alpha_dst is not None:
            alpha = alpha + alpha_dst.index_select(0, dst)                 # conv.py:205
        alpha = F.leaky_relu(alpha, self.negative_slope)                   # conv.py:217
        if self.sigmoid_gat:
            alpha = torch.sigmoid(alpha / self.temperature)                # conv.py:220
        elif not raw:
            alpha = pyg_softmax(alpha / self.temperature, dst, n_dst)      # conv.py:223
        # message + 'add' aggregation (conv.py:227-228, :54)
        msg = alpha.unsqueeze(-1) * x_src.index_select(0, src)
        out = scatter_sum(msg, dst, n_dst).view(-1, H * C)                 # conv.py:185
        if self.bias is not None:
            out = out + self.bias                                          # conv.py:190
        if isinstance(return_attention_weights, bool):                     # conv.py:192-194
            return out, (edge_index, alpha)
        return out


# ----------------------------------------------------------------------------------------
# HeteroConv (PyG; in-tree echoes kgwas/conv.py:17-32, kgwas/utils.py:53-71) -- Appendix A.1
# ----------------------------------------------------------------------------------------


def group(xs: List, aggr: Optional[str]):
    """PyG ``group`` with the patch KGWAS documents at kgwas/utils.py:53-71 (tuple outputs when
    attention weights are requested).  NB: a destination type fed by a single relation gets
    ``xs[0]`` back unchanged (a bare tuple, not (out, [att])) -- kept as in the reference."""
    if len(xs) == 0:
        return None
    elif aggr is None:
        return torch.stack(xs, dim=1)
    elif len(xs) == 1:
        return xs[0]
    elif isinstance(xs, list) and isinstance(xs[0], tuple):
        out = torch.stack([i[0] for i in xs], dim=0)
        out = getattr(torch, aggr)(out, dim=0)
        out = out[0] if isinstance(out, tuple) else out
        return (out, [i[1] for i in xs])
    else:
        out = torch.stack(xs, dim=0)
        out = getattr(torch, aggr)(out, dim=0)
        return out[0] if isinstance(out, tuple) else out


class HeteroConv(nn.Module):
    def __init__(self, convs: Dict[EdgeType, nn.Module], aggr: Optional[str] = "sum"):
        super().__init__()
        self.convs = nn.ModuleDict({"__".join(k): v for k, v in convs.items()})
        self.aggr = aggr

    def forward(self, x_dict, edge_index_dict, **kwargs_dict):
        out_dict: Dict[str, List] = {}
        for edge_type, edge_index in edge_index_dict.items():
            src, _, dst = edge_type
            key = "__".join(edge_type)
            if key not in self.convs:
                continue
            kwargs = {}
            for arg, value_dict in kwargs_dict.items():
                arg = arg[:-5]                                 # strip '_dict'
                if edge_type in value_dict:
                    kwargs[arg] = value_dict[edge_type]
            conv = self.convs[key]
            if src == dst:
                out = conv(x_dict[src], edge_index, **kwargs)
            else:
                out = conv((x_dict[src], x_dict[dst]), edge_index, **kwargs)
            out_dict.setdefault(dst, []).append(out)
        for k, v in out_dict.items():
            out_dict[k] = group(v, self.aggr)
        return out_dict


# ----------------------------------------------------------------------------------------
# HeteroGNN -- restatement of kgwas/model.py:10-86
# ----------------------------------------------------------------------------------------


class SimpleMLP(nn.Module):                                    # model.py:10-22
    def __init__(self, input_dim, hidden_dim, output_dim):
        super().__init__()
        self.FC_hidden = nn.Linear(input_dim, hidden_dim)
        self.FC_hidden2 = nn.Linear(hidden_dim, hidden_dim)
        self.FC_output = nn.Linear(hidden_dim, output_dim)
        self.ReLU = nn.ReLU()

    def forward(self, x):
        h = self.ReLU(self.FC_hidden(x))
        h = self.ReLU(self.FC_hidden2(h))
        return self.FC_output(h)


class HeteroGNN(nn.Module):                                    # model.py:24-86
    def __init__(self, pyg_data, hidden_channels, out_channels, num_layers, gnn_backbone,
                 gnn_aggr, snp_init_dim_size, gene_init_dim_size, go_init_dim_size,
                 gat_num_head, no_relu=False, lazy=True):
        super().__init__()
        edge_types = pyg_data.edge_types
        self.convs = nn.ModuleList()
        self.snp_feat_mlp = SimpleMLP(snp_init_dim_size, hidden_channels, hidden_channels)
        self.go_feat_mlp = SimpleMLP(go_init_dim_size, hidden_channels, hidden_channels)
        self.gene_feat_mlp = SimpleMLP(gene_init_dim_size, hidden_channels, hidden_channels)
        self.ReLU = nn.ReLU()
        in_ch = (-1, -1) if lazy else (hidden_channels, hidden_channels)
        for _ in range(num_layers):
            conv_layer = {}
            for i in edge_types:
                if gnn_backbone == "SAGE":
                    conv_layer[i] = SAGEConv(in_ch, hidden_channels)
                elif gnn_backbone == "GAT":
                    conv_layer[i] = GATConv(in_ch, hidden_channels, heads=gat_num_head,
                                            add_self_loops=False)
                else:
                    raise NotImplementedError(gnn_backbone)
            self.convs.append(HeteroConv(conv_layer, aggr=gnn_aggr))
        self.lin = PygLinear(hidden_channels, out_channels)
        self.no_relu = no_relu

    def forward(self, x_dict, edge_index_dict, batch_size, genotype=None, return_h=False,
                return_attention_weights=False):
        x_dict = dict(x_dict)   # the reference mutates the caller's dict (model.py:56); values equal
        x_dict["SNP"] = self.snp_feat_mlp(x_dict["SNP"])
        x_dict["Gene"] = self.gene_feat_mlp(x_dict["Gene"])
        x_dict["CellularComponent"] = self.go_feat_mlp(x_dict["CellularComponent"])
        x_dict["BiologicalProcess"] = self.go_feat_mlp(x_dict["BiologicalProcess"])
        x_dict["MolecularFunction"] = self.go_feat_mlp(x_dict["MolecularFunction"])
        attention_all_layers = []
        for conv in self.convs:
            if return_attention_weights:                       # model.py:65-72
                keys = list(edge_index_dict.keys())
                out = conv(x_dict, edge_index_dict,
                           return_attention_weights_dict=dict(zip(keys, [True] * len(keys))))
                mean_attention = torch.mean(torch.vstack(
                    [torch.vstack([x[1] for x in j[1]]) for i, j in out.items()]))
                x_dict = {i: j[0] for i, j in out.items()}
                attention_all_layers.append(mean_attention)
            else:
                x_dict = conv(x_dict, edge_index_dict)         # model.py:74
            x_dict = {key: x.relu() for key, x in x_dict.items()}   # model.py:75
        if return_h:
            return self.ReLU(self.lin(x_dict["SNP"]))[:batch_size], x_dict["SNP"][:batch_size]
        if return_attention_weights:
            return self.ReLU(self.lin(x_dict["SNP"]))[:batch_size], attention_all_layers
        if self.no_relu:
            return self.lin(x_dict["SNP"])[:batch_size]
        return self.ReLU(self.lin(x_dict["SNP"]))[:batch_size]


def conv_stack_forward(convs, x_dict, edge_index_dict):
    """The L x (HeteroConv -> ReLU) core of model.py:64-75 on already-projected features --
    exactly the region the throughput metric (edges aggregated / s) is defined on."""
    for conv in convs:
        x_dict = conv(x_dict, edge_index_dict)
        x_dict = {k: v.relu() for k, v in x_dict.items()}
    return x_dict


def weighted_mse(pred: Tensor, y: Tensor, w: Tensor) -> Tensor:
    """kgwas/kgwas.py:145 -- ``torch.mean(ld_weight * (pred - y_batch)**2)`` (w is float64)."""
    return torch.mean(w * (pred - y) ** 2)

from typing import Dict, List, Optional

import torch
from torch import nn

from .conv.message_passing import MessagePassing
from .dense.linear import Linear  # noqa: F401


class SAGEConv(MessagePassing):
    """PyG SAGEConv defaults: mean aggregation, root weight, bias on lin_l only."""

    def __init__(self, in_channels, out_channels, aggr="mean", **kwargs):
        super().__init__(aggr=aggr)
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.lin_l = Linear(in_channels[0], out_channels, bias=True)
        self.lin_r = Linear(in_channels[1], out_channels, bias=False)

    def message(self, x_j):
        return x_j

    def forward(self, x, edge_index, size=None):
        if torch.is_tensor(x):
            x = (x, x)
        out = self.lin_l(self.propagate(edge_index, x=x, size=size))
        if x[1] is not None:
            out = out + self.lin_r(x[1])
        return out


class _Unused(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not used by the KGWAS configurations under test")


GCNConv = SGConv = Sequential = _Unused


def to_hetero(*a, **k):
    raise NotImplementedError


def group(xs: List, aggr: Optional[str]):
    """PyG's group() with the patch KGWAS asks its users to apply (kgwas/utils.py:53-71)."""
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
    def __init__(self, convs: Dict, aggr: Optional[str] = "sum"):
        super().__init__()
        self.convs = nn.ModuleDict({"__".join(k): v for k, v in convs.items()})
        self.aggr = aggr

    def forward(self, x_dict, edge_index_dict, *args_dict, **kwargs_dict):
        out_dict = {}
        for edge_type, edge_index in edge_index_dict.items():
            src, rel, dst = edge_type
            str_edge_type = "__".join(edge_type)
            if str_edge_type not in self.convs:
                continue
            kwargs = {arg[:-5]: value_dict[edge_type] for arg, value_dict in kwargs_dict.items() if edge_type in value_dict}
            conv = self.convs[str_edge_type]
            if src == dst:
                out = conv(x_dict[src], edge_index, **kwargs)
            else:
                out = conv((x_dict[src], x_dict[dst]), edge_index, **kwargs)
            out_dict.setdefault(dst, []).append(out)
        for key, value in out_dict.items():
            out_dict[key] = group(value, self.aggr)
        return out_dict

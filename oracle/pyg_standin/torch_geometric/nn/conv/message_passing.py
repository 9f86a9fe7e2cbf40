import inspect

import torch

from ...utils import scatter


class MessagePassing(torch.nn.Module):
    """The slice of PyG's MessagePassing that kgwas/conv.py drives: ``propagate`` (message -> aggregate) and
    ``edge_updater`` (edge_update), both with PyG's ``<name>_i`` / ``<name>_j`` argument collection
    (flow source_to_target: j = edge_index[0], i = edge_index[1])."""

    def __init__(self, aggr="add", node_dim=0, **kwargs):
        super().__init__()
        self.aggr, self.node_dim = aggr, node_dim

    def _collect(self, fn, edge_index, size, kwargs):
        src, dst = edge_index[0], edge_index[1]
        out = {}
        for name in inspect.signature(fn).parameters:
            if name.endswith("_i") or name.endswith("_j"):
                data = kwargs.get(name[:-2])
                idx, side = (dst, 1) if name.endswith("_i") else (src, 0)
                if isinstance(data, (tuple, list)):
                    data = data[side]
                out[name] = None if data is None else data.index_select(self.node_dim, idx)
            elif name == "index":
                out[name] = dst
            elif name == "ptr":
                out[name] = None
            elif name == "size_i":
                out[name] = size[1]
            elif name == "size_j":
                out[name] = size[0]
            elif name in kwargs:
                out[name] = kwargs[name]
        return out

    @staticmethod
    def _sizes(size, kwargs):
        size = [None, None] if size is None else list(size)
        for v in kwargs.values():
            if isinstance(v, (tuple, list)):
                for side in (0, 1):
                    if size[side] is None and torch.is_tensor(v[side]):
                        size[side] = v[side].size(0)
            elif torch.is_tensor(v) and v.dim() > 1:
                size = [s if s is not None else v.size(0) for s in size]
        if size[0] is None:
            size[0] = size[1]
        if size[1] is None:
            size[1] = size[0]
        return size

    def propagate(self, edge_index, size=None, **kwargs):
        size = self._sizes(size, {k: v for k, v in kwargs.items() if k == "x"})
        msg = self.message(**self._collect(self.message, edge_index, size, kwargs))
        out = scatter(msg, edge_index[1], size[1], reduce="sum" if self.aggr == "add" else self.aggr)
        return self.update(out)

    def edge_updater(self, edge_index, size=None, **kwargs):
        size = self._sizes(size, {k: v for k, v in kwargs.items() if isinstance(v, (tuple, list))})
        return self.edge_update(**self._collect(self.edge_update, edge_index, size, kwargs))

    def update(self, inputs):
        return inputs

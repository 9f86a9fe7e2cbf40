import torch


def scatter(src, index, dim_size, reduce="sum"):
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    if reduce in ("sum", "add"):
        return out.index_add_(0, index, src)
    if reduce == "mean":
        out.index_add_(0, index, src)
        cnt = src.new_zeros(dim_size).index_add_(0, index, torch.ones_like(index, dtype=src.dtype)).clamp_(min=1)
        return out / cnt.view(-1, *([1] * (src.dim() - 1)))
    if reduce == "max":
        out = src.new_full((dim_size,) + tuple(src.shape[1:]), float("-inf"))
        idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
        return out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    raise ValueError(reduce)


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    src_max = scatter(src.detach(), index, n, reduce="max")
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = scatter(out, index, n, reduce="sum") + 1e-16
    return out / out_sum.index_select(0, index)


def remove_self_loops(edge_index, edge_attr=None):
    raise NotImplementedError("KGWAS instantiates GATConv with add_self_loops=False")


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    raise NotImplementedError("KGWAS instantiates GATConv with add_self_loops=False")

"""Drop-in ``HeteroGNN`` (kgwas/model.py:24-86): same constructor, ``forward`` signature, attribute
names (``snp_feat_mlp`` / ``gene_feat_mlp`` / ``go_feat_mlp`` / ``convs`` / ``lin``) and state-dict
keys; the L x (HeteroConv -> ReLU) core runs in the fused CUDA engine."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .conv import HeteroConv, Linear, SAGEConv


class SimpleMLP(nn.Module):
    """kgwas/model.py:10-22.  Dense GEMMs on the raw features: kept on torch.nn.Linear (cuBLAS) --
    adjacent to the hot path, SURVEY.md section 8 f-1."""

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


class _HeadFn(torch.autograd.Function):
    """logit[i] = <h[i,:], w> + b  for the single-output head."""

    @staticmethod
    @_lib.on_device_of
    def forward(ctx, hid, w, b):
        hid = hid.contiguous()
        out = torch.empty((hid.size(0), 1), dtype=torch.float32, device=hid.device)
        _lib.rowdot(hid, w, out, w.size(1), 1, 0)
        if b is not None:
            out += b
        ctx.save_for_backward(hid, w)
        ctx.has_bias = b is not None
        return out

    @staticmethod
    @_lib.on_device_of
    def backward(ctx, g):
        hid, w = ctx.saved_tensors
        g = g.contiguous()
        hdim = w.size(1)
        d_hid = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            d_hid = torch.empty_like(hid)
            _lib.rank_update(g, w, d_hid, hdim, 1, 0.0)
        if ctx.needs_input_grad[1]:
            d_w = torch.empty_like(w)
            _lib.wcolsum(hid, hdim, d_w, w=g, n_slots=1)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_b = g.sum(0)
        return d_hid, d_w, d_b


def _head_is_fusable(w_lin) -> bool:
    """single-output fp32 head on the GPU: evaluated in the epilogue of the last layer's last SNP-row kernel"""
    return w_lin.size(0) == 1 and w_lin.is_cuda and w_lin.dtype == torch.float32


class HeteroGNN(nn.Module):
    def __init__(self, pyg_data, hidden_channels, out_channels, num_layers, gnn_backbone, gnn_aggr,
                 snp_init_dim_size, gene_init_dim_size, go_init_dim_size, gat_num_head, no_relu=False):
        super().__init__()
        edge_types = pyg_data.edge_types                                   # model.py:27
        self.convs = nn.ModuleList()
        self.snp_feat_mlp = SimpleMLP(snp_init_dim_size, hidden_channels, hidden_channels)
        self.go_feat_mlp = SimpleMLP(go_init_dim_size, hidden_channels, hidden_channels)
        self.gene_feat_mlp = SimpleMLP(gene_init_dim_size, hidden_channels, hidden_channels)
        self.ReLU = nn.ReLU()
        for _ in range(num_layers):
            conv_layer = {}
            for i in edge_types:
                if gnn_backbone == "SAGE":
                    conv_layer[i] = SAGEConv((-1, -1), hidden_channels)
                elif gnn_backbone == "GAT":
                    from .gat import GATConv
                    conv_layer[i] = GATConv((-1, -1), hidden_channels, heads=gat_num_head, add_self_loops=False)
                else:
                    # model.py:43-46 offers GCN / SGC, but PyG's GCNConv / SGConv are not defined for the
                    # bipartite (x_src, x_dst) inputs HeteroConv passes on this KG
                    raise NotImplementedError(f"gnn_backbone={gnn_backbone!r}: only 'SAGE' and 'GAT' are supported")
            self.convs.append(HeteroConv(conv_layer, aggr=gnn_aggr))
        self.lin = Linear(hidden_channels, out_channels)                   # model.py:50 (PyG Linear)
        self.no_relu = no_relu

    def encode(self, x_dict):
        """Input MLPs (model.py:56-60); the GO MLP is shared by the three GO node types."""
        out = dict(x_dict)
        out["SNP"] = self.snp_feat_mlp(x_dict["SNP"])
        out["Gene"] = self.gene_feat_mlp(x_dict["Gene"])
        out["CellularComponent"] = self.go_feat_mlp(x_dict["CellularComponent"])
        out["BiologicalProcess"] = self.go_feat_mlp(x_dict["BiologicalProcess"])
        out["MolecularFunction"] = self.go_feat_mlp(x_dict["MolecularFunction"])
        return out

    def head(self, h_snp):
        """``self.lin`` (model.py:50,83-86).  KGWAS uses out_channels = 1: a row-dot per SNP, run by kgb_rowdot /
        kgb_rank_update / kgb_wcolsum; other widths go through torch (plain library GEMM)."""
        if self.lin.weight.size(0) == 1 and h_snp.is_cuda and h_snp.dtype == torch.float32:
            return _HeadFn.apply(h_snp, self.lin.weight, self.lin.bias)
        return torch.nn.functional.linear(h_snp, self.lin.weight, self.lin.bias)

    def forward(self, x_dict, edge_index_dict, batch_size, genotype=None, return_h=False,
                return_attention_weights=False):
        return self.forward_from_hidden(self.encode(x_dict), edge_index_dict, batch_size, return_h,
                                        return_attention_weights)

    def forward_from_hidden(self, x_dict, edge_index_dict, batch_size, return_h=False,
                            return_attention_weights=False):
        """model.py:62-86 on already-projected ``[N, hidden]`` features: L x (HeteroConv -> ReLU),
        head, slice.  This is the region the edges-aggregated/s metric is defined on."""
        attention_all_layers = []
        w_lin = self.lin.weight
        fuse_head = not return_attention_weights and _head_is_fusable(w_lin)
        logits_all = None
        for li, conv in enumerate(self.convs):
            if return_attention_weights:                                   # model.py:65-72
                keys = list(edge_index_dict.keys())
                out = conv(x_dict, edge_index_dict,
                           return_attention_weights_dict=dict(zip(keys, [True] * len(keys))))
                mean_attention = torch.mean(torch.vstack(
                    [torch.vstack([x[1] for x in j[1]]) for i, j in out.items()]))
                x_dict = {i: j[0].relu() for i, j in out.items()}
                attention_all_layers.append(mean_attention)
            else:
                # model.py:74-75 (ReLU fused); the last layer also evaluates the single-output head in the epilogue
                # of the kernel that finishes the SNP rows
                head = ("SNP", w_lin) if fuse_head and li == len(self.convs) - 1 else None
                x_dict = conv(x_dict, edge_index_dict, _fuse_relu=True, _head=head)
                logits_all = x_dict.pop(("head", "SNP"), None)
        h_snp = x_dict["SNP"]
        if batch_size < h_snp.size(0):         # rows are independent: slicing before the head is exact
            h_snp = h_snp[:batch_size]
        if logits_all is not None:
            out = logits_all[:batch_size] if batch_size < logits_all.size(0) else logits_all
            if self.lin.bias is not None:
                out = out + self.lin.bias
        else:
            out = self.head(h_snp)
        if return_h:
            return self.ReLU(out), h_snp
        if return_attention_weights:
            return self.ReLU(out), attention_all_layers
        if self.no_relu:
            return out
        return self.ReLU(out)

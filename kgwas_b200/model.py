"""Drop-in ``HeteroGNN`` (kgwas/model.py:24-86): same constructor, ``forward`` signature, attribute
names (``snp_feat_mlp`` / ``gene_feat_mlp`` / ``go_feat_mlp`` / ``convs`` / ``lin``) and state-dict
keys; the L x (HeteroConv -> ReLU) core runs in the fused CUDA engine."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import KGB_NN, KGB_NT, KGB_TN
from .conv import HeteroConv, Linear, SAGEConv


class _MlpFn(torch.autograd.Function):
    """SimpleMLP (kgwas/model.py:10-22) on the engine's kernels: three kgb_gemm calls with the bias and the ReLU in the
    GEMM epilogue; backward = input-gradient / weight-gradient GEMMs, the ReLU mask and the bias column sums fused in one
    pass per hidden layer (kgb_relu_bwd_fused)."""

    @staticmethod
    @_lib.on_device_of
    def forward(ctx, x, w1, b1, w2, b2, w3, b3):
        x = x.contiguous()
        n, hid, out_dim = x.size(0), w1.size(0), w3.size(0)
        h1 = torch.empty((n, hid), dtype=torch.float32, device=x.device)
        h2 = torch.empty((n, hid), dtype=torch.float32, device=x.device)
        out = torch.empty((n, out_dim), dtype=torch.float32, device=x.device)
        _lib.gemm(KGB_NT, x, w1, h1, n, hid, x.size(1), bias=b1, relu=True)
        _lib.gemm(KGB_NT, h1, w2, h2, n, hid, hid, bias=b2, relu=True)
        _lib.gemm(KGB_NT, h2, w3, out, n, out_dim, hid, bias=b3)
        ctx.save_for_backward(x, w1, w2, w3, h1, h2)
        return out

    @staticmethod
    @_lib.on_device_of
    def backward(ctx, g):
        x, w1, w2, w3, h1, h2 = ctx.saved_tensors
        g = g.contiguous()
        n, hid, out_dim = x.size(0), w1.size(0), w3.size(0)
        dev = x.device
        need = ctx.needs_input_grad

        def new(*shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)
        dw3, db3 = new(out_dim, hid), new(out_dim)
        _lib.gemm(KGB_TN, g, h2, dw3, out_dim, hid, n)
        _lib.wcolsum(g, out_dim, db3)
        g2 = new(n, hid)
        _lib.gemm(KGB_NN, g, w3, g2, n, hid, out_dim)
        s2, g2m = new(2, hid), new(n, hid)
        _lib.relu_bwd_fused(g2m, hid, dy=g2, y=h2, sums=s2)           # ReLU mask + bias gradient in one pass
        g2 = g2m
        dw2 = new(hid, hid)
        _lib.gemm(KGB_TN, g2, h1, dw2, hid, hid, n)
        g1 = new(n, hid)
        _lib.gemm(KGB_NN, g2, w2, g1, n, hid, hid)
        s1, g1m = new(2, hid), new(n, hid)
        _lib.relu_bwd_fused(g1m, hid, dy=g1, y=h1, sums=s1)
        g1 = g1m
        dw1 = new(hid, x.size(1))
        _lib.gemm(KGB_TN, g1, x, dw1, hid, x.size(1), n)
        dx = None
        if need[0]:
            dx = new(n, x.size(1))
            _lib.gemm(KGB_NN, g1, w1, dx, n, x.size(1), hid)
        return dx, dw1, s1[0].clone(), dw2, s2[0].clone(), dw3, db3


class SimpleMLP(nn.Module):
    """kgwas/model.py:10-22: same sub-module names (``FC_hidden`` / ``FC_hidden2`` / ``FC_output``), hence the same
    state-dict keys; the ``nn.Linear`` modules only hold the parameters.  CUDA fp32 inputs of a supported shape run on
    the engine's GEMM (bias + ReLU in the epilogue); anything else -- CPU tensors (config 1,
    the plumbing run), widths the GEMM cannot take -- goes through ``torch.nn.functional.linear``."""

    def __init__(self, input_dim, hidden_dim, output_dim):
        super().__init__()
        self.FC_hidden = nn.Linear(input_dim, hidden_dim)
        self.FC_hidden2 = nn.Linear(hidden_dim, hidden_dim)
        self.FC_output = nn.Linear(hidden_dim, output_dim)
        self.ReLU = nn.ReLU()

    def _engine_ok(self, x):
        dims = (x.size(-1), self.FC_hidden.out_features, self.FC_output.out_features)
        return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.size(0) > 0 and USE_ENGINE_MLP
                and all(d % 4 == 0 for d in dims) and self.FC_hidden.out_features in (32, 64, 128, 256, 384, 512))

    def forward(self, x):
        if self._engine_ok(x):
            return _MlpFn.apply(x, self.FC_hidden.weight, self.FC_hidden.bias, self.FC_hidden2.weight,
                                self.FC_hidden2.bias, self.FC_output.weight, self.FC_output.bias)
        h = self.ReLU(self.FC_hidden(x))
        h = self.ReLU(self.FC_hidden2(h))
        return self.FC_output(h)


USE_ENGINE_MLP = True


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

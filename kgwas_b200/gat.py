"""Drop-in ``GATConv`` (the in-tree fork kgwas/conv.py:36-232) and the fused heterogeneous GAT layer.

Arithmetic, heads = 1 as KGWAS effectively uses it:
    H_s = X_s W_src^T,  a_s = <H_s, att_src>,  a_t = <H_t, att_dst>   (H_t = H_s for a same-type relation)
    alpha_e = softmax_{e in in(t)}( leaky_relu(a_s[src] + a_t[dst]) / T )     (or sigmoid / raw)
    out[t]  = sum_e alpha_e H_s[src(e)] + bias
The node logits are folded (a_s = X_s . (W_src^T att_src)), so H_t is never formed, and the
W_src product runs on whichever side of the relation has fewer rows (plan.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from ._lib import ATT_RAW, ATT_SIGMOID, ATT_SOFTMAX, KGB_NN, KGB_NT, KGB_TN
import functools

from .conv import Linear, glorot_
from .ops import _Sched
from .plan import BIG_EDGES, BIG_ROWS, LayerPlan, get_plan

EdgeType = Tuple[str, str, str]


class GATConv(nn.Module):
    """Same constructor / forward signature / parameter names as kgwas/conv.py:37-53,122-124."""

    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int, heads: int = 1,
                 concat: bool = True, negative_slope: float = 0.2, dropout: float = 0.0,
                 add_self_loops: bool = True, edge_dim: Optional[int] = None, fill_value="mean",
                 bias: bool = True, sigmoid_gat: bool = False, temperature: float = 1,
                 pheno_condition: bool = False, **kwargs):
        super().__init__()
        if heads != 1:
            raise NotImplementedError("kgwas_b200.GATConv supports heads=1 (KGWAS's Linear(hidden, 1) head only "
                                      "works with one head: kgwas/model.py:50; SURVEY.md section 8 a5)")
        if add_self_loops:
            raise NotImplementedError("add_self_loops=True is not used by KGWAS (kgwas/model.py:42 passes False)")
        if edge_dim is not None or pheno_condition or dropout != 0.0 or not concat or not bias:
            raise NotImplementedError("only the configuration KGWAS instantiates is implemented "
                                      "(edge_dim=None, pheno_condition=False, dropout=0, concat=True, bias=True)")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops, self.edge_dim, self.fill_value = add_self_loops, edge_dim, fill_value
        self.sigmoid_gat, self.temperature, self.pheno_condition = sigmoid_gat, temperature, pheno_condition
        if isinstance(in_channels, int):                                     # conv.py:81-84
            self.lin_src = Linear(in_channels, heads * out_channels, bias=False, weight_initializer="glorot")
            self.lin_dst = self.lin_src
        else:                                                                # conv.py:85-89
            self.lin_src = Linear(in_channels[0], heads * out_channels, False, weight_initializer="glorot")
            self.lin_dst = Linear(in_channels[1], heads * out_channels, False, weight_initializer="glorot")
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))     # conv.py:92-93
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.lin_edge = None
        self.register_parameter("att_edge", None)
        self.bias = nn.Parameter(torch.empty(heads * out_channels))
        self.reset_parameters()

    def reset_parameters(self):                                              # conv.py:112-120
        self.lin_src.reset_parameters()
        self.lin_dst.reset_parameters()
        glorot_(self.att_src)
        glorot_(self.att_dst)
        nn.init.zeros_(self.bias)

    def forward(self, x: Union[Tensor, Tuple[Tensor, Tensor]], edge_index: Tensor, edge_attr=None, size=None,
                pheno_emb=None, return_attention_weights=None, return_raw_attention_weights=None):
        if edge_attr is not None or pheno_emb is not None:
            raise NotImplementedError("edge_attr / pheno_emb are dead paths in KGWAS (edge_dim=None)")
        same = isinstance(x, Tensor)
        x_src, x_dst = (x, x) if same else x
        et = ("src", "to", "src") if same else ("src", "to", "dst")
        x_dict = {"src": x_src} if same else {"src": x_src, "dst": x_dst}
        kw = {}
        if isinstance(return_attention_weights, bool):
            kw["return_attention_weights_dict"] = {et: return_attention_weights}
        if return_raw_attention_weights:
            kw["return_raw_attention_weights_dict"] = {et: True}
        out = hetero_gat({et: self}, x_dict, {et: edge_index}, "sum", False, kw)
        return out[et[2]]

    def __repr__(self):
        return f"GATConv({self.in_channels}, {self.out_channels}, heads={self.heads})"


class _GatMeta:
    def __init__(self, plan: LayerPlan, node_types, h, relu, rel_scale, bip, slope, temperature, mode, want_alpha,
                 shard=None):
        self.plan, self.node_types, self.h, self.relu, self.rel_scale = plan, node_types, h, relu, rel_scale
        self.bip, self.slope, self.temperature, self.mode, self.want_alpha = bip, slope, temperature, mode, want_alpha
        # SNP-sharded execution (dist.py): rows of a shared destination type are partial sums over ranks; this rank adds
        # the bias (and takes the bias gradient) only on the rows it owns, the ReLU runs after the cross-rank sum, and a
        # softmax group whose in-edges come from the sharded type (SNP -> Gene) is normalised ACROSS ranks
        self.shard = shard
        self.root_range = shard.root_range if shard is not None else {}
        self.sharded_type = shard.sharded_type if shard is not None else None

    def fused_relu(self, T):
        return self.relu and T not in self.root_range

    def cross_rank_softmax(self, T, job):
        return (self.shard is not None and self.mode == ATT_SOFTMAX and T in self.root_range
                and job.src_type == self.sharded_type)


def _seg_sum(job, v):
    """Per softmax group: sum of a per-slot scalar (slot order), through kgb_spmm on a one-column gather table."""
    g0, ones = job.group_sum_csr()
    y = torch.empty((g0.n_rows, 32), dtype=torch.float32, device=v.device)
    _lib.spmm(g0, ones, y, 32, ew=v.contiguous())
    return y[:, 0].contiguous()


def _global_softmax(job, alpha_loc, z_raw, temperature):
    """alpha_loc = softmax over THIS rank's slots of each group; returns the softmax over the slots of all ranks.
    Per group, log-sum-exp_local = sum_j alpha_j (s_j - log alpha_j) (every term equals it; the alpha-weighted mean is
    exact and immune to underflowing alphas); the global normaliser follows from two all-reduces of [n_groups]."""
    import torch.distributed as dist
    s = z_raw / temperature
    lse = _seg_sum(job, alpha_loc * s - torch.xlogy(alpha_loc, alpha_loc))
    has = job.local_group_mask()
    neg_inf = torch.full_like(lse, float("-inf"))
    lse = torch.where(has, lse, neg_inf)
    m = lse.clone()
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    se = torch.exp(lse - m)
    dist.all_reduce(se, op=dist.ReduceOp.SUM)
    f = torch.where(has, torch.exp(lse - m - torch.log(se.clamp_min(1e-38))), torch.zeros_like(lse))
    return alpha_loc * f[job.slot_group()]


def _f(rows, cols, dev):
    return torch.empty((rows, cols), dtype=torch.float32, device=dev)


class HeteroGatLayerFn(torch.autograd.Function):
    """inputs: x per node type; per relation (plan.rel_order): lin_src.weight, att_src, att_dst, bias;
    then lin_dst.weight of every bipartite (src type != dst type) relation.
    outputs: out per destination type [, alpha per relation in COO order (not differentiable)].

    Like the SAGE layer, every launch is only RECORDED with ``ops._Sched`` (what it reads, what it writes, SNP-sized or
    not) and issued by the scheduler: SNP-sized kernels back to back on the caller's stream, the gene / GO sized chains on
    high-priority side streams.  Steps that need ordinary torch ops with allocations (the cross-rank softmax of a
    sharded run) flush the recorded launches first and run on the caller's stream."""

    @staticmethod
    @_lib.on_device_of
    def forward(ctx, meta: _GatMeta, *tensors):
        plan, h = meta.plan, meta.h
        nt, nr = len(meta.node_types), len(plan.rel_order)
        x = dict(zip(meta.node_types, [t.contiguous() for t in tensors[:nt]]))
        p = tensors[nt:]
        Wsrc = torch.stack(p[0:nr])
        As = torch.stack([t.reshape(h) for t in p[nr:2 * nr]])
        Ad = torch.stack([t.reshape(h) for t in p[2 * nr:3 * nr]])
        bias = torch.stack(p[3 * nr:4 * nr])
        wd = list(p[4 * nr:])
        Wdst = torch.stack([wd.pop(0) if meta.bip[i] else p[i] for i in range(nr)])   # same-type: H_t = H_s
        Vs = torch.bmm(As.unsqueeze(1), Wsrc).squeeze(1).contiguous()                 # v = W^T att   [nr, h]
        Vd = torch.bmm(Ad.unsqueeze(1), Wdst).squeeze(1).contiguous()
        ctx.set_materialize_grads(False)
        dev = Wsrc.device
        outs, saved = [], {}
        alphas: List[Optional[Tensor]] = [None] * nr
        sch = _Sched(dev)
        P = functools.partial
        keep = []
        for T in plan.dst_types:
            a, b = plan.rel_range[T]
            scale = meta.rel_scale[T]
            n_t = plan.num_nodes[T]
            out = _f(n_t, h, dev)
            bias_T = bias[a:b].sum(0) * scale
            jobs = plan.jobs[T]
            shared = T in meta.root_range                 # partial rows: bias on owned rows only, ReLU after the sum
            relu_T = meta.fused_relu(T)
            big_T = n_t >= BIG_ROWS
            for ji, job in enumerate(jobs):
                lo, hi, R, S = job.rel_ids[0], job.rel_ids[-1] + 1, job.R, job.src_type
                first, last = ji == 0, ji == len(jobs) - 1
                job.schedule(h)
                big_S, big_e = job.n_src >= BIG_ROWS, job.n_edges >= BIG_EDGES
                a_s, a_d = _f(job.n_src, R, dev), _f(n_t, R, dev)
                vs, vd = Vs[lo:hi], Vd[lo:hi]
                sch.run(big_S, P(_lib.rowdot, x[S], vs, a_s, h, R, 0), (x[S], vs), (a_s,), f"fwd a_s {S}->{T}", T)
                sch.run(big_T, P(_lib.rowdot, x[T], vd, a_d, h, R, 0), (x[T], vd), (a_d,), f"fwd a_d {S}->{T}", T)
                alpha = torch.empty(job.n_edges, dtype=torch.float32, device=dev)
                sch.run(big_e, P(_lib.gat_alpha, job.gcsr, a_s, a_d, R, job.mode == "af", alpha, meta.slope,
                                 meta.temperature, meta.mode), (a_s, a_d), (alpha,), f"fwd alpha {S}->{T}", T)
                if meta.cross_rank_softmax(T, job):
                    sch.join()                                       # torch ops + collectives: caller's stream
                    z_raw = torch.empty_like(alpha)
                    _lib.gat_alpha(job.gcsr, a_s, a_d, R, job.mode == "af", z_raw, meta.slope, meta.temperature, ATT_RAW)
                    alpha = _global_softmax(job, alpha, z_raw, meta.temperature)
                beta = 0.0 if first else 1.0
                bias_arg = bias_T if (last and not shared) else None
                if job.mode == "xf":
                    z = _f(job.n_src, R * h, dev)
                    wcat = Wsrc[lo:hi].reshape(R * h, h)
                    sch.run(big_S, P(_lib.gemm, KGB_NT, x[S], wcat, z, job.n_src, R * h, h, alpha=scale), (x[S], wcat), (z,),
                            f"fwd Z {S}->{T}", T)
                    sch.run(big_T or big_e, P(_lib.spmm, job.csr, z.view(job.n_src * R, h), out, h, ew=alpha, beta=beta,
                                              bias=bias_arg, relu=relu_T and last),
                            (z, alpha, out, bias_arg), (out,), f"fwd spmm xf {S}->{T}", T)
                    A = None
                    keep.append((z, wcat))
                else:
                    A = _f(n_t, R * h, dev)
                    sch.run(big_e or big_S, P(_lib.spmm, job.csr, x[S], A.view(n_t * R, h), h, ew=alpha), (x[S], alpha), (A,),
                            f"fwd spmm af {S}->{T}", T)
                    wcat_t = Wsrc[lo:hi].permute(1, 0, 2).reshape(h, R * h)
                    sch.run(big_T, P(_lib.gemm, KGB_NT, A, wcat_t, out, n_t, h, R * h, alpha=scale, beta=beta,
                                     bias=bias_arg, relu=relu_T and last), (A, wcat_t, out, bias_arg), (out,),
                            f"fwd gemm af {S}->{T}", T)
                    keep.append(wcat_t)
                saved[(T, ji)] = (a_s, a_d, alpha, A)
                if meta.want_alpha:
                    coo = torch.empty_like(alpha)
                    eperm = job.eperm_long()

                    def to_coo(coo=coo, eperm=eperm, alpha=alpha):
                        coo[eperm] = alpha                                       # slot order -> COO order
                    sch.run(big_e, to_coo, (alpha, eperm), (coo,), f"fwd alpha coo {S}->{T}", T)
                    for k, rid in enumerate(job.rel_ids):
                        alphas[rid] = coo[job.edge_offsets[k]:job.edge_offsets[k + 1]].unsqueeze(-1)
            if shared:
                r0, r1 = meta.root_range[T]

                def add_bias(o=out[r0:r1], bias_T=bias_T):
                    o.add_(bias_T)
                sch.run(big_T, add_bias, (out, bias_T), (out,), f"fwd bias {T}", T)
            keep.append(bias_T)
            outs.append(out)
        sch.keep.append((keep, saved, Vs, Vd))
        sch.join()
        ctx.meta, ctx.saved = meta, saved
        ctx.save_for_backward(Wsrc, Wdst, As, Ad, Vs, Vd, *[x[t] for t in meta.node_types], *outs)
        if meta.want_alpha:
            ctx.mark_non_differentiable(*alphas)
            return (*outs, *alphas)
        return tuple(outs)

    @staticmethod
    @_lib.on_device_of
    def backward(ctx, *grads):
        meta: _GatMeta = ctx.meta
        plan, h = meta.plan, meta.h
        sv = ctx.saved_tensors
        Wsrc, Wdst, As, Ad, Vs, Vd = sv[:6]
        nt, nr = len(meta.node_types), len(plan.rel_order)
        x = dict(zip(meta.node_types, sv[6:6 + nt]))
        outs = dict(zip(plan.dst_types, sv[6 + nt:]))
        need_x = dict(zip(meta.node_types, ctx.needs_input_grad[1:1 + nt]))
        dev = Wsrc.device
        # one gradient tensor per job / destination type (NOT slices of one stacked tensor): the scheduler tracks
        # dependencies per storage, and slices of a shared buffer would chain every job behind the previous one
        dW_parts: Dict[int, Tensor] = {}      # first relation id of the job -> [R, h, h]
        dVs_parts: Dict[int, Tensor] = {}
        dVd_parts: Dict[int, Tensor] = {}
        dbias_parts: Dict[int, Tensor] = {}   # first relation id of the destination type -> [b - a, h]
        used = [False] * nr
        dx: Dict[str, Optional[Tensor]] = {t: None for t in meta.node_types}

        def target(t):
            if dx[t] is None:
                dx[t] = _f(plan.num_nodes[t], h, dev)
                return dx[t], 0.0
            return dx[t], 1.0

        sch = _Sched(dev)
        P = functools.partial
        keep = []
        order = sorted(range(len(plan.dst_types)), key=lambda i: -plan.num_nodes[plan.dst_types[i]])
        for T, d_out in [(plan.dst_types[i], grads[i]) for i in order]:
            if d_out is None:
                continue
            a, b = plan.rel_range[T]
            scale = meta.rel_scale[T]
            n_t = plan.num_nodes[T]
            big_T = n_t >= BIG_ROWS
            for i in range(a, b):
                used[i] = True
            if meta.fused_relu(T):
                # ReLU mask, aggregation scale and the bias gradient (column sums) in one pass over the rows
                g = _f(n_t, h, dev)
                sums = _f(2, h, dev)
                dy = d_out.contiguous()

                dbias_T = dbias_parts[a] = _f(b - a, h, dev)

                def relu_bias(g=g, dy=dy, y=outs[T], scale=scale, sums=sums, dst=dbias_T):
                    _lib.relu_bwd_fused(g, h, dy=dy, y=y, scale=scale, sums=sums)
                    dst.copy_(sums[0].expand_as(dst))
                sch.run(big_T, relu_bias, (dy, outs[T]), (g, sums, dbias_T), f"bwd relu {T}", T)
            else:
                g = d_out.contiguous()
                if scale != 1.0:
                    g = g * scale
                sch.main_made(g)
                db = torch.empty(h, dtype=torch.float32, device=dev)
                r0, r1 = meta.root_range.get(T, (0, n_t))      # the bias lives on the rows this rank owns
                dbias_T = dbias_parts[a] = _f(b - a, h, dev)

                def bias_grad(gr=g[r0:r1], db=db, dst=dbias_T):
                    _lib.wcolsum(gr, h, db)
                    dst.copy_(db.expand_as(dst))
                sch.run(big_T, bias_grad, (g,), (db, dbias_T), f"bwd bias {T}", T)
            for ji, job in enumerate(plan.jobs[T]):
                lo, hi, R, S = job.rel_ids[0], job.rel_ids[-1] + 1, job.R, job.src_type
                a_s, a_d, alpha, A = ctx.saved[(T, ji)]
                E = job.n_edges
                big_S, big_e = job.n_src >= BIG_ROWS, E >= BIG_EDGES
                dalpha = torch.empty(E, dtype=torch.float32, device=dev)
                du = torch.empty(E, dtype=torch.float32, device=dev)
                da_d = _f(n_t, R, dev)
                da_s = _f(job.n_src, R, dev)
                dW_j = dW_parts[lo] = _f(R * h, h, dev).view(R, h, h)
                dvs = dVs_parts[lo] = _f(R, h, dev)
                dvd = dVd_parts[lo] = _f(R, h, dev)
                cross = meta.cross_rank_softmax(T, job)
                if job.mode == "xf":
                    wcat = Wsrc[lo:hi].reshape(R * h, h)
                    z = _f(job.n_src, R * h, dev)
                    sch.run(big_S, P(_lib.gemm, KGB_NT, x[S], wcat, z, job.n_src, R * h, h), (x[S], wcat), (z,),
                            f"bwd Z {T}->{S}", T)                                     # recompute H_s (small side)
                    sch.run(big_T or big_e, P(_lib.sddmm, job.csr, g, z.view(job.n_src * R, h), h, dalpha), (g, z), (dalpha,),
                            f"bwd sddmm xf {T}->{S}", T)
                    if cross:
                        sch.join()
                        _dsoftmax(meta, T, job, a_s, a_d, R, False, alpha, dalpha, du, da_d)
                    else:
                        sch.run(big_e, P(_dsoftmax, meta, T, job, a_s, a_d, R, False, alpha, dalpha, du, da_d),
                                (a_s, a_d, alpha, dalpha), (du, da_d), f"bwd dsoftmax {T}->{S}", T)
                    dz = z                                                            # reuse the buffer
                    sch.run(big_T or big_e, P(_lib.spmm, job.tcsr, g, dz.view(job.n_src * R, h), h, ew=alpha,
                                              wperm=job.t_eperm, ew2=du, rowsum2=da_s, bins=1),
                            (g, alpha, du), (dz, da_s), f"bwd spmm xf {T}->{S}", T)
                    sch.run(big_S, P(_lib.gemm, KGB_TN, dz, x[S], dW_j.view(R * h, h), R * h, h, job.n_src), (dz, x[S]),
                            (dW_j,), f"bwd dW xf {T}->{S}", T)
                    if need_x[S]:
                        buf, beta = target(S)
                        sch.run(big_S, P(_lib.gemm, KGB_NN, dz, wcat, buf, job.n_src, h, R * h, beta=beta), (dz, wcat, buf),
                                (buf,), f"bwd dx xf {T}->{S}", T)
                    keep.append(wcat)
                else:
                    wcat_t = Wsrc[lo:hi].permute(1, 0, 2).reshape(h, R * h)
                    gp = _f(n_t, R * h, dev)
                    sch.run(big_T, P(_lib.gemm, KGB_NN, g, wcat_t, gp, n_t, R * h, h), (g, wcat_t), (gp,),
                            f"bwd G' af {T}->{S}", T)                                 # G' = g . W_src per slot
                    sch.run(big_e or big_S, P(_lib.sddmm, job.csr, gp.view(n_t * R, h), x[S], h, dalpha), (gp, x[S]),
                            (dalpha,), f"bwd sddmm af {T}->{S}", T)
                    if cross:
                        sch.join()
                        _dsoftmax(meta, T, job, a_s, a_d, R, True, alpha, dalpha, du, da_d)
                    else:
                        sch.run(big_e, P(_dsoftmax, meta, T, job, a_s, a_d, R, True, alpha, dalpha, du, da_d),
                                (a_s, a_d, alpha, dalpha), (du, da_d), f"bwd dsoftmax {T}->{S}", T)
                    buf, beta = target(S)          # needed as the spmm output even if x[S] wants no grad
                    sch.run(big_e or big_S, P(_lib.spmm, job.tcsr, gp.view(n_t * R, h), buf, h, ew=alpha, wperm=job.t_eperm,
                                              ew2=du, rowsum2=da_s, bins=R, beta=beta),
                            (gp, alpha, du, buf), (buf, da_s), f"bwd spmm af {T}->{S}", T)
                    dwt = _f(h, R * h, dev)

                    def af_wgrad(g=g, A=A, dwt=dwt, n_t=n_t, R=R, dst=dW_j):
                        _lib.gemm(KGB_TN, g, A, dwt, h, R * h, n_t)
                        dst.copy_(dwt.view(h, R, h).permute(1, 0, 2))
                    sch.run(big_T, af_wgrad, (g, A), (dwt, dW_j), f"bwd dW af {T}->{S}", T)
                    keep.append(wcat_t)
                # node-logit paths: a_s = X_S . Vs^T, a_d = X_T . Vd^T
                vs, vd = Vs[lo:hi], Vd[lo:hi]
                sch.run(big_S, P(_lib.wcolsum, x[S], h, dvs, w=da_s, n_slots=R), (x[S], da_s), (dvs,), f"bwd dVs {T}->{S}", T)
                sch.run(big_T, P(_lib.wcolsum, x[T], h, dvd, w=da_d, n_slots=R), (x[T], da_d), (dvd,), f"bwd dVd {T}->{S}", T)
                if need_x[S]:
                    buf, beta = target(S)
                    sch.run(big_S, P(_lib.rank_update, da_s, vs, buf, h, R, beta), (da_s, vs, buf), (buf,),
                            f"bwd dx a_s {T}->{S}", T)
                if need_x[T]:
                    buf, beta = target(T)
                    sch.run(big_T, P(_lib.rank_update, da_d, vd, buf, h, R, beta), (da_d, vd, buf), (buf,),
                            f"bwd dx a_d {T}->{S}", T)
                keep.append((dalpha, du, da_d, da_s))
        sch.keep.append((keep, ctx.saved))
        sch.join()
        ctx.saved = None

        def assemble(parts, width, shape):
            """per-job pieces -> one [nr, ...] tensor in relation order (zeros where no gradient arrived)"""
            pieces, pos = [], 0
            for lo_ in sorted(parts):
                if lo_ > pos:
                    pieces.append(torch.zeros((lo_ - pos, *shape), dtype=torch.float32, device=dev))
                pieces.append(parts[lo_].reshape(-1, *shape))
                pos = lo_ + pieces[-1].size(0)
            if pos < nr:
                pieces.append(torch.zeros((nr - pos, *shape), dtype=torch.float32, device=dev))
            return torch.cat(pieces) if pieces else torch.zeros((nr, *shape), dtype=torch.float32, device=dev)

        dWsrc = assemble(dW_parts, h, (h, h))
        dVs, dVd = assemble(dVs_parts, h, (h,)), assemble(dVd_parts, h, (h,))
        dbias = assemble(dbias_parts, h, (h,))
        # fold the logit vectors back onto the parameters:  v = W^T att
        bip = getattr(plan, "_bip_mask", None)        # built once per plan (a host-to-device copy: not inside a graph capture)
        if bip is None or bip.device != dev or bip.numel() != len(meta.bip):
            bip = plan._bip_mask = torch.tensor(meta.bip, device=dev)
        outer_s = As.unsqueeze(2) * dVs.unsqueeze(1)                                  # d W_src from a_s
        outer_d = Ad.unsqueeze(2) * dVd.unsqueeze(1)                                  # d W_dst (or W_src) from a_t
        dWsrc = dWsrc + outer_s + torch.where(bip.view(-1, 1, 1), torch.zeros_like(outer_d), outer_d)
        dAs = torch.bmm(Wsrc, dVs.unsqueeze(2)).squeeze(2)
        dAd = torch.bmm(Wdst, dVd.unsqueeze(2)).squeeze(2)
        need_p = ctx.needs_input_grad[1 + nt:]
        gx = [dx[t] if need_x[t] else None for t in meta.node_types]
        gp_: List[Optional[Tensor]] = []
        for stacked, shape in ((dWsrc, None), (dAs, (1, 1, h)), (dAd, (1, 1, h)), (dbias, None)):
            for i in range(nr):
                if used[i] and need_p[len(gp_)]:
                    gp_.append(stacked[i] if shape is None else stacked[i].view(shape))
                else:
                    gp_.append(None)
        for i in range(nr):
            if meta.bip[i]:
                gp_.append(outer_d[i] if (used[i] and need_p[len(gp_)]) else None)
        return (None, *gx, *gp_)


def _dsoftmax(meta, T, job, a_s, a_d, R, src_is_node, alpha, dalpha, du, da_d):
    """Backward of the attention normalisation.  Groups that span ranks: S_g = sum_j alpha_j dalpha_j is summed over
    ranks first, dz_j = alpha_j (dalpha_j - S_g) / T is formed here and the kernel only applies the leaky-relu slope
    and the per-group sum (its RAW mode: dz = dalpha)."""
    if meta.cross_rank_softmax(T, job):
        import torch.distributed as dist
        s_g = _seg_sum(job, alpha * dalpha)
        dist.all_reduce(s_g, op=dist.ReduceOp.SUM)
        dz = (alpha * (dalpha - s_g[job.slot_group()]) / meta.temperature).contiguous()
        _lib.gat_dsoftmax(job.gcsr, a_s, a_d, R, src_is_node, alpha, dz, du, da_d, meta.slope, meta.temperature, ATT_RAW)
    else:
        _lib.gat_dsoftmax(job.gcsr, a_s, a_d, R, src_is_node, alpha, dalpha, du, da_d, meta.slope, meta.temperature,
                          meta.mode)


def hetero_gat(convs: Dict[EdgeType, GATConv], x_dict, edge_index_dict, aggr: str, relu: bool, kwargs_dict, shard=None):
    """Fused multi-relation GAT layer; mirrors HeteroConv + the patched ``group`` (kgwas/utils.py:53-71)."""
    ret = kwargs_dict.get("return_attention_weights_dict", {}) or {}
    raw = kwargs_dict.get("return_raw_attention_weights_dict", {}) or {}
    extra = set(kwargs_dict) - {"return_attention_weights_dict", "return_raw_attention_weights_dict"}
    if extra:
        raise NotImplementedError(f"unsupported HeteroConv kwargs for GAT: {sorted(extra)}")
    node_types = list(x_dict.keys())
    num_nodes = {t: int(v.size(0)) for t, v in x_dict.items()}
    if shard is not None:
        with shard.building_plan():
            plan = get_plan(edge_index_dict, num_nodes, frozenset(convs.keys()))
    else:
        plan = get_plan(edge_index_dict, num_nodes, frozenset(convs.keys()))
    if not plan.rel_order:
        return {}
    cs = [convs[et] for et in plan.rel_order]
    h = cs[0].out_channels
    want = [isinstance(ret.get(et), bool) for et in plan.rel_order]
    raws = [bool(raw.get(et)) for et in plan.rel_order]
    uniform = (len(set(want)) == 1 and len(set(raws)) == 1
               and len({(c.negative_slope, float(c.temperature), bool(c.sigmoid_gat)) for c in cs}) == 1)
    if not uniform:
        raise NotImplementedError("the fused GAT layer needs the same attention flags / slope / temperature on every "
                                  "relation of the layer (KGWAS sets them uniformly: model.py:40-42, utils.py:453-458)")
    for t in node_types:
        if x_dict[t].size(-1) != h or x_dict[t].dtype != torch.float32:
            raise NotImplementedError("fused hetero-GAT needs fp32 features of width hidden on every node type")
    bip = [et[0] != et[2] for et in plan.rel_order]
    for c, is_bip in zip(cs, bip):
        c.lin_src.materialize(h)
        if is_bip:
            c.lin_dst.materialize(h)      # same-type relations never touch lin_dst (conv.py:136-138)
    mode = ATT_SIGMOID if cs[0].sigmoid_gat else (ATT_RAW if raws[0] else ATT_SOFTMAX)
    rel_scale = {}
    for T in plan.dst_types:
        a, b = plan.rel_range[T]
        rel_scale[T] = 1.0 if aggr == "sum" else 1.0 / (b - a)
    meta = _GatMeta(plan, node_types, h, relu, rel_scale, bip, float(cs[0].negative_slope), float(cs[0].temperature),
                    mode, want[0], shard)
    args = [x_dict[t] for t in node_types]
    args += [c.lin_src.weight for c in cs] + [c.att_src for c in cs] + [c.att_dst for c in cs] + [c.bias for c in cs]
    args += [c.lin_dst.weight for c, is_bip in zip(cs, bip) if is_bip]
    res = HeteroGatLayerFn.apply(meta, *args)
    nd = len(plan.dst_types)
    out = dict(zip(plan.dst_types, res[:nd]))
    if shard is not None:
        out = shard.combine(out, relu)     # sum the partial rows of shared node types across ranks, then ReLU
    if not want[0]:
        return out
    alphas = dict(zip(plan.rel_order, res[nd:]))
    final = {}
    for T in plan.dst_types:
        rels = [et for et in edge_index_dict if et in alphas and et[2] == T]     # dict order, as HeteroConv appends
        atts = [(edge_index_dict[et], alphas[et]) for et in rels]
        final[T] = (out[T], atts[0]) if len(atts) == 1 else (out[T], atts)      # len==1 quirk of `group`
    return final

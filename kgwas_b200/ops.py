"""Autograd boundary of the engine: one ``torch.autograd.Function`` per heterogeneous layer.

Forward and backward call only the C-ABI kernels of libkgwas_b200 (via ``_lib``); PyTorch is used
for allocation, tiny parameter reshuffles ([h,h]-sized stack / permute / sum) and autograd
plumbing.  Semantics follow PyG ``HeteroConv`` + ``SAGEConv`` as used by kgwas/model.py:34-48,74
(SURVEY.md Appendix A.1-A.2); the per-relation Python loop of the reference becomes one merged
gather-reduce per (destination type, source type) pair -- see plan.py.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import KGB_NN, KGB_NT, KGB_TN
from .plan import LayerPlan


def _empty(rows, cols, like):
    return torch.empty((rows, cols), dtype=torch.float32, device=like.device)


MULTI_STREAM = True      # run independent kernel chains of a layer on a side stream (bench.py turns it off while it
                         # times single kernels for the roofline, so that their durations are not shared)
_SIDE = {}


class _Fork:
    """Fork / join of one side CUDA stream inside a layer.  Every buffer is allocated on the main stream BEFORE it is
    used on the side stream and kept alive until ``join`` (the caching allocator ties a block to its allocation
    stream), so no ``record_stream`` bookkeeping is needed."""

    def __init__(self, device):
        self.main = torch.cuda.current_stream(device)
        self.side = None
        self.keep = []
        if MULTI_STREAM:
            key = (device.index, self.main.cuda_stream)
            if key not in _SIDE:
                _SIDE[key] = torch.cuda.Stream(device)
            self.side = _SIDE[key]
            self.side.wait_stream(self.main)

    def stream(self, on_side: bool):
        return torch.cuda.stream(self.side if (on_side and self.side is not None) else self.main)

    def sync_side_after_main(self):
        """everything enqueued on main so far happens-before what the side stream is given next"""
        if self.side is not None:
            self.side.wait_stream(self.main)

    def join(self):
        if self.side is not None:
            self.main.wait_stream(self.side)
        self.keep.clear()


class _SageLayerCtx:
    """Static description handed to the autograd Function (not a tensor)."""

    def __init__(self, plan: LayerPlan, node_types: List[str], h: int, relu: bool, rel_scale: Dict[str, float],
                 root_range: Optional[Dict[str, tuple]] = None, head_type: Optional[str] = None):
        self.plan, self.node_types, self.h, self.relu, self.rel_scale = plan, node_types, h, relu, rel_scale
        # single-output head (kgwas/model.py:50,83) fused into this layer: the Function takes the head weight [1,h]
        # as one more input and returns <relu(out[head_type]), w> as one more output [N,1]
        self.head_type = head_type
        # SNP-sharded execution (dist.py): for a destination type whose rows are shared by all ranks, this rank
        # adds the root term (and bias) only for rows [lo, hi); the partial outputs are summed across ranks and
        # the ReLU runs after that sum, outside this Function.
        self.root_range = root_range or {}

    def fused_relu(self, T: str) -> bool:
        return self.relu and T not in self.root_range


class HeteroSageLayerFn(torch.autograd.Function):
    """out[T] = act( scale_T * sum_{r into T} [ lin_l^r(mean_{e in r} x_src) + lin_r^r(x_T) ] )

    inputs : one [N,h] tensor per node type in ``meta.node_types``, then per relation (in
             ``plan.rel_order`` order) lin_l.weight [h,h] x R, lin_l.bias [h] x R, lin_r.weight [h,h] x R.
    outputs: one [N_T,h] tensor per destination type in ``plan.dst_types`` order.
    Relations whose destination type receives no gradient get ``None`` (not zeros), exactly like
    autograd in the reference (those parameters are then skipped by Adam, kgwas/kgwas.py:116,151).
    With ``meta.head_type`` set: one more input (head weight ``[1,h]``) and one more output
    ``[N_head, 1] = out[head_type] . w^T`` written by the epilogue of the last kernel that touches those rows; its
    backward is folded into the ReLU-mask / bias-gradient pass (kgb_relu_bwd_fused).
    """

    @staticmethod
    def forward(ctx, meta: _SageLayerCtx, *tensors):
        plan, h = meta.plan, meta.h
        nt, nr = len(meta.node_types), len(plan.rel_order)
        xs = tensors[:nt]
        Wl = torch.stack(tensors[nt:nt + nr])
        bl = torch.stack(tensors[nt + nr:nt + 2 * nr])
        Wr = torch.stack(tensors[nt + 2 * nr:nt + 3 * nr])
        head_T = meta.head_type
        w_head = tensors[nt + 3 * nr].contiguous() if head_T is not None else None
        pred = None
        ctx.set_materialize_grads(False)
        x = dict(zip(meta.node_types, [t.contiguous() for t in xs]))
        outs, saved_A = [], {}
        # ---- phase 1 (main stream): allocate every buffer, do the [h,h]-sized parameter reshuffles -------------
        prep = {}
        for T in plan.dst_types:
            a, b = plan.rel_range[T]
            scale = meta.rel_scale[T]
            n_t = plan.num_nodes[T]
            bias = bl[a:b].sum(0)
            if scale != 1.0:
                bias = bias * scale
            job_bufs = []
            for job in plan.jobs[T]:
                lo, hi = job.rel_ids[0], job.rel_ids[-1] + 1
                job.schedule(h)
                if job.mode == "xf":       # Z = X_src . [W_1;..;W_R]^T  (view: rows k*h.. = W_l^k)
                    job_bufs.append((Wl[lo:hi].reshape(job.R * h, h), _empty(job.n_src, job.R * h, Wl)))
                else:                      # A = gather-reduce, then out += A . [W_1|..|W_R]^T
                    job_bufs.append((Wl[lo:hi].permute(1, 0, 2).reshape(h, job.R * h), _empty(n_t, job.R * h, Wl)))
            prep[T] = (_empty(n_t, h, Wl), Wr[a:b].sum(0), bias, job_bufs)
            if T == head_T:
                pred = _empty(n_t, 1, Wl)
        # ---- phase 2: one kernel chain per destination type; the largest type (SNP) on the main stream, the
        # others on the side stream: their many small launches hide behind the big gather-reduce kernels ------------
        fork = _Fork(Wl.device)
        big = max(plan.dst_types, key=lambda t: plan.num_nodes[t])
        for T in plan.dst_types:
            scale = meta.rel_scale[T]
            n_t = plan.num_nodes[T]
            out, w_root, bias, job_bufs = prep[T]
            jobs = plan.jobs[T]
            relu_T = meta.fused_relu(T)
            # the head dot product rides on the last gather-reduce into these rows (else: one rowdot pass below)
            head_in_spmm = T == head_T and bool(jobs) and jobs[-1].mode == "xf"
            with fork.stream(T != big):
                # root term first (dense, overwrites), then every job accumulates; the last writer applies the ReLU.
                # (A gather-reduce that accumulates re-reads one row per warp, which hides latency far better than
                # a GEMM epilogue re-reading C.)
                if T in meta.root_range:
                    r0, r1 = meta.root_range[T]
                    out.zero_()
                    _lib.gemm(KGB_NT, x[T][r0:r1], w_root, out[r0:r1], r1 - r0, h, h, alpha=scale, bias=bias)
                else:
                    _lib.gemm(KGB_NT, x[T], w_root, out, n_t, h, h, alpha=scale, bias=bias, relu=relu_T and not jobs)
                for ji, job in enumerate(jobs):
                    R, xs_ = job.R, x[job.src_type]
                    last_relu = relu_T and ji == len(jobs) - 1
                    w_job, buf = job_bufs[ji]
                    if job.mode == "xf":
                        _lib.gemm(KGB_NT, xs_, w_job, buf, job.n_src, R * h, h, alpha=scale)
                        hd = head_in_spmm and ji == len(jobs) - 1
                        _lib.spmm(job.csr, buf.view(job.n_src * R, h), out, h, ew=job.w_mean, beta=1.0, relu=last_relu,
                                  dot_w=w_head if hd else None, dot_out=pred if hd else None)
                    else:
                        _lib.spmm(job.csr, xs_, buf.view(n_t * R, h), h, ew=job.w_mean)
                        _lib.gemm(KGB_NT, buf, w_job, out, n_t, h, R * h, alpha=scale, beta=1.0, relu=last_relu)
                        saved_A[(T, ji)] = buf
                if T == head_T and not head_in_spmm:
                    _lib.rowdot(out, w_head, pred, h, 1, 0)
            outs.append(out)
        fork.keep.append(prep)
        fork.join()
        ctx.meta = meta
        ctx.saved_A = saved_A
        ctx.save_for_backward(Wl, Wr, *[x[t] for t in meta.node_types], *outs, *([w_head] if head_T is not None else []))
        if head_T is not None:
            if head_T not in plan.dst_types or not meta.fused_relu(head_T):
                raise _lib.KgbError("fused head needs its node type to be a destination with the ReLU fused")
            return (*outs, pred)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *d_outs):
        meta: _SageLayerCtx = ctx.meta
        plan, h = meta.plan, meta.h
        saved = ctx.saved_tensors
        Wl, Wr = saved[0], saved[1]
        n_types, nr = len(meta.node_types), len(plan.rel_order)
        x = dict(zip(meta.node_types, saved[2:2 + n_types]))
        outs = dict(zip(plan.dst_types, saved[2 + n_types:]))
        need_x = dict(zip(meta.node_types, ctx.needs_input_grad[1:1 + n_types]))
        need_p = ctx.needs_input_grad[1 + n_types:1 + n_types + 3 * nr]
        need_w = any(need_p)
        head_T = meta.head_type
        w_head = saved[-1] if head_T is not None else None
        d_pred = d_outs[len(plan.dst_types)] if head_T is not None else None
        d_w_head = None

        dWl = torch.zeros_like(Wl) if need_w else None
        dbl = torch.zeros((Wl.size(0), h), dtype=torch.float32, device=Wl.device) if need_w else None
        dWr = torch.zeros_like(Wr) if need_w else None
        used = [False] * nr                                      # relations that received a gradient
        dx: Dict[str, Optional[torch.Tensor]] = {t: None for t in meta.node_types}

        def dx_target(t):
            """(buffer, beta) for accumulating into d x[t]."""
            if dx[t] is None:
                dx[t] = _empty(plan.num_nodes[t], h, Wl)
                return dx[t], 0.0
            return dx[t], 1.0

        # largest destination type first: its dense d x = g . W_root then is the first (overwriting) writer.
        # Main stream: the d x chain (ReLU mask, NN GEMMs, transposed gather-reduces).  Side stream: everything that only
        # produces parameter gradients (column sums, TN split-K GEMMs) -- it reads g / dz / A and writes disjoint slices.
        fork = _Fork(Wl.device)
        order = sorted(range(len(plan.dst_types)), key=lambda i: -plan.num_nodes[plan.dst_types[i]])
        for T, d_out in [(plan.dst_types[i], d_outs[i]) for i in order]:
            dp = d_pred if T == head_T else None
            if d_out is None and dp is None:
                continue
            a, b = plan.rel_range[T]
            scale = meta.rel_scale[T]
            n_t = plan.num_nodes[T]
            sums = None
            if meta.fused_relu(T):
                # one pass: ReLU mask (+ the head's rank-1 gradient) -> g, bias gradient, head-weight gradient
                g = _empty(n_t, h, Wl)
                if need_w or dp is not None:
                    sums = _empty(2, h, Wl)
                _lib.relu_bwd_fused(g, h, dy=d_out.contiguous() if d_out is not None else None, y=outs[T],
                                    dp=dp.contiguous() if dp is not None else None, wv=w_head if dp is not None else None,
                                    scale=scale, sums=sums)
                if dp is not None:
                    d_w_head = sums[1:2]
            else:
                g = d_out.contiguous()
                if scale != 1.0:
                    g = g * scale
            fork.keep += [g, sums]
            for i in range(a, b):
                used[i] = True
            r0, r1 = meta.root_range.get(T, (0, n_t))          # rows whose root term this rank owns
            if need_w:
                db = torch.empty(h, dtype=torch.float32, device=g.device)
                dwr = _empty(h, h, g)
                fork.keep += [db, dwr]
                fork.sync_side_after_main()
                with fork.stream(True):
                    if sums is not None:
                        dbl[a:b] = sums[0]
                    else:
                        _lib.wcolsum(g[r0:r1], h, db)
                        dbl[a:b] = db
                    _lib.gemm(KGB_TN, g[r0:r1], x[T][r0:r1], dwr, h, h, r1 - r0)
                    dWr[a:b] = dwr
            if need_x[T]:
                buf, beta = dx_target(T)
                if (r0, r1) != (0, n_t) and beta == 0.0:
                    buf.zero_()
                    beta = 1.0
                _lib.gemm(KGB_NN, g[r0:r1], Wr[a:b].sum(0), buf[r0:r1], r1 - r0, h, h, beta=beta)
            for ji, job in enumerate(plan.jobs[T]):
                lo, hi = job.rel_ids[0], job.rel_ids[-1] + 1
                R, S = job.R, job.src_type
                if job.mode == "xf":
                    dz = _empty(job.n_src, R * h, g)
                    fork.keep.append(dz)
                    _lib.spmm(job.tcsr, g, dz.view(job.n_src * R, h), h, ew=job.w_mean_t)
                    if need_w:
                        fork.sync_side_after_main()
                        with fork.stream(True):
                            _lib.gemm(KGB_TN, dz, x[S], dWl[lo:hi].view(R * h, h), R * h, h, job.n_src)
                    if need_x[S]:
                        buf, beta = dx_target(S)
                        _lib.gemm(KGB_NN, dz, Wl[lo:hi].reshape(R * h, h), buf, job.n_src, h, R * h, beta=beta)
                else:
                    A = ctx.saved_A[(T, ji)]
                    if need_w:
                        dwt = _empty(h, R * h, g)                              # [h_out, R*h_in]
                        fork.keep.append(dwt)
                        fork.sync_side_after_main()
                        with fork.stream(True):
                            _lib.gemm(KGB_TN, g, A, dwt, h, R * h, n_t)
                            dWl[lo:hi] = dwt.view(h, R, h).permute(1, 0, 2)
                    if need_x[S]:
                        wcat_t = Wl[lo:hi].permute(1, 0, 2).reshape(h, R * h)
                        dA = _empty(n_t, R * h, g)
                        fork.keep.append(dA)
                        _lib.gemm(KGB_NN, g, wcat_t, dA, n_t, R * h, h)
                        buf, beta = dx_target(S)
                        _lib.spmm(job.tcsr, dA.view(n_t * R, h), buf, h, ew=job.w_mean_t, beta=beta)
        fork.keep.append(ctx.saved_A)
        fork.join()
        ctx.saved_A = None
        grads_x = []
        for t in meta.node_types:
            if need_x[t] and dx[t] is None:      # no gradient reached this type: autograd treats None as zero
                grads_x.append(None)
            else:
                grads_x.append(dx[t] if need_x[t] else None)
        grads_p = []
        for k, stacked in enumerate((dWl, dbl, dWr)):
            for i in range(nr):
                grads_p.append(stacked[i] if (used[i] and need_p[k * nr + i]) else None)
        if head_T is not None:
            return (None, *grads_x, *grads_p, d_w_head if ctx.needs_input_grad[1 + n_types + 3 * nr] else None)
        return (None, *grads_x, *grads_p)

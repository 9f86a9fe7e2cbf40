"""Autograd boundary of the engine: one ``torch.autograd.Function`` per heterogeneous layer.

Forward and backward call only the C-ABI kernels of libkgwas_b200 (via ``_lib``); PyTorch is used
for allocation, tiny parameter reshuffles ([h,h]-sized stack / permute / sum) and autograd
plumbing.  Semantics follow PyG ``HeteroConv`` + ``SAGEConv`` as used by kgwas/model.py:34-48,74
(SURVEY.md Appendix A.1-A.2); the per-relation Python loop of the reference becomes one merged
gather-reduce per (destination type, source type) pair -- see plan.py.
"""
from __future__ import annotations

import functools
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import KGB_NN, KGB_NT, KGB_TN
from .plan import BIG_EDGES, BIG_ROWS, LayerPlan


def _empty(rows, cols, like):
    return torch.empty((rows, cols), dtype=torch.float32, device=like.device)


MULTI_STREAM = True      # run the small kernels of a layer on a high-priority side stream (bench.py turns it off while
                         # it times single kernels for the roofline, so that their durations are not shared)
_SIDE = {}
_EVENTS: List["torch.cuda.Event"] = []
ISSUE_ORDER_ALWAYS = False   # tests: run the recorded launches in the scheduler's issue order even on a single stream
CHAIN_PRIORITY = __import__("os").environ.get("KGB_CHAIN_PRIORITY") == "1"   # issue_order: longest-chain-first among ready launches of one class (measured on B200, round 2:
                         # 6.44 ms/step with it vs 6.22 without -- kept as an experiment knob, off by default)
TRACE = None             # scratch/trace_step.py sets this to a list: (label, stream name, start event, end event) per launch


def issue_order(ops, chain_priority=None):
    """Pure scheduling logic of ``_Sched.join`` (unit-tested on CPU).

    ``ops``: program-ordered list of ``(big, read_keys, write_keys)``.  Returns ``(preds, succs, order)``: the
    dependency DAG implied by program order -- launch i depends on the last earlier writer of everything it reads or
    writes (RAW, WAW) and on every earlier reader of what it writes since that writer (WAR) -- and a topological issue
    order with three classes: "critical" launches (big ones and everything a big one transitively depends on) as early as
    their inputs allow; then the launches that do not depend on any big launch; "late" ones (small launches that
    transitively wait for a big kernel) last -- a side stream is a FIFO, a late launch queued early would hold up
    everything behind it for the duration of that big kernel.  Ties keep program order."""
    import heapq
    n = len(ops)
    preds = [set() for _ in range(n)]

    def overlap(a, b):
        """keys are opaque hashables, or (storage, first byte, end byte) intervals of one allocation"""
        if isinstance(a, tuple) and isinstance(b, tuple) and len(a) == 3 and len(b) == 3:
            return a[0] == b[0] and a[1] < b[2] and b[1] < a[2]
        return a == b

    def inside(a, b):      # a fully covered by b
        if isinstance(a, tuple) and isinstance(b, tuple) and len(a) == 3 and len(b) == 3:
            return a[0] == b[0] and b[1] <= a[1] and a[2] <= b[2]
        return a == b

    live = []              # [key, last writer, readers since]: regions written so far (a superseded region is dropped)
    for i, (_, rk, wk) in enumerate(ops):
        for k in rk:
            for e in live:
                if overlap(k, e[0]):
                    preds[i].add(e[1])
        for k in wk:
            for e in live:
                if overlap(k, e[0]):
                    preds[i].add(e[1])
                    preds[i].update(r for r, rkey in e[2] if overlap(k, rkey))
        for k in wk:
            live = [e for e in live if not inside(e[0], k)]
            live.append([k, i, []])
        for k in rk:
            hit = False
            for e in live:
                if overlap(k, e[0]) and e[1] != i:
                    e[2].append((i, k))
                    hit = True
            if not hit and not any(overlap(k, e[0]) for e in live):
                live.append([k, -1, [(i, k)]])     # read of something nobody wrote here: remember the reader (WAR)
        preds[i].discard(i)
        preds[i].discard(-1)
    succs = [[] for _ in range(n)]
    for i in range(n):
        for p in preds[i]:
            succs[p].append(i)
    crit = [bool(ops[i][0]) for i in range(n)]                 # big, or feeds a big launch
    for i in range(n - 1, -1, -1):
        if not crit[i]:
            crit[i] = any(crit[j] for j in succs[i])
    late = [False] * n                                         # waits (transitively) for a big launch
    for i in range(n):
        late[i] = any(ops[p][0] or late[p] for p in preds[i])
    cls = [0 if crit[i] else (2 if late[i] else 1) for i in range(n)]
    # Among launches of one class that are ready at the same time, the one with the longest chain of work behind it goes
    # first (big = 10, small = 1; ties: more dependent launches, then program order).  This is what puts the big
    # SNP -> Gene gather-reduce -- whose result still has to pass through a gene-sized GEMM -- ahead of the big
    # Gene -> SNP one that nothing waits for, so that the small GEMM runs under a big kernel instead of after the last one.
    tail = [0] * n
    desc = [0] * n
    for i in range(n - 1, -1, -1):
        tail[i] = (10 if ops[i][0] else 1) + max((tail[j] for j in succs[i]), default=0)
        desc[i] = sum(1 + desc[j] for j in succs[i])
    indeg = [len(preds[i]) for i in range(n)]
    if not (CHAIN_PRIORITY if chain_priority is None else chain_priority):
        tail = desc = [0] * n
    heap = [(cls[i], -tail[i], -desc[i], i) for i in range(n) if indeg[i] == 0]
    heapq.heapify(heap)
    order = []
    while heap:
        i = heapq.heappop(heap)[3]
        order.append(i)
        for j in succs[i]:
            indeg[j] -= 1
            if indeg[j] == 0:
                heapq.heappush(heap, (cls[j], -tail[j], -desc[j], j))
    return preds, succs, order


def _region(t):
    """(storage, first byte, end byte) touched by a tensor view: launches on disjoint row-slices of one buffer (the flat
    buffer that carries every shared node type through ONE all-reduce) do not depend on each other."""
    es = t.element_size()
    lo = t.storage_offset() * es
    extent = 1
    for size, stride in zip(t.shape, t.stride()):
        if size == 0:
            extent = 0
            break
        extent += (size - 1) * abs(stride)
    return (t.untyped_storage().data_ptr(), lo, lo + max(extent, 1) * es)


class _Sched:
    """Deferred multi-stream launch scheduler of one layer pass.

    A KGWAS layer is a handful of big kernels (everything that touches the 784 k SNP rows or the 8 M SNP<->Gene edges)
    and several dozen small ones (gene / GO sized GEMMs and gather-reduces, [h,h] parameter-gradient updates).  Run
    back to back the small ones cost as much wall time as the big ones while using a fraction of the GPU.

    ``run`` only RECORDS a launch (a closure with bound arguments, what it reads, what it writes, big or small, the
    chain it belongs to).  ``join`` derives the dependency DAG from program order (RAW / WAW / WAR on whole buffers,
    tracked per storage), then issues the launches in a topological order that puts first whatever a big kernel is
    (transitively) waiting for: big launches go to the caller's stream back to back, small ones to a few HIGH-PRIORITY
    side streams (one per chain, round-robin), so that the block scheduler slips their few CTAs in between the waves of
    whatever big kernel is running; cross-stream edges of the DAG become event waits.  Dependencies are never spelled
    out by hand and a stream is never blocked behind a launch it does not depend on.

    Every buffer is allocated on the caller's stream BEFORE ``join`` and kept alive until it returns (the caching
    allocator ties a block to its allocation stream), so no ``record_stream`` bookkeeping is needed.  Ordinary torch ops
    issued on the caller's stream while the list is being built are complete (in stream order) before any recorded
    launch starts."""

    N_SMALL = 3

    def __init__(self, device, chain_priority=None):
        self.device = device
        self.chain_priority = chain_priority
        on_gpu = device.type == "cuda"     # (CPU tensors never reach a kernel: _lib raises; tests stub _lib)
        self.main = torch.cuda.current_stream(device) if on_gpu else None
        self.ops = []          # (big, fn, read keys, write keys, label, chain)
        self.keep = []
        self.smalls = []
        if MULTI_STREAM and on_gpu:
            key = (device.index, self.main.cuda_stream)
            if key not in _SIDE:
                _SIDE[key] = [torch.cuda.Stream(device, priority=-1) for _ in range(self.N_SMALL)]
            self.smalls = _SIDE[key]

    def run(self, big: bool, fn, reads=(), writes=(), label="", chain=None):
        """Record a launch.  ``fn`` takes no arguments, launches kernels only (no allocation) and must not depend on
        variables that change after this call (bind them with functools.partial / default arguments)."""
        self.keep.append((reads, writes, fn))
        rk = [_region(t) for t in reads if t is not None]
        wk = [_region(t) for t in writes if t is not None]
        self.ops.append((bool(big), fn, rk, wk, label, chain))

    def main_made(self, *ts):
        """kept for call sites that produce operands with ordinary torch ops on the caller's stream: those are ordered
        before every recorded launch by construction (see class docstring); only keep the tensors alive."""
        self.keep.append(ts)

    def join(self):
        """Issue everything recorded so far; on return the caller's stream is ordered after all of it."""
        ops, self.ops = self.ops, []
        n = len(ops)
        if n == 0:
            self.keep.clear()
            return
        if not self.smalls:                                   # single stream: program order
            seq = issue_order([(o[0], o[2], o[3]) for o in ops], self.chain_priority)[2] if ISSUE_ORDER_ALWAYS else range(n)
            for i in seq:
                self._launch(ops[i][1], self.main, ops[i][4], True)
            self.keep.clear()
            return
        preds, succs, order = issue_order([(o[0], o[2], o[3]) for o in ops], self.chain_priority)
        chains, ev_of, st_of = {}, [None] * n, [None] * n
        start = _event_from_pool(self, 0)
        start.record(self.main)
        for st in self.smalls:
            st.wait_event(start)
        n_ev = 1
        for i in order:
            big, fn, _, _, label, chain = ops[i]
            if big:
                st = self.main
            else:
                if chain not in chains:
                    chains[chain] = self.smalls[len(chains) % len(self.smalls)]
                st = chains[chain]
            for p in preds[i]:
                if st_of[p] is not st:
                    st.wait_event(ev_of[p])
            self._launch(fn, st, label, big)
            ev = _event_from_pool(self, n_ev)
            n_ev += 1
            ev.record(st)
            ev_of[i], st_of[i] = ev, st
        for st in self.smalls:
            self.main.wait_stream(st)
        self.keep.clear()

    def _launch(self, fn, st, label, big):
        if st is None:                                        # no CUDA stream (stubbed kernels in the CPU tests)
            fn()
            return
        if TRACE is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
        if st is self.main:
            fn()
        else:
            with torch.cuda.stream(st):
                fn()
        if TRACE is not None:
            b.record(st)
            TRACE.append((label, "big" if st is self.main else f"small{self.smalls.index(st)}", a, b))


def _event_from_pool(_sched, i):
    while len(_EVENTS) <= i:
        _EVENTS.append(torch.cuda.Event())
    return _EVENTS[i]


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
        # the cross-rank sum of a shared type's partial rows (forward) / of its masked gradient (backward) is issued by
        # the layer itself, as one more recorded launch on a side stream, so that it overlaps the SNP-row kernels of the
        # same pass instead of sitting between two layers
        self.exchange = bool(self.root_range)

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
    @_lib.on_device_of
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
        saved_A = {}
        # ---- phase 1: allocate every buffer; the [h,h]-sized parameter reshuffles (summed root weights and biases,
        # transposed job weights) are RECORDED like any other launch -- ~20 tiny kernels that used to run on the caller's
        # stream ahead of the first SNP-sized kernel of every pass (~100 us per pass) now sit on the side streams
        sch = _Sched(Wl.device, chain_priority=True if meta.exchange else None)
        prep = {}
        shared = [T for T in plan.dst_types if meta.exchange and T in meta.root_range]
        flat_out, flat_off = None, {}
        if shared:        # one buffer for the partial rows of every shared type: ONE all-reduce per pass
            off = 0
            for T in shared:
                flat_off[T] = off
                off += plan.num_nodes[T]
            flat_out = _empty(off, h, Wl)
        for T in plan.dst_types:
            a, b = plan.rel_range[T]
            scale = meta.rel_scale[T]
            n_t = plan.num_nodes[T]
            bias = torch.empty(h, dtype=torch.float32, device=Wl.device)
            w_root = _empty(h, h, Wl)

            def root_operands(bias=bias, w_root=w_root, bl_s=bl[a:b], wr_s=Wr[a:b], scale=scale):
                torch.sum(bl_s, dim=0, out=bias)
                if scale != 1.0:
                    bias.mul_(scale)
                torch.sum(wr_s, dim=0, out=w_root)
            sch.run(False, root_operands, (bl[a:b], Wr[a:b]), (bias, w_root), f"prep root {T}", T)
            job_bufs = []
            for job in plan.jobs[T]:
                job.schedule(h)
                if getattr(job, "multi", False):               # one Z table for several source types
                    job_bufs.append((None, _empty(job.total_rows, h, Wl)))
                    continue
                lo, hi = job.rel_ids[0], job.rel_ids[-1] + 1
                if job.mode == "xf":       # Z = X_src . [W_1;..;W_R]^T  (view: rows k*h.. = W_l^k)
                    job_bufs.append((Wl[lo:hi].reshape(job.R * h, h), _empty(job.n_src, job.R * h, Wl)))
                else:                      # A = gather-reduce, then out += A . [W_1|..|W_R]^T
                    w_t = _empty(h, job.R * h, Wl)

                    def transpose_w(w_t=w_t, src=Wl[lo:hi], R=job.R):
                        w_t.view(h, R, h).copy_(src.permute(1, 0, 2))
                    sch.run(False, transpose_w, (Wl[lo:hi],), (w_t,), f"prep W af {job.src_type}->{T}", T)
                    job_bufs.append((w_t, _empty(n_t, job.R * h, Wl)))
            out_T = flat_out[flat_off[T]:flat_off[T] + n_t] if T in flat_off else _empty(n_t, h, Wl)
            prep[T] = (out_T, w_root, bias, job_bufs)
            if T == head_T:
                pred = _empty(n_t, 1, Wl)
        # ---- phase 2: record every launch with the scheduler (nothing runs yet), then let it issue them: big kernels
        # back to back on this stream, small chains on high-priority side streams, ordered by the data they touch ----
        # (sharded runs: the gather with a cross-rank sum behind it goes first -- measured at N = 2: 4.27 -> 4.16 ms/step; on one
        #  GPU the same order only moves the contention, 6.22 -> 6.37 ms)
        P = functools.partial
        for T in sorted(plan.dst_types, key=lambda t: -plan.num_nodes[t]):
            scale = meta.rel_scale[T]
            n_t = plan.num_nodes[T]
            out, w_root, bias, job_bufs = prep[T]
            jobs = plan.jobs[T]
            relu_T = meta.fused_relu(T)
            big_T = n_t >= BIG_ROWS
            # the head dot product rides on the last gather-reduce into these rows (else: one rowdot pass below)
            head_in_spmm = T == head_T and bool(jobs) and jobs[-1].mode == "xf" and h % 128 == 0
            # root term first (dense, overwrites), then every job accumulates; the last writer applies the ReLU.
            # (A gather-reduce that accumulates re-reads one row per warp, which hides latency far better than
            # a GEMM epilogue re-reading C.)
            if T in meta.root_range:
                r0, r1 = meta.root_range[T]

                def root(out=out, xr=x[T][r0:r1], w_root=w_root, o=out[r0:r1], m=r1 - r0, scale=scale, bias=bias):
                    out.zero_()
                    _lib.gemm(KGB_NT, xr, w_root, o, m, h, h, alpha=scale, bias=bias)
                sch.run(big_T, root, (x[T], w_root, bias), (out,), f"fwd root {T}", T)
            else:
                sch.run(big_T, P(_lib.gemm, KGB_NT, x[T], w_root, out, n_t, h, h, alpha=scale, bias=bias,
                                 relu=relu_T and not jobs), (x[T], w_root, bias), (out,), f"fwd root {T}", T)
            for ji, job in enumerate(jobs):
                last_relu = relu_T and ji == len(jobs) - 1
                w_job, buf = job_bufs[ji]
                big_e = job.n_edges >= BIG_EDGES
                if getattr(job, "multi", False):
                    for (S, R_i, lo_i, hi_i, n_i, off) in job.parts:
                        w_i = Wl[lo_i:hi_i].reshape(R_i * h, h)
                        z_i = buf[off:off + n_i * R_i].view(n_i, R_i * h)
                        sch.run(False, P(_lib.gemm, KGB_NT, x[S], w_i, z_i, n_i, R_i * h, h, alpha=scale), (x[S], w_i), (buf,),
                                f"fwd Z {S}->{T}", T)
                        sch.keep.append(w_i)
                    hd = head_in_spmm and ji == len(jobs) - 1
                    sch.run(big_T or big_e,
                            P(_lib.spmm, job.csr, buf, out, h, ew=job.w_mean, beta=1.0, relu=last_relu,
                              dot_w=w_head if hd else None, dot_out=pred if hd else None),
                            (buf, out, w_head if hd else None), (out, pred if hd else None), f"fwd spmm xf {job.src_type}->{T}", T)
                    continue
                R, xs_ = job.R, x[job.src_type]
                if job.mode == "xf":
                    sch.run(job.n_src >= BIG_ROWS, P(_lib.gemm, KGB_NT, xs_, w_job, buf, job.n_src, R * h, h, alpha=scale),
                            (xs_, w_job), (buf,), f"fwd Z {job.src_type}->{T}", T)
                    hd = head_in_spmm and ji == len(jobs) - 1
                    sch.run(big_T or big_e,
                            P(_lib.spmm, job.csr, buf.view(job.n_src * R, h), out, h, ew=job.w_mean, beta=1.0,
                              relu=last_relu, dot_w=w_head if hd else None, dot_out=pred if hd else None),
                            (buf, out, w_head if hd else None), (out, pred if hd else None),
                            f"fwd spmm xf {job.src_type}->{T}", T)
                else:
                    sch.run(big_e or job.n_src >= BIG_ROWS,
                            P(_lib.spmm, job.csr, xs_, buf.view(n_t * R, h), h, ew=job.w_mean), (xs_,), (buf,),
                            f"fwd spmm af {job.src_type}->{T}", T)
                    sch.run(big_T, P(_lib.gemm, KGB_NT, buf, w_job, out, n_t, h, R * h, alpha=scale, beta=1.0,
                                     relu=last_relu), (buf, w_job, out), (out,), f"fwd gemm af {job.src_type}->{T}", T)
                    saved_A[(T, ji)] = buf
            if T == head_T and not head_in_spmm:
                sch.run(big_T, P(_lib.rowdot, out, w_head, pred, h, 1, 0), (out, w_head), (pred,), "fwd head", T)

        if flat_out is not None:
            def exchange(buf=flat_out, relu=meta.relu):
                import torch.distributed as dist
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)          # partial rows of every rank -> the full rows
                if relu:
                    buf.relu_()
            sch.run(False, exchange, (flat_out,), (flat_out,), "fwd all-reduce shared types", ("comm",))
        outs = [prep[T][0] for T in plan.dst_types]
        sch.keep.append(prep)
        sch.join()
        ctx.meta = meta
        ctx.saved_A = saved_A
        # [h,h]-sized derived operands the backward needs again (pre-summed root weights, concatenated job weights):
        # kept instead of being recomputed by a dozen tiny launches at the head of the backward pass
        ctx.derived = {T: (prep[T][1], [wb[0] for wb in prep[T][3]]) for T in plan.dst_types}
        ctx.save_for_backward(Wl, Wr, *[x[t] for t in meta.node_types], *outs, *([w_head] if head_T is not None else []))
        if head_T is not None:
            if head_T not in plan.dst_types or not meta.fused_relu(head_T):
                raise _lib.KgbError("fused head needs its node type to be a destination with the ReLU fused")
            return (*outs, pred)
        return tuple(outs)

    @staticmethod
    @_lib.on_device_of
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

        # Largest destination type first: its dense d x = g . W_root then is the first (overwriting) writer.  Launches
        # are only recorded here; placement (SNP-sized kernels on this stream, everything else on high-priority side
        # streams), issue order and cross-stream ordering are the scheduler's.
        sch = _Sched(Wl.device, chain_priority=True if meta.exchange else None)
        P = functools.partial
        order = sorted(range(len(plan.dst_types)), key=lambda i: -plan.num_nodes[plan.dst_types[i]])
        late = []       # recorded after everything else: small consumers of a big gather-reduce (see below)
        # Sharded runs: d_out of a shared type is this rank's gradient w.r.t. the (ReLU-ed) SUM over ranks.  Mask every such
        # gradient into one flat buffer, sum it over ranks with ONE all-reduce -- every rank's partial rows entered the
        # forward sum with weight one -- and only then use the slices.
        g_shared = {}
        if meta.exchange:
            sh = [(plan.dst_types[i], d_outs[i]) for i in order
                  if plan.dst_types[i] in meta.root_range and d_outs[i] is not None]
            if sh:
                gflat = _empty(sum(plan.num_nodes[T] for T, _ in sh), h, Wl)
                off = 0
                for T, d_out in sh:
                    n_t = plan.num_nodes[T]
                    gT = g_shared[T] = gflat[off:off + n_t]
                    off += n_t
                    dy = d_out.contiguous()

                    def mask(g=gT, dy=dy, y=outs[T], relu=meta.relu, scale=meta.rel_scale[T]):
                        if relu:
                            _lib.relu_bwd_fused(g, h, dy=dy, y=y, scale=scale)
                        else:
                            torch.mul(dy, scale, out=g)
                    sch.run(False, mask, (dy, outs[T]), (gT,), f"bwd mask {T}", T)

                def exchange_bwd(buf=gflat):
                    import torch.distributed as dist
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                sch.run(False, exchange_bwd, (gflat,), (gflat,), "bwd all-reduce shared types", ("comm",))
        for T, d_out in [(plan.dst_types[i], d_outs[i]) for i in order]:
            dp = d_pred if T == head_T else None
            if d_out is None and dp is None:
                continue
            a, b = plan.rel_range[T]
            scale = meta.rel_scale[T]
            n_t = plan.num_nodes[T]
            big_T = n_t >= BIG_ROWS
            sums = None
            if meta.fused_relu(T):
                # one pass: ReLU mask (+ the head's rank-1 gradient) -> g, bias gradient, head-weight gradient
                g = _empty(n_t, h, Wl)
                if need_w or dp is not None:
                    sums = _empty(2, h, Wl)
                dy = d_out.contiguous() if d_out is not None else None
                dpc = dp.contiguous() if dp is not None else None
                sch.run(big_T, P(_lib.relu_bwd_fused, g, h, dy=dy, y=outs[T], dp=dpc, wv=w_head if dp is not None else None,
                                 scale=scale, sums=sums),
                        (dy, outs[T], dpc, w_head), (g, sums), f"bwd relu {T}", T)
                if dp is not None:
                    d_w_head = sums[1:2]
            elif T in g_shared:
                g = g_shared[T]                                   # masked and summed over ranks above
            else:
                g = d_out.contiguous()
                if scale != 1.0:
                    g = g * scale
                sch.main_made(g)
            for i in range(a, b):
                used[i] = True
            r0, r1 = meta.root_range.get(T, (0, n_t))          # rows whose root term this rank owns
            if need_w:
                db = torch.empty(h, dtype=torch.float32, device=g.device)
                dwr = _empty(h, h, g)

                def root_grads(sums=sums, gr=g[r0:r1], db=db, dbl_s=dbl[a:b], dWr_s=dWr[a:b], dwr=dwr):
                    if sums is not None:
                        dbl_s.copy_(sums[0].expand_as(dbl_s))
                    else:
                        _lib.wcolsum(gr, h, db)
                        dbl_s.copy_(db.expand_as(dbl_s))
                    dWr_s.copy_(dwr.expand_as(dWr_s))
                sch.run(big_T, P(_lib.gemm, KGB_TN, g[r0:r1], x[T][r0:r1], dwr, h, h, r1 - r0), (g, x[T]), (dwr,),
                        f"bwd dWr {T}", T)
                sch.run(big_T and sums is None, root_grads, (g, sums, dwr), (db, dbl, dWr), f"bwd rootgrads {T}", T)
            if need_x[T]:
                buf, beta = dx_target(T)
                w_root = ctx.derived[T][0]                                   # sum of the relations' root weights
                zero_first = (r0, r1) != (0, n_t) and beta == 0.0

                def root_dx(buf=buf, gr=g[r0:r1], w_root=w_root, o=buf[r0:r1], m=r1 - r0, zero_first=zero_first,
                            beta=1.0 if zero_first else beta):
                    if zero_first:
                        buf.zero_()
                    _lib.gemm(KGB_NN, gr, w_root, o, m, h, h, beta=beta)
                sch.run(big_T, root_dx, (g, w_root, buf), (buf,), f"bwd dx root {T}", T)
            for ji, job in enumerate(plan.jobs[T]):
                big_e = job.n_edges >= BIG_EDGES
                if getattr(job, "multi", False):
                    dz = _empty(job.total_rows, h, g)
                    sch.run(big_T or big_e, P(_lib.spmm, job.tcsr, g, dz, h, ew=job.w_mean_t), (g,), (dz,),
                            f"bwd spmm xf {T}->{job.src_type}", T)
                    for (S, R_i, lo_i, hi_i, n_i, off) in job.parts:
                        dz_i = dz[off:off + n_i * R_i].view(n_i, R_i * h)
                        if need_w:
                            sch.run(False, P(_lib.gemm, KGB_TN, dz_i, x[S], dWl[lo_i:hi_i].view(R_i * h, h), R_i * h, h, n_i),
                                    (dz, x[S]), (dWl,), f"bwd dWl xf {T}->{S}", T)
                        if need_x[S]:
                            buf, beta = dx_target(S)
                            w_nn = Wl[lo_i:hi_i].reshape(R_i * h, h)
                            sch.run(False, P(_lib.gemm, KGB_NN, dz_i, w_nn, buf, n_i, h, R_i * h, beta=beta), (dz, w_nn, buf),
                                    (buf,), f"bwd dx xf {T}->{S}", T)
                    continue
                lo, hi = job.rel_ids[0], job.rel_ids[-1] + 1
                R, S = job.R, job.src_type
                big_S = plan.num_nodes[S] >= BIG_ROWS
                if job.mode == "xf":
                    dz = _empty(job.n_src, R * h, g)
                    sch.run(big_T or big_e, P(_lib.spmm, job.tcsr, g, dz.view(job.n_src * R, h), h, ew=job.w_mean_t),
                            (g,), (dz,), f"bwd spmm xf {T}->{S}", T)
                    def consumers(T=T, S=S, job=job, dz=dz, lo=lo, hi=hi, R=R, big_S=big_S):
                        if need_w:
                            sch.run(big_S, P(_lib.gemm, KGB_TN, dz, x[S], dWl[lo:hi].view(R * h, h), R * h, h, job.n_src),
                                    (dz, x[S]), (dWl,), f"bwd dWl xf {T}->{S}", T)
                        if need_x[S]:
                            buf, beta = dx_target(S)
                            w_nn = Wl[lo:hi].reshape(R * h, h)
                            sch.run(big_S, P(_lib.gemm, KGB_NN, dz, w_nn, buf, job.n_src, h, R * h, beta=beta),
                                    (dz, w_nn, buf), (buf,), f"bwd dx xf {T}->{S}", T)
                    if (big_T or big_e) and not big_S:
                        # d x[S] accumulates contributions one after the other: a small GEMM that waits for this big
                        # gather-reduce must be the LAST writer of its buffer, not the first, or every other small
                        # contribution to d x[S] would be chained behind the big kernel
                        late.append(consumers)
                    else:
                        consumers()
                else:
                    A = ctx.saved_A[(T, ji)]
                    if need_w:
                        dwt = _empty(h, R * h, g)                              # [h_out, R*h_in]

                        def af_wgrad(g=g, A=A, dwt=dwt, n_t=n_t, R=R, dst=dWl[lo:hi]):
                            _lib.gemm(KGB_TN, g, A, dwt, h, R * h, n_t)
                            dst.copy_(dwt.view(h, R, h).permute(1, 0, 2))
                        sch.run(big_T, af_wgrad, (g, A), (dwt, dWl), f"bwd dWl af {T}->{S}", T)
                    if need_x[S]:
                        wcat_t = ctx.derived[T][1][ji]                         # [h, R*h] = [W_1 | .. | W_R]
                        dA = _empty(n_t, R * h, g)
                        sch.run(big_T, P(_lib.gemm, KGB_NN, g, wcat_t, dA, n_t, R * h, h), (g, wcat_t), (dA,),
                                f"bwd dA af {T}->{S}", T)
                        buf, beta = dx_target(S)
                        sch.run(big_S or big_e, P(_lib.spmm, job.tcsr, dA.view(n_t * R, h), buf, h, ew=job.w_mean_t, beta=beta),
                                (dA, buf), (buf,), f"bwd spmm af {T}->{S}", T)
        for rec in late:
            rec()
        sch.keep.append(ctx.saved_A)
        sch.join()
        ctx.saved_A = None
        ctx.derived = None
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

"""Execution plan of one heterogeneous convolution over a fixed ``edge_index_dict``.

The reference loops over relations in Python and materialises one ``[N_dst, h]`` tensor per
relation before stacking and summing them (PyG ``HeteroConv`` + ``group``; in-tree echo
kgwas/conv.py:17-32).  Here all relations between one (destination type, source type) pair are
merged into ONE bipartite job with a single CSR (+ transposed CSR for the backward pass):

* ``xf`` ("transform first", chosen when N_src <= N_dst, e.g. Gene -> SNP): the per-relation linear
  map is applied on the small source side, ``Z = X_src . [W_1; ..; W_R]^T`` (``[N_src, R*h]`` viewed
  as ``[N_src*R, h]``), and the job gathers row ``s*R + k`` of Z straight into the destination row;
* ``af`` ("aggregate first", N_src > N_dst, e.g. SNP -> Gene): the job reduces into virtual rows
  ``t*R + k`` (``A = [N_dst, R*h]``) and the linear map ``A . [W_1 | .. | W_R]^T`` runs on the small
  destination side.

Both are legal because mean / sum / attention-weighted aggregation commute with the linear map.
Edges of a job are sorted by (t, k) with ties in original edge order, so a softmax group
(relation k, destination t) is a contiguous slot range in either mode.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

from . import _lib

EdgeType = Tuple[str, str, str]
MAX_SLOTS = 8      # KGB_MAX_BINS of the C ABI
BIG_ROWS = 100_000       # an operand with at least this many rows (SNP-sized) makes a launch "big" (ops._Sched)
BIG_EDGES = 2_000_000    # ... and so does a gather-reduce over at least this many edges
DEG_REDUCE = None  # set by kgwas_b200.dist while a sharded plan is built (all-reduce of the group degrees)


class PairJob:
    def __init__(self, dst_type: str, src_type: str, rels: List[EdgeType], rel_ids: List[int],
                 edge_indices: List[torch.Tensor], n_src: int, n_dst: int):
        self.dst_type, self.src_type, self.rels, self.rel_ids = dst_type, src_type, rels, rel_ids
        self.R, self.n_src, self.n_dst = len(rels), n_src, n_dst
        self.mode = "xf" if n_src <= n_dst else "af"
        R = self.R
        dev = edge_indices[0].device
        sizes = [int(ei.size(1)) for ei in edge_indices]
        self.edge_offsets = [0]
        for s in sizes:
            self.edge_offsets.append(self.edge_offsets[-1] + s)
        self.n_edges = self.edge_offsets[-1]
        src = torch.cat([ei[0] for ei in edge_indices])
        dst = torch.cat([ei[1] for ei in edge_indices])
        slot = torch.cat([torch.full((s,), k, dtype=torch.int64, device=dev) for k, s in enumerate(sizes)])
        group = dst * R + slot                                   # softmax / mean group of every edge
        self._coo = None
        if self.mode == "xf":
            job_src, job_dst, n_js, n_jd = src * R + slot, dst, n_src * R, n_dst
            self._coo = (job_src, dst, slot)   # kept for LayerPlan(merge_xf=True): MultiXfJob re-bases the table rows
            presort = slot * n_src + src       # slots of a row ordered by (relation slot, source): groups stay contiguous
        else:
            job_src, job_dst, n_js, n_jd = src, group, n_src, n_dst * R
            presort = None                     # ordered by source: heavy rows sweep the gathered table front to back
        self.csr, self.eperm, self.tcsr, self.t_eperm = _lib.csr_build(job_src, job_dst, n_js, n_jd, sort_cols=True,
                                                                        presort_key=presort)
        # group bookkeeping (exact integer work; plan time only)
        deg = torch.bincount(group, minlength=n_dst * R)
        if DEG_REDUCE is not None:      # SNP-sharded graphs: in-degrees of shared destination nodes are global
            deg = DEG_REDUCE(deg, dst_type)
        self.group_deg = deg.to(torch.int32)
        w = (1.0 / deg.clamp(min=1).to(torch.float32))[group]    # SAGE mean: 1 / max(in-degree, 1)
        local_deg = torch.bincount(group, minlength=n_dst * R) if DEG_REDUCE is not None else deg
        self.w_mean = w[self.eperm.long()].contiguous()          # CSR slot order
        # the same weights in transposed-CSR slot order: static for SAGE, so the backward gather-reduce reads them
        # directly instead of through the t_eperm indirection (one dependent load less per slice)
        self.w_mean_t = self.w_mean[self.t_eperm.long()].contiguous()
        if self.mode == "xf":
            # slots are (t, k)-sorted; group row pointers over the same slot order
            gp = torch.zeros(n_dst * R + 1, dtype=torch.int64, device=dev)
            torch.cumsum(local_deg, 0, out=gp[1:])
            self.group_rowptr = gp.to(torch.int32)
        else:
            self.group_rowptr = self.csr.rowptr
        self._gcsr = None
        self._scheduled_h = None

    def schedule(self, h: int):
        """L2-window scheduling of the heavy segments for feature width h (idempotent)."""
        if self._scheduled_h != h:
            self.csr.schedule_for_l2(4 * h)
            self.tcsr.schedule_for_l2(4 * h)
            # hub rows of launches that gather from a table too big for L2 (the SNP rows: aggregate-first forward,
            # transform-first backward) are reduced from shared-memory tiles; the mean weights are static, so they are
            # baked into the plan (csrc/kgb_spmm_hub.cuh).  No-op for every other job.
            self.csr.build_hub(self.w_mean, h)
            self.tcsr.build_hub(self.w_mean_t, h)
            self._scheduled_h = h
        return self

    @property
    def gcsr(self) -> "_lib.Csr":
        """CSR whose rows are the softmax groups (t, k), over the same slot order as ``csr``."""
        if self._gcsr is None:
            self._gcsr = self.csr if self.mode == "af" else _lib.Csr(self.group_rowptr, self.csr.col,
                                                                      self.n_dst * self.R, self.csr.n_cols)
        return self._gcsr

    def eperm_long(self) -> torch.Tensor:
        """int64 copy of ``eperm`` (CSR slot -> COO edge), cached: index for the attention export."""
        if getattr(self, "_eperm_long", None) is None:
            self._eperm_long = self.eperm.long()
        return self._eperm_long

    def slot_group(self) -> torch.Tensor:
        """int64 [E]: softmax / mean group (t*R + k) of every CSR slot."""
        if getattr(self, "_slot_group", None) is None:
            rp = self.gcsr.rowptr.long()
            self._slot_group = torch.repeat_interleave(torch.arange(rp.numel() - 1, device=rp.device), rp[1:] - rp[:-1])
        return self._slot_group

    def local_group_mask(self) -> torch.Tensor:
        """bool [n_groups]: the group has at least one slot on this rank."""
        rp = self.gcsr.rowptr
        return (rp[1:] - rp[:-1]) > 0

    def group_sum_csr(self):
        """(CSR with the group row pointers and an all-zero column array, ones[1, 32]): kgb_spmm over it with a
        per-slot scalar as edge weight is the per-group sum of that scalar (cross-rank softmax statistics, gat.py)."""
        if getattr(self, "_gsum", None) is None:
            g = self.gcsr
            self._gsum = (_lib.Csr(g.rowptr, torch.zeros_like(g.col), g.n_rows, 1),
                          torch.ones((1, 32), dtype=torch.float32, device=g.col.device))
        return self._gsum

    def __repr__(self):
        return (f"PairJob({self.src_type}->{self.dst_type}, R={self.R}, mode={self.mode}, E={self.n_edges}, "
                f"heavy_rows={self.csr.n_hrows}/{self.tcsr.n_hrows})")


class MultiXfJob:
    """Several transform-first pair jobs into ONE destination type merged into one gather-reduce (SAGE only).

    The destination rows of e.g. ``Gene`` are fed by Gene->Gene, BP->Gene, MF->Gene and CC->Gene relations; as separate jobs
    they form a chain of dependent gene-sized launches (Z GEMM, gather-reduce, Z GEMM, gather-reduce, ...) that every
    rank of a sharded run repeats and that the SNP-row kernels have to share the SMs with.  Here the ``Z`` products of all
    sources are written into ONE table (part i owns rows ``[row_off_i, row_off_i + n_src_i * R_i)``, viewed as
    ``[n_src_i, R_i * h]``) and ONE CSR over the destination rows gathers from it: chain depth 2 instead of 2 per source
    type.  Mean weights are per (destination, relation), exactly as in the separate jobs."""

    mode = "xf"
    multi = True

    def __init__(self, dst_type: str, parts: List["PairJob"], n_dst: int):
        self.dst_type, self.n_dst = dst_type, n_dst
        self.parts = []                     # (src_type, R, lo, hi, n_src, row_off)
        dev = parts[0].csr.col.device
        srcs, dsts, groups = [], [], []
        row_off, slot_base = 0, 0
        r_tot = sum(p.R for p in parts)
        for p in parts:
            self.parts.append((p.src_type, p.R, p.rel_ids[0], p.rel_ids[-1] + 1, p.n_src, row_off))
            src_rows, dst, slot = p._coo                       # table row s*R+k inside the part, destination, slot k
            srcs.append(src_rows + row_off)
            dsts.append(dst)
            groups.append(dst * r_tot + slot_base + slot)
            row_off += p.n_src * p.R
            slot_base += p.R
        self.total_rows = row_off
        self.n_edges = sum(p.n_edges for p in parts)
        job_src, job_dst, group = torch.cat(srcs), torch.cat(dsts), torch.cat(groups)
        self.csr, self.eperm, self.tcsr, self.t_eperm = _lib.csr_build(job_src, job_dst, self.total_rows, n_dst, sort_cols=True)
        deg = torch.bincount(group, minlength=n_dst * r_tot)
        if DEG_REDUCE is not None:
            deg = DEG_REDUCE(deg, dst_type)
        w = (1.0 / deg.clamp(min=1).to(torch.float32))[group]
        self.w_mean = w[self.eperm.long()].contiguous()
        self.w_mean_t = self.w_mean[self.t_eperm.long()].contiguous()
        self.src_type = "+".join(p[0] for p in self.parts)
        self.n_src = max(p[4] for p in self.parts)
        self.rels = [et for p in parts for et in p.rels]
        self._scheduled_h = None

    def schedule(self, h: int):
        if self._scheduled_h != h:
            self.csr.schedule_for_l2(4 * h)
            self.tcsr.schedule_for_l2(4 * h)
            self._scheduled_h = h
        return self

    def __repr__(self):
        return f"MultiXfJob({self.src_type}->{self.dst_type}, parts={len(self.parts)}, E={self.n_edges})"


class LayerPlan:
    """All jobs of one ``edge_index_dict``; shared by every layer of the model."""

    def __init__(self, edge_index_dict: Dict[EdgeType, torch.Tensor], num_nodes: Dict[str, int],
                 conv_keys=None, merge_xf: bool = False):
        self.merge_xf = merge_xf
        self.edge_types: List[EdgeType] = [et for et in edge_index_dict
                                           if (conv_keys is None or et in conv_keys)
                                           and et[0] in num_nodes and et[2] in num_nodes]
        self.num_nodes = dict(num_nodes)
        # relation order: grouped by (dst type, src type) in first-appearance order, so that a job's
        # relations are contiguous in the stacked parameter tensors
        pair_order: "OrderedDict[Tuple[str, str], List[EdgeType]]" = OrderedDict()
        dst_order: List[str] = []
        for et in self.edge_types:
            if et[2] not in dst_order:
                dst_order.append(et[2])
        for T in dst_order:
            for et in self.edge_types:
                if et[2] == T:
                    pair_order.setdefault((T, et[0]), []).append(et)
        self.dst_types = dst_order
        self.rel_order: List[EdgeType] = [et for rels in pair_order.values() for et in rels]
        self.rel_index = {et: i for i, et in enumerate(self.rel_order)}
        self.jobs: Dict[str, List[PairJob]] = {T: [] for T in dst_order}
        self.rel_range: Dict[str, Tuple[int, int]] = {}
        pos = 0
        for (T, S), rels_all in pair_order.items():
            for c in range(0, len(rels_all), MAX_SLOTS):          # at most MAX_SLOTS relations per job
                rels = rels_all[c:c + MAX_SLOTS]
                ids = [self.rel_index[et] for et in rels]
                self.jobs[T].append(PairJob(T, S, rels, ids, [edge_index_dict[et] for et in rels],
                                            num_nodes[S], num_nodes[T]))
        for T in dst_order:
            n = sum(j.R for j in self.jobs[T])
            self.rel_range[T] = (pos, pos + n)
            pos += n
            # Jobs accumulate into the destination rows one after the other.  An aggregate-first job over a big edge
            # set (SNP -> Gene) ends with a small GEMM that has to wait for its big gather-reduce: put such jobs LAST,
            # so that the small jobs of this destination type are not chained behind a big kernel (ops._Sched).
            self.jobs[T].sort(key=lambda j: j.mode == "af" and (j.n_edges >= BIG_EDGES or j.n_src >= BIG_ROWS))
            if merge_xf:
                # the small transform-first jobs of this destination type become one gather-reduce over one Z table
                small = [j for j in self.jobs[T] if j.mode == "xf" and j.n_edges < BIG_EDGES and j.n_src < BIG_ROWS]
                if len(small) >= 2:
                    merged = MultiXfJob(T, small, num_nodes[T])
                    first = self.jobs[T].index(small[0])
                    rest = [j for j in self.jobs[T] if j not in small]
                    self.jobs[T] = rest[:first] + [merged] + rest[first:]
        for js in self.jobs.values():
            for j in js:
                if getattr(j, "_coo", None) is not None:
                    j._coo = None                                  # the COO copies were only needed for merging
        self.n_edges = sum(j.n_edges for js in self.jobs.values() for j in js)
        self._tensors = [edge_index_dict[et] for et in self.edge_types]   # identity anchors for the cache
        self._versions = [t._version for t in self._tensors]

    def matches(self, edge_index_dict, num_nodes, conv_keys, merge_xf=False) -> bool:
        if merge_xf != self.merge_xf:
            return False
        ets = [et for et in edge_index_dict if (conv_keys is None or et in conv_keys)
               and et[0] in num_nodes and et[2] in num_nodes]
        if ets != self.edge_types or num_nodes != self.num_nodes:
            return False
        return all(edge_index_dict[et] is t and t._version == v
                   for et, t, v in zip(self.edge_types, self._tensors, self._versions))


_CACHE: List[LayerPlan] = []
_CACHE_SIZE = 4
plan_builds = 0


def get_plan(edge_index_dict, num_nodes: Dict[str, int], conv_keys=None, merge_xf: bool = False) -> LayerPlan:
    """Plans are cached on tensor identity (+ in-place version), most recent first.  ``merge_xf``: the SAGE layer's
    variant with one multi-source gather-reduce per destination type (``MultiXfJob``)."""
    global plan_builds
    for i, p in enumerate(_CACHE):
        if p.matches(edge_index_dict, num_nodes, conv_keys, merge_xf):
            if i:
                _CACHE.insert(0, _CACHE.pop(i))
            return p
    for et, ei in edge_index_dict.items():
        if ei.dtype != torch.int64 or ei.dim() != 2 or ei.size(0) != 2:
            raise ValueError(f"edge_index of {et} must be int64 [2, E], got {ei.dtype} {tuple(ei.shape)}")
    # the plan's kernels (CSR build, heavy-row tables) launch on the CURRENT device: build under the graph's device so that
    # `KGWAS(data, device='cuda:1')` works without `torch.cuda.set_device` (kgwas/kgwas.py:38-39)
    dev = next((ei.device for ei in edge_index_dict.values() if ei.is_cuda), None)
    if dev is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            p = LayerPlan(edge_index_dict, num_nodes, conv_keys, merge_xf)
    else:
        p = LayerPlan(edge_index_dict, num_nodes, conv_keys, merge_xf)
    plan_builds += 1
    _CACHE.insert(0, p)
    del _CACHE[_CACHE_SIZE:]
    return p


def clear_plan_cache():
    _CACHE.clear()

"""Full-neighbour mini-batch loader: a PyG-free stand-in for
``NeighborLoader(data, num_neighbors=[-1] * L, input_nodes=('SNP', ids), batch_size, drop_last)``
as used by kgwas/kgwas.py:99-113.

A batch holds the L-hop *directional* subgraph of its seed SNPs: seeds come first in the SNP node
list (``batch['SNP'].batch_size``), hop k expands -- for every relation -- each destination node first
discovered in hop k-1 to ALL its in-neighbours, only traversed edges are kept, node attributes
(``x``, ``y``, ``n_id``) are row-sliced and edges relabelled to batch-local ids.  With L conv layers on
an L-hop batch the seed outputs equal the full-graph outputs.  Integer bookkeeping only; bit-exact
against oracle/bookkeeping.full_neighbor_subgraph_ref.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .graph import HeteroData

EdgeType = Tuple[str, str, str]


class _InAdj:
    """In-neighbour lists of one relation: CSC by destination, edges of a destination in original order."""

    def __init__(self, edge_index: np.ndarray, n_dst: int):
        dst = edge_index[1]
        self.order = np.argsort(dst, kind="stable")
        self.src_sorted = edge_index[0][self.order]
        self.ptr = np.zeros(n_dst + 1, dtype=np.int64)
        np.add.at(self.ptr, dst + 1, 1)
        np.cumsum(self.ptr, out=self.ptr)

    def expand(self, frontier: np.ndarray):
        """(edge ids, srcs, dst-of-edge) of all in-edges of the frontier nodes, frontier order then edge order."""
        start, stop = self.ptr[frontier], self.ptr[frontier + 1]
        cnt = stop - start
        total = int(cnt.sum())
        if total == 0:
            z = np.zeros(0, dtype=np.int64)
            return z, z, z
        offs = np.repeat(start - np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt)
        slots = np.arange(total, dtype=np.int64) + offs
        return self.order[slots], self.src_sorted[slots], np.repeat(frontier, cnt)


class FullNeighborSampler:
    def __init__(self, data: HeteroData, num_hops: int):
        self.data, self.num_hops = data, num_hops
        self.num_nodes = {t: data[t].num_nodes for t in data.node_types}
        self.edge_types: List[EdgeType] = list(data.edge_types)
        self.edges = {et: data[et].edge_index.cpu().numpy() for et in self.edge_types}
        self.adj = {et: _InAdj(self.edges[et], self.num_nodes[et[2]]) for et in self.edge_types}

    def sample(self, seed_type: str, seeds: np.ndarray):
        seeds = np.asarray(seeds, dtype=np.int64)
        local = {t: np.full(n, -1, dtype=np.int64) for t, n in self.num_nodes.items()}
        nodes: Dict[str, List[np.ndarray]] = {t: [] for t in self.num_nodes}
        count = {t: 0 for t in self.num_nodes}

        def add(t, cand):
            """append the not-yet-seen nodes of cand in first-occurrence order; returns them"""
            if cand.size == 0:
                return cand
            fresh = cand[local[t][cand] < 0]
            if fresh.size == 0:
                return fresh
            _, first = np.unique(fresh, return_index=True)
            new = fresh[np.sort(first)]
            local[t][new] = count[t] + np.arange(new.size)
            count[t] += new.size
            nodes[t].append(new)
            return new

        frontier = {t: np.zeros(0, dtype=np.int64) for t in self.num_nodes}
        frontier[seed_type] = add(seed_type, seeds)
        kept: Dict[EdgeType, List[np.ndarray]] = {et: [] for et in self.edge_types}
        for _ in range(self.num_hops):
            new_frontier: Dict[str, List[np.ndarray]] = {t: [] for t in self.num_nodes}
            for et in self.edge_types:
                s_t, _, d_t = et
                if frontier[d_t].size == 0:
                    continue
                eids, srcs, _ = self.adj[et].expand(frontier[d_t])
                kept[et].append(eids)
                new = add(s_t, srcs)
                if new.size:
                    new_frontier[s_t].append(new)
            frontier = {t: (np.concatenate(v) if v else np.zeros(0, dtype=np.int64)) for t, v in new_frontier.items()}
        node_ids = {t: (np.concatenate(v) if v else np.zeros(0, dtype=np.int64)) for t, v in nodes.items()}
        sub_edges, edge_ids = {}, {}
        for et in self.edge_types:
            s_t, _, d_t = et
            ids = np.concatenate(kept[et]) if kept[et] else np.zeros(0, dtype=np.int64)
            edge_ids[et] = ids
            ei = self.edges[et]
            sub_edges[et] = np.stack([local[s_t][ei[0][ids]], local[d_t][ei[1][ids]]]) if ids.size else \
                np.zeros((2, 0), dtype=np.int64)
        return node_ids, sub_edges, edge_ids


class GpuFullNeighborSampler:
    """The same L-hop expansion on the GPU (csrc/kgb_sampler.cu through the C ABI): per hop and relation -- in the
    reference's order, the local-id table updated after every relation -- ``kgb_frontier_count`` /
    ``kgb_frontier_expand`` list every in-edge of the frontier and ``kgb_frontier_add`` hands out batch-local ids in
    first-occurrence order.  Node lists, kept edge ids and relabelled edges are bit-identical to ``FullNeighborSampler``
    (tests/test_sampler_gpu.py); they stay on the device, so a batch is assembled without the host touching the graph."""

    def __init__(self, data: HeteroData, num_hops: int):
        from . import _lib
        self._lib = _lib
        self.num_hops = num_hops
        self.num_nodes = {t: data[t].num_nodes for t in data.node_types}
        self.edge_types: List[EdgeType] = list(data.edge_types)
        self.edges = {et: data[et].edge_index for et in self.edge_types}
        dev = next(iter(self.edges.values())).device
        self.device = dev
        self.adj = {}
        with torch.cuda.device(dev):                    # the C ABI launches on the current device: follow the data
            for et in self.edge_types:
                ei = self.edges[et]
                csr, eperm, _, _ = _lib.csr_build(ei[0], ei[1], self.num_nodes[et[0]], self.num_nodes[et[2]], transposed=False)
                self.adj[et] = (csr.rowptr, csr.col, eperm)             # in-edges of a node in original edge order
        self.local = {t: torch.full((n,), -1, dtype=torch.int32, device=dev) for t, n in self.num_nodes.items()}
        self.firstpos = {t: torch.full((n,), 2 ** 31 - 1, dtype=torch.int32, device=dev) for t, n in self.num_nodes.items()}

    def sample(self, seed_type: str, seeds):
        with torch.cuda.device(self.device):
            return self._sample(seed_type, seeds)

    def _sample(self, seed_type: str, seeds):
        lib, dev = self._lib, self.device
        seeds = torch.as_tensor(seeds, device=dev).to(torch.int32).contiguous()
        count = {t: 0 for t in self.num_nodes}
        nodes: Dict[str, List[torch.Tensor]] = {t: [] for t in self.num_nodes}

        def add(t, cand):
            new = lib.frontier_add(cand, self.local[t], self.firstpos[t], count[t])
            if new.numel():
                count[t] += int(new.numel())
                nodes[t].append(new)
            return new

        empty = torch.empty(0, dtype=torch.int32, device=dev)
        frontier = {t: empty for t in self.num_nodes}
        frontier[seed_type] = add(seed_type, seeds)
        kept: Dict[EdgeType, List[torch.Tensor]] = {et: [] for et in self.edge_types}
        for _ in range(self.num_hops):
            new_frontier: Dict[str, List[torch.Tensor]] = {t: [] for t in self.num_nodes}
            for et in self.edge_types:
                s_t, _, d_t = et
                if frontier[d_t].numel() == 0:
                    continue
                ptr, col, eperm = self.adj[et]
                eids, srcs = lib.frontier_expand(ptr, col, eperm, frontier[d_t])
                kept[et].append(eids)
                new = add(s_t, srcs)
                if new.numel():
                    new_frontier[s_t].append(new)
            frontier = {t: (torch.cat(v) if v else empty) for t, v in new_frontier.items()}
        node_ids = {t: (torch.cat(v) if v else empty).long() for t, v in nodes.items()}
        sub_edges, edge_ids = {}, {}
        for et in self.edge_types:
            s_t, _, d_t = et
            ids = (torch.cat(kept[et]) if kept[et] else empty).long()
            edge_ids[et] = ids
            ei = self.edges[et]
            if ids.numel():
                sub_edges[et] = torch.stack([self.local[s_t][ei[0][ids]].long(), self.local[d_t][ei[1][ids]].long()])
            else:
                sub_edges[et] = torch.zeros((2, 0), dtype=torch.int64, device=dev)
        for t, ids in node_ids.items():                 # leave the local-id tables clean for the next batch
            if ids.numel():
                self.local[t][ids] = -1
        return node_ids, sub_edges, edge_ids


class NeighborLoader:
    """Iterable of ``HeteroData`` mini-batches (same constructor keywords KGWAS passes)."""

    def __init__(self, data: HeteroData, num_neighbors: Sequence[int], input_nodes, batch_size: int = 1,
                 shuffle: bool = False, drop_last: bool = False, num_workers: int = 0, sampler=None, **kwargs):
        if any(n != -1 for n in num_neighbors):
            raise NotImplementedError("KGWAS samples full neighbourhoods (num_neighbors=[-1]*L, kgwas.py:99)")
        if shuffle:
            raise NotImplementedError("KGWAS never shuffles (kgwas.py:93-94)")
        self.data = data
        self.seed_type, ids = input_nodes
        self.ids = np.asarray(ids.cpu() if torch.is_tensor(ids) else ids, dtype=np.int64)
        self.batch_size, self.drop_last = batch_size, drop_last
        on_gpu = any(t.is_cuda for t in data.edge_index_dict.values())
        # graph resident on the GPU (KGWAS.train(data_to_cuda=True), kgwas.py:96-97): the batches are assembled there
        self.sampler = GpuFullNeighborSampler(data, len(num_neighbors)) if on_gpu else \
            FullNeighborSampler(data, len(num_neighbors))
        self.on_gpu = on_gpu

    def __len__(self):
        n = len(self.ids)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for b in range(len(self)):
            yield self.make_batch(self.ids[b * self.batch_size:(b + 1) * self.batch_size])

    def make_batch(self, seeds: np.ndarray) -> HeteroData:
        node_ids, sub_edges, edge_ids = self.sampler.sample(self.seed_type, seeds)
        as_t = (lambda a: a) if self.on_gpu else torch.from_numpy
        batch = HeteroData()
        for t in self.data.node_types:
            store = self.data[t]
            idx = as_t(node_ids[t])
            for key, val in store.items():
                if torch.is_tensor(val) and val.dim() >= 1 and val.size(0) == store.num_nodes:
                    batch[t][key] = val[idx.to(val.device)]
            if "n_id" not in batch[t]:
                batch[t].n_id = idx
        batch[self.seed_type].batch_size = int(len(seeds))
        batch[self.seed_type].input_id = torch.from_numpy(np.asarray(seeds))
        for et in self.data.edge_types:
            batch[et].edge_index = as_t(sub_edges[et])
            batch[et].e_id = as_t(edge_ids[et])
        return batch

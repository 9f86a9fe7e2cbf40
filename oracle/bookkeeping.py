"""Integer bookkeeping oracle (numpy / pure Python).  TEST INFRASTRUCTURE ONLY.

Restates, for bit-exact comparison with the CUDA / torch product code:
  * COO -> CSR / transposed CSR with stable edge order (the index handling PyG's
    MessagePassing.propagate + scatter perform implicitly; reached from kgwas/conv.py:182),
  * ToUndirected / AddSelfLoops as applied by kgwas/kgwas_data.py:271-272,
  * full-neighbour L-hop mini-batch extraction = NeighborLoader(num_neighbors=[-1]*L)
    (kgwas/kgwas.py:99-113) as a slow Python reference.
Parity status: PyG is absent from /root/reference (un-vendored dependency), so these follow the
published PyG 2.1-2.6 semantics; "parity unpinned" beyond the hand-made cases in tests/.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def csr_from_coo_ref(src: np.ndarray, dst: np.ndarray, n_src: int, n_dst: int, sort_cols: bool = False):
    """Returns rowptr, col, eperm, t_rowptr, t_col, t_eperm (see include/kgwas_b200.h)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    if sort_cols:      # slots of a row ordered by (source, original edge id)
        by_src = np.argsort(src, kind="stable")
        eperm = by_src[np.argsort(dst[by_src], kind="stable")]
    else:
        eperm = np.argsort(dst, kind="stable")
    col = src[eperm]
    rowptr = np.zeros(n_dst + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    t_eperm = np.argsort(col, kind="stable")          # stable w.r.t. CSR slot order
    t_col = dst[eperm][t_eperm]
    t_rowptr = np.zeros(n_src + 1, dtype=np.int64)
    np.add.at(t_rowptr, src + 1, 1)
    t_rowptr = np.cumsum(t_rowptr)
    i32 = lambda a: a.astype(np.int32)
    return i32(rowptr), i32(col), i32(eperm), i32(t_rowptr), i32(t_col), i32(t_eperm)


def heavy_segments_ref(rowptr: np.ndarray, seg_len: int):
    hrow_id, segptr, hseg_hrow = [], [0], []
    for r in range(len(rowptr) - 1):
        deg = int(rowptr[r + 1] - rowptr[r])
        if deg > seg_len:
            ns = (deg + seg_len - 1) // seg_len
            hseg_hrow += [len(hrow_id)] * ns
            hrow_id.append(r)
            segptr.append(segptr[-1] + ns)
    return (np.array(hrow_id, dtype=np.int32), np.array(segptr, dtype=np.int32),
            np.array(hseg_hrow, dtype=np.int32))


def coalesce_ref(row: np.ndarray, col: np.ndarray):
    pairs = sorted(set(zip(row.tolist(), col.tolist())))
    if not pairs:
        return np.zeros((2, 0), dtype=np.int64)
    return np.array(pairs, dtype=np.int64).T


def to_undirected_ref(edges: Dict[Tuple[str, str, str], np.ndarray]):
    """kgwas_data.py:271 ``T.ToUndirected()`` (merge=True): bipartite -> rev_ twin appended after all
    original types; same-type -> symmetrise + coalesce (sort by (row, col), drop duplicates)."""
    out, rev = {}, {}
    for (s, rel, t), ei in edges.items():
        ei = np.asarray(ei, dtype=np.int64)
        if s != t:
            out[(s, rel, t)] = ei
            rev[(t, "rev_" + rel, s)] = ei[::-1].copy()
        else:
            out[(s, rel, t)] = coalesce_ref(np.concatenate([ei[0], ei[1]]), np.concatenate([ei[1], ei[0]]))
    out.update(rev)
    return out


def add_self_loops_ref(edges, num_nodes: Dict[str, int]):
    """kgwas_data.py:272 ``T.AddSelfLoops()``: same-type relations get arange(N) appended."""
    out = {}
    for (s, rel, t), ei in edges.items():
        if s == t:
            loop = np.arange(num_nodes[s], dtype=np.int64)
            ei = np.concatenate([ei, np.stack([loop, loop])], axis=1)
        out[(s, rel, t)] = ei
    return out


def full_neighbor_subgraph_ref(edges: Dict[Tuple[str, str, str], np.ndarray], num_nodes: Dict[str, int],
                               seed_type: str, seeds: np.ndarray, num_hops: int):
    """NeighborLoader(num_neighbors=[-1]*L) for one batch.

    Seeds come first in their type's node list; hop k expands, for every relation, each destination
    node first discovered in hop k-1 to all its in-neighbours; only traversed edges are kept; new
    nodes are appended in discovery order (relations in dict order, edges in original order).
    Returns (nodes: {type: global ids}, sub_edges: {edge_type: [2, E_sub] local ids},
             edge_ids: {edge_type: original edge positions})."""
    nodes: Dict[str, List[int]] = {t: [] for t in num_nodes}
    local: Dict[str, Dict[int, int]] = {t: {} for t in num_nodes}
    for s in np.asarray(seeds).tolist():
        if s not in local[seed_type]:
            local[seed_type][s] = len(nodes[seed_type])
            nodes[seed_type].append(s)
    frontier = {t: list(nodes[t]) for t in num_nodes}
    in_edges = {}
    for et, ei in edges.items():
        by_dst: Dict[int, List[int]] = {}
        for e, d in enumerate(np.asarray(ei[1]).tolist()):
            by_dst.setdefault(d, []).append(e)
        in_edges[et] = by_dst
    kept: Dict[Tuple[str, str, str], List[int]] = {et: [] for et in edges}
    for _ in range(num_hops):
        new_frontier = {t: [] for t in num_nodes}
        for et, ei in edges.items():
            s_t, _, d_t = et
            for d in frontier[d_t]:
                for e in in_edges[et].get(d, []):
                    kept[et].append(e)
                    s = int(ei[0][e])
                    if s not in local[s_t]:
                        local[s_t][s] = len(nodes[s_t])
                        nodes[s_t].append(s)
                        new_frontier[s_t].append(s)
        frontier = new_frontier
    sub_edges, edge_ids = {}, {}
    for et, ei in edges.items():
        s_t, _, d_t = et
        ids = np.array(kept[et], dtype=np.int64)
        edge_ids[et] = ids
        if len(ids):
            sub_edges[et] = np.stack([np.array([local[s_t][int(ei[0][e])] for e in ids], dtype=np.int64),
                                      np.array([local[d_t][int(ei[1][e])] for e in ids], dtype=np.int64)])
        else:
            sub_edges[et] = np.zeros((2, 0), dtype=np.int64)
    return {t: np.array(v, dtype=np.int64) for t, v in nodes.items()}, sub_edges, edge_ids

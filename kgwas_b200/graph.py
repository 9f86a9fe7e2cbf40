"""PyG-free heterogeneous-graph substrate for KGWAS.

``HeteroData`` duck-types the subset of ``torch_geometric.data.HeteroData`` that KGWAS touches
(kgwas/kgwas_data.py:129-273, :522-545; kgwas/kgwas.py:99-143; kgwas/utils.py:20-39, :437-458):
``data['SNP'].x``, ``data[(s, rel, t)].edge_index``, ``.node_types``, ``.edge_types``,
``.x_dict``, ``.edge_index_dict``, ``.to()``, free attributes (``data.train_mask = ...``).
``ToUndirected`` / ``AddSelfLoops`` restate the two transforms ``load_kg`` applies
(kgwas_data.py:271-272) -- integer bookkeeping, bit-exact against
oracle/bookkeeping.py.  ``make_synth_kg`` generates the ``kgwas-synth-v1`` knowledge graph
 the benchmark and the parity tests run on.

If a real ``torch_geometric`` is installed its ``HeteroData`` works with the engine as well: the
engine only reads ``x_dict`` / ``edge_index_dict`` / ``edge_types``.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

EdgeType = Tuple[str, str, str]
NODE_TYPES = ["SNP", "Gene", "CellularComponent", "BiologicalProcess", "MolecularFunction"]


class _Store:
    """Attribute bag for one node type or one edge type (``data['SNP'].x``, ``data['SNP']['n_id']``)."""

    def __init__(self, key):
        object.__setattr__(self, "_key", key)
        object.__setattr__(self, "_d", {})

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "_d")[name]
        except KeyError:
            raise AttributeError(f"{self._key!r} has no attribute {name!r}") from None

    def __setattr__(self, name, value):
        self._d[name] = value

    def __getitem__(self, name):
        return self._d[name]

    def __setitem__(self, name, value):
        self._d[name] = value

    def __contains__(self, name):
        return name in self._d

    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    @property
    def num_nodes(self):
        for k in ("x", "n_id", "y"):
            if k in self._d and torch.is_tensor(self._d[k]):
                return self._d[k].size(0)
        if "num_nodes" in self._d:
            return self._d["num_nodes"]
        raise AttributeError(f"cannot infer num_nodes of {self._key!r}")

    def is_bipartite(self):
        return isinstance(self._key, tuple) and self._key[0] != self._key[-1]

    def _apply(self, fn):
        out = _Store(self._key)
        for k, v in self._d.items():
            out._d[k] = fn(v) if torch.is_tensor(v) else v
        return out


class HeteroData:
    def __init__(self):
        object.__setattr__(self, "_nodes", {})
        object.__setattr__(self, "_edges", {})
        object.__setattr__(self, "_globals", {})

    # -- store access ---------------------------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, tuple):
            if len(key) != 3:
                raise KeyError(f"edge types are (src, rel, dst) triples, got {key!r}")
            return self._edges.setdefault(key, _Store(key))
        return self._nodes.setdefault(key, _Store(key))

    def __getattr__(self, name):
        g = object.__getattribute__(self, "_globals")
        if name in g:
            return g[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._globals[name] = value

    # -- views ----------------------------------------------------------------------------
    @property
    def node_types(self) -> List[str]:
        return list(self._nodes.keys())

    @property
    def edge_types(self) -> List[EdgeType]:
        return list(self._edges.keys())

    @property
    def node_stores(self):
        return list(self._nodes.values())

    @property
    def edge_stores(self):
        return list(self._edges.values())

    def _collect(self, stores, attr):
        return {k: s[attr] for k, s in stores.items() if attr in s}

    @property
    def x_dict(self) -> Dict[str, torch.Tensor]:
        return self._collect(self._nodes, "x")

    @property
    def edge_index_dict(self) -> Dict[EdgeType, torch.Tensor]:
        return self._collect(self._edges, "edge_index")

    def num_nodes_dict(self) -> Dict[str, int]:
        return {k: s.num_nodes for k, s in self._nodes.items()}

    @property
    def num_edges(self) -> int:
        return sum(int(s["edge_index"].size(1)) for s in self._edges.values() if "edge_index" in s)

    def to(self, device, *args, **kwargs):
        out = HeteroData()
        for k, s in self._nodes.items():
            out._nodes[k] = s._apply(lambda t: t.to(device, *args, **kwargs))
        for k, s in self._edges.items():
            out._edges[k] = s._apply(lambda t: t.to(device, *args, **kwargs))
        out._globals.update(self._globals)
        return out

    def __repr__(self):
        n = ", ".join(f"{k}={s.num_nodes}" for k, s in self._nodes.items())
        return f"HeteroData(nodes[{n}], edge_types={len(self._edges)}, edges={self.num_edges})"


# -------------------------------------------------------------------------------------------------
# transforms (kgwas/kgwas_data.py:271-272)
# -------------------------------------------------------------------------------------------------


def coalesce(edge_index: torch.Tensor) -> torch.Tensor:
    """Sort by (row, col) and drop duplicate pairs (PyG ``coalesce``)."""
    if edge_index.numel() == 0:
        return edge_index
    n = int(edge_index.max()) + 1
    key = edge_index[0] * n + edge_index[1]
    key = torch.unique(key, sorted=True)
    return torch.stack([torch.div(key, n, rounding_mode="floor"), key % n])


def to_undirected(edge_index: torch.Tensor) -> torch.Tensor:
    row, col = edge_index[0], edge_index[1]
    return coalesce(torch.stack([torch.cat([row, col]), torch.cat([col, row])]))


class ToUndirected:
    """Bipartite relation (s != t): add ``(t, 'rev_' + rel, s)`` with rows swapped, same edge order.
    Same-type relation: symmetrise + coalesce.  Reverse types come after all original types."""

    def __call__(self, data: HeteroData) -> HeteroData:
        for store in data.edge_stores:
            if "edge_index" not in store:
                continue
            src, rel, dst = store._key
            if store.is_bipartite():
                data[dst, f"rev_{rel}", src].edge_index = store.edge_index.flip([0])
            else:
                store.edge_index = to_undirected(store.edge_index)
        return data


class AddSelfLoops:
    """Same-type relations only: append ``arange(N)`` self loops after the existing edges
    (existing self loops are kept, so duplicates are possible)."""

    def __call__(self, data: HeteroData) -> HeteroData:
        for store in data.edge_stores:
            if store.is_bipartite() or "edge_index" not in store:
                continue
            n = data[store._key[0]].num_nodes
            ei = store.edge_index
            loop = torch.arange(n, dtype=ei.dtype, device=ei.device)
            store.edge_index = torch.cat([ei, loop.unsqueeze(0).repeat(2, 1)], dim=1)
        return data


# -------------------------------------------------------------------------------------------------
# kgwas-synth-v1
# -------------------------------------------------------------------------------------------------

SYNTH_NODES = {"SNP": 784_256, "Gene": 20_371, "BiologicalProcess": 17_411, "MolecularFunction": 4_563,
               "CellularComponent": 1_237}
SYNTH_RELATIONS: List[Tuple[EdgeType, int]] = [
    (("SNP", "TSS", "Gene"), 3_000_000),
    (("SNP", "PCHi-C", "Gene"), 2_000_000),
    (("SNP", "ABC", "Gene"), 1_500_000),
    (("SNP", "eQTL", "Gene"), 1_000_000),
    (("SNP", "VEP", "Gene"), 300_000),
    (("SNP", "Exon", "Gene"), 200_000),
    (("Gene", "Gene-PhysicalAssociation-Gene", "Gene"), 1_200_000),
    (("Gene", "Gene-Literature-Gene", "Gene"), 600_000),
    (("Gene", "Gene-Signaling-Gene", "Gene"), 300_000),
    (("Gene", "Gene-Reaction-Gene", "Gene"), 250_000),
    (("Gene", "Gene-DosageLethality-Gene", "Gene"), 50_000),
    (("Gene", "Gene-Associates-BiologicalProcess", "BiologicalProcess"), 140_000),
    (("Gene", "Gene-Enables-MolecularFunction", "MolecularFunction"), 60_000),
    (("Gene", "Gene-LocatedIn-CellularComponent", "CellularComponent"), 40_000),
    (("Gene", "Gene-NotContributes-MolecularFunction", "MolecularFunction"), 5_000),
    (("Gene", "Gene-NotColocalizes-CellularComponent", "CellularComponent"), 5_000),
]


def _zipf_sampler(rng: np.random.Generator, n: int, alpha: float):
    """Zipf(alpha) over a seeded permutation of n items: returns a function size -> ids."""
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** alpha
    cdf = np.cumsum(p / p.sum())
    perm = rng.permutation(n)

    def draw(size, rng_=None):
        r = np.searchsorted(cdf, (rng_ or rng).random(size), side="left")
        return perm[np.minimum(r, n - 1)]

    return draw


def make_synth_edges(scale: float = 1.0, seed: int = 42, node_scale: float | None = None, snp_block: int = 0):
    """Raw (pre-transform) ``edge_index_all`` dict + node counts of kgwas-synth-v1.
    ``scale`` shrinks edge counts, ``node_scale`` (default = scale) shrinks node counts.  ``snp_block`` > 0 draws a
    different block of SNP nodes (and SNP->Gene edges) against the SAME gene / GO graph: block b of a multi-GPU run."""
    rng = np.random.default_rng(seed)
    rng_snp = rng if snp_block == 0 else np.random.default_rng(seed + 7919 * snp_block)
    node_scale = scale if node_scale is None else node_scale
    nodes = {k: max(8, int(round(v * node_scale))) for k, v in SYNTH_NODES.items()}
    gene_draw = _zipf_sampler(rng, nodes["Gene"], 1.1)
    go_draw = {t: _zipf_sampler(rng, nodes[t], 1.0)
               for t in ("BiologicalProcess", "MolecularFunction", "CellularComponent")}
    edges: Dict[EdgeType, np.ndarray] = {}
    for (s, rel, t), e_raw in SYNTH_RELATIONS:
        e = max(1, int(round(e_raw * scale)))
        if s == "SNP":
            src = rng_snp.integers(0, nodes["SNP"], size=e)      # uniform => Poisson degrees, many isolated SNPs
            dst = gene_draw(e) if snp_block == 0 else gene_draw(e, rng_snp)
        elif t == "Gene":
            src, dst = gene_draw(e), gene_draw(e)
        else:
            src, dst = gene_draw(e), go_draw[t](e)
        edges[(s, rel, t)] = np.stack([src, dst]).astype(np.int64)
    return edges, nodes


def make_synth_kg(scale: float = 1.0, seed: int = 42, hidden: int | None = None, node_scale: float | None = None,
                  feature_dims: Dict[str, int] | None = None, snp_block: int = 0) -> HeteroData:
    """kgwas-synth-v1 as a transformed ``HeteroData`` (27 edge types at any scale).

    ``hidden`` given: every node type gets ``[N, hidden]`` N(0,1) features (the conv path starts after
    the input MLPs).  Otherwise fast-mode raw widths (SNP 20, Gene 5120, GO 128) or ``feature_dims``."""
    edges, nodes = make_synth_edges(scale, seed, node_scale, snp_block)
    g = torch.Generator().manual_seed(seed)
    g_snp = g if snp_block == 0 else torch.Generator().manual_seed(seed + 7919 * snp_block)
    data = HeteroData()
    dims = {"SNP": 20, "Gene": 5120, "CellularComponent": 128, "BiologicalProcess": 128, "MolecularFunction": 128}
    if feature_dims:
        dims.update(feature_dims)
    for t in NODE_TYPES:
        d = hidden if hidden is not None else dims[t]
        if hidden is None and t in ("CellularComponent", "BiologicalProcess", "MolecularFunction"):
            data[t].x = torch.rand((nodes[t], d), generator=g)     # kgwas_data.py:190
        else:
            data[t].x = torch.randn((nodes[t], d), generator=g_snp if t == "SNP" else g)
    for k, ei in edges.items():
        data[k].edge_index = torch.from_numpy(ei)
    data = ToUndirected()(data)
    data = AddSelfLoops()(data)
    return data

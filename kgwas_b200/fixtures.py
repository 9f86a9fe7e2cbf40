"""Synthetic on-disk data tree in KGWAS_Data's file formats -- BASELINE config 1:
fast-mode KG at a chosen scale, random SNP / gene embeddings, synthetic sum-stats with columns CHR, SNP, P, N,
BETA, SE.  Used by the plumbing tests; the reference ships no data."""
from __future__ import annotations

import os
import pickle

import numpy as np
import pandas as pd

from .graph import NODE_TYPES, make_synth_edges
from .kgwas_data import REQUIRED_FILES

_PREFIX = {"SNP": "rs", "Gene": "ENSG", "CellularComponent": "GO:CC", "BiologicalProcess": "GO:BP",
           "MolecularFunction": "GO:MF"}


def write_fixture_tree(root: str, scale: float = 0.01, seed: int = 42, n_sumstats: int = 10_000,
                       snp_dim: int = 20, gene_dim: int = 64, missing_emb_fraction: float = 0.05) -> str:
    rng = np.random.default_rng(seed)
    edges, nodes = make_synth_edges(scale=scale, seed=seed)
    idx2id = {t: {i: f"{_PREFIX[t]}{i:07d}" for i in range(nodes[t])} for t in NODE_TYPES}
    id2idx = {t: {v: k for k, v in idx2id[t].items()} for t in NODE_TYPES}

    def dump(rel, obj):
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            pickle.dump(obj, f)

    dump("cell_kg/network/node_idx2id.pkl", idx2id)
    dump("cell_kg/network/node_id2idx.pkl", id2idx)
    dump("cell_kg/network/edge_index.pkl", {k: v for k, v in edges.items()})
    keep = lambda n: rng.random(n) >= missing_emb_fraction          # some ids have no embedding -> torch.rand fallback
    dump("cell_kg/node_emb/variant_emb/enformer_feat.pkl",
         {idx2id["SNP"][i]: rng.standard_normal(snp_dim).astype(np.float32) for i in np.nonzero(keep(nodes["SNP"]))[0]})
    dump("cell_kg/node_emb/gene_emb/esm_feat.pkl",
         {idx2id["Gene"][i]: rng.standard_normal(gene_dim).astype(np.float32) for i in np.nonzero(keep(nodes["Gene"]))[0]})
    snp_ids = [idx2id["SNP"][i] for i in range(nodes["SNP"])]
    sub = rng.random(nodes["SNP"]) < 0.9                            # LD files miss some SNPs -> minimum score
    os.makedirs(os.path.join(root, "ld_score"), exist_ok=True)
    for name in ("filter_genotyped_ldscores.csv", "ldscores_from_data.csv"):
        pd.DataFrame({"SNP": np.array(snp_ids)[sub], "L2": rng.gamma(2.0, 20.0, int(sub.sum()))}).to_csv(
            os.path.join(root, "ld_score", name), index=False)
    for rel in REQUIRED_FILES:                                     # clumping inputs: empty stubs suffice for train()
        path = os.path.join(root, rel)
        if not os.path.exists(path):
            os.makedirs(os.path.dirname(path), exist_ok=True)
            open(path, "wb").close()
    n = min(n_sumstats, nodes["SNP"])
    chosen = rng.choice(nodes["SNP"], size=n, replace=False)
    p = rng.random(n)
    p[: max(2, n // 50)] = rng.uniform(1.1e-3, 9e-3, max(2, n // 50))    # keep find_closest_x's denominator non-zero
    beta, se = rng.standard_normal(n) * 0.05, rng.uniform(0.02, 0.08, n)
    df = pd.DataFrame({"CHR": rng.integers(1, 23, n), "SNP": [snp_ids[i] for i in chosen], "P": p, "N": 10000,
                       "BETA": beta, "SE": se})
    df.to_csv(os.path.join(root, "synthetic_sumstats.fastGWA"), sep="\t", index=False)
    return os.path.join(root, "synthetic_sumstats.fastGWA")

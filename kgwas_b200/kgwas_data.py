"""``KGWAS_Data`` with the reference's surface (kgwas/kgwas_data.py:19-558): same constructor, ``load_kg`` /
``load_external_gwas`` / ``load_full_gwas`` / ``load_gwas_subsample`` / ``load_simulation_gwas`` /
``process_gwas_file`` / ``prepare_split`` and the attributes ``KGWAS.train`` reads (``data``, ``id2idx``,
``idx2id``, ``*_init_dim_size``, ``lr_uni``, ``rs_id_to_ldsc_weight``, ``*_input_nodes``, ``all_ids``, ``y``),
on the PyG-free ``HeteroData`` of graph.py.  On-disk formats: SURVEY.md Appendix B."""
from __future__ import annotations

import os
import pickle

import numpy as np
import pandas as pd
import torch
from sklearn.model_selection import train_test_split

from .graph import AddSelfLoops, HeteroData, ToUndirected
from .postprocess import ldsc_regression_weights
from .utils import load_dict

REQUIRED_FILES = [
    "cell_kg/network/node_idx2id.pkl", "cell_kg/network/edge_index.pkl", "cell_kg/network/node_id2idx.pkl",
    "cell_kg/node_emb/variant_emb/enformer_feat.pkl", "cell_kg/node_emb/gene_emb/esm_feat.pkl",
    "ld_score/filter_genotyped_ldscores.csv", "ld_score/ldscores_from_data.csv",
    "ld_score/ukb_white_ld_10MB_no_hla.pkl", "ld_score/ukb_white_ld_10MB.pkl", "misc_data/ukb_white_with_cm.bim",
]
GO_TYPES = ["CellularComponent", "BiologicalProcess", "MolecularFunction"]
BINARY_TRAITS = ["body_BALDING1", "cancer_BREAST", "disease_ALLERGY_ECZEMA_DIAGNOSED",
                 "disease_HYPOTHYROIDISM_SELF_REP", "other_MORNINGPERSON", "pigment_SUNBURN"]
# (file below node_emb/, vector width) per embedding option -- kgwas_data.py:133-252
SNP_EMB = {"baselineLD": ("variant_emb/baselineld_feat.pkl", 70), "SLDSC": ("variant_emb/sldsc_feat.pkl", 165),
           "enformer": ("variant_emb/enformer_feat.pkl", 20)}
GENE_EMB = {"esm": ("gene_emb/esm_feat.pkl", 5120), "pops": ("gene_emb/pops_feat.pkl", 57742),
            "pops_expression": ("gene_emb/pops_expression_feat.pkl", 40546)}


class KGWAS_Data:
    def __init__(self, data_path="./data/"):
        self.data_path = data_path
        os.makedirs(data_path, exist_ok=True)
        missing = [f for f in REQUIRED_FILES if not os.path.exists(os.path.join(data_path, f))]
        if missing:
            # the reference downloads kgwas_core_data from Harvard Dataverse here (kgwas_data.py:41-46)
            raise FileNotFoundError(
                "KGWAS core data not found under %r (missing %d files, e.g. %s). The reference would download it from "
                "dataverse.harvard.edu; fetch it yourself or build a synthetic tree with "
                "kgwas_b200.fixtures.write_fixture_tree()." % (data_path, len(missing), missing[0]))
        print("All required data files are present.")

    # ---- knowledge graph ----------------------------------------------------------------------
    def _emb_matrix(self, node_map, table, width):
        """[N, width] rows from a {id: vector} table; ids without an entry get torch.rand (kgwas_data.py:167,183)."""
        rows = [torch.as_tensor(np.asarray(table[node_map[i]])) if node_map[i] in table else torch.rand(width)
                for i in range(len(node_map))]
        return torch.vstack(rows).float()

    def load_kg(self, snp_init_emb="enformer", go_init_emb="random", gene_init_emb="esm", sample_edges=False,
                sample_ratio=1):
        emb_dir = os.path.join(self.data_path, "cell_kg/node_emb")
        print("--loading KG---")
        idx2id = load_dict(os.path.join(self.data_path, "cell_kg/network/node_idx2id.pkl"))
        edge_index_all = load_dict(os.path.join(self.data_path, "cell_kg/network/edge_index.pkl"))
        self.id2idx = load_dict(os.path.join(self.data_path, "cell_kg/network/node_id2idx.pkl"))
        self.idx2id = idx2id
        data = HeteroData()

        def transe():
            return (load_dict(os.path.join(emb_dir, "transe_emb/transe_emb_id2idx_kg.pkl")),
                    load_dict(os.path.join(emb_dir, "transe_emb/transe_emb_inverse_triplets.pkl")))

        def kg_rows(node_map):
            id2idx_kg, kg_emb = transe()
            return torch.vstack([torch.as_tensor(kg_emb[id2idx_kg[node_map[i]]]) if node_map[i] in id2idx_kg
                                 else torch.rand(50) for i in range(len(node_map))])

        print(f"--using {snp_init_emb} SNP embedding--")
        if snp_init_emb == "random":
            data["SNP"].x, snp_dim = torch.rand((len(idx2id["SNP"]), 128)), 128
        elif snp_init_emb == "kg":
            data["SNP"].x, snp_dim = kg_rows(idx2id["SNP"]), 50
        elif snp_init_emb == "cadd":
            df = pd.read_csv(os.path.join(emb_dir, "variant_emb/cadd_feat.csv")).set_index("Unnamed: 0")
            table = dict(zip(df.index.values, df.values))
            data["SNP"].x, snp_dim = self._emb_matrix(idx2id["SNP"], table, 64), 64
        elif snp_init_emb in SNP_EMB:
            rel, snp_dim = SNP_EMB[snp_init_emb]
            data["SNP"].x = self._emb_matrix(idx2id["SNP"], load_dict(os.path.join(emb_dir, rel)), snp_dim)
        else:
            raise ValueError(f"unknown snp_init_emb {snp_init_emb!r}")

        print(f"--using {go_init_emb} go embedding--")
        if go_init_emb == "random":
            for t in GO_TYPES:
                data[t].x = torch.rand((len(idx2id[t]), 128))
            go_dim = 128
        elif go_init_emb == "kg":
            for t in GO_TYPES:
                data[t].x = kg_rows(idx2id[t])
            go_dim = 50
        elif go_init_emb == "biogpt":
            table = load_dict(os.path.join(emb_dir, "program_emb/biogpt_feat.pkl"))
            for t in GO_TYPES:
                data[t].x = self._emb_matrix(idx2id[t], table, 1600)
            go_dim = 1600
        else:
            raise ValueError(f"unknown go_init_emb {go_init_emb!r}")

        print(f"--using {gene_init_emb} gene embedding--")
        if gene_init_emb == "random":
            data["Gene"].x, gene_dim = torch.rand((len(idx2id["Gene"]), 128)), 128
        elif gene_init_emb == "kg":
            data["Gene"].x, gene_dim = kg_rows(idx2id["Gene"]), 50
        elif gene_init_emb in GENE_EMB:
            rel, gene_dim = GENE_EMB[gene_init_emb]
            data["Gene"].x = self._emb_matrix(idx2id["Gene"], load_dict(os.path.join(emb_dir, rel)), gene_dim)
        else:
            raise ValueError(f"unknown gene_init_emb {gene_init_emb!r}")

        self.gene_init_dim_size, self.go_init_dim_size, self.snp_init_dim_size = gene_dim, go_dim, snp_dim
        for et, ei in edge_index_all.items():
            edge_index = torch.as_tensor(np.asarray(ei), dtype=torch.int64)
            if sample_edges:                                       # kgwas_data.py:261-268
                n = edge_index.size(1)
                keep = torch.randperm(n)[:int(n * sample_ratio)]
                print(et, " sampling ratio ", sample_ratio, " from ", n, " to ", keep.numel())
                edge_index = edge_index[:, keep]
            data[et].edge_index = edge_index
        data = ToUndirected()(data)                                # kgwas_data.py:271
        data = AddSelfLoops()(data)                                # kgwas_data.py:272
        self.data = data

    # ---- summary statistics --------------------------------------------------------------------
    def load_external_gwas(self, path=None, seed=42, example_file=False):
        if example_file:
            path = os.path.join(self.data_path, "biochemistry_Creatinine_fastgwa_full_10000_1.fastGWA")
            if not os.path.exists(path):
                raise FileNotFoundError("example GWAS file not present (the reference downloads it from Dataverse)")
        if path is None:
            raise ValueError("A valid path must be provided or example_file must be set to True.")
        print(f"Loading GWAS file from {path}...")
        lr_uni = pd.read_csv(path, sep=None, engine="python")
        for col, msg in (("CHR", "CHR chromosome not in the file!"), ("SNP", "SNP column not in the file!"),
                         ("P", "P column not in the file!"), ("N", "N column number of sample size not in the file!")):
            if col not in lr_uni.columns.values:
                raise ValueError(msg)
        lr_uni = lr_uni.rename(columns={"CHR": "#CHROM", "SNP": "ID"})
        n_before = len(lr_uni)
        lr_uni = lr_uni[lr_uni.ID.isin(set(self.idx2id["SNP"].values()))]
        print("Number of SNPs in the KG:", len(self.idx2id["SNP"]))
        print("Number of SNPs in the GWAS:", n_before)
        print("Number of SNPs in the KG variant set:", len(lr_uni))
        self.lr_uni, self.sample_size, self.pheno, self.seed = lr_uni, lr_uni.N.values[0], "EXTERNAL", seed

    def load_full_gwas(self, pheno, seed=42):
        self.pheno, self.seed = pheno, seed
        lr_uni = pd.read_csv(os.path.join(self.data_path, "full_gwas", f"{pheno}_with_rel_fastgwa.fastGWA"), sep="\t")
        self.lr_uni = lr_uni.rename(columns={"CHR": "#CHROM", "SNP": "ID"})
        self.sample_size = 387113

    def load_gwas_subsample(self, pheno, sample_size, seed):
        self.sample_size, self.pheno, self.seed = sample_size, pheno, seed
        base = os.path.join(self.data_path, "subsample_gwas")
        if sample_size > 3000:
            lr_uni = pd.read_csv(os.path.join(base, f"{pheno}_fastgwa_full_{sample_size}_{seed}.fastGWA"), sep="\t")
            lr_uni = lr_uni.rename(columns={"CHR": "#CHROM", "SNP": "ID"})
        else:                                                      # PLINK output below 3000 samples
            suffix = "PHENO1.glm.logistic.hybrid" if pheno in BINARY_TRAITS else "PHENO1.glm.linear"
            lr_uni = pd.read_csv(os.path.join(base, f"{pheno}_plink_{sample_size}_{seed}.{suffix}"), sep="\t")
        self.lr_uni = lr_uni

    def load_simulation_gwas(self, simulation_type, seed):
        small_cohort, hits, h2 = 5000, 20000, 0.3
        base = os.path.join(self.data_path, "simulation_gwas")
        name = {"causal_link": f"causal_link_simulation/{hits}_{seed}_{h2}_graph_funct_v2_ggi.fastGWA",
                "causal": f"causal_simulation/{hits}_{seed}_{h2}_{small_cohort}_graph_funct_v2.fastGWA",
                "null": f"null_simulation/{hits}_{seed}_{h2}_{small_cohort}.fastGWA"}[simulation_type]
        lr_uni = pd.read_csv(os.path.join(base, name), sep="\t")
        ren = {"CHR": "#CHROM"} if ("SNP" in lr_uni.columns and "ID" in lr_uni.columns) else {"CHR": "#CHROM", "SNP": "ID"}
        self.lr_uni, self.sample_size, self.seed, self.pheno = lr_uni.rename(columns=ren), small_cohort, seed, "simulation"

    def process_gwas_file(self, label="chi"):
        """LDSC regression weights (normalised to mean 1) + chi-square labels (kgwas_data.py:391-520)."""
        lr_uni = self.lr_uni
        ld_tab = dict(pd.read_csv(os.path.join(self.data_path, "ld_score/filter_genotyped_ldscores.csv")).values)
        wld_tab = dict(pd.read_csv(os.path.join(self.data_path, "ld_score/ldscores_from_data.csv")).values)
        m, h_g_2 = 15000000, 0.5
        n = self.sample_size if "N" not in lr_uni.columns.values else np.mean(lr_uni.N)
        min_ld, min_wld = min(ld_tab.values()), min(wld_tab.values())
        lr_uni["ld_score"] = lr_uni.ID.map(lambda s: ld_tab.get(s, min_ld))        # missing SNPs get the minimum score
        lr_uni["w_ld_score"] = 1 + lr_uni.ID.map(lambda s: wld_tab.get(s, min_wld))  # data LD excludes the SNP itself
        print("Using ldsc weight...")
        w = ldsc_regression_weights(lr_uni["ld_score"].to_numpy(float), lr_uni["w_ld_score"].to_numpy(float), n, m, h_g_2)
        w = w / np.mean(w)
        print("ldsc_weight mean: ", np.mean(w))
        self.rs_id_to_ldsc_weight = dict(zip(lr_uni.ID.values, w))
        if label != "chi":
            raise NotImplementedError("residual-* labels need statsmodels (absent here); KGWAS's default label is 'chi'")
        if "chi" in lr_uni.columns.values:
            print("chi pre-computed...")
            lr_uni["y"] = lr_uni["chi"].values
        elif self.pheno in BINARY_TRAITS and self.sample_size <= 3000:
            lr_uni["y"] = (lr_uni["Z_STAT"].values ** 2)
        elif "BETA" in lr_uni.columns.values and "SE" in lr_uni.columns.values:
            lr_uni["y"] = (lr_uni["BETA"] / lr_uni["SE"]).values ** 2
        else:
            from scipy.stats import chi2
            lr_uni["y"] = chi2.ppf(1 - lr_uni["P"].values, 1)
        lr_uni["y"] = lr_uni.y.fillna(0)
        self.all_ids = np.array([self.id2idx["SNP"][i] for i in lr_uni.ID.values])
        self.y = lr_uni.y.values
        self.lr_uni = lr_uni

    def prepare_split(self, test_set_fraction_data=0.05):
        """5 % test, then 5 % of the rest validation, both with random_state = seed (kgwas_data.py:522-545)."""
        tv_ids, test_ids, y_tv, y_test = train_test_split(self.all_ids, self.y, test_size=test_set_fraction_data,
                                                          random_state=self.seed)
        train_ids, val_ids, y_train, y_val = train_test_split(tv_ids, y_tv, test_size=0.05, random_state=self.seed)
        self.train_input_nodes, self.val_input_nodes, self.test_input_nodes = \
            ("SNP", train_ids), ("SNP", val_ids), ("SNP", test_ids)
        y_snp = torch.zeros(self.data["SNP"].x.shape[0]) - 1           # -1 marks SNPs without a label
        y_snp[train_ids] = torch.tensor(y_train).float()
        y_snp[val_ids] = torch.tensor(y_val).float()
        y_snp[test_ids] = torch.tensor(y_test).float()
        self.data["SNP"].y = y_snp
        for t in self.data.node_types:
            self.data[t].n_id = torch.arange(self.data[t].x.shape[0])
        self.data.train_mask, self.data.val_mask, self.data.test_mask = train_ids, val_ids, test_ids
        self.data.all_mask = self.all_ids

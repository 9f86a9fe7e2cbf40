"""CPU plumbing (BASELINE config 1): synthetic on-disk tree -> KGWAS_Data -> KGWAS.train(epoch=1) with
sample_edges=True ratio=0.01-style sub-sampling, 2-layer SAGE, device='cpu'.  The host logic under test is
kgwas_b200's; the arithmetic engine is the CPU oracle, injected here because the product engine is CUDA-only."""
import os

import numpy as np
import pandas as pd
import pytest
import torch


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    from kgwas_b200.fixtures import write_fixture_tree
    root = str(tmp_path_factory.mktemp("kgwas_data"))
    gwas = write_fixture_tree(root, scale=0.004, seed=1, n_sumstats=1500, gene_dim=48)
    return root, gwas


def test_kgwas_data_surface(tree, monkeypatch):
    from kgwas_b200.kgwas_data import GENE_EMB, KGWAS_Data
    root, gwas = tree
    monkeypatch.setitem(GENE_EMB, "esm", ("gene_emb/esm_feat.pkl", 48))      # the fixture's synthetic 'esm' width
    torch.manual_seed(0)
    d = KGWAS_Data(data_path=root)
    d.load_kg(snp_init_emb="enformer", go_init_emb="random", gene_init_emb="esm", sample_edges=True, sample_ratio=0.5)
    assert (d.snp_init_dim_size, d.gene_init_dim_size, d.go_init_dim_size) == (20, 48, 128)
    assert len(d.data.edge_types) == 27 and d.data["SNP"].x.shape[1] == 20
    d.load_external_gwas(gwas, seed=42)
    assert {"#CHROM", "ID", "P", "N"} <= set(d.lr_uni.columns)
    d.process_gwas_file()
    w = np.array(list(d.rs_id_to_ldsc_weight.values()))
    assert abs(w.mean() - 1) < 1e-9 and (w > 0).all()
    assert np.allclose(d.y, (d.lr_uni.BETA / d.lr_uni.SE).values ** 2)
    d.prepare_split()
    n = len(d.all_ids)
    assert len(d.test_input_nodes[1]) == int(np.ceil(0.05 * n))
    assert (d.data["SNP"].y >= -1).all() and (d.data["SNP"].y[d.train_input_nodes[1]] >= 0).all()
    assert torch.equal(d.data["Gene"].n_id, torch.arange(d.data["Gene"].x.shape[0]))
    with pytest.raises(FileNotFoundError):
        KGWAS_Data(data_path=os.path.join(root, "nope"))


def test_train_one_epoch_cpu(tree, monkeypatch):
    from kgwas_b200.kgwas import KGWAS
    from kgwas_b200.kgwas_data import GENE_EMB, KGWAS_Data
    from oracle import kgwas_oracle as O
    root, gwas = tree
    monkeypatch.setitem(GENE_EMB, "esm", ("gene_emb/esm_feat.pkl", 48))      # small synthetic width keeps the test fast
    torch.manual_seed(0)
    d = KGWAS_Data(data_path=root)
    d.load_kg(sample_edges=True, sample_ratio=0.3)
    d.load_external_gwas(gwas)
    d.process_gwas_file()
    d.prepare_split()
    monkeypatch.setattr(KGWAS, "_model_cls", O.HeteroGNN)
    run = KGWAS(d, device="cpu", exp_name="plumb", seed=42)
    run.initialize_model(gnn_num_layers=2, gnn_hidden_dim=16, gnn_backbone="SAGE")
    before = {k: v.clone() for k, v in run.model.state_dict().items()
              if not isinstance(v, torch.nn.parameter.UninitializedParameter)}
    run.train(batch_size=32, epoch=1, save_best_model=True, save_name="plumb")
    assert os.path.exists(os.path.join(root, "model/plumb/model.pt"))
    res = pd.read_csv(os.path.join(root, "model_pred/new_experiments/plumb_pred.csv"), sep="\t")
    assert {"pred", "P_weighted", "KGWAS_P"} <= set(res.columns) and len(res) == len(d.lr_uni)
    assert res.KGWAS_P.between(0, 1).all() and np.isfinite(res.pred).all()
    after = run.model.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before)       # the optimiser moved the weights
    # checkpoint round trip through the reference's load_pretrained surface
    run2 = KGWAS(d, device="cpu", exp_name="plumb2")
    run2.load_pretrained(os.path.join(root, "model/plumb"))
    assert len(run2.kgwas_res) == len(res)
    # vectorised LDSC weight lookup == the reference's per-SNP dict lookups (kgwas.py:142-143)
    n_id = torch.from_numpy(d.train_input_nodes[1][:50])
    ref = torch.tensor([d.rs_id_to_ldsc_weight[d.idx2id["SNP"][i.item()]] for i in n_id])
    assert torch.equal(run._ld_weights(n_id), ref) and ref.dtype == torch.float64


def test_postprocess_reweighting():
    from kgwas_b200.postprocess import find_closest_x, storey_ribshirani_integrate
    rng = np.random.default_rng(0)
    n = 20000
    signal = rng.random(n) < 0.1
    p = np.where(signal, rng.beta(0.3, 4, n), rng.random(n))
    pred = np.where(signal, rng.normal(3, 1, n), rng.normal(0, 1, n))
    df = pd.DataFrame({"P": p, "abs_pred": np.abs(pred)})
    pw = storey_ribshirani_integrate(df, column="abs_pred", num_bins=50)
    assert pw.shape == (n,) and np.isfinite(pw).all() and (pw >= 0).all() and (pw <= 1).all()
    assert abs(df["weights"].mean() - 1) < 1e-9
    assert (pw[signal] < p[signal]).mean() > (pw[~signal] < p[~signal]).mean() + 0.1        # signal bins gain power
    df["P_weighted"] = pw
    s = find_closest_x(df)
    assert 0 <= s <= 200

"""GPU: the drop-in surface end to end -- mini-batch equivalence property and KGWAS.train on the CUDA engine."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import kgwas_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("backbone", ["SAGE", "GAT"])
def test_minibatch_equals_full_graph_on_gpu(cuda, backbone):
    """L layers on an L-hop full-neighbour batch == full-graph forward, and both equal
    the CPU oracle on the same weights."""
    import kgwas_b200
    from kgwas_b200 import make_synth_kg
    from kgwas_b200.loader import NeighborLoader
    h = 64
    data = make_synth_kg(scale=0.003, seed=21, hidden=h)
    for t in data.node_types:
        data[t].n_id = torch.arange(data[t].num_nodes)
    torch.manual_seed(1)
    ref = O.HeteroGNN(data, h, 1, 2, backbone, "sum", h, h, h, 1, no_relu=True)
    full_ref = ref({k: v.clone() for k, v in data.x_dict.items()}, data.edge_index_dict, data["SNP"].num_nodes)
    model = kgwas_b200.HeteroGNN(data, h, 1, 2, backbone, "sum", h, h, h, 1, no_relu=True)
    model.load_state_dict(ref.state_dict())
    model = model.to(cuda).eval()
    g = data.to(cuda)
    with torch.no_grad():
        full = model(g.x_dict, g.edge_index_dict, data["SNP"].num_nodes)
        scale = full_ref.abs().max().item()
        assert (full.cpu() - full_ref).abs().max().item() / scale < 1e-4
        seeds = np.random.default_rng(0).choice(data["SNP"].num_nodes, 100, replace=False)
        loader = NeighborLoader(data, [-1, -1], ("SNP", seeds), batch_size=40)
        outs = []
        for batch in loader:
            b = batch.to(cuda)
            outs.append(model(b.x_dict, b.edge_index_dict, b["SNP"].batch_size))
        out = torch.cat(outs)
        assert (out - full[torch.from_numpy(seeds).to(cuda)]).abs().max().item() / scale < 1e-5


def test_kgwas_train_one_epoch_on_gpu(cuda, tmp_path, monkeypatch):
    """BASELINE config 1 on the CUDA engine: fixture tree -> KGWAS_Data -> KGWAS.train(epoch=1) -> predictions equal
    the oracle's when it is given the trained weights."""
    from kgwas_b200.fixtures import write_fixture_tree
    from kgwas_b200.kgwas import KGWAS
    from kgwas_b200.kgwas_data import GENE_EMB, KGWAS_Data
    from kgwas_b200.loader import NeighborLoader
    root = str(tmp_path)
    gwas = write_fixture_tree(root, scale=0.004, seed=2, n_sumstats=1500, gene_dim=64)
    monkeypatch.setitem(GENE_EMB, "esm", ("gene_emb/esm_feat.pkl", 64))
    torch.manual_seed(0)
    d = KGWAS_Data(data_path=root)
    d.load_kg(sample_edges=True, sample_ratio=0.5)
    d.load_external_gwas(gwas)
    d.process_gwas_file()
    d.prepare_split()
    for backbone in ("SAGE", "GAT"):
        run = KGWAS(d, device="cuda:0", exp_name=f"gpu_{backbone}")
        run.initialize_model(gnn_num_layers=2, gnn_hidden_dim=32, gnn_backbone=backbone)
        run.train(batch_size=64, epoch=1, save_best_model=True)
        res = pd.read_csv(os.path.join(root, f"model_pred/new_experiments/gpu_{backbone}_pred.csv"), sep="\t")
        assert len(res) == len(d.lr_uni) and np.isfinite(res.pred).all() and res.KGWAS_P.between(0, 1).all()
        ref = O.HeteroGNN(d.data, 32, 1, 2, backbone, "sum", d.snp_init_dim_size, d.gene_init_dim_size, d.go_init_dim_size, 1)
        ref.load_state_dict(torch.load(os.path.join(root, f"model/gpu_{backbone}/model.pt"), weights_only=False))
        ids = d.test_input_nodes[1][:60]
        batch = next(iter(NeighborLoader(d.data, [-1, -1], ("SNP", ids), batch_size=60)))
        with torch.no_grad():
            p_ref = ref(batch.x_dict, batch.edge_index_dict, 60).reshape(-1)
            p_gpu = run.best_model(batch.to("cuda:0").x_dict, batch.to("cuda:0").edge_index_dict, 60).reshape(-1).cpu()
        assert (p_ref - p_gpu).abs().max().item() <= 1e-4 * max(1e-3, p_ref.abs().max().item())

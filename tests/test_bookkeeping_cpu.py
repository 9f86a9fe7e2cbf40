"""CPU: integer bookkeeping of the host side (transforms, synthetic KG, full-neighbour loader) is
bit-exact against the numpy / pure-Python oracle."""
import numpy as np
import torch

from oracle import bookkeeping as B


def _raw_graph(seed=0, scale=0.0015):
    from kgwas_b200.graph import make_synth_edges
    return make_synth_edges(scale=scale, seed=seed)


def test_transforms_bit_exact():
    from kgwas_b200 import AddSelfLoops, HeteroData, ToUndirected
    edges, nodes = _raw_graph()
    # a Gene self-loop already present in the raw data must survive and be duplicated (SURVEY App. A.8 iii)
    k = ("Gene", "Gene-Signaling-Gene", "Gene")
    edges[k] = np.concatenate([edges[k], np.array([[3, 3], [3, 3]])], axis=1)
    data = HeteroData()
    for t, n in nodes.items():
        data[t].x = torch.zeros(n, 4)
    for et, ei in edges.items():
        data[et].edge_index = torch.from_numpy(ei)
    data = AddSelfLoops()(ToUndirected()(data))
    ref = B.add_self_loops_ref(B.to_undirected_ref(edges), nodes)
    assert list(data.edge_types) == list(ref.keys())
    assert len(data.edge_types) == 27
    for et in ref:
        assert np.array_equal(data[et].edge_index.numpy(), ref[et]), et
    loops = data[k].edge_index[:, data[k].edge_index[0] == data[k].edge_index[1]]
    assert (loops[0] == 3).sum() == 2


def test_synth_kg_shape_and_determinism():
    from kgwas_b200 import make_synth_kg
    a = make_synth_kg(scale=0.002, seed=42, hidden=32)
    b = make_synth_kg(scale=0.002, seed=42, hidden=32)
    assert a.node_types == ["SNP", "Gene", "CellularComponent", "BiologicalProcess", "MolecularFunction"]
    assert len(a.edge_types) == 27
    for et in a.edge_types:
        assert torch.equal(a[et].edge_index, b[et].edge_index)
        assert a[et].edge_index.dtype == torch.int64
        assert int(a[et].edge_index[0].max()) < a[et[0]].num_nodes and int(a[et].edge_index[1].max()) < a[et[2]].num_nodes
    raw = make_synth_kg(scale=0.002, seed=42)
    assert raw["SNP"].x.size(1) == 20 and raw["Gene"].x.size(1) == 5120 and raw["MolecularFunction"].x.size(1) == 128


def test_full_neighbor_loader_bit_exact():
    from kgwas_b200 import make_synth_kg
    from kgwas_b200.loader import NeighborLoader
    data = make_synth_kg(scale=0.002, seed=7, hidden=8)
    for t in data.node_types:
        data[t].n_id = torch.arange(data[t].num_nodes)
    data["SNP"].y = torch.arange(data["SNP"].num_nodes, dtype=torch.float32)
    rng = np.random.default_rng(0)
    seeds = rng.choice(data["SNP"].num_nodes, size=70, replace=False)
    nodes = {t: data[t].num_nodes for t in data.node_types}
    edges = {et: data[et].edge_index.numpy() for et in data.edge_types}
    for hops in (1, 2, 3):
        loader = NeighborLoader(data, num_neighbors=[-1] * hops, input_nodes=("SNP", seeds), batch_size=32, drop_last=False)
        assert len(loader) == 3
        batches = list(loader)
        assert [b["SNP"].batch_size for b in batches] == [32, 32, 6]
        for bi, batch in enumerate(batches):
            s = seeds[bi * 32:(bi + 1) * 32]
            ref_nodes, ref_edges, ref_eids = B.full_neighbor_subgraph_ref(edges, nodes, "SNP", s, hops)
            for t in data.node_types:
                assert np.array_equal(batch[t].n_id.numpy(), ref_nodes[t]), (hops, t)
                assert torch.equal(batch[t].x, data[t].x[batch[t].n_id])
            assert torch.equal(batch["SNP"].y[:len(s)], torch.from_numpy(s).float())
            for et in data.edge_types:
                assert np.array_equal(batch[et].edge_index.numpy(), ref_edges[et]), (hops, et)
                assert np.array_equal(batch[et].e_id.numpy(), ref_eids[et])
    dl = NeighborLoader(data, num_neighbors=[-1, -1], input_nodes=("SNP", seeds), batch_size=32, drop_last=True)
    assert len(dl) == 2 and len(list(dl)) == 2


def test_seed_outputs_equal_full_graph_outputs_oracle():
    """L layers on an L-hop full-neighbour batch == full-graph forward (oracle, CPU)."""
    from kgwas_b200 import make_synth_kg
    from kgwas_b200.loader import NeighborLoader
    from oracle import kgwas_oracle as O
    h = 16
    data = make_synth_kg(scale=0.002, seed=9, hidden=h)
    for t in data.node_types:
        data[t].n_id = torch.arange(data[t].num_nodes)
    torch.manual_seed(0)
    for backbone in ("SAGE", "GAT"):
        model = O.HeteroGNN(data, h, 1, 2, backbone, "sum", h, h, h, 1, no_relu=True).double()
        xd = {k: v.double() for k, v in data.x_dict.items()}
        full = model(xd, data.edge_index_dict, data["SNP"].num_nodes)
        seeds = np.array([5, 17, 3, 900, 42])
        batch = next(iter(NeighborLoader(data, [-1, -1], ("SNP", seeds), batch_size=5)))
        out = model({k: v.double() for k, v in batch.x_dict.items()}, batch.edge_index_dict, 5)
        assert torch.allclose(out, full[torch.from_numpy(seeds)], rtol=1e-10, atol=1e-12)

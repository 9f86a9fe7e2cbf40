"""GPU: the full-neighbour sampler on the device (kgb_frontier_count / _expand / _add through the C ABI) is bit-identical
to the CPU sampler -- itself bit-exact against oracle/bookkeeping.full_neighbor_subgraph_ref (tests/test_bookkeeping_cpu.py)
-- in node order, kept edge ids and relabelled edges; duplicate seeds, empty frontiers, 1 to 3 hops; and a NeighborLoader
over a CUDA-resident graph yields the same batches as over the CPU copy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check_same(cpu_out, gpu_out, edge_types):
    n_c, e_c, i_c = cpu_out
    n_g, e_g, i_g = gpu_out
    for t in n_c:
        assert np.array_equal(n_c[t], n_g[t].cpu().numpy()), t
    for et in edge_types:
        assert np.array_equal(i_c[et], i_g[et].cpu().numpy()), et
        assert np.array_equal(e_c[et], e_g[et].cpu().numpy()), et


@pytest.mark.parametrize("hops", [1, 2, 3])
def test_gpu_sampler_is_bit_identical_to_the_cpu_sampler(cuda, hops):
    from kgwas_b200 import make_synth_kg
    from kgwas_b200.loader import FullNeighborSampler, GpuFullNeighborSampler
    data = make_synth_kg(scale=0.01, seed=5, hidden=32)
    cpu = FullNeighborSampler(data, hops)
    gpu = GpuFullNeighborSampler(data.to(cuda), hops)
    rng = np.random.default_rng(hops)
    n_snp = data["SNP"].num_nodes
    for seeds in (rng.integers(0, n_snp, 64), np.array([3, 3, 7, 3, 0]), np.arange(200), rng.integers(0, n_snp, 1)):
        _check_same(cpu.sample("SNP", seeds), gpu.sample("SNP", seeds), data.edge_types)
        for t in gpu.local:                                   # scratch tables are clean again
            assert int((gpu.local[t] != -1).sum()) == 0 and int((gpu.firstpos[t] != 2 ** 31 - 1).sum()) == 0
    # a seed type other than SNP, and a frontier that dies out (isolated node)
    _check_same(cpu.sample("Gene", np.array([1, 5, 1])), gpu.sample("Gene", np.array([1, 5, 1])), data.edge_types)


def test_neighbor_loader_on_a_cuda_graph_matches_the_cpu_loader(cuda):
    from kgwas_b200 import make_synth_kg
    from kgwas_b200.loader import NeighborLoader
    data = make_synth_kg(scale=0.004, seed=2, hidden=32)
    n_snp = data["SNP"].num_nodes
    data["SNP"].y = torch.arange(n_snp, dtype=torch.float32)
    ids = np.random.default_rng(0).permutation(n_snp)[:700]
    a = NeighborLoader(data, num_neighbors=[-1, -1], input_nodes=("SNP", ids), batch_size=256, drop_last=False)
    b = NeighborLoader(data.to(cuda), num_neighbors=[-1, -1], input_nodes=("SNP", ids), batch_size=256, drop_last=False)
    assert len(a) == len(b) == 3
    for ba, bb in zip(a, b):
        assert ba["SNP"].batch_size == bb["SNP"].batch_size
        for t in data.node_types:
            assert torch.equal(ba[t].n_id, bb[t].n_id.cpu())
            assert torch.equal(ba[t].x, bb[t].x.cpu())
        assert torch.equal(ba["SNP"].y, bb["SNP"].y.cpu())
        for et in data.edge_types:
            assert bb[et].edge_index.is_cuda
            assert torch.equal(ba[et].edge_index, bb[et].edge_index.cpu())
            assert torch.equal(ba[et].e_id, bb[et].e_id.cpu())

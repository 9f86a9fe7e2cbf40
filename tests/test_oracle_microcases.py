"""Hand-computable micro-cases for the PyG half of the path, written as explicit per-edge
Python loops in float64 -- independent of both the oracle's vectorised ops and the CUDA kernels.  The oracle is checked
on the CPU (``-m "not gpu"``), the CUDA modules on the GPU.  Semantics pinned here: PyG ``SAGEConv`` mean aggregation
with clamped degree, ``HeteroConv`` grouping (sum / mean over relations), ``softmax`` with the 1e-16 epsilon, duplicate
edges, isolated destinations, ``ToUndirected`` + ``AddSelfLoops`` bookkeeping, and the in-tree GATConv modes
(kgwas/conv.py:200-228: leaky-relu 0.2, temperature, sigmoid, raw)."""
import math

import pytest
import torch

from oracle import kgwas_oracle as O


def _loop_sage(x_src, x_dst, ei, Wl, bl, Wr):
    n_t, h = x_dst.shape[0], Wl.shape[0]
    agg = [[0.0] * x_src.shape[1] for _ in range(n_t)]
    deg = [0] * n_t
    for s, t in zip(ei[0].tolist(), ei[1].tolist()):
        deg[t] += 1
        for c in range(x_src.shape[1]):
            agg[t][c] += float(x_src[s, c])
    out = torch.zeros(n_t, h, dtype=torch.float64)
    for t in range(n_t):
        a = torch.tensor(agg[t], dtype=torch.float64) / max(deg[t], 1)
        out[t] = Wl.double() @ a + bl.double() + Wr.double() @ x_dst[t].double()
    return out


def _loop_gat(x_src, x_dst, ei, Ws, Wd, a_s, a_d, bias, temperature=1.0, mode="softmax"):
    Hs, Ht = x_src.double() @ Ws.double().t(), x_dst.double() @ Wd.double().t()
    als, ald = (Hs * a_s.double().view(1, -1)).sum(1), (Ht * a_d.double().view(1, -1)).sum(1)
    E, n_t = ei.shape[1], x_dst.shape[0]
    z = []
    for s, t in zip(ei[0].tolist(), ei[1].tolist()):
        u = float(als[s] + ald[t])
        z.append(u if u > 0 else 0.2 * u)
    if mode == "raw":
        alpha = list(z)
    elif mode == "sigmoid":
        alpha = [1.0 / (1.0 + math.exp(-v / temperature)) for v in z]
    else:
        alpha = [0.0] * E
        for t in range(n_t):
            idx = [e for e in range(E) if int(ei[1, e]) == t]
            if not idx:
                continue
            m = max(z[e] / temperature for e in idx)
            ex = {e: math.exp(z[e] / temperature - m) for e in idx}
            den = sum(ex.values()) + 1e-16
            for e in idx:
                alpha[e] = ex[e] / den
    out = bias.double().repeat(n_t, 1).clone()
    for e, (s, t) in enumerate(zip(ei[0].tolist(), ei[1].tolist())):
        out[t] += alpha[e] * Hs[s]
    return out, torch.tensor(alpha, dtype=torch.float64)


def _mk(seed=0, h=32):
    g = torch.Generator().manual_seed(seed)
    x_src, x_dst = torch.randn(5, h, generator=g), torch.randn(4, h, generator=g)
    # destination 3 has no in-edge (i); edge (1 -> 0) appears twice (ii)
    ei = torch.tensor([[1, 1, 0, 2, 4, 3, 2], [0, 0, 0, 1, 1, 2, 2]])
    return g, x_src, x_dst, ei, h


def _sage_pair(make, dev):
    g, x_src, x_dst, ei, h = _mk()
    conv = make((h, h), h)
    conv.lin_l.materialize(h) if hasattr(conv.lin_l, "materialize") else None
    conv.lin_r.materialize(h) if hasattr(conv.lin_r, "materialize") else None
    with torch.no_grad():
        conv.lin_l.weight.copy_(torch.randn(h, h, generator=g))
        conv.lin_l.bias.copy_(torch.randn(h, generator=g))
        conv.lin_r.weight.copy_(torch.randn(h, h, generator=g))
    want = _loop_sage(x_src, x_dst, ei, conv.lin_l.weight.detach().cpu(), conv.lin_l.bias.detach().cpu(),
                      conv.lin_r.weight.detach().cpu())
    conv = conv.to(dev)
    got = conv((x_src.to(dev), x_dst.to(dev)), ei.to(dev)).detach().cpu().double()
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    # (i) isolated destination: b_l + W_r x_t exactly
    iso = conv.lin_l.bias.detach().cpu().double() + conv.lin_r.weight.detach().cpu().double() @ x_dst[3].double()
    assert torch.allclose(got[3], iso, rtol=1e-5, atol=1e-5)


def _gat_pair(make, dev):
    g, x_src, x_dst, ei, h = _mk(1)
    for kw, mode, raw in (({}, "softmax", None), ({"temperature": 0.5}, "softmax", None),
                          ({"sigmoid_gat": True, "temperature": 2.0}, "sigmoid", None), ({}, "raw", True)):
        conv = make((h, h), h, heads=1, add_self_loops=False, **kw)
        for lin in (conv.lin_src, conv.lin_dst):
            lin.materialize(h) if hasattr(lin, "materialize") else None
        with torch.no_grad():
            conv.lin_src.weight.copy_(torch.randn(h, h, generator=g))
            conv.lin_dst.weight.copy_(torch.randn(h, h, generator=g))
            conv.att_src.copy_(torch.randn(1, 1, h, generator=g))
            conv.att_dst.copy_(torch.randn(1, 1, h, generator=g))
            conv.bias.copy_(torch.randn(h, generator=g))
        want, alpha = _loop_gat(x_src, x_dst, ei, conv.lin_src.weight.detach(), conv.lin_dst.weight.detach(),
                                conv.att_src.detach().view(-1), conv.att_dst.detach().view(-1), conv.bias.detach(),
                                float(kw.get("temperature", 1.0)), mode)
        conv = conv.to(dev)
        out, (_, a) = conv((x_src.to(dev), x_dst.to(dev)), ei.to(dev), return_attention_weights=True,
                           return_raw_attention_weights=raw)
        assert torch.allclose(a.detach().cpu().double().view(-1), alpha, rtol=1e-5, atol=1e-6), (kw, mode)
        assert torch.allclose(out.detach().cpu().double(), want, rtol=1e-4, atol=1e-5), (kw, mode)
        if mode == "softmax":        # (i) isolated destination -> bias; (ii) the duplicate edge has two equal weights
            assert torch.allclose(out[3].detach().cpu(), conv.bias.detach().cpu(), atol=1e-6)
            assert abs(float(a[0]) - float(a[1])) < 1e-7
    # (v) single-tensor path: H_t = H_s, lin_dst is never used (conv.py:136-138)
    conv = make(h, h, heads=1, add_self_loops=False)
    with torch.no_grad():
        conv.lin_src.weight.copy_(torch.randn(h, h, generator=g))
        conv.att_src.copy_(torch.randn(1, 1, h, generator=g))
        conv.att_dst.copy_(torch.randn(1, 1, h, generator=g))
        conv.bias.copy_(torch.randn(h, generator=g))
    ei1 = torch.tensor([[0, 1, 2, 3, 3, 4], [1, 1, 0, 0, 3, 2]])
    want, _ = _loop_gat(x_src, x_src, ei1, conv.lin_src.weight.detach(), conv.lin_src.weight.detach(),
                        conv.att_src.detach().view(-1), conv.att_dst.detach().view(-1), conv.bias.detach())
    conv = conv.to(dev)
    got = conv(x_src.to(dev), ei1.to(dev)).detach().cpu().double()
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)


def _hetero_pair(sage_cls, hetero_cls, dev):
    """(iv) one destination type fed by three relations, aggr = sum vs mean (PyG ``group``: stack + reduce)."""
    g = torch.Generator().manual_seed(2)
    h = 32
    x = {"A": torch.randn(5, h, generator=g), "B": torch.randn(4, h, generator=g)}
    eis = {("A", "r1", "B"): torch.tensor([[0, 1, 2], [0, 0, 1]]), ("A", "r2", "B"): torch.tensor([[3, 4], [1, 2]]),
           ("B", "r3", "B"): torch.tensor([[0, 1, 2, 3], [1, 2, 3, 3]])}
    for aggr in ("sum", "mean"):
        convs = {et: sage_cls((h, h), h) for et in eis}
        for c in convs.values():
            for lin in (c.lin_l, c.lin_r):
                lin.materialize(h) if hasattr(lin, "materialize") else None
            with torch.no_grad():
                c.lin_l.weight.copy_(torch.randn(h, h, generator=g))
                c.lin_l.bias.copy_(torch.randn(h, generator=g))
                c.lin_r.weight.copy_(torch.randn(h, h, generator=g))
        parts = [_loop_sage(x[et[0]], x[et[2]], ei, c.lin_l.weight.detach(), c.lin_l.bias.detach(), c.lin_r.weight.detach())
                 for (et, ei), c in zip(eis.items(), convs.values())]
        want = torch.stack(parts).sum(0) if aggr == "sum" else torch.stack(parts).mean(0)
        layer = hetero_cls(convs, aggr=aggr).to(dev)
        out = layer({k: v.to(dev) for k, v in x.items()}, {k: v.to(dev) for k, v in eis.items()})
        assert set(out.keys()) == {"B"}                   # 'A' is never a destination: it disappears (Appendix A.1)
        assert torch.allclose(out["B"].detach().cpu().double(), want, rtol=1e-4, atol=1e-5), aggr


def test_graph_transforms_microcase():
    """(iii) a gene self-loop in the raw data appears twice after ToUndirected + AddSelfLoops; bipartite relations get
    a 'rev_' twin with rows swapped in the same order (kgwas_data.py:259-272, Appendix A.5).  Pure integer work."""
    from kgwas_b200.graph import AddSelfLoops, HeteroData, ToUndirected
    from oracle import bookkeeping as B
    d = HeteroData()
    d["SNP"].x, d["Gene"].x = torch.zeros(3, 1), torch.zeros(3, 1)
    d["Gene", "g2g", "Gene"].edge_index = torch.tensor([[0, 1, 1], [1, 1, 2]])       # (1 -> 1) is a raw self-loop
    d["SNP", "s2g", "Gene"].edge_index = torch.tensor([[2, 0, 2], [0, 1, 0]])        # duplicate (2 -> 0) kept
    d = AddSelfLoops()(ToUndirected()(d))
    gg = d["Gene", "g2g", "Gene"].edge_index.tolist()
    assert gg == [[0, 1, 1, 1, 2, 0, 1, 2], [1, 0, 1, 2, 1, 0, 1, 2]]               # coalesced, then arange(3) appended
    assert list(zip(*gg)).count((1, 1)) == 2
    assert d["SNP", "s2g", "Gene"].edge_index.tolist() == [[2, 0, 2], [0, 1, 0]]
    assert d["Gene", "rev_s2g", "SNP"].edge_index.tolist() == [[0, 1, 0], [2, 0, 2]]
    assert d.edge_types == [("Gene", "g2g", "Gene"), ("SNP", "s2g", "Gene"), ("Gene", "rev_s2g", "SNP")]
    import numpy as np
    raw = {("Gene", "g2g", "Gene"): np.array([[0, 1, 1], [1, 1, 2]]), ("SNP", "s2g", "Gene"): np.array([[2, 0, 2], [0, 1, 0]])}
    ref = B.add_self_loops_ref(B.to_undirected_ref(raw), {"SNP": 3, "Gene": 3})
    assert {k: np.asarray(v).tolist() for k, v in ref.items()} == {k: d[k].edge_index.tolist() for k in d.edge_types}


def test_oracle_sage_microcases():
    _sage_pair(lambda ch, h: O.SAGEConv(ch, h), "cpu")


def test_oracle_gat_microcases():
    _gat_pair(lambda ch, h, **kw: O.GATConv(ch, h, **kw), "cpu")


def test_oracle_heteroconv_microcases():
    _hetero_pair(lambda ch, h: O.SAGEConv(ch, h), O.HeteroConv, "cpu")


@pytest.mark.gpu
def test_cuda_sage_microcases(cuda):
    import kgwas_b200
    _sage_pair(lambda ch, h: kgwas_b200.SAGEConv(ch, h), cuda)


@pytest.mark.gpu
def test_cuda_gat_microcases(cuda):
    import kgwas_b200
    _gat_pair(lambda ch, h, **kw: kgwas_b200.GATConv(ch, h, **kw), cuda)


@pytest.mark.gpu
def test_cuda_heteroconv_microcases(cuda):
    import kgwas_b200
    _hetero_pair(lambda ch, h: kgwas_b200.SAGEConv(ch, h), kgwas_b200.HeteroConv, cuda)

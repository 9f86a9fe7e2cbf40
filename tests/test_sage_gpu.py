"""GPU parity: drop-in SAGE modules vs the CPU oracle (fp32 and fp64) on seeded synthetic KGs.
Tolerance: north_star asks for per-SNP logits within 1e-4 relative (fp32)."""
import copy

import numpy as np
import pytest
import torch

from oracle import kgwas_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-12)).item()


def _models(data, h, backbone, aggr, layers=2, seed=0, no_relu=False):
    import kgwas_b200
    torch.manual_seed(seed)
    ref = O.HeteroGNN(data, h, 1, layers, backbone, aggr, h, h, h, 1, no_relu=no_relu)
    # a first forward materialises the lazy (-1, -1) weights
    ref({k: v.clone() for k, v in data.x_dict.items()}, data.edge_index_dict, 4)
    ours = kgwas_b200.HeteroGNN(data, h, 1, layers, backbone, aggr, h, h, h, 1, no_relu=no_relu)
    ours.load_state_dict(ref.state_dict())
    return ref, ours


@pytest.mark.parametrize("h,scale,aggr", [(32, 0.002, "sum"), (64, 0.003, "mean"), (128, 0.004, "sum"), (256, 0.002, "sum")])
def test_hetero_sage_forward_backward(cuda, h, scale, aggr):
    from kgwas_b200 import make_synth_kg
    data = make_synth_kg(scale=scale, seed=3, hidden=h)
    ref, ours = _models(data, h, "SAGE", aggr, no_relu=True)
    ref64 = copy.deepcopy(ref).double()
    ours = ours.to(cuda)
    gdata = data.to(cuda)
    bs = 200
    w = torch.rand(bs, dtype=torch.float64)
    yt = torch.randn(bs)

    def loss_of(model, x_dict, ei, dev, dt):
        out = model(x_dict, ei, bs).reshape(-1)
        return out, torch.mean(w.to(dev) * (out - yt.to(dev, dt)) ** 2)       # kgwas.py:145

    out_r, loss_r = loss_of(ref, {k: v.clone() for k, v in data.x_dict.items()}, data.edge_index_dict, "cpu", torch.float32)
    out_64, loss_64 = loss_of(ref64, {k: v.double() for k, v in data.x_dict.items()}, data.edge_index_dict, "cpu", torch.float64)
    out_g, loss_g = loss_of(ours, gdata.x_dict, gdata.edge_index_dict, cuda, torch.float32)
    loss_r.backward(); loss_64.backward(); loss_g.backward()
    torch.cuda.synchronize()
    assert out_g.shape == out_r.shape
    e_ours, e_ref = _rel_err(out_g, out_64), _rel_err(out_r, out_64)
    assert _rel_err(out_g, out_r) < RTOL, (e_ours, e_ref)
    assert e_ours < RTOL
    p_ref, p_64, p_g = dict(ref.named_parameters()), dict(ref64.named_parameters()), dict(ours.named_parameters())
    assert p_ref.keys() == p_g.keys()
    worst = 0.0
    floor = 1e-5 * max(p.grad.abs().max().item() for p in p_64.values() if p.grad is not None)
    for k in p_ref:
        if p_ref[k].grad is None:
            assert p_g[k].grad is None, f"{k}: reference leaves grad None (unused relation), ours does not"
            continue
        assert p_g[k].grad is not None, k
        scale_ = p_64[k].grad.abs().max().item()
        if scale_ == 0:
            assert p_g[k].grad.abs().max().item() == 0
            continue
        diff = (p_g[k].grad.double().cpu() - p_64[k].grad).abs().max().item()
        assert diff <= 1e-3 * scale_ + floor, (k, diff, scale_)


def test_sage_microcases(cuda):
    """Hand-computed anchors: isolated destination, duplicate edge."""
    import kgwas_b200
    h = 32
    conv = kgwas_b200.SAGEConv((h, h), h).to(cuda)
    x_src = torch.randn(3, h, device=cuda)
    x_dst = torch.randn(3, h, device=cuda)
    ei = torch.tensor([[0, 0, 1], [1, 1, 1]], device=cuda)        # dst 1 gets src0 twice + src1; dst 0, 2 isolated
    out = conv((x_src, x_dst), ei)
    Wl, bl, Wr = conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight
    exp = x_dst @ Wr.T + bl
    exp[1] += ((2 * x_src[0] + x_src[1]) / 3) @ Wl.T
    assert torch.allclose(out, exp, rtol=1e-5, atol=1e-5)
    # single-tensor form == tuple form with x_dst = x_src
    out_same = conv(x_src, ei)
    assert torch.allclose(out_same, conv((x_src, x_src), ei), rtol=0, atol=0)


def test_hetero_conv_min_max_and_relations(cuda):
    """dst type fed by 3 relations, aggr in {sum, mean, max, min} vs the oracle's HeteroConv."""
    import kgwas_b200
    h = 32
    torch.manual_seed(5)
    ets = [("a", "r1", "b"), ("a", "r2", "b"), ("b", "r3", "b"), ("b", "r4", "a")]
    n = {"a": 50, "b": 30}
    ei = {et: torch.stack([torch.randint(0, n[et[0]], (200,)), torch.randint(0, n[et[2]], (200,))]) for et in ets}
    ei[("b", "r4", "a")] = torch.zeros((2, 0), dtype=torch.int64)        # an empty relation still runs
    x = {k: torch.randn(v, h) for k, v in n.items()}
    for aggr in ("sum", "mean", "max", "min"):
        ref = O.HeteroConv({et: O.SAGEConv((h, h), h) for et in ets}, aggr=aggr)
        ours = kgwas_b200.HeteroConv({et: kgwas_b200.SAGEConv((h, h), h) for et in ets}, aggr=aggr)
        ours.load_state_dict(ref.state_dict())
        ours = ours.to(cuda)
        out_r = ref(x, ei)
        out_g = ours({k: v.to(cuda) for k, v in x.items()}, {k: v.to(cuda) for k, v in ei.items()})
        assert out_r.keys() == out_g.keys()
        for k in out_r:
            assert _rel_err(out_g[k], out_r[k]) < RTOL, (aggr, k)


def test_state_dict_roundtrip_and_pyg24_keys(cuda):
    import kgwas_b200
    from kgwas_b200 import make_synth_kg
    data = make_synth_kg(scale=0.001, seed=3, hidden=32)
    m = kgwas_b200.HeteroGNN(data, 32, 1, 2, "SAGE", "sum", 32, 32, 32, 1)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=5e-4)    # created BEFORE the first forward (kgwas.py:116)
    m = m.to(cuda)
    g = data.to(cuda)
    before = copy.deepcopy(m)                                               # deepcopy with lazy params (kgwas.py:124)
    out = m(g.x_dict, g.edge_index_dict, 16)
    out.sum().backward()
    opt.step()
    sd = m.state_dict()
    renamed = {}
    for k, v in sd.items():
        if ".convs." in k:
            pre, rest = k.split(".convs.", 1)
            name, tail = rest.split(".", 1) if not rest.startswith("<") else (rest, "")
            et = name.split("__")
            renamed[f"{pre}.convs.<{'___'.join(et)}>.{tail}"] = v
        else:
            renamed[k] = v
    m2 = kgwas_b200.HeteroGNN(data, 32, 1, 2, "SAGE", "sum", 32, 32, 32, 1).to(cuda)
    m2.load_state_dict(renamed)
    assert torch.equal(m2(g.x_dict, g.edge_index_dict, 16), m(g.x_dict, g.edge_index_dict, 16))
    # optimiser really updated lazily-materialised conv weights
    first = m.convs[0].convs["Gene__rev_TSS__SNP"].lin_l.weight
    assert first.grad is not None and opt.state[first]["step"] >= 1


def test_graphed_step_matches_eager(cuda):
    """The CUDA-graph capture of a full-graph training step (kgwas_b200.graphed) replays the same kernels in the same
    order: after 3 optimiser steps the parameters equal those of 3 eager steps."""
    import kgwas_b200
    from kgwas_b200 import make_synth_kg
    from kgwas_b200.graphed import GraphedStep
    h = 128
    data = make_synth_kg(scale=0.004, seed=5, hidden=h)
    n_snp = data["SNP"].x.size(0)
    gdata = data.to(cuda)
    yt = torch.randn(n_snp, device=cuda)
    w = torch.rand(n_snp, device=cuda, dtype=torch.float64)
    results = []
    for use_graph in (False, True):
        torch.manual_seed(11)
        model = kgwas_b200.HeteroGNN(data, h, 1, 2, "SAGE", "sum", h, h, h, 1)
        model = model.to(cuda)
        x = {k: v.clone().requires_grad_() for k, v in gdata.x_dict.items()}
        with torch.no_grad():                                    # materialise the lazy weights identically in both runs
            torch.manual_seed(12)
            model.forward_from_hidden({k: v.detach() for k, v in x.items()}, gdata.edge_index_dict, 4)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4, capturable=True)
        state0 = {k: v.clone() for k, v in model.state_dict().items()}

        def step(xd):
            opt.zero_grad(set_to_none=True)
            for v in xd.values():
                v.grad = None
            pred = model.forward_from_hidden(xd, gdata.edge_index_dict, n_snp).reshape(-1)
            loss = torch.mean(w * (pred - yt) ** 2)
            loss.backward()
            opt.step()
            return pred, loss

        if use_graph:
            gs = GraphedStep(step, x, warmup=2)
            # the warm-up and the capture pass moved the parameters / optimiser state: rewind both
            model.load_state_dict(state0)
            for st in opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
            assert gs.kernels_per_replay > 20
            for _ in range(3):
                pred, loss = gs()
        else:
            for _ in range(3):
                pred, loss = step(x)
        torch.cuda.synchronize()
        results.append((pred.detach().clone(), loss.detach().clone(), {k: v.clone() for k, v in model.state_dict().items()}))
    (p0, l0, s0), (p1, l1, s1) = results
    assert torch.allclose(p0, p1, rtol=1e-5, atol=1e-6 * p0.abs().max().item())
    assert torch.allclose(l0, l1, rtol=1e-6)
    for k in s0:
        assert torch.allclose(s0[k], s1[k], rtol=1e-5, atol=1e-7), k


@pytest.mark.parametrize("backbone", ["SAGE", "GAT"])
def test_a_node_type_without_rows_in_the_batch(cuda, backbone):
    """A mini-batch may lack a node type entirely (edge-sampled or tiny KGs): zero-row operands must flow through
    forward AND backward (empty contractions write zeros) and still match the oracle."""
    import kgwas_b200
    from oracle import kgwas_oracle as O
    h = 64
    g = torch.Generator().manual_seed(0)
    n = {"SNP": 40, "Gene": 9, "CellularComponent": 0}
    ei = {("SNP", "a", "Gene"): torch.stack([torch.randint(0, 40, (120,), generator=g), torch.randint(0, 9, (120,), generator=g)]),
          ("Gene", "rev_a", "SNP"): torch.stack([torch.randint(0, 9, (120,), generator=g), torch.randint(0, 40, (120,), generator=g)]),
          ("Gene", "g", "Gene"): torch.stack([torch.randint(0, 9, (30,), generator=g), torch.randint(0, 9, (30,), generator=g)]),
          ("Gene", "c", "CellularComponent"): torch.zeros((2, 0), dtype=torch.int64),
          ("CellularComponent", "rev_c", "Gene"): torch.zeros((2, 0), dtype=torch.int64)}
    x = {t: torch.randn(c, h, generator=g) for t, c in n.items()}
    mk = (lambda m: m.SAGEConv((-1, -1), h)) if backbone == "SAGE" else (lambda m: m.GATConv((-1, -1), h, heads=1, add_self_loops=False))
    torch.manual_seed(1)
    ref = O.HeteroConv({et: mk(O) for et in ei}, aggr="sum")
    xr = {k: v.clone().requires_grad_() for k, v in x.items()}
    out_r = ref(xr, ei)
    sum(v.relu().sum() for v in out_r.values()).backward()
    ours = kgwas_b200.HeteroConv({et: mk(kgwas_b200) for et in ei}, aggr="sum")
    missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=False)
    assert not unexpected
    ours = ours.to(cuda)
    xc = {k: v.to(cuda).requires_grad_() for k, v in x.items()}
    out = ours(xc, {k: v.to(cuda) for k, v in ei.items()})
    sum(v.relu().sum() for v in out.values()).backward()
    assert set(out.keys()) == set(out_r.keys()) and out["CellularComponent"].shape == (0, h)
    for t in ("SNP", "Gene"):
        assert torch.allclose(out[t].detach().cpu(), out_r[t].detach(), rtol=1e-4, atol=1e-5), t
        assert torch.allclose(xc[t].grad.cpu(), xr[t].grad, rtol=1e-3, atol=1e-5), t


@pytest.mark.parametrize("backbone", ["SAGE", "GAT"])
def test_model_on_a_non_current_device(cuda, backbone):
    """``KGWAS(data, device='cuda:1')`` without ``torch.cuda.set_device`` (kgwas/kgwas.py:38-39): the autograd Functions
    enter the device of their inputs, so launches, streams and allocations follow the data; a direct C-ABI call with a
    tensor of another device raises instead of launching on the wrong GPU.  Needs two visible GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import kgwas_b200
    from kgwas_b200 import _lib, make_synth_kg
    h = 128
    data = make_synth_kg(scale=0.003, seed=4, hidden=h)
    torch.manual_seed(0)
    m0 = kgwas_b200.HeteroGNN(data, h, 1, 2, backbone, "sum", h, h, h, 1, no_relu=True)
    d0 = data.to("cuda:0")
    m0 = m0.to("cuda:0")
    out0 = m0.forward_from_hidden({k: v.clone().requires_grad_() for k, v in d0.x_dict.items()}, d0.edge_index_dict, 50)
    out0.sum().backward()
    kgwas_b200.plan.clear_plan_cache()
    assert torch.cuda.current_device() == 0
    m1 = kgwas_b200.HeteroGNN(data, h, 1, 2, backbone, "sum", h, h, h, 1, no_relu=True)
    m1.load_state_dict(m0.state_dict())
    m1 = m1.to("cuda:1")
    d1 = data.to("cuda:1")
    out1 = m1.forward_from_hidden({k: v.clone().requires_grad_() for k, v in d1.x_dict.items()}, d1.edge_index_dict, 50)
    out1.sum().backward()
    torch.cuda.synchronize("cuda:1")
    assert torch.cuda.current_device() == 0
    assert torch.allclose(out1.cpu(), out0.cpu(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(m1.lin.weight.grad.cpu(), m0.lin.weight.grad.cpu(), rtol=1e-4, atol=1e-6)
    x1 = torch.randn(64, h, device="cuda:1")
    with pytest.raises(_lib.KgbError):
        _lib.rowdot(x1, torch.randn(1, h, device="cuda:1"), torch.empty(64, 1, device="cuda:1"), h, 1, 0)
    kgwas_b200.plan.clear_plan_cache()

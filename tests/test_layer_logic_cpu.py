"""CPU: host logic of the fused hetero-SAGE layer (plan.py / ops.py / conv.py / model.py) against the oracle, with the
CUDA kernels replaced by the plain-torch stand-ins of tests/_cpu_kernels.py.  The GPU suite checks the same thing
through the real kernels; this one keeps the relation merging, accumulate / ReLU placement, head fusion, gradient
wiring and the scheduler's reordering covered where no GPU is available."""
import os
import sys

import pytest
import torch

from oracle import kgwas_oracle as O

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _cpu_kernels  # noqa: E402  (tests/_cpu_kernels.py)


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-12)).item()


def _pair(data, h, aggr, no_relu):
    import kgwas_b200
    torch.manual_seed(0)
    ref = O.HeteroGNN(data, h, 1, 2, "SAGE", aggr, h, h, h, 1, no_relu=no_relu)
    ref({k: v.clone() for k, v in data.x_dict.items()}, data.edge_index_dict, 4)      # materialise lazy weights
    ours = kgwas_b200.HeteroGNN(data, h, 1, 2, "SAGE", aggr, h, h, h, 1, no_relu=no_relu)
    ours.load_state_dict(ref.state_dict())
    return ref, ours


@pytest.mark.parametrize("issue_order", [False, True])
@pytest.mark.parametrize("h,aggr", [(32, "mean"), (128, "sum")])
def test_fused_layer_host_logic_matches_oracle(monkeypatch, h, aggr, issue_order):
    from kgwas_b200 import make_synth_kg, ops
    _cpu_kernels.install(monkeypatch)
    monkeypatch.setattr(ops, "ISSUE_ORDER_ALWAYS", issue_order)
    data = make_synth_kg(scale=0.002, seed=3, hidden=h)
    ref, ours = _pair(data, h, aggr, no_relu=True)
    bs = 150
    w = torch.rand(bs, dtype=torch.float64)
    yt = torch.randn(bs)

    def run(model):
        x = {k: v.clone().requires_grad_() for k, v in data.x_dict.items()}
        out = model(x, data.edge_index_dict, bs).reshape(-1)
        torch.mean(w * (out - yt) ** 2).backward()
        return out, x

    out_r, x_r = run(ref)
    out_o, x_o = run(ours)
    assert out_o.shape == out_r.shape and _rel(out_o, out_r) < 1e-4
    p_r, p_o = dict(ref.named_parameters()), dict(ours.named_parameters())
    assert p_r.keys() == p_o.keys()
    scale = max(p.grad.abs().max().item() for p in p_r.values() if p.grad is not None)
    for k in p_r:
        assert (p_r[k].grad is None) == (p_o[k].grad is None), k          # unused last-layer relations: None, not zeros
        if p_r[k].grad is not None:
            assert (p_r[k].grad - p_o[k].grad).abs().max().item() <= 2e-4 * scale, k
    for t in x_r:
        if x_r[t].grad is None:
            assert x_o[t].grad is None or x_o[t].grad.abs().max() == 0
        else:
            assert _rel(x_o[t].grad, x_r[t].grad) < 2e-4, t


def test_fused_head_epilogue_host_logic(monkeypatch):
    """conv(..., _head=('SNP', w)): the extra output equals relu(out['SNP']) . w^T and its gradient reaches w, the
    layer parameters and the inputs exactly as the un-fused head does."""
    import kgwas_b200
    from kgwas_b200 import make_synth_kg
    _cpu_kernels.install(monkeypatch)
    h = 128
    data = make_synth_kg(scale=0.002, seed=4, hidden=h)
    torch.manual_seed(1)
    model = kgwas_b200.HeteroGNN(data, h, 1, 1, "SAGE", "sum", h, h, h, 1)
    conv = model.convs[0]
    w_head = torch.randn(1, h, requires_grad=True)
    res = {}
    for fused in (True, False):
        x = {k: v.clone().requires_grad_() for k, v in data.x_dict.items()}
        out = conv(x, data.edge_index_dict, _fuse_relu=True, _head=("SNP", w_head) if fused else None)
        logits = out.pop(("head", "SNP")) if fused else out["SNP"] @ w_head.T
        assert set(out) == {"SNP", "Gene", "CellularComponent", "BiologicalProcess", "MolecularFunction"}
        model.zero_grad(set_to_none=True)
        w_head.grad = None
        (logits.sum() + 0.5 * out["Gene"].sum()).backward()
        res[fused] = (logits.detach().clone(), w_head.grad.clone(), x["Gene"].grad.clone(),
                      {k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()})
    assert _rel(res[True][0], res[False][0]) < 1e-5
    assert _rel(res[True][1], res[False][1]) < 1e-5
    assert _rel(res[True][2], res[False][2]) < 1e-5
    for k, g in res[False][3].items():
        assert (g is None) == (res[True][3][k] is None), k
        if g is not None and g.abs().max() > 0:
            assert _rel(res[True][3][k], g) < 1e-4, k


@pytest.mark.parametrize("h,aggr", [(32, "sum"), (64, "mean")])
def test_fused_gat_layer_host_logic_matches_oracle(monkeypatch, h, aggr):
    """Same for the hetero-GAT layer (gat.py): folded attention logits, transform-first / aggregate-first jobs,
    softmax-group bookkeeping, parameter-gradient folding -- against the oracle, kernels stubbed."""
    import kgwas_b200
    from kgwas_b200 import make_synth_kg
    _cpu_kernels.install(monkeypatch)
    data = make_synth_kg(scale=0.002, seed=5, hidden=h)
    torch.manual_seed(0)
    ref = O.HeteroGNN(data, h, 1, 2, "GAT", aggr, h, h, h, 1, no_relu=True)
    ref({k: v.clone() for k, v in data.x_dict.items()}, data.edge_index_dict, 4)
    ours = kgwas_b200.HeteroGNN(data, h, 1, 2, "GAT", aggr, h, h, h, 1, no_relu=True)
    ours.load_state_dict(ref.state_dict())
    bs = 150
    w = torch.rand(bs, dtype=torch.float64)
    yt = torch.randn(bs)

    def run(model):
        out = model({k: v.clone() for k, v in data.x_dict.items()}, data.edge_index_dict, bs).reshape(-1)
        torch.mean(w * (out - yt) ** 2).backward()
        return out

    out_r, out_o = run(ref), run(ours)
    assert _rel(out_o, out_r) < 1e-4
    p_r, p_o = dict(ref.named_parameters()), dict(ours.named_parameters())
    assert p_r.keys() == p_o.keys()
    lazy = torch.nn.parameter.UninitializedParameter
    scale = max(p.grad.abs().max().item() for p in p_r.values() if not isinstance(p, lazy) and p.grad is not None)
    for k in p_r:
        assert isinstance(p_r[k], lazy) == isinstance(p_o[k], lazy), k
        if isinstance(p_r[k], lazy):
            continue
        assert (p_r[k].grad is None) == (p_o[k].grad is None), k
        if p_r[k].grad is not None:
            assert (p_r[k].grad - p_o[k].grad).abs().max().item() <= 1e-3 * scale, k


@pytest.mark.parametrize("bs", [90, None])
def test_model_with_fused_head_matches_oracle(monkeypatch, bs):
    """HeteroGNN.forward with the head evaluated inside the last layer (the GPU configuration), final ReLU on:
    logits, head and layer gradients against the oracle; batch_size < N and == N."""
    import kgwas_b200
    from kgwas_b200 import make_synth_kg, model as kmodel
    _cpu_kernels.install(monkeypatch)
    monkeypatch.setattr(kmodel, "_head_is_fusable", lambda w: w.size(0) == 1)
    h = 128
    data = make_synth_kg(scale=0.002, seed=8, hidden=h)
    n_snp = data["SNP"].x.size(0)
    bs = n_snp if bs is None else bs
    ref, ours = _pair(data, h, "sum", no_relu=False)
    with torch.no_grad():                      # a positive head bias keeps the final ReLU from zeroing every logit
        ref.lin.bias.fill_(0.5)
        ours.lin.bias.fill_(0.5)
    w = torch.rand(bs, dtype=torch.float64)
    yt = torch.randn(bs)
    outs = []
    for model in (ref, ours):
        out = model({k: v.clone() for k, v in data.x_dict.items()}, data.edge_index_dict, bs).reshape(-1)
        torch.mean(w * (out - yt) ** 2).backward()
        outs.append(out)
    assert float(outs[0].detach().abs().max()) > 0
    assert _rel(outs[1], outs[0]) < 1e-4
    p_r, p_o = dict(ref.named_parameters()), dict(ours.named_parameters())
    scale = max(p.grad.abs().max().item() for p in p_r.values() if p.grad is not None)
    for k in p_r:
        assert (p_r[k].grad is None) == (p_o[k].grad is None), k
        if p_r[k].grad is not None:
            assert (p_r[k].grad - p_o[k].grad).abs().max().item() <= 2e-4 * scale, k


def test_multi_source_job_is_the_union_of_its_parts(monkeypatch):
    """plan.MultiXfJob (one gather-reduce over one Z table for all small transform-first jobs into a destination type):
    integer bookkeeping -- every destination row gathers exactly the union of what the separate jobs gather, from the
    re-based table rows, with the same mean weights; the transposed CSR is its exact transpose."""
    import numpy as np
    _cpu_kernels.install(monkeypatch)
    from kgwas_b200 import make_synth_kg
    from kgwas_b200 import plan as P
    P.clear_plan_cache()
    data = make_synth_kg(0.01, 3, hidden=32)
    num_nodes = {t: int(x.size(0)) for t, x in data.x_dict.items()}
    sep = P.get_plan(data.edge_index_dict, num_nodes, merge_xf=False)
    mer = P.get_plan(data.edge_index_dict, num_nodes, merge_xf=True)
    assert sep is not mer and mer.n_edges == sep.n_edges
    merged = [j for j in mer.jobs["Gene"] if getattr(j, "multi", False)]
    assert len(merged) == 1 and len(merged[0].parts) >= 2
    job = merged[0]
    parts = {j.src_type: j for j in sep.jobs["Gene"] if j.mode == "xf" and j.src_type in [p[0] for p in job.parts]}
    assert job.n_edges == sum(p.n_edges for p in parts.values())
    rp, col, w = job.csr.rowptr.numpy(), job.csr.col.numpy(), job.w_mean.numpy()
    for t in range(num_nodes["Gene"]):
        got = sorted(zip(col[rp[t]:rp[t + 1]].tolist(), np.round(w[rp[t]:rp[t + 1]], 7).tolist()))
        want = []
        for (S, R, lo, hi, n_src, off) in job.parts:
            pj = parts[S]
            assert (pj.R, pj.n_src, pj.rel_ids[0], pj.rel_ids[-1] + 1) == (R, n_src, lo, hi)
            prp, pcol, pw = pj.csr.rowptr.numpy(), pj.csr.col.numpy(), pj.w_mean.numpy()
            want += list(zip((pcol[prp[t]:prp[t + 1]] + off).tolist(), np.round(pw[prp[t]:prp[t + 1]], 7).tolist()))
        assert got == sorted(want), t
    # transposed CSR: same (row, column, weight) triples
    trp, tcol, tw = job.tcsr.rowptr.numpy(), job.tcsr.col.numpy(), job.w_mean_t.numpy()
    fwd = sorted((int(c), t, round(float(x), 7)) for t in range(num_nodes["Gene"]) for c, x in zip(col[rp[t]:rp[t + 1]], w[rp[t]:rp[t + 1]]))
    bwd = sorted((r, int(c), round(float(x), 7)) for r in range(job.total_rows) for c, x in zip(tcol[trp[r]:trp[r + 1]], tw[trp[r]:trp[r + 1]]))
    assert fwd == bwd
    P.clear_plan_cache()

"""Generate tests/golden/ref_*.pt by executing the reference's in-tree files VERBATIM.

    python oracle/gen_golden_from_reference.py          (needs /root/reference; CPU only)

/root/reference/kgwas/model.py and /root/reference/kgwas/conv.py are loaded as they lie (nothing is copied) on
top of oracle/pyg_standin -- a minimal stand-in for the PyG base classes those files import, because
torch_geometric is an un-vendored dependency that cannot be installed offline.  The fixtures therefore pin the
in-tree half of the hot path (HeteroGNN wiring, GATConv arithmetic: leaky-relu / softmax / temperature / sigmoid /
raw modes, bias, attention export); the PyG half (SAGEConv, HeteroConv, scatter, softmax) is pinned only as far
as the stand-in restates PyG faithfully.  The fixtures travel; this script and /root/reference do not need to.
"""
import importlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/kgwas"
OUT = os.path.join(ROOT, "tests", "golden")


def load_reference():
    sys.path.insert(0, os.path.join(HERE, "pyg_standin"))
    pkg = types.ModuleType("kgwas_ref")          # package shell: kgwas/__init__.py (data loaders, pandas, ...) is NOT run
    pkg.__path__ = [REF]
    sys.modules["kgwas_ref"] = pkg
    return importlib.import_module("kgwas_ref.model"), importlib.import_module("kgwas_ref.conv")


class _Graph:
    def __init__(self, edge_types):
        self.edge_types = edge_types


def tiny_kg(seed, h):
    g = torch.Generator().manual_seed(seed)
    n = {"SNP": 40, "Gene": 12, "CellularComponent": 5, "BiologicalProcess": 7, "MolecularFunction": 6}
    ets = [("SNP", "TSS", "Gene"), ("SNP", "eQTL", "Gene"), ("Gene", "Gene-Signaling-Gene", "Gene"),
           ("Gene", "Gene-Reaction-Gene", "Gene"),
           ("Gene", "Gene-Associates-BiologicalProcess", "BiologicalProcess"),
           ("Gene", "Gene-Regulates-BiologicalProcess", "BiologicalProcess"),
           ("Gene", "Gene-Enables-MolecularFunction", "MolecularFunction"),
           ("Gene", "Gene-NotContributes-MolecularFunction", "MolecularFunction"),
           ("Gene", "Gene-LocatedIn-CellularComponent", "CellularComponent"),
           ("Gene", "Gene-NotColocalizes-CellularComponent", "CellularComponent")]
    ei = {}
    for i, et in enumerate(ets):
        e = [90, 60, 30, 25, 20, 12, 14, 9, 10, 8][i]
        ei[et] = torch.stack([torch.randint(0, n[et[0]], (e,), generator=g), torch.randint(0, n[et[2]], (e,), generator=g)])
    for et in list(ei):                                   # reverse twins of bipartite relations, as load_kg produces
        if et[0] != et[2]:
            ei[(et[2], "rev_" + et[1], et[0])] = ei[et].flip([0])
    x = {t: torch.randn(c, h, generator=g) for t, c in n.items()}
    return n, ei, x


def mid_kg(seed, h, n_snp=3000, n_gene=260):
    """A KG big enough for the kernels the benchmark runs: h % 128 == 0 -> lean gather-reduce, >= 916 SNP rows ->
    tcgen05 GEMMs, hub genes with several hundred in-edges -> heavy-row segments and the hub-tile path."""
    g = torch.Generator().manual_seed(seed)
    n = {"SNP": n_snp, "Gene": n_gene, "CellularComponent": 9, "BiologicalProcess": 40, "MolecularFunction": 17}
    spec = [(("SNP", "TSS", "Gene"), 9000), (("SNP", "eQTL", "Gene"), 5000), (("SNP", "ABC", "Gene"), 2500),
            (("Gene", "Gene-Signaling-Gene", "Gene"), 1500), (("Gene", "Gene-Reaction-Gene", "Gene"), 900),
            (("Gene", "Gene-Associates-BiologicalProcess", "BiologicalProcess"), 500),
            (("Gene", "Gene-Enables-MolecularFunction", "MolecularFunction"), 300),
            (("Gene", "Gene-NotContributes-MolecularFunction", "MolecularFunction"), 60),
            (("Gene", "Gene-LocatedIn-CellularComponent", "CellularComponent"), 200)]
    ei = {}
    for et, e in spec:
        src = torch.randint(0, n[et[0]], (e,), generator=g)
        if et[2] == "Gene":                          # skewed destinations: a few hub genes
            u = torch.rand(e, generator=g)
            dst = (n["Gene"] * u ** 3).long().clamp_(max=n["Gene"] - 1)
        else:
            dst = torch.randint(0, n[et[2]], (e,), generator=g)
        ei[et] = torch.stack([src, dst])
    for et in list(ei):
        if et[0] != et[2]:
            ei[(et[2], "rev_" + et[1], et[0])] = ei[et].flip([0])
    return n, ei


def mid_features(n, h, seed):
    from oracle.seeded import seeded_tensor
    return {t: seeded_tensor("x." + t, (c, h), seed, 1.0) for t, c in n.items()}


def big_cases(model_mod):
    """Reference-executed fixtures at the benchmark's widths (weights and features are regenerated from seeds by the
    tests, see oracle/seeded.py; the fixture holds edge lists, logits, hidden rows, loss and gradient samples)."""
    from oracle.seeded import canonical_key, fill_parameters, grad_sample
    for backbone, L, h in (("SAGE", 2, 128), ("GAT", 2, 128), ("GAT", 3, 256)):
        n, eid = mid_kg(5, h)
        x = mid_features(n, h, 11)
        torch.manual_seed(7)
        m = model_mod.HeteroGNN(_Graph(list(eid.keys())), h, 1, L, backbone, "sum", h, h, h, 1)
        bs = 1500
        m({k: v.clone() for k, v in x.items()}, eid, bs)            # materialise the lazy weights
        with torch.no_grad():
            for name, p in m.named_parameters():
                if isinstance(p, torch.nn.parameter.UninitializedParameter):
                    continue
                scale = 1.0 / (p.size(-1) ** 0.5) if p.dim() >= 2 else 0.1
                from oracle.seeded import seeded_tensor
                p.copy_(seeded_tensor(canonical_key(name), p.shape, 21, scale))
            _, hid = m({k: v.clone() for k, v in x.items()}, eid, bs, return_h=True)
            # centre the head so that about half of the logits survive the final ReLU (model.py:86) -- between two
            # well separated logits, never ON one: a logit at 0 +- 1e-7 would make the ReLU mask (and with it every
            # gradient) depend on the last bit of the arithmetic
            pre = (hid @ m.lin.weight.t()).reshape(-1).sort().values
            gaps = pre[1:] - pre[:-1]
            mid = pre.numel() // 2
            j = mid - 50 + int(gaps[mid - 50:mid + 50].argmax())
            assert float(gaps[j]) > 1e-3 * float(pre.abs().max()), float(gaps[j])
            m.lin.bias.copy_(-0.5 * (pre[j] + pre[j + 1]).reshape(1))
        out, hid = m({k: v.clone() for k, v in x.items()}, eid, bs, return_h=True)
        w = torch.rand(bs, dtype=torch.float64, generator=torch.Generator().manual_seed(9))
        y = torch.randn(bs, generator=torch.Generator().manual_seed(10))
        loss = torch.mean(w * (out.reshape(-1) - y) ** 2)
        loss.backward()
        shapes = {canonical_key(k): tuple(p.shape) for k, p in m.named_parameters()
                  if not isinstance(p, torch.nn.parameter.UninitializedParameter)}
        grads = {canonical_key(k): (None if p.grad is None else grad_sample(p.grad)) for k, p in m.named_parameters()
                 if not isinstance(p, torch.nn.parameter.UninitializedParameter)}
        lazy = [canonical_key(k) for k, v in m.state_dict().items()
                if isinstance(v, torch.nn.parameter.UninitializedParameter)]
        rec = {"num_nodes": n, "edge_index": {k: v.to(torch.int32) for k, v in eid.items()}, "batch_size": bs,
               "feature_seed": 11, "param_seed": 21, "hidden_dim": h, "layers": L, "backbone": backbone,
               "w": w, "y": y, "out": out.detach(), "hidden": hid.detach()[:, :16].clone(), "loss": loss.detach(),
               "grads": grads, "lazy_keys": lazy, "param_shapes": shapes, "lin_bias": m.lin.bias.detach().clone(),
               "frac_positive_logits": float((out > 0).float().mean())}
        torch.save(rec, os.path.join(OUT, f"ref_mid_{backbone.lower()}_L{L}_h{h}.pt"))
        print(backbone, L, h, "loss", float(loss.detach()), "positive logits", rec["frac_positive_logits"])


def _state(module):
    """state_dict split into plain tensors + the keys that are still lazy (never materialised in the reference:
    e.g. GATConv.lin_dst of a same-type relation, kgwas/conv.py:136-138)."""
    sd = module.state_dict()
    lazy = [k for k, v in sd.items() if isinstance(v, torch.nn.parameter.UninitializedParameter)]
    return {"state": {k: v.detach().clone() for k, v in sd.items() if k not in lazy}, "lazy_keys": lazy}


def main():
    os.makedirs(OUT, exist_ok=True)
    model_mod, conv_mod = load_reference()
    h = 32
    # ---- GATConv alone: every attention mode, tuple and single-tensor inputs -----------------------------
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)
    x_src, x_dst = torch.randn(9, h, generator=g), torch.randn(6, h, generator=g)
    ei = torch.stack([torch.randint(0, 9, (40,), generator=g), torch.randint(0, 6, (40,), generator=g)])
    ei[:, :3] = torch.tensor([[0, 0, 0], [2, 2, 2]])              # a triple edge
    ei[1][ei[1] == 5] = 4                                          # destination 5 isolated
    cases = []
    for kw in ({}, {"temperature": 0.5}, {"sigmoid_gat": True, "temperature": 2.0}):
        conv = conv_mod.GATConv((-1, -1), h, heads=1, add_self_loops=False, **kw)
        for raw in (None, True):
            out, (_, alpha) = conv((x_src, x_dst), ei, return_attention_weights=True, return_raw_attention_weights=raw)
            cases.append({"kwargs": kw, "raw": raw, "out": out.detach(), "alpha": alpha.detach(), **_state(conv)})
    conv1 = conv_mod.GATConv(h, h, heads=1, add_self_loops=False)
    ei1 = torch.stack([torch.randint(0, 9, (30,), generator=g), torch.randint(0, 9, (30,), generator=g)])
    out1 = conv1(x_src, ei1)
    torch.save({"x_src": x_src, "x_dst": x_dst, "edge_index": ei, "cases": cases,
                "single": {"edge_index": ei1, "out": out1.detach(), **_state(conv1)}},
               os.path.join(OUT, "ref_gatconv.pt"))
    # ---- HeteroGNN (model.py) SAGE and GAT: logits, hidden, attention summary, parameter gradients ------------
    n, eid, x = tiny_kg(3, h)
    for backbone in ("SAGE", "GAT"):
        for aggr in ("sum", "mean"):
            torch.manual_seed(7)
            m = model_mod.HeteroGNN(_Graph(list(eid.keys())), h, 1, 2, backbone, aggr, h, h, h, 1)
            bs = 25
            out, hid = m({k: v.clone() for k, v in x.items()}, eid, bs, return_h=True)
            with torch.no_grad():                        # biases start at zero: make them matter
                for name, p in m.named_parameters():
                    if name.endswith(".bias") and ".convs." in name:
                        p.normal_(0, 0.1)
            out, hid = m({k: v.clone() for k, v in x.items()}, eid, bs, return_h=True)
            w = torch.rand(bs, dtype=torch.float64, generator=torch.Generator().manual_seed(9))
            y = torch.randn(bs, generator=torch.Generator().manual_seed(10))
            loss = torch.mean(w * (out.reshape(-1) - y) ** 2)          # kgwas/kgwas.py:145
            loss.backward()
            grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()
                     if not isinstance(p, torch.nn.parameter.UninitializedParameter)}
            rec = {"x": x, "edge_index": eid, "batch_size": bs, "w": w, "y": y, "out": out.detach(), "hidden": hid.detach(),
                   "loss": loss.detach(), "grads": grads, **_state(m)}
            if backbone == "GAT":
                m.zero_grad()
                o2, att = m({k: v.clone() for k, v in x.items()}, eid, bs, return_attention_weights=True)
                rec["att_out"], rec["att_mean"] = o2.detach(), [a.detach() for a in att]
            torch.save(rec, os.path.join(OUT, f"ref_heterognn_{backbone.lower()}_{aggr}.pt"))
    sys.path.insert(0, ROOT)
    big_cases(model_mod)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()

"""Per-parameter gradient error of the CUDA path on the reference-executed mid fixtures (tests/golden/ref_mid_*)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kgwas_b200
from oracle.seeded import seeded_tensor
from oracle import kgwas_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class G:
    def __init__(self, ets):
        self.edge_types = ets


def run(cls, f, dev, dtype=torch.float32):
    h, L = f["hidden_dim"], f["layers"]
    ei = {k: v.long() for k, v in f["edge_index"].items()}
    x = {t: seeded_tensor("x." + t, (c, h), f["feature_seed"], 1.0) for t, c in f["num_nodes"].items()}
    m = cls(G(list(ei.keys())), h, 1, L, f["backbone"], "sum", h, h, h, 1)
    state = {}
    for k, shape in f["param_shapes"].items():
        scale = 1.0 / (shape[-1] ** 0.5) if len(shape) >= 2 else 0.1
        state[k] = seeded_tensor(k, shape, f["param_seed"], scale)
    state["lin.bias"] = f["lin_bias"]
    m.load_state_dict(state, strict=False)
    m = m.to(dev).to(dtype)
    out, hid = m({k: v.to(dev).to(dtype) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()}, f["batch_size"], return_h=True)
    loss = torch.mean(f["w"].to(dev) * (out.reshape(-1) - f["y"].to(dev).to(dtype)) ** 2)
    loss.backward()
    return out.detach().cpu().double(), {k: (p.grad.detach().cpu().double() if p.grad is not None else None) for k, p in m.named_parameters()
                                         if not isinstance(p, torch.nn.parameter.UninitializedParameter)}


for name in sys.argv[1:] or ["sage_L2_h128", "gat_L2_h128", "gat_L3_h256"]:
    f = torch.load(os.path.join(GOLD, f"ref_mid_{name}.pt"), weights_only=True)
    o64, g64 = run(O.HeteroGNN, f, "cpu", torch.float64)
    o32, g32 = run(O.HeteroGNN, f, "cpu", torch.float32)
    oc, gc = run(kgwas_b200.HeteroGNN, f, "cuda")
    print(name, "logits: cuda vs fp64", float((oc - o64).abs().max() / o64.abs().max()), "fp32 oracle vs fp64", float((o32 - o64).abs().max() / o64.abs().max()))
    gscale = max(float(g.abs().max()) for g in g64.values() if g is not None)
    rows = []
    for k, g in g64.items():
        if g is None:
            continue
        a = float(g.abs().max())
        rows.append((float((gc[k] - g).abs().max()) / a, float((g32[k] - g).abs().max()) / a, a / gscale, k))
    rows.sort(reverse=True)
    for r in rows[:12]:
        print("  cuda %.2e  fp32-oracle %.2e  absmax/gscale %.2e  %s" % r)

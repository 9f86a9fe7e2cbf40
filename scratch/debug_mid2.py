"""Bisect the gradient error on the mid SAGE fixture graph: batch size, MLPs, layer count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kgwas_b200
from oracle.seeded import seeded_tensor
from oracle import kgwas_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
f = torch.load(os.path.join(GOLD, "ref_mid_sage_L2_h128.pt"), weights_only=True)


class G:
    def __init__(self, ets):
        self.edge_types = ets


def run(cls, dev, dtype, bs, L, from_hidden, relu_out=True):
    h = f["hidden_dim"]
    ei = {k: v.long() for k, v in f["edge_index"].items()}
    x = {t: seeded_tensor("x." + t, (c, h), f["feature_seed"], 1.0) for t, c in f["num_nodes"].items()}
    m = cls(G(list(ei.keys())), h, 1, L, "SAGE", "sum", h, h, h, 1, no_relu=not relu_out)
    state = {}
    for k, shape in f["param_shapes"].items():
        if k.startswith("convs.") and int(k.split(".")[1]) >= L:
            continue
        scale = 1.0 / (shape[-1] ** 0.5) if len(shape) >= 2 else 0.1
        state[k] = seeded_tensor(k, shape, f["param_seed"], scale)
    state["lin.bias"] = f["lin_bias"]
    m.load_state_dict(state, strict=False)
    m = m.to(dev).to(dtype)
    xx = {k: v.to(dev).to(dtype).requires_grad_() for k, v in x.items()}
    eid = {k: v.to(dev) for k, v in ei.items()}
    if from_hidden:
        if cls is O.HeteroGNN:
            xd = O.conv_stack_forward(m.convs, xx, eid)
            out = m.lin(xd["SNP"])[:bs]
            out = out.relu() if relu_out else out
        else:
            out = m.forward_from_hidden(xx, eid, bs)
    else:
        out = m(dict(xx), eid, bs)
    w = f["w"][:bs] if bs <= 1500 else torch.ones(bs, dtype=torch.float64)
    y = f["y"][:bs] if bs <= 1500 else torch.zeros(bs)
    loss = torch.mean(w.to(dev) * (out.reshape(-1) - y.to(dev).to(dtype)) ** 2)
    loss.backward()
    g = {k: (p.grad.detach().cpu().double() if p.grad is not None else None) for k, p in m.named_parameters()
         if not isinstance(p, torch.nn.parameter.UninitializedParameter)}
    for k, v in xx.items():
        g["x." + k] = v.grad.detach().cpu().double() if v.grad is not None else None
    return out.detach().cpu().double(), g


for (bs, L, fh, ro) in ((1500, 2, False, True), (1500, 2, True, True), (3000, 2, True, True), (1500, 2, True, False),
                        (1500, 1, True, True), (3000, 1, True, False)):
    o64, g64 = run(O.HeteroGNN, "cpu", torch.float64, bs, L, fh, ro)
    oc, gc = run(kgwas_b200.HeteroGNN, "cuda", torch.float32, bs, L, fh, ro)
    rows = []
    for k, g in g64.items():
        if g is None or gc.get(k) is None:
            continue
        a = float(g.abs().max())
        if a == 0:
            continue
        rows.append((float((gc[k] - g).abs().max()) / a, k))
    rows.sort(reverse=True)
    print(f"bs={bs} L={L} from_hidden={fh} relu_out={ro}: logits {float((oc - o64).abs().max() / o64.abs().max()):.2e}; worst grads:",
          ", ".join("%.1e %s" % r for r in rows[:4]))

"""Final-gradient error of the full model (MLPs + conv stack) on the mid SAGE fixture, by scheduler mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kgwas_b200
from kgwas_b200 import ops
from oracle.seeded import seeded_tensor
from oracle import kgwas_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
f = torch.load(os.path.join(GOLD, "ref_mid_sage_L2_h128.pt"), weights_only=True)


class G:
    def __init__(self, ets):
        self.edge_types = ets


def run(cls, dev, dtype, mode):
    h, L, bs = 128, 2, 1500
    ei = {k: v.long() for k, v in f["edge_index"].items()}
    x = {t: seeded_tensor("x." + t, (c, h), f["feature_seed"], 1.0) for t, c in f["num_nodes"].items()}
    m = cls(G(list(ei.keys())), h, 1, L, "SAGE", "sum", h, h, h, 1)
    state = {}
    for k, shape in f["param_shapes"].items():
        scale = 1.0 / (shape[-1] ** 0.5) if len(shape) >= 2 else 0.1
        state[k] = seeded_tensor(k, shape, f["param_seed"], scale)
    state["lin.bias"] = f["lin_bias"]
    m.load_state_dict(state, strict=False)
    m = m.to(dev).to(dtype)
    xx = {k: v.to(dev).to(dtype).requires_grad_() for k, v in x.items()}
    eid = {k: v.to(dev) for k, v in ei.items()}
    if mode == "split":          # encode, detach, conv stack; then push the gradient through the MLPs by hand
        enc = m.encode(dict(xx))
        enc_d = {k: v.detach().requires_grad_() for k, v in enc.items()}
        out = m.forward_from_hidden(enc_d, eid, bs)
    else:
        out = m(dict(xx), eid, bs)
    loss = torch.mean(f["w"].to(dev) * (out.reshape(-1) - f["y"].to(dev).to(dtype)) ** 2)
    loss.backward()
    if mode == "split":
        torch.cuda.synchronize()
        torch.autograd.backward([enc[k] for k in enc], [enc_d[k].grad for k in enc])
    g = {k: (p.grad.detach().cpu().double() if p.grad is not None else None) for k, p in m.named_parameters()
         if not isinstance(p, torch.nn.parameter.UninitializedParameter)}
    for k, v in xx.items():
        g["x." + k] = v.grad.detach().cpu().double() if v.grad is not None else None
    return g


g64 = run(O.HeteroGNN, "cpu", torch.float64, "full")
for ms, mode in ((True, "full"), (False, "full"), (True, "split")):
    ops.MULTI_STREAM = ms
    gc = run(kgwas_b200.HeteroGNN, "cuda", torch.float32, mode)
    rows = sorted(((float((gc[k] - g).abs().max()) / float(g.abs().max()), k) for k, g in g64.items()
                   if g is not None and gc.get(k) is not None and float(g.abs().max()) > 0), reverse=True)
    print(f"MULTI_STREAM={ms} mode={mode} blocking={os.environ.get('CUDA_LAUNCH_BLOCKING')}: worst grads:", ", ".join("%.1e %s" % r for r in rows[:4]))

"""Where does the conv-stack input gradient go wrong on MLP-produced features?  d loss / d enc vs the fp64 oracle, per
node type and per row, plus the largest per-call error of every kgb_gemm / kgb_spmm / relu_bwd_fused call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kgwas_b200
from kgwas_b200 import _lib, ops
from oracle.seeded import seeded_tensor
from oracle import kgwas_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
f = torch.load(os.path.join(GOLD, "ref_mid_sage_L2_h128.pt"), weights_only=True)
h, L, bs = 128, 2, 1500


class G:
    def __init__(self, ets):
        self.edge_types = ets


def build(cls, dev, dtype):
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
    return m, {k: v.to(dev).to(dtype) for k, v in x.items()}, {k: v.to(dev) for k, v in ei.items()}


# oracle fp64: encoded features and d loss / d enc
m64, x64, ei64 = build(O.HeteroGNN, "cpu", torch.float64)
enc64 = {"SNP": m64.snp_feat_mlp(x64["SNP"]), "Gene": m64.gene_feat_mlp(x64["Gene"])}
for t in ("CellularComponent", "BiologicalProcess", "MolecularFunction"):
    enc64[t] = m64.go_feat_mlp(x64[t])
enc64 = {k: v.detach().requires_grad_() for k, v in enc64.items()}
xd = O.conv_stack_forward(m64.convs, enc64, ei64)
out64 = m64.lin(xd["SNP"])[:bs].relu()
l64 = torch.mean(f["w"] * (out64.reshape(-1) - f["y"].double()) ** 2)
l64.backward()

mc, xc, eic = build(kgwas_b200.HeteroGNN, "cuda", torch.float32)
for feed in ("oracle-enc", "own-enc"):
    if feed == "oracle-enc":
        enc = {k: v.detach().float().cuda().requires_grad_() for k, v in enc64.items()}
    else:
        with torch.no_grad():
            e = mc.encode(dict(xc))
        enc = {k: v.detach().requires_grad_() for k, v in e.items()}
        for k in enc:
            print("  enc diff", k, float((enc[k].detach().cpu().double() - enc64[k].detach()).abs().max()))
    out = mc.forward_from_hidden(enc, eic, bs)
    loss = torch.mean(f["w"].cuda() * (out.reshape(-1) - f["y"].cuda()) ** 2)
    mc.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    print(feed, "loss", float(loss), float(l64), "logit err", float((out.detach().cpu().double() - out64.detach()).abs().max() / out64.detach().abs().max()))
    for k in enc:
        g, r = enc[k].grad.detach().cpu().double(), enc64[k].grad
        rowerr = (g - r).abs().max(1).values
        print(f"  d enc[{k}]: max err / absmax = {float(rowerr.max() / r.abs().max()):.2e}; worst row {int(rowerr.argmax())}, "
              f"|ref row| {float(r[int(rowerr.argmax())].abs().max()):.2e}, ref absmax {float(r.abs().max()):.2e}")

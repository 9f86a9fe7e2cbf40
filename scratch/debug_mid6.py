"""Layer-by-layer: forward outputs, ReLU masks and single-layer backward on MLP-produced features (mid SAGE fixture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kgwas_b200
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


m64, x64, ei64 = build(O.HeteroGNN, "cpu", torch.float64)
with torch.no_grad():
    enc64 = {"SNP": m64.snp_feat_mlp(x64["SNP"]), "Gene": m64.gene_feat_mlp(x64["Gene"])}
    for t in ("CellularComponent", "BiologicalProcess", "MolecularFunction"):
        enc64[t] = m64.go_feat_mlp(x64[t])
mc, xc, eic = build(kgwas_b200.HeteroGNN, "cuda", torch.float32)
gen = torch.Generator().manual_seed(3)
for which, feats in (("mlp", enc64), ("randn", {k: torch.randn(v.shape, generator=gen, dtype=torch.float64) for k, v in enc64.items()})):
    cur64 = {k: v.clone().requires_grad_() for k, v in feats.items()}
    curc = {k: v.float().cuda().requires_grad_() for k, v in feats.items()}
    for li in range(L):
        o64 = {k: v.relu() for k, v in m64.convs[li](cur64, ei64).items()}
        oc = mc.convs[li](curc, eic, _fuse_relu=True)
        up = {k: torch.randn(v.shape, generator=gen, dtype=torch.float64) for k, v in o64.items()}
        g64 = torch.autograd.grad([o64[k] for k in o64], [cur64[k] for k in cur64], [up[k] for k in o64], allow_unused=True)
        gc = torch.autograd.grad([oc[k] for k in o64], [curc[k] for k in cur64], [up[k].float().cuda() for k in o64], allow_unused=True)
        for k in o64:
            a, b = oc[k].detach().cpu().double(), o64[k].detach()
            mism = int(((a > 0) != (b > 0)).sum())
            print(f"{which} layer {li} out[{k}]: err {float((a - b).abs().max() / b.abs().max()):.2e}, mask mismatches {mism} of {a.numel()}, zeros {int((b == 0).sum())}")
        for k, ga, gb in zip(cur64, gc, g64):
            if gb is None:
                continue
            e = (ga.detach().cpu().double() - gb).abs().max(1).values
            print(f"{which} layer {li} d in[{k}]: err/absmax {float(e.max() / gb.abs().max()):.2e} worst row {int(e.argmax())} rows>1e-3: {int((e > 1e-3 * gb.abs().max()).sum())}")
        cur64 = {k: v.detach().requires_grad_() for k, v in o64.items()}
        curc = {k: v.detach().cpu().float().cuda().requires_grad_() for k, v in o64.items()}    # same inputs for the next layer

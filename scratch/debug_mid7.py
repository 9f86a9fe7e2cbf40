"""Composite model on MLP features: d loss / d x1 (input of the last layer) with and without the fused head."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kgwas_b200
from kgwas_b200 import model as kmodel
from oracle import kgwas_oracle as O
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from debug_mid6 import build, m64, enc64, ei64, mc, eic, f, bs, L   # noqa: E402  (re-runs the layer checks: ignore)

print("=" * 30)
for fused in (True, False):
    kmodel._head_is_fusable = (lambda w: w.size(0) == 1 and w.is_cuda) if fused else (lambda w: False)
    # oracle
    e64 = {k: v.clone().requires_grad_() for k, v in enc64.items()}
    x1_64 = {k: v.relu() for k, v in m64.convs[0](e64, ei64).items()}
    for v in x1_64.values():
        v.retain_grad()
    x2_64 = {k: v.relu() for k, v in m64.convs[1](x1_64, ei64).items()}
    out64 = m64.lin(x2_64["SNP"])[:bs].relu()
    l64 = torch.mean(f["w"] * (out64.reshape(-1) - f["y"].double()) ** 2)
    l64.backward()
    # ours, same structure as HeteroGNN.forward_from_hidden
    ec = {k: v.float().cuda().requires_grad_() for k, v in enc64.items()}
    x1 = mc.convs[0](ec, eic, _fuse_relu=True)
    for v in x1.values():
        v.retain_grad()
    head = ("SNP", mc.lin.weight) if fused else None
    x2 = mc.convs[1](x1, eic, _fuse_relu=True, _head=head)
    logits = x2.pop(("head", "SNP"), None)
    if logits is not None:
        out = (logits[:bs] + mc.lin.bias).relu()
    else:
        out = mc.head(x2["SNP"][:bs]).relu()
    loss = torch.mean(f["w"].cuda() * (out.reshape(-1) - f["y"].cuda()) ** 2)
    mc.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    print(f"fused head = {fused} (got fused logits: {logits is not None}); loss {float(loss):.6f} vs {float(l64):.6f}")
    for k in x1_64:
        if x1_64[k].grad is None:
            continue
        g, r = x1[k].grad.detach().cpu().double(), x1_64[k].grad
        e = (g - r).abs().max(1).values
        print(f"  d x1[{k}]: err/absmax {float(e.max() / r.abs().max()):.2e}, rows > 1e-3: {int((e > 1e-3 * r.abs().max()).sum())}, worst row {int(e.argmax())}")
    for k in e64:
        g, r = ec[k].grad.detach().cpu().double(), e64[k].grad
        e = (g - r).abs().max(1).values
        print(f"  d enc[{k}]: err/absmax {float(e.max() / r.abs().max()):.2e}, rows > 1e-3: {int((e > 1e-3 * r.abs().max()).sum())}")

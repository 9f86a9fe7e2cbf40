"""Check every kgb_gemm / kgb_spmm call of one fwd+bwd on the mid SAGE fixture (full model incl. MLPs) against fp64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kgwas_b200
from kgwas_b200 import _lib, ops
from oracle.seeded import seeded_tensor

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
f = torch.load(os.path.join(GOLD, "ref_mid_sage_L2_h128.pt"), weights_only=True)
ops.MULTI_STREAM = False
real_gemm, real_spmm = _lib.gemm, _lib.spmm
bad = []


def gemm(layout, a, b, c, m, n, k, *, alpha=1.0, beta=0.0, bias=None, relu=False):
    c0 = c[:m, :n].double().clone() if beta != 0 else None
    real_gemm(layout, a, b, c, m, n, k, alpha=alpha, beta=beta, bias=bias, relu=relu)
    A, B = a.double(), b.double()
    if layout == _lib.KGB_NT:
        ref = A[:m, :k] @ B[:n, :k].t()
    elif layout == _lib.KGB_NN:
        ref = A[:m, :k] @ B[:k, :n]
    else:
        ref = A[:k, :m].t() @ B[:k, :n]
    ref = alpha * ref
    if c0 is not None:
        ref = ref + beta * c0
    if bias is not None:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp(min=0)
    err = float((c[:m, :n].double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    tag = f"gemm layout={layout} m={m} n={n} k={k} alpha={alpha} beta={beta} bias={bias is not None} relu={relu} strides a={a.stride(0)} b={b.stride(0)} c={c.stride(0)} |a|={float(a.abs().max()):.2e} |b|={float(b.abs().max()):.2e}"
    if err > 1e-4:
        bad.append((err, tag))
    return c


def spmm(csr, x, y, h, **kw):
    beta = kw.get("beta", 0.0)
    y0 = y.double().clone() if beta != 0 else None
    real_spmm(csr, x, y, h, **kw)
    ew = kw.get("ew")
    rows = torch.repeat_interleave(torch.arange(csr.n_rows, device=x.device), (csr.rowptr[1:] - csr.rowptr[:-1]).long())
    w = ew.double() if ew is not None else torch.ones(csr.n_edges, dtype=torch.float64, device=x.device)
    ref = torch.zeros(csr.n_rows, h, dtype=torch.float64, device=x.device).index_add_(0, rows, x.double()[csr.col.long(), :h] * w[:, None])
    if y0 is not None:
        ref = ref + beta * y0[:, :h]
    if kw.get("bias") is not None:
        ref = ref + kw["bias"].double()
    if kw.get("relu"):
        ref = ref.clamp(min=0)
    err = float((y[:, :h].double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    if err > 1e-4:
        bad.append((err, f"spmm rows={csr.n_rows} cols={csr.n_cols} E={csr.n_edges} beta={beta} hsegs={csr.n_hsegs}"))
    return y


_lib.gemm, _lib.spmm = gemm, spmm


class G:
    def __init__(self, ets):
        self.edge_types = ets


h, L = 128, 2
ei = {k: v.long().cuda() for k, v in f["edge_index"].items()}
x = {t: seeded_tensor("x." + t, (c, h), f["feature_seed"], 1.0).cuda().requires_grad_() for t, c in f["num_nodes"].items()}
m = kgwas_b200.HeteroGNN(G(list(ei.keys())), h, 1, L, "SAGE", "sum", h, h, h, 1)
state = {}
for k, shape in f["param_shapes"].items():
    scale = 1.0 / (shape[-1] ** 0.5) if len(shape) >= 2 else 0.1
    state[k] = seeded_tensor(k, shape, f["param_seed"], scale)
state["lin.bias"] = f["lin_bias"]
m.load_state_dict(state, strict=False)
m = m.cuda()
print("allow_tf32 matmul:", torch.backends.cuda.matmul.allow_tf32, "precision:", torch.get_float32_matmul_precision())
enc = m.encode(dict(x))
for k, v in enc.items():
    v.retain_grad()
    print("encoded", k, "mean %.3f std %.3f absmax %.3f" % (float(v.mean()), float(v.std()), float(v.abs().max())))
out = m.forward_from_hidden(enc, ei, 1500)
loss = torch.mean(f["w"].cuda() * (out.reshape(-1) - f["y"].cuda()) ** 2)
loss.backward()
torch.cuda.synchronize()
print("calls with error > 1e-4:", len(bad))
for e, t in sorted(bad, reverse=True)[:20]:
    print("  %.2e %s" % (e, t))
# MLP backward check in fp64: d loss / d raw x from the retained d loss / d enc
for t, mlp in (("SNP", m.snp_feat_mlp), ("Gene", m.gene_feat_mlp)):
    m64 = type(mlp)(h, h, h).double().cuda()
    m64.load_state_dict({k: v.double() for k, v in mlp.state_dict().items()})
    xin = x[t].detach().double().requires_grad_()
    o = m64(xin)
    o.backward(enc[t].grad.double())
    print(t, "MLP backward (cuBLAS) vs fp64 given the same upstream grad: dx err %.2e, dW err %.2e" % (
        float((x[t].grad.double() - xin.grad).abs().max() / xin.grad.abs().max()),
        float((mlp.FC_hidden.weight.grad.double() - m64.FC_hidden.weight.grad).abs().max() / m64.FC_hidden.weight.grad.abs().max())))

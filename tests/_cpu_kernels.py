"""TEST INFRASTRUCTURE ONLY: plain-torch CPU stand-ins for the C-ABI kernels, with the signatures of kgwas_b200._lib.

They let the ``-m "not gpu"`` suite run the HOST logic of the fused layer (relation merging, transform-first /
aggregate-first jobs, pre-summed root weights, accumulate / ReLU placement, head fusion, gradient wiring, None
gradients, the launch scheduler's reordering) against the oracle without a GPU.  Nothing outside tests/ may import this
module; the product path has no CPU implementation (kgwas_b200._lib raises on CPU tensors).
"""
import numpy as np
import torch


class FakeCsr:
    def __init__(self, rowptr, col, n_rows, n_cols, seg_len=128):
        self.rowptr, self.col, self.n_rows, self.n_cols, self.seg_len = rowptr, col, n_rows, n_cols, seg_len
        self.n_hrows = self.n_hsegs = self.n_hgroups = 0

    @property
    def n_edges(self):
        return self.col.numel()

    def schedule_for_l2(self, row_bytes, window_bytes=0):
        return self

    def build_hub(self, ew, h, **kw):
        return False


def csr_build(src, dst, n_src, n_dst, transposed=True, seg_len=128, sort_cols=False, presort_key=None):
    """Same contract as kgb_csr_build (include/kgwas_b200.h): stable edge order, optional in-row column order."""
    s, d = src.numpy(), dst.numpy()
    if presort_key is not None or sort_cols:
        k2 = presort_key.numpy() if presort_key is not None else s
        first = np.argsort(k2, kind="stable")
        eperm = first[np.argsort(d[first], kind="stable")]
    else:
        eperm = np.argsort(d, kind="stable")
    col = s[eperm]
    rowptr = np.zeros(n_dst + 1, dtype=np.int64)
    np.add.at(rowptr, d + 1, 1)
    rowptr = np.cumsum(rowptr)
    i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a).astype(np.int32))
    fwd = FakeCsr(i32(rowptr), i32(col), n_dst, n_src, seg_len)
    if not transposed:
        return fwd, i32(eperm), None, None
    t_eperm = np.argsort(col, kind="stable")
    t_col = d[eperm][t_eperm]
    t_rowptr = np.zeros(n_src + 1, dtype=np.int64)
    np.add.at(t_rowptr, s + 1, 1)
    t_rowptr = np.cumsum(t_rowptr)
    return fwd, i32(eperm), FakeCsr(i32(t_rowptr), i32(t_col), n_src, n_dst, seg_len), i32(t_eperm)


def spmm(csr, x, y, h, *, ew=None, wperm=None, ew2=None, rowsum2=None, bins=1, beta=0.0, bias=None, relu=False,
         dot_w=None, dot_out=None):
    deg = (csr.rowptr[1:] - csr.rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(csr.n_rows), deg)
    col = csr.col.long()
    w = torch.ones(col.numel()) if ew is None else (ew[wperm.long()] if wperm is not None else ew)
    acc = torch.zeros(csr.n_rows, h).index_add_(0, rows, x[col, :h] * w[:, None])
    out = acc + (beta * y[:, :h] if beta != 0.0 else 0.0)
    if bias is not None:
        out = out + bias
    if relu:
        out = out.relu()
    y[:, :h] = out
    if rowsum2 is not None:
        w2 = ew2[wperm.long()] if wperm is not None else ew2
        rowsum2.zero_()
        rowsum2.view(-1).index_add_(0, rows * bins + (col % bins if bins > 1 else 0), w2)
    if dot_w is not None:
        dot_out.view(-1).copy_(y[:, :h] @ dot_w.reshape(-1))
    return y


def gemm(layout, a, b, c, m, n, k, alpha=1.0, beta=0.0, bias=None, relu=False):
    if layout == 0:
        p = a[:m, :k] @ b[:n, :k].T
    elif layout == 1:
        p = a[:m, :k] @ b[:k, :n]
    else:
        p = a[:k, :m].T @ b[:k, :n]
    out = alpha * p + (beta * c[:m, :n] if beta != 0.0 else 0.0)
    if bias is not None:
        out = out + bias
    if relu:
        out = out.relu()
    c[:m, :n] = out
    return c


def relu_bwd(dy, y, out=None):
    r = dy * (y > 0)
    if out is not None:
        out.copy_(r)
        return out
    return r


def relu_bwd_fused(g, h, *, dy=None, y=None, dp=None, wv=None, scale=1.0, sums=None):
    t = torch.zeros(g.size(0), h)
    if dy is not None:
        t = t + dy[:, :h]
    if dp is not None:
        t = t + dp.reshape(-1, 1) * wv.reshape(1, -1)
    if y is not None:
        t = t * (y[:, :h] > 0)
    g[:, :h] = scale * t
    if sums is not None:
        sums[0] = g[:, :h].sum(0)
        sums[1] = (dp.reshape(-1, 1) * y[:, :h]).sum(0) if (dp is not None and y is not None) else 0.0
    return g


def wcolsum(x, h, out, *, w=None, n_slots=1, beta=0.0):
    r = x[:, :h].sum(0, keepdim=True) if w is None else w[:, :n_slots].T @ x[:, :h]
    out.view(n_slots, h).copy_(r + (beta * out.view(n_slots, h) if beta != 0.0 else 0.0))
    return out


def rowdot(x, v, a, h, n_slots, slot_stride):
    for s in range(n_slots):
        a[:, s] = x[:, s * slot_stride:s * slot_stride + h] @ v[s]
    return a


def install(monkeypatch):
    """Swap the kernel entry points of kgwas_b200._lib for the stand-ins above (pytest monkeypatch: undone per test)."""
    from kgwas_b200 import _lib, ops, plan
    for name, fn in (("csr_build", csr_build), ("spmm", spmm), ("gemm", gemm), ("relu_bwd", relu_bwd),
                     ("relu_bwd_fused", relu_bwd_fused), ("wcolsum", wcolsum), ("rowdot", rowdot), ("Csr", FakeCsr)) + _GAT:
        monkeypatch.setattr(_lib, name, fn)
    monkeypatch.setattr(ops, "MULTI_STREAM", False)
    plan.clear_plan_cache()


# ---- GAT edge kernels (include/kgwas_b200.h: kgb_gat_alpha / kgb_sddmm / kgb_gat_dsoftmax) ----------------------------

def _groups(groups, n_slots, src_is_node):
    deg = (groups.rowptr[1:] - groups.rowptr[:-1]).long()
    g = torch.repeat_interleave(torch.arange(groups.n_rows), deg)
    col = groups.col.long()
    return g, (col * n_slots + g % n_slots) if src_is_node else col


def _seg_sum(v, g, n):
    return torch.zeros(n, dtype=v.dtype).index_add_(0, g, v)


def gat_alpha(groups, a_src, a_dst, n_slots, src_is_node, alpha, slope, temperature, mode):
    if groups.n_edges == 0:
        return alpha
    g, si = _groups(groups, n_slots, src_is_node)
    u = a_src.reshape(-1)[si] + a_dst.reshape(-1)[g]
    z = torch.nn.functional.leaky_relu(u, slope)
    if mode == 2:
        alpha.copy_(z)
    elif mode == 1:
        alpha.copy_(torch.sigmoid(z / temperature))
    else:
        s = z / temperature
        m = torch.full((groups.n_rows,), float("-inf")).scatter_reduce_(0, g, s, "amax")
        e = torch.exp(s - m[g])
        alpha.copy_(e / (_seg_sum(e, g, groups.n_rows) + 1e-16)[g])
    return alpha


def sddmm(csr, xrow, x, h, out):
    if csr.n_edges == 0:
        return out
    deg = (csr.rowptr[1:] - csr.rowptr[:-1]).long()
    r = torch.repeat_interleave(torch.arange(csr.n_rows), deg)
    out.copy_((xrow[r, :h] * x[csr.col.long(), :h]).sum(-1))
    return out


def gat_dsoftmax(groups, a_src, a_dst, n_slots, src_is_node, alpha, dalpha, du, da_dst, slope, temperature, mode):
    if groups.n_rows == 0:
        return du, da_dst
    if groups.n_edges == 0:
        da_dst.zero_()
        return du, da_dst
    g, si = _groups(groups, n_slots, src_is_node)
    u = a_src.reshape(-1)[si] + a_dst.reshape(-1)[g]
    if mode == 2:
        dz = dalpha.clone()
    elif mode == 1:
        dz = alpha * (1 - alpha) * dalpha / temperature
    else:
        dz = alpha * (dalpha - _seg_sum(alpha * dalpha, g, groups.n_rows)[g]) / temperature
    du.copy_(dz * torch.where(u > 0, torch.ones_like(u), torch.full_like(u, slope)))
    da_dst.view(-1).copy_(_seg_sum(du, g, groups.n_rows))
    return du, da_dst


def rank_update(a, v, y, h, n_slots, beta):
    y[:, :h] = a[:, :n_slots] @ v[:n_slots, :h] + (beta * y[:, :h] if beta != 0.0 else 0.0)
    return y


def permute_f32(w, perm, out=None):
    r = w[perm.long()]
    if out is not None:
        out.copy_(r)
        return out
    return r


_GAT = (("gat_alpha", gat_alpha), ("sddmm", sddmm), ("gat_dsoftmax", gat_dsoftmax), ("rank_update", rank_update),
        ("permute_f32", permute_f32))

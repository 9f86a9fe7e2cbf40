"""GPU: every kernel reached through the C ABI vs numpy / torch-fp64 on the same seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import bookkeeping as B

pytestmark = pytest.mark.gpu


def _rand_coo(rng, n_src, n_dst, e, hub=False):
    src = rng.integers(0, n_src, size=e)
    dst = rng.integers(0, n_dst, size=e)
    if hub and e > 10:
        dst[: e // 2] = int(rng.integers(0, n_dst))      # one destination takes half of all edges
        src[e // 4: e // 2] = int(rng.integers(0, n_src))  # and one source is a hub as well
    return src.astype(np.int64), dst.astype(np.int64)


@pytest.mark.parametrize("n_src,n_dst,e,hub", [(1, 1, 0, False), (5, 7, 1, False), (100, 37, 1000, False),
                                               (3000, 50, 40000, True), (17, 90001, 250000, True)])
@pytest.mark.parametrize("sort_cols", [False, True])
def test_csr_build_bit_exact(cuda, n_src, n_dst, e, hub, sort_cols):
    from kgwas_b200 import _lib
    rng = np.random.default_rng(e + n_src)
    src, dst = _rand_coo(rng, n_src, n_dst, e, hub)
    fwd, eperm, bwd, t_eperm = _lib.csr_build(torch.from_numpy(src).to(cuda), torch.from_numpy(dst).to(cuda), n_src, n_dst,
                                              seg_len=64, sort_cols=sort_cols)
    ref = B.csr_from_coo_ref(src, dst, n_src, n_dst, sort_cols)
    got = [fwd.rowptr, fwd.col, eperm, bwd.rowptr, bwd.col, t_eperm]
    for name, g, r in zip(["rowptr", "col", "eperm", "t_rowptr", "t_col", "t_eperm"], got, ref):
        assert np.array_equal(g.cpu().numpy(), r), name
    for csr in (fwd, bwd):
        hid, segptr, hseg = B.heavy_segments_ref(csr.rowptr.cpu().numpy(), 64)
        assert csr.n_hrows == len(hid) and csr.n_hsegs == len(hseg)
        if len(hid):
            assert np.array_equal(csr.hrow_id.cpu().numpy(), hid)
            assert np.array_equal(csr.hrow_segptr.cpu().numpy(), segptr)
            assert np.array_equal(csr.hseg_hrow.cpu().numpy(), hseg)


@pytest.mark.parametrize("h", [32, 64, 128, 256])
@pytest.mark.parametrize("weighted", [False, True])
def test_spmm_matches_fp64(cuda, h, weighted):
    from kgwas_b200 import _lib
    rng = np.random.default_rng(h)
    torch.manual_seed(h)
    n_src, n_dst, e = 700, 300, 30000
    src, dst = _rand_coo(rng, n_src, n_dst, e, hub=True)
    dst[dst == 5] = 6                                      # an isolated destination row
    s, d = torch.from_numpy(src).to(cuda), torch.from_numpy(dst).to(cuda)
    fwd, eperm, bwd, t_eperm = _lib.csr_build(s, d, n_src, n_dst, seg_len=64)
    assert fwd.n_hsegs > 0 and bwd.n_hsegs > 0
    x = torch.randn(n_src, h, device=cuda)
    w_coo = torch.rand(e, device=cuda) if weighted else None
    w_csr = w_coo[eperm.long()].contiguous() if weighted else None
    y = torch.full((n_dst, h), 7.0, device=cuda)
    _lib.spmm(fwd, x, y, h, ew=w_csr)
    msg = x.double()[s] * (w_coo.double()[:, None] if weighted else 1.0)
    ref = torch.zeros(n_dst, h, device=cuda, dtype=torch.float64).index_add_(0, d, msg)
    mag = torch.zeros(n_dst, h, device=cuda, dtype=torch.float64).index_add_(0, d, msg.abs())

    def close(a, b, m):      # fp32 summation error is bounded relative to sum |terms| (hub rows have 15k terms)
        return bool(((a.double() - b).abs() <= 2e-6 * m + 1e-6).all())

    assert close(y, ref, mag)
    assert y[5].abs().max() == 0
    # accumulate + relu epilogue, twice (the ticket counters must come back clean)
    for _ in range(2):
        y2 = torch.ones(n_dst, h, device=cuda)
        _lib.spmm(fwd, x, y2, h, ew=w_csr, beta=1.0, relu=True)
        assert close(y2, (ref + 1).clamp(min=0), mag + 1)
    # transposed pass with weights still in CSR order + second scalar row-summed
    g = torch.randn(n_dst, h, device=cuda)
    dx = torch.empty(n_src, h, device=cuda)
    ew2 = torch.randn(e, device=cuda)
    rs2 = torch.empty(n_src, device=cuda)
    _lib.spmm(bwd, g, dx, h, ew=w_csr if weighted else None, wperm=t_eperm, ew2=ew2, rowsum2=rs2)
    msg = g.double()[d] * (w_coo.double()[:, None] if weighted else 1.0)
    ref_dx = torch.zeros(n_src, h, device=cuda, dtype=torch.float64).index_add_(0, s, msg)
    mag_dx = torch.zeros(n_src, h, device=cuda, dtype=torch.float64).index_add_(0, s, msg.abs())
    assert close(dx, ref_dx, mag_dx)
    ew2_coo = torch.empty(e, device=cuda, dtype=torch.float64)
    ew2_coo[eperm.long()] = ew2.double()
    ref_rs = torch.zeros(n_src, device=cuda, dtype=torch.float64).index_add_(0, s, ew2_coo)
    mag_rs = torch.zeros(n_src, device=cuda, dtype=torch.float64).index_add_(0, s, ew2_coo.abs())
    assert close(rs2, ref_rs, mag_rs)
    # binned row sums: bins by (column index % 3)
    rs3 = torch.empty(n_src, 3, device=cuda)
    _lib.spmm(bwd, g, dx, h, wperm=t_eperm, ew2=ew2, rowsum2=rs3, bins=3)
    ref3 = torch.zeros(n_src * 3, device=cuda, dtype=torch.float64).index_add_(0, s * 3 + d % 3, ew2_coo)
    assert close(rs3.view(-1), ref3, mag_rs.repeat_interleave(3))


def test_spmm_is_deterministic_and_l2_schedule_is_invisible(cuda):
    from kgwas_b200 import _lib
    rng = np.random.default_rng(3)
    src, dst = _rand_coo(rng, 5000, 40, 200000, hub=True)
    fwd, *_ = _lib.csr_build(torch.from_numpy(src).to(cuda), torch.from_numpy(dst).to(cuda), 5000, 40, transposed=False,
                             sort_cols=True)
    x = torch.randn(5000, 128, device=cuda)
    outs = [_lib.spmm(fwd, x, torch.empty(40, 128, device=cuda), 128).clone() for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    # window scheduling only permutes which warp runs which segment: results are bit-identical
    fwd.schedule_for_l2(512, window_bytes=64 * 512)
    assert fwd.hseg_order is not None and sorted(fwd.hseg_order.tolist()) == list(range(fwd.n_hsegs))
    out2 = _lib.spmm(fwd, x, torch.empty(40, 128, device=cuda), 128)
    assert torch.equal(outs[0], out2)


@pytest.mark.parametrize("m,n,k", [(1, 4, 4), (130, 128, 128), (1000, 768, 128), (777, 128, 768), (5, 32, 64)])
def test_gemm_nt_nn(cuda, m, n, k):
    from kgwas_b200 import _lib
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device=cuda)
    b_nt = torch.randn(n, k, device=cuda)
    bias = torch.randn(n, device=cuda)
    c = torch.randn(m, n, device=cuda)
    c0 = c.clone()
    _lib.gemm(_lib.KGB_NT, a, b_nt, c, m, n, k, alpha=0.5, beta=1.0, bias=bias, relu=True)
    ref = (0.5 * a.double() @ b_nt.double().T + c0.double() + bias.double()).clamp(min=0)
    assert torch.allclose(c.double(), ref, rtol=1e-5, atol=1e-4 * k ** 0.5)
    b_nn = torch.randn(k, n, device=cuda)
    c = torch.empty(m, n, device=cuda)
    _lib.gemm(_lib.KGB_NN, a, b_nn, c, m, n, k)
    assert torch.allclose(c.double(), a.double() @ b_nn.double(), rtol=1e-5, atol=1e-4 * k ** 0.5)


@pytest.mark.parametrize("rows,m,n", [(1, 32, 32), (100, 128, 128), (50001, 128, 128), (20371, 768, 128), (3, 128, 768)])
def test_gemm_tn_splitk(cuda, rows, m, n):
    from kgwas_b200 import _lib
    torch.manual_seed(rows)
    a = torch.randn(rows, m, device=cuda)
    b = torch.randn(rows, n, device=cuda)
    c = torch.empty(m, n, device=cuda)
    _lib.gemm(_lib.KGB_TN, a, b, c, m, n, rows)
    ref = a.double().T @ b.double()
    assert torch.allclose(c.double(), ref, rtol=1e-5, atol=2e-5 * rows ** 0.5 + 1e-5)
    c2 = torch.empty(m, n, device=cuda)
    _lib.gemm(_lib.KGB_TN, a, b, c2, m, n, rows)
    assert torch.equal(c, c2)                               # split-K reduction order is fixed


def test_gemm_strided_views(cuda):
    """Operands are views into wider buffers (row stride > width), as the engine uses them."""
    from kgwas_b200 import _lib
    torch.manual_seed(0)
    big = torch.randn(300, 3 * 128, device=cuda)
    a = big[:, 128:256]
    w = torch.randn(128, 128, device=cuda)
    out = torch.zeros(300, 2 * 128, device=cuda)
    _lib.gemm(_lib.KGB_NT, a, w, out[:, 128:], 300, 128, 128)
    assert torch.allclose(out[:, 128:].double(), a.double() @ w.double().T, rtol=1e-5, atol=1e-3)
    assert out[:, :128].abs().max() == 0


def test_elementwise_helpers(cuda):
    from kgwas_b200 import _lib
    torch.manual_seed(1)
    for n in (1, 3, 4, 1000, 128 * 1001 + 3):
        dy, y = torch.randn(n, device=cuda), torch.randn(n, device=cuda)
        assert torch.equal(_lib.relu_bwd(dy, y), dy * (y > 0))
    for h in (32, 128, 256):
        x = torch.randn(5003, h, device=cuda)
        out = torch.empty(h, device=cuda)
        _lib.wcolsum(x, h, out)
        assert torch.allclose(out.double(), x.double().sum(0), rtol=1e-5, atol=1e-3)
        w = torch.randn(5003, 8, device=cuda)
        out6 = torch.zeros(6, h, device=cuda)
        _lib.wcolsum(x, h, out6, w=w, n_slots=6)
        assert torch.allclose(out6.double(), w[:, :6].double().T @ x.double(), rtol=1e-5, atol=1e-3)
        v = torch.randn(6, h, device=cuda)
        a = torch.zeros(5003, 8, device=cuda)
        _lib.rowdot(x, v, a, h, 6, 0)
        assert torch.allclose(a[:, :6].double(), x.double() @ v.double().T, rtol=1e-5, atol=1e-3)
        assert a[:, 6:].abs().max() == 0
        xs = torch.randn(301, 6 * h, device=cuda)
        a2 = torch.zeros(301, 6, device=cuda)
        _lib.rowdot(xs, v, a2, h, 6, h)
        ref = (xs.double().view(301, 6, h) * v.double()).sum(-1)
        assert torch.allclose(a2.double(), ref, rtol=1e-5, atol=1e-3)
        yy = torch.randn(5003, h, device=cuda)
        y0 = yy.clone()
        _lib.rank_update(w, v, yy, h, 6, 1.0)
        assert torch.allclose(yy.double(), y0.double() + w[:, :6].double() @ v.double(), rtol=1e-5, atol=1e-3)
    w = torch.randn(1000, device=cuda)
    p = torch.randperm(1000, device=cuda).to(torch.int32)
    assert torch.equal(_lib.permute_f32(w, p), w[p.long()])


@pytest.mark.parametrize("h", [128, 256])
def test_spmm_dot_epilogue(cuda, h):
    """dot_out[i] = <row i just written, dot_w> for ordinary rows, heavy (segmented) rows and empty rows."""
    from kgwas_b200 import _lib
    rng = np.random.default_rng(7 + h)
    torch.manual_seed(h)
    n_src, n_dst, e = 500, 260, 20000
    src, dst = _rand_coo(rng, n_src, n_dst, e, hub=True)
    dst[dst == 9] = 10
    s, d = torch.from_numpy(src).to(cuda), torch.from_numpy(dst).to(cuda)
    fwd, eperm, _, _ = _lib.csr_build(s, d, n_src, n_dst, seg_len=64)
    assert fwd.n_hsegs > 0
    x = torch.randn(n_src, h, device=cuda)
    w_csr = torch.rand(e, device=cuda)
    bias = torch.randn(h, device=cuda)
    wv = torch.randn(1, h, device=cuda)
    y0 = torch.randn(n_dst, h, device=cuda)
    y_plain = _lib.spmm(fwd, x, y0.clone(), h, ew=w_csr, beta=1.0, bias=bias, relu=True)
    y, dots = y0.clone(), torch.full((n_dst, 1), 3.0, device=cuda)
    _lib.spmm(fwd, x, y, h, ew=w_csr, beta=1.0, bias=bias, relu=True, dot_w=wv, dot_out=dots)
    assert torch.equal(y, y_plain)                                 # the epilogue does not change the rows
    ref = y.double() @ wv.double().T
    assert torch.allclose(dots.double(), ref, rtol=1e-5, atol=1e-4 * ref.abs().max().item())


@pytest.mark.parametrize("h", [32, 128, 256])
@pytest.mark.parametrize("m", [1, 17, 5003])
def test_relu_bwd_fused(cuda, h, m):
    """g = scale * (y > 0) * (dy + dp (x) wv), sums = [colsum(g), sum_i dp_i y_i]: every nullable combination."""
    from kgwas_b200 import _lib
    torch.manual_seed(m + h)
    dy, y = torch.randn(m, h, device=cuda), torch.randn(m, h, device=cuda)
    dp, wv = torch.randn(m, 1, device=cuda), torch.randn(1, h, device=cuda)
    for use_dy, use_y, use_dp, scale in [(True, True, False, 1.0), (False, True, True, 1.0), (True, True, True, 0.5),
                                         (True, False, False, 2.0)]:
        g = torch.full((m, h), 9.0, device=cuda)
        sums = torch.full((2, h), 9.0, device=cuda)
        _lib.relu_bwd_fused(g, h, dy=dy if use_dy else None, y=y if use_y else None, dp=dp if use_dp else None,
                            wv=wv if use_dp else None, scale=scale, sums=sums)
        t = torch.zeros(m, h, device=cuda, dtype=torch.float64)
        if use_dy:
            t += dy.double()
        if use_dp:
            t += dp.double() * wv.double()
        ref = scale * t * ((y > 0).double() if use_y else 1.0)
        tol = 1e-5 * max(1.0, ref.abs().max().item())
        assert torch.allclose(g.double(), ref, rtol=1e-5, atol=tol)
        assert torch.allclose(sums[0].double(), ref.sum(0), rtol=1e-5, atol=1e-4 * max(1.0, ref.sum(0).abs().max().item()))
        if use_dp and use_y:
            ref1 = (dp.double() * y.double()).sum(0)
            assert torch.allclose(sums[1].double(), ref1, rtol=1e-5, atol=1e-4 * max(1.0, ref1.abs().max().item()))
        g2 = torch.empty_like(g)
        _lib.relu_bwd_fused(g2, h, dy=dy if use_dy else None, y=y if use_y else None, dp=dp if use_dp else None,
                            wv=wv if use_dp else None, scale=scale)      # without the sums
        assert torch.equal(g, g2)
    if m > 1:   # pure mask case is bit-exact and strided inputs are honoured
        wide = torch.randn(m, 2 * h, device=cuda)
        g = torch.empty(m, h, device=cuda)
        _lib.relu_bwd_fused(g, h, dy=wide[:, h:], y=y)
        assert torch.equal(g, wide[:, h:] * (y > 0))

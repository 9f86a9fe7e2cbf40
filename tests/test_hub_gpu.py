"""GPU: the hub-tile path of kgb_spmm (csrc/kgb_spmm_hub.cuh: hub rows reduced from TMA-staged shared-memory tiles, the
rest through the pull kernel) against fp64 and against the pull-only result of the same CSR -- hub rows cut into parts,
ragged last tile, epilogue (beta, bias, ReLU, dot product) on hub rows, run-to-run bit-reproducibility."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _graph(n_rows, n_cols, n_edges, seed, dev):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n_cols, (n_edges,), generator=g)
    u = torch.rand(n_edges, generator=g)
    dst = (n_rows * u ** 4).long().clamp_(max=n_rows - 1)           # a few rows take most of the edges
    return src.to(dev), dst.to(dev)


@pytest.mark.parametrize("h,n_rows,n_cols,n_edges,tile_rows,n_cta", [
    (128, 700, 5000, 90000, None, None), (128, 300, 4099, 60000, 64, 5), (256, 500, 3000, 50000, None, 7),
    (128, 2000, 20011, 400000, 128, None)])
def test_hub_tiles_match_pull_kernel_and_fp64(cuda, h, n_rows, n_cols, n_edges, tile_rows, n_cta):
    from kgwas_b200 import _lib
    src, dst = _graph(n_rows, n_cols, n_edges, 3 + h, cuda)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n_cols, h, generator=g).to(cuda)
    csr, eperm, _, _ = _lib.csr_build(src, dst, n_cols, n_rows, transposed=False, sort_cols=True)
    ew = torch.rand(n_edges, generator=g).to(cuda)[eperm.long()].contiguous()
    y_pull = _lib.spmm(csr, x, torch.empty(n_rows, h, device=cuda), h, ew=ew)
    assert csr.build_hub(ew, h, min_table_bytes=0, tile_rows=tile_rows, n_cta=n_cta)
    hub = csr.hub
    assert hub.n >= 1 and hub.nv >= hub.n and int(hub.hitem_tail.size(0)) < csr.n_hsegs
    y_hub = _lib.spmm(csr, x, torch.empty(n_rows, h, device=cuda), h, ew=ew)
    rows = torch.repeat_interleave(torch.arange(n_rows, device=cuda), (csr.rowptr[1:] - csr.rowptr[:-1]).long())
    ref = torch.zeros(n_rows, h, dtype=torch.float64, device=cuda).index_add_(
        0, rows, x.double()[csr.col.long()] * ew.double()[:, None])
    scale = ref.abs().max().item()
    assert (y_pull.double() - ref).abs().max().item() <= 2e-6 * scale
    assert (y_hub.double() - ref).abs().max().item() <= 2e-6 * scale
    hub_rows = hub.row.long()
    assert not torch.equal(y_hub[hub_rows], torch.zeros_like(y_hub[hub_rows]))
    other = torch.ones(n_rows, dtype=torch.bool, device=cuda)
    other[hub_rows] = False
    assert torch.equal(y_hub[other], y_pull[other])                 # non-hub rows: same kernel, same order
    # bit-reproducible
    y_again = _lib.spmm(csr, x, torch.empty(n_rows, h, device=cuda), h, ew=ew)
    assert torch.equal(y_hub, y_again)
    # epilogue on hub rows: beta, bias, relu, dot
    y0 = torch.randn(n_rows, h, generator=g).to(cuda)
    bias = torch.randn(h, generator=g).to(cuda)
    dw = torch.randn(h, generator=g).to(cuda)
    dot = torch.empty(n_rows, device=cuda)
    y2 = _lib.spmm(csr, x, y0.clone(), h, ew=ew, beta=1.0, bias=bias, relu=True, dot_w=dw, dot_out=dot)
    ref2 = (ref + y0.double() + bias.double()).clamp(min=0)
    assert (y2.double() - ref2).abs().max().item() <= 2e-6 * ref2.abs().max().item()
    assert (dot.double() - ref2 @ dw.double()).abs().max().item() <= 1e-5 * (ref2 @ dw.double()).abs().max().item()
    # other weights than the ones the plan was built for: the pull kernel handles every row
    ew2 = (ew * 0.5).contiguous()
    y3 = _lib.spmm(csr, x, torch.empty(n_rows, h, device=cuda), h, ew=ew2)
    assert (y3.double() - 0.5 * ref).abs().max().item() <= 2e-6 * scale


def test_hub_plan_bookkeeping_is_exact(cuda):
    """Every hub edge appears exactly once in the chunks, with its weight, in a slot owned by one warp."""
    from kgwas_b200 import _lib
    import numpy as np
    h, n_rows, n_cols, n_edges = 128, 400, 3001, 70000
    src, dst = _graph(n_rows, n_cols, n_edges, 9, cuda)
    csr, eperm, _, _ = _lib.csr_build(src, dst, n_cols, n_rows, transposed=False, sort_cols=True)
    ew = torch.rand(n_edges, device=cuda)
    assert csr.build_hub(ew, h, min_table_bytes=0, tile_rows=128, n_cta=4)
    hub = csr.hub
    chunks = hub.chunks.cpu().numpy()
    off = (hub.tile_off.cpu().numpy() // 4).astype(np.int64)
    vptr = hub.vptr.cpu().numpy()
    hub_row = hub.row.cpu().numpy()
    slot_hub = np.repeat(np.arange(hub.n), np.diff(vptr))
    rp, col, w = csr.rowptr.cpu().numpy(), csr.col.cpu().numpy(), ew.cpu().numpy()
    seen = {}
    slot_owner = {}
    n_rec = 0
    for t in range(hub.n_tiles):
        hdr = chunks[off[t]:off[t] + 24]
        assert hdr[0] == 0 and (np.diff(hdr[:17]) >= 0).all()
        recs = chunks[off[t] + 24: off[t] + 24 + 2 * hdr[16]].reshape(-1, 2)
        for wi in range(16):
            part = recs[hdr[wi]:hdr[wi + 1]]
            for k, (pk, wb) in enumerate(part):
                slot, loc = (int(pk) >> 8) & 0x7fffff, int(pk) & 0xff
                flush = int(pk) < 0
                last = k == len(part) - 1 or ((int(part[k + 1][0]) >> 8) & 0x7fffff) != slot
                assert flush == last
                assert slot_owner.setdefault(slot, wi) == wi
                key = (hub_row[slot_hub[slot]], t * hub.tile_rows + loc)
                seen.setdefault(key, []).append(np.int32(wb).view(np.float32))
                n_rec += 1
    assert n_rec == hub.n_edges
    for r in hub_row:
        cs, ws = col[rp[r]:rp[r + 1]], w[rp[r]:rp[r + 1]]
        for c in np.unique(cs):
            assert sorted(seen[(r, c)]) == sorted(ws[cs == c].tolist())

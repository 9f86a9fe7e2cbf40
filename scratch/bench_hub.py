"""Hub-tile path vs pull-only path on the two (gene, relation)-row launches of kgwas-synth-v1 (h given), CUDA events,
median of 9.  The pull-only timing passes a CLONE of the edge weights (another pointer: kgb_spmm then ignores the plan)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kgwas_b200 import _lib, make_synth_kg  # noqa: E402
from kgwas_b200.plan import get_plan  # noqa: E402


def timeit(fn, n=9):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    h = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dev = torch.device("cuda:0")
    data = make_synth_kg(1.0, 42, hidden=h).to(dev)
    num_nodes = {t: int(x.size(0)) for t, x in data.x_dict.items()}
    plan = get_plan(data.edge_index_dict, num_nodes)
    xf = [j for j in plan.jobs["SNP"] if j.src_type == "Gene"][0]
    af = [j for j in plan.jobs["Gene"] if j.src_type == "SNP"][0]
    n_snp = num_nodes["SNP"]
    torch.manual_seed(0)
    x_snp = torch.randn(n_snp, h, device=dev)
    for name, csr, w in (("af_fwd (gene,k rows <- x_snp)", af.csr, af.w_mean), ("xf_bwd (gene,k rows <- g_snp)", xf.tcsr, xf.w_mean_t)):
        y = torch.empty(csr.n_rows, h, device=dev)
        wc = w.clone()
        csr.schedule_for_l2(4 * h)
        for kw in ({}, {"max_slots": 96}, {"max_slots": 128}, {"tile_rows": 64}, {"tile_rows": 96}):
            csr.hub = None
            ok = csr.build_hub(w, h, **kw)
            if not ok:
                print(name, kw, "no plan")
                continue
            hub = csr.hub
            t_hub = timeit(lambda: _lib.spmm(csr, x_snp, y, h, ew=w))
            y1 = y.clone()
            t_pull = timeit(lambda: _lib.spmm(csr, x_snp, y, h, ew=wc))
            err = float((y - y1).abs().max() / y.abs().max())
            print(f"{name} {kw}: hub path {t_hub:.1f} us, pull-only {t_pull:.1f} us; hubs {hub.n}, slots {hub.nv}, tile {hub.tile_rows}, "
                  f"hub edges {hub.n_edges} of {csr.n_edges} ({hub.n_edges / csr.n_edges:.3f}), chunk_cap {hub.chunk_cap}, "
                  f"tail items {int(hub.hitem_tail.size(0))} of {csr.n_hsegs}, warp load max/mean {hub.warp_loads.max() / hub.warp_loads.mean():.3f}, "
                  f"max diff {err:.1e}", flush=True)


if __name__ == "__main__":
    main()

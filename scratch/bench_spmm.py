"""Micro-benchmark of the four big gather-reduce launches of one SAGE layer on kgwas-synth-v1 (h=128), per kernel
variant (KGB_SPMM_VARIANT).  `--ncu V[,V..]`: run only xf_fwd / xf_bwd once per listed variant inside a
cudaProfilerStart/Stop range (for `ncu --profile-from-start off`)."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kgwas_b200 import _lib, make_synth_kg  # noqa: E402
from kgwas_b200.plan import get_plan  # noqa: E402


def main():
    ncu = None
    if "--ncu" in sys.argv:
        ncu = [int(v) for v in sys.argv[sys.argv.index("--ncu") + 1].split(",")]
    variants = [0, 1, 2, 3, 4]
    if "--variants" in sys.argv:
        variants = [int(v) for v in sys.argv[sys.argv.index("--variants") + 1].split(",")]
    h = 128
    dev = torch.device("cuda:0")
    data = make_synth_kg(1.0, 42, hidden=h).to(dev)
    num_nodes = {t: int(x.size(0)) for t, x in data.x_dict.items()}
    plan = get_plan(data.edge_index_dict, num_nodes)
    xf = [j for j in plan.jobs["SNP"] if j.src_type == "Gene"][0]
    af = [j for j in plan.jobs["Gene"] if j.src_type == "SNP"][0]
    gg = [j for j in plan.jobs["Gene"] if j.src_type == "Gene"][0]
    for j in (xf, af, gg):
        j.schedule(h)
        print(j, flush=True)
    n_snp, n_gene = num_nodes["SNP"], num_nodes["Gene"]
    torch.manual_seed(0)
    Z = torch.randn(xf.n_src * xf.R, h, device=dev)
    out = torch.randn(n_snp, h, device=dev)
    x_snp = torch.randn(n_snp, h, device=dev)
    A = torch.empty(af.n_dst * af.R, h, device=dev)
    g = torch.randn(n_snp, h, device=dev)
    dz = torch.empty(xf.n_src * xf.R, h, device=dev)
    dA = torch.randn(af.n_dst * af.R, h, device=dev)
    dx = torch.randn(n_snp, h, device=dev)
    Zg = torch.randn(gg.n_src * gg.R, h, device=dev)
    og = torch.randn(n_gene, h, device=dev)
    ops = {
        "xf_fwd (SNP rows <- Z)": lambda: _lib.spmm(xf.csr, Z, out, h, ew=xf.w_mean, beta=1.0, relu=True),
        "af_fwd (gene,k rows <- x_snp)": lambda: _lib.spmm(af.csr, x_snp, A, h, ew=af.w_mean),
        "xf_bwd (gene,k rows <- g_snp)": lambda: _lib.spmm(xf.tcsr, g, dz, h, ew=xf.w_mean_t),
        "af_bwd (SNP rows <- dA)": lambda: _lib.spmm(af.tcsr, dA, dx, h, ew=af.w_mean_t, beta=1.0),
        "gg_fwd (gene rows <- Zg)": lambda: _lib.spmm(gg.csr, Zg, og, h, ew=gg.w_mean, beta=1.0),
    }
    if ncu is not None:
        for v in ncu:
            os.environ["KGB_SPMM_VARIANT"] = str(v)
            for name in list(ops)[:4]:
                ops[name]()
            torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for v in ncu:
            os.environ["KGB_SPMM_VARIANT"] = str(v)
            for name in list(ops)[:4]:
                ops[name]()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    ref = {}
    for v in variants:
        os.environ["KGB_SPMM_VARIANT"] = str(v)
        for name, f in ops.items():
            try:
                for _ in range(3):
                    f()
                torch.cuda.synchronize()
                ts = []
                for _ in range(7):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    f()
                    b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) * 1e3)
                ts.sort()
                print(f"variant {v}  {name:32s} median {ts[3]:8.1f} us   min {ts[0]:8.1f} us", flush=True)
            except Exception:
                traceback.print_exc()
        # results must not depend on the variant (same summation order)
        os.environ["KGB_SPMM_VARIANT"] = str(v)
        chk = _lib.spmm(xf.tcsr, g, torch.empty_like(dz), h, ew=xf.w_mean_t)
        chk2 = _lib.spmm(xf.csr, Z, torch.zeros_like(out), h, ew=xf.w_mean)
        if not ref:
            ref = {"a": chk.clone(), "b": chk2.clone()}
        else:
            print(f"variant {v}: bit-identical to variant {variants[0]}: {torch.equal(chk, ref['a'])} {torch.equal(chk2, ref['b'])}")


if __name__ == "__main__":
    main()

"""Size-independent properties of the hot path at BASELINE.json's FULL size (kgwas-synth-v1, scale 1.0: 784 256 SNP,
18.4 M typed edges, h = 128) on the GPU -- the oracle cannot follow at this size, so parity is asserted through
invariants: exact integer bookkeeping, the mean aggregator's checksum (weights of every non-empty group sum to 1),
linearity, the adjoint identity between the forward gather-reduce and its backward (transposed) one, and run-to-run
bit-reproducibility.  The same body runs on the CPU at a small scale with the stand-in kernels (it then checks the
test itself and the plan logic).  Named ``zz`` so that it runs after every other test file."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _cpu_kernels  # noqa: E402


def _check_properties(scale, dev, h=128):
    from kgwas_b200 import _lib, make_synth_kg
    from kgwas_b200.plan import get_plan
    data = make_synth_kg(scale, 42, hidden=h).to(dev) if dev.type == "cuda" else make_synth_kg(scale, 42, hidden=h)
    num_nodes = {t: int(x.size(0)) for t, x in data.x_dict.items()}
    plan = get_plan(data.edge_index_dict, num_nodes)
    n_edges = sum(int(ei.size(1)) for ei in data.edge_index_dict.values())
    assert plan.n_edges == n_edges
    gen = torch.Generator().manual_seed(1)
    for T, jobs in plan.jobs.items():
        for job in jobs:
            job.schedule(h)
            csr, tcsr = job.csr, job.tcsr
            E = job.n_edges
            # ---- integer bookkeeping, exact
            for c in (csr, tcsr):
                rp = c.rowptr.long()
                assert rp.numel() == c.n_rows + 1 and int(rp[0]) == 0 and int(rp[-1]) == E
                assert bool((rp[1:] >= rp[:-1]).all())
                if E:
                    assert int(c.col.min()) >= 0 and int(c.col.max()) < c.n_cols
            assert torch.equal(torch.sort(job.eperm.long())[0], torch.arange(E, device=job.eperm.device))
            assert torch.equal(torch.sort(job.t_eperm.long())[0], torch.arange(E, device=job.eperm.device))
            assert torch.equal(csr.col.long()[job.t_eperm.long()].sort()[0], csr.col.long().sort()[0])
            deg = (csr.rowptr[1:] - csr.rowptr[:-1]).long()
            if E == 0:
                continue
            x = torch.randn(csr.n_cols, h, generator=gen).to(dev)
            x2 = torch.randn(csr.n_cols, h, generator=gen).to(dev)
            g = torch.randn(csr.n_rows, h, generator=gen).to(dev)
            y = _lib.spmm(csr, x, torch.empty(csr.n_rows, h, device=dev), h, ew=job.w_mean)
            # ---- mean checksum: constant features aggregate to that constant (per relation slot of a destination row)
            ones = torch.ones(csr.n_cols, h, device=dev)
            s = _lib.spmm(csr, ones, torch.empty(csr.n_rows, h, device=dev), h, ew=job.w_mean)
            if job.mode == "af":           # rows are (destination, relation) groups: weight sum is 1 or the row is empty
                want = (deg > 0).to(torch.float32)
            else:                          # rows are destinations: one unit per relation slot that has an in-edge
                grp = (job.group_rowptr[1:] - job.group_rowptr[:-1]).long().view(job.n_dst, job.R)
                want = (grp > 0).sum(1).to(torch.float32)
            assert float((s[:, 0] - want).abs().max()) < 2e-3, (T, job)
            assert bool((s[deg == 0] == 0).all())
            # ---- linearity
            y12 = _lib.spmm(csr, 0.5 * x + x2, torch.empty(csr.n_rows, h, device=dev), h, ew=job.w_mean)
            y2 = _lib.spmm(csr, x2, torch.empty(csr.n_rows, h, device=dev), h, ew=job.w_mean)
            ref = 0.5 * y + y2
            assert float((y12 - ref).abs().max()) <= 1e-3 * max(1.0, float(ref.abs().max())), (T, job)
            # ---- adjoint: <A x, g> == <x, A^T g> with A^T run on the transposed CSR (the backward launch)
            atg = _lib.spmm(tcsr, g, torch.empty(csr.n_cols, h, device=dev), h, ew=job.w_mean_t)
            lhs = float((y.double() * g.double()).sum())
            rhs = float((x.double() * atg.double()).sum())
            norm = float(y.double().norm() * g.double().norm()) + 1e-30
            assert abs(lhs - rhs) <= 1e-4 * norm, (T, job, lhs, rhs)
            # ---- bit-reproducible
            y_again = _lib.spmm(csr, x, torch.empty(csr.n_rows, h, device=dev), h, ew=job.w_mean)
            assert torch.equal(y, y_again), (T, job)
    return plan


def test_properties_small_scale_with_cpu_stand_ins(monkeypatch):
    _cpu_kernels.install(monkeypatch)
    _check_properties(0.01, torch.device("cpu"), h=32)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_properties_at_full_size_on_gpu(cuda):
    from kgwas_b200 import plan as _plan
    _plan.clear_plan_cache()
    p = _check_properties(1.0, cuda)
    assert p.n_edges == 18400753 and p.num_nodes["SNP"] == 784256      # kgwas-synth-v1, coalesced
    _plan.clear_plan_cache()
    torch.cuda.empty_cache()

"""GPU: SNP-sharded hetero-GAT (cross-rank softmax over SNP -> Gene groups, owned-row bias, un-fused ReLU) reproduces
the single-GPU logits and parameter gradients -- 2 NCCL ranks when 2 GPUs are visible, else a 1-rank group through the
same code path.  The host logic is covered on 2 gloo ranks in tests/test_dist_cpu.py.  Runs after the other test files."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_dist_gpu import _free_port, _worker  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(300)
@pytest.mark.parametrize("h", [64, 128])
def test_sharded_gat_matches_single(cuda, h):
    import torch.multiprocessing as mp
    world = 2 if torch.cuda.device_count() >= 2 else 1
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret, h, "GAT")) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=200)
    for p in procs:
        if p.is_alive():
            p.terminate()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in range(world):
        err, gerr = ret[r]
        assert err < 1e-4 and gerr < 2e-3, (r, err, gerr)

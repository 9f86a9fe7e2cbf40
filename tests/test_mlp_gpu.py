"""GPU: the input MLPs (kgwas/model.py:10-22) on the engine's GEMM -- forward, input gradient
and every parameter gradient against an fp64 torch.nn reference with the same weights, at the fast-mode raw widths
(SNP 20, Gene 5120, GO 128) and one width the GEMM cannot take (70: falls back to torch, same results)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,in_dim,hid", [(3001, 20, 128), (2037, 5120, 128), (777, 128, 128), (1500, 128, 256), (900, 70, 128)])
def test_simple_mlp_matches_fp64(cuda, n, in_dim, hid):
    import kgwas_b200
    from kgwas_b200 import _lib
    torch.manual_seed(n + in_dim)
    mlp = kgwas_b200.SimpleMLP(in_dim, hid, hid)
    ref = torch.nn.Sequential(torch.nn.Linear(in_dim, hid), torch.nn.ReLU(), torch.nn.Linear(hid, hid), torch.nn.ReLU(),
                              torch.nn.Linear(hid, hid)).double()
    with torch.no_grad():
        for src, dst in ((mlp.FC_hidden, ref[0]), (mlp.FC_hidden2, ref[2]), (mlp.FC_output, ref[4])):
            dst.weight.copy_(src.weight.double())
            dst.bias.copy_(src.bias.double())
    assert list(mlp.state_dict().keys()) == ["FC_hidden.weight", "FC_hidden.bias", "FC_hidden2.weight", "FC_hidden2.bias",
                                             "FC_output.weight", "FC_output.bias"]
    mlp = mlp.to(cuda)
    x = torch.randn(n, in_dim)
    # keep every hidden pre-activation away from the ReLU kink: an element within fp32 rounding of 0 may fall on the other
    # side in fp32 than in fp64 and then moves a whole gradient row by O(1e-3) -- a property of ReLU, not of the kernels
    for _ in range(50):
        with torch.no_grad():
            p1 = ref[0](x.double())
            p2 = ref[2](p1.relu())
        bad = ((p1.abs() < 1e-3).any(1) | (p2.abs() < 1e-3).any(1)).nonzero().reshape(-1)
        if bad.numel() == 0:
            break
        x[bad] = torch.randn(bad.numel(), in_dim)
    assert bad.numel() == 0
    up = torch.randn(n, hid)
    xc = x.to(cuda).requires_grad_()
    k0 = _lib.kernel_launch_count()
    out = mlp(xc)
    out.backward(up.to(cuda))
    used_engine = _lib.kernel_launch_count() > k0
    assert used_engine == (in_dim % 4 == 0)
    xr = x.double().requires_grad_()
    outr = ref(xr)
    outr.backward(up.double())

    def err(a, b):
        return ((a.detach().cpu().double() - b.detach()).abs().max() / b.detach().abs().max()).item()
    # fp32-class accuracy (3xTF32 products, fp32 accumulation; K up to 5120): an order below north_star's 1e-4
    assert err(out, outr) < 2e-5, err(out, outr)
    assert err(xc.grad, xr.grad) < 1e-4, err(xc.grad, xr.grad)
    for (a, b) in ((mlp.FC_hidden, ref[0]), (mlp.FC_hidden2, ref[2]), (mlp.FC_output, ref[4])):
        assert err(a.weight.grad, b.weight.grad) < 1e-4, err(a.weight.grad, b.weight.grad)
        assert err(a.bias.grad, b.bias.grad) < 1e-4, err(a.bias.grad, b.bias.grad)

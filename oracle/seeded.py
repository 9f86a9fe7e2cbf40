"""TEST INFRASTRUCTURE ONLY: deterministic parameter / feature values shared by the golden-fixture generator
(oracle/gen_golden_from_reference.py, which fills the REFERENCE's modules with them) and the parity tests (which fill
the oracle and the CUDA modules with the same values).  Keeping multi-megabyte weight tensors out of tests/golden/
is the point: a fixture stores edge lists, outputs and gradient samples only."""
import zlib

import torch


def seeded_tensor(name: str, shape, seed: int, scale: float) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 1000003 * seed) % (2 ** 31))
    return torch.randn(tuple(shape), generator=g) * scale


def fill_parameters(module: torch.nn.Module, seed: int):
    """Every materialised parameter <- N(0, 1/fan_in) (matrices) or N(0, 0.1^2) (biases, attention vectors), keyed by
    its state-dict name: independent of construction / materialisation order."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if isinstance(p, torch.nn.parameter.UninitializedParameter):
                continue
            if p.dim() == 2:
                scale = 1.0 / (p.size(1) ** 0.5)
            elif p.dim() == 3:                      # att_src / att_dst (1, heads, C)
                scale = 1.0 / (p.size(-1) ** 0.5)
            else:
                scale = 0.1
            p.copy_(seeded_tensor(name, p.shape, seed, scale))
    return module


def canonical_key(name: str) -> str:
    """State-dict key with PyG >= 2.4 HeteroConv spelling ('<a___b___c>') folded to the PyG <= 2.3 one ('a__b__c'),
    so that values do not depend on which spelling a module uses."""
    if ".convs.<" in name:
        head, _, rest = name.partition(".convs.<")
        key, _, tail = rest.partition(">")
        return head + ".convs." + key.replace("___", "__").replace("#", ".") + tail
    return name


def grad_sample(g: torch.Tensor, stride: int = 61):
    """Compact fingerprint of a gradient tensor: strided sample + sum + abs-max."""
    f = g.detach().reshape(-1).double()
    return {"sample": f[::stride].float().clone(), "sum": float(f.sum()), "absmax": float(f.abs().max()),
            "numel": int(f.numel())}

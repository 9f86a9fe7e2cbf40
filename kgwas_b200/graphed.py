"""CUDA-graph capture of a full-graph training step.

The fused layer launches ~170 kernels per step (two streams of big gather-reduces / SNP-row GEMMs and several dozen
gene / GO sized ones); on a B200 they take ~5 ms of GPU time while Python + ctypes + autograd need more than 6 ms to
issue them, so an eagerly driven step is bound by the host.  For the full-graph regime (fixed ``edge_index_dict``, fixed
shapes: every optimiser step sees the same graph -- kgwas/kgwas.py:126-151 with the whole KG as one batch) the step is
captured ONCE, multi-stream scheduling and all (the scheduler's fork / join of its side streams is a legal capture
topology), and replayed with a single launch.  Mini-batch training (``NeighborLoader`` batches of varying shape) stays
eager.  The same kernels run in the same order on the same buffers: results are bit-identical to the eager step.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch


class GraphedStep:
    """``step_fn(x_dict) -> (pred, loss)`` (forward + backward + optimiser step, reading the static tensors in
    ``x_static``) captured in a CUDA graph.

    ``__call__(x_dict=None)`` copies new feature values into the static inputs (device-to-device, stream ordered) when
    given, replays the graph and returns the static ``(pred, loss)`` tensors (overwritten by the next replay).
    The optimiser must have been built with ``capturable=True``."""

    def __init__(self, step_fn: Callable, x_static: Dict[str, torch.Tensor], warmup: int = 3):
        self.x_static = x_static
        dev = next(iter(x_static.values())).device
        self.stream = torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream(dev)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):              # warm-up ON the capture stream: plans, workspaces, side
            for _ in range(max(1, warmup)):               # streams and optimiser state are created here, not in capture
                step_fn(x_static)
        cur.wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        from . import _lib
        k0 = _lib.kernel_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.pred, self.loss = step_fn(x_static)
        self.kernels_per_replay = _lib.kernel_launch_count() - k0     # libkgwas_b200 kernels inside one replay

    def __call__(self, x_dict: Optional[Dict[str, torch.Tensor]] = None):
        if x_dict is not None:
            with torch.no_grad():
                for k, v in x_dict.items():
                    if v is not self.x_static[k]:
                        self.x_static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.pred, self.loss

"""CUDA-graph capture of one head training step.

The eager step (forward + `calculate_loss` + backward) issues ~20 small kernels and a lot of Python /
autograd bookkeeping around four large ones; on B200 the host side is then the bottleneck (the kernels of a
cub27 batch-64 step take ~0.6 ms, the Python around them >2 ms).  For fixed shapes the whole step is captured
once into a CUDA graph and replayed: same kernels, same results, no host work per step.  This is the
B200-native replacement for the reference's eager per-node loops (`pipnet/train.py:229-264`).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class GraphedHeadStep:
    """Capture `fn(x, ys) -> loss` (+ backward) for static shapes.

    fn must be capture-safe: no host synchronisation, no `.item()`; every tensor it creates comes from the
    torch caching allocator (our C ABI never allocates).  After `replay(x, ys)`: `self.loss` (0-dim),
    `self.grad_x` and every parameter's `.grad` hold the step's results in static storage.
    """

    def __init__(self, fn: Callable, params, x_example: torch.Tensor, ys_example: torch.Tensor, warmup: int = 3,
                 need_grad_x: bool = True):
        self.params = [p for p in params]
        self.static_x = x_example.detach().clone().requires_grad_(need_grad_x)
        self.static_y = ys_example.detach().clone()
        self.loss: Optional[torch.Tensor] = None
        self.grad_x: Optional[torch.Tensor] = None
        from . import ops
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # the captured step owns its gradient storage: head-parameter gradients go through the flat bucket
        # (ops.local_grad_bucket: no per-producer zero fills, no autograd add of the orth gradient and dW)
        one = self._one = torch.ones((), device=x_example.device)      # root gradient: static, instead of a ones_like fill per step
        with ops.local_grad_bucket():
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._zero()
                    fn(self.static_x, self.static_y).backward(gradient=one)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._zero()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = fn(self.static_x, self.static_y)
                self.loss.backward(gradient=one)
        self.grad_x = self.static_x.grad

    def _zero(self):
        for p in self.params:
            p.grad = None
        self.static_x.grad = None

    def replay(self, x: Optional[torch.Tensor] = None, ys: Optional[torch.Tensor] = None):
        if x is not None:
            self.static_x.detach().copy_(x, non_blocking=True)
        if ys is not None:
            self.static_y.copy_(ys, non_blocking=True)
        self.graph.replay()
        return self.loss, self.grad_x

"""CUDA-graph replay of one log-likelihood + gradient evaluation.

A NUTS leapfrog or an optimiser step evaluates the same computation -- parameters -> model matrices (a few dozen tiny
torch kernels) -> fused forward+adjoint (``liblqgk.so``) -> chain back to the parameters -- thousands of times with the same
shapes.  At small sample counts the host-side launch overhead of the torch glue (~2 ms) is a third of the evaluation; the
whole evaluation is therefore captured ONCE into a CUDA graph and replayed (the library's entry points are capture-safe after
``lqgk_init``: no CUDA object is created inside a call).  In the reference this role is played by ``jax.jit`` of
``value_and_grad`` (numpyro jit-compiles the potential energy).
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch

from lqg_b200 import abi


class GraphedValueAndGrad:
    """``fn(theta) -> ll`` (any shape; its sum is differentiated) captured with its backward pass.

    ``__call__(theta)`` copies ``theta`` into the static input, replays the graph and returns ``(ll, grad)`` -- static
    tensors that the next call overwrites."""

    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor], theta: torch.Tensor, warmup: int = 3):
        if not theta.is_cuda:
            raise RuntimeError("GraphedValueAndGrad needs CUDA tensors")
        abi.load_library().init(1)
        self.theta = theta.detach().clone().requires_grad_()
        cur = torch.cuda.current_stream(theta.device)
        side = torch.cuda.Stream(theta.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                       # warm-up off the capture: workspace, kernel attributes, autotuning
            for _ in range(warmup):
                self.theta.grad = None
                fn(self.theta).sum().backward()
        cur.wait_stream(side)
        torch.cuda.synchronize(theta.device)
        self.theta.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn(self.theta)
            self.out.sum().backward()
        self.grad = self.theta.grad

    def __call__(self, theta: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        with torch.no_grad():
            self.theta.copy_(theta)
        self.graph.replay()
        return self.out, self.grad

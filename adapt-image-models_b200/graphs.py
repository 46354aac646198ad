"""CUDA-graph capture of a whole training (or inference) step.

The step is a fixed sequence of ~650 kernel launches (C-ABI calls + a handful of torch ops for the head,
loss and AdamW); replayed as one graph it is no longer bound by Python / launch latency.  Inputs are copied
into static device buffers before each replay (that copy is the step's H2D transfer in the e2e timing)."""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    def __init__(self, fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor], warmup: int = 3,
                 before_capture: Callable[[], None] = None):
        """fn(*static_inputs) -> tensor (e.g. the loss); must not synchronise with the host."""
        self.static_in = [torch.empty_like(t) for t in example_inputs]
        for s, t in zip(self.static_in, example_inputs):
            s.copy_(t)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if before_capture is not None:
            before_capture()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        for s, t in zip(self.static_in, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_out

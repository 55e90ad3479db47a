"""CUDA-graph capture of a whole hot-path micro-step.

At the reference's batch sizes (train_hypernet: B=4, K=128; few-shot: B=256) one micro-step is ~60 short kernels, so the cost is
launch latency and Python/autograd overhead, not GPU work.  Every C-ABI entry point of libdmi_b200 is capture-safe (no
allocation, no synchronisation, TMA descriptors are built on the host and baked into the launch), so the complete
forward + backward of a micro-step can be recorded once and replayed with a single ``cudaGraphLaunch``.

    step = GraphedStep(fn, static_inputs={"mm": mm, "R": R, ...})     # fn() reads the static tensors, calls .backward()
    out = step(mm=new_mm, R=new_R, ...)                                # copies into the static buffers, replays

Gradients: parameters that have ``.grad`` allocated before capture are accumulated into in place on every replay (gradient
accumulation across micro-steps, train_hypernet.py:119-149 semantics); zero them with ``p.grad.zero_()`` between optimizer steps.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Optional

import torch


class GraphedStep:
    def __init__(self, fn: Callable[[], object], static_inputs: Dict[str, torch.Tensor], params: Optional[Iterable[torch.nn.Parameter]] = None,
                 accumulate_grads: bool = True, warmup: int = 3):
        self.fn = fn
        self.static_inputs = static_inputs
        self.params = list(params) if params is not None else []
        # Warm-up, gradient probing and capture all run on ONE side stream: autograd's AccumulateGrad nodes remember the stream
        # they were created on, and a node created on the legacy default stream would make the captured backward depend on it.
        # (Drop every reference to earlier autograd graphs of these parameters before building a GraphedStep.)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for p in self.params:
                p.grad = None
            for _ in range(warmup):                 # lazy inits: cudaFuncSetAttribute, operand caches, allocator pools
                fn()
                for p in self.params:
                    p.grad = None
            if accumulate_grads and self.params:
                # one eager run tells which parameters actually receive a gradient (e.g. generators.1 never does in as-written
                # mode); those keep a persistent zeroed .grad so that the captured backward accumulates into it in place
                fn()
                for p in self.params:
                    if p.grad is not None:
                        p.grad.zero_()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.static_output = fn()
        torch.cuda.synchronize()

    def __call__(self, **new_inputs):
        for k, v in new_inputs.items():
            self.static_inputs[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_output

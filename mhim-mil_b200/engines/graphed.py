"""CUDA-graph replay of a whole training (or inference) step.

The per-bag step of the reference's trainer (engines/base_engine.py:78-120: `forward_func` -> loss -> backward) issues 150-180 kernel
launches for 1-2 ms of GPU work at N = 10 000: it is bound by the host (Python + launch latency), not by the GPU
(profiles/round2_train_step_*.txt).  `GraphedStep` captures the step ONCE per input shape into a CUDA graph and replays it: one
launch per step, GPU-bound.

What makes the capture sound here (and not in general):
  * every mhimk kernel is enqueued on torch's current stream through the C ABI, allocates nothing, and takes its sizes from the
    host-side tensor shapes -- which are part of the cache key;
  * mask sizes (`len_keep`, top-k counts, Merge's keep count) are host integers derived from N, never read back from the device;
  * randomness stays fresh across replays: torch's own generators are graph-safe, and the in-kernel dropout switches to
    device-resident (seed, offset) words that a captured generator kernel refreshes (mil_dropout_t mode 3, ops.next_dropout);
  * weight images are rebuilt inside the graph in train mode (`volatile`), so optimizer / EMA updates between replays are seen.
Bags differ in size: one graph per distinct N, all sharing one memory pool (they never run concurrently).
"""
from typing import Callable, Dict, Tuple

import torch


class GraphedStep:
    """step = GraphedStep(fn);  out = step(x, ...)

    `fn(*tensors)` runs one step (e.g. zero_grad(set_to_none=True) -> teacher -> student -> loss -> backward) and returns a tensor
    or a tuple/list of tensors (python scalars are passed through as captured).  Inputs are copied into static buffers; outputs
    are static tensors overwritten by the next replay (clone them to keep).  Gradients live in the graph's pool: read `.grad` after
    the call, step the optimizer outside (or inside `fn` with a capturable optimizer)."""

    def __init__(self, fn: Callable, warmup: int = 3, max_graphs: int = 64):
        self.fn, self.warmup, self.max_graphs = fn, warmup, max_graphs
        self._graphs: Dict[Tuple, Tuple] = {}
        self._pool = None

    def _key(self, args):
        return tuple((tuple(a.shape), a.dtype, a.device) if isinstance(a, torch.Tensor) else ("py", a) for a in args)

    def __call__(self, *args):
        key = self._key(args)
        hit = self._graphs.get(key)
        if hit is None:
            hit = self._capture(key, args)
        graph, static_in, static_out = hit
        for s, a in zip(static_in, args):
            if isinstance(a, torch.Tensor) and s.data_ptr() != a.data_ptr():
                s.copy_(a, non_blocking=True)
        graph.replay()
        return static_out

    def _capture(self, key, args):
        if len(self._graphs) >= self.max_graphs:
            self._graphs.clear()
        static_in = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the capture: lazy initialisations, caches, allocator
            for _ in range(self.warmup):
                self.fn(*static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        with torch.cuda.graph(graph, pool=self._pool):
            static_out = self.fn(*static_in)
        self._graphs[key] = (graph, static_in, static_out)
        return self._graphs[key]

    def buffers(self, *args):
        """The static input buffers of the graph for these shapes (captured on first use): fill them in place (e.g. the H2D copy of the
        next bag) and call the step with them to skip the per-replay input copy."""
        key = self._key(args)
        hit = self._graphs.get(key)
        if hit is None:
            hit = self._capture(key, args)
        return hit[1]

    @property
    def n_graphs(self) -> int:
        return len(self._graphs)

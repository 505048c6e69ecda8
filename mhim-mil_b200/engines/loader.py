"""Bag loader: `.pt` [N, D] fp32 feature files -> pinned host ring -> HBM, one bag ahead of the compute stream, sharded over the
ranks of a node.

Reference: datasets/dataset_feat.py:86-91 (`torch.load(<root>/pt_files/<slide>.pt)` per item; the on-disk layout is what
CLAM/extract_features_fp.py:169-171 writes: one float32 [N, D] tensor per slide) behind a DataLoader(batch_size=1) and
datasets/data_utils.py:484-521 (`PrefetchLoader`: a side stream that moves the next batch while the current one is computed).
The kernels drain a 205 MB bag in ~0.15 ms, so feeding is the end-to-end bound (PCIe: ~3.7 ms per such bag).  What this loader
does about it:
  * a background thread reads files ahead into a ring of PINNED host buffers (pageable memory would halve the PCIe rate and make
    the copy synchronous);
  * the host->device copy of bag i+1 runs on a copy stream while bag i is computed; the compute stream waits on an event only;
  * device buffers are recycled through events (no allocation per bag; variable N uses the front of a max-size buffer);
  * multi-GPU: bags are dealt to ranks in contiguous slices (mhimk.dist.bag_slice) -- bag-parallel, no collective.
Yields (bag [1, N, D] on the device, label int64 [1] on the device, name).
"""
import os
import queue
import threading
from typing import Callable, List, Optional, Sequence, Tuple, Union

import torch

from ..dist import bag_slice

Item = Union[str, torch.Tensor]


def _read(item: Item) -> torch.Tensor:
    if isinstance(item, torch.Tensor):
        t = item
    else:
        try:
            t = torch.load(item, weights_only=True)
        except Exception:                                           # older files pickled with numpy arrays (dataset_feat.py:88-91)
            t = torch.load(item, weights_only=False)
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
    if t.dim() == 3 and t.shape[0] == 1:
        t = t[0]
    if t.dim() != 2:
        raise RuntimeError(f"mhimk BagLoader: a bag must be [N, D] (got {tuple(t.shape)})")
    return t.to(torch.float32)


class BagLoader:
    def __init__(self, items: Sequence[Item], labels: Sequence[int], device=None, rank: int = 0, world: int = 1, ring: int = 3,
                 max_rows: Optional[int] = None, shuffle_seed: Optional[int] = None, reader: Callable[[Item], torch.Tensor] = _read):
        if len(items) != len(labels):
            raise ValueError("BagLoader: items and labels differ in length")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mhimk BagLoader: the target must be a CUDA device -- no CPU path")
        order = list(range(len(items)))
        if shuffle_seed is not None:
            order = torch.randperm(len(items), generator=torch.Generator().manual_seed(shuffle_seed)).tolist()
        mine = [order[i] for i in bag_slice(len(order), rank, world)]
        self.items, self.labels = [items[i] for i in mine], [int(labels[i]) for i in mine]
        self.ring, self.max_rows, self.reader = max(2, ring), max_rows, reader
        self._pinned: List[Optional[torch.Tensor]] = [None] * self.ring
        self._dev: List[Optional[torch.Tensor]] = [None] * 2
        self.h2d_bytes = 0

    def __len__(self):
        return len(self.items)

    def _names(self):
        return [os.path.basename(it) if isinstance(it, str) else f"bag{i}" for i, it in enumerate(self.items)]

    def _fit(self, bufs, slot, numel, make):
        b = bufs[slot]
        if b is None or b.numel() < numel:
            want = max(numel, (self.max_rows or 0))
            bufs[slot] = b = make(want)
        return b

    def __iter__(self):
        n = len(self.items)
        if n == 0:
            return
        names = self._names()
        free: "queue.Queue[int]" = queue.Queue()
        ready: "queue.Queue[Tuple[int, int, Tuple[int, int]]]" = queue.Queue()
        for s in range(self.ring):
            free.put(s)
        err: List[BaseException] = []

        def produce():                                               # file -> pinned ring (host thread; torch.load releases the GIL in IO)
            try:
                for i, it in enumerate(self.items):
                    t = self.reader(it)
                    slot = free.get()
                    pin = self._fit(self._pinned, slot, t.numel(), lambda k: torch.empty(k, dtype=torch.float32).pin_memory())
                    pin[: t.numel()].view(t.shape).copy_(t)
                    ready.put((i, slot, tuple(t.shape)))
            except BaseException as e:                               # surfaced on the consumer side
                err.append(e)
                ready.put((-1, -1, (0, 0)))

        th = threading.Thread(target=produce, daemon=True)
        th.start()
        copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        consumed = [torch.cuda.Event() for _ in range(2)]
        for e in consumed:
            e.record(main)

        def stage(k):                                                # pinned -> device buffer k & 1 on the copy stream
            i, slot, shape = ready.get()
            if i < 0:
                raise err[0]
            numel = shape[0] * shape[1]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[k & 1])              # the compute that read this device buffer two bags ago is done
                dev = self._fit(self._dev, k & 1, numel, lambda m: torch.empty(m, dtype=torch.float32, device=self.device))
                view = dev[:numel].view(1, shape[0], shape[1])
                view.copy_(self._pinned[slot][:numel].view(1, shape[0], shape[1]), non_blocking=True)
                done = torch.cuda.Event()
                done.record(copy_stream)
            self.h2d_bytes += numel * 4
            return view, done, slot

        nxt = stage(0)
        for k in range(n):
            view, done, slot = nxt
            if k + 1 < n:
                nxt = stage(k + 1)                                   # bag k+1 moves while bag k is computed
            main.wait_event(done)
            done.synchronize()                                       # the pinned slot may be refilled once its copy has finished
            free.put(slot)
            label = torch.tensor([self.labels[k]], dtype=torch.int64, device=self.device)
            yield view, label, names[k]
            consumed[k & 1].record(main)
        th.join()

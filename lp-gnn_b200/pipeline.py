"""Host -> device -> host pipeline for streams of LPs (the shape of the reference's prediction sweep,
scripts/pred_basis.py:153-154: for every LP, move it to the GPU, run the model, decide the basis, fetch it).

Each LP travels as ONE pinned, packed host buffer ``[row | col | val | x_s | x_t]`` (4-byte words) -> one H2D
copy; the basis statuses come back as one uint8 D2H copy.  The copy of LP i+1 runs on a side stream while LP i is
computed, so PCIe time hides behind the kernels; every LP's copies still happen inside the loop.
"""
from __future__ import annotations

import types
from dataclasses import dataclass

import numpy as np
import torch

from .graph import BipartiteCSR


@dataclass
class HostLP:
    pack: torch.Tensor          # pinned int32 [words]
    offs: tuple                 # word offsets of row, col, val, x_s, x_t, end
    m: int
    n: int
    p: int
    q: int
    sorted: bool = True

    @property
    def nbytes(self):
        return self.pack.numel() * 4


def pack_lp(row, col, val, x_s, x_t, is_sorted=True, pin=True) -> HostLP:
    """COO (any integer dtype) + features -> one packed (pinned) staging buffer."""
    parts = [np.ascontiguousarray(row, dtype=np.int32), np.ascontiguousarray(col, dtype=np.int32),
             np.ascontiguousarray(val, dtype=np.float32).view(np.int32),
             np.ascontiguousarray(x_s, dtype=np.float32).reshape(-1).view(np.int32),
             np.ascontiguousarray(x_t, dtype=np.float32).reshape(-1).view(np.int32)]
    offs = tuple(np.cumsum([0] + [a.shape[0] for a in parts]).tolist())
    pack = torch.from_numpy(np.concatenate(parts))
    if pin and torch.cuda.is_available():
        pack = pack.pin_memory()
    return HostLP(pack, offs, int(np.shape(x_s)[0]), int(np.shape(x_t)[0]), int(np.shape(x_s)[1]), int(np.shape(x_t)[1]),
                  bool(is_sorted))


def unpack_device(d_pack, lp: HostLP):
    o = lp.offs
    f32 = lambda a, b: d_pack[a:b].view(torch.float32)
    return (d_pack[o[0]:o[1]], d_pack[o[1]:o[2]], f32(o[2], o[3]), f32(o[3], o[4]).view(lp.m, lp.p),
            f32(o[4], o[5]).view(lp.n, lp.q))


class BasisPipeline:
    """Double-buffered predict loop: ``for idx, status in pipe.run(host_lps)`` yields the uint8 status vector
    (constraints first) of every LP, in order.  The yielded array is a view of a pinned slot that is reused two
    LPs later -- copy it if it must outlive the next iteration."""

    def __init__(self, model, device):
        if not torch.cuda.is_available():
            raise RuntimeError("BasisPipeline needs a CUDA device (no CPU fallback)")
        self.model, self.dev = model, torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.d_buf = [None, None]
        self.h_status = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]      # H2D of the slot finished
        self.done = [torch.cuda.Event(), torch.cuda.Event()]       # compute + D2H of the slot finished

    def _copy_in(self, slot, lp: HostLP, first_use):
        words = lp.pack.numel()
        if self.d_buf[slot] is None or self.d_buf[slot].numel() < words:
            self.d_buf[slot] = torch.empty(int(words * 1.25) + 64, dtype=torch.int32, device=self.dev)
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self.done[slot])       # the slot's previous LP has been consumed
            self.d_buf[slot][:lp.pack.numel()].copy_(lp.pack, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    @torch.no_grad()
    def _compute(self, slot, lp: HostLP):
        cur = torch.cuda.current_stream(self.dev)
        # the slot's previous result (LP i-2) was handed out one iteration ago, so its host buffer may be replaced
        nodes = lp.m + lp.n
        if self.h_status[slot] is None or self.h_status[slot].numel() < nodes:
            self.h_status[slot] = torch.empty(int(nodes * 1.25) + 64, dtype=torch.uint8).pin_memory()
        cur.wait_event(self.ready[slot])
        row, col, val, x_s, x_t = unpack_device(self.d_buf[slot], lp)
        st = self.model.predict_basis_coo(row, col, val, lp.m, lp.n, x_s, x_t, is_sorted=lp.sorted)   # one native call
        self.h_status[slot][:lp.m + lp.n].copy_(st, non_blocking=True)
        self.done[slot].record(cur)

    def run(self, host_lps):
        lps = list(host_lps)
        if not lps:
            return
        self._copy_in(0, lps[0], True)
        for i, lp in enumerate(lps):
            slot = i & 1
            if i + 1 < len(lps):
                self._copy_in(slot ^ 1, lps[i + 1], i + 1 < 2)     # prefetch the next LP while this one computes
            self._compute(slot, lp)
            if i > 0:
                self.done[slot ^ 1].synchronize()
                prev = lps[i - 1]
                yield i - 1, self.h_status[slot ^ 1][:prev.m + prev.n].numpy()
        last = len(lps) - 1
        self.done[last & 1].synchronize()
        yield last, self.h_status[last & 1][:lps[last].m + lps[last].n].numpy()

"""Host -> device -> host pipeline for streams of LPs (the shape of the reference's prediction sweep,
scripts/pred_basis.py:153-154: for every LP, move it to the GPU, run the model, decide the basis, fetch it).

Each LP travels as ONE pinned, packed host buffer ``[row | col | val | x_s | x_t]`` (4-byte words) -> one H2D
copy; the basis statuses come back as one uint8 D2H copy.  The copy of LP i+1 runs on a side stream while LP i is
computed, so PCIe time hides behind the kernels; every LP's copies still happen inside the loop.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


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
    """Multi-buffered predict loop: ``for idx, status in pipe.run(host_lps)`` yields the uint8 status vector
    (constraints first) of every LP, in order.  The yielded array is a view of a pinned slot that is reused
    ``depth`` LPs later -- copy it if it must outlive the next iteration."""

    def __init__(self, model, device, compute_streams=3):
        """``compute_streams=k > 1``: k LPs in flight, consecutive LPs on k alternating streams (one staging slot
        each), so the latency-bound small kernels of one LP (graph build, basis selection) overlap the neighbouring
        LPs' work; 1 = strictly one LP at a time on the caller's current stream (two staging slots)."""
        if not torch.cuda.is_available():
            raise RuntimeError("BasisPipeline needs a CUDA device (no CPU fallback)")
        self.model, self.dev = model, torch.device(device)
        self.depth = D = max(2, int(compute_streams))
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.compute = [torch.cuda.Stream(self.dev) for _ in range(D)] if compute_streams > 1 else None
        self.d_buf = [None] * D
        self.h_status = [None] * D
        self.ready = [torch.cuda.Event() for _ in range(D)]        # H2D of the slot finished
        self.done = [torch.cuda.Event() for _ in range(D)]         # compute + D2H of the slot finished

    def _copy_in(self, slot, lp: HostLP, first_use):
        words = lp.pack.numel()
        if self.d_buf[slot] is None or self.d_buf[slot].numel() < words:
            if not first_use:
                # the slot's previous LP may still be computing on another stream: it must be finished before its
                # staging buffer goes back to the allocator (growth is rare, the host wait is not on the steady path)
                self.done[slot].synchronize()
            self.d_buf[slot] = torch.empty(int(words * 1.25) + 64, dtype=torch.int32, device=self.dev)
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self.done[slot])       # the slot's previous LP has been consumed
            self.d_buf[slot][:lp.pack.numel()].copy_(lp.pack, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    @torch.no_grad()
    def _compute(self, slot, lp: HostLP):
        if self.compute is None:
            return self._compute_on(slot, lp, torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.compute[slot]):
            return self._compute_on(slot, lp, self.compute[slot])

    def _compute_on(self, slot, lp: HostLP, cur):
        # the slot's previous result (LP i-depth) was handed out at least one iteration ago, so its host buffer may be replaced
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
        D = self.depth
        # (re)build the cached 16-bit / x2 weight copies on the CALLER's stream before the compute streams fork from it:
        # built lazily inside the first predict call they would be written on one compute stream and read on the others
        self.model._native_weights()
        if self.compute is not None:                               # weights etc. were produced on the caller's stream
            for st in self.compute:
                st.wait_stream(torch.cuda.current_stream(self.dev))
        self._copy_in(0, lps[0], True)
        for i, lp in enumerate(lps):
            if i + 1 < len(lps):
                self._copy_in((i + 1) % D, lps[i + 1], i + 1 < D)  # prefetch the next LP while this one computes
            self._compute(i % D, lp)
            j = i - (D - 1)                                        # oldest LP in flight: its slot is needed next
            if j >= 0:
                self.done[j % D].synchronize()
                yield j, self.h_status[j % D][:lps[j].m + lps[j].n].numpy()
        for j in range(max(0, len(lps) - (D - 1)), len(lps)):
            self.done[j % D].synchronize()
            yield j, self.h_status[j % D][:lps[j].m + lps[j].n].numpy()


class PackedBasisPipeline:
    """Sweep over many small/medium LPs (the reference's pred_basis workload, BASELINE config C5): consecutive LPs
    are packed block-diagonally up to a node / nonzero budget, one pack = one H2D burst (ONE copy per LP: its pinned
    pack goes verbatim into a device staging area; ``lpgnn_pack_scatter`` builds the pack layout from it), ONE native
    forward over the pack and a per-LP (segmented) basis decision, one D2H copy of the pack's statuses.  Packs are double-buffered like ``BasisPipeline``.

    ``for idx, status in pipe.run(host_lps)`` yields a fresh uint8 array [m+n] (constraints first) per LP, in order.
    """

    def __init__(self, model, device, max_nodes=600_000, max_nnz=3_000_000, max_lps=128, compute_streams=1):
        if not torch.cuda.is_available():
            raise RuntimeError("PackedBasisPipeline needs a CUDA device (no CPU fallback)")
        self.model, self.dev = model, torch.device(device)
        self.max_nodes, self.max_nnz, self.max_lps = max_nodes, max_nnz, max_lps
        self.copy_stream = torch.cuda.Stream(self.dev)
        # optional alternating compute streams as in BasisPipeline; off by default: packs already fill the GPU and
        # their varying workspace sizes defeat the per-stream allocator caches (measured 2.7x slower on C5)
        self.compute = [torch.cuda.Stream(self.dev) for _ in range(2)] if compute_streams > 1 else None
        self.d_buf, self.d_stage, self.h_status, self.h_ptr = [None, None], [None, None], [None, None], [None, None]
        self.bufs = [dict(), dict()]          # per slot: grow-only workspace / status buffers of the native call
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]

    def _plan(self, lps):
        packs, cur, nodes, nnz = [], [], 0, 0
        for i, lp in enumerate(lps):
            z = lp.offs[1] - lp.offs[0]
            if cur and (nodes + lp.m + lp.n > self.max_nodes or nnz + z > self.max_nnz or len(cur) >= self.max_lps):
                packs.append(cur)
                cur, nodes, nnz = [], 0, 0
            cur.append(i)
            nodes += lp.m + lp.n
            nnz += z
        if cur:
            packs.append(cur)
        return packs

    def _stage(self, slot, lps, ids, first_use):
        """H2D of one pack on the copy stream: ONE copy per LP (its pinned pack, verbatim, into a device staging area,
        LP after LP) + one copy of the pack's offset tables.  Staging layout (int32 words): LP b at ``stage_off[b]``:
        [row z_b | col z_b | val z_b | x_s m_b*p | x_t n_b*q]; tables: [stage_off | edge_ptr | cons_ptr | vars_ptr], B+1
        words each.  ``lpgnn_pack_scatter`` (compute stream) turns that into the pack layout."""
        sub = [lps[i] for i in ids]
        B = len(sub)
        words = np.fromiter((lp.offs[5] for lp in sub), dtype=np.int64, count=B)
        zs = np.fromiter((lp.offs[1] for lp in sub), dtype=np.int64, count=B)
        ms = np.fromiter((lp.m for lp in sub), dtype=np.int64, count=B)
        ns = np.fromiter((lp.n for lp in sub), dtype=np.int64, count=B)
        tables = np.zeros((4, B + 1), dtype=np.int64)
        np.cumsum(words, out=tables[0, 1:]); np.cumsum(zs, out=tables[1, 1:])
        np.cumsum(ms, out=tables[2, 1:]); np.cumsum(ns, out=tables[3, 1:])
        W, Z, M, N = (int(tables[k, -1]) for k in range(4))
        p, q = sub[0].p, sub[0].q
        need = W + 4 * (B + 1)
        if self.d_stage[slot] is None or self.d_stage[slot].numel() < need:
            if not first_use:
                self.done[slot].synchronize()   # the slot's previous pack may still be computing on this buffer
            self.d_stage[slot] = torch.empty(int(need * 1.25) + 64, dtype=torch.int32, device=self.dev)
        if self.h_ptr[slot] is None:
            self.h_ptr[slot] = torch.empty(4 * (self.max_lps + 1), dtype=torch.int32).pin_memory()
        hp, d = self.h_ptr[slot], self.d_stage[slot]
        if not first_use:
            self.ready[slot].synchronize()      # the slot's previous H2D has consumed the pinned table buffer
        hp.numpy()[:4 * (B + 1)] = tables.reshape(-1)
        base = d.data_ptr()
        src = np.empty(B + 1, dtype=np.uint64); dst = np.empty(B + 1, dtype=np.uint64); nby = np.empty(B + 1, dtype=np.uint64)
        src[:B] = np.fromiter((lp.pack.data_ptr() for lp in sub), dtype=np.uint64, count=B)
        dst[:B] = np.uint64(base) + (4 * tables[0, :B]).astype(np.uint64)
        nby[:B] = (4 * words).astype(np.uint64)
        src[B], dst[B], nby[B] = hp.data_ptr(), base + 4 * W, 16 * (B + 1)
        from . import _lib
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self.done[slot])
            with torch.cuda.device(self.dev):
                rc = _lib.load().lpgnn_copy_many_h2d(dst.ctypes.data, src.ctypes.data, nby.ctypes.data, B + 1,
                                                     self.copy_stream.cuda_stream)
            _lib.check(rc, "lpgnn_copy_many_h2d")
            self.ready[slot].record(self.copy_stream)
        return dict(B=B, W=W, Z=Z, M=M, N=N, p=p, q=q, c_ptr=tables[2], v_ptr=tables[3], sorted=all(lp.sorted for lp in sub))

    @torch.no_grad()
    def _compute(self, slot, meta):
        if self.compute is None:
            return self._compute_on(slot, meta, torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.compute[slot]):
            return self._compute_on(slot, meta, self.compute[slot])

    def _compute_on(self, slot, meta, cur):
        from . import _lib
        M, N, Z, B, W, p, q = meta["M"], meta["N"], meta["Z"], meta["B"], meta["W"], meta["p"], meta["q"]
        if self.h_status[slot] is None or self.h_status[slot].numel() < M + N:
            self.h_status[slot] = torch.empty(int((M + N) * 1.25) + 64, dtype=torch.uint8).pin_memory()
        words = 3 * Z + M * p + N * q
        if self.d_buf[slot] is None or self.d_buf[slot].numel() < words:      # same stream as its previous users: safe to replace
            self.d_buf[slot] = torch.empty(int(words * 1.25) + 64, dtype=torch.int32, device=self.dev)
        cur.wait_event(self.ready[slot])
        st, d = self.d_stage[slot], self.d_buf[slot]
        row, col = d[0:Z], d[Z:2 * Z]
        val = d[2 * Z:3 * Z].view(torch.float32)
        x_s = d[3 * Z:3 * Z + M * p].view(torch.float32).view(M, p)
        x_t = d[3 * Z + M * p:words].view(torch.float32).view(N, q)
        s_off, e_ptr, c_ptr, v_ptr = (st[W + k * (B + 1):W + (k + 1) * (B + 1)] for k in range(4))
        with torch.cuda.device(self.dev):
            rc = _lib.load().lpgnn_pack_scatter(st.data_ptr(), s_off.data_ptr(), e_ptr.data_ptr(), c_ptr.data_ptr(), v_ptr.data_ptr(),
                                                B, p, q, W, row.data_ptr(), col.data_ptr(), val.data_ptr(), x_s.data_ptr(),
                                                x_t.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "lpgnn_pack_scatter")
        stt = self.model.predict_basis_packed(row, col, val, M, N, x_s, x_t, c_ptr, v_ptr, is_sorted=meta["sorted"],
                                              buffers=self.bufs[slot], lp_major=True)
        self.h_status[slot][:M + N].copy_(stt, non_blocking=True)
        self.done[slot].record(cur)

    def _emit(self, slot, meta, ids):
        """Statuses arrive LP by LP (constraints, then variables of each LP): every yield is a copy of one slice."""
        self.done[slot].synchronize()
        h = self.h_status[slot].numpy()
        start = meta["c_ptr"] + meta["v_ptr"]
        for b, i in enumerate(ids):
            yield i, h[start[b]:start[b + 1]].copy()

    def run(self, host_lps):
        lps = list(host_lps)
        packs = self._plan(lps)
        if not packs:
            return
        self.model._native_weights()            # on the caller's stream, before the compute streams fork (see BasisPipeline.run)
        if self.compute is not None:
            for st in self.compute:
                st.wait_stream(torch.cuda.current_stream(self.dev))
        metas = [None] * len(packs)
        metas[0] = self._stage(0, lps, packs[0], True)
        for k, ids in enumerate(packs):
            slot = k & 1
            if k + 1 < len(packs):
                metas[k + 1] = self._stage(slot ^ 1, lps, packs[k + 1], k + 1 < 2)
            self._compute(slot, metas[k])
            if k > 0:
                yield from self._emit(slot ^ 1, metas[k - 1], packs[k - 1])
        last = len(packs) - 1
        yield from self._emit(last & 1, metas[last], packs[last])

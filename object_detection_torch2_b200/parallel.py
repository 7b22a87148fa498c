"""Image-sharded data parallelism for the head path: one process per GPU, no data-path collective.

Matching, mining, normalisation, NMS and TP/FP assignment are all per image (reference ssd.py:222-227,
utils.py:113, evaluate.py:134), so rank r simply owns a contiguous block of the batch.  The only exchanges are
  * training: one all-reduce(sum) of a packed fp64 vector [sum_i loss_i / N_global, n_images, sum pos, ...],
    kept OFF the critical path (the gradient needs only the constant 1 / N_global, never a reduced value);
  * evaluation: one all-reduce(sum) of the int64 (20, 3) tallies at the end (exact for any world size).
The reference has no multi-GPU code; single-GPU results are the comparator.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``; the first ``n_items % world_size`` ranks get one extra item."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(outputs: torch.Tensor, targets: torch.Tensor, world_size: int, rank: int):
    lo, hi = shard_bounds(outputs.shape[0], world_size, rank)
    return outputs[lo:hi], targets[lo:hi]


class ScalarAllReducer:
    """Packs per-step scalars and all-reduces them every ``window`` steps on a side stream (CUDA) so that the
    latency of a tiny NCCL message never serialises with the step kernels.  ``flush()`` returns the reduced rows."""

    def __init__(self, width: int, window: int = 16, device: Optional[torch.device] = None, group=None,
                 dtype: torch.dtype = torch.float64):
        self.width, self.window, self.group = width, max(1, window), group
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.buf = torch.zeros(self.window, width, dtype=dtype, device=self.device)
        self.fill = 0
        self.pending: List[Tuple[torch.Tensor, object]] = []
        self.side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def push(self, row: torch.Tensor) -> None:
        """``row``: (width,) tensor on ``device``; copied (stream-ordered) into the current window."""
        self.buf[self.fill].copy_(row.to(self.buf.dtype), non_blocking=True)
        self.fill += 1
        if self.fill == self.window:
            self._launch()

    def _launch(self) -> None:
        if self.fill == 0:
            return
        chunk = self.buf[: self.fill].clone()
        self.fill = 0
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            self.pending.append((chunk, None))
            return
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.side):
                chunk.record_stream(self.side)
                work = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((chunk, work))

    def flush(self) -> torch.Tensor:
        """Launch the partial window, wait for everything in flight, return all reduced rows (steps, width)."""
        self._launch()
        rows = []
        for chunk, work in self.pending:
            if work is not None:
                work.wait()
            rows.append(chunk)
        self.pending = []
        if self.side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        return torch.cat(rows) if rows else self.buf[:0].clone()


class ScalarExchange:
    """The step's scalar all-reduce without a collective launch (csrc/exchange.cu): every rank owns an inbox in device memory
    that its peers map through CUDA IPC; ``ops.multibox_loss_raw(..., exchange=self)`` makes the loss kernel's last CTA store
    the step's scalar into every rank's inbox over NVLink, and ``reduce(count)`` (one tiny kernel, CUDA-graph capturable)
    returns the rank-ordered sums of the next ``count`` steps -- bit-identical on every rank.

    torch.distributed is only the transport of the 64-byte IPC handles at construction.  Raises if CUDA IPC is not
    available (callers fall back to ``ScalarAllReducer`` / NCCL)."""

    def __init__(self, device, group=None):
        import ctypes

        from . import _lib
        self._lib = lib = _lib.load()
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > _lib.MAX_RANKS:
            raise ValueError(f"ScalarExchange supports up to {_lib.MAX_RANKS} ranks")
        self._peers, self._inbox = [], None
        with torch.cuda.device(self.device):
            # Every step below is executed by EVERY rank whatever happened locally (a rank that cannot create or open an inbox
            # still takes part in the collectives and reports through the flag), so a failure never strands the others.
            inbox, handle = ctypes.c_void_p(), _lib.IpcHandle()
            failed = 0
            if lib.ssdh_scalar_exchange_create(self.world, ctypes.byref(inbox), ctypes.byref(handle)) != 0:
                failed = 1
            self._inbox = inbox.value
            mine = torch.tensor(list(bytes(handle.bytes)), dtype=torch.uint8, device=self.device)
            handles = [torch.zeros_like(mine) for _ in range(self.world)]
            flag = torch.tensor([failed], device=self.device)
            if self.world > 1:
                dist.all_gather(handles, mine, group=group)
                dist.all_reduce(flag, group=group)
            else:
                handles[0] = mine
            desc = _lib.ScalarExchange()
            desc.world, desc.rank, desc.ring = self.world, self.rank, _lib.XCHG_RING
            failed = int(flag)
            if not failed:
                for r in range(self.world):
                    if r == self.rank:
                        desc.inbox[r] = self._inbox
                        continue
                    h = _lib.IpcHandle()
                    ctypes.memmove(h.bytes, bytes(handles[r].cpu().tolist()), 64)
                    ptr = ctypes.c_void_p()
                    if lib.ssdh_scalar_exchange_open(ctypes.byref(h), ctypes.byref(ptr)) != 0:
                        failed = 1
                        break
                    self._peers.append(ptr.value)
                    desc.inbox[r] = ptr.value
            flag = torch.tensor([failed], device=self.device)
            if self.world > 1:
                dist.all_reduce(flag, group=group)                     # all ranks agree on whether the exchange is usable
            if int(flag) != 0:
                msg = lib.ssdh_last_error()
                if self.world > 1:
                    dist.barrier(group=group)                          # nobody unmaps while a peer is still opening
                self.close()
                raise RuntimeError(f"CUDA IPC is not available between the GPUs of this job ({msg.decode() if msg else 'a peer failed'})")
            desc.counters = self._inbox + self.world * _lib.XCHG_RING * 8
            self.desc = desc
            self.status = torch.zeros(1, dtype=torch.int32, device=self.device)

    def reduce(self, count: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Sums over ranks of the next ``count`` steps' scalars (count <= ring / 2), on the current stream."""
        import ctypes

        from . import _lib
        if out is None:
            out = torch.empty(count, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ssdh_scalar_exchange_reduce(ctypes.byref(self.desc), int(count), out.data_ptr(), self.status.data_ptr(),
                                                             torch.cuda.current_stream().cuda_stream), "ssdh_scalar_exchange_reduce")
        return out

    def ok(self) -> bool:
        """False if a ``reduce`` gave up on a peer (synchronises)."""
        return int(self.status.item()) == 0

    def close(self) -> None:
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self._peers:
                self._lib.ssdh_scalar_exchange_close(p)
            self._peers = []
            if self._inbox:
                self._lib.ssdh_scalar_exchange_destroy(self._inbox)
                self._inbox = None


def all_reduce_tallies(tallies: torch.Tensor, group=None) -> torch.Tensor:
    """Exact integer sum of the (C-1, 3) TP / detection / ground-truth tallies over all ranks (in place)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tallies, op=dist.ReduceOp.SUM, group=group)
    return tallies


def global_loss(local_loss_sum_over_n_global: torch.Tensor, group=None) -> torch.Tensor:
    """Synchronous variant: each rank holds sum_i(local) loss_i / N_global (what ssdh_multibox_loss writes when
    called with n_global = global batch); their sum is the reference's batch mean (ssd.py:227)."""
    out = local_loss_sum_over_n_global.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out

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

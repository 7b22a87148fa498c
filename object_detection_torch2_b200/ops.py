"""Tensor-level wrappers over the C ABI: argument checks, caller-owned workspaces, current-stream launches.

Every function requires CUDA fp32 tensors and raises otherwise -- the product has no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, NamedTuple, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import ImageStats, check

NUM_PRIORS = 8732
_ws_cache: Dict[Tuple, torch.Tensor] = {}
_ws_retired = []      # outgrown buffers stay alive: a captured CUDA graph may still point at them


def _need_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError("ssdhead kernels run on CUDA tensors only (no CPU fallback); got "
                               f"{'a non-tensor' if not isinstance(t, torch.Tensor) else t.device}")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _workspace(kind: str, nbytes: int, device: torch.device, zero: bool = False) -> torch.Tensor:
    """Scratch buffers are cached per (kind, device, stream) and grown on demand; the library itself never allocates."""
    key = (kind, device.index, _stream())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ws_retired.append(buf)
        buf = (torch.zeros if zero else torch.empty)(max(nbytes, 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def device_info() -> dict:
    lib = _lib.load()
    vals = [ctypes.c_int(0) for _ in range(4)]
    check(lib.ssdh_device_info(*[ctypes.byref(v) for v in vals]), "ssdh_device_info")
    return dict(sm_count=vals[0].value, max_smem_optin=vals[1].value, loss_cluster_size=vals[2].value,
                loss_max_active_clusters=vals[3].value)


# ------------------------------------------------------------------------------------------------ P1
def default_boxes(device="cuda") -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(NUM_PRIORS, 4, dtype=torch.float32, device=device)
    _need_cuda(out)
    with torch.cuda.device(out.device):
        check(lib.ssdh_default_boxes(out.data_ptr(), _stream()), "ssdh_default_boxes")
    return out


# ------------------------------------------------------------------------------------------------ L1-L6 probes
class MatchResult(NamedTuple):
    bits: Optional[torch.Tensor]
    mask: Optional[torch.Tensor]
    best_gt: Optional[torch.Tensor]
    best_iou: Optional[torch.Tensor]
    best_prior: Optional[torch.Tensor]
    best_prior_iou: Optional[torch.Tensor]


def match(gt: torch.Tensor, priors: torch.Tensor, threshold: float = 0.25, want_bits: bool = False, want_mask: bool = True,
          want_best_gt: bool = False, want_best_prior: bool = False, force_best_prior: bool = False) -> MatchResult:
    """``force_best_prior`` (north_star extension, off = the reference): every real ground-truth box also claims the prior
    ssdh_match reports as its arg-max IoU (lowest index on ties) when that IoU is positive."""
    lib = _lib.load()
    want_best_prior = want_best_prior or force_best_prior
    _need_cuda(gt, priors)
    gt, priors = _f32c(gt), _f32c(priors)
    N, G, stride = gt.shape
    P = priors.shape[0]
    dev = gt.device
    bits = torch.empty(N, P, dtype=torch.int64, device=dev) if want_bits else None
    mask = torch.empty(N, P, G, dtype=torch.bool, device=dev) if want_mask else None
    bg = torch.empty(N, P, dtype=torch.int32, device=dev) if want_best_gt else None
    bi = torch.empty(N, P, dtype=torch.float32, device=dev) if want_best_gt else None
    bp = torch.empty(N, G, dtype=torch.int32, device=dev) if want_best_prior else None
    bpi = torch.empty(N, G, dtype=torch.float32, device=dev) if want_best_prior else None
    if N > 0 and P > 0:
        with torch.cuda.device(dev):
            check(lib.ssdh_match(gt.data_ptr(), stride, N, G, priors.data_ptr(), P, float(threshold), _ptr(bits), _ptr(mask),
                                 _ptr(bg), _ptr(bi), _ptr(bp), _ptr(bpi), _stream()), "ssdh_match")
    if force_best_prior and N > 0 and G > 0 and P > 0:
        real = ((gt[:, :, 2] * gt[:, :, 3]) > 0) & (bpi > 0)
        nn, gg = torch.nonzero(real, as_tuple=True)
        rows = bp[nn, gg].long()
        if mask is not None:
            mask[nn, rows, gg] = True
        if bits is not None:
            add = torch.zeros_like(bits)
            add.index_put_((nn, rows), torch.ones_like(gg, dtype=torch.int64) << gg, accumulate=True)     # distinct gt rows: sum == OR
            bits |= add
    return MatchResult(bits, mask, bg, bi, bp, bpi)


def encode(gt: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(gt, priors)
    gt, priors = _f32c(gt), _f32c(priors)
    N, G, stride = gt.shape
    P = priors.shape[0]
    out = torch.empty(N, P, G, 4, dtype=torch.float32, device=gt.device)
    if out.numel():
        with torch.cuda.device(gt.device):
            check(lib.ssdh_encode(gt.data_ptr(), stride, N, G, priors.data_ptr(), P, out.data_ptr(), _stream()), "ssdh_encode")
    return out


def smooth_l1(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(x)
    x = _f32c(x)
    out = torch.empty_like(x)
    if x.numel():
        with torch.cuda.device(x.device):
            check(lib.ssdh_smooth_l1(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "ssdh_smooth_l1")
    return out


def softmax_cross_entropy(pr: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """pr (N, P, C) logits, gt (N, G, C) class weights (G may be 1 and N may broadcast) -> (N, P, G)."""
    lib = _lib.load()
    _need_cuda(pr, gt)
    if pr.stride(-1) != 1 or pr.stride(0) != pr.shape[1] * pr.stride(1) or pr.dtype != torch.float32:
        pr = _f32c(pr)
    if gt.shape[0] != pr.shape[0]:
        gt = gt.expand(pr.shape[0], -1, -1)
    if gt.stride(-1) != 1 or gt.stride(0) != gt.shape[1] * gt.stride(1) or gt.dtype != torch.float32:
        gt = _f32c(gt)
    N, P, C = pr.shape
    G = gt.shape[1]
    out = torch.empty(N, P, G, dtype=torch.float32, device=pr.device)
    if out.numel():
        with torch.cuda.device(pr.device):
            check(lib.ssdh_softmax_cross_entropy(pr.data_ptr(), pr.stride(1), gt.data_ptr(), gt.stride(1), N, P, G, C,
                                                 out.data_ptr(), _stream()), "ssdh_softmax_cross_entropy")
    return out


def split_pos_neg(pos: torch.Tensor, neg: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    _need_cuda(pos, neg)
    pos, neg = pos.long().contiguous(), neg.long().contiguous()
    po, no = torch.empty_like(pos), torch.empty_like(neg)
    if pos.numel():
        with torch.cuda.device(pos.device):
            check(lib.ssdh_split_pos_neg(pos.data_ptr(), neg.data_ptr(), po.data_ptr(), no.data_ptr(), pos.numel(), _stream()),
                  "ssdh_split_pos_neg")
    return po, no


def kplus1_value(values: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """values (rows, len) or (len,), k (rows,) or 0-d int64 -> (k+1)-th largest per row."""
    lib = _lib.load()
    _need_cuda(values, k)
    squeeze = values.dim() == 1
    v = _f32c(values.unsqueeze(0) if squeeze else values)
    kk = k.reshape(-1).long().contiguous()
    out = torch.empty(v.shape[0], dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        check(lib.ssdh_kplus1_value(v.data_ptr(), v.shape[0], v.shape[1], kk.data_ptr(), out.data_ptr(), _stream()), "ssdh_kplus1_value")
    return out[0] if squeeze else out


# ------------------------------------------------------------------------------------------------ L1-L7 fused
STATS_DTYPE = np.dtype([("loss", "<f4"), ("thr_pos", "<f4"), ("thr_neg", "<f4"), ("pos_raw", "<i4"), ("k_pos", "<i4"),
                        ("k_neg", "<i4"), ("pos_sel", "<i4"), ("neg_sel", "<i4")])
assert STATS_DTYPE.itemsize == ctypes.sizeof(ImageStats)


def multibox_loss_raw(outputs: torch.Tensor, targets: torch.Tensor, priors: torch.Tensor, a: float = 1.0, threshold: float = 0.25,
                      n_global: Optional[int] = None, want_grad: bool = True, want_stats: bool = False,
                      loss_out: Optional[torch.Tensor] = None, grad_out: Optional[torch.Tensor] = None,
                      stats_out: Optional[torch.Tensor] = None, next_outputs: Optional[torch.Tensor] = None,
                      next_targets: Optional[torch.Tensor] = None, inputs_stable: bool = False, force_best_prior: bool = False,
                      exact_math: bool = False, ce_override: Optional[torch.Tensor] = None, exchange=None):
    """One launch: loss (0-d), d loss / d outputs (or None) and per-image stats (uint8 (N, 32) view of ssdh_image_stats, or None).

    Inputs must already be contiguous fp32 CUDA tensors (this is the graph-capturable hot call).  ``inputs_stable`` (implied
    by ``next_outputs`` / ``next_targets``): the caller vouches that the inputs were not written by the kernel that precedes
    this call in the stream, see ssdh_multibox_loss_pipelined.  ``force_best_prior`` (north_star extension, off = the
    reference), ``exact_math``, ``ce_override`` and ``exchange`` (a ``parallel.ScalarExchange``: the step's loss scalar is stored into every
    rank's inbox over NVLink by the kernel itself) go through ssdh_multibox_loss_ex (see include/ssdhead.h)."""
    lib = _lib.load()
    _need_cuda(outputs, targets, priors)
    _check_loss_args(outputs, targets, priors, next_outputs, next_targets, loss_out, grad_out)
    N, P, row = outputs.shape
    C = row - 4
    G = targets.shape[1]
    dev = outputs.device
    if loss_out is None:
        loss_out = torch.empty((), dtype=torch.float32, device=dev)
    if want_grad and grad_out is None:
        grad_out = torch.empty_like(outputs)
    if want_stats and stats_out is None:
        stats_out = torch.empty(N, ctypes.sizeof(ImageStats), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib.ssdh_multibox_loss_workspace_bytes(N, P, C, G)
        ws = _workspace("loss", nbytes, dev, zero=True)
        args = (outputs.data_ptr(), targets.data_ptr() if G > 0 else None, priors.data_ptr(), N, P, C, G,
                float(a), float(threshold), int(n_global or N), loss_out.data_ptr(),
                _ptr(grad_out) if want_grad else None, _ptr(stats_out) if want_stats else None,
                ws.data_ptr(), ws.numel(), _stream())
        if force_best_prior or exact_math or ce_override is not None or exchange is not None:
            _need_cuda(next_outputs, next_targets, ce_override)
            if ce_override is not None and (ce_override.dtype != torch.float32 or tuple(ce_override.shape) != (N, P) or not ce_override.is_contiguous()):
                raise ValueError("multibox_loss: ce_override must be a contiguous fp32 (N, P) tensor")
            opt = _lib.LossOptions(ctypes.sizeof(_lib.LossOptions), int(bool(force_best_prior)),
                                   int(bool(inputs_stable or next_outputs is not None or next_targets is not None)), int(bool(exact_math)),
                                   _ptr(next_outputs), _ptr(next_targets), _ptr(ce_override),
                                   ctypes.pointer(exchange.desc) if exchange is not None else None)
            check(lib.ssdh_multibox_loss_ex(*args, ctypes.byref(opt)), "ssdh_multibox_loss_ex")
        elif next_outputs is None and next_targets is None and not inputs_stable:
            check(lib.ssdh_multibox_loss(*args), "ssdh_multibox_loss")
        else:                                   # L2 prefetch of the next micro-batch (same shapes) from inside the kernel
            _need_cuda(next_outputs, next_targets)
            check(lib.ssdh_multibox_loss_pipelined(*args, _ptr(next_outputs), _ptr(next_targets)), "ssdh_multibox_loss_pipelined")
    return loss_out, (grad_out if want_grad else None), (stats_out if want_stats else None)


def _check_loss_args(outputs, targets, priors, next_outputs, next_targets, loss_out, grad_out) -> None:
    """The kernel trusts these shapes: a mismatched class count or stride would make it read out of bounds."""
    if outputs.dim() != 3 or targets.dim() != 3 or priors.dim() != 2:
        raise ValueError("multibox_loss: outputs (N, P, 4+C), targets (N, G, 4+C), priors (P, 4) expected")
    N, P, row = outputs.shape
    if row < 5:
        raise ValueError("multibox_loss: outputs rows are [4 offsets, C >= 1 logits]")
    if targets.shape[0] != N or targets.shape[2] != row:
        raise ValueError(f"multibox_loss: targets {tuple(targets.shape)} do not match outputs {tuple(outputs.shape)} (need (N, G, {row}))")
    if tuple(priors.shape) != (P, 4):
        raise ValueError(f"multibox_loss: priors {tuple(priors.shape)} do not match outputs (need ({P}, 4))")
    for name, t, shape in (("outputs", outputs, None), ("targets", targets, None), ("priors", priors, None),
                           ("next_outputs", next_outputs, outputs.shape), ("next_targets", next_targets, targets.shape),
                           ("grad_out", grad_out, outputs.shape)):
        if t is None:
            continue
        if t.dtype != torch.float32 or not t.is_contiguous() or t.device != outputs.device:
            raise ValueError(f"multibox_loss: {name} must be a contiguous fp32 tensor on {outputs.device}")
        if shape is not None and t.shape != shape:
            raise ValueError(f"multibox_loss: {name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    if loss_out is not None and (loss_out.dtype != torch.float32 or loss_out.numel() != 1 or loss_out.device != outputs.device):
        raise ValueError("multibox_loss: loss_out must be a one-element fp32 tensor on the device of outputs")


def stats_to_numpy(stats: torch.Tensor) -> np.ndarray:
    return stats.cpu().numpy().view(STATS_DTYPE).reshape(-1)


class _MultiBoxLossFn(torch.autograd.Function):
    """SSD.loss with its analytic gradient produced by the same launch as the forward value."""

    @staticmethod
    def forward(ctx, outputs, targets, priors, a, threshold, n_global, force_best_prior):
        want_grad = ctx.needs_input_grad[0]
        loss, grad, _ = multibox_loss_raw(outputs, targets, priors, a, threshold, n_global, want_grad=want_grad,
                                          force_best_prior=force_best_prior)
        ctx.grad = grad
        ctx.consumed = False
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        grad = ctx.grad
        if grad is None:
            return (None,) * 7
        if ctx.consumed:
            # the chain-rule factor is applied in place on the gradient the forward launch wrote (no second 28 MB buffer on
            # the hot path), so that buffer cannot serve a second backward
            raise RuntimeError("SSD.loss: backward through the fused MultiBox loss a second time (retain_graph=True or several "
                               "heads over one loss); call loss() again instead -- the gradient buffer was scaled in place")
        ctx.consumed = True
        lib = _lib.load()
        g = _f32c(g.reshape(1))
        with torch.cuda.device(grad.device):
            check(lib.ssdh_scale_inplace(grad.data_ptr(), grad.numel(), g.data_ptr(), _stream()), "ssdh_scale_inplace")
        return grad, None, None, None, None, None, None


def multibox_loss(outputs: torch.Tensor, targets: torch.Tensor, priors: torch.Tensor, a: float = 1.0, threshold: float = 0.25,
                  n_global: Optional[int] = None, force_best_prior: bool = False) -> torch.Tensor:
    """Differentiable (w.r.t. ``outputs``) MultiBox loss, 0-dim tensor."""
    _need_cuda(outputs, targets, priors)
    o = outputs if (outputs.dtype == torch.float32 and outputs.is_contiguous()) else outputs.float().contiguous()
    return _MultiBoxLossFn.apply(o, _f32c(targets.detach()), _f32c(priors.detach()), float(a), float(threshold), n_global,
                                 bool(force_best_prior))


def prefetch_l2(t: torch.Tensor) -> None:
    """Pull a tensor from HBM into L2 on the current stream (software pipelining of the next micro-batch)."""
    lib = _lib.load()
    _need_cuda(t)
    if t.numel():
        with torch.cuda.device(t.device):
            check(lib.ssdh_prefetch_l2(t.data_ptr(), t.numel() * t.element_size(), _stream()), "ssdh_prefetch_l2")


# ------------------------------------------------------------------------------------------------ 8f-1 head producer
def _pack_call(levels, slab, width, unpack, nhwc=False):
    lib = _lib.load()
    n_levels = len(levels)
    N = slab.shape[0]
    ptrs = (ctypes.c_void_p * n_levels)(*[t.data_ptr() for t in levels])
    ch = (ctypes.c_int * n_levels)(*[t.shape[1] for t in levels])
    hw = (ctypes.c_int * n_levels)(*[t.shape[2] * t.shape[3] for t in levels])
    sfx = "_nhwc" if nhwc else ""
    with torch.cuda.device(slab.device):
        if unpack:
            check(getattr(lib, "ssdh_unpack_head" + sfx)(slab.data_ptr(), ptrs, ch, hw, n_levels, N, width, slab.shape[1], _stream()), "ssdh_unpack_head" + sfx)
        else:
            check(getattr(lib, "ssdh_pack_head" + sfx)(ptrs, ch, hw, n_levels, N, width, slab.data_ptr(), slab.shape[1], _stream()), "ssdh_pack_head" + sfx)


def _channels_last(t: torch.Tensor) -> bool:
    """The tensor's memory is (N, H*W, C): torch.channels_last (a 1 x 1 map is both layouts at once)."""
    return t.dtype == torch.float32 and t.is_contiguous(memory_format=torch.channels_last)


class _PackHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, width, *levels):
        _need_cuda(*levels)
        # channels-last producers (what cuDNN's tensor-core convolutions write): every (level, image) block is already in slab
        # order, the pass is a plain copy; otherwise NCHW, the transposing kernel
        nhwc = all(_channels_last(t) for t in levels)
        if not nhwc:
            levels = [_f32c(t) for t in levels]
        N = levels[0].shape[0]
        rows = sum(t.shape[1] // width * t.shape[2] * t.shape[3] for t in levels)
        out = torch.empty((N, rows, width), dtype=torch.float32, device=levels[0].device)
        _pack_call(levels, out, width, unpack=False, nhwc=nhwc)
        ctx.width = width
        ctx.shapes = [t.shape for t in levels]
        ctx.nhwc = nhwc
        return out

    @staticmethod
    def backward(ctx, grad):
        grad = _f32c(grad)
        fmt = torch.channels_last if ctx.nhwc else torch.contiguous_format
        grads = [torch.empty(s, dtype=torch.float32, device=grad.device, memory_format=fmt) for s in ctx.shapes]
        _pack_call(grads, grad, ctx.width, unpack=True, nhwc=ctx.nhwc)
        return (None, *grads)


def pack_head(levels, width: int) -> torch.Tensor:
    """[(N, A_k * width, H_k, W_k)] detector outputs -> (N, sum_k H_k W_k A_k, width): the permute / reshape / cat tail of
    SSD.forward (reference ssd.py:96-104) as ONE pass; differentiable (the backward is the same kernel run in reverse)."""
    for t in levels:
        if t.dim() != 4 or t.shape[1] % width != 0 or t.shape[0] != levels[0].shape[0]:
            raise ValueError("pack_head: every level must be (N, anchors * width, H, W)")
    return _PackHeadFn.apply(width, *levels)


# ------------------------------------------------------------------------------------------------ 8f-3 ground-truth ingest
def expand_targets(compact: torch.Tensor, lengths: Optional[torch.Tensor], num_classes: int) -> torch.Tensor:
    """(N, G, 5) rows [cx, cy, w, h, label] (+ per-image row counts) -> the dense zero-padded one-hot (N, G, 4 + C) tensor
    the reference's collate_fn builds on the host (src/utils.py:8-16)."""
    lib = _lib.load()
    _need_cuda(compact)
    compact = _f32c(compact)
    N, G, five = compact.shape
    if five != 5:
        raise ValueError("expand_targets: compact rows are [cx, cy, w, h, label]")
    if lengths is not None:
        _need_cuda(lengths)
        lengths = lengths.to(torch.int32).contiguous()
        if lengths.numel() != N:
            raise ValueError("expand_targets: one length per image")
    out = torch.empty((N, G, 4 + num_classes), dtype=torch.float32, device=compact.device)
    with torch.cuda.device(compact.device):
        check(lib.ssdh_expand_targets(compact.data_ptr(), _ptr(lengths), N, G, num_classes, out.data_ptr(), _stream()), "ssdh_expand_targets")
    return out


# ------------------------------------------------------------------------------------------------ I1-I4
def decode(pr: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(pr, priors)
    pr, priors = _f32c(pr), _f32c(priors)
    N, P, stride = pr.shape
    out = torch.empty(N, P, 4, dtype=torch.float32, device=pr.device)
    if out.numel():
        with torch.cuda.device(pr.device):
            check(lib.ssdh_decode(pr.data_ptr(), stride, priors.data_ptr(), N, P, out.data_ptr(), _stream()), "ssdh_decode")
    return out


def score(pr: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(pr)
    pr = _f32c(pr)
    N, P, stride = pr.shape
    C = stride - 4
    out = torch.empty(N, P, C, dtype=torch.float32, device=pr.device)
    if out.numel():
        with torch.cuda.device(pr.device):
            check(lib.ssdh_score(pr.data_ptr(), stride, N, P, C, out.data_ptr(), _stream()), "ssdh_score")
    return out


def iou(t: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(t, s)
    t, s = _f32c(t), _f32c(s)
    N, T, ts = t.shape
    S, ss = s.shape[1], s.shape[2]
    out = torch.empty(N, T, S, dtype=torch.float32, device=t.device)
    if out.numel():
        with torch.cuda.device(t.device):
            check(lib.ssdh_iou(t.data_ptr(), ts, T, s.data_ptr(), ss, S, N, out.data_ptr(), _stream()), "ssdh_iou")
    return out


class NmsResult(NamedTuple):
    outputs: torch.Tensor
    order: Optional[torch.Tensor]
    order_cnt: Optional[torch.Tensor]
    keep: Optional[torch.Tensor]
    keep_cnt: Optional[torch.Tensor]


def _nms_call(fn_name: str, outputs: torch.Tensor, priors: Optional[torch.Tensor], iou_thresh: float, score_thresh: float,
              top_k: Optional[int], per_class: bool, want_lists: bool) -> NmsResult:
    lib = _lib.load()
    _need_cuda(outputs, priors)
    if outputs.dtype != torch.float32 or not outputs.is_contiguous():
        raise RuntimeError(f"{fn_name} works in place and needs a contiguous fp32 tensor")
    N, P, row = outputs.shape
    C = row - 4
    dev = outputs.device
    order = order_cnt = keep = keep_cnt = None
    if want_lists:
        order = torch.empty(N, P, dtype=torch.int32, device=dev)
        keep = torch.empty(N, P, dtype=torch.int32, device=dev)
        order_cnt = torch.empty(N, dtype=torch.int32, device=dev)
        keep_cnt = torch.empty(N, dtype=torch.int32, device=dev)
    if N == 0 or P == 0:
        return NmsResult(outputs, order, order_cnt, keep, keep_cnt)
    with torch.cuda.device(dev):
        ws = _workspace("nms", lib.ssdh_nms_workspace_bytes(N, P, C), dev)
        tail = (float(iou_thresh), float(score_thresh), int(top_k or 0), int(bool(per_class)), _ptr(order), _ptr(order_cnt),
                _ptr(keep), _ptr(keep_cnt), ws.data_ptr(), ws.numel(), _stream())
        if priors is None:
            check(lib.ssdh_nms(outputs.data_ptr(), N, P, C, *tail), "ssdh_nms")
        else:
            priors = _f32c(priors)
            check(lib.ssdh_postprocess(outputs.data_ptr(), priors.data_ptr(), N, P, C, *tail), "ssdh_postprocess")
    return NmsResult(outputs, order, order_cnt, keep, keep_cnt)


def nms_(outputs: torch.Tensor, iou_thresh: float = 0.5, score_thresh: float = 0.0, top_k: Optional[int] = None,
         per_class: bool = False, want_lists: bool = False) -> NmsResult:
    return _nms_call("nms_", outputs, None, iou_thresh, score_thresh, top_k, per_class, want_lists)


def postprocess_(outputs: torch.Tensor, priors: torch.Tensor, iou_thresh: float = 0.5, score_thresh: float = 0.0,
                 top_k: Optional[int] = None, per_class: bool = False, want_lists: bool = False) -> NmsResult:
    return _nms_call("postprocess_", outputs, priors, iou_thresh, score_thresh, top_k, per_class, want_lists)


def gather_detections(outputs: torch.Tensor, keep: torch.Tensor, keep_cnt: torch.Tensor, max_det: int = 200):
    """Compact per-image detection lists: (dets (N, max_det, 6) = [cx, cy, w, h, score, label], det_cnt (N,) int32)."""
    lib = _lib.load()
    _need_cuda(outputs, keep, keep_cnt)
    outputs = _f32c(outputs)
    N, P, row = outputs.shape
    dets = torch.empty(N, max_det, 6, dtype=torch.float32, device=outputs.device)
    cnt = torch.empty(N, dtype=torch.int32, device=outputs.device)
    if N > 0:
        with torch.cuda.device(outputs.device):
            check(lib.ssdh_gather_detections(outputs.data_ptr(), keep.contiguous().data_ptr(), keep_cnt.contiguous().data_ptr(), N, P, row - 4,
                                             int(max_det), dets.data_ptr(), cnt.data_ptr(), _stream()), "ssdh_gather_detections")
    return dets, cnt


# ------------------------------------------------------------------------------------------------ E1-E2
def eval_accumulate(outputs: torch.Tensor, gts: torch.Tensor, tallies: Optional[torch.Tensor] = None, iou_thresh: float = 0.5,
                    want_flags: bool = False, keep: Optional[torch.Tensor] = None, keep_cnt: Optional[torch.Tensor] = None,
                    check_status: bool = True):
    """Adds this batch's {TP, detections, ground truths} per class into ``tallies`` (int64 (C-1, 3)).

    With ``keep`` / ``keep_cnt`` (the lists ``nms_`` / ``postprocess_`` return under ``want_lists=True``) only the kept rows
    are read instead of the whole tensor.  ``check_status`` reads the kernel's overflow word back after the launch (one
    4-byte synchronising copy; skipped while a CUDA graph is being captured) and raises if detections were dropped --
    pass False in a pipelined loop and call ``eval_status`` once at the end."""
    lib = _lib.load()
    _need_cuda(outputs, gts, tallies, keep, keep_cnt)
    outputs, gts = _f32c(outputs), _f32c(gts)
    N, P, row = outputs.shape
    C = row - 4
    G = gts.shape[1]
    dev = outputs.device
    if gts.shape[0] != N or gts.shape[2] != row:
        raise ValueError(f"eval_accumulate: gts {tuple(gts.shape)} do not match outputs {tuple(outputs.shape)}")
    if (keep is None) != (keep_cnt is None):
        raise ValueError("eval_accumulate: keep and keep_cnt go together")
    if keep is not None:
        if keep.dtype != torch.int32 or keep_cnt.dtype != torch.int32 or tuple(keep.shape) != (N, P) or keep_cnt.numel() != N:
            raise ValueError("eval_accumulate: keep (N, P) int32 and keep_cnt (N,) int32 expected")
        keep, keep_cnt = keep.contiguous(), keep_cnt.contiguous()
    if tallies is None:
        tallies = torch.zeros(C - 1, 3, dtype=torch.int64, device=dev)
    flags = torch.empty(N, P, dtype=torch.uint8, device=dev) if want_flags else None
    if N > 0:
        with torch.cuda.device(dev):
            ws = _workspace("eval", lib.ssdh_eval_workspace_bytes(N, P, C, G), dev, zero=True)
            gptr = gts.data_ptr() if G > 0 else None
            if keep is None:
                check(lib.ssdh_eval_accumulate(outputs.data_ptr(), gptr, N, P, C, G, float(iou_thresh),
                                               tallies.data_ptr(), _ptr(flags), ws.data_ptr(), ws.numel(), _stream()),
                      "ssdh_eval_accumulate")
            else:
                check(lib.ssdh_eval_accumulate_kept(outputs.data_ptr(), keep.data_ptr(), keep_cnt.data_ptr(), gptr, N, P, C, G,
                                                    float(iou_thresh), tallies.data_ptr(), _ptr(flags), ws.data_ptr(), ws.numel(),
                                                    _stream()), "ssdh_eval_accumulate_kept")
        if check_status and keep is None and not torch.cuda.is_current_stream_capturing():
            eval_status(dev)
    return tallies, flags


def eval_status(device) -> None:
    """Raises if any ``eval_accumulate`` call on the current stream since the last check dropped detections (an image
    with more than P positive score entries: rows with several positive classes, i.e. input that skipped ``calc_score``)."""
    lib = _lib.load()
    device = torch.device(device)
    key = ("eval", device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _ws_cache.get(key)
    if ws is None:
        return
    status = ctypes.c_int(0)
    with torch.cuda.device(device):
        check(lib.ssdh_eval_status(ws.data_ptr(), ctypes.byref(status), _stream()), "ssdh_eval_status")
        if status.value != 0:
            ws[:4].zero_()
            raise _lib.SsdHeadError("ssdh_eval_accumulate: an image had more than P positive score entries (rows with several "
                                    "positive classes); detections were dropped.  Feed calc_score / postprocess output.")


# ------------------------------------------------------------------------------------------------ 8f-4 true VOC AP
def voc_ap(scores: torch.Tensor, tp: torch.Tensor, cls: torch.Tensor, tallies: torch.Tensor, use_07_metric: bool = False) -> torch.Tensor:
    """PASCAL VOC AP per class on the device (ssdh_voc_ap): ``scores`` (D,) fp32, ``tp`` (D,) 1 = true positive, ``cls`` (D,)
    0-based class, ``tallies`` (C-1, 3) int64 (ground-truth counts in column 2) -> (C-1,) fp32, NaN without ground truth."""
    lib = _lib.load()
    _need_cuda(scores, tp, cls, tallies)
    D = int(scores.numel())
    NC = int(tallies.shape[0])
    if tp.numel() != D or cls.numel() != D or tallies.dim() != 2 or tallies.shape[1] != 3 or tallies.dtype != torch.int64:
        raise ValueError("voc_ap: scores / tp / cls must have one entry per detection, tallies (C-1, 3) int64")
    scores = _f32c(scores.reshape(-1))
    tp8 = tp.reshape(-1).to(torch.uint8).contiguous()
    cls32 = cls.reshape(-1).to(torch.int32).contiguous()
    tallies = tallies.contiguous()
    out = torch.empty(NC, dtype=torch.float32, device=tallies.device)
    with torch.cuda.device(tallies.device):
        ws = _workspace("vocap", lib.ssdh_voc_ap_workspace_bytes(D, NC), tallies.device)
        check(lib.ssdh_voc_ap(scores.data_ptr() if D else None, tp8.data_ptr() if D else None, cls32.data_ptr() if D else None, D,
                              tallies.data_ptr(), NC, int(bool(use_07_metric)), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
              "ssdh_voc_ap")
    return out

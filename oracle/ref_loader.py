"""Import the UNMODIFIED reference from /root/reference/src for oracle validation.  TEST INFRASTRUCTURE ONLY.

Only usable in the build container (the reference tree does not travel to the GPU box);
callers must check ``available()`` first.  The reference uses bare top-level imports
(``from model import VGG16`` at src/model/ssd.py:5, ``from utils import ...`` at
src/evaluate.py:4), so the modules are imported with the reference's src/ at the front of
``sys.path`` and then evicted from ``sys.modules`` so that they cannot shadow anything here.
"""
from __future__ import annotations

import importlib
import os
import sys
from types import SimpleNamespace

import torch

REF_SRC = os.environ.get("SSDH_REFERENCE_SRC", "/root/reference/src")
_SHADOWED = ("model", "model.ssd", "model.vgg16", "utils", "evaluate", "dataset", "augmentation",
             "augmentation.compose", "augmentation.to_tensor", "augmentation.random")
_cache = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "model", "ssd.py"))


def load() -> SimpleNamespace:
    """Returns namespace(net, SSD, utils, evaluate).  ``net`` is an ``SSD`` built WITHOUT running its
    constructor (which downloads VGG weights, src/model/vgg16.py:68); the head methods use no state."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_SRC}")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in _SHADOWED}
    sys.path.insert(0, REF_SRC)
    try:
        ssd_mod = importlib.import_module("model.ssd")
        utils_mod = importlib.import_module("utils")
        try:
            eval_mod = importlib.import_module("evaluate")
        except Exception:            # optional deps of the script shell (tqdm, PIL); functions restated in tests
            eval_mod = None
    finally:
        sys.path.remove(REF_SRC)
        for k in list(sys.modules):
            if k in _SHADOWED:
                sys.modules.pop(k)
        sys.modules.update(saved)
    net = ssd_mod.SSD.__new__(ssd_mod.SSD)
    torch.nn.Module.__init__(net)
    _cache = SimpleNamespace(net=net, SSD=ssd_mod.SSD, utils=utils_mod, evaluate=eval_mod)
    return _cache


def reference_eval_loop(ref, outputs: torch.Tensor, gts: torch.Tensor, n_classes: int = 20):
    """Drives the reference's OWN functions through the loop body of src/evaluate.py:132-151, which is
    script code under ``__main__`` and cannot be imported.  Returns (result_correct, result_count)."""
    ious = ref.utils.calc_iou(outputs, gts)
    result_correct, result_count = {}, {c: 0 for c in range(n_classes)}
    for i, (output, gt, iou) in enumerate(zip(outputs, gts, ious)):
        result_correct[i] = {}
        for c in range(n_classes):
            po, go = ref.evaluate.get_order(output, c), ref.evaluate.get_order(gt, c)
            if len(po) == 0 and len(go) == 0:
                continue
            if len(po) == 0:
                result_count[c] += len(go)
                continue
            if len(go) == 0:
                correct = torch.zeros(len(po), 1)
            else:
                sub = iou[po][:, go]
                valid = torch.eye(len(go))[sub.max(dim=1).indices] * (sub > 0.5)
                correct = ((valid.cumsum(dim=0) == valid) * valid).sum(dim=1, keepdims=True)
            result_correct[i][c] = torch.cat([correct, output[po][:, [5 + c]]], dim=1)
            result_count[c] += len(go)
    return result_correct, result_count

"""CPU oracle for the SSD300 detection-head hot path.  TEST INFRASTRUCTURE ONLY.

This package is a from-scratch CPU restatement (torch fp32 on the host, loops over the
ground-truth axis instead of the reference's 4-D broadcasts) of the reference algorithms
in ``/root/reference/src/model/ssd.py``, ``src/utils.py`` and ``src/evaluate.py``.
Every function cites the reference lines it follows.

Who may use it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs -- and there only as the checker or the timed CPU baseline.
The product (``object_detection_torch2_b200``) never imports it and has no CPU fallback.

Pinning: the reference ships NO tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8c), so reference-side parity is *unpinned by reference tests*.  The pin we do have
is the reference's own code executed on CPU in the build container (torch 2.11.0+cu128):
``tests/golden/make_golden.py`` imports ``/root/reference/src`` and writes the fixtures in
``tests/golden/``; ``tests/test_oracle_vs_reference.py`` re-proves restatement == reference
whenever the reference tree is present, and ``tests/test_oracle_golden.py`` checks the
restatement against the committed fixtures everywhere else (including the GPU box).
"""
from .head import (  # noqa: F401
    default_boxes, pair_iou_match, match_mask, encode_offsets, smooth_l1, split_pos_neg,
    kplus1_threshold, multibox_loss, decode_boxes, class_scores, pair_iou, greedy_nms,
    nms_inplace, class_order, eval_image_class, eval_batch, average_precision, voc_ap_numpy,
)

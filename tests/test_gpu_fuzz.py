"""GPU property / fuzz tests (hypothesis): small random shapes, ties, degenerate boxes and option combinations against the
oracle -- the edge cases a fixed seed does not visit ("property" level).  Everything goes through the C ABI."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from object_detection_torch2_b200 import evaluate, ops, synth
from oracle import head
from test_gpu_loss import run_and_compare

pytestmark = pytest.mark.gpu
DEV = "cuda"
COMMON = dict(deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))


@settings(max_examples=80, **COMMON)
@given(n=st.integers(1, 3), p=st.integers(1, 420), c=st.sampled_from([2, 5, 21, 33]), g=st.integers(0, 9),
       thr=st.sampled_from([0.1, 0.25, 0.5, 0.7]), a=st.sampled_from([0.5, 1.0, 2.0]), seed=st.integers(0, 10_000),
       crowded=st.booleans())
def test_loss_fuzz(n, p, c, g, thr, a, seed, crowded, priors_cpu):
    gen = torch.Generator().manual_seed(seed)
    pri = priors_cpu[torch.randperm(8732, generator=gen)[:p]].contiguous()
    o = torch.randn(n, p, 4 + c, generator=gen)
    t = torch.zeros(n, g, 4 + c)
    if g:
        base = synth.make_targets(n, seed, max_boxes=g, num_classes=21)
        k = min(g, base.shape[1])
        t[:, :k, :4] = base[:, :k, :4]
        if crowded:
            t[:, :k, 2:4] = t[:, :k, 2:4].clamp(min=0.6)                    # big boxes: many positives, the 3 * pos > neg branch
        lab = torch.randint(0, c, (n, g), generator=gen)
        t[torch.arange(n)[:, None], torch.arange(g)[None, :], 4 + lab] = 1.0
        t[:, :, 4:] *= (t[:, :, 2:3] * t[:, :, 3:4] > 0)                     # padding rows stay all-zero
    run_and_compare(o, t, pri, pri.to(DEV), a=a, thr=thr, max_flips=2, check_ambiguous=False)


def _random_scored(gen, n, p, ties, degenerate):
    """Decoded + scored rows: boxes, one positive class score per candidate row, void rows elsewhere."""
    x = torch.zeros(n, p, 25)
    x[:, :, 0:2] = torch.rand(n, p, 2, generator=gen)
    x[:, :, 2:4] = 0.02 + 0.5 * torch.rand(n, p, 2, generator=gen) ** 2
    if ties:                                                                 # coarse grid: duplicate boxes and duplicate scores
        x[:, :, :4] = (x[:, :, :4] * 8).round() / 8 + 0.0625
    score = torch.rand(n, p, generator=gen)
    if ties:
        score = (score * 6).ceil() / 6
    cand = torch.rand(n, p, generator=gen) < 0.7
    cls = torch.randint(1, 21, (n, p), generator=gen)
    x[:, :, 4:].scatter_(2, (cls * cand).unsqueeze(2), torch.where(cand, score, torch.ones_like(score)).unsqueeze(2))
    if degenerate and p > 3:
        x[:, 0, 2] = 0.0                                                     # zero width
        x[:, 1, 3] = -0.1                                                    # negative height
        x[:, 2, :4] = torch.tensor([5.0, 5.0, 0.2, 0.2])                     # far outside the image
    return x


@settings(max_examples=80, **COMMON)
@given(n=st.integers(1, 3), p=st.integers(1, 700), thr=st.sampled_from([0.0, 0.3, 0.45, 0.5, 0.9]), ties=st.booleans(),
       degenerate=st.booleans(), per_class=st.booleans(), top_k=st.sampled_from([None, 1, 7, 200]),
       score_thresh=st.sampled_from([0.0, 0.01, 0.5]), seed=st.integers(0, 10_000))
def test_nms_fuzz(n, p, thr, ties, degenerate, per_class, top_k, score_thresh, seed):
    gen = torch.Generator().manual_seed(seed)
    x = _random_scored(gen, n, p, ties, degenerate)
    kw = dict(iou_thresh=thr, score_thresh=score_thresh, top_k=top_k, per_class=per_class)
    xd = x.to(DEV)
    res = ops.nms_(xd, want_lists=True, **kw)
    want = x.clone()
    want, _ = head.nms_inplace(want, **kw)
    for i in range(n):
        o_want, k_want = head.greedy_nms(x[i], **kw)
        assert int(res.order_cnt[i]) == o_want.numel() and int(res.keep_cnt[i]) == k_want.numel(), (kw, i)
        assert torch.equal(res.order[i, :o_want.numel()].cpu().long(), o_want), (kw, i)
        assert torch.equal(res.keep[i, :k_want.numel()].cpu().long(), k_want), (kw, i)
    assert torch.equal(xd.cpu(), want), kw
    # evaluation on the result: dense scan == kept lists == oracle
    t = synth.pad_targets(synth.make_targets(n, seed, 6), 6).to(DEV)
    dense, fl = evaluate.accumulate(xd, t, want_flags=True)
    kept, fk = evaluate.accumulate(xd, t, want_flags=True, keep=res.keep, keep_cnt=res.keep_cnt)
    want_t, _ = head.eval_batch(want, t.cpu())
    assert torch.equal(dense.cpu(), want_t) and torch.equal(kept.cpu(), want_t) and torch.equal(fl, fk)


@settings(max_examples=30, **COMMON)
@given(n=st.integers(1, 2), g=st.integers(1, 12), thr=st.sampled_from([0.25, 0.5, 0.75]), scale=st.sampled_from([0.15, 0.4, 1.0]),
       seed=st.integers(0, 10_000))
def test_force_best_prior_fuzz(n, g, thr, scale, seed, priors_cpu):
    priors_gpu = priors_cpu.to(DEV)
    t = synth.make_targets(n, seed, max_boxes=g)
    t[:, :, 2:4] *= scale
    o = synth.make_outputs(n, seed, "D2")
    want = head.match_mask(t, priors_cpu, thr, force_best_prior=True)
    got = ops.match(t.to(DEV), priors_gpu, thr, force_best_prior=True).mask.cpu()
    assert torch.equal(got, want)
    ref = head.multibox_loss(o, t, priors_cpu, threshold=thr, force_best_prior=True)
    loss, _, stats = ops.multibox_loss_raw(o.to(DEV), t.to(DEV).contiguous(), priors_gpu, threshold=thr, want_stats=True, force_best_prior=True)
    sn = ops.stats_to_numpy(stats)
    assert np.array_equal(sn["pos_raw"], ref["pos_raw"].numpy()) and np.array_equal(sn["k_pos"], ref["k_pos"].numpy())
    np.testing.assert_allclose(float(loss), float(ref["loss"]), rtol=1e-5, atol=1e-7)

"""CPU: the oracle restatement against the fixtures the reference itself produced (tests/golden)."""
import numpy as np
import pytest
import torch

import cases
from conftest import sha, unpack_bits
from oracle import head


def test_priors_bit_exact(golden, priors_cpu):
    assert priors_cpu.shape == (8732, 4)
    assert np.array_equal(priors_cpu.numpy().view(np.uint32), golden["priors"].view(np.uint32))


@pytest.mark.parametrize("case", cases.LOSS_CASES, ids=[c[0] for c in cases.LOSS_CASES])
def test_loss_cases(case, golden, priors_cpu):
    k = f"loss/{case[0]}/"
    o, t = cases.loss_inputs(case, priors_cpu)
    assert sha(o) == str(golden[k + "sha_outputs"]) and sha(t) == str(golden[k + "sha_targets"]), "input generator drifted"
    r = head.multibox_loss(o, t, priors_cpu, a=case[5], want_grad=True)
    want_match = unpack_bits(golden[k + "match_bits"], golden[k + "match_shape"])
    assert torch.equal(r["match"], want_match)
    assert np.array_equal(r["pos_raw"].numpy(), golden[k + "pos_raw"])
    assert np.array_equal(r["k_pos"].numpy(), golden[k + "k_pos"])
    assert np.array_equal(r["k_neg"].numpy(), golden[k + "k_neg"])
    np.testing.assert_allclose(r["loss"].numpy(), golden[k + "loss"], rtol=1e-6)
    np.testing.assert_allclose(r["loss"].numpy(), golden[k + "loss_nograd"], rtol=1e-6)
    g = r["grad"]
    np.testing.assert_allclose(g.reshape(-1)[::cases.GRAD_STRIDE].numpy(), golden[k + "grad_sample"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(g.abs().sum(dim=(1, 2)).double().numpy(), golden[k + "grad_abs_sum"], rtol=1e-5)
    rows = unpack_bits(golden[k + "grad_row_nonzero"], g.shape[:2])
    assert torch.equal(g.abs().sum(dim=2) > 0, rows)
    # every row with a non-zero gradient is a selected positive or negative and vice versa
    assert torch.equal(rows, r["pos_valid"] | r["neg_valid"])
    delta = torch.stack([head.encode_offsets(t[:, j, :4], priors_cpu) for j in range(t.shape[1])], dim=2)
    got = delta.reshape(-1)[::cases.GRAD_STRIDE].numpy()
    assert np.array_equal(got.view(np.uint32), golden[k + "delta_sample"].view(np.uint32))


@pytest.mark.parametrize("case", cases.POST_CASES, ids=[c[0] for c in cases.POST_CASES])
def test_post_cases(case, golden, priors_cpu):
    k = f"post/{case[0]}/"
    o, t = cases.post_inputs(case, priors_cpu)
    assert sha(o) == str(golden[k + "sha_outputs"]) and sha(t) == str(golden[k + "sha_targets"]), "input generator drifted"
    x = o.clone()
    box = head.decode_boxes(x, priors_cpu)
    x[:, :, :4] = box
    sc = head.class_scores(x)
    x[:, :, 4:] = sc
    assert np.array_equal(box.reshape(-1)[::7].numpy().view(np.uint32), golden[k + "box_sample"].view(np.uint32))
    assert np.array_equal(sc.argmax(dim=2).numpy().astype(np.uint8), golden[k + "score_argmax"])
    np.testing.assert_allclose(sc.max(dim=2).values.numpy(), golden[k + "score_max"], rtol=1e-6)
    np.testing.assert_array_equal(head.pair_iou(x[:, :64], t).numpy(), golden[k + "iou_gt_64"])
    y, keeps = head.nms_inplace(x, iou_thresh=case[4])
    assert y is x
    kept_rows = y[:, :, 4:].sum(dim=2) > 0
    assert torch.equal(kept_rows, unpack_bits(golden[k + "kept_rows"], kept_rows.shape))
    assert [len(q) for q in keeps] == golden[k + "kept_count"].tolist()
    assert sha(y) == str(golden[k + "sha_after_nms"])
    tallies, results = head.eval_batch(y, t)
    assert np.array_equal(tallies.numpy(), golden[k + "tallies"])
    flags = [torch.cat(results[c])[:, 0] for c in range(20) if results[c]]
    flat = torch.cat(flags).numpy().astype(np.uint8) if flags else np.zeros(0, np.uint8)
    assert np.array_equal(flat, golden[k + "tp_flags"])
    for c in range(20):
        if results[c]:
            ap = float(head.average_precision(torch.cat(results[c]), int(tallies[c, 2])))
            want = float(golden[k + "ap"][c])
            assert (np.isnan(ap) and np.isnan(want)) or ap == pytest.approx(want, rel=1e-6)
            if tallies[c, 2] > 0:       # closed form: the reference's AP is TP / #gt
                assert ap == pytest.approx(tallies[c, 0].item() / tallies[c, 2].item(), rel=1e-6)


def test_selection_edge_cases():
    # strict '>' against the (k+1)-th value: ties at the threshold select fewer than k (ssd.py:222-223)
    v = torch.tensor([3.0, 1.0, 1.0, 1.0, 0.5])
    assert float(head.kplus1_threshold(v, 2)) == 1.0 and int((v > head.kplus1_threshold(v, 2)).sum()) == 1
    assert float(head.kplus1_threshold(v, 0)) == 3.0 and int((v > head.kplus1_threshold(v, 0)).sum()) == 0
    kp, kn = head.split_pos_neg(torch.tensor([0, 10, 2183, 2184, 8732]), 8732)
    assert kp.tolist() == [0, 10, 2183, 2182, 0] and kn.tolist() == [0, 30, 6549, 6548, 0]


def test_nms_edge_cases(priors_cpu):
    rows = torch.zeros(6, 25)
    rows[:, :4] = torch.tensor([[.5, .5, .2, .2], [.5, .5, .2, .2], [.8, .8, .1, .1], [.5, .52, .2, .2], [.1, .1, .1, .1], [.1, .1, .1, .1]])
    rows[0, 6] = 0.9
    rows[1, 7] = 0.9          # same score as row 0: lower index wins, row 1 suppressed (class-agnostic)
    rows[2, 5] = 0.3
    rows[3, 6] = 0.95
    rows[4, 4] = 1.0          # void arg-max: never a candidate
    order, keep = head.greedy_nms(rows)
    assert order.tolist() == [3, 0, 1, 2] and keep.tolist() == [3, 2]
    order, keep = head.greedy_nms(rows, per_class=True)
    assert keep.tolist() == [3, 1, 2]
    order, keep = head.greedy_nms(rows, top_k=1)
    assert keep.tolist() == [3]
    order, keep = head.greedy_nms(torch.zeros(4, 25))
    assert order.numel() == 0 and keep.numel() == 0

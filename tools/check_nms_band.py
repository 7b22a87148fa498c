"""Brute-force check of the dense NMS sweep's division-free verdict (csrc/post.cu, nms_kernel step 1) against the reference's
own fp32 IoU arithmetic (src/utils.py:58-77 followed by `> thr`, utils.py:108).

The sweep computes, per (kept box k, candidate me), e = fma(w * h, 1 + thr, -(thr * area_me) - thr * area_k) with only the
x-extent clamped at zero, and decides from the running maximum of e over the kept list: > m -> suppressed, < -m -> not,
otherwise the IEEE division; m = area_me * (1 + thr) * (1 + 2 / thr) * 2^-18.  This script restates that arithmetic in numpy
float32 (the FMA as an exactly rounded float64 product-sum) and counts verdicts that contradict the reference on pairs placed
within +-3e-5 of the threshold (shifted copies, nested boxes either way, three scales) and on random pairs of very different
sizes.  Any WRONG count above zero is a bug.   python tools/check_nms_band.py [pairs per case]
"""
import sys
import numpy as np

f32 = np.float32


def corners(cx, cy, w, h):
    return ((cx - w / f32(2)).astype(f32), (cx + w / f32(2)).astype(f32), (cy - h / f32(2)).astype(f32),
            (cy + h / f32(2)).astype(f32), (w * h).astype(f32))


def verdicts(thr, kept, me):
    """kept / me: (cx, cy, w, h) float32 arrays.  Returns (reference hit, sure hit, sure miss)."""
    thr = f32(thr)
    ax1, ax2, ay1, ay2, aa = corners(*kept)
    bx1, bx2, by1, by2, ba = corners(*me)
    wd = np.maximum(np.minimum(ax2, bx2) - np.maximum(ax1, bx1), f32(0)).astype(f32)
    hc = np.maximum(np.minimum(ay2, by2) - np.maximum(ay1, by1), f32(0)).astype(f32)
    inter_ref = (wd * hc).astype(f32)
    uni_ref = ((aa + ba).astype(f32) - inter_ref).astype(f32)
    with np.errstate(all="ignore"):
        iou = np.where(inter_ref > 0, (inter_ref / uni_ref).astype(f32), inter_ref)
    ref_hit = iou > thr
    ht = (np.minimum(ay2, by2) - np.maximum(ay1, by1)).astype(f32)          # not clamped
    inter = (wd * ht).astype(f32)
    c1 = f32(f32(1) + thr)
    s = ((-(thr * ba)).astype(f32) - (thr * aa).astype(f32)).astype(f32)
    e = (inter.astype(np.float64) * np.float64(c1) + s.astype(np.float64)).astype(f32)
    mk = f32(f32(c1 * f32(f32(1) + f32(2) / thr)) * f32(3.814697265625e-06))
    m = (ba * mk).astype(f32)
    return ref_hit, e > m, e < -m


def run(n=1_000_000, seed=1, out=print):
    rng = np.random.default_rng(seed)
    wrong = 0
    for thr in (0.45, 0.5, 0.3, 0.9, 0.05, 0.01, 0.001):
        for sc in (1.0, 0.003, 30.0):
            cx = rng.uniform(0, 1, n).astype(f32); cy = rng.uniform(0, 1, n).astype(f32)
            w = (rng.uniform(0.01, 0.9, n) * sc).astype(f32); h = (rng.uniform(0.01, 0.9, n) * sc).astype(f32)
            t = np.float64(thr) * (1 + rng.uniform(-3e-5, 3e-5, n))
            d = w.astype(np.float64) * (1 - t) / (1 + t)
            r = np.sqrt(t)
            cases = {"shifted copy": ((cx, cy, w, h), ((cx.astype(np.float64) + d).astype(f32), cy, w, h)),
                     "candidate inside kept": ((cx, cy, w, h), (cx, cy, (w * r).astype(f32), (h * r).astype(f32))),
                     "kept inside candidate": ((cx, cy, (w * r).astype(f32), (h * r).astype(f32)), (cx, cy, w, h))}
            for tag, (kept, me) in cases.items():
                ref, hit, miss = verdicts(thr, kept, me)
                bad = int(((hit & ~ref) | (miss & ref)).sum()); wrong += bad
                out(f"thr {thr:<6} scale {sc:<6} {tag:22s}: reference hits {ref.mean():.3f}, decided without division {np.mean(hit | miss):.4f}, WRONG {bad}")
        a = [rng.uniform(0, 1, n).astype(f32) for _ in range(4)]
        s1 = (10.0 ** rng.uniform(-3.5, 0.5, n)).astype(f32); s2 = (10.0 ** rng.uniform(-3.5, 0.5, n)).astype(f32)
        ref, hit, miss = verdicts(thr, (a[0], a[1], s1, s1), (a[2], a[3], s2, s2))
        bad = int(((hit & ~ref) | (miss & ref)).sum()); wrong += bad
        out(f"thr {thr:<6} random pairs, sizes over four decades: reference hits {ref.mean():.3f}, decided without division {np.mean(hit | miss):.4f}, WRONG {bad}")
    return wrong


if __name__ == "__main__":
    total = run(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000)
    print("total WRONG", total)
    sys.exit(1 if total else 0)

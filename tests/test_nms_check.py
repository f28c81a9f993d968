"""The NMS keep-list checker itself (CPU): it must accept only differences that hinge on a near-threshold pair."""
import numpy as np

from nms_check import replay_explains


def _greedy(iou, thr):
    kept = []
    for i in range(iou.shape[0]):
        if not kept or iou[kept, i].max() <= thr:
            kept.append(i)
    return kept


def test_replay_accepts_exact_and_near_threshold_only():
    rng = np.random.default_rng(0)
    n = 300
    iou = (rng.uniform(0, 1, (n, n)) * (rng.random((n, n)) < 0.02)).astype(np.float32)  # sparse overlaps
    iou = np.triu(iou, 1); iou = iou + iou.T
    iou[np.abs(iou - 0.5) < 1e-3] = 0.6   # nothing near the threshold
    order = rng.permutation(n)
    want = _greedy(iou, 0.5)
    ok, amb, _ = replay_explains([order[i] for i in want], iou, order, 0.5)
    assert ok and amb == 0
    # a wrong list (one suppressed box kept) is rejected even though the matrix is large
    sup = next(i for i in range(n) if i not in want)
    bad = sorted(want + [sup])
    ok, _, msg = replay_explains([order[i] for i in bad], iou, order, 0.5)
    assert not ok and "kept" in msg
    # dropping a kept box is rejected as well
    ok, _, _ = replay_explains([order[i] for i in want[:5] + want[6:]], iou, order, 0.5)
    assert not ok
    # a pair inside the band may go either way, and the rest of the list must follow from that decision
    j = want[3]
    k = next(i for i in range(j + 1, n) if i in want and iou[[x for x in want if x < i], i].max() < 0.4 and
             any(iou[i, m] > 0.5 for m in range(i + 1, n)))  # k's fate changes what follows
    iou2 = iou.copy(); iou2[j, k] = iou2[k, j] = 0.5 + 4e-7
    alt = _greedy(np.where(np.abs(iou2 - 0.5) < 1e-6, 0.0, iou2), 0.5)    # k kept
    alt2 = _greedy(np.where(np.abs(iou2 - 0.5) < 1e-6, 1.0, iou2), 0.5)   # k suppressed
    assert alt != alt2
    for lst in (alt, alt2):
        ok, amb, msg = replay_explains([order[i] for i in lst], iou2, order, 0.5)
        assert ok and amb >= 1, msg
    # post_max_size truncation
    ok, _, _ = replay_explains([order[i] for i in want[:7]], iou, order, 0.5, post_max_size=7)
    assert ok

"""Checker for NMS keep lists that differ from the oracle's (test infrastructure).

A greedy NMS result may legitimately differ from the oracle only through pairs whose IoU lies within the float
tolerance of the threshold (north_star: 1e-6).  Instead of accepting any difference when SOME near-threshold pair
exists, the sweep is replayed box by box: a box with a kept predecessor of IoU > thr + tol must be suppressed, a box
whose kept predecessors all have IoU < thr - tol must be kept, and only in between -- the decision hinges on a
specific (kept, candidate) pair inside the tolerance band -- may the result follow either way.  The replay has to
reproduce the checked list exactly, so every divergence is tied to the pair that explains it and the lists agree
again once that pair's decision is taken as given.
"""
import numpy as np


def replay_explains(got, iou_sorted, order, thr, tol=1e-6, post_max_size=None):
    """got: kept original indices in keep order; iou_sorted[i, j]: IoU of the boxes at sorted positions i, j
    (descending score, after pre_max_size); order[i]: original index at sorted position i.
    Returns (ok, n_ambiguous, message)."""
    n = len(order)
    want_pos = {int(o): i for i, o in enumerate(order)}
    try:
        gpos = [want_pos[int(i)] for i in got]
    except KeyError as e:  # kept a box that is not among the candidates
        return False, 0, f"kept index {e} is not a candidate"
    gset = set(gpos)
    kept, ambiguous = [], 0
    limit = n if post_max_size is None else int(post_max_size)
    for i in range(n):
        if len(kept) >= limit:
            break
        m = float(iou_sorted[kept, i].max()) if kept else -1.0
        if m > thr + tol:
            keep = False
        elif m < thr - tol:
            keep = True
        else:
            keep = i in gset  # hinges on the pair (argmax kept box, i), |IoU - thr| <= tol
            ambiguous += 1
        if keep != (i in gset):
            j = kept[int(np.argmax(iou_sorted[kept, i]))] if kept else -1
            return False, ambiguous, (f"box at sorted position {i} (index {int(order[i])}) "
                                      f"{'kept' if i in gset else 'suppressed'} but max IoU with the kept boxes is {m!r} "
                                      f"(pair {j},{i}), threshold {thr}")
        if keep:
            kept.append(i)
    if gpos != kept:
        return False, ambiguous, "keep order differs from score order"
    return True, ambiguous, ""


def assert_keep_lists_agree(got, want, dets, thr, oracle, tol=1e-6, pre_max_size=None, post_max_size=None):
    """dets [N,6] (x,y,w,l,angle,score).  Equal lists pass; different lists must be explained pair by pair."""
    if list(got) == list(want):
        return 0
    order = oracle.argsort_desc(dets[:, 5])
    if pre_max_size is not None:
        order = order[:int(pre_max_size)]
    ds = dets[order]
    iou = oracle.rotate_iou_gpu_eval(ds[:, :5], ds[:, :5], -1)
    ok, amb, msg = replay_explains(got, iou, order, float(np.float32(thr)), tol, post_max_size)
    assert ok, "keep lists differ and no near-threshold pair explains it: " + msg
    assert amb > 0
    return amb

"""GPU: randomised configurations of the voxelizer / scatter / NMS against the CPU oracle (small sizes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_voxelizer_random_config(pp, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    D = int(rng.integers(3, 7))
    dt = np.float64 if rng.random() < 0.5 else np.float32
    grid = rng.integers(1, 40, size=3)
    vs = rng.uniform(0.05, 2.0, size=3)
    lo = rng.uniform(-20, 5, size=3)
    # ranges that are exact multiples, half-cell remainders (np.round half-to-even) or arbitrary
    extra = rng.choice([0.0, 0.5, 0.37])
    hi = lo + (grid + extra) * vs
    pcr = np.concatenate([lo, hi])
    N = int(rng.choice([0, 1, 31, 32, 33, 1000, 5000, 20000]))
    P = int(rng.choice([1, 2, 5, 31, 32, 33, 64, 65, 100]))
    cap = int(rng.choice([0, 1, 7, 100, 5000]))
    # clustered points so that some cells overflow max_points and the cap binds
    centers = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), size=(max(1, N // 50), 3))
    pts = centers[rng.integers(0, centers.shape[0], N)] + rng.normal(0, 1.5, size=(N, 3)) * vs
    if D > 3:
        pts = np.concatenate([pts, rng.random((N, D - 3))], axis=1)
    pts = np.ascontiguousarray(pts.astype(dt))
    rev = bool(rng.random() < 0.5)
    params = [(vs, pcr)]
    if dt == np.float32:
        params.append((vs.astype(np.float32), pcr.astype(np.float32)))  # float32 arithmetic
        params.append((vs.tolist(), pcr.tolist()))                      # lists -> cast to points.dtype
    for v_, r_ in params:
        got = pp.points_to_voxel(pts, v_, r_, P, rev, cap, return_point_slots=True)
        want = oracle.points_to_voxel(pts, v_, r_, P, rev, cap, return_slots=True)
        for a, b in zip(got, want):
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(8))
def test_scatter_and_nms_random(pp, oracle, synth, seed):
    rng = np.random.default_rng(2000 + seed)
    B, ny, nx, C = int(rng.integers(1, 5)), int(rng.integers(1, 70)), int(rng.integers(1, 90)), int(rng.choice([1, 3, 4, 17, 64, 129]))
    M = int(rng.choice([0, 1, 50, 3000]))
    coords = np.stack([rng.integers(0, B, M), rng.integers(0, 3, M), rng.integers(0, ny, M), rng.integers(0, nx, M)], axis=1).astype(np.int32)
    feats = rng.normal(size=(M, C)).astype(np.float32)
    for layout in ("NCHW", "NHWC"):
        assert np.array_equal(pp.scatter(feats, coords, B, ny, nx, layout), oracle.scatter(feats, coords, B, ny, nx, layout))
    n = int(rng.choice([1, 2, 63, 64, 65, 128, 129, 700, 1500]))
    d = synth.rotated_boxes(n, 3000 + seed, clustered=True)
    thr = float(rng.choice([0.05, 0.3, 0.5, 0.7]))
    pre = rng.choice([None, 10, 100, 128, 129, 1024, 1100])
    post = rng.choice([None, 1, 50, 300])
    pre = None if pre is None else int(pre)
    post = None if post is None else int(post)
    got = pp.rotate_nms_gpu(d, thr, pre_max_size=pre, post_max_size=post)
    want = oracle.rotate_nms_gpu(d, thr, pre, post)
    if got != want:
        ds = d[oracle.argsort_desc(d[:, 5])]
        iou = oracle.rotate_iou_gpu_eval(ds[:, :5], ds[:, :5], -1)
        assert (np.abs(iou - thr) < 1e-5).any(), (n, thr, pre, post)
    sb = oracle.rbox_to_standup(d[:, :5]) * np.float32(rng.choice([1.0, 20.0]))
    g2 = pp.nms(sb, d[:, 5], pre, post, thr)
    w2 = oracle.nms(sb, d[:, 5], pre, post, thr)
    assert (g2 is None and w2 is None) or g2.tolist() == w2.tolist()

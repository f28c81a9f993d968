"""Extended randomized parity run of the voxelizer (both implementations behind pp_voxelize_dev) against the CPU oracle:
random grids / caps / dtypes / point counts, single frames through the numpy drop-in and ragged multi-frame batches through the
device entry point (frames packed back to back, decoration fused).  usage: python tools/vox_fuzz.py [n_cases] [seed0]"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
PKG = "3d-object-detection-for-autonomous-navigation_b200"
pp = importlib.import_module(PKG); _lib = importlib.import_module(PKG + "._lib"); interop = importlib.import_module(PKG + ".interop")
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = time.time(); checked = {"table_path": 0, "anygrid_path": 0}; batches = 0
for case in range(n_cases):
    rng = np.random.default_rng(50_000 + seed0 + case)
    D = int(rng.integers(3, 5))
    dt = np.float64 if rng.random() < 0.5 else np.float32
    grid = rng.integers(1, 30, size=3)
    if rng.random() < 0.3: grid = np.array([80, 64, 2])
    vs = rng.uniform(0.05, 2.0, size=3)
    lo = rng.uniform(-20, 5, size=3)
    hi = lo + (grid + rng.choice([0.0, 0.5, 0.37])) * vs
    pcr = np.concatenate([lo, hi])
    P = int(rng.choice([1, 2, 5, 31, 32, 33, 50, 64, 65, 100, 128, 129, 200, 254]))
    cap = int(rng.choice([0, 1, 7, 100, 5000, 12000]))
    B = int(rng.choice([1, 1, 2, 3, 5]))
    frames = []
    for b in range(B):
        N = int(rng.choice([0, 1, 33, 1000, 20000, 40000, 70000]))
        centers = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), size=(max(1, N // 50), 3))
        pts = centers[rng.integers(0, centers.shape[0], N)] + rng.normal(0, 1.5, size=(N, 3)) * vs
        if rng.random() < 0.3 and N: pts = np.sort(pts, axis=0)           # coherent clouds: long runs of one cell
        if rng.random() < 0.2 and N: pts[rng.integers(0, N, max(1, N // 100))] = np.nan
        if D > 3: pts = np.concatenate([pts, rng.random((N, D - 3))], axis=1)
        frames.append(np.ascontiguousarray(pts.astype(dt)))
    rev = bool(rng.random() < 0.5)
    want = [oracle.points_to_voxel(f, vs, pcr, P, rev, cap, return_slots=True) for f in frames]
    for path, thr in (("table_path", 0), ("anygrid_path", 1 << 62)):
        _lib.check(_lib.lib().pp_voxelize_set_small_path_min_points(thr))
        for f, w in zip(frames, want):   # numpy drop-in, one frame
            got = pp.points_to_voxel(f, vs, pcr, P, rev, cap, return_point_slots=True)
            for a, b in zip(got, w):
                assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True), (case, path)
        if B > 1 and cap > 0 and sum(f.shape[0] for f in frames) > 0:   # device entry point, ragged batch (empty outputs have no pointer)
            allp = torch.from_numpy(np.concatenate(frames)).cuda()
            off = torch.from_numpy(np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int64)).cuda()
            out = interop.points_to_voxel(allp, vs, pcr, P, rev, cap, frame_offsets=off, decorate=True,
                                          max_frame_points=max(f.shape[0] for f in frames))
            torch.cuda.synchronize()
            vb = out["voxel_base"].cpu().numpy()
            for b, w in enumerate(want):
                lo_, hi_ = int(vb[b]), int(vb[b + 1])
                assert hi_ - lo_ == w[0].shape[0], (case, path, b)
                assert np.array_equal(out["voxels"][lo_:hi_].cpu().numpy(), w[0].astype(np.float32), equal_nan=True), (case, path, b)
                assert np.array_equal(out["coors"][lo_:hi_, 1:].cpu().numpy(), w[1]) and np.array_equal(out["num_points"][lo_:hi_].cpu().numpy(), w[2])
            batches += 1
        checked[path] += len(frames)
_lib.check(_lib.lib().pp_voxelize_set_small_path_min_points(-1))
print(f"vox_fuzz: {n_cases} random configurations, frames checked per path {checked}, {batches} ragged batches through the device entry point, "
      f"all bit-exact against the oracle ({time.time() - t0:.0f} s)")

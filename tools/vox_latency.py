"""Whole-call time of pp_voxelize_dev (CUDA events) for small batches: frames x points per frame, table path against
any-grid path (pp_voxelize_set_small_path_min_points).  Used to place the path thresholds of voxelize_small.cu."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "3d-object-detection-for-autonomous-navigation_b200"
pp = importlib.import_module(PKG)
_lib = importlib.import_module(PKG + "._lib")
pipeline = importlib.import_module(PKG + ".pipeline")
synth = pp.synth
cfg = synth.D435
L = _lib.lib()
full = synth.d435_cloud(0)
cases = [(1, n) for n in (5000, 20000, 50000, 100000, 200000, full.shape[0])] + [(f, full.shape[0]) for f in (2, 4, 8, 12, 16, 24, 32, 64)]
if len(sys.argv) > 1:
    cases = [(int(a.split("x")[0]), int(a.split("x")[1])) for a in sys.argv[1:]]
for F, n in cases:
    step = max(1, full.shape[0] // n)
    fr = np.ascontiguousarray(full[::step][:n])
    n = fr.shape[0]
    pts = torch.from_numpy(np.concatenate([fr] * F)).cuda()
    off = (torch.arange(F + 1, dtype=torch.int64) * n).cuda()
    L.pp_voxelize_set_small_path_min_points(C.c_int64(0))  # the workspace is sized for the paths the batch is eligible for
    pipe = pipeline.FramePipeline(cfg, max_frames=F, max_total_points=F * n, max_frame_points=n)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for name, thr in (("table", 0), ("anygrid", 1 << 60)):
        L.pp_voxelize_set_small_path_min_points(C.c_int64(thr))
        for _ in range(3):
            pipe.voxelize(pts, off, F, F * n, n, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            pipe.voxelize(pts, off, F, F * n, n, st)
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 20 * 1000
    print(f"frames {F:3d} x {n:7d} points: table {out['table']:7.1f} us   any-grid {out['anygrid']:7.1f} us")
    del pipe

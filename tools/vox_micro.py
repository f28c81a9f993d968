"""Micro-benchmark of pp_voxelize_dev on a batch of D435 frames (per-kernel CUDA-event times)."""
import ctypes as C, importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "3d-object-detection-for-autonomous-navigation_b200"
pp = importlib.import_module(PKG); _lib = importlib.import_module(PKG + "._lib"); pipeline = importlib.import_module(PKG + ".pipeline")
synth = pp.synth; cfg = synth.D435; F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
fr = [synth.d435_cloud(i) for i in range(4)]; n = fr[0].shape[0]
pts = torch.from_numpy(np.concatenate([fr[i % 4] for i in range(F)])).cuda()
off = (torch.arange(F + 1, dtype=torch.int64) * n).cuda()
pipe = pipeline.FramePipeline(cfg, max_frames=F, max_total_points=F * n, max_frame_points=n)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _once in (0,):
    for _ in range(3): pipe.voxelize(pts, off, F, F * n, n, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): pipe.voxelize(pts, off, F, F * n, n, st)
    e1.record(); torch.cuda.synchronize()
    whole = e0.elapsed_time(e1) / 10 * 1000
    _lib.profile_start()
    for _ in range(3): pipe.voxelize(pts, off, F, F * n, n, st)
    acc = {}
    for k, v in _lib.profile_stop(): acc.setdefault(k, []).append(v)
    print("whole call %.0f us;" % whole, "per launch:", {k: round(1000 * float(np.mean(v))) for k, v in acc.items()})

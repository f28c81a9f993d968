"""Per-kernel CUDA-event times of the whole device-resident path for a named config
(d435 | kitti), batch F frames.  BASELINE.json configs[2] = `kitti 64`."""
import importlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "3d-object-detection-for-autonomous-navigation_b200"
pp = importlib.import_module(PKG); _lib = importlib.import_module(PKG + "._lib"); pipeline = importlib.import_module(PKG + ".pipeline")
synth = pp.synth
name = sys.argv[1] if len(sys.argv) > 1 else "kitti"; F = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = synth.KITTI if name == "kitti" else synth.D435
gen = (lambda i: synth.kitti_cloud(i, shuffled=bool(i & 1))) if name == "kitti" else synth.d435_cloud
fr = [gen(i) for i in range(8)]; n = fr[0].shape[0]
pts = torch.from_numpy(np.concatenate([fr[i % 8] for i in range(F)])).cuda()
off = (torch.arange(F + 1, dtype=torch.int64) * n).cuda()
amask = len(sys.argv) > 3 and sys.argv[3] == "amask"
pipe = pipeline.FramePipeline(cfg, max_frames=F, max_total_points=F * n, max_frame_points=n, overlap_post=False, anchor_area_threshold=1 if amask else None)
A = pipe.A
box = torch.from_numpy(np.stack([synth.rpn_standin(A, i % 8)[0] for i in range(F)])).cuda()
sco = torch.from_numpy(np.stack([synth.rpn_standin(A, i % 8)[1] for i in range(F)])).cuda()
feats = torch.from_numpy(synth.pfn_standin(pipe.cap_rows, cfg["num_filters"], 0)).cuda()
run = lambda: pipe.run(pts, off, F, F * n, n, feats, box, sco)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
_lib.profile_start()
for _ in range(3): run()
acc = {}
for k, v in _lib.profile_stop(): acc.setdefault(k, []).append(v)
M = int(pipe.voxel_base[F].item()) / F
nx, ny, _ = synth.grid_size(cfg); P, D, C = cfg["max_points"], cfg["num_point_features"], cfg["num_filters"]
s_in = 8 if cfg["point_dtype"] == "float64" else 4
alg = n * D * s_in + M * P * D * 4 + M * 16 + M * P * (D + 5) * 4 + M * C * 4 + M * 16 + C * ny * nx * 4
km = {k: round(1000 * float(np.mean(v))) for k, v in acc.items()}
vs_us = sum(v for k, v in km.items() if k.startswith("vox_") or k.startswith("scatter_"))
print(json.dumps({"config": f"{name} x{F}", "ms_per_step": ms, "frames_per_s": F / ms * 1e3, "points_per_s": F * n / ms * 1e3, "pillars_per_frame": M,
                  "kernel_us": km, "voxelize_scatter_alg_MB_per_frame": alg / 1e6, "voxelize_scatter_GBs": alg * F / (vs_us * 1e-6) / 1e9}))

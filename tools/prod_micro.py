"""Per-kernel CUDA-event times of the production chain (ingest -> voxelize -> scatter -> anchor mask -> predict)."""
import importlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "3d-object-detection-for-autonomous-navigation_b200"
pp = importlib.import_module(PKG); _lib = importlib.import_module(PKG + "._lib"); pipeline = importlib.import_module(PKG + ".pipeline")
synth = pp.synth; cfg = synth.D435; F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rot = len(sys.argv) > 2 and sys.argv[2] == "rotated"
n = 848 * 480
pipe = pipeline.FramePipeline(cfg, max_frames=F, rotated_nms=rot, anchor_area_threshold=1, production=True, sensor_points=n)
A = pipe.A
sens = [synth.d435_sensor_cloud(i) for i in range(4)]
cloud = torch.from_numpy(np.stack([sens[i % 4] for i in range(F)])).cuda()
rng = np.random.default_rng(7)
bp = torch.from_numpy(rng.normal(0, 0.1, (F, A, 7)).astype(np.float32)).cuda()
cl = torch.from_numpy(rng.normal(-2, 1, (F, A, 1)).astype(np.float32)).cuda()
dr = torch.from_numpy(rng.normal(0, 1, (F, A, 2)).astype(np.float32)).cuda()
rect = torch.eye(4).repeat(F, 1, 1).cuda(); trv = torch.eye(4).repeat(F, 1, 1).cuda()
feats = torch.from_numpy(synth.pfn_standin(pipe.cap_rows, cfg["num_filters"], 0)).cuda()
run = lambda: pipe.run_production(cloud, F, 12, (0, 4, 8), feats, bp, cl, dr, rect, trv)
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
print(json.dumps({"config": f"production d435 x{F}", "ms_per_step": ms, "frames_per_s": F / ms * 1e3,
                  "kernel_us": {k: round(1000 * float(np.mean(v))) for k, v in acc.items()}}))

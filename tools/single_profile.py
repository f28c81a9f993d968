"""Per-launch CUDA-event times of ONE d435i frame through the whole path (voxelize + decorate, scatter, decode + NMS)."""
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
cloud = synth.d435_cloud(0)
n = cloud.shape[0]
pts = torch.from_numpy(cloud).cuda()
off = torch.tensor([0, n], dtype=torch.int64).cuda()
pipe = pipeline.FramePipeline(cfg, max_frames=1, max_total_points=n, rotated_nms=True, layout="NCHW", fused_decorate=True, keep_voxels=True)
A = pipe.A
box, sco = synth.rpn_standin(A, 0)
d_box = torch.from_numpy(box[None]).cuda()
d_sco = torch.from_numpy(sco[None]).cuda()
d_feats = torch.from_numpy(synth.pfn_standin(pipe.cap_rows, cfg["num_filters"] if "num_filters" in cfg else 128, 0)).cuda()


def step():
    pipe.run(pts, off, 1, n, n, d_feats, d_box, d_sco)


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    step()
e1.record(); torch.cuda.synchronize()
print("step %.1f us" % (e0.elapsed_time(e1) / 50 * 1000))
_lib.profile_start()
for _ in range(5):
    step()
acc = {}
for k, v in _lib.profile_stop():
    acc.setdefault(k, []).append(v)
print({k: round(1000 * float(np.mean(v)), 1) for k, v in acc.items()})

"""Per-launch CUDA-event times of the large-N NMS stripe kernels (select / mask / sweep / push) for 100 000 boxes x 32 frames."""
import ctypes as C, importlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "3d-object-detection-for-autonomous-navigation_b200"
pp = importlib.import_module(PKG); _lib = importlib.import_module(PKG + "._lib")
N, B = 100000, 32
L = _lib.lib(); dev = torch.device("cuda", 0)
dets = np.stack([pp.synth.rotated_boxes(N, 500 + i, False) for i in range(4)])
dets = np.concatenate([dets] * (B // 4 + 1))[:B]
boxes = torch.from_numpy(np.ascontiguousarray(dets[:, :, :5])).to(dev); scores = torch.from_numpy(np.ascontiguousarray(dets[:, :, 5])).to(dev)
ws_bytes = int(L.pp_nms_workspace_bytes(1, B, N, -1)); ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
keep = torch.empty((B, N), dtype=torch.int32, device=dev); cnt = torch.zeros(B, dtype=torch.int32, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    _lib.check(L.pp_nms_dev(1, C.c_void_p(boxes.data_ptr()), 5, C.c_void_p(scores.data_ptr()), None, B, N, -1, -1, 0.5, C.c_void_p(keep.data_ptr()), N, C.c_void_p(cnt.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes, st))
run(); torch.cuda.synchronize()
_lib.profile_start(); run(); rec = _lib.profile_stop()
for name in ("nms_push", "nms_stripe_mask", "nms_stripe_sweep", "nms_select"):
    v = [round(t * 1000) for k, t in rec if k == name]
    print(name, len(v), v[:40], "...", v[-5:])

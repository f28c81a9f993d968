"""BASELINE.json configs[3]: rotated BEV NMS stress, 100k boxes per frame, IoU 0.5, batch 32.
Device-resident timing of pp_nms_dev (sort + prep + mask + sweep), frames processed in chunks that
fit the mask (1.25 GB per frame)."""
import ctypes as C, importlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "3d-object-detection-for-autonomous-navigation_b200"
pp = importlib.import_module(PKG); _lib = importlib.import_module(PKG + "._lib")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
CH = int(sys.argv[3]) if len(sys.argv) > 3 else 8
clustered = len(sys.argv) > 4 and sys.argv[4] == "clustered"
L = _lib.lib(); dev = torch.device("cuda", 0)
dets = np.stack([pp.synth.rotated_boxes(N, 500 + i, clustered) for i in range(min(B, 4))])
dets = np.concatenate([dets] * (B // dets.shape[0] + 1))[:B]
boxes = torch.from_numpy(np.ascontiguousarray(dets[:, :, :5])).to(dev); scores = torch.from_numpy(np.ascontiguousarray(dets[:, :, 5])).to(dev)
ws_bytes = int(L.pp_nms_workspace_bytes(_lib.PP_NMS_ROTATED, CH, N, -1))
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
keep = torch.empty((B, N), dtype=torch.int32, device=dev); cnt = torch.zeros(B, dtype=torch.int32, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    for b0 in range(0, B, CH):
        nb = min(CH, B - b0)
        _lib.check(L.pp_nms_dev(_lib.PP_NMS_ROTATED, C.c_void_p(boxes[b0].data_ptr()), 5, C.c_void_p(scores[b0].data_ptr()), None, nb, N,
                                -1, -1, 0.5, C.c_void_p(keep[b0].data_ptr()), N, C.c_void_p(cnt[b0:].data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes, st))
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
_lib.profile_start(); run(); rec = _lib.profile_stop()
acc = {}
for k, v in rec: acc[k] = acc.get(k, 0.0) + v
print(json.dumps({"config": f"rotated NMS {N} boxes x batch {B} (chunks of {CH}), clustered={clustered}", "ms_total": ms, "ms_per_frame": ms / B,
                  "frames_per_s": B / ms * 1e3, "pairs_per_s": B * (N * (N - 1) / 2) / ms * 1e3, "kept_per_frame": float(cnt.float().mean()),
                  "mask_bytes_per_frame": N * ((N + 63) // 64) * 8, "kernel_ms": {k: round(v, 3) for k, v in acc.items()}}))

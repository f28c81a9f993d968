"""Timing of the NMS order stage (top-k of long score lists) alone: pp_nms_dev with pre_max 1000 on N anchors x B frames,
per-kernel CUDA-event times."""
import ctypes as C, importlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "3d-object-detection-for-autonomous-navigation_b200"
pp = importlib.import_module(PKG); _lib = importlib.import_module(PKG + "._lib")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 107136
L = _lib.lib(); dev = torch.device("cuda", 0)
for B in (1, 8, 64):
    d = pp.synth.rotated_boxes(N, 5, False)
    boxes = torch.from_numpy(np.ascontiguousarray(np.tile(d[None, :, :5], (B, 1, 1)))).to(dev)
    scores = torch.from_numpy(np.ascontiguousarray(np.tile(d[None, :, 5], (B, 1)))).to(dev)
    wsb = int(L.pp_nms_workspace_bytes(1, B, N, 1000)); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    keep = torch.empty((B, 300), dtype=torch.int32, device=dev); cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    run = lambda: _lib.check(L.pp_nms_dev(1, C.c_void_p(boxes.data_ptr()), 5, C.c_void_p(scores.data_ptr()), None, B, N, 1000, 300, 0.5,
                                          C.c_void_p(keep.data_ptr()), 300, C.c_void_p(cnt.data_ptr()), C.c_void_p(ws.data_ptr()), wsb, st))
    for _ in range(3): run()
    torch.cuda.synchronize(); _lib.profile_start()
    for _ in range(5): run()
    acc = {}
    for k, v in _lib.profile_stop(): acc.setdefault(k, []).append(v)
    print(B, {k: round(1000 * float(np.mean(v))) for k, v in acc.items()})

"""One d435i frame through pp_decode_nms_dev (the single-frame post stage): CUDA-event time per call; meant to be run
under ncu for the per-line profile of nms_small (tools/ncu_lines.py)."""
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
synth = pp.synth
L = _lib.lib()
cfg = synth.D435
an = torch.from_numpy(synth.anchors_stride(cfg)).cuda()
A = an.shape[0]
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
enc, sc = synth.rpn_standin(A, 0)
enc = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(enc, (F, A, 7)))).cuda()
sc = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(sc, (F, A)))).cuda()
p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
for kind, name in ((_lib.PP_NMS_ROTATED, "rotated"), (_lib.PP_NMS_STANDUP, "standup")):
    for pre, post in ((100, 50), (128, 128)):
        nb = int(L.pp_nms_workspace_bytes(kind, F, A, pre))
        ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
        keep = torch.empty((F, post), dtype=torch.int32, device="cuda")
        cnt = torch.empty((F,), dtype=torch.int32, device="cuda")
        dets = torch.empty((F, post, 8), device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def run():
            _lib.check(L.pp_decode_nms_dev(kind, p(enc), p(an), A, p(sc), None, F, A, pre, post, 0.5, p(keep), post, p(cnt), None, 0,
                                           p(ws), nb, st))
        for _ in range(5):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            run()
        e1.record(); torch.cuda.synchronize()
        print(f"{name} pre={pre} post={post} frames={F}: {e0.elapsed_time(e1) / reps * 1000:.1f} us/call, kept {cnt.tolist()[:2]}, first {keep[0, :5].tolist()}")

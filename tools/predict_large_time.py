"""Times pp_predict_dev on KITTI-sized heads (A = 107 136 anchors) for the reference's 100-box selection and for a
KITTI-style 1000 / 300 selection (the general decode + NMS path).  Device tensors, CUDA events."""
import ctypes as C
import importlib
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
PKG = "3d-object-detection-for-autonomous-navigation_b200"
_lib = importlib.import_module(PKG + "._lib")
pred = importlib.import_module(PKG + ".predict")
L = _lib.lib()
A = 107136
for B in (1, 8):
    g = torch.Generator(device="cuda").manual_seed(1)
    bp = torch.randn((B, A, 7), device="cuda", generator=g) * 0.2
    cl = torch.randn((B, A, 1), device="cuda", generator=g) * 1.5 - 1
    dr = torch.randn((B, A, 2), device="cuda", generator=g)
    xy = torch.rand((A, 2), device="cuda", generator=g) * torch.tensor([69.12, 79.36], device="cuda") + torch.tensor([0, -39.68], device="cuda")
    an = torch.cat([xy, torch.full((A, 1), -1.0, device="cuda"), torch.tensor([[1.6, 3.9, 1.56]], device="cuda").expand(A, 3),
                    (torch.arange(A, device="cuda") % 2).float()[:, None] * 1.5708], 1).contiguous()
    rect = torch.eye(4, device="cuda").repeat(B, 1, 1).contiguous()
    trv = torch.eye(4, device="cuda").repeat(B, 1, 1).contiguous()
    for top_k, pre, post in ((100, 100, 50), (1000, 1000, 300)):
        for rot in (False, True):
            cfg = pred.make_cfg(1, True, top_k, pre, post, 0.5, 0.05, rot, False)
            K = post
            ws_bytes = int(L.pp_predict_workspace_bytes(C.byref(cfg), B, A, K))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
            lid = torch.empty((B, K, 7), device="cuda"); cam = torch.empty((B, K, 7), dtype=torch.float64, device="cuda")
            sc = torch.empty((B, K), device="cuda"); lab = torch.empty((B, K), dtype=torch.int32, device="cuda")
            idx = torch.empty((B, K), dtype=torch.int32, device="cuda"); cnt = torch.empty((B,), dtype=torch.int32, device="cuda")
            p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
            st = torch.cuda.current_stream().cuda_stream

            def run():
                _lib.check(L.pp_predict_dev(C.byref(cfg), p(bp), p(cl), p(dr), p(an), None, p(rect), p(trv), B, A, K, p(lid), p(cam),
                                            p(sc), p(lab), p(idx), p(cnt), p(ws), ws_bytes, C.c_void_p(st)))
            for _ in range(5):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                run()
            e1.record(); torch.cuda.synchronize()
            print(f"B={B} top_k={top_k} pre={pre} post={post} rotated={rot}: {e0.elapsed_time(e1) / 50 * 1000:.1f} us/call, kept {cnt.tolist()[:4]}")

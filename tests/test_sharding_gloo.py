"""CPU, world_size 2 over gloo: frame sharding and the host-side gather of final detections (the only
cross-rank step of the path).  No GPU, no compute kernels."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def test_shard_frames_partition():
    pipeline = importlib.import_module(PKG + ".pipeline")
    for n in (0, 1, 7, 64, 512, 513):
        for w in (1, 2, 3, 4, 8):
            spans = [pipeline.shard_frames(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, n_frames, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pipeline = importlib.import_module(PKG + ".pipeline")
    first, count = pipeline.shard_frames(n_frames, world, rank)
    # stand-in for the per-rank GPU result: detection k of frame f carries (f, k)
    dets = np.zeros((count, 5, 8), np.float32)
    cnts = np.zeros((count,), np.int32)
    for i in range(count):
        f = first + i
        cnts[i] = f % 5
        for k in range(cnts[i]):
            dets[i, k] = [f, k, 0, 0, 0, 0, 0, 0.5]
    all_d, all_c = pipeline.gather_detections(dets, cnts, world)
    if rank == 0:
        d = np.concatenate(all_d); c = np.concatenate(all_c)
        np.savez(out_path, d=d, c=c)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_detections_world2(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "gathered.npz")
    n_frames = 11
    mp.spawn(_worker, args=(2, port, n_frames, out), nprocs=2, join=True)
    g = np.load(out)
    assert g["d"].shape == (n_frames, 5, 8)
    for f in range(n_frames):
        assert g["c"][f] == f % 5
        for k in range(f % 5):
            assert g["d"][f, k, 0] == f and g["d"][f, k, 1] == k

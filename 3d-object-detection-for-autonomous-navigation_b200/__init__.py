"""B200-native PointPillars pre/post-processing hot path (sm_100a CUDA behind a C ABI).

Drop-in Python surface (same names, argument meaning and return types as the reference):

    points_to_voxel        load_data.py:695-771
    PillarFeatureNet.decorate / pillar_decorate   model/pointpillars.py:143-203
    PointPillarsScatter    model/pointpillars.py:240-341
    second_box_decode      libraries/eval_helper_functions.py:388-461
    rbox_to_standup        load_data.py:1525-1594 + 1330-1341 (as used at model/voxelnet.py:1233-1249)
    nms, nms_gpu           libraries/eval_helper_functions.py:463-527
    rotate_nms_gpu, rotate_iou_gpu, rotate_iou_gpu_eval   second/core/non_max_suppression/nms_gpu.py
    anchors_mask           load_data.py:3043-3072 ("next" row N1)
    pointcloud2_to_lidar   load_data.py:2434-2443, sensor ingest of the production path ("next" row N3)
    predict                model/voxelnet.py:1060-1389, post-network half of VoxelNet.predict ("next" row N2)

The compute lives in libpp_b200.so (csrc/*.cu, include/pp_b200.h).  There is no CPU fallback.
"""
from ._lib import PPError, Ctx, ctx, grid_size, launch_count, lib, pinned_empty, set_device  # noqa: F401
from .anchors import anchors_mask  # noqa: F401
from .boxes import rbox_to_standup, second_box_decode  # noqa: F401
from .nms import (bev_box_overlap, d3_box_overlap, nms, nms_gpu, rotate_iou_gpu, rotate_iou_gpu_eval,  # noqa: F401
                  rotate_nms_gpu)
from .pillars import PillarFeatureNet, PointPillarsScatter, pillar_decorate, scatter  # noqa: F401
from .ingest import pointcloud2_to_lidar  # noqa: F401
from .predict import predict, predict_arrays  # noqa: F401
from .voxelizer import points_to_voxel  # noqa: F401
from . import synth  # noqa: F401

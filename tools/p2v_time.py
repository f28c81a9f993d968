import importlib, sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pp = importlib.import_module("3d-object-detection-for-autonomous-navigation_b200")
synth = pp.synth; cfg = synth.D435
vs, pcr = np.array(cfg["voxel_size"]), np.array(cfg["point_cloud_range"])
full = synth.d435_cloud(0)
for _ in range(5): pp.points_to_voxel(full, vs, pcr, 50, True, 12000)
t = []
for _ in range(40):
    t0 = time.perf_counter(); pp.points_to_voxel(full, vs, pcr, 50, True, 12000); t.append(time.perf_counter() - t0)
print("points_to_voxel pageable median %.3f ms  min %.3f ms" % (np.median(t) * 1e3, np.min(t) * 1e3))

#!/usr/bin/env python
"""Per-point cycle breakdown of k_sv_points (library built with -DISDF_PHASE_TIMING, path in ISDF_B200_LIB).
usage: dbg_swept_phases.py <robot: mesh|ShapeName>"""
import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import isdf_b200 as I, workloads as W
robot = sys.argv[1] if len(sys.argv) > 1 else "mesh"
X = 256
occ = W.random_map(X, X, X, p=0.02, seed=2, slabs=3)
cfg = I.default_config_values(); cfg.flags = I.WITH_DYNAMICS
T, Cc, wp = W.make_trajectory(64, [0, 0, 0], [X, X, X], seed=11, jitter=0.2)
pts = W.gather_obstacle_points(occ, [0, 0, 0], 1.0, wp, cfg.kernel_size / 3.0)
V, F = W.rounded_cone_mesh()
ev = I.Evaluator(cfg)
if robot == "mesh": ev.set_shape_mesh(V, F, [0, 0, 0, 120, 0, 0])
else: ev.set_shape_named(robot)
ev.set_points(pts)
ev.lib.isdf_dbg_enable(ev.h, 1)
for _ in range(3): ev.eval_swept(T, Cc)
P = len(pts)
out = np.zeros(8 * P, dtype=np.uint64)
ev.lib.isdf_dbg_swept_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
rc = ev.lib.isdf_dbg_swept_stats(ev.h, out.ctypes.data, 8 * P)
d = out.reshape(P, 8).astype(float)
print("rc", rc, "P", P, "kernel ms", ev.stats().last_kernel_ms)
names = ["total", "coarse", "bracket", "exact", "descent", "intervals", "exact searches"]
for k, n in enumerate(names): print(f"{n:>15}: mean {d[:, k].mean():12.0f}  p50 {np.median(d[:, k]):12.0f}  p90 {np.percentile(d[:, k], 90):12.0f}  max {d[:, k].max():12.0f}")
w = int(np.argmax(d[:, 0])); print("slowest point:", dict(zip(names, d[w, :7])))

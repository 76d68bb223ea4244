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
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if world > 1: ev.set_shard(0, world)
ev.lib.isdf_dbg_enable(ev.h, 1)
for _ in range(3): ev.eval_swept(T, Cc)
P = len(pts)
out = np.zeros(12 * P, dtype=np.uint64)
ev.lib.isdf_dbg_swept_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
rc = ev.lib.isdf_dbg_swept_stats(ev.h, out.ctypes.data, 12 * P)
d = out[:8 * P].reshape(P, 8).astype(float)
d2 = out[8 * P:].reshape(P, 4).astype(float)
requery = (out[:8 * P].reshape(P, 8)[:, 5] >> 16).astype(float)
d[:, 5] = (out[:8 * P].reshape(P, 8)[:, 5] & 0xffff).astype(float)
keep = d[:, 0] > 0
d, d2, requery = d[keep], d2[keep], requery[keep]
print("rc", rc, "P", P, "kernel ms", ev.stats().last_kernel_ms)
names = ["total", "coarse", "bracket", "exact", "descent", "intervals", "exact searches", "descent rounds"]
for k, n in enumerate(names): print(f"{n:>15}: mean {d[:, k].mean():12.0f}  p50 {np.median(d[:, k]):12.0f}  p90 {np.percentile(d[:, k], 90):12.0f}  max {d[:, k].max():12.0f}")
w = int(np.argmax(d[:, 0])); print("slowest point:", dict(zip(names, d[w, :8])))
tot = np.sort(d[:, 0])[::-1]
print("top 12 totals (k cycles):", (tot[:12] / 1e3).round(0).tolist())
print("sum of totals / (148 SMs) = %.0f k cycles per SM if perfectly packed at one CTA per SM; points with total > 200k: %d" % (tot.sum() / 148 / 1e3, int((tot > 2e5).sum())))
order = np.argsort(-d[:, 0])[:12]
for w in order: print("  point %4d:" % w, {n: int(v) for n, v in zip(names, d[w, :8])})
hv = d[:, 7] > 0
print("descent: cycles per round (points with a descent): mean %.0f ; rounds mean %.1f max %.0f" % ((d[hv, 4] / d[hv, 7]).mean(), d[hv, 7].mean(), d[:, 7].max()))
he = d[:, 6] > 0
print("exact pass: cycles per exact search x 8 warps (points with one): mean %.0f ; searches mean %.1f max %.0f" % ((8 * d[he, 3] / d[he, 6]).mean(), d[he, 6].mean(), d[:, 6].max()))
hv = d[:, 7] > 0
print("descent sub-phases of warp 0 (cycles per batch): pose at x %.0f, re-query at x %.0f, candidate pose+query %.0f, wait+replay %.0f ; re-queries per batch %.2f" % tuple(
    [(d2[hv, q] / d[hv, 7]).mean() for q in range(4)] + [(requery[hv] / d[hv, 7]).mean()]))

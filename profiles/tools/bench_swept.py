#!/usr/bin/env python
"""Diagnostic timing of the swept-volume path (BASELINE configs[3]: 256^3 map, 64-piece trajectory, obstacle points gathered
around the waypoints). usage: bench_swept.py <robot: mesh|ShapeName> [max_points] [cpu]"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import isdf_b200 as I, workloads as W
robot = sys.argv[1] if len(sys.argv) > 1 else "Torus"
maxp = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
X = 256
occ = W.random_map(X, X, X, p=0.02, seed=2, slabs=3)
cfg = I.default_config_values(); cfg.flags = I.WITH_DYNAMICS
T, Cc, wp = W.make_trajectory(64, [0, 0, 0], [X, X, X], seed=11, jitter=0.2)
pts = W.gather_obstacle_points(occ, [0, 0, 0], 1.0, wp, cfg.kernel_size / 3.0)[:maxp]
V, F = W.rounded_cone_mesh()
ev = I.Evaluator(cfg)
if robot == "mesh": ev.set_shape_mesh(V, F, [0, 0, 0, 120, 0, 0])
else: ev.set_shape_named(robot)
ev.set_points(pts)
for _ in range(3): r = ev.eval_swept(T, Cc)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); r = ev.eval_swept(T, Cc); ts.append(time.perf_counter() - t0)
st = ev.stats()
print(f"{robot}: P={len(pts)} Ttot={T.sum():.1f}s  e2e {1e3*np.median(ts):.3f} ms  kernels {st.last_kernel_ms:.3f} ms  sdf evals {st.last_sdf_evals:,}  -> {st.last_sdf_evals/st.last_kernel_ms/1e6:.2f} G ref-equivalent evals/s  cost {r[0]:.6g}")
if len(sys.argv) > 3:
    import oracle_lib as O
    oc = O.config_from(cfg); oc.threads_num = os.cpu_count()
    sh = O.Shape.mesh(V, F, [0, 0, 0, 120, 0, 0], wn_mode=O.WN_BH) if robot == "mesh" else O.Shape.named(robot)
    n = min(len(pts), 400)
    t0 = time.perf_counter(); o = O.eval_swept(oc, sh, T, Cc, pts[:n], use_omp=True); dt = time.perf_counter() - t0
    print(f"   CPU OpenMP x{oc.threads_num}: {n} points {dt:.2f} s -> full {dt*len(pts)/n:.2f} s/eval ; GPU speed-up {dt*len(pts)/n/(st.last_kernel_ms*1e-3):.0f}x")

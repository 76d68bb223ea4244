import sys, ctypes as C, numpy as np
sys.path.insert(0, 'implicit-sdf-planner_b200/py'); sys.path.insert(0, '.')
import isdf_b200 as I, bench
w, cfg, occ, T, Cc, V, F = bench.make_workload(False)
ev = I.Evaluator(cfg); ev.set_map_u8(occ, [0,0,0], 1.0)
robot = sys.argv[1] if len(sys.argv) > 1 else "mesh"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ev.set_shard(0, world)
if robot == "mesh": ev.set_shape_mesh(V, F, w["poly_params"])
else: ev.set_shape_named(robot)
ev.lib.isdf_dbg_enable(ev.h, 1)
for _ in range(4): ev.eval_discrete(T, Cc); print("kernel ms", ev.stats().last_kernel_ms)
S = w["pieces"] * (w["samples_per_piece"] + 1)
out = np.zeros(3 * S, dtype=np.uint64)
ev.lib.isdf_dbg_sample_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
r = ev.lib.isdf_dbg_sample_stats(ev.h, out.ctypes.data, 3 * S)
d = out.reshape(S, 3).astype(np.float64)[0::world]
cyc, pairs, q = d[:, 0], d[:, 1], d[:, 2]
print("rc", r, "kernel ms", ev.stats().last_kernel_ms)
print("cycles: mean %.0f median %.0f p99 %.0f max %.0f  (us at 1.9GHz: mean %.1f max %.1f)" % (cyc.mean(), np.median(cyc), np.percentile(cyc, 99), cyc.max(), cyc.mean()/1900, cyc.max()/1900))
print("pairs: mean %.1f max %.0f ; queries: mean %.2f median %.0f p99 %.0f max %.0f total %.0f" % (pairs.mean(), pairs.max(), q.mean(), np.median(q), np.percentile(q,99), q.max(), q.sum()))
print("sum cycles / (148*12 slots) -> %.1f us ; corr(cycles, queries) %.3f" % (cyc.sum()/(148*12)/1900, np.corrcoef(cyc, q)[0,1] if q.std()>0 else 0))
if q.sum() > 0:
    A = np.vstack([q, np.ones_like(q)]).T
    sl, ic = np.linalg.lstsq(A, cyc, rcond=None)[0]
    print("cycles ~ %.0f * queries + %.0f" % (sl, ic))

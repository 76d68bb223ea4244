import sys, ctypes as C, numpy as np
sys.path.insert(0, 'implicit-sdf-planner_b200/py'); sys.path.insert(0, '.')
import isdf_b200 as I, bench
w, cfg, occ, T, Cc, V, F = bench.make_workload(False)
ev = I.Evaluator(cfg); ev.set_map_u8(occ, [0,0,0], 1.0); ev.set_shape_mesh(V, F, w["poly_params"])
ev.lib.isdf_dbg_enable(ev.h, 1)
for _ in range(4): ev.eval_discrete(T, Cc)
S = w["pieces"] * (w["samples_per_piece"] + 1)
out = np.zeros(3 * S, dtype=np.uint64)
ev.lib.isdf_dbg_sample_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
ev.lib.isdf_dbg_sample_stats(ev.h, out.ctypes.data, 3 * S)
d = out.reshape(S, 3)
tot = d[:, 0].astype(float); rest = (d[:, 1] & ((1 << 40) - 1)).astype(float); nq = (d[:, 1] >> 40).astype(float)
tq = (d[:, 2] & ((1 << 40) - 1)).astype(float); npairs = (d[:, 2] >> 40).astype(float)
ok = (rest > 0)
print("kernel ms", ev.stats().last_kernel_ms, "whole samples", ok.sum())
print("mean cycles: total %.0f  pose+scan %.0f  queries %.0f  finish %.0f" % (tot[ok].mean(), rest[ok].mean(), tq[ok].mean(), (tot - rest - tq)[ok].mean()))
print("per query %.0f cycles (sum tq / sum nq), queries/sample %.2f, pairs/sample %.1f" % (tq[ok].sum() / max(nq[ok].sum(), 1), nq[ok].mean(), npairs[ok].mean()))

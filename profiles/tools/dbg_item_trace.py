"""Per-work-item timeline of the discrete scan kernel (library built as is; isdf_dbg_enable turns the trace on).
usage: dbg_item_trace.py [world]"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, 'implicit-sdf-planner_b200/py'); sys.path.insert(0, '.')
import isdf_b200 as I, bench
import torch
w, cfg, occ, T, Cc, V, F = bench.make_workload(False)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ev = I.Evaluator(cfg); ev.set_map_u8(occ, [0, 0, 0], 1.0)
ev.set_shard(0, world)
ev.set_shape_mesh(V, F, w["poly_params"])
iters = bench.make_iterates(w, T, Cc, 6)
for k in range(4): ev.eval_discrete(T, iters[k])
ev.lib.isdf_dbg_enable(ev.h, 1)
ev.eval_discrete(T, iters[4])
print("kernel ms (scan+epilogue)", ev.stats().last_kernel_ms)
n, parts = ev.dbg_item_stats()
st = ev.lib.isdf_dbg_trace_stride()
out = np.zeros(st * n, dtype=np.uint64)
ev.lib.isdf_dbg_item_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
print("rc", ev.lib.isdf_dbg_item_trace(ev.h, out.ctypes.data, st * n))
raw = out.reshape(n, st)
t = raw[:, :2].astype(np.float64)
ok = t[:, 0] > 0
t0 = t[ok, 0].min()
b, e = (t[ok, 0] - t0) / 1e3, (t[ok, 1] - t0) / 1e3
d = e - b
print(f"items {n} (split parts {parts}), traced {ok.sum()}; kernel span {e.max():.1f} us; item duration: mean {d.mean():.1f} median {np.median(d):.1f} p99 {np.percentile(d, 99):.1f} max {d.max():.1f} us")
print("sum of item durations / span = avg busy warps:", d.sum() / e.max())
order = np.argsort(-e)[:12]
print("last finishers (slot, begin, end, dur):", [(int(np.nonzero(ok)[0][i]), round(b[i], 1), round(e[i], 1), round(d[i], 1)) for i in order])
print("begin of last-started item:", b.max(), "; items started after 50% of span:", int((b > 0.5 * e.max()).sum()))
hist, edges = np.histogram(e, bins=10, range=(0, e.max()))
print("finish-time histogram:", hist.tolist())
hist, edges = np.histogram(b, bins=10, range=(0, e.max()))
print("start-time histogram:", hist.tolist())

if st >= 8:   # -DISDF_PHASE_TIMING build: cycles per phase
    r = raw[ok].astype(np.float64)
    pose, cull, search, tail, allc = r[:, 2], r[:, 3], r[:, 4], r[:, 5], r[:, 6]
    nq = (raw[ok, 7] >> np.uint64(32)).astype(np.float64); npairs = (raw[ok, 7] & np.uint64(0xffffffff)).astype(np.float64)
    prod = allc - pose - cull - search - tail
    f = lambda x: f"mean {x.mean():8.0f} median {np.median(x):8.0f} p99 {np.percentile(x, 99):8.0f}"
    print("cycles per item  total   :", f(allc)); print("                 pose    :", f(pose)); print("                 producer:", f(prod))
    print("                 cull    :", f(cull)); print("                 search  :", f(search)); print("                 tails   :", f(tail))
    print("queries per item: mean %.2f ; pairs per item mean %.1f ; search cycles per query %.0f" % (nq.mean(), npairs.mean(), search.sum() / max(nq.sum(), 1)))

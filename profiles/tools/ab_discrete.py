#!/usr/bin/env python
"""A/B of discrete-kernel build variants on the bench workload (BASELINE configs[2]) in ONE process per variant.
usage (on the GPU box): python profiles/tools/ab_discrete.py lib1.so[:label] lib2.so ...   -> one line per variant:
  ms/eval steady state (work items learned from the previous identical evaluation), ms/eval natural order (first-evaluation state),
  analytic Ball ms, parity of cost / gradient against the first variant."""
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(lib):
    os.environ["ISDF_B200_LIB"] = lib
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py"))
    import numpy as np
    import torch
    import bench as B
    import isdf_b200 as I
    w, cfg, occ, T, Cc, V, F = B.make_workload(False)
    N = w["pieces"]
    dev = torch.device("cuda", 0)
    ev = I.Evaluator(cfg, device=0)
    ev.set_map_u8(occ, [0, 0, 0], 1.0)
    d_T, d_C = torch.from_numpy(T).to(dev), torch.from_numpy(Cc).to(dev)
    d_out = torch.zeros(19 * N + 1, dtype=torch.float64, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def timeit(steps=20, warm=4):
        for _ in range(warm):
            ev.eval_discrete_device(N, d_T.data_ptr(), d_C.data_ptr(), d_out.data_ptr(), stream)
        ts = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(); ev.eval_discrete_device(N, d_T.data_ptr(), d_C.data_ptr(), d_out.data_ptr(), stream); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.mean(ts), min(ts)
    res = {"lib": os.path.basename(lib)}
    ev.set_shape_mesh(V, F, w["poly_params"])
    res["mesh_ms"], res["mesh_ms_min"] = timeit()
    out = d_out.cpu().numpy().copy()
    try:
        res["items"] = ev.dbg_item_stats()
    except Exception:
        res["items"] = None
    try:
        ev.dbg_schedule(natural_order=True)
        res["mesh_natural_ms"], _ = timeit(steps=10, warm=2)
        nat = d_out.cpu().numpy().copy()
        res["natural_vs_items_rel"] = float(np.linalg.norm(nat - out) / np.linalg.norm(out))
        ev.dbg_schedule(natural_order=False)
    except Exception as e:
        res["mesh_natural_ms"] = repr(e)
    import workloads as W
    R, t = W.rotation_from_poly_params(w["poly_params"])
    ev.set_shape_named("Ball", R, t)
    res["ball_ms"], _ = timeit(steps=10, warm=3)
    res["cost"] = float(out[0]); res["gnorm"] = float(np.linalg.norm(out[1:]))
    np.save(os.path.join(ROOT, "gpurun_out", "ab_" + os.path.basename(lib) + ".npy"), out)
    print("ABRESULT " + json.dumps(res), flush=True)


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2])
    import numpy as np
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rows, base = [], None
    for lib in sys.argv[1:]:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", os.path.abspath(lib)], capture_output=True, text=True)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("ABRESULT ")]
        if not line:
            print(f"{lib}: FAILED\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}")
            continue
        r = json.loads(line[0][9:])
        out = np.load(os.path.join(ROOT, "gpurun_out", "ab_" + os.path.basename(lib) + ".npy"))
        if base is None:
            base = out
        r["rel_vs_first"] = float(np.linalg.norm(out - base) / np.linalg.norm(base))
        rows.append(r)
        print(json.dumps(r), flush=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ab_discrete.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

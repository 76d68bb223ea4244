#!/usr/bin/env python
"""Swept-volume evaluation (BASELINE configs[3], mesh robot) under the four timing protocols {same, distinct iterate} x {warm, flushed L2},
for the whole point set and for one shard of a world of 8 (no exchange). usage: swept_protocols.py [steps]"""
import sys, os, statistics
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import isdf_b200 as I
import bench as B
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
w, cfg, T, Cc, pts, V, F = B.swept_workload()
N = w["pieces"]
dev = torch.device("cuda", 0)
iters = B.make_iterates(w, T, Cc, steps + 4)
d_T, d_Cs = torch.from_numpy(T).to(dev), torch.from_numpy(iters).to(dev)
d_out = torch.zeros(19 * N + 1, dtype=torch.float64, device=dev)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for world in (1, 8):
    ev = I.Evaluator(cfg, device=0)
    ev.set_shape_mesh(V, F, w["poly_params"])
    ev.set_points(pts)
    ev.set_shard(0, world)
    for k in range(3): ev.eval_swept_device(N, d_T.data_ptr(), d_Cs[k].data_ptr(), d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    for distinct in (False, True):
        for fl in (False, True):
            ts = []
            for k in range(steps):
                if fl: flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(); ev.eval_swept_device(N, d_T.data_ptr(), d_Cs[(3 + k) if distinct else 0].data_ptr(), d_out.data_ptr(), stream); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(f"world {world} shard 0: {'distinct' if distinct else 'same    '} iterate, L2 {'flushed' if fl else 'warm   '}: mean {statistics.mean(ts):.3f} ms  min {min(ts):.3f}  max {max(ts):.3f}", flush=True)

#!/usr/bin/env python
"""Minimal driver for ncu: the swept-volume evaluation of BASELINE configs[3] (mesh robot), `steps` times, nothing else.
usage: prof_swept.py [steps] [--shard=rank/world]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py"))
import torch          # noqa: E402
import bench as B     # noqa: E402
import isdf_b200 as I  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
w, cfg, T, Cc, pts, V, F = B.swept_workload()
N = w["pieces"]
dev = torch.device("cuda", 0)
ev = I.Evaluator(cfg, device=0)
ev.set_shape_mesh(V, F, w["poly_params"])
ev.set_points(pts)
for a in sys.argv:
    if a.startswith("--shard="):
        r_, w_ = a[8:].split("/")
        ev.set_shard(int(r_), int(w_))
iters = B.make_iterates(w, T, Cc, steps)
d_T, d_Cs = torch.from_numpy(T).to(dev), torch.from_numpy(iters).to(dev)
d_out = torch.zeros(19 * N + 1, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for k in range(steps):
    ev.eval_swept_device(N, d_T.data_ptr(), d_Cs[k].data_ptr(), d_out.data_ptr(), stream)
    torch.cuda.synchronize()
print("cost", float(d_out[0].item()))
ev.close()

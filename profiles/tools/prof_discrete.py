#!/usr/bin/env python
"""Minimal driver for ncu: the bench workload's discrete evaluation (device-resident), `steps` times, nothing else.
usage: prof_discrete.py [steps] [robot: mesh | <analytic name>] [--natural]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py"))
import torch          # noqa: E402
import bench as B     # noqa: E402
import isdf_b200 as I  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
robot = sys.argv[2] if len(sys.argv) > 2 else "mesh"
w, cfg, occ, T, Cc, V, F = B.make_workload(False)
N = w["pieces"]
dev = torch.device("cuda", 0)
ev = I.Evaluator(cfg, device=0)
ev.set_map_u8(occ, [0, 0, 0], 1.0)
if robot == "mesh":
    ev.set_shape_mesh(V, F, w["poly_params"])
else:
    import workloads as W
    R, t = W.rotation_from_poly_params(w["poly_params"])
    ev.set_shape_named(robot, R, t)
for a in sys.argv:
    if a.startswith("--shard="):   # --shard=rank/world: one rank's shard, no exchange
        r_, w_ = a[8:].split("/")
        ev.set_shard(int(r_), int(w_))
if "--natural" in sys.argv:
    ev.dbg_schedule(natural_order=True)
iters = B.make_iterates(w, T, Cc, steps)
d_T, d_Cs = torch.from_numpy(T).to(dev), torch.from_numpy(iters).to(dev)
d_out = torch.zeros(19 * N + 1, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for k in range(steps):
    ev.eval_discrete_device(N, d_T.data_ptr(), d_Cs[k].data_ptr(), d_out.data_ptr(), stream)
    torch.cuda.synchronize()
print("cost", float(d_out[0].item()))
ev.close()

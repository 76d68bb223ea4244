#!/usr/bin/env python
"""Front-end feasibility of the bench map (512^3 x 121 attitudes): two-pass (default) vs one-pass (ISDF_FE_ONE_PASS=1) kernels."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py"))
import torch
import bench as B
import isdf_b200 as I
w, cfg, occ, T, Cc, V, F = B.make_workload(False)
dev = torch.device("cuda", 0)
ev = I.Evaluator(cfg, device=0)
ev.set_map_u8(occ, [0, 0, 0], 1.0)
ev.set_shape_mesh(V, F, w["poly_params"])
for mode in ("tables", "two-pass", "one-pass"):
    if mode == "two-pass": os.environ["ISDF_FE_NO_TABLES"] = "1"
    if mode == "one-pass": os.environ["ISDF_FE_ONE_PASS"] = "1"
    r = B.frontend_bench(ev, w, occ, V, F, dev, False)
    print(mode, json.dumps({k: r[k] for k in ("ms", "voxels_per_s", "voxels_with_a_fitting_attitude")}), "frac", r["roofline"]["frac"], flush=True)

import os, statistics, sys
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py"))
import torch, bench as B, isdf_b200 as I
w, cfg, occ, T, Cc, V, F = B.make_workload(False)
N = w["pieces"]; dev = torch.device("cuda", 0)
ev = I.Evaluator(cfg, device=0); ev.set_map_u8(occ, [0, 0, 0], 1.0); ev.set_shape_mesh(V, F, w["poly_params"])
iters = B.make_iterates(w, T, Cc, 40)
d_T, d_Cs = torch.from_numpy(T).to(dev), torch.from_numpy(iters).to(dev)
d_out = torch.zeros(19 * N + 1, dtype=torch.float64, device=dev)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for world in (1, 2, 4, 8):
    for mult in (1, 2, 4, 8):
        ev.set_shard(0, world)
        ev.dbg_schedule(warp_slots=2368 * mult)
        for k in range(4): ev.eval_discrete_device(N, d_T.data_ptr(), d_Cs[k].data_ptr(), d_out.data_ptr(), stream)
        ts = []
        for k in range(4, 24):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(); ev.eval_discrete_device(N, d_T.data_ptr(), d_Cs[k].data_ptr(), d_out.data_ptr(), stream); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"world {world} slots x{mult}: {statistics.mean(ts)*1e3:7.1f} us  items {ev.dbg_item_stats()}", flush=True)

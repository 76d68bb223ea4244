"""Multi-GPU check of the peer-memory reduction (isdf_peer_*): run with torchrun, one rank per GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_peer.py
Every rank evaluates its shard; the fused exchange must give the same all-rank sum as NCCL (to rounding), the same as the
unsharded evaluation (to rounding), and bit-identical vectors on every rank. Prints 'PEER OK' on success (rank 0)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "implicit-sdf-planner_b200", "py"))
import isdf_b200 as I
import workloads as W
from common import small_case, rel_l2, BMIN, MESHES


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg, occ, _, _, _ = small_case(N=6, K=32, seed=3)
    N = 6
    ev = I.Evaluator(cfg, device=local)
    ev.set_map_u8(occ, BMIN, 1.0)
    V, F = MESHES["rcone"]()
    ev.set_shape_mesh(V, F, [0, 0, 0, 120, 0, 0])
    stream = torch.cuda.current_stream().cuda_stream
    n = 19 * N + 1
    trajs = [W.make_trajectory(N, [0, 0, 0], [50, 50, 34], seed=40 + k, jitter=0.3) for k in range(5)]
    # references: unsharded, and sharded + NCCL
    full, nccl = [], []
    for (T, Cc, _) in trajs:
        dT, dC = torch.from_numpy(T).to(dev), torch.from_numpy(Cc).to(dev)
        out = torch.zeros(n, dtype=torch.float64, device=dev)
        ev.set_shard(0, 1)
        ev.eval_discrete_device(N, dT.data_ptr(), dC.data_ptr(), out.data_ptr(), stream)
        torch.cuda.synchronize()
        full.append(out.cpu().numpy().copy())
        ev.set_shard(rank, world)
        ev.eval_discrete_device(N, dT.data_ptr(), dC.data_ptr(), out.data_ptr(), stream)
        dist.all_reduce(out)
        torch.cuda.synchronize()
        nccl.append(out.cpu().numpy().copy())
    # connect
    h = ev.peer_export(world, n)
    handles = [None] * world
    dist.all_gather_object(handles, h)
    ev.peer_connect(world, rank, handles, fuse=True)
    dist.barrier()
    ok = True
    for rep in range(3):                                  # several epochs: both buffer parities, repeated
        for k, (T, Cc, _) in enumerate(trajs):
            dT, dC = torch.from_numpy(T).to(dev), torch.from_numpy(Cc).to(dev)
            out = torch.zeros(n, dtype=torch.float64, device=dev)
            ev.eval_discrete_device(N, dT.data_ptr(), dC.data_ptr(), out.data_ptr(), stream)
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            r1, r2 = rel_l2(got, nccl[k]), rel_l2(got, full[k])
            gathered = [torch.zeros_like(out) for _ in range(world)]
            dist.all_gather(gathered, out)
            same = all(torch.equal(gathered[0], g) for g in gathered)
            if not (r1 < 1e-13 and r2 < 1e-12 and same):
                ok = False
                print(f"rank {rank} rep {rep} traj {k}: vs nccl {r1:.2e} vs full {r2:.2e} identical-across-ranks {same}", flush=True)
    # host-buffer API accumulates the all-rank sums
    T, Cc, _ = trajs[0]
    c, gC, gT = ev.eval_discrete(T, Cc)
    if not (abs(c - full[0][0]) <= 1e-12 * abs(full[0][0]) and rel_l2(np.concatenate([gC, gT]), full[0][1:]) < 1e-12):
        ok = False; print(f"rank {rank}: host-buffer call mismatch", flush=True)
    # swept-volume path
    cfg2, occ2, T2, C2, wp2 = small_case(N=4, K=16, seed=5, noise=0.05)
    pts = W.gather_obstacle_points(occ2, BMIN, 1.0, wp2, cfg2.kernel_size * cfg2.occupancy_resolution / 3.0)[:300]
    ev.set_points(pts)
    # (unsharded reference needs the exchange off: use a second context)
    ev2 = I.Evaluator(cfg, device=local)
    ev2.set_shape_mesh(V, F, [0, 0, 0, 120, 0, 0]); ev2.set_points(pts)
    ref = ev2.eval_swept(T2, C2)
    ev2.close()
    ev.set_shard(rank, world)
    c, gC, gT = ev.eval_swept(T2, C2)
    if not (abs(c - ref[0]) <= 1e-12 * max(abs(ref[0]), 1) and rel_l2(np.concatenate([gC, gT]), np.concatenate([ref[1], ref[2]])) < 1e-12):
        ok = False; print(f"rank {rank}: swept mismatch {c} vs {ref[0]}", flush=True)
    # stand-alone exchange of an arbitrary vector
    v = torch.arange(n, dtype=torch.float64, device=dev) * (rank + 1) + 0.25 * rank
    w = v.clone()
    ev.peer_allreduce_device(v.data_ptr(), n, stream)
    dist.all_reduce(w)
    torch.cuda.synchronize()
    if not torch.allclose(v, w, rtol=1e-15, atol=0):
        ok = False; print(f"rank {rank}: stand-alone exchange mismatch", flush=True)
    ev.peer_status()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    ev.peer_disconnect()
    ev.close()
    if rank == 0:
        print("PEER OK" if int(flag.item()) == 1 else "PEER FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())

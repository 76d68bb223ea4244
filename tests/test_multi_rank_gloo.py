"""N>1 host logic on CPU: world_size-2 gloo run of the sharding + all-reduce harness that bench.py uses for --gpus N.
Each rank evaluates the samples s % world == rank (here through the oracle's shard mode standing in for the GPU),
all-reduces the 19N+1 vector and must reproduce the unsharded result."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from common import small_case, BMIN
    import bench
    cfg, occ, T, Cc, _ = small_case(N=3, K=10, seed=4)
    oc = O.config_from(cfg)
    sh = O.Shape.named("Ball")
    c, gC, gT, _ = O.eval_discrete(oc, occ, BMIN, 1.0, sh, T, Cc, rank=rank, world=world)
    out = torch.from_numpy(np.concatenate([[c], gC, gT]))
    bench.allreduce_partials(out)                      # the same call bench.py makes on the GPU tensors
    full = O.eval_discrete(oc, occ, BMIN, 1.0, sh, T, Cc)
    ref = np.concatenate([[full[0]], full[1], full[2]])
    err = float(np.linalg.norm(out.numpy() - ref) / np.linalg.norm(ref))
    tmax = bench.max_over_ranks(float(rank + 1))
    if rank == 0:
        q.put((err, tmax))
    dist.destroy_process_group()


def test_two_rank_shard_allreduce_reproduces_full_eval():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    err, tmax = q.get(timeout=5)
    assert err < 1e-12
    assert tmax == 2.0      # max over ranks, not rank 0's own time

#!/usr/bin/env python
"""bench.py — collision cost+gradient evaluations/s of the discrete hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through libisdf_b200.so)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (OpenMP structure) on host cores

Workload (N=1 and every N: strong scaling, total work fixed): BASELINE.json configs[2] — the configuration the
north_star target is quoted on: 512^3 random voxel map (Bernoulli 5 % + wall slabs), 64-piece MINCO trajectory,
256 samples/piece (S = 16448 pose samples), robot = 3900-triangle closed mesh (rounded cone, the mesh-SDF path).
One "step" = one cost + gradC (6N x 3) + gradT (N) evaluation of the collision term over the whole trajectory.
At N > 1 every rank evaluates the pose samples s % N == rank and the 19N+1 doubles are all-reduced over NCCL.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py"))

METRIC = "collision_cost_grad_evals_per_s"
UNIT = "evals/s"
WORKLOAD = dict(map_dim=512, occupancy=0.05, pieces=64, samples_per_piece=256, kernel_size=13, mesh="rounded_cone_3900tri",
                poly_params=[0.0, 0.0, 0.0, 120.0, 0.0, 0.0])


# ---- distributed helpers (also exercised by tests/test_multi_rank_gloo.py on CPU/gloo) ------------------------------
def dist_ready():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def allreduce_partials(t):
    """sum the per-rank partial [cost | gradC | gradT] vectors in place (NCCL over NVLink on GPU tensors, gloo on CPU)."""
    if dist_ready():
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def max_over_ranks(x):
    if not dist_ready():
        return float(x)
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist_ready():
        import torch.distributed as dist
        dist.barrier()


# ---- workload ------------------------------------------------------------------------------------------------------------
def make_workload(small=False):
    import isdf_b200 as I
    import workloads as W
    w = dict(WORKLOAD)
    if small:
        w.update(map_dim=128, pieces=8, samples_per_piece=32)
    X = w["map_dim"]
    occ = W.random_map(X, X, X, p=w["occupancy"], seed=1, slabs=3)
    cfg = I.default_config_values()
    cfg.integral_intervs = w["samples_per_piece"]
    cfg.kernel_size = w["kernel_size"]
    cfg.flags = I.WITH_COLLISION | I.WITH_DYNAMICS
    T, Cc, wp = W.make_trajectory(w["pieces"], [0, 0, 0], [X, X, X], seed=11, jitter=0.2)
    V, F = W.rounded_cone_mesh()
    return w, cfg, occ, T, Cc, V, F


def algorithmic_bytes(w):
    """S*W^3 bytes of occupancy (1 B/voxel, each sample's window counted once) + coeff/T read + grads written."""
    N, K, Wk = w["pieces"], w["samples_per_piece"], w["kernel_size"]
    S = N * (K + 1)
    return S * Wk ** 3 + 8 * 19 * N + 8 * (19 * N + 1)


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            txt, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            return out
        sm, mx, reasons = [], [], set()
        for line in txt.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            # the first samples precede the GPU work: report the median of the upper half (clocks under load)
            sm_sorted = sorted(sm)
            out = {"sm_mhz": statistics.median(sm_sorted[len(sm_sorted) // 2:]), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "sm_mhz_min": min(sm)}
        return out


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_capture(key="k_discrete_dram_bytes_per_launch"):
    """a figure of the dominant kernel from the committed ncu --set full capture (profiles/roofline_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


def ncu_traffic():
    return ncu_capture("k_discrete_dram_bytes_per_launch")


# ---- CPU arm: the reference's algorithm (oracle port, OpenMP `parallel for schedule(dynamic)` + `critical`) -----------------
def usable_cores():
    """threads the CPU arms may really use: logical CPUs, narrowed by the affinity mask and a cgroup CPU quota (a container that sees 128
    CPUs but is throttled to a few cores' worth of time runs SLOWER with 128 threads — the CPU arm should get its best shot)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return max(1, n)


def cpu_sample_eval(w, cfg, occ, T, Cc, V, F, pieces, threads):
    """one evaluation restricted to `pieces` pieces spread evenly over the trajectory (a bounded, representative sample of the same
    workload — the term couples nothing across pieces); returns seconds"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O     # bench's CPU legs are one of the three places allowed to execute oracle/
    N = w["pieces"]
    sel = np.unique(np.linspace(0, N - 1, pieces).round().astype(int)) if pieces < N else np.arange(N)
    assert len(sel) == pieces
    Cm = Cc.reshape(3, 6 * N)
    sub = np.concatenate([np.concatenate([Cm[ax, 6 * i:6 * i + 6] for i in sel]) for ax in range(3)])
    oc = O.config_from(cfg)
    oc.threads_num = threads
    if not hasattr(cpu_sample_eval, "shape"):
        cpu_sample_eval.shape = O.Shape.mesh(V, F, w["poly_params"], wn_mode=O.WN_BH)
    t0 = time.perf_counter()
    r = O.eval_discrete(oc, occ, [0, 0, 0], 1.0, cpu_sample_eval.shape, np.ascontiguousarray(T[sel]), sub, use_omp=True)
    return time.perf_counter() - t0, r


def cpu_baseline(w, cfg, occ, T, Cc, V, F, budget_s=20.0):
    cores = usable_cores()
    pieces = 1
    dt, _ = cpu_sample_eval(w, cfg, occ, T, Cc, V, F, pieces, cores)
    # grow the sample until it is worth ~budget/2 of CPU time, never beyond the full trajectory
    while dt < budget_s / 4 and pieces < w["pieces"]:
        pieces = min(w["pieces"], pieces * 2)
        dt, _ = cpu_sample_eval(w, cfg, occ, T, Cc, V, F, pieces, cores)
    evals_per_s = 1.0 / (dt * w["pieces"] / pieces)
    # context: one thread, and the README's 1.5 x nproc oversubscription (README.md:148), on bounded samples of the same workload
    dt1, _ = cpu_sample_eval(w, cfg, occ, T, Cc, V, F, 1, 1)
    dto, _ = cpu_sample_eval(w, cfg, occ, T, Cc, V, F, pieces, int(1.5 * cores))
    return {"value": evals_per_s, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{pieces} of {w['pieces']} pieces, evenly spread ({pieces * (w['samples_per_piece'] + 1)} pose samples) of the same workload, "
                      f"{dt:.2f} s wall with {cores} OpenMP threads (schedule(dynamic) + critical, g++ -O3), scaled by pieces",
            "one_thread_evals_per_s": 1.0 / (dt1 * w["pieces"]), "oversubscribed_1p5x_evals_per_s": 1.0 / (dto * w["pieces"] / pieces),
            "note": "the critical section is 20 additions per pose sample against ~10^2 us of SDF work: a lock-free CPU variant would not move these numbers"}


# ---- secondary metric: L-BFGS iterations/s over the whole callback (MINCO -> swept-volume term -> time integral -> adjoint) -------
def lbfgs_workload():
    import isdf_b200 as I
    import workloads as W
    X = 256
    occ = W.random_map(X, X, X, p=0.05, seed=2, slabs=3)
    cfg = I.default_config_values()
    cfg.flags = I.WITH_DYNAMICS                      # the live reference callback: collision through the swept-volume term
    N = 64
    wp = W.random_walk_waypoints(N, [0, 0, 0], [X, X, X], seed=11)
    pts = W.gather_obstacle_points(occ, [0, 0, 0], 1.0, wp, cfg.kernel_size * cfg.occupancy_resolution / 3.0)
    head, tail = np.zeros((3, 3)), np.zeros((3, 3))
    head[:, 0], tail[:, 0] = wp[0], wp[-1]
    x0 = np.concatenate([np.full(N, 1.2), wp[1:-1].reshape(-1)])   # tau = 1.2 -> T = 2.92 s per piece
    return cfg, N, wp, pts, head, tail, x0, "Torus_big"


def lbfgs_ours(device, max_iterations=40):
    import ctypes as CT
    import isdf_b200 as I
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import host_lib as H
    cfg, N, wp, pts, head, tail, x0, shape = lbfgs_workload()
    ev = I.Evaluator(cfg, device=device)
    ev.set_shape_named(shape)
    ev.set_points(pts)
    L = H.lib()
    hh, tt = np.asfortranarray(head), np.asfortranarray(tail)
    out = {}
    for rep in range(2):                                           # first run warms up allocations / work-item tables
        be = L.isdf_host_backend_create(ev.h, N, hh.ctypes.data_as(H.dp), tt.ctypes.data_as(H.dp), 20.0, 1, 1)
        x = x0.copy()
        fx, it, evs = CT.c_double(0), CT.c_int(0), CT.c_int(0)
        t0 = time.perf_counter()
        r = L.isdf_host_lbfgs_backend(be, x.ctypes.data_as(H.dp), x.size, CT.byref(fx), 16, 10, 1e-6, 0.0, max_iterations, CT.byref(it), CT.byref(evs))
        dt = time.perf_counter() - t0
        L.isdf_host_backend_destroy(be)
        out = {"iters_per_s": it.value / dt, "callback_evals_per_s": evs.value / dt, "iterations": it.value, "evaluations": evs.value,
               "seconds": dt, "ret": r, "final_cost": fx.value,
               "config": f"256^3 map p=0.05, 64 pieces, {len(pts)} obstacle points, robot {shape}, callback = MINCO + swept-volume term + "
                         "time-integral dynamics + adjoint (back_end_optimizer.hpp:358-430), L-BFGS mem 16 past 10 (config_CappedCone.yaml:99-102)"}
    ev.close()
    return out


def lbfgs_cpu(max_iterations=2):
    """same driver, callback assembled from the oracle (OpenMP, all host threads)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import host_lib as H
    import oracle_lib as O
    cfg, N, wp, pts, head, tail, x0, shape = lbfgs_workload()
    oc = O.config_from(cfg)
    oc.threads_num = usable_cores()
    sh = O.Shape.named(shape)
    rho = 20.0

    def fun(x):
        tau = x[:N]
        T = np.where(tau > 0, (0.5 * tau + 1) * tau + 1, 1 / ((0.5 * tau - 1) * tau + 1))
        inP = x[N:].reshape(-1, 3).T
        co, energy, gC, gT = O.minco_forward(head, tail, inP, T)
        sv = O.eval_swept(oc, sh, T, co, pts, use_omp=True)
        di = O.eval_discrete(oc, None, [0, 0, 0], 1.0, None, T, co, use_omp=True)
        cost = energy + sv["cost"] + di[0] + rho * T.sum()
        gp, gt = O.minco_backward(head, tail, inP, T, gC + sv["gradC"] + di[1], gT + sv["gradT"] + di[2])
        gt = gt + rho
        gtau = np.where(tau > 0, gt * (tau + 1), gt * (1 - tau) / ((0.5 * tau - 1) * tau + 1) ** 2)
        return cost, np.concatenate([gtau, gp.T.reshape(-1)])
    t0 = time.perf_counter()
    r = H.lbfgs_minimize(fun, x0, mem_size=16, past=10, delta=1e-6, g_epsilon=0.0, max_iterations=max_iterations)
    dt = time.perf_counter() - t0
    return {"iters_per_s": r["iterations"] / dt, "callback_evals_per_s": r["evaluations"] / dt, "iterations": r["iterations"],
            "evaluations": r["evaluations"], "seconds": dt, "cores": oc.threads_num}


# ---- secondary metric: the swept-volume term alone (BASELINE configs[3]), mesh robot = the reference's live default ---------------
def swept_ours(device, with_cpu, cpu_points=423):
    import isdf_b200 as I
    import workloads as W
    X = 256
    occ = W.random_map(X, X, X, p=0.02, seed=2, slabs=3)
    cfg = I.default_config_values()
    cfg.flags = I.WITH_DYNAMICS
    T, Cc, wp = W.make_trajectory(64, [0, 0, 0], [X, X, X], seed=11, jitter=0.2)
    pts = W.gather_obstacle_points(occ, [0, 0, 0], 1.0, wp, cfg.kernel_size / 3.0)
    V, F = W.rounded_cone_mesh()
    poly = [0, 0, 0, 120, 0, 0]
    ev = I.Evaluator(cfg, device=device)
    ev.set_shape_mesh(V, F, poly)
    ev.set_points(pts)
    for _ in range(3):
        ev.eval_swept(T, Cc)
    ks, es = [], []
    for _ in range(10):
        t0 = time.perf_counter(); ev.eval_swept(T, Cc); es.append(time.perf_counter() - t0)
        ks.append(ev.stats().last_kernel_ms)
    st = ev.stats()
    out = {"config": f"BASELINE configs[3]: 256^3 map, 64-piece trajectory ({T.sum():.0f} s), {len(pts)} obstacle points, 3900-triangle mesh robot; "
                     "SV-SDF query (coarse 0.2 s scan, fine 0.02 s scan, sign descent) + chain rule per point",
           "kernel_ms": statistics.median(ks), "e2e_ms": 1e3 * statistics.median(es), "points_per_s": len(pts) / (statistics.median(ks) * 1e-3),
           "reference_equivalent_sdf_evals": int(st.last_sdf_evals)}
    ev.close()
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        oc = O.config_from(cfg); oc.threads_num = usable_cores()
        sh = O.Shape.mesh(V, F, poly, wn_mode=O.WN_BH)
        n = min(len(pts), cpu_points)
        idx = np.linspace(0, len(pts) - 1, n).astype(int)          # spread over the trajectory: per-point work is very uneven
        O.eval_swept(oc, sh, T, Cc, pts[idx[:32]], use_omp=True)      # wake the OpenMP pool
        dts = []
        for _ in range(2):
            t0 = time.perf_counter(); O.eval_swept(oc, sh, T, Cc, pts[idx], use_omp=True); dts.append(time.perf_counter() - t0)
        dt = min(dts)                                                 # the faster run: the CPU arm gets the benefit of the doubt
        out["cpu"] = {"ms_full_estimate": 1e3 * dt * len(pts) / n, "cores": oc.threads_num, "sample": f"{n} of {len(pts)} points, OpenMP oracle port, best of 2"}
        out["speedup_kernel_vs_cpu"] = out["cpu"]["ms_full_estimate"] / out["kernel_ms"]
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w, cfg, occ, T, Cc, V, F = make_workload(args.small)
    cores = usable_cores()
    pieces = max(1, min(args.ref_pieces, w["pieces"]))
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_sample_eval(w, cfg, occ, T, Cc, V, F, pieces, cores)
    times = []
    for _ in range(args.steps):
        dt, _ = cpu_sample_eval(w, cfg, occ, T, Cc, V, F, pieces, cores)
        times.append(dt)
    ms = 1e3 * statistics.mean(times) * w["pieces"] / pieces
    val = 1e3 / ms
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(w), **w},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": (f"each step = the whole workload ({pieces} pieces); " if pieces == w["pieces"] else
                                        f"each step = {pieces} of {w['pieces']} pieces, evenly spread, time scaled by {w['pieces']}/{pieces}; ") +
                                       f"{cores} OpenMP threads, reference loop structure (parallel for dynamic + critical)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def batch_callback_bench(ev, w, cfg, dev, rank, world, B, barrier, max_over_ranks, reps=3):
    """B problems per GPU through isdf_callback_batch_device: x -> MINCO -> time-integral term with the discrete collision loop over the
    concatenated B*N pieces -> adjoint -> grad(x). Problems differ per rank (seeds), nothing is exchanged: weak scaling by construction."""
    import torch
    import workloads as W
    N0, X = w["pieces"], w["map_dim"]
    dim = 4 * N0 - 3
    xs, heads, tails = np.zeros((B, dim)), np.zeros((B, 9)), np.zeros((B, 9))
    for b in range(B):
        wp = W.random_walk_waypoints(N0, [0, 0, 0], [X, X, X], seed=1000 + rank * B + b)
        xs[b, :N0] = 1.0                                   # tau = 1 -> T = 2.5 s (inittime)
        xs[b, N0:] = wp[1:-1].reshape(-1)
        heads[b, 0:3], tails[b, 0:3] = wp[0], wp[-1]       # column-major 3x3: first column = position
    d_x, d_h, d_t = (torch.from_numpy(a).to(dev) for a in (xs, heads, tails))
    d_cost = torch.zeros(B, dtype=torch.float64, device=dev)
    d_grad = torch.zeros(B, dim, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ev.set_shard(0, 1)

    def step():
        ev.callback_batch_device(B, N0, d_h.data_ptr(), d_t.data_ptr(), 1, 20.0, d_x.data_ptr(), d_cost.data_ptr(), d_grad.data_ptr(), stream)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); barrier()
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        ts.append(max_over_ranks(e0.elapsed_time(e1)))
    ms = statistics.mean(ts)
    c = d_cost.cpu().numpy()
    per_problem = None
    if world == 1:
        # the same problems one at a time through the host adapter (host MINCO port + isdf_eval_discrete with host buffers)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import host_lib as H
        from ctypes import c_int as C_int
        L = H.lib()
        nb = min(B, 16)
        g = np.zeros(dim)
        t0 = time.perf_counter()
        cs = []
        for b in range(nb):
            be = L.isdf_host_backend_create(ev.h, N0, heads[b].ctypes.data_as(H.dp), tails[b].ctypes.data_as(H.dp), 20.0, 0, 1)
            cs.append(L.isdf_host_backend_cost(be, xs[b].ctypes.data_as(H.dp), g.ctypes.data_as(H.dp), dim))
            L.isdf_host_backend_destroy(be)
        dt = time.perf_counter() - t0
        per_problem = {"callbacks_per_s": nb / dt, "problems": nb, "max_rel_cost_diff_vs_batched": float(np.max(np.abs(np.array(cs) - c[:nb]) / np.abs(c[:nb])))}
    lockstep = None
    if world == 1:
        # B restarts under the lock-step batched L-BFGS driver (host/isdf_lbfgs.hpp): every round = one batched device callback
        try:
            nb = min(B, 32)
            Xl = np.ascontiguousarray(xs[:nb]).copy().reshape(-1)
            fl, rl = np.zeros(nb), np.zeros(nb, dtype=np.int32)
            itl, evl, stl = np.zeros(nb, dtype=np.int32), np.zeros(nb, dtype=np.int32), C_int()
            ipt = H.C.POINTER(H.C.c_int)
            t0 = time.perf_counter()
            rounds = L.isdf_host_lbfgs_batch_backend(ev.h, nb, N0, np.ascontiguousarray(heads[:nb]).ctypes.data_as(H.dp), np.ascontiguousarray(tails[:nb]).ctypes.data_as(H.dp),
                                                     20.0, Xl.ctypes.data_as(H.dp), fl.ctypes.data_as(H.dp), rl.ctypes.data_as(ipt), 16, 10, 1e-6, 0.0, 12,
                                                     itl.ctypes.data_as(ipt), evl.ctypes.data_as(ipt), H.C.byref(stl))
            dt = time.perf_counter() - t0
            lockstep = {"problems": nb, "rounds": int(rounds), "iterations_total": int(itl.sum()), "evaluations_total": int(evl.sum()), "seconds": dt,
                        "iters_per_s": float(itl.sum() / dt), "callback_evals_per_s": float(evl.sum() / dt), "status": int(stl.value),
                        "mean_cost_start": float(c[:nb].mean()), "mean_cost_end": float(fl.mean()), "max_iterations": 12}
        except Exception as e:
            lockstep = {"error": repr(e)}
    return {"callbacks_per_s": world * B * 1e3 / ms, "host_adapter_one_at_a_time": per_problem, "lockstep_lbfgs": lockstep, "ms_per_batch_max_over_ranks": ms, "problems_per_gpu": B, "problems_total": world * B,
            "scaling": "weak", "finite_costs": bool(np.all(np.isfinite(c))), "mean_cost": float(c.mean()),
            "what": f"BASELINE configs[4]: {world * B} random-restart problems ({N0} pieces x {w['samples_per_piece']} samples, shared {X}^3 map, mesh robot), "
                    "decision vector in -> cost and gradient out, MINCO forward/adjoint + time-integral/collision term all on the device; no collective"}


def frontend_bench(ev, w, occ, V, F, dev, with_cpu, reps=3):
    """SURVEY 8f row 4 on the bench map: which of the 121 (roll, pitch) attitudes of the mesh robot fit at EVERY voxel (128-bit mask per voxel)."""
    import torch
    X = w["map_dim"]
    t0 = time.perf_counter()
    xk, yk = ev.frontend_build_kernels(45.0, 45.0, 9.0, 0.0)          # config_CappedCone.yaml:62-64
    t_build = time.perf_counter() - t0
    nvox = X ** 3
    d_masks = torch.empty(nvox * 4, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ev.frontend_feasibility_device(d_masks.data_ptr(), stream)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ev.frontend_feasibility_device(d_masks.data_ptr(), stream); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    fit = int((d_masks.view(-1, 4) != 0).any(dim=1).sum().item())
    peak, peak_src = measured_peak_hbm()
    nbytes = nvox * 16 + nvox // 8
    out = {"what": f"attitude-kernel feasibility of all {X}^3 voxels x {xk * yk} attitudes (kernel {w['kernel_size']}^3, mesh robot): kernelConv<true> of "
                   "sw_manager.hpp:821-846 for every voxel, 128-bit mask out",
           "ms": ms, "voxels_per_s": nvox / (ms * 1e-3), "voxel_attitude_checks_per_s": nvox * xk * yk / (ms * 1e-3), "kernel_build_s": t_build,
           "voxels_with_a_fitting_attitude": fit,
           "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                        "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                        "note": "16 B mask written + 1 bit occupancy read per voxel; the window reads hit L1/L2"}}
    del d_masks
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        sub = np.ascontiguousarray(occ[:96, :96, :96])
        fe = O.FrontEnd(O.Shape.mesh(V, F, w["poly_params"], wn_mode=O.WN_BH), sub, ks=w["kernel_size"])
        rng = np.random.default_rng(0)
        n = 200000
        ind = rng.integers(0, 96, (n, 3))
        O.lib().orc_set_num_threads(usable_cores())                   # the front-end loop uses OpenMP's default team
        fe.feasibility(ind[:2000])                                    # wake the OpenMP pool
        dts = []
        for _ in range(2):
            t0 = time.perf_counter(); fe.feasibility(ind); dts.append(time.perf_counter() - t0)
        dt = min(dts)
        out["cpu"] = {"voxels_per_s": n / dt, "cores": usable_cores(), "sample": f"{n} random voxels of a 96^3 corner of the same map, oracle byte-kernel port, OpenMP, best of 2"}
        out["speedup_vs_cpu"] = out["voxels_per_s"] / out["cpu"]["voxels_per_s"]
    return out


def workload_name(w):
    return (f"BASELINE configs[2]: random {w['map_dim']}^3 voxel map (p={w['occupancy']}, wall slabs), {w['pieces']}-piece MINCO traj, "
            f"{w['samples_per_piece']} samples/piece, mesh-SDF robot ({w['mesh']}), discrete collision cost+grad")


# ---- GPU arm ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import isdf_b200 as I
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, cfg, occ, T, Cc, V, F = make_workload(args.small)
    N = w["pieces"]
    ev = I.Evaluator(cfg, device=local)
    ev.set_map_u8(occ, [0, 0, 0], 1.0)
    if args.robot == "mesh":
        ev.set_shape_mesh(V, F, w["poly_params"])
    else:
        import workloads as W
        R, t = W.rotation_from_poly_params(w["poly_params"])
        ev.set_shape_named(args.robot, R, t)
        w["mesh"] = "analytic:" + args.robot
        args.no_cpu_baseline = True
    ev.set_shard(rank, world)
    dev = torch.device("cuda", local)
    d_T = torch.from_numpy(T).to(dev)
    d_C = torch.from_numpy(Cc).to(dev)
    d_out = torch.zeros(19 * N + 1, dtype=torch.float64, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream
    ref_nccl = None
    if world > 1:                                         # untimed cross-check for the fused exchange: same shards, NCCL sum
        ev.eval_discrete_device(N, d_T.data_ptr(), d_C.data_ptr(), d_out.data_ptr(), stream)
        allreduce_partials(d_out)
        torch.cuda.synchronize()
        ref_nccl = d_out.cpu().numpy().copy()
    # multi-GPU reduction: fused into the evaluation through NVLink peer memory (isdf_peer_*); NCCL all-reduce if that cannot be set up
    collective = "none"
    if world > 1:
        import torch.distributed as dist
        collective = "nccl"
        if args.collective == "peer":
            try:
                handles = [None] * world
                dist.all_gather_object(handles, ev.peer_export(world, 19 * N + 1))
                ev.peer_connect(world, rank, handles, fuse=True)
                okf = torch.ones(1, device=dev)
            except Exception as e:
                okf = torch.zeros(1, device=dev)
                if rank == 0:
                    print(f"bench: peer-memory exchange unavailable ({e}); using NCCL", file=sys.stderr)
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if okf.item() == 1:
                collective = "peer"
            else:
                try:
                    ev.peer_disconnect()
                except Exception:
                    pass
            dist.barrier()

    def step_device():
        ev.eval_discrete_device(N, d_T.data_ptr(), d_C.data_ptr(), d_out.data_ptr(), stream)   # collective == "peer": includes the exchange
        if collective == "nccl":
            allreduce_partials(d_out)

    sampler = ClockSampler(local) if rank == 0 else None   # nvidia-smi needs ~0.5 s to start: begin before the warm-up
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    barrier()
    l0 = ev.stats().kernel_launches
    ev_pairs = []
    times = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        barrier()
        e0.record()
        step_device()
        e1.record()
        torch.cuda.synchronize()
        times.append(max_over_ranks(e0.elapsed_time(e1)))
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = ev.stats().kernel_launches - l0
    # warm-L2 figure for context (steady state of an optimiser loop: map stays L2 resident)
    warm = []
    for _ in range(min(args.steps, 10)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); barrier()
        e0.record(); step_device(); e1.record()
        torch.cuda.synchronize()
        warm.append(max_over_ranks(e0.elapsed_time(e1)))
    # the timed region is only tens of milliseconds: keep the same kernel running (untimed) until nvidia-smi has sampled it
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 1.5:
        for _ in range(50):
            step_device()
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = statistics.mean(times)
    result = d_out.cpu().numpy().copy()

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (H2D + kernel + D2H inside the timed region) ---------
    e2e_times = []
    h_part = torch.empty(19 * N + 1, dtype=torch.float64).pin_memory()
    for it in range(3 + args.steps):
        flush.zero_()
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        c, gC, gT = ev.eval_discrete(T, Cc)               # isdf_eval_discrete: pinned staging, H2D, kernel, D2H, sync
        if collective == "nccl":
            h_part[0] = c; h_part[1:1 + 18 * N] = torch.from_numpy(gC); h_part[1 + 18 * N:] = torch.from_numpy(gT)
            d_tmp = h_part.to(dev, non_blocking=True)
            allreduce_partials(d_tmp)
            h_part.copy_(d_tmp)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if it >= 3:
            e2e_times.append(max_over_ranks(dt))
    pairs = ev.stats().last_pairs
    kernel_ms_alone = ev.stats().last_kernel_ms
    ms_e2e = statistics.mean(e2e_times)

    # ---- context for N > 1: batch of random restarts, one trajectory per GPU, no collective (BASELINE configs[4] in miniature) ----
    batch_weak = None
    if world > 1:
        import workloads as W
        Tb, Cb, _ = W.make_trajectory(N, [0, 0, 0], [w["map_dim"]] * 3, seed=11 + rank, jitter=0.2)
        ev.set_shard(0, 1)
        d_Tb, d_Cb = torch.from_numpy(Tb).to(dev), torch.from_numpy(Cb).to(dev)
        for _ in range(4):
            ev.eval_discrete_device(N, d_Tb.data_ptr(), d_Cb.data_ptr(), d_out.data_ptr(), stream)
        tb = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); barrier()
            e0.record(); ev.eval_discrete_device(N, d_Tb.data_ptr(), d_Cb.data_ptr(), d_out.data_ptr(), stream); e1.record()
            torch.cuda.synchronize()
            tb.append(max_over_ranks(e0.elapsed_time(e1)))
        ev.set_shard(rank, world)
        batch_weak = {"evals_per_s": world * 1e3 / statistics.mean(tb), "ms_per_step_max_over_ranks": statistics.mean(tb), "scaling": "weak",
                      "what": f"{world} different trajectories (seed 11+rank), one per GPU, evaluated concurrently; no collective"}

    # ---- BASELINE configs[4]: batch of random restarts, whole callback on the device, B_local problems per GPU, no collective ----
    batch_cb = None
    if not args.no_batch and not args.small:
        try:
            batch_cb = batch_callback_bench(ev, w, cfg, dev, rank, world, args.batch_per_gpu, barrier, max_over_ranks)
        except Exception as e:
            batch_cb = {"error": repr(e)}
        ev.set_shard(rank, world)

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        ab = algorithmic_bytes(w)
        achieved = ab / (ms * 1e-3) / 1e9 if world == 1 else ab / world / (ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(w), **w, "l2": "flushed (512 MiB write) between timed steps",
                           "parallelism": (f"sample-interleaved shards x{world} + " + (f"rank-ordered sum of {19 * N + 1} doubles over NVLink peer memory, fused into the epilogue kernel"
                                                                                      if collective == "peer" else f"1 NCCL all-reduce of {19 * N + 1} doubles")) if world > 1 else "1 GPU"},
                "clocks": clocks,
                "e2e": {"value": 1e3 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * 19 * N, "d2h_bytes_per_step": 8 * (19 * N + 1) + 8,
                        "ms_per_step": ms_e2e, "api": "isdf_eval_discrete (host buffers)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                             "peak_source": peak_src, "kernel": "k_discrete_mesh", "algorithmic_bytes_per_launch": ab // world,
                             "fp64_pipe_active_pct_ncu": ncu_capture("fp64_pipe_active_pct"), "issue_slots_busy_pct_ncu": ncu_capture("issue_slots_busy_pct"),
                             "note": "window bytes counted at 1 B/voxel per sample (SURVEY §8d); the kernel is FP64/latency bound, see DESIGN.md"},
                "extra": {"pairs_per_eval": int(pairs), "pairs_per_s": pairs / (ms * 1e-3), "ms_per_step_warm_l2": statistics.mean(warm),
                          "evals_per_s_warm_l2": 1e3 / statistics.mean(warm), "ms_min": min(times), "ms_max": max(times),
                          "kernel_ms_in_host_call": kernel_ms_alone, "wall_s_timed_loop": wall,
                          "cost": float(result[0]), "grad_norm": float(np.linalg.norm(result[1:])), "batch_weak": batch_weak,
                          "collective": collective,
                          "collective_vs_nccl_rel_l2": (float(np.linalg.norm(result - ref_nccl) / np.linalg.norm(ref_nccl)) if ref_nccl is not None else None),
                          "batch_callback": batch_cb}}
        if world == 1 and not args.no_lbfgs:
            try:
                line["extra"]["lbfgs"] = lbfgs_ours(local)
                if not args.no_cpu_baseline:
                    line["extra"]["lbfgs"]["cpu"] = lbfgs_cpu()
            except Exception as e:   # secondary metric must never take the headline line down
                line["extra"]["lbfgs"] = {"error": repr(e)}
        if world == 1 and not args.no_frontend and args.robot == "mesh" and not args.small:
            try:
                line["extra"]["frontend"] = frontend_bench(ev, w, occ, V, F, dev, not args.no_cpu_baseline)
            except Exception as e:
                line["extra"]["frontend"] = {"error": repr(e)}
        if world == 1 and not args.no_swept:
            try:
                line["extra"]["swept"] = swept_ours(local, not args.no_cpu_baseline)
            except Exception as e:
                line["extra"]["swept"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(w, cfg, occ, T, Cc, V, F, budget_s=args.cpu_budget)
            line["extra"]["speedup_kernel_vs_cpu"] = line["value"] / line["cpu_baseline"]["value"]
            line["extra"]["speedup_e2e_vs_cpu"] = line["e2e"]["value"] / line["cpu_baseline"]["value"]
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        if collective == "peer":
            ev.peer_status()
            ev.peer_disconnect()
    ev.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="tiny workload for plumbing checks (not a bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lbfgs", action="store_true", help="skip the secondary L-BFGS iterations/s measurement")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"], help="N > 1: reduction of the sharded evaluation")
    ap.add_argument("--no-batch", action="store_true", help="skip the batched device-callback measurement (configs[4])")
    ap.add_argument("--batch-per-gpu", type=int, default=128, help="problems per GPU in the batched callback measurement (1024 / 8 GPUs)")
    ap.add_argument("--no-frontend", action="store_true", help="skip the front-end attitude-kernel feasibility measurement")
    ap.add_argument("--no-swept", action="store_true", help="skip the secondary swept-volume (SV-SDF) measurement")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--ref-pieces", type=int, default=64, help="--impl reference: pieces per step sample (of 64; evenly spread; 64 = the whole workload, ~1-2 s per step)")
    ap.add_argument("--robot", default="mesh", help="mesh (headline) or an analytic shape name, e.g. SmoothIntersection (diagnostic runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

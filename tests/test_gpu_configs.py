"""GPU parity at the size of EVERY BASELINE.json config, on the code path the bench times.

Each case compares, against the oracle on identical inputs (tolerance 1e-6 relative cost / rel-L2 gradient, the north star's bar):
  * the FIRST evaluation of a context (natural sample order, nothing split), and
  * the THIRD evaluation (longest-first work items learned from the previous evaluation, heavy samples split into row-class parts) —
    the steady state `bench.py` times — which must also be BIT-identical to the first;
  * one case forces splitting on a single GPU (isdf_dbg_schedule with a huge pretended warp-slot count) so that 2..32-way split samples are checked
    against the oracle directly.
configs[0] ball robot, 8 pieces x 32 samples, 64^3 map            configs[1] three-slit map, the reference's RoundedCone.obj, 32 x 128
configs[2] random 512^3 map, 64 x 256, mesh robot — ALL 64 pieces  configs[3] swept-volume path, 256^3 map, 846 obstacle points, mesh robot
configs[4] 1024 random-restart problems on the 512^3 map through the batched device callback; 8 sampled problems against the oracle
The oracle runs in its OpenMP mode (same arithmetic per sample; the summation order over samples differs by ~1e-16)."""
import os
import numpy as np
import pytest
import isdf_b200 as I
import oracle_lib as O
import workloads as W
from common import rel_l2, BMIN

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-6
NTHREADS = max(1, min(len(os.sched_getaffinity(0)), 16))


def grads(r):
    return np.concatenate([r[1], r[2]])


def ocfg_of(cfg):
    oc = O.config_from(cfg)
    oc.threads_num = NTHREADS
    return oc


def check(got, exp, what):
    relc = abs(got[0] - exp[0]) / max(abs(exp[0]), 1e-300)
    relg = rel_l2(grads(got), np.concatenate([exp[1], exp[2]]))
    assert relc <= TOL and relg <= TOL, f"{what}: cost rel {relc:.3e}, grad rel-L2 {relg:.3e}"
    return relc, relg


def first_and_steady(ev, T, Cc, exp, what):
    """1st evaluation (natural order) and 3rd (work items + splits) against the oracle; the two must agree bit for bit"""
    a = ev.eval_discrete(T, Cc)
    check(a, exp, what + " [first evaluation]")
    ev.eval_discrete(T, Cc)
    c = ev.eval_discrete(T, Cc)
    check(c, exp, what + " [third evaluation: work items]")
    assert a[0] == c[0] and np.array_equal(grads(a), grads(c)), what + ": the work-item schedule changed a bit of the result"
    assert ev.stats().last_pairs == exp[3], what + ": (pose, voxel) pair count differs from the oracle's"
    return a


def test_config0_ball_8x32_64cube():
    cfg = I.default_config_values()
    cfg.flags = I.WITH_DYNAMICS | I.WITH_COLLISION
    cfg.integral_intervs = 32
    occ = W.three_slit_map(64, 64, 64, noise=0.05, seed=1)
    T, Cc, _ = W.make_trajectory(8, [0, 0, 0], [50, 50, 34], seed=4, jitter=0.2)
    exp = O.eval_discrete(O.config_from(cfg), occ, BMIN, 1.0, O.Shape.named("Ball"), T, Cc)      # serial oracle
    assert exp[0] > 0 and exp[3] > 1000
    ev = I.Evaluator(cfg)
    ev.set_map_u8(occ, BMIN, 1.0)
    ev.set_shape_named("Ball")
    first_and_steady(ev, T, Cc, exp, "configs[0]")
    ev.close()


def test_config1_three_slit_reference_mesh_32x128():
    z = np.load(os.path.join(G, "ref_meshes.npz"))
    V, F, pp = z["RoundedCone_V"], z["RoundedCone_F"], z["RoundedCone_pp"]                        # config_CappedCone.yaml:8-10
    cfg = I.default_config_values()
    cfg.flags = I.WITH_DYNAMICS | I.WITH_COLLISION
    cfg.integral_intervs = 128
    occ = W.three_slit_map(64, 64, 64, noise=0.02, seed=2)
    T, Cc, _ = W.make_trajectory(32, [0, 0, 0], [50, 50, 34], seed=6, jitter=0.2)
    exp = O.eval_discrete(ocfg_of(cfg), occ, BMIN, 1.0, O.Shape.mesh(V, F, pp, wn_mode=O.WN_BH), T, Cc, use_omp=True)
    assert exp[0] > 0 and exp[3] > 100000
    ev = I.Evaluator(cfg)
    ev.set_map_u8(occ, BMIN, 1.0)
    ev.set_shape_mesh(V, F, pp)
    first_and_steady(ev, T, Cc, exp, "configs[1]")
    # forced splitting: pretend the device has 5e7 resident warps -> the balanced share per warp is tiny and every sample heavier than
    # SPLIT_WORK_MIN is split 2..32 ways
    ev.dbg_schedule(warp_slots=50000000)
    ev.eval_discrete(T, Cc)
    s = ev.eval_discrete(T, Cc)
    nitems, nparts = ev.dbg_item_stats()
    assert nparts > 200, f"expected many split parts, got {nparts} of {nitems} items"
    check(s, exp, f"configs[1] forced splits ({nparts} parts)")
    ev.dbg_schedule(natural_order=True)
    n = ev.eval_discrete(T, Cc)
    assert n[0] == s[0] and np.array_equal(grads(n), grads(s))
    if O.ref_fwn_available():   # deviation of the ±1 sign policy from the reference-faithful s = 1 - 2 w_FWN at this config's size (reported, not a parity claim)
        rf = O.eval_discrete(ocfg_of(cfg), occ, BMIN, 1.0, O.Shape.mesh(V, F, pp, wn_mode=O.WN_REF), T, Cc, use_omp=True)
        print(f"configs[1] vs reference-faithful FWN sign: |dcost|/cost = {abs(s[0] - rf[0]) / rf[0]:.3e}, grad rel-L2 = {rel_l2(grads(s), np.concatenate([rf[1], rf[2]])):.3e}")
    ev.close()


@pytest.fixture(scope="module")
def big():
    X = 512
    occ = W.random_map(X, X, X, p=0.05, seed=1, slabs=3)
    T, Cc, wp = W.make_trajectory(64, [0, 0, 0], [X, X, X], seed=11, jitter=0.2)
    V, F = W.rounded_cone_mesh()
    return occ, T, Cc, V, F, [0.0, 0.0, 0.0, 120.0, 0.0, 0.0]


def test_config2_random_512cube_64x256_mesh_all_pieces(big):
    """the bench workload itself, every one of its 64 pieces (16448 pose samples)"""
    occ, T, Cc, V, F, pp = big
    cfg = I.default_config_values()
    cfg.flags = I.WITH_DYNAMICS | I.WITH_COLLISION
    cfg.integral_intervs = 256
    exp = O.eval_discrete(ocfg_of(cfg), occ, BMIN, 1.0, O.Shape.mesh(V, F, pp, wn_mode=O.WN_BH), T, Cc, use_omp=True)
    assert exp[3] > 1000000
    ev = I.Evaluator(cfg)
    ev.set_map_u8(occ, BMIN, 1.0)
    ev.set_shape_mesh(V, F, pp)
    first_and_steady(ev, T, Cc, exp, "configs[2]")
    nitems, nparts = ev.dbg_item_stats()
    assert nparts > 0                                                # the steady state of this workload does split its heaviest samples
    # a perturbed trajectory evaluated with the schedule learned from the unperturbed one (what an optimiser step does)
    rng = np.random.default_rng(3)
    Cp = Cc + 1e-3 * rng.normal(size=Cc.size)
    expp = O.eval_discrete(ocfg_of(cfg), occ, BMIN, 1.0, O.Shape.mesh(V, F, pp, wn_mode=O.WN_BH), T, Cp, use_omp=True)
    check(ev.eval_discrete(T, Cp), expp, "configs[2] perturbed iterate on a stale schedule")
    ev.close()


def test_config3_swept_256cube_846_points_mesh():
    X = 256
    occ = W.random_map(X, X, X, p=0.02, seed=2, slabs=3)
    cfg = I.default_config_values()
    cfg.flags = I.WITH_DYNAMICS
    T, Cc, wp = W.make_trajectory(64, [0, 0, 0], [X, X, X], seed=11, jitter=0.2)
    pts = W.gather_obstacle_points(occ, BMIN, 1.0, wp, cfg.kernel_size / 3.0)
    assert len(pts) == 846
    V, F = W.rounded_cone_mesh()
    pp = [0, 0, 0, 120, 0, 0]
    ref = O.eval_swept(ocfg_of(cfg), O.Shape.mesh(V, F, pp, wn_mode=O.WN_BH), T, Cc, pts, use_omp=True)
    assert ref["cost"] > 0
    ev = I.Evaluator(cfg)
    ev.set_shape_mesh(V, F, pp)
    ev.set_points(pts)
    for rep in range(2):
        c, gC, gT = ev.eval_swept(T, Cc)
        relc = abs(c - ref["cost"]) / ref["cost"]
        relg = rel_l2(np.concatenate([gC, gT]), np.concatenate([ref["gradC"], ref["gradT"]]))
        assert relc <= TOL and relg <= TOL, f"configs[3] evaluation {rep}: cost rel {relc:.3e}, grad rel-L2 {relg:.3e}"
    ts, sd, _ = ev.swept_results()
    hit = ref["sdf"] < 9.0                                           # points that came into range at all (10.0 = never, swm:733)
    assert np.array_equal(sd < 9.0, hit)
    assert np.abs(ts[hit] - ref["tstar"][hit]).max() <= 8e-5 and np.abs(sd[hit] - ref["sdf"][hit]).max() <= 1e-9
    assert ev.stats().last_sdf_evals == ref["nsdf"]                  # reference-equivalent SDF evaluation count
    ev.close()


def test_config4_1024_restarts_on_512cube_sampled_against_oracle(big):
    occ, _, _, V, F, pp = big
    X, N0, B, rho = 512, 64, 1024, 20.0
    cfg = I.default_config_values()
    cfg.flags = I.WITH_DYNAMICS | I.WITH_COLLISION
    cfg.integral_intervs = 256
    dim = 4 * N0 - 3
    xs, heads, tails = np.zeros((B, dim)), np.zeros((B, 3, 3)), np.zeros((B, 3, 3))
    for b in range(B):
        wp = W.random_walk_waypoints(N0, [0, 0, 0], [X, X, X], seed=1000 + b)
        xs[b, :N0] = 1.0
        xs[b, N0:] = wp[1:-1].reshape(-1)
        heads[b, :, 0], tails[b, :, 0] = wp[0], wp[-1]
    ev = I.Evaluator(cfg)
    ev.set_map_u8(occ, BMIN, 1.0)
    ev.set_shape_mesh(V, F, pp)
    cost, grad = ev.callback_batch(heads, tails, rho, xs)
    cost2, grad2 = ev.callback_batch(heads, tails, rho, xs)          # second call: work items over the 16.8 M concatenated samples
    assert np.array_equal(cost, cost2) and np.array_equal(grad, grad2)
    assert np.all(np.isfinite(cost)) and np.all(np.isfinite(grad))
    ev.close()
    oc = ocfg_of(cfg)
    osh = O.Shape.mesh(V, F, pp, wn_mode=O.WN_BH)
    for b in [0, 1, 127, 128, 500, 511, 777, 1023]:
        tau = xs[b, :N0]
        Tt = np.where(tau > 0, (0.5 * tau + 1) * tau + 1, 1 / ((0.5 * tau - 1) * tau + 1))
        inP = xs[b, N0:].reshape(-1, 3).T
        co, energy, gC, gT = O.minco_forward(heads[b], tails[b], inP, Tt)
        di = O.eval_discrete(oc, occ, BMIN, 1.0, osh, Tt, co, use_omp=True)
        c_ref = energy + di[0] + rho * Tt.sum()
        gp, gt = O.minco_backward(heads[b], tails[b], inP, Tt, gC + di[1], gT + di[2])
        gt = gt + rho
        gtau = np.where(tau > 0, gt * (tau + 1), gt * (1 - tau) / ((0.5 * tau - 1) * tau + 1) ** 2)
        g_ref = np.concatenate([gtau, gp.T.reshape(-1)])
        relc, relg = abs(cost[b] - c_ref) / abs(c_ref), rel_l2(grad[b], g_ref)
        assert relc <= TOL and relg <= TOL, f"configs[4] problem {b}: cost rel {relc:.3e}, grad rel-L2 {relg:.3e}"

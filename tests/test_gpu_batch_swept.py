"""Batched LIVE callback: isdf_callback_batch with per-problem obstacle point sets (isdf_set_points_batch) = swept-volume term + time-integral
term for B problems at once (back_end_optimizer.hpp:386-405) == the per-problem host adapter (host MINCO + isdf_eval_swept + isdf_eval_discrete)."""
import numpy as np
import pytest
import isdf_b200 as I
import oracle_lib as O
import workloads as W
from common import small_case, rel_l2, BMIN, MESHES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("robot", ["Torus", "mesh"])
def test_batched_callback_with_swept_term_matches_host_adapter(robot):
    import host_lib as H
    cfg, occ, _, _, _ = small_case(N=5, K=16, seed=3, flags=I.WITH_DYNAMICS)        # the live composition: collision through the swept term only
    cfg.vmax, cfg.omgmax = 1.2, 0.5
    B, N0, rho = 6, 5, 20.0
    rng = np.random.default_rng(4)
    heads, tails, X, psets = [], [], [], []
    for b in range(B):
        wp = W.random_walk_waypoints(N0, [0, 0, 0], [50, 50, 34], seed=300 + b)
        h, t = np.zeros((3, 3)), np.zeros((3, 3))
        h[:, 0], t[:, 0] = wp[0], wp[-1]
        tau = rng.normal(size=N0) * 0.3 + 0.9
        X.append(np.concatenate([tau, wp[1:-1].reshape(-1)]))
        heads.append(h); tails.append(t)
        pts = W.gather_obstacle_points(occ, BMIN, 1.0, wp, cfg.kernel_size * cfg.occupancy_resolution / 3.0)
        psets.append(pts[:: max(1, len(pts) // (40 + 25 * b))][: 40 + 25 * b] if b != 2 else pts[:0])    # different sizes, one EMPTY set
    X = np.array(X)
    ev = I.Evaluator(cfg)
    ev.set_map_u8(occ, BMIN, 1.0)
    if robot == "mesh":
        V, F = MESHES["rcone"]()
        ev.set_shape_mesh(V, F, [0, 0, 0, 120, 0, 0])
    else:
        ev.set_shape_named(robot)
    L = H.lib()
    ref_c, ref_g = [], []
    for b in range(B):
        g = np.zeros(X.shape[1])
        hh, tt = np.asfortranarray(heads[b]), np.asfortranarray(tails[b])
        if len(psets[b]):
            ev.set_points(psets[b])
            be = L.isdf_host_backend_create(ev.h, N0, hh.ctypes.data_as(H.dp), tt.ctypes.data_as(H.dp), rho, 1, 1)
        else:
            be = L.isdf_host_backend_create(ev.h, N0, hh.ctypes.data_as(H.dp), tt.ctypes.data_as(H.dp), rho, 0, 1)
        ref_c.append(L.isdf_host_backend_cost(be, X[b].ctypes.data_as(H.dp), g.ctypes.data_as(H.dp), g.size))
        ref_g.append(g)
        L.isdf_host_backend_destroy(be)
    ev.set_points_batch(psets)
    cost, grad = ev.callback_batch(np.array(heads), np.array(tails), rho, X)
    c_no, _ = None, None
    for b in range(B):
        assert abs(cost[b] - ref_c[b]) <= 1e-12 * abs(ref_c[b]), (b, cost[b], ref_c[b])
        assert rel_l2(grad[b], ref_g[b]) <= 1e-11, (b, rel_l2(grad[b], ref_g[b]))
    # the swept term is really in there: switching it off changes the problems that have points in range
    ev.set_points_batch([])
    cost0, _ = ev.callback_batch(np.array(heads), np.array(tails), rho, X)
    assert cost0[2] == cost[2] and np.any(cost0 != cost)
    # second evaluation with the sets registered again: t* persistence does not change the result (set_ts = false, swm:576)
    ev.set_points_batch(psets)
    c1, g1 = ev.callback_batch(np.array(heads), np.array(tails), rho, X)
    c2, g2 = ev.callback_batch(np.array(heads), np.array(tails), rho, X)
    assert np.array_equal(c1, cost) and np.array_equal(c2, cost) and np.array_equal(g2, grad)
    # device-resident L-BFGS over the live callback runs and decreases every cost
    r = ev.lbfgs_batch(np.array(heads), np.array(tails), rho, X, ev.lbfgs_params(max_iterations=5))
    assert np.all(r["f"] <= cost) and np.all(np.isfinite(r["f"]))
    ev.close()

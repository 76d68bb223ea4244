"""Host-side C++ adapters (no GPU needed): the product's MINCO port against the oracle's and a dense numpy solve, and the
tau <-> T decision-variable maps (back_end_optimizer.hpp:214-256)."""
import numpy as np
import host_lib as H
import oracle_lib as O
import workloads as W
from common import rel_l2


def _problem(N=6, seed=0):
    rng = np.random.default_rng(seed)
    wp = W.random_walk_waypoints(N, [0, 0, 0], [40, 40, 30], seed=seed + 1)
    T = 2.5 * (1 + 0.4 * (rng.random(N) - 0.5))
    head, tail = np.zeros((3, 3)), np.zeros((3, 3))
    head[:, 0], tail[:, 0] = wp[0], wp[-1]
    head[:, 1], head[:, 2], tail[:, 1] = rng.normal(size=3) * 0.3, rng.normal(size=3) * 0.1, rng.normal(size=3) * 0.2
    return wp, T, head, tail, rng


def test_host_minco_equals_oracle_and_dense_solve():
    for N in (1, 2, 6, 17):
        wp, T, head, tail, rng = _problem(N, seed=N)
        inP = wp[1:-1].T if N > 1 else np.zeros((3, 0))
        co, e, gc, gt = H.minco_forward(head, tail, inP, T)
        ref = W.minco_s3(wp, T, head_va=[head[:, 1], head[:, 2]], tail_va=[tail[:, 1], tail[:, 2]])
        assert rel_l2(co, ref) < 1e-9
        oco, oe, ogc, ogt = O.minco_forward(head, tail, inP, T)
        assert rel_l2(co, oco) < 1e-12 and abs(e - oe) <= 1e-12 * abs(oe) and rel_l2(gc, ogc) < 1e-12 and rel_l2(gt, ogt) < 1e-12
        if N > 1:
            gC, gT = rng.normal(size=18 * N), rng.normal(size=N)
            gp, gtt = H.minco_backward(head, tail, inP, T, gC, gT)
            ogp, ogtt = O.minco_backward(head, tail, inP, T, gC, gT)
            assert rel_l2(gp, ogp) < 1e-11 and rel_l2(gtt, ogtt) < 1e-11


def test_tau_maps_round_trip():
    tau = np.array([-3.0, -0.5, 0.0, 0.2, 1.7, 4.0])
    T, back = H.tau_maps(tau)
    exp = np.where(tau > 0, (0.5 * tau + 1) * tau + 1, 1 / ((0.5 * tau - 1) * tau + 1))
    assert np.allclose(T, exp, rtol=1e-15) and np.allclose(back, tau, atol=1e-12) and np.all(T > 0)


def test_lbfgs_driver_known_answers():
    """Rosenbrock (minimum 0 at (1,..,1)) and a convex quadratic: the driver with the reference fork's semantics converges."""
    def rosen(x):
        f = np.sum(100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2)
        g = np.zeros_like(x)
        g[:-1] = -400 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
        g[1:] += 200 * (x[1:] - x[:-1] ** 2)
        return f, g
    r = H.lbfgs_minimize(rosen, np.array([-1.2, 1.0, -0.5, 0.8]), past=0, g_epsilon=1e-8, max_iterations=5000)
    assert r["ret"] in (0, 1) and np.allclose(r["x"], 1.0, atol=1e-4) and r["f"] < 1e-8
    assert r["evaluations"] >= r["iterations"] >= 5
    A = np.diag([1.0, 10.0, 100.0]); b = np.array([1.0, -2.0, 3.0])
    r = H.lbfgs_minimize(lambda x: (0.5 * x @ A @ x - b @ x, A @ x - b), np.zeros(3), past=0, g_epsilon=1e-10, max_iterations=500)
    assert np.allclose(r["x"], np.linalg.solve(A, b), atol=1e-6)
    # argument validation mirrors the reference's return codes
    assert H.lbfgs_minimize(rosen, np.zeros(2), mem_size=0)["ret"] == -1022    # LBFGSERR_INVALID_MEMSIZE
    # NaN cost aborts with LBFGSERR_INVALID_FUNCVAL: what the C ABI's NaN-on-error relies on
    assert H.lbfgs_minimize(lambda x: (float("nan"), x), np.ones(2))["ret"] == -1012


def test_batched_lbfgs_driver_equals_sequential_instances():
    """lbfgs_optimize_batch: B instances in lock step, each bit-identical (iterates, value, iteration / evaluation counts, return
    code) to the sequential driver run alone — including instances that finish early, hit the direction-reset re-evaluation,
    fail in the line search, or start at a stationary point."""
    def rosen(x):
        f = np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2)
        g = np.zeros_like(x)
        g[:-1] = -400 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
        g[1:] += 200 * (x[1:] - x[:-1] ** 2)
        return f, g
    A = np.diag([1.0, 10.0, 100.0, 3.0])

    def quad(x):
        return 0.5 * x @ A @ x, A @ x

    def hinge(x):       # smoothed-L1 like kink: piecewise cubic / linear, exercises long line searches
        f, g = 0.0, np.zeros_like(x)
        for i, v in enumerate(x):
            a = abs(v)
            if a > 0.1:
                f += a - 0.05; g[i] = np.sign(v)
            else:
                f += (0.1 - a / 2) * (a / 0.1) ** 3
                g[i] = np.sign(v) * ((a / 0.1) ** 2 * (-a / 0.2 + 3 * (0.1 - a / 2) / 0.1))
        return f, g
    funs = [rosen, quad, hinge, rosen, quad, lambda x: (float("nan"), x), quad]
    rng = np.random.default_rng(4)
    X0 = rng.normal(size=(len(funs), 4))
    X0[6] = 0.0                                                      # stationary start: converges without a line search
    X0[3] = [-1.2, 1.0, -0.5, 0.8]
    kw = dict(mem_size=6, past=3, delta=1e-9, g_epsilon=1e-7, max_iterations=200)
    rb = H.lbfgs_minimize_batch(lambda i, x: funs[i](x), X0, **kw)
    assert rb["rounds"] >= max(rb["evaluations"])
    kinds = set()
    for b, fn in enumerate(funs):
        rs = H.lbfgs_minimize(fn, X0[b], **kw)
        assert rs["ret"] == rb["ret"][b], (b, rs["ret"], rb["ret"][b])
        assert rs["iterations"] == rb["iterations"][b] and rs["evaluations"] == rb["evaluations"][b], b
        assert np.array_equal(rs["x"], rb["x"][b]), b
        if np.isfinite(rs["f"]):
            assert rs["f"] == rb["f"][b], b
        kinds.add(int(rs["ret"]))
    assert len(kinds) >= 3 and min(rb["iterations"]) == 0 and max(rb["iterations"]) > 10

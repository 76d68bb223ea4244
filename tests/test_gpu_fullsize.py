"""Full-size (BASELINE.json configs[2]: 512^3 map, 64 pieces, 256 samples/piece) checks through size-independent
properties — the oracle would need minutes here, so parity at this size is pinned by invariants:
  * linearity: doubling weight_p doubles cost and gradient bit-exactly (power-of-two scaling commutes with rounding)
  * shard additivity: partial sums over 8 interleaved shards reproduce the unsharded result
  * run-to-run bit determinism
  * the gradient is the directional derivative of the cost (central differences along random directions)
  * an empty map gives exactly zero
plus a sampled oracle comparison on a sub-trajectory of the same workload."""
import numpy as np
import pytest
import isdf_b200 as I
import oracle_lib as O
import workloads as W
from common import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    X = 512
    occ = W.random_map(X, X, X, p=0.05, seed=1, slabs=3)
    cfg = I.default_config_values()
    cfg.integral_intervs = 256
    cfg.flags = I.WITH_COLLISION
    T, Cc, wp = W.make_trajectory(64, [0, 0, 0], [X, X, X], seed=11, jitter=0.2)
    return cfg, occ, T, Cc


def grads(r):
    return np.concatenate([r[1], r[2]])


def test_full_size_invariants(big):
    cfg, occ, T, Cc = big
    ev = I.Evaluator(cfg)
    ev.set_map_u8(occ, [0, 0, 0], 1.0)
    ev.set_shape_named("RoundedCone")
    a = ev.eval_discrete(T, Cc)
    b = ev.eval_discrete(T, Cc)
    assert a[0] > 0 and a[0] == b[0] and np.array_equal(grads(a), grads(b))
    pairs = ev.stats().last_pairs
    assert pairs > 100000
    # shard additivity
    acc = [0.0, 0.0]
    for r in range(8):
        ev.set_shard(r, 8)
        p = ev.eval_discrete(T, Cc)
        acc[0] += p[0]; acc[1] = acc[1] + grads(p)
    ev.set_shard(0, 1)
    assert abs(acc[0] - a[0]) <= 1e-12 * a[0] and rel_l2(acc[1], grads(a)) < 1e-12
    # directional derivative
    rng = np.random.default_rng(0)
    e = 1e-6
    okc = 0
    for _ in range(4):
        dC, dT = rng.normal(size=Cc.size), rng.normal(size=T.size) * 0.1
        cp = ev.eval_discrete(T + e * dT, Cc + e * dC)[0]
        cm = ev.eval_discrete(T - e * dT, Cc - e * dC)[0]
        num, ana = (cp - cm) / (2 * e), a[1] @ dC + a[2] @ dT
        okc += abs(num - ana) <= 2e-3 * abs(ana)
    assert okc >= 3
    ev.close()
    # linearity in weight_p
    cfg2 = cfg.copy()
    cfg2.weight_p = 2 * cfg.weight_p
    ev2 = I.Evaluator(cfg2)
    ev2.set_map_u8(occ, [0, 0, 0], 1.0)
    ev2.set_shape_named("RoundedCone")
    d = ev2.eval_discrete(T, Cc)
    assert d[0] == 2 * a[0] and np.array_equal(grads(d), 2 * grads(a))
    ev2.set_map_u8(np.zeros((64, 64, 64), np.uint8), [0, 0, 0], 8.0)
    z = ev2.eval_discrete(T, Cc)
    assert z[0] == 0 and not grads(z).any()
    ev2.close()


def test_full_size_sampled_oracle_mesh(big):
    """oracle on the first 2 pieces (K=256) of the full-size workload with the 3900-triangle mesh robot"""
    cfg, occ, T, Cc = big
    N = 64
    sub = np.concatenate([Cc.reshape(3, 6 * N)[ax, :12] for ax in range(3)])
    V, F = W.rounded_cone_mesh()
    ev = I.Evaluator(cfg)
    ev.set_map_u8(occ, [0, 0, 0], 1.0)
    ev.set_shape_mesh(V, F, [0, 0, 0, 120, 0, 0])
    got = ev.eval_discrete(T[:2], sub)
    osh = O.Shape.mesh(V, F, [0, 0, 0, 120, 0, 0], wn_mode=O.WN_BH)
    exp = O.eval_discrete(O.config_from(cfg), occ, [0, 0, 0], 1.0, osh, T[:2], sub)
    assert abs(got[0] - exp[0]) <= 1e-6 * abs(exp[0]) and rel_l2(grads(got), np.concatenate([exp[1], exp[2]])) <= 1e-6
    assert ev.stats().last_pairs == exp[3]
    ev.close()

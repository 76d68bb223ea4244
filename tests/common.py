"""Shared fixtures/helpers for the parity tests."""
import numpy as np
import isdf_b200 as I
import workloads as W

BMIN = [0.0, 0.0, 0.0]


def rel_l2(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / d if d > 0 else np.linalg.norm(a - b)


def small_case(N=4, K=16, seed=3, noise=0.03, flags=None, kernel_size=13):
    cfg = I.default_config_values()
    cfg.integral_intervs = K
    cfg.kernel_size = kernel_size
    cfg.flags = flags if flags is not None else (I.WITH_DYNAMICS | I.WITH_COLLISION)   # the library default is dynamics only
    occ = W.three_slit_map(64, 64, 64, noise=noise, seed=seed)
    T, Cc, wp = W.make_trajectory(N, [0, 0, 0], [50, 50, 34], seed=seed, jitter=0.3)
    return cfg, occ, T, Cc, wp


def tilted(seed=0):
    """a non-trivial Rotate / trans (poly_params) for analytic shapes"""
    R, t = W.rotation_from_poly_params([0.3, -0.2, 0.1, 30.0, -20.0, 45.0])
    return R, t


MESHES = {
    "box": lambda: W.box_mesh(1.2, 0.7, 0.4),
    "lprism": lambda: W.l_prism_mesh(),
    "ico": lambda: W.icosphere(1.3, 2),
    "rcone": lambda: W.rounded_cone_mesh(n_theta=24, n_prof=16),
}

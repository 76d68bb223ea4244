"""Oracle pinned to REFERENCE-COMPILED code (kind "reference"), CPU only.

oracle/_ref/libref_flat.so is the reference's own utils/flatness.hpp (optimizated_forward :53-86 / :88-148, backwardthreadsafe
:230-406) compiled unmodified against an element-access-only Eigen stand-in (oracle/Makefile `ref`). The oracle's flatness map and its
HAND-DERIVED adjoint (oracle_math.hpp orc::Flat) must reproduce it to rounding; tests/golden/flat_reference.npz holds outputs of the
same reference build (made by tests/golden/make_reference_golden.py) so the pin also holds where oracle/_ref is absent."""
import os
import numpy as np
import pytest
import isdf_b200 as I
import oracle_lib as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_RANDOM = 10000


def flat_inputs(n, seed=0):
    rng = np.random.default_rng(seed)
    scale = 10.0 ** rng.uniform(-2, 1.3, size=(n, 1))                 # speeds / accelerations from cm/s to ~20 m/s
    v, a, j = rng.normal(size=(n, 3)) * scale, rng.normal(size=(n, 3)) * scale * 1.5, rng.normal(size=(n, 3)) * scale * 3
    a[:, 2] = np.abs(a[:, 2]) * 0.3 - 2.0                             # keeps thrust away from the singular zu = -|zu| e_z
    pg, vg, og = rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    qg = rng.normal(size=(n, 4))
    return v, a, j, pg, vg, qg, og


def relerr(x, ref):
    return np.max(np.abs(x - ref) / np.maximum(1.0, np.max(np.abs(ref), axis=-1, keepdims=True)))


def test_flatness_oracle_equals_reference_compiled_flatness_hpp():
    if not O.ref_flat_available():
        pytest.skip("oracle/_ref/libref_flat.so not built (needs /root/reference)")
    cfg = O.config_from(I.default_config_values())
    ref = O.RefFlat(cfg)
    v, a, j, pg, vg, qg, og = flat_inputs(N_RANDOM)
    q_ref, o_ref = ref.forward(v, a, j)
    q_only = ref.forward_quat(v, a, j)
    q, o = O.flat_forward_batch(cfg, v, a, j)
    assert np.array_equal(q_only, q_ref)
    assert relerr(q, q_ref) <= 1e-15 and relerr(o, o_ref) <= 1e-12
    b_ref = ref.backward(v, a, j, pg, vg, qg, og)
    b = O.flat_backward_batch(cfg, v, a, j, pg, vg, qg, og)
    assert relerr(b, b_ref) <= 1e-12, relerr(b, b_ref)
    assert np.array_equal(b[:, 0:3], pg)                               # pos_total_grad = pos_grad (flat:402-404)


def test_flatness_oracle_equals_committed_reference_outputs():
    z = np.load(os.path.join(G, "flat_reference.npz"))
    cfg = O.config_from(I.default_config_values())
    assert np.allclose(z["par"], [cfg.vehicle_mass, cfg.grav_acc, cfg.horiz_drag, cfg.vert_drag, cfg.paras_drag, cfg.speed_eps])
    q, o = O.flat_forward_batch(cfg, z["v"], z["a"], z["j"])
    b = O.flat_backward_batch(cfg, z["v"], z["a"], z["j"], z["pg"], z["vg"], z["qg"], z["og"])
    assert relerr(q, z["quat"]) <= 1e-15 and relerr(o, z["omg"]) <= 1e-12 and relerr(b, z["back"]) <= 1e-12

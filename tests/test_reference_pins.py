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


# ---- the robot meshes the reference ships (src/plan_manager/shapes/*.obj), committed as INPUT fixtures with reference-computed winding numbers
def ref_mesh_cases():
    z = np.load(os.path.join(G, "ref_meshes.npz"))
    return z, [str(n) for n in z["names"]]


@pytest.mark.parametrize("name", ["Lthick", "RoundedCone", "mybox", "drone", "kuang", "box"])
def test_reference_meshes_oracle_sign_and_distance(name):
    """oracle on the reference's own robot meshes: inside/outside == the reference-compiled FWN's (rounded), exact winding within the FWN's
    approximation error, BVH distance == brute force; and the reference-faithful mode s = 1 - 2 w_FWN differs from ±1 by ~1e-3."""
    z, _ = ref_mesh_cases()
    V, F, pp, q, w_ref = z[name + "_V"], z[name + "_F"], z[name + "_pp"], z[name + "_q"], z[name + "_w"]
    sh = O.Shape.mesh(V, F, pp)
    r = sh.mesh_query(q)
    assert np.array_equal(r["d2"], r["d2_brute"])
    keep = np.sqrt(r["d2"]) > 1e-3
    assert np.abs(r["w_exact"] - np.round(r["w_exact"])).max() < 1e-9          # closed, consistently oriented (what the product requires)
    assert np.abs(w_ref - r["w_exact"])[keep].max() < 5e-3
    sdf, _ = sh.query(q)
    assert np.array_equal(sdf[keep] < 0, w_ref[keep] > 0.5)
    if O.ref_fwn_available():
        sdf_ref, g_ref = O.Shape.mesh(V, F, pp, wn_mode=O.WN_REF).query(q)
        assert np.allclose(sdf_ref, (1 - 2 * w_ref) * np.sqrt(r["d2"]), rtol=1e-12, atol=1e-15)   # WN_REF mode = Shape.cpp:110-113 with the fixture's w
        dev = np.abs(sdf_ref - sdf)[keep] / np.abs(sdf)[keep]
        assert 1e-7 < dev.max() < 1e-2                                            # the FP32 order-2 tree's error reaches the SDF VALUE ...
        _, g = sh.query(q)
        assert np.abs(g_ref - g)[keep].max() < 1e-12                               # ... but not the (normalised) gradient


def test_obj_reader_matches_fixture_on_reference_files():
    """host/isdf_obj.hpp (read_triangle_mesh counterpart, Shape.cpp:36) on the reference's OBJ files — only where /root/reference exists"""
    import host_lib as H
    d = "/root/reference/src/plan_manager/shapes"
    if not os.path.isdir(d):
        pytest.skip("reference tree absent")
    z, names = ref_mesh_cases()
    for name in names:
        V, F = H.read_obj(os.path.join(d, name + ".obj"))
        assert np.array_equal(V, z[name + "_V"]) and np.array_equal(F, z[name + "_F"])
    # independent numpy parse of one file
    rows = [ln.split() for ln in open(os.path.join(d, "RoundedCone.obj")) if ln[:2] in ("v ", "f ")]
    Vn = np.array([[float(x) for x in r[1:4]] for r in rows if r[0] == "v"])
    Fn = np.array([[int(x.split("/")[0]) - 1 for x in r[1:4]] for r in rows if r[0] == "f"], dtype=np.int32)
    assert np.array_equal(Vn, z["RoundedCone_V"]) and np.array_equal(Fn, z["RoundedCone_F"])


def test_obj_reader_forms(tmp_path):
    """index forms i, i/t, i/t/n, i//n, negative (relative) indices, polygon fans, comments"""
    import host_lib as H
    p = tmp_path / "t.obj"
    p.write_text("# c\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0 1.0\nvn 0 0 1\nvt 0 0\nf 1/1/1 2/1/1 3/1/1 4/1/1\nv 0 0 1\nf -1 1//1 2\nf 1/1 3/1 5/1\n")
    V, F = H.read_obj(str(p))
    assert V.shape == (5, 3) and np.array_equal(F, [[0, 1, 2], [0, 2, 3], [4, 0, 1], [0, 2, 4]])

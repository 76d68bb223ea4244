"""The oracle against the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import os
import numpy as np
import isdf_b200 as I
import oracle_lib as O
from common import small_case, rel_l2, BMIN, MESHES

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_exact_winding_vs_reference_fwn_golden():
    """fwn_reference.npz holds answers computed by the reference's own FastWindingNumberForSoups.h."""
    z = np.load(os.path.join(G, "fwn_reference.npz"))
    for name in MESHES:
        sh = O.Shape.mesh(z[f"{name}_V"], z[f"{name}_F"])
        r = sh.mesh_query(z[f"{name}_q"], brute=False)
        d = np.sqrt(r["d2"])
        keep = d > 1e-3
        assert np.abs(r["w_exact"] - z[f"{name}_w"])[keep].max() < 5e-3
        assert np.array_equal(r["w_exact"][keep] > 0.5, z[f"{name}_w"][keep] > 0.5)


def test_shapes_golden():
    z = np.load(os.path.join(G, "shapes.npz"))
    for name in I.NAMED_SHAPES:
        s, g = O.Shape.named(name, z["R"], z["t"]).query(z["p"])
        assert np.allclose(s, z[f"{name}_sdf"], rtol=1e-13, atol=1e-13) and np.allclose(g, z[f"{name}_grad"], atol=1e-9)


def test_discrete_golden():
    z = np.load(os.path.join(G, "discrete.npz"))
    cfg, occ, T, Cc, _ = small_case(N=4, K=16, seed=3)
    assert np.array_equal(np.packbits(occ), z["occ_bits"]) and np.array_equal(T, z["T"]) and np.allclose(Cc, z["C"], rtol=1e-12)
    for name in ["Ball", "CSG", "Trefoil", "SmoothIntersection"]:
        c, gC, gT, npairs = O.eval_discrete(O.config_from(cfg), occ, BMIN, 1.0, O.Shape.named(name, z_R(), z_t()), z["T"], z["C"])
        assert npairs == int(z[f"{name}_pairs"])
        assert abs(c - float(z[f"{name}_cost"])) <= 1e-10 * abs(c) and rel_l2(gC, z[f"{name}_gradC"]) < 1e-10


def z_R():
    return np.load(os.path.join(G, "shapes.npz"))["R"]


def z_t():
    return np.load(os.path.join(G, "shapes.npz"))["t"]


def test_swept_golden():
    z = np.load(os.path.join(G, "swept.npz"))
    cfg, *_ = small_case(N=4, K=16, seed=3)
    for name in ["Ball", "Torus", "SmoothIntersection"]:
        r = O.eval_swept(O.config_from(cfg), O.Shape.named(name), z["T"], z["C"], z["pts"])
        assert np.abs(r["tstar"] - z[f"{name}_tstar"]).max() < 1e-9
        assert abs(r["cost"] - float(z[f"{name}_cost"])) <= 1e-9 * max(abs(r["cost"]), 1.0)

"""Generates the committed golden fixtures in tests/golden/.

  * fwn_reference.npz  — winding numbers produced by the REFERENCE's own code (igl/FastWindingNumberForSoups.h compiled
                         from /root/reference into oracle/_ref by oracle/Makefile). Only regenerable where the reference
                         tree exists; it is the one piece of the path the reference can answer for itself here.
  * shapes.npz, discrete.npz, swept.npz — outputs of the oracle (the CPU restatement) on seeded inputs; they pin the
                         oracle against silent drift and give the GPU tests an input/output pair that does not depend
                         on building the oracle.
Run:  python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py"))
import isdf_b200 as I      # noqa: E402
import oracle_lib as O     # noqa: E402
import workloads as W      # noqa: E402
from common import small_case, tilted, MESHES, BMIN   # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    # ---- reference winding numbers
    if O.ref_fwn_available():
        out = {}
        for name, gen in MESHES.items():
            V, F = gen()
            q = rng.uniform(-3, 5, size=(200, 3))
            out[f"{name}_V"], out[f"{name}_F"], out[f"{name}_q"] = V, F, q
            out[f"{name}_w"] = O.RefFwn(V, F, order=2).query(q, 2.0)
        np.savez_compressed(os.path.join(HERE, "fwn_reference.npz"), **out)
    # ---- shape queries
    R, t = tilted()
    out = {"R": R, "t": t}
    p = rng.uniform(-6, 6, size=(96, 3))
    out["p"] = p
    for name in I.NAMED_SHAPES:
        s, g = O.Shape.named(name, R, t).query(p)
        out[f"{name}_sdf"], out[f"{name}_grad"] = s, g
    s, g = O.Shape.analytic(I.SHAPE_KINDS["BOX"], [1.5, 0.15, 0.15], R, t).query(p)
    out["Box_sdf"], out["Box_grad"] = s, g
    for name, gen in MESHES.items():
        V, F = gen()
        s, g = O.Shape.mesh(V, F).query(p)
        out[f"mesh_{name}_sdf"], out[f"mesh_{name}_grad"] = s, g
    np.savez_compressed(os.path.join(HERE, "shapes.npz"), **out)
    # ---- discrete path
    cfg, occ, T, Cc, wp = small_case(N=4, K=16, seed=3)
    out = {"occ_bits": np.packbits(occ), "occ_shape": np.array(occ.shape), "T": T, "C": Cc, "K": 16, "seed": 3}
    for name in ["Ball", "CSG", "Trefoil", "SmoothIntersection"]:
        c, gC, gT, npairs = O.eval_discrete(O.config_from(cfg), occ, BMIN, 1.0, O.Shape.named(name, R, t), T, Cc)
        out[f"{name}_cost"], out[f"{name}_gradC"], out[f"{name}_gradT"], out[f"{name}_pairs"] = c, gC, gT, npairs
    V, F = MESHES["lprism"]()
    c, gC, gT, npairs = O.eval_discrete(O.config_from(cfg), occ, BMIN, 1.0, O.Shape.mesh(V, F, [0, 0, 0, 120, 0, 0]), T, Cc)
    out["mesh_cost"], out["mesh_gradC"], out["mesh_gradT"], out["mesh_pairs"] = c, gC, gT, npairs
    np.savez_compressed(os.path.join(HERE, "discrete.npz"), **out)
    # ---- swept path
    pts = W.gather_obstacle_points(occ, BMIN, 1.0, wp, cfg.kernel_size * cfg.occupancy_resolution / 3.0)[:160]
    out = {"pts": pts, "T": T, "C": Cc}
    for name in ["Ball", "Torus", "SmoothIntersection"]:
        r = O.eval_swept(O.config_from(cfg), O.Shape.named(name), T, Cc, pts)
        for k in ("cost", "gradC", "gradT", "tstar", "sdf", "grel", "nsdf"):
            out[f"{name}_{k}"] = r[k]
    np.savez_compressed(os.path.join(HERE, "swept.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()

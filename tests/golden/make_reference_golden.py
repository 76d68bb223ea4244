"""Golden vectors computed by REFERENCE-COMPILED code (oracle/_ref/*.so built from /root/reference by `make -C oracle ref`).
Run in the build container (where /root/reference exists):  python tests/golden/make_reference_golden.py
  flat_reference.npz — inputs and outputs of the reference's utils/flatness.hpp (optimizated_forward, backwardthreadsafe)."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "implicit-sdf-planner_b200", "py"))
import isdf_b200 as I          # noqa: E402
import oracle_lib as O         # noqa: E402
from test_reference_pins import flat_inputs   # noqa: E402


def main():
    O.build()
    cfg = O.config_from(I.default_config_values())
    ref = O.RefFlat(cfg)
    v, a, j, pg, vg, qg, og = flat_inputs(256, seed=123)
    q, o = ref.forward(v, a, j)
    b = ref.backward(v, a, j, pg, vg, qg, og)
    np.savez(os.path.join(HERE, "flat_reference.npz"), par=ref.par, v=v, a=a, j=j, pg=pg, vg=vg, qg=qg, og=og, quat=q, omg=o, back=b)
    print("flat_reference.npz written")


if __name__ == "__main__":
    main()

"""Generates tests/golden/tsdf_golden.npz by running the REFERENCE'S OWN kernel text (compiled for the
host by oracle/Makefile into oracle/_ref/libref_tsdf_host.so) on the seeded scenes of
tests/test_oracle_pinning.py, and tests/golden/color_golden.npz by running the reference's own
MeshColoration.cxx / ReconstructionData.cxx / Helper.h (oracle/_ref/libref_coloration.so, built against the VTK
stand-in of oracle/vtk_shim/) on the cases of tests/test_coloration_pinning.py.
Needs /root/reference (to build oracle/_ref); the fixtures it writes do not.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import _oracle  # noqa: E402
from tests.test_oracle_pinning import CASES, make_case  # noqa: E402


def main():
    ref = _oracle.load_ref_host()
    if ref is None:
        raise SystemExit("oracle/_ref/libref_tsdf_host.so unavailable (no reference tree?)")
    orc = _oracle.load_oracle()
    out = {}
    for name in sorted(CASES):
        s, dtype = make_case(name)
        filtered = orc.apply_depth_threshold(s.depths, s.best_cost, 0.14)   # ReconstructionData.cxx:159-166, trivial
        out[name] = ref.run(s.grid, s.rp, s.W, s.H, filtered, s.K, s.RT, s.zeros(dtype))
        print(name, out[name].dtype, out[name].shape, "nonzero:", np.count_nonzero(out[name]),
              "sum:", float(out[name].astype(np.float64).sum()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tsdf_golden.npz"), **out)

    from tests import test_coloration_pinning as tcp
    refc = _oracle.load_ref_coloration()
    if refc is None:
        raise SystemExit("oracle/_ref/libref_coloration.so unavailable (no reference tree?)")
    cout = {}
    for name in sorted(tcp.CASES):
        pts, colors, K, RT, W, H = tcp.CASES[name]()
        cout[name] = tcp.pack(*refc.colorize(pts, colors, K, RT, W, H))
        print(name, cout[name].shape, "seen:", int((cout[name][:, 6] > 0).sum()), "max views:", int(cout[name][:, 6].max()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "color_golden.npz"), **cout)


if __name__ == "__main__":
    main()

"""Regenerates tests/golden/*.npz from the from-paper oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference repository has no golden vectors
(/root/reference/README.md:1-2 is all there is), so these pin the ORACLE against regressions
and give the GPU tests a fixture that does not need the oracle to be rebuilt."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sac_cot_b200 import _abi, synth  # noqa: E402
from sac_cot_b200.api import Registrar  # noqa: E402

CASES = {
    # name: (N, inlier ratio, seed, K_e, m, score_mode, box, tau)
    "g256_r20": (256, 0.20, 901, 64, 4, 0, (3.0, 3.0, 3.0), 0.1),
    "g300_r10_mode1": (300, 0.10, 902, 48, 3, 1, (3.0, 3.0, 3.0), 0.1),
    "g200_kitti": (200, 0.15, 903, 32, 2, 0, (60.0, 60.0, 6.0), 0.6),
}
STAGES = {"adj": _abi.DBG_ADJ, "t_node": _abi.DBG_T_NODE, "top_edges": _abi.DBG_TOP_EDGES,
          "triangles": _abi.DBG_TRIANGLES, "hyp_rt": _abi.DBG_HYP_RT, "hyp_score": _abi.DBG_HYP_SCORE,
          "best_key": _abi.DBG_BEST_KEY, "mask": _abi.DBG_MASK, "hist": _abi.DBG_HIST}


def main():
    lib = _abi.bind(ctypes.CDLL(os.path.join(ROOT, "oracle", "libsaccot_oracle.so")))
    for name, (N, ratio, seed, Ke, m, mode, box, tau) in CASES.items():
        p = synth.make_pair(N, ratio, seed, box=box, tau_compat=tau)
        with Registrar(lib=lib, tau_compat=tau, tau_inlier=tau, num_edges=Ke, apex_per_edge=m, score_mode=mode) as reg:
            reg.set("keep_debug", 1)
            R, t, inl = reg.register(p.src, p.dst)
            out = {k: reg.debug(0, w) for k, w in STAGES.items()}
            out["edge_keys_sorted"] = np.sort(reg.debug(0, _abi.DBG_EDGE_KEYS))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), src=p.src, dst=p.dst, R=R, t=t,
                            inliers=np.int32(inl), params=np.array([tau, tau, Ke, m, mode], np.float64), **out)
        print(name, "inliers", inl)


if __name__ == "__main__":
    main()

"""The committed golden fixtures: the oracle must keep reproducing them (CPU), and the CUDA
path must match them bit for bit (GPU)."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT
from sac_cot_b200 import _abi, synth

FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
STAGES = {"adj": _abi.DBG_ADJ, "t_node": _abi.DBG_T_NODE, "top_edges": _abi.DBG_TOP_EDGES,
          "triangles": _abi.DBG_TRIANGLES, "hyp_rt": _abi.DBG_HYP_RT, "hyp_score": _abi.DBG_HYP_SCORE,
          "best_key": _abi.DBG_BEST_KEY, "mask": _abi.DBG_MASK, "hist": _abi.DBG_HIST}


def check(reg, path, exact_pose):
    g = np.load(path)
    tau_c, tau_i, Ke, m, mode = g["params"]
    reg.params.tau_compat, reg.params.tau_inlier = float(tau_c), float(tau_i)
    reg.params.num_edges, reg.params.apex_per_edge, reg.params.score_mode = int(Ke), int(m), int(mode)
    R, t, inl = reg.register(g["src"], g["dst"])
    for name, which in STAGES.items():
        got = reg.debug(0, which)
        assert got.shape == g[name].shape, name
        assert (got.view(np.uint8) == g[name].view(np.uint8)).all(), name
    np.testing.assert_array_equal(np.sort(reg.debug(0, _abi.DBG_EDGE_KEYS)), g["edge_keys_sorted"])
    assert inl == int(g["inliers"])
    if exact_pose:
        np.testing.assert_array_equal(R, g["R"])
        np.testing.assert_array_equal(t, g["t"])
    else:  # fp64 refit: 1e-5 rad / 1e-5 units (BASELINE.json north_star)
        ang, dt = synth.pose_error(R, t, g["R"].astype(np.float64), g["t"].astype(np.float64))
        assert ang < 1e-5 and dt < 1e-5


def test_fixtures_exist():
    assert len(FILES) >= 3


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_golden(oracle, path):
    check(oracle, path, exact_pose=True)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_matches_golden(gpu, path):
    check(gpu, path, exact_pose=False)

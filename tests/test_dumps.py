"""Correspondence dumps + recall evaluation.  The CPU tests run the evaluation on the oracle (test
infrastructure); the GPU test runs the same on the CUDA library."""
import numpy as np
import pytest

from sac_cot_b200 import dumps, synth
from sac_cot_b200.api import Registrar


def test_dump_round_trip_and_gt_log(tmp_path):
    paths = dumps.write_synthetic(str(tmp_path / "c"), "cfg1_single_n1000", 2)
    d = dumps.read_dump(paths[0])
    p = synth.make_config_pair("cfg1_single_n1000", 0)
    assert d.kind == "correspondences"
    np.testing.assert_array_equal(d.src, p.src)
    np.testing.assert_array_equal(d.dst, p.dst)
    np.testing.assert_allclose(d.T_gt[:3, :3], p.R_gt)
    assert d.labels.sum() == len(p.inlier_idx)
    poses = dumps.read_gt_log(str(tmp_path / "c" / "gt.log"))
    assert sorted(poses) == [(0, 1), (1, 2)]
    np.testing.assert_allclose(poses[(0, 1)], d.T_gt, rtol=0, atol=1e-9)
    paths = dumps.write_synthetic(str(tmp_path / "d"), "cfg1_single_n1000", 1, descriptors=True)
    d = dumps.read_dump(paths[0])
    assert d.kind == "descriptors" and d.feat0.shape == (1000, 33) and d.xyz1.shape == (1000, 3)
    assert len(list(dumps.iter_dumps(str(tmp_path / "d")))) == 1
    np.savez(tmp_path / "bad.npz", x=np.zeros(3))
    with pytest.raises(ValueError):
        dumps.read_dump(str(tmp_path / "bad.npz"))
    (tmp_path / "short.log").write_text("0 1 2\n1 0 0 0\n")
    with pytest.raises(ValueError):
        dumps.read_gt_log(str(tmp_path / "short.log"))


def test_recall_report_on_the_oracle(tmp_path, oracle_lib):
    dumps.write_synthetic(str(tmp_path / "c"), "cfg1_single_n1000", 3)
    dumps.write_synthetic(str(tmp_path / "d"), "cfg1_single_n1000", 2, descriptors=True)
    with Registrar(lib=oracle_lib) as reg:
        rep = dumps.evaluate(reg, dumps.iter_dumps(str(tmp_path / "c")), "3dmatch", tau=0.1)
        assert rep["pairs"] == rep["pairs_with_ground_truth"] == 3 and rep["recall"] == 1.0
        assert rep["mean_re_deg_success"] < 1.0 and rep["mean_te_success"] < 0.02
        assert 0.09 < rep["mean_inlier_ratio"] < 0.13           # 10 % inliers plus a few chance hits
        rep_d = dumps.evaluate(reg, dumps.iter_dumps(str(tmp_path / "d")), "kitti", tau=0.1)
        assert rep_d["recall"] == 1.0 and all(r["kind"] == "descriptors" for r in rep_d["per_pair"])
        # a wrong ground truth is reported as a failure, not hidden
        d = dumps.read_dump(str(tmp_path / "c" / "pair_0000.npz"))
        d.T_gt = np.eye(4)
        rep_bad = dumps.evaluate(reg, [d], "3dmatch")
        assert rep_bad["recall"] == 0.0 and rep_bad["mean_re_deg_success"] is None
        with pytest.raises(ValueError):
            dumps.evaluate(reg, [], "nuscenes")


@pytest.mark.gpu
def test_recall_report_on_the_cuda_library(tmp_path, gpu_lib, oracle_lib):
    dumps.write_synthetic(str(tmp_path / "c"), "cfg2_3dmatch_256x5000", 3)
    dumps.write_synthetic(str(tmp_path / "d"), "cfg2_3dmatch_256x5000", 2, descriptors=True)
    with Registrar(lib=gpu_lib) as reg, Registrar(lib=oracle_lib) as ora:
        for d in ("c", "d"):
            rep = dumps.evaluate(reg, dumps.iter_dumps(str(tmp_path / d)), "3dmatch", tau=0.1)
            ref = dumps.evaluate(ora, dumps.iter_dumps(str(tmp_path / d)), "3dmatch", tau=0.1)
            assert rep["recall"] == ref["recall"] == 1.0
            assert [r["inliers"] for r in rep["per_pair"]] == [r["inliers"] for r in ref["per_pair"]]

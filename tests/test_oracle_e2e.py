"""End-to-end behaviour of the oracle through the C ABI: pose recovery, argument checking,
batch/packed equivalence, degenerate inputs, sharded-phase equivalence.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from sac_cot_b200 import _abi, synth
from sac_cot_b200.api import Registrar, SacCotError


@pytest.mark.parametrize("N,ratio,seed", [(1000, 0.10, 1), (1000, 0.05, 2), (2000, 0.03, 3), (2000, 0.02, 4)])
def test_recovers_ground_truth_pose(oracle, N, ratio, seed):
    p = synth.make_pair(N, ratio, seed)
    R, t, inl = oracle.register(p.src, p.dst)
    ang, dt = synth.pose_error(R, t, p.R_gt, p.t_gt)
    assert ang < np.deg2rad(1.0) and dt < 0.02, (ang, dt)
    assert inl >= len(p.inlier_idx) * 0.95


def test_all_inlier_noise_free_is_complete_graph_and_exact(oracle):
    N = 200
    p = synth.make_pair(N, 1.0, 9)
    dst = (p.src.astype(np.float64) @ p.R_gt.T + p.t_gt).astype(np.float32)
    R, t, inl = oracle.register(p.src, dst)
    assert inl == N
    assert int(oracle.debug(0, _abi.DBG_NUM_EDGES)[0]) == N * (N - 1) // 2
    t_node = oracle.debug(0, _abi.DBG_T_NODE)
    assert (t_node == (N - 1) * (N - 2) // 2).all()
    ang, dt = synth.pose_error(R, t, p.R_gt, p.t_gt)
    assert ang < 1e-5 and dt < 1e-5


def test_no_triangle_is_a_result_not_an_error(oracle):
    # points far apart with inconsistent lengths: empty graph -> R=I, t=0, inliers=0, status OK
    rng = np.random.default_rng(0)
    src = (rng.random((40, 3)) * 100).astype(np.float32)
    dst = (rng.random((40, 3)) * 100).astype(np.float32)
    oracle.params.tau_compat = 1e-4
    R, t, inl = oracle.register(src, dst)
    if int(oracle.debug(0, _abi.DBG_BEST_KEY)[0]) == 0:
        np.testing.assert_array_equal(R, np.eye(3, dtype=np.float32))
        np.testing.assert_array_equal(t, np.zeros(3, np.float32))
        assert inl == 0
    else:  # extremely unlikely, but then the result must still be a rotation
        assert abs(np.linalg.det(R.astype(np.float64)) - 1) < 1e-4


def test_argument_checking(oracle_lib):
    lib = oracle_lib
    p = _abi.default_params(lib)
    src = np.zeros((10, 3), np.float32)
    R = np.zeros(9, np.float32)
    t = np.zeros(3, np.float32)
    inl = C.c_int32()
    f = _abi.fptr
    assert lib.sac_cot_register(None, f(src), 10, C.byref(p), f(R), f(t), C.byref(inl)) == _abi.E_NULL
    assert lib.sac_cot_register(f(src), f(src), 2, C.byref(p), f(R), f(t), C.byref(inl)) == _abi.E_SIZE
    assert lib.sac_cot_register(f(src), f(src), _abi.MAX_N + 1, C.byref(p), f(R), f(t), C.byref(inl)) == _abi.E_SIZE
    assert lib.sac_cot_register(f(src), f(src), 10, None, f(R), f(t), C.byref(inl)) == _abi.E_NULL
    for field, bad in [("struct_size", 4), ("tau_compat", 0.0), ("tau_inlier", -1.0), ("num_edges", 0),
                       ("num_edges", _abi.MAX_EDGES + 1), ("apex_per_edge", 0), ("apex_per_edge", 9),
                       ("score_mode", 2), ("refit", 2), ("reserved", 1)]:
        q = _abi.default_params(lib, **{field: bad})
        assert lib.sac_cot_register(f(src), f(src), 10, C.byref(q), f(R), f(t), C.byref(inl)) == _abi.E_PARAMS, field
    assert lib.sac_cot_strerror(_abi.E_PARAMS)
    assert b"oracle" in lib.sac_cot_version()
    with Registrar(lib=lib) as reg:
        with pytest.raises(SacCotError):
            reg.set("no_such_knob", 1)
        with pytest.raises(SacCotError):
            reg.debug(5, _abi.DBG_ADJ)


def test_batch_equals_singles_and_packed_equals_pointer(oracle):
    pairs = [synth.make_pair(n, 0.1, 100 + k) for k, n in enumerate((300, 517, 64, 1000))]
    res = oracle.register_batch([p.src for p in pairs], [p.dst for p in pairs])
    res2 = oracle.register_pointer_batch([p.src for p in pairs], [p.dst for p in pairs])
    np.testing.assert_array_equal(res.R, res2.R)
    np.testing.assert_array_equal(res.t, res2.t)
    np.testing.assert_array_equal(res.inliers, res2.inliers)
    for b, p in enumerate(pairs):
        R, t, inl = oracle.register(p.src, p.dst)
        np.testing.assert_array_equal(R, res.R[b])
        np.testing.assert_array_equal(t, res.t[b])
        assert inl == res.inliers[b]
    empty = oracle.register_batch([], [])
    assert empty.R.shape == (0, 3, 3)


def test_openmp_build_is_thread_count_independent(oracle_lib, oracle_omp_lib):
    p = synth.make_pair(1200, 0.05, 5)
    outs = []
    for lib in (oracle_lib, oracle_omp_lib):
        with Registrar(lib=lib) as reg:
            reg.set("keep_debug", 1)
            R, t, inl = reg.register(p.src, p.dst)
            outs.append((R.copy(), t.copy(), inl, reg.debug(0, _abi.DBG_ADJ), reg.debug(0, _abi.DBG_T_NODE),
                         reg.debug(0, _abi.DBG_TRIANGLES), reg.debug(0, _abi.DBG_HYP_SCORE)))
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_phases_equal_unsharded(oracle_lib, world):
    N = 1500
    p = synth.make_pair(N, 0.05, 66)
    with Registrar(lib=oracle_lib) as ref:
        R0, t0, inl0 = ref.register(p.src, p.dst)
    regs = [Registrar(lib=oracle_lib) for _ in range(world)]
    try:
        ph1 = [regs[g].sharded_phase1(p.src, p.dst, g, world) for g in range(world)]
        t_all = np.stack([x[0] for x in ph1])
        cand_all = np.stack([x[1] for x in ph1])
        # every edge is owned by exactly one rank: partial sums add up to 2*t_i
        with Registrar(lib=oracle_lib) as one:
            one.set("keep_debug", 1)
            one.register(p.src, p.dst)
            np.testing.assert_array_equal(t_all.sum(0), 2 * one.debug(0, _abi.DBG_T_NODE).astype(np.uint64))
        best = max(regs[g].sharded_phase2(t_all, cand_all) for g in range(world))
        for g in range(world):
            R, t, inl = regs[g].sharded_phase3(best)
            np.testing.assert_array_equal(R, R0)
            np.testing.assert_array_equal(t, t0)
            assert inl == inl0
    finally:
        for r in regs:
            r.close()

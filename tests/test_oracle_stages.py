"""Pins the from-paper oracle against independent witnesses.  CPU only."""
import networkx as nx
import numpy as np
import pytest

import witness as W
from sac_cot_b200 import _abi, synth


def run(oracle, pair, **prm):
    for k, v in prm.items():
        setattr(oracle.params, k, v)
    oracle.params.tau_compat = pair.tau_compat
    oracle.params.tau_inlier = pair.tau_inlier
    return oracle.register(pair.src, pair.dst)


@pytest.mark.parametrize("N,seed", [(3, 5), (31, 6), (32, 7), (33, 8), (127, 9), (128, 10), (129, 11), (257, 12), (600, 13)])
def test_graph_matches_numpy_same_order(oracle, N, seed):
    p = synth.make_pair(N, 0.2, seed)
    run(oracle, p, num_edges=8, apex_per_edge=2)
    words = oracle.debug(0, _abi.DBG_ADJ)
    assert words.size == N * W.stride_words(N)
    A, pad = W.unpack_adj(words, N)
    assert not pad.any(), "pad bits must be zero"
    assert not A.diagonal().any()
    assert (A == A.T).all()
    np.testing.assert_array_equal(A, W.graph_dense(p.src, p.dst, p.tau_compat))
    np.testing.assert_array_equal(words.reshape(N, -1), W.pack_adj(A))


@pytest.mark.parametrize("N,ratio,seed", [(64, 0.3, 1), (200, 0.1, 2), (500, 0.05, 3)])
def test_triangle_counts_match_networkx_and_matmul(oracle, N, ratio, seed):
    p = synth.make_pair(N, ratio, seed)
    run(oracle, p, num_edges=16, apex_per_edge=2)
    A, _ = W.unpack_adj(oracle.debug(0, _abi.DBG_ADJ), N)
    T, t = W.triangle_counts(A)
    t_or = oracle.debug(0, _abi.DBG_T_NODE)
    np.testing.assert_array_equal(t_or, t)
    G = nx.from_numpy_array(A.astype(np.uint8))
    tri = nx.triangles(G)
    np.testing.assert_array_equal(t_or, np.array([tri[i] for i in range(N)]))
    assert int(t_or.sum()) % 3 == 0
    assert int(t_or.sum()) // 3 == int(np.trace(np.linalg.matrix_power(A.astype(np.int64), 3))) // 6
    keys = np.sort(oracle.debug(0, _abi.DBG_EDGE_KEYS))
    np.testing.assert_array_equal(keys, np.sort(W.edge_keys(A, T)))
    assert int(oracle.debug(0, _abi.DBG_NUM_EDGES)[0]) == keys.size == int(np.triu(A, 1).sum())
    hist = oracle.debug(0, _abi.DBG_HIST)
    np.testing.assert_array_equal(hist, np.bincount((keys >> np.uint64(36)).astype(np.int64), minlength=4096))


@pytest.mark.parametrize("Ke,m", [(1, 1), (7, 3), (64, 4), (4096, 8)])
def test_edge_ranking_and_apex_selection(oracle, Ke, m):
    N = 300
    p = synth.make_pair(N, 0.1, 21)
    run(oracle, p, num_edges=Ke, apex_per_edge=m)
    A, _ = W.unpack_adj(oracle.debug(0, _abi.DBG_ADJ), N)
    T, t = W.triangle_counts(A)
    keys = W.edge_keys(A, T)
    want = np.sort(keys)[::-1][:Ke]
    top = oracle.debug(0, _abi.DBG_TOP_EDGES)
    np.testing.assert_array_equal(top, want)
    # order really is (T desc, i asc, j asc)
    dec = [W.decode_edge(k) for k in top]
    assert dec == sorted(dec, key=lambda e: (-e[0], e[1], e[2]))
    tri = oracle.debug(0, _abi.DBG_TRIANGLES).reshape(-1, 3)
    assert tri.shape[0] == Ke * m
    np.testing.assert_array_equal(tri, W.select_triangles(A, t, top, m, Ke * m))


def test_kabsch3_matches_fp64_svd(oracle):
    N = 400
    p = synth.make_pair(N, 0.3, 33)
    run(oracle, p, num_edges=128, apex_per_edge=4)
    tri = oracle.debug(0, _abi.DBG_TRIANGLES).reshape(-1, 3)
    rt = oracle.debug(0, _abi.DBG_HYP_RT).reshape(-1, 12)
    checked = 0
    for h in range(tri.shape[0]):
        if tri[h, 0] < 0:
            assert not rt[h].any()
            continue
        idx = tri[h]
        P, Q = p.src[idx].astype(np.float64), p.dst[idx].astype(np.float64)
        R = rt[h, :9].astype(np.float64).reshape(3, 3)
        # proper rotation to fp32 accuracy
        assert np.abs(R @ R.T - np.eye(3)).max() < 2e-6
        assert abs(np.linalg.det(R) - 1.0) < 2e-6
        # conditioning of the 3-point problem: skip near-collinear triangles for the angle check
        sv = np.linalg.svd(P - P.mean(0), compute_uv=False)
        if sv[1] < 0.05 * sv[0] or sv[1] < 0.05:
            continue
        Rw, tw = W.kabsch_svd(P, Q)
        assert W.rot_angle(R, Rw) < 2e-4, (h, W.rot_angle(R, Rw))
        assert np.abs(rt[h, 9:] - tw).max() < 2e-3
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("mode", [0, 1])
def test_scoring_and_argmax(oracle, mode):
    N = 500
    p = synth.make_pair(N, 0.1, 44)
    run(oracle, p, num_edges=32, apex_per_edge=4, score_mode=mode)
    rt = oracle.debug(0, _abi.DBG_HYP_RT).reshape(-1, 12)
    tri = oracle.debug(0, _abi.DBG_TRIANGLES).reshape(-1, 3)
    keys = oracle.debug(0, _abi.DBG_HYP_SCORE)
    best = int(oracle.debug(0, _abi.DBG_BEST_KEY)[0])
    assert best == int(keys.max())
    tau2 = float(np.float32(p.tau_inlier) * np.float32(p.tau_inlier))
    for h in range(len(keys)):
        k = int(keys[h])
        if tri[h, 0] < 0:
            assert k == 0
            continue
        assert (k & 0xFFFF) == 0xFFFF - h
        score = k >> 16
        if mode == 0:
            lo, hi = W.count_bounds(rt[h], p.src, p.dst, p.tau_inlier)
            assert lo <= score - 1 <= hi, (h, lo, score - 1, hi)
        else:
            r2 = W.residual2_f64(rt[h], p.src, p.dst)
            q = np.minimum(r2, tau2) / tau2
            want = float(np.floor(q * 2 ** 20).sum())
            got = (N << 20) - (score - 1)
            assert abs(got - want) <= N * 4 + want * 1e-5, (h, got, want)  # floor + fp32 rounding slack
    # ties -> lowest hypothesis id: the winner is the first index holding the maximal score
    scores = keys >> np.uint64(16)
    h_best = 0xFFFF - (best & 0xFFFF)
    assert h_best == int(np.argmax(scores))


def test_inlier_mask_and_refit(oracle):
    N = 800
    p = synth.make_pair(N, 0.1, 55)
    R, t, inl = run(oracle, p)
    mask = oracle.debug(0, _abi.DBG_MASK)
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:N].astype(bool)
    assert bits.sum() == inl
    # the true inliers are found (a few outliers may land inside tau by chance)
    assert bits[p.inlier_idx].mean() > 0.95
    Rw, tw = W.kabsch_svd(p.src[bits], p.dst[bits])
    assert W.rot_angle(R, Rw) < 1e-5
    assert np.abs(t - tw).max() < 1e-5
    # refit = 0 returns the winning hypothesis unchanged
    R0, t0, inl0 = run(oracle, p, refit=0)
    best = int(oracle.debug(0, _abi.DBG_BEST_KEY)[0])
    rt = oracle.debug(0, _abi.DBG_HYP_RT).reshape(-1, 12)[0xFFFF - (best & 0xFFFF)]
    np.testing.assert_array_equal(np.r_[R0.ravel(), t0], rt)
    assert inl0 == inl


def test_inliers_form_a_clique_in_fp32(oracle):
    # sigma = tau_c/4 must make the true inliers pairwise compatible, fp32 rounding included
    for name in ("cfg2_3dmatch_256x5000", "cfg4_kitti_128x10000"):
        c = synth.CONFIGS[name]
        p = synth.make_pair(1500, 0.1, 77, box=c["box"], tau_compat=c["tau"])
        A = W.graph_dense(p.src, p.dst, p.tau_compat)
        sub = A[np.ix_(p.inlier_idx, p.inlier_idx)]
        assert sub.sum() == len(p.inlier_idx) * (len(p.inlier_idx) - 1)


# ---- second-order compatibility: A2 = A and ((A.A) o A >= c) ---------------------------
@pytest.mark.parametrize("N,ratio,cmin,seed", [(200, 0.2, 0, 31), (200, 0.2, 12, 32), (500, 0.05, 6, 33), (300, 0.1, 10 ** 4, 34)])
def test_second_order_graph_matches_matmul_witness(oracle, N, ratio, cmin, seed):
    p = synth.make_pair(N, ratio, seed)
    R, t, inl = run(oracle, p, num_edges=64, apex_per_edge=4, compat_mode=_abi.COMPAT_SECOND_ORDER, so_min_common=cmin)
    A, _ = W.unpack_adj(oracle.debug(0, _abi.DBG_ADJ_FIRST), N)
    np.testing.assert_array_equal(A, W.graph_dense(p.src, p.dst, p.tau_compat))
    C1, _ = W.triangle_counts(A)                      # (A.A) o A: common neighbours on the edges of A
    A2_want = A & (C1 >= cmin)
    A2, pad = W.unpack_adj(oracle.debug(0, _abi.DBG_ADJ), N)
    assert not pad.any()
    np.testing.assert_array_equal(A2, A2_want)
    if cmin == 0:
        np.testing.assert_array_equal(A2, A)
    # every later stage ran on A2
    T2, t2 = W.triangle_counts(A2)
    np.testing.assert_array_equal(oracle.debug(0, _abi.DBG_T_NODE), t2)
    keys = np.sort(oracle.debug(0, _abi.DBG_EDGE_KEYS))
    np.testing.assert_array_equal(keys, np.sort(W.edge_keys(A2, T2)))
    want_top = np.sort(W.edge_keys(A2, T2))[::-1][:64]
    np.testing.assert_array_equal(oracle.debug(0, _abi.DBG_TOP_EDGES), want_top)
    tri = oracle.debug(0, _abi.DBG_TRIANGLES).reshape(-1, 3)
    np.testing.assert_array_equal(tri, W.select_triangles(A2, t2, want_top, 4, 64 * 4))
    if not A2.any():
        assert inl == 0   # no edge survives: "no valid triangle" is a result


def test_second_order_mode_filters_outlier_edges_and_recovers_the_pose(oracle):
    # at 5 % inliers the inlier-inlier edges share ~N_in neighbours, random edges ~p^2 N: a threshold in between keeps
    # the inlier clique and drops most of the rest
    N = 1500
    p = synth.make_pair(N, 0.05, 35)
    run(oracle, p, compat_mode=_abi.COMPAT_FIRST_ORDER)
    e1 = int(oracle.debug(0, _abi.DBG_NUM_EDGES)[0])
    R, t, inl = run(oracle, p, compat_mode=_abi.COMPAT_SECOND_ORDER, so_min_common=40)
    e2 = int(oracle.debug(0, _abi.DBG_NUM_EDGES)[0])
    assert e2 < e1 / 5
    A2, _ = W.unpack_adj(oracle.debug(0, _abi.DBG_ADJ), N)
    idx = p.inlier_idx
    assert A2[np.ix_(idx, idx)].sum() == len(idx) * (len(idx) - 1)   # the inlier clique survives
    ang, dt = synth.pose_error(R, t, p.R_gt, p.t_gt)
    assert ang < np.deg2rad(1.0) and dt < 0.02 and inl >= len(idx)


def test_version1_params_struct_still_accepted(oracle_lib):
    import ctypes as C
    p = synth.make_pair(200, 0.2, 36)
    R = np.zeros(9, np.float32); t = np.zeros(3, np.float32); inl = C.c_int32()
    prm = _abi.default_params(oracle_lib)
    want = oracle_lib.sac_cot_register(_abi.fptr(p.src), _abi.fptr(p.dst), 200, C.byref(prm), _abi.fptr(R), _abi.fptr(t), C.byref(inl))
    assert want == 0
    R2 = np.zeros(9, np.float32); t2 = np.zeros(3, np.float32); inl2 = C.c_int32()
    prm.struct_size = _abi.PARAMS_SIZE_V1
    prm.so_min_common = 12345          # beyond the version-1 struct: must be ignored
    assert oracle_lib.sac_cot_register(_abi.fptr(p.src), _abi.fptr(p.dst), 200, C.byref(prm), _abi.fptr(R2), _abi.fptr(t2), C.byref(inl2)) == 0
    np.testing.assert_array_equal(R, R2)
    assert inl.value == inl2.value
    prm.compat_mode = 1                # version 1 called this field `reserved`: must be 0
    assert oracle_lib.sac_cot_register(_abi.fptr(p.src), _abi.fptr(p.dst), 200, C.byref(prm), _abi.fptr(R2), _abi.fptr(t2), C.byref(inl2)) == _abi.E_PARAMS
    prm.struct_size = 36
    assert oracle_lib.sac_cot_register(_abi.fptr(p.src), _abi.fptr(p.dst), 200, C.byref(prm), _abi.fptr(R2), _abi.fptr(t2), C.byref(inl2)) == _abi.E_PARAMS

"""Pins the oracle's correspondence front end (descriptor nearest-neighbour matching) against a
float64 numpy witness that shares no code with it.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from sac_cot_b200 import _abi, synth


def witness_d64(f, g):
    f = np.asarray(f, np.float64)
    g = np.asarray(g, np.float64)
    return ((f[:, None, :] - g[None, :, :]) ** 2).sum(-1)


def fp32_chain(f, g):
    """The specified chain D = fma(e, e, D), c ascending, emulated with float64 products (exact) and one rounding to
    float32 per step (a double rounding could differ from a true fma by one ulp in rare cases: callers allow it)."""
    f = np.asarray(f, np.float32)
    g = np.asarray(g, np.float32)
    D = np.zeros((f.shape[0], g.shape[0]), np.float32)
    for c in range(f.shape[1]):
        e = (f[:, None, c] - g[None, :, c]).astype(np.float32)
        D = (e.astype(np.float64) * e.astype(np.float64) + D.astype(np.float64)).astype(np.float32)
    return D


@pytest.mark.parametrize("Ns,Nd,dim,seed", [(1, 1, 33, 1), (7, 300, 33, 2), (300, 7, 33, 3), (500, 640, 33, 4), (200, 200, 8, 5),
                                            (100, 150, 64, 6)])
def test_nearest_neighbour_matches_float64_witness(oracle, Ns, Nd, dim, seed):
    rng = np.random.default_rng(seed)
    f = (rng.random((Ns, dim)) * 10).astype(np.float32)
    g = (rng.random((Nd, dim)) * 10).astype(np.float32)
    xs = rng.random((Ns, 3)).astype(np.float32)
    xd = rng.random((Nd, 3)).astype(np.float32)
    nn, cs, cd = oracle.match(f, xs, g, xd)
    assert nn.shape == (Ns,) and nn.dtype == np.int32 and (0 <= nn).all() and (nn < Nd).all()
    D = witness_d64(f, g)
    got = D[np.arange(Ns), nn]
    assert (got <= D.min(axis=1) * (1 + 1e-5) + 1e-12).all()
    # against the emulated fp32 chain: identical wherever the emulated minimum is unique by more than one ulp
    D32 = fp32_chain(f, g)
    arg = D32.argmin(axis=1)
    second = np.partition(D32, 1, axis=1)[:, 1] if Nd > 1 else np.full(Ns, np.inf)
    clear = second > D32.min(axis=1) * (1 + 3e-7)
    np.testing.assert_array_equal(nn[clear], arg[clear])
    np.testing.assert_array_equal(cs, xs)
    np.testing.assert_array_equal(cd, xd[nn])


def test_exact_ties_go_to_the_lowest_index(oracle):
    rng = np.random.default_rng(7)
    g = rng.random((50, 33)).astype(np.float32)
    g[40] = g[11]
    g[23] = g[11]
    f = g[[11, 3, 40]].copy()
    nn, _, _ = oracle.match(f, np.zeros((3, 3), np.float32), g, np.zeros((50, 3), np.float32))
    np.testing.assert_array_equal(nn, [11, 3, 11])
    # all descriptors identical: every distance ties, index 0 wins
    nn, _, _ = oracle.match(np.ones((4, 33), np.float32), np.zeros((4, 3), np.float32), np.ones((9, 33), np.float32),
                            np.zeros((9, 3), np.float32))
    np.testing.assert_array_equal(nn, [0, 0, 0, 0])


def test_matching_feeds_registration(oracle):
    p = synth.make_pair(800, 0.2, 41)
    f, g = synth.make_descriptors(p, 33, seed=41)
    nn, cs, cd = oracle.match(f, p.src, g, p.dst)
    assert (nn[p.inlier_idx] == p.inlier_idx).all()          # inliers find their true partner
    wrong = np.setdiff1d(np.arange(800), p.inlier_idx)
    assert (nn[wrong] != wrong).mean() > 0.95                # outliers get arbitrary partners
    oracle.params.tau_compat = oracle.params.tau_inlier = 0.1
    R, t, inl = oracle.register(cs, cd)
    ang, dt = synth.pose_error(R, t, p.R_gt, p.t_gt)
    assert ang < np.deg2rad(1.0) and dt < 0.02 and inl >= len(p.inlier_idx)


def test_batch_and_argument_checks(oracle, oracle_lib):
    rng = np.random.default_rng(8)
    sizes = [(30, 40), (1, 5), (64, 3)]
    fs = [rng.random((a, 16)).astype(np.float32) for a, _ in sizes]
    gs = [rng.random((b, 16)).astype(np.float32) for _, b in sizes]
    xs = [rng.random((a, 3)).astype(np.float32) for a, _ in sizes]
    xd = [rng.random((b, 3)).astype(np.float32) for _, b in sizes]
    nn, cs, cd, offs = oracle.match_batch(fs, xs, gs, xd)
    np.testing.assert_array_equal(offs, [0, 30, 31, 95])
    for b in range(3):
        one, c1, c2 = oracle.match(fs[b], xs[b], gs[b], xd[b])
        np.testing.assert_array_equal(nn[offs[b]:offs[b + 1]], one)
        np.testing.assert_array_equal(cd[offs[b]:offs[b + 1]], c2)
    # argument checking
    ctx = C.c_void_p()
    assert oracle_lib.sac_cot_ctx_create(C.byref(ctx), 0, None) == 0
    offs2 = np.array([0, 4], np.int64)
    o = offs2.ctypes.data_as(C.POINTER(C.c_int64))
    buf = np.zeros(4 * 33, np.float32)
    out = np.zeros(64, np.float32)
    nnb = np.zeros(4, np.int32)
    call = lambda dim, B=1, p=buf.ctypes.data: oracle_lib.sac_cot_match_packed(  # noqa: E731
        ctx, p, buf.ctypes.data, o, buf.ctypes.data, buf.ctypes.data, o, B, dim, nnb.ctypes.data, out.ctypes.data,
        out.ctypes.data, _abi.LOC_HOST)
    assert call(33) == 0
    assert call(0) == _abi.E_SIZE and call(257) == _abi.E_SIZE and call(33, -1) == _abi.E_SIZE
    assert call(33, 1, None) == _abi.E_NULL
    empty = np.array([0, 0], np.int64).ctypes.data_as(C.POINTER(C.c_int64))
    assert oracle_lib.sac_cot_match_packed(ctx, buf.ctypes.data, buf.ctypes.data, empty, buf.ctypes.data, buf.ctypes.data, o, 1, 33,
                                           nnb.ctypes.data, out.ctypes.data, out.ctypes.data, _abi.LOC_HOST) == _abi.E_SIZE
    oracle_lib.sac_cot_ctx_destroy(ctx)


def test_mutual_filter_matches_numpy_and_raises_the_inlier_ratio(oracle):
    ps = [synth.make_pair(600, 0.2, 50 + k) for k in range(3)]
    ds = [synth.make_descriptors(p, 33, seed=k) for k, p in enumerate(ps)]
    F, G = [d[0] for d in ds], [d[1] for d in ds]
    XS, XD = [p.src for p in ps], [p.dst for p in ps]
    cs, cd, offs = oracle.match_mutual_batch(F, XS, G, XD)
    for b, p in enumerate(ps):
        D = witness_d64(F[b], G[b])
        nn, nb = D.argmin(axis=1), D.argmin(axis=0)
        keep = np.nonzero(nb[nn] == np.arange(600))[0]
        assert offs[b + 1] - offs[b] == len(keep)
        np.testing.assert_array_equal(cs[offs[b]:offs[b + 1]], p.src[keep])
        np.testing.assert_array_equal(cd[offs[b]:offs[b + 1]], p.dst[nn[keep]])
        assert np.isin(p.inlier_idx, keep).all()                       # every true match is mutual
        assert len(p.inlier_idx) / len(keep) > 0.3                     # 20 % inliers before the filter
        oracle.params.tau_compat = oracle.params.tau_inlier = 0.1
        R, t, inl = oracle.register(cs[offs[b]:offs[b + 1]], cd[offs[b]:offs[b + 1]])
        ang, dt = synth.pose_error(R, t, p.R_gt, p.t_gt)
        assert ang < np.deg2rad(1.0) and dt < 0.02

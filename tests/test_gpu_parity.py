"""GPU parity: the CUDA path through the C ABI against the from-paper oracle on the same seeded
inputs.  Bit-exact for the compatibility graph, triangle counts, edge ranking, triangles,
per-hypothesis (R,t) bits, per-hypothesis scores, the winning id and its inlier mask; the fp64
refit within 1e-5 rad / 1e-5 units (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

from sac_cot_b200 import _abi, synth
from sac_cot_b200.api import Registrar, SacCotError

pytestmark = pytest.mark.gpu

EXACT = [("adj", _abi.DBG_ADJ), ("num_edges", _abi.DBG_NUM_EDGES), ("hist", _abi.DBG_HIST),
         ("t_node", _abi.DBG_T_NODE), ("top_edges", _abi.DBG_TOP_EDGES), ("triangles", _abi.DBG_TRIANGLES),
         ("hyp_rt", _abi.DBG_HYP_RT), ("hyp_score", _abi.DBG_HYP_SCORE), ("best_key", _abi.DBG_BEST_KEY),
         ("mask", _abi.DBG_MASK)]
TOL_ANG = 1e-5   # rad
TOL_T = 1e-5     # translation units


def set_params(reg, **kw):
    for k, v in kw.items():
        setattr(reg.params, k, v)


def compare_stages(gpu, oracle, pair_idx=0, stages=EXACT, edge_keys=True):
    bad = []
    for name, which in stages:
        a, b = gpu.debug(pair_idx, which), oracle.debug(pair_idx, which)
        if a.shape != b.shape:
            bad.append(f"{name}: shape {a.shape} vs {b.shape}")
        elif not (a.view(np.uint8) == b.view(np.uint8)).all():
            n = int((a.view(np.uint8) != b.view(np.uint8)).sum())
            bad.append(f"{name}: {n} differing bytes of {a.nbytes}")
    if edge_keys:
        a = np.sort(gpu.debug(pair_idx, _abi.DBG_EDGE_KEYS))
        b = np.sort(oracle.debug(pair_idx, _abi.DBG_EDGE_KEYS))
        if a.shape != b.shape or not (a == b).all():
            bad.append("edge_keys (sorted) differ")
    assert not bad, "; ".join(bad)


def compare_pose(Rg, tg, ig, Ro, to, io):
    assert ig == io
    ang, dt = synth.pose_error(Rg, tg, np.asarray(Ro, np.float64), np.asarray(to, np.float64))
    assert ang < TOL_ANG and dt < TOL_T, (ang, dt)


def run_both(gpu, oracle, p, **kw):
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier, **kw)
    out_g = gpu.register(p.src, p.dst)
    out_o = oracle.register(p.src, p.dst)
    compare_stages(gpu, oracle)
    compare_pose(*out_g, *out_o)
    return out_g


@pytest.mark.parametrize("N", [3, 4, 31, 32, 33, 100, 127, 128, 129, 255, 256, 257, 300, 1000, 1025])
def test_stage_parity_small_and_ragged_sizes(gpu, oracle, N):
    p = synth.make_pair(N, 0.3, 7000 + N)
    run_both(gpu, oracle, p, num_edges=64, apex_per_edge=4)


@pytest.mark.parametrize("Ke,m", [(1, 1), (5, 8), (1024, 4), (4096, 8)])
def test_stage_parity_selection_sizes(gpu, oracle, Ke, m):
    p = synth.make_pair(700, 0.1, 7100 + Ke)
    run_both(gpu, oracle, p, num_edges=Ke, apex_per_edge=m)


def test_cfg1_single_pair_n1000(gpu, oracle):
    p = synth.make_config_pair("cfg1_single_n1000", 0)
    R, t, inl = run_both(gpu, oracle, p)
    ang, dt = synth.pose_error(R, t, p.R_gt, p.t_gt)
    assert ang < np.deg2rad(1.0) and dt < 0.02


@pytest.mark.parametrize("cfg,b", [("cfg2_3dmatch_256x5000", 0), ("cfg2_3dmatch_256x5000", 1),
                                   ("cfg3_3dlomatch_256x5000", 0), ("cfg3_3dlomatch_256x5000", 1)])
def test_n5000_pairs_full_stage_parity(gpu, oracle, cfg, b):
    p = synth.make_config_pair(cfg, b)
    R, t, inl = run_both(gpu, oracle, p)
    ang, dt = synth.pose_error(R, t, p.R_gt, p.t_gt)
    assert ang < np.deg2rad(2.0) and dt < 0.05
    assert inl >= len(p.inlier_idx)


def test_kitti_scale_pair_n10000(gpu, oracle):
    p = synth.make_config_pair("cfg4_kitti_128x10000", 0)
    run_both(gpu, oracle, p)


def test_chunked_row_path_n13000(gpu, oracle):
    # stride 408 words > 352: the triangle kernel runs its 256-word chunk / accumulate mode
    p = synth.make_pair(13000, 0.03, 7300)
    run_both(gpu, oracle, p)


def test_truncated_residual_mode(gpu, oracle):
    p = synth.make_pair(1500, 0.08, 7400)
    run_both(gpu, oracle, p, score_mode=1, num_edges=256)


def test_refit_off_returns_winner_bits(gpu, oracle):
    p = synth.make_pair(900, 0.1, 7500)
    for r in (gpu, oracle):
        set_params(r, refit=0)
    Rg, tg, ig = gpu.register(p.src, p.dst)
    Ro, to, io = oracle.register(p.src, p.dst)
    np.testing.assert_array_equal(Rg, Ro)
    np.testing.assert_array_equal(tg, to)
    assert ig == io


def test_complete_graph_and_exact_pose(gpu, oracle):
    # noise-free all-inlier input: complete graph, every T equal -> the tie bucket holds all E keys
    # (in-place radix select path), pose exact
    N = 400
    p = synth.make_pair(N, 1.0, 7600)
    dst = (p.src.astype(np.float64) @ p.R_gt.T + p.t_gt).astype(np.float32)
    for r in (gpu, oracle):
        set_params(r, num_edges=512, apex_per_edge=4)
    out_g = gpu.register(p.src, dst)
    out_o = oracle.register(p.src, dst)
    compare_stages(gpu, oracle)
    compare_pose(*out_g, *out_o)
    assert out_g[2] == N
    assert int(gpu.debug(0, _abi.DBG_NUM_EDGES)[0]) == N * (N - 1) // 2


def test_empty_graph_is_a_result(gpu, oracle):
    rng = np.random.default_rng(0)
    src = (rng.random((64, 3)) * 100).astype(np.float32)
    dst = (rng.random((64, 3)) * 100).astype(np.float32)
    for r in (gpu, oracle):
        set_params(r, tau_compat=1e-5)
    out_g = gpu.register(src, dst)
    out_o = oracle.register(src, dst)
    compare_stages(gpu, oracle)
    np.testing.assert_array_equal(out_g[0], out_o[0])
    np.testing.assert_array_equal(out_g[1], out_o[1])
    assert out_g[2] == out_o[2]


def test_fewer_edges_than_requested(gpu, oracle):
    # tiny graph: E < K_e, triangles missing for most slots
    p = synth.make_pair(12, 0.5, 7700)
    run_both(gpu, oracle, p, num_edges=256, apex_per_edge=8)


def test_batch_with_ragged_sizes_matches_oracle(gpu, oracle):
    sizes = (300, 1000, 129, 2048, 64, 777)
    pairs = [synth.make_pair(n, 0.1, 7800 + k) for k, n in enumerate(sizes)]
    rg = gpu.register_batch([p.src for p in pairs], [p.dst for p in pairs])
    ro = oracle.register_batch([p.src for p in pairs], [p.dst for p in pairs])
    for b in range(len(pairs)):
        compare_stages(gpu, oracle, b)
        compare_pose(rg.R[b], rg.t[b], rg.inliers[b], ro.R[b], ro.t[b], ro.inliers[b])
    # pointer-array entry point and chunked execution give the same bits
    rp = gpu.register_pointer_batch([p.src for p in pairs], [p.dst for p in pairs])
    np.testing.assert_array_equal(rp.R, rg.R)
    np.testing.assert_array_equal(rp.t, rg.t)
    np.testing.assert_array_equal(rp.inliers, rg.inliers)
    gpu.set("keep_debug", 0)
    gpu.set("chunk_pairs", 2)
    rc = gpu.register_batch([p.src for p in pairs], [p.dst for p in pairs])
    np.testing.assert_array_equal(rc.R, rg.R)
    np.testing.assert_array_equal(rc.t, rg.t)
    np.testing.assert_array_equal(rc.inliers, rg.inliers)


@pytest.mark.parametrize("path", [0, 2])
def test_key_pool_growth_retry(gpu_lib, oracle, path):
    # a dense graph (large tau) exceeds the 12.5 % initial key-pool guess: the library must grow
    # the pool and re-run transparently
    p = synth.make_pair(1200, 0.2, 7900)
    with Registrar(lib=gpu_lib, tau_compat=1.5, tau_inlier=0.1) as g:
        g.set("keep_debug", 1)
        g.set("triangle_path", path)
        set_params(oracle, tau_compat=1.5, tau_inlier=0.1)
        out_g = g.register(p.src, p.dst)
        out_o = oracle.register(p.src, p.dst)
        assert g.get("retries") >= 1
        assert g.get("triangle_path_used") == (1 if path else 0)   # dense graph: the automatic choice is the tensor cores
        (compare_pruned if path else compare_stages)(g, oracle)
        compare_pose(*out_g, *out_o)


def test_device_resident_entry_point(gpu_lib, oracle):
    import torch
    pairs = [synth.make_pair(1000, 0.1, 8000 + k) for k in range(5)]
    src = np.concatenate([p.src for p in pairs])
    dst = np.concatenate([p.dst for p in pairs])
    offsets = np.arange(6, dtype=np.int64) * 1000
    dev = torch.device("cuda", 0)
    d_src, d_dst = torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev)
    d_R = torch.empty((5, 3, 3), dtype=torch.float32, device=dev)
    d_t = torch.empty((5, 3), dtype=torch.float32, device=dev)
    d_i = torch.empty(5, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)
    with Registrar(lib=gpu_lib, device=0, stream=stream.cuda_stream) as g:
        g.set("triangle_path", 0)
        g.set("lanes", 2)
        g.set("chunk_pairs", 3)   # (five pairs this small would otherwise run as one chunk)
        g.register_packed_ptr(d_src.data_ptr(), d_dst.data_ptr(), offsets, d_R.data_ptr(), d_t.data_ptr(),
                              d_i.data_ptr(), _abi.LOC_DEVICE)
        stream.synchronize()
        assert g.get("last_status") == 0
        assert g.get("launches") == 22   # 2 chunks (one per lane) x 11 kernels
    ro = oracle.register_batch([p.src for p in pairs], [p.dst for p in pairs])
    for b in range(5):
        compare_pose(d_R[b].cpu().numpy(), d_t[b].cpu().numpy(), int(d_i[b]), ro.R[b], ro.t[b], ro.inliers[b])


@pytest.mark.parametrize("path", [0, 2])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_phases_match_unsharded_and_oracle(gpu_lib, oracle_lib, world, path):
    # ranks emulated one after the other on the single GPU (no kernel waits on another).
    # path 0: POPC kernels; path 2: the library's choice, the tensor-core kernel at this size and density
    # (a rank then runs the 240-column tiles of its blocks of 1920 columns).
    N = 6000
    p = synth.make_pair(N, 0.05, 8100)
    with Registrar(lib=gpu_lib) as one:
        R1, t1, i1 = one.register(p.src, p.dst)
    ranks = [Registrar(lib=gpu_lib) for _ in range(world)]
    oranks = [Registrar(lib=oracle_lib) for _ in range(world)]
    try:
        for r in ranks:
            r.set("triangle_path", path)
        ph = [r.sharded_phase1(p.src, p.dst, g, world) for g, r in enumerate(ranks)]
        oph = [r.sharded_phase1(p.src, p.dst, g, world) for g, r in enumerate(oranks)]
        used = [r.get("triangle_path_used") for r in ranks]
        assert used == [1 if path else 0] * world   # decided by the density of the whole graph: the same on every rank
        # the global K_e-th key: what a rank's list must contain at least
        union = np.sort(np.concatenate([x[1] for x in oph]))[::-1]
        kth = union[len(ph[0][1]) - 1]
        for g in range(world):
            np.testing.assert_array_equal(ph[g][0], oph[g][0])   # partial node sums, bit-exact per rank
            if path == 0:
                np.testing.assert_array_equal(ph[g][1], oph[g][1])   # local top-K_e candidates
            else:
                # pruned: a rank keeps its edges at or above a threshold that at least K_e edges of the WHOLE graph
                # reach: its list is sorted, duplicate-free, and holds every local candidate that can be selected
                mine, ref = ph[g][1], oph[g][1]
                nz = mine[mine != 0]
                assert (np.diff(nz.astype(np.int64)) < 0).all() if len(nz) > 1 else True
                assert np.isin(ref[ref >= kth], nz).all()
        t_all = np.stack([x[0] for x in ph])
        c_all = np.stack([x[1] for x in ph])
        keys = [r.sharded_phase2(t_all, c_all) for r in ranks]
        okeys = [r.sharded_phase2(np.stack([x[0] for x in oph]), np.stack([x[1] for x in oph])) for r in oranks]
        assert keys == okeys   # merged selection, hypotheses and scores are the same with either candidate lists
        best = max(keys)
        for r in ranks:
            R, t, inl = r.sharded_phase3(best)
            np.testing.assert_array_equal(R, R1)   # identical to the unsharded GPU result, bit for bit
            np.testing.assert_array_equal(t, t1)
            assert inl == i1
    finally:
        for r in ranks + oranks:
            r.close()


def test_argument_checking_matches_oracle(gpu_lib, oracle_lib):
    src = np.zeros((10, 3), np.float32)
    R = np.zeros(9, np.float32)
    t = np.zeros(3, np.float32)
    inl = C.c_int32()
    f = _abi.fptr
    for lib in (gpu_lib, oracle_lib):
        p = _abi.default_params(lib)
        assert lib.sac_cot_register(None, f(src), 10, C.byref(p), f(R), f(t), C.byref(inl)) == _abi.E_NULL
        assert lib.sac_cot_register(f(src), f(src), 2, C.byref(p), f(R), f(t), C.byref(inl)) == _abi.E_SIZE
        assert lib.sac_cot_register(f(src), f(src), 70000, C.byref(p), f(R), f(t), C.byref(inl)) == _abi.E_SIZE
        for field, bad in [("struct_size", 4), ("tau_compat", 0.0), ("num_edges", 0), ("num_edges", 5000),
                           ("apex_per_edge", 9), ("score_mode", 2), ("refit", 2), ("reserved", 1)]:
            q = _abi.default_params(lib, **{field: bad})
            assert lib.sac_cot_register(f(src), f(src), 10, C.byref(q), f(R), f(t), C.byref(inl)) == _abi.E_PARAMS
    with Registrar(lib=gpu_lib) as g:
        with pytest.raises(SacCotError):
            g.set("no_such_knob", 1)
        with pytest.raises(SacCotError):
            g.debug(0, _abi.DBG_ADJ)  # nothing resident yet


def test_module_level_register_uses_cuda(gpu_lib):
    import sac_cot_b200
    p = synth.make_pair(800, 0.1, 8200)
    R, t, inl = sac_cot_b200.register(p.src, p.dst)
    ang, dt = synth.pose_error(R, t, p.R_gt, p.t_gt)
    assert ang < np.deg2rad(1.0) and dt < 0.02 and inl >= 75


def test_run_to_run_determinism(gpu):
    p = synth.make_pair(3000, 0.05, 8300)
    outs = []
    for _ in range(3):
        R, t, inl = gpu.register(p.src, p.dst)
        outs.append((R.copy(), t.copy(), inl, gpu.debug(0, _abi.DBG_TOP_EDGES), gpu.debug(0, _abi.DBG_HYP_SCORE)))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            np.testing.assert_array_equal(a, b)


# ---- adversarial inputs for the graph kernel's exact sqrt-free filter ------------------------
def _adj_parity(gpu, oracle, src, dst, tau):
    for r in (gpu, oracle):
        set_params(r, tau_compat=tau, tau_inlier=tau, num_edges=16, apex_per_edge=2)
    gpu.register(src, dst)
    oracle.register(src, dst)
    compare_stages(gpu, oracle, stages=[("adj", _abi.DBG_ADJ), ("num_edges", _abi.DBG_NUM_EDGES),
                                        ("t_node", _abi.DBG_T_NODE)])
    return int(gpu.debug(0, _abi.DBG_NUM_EDGES)[0])


def test_graph_near_threshold_scaled_cloud(gpu, oracle):
    # dst = rotated (1+eps) * src: |ld - ls| = eps * ls, so every pair with ls ~ tau/eps sits on the
    # decision boundary; thousands of pairs land inside the filter's rounding band
    rng = np.random.default_rng(5)
    N = 1500
    src = (rng.random((N, 3)) * 4 - 2).astype(np.float32)
    p = synth.make_pair(8, 0.5, 1)  # only for a rotation
    for eps in (1.0 / 64, 0.03, 0.1):
        dst = ((1 + eps) * src.astype(np.float64) @ p.R_gt.T).astype(np.float32)
        tau = float(np.float32(eps * 2.0))
        E = _adj_parity(gpu, oracle, src, dst, tau)
        assert 0 < E < N * (N - 1) // 2


def test_graph_short_lengths_around_the_filters_lower_limit(gpu, oracle):
    # The filter's threshold Theta = U*U * 48u (U = x + y - tau^2) is smallest relative to (x + y)^2 where x + y is just
    # above its lower limit 4 tau^2, i.e. for pairs a few tau apart in both clouds.  Tight clusters (radius ~ 1.5 tau,
    # independent offsets in src and dst) put most intra-cluster pairs there, with |ls - ld| on both sides of tau; pairs
    # of different clusters are ordinary.  The same cloud scaled so that tau sits at 2^-9 and 2^7 as well.
    rng = np.random.default_rng(21)
    N, C = 1536, 24
    for tau in (0.1, 2.0 ** -9, 128.0):
        centres = (rng.random((C, 3)) * 40 - 20) * tau
        which = rng.integers(0, C, N)
        src = (centres[which] + (rng.random((N, 3)) * 3 - 1.5) * tau).astype(np.float32)
        dst = (centres[which] + (rng.random((N, 3)) * 3 - 1.5) * tau).astype(np.float32)
        E = _adj_parity(gpu, oracle, src, dst, float(np.float32(tau)))
        assert 0 < E < N * (N - 1) // 2
        # a cloud only 4 tau wide and dst = rotated 1.7 x src: |ld - ls| = 0.7 ls, the decision boundary lies at
        # ls = 1.43 tau, where x + y = 7.9 tau^2
        src = ((rng.random((N, 3)) * 4 - 2) * tau).astype(np.float32)
        dst = (1.7 * src.astype(np.float64) @ synth.make_pair(8, 0.5, 1).R_gt.T).astype(np.float32)
        E = _adj_parity(gpu, oracle, src, dst, float(np.float32(tau)))
        assert 0 < E < N * (N - 1) // 2


def test_graph_exact_ties_on_a_lattice(gpu, oracle):
    # collinear integer lattice, dst stretched by 1 + 2^-6: ld - ls = |i-j| * 2^-6 exactly, so with
    # tau = k * 2^-6 every pair at lattice distance k is an exact tie (strict '<' => no edge)
    N = 700
    n = np.arange(N, dtype=np.float64)
    src = np.stack([n, np.zeros(N), np.zeros(N)], 1).astype(np.float32)
    dst = np.stack([n * (1 + 2.0 ** -6), np.zeros(N), np.zeros(N)], 1).astype(np.float32)
    for k in (1, 7, 64):
        E = _adj_parity(gpu, oracle, src, dst, float(k * 2.0 ** -6))
        A = np.abs(n[:, None] - n[None, :])
        assert E == int(((A < k) & (A > 0)).sum()) // 2


@pytest.mark.parametrize("scale", [1e-18, 1e-9, 1e-3, 1e6, 1e15])
def test_graph_extreme_magnitudes_and_duplicates(gpu, oracle, scale):
    rng = np.random.default_rng(9)
    N = 300
    src = (rng.random((N, 3)) * scale).astype(np.float32)
    dst = (rng.random((N, 3)) * scale).astype(np.float32)
    src[10:20] = src[0]          # duplicated points: zero lengths
    dst[15:40] = dst[1]
    dst[50:60] = 0.0
    _adj_parity(gpu, oracle, src, dst, float(np.float32(0.3 * scale)))
    _adj_parity(gpu, oracle, src, dst, float(np.float32(1e-3 * scale)))


def test_graph_huge_threshold_takes_literal_path(gpu, oracle):
    p = synth.make_pair(257, 0.2, 77)
    E = _adj_parity(gpu, oracle, p.src, p.dst, 1e30)   # 4*tau^2 overflows: every pair decided literally
    assert E == 257 * 256 // 2


def test_cfg5_single_pair_n50000_full_stage_parity(gpu_lib, oracle_omp_lib):
    # BASELINE.json configs[4]: 2.5e9-entry compatibility matrix, 9.4e7 edges, chunked-row triangle
    # kernel.  The OpenMP build of the oracle (thread-count independent by construction) keeps the
    # CPU side to a few seconds per core-minute.
    p = synth.make_config_pair("cfg5_single_n50000", 0)
    with Registrar(lib=gpu_lib) as g, Registrar(lib=oracle_omp_lib) as o:
        for r in (g, o):
            r.set("keep_debug", 1)
            set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
        out_o = o.register(p.src, p.dst)
        # the automatic choice at this density is the tensor-core kernel (390 tiles of 256 x 240 x 50176 MACs)
        out_g = g.register(p.src, p.dst)
        assert g.get("triangle_path_used") == 1
        compare_pruned(g, o, stages=[s for s in PRUNED if s[0] != "adj"])
        compare_pose(*out_g, *out_o)
        # POPC kernels (chunked rows): every key and the whole histogram
        g.set("triangle_path", 0)
        out_g = g.register(p.src, p.dst)
        compare_stages(g, o, stages=[s for s in EXACT if s[0] != "adj"])
        # adjacency: 313 MB per side; compare through a checksum of 64-bit words plus the counts above
        ag = g.debug(0, _abi.DBG_ADJ).view(np.uint64)
        ao = o.debug(0, _abi.DBG_ADJ).view(np.uint64)
        assert ag.shape == ao.shape and bool((ag == ao).all())
        compare_pose(*out_g, *out_o)
        ang, dt = synth.pose_error(out_g[0], out_g[1], p.R_gt, p.t_gt)
        assert ang < np.deg2rad(1.0) and dt < 0.02


def test_many_tiny_pairs_and_repeated_calls_on_one_ctx(gpu_lib, oracle_lib):
    # thousands of tiny pairs (several chunks over two lanes), then a larger call on the same ctx
    # (workspace regrowth), then the tiny batch again: results never depend on call history
    rng = np.random.default_rng(3)
    sizes = rng.integers(3, 40, size=3000)
    pairs = [synth.make_pair(int(n), 0.5, 9000 + k) for k, n in enumerate(sizes)]
    S, D = [p.src for p in pairs], [p.dst for p in pairs]
    with Registrar(lib=gpu_lib, num_edges=32, apex_per_edge=2) as g, \
            Registrar(lib=oracle_lib, num_edges=32, apex_per_edge=2) as o:
        g.set("chunk_pairs", 700)
        r1 = g.register_batch(S, D)
        ro = o.register_batch(S, D)
        np.testing.assert_array_equal(r1.inliers, ro.inliers)
        for b in range(0, 3000, 97):
            compare_pose(r1.R[b], r1.t[b], r1.inliers[b], ro.R[b], ro.t[b], ro.inliers[b])
        big = synth.make_pair(4000, 0.05, 9999)
        g.register(big.src, big.dst)
        r2 = g.register_batch(S, D)
        np.testing.assert_array_equal(r1.R, r2.R)
        np.testing.assert_array_equal(r1.t, r2.t)
        np.testing.assert_array_equal(r1.inliers, r2.inliers)


# ---- tensor-core triangle path (tcgen05 kind::mxf4, TMEM accumulators) ---------------------------
PRUNED = [st for st in EXACT if st[0] != "hist"]


def compare_pruned(gpu, oracle, pair_idx=0, stages=None):
    """With pruning the kernel keeps only edges whose T reaches a per-pair threshold that provably lies at
    or below the K_e-th largest T: every stage downstream is identical, the key list is a subset that
    contains every edge at or above the K_e-th key, and the histogram is exact above that key's digit."""
    compare_stages(gpu, oracle, pair_idx, stages=PRUNED if stages is None else stages, edge_keys=False)
    kg = np.sort(gpu.debug(pair_idx, _abi.DBG_EDGE_KEYS))
    ko = np.sort(oracle.debug(pair_idx, _abi.DBG_EDGE_KEYS))
    assert len(np.unique(kg)) == len(kg) and np.isin(kg, ko).all(), "pruned keys are not a subset of the oracle's"
    top = oracle.debug(pair_idx, _abi.DBG_TOP_EDGES)
    if len(top):
        t_min = int(top[-1] >> np.uint64(32))
        need = ko[(ko >> np.uint64(32)) >= np.uint64(t_min)]
        assert np.isin(need, kg).all(), "an edge at or above the K_e-th count was pruned"
        hg, ho = gpu.debug(pair_idx, _abi.DBG_HIST), oracle.debug(pair_idx, _abi.DBG_HIST)
        d = t_min >> 4
        assert (hg[d + 1:] == ho[d + 1:]).all(), "histogram differs above the threshold digit"
        assert int((need >> np.uint64(36) == np.uint64(d)).sum()) <= hg[d] <= ho[d]
    if len(ko) and len(kg) < len(ko):
        assert len(kg) >= min(len(ko), len(top))


@pytest.mark.parametrize("prune", [0, 1])
@pytest.mark.parametrize("N,ratio", [(3, 1.0), (64, 0.3), (129, 0.3), (241, 0.2), (500, 0.2), (1000, 0.1), (2048, 0.1),
                                     (5000, 0.05)])
def test_tensor_core_triangle_path_matches_oracle(gpu, oracle, N, ratio, prune):
    gpu.set("triangle_path", 1)
    gpu.set("triangle_prune", prune)
    assert gpu.get("triangle_path") == 1 and gpu.get("triangle_prune") == prune
    p = synth.make_pair(N, ratio, 8800 + N)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier, num_edges=256, apex_per_edge=4)
    out_g = gpu.register(p.src, p.dst)
    out_o = oracle.register(p.src, p.dst)
    if prune:
        compare_pruned(gpu, oracle)
    else:
        compare_stages(gpu, oracle)
    compare_pose(*out_g, *out_o)


def test_tensor_core_path_prunes_most_edges_at_headline_scale(gpu, oracle):
    gpu.set("triangle_path", 1)
    p = synth.make_config_pair("cfg2_3dmatch_256x5000", 3)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
    out_g = gpu.register(p.src, p.dst)
    out_o = oracle.register(p.src, p.dst)
    compare_pruned(gpu, oracle)
    compare_pose(*out_g, *out_o)
    kept = len(gpu.debug(0, _abi.DBG_EDGE_KEYS))
    E = int(gpu.debug(0, _abi.DBG_NUM_EDGES)[0])
    assert kept >= 1024 and kept < E // 4, (kept, E)


def test_tensor_core_path_no_inliers_keeps_selection_exact(gpu, oracle):
    # no clique: the sample of high-degree nodes holds few edges, the threshold falls back towards 0
    gpu.set("triangle_path", 1)
    p = synth.make_pair(3000, 0.0, 8870)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier, num_edges=1024, apex_per_edge=2)
    gpu.register(p.src, p.dst)
    oracle.register(p.src, p.dst)
    compare_pruned(gpu, oracle)


@pytest.mark.parametrize("prune", [0, 1])
def test_tensor_core_path_batch_and_complete_graph(gpu, oracle, prune):
    gpu.set("triangle_path", 1)
    gpu.set("triangle_prune", prune)
    check = compare_pruned if prune else compare_stages
    # complete graph: every T equals N-2 (exact integers out of the fp32 accumulators)
    N = 300
    p = synth.make_pair(N, 1.0, 8900)
    dst = (p.src.astype(np.float64) @ p.R_gt.T + p.t_gt).astype(np.float32)
    for r in (gpu, oracle):
        set_params(r, num_edges=128, apex_per_edge=4)
    out_g = gpu.register(p.src, dst)
    out_o = oracle.register(p.src, dst)
    check(gpu, oracle)
    compare_pose(*out_g, *out_o)
    # ragged batch
    pairs = [synth.make_pair(n, 0.1, 8950 + k) for k, n in enumerate((300, 1000, 129, 2048, 64, 777))]
    rg = gpu.register_batch([q.src for q in pairs], [q.dst for q in pairs])
    ro = oracle.register_batch([q.src for q in pairs], [q.dst for q in pairs])
    for b in range(len(pairs)):
        check(gpu, oracle, b)
        compare_pose(rg.R[b], rg.t[b], rg.inliers[b], ro.R[b], ro.t[b], ro.inliers[b])


# ---- automatic choice of the S2 kernels from the measured edge density ---------------------------
def test_auto_path_picks_tensor_cores_for_dense_graphs_and_popc_for_sparse(gpu, oracle):
    gpu.set("triangle_path", 2)
    # indoor scale: ~7.6 % density at N = 5000 -> tensor cores
    p = synth.make_config_pair("cfg2_3dmatch_256x5000", 5)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
    out_g, out_o = gpu.register(p.src, p.dst), oracle.register(p.src, p.dst)
    assert gpu.get("triangle_path_used") == 1
    compare_pruned(gpu, oracle)
    compare_pose(*out_g, *out_o)
    # outdoor scale: ~2.3 % density -> still the tensor cores (measured at KITTI scale: 21.6 ms against 24.4 ms per
    # 128-pair step for the POPC kernels, whose time falls with the density while the dense kernel's does not)
    p = synth.make_pair(4000, 0.03, 9911, box=(60.0, 60.0, 6.0), tau_compat=0.6)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
    out_g, out_o = gpu.register(p.src, p.dst), oracle.register(p.src, p.dst)
    assert gpu.get("triangle_path_used") == 1
    compare_pruned(gpu, oracle)
    compare_pose(*out_g, *out_o)
    # a sparser graph (tighter threshold in the same scene, well under 2 % density) -> POPC kernels, every key kept
    p = synth.make_pair(4000, 0.03, 9911, box=(60.0, 60.0, 6.0), tau_compat=0.2)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
    out_g, out_o = gpu.register(p.src, p.dst), oracle.register(p.src, p.dst)
    assert int(gpu.debug(0, _abi.DBG_NUM_EDGES)[0]) < 0.02 * 4000 * 3999 / 2
    assert gpu.get("triangle_path_used") == 0
    compare_stages(gpu, oracle)
    compare_pose(*out_g, *out_o)
    # small pairs stay on the POPC kernels whatever their density
    p = synth.make_pair(500, 0.5, 9912)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
    gpu.register(p.src, p.dst), oracle.register(p.src, p.dst)
    assert gpu.get("triangle_path_used") == 0
    compare_stages(gpu, oracle)


def test_default_ctx_uses_auto_path_and_matches_oracle_batch(gpu_lib, oracle):
    ps = [synth.make_config_pair("cfg2_3dmatch_256x5000", b) for b in range(3)]
    with Registrar(lib=gpu_lib, device=0) as reg:
        assert reg.get("triangle_path") == 2
        set_params(oracle, tau_compat=ps[0].tau_compat, tau_inlier=ps[0].tau_inlier)
        set_params(reg, tau_compat=ps[0].tau_compat, tau_inlier=ps[0].tau_inlier)
        rg = reg.register_batch([q.src for q in ps], [q.dst for q in ps])
        ro = oracle.register_batch([q.src for q in ps], [q.dst for q in ps])
        assert reg.get("triangle_path_used") == 1
        for b in range(len(ps)):
            compare_pose(rg.R[b], rg.t[b], rg.inliers[b], ro.R[b], ro.t[b], ro.inliers[b])


# ---------------------------------------------------------------------------------------------
# apex selection: rank list, exhaustive scan and the global-lookup fallback give the same triangles
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("apex_path", [0, 1, 2])
@pytest.mark.parametrize("N,ratio,m", [(700, 0.1, 4), (5000, 0.05, 4), (5000, 0.01, 8), (2000, 0.0, 2)])
def test_apex_paths_match_oracle(gpu, oracle, apex_path, N, ratio, m):
    # 1 % / 0 % inliers: most selected edges have fewer than m common neighbours among the ~256 best-ranked
    # nodes, so the rank-list kernel takes its exhaustive path for them
    gpu.set("apex_path", apex_path)
    p = synth.make_pair(N, ratio, 9100 + N + m)
    run_both(gpu, oracle, p, num_edges=512, apex_per_edge=m)


@pytest.mark.parametrize("apex_path", [0, 1, 2])
def test_apex_paths_on_ties(gpu, oracle, apex_path):
    # complete graph: every node count is equal, the rank list is cut inside one huge tie (ties go to the lowest k)
    gpu.set("apex_path", apex_path)
    N = 600
    rng = np.random.default_rng(77)
    src = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
    for r in (gpu, oracle):
        set_params(r, tau_compat=0.5, tau_inlier=0.1, num_edges=300, apex_per_edge=8)
    out_g = gpu.register(src, src.copy())
    out_o = oracle.register(src, src.copy())
    compare_stages(gpu, oracle)
    compare_pose(*out_g, *out_o)


def test_apex_path_rejects_unknown_value(gpu):
    from sac_cot_b200.api import SacCotError
    with pytest.raises(SacCotError):
        gpu.set("apex_path", 3)


# ---------------------------------------------------------------------------------------------
# tensor-core path: the tile schedule (runs vs one at a time) does not change anything
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tile_runs", [0, 1])
def test_tensor_core_tile_schedule_is_invisible(gpu_lib, oracle_lib, tile_runs):
    from sac_cot_b200.api import Registrar
    # enough tiles for several runs per CTA pair, ragged sizes so that runs straddle pairs of different K
    sizes = [1500, 2600, 1030, 3100, 2048, 1800, 2900, 1200, 2500, 1700, 3000, 1100]
    ps = [synth.make_pair(n, 0.08, 9300 + k) for k, n in enumerate(sizes)]
    with Registrar(lib=gpu_lib, device=0) as g, Registrar(lib=oracle_lib) as o:
        for r in (g, o):
            r.set("keep_debug", 1)
            set_params(r, tau_compat=ps[0].tau_compat, tau_inlier=ps[0].tau_inlier, num_edges=256, apex_per_edge=4)
        g.set("triangle_path", 1)
        g.set("tile_runs", tile_runs)
        rg = g.register_batch([p.src for p in ps], [p.dst for p in ps])
        ro = o.register_batch([p.src for p in ps], [p.dst for p in ps])
        assert g.get("triangle_path_used") == 1
        assert (rg.inliers == ro.inliers).all()
        for b in range(len(ps)):
            compare_pruned(g, o, pair_idx=b)


# ---------------------------------------------------------------------------------------------
# Sharded single pair with the collectives inside the library (sac_cot_register_sharded)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path,N", [(0, 3000), (2, 6000)])
def test_register_sharded_in_library_world1_matches_unsharded(gpu_lib, oracle, path, N):
    # a one-rank communicator on the single GPU: pack -> ncclAllGather -> merge -> ... -> ncclAllReduce(max) all run
    p = synth.make_pair(N, 0.05, 8400)
    ident = (C.c_ubyte * _abi.COMM_ID_BYTES)()
    assert gpu_lib.sac_cot_comm_unique_id(ident) == _abi.OK
    with Registrar(lib=gpu_lib) as one, Registrar(lib=gpu_lib) as sh:
        for r in (one, sh):
            r.set("triangle_path", path)
            r.set("keep_debug", 1)
        with pytest.raises(SacCotError) as ei:   # no communicator yet
            sh.register_sharded_ptr(p.src.ctypes.data, p.dst.ctypes.data, N, 0, 0, 0, _abi.LOC_HOST)
        assert ei.value.status in (_abi.E_NULL, _abi.E_COMM)
        sh.comm_init(rank=0, world=1, unique_id=bytes(ident))
        assert sh.get("comm_world") == 1 and sh.get("comm_rank") == 0
        R1, t1, i1 = one.register(p.src, p.dst)
        launches0 = sh.get("launches")
        R, t, inl = sh.register_sharded(p.src, p.dst)
        np.testing.assert_array_equal(R, R1)
        np.testing.assert_array_equal(t, t1)
        assert inl == i1
        assert sh.get("launches") - launches0 >= 12   # the pipeline plus the record and merge kernels
        for name, which in [("t_node", _abi.DBG_T_NODE), ("top_edges", _abi.DBG_TOP_EDGES), ("triangles", _abi.DBG_TRIANGLES),
                            ("hyp_score", _abi.DBG_HYP_SCORE), ("best_key", _abi.DBG_BEST_KEY), ("mask", _abi.DBG_MASK)]:
            np.testing.assert_array_equal(sh.debug(0, which), one.debug(0, which), err_msg=name)
        set_params(oracle, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
        Ro, to, io = oracle.register(p.src, p.dst)
        compare_pose(R, t, inl, Ro, to, io)
        # a second call re-uses the resident plan (no descriptor upload) and gives the same bits
        R2, t2, i2 = sh.register_sharded(p.src, p.dst)
        np.testing.assert_array_equal(R2, R1)
        assert i2 == i1


def test_register_sharded_in_library_grows_key_pool_in_step(gpu_lib):
    # dense graph: the first attempt overflows the 12.5 % key-pool guess; the merged summary tells every rank to
    # grow and re-run
    p = synth.make_pair(1500, 0.2, 8500)
    ident = (C.c_ubyte * _abi.COMM_ID_BYTES)()
    assert gpu_lib.sac_cot_comm_unique_id(ident) == _abi.OK
    with Registrar(lib=gpu_lib, tau_compat=1.5) as one, Registrar(lib=gpu_lib, tau_compat=1.5) as sh:
        sh.comm_init(rank=0, world=1, unique_id=bytes(ident))
        R1, t1, i1 = one.register(p.src, p.dst)
        R, t, inl = sh.register_sharded(p.src, p.dst)
        assert sh.get("retries") >= 1
        np.testing.assert_array_equal(R, R1)
        np.testing.assert_array_equal(t, t1)
        assert inl == i1


def _sharded_worker(rank, world, port, N, ratio, seed, path, out_dir):
    import os
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # carries the 128-byte NCCL id only
    p = synth.make_pair(N, ratio, seed)
    with Registrar(device=rank) as reg:
        reg.set("triangle_path", path)
        reg.comm_init()
        R, t, inl = reg.register_sharded(p.src, p.dst)
        used = reg.get("triangle_path_used")
        # device-resident form: enqueue only, outputs in device buffers
        dev = torch.device("cuda", rank)
        d_src, d_dst = torch.from_numpy(p.src).to(dev), torch.from_numpy(p.dst).to(dev)
        d_R = torch.zeros(9, dtype=torch.float32, device=dev)
        d_t = torch.zeros(3, dtype=torch.float32, device=dev)
        d_i = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        reg.register_sharded_ptr(d_src.data_ptr(), d_dst.data_ptr(), N, d_R.data_ptr(), d_t.data_ptr(), d_i.data_ptr(),
                                 _abi.LOC_DEVICE)
        status = reg.get("last_status")   # synchronises
        R1, t1, i1 = reg.register(p.src, p.dst)   # unsharded, same GPU
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), R=R, t=t, inl=inl, R1=R1, t1=t1, i1=i1, used=used, status=status,
             dR=d_R.cpu().numpy().reshape(3, 3), dt=d_t.cpu().numpy(), di=int(d_i.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("path,N,ratio", [(0, 4000, 0.05), (2, 9000, 0.05)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_register_sharded_in_library_multi_gpu(gpu_lib, tmp_path, world, path, N, ratio):
    # one process per GPU, NCCL inside the library; needs `world` GPUs (gpurun --gpus N)
    import socket

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_sharded_worker, args=(world, port, N, ratio, 8600 + world, path, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        np.testing.assert_array_equal(got["R"], got["R1"])   # every rank: bit-identical to the unsharded result
        np.testing.assert_array_equal(got["t"], got["t1"])
        assert int(got["inl"]) == int(got["i1"])
        assert int(got["status"]) == 0
        np.testing.assert_array_equal(got["dR"], got["R1"])
        np.testing.assert_array_equal(got["dt"], got["t1"])
        assert int(got["di"]) == int(got["i1"])
        assert int(got["used"]) == (1 if path else 0)


def test_overflowing_device_call_leaves_defined_outputs_and_reports(gpu_lib):
    # a key-pool overflow must stop every later kernel of the chunk.  The arena is poisoned by a larger,
    # different batch first; the overflowing device-location call must then neither fault nor write garbage:
    # its outputs read "no result" (R = I, t = 0, inliers = 0), last_status reports the overflow once, and the next
    # call (pool grown) is correct.
    import torch
    dev = torch.device("cuda", 0)
    # three N=700 pairs: a larger arena than the single N=900 pair needs (so it is re-used, stale contents and
    # all) but a smaller key pool (12.5 % guess: 3 x 30 k keys) than its near-complete graph fills (~400 k edges)
    big = [synth.make_pair(700, 0.1, 8700 + k) for k in range(3)]
    p = synth.make_pair(900, 0.3, 8710)
    stream = torch.cuda.current_stream(dev)
    with Registrar(lib=gpu_lib, device=0, stream=stream.cuda_stream) as g, Registrar(lib=gpu_lib) as ref:
        g.set("lanes", 1)
        g.register_batch([q.src for q in big], [q.dst for q in big])      # fills the arena with other pairs' data
        set_params(g, tau_compat=2.0)
        set_params(ref, tau_compat=2.0)
        d_src, d_dst = torch.from_numpy(p.src).to(dev), torch.from_numpy(p.dst).to(dev)
        d_R = torch.full((9,), 7.0, dtype=torch.float32, device=dev)
        d_t = torch.full((3,), 7.0, dtype=torch.float32, device=dev)
        d_i = torch.full((1,), 7, dtype=torch.int32, device=dev)
        offsets = np.array([0, 900], dtype=np.int64)
        g.register_packed_ptr(d_src.data_ptr(), d_dst.data_ptr(), offsets, d_R.data_ptr(), d_t.data_ptr(), d_i.data_ptr(),
                              _abi.LOC_DEVICE)
        assert g.get("last_status") == _abi.E_NOMEM
        assert g.get("last_status") == 0                                   # reported once
        np.testing.assert_array_equal(d_R.cpu().numpy(), np.eye(3, dtype=np.float32).ravel())
        np.testing.assert_array_equal(d_t.cpu().numpy(), np.zeros(3, np.float32))
        assert int(d_i.item()) == 0
        g.register_packed_ptr(d_src.data_ptr(), d_dst.data_ptr(), offsets, d_R.data_ptr(), d_t.data_ptr(), d_i.data_ptr(),
                              _abi.LOC_DEVICE)
        assert g.get("last_status") == 0
        R1, t1, i1 = ref.register(p.src, p.dst)
        np.testing.assert_array_equal(d_R.cpu().numpy().reshape(3, 3), R1)
        assert int(d_i.item()) == i1


# ---------------------------------------------------------------------------------------------
# Device groups: one batch dealt round-robin over the member contexts
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sizes", [(1000,) * 7, (300, 1000, 129, 2048, 64, 777, 1500)])
def test_group_round_robin_matches_single_ctx(gpu_lib, sizes):
    # members may share a device: two contexts on GPU 0 exercise the strided gather / scatter (uniform sizes: one 2-D
    # copy per side; ragged sizes: one copy per pair) on a one-GPU box; with more GPUs the members spread out
    import torch
    from sac_cot_b200.api import Group
    ndev = torch.cuda.device_count()
    pairs = [synth.make_pair(n, 0.1, 8800 + k) for k, n in enumerate(sizes)]
    with Registrar(lib=gpu_lib) as one:
        ref = one.register_batch([p.src for p in pairs], [p.dst for p in pairs])
    for devices in ([0], [0, 0], [0, 0, 0], list(range(min(ndev, 4))) if ndev > 1 else [0, 0, 0, 0]):
        with Group(devices, lib=gpu_lib) as grp:
            assert len(grp) == len(devices)
            grp.set("chunk_pairs", 2)
            got = grp.register_batch([p.src for p in pairs], [p.dst for p in pairs])
            np.testing.assert_array_equal(got.R, ref.R)
            np.testing.assert_array_equal(got.t, ref.t)
            np.testing.assert_array_equal(got.inliers, ref.inliers)
            assert sum(grp.get(g, "launches") for g in range(len(devices))) > 0
            # fewer pairs than members: some members get nothing
            few = grp.register_batch([p.src for p in pairs[:2]], [p.dst for p in pairs[:2]])
            np.testing.assert_array_equal(few.R, ref.R[:2])
            np.testing.assert_array_equal(few.inliers, ref.inliers[:2])


def test_group_rejects_bad_arguments(gpu_lib):
    from sac_cot_b200.api import Group
    with pytest.raises(SacCotError):
        Group([], lib=gpu_lib)
    with pytest.raises(SacCotError):
        Group([99], lib=gpu_lib)


# ---------------------------------------------------------------------------------------------
# Tensor-native front (sac_cot_b200.torch_api): device and stream inferred from the inputs
# ---------------------------------------------------------------------------------------------
def test_torch_api_matches_registrar(gpu_lib):
    import torch
    from sac_cot_b200 import torch_api
    pairs = [synth.make_pair(1500, 0.1, 8900 + k) for k in range(4)]
    with Registrar(lib=gpu_lib) as one:
        ref = one.register_batch([p.src for p in pairs], [p.dst for p in pairs])
    dev = torch.device("cuda", 0)
    src = torch.from_numpy(np.stack([p.src for p in pairs])).to(dev)
    dst = torch.from_numpy(np.stack([p.dst for p in pairs])).to(dev)
    try:
        # (B, N, 3) CUDA tensors on a side stream: enqueue only, results on the same device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            R, t, inl = torch_api.register(src, dst)
            assert R.is_cuda and R.shape == (4, 3, 3) and inl.dtype == torch.int32
            side.synchronize()
            assert torch_api.last_status(src) == 0
        np.testing.assert_array_equal(R.cpu().numpy(), ref.R)
        np.testing.assert_array_equal(t.cpu().numpy(), ref.t)
        np.testing.assert_array_equal(inl.cpu().numpy(), ref.inliers)
        # one pair, ragged list, CPU tensors
        R0, t0, i0 = torch_api.register(src[0], dst[0])
        torch.cuda.synchronize(dev)
        np.testing.assert_array_equal(R0.cpu().numpy(), ref.R[0])
        assert int(i0) == int(ref.inliers[0])
        Rr, tr, ir = torch_api.register([src[0], src[1][:1000]], [dst[0], dst[1][:1000]])
        torch.cuda.synchronize(dev)
        np.testing.assert_array_equal(Rr[0].cpu().numpy(), ref.R[0])
        Rc, tc, ic = torch_api.register(src.cpu(), dst.cpu())
        assert not Rc.is_cuda
        np.testing.assert_array_equal(Rc.numpy(), ref.R)
        np.testing.assert_array_equal(ic.numpy(), ref.inliers)
        with pytest.raises(ValueError):
            torch_api.register(src, dst[:, :100])
    finally:
        torch_api.release_contexts()


# ---------------------------------------------------------------------------------------------
# Second-order compatibility (params.compat_mode = SAC_COT_COMPAT_SECOND_ORDER)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", [0, 1, 2])
@pytest.mark.parametrize("N,ratio,cmin", [(64, 0.3, 3), (500, 0.2, 0), (1000, 0.1, 12), (2048, 0.1, 30), (5000, 0.05, 60),
                                          (700, 0.1, 60000)])
def test_second_order_mode_matches_oracle(gpu, oracle, path, N, ratio, cmin):
    gpu.set("triangle_path", path)
    p = synth.make_pair(N, ratio, 9000 + N)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier, num_edges=256, apex_per_edge=4,
                   compat_mode=_abi.COMPAT_SECOND_ORDER, so_min_common=cmin)
    out_g = gpu.register(p.src, p.dst)
    out_o = oracle.register(p.src, p.dst)
    np.testing.assert_array_equal(gpu.debug(0, _abi.DBG_ADJ_FIRST), oracle.debug(0, _abi.DBG_ADJ_FIRST))
    np.testing.assert_array_equal(gpu.debug(0, _abi.DBG_ADJ), oracle.debug(0, _abi.DBG_ADJ))   # A2
    if gpu.get("triangle_path_used") == 1:
        compare_pruned(gpu, oracle)
    else:
        compare_stages(gpu, oracle)
    compare_pose(*out_g, *out_o)
    if cmin == 0:
        np.testing.assert_array_equal(gpu.debug(0, _abi.DBG_ADJ), gpu.debug(0, _abi.DBG_ADJ_FIRST))


def test_second_order_mode_batch_and_pool_growth(gpu_lib, oracle):
    sizes = (300, 1000, 129, 2048, 777)
    pairs = [synth.make_pair(n, 0.15, 9100 + k) for k, n in enumerate(sizes)]
    with Registrar(lib=gpu_lib, compat_mode=_abi.COMPAT_SECOND_ORDER, so_min_common=8, tau_compat=0.6) as g:
        g.set("keep_debug", 1)
        set_params(oracle, compat_mode=_abi.COMPAT_SECOND_ORDER, so_min_common=8, tau_compat=0.6, tau_inlier=0.1)
        rg = g.register_batch([p.src for p in pairs], [p.dst for p in pairs])   # tau 0.6 in a 3 m box: dense, the pool grows
        ro = oracle.register_batch([p.src for p in pairs], [p.dst for p in pairs])
        assert g.get("retries") >= 1
        for b in range(len(pairs)):
            np.testing.assert_array_equal(g.debug(b, _abi.DBG_ADJ), oracle.debug(b, _abi.DBG_ADJ))
            np.testing.assert_array_equal(g.debug(b, _abi.DBG_T_NODE), oracle.debug(b, _abi.DBG_T_NODE))
            np.testing.assert_array_equal(g.debug(b, _abi.DBG_TOP_EDGES), oracle.debug(b, _abi.DBG_TOP_EDGES))
            np.testing.assert_array_equal(g.debug(b, _abi.DBG_HYP_SCORE), oracle.debug(b, _abi.DBG_HYP_SCORE))
            compare_pose(rg.R[b], rg.t[b], rg.inliers[b], ro.R[b], ro.t[b], ro.inliers[b])
        # chunked / multi-lane execution gives the same bits
        g.set("keep_debug", 0)
        g.set("chunk_pairs", 2)
        rc = g.register_batch([p.src for p in pairs], [p.dst for p in pairs])
        np.testing.assert_array_equal(rc.R, rg.R)
        np.testing.assert_array_equal(rc.inliers, rg.inliers)


def test_second_order_mode_rejected_where_unsupported_and_v1_struct_accepted(gpu_lib):
    p = synth.make_pair(600, 0.1, 9200)
    with Registrar(lib=gpu_lib, compat_mode=_abi.COMPAT_SECOND_ORDER, so_min_common=5) as g:
        with pytest.raises(SacCotError) as ei:
            g.sharded_phase1(p.src, p.dst, 0, 2)     # A2 needs every rank's counts: not available sharded
        assert ei.value.status == _abi.E_UNSUPPORTED
        g.params.so_min_common = -1
        with pytest.raises(SacCotError) as ei:
            g.register(p.src, p.dst)
        assert ei.value.status == _abi.E_PARAMS
    with Registrar(lib=gpu_lib) as g:
        R1, t1, i1 = g.register(p.src, p.dst)
        g.params.struct_size = _abi.PARAMS_SIZE_V1   # a caller compiled against version 1 of the struct
        g.params.so_min_common = 777                 # lies beyond it: ignored
        R2, t2, i2 = g.register(p.src, p.dst)
        np.testing.assert_array_equal(R1, R2)
        assert i1 == i2


# ---------------------------------------------------------------------------------------------
# exact node pruning of S2 (kernels_prune.cu): triangle counts for the rows of the high-degree nodes only
# ---------------------------------------------------------------------------------------------
NODE_PRUNED = [st for st in EXACT if st[0] not in ("hist", "t_node")]


def compare_node_pruned(gpu, oracle, pair_idx=0):
    """Everything downstream of S2 is identical; keys / histogram obey the same subset rules as with the threshold
    pruning; node counts are exact where they were computed (0 elsewhere) and computed for every node a selected
    edge touches."""
    compare_pruned(gpu, oracle, pair_idx, stages=NODE_PRUNED)
    tg, to = gpu.debug(pair_idx, _abi.DBG_T_NODE), oracle.debug(pair_idx, _abi.DBG_T_NODE)
    assert ((tg == to) | (tg == 0)).all(), "a computed node count differs from the oracle's"
    top = oracle.debug(pair_idx, _abi.DBG_TOP_EDGES)
    ends = np.unique(np.concatenate([0xFFFF - ((top >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64),
                                     0xFFFF - (top & np.uint64(0xFFFF)).astype(np.int64)]))
    assert (tg[ends] == to[ends]).all(), "an endpoint of a selected edge has no node count"


KITTI = dict(box=(60.0, 60.0, 6.0), tau_compat=0.6)   # ~2.3 % edge density: outlier degrees stay below the clique's counts


@pytest.mark.parametrize("rect", [1, 0])
@pytest.mark.parametrize("apex_path", [0, 1])
@pytest.mark.parametrize("N,ratio,Ke,m,kw", [
    (1200, 0.10, 256, 4, {}), (2048, 0.10, 256, 4, {}), (3000, 0.30, 512, 2, {}), (5000, 0.05, 1024, 4, KITTI),
    (5000, 0.05, 4096, 8, KITTI), (4999, 0.04, 64, 8, KITTI), (1500, 0.08, 1, 1, {}), (6000, 0.03, 1024, 4, KITTI),
    (10000, 0.03, 1024, 4, KITTI),
])
def test_node_pruned_path_matches_oracle(gpu, oracle, N, ratio, Ke, m, kw, apex_path, rect):
    """rect = 1: the kept rows of pairs with N >= 1921 (Npad >= 2048) and at most 1024 kept nodes go through the RECT
    instance of the tensor-core kernel, the others (and everything with rect = 0) through the kept-row POPC kernel."""
    if rect == 0 and apex_path == 1 and N > 3000:
        pytest.skip("the apex paths do not depend on which kernel counted the kept rows")
    gpu.set("triangle_path", 1)
    gpu.set("node_prune", 2)
    gpu.set("node_prune_rect", 2 * rect)   # 2: also for a single pair
    gpu.set("apex_path", apex_path)   # 1: no rank list, every edge evaluates its unknown candidates on demand
    assert gpu.get("node_prune") == 2 and gpu.get("node_prune_rect") == 2 * rect
    p = synth.make_pair(N, ratio, 9300 + N + Ke, **kw)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier, num_edges=Ke, apex_per_edge=m)
    out_g = gpu.register(p.src, p.dst)
    out_o = oracle.register(p.src, p.dst)
    assert gpu.get("pruned_pairs") == 1, "the pair was expected to take the kept-row kernel"
    assert 2 <= gpu.get("kept_nodes") <= 2048
    assert gpu.get("rect_pairs") == (1 if rect and N > 1920 and gpu.get("kept_nodes") <= 1024 else 0)
    compare_node_pruned(gpu, oracle)
    compare_pose(*out_g, *out_o)


def test_node_pruning_leaves_dense_outlier_graphs_alone(gpu, oracle):
    """Indoor scale (7.6 % density): the outliers' degrees exceed the clique's counts, every node would be kept."""
    gpu.set("triangle_path", 1)
    for mode in (1, 2):
        gpu.set("node_prune", mode)
        p = synth.make_config_pair("cfg2_3dmatch_256x5000", 11)
        for r in (gpu, oracle):
            set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
        out_g, out_o = gpu.register(p.src, p.dst), oracle.register(p.src, p.dst)
        assert gpu.get("pruned_pairs") == 0
        compare_pruned(gpu, oracle)
        compare_pose(*out_g, *out_o)


def test_node_pruning_mixed_batch_and_cost_model(gpu, oracle):
    """Pairs that prune and pairs that do not share a chunk: the tensor-core kernel runs the compacted tile list."""
    gpu.set("triangle_path", 1)
    gpu.set("node_prune", 2)
    specs = [(1500, 0.10), (1400, 0.0), (2048, 0.10), (1300, 0.9), (3000, 0.02), (1100, 0.0)]
    pairs = [synth.make_pair(n, r, 9400 + k) for k, (n, r) in enumerate(specs)]
    for r in (gpu, oracle):
        set_params(r, num_edges=256, apex_per_edge=4)
    rg = gpu.register_batch([q.src for q in pairs], [q.dst for q in pairs])
    ro = oracle.register_batch([q.src for q in pairs], [q.dst for q in pairs])
    npr = gpu.get("pruned_pairs")
    assert 1 <= npr < len(pairs), npr
    assert 1 <= gpu.get("rect_pairs") < npr   # the N = 2048 pair; the shorter ones take the kept-row POPC kernel
    for b in range(len(pairs)):
        compare_node_pruned(gpu, oracle, b)
        compare_pose(rg.R[b], rg.t[b], rg.inliers[b], ro.R[b], ro.t[b], ro.inliers[b])
    # the cost model (node_prune = 1 would apply it; keep_debug turns that mode off, so drive it through the knob):
    # an absurd cost prunes nothing, and the results do not move
    gpu.set("node_prune_cost", 1000000)
    gpu.set("node_prune", 2)
    rg2 = gpu.register_batch([q.src for q in pairs], [q.dst for q in pairs])
    np.testing.assert_array_equal(rg.R, rg2.R)
    np.testing.assert_array_equal(rg.inliers, rg2.inliers)


def test_node_pruning_is_the_default_and_invisible(gpu_lib, oracle):
    """Default ctx (no keep_debug): outdoor-scale pairs prune down to their inlier cliques; poses equal the oracle's and
    those of a ctx with the pruning switched off, bit for bit."""
    ps = [synth.make_pair(6000, 0.03, 9450 + b, **KITTI) for b in range(5)]   # >= 4 pairs: RECT instance for the kept rows
    set_params(oracle, tau_compat=ps[0].tau_compat, tau_inlier=ps[0].tau_inlier)
    ro = oracle.register_batch([q.src for q in ps], [q.dst for q in ps])
    outs = []
    for mode in (1, 0):
        with Registrar(lib=gpu_lib, device=0) as reg:
            assert reg.get("node_prune") == 1
            reg.set("node_prune", mode)
            set_params(reg, tau_compat=ps[0].tau_compat, tau_inlier=ps[0].tau_inlier)
            rg = reg.register_batch([q.src for q in ps], [q.dst for q in ps])
            assert reg.get("triangle_path_used") == 1
            assert reg.get("pruned_pairs") == (len(ps) if mode else 0)
            assert reg.get("rect_pairs") == (len(ps) if mode else 0)
            if mode:   # the inlier cliques and a few central (high-degree) outliers
                n_in = sum(len(q.inlier_idx) for q in ps)
                assert n_in <= reg.get("kept_nodes") <= 3 * n_in
            for b in range(len(ps)):
                compare_pose(rg.R[b], rg.t[b], rg.inliers[b], ro.R[b], ro.t[b], ro.inliers[b])
            outs.append(rg)
    np.testing.assert_array_equal(outs[0].R, outs[1].R)
    np.testing.assert_array_equal(outs[0].t, outs[1].t)
    np.testing.assert_array_equal(outs[0].inliers, outs[1].inliers)


def test_node_pruning_off_for_second_order_and_debug_dumps(gpu, oracle):
    gpu.set("triangle_path", 1)   # keep_debug is set by the fixture, node_prune is at its default (1)
    p = synth.make_pair(2048, 0.1, 9500)
    for r in (gpu, oracle):
        set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier, num_edges=256, apex_per_edge=4)
    gpu.register(p.src, p.dst)
    oracle.register(p.src, p.dst)
    assert gpu.get("pruned_pairs") == 0
    compare_pruned(gpu, oracle)   # full node counts
    gpu.set("node_prune", 2)
    for r in (gpu, oracle):
        set_params(r, compat_mode=_abi.COMPAT_SECOND_ORDER, so_min_common=10)
    out_g, out_o = gpu.register(p.src, p.dst), oracle.register(p.src, p.dst)
    assert gpu.get("pruned_pairs") == 0
    compare_pose(*out_g, *out_o)


def test_node_pruning_stays_out_of_the_sharded_phases(gpu_lib, oracle):
    """A pair that WOULD prune, through the in-library sharded entry point with a one-rank communicator and without
    keep_debug: the sharded phases rank apexes from complete node sums, so the pruning must stay off there."""
    p = synth.make_pair(6000, 0.03, 9470, **KITTI)
    ident = (C.c_ubyte * _abi.COMM_ID_BYTES)()
    assert gpu_lib.sac_cot_comm_unique_id(ident) == _abi.OK
    set_params(oracle, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
    Ro, to, io = oracle.register(p.src, p.dst)
    with Registrar(lib=gpu_lib) as one, Registrar(lib=gpu_lib) as sh:
        for r in (one, sh):
            set_params(r, tau_compat=p.tau_compat, tau_inlier=p.tau_inlier)
        R1, t1, i1 = one.register(p.src, p.dst)
        assert one.get("pruned_pairs") == 1
        sh.comm_init(rank=0, world=1, unique_id=bytes(ident))
        R, t, inl = sh.register_sharded(p.src, p.dst)
        assert sh.get("pruned_pairs") == 0
        np.testing.assert_array_equal(R, R1)
        np.testing.assert_array_equal(t, t1)
        assert inl == i1
        compare_pose(R, t, inl, Ro, to, io)


def test_node_prune_probe_backs_off_and_comes_back(gpu_lib, oracle):
    """A ctx whose calls prune nothing stops trying for node_prune_probe calls (five near-empty launches per chunk
    saved), then tries again; the results never depend on it."""
    dense = synth.make_config_pair("cfg2_3dmatch_256x5000", 12)     # indoor scale: nothing to prune
    sparse = synth.make_pair(6000, 0.03, 9480, **KITTI)              # outdoor scale: prunes to the inlier clique
    set_params(oracle, tau_compat=sparse.tau_compat, tau_inlier=sparse.tau_inlier)
    Ro, to, io = oracle.register(sparse.src, sparse.dst)
    with Registrar(lib=gpu_lib, device=0) as reg:
        reg.set("node_prune_probe", 3)
        assert reg.get("node_prune_probe") == 3 and reg.get("node_prune_trying") == 1
        set_params(reg, tau_compat=dense.tau_compat, tau_inlier=dense.tau_inlier)
        for _ in range(2):
            reg.register(dense.src, dense.dst)
            assert reg.get("pruned_pairs") == 0
        assert reg.get("node_prune_trying") == 0
        set_params(reg, tau_compat=sparse.tau_compat, tau_inlier=sparse.tau_inlier)
        outs = []
        for _ in range(3):   # inside the back-off: the dense kernel counts every triangle of the pair
            outs.append(reg.register(sparse.src, sparse.dst))
            assert reg.get("pruned_pairs") == 0
        assert reg.get("node_prune_trying") == 1
        outs.append(reg.register(sparse.src, sparse.dst))
        assert reg.get("pruned_pairs") == 1
        for R, t, inl in outs:
            np.testing.assert_array_equal(R, outs[0][0])
            np.testing.assert_array_equal(t, outs[0][1])
            assert inl == outs[0][2]
        compare_pose(*outs[-1], Ro, to, io)

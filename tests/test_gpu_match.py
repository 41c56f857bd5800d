"""GPU parity of the correspondence front end (sac_cot_match_packed): tensor-core search + exact fp32 decision against
the oracle's brute force.  Indices are bit-exact by construction of the candidate margin (DESIGN.md §4 "match")."""
import numpy as np
import pytest

from sac_cot_b200 import _abi, synth
from sac_cot_b200.api import Registrar, SacCotError

pytestmark = pytest.mark.gpu


def both(gpu, oracle, f, xs, g, xd):
    ng, csg, cdg = gpu.match(f, xs, g, xd)
    no, cso, cdo = oracle.match(f, xs, g, xd)
    np.testing.assert_array_equal(ng, no)
    np.testing.assert_array_equal(csg, cso)
    np.testing.assert_array_equal(cdg, cdo)
    return ng


@pytest.fixture()
def mgpu(gpu_lib):
    reg = Registrar(lib=gpu_lib, device=0)
    yield reg
    reg.close()


@pytest.mark.parametrize("path", [1, 0])
@pytest.mark.parametrize("Ns,Nd,dim", [(1, 1, 33), (5, 300, 33), (129, 257, 33), (1000, 777, 33), (128, 256, 33), (300, 513, 8),
                                       (200, 900, 40), (64, 100, 1), (150, 333, 41), (100, 300, 64), (5000, 5000, 33)])
def test_random_descriptors_match_oracle(mgpu, oracle, path, Ns, Nd, dim):
    if path == 0 and Ns * Nd > 2_000_000:
        pytest.skip("exhaustive scan of the large case is covered by the tensor path's fallback tests")
    mgpu.set("match_path", path)
    rng = np.random.default_rng(Ns * 31 + Nd)
    f = (rng.random((Ns, dim)) * 10).astype(np.float32)
    g = (rng.random((Nd, dim)) * 10).astype(np.float32)
    both(mgpu, oracle, f, rng.random((Ns, 3)).astype(np.float32), g, rng.random((Nd, 3)).astype(np.float32))


def test_fpfh_like_descriptors_and_handoff_to_registration(mgpu, oracle):
    p = synth.make_config_pair("cfg2_3dmatch_256x5000", 2)
    f, g = synth.make_descriptors(p, 33, seed=5)
    nn = both(mgpu, oracle, f, p.src, g, p.dst)
    assert (nn[p.inlier_idx] == p.inlier_idx).all()
    _, cs, cd = mgpu.match(f, p.src, g, p.dst)
    set_ = lambda r: (setattr(r.params, "tau_compat", p.tau_compat), setattr(r.params, "tau_inlier", p.tau_inlier))  # noqa: E731
    set_(mgpu), set_(oracle)
    Rg, tg, ig = mgpu.register(cs, cd)
    Ro, to, io = oracle.register(cs, cd)
    assert ig == io
    ang, dt = synth.pose_error(Rg, tg, p.R_gt, p.t_gt)
    assert ang < np.deg2rad(1.0) and dt < 0.02


def test_exact_ties_near_ties_and_overflowing_candidate_lists(mgpu, oracle):
    rng = np.random.default_rng(3)
    Nd, dim = 700, 33
    g = (rng.random((Nd, dim)) * 5).astype(np.float32)
    g[600] = g[17]                       # exact duplicate: lowest index wins
    g[300] = g[17]
    g[450] = np.nextafter(g[18], np.float32(np.inf))   # one ulp away in every component: decided by the exact chain
    for k in range(40):                  # 40 copies of one descriptor: the candidate list (8) overflows -> exhaustive scan
        g[100 + k] = g[99]
    f = np.concatenate([g[[17, 18, 450, 99, 120]], (rng.random((60, dim)) * 5).astype(np.float32)])
    nn = both(mgpu, oracle, f, np.zeros((len(f), 3), np.float32), g, rng.random((Nd, 3)).astype(np.float32))
    assert nn[0] == 17 and nn[3] == 99 and nn[4] == 99
    # every descriptor identical: every column is a candidate of every row
    both(mgpu, oracle, np.ones((200, dim), np.float32), np.zeros((200, 3), np.float32), np.ones((600, dim), np.float32),
         np.zeros((600, 3), np.float32))


@pytest.mark.parametrize("scale", [1e-6, 1e-2, 1e3, 1e6])
def test_descriptor_magnitudes(mgpu, oracle, scale):
    rng = np.random.default_rng(11)
    g = (rng.standard_normal((900, 33)) * scale).astype(np.float32)          # signed values too
    f = (g[rng.integers(0, 900, 400)] + rng.standard_normal((400, 33)).astype(np.float32) * np.float32(scale * 1e-3)).astype(np.float32)
    both(mgpu, oracle, f, np.zeros((400, 3), np.float32), g, np.zeros((900, 3), np.float32))


def test_clustered_descriptors_with_many_close_neighbours(mgpu, oracle):
    # 20 tight clusters: each row has ~45 neighbours at almost the same distance (inside the candidate margin or close)
    rng = np.random.default_rng(12)
    centres = (rng.random((20, 33)) * 10).astype(np.float32)
    g = (centres[rng.integers(0, 20, 900)] + rng.standard_normal((900, 33)).astype(np.float32) * np.float32(1e-3)).astype(np.float32)
    f = (centres[rng.integers(0, 20, 300)] + rng.standard_normal((300, 33)).astype(np.float32) * np.float32(1e-3)).astype(np.float32)
    both(mgpu, oracle, f, np.zeros((300, 3), np.float32), g, np.zeros((900, 3), np.float32))


def test_ragged_batch_and_device_resident_handoff(gpu_lib, oracle):
    import torch
    rng = np.random.default_rng(13)
    sizes = [(300, 400), (1, 9), (1000, 700), (129, 129)]
    fs = [(rng.random((a, 33)) * 10).astype(np.float32) for a, _ in sizes]
    gs = [(rng.random((b, 33)) * 10).astype(np.float32) for _, b in sizes]
    xs = [rng.random((a, 3)).astype(np.float32) for a, _ in sizes]
    xd = [rng.random((b, 3)).astype(np.float32) for _, b in sizes]
    no, cso, cdo, offs = oracle.match_batch(fs, xs, gs, xd)
    with Registrar(lib=gpu_lib) as g:
        ng, csg, cdg, offs_g = g.match_batch(fs, xs, gs, xd)
        np.testing.assert_array_equal(offs, offs_g)
        np.testing.assert_array_equal(ng, no)
        np.testing.assert_array_equal(cdg, cdo)
        assert g.get("launches") == 5   # two operand-image kernels, the tensor-core sweep, the exact decision, the (idle) scan
    # device-resident: descriptors and keypoints in HBM, correspondences stay there
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    od = np.zeros(len(sizes) + 1, np.int64)
    np.cumsum([b for _, b in sizes], out=od[1:])
    T = lambda parts: torch.from_numpy(np.concatenate(parts)).to(dev)  # noqa: E731
    dfs, dgs, dxs, dxd = T(fs), T(gs), T(xs), T(xd)
    d_nn = torch.empty(int(offs[-1]), dtype=torch.int32, device=dev)
    d_cs = torch.empty((int(offs[-1]), 3), dtype=torch.float32, device=dev)
    d_cd = torch.empty((int(offs[-1]), 3), dtype=torch.float32, device=dev)
    with Registrar(lib=gpu_lib, device=0, stream=stream.cuda_stream) as g:
        g.match_packed_ptr(dfs.data_ptr(), dxs.data_ptr(), offs, dgs.data_ptr(), dxd.data_ptr(), od, 33, d_nn.data_ptr(),
                           d_cs.data_ptr(), d_cd.data_ptr(), _abi.LOC_DEVICE)
        stream.synchronize()
    np.testing.assert_array_equal(d_nn.cpu().numpy(), no)
    np.testing.assert_array_equal(d_cd.cpu().numpy(), cdo)


def test_argument_checks(mgpu):
    f = np.zeros((4, 33), np.float32)
    x = np.zeros((4, 3), np.float32)
    with pytest.raises(SacCotError) as ei:
        mgpu.match(np.zeros((4, 300), np.float32), x, np.zeros((4, 300), np.float32), x)
    assert ei.value.status == _abi.E_SIZE
    with pytest.raises(ValueError):
        mgpu.match(f, x, np.zeros((4, 32), np.float32), x)
    with pytest.raises(SacCotError):
        mgpu.set("match_path", 2)


def test_mutual_filter_matches_oracle(gpu_lib, oracle):
    import torch
    sizes = [(500, 640), (129, 77), (1000, 1000)]
    rng = np.random.default_rng(21)
    G = [(rng.random((b, 33)) * 10).astype(np.float32) for _, b in sizes]
    F = []
    for (a, b), g in zip(sizes, G):
        f = (rng.random((a, 33)) * 10).astype(np.float32)
        k = min(a, b) // 2
        f[:k] = g[:k] + (rng.standard_normal((k, 33)) * 0.01).astype(np.float32)   # half of the rows have a true partner
        F.append(f)
    XS = [rng.random((a, 3)).astype(np.float32) for a, _ in sizes]
    XD = [rng.random((b, 3)).astype(np.float32) for _, b in sizes]
    cso, cdo, offo = oracle.match_mutual_batch(F, XS, G, XD)
    with Registrar(lib=gpu_lib) as g:
        csg, cdg, offg = g.match_mutual_batch(F, XS, G, XD)
    np.testing.assert_array_equal(offg, offo)
    np.testing.assert_array_equal(csg, cso)
    np.testing.assert_array_equal(cdg, cdo)
    assert all(offo[b + 1] - offo[b] >= min(a, c) // 2 for b, (a, c) in enumerate(sizes))
    # device-resident form: everything stays in HBM, the offsets come back with one small copy
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    os_ = np.zeros(len(sizes) + 1, np.int64); np.cumsum([a for a, _ in sizes], out=os_[1:])
    od_ = np.zeros(len(sizes) + 1, np.int64); np.cumsum([b for _, b in sizes], out=od_[1:])
    T = lambda parts: torch.from_numpy(np.concatenate(parts)).to(dev)  # noqa: E731
    dF, dG, dXS, dXD = T(F), T(G), T(XS), T(XD)
    ns, nd = int(os_[-1]), int(od_[-1])
    d_nn = torch.empty(ns, dtype=torch.int32, device=dev); d_nb = torch.empty(nd, dtype=torch.int32, device=dev)
    d_cs = torch.empty((ns, 3), device=dev); d_cd = torch.empty((ns, 3), device=dev)
    d_bs = torch.empty((nd, 3), device=dev); d_bd = torch.empty((nd, 3), device=dev)
    d_os = torch.empty((ns, 3), device=dev); d_od = torch.empty((ns, 3), device=dev)
    d_off = torch.empty(len(sizes) + 1, dtype=torch.int64, device=dev)
    i64p = lambda a: a.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_int64))  # noqa: E731
    with Registrar(lib=gpu_lib, device=0, stream=stream.cuda_stream) as g:
        g.match_packed_ptr(dF.data_ptr(), dXS.data_ptr(), os_, dG.data_ptr(), dXD.data_ptr(), od_, 33, d_nn.data_ptr(),
                           d_cs.data_ptr(), d_cd.data_ptr(), _abi.LOC_DEVICE)
        g.match_packed_ptr(dG.data_ptr(), dXD.data_ptr(), od_, dF.data_ptr(), dXS.data_ptr(), os_, 33, d_nb.data_ptr(),
                           d_bs.data_ptr(), d_bd.data_ptr(), _abi.LOC_DEVICE)
        rc = gpu_lib.sac_cot_match_mutual(g._ctx, d_nn.data_ptr(), d_nb.data_ptr(), d_cs.data_ptr(), d_cd.data_ptr(), i64p(os_),
                                          i64p(od_), len(sizes), d_os.data_ptr(), d_od.data_ptr(), d_off.data_ptr(), _abi.LOC_DEVICE)
        assert rc == 0
        stream.synchronize()
    np.testing.assert_array_equal(d_off.cpu().numpy(), offo)
    n = int(offo[-1])
    np.testing.assert_array_equal(d_os.cpu().numpy()[:n], cso)
    np.testing.assert_array_equal(d_od.cpu().numpy()[:n], cdo)

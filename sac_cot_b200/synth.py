"""Deterministic synthetic correspondence sets with a known ground-truth pose.

Follows SURVEY.md §8d ("Synthetic inputs"): the reference (/root/reference/README.md:1-2)
ships no data and 3DMatch / 3DLoMatch / KITTI are not available offline, so every config of
BASELINE.json is realised as a seeded synthetic pair:

  src_n ~ U(box);  GT rotation = uniform random unit quaternion;  GT t ~ U([-L/2, L/2]^3)
  inlier  n: dst_n = R src_n    + t + eps,  eps ~ U([-sigma, sigma]^3), sigma = tau_c / 4
  outlier n: dst_n = R src_pi(n) + t + eps  with pi(n) a random *other* index
             ("FPFH-like" wrong match to a real keypoint)
  inlier positions are a random subset of the indices.

With sigma = tau_c/4 the inliers are pairwise compatible in exact arithmetic
(| |d_i-d_j| - |s_i-s_j| | <= |eps_i - eps_j| <= 2*sqrt(3)*sigma ~= 0.87 tau_c).

The PRNG is a counter-based splitmix64 evaluated with numpy uint64 arithmetic, so a given
(seed, N, ...) yields the same arrays on every machine and numpy version; only + - * / sqrt
are used (no libm transcendentals).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 counters."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


class Stream:
    """Counter-based random stream: value k of stream (seed, lane) is splitmix64 of a hash."""

    def __init__(self, seed: int, lane: int):
        base = _splitmix64(np.array([seed & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
        with np.errstate(over="ignore"):
            self._base = _splitmix64(np.array([base ^ np.uint64(lane * 0x9E3779B1 + 1)], dtype=np.uint64))[0]
        self._ctr = 0

    def bits(self, n: int) -> np.ndarray:
        idx = np.arange(self._ctr, self._ctr + n, dtype=np.uint64)
        self._ctr += n
        with np.errstate(over="ignore"):
            return _splitmix64((idx * np.uint64(0xD1342543DE82EF95) + self._base) & _M64)

    def uniform(self, n: int) -> np.ndarray:
        """n doubles in [0, 1) with 53 random bits (exact int -> float conversion)."""
        return (self.bits(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def below(self, n: int, bound: int) -> np.ndarray:
        """n integers in [0, bound) (multiply-shift; bias < 2^-32, irrelevant here)."""
        hi = (self.bits(n) >> np.uint64(32)).astype(np.uint64)
        return ((hi * np.uint64(bound)) >> np.uint64(32)).astype(np.int64)


@dataclass
class Pair:
    src: np.ndarray          # (N,3) float32
    dst: np.ndarray          # (N,3) float32
    R_gt: np.ndarray         # (3,3) float64
    t_gt: np.ndarray         # (3,)  float64
    inlier_idx: np.ndarray   # sorted int64 indices of the true inliers
    tau_compat: float
    tau_inlier: float


def _random_rotation(st: Stream) -> np.ndarray:
    # uniform unit quaternion by rejection from the 4-cube (only + * sqrt)
    while True:
        q = st.uniform(4) * 2.0 - 1.0
        n2 = float(q @ q)
        if 1e-3 < n2 <= 1.0:
            break
    w, x, y, z = q / np.sqrt(n2)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)


def make_pair(N: int, inlier_ratio: float, seed: int, *, box=(3.0, 3.0, 3.0), tau_compat: float = 0.1,
              tau_inlier: float | None = None, outliers: str = "perm", n_inliers: int | None = None) -> Pair:
    """One synthetic correspondence set.  outliers: 'perm' (wrong match to a real keypoint)
    or 'uniform' (dst uniform in the transformed box)."""
    if tau_inlier is None:
        tau_inlier = tau_compat
    box = np.asarray(box, dtype=np.float64)
    s_pts, s_rot, s_in, s_eps, s_pi = (Stream(seed, k) for k in range(5))
    src = (s_pts.uniform(3 * N).reshape(N, 3) - 0.5) * box
    R = _random_rotation(s_rot)
    t = (s_rot.uniform(3) - 0.5) * box
    n_in = int(round(N * inlier_ratio)) if n_inliers is None else int(n_inliers)
    n_in = max(0, min(N, n_in))
    # random subset: the n_in smallest of N random keys (stable argsort => deterministic)
    order = np.argsort(s_in.bits(N), kind="stable")
    inlier_idx = np.sort(order[:n_in])
    is_in = np.zeros(N, dtype=bool)
    is_in[inlier_idx] = True
    sigma = tau_compat / 4.0
    eps = (s_eps.uniform(3 * N).reshape(N, 3) * 2.0 - 1.0) * sigma
    match = np.arange(N)
    if outliers == "perm":
        wrong = (match + 1 + s_pi.below(N, N - 1)) % N  # any index except n itself
        match = np.where(is_in, match, wrong)
        dst = src[match] @ R.T + t + eps
    elif outliers == "uniform":
        dst = src @ R.T + t + eps
        rnd = (s_pi.uniform(3 * N).reshape(N, 3) - 0.5) * box
        dst = np.where(is_in[:, None], dst, rnd @ R.T + t)
    else:
        raise ValueError(outliers)
    return Pair(np.ascontiguousarray(src, dtype=np.float32), np.ascontiguousarray(dst, dtype=np.float32),
                R, t, inlier_idx, float(tau_compat), float(tau_inlier))


# BASELINE.json `configs`, realised as SURVEY.md §8d lays out.
CONFIGS = {
    # name: (pairs, N, inlier ratio(s), box, tau_c, first seed)
    "cfg1_single_n1000": dict(pairs=1, N=1000, ratios=(0.10,), box=(3.0, 3.0, 3.0), tau=0.1, seed0=1),
    "cfg2_3dmatch_256x5000": dict(pairs=256, N=5000, ratios=(0.05,), box=(3.0, 3.0, 3.0), tau=0.1, seed0=1000),
    "cfg3_3dlomatch_256x5000": dict(pairs=256, N=5000, ratios=(0.01, 0.02), box=(3.0, 3.0, 3.0), tau=0.1, seed0=3000),
    "cfg4_kitti_128x10000": dict(pairs=128, N=10000, ratios=(0.03,), box=(60.0, 60.0, 6.0), tau=0.6, seed0=5000),
    "cfg5_single_n50000": dict(pairs=1, N=50000, ratios=(0.05,), box=(3.0, 3.0, 3.0), tau=0.1, seed0=7000),
}


def make_config_pair(name: str, b: int, seed_shift: int = 0) -> Pair:
    c = CONFIGS[name]
    ratio = c["ratios"][b % len(c["ratios"])]
    return make_pair(c["N"], ratio, c["seed0"] + b + seed_shift, box=c["box"], tau_compat=c["tau"])


def pose_error(R: np.ndarray, t: np.ndarray, R_gt: np.ndarray, t_gt: np.ndarray) -> tuple[float, float]:
    """(rotation angle error [rad], translation error [units]) of an estimate vs ground truth."""
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    dR = R @ np.asarray(R_gt, dtype=np.float64).T
    c = max(-1.0, min(1.0, (np.trace(dR) - 1.0) / 2.0))
    # angle from both the trace and the skew part: accurate near 0 where acos is ill-conditioned
    sk = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    ang = float(np.arctan2(np.linalg.norm(sk), c))
    return ang, float(np.linalg.norm(np.asarray(t, dtype=np.float64).ravel() - np.asarray(t_gt).ravel()))


def make_descriptors(pair: Pair, dim: int = 33, seed: int = 0, noise: float = 0.05):
    """FPFH-like descriptors for a synthetic pair (no descriptor data is available offline).

    Target keypoint j gets a random non-negative histogram-like descriptor g_j (uniform in [0, 1), scaled to sum 100 as
    FPFH bins are); source keypoint i gets the descriptor of the target keypoint it truly corresponds to plus uniform
    noise of relative size `noise` for the inliers, and an unrelated random descriptor otherwise.  Returns
    (desc_src (N, dim), desc_dst (N, dim)) float32: brute-force nearest-neighbour matching src -> dst then recovers
    the inlier correspondences and assigns the outliers arbitrary (wrong) partners, like a real front end."""
    N = pair.src.shape[0]
    st_g, st_n, st_o = (Stream(seed * 7919 + 17, k) for k in range(3))
    g = st_g.uniform(N * dim).reshape(N, dim)
    g = g / g.sum(axis=1, keepdims=True) * 100.0
    f = st_o.uniform(N * dim).reshape(N, dim)
    f = f / f.sum(axis=1, keepdims=True) * 100.0
    eps = (st_n.uniform(N * dim).reshape(N, dim) * 2.0 - 1.0) * noise * (100.0 / dim)
    f[pair.inlier_idx] = g[pair.inlier_idx] + eps[pair.inlier_idx]
    return np.ascontiguousarray(f, dtype=np.float32), np.ascontiguousarray(g, dtype=np.float32)

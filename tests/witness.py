"""Independent numpy/networkx restatements used to pin the oracle.

None of this shares code with oracle/sac_cot_oracle.cpp: numpy elementwise fp32 ops are
individually rounded IEEE operations, so the graph witness follows the same operation order
bit for bit; triangle counts come from integer matrix products / networkx; Kabsch from
numpy's fp64 SVD; scoring from an fp64 evaluation with an explicit ambiguity band.
"""
import numpy as np


def stride_words(N):
    return ((N + 127) // 128) * 4


def graph_dense(src, dst, tau):
    """bool (N,N): |  ||s_i-s_j|| - ||d_i-d_j||  | < tau in fp32, ops in SURVEY §8a S1 order."""
    s = np.asarray(src, np.float32)
    d = np.asarray(dst, np.float32)

    def length(p):
        a = p[:, None, 0] - p[None, :, 0]
        b = p[:, None, 1] - p[None, :, 1]
        c = p[:, None, 2] - p[None, :, 2]
        s2 = (a * a + b * b) + c * c
        assert s2.dtype == np.float32
        return np.sqrt(s2)

    A = np.abs(length(s) - length(d)) < np.float32(tau)
    np.fill_diagonal(A, False)
    return A


def pack_adj(A):
    N = A.shape[0]
    W = stride_words(N)
    bits = np.zeros((N, W * 32), dtype=np.uint8)
    bits[:, :N] = A
    return np.packbits(bits, axis=1, bitorder="little").view(np.uint32).reshape(N, W)


def unpack_adj(words, N):
    W = stride_words(N)
    w = np.asarray(words, np.uint32).reshape(N, W)
    bits = np.unpackbits(w.view(np.uint8), axis=1, bitorder="little")
    return bits[:, :N].astype(bool), bits[:, N:]


def triangle_counts(A):
    """(T (N,N) int64 with T_ij = #common neighbours on edges, t (N,) triangles per node)."""
    Ai = A.astype(np.int64)
    T = (Ai @ Ai) * Ai
    t = T.sum(axis=1) // 2
    return T, t


def edge_keys(A, T):
    i, j = np.nonzero(np.triu(A, 1))
    return (T[i, j].astype(np.uint64) << np.uint64(32)) | ((0xFFFF - i).astype(np.uint64) << np.uint64(16)) \
        | (0xFFFF - j).astype(np.uint64)


def decode_edge(key):
    key = int(key)
    return key >> 32, 0xFFFF - ((key >> 16) & 0xFFFF), 0xFFFF - (key & 0xFFFF)


def select_triangles(A, t, top_edges, m, K):
    tri = -np.ones((K, 3), np.int32)
    for r, key in enumerate(top_edges):
        _, i, j = decode_edge(key)
        ks = np.nonzero(A[i] & A[j])[0]
        order = sorted(ks.tolist(), key=lambda k: (-int(t[k]), k))[:m]
        for q, k in enumerate(order):
            tri[r * m + q] = (i, j, k)
    return tri


def kabsch_svd(P, Q):
    """fp64 least-squares rigid transform Q ~= R P + t (rows are points)."""
    P = np.asarray(P, np.float64)
    Q = np.asarray(Q, np.float64)
    pc, qc = P.mean(0), Q.mean(0)
    H = (P - pc).T @ (Q - qc)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    return R, qc - R @ pc


def rot_angle(Ra, Rb):
    dR = np.asarray(Ra, np.float64).reshape(3, 3) @ np.asarray(Rb, np.float64).reshape(3, 3).T
    sk = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    return float(np.arctan2(np.linalg.norm(sk), (np.trace(dR) - 1.0) / 2.0))


def residual2_f64(rt, src, dst):
    R = np.asarray(rt[:9], np.float64).reshape(3, 3)
    t = np.asarray(rt[9:], np.float64)
    e = np.asarray(src, np.float64) @ R.T + t - np.asarray(dst, np.float64)
    return (e * e).sum(axis=1)


def count_bounds(rt, src, dst, tau_in, rel=1e-4):
    """(lo, hi): inlier counts excluding / including the points within `rel` of the threshold."""
    tau2 = float(np.float32(tau_in) * np.float32(tau_in))
    r2 = residual2_f64(rt, src, dst)
    return int((r2 < tau2 * (1 - rel)).sum()), int((r2 < tau2 * (1 + rel)).sum())

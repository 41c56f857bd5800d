"""CPU check of the facts the exact node pruning of S2 rests on (sac_cot_b200/csrc/kernels_prune.cu, the PRUNED
instance of the apex kernel in kernels_select.cu), on graphs the from-paper oracle builds:

  (1) an edge's triangle count is at most the smaller endpoint degree minus one, hence every selected edge joins two
      nodes of degree >= theta + 1 for any theta at or below the K_e-th largest count (the "kept" nodes);
  (2) restricted to the kept rows, the loop "for every kept k, for every neighbour j: popc(row_k & row_j)" yields every
      selectable edge key and the exact node count of every kept node;
  (3) t_k <= deg_k (deg_k - 1) / 2 and t_k <= 1/2 sum_{n in N(k)} (min(deg_k, deg_n) - 1): the two bounds the apex kernel
      uses for candidates outside the kept set before it evaluates one exactly;
  (4) with those bounds the first m hits of the kept-node rank list are the m best candidates of a selected edge
      whenever the m-th hit beats every outside candidate's bound — checked against the oracle's triangles.
"""
import numpy as np
import pytest

from sac_cot_b200 import _abi, synth

KITTI = dict(box=(60.0, 60.0, 6.0), tau_compat=0.6)


def unpack_adj(words, N):
    bits = np.unpackbits(words.view(np.uint8).reshape(N, -1), axis=1, bitorder="little")
    return bits[:, :N].astype(bool)


@pytest.mark.parametrize("N,ratio,Ke,m,kw", [(600, 0.10, 128, 4, {}), (900, 0.05, 256, 4, KITTI), (700, 0.0, 64, 2, {}),
                                             (500, 0.3, 512, 8, {})])
def test_degree_bounds_behind_node_pruning(oracle, N, ratio, Ke, m, kw):
    p = synth.make_pair(N, ratio, 7700 + N, **kw)
    oracle.params.tau_compat, oracle.params.tau_inlier = p.tau_compat, p.tau_inlier
    oracle.params.num_edges, oracle.params.apex_per_edge = Ke, m
    oracle.register(p.src, p.dst)
    A = unpack_adj(oracle.debug(0, _abi.DBG_ADJ), N)
    assert (A == A.T).all() and not A.diagonal().any()
    deg = A.sum(1).astype(np.int64)
    Af = A.astype(np.int64)
    T = (Af @ Af) * Af                                   # T_ij on the edges
    t = oracle.debug(0, _abi.DBG_T_NODE).astype(np.int64)
    assert (T.sum(1) // 2 == t).all()                    # t_i = 1/2 sum_j A_ij T_ij
    # (1)
    ii, jj = np.nonzero(np.triu(A, 1))
    assert (T[ii, jj] <= np.minimum(deg[ii], deg[jj]) - 1).all()
    top = oracle.debug(0, _abi.DBG_TOP_EDGES)
    if len(top) == 0:
        return
    Ttop = (top >> np.uint64(32)).astype(np.int64)
    ti = 0xFFFF - ((top >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
    tj = 0xFFFF - (top & np.uint64(0xFFFF)).astype(np.int64)
    for theta in {int(Ttop[-1]), max(1, int(Ttop[-1]) // 2)}:     # the exact K_e-th count and a looser certified bound
        kept = deg >= theta + 1
        assert kept[ti].all() and kept[tj].all()
        # (2) keys with T >= theta from the kept rows alone == all such keys of the graph
        K = np.nonzero(kept)[0]
        sub = T[np.ix_(K, np.arange(N))]
        keys_kept = {(int(sub[a, j]), int(K[a]), int(j)) for a in range(len(K)) for j in np.nonzero(A[K[a]])[0]
                     if kept[j] and j > K[a] and sub[a, j] >= theta}
        keys_all = {(int(T[i, j]), int(i), int(j)) for i, j in zip(ii, jj) if T[i, j] >= theta}
        assert keys_kept == keys_all
        assert (sub.sum(1) // 2 == t[K]).all()
    # (3)
    b1 = deg * (deg - 1) // 2
    mind = np.minimum(deg[:, None], deg[None, :]) - 1
    b2 = (mind * Af).sum(1) // 2
    assert (t <= b2).all() and (b2 <= b1).all()
    # (4) rank the kept nodes only, accept an edge's first m hits iff the m-th beats every outside candidate's bound
    theta = int(Ttop[-1])
    kept = deg >= theta + 1
    order = sorted(np.nonzero(kept)[0], key=lambda k: (-t[k], k))
    tri = oracle.debug(0, _abi.DBG_TRIANGLES).reshape(-1, m, 3)
    accepted = 0
    for r in range(len(top)):
        cand = A[ti[r]] & A[tj[r]]
        hits = [k for k in order if cand[k]][:m]
        outside = np.nonzero(cand & ~kept)[0]
        if len(hits) == m and (len(outside) == 0 or t[hits[-1]] > b2[outside].max()):
            accepted += 1
            assert [int(x) for x in tri[r, :, 2]] == [int(k) for k in hits]
    assert accepted > 0 or ratio == 0.0

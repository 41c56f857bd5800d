// sac_cot_oracle.cpp — from-paper CPU oracle of the SAC-COT hot path.
//
// *** TEST INFRASTRUCTURE.  NOT PRODUCT CODE. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (sac_cot_b200/) never links, imports or calls it.
//
// *** FROM THE PAPER, NOT FROM THE REPO — PARITY UNPINNED AT THE REFERENCE BOUNDARY. ***
// /root/reference/README.md:1-2 is the entire upstream repository: a title and a one-line
// pointer to the paper "SAC-COT: Sample Consensus by Sampling Compatibility Triangles in
// Graphs for 3-D Point Cloud Registration".  It holds no code, tests, golden vectors or
// dependencies, so there is nothing upstream to compile, import or check against.  This file
// restates the method's stages as fixed by BASELINE.json `north_star` and SURVEY.md §8a
// (S1..S7; each function below cites its row).  It is pinned instead by independent
// witnesses in tests/test_oracle_*.py (numpy fp32 same-order graph, networkx triangle
// counts, numpy fp64 SVD Kabsch, brute-force numpy scoring), by ground-truth pose recovery
// on synthetic data, and by committed golden files in tests/golden/.
//
// Arithmetic contract (what makes CPU and GPU bit-identical):
//   * fp32 stages use only IEEE-754 correctly rounded  + - * / sqrt fma, in the order written.
//   * build with  -O2 -ffp-contract=off -fno-fast-math  so a*b+c is never fused; every
//     intended fused multiply-add is an explicit fmaf().
//   * all sums over correspondences are integer (counts / fixed point) => order free.
//   * the fp64 refit (S7) is compared with a tolerance (1e-5 rad / 1e-5 units), not bits.
//
// Build: see oracle/Makefile (single thread by default; -fopenmp variant parallelises only
// over independent rows / edges / hypotheses with integer reductions, so results do not
// depend on the thread count).

#include "../include/sac_cot.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int kJacobiSweeps32 = 8;   // S4: fixed sweep count, fp32
constexpr int kJacobiSweeps64 = 12;  // S7: fixed sweep count, fp64

inline int stride_words_for(int N) { return ((N + 127) / 128) * 4; }

// ---------------------------------------------------------------------------------------
// Per-pair state (everything the debug getter can hand out).
// ---------------------------------------------------------------------------------------
struct PairState {
  int N = 0;
  int stride = 0;                     // 32-bit words per adjacency row
  sac_cot_params prm{};
  std::vector<float> src, dst;        // copies (sharded phases need them later)
  std::vector<uint32_t> adj;          // N x stride: the graph S2/S3 run on (A, or A2 in second-order mode)
  std::vector<uint32_t> adj_first;    // second-order mode: the first-order graph A
  std::vector<uint32_t> t_node;       // N
  std::vector<uint64_t> t2;           // N: sum of T over evaluated incident edges (= 2 t_i if world == 1)
  std::vector<uint64_t> edge_keys;    // E (row-major order of (i,j), i<j)
  std::vector<uint32_t> hist;         // 4096 bins of T>>4
  std::vector<uint64_t> top_edges;    // K_e' sorted descending
  std::vector<int32_t> tri;           // K x 3
  std::vector<float> hyp_rt;          // K x 12
  std::vector<uint64_t> hyp_key;      // K
  uint64_t best_key = 0;
  std::vector<uint32_t> mask;         // ceil(N/32)
  // sharded bookkeeping
  int rank = 0, world = 1;
};

// ---------------------------------------------------------------------------------------
// S1 — first-order length-consistency compatibility graph.
//   a=sx_i-sx_j; b=sy_i-sy_j; c=sz_i-sz_j; s2=(a*a+b*b)+c*c; ls=sqrt(s2); same for dst -> ld;
//   A_ij = |ls-ld| < tau_c, A_ii = 0.  Every operation individually rounded.
// Only i<j is evaluated and mirrored: negating a,b,c leaves the squares unchanged, so
// A_ji computed directly is bit-identical (tests/test_oracle_graph.py checks this against
// a full numpy evaluation).
// ---------------------------------------------------------------------------------------
inline bool compatible(const float* s, const float* d, int i, int j, float tau) {
  const float a = s[3 * i + 0] - s[3 * j + 0];
  const float b = s[3 * i + 1] - s[3 * j + 1];
  const float c = s[3 * i + 2] - s[3 * j + 2];
  const float aa = a * a, bb = b * b, cc = c * c;
  const float s2 = (aa + bb) + cc;
  const float ls = std::sqrt(s2);
  const float u = d[3 * i + 0] - d[3 * j + 0];
  const float v = d[3 * i + 1] - d[3 * j + 1];
  const float w = d[3 * i + 2] - d[3 * j + 2];
  const float uu = u * u, vv = v * v, ww = w * w;
  const float d2 = (uu + vv) + ww;
  const float ld = std::sqrt(d2);
  const float diff = ls - ld;
  return std::fabs(diff) < tau;
}

void build_graph(PairState& st) {
  const int N = st.N, W = st.stride;
  st.adj.assign(static_cast<size_t>(N) * W, 0u);
  const float* s = st.src.data();
  const float* d = st.dst.data();
  const float tau = st.prm.tau_compat;
  uint32_t* A = st.adj.data();
  // upper triangle, row i owns its words exclusively => race free under OpenMP
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int i = 0; i < N; ++i) {
    uint32_t* row = A + static_cast<size_t>(i) * W;
    for (int j = i + 1; j < N; ++j)
      if (compatible(s, d, i, j, tau)) row[j >> 5] |= 1u << (j & 31);
  }
  // mirror (serial over i so that each target row is written by one iteration at a time)
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int j = 0; j < N; ++j) {
    uint32_t* rowj = A + static_cast<size_t>(j) * W;
    for (int i = 0; i < j; ++i)
      if ((A[static_cast<size_t>(i) * W + (j >> 5)] >> (j & 31)) & 1u) rowj[i >> 5] |= 1u << (i & 31);
  }
}

// ---------------------------------------------------------------------------------------
// S2 — compatibility-triangle counts.
//   T_ij = popc(row_i & row_j) for every edge i<j;  t_i = 1/2 * sum_j A_ij T_ij.
// Sharded mode ("S2 partition") evaluates only the edges this rank owns: edge (i < j) lies in cell
// (cb, ic) = (j / 1920, i / 256) and cell (cb, ic) belongs to rank (257 cb + ic) mod world.  The partition only
// decides WHO counts an edge, never a value: summed over the ranks the results are those of world = 1.  With
// world = 1 every edge is evaluated.  t_node then holds this rank's partial contribution.
// ---------------------------------------------------------------------------------------
inline uint32_t edge_owner(int i, int j, int world) {
  const uint32_t cb = static_cast<uint32_t>(j) / 1920u, ic = static_cast<uint32_t>(i) / 256u;
  return (257u * cb + ic) % static_cast<uint32_t>(world);
}

void count_triangles(PairState& st) {
  const int N = st.N, W = st.stride;
  const uint32_t* A = st.adj.data();
  st.t_node.assign(N, 0u);
  st.hist.assign(4096, 0u);
  st.edge_keys.clear();
  const int W64 = W / 2;  // stride is a multiple of 4 words => rows are 8-byte aligned pairs
  std::vector<std::vector<uint64_t>> per_row(N);
  std::vector<uint64_t>& t2 = st.t2;
  t2.assign(N, 0);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8)
#endif
  for (int i = 0; i < N; ++i) {
    const uint64_t* ri = reinterpret_cast<const uint64_t*>(A + static_cast<size_t>(i) * W);
    std::vector<uint64_t>& out = per_row[i];
    for (int j = i + 1; j < N; ++j) {
      if (!((A[static_cast<size_t>(i) * W + (j >> 5)] >> (j & 31)) & 1u)) continue;
      if (st.world > 1 && edge_owner(i, j, st.world) != static_cast<uint32_t>(st.rank)) continue;
      const uint64_t* rj = reinterpret_cast<const uint64_t*>(A + static_cast<size_t>(j) * W);
      uint32_t T = 0;
      for (int w = 0; w < W64; ++w) T += static_cast<uint32_t>(__builtin_popcountll(ri[w] & rj[w]));
      out.push_back((static_cast<uint64_t>(T) << 32) |
                    (static_cast<uint64_t>(0xFFFFu - static_cast<uint32_t>(i)) << 16) |
                    static_cast<uint64_t>(0xFFFFu - static_cast<uint32_t>(j)));
    }
  }
  for (int i = 0; i < N; ++i)
    for (uint64_t k : per_row[i]) {
      const uint32_t T = static_cast<uint32_t>(k >> 32);
      const uint32_t j = 0xFFFFu - static_cast<uint32_t>(k & 0xFFFFu);
      t2[i] += T;
      t2[j] += T;
      st.hist[T >> 4] += 1;
      st.edge_keys.push_back(k);
    }
  // each triangle at node i is seen on two of its incident edges => t_i = t2/2 (world == 1).
  // Sharded: the partial sums t2 are exchanged and halved after the merge (phase 2).
  for (int i = 0; i < N; ++i) st.t_node[i] = static_cast<uint32_t>(t2[i] / 2);
}

// ---------------------------------------------------------------------------------------
// S3a — edge ranking: first K_e edges by key descending
//   key = T<<32 | (0xFFFF-i)<<16 | (0xFFFF-j)  <=>  order (T desc, i asc, j asc).
// ---------------------------------------------------------------------------------------
void top_k_desc(std::vector<uint64_t> keys, size_t k, std::vector<uint64_t>& out) {
  k = std::min(k, keys.size());
  std::partial_sort(keys.begin(), keys.begin() + static_cast<std::ptrdiff_t>(k), keys.end(),
                    [](uint64_t a, uint64_t b) { return a > b; });
  out.assign(keys.begin(), keys.begin() + static_cast<std::ptrdiff_t>(k));
}

// ---------------------------------------------------------------------------------------
// S3b — apex selection (SURVEY.md §8a row S3 (2),(3)): for the r-th edge (i,j) the apex
// candidates k in N(i) ∩ N(j), ordered (t_k desc, k asc); first m.  Hypothesis id h = r*m+q.
// ---------------------------------------------------------------------------------------
void select_triangles(PairState& st) {
  const int N = st.N, W = st.stride;
  const int Ke = st.prm.num_edges, m = st.prm.apex_per_edge;
  const int K = Ke * m;
  st.tri.assign(static_cast<size_t>(K) * 3, -1);
  const uint32_t* A = st.adj.data();
  const int nsel = static_cast<int>(st.top_edges.size());
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int r = 0; r < nsel; ++r) {
    const uint64_t key = st.top_edges[r];
    const int i = static_cast<int>(0xFFFFu - ((key >> 16) & 0xFFFFu));
    const int j = static_cast<int>(0xFFFFu - (key & 0xFFFFu));
    std::vector<uint64_t> cand;  // (t_k << 32) | (0xFFFFFFFF - k): larger = better
    for (int w = 0; w < W; ++w) {
      uint32_t bits = A[static_cast<size_t>(i) * W + w] & A[static_cast<size_t>(j) * W + w];
      while (bits) {
        const int b = __builtin_ctz(bits);
        bits &= bits - 1;
        const uint32_t k = static_cast<uint32_t>(w * 32 + b);
        cand.push_back((static_cast<uint64_t>(st.t_node[k]) << 32) | (0xFFFFFFFFu - k));
      }
    }
    (void)N;
    const size_t take = std::min<size_t>(static_cast<size_t>(m), cand.size());
    std::partial_sort(cand.begin(), cand.begin() + static_cast<std::ptrdiff_t>(take), cand.end(),
                      [](uint64_t a, uint64_t b) { return a > b; });
    for (size_t q = 0; q < take; ++q) {
      const int k = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(cand[q] & 0xFFFFFFFFu));
      int32_t* o = &st.tri[(static_cast<size_t>(r) * m + q) * 3];
      o[0] = i; o[1] = j; o[2] = k;
    }
  }
}

// ---------------------------------------------------------------------------------------
// S4 — 3-point rigid hypothesis: Horn's quaternion form of Kabsch.
//   centroids (sum then /3), centred cross-covariance S (3 terms, explicit fma chain),
//   Horn's symmetric 4x4 matrix, cyclic Jacobi with a FIXED number of sweeps (skip a_pq==0),
//   eigenvector of the largest diagonal entry (first on ties) -> unit quaternion -> R,
//   t = qc - R pc.  Only + - * / sqrt and explicit fmaf; order exactly as written.
// Templated on the scalar so the fp64 refit (S7) reuses the algebra; the GPU library has its
// own, separately written, implementation of the same operation sequence.
// ---------------------------------------------------------------------------------------
template <typename F>
struct Jacobi4 {
  F a[4][4];
  F v[4][4];
  void rotate(int p, int q) {
    const F apq = a[p][q];
    if (apq == F(0)) return;
    const F theta = (a[q][q] - a[p][p]) / (F(2) * apq);
    const F ath = std::fabs(theta);
    const F rad = std::sqrt(theta * theta + F(1));
    F t = F(1) / (ath + rad);
    if (theta < F(0)) t = -t;
    const F c = F(1) / std::sqrt(t * t + F(1));
    const F s = t * c;
    const F tapq = t * apq;
    a[p][p] = a[p][p] - tapq;
    a[q][q] = a[q][q] + tapq;
    a[p][q] = F(0);
    a[q][p] = F(0);
    for (int r = 0; r < 4; ++r) {
      if (r == p || r == q) continue;
      const F arp = a[r][p], arq = a[r][q];
      const F np_ = c * arp - s * arq;
      const F nq_ = s * arp + c * arq;
      a[r][p] = np_; a[p][r] = np_;
      a[r][q] = nq_; a[q][r] = nq_;
    }
    for (int r = 0; r < 4; ++r) {
      const F vrp = v[r][p], vrq = v[r][q];
      v[r][p] = c * vrp - s * vrq;
      v[r][q] = s * vrp + c * vrq;
    }
  }
};

// S: cross-covariance, S[r][c] = sum_k p~_k[r] * q~_k[c].  Writes R (row-major) only.
template <typename F>
void horn_rotation(const F S[3][3], int sweeps, F R[9]) {
  Jacobi4<F> J;
  const F Sxx = S[0][0], Sxy = S[0][1], Sxz = S[0][2];
  const F Syx = S[1][0], Syy = S[1][1], Syz = S[1][2];
  const F Szx = S[2][0], Szy = S[2][1], Szz = S[2][2];
  J.a[0][0] = (Sxx + Syy) + Szz;
  J.a[1][1] = (Sxx - Syy) - Szz;
  J.a[2][2] = (Syy - Sxx) - Szz;
  J.a[3][3] = (Szz - Sxx) - Syy;
  J.a[0][1] = J.a[1][0] = Syz - Szy;
  J.a[0][2] = J.a[2][0] = Szx - Sxz;
  J.a[0][3] = J.a[3][0] = Sxy - Syx;
  J.a[1][2] = J.a[2][1] = Sxy + Syx;
  J.a[1][3] = J.a[3][1] = Szx + Sxz;
  J.a[2][3] = J.a[3][2] = Syz + Szy;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) J.v[r][c] = (r == c) ? F(1) : F(0);
  for (int sweep = 0; sweep < sweeps; ++sweep) {
    J.rotate(0, 1); J.rotate(0, 2); J.rotate(0, 3);
    J.rotate(1, 2); J.rotate(1, 3); J.rotate(2, 3);
  }
  int best = 0;
  for (int k = 1; k < 4; ++k)
    if (J.a[k][k] > J.a[best][best]) best = k;
  F w = J.v[0][best], x = J.v[1][best], y = J.v[2][best], z = J.v[3][best];
  const F n2 = ((w * w + x * x) + y * y) + z * z;
  if (!(n2 > F(0))) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    return;
  }
  const F n = std::sqrt(n2);
  w = w / n; x = x / n; y = y / n; z = z / n;
  const F xx = x * x, yy = y * y, zz = z * z;
  const F xy = x * y, xz = x * z, yz = y * z;
  const F wx = w * x, wy = w * y, wz = w * z;
  R[0] = F(1) - F(2) * (yy + zz);
  R[1] = F(2) * (xy - wz);
  R[2] = F(2) * (xz + wy);
  R[3] = F(2) * (xy + wz);
  R[4] = F(1) - F(2) * (xx + zz);
  R[5] = F(2) * (yz - wx);
  R[6] = F(2) * (xz - wy);
  R[7] = F(2) * (yz + wx);
  R[8] = F(1) - F(2) * (xx + yy);
}

void kabsch3(const float* src, const float* dst, int ia, int ib, int ic, float* rt /*12*/) {
  const float* p[3] = {src + 3 * ia, src + 3 * ib, src + 3 * ic};
  const float* q[3] = {dst + 3 * ia, dst + 3 * ib, dst + 3 * ic};
  float pc[3], qc[3];
  for (int a = 0; a < 3; ++a) {
    pc[a] = ((p[0][a] + p[1][a]) + p[2][a]) / 3.0f;
    qc[a] = ((q[0][a] + q[1][a]) + q[2][a]) / 3.0f;
  }
  float pt[3][3], qt[3][3];
  for (int k = 0; k < 3; ++k)
    for (int a = 0; a < 3; ++a) {
      pt[k][a] = p[k][a] - pc[a];
      qt[k][a] = q[k][a] - qc[a];
    }
  float S[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      S[r][c] = std::fmaf(pt[2][r], qt[2][c], std::fmaf(pt[1][r], qt[1][c], pt[0][r] * qt[0][c]));
  float R[9];
  horn_rotation<float>(S, kJacobiSweeps32, R);
  for (int a = 0; a < 9; ++a) rt[a] = R[a];
  for (int a = 0; a < 3; ++a) {
    const float rp = (R[3 * a + 0] * pc[0] + R[3 * a + 1] * pc[1]) + R[3 * a + 2] * pc[2];
    rt[9 + a] = qc[a] - rp;
  }
}

void make_hypotheses(PairState& st) {
  const int K = st.prm.num_edges * st.prm.apex_per_edge;
  st.hyp_rt.assign(static_cast<size_t>(K) * 12, 0.0f);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int h = 0; h < K; ++h) {
    const int32_t* tr = &st.tri[static_cast<size_t>(h) * 3];
    if (tr[0] < 0) continue;
    kabsch3(st.src.data(), st.dst.data(), tr[0], tr[1], tr[2], &st.hyp_rt[static_cast<size_t>(h) * 12]);
  }
}

// ---------------------------------------------------------------------------------------
// S5 — hypothesis scoring over all N correspondences.
//   x' = fma(R00,px, fma(R01,py, fma(R02,pz, tx)))  (same for y', z');  e = x' - qx ...;
//   r2 = fma(ez,ez, fma(ey,ey, ex*ex));  inlier iff r2 < tau_in^2 (tau_in^2 = one fp32 multiply).
//   mode 0: score = count.  mode 1: sum of floor(min(r2,tau2)/tau2 * 2^20) in u64 (lower=better).
// S6 — selection key  score' << 16 | (0xFFFF - h)  with score' = count+1 (mode 0) or
//   N*2^20 - sum + 1 (mode 1); invalid hypothesis => 0.  argmax => ties go to the lowest h.
// ---------------------------------------------------------------------------------------
inline float residual2(const float* rt, const float* p, const float* q) {
  const float xp = std::fmaf(rt[0], p[0], std::fmaf(rt[1], p[1], std::fmaf(rt[2], p[2], rt[9])));
  const float yp = std::fmaf(rt[3], p[0], std::fmaf(rt[4], p[1], std::fmaf(rt[5], p[2], rt[10])));
  const float zp = std::fmaf(rt[6], p[0], std::fmaf(rt[7], p[1], std::fmaf(rt[8], p[2], rt[11])));
  const float ex = xp - q[0], ey = yp - q[1], ez = zp - q[2];
  return std::fmaf(ez, ez, std::fmaf(ey, ey, ex * ex));
}

uint64_t score_one(const PairState& st, int h) {
  const int N = st.N;
  const float tau2 = st.prm.tau_inlier * st.prm.tau_inlier;
  const float* rt = &st.hyp_rt[static_cast<size_t>(h) * 12];
  const float* s = st.src.data();
  const float* d = st.dst.data();
  uint64_t score;
  if (st.prm.score_mode == SAC_COT_SCORE_INLIER_COUNT) {
    uint32_t cnt = 0;
    for (int n = 0; n < N; ++n) cnt += residual2(rt, s + 3 * n, d + 3 * n) < tau2 ? 1u : 0u;
    score = static_cast<uint64_t>(cnt) + 1;
  } else {
    uint64_t sum = 0;
    for (int n = 0; n < N; ++n) {
      const float r2 = residual2(rt, s + 3 * n, d + 3 * n);
      const float m = r2 < tau2 ? r2 : tau2;  // NaN r2 -> tau2 (treated as a full-cost outlier)
      const float qn = m / tau2;
      sum += static_cast<uint32_t>(qn * 1048576.0f);
    }
    score = (static_cast<uint64_t>(N) << 20) - sum + 1;
  }
  return (score << 16) | static_cast<uint64_t>(0xFFFFu - static_cast<uint32_t>(h));
}

void score_hypotheses(PairState& st, int h_begin, int h_end) {
  const int K = st.prm.num_edges * st.prm.apex_per_edge;
  st.hyp_key.assign(K, 0);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 32)
#endif
  for (int h = h_begin; h < h_end; ++h) {
    if (st.tri[static_cast<size_t>(h) * 3] < 0) continue;
    st.hyp_key[h] = score_one(st, h);
  }
  uint64_t best = 0;
  for (int h = h_begin; h < h_end; ++h) best = std::max(best, st.hyp_key[h]);
  st.best_key = best;
}

// ---------------------------------------------------------------------------------------
// S7 — final inlier refit: inlier set of the winner (bit-exact fp32
// test), one Kabsch over it with fp64 accumulation and an fp64 Horn/Jacobi solve, rounded to
// fp32.  Fewer than 3 inliers or refit == 0 => the winning hypothesis itself is returned.
// ---------------------------------------------------------------------------------------
void finalize(PairState& st, uint64_t best_key, float R[9], float t[3], int32_t* inliers) {
  const int N = st.N;
  st.mask.assign((N + 31) / 32, 0u);
  const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (best_key == 0) {
    std::memcpy(R, I, sizeof(I));
    t[0] = t[1] = t[2] = 0.0f;
    *inliers = 0;
    return;
  }
  const int h = static_cast<int>(0xFFFFu - static_cast<uint32_t>(best_key & 0xFFFFu));
  const float* rt = &st.hyp_rt[static_cast<size_t>(h) * 12];
  const float tau2 = st.prm.tau_inlier * st.prm.tau_inlier;
  const float* s = st.src.data();
  const float* d = st.dst.data();
  int cnt = 0;
  for (int n = 0; n < N; ++n)
    if (residual2(rt, s + 3 * n, d + 3 * n) < tau2) {
      st.mask[n >> 5] |= 1u << (n & 31);
      ++cnt;
    }
  *inliers = cnt;
  if (!st.prm.refit || cnt < 3) {
    for (int a = 0; a < 9; ++a) R[a] = rt[a];
    for (int a = 0; a < 3; ++a) t[a] = rt[9 + a];
    return;
  }
  double pc[3] = {0, 0, 0}, qc[3] = {0, 0, 0};
  for (int n = 0; n < N; ++n)
    if ((st.mask[n >> 5] >> (n & 31)) & 1u)
      for (int a = 0; a < 3; ++a) {
        pc[a] += static_cast<double>(s[3 * n + a]);
        qc[a] += static_cast<double>(d[3 * n + a]);
      }
  for (int a = 0; a < 3; ++a) { pc[a] /= cnt; qc[a] /= cnt; }
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int n = 0; n < N; ++n)
    if ((st.mask[n >> 5] >> (n & 31)) & 1u) {
      double pt[3], qt[3];
      for (int a = 0; a < 3; ++a) {
        pt[a] = static_cast<double>(s[3 * n + a]) - pc[a];
        qt[a] = static_cast<double>(d[3 * n + a]) - qc[a];
      }
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) S[r][c] += pt[r] * qt[c];
    }
  double Rd[9];
  horn_rotation<double>(S, kJacobiSweeps64, Rd);
  for (int a = 0; a < 9; ++a) R[a] = static_cast<float>(Rd[a]);
  for (int a = 0; a < 3; ++a) {
    const double rp = (Rd[3 * a + 0] * pc[0] + Rd[3 * a + 1] * pc[1]) + Rd[3 * a + 2] * pc[2];
    t[a] = static_cast<float>(qc[a] - rp);
  }
}

// ---------------------------------------------------------------------------------------
// argument checks shared by every entry point
// ---------------------------------------------------------------------------------------
// struct_size versioning: version 1 (32 bytes) ends with compat_mode (then called `reserved`, must be 0)
int check_params(const sac_cot_params* p) {
  if (!p) return SAC_COT_E_NULL;
  if (p->struct_size != sizeof(sac_cot_params) && p->struct_size != SAC_COT_PARAMS_SIZE_V1) return SAC_COT_E_PARAMS;
  if (!(p->tau_compat > 0.0f) || !(p->tau_inlier > 0.0f)) return SAC_COT_E_PARAMS;
  if (p->num_edges < 1 || p->num_edges > SAC_COT_MAX_EDGES) return SAC_COT_E_PARAMS;
  if (p->apex_per_edge < 1 || p->apex_per_edge > SAC_COT_MAX_APEX) return SAC_COT_E_PARAMS;
  if (p->num_edges * p->apex_per_edge > SAC_COT_MAX_HYPOTHESES) return SAC_COT_E_PARAMS;
  if (p->score_mode != 0 && p->score_mode != 1) return SAC_COT_E_PARAMS;
  if (p->refit != 0 && p->refit != 1) return SAC_COT_E_PARAMS;
  if (p->struct_size == SAC_COT_PARAMS_SIZE_V1) return p->compat_mode == 0 ? SAC_COT_OK : SAC_COT_E_PARAMS;
  if (p->compat_mode != SAC_COT_COMPAT_FIRST_ORDER && p->compat_mode != SAC_COT_COMPAT_SECOND_ORDER) return SAC_COT_E_PARAMS;
  if (p->so_min_common < 0 || p->so_min_common > 65535) return SAC_COT_E_PARAMS;
  if (p->reserved != 0) return SAC_COT_E_PARAMS;
  return SAC_COT_OK;
}

// the caller's struct (either version) as the current one
sac_cot_params normalized(const sac_cot_params& in) {
  sac_cot_params out{};
  std::memcpy(&out, &in, std::min<size_t>(in.struct_size, sizeof(out)));
  out.struct_size = sizeof(out);
  return out;
}

void load_pair(PairState& st, const float* src, const float* dst, int N, const sac_cot_params& prm) {
  st.N = N;
  st.stride = stride_words_for(N);
  st.prm = normalized(prm);
  st.src.assign(src, src + static_cast<size_t>(N) * 3);
  st.dst.assign(dst, dst + static_cast<size_t>(N) * 3);
  st.rank = 0;
  st.world = 1;
}

// Second-order graph: A2_ij = A_ij and popc(row_i(A) & row_j(A)) >= so_min_common.  The counts of
// pass 1 are exactly the per-edge triangle counts T_ij of the first-order graph (the SC^2 measure (A.A) o A).
void second_order_graph(PairState& st) {
  const int N = st.N, W = st.stride;
  count_triangles(st);  // T of every edge of A (world = 1 here)
  const uint32_t cmin = static_cast<uint32_t>(st.prm.so_min_common);
  st.adj_first = st.adj;
  std::vector<uint32_t> a2(static_cast<size_t>(N) * W, 0u);
  for (uint64_t key : st.edge_keys) {
    if (static_cast<uint32_t>(key >> 32) < cmin) continue;
    const int i = static_cast<int>(0xFFFFu - ((key >> 16) & 0xFFFFu)), j = static_cast<int>(0xFFFFu - (key & 0xFFFFu));
    a2[static_cast<size_t>(i) * W + (j >> 5)] |= 1u << (j & 31);
    a2[static_cast<size_t>(j) * W + (i >> 5)] |= 1u << (i & 31);
  }
  st.adj.swap(a2);
}

void run_pair(PairState& st, float R[9], float t[3], int32_t* inliers) {
  build_graph(st);
  st.adj_first.clear();
  if (st.prm.compat_mode == SAC_COT_COMPAT_SECOND_ORDER) second_order_graph(st);
  count_triangles(st);
  top_k_desc(st.edge_keys, static_cast<size_t>(st.prm.num_edges), st.top_edges);
  select_triangles(st);
  make_hypotheses(st);
  score_hypotheses(st, 0, st.prm.num_edges * st.prm.apex_per_edge);
  finalize(st, st.best_key, R, t, inliers);
}

}  // namespace

// =======================================================================================
// C ABI
// =======================================================================================
struct sac_cot_ctx {
  bool keep_debug = false;
  std::vector<PairState> pairs;  // retained when keep_debug (else only the last pair)
  PairState sharded;             // state of the sharded call in flight
};

extern "C" {

int sac_cot_params_default(sac_cot_params* p) {
  if (!p) return SAC_COT_E_NULL;
  p->struct_size = sizeof(sac_cot_params);
  p->tau_compat = 0.1f;
  p->tau_inlier = 0.1f;
  p->num_edges = 1024;
  p->apex_per_edge = 4;
  p->score_mode = SAC_COT_SCORE_INLIER_COUNT;
  p->refit = 1;
  p->compat_mode = SAC_COT_COMPAT_FIRST_ORDER;
  p->so_min_common = 0;
  p->reserved = 0;
  return SAC_COT_OK;
}

int sac_cot_ctx_create(sac_cot_ctx** out, int32_t /*device*/, void* /*stream*/) {
  if (!out) return SAC_COT_E_NULL;
  *out = new (std::nothrow) sac_cot_ctx();
  return *out ? SAC_COT_OK : SAC_COT_E_NOMEM;
}

int sac_cot_ctx_destroy(sac_cot_ctx* ctx) {
  delete ctx;
  return SAC_COT_OK;
}

int sac_cot_ctx_set(sac_cot_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return SAC_COT_E_NULL;
  if (!std::strcmp(name, "keep_debug")) { ctx->keep_debug = value != 0; return SAC_COT_OK; }
  if (!std::strcmp(name, "chunk_pairs") || !std::strcmp(name, "triangle_path") || !std::strcmp(name, "match_path")) return SAC_COT_OK;
  if (!std::strcmp(name, "threads")) {  // OpenMP build: worker threads of the calls that follow (a launcher may have
                                        // exported OMP_NUM_THREADS=1); ignored by the single-thread build
    if (value < 1) return SAC_COT_E_SIZE;
#ifdef _OPENMP
    omp_set_num_threads(static_cast<int>(value));
#endif
    return SAC_COT_OK;
  }
  return SAC_COT_E_WHICH;
}

int sac_cot_ctx_get(sac_cot_ctx* ctx, const char* name, int64_t* value) {
  if (!ctx || !name || !value) return SAC_COT_E_NULL;
  if (!std::strcmp(name, "launches") || !std::strcmp(name, "workspace_bytes") ||
      !std::strcmp(name, "retries") || !std::strcmp(name, "sm_count")) { *value = 0; return SAC_COT_OK; }
  if (!std::strcmp(name, "device")) { *value = -1; return SAC_COT_OK; }
  if (!std::strcmp(name, "threads")) {
#ifdef _OPENMP
    *value = omp_get_max_threads();
#else
    *value = 1;
#endif
    return SAC_COT_OK;
  }
  return SAC_COT_E_WHICH;
}

int sac_cot_register_packed(sac_cot_ctx* ctx, const float* src, const float* dst,
                            const int64_t* offsets, int32_t B, const sac_cot_params* params,
                            float* R, float* t, int32_t* inliers, int32_t location) {
  if (!ctx || !offsets) return SAC_COT_E_NULL;
  if (location != SAC_COT_LOC_HOST) return SAC_COT_E_UNSUPPORTED;
  if (B < 0) return SAC_COT_E_SIZE;
  if (int rc = check_params(params)) return rc;
  if (B > 0 && (!src || !dst || !R || !t || !inliers)) return SAC_COT_E_NULL;
  for (int b = 0; b < B; ++b) {
    const int64_t n = offsets[b + 1] - offsets[b];
    if (n < 3 || n > SAC_COT_MAX_N) return SAC_COT_E_SIZE;
  }
  try {
    ctx->pairs.clear();
    ctx->pairs.resize(ctx->keep_debug ? static_cast<size_t>(B) : (B > 0 ? 1 : 0));
    for (int b = 0; b < B; ++b) {
      PairState& st = ctx->pairs[ctx->keep_debug ? b : 0];
      load_pair(st, src + 3 * offsets[b], dst + 3 * offsets[b],
                static_cast<int>(offsets[b + 1] - offsets[b]), *params);
      run_pair(st, R + 9 * static_cast<size_t>(b), t + 3 * static_cast<size_t>(b), inliers + b);
    }
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
  return SAC_COT_OK;
}

int sac_cot_register_batch(sac_cot_ctx* ctx, const float* const* src, const float* const* dst,
                           const int32_t* N, int32_t B, const sac_cot_params* params,
                           float* R, float* t, int32_t* inliers) {
  if (!ctx) return SAC_COT_E_NULL;
  if (B < 0) return SAC_COT_E_SIZE;
  if (int rc = check_params(params)) return rc;
  if (B > 0 && (!src || !dst || !N || !R || !t || !inliers)) return SAC_COT_E_NULL;
  for (int b = 0; b < B; ++b) {
    if (!src[b] || !dst[b]) return SAC_COT_E_NULL;
    if (N[b] < 3 || N[b] > SAC_COT_MAX_N) return SAC_COT_E_SIZE;
  }
  try {
    ctx->pairs.clear();
    ctx->pairs.resize(ctx->keep_debug ? static_cast<size_t>(B) : (B > 0 ? 1 : 0));
    for (int b = 0; b < B; ++b) {
      PairState& st = ctx->pairs[ctx->keep_debug ? b : 0];
      load_pair(st, src[b], dst[b], N[b], *params);
      run_pair(st, R + 9 * static_cast<size_t>(b), t + 3 * static_cast<size_t>(b), inliers + b);
    }
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
  return SAC_COT_OK;
}

int sac_cot_register(const float* src, const float* dst, int32_t N, const sac_cot_params* params,
                     float R[9], float t[3], int32_t* inliers) {
  static sac_cot_ctx global_ctx;
  if (!src || !dst || !R || !t || !inliers) return SAC_COT_E_NULL;
  const float* s[1] = {src};
  const float* d[1] = {dst};
  return sac_cot_register_batch(&global_ctx, s, d, &N, 1, params, R, t, inliers);
}

// ---- sharded single pair -------------------------------------------------------------
int sac_cot_sharded_phase1(sac_cot_ctx* ctx, const float* src, const float* dst, int32_t N,
                           const sac_cot_params* params, int32_t rank, int32_t world,
                           uint64_t* t_partial, uint64_t* cand) {
  if (!ctx || !src || !dst || !t_partial || !cand) return SAC_COT_E_NULL;
  if (N < 3 || N > SAC_COT_MAX_N) return SAC_COT_E_SIZE;
  if (world < 1 || rank < 0 || rank >= world) return SAC_COT_E_SIZE;
  if (int rc = check_params(params)) return rc;
  if (normalized(*params).compat_mode != SAC_COT_COMPAT_FIRST_ORDER) return SAC_COT_E_UNSUPPORTED;  // A2 needs every rank's counts
  try {
    PairState& st = ctx->sharded;
    load_pair(st, src, dst, N, *params);
    st.rank = rank;
    st.world = world;
    build_graph(st);
    count_triangles(st);
    std::memcpy(t_partial, st.t2.data(), sizeof(uint64_t) * static_cast<size_t>(N));
    std::vector<uint64_t> local;
    top_k_desc(st.edge_keys, static_cast<size_t>(params->num_edges), local);
    for (int k = 0; k < params->num_edges; ++k) cand[k] = k < static_cast<int>(local.size()) ? local[k] : 0;
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
  return SAC_COT_OK;
}

int sac_cot_sharded_phase2(sac_cot_ctx* ctx, const uint64_t* t_all, const uint64_t* cand_all,
                           uint64_t* best_key) {
  if (!ctx || !t_all || !cand_all || !best_key) return SAC_COT_E_NULL;
  PairState& st = ctx->sharded;
  if (st.N < 3) return SAC_COT_E_SIZE;
  try {
    const int N = st.N, world = st.world, Ke = st.prm.num_edges, m = st.prm.apex_per_edge;
    for (int i = 0; i < N; ++i) {
      uint64_t sum = 0;
      for (int g = 0; g < world; ++g) sum += t_all[static_cast<size_t>(g) * N + i];
      st.t_node[i] = static_cast<uint32_t>(sum / 2);
    }
    std::vector<uint64_t> all;
    for (size_t k = 0; k < static_cast<size_t>(world) * Ke; ++k)
      if (cand_all[k]) all.push_back(cand_all[k]);
    top_k_desc(all, static_cast<size_t>(Ke), st.top_edges);
    select_triangles(st);
    make_hypotheses(st);
    const int K = Ke * m;
    const int per = (K + world - 1) / world;
    const int h0 = std::min(K, st.rank * per), h1 = std::min(K, h0 + per);
    score_hypotheses(st, h0, h1);
    *best_key = st.best_key;
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
  return SAC_COT_OK;
}

int sac_cot_sharded_phase3(sac_cot_ctx* ctx, uint64_t best_key_global, float R[9], float t[3],
                           int32_t* inliers) {
  if (!ctx || !R || !t || !inliers) return SAC_COT_E_NULL;
  PairState& st = ctx->sharded;
  if (st.N < 3) return SAC_COT_E_SIZE;
  st.best_key = best_key_global;
  finalize(st, best_key_global, R, t, inliers);
  return SAC_COT_OK;
}

// ---- correspondence front end (DESIGN.md §2 "S-1") --------------------------
//   D_ij = sum_c (f_ic - g_jc)^2 in fp32: D = 0; for c ascending: e = f_ic - g_jc; D = fma(e, e, D)
//   nn[i] = argmin_j D_ij, ties -> lowest j (strict < while j ascends); corr = (xyz_src[i], xyz_dst[nn[i]])
int sac_cot_match_packed(sac_cot_ctx* ctx, const float* desc_src, const float* xyz_src, const int64_t* offs_src,
                         const float* desc_dst, const float* xyz_dst, const int64_t* offs_dst, int32_t B, int32_t dim,
                         int32_t* nn, float* corr_src, float* corr_dst, int32_t location) {
  if (!ctx || !offs_src || !offs_dst) return SAC_COT_E_NULL;
  if (location != SAC_COT_LOC_HOST) return SAC_COT_E_UNSUPPORTED;
  if (B < 0 || B > 65535) return SAC_COT_E_SIZE;
  if (dim < 1 || dim > SAC_COT_MAX_DESC_DIM) return SAC_COT_E_SIZE;
  if (B > 0 && (!desc_src || !xyz_src || !desc_dst || !xyz_dst || !nn || !corr_src || !corr_dst)) return SAC_COT_E_NULL;
  for (int b = 0; b < B; ++b) {
    const int64_t ns = offs_src[b + 1] - offs_src[b], nd = offs_dst[b + 1] - offs_dst[b];
    if (ns < 1 || nd < 1 || ns > SAC_COT_MAX_KEYPOINTS || nd > SAC_COT_MAX_KEYPOINTS) return SAC_COT_E_SIZE;
  }
  for (int b = 0; b < B; ++b) {
    const int64_t s0 = offs_src[b], d0 = offs_dst[b];
    const int ns = static_cast<int>(offs_src[b + 1] - s0), nd = static_cast<int>(offs_dst[b + 1] - d0);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int i = 0; i < ns; ++i) {
      const float* f = desc_src + static_cast<size_t>(s0 + i) * dim;
      float best = 0.0f;
      int arg = 0;
      for (int j = 0; j < nd; ++j) {
        const float* g = desc_dst + static_cast<size_t>(d0 + j) * dim;
        float D = 0.0f;
        for (int c = 0; c < dim; ++c) {
          const float e = f[c] - g[c];
          D = std::fmaf(e, e, D);
        }
        if (j == 0 || D < best) {
          best = D;
          arg = j;
        }
      }
      nn[s0 + i] = arg;
      for (int a = 0; a < 3; ++a) {
        corr_src[static_cast<size_t>(s0 + i) * 3 + a] = xyz_src[static_cast<size_t>(s0 + i) * 3 + a];
        corr_dst[static_cast<size_t>(s0 + i) * 3 + a] = xyz_dst[static_cast<size_t>(d0 + arg) * 3 + a];
      }
    }
  }
  return SAC_COT_OK;
}

int sac_cot_match_mutual(sac_cot_ctx* ctx, const int32_t* nn, const int32_t* nn_back, const float* corr_src,
                         const float* corr_dst, const int64_t* offs_src, const int64_t* offs_dst, int32_t B, float* out_src,
                         float* out_dst, int64_t* out_offsets, int32_t location) {
  if (!ctx || !offs_src || !offs_dst || !out_offsets) return SAC_COT_E_NULL;
  if (location != SAC_COT_LOC_HOST) return SAC_COT_E_UNSUPPORTED;
  if (B < 0 || B > 65535) return SAC_COT_E_SIZE;
  if (B > 0 && (!nn || !nn_back || !corr_src || !corr_dst || !out_src || !out_dst)) return SAC_COT_E_NULL;
  for (int b = 0; b < B; ++b) {
    const int64_t ns = offs_src[b + 1] - offs_src[b], nd = offs_dst[b + 1] - offs_dst[b];
    if (ns < 1 || nd < 1 || ns > SAC_COT_MAX_KEYPOINTS || nd > SAC_COT_MAX_KEYPOINTS) return SAC_COT_E_SIZE;
  }
  int64_t o = 0;
  for (int b = 0; b < B; ++b) {
    out_offsets[b] = o;
    const int64_t s0 = offs_src[b], d0 = offs_dst[b];
    const int64_t ns = offs_src[b + 1] - s0, nd = offs_dst[b + 1] - d0;
    for (int64_t i = 0; i < ns; ++i) {
      const int64_t j = nn[s0 + i];
      if (j < 0 || j >= nd || nn_back[d0 + j] != i) continue;
      for (int a = 0; a < 3; ++a) {
        out_src[o * 3 + a] = corr_src[(s0 + i) * 3 + a];
        out_dst[o * 3 + a] = corr_dst[(s0 + i) * 3 + a];
      }
      ++o;
    }
  }
  out_offsets[B] = o;
  return SAC_COT_OK;
}

int sac_cot_match(const float* desc_src, const float* xyz_src, int32_t Ns, const float* desc_dst, const float* xyz_dst,
                  int32_t Nd, int32_t dim, int32_t* nn, float* corr_src, float* corr_dst) {
  static sac_cot_ctx global_ctx;
  const int64_t os[2] = {0, Ns}, od[2] = {0, Nd};
  return sac_cot_match_packed(&global_ctx, desc_src, xyz_src, os, desc_dst, xyz_dst, od, 1, dim, nn, corr_src, corr_dst,
                              SAC_COT_LOC_HOST);
}

// ---- device groups: GPU library only -----------------------------------------------------
int sac_cot_group_create(sac_cot_group** out, const int32_t*, int32_t) {
  if (out) *out = nullptr;
  return out ? SAC_COT_E_UNSUPPORTED : SAC_COT_E_NULL;
}
int sac_cot_group_destroy(sac_cot_group*) { return SAC_COT_OK; }
int32_t sac_cot_group_size(const sac_cot_group*) { return 0; }
sac_cot_ctx* sac_cot_group_ctx(sac_cot_group*, int32_t) { return nullptr; }
int sac_cot_group_set(sac_cot_group*, const char*, int64_t) { return SAC_COT_E_UNSUPPORTED; }
int sac_cot_group_register_packed(sac_cot_group*, const float*, const float*, const int64_t*, int32_t,
                                  const sac_cot_params*, float*, float*, int32_t*) {
  return SAC_COT_E_UNSUPPORTED;
}

// ---- in-library collectives: the oracle is a single process; it has no communicator ---------
int sac_cot_comm_unique_id(void* id_out) { return id_out ? SAC_COT_E_UNSUPPORTED : SAC_COT_E_NULL; }
int sac_cot_ctx_comm_init(sac_cot_ctx* ctx, const void* id, int32_t, int32_t) {
  return (ctx && id) ? SAC_COT_E_UNSUPPORTED : SAC_COT_E_NULL;
}
int sac_cot_ctx_set_comm(sac_cot_ctx* ctx, void*, int32_t, int32_t) { return ctx ? SAC_COT_E_UNSUPPORTED : SAC_COT_E_NULL; }

// world = 1: the three parts back to back (the partition then holds every edge and every hypothesis)
int sac_cot_register_sharded(sac_cot_ctx* ctx, const float* src, const float* dst, int32_t N,
                             const sac_cot_params* params, float R[9], float t[3], int32_t* inliers,
                             int32_t location) {
  if (!ctx || !src || !dst || !R || !t || !inliers) return SAC_COT_E_NULL;
  if (location != SAC_COT_LOC_HOST) return SAC_COT_E_UNSUPPORTED;
  if (N < 3 || N > SAC_COT_MAX_N) return SAC_COT_E_SIZE;
  if (int rc = check_params(params)) return rc;
  try {
    std::vector<uint64_t> t_partial(static_cast<size_t>(N)), cand(static_cast<size_t>(params->num_edges));
    if (int rc = sac_cot_sharded_phase1(ctx, src, dst, N, params, 0, 1, t_partial.data(), cand.data())) return rc;
    uint64_t best = 0;
    if (int rc = sac_cot_sharded_phase2(ctx, t_partial.data(), cand.data(), &best)) return rc;
    return sac_cot_sharded_phase3(ctx, best, R, t, inliers);
  } catch (const std::bad_alloc&) {
    return SAC_COT_E_NOMEM;
  }
}

// ---- debug getter -------------------------------------------------------------------
int sac_cot_debug_get(sac_cot_ctx* ctx, int32_t pair, int32_t which, void* out, size_t cap,
                      size_t* written) {
  if (!ctx || !written) return SAC_COT_E_NULL;
  const PairState* st = nullptr;
  if (pair == -1) st = &ctx->sharded;  // the sharded call's state
  else if (pair >= 0 && static_cast<size_t>(pair) < ctx->pairs.size()) st = &ctx->pairs[pair];
  else return SAC_COT_E_WHICH;
  const void* p = nullptr;
  size_t bytes = 0;
  uint64_t scalar = 0;
  switch (which) {
    case SAC_COT_DBG_ADJ: p = st->adj.data(); bytes = st->adj.size() * 4; break;
    case SAC_COT_DBG_ADJ_FIRST: {
      const std::vector<uint32_t>& a = st->adj_first.empty() ? st->adj : st->adj_first;
      p = a.data();
      bytes = a.size() * 4;
      break;
    }
    case SAC_COT_DBG_T_NODE: p = st->t_node.data(); bytes = st->t_node.size() * 4; break;
    case SAC_COT_DBG_NUM_EDGES: scalar = st->edge_keys.size(); p = &scalar; bytes = 8; break;
    case SAC_COT_DBG_EDGE_KEYS: p = st->edge_keys.data(); bytes = st->edge_keys.size() * 8; break;
    case SAC_COT_DBG_TOP_EDGES: p = st->top_edges.data(); bytes = st->top_edges.size() * 8; break;
    case SAC_COT_DBG_TRIANGLES: p = st->tri.data(); bytes = st->tri.size() * 4; break;
    case SAC_COT_DBG_HYP_RT: p = st->hyp_rt.data(); bytes = st->hyp_rt.size() * 4; break;
    case SAC_COT_DBG_HYP_SCORE: p = st->hyp_key.data(); bytes = st->hyp_key.size() * 8; break;
    case SAC_COT_DBG_BEST_KEY: scalar = st->best_key; p = &scalar; bytes = 8; break;
    case SAC_COT_DBG_MASK: p = st->mask.data(); bytes = st->mask.size() * 4; break;
    case SAC_COT_DBG_HIST: p = st->hist.data(); bytes = st->hist.size() * 4; break;
    default: return SAC_COT_E_WHICH;
  }
  *written = bytes;
  if (bytes > cap) return SAC_COT_E_CAPACITY;
  if (bytes && !out) return SAC_COT_E_NULL;
  if (bytes) std::memcpy(out, p, bytes);
  return SAC_COT_OK;
}

const char* sac_cot_strerror(int status) {
  switch (status) {
    case SAC_COT_OK: return "ok";
    case SAC_COT_E_NULL: return "null pointer argument";
    case SAC_COT_E_SIZE: return "size out of range (3 <= N <= 65535, B >= 0)";
    case SAC_COT_E_PARAMS: return "invalid sac_cot_params";
    case SAC_COT_E_NODEVICE: return "no usable CUDA device";
    case SAC_COT_E_UNSUPPORTED: return "not supported by this implementation";
    case SAC_COT_E_WHICH: return "unknown selector / index";
    case SAC_COT_E_CAPACITY: return "output buffer too small";
    case SAC_COT_E_NOMEM: return "out of memory";
    case SAC_COT_E_COMM: return "no communicator (the oracle is a single process)";
    default: return "unknown status";
  }
}

const char* sac_cot_version(void) { return "sac-cot-b200 0.1 (oracle, from-paper CPU restatement)"; }

}  // extern "C"

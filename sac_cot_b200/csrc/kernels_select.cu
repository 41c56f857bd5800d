// kernels_select.cu — S3, guided selection of the top-ranked compatibility triangles
// (SURVEY.md §8a row S3).
//
//  (1) edges ordered by key = T<<32 | (0xFFFF-i)<<16 | (0xFFFF-j) descending, i.e. (T desc, i asc,
//      j asc); the first K_e are kept.  Exact top-K_e without sorting E keys:
//        a. the triangle kernel left a 4096-bin histogram of the top digit T>>4;
//        b. select_scatter finds the threshold digit d* (largest d with #{digit >= d} >= K_e),
//           appends every key with digit > d* to the selected list and every key with digit == d*
//           to a tie list (or flags an in-place scan if that bucket exceeds the tie list);
//        c. select_final (one CTA per pair) radix-selects, over the remaining 36 key bits, the
//           exact key threshold among the ties (keys are unique, so the r-th largest is unique),
//           appends the ties >= threshold, and bitonic-sorts the <= 4096 selected keys.
//      The result does not depend on the (atomic, unordered) placement of keys in any list.
//  (2) select_apex: one warp per selected edge (i,j) enumerates k in N(i) ∩ N(j) and keeps the
//      m best by (t_k desc, k asc); hypothesis id h = r*m + q.
#include "common.cuh"

#include <algorithm>

namespace saccot {

constexpr int kSelThreads = 256;

__device__ __forceinline__ unsigned int key_digit(unsigned long long key) {
  return static_cast<unsigned int>(key >> 36);  // T >> 4 (T < 65536 => digit < 4096)
}

// Threshold digit from the pair's histogram.  All threads of a 256-thread CTA call this.
// Thread t owns bins [4095-16t-15, 4095-16t] (descending order => thread 0 owns the top bins).
__device__ void find_threshold_digit(const uint32_t* __restrict__ histp, unsigned int Ke, unsigned int* s_scan,
                                     unsigned int* s_out /* [3]: dstar, n_above, n_tie */) {
  const int t = threadIdx.x;
  unsigned int bins[16];
  unsigned int mine = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    bins[k] = histp[kHistBins - 1 - (16 * t + k)];  // descending digits
    mine += bins[k];
  }
  s_scan[t] = mine;
  __syncthreads();
  for (int o = 1; o < kSelThreads; o <<= 1) {
    const unsigned int add = t >= o ? s_scan[t - o] : 0u;
    __syncthreads();
    s_scan[t] += add;
    __syncthreads();
  }
  const unsigned int incl = s_scan[t];
  const unsigned int excl = incl - mine;
  const unsigned int total = s_scan[kSelThreads - 1];
  if (total < Ke) {
    // fewer edges than K_e: everything is selected; treat digit 0 as the tie bucket
    if (t == kSelThreads - 1) {
      s_out[0] = 0;
      s_out[1] = total - bins[15];
      s_out[2] = bins[15];
    }
  } else if (excl < Ke && incl >= Ke) {
    unsigned int cum = excl;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (cum < Ke && cum + bins[k] >= Ke) {
        s_out[0] = static_cast<unsigned int>(kHistBins - 1 - (16 * t + k));
        s_out[1] = cum;
        s_out[2] = bins[k];
      }
      cum += bins[k];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSelThreads) select_scatter_kernel(
    PairDev* __restrict__ state, const ChunkDev* __restrict__ chunk, const unsigned long long* __restrict__ keys,
    const uint32_t* __restrict__ hist, unsigned long long* __restrict__ sel, unsigned long long* __restrict__ tie,
    int Ke) {
  if (chunk->overflow) return;
  const int pair = blockIdx.y;
  __shared__ unsigned int s_scan[kSelThreads];
  __shared__ unsigned int s_out[3];
  find_threshold_digit(hist + static_cast<size_t>(pair) * kHistBins, static_cast<unsigned int>(Ke), s_scan, s_out);
  const unsigned int dstar = s_out[0], n_above = s_out[1], n_tie = s_out[2];
  const bool inplace = n_tie > static_cast<unsigned int>(kTieCap);
  PairDev* st = state + pair;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->dstar = dstar;
    st->n_above = n_above;
    st->n_tie = n_tie;
    st->tie_inplace = inplace ? 1u : 0u;
  }
  const unsigned long long E = st->key_count;
  const unsigned long long* kp = keys + st->key_base;
  unsigned long long* selp = sel + static_cast<size_t>(pair) * Ke;
  unsigned long long* tiep = tie + static_cast<size_t>(pair) * kTieCap;
  for (unsigned long long idx = static_cast<unsigned long long>(blockIdx.x) * kSelThreads + threadIdx.x; idx < E;
       idx += static_cast<unsigned long long>(gridDim.x) * kSelThreads) {
    const unsigned long long key = kp[idx];
    const unsigned int dg = key_digit(key);
    if (dg > dstar) {
      const unsigned int pos = atomicAdd(&st->sel_count, 1u);
      selp[pos] = key;  // pos < n_above < K_e by construction
    } else if (dg == dstar && !inplace) {
      const unsigned int pos = atomicAdd(&st->tie_count, 1u);
      tiep[pos] = key;  // pos < n_tie <= kTieCap
    }
  }
}

// One CTA (1024 threads) per pair.
__global__ void __launch_bounds__(1024) select_final_kernel(
    PairDev* __restrict__ state, const ChunkDev* __restrict__ chunk, const unsigned long long* __restrict__ keys,
    unsigned long long* __restrict__ sel, const unsigned long long* __restrict__ tie,
    unsigned long long* __restrict__ top, int Ke) {
  const int pair = blockIdx.x;
  unsigned long long* topp = top + static_cast<size_t>(pair) * Ke;
  if (chunk->overflow) {
    // key pool too small: the triangle kernels did nothing and `top` still holds whatever the arena held.  Every
    // slot becomes "unused" (key 0), so the apex / Kabsch / scoring / refit kernels that follow touch nothing outside
    // the pair's buffers and the chunk's outputs read R = I, t = 0, inliers = 0 until the re-run replaces them.
    for (int k = threadIdx.x; k < Ke; k += 1024) topp[k] = 0ull;
    return;
  }
  PairDev* st = state + pair;
  __shared__ unsigned long long sbuf[kMaxEdges];
  __shared__ unsigned int h256[256];
  __shared__ unsigned long long s_pref;
  __shared__ unsigned int s_need;

  const int t = threadIdx.x;
  const unsigned int dstar = st->dstar, n_above = st->n_above, n_tie = st->n_tie;
  const bool inplace = st->tie_inplace != 0;
  const unsigned long long* srcp = inplace ? keys + st->key_base : tie + static_cast<size_t>(pair) * kTieCap;
  const unsigned long long n_src = inplace ? st->key_count : static_cast<unsigned long long>(n_tie);
  unsigned int need = static_cast<unsigned int>(Ke) - n_above;  // n_above < K_e
  if (need > n_tie) need = n_tie;
  unsigned long long* selp = sel + static_cast<size_t>(pair) * Ke;

  unsigned long long theta = static_cast<unsigned long long>(dstar) << 36;  // take every tie by default
  if (need < n_tie) {
    // radix select of the `need`-th largest key among {key : digit == dstar} over bits [35:0]
    if (t == 0) {
      s_pref = static_cast<unsigned long long>(dstar) << 36;
      s_need = need;
    }
    __syncthreads();
    int known = 36;  // bits [63:known] of the threshold are fixed in s_pref
    while (known > 0) {
      const int width = known >= 8 ? 8 : known;
      const int shift = known - width;
      if (t < 256) h256[t] = 0;
      __syncthreads();
      const unsigned long long pref = s_pref;
      for (unsigned long long idx = t; idx < n_src; idx += 1024) {
        const unsigned long long key = srcp[idx];
        if ((key >> known) == (pref >> known)) atomicAdd(&h256[(key >> shift) & ((1u << width) - 1u)], 1u);
      }
      __syncthreads();
      if (t == 0) {
        unsigned int remaining = s_need, cum = 0;
        int dsel = 0;
        for (int dgt = (1 << width) - 1; dgt >= 0; --dgt) {
          if (cum + h256[dgt] >= remaining) {
            dsel = dgt;
            break;
          }
          cum += h256[dgt];
        }
        s_need = remaining - cum;  // still needed inside the chosen bucket (>= 1)
        s_pref = pref | (static_cast<unsigned long long>(dsel) << shift);
      }
      __syncthreads();
      known = shift;
    }
    theta = s_pref;  // exact key of the `need`-th largest tie
  }
  // append the ties at or above the threshold
  if (need > 0) {
    for (unsigned long long idx = t; idx < n_src; idx += 1024) {
      const unsigned long long key = srcp[idx];
      if (key_digit(key) == dstar && key >= theta) {
        const unsigned int pos = atomicAdd(&st->sel_count, 1u);
        if (pos < static_cast<unsigned int>(Ke)) selp[pos] = key;
      }
    }
  }
  __syncthreads();
  const unsigned int n_sel = n_above + need;
  // bitonic sort (descending) of the selected keys, zero padded to a power of two >= Ke
  int P = 1;
  while (P < Ke) P <<= 1;
  for (int k = t; k < P; k += 1024) sbuf[k] = static_cast<unsigned int>(k) < n_sel ? selp[k] : 0ull;
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int strd = size >> 1; strd > 0; strd >>= 1) {
      for (int k = t; k < P; k += 1024) {
        const int partner = k ^ strd;
        if (partner > k) {
          const bool desc = (k & size) == 0;
          const unsigned long long a = sbuf[k], b = sbuf[partner];
          if (desc ? (a < b) : (a > b)) {
            sbuf[k] = b;
            sbuf[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int k = t; k < Ke; k += 1024) topp[k] = sbuf[k];
  if (t == 0) st->n_sel = n_sel;
}

int launch_select_edges(const LaunchCtx& lc, int pairs, PairDev* d_state, const ChunkDev* d_chunk,
                        const unsigned long long* d_keys, const uint32_t* d_hist, unsigned long long* d_sel,
                        unsigned long long* d_tie, unsigned long long* d_top, int Ke) {
  int gx = (4 * lc.sm_count + pairs - 1) / pairs;
  if (gx < 16) gx = 16;
  if (gx > 4 * lc.sm_count) gx = 4 * lc.sm_count;
  select_scatter_kernel<<<dim3(gx, pairs), kSelThreads, 0, lc.stream>>>(d_state, d_chunk, d_keys, d_hist, d_sel,
                                                                        d_tie, Ke);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  select_final_kernel<<<pairs, 1024, 0, lc.stream>>>(d_state, d_chunk, d_keys, d_sel, d_tie, d_top, Ke);
  e = cudaGetLastError();
  return e == cudaSuccess ? 2 : -static_cast<int>(e);
}

// ------------------------------------------------------------------------------------------
// Apex selection: one warp per selected edge.
// ------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(128) select_apex_kernel(const PairDesc* __restrict__ descs,
                                                          const uint32_t* __restrict__ adj,
                                                          const unsigned long long* __restrict__ t2,
                                                          const unsigned long long* __restrict__ top,
                                                          int32_t* __restrict__ tri, int Ke, int m) {
  const int pair = blockIdx.y;
  const PairDesc d = descs[pair];
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= Ke) return;
  int32_t* out = tri + (static_cast<size_t>(pair) * Ke + r) * m * 3;
  const unsigned long long ekey = top[static_cast<size_t>(pair) * Ke + r];
  if (ekey == 0) {  // fewer than K_e edges: slot unused
    for (int k = lane; k < m * 3; k += 32) out[k] = -1;
    return;
  }
  const int i = static_cast<int>(0xFFFFu - static_cast<unsigned int>((ekey >> 16) & 0xFFFFu));
  const int j = static_cast<int>(0xFFFFu - static_cast<unsigned int>(ekey & 0xFFFFu));
  const uint32_t* ri = adj + d.adj_off + static_cast<size_t>(i) * d.stride;
  const uint32_t* rj = adj + d.adj_off + static_cast<size_t>(j) * d.stride;
  const unsigned long long* t2p = t2 + d.node_off;

  // per-lane best-M candidates, sorted descending; candidate key = t_k << 32 | (0xFFFFFFFF - k)
  unsigned long long best[M];
#pragma unroll
  for (int q = 0; q < M; ++q) best[q] = 0ull;
  for (int w = lane; w < d.stride; w += 32) {
    uint32_t bits = ri[w] & rj[w];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const unsigned int k = static_cast<unsigned int>(w * 32 + b);
      unsigned long long c = ((t2p[k] >> 1) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - k);
      if (c > best[M - 1]) {  // most candidates fail this test once the array has warmed up
#pragma unroll
        for (int q = 0; q < M; ++q) {
          if (c > best[q]) {
            const unsigned long long tmp = best[q];
            best[q] = c;
            c = tmp;
          }
        }
      }
    }
  }
  // warp merge: m rounds of "global max, owner pops"
  for (int q = 0; q < m; ++q) {
    const unsigned long long mine = best[0];
    const unsigned long long wmax = warp_max_u64(mine);
    if (wmax == 0ull) {
      if (lane == 0) { out[q * 3 + 0] = -1; out[q * 3 + 1] = -1; out[q * 3 + 2] = -1; }
      continue;
    }
    // candidate keys are unique (distinct k), so exactly one lane owns the maximum
    if (mine == wmax) {
      out[q * 3 + 0] = i;
      out[q * 3 + 1] = j;
      out[q * 3 + 2] = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned int>(wmax & 0xFFFFFFFFull));
#pragma unroll
      for (int s = 0; s < M - 1; ++s) best[s] = best[s + 1];
      best[M - 1] = 0ull;
    }
  }
}

// Node-pruned pairs (kernels_prune.cu), one warp: looks at the candidates of edge (ri, rj) that lie OUTSIDE the kept
// set (kb = the pair's kept-node bit mask, degs = its exact degrees, both staged in shared memory) and returns
// whether one of them has a count >= bound.  ts[k] of such a node is 0 (nothing known), an upper bound with
// kApexBoundFlag set, or the exact count (evaluated for an earlier edge of this CTA).  Bounds, cheapest first:
//   t_k <= deg_k (deg_k - 1) / 2                                   (every T_kn <= deg_k - 1)
//   t_k <= 1/2 sum_{n in N(k)} (min(deg_k, deg_n) - 1)             (one degree lookup per neighbour)
// and only a node whose second bound still reaches `bound` is evaluated exactly,
//   t_k = 1/2 sum_{n in N(k)} popc(row_k & row_n).
// Phase 1 is lane-parallel and free of warp-level primitives: every lane classifies the candidates in its own words
// (all loads independent: one memory latency per edge).  Phase 2 serves the few that need the second bound or the
// exact count, with warp-uniform control flow.  Only the PRUNED instance of the apex kernel carries this code.
constexpr uint32_t kApexBoundFlag = 0x80000000u;  // counts are < 2^31
constexpr int kApexPrunedWords = kNodePruneMaxNpad / 1024;  // adjacency words per lane: 10
__device__ __forceinline__ bool apex_outside_candidates(const uint32_t* __restrict__ adjp, int stride,
                                                        const uint32_t* __restrict__ ri, const uint32_t* __restrict__ rj,
                                                        const uint32_t* kb, const unsigned short* degs, uint32_t* ts,
                                                        uint32_t bound) {
  const int lane = threadIdx.x & 31;
  uint32_t need[kApexPrunedWords];
  bool upset = false;
#pragma unroll
  for (int s = 0; s < kApexPrunedWords; ++s) {
    const int w = 32 * s + lane;
    need[s] = w < stride ? (ri[w] & rj[w] & ~kb[w]) : 0u;
  }
#pragma unroll
  for (int s = 0; s < kApexPrunedWords; ++s) {
    uint32_t bits = need[s], keep = 0u;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1u;
      const int k = 32 * (32 * s + lane) + b;
      const uint32_t v = ts[k];
      if (v == 0u) {
        const uint32_t dk = degs[k];
        if (dk * (dk - 1u) / 2u >= bound) keep |= 1u << b;
      } else if (v & kApexBoundFlag) {
        if ((v & ~kApexBoundFlag) >= bound) keep |= 1u << b;
      } else {
        upset |= v >= bound;
      }
    }
    need[s] = keep;
  }
  upset = __any_sync(0xffffffffu, upset);
#pragma unroll
  for (int s = 0; s < kApexPrunedWords; ++s) {
    unsigned act = __ballot_sync(0xffffffffu, need[s] != 0u);
    while (act) {
      const int src = __ffs(act) - 1;
      act &= act - 1u;
      uint32_t bits = __shfl_sync(0xffffffffu, need[s], src);
      while (bits) {
        const int k = 32 * (32 * s + src) + __ffs(bits) - 1;
        bits &= bits - 1u;
        uint32_t v = ts[k];  // one shared-memory word: the same value in every lane (it may have moved on since phase 1)
        const uint32_t* rk = adjp + static_cast<size_t>(k) * stride;
        if (v == 0u) {
          const uint32_t dk = degs[k];
          uint32_t ub = 0;  // second bound, per lane: <= 10 words x 32 neighbours x 65534
          for (int x = lane; x < stride; x += 32) {
            uint32_t nb = rk[x];
            while (nb) {
              const int b2 = __ffs(nb) - 1;
              nb &= nb - 1u;
              ub += min(dk, static_cast<uint32_t>(degs[32 * x + b2])) - 1u;  // an edge: both degrees >= 1
            }
          }
          unsigned long long ub64 = ub;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) ub64 += __shfl_xor_sync(0xffffffffu, ub64, o);
          v = static_cast<uint32_t>(ub64 >> 1) | kApexBoundFlag;
          __syncwarp();
          if (lane == 0) ts[k] = v;  // another warp of the CTA may store this or the exact count: either is valid
          __syncwarp();
        }
        if (v & kApexBoundFlag) {
          if ((v & ~kApexBoundFlag) < bound) continue;
          unsigned long long sum = 0;  // exact count
          for (int w2 = 0; w2 < stride; ++w2) {
            uint32_t nb = rk[w2];  // warp-uniform
            while (nb) {
              const int b2 = __ffs(nb) - 1;
              nb &= nb - 1u;
              const uint32_t* rn = adjp + static_cast<size_t>(32 * w2 + b2) * stride;
              uint32_t c = 0;
              for (int x = lane; x < stride; x += 32) c += __popc(rk[x] & rn[x]);
              sum += c;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          v = static_cast<uint32_t>(sum >> 1);
          __syncwarp();
          if (lane == 0) ts[k] = v;
          __syncwarp();
        }
        upset |= v >= bound;
      }
    }
  }
  return upset;
}

// Staged variant (the one normally used): the pair's node counts t_k = t2_k / 2 are copied once per CTA into
// shared memory as 32-bit values (t_k <= (N-1)(N-2)/2 < 2^31) and a CTA of 32 warps walks many edges of ONE pair.
// The warp-per-edge kernel above looks every candidate up in global memory: 8 useful bytes per 32-byte L2
// sector, ~290 candidates per edge.  ncu showed that kernel (and a first staged version) bound by the ALU pipe
// (88 %): 1600 warp instructions per edge, almost all of them the 64-bit insertion network, run divergently for
// every candidate.  Hence two passes: a 32-bit maximum per lane first, the insertion only for candidates that
// reach the resulting lower bound.  Rows are read eight words per lane at a time (independent loads).
constexpr int kApexThreads = 1024;
constexpr int kApexRank = 256;   // nodes the rank list aims for
constexpr int kApexCap = 512;    // its capacity (ties at the cut)
constexpr int kApexMaxSmem = 200 * 1024;
template <int M, bool PRUNED>
__global__ void __launch_bounds__(kApexThreads) select_apex_staged_kernel(const PairDesc* __restrict__ descs,
                                                                          const uint32_t* __restrict__ adj,
                                                                          const unsigned long long* __restrict__ t2,
                                                                          const unsigned long long* __restrict__ top,
                                                                          int32_t* __restrict__ tri, int Ke, int m,
                                                                          int use_rank_list,
                                                                          const NodePlan* __restrict__ plan,
                                                                          const unsigned short* __restrict__ deg,
                                                                          const uint32_t* __restrict__ keptbits) {
  extern __shared__ uint32_t apex_ts[];  // [Npad] node counts, then the rank list
  __shared__ int s_cnt, s_len;
  __shared__ uint32_t s_max;
  const int pair = blockIdx.y;
  const PairDesc d = descs[pair];
  // Node-pruned pair (kernels_prune.cu): t2 holds the sums of the kept nodes only, the others read 0 and their
  // t_k is at most ub_rest.  A candidate apex is in a triangle with its edge, so its true t_k is >= 1: among the
  // candidates 0 means "not evaluated".
  // The PRUNED instance handles those pairs, the plain one all the others (both are launched when the chunk may
  // hold pruned pairs; a CTA of the wrong kind leaves at once).
  if ((plan != nullptr && plan[pair].pruned != 0u) != PRUNED) return;
  // PRUNED: the pair's degrees and kept-node bits follow the rank list in shared memory
  unsigned short* deg_s = reinterpret_cast<unsigned short*>(apex_ts + d.Npad + 2 * kApexCap);
  uint32_t* kb_s = reinterpret_cast<uint32_t*>(deg_s + d.Npad);
  if constexpr (PRUNED) {
    for (int k = threadIdx.x; k < d.Npad; k += kApexThreads) deg_s[k] = k < d.N ? deg[d.node_off + k] : static_cast<unsigned short>(0);
    for (int k = threadIdx.x; k < d.stride; k += kApexThreads) kb_s[k] = keptbits[d.mask_off + k];
  }
  const int lane = threadIdx.x & 31, tid = threadIdx.x;
  unsigned long long* rank_list = reinterpret_cast<unsigned long long*>(apex_ts + d.Npad);  // [kApexCap]
  {
    const unsigned long long* t2p = t2 + d.node_off;
    uint32_t mx = 0;
    for (int k = tid; k < d.Npad; k += kApexThreads) {
      const uint32_t v = static_cast<uint32_t>(t2p[k] >> 1);
      apex_ts[k] = v;
      mx = max(mx, v);
    }
    if (tid == 0) {
      s_max = 0u;
      s_len = 0;
    }
    for (int k = tid; k < kApexCap; k += kApexThreads) rank_list[k] = 0ull;
    __syncthreads();
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0 && mx) atomicMax(&s_max, mx);
  }
  __syncthreads();
  // ---- rank list: every node with t_k >= tcut, sorted by (t_k descending, k ascending) — the order apexes are
  //      ranked in.  It holds ALL nodes down to tcut, so the first m common neighbours found while walking it are
  //      exactly the m best ones; an edge that finds fewer than m takes the exhaustive path below. ----
  auto count_ge = [&](uint32_t v) {
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    for (int k = tid; k < d.Npad; k += kApexThreads) c += apex_ts[k] >= v ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    const int r = s_cnt;
    __syncthreads();
    return r;
  };
  uint32_t tcut = 1;
  if (!use_rank_list) tcut = 0xFFFFFFFFu;  // tests: empty list, every edge takes the exhaustive path
  else if (count_ge(1u) > kApexRank) {
    uint32_t lo = 1, hi = s_max + 1;  // count_ge(lo) >= kApexRank > count_ge(hi)
    while (hi - lo > 1) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (count_ge(mid) >= kApexRank) lo = mid;
      else hi = mid;
    }
    tcut = count_ge(lo) <= kApexCap ? lo : lo + 1;  // many ties at lo: keep only what is strictly above
  }
  for (int k = tid; k < d.Npad; k += kApexThreads) {
    const uint32_t v = apex_ts[k];
    if (v >= tcut)
      rank_list[atomicAdd(&s_len, 1)] = (static_cast<unsigned long long>(v) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned int>(k));
  }
  __syncthreads();
  const int nrank = s_len;
  // bitonic sort, descending, kApexCap slots (empty slots are 0 and sink to the end)
  for (int sz = 2; sz <= kApexCap; sz <<= 1) {
    for (int st = sz >> 1; st > 0; st >>= 1) {
      for (int idx = tid; idx < kApexCap / 2; idx += kApexThreads) {
        const int a = 2 * idx - (idx & (st - 1)), b = a + st;
        const unsigned long long x = rank_list[a], y = rank_list[b];
        const bool desc = (a & sz) == 0;
        if (desc ? x < y : x > y) {
          rank_list[a] = y;
          rank_list[b] = x;
        }
      }
      __syncthreads();
    }
  }
  const int wstep = gridDim.x * (kApexThreads >> 5);
  for (int r = blockIdx.x * (kApexThreads >> 5) + (threadIdx.x >> 5); r < Ke; r += wstep) {
    int32_t* out = tri + (static_cast<size_t>(pair) * Ke + r) * m * 3;
    const unsigned long long ekey = top[static_cast<size_t>(pair) * Ke + r];
    if (ekey == 0) {  // fewer than K_e edges: slot unused
      for (int k = lane; k < m * 3; k += 32) out[k] = -1;
      continue;
    }
    const int i = static_cast<int>(0xFFFFu - static_cast<unsigned int>((ekey >> 16) & 0xFFFFu));
    const int j = static_cast<int>(0xFFFFu - static_cast<unsigned int>(ekey & 0xFFFFu));
    const uint32_t* ri = adj + d.adj_off + static_cast<size_t>(i) * d.stride;
    const uint32_t* rj = adj + d.adj_off + static_cast<size_t>(j) * d.stride;
    // ---- fast path: walk the rank list, 32 nodes at a time ----
    int found = 0;
    uint32_t t_mth = 0u;  // count of the m-th hit
    for (int base = 0; base < nrank && found < m; base += 32) {
      const unsigned long long key = base + lane < nrank ? rank_list[base + lane] : 0ull;
      bool hit = false;
      const unsigned int k = 0xFFFFFFFFu - static_cast<unsigned int>(key & 0xFFFFFFFFull);
      if (key) hit = (((ri[k >> 5] & rj[k >> 5]) >> (k & 31)) & 1u) != 0u;
      const unsigned hm = __ballot_sync(0xffffffffu, hit);
      const int q = found + __popc(hm & ((1u << lane) - 1u));
      if (hit && q < m) {
        out[q * 3 + 0] = i;
        out[q * 3 + 1] = j;
        out[q * 3 + 2] = static_cast<int>(k);
      }
      const unsigned last = __ballot_sync(0xffffffffu, hit && q == m - 1);
      if (last) t_mth = __shfl_sync(0xffffffffu, static_cast<uint32_t>(key >> 32), __ffs(last) - 1);
      found += __popc(hm);
    }
    if constexpr (!PRUNED) {
      if (found >= m) continue;
    } else {
      // Pruned pair: the list ranks the kept nodes only.  The m hits stand iff no candidate OUTSIDE the kept set
      // reaches the m-th hit's count (a tie would go to the lower index: be strict).  Without m hits every unknown
      // candidate is evaluated (bound 0).
      const uint32_t bound = found >= m ? t_mth : 0u;
      const bool upset = apex_outside_candidates(adj + d.adj_off, d.stride, ri, rj, kb_s, deg_s, apex_ts, bound);
      if (found >= m && !upset) continue;
      // otherwise the exhaustive path below decides among the known candidates: with m hits, the ones left unknown
      // (count 0) lie strictly below the m-th hit and cannot be selected
    }
    // ---- exhaustive path (the rank list held fewer than m common neighbours of this edge) ----
    // Pass 1: per-lane maximum of t_k over the candidates (one IMNMX per candidate), then a lower bound L on the
    // m-th best count: the value at which the lane maxima, taken from the top, cover m lanes (each of them holds
    // a candidate >= L).  Every candidate is in a triangle with (i, j), so t_k >= 1 and 0 means "none".
    uint32_t lmax = 0;
    for (int w0 = 0; w0 < d.stride; w0 += 256) {
      uint32_t bw[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int w = w0 + 32 * c + lane;
        bw[c] = w < d.stride ? (ri[w] & rj[w]) : 0u;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t bits = bw[c];
        const uint32_t* tw = apex_ts + (w0 + 32 * c + lane) * 32;
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          uint32_t tk = tw[b];
          if (PRUNED && (tk & kApexBoundFlag)) tk = 0u;  // only an upper bound is known: below the m-th hit (see above)
          lmax = max(lmax, tk);
        }
      }
    }
    uint32_t L = 1;
    {
      uint32_t v = lmax;
      int got = 0;
      for (int q = 0; q < m && got < m; ++q) {
        const uint32_t wm = __reduce_max_sync(0xffffffffu, v);
        if (wm == 0) break;
        L = wm;
        got += __popc(__ballot_sync(0xffffffffu, v == wm));
        if (v == wm) v = 0;
      }
      if (got < m) L = 1;  // fewer than m lanes hold candidates: keep them all
    }
    // Pass 2: per-lane best-M among the few candidates with t_k >= L, sorted descending;
    // candidate key = t_k << 32 | (0xFFFFFFFF - k)
    unsigned long long best[M];
#pragma unroll
    for (int q = 0; q < M; ++q) best[q] = 0ull;
    for (int w0 = 0; w0 < d.stride; w0 += 256) {
      uint32_t bw[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int w = w0 + 32 * c + lane;
        bw[c] = w < d.stride ? (ri[w] & rj[w]) : 0u;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t bits = bw[c];
        const unsigned int kb = static_cast<unsigned int>((w0 + 32 * c + lane) * 32);
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const unsigned int k = kb + static_cast<unsigned int>(b);
          uint32_t tk = apex_ts[k];
          if (PRUNED && (tk & kApexBoundFlag)) tk = 0u;
          if (tk >= L) {
            unsigned long long cnd = (static_cast<unsigned long long>(tk) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - k);
#pragma unroll
            for (int q = 0; q < M; ++q) {
              if (cnd > best[q]) {
                const unsigned long long tmp = best[q];
                best[q] = cnd;
                cnd = tmp;
              }
            }
          }
        }
      }
    }
    // warp merge: m rounds of "global max, owner pops"
    for (int q = 0; q < m; ++q) {
      const unsigned long long mine = best[0];
      const unsigned long long wmax = warp_max_u64(mine);
      if (wmax == 0ull) {
        if (lane == 0) { out[q * 3 + 0] = -1; out[q * 3 + 1] = -1; out[q * 3 + 2] = -1; }
        continue;
      }
      // candidate keys are unique (distinct k), so exactly one lane owns the maximum
      if (mine == wmax) {
        out[q * 3 + 0] = i;
        out[q * 3 + 1] = j;
        out[q * 3 + 2] = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned int>(wmax & 0xFFFFFFFFull));
#pragma unroll
        for (int s = 0; s < M - 1; ++s) best[s] = best[s + 1];
        best[M - 1] = 0ull;
      }
    }
  }
}

template <int M>
static cudaError_t apex_configure_one() {
  cudaError_t e = cudaFuncSetAttribute(select_apex_staged_kernel<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kApexMaxSmem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(select_apex_staged_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kApexMaxSmem);
  return e;
}
int select_configure() {
  cudaError_t e = apex_configure_one<1>();
  if (e == cudaSuccess) e = apex_configure_one<2>();
  if (e == cudaSuccess) e = apex_configure_one<4>();
  if (e == cudaSuccess) e = apex_configure_one<8>();
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

int launch_select_apex(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_npad, const uint32_t* d_adj,
                       const unsigned long long* d_t2, const unsigned long long* d_top, int32_t* d_tri, int Ke, int m,
                       int apex_path, const NodePlan* d_plan, const unsigned short* d_deg, const uint32_t* d_keptbits) {
  // apex_path (tests): 0 = staged kernel with the rank list (default), 1 = staged kernel, exhaustive path only,
  // 2 = the global-lookup kernel that large N falls back to
  const size_t smem = static_cast<size_t>(max_npad) * 4 + kApexCap * 8;
  const int rl = apex_path == 0 ? 1 : 0;
  int launched = 1;
  if (smem <= static_cast<size_t>(kApexMaxSmem) && apex_path != 2) {
    // enough CTAs to fill the GPU (every CTA of a pair rebuilds the pair's rank list), at most one per 32 edges
    int split = (lc.sm_count + pairs - 1) / pairs;
    split = std::max(1, std::min(split, (Ke + 31) / 32));
    dim3 grid(split, pairs);
    auto go = [&](auto kernel_plain, auto kernel_pruned) {
      kernel_plain<<<grid, kApexThreads, smem, lc.stream>>>(d_desc, d_adj, d_t2, d_top, d_tri, Ke, m, rl, d_plan, d_deg, d_keptbits);
      if (d_plan) {
        const size_t smem_pruned = smem + static_cast<size_t>(max_npad) * 2 + static_cast<size_t>(max_npad / 32) * 4;
        kernel_pruned<<<grid, kApexThreads, smem_pruned, lc.stream>>>(d_desc, d_adj, d_t2, d_top, d_tri, Ke, m, rl, d_plan, d_deg, d_keptbits);
        launched = 2;
      }
    };
    if (m <= 1) go(select_apex_staged_kernel<1, false>, select_apex_staged_kernel<1, true>);
    else if (m <= 2) go(select_apex_staged_kernel<2, false>, select_apex_staged_kernel<2, true>);
    else if (m <= 4) go(select_apex_staged_kernel<4, false>, select_apex_staged_kernel<4, true>);
    else go(select_apex_staged_kernel<8, false>, select_apex_staged_kernel<8, true>);
  } else {  // node counts do not fit shared memory (N > 51200): look them up in global memory
    dim3 grid((Ke + 3) / 4, pairs);
    if (m <= 1) select_apex_kernel<1><<<grid, 128, 0, lc.stream>>>(d_desc, d_adj, d_t2, d_top, d_tri, Ke, m);
    else if (m <= 2) select_apex_kernel<2><<<grid, 128, 0, lc.stream>>>(d_desc, d_adj, d_t2, d_top, d_tri, Ke, m);
    else if (m <= 4) select_apex_kernel<4><<<grid, 128, 0, lc.stream>>>(d_desc, d_adj, d_t2, d_top, d_tri, Ke, m);
    else select_apex_kernel<8><<<grid, 128, 0, lc.stream>>>(d_desc, d_adj, d_t2, d_top, d_tri, Ke, m);
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? launched : -static_cast<int>(e);
}

// ------------------------------------------------------------------------------------------
// Sharded single pair: the two device-side ends of exchange #1.
//   record of one rank (u64 units): [0, Npad) partial node sums | [Npad, Npad + Ke) its top-K_e edge keys, descending,
//   0-padded | overflow flag | key-pool demand.  shard_pack builds the record next to the pipeline's buffers (one
//   contiguous send buffer for ncclAllGather); shard_merge turns the `world` gathered records into the pair's node
//   sums and the global top-K_e list:
//     - node sums add up (integers: order-free);
//     - the global top-K_e restricted to a rank lies inside that rank's top-K_e, so the top-K_e of the union is
//       exact.  Keys are unique (an edge is owned by exactly one rank) and every list is sorted, so the position of a
//       key in the merged order is the number of keys above it, found by one binary search per list: no sort, no
//       atomics, no dependence on the gather order.
//   If any rank ran out of key-pool space the merged list stays empty (every slot unused), the chunk is flagged on
//   EVERY rank alike and the sticky record takes the largest demand: all ranks then grow and re-run in step.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) shard_pack_kernel(const unsigned long long* __restrict__ t2,
                                                         const unsigned long long* __restrict__ top,
                                                         const ChunkDev* __restrict__ chunk,
                                                         unsigned long long* __restrict__ rec, int npad, int Ke) {
  const int n = npad + Ke;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    rec[k] = k < npad ? t2[k] : top[k - npad];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    rec[n] = chunk->overflow;
    rec[n + 1] = chunk->total_edges;
  }
}

__global__ void __launch_bounds__(256) shard_merge_kernel(const unsigned long long* __restrict__ recs, int world,
                                                          int npad, int Ke, unsigned long long* __restrict__ t2,
                                                          unsigned long long* __restrict__ top /* zeroed */,
                                                          PairDev* __restrict__ state, ChunkDev* __restrict__ chunk,
                                                          StickyDev* __restrict__ sticky,
                                                          unsigned long long* __restrict__ summary /* [2] */) {
  const size_t rec_len = static_cast<size_t>(npad) + Ke + 2;
  unsigned long long any = 0, demand = 0;
  for (int g = 0; g < world; ++g) {
    const unsigned long long* r = recs + g * rec_len + npad + Ke;
    any |= r[0];
    demand = r[1] > demand ? r[1] : demand;
  }
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstep = gridDim.x * blockDim.x;
  for (int i = gtid; i < npad; i += gstep) {
    unsigned long long s = 0;
    for (int g = 0; g < world; ++g) s += recs[g * rec_len + i];
    t2[i] = s;
  }
  // number of keys of a (descending, 0-padded) list that are larger than `key`
  auto count_above = [&](const unsigned long long* list, unsigned long long key) {
    int lo = 0, hi = Ke;  // list[lo - 1] > key >= list[hi]
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (list[mid] > key) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  };
  if (!any) {
    for (int c = gtid; c < world * Ke; c += gstep) {
      const unsigned long long key = recs[(c / Ke) * rec_len + npad + (c % Ke)];
      if (key == 0ull) continue;
      int pos = 0;
      for (int g = 0; g < world && pos < Ke; ++g) pos += count_above(recs + g * rec_len + npad, key);
      if (pos < Ke) top[pos] = key;
    }
  }
  if (gtid == 0) {
    int total = 0;
    for (int g = 0; g < world; ++g) total += count_above(recs + g * rec_len + npad, 0ull);
    state[0].n_sel = any ? 0u : static_cast<uint32_t>(total < Ke ? total : Ke);
    summary[0] = any;
    summary[1] = demand;
    if (any) {
      chunk->overflow = 1u;
      chunk->total_edges = demand;
      if (sticky) {
        atomicMax(&sticky->max_total_edges, demand);
        atomicAdd(&sticky->overflow_count, 1u);
      }
    }
  }
}

int launch_shard_pack(const LaunchCtx& lc, const unsigned long long* d_t2, const unsigned long long* d_top,
                      const ChunkDev* d_chunk, unsigned long long* d_rec, int npad, int Ke) {
  const int blocks = std::min(lc.sm_count, (npad + Ke + 255) / 256);
  shard_pack_kernel<<<blocks, 256, 0, lc.stream>>>(d_t2, d_top, d_chunk, d_rec, npad, Ke);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

int launch_shard_merge(const LaunchCtx& lc, const unsigned long long* d_recs, int world, int npad, int Ke,
                       unsigned long long* d_t2, unsigned long long* d_top, PairDev* d_state, ChunkDev* d_chunk,
                       StickyDev* d_sticky, unsigned long long* d_summary) {
  const int work = std::max(npad, world * Ke);
  const int blocks = std::min(lc.sm_count, (work + 255) / 256);
  shard_merge_kernel<<<blocks, 256, 0, lc.stream>>>(d_recs, world, npad, Ke, d_t2, d_top, d_state, d_chunk, d_sticky,
                                                    d_summary);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

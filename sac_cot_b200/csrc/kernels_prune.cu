// kernels_prune.cu — exact node pruning of S2 (tensor-core path).
//
// tri_theta_kernel certifies, per pair, a lower bound theta0 on the K_e-th largest edge count: at least K_e edges with
// T >= theta0 exist, so no edge below theta0 can be selected.  An edge (i, j) has T_ij = |N(i) ∩ N(j)| <= deg_i - 1
// (j is a neighbour of i but not of itself) and likewise <= deg_j - 1.  Hence every selectable edge joins two nodes
// of degree >= theta0 + 1 — the KEPT nodes.  When they are few (a 5 % inlier clique in a sparse outlier graph keeps
// exactly the inliers), S2 shrinks from the dense N x N x N contraction to
//     for every kept node k, for every neighbour j of k:   T_kj = popc(row_k & row_j)
// which yields (a) the exact key of every selectable edge (j kept too, j > k, T >= theta0) and (b) the exact node sum
// t2_k = sum_j T_kj of every kept node.  Node sums of the other nodes are NOT computed; S3's apex ranking needs them
// only when a candidate apex outside the kept set could displace a kept one, which the apex kernel decides exactly
// from the bound t_k <= D (D - 1) / 2, D = largest degree outside the kept set, and otherwise evaluates on demand
// (kernels_select.cu).  Every result of the pipeline is bit-identical to the unpruned one.
//
//   node_degree_kernel   exact degrees (one warp per row)
//   node_plan_kernel     per pair: kept list (ascending), D, and the decision (cost model below)
//   tile_compact_kernel  tile list of the tensor-core kernel without the tiles of pruned pairs
//   triangles_kept_kernel the loop above on the CUDA cores: a warp holds 4 kept rows in registers, all rows of the pair
//                        stream through a shared-memory ring in batches of 32 (bulk copies, mbarriers)
//   kept_panel_kernel    compact copy of the kept nodes' K-panel records: long pairs in chunks of >= 4 run the same
//                        loop as tiles "kept nodes x all columns" of the tensor-core kernel's RECT instance
//                        (kernels_triangles_mma.cu), whose tile list tile_compact_kernel also writes
#include "common.cuh"

#include <algorithm>

namespace saccot {

namespace {

constexpr int kKeptThreads = 512;   // 16 warps
constexpr int kKeptRows = 4;        // kept rows per warp
constexpr int kKeptRowsPerCta = (kKeptThreads / 32) * kKeptRows;  // 64

// A pair whose AVERAGE degree already reaches theta0 + 1 keeps most of its nodes: not worth a look (the degree pass
// reads the whole adjacency).  Skipping a pair is always exact — it then takes the tensor-core kernel.
__device__ __forceinline__ bool prune_hopeless(const PairDesc& d, const PairDev& st, uint32_t theta_raw, int force) {
  const uint32_t th = theta_raw & 0x7FFFFFFFu;
  if (th == 0u || d.Npad > kNodePruneMaxNpad) return true;
  return force < 2 && 2ull * st.all_edges >= static_cast<unsigned long long>(th + 1u) * static_cast<unsigned long long>(d.N);
}

__global__ void __launch_bounds__(256) node_degree_kernel(const PairDesc* __restrict__ descs,
                                                          const uint32_t* __restrict__ adj,
                                                          const ChunkDev* __restrict__ chunk,
                                                          const PairDev* __restrict__ state,
                                                          const uint32_t* __restrict__ theta,
                                                          unsigned short* __restrict__ deg, int force) {
  if (chunk->overflow || !chunk->use_tensor) return;
  const PairDesc d = descs[blockIdx.y];
  if (prune_hopeless(d, state[blockIdx.y], theta[blockIdx.y], force)) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint4* base = reinterpret_cast<const uint4*>(adj + d.adj_off);
  const int q4 = d.stride >> 2;  // uint4 per row
  unsigned short* out = deg + d.node_off;
  // four rows per warp and step: the pass is bound by HBM (the whole adjacency is read once), so every lane keeps
  // several independent 16-byte loads in flight
  constexpr int kRows = 4;
  for (int r0 = (blockIdx.x * 8 + warp) * kRows; r0 < d.N; r0 += gridDim.x * 8 * kRows) {
    int c[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) c[k] = 0;
    for (int q = lane; q < q4; q += 32) {
      uint4 w[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k)
        w[k] = r0 + k < d.N ? __ldg(base + static_cast<size_t>(r0 + k) * q4 + q) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int k = 0; k < kRows; ++k) c[k] += __popc(w[k].x) + __popc(w[k].y) + __popc(w[k].z) + __popc(w[k].w);
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int v = __reduce_add_sync(0xffffffffu, c[k]);
      if (lane == 0 && r0 + k < d.N) out[r0 + k] = static_cast<unsigned short>(v);  // deg <= N - 1 <= 65534
    }
  }
}

// One CTA per pair.  Decision: the kept-row kernel costs ~ sum of the kept degrees x row length, the tensor-core
// kernel ~ Npad^3, so the pair is pruned if (sum of kept degrees) x cost <= Npad^2 (cost: api.cu).
__global__ void __launch_bounds__(1024) node_plan_kernel(const PairDesc* __restrict__ descs,
                                                         const ChunkDev* __restrict__ chunk,
                                                         StickyDev* __restrict__ sticky,
                                                         const PairDev* __restrict__ state,
                                                         const uint32_t* __restrict__ theta,
                                                         const unsigned short* __restrict__ deg,
                                                         NodePlan* __restrict__ plan, unsigned short* __restrict__ kept,
                                                         uint32_t* __restrict__ keptbits, int* __restrict__ n_pruned,
                                                         int cost, int force, int rect) {
  const int pair = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  NodePlan* pl = plan + pair;
  if (chunk->overflow || !chunk->use_tensor) return;  // plan is zeroed with the chunk state: not pruned
  const PairDesc d = descs[pair];
  if (prune_hopeless(d, state[pair], theta[pair], force)) return;
  const uint32_t min_deg = (theta[pair] & 0x7FFFFFFFu) + 1u;
  const unsigned short* dg = deg + d.node_off;
  __shared__ int s_keep, s_dmax;
  __shared__ unsigned long long s_sum;
  __shared__ int s_wc[32];
  if (t == 0) {
    s_keep = 0;
    s_dmax = 0;
    s_sum = 0ull;
  }
  __syncthreads();
  // every warp takes a contiguous segment of the rows (a multiple of 32): counts first, slots after a prefix
  const int seg = ((d.N + 31) / 32 + 31) & ~31;
  const int r0 = warp * seg, r1 = min(d.N, r0 + seg);
  int nk = 0, dmax = 0;
  unsigned long long sum = 0;
  for (int i0 = r0; i0 < r1; i0 += 32) {
    const int i = i0 + lane;
    const uint32_t v = i < r1 ? dg[i] : 0u;
    const bool k = i < r1 && v >= min_deg;
    nk += __popc(__ballot_sync(0xffffffffu, k));
    if (k) sum += v;
    else dmax = max(dmax, static_cast<int>(v));
  }
  dmax = __reduce_max_sync(0xffffffffu, dmax);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) {
    s_wc[warp] = nk;
    if (nk) atomicAdd(&s_keep, nk);
    if (dmax) atomicMax(&s_dmax, dmax);
    if (sum) atomicAdd(&s_sum, sum);
  }
  __syncthreads();
  const int n_keep = s_keep;
  const unsigned long long npad2 = static_cast<unsigned long long>(d.Npad) * d.Npad;
  const bool ok = n_keep >= 2 && n_keep <= kNodeKeepMax && (force >= 2 || s_sum * static_cast<unsigned long long>(cost) <= npad2);
  if (!ok) return;
  int before = 0;
  for (int w = 0; w < warp; ++w) before += s_wc[w];
  unsigned short* kp = kept + static_cast<size_t>(pair) * kNodeKeepMax;
  uint32_t* kbp = keptbits + d.mask_off;  // bit i & 31 of word i >> 5: node i is kept (segments start at multiples of 32)
  for (int i0 = r0; i0 < r1; i0 += 32) {
    const int i = i0 + lane;
    const bool k = i < r1 && dg[i] >= min_deg;
    const unsigned m = __ballot_sync(0xffffffffu, k);
    if (k) kp[before + __popc(m & ((1u << lane) - 1u))] = static_cast<unsigned short>(i);
    if (lane == 0) kbp[i0 >> 5] = m;
    before += __popc(m);
  }
  if (t == 0) {
    const unsigned long long D = static_cast<unsigned long long>(s_dmax);
    pl->n_keep = static_cast<uint32_t>(n_keep);
    pl->ub_rest = static_cast<uint32_t>(D ? D * (D - 1ull) / 2ull : 0ull);  // D <= 65534: < 2^31
    pl->min_deg = min_deg;
    // 2: the kept rows go through the tensor-core kernel's RECT instance (long rows, few enough kept nodes for its
    // compact panel copy); 1: through the kept-row POPC kernel
    pl->pruned = (rect && n_keep <= kRectRows && d.Npad >= kRectMinNpad) ? 2u : 1u;
    atomicAdd(n_pruned, 1);
    atomicAdd(&sticky->pruned_total, 1u);  // the host learns from it whether trying pays on this ctx's workload
  }
}

// Order-preserving compaction of the tile list (the runs the host dealt stay runs) and the tile list of the RECT
// instance.  One CTA.  out_total[0] = tiles left, or -1 if no pair was pruned (the tensor-core kernel then walks the
// original list); out_total[1] = pruned pairs (zeroed with the chunk state, counted by node_plan_kernel);
// out_total[2] = RECT tiles: for every pair with plan.pruned == 2, column block by column block, its row blocks of
// 256 kept nodes — entries (pair, row block << 16 | column block) like the square tiles.
__global__ void __launch_bounds__(1024) tile_compact_kernel(const PairDesc* __restrict__ descs, int pairs,
                                                            const uint2* __restrict__ tiles, int total,
                                                            const NodePlan* __restrict__ plan,
                                                            uint2* __restrict__ out, int* __restrict__ out_total,
                                                            uint2* __restrict__ rect) {
  __shared__ int s_w[32];
  __shared__ int s_base;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (out_total[1] == 0) {
    if (t == 0) out_total[0] = -1;
    return;
  }
  if (t == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < total; i0 += 1024) {
    const int i = i0 + t;
    uint2 e = make_uint2(0u, 0u);
    bool keep = false;
    if (i < total) {
      e = tiles[i];
      keep = plan[e.x].pruned == 0u;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    if (keep) out[before + __popc(m & ((1u << lane) - 1u))] = e;
    __syncthreads();
    if (t == 0) {
      int add = 0;
      for (int w = 0; w < 32; ++w) add += s_w[w];
      s_base += add;
    }
    __syncthreads();
  }
  if (t == 0) {
    out_total[0] = s_base;
    s_base = 0;
  }
  __syncthreads();
  for (int b0 = 0; b0 < pairs; b0 += 1024) {  // RECT tiles: one thread per pair, offsets by a block scan
    const int b = b0 + t;
    int nI = 0, nJ = 0;
    if (b < pairs && plan[b].pruned == 2u) {
      nI = (static_cast<int>(plan[b].n_keep) + kMmaTileM - 1) / kMmaTileM;
      nJ = (descs[b].N + kMmaTileN - 1) / kMmaTileN;
    }
    const int cnt = nI * nJ;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    uint2* dst = rect + before + incl - cnt;
    for (int jq = 0; jq < nJ; ++jq)
      for (int ib = 0; ib < nI; ++ib) *dst++ = make_uint2(static_cast<unsigned>(b), (static_cast<unsigned>(ib) << 16) | static_cast<unsigned>(jq));
    __syncthreads();
    if (t == 0) {
      int add = 0;
      for (int w = 0; w < 32; ++w) add += s_w[w];
      s_base += add;
    }
    __syncthreads();
  }
  if (t == 0) out_total[2] = s_base;
}

// Compact copy of the kept nodes' K-panel records for the RECT instance: [panel][kRectRows rows][8 words] per pair,
// row a = kept node a; the rows up to the next multiple of 256 are zero.  grid (panels, pairs).
__global__ void __launch_bounds__(256) kept_panel_kernel(const PairDesc* __restrict__ descs, const uint32_t* __restrict__ panel,
                                                         const NodePlan* __restrict__ plan, const unsigned short* __restrict__ kept,
                                                         uint32_t* __restrict__ kpanel, long long kpanel_pair_words) {
  const int pair = blockIdx.y;
  const NodePlan pl = plan[pair];
  if (pl.pruned != 2u) return;
  const PairDesc d = descs[pair];
  const int p = blockIdx.x;
  if (p >= ((d.npanel + 1) & ~1)) return;
  const int rows = (static_cast<int>(pl.n_keep) + kMmaTileM - 1) / kMmaTileM * kMmaTileM;
  const uint4* src = reinterpret_cast<const uint4*>(panel + d.panel_off) + static_cast<size_t>(p) * d.Npad * 2;
  uint4* dst = reinterpret_cast<uint4*>(kpanel + static_cast<long long>(pair) * kpanel_pair_words) + static_cast<size_t>(p) * kRectRows * 2;
  const unsigned short* kp = kept + static_cast<size_t>(pair) * kNodeKeepMax;
  for (int k = threadIdx.x; k < 2 * rows; k += 256) {
    const int a = k >> 1, h = k & 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (a < static_cast<int>(pl.n_keep) && p < d.npanel) v = __ldg(src + static_cast<size_t>(kp[a]) * 2 + h);
    dst[static_cast<size_t>(a) * 2 + h] = v;
  }
}

// popc(x & y) summed over WPL words per lane.  POPC runs on the XU pipe at a quarter of the ALU rate (16 lanes per
// clock and SM), so the words first go through a carry-save adder tree (two LOP3 per adder): 10 words need 4 POPC
// instead of 10, 5 words need 3.
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
  sum = a ^ b ^ c;
  carry = (a & b) | (c & (a ^ b));
}
template <int WPL>
__device__ __forceinline__ uint32_t and_popc(const uint32_t (&x)[WPL], const uint32_t (&y)[WPL]) {
  uint32_t a[WPL];
#pragma unroll
  for (int s = 0; s < WPL; ++s) a[s] = x[s] & y[s];
  if constexpr (WPL == 10) {
    uint32_t s0, c0, s1, c1, s2, c2, S, c3, t0, f0, T, f1;
    csa(a[0], a[1], a[2], s0, c0);
    csa(a[3], a[4], a[5], s1, c1);
    csa(a[6], a[7], a[8], s2, c2);
    csa(s0, s1, s2, S, c3);
    const uint32_t ones = S ^ a[9], c4 = S & a[9];
    csa(c0, c1, c2, t0, f0);   // weight 2 -> sum of weight 2, carry of weight 4
    csa(c3, c4, t0, T, f1);
    return __popc(ones) + 2 * __popc(T) + 4 * (__popc(f0) + __popc(f1));
  } else if constexpr (WPL == 5) {
    uint32_t s0, c0, ones, c1;
    csa(a[0], a[1], a[2], s0, c0);
    csa(s0, a[3], a[4], ones, c1);
    return __popc(ones) + 2 * (__popc(c0) + __popc(c1));
  } else {
    uint32_t c = 0;
#pragma unroll
    for (int s = 0; s < WPL; ++s) c += __popc(a[s]);
    return c;
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// grid (row slices, pairs, block lanes): a CTA takes one slice of the pair's rows as the streamed operand and walks the
// kept rows of the pair against it, 64 at a time (blocks blockIdx.z, blockIdx.z + gridDim.z, ...); node sums are added
// atomically.  Slices and block lanes follow from the pairs in the chunk (about four CTAs per SM): a single pair
// fills the device, and a chunk without any pruned pair costs one wave of CTAs that leave at once.
// 16 warps with 4 kept rows each in registers; the pair's rows stream through a ring of kKeptStages shared-memory
// buffers, 32 at a time (contiguous in memory: one bulk copy, SASS UBLKCP, issued by one lane two batches ahead),
// guarded by full / empty mbarriers — no CTA-wide barrier inside the loop.  For kept row k and a neighbour j of it the consumers form
// popc(row_k & row_j) per lane; the warp-wide sum T_kj is only needed where a key can arise (j kept too, j > k),
// everything else goes into per-lane partial node sums that are reduced once at the end.  Keys are staged per warp.
// Shared memory: ring [kKeptStages][32 * max_stride] | key buffers [16][kKeptWarpKeys] | histogram | barriers.
constexpr int kKeptStages = 3;
constexpr int kKeptWarpKeys = 64;
constexpr int kKeptConsumers = kKeptThreads / 32;       // 16

template <int WPL>
__global__ void __launch_bounds__(kKeptThreads, 1) triangles_kept_kernel(
    const PairDesc* __restrict__ descs, const uint32_t* __restrict__ adj, const NodePlan* __restrict__ plan,
    const unsigned short* __restrict__ kept, const uint32_t* __restrict__ keptbits, const ChunkDev* __restrict__ chunk,
    PairDev* __restrict__ state, unsigned long long* __restrict__ keys, uint32_t* __restrict__ hist,
    unsigned long long* __restrict__ t2, int max_stride, int bins) {
  if (chunk->overflow || !chunk->use_tensor) return;
  const int pair = blockIdx.y;
  const NodePlan pl = plan[pair];
  if (pl.pruned != 1u) return;  // 2: the tensor-core kernel's RECT instance takes the pair
  const PairDesc d = descs[pair];
  if (d.stride > 32 * WPL) return;  // launched with the instance that fits the chunk's longest row
  // batches of 32 rows: row block b <-> adjacency word b of a kept row; this CTA's slice of them
  const int b_begin = static_cast<int>((static_cast<long long>(d.stride) * blockIdx.x) / gridDim.x);
  const int b_end = static_cast<int>((static_cast<long long>(d.stride) * (blockIdx.x + 1)) / gridDim.x);
  if (b_begin >= b_end) return;
  const uint32_t* adjp = adj + d.adj_off;
  extern __shared__ __align__(128) unsigned char kp_smem[];
  uint32_t* ring = reinterpret_cast<uint32_t*>(kp_smem);
  const size_t slot_words = static_cast<size_t>(32) * max_stride;
  unsigned long long* kbuf = reinterpret_cast<unsigned long long*>(ring + kKeptStages * slot_words);
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(kbuf + kKeptConsumers * kKeptWarpKeys);
  uint64_t* full = reinterpret_cast<uint64_t*>(hist_s + ((bins + 1) & ~1));
  uint64_t* empty = full + kKeptStages;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int my_bins = min(bins, (d.N >> 4) + 1);
  for (int k = t; k < my_bins; k += kKeptThreads) hist_s[k] = 0u;
  if (t == 0) {
    for (int s = 0; s < kKeptStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kKeptConsumers);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t batch_bytes = static_cast<uint32_t>(32 * d.stride * 4);

  // The CTA walks the pair's kept rows 64 at a time and streams its slice of the rows once per such block (from
  // the L2 after the first time).  The ring's use counter runs on across the blocks (q0 = batches streamed so far).
  // lane 0 of warp 0 also issues the bulk copies, kKeptStages - 1 batches ahead of the batch the warps work on
  int q0 = 0;
  auto issue = [&](int bn) {
    const int qn = q0 + bn - b_begin, sn = qn % kKeptStages;
    if (qn >= kKeptStages) mbar_wait(&empty[sn], static_cast<uint32_t>((qn / kKeptStages - 1) & 1));
    mbar_arrive_expect_tx(&full[sn], batch_bytes);
    bulk_g2s(ring + sn * slot_words, adjp + static_cast<size_t>(bn) * 32 * d.stride, batch_bytes, &full[sn]);
  };
  for (int a0 = static_cast<int>(blockIdx.z) * kKeptRowsPerCta; a0 < static_cast<int>(pl.n_keep);
       a0 += static_cast<int>(gridDim.z) * kKeptRowsPerCta, q0 += b_end - b_begin) {
    if (t == 0)
      for (int bn = b_begin; bn < b_end && bn < b_begin + kKeptStages - 1; ++bn) issue(bn);
    int krow[kKeptRows];
    uint32_t rk[kKeptRows][WPL];
#pragma unroll
    for (int r = 0; r < kKeptRows; ++r) {
      const int a = a0 + warp * kKeptRows + r;
      const bool valid = a < static_cast<int>(pl.n_keep);
      krow[r] = valid ? static_cast<int>(kept[static_cast<size_t>(pair) * kNodeKeepMax + a]) : -1;
#pragma unroll
      for (int s = 0; s < WPL; ++s) {
        const int w = lane + 32 * s;
        rk[r][s] = (valid && w < d.stride) ? __ldg(adjp + static_cast<size_t>(krow[r]) * d.stride + w) : 0u;
      }
    }
    unsigned long long acc[kKeptRows];  // warp-uniform part of the node sums (the reduced counts)
    uint32_t accl[kKeptRows];           // per-lane part (counts that were never reduced); < 2^32: see the header
#pragma unroll
    for (int r = 0; r < kKeptRows; ++r) {
      acc[r] = 0ull;
      accl[r] = 0u;
    }
    const uint32_t thr = pl.min_deg - 1u;  // theta0
    const uint32_t* kbp = keptbits + d.mask_off;
    unsigned long long* kw = kbuf + warp * kKeptWarpKeys;
    unsigned long long* const keyp = keys + state[pair].key_base;
    int nk = 0;  // staged keys (warp-uniform)
    auto flush = [&]() {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&state[pair].key_count, static_cast<unsigned long long>(nk));
      base = __shfl_sync(0xffffffffu, base, 0);
      __syncwarp();
      for (int k = lane; k < nk; k += 32) keyp[base + k] = kw[k];
      __syncwarp();
      nk = 0;
    };
    // one (kept row, neighbour row) item: count, node sum, key
    auto item = [&](int r, uint32_t c, int j, bool jkept) {
      if (jkept && j > krow[r]) {  // a key can arise: the count itself is needed
        const uint32_t T = __reduce_add_sync(0xffffffffu, c);
        acc[r] += T;
        if (T >= thr) {
          if (lane == 0) {
            kw[nk] = (static_cast<unsigned long long>(T) << 32) |
                     (static_cast<unsigned long long>(0xFFFFu - static_cast<uint32_t>(krow[r])) << 16) |
                     static_cast<unsigned long long>(0xFFFFu - static_cast<uint32_t>(j));
            atomicAdd(&hist_s[T >> 4], 1u);
          }
          if (++nk == kKeptWarpKeys) flush();
        }
      } else {
        accl[r] += c;
      }
    };
    // adjacency words of the kept rows for 32 batches at a time: lane l holds word blk + l (one coalesced load per
    // row and 32 batches; the batch's word then comes from a shuffle)
    uint32_t mw[kKeptRows], kbw32 = 0u;
    int blk = -1;
    for (int b = b_begin; b < b_end; ++b) {
      const int q = q0 + b - b_begin, s = q % kKeptStages;
      if (warp == 0) {
        if (lane == 0 && b + kKeptStages - 1 < b_end) issue(b + kKeptStages - 1);
        __syncwarp();
      }
      if ((b & ~31) != blk) {
        blk = b & ~31;
        const int w = blk + lane;
#pragma unroll
        for (int r = 0; r < kKeptRows; ++r)
          mw[r] = (krow[r] >= 0 && w < d.stride) ? __ldg(adjp + static_cast<size_t>(krow[r]) * d.stride + w) : 0u;
        kbw32 = w < d.stride ? __ldg(kbp + w) : 0u;
      }
      uint32_t m[kKeptRows];
#pragma unroll
      for (int r = 0; r < kKeptRows; ++r) m[r] = __shfl_sync(0xffffffffu, mw[r], b & 31);
      const uint32_t kbw = __shfl_sync(0xffffffffu, kbw32, b & 31);
      uint32_t common = m[0];
#pragma unroll
      for (int r = 1; r < kKeptRows; ++r) common &= m[r];
      mbar_wait(&full[s], static_cast<uint32_t>((q / kKeptStages) & 1));
      const uint32_t* rows = ring + s * slot_words + lane;
      // neighbours of all four rows (the clique's rows against each other): the streamed row is read once and the
      // four counts are independent instruction streams
      for (uint32_t bits = common; bits;) {
        const int bb = __ffs(bits) - 1;
        bits &= bits - 1u;
        const uint32_t* y = rows + bb * d.stride;
        uint32_t yv[WPL];
#pragma unroll
        for (int w = 0; w < WPL; ++w) yv[w] = y[32 * w];
        uint32_t c[kKeptRows];
#pragma unroll
        for (int r = 0; r < kKeptRows; ++r) c[r] = and_popc<WPL>(rk[r], yv);
        const int j = 32 * b + bb;
        const bool jkept = (kbw >> bb) & 1u;
#pragma unroll
        for (int r = 0; r < kKeptRows; ++r) item(r, c[r], j, jkept);
      }
      // the others, row by row (one neighbour of every row per step, to have four chains in flight, was measured
      // slower: 5.5 ms against 4.0 ms per 128-pair KITTI-scale step — the rows' neighbour counts differ too much)
#pragma unroll
      for (int r = 0; r < kKeptRows; ++r) {
        for (uint32_t bits = m[r] & ~common; bits;) {
          const int bb = __ffs(bits) - 1;
          bits &= bits - 1u;
          const uint32_t* y = rows + bb * d.stride;
          uint32_t yv[WPL];
#pragma unroll
          for (int w = 0; w < WPL; ++w) yv[w] = y[32 * w];
          item(r, and_popc<WPL>(rk[r], yv), 32 * b + bb, (kbw >> bb) & 1u);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (nk) flush();
#pragma unroll
    for (int r = 0; r < kKeptRows; ++r) {
      const unsigned long long tot = acc[r] + __reduce_add_sync(0xffffffffu, accl[r]);
      if (lane == 0 && krow[r] >= 0 && tot) atomicAdd(&t2[d.node_off + krow[r]], tot);  // t2 starts at zero
    }
  }
  __syncthreads();
  uint32_t* histp = hist + static_cast<size_t>(pair) * kHistBins;
  for (int k = t; k < my_bins; k += kKeptThreads) {
    const uint32_t v = hist_s[k];
    if (v) atomicAdd(&histp[k], v);
  }
}

size_t kept_smem_bytes(int max_stride, int bins) {
  return static_cast<size_t>(kKeptStages) * 32 * max_stride * 4 + static_cast<size_t>(kKeptConsumers) * kKeptWarpKeys * 8 +
         static_cast<size_t>((bins + 1) & ~1) * 4 + 2 * kKeptStages * 8;
}

}  // namespace

int node_prune_configure() {
  const int bins = (kNodePruneMaxNpad >> 4) + 1;
  cudaError_t e = cudaFuncSetAttribute(triangles_kept_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kept_smem_bytes(160, bins)));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(triangles_kept_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(kept_smem_bytes(320, bins)));
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

int launch_node_plan(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_npad, const uint32_t* d_adj,
                     const ChunkDev* d_chunk, StickyDev* d_sticky, const PairDev* d_state, const uint32_t* d_theta, unsigned short* d_deg, NodePlan* d_plan,
                     unsigned short* d_kept, uint32_t* d_keptbits, const uint2* d_tiles, int total_tiles, uint2* d_tiles_out,
                     int* d_total, int cost, int force, uint2* d_rect_tiles, const uint32_t* d_panel, uint32_t* d_kpanel,
                     long long kpanel_pair_words, int max_npanel) {
  int gx = (8 * lc.sm_count + pairs - 1) / pairs;
  gx = std::max(1, std::min(gx, (max_npad + 31) / 32));
  node_degree_kernel<<<dim3(gx, pairs), 256, 0, lc.stream>>>(d_desc, d_adj, d_chunk, d_state, d_theta, d_deg, force);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  node_plan_kernel<<<pairs, 1024, 0, lc.stream>>>(d_desc, d_chunk, d_sticky, d_state, d_theta, d_deg, d_plan, d_kept, d_keptbits, d_total + 1, cost, force, d_kpanel != nullptr ? 1 : 0);
  e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  tile_compact_kernel<<<1, 1024, 0, lc.stream>>>(d_desc, pairs, d_tiles, total_tiles, d_plan, d_tiles_out, d_total, d_rect_tiles);
  e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  if (d_kpanel == nullptr) return 3;
  kept_panel_kernel<<<dim3((max_npanel + 1) & ~1, pairs), 256, 0, lc.stream>>>(d_desc, d_panel, d_plan, d_kept, d_kpanel, kpanel_pair_words);
  e = cudaGetLastError();
  return e == cudaSuccess ? 4 : -static_cast<int>(e);
}

int launch_triangles_kept(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_n, int max_stride,
                          const uint32_t* d_adj, const NodePlan* d_plan, const unsigned short* d_kept,
                          const uint32_t* d_keptbits, const ChunkDev* d_chunk, PairDev* d_state, unsigned long long* d_keys, uint32_t* d_hist,
                          unsigned long long* d_t2) {
  if (max_stride > 320) return 0;  // no pair of the chunk can be pruned (node_plan_kernel left them alone)
  const int bins = (max_n >> 4) + 1;
  const size_t smem = kept_smem_bytes(max_stride, bins);
  const int want = (4 * lc.sm_count + pairs - 1) / pairs;            // CTAs per pair
  const int slices = std::max(1, std::min(want, max_stride / 8));    // at least eight batches of 32 rows per slice
  const int lanes = std::max(1, std::min((want + slices - 1) / slices, kNodeKeepMax / kKeptRowsPerCta));
  const dim3 grid(slices, pairs, lanes);
  if (max_stride <= 160)
    triangles_kept_kernel<5><<<grid, kKeptThreads, smem, lc.stream>>>(d_desc, d_adj, d_plan, d_kept, d_keptbits, d_chunk,
                                                                         d_state, d_keys, d_hist, d_t2, max_stride, bins);
  else
    triangles_kept_kernel<10><<<grid, kKeptThreads, smem, lc.stream>>>(d_desc, d_adj, d_plan, d_kept, d_keptbits, d_chunk,
                                                                          d_state, d_keys, d_hist, d_t2, max_stride, bins);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

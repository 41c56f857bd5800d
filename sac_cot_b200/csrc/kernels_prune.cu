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
//   triangles_kept_kernel the loop above: a warp holds R kept rows in registers, the CTA streams all rows of the pair
//                        through shared memory in batches of 32 (cp.async, double buffered)
#include "common.cuh"

#include <algorithm>

namespace saccot {

namespace {

constexpr int kKeptThreads = 512;   // 16 warps
constexpr int kKeptRows = 4;        // kept rows per warp
constexpr int kKeptRowsPerCta = (kKeptThreads / 32) * kKeptRows;  // 64
constexpr int kKeptKeyCap = 2 * kKeptRowsPerCta * 32;             // staged keys: two batches' worst case

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
  for (int r = blockIdx.x * 8 + warp; r < d.N; r += gridDim.x * 8) {
    const uint4* rp = base + static_cast<size_t>(r) * q4;
    int c = 0;
    for (int q = lane; q < q4; q += 32) {
      const uint4 w = __ldg(rp + q);
      c += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) out[r] = static_cast<unsigned short>(c);  // deg <= N - 1 <= 65534
  }
}

// One CTA per pair.  Decision: the kept-row kernel costs ~ sum of the kept degrees x row length, the tensor-core
// kernel ~ Npad^3, so the pair is pruned if (sum of kept degrees) x cost <= Npad^2 (cost: api.cu).
__global__ void __launch_bounds__(1024) node_plan_kernel(const PairDesc* __restrict__ descs,
                                                         const ChunkDev* __restrict__ chunk,
                                                         const PairDev* __restrict__ state,
                                                         const uint32_t* __restrict__ theta,
                                                         const unsigned short* __restrict__ deg,
                                                         NodePlan* __restrict__ plan, unsigned short* __restrict__ kept,
                                                         uint32_t* __restrict__ keptbits, int* __restrict__ n_pruned,
                                                         int cost, int force) {
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
    pl->pruned = 1u;
    atomicAdd(n_pruned, 1);
  }
}

// Order-preserving compaction of the tile list (the runs the host dealt stay runs).  One CTA.  out_total[0] = tiles
// left, or -1 if no pair was pruned (the tensor-core kernel then walks the original list); out_total[1] = pruned
// pairs (zeroed with the chunk state, counted by node_plan_kernel).
__global__ void __launch_bounds__(1024) tile_compact_kernel(const uint2* __restrict__ tiles, int total,
                                                            const NodePlan* __restrict__ plan,
                                                            uint2* __restrict__ out, int* __restrict__ out_total) {
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
  if (t == 0) *out_total = s_base;
}

// popc(a & b) summed over WPL words per lane with carry-save adders (POPC shares the slow XU pipe with REDUX)
template <int WPL>
__device__ __forceinline__ int and_popc(const uint32_t (&x)[WPL], const uint32_t* __restrict__ y) {
  uint32_t a[WPL];
#pragma unroll
  for (int s = 0; s < WPL; ++s) a[s] = x[s] & y[32 * s];
  int c = 0;
  int s = 0;
#pragma unroll
  for (; s + 2 < WPL; s += 3) {
    const uint32_t lo = a[s] ^ a[s + 1] ^ a[s + 2];
    const uint32_t hi = (a[s] & a[s + 1]) | (a[s + 2] & (a[s] ^ a[s + 1]));
    c += __popc(lo) + 2 * __popc(hi);
  }
#pragma unroll
  for (; s < WPL; ++s) c += __popc(a[s]);
  return c;
}

// grid (row blocks, pairs, row slices): with few pairs in the chunk the streamed rows are split over gridDim.z CTAs
// (a single pair would otherwise keep a handful of SMs busy) and the node sums are added atomically.
// Shared memory: row buffers [2][32][stride] | staged keys [kKeptKeyCap] | histogram.
template <int WPL>
__global__ void __launch_bounds__(kKeptThreads) triangles_kept_kernel(
    const PairDesc* __restrict__ descs, const uint32_t* __restrict__ adj, const NodePlan* __restrict__ plan,
    const unsigned short* __restrict__ kept, const ChunkDev* __restrict__ chunk, PairDev* __restrict__ state,
    unsigned long long* __restrict__ keys, uint32_t* __restrict__ hist, unsigned long long* __restrict__ t2,
    int max_stride, int bins) {
  if (chunk->overflow || !chunk->use_tensor) return;
  const int pair = blockIdx.y;
  const NodePlan pl = plan[pair];
  const int a0 = blockIdx.x * kKeptRowsPerCta;
  if (!pl.pruned || a0 >= static_cast<int>(pl.n_keep)) return;
  const PairDesc d = descs[pair];
  if (d.stride > 32 * WPL) return;  // launched with the instance that fits the chunk's longest row
  const uint32_t* adjp = adj + d.adj_off;
  extern __shared__ __align__(16) unsigned char kp_smem[];
  uint32_t* buf = reinterpret_cast<uint32_t*>(kp_smem);  // [2][32 * stride]
  unsigned long long* kst = reinterpret_cast<unsigned long long*>(buf + static_cast<size_t>(2) * 32 * max_stride);
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(kst + kKeptKeyCap);
  __shared__ uint32_t s_nkeys;
  __shared__ unsigned long long s_kbase;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int my_bins = min(bins, (d.N >> 4) + 1);
  for (int k = t; k < my_bins; k += kKeptThreads) hist_s[k] = 0u;
  if (t == 0) s_nkeys = 0u;

  // this warp's kept rows, resident in registers
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
  unsigned long long acc[kKeptRows];
#pragma unroll
  for (int r = 0; r < kKeptRows; ++r) acc[r] = 0ull;
  const uint32_t thr = pl.min_deg - 1u;  // theta0
  // batches of 32 rows: row block b <-> adjacency word b of a kept row; this CTA's slice of them
  const int nb_all = d.stride;
  const int b_begin = static_cast<int>((static_cast<long long>(nb_all) * blockIdx.z) / gridDim.z);
  const int nb = static_cast<int>((static_cast<long long>(nb_all) * (blockIdx.z + 1)) / gridDim.z);
  if (b_begin >= nb) return;
  const int q4 = (32 * d.stride) >> 2;   // uint4 per batch (the 32 rows are contiguous in memory)
  auto issue = [&](int b) {
    const uint4* src = reinterpret_cast<const uint4*>(adjp + static_cast<size_t>(b) * 32 * d.stride);
    const uint32_t dst = smem_u32(buf + static_cast<size_t>((b - b_begin) & 1) * 32 * max_stride);
    for (int q = t; q < q4; q += kKeptThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * q), "l"(src + q) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto flush = [&]() {  // all threads; s_nkeys is stable (read after a barrier)
    const uint32_t n = s_nkeys;
    if (n == 0u) return;
    if (t == 0) s_kbase = state[pair].key_base + atomicAdd(&state[pair].key_count, static_cast<unsigned long long>(n));
    __syncthreads();
    unsigned long long* dst = keys + s_kbase;
    for (uint32_t k = t; k < n; k += kKeptThreads) dst[k] = kst[k];
    __syncthreads();
    if (t == 0) s_nkeys = 0u;
    __syncthreads();
  };
  uint32_t mnext[kKeptRows];
#pragma unroll
  for (int r = 0; r < kKeptRows; ++r) mnext[r] = krow[r] >= 0 ? __ldg(adjp + static_cast<size_t>(krow[r]) * d.stride + b_begin) : 0u;
  issue(b_begin);
  for (int b = b_begin; b < nb; ++b) {
    uint32_t m[kKeptRows];
#pragma unroll
    for (int r = 0; r < kKeptRows; ++r) {
      m[r] = mnext[r];
      mnext[r] = (krow[r] >= 0 && b + 1 < nb) ? __ldg(adjp + static_cast<size_t>(krow[r]) * d.stride + b + 1) : 0u;
    }
    if (b + 1 < nb) {
      issue(b + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t* rows = buf + static_cast<size_t>((b - b_begin) & 1) * 32 * max_stride + lane;
    uint32_t any = 0u;
#pragma unroll
    for (int r = 0; r < kKeptRows; ++r) any |= m[r];
    while (any) {
      const int bb = __ffs(any) - 1;
      any &= any - 1u;
      const uint32_t* y = rows + bb * d.stride;
      const int j = 32 * b + bb;
#pragma unroll
      for (int r = 0; r < kKeptRows; ++r) {
        if ((m[r] >> bb) & 1u) {  // warp-uniform
          const uint32_t T = static_cast<uint32_t>(__reduce_add_sync(0xffffffffu, and_popc<WPL>(rk[r], y)));
          acc[r] += T;
          if (lane == 0 && j > krow[r] && T >= thr) {
            const uint32_t pos = atomicAdd(&s_nkeys, 1u);
            kst[pos] = (static_cast<unsigned long long>(T) << 32) |
                       (static_cast<unsigned long long>(0xFFFFu - static_cast<uint32_t>(krow[r])) << 16) |
                       static_cast<unsigned long long>(0xFFFFu - static_cast<uint32_t>(j));
            atomicAdd(&hist_s[T >> 4], 1u);
          }
        }
      }
    }
    __syncthreads();  // the batch buffer may be overwritten; every key of this batch is staged
    if (s_nkeys > static_cast<uint32_t>(kKeptKeyCap / 2)) flush();
  }
  flush();
#pragma unroll
  for (int r = 0; r < kKeptRows; ++r)
    if (lane == 0 && krow[r] >= 0 && acc[r]) atomicAdd(&t2[d.node_off + krow[r]], acc[r]);  // t2 starts at zero
  uint32_t* histp = hist + static_cast<size_t>(pair) * kHistBins;
  for (int k = t; k < my_bins; k += kKeptThreads) {
    const uint32_t v = hist_s[k];
    if (v) atomicAdd(&histp[k], v);
  }
}

size_t kept_smem_bytes(int max_stride, int bins) {
  return static_cast<size_t>(2) * 32 * max_stride * 4 + static_cast<size_t>(kKeptKeyCap) * 8 + static_cast<size_t>(bins) * 4;
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
                     const ChunkDev* d_chunk, const PairDev* d_state, const uint32_t* d_theta, unsigned short* d_deg, NodePlan* d_plan,
                     unsigned short* d_kept, uint32_t* d_keptbits, const uint2* d_tiles, int total_tiles, uint2* d_tiles_out,
                     int* d_total, int cost, int force) {
  int gx = (8 * lc.sm_count + pairs - 1) / pairs;
  gx = std::max(1, std::min(gx, (max_npad + 7) / 8));
  node_degree_kernel<<<dim3(gx, pairs), 256, 0, lc.stream>>>(d_desc, d_adj, d_chunk, d_state, d_theta, d_deg, force);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  node_plan_kernel<<<pairs, 1024, 0, lc.stream>>>(d_desc, d_chunk, d_state, d_theta, d_deg, d_plan, d_kept, d_keptbits, d_total + 1, cost, force);
  e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  tile_compact_kernel<<<1, 1024, 0, lc.stream>>>(d_tiles, total_tiles, d_plan, d_tiles_out, d_total);
  e = cudaGetLastError();
  return e == cudaSuccess ? 3 : -static_cast<int>(e);
}

int launch_triangles_kept(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_n, int max_stride,
                          const uint32_t* d_adj, const NodePlan* d_plan, const unsigned short* d_kept,
                          const ChunkDev* d_chunk, PairDev* d_state, unsigned long long* d_keys, uint32_t* d_hist,
                          unsigned long long* d_t2) {
  if (max_stride > 320) return 0;  // no pair of the chunk can be pruned (node_plan_kernel left them alone)
  const int bins = (max_n >> 4) + 1;
  const size_t smem = kept_smem_bytes(max_stride, bins);
  // a pruned pair keeps a few hundred nodes (~6 row blocks): aim at two CTAs per SM
  int slices = (2 * lc.sm_count + 6 * pairs - 1) / (6 * pairs);
  slices = std::max(1, std::min(slices, std::min(32, max_stride / 4)));
  const dim3 grid((kNodeKeepMax + kKeptRowsPerCta - 1) / kKeptRowsPerCta, pairs, slices);
  if (max_stride <= 160)
    triangles_kept_kernel<5><<<grid, kKeptThreads, smem, lc.stream>>>(d_desc, d_adj, d_plan, d_kept, d_chunk, d_state, d_keys,
                                                                      d_hist, d_t2, max_stride, bins);
  else
    triangles_kept_kernel<10><<<grid, kKeptThreads, smem, lc.stream>>>(d_desc, d_adj, d_plan, d_kept, d_chunk, d_state, d_keys,
                                                                       d_hist, d_t2, max_stride, bins);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

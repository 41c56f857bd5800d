// common.cuh — shared declarations of the sm_100a SAC-COT kernels and their host launchers.
//
// Stage ids (S1..S7) refer to SURVEY.md §8a; the reference repository holds no code to cite
// (/root/reference/README.md:1-2), so the normative text is BASELINE.json `north_star`.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace saccot {

// ------------------------------------------------------------------------------------------
// Per-pair descriptors.  PairDesc is built on the host and read-only on the device; PairDev
// is device state zeroed at the start of every run.
// ------------------------------------------------------------------------------------------
struct PairDesc {
  int32_t N;         // correspondences
  int32_t Npad;      // N rounded up to a multiple of 128
  int32_t stride;    // 32-bit words per adjacency row = Npad / 32 (multiple of 4 => 16 B rows)
  int32_t nblk;      // Npad / 128
  int64_t pt_off;    // first point of this pair in the packed AoS inputs
  int64_t soa_off;   // first float of this pair's 6 x Npad SoA block (sx,sy,sz,dx,dy,dz)
  int64_t adj_off;   // first u32 word of this pair's Npad x stride adjacency block
  int64_t node_off;  // first entry of this pair in per-node arrays (t2), sum of Npad
  int64_t mask_off;  // first u32 word of this pair's inlier mask (Npad/32 words)
  int64_t panel_off; // first u32 word of this pair's K-panel copy of the adjacency (tensor-core path)
  int32_t npanel;    // 256-column panels = ceil(Npad / 256); panel p holds [Npad rows][8 words]
  int32_t tile_base; // first tile index of this pair in the chunk's tile list (tensor-core path)
};

struct PairDev {
  unsigned long long num_edges;  // evaluated edges (all of E unless sharded), from the unit scan
  unsigned long long key_base;   // offset of this pair's edge keys in the key pool (scan)
  unsigned long long key_count;  // keys in this pair's slice of the pool (== num_edges)
  unsigned long long best_key;   // max hypothesis key
  uint32_t sel_count;            // keys appended to the selected list so far
  uint32_t tie_count;            // keys appended to the tie list so far
  uint32_t dstar;                // threshold digit (T >> 4) of the top-K_e selection
  uint32_t n_above;              // #keys with digit > dstar
  uint32_t n_tie;                // #keys with digit == dstar
  uint32_t tie_inplace;          // 1: tie bucket too large for the tie list, scan keys in place
  uint32_t n_sel;                // number of selected edges = min(K_e, evaluated edges)
  uint32_t pad_;
  unsigned long long all_edges;  // edges of the whole pair (== num_edges unless sharded): decides the S2 path
};

struct ChunkDev {
  unsigned long long total_edges;  // sum of E over the chunk (key pool demand)
  uint32_t overflow;               // 1: key pool too small; downstream kernels do nothing
  uint32_t use_tensor;             // 1: S2 runs on the tensor cores for this chunk, 0: POPC bitset kernels
};

// Survives across calls (never zeroed by the pipeline): lets device-location calls report a
// key-pool overflow without a host synchronisation.
struct StickyDev {
  unsigned long long max_total_edges;  // largest chunk demand seen in an overflowing chunk
  uint32_t overflow_count;             // number of chunks that overflowed so far
  uint32_t pruned_total;               // pairs that took the kept-row kernel so far (kernels_prune.cu)
};

constexpr int kMaxEdges = 4096;     // == SAC_COT_MAX_EDGES
constexpr int kMaxApex = 8;         // == SAC_COT_MAX_APEX
constexpr int kHistBins = 4096;     // histogram of T >> 4
constexpr int kTieCap = 16384;      // tie-list entries per pair
constexpr int kTriJ = 128;          // columns per triangle work unit (unit = 128 columns x 256 rows)
constexpr int kTriI = 256;          // rows of i per triangle work unit
constexpr int kTriThreads = 512;
constexpr int kTriMaxR = 11;        // max words per lane per row chunk (352 words)
constexpr int kTriChunkR = 8;       // words per lane per chunk when rows do not fit (256 words)

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk asynchronous copy (TMA, 1-D form; SASS UBLKCP).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other > v ? other : v;
  }
  return v;
}

// ------------------------------------------------------------------------------------------
// Host launchers (one per kernel file).  Every launcher returns the number of kernels it
// enqueued on `stream` (for the ctx "launches" counter) or a negative cudaError_t.
// ------------------------------------------------------------------------------------------
// Triangle work units: unit (jb, ic) = columns [128 jb, 128 jb + 128) x rows [256 ic, 256 ic + 256)
// with ic <= jb/2 (only i < j matters).  unit id = unit_offset(jb) + ic; the oracle uses the same
// numbering; in sharded mode a unit belongs to the rank owner_of_unit() names.
__host__ __device__ inline unsigned int unit_offset(unsigned int jb) {
  const unsigned int h = jb >> 1;  // sum_{b<jb} (b/2+1) = h(h+1) for jb=2h, (h+1)^2 for jb=2h+1
  return (jb & 1u) ? (h + 1) * (h + 1) : h * (h + 1);
}
__host__ __device__ inline unsigned int unit_count(unsigned int nblk) { return unit_offset(nblk); }
// Sharded single-pair runs (normative, DESIGN.md §2 "S2 partition"): the owner of an edge (i < j) is decided by the
// cell it lies in — cell (cb, ic) = columns [1920 cb, 1920 cb + 1920) x rows [256 ic, 256 ic + 256) — and cells are
// dealt to the ranks as (257 cb + ic) mod world.  1920 = lcm(128, 240) columns are whole 128-column units of the
// bitset kernels and whole 240-column tiles of the tensor-core kernel; 256 rows are one row block of either.  A
// column block holds ~7.5 (cb + 1/2) cells, so every rank gets the same share of every column block to within one
// cell (the triangular work profile over j does not unbalance the ranks), and 257 = 1 (mod 2, 4, 8) shifts the deal
// from one column block to the next.  The oracle restates the same rule.
constexpr unsigned int kOwnerCols = 1920;
constexpr unsigned int kOwnerRows = 256;
__host__ __device__ inline unsigned int owner_of_cell(unsigned int cb, unsigned int ic, unsigned int world) {
  return (257u * cb + ic) % world;
}
// column block jb of unit id u (inverse of unit_offset)
__host__ __device__ inline unsigned int unit_jb(unsigned int u) {
  unsigned int jb = static_cast<unsigned int>(2.0f * sqrtf(static_cast<float>(u)));
  while (unit_offset(jb) > u) --jb;
  while (unit_offset(jb + 1) <= u) ++jb;
  return jb;
}
__host__ __device__ inline unsigned int owner_of_unit(unsigned int u, unsigned int world) {
  const unsigned int jb = unit_jb(u);
  return owner_of_cell(jb / (kOwnerCols / 128u), u - unit_offset(jb), world);
}

struct LaunchCtx {
  cudaStream_t stream;
  int sm_count;
};

// kernels_graph.cu — S0 repack + S1 compatibility graph
int launch_pack_soa(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_npad, const float* d_src,
                    const float* d_dst, float* d_soa);
int launch_graph(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_nblk, const float* d_soa,
                 uint32_t* d_adj, uint32_t* d_panel, uint32_t* d_ucount, int unit_pitch, float tau);
int launch_unit_scan(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, PairDev* d_state,
                     const uint32_t* d_ucount, uint32_t* d_ubase, int unit_pitch, int rank, int world);
// tri_mode: 0 = POPC kernels, 1 = tensor-core kernel, 2 = decided here from the chunk's edge density
// d_prev (or null): chunk header of an earlier pass over the same chunk whose overflow carries over
int launch_key_scan(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, PairDev* d_state, ChunkDev* d_chunk,
                    StickyDev* d_sticky, unsigned long long key_cap, int tri_mode, const ChunkDev* d_prev);
// second-order compatibility: key list of the first pass -> A2 (+ its K-panel copy) and A2's per-unit edge counts
int launch_second_order_scatter(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, const PairDev* d_state1,
                                const ChunkDev* d_chunk1, const unsigned long long* d_keys, uint32_t cmin,
                                uint32_t* d_adj2, uint32_t* d_panel2, uint32_t* d_ucount, int unit_pitch);
int launch_fill_u32(const LaunchCtx& lc, uint32_t* d_p, uint32_t v, int n);
// density (edges per node pair) from which the tensor-core triangle kernel beats the POPC kernels, and the
// smallest pair it is worth starting for (measured on B200)
constexpr float kTensorMinDensity = 0.02f;
constexpr int kTensorMinN = 1024;

// kernels_triangles.cu — S2 triangle counts (POPC bitset path)
int launch_triangles(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_nblk, int max_stride,
                     const uint32_t* d_adj, PairDev* d_state, const ChunkDev* d_chunk, unsigned long long* d_keys,
                     const uint32_t* d_ubase, int unit_pitch, uint32_t* d_hist, unsigned long long* d_t2, int rank,
                     int world);
int triangles_configure();  // opt-in dynamic shared memory; call once per device

struct NodePlan;
// kernels_triangles_mma.cu — S2 triangle counts on the tensor cores (tcgen05 kind::mxf4, TMEM)
constexpr int kMmaTileM = 256;      // rows of i per tile: a CTA pair (cta_group::2), 128 rows = TMEM lanes per CTA
constexpr int kMmaTileN = 240;      // columns of j per tile (UMMA N = 240; two accumulators = 480 TMEM columns)
// tiles of one pair: J-blocks jq = 0 .. ceil(N/240)-1, each with the row blocks that hold some i < j
__host__ __device__ inline int mma_tiles_of_jblock(int N, int jq) {
  const int nI = (N + kMmaTileM - 1) / kMmaTileM;
  const int c = (kMmaTileN * jq + kMmaTileN - 2) / kMmaTileM + 1;
  return c < nI ? c : nI;
}
__host__ __device__ inline int mma_tiles_of_pair(int N) {
  const int nJ = (N + kMmaTileN - 1) / kMmaTileN;
  int t = 0;
  for (int jq = 0; jq < nJ; ++jq) t += mma_tiles_of_jblock(N, jq);
  return t;
}
// CTA pairs (clusters) the tensor-core kernel runs with; cluster c walks entries c, c + n, c + 2 n, ... of the tile list
inline int mma_clusters(int total_tiles, int sm_count) { return total_tiles < sm_count / 2 ? total_tiles : sm_count / 2; }
// d_scratch (theta_scratch_bytes(pairs) bytes, or null): chunks of fewer pairs than the device has SMs spread the
// sample evaluation over many CTAs (three launches instead of one)
constexpr int kThetaSplitMaxPairs = 147;
size_t theta_scratch_bytes(int pairs);
int launch_tri_theta(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_npad, const uint32_t* d_adj,
                     const ChunkDev* d_chunk, uint32_t* d_theta, void* d_scratch, int Ke, int prune);
// d_total / d_tiles2 (or null): if *d_total >= 0 the kernel walks the *d_total entries of d_tiles2 instead (the list
// without the tiles of node-pruned pairs, compacted on the device)
int launch_triangles_mma(const LaunchCtx& lc, const PairDesc* d_desc, const uint2* d_tiles, int total_tiles,
                         const int* d_total, const uint2* d_tiles2, const uint32_t* d_adj, const uint32_t* d_panel, PairDev* d_state,
                         const ChunkDev* d_chunk, unsigned long long* d_keys, uint32_t* d_theta, uint32_t* d_hist,
                         unsigned long long* d_t2, int Ke, int raise, int dbg);
// the RECT instance: tiles of node-pruned pairs (256 kept nodes x 240 columns; list and length on the device)
constexpr int kRectRows = 1024;     // kept nodes a pair may have to take this path (4 row blocks); rows of its compact panel copy
constexpr int kRectMinNpad = 2048;  // shorter rows stay with the kept-row POPC kernel
int launch_triangles_mma_rect(const LaunchCtx& lc, const PairDesc* d_desc, const uint2* d_rect_tiles, int max_tiles,
                              const int* d_rect_total, const uint32_t* d_adj, const uint32_t* d_panel, PairDev* d_state,
                              const ChunkDev* d_chunk, unsigned long long* d_keys, uint32_t* d_theta, uint32_t* d_hist,
                              unsigned long long* d_t2, int Ke, int raise, int dbg, const struct NodePlan* d_plan,
                              const unsigned short* d_kept, const uint32_t* d_kpanel, long long kpanel_pair_words);
int triangles_mma_configure();
// tensor-pipe peak probe (the triangle kernel's MMA shape, issued back to back): bench.py's roofline denominator
int launch_mma_peak_probe(const LaunchCtx& lc, int clusters, int stage_pairs);
double mma_peak_probe_flops(int clusters, int stage_pairs);

// kernels_prune.cu — exact node pruning of S2 on the tensor-core path: pairs whose selectable edges join few
// high-degree nodes count triangles for those nodes' rows only
struct NodePlan {
  uint32_t pruned;   // the dense kernel skips the pair's tiles; its kept rows take 1: the POPC kept-row kernel, 2: the RECT instance
  uint32_t n_keep;   // nodes of degree >= min_deg, listed ascending in the pair's slice of the kept list
  uint32_t ub_rest;  // D (D - 1) / 2 >= t_k of every node outside the kept set, D = their largest degree (diagnostic)
  uint32_t min_deg;  // theta0 + 1
};
constexpr int kNodeKeepMax = 2048;         // kept-list slots per pair
constexpr int kNodePruneMaxNpad = 10240;   // longest row the kept-row kernel holds in registers (10 words per lane)
int node_prune_configure();
// exact degrees -> per-pair plan + kept list -> tile list without the pruned pairs' tiles (three launches).
// cost: a pair is pruned if (sum of kept degrees) x cost <= Npad^2; force >= 2: whenever the kept list fits (tests)
// d_total: int[4] in the chunk's zero region: [0] tiles left (-1: nothing pruned, use the original list), [1] pruned
// pairs, [2] tiles of the RECT list d_rect_tiles (pairs with plan.pruned == 2; d_kpanel = null: none gets that mode)
int launch_node_plan(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_npad, const uint32_t* d_adj,
                     const ChunkDev* d_chunk, StickyDev* d_sticky, const PairDev* d_state, const uint32_t* d_theta, unsigned short* d_deg, NodePlan* d_plan,
                     unsigned short* d_kept, uint32_t* d_keptbits, const uint2* d_tiles, int total_tiles, uint2* d_tiles_out,
                     int* d_total, int cost, int force, uint2* d_rect_tiles, const uint32_t* d_panel, uint32_t* d_kpanel,
                     long long kpanel_pair_words, int max_npanel);
int launch_triangles_kept(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_n, int max_stride,
                          const uint32_t* d_adj, const NodePlan* d_plan, const unsigned short* d_kept,
                          const uint32_t* d_keptbits, const ChunkDev* d_chunk, PairDev* d_state, unsigned long long* d_keys, uint32_t* d_hist,
                          unsigned long long* d_t2);

// kernels_select.cu — S3 edge ranking + apex selection
int launch_select_edges(const LaunchCtx& lc, int pairs, PairDev* d_state, const ChunkDev* d_chunk,
                        const unsigned long long* d_keys, const uint32_t* d_hist, unsigned long long* d_sel,
                        unsigned long long* d_tie, unsigned long long* d_top, int Ke);
int select_configure();  // opt-in dynamic shared memory; call once per device
// d_plan / d_deg / d_keptbits (or null): pairs whose node sums cover the kept nodes only, the exact degrees and the
// kept-node bit masks (kernels_prune.cu)
int launch_select_apex(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_npad, const uint32_t* d_adj,
                       const unsigned long long* d_t2, const unsigned long long* d_top, int32_t* d_tri, int Ke, int m,
                       int apex_path, const NodePlan* d_plan, const unsigned short* d_deg, const uint32_t* d_keptbits);

// sharded single pair: record of a rank for exchange #1 and the merge of the gathered records (kernels_select.cu)
int launch_shard_pack(const LaunchCtx& lc, const unsigned long long* d_t2, const unsigned long long* d_top,
                      const ChunkDev* d_chunk, unsigned long long* d_rec, int npad, int Ke);
int launch_shard_merge(const LaunchCtx& lc, const unsigned long long* d_recs, int world, int npad, int Ke,
                       unsigned long long* d_t2, unsigned long long* d_top, PairDev* d_state, ChunkDev* d_chunk,
                       StickyDev* d_sticky, unsigned long long* d_summary);

// kernels_match.cu — correspondence front end (descriptor nearest-neighbour matching)
struct MatchPair {
  int32_t Ns, Nd;          // source / target keypoints of the pair
  int32_t s_tiles, d_tiles;  // ceil(Ns / 128), ceil(Nd / 256)
  int64_t s_off, d_off;    // first row of the pair in the packed source / target arrays
  int64_t s_img, d_img;    // byte offsets of the pair's operand images
  int64_t s_norm;          // first entry of the pair's source norms (rows padded to whole tiles)
  int64_t cand_off;        // first row of the pair in the candidate lists (== s_off)
};
constexpr int kMatchTileM = 128;    // source rows per CTA (TMEM lanes)
constexpr int kMatchTileN = 256;    // target rows per tile (TMEM columns of one accumulator)
constexpr int kMatchMaxDim = 40;    // widest descriptor the tensor-core sweep takes (K = 3 dim + 3 <= 128); wider ones are scanned exhaustively
constexpr int kMatchCand = 8;       // candidate columns one epilogue thread keeps (four threads sweep a row)
constexpr int kMatchUnion = 16;     // candidate columns kept per source row (union of the four lists)
int match_chunks(int dim);          // 16-byte K chunks per operand row
int match_configure();
int launch_match_prep(const LaunchCtx& lc, const MatchPair* d_pairs, int pairs, int max_tiles_rows, const float* d_desc,
                      int side, int dim, unsigned char* d_img, float* d_norms, uint32_t* d_bmax);
int launch_match_mma(const LaunchCtx& lc, const MatchPair* d_pairs, int pairs, int max_s_tiles, int dim,
                     const unsigned char* d_img, const float* d_norms, const uint32_t* d_bmax, int32_t* d_cand,
                     int32_t* d_cand_cnt, int dbg);
int launch_match_exact(const LaunchCtx& lc, const MatchPair* d_pairs, int pairs, int max_ns, const float* d_desc_src,
                       const float* d_desc_dst, const float* d_xyz_src, const float* d_xyz_dst, int dim,
                       const int32_t* d_cand, const int32_t* d_cand_cnt, int32_t* d_nn, float* d_corr_src,
                       float* d_corr_dst, int32_t* d_work_list, uint32_t* d_work_count);

// mutual-nearest-neighbour filter of the matched correspondences (kernels_match.cu)
struct MutualPair {
  int32_t Ns, Nd;
  int64_t s_off, d_off;  // first row of the pair in the packed source / target arrays
};
int launch_match_mutual(const LaunchCtx& lc, const MutualPair* d_pairs, int B, const int32_t* d_nn, const int32_t* d_nn_back,
                        const float* d_corr_src, const float* d_corr_dst, unsigned long long* d_kept, long long* d_out_offsets,
                        float* d_out_src, float* d_out_dst);

// kernels_hypo.cu — S4 Kabsch, S5/S6 scoring + argmax, S7 refit
int launch_kabsch(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, const float* d_soa, const int32_t* d_tri,
                  float* d_rt, int K);
int launch_score(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_n, const float* d_soa,
                 const int32_t* d_tri, const float* d_rt, unsigned long long* d_hyp_key, PairDev* d_state, float tau2,
                 int K, int h_begin, int h_end, int mode);
int launch_finalize(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, const float* d_soa, const float* d_rt,
                    const PairDev* d_state, const unsigned long long* d_best_override, uint32_t* d_mask, float* d_R,
                    float* d_t, int32_t* d_inl, float tau2, int K, int refit);

}  // namespace saccot

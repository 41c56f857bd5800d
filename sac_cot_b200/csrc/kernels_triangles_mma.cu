// kernels_triangles_mma.cu — S2 on the tensor cores (K2b of SURVEY.md §2b): the dense contraction
// T = (A·Aᵀ) ∘ A with tcgen05.mma, accumulators in TMEM.  Integer-exact: operands are 0/1 products.
//
// north_star names int8 tiles; profiles/microbench/umma_probe.cu measured one SM at 7.9 k MAC/clk for
// kind::i8 and 15.6 k MAC/clk for kind::mxf4.block_scale, so the kernel uses the 4-bit path.
//
// Operands are produced ON CHIP from the 1-bit adjacency (pre-expanded operands would need > 400 MB of
// L2->SMEM traffic per pair).  The contraction index may be permuted freely as long as both operands use
// the same permutation, so a 32-bit adjacency word w is turned into four words of eight e2m1 nibbles with
//     w & 0x22222222  (nibble 0b0010 = 1.0)      w & 0x11111111  (nibble 0b0001 = 0.5)
//     (w>>2) & 0x22222222                         (w>>2) & 0x11111111
// and the 0.5 blocks carry the UE8M0 block scale 2.0 on BOTH operands: every product is exactly 0 or 1.
// That is 5 ALU instructions per 32 adjacency bits (profiles/microbench/umma_contend.cu checks the result
// against popcounts and shows the MMA pipe at 99.6 % with the shared-memory stores running flat out).
//
// Kernel structure (persistent, one CTA PAIR per two SMs, cta_group::2; 15 warps = 480 threads per CTA so
// that every thread may use 128 registers):
//   tile        = 256 rows (i; 128 per CTA = TMEM lanes) x 240 columns (j), full K; only tiles that hold some
//                 i < j are visited.  The host deals the (pair, J, I)-sorted tile list to the CTA pairs in runs
//                 (api.cu::interleave_tile_runs): few pair changes per CTA pair, L2-resident panels.
//   warp  8     MMA issuer (leader CTA): per stage pair (2 x 256 columns of K) 8 x tcgen05.mma M=256 N=240
//               K=64 by one elected lane, two TMEM accumulators (2 x 240 columns) that start every tile at
//               2^23 (kBias): the epilogue of tile n overlaps the MMAs of tile n+1 and works on integers
//   warps 9-14  expansion (3 groups of 2 warps, group = stage % 3): raw rows of the K-panel copy of A straight
//               from L2 into registers one group-stage ahead -> 6-deep ring of canonical (no-swizzle,
//               K-major) operand stages, handed to the issuer in pairs
//   warps 0-7   epilogue: tcgen05.ld 16x256b (accumulator-fragment layout: a thread holds 4 rows of each of
//               its 8 columns of a 32 x 32 block); mask with the edge bits (j > i, prefetched into shared
//               memory with cp.async one tile ahead), row sums (t2_i), column sums by a 3-level shuffle
//               reduce-scatter (t2_j), accumulator reset to kBias; edges whose T reaches the pair's pruning
//               threshold are staged (key, histogram) in a per-warp shared-memory buffer
//
// Measured limits (profiles/README.md): the MMAs alone would take 15.6 us per N=5000 pair; the kernel runs
// at ~24 us.  With the epilogue switched off it is 22 us and with three quarters of the operand stores
// removed 17.8 us: the expansion warps' STS.128 stream (31.7 KB per stage and SM, next to the tensor cores'
// own operand reads) is what the tensor pipe waits for, not the ALU work and not the epilogue.
//
// Pruning threshold (tri_theta_kernel): exact T of the edges among the ~128 highest-degree nodes; if at
// least K_e of them exist, theta = the K_e-th largest of those counts.  Then at least K_e edges of the
// graph have T >= theta, so no edge below theta can be among the top K_e: dropping them changes nothing
// downstream (selection is exact), but removes ~98 % of the key traffic.  theta = 0 keeps every edge.
#include "common.cuh"

#include <algorithm>
#include <cstdio>

namespace saccot {

namespace {

constexpr int kEpiWarps = 8;                          // warps 0..7: TMEM lane quadrant = warp & 3, column half = warp >> 2
constexpr int kMmaWarp = 8;
constexpr int kProdWarp0 = 9;
constexpr int kProducerWarps = 6;                     // warps 9..14: 15 warps => 128 registers per thread
constexpr int kThreads = 32 * (kProdWarp0 + kProducerWarps);  // 480
constexpr int kCtaM = 128;                            // rows of the pair tile owned by one CTA (one TMEM lane each)
constexpr int kCtaNB = kMmaTileN / 2;                 // 120 rows of the B operand staged by one CTA
constexpr int kStageK = 256;                          // K elements (adjacency columns) per stage
constexpr int kStages = 6;                            // expanded operand stages (handed over in pairs)
constexpr int kPairs = kStages / 2;
constexpr int kRows = kCtaM + kCtaNB;                 // 248 rows expanded per stage and CTA
constexpr int kGroups = kRows / 8;                    // 31 groups of 8 rows
constexpr int kLBO = 128;                             // next 16-byte K chunk (core matrices contiguous along K)
constexpr int kSBO = (kStageK / 2 / 16) * 128;        // next 8-row group: 8 core matrices = 1024 B
constexpr int kStageBytes = kGroups * kSBO;           // 31744
constexpr int kWarpsPerGroup = 2;                     // producer warps that share a stage
constexpr int kGroupsP = kProducerWarps / kWarpsPerGroup;  // producer groups; group = stage % kGroupsP
constexpr int kTasksPerWarp = 16 / kWarpsPerGroup;    // 16 warp tasks (16 rows x 2 quads) per stage, the last one half
// A group must meet every phase of the barriers it waits on in order (mbarrier waits only know the phase
// parity: a waiter two phases early passes at once).  With the group count dividing both ring sizes, a
// group always returns to the same slots.
static_assert(kStages % kGroupsP == 0, "producer groups must divide the ring size");
constexpr int kKeyBuf = 128;                          // staged keys per epilogue warp
constexpr uint32_t kSfCol = 480;                      // scale factors: TMEM columns [480, 512)
constexpr uint32_t kSfWord = 0x807F807Fu;             // UE8M0 per K block of 32: {1.0, 2.0, 1.0, 2.0}
// Every accumulator entry starts at 2^23 (written by the epilogue when it hands a buffer back) and every MMA
// accumulates: the fp32 value 2^23 + T has the bit pattern kBias + T, so the epilogue works on integers
// without a single conversion (T <= 65535 < 2^23: exact).
constexpr uint32_t kBias = 0x4B000000u;

// shared-memory carve-up (dynamic)
constexpr int kOffHist = kStages * kStageBytes;                   // 190464
constexpr int kOffKeys = kOffHist + kHistBins * 4;                // +16384
constexpr int kOffWin = kOffKeys + kEpiWarps * (kKeyBuf + 16) * 8;  // +9216
constexpr int kWinBytes = 2 * kEpiWarps * 5 * 32 * 4;             // edge-bit windows [2][warp][5][lane]
constexpr int kOffBars = kOffWin + kWinBytes;                     // +10240
// barriers: full2[kPairs] empty2[kPairs] tmem_full[2] tmem_empty[2]
constexpr int kNumBars = 2 * kPairs + 2 + 2;
constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16 + kEpiWarps * 32;  // 32 = sizeof(EpiCtx)
static_assert(kSmemBytes <= 232448, "shared memory budget");

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((kLBO >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((kSBO >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell); SWIZZLE_NONE, base offset 0
  return d;
}

// Barrier wait with a watchdog: a protocol bug must end the kernel with an error, not hang the GPU.
// Default (CTA-scope) semantics also for barriers that peer-CTA threads arrive on, as in CUTLASS's
// ClusterBarrier: cluster-scope release/acquire compiles to MEMBAR.ALL.GPU / CCTL.IVALL per arrival
// (measured: 3x slower main loop); what the tensor cores read was made visible by the writers'
// fence.proxy.async before their arrival.
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity, int tag, uint32_t info) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (int spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spins == 0) {
      t0 = clock64();
    } else {
      const long long dt = clock64() - t0;
      if (dt > 1000000000LL && (spins & 0x40000000) == 0) {  // report once, give the other roles time to report too
        if ((threadIdx.x & 31) == 0)
          printf("sac_cot: barrier timeout, cta %d warp %d tag %d parity %u info %u\n", blockIdx.x, threadIdx.x >> 5, tag,
                 parity, info);
        spins |= 0x40000000;
      }
      if (dt > 3000000000LL) __trap();
    }
  }
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// completion of every tcgen05 operation issued so far -> one arrival on `bar` in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair_if(uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3)), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 128 adjacency bits of one row -> four 16-byte K chunks (see the header comment)
__device__ __forceinline__ void expand_quad(const uint4 w, unsigned char* dst) {
  const uint32_t m2 = 0x22222222u, m1 = 0x11111111u;
  const uint32_t sx = w.x >> 2, sy = w.y >> 2, sz = w.z >> 2, sw = w.w >> 2;
  *reinterpret_cast<uint4*>(dst) = make_uint4(w.x & m2, w.y & m2, w.z & m2, w.w & m2);
  *reinterpret_cast<uint4*>(dst + kLBO) = make_uint4(w.x & m1, w.y & m1, w.z & m1, w.w & m1);
  *reinterpret_cast<uint4*>(dst + 2 * kLBO) = make_uint4(sx & m2, sy & m2, sz & m2, sw & m2);
  *reinterpret_cast<uint4*>(dst + 3 * kLBO) = make_uint4(sx & m1, sy & m1, sz & m1, sw & m1);
}

// The fields of a PairDesc a role needs (a full copy per role costs ~20 registers and spills the epilogue).
struct PairLite {
  int32_t N, Npad, stride, npanel;
  int64_t adj_off, node_off, panel_off;
  int32_t nkeep;  // RECT instance: kept nodes of the pair (rows of its tiles index the kept list)
};
// tile entry e = (pair, row block << 16 | column block); d follows the pair.  The roles load the entry of
// their NEXT tile one tile ahead (a dependent L2 round trip per tile otherwise sits in front of every role).
__device__ __forceinline__ void decode_tile(const uint2 e, const PairDesc* __restrict__ descs, int& pair, PairLite& d,
                                            int& I0, int& J0, const NodePlan* __restrict__ plan = nullptr) {
  if (static_cast<int>(e.x) != pair) {
    pair = static_cast<int>(e.x);
    const PairDesc* pd = descs + pair;
    d.nkeep = plan ? static_cast<int32_t>(plan[pair].n_keep) : 0;
    d.N = pd->N;
    d.Npad = pd->Npad;
    d.stride = pd->stride;
    d.npanel = pd->npanel;
    d.adj_off = pd->adj_off;
    d.node_off = pd->node_off;
    d.panel_off = pd->panel_off;
  }
  I0 = static_cast<int>(e.y >> 16) * kMmaTileM;
  J0 = static_cast<int>(e.y & 0xFFFFu) * kMmaTileN;
}

#define SACCOT_TMEM_LD16(v, taddr)                                                                                   \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"       \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                      \
      : "r"(taddr))

#define SACCOT_TMEM_ST16_BIAS(taddr)                                                                          \
  asm volatile(                                                                                               \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), \
      "r"(kBias)                                                                                              \
      : "memory")

// 16 TMEM lanes x 32 columns in the accumulator-fragment layout (tcgen05.ld 16x256b.x4): with t4 = lane & 3,
// t8 = lane >> 2, register 4 b + 2 rh + e holds lane (row) t8 + 8 rh, column 8 b + 2 t4 + e.  A thread then
// owns 2 (4 with both lane halves) rows of each of its 8 columns: column sums need a quarter of the adds and
// three shuffle levels instead of five.
#define SACCOT_TMEM_LDF4(v, taddr)                                                                                   \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"       \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                      \
      : "r"(taddr))

// Per-warp state of the epilogue's rare path (edges that reach the pruning threshold), kept in shared memory:
// it costs the hot loop no registers.
struct EpiCtx {
  unsigned long long* keyp;       // the pair's key list
  unsigned long long* kcount;     // the pair's key counter
  uint32_t thb;                   // pruning threshold + kBias
  uint32_t fill;                  // slots reserved in the warp's buffer (reservations past the end included)
  uint32_t limit;                 // first reserved slot that was turned away (buffer full), else 0xFFFFFFFF
  uint32_t pad_;
};
constexpr int kKeySlots = kKeyBuf + 16;  // a reservation (<= 16 keys) is taken if it STARTS below kKeyBuf

// staged keys of one epilogue warp -> the pair's key list (one reservation, coalesced copy)
__device__ __noinline__ void flush_keys(unsigned long long* kb, EpiCtx* ctx) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  const uint32_t fill = min(ctx->fill, ctx->limit);  // reservations are contiguous: [0, first one turned away)
  if (fill) {
    unsigned long long* keyp = ctx->keyp;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(ctx->kcount, static_cast<unsigned long long>(fill));
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t k = lane; k < fill; k += 32) keyp[base + k] = kb[k];
    __syncwarp();
    if (lane == 0) {
      ctx->fill = 0u;
      ctx->limit = 0xFFFFFFFFu;
    }
  }
  __syncwarp();
}

// Rare path of the epilogue (with a tight theta): some of this thread's 16 masked values m (kBias + T) of a
// fragment half reach the pruning threshold.  One slot reservation per thread in the warp's staging buffer
// (shared-memory atomic); if the buffer is full the keys go straight to the pair's list.  lo0 = key bits of
// the thread's first row and column, (0xFFFF - i) << 16 | (0xFFFF - j); fa / fb = edge bits of its two rows
// (with theta = 0 a masked-off entry, kBias, passes the threshold test too).
// loA / loB: key bits (0xFFFF - i) << 16 | (0xFFFF - j0) of the thread's two rows and first column.  RECT (tiles of a
// node-pruned pair: the rows are kept nodes, every column is visited): an edge shows up from both of its ends, only
// the end with j > i emits it.
template <bool RECT>
__device__ __forceinline__ void push16(const uint32_t (&m)[16], uint32_t fa, uint32_t fb, uint32_t loA, uint32_t loB,
                                       unsigned long long* kb, EpiCtx* ctx, uint32_t* hist_s) {
  const uint32_t thb = ctx->thb;
  uint32_t cand = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (m[k] >= thb && ((((k & 2) ? fb : fa) >> (8 * (k >> 2) + (k & 1))) & 1u)) {
      if constexpr (RECT) {
        const uint32_t lo = ((k & 2) ? loB : loA) - static_cast<uint32_t>(8 * (k >> 2) + (k & 1));
        if ((lo & 0xFFFFu) >= (lo >> 16)) continue;  // j <= i
      }
      cand |= 1u << k;
    }
  if (cand == 0) return;
  const uint32_t nc = __popc(cand);
  const uint32_t pos = atomicAdd(&ctx->fill, nc);
  unsigned long long* dst;  // generic: shared (staging buffer) or global (key list)
  if (pos < static_cast<uint32_t>(kKeyBuf)) {
    dst = kb + pos;
  } else {
    atomicMin(&ctx->limit, pos);
    dst = ctx->keyp + atomicAdd(ctx->kcount, static_cast<unsigned long long>(nc));
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (cand & (1u << k)) {
      const uint32_t T = m[k] - kBias;
      const uint32_t lo = ((k & 2) ? loB : loA) - static_cast<uint32_t>(8 * (k >> 2) + (k & 1));
      *dst++ = (static_cast<unsigned long long>(T) << 32) | lo;
      atomicAdd(&hist_s[T >> 4], 1u);
    }
  }
}

// One warp: moves this CTA's histogram of kept keys (by T >> 4) into the pair's global histogram and finds, in the
// GLOBAL counts, the highest bin b with at least Ke keys in bins >= b: then Ke edges with T >= 16 b exist (every
// counted key is in a key list or in a staging buffer on its way there) and the pair's threshold becomes
// max(threshold, 16 b).  Out of line: it runs once per tile in one warp and must not fatten the hot loop.
__device__ __noinline__ void raise_threshold(uint32_t* hist_s, uint32_t* histp, int cur_bins, int Ke, uint32_t thb,
                                             uint32_t* theta_pair) {
  const int lane = threadIdx.x & 31;
  int carry = 0, found = -1;
  for (int c0 = (cur_bins - 1) & ~31; c0 >= 0 && found < 0; c0 -= 32) {
    const int k = c0 + lane;
    int v = 0;
    if (k < cur_bins) {
      const uint32_t delta = hist_s[k] ? atomicExch(&hist_s[k], 0u) : 0u;
      v = static_cast<int>(delta ? atomicAdd(&histp[k], delta) + delta : __ldcg(&histp[k]));
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {  // suffix sums: lane l gets bins c0 + l .. c0 + 31
      const int u = __shfl_down_sync(0xffffffffu, v, o);
      if (lane + o < 32) v += u;
    }
    const unsigned ok = __ballot_sync(0xffffffffu, carry + v >= Ke);
    if (ok) found = c0 + 31 - __clz(ok);
    carry += __shfl_sync(0xffffffffu, v, 0);
  }
  if (found > 0 && 16u * static_cast<uint32_t>(found) + kBias > thb && lane == 0) atomicMax(theta_pair, 16u * static_cast<uint32_t>(found));
}

}  // namespace

// PROF: experiments only — lane 0 of one warp per role accumulates the cycles it spends waiting on each
// barrier and CTAs 0/1 print them (dbg bit 6).
#define SACCOT_TIMED_WAIT(acc, call)       \
  do {                                     \
    if (PROF) {                            \
      const long long t0__ = clock64();    \
      call;                                \
      acc += clock64() - t0__;             \
    } else {                               \
      call;                                \
    }                                      \
  } while (0)

// RECT instance: tiles of node-pruned pairs (kernels_prune.cu).  A tile is 256 KEPT nodes (positions I0 .. I0 + 255
// of the pair's kept list) x 240 columns; every column block is visited.  The A operand rows come from the compact
// copy `kpanel` of the kept nodes' K-panel records ([panel][kRectRows rows][8 words] per pair, rows past the kept
// count zero), the B operand rows from the pair's panel copy as usual.  The epilogue maps its rows through the kept
// list: edge windows, keys (j > i only), node sums (row sums over ALL columns; no column sums).
template <bool PROF, bool RECT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) triangles_mma_kernel(
    const PairDesc* __restrict__ descs, const uint2* __restrict__ tiles, int total_tiles, const int* __restrict__ d_total,
    const uint2* __restrict__ tiles2,
    const uint32_t* __restrict__ adj, const uint32_t* __restrict__ panel, PairDev* __restrict__ state,
    const ChunkDev* __restrict__ chunk, unsigned long long* __restrict__ keys, uint32_t* __restrict__ theta,
    uint32_t* __restrict__ hist, unsigned long long* __restrict__ t2, int Ke, int raise, int dbg,
    const NodePlan* __restrict__ plan, const unsigned short* __restrict__ kept, const uint32_t* __restrict__ kpanel,
    long long kpanel_pair_words) {
  if (chunk->overflow || !chunk->use_tensor) return;
  // the list was compacted on the device (kernels_prune.cu): fewer tiles than the grid was sized for, possibly none
  if (d_total && *d_total >= 0) {
    total_tiles = *d_total;
    tiles = tiles2;
  }
  if (total_tiles == 0) return;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* stage_base = smem_raw;
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(smem_raw + kOffHist);
  unsigned long long* kbuf = reinterpret_cast<unsigned long long*>(smem_raw + kOffKeys);
  uint32_t* wbuf = reinterpret_cast<uint32_t*>(smem_raw + kOffWin);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kOffBars);
  uint64_t* full2 = bars;                                // [kPairs] producers of both CTAs -> MMA (leader's copy); per stage pair
  uint64_t* empty2 = full2 + kPairs;                     // [kPairs] MMA commit (multicast) -> producers; per stage pair
  uint64_t* tmem_full = empty2 + kPairs;                 // [2] MMA commit (multicast) -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;                  // [2] epilogue warps of both CTAs -> MMA (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  EpiCtx* ectx = reinterpret_cast<EpiCtx*>(tmem_slot + 4);  // [kEpiWarps] rare-path state of the epilogue warps

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;  // cluster = CTA pair = one 256 x 240 tile at a time

  for (int k = tid; k < kHistBins; k += kThreads) hist_s[k] = 0;
  if (tid < kEpiWarps) {
    ectx[tid].fill = 0u;
    ectx[tid].limit = 0xFFFFFFFFu;
  }
  if (tid == 0) {
    for (int s = 0; s < kPairs; ++s) {
      mbar_init(&full2[s], 4 * kWarpsPerGroup);  // two stages x the warps of each stage's group x two CTAs
      mbar_init(&empty2[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;
  if (warp < 4) {  // scale factors: every 32-bit word of columns [480, 512) = {1.0, 2.0, 1.0, 2.0}
    const uint32_t taddr = tmem + ((32u * warp) << 16) + kSfCol;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
            taddr),
        "r"(kSfWord)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if (warp < kEpiWarps) {  // both accumulator buffers start at 2^23 (see kBias)
    const int q = warp & 3, h = warp >> 2;
    const int nch = h ? kMmaTileN / 16 - 8 : 8;
    for (int b = 0; b < 2; ++b)
      for (int c = 0; c < nch; ++c)
        SACCOT_TMEM_ST16_BIAS(tmem + ((32u * q) << 16) + static_cast<uint32_t>(kMmaTileN * b + 128 * h + 16 * c));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // barriers initialised, TMEM allocated, scale factors and accumulator bias written in BOTH CTAs before
  // anyone proceeds
  asm volatile("tcgen05.fence::before_thread_sync;");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;");

  if (warp < kEpiWarps) {
    // ================================== epilogue warps ==================================
    // warp = 4 h + q: TMEM lanes 32 q .. 32 q + 31 (rows), columns 128 h .. 128 h + 127 of the accumulator
    // (h = 1: the last 16 lie outside the tile and are masked off), four units of 32 columns each.
    const int q = warp & 3, h = warp >> 2;
    const int t4 = lane & 3, t8 = lane >> 2;
    unsigned long long* kb = kbuf + warp * kKeySlots;
    EpiCtx* ctx = ectx + warp;
    int cur_pair = -1, cur_bins = 0;
    bool dyn = false;  // this pair's threshold may still be raised while the kernel runs
    uint32_t thb = kBias;
    auto flush_pair = [&]() {  // all epilogue warps
      flush_keys(kb, ctx);
      asm volatile("bar.sync 2, 256;" ::: "memory");
      uint32_t* histp = hist + static_cast<size_t>(cur_pair) * kHistBins;
      for (int k = tid; k < cur_bins; k += 32 * kEpiWarps) {  // T <= N - 2: higher bins are empty
        const uint32_t v = hist_s[k];
        if (v) {
          atomicAdd(&histp[k], v);
          hist_s[k] = 0;
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
    };
    // Raw edge-bit window of this thread's row: 5 words starting at the word that holds column J0 + 128 h.
    // Copied asynchronously into the thread's own slots of a double-buffered shared-memory area one tile
    // ahead: as register loads the compiler spilled them right after issue, which exposed the full
    // global-load latency on every tile.
    uint32_t* wmine = wbuf + warp * (5 * 32) + lane;
    int i_next = 0;  // RECT: original index of this thread's row in the tile whose window is being fetched
    auto load_window = [&](const PairLite& pd, int I0, int J0, int slot, int pair_of) {
      int i = I0 + kCtaM * static_cast<int>(rank) + 32 * q + lane;
      bool row_ok = i < pd.Npad;
      if constexpr (RECT) {
        row_ok = i < pd.nkeep;
        i = row_ok ? static_cast<int>(__ldg(kept + static_cast<size_t>(pair_of) * kNodeKeepMax + i)) : 0;
        i_next = i;
      }
      const uint32_t* rowp = adj + pd.adj_off + static_cast<size_t>(row_ok ? i : 0) * pd.stride;
      const int w0 = (J0 >> 5) + 4 * h;
      const uint32_t dst = smem_u32(wmine + slot * (kEpiWarps * 5 * 32));
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const bool ok = row_ok && w0 + k < pd.stride;  // not ok: nothing is read, the slot is filled with zeros
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 128u * k), "l"(rowp + (ok ? w0 + k : 0)),
                     "r"(ok ? 4u : 0u)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int n = 0, pair = -1;
    PairLite d;
    long long w_tfull = 0, t_ldw = 0, t_emit = 0, t_pairchg = 0, t_win = 0, t_fetch = 0, t_tail = 0;
    const long long t_begin = PROF ? clock64() : 0;
    int t = cid, I0 = 0, J0 = 0;
    uint2 en = make_uint2(0u, 0u);  // entry of the tile after this one
    if (t < total_tiles) {
      decode_tile(tiles[t], descs, pair, d, I0, J0, RECT ? plan : nullptr);
      load_window(d, I0, J0, 0, pair);
      if (t + ncl < total_tiles) en = tiles[t + ncl];
    }
    for (; t < total_tiles; ++n) {
      // ---- this tile ----
      const int tJ0 = J0;
      const long long node_off = d.node_off;
      const int irow = i_next;  // RECT: this tile's row index (fetched with its window one tile ago)
      const long long tq0 = PROF ? clock64() : 0;
      if (pair != cur_pair) {
        if (cur_pair >= 0) flush_pair();
        cur_pair = pair;
        cur_bins = min(kHistBins, (d.N >> 4) + 1);
        // bit 31: tri_theta_kernel certified the threshold with a fat sample (>= 4 K_e edges): tight, leave it alone
        const uint32_t traw = theta[pair];
        dyn = raise != 0 && (traw >> 31) == 0u;
        thb = (traw & 0x7FFFFFFFu) + kBias;
        __syncwarp();
        if (lane == 0) {
          ctx->keyp = keys + state[pair].key_base;
          ctx->kcount = &state[pair].key_count;
          ctx->thb = thb;
        }
        __syncwarp();
      }
      // the pair's threshold as it stands now: other CTA pairs (and this one, below) raise it while they run
      const uint32_t theta_now = dyn ? __ldcg(theta + cur_pair) : 0u;
      const long long tq1 = PROF ? clock64() : 0;
      uint32_t win[4];
      uint32_t wraw[5];
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 5; ++k) wraw[k] = wmine[(n & 1) * (kEpiWarps * 5 * 32) + 32 * k];
      if (tJ0 & 16) {
#pragma unroll
        for (int k = 0; k < 4; ++k) win[k] = __funnelshift_r(wraw[k], wraw[k + 1], 16);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) win[k] = wraw[k];
      }
      if (h) win[3] &= 0xFFFFu;  // columns 240..255 of the window belong to the next J-block
      const int rbase = I0 + kCtaM * static_cast<int>(rank) + 32 * q;  // first row of this warp
      const int i = rbase + lane;
      const int cbase = tJ0 + 128 * h;  // first column of this warp's half
      if constexpr (!RECT) {
        const int dd = i - cbase;  // keep j > i: clear window bits 0..dd
        if (dd >= 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int lo = 32 * k;
            if (dd >= lo + 31) win[k] = 0u;
            else if (dd >= lo) win[k] &= 0xFFFFFFFEu << (dd - lo);
          }
        }
      }
      const long long tq2 = PROF ? clock64() : 0;
      // ---- next tile: its entry arrived during the previous tile; start loading its window now (hidden
      //      behind this tile's work) and fetch the entry of the tile after it ----
      t += ncl;
      if (t < total_tiles) {
        decode_tile(en, descs, pair, d, I0, J0, RECT ? plan : nullptr);
        load_window(d, I0, J0, (n + 1) & 1, pair);
        if (t + ncl < total_tiles) en = tiles[t + ncl];
      }

      const long long tq3 = PROF ? clock64() : 0;
      if (PROF) {
        t_pairchg += tq1 - tq0;
        t_win += tq2 - tq1;
        t_fetch += tq3 - tq2;
      }
      // ---- accumulators of tile n ----
      const int buf = n & 1;
      SACCOT_TIMED_WAIT(w_tfull, mbar_wait_wd(&tmem_full[buf], static_cast<uint32_t>((n >> 1) & 1), 1, n));
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t tbase = tmem + ((32u * q) << 16) + static_cast<uint32_t>(kMmaTileN * buf + 128 * h);
      if (theta_now + kBias > thb) {  // warp-uniform
        thb = theta_now + kBias;
        __syncwarp();
        if (lane == 0) ctx->thb = thb;
        __syncwarp();
      }
      // biased row sums of the four fragment rows t8 + 8 s of this thread (32 values each per tile)
      uint32_t rs0 = 0, rs1 = 0, rs2 = 0, rs3 = 0;
      uint32_t v[2][16];
      // 16 rows x 32 columns in fragment layout: masked values (kBias + T on an edge j > i, kBias elsewhere),
      // row sums, the sums of this thread's two rows per column.  lo0 = key bits of (row t8 of the half, column
      // 2 t4 of the unit): (0xFFFF - i) << 16 | (0xFFFF - j).
      auto half16 = [&](const uint32_t(&vv)[16], uint32_t fa, uint32_t fb, uint32_t& ra, uint32_t& rb, uint32_t(&c)[8],
                        uint32_t loA, uint32_t loB, bool first) {
        uint32_t m[16];
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            m[4 * b + e] = ((fa >> (8 * b + e)) & 1u) ? vv[4 * b + e] : kBias;
            m[4 * b + 2 + e] = ((fb >> (8 * b + e)) & 1u) ? vv[4 * b + 2 + e] : kBias;
          }
        ra += ((m[0] + m[1] + m[4]) + (m[5] + m[8] + m[9])) + (m[12] + m[13]);
        rb += ((m[2] + m[3] + m[6]) + (m[7] + m[10] + m[11])) + (m[14] + m[15]);
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t cc = m[4 * b + e] + m[4 * b + 2 + e];
            c[2 * b + e] = first ? cc : c[2 * b + e] + cc;
          }
        const uint32_t a = max(max(max(m[0], m[1]), max(m[2], m[3])), max(max(m[4], m[5]), max(m[6], m[7])));
        const uint32_t bb = max(max(max(m[8], m[9]), max(m[10], m[11])), max(max(m[12], m[13]), max(m[14], m[15])));
        // rare (with a tight theta): one of this thread's 16 edges reaches the pruning threshold
        if (max(a, bb) >= thb) {
          const long long te0 = PROF ? clock64() : 0;
          push16<RECT>(m, fa, fb, loA, loB, kb, ctx, hist_s);
          if (PROF) t_emit += clock64() - te0;
        }
      };
      // RECT: key bits of the original indices of the thread's four fragment rows (rows t8 + 8 s of the warp's 32)
      uint32_t rowbits[4] = {0u, 0u, 0u, 0u};
      if constexpr (RECT) {
#pragma unroll
        for (int sidx = 0; sidx < 4; ++sidx)
          rowbits[sidx] = (0xFFFFu - static_cast<uint32_t>(__shfl_sync(0xffffffffu, irow, t8 + 8 * sidx))) << 16;
      }
      if (!(dbg & 1)) {
        SACCOT_TMEM_LDF4(v[0], tbase);
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          // edge bits of the four fragment rows for these 32 columns, aligned to this thread's columns
          const uint32_t wown = win[0];
          const uint32_t f0 = __shfl_sync(0xffffffffu, wown, t8) >> (2 * t4);
          const uint32_t f1 = __shfl_sync(0xffffffffu, wown, t8 + 8) >> (2 * t4);
          const uint32_t f2 = __shfl_sync(0xffffffffu, wown, t8 + 16) >> (2 * t4);
          const uint32_t f3 = __shfl_sync(0xffffffffu, wown, t8 + 24) >> (2 * t4);
          uint32_t c[8];
          // key bits of (row t8 of the half, column 2 t4 of the unit); consecutive rows in the square tiles
          const uint32_t lo0 = ((0xFFFFu - static_cast<uint32_t>(rbase + t8)) << 16) |
                               (0xFFFFu - static_cast<uint32_t>(cbase + 32 * u + 2 * t4));
          const uint32_t locol = lo0 & 0xFFFFu;
          SACCOT_TIMED_WAIT(t_ldw, asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"));
          SACCOT_TMEM_LDF4(v[1], tbase + (16u << 16) + 32u * u);
          if constexpr (RECT) half16(v[0], f0, f1, rs0, rs1, c, rowbits[0] | locol, rowbits[1] | locol, true);
          else half16(v[0], f0, f1, rs0, rs1, c, lo0, lo0 - (8u << 16), true);
          SACCOT_TIMED_WAIT(t_ldw, asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"));
          if (u + 1 < 4) SACCOT_TMEM_LDF4(v[0], tbase + 32u * (u + 1));
          if constexpr (RECT) half16(v[1], f2, f3, rs2, rs3, c, rowbits[2] | locol, rowbits[3] | locol, false);
          else half16(v[1], f2, f3, rs2, rs3, c, lo0 - (16u << 16), lo0 - (24u << 16), false);
          // column sums (t2_j): c[2 b + e] covers 4 of the 32 rows; reduce-scatter over the lanes that share t4
          // (RECT: a kept node's sum is its row sum over all columns, nothing is collected per column)
          if constexpr (!RECT) {
          {
            const bool hi = (lane & 16) != 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t send = hi ? c[k] : c[k + 4], keep = hi ? c[k + 4] : c[k];
              c[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
          }
          {
            const bool hi = (lane & 8) != 0;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint32_t send = hi ? c[k] : c[k + 2], keep = hi ? c[k + 2] : c[k];
              c[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
          }
          {
            const bool hi = (lane & 4) != 0;
            const uint32_t send = hi ? c[0] : c[1], keep = hi ? c[1] : c[0];
            // 32 rows x kBias = 0x60000000 (mod 2^32); the thread ends up with column 2 b + e = t8
            const uint32_t cs = keep + __shfl_xor_sync(0xffffffffu, send, 4) - 32u * kBias;
            if (cs)
              atomicAdd(&t2[node_off + cbase + 32 * u + 8 * (t8 >> 1) + 2 * t4 + (t8 & 1)], static_cast<unsigned long long>(cs));
          }
          }
          // staged keys: drain the buffer once it is half full (warp-uniform: every push happened before the barrier)
          __syncwarp();
          if (ctx->fill >= static_cast<uint32_t>(kKeyBuf / 2)) flush_keys(kb, ctx);
          // the columns go back to 2^23 for the tile after next (h = 1: the last 16 are not this tile's)
          SACCOT_TMEM_ST16_BIAS(tbase + 32u * u);
          if (!(h && u == 3)) SACCOT_TMEM_ST16_BIAS(tbase + 32u * u + 16u);
          win[0] = win[1];  // next 32 columns of the window
          win[1] = win[2];
          win[2] = win[3];
        }
      }
      // accumulator buffer is free (and back at its bias) again; the MMA issuer lives in the leader CTA
      const long long tq4 = PROF ? clock64() : 0;
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[buf], 0u);
      // Dynamic pruning threshold.  hist_s counts (by T >> 4) the keys this CTA has kept for the pair since it last
      // changed pairs; if K_e of them lie in bins >= b, then K_e edges with T >= 16 b exist and nothing below 16 b can
      // be selected: raise the pair's threshold for everybody.  One warp per tile takes its turn.  (The sample-based
      // threshold of tri_theta_kernel is loose when the inliers are few or the 512-column degree proxy misses them.)
      if (dyn && warp == (n & (kEpiWarps - 1))) raise_threshold(hist_s, hist + static_cast<size_t>(cur_pair) * kHistBins, cur_bins, Ke, thb, theta + cur_pair);
      if (!(dbg & 1)) {
        // row sums (t2_i): reduce-scatter over the four lanes that share t8; the thread ends up with row t8 + 8 t4
        const bool hi2 = (lane & 2) != 0, hi1 = (lane & 1) != 0;
        const uint32_t a0 = (hi2 ? rs2 : rs0) + __shfl_xor_sync(0xffffffffu, hi2 ? rs0 : rs2, 2);
        const uint32_t a1 = (hi2 ? rs3 : rs1) + __shfl_xor_sync(0xffffffffu, hi2 ? rs1 : rs3, 2);
        // 128 values x kBias = 0x80000000 (mod 2^32)
        const uint32_t rowsum = (hi1 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, hi1 ? a0 : a1, 1) - 128u * kBias;
        int trow = rbase + t8 + 8 * t4;
        if constexpr (RECT) trow = __shfl_sync(0xffffffffu, irow, t8 + 8 * t4);
        if (rowsum) atomicAdd(&t2[node_off + trow], static_cast<unsigned long long>(rowsum));
      }
      if (PROF) t_tail += clock64() - tq4;
    }
    if (cur_pair >= 0) flush_pair();
    if (PROF && blockIdx.x < 2 && tid == 0)
      printf("cta %d epilogue: tiles %d total %lld wait_tmem_full %lld wait_ld %lld emit %lld pairchg %lld win %lld "
             "fetch %lld tail %lld\n",
             blockIdx.x, n, clock64() - t_begin, w_tfull, t_ldw, t_emit, t_pairchg, t_win, t_fetch, t_tail);
  } else if (warp == kMmaWarp) {
    // ================================== MMA issuer (leader CTA only) ==================================
    // One barrier wait and one commit per PAIR of stages: the issuing thread (not the tensor pipe) was the
    // limit with a hand-over per stage.
    if (rank == 0) {
      // instruction descriptor: block-scaled, A/B = E2M1, UE8M0 scales, N = 240, M = 256 (2 x 128), K-major both
      const uint32_t idesc = (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kMmaTileN >> 3) << 17) | (1u << 23) | ((256u >> 4) << 24);
      const uint64_t desc0 = umma_desc(smem_u32(stage_base));
      const uint32_t desc_lo = static_cast<uint32_t>(desc0), desc_hi = static_cast<uint32_t>(desc0 >> 32);
      int pair = -1, n = 0;
      uint32_t g = 0;
      PairLite d;
      uint32_t leader;  // 1 in the one lane that issues the tensor-core instructions
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(leader));
      long long w_full = 0, w_tempty = 0, t_begin = PROF ? clock64() : 0;
      uint2 en = cid < total_tiles ? tiles[cid] : make_uint2(0u, 0u);
      for (int t = cid; t < total_tiles; t += ncl, ++n) {
        int I0, J0;
        const uint2 e = en;
        if (t + ncl < total_tiles) en = tiles[t + ncl];
        decode_tile(e, descs, pair, d, I0, J0);
        const int np2 = (d.npanel + 1) >> 1;  // stage pairs (an odd last stage is padded with zeros)
        const int buf = n & 1;
        if (n >= 2) {
          SACCOT_TIMED_WAIT(w_tempty, mbar_wait_wd(&tmem_empty[buf], static_cast<uint32_t>(((n >> 1) - 1) & 1), 2, n));
          asm volatile("tcgen05.fence::after_thread_sync;");
        }
        for (int ip = 0; ip < np2; ++ip, g += 2) {
          const uint32_t pr = (g >> 1) % kPairs;
          SACCOT_TIMED_WAIT(w_full, mbar_wait_wd(&full2[pr], (g / kStages) & 1u, 3, g));
          asm volatile("tcgen05.fence::after_thread_sync;");
          // Every lane runs this code and one elected lane issues: inside an `if (lane == 0)` branch the
          // compiler paid an ELECT + PLOP3 + R2UR.BROADCAST sequence per operand, ~17 instructions per MMA
          // (now ~11), which made the issuing thread as slow as the tensor pipe itself.
          // shared-memory descriptors: only the 14-bit start-address field (16-byte units) changes
          const uint32_t lo0 = desc_lo + pr * ((2 * kStageBytes) >> 4);
#pragma unroll
          for (int ks = 0; ks < 2 * kStageK / 64; ++ks) {
            const uint32_t loA = lo0 + static_cast<uint32_t>(((ks >> 2) * kStageBytes + (ks & 3) * 2 * kLBO) >> 4);
            const uint32_t loB = loA + static_cast<uint32_t>(((kCtaM / 8) * kSBO) >> 4);
            asm volatile(
                "{\n.reg .b64 da, db;\n.reg .pred p, q;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
                "setp.ne.b32 q, %5, 0;\nsetp.eq.u32 p, 1, 1;\n"
                "@q tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], da, db, %4, [%6], [%7], p;\n}" ::"r"(
                    tmem + static_cast<uint32_t>(kMmaTileN * buf)),
                "r"(loA), "r"(loB), "r"(desc_hi), "r"(idesc), "r"(leader), "r"(tmem + kSfCol), "r"(tmem + kSfCol + 16u)
                : "memory");
          }
          umma_commit_pair_if(&empty2[pr], leader);                         // both stages reusable in both CTAs
          if (ip == np2 - 1) umma_commit_pair_if(&tmem_full[buf], leader);  // accumulators complete in both CTAs
          __syncwarp();
        }
      }
      if (PROF && blockIdx.x < 2 && lane == 0)
        printf("cta %d mma: stages %u total %lld wait_full %lld wait_tmem_empty %lld\n", blockIdx.x, g, clock64() - t_begin,
               w_full, w_tempty);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
  } else {
    // ================================== expansion warps ==================================
    // Stage g is expanded by group g % kGroupsP (kWarpsPerGroup warps, kTasksPerWarp 16-row tasks each): the
    // fixed cost of a stage hand-over (barrier wait, proxy fence, arrival; ~350 cycles measured) is paid once
    // per kGroupsP stages by every warp instead of once per stage.
    // The raw adjacency bits come straight from the K-panel copy in L2 into registers, one group-stage ahead
    // (two register sets, loop unrolled by two).  An earlier version staged them in shared memory with bulk
    // copies; the kernel is bound by shared-memory bandwidth (operand stores 31.7 KB + tensor-core operand reads
    // 31.7 KB per stage and SM against 128 B/clk), and the raw ring added a quarter to that traffic.
    const int pw = warp - kProdWarp0;
    const int grp = pw / kWarpsPerGroup, wq = pw % kWarpsPerGroup;
    const int r8 = lane & 7, rg2 = (lane >> 3) & 1, q = lane >> 4;
    long long w_empty = 0, t_sts = 0, t_fence = 0, t_arr = 0, t_begin = PROF ? clock64() : 0;
    // ---- load cursor: the group's next stage (tile lt, stage lit of it) ----
    // Task k of this lane is row rbase + 32 k of the stage image: k < 4 are A rows, k >= 4 the B rows rbase + 32 (k - 4).
    // Everything that depends on the tile only (this lane's two row pointers, which of its rows exist) is worked out
    // when the tile is entered; a stage then costs two pointer offsets and eight predicated loads at constant
    // distances.  (Before, every task re-derived its row, operand and pointer: ~170 instructions per stage, a third
    // of an expansion warp's time by the in-kernel counters, on the warps the tensor pipe waits for.)
    static_assert(kWarpsPerGroup == 2 && kTasksPerWarp == 8 && kCtaM == 128 && kCtaNB <= 128, "task -> row mapping");
    const int rbase = 8 * (2 * wq + rg2) + r8;
    int lt = cid, lit = grp, lnp = 0, lnpanel = 0, lNpad = 0, lpair = -1;
    uint32_t strideA = 0, strideB = 0;  // uint4 units from one K panel to the next
    bool okA = false;                   // this CTA's A rows exist (a whole row block does or does not)
    uint32_t okB = 0;                   // bit k: B row rbase + 32 k exists
    const uint4* lbase = nullptr;       // the pair's K-panel copy
    const uint4* pA = nullptr;          // panel 0, this lane's first A row and quad
    const uint4* pB = nullptr;          // panel 0, this lane's first B row and quad
    uint2 len = make_uint2(0u, 0u);
    auto enter_tile = [&](const uint2 e) {  // tile lt
      if (static_cast<int>(e.x) != lpair) {
        lpair = static_cast<int>(e.x);
        const PairDesc* pd = descs + lpair;
        lNpad = pd->Npad;
        lnpanel = pd->npanel;
        lnp = (lnpanel + 1) & ~1;
        lbase = reinterpret_cast<const uint4*>(panel + pd->panel_off);
        strideB = 2u * static_cast<uint32_t>(lNpad);
        strideA = RECT ? 2u * static_cast<uint32_t>(kRectRows) : strideB;
      }
      const int I0 = static_cast<int>(e.y >> 16) * kMmaTileM, J0 = static_cast<int>(e.y & 0xFFFFu) * kMmaTileN;
      const int a0 = I0 + kCtaM * static_cast<int>(rank), b0 = J0 + kCtaNB * static_cast<int>(rank);
      // rows past the end of the pair expand to zeros
      const int rowsB = min(kCtaNB, lNpad - b0) - rbase;  // B row rbase + 32 k exists iff 32 k < rowsB
      okB = (rowsB > 0 ? 1u : 0u) | (rowsB > 32 ? 2u : 0u) | (rowsB > 64 ? 4u : 0u) | (rowsB > 96 ? 8u : 0u);
      pB = lbase + static_cast<size_t>(b0 + rbase) * 2 + q;
      if constexpr (RECT) {  // A rows: the compact copy of the kept nodes' records (whole row blocks, zero padded)
        okA = true;
        pA = reinterpret_cast<const uint4*>(kpanel + static_cast<long long>(lpair) * kpanel_pair_words) +
             static_cast<size_t>(a0 + rbase) * 2 + q;
      } else {
        okA = a0 < lNpad;
        pA = lbase + static_cast<size_t>(a0 + rbase) * 2 + q;
      }
    };
    if (lt < total_tiles) {
      enter_tile(tiles[lt]);
      if (lt + ncl < total_tiles) len = tiles[lt + ncl];
      while (lt < total_tiles && lit >= lnp) {  // (lnp >= 2 > grp only fails for kGroupsP > 2)
        lit -= lnp;
        lt += ncl;
        if (lt < total_tiles) {
          enter_tile(len);
          if (lt + ncl < total_tiles) len = tiles[lt + ncl];
        }
      }
    }
    // loads the raw bits of the group's next stage; false when the group has no stage left
    auto load_stage = [&](uint4(&w)[kTasksPerWarp]) -> bool {
      if (lt >= total_tiles) return false;
      const bool real = lit < lnpanel;  // the padding stage of an odd panel count carries no data
      const uint4* a = pA + static_cast<size_t>(static_cast<uint32_t>(lit)) * strideA;  // panel lit
      const uint4* b = pB + static_cast<size_t>(static_cast<uint32_t>(lit)) * strideB;
      const bool ra = real && okA;
      const uint32_t rb = real ? okB : 0u;
      if (__all_sync(0xffffffffu, ra && rb == 0xFu)) {  // every row of every lane exists (all but the edge tiles)
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 32 rows further = 64 uint4
          w[k] = __ldg(a + 64 * k);
          w[4 + k] = __ldg(b + 64 * k);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          w[k] = ra ? __ldg(a + 64 * k) : make_uint4(0u, 0u, 0u, 0u);
          w[4 + k] = ((rb >> k) & 1u) ? __ldg(b + 64 * k) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
      lit += kGroupsP;
      while (lt < total_tiles && lit >= lnp) {
        lit -= lnp;
        lt += ncl;
        if (lt < total_tiles) {
          enter_tile(len);
          if (lt + ncl < total_tiles) len = tiles[lt + ncl];
        }
      }
      return true;
    };
    uint32_t g = static_cast<uint32_t>(grp);
    auto store_stage = [&](const uint4(&w)[kTasksPerWarp]) {
      const uint32_t s = g % kStages;
      if (g >= kStages) SACCOT_TIMED_WAIT(w_empty, mbar_wait_wd(&empty2[s >> 1], ((g / kStages) - 1) & 1u, 6, g));
      unsigned char* st = stage_base + s * kStageBytes;
      const long long tp0 = PROF ? clock64() : 0;
#pragma unroll
      for (int k = 0; k < kTasksPerWarp; ++k) {
        const int gr = 2 * (wq + kWarpsPerGroup * k) + rg2;
        if (gr < kGroups) expand_quad(w[k], st + gr * kSBO + 4 * q * kLBO + r8 * 16);
      }
      const long long tp1 = PROF ? clock64() : 0;
      // generic-proxy writes -> visible to the tensor cores (async proxy), then one arrival per warp on the
      // leader's barrier: its MMAs read this CTA's stage too
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      const long long tp2 = PROF ? clock64() : 0;
      if (lane == 0) mbar_arrive_cluster(&full2[s >> 1], 0u);
      if (PROF) {
        t_sts += tp1 - tp0;
        t_fence += tp2 - tp1;
        t_arr += clock64() - tp2;
      }
      g += kGroupsP;
    };
    uint4 w0[kTasksPerWarp], w1[kTasksPerWarp];
    bool have = load_stage(w0);
    while (have) {
      const bool have1 = load_stage(w1);
      store_stage(w0);
      if (!have1) break;
      have = load_stage(w0);
      store_stage(w1);
    }
    if (PROF && blockIdx.x < 2 && lane == 0 && (pw == 0 || pw == kProducerWarps - 1))
      printf("cta %d producer %d: stages %u total %lld wait_empty %lld expand+sts %lld fence %lld arrive %lld\n", blockIdx.x, pw,
             g, clock64() - t_begin, w_empty, t_sts, t_fence, t_arr);
  }

  // no CTA of the pair may exit (or free its TMEM) while the other can still touch its memory
  asm volatile("tcgen05.fence::before_thread_sync;");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// ------------------------------------------------------------------------------------------
// Pruning threshold per pair (see the header comment).  One CTA of 1024 threads per pair; every phase is
// arranged so that its global loads are independent (the kernel is latency-, not throughput-bound).
// ------------------------------------------------------------------------------------------
constexpr int kThetaNodes = 128;
constexpr int kThetaSamples = kThetaNodes * kThetaNodes;            // dense [x][y] table, x < y used
constexpr int kThetaChunkW = 192;                                   // adjacency words per staged row chunk

static size_t theta_smem_bytes(int max_npad) {
  const int cw = std::min(max_npad / 32, kThetaChunkW);
  return static_cast<size_t>(max_npad) + static_cast<size_t>(kThetaNodes) * cw * 4 + kThetaSamples * 2 + kThetaNodes * 16 + 16;
}

// block-wide count of a per-element predicate over n entries (1024 threads)
template <typename Pred>
__device__ __forceinline__ int theta_block_count(Pred pred, int n, int* s_cnt) {
  const int t = threadIdx.x, lane = t & 31;
  if (t == 0) *s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int k = t; k < n; k += 1024) c += pred(k) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0 && c) atomicAdd(s_cnt, c);
  __syncthreads();
  const int r = *s_cnt;
  __syncthreads();
  return r;
}

// Steps 1-4a: the sample nodes (the S = min(N, 128) rows with the largest proxy degree, ties by index), the edges
// among them as a bit mask per sample row (y > x) and their number.  Returns the number of sample nodes; *nts_out =
// the number of sample edges.  deg: [Npad] bytes, nodes: [128], emask: [128][4], all in shared memory.
__device__ __forceinline__ int theta_sample(const PairDesc& d, const uint32_t* __restrict__ adjp, unsigned char* deg,
                                            int* nodes, uint32_t* emask, int* s_cnt, int* s_nsel, int* s_nts,
                                            int* nts_out) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int S = d.N < kThetaNodes ? d.N : kThetaNodes;
  // 1. proxy degree: popcount of the first min(stride, 16) words (512 columns) of every row
  {
    const int nq = min(d.stride, 16) / 4;  // stride is a multiple of 4
    for (int i = t; i < d.Npad; i += 1024) {
      int c = 0;
      if (i < d.N) {
        const uint4* rp = reinterpret_cast<const uint4*>(adjp + static_cast<size_t>(i) * d.stride);
        for (int k = 0; k < nq; ++k) {
          const uint4 w = rp[k];
          c += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
        }
      }
      deg[i] = static_cast<unsigned char>(min(c, 255));
    }
  }
  if (t == 0) *s_nts = 0;
  __syncthreads();
  // 2. largest degree threshold that still leaves >= S nodes, and the number of nodes above it: one histogram
  //    pass over the (saturated, 8-bit) proxy degrees and a suffix scan by one warp.  (Eight block-wide counting
  //    passes of a binary search before: ~15 us of a single N = 5000 pair.)
  __shared__ int s_dh[256];
  __shared__ int s_thr, s_above;
  if (t < 256) s_dh[t] = 0;
  __syncthreads();
  for (int k = t; k < d.N; k += 1024) atomicAdd(&s_dh[deg[k]], 1);
  __syncthreads();
  if (warp == 0) {
    // lane l owns degrees 8 l .. 8 l + 7; suffix sums from the top
    int mine[8], tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mine[k] = s_dh[8 * lane + k];
      tot += mine[k];
    }
    int above = 0;  // nodes with a degree in a higher lane's range
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_down_sync(0xffffffffu, tot, o);
      if (lane + o < 32) tot += u;
    }
    above = tot;
#pragma unroll
    for (int k = 0; k < 8; ++k) above -= mine[k];
    // the threshold lies in the one lane where the suffix count crosses S
    if (above < S && tot >= S) {
      int cum = above;
      for (int k = 7; k >= 0; --k) {
        if (cum + mine[k] >= S) {
          s_thr = 8 * lane + k;
          s_above = cum;
          break;
        }
        cum += mine[k];
      }
    }
  }
  __syncthreads();
  const int thr = s_thr;
  const int n_above = s_above;
  // 3. the S sample nodes in index order: every node above the threshold, then ties until S are taken.  The slot of
  //    a selected node = nodes above the threshold before it + min(ties before it, quota).  Every warp takes a
  //    contiguous segment of the rows: counts first, then (after a prefix over the warps) the slots.  (One warp
  //    walking all N rows was ~100 us of the single N = 50 000 pair's 380.)
  {
    __shared__ int s_wa[32], s_wt[32];
    const int seg = ((d.N + 31) / 32 + 31) & ~31;  // rows per warp, a multiple of 32
    const int r0 = warp * seg, r1 = min(d.N, r0 + seg);
    int na = 0, nt = 0;
    for (int i0 = r0; i0 < r1; i0 += 32) {
      const int i = i0 + lane;
      const int dg = i < r1 ? deg[i] : -1;
      na += __popc(__ballot_sync(0xffffffffu, dg > thr));
      nt += __popc(__ballot_sync(0xffffffffu, dg == thr));
    }
    if (lane == 0) {
      s_wa[warp] = na;
      s_wt[warp] = nt;
    }
    __syncthreads();
    int a_before = 0, t_before = 0;
    for (int w = 0; w < warp; ++w) {
      a_before += s_wa[w];
      t_before += s_wt[w];
    }
    const int quota = S - n_above;  // ties to take (>= 0 by the choice of thr)
    for (int i0 = r0; i0 < r1; i0 += 32) {
      const int i = i0 + lane;
      const int dg = i < r1 ? deg[i] : -1;
      const unsigned amask = __ballot_sync(0xffffffffu, dg > thr);
      const unsigned tmask = __ballot_sync(0xffffffffu, dg == thr);
      const unsigned below = (1u << lane) - 1u;
      const int a_here = a_before + __popc(amask & below), t_here = t_before + __popc(tmask & below);
      const bool take = dg > thr || (dg == thr && t_here < quota);
      const int slot = a_here + min(t_here, quota);
      if (take && slot < kThetaNodes) nodes[slot] = i;
      a_before += __popc(amask);
      t_before += __popc(tmask);
    }
    if (t == 0) {
      int ta = 0, tt = 0;
      for (int w = 0; w < 32; ++w) {
        ta += s_wa[w];
        tt += s_wt[w];
      }
      *s_nsel = min(kThetaNodes, ta + min(tt, quota));
    }
  }
  __syncthreads();
  const int nsel = *s_nsel;
  // 4a. edges among the sample nodes: a bit mask per sample row
  for (int k = t; k < kThetaNodes * 4; k += 1024) emask[k] = 0u;
  __syncthreads();
  for (int base = 0; base < nsel * nsel; base += 1024) {
    const int idx = base + t;
    const int x = idx / nsel, y = idx - x * nsel;
    bool is_edge = false;
    if (idx < nsel * nsel && x < y) {
      const int a = nodes[x], b = nodes[y];
      is_edge = ((adjp[static_cast<size_t>(a) * d.stride + (b >> 5)] >> (b & 31)) & 1u) != 0u;
    }
    if (is_edge) atomicOr(&emask[x * 4 + (y >> 5)], 1u << (y & 31));
    const unsigned em = __ballot_sync(0xffffffffu, is_edge);
    if (lane == 0 && em) atomicAdd(s_nts, __popc(em));
  }
  __syncthreads();
  *nts_out = *s_nts;
  return nsel;
}

// Step 4b for the adjacency words [c0, c0 + cw): the sample rows are staged in shared memory; a warp keeps sample
// row x in registers (six words per lane at most) and walks its partners y > x: per edge 6 LDS, AND, a carry-save
// adder tree (3 POPC instead of 6: POPC and REDUX share the slow XU pipe) and one warp reduction.  Rows are dealt so
// that every warp gets the same number of partners (x, 63 - x, 64 + x, 127 - x).  add(x, y, c) receives the count.
template <typename Add>
__device__ __forceinline__ void theta_count_chunk(const PairDesc& d, const uint32_t* __restrict__ adjp, const int* nodes,
                                                  const uint32_t* emask, uint32_t* rows_s, int nsel, int c0, int cw,
                                                  int cw_max, Add add, int j_begin = 0, int j_end = 4) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int r = warp; r < nsel; r += 32) {
    const uint32_t* rp = adjp + static_cast<size_t>(nodes[r]) * d.stride + c0;
    for (int w = lane; w < cw_max; w += 32) rows_s[r * cw_max + w] = w < cw ? rp[w] : 0u;
  }
  __syncthreads();
  for (int j = j_begin; j < j_end; ++j) {  // (a caller may take only some of the four groups of sample rows)
    const int x = (j == 0) ? warp : (j == 1) ? 63 - warp : (j == 2) ? 64 + warp : 127 - warp;
    if (x >= nsel) continue;
    uint32_t rx[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) rx[r] = lane + 32 * r < cw_max ? rows_s[x * cw_max + lane + 32 * r] : 0u;
    for (int wq = 0; wq < 4; ++wq) {
      uint32_t bits = emask[x * 4 + wq];
      while (bits) {
        const int y = 32 * wq + __ffs(bits) - 1;
        bits &= bits - 1;
        const uint32_t* ry = rows_s + y * cw_max + lane;
        uint32_t a[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) a[r] = lane + 32 * r < cw_max ? (rx[r] & ry[32 * r]) : 0u;
        // two carry-save adders: a0+a1+a2 -> (s1, c1), a3+a4+a5 -> (s2, c2); sum = popc(s1)+popc(s2) + 2 popc(c1)+2 popc(c2)
        const uint32_t s1 = a[0] ^ a[1] ^ a[2], c1 = (a[0] & a[1]) | (a[2] & (a[0] ^ a[1]));
        const uint32_t s2 = a[3] ^ a[4] ^ a[5], c2 = (a[3] & a[4]) | (a[5] & (a[3] ^ a[4]));
        // third adder over (s1, s2, 0) and the carries: sum = popc(s1 ^ s2) + 2 (popc(s1 & s2) + popc(c1) + popc(c2))
        const uint32_t lo1 = s1 ^ s2, hi1 = s1 & s2;
        // carries c1, c2, hi1 all weigh 2: one more adder -> (s3, c3) with weights 2 and 4
        const uint32_t s3 = c1 ^ c2 ^ hi1, c3 = (c1 & c2) | (hi1 & (c1 ^ c2));
        int c = __popc(lo1) + 2 * __popc(s3) + 4 * __popc(c3);
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) add(x, y, c);
      }
    }
  }
  __syncthreads();
}

// Step 5: theta = K_e-th largest sample count (0 if the sample holds fewer than K_e edges); ts entries are T + 1,
// 0 = no edge.  Two histogram passes (high byte, then the low byte inside the bucket that holds the K_e-th value)
// instead of a 16-step binary search over the value range.  bit 31: certified by a fat sample, the triangle kernel
// need not try to raise it.
__device__ __forceinline__ uint32_t theta_pick(const unsigned short* ts, int nts, int Ke, int* s_cnt) {
  __shared__ int s_h[256];
  __shared__ int s_hi, s_need;
  const int t = threadIdx.x;
  uint32_t th = 0;
  if (nts >= Ke) {  // block-uniform
    // which of the 256 bins holds the need-th largest entry: the highest bin b with at least `need` entries in bins
    // >= b (bin 0 if there is none).  One warp, lane l owns bins 8 l .. 8 l + 7, suffix sums from the top.
    auto pick_bucket = [&](int need) {
      __syncthreads();
      if (t < 32) {
        int mine[8], tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          mine[k] = s_h[8 * t + k];
          tot += mine[k];
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_down_sync(0xffffffffu, tot, o);
          if (t + o < 32) tot += u;
        }
        int above = tot;  // entries in higher lanes' bins
#pragma unroll
        for (int k = 0; k < 8; ++k) above -= mine[k];
        const bool here = above < need && tot >= need;  // true in exactly one lane if the total reaches need
        if (here) {
          int cum = above;
          for (int k = 7; k >= 0; --k) {
            if (cum + mine[k] >= need) {
              s_hi = 8 * t + k;
              s_need = need - cum;  // still needed inside that bin
              break;
            }
            cum += mine[k];
          }
        }
        if (!__any_sync(0xffffffffu, here) && t == 0) {  // fewer than `need` entries in all: the walk used to end at bin 0
          s_hi = 0;
          s_need = need - (tot - mine[0]);
        }
      }
      __syncthreads();
    };
    if (t < 256) s_h[t] = 0;
    __syncthreads();
    for (int k = t; k < kThetaSamples; k += 1024) {
      const unsigned int v = ts[k];
      if (v) atomicAdd(&s_h[v >> 8], 1);
    }
    pick_bucket(Ke);
    const int hi = s_hi, need = s_need;
    __syncthreads();
    if (t < 256) s_h[t] = 0;
    __syncthreads();
    for (int k = t; k < kThetaSamples; k += 1024) {
      const unsigned int v = ts[k];
      if (v && static_cast<int>(v >> 8) == hi) atomicAdd(&s_h[v & 255u], 1);
    }
    pick_bucket(need);
    // the K_e-th largest entry is T + 1 = 256 hi + lo; theta = T
    th = static_cast<uint32_t>(256 * hi + s_hi) - 1u;
  }
  (void)s_cnt;
  return th | (nts >= 4 * Ke ? 0x80000000u : 0u);
}

__global__ void __launch_bounds__(1024) tri_theta_kernel(const PairDesc* __restrict__ descs,
                                                         const uint32_t* __restrict__ adj,
                                                         const ChunkDev* __restrict__ chunk,
                                                         uint32_t* __restrict__ theta, int Ke, int prune, int max_npad) {
  if (chunk->overflow || !chunk->use_tensor) return;
  const int pair = blockIdx.x;
  if (!prune) {
    if (threadIdx.x == 0) theta[pair] = 0u;
    return;
  }
  const PairDesc d = descs[pair];
  extern __shared__ __align__(16) unsigned char th_smem[];
  const int cw_max = min(max_npad / 32, kThetaChunkW);
  uint32_t* rows_s = reinterpret_cast<uint32_t*>(th_smem);                         // [128][cw]
  uint32_t* emask = rows_s + static_cast<size_t>(kThetaNodes) * cw_max;            // [128][4] edge bits among the sample nodes (y > x)
  unsigned short* ts = reinterpret_cast<unsigned short*>(emask + kThetaNodes * 4); // [128][128] exact T + 1 of sample edge (x, y), 0 = no edge
  unsigned char* deg = reinterpret_cast<unsigned char*>(ts + kThetaSamples);      // [Npad] proxy degrees, saturated
  __shared__ int nodes[kThetaNodes];
  __shared__ int s_cnt, s_nsel, s_nts;
  const int t = threadIdx.x;
  const uint32_t* adjp = adj + d.adj_off;
  int nts = 0;
  const int nsel = theta_sample(d, adjp, deg, nodes, emask, &s_cnt, &s_nsel, &s_nts, &nts);
  // dense table of counts (T + 1, 0 = no edge)
  for (int k = t; k < kThetaSamples; k += 1024) {
    const int x = k >> 7, y = k & 127;
    ts[k] = static_cast<unsigned short>((emask[x * 4 + (y >> 5)] >> (y & 31)) & 1u);
  }
  __syncthreads();
  // 4b. exact T of those edges, the sample rows staged chunk by chunk
  for (int c0 = 0; c0 < d.stride; c0 += cw_max)
    theta_count_chunk(d, adjp, nodes, emask, rows_s, nsel, c0, min(cw_max, d.stride - c0), cw_max,
                      [&](int x, int y, int c) { ts[x * kThetaNodes + y] = static_cast<unsigned short>(ts[x * kThetaNodes + y] + c); });
  const uint32_t th = theta_pick(ts, nts, Ke, &s_cnt);
  if (t == 0) theta[pair] = th;
}

// The same computation for calls with very few (large) pairs, where one CTA per pair leaves the device idle for a
// millisecond at N = 50 000: (a) one CTA per pair samples, (b) one CTA per (64-word chunk of the rows, pair) adds its
// share of the exact counts into a global table, (c) one CTA per pair picks the threshold.  Identical result: the
// counts are integers.
struct ThetaScratch {
  int nodes[kThetaNodes];
  uint32_t emask[kThetaNodes * 4];
  int nsel, nts, pad0, pad1;
  uint32_t ts[kThetaSamples];  // exact T of sample edge (x, y)
};
constexpr int kThetaSplitW = 64;

__global__ void __launch_bounds__(1024) tri_theta_sample_kernel(const PairDesc* __restrict__ descs,
                                                                const uint32_t* __restrict__ adj,
                                                                const ChunkDev* __restrict__ chunk,
                                                                ThetaScratch* __restrict__ scratch, int prune) {
  if (chunk->overflow || !chunk->use_tensor || !prune) return;
  const int pair = blockIdx.x;
  const PairDesc d = descs[pair];
  extern __shared__ __align__(16) unsigned char th_smem[];
  unsigned char* deg = th_smem;  // [Npad]
  __shared__ int nodes[kThetaNodes];
  __shared__ uint32_t emask[kThetaNodes * 4];
  __shared__ int s_cnt, s_nsel, s_nts;
  int nts = 0;
  const int nsel = theta_sample(d, adj + d.adj_off, deg, nodes, emask, &s_cnt, &s_nsel, &s_nts, &nts);
  ThetaScratch* sc = scratch + pair;
  const int t = threadIdx.x;
  if (t < kThetaNodes) sc->nodes[t] = t < nsel ? nodes[t] : 0;
  if (t < kThetaNodes * 4) sc->emask[t] = emask[t];
  if (t == 0) {
    sc->nsel = nsel;
    sc->nts = nts;
  }
  for (int k = t; k < kThetaSamples; k += 1024) sc->ts[k] = 0u;
}

__global__ void __launch_bounds__(1024) tri_theta_count_kernel(const PairDesc* __restrict__ descs,
                                                               const uint32_t* __restrict__ adj,
                                                               const ChunkDev* __restrict__ chunk,
                                                               ThetaScratch* __restrict__ scratch, int prune) {
  if (chunk->overflow || !chunk->use_tensor || !prune) return;
  const int pair = blockIdx.y;
  const PairDesc d = descs[pair];
  const int c0 = static_cast<int>(blockIdx.x) * kThetaSplitW;
  if (c0 >= d.stride) return;
  __shared__ uint32_t rows_s[kThetaNodes * kThetaSplitW];  // 32 KB
  __shared__ int nodes[kThetaNodes];
  __shared__ uint32_t emask[kThetaNodes * 4];
  ThetaScratch* sc = scratch + pair;
  const int t = threadIdx.x;
  if (t < kThetaNodes) nodes[t] = sc->nodes[t];
  if (t < kThetaNodes * 4) emask[t] = sc->emask[t];
  __syncthreads();
  // gridDim.z = 4 (calls with very few pairs): one group of sample rows per CTA, so that a single pair of N = 5000
  // runs on 12 CTAs instead of 3
  const int j0 = gridDim.z == 4 ? static_cast<int>(blockIdx.z) : 0, j1 = gridDim.z == 4 ? j0 + 1 : 4;
  theta_count_chunk(d, adj + d.adj_off, nodes, emask, rows_s, sc->nsel, c0, min(kThetaSplitW, d.stride - c0), kThetaSplitW,
                    [&](int x, int y, int c) { if (c) atomicAdd(&sc->ts[x * kThetaNodes + y], static_cast<uint32_t>(c)); }, j0, j1);
}

__global__ void __launch_bounds__(1024) tri_theta_pick_kernel(const ChunkDev* __restrict__ chunk,
                                                              const ThetaScratch* __restrict__ scratch,
                                                              uint32_t* __restrict__ theta, int Ke, int prune) {
  if (chunk->overflow || !chunk->use_tensor) return;
  const int pair = blockIdx.x;
  if (!prune) {
    if (threadIdx.x == 0) theta[pair] = 0u;
    return;
  }
  __shared__ unsigned short ts[kThetaSamples];
  __shared__ int s_cnt;
  const ThetaScratch* sc = scratch + pair;
  for (int k = threadIdx.x; k < kThetaSamples; k += 1024) {
    const int x = k >> 7, y = k & 127;
    const uint32_t e = (sc->emask[x * 4 + (y >> 5)] >> (y & 31)) & 1u;
    ts[k] = static_cast<unsigned short>(e ? sc->ts[k] + 1u : 0u);
  }
  __syncthreads();
  const uint32_t th = theta_pick(ts, sc->nts, Ke, &s_cnt);
  if (threadIdx.x == 0) theta[pair] = th;
}

// ------------------------------------------------------------------------------------------
// Tensor-pipe peak probe: the triangle kernel's MMA (cta_group::2, kind::mxf4.block_scale, M = 256, N = 240,
// K = 64, the same canonical K-major shared-memory operand layout) issued back to back by one elected thread per
// CTA pair, nothing else running.  bench.py divides the triangle kernel's rate by THIS number, measured in the
// same process on the same GPU (MEASURED_PEAKS.json carries no 4-bit figure).  Operands are zeros: the MMA rate
// does not depend on the data.
// ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_peak_kernel(int stage_pairs) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];  // two operand stages
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int k = tid; k < 2 * kStageBytes / 16; k += 128) reinterpret_cast<uint4*>(smem_raw)[k] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  {
    const uint32_t taddr = tmem + ((32u * warp) << 16) + kSfCol;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
            taddr),
        "r"(kSfWord)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kMmaTileN >> 3) << 17) | (1u << 23) | ((256u >> 4) << 24);
    const uint64_t desc0 = umma_desc(smem_u32(smem_raw));
    const uint32_t desc_lo = static_cast<uint32_t>(desc0), desc_hi = static_cast<uint32_t>(desc0 >> 32);
    uint32_t leader;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(leader));
    for (int ip = 0; ip < stage_pairs; ++ip) {
#pragma unroll
      for (int ks = 0; ks < 2 * kStageK / 64; ++ks) {
        const uint32_t loA = desc_lo + static_cast<uint32_t>(((ks >> 2) * kStageBytes + (ks & 3) * 2 * kLBO) >> 4);
        const uint32_t loB = loA + static_cast<uint32_t>(((kCtaM / 8) * kSBO) >> 4);
        asm volatile(
            "{\n.reg .b64 da, db;\n.reg .pred p, q;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
            "setp.ne.b32 q, %5, 0;\nsetp.eq.u32 p, 1, 1;\n"
            "@q tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], da, db, %4, [%6], [%7], p;\n}" ::"r"(
                tmem + static_cast<uint32_t>(kMmaTileN * (ip & 1))),
            "r"(loA), "r"(loB), "r"(desc_hi), "r"(idesc), "r"(leader), "r"(tmem + kSfCol), "r"(tmem + kSfCol + 16u)
            : "memory");
      }
      __syncwarp();
    }
    umma_commit_pair_if(&done_bar, leader);
  }
  if (warp == 0) mbar_wait_wd(&done_bar, 0u, 9, 0u);
  asm volatile("tcgen05.fence::before_thread_sync;");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// flops of one launch: clusters x stage_pairs x 8 MMAs x (256 x 240 x 64 MAC) x 2
double mma_peak_probe_flops(int clusters, int stage_pairs) {
  return static_cast<double>(clusters) * stage_pairs * 8.0 * 256.0 * kMmaTileN * 64.0 * 2.0;
}
int launch_mma_peak_probe(const LaunchCtx& lc, int clusters, int stage_pairs) {
  mma_peak_kernel<<<2 * clusters, 128, 2 * kStageBytes, lc.stream>>>(stage_pairs);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

int triangles_mma_configure() {
  cudaError_t e = cudaFuncSetAttribute(triangles_mma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(triangles_mma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(triangles_mma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kStageBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tri_theta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(theta_smem_bytes(65536)));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tri_theta_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

size_t theta_scratch_bytes(int pairs) { return pairs <= kThetaSplitMaxPairs ? sizeof(ThetaScratch) * static_cast<size_t>(pairs) : 0; }

int launch_tri_theta(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_npad, const uint32_t* d_adj,
                     const ChunkDev* d_chunk, uint32_t* d_theta, void* d_scratch, int Ke, int prune) {
  // measured (us per step): 1 pair N=50000 1102 -> 379; 43-pair chunks N=10000 992 -> 778; 32 pairs N=5000 138 -> 128;
  // but 86-pair chunks N=5000 290 -> 423: many pairs of moderate size are better off with one CTA each
  const bool split = d_scratch && pairs <= kThetaSplitMaxPairs && max_npad / 32 > 2 * kThetaSplitW &&
                     (pairs <= 32 || max_npad > 8192);
  if (split) {
    ThetaScratch* sc = static_cast<ThetaScratch*>(d_scratch);
    tri_theta_sample_kernel<<<pairs, 1024, static_cast<size_t>(max_npad), lc.stream>>>(d_desc, d_adj, d_chunk, sc, prune);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -static_cast<int>(e);
    const int nchunks = (max_npad / 32 + kThetaSplitW - 1) / kThetaSplitW;
    const int groups = nchunks * pairs <= lc.sm_count ? 4 : 1;  // few CTAs: one group of sample rows each
    tri_theta_count_kernel<<<dim3(nchunks, pairs, groups), 1024, 0, lc.stream>>>(d_desc, d_adj, d_chunk, sc, prune);
    e = cudaGetLastError();
    if (e != cudaSuccess) return -static_cast<int>(e);
    tri_theta_pick_kernel<<<pairs, 1024, 0, lc.stream>>>(d_chunk, sc, d_theta, Ke, prune);
    e = cudaGetLastError();
    return e == cudaSuccess ? 3 : -static_cast<int>(e);
  }
  tri_theta_kernel<<<pairs, 1024, theta_smem_bytes(max_npad), lc.stream>>>(d_desc, d_adj, d_chunk, d_theta, Ke, prune, max_npad);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

int launch_triangles_mma(const LaunchCtx& lc, const PairDesc* d_desc, const uint2* d_tiles, int total_tiles,
                         const int* d_total, const uint2* d_tiles2, const uint32_t* d_adj, const uint32_t* d_panel, PairDev* d_state,
                         const ChunkDev* d_chunk, unsigned long long* d_keys, uint32_t* d_theta, uint32_t* d_hist,
                         unsigned long long* d_t2, int Ke, int raise, int dbg) {
  const int grid = 2 * mma_clusters(total_tiles, lc.sm_count);  // CTA pairs
  if (grid > 0) {
    if (dbg & 64)
      triangles_mma_kernel<true, false><<<grid, kThreads, kSmemBytes, lc.stream>>>(
          d_desc, d_tiles, total_tiles, d_total, d_tiles2, d_adj, d_panel, d_state, d_chunk, d_keys, d_theta, d_hist, d_t2, Ke,
          raise, dbg, nullptr, nullptr, nullptr, 0);
    else
      triangles_mma_kernel<false, false><<<grid, kThreads, kSmemBytes, lc.stream>>>(
          d_desc, d_tiles, total_tiles, d_total, d_tiles2, d_adj, d_panel, d_state, d_chunk, d_keys, d_theta, d_hist, d_t2, Ke,
          raise, dbg, nullptr, nullptr, nullptr, 0);
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

// tiles of node-pruned pairs (kernels_prune.cu): the list and its length live on the device; max_tiles bounds it
int launch_triangles_mma_rect(const LaunchCtx& lc, const PairDesc* d_desc, const uint2* d_rect_tiles, int max_tiles,
                              const int* d_rect_total, const uint32_t* d_adj, const uint32_t* d_panel, PairDev* d_state,
                              const ChunkDev* d_chunk, unsigned long long* d_keys, uint32_t* d_theta, uint32_t* d_hist,
                              unsigned long long* d_t2, int Ke, int raise, int dbg, const NodePlan* d_plan,
                              const unsigned short* d_kept, const uint32_t* d_kpanel, long long kpanel_pair_words) {
  const int grid = 2 * mma_clusters(max_tiles, lc.sm_count);
  if (grid <= 0) return 0;
  triangles_mma_kernel<false, true><<<grid, kThreads, kSmemBytes, lc.stream>>>(
      d_desc, d_rect_tiles, max_tiles, d_rect_total, d_rect_tiles, d_adj, d_panel, d_state, d_chunk, d_keys, d_theta, d_hist,
      d_t2, Ke, raise, dbg, d_plan, d_kept, d_kpanel, kpanel_pair_words);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

// kernels_triangles.cu — S2, compatibility-triangle counts on the POPC bitset path
// (SURVEY.md §8a row S2):  T_ij = popc(row_i & row_j) for every edge i<j, and the per-node sums
// t2_i = sum_j A_ij T_ij (= 2 t_i).  Integer, exact, order free.
//
// Main kernel (rows of up to 352 words, N <= 11264): one CTA per J-block of JB = 128 or 256
// adjacency rows, staged once in shared memory by bulk copies (TMA, one per row, completion on
// an mbarrier).  Warps take rows i < J0+JB-1 from a shared cursor; the lanes hold row i's words
// in registers, the bits A[i][J-block] (j > i) say which staged rows to intersect, and every
// intersection is (vector LDS + AND) per word, a carry-save-adder popcount (5 words -> 3 POPC on
// the 4-lane/clk XU pipe) and one warp REDUX.  The next row (index, edge bits, words, key
// offsets) is prefetched while the current one is processed.  Key emission, histogram and
// node-sum updates are batched 32 edges at a time, one edge per lane.
//
// Keys go to exact positions: the graph kernel counted the edges of every work unit
// (128 columns x 256 rows), the unit scan turned the counts into offsets, and a shared-memory
// cursor per unit hands out sub-ranges to row visits — no global reservation atomics.
// Per edge the kernel emits  T<<32 | (0xFFFF-i)<<16 | (0xFFFF-j)  and updates a 4096-bin
// shared histogram of T>>4 (input of the top-K_e selection), flushed once per CTA.
//
// Chunked kernel (longer rows): one CTA per work unit, rows processed in 256-word chunks with
// per-edge u16 accumulators in shared memory, then an emission pass.
#include "common.cuh"

namespace saccot {

// ---- carry-save-adder popcount -------------------------------------------------------------
__device__ __forceinline__ void csa(uint32_t& ones, uint32_t& carry, uint32_t a, uint32_t b, uint32_t c) {
  ones = a ^ b ^ c;
  carry = (a & b) | (a & c) | (b & c);
}
template <int N>
struct PopcSum {
  static __device__ __forceinline__ int run(const uint32_t (&x)[N]) {
    constexpr int M = N / 2;  // carries: pairs (x[1],x[2]), (x[3],x[4]), ... plus a half adder if N is even
    uint32_t ones = x[0];
    uint32_t carries[M];
#pragma unroll
    for (int k = 0; k < (N - 1) / 2; ++k) {
      uint32_t o, c;
      csa(o, c, ones, x[2 * k + 1], x[2 * k + 2]);
      ones = o;
      carries[k] = c;
    }
    if ((N - 1) % 2 == 1) {  // one word left: half adder
      carries[M - 1] = ones & x[N - 1];
      ones ^= x[N - 1];
    }
    return __popc(ones) + 2 * PopcSum<M>::run(carries);
  }
};
template <>
struct PopcSum<1> {
  static __device__ __forceinline__ int run(const uint32_t (&x)[1]) { return __popc(x[0]); }
};
template <>
struct PopcSum<2> {
  static __device__ __forceinline__ int run(const uint32_t (&x)[2]) { return __popc(x[0]) + __popc(x[1]); }
};

// Row words of one lane.  Word -> lane mapping inside a PITCH = 32 R word row: the first
// A = R/4 rounds are 128-bit (lane holds words 128a+4*lane .. +3), the last B = R%4 rounds are
// 32-bit (word 128A + 32b + lane): R = 5 costs one LDS.128 + one LDS.32 per intersection.
template <int R>
struct LaneWords {
  static constexpr int A = R / 4, B = R % 4;
  uint32_t w[R];
  __device__ __forceinline__ void load_global(const uint32_t* __restrict__ row, int lane, int wc) {
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const int idx = 128 * a + 4 * lane;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (idx < wc) v = *reinterpret_cast<const uint4*>(row + idx);
      w[4 * a + 0] = v.x; w[4 * a + 1] = v.y; w[4 * a + 2] = v.z; w[4 * a + 3] = v.w;
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int idx = 128 * A + 32 * b + lane;
      w[4 * A + b] = idx < wc ? row[idx] : 0u;
    }
  }
  // same as and_popc, addressing the staged row by 32-bit shared-window addresses: `a128` =
  // row address + 16*lane (128-bit rounds), `a32` = row address + 4*(128*A + lane) (32-bit rounds)
  __device__ __forceinline__ int and_popc_saddr(uint32_t a128, uint32_t a32) const {
    uint32_t x[R];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      uint32_t v0, v1, v2, v3;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a128 + 512u * a));
      x[4 * a + 0] = w[4 * a + 0] & v0; x[4 * a + 1] = w[4 * a + 1] & v1;
      x[4 * a + 2] = w[4 * a + 2] & v2; x[4 * a + 3] = w[4 * a + 3] & v3;
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      uint32_t v;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a32 + 128u * b));
      x[4 * A + b] = w[4 * A + b] & v;
    }
    return PopcSum<R>::run(x);
  }
  __device__ __forceinline__ int and_popc(const uint32_t* __restrict__ srow, int lane) const {
    uint32_t x[R];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const uint4 v = *reinterpret_cast<const uint4*>(srow + 128 * a + 4 * lane);
      x[4 * a + 0] = w[4 * a + 0] & v.x; x[4 * a + 1] = w[4 * a + 1] & v.y;
      x[4 * a + 2] = w[4 * a + 2] & v.z; x[4 * a + 3] = w[4 * a + 3] & v.w;
    }
#pragma unroll
    for (int b = 0; b < B; ++b) x[4 * A + b] = w[4 * A + b] & srow[128 * A + 32 * b + lane];
    return PopcSum<R>::run(x);
  }
};

// Edge bits of row i inside a block of NB x 128 columns starting at column J0, restricted to
// j > i (sharded mode: the callers skip the units this rank does not own).
template <int NB>
struct EdgeBits {
  uint32_t w[4 * NB];
  // nsub = number of 128-column sub-blocks that exist in the row (the last block of a pair whose
  // Npad is not a multiple of 128*NB has fewer)
  __device__ __forceinline__ void load(const uint32_t* __restrict__ row_words /* at word J0/32 */, int nsub) {
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (s < nsub) v = *reinterpret_cast<const uint4*>(row_words + 4 * s);
      w[4 * s + 0] = v.x; w[4 * s + 1] = v.y; w[4 * s + 2] = v.z; w[4 * s + 3] = v.w;
    }
  }
  __device__ __forceinline__ void restrict_to(int i, int J0) {
    const int li = i - J0;  // clear bits <= li
#pragma unroll
    for (int k = 0; k < 4 * NB; ++k) {
      if (li >= 32 * k + 31) w[k] = 0u;
      else if (li >= 32 * k) w[k] &= 0xffffffffu << ((li & 31) + 1);
    }
  }
  __device__ __forceinline__ int count(int s) const {
    return __popc(w[4 * s]) + __popc(w[4 * s + 1]) + __popc(w[4 * s + 2]) + __popc(w[4 * s + 3]);
  }
};

// ------------------------------------------------------------------------------------------
// Main kernel.
// ------------------------------------------------------------------------------------------
template <int R, int JB, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) triangles_block_kernel(
    const PairDesc* __restrict__ descs, const uint32_t* __restrict__ adj, const PairDev* __restrict__ state,
    const ChunkDev* __restrict__ chunk, unsigned long long* __restrict__ keys, const uint32_t* __restrict__ ubase,
    int unit_pitch, uint32_t* __restrict__ hist, unsigned long long* __restrict__ t2, int rank, int world) {
  if (chunk->overflow || chunk->use_tensor) return;
  const int pair = blockIdx.y;
  const PairDesc d = descs[pair];
  constexpr int NB = JB / 128;
  const int nJ = (d.Npad + JB - 1) / JB;
  if (static_cast<int>(blockIdx.x) >= nJ) return;
  const int jblk = nJ - 1 - static_cast<int>(blockIdx.x);  // largest (most work) first
  const int J0 = jblk * JB;
  const unsigned int jb0 = static_cast<unsigned int>(jblk) * NB;  // first 128-column block
  const int jrows = min(JB, d.Npad - J0);                          // staged rows (128 or 256)

  constexpr int PITCH = 32 * R;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw);                 // [JB][PITCH]
  uint32_t* hist_s = rows + JB * PITCH;                                   // [4096]
  uint32_t* tJ = hist_s + kHistBins;                                      // [JB]
  uint32_t* cur = tJ + JB;                                                // [NB][256] unit cursors
  uint64_t* bar = reinterpret_cast<uint64_t*>(cur + NB * 256);            // mbarrier
  uint16_t* elist_all = reinterpret_cast<uint16_t*>(bar + 2);             // [NWARP][JB]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint16_t* elist = elist_all + warp * JB;
  for (int k = tid; k < kHistBins; k += THREADS) hist_s[k] = 0;
  for (int k = tid; k < JB + NB * 256; k += THREADS) tJ[k] = 0;  // tJ and cur are contiguous
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t* adjp = adj + d.adj_off;
  const int stride = d.stride;
  if (warp == 0) {
    if (lane == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(jrows * stride * 4));
    __syncwarp();
    for (int rr = lane; rr < jrows; rr += 32)
      bulk_g2s(rows + rr * PITCH, adjp + static_cast<size_t>(J0 + rr) * stride, static_cast<uint32_t>(stride * 4), bar);
  }

  unsigned long long* keyp = keys + state[pair].key_base;
  const uint32_t* ubp = ubase + static_cast<size_t>(pair) * unit_pitch;
  const int row_end = min(d.N, J0 + jrows - 1);  // rows i >= row_end have no neighbour j > i in the block

  // ---- row visits: warp w takes rows w, w+NWARP, ... (edge counts per row vary little, so a static
  // interleave balances well and costs no atomics); the next visit's data is prefetched into a
  // second register set while the current one is processed (two copies of the body, no moves) ----
  // Edge bits travel one byte per lane: lane l holds the bits of columns J0+8l .. J0+8l+7
  // (JB = 256 -> 32 bytes, JB = 128 -> lanes 0..15).
  constexpr int NWARP = THREADS / 32;
  const uint32_t unit0 = unit_offset(jb0), unit1 = NB == 2 ? unit_offset(jb0 + 1) : 0u;
  const bool sub1 = NB == 2 && jrows > 128;  // second 128-column sub-block exists
  // per-lane base pointers (row 0) and validity, loop invariant
  const uint8_t* pbyte = reinterpret_cast<const uint8_t*>(adjp + J0 / 32) + lane;
  const bool has_byte = lane < jrows / 8;
  const uint32_t* pw128 = adjp + 4 * lane;
  const uint32_t* pw32 = adjp + 128 * LaneWords<R>::A + lane;
  bool ok128[LaneWords<R>::A > 0 ? LaneWords<R>::A : 1], ok32[LaneWords<R>::B > 0 ? LaneWords<R>::B : 1];
#pragma unroll
  for (int a = 0; a < LaneWords<R>::A; ++a) ok128[a] = 128 * a + 4 * lane < stride;
#pragma unroll
  for (int b = 0; b < LaneWords<R>::B; ++b) ok32[b] = 128 * LaneWords<R>::A + 32 * b + lane < stride;

  struct Visit {
    uint32_t byte;
    LaneWords<R> ri;
    uint32_t ub[NB];
  };
  auto fetch = [&](int i, Visit& v) {
    if (i >= row_end) return;
    const size_t off = static_cast<size_t>(i) * stride;  // words
    v.byte = has_byte ? pbyte[off * 4] : 0u;
#pragma unroll
    for (int a = 0; a < LaneWords<R>::A; ++a) {
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (ok128[a]) q = *reinterpret_cast<const uint4*>(pw128 + off + 128 * a);
      v.ri.w[4 * a + 0] = q.x; v.ri.w[4 * a + 1] = q.y; v.ri.w[4 * a + 2] = q.z; v.ri.w[4 * a + 3] = q.w;
    }
#pragma unroll
    for (int b = 0; b < LaneWords<R>::B; ++b) v.ri.w[4 * LaneWords<R>::A + b] = ok32[b] ? pw32[off + 32 * b] : 0u;
    v.ub[0] = ubp[unit0 + (static_cast<unsigned int>(i) >> 8)];
    if (NB == 2) v.ub[NB - 1] = sub1 ? ubp[unit1 + (static_cast<unsigned int>(i) >> 8)] : 0u;
  };
  const uint32_t sa128 = smem_u32(rows) + 16u * lane;
  const uint32_t sa32 = smem_u32(rows) + 4u * (128u * LaneWords<R>::A + lane);
  const unsigned int jkey0 = 0xFFFFu - static_cast<unsigned int>(J0);  // (0xFFFF - j) = jkey0 - jl

  auto process = [&](int i, const Visit& v) {
    uint32_t bits = v.byte;
    if (i >= J0) {  // diagonal region: keep only columns j > i
      const int li = i - J0;
      if (8 * lane + 7 <= li) bits = 0u;
      else if (8 * lane <= li) bits &= (0xFFu << ((li & 7) + 1)) & 0xFFu;
    }
    if (world > 1) {  // sharded: keep only the sub-blocks whose unit this rank owns
      const uint32_t unit = ((NB == 2 && lane >= 16) ? unit1 : unit0) + (static_cast<unsigned int>(i) >> 8);
      if (owner_of_unit(unit, static_cast<unsigned int>(world)) != static_cast<unsigned int>(rank)) bits = 0u;
    }
    // ordinal of each lane's first edge: inclusive warp scan of the per-byte counts
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    const int n = __shfl_sync(0xffffffffu, incl, 31);
    if (n == 0) return;
    const int n0 = NB == 2 ? __shfl_sync(0xffffffffu, incl, 15) : n;  // edges in the first sub-block

    // sub-range of each unit's key region for this row visit (lane 0 asks the unit cursors)
    uint32_t pos[NB];
    {
      uint32_t off0 = 0, off1 = 0;
      if (lane == 0) {
        if (n0) off0 = atomicAdd(&cur[i >> 8], static_cast<uint32_t>(n0));
        if (NB == 2 && n > n0) off1 = atomicAdd(&cur[256 + (i >> 8)], static_cast<uint32_t>(n - n0));
      }
      pos[0] = v.ub[0] + off0;
      if (NB == 2) pos[NB - 1] = v.ub[NB - 1] + off1;
    }
    // edge columns in ascending order -> elist (one u16 per edge: the local column jl)
    __syncwarp();  // the previous row's elist has been consumed by every lane
    {
      int o = incl - cnt;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        elist[o++] = static_cast<uint16_t>(8 * lane + b);
      }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NB; ++s) pos[s] = __shfl_sync(0xffffffffu, pos[s], 0);

    const unsigned long long ikey = static_cast<unsigned long long>(0xFFFFu - static_cast<unsigned int>(i)) << 16;
    unsigned int tsum = 0;
    // edges in groups of 32: lane k of the group keeps T of the group's k-th edge, then all lanes
    // do the bookkeeping (key store, histogram, J-side node sum) for their edge at once
    for (int q0 = 0; q0 < n; q0 += 32) {
      const int cntq = min(32, n - q0);
      unsigned int myT = 0;
      const uint16_t* el = elist + q0;
#pragma unroll 4
      for (int k = 0; k < cntq; ++k) {
        const uint32_t roff = static_cast<uint32_t>(el[k]) * (PITCH * 4u);
        const int s = v.ri.and_popc_saddr(roff + sa128, roff + sa32);
        const unsigned int T = static_cast<unsigned int>(__reduce_add_sync(0xffffffffu, s));
        if (lane == k) myT = T;
      }
      if (lane < cntq) {
        const int q = q0 + lane;
        const unsigned int myJ = el[lane];
        uint32_t at = pos[0] + static_cast<uint32_t>(q);
        if (NB == 2 && q >= n0) at = pos[NB - 1] + static_cast<uint32_t>(q - n0);
        keyp[at] = (static_cast<unsigned long long>(myT) << 32) | ikey | static_cast<unsigned long long>(jkey0 - myJ);
        atomicAdd(&hist_s[myT >> 4], 1u);
        atomicAdd(&tJ[myJ], myT);
        tsum += myT;
      }
    }
    tsum = __reduce_add_sync(0xffffffffu, tsum);
    if (lane == 0) atomicAdd(&t2[d.node_off + i], static_cast<unsigned long long>(tsum));
  };

  Visit va, vb;
  int i = warp;
  fetch(i, va);
  mbar_wait(bar, 0);
  while (i < row_end) {
    fetch(i + NWARP, vb);
    process(i, va);
    i += NWARP;
    if (i >= row_end) break;
    fetch(i + NWARP, va);
    process(i, vb);
    i += NWARP;
  }
  __syncthreads();

  // flush the block's histogram and J-side node sums
  uint32_t* histp = hist + static_cast<size_t>(pair) * kHistBins;
  for (int k = tid; k < kHistBins; k += THREADS) {
    const uint32_t v = hist_s[k];
    if (v) atomicAdd(&histp[k], v);
  }
  for (int k = tid; k < jrows; k += THREADS) {
    const uint32_t v = tJ[k];
    if (v) atomicAdd(&t2[d.node_off + J0 + k], static_cast<unsigned long long>(v));
  }
}

// ------------------------------------------------------------------------------------------
// Chunked kernel: one CTA per work unit (128 columns x 256 rows), rows in 32*R-word chunks.
// ------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kTriThreads) triangles_chunked_kernel(
    const PairDesc* __restrict__ descs, const uint32_t* __restrict__ adj, const PairDev* __restrict__ state,
    const ChunkDev* __restrict__ chunk, unsigned long long* __restrict__ keys, const uint32_t* __restrict__ ubase,
    int unit_pitch, uint32_t* __restrict__ hist, unsigned long long* __restrict__ t2, int rank, int world) {
  if (chunk->overflow || chunk->use_tensor) return;
  const int pair = blockIdx.y;
  const PairDesc d = descs[pair];
  const unsigned int unit = blockIdx.x;
  if (unit >= unit_count(static_cast<unsigned int>(d.nblk))) return;
  if (world > 1 && owner_of_unit(unit, static_cast<unsigned int>(world)) != static_cast<unsigned int>(rank)) return;
  unsigned int jb = static_cast<unsigned int>(2.0f * sqrtf(static_cast<float>(unit)));
  if (jb >= static_cast<unsigned int>(d.nblk)) jb = d.nblk - 1;
  while (unit_offset(jb) > unit) --jb;
  while (unit_offset(jb + 1) <= unit) ++jb;
  const unsigned int ic = unit - unit_offset(jb);
  const int J0 = static_cast<int>(jb) * kTriJ;
  const int I0 = static_cast<int>(ic) * kTriI;

  constexpr int PITCH = 32 * R;
  constexpr int NWARP = kTriThreads / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw);                     // [128][PITCH]
  uint32_t* hist_s = rows + kTriJ * PITCH;                                    // [4096]
  uint32_t* tJ = hist_s + kHistBins;                                          // [128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(tJ + kTriJ);                    // mbarrier
  int* next_row = reinterpret_cast<int*>(bar + 1);                            // dynamic row cursor
  uint32_t* cursor = reinterpret_cast<uint32_t*>(bar + 1) + 1;                // key cursor of the unit
  uint8_t* elist_all = reinterpret_cast<uint8_t*>(bar + 2);                   // [NWARP][128]
  uint16_t* acc = reinterpret_cast<uint16_t*>(elist_all + NWARP * kTriJ);     // [256][128]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* elist = elist_all + warp * kTriJ;
  for (int k = tid; k < kHistBins; k += kTriThreads) hist_s[k] = 0;
  if (tid < kTriJ) tJ[tid] = 0;
  if (tid == 0) {
    *cursor = 0;
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t* adjp = adj + d.adj_off;
  const int stride = d.stride;
  unsigned long long* keyp = keys + state[pair].key_base + ubase[static_cast<size_t>(pair) * unit_pitch + unit];
  const unsigned int jkey0 = 0xFFFFu - static_cast<unsigned int>(J0);
  const int row_end = min(min(kTriI, d.N - I0), J0 + kTriJ - 1 - I0);
  uint32_t phase = 0;
  const int my_chunks = (stride + PITCH - 1) / PITCH;

  for (int c = 0; c < my_chunks; ++c) {
    const int c0 = c * PITCH;
    const int wc = min(PITCH, stride - c0);  // words of this chunk (multiple of 4)
    if (tid == 0) *next_row = 0;
    if (warp == 0) {
      if (lane == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(kTriJ * wc * 4));
      __syncwarp();
      for (int rr = lane; rr < kTriJ; rr += 32)
        bulk_g2s(rows + rr * PITCH, adjp + static_cast<size_t>(J0 + rr) * stride + c0, static_cast<uint32_t>(wc * 4), bar);
    }
    __syncthreads();  // next_row reset visible
    mbar_wait(bar, phase);
    phase ^= 1u;

    for (;;) {
      int rr = 0;
      if (lane == 0) rr = atomicAdd(next_row, 1);
      rr = __shfl_sync(0xffffffffu, rr, 0);
      if (rr >= row_end) break;
      const int i = I0 + rr;
      EdgeBits<1> eb;
      eb.load(adjp + static_cast<size_t>(i) * stride + jb * 4, 1);
      eb.restrict_to(i, J0);
      const int n = eb.count(0);
      if (n == 0) continue;
      LaneWords<R> ri;
      ri.load_global(adjp + static_cast<size_t>(i) * stride + c0, lane, wc);
      __syncwarp();
      {
        const uint32_t lt = (1u << lane) - 1u;
        int before = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if ((eb.w[k] >> lane) & 1u) elist[before + __popc(eb.w[k] & lt)] = static_cast<uint8_t>(k * 32 + lane);
          before += __popc(eb.w[k]);
        }
      }
      __syncwarp();
      unsigned int myT = 0, myJ = 0;
      auto flush = [&](int cnt) {
        if (lane < cnt) {
          uint16_t* a = acc + rr * kTriJ + myJ;
          *a = static_cast<uint16_t>(c == 0 ? myT : static_cast<unsigned int>(*a) + myT);
        }
      };
#pragma unroll 2
      for (int q = 0; q < n; ++q) {
        const int jl = elist[q];
        const int s = ri.and_popc(rows + jl * PITCH, lane);
        const unsigned int T = static_cast<unsigned int>(__reduce_add_sync(0xffffffffu, s));
        if (lane == (q & 31)) { myT = T; myJ = static_cast<unsigned int>(jl); }
        if ((q & 31) == 31) flush(32);
      }
      if (n & 31) flush(n & 31);
    }
    __syncthreads();  // everyone is done with `rows` (and next_row) before the next chunk
  }

  // emission pass from the accumulated per-edge counts: lanes take the set bits of the row
  if (tid == 0) *next_row = 0;
  __syncthreads();
  for (;;) {
    int rr = 0;
    if (lane == 0) rr = atomicAdd(next_row, 1);
    rr = __shfl_sync(0xffffffffu, rr, 0);
    if (rr >= row_end) break;
    const int i = I0 + rr;
    EdgeBits<1> eb;
    eb.load(adjp + static_cast<size_t>(i) * stride + jb * 4, 1);
    eb.restrict_to(i, J0);
    const int n = eb.count(0);
    if (n == 0) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(cursor, static_cast<uint32_t>(n));
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned long long ikey = static_cast<unsigned long long>(0xFFFFu - static_cast<unsigned int>(i)) << 16;
    const uint32_t lt = (1u << lane) - 1u;
    unsigned int tsum = 0;
    int before = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      if ((eb.w[w] >> lane) & 1u) {
        const unsigned int jl = static_cast<unsigned int>(w * 32 + lane);
        const unsigned int o = static_cast<unsigned int>(before + __popc(eb.w[w] & lt));
        const unsigned int T = acc[rr * kTriJ + jl];
        keyp[base + o] = (static_cast<unsigned long long>(T) << 32) | ikey | static_cast<unsigned long long>(jkey0 - jl);
        atomicAdd(&hist_s[T >> 4], 1u);
        atomicAdd(&tJ[jl], T);
        tsum += T;
      }
      before += __popc(eb.w[w]);
    }
    tsum = __reduce_add_sync(0xffffffffu, tsum);
    if (lane == 0) atomicAdd(&t2[d.node_off + i], static_cast<unsigned long long>(tsum));
  }
  __syncthreads();

  uint32_t* histp = hist + static_cast<size_t>(pair) * kHistBins;
  for (int k = tid; k < kHistBins; k += kTriThreads) {
    const uint32_t v = hist_s[k];
    if (v) atomicAdd(&histp[k], v);
  }
  if (tid < kTriJ) {
    const uint32_t v = tJ[tid];
    if (v) atomicAdd(&t2[d.node_off + J0 + tid], static_cast<unsigned long long>(v));
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static size_t block_smem_bytes(int R, int JB, int threads) {
  return static_cast<size_t>(JB) * 32 * R * 4 + kHistBins * 4 + JB * 4 + (JB / 128) * 256 * 4 + 16 +
         static_cast<size_t>(threads / 32) * JB * 2;
}
static size_t chunked_smem_bytes(int R) {
  return static_cast<size_t>(kTriJ) * 32 * R * 4 + kHistBins * 4 + kTriJ * 4 + 16 + (kTriThreads / 32) * kTriJ +
         static_cast<size_t>(kTriI) * kTriJ * 2;
}

// Configuration table: rows of 32 R words; JB = 256 staged rows while they fit in shared memory
// next to 1024 threads (R <= 5, i.e. N <= 5120), else 128 rows; 512 threads once the per-lane
// row words (2 R registers with the prefetch) no longer fit the 64-register budget of 1024.
#define SACCOT_TRI_CONFIGS(X) \
  X(1, 256, 1024) X(2, 256, 1024) X(3, 256, 1024) X(4, 256, 1024) X(5, 256, 1024) \
  X(6, 128, 512) X(7, 128, 512) X(8, 128, 512) X(9, 128, 512) X(10, 128, 512) X(11, 128, 512)
// sharded runs (world > 1) use 128-row blocks with the same thread counts
#define SACCOT_TRI_CONFIGS_128(X) \
  X(1, 128, 1024) X(2, 128, 1024) X(3, 128, 1024) X(4, 128, 1024) X(5, 128, 1024)

int triangles_configure() {
  cudaError_t e;
#define SACCOT_SET(R, JB, TH)                                                                                   \
  if ((e = cudaFuncSetAttribute(triangles_block_kernel<R, JB, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                static_cast<int>(block_smem_bytes(R, JB, TH)))) != cudaSuccess)                 \
    return -static_cast<int>(e);
  SACCOT_TRI_CONFIGS(SACCOT_SET)
  SACCOT_TRI_CONFIGS_128(SACCOT_SET)
#undef SACCOT_SET
  if ((e = cudaFuncSetAttribute(triangles_chunked_kernel<kTriChunkR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(chunked_smem_bytes(kTriChunkR)))) != cudaSuccess)
    return -static_cast<int>(e);
  return 0;
}

int launch_triangles(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_nblk, int max_stride,
                     const uint32_t* d_adj, PairDev* d_state, const ChunkDev* d_chunk, unsigned long long* d_keys,
                     const uint32_t* d_ubase, int unit_pitch, uint32_t* d_hist, unsigned long long* d_t2, int rank,
                     int world) {
  const int R = (max_stride + 31) / 32;
  if (R > kTriMaxR) {
    dim3 grid(unit_count(static_cast<unsigned int>(max_nblk)), pairs);
    triangles_chunked_kernel<kTriChunkR><<<grid, kTriThreads, chunked_smem_bytes(kTriChunkR), lc.stream>>>(
        d_desc, d_adj, d_state, d_chunk, d_keys, d_ubase, unit_pitch, d_hist, d_t2, rank, world);
  } else {
    bool launched = false;
#define SACCOT_LAUNCH(RR, JB, TH)                                                                 \
  if (!launched && R == RR && (JB == 128 || world <= 1)) {                                       \
    dim3 grid((max_nblk * 128 + JB - 1) / JB, pairs);                                             \
    triangles_block_kernel<RR, JB, TH><<<grid, TH, block_smem_bytes(RR, JB, TH), lc.stream>>>(    \
        d_desc, d_adj, d_state, d_chunk, d_keys, d_ubase, unit_pitch, d_hist, d_t2, rank, world); \
    launched = true;                                                                              \
  }
    SACCOT_TRI_CONFIGS(SACCOT_LAUNCH)
    SACCOT_TRI_CONFIGS_128(SACCOT_LAUNCH)
#undef SACCOT_LAUNCH
    if (!launched) return -static_cast<int>(cudaErrorInvalidValue);
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

// kernels_match.cu — correspondence front end: nearest-neighbour matching of local descriptors
// (33-D FPFH-like), the stage that PRODUCES the N putative correspondences the hot path registers.  The reference
// has no such code either (/root/reference/README.md:1-2); the specification is DESIGN.md §2 "S-1":
//
//     D_ij = sum_c (f_ic - g_jc)^2   in fp32:  D = 0;  for c = 0 .. dim-1:  e = f_ic - g_jc;  D = fma(e, e, D)
//     nn(i) = argmin_j D_ij, ties -> lowest j;   correspondence i = (xyz_src[i], xyz_dst[nn(i)])
//
// The search is a dense contraction (|f|^2 + |g|^2 - 2 f.g), so it runs on the tensor cores; the DECISION is exact:
//   1. match_prep_kernel   descriptors -> bf16 operand images in the UMMA canonical (no-swizzle, K-major) layout, one
//                          contiguous image per tile, so the matching kernel stages a tile with ONE bulk copy.  Every
//                          fp32 value is split into bf16 hi + lo; the K dimension carries [f_hi | f_lo | f_hi] against
//                          [g_hi | g_hi | g_lo] (the three products that matter) plus three columns (-1,-1,-1) against
//                          the three bf16 pieces of |g_j|^2 / 2: the accumulator is C_ij ~ f_i.g_j - |g_j|^2/2, and
//                          maximising C_ij over j is minimising D_ij.
//   2. match_mma_kernel    one CTA per 128 source rows: tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), M 128 x
//                          N 256 per tile, K = 3 dim + 3 padded to 16; a producer warp feeds a 3-stage bulk-copy ring
//                          and issues the MMAs into two TMEM accumulators; four epilogue warps (thread = row = TMEM
//                          lane) read them back with tcgen05.ld and keep, per row, the running maximum and the SHORT
//                          LIST of columns within a margin of it.  |C~ - C| <= 2.6e-5 (|f|^2 + |g|^2) (bf16x3 split
//                          2^-15.5, tensor-core accumulation <= 128 x 2^-22) and the fp32 chain of the specification
//                          is within 4.2e-6 of the real distance, so the true nearest neighbour under the SPECIFIED
//                          arithmetic lies within 2^-12 (|f_i|^2 + max_j |g_j|^2) of the best approximate value: the
//                          list contains it (DESIGN.md §4 "match").
//   3. match_exact_kernel  the specified fp32 chain for the listed columns (typically one or two per row), lowest j on
//                          ties; a row whose list overflowed (many near-identical descriptors) is scanned exhaustively
//                          with the same chain.  Then the gather of the matched points.
// Descriptors wider than kMatchMaxDim skip 1-2: every row is scanned exhaustively (CUDA cores).
#include "common.cuh"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>

namespace saccot {

namespace {

constexpr int kStagesB = 3;                           // most B stages (wide descriptors get 2: shared-memory budget)
constexpr float kHugeNorm = 1.0e38f;                 // |g|^2 / 2 of a pad row: its C is -1e38, never a candidate
constexpr float kMarginRel = 1.0f / 4096.0f;         // 2^-12, in units of C (see the header comment)

__device__ __forceinline__ uint64_t match_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((128u >> 4) & 0x3FFF) << 16;        // LBO: next 16-byte K chunk
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;   // SBO: next 8-row group
  d |= 1ull << 46;                                               // descriptor version (Blackwell); SWIZZLE_NONE
  return d;
}

// bounded barrier wait: a protocol bug must end the kernel, not hang the GPU
__device__ __forceinline__ void match_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (int spins = 0;; ++spins) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
    if (done) return;
    if (spins == 0) t0 = clock64();
    else if (clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ void match_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void match_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#define SACCOT_MATCH_LD32(v, taddr)                                                                                    \
  asm volatile(                                                                                                        \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19," \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                       \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),    \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),          \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),         \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                       \
      : "r"(taddr))

}  // namespace

// ------------------------------------------------------------------------------------------
// 1. operand images.  side 0: source rows, tiles of 128, K sections [hi | lo | hi | -1 -1 -1]; also |f_i|^2.
//                     side 1: target rows, tiles of 256, K sections [hi | hi | lo | pieces of |g_j|^2 / 2]; also the
//                             pair's max |g_j|^2 (atomicMax on the bit pattern: norms are non-negative floats).
//    One thread per (row, 16-byte chunk); image of a tile: [8-row group][chunk][8 rows][16 B].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) match_prep_kernel(const MatchPair* __restrict__ pairs, const float* __restrict__ desc,
                                                         int side, int dim, int chunks, unsigned char* __restrict__ img,
                                                         float* __restrict__ norms, uint32_t* __restrict__ bmax) {
  const MatchPair mp = pairs[blockIdx.y];
  const int TR = side ? kMatchTileN : kMatchTileM;
  const int n = side ? mp.Nd : mp.Ns;
  const int tiles = side ? mp.d_tiles : mp.s_tiles;
  const long long item = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (item >= static_cast<long long>(tiles) * TR * chunks) return;
  const int c = static_cast<int>(item % chunks);
  const int r = static_cast<int>(item / chunks);  // row of the pair, pad rows included
  const float* row = desc + (static_cast<size_t>(side ? mp.d_off : mp.s_off) + static_cast<size_t>(r < n ? r : 0)) * dim;
  // the row's squared norm (every chunk thread of a target row needs its pieces only in the chunk that holds them;
  // computing it in the one thread that stores it keeps the chain's order fixed: c ascending, one fma each)
  const int k_norm = 3 * dim;  // first of the three norm columns
  float nrm = 0.0f;
  const bool need_norm = c == 0 || (8 * c <= k_norm + 2 && 8 * c + 7 >= k_norm);
  if (need_norm && r < n)
    for (int k = 0; k < dim; ++k) nrm = __fmaf_rn(row[k], row[k], nrm);
  float half = side ? (r < n ? __fmul_rn(nrm, 0.5f) : kHugeNorm) : 0.0f;
  // three bf16 pieces of |g|^2 / 2 (hi, mid, lo: 24 mantissa bits)
  const __nv_bfloat16 p0 = __float2bfloat16_rn(half);
  const float r1 = __fsub_rn(half, __bfloat162float(p0));
  const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
  const __nv_bfloat16 p2 = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(p1)));
  __nv_bfloat16 out[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = 8 * c + e;
    __nv_bfloat16 v = __float2bfloat16_rn(0.0f);
    if (k < k_norm) {
      if (r < n) {
        const int sec = k / dim, col = k - sec * dim;
        const float f = row[col];
        const __nv_bfloat16 hi = __float2bfloat16_rn(f);
        const __nv_bfloat16 lo = __float2bfloat16_rn(__fsub_rn(f, __bfloat162float(hi)));
        const bool want_lo = side ? sec == 2 : sec == 1;
        v = want_lo ? lo : hi;
      }
    } else if (k < k_norm + 3) {
      if (side) v = k == k_norm ? p0 : (k == k_norm + 1 ? p1 : p2);
      else v = __float2bfloat16_rn(-1.0f);
    }
    out[e] = v;
  }
  const int tt = r / TR, rt = r - tt * TR;
  unsigned char* dst = img + (side ? mp.d_img : mp.s_img) + static_cast<size_t>(tt) * TR * chunks * 16 +
                       static_cast<size_t>(rt >> 3) * chunks * 128 + static_cast<size_t>(c) * 128 + (rt & 7) * 16;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(out);
  if (c == 0) {
    if (!side) norms[mp.s_norm + r] = r < n ? nrm : 0.0f;
    else if (r < n) atomicMax(bmax + blockIdx.y, __float_as_uint(nrm));
  }
}

// ------------------------------------------------------------------------------------------
// 2. tensor-core sweep.  grid (max source tiles, pairs).
//    Warps 0-15: epilogue.  Warp w reads TMEM lanes 32 (w % 4) .. + 31 (the rows a warp may touch are fixed by its
//    index modulo 4) and the column block w / 4 (64 of the tile's 256 columns): a row is swept by four threads, each
//    with its own running best and short list — a thread's list holds everything within the margin of ITS best, which
//    is at most the row's best, so the union of the four lists holds everything within the margin of the row's best.
//    Per 32 columns a thread takes the maximum (FMNMX3 tree) and compares it with its threshold; only then does it
//    look at the individual values (predicated scan, inline).  With 32 rows per warp SOME lane sets a new record in
//    most steps, so that scan must be cheap: two earlier versions — the scan through a local-memory array in an
//    out-of-line function, and one warp per scheduler — ran at 7-10 k cycles per tile against 950 for the tile's
//    MMAs (in-kernel cycle counters, `match_dbg`).
//    Warp 16: producer + MMA issuer (one lane).
// ------------------------------------------------------------------------------------------
constexpr int kEpiBlocks = 4;                               // column blocks of a tile = epilogue warps per lane quadrant
constexpr int kEpiCols = kMatchTileN / kEpiBlocks;          // 64
constexpr int kMatchEpiThreads = kMatchTileM * kEpiBlocks;  // 512
constexpr int kMatchThreadsAll = kMatchEpiThreads + 32;     // 544

__global__ void __launch_bounds__(kMatchThreadsAll, 1) match_mma_kernel(const MatchPair* __restrict__ pairs,
                                                                        const unsigned char* __restrict__ img,
                                                                        const float* __restrict__ norms,
                                                                        const uint32_t* __restrict__ bmax, int chunks, int S,
                                                                        int32_t* __restrict__ cand, int32_t* __restrict__ cand_cnt,
                                                                        int dbg) {
  const MatchPair mp = pairs[blockIdx.y];
  const int st = blockIdx.x;
  if (st >= mp.s_tiles) return;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int a_bytes = kMatchTileM * chunks * 16, b_bytes = kMatchTileN * chunks * 16;
  unsigned char* sA = smem;
  unsigned char* sB = smem + a_bytes;                                      // S stages
  int32_t* lists = reinterpret_cast<int32_t*>(sB + S * b_bytes);           // [512 threads][kMatchCand]
  int32_t* counts = lists + kMatchEpiThreads * kMatchCand;                 // [512]
  float* bests = reinterpret_cast<float*>(counts + kMatchEpiThreads);      // [512] each epilogue thread's best value
  uint64_t* bars = reinterpret_cast<uint64_t*>(bests + kMatchEpiThreads);
  uint64_t* a_full = bars;                 // A image landed
  uint64_t* full = bars + 1;               // [3] B stage landed
  uint64_t* stage_free = full + kStagesB;  // [3] MMAs that read the stage have completed
  uint64_t* mma_done = stage_free + kStagesB;  // [2] accumulator complete
  uint64_t* acc_free = mma_done + 2;       // [2] epilogue has drained the accumulator (16 warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kProdWarp = kMatchEpiThreads / 32;
  if (tid == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < kStagesB; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&stage_free[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&mma_done[b], 1);
      mbar_init(&acc_free[b], kProdWarp);
    }
    mbar_fence_init();
  }
  if (warp == kProdWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;
  const int T = mp.d_tiles;
  const unsigned char* gA = img + mp.s_img + static_cast<size_t>(st) * a_bytes;
  const unsigned char* gB = img + mp.d_img;

  if (warp == kProdWarp) {
    // ================================ producer + MMA issuer (one lane) ================================
    if (lane == 0) {
      // instruction descriptor (kind::f16): D = F32, A = B = BF16, both K-major, N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kMatchTileN >> 3) << 17) |
                             (static_cast<uint32_t>(kMatchTileM >> 4) << 24);
      const uint32_t sbo = static_cast<uint32_t>(chunks) * 128u;
      mbar_arrive_expect_tx(a_full, static_cast<uint32_t>(a_bytes));
      bulk_g2s(sA, gA, static_cast<uint32_t>(a_bytes), a_full);
      for (int n = 0; n < 2 && n < T; ++n) {
        mbar_arrive_expect_tx(&full[n], static_cast<uint32_t>(b_bytes));
        bulk_g2s(sB + n * b_bytes, gB + static_cast<size_t>(n) * b_bytes, static_cast<uint32_t>(b_bytes), &full[n]);
      }
      long long w_a = 0, w_full = 0, w_acc = 0, w_stage = 0;  // dbg: cycles waited on each barrier
      const long long t_begin = clock64();
      match_wait(a_full, 0u);
      w_a = clock64() - t_begin;
      const uint64_t dA0 = match_desc(smem_u32(sA), sbo);
      for (int n = 0; n < T; ++n) {
        const int s = n % S, b = n & 1;
        long long tq = clock64();
        match_wait(&full[s], static_cast<uint32_t>((n / S) & 1));
        w_full += clock64() - tq;
        tq = clock64();
        if (n >= 2) match_wait(&acc_free[b], static_cast<uint32_t>(((n >> 1) - 1) & 1));
        w_acc += clock64() - tq;
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint64_t dB0 = match_desc(smem_u32(sB + s * b_bytes), sbo);
        for (int k = 0; k < chunks / 2; ++k) {  // K = 16 per instruction = two 16-byte chunks
          const uint64_t da = dA0 + static_cast<uint64_t>((k * 256) >> 4), db = dB0 + static_cast<uint64_t>((k * 256) >> 4);
          const uint32_t acc = k ? 1u : 0u;      // the first instruction of a tile overwrites the accumulator
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem + static_cast<uint32_t>(kMatchTileN * b)),
              "l"(da), "l"(db), "r"(idesc), "r"(acc)
              : "memory");
        }
        match_commit(&mma_done[b]);
        match_commit(&stage_free[s]);
        // tile n + 2 goes into the stage tile n + 2 - S used, as soon as the MMAs that read it have completed
        if (n + 2 < T) {
          const int s2 = (n + 2) % S, prev = n + 2 - S;
          const long long tq2 = clock64();
          if (prev >= 0) match_wait(&stage_free[s2], static_cast<uint32_t>((prev / S) & 1));
          w_stage += clock64() - tq2;
          mbar_arrive_expect_tx(&full[s2], static_cast<uint32_t>(b_bytes));
          bulk_g2s(sB + s2 * b_bytes, gB + static_cast<size_t>(n + 2) * b_bytes, static_cast<uint32_t>(b_bytes), &full[s2]);
        }
      }
      if (dbg && blockIdx.x == 0 && blockIdx.y == 0)
        printf("match producer: tiles %d total %lld wait_A %lld wait_full %lld wait_acc_free %lld wait_stage_free %lld\n", T,
               clock64() - t_begin, w_a, w_full, w_acc, w_stage);
    }
  } else {
    // ================================ epilogue ================================
    const int q = warp & 3, eb = warp >> 2;          // lane quadrant, column block
    const int trow = 32 * q + lane;                  // row of the tile = TMEM lane
    const int row = st * kMatchTileM + trow;         // row of the pair (pad rows of the last tile run along, results unused)
    const float a_i = norms[mp.s_norm + row];
    int32_t* lst = lists + (trow * kEpiBlocks + eb) * kMatchCand;
    const float margin = __fmul_rn(kMarginRel, __fadd_rn(a_i, __uint_as_float(bmax[blockIdx.y])));
    float best = -3.0e38f, thr = -3.0e38f;  // thr = best - margin
    int cnt = 0;
    long long w_mma = 0, w_ld = 0;
    const long long t_begin = clock64();
    for (int n = 0; n < T; ++n) {
      const int b = n & 1;
      const long long tq = clock64();
      match_wait(&mma_done[b], static_cast<uint32_t>((n >> 1) & 1));
      w_mma += clock64() - tq;
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t tbase = tmem + ((32u * q) << 16) + static_cast<uint32_t>(kMatchTileN * b + kEpiCols * eb);
      uint32_t v[2][32];
      SACCOT_MATCH_LD32(v[0], tbase);
#pragma unroll
      for (int u = 0; u < kEpiCols / 32; ++u) {
        const long long tl = dbg ? clock64() : 0;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (dbg) w_ld += clock64() - tl;
        if (u + 1 < kEpiCols / 32) SACCOT_MATCH_LD32(v[(u + 1) & 1], tbase + 32u * (u + 1));
        const uint32_t* vv = v[u & 1];
        float m0 = __uint_as_float(vv[0]), m1 = __uint_as_float(vv[1]), m2 = __uint_as_float(vv[2]), m3 = __uint_as_float(vv[3]);
#pragma unroll
        for (int e = 4; e < 32; e += 8) {
          m0 = fmaxf(fmaxf(m0, __uint_as_float(vv[e])), __uint_as_float(vv[e + 4]));
          m1 = fmaxf(fmaxf(m1, __uint_as_float(vv[e + 1])), __uint_as_float(vv[e + 5]));
          m2 = fmaxf(fmaxf(m2, __uint_as_float(vv[e + 2])), __uint_as_float(vv[e + 6]));
          m3 = fmaxf(fmaxf(m3, __uint_as_float(vv[e + 3])), __uint_as_float(vv[e + 7]));
        }
        if (fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) >= thr) {
          // some lane of the warp gets here in most steps (32 rows, each setting ~8 records per sweep), so only the
          // group(s) of eight values (e = k mod 4) whose maximum reaches the threshold are looked at
          const int col0 = n * kMatchTileN + kEpiCols * eb + 32 * u;
          const float mk[4] = {m0, m1, m2, m3};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (mk[k] >= thr) {
#pragma unroll
              for (int e = k; e < 32; e += 4) {
                const float c = __uint_as_float(vv[e]);
                if (c >= thr) {
                  if (c > best) {
                    if (c - margin > best) cnt = 0;  // everything listed so far is now out of range
                    best = c;
                    thr = c - margin;
                  }
                  if (cnt < kMatchCand) lst[cnt] = col0 + e;
                  ++cnt;
                }
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) match_arrive(&acc_free[b]);
    }
    counts[trow * kEpiBlocks + eb] = cnt;
    bests[trow * kEpiBlocks + eb] = best;
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)
      printf("match epilogue: total %lld wait_mma_done %lld wait_tmem_ld %lld\n", clock64() - t_begin, w_mma, w_ld);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == kProdWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  // union of the row's four lists (an overflowing list, or more than kMatchUnion in total, sends the row to the
  // exhaustive scan)
  if (tid < kMatchTileM) {
    const int row = st * kMatchTileM + tid;
    if (row < mp.Ns) {
      const size_t o = static_cast<size_t>(mp.cand_off) + row;
      int total = 0;
      bool over = false;
      // a list whose best lies more than the margin below the row's best holds nothing within the margin of it
      float rb = bests[tid * kEpiBlocks];
      for (int eb = 1; eb < kEpiBlocks; ++eb) rb = fmaxf(rb, bests[tid * kEpiBlocks + eb]);
      const float cut = rb - __fmul_rn(kMarginRel, __fadd_rn(norms[mp.s_norm + row], __uint_as_float(bmax[blockIdx.y])));
      for (int eb = 0; eb < kEpiBlocks; ++eb) {
        if (bests[tid * kEpiBlocks + eb] < cut) continue;
        const int c = counts[tid * kEpiBlocks + eb];
        over |= c > kMatchCand;
        for (int k = 0; k < c && k < kMatchCand; ++k) {
          if (total < kMatchUnion) cand[o * kMatchUnion + total] = lists[(tid * kEpiBlocks + eb) * kMatchCand + k];
          ++total;
        }
      }
      for (int k = total; k < kMatchUnion; ++k) cand[o * kMatchUnion + k] = -1;
      cand_cnt[o] = over ? kMatchUnion + 1 : total;
    }
  }
}

// ------------------------------------------------------------------------------------------
// 3. exact decision + gather.
//    match_exact_kernel: one warp per source row; lane k evaluates the k-th listed candidate with the specified
//    chain (D = fma(e, e, D), c ascending — sequential by specification, so one lane per target row), reading the
//    target row straight from global memory eight independent loads at a time.  Rows whose list overflowed are
//    appended to a work list.
//    match_scan_kernel: the exhaustive scan, one CTA per row of the work list (or per row of the batch when there are
//    no lists at all: descriptors too wide for the sweep, or match_path = 0).  Eight warps split the target rows;
//    32 consecutive rows are one contiguous run, fetched coalesced into a shared-memory stage 64 columns at a time.
// ------------------------------------------------------------------------------------------
constexpr int kExactWarps = 8;
constexpr int kExactCols = 64;                       // columns staged at a time (scan kernel)
constexpr int kExactPitch = kExactCols + 1;          // odd: lanes reading different rows hit different banks
constexpr int kExactStageFloats = 32 * kExactPitch;  // 32 target rows
static size_t match_scan_smem(int dim) { return (static_cast<size_t>(kExactWarps) * kExactStageFloats + dim) * sizeof(float); }

__device__ __forceinline__ void match_write(const MatchPair& mp, int i, int bestJ, const float* __restrict__ xyz_src,
                                            const float* __restrict__ xyz_dst, int32_t* __restrict__ nn,
                                            float* __restrict__ corr_src, float* __restrict__ corr_dst, int lane) {
  if (lane == 0) nn[mp.s_off + i] = bestJ;
  if (lane < 3) {
    corr_src[(static_cast<size_t>(mp.s_off) + i) * 3 + lane] = xyz_src[(static_cast<size_t>(mp.s_off) + i) * 3 + lane];
    corr_dst[(static_cast<size_t>(mp.s_off) + i) * 3 + lane] = xyz_dst[(static_cast<size_t>(mp.d_off) + bestJ) * 3 + lane];
  }
}
// lexicographic minimum of (D, j) over the warp (0x7fffffff = no entry).  A NaN distance never wins against a number.
__device__ __forceinline__ void match_warp_min(float& bestD, int& bestJ) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oD = __shfl_xor_sync(0xffffffffu, bestD, o);
    const int oJ = __shfl_xor_sync(0xffffffffu, bestJ, o);
    const bool take = oJ != 0x7fffffff && (bestJ == 0x7fffffff || oD < bestD || (oD == bestD && oJ < bestJ) ||
                                           (bestD != bestD && oD == oD));
    if (take) {
      bestD = oD;
      bestJ = oJ;
    }
  }
}

__global__ void __launch_bounds__(kExactWarps * 32) match_exact_kernel(const MatchPair* __restrict__ pairs,
                                                                       const float* __restrict__ desc_src,
                                                                       const float* __restrict__ desc_dst,
                                                                       const float* __restrict__ xyz_src,
                                                                       const float* __restrict__ xyz_dst, int dim,
                                                                       const int32_t* __restrict__ cand,
                                                                       const int32_t* __restrict__ cand_cnt,
                                                                       int32_t* __restrict__ nn, float* __restrict__ corr_src,
                                                                       float* __restrict__ corr_dst,
                                                                       int32_t* __restrict__ work_list,
                                                                       uint32_t* __restrict__ work_count) {
  const MatchPair mp = pairs[blockIdx.y];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int i = blockIdx.x * kExactWarps + wib;
  extern __shared__ float ex_smem[];  // [kExactWarps][dim]
  if (i >= mp.Ns) return;
  const int cnt = cand_cnt[mp.cand_off + i];
  if (cnt > kMatchUnion) {  // overflowed list: the scan kernel takes the row
    if (lane == 0) {
      const uint32_t k = atomicAdd(work_count, 1u);
      work_list[2 * k] = static_cast<int32_t>(blockIdx.y);
      work_list[2 * k + 1] = i;
    }
    return;
  }
  float* f = ex_smem + wib * dim;
  const float* fi = desc_src + (static_cast<size_t>(mp.s_off) + i) * dim;
  for (int c = lane; c < dim; c += 32) f[c] = fi[c];
  __syncwarp();
  float bestD = __int_as_float(0x7f800000);  // +inf
  int bestJ = 0x7fffffff;
  if (lane < cnt) {
    const int j = cand[(static_cast<size_t>(mp.cand_off) + i) * kMatchUnion + lane];
    if (j >= 0 && j < mp.Nd) {
      const float* g = desc_dst + (static_cast<size_t>(mp.d_off) + j) * dim;
      float D = 0.0f;
      int c0 = 0;
      for (; c0 + 8 <= dim; c0 += 8) {
        float gv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[k] = g[c0 + k];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float e = __fsub_rn(f[c0 + k], gv[k]);
          D = __fmaf_rn(e, e, D);
        }
      }
      for (; c0 < dim; ++c0) {
        const float e = __fsub_rn(f[c0], g[c0]);
        D = __fmaf_rn(e, e, D);
      }
      bestD = D;
      bestJ = j;
    }
  }
  match_warp_min(bestD, bestJ);
  if (bestJ == 0x7fffffff) bestJ = 0;
  match_write(mp, i, bestJ, xyz_src, xyz_dst, nn, corr_src, corr_dst, lane);
}

__global__ void __launch_bounds__(kExactWarps * 32) match_scan_kernel(const MatchPair* __restrict__ pairs,
                                                                      const float* __restrict__ desc_src,
                                                                      const float* __restrict__ desc_dst,
                                                                      const float* __restrict__ xyz_src,
                                                                      const float* __restrict__ xyz_dst, int dim,
                                                                      const int32_t* __restrict__ work_list,
                                                                      const uint32_t* __restrict__ work_count,
                                                                      int32_t* __restrict__ nn, float* __restrict__ corr_src,
                                                                      float* __restrict__ corr_dst) {
  extern __shared__ float sc_smem[];  // [kExactWarps][32][kExactPitch] stages, then f[dim]
  __shared__ float s_D[kExactWarps];
  __shared__ int s_J[kExactWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* stage = sc_smem + wib * kExactStageFloats;
  float* f = sc_smem + kExactWarps * kExactStageFloats;
  const uint32_t items = work_list ? *work_count : gridDim.x * gridDim.y;
  for (uint32_t it = work_list ? blockIdx.x : blockIdx.y * gridDim.x + blockIdx.x; it < items;
       it += work_list ? gridDim.x : items) {
    const int pair = work_list ? work_list[2 * it] : static_cast<int>(blockIdx.y);
    const int i = work_list ? work_list[2 * it + 1] : static_cast<int>(blockIdx.x);
    const MatchPair mp = pairs[pair];
    if (i >= mp.Ns) break;  // (only without a work list: the grid spans the largest pair)
    __syncthreads();        // the previous row's f and reduction slots have been consumed
    const float* fi = desc_src + (static_cast<size_t>(mp.s_off) + i) * dim;
    for (int c = threadIdx.x; c < dim; c += kExactWarps * 32) f[c] = fi[c];
    __syncthreads();
    const float* G = desc_dst + static_cast<size_t>(mp.d_off) * dim;
    float bestD = __int_as_float(0x7f800000);
    int bestJ = 0x7fffffff;
    for (int j0 = 32 * wib; j0 < mp.Nd; j0 += 32 * kExactWarps) {  // ascending j per lane: strict < keeps the lowest j
      const int nrows = min(32, mp.Nd - j0);
      const int jl = lane < nrows ? j0 + lane : -1;
      float D = 0.0f;
      for (int c0 = 0; c0 < dim; c0 += kExactCols) {
        const int cw = min(kExactCols, dim - c0);
        __syncwarp();
        if (cw == dim) {  // whole rows: one contiguous run
          const float* g = G + static_cast<size_t>(j0) * dim;
          for (int t = lane; t < nrows * dim; t += 32) stage[(t / dim) * kExactPitch + t % dim] = g[t];
        } else {
          for (int r = 0; r < nrows; ++r) {
            const float* g = G + static_cast<size_t>(j0 + r) * dim + c0;
            if (lane < cw) stage[r * kExactPitch + lane] = g[lane];
            if (lane + 32 < cw) stage[r * kExactPitch + lane + 32] = g[lane + 32];
          }
        }
        __syncwarp();
        if (jl >= 0) {
          const float* sr = stage + lane * kExactPitch;
          for (int c = 0; c < cw; ++c) {
            const float e = __fsub_rn(f[c0 + c], sr[c]);
            D = __fmaf_rn(e, e, D);
          }
        }
      }
      if (jl >= 0 && (D < bestD || bestJ == 0x7fffffff)) {
        bestD = D;
        bestJ = jl;
      }
    }
    match_warp_min(bestD, bestJ);
    if (lane == 0) {
      s_D[wib] = bestD;
      s_J[wib] = bestJ;
    }
    __syncthreads();
    if (wib == 0) {
      bestD = lane < kExactWarps ? s_D[lane] : __int_as_float(0x7f800000);
      bestJ = lane < kExactWarps ? s_J[lane] : 0x7fffffff;
      match_warp_min(bestD, bestJ);
      if (bestJ == 0x7fffffff) bestJ = 0;
      match_write(mp, i, bestJ, xyz_src, xyz_dst, nn, corr_src, corr_dst, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------
// 4. mutual-nearest-neighbour filter: correspondence i of a pair is kept iff nn_back[nn[i]] == i (nn: source ->
//    target, nn_back: target -> source, both local to the pair).  Ordered, packed across the pairs:
//      mutual_count_kernel  one CTA per pair: kept[b]
//      mutual_scan_kernel   one CTA: out_offsets = exclusive scan of kept (B + 1 entries)
//      mutual_write_kernel  one CTA per pair: flags again, block-wide ordered compaction into out_src / out_dst
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mutual_keep(const MutualPair& mp, const int32_t* __restrict__ nn,
                                            const int32_t* __restrict__ nn_back, int i) {
  const int j = nn[mp.s_off + i];
  return j >= 0 && j < mp.Nd && nn_back[mp.d_off + j] == i;
}

__global__ void __launch_bounds__(1024) mutual_count_kernel(const MutualPair* __restrict__ pairs,
                                                            const int32_t* __restrict__ nn,
                                                            const int32_t* __restrict__ nn_back,
                                                            unsigned long long* __restrict__ kept) {
  const MutualPair mp = pairs[blockIdx.x];
  __shared__ unsigned int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  unsigned int c = 0;
  for (int i = threadIdx.x; i < mp.Ns; i += 1024) c += mutual_keep(mp, nn, nn_back, i) ? 1u : 0u;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) kept[blockIdx.x] = s_cnt;
}

__global__ void __launch_bounds__(1024) mutual_scan_kernel(const unsigned long long* __restrict__ kept, int B,
                                                           long long* __restrict__ out_offsets) {
  __shared__ unsigned long long part[1024];
  __shared__ unsigned long long carry;
  const int t = threadIdx.x;
  if (t == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + t;
    const unsigned long long v = b < B ? kept[b] : 0ull;
    part[t] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const unsigned long long add = t >= o ? part[t - o] : 0ull;
      __syncthreads();
      part[t] += add;
      __syncthreads();
    }
    if (b < B) out_offsets[b] = static_cast<long long>(carry + part[t] - v);
    __syncthreads();
    if (t == 1023) carry += part[1023];
    __syncthreads();
  }
  if (t == 0) out_offsets[B] = static_cast<long long>(carry);
}

__global__ void __launch_bounds__(1024) mutual_write_kernel(const MutualPair* __restrict__ pairs,
                                                            const int32_t* __restrict__ nn,
                                                            const int32_t* __restrict__ nn_back,
                                                            const float* __restrict__ corr_src,
                                                            const float* __restrict__ corr_dst,
                                                            const long long* __restrict__ out_offsets,
                                                            float* __restrict__ out_src, float* __restrict__ out_dst) {
  const MutualPair mp = pairs[blockIdx.x];
  __shared__ unsigned int wsum[32];
  __shared__ unsigned int s_base;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_base = 0;
  __syncthreads();
  const long long o0 = out_offsets[blockIdx.x];
  for (int i0 = 0; i0 < mp.Ns; i0 += 1024) {
    const int i = i0 + t;
    const bool keep = i < mp.Ns && mutual_keep(mp, nn, nn_back, i);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(m);
    __syncthreads();
    unsigned int before = s_base;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (keep) {
      const long long o = o0 + before + __popc(m & ((1u << lane) - 1u));
      const size_t si = (static_cast<size_t>(mp.s_off) + i) * 3;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        out_src[o * 3 + a] = corr_src[si + a];
        out_dst[o * 3 + a] = corr_dst[si + a];
      }
    }
    __syncthreads();
    if (t == 0) {
      unsigned int tot = 0;
      for (int w = 0; w < 32; ++w) tot += wsum[w];
      s_base += tot;
    }
    __syncthreads();
  }
}

int launch_match_mutual(const LaunchCtx& lc, const MutualPair* d_pairs, int B, const int32_t* d_nn, const int32_t* d_nn_back,
                        const float* d_corr_src, const float* d_corr_dst, unsigned long long* d_kept, long long* d_out_offsets,
                        float* d_out_src, float* d_out_dst) {
  mutual_count_kernel<<<B, 1024, 0, lc.stream>>>(d_pairs, d_nn, d_nn_back, d_kept);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  mutual_scan_kernel<<<1, 1024, 0, lc.stream>>>(d_kept, B, d_out_offsets);
  e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  mutual_write_kernel<<<B, 1024, 0, lc.stream>>>(d_pairs, d_nn, d_nn_back, d_corr_src, d_corr_dst, d_out_offsets, d_out_src, d_out_dst);
  e = cudaGetLastError();
  return e == cudaSuccess ? 3 : -static_cast<int>(e);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int match_chunks(int dim) { return ((3 * dim + 3 + 15) / 16) * 2; }  // 16-byte chunks of 8 bf16; K padded to 16

int match_stages(int chunks) { return chunks <= 14 ? kStagesB : 2; }  // 3 x 64 KB stages + the A tile exceed 227 KB
size_t match_smem_bytes(int chunks) {
  return static_cast<size_t>(kMatchTileM + match_stages(chunks) * kMatchTileN) * chunks * 16 +
         static_cast<size_t>(kMatchEpiThreads) * (kMatchCand + 2) * 4 + 16 * 8 + 16;
}

int match_configure() {
  size_t most = 0;
  for (int dim = 1; dim <= kMatchMaxDim; ++dim) most = std::max(most, match_smem_bytes(match_chunks(dim)));
  cudaError_t e = cudaFuncSetAttribute(match_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(most));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(match_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(match_scan_smem(256)));
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

int launch_match_prep(const LaunchCtx& lc, const MatchPair* d_pairs, int pairs, int max_tiles_rows, const float* d_desc,
                      int side, int dim, unsigned char* d_img, float* d_norms, uint32_t* d_bmax) {
  const int chunks = match_chunks(dim);
  const long long items = static_cast<long long>(max_tiles_rows) * chunks;
  match_prep_kernel<<<dim3(static_cast<unsigned>((items + 255) / 256), pairs), 256, 0, lc.stream>>>(d_pairs, d_desc, side, dim,
                                                                                                 chunks, d_img, d_norms, d_bmax);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

int launch_match_mma(const LaunchCtx& lc, const MatchPair* d_pairs, int pairs, int max_s_tiles, int dim,
                     const unsigned char* d_img, const float* d_norms, const uint32_t* d_bmax, int32_t* d_cand,
                     int32_t* d_cand_cnt, int dbg) {
  const int chunks = match_chunks(dim);
  match_mma_kernel<<<dim3(max_s_tiles, pairs), kMatchThreadsAll, match_smem_bytes(chunks), lc.stream>>>(
      d_pairs, d_img, d_norms, d_bmax, chunks, match_stages(chunks), d_cand, d_cand_cnt, dbg);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

int launch_match_exact(const LaunchCtx& lc, const MatchPair* d_pairs, int pairs, int max_ns, const float* d_desc_src,
                       const float* d_desc_dst, const float* d_xyz_src, const float* d_xyz_dst, int dim,
                       const int32_t* d_cand, const int32_t* d_cand_cnt, int32_t* d_nn, float* d_corr_src,
                       float* d_corr_dst, int32_t* d_work_list, uint32_t* d_work_count) {
  if (d_cand == nullptr) {  // no lists: every row is scanned
    match_scan_kernel<<<dim3(max_ns, pairs), kExactWarps * 32, match_scan_smem(dim), lc.stream>>>(
        d_pairs, d_desc_src, d_desc_dst, d_xyz_src, d_xyz_dst, dim, nullptr, nullptr, d_nn, d_corr_src, d_corr_dst);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -static_cast<int>(e);
  }
  cudaError_t e = cudaMemsetAsync(d_work_count, 0, sizeof(uint32_t), lc.stream);
  if (e != cudaSuccess) return -static_cast<int>(e);
  match_exact_kernel<<<dim3((max_ns + kExactWarps - 1) / kExactWarps, pairs), kExactWarps * 32,
                       static_cast<size_t>(kExactWarps) * dim * sizeof(float), lc.stream>>>(
      d_pairs, d_desc_src, d_desc_dst, d_xyz_src, d_xyz_dst, dim, d_cand, d_cand_cnt, d_nn, d_corr_src, d_corr_dst, d_work_list,
      d_work_count);
  e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  // rows whose list overflowed (rare: near-identical descriptors); the CTAs find the count on the device
  match_scan_kernel<<<dim3(2 * lc.sm_count, 1), kExactWarps * 32, match_scan_smem(dim), lc.stream>>>(
      d_pairs, d_desc_src, d_desc_dst, d_xyz_src, d_xyz_dst, dim, d_work_list, d_work_count, d_nn, d_corr_src, d_corr_dst);
  e = cudaGetLastError();
  return e == cudaSuccess ? 2 : -static_cast<int>(e);
}

}  // namespace saccot

// kernels_graph_mma.cu — S1 with the squared distances taken from the tensor cores.
//
// The compatibility test needs x = |s_i - s_j|^2 and y = |d_i - d_j|^2 for every pair of correspondences: 12 of the
// 19 FP32-pipe cycles the CUDA-core kernel (kernels_graph.cu) spends per pair test.  Both are dense contractions,
//     |p_i - p_j|^2 = |p_i|^2 + |p_j|^2 - 2 p_i . p_j,
// so this kernel takes them from tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) and keeps only the decision on the CUDA
// cores.  The DECISION stays bit-exact against the specified predicate | RN(RN(sqrt x) - RN(sqrt y)) | < tau:
//   * every fp32 coordinate is split into three bf16 pieces h + m + l (24 mantissa bits); K carries the products
//     hh, hm, mh, hl, lh, mm of each coordinate (what is dropped is below 4 x 2^-27 |p_i||p_j|), the three bf16
//     pieces of |p_i|^2 against 1 and 1 against the pieces of |p_j|^2, with the factor -2 folded into the row operand:
//     K = 24, padded to 32 = two MMAs of K = 16 per cloud and tile.  The accumulator is x~ with
//     |x~ - |p_i - p_j|^2| <= eps := kappa (|p_i|^2 + |p_j|^2): the fp32 accumulation of 24 terms of total magnitude
//     <= 2.1 (|p_i|^2 + |p_j|^2) is charged 2^-21 per term, the split 3 x 2^-26; kappa = 2^-15 leaves a factor of two
//     (`graph_dbg` measures the actual error against exact arithmetic: tests/test_gpu_parity.py).
//   * the sqrt-free filter of kernels_graph.cu runs on (x~, y~) with a wider threshold:
//         S = x~ + y~;  U = S - tau^2;  Q = fma(x~ y~, -4, U U);  Theta' = 2^-18 S^2 + c1 S + c0
//     Q*(x, y) = ((a-b)^2 - tau^2)((a+b)^2 - tau^2) is quadratic: Q*(x+dx, y+dy) - Q*(x, y) = 2(x-y-tau^2) dx +
//     2(y-x-tau^2) dy + (dx-dy)^2, so with |dx|, |dy| <= eps + 3u S (x~ approximates the exact squared distance, the
//     specified x is within 3u of it) the shift is at most 4 (S + tau^2) eps + 4 eps^2 + 12u S^2.  c1 = 4.2 eps_t,
//     c0 = 4.2 eps_t (tau^2 + 2 eps_t) with eps_t = kappa (largest |p|^2 of the row tile + of the column tile, both
//     clouds) cover the eps terms; 2^-18 S^2 = 63.8u S^2 covers 12u + the 7.6u of the filter's own rounding + the
//     8.5u band in which the literal predicate may differ from the real-arithmetic one.  sign(Q) is trusted iff
//     |Q| > Theta' and S > max(4 tau^2, 2^-50) + 2 eps_t; everything else (about 1 % of the pairs at kappa = 2^-15,
//     pad / NaN / out-of-range values) takes the literal sqrt.rn evaluation on the fp32 coordinates, four columns at a
//     time, exactly as in the CUDA-core kernel.
//
// Structure: one CTA per 128-row block I of a pair walks the column blocks J = I .. nblk-1 (upper triangle; the
// mirrored tile comes from the same bits).  Warp 16: bulk copies of the column block's operand images and raw
// coordinates into a 3-stage ring, four MMAs per tile (two clouds x K 32) into double-buffered TMEM accumulators.
// Warps 0-15: thread (row r = 32 (w % 4) + lane, word w / 4) reads 32 values of x~ and y~ with tcgen05.ld, decides
// 32 columns and owns one 32-bit word of its row; words are gathered through shared memory into the 128-bit row
// stores, the 32 x 32 shuffle transpose gives the mirrored tile, edge counts go to the triangle work units.
#include "common.cuh"

#include <cuda_bf16.h>

#include <cstdio>

namespace saccot {

namespace {

constexpr int kGmEpiWarps = 16;
constexpr int kGmThreads = 32 * (kGmEpiWarps + 1);
constexpr int kGmStages = 3;
constexpr int kGmImgTile = 128 * 64;          // one cloud's operand image of a 128-point block: 32 bf16 per point
constexpr int kGmRawBytes = 6 * 128 * 4;      // raw coordinates of a column block (literal fallback)
constexpr int kGmStageBytes = 2 * kGmImgTile + kGmRawBytes;
constexpr float kGmKappa = 1.0f / 32768.0f;   // 2^-15

__device__ __forceinline__ uint64_t gm_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((128u >> 4) & 0x3FFF) << 16;   // LBO: next 16-byte K chunk
  d |= static_cast<uint64_t>((512u >> 4) & 0x3FFF) << 32;   // SBO: next 8-row group (4 chunks x 128 B)
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void gm_wait(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void gm_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gm_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#define SACCOT_GM_LD32(v, taddr)                                                                                       \
  asm volatile(                                                                                                        \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19," \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                       \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),    \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),          \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),         \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                       \
      : "r"(taddr))

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 gpack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void gunpack(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 gadd2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 gmul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 gfma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// literal evaluation of the specified predicate for four consecutive columns of the staged column block
__device__ __noinline__ uint32_t gm_literal4(const float* __restrict__ cs /* [6][128] */, int c, float sxi, float syi,
                                             float szi, float dxi, float dyi, float dzi, float tau) {
  uint32_t bits = 0;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const float a = __fsub_rn(sxi, cs[c + k]), b = __fsub_rn(syi, cs[128 + c + k]), cc = __fsub_rn(szi, cs[256 + c + k]);
    const float x = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(cc, cc));
    const float e = __fsub_rn(dxi, cs[384 + c + k]), f = __fsub_rn(dyi, cs[512 + c + k]), g = __fsub_rn(dzi, cs[640 + c + k]);
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(e, e), __fmul_rn(f, f)), __fmul_rn(g, g));
    if (fabsf(__fsub_rn(__fsqrt_rn(x), __fsqrt_rn(y))) < tau) bits |= 1u << k;
  }
  return bits;
}

__device__ __forceinline__ uint32_t gm_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
    x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
  }
  return x;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Operand images.  One thread per (point, cloud).  Row image (M side) and column image (N side) of a 128-point
// block are 8 KB each, laid out [8-row group][16-byte K chunk][8 rows][16 B] (UMMA canonical, no swizzle, K-major);
// block t of the pair holds [cloud 0 | cloud 1] back to back in each of the two image arrays.  Pad points (n >= N)
// give zero rows (their bits are skipped, not computed).  nmax[t] = largest |p|^2 of either cloud in block t.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) graph_prep_kernel(const PairDesc* __restrict__ descs, const float* __restrict__ soa,
                                                         unsigned char* __restrict__ img_row,
                                                         unsigned char* __restrict__ img_col, uint32_t* __restrict__ nmax,
                                                         int nmax_pitch) {
  const PairDesc d = descs[blockIdx.y];
  const int item = blockIdx.x * 256 + threadIdx.x;
  const int n = item >> 1, cloud = item & 1;
  if (n >= d.Npad) return;
  const float* base = soa + d.soa_off + static_cast<size_t>(3 * cloud) * d.Npad;
  float p[3] = {0.0f, 0.0f, 0.0f};
  const bool valid = n < d.N;
  if (valid) {
    p[0] = base[n];
    p[1] = base[d.Npad + n];
    p[2] = base[2 * static_cast<size_t>(d.Npad) + n];
  }
  const float nrm = __fmaf_rn(p[2], p[2], __fmaf_rn(p[1], p[1], __fmul_rn(p[0], p[0])));
  __nv_bfloat16 h[3], m[3], l[3], np[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    h[c] = __float2bfloat16_rn(p[c]);
    const float r1 = __fsub_rn(p[c], __bfloat162float(h[c]));
    m[c] = __float2bfloat16_rn(r1);
    l[c] = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(m[c])));
  }
  {
    np[0] = __float2bfloat16_rn(nrm);
    const float r1 = __fsub_rn(nrm, __bfloat162float(np[0]));
    np[1] = __float2bfloat16_rn(r1);
    np[2] = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(np[1])));
  }
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.0f), one = __float2bfloat16_rn(valid ? 1.0f : 0.0f);
  __nv_bfloat16 row[32], col[32];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const __nv_bfloat16 h2 = __float2bfloat16_rn(-2.0f * __bfloat162float(h[c]));  // exact: a power-of-two factor
    const __nv_bfloat16 m2 = __float2bfloat16_rn(-2.0f * __bfloat162float(m[c]));
    const __nv_bfloat16 l2 = __float2bfloat16_rn(-2.0f * __bfloat162float(l[c]));
    // products hh, hm, mh, hl, lh, mm
    row[6 * c + 0] = h2; col[6 * c + 0] = h[c];
    row[6 * c + 1] = h2; col[6 * c + 1] = m[c];
    row[6 * c + 2] = m2; col[6 * c + 2] = h[c];
    row[6 * c + 3] = h2; col[6 * c + 3] = l[c];
    row[6 * c + 4] = l2; col[6 * c + 4] = h[c];
    row[6 * c + 5] = m2; col[6 * c + 5] = m[c];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    row[18 + k] = np[k]; col[18 + k] = one;   // |p_i|^2 x 1
    row[21 + k] = one;   col[21 + k] = np[k]; // 1 x |p_j|^2
  }
#pragma unroll
  for (int k = 24; k < 32; ++k) { row[k] = zero; col[k] = zero; }
  const int t = n >> 7, r = n & 127;
  const size_t off = static_cast<size_t>(d.gimg_off) + (static_cast<size_t>(t) * 2 + cloud) * kGmImgTile +
                     static_cast<size_t>(r >> 3) * 512 + (r & 7) * 16;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    *reinterpret_cast<uint4*>(img_row + off + c * 128) = *reinterpret_cast<const uint4*>(row + 8 * c);
    *reinterpret_cast<uint4*>(img_col + off + c * 128) = *reinterpret_cast<const uint4*>(col + 8 * c);
  }
  if (valid) atomicMax(nmax + static_cast<size_t>(blockIdx.y) * nmax_pitch + t, __float_as_uint(nrm));
}

// ------------------------------------------------------------------------------------------
// The graph kernel.  grid (max row blocks, pairs); block 544 threads.
// ------------------------------------------------------------------------------------------
template <bool DBG>
__global__ void __launch_bounds__(kGmThreads, 1) graph_mma_kernel(
    const PairDesc* __restrict__ descs, const float* __restrict__ soa, const unsigned char* __restrict__ img_row,
    const unsigned char* __restrict__ img_col, const uint32_t* __restrict__ nmax, int nmax_pitch,
    uint32_t* __restrict__ adj, uint32_t* __restrict__ panel, uint32_t* __restrict__ ucount, int unit_pitch, float tau,
    float tau2f, float lo, float* __restrict__ dbg_err) {
  const PairDesc d = descs[blockIdx.y];
  const int I = blockIdx.x;
  if (I >= d.nblk) return;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                                   // [2 clouds][8 KB] row images of block I
  unsigned char* sB = smem + 2 * kGmImgTile;                  // kGmStages x (2 x 8 KB column images + raw coordinates)
  uint32_t* tsm = reinterpret_cast<uint32_t*>(sB + kGmStages * kGmStageBytes);  // [2 tile parities][2 (direct, mirrored)][128][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + 2 * 2 * 128 * 4);
  uint64_t* a_full = bars;
  uint64_t* full = bars + 1;                 // [3] column block landed
  uint64_t* stage_free = full + kGmStages;   // [3] epilogue done with the stage (16 warps)
  uint64_t* mma_done = stage_free + kGmStages;  // [2]
  uint64_t* acc_free = mma_done + 2;         // [2] epilogue done with the accumulators (16 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);
  unsigned int* tile_cnt = tmem_slot + 2;    // [2] edges (i < j) of the tile, per tile parity

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < kGmStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&stage_free[s], kGmEpiWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&mma_done[b], 1);
      mbar_init(&acc_free[b], kGmEpiWarps);
    }
    tile_cnt[0] = 0u;
    tile_cnt[1] = 0u;
    mbar_fence_init();
  }
  if (warp == kGmEpiWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;
  const int T = d.nblk - I;  // column blocks J = I + n
  const unsigned char* gA = img_row + d.gimg_off + static_cast<size_t>(I) * 2 * kGmImgTile;
  const unsigned char* gB = img_col + d.gimg_off;
  const float* base = soa + d.soa_off;

  if (warp == kGmEpiWarps) {
    // ================================ producer + MMA issuer (one lane) ================================
    if (lane == 0) {
      // kind::f16: D = F32, A = B = BF16, K-major, N = 128, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      auto load = [&](int n) {
        const int s = n % kGmStages, J = I + n;
        unsigned char* st = sB + s * kGmStageBytes;
        mbar_arrive_expect_tx(&full[s], static_cast<uint32_t>(kGmStageBytes));
        bulk_g2s(st, gB + static_cast<size_t>(J) * 2 * kGmImgTile, 2 * kGmImgTile, &full[s]);
#pragma unroll
        for (int c = 0; c < 6; ++c)
          bulk_g2s(st + 2 * kGmImgTile + c * 512, base + static_cast<size_t>(c) * d.Npad + J * 128, 512, &full[s]);
      };
      mbar_arrive_expect_tx(a_full, 2 * kGmImgTile);
      bulk_g2s(sA, gA, 2 * kGmImgTile, a_full);
      for (int n = 0; n < 2 && n < T; ++n) load(n);
      gm_wait(a_full, 0u);
      for (int n = 0; n < T; ++n) {
        const int s = n % kGmStages, b = n & 1;
        gm_wait(&full[s], static_cast<uint32_t>((n / kGmStages) & 1));
        if (n >= 2) gm_wait(&acc_free[b], static_cast<uint32_t>(((n >> 1) - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
        for (int cloud = 0; cloud < 2; ++cloud) {
          const uint64_t dA = gm_desc(smem_u32(sA + cloud * kGmImgTile));
          const uint64_t dB = gm_desc(smem_u32(sB + s * kGmStageBytes + cloud * kGmImgTile));
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t acc = k ? 1u : 0u;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem + static_cast<uint32_t>(256 * b + 128 * cloud)),
                "l"(dA + static_cast<uint64_t>((k * 256) >> 4)), "l"(dB + static_cast<uint64_t>((k * 256) >> 4)), "r"(idesc), "r"(acc)
                : "memory");
          }
        }
        gm_commit(&mma_done[b]);
        if (n + 2 < T) {  // the stage tile n - 1 used is free once the epilogue has finished with its raw coordinates
          const int s2 = (n + 2) % kGmStages, prev = n + 2 - kGmStages;
          if (prev >= 0) gm_wait(&stage_free[s2], static_cast<uint32_t>((prev / kGmStages) & 1));
          load(n + 2);
        }
      }
    }
  } else {
    // ================================ decision warps ================================
    const int q = warp & 3, wb = warp >> 2;   // lane quadrant, 32-column word of the tile
    const int r = 32 * q + lane;              // row of the block = TMEM lane
    const int I0 = I * 128;
    const bool row_valid = I0 + r < d.N;
    float rv[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) rv[c] = base[static_cast<size_t>(c) * d.Npad + I0 + r];
    const float nI = __uint_as_float(nmax[static_cast<size_t>(blockIdx.y) * nmax_pitch + I]);
    uint32_t* adjp = adj + d.adj_off;
    uint32_t* pp = panel != nullptr ? panel + d.panel_off : nullptr;
    const f32x2 ntau2 = gpack(-tau2f, -tau2f), m4 = gpack(-4.0f, -4.0f), kth = gpack(3.814697265625e-06f, 3.814697265625e-06f);  // 2^-18
    float err_max = 0.0f;
    unsigned int n_groups = 0, n_unsure = 0;
    for (int n = 0; n < T; ++n) {
      const int b = n & 1, s = n % kGmStages, J = I + n, J0 = J * 128;
      const float* cs = reinterpret_cast<const float*>(sB + s * kGmStageBytes + 2 * kGmImgTile);  // [6][128] raw columns
      const float eps = __fmul_rn(kGmKappa, __fadd_rn(nI, __uint_as_float(nmax[static_cast<size_t>(blockIdx.y) * nmax_pitch + J])));
      const float c1s = __fmul_rn(4.2f, eps), c0s = __fmul_rn(c1s, __fadd_rn(tau2f, __fadd_rn(eps, eps)));
      const float lo2 = __fadd_rn(lo, __fadd_rn(eps, eps));
      const f32x2 c1 = gpack(c1s, c1s), c0 = gpack(c0s, c0s);
      gm_wait(&mma_done[b], static_cast<uint32_t>((n >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t tb = tmem + ((32u * q) << 16) + static_cast<uint32_t>(256 * b + 32 * wb);
      uint32_t xv[32], yv[32];
      SACCOT_GM_LD32(xv, tb);
      SACCOT_GM_LD32(yv, tb + 128u);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int ncv = min(128, d.N - J0) - 32 * wb;  // valid columns of this word (may be <= 0)
      uint32_t wbits = 0;
#pragma unroll
      for (int g = 28; g >= 0; g -= 4) {   // groups of four columns, from the top of the word down
        if (g >= ncv || !row_valid) {     // pad columns (uniform over the warp) / pad rows: zero bits, nothing computed
          wbits <<= 4;
          continue;
        }
        bool sure = true;
#pragma unroll
        for (int h2 = 1; h2 >= 0; --h2) {
          const int e = g + 2 * h2;
          const f32x2 x = gpack(__uint_as_float(xv[e]), __uint_as_float(xv[e + 1]));
          const f32x2 y = gpack(__uint_as_float(yv[e]), __uint_as_float(yv[e + 1]));
          const f32x2 S = gadd2(x, y);
          const f32x2 U = gadd2(S, ntau2);
          const f32x2 Q = gfma2(gmul2(x, y), m4, gmul2(U, U));
          const f32x2 Th = gfma2(S, c1, gfma2(gmul2(S, S), kth, c0));
          float q0, q1, t0, t1, s0, s1;
          gunpack(Q, q0, q1);
          gunpack(Th, t0, t1);
          gunpack(S, s0, s1);
          sure = sure && fabsf(q0) > t0 && s0 > lo2 && fabsf(q1) > t1 && s1 > lo2;
          wbits = __funnelshift_l(__float_as_uint(q1), wbits, 1);
          wbits = __funnelshift_l(__float_as_uint(q0), wbits, 1);
        }
        if (!sure) wbits = (wbits & ~0xFu) | gm_literal4(cs, 32 * wb + g, rv[0], rv[1], rv[2], rv[3], rv[4], rv[5], tau);
        if (DBG) {
          ++n_groups;
          n_unsure += sure ? 0u : 1u;
        }
      }
      if (DBG && row_valid) {  // tests: largest |x~ - exact| / (|p_i|^2 + |p_j|^2) seen (static indices: the
                               // accumulator values must stay in registers in the product instantiation)
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (e >= ncv) continue;
          const int c = 32 * wb + e;
          double ex = 0.0, ey = 0.0, ni = 0.0, nj = 0.0, mi = 0.0, mj = 0.0;
          for (int a = 0; a < 3; ++a) {
            const double pi = rv[a], pj = cs[a * 128 + c], qi = rv[3 + a], qj = cs[(3 + a) * 128 + c];
            ex += (pi - pj) * (pi - pj);
            ey += (qi - qj) * (qi - qj);
            ni += pi * pi; nj += pj * pj; mi += qi * qi; mj += qj * qj;
          }
          if (ni + nj > 0.0) err_max = fmaxf(err_max, static_cast<float>(fabs(static_cast<double>(__uint_as_float(xv[e])) - ex) / (ni + nj)));
          if (mi + mj > 0.0) err_max = fmaxf(err_max, static_cast<float>(fabs(static_cast<double>(__uint_as_float(yv[e])) - ey) / (mi + mj)));
        }
      }
      // the accumulators and the stage's raw coordinates are no longer needed by this warp
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) {
        gm_arrive(&acc_free[b]);
        gm_arrive(&stage_free[s]);
      }
      // ---- diagonal tile: A_ii = 0; edge count of the tile (i < j only) ----
      unsigned int cnt;
      if (J == I) {
        if (wb == (r >> 5)) wbits &= ~(1u << (r & 31));
        uint32_t upper;
        if (wb > (r >> 5)) upper = 0xffffffffu;
        else if (wb < (r >> 5)) upper = 0u;
        else upper = (r & 31) == 31 ? 0u : (0xffffffffu << ((r & 31) + 1));
        cnt = __popc(wbits & upper);
      } else {
        cnt = __popc(wbits);
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0 && cnt) atomicAdd(&tile_cnt[n & 1], cnt);
      // ---- gather the row's four words (direct tile) and the transposed words (mirrored tile) ----
      uint32_t* td = tsm + (n & 1) * 1024;   // [128][4] direct
      uint32_t* tm = td + 512;               // [128][4] mirrored: row = column of the tile, word = this row quadrant
      td[r * 4 + wb] = wbits;
      if (J != I) tm[(32 * wb + lane) * 4 + q] = gm_transpose32(wbits, lane);
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (wb == 0) {
        const uint4 w4 = *reinterpret_cast<const uint4*>(td + r * 4);
        *reinterpret_cast<uint4*>(adjp + static_cast<size_t>(I0 + r) * d.stride + J0 / 32) = w4;
        if (pp != nullptr) {
          uint32_t* prow = pp + (static_cast<size_t>(J >> 1) * d.Npad + I0 + r) * 8;
          *reinterpret_cast<uint4*>(prow + (J & 1) * 4) = w4;
          if (J == d.nblk - 1 && (J & 1) == 0) *reinterpret_cast<uint4*>(prow + 4) = make_uint4(0u, 0u, 0u, 0u);  // half panel
        }
      } else if (wb == 1 && J != I) {
        const uint4 w4 = *reinterpret_cast<const uint4*>(tm + r * 4);
        *reinterpret_cast<uint4*>(adjp + static_cast<size_t>(J0 + r) * d.stride + I0 / 32) = w4;
        if (pp != nullptr) *reinterpret_cast<uint4*>(pp + (static_cast<size_t>(I >> 1) * d.Npad + J0 + r) * 8 + (I & 1) * 4) = w4;
      } else if (wb == 2 && r == 0) {
        const unsigned int total = tile_cnt[n & 1];
        tile_cnt[n & 1] = 0u;  // next use: tile n + 2, after another bar.sync
        if (total)
          atomicAdd(&ucount[static_cast<size_t>(blockIdx.y) * unit_pitch + unit_offset(static_cast<unsigned int>(J)) + (I >> 1)], total);
      }
    }
    if (DBG) {
      err_max = fmaxf(err_max, __shfl_xor_sync(0xffffffffu, err_max, 16));
      err_max = fmaxf(err_max, __shfl_xor_sync(0xffffffffu, err_max, 8));
      err_max = fmaxf(err_max, __shfl_xor_sync(0xffffffffu, err_max, 4));
      err_max = fmaxf(err_max, __shfl_xor_sync(0xffffffffu, err_max, 2));
      err_max = fmaxf(err_max, __shfl_xor_sync(0xffffffffu, err_max, 1));
      if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(dbg_err), __float_as_uint(err_max));
      n_groups = __reduce_add_sync(0xffffffffu, n_groups);
      n_unsure = __reduce_add_sync(0xffffffffu, n_unsure);
      if (lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(dbg_err) + 1, static_cast<unsigned long long>(n_groups));
        atomicAdd(reinterpret_cast<unsigned long long*>(dbg_err) + 2, static_cast<unsigned long long>(n_unsure));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == kGmEpiWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// the CTA owns all 512 TMEM columns: ask for more than half of the shared memory so that no second CTA becomes
// resident on the SM only to wait for the allocation
static size_t graph_mma_smem() {
  const size_t need = 2 * kGmImgTile + kGmStages * kGmStageBytes + 2 * 2 * 128 * 4 * 4 + 16 * 8 + 32;
  return need > 120 * 1024 ? need : 120 * 1024;
}

size_t graph_mma_image_bytes(int nblk) { return static_cast<size_t>(nblk) * 2 * kGmImgTile; }

int graph_mma_configure() {
  cudaError_t e = cudaFuncSetAttribute(graph_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(graph_mma_smem()));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(graph_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(graph_mma_smem()));
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

int launch_graph_mma(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_nblk, const float* d_soa,
                     unsigned char* d_img_row, unsigned char* d_img_col, uint32_t* d_nmax, int nmax_pitch, uint32_t* d_adj,
                     uint32_t* d_panel, uint32_t* d_ucount, int unit_pitch, float tau, float* d_dbg_err) {
  const float tau2f = tau * tau;
  float lo = 4.0f * tau2f;
  if (!(lo >= 8.8817841970012523e-16f)) lo = 8.8817841970012523e-16f;  // 2^-50; also replaces NaN
  cudaError_t e = cudaMemsetAsync(d_nmax, 0, sizeof(uint32_t) * static_cast<size_t>(nmax_pitch) * pairs, lc.stream);
  if (e != cudaSuccess) return -static_cast<int>(e);
  graph_prep_kernel<<<dim3((2 * max_nblk * 128 + 255) / 256, pairs), 256, 0, lc.stream>>>(d_desc, d_soa, d_img_row, d_img_col,
                                                                                         d_nmax, nmax_pitch);
  e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  if (d_dbg_err != nullptr)
    graph_mma_kernel<true><<<dim3(max_nblk, pairs), kGmThreads, graph_mma_smem(), lc.stream>>>(
        d_desc, d_soa, d_img_row, d_img_col, d_nmax, nmax_pitch, d_adj, d_panel, d_ucount, unit_pitch, tau, tau2f, lo, d_dbg_err);
  else
    graph_mma_kernel<false><<<dim3(max_nblk, pairs), kGmThreads, graph_mma_smem(), lc.stream>>>(
        d_desc, d_soa, d_img_row, d_img_col, d_nmax, nmax_pitch, d_adj, d_panel, d_ucount, unit_pitch, tau, tau2f, lo, d_dbg_err);
  e = cudaGetLastError();
  return e == cudaSuccess ? 2 : -static_cast<int>(e);
}

}  // namespace saccot

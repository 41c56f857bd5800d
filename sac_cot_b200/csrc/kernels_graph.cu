// kernels_graph.cu — S0 (AoS -> padded SoA repack) and S1 (first-order length-consistency
// compatibility graph), plus the tiny per-chunk key-pool scan.
//
// S1 arithmetic contract (bit-exact against the oracle): every fp32 operation is an
// individually rounded IEEE op (explicit .rn PTX / __f*_rn intrinsics, never contracted into
// FMAs: equivalent to -fmad=false on this kernel), in the order
//   a=sx_i-sx_j; b=..; c=..; s2=(a*a+b*b)+c*c; ls=sqrt(s2); (same for dst -> ld);
//   A_ij = |ls-ld| < tau_c
// where the final decision is taken by an exact sqrt-free filter with a literal fallback (below).
// Only tiles on or above the diagonal are evaluated; the mirrored tile is emitted from the
// same predicate bits (negating a,b,c leaves the squares unchanged, so the mirrored entry is
// bit-identical to evaluating it directly).
#include "common.cuh"

namespace saccot {

// ------------------------------------------------------------------------------------------
// S0: pack the caller's AoS points into six padded SoA arrays.  Pad lanes hold NaN, so every
// predicate that involves a pad point is false and pad bits/rows of the adjacency are zero.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_soa_kernel(const PairDesc* __restrict__ descs,
                                                       const float* __restrict__ src,
                                                       const float* __restrict__ dst, float* __restrict__ soa) {
  const PairDesc d = descs[blockIdx.y];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.Npad) return;
  float v[6];
  if (n < d.N) {
    const float* s = src + 3 * (d.pt_off + n);
    const float* t = dst + 3 * (d.pt_off + n);
    v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
    v[3] = t[0]; v[4] = t[1]; v[5] = t[2];
  } else {
    const float nan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int c = 0; c < 6; ++c) v[c] = nan;
  }
  float* o = soa + d.soa_off + n;
#pragma unroll
  for (int c = 0; c < 6; ++c) o[static_cast<size_t>(c) * d.Npad] = v[c];
}

int launch_pack_soa(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_npad, const float* d_src,
                    const float* d_dst, float* d_soa) {
  dim3 grid((max_npad + 255) / 256, pairs);
  pack_soa_kernel<<<grid, 256, 0, lc.stream>>>(d_desc, d_src, d_dst, d_soa);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

// ------------------------------------------------------------------------------------------
// S1: one CTA per 128 x 128 tile (I <= J) of one pair.  Thread r owns row I0+r: its point lives
// in registers, the 128 column points are staged in shared memory by six 512-byte bulk copies
// (TMA, completion on an mbarrier) and read as warp-wide broadcasts, two columns per load.
// Each thread builds the four 32-bit words of its row and stores them as one 128-bit word; the
// mirrored tile is produced by a shuffle-based 32x32 bit transpose per warp, staged through
// shared memory and stored the same way.
//
// Exact decision without square roots.  The specified predicate is
//     P(x, y) = | RN(RN(sqrt x) - RN(sqrt y)) | < tau          x = s2, y = d2 (fp32, as above)
// Evaluating it literally costs two sqrt.rn expansions (~20 instructions).  Instead the kernel
// computes, in fp32 (u = 2^-24, every op individually rounded),
//     S = x + y;  U = S - tau2f;  Q = U*U - 4*(x*y);  Theta = (U*U) * (3 * 2^-20)
// with tau2f = RN(tau*tau).  In real arithmetic  Q* = ((a-b)^2 - tau^2)((a+b)^2 - tau^2), a = sqrt x,
// b = sqrt y, so for (a+b) > 2 tau its sign is the sign of |a-b| - tau.  Error analysis (DESIGN.md
// "S1 exact filter"): if S > lo = max(4 tau2f, 2^-50) then |Q - Q*| <= 7.6u (x+y)^2 (+ 8u (x+y)^2 because the
// filter forms x and y with fused multiply-adds, within 4u of the specified values), and whenever
// the literal predicate could disagree with sign(|a-b| - tau) (rounding of the two square roots
// and of their difference, at most 2.1u (a+b) in total) one has |Q*| <= 8.5u (x+y)^2.  Hence
//     |Q| > Theta  (S > 4 tau2f gives U >= 0.75 S, so Theta >= 26.9u (x+y)^2 > 7.6u + 8u + 8.5u)   ==>   P(x, y) == (Q < 0)
// exactly.  (Theta re-uses the product U*U that Q needs anyway: one packed multiply less than (S*S) * 2^-19; the
// factor 48u keeps the band of undecided pairs at 1.5x that version's where U ~ S, the usual case.)
// Everything else — the ~1e-4 fraction of pairs inside the band, tiny/huge/NaN inputs, tau so
// large that lo overflows — takes the literal sqrt.rn evaluation.  The result is bit-identical to
// the oracle's for every input; tests/test_gpu_parity.py has adversarial near-threshold sets.
// ------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;  // two packed fp32 lanes (sm_100 FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Squared length for the FILTER only: a*a, then two fused multiply-adds (3 packed instructions instead of 3 packed
// multiplies and 4 scalar additions: the kernel is bound by the FP32 pipe, 23 -> 19 pipe cycles per pair test).  It
// differs from the specified sum ((a*a + b*b) + c*c, every operation rounded) by at most 4u relative; the filter's
// threshold Theta absorbs that, and whatever the filter is not sure about is decided by
// compat_literal4, which computes the SPECIFIED squared lengths itself.
__device__ __forceinline__ f32x2 sum_sq_fused(f32x2 a, f32x2 b, f32x2 c) { return fma2(c, c, fma2(b, b, mul2(a, a))); }

// Literal evaluation of the specified predicate for the four columns c..c+3 of the staged tile:
// the rare fallback of the filter.  Out of line and self-contained (recomputes the squared lengths
// with the same individually rounded operations) so that it costs the hot loop nothing but a flag.
__device__ __noinline__ uint32_t compat_literal4(const float (*cs)[128], int c, float sxi, float syi, float szi,
                                                 float dxi, float dyi, float dzi, float tau) {
  uint32_t bits = 0;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const float a = __fsub_rn(sxi, cs[0][c + k]), b = __fsub_rn(syi, cs[1][c + k]), cc = __fsub_rn(szi, cs[2][c + k]);
    const float x = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(cc, cc));
    const float e = __fsub_rn(dxi, cs[3][c + k]), f = __fsub_rn(dyi, cs[4][c + k]), g = __fsub_rn(dzi, cs[5][c + k]);
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(e, e), __fmul_rn(f, f)), __fmul_rn(g, g));
    if (fabsf(__fsub_rn(__fsqrt_rn(x), __fsqrt_rn(y))) < tau) bits |= 1u << k;
  }
  return bits;
}

// out[lane] bit r = in[r] bit lane, for the 32 lanes of a warp
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
    x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
  }
  return x;
}

__global__ void __launch_bounds__(128) graph_kernel(const PairDesc* __restrict__ descs,
                                                    const float* __restrict__ soa, uint32_t* __restrict__ adj,
                                                    uint32_t* __restrict__ panel, uint32_t* __restrict__ ucount,
                                                    int unit_pitch, float tau,
                                                    float tau2f, float lo) {
  const PairDesc d = descs[blockIdx.y];
  const int ntiles = d.nblk * (d.nblk + 1) / 2;
  const int tile = blockIdx.x;
  if (tile >= ntiles) return;
  // tile = J(J+1)/2 + I with I <= J
  int J = static_cast<int>((sqrtf(8.0f * static_cast<float>(tile) + 1.0f) - 1.0f) * 0.5f);
  while (J * (J + 1) / 2 > tile) --J;
  while ((J + 1) * (J + 2) / 2 <= tile) ++J;
  const int I = tile - J * (J + 1) / 2;
  const int I0 = I * 128, J0 = J * 128;

  __shared__ __align__(128) float cs[6][128];
  __shared__ __align__(16) uint32_t tsm[128][4];
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned int wcnt[4];

  const int r = threadIdx.x;
  const int lane = r & 31, warp = r >> 5;
  const float* base = soa + d.soa_off;

  if (r == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (r == 0) {
    mbar_arrive_expect_tx(&bar, 6 * 128 * sizeof(float));
#pragma unroll
    for (int c = 0; c < 6; ++c) bulk_g2s(&cs[c][0], base + static_cast<size_t>(c) * d.Npad + J0, 128 * sizeof(float), &bar);
  }
  // own row point (coalesced), overlapped with the bulk copies
  float rv[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) rv[c] = base[static_cast<size_t>(c) * d.Npad + I0 + r];
  f32x2 ri[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) ri[c] = pack2(rv[c], rv[c]);  // SASS uses the scalar-broadcast operand form
  const f32x2 ntau2 = pack2(-tau2f, -tau2f);
  const f32x2 m4 = pack2(-4.0f, -4.0f);
  const f32x2 kth = pack2(2.86102294921875e-06f, 2.86102294921875e-06f);  // 3 * 2^-20 = 48u
  mbar_wait(&bar, 0);

  // Pad rows (i >= N) and pad columns (j >= N) hold NaN and give zero bits through the literal path; that path
  // costs ~3x a filtered group, and with N = 5000 in Npad = 5120 the 30 pad groups of the last column tile
  // were 6 % of the kernel's instructions.  They are skipped instead: same (zero) bits.
  const int ncv = min(128, d.N - J0);        // valid columns of this tile (> 0)
  const bool row_valid = I0 + r < d.N;
  // One word = 32 columns = eight groups of four (one 128-bit broadcast load per staged array and group).  The word
  // is assembled by shifting one sign bit in per test (one funnel shift each): groups from the top of the word
  // down, and the four columns of a group from the last to the first, so that every bit ends up at its column's
  // position.  A group the filter is not sure about only leaves a flag; the flagged groups (rare: inside the
  // rounding band, out-of-range magnitudes, or NaN pad columns) are re-evaluated literally after the word.  The
  // loop over the words stays rolled and the eight groups of a full word are straight-line code without any
  // per-group test: the kernel is bound by instruction issue, and the per-group loop overhead (validity tests,
  // re-derived shared-memory addresses, the branch around the fallback) was 3 of 17 instructions per pair test
  // (2.64 -> 2.43 ms per 256-pair step).  Measured slower (profiles/experiments/README.md): ONE predicate per word and
  // the whole word redone group by group when it fails (+9 %; +18 % at KITTI scale) — per test the filter is
  // undecided 2e-4 of the time, but a warp visits 1024 tests per word, so every fifth warp-word took the slow path.
  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll 1
  for (int cw = 0; cw < 4; ++cw) {
    const int c0 = cw * 32;
    const int nv = ncv - c0;  // valid columns of this word (uniform over the CTA)
    uint32_t wbits = 0, unsure = 0;
    // columns c0 + 4 g .. c0 + 4 g + 3: shifts their four sign bits into wbits; `sure` stays true iff the filter is
    // sure about all four
    auto group4 = [&](const int g, bool& sure) {
      const int c = c0 + 4 * g;
      float4 cj[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) cj[k] = *reinterpret_cast<const float4*>(&cs[k][c]);
#pragma unroll
      for (int h = 1; h >= 0; --h) {  // two packed column pairs: columns c+2, c+3 first
        f32x2 col[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) col[k] = h == 0 ? pack2(cj[k].x, cj[k].y) : pack2(cj[k].z, cj[k].w);
        // squared lengths for the filter (fused; within 4u of the specified (a*a + b*b) + c*c)
        const f32x2 x = sum_sq_fused(sub2(ri[0], col[0]), sub2(ri[1], col[1]), sub2(ri[2], col[2]));
        const f32x2 y = sum_sq_fused(sub2(ri[3], col[3]), sub2(ri[4], col[4]), sub2(ri[5], col[5]));
        // exact filter: trust sign(Q) iff |Q| > Theta and S > lo
        const f32x2 S = add2(x, y);
        const f32x2 U = add2(S, ntau2);
        const f32x2 UU = mul2(U, U);
        const f32x2 Q = fma2(mul2(x, y), m4, UU);
        const f32x2 T = mul2(UU, kth);
        float q0, q1, t0, t1, s0, s1;
        unpack2(Q, q0, q1);
        unpack2(T, t0, t1);
        unpack2(S, s0, s1);
        sure = sure && fabsf(q0) > t0 && s0 > lo && fabsf(q1) > t1 && s1 > lo;
        wbits = __funnelshift_l(__float_as_uint(q1), wbits, 1);  // bit = (Q < 0) when sure
        wbits = __funnelshift_l(__float_as_uint(q0), wbits, 1);
      }
    };
    if (row_valid && nv > 0) {
      if (nv >= 32) {
#pragma unroll
        for (int g = 7; g >= 0; --g) {
          bool sure = true;
          group4(g, sure);
          if (!sure) unsure |= 1u << g;
        }
      } else {  // last column tile of the pair: only the groups that hold a valid column
#pragma unroll 1
        for (int g = (nv - 1) >> 2; g >= 0; --g) {
          bool sure = true;
          group4(g, sure);
          if (!sure) unsure |= 1u << g;
        }
      }
      while (unsure) {
        const int g = __ffs(static_cast<int>(unsure)) - 1;
        unsure &= unsure - 1u;
        const uint32_t lit = compat_literal4(cs, c0 + 4 * g, rv[0], rv[1], rv[2], rv[3], rv[4], rv[5], tau);
        wbits = (wbits & ~(0xFu << (4 * g))) | (lit << (4 * g));
      }
    }
    if (cw == 0) w0 = wbits;
    else if (cw == 1) w1 = wbits;
    else if (cw == 2) w2 = wbits;
    else w3 = wbits;
  }
  uint32_t words[4] = {w0, w1, w2, w3};

  uint32_t* adjp = adj + d.adj_off;
  unsigned int cnt = 0;
  if (I == J) {
    // A_ii = 0; count only j > i for the edge total
#pragma unroll
    for (int cw = 0; cw < 4; ++cw) {
      if (cw == (r >> 5)) words[cw] &= ~(1u << (r & 31));
      uint32_t upper;  // bits with column index > r
      if (cw > (r >> 5)) upper = 0xffffffffu;
      else if (cw < (r >> 5)) upper = 0u;
      else upper = (r & 31) == 31 ? 0u : (0xffffffffu << ((r & 31) + 1));
      cnt += __popc(words[cw] & upper);
    }
  } else {
#pragma unroll
    for (int cw = 0; cw < 4; ++cw) cnt += __popc(words[cw]);
  }
  *reinterpret_cast<uint4*>(adjp + static_cast<size_t>(I0 + r) * d.stride + J0 / 32) =
      make_uint4(words[0], words[1], words[2], words[3]);
  // K-panel copy for the tensor-core triangle kernel: panel p = 256 columns, [Npad rows][8 words], so a
  // block of rows of one panel is contiguous (one bulk copy per operand block and K stage)
  uint32_t* pp = panel != nullptr ? panel + d.panel_off : nullptr;
  if (pp != nullptr) {
    uint32_t* prow = pp + (static_cast<size_t>(J >> 1) * d.Npad + I0 + r) * 8;
    *reinterpret_cast<uint4*>(prow + (J & 1) * 4) = make_uint4(words[0], words[1], words[2], words[3]);
    if (J == d.nblk - 1 && (J & 1) == 0) *reinterpret_cast<uint4*>(prow + 4) = make_uint4(0u, 0u, 0u, 0u);  // half panel
  }

  if (I != J) {
    // mirrored tile: word for row J0+c, column word I0/32+warp, bit lane = A[I0+32*warp+lane][J0+c]
#pragma unroll
    for (int cw = 0; cw < 4; ++cw) tsm[cw * 32 + lane][warp] = transpose32(words[cw], lane);
    __syncthreads();
    const uint4 tw = *reinterpret_cast<const uint4*>(&tsm[r][0]);
    *reinterpret_cast<uint4*>(adjp + static_cast<size_t>(J0 + r) * d.stride + I0 / 32) = tw;
    if (pp != nullptr) *reinterpret_cast<uint4*>(pp + (static_cast<size_t>(I >> 1) * d.Npad + J0 + r) * 8 + (I & 1) * 4) = tw;
  }

  // edges (i<j) of this tile, added to the count of its triangle work unit (J, I/2): the unit
  // scan turns these into exact key-pool offsets, so the triangle kernel needs no reservation
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) wcnt[warp] = cnt;
  __syncthreads();
  if (r == 0) {
    const unsigned int total = wcnt[0] + wcnt[1] + wcnt[2] + wcnt[3];
    if (total)
      atomicAdd(&ucount[static_cast<size_t>(blockIdx.y) * unit_pitch + unit_offset(static_cast<unsigned int>(J)) + (I >> 1)],
                total);
  }
}

int launch_graph(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_nblk, const float* d_soa,
                 uint32_t* d_adj, uint32_t* d_panel, uint32_t* d_ucount, int unit_pitch, float tau) {
  dim3 grid(max_nblk * (max_nblk + 1) / 2, pairs);
  const float tau2f = tau * tau;  // one fp32 multiply
  float lo = 4.0f * tau2f;        // exact scaling (or +inf: then every pair takes the literal path)
  if (!(lo >= 8.8817841970012523e-16f)) lo = 8.8817841970012523e-16f;  // 2^-50; also replaces NaN
  graph_kernel<<<grid, 128, 0, lc.stream>>>(d_desc, d_soa, d_adj, d_panel, d_ucount, unit_pitch, tau, tau2f, lo);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

// ------------------------------------------------------------------------------------------
// Per pair: exclusive scan of the unit edge counts over the units this rank owns (all of them
// unless sharded) -> ubase[unit] = offset of the unit's keys inside the pair's key slice, and
// the pair's evaluated-edge total.  One CTA per pair.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) unit_scan_kernel(const PairDesc* __restrict__ descs,
                                                         PairDev* __restrict__ state,
                                                         const uint32_t* __restrict__ ucount,
                                                         uint32_t* __restrict__ ubase, int unit_pitch, int rank,
                                                         int world) {
  const int pair = blockIdx.x;
  const unsigned int U = unit_count(static_cast<unsigned int>(descs[pair].nblk));
  const uint32_t* cnt = ucount + static_cast<size_t>(pair) * unit_pitch;
  uint32_t* out = ubase + static_cast<size_t>(pair) * unit_pitch;
  __shared__ unsigned int wsum[32];
  __shared__ unsigned long long carry, all_edges;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) {
    carry = 0;
    all_edges = 0;
  }
  __syncthreads();
  {  // edges of the whole pair, whoever owns them
    unsigned long long a = 0;
    for (unsigned int u = t; u < U; u += 1024) a += cnt[u];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0 && a) atomicAdd(&all_edges, a);
  }
  for (unsigned int u0 = 0; u0 < U; u0 += 1024) {
    const unsigned int u = u0 + t;
    const bool owned = u < U && (world <= 1 || owner_of_unit(u, static_cast<unsigned int>(world)) == static_cast<unsigned int>(rank));
    const unsigned int v = owned ? cnt[u] : 0u;
    unsigned int incl = v;  // warp inclusive scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned int w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      wsum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long base = carry + (warp ? wsum[warp - 1] : 0u);
    if (u < U) out[u] = static_cast<uint32_t>(base + incl - v);
    __syncthreads();
    if (t == 1023) carry = base + incl;
    __syncthreads();
  }
  if (t == 0) {
    state[pair].num_edges = carry;
    state[pair].all_edges = all_edges;
    state[pair].key_count = carry;  // the key scan zeroes it again if the tensor-core kernel is chosen
  }
}

int launch_unit_scan(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, PairDev* d_state,
                     const uint32_t* d_ucount, uint32_t* d_ubase, int unit_pitch, int rank, int world) {
  unit_scan_kernel<<<pairs, 1024, 0, lc.stream>>>(d_desc, d_state, d_ucount, d_ubase, unit_pitch, rank, world);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

// ------------------------------------------------------------------------------------------
// Key-pool layout for the chunk: exclusive scan of the per-pair edge totals.  If the pool is
// too small the overflow flag is raised; every later kernel of the chunk then returns at once
// and the host grows the pool and re-runs the chunk.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) key_scan_kernel(const PairDesc* __restrict__ descs, int pairs,
                                                        PairDev* __restrict__ state, ChunkDev* __restrict__ chunk,
                                                        StickyDev* __restrict__ sticky, unsigned long long key_cap,
                                                        int tri_mode, const ChunkDev* __restrict__ prev) {
  __shared__ unsigned long long part[1024];
  __shared__ unsigned long long carry;
  __shared__ unsigned long long node_pairs;  // sum of N (N - 1) / 2 over the chunk
  __shared__ unsigned long long all_total;   // sum of the pairs' edge counts (owned or not)
  __shared__ int max_n;
  __shared__ uint32_t s_tensor;
  const int t = threadIdx.x;
  if (t == 0) {
    carry = 0;
    node_pairs = 0;
    all_total = 0;
    max_n = 0;
  }
  __syncthreads();
  if (tri_mode == 2) {
    unsigned long long np = 0, ae = 0;
    int mn = 0;
    for (int b = t; b < pairs; b += 1024) {
      const unsigned long long n = static_cast<unsigned long long>(descs[b].N);
      np += n * (n - 1) / 2;
      mn = max(mn, descs[b].N);
      ae += state[b].all_edges;
    }
    if (np) atomicAdd(&node_pairs, np);
    if (ae) atomicAdd(&all_total, ae);
    if (mn) atomicMax(&max_n, mn);
  }
  for (int b0 = 0; b0 < pairs; b0 += 1024) {
    const int b = b0 + t;
    const unsigned long long e = b < pairs ? state[b].num_edges : 0ull;
    part[t] = e;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const unsigned long long add = t >= o ? part[t - o] : 0ull;
      __syncthreads();
      part[t] += add;
      __syncthreads();
    }
    if (b < pairs) state[b].key_base = carry + part[t] - e;
    __syncthreads();
    if (t == 1023) carry += part[1023];
    __syncthreads();
  }
  if (t == 0) {
    // S2 path of the chunk: the dense tensor-core kernel costs ~N^3, the POPC kernels ~E N / 32
    uint32_t tensor = tri_mode == 1 ? 1u : 0u;
    if (tri_mode == 2)
      tensor = (max_n >= kTensorMinN && static_cast<float>(all_total) >= kTensorMinDensity * static_cast<float>(node_pairs)) ? 1u : 0u;  // density of the whole graph: the same decision on every rank of a sharded run
    s_tensor = tensor;
    chunk->use_tensor = tensor;
    chunk->total_edges = carry;
    chunk->overflow = carry > key_cap ? 1u : 0u;
    if (carry > key_cap) {
      atomicMax(&sticky->max_total_edges, carry);
      atomicAdd(&sticky->overflow_count, 1u);
    }
    // second pass of the second-order mode: an overflow of the first pass (already recorded) voids this one too
    if (prev && prev->overflow) {
      chunk->overflow = 1u;
      chunk->total_edges = prev->total_edges > carry ? prev->total_edges : carry;
    }
  }
  __syncthreads();
  // the tensor-core kernel appends only the keys that pass its pruning threshold and counts them itself
  if (s_tensor)
    for (int b = t; b < pairs; b += 1024) state[b].key_count = 0ull;
}

int launch_key_scan(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, PairDev* d_state, ChunkDev* d_chunk,
                    StickyDev* d_sticky, unsigned long long key_cap, int tri_mode, const ChunkDev* d_prev) {
  key_scan_kernel<<<1, 1024, 0, lc.stream>>>(d_desc, pairs, d_state, d_chunk, d_sticky, key_cap, tri_mode, d_prev);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

// ------------------------------------------------------------------------------------------
// Second-order compatibility: A2_ij = A_ij and C_ij >= cmin with C_ij = popc(row_i(A) & row_j(A)),
// which is exactly the per-edge count T_ij the triangle kernels leave in the key list of a first pass over A.
// This kernel turns that key list into A2 (both orientations, plus the K-panel copy the tensor-core kernel reads)
// and into the per-unit edge counts of A2 the second pass starts from.  Bits are set with atomicOr: the result does
// not depend on the (unordered) key list.  adj2 / panel2 / ucount arrive zeroed.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) second_order_scatter_kernel(
    const PairDesc* __restrict__ descs, const PairDev* __restrict__ state1, const ChunkDev* __restrict__ chunk1,
    const unsigned long long* __restrict__ keys, uint32_t cmin, uint32_t* __restrict__ adj2, uint32_t* __restrict__ panel2,
    uint32_t* __restrict__ ucount, int unit_pitch) {
  if (chunk1->overflow) return;
  const int pair = blockIdx.y;
  const PairDesc d = descs[pair];
  const unsigned long long n = state1[pair].key_count;
  const unsigned long long* kp = keys + state1[pair].key_base;
  uint32_t* a2 = adj2 + d.adj_off;
  uint32_t* p2 = panel2 ? panel2 + d.panel_off : nullptr;
  uint32_t* uc = ucount + static_cast<size_t>(pair) * unit_pitch;
  for (unsigned long long idx = static_cast<unsigned long long>(blockIdx.x) * 256 + threadIdx.x; idx < n;
       idx += static_cast<unsigned long long>(gridDim.x) * 256) {
    const unsigned long long key = kp[idx];
    if (static_cast<uint32_t>(key >> 32) < cmin) continue;
    const unsigned int i = 0xFFFFu - static_cast<unsigned int>((key >> 16) & 0xFFFFu);
    const unsigned int j = 0xFFFFu - static_cast<unsigned int>(key & 0xFFFFu);
    atomicOr(&a2[static_cast<size_t>(i) * d.stride + (j >> 5)], 1u << (j & 31));
    atomicOr(&a2[static_cast<size_t>(j) * d.stride + (i >> 5)], 1u << (i & 31));
    if (p2) {  // panel p = columns [256 p, 256 p + 256): [Npad rows][8 words]
      atomicOr(&p2[(static_cast<size_t>(j >> 8) * d.Npad + i) * 8 + ((j & 255u) >> 5)], 1u << (j & 31));
      atomicOr(&p2[(static_cast<size_t>(i >> 8) * d.Npad + j) * 8 + ((i & 255u) >> 5)], 1u << (i & 31));
    }
    atomicAdd(&uc[unit_offset(j >> 7) + (i >> 8)], 1u);  // unit = 128 columns x 256 rows, i < j
  }
}

int launch_second_order_scatter(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, const PairDev* d_state1,
                                const ChunkDev* d_chunk1, const unsigned long long* d_keys, uint32_t cmin,
                                uint32_t* d_adj2, uint32_t* d_panel2, uint32_t* d_ucount, int unit_pitch) {
  int gx = (8 * lc.sm_count + pairs - 1) / pairs;
  if (gx < 8) gx = 8;
  second_order_scatter_kernel<<<dim3(gx, pairs), 256, 0, lc.stream>>>(d_desc, d_state1, d_chunk1, d_keys, cmin, d_adj2,
                                                                      d_panel2, d_ucount, unit_pitch);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, uint32_t v, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) p[k] = v;
}
int launch_fill_u32(const LaunchCtx& lc, uint32_t* d_p, uint32_t v, int n) {
  fill_u32_kernel<<<(n + 255) / 256, 256, 0, lc.stream>>>(d_p, v, n);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

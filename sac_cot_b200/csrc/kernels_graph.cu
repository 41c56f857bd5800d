// kernels_graph.cu — S0 (AoS -> padded SoA repack) and S1 (first-order length-consistency
// compatibility graph), plus the tiny per-chunk key-pool scan.
//
// S1 arithmetic contract (bit-exact against the oracle): every fp32 operation is an
// individually rounded IEEE op issued through __f*_rn intrinsics, which the compiler never
// contracts into FMAs (equivalent to -fmad=false on this kernel), in the order
//   a=sx_i-sx_j; b=..; c=..; s2=(a*a+b*b)+c*c; ls=sqrt(s2); (same for dst -> ld);
//   A_ij = |ls-ld| < tau_c.
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
// (TMA, completion on an mbarrier) and read as warp-wide broadcasts.  Each thread builds the
// four 32-bit words of its row with one predicate per pair and stores them as one 128-bit
// word; the mirrored tile is produced with warp ballots, staged through shared memory, and
// stored the same way.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool compat_pred(float sxi, float syi, float szi, float dxi, float dyi, float dzi,
                                            float sxj, float syj, float szj, float dxj, float dyj, float dzj,
                                            float tau) {
  const float a = __fsub_rn(sxi, sxj);
  const float b = __fsub_rn(syi, syj);
  const float c = __fsub_rn(szi, szj);
  const float s2 = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
  const float ls = __fsqrt_rn(s2);
  const float u = __fsub_rn(dxi, dxj);
  const float v = __fsub_rn(dyi, dyj);
  const float w = __fsub_rn(dzi, dzj);
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)), __fmul_rn(w, w));
  const float ld = __fsqrt_rn(d2);
  return fabsf(__fsub_rn(ls, ld)) < tau;
}

__global__ void __launch_bounds__(128) graph_kernel(const PairDesc* __restrict__ descs,
                                                    const float* __restrict__ soa, uint32_t* __restrict__ adj,
                                                    uint32_t* __restrict__ ucount, int unit_pitch, float tau) {
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
  const float sxi = base[0 * static_cast<size_t>(d.Npad) + I0 + r];
  const float syi = base[1 * static_cast<size_t>(d.Npad) + I0 + r];
  const float szi = base[2 * static_cast<size_t>(d.Npad) + I0 + r];
  const float dxi = base[3 * static_cast<size_t>(d.Npad) + I0 + r];
  const float dyi = base[4 * static_cast<size_t>(d.Npad) + I0 + r];
  const float dzi = base[5 * static_cast<size_t>(d.Npad) + I0 + r];
  mbar_wait(&bar, 0);

  uint32_t words[4];
#pragma unroll
  for (int cw = 0; cw < 4; ++cw) {
    uint32_t wbits = 0;
#pragma unroll 8
    for (int b = 0; b < 32; ++b) {
      const int c = cw * 32 + b;
      const bool p = compat_pred(sxi, syi, szi, dxi, dyi, dzi, cs[0][c], cs[1][c], cs[2][c], cs[3][c], cs[4][c],
                                 cs[5][c], tau);
      wbits |= (p ? 1u : 0u) << b;
    }
    words[cw] = wbits;
  }

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

  if (I != J) {
    // mirrored tile: word for row J0+c, column word I0/32+warp, bit lane = A[I0+32*warp+lane][J0+c]
#pragma unroll
    for (int cw = 0; cw < 4; ++cw) {
      uint32_t mine = 0;
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        const uint32_t v = __ballot_sync(0xffffffffu, (words[cw] >> b) & 1u);
        if (lane == b) mine = v;
      }
      tsm[cw * 32 + lane][warp] = mine;
    }
    __syncthreads();
    const uint4 tw = *reinterpret_cast<const uint4*>(&tsm[r][0]);
    *reinterpret_cast<uint4*>(adjp + static_cast<size_t>(J0 + r) * d.stride + I0 / 32) = tw;
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
                 uint32_t* d_adj, uint32_t* d_ucount, int unit_pitch, float tau) {
  dim3 grid(max_nblk * (max_nblk + 1) / 2, pairs);
  graph_kernel<<<grid, 128, 0, lc.stream>>>(d_desc, d_soa, d_adj, d_ucount, unit_pitch, tau);
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
  __shared__ unsigned long long carry;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) carry = 0;
  __syncthreads();
  for (unsigned int u0 = 0; u0 < U; u0 += 1024) {
    const unsigned int u = u0 + t;
    const bool owned = u < U && (world <= 1 || (u % static_cast<unsigned int>(world)) == static_cast<unsigned int>(rank));
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
    state[pair].key_count = carry;
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
__global__ void __launch_bounds__(1024) key_scan_kernel(int pairs, PairDev* __restrict__ state,
                                                        ChunkDev* __restrict__ chunk, StickyDev* __restrict__ sticky,
                                                        unsigned long long key_cap) {
  __shared__ unsigned long long part[1024];
  __shared__ unsigned long long carry;
  const int t = threadIdx.x;
  if (t == 0) carry = 0;
  __syncthreads();
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
    chunk->total_edges = carry;
    chunk->overflow = carry > key_cap ? 1u : 0u;
    if (carry > key_cap) {
      atomicMax(&sticky->max_total_edges, carry);
      atomicAdd(&sticky->overflow_count, 1u);
    }
  }
}

int launch_key_scan(const LaunchCtx& lc, int pairs, PairDev* d_state, ChunkDev* d_chunk, StickyDev* d_sticky,
                    unsigned long long key_cap) {
  key_scan_kernel<<<1, 1024, 0, lc.stream>>>(pairs, d_state, d_chunk, d_sticky, key_cap);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

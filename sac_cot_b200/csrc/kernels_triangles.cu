// kernels_triangles.cu — S2, compatibility-triangle counts on the POPC bitset path
// (SURVEY.md §8a row S2):  T_ij = popc(row_i & row_j) for every edge i<j, and the per-node sums
// t2_i = sum_j A_ij T_ij (= 2 t_i).  Integer, exact, order free.
//
// Work unit = (J-block of 128 adjacency rows, chunk of 256 rows i) with i-chunk <= J-block, the
// same numbering the oracle uses to assign edges to ranks in sharded mode.  The 128 J rows are
// staged in shared memory by bulk copies (TMA, one per row, completion on an mbarrier).  Each
// warp walks rows i of the chunk: the lanes hold row i's words in registers (word lane+32k),
// the edge bits A[i][J-block] say which staged rows to intersect, and every intersection is
// R x (LDS + AND + POPC) per lane followed by one warp REDUX.  Rows longer than 352 words
// (N > 11264) are processed in 256-word chunks with per-edge u16 accumulators in shared
// memory.  Per edge the kernel emits a 64-bit key  T<<32 | (0xFFFF-i)<<16 | (0xFFFF-j)  into
// the pair's slice of the key pool and updates a 4096-bin shared histogram of T>>4 that is
// flushed once per unit (input of the top-K_e edge selection).
#include "common.cuh"

namespace saccot {

// unit id -> (jb, ic);  offset(jb) = sum_{b<jb} (b/2+1) = h(h+1) for jb=2h, (h+1)^2 for jb=2h+1
__device__ __forceinline__ unsigned int unit_offset(unsigned int jb) {
  const unsigned int h = jb >> 1;
  return (jb & 1u) ? (h + 1) * (h + 1) : h * (h + 1);
}
__host__ __device__ inline unsigned int unit_count(unsigned int nblk) {
  const unsigned int h = nblk >> 1;
  return (nblk & 1u) ? (h + 1) * (h + 1) : h * (h + 1);
}

template <int R>
__global__ void __launch_bounds__(kTriThreads) triangles_kernel(
    const PairDesc* __restrict__ descs, const uint32_t* __restrict__ adj, PairDev* __restrict__ state,
    const ChunkDev* __restrict__ chunk, unsigned long long* __restrict__ keys, uint32_t* __restrict__ hist,
    unsigned long long* __restrict__ t2, int nchunks, int rank, int world) {
  if (chunk->overflow) return;
  const int pair = blockIdx.y;
  const PairDesc d = descs[pair];
  const unsigned int unit = blockIdx.x;
  if (unit >= unit_count(static_cast<unsigned int>(d.nblk))) return;
  if (world > 1 && (unit % static_cast<unsigned int>(world)) != static_cast<unsigned int>(rank)) return;
  // decode the unit
  unsigned int jb = static_cast<unsigned int>(2.0f * sqrtf(static_cast<float>(unit)));
  if (jb >= static_cast<unsigned int>(d.nblk)) jb = d.nblk - 1;
  while (unit_offset(jb) > unit) --jb;
  while (unit_offset(jb + 1) <= unit) ++jb;
  const unsigned int ic = unit - unit_offset(jb);
  const int J0 = static_cast<int>(jb) * kTriJ;
  const int I0 = static_cast<int>(ic) * kTriI;

  constexpr int PITCH = 32 * R;  // words per staged row (bank-conflict free: bank == lane)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw);                     // [128][PITCH]
  uint32_t* hist_s = rows + kTriJ * PITCH;                                    // [4096]
  uint32_t* tJ = hist_s + kHistBins;                                          // [128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(tJ + kTriJ);                    // 8 bytes (8-aligned)
  uint16_t* acc = reinterpret_cast<uint16_t*>(bar + 2);                       // [256][128], nchunks > 1 only

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = kTriThreads / 32;
  for (int k = tid; k < kHistBins; k += kTriThreads) hist_s[k] = 0;
  if (tid < kTriJ) tJ[tid] = 0;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t* adjp = adj + d.adj_off;
  const int stride = d.stride;
  unsigned long long* keyp = keys + state[pair].key_base;
  uint32_t phase = 0;
  // `nchunks` (launch-wide, from the longest row in the launch) selects the accumulate-then-emit
  // mode; this pair's own rows may need fewer chunks.
  const int my_chunks = (stride + PITCH - 1) / PITCH;

  for (int c = 0; c < my_chunks; ++c) {
    const int c0 = c * PITCH;
    const int wc = min(PITCH, stride - c0);  // words of this chunk (multiple of 4)
    if (warp == 0) {
      if (lane == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(kTriJ * wc * 4));
      __syncwarp();
      for (int rr = lane; rr < kTriJ; rr += 32)
        bulk_g2s(rows + rr * PITCH, adjp + static_cast<size_t>(J0 + rr) * stride + c0, static_cast<uint32_t>(wc * 4), bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;

    for (int rr = warp; rr < kTriI; rr += NWARP) {
      const int i = I0 + rr;
      if (i >= d.N || i >= J0 + kTriJ - 1) break;  // rows are visited in increasing i per warp
      // edge bits A[i][J0 .. J0+127] restricted to j > i
      const uint4 e4 = *reinterpret_cast<const uint4*>(adjp + static_cast<size_t>(i) * stride + jb * 4);
      uint32_t eb[4] = {e4.x, e4.y, e4.z, e4.w};
      if (i >= J0) {
        const int li = i - J0;  // clear bits <= li
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          if (w < (li >> 5)) eb[w] = 0;
          else if (w == (li >> 5)) eb[w] &= (li & 31) == 31 ? 0u : (0xffffffffu << ((li & 31) + 1));
        }
      }
      const int n = __popc(eb[0]) + __popc(eb[1]) + __popc(eb[2]) + __popc(eb[3]);
      if (n == 0) continue;
      uint32_t ri[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int idx = lane + 32 * k;
        ri[k] = idx < wc ? adjp[static_cast<size_t>(i) * stride + c0 + idx] : 0u;
      }
      unsigned long long base = 0;
      if (nchunks == 1) {
        if (lane == 0) base = atomicAdd(&state[pair].key_count, static_cast<unsigned long long>(n));
        base = __shfl_sync(0xffffffffu, base, 0);
      }
      unsigned int q = 0;
      unsigned long long tsum = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        uint32_t bits = eb[w];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int jl = w * 32 + b;
          const uint32_t* rj = rows + jl * PITCH + lane;
          int s = 0;
#pragma unroll
          for (int k = 0; k < R; ++k) s += __popc(ri[k] & rj[32 * k]);
          const unsigned int T = static_cast<unsigned int>(__reduce_add_sync(0xffffffffu, s));
          if (nchunks == 1) {
            if (lane == 0) {
              keyp[base + q] = (static_cast<unsigned long long>(T) << 32) |
                               (static_cast<unsigned long long>(0xFFFFu - static_cast<unsigned int>(i)) << 16) |
                               static_cast<unsigned long long>(0xFFFFu - static_cast<unsigned int>(J0 + jl));
              atomicAdd(&hist_s[T >> 4], 1u);
              atomicAdd(&tJ[jl], T);
            }
            tsum += T;
            ++q;
          } else if (lane == 0) {
            uint16_t* a = acc + rr * kTriJ + jl;
            *a = static_cast<uint16_t>(c == 0 ? T : static_cast<unsigned int>(*a) + T);
          }
        }
      }
      if (nchunks == 1 && lane == 0) atomicAdd(&t2[d.node_off + i], tsum);
    }
    __syncthreads();  // everyone is done with `rows` before the next chunk overwrites it
  }

  if (nchunks > 1) {
    // emission pass from the accumulated per-edge counts
    for (int rr = warp; rr < kTriI; rr += NWARP) {
      const int i = I0 + rr;
      if (i >= d.N || i >= J0 + kTriJ - 1) break;
      const uint4 e4 = *reinterpret_cast<const uint4*>(adjp + static_cast<size_t>(i) * stride + jb * 4);
      uint32_t eb[4] = {e4.x, e4.y, e4.z, e4.w};
      if (i >= J0) {
        const int li = i - J0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          if (w < (li >> 5)) eb[w] = 0;
          else if (w == (li >> 5)) eb[w] &= (li & 31) == 31 ? 0u : (0xffffffffu << ((li & 31) + 1));
        }
      }
      const int n = __popc(eb[0]) + __popc(eb[1]) + __popc(eb[2]) + __popc(eb[3]);
      if (n == 0) continue;
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&state[pair].key_count, static_cast<unsigned long long>(n));
      base = __shfl_sync(0xffffffffu, base, 0);
      // lanes take the set bits round-robin: ordinal o of bit (w,b) = popc of lower set bits
      unsigned long long tsum = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t bits = eb[w];
        const int before = (w > 0 ? __popc(eb[0]) : 0) + (w > 1 ? __popc(eb[1]) : 0) + (w > 2 ? __popc(eb[2]) : 0);
        if ((bits >> lane) & 1u) {
          const int jl = w * 32 + lane;
          const unsigned int o = before + __popc(bits & ((1u << lane) - 1u));
          const unsigned int T = acc[rr * kTriJ + jl];
          keyp[base + o] = (static_cast<unsigned long long>(T) << 32) |
                           (static_cast<unsigned long long>(0xFFFFu - static_cast<unsigned int>(i)) << 16) |
                           static_cast<unsigned long long>(0xFFFFu - static_cast<unsigned int>(J0 + jl));
          atomicAdd(&hist_s[T >> 4], 1u);
          atomicAdd(&tJ[jl], T);
          tsum += T;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
      if (lane == 0) atomicAdd(&t2[d.node_off + i], tsum);
    }
    __syncthreads();
  }

  // flush the unit's histogram and J-side node sums
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

static size_t tri_smem_bytes(int R, bool chunked) {
  size_t b = static_cast<size_t>(kTriJ) * 32 * R * 4 + kHistBins * 4 + kTriJ * 4 + 16;
  if (chunked) b += static_cast<size_t>(kTriI) * kTriJ * 2;
  return b;
}

template <int R>
static cudaError_t tri_set_attr() {
  return cudaFuncSetAttribute(triangles_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              static_cast<int>(tri_smem_bytes(R, R == kTriChunkR)));
}

int triangles_configure() {
  cudaError_t e;
#define SACCOT_SET(R) if ((e = tri_set_attr<R>()) != cudaSuccess) return -static_cast<int>(e);
  SACCOT_SET(1) SACCOT_SET(2) SACCOT_SET(3) SACCOT_SET(4) SACCOT_SET(5) SACCOT_SET(6)
  SACCOT_SET(7) SACCOT_SET(8) SACCOT_SET(9) SACCOT_SET(10) SACCOT_SET(11)
#undef SACCOT_SET
  return 0;
}

int launch_triangles(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_nblk, int max_stride,
                     const uint32_t* d_adj, PairDev* d_state, const ChunkDev* d_chunk, unsigned long long* d_keys,
                     uint32_t* d_hist, unsigned long long* d_t2, int rank, int world) {
  // R = words per lane; rows of up to 32*kTriMaxR words are staged whole, longer rows in chunks
  int R = (max_stride + 31) / 32;
  int nchunks = 1;
  if (R > kTriMaxR) {
    R = kTriChunkR;
    nchunks = (max_stride + 32 * R - 1) / (32 * R);
  }
  dim3 grid(unit_count(static_cast<unsigned int>(max_nblk)), pairs);
  const size_t smem = tri_smem_bytes(R, nchunks > 1);
#define SACCOT_LAUNCH(RR)                                                                                      \
  case RR:                                                                                                     \
    triangles_kernel<RR><<<grid, kTriThreads, smem, lc.stream>>>(d_desc, d_adj, d_state, d_chunk, d_keys,     \
                                                                 d_hist, d_t2, nchunks, rank, world);          \
    break;
  switch (R) {
    SACCOT_LAUNCH(1) SACCOT_LAUNCH(2) SACCOT_LAUNCH(3) SACCOT_LAUNCH(4) SACCOT_LAUNCH(5) SACCOT_LAUNCH(6)
    SACCOT_LAUNCH(7) SACCOT_LAUNCH(8) SACCOT_LAUNCH(9) SACCOT_LAUNCH(10) SACCOT_LAUNCH(11)
    default: return -static_cast<int>(cudaErrorInvalidValue);
  }
#undef SACCOT_LAUNCH
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

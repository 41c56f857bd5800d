// kernels_triangles_mma.cu — S2 on the tensor cores (K2b of SURVEY.md §2b): the dense contraction
// T = (A·Aᵀ) ∘ A with tcgen05.mma, accumulators in TMEM.  Integer-exact: operands are 0/1.
//
// north_star names int8 tiles; the probe in profiles/microbench/umma_probe.cu measured one SM at
// 7.9 k MAC/clk for kind::i8 and 15.6 k MAC/clk for kind::mxf4.block_scale, so the kernel uses
// the 4-bit path: bit b becomes the e2m1 nibble 0b0010 (1.0) or 0, every UE8M0 scale factor is
// 0x7F (1.0) — one TMEM region filled once and shared by all MMAs — and the fp32 accumulator holds
// exact integers (counts < 2^24).
//
// Feeding the tensor core is the problem, not the MMA: pre-expanded operands would need > 400 MB
// of L2->SMEM traffic per pair (L2-bound, slower than the POPC kernel), so the 1-bit adjacency rows
// are expanded ON CHIP: producer warps read 32-bit words of the rows (L2), spread them to 16 bytes of
// nibbles with a byte-permute LUT, and write them straight into the no-swizzle K-major canonical
// layout the UMMA descriptors describe (core matrix = 8 rows x 16 B).
//
// One CTA per work item = (J-block of 224 columns, 256 rows = two 128-row A blocks).  TMEM: two
// 128 x 224 fp32 accumulators (448 columns) + the scale-factor region (64 columns).  Warp roles:
//   warps 0-3   epilogue: while the MMAs run they read the edge bits A[i][J-block], reserve key
//               ranges (one warp-aggregated atomic per 32-column chunk) and then, when the
//               accumulators are complete, tcgen05.ld them, keep T_ij where A_ij = 1 and j > i,
//               emit keys / histogram / node sums exactly like the POPC kernel
//   warp  4     MMA issuer (one lane): per stage 2 A-blocks x 4 k-steps of M=128, N=224, K=64
//   warps 5-15  producers: expansion of 480 rows x 256 K-elements per stage, 3-stage ring
#include "common.cuh"

#include <algorithm>

namespace saccot {

constexpr int kMmaThreads = 512;
constexpr int kMmaJ = 224;                 // columns per item (7 words)
constexpr int kMmaI = 256;                 // rows per item (two A blocks of 128)
constexpr int kMmaStageK = 256;            // K elements per stage (8 words, 128 bytes of nibbles per row)
constexpr int kMmaStages = 3;
constexpr int kMmaRows = kMmaI + kMmaJ;    // 480 rows expanded per stage
constexpr int kMmaLBO = 128;               // next 16-byte K chunk (core matrices contiguous along K)
constexpr int kMmaSBO = (kMmaStageK / 2 / 16) * 128;  // next 8-row group: 8 core matrices = 1024 B
constexpr int kMmaStageBytes = (kMmaRows / 8) * kMmaSBO;  // 60 groups x 1024 B = 61440
constexpr int kMmaProducerWarps = 11;
constexpr uint32_t kSfCol = 448;           // scale-factor region: TMEM columns [448, 512)

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((kMmaLBO >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((kMmaSBO >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell); SWIZZLE_NONE, base offset 0
  return d;
}

// 16 adjacency bits -> 16 e2m1 nibbles (two 32-bit words): spread the 2-bit groups into nibbles,
// then a 4-entry byte LUT {00,01,10,11} -> {0x00,0x02,0x20,0x22} through PRMT.
__device__ __forceinline__ void expand16(uint32_t x16, uint32_t& w0, uint32_t& w1) {
  uint32_t t = __byte_perm(x16, 0u, 0x4140);          // byte0 -> byte0, byte1 -> byte2
  t = (t | (t << 4)) & 0x0F0F0F0Fu;
  t = (t | (t << 2)) & 0x33333333u;                   // nibble k = bits 2k, 2k+1
  const uint32_t lut = 0x22200200u;
  w0 = __byte_perm(lut, 0u, t & 0xFFFFu);
  w1 = __byte_perm(lut, 0u, t >> 16);
}

__global__ void __launch_bounds__(kMmaThreads, 1) triangles_mma_kernel(
    const PairDesc* __restrict__ descs, const uint32_t* __restrict__ adj, const PairDev* __restrict__ state,
    const ChunkDev* __restrict__ chunk, unsigned long long* __restrict__ keys, const uint32_t* __restrict__ ubase,
    uint32_t* __restrict__ ucursor, int unit_pitch, uint32_t* __restrict__ hist, unsigned long long* __restrict__ t2,
    int rank, int world) {
  if (chunk->overflow) return;
  const int pair = blockIdx.y;
  const PairDesc d = descs[pair];
  // item -> (J, ip): J-block jq covers columns [224 jq, 224 jq + 224); row pairs ip = 0 .. last(jq)
  // with last(jq) = (224 jq + 222) / 256 clipped to the rows that exist
  const int nJ = (d.Npad + kMmaJ - 1) / kMmaJ;
  const int nIp = (d.Npad + kMmaI - 1) / kMmaI;
  int item = blockIdx.x, jq = nJ - 1, ip = -1;
  for (; jq >= 0; --jq) {  // largest J-blocks (most row pairs) first
    const int cnt = min(nIp, (kMmaJ * jq + kMmaJ - 2) / kMmaI + 1);
    if (item < cnt) { ip = item; break; }
    item -= cnt;
  }
  if (ip < 0) return;
  const int J0 = jq * kMmaJ, I0 = ip * kMmaI;
  const int stride = d.stride;
  const int nstages = (stride + 7) / 8;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* stage_base = smem_raw;                                           // [3][61440]
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(smem_raw + kMmaStages * kMmaStageBytes);  // [4096]
  uint32_t* tJ = hist_s + kHistBins;                                              // [224]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tJ + kMmaJ);                       // full[3], empty[3], acc
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMmaStages;
  uint64_t* acc_full = bars + 2 * kMmaStages;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < kHistBins + kMmaJ; k += kMmaThreads) hist_s[k] = 0;  // hist and tJ are contiguous
  if (tid == 0) {
    for (int s = 0; s < kMmaStages; ++s) {
      mbar_init(&full[s], kMmaProducerWarps);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;
  const uint32_t* adjp = adj + d.adj_off;

  if (warp < 4) {
    // ============================ epilogue warps ============================
    // scale factors: every byte 0x7F (UE8M0 1.0) in columns [448, 512) of this warp's 32 lanes
    {
      const uint32_t v = 0x7F7F7F7Fu;
      const uint32_t taddr = tmem + ((32u * warp) << 16) + kSfCol;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
                taddr + 32u * h),
            "r"(v)
            : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
    }
    // named barrier 1: the four epilogue warps + the MMA warp (160 threads): SF region is ready
    asm volatile("bar.sync 1, 160;" ::: "memory");

    // ---- while the MMAs run: edge bits, counts and key positions of this thread's rows ----
    const int m = 32 * warp + lane;  // row inside an A block
    uint32_t ebits[2][7];
    uint32_t epos[2][7];
    const uint32_t* ubp = ubase + static_cast<size_t>(pair) * unit_pitch;
    uint32_t* ucp = ucursor + static_cast<size_t>(pair) * unit_pitch;
    const unsigned int ic = static_cast<unsigned int>(ip);  // rows [256 ip, 256 ip + 256) => i >> 8 == ip
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int i = I0 + 128 * a + m;
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        const int jc = J0 + 32 * c;  // first column of the chunk
        uint32_t bits = 0;
        if (i < d.N && jc < d.Npad) bits = adjp[static_cast<size_t>(i) * stride + (jc >> 5)];
        // keep j > i
        if (i >= jc + 31) bits = 0;
        else if (i >= jc) bits &= 0xFFFFFFFEu << (i - jc);
        const unsigned int unit = unit_offset(static_cast<unsigned int>(jc) >> 7) + ic;
        if (world > 1 && (unit % static_cast<unsigned int>(world)) != static_cast<unsigned int>(rank)) bits = 0;
        ebits[a][c] = bits;
        // warp-aggregated reservation inside the unit's key region
        const int n = __popc(bits);
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += u;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = 0;
        if (total > 0) {
          if (lane == 31) base = ubp[unit] + atomicAdd(&ucp[unit], static_cast<uint32_t>(total));
          base = __shfl_sync(0xffffffffu, base, 31);
        }
        epos[a][c] = base + static_cast<uint32_t>(incl - n);
      }
    }

    // ---- accumulators complete ----
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    unsigned long long* keyp = keys + state[pair].key_base;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int i = I0 + 128 * a + m;
      const unsigned long long ikey = static_cast<unsigned long long>(0xFFFFu - static_cast<unsigned int>(i)) << 16;
      unsigned int tsum = 0;
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((32u * warp) << 16) + static_cast<uint32_t>(kMmaJ * a + 32 * c);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint32_t bits = ebits[a][c];
        if (bits) {
          uint32_t at = epos[a][c];
          const unsigned int jkey = 0xFFFFu - static_cast<unsigned int>(J0 + 32 * c);
#pragma unroll
          for (int b = 0; b < 32; ++b) {
            if ((bits >> b) & 1u) {
              const unsigned int T = static_cast<unsigned int>(__uint_as_float(v[b]));  // exact integer in fp32
              keyp[at++] = (static_cast<unsigned long long>(T) << 32) | ikey | static_cast<unsigned long long>(jkey - b);
              atomicAdd(&hist_s[T >> 4], 1u);
              atomicAdd(&tJ[32 * c + b], T);
              tsum += T;
            }
          }
        }
      }
      if (tsum) atomicAdd(&t2[d.node_off + i], static_cast<unsigned long long>(tsum));
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
  } else if (warp == 4) {
    // ============================ MMA issuer ============================
    asm volatile("bar.sync 1, 160;" ::: "memory");  // scale-factor region written
    asm volatile("tcgen05.fence::after_thread_sync;");
    // instruction descriptor: block-scaled, A/B = E2M1 (1), UE8M0 scales, N = 224, M = 128, K-major both
    const uint32_t idesc = (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kMmaJ >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    const uint32_t sbase = smem_u32(stage_base);
    for (int it = 0; it < nstages; ++it) {
      const int s = it % kMmaStages;
      mbar_wait(&full[s], static_cast<uint32_t>((it / kMmaStages) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      if (lane == 0) {
        const uint32_t st = sbase + s * kMmaStageBytes;
        const uint32_t bA0 = st, bA1 = st + 16 * kMmaSBO, bB = st + 32 * kMmaSBO;
#pragma unroll
        for (int ks = 0; ks < kMmaStageK / 64; ++ks) {
          const uint64_t db = umma_desc(bB + ks * 2 * kMmaLBO);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const uint64_t da = umma_desc((a == 0 ? bA0 : bA1) + ks * 2 * kMmaLBO);
            const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n}" ::"r"(
                    tmem + static_cast<uint32_t>(kMmaJ * a)),
                "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(tmem + kSfCol), "r"(tmem + kSfCol + 32u)
                : "memory");
          }
        }
        // the stage may be overwritten once these MMAs have read it
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        if (it == nstages - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(acc_full)) : "memory");
      }
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
  } else {
    // ============================ producers ============================
    // Task = (row r of the 480, word j of the 8 in the stage): 32 bits -> 16 bytes of nibbles, one
    // 128-bit store into the canonical layout.  A warp covers 8 rows x 4 words per step: its lanes
    // write 512 contiguous bytes (conflict free) and read 16 contiguous bytes per row.
    const int pw = warp - 5;               // 0..10
    const int r8 = lane & 7, j4 = lane >> 3;
    constexpr int kGroups = (kMmaRows / 8) * 2;  // 120 (row group, word half) pairs per stage
    constexpr int kMaxT = (kGroups + kMmaProducerWarps - 1) / kMmaProducerWarps;  // 11
    const uint32_t* rowp[kMaxT];
    uint32_t soff[kMaxT];
    int jword[kMaxT];
#pragma unroll
    for (int t = 0; t < kMaxT; ++t) {
      const int g = pw + kMmaProducerWarps * t;
      rowp[t] = nullptr;
      soff[t] = 0;
      jword[t] = 0;
      if (g < kGroups) {
        const int rg = g >> 1, jh = g & 1;
        const int r = 8 * rg + r8;
        const int grow = r < kMmaI ? I0 + r : J0 + (r - kMmaI);   // global adjacency row
        jword[t] = 4 * jh + j4;
        soff[t] = static_cast<uint32_t>(rg * kMmaSBO + jword[t] * kMmaLBO + r8 * 16);
        if (grow < d.Npad) rowp[t] = adjp + static_cast<size_t>(grow) * stride;
      }
    }
    uint32_t cur[kMaxT];
    auto load_stage = [&](int it, uint32_t (&w)[kMaxT]) {
#pragma unroll
      for (int t = 0; t < kMaxT; ++t) {
        const int widx = it * 8 + jword[t];
        w[t] = (rowp[t] != nullptr && widx < stride) ? rowp[t][widx] : 0u;
      }
    };
    load_stage(0, cur);
    for (int it = 0; it < nstages; ++it) {
      const int s = it % kMmaStages;
      uint32_t nxt[kMaxT];
      if (it + 1 < nstages) load_stage(it + 1, nxt);
      if (it >= kMmaStages) mbar_wait(&empty[s], static_cast<uint32_t>(((it / kMmaStages) - 1) & 1));
      unsigned char* st = stage_base + s * kMmaStageBytes;
#pragma unroll
      for (int t = 0; t < kMaxT; ++t) {
        if (pw + kMmaProducerWarps * t < kGroups) {
          uint4 o;
          expand16(cur[t] & 0xFFFFu, o.x, o.y);
          expand16(cur[t] >> 16, o.z, o.w);
          *reinterpret_cast<uint4*>(st + soff[t]) = o;
        }
      }
      // generic-proxy writes -> visible to the tensor core (async proxy), then one arrival per warp
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
      if (it + 1 < nstages) {
#pragma unroll
        for (int t = 0; t < kMaxT; ++t) cur[t] = nxt[t];
      }
    }
  }

  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  // flush the item's histogram and J-side node sums
  uint32_t* histp = hist + static_cast<size_t>(pair) * kHistBins;
  for (int k = tid; k < kHistBins; k += kMmaThreads) {
    const uint32_t v = hist_s[k];
    if (v) atomicAdd(&histp[k], v);
  }
  for (int k = tid; k < kMmaJ; k += kMmaThreads) {
    const uint32_t v = tJ[k];
    if (v && J0 + k < d.Npad) atomicAdd(&t2[d.node_off + J0 + k], static_cast<unsigned long long>(v));
  }
}

static size_t mma_smem_bytes() { return static_cast<size_t>(kMmaStages) * kMmaStageBytes + kHistBins * 4 + kMmaJ * 4 + 8 * 8 + 16; }

int triangles_mma_configure() {
  const cudaError_t e = cudaFuncSetAttribute(triangles_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(mma_smem_bytes()));
  return e == cudaSuccess ? 0 : -static_cast<int>(e);
}

int launch_triangles_mma(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_npad, const uint32_t* d_adj,
                         PairDev* d_state, const ChunkDev* d_chunk, unsigned long long* d_keys, const uint32_t* d_ubase,
                         uint32_t* d_ucursor, int unit_pitch, uint32_t* d_hist, unsigned long long* d_t2, int rank,
                         int world) {
  const int nJ = (max_npad + kMmaJ - 1) / kMmaJ, nIp = (max_npad + kMmaI - 1) / kMmaI;
  int items = 0;
  for (int jq = 0; jq < nJ; ++jq) items += std::min(nIp, (kMmaJ * jq + kMmaJ - 2) / kMmaI + 1);
  dim3 grid(items, pairs);
  triangles_mma_kernel<<<grid, kMmaThreads, mma_smem_bytes(), lc.stream>>>(d_desc, d_adj, d_state, d_chunk, d_keys, d_ubase,
                                                                          d_ucursor, unit_pitch, d_hist, d_t2, rank, world);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

// umma_contend.cu — what limits a tcgen05 (kind::mxf4) main loop whose operands are produced ON CHIP:
//   * correctness of the cheap bit->e2m1 expansion: the contraction index may be permuted freely, so a
//     32-bit adjacency word w becomes four words  w&0x22222222 (nibble 0b0010 = 1.0),  w&0x11111111
//     (0b0001 = 0.5),  (w>>2)&0x22222222,  (w>>2)&0x11111111;  the 0.5 blocks carry UE8M0 scale 2.0 on
//     both operands, so every product is exactly 0 or 1.  5 ALU ops per 32 bits instead of 18.
//   * MMA rate of one SM alone, with concurrent STS.128 streaming, with the full expansion running
//     unsynchronised, and with the real full/empty mbarrier ring (the main loop of the triangle kernel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_contend umma_contend.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}

constexpr int kThreads = 512;
constexpr int kStages = 3;
constexpr int kStageK = 256;                       // K elements per stage = 8 words of bits = 128 B of nibbles per row
constexpr int kLBO = 128;                          // next 16-byte K chunk
constexpr int kSBO = (kStageK / 2 / 16) * 128;     // next 8-row group = 1024 B
constexpr int kProducerWarps = 10;
constexpr uint32_t kSfCol = 448;

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((kLBO >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((kSBO >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}

// one task = (row, quad of 4 words) -> 4 chunks of 16 B
__device__ __forceinline__ void expand_quad(const uint4 w, unsigned char* dst /* chunk 0 of the quad for this row */) {
  const uint32_t m2 = 0x22222222u, m1 = 0x11111111u;
  uint4 c0, c1, c2, c3;
  c0.x = w.x & m2; c0.y = w.y & m2; c0.z = w.z & m2; c0.w = w.w & m2;
  c1.x = w.x & m1; c1.y = w.y & m1; c1.z = w.z & m1; c1.w = w.w & m1;
  const uint32_t sx = w.x >> 2, sy = w.y >> 2, sz = w.z >> 2, sw = w.w >> 2;
  c2.x = sx & m2; c2.y = sy & m2; c2.z = sz & m2; c2.w = sw & m2;
  c3.x = sx & m1; c3.y = sy & m1; c3.z = sz & m1; c3.w = sw & m1;
  *reinterpret_cast<uint4*>(dst) = c0;
  *reinterpret_cast<uint4*>(dst + kLBO) = c1;
  *reinterpret_cast<uint4*>(dst + 2 * kLBO) = c2;
  *reinterpret_cast<uint4*>(dst + 3 * kLBO) = c3;
}

// MODE 0: MMA alone          MODE 1: + STS.128 streaming (same bytes as the expansion, no ALU, no loads)
// MODE 2: + full expansion, unsynchronised     MODE 3: full/empty ring (real main loop)
// rows of the stage: [0,128) A block 0, [128,256) A block 1, [256, 256+N) B.
template <int N, int MODE>
__global__ void __launch_bounds__(kThreads, 1) contend(const uint32_t* __restrict__ bits /* rows x stride words */, int stride,
                                                      int nstages, float* __restrict__ gD, long long* cyc_mma,
                                                      long long* cyc_prod, uint32_t sfword) {
  constexpr int kRows = 256 + N;
  constexpr int kStageBytes = (kRows / 8) * kSBO;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* bitp = bits + static_cast<size_t>(blockIdx.x) * kRows * stride;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], kProducerWarps); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // unsynchronised modes read whatever is in the stages: zero them so the values stay finite
  for (int k = tid; k < kStages * kStageBytes / 16; k += kThreads) reinterpret_cast<uint4*>(smem)[k] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    const uint32_t taddr = tmem + ((32u * warp) << 16) + kSfCol;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
              taddr + 32u * h),
          "r"(sfword)
          : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    asm volatile("bar.sync 1, 160;" ::: "memory");
    // wait for the accumulators, dump them
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (gD != nullptr && blockIdx.x == 0) {
      for (int a = 0; a < 2; ++a)
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          const uint32_t ta = tmem + ((32u * warp) << 16) + static_cast<uint32_t>(N * a + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(ta));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int k = 0; k < 32; ++k) gD[static_cast<size_t>(128 * a + 32 * warp + lane) * N + c0 + k] = __uint_as_float(v[k]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
  } else if (warp == 4) {
    asm volatile("bar.sync 1, 160;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t idesc = (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    const uint32_t sbase = smem_u32(stage_base);
    const long long t0 = clock64();
    for (int it = 0; it < nstages; ++it) {
      const int s = it % kStages;
      if (MODE == 3) {
        mbar_wait(&full[s], static_cast<uint32_t>((it / kStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");
      }
      if (lane == 0) {
        const uint32_t st = sbase + s * kStageBytes;
        const uint32_t bA0 = st, bA1 = st + 16 * kSBO, bB = st + 32 * kSBO;
#pragma unroll
        for (int ks = 0; ks < kStageK / 64; ++ks) {
          const uint64_t db = umma_desc(bB + ks * 2 * kLBO);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const uint64_t da = umma_desc((a == 0 ? bA0 : bA1) + ks * 2 * kLBO);
            const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n}" ::"r"(
                    tmem + static_cast<uint32_t>(N * a)),
                "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(tmem + kSfCol), "r"(tmem + kSfCol + 32u)
                : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        if (it == nstages - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(acc_full)) : "memory");
      }
      __syncwarp();
      if (MODE != 3 && it >= 2) {
        // keep at most 2 stages of MMAs in flight so that the clock reading below is meaningful
        mbar_wait(&empty[(it - 2) % kStages], static_cast<uint32_t>((((it - 2) / kStages)) & 1));
      }
    }
    mbar_wait(acc_full, 0);
    const long long t1 = clock64();
    if (lane == 0) cyc_mma[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
  } else if (warp >= 5 && warp < 5 + kProducerWarps && MODE != 0) {
    const int pw = warp - 5;
    const int r8 = lane & 7, rg2 = (lane >> 3) & 1, q = lane >> 4;
    constexpr int kTasks = kRows / 16;                       // warp task = 16 rows x 2 quads
    constexpr int kMaxT = (kTasks + kProducerWarps - 1) / kProducerWarps;
    const long long t0 = clock64();
    uint4 cur[kMaxT];
    auto load_stage = [&](int it, uint4 (&w)[kMaxT]) {
#pragma unroll
      for (int t = 0; t < kMaxT; ++t) {
        const int task = pw + kProducerWarps * t;
        const int r = 16 * task + 8 * rg2 + r8;
        w[t] = make_uint4(0, 0, 0, 0);
        if (task < kTasks && MODE >= 2) w[t] = *reinterpret_cast<const uint4*>(bitp + static_cast<size_t>(r) * stride + it * 8 + 4 * q);
      }
    };
    load_stage(0, cur);
    for (int it = 0; it < nstages; ++it) {
      const int s = it % kStages;
      uint4 nxt[kMaxT];
      if (it + 1 < nstages) load_stage(it + 1, nxt);
      if (MODE == 3 && it >= kStages) mbar_wait(&empty[s], static_cast<uint32_t>(((it / kStages) - 1) & 1));
      unsigned char* st = stage_base + s * kStageBytes;
#pragma unroll
      for (int t = 0; t < kMaxT; ++t) {
        const int task = pw + kProducerWarps * t;
        if (task < kTasks) {
          unsigned char* dst = st + (2 * task + rg2) * kSBO + (4 * q) * kLBO + r8 * 16;
          if (MODE == 1) {
            const uint4 z = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(dst) = z;
            *reinterpret_cast<uint4*>(dst + kLBO) = z;
            *reinterpret_cast<uint4*>(dst + 2 * kLBO) = z;
            *reinterpret_cast<uint4*>(dst + 3 * kLBO) = z;
          } else {
            expand_quad(cur[t], dst);
          }
        }
      }
      if (MODE == 3) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
      }
      if (it + 1 < nstages) {
#pragma unroll
        for (int t = 0; t < kMaxT; ++t) cur[t] = nxt[t];
      }
    }
    const long long t1 = clock64();
    if (lane == 0 && pw == 0) cyc_prod[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int N, int MODE>
int run(int K, int grid, bool check, uint32_t sfword, const char* label) {
  constexpr int kRows = 256 + N;
  const int stride = K / 32, nstages = K / kStageK;
  std::vector<uint32_t> bits(static_cast<size_t>(grid) * kRows * stride);
  srand(99);
  for (auto& w : bits) w = (rand() & 0xFFFF) | (static_cast<uint32_t>(rand() & 0xFFFF) << 16);
  for (auto& w : bits) w &= (rand() & 0xFFFF) | (static_cast<uint32_t>(rand() & 0xFFFF) << 16);  // density 1/4
  uint32_t* dB; float* dD; long long *dC, *dP;
  CK(cudaMalloc(&dB, bits.size() * 4)); CK(cudaMalloc(&dD, 256 * N * 4)); CK(cudaMalloc(&dC, 8 * grid)); CK(cudaMalloc(&dP, 8 * grid));
  CK(cudaMemcpy(dB, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dP, 0, 8 * grid));
  const size_t smem = static_cast<size_t>(kStages) * (kRows / 8) * kSBO + 8 * 8 + 16;
  CK(cudaFuncSetAttribute(contend<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  for (int rep = 0; rep < 2; ++rep) {  // second launch is the warm one
    contend<N, MODE><<<grid, kThreads, smem>>>(dB, stride, nstages, check ? dD : nullptr, dC, dP, sfword);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> c(grid), p(grid);
  CK(cudaMemcpy(c.data(), dC, 8 * grid, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(p.data(), dP, 8 * grid, cudaMemcpyDeviceToHost));
  double cm = 0, pm = 0;
  for (int g = 0; g < grid; ++g) { cm += c[g]; pm += p[g]; }
  cm /= grid; pm /= grid;
  int bad = 0;
  if (check) {
    std::vector<float> D(256 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        int ref = 0;
        for (int w = 0; w < stride; ++w) ref += __builtin_popcount(bits[static_cast<size_t>(m) * stride + w] & bits[static_cast<size_t>(256 + n) * stride + w]);
        if (D[m * N + n] != static_cast<float>(ref)) { if (bad < 4) printf("  mismatch (%d,%d): got %g want %d\n", m, n, D[m * N + n], ref); ++bad; }
      }
  }
  const double macs = 256.0 * N * K;
  printf("%-34s N=%d K=%d grid=%d sf=%08x: %s mma %.0f clk (%.0f MAC/clk/SM, %.1f%% of 16384), producers %.0f clk\n", label, N, K,
         grid, sfword, check ? (bad ? "MISMATCH" : "exact") : "unchecked", cm, macs / cm, 100.0 * macs / cm / 16384.0, pm);
  cudaFree(dB); cudaFree(dD); cudaFree(dC); cudaFree(dP);
  return bad;
}

int main() {
  int bad = 0;
  // scale-factor byte order: which of the two patterns makes the 0.5 blocks count as 1?
  const int b1 = run<224, 3>(512, 1, true, 0x807F807Fu, "ring, SF bytes {7F,80,7F,80}");
  const int b2 = run<224, 3>(512, 1, true, 0x7F807F80u, "ring, SF bytes {80,7F,80,7F}");
  const uint32_t sf = b1 == 0 ? 0x807F807Fu : 0x7F807F80u;
  bad += (b1 != 0 && b2 != 0);
  bad += run<224, 3>(5120, 1, true, sf, "ring, full K, checked");
  for (int grid : {1, 148}) {
    run<224, 0>(5120 * 4, grid, false, sf, "MMA alone");
    run<224, 1>(5120 * 4, grid, false, sf, "MMA + STS.128 stream");
    run<224, 2>(5120 * 4, grid, false, sf, "MMA + expansion (unsynchronised)");
    run<224, 3>(5120 * 4, grid, false, sf, "MMA + expansion (full/empty ring)");
    run<256, 0>(5120 * 4, grid, false, sf, "MMA alone");
    run<256, 3>(5120 * 4, grid, false, sf, "MMA + expansion (full/empty ring)");
  }
  return bad ? 1 : 0;
}

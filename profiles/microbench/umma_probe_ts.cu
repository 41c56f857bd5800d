// umma_probe_ts.cu — probe of tcgen05.mma kind::mxf4.block_scale with the A operand in TENSOR MEMORY (TS form):
// which K element sits where in a TMEM column, and what one SM sustains.  Derived from umma_probe.cu
// (kind::i8 and kind::mxf4.block_scale)
// with operands written by threads into the no-swizzle K-major canonical shared-memory layout.
// Groundwork for the tensor-core triangle-count variant (K2b): establishes that
//   * 0/1 operands expanded on chip (no TMA tensor maps, SWIZZLE_NONE descriptors) give exact counts,
//   * an all-0x7F (UE8M0 = 1.0) scale-factor region in TMEM makes kind::mxf4 a plain 0/1 dot product,
//   * what one SM sustains for M=128, N=256 tiles in SS mode (operands in shared memory).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes (contiguous 128 B);
// byte address of (row m, K byte kb) = (m/8)*SBO + (kb/16)*LBO + (m%8)*16 + kb%16
__host__ __device__ inline uint32_t canon_off(int m, int kb, int lbo, int sbo) {
  return (m >> 3) * sbo + (kb >> 4) * lbo + (m & 7) * 16 + (kb & 15);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

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

// MODE 0: kind::i8 (u8 x u8 -> s32), K = 32 per MMA (32 bytes);  MODE 1: kind::mxf4.block_scale (e2m1, K = 64 per MMA, 32 bytes)
constexpr uint32_t kAcol = 320;  // A operand: TMEM columns [320, 320 + kbytes / 4)
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) probe(const uint8_t* __restrict__ gA, const uint8_t* __restrict__ gB,
                                               uint32_t* __restrict__ gD, int kbytes, int reps, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  // kbytes of K per row for both operands; LBO = 128 (next 16-byte K chunk), SBO = kbytes/16*128 (next 8 rows)
  const int lbo = 128, sbo = (kbytes / 16) * 128;
  uint8_t* sA = smem;                      // 128 rows
  uint8_t* sB = smem + 128 * kbytes;       // N rows
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int idx = tid; idx < 128 * kbytes; idx += 128) sA[canon_off(idx / kbytes, idx % kbytes, lbo, sbo)] = gA[idx];
  for (int idx = tid; idx < N * kbytes; idx += 128) sB[canon_off(idx / kbytes, idx % kbytes, lbo, sbo)] = gB[idx];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // make the generic-proxy smem writes visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  // A operand: row m = TMEM lane m, 32-bit word j of the row's K bytes = TMEM column kAcol + j
  {
    const uint32_t* rowp = reinterpret_cast<const uint32_t*>(gA + static_cast<size_t>(tid) * kbytes);
    for (int j = 0; j < kbytes / 4; j += 8) {
      uint32_t w[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) w[c] = rowp[j + c];
      const uint32_t taddr = tmem + ((32u * warp) << 16) + kAcol + j;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(w[0]), "r"(w[1]),
                   "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if (MODE == 1) {
    // scale factors: every byte 0x7F (UE8M0 1.0) in columns [N, N+32) of every lane
    uint32_t v = 0x7F7F7F7Fu;
    const uint32_t taddr = tmem + ((32u * warp) << 16) + 256u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
            taddr),
        "r"(v)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }

  uint32_t idesc;
  if (MODE == 0) idesc = (2u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | ((128u >> 4) << 24);
  else idesc = (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);

  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const int ksteps = kbytes / 32;  // one MMA consumes 32 bytes of K per row
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t da = make_desc(a0 + ks * 2 * lbo, lbo, sbo);
        const uint64_t db = make_desc(b0 + ks * 2 * lbo, lbo, sbo);
        const uint32_t acc = (r > 0 || ks > 0) ? 1u : 0u;
        if ((tid & 31) == 0) {
          if (MODE == 0) {
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
                : "memory");
          } else {
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n}" ::"r"(tmem),
                "r"(tmem + kAcol + static_cast<uint32_t>(ks) * 8u), "l"(db), "r"(idesc), "r"(acc), "r"(tmem + 256u), "r"(tmem + 256u + 16u)
                : "memory");
          }
        }
        __syncwarp();
      }
    }
    if ((tid & 31) == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) *cycles = t1 - t0;

  // D: row m = TMEM lane m, column n = TMEM column n.  Warp w reads lanes 32w..32w+31.
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((32u * warp) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 32; ++k) gD[static_cast<size_t>(tid) * N + c0 + k] = v[k];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int MODE, int N>
int run(int kbytes, int reps, bool check) {
  const int kelem = MODE == 0 ? kbytes : kbytes * 2;
  std::vector<uint8_t> bitsA(128 * kelem), bitsB(N * kelem);
  srand(1234 + MODE);
  for (auto& b : bitsA) b = rand() % 3 == 0;
  for (auto& b : bitsB) b = rand() % 3 == 0;
  std::vector<uint8_t> hA(128 * kbytes), hB(N * kbytes);
  auto packrow = [&](const std::vector<uint8_t>& bits, std::vector<uint8_t>& out, int rows) {
    for (int m = 0; m < rows; ++m)
      for (int kb = 0; kb < kbytes; ++kb) {
        if (MODE == 0) out[m * kbytes + kb] = bits[m * kelem + kb];
        else out[m * kbytes + kb] = (bits[m * kelem + 2 * kb] ? 0x2 : 0) | (bits[m * kelem + 2 * kb + 1] ? 0x20 : 0);  // e2m1 1.0 = 0b0010
      }
  };
  packrow(bitsA, hA, 128);
  packrow(bitsB, hB, N);
  uint8_t *dA, *dB; uint32_t* dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  const size_t smem = static_cast<size_t>(128 + N) * kbytes;
  CK(cudaFuncSetAttribute(probe<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  probe<MODE, N><<<1, 128, smem>>>(dA, dB, dD, kbytes, reps, dC);
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> hD(128 * N);
  long long cyc = 0;
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  if (check) {
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        int ref = 0;
        for (int k = 0; k < kelem; ++k) ref += bitsA[m * kelem + k] & bitsB[n * kelem + k];
        ref *= reps;
        int got;
        if (MODE == 0) got = static_cast<int>(hD[m * N + n]);
        else { float f; memcpy(&f, &hD[m * N + n], 4); got = static_cast<int>(f); if (static_cast<float>(got) != f) got = -1; }
        if (got != ref) { if (bad < 5) printf("  mismatch (%d,%d): got %d want %d\n", m, n, got, ref); ++bad; }
      }
  }
  const double macs = 128.0 * N * kelem * reps;
  printf("%s N=%d K=%d elems reps=%d: %s, %lld cycles, %.0f MAC/clk (one SM)\n", MODE == 0 ? "kind::i8   " : "kind::mxf4 ", N,
         kelem, reps, check ? (bad ? "MISMATCH" : "exact") : "unchecked", cyc, macs / cyc);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad;
}

// one-hot test: A row m has its only 1 at K element m % kelem, B row n at K element n: D[m][n] = 1 iff the
// hardware pairs A's element m with B's element n -> prints the K permutation between TMEM-A and SMEM-B
template <int N>
void onehot(int kbytes) {
  const int kelem = kbytes * 2;
  std::vector<uint8_t> hA(128 * kbytes, 0), hB(N * kbytes, 0);
  for (int m = 0; m < 128; ++m) { const int k = m % kelem; hA[m * kbytes + k / 2] = (k & 1) ? 0x20 : 0x2; }
  for (int n = 0; n < N; ++n) { const int k = n % kelem; hB[n * kbytes + k / 2] = (k & 1) ? 0x20 : 0x2; }
  uint8_t *dA, *dB; uint32_t* dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  const size_t smem = static_cast<size_t>(128 + N) * kbytes;
  CK(cudaFuncSetAttribute(probe<1, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  probe<1, N><<<1, 128, smem>>>(dA, dB, dD, kbytes, 1, dC);
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> hD(128 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  printf("one-hot K=%d N=%d: A element -> B element(s) with a non-zero product\n", kelem, N);
  for (int m = 0; m < kelem && m < 128; ++m) {
    printf("  %3d ->", m);
    for (int n = 0; n < N; ++n) { float f; memcpy(&f, &hD[m * N + n], 4); if (f != 0.0f) printf(" %d(%g)", n, f); }
    printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
}

int main() {
  onehot<128>(64);                      // K = 128: two MMAs
  int bad = 0;
  bad += run<1, 64>(64, 1, true);       // mxf4, K = 128
  bad += run<1, 256>(256, 1, true);     // mxf4, N = 256, K = 512 (A needs 64 TMEM columns)
  run<1, 256>(384, 64, false);          // throughput, A resident in TMEM
  run<1, 192>(384, 64, false);
  return bad ? 1 : 0;
}

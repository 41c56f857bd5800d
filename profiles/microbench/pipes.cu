// pipes.cu — issue-rate microbenchmarks for the pipes the SAC-COT kernels are bound by
// ("FP32 non-FMA issue peak, POPC throughput ... to be microbenchmarked by the
// builder").  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Prints lane-operations per clock per SM for each instruction class.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int OP>
__global__ void __launch_bounds__(1024) bench(float* out, unsigned int* outi, float seed) {
  float a[UNROLL], b = seed * 1.0001f, c = seed * 0.5f;
  unsigned int u[UNROLL];
  unsigned long long p[UNROLL];
#pragma unroll
  for (int k = 0; k < UNROLL; ++k) {
    a[k] = seed + k + threadIdx.x;
    u[k] = threadIdx.x * 2654435761u + k;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p[k]) : "f"(a[k]), "f"(a[k] + 1.f));
  }
  unsigned long long pb, pc;
  asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
  asm("mov.b64 %0, {%1,%2};" : "=l"(pc) : "f"(c), "f"(c));
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      if (OP == 0) a[k] = __fmaf_rn(a[k], b, c);                                   // FFMA
      if (OP == 1) a[k] = __fadd_rn(a[k], b);                                      // FADD
      if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[k]) : "l"(pb), "l"(pc));  // FFMA2
      if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pb));              // FADD2
      if (OP == 4) u[k] = __popc(u[k]) + u[k];                                     // POPC (+IADD)
      if (OP == 5) u[k] = (u[k] & 0x5555aaaau) ^ (u[k] >> 3);                      // LOP3/SHF
      if (OP == 6) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k]));      // MUFU.RSQ
      if (OP == 7) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pb));              // FMUL2
      if (OP == 8) u[k] = __reduce_add_sync(0xffffffffu, u[k]);                    // REDUX
      if (OP == 9) u[k] = __brev(u[k]) + 1;                                        // BREV
      if (OP == 10) u[k] = __ffs(u[k]) + u[k];                                     // FLO
    }
  }
  float s = 0; unsigned int su = 0;
#pragma unroll
  for (int k = 0; k < UNROLL; ++k) {
    float lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[k]));
    s += a[k] + lo + hi; su += u[k];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  outi[blockIdx.x * blockDim.x + threadIdx.x] = su;
}

template <int OP>
void run(const char* name, int flops_per_op, float* out, unsigned int* outi, int sms, float mhz) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * 2;
  bench<OP><<<blocks, 1024>>>(out, outi, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<OP><<<blocks, 1024>>>(out, outi, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = double(blocks) * 1024 * ITERS * UNROLL;       // lane-instructions
  const double per_clk_sm = ops / (ms * 1e-3) / (mhz * 1e6) / sms;
  printf("%-10s %8.3f ms  %7.1f lane-instr/clk/SM  (%6.2f T lane-instr/s, x%d elementary ops)\n", name, ms, per_clk_sm,
         ops / (ms * 1e-3) / 1e12, flops_per_op);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float mhz = khz / 1000.0f;
  printf("%s, %d SMs, nominal max %0.f MHz (rates below assume that clock)\n", p.name, p.multiProcessorCount, mhz);
  float* out; unsigned int* outi;
  cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 2 * 1024);
  cudaMalloc(&outi, sizeof(int) * p.multiProcessorCount * 2 * 1024);
  const int sms = p.multiProcessorCount;
  run<0>("FFMA", 1, out, outi, sms, mhz);
  run<1>("FADD", 1, out, outi, sms, mhz);
  run<2>("FFMA2", 2, out, outi, sms, mhz);
  run<3>("FADD2", 2, out, outi, sms, mhz);
  run<7>("FMUL2", 2, out, outi, sms, mhz);
  run<4>("POPC+IADD", 1, out, outi, sms, mhz);
  run<5>("LOP3+SHF", 1, out, outi, sms, mhz);
  run<6>("MUFU.RSQ", 1, out, outi, sms, mhz);
  run<8>("REDUX", 1, out, outi, sms, mhz);
  run<9>("BREV+IADD", 1, out, outi, sms, mhz);
  run<10>("FLO+IADD", 1, out, outi, sms, mhz);
  return 0;
}

// kernels_hypo.cu — S4 (3-point Kabsch per triangle), S5/S6 (K x N scoring fused with the
// block/warp argmax) and S7 (inlier mask of the winner + fp64 refit); SURVEY.md §8a rows S4-S7.
//
// Bit-exactness against the oracle: every fp32 operation below is a single correctly rounded
// IEEE op issued through an __f*_rn intrinsic (never contracted, never reassociated); fused
// multiply-adds appear only where the specification has them (the cross-covariance chain in S4
// and the transform / squared-residual chains in S5) and are explicit __fmaf_rn.  The scoring
// runs on the FP32 CUDA cores (FFMA), not on TF32 tensor cores.  Sums over correspondences are
// integers.  Only the fp64 refit is tolerance-compared.
#include "common.cuh"

#include <algorithm>

namespace saccot {

#define FADD(a, b) __fadd_rn((a), (b))
#define FSUB(a, b) __fsub_rn((a), (b))
#define FMUL(a, b) __fmul_rn((a), (b))
#define FDIV(a, b) __fdiv_rn((a), (b))
#define FSQRT(a) __fsqrt_rn((a))
#define FFMA(a, b, c) __fmaf_rn((a), (b), (c))

// ------------------------------------------------------------------------------------------
// Horn's quaternion solution of the absolute-orientation problem with a fixed-sweep cyclic
// Jacobi eigen-solver on the symmetric 4x4 matrix.  Scalar policy objects give the fp32
// (intrinsics) and fp64 (plain) arithmetic.
// ------------------------------------------------------------------------------------------
struct OpsF32 {
  typedef float T;
  static __device__ __forceinline__ T add(T a, T b) { return FADD(a, b); }
  static __device__ __forceinline__ T sub(T a, T b) { return FSUB(a, b); }
  static __device__ __forceinline__ T mul(T a, T b) { return FMUL(a, b); }
  static __device__ __forceinline__ T div(T a, T b) { return FDIV(a, b); }
  static __device__ __forceinline__ T sqrt(T a) { return FSQRT(a); }
  static __device__ __forceinline__ T abs(T a) { return fabsf(a); }
};
struct OpsF64 {
  typedef double T;
  static __device__ __forceinline__ T add(T a, T b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ T sub(T a, T b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ T mul(T a, T b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ T div(T a, T b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ T sqrt(T a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ T abs(T a) { return fabs(a); }
};

template <typename O, int P, int Q>
__device__ __forceinline__ void jacobi_rotate(typename O::T (&a)[4][4], typename O::T (&v)[4][4]) {
  typedef typename O::T T;
  const T apq = a[P][Q];
  if (apq == T(0)) return;
  const T theta = O::div(O::sub(a[Q][Q], a[P][P]), O::mul(T(2), apq));
  const T ath = O::abs(theta);
  const T rad = O::sqrt(O::add(O::mul(theta, theta), T(1)));
  T t = O::div(T(1), O::add(ath, rad));
  if (theta < T(0)) t = -t;
  const T c = O::div(T(1), O::sqrt(O::add(O::mul(t, t), T(1))));
  const T s = O::mul(t, c);
  const T tapq = O::mul(t, apq);
  a[P][P] = O::sub(a[P][P], tapq);
  a[Q][Q] = O::add(a[Q][Q], tapq);
  a[P][Q] = T(0);
  a[Q][P] = T(0);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (r == P || r == Q) continue;
    const T arp = a[r][P], arq = a[r][Q];
    const T np_ = O::sub(O::mul(c, arp), O::mul(s, arq));
    const T nq_ = O::add(O::mul(s, arp), O::mul(c, arq));
    a[r][P] = np_; a[P][r] = np_;
    a[r][Q] = nq_; a[Q][r] = nq_;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const T vrp = v[r][P], vrq = v[r][Q];
    v[r][P] = O::sub(O::mul(c, vrp), O::mul(s, vrq));
    v[r][Q] = O::add(O::mul(s, vrp), O::mul(c, vrq));
  }
}

// S = cross-covariance (S[r][c] = sum p~[r] q~[c]); writes the rotation R (row-major).
template <typename O, int SWEEPS>
__device__ void horn_rotation(const typename O::T (&S)[3][3], typename O::T (&R)[9]) {
  typedef typename O::T T;
  T a[4][4], v[4][4];
  const T Sxx = S[0][0], Sxy = S[0][1], Sxz = S[0][2];
  const T Syx = S[1][0], Syy = S[1][1], Syz = S[1][2];
  const T Szx = S[2][0], Szy = S[2][1], Szz = S[2][2];
  a[0][0] = O::add(O::add(Sxx, Syy), Szz);
  a[1][1] = O::sub(O::sub(Sxx, Syy), Szz);
  a[2][2] = O::sub(O::sub(Syy, Sxx), Szz);
  a[3][3] = O::sub(O::sub(Szz, Sxx), Syy);
  a[0][1] = a[1][0] = O::sub(Syz, Szy);
  a[0][2] = a[2][0] = O::sub(Szx, Sxz);
  a[0][3] = a[3][0] = O::sub(Sxy, Syx);
  a[1][2] = a[2][1] = O::add(Sxy, Syx);
  a[1][3] = a[3][1] = O::add(Szx, Sxz);
  a[2][3] = a[3][2] = O::add(Syz, Szy);
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) v[r][c] = (r == c) ? T(1) : T(0);
#pragma unroll 1
  for (int sweep = 0; sweep < SWEEPS; ++sweep) {
    jacobi_rotate<O, 0, 1>(a, v);
    jacobi_rotate<O, 0, 2>(a, v);
    jacobi_rotate<O, 0, 3>(a, v);
    jacobi_rotate<O, 1, 2>(a, v);
    jacobi_rotate<O, 1, 3>(a, v);
    jacobi_rotate<O, 2, 3>(a, v);
  }
  // eigenvector of the largest diagonal entry (first one on ties)
  T lam = a[0][0];
  T w = v[0][0], x = v[1][0], y = v[2][0], z = v[3][0];
  if (a[1][1] > lam) { lam = a[1][1]; w = v[0][1]; x = v[1][1]; y = v[2][1]; z = v[3][1]; }
  if (a[2][2] > lam) { lam = a[2][2]; w = v[0][2]; x = v[1][2]; y = v[2][2]; z = v[3][2]; }
  if (a[3][3] > lam) { lam = a[3][3]; w = v[0][3]; x = v[1][3]; y = v[2][3]; z = v[3][3]; }
  const T n2 = O::add(O::add(O::add(O::mul(w, w), O::mul(x, x)), O::mul(y, y)), O::mul(z, z));
  if (!(n2 > T(0))) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    return;
  }
  const T n = O::sqrt(n2);
  w = O::div(w, n); x = O::div(x, n); y = O::div(y, n); z = O::div(z, n);
  const T xx = O::mul(x, x), yy = O::mul(y, y), zz = O::mul(z, z);
  const T xy = O::mul(x, y), xz = O::mul(x, z), yz = O::mul(y, z);
  const T wx = O::mul(w, x), wy = O::mul(w, y), wz = O::mul(w, z);
  R[0] = O::sub(T(1), O::mul(T(2), O::add(yy, zz)));
  R[1] = O::mul(T(2), O::sub(xy, wz));
  R[2] = O::mul(T(2), O::add(xz, wy));
  R[3] = O::mul(T(2), O::add(xy, wz));
  R[4] = O::sub(T(1), O::mul(T(2), O::add(xx, zz)));
  R[5] = O::mul(T(2), O::sub(yz, wx));
  R[6] = O::mul(T(2), O::sub(xz, wy));
  R[7] = O::mul(T(2), O::add(yz, wx));
  R[8] = O::sub(T(1), O::mul(T(2), O::add(xx, yy)));
}

// ------------------------------------------------------------------------------------------
// S4: one thread per hypothesis.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) kabsch_kernel(const PairDesc* __restrict__ descs,
                                                     const float* __restrict__ soa, const int32_t* __restrict__ tri,
                                                     float* __restrict__ rt, int K) {
  const int pair = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= K) return;
  const PairDesc d = descs[pair];
  const int32_t* tr = tri + (static_cast<size_t>(pair) * K + h) * 3;
  float* out = rt + (static_cast<size_t>(pair) * K + h) * 12;
  const int ia = tr[0], ib = tr[1], ic = tr[2];
  if (ia < 0) {
#pragma unroll
    for (int k = 0; k < 12; ++k) out[k] = 0.0f;
    return;
  }
  const float* base = soa + d.soa_off;
  const size_t np = static_cast<size_t>(d.Npad);
  float p[3][3], q[3][3];
  const int idx[3] = {ia, ib, ic};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      p[k][a] = base[a * np + idx[k]];
      q[k][a] = base[(3 + a) * np + idx[k]];
    }
  float pc[3], qc[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    pc[a] = FDIV(FADD(FADD(p[0][a], p[1][a]), p[2][a]), 3.0f);
    qc[a] = FDIV(FADD(FADD(q[0][a], q[1][a]), q[2][a]), 3.0f);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      p[k][a] = FSUB(p[k][a], pc[a]);
      q[k][a] = FSUB(q[k][a], qc[a]);
    }
  float S[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) S[r][c] = FFMA(p[2][r], q[2][c], FFMA(p[1][r], q[1][c], FMUL(p[0][r], q[0][c])));
  float R[9];
  horn_rotation<OpsF32, 8>(S, R);
#pragma unroll
  for (int a = 0; a < 9; ++a) out[a] = R[a];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float rp = FADD(FADD(FMUL(R[3 * a + 0], pc[0]), FMUL(R[3 * a + 1], pc[1])), FMUL(R[3 * a + 2], pc[2]));
    out[9 + a] = FSUB(qc[a], rp);
  }
}

int launch_kabsch(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, const float* d_soa, const int32_t* d_tri,
                  float* d_rt, int K) {
  dim3 grid((K + 127) / 128, pairs);
  kabsch_kernel<<<grid, 128, 0, lc.stream>>>(d_desc, d_soa, d_tri, d_rt, K);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

// ------------------------------------------------------------------------------------------
// S5/S6: each thread scores two hypotheses (held in registers, duplicated into packed lanes)
// against all N correspondences, which the CTA stages through shared memory in 256-point slabs
// laid out as three float4 per PAIR of correspondences, so the inner loop issues three broadcast
// LDS.128 per two points for 2 x 15 packed FP32 instructions (FFMA2/FADD2/FMUL2).
// The packed selection key  score<<16 | (0xFFFF-h)  is max-reduced per warp, per block, and
// with one 64-bit atomicMax per block into the pair's best key (ties -> lowest h).
// ------------------------------------------------------------------------------------------
constexpr int kScoreThreads = 128;
constexpr int kScoreHyp = 2;     // hypotheses per thread
constexpr int kScoreSlab = 256;  // points per shared-memory slab (even)

__device__ __forceinline__ float residual2(const float (&rt)[12], const float4 a, const float4 b) {
  // a = (sx, sy, sz, dx), b = (dy, dz, -, -)
  const float xp = FFMA(rt[0], a.x, FFMA(rt[1], a.y, FFMA(rt[2], a.z, rt[9])));
  const float yp = FFMA(rt[3], a.x, FFMA(rt[4], a.y, FFMA(rt[5], a.z, rt[10])));
  const float zp = FFMA(rt[6], a.x, FFMA(rt[7], a.y, FFMA(rt[8], a.z, rt[11])));
  const float ex = FSUB(xp, a.w), ey = FSUB(yp, b.x), ez = FSUB(zp, b.y);
  return FFMA(ez, ez, FFMA(ey, ey, FMUL(ex, ex)));
}

// Packed fp32x2 arithmetic (sm_100 FFMA2/FADD2/FMUL2): two correspondences per instruction.  Each
// lane of a packed op is an individually rounded IEEE operation, so the results are bit-identical
// to the scalar chain above; every multiply-add here is a specified fma (nothing for ptxas to
// contract).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
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
// squared residuals of two correspondences under one hypothesis (rt2[k] = {rt[k], rt[k]})
__device__ __forceinline__ f32x2 residual2_x2(const f32x2 (&rt2)[12], f32x2 px, f32x2 py, f32x2 pz, f32x2 qx, f32x2 qy,
                                              f32x2 qz) {
  const f32x2 xp = fma2(rt2[0], px, fma2(rt2[1], py, fma2(rt2[2], pz, rt2[9])));
  const f32x2 yp = fma2(rt2[3], px, fma2(rt2[4], py, fma2(rt2[5], pz, rt2[10])));
  const f32x2 zp = fma2(rt2[6], px, fma2(rt2[7], py, fma2(rt2[8], pz, rt2[11])));
  const f32x2 ex = sub2(xp, qx), ey = sub2(yp, qy), ez = sub2(zp, qz);
  return fma2(ez, ez, fma2(ey, ey, mul2(ex, ex)));
}

// Shared-memory loads through explicit 32-bit shared addresses (phase 2 of the scoring kernel): with generic
// pointers the compiler re-derives the shared window base (S2UR SR_CgaCtaId, UMOV, ULEA, ...) in every iteration
// of that divergent loop.
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}

// Phase 2 of the scoring kernel for one hypothesis and one chunk: the specified chain for every candidate in `bits`.
// q4 / q2: shared addresses of correspondence 31 of the chunk in pts4 / pts2.  Point j of the chunk sits at mask bit
// 31 - j: the highest set bit, at position pos, is point 31 - pos, at q - size * pos.
template <int MODE>
__device__ __forceinline__ void score_bits(const float (&rt)[12], uint32_t bits, const uint32_t q4, const uint32_t q2,
                                           const float tau2, unsigned int& cnt, unsigned long long& fsum) {
  while (bits) {
    uint32_t pos;  // one FLO (31 - __clz costs three more instructions here)
    asm("bfind.u32 %0, %1;" : "=r"(pos) : "r"(bits));
    bits ^= 1u << pos;
    const float4 pa = lds_f32x4(q4 - 16u * pos);  // sx, sy, sz, dx
    const float2 pq = lds_f32x2(q2 - 8u * pos);   // dy, dz
    const float r2 = residual2(rt, pa, make_float4(pq.x, pq.y, 0.0f, 0.0f));
    if (MODE == 0) {
      asm("{\n.reg .pred p;\nsetp.lt.f32 p, %1, %2;\n@p add.u32 %0, %0, 1;\n}" : "+r"(cnt) : "f"(r2), "f"(tau2));
    } else {
      const float mm = r2 < tau2 ? r2 : tau2;  // NaN -> tau2
      fsum += static_cast<unsigned int>(FMUL(FDIV(mm, tau2), 1048576.0f));
    }
  }
}

// Two phases per 32 correspondences.  Phase 1 (packed, every correspondence): x' and the sign of ex^2 - tau^2 only —
// 5 packed FP instructions per two points instead of 15.  r2 = fma(ez,ez,fma(ey,ey,ex*ex)) >= RN(ex*ex) in fp32 as
// in real arithmetic (each fma adds a non-negative term and rounding is monotone), so a point with ex^2 >= tau^2
// cannot be an inlier and contributes the constant 2^20 in mode 1; ~89 % of the points of a 3 m scene leave here.
// Phase 2 (scalar, the survivors, one bit per point in a register mask): the full specified chain, bit-identical to
// the oracle's.  On the headline workload 10.6 % of the (hypothesis, correspondence) pairs get here (the 5 % inliers
// and as many near misses); a warp runs for the largest count among its lanes (4.6 per 32 correspondences against
// a mean of 3.4), which made this loop more than half of the kernel's instructions at 38 per iteration.  It reads
// the correspondence from a second, point-by-point copy of the slab (one LDS.128 + one LDS.64 at explicit shared
// addresses instead of six LDS.32 and a re-derived base) and finds it with one FLO: 29 instructions per iteration,
// 56 registers (nine CTAs per SM instead of seven), 1.99 -> 1.77 ms per 256-pair step.  Tried and measured slower
// (profiles/experiments/README.md, round 2): deferring phase 2 to the end of the slab with the masks in shared memory
// (every lane walking its own list: fewer iterations, but 66 registers and a word-advance in the loop: +5 %); the
// warp's COMMON candidates in a uniform loop first (+2 %); two candidates per iteration in packed lanes (the
// packing moves cost more than the halved FP issue: +5 %).  ncu of the one-phase version
// (`profiles/ncu_score_r01a.txt`): FP32 pipe saturated by 15 FFMA2-class instructions per point pair and hypothesis,
// issue slots 56 % used.
//
// nsplit > 1 (few pairs, many points: the single-pair calls): the hypothesis blocks alone would leave most SMs idle,
// so the correspondences are cut into nsplit ranges of whole slabs, one CTA per (hypothesis block, range); a CTA adds
// its partial integer score to hyp_key[h] (zeroed by the launcher; integer adds commute, so the total is exact and
// order-free) and score_finish_kernel turns the totals into keys and takes the argmax.
template <int MODE>
__global__ void __launch_bounds__(kScoreThreads) score_kernel(
    const PairDesc* __restrict__ descs, const float* __restrict__ soa, const int32_t* __restrict__ tri,
    const float* __restrict__ rt_all, unsigned long long* __restrict__ hyp_key, PairDev* __restrict__ state,
    float tau2, int K, int h_begin, int h_end, int nsplit) {
  const int pair = blockIdx.y;
  const PairDesc d = descs[pair];
  const int hblock = static_cast<int>(blockIdx.x) / nsplit, split = static_cast<int>(blockIdx.x) % nsplit;
  int n_begin = 0, n_end = d.N;
  if (nsplit > 1) {
    const int slabs = (d.N + kScoreSlab - 1) / kScoreSlab, per = (slabs + nsplit - 1) / nsplit;
    n_begin = min(d.N, split * per * kScoreSlab);
    n_end = min(d.N, n_begin + per * kScoreSlab);
  }
  // slab[p] holds correspondences 2p and 2p+1:  {sx0,sx1,sy0,sy1} {sz0,sz1,dx0,dx1} {dy0,dy1,dz0,dz1}  (phase 1)
  // pts4[n] / pts2[n]: the same correspondences one by one, {sx,sy,sz,dx} / {dy,dz}                    (phase 2)
  __shared__ float4 slab[kScoreSlab / 2][3];
  __shared__ float4 pts4[kScoreSlab];
  __shared__ __align__(16) float2 pts2[kScoreSlab];
  __shared__ unsigned long long wbest[kScoreThreads / 32];

  const int tid = threadIdx.x;
  int hid[kScoreHyp];
  bool valid[kScoreHyp];
  float rt[kScoreHyp][12];
  f32x2 rx2[kScoreHyp][4];  // packed copies of R00, R01, R02, tx for phase 1
#pragma unroll
  for (int u = 0; u < kScoreHyp; ++u) {
    hid[u] = h_begin + hblock * (kScoreThreads * kScoreHyp) + u * kScoreThreads + tid;
    valid[u] = hid[u] < h_end && tri[(static_cast<size_t>(pair) * K + (hid[u] < K ? hid[u] : 0)) * 3] >= 0;
    const float* src_rt = rt_all + (static_cast<size_t>(pair) * K + (hid[u] < K ? hid[u] : 0)) * 12;
#pragma unroll
    for (int k = 0; k < 12; ++k) rt[u][k] = valid[u] ? src_rt[k] : 0.0f;
    rx2[u][0] = pk(rt[u][0], rt[u][0]);
    rx2[u][1] = pk(rt[u][1], rt[u][1]);
    rx2[u][2] = pk(rt[u][2], rt[u][2]);
    rx2[u][3] = pk(rt[u][9], rt[u][9]);
  }
  unsigned int cnt[kScoreHyp];      // mode 0: inliers; mode 1: phase-2 candidates
  unsigned long long fsum[kScoreHyp];
#pragma unroll
  for (int u = 0; u < kScoreHyp; ++u) { cnt[u] = 0; fsum[u] = 0; }

  const float* base = soa + d.soa_off;
  const size_t np = static_cast<size_t>(d.Npad);
  const float nan = __int_as_float(0x7fc00000);
  const uint32_t q4_0 = smem_u32(&pts4[31]), q2_0 = smem_u32(&pts2[31]);  // correspondence 31 of the slab's first chunk
  for (int n0 = n_begin; n0 < n_end; n0 += kScoreSlab) {
    const int nn = min(kScoreSlab, n_end - n0);
    const int npairs = (nn + 1) / 2;
    __syncthreads();  // previous slab fully consumed
    for (int k = tid; k < kScoreSlab / 2; k += kScoreThreads) {
      const int n = n0 + 2 * k;
      float2 sx, sy, sz, dx, dy, dz;
      if (k < npairs) {
        // the SoA arrays are padded to Npad (even) with NaN, so n+1 is always readable; a NaN
        // correspondence never passes phase 2
        sx = *reinterpret_cast<const float2*>(base + n);
        sy = *reinterpret_cast<const float2*>(base + np + n);
        sz = *reinterpret_cast<const float2*>(base + 2 * np + n);
        dx = *reinterpret_cast<const float2*>(base + 3 * np + n);
        dy = *reinterpret_cast<const float2*>(base + 4 * np + n);
        dz = *reinterpret_cast<const float2*>(base + 5 * np + n);
        if (n + 1 >= n_end) { dx.y = nan; dy.y = nan; dz.y = nan; }
      } else {  // tail of the last slab: NaN, never an inlier
        sx = sy = sz = dx = dy = dz = make_float2(nan, nan);
      }
      slab[k][0] = make_float4(sx.x, sx.y, sy.x, sy.y);
      slab[k][1] = make_float4(sz.x, sz.y, dx.x, dx.y);
      slab[k][2] = make_float4(dy.x, dy.y, dz.x, dz.y);
      pts4[2 * k] = make_float4(sx.x, sy.x, sz.x, dx.x);
      pts4[2 * k + 1] = make_float4(sx.y, sy.y, sz.y, dx.y);
      *reinterpret_cast<float4*>(&pts2[2 * k]) = make_float4(dy.x, dz.x, dy.y, dz.y);
    }
    __syncthreads();
    const int nchunks = (npairs + 15) >> 4;
    for (int ch = 0; ch < nchunks; ++ch) {
      const int kc = 16 * ch;
      // ---- phase 1: 32 correspondences.  The candidate bit of a point is the sign of ex*ex - tau^2, shifted into the
      //      mask with one funnel shift: point 2 j + p of the chunk ends up at bit 31 - (2 j + p).  (A NaN may land
      //      either way; phase 2 is exact for whatever it is given.) ----
      uint32_t m[kScoreHyp];
#pragma unroll
      for (int u = 0; u < kScoreHyp; ++u) m[u] = 0u;
      const f32x2 ntau2 = pk(-tau2, -tau2);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 a = slab[kc + j][0], b = slab[kc + j][1];
        const f32x2 px = pk(a.x, a.y), py = pk(a.z, a.w), pz = pk(b.x, b.y), qx = pk(b.z, b.w);
#pragma unroll
        for (int u = 0; u < kScoreHyp; ++u) {
          const f32x2 xp = fma2(rx2[u][0], px, fma2(rx2[u][1], py, fma2(rx2[u][2], pz, rx2[u][3])));
          const f32x2 ex = sub2(xp, qx);
          // RN(ex^2 - tau^2) with ONE rounding: its sign is the sign of the exact difference, and
          // RN(ex*ex) < tau^2 implies ex^2 < tau^2 (rounding is monotone, tau^2 is a float): no inlier is lost
          float d0, d1;
          unpk(fma2(ex, ex, ntau2), d0, d1);
          m[u] = __funnelshift_l(__float_as_uint(d0), m[u], 1);
          m[u] = __funnelshift_l(__float_as_uint(d1), m[u], 1);
        }
      }
      // ---- phase 2: the full chain for the candidates ----
#pragma unroll
      for (int u = 0; u < kScoreHyp; ++u) {
        if (MODE == 1) cnt[u] += __popc(m[u]);
        score_bits<MODE>(rt[u], m[u], q4_0 + 512u * static_cast<uint32_t>(ch), q2_0 + 256u * static_cast<uint32_t>(ch), tau2,
                         cnt[u], fsum[u]);
      }
    }
  }
  if (MODE == 1) {
    // every correspondence that did not reach phase 2 has r2 >= tau^2 (or NaN) and contributes what the chain
    // gives for min(r2, tau2) = tau2 (2^20 for any sane tau)
    const unsigned int far = static_cast<unsigned int>(FMUL(FDIV(tau2, tau2), 1048576.0f));
#pragma unroll
    for (int u = 0; u < kScoreHyp; ++u)
      fsum[u] += static_cast<unsigned long long>(static_cast<unsigned int>(n_end - n_begin) - cnt[u]) * far;
  }
  if (nsplit > 1) {  // partial integer scores; score_finish_kernel builds the keys
#pragma unroll
    for (int u = 0; u < kScoreHyp; ++u) {
      const unsigned long long part = MODE == 0 ? static_cast<unsigned long long>(cnt[u]) : fsum[u];
      if (valid[u] && part) atomicAdd(&hyp_key[static_cast<size_t>(pair) * K + hid[u]], part);
    }
    return;
  }

  unsigned long long best = 0;
#pragma unroll
  for (int u = 0; u < kScoreHyp; ++u) {
    unsigned long long key = 0;
    if (valid[u]) {
      const unsigned long long score = MODE == 0 ? static_cast<unsigned long long>(cnt[u]) + 1ull
                                                 : (static_cast<unsigned long long>(d.N) << 20) - fsum[u] + 1ull;
      key = (score << 16) | static_cast<unsigned long long>(0xFFFFu - static_cast<unsigned int>(hid[u]));
    }
    if (hid[u] < h_end) hyp_key[static_cast<size_t>(pair) * K + hid[u]] = key;
    best = key > best ? key : best;
  }
  best = warp_max_u64(best);
  if ((tid & 31) == 0) wbest[tid >> 5] = best;
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int w = 1; w < kScoreThreads / 32; ++w) best = wbest[w] > best ? wbest[w] : best;
    if (best) atomicMax(&state[pair].best_key, best);
  }
}

// Split scoring, second half: hyp_key[h] holds the integer total (mode 0: inliers, mode 1: sum of the fixed-point
// residuals) of hypothesis h; turn it into the selection key and max-reduce into the pair's best key.
template <int MODE>
__global__ void __launch_bounds__(256) score_finish_kernel(const PairDesc* __restrict__ descs,
                                                           const int32_t* __restrict__ tri,
                                                           unsigned long long* __restrict__ hyp_key,
                                                           PairDev* __restrict__ state, int K, int h_begin, int h_end) {
  const int pair = blockIdx.y;
  const int h = h_begin + blockIdx.x * 256 + threadIdx.x;
  __shared__ unsigned long long wbest[8];
  unsigned long long key = 0;
  if (h < h_end) {
    if (tri[(static_cast<size_t>(pair) * K + h) * 3] >= 0) {
      const unsigned long long total = hyp_key[static_cast<size_t>(pair) * K + h];
      const unsigned long long score = MODE == 0 ? total + 1ull
                                                 : (static_cast<unsigned long long>(descs[pair].N) << 20) - total + 1ull;
      key = (score << 16) | static_cast<unsigned long long>(0xFFFFu - static_cast<unsigned int>(h));
    }
    hyp_key[static_cast<size_t>(pair) * K + h] = key;
  }
  unsigned long long best = warp_max_u64(key);
  if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) best = wbest[w] > best ? wbest[w] : best;
    if (best) atomicMax(&state[pair].best_key, best);
  }
}

int launch_score(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, int max_n, const float* d_soa,
                 const int32_t* d_tri, const float* d_rt, unsigned long long* d_hyp_key, PairDev* d_state, float tau2,
                 int K, int h_begin, int h_end, int mode) {
  const int span = h_end - h_begin;
  if (span <= 0) return 0;
  const int hblocks = (span + kScoreThreads * kScoreHyp - 1) / (kScoreThreads * kScoreHyp);
  // fewer than two CTAs per SM (single pairs, chunks of a dozen pairs): split the correspondences as well, into
  // ranges of at least 2 slabs, aiming at four CTAs per SM.  Measured (score us per chunk, N = 5000): 8 pairs
  // 125 -> 109, 11 pairs 232 -> 138, 16 pairs 233 -> 210; from 24 pairs (384 CTAs) on splitting is slower
  // (280 -> 298; 32 pairs 334 -> 356).
  int nsplit = 1;
  const int slabs = (max_n + kScoreSlab - 1) / kScoreSlab;
  if (hblocks * pairs < 2 * lc.sm_count && slabs >= 8) {
    nsplit = std::min((4 * lc.sm_count + hblocks * pairs - 1) / (hblocks * pairs), slabs / 2);
    if (nsplit < 2) nsplit = 1;
  }
  if (nsplit > 1) {  // partial scores add up in hyp_key: start from zero (one memset when the range is all of K)
    const int nset = span == K ? 1 : pairs;
    const size_t bytes = sizeof(unsigned long long) * (span == K ? static_cast<size_t>(pairs) * K : static_cast<size_t>(span));
    for (int b = 0; b < nset; ++b) {
      const cudaError_t me = cudaMemsetAsync(d_hyp_key + static_cast<size_t>(b) * K + h_begin, 0, bytes, lc.stream);
      if (me != cudaSuccess) return -static_cast<int>(me);
    }
  }
  dim3 grid(hblocks * nsplit, pairs);
  if (mode == 0)
    score_kernel<0><<<grid, kScoreThreads, 0, lc.stream>>>(d_desc, d_soa, d_tri, d_rt, d_hyp_key, d_state, tau2, K,
                                                           h_begin, h_end, nsplit);
  else
    score_kernel<1><<<grid, kScoreThreads, 0, lc.stream>>>(d_desc, d_soa, d_tri, d_rt, d_hyp_key, d_state, tau2, K,
                                                           h_begin, h_end, nsplit);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -static_cast<int>(e);
  if (nsplit == 1) return 1;
  dim3 fgrid((span + 255) / 256, pairs);
  if (mode == 0) score_finish_kernel<0><<<fgrid, 256, 0, lc.stream>>>(d_desc, d_tri, d_hyp_key, d_state, K, h_begin, h_end);
  else score_finish_kernel<1><<<fgrid, 256, 0, lc.stream>>>(d_desc, d_tri, d_hyp_key, d_state, K, h_begin, h_end);
  e = cudaGetLastError();
  return e == cudaSuccess ? 2 : -static_cast<int>(e);
}

// ------------------------------------------------------------------------------------------
// S7: one CTA per pair.  Inlier mask of the winning hypothesis (same fp32 test as S5), then a
// Kabsch over the inliers with fp64 accumulation (fixed-order block reductions => run-to-run
// deterministic) and an fp64 Horn/Jacobi solve by thread 0; output rounded to fp32.
// ------------------------------------------------------------------------------------------
constexpr int kFinThreads = 512;  // (256 before: 132 us for the single N = 50 000 pair, two dependent passes over the points)

__device__ void block_sum_f64(double* vals, int count, double* scratch /* [count][kFinThreads/32] */) {
  // reduces `count` per-thread doubles across the CTA in a fixed order; result in vals[] of thread 0
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = 0; c < count; ++c) {
    double v = vals[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) scratch[c * (kFinThreads / 32) + warp] = v;
  }
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < count; ++c) {
      double v = 0.0;
      for (int w = 0; w < kFinThreads / 32; ++w) v += scratch[c * (kFinThreads / 32) + w];
      vals[c] = v;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kFinThreads) finalize_kernel(
    const PairDesc* __restrict__ descs, const float* __restrict__ soa, const float* __restrict__ rt_all,
    const PairDev* __restrict__ state, const unsigned long long* __restrict__ best_override,
    uint32_t* __restrict__ mask, float* __restrict__ outR, float* __restrict__ outT, int32_t* __restrict__ outInl,
    float tau2, int K, int refit) {
  const int pair = blockIdx.x;
  const PairDesc d = descs[pair];
  const int tid = threadIdx.x, lane = tid & 31;
  __shared__ double scratch[9 * (kFinThreads / 32)];
  __shared__ double s_c[6];
  __shared__ int s_cnt;

  float* R = outR + static_cast<size_t>(pair) * 9;
  float* T = outT + static_cast<size_t>(pair) * 3;
  uint32_t* maskp = mask + d.mask_off;
  const unsigned long long best = best_override ? *best_override : state[pair].best_key;
  if (best == 0ull) {
    for (int w = tid; w < d.Npad / 32; w += kFinThreads) maskp[w] = 0u;
    if (tid == 0) {
      R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
      T[0] = 0; T[1] = 0; T[2] = 0;
      outInl[pair] = 0;
    }
    return;
  }
  const int h = static_cast<int>(0xFFFFu - static_cast<unsigned int>(best & 0xFFFFull));
  float rt[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) rt[k] = rt_all[(static_cast<size_t>(pair) * K + h) * 12 + k];

  const float* base = soa + d.soa_off;
  const size_t np = static_cast<size_t>(d.Npad);
  // pass 1: mask, count, fp64 coordinate sums
  double acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0;
  int cnt = 0;
  for (int n0 = 0; n0 < d.Npad; n0 += kFinThreads) {
    const int n = n0 + tid;  // Npad is a multiple of 128, not of kFinThreads: guard n < Npad
    bool in = false;
    float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
    if (n < d.N) {
      a = make_float4(base[n], base[np + n], base[2 * np + n], base[3 * np + n]);
      b = make_float4(base[4 * np + n], base[5 * np + n], 0.0f, 0.0f);
      in = residual2(rt, a, b) < tau2;
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, in);
    if (lane == 0 && n < d.Npad) maskp[n >> 5] = bits;
    if (in) {
      ++cnt;
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z;
      acc[3] += a.w; acc[4] += b.x; acc[5] += b.y;
    }
  }
  int wc = __reduce_add_sync(0xffffffffu, cnt);
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  if (lane == 0 && wc) atomicAdd(&s_cnt, wc);
  block_sum_f64(acc, 6, scratch);  // contains __syncthreads
  const int total = s_cnt;
  if (tid == 0) {
    outInl[pair] = total;
    if (total > 0)
      for (int k = 0; k < 6; ++k) s_c[k] = acc[k] / static_cast<double>(total);
  }
  __syncthreads();
  if (!refit || total < 3) {
    if (tid < 9) R[tid] = rt[tid];
    if (tid < 3) T[tid] = rt[9 + tid];
    return;
  }
  // pass 2: centred cross-covariance over the inliers
  const double pcx = s_c[0], pcy = s_c[1], pcz = s_c[2], qcx = s_c[3], qcy = s_c[4], qcz = s_c[5];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0;
  for (int n = tid; n < d.N; n += kFinThreads) {
    if ((maskp[n >> 5] >> (n & 31)) & 1u) {
      const double px = static_cast<double>(base[n]) - pcx, py = static_cast<double>(base[np + n]) - pcy,
                   pz = static_cast<double>(base[2 * np + n]) - pcz;
      const double qx = static_cast<double>(base[3 * np + n]) - qcx, qy = static_cast<double>(base[4 * np + n]) - qcy,
                   qz = static_cast<double>(base[5 * np + n]) - qcz;
      acc[0] += px * qx; acc[1] += px * qy; acc[2] += px * qz;
      acc[3] += py * qx; acc[4] += py * qy; acc[5] += py * qz;
      acc[6] += pz * qx; acc[7] += pz * qy; acc[8] += pz * qz;
    }
  }
  block_sum_f64(acc, 9, scratch);
  if (tid == 0) {
    double S[3][3] = {{acc[0], acc[1], acc[2]}, {acc[3], acc[4], acc[5]}, {acc[6], acc[7], acc[8]}};
    double Rd[9];
    horn_rotation<OpsF64, 12>(S, Rd);
    for (int k = 0; k < 9; ++k) R[k] = static_cast<float>(Rd[k]);
    const double pc[3] = {pcx, pcy, pcz}, qc[3] = {qcx, qcy, qcz};
    for (int a = 0; a < 3; ++a) {
      const double rp = (Rd[3 * a + 0] * pc[0] + Rd[3 * a + 1] * pc[1]) + Rd[3 * a + 2] * pc[2];
      T[a] = static_cast<float>(qc[a] - rp);
    }
  }
}

int launch_finalize(const LaunchCtx& lc, const PairDesc* d_desc, int pairs, const float* d_soa, const float* d_rt,
                    const PairDev* d_state, const unsigned long long* d_best_override, uint32_t* d_mask, float* d_R,
                    float* d_t, int32_t* d_inl, float tau2, int K, int refit) {
  finalize_kernel<<<pairs, kFinThreads, 0, lc.stream>>>(d_desc, d_soa, d_rt, d_state, d_best_override, d_mask, d_R,
                                                        d_t, d_inl, tau2, K, refit);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -static_cast<int>(e);
}

}  // namespace saccot

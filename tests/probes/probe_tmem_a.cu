// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (A staged smem -> TMEM once with tcgen05.cp, then reused by several
// MMAs that fetch only B from shared memory).  Checks the result against the ordinary smem-A MMA and measures the rate of
// N=96 MMAs (the ring wgrad's shape) in both forms.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_tmem_a.bin
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
  d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(taddr), "l"(sdesc));
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int N = 96, ACOL = 480;

// smem: A [2 k-chunks][128 rows][8 halves], B [2][N rows][8 halves]
__global__ void __launch_bounds__(384) probe(const __half* A, const __half* B, float* Dref, float* Dts, int reps, long long* cyc, int mode) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  __half* sA = reinterpret_cast<__half*>(smem);
  __half* sB = sA + 2 * 128 * 8;
  const int tid = threadIdx.x, warp = tid / 32;
  if (tid < 128) {
  for (int i = tid; i < 2 * 128 * 8; i += 128) sA[i] = A[i];
  for (int i = tid; i < 2 * N * 8; i += 128) sB[i] = B[i];
  }
  if (tid == 0) for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base_s;
  const uint32_t idesc = idesc_f16(128, N);
  const uint64_t ad = make_desc(smem_u32(sA), 128, 8), bd = make_desc(smem_u32(sB), N, 8);
  if (tid == 0) {
    mma_ss(tm, ad, bd, idesc, 0);                     // reference: columns 0..95
    cp_128x256b(tm + ACOL, ad);                       // A -> TMEM columns 480..487
    mma_ts(tm + 128, tm + ACOL, bd, idesc, 0);        // columns 128..223
    commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    // rates: 5 accumulators (96 columns each) round robin, as in the ring wgrad kernel
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
#pragma unroll
      for (int q = 0; q < 10; ++q) mma_ss(tm + 96 * (q % 5), ad, (mode & 2) ? bd + (uint64_t)(((r * 10 + q) * 7 % 50) * 192) : bd, idesc, 1);
    commit(&bar[1]);
    mbar_wait(&bar[1], 0);
    long long t1 = clock64();
    for (int r = 0; r < reps; ++r) {
      cp_128x256b(tm + ACOL, ad);
#pragma unroll
      for (int q = 0; q < 10; ++q) mma_ts(tm + 96 * (q % 5), tm + ACOL, (mode & 2) ? bd + (uint64_t)(((r * 10 + q) * 7 % 50) * 192) : bd, idesc, 1);
    }
    commit(&bar[2]);
    mbar_wait(&bar[2], 0);
    long long t2 = clock64();
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = 10LL * reps;
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(smem_u32(&bar[3])) : "memory");
  }
  if (tid >= 128 && (mode & 1)) mbar_wait(&bar[3], 0);   // 8 warps polling a barrier for the whole timed region, every lane
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (reps == 0 && tid < 128)
    for (int half = 0; half < 2; ++half)
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + half * 128 + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* D = half ? Dts : Dref;
        for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
      }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
}


__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_ss_w(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t}\n"
               :: "r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
// how: 0 = one thread (if (tid == 0) region), 1 = whole warp + lane 0 predicate, 2 = whole warp + elect.sync predicate.
// B descriptor varies per MMA (ring-like slots computed from the loop counter), 5 accumulators round robin.
template <int how>
__global__ void __launch_bounds__(128) issue_forms(int reps, long long* cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;
  uint32_t* w32 = reinterpret_cast<uint32_t*>(smem);
  for (int i = tid; i < (2 * 128 * 16 + 58 * 192 * 16) / 4; i += 128) w32[i] = 0x3c003c00u;
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  const uint32_t idesc = idesc_f16(128, N);
  const uint64_t ad = make_desc(smem_u32(smem), 128, 8), bd = make_desc(smem_u32(smem) + 2 * 128 * 16, 192, 8);
  const uint32_t alo = (uint32_t)ad, ahi = (uint32_t)(ad >> 32), blo = (uint32_t)bd, bhi = (uint32_t)(bd >> 32);
  if (warp == 0 && (how != 0 || lane == 0)) {
    const bool leader = how == 2 ? elect_one() : lane == 0;
    long long t0 = clock64();
    int c0 = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        int slot = c0 + 7 * q;
        if (slot >= 56) slot -= 56;
        if (leader) mma_ss_w(tm + 96 * (q % 5), alo, ahi, blo + (uint32_t)(slot * 192), bhi, idesc);
      }
      c0 += 8;
      if (c0 >= 56) c0 -= 56;
    }
    if (leader) { commit(&bar); mbar_wait(&bar, 0); cyc[0] = clock64() - t0; cyc[1] = 10LL * reps; }
    if (how != 0) __syncwarp();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
}
int main() {
  std::vector<__half> hA(2 * 128 * 8), hB(2 * N * 8);
  std::vector<float> fA(128 * 16), fB(N * 16);
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) { float v = (float)((m * 7 + k * 3) % 11 - 5); fA[m * 16 + k] = v; hA[((k / 8) * 128 + m) * 8 + k % 8] = __float2half(v); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < 16; ++k) { float v = (float)((n * 5 + k * 2) % 7 - 3); fB[n * 16 + k] = v; hB[((k / 8) * N + n) * 8 + k % 8] = __float2half(v); }
  __half *dA, *dB; float *dR, *dT; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dR, 128 * N * 4)); CK(cudaMalloc(&dT, 128 * N * 4)); CK(cudaMalloc(&dC, 64));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dR, 0, 128 * N * 4)); CK(cudaMemset(dT, 0xff, 128 * N * 4));
  size_t smem = 2 * 128 * 16 + 58 * 192 * 16 + 256;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<<<1, 128, smem>>>(dA, dB, dR, dT, 0, dC, 0);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  std::vector<float> R(128 * N), T(128 * N);
  CK(cudaMemcpy(R.data(), dR, R.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(T.data(), dT, T.size() * 4, cudaMemcpyDeviceToHost));
  double e_ref = 0, e_ts = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
    float s = 0; for (int k = 0; k < 16; ++k) s += fA[m * 16 + k] * fB[n * 16 + k];
    e_ref = fmax(e_ref, fabs(R[m * N + n] - s)); e_ts = fmax(e_ts, fabs(T[m * N + n] - s));
  }
  printf("max |D - host|: smem-A MMA %.3g, TMEM-A MMA (tcgen05.cp 128x256b) %.3g\n", e_ref, e_ts);
  if (e_ts != 0) { printf("first rows of TMEM-A result vs host:\n"); for (int m = 0; m < 4; ++m) { for (int n = 0; n < 6; ++n) { float s = 0; for (int k = 0; k < 16; ++k) s += fA[m * 16 + k] * fB[n * 16 + k]; printf(" %7.1f/%-7.1f", T[m * N + n], s); } printf("\n"); } }
  for (int mode = 0; mode < 4; ++mode) {
    const int grid = 148;
    probe<<<grid, (mode & 1) ? 384 : 128, smem>>>(dA, dB, dR, dT, 400, dC, mode);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    long long c[3]; CK(cudaMemcpy(c, dC, 24, cudaMemcpyDeviceToHost));
    printf("pollers %d ring-like B %d: N=96 MMA, 10 per A block over 5 accumulators: A from smem %.1f cyc/MMA, A from TMEM (+1 cp per 10 MMAs) %.1f cyc/MMA\n",
           mode & 1, (mode >> 1) & 1, (double)c[0] / c[2], (double)c[1] / c[2]);
  }
  CK(cudaFuncSetAttribute(issue_forms<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(issue_forms<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(issue_forms<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* names[3] = {"one thread inside if (tid == 0)", "whole warp, lane-0 predicate", "whole warp, elect.sync predicate"};
  for (int how = 0; how < 3; ++how) {
    if (how == 0) issue_forms<0><<<148, 128, smem>>>(400, dC); else if (how == 1) issue_forms<1><<<148, 128, smem>>>(400, dC); else issue_forms<2><<<148, 128, smem>>>(400, dC);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    long long c[2]; CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
    printf("issue form: %-36s N=96 smem-A MMAs with per-MMA varying B descriptors: %.1f cyc/MMA\n", names[how], (double)c[0] / c[1]);
  }
  return 0;
}

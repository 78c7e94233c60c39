// Probe 2: MN-major (transposed) tf32 operands and the M=64 accumulator layout, as needed by the tcgen05 weight-gradient
// kernel:  D[m][n] = sum_k A[k][m] * B[k][n]  with both operands stored [chunk of 4 m|n][k rows][4 floats].
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// variant: 0 -> (LBO = 8, SBO = KR) ; 1 -> (LBO = KR, SBO = 8)
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int M, int N, int KR, int K, int variant, long long* cyc, int reps, int majA, int majB, int layoutType) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  float* sA = (float*)smem; float* sB = sA + (M / 4) * KR * 4;
  int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < (M / 4) * KR * 4; i += 128) sA[i] = A[i];
  for (int i = tid; i < (N / 4) * KR * 4; i += 128) sB[i] = B[i];
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tb)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem = tb;
  uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)majA << 15) | ((uint32_t)majB << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t aa = smem_u32(sA) + k0 * 16, bb = smem_u32(sB) + k0 * 16;
        uint64_t ad = variant == 0 ? make_desc(aa, 8, KR) : make_desc(aa, KR, 8);
        uint64_t bd = variant == 0 ? make_desc(bb, 8, KR) : make_desc(bb, KR, 8);
        ad |= (uint64_t)layoutType << 61; bd |= (uint64_t)layoutType << 61;
        uint32_t acc = (r > 0 || k0 > 0) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    unsigned long long spins = 0;
    while (!mbar_try(&bar, 0)) { if (++spins > (1ull << 26)) __trap(); }
    if (cyc) *cyc = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}
static float tr(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
int main() {
  const int KR = 72, K = 64;
  for (int M : {128}) for (int N : {32}) {
    std::vector<float> LA(M * KR), LB(N * KR), hA((M / 4) * KR * 4), hB((N / 4) * KR * 4);
    srand(7);
    for (auto& v : LA) v = (rand() / (float)RAND_MAX) * 2 - 1;
    for (auto& v : LB) v = (rand() / (float)RAND_MAX) * 2 - 1;
    for (int m = 0; m < M; ++m) for (int k = 0; k < KR; ++k) hA[((m / 4) * KR + k) * 4 + m % 4] = LA[m * KR + k];
    for (int n = 0; n < N; ++n) for (int k = 0; k < KR; ++k) hB[((n / 4) * KR + k) * 4 + n % 4] = LB[n * KR + k];
    float *dA, *dB, *dD; long long* dC;
    CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    size_t smem = (hA.size() + hB.size()) * 4 + 256;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int majA = 0; majA < 2; ++majA) for (int majB = 0; majB < 2; ++majB) for (int variant = 0; variant < 2; ++variant) {
      CK(cudaMemset(dD, 0, 128 * N * 4));
      probe<<<1, 128, smem>>>(dA, dB, dD, M, N, KR, K, variant, nullptr, 1, majA, majB, 0);
      CK(cudaGetLastError());
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("M=%d N=%d variant=%d: CUDA error %s\n", M, N, variant, cudaGetErrorString(e)); return 1; }
      std::vector<float> hD(128 * N);
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      // expected rows
      std::vector<double> E(M * N);
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)tr(LA[m * KR + k]) * tr(LB[n * KR + k]); E[m * N + n] = s; }
      // identity lane mapping error, and discovered mapping for M=64
      double err = 0; for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) err = fmax(err, fabs(hD[m * N + n] - E[m * N + n]));
      { int nz = 0; for (auto v : hD) nz += (v != 0.f); printf("majA=%d majB=%d nonzero=%d ", majA, majB, nz); }
      printf("M=%d N=%d variant=%d: identity-lane max err %.3e  D[0][0..3]=%.4f %.4f %.4f %.4f  E=%.4f %.4f %.4f %.4f  D[1][0]=%.4f E[1][0]=%.4f", M, N, variant, err,
             hD[0], hD[1], hD[2], hD[3], E[0], E[1], E[2], E[3], hD[N], E[N]);
      { // try to explain D[0][0] as a dot product of some A row / B row pairing
        int bm = -1, bn = -1; for (int m = 0; m < M && bm < 0; ++m) for (int n = 0; n < N; ++n) if (fabs(E[m * N + n] - hD[0]) < 1e-4) { bm = m; bn = n; break; }
        printf("  D[0][0]==E[%d][%d]", bm, bn); }
      if (M == 64) {
        printf("  row->lane:");
        for (int m = 0; m < M; m += 8) { int best = -1; for (int l = 0; l < 128; ++l) { double e2 = 0; for (int n = 0; n < N; ++n) e2 = fmax(e2, fabs(hD[l * N + n] - E[m * N + n])); if (e2 < 1e-4) best = l; } printf(" %d->%d", m, best); }
      }
      printf("\n");
    }
    for (int reps : {64}) {
      probe<<<1, 128, smem>>>(dA, dB, dD, M, N, KR, K, 0, dC, reps, 1, 1, 0);
      CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
      printf("M=%d N=%d: %d MMAs in %lld cycles -> %.1f cyc/MMA (same accumulator, dependent)\n", M, N, reps * K / 8, c, (double)c / (reps * K / 8));
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  }
  return 0;
}

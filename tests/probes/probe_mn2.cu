// Probe 3: tf32 MN-major operands with swizzled canonical layouts (SW128 / SW64 / SW32) and arbitrary K-row shifts.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t ltype, uint32_t base_off) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(base_off & 7) << 49) | ((uint64_t)ltype << 61);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int a_floats, int b_floats, int M, int N, int K,
                                             int W, int KR, int ltype, int shiftA, int use_base_off) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  float* sA = (float*)smem; float* sB = sA + a_floats;
  int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < a_floats; i += 128) sA[i] = A[i];
  for (int i = tid; i < b_floats; i += 128) sB[i] = B[i];
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
  uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t aa = smem_u32(sA) + (k0 + shiftA) * W * 16, bb = smem_u32(sB) + k0 * W * 16;
      uint32_t boA = use_base_off ? ((aa >> 7) & 7) : 0;
      uint64_t ad = make_desc(aa, KR * W, 8 * W, ltype, boA);
      uint64_t bd = make_desc(bb, KR * W, 8 * W, ltype, 0);
      uint32_t acc = k0 > 0 ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                   :: "r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    unsigned long long spins = 0;
    while (!mbar_try(&bar, 0)) { if (++spins > (1ull << 26)) __trap(); }
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
static int swz(int c, int r, int W) { return W == 8 ? (c ^ (r & 7)) : (W == 4 ? (c ^ ((r >> 1) & 3)) : (c ^ ((r >> 2) & 1))); }
// element (m, k) -> float index in a buffer of [atoms][KR rows][W chunks][4]
static size_t pack_idx(int m, int k, int W, int KR) {
  int ch = m / 4, atom = ch / W, c = ch % W;
  return ((size_t)atom * KR * W + (size_t)k * W + swz(c, k, W)) * 4 + m % 4;
}
int main() {
  const int KR = 80, K = 64;
  for (int M : {128, 64}) for (int N : {32, 64}) for (int W : {8, 4, 2}) {
    int ltype = W == 8 ? 2 : (W == 4 ? 4 : 6);
    int atomsA = (M / 4 + W - 1) / W, atomsB = (N / 4 + W - 1) / W;
    size_t af = (size_t)atomsA * KR * W * 4, bf = (size_t)atomsB * KR * W * 4;
    std::vector<float> LA(M * KR), LB(N * KR), hA(af, 0.f), hB(bf, 0.f);
    srand(7);
    for (auto& v : LA) v = (rand() / (float)RAND_MAX) * 2 - 1;
    for (auto& v : LB) v = (rand() / (float)RAND_MAX) * 2 - 1;
    for (int m = 0; m < M; ++m) for (int k = 0; k < KR; ++k) hA[pack_idx(m, k, W, KR)] = LA[m * KR + k];
    for (int n = 0; n < N; ++n) for (int k = 0; k < KR; ++k) hB[pack_idx(n, k, W, KR)] = LB[n * KR + k];
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, af * 4)); CK(cudaMalloc(&dB, bf * 4)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), af * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), bf * 4, cudaMemcpyHostToDevice));
    size_t smem = (af + bf) * 4 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int shift : {0, 3, 8}) for (int ubo : {0, 1}) {
      if (shift == 0 && ubo) continue;
      CK(cudaMemset(dD, 0, 128 * N * 4));
      probe<<<1, 128, smem>>>(dA, dB, dD, (int)af, (int)bf, M, N, K, W, KR, ltype, shift, ubo);
      CK(cudaGetLastError());
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("M=%d N=%d W=%d: CUDA error %s\n", M, N, W, cudaGetErrorString(e)); return 1; }
      std::vector<float> hD(128 * N);
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      std::vector<double> E(M * N);
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)tr(LA[m * KR + k + shift]) * tr(LB[n * KR + k]); E[m * N + n] = s; }
      double err = 0; int nz = 0; for (auto v : hD) nz += v != 0.f;
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) err = fmax(err, fabs(hD[m * N + n] - E[m * N + n]));
      printf("M=%3d N=%2d SW%-3d shift=%d base_off=%d: nonzero=%5d identity-lane max err %.3e", M, N, W * 16, shift, ubo, nz, err);
      if (M == 64 && err > 1e-3) {
        printf("  row->lane:");
        for (int m = 0; m < M; m += 16) { int best = -1; for (int l = 0; l < 128; ++l) { double e2 = 0; for (int n = 0; n < N; ++n) e2 = fmax(e2, fabs(hD[l * N + n] - E[m * N + n])); if (e2 < 1e-4) best = l; } printf(" %d->%d", m, best); }
      }
      printf("\n");
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}

// Probe: validates the tcgen05 / TMEM / smem-descriptor conventions used by the implicit-GEMM conv kernel and
// measures the issue rate of narrow-N tf32 MMAs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
  d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity));
}

// mode: 0 = LBO is the K-chunk plane stride, SBO = 8 (128 B between 8-row groups); 1 = swapped
// A smem layout: [kchunk (K/4)][rows RA][4 floats]; B: [kchunk][N][4 floats]; D[m][n] = sum_k A[m+shift][k] * B[n][k]
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                    int RA, int N, int K, int shift, int mode, int reps, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = sA + (size_t)(K / 4) * RA * 4;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < (K / 4) * RA * 4; i += 128) sA[i] = A[i];
  for (int i = tid; i < (K / 4) * N * 4; i += 128) sB[i] = B[i];
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_tf32(128, N);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int k8 = 0; k8 < K / 8; ++k8) {
        uint32_t a_addr = smem_u32(sA) + (uint32_t)((2 * k8) * RA + shift) * 16;
        uint32_t b_addr = smem_u32(sB) + (uint32_t)((2 * k8) * N) * 16;
        uint64_t ad = mode == 0 ? make_desc(a_addr, RA, 8) : make_desc(a_addr, 8, RA);
        uint64_t bd = mode == 0 ? make_desc(b_addr, N, 8) : make_desc(b_addr, 8, N);
        mma_tf32(tmem_base, ad, bd, idesc, (r > 0 || k8 > 0) ? 1u : 0u);
      }
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // each warp reads its 32 lanes, N columns (in chunks of 32 / 16)
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256));
}

// conv-like issue pattern: G accumulator tiles x 25 taps per K=8 step; A window shifts per tap, B slab per tap.
// Descriptors are built incrementally (only the 14-bit start-address field changes) and the tap loop is unrolled so the
// issuing thread spends a couple of instructions per MMA.
__device__ __forceinline__ void mma_tf32_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, 1;\n\t}\n"
      :: "r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc));
}
template <int N, int G, int ORDER>
__global__ void __launch_bounds__(128) issue_kernel(int reps, int ncols, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  constexpr int PH = 128 * G + 276;
  uint32_t* w32 = reinterpret_cast<uint32_t*>(smem);
  constexpr int total_words = (2 * PH * 16 + 25 * 2 * N * 16) / 4;
  for (int i = tid; i < total_words; i += 128) w32[i] = 0x3c000000u + (i * 2654435761u & 0xFFFFF);
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_tf32(128, N);
  const uint32_t a_base = smem_u32(smem), b_base = a_base + 2 * PH * 16;
  if (tid == 0) {
    const uint64_t ad0 = make_desc(a_base, PH, 8), bd0 = make_desc(b_base, N, 8);
    const uint32_t alo = (uint32_t)ad0, ahi = (uint32_t)(ad0 >> 32), blo = (uint32_t)bd0, bhi = (uint32_t)(bd0 >> 32);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (ORDER == 0) {
#pragma unroll
        for (int t = 0; t < G; ++t)
#pragma unroll
          for (int tap = 0; tap < 25; ++tap)
            mma_tf32_lohi(tmem_base + t * N, alo + 128 * t + (tap / 5) * 68 + tap % 5, ahi, blo + tap * 2 * N, bhi, idesc);
      } else {
#pragma unroll
        for (int tap = 0; tap < 25; ++tap)
#pragma unroll
          for (int t = 0; t < G; ++t)
            mma_tf32_lohi(tmem_base + t * N, alo + 128 * t + (tap / 5) * 68 + tap % 5, ahi, blo + tap * 2 * N, bhi, idesc);
      }
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    *cycles = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(ncols));
}

template <int N, int G, int ORDER>
void run_issue(long long* dC) {
  int ncols = 32; while (ncols < N * G) ncols *= 2;
  size_t smem = (size_t)2 * (128 * G + 276) * 16 + 25 * 2 * N * 16 + 128;
  CK(cudaFuncSetAttribute(issue_kernel<N, G, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int grid : {1, 148}) {
    CK(cudaMemset(dC, 0, 8));
    issue_kernel<N, G, ORDER><<<grid, 128, smem>>>(40, ncols, dC);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
    int n_mma = 40 * 25 * G;
    printf("issue N=%3d G=%d order=%d grid=%3d: %6d MMAs in %8lld cyc -> %6.1f cyc/MMA (ideal %5.1f) smemB/MMA=%d\n", N, G, ORDER, grid,
           n_mma, c, (double)c / n_mma, 128.0 * N / 256, 128 * 32 + N * 32);
  }
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  const int RA = 160, K = 32;
  for (int N : {32, 64}) {
    std::vector<float> hA((K / 4) * RA * 4), hB((K / 4) * N * 4);
    srand(1);
    // logical A[row][k], B[n][k]
    std::vector<float> LA(RA * K), LB(N * K);
    for (auto& v : LA) v = (rand() / (float)RAND_MAX) * 2 - 1;
    for (auto& v : LB) v = (rand() / (float)RAND_MAX) * 2 - 1;
    for (int r = 0; r < RA; ++r) for (int k = 0; k < K; ++k) hA[((k / 4) * RA + r) * 4 + (k % 4)] = LA[r * K + k];
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[((k / 4) * N + n) * 4 + (k % 4)] = LB[n * K + k];
    float *dA, *dB, *dD; long long* dC;
    CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    size_t smem = (hA.size() + hB.size()) * 4 + 128;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int mode = 0; mode < 1; ++mode) for (int shift : {0, 3, 17}) {
      CK(cudaMemset(dD, 0, 128 * N * 4));
      probe_kernel<<<1, 128, smem>>>(dA, dB, dD, RA, N, K, shift, mode, 1, nullptr);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d mode=%d shift=%d: CUDA error %s\n", N, mode, shift, cudaGetErrorString(e)); return 1; }
      std::vector<float> hD(128 * N);
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      double e_full = 0, e_tr = 0, e_rn = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double s = 0, st = 0, sr = 0;
        for (int k = 0; k < K; ++k) {
          float a = LA[(m + shift) * K + k], b = LB[n * K + k];
          s += (double)a * b; st += (double)tf32_trunc(a) * tf32_trunc(b); sr += (double)tf32_rn(a) * tf32_rn(b);
        }
        double v = hD[m * N + n];
        e_full = fmax(e_full, fabs(v - s)); e_tr = fmax(e_tr, fabs(v - st)); e_rn = fmax(e_rn, fabs(v - sr));
      }
      printf("N=%d mode=%d shift=%2d: max|D-fp32|=%.3e  max|D-tf32trunc|=%.3e  max|D-tf32rn|=%.3e\n", N, mode, shift, e_full, e_tr, e_rn);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  }
  {
    long long* dC; CK(cudaMalloc(&dC, 8 * 512));
    run_issue<16, 1, 0>(dC); run_issue<16, 8, 0>(dC); run_issue<16, 8, 1>(dC);
    run_issue<32, 1, 0>(dC); run_issue<32, 8, 0>(dC); run_issue<32, 8, 1>(dC);
    run_issue<64, 1, 0>(dC); run_issue<64, 8, 0>(dC); run_issue<64, 8, 1>(dC);
    run_issue<128, 1, 0>(dC); run_issue<128, 4, 0>(dC); run_issue<128, 4, 1>(dC);
    run_issue<256, 1, 0>(dC); run_issue<256, 2, 1>(dC);
  }
  return 0;
}

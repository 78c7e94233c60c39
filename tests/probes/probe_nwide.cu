// Probe (round 1, session 2): MMA cost vs N for back-to-back MMAs on ONE accumulator (ORDER 0) vs alternating accumulators (ORDER 1).
// Probe: issue rate of kind::f16 (K=16) vs kind::tf32 (K=8) MMAs with the conv kernel's operand pattern (same smem bytes per MMA).
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

// conv-like issue pattern: G accumulator tiles x 25 taps per K=8 step; A window shifts per tap, B slab per tap.
// Descriptors are built incrementally (only the 14-bit start-address field changes) and the tap loop is unrolled so the
// issuing thread spends a couple of instructions per MMA.
__device__ __forceinline__ void mma_tf32_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, 1;\n\t}\n"
      :: "r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc));
}
__device__ __forceinline__ void mma_f16_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t}\n"
      :: "r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc));
}
template <int N, int G, int ORDER, int F16>
__global__ void __launch_bounds__(128) issue_kernel(int reps, int ncols, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  constexpr int PH = 128 * G + 276;
  uint32_t* w32 = reinterpret_cast<uint32_t*>(smem);
  constexpr int total_words = (2 * PH * 16 + 25 * 2 * N * 16) / 4;
  for (int i = tid; i < total_words; i += 128) w32[i] = (F16 ? 0x3c003c00u : 0x3c000000u + (i * 2654435761u & 0xFFFFF));
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
  const uint32_t idesc = F16 ? ((1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)) : make_idesc_tf32(128, N);
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
            { if (F16) mma_f16_lohi(tmem_base + t * N, alo + 128 * t + (tap / 5) * 68 + tap % 5, ahi, blo + tap * 2 * N, bhi, idesc); else mma_tf32_lohi(tmem_base + t * N, alo + 128 * t + (tap / 5) * 68 + tap % 5, ahi, blo + tap * 2 * N, bhi, idesc); }
      } else {
#pragma unroll
        for (int tap = 0; tap < 25; ++tap)
#pragma unroll
          for (int t = 0; t < G; ++t)
            { if (F16) mma_f16_lohi(tmem_base + t * N, alo + 128 * t + (tap / 5) * 68 + tap % 5, ahi, blo + tap * 2 * N, bhi, idesc); else mma_tf32_lohi(tmem_base + t * N, alo + 128 * t + (tap / 5) * 68 + tap % 5, ahi, blo + tap * 2 * N, bhi, idesc); }
      }
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    *cycles = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(ncols));
}

template <int N, int G, int ORDER, int F16>
void run_issue(long long* dC) {
  int ncols = 32; while (ncols < N * G) ncols *= 2;
  size_t smem = (size_t)2 * (128 * G + 276) * 16 + 25 * 2 * N * 16 + 128;
  CK(cudaFuncSetAttribute(issue_kernel<N, G, ORDER, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int grid : {1, 148}) {
    CK(cudaMemset(dC, 0, 8));
    issue_kernel<N, G, ORDER, F16><<<grid, 128, smem>>>(40, ncols, dC);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
    int n_mma = 40 * 25 * G;
    printf("%s issue N=%3d G=%d order=%d grid=%3d: %6d MMAs in %8lld cyc -> %6.1f cyc/MMA (ideal %5.1f) smemB/MMA=%d\n", F16 ? "f16 K=16" : "tf32 K=8", N, G, ORDER, grid,
           n_mma, c, (double)c / n_mma, 128.0 * N / 256, 128 * 32 + N * 32);
  }
}



// The quint-tap conv kernel's exact operand pattern: A = [hi/lo][4 k-chunk planes][872 rows] (LBO = 872), B slabs of 160 rows,
// per tile 2 K groups x 5 kernel rows x 3 MMAs into ONE 160-column accumulator; tiles advance by 124 rows in a 744-row ring.
template <int MODE>
__global__ void __launch_bounds__(128) quint_kernel(int reps, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  constexpr int ROWS = 872, NQ = 160;
  uint32_t* w32 = reinterpret_cast<uint32_t*>(smem);
  constexpr int a_bytes = 8 * ROWS * 16, w_bytes = 2 * 5 * 2 * 2 * NQ * 16;
  for (int i = tid; i < (a_bytes + w_bytes) / 4; i += 128) w32[i] = 0x3c003c00u;
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(NQ >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_base = smem_u32(smem), b_base = a_base + a_bytes;
  if (tid == 0) {
    const uint64_t ad0 = make_desc(a_base, ROWS, 8), bd0 = make_desc(b_base, NQ, 8);
    const uint32_t alo = (uint32_t)ad0, ahi = (uint32_t)(ad0 >> 32), blo = (uint32_t)bd0, bhi = (uint32_t)(bd0 >> 32);
    long long t0 = clock64();
    int n = 0;
    for (int r = 0; r < reps; ++r)
      for (int t = 0; t < 6; ++t) {
        const uint32_t dcol = tmem_base + (MODE == 1 ? 0 : (t % 3) * NQ);
        for (int cg = 0; cg < 2; ++cg)
#pragma unroll
          for (int ky = 0; ky < 5; ++ky) {
            int row = t * 124 + ky * 68;
            if (row >= 744) row -= 744;
            const uint32_t a_hi_op = alo + (2 * cg) * ROWS + row, a_lo_op = alo + (4 + 2 * cg) * ROWS + row;
            const uint32_t b_hi_op = blo + ((cg * 5 + ky) * 2 + 0) * 2 * NQ, b_lo_op = blo + ((cg * 5 + ky) * 2 + 1) * 2 * NQ;
            mma_f16_lohi(dcol, a_hi_op, ahi, b_hi_op, bhi, idesc);
            mma_f16_lohi(dcol, a_hi_op, ahi, b_lo_op, bhi, idesc);
            mma_f16_lohi(dcol, a_lo_op, ahi, b_hi_op, bhi, idesc);
            n += 3;
          }
      }
    commit(&bar);
    mbar_wait(&bar, 0);
    cycles[0] = clock64() - t0;
    cycles[1] = n;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512));
}

void run_quint(long long* dC) {
  size_t smem = 8 * 872 * 16 + 2 * 5 * 2 * 2 * 160 * 16 + 128;
  CK(cudaFuncSetAttribute(quint_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int grid : {1, 148}) {
    CK(cudaMemset(dC, 0, 16));
    quint_kernel<0><<<grid, 128, smem>>>(10, dC);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long c[2]; CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
    printf("quint pattern (LBO 872, N=160, 3 MMAs per kernel row) grid=%3d: %lld MMAs in %lld cyc -> %.1f cyc/MMA\n", grid, c[1], c[0], (double)c[0] / c[1]);
  }
}

int main() {
  long long* dC; CK(cudaMalloc(&dC, 8 * 512));
  run_quint(dC);
  // ORDER 0: all 25 taps of a tile back to back into ONE accumulator; ORDER 1: consecutive MMAs alternate over G accumulators
  run_issue<32, 4, 0, 1>(dC); run_issue<32, 4, 1, 1>(dC);
  run_issue<64, 4, 0, 1>(dC); run_issue<64, 4, 1, 1>(dC);
  run_issue<96, 4, 0, 1>(dC); run_issue<96, 4, 1, 1>(dC);
  run_issue<128, 4, 0, 1>(dC); run_issue<128, 4, 1, 1>(dC);
  run_issue<160, 3, 0, 1>(dC); run_issue<160, 3, 1, 1>(dC);
  run_issue<192, 2, 0, 1>(dC); run_issue<192, 2, 1, 1>(dC);
  run_issue<256, 2, 0, 1>(dC); run_issue<256, 2, 1, 1>(dC);
  return 0;
}

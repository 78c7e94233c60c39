// Probe: MMA issue/execute rate of the ring-buffered wgrad's operand pattern (no loaders).
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


// ring-wgrad issue pattern: per step 5 accumulators (columns 96*ky) x 4 K-steps x 2 (B hi rows, B lo rows); A = 8 dense chunks of
// 128 rows, B = ring of 57 chunks of 192 rows.  VAR 0: kernel order (ky outer), 1: ky inner, 2: N=192 single MMA per (ky,ks)
__device__ __forceinline__ void mma_f16_acc(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t}\n"
      :: "r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc));
}
template <int VAR, int RND, int STORES>
__global__ void __launch_bounds__(256) ring_kernel(int steps, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  constexpr int A_ROWS = 128, B_ROWS = 192, RING = 56;
  uint32_t* w32 = reinterpret_cast<uint32_t*>(smem);
  constexpr int total_words = (8 * A_ROWS * 16 + (RING + 1) * B_ROWS * 16) / 4;
  for (int i = tid; i < total_words; i += 256) { uint32_t h = i * 2654435761u; w32[i] = RND ? ((h & 0x83ff83ffu) | 0x38003800u) : 0x3c003c00u; }
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
  const uint32_t idesc96 = (1u << 4) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc192 = (1u << 4) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_base = smem_u32(smem), b_base = a_base + 8 * A_ROWS * 16;
  if (tid == 0) {
    const uint64_t ad0 = make_desc(a_base, A_ROWS, 8), bd0 = make_desc(b_base, B_ROWS, 8);
    const uint32_t alo = (uint32_t)ad0, ahi = (uint32_t)(ad0 >> 32), blo = (uint32_t)bd0, bhi = (uint32_t)(bd0 >> 32);
    long long t0 = clock64();
    for (int t = 0; t < steps; ++t) {
      const int c0 = 8 * (t % 7);
      uint32_t boff[5][4];
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        int slot = c0 + (5 * RING - 9 * ky) % RING;
        if (slot >= RING) slot -= RING;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) { boff[ky][ks] = blo + slot * B_ROWS; slot += 2; if (slot >= RING) slot -= RING; }
      }
      if (VAR == 0) {
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_f16_acc(tmem_base + 96 * ky, alo + 2 * ks * A_ROWS, ahi, boff[ky][ks], bhi, idesc96);
            mma_f16_acc(tmem_base + 96 * ky, alo + 2 * ks * A_ROWS, ahi, boff[ky][ks] + 96, bhi, idesc96);
          }
      } else if (VAR == 1) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int hl = 0; hl < 2; ++hl)
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
              mma_f16_acc(tmem_base + 96 * ky, alo + 2 * ks * A_ROWS, ahi, boff[ky][ks] + 96 * hl, bhi, idesc96);
      } else {
#pragma unroll
        for (int ky = 0; ky < 2; ++ky)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_f16_acc(tmem_base + 192 * ky, alo + 2 * ks * A_ROWS, ahi, boff[ky][ks], bhi, idesc192);
      }
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    *cycles = clock64() - t0;
  }
  else if (STORES && warp >= 4) {
    // background shared-memory store traffic like the loaders': each of 4 warps writes 16-byte rows into a scratch area
    uint4* scr = reinterpret_cast<uint4*>(smem + 8 * 128 * 16 + 57 * 192 * 16 + 128);
    volatile long long* cyc = cycles;
    for (int it = 0; it < steps * STORES; ++it) {
#pragma unroll
      for (int r = 0; r < 16; ++r) scr[(r * 128 + (tid - 128)) & 1023] = make_uint4(it, r, tid, 0);
      __syncwarp();
      if (*cyc != 0) break;
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512));
}
template <int VAR, int RND, int STORES>
void run_ring(long long* dC) {
  size_t smem = 8 * 128 * 16 + 57 * 192 * 16 + 128 + 16384;
  CK(cudaFuncSetAttribute(ring_kernel<VAR, RND, STORES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int grid : {1, 148}) {
    CK(cudaMemset(dC, 0, 8));
    ring_kernel<VAR, RND, STORES><<<grid, 256, smem>>>(100, dC);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
    int n_mma = 100 * (VAR == 2 ? 8 : 40);
    printf("ring pattern RND=%d STORES=%d VAR=%d grid=%3d: %5d MMAs in %8lld cyc -> %6.1f cyc/MMA\n", RND, STORES, VAR, grid, n_mma, c, (double)c / n_mma);
  }
}
int main() {
  long long* dC; CK(cudaMalloc(&dC, 8 * 512));
  run_ring<0, 0, 0>(dC); run_ring<0, 1, 0>(dC); run_ring<0, 1, 20>(dC); run_ring<0, 0, 20>(dC); run_ring<0, 1, 200>(dC);
  return 0;
}

// Probe (round 2): tcgen05 cta_group::2 conventions for the quint-tap conv kernel -- one M=256 x N=160 x K=16 kind::f16 MMA issued
// by the leader CTA of a 2-CTA cluster.  Checks against a host reference:
//   * tcgen05.alloc.cta_group::2 executed by one warp of EACH CTA;
//   * A: each CTA's own 128 rows at the SAME shared-memory offset; B: each CTA holds N/2 = 80 rows (rank 0 -> columns 0..79);
//   * D: each CTA's TMEM receives its 128 rows x 160 columns;
//   * operand-ready handshake: every thread of both CTAs arrives on the LEADER's mbarrier (mapa + arrive.release.cluster);
//   * tcgen05.commit.cta_group::2 ... multicast::cluster arrives on the same barrier offset in both CTAs;
//   * MODE 1: x staged in TMEM with tcgen05.cp.cta_group::2.128x256b and the MMA taking A from TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_2cta.bin probe_2cta.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

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
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  unsigned long long spins = 0;
  while (!mbar_try(bar, parity)) if (++spins > (1ull << 26)) { printf("mbar timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); }
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
}

constexpr int N = 160, NH = 80;

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe_kernel(float* out /*[2][128][160]*/) {
  __shared__ __align__(128) uint8_t a_s[2 * 128 * 16];     // [k-chunk][row][16 B]
  __shared__ __align__(128) uint8_t b_s[2 * NH * 16];      // [k-chunk][80 rows][16 B]: this CTA's half of B
  __shared__ uint64_t ready_bar, done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // operands: A_rank[row][k] = (row % 7) + (k % 3) + rank,  B[n][k] = ((3 n + k) % 5) - 2   (exact in fp16, sums exact in fp32)
  for (int e = tid; e < 128 * 16; e += 128) {
    const int row = e / 16, k = e % 16;
    reinterpret_cast<__half*>(a_s)[((k / 8) * 128 + row) * 8 + (k % 8)] = __float2half((float)(row % 7 + k % 3 + (int)rank));
  }
  for (int e = tid; e < NH * 16; e += 128) {
    const int nl = e / 16, k = e % 16, n = nl + NH * (int)rank;
    reinterpret_cast<__half*>(b_s)[((k / 8) * NH + nl) * 8 + (k % 8)] = __float2half((float)((3 * n + k) % 5 - 2));
  }
  if (tid == 0) {
    mbar_init(&ready_bar, 256);
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // barriers + TMEM ready everywhere
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  // "loaders": operands written -> visible to the async proxy -> arrive on the LEADER's barrier
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_arrive_cluster(&ready_bar, 0);
  if (rank == 0 && warp == 1) {
    mbar_wait(&ready_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    if (elected) {
      const uint64_t ad = make_desc(smem_u32(a_s), 128, 8), bd = make_desc(smem_u32(b_s), NH, 8);
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      if (MODE == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 1;\n\tsetp.eq.b32 p, 0, 1;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc) : "memory");
      } else {
        asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" :: "r"(tmem_base + 480u), "l"(ad) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 1;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                     :: "r"(tmem_base), "r"(tmem_base + 480u), "l"(bd), "r"(idesc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   :: "r"(smem_u32(&done_bar)), "h"((uint16_t)3) : "memory");
    }
    __syncwarp();
  }
  mbar_wait(&done_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[((size_t)rank * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
}

template <int MODE>
int run() {
  float* d;
  CK(cudaMalloc(&d, 2 * 128 * N * sizeof(float)));
  CK(cudaMemset(d, 0xff, 2 * 128 * N * sizeof(float)));
  probe_kernel<MODE><<<2, 128>>>(d);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> h(2 * 128 * N);
  CK(cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int rank = 0; rank < 2; ++rank)
    for (int row = 0; row < 128; ++row)
      for (int n = 0; n < N; ++n) {
        float ref = 0.f;
        for (int k = 0; k < 16; ++k) ref += (float)(row % 7 + k % 3 + rank) * (float)((3 * n + k) % 5 - 2);
        const float got = h[((size_t)rank * 128 + row) * N + n];
        if (got != ref && bad++ < 8) printf("  MODE %d mismatch rank %d row %d col %d: got %g want %g\n", MODE, rank, row, n, got, ref);
      }
  printf("MODE %d (%s): %s (%d mismatches of %d)\n", MODE, MODE ? "A staged in TMEM by tcgen05.cp.cta_group::2" : "A from shared memory",
         bad ? "FAIL" : "OK", bad, 2 * 128 * N);
  CK(cudaFree(d));
  return bad;
}

int main() {
  int bad = run<0>();
  bad += run<1>();
  return bad ? 1 : 0;
}

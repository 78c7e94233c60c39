// Probe (round 2): the quint-tap kernel's resident-weight prologue as TMA tensor copies.  The fp16 weight slab of a layer is
// the contiguous 5-D array [K group cg][ky][kx][j = k-chunk*2 + hi/lo][256 halves = 32 output channels x 8 input channels]; the
// kernel wants, per (cg, ky, hi/lo, k-chunk), the 160 rows [kx*32 + co] of 16 bytes.  One box {256, 1, 5, 1, 1} per (cg, ky, j)
// lands exactly that slab: 20 * ngroups cp.async.bulk.tensor.5d copies instead of 100 loads + index arithmetic + stores per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_tma_weights.bin probe_tma_weights.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int NQ = 160;

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmap, int ngroups, const uint4* __restrict__ src,
                                                    int* __restrict__ mismatches) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint4* w_s = reinterpret_cast<uint4*>(smem);             // [cg][ky][hl][kc][160] x 16 B
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)ngroups * 20u * 2560u;
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" :: "r"(smem_u32(&bar)), "r"(bytes) : "memory");
    for (int cg = 0; cg < ngroups; ++cg)
      for (int ky = 0; ky < 5; ++ky)
        for (int j = 0; j < 4; ++j) {
          const int kc = j >> 1, hl = j & 1;
          uint4* dst = w_s + ((((cg * 5 + ky) * 2 + hl) * 2 + kc) * NQ);
          asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                       :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(j), "r"(0), "r"(ky), "r"(cg),
                          "r"(smem_u32(&bar)) : "memory");
        }
  }
  // wait
  uint32_t ok = 0;
  unsigned long long spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (++spins > (1ull << 26)) { printf("timeout\n"); __trap(); }
  }
  // compare with the hand re-layout of the kernel's current prologue
  int bad = 0;
  for (int e = tid; e < ngroups * 25 * 2 * 64; e += 128) {
    const int row64 = e & 63, kc = (e >> 6) & 1, tap = (e >> 7) % 25, cg = e / (128 * 25);
    const int hl = row64 >> 5, co = row64 & 31, ky = tap / 5, kx = tap - 5 * ky;
    const uint4 want = src[e];
    const uint4 got = w_s[((((cg * 5 + ky) * 2 + hl) * 2 + kc) * NQ) + kx * 32 + co];
    if (want.x != got.x || want.y != got.y || want.z != got.z || want.w != got.w) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int ngroups = 2;
  const size_t n16 = (size_t)ngroups * 25 * 2 * 64;        // 16-byte elements
  std::vector<uint32_t> h(n16 * 4);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)(i * 2654435761u);
  uint4* d;
  int* dm;
  CK(cudaMalloc(&d, n16 * 16));
  CK(cudaMalloc(&dm, 4));
  CK(cudaMemcpy(d, h.data(), n16 * 16, cudaMemcpyHostToDevice));
  CK(cudaMemset(dm, 0, 4));
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
  if (!encode || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tmap;
  const cuuint64_t gdim[5] = {256, 4, 5, 5, (cuuint64_t)ngroups};
  const cuuint64_t gstr[4] = {512, 2048, 10240, 51200};    // bytes, dims 1..4
  const cuuint32_t box[5] = {256, 1, 5, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  const size_t smem = (size_t)ngroups * 5 * 2 * 2 * NQ * 16;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(tmap, ngroups, d, dm);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int bad = -1;
  CK(cudaMemcpy(&bad, dm, 4, cudaMemcpyDeviceToHost));
  printf("TMA weight prologue: %s (%d mismatching 16-byte rows of %zu)\n", bad == 0 ? "OK" : "FAIL", bad, n16);
  return bad != 0;
}

// tcgen05 / TMEM / mbarrier helpers (inline PTX for sm_100a).  Conventions verified by tests/probes/probe_tcgen05.cu:
//  * K-major, no-swizzle smem operand: rows of 16 bytes (4 tf32), 8-row core matrices contiguous (SBO = 8 x 16 B),
//    the two 16-byte K-chunks of one K=8 MMA are LBO apart; the descriptor start address may point at any 16-byte row;
//  * kind::tf32 TRUNCATES fp32 operands to 19 bits;
//  * a single thread can issue one M=128 x K=8 MMA every ~48 cycles for N <= 64 (operand-fetch floor), N/4 beyond.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace op3 {

#define OP3_PREC_FP32 0      /* CUDA-core fp32 kernels (exact mode)          */
#define OP3_PREC_TF32X3 1    /* tcgen05 3xTF32 split (fp32-grade, parity)    */
#define OP3_PREC_TF32 2      /* tcgen05 1xTF32 (speed mode)                  */
#define OP3_PREC_F16X3 3     /* tcgen05 fp16 hi/lo split with per-image power-of-two operand scaling: the same 11-bit
                                significands as 3xTF32 at half the operand bytes (forward/dgrad; wgrad stays 3xTF32) */

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_lo(float x) { return x - tf32_hi(x); }
// round to nearest tf32 (ties away from zero in magnitude); result has its low 13 mantissa bits clear
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
  d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell); layout_type 0 = no swizzle, base_offset 0
  return d;
}
// instruction descriptor: D fp32, A/B tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}\n"
      :: "r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16: A/B fp16 (format 0), D fp32, K = 16 per MMA (two 16-byte K-chunks of 8 halves, LBO apart)
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
      :: "r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
// accumulating form without the predicate set-up (shorter issue sequence)
__device__ __forceinline__ void mma_f16_acc(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                            uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t}\n"
      :: "r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
// A operand in TENSOR MEMORY (tests/probes/probe_tmem_a.cu: bit-identical to the smem-A form): `tcgen05.cp.128x256b` copies a
// 128-row x 32-byte K-major smem tile (one K=16 fp16 step, the same descriptor an MMA would take) into 8 TMEM columns, in
// issue order with the MMAs; MMAs that name it fetch only B from shared memory.
__device__ __forceinline__ void tc_cp_a128(uint32_t a_tmem, uint32_t alo, uint32_t ahi) {
  asm volatile("{\n\t.reg .b64 da;\n\tmov.b64 da, {%1, %2};\n\ttcgen05.cp.cta_group::1.128x256b [%0], da;\n\t}\n"
               :: "r"(a_tmem), "r"(alo), "r"(ahi) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_f16_ts_acc(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, 1;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
// ---- cta_group::2: one MMA over a CTA PAIR (tests/probes/probe_2cta.cu).  M = 256: each CTA supplies its own 128 A rows
// at the same shared-memory offset and receives its 128 accumulator rows in its own TMEM; B has N rows in total, each CTA
// supplies N/2 of them (rank 0 the first half) from the same offset.  Issued by the leader CTA (cluster rank 0) only.
__device__ __forceinline__ void mma2_f16(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n"
      :: "r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma2_f16_acc(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                             uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n\t}\n"
      :: "r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void tc2_cp_a128(uint32_t a_tmem, uint32_t alo, uint32_t ahi) {     // in both CTAs, each from its own smem
  asm volatile("{\n\t.reg .b64 da;\n\tmov.b64 da, {%1, %2};\n\ttcgen05.cp.cta_group::2.128x256b [%0], da;\n\t}\n"
               :: "r"(a_tmem), "r"(alo), "r"(ahi) : "memory");
}
__device__ __forceinline__ void mma2_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma2_f16_ts_acc(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, 1;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the preceding MMAs / copies completed
__device__ __forceinline__ void tc2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// arrive (release at cluster scope) on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {  // one whole warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}

// Power-of-two scale that brings a tensor whose largest magnitude has the fp32 bit pattern `maxbits` into [2^13, 2^14)
// (fp16 overflows at 2^16; products are accumulated in fp32).  Zero / denormal maxima get scale 1.
__host__ __device__ __forceinline__ uint32_t f16_scale_bits(uint32_t maxbits) {
  const uint32_t e = (maxbits >> 23) & 0xFFu;
  if (e == 0u) return 0x3F800000u;
  uint32_t se = 267u - e;                       // 127 + 13 - (e - 127)
  if (se > 254u) se = 254u;
  if (se < 1u) se = 1u;
  return se << 23;
}
// (a, b) -> packed fp16 pairs hi = rn(a), lo = rn(a - hi): a_hi + a_lo carries 22 significand bits (exact residual)
__device__ __forceinline__ void f16_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// f16_split2(a * sc, b * sc, ...) with the scale and the residual as packed fp32 pairs (FMUL2 / FADD2: 6 instead of 8
// instructions per pair when a and b sit in an aligned register pair, e.g. the .x/.y or .z/.w of a loaded float4); same
// roundings, bit-identical results
__device__ __forceinline__ void f16_split2_scaled(float a, float b, float sc, uint32_t& hi, uint32_t& lo) {
  float sa, sb;
  asm("{\n\t.reg .b64 p, s, r;\n\tmov.b64 p, {%2, %3};\n\tmov.b64 s, {%4, %4};\n\tmul.rn.f32x2 r, p, s;\n\tmov.b64 {%0, %1}, r;\n\t}"
      : "=f"(sa), "=f"(sb) : "f"(a), "f"(b), "f"(sc));
  const __half2 h = __floats2half2_rn(sa, sb);
  const float2 hf = __half22float2(h);
  float ra, rb;
  asm("{\n\t.reg .b64 p, q, r;\n\tmov.b64 p, {%2, %3};\n\tmov.b64 q, {%4, %5};\n\tsub.rn.f32x2 r, p, q;\n\tmov.b64 {%0, %1}, r;\n\t}"
      : "=f"(ra), "=f"(rb) : "f"(sa), "f"(sb), "f"(hf.x), "f"(hf.y));
  const __half2 l = __floats2half2_rn(ra, rb);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  unsigned long long spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1ull << 28)) __trap();
  }
}

// elect.sync: true in exactly one lane of the (fully converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
// Warp-convergent wait: ONE lane polls the barrier, the others park at the warp barrier.  With every lane polling, the
// eight-plus waiting warps of a warp-specialised kernel keep the shared-memory pipe busy with try_wait probes next to the
// tensor core's operand fetches.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}
// 32 lanes x 16 consecutive columns: thread i of the warp receives lane (taddr.lane + i), columns [col, col+16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ const float4* tc_src_chunk_ptr(const ConvInput& in, int c, int n, int D) {
  int base = 0;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    if (s < in.nsrc) {
      int nc = in.src[s].nchunks;
      if (c < base + nc) {
        int idx = in.src[s].div > 0 ? n / in.src[s].div : 0;
        return in.src[s].ptr + ((size_t)idx * (in.src[s].pitch ? in.src[s].pitch : nc) + (c - base)) * D * D;
      }
      base += nc;
    }
  }
  return nullptr;
}

// conv5x5_tc.cu
int tc_nout(int cout);
bool tc_quint_fits(int D, int total_chunks);    // the quint-tap fp16 kernel applies (see conv5x5_tc.cu)
size_t conv5x5_tc_weight_floats(int cin_pad, int cout);
size_t conv5x5_tc_f16_slab_floats(int cin_pad, int nout);   // one fp16 slab (+ its scale word), in floats
size_t tc_scratch_floats(int N);                             // 8 slots of per-image operand maxima (fp16 mode)
inline size_t tc_slot_floats(int N) { return ((size_t)N + 3) & ~(size_t)3; }
inline uint32_t* tc_slot(float* base, int N, int slot) { return reinterpret_cast<uint32_t*>(base + slot * tc_slot_floats(N)); }
inline const uint32_t* tc_slot(const float* base, int N, int slot) { return reinterpret_cast<const uint32_t*>(base + slot * tc_slot_floats(N)); }
int conv5x5_tc_forward(cudaStream_t st, int precision, int N, int D, const ConvInput& in, int cout, const float* wt,
                       const float* bias, float4* out, int epi, const float4* elu_ref, float* tc_scratch);

// conv5x5_wgrad_tc.cu
size_t conv5x5_wgrad_tc_scratch_floats(int cin_pad, int cout);
// dpre_maxbits: per-image maxima of dpre if its producer recorded them (fp16 mode), else NULL
int conv5x5_wgrad_tc(cudaStream_t st, int precision, int N, int D, const ConvInput& in, int cout, const float4* dpre,
                     float* scratch, float* dwt, float* dbias, const uint32_t* dpre_maxbits);
int conv5x5_wgrad_auto(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float4* dpre, float* scratch,
                       float* dwt, float* dbias, const uint32_t* dpre_maxbits);
// tc_scratch: tc_scratch_floats(N) floats of caller scratch (only touched in the fp16 mode)
int conv5x5_auto(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float* wt_ffma, const float* wt_tc,
                 const float* bias, float4* out, int epi, const float4* elu_ref, float* tc_scratch);

}  // namespace op3

// tcgen05 implicit-GEMM 5x5 convolution (forward; also dgrad through transformed weights) for sm_100a.
//
//   D[pixels, cout] += A[pixels, (tap, ci)] * W[(tap, ci), cout]      M = 128 pixels / MMA, N = cout, K = 8 (tf32)
//
// * The activation halo of a group of G 128-pixel tiles sits in shared memory as [k-chunk][padded-linear pixel][4 ch]
//   (16-byte rows, no swizzle).  In padded-linear pixel space (row pitch D+4) every tap is a pure shift, so the A
//   operand of tap (ky,kx) is the SAME buffer with the descriptor start address advanced by (ky*(D+4)+kx)*16 bytes
//   (verified on B200 by tests/probes/probe_tcgen05.cu): no im2col copy is ever made; the 4 pad columns of each
//   row produce garbage outputs that the epilogue drops.
// * Persistent, warp-specialised: one CTA per SM loops over (image, tile-group) work items.
//     warps 0-7  epilogue   TMEM -> registers -> bias/ELU (or x ELU' for dgrad) -> coalesced float4 stores; two warps
//                           per TMEM lane quadrant alternate over the tiles of an item
//     warp  8    MMA issuer one elected thread issues every tcgen05.mma; tcgen05.commit releases smem stages and
//                           publishes accumulators
//     warps 9-12 loaders    cp.async (zero fill = conv padding) -> round to tf32 in place -> fence.proxy.async -> mbarrier
//   Accumulators are double buffered in TMEM (2 x G tiles x N columns <= 512) so the epilogue of item i overlaps the
//   MMAs of item i+1; operand stages form an NSTAGE ring that runs across item boundaries.
// * kind::tf32 truncates fp32 operands to their top 19 bits (measured), which biases products; all operands are
//   therefore rounded to nearest tf32 before the tensor core sees them (weights in op3_prepare, activations by the
//   loaders).  3xTF32 ("parity mode"): a_hi*[w_hi;w_lo] is ONE MMA with N = 2*cout (the halves are summed in the
//   epilogue), a_lo*w_hi a second one.  1xTF32 ("speed mode") issues a_hi*w_hi only.  Narrow-N MMAs are operand-fetch
//   bound (~48 cycles for any N <= 64, measured), so the N = 2*cout stacking is free.
#include <cuda.h>
#include <cstdlib>
#include <unordered_map>
#include "common.cuh"
#include "tc_common.cuh"

namespace op3 {

namespace {

constexpr int EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP0 = 9, LOAD_WARPS = 7, LOAD_THREADS = LOAD_WARPS * 32;
constexpr int TC_THREADS = (LOAD_WARP0 + LOAD_WARPS) * 32;
constexpr int TC_LD_U = 4;   // pixels per loader thread and pass in the fp16 mode (x 4 channel chunks = 16 loads in flight)

// MODE: 0 = 1xTF32, 1 = 3xTF32 (hi/lo split), 2 = fp16 hi/lo split (K = 16 per MMA, per-image operand scaling)
#ifdef OP3_TC_TRACE
// debug timeline of the MMA issuer thread: [sm][0] cycles waiting for operand stages, [1] waiting for a free
// accumulator buffer, [2] total cycles in the item loop, [3] number of MMAs issued
__device__ unsigned long long g_tc_trace[256][4];
__device__ unsigned long long g_tcq_epi[8];          // quint epilogue warp 0: acc wait, loads+shuffles, barrier, tiles, fix-up, ELU+stores, barrier
#endif

template <int NOUT, int MODE>
struct TcCfg {
  static constexpr bool SPLIT3 = MODE != 0;
  static constexpr bool F16 = MODE == 2;
  static constexpr int NB = SPLIT3 ? 2 * NOUT : NOUT;            // B rows per tap slab = TMEM columns per tile
  static constexpr int GMAX = 256 / NB;                          // tiles per item: two accumulator buffers in 512 columns
  static constexpr int NSTAGE = SPLIT3 ? 2 : (NOUT == 64 ? 2 : 3);
  static_assert(GMAX >= 1 && 2 * GMAX * NB <= 512, "accumulators exceed TMEM");
};

template <int NOUT, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv5x5_tc_kernel(ConvInput in, int N, int D, int G, const float* __restrict__ wt, const float* __restrict__ bias,
                  int cout_real, float4* __restrict__ out, int out_planes, int out_chunk0, int out_chunks, int epi,
                  const float4* __restrict__ ref, const uint32_t* __restrict__ maxbits, const float* __restrict__ wscale) {
  using Cfg = TcCfg<NOUT, MODE>;
  constexpr int NB = Cfg::NB, NSTAGE = Cfg::NSTAGE;
  constexpr bool SPLIT3 = Cfg::SPLIT3, F16 = Cfg::F16;
  const int P = D + 4;                                   // padded row pitch
  const int PH = 128 * G + 4 * P + 4;                    // halo pixels of one group
  const int tiles_img = (D * P + 127) / 128;
  const int groups_img = (tiles_img + G - 1) / G;
  const int n_items = N * groups_img;
  const int ngroups = F16 ? (in.total_chunks + 3) / 4 : (in.total_chunks + 1) / 2;   // K groups of 16 / 8 channels
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;   // provably warp-uniform roles

  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int a_bytes = 2 * PH * 16;
  const int w_bytes = 25 * 2 * NB * 16;
  const int stage_bytes = a_bytes * (SPLIT3 ? 2 : 1) + w_bytes;
  __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[NOUT];

  if (tid < NOUT) s_bias[tid] = (bias != nullptr && tid < cout_real) ? bias[tid] : 0.f;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], LOAD_THREADS); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EPI_WARPS * 32); }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp >= LOAD_WARP0) {
    // ===================== loaders =====================
    const int lt = tid - LOAD_WARP0 * 32;
    int it = 0;                                          // running stage counter across items
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int n = item / groups_img, grp = item % groups_img;
      const int q0 = grp * G * 128;
      for (int cg = 0; cg < ngroups; ++cg, ++it) {
        const int s = it % NSTAGE;
        if (it >= NSTAGE) mbar_wait(&empty_bar[s], ((it / NSTAGE) - 1) & 1);
        uint8_t* st = smem_raw + (size_t)s * stage_bytes;
        float4* a_raw = reinterpret_cast<float4*>(st);
        float4* a_lo = reinterpret_cast<float4*>(st + a_bytes);
        float4* w_s = reinterpret_cast<float4*>(st + a_bytes * (SPLIT3 ? 2 : 1));
        const float4* wsrc = reinterpret_cast<const float4*>(wt) + (size_t)cg * (25 * 2 * NB);
        for (int e = lt; e < 25 * 2 * NB; e += LOAD_THREADS) cp_async16(&w_s[e], &wsrc[e], true);
        if (F16) {
          // fp32 planes -> registers -> scaled fp16 hi / lo, 8 channels (two float4 chunks) per 16-byte row
          const float sc = __uint_as_float(f16_scale_bits(maxbits[n]));
          uint4* a_hi16 = reinterpret_cast<uint4*>(st);
          uint4* a_lo16 = reinterpret_cast<uint4*>(st + a_bytes);
          // all loads of a pass (both 8-channel k-chunks x TC_LD_U pixels) are in flight before the first conversion: one
          // exposed memory latency per pass instead of one per k-chunk
          const float4* src[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) src[j] = cg * 4 + j < in.total_chunks ? tc_src_chunk_ptr(in, cg * 4 + j, n, D) : nullptr;
          int q = q0 + lt;
          int iy = q / P - 2, ix = q % P - 2;
          for (int h0 = lt; h0 < PH; h0 += TC_LD_U * LOAD_THREADS) {
            float4 v[TC_LD_U][4];
#pragma unroll
            for (int u = 0; u < TC_LD_U; ++u) {
              const bool ok = h0 + u * LOAD_THREADS < PH && iy >= 0 && iy < D && ix >= 0 && ix < D;
              const size_t off = (size_t)iy * D + ix;
#pragma unroll
              for (int j = 0; j < 4; ++j) v[u][j] = (ok && src[j]) ? __ldg(src[j] + off) : make_float4(0.f, 0.f, 0.f, 0.f);
              ix += LOAD_THREADS;
              while (ix >= P - 2) { ix -= P; ++iy; }
            }
#pragma unroll
            for (int u = 0; u < TC_LD_U; ++u) {
              const int h = h0 + u * LOAD_THREADS;
              if (h < PH) {
#pragma unroll
                for (int kc = 0; kc < 2; ++kc) {
                  uint4 hi, lo;
                  f16_split2_scaled(v[u][2 * kc].x, v[u][2 * kc].y, sc, hi.x, lo.x);
                  f16_split2_scaled(v[u][2 * kc].z, v[u][2 * kc].w, sc, hi.y, lo.y);
                  f16_split2_scaled(v[u][2 * kc + 1].x, v[u][2 * kc + 1].y, sc, hi.z, lo.z);
                  f16_split2_scaled(v[u][2 * kc + 1].z, v[u][2 * kc + 1].w, sc, hi.w, lo.w);
                  a_hi16[kc * PH + h] = hi;
                  a_lo16[kc * PH + h] = lo;
                }
              }
            }
          }
          cp_async_commit();
          cp_async_wait<0>();
        } else {
          for (int kc = 0; kc < 2; ++kc) {
            const int c = cg * 2 + kc;
            const float4* src = c < in.total_chunks ? tc_src_chunk_ptr(in, c, n, D) : nullptr;
            int q = q0 + lt;
            int iy = q / P - 2, ix = q % P - 2;
            for (int h = lt; h < PH; h += LOAD_THREADS) {
              bool ok = src != nullptr && iy >= 0 && iy < D && ix >= 0 && ix < D;
              cp_async16(&a_raw[kc * PH + h], ok ? src + (size_t)iy * D + ix : reinterpret_cast<const float4*>(wt), ok);
              ix += LOAD_THREADS;                          // advance LOAD_THREADS pixels in padded-linear space
              while (ix >= P - 2) { ix -= P; ++iy; }
            }
          }
          cp_async_commit();
          cp_async_wait<0>();
          // round the granules this thread loaded itself to tf32 (and split off a_lo in 3xTF32 mode)
          for (int kc = 0; kc < 2; ++kc)
            for (int h = lt; h < PH; h += LOAD_THREADS) {
              float4 v = a_raw[kc * PH + h];
              float4 hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
              a_raw[kc * PH + h] = hi;
              if (SPLIT3)
                a_lo[kc * PH + h] = make_float4(tf32_rn(v.x - hi.x), tf32_rn(v.y - hi.y), tf32_rn(v.z - hi.z), tf32_rn(v.w - hi.w));
            }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer: whole warp in uniform control flow, tcgen05 instructions on the elect.sync leader
    // (descriptor words stay in uniform registers: no ELECT / R2UR waterfall per MMA) =====================
    {
      const bool leader = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc1 = F16 ? make_idesc_f16(128, NB) : make_idesc_tf32(128, NB);
      const uint32_t idesc2 = F16 ? make_idesc_f16(128, NOUT) : make_idesc_tf32(128, NOUT);
      int it = 0, j = 0;
#ifdef OP3_TC_TRACE
      unsigned long long t_full = 0, t_acc = 0, n_mma = 0;
      const unsigned long long t_begin = clock64();
#endif
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
        const int grp = item % groups_img;
        const int ntile = min(G, tiles_img - grp * G);
        const int buf = j & 1;
#ifdef OP3_TC_TRACE
        const unsigned long long ta0 = clock64();
#endif
        if (j >= 2) mbar_wait(&acc_empty[buf], ((j >> 1) - 1) & 1);
#ifdef OP3_TC_TRACE
        t_acc += clock64() - ta0;
#endif
        tc_fence_after();
        const uint32_t dbase = tmem_u + (uint32_t)(buf * 256);
        for (int cg = 0; cg < ngroups; ++cg, ++it) {
          const int s = it % NSTAGE;
#ifdef OP3_TC_TRACE
          const unsigned long long tf0 = clock64();
#endif
          mbar_wait(&full_bar[s], (it / NSTAGE) & 1);
#ifdef OP3_TC_TRACE
          t_full += clock64() - tf0;
          n_mma += (unsigned long long)ntile * 25 * (SPLIT3 ? 2 : 1);
#endif
          tc_fence_after();
          const uint32_t st = smem_u32(smem_raw + (size_t)s * stage_bytes);
          const uint64_t ad = make_smem_desc(st, PH, 8);
          const uint64_t ald = make_smem_desc(st + a_bytes, PH, 8);
          const uint64_t bd = make_smem_desc(st + a_bytes * (SPLIT3 ? 2 : 1), NB, 8);
          const uint32_t a_lo32 = (uint32_t)ad, a_hi32 = (uint32_t)(ad >> 32);
          const uint32_t al_lo32 = (uint32_t)ald, al_hi32 = (uint32_t)(ald >> 32);
          const uint32_t b_lo32 = (uint32_t)bd, b_hi32 = (uint32_t)(bd >> 32);
          for (int t = 0; t < ntile; ++t) {
            const uint32_t dcol = dbase + (uint32_t)(t * NB);
#pragma unroll
            for (int tap = 0; tap < 25; ++tap) {
#ifdef OP3_TC_ALIGNED_PROBE   /* timing experiment only (wrong results): every tap starts on a 128-byte line */
              const uint32_t aoff = (uint32_t)(128 * t + (tap / 5) * P + (tap % 5)) & ~7u;
#else
              const uint32_t aoff = (uint32_t)(128 * t + (tap / 5) * P + (tap % 5));   // 16-byte units
#endif
              const uint32_t boff = (uint32_t)(tap * 2 * NB);
              if (leader) {
                if (F16) {
                  mma_f16(dcol, a_lo32 + aoff, a_hi32, b_lo32 + boff, b_hi32, idesc1, (cg > 0 || tap > 0) ? 1u : 0u);
                  mma_f16(dcol, al_lo32 + aoff, al_hi32, b_lo32 + boff, b_hi32, idesc2, 1u);
                } else {
                  mma_tf32(dcol, a_lo32 + aoff, a_hi32, b_lo32 + boff, b_hi32, idesc1, (cg > 0 || tap > 0) ? 1u : 0u);
                  if (SPLIT3) mma_tf32(dcol, al_lo32 + aoff, al_hi32, b_lo32 + boff, b_hi32, idesc2, 1u);
                }
              }
            }
          }
          if (leader) tc_commit(&empty_bar[s]);          // smem stage free once these MMAs have read it
          __syncwarp();
        }
        if (leader) tc_commit(&acc_full[buf]);           // accumulators of this item complete
      }
#ifdef OP3_TC_TRACE
      if (leader && blockIdx.x < 256) {
        atomicAdd(&g_tc_trace[blockIdx.x][0], t_full); atomicAdd(&g_tc_trace[blockIdx.x][1], t_acc);
        atomicAdd(&g_tc_trace[blockIdx.x][2], clock64() - t_begin); atomicAdd(&g_tc_trace[blockIdx.x][3], n_mma);
      }
#endif
    }
  } else {
    // ===================== epilogue (warp % 4 = TMEM lane quadrant, warp / 4 = tile parity) =====================
    const int quad = warp & 3, half = warp >> 2;
    int j = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
      const int n = item / groups_img, grp = item % groups_img;
      const int ntile = min(G, tiles_img - grp * G);
      const int q0 = grp * G * 128;
      const int buf = j & 1;
      float omax = 0.f;
      float inv_scale = 1.f;
      if (F16) inv_scale = 1.f / (__uint_as_float(f16_scale_bits(maxbits[n])) * __ldg(wscale));   // powers of two: exact
      mbar_wait(&acc_full[buf], (j >> 1) & 1);
      tc_fence_after();
      for (int t = half; t < ntile; t += 2) {
        const int p = q0 + 128 * t + quad * 32 + lane;
        const int y = p / P, x = p % P;
        const bool valid = y < D && x < D;
        const size_t obase = (((size_t)n * out_planes + out_chunk0) * D + (valid ? y : 0)) * D + (valid ? x : 0);
        const size_t plane = (size_t)D * D;
        // prefetch the ELU' reference values of this pixel (independent loads, latency overlapped with the TMEM reads)
        float4 ra[NOUT / 4];
        if (epi == EPI_MUL_ELUGRAD) {
#pragma unroll
          for (int c4 = 0; c4 < NOUT / 4; ++c4)
            ra[c4] = (valid && c4 < out_chunks) ? __ldg(&ref[obase + c4 * plane]) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256 + t * NB);
        float v[NOUT];
#pragma unroll
        for (int c0 = 0; c0 < NOUT; c0 += 16) tmem_ld16(trow + c0, v + c0);
        if (SPLIT3) {
          float u[NOUT];
#pragma unroll
          for (int c0 = 0; c0 < NOUT; c0 += 16) tmem_ld16(trow + NOUT + c0, u + c0);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < NOUT; ++i) v[i] = F16 ? (v[i] + u[i]) * inv_scale : v[i] + u[i];
        } else {
          tmem_ld_wait();
        }
        if (valid) {
#pragma unroll
          for (int c4 = 0; c4 < NOUT / 4; ++c4) {
            if (c4 >= out_chunks) continue;
            float4 r = make_float4(v[c4 * 4] + s_bias[c4 * 4], v[c4 * 4 + 1] + s_bias[c4 * 4 + 1],
                                   v[c4 * 4 + 2] + s_bias[c4 * 4 + 2], v[c4 * 4 + 3] + s_bias[c4 * 4 + 3]);
            if (epi == EPI_ELU) {
              r.x = elu_f(r.x); r.y = elu_f(r.y); r.z = elu_f(r.z); r.w = elu_f(r.w);
            } else if (epi == EPI_MUL_ELUGRAD) {
              r.x *= elu_grad_from_out(ra[c4].x); r.y *= elu_grad_from_out(ra[c4].y);
              r.z *= elu_grad_from_out(ra[c4].z); r.w *= elu_grad_from_out(ra[c4].w);
            }
            out[obase + c4 * plane] = r;
            if (F16) omax = fmaxf(omax, fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w))));
          }
        }
      }
      if (F16 && in.out_maxbits != nullptr) {             // magnitudes of the written tensor, for its consumer's operand scale
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, o));
        if (lane == 0 && __float_as_uint(omax) > __ldcg(&in.out_maxbits[n])) atomicMax(&in.out_maxbits[n], __float_as_uint(omax));
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);                      // this thread is done reading the buffer
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<512>(tmem_base);
}

// =====================================================================================================================
// "Quint-tap" fp16 variant (cin <= 32, 32 output channels per launch): the tensor core accepts one M=128 MMA per ~50 cycles
// for ANY N <= 64 (measured in situ: 107 cycles for the N=64 + N=32 pair of one tap), so the kernel above pays per tap, not
// per byte.  Here the five kx taps of a kernel row share ONE activation fetch: B = [w(kx=0); ...; w(kx=4)] is stacked along
// N (N = 160) and the tap shift moves to the epilogue,
//     D_kx[r] = sum_{ky,ci} w(ky,kx)[co][ci] * x[ci][r + ky*P]   (A operand: the tile shifted by ky*P only)
//     out[p]  = sum_kx D_kx[p + kx]                               (lane shift inside the epilogue warp)
// so a kernel row costs 3 MMAs (x_hi*w_hi, x_hi*w_lo, x_lo*w_hi, all into the same 160 TMEM columns) instead of 10.
// Consequences: * a tile is 128 accumulator rows but only 124 outputs (rows 124..127 exist to be shifted into 120..123),
//   tiles advance by 124 padded-linear pixels; * 160 columns per tile: a 3-slot accumulator ring filled TILE BY TILE, so
//   both 16-channel K groups of a tile must be resident: the activations live in a RING in padded-linear pixel space
//   ([hi/lo][4 k-chunks of 8 channels][6 chunk slots x 124 rows + 128 guard rows] x 16 B; every pixel is loaded, scaled and
//   split once per image segment instead of 1.5x per halo group) and the weights are RESIDENT (100 KB, re-laid out from the
//   ordinary fp16 slab by the prologue: [K group][ky][hi/lo][k-chunk][kx*32 + co]); * each CTA owns a contiguous range of
//   the global tile sequence (balanced to one tile) and restarts the ring at image boundaries (3 extra chunks).
// =====================================================================================================================
// ELU of the quint-tap epilogue: ex2.approx(x * log2 e) - 1 (absolute error <= ~2e-7 for x <= 0, below the 2e-6 floor of the
// fp16-split products feeding it) instead of expm1f, whose ~20 instructions per channel made the epilogue -- 32 channels x
// 124 pixels per tile -- the kernel's bottleneck once the tensor core needed only 2400 cycles per tile.
__device__ __forceinline__ float elu_tc(float x) { return x > 0.f ? x : __expf(x) - 1.f; }
#ifdef OP3_TC_TRACE
__device__ int g_tcq_dbg = 0;      // timing-only experiments (results are WRONG when set): 16 = loaders skip loads, conversion and stores
#endif

#ifndef OP3_TCQ_SMEM_A
#define OP3_TCQ_TMEM_A 1      /* x_hi staged in TMEM once per kernel row (define OP3_TCQ_SMEM_A for the all-smem form) */
#endif
constexpr int TQ_PAD = 2;                           // pad columns per row of the padded-linear pixel space
constexpr int TQ_TS = 124;                          // outputs per tile
constexpr int TQ_NSLOT = 6;                         // chunk slots in the ring: 4 in use by the current tile + 2 ahead
constexpr int TQ_RING = TQ_NSLOT * TQ_TS;           // 744 rows
constexpr int TQ_ROWS = TQ_RING + 128;              // + guard copy of rows [0, 128) for operands that wrap
constexpr int TQ_PLANE = TQ_ROWS * 16;              // bytes
constexpr int TQ_A_BYTES = 8 * TQ_PLANE;            // planes [hi/lo][k-chunk 0..3]
constexpr int TQ_NQ = 160;                          // 5 taps x 32 output channels
constexpr int TQ_W_SLAB = 2 * TQ_NQ * 16;           // one (K group, ky, hi/lo) slab: [k-chunk 2][160 rows][16 B]
constexpr int TQ_W_BYTES = 2 * 5 * 2 * TQ_W_SLAB;
constexpr int TQ_XCH = 4 * 2 * 10 * 16;             // floats per exchange buffer: [quad][half][(kx, lane < kx)][16]
constexpr int TQ_SMEM = TQ_A_BYTES + TQ_W_BYTES + 2 * TQ_XCH * 4 + 8 * 128 * 4;          // single-CTA form, two epilogue groups
#ifndef OP3_TQ_LEAN
#define OP3_TQ_LEAN 0
#endif
constexpr bool TQ_LEAN = OP3_TQ_LEAN != 0;          // half-wise epilogue (see the kernel)
#ifndef OP3_TQ_PAIR_NEG
#define OP3_TQ_PAIR_NEG 2
#endif
constexpr int TQ_PAIR_NEG = OP3_TQ_PAIR_NEG;                      // epilogue groups of the pair form (3 fits its shared memory but not the issue slots, see below)
constexpr int TQ_SMEM_PAIR = TQ_A_BYTES + TQ_W_BYTES / 2 + TQ_PAIR_NEG * TQ_XCH * 4 + 4 * TQ_PAIR_NEG * 128 * 4;

// PAIR: the kernel runs as clusters of two CTAs driving ONE tcgen05.mma.cta_group::2 stream (M = 256): the two CTAs work on
// two different items (images / strips) at the same tile position, so their operand rows sit at the same ring offsets; each
// holds HALF of every weight slab (80 of the 160 stacked rows), which halves the weight bytes the tensor core fetches from each
// CTA's shared memory (15.5 instead of 23 KB per kernel row and K group) -- the crossbar the loaders' stores and the
// epilogue's shuffles compete for.  The leader (cluster rank 0) issues; loaders and epilogue warps of both CTAs report to the
// leader's barriers (mapa + arrive.release.cluster), tcgen05.commit multicasts the completions back to both.
// NEG epilogue groups of four warps (one per TMEM lane quadrant) take the tiles round-robin.  The single-CTA form has room for
// two exchange buffer sets in its 227 KB.  Measured for the pair form (tools/tc_trace.py, decoder forward at N = 256): with two
// groups the leader issues at 62 cycles per MMA (85 single-CTA) but waits 31 % of its loop for accumulators and 19 % for
// operands -- 0.647 vs 0.615 ms; with three groups (640 threads, 96 registers) the accumulator wait disappears and the six
// loader warps starve (64 % operand wait, 0.93 ms): loaders + epilogue are bound by instruction issue, not by the tensor pipe.
#ifndef OP3_TQ_NLOAD
#define OP3_TQ_NLOAD 6
#endif
#ifndef OP3_TQ_IDLE
#define OP3_TQ_IDLE 1
#endif
// out += v as one 16-byte reduction (REDG.E.ADD.F32x4): no return value, no register or latency cost for the issuing warp
__device__ __forceinline__ void red_add_f4(float4* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int NEG> struct TqRoles {
  static constexpr bool HAS_IDLE = OP3_TQ_IDLE != 0;
  static constexpr int EPI = 4 * NEG, MMA = EPI, IDLE = HAS_IDLE ? MMA + 4 : 1000, LOAD0 = MMA + 1, NLOAD = OP3_TQ_NLOAD;   // IDLE: the loader-range warp on the issuer's scheduler
  static constexpr int NROUND = (TQ_TS * 4 + NLOAD * 32 - 1) / (NLOAD * 32);                  // (pixel, k-chunk) tasks per loader thread
  static constexpr int THREADS = (LOAD0 + NLOAD + (HAS_IDLE ? 1 : 0)) * 32;
  static constexpr int SMEM_X = NEG * TQ_XCH * 4 + EPI * 128 * 4;                              // exchange buffers + per-warp fix-up rows
};
template <bool PAIR, int NEG, bool LEAN, bool STRIPS, bool ACCUM>
__global__ void __launch_bounds__(TqRoles<NEG>::THREADS, 1)
conv5x5_tcq_kernel(ConvInput in, int N, int D, const float* __restrict__ wt, const float* __restrict__ bias, int cout_real,
                   float4* __restrict__ out, int out_planes, int out_chunk0, int out_chunks, int epi,
                   const float4* __restrict__ ref, const uint32_t* __restrict__ maxbits, const float* __restrict__ wscale,
                   int W_arg, int nstrips_arg, const __grid_constant__ CUtensorMap wmap, int use_tma) {
  // STRIPS = false is the whole-image instance (every 64 x 64 configuration): W = D, one item per image and x0 = 0 are
  // compile-time facts there -- as runtime values they cost the loaders and the epilogue ~10 % (registers, index arithmetic)
  const int W = STRIPS ? W_arg : D, nstrips = STRIPS ? nstrips_arg : 1;
  // An "item" is one image (nstrips == 1, W == D) or one vertical strip of W output columns of an image (D = 128: two
  // strips of 64, so that 128 + 4 P rows still fit four chunks of the ring).  Whole images use pitch D + 2: the two
  // right-hand pad columns of a row double as the left-hand pad of the next (3 % padding rows instead of 6 %).  Strips carry
  // REAL neighbour pixels in their halo columns, so both halos are stored: pitch W + 4.
  const int P = nstrips == 1 ? D + TQ_PAD : W + 4;
  const int tiles_img = (D * P + TQ_TS - 1) / TQ_TS;       // tiles per item
  // PAIR: the tile sequence counts item PAIRS; CTA `rank` of pair blockIdx.x / 2 takes item 2 * (g / tiles_img) + rank
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int n_workers = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x, worker = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const long long total = (long long)(PAIR ? N * nstrips / 2 : N * nstrips) * tiles_img;
  const int g0 = (int)(total * worker / n_workers), g1 = (int)(total * (worker + 1) / n_workers);
  constexpr int NQL = PAIR ? TQ_NQ / 2 : TQ_NQ;            // weight rows per slab held by this CTA
  const int ngroups = (in.total_chunks + 3) / 4;           // K groups of 16 channels (1 or 2)
  // the warp index is broadcast with a shuffle so that the compiler KNOWS the role branches are warp-uniform (it then keeps
  // the issuer's descriptor arithmetic in uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;
  using Roles = TqRoles<NEG>;
  constexpr int EPI_WARPS = Roles::EPI, MMA_WARP = Roles::MMA, TQ_IDLE_WARP = Roles::IDLE, LOAD_WARP0 = Roles::LOAD0,
                TQ_LOAD_THREADS = Roles::NLOAD * 32, TC_THREADS = Roles::THREADS;
  constexpr int W_BYTES = PAIR ? TQ_W_BYTES / 2 : TQ_W_BYTES;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* a_s = smem_raw;
  uint4* w_s = reinterpret_cast<uint4*>(smem_raw + TQ_A_BYTES);
  float* xch = reinterpret_cast<float*>(smem_raw + TQ_A_BYTES + W_BYTES);
  float* fixb = xch + NEG * TQ_XCH;                            // per epilogue warp: [4 target lanes][32] boundary sums
  __shared__ uint64_t full_bar[TQ_NSLOT], empty_bar[TQ_NSLOT], acc_full[3], acc_empty[3], w_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[32];

  if (tid < 32) s_bias[tid] = (bias != nullptr && tid < cout_real) ? bias[tid] : 0.f;
  if (tid == 0) {
    // PAIR: the leader's full / acc_empty barriers collect the arrivals of BOTH CTAs (the peer's own copies stay unused)
    for (int s = 0; s < TQ_NSLOT; ++s) { mbar_init(&full_bar[s], (PAIR ? 2 : 1) * TQ_LOAD_THREADS); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 3; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], (PAIR ? 2 : 1) * 128); }
    mbar_init(&w_bar, 1);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    if (PAIR) tmem_alloc2<512>(&tmem_base_s);
    else tmem_alloc<512>(&tmem_base_s);
  }
  if (PAIR || !use_tma) {
    // resident weights: ordinary slab [K group][tap][k-chunk][64 rows: hi 0..31, lo 32..63] -> [K group][ky][hi/lo][k-chunk][kx*32 + co]
    // (PAIR: only this CTA's half of the 160 stacked rows, [rank * 80, rank * 80 + 80)).  The single-CTA form fetches them with
    // TMA tensor copies instead (issued by the MMA warp below).
    const uint4* wsrc = reinterpret_cast<const uint4*>(wt);
    for (int e = tid; e < ngroups * 25 * 2 * 64; e += TC_THREADS) {
      const int row64 = e & 63, kc = (e >> 6) & 1, tap = (e >> 7) % 25, cg = e / (128 * 25);
      const int hl = row64 >> 5, co = row64 & 31, ky = tap / 5, kx = tap - 5 * ky;
      const int r160 = kx * 32 + co, rl = r160 - (int)rank * NQL;
      if (rl >= 0 && rl < NQL) w_s[((((cg * 5 + ky) * 2 + hl) * 2 + kc) * NQL) + rl] = __ldg(wsrc + e);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  if (PAIR) cluster_sync_all();                            // both CTAs' barriers and TMEM exist before any remote arrive / MMA
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == TQ_IDLE_WARP) {
    // shares the issuer's scheduler (warp % 4): left idle so that the issuing thread is not starved of issue slots
  } else if (warp >= LOAD_WARP0) {
    // ===================== loaders: one chunk of 124 pixels x (2 * ngroups) k-chunks per step =====================
    // A thread owns up to three fixed (pixel-in-chunk, k-chunk) tasks; the loads of chunk c+1 are in flight while chunk c
    // is converted and stored (one exposed global latency per item, not per chunk); pixel coordinates advance incrementally.
    const int lt = tid - LOAD_WARP0 * 32 - (warp > TQ_IDLE_WARP ? 32 : 0);
    const int ntask = TQ_TS * 2 * ngroups;                 // task = (pixel, k-chunk of 8 channels)
    constexpr int NR = Roles::NROUND;
    int ti[NR], tk[NR];
    bool tact[NR];
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      const int e = lt + u * TQ_LOAD_THREADS;
      tact[u] = e < ntask;
      tk[u] = tact[u] ? e / TQ_TS : 0;
      ti[u] = e - (e / TQ_TS) * TQ_TS;
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = 0;                                             // running chunk counter of this CTA
    for (int g = g0; g < g1;) {
      const int itw = g / tiles_img, ta = g - itw * tiles_img, it = PAIR ? 2 * itw + (int)rank : itw;
      const int n = it / nstrips, x0 = (it - n * nstrips) * W;
      const int T = min(tiles_img - ta, g1 - g);
      const float sc = __uint_as_float(f16_scale_bits(maxbits[n]));
      const float4* s0[NR];
      const float4* s1[NR];
      int iy[NR], ix[NR];
#pragma unroll
      for (int u = 0; u < NR; ++u) {
        s0[u] = (tact[u] && 2 * tk[u] < in.total_chunks) ? tc_src_chunk_ptr(in, 2 * tk[u], n, D) : nullptr;
        s1[u] = (tact[u] && 2 * tk[u] + 1 < in.total_chunks) ? tc_src_chunk_ptr(in, 2 * tk[u] + 1, n, D) : nullptr;
        const int q = TQ_TS * ta + ti[u];
        const int r = q / P;
        iy[u] = r - 2; ix[u] = q - r * P - 2;
      }
      float4 v[NR][2], vn[NR][2];
      auto load_chunk = [&](float4 (*dst)[2]) {            // loads the chunk at the current coordinates, then advances them
#pragma unroll
        for (int u = 0; u < NR; ++u) {
          const bool ok = (unsigned)iy[u] < (unsigned)D && (unsigned)(ix[u] + x0) < (unsigned)D;
          const size_t off = (size_t)iy[u] * D + ix[u] + x0;
          dst[u][0] = (ok && s0[u]) ? __ldg(s0[u] + off) : zero4;
          dst[u][1] = (ok && s1[u]) ? __ldg(s1[u] + off) : zero4;
          ix[u] += TQ_TS;
          while (ix[u] >= P - 2) { ix[u] -= P; ++iy[u]; }
        }
      };
#ifdef OP3_TC_TRACE
      const bool skip = (g_tcq_dbg & 16) != 0;
#else
      constexpr bool skip = false;
#endif
      if (!skip) load_chunk(v);
      for (int jc = 0; jc < T + 3; ++jc, ++c) {
        const int slot = c % TQ_NSLOT;
        if (jc + 1 < T + 3 && !skip) load_chunk(vn);
        if (c >= TQ_NSLOT) mbar_wait_warp(&empty_bar[slot], ((c / TQ_NSLOT) - 1) & 1);
#pragma unroll
        for (int u = 0; u < NR; ++u)
          if (tact[u] && !skip) {
            uint4 hi, lo;
            f16_split2_scaled(v[u][0].x, v[u][0].y, sc, hi.x, lo.x);
            f16_split2_scaled(v[u][0].z, v[u][0].w, sc, hi.y, lo.y);
            f16_split2_scaled(v[u][1].x, v[u][1].y, sc, hi.z, lo.z);
            f16_split2_scaled(v[u][1].z, v[u][1].w, sc, hi.w, lo.w);
            const int row = slot * TQ_TS + ti[u];
            uint4* ph = reinterpret_cast<uint4*>(a_s + (size_t)tk[u] * TQ_PLANE);
            uint4* pl = reinterpret_cast<uint4*>(a_s + (size_t)(4 + tk[u]) * TQ_PLANE);
            ph[row] = hi; pl[row] = lo;
            if (row < 128) { ph[row + TQ_RING] = hi; pl[row + TQ_RING] = lo; }
          }
        fence_proxy_async();
        if (PAIR) mbar_arrive_cluster(&full_bar[slot], 0);
        else mbar_arrive(&full_bar[slot]);
#pragma unroll
        for (int u = 0; u < NR; ++u) { v[u][0] = vn[u][0]; v[u][1] = vn[u][1]; }
      }
      g += T;
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    // The WHOLE warp runs this loop in uniform control flow and only the tcgen05 instructions are predicated on one lane:
    // descriptor words computed by all lanes from loop counters and kernel parameters stay in UNIFORM registers.  Issued
    // from inside an `if (lane == 0)` region instead, every MMA is wrapped in an ELECT / R2UR.BROADCAST waterfall of ~12
    // instructions (~25 cycles per MMA: the whole gap between the isolated and the in-situ MMA rate).
    if (!PAIR || rank == 0) {
      const bool leader = elect_one();
      if (!PAIR && use_tma) {
        // The slab is the contiguous array [K group][ky][kx][j = k-chunk*2 + hi/lo][32 co x 8 ci halves]: one box
        // {256 halves, 1, 5 kx, 1, 1} per (K group, ky, j) is exactly the 160-row operand slab the MMAs take
        // (tests/probes/probe_tma_weights.cu).  20 x ngroups cp.async.bulk.tensor copies by one thread replace ~1500
        // instructions per thread of loads, index arithmetic and stores; loaders and epilogue warps never wait for them.
        if (leader)
          asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n"
                       :: "r"(smem_u32(&w_bar)), "r"((uint32_t)ngroups * 20u * (uint32_t)(TQ_NQ * 16)) : "memory");
        for (int e = lane; e < ngroups * 20; e += 32) {        // one copy per lane and round: a single thread needs ~150 cycles per issue
          const int cg = e / 20, ky = (e - cg * 20) >> 2, jj = e & 3;
          const uint4* dst = w_s + ((((cg * 5 + ky) * 2 + (jj & 1)) * 2 + (jj >> 1)) * TQ_NQ);
          asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                       :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&wmap)), "r"(0), "r"(jj), "r"(0), "r"(ky), "r"(cg),
                          "r"(smem_u32(&w_bar)) : "memory");
        }
        mbar_wait(&w_bar, 0);
      }
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, TQ_NQ);
      const uint64_t ad = make_smem_desc(smem_u32(a_s), TQ_ROWS, 8);
      const uint64_t bd = make_smem_desc(smem_u32(w_s), NQL, 8);
      const uint32_t a_lo32 = (uint32_t)ad, a_hi32 = (uint32_t)(ad >> 32);
      const uint32_t b_lo32 = (uint32_t)bd, b_hi32 = (uint32_t)(bd >> 32);
      int c_run = 0, j = 0;
#ifdef OP3_TC_TRACE
      unsigned long long t_full = 0, t_acc = 0, n_mma = 0;
      const unsigned long long t_begin = clock64();
#endif
      for (int g = g0; g < g1;) {
        const int ta = g % tiles_img;
        const int T = min(tiles_img - ta, g1 - g);
        const int c0 = c_run;
        int ready = c0 - 1;                                // last chunk known to be in shared memory
        for (int tau = 0; tau < T; ++tau, ++j) {
          const int slot = j % 3;
#ifdef OP3_TC_TRACE
          unsigned long long tw = clock64();
#endif
          if (j >= 3) mbar_wait(&acc_empty[slot], ((j / 3) - 1) & 1);
#ifdef OP3_TC_TRACE
          t_acc += clock64() - tw; tw = clock64();
#endif
          while (ready < c0 + tau + 3) {
            ++ready;
            mbar_wait(&full_bar[ready % TQ_NSLOT], (ready / TQ_NSLOT) & 1);
          }
#ifdef OP3_TC_TRACE
          t_full += clock64() - tw; n_mma += 15 * ngroups;
#endif
          tc_fence_after();
          const uint32_t dcol = tmem_u + (uint32_t)(slot * TQ_NQ);
          const int rowbase = ((c0 + tau) % TQ_NSLOT) * TQ_TS;
          uint32_t arow[5];                                // operand start rows of the five kernel rows (ring wrap)
#pragma unroll
          for (int ky = 0; ky < 5; ++ky) {
            int row = rowbase + ky * P;
            if (row >= TQ_RING) row -= TQ_RING;
            arow[ky] = a_lo32 + (uint32_t)row;
          }
#pragma unroll
          for (int cg = 0; cg < 2; ++cg) {
            if (cg < ngroups) {
#pragma unroll
              for (int ky = 0; ky < 5; ++ky) {
                const uint32_t a_hi_op = arow[ky] + (uint32_t)((2 * cg) * TQ_ROWS);             // 16-byte units
                const uint32_t a_lo_op = arow[ky] + (uint32_t)((4 + 2 * cg) * TQ_ROWS);
                const uint32_t b_hi_op = b_lo32 + (uint32_t)(((cg * 5 + ky) * 2 + 0) * 2 * NQL);
                const uint32_t b_lo_op = b_lo32 + (uint32_t)(((cg * 5 + ky) * 2 + 1) * 2 * NQL);
                if (leader) {
                  if (PAIR) {
                    tc2_cp_a128(tmem_u + 480u, a_hi_op, a_hi32);
                    if (cg == 0 && ky == 0) mma2_f16_ts(dcol, tmem_u + 480u, b_hi_op, b_hi32, idesc, 0u);
                    else mma2_f16_ts_acc(dcol, tmem_u + 480u, b_hi_op, b_hi32, idesc);
                    mma2_f16_ts_acc(dcol, tmem_u + 480u, b_lo_op, b_hi32, idesc);
                    mma2_f16_acc(dcol, a_lo_op, a_hi32, b_hi_op, b_hi32, idesc);
                  } else {
#ifdef OP3_TCQ_TMEM_A
                    // x_hi feeds two MMAs: staged once in the 8 TMEM columns behind the accumulators (tcgen05.cp, in issue order
                    // with the MMAs), so they fetch only their 5 KB of weights from shared memory
                    tc_cp_a128(tmem_u + 480u, a_hi_op, a_hi32);
                    if (cg == 0 && ky == 0) mma_f16_ts(dcol, tmem_u + 480u, b_hi_op, b_hi32, idesc, 0u);
                    else mma_f16_ts_acc(dcol, tmem_u + 480u, b_hi_op, b_hi32, idesc);
                    mma_f16_ts_acc(dcol, tmem_u + 480u, b_lo_op, b_hi32, idesc);
#else
                    if (cg == 0 && ky == 0) mma_f16(dcol, a_hi_op, a_hi32, b_hi_op, b_hi32, idesc, 0u);
                    else mma_f16_acc(dcol, a_hi_op, a_hi32, b_hi_op, b_hi32, idesc);
                    mma_f16_acc(dcol, a_hi_op, a_hi32, b_lo_op, b_hi32, idesc);
#endif
                    mma_f16_acc(dcol, a_lo_op, a_hi32, b_hi_op, b_hi32, idesc);
                  }
                }
              }
            }
          }
          if (leader) {
            if (PAIR) {                                    // the same barrier offsets in both CTAs
              tc2_commit(&empty_bar[(c0 + tau) % TQ_NSLOT]);
              if (tau == T - 1)
                for (int x = 0; x < 3; ++x) tc2_commit(&empty_bar[(c0 + T + x) % TQ_NSLOT]);
              tc2_commit(&acc_full[slot]);
            } else {
              tc_commit(&empty_bar[(c0 + tau) % TQ_NSLOT]);  // chunk c0+tau is dead once these MMAs have read it
              if (tau == T - 1)
                for (int x = 0; x < 3; ++x) tc_commit(&empty_bar[(c0 + T + x) % TQ_NSLOT]);   // the tail halo chunks of the item
              tc_commit(&acc_full[slot]);
            }
          }
          __syncwarp();
        }
        c_run = c0 + T + 3;
        g += T;
      }
#ifdef OP3_TC_TRACE
      if (leader && blockIdx.x < 256) {
        atomicAdd(&g_tc_trace[blockIdx.x][0], t_full); atomicAdd(&g_tc_trace[blockIdx.x][1], t_acc);
        atomicAdd(&g_tc_trace[blockIdx.x][2], clock64() - t_begin); atomicAdd(&g_tc_trace[blockIdx.x][3], n_mma);
      }
#endif
    }
  } else {
    // ===================== epilogue: two groups of four warps (one per TMEM lane quadrant) take alternate tiles, so a
    // tile's epilogue may last two tile times; each warp handles the 32 output channels in two passes of 16 =====================
    const int quad = warp & 3, grp = warp >> 2;
    const size_t plane = (size_t)D * D;
    float* xb = xch + grp * TQ_XCH;
    const bool want_max = in.out_maxbits != nullptr;
    // EPI_ACCUM: this launch is the second K pass of a contraction over more than 32 channels -- its results are ADDED to what
    // the first pass stored (one fire-and-forget 16-byte reduction per float4: each element has exactly one writer, so the sum
    // is deterministic).  EPI_MAX2: record twice the magnitudes, so that the two passes' maxima bound the magnitude of the sum.
    const int epi_b = epi & 3;
    constexpr bool accum = ACCUM;                          // (a separate instance: as a runtime flag it cost the main path ~5 %)
    const float max_mul = (epi & EPI_MAX2) ? 2.f : 1.f;
    const float m1 = lane < 31 ? 1.f : 0.f, m2 = lane < 30 ? 1.f : 0.f, m3 = lane < 29 ? 1.f : 0.f, m4 = lane < 28 ? 1.f : 0.f;
    int j = 0;
    for (int g = g0; g < g1;) {
      const int itw = g / tiles_img, ta = g - itw * tiles_img, it = PAIR ? 2 * itw + (int)rank : itw;
      const int n = it / nstrips, x0 = (it - n * nstrips) * W;
      const int T = min(tiles_img - ta, g1 - g);
      const float inv_scale = 1.f / (__uint_as_float(f16_scale_bits(maxbits[n])) * __ldg(wscale));   // powers of two: exact
      float omax = 0.f;
      for (int tau = 0; tau < T; ++tau, ++j) {
        if ((j % NEG) != grp) continue;
        const int slot = j % 3;
        const int r = quad * 32 + lane;
        const int p = TQ_TS * (ta + tau) + r;
        const int y = p / P, x = p - y * P;
        const bool valid = r < TQ_TS && y < D && x < W;
        const size_t obase = (((size_t)n * out_planes + out_chunk0) * D + (valid ? y : 0)) * D + (valid ? x + x0 : 0);
#ifdef OP3_TC_TRACE
        const unsigned long long te0 = clock64();
#endif
        mbar_wait_warp(&acc_full[slot], (j / 3) & 1);
#ifdef OP3_TC_TRACE
        const unsigned long long te1 = clock64();
#endif
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * TQ_NQ);
        // Every accumulator value is read from TMEM once, two 16-column groups per wait (TMEM load latency is long while
        // the tensor core is busy; one wait per load made the epilogue the bottleneck).  out[p] = sum_kx D_kx[p + kx]: the
        // shift is a warp shuffle; the rows that cross a quadrant boundary (lane l < kx of D_kx, read by lane 32 - kx + l of
        // the previous quadrant) travel through shared memory and are added after the group barrier.
        if constexpr (LEAN) {
          // Half-wise form: 16 output channels at a time run through the WHOLE pipeline (TMEM -> tap shift -> boundary exchange
          // -> bias / ELU -> stores) with the 16-column units double-buffered against the shuffles: 48 live accumulator
          // registers instead of 96, which is what lets three epilogue groups share the register file with eight loader warps.
          // Same arithmetic in the same order as the form below (bit-identical results).  The two halves use the two halves of
          // the exchange buffer, so one group barrier per half is enough: a region is rewritten only after every warp of the
          // group has passed the other half's barrier, i.e. after its last read.
          float* fx = fixb + warp * 128;                     // [4 target lanes][16 columns]
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float acc[16], va[16], vb[16];
            float4* dst4 = reinterpret_cast<float4*>(xb + ((quad * 2 + half) * 10) * 16);
            tmem_ld16(trow + 16 * half, acc);
            tmem_ld16(trow + 32 + 16 * half, va);
            tmem_ld_wait();
            tmem_ld16(trow + 64 + 16 * half, vb);
            if (lane < 1) {
#pragma unroll
              for (int i = 0; i < 4; ++i) dst4[(0 + lane) * 4 + i] = make_float4(va[4 * i], va[4 * i + 1], va[4 * i + 2], va[4 * i + 3]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(__shfl_down_sync(0xffffffffu, va[i], 1), m1, acc[i]);
            tmem_ld_wait();
            tmem_ld16(trow + 96 + 16 * half, va);
            if (lane < 2) {
#pragma unroll
              for (int i = 0; i < 4; ++i) dst4[(1 + lane) * 4 + i] = make_float4(vb[4 * i], vb[4 * i + 1], vb[4 * i + 2], vb[4 * i + 3]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(__shfl_down_sync(0xffffffffu, vb[i], 2), m2, acc[i]);
            tmem_ld_wait();
            tmem_ld16(trow + 128 + 16 * half, vb);
            if (lane < 3) {
#pragma unroll
              for (int i = 0; i < 4; ++i) dst4[(3 + lane) * 4 + i] = make_float4(va[4 * i], va[4 * i + 1], va[4 * i + 2], va[4 * i + 3]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(__shfl_down_sync(0xffffffffu, va[i], 3), m3, acc[i]);
            tmem_ld_wait();
            if (lane < 4) {
#pragma unroll
              for (int i = 0; i < 4; ++i) dst4[(6 + lane) * 4 + i] = make_float4(vb[4 * i], vb[4 * i + 1], vb[4 * i + 2], vb[4 * i + 3]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(__shfl_down_sync(0xffffffffu, vb[i], 4), m4, acc[i]);
            if (half == 1) {
              tc_fence_before();
              if (PAIR) mbar_arrive_cluster(&acc_empty[slot], 0);    // every TMEM read of this tile is done
              else mbar_arrive(&acc_empty[slot]);
            }
            float4 rh[4];                                    // dgrad: activation outputs whose ELU' multiplies this half
            if (epi_b == EPI_MUL_ELUGRAD) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                rh[i] = (valid && 4 * half + i < out_chunks) ? __ldg(&ref[obase + (4 * half + i) * plane]) : make_float4(1.f, 1.f, 1.f, 1.f);
            }
            asm volatile("bar.sync %0, 128;" :: "r"(1 + grp) : "memory");      // the four warps of this group
            if (quad < 3) {
              // rows whose shifted sources live in the next quadrant (lanes 28..31): lane = (target pair, column) sums the kx
              // terms two targets are missing, the targets then pick up their 16 sums with four LDS.128
              const float* src = xb + (((quad + 1) * 2 + half) * 10) * 16 + (lane & 15);
#pragma unroll
              for (int t2 = 0; t2 < 2; ++t2) {
                const int tt = 2 * (lane >> 4) + t2;         // target lane 28 + tt misses kx = 4 - tt .. 4 (source lane tt + kx - 4)
                float sl = 0.f;
#pragma unroll
                for (int kx = 1; kx <= 4; ++kx)
                  if (kx >= 4 - tt) sl += src[((kx * (kx - 1)) / 2 + (tt + kx - 4)) * 16];
                fx[tt * 16 + (lane & 15)] = sl;
              }
              __syncwarp();
              if (lane >= 28) {
                const float4* f4 = reinterpret_cast<const float4*>(fx + (lane - 28) * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 t = f4[i];
                  acc[4 * i] += t.x; acc[4 * i + 1] += t.y; acc[4 * i + 2] += t.z; acc[4 * i + 3] += t.w;
                }
              }
              __syncwarp();
            }
            if (valid) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int c4 = 4 * half + i;
                if (c4 >= out_chunks) continue;
                const float4 bb = *reinterpret_cast<const float4*>(&s_bias[16 * half + 4 * i]);
                float4 o = make_float4(fmaf(acc[4 * i], inv_scale, bb.x), fmaf(acc[4 * i + 1], inv_scale, bb.y),
                                       fmaf(acc[4 * i + 2], inv_scale, bb.z), fmaf(acc[4 * i + 3], inv_scale, bb.w));
                if (epi_b == EPI_ELU) {
                  o.x = elu_tc(o.x); o.y = elu_tc(o.y); o.z = elu_tc(o.z); o.w = elu_tc(o.w);
                } else if (epi_b == EPI_MUL_ELUGRAD) {
                  o.x *= elu_grad_from_out(rh[i].x); o.y *= elu_grad_from_out(rh[i].y);
                  o.z *= elu_grad_from_out(rh[i].z); o.w *= elu_grad_from_out(rh[i].w);
                }
                if (accum) red_add_f4(&out[obase + c4 * plane], o); else out[obase + c4 * plane] = o;
                if (want_max) omax = fmaxf(omax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
              }
            }
          }
        } else {
        float acc[2][16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v1[16], v2[16];
          tmem_ld16(trow + 16 * half, acc[half]);
          tmem_ld16(trow + 32 + 16 * half, v1);
          tmem_ld16(trow + 64 + 16 * half, v2);
          tmem_ld_wait();
          float4* dst4 = reinterpret_cast<float4*>(xb + ((quad * 2 + half) * 10) * 16);
          if (lane < 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dst4[(0 + lane) * 4 + i] = make_float4(v1[4 * i], v1[4 * i + 1], v1[4 * i + 2], v1[4 * i + 3]);
          }
          if (lane < 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dst4[(1 + lane) * 4 + i] = make_float4(v2[4 * i], v2[4 * i + 1], v2[4 * i + 2], v2[4 * i + 3]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {         // out-of-range shuffles return the caller's own (finite) value: masked by 0
            acc[half][i] = fmaf(__shfl_down_sync(0xffffffffu, v1[i], 1), m1, acc[half][i]);
            acc[half][i] = fmaf(__shfl_down_sync(0xffffffffu, v2[i], 2), m2, acc[half][i]);
          }
          tmem_ld16(trow + 96 + 16 * half, v1);
          tmem_ld16(trow + 128 + 16 * half, v2);
          tmem_ld_wait();
          if (lane < 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dst4[(3 + lane) * 4 + i] = make_float4(v1[4 * i], v1[4 * i + 1], v1[4 * i + 2], v1[4 * i + 3]);
          }
          if (lane < 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dst4[(6 + lane) * 4 + i] = make_float4(v2[4 * i], v2[4 * i + 1], v2[4 * i + 2], v2[4 * i + 3]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            acc[half][i] = fmaf(__shfl_down_sync(0xffffffffu, v1[i], 3), m3, acc[half][i]);
            acc[half][i] = fmaf(__shfl_down_sync(0xffffffffu, v2[i], 4), m4, acc[half][i]);
          }
        }
        tc_fence_before();
        if (PAIR) mbar_arrive_cluster(&acc_empty[slot], 0);  // every TMEM read of this tile is done (the leader collects both CTAs)
        else mbar_arrive(&acc_empty[slot]);
        // dgrad: the activation outputs whose ELU' multiplies the result -- requested only now (the tap-shift temporaries are
        // dead), their latency hides behind the group barrier and the boundary fix-up
        float4 ra[8];
        if (epi_b == EPI_MUL_ELUGRAD) {
#pragma unroll
          for (int i = 0; i < 8; ++i) ra[i] = (valid && i < out_chunks) ? __ldg(&ref[obase + i * plane]) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
#ifdef OP3_TC_TRACE
        const unsigned long long te2 = clock64();
#endif
        asm volatile("bar.sync %0, 128;" :: "r"(1 + grp) : "memory");      // the four warps of this group
#ifdef OP3_TC_TRACE
        const unsigned long long te3 = clock64();
#endif
        if (quad < 3) {
          // Rows whose shifted sources live in the next quadrant (lanes 28..31).  All 32 lanes gather: lane = (half, column)
          // sums, for each of the four target lanes, the kx terms it is missing (10 conflict-free LDS instead of 160 issue
          // slots spent on four active lanes); the targets then pick up their 32 sums with eight LDS.128.
          float* fx = fixb + warp * 128;                     // [4 targets][32 (half, column)]
          const float* src = xb + (((quad + 1) * 2 + (lane >> 4)) * 10) * 16 + (lane & 15);
#pragma unroll
          for (int tt = 0; tt < 4; ++tt) {                   // target lane 28 + tt misses kx = 4 - tt .. 4 (source lane tt + kx - 4)
            float sl = 0.f;
#pragma unroll
            for (int kx = 4 - tt; kx <= 4; ++kx) sl += src[((kx * (kx - 1)) / 2 + (tt + kx - 4)) * 16];
            fx[tt * 32 + lane] = sl;
          }
          __syncwarp();
          if (lane >= 28) {
            const float4* f4 = reinterpret_cast<const float4*>(fx + (lane - 28) * 32);
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 t = f4[half * 4 + i];
                acc[half][4 * i] += t.x; acc[half][4 * i + 1] += t.y; acc[half][4 * i + 2] += t.z; acc[half][4 * i + 3] += t.w;
              }
          }
          __syncwarp();
        }
#ifdef OP3_TC_TRACE
        const unsigned long long te4 = clock64();
#endif
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c4 = 4 * half + i;
              if (c4 >= out_chunks) continue;
              const float4 bb = *reinterpret_cast<const float4*>(&s_bias[16 * half + 4 * i]);
              float4 o = make_float4(fmaf(acc[half][4 * i], inv_scale, bb.x), fmaf(acc[half][4 * i + 1], inv_scale, bb.y),
                                     fmaf(acc[half][4 * i + 2], inv_scale, bb.z), fmaf(acc[half][4 * i + 3], inv_scale, bb.w));
              if (epi_b == EPI_ELU) {
                o.x = elu_tc(o.x); o.y = elu_tc(o.y); o.z = elu_tc(o.z); o.w = elu_tc(o.w);
              } else if (epi_b == EPI_MUL_ELUGRAD) {
                o.x *= elu_grad_from_out(ra[c4].x); o.y *= elu_grad_from_out(ra[c4].y);
                o.z *= elu_grad_from_out(ra[c4].z); o.w *= elu_grad_from_out(ra[c4].w);
              }
              if (accum) red_add_f4(&out[obase + c4 * plane], o); else out[obase + c4 * plane] = o;
              if (want_max) omax = fmaxf(omax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
            }
          }
        }
#ifdef OP3_TC_TRACE
        const unsigned long long te5 = clock64();
#endif
        asm volatile("bar.sync %0, 128;" :: "r"(1 + grp) : "memory");      // exchange rows read before the next tile rewrites them
#ifdef OP3_TC_TRACE
        if (tid == 0) {     // per-tile timeline of one epilogue warp: see tools/tc_trace.py
          atomicAdd(&g_tcq_epi[0], te1 - te0); atomicAdd(&g_tcq_epi[1], te2 - te1); atomicAdd(&g_tcq_epi[2], te3 - te2); atomicAdd(&g_tcq_epi[3], 1ull);
          atomicAdd(&g_tcq_epi[4], te4 - te3); atomicAdd(&g_tcq_epi[5], te5 - te4); atomicAdd(&g_tcq_epi[6], clock64() - te5);
        }
#endif
        }   // !LEAN
      }
      if (in.out_maxbits != nullptr) {                      // magnitudes of the written tensor, for its consumer's operand scale
        omax *= max_mul;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, o));
        if (lane == 0 && __float_as_uint(omax) > __ldcg(&in.out_maxbits[n])) atomicMax(&in.out_maxbits[n], __float_as_uint(omax));
      }
      g += T;
    }
  }
  tc_fence_before();
  if (PAIR) {
    cluster_sync_all();                                    // nobody leaves while its peer may still arrive on its barriers / use its TMEM
    if (warp == MMA_WARP) tmem_dealloc2<512>(tmem_base);
  } else {
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc<512>(tmem_base);
  }
}


// per-image largest magnitude of a conv input (fp32 bit pattern; non-negative floats order like unsigned ints)
__global__ void __launch_bounds__(256) tc_absmax_kernel(ConvInput in, int D, uint32_t* __restrict__ maxbits) {
  const int n = blockIdx.x, c = blockIdx.y;
  const float4* src = tc_src_chunk_ptr(in, c, n, D);
  float m = 0.f;
  if (src)
    for (int i = threadIdx.x; i < D * D; i += 256) {
      const float4 v = __ldg(src + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(&maxbits[n], __float_as_uint(m));
}

template <int NOUT, int MODE>
int launch_tc(cudaStream_t st, int N, int D, const ConvInput& in, const float* wt, const float* bias, int cout_real,
              float4* out, int out_planes, int out_chunk0, int out_chunks, int epi, const float4* ref,
              const uint32_t* maxbits, const float* wscale) {
  using Cfg = TcCfg<NOUT, MODE>;
  const int P = D + 4;
  int G = Cfg::GMAX;
  size_t smem = 0;
  for (; G >= 1; --G) {
    const int PH = 128 * G + 4 * P + 4;
    smem = (size_t)Cfg::NSTAGE * ((size_t)2 * PH * 16 * (Cfg::SPLIT3 ? 2 : 1) + 25 * 2 * Cfg::NB * 16);
    if (smem <= 226 * 1024) break;
  }
  OP3_REQUIRE(G >= 1, "conv5x5_tc: D=%d does not fit in shared memory", D);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_tc_kernel<NOUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  }
  const int g_num_sms = device_sm_count();
  const int tiles_img = (D * P + 127) / 128;
  const int groups_img = (tiles_img + G - 1) / G;
  const int n_items = N * groups_img;
  const int grid = n_items < g_num_sms ? n_items : g_num_sms;      // persistent: one CTA per SM
  conv5x5_tc_kernel<NOUT, MODE><<<grid, TC_THREADS, smem, st>>>(in, N, D, G, wt, bias, cout_real, out, out_planes,
                                                                out_chunk0, out_chunks, epi, ref, maxbits, wscale);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

// Strip decomposition of a D x D image for the quint-tap kernel: one item if 128 + 4 (D + 2) rows fit four chunks of the
// activation ring (D <= 90), else the fewest equal strips of W columns (W a multiple of 8) with 128 + 4 (W + 4) <= 4 * 124.
bool tq_strips(int D, int* W, int* nstrips) {
  if (128 + 4 * (D + TQ_PAD) <= 4 * TQ_TS) { *W = D; *nstrips = 1; return true; }
  for (int ns = 2; ns <= 16; ++ns) {
    if (D % ns) continue;
    const int w = D / ns;
    if (w % 8 == 0 && 128 + 4 * (w + 4) <= 4 * TQ_TS) { *W = w; *nstrips = ns; return true; }
  }
  return false;
}

// Tensor map of a layer's fp16 weight slab for the quint-tap kernel's TMA prologue (see the kernel), cached per slab.
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool tq_weight_tmap(const float* wt, int ngroups, CUtensorMap* out) {
  static TmapEncodeFn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (getenv("OP3_TCQ_NO_TMA") != nullptr) return false;          // A/B switch: the cooperative re-layout prologue
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<TmapEncodeFn>(fn);
  }
  if (encode == nullptr || (reinterpret_cast<uintptr_t>(wt) & 15u) != 0) return false;
  static std::unordered_map<uint64_t, CUtensorMap> cache;           // key: slab address ^ K groups (host single-threaded like the other caches)
  const uint64_t key = reinterpret_cast<uint64_t>(wt) ^ ((uint64_t)ngroups << 60);
  auto it = cache.find(key);
  if (it == cache.end()) {
    CUtensorMap m;
    const cuuint64_t gdim[5] = {256, 4, 5, 5, (cuuint64_t)ngroups};
    const cuuint64_t gstr[4] = {512, 2048, 10240, 51200};            // bytes, dimensions 1..4
    const cuuint32_t box[5] = {256, 1, 5, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<float*>(wt), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
    it = cache.emplace(key, m).first;
  }
  *out = it->second;
  return true;
}

int launch_tcq(cudaStream_t st, int N, int D, const ConvInput& in, const float* wt, const float* bias, int cout_real,
               float4* out, int out_planes, int out_chunk0, int out_chunks, int epi, const float4* ref,
               const uint32_t* maxbits, const float* wscale) {
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_tcq_kernel<false, 2, TQ_LEAN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TQ_SMEM));
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_tcq_kernel<false, 2, TQ_LEAN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TQ_SMEM));
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_tcq_kernel<false, 2, TQ_LEAN, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TQ_SMEM));
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_tcq_kernel<true, TQ_PAIR_NEG, TQ_LEAN, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TQ_SMEM_PAIR));
  }
  const int g_num_sms = device_sm_count();
  int W = 0, nstrips = 0;
  OP3_REQUIRE(tq_strips(D, &W, &nstrips), "conv5x5_tcq: D=%d has no strip decomposition", D);
  const int P = nstrips == 1 ? D + TQ_PAD : W + 4;
  const int tiles_img = (D * P + TQ_TS - 1) / TQ_TS;
  const long long items = (long long)N * nstrips;
  if (tcq_pair_enabled() && items % 2 == 0 && items >= 2 && !(epi & EPI_ACCUM)) {
    // CTA pairs (cta_group::2): one cluster of two CTAs per TPC, each pair owns a contiguous range of item-pair tiles
    const long long total = items / 2 * tiles_img;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(TqRoles<TQ_PAIR_NEG>::THREADS);
    cfg.dynamicSmemBytes = TQ_SMEM_PAIR;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent pairs: as many clusters as can be resident at once (a pair needs both SMs of a TPC), cached per device
    static int max_pairs[OP3_MAX_DEVICES] = {};
    int dev = 0;
    OP3_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < OP3_MAX_DEVICES && max_pairs[dev] == 0) {
      cfg.gridDim = dim3(2 * (g_num_sms / 2));
      int nc = 0;
      OP3_CHECK(cudaOccupancyMaxActiveClusters(&nc, conv5x5_tcq_kernel<true, TQ_PAIR_NEG, TQ_LEAN, true, false>, &cfg));
      max_pairs[dev] = nc > 0 ? nc : 1;
    }
    int pairs = (dev >= 0 && dev < OP3_MAX_DEVICES) ? max_pairs[dev] : g_num_sms / 2;
    if (total < pairs) pairs = (int)total;
    cfg.gridDim = dim3(2 * pairs);
    CUtensorMap nomap{};
    OP3_CHECK(cudaLaunchKernelEx(&cfg, conv5x5_tcq_kernel<true, TQ_PAIR_NEG, TQ_LEAN, true, false>, in, N, D, wt, bias, cout_real, out, out_planes, out_chunk0,
                                 out_chunks, epi, ref, maxbits, wscale, W, nstrips, nomap, 0));
    OP3_COUNT_LAUNCH();
    return OP3_OK;
  }
  const long long total = items * tiles_img;
  const int grid = total < g_num_sms ? (int)total : g_num_sms;
  CUtensorMap wmap{};
  const int use_tma = tq_weight_tmap(wt, (in.total_chunks + 3) / 4, &wmap) ? 1 : 0;
  if (epi & EPI_ACCUM) {
    OP3_REQUIRE(nstrips == 1, "conv5x5_tcq: the accumulating pass is built for whole images only (D=%d)", D);
    conv5x5_tcq_kernel<false, 2, TQ_LEAN, false, true><<<grid, TqRoles<2>::THREADS, TQ_SMEM, st>>>(
        in, N, D, wt, bias, cout_real, out, out_planes, out_chunk0, out_chunks, epi, ref, maxbits, wscale, W, nstrips, wmap, use_tma);
  } else if (nstrips == 1)
    conv5x5_tcq_kernel<false, 2, TQ_LEAN, false, false><<<grid, TqRoles<2>::THREADS, TQ_SMEM, st>>>(
        in, N, D, wt, bias, cout_real, out, out_planes, out_chunk0, out_chunks, epi, ref, maxbits, wscale, W, nstrips, wmap, use_tma);
  else
    conv5x5_tcq_kernel<false, 2, TQ_LEAN, true, false><<<grid, TqRoles<2>::THREADS, TQ_SMEM, st>>>(
        in, N, D, wt, bias, cout_real, out, out_planes, out_chunk0, out_chunks, epi, ref, maxbits, wscale, W, nstrips, wmap, use_tma);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}
}  // namespace

// the quint-tap kernel needs 128 + 4*P rows inside four 124-row chunks (whole images up to D = 90, column strips beyond)
// and at most two 16-channel K groups
bool tc_quint_fits(int D, int total_chunks) {
  int W, ns;
  return tq_strips(D, &W, &ns) && total_chunks <= 8;
}

#ifdef OP3_TC_TRACE
extern "C" int op3_debug_tcq_flags(int flags) { cudaMemcpyToSymbol(g_tcq_dbg, &flags, sizeof(int)); return 0; }
extern "C" int op3_debug_tcq_epi(unsigned long long* out /*[8] host*/, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_tcq_epi, sizeof(unsigned long long) * 8);
  if (reset) { static unsigned long long z[8]; cudaMemcpyToSymbol(g_tcq_epi, z, sizeof(z)); }
  return 0;
}
extern "C" int op3_debug_tc_trace(unsigned long long* out /*[256][4] host*/, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(unsigned long long) * 256 * 4);
  if (reset) { static unsigned long long z[256][4]; cudaMemcpyToSymbol(g_tc_trace, z, sizeof(z)); }
  return 0;
}
#endif

int tc_nout(int cout) { return cout <= 16 ? 16 : (cout <= 32 ? 32 : 64); }

size_t conv5x5_tc_weight_floats(int cin_pad, int cout) {
  const int groups = (cin_pad / 4 + 1) / 2;
  return (size_t)groups * 25 * 2 * (2 * tc_nout(cout)) * 4 + 8;   // sized for the 3x layout (hi and lo rows) / 2 fp16 slabs + scales
}

// one fp16 slab: [K group of 16 ci][tap 25][k-chunk 2][row 2*nout][8 halves], then 4 floats whose first is the weight scale
size_t conv5x5_tc_f16_slab_floats(int cin_pad, int nout) {
  const int groups16 = (cin_pad / 4 + 3) / 4;
  return (size_t)groups16 * 25 * 2 * (2 * nout) * 4 + 4;
}

size_t tc_scratch_floats(int N) { return 8 * tc_slot_floats(N); }

// Weight slab layout consumed by the kernel (built by op3_prepare): [K group of 8 ci][tap 25][k-chunk 2][row NB][4 ci]
// with NB = nout rows (1xTF32: rn_tf32(w)) or 2*nout rows (3xTF32: hi rows, then lo rows).  cout = 64 in the split modes is
// stored and run as two independent halves of 32 output channels (TMEM / smem budget).
int conv5x5_tc_forward(cudaStream_t st, int precision, int N, int D, const ConvInput& in, int cout, const float* wt,
                       const float* bias, float4* out, int epi, const float4* elu_ref, float* tc_scratch) {
  if (N <= 0) return OP3_OK;
  OP3_REQUIRE(cout % 4 == 0 && cout <= 64, "conv5x5_tc: unsupported cout=%d", cout);
  const int cin_real = in.real_cin ? in.real_cin : in.total_chunks * 4;
  ProfScope prof(st, KID_CONV_FWD, 2.0 * 25 * cin_real * (cout == 20 ? 17 : cout) * (double)N * D * D);
  const int nout = tc_nout(cout);
  const int planes = cout / 4;
  if (precision == OP3_PREC_F16X3) {
    OP3_REQUIRE(tc_scratch != nullptr, "conv5x5_tc: the fp16 mode needs tc_scratch");
    const uint32_t* maxbits = in.maxbits;
    if (maxbits == nullptr) {                              // the producer did not record them: reduce here (slot 0)
      uint32_t* mb = reinterpret_cast<uint32_t*>(tc_scratch);
      OP3_CHECK(cudaMemsetAsync(mb, 0, (size_t)N * sizeof(uint32_t), st));
      tc_absmax_kernel<<<dim3(N, in.total_chunks), 256, 0, st>>>(in, D, mb);
      OP3_COUNT_LAUNCH();
      OP3_LAUNCH_CHECK();
      maxbits = mb;
    }
    const int h_nout = (nout == 64 || (nout == 16 && tc_quint_fits(D, in.total_chunks))) ? 32 : nout;
    const size_t slab = conv5x5_tc_f16_slab_floats(in.total_chunks * 4, h_nout);
    const bool quint = tc_quint_fits(D, in.total_chunks);
    // narrow layers (cout <= 16) run on the quint-tap kernel with zero rows up to 32 output channels (op3_prepare builds
    // their slab that way when the kernel fits): its cost does not depend on cout, and it beats both the 16-column
    // per-tap kernel and the FFMA kernel
    if (nout == 16 && !quint)
      return launch_tc<16, 2>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref, maxbits, wt + slab - 4);
    if (nout == 16) return launch_tcq(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref, maxbits, wt + slab - 4);
    int tq_w = 0, tq_ns = 0;
    if (nout == 32 && in.total_chunks == 16 && in.nsrc == 1 && epi != EPI_ELU && tq_strips(D, &tq_w, &tq_ns) && tq_ns == 1) {
      // 64 input channels (dgrad of the refinement net's third conv): two K passes of 32 channels on the quint-tap kernel, whose
      // resident weights and activation ring hold two 16-channel K groups.  Pass 1 stores, pass 2 adds (EPI_ACCUM); both apply the
      // epilogue factor, which distributes over the sum ((a + b) * ELU' = a * ELU' + b * ELU').
      ConvInput lo = in, hi = in;
      lo.src[0].nchunks = hi.src[0].nchunks = 8;
      lo.src[0].pitch = hi.src[0].pitch = in.src[0].pitch ? in.src[0].pitch : 16;
      hi.src[0].ptr = in.src[0].ptr + (size_t)8 * D * D;
      lo.total_chunks = hi.total_chunks = 8;
      lo.real_cin = hi.real_cin = 0;
      const size_t kgroup = (size_t)25 * 2 * 64 * 4;      // floats per 16-channel K group of the fp16 slab
      OP3_TRY(launch_tcq(st, N, D, lo, wt, bias, cout, out, planes, 0, planes, epi | EPI_MAX2, elu_ref, maxbits, wt + slab - 4));
      return launch_tcq(st, N, D, hi, wt + 2 * kgroup, nullptr, cout, out, planes, 0, planes, epi | EPI_ACCUM | EPI_MAX2, elu_ref, maxbits,
                        wt + slab - 4);
    }
    if (nout == 32)
      return quint ? launch_tcq(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref, maxbits, wt + slab - 4)
                   : launch_tc<32, 2>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref, maxbits, wt + slab - 4);
    if (quint) {
      OP3_TRY(launch_tcq(st, N, D, in, wt, bias, 32, out, planes, 0, 8, epi, elu_ref, maxbits, wt + slab - 4));
      return launch_tcq(st, N, D, in, wt + slab, bias ? bias + 32 : nullptr, cout - 32, out, planes, 8, planes - 8, epi, elu_ref,
                        maxbits, wt + 2 * slab - 4);
    }
    OP3_TRY((launch_tc<32, 2>(st, N, D, in, wt, bias, 32, out, planes, 0, 8, epi, elu_ref, maxbits, wt + slab - 4)));
    return launch_tc<32, 2>(st, N, D, in, wt + slab, bias ? bias + 32 : nullptr, cout - 32, out, planes, 8, planes - 8, epi, elu_ref,
                            maxbits, wt + 2 * slab - 4);
  }
  if (precision == OP3_PREC_TF32X3) {
    if (nout == 16) return launch_tc<16, 1>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref, nullptr, nullptr);
    if (nout == 32) return launch_tc<32, 1>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref, nullptr, nullptr);
    const size_t half = (size_t)((in.total_chunks + 1) / 2) * 25 * 2 * 64 * 4;
    OP3_TRY((launch_tc<32, 1>(st, N, D, in, wt, bias, 32, out, planes, 0, 8, epi, elu_ref, nullptr, nullptr)));
    return launch_tc<32, 1>(st, N, D, in, wt + half, bias ? bias + 32 : nullptr, cout - 32, out, planes, 8, planes - 8, epi, elu_ref,
                            nullptr, nullptr);
  }
  if (nout == 16) return launch_tc<16, 0>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref, nullptr, nullptr);
  if (nout == 32) return launch_tc<32, 0>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref, nullptr, nullptr);
  return launch_tc<64, 0>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref, nullptr, nullptr);
}

// fp32 CUDA-core kernel in exact mode, tcgen05 kernel otherwise.  The 4-channel output layer of the decoder stays on the
// CUDA cores in every mode: an MMA costs the same shared-memory operand traffic for N = 8 as for N = 64, so the tensor
// kernel needs 0.38 ms where the FFMA kernel needs 0.21 ms (profiles/r01_launches_*.txt).
int conv5x5_auto(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float* wt_ffma, const float* wt_tc,
                 const float* bias, float4* out, int epi, const float4* elu_ref, float* tc_scratch) {
  const int prec = op3_get_precision();
  const bool narrow_on_tc = prec == OP3_PREC_F16X3 && tc_quint_fits(D, in.total_chunks);   // 0.12 ms vs 0.22 ms on the FMA pipe
  if (prec == OP3_PREC_FP32 || (cout == 4 && !narrow_on_tc)) return conv5x5_forward(st, N, D, in, cout, wt_ffma, bias, out, epi, elu_ref);
  return conv5x5_tc_forward(st, prec, N, D, in, cout, wt_tc, bias, out, epi, elu_ref, tc_scratch);
}

}  // namespace op3

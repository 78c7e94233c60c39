// tcgen05 weight gradient of the 5x5 convolution:
//     dW[tap][ci][co] = sum_{n,p} X[n][p + o_tap][ci] * G[n][p][co]        (p, o in padded-linear pixel space, pitch D+4)
// The contraction runs over PIXELS, so both operands must be K-major with pixels along K.  Measured on B200
// (tests/probes/probe_mnmajor.cu, probe_mn2.cu): kind::tf32 with MN-major (transposed) operands silently produces zeros
// for every swizzle mode, so the loaders transpose 4 pixels x 4 channels blocks in registers into
//     [pixel chunk of 4][channel row][4 pixels]            (16-byte rows, K-major, no swizzle).
// A 16-byte K chunk holds 4 pixels, so a tap offset can only be applied in whole chunks; the residual phase
// (o mod 4 = kx mod 4, the pitch is a multiple of 4) is absorbed by shifted COPIES of the operands:
//     X_sx[q] = X[q + sx], sx in {0,1}      G_sg[m] = G[m - sg], sg in {0,2}      =>  o = 4a + sg + sx covers every phase.
// The two X copies are stacked along M and the two G copies along N, so ONE M=128 MMA  [X_0; X_1] x [G_0 | G_2]
// accumulates FOUR taps (kx = 0..3) and a second one ([X_0; X_1] x G_0 one chunk lower) the fifth:
//     lanes   0..31  X_0 hi (ci)    lanes 32..63  X_0 lo      -> tap kx = sg      (3xTF32: columns [G_hi | G_lo])
//     lanes  64..95  X_1 hi         lanes 96..127 X_1 lo      -> tap kx = sg + 1
//   * One CTA owns one kernel ROW ky: its X window needs no halo; grid = 5 x S persistent CTAs over (image, 64-pixel
//     block) items; per-CTA partial slabs are reduced in a fixed order by a second kernel (deterministic).
//   * warps 0-15 load (two groups, one per smem stage; register double-buffered across items), transpose, round to
//     tf32, split hi/lo; warp 16 issues the MMAs; warps 0-3 run the epilogue (quadrant sums).  The bias gradient is
//     summed by the G loaders of the ky = 0 CTAs.
#include "common.cuh"
#include "tc_common.cuh"

namespace op3 {

namespace {

constexpr int WG_BP = 64;                    // pixels per block (8 K-steps of 8)
constexpr int WG_CH = WG_BP / 4;             // 16 pixel chunks per block
constexpr int WG_GCH = WG_CH + 1;            // G copies need one extra chunk (tap kx = 4 uses chunk offset a + 1)
constexpr int WG_LOADERS = 512;             // two groups of 8 loader warps, one per smem stage
constexpr int WG_THREADS = WG_LOADERS + 32;  // + the MMA issuer warp
constexpr int WG_A_ROWS = 129;               // 128 rows per A chunk, +1: chunk stride == 1 mod 8 rows -> conflict-free stores

template <int NCO, bool SPLIT3>
struct WgCfg {
  static constexpr int NB = SPLIT3 ? 2 * NCO : NCO;          // rows of one G copy = accumulator columns per tap pair
  static constexpr int B_ROWS = 2 * NB + 1;                  // [G_0 rows | G_2 rows] per chunk; == 1 mod 8
  static constexpr int A_BYTES = WG_CH * WG_A_ROWS * 16;
  static constexpr int G_BYTES = WG_GCH * B_ROWS * 16;
  static constexpr int STAGE_BYTES = A_BYTES + G_BYTES;
  static constexpr int TMEM_COLS = 3 * NB <= 64 ? 64 : (3 * NB <= 128 ? 128 : 256);
};

__device__ __forceinline__ void split_store(float4* hi_dst, float4* lo_dst, float a, float b, float c, float d, bool split) {
  float4 h = make_float4(tf32_rn(a), tf32_rn(b), tf32_rn(c), tf32_rn(d));
  *hi_dst = h;
  if (split) *lo_dst = make_float4(a - h.x, b - h.y, c - h.z, d - h.w);   // exact; the tensor core truncates it (2^-22 rel.)
}

template <int NCO, bool SPLIT3>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv5x5_wgrad_tc_kernel(ConvInput in, int N, int D, const float4* __restrict__ dpre, int g_planes, int g_chunk0,
                        int g_chunks, int cout_total, int co0, float* __restrict__ partial) {
  using Cfg = WgCfg<NCO, SPLIT3>;
  constexpr int NB = Cfg::NB, B_ROWS = Cfg::B_ROWS;
  const int P = D + 4;                                     // multiple of 4 (D is a multiple of 8)
  const int QN = (D + 4) * P;                              // padded-linear input pixels per image
  const int NBLK = (QN + WG_BP - 1) / WG_BP;
  const int ky = blockIdx.y;
  const int n_items = N * NBLK;
  const int cin_chunks = in.total_chunks;
  const int cin_pad = cin_chunks * 4;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;   // provably warp-uniform roles
  const int a0 = (ky * P) >> 2;                            // chunk offset of taps kx = 0..3; kx = 4 uses a0 + 1
  const bool have = n_items > (int)blockIdx.x;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[2], empty_bar[2], acc_bar;
  __shared__ uint32_t tmem_base_s;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], WG_LOADERS / 2); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == WG_LOADERS / 32) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  float4 bias_acc = make_float4(0.f, 0.f, 0.f, 0.f);       // loaders of ky == 0: sum_p G[p][4*warp ..]

  if (warp < WG_LOADERS / 32) {
    // ===================== loaders =====================
    // Two groups of 8 warps; group g fills stage g with every second item of this CTA.  Inside a group warps 0-3 build
    // the X operand and warps 4-7 the G operand; a warp covers TWO channel chunks (half-warps) x 16 pixel chunks, so all
    // 32 lanes work.  Every lane reads its own shifted pixels from global memory (L1 hits), there are no shuffles.
    const int grp = warp >> 3, wl = warp & 7;
    const int half = lane >> 4, j = lane & 15;
    const uint32_t magic = (uint32_t)((0x100000000ull + (uint32_t)P - 1) / (uint32_t)P);      // exact n / P for n < 2^24
    uint8_t* st = smem_raw + (size_t)grp * Cfg::STAGE_BYTES;
    const int first = blockIdx.x + grp * gridDim.x, stride = 2 * gridDim.x;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wl < 4) {
      const int cc = 2 * wl + half;                        // X channel chunk of this half-warp
      const bool act = cc < cin_chunks;
      float4* row = reinterpret_cast<float4*>(st) + (size_t)j * WG_A_ROWS + 4 * cc;
      auto load_x = [&](int item, float4* xv) {            // pixels q .. q+4 of the padded-linear input space
#pragma unroll
        for (int e = 0; e < 5; ++e) xv[e] = zero4;
        if (!act) return;
        const int n = item / NBLK, q = (item - n * NBLK) * WG_BP + 4 * j;
        const float4* src = tc_src_chunk_ptr(in, cc, n, D);
        const int r = (int)__umulhi((uint32_t)q, magic);
        const int iy = r - 2, ix = q - r * P - 2;          // q..q+3 share a row (q and P are multiples of 4)
        if (q < QN && (unsigned)iy < (unsigned)D) {
          const float4* rp = src + (size_t)iy * D + ix;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if ((unsigned)(ix + e) < (unsigned)D) xv[e] = __ldg(rp + e);
        }
        int ix4 = ix + 4, iy4 = iy;
        if (ix4 >= P - 2) { ix4 -= P; ++iy4; }
        if (q + 4 < QN && (unsigned)iy4 < (unsigned)D && (unsigned)ix4 < (unsigned)D) xv[4] = __ldg(src + (size_t)iy4 * D + ix4);
      };
      float4 xv[5], xn[5];
      if (first < n_items) load_x(first, xv);
      int k = 0;
      for (int item = first; item < n_items; item += stride, ++k) {
        if (item + stride < n_items) load_x(item + stride, xn);      // in flight while this item is stored
        if (k >= 1) mbar_wait(&empty_bar[grp], (k - 1) & 1);
        if (act) {
          // X_0 chunk j = pixels 4j .. 4j+3, X_1 chunk j = pixels 4j+1 .. 4j+4 (transposed: one 16-byte row per channel)
          split_store(row + 0, row + 32 + 0, xv[0].x, xv[1].x, xv[2].x, xv[3].x, SPLIT3);
          split_store(row + 1, row + 32 + 1, xv[0].y, xv[1].y, xv[2].y, xv[3].y, SPLIT3);
          split_store(row + 2, row + 32 + 2, xv[0].z, xv[1].z, xv[2].z, xv[3].z, SPLIT3);
          split_store(row + 3, row + 32 + 3, xv[0].w, xv[1].w, xv[2].w, xv[3].w, SPLIT3);
          split_store(row + 64 + 0, row + 96 + 0, xv[1].x, xv[2].x, xv[3].x, xv[4].x, SPLIT3);
          split_store(row + 64 + 1, row + 96 + 1, xv[1].y, xv[2].y, xv[3].y, xv[4].y, SPLIT3);
          split_store(row + 64 + 2, row + 96 + 2, xv[1].z, xv[2].z, xv[3].z, xv[4].z, SPLIT3);
          split_store(row + 64 + 3, row + 96 + 3, xv[1].w, xv[2].w, xv[3].w, xv[4].w, SPLIT3);
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[grp]);
#pragma unroll
        for (int e = 0; e < 5; ++e) xv[e] = xn[e];
      }
    } else {
      const int cc = 2 * (wl - 4) + half;                  // G channel chunk of this half-warp
      const bool act = cc < g_chunks;
      const int lj = j + 1;                                // local chunks 1..16; lane j == 0 also stores chunk 0 of G_0
      float4* G0 = reinterpret_cast<float4*>(st + Cfg::A_BYTES) + (size_t)lj * B_ROWS + 4 * cc;
      float4* G2 = G0 + NB;
      auto load_g = [&](int item, float4* gv, float4* gm) {  // gv: pixels p-2 .. p+3, gm: pixels p-4, p-3 (lane j == 0)
#pragma unroll
        for (int e = 0; e < 6; ++e) gv[e] = zero4;
        gm[0] = zero4; gm[1] = zero4;
        if (!act) return;
        const int n = item / NBLK, Q0 = (item - n * NBLK) * WG_BP;
        const float4* src = dpre + ((size_t)n * g_planes + g_chunk0 + cc) * D * D;
        const int p = Q0 - 4 * (a0 + 1) + 4 * lj;
        if (p >= 0) {
          const int y = (int)__umulhi((uint32_t)p, magic), x = p - y * P;
          if (y < D) {
            const float4* rp = src + (size_t)y * D + x;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (x + e < D) gv[2 + e] = __ldg(rp + e);
          }
        }
        const int pm = p - 4;
        if (pm >= 0) {
          const int y = (int)__umulhi((uint32_t)pm, magic), x = pm - y * P;
          if (y < D) {
            const float4* rp = src + (size_t)y * D + x;
            if (x + 2 < D) gv[0] = __ldg(rp + 2);
            if (x + 3 < D) gv[1] = __ldg(rp + 3);
            if (j == 0) {
              if (x < D) gm[0] = __ldg(rp);
              if (x + 1 < D) gm[1] = __ldg(rp + 1);
            }
          }
        }
      };
      float4 gv[6], gm[2], gn[6], gmn[2];
      if (first < n_items) load_g(first, gv, gm);
      int k = 0;
      for (int item = first; item < n_items; item += stride, ++k) {
        if (item + stride < n_items) load_g(item + stride, gn, gmn);
        if (k >= 1) mbar_wait(&empty_bar[grp], (k - 1) & 1);
        if (act) {
          // G_0 chunk lj = pixels p .. p+3, G_2 chunk lj = pixels p-2 .. p+1
          split_store(G0 + 0, G0 + NCO + 0, gv[2].x, gv[3].x, gv[4].x, gv[5].x, SPLIT3);
          split_store(G0 + 1, G0 + NCO + 1, gv[2].y, gv[3].y, gv[4].y, gv[5].y, SPLIT3);
          split_store(G0 + 2, G0 + NCO + 2, gv[2].z, gv[3].z, gv[4].z, gv[5].z, SPLIT3);
          split_store(G0 + 3, G0 + NCO + 3, gv[2].w, gv[3].w, gv[4].w, gv[5].w, SPLIT3);
          split_store(G2 + 0, G2 + NCO + 0, gv[0].x, gv[1].x, gv[2].x, gv[3].x, SPLIT3);
          split_store(G2 + 1, G2 + NCO + 1, gv[0].y, gv[1].y, gv[2].y, gv[3].y, SPLIT3);
          split_store(G2 + 2, G2 + NCO + 2, gv[0].z, gv[1].z, gv[2].z, gv[3].z, SPLIT3);
          split_store(G2 + 3, G2 + NCO + 3, gv[0].w, gv[1].w, gv[2].w, gv[3].w, SPLIT3);
          if (j == 0) {                                    // chunk 0 of G_0 (tap kx = 4 reads G_0 one chunk lower)
            float4* Gz = G0 - B_ROWS;
            split_store(Gz + 0, Gz + NCO + 0, gm[0].x, gm[1].x, gv[0].x, gv[1].x, SPLIT3);
            split_store(Gz + 1, Gz + NCO + 1, gm[0].y, gm[1].y, gv[0].y, gv[1].y, SPLIT3);
            split_store(Gz + 2, Gz + NCO + 2, gm[0].z, gm[1].z, gv[0].z, gv[1].z, SPLIT3);
            split_store(Gz + 3, Gz + NCO + 3, gm[0].w, gm[1].w, gv[0].w, gv[1].w, SPLIT3);
          }
          if (ky == 0) {                                   // chunks 1..16 tile the pixel space exactly once over the blocks
#pragma unroll
            for (int e = 2; e < 6; ++e) {
              bias_acc.x += gv[e].x; bias_acc.y += gv[e].y; bias_acc.z += gv[e].z; bias_acc.w += gv[e].w;
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[grp]);
#pragma unroll
        for (int e = 0; e < 6; ++e) gv[e] = gn[e];
        gm[0] = gmn[0]; gm[1] = gmn[1];
      }
    }
  } else {
    // whole warp in uniform control flow, tcgen05 instructions on the elect.sync leader (uniform-register descriptors)
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    // ===================== MMA issuer: per 8 pixels, [X_0;X_1] x [G_0|G_2](a0) (N = 2 NB) and [X_0;X_1] x G_0(a0+1) ==========
    // An MMA's cost is its shared-memory operand traffic ((4 KB of A + 32 B x N of B) / 128 B per clock for N <= 128), so
    // the two G copies share one read of A.
    const uint32_t idesc2 = make_idesc_tf32(128, 2 * NB), idesc1 = make_idesc_tf32(128, NB);
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait(&full_bar[s], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem_raw + (size_t)s * Cfg::STAGE_BYTES);
      const uint64_t ad = make_smem_desc(st, WG_A_ROWS, 8);
      const uint32_t a_lo32 = (uint32_t)ad, a_hi32 = (uint32_t)(ad >> 32);
      // local G chunk of X chunk jx is jx + (a0 + 1) - a:  taps kx = 0..3 -> +1, tap kx = 4 -> +0
      const uint64_t bd2 = make_smem_desc(st + Cfg::A_BYTES + B_ROWS * 16, B_ROWS, 8);
      const uint64_t bd1 = make_smem_desc(st + Cfg::A_BYTES, B_ROWS, 8);
      const uint32_t b2_lo32 = (uint32_t)bd2, b2_hi32 = (uint32_t)(bd2 >> 32);
      const uint32_t b1_lo32 = (uint32_t)bd1, b1_hi32 = (uint32_t)(bd1 >> 32);
#pragma unroll
      for (int ks = 0; ks < WG_BP / 8; ++ks) {
        const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
        if (leader) mma_tf32(tmem_u, a_lo32 + (uint32_t)(2 * ks * WG_A_ROWS), a_hi32, b2_lo32 + (uint32_t)(2 * ks * B_ROWS), b2_hi32,
                 idesc2, acc);
        if (leader) mma_tf32(tmem_u + (uint32_t)(2 * NB), a_lo32 + (uint32_t)(2 * ks * WG_A_ROWS), a_hi32,
                 b1_lo32 + (uint32_t)(2 * ks * B_ROWS), b1_hi32, idesc1, acc);
      }
      if (leader) tc_commit(&empty_bar[s]);
      __syncwarp();
    }
    if (leader) tc_commit(&acc_bar);
  }
  // ===================== epilogue: quadrant sums -> per-CTA partial slab =====================
  __syncthreads();                      // loaders are done with smem; reuse it as the combine buffer
  float* red = reinterpret_cast<float*>(smem_raw);         // [3 MMAs][2 taps][32 ci][NCO]: sums of the X_lo rows
  float* s_bias = red + 3 * 2 * 32 * NCO;                  // [2 groups][8 chunks][4]
  float* slab = partial + (size_t)blockIdx.x * ((size_t)25 * cin_pad * cout_total + cout_total);
  if (warp < WG_LOADERS / 32 && have) {  // nobody may touch smem before the last MMAs have consumed their operands
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
  }
  if (warp < WG_LOADERS / 32 && (warp & 7) >= 4) {   // bias partial of this CTA (ky == 0 only; fixed shuffle tree per half-warp)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      bias_acc.x += __shfl_xor_sync(0xffffffffu, bias_acc.x, o); bias_acc.y += __shfl_xor_sync(0xffffffffu, bias_acc.y, o);
      bias_acc.z += __shfl_xor_sync(0xffffffffu, bias_acc.z, o); bias_acc.w += __shfl_xor_sync(0xffffffffu, bias_acc.w, o);
    }
    if ((lane & 15) == 0) {
      float* dst = s_bias + ((warp >> 3) * 8 + 2 * ((warp & 7) - 4) + (lane >> 4)) * 4;
      dst[0] = bias_acc.x; dst[1] = bias_acc.y; dst[2] = bias_acc.z; dst[3] = bias_acc.w;
    }
  }
  // lanes 32..63 (warp 1) and 96..127 (warp 3) hold the X_lo rows: park their sums in smem
  if (SPLIT3 && (warp == 1 || warp == 3)) {
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int c0 = 0; c0 < NCO; c0 += 16) {
        float v[16], u[16];
        const uint32_t t = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(m * NB + c0);
        if (have) { tmem_ld16(t, v); tmem_ld16(t + NCO, u); tmem_ld_wait(); }
#pragma unroll
        for (int i = 0; i < 16; ++i) red[((size_t)(m * 2 + (warp >> 1)) * 32 + lane) * NCO + c0 + i] = have ? v[i] + u[i] : 0.f;
      }
  }
  __syncthreads();
  if (warp == 0 || warp == 2) {          // lanes 0..31 / 64..95: X_hi rows of tap kx = 2m / 2m + 1
    const int sub = warp >> 1;
    for (int m = 0; m < 3; ++m) {
      const int kx = 2 * m + sub;
      if (kx > 4) continue;              // the upper half of the third MMA is a non-existent tap
#pragma unroll
      for (int c0 = 0; c0 < NCO; c0 += 16) {
        float v[16], u[16];
        const uint32_t t = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(m * NB + c0);
        if (have) {
          tmem_ld16(t, v);
          if (SPLIT3) tmem_ld16(t + NCO, u);
          tmem_ld_wait();
        }
        if (lane < cin_pad) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int co = c0 + i;
            if (co < g_chunks * 4) {
              float r = have ? v[i] : 0.f;
              if (SPLIT3) r += (have ? u[i] : 0.f) + red[((size_t)(m * 2 + sub) * 32 + lane) * NCO + co];
              slab[((size_t)(ky * 5 + kx) * cin_pad + lane) * cout_total + co0 + co] = r;
            }
          }
        }
      }
    }
  }
  if (ky == 0 && tid < g_chunks * 4)
    slab[(size_t)25 * cin_pad * cout_total + co0 + tid] = s_bias[tid] + s_bias[32 + tid];   // the two loader groups
  tc_fence_before();
  __syncthreads();
  if (warp == WG_LOADERS / 32) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// Sum of the per-CTA partial slabs.  A block owns 64 outputs; its four 64-thread groups each walk a quarter of the slabs
// (coalesced 256-byte rows, two chains in flight) and the first group adds the four partials in a fixed order, so the
// result is deterministic and four times as many loads are in flight as with one thread per output.
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float* __restrict__ partial, int S, int n_w, int n_b,
                                                              float* __restrict__ dwt, float* __restrict__ dbias) {
  __shared__ float s_part[4][64];
  const int q = threadIdx.x >> 6, j = threadIdx.x & 63;
  const int i = blockIdx.x * 64 + j;
  const int tot = n_w + n_b;
  const int per = (S + 3) / 4, k0 = q * per, k1 = min(S, k0 + per);
  float s0 = 0.f, s1 = 0.f;
  if (i < tot) {
    int k = k0;
    for (; k + 2 <= k1; k += 2) {
      s0 += partial[(size_t)k * tot + i];
      s1 += partial[(size_t)(k + 1) * tot + i];
    }
    if (k < k1) s0 += partial[(size_t)k * tot + i];
  }
  s_part[q][j] = s0 + s1;
  __syncthreads();
  if (q == 0 && i < tot) {
    const float s = (s_part[0][j] + s_part[1][j]) + (s_part[2][j] + s_part[3][j]);
    if (i < n_w) dwt[i] += s;
    else if (dbias) dbias[i - n_w] += s;
  }
}


// =====================================================================================================================
// fp16 hi/lo-split variant (precision mode f16x3).  Same contraction, K = 16 pixels per MMA: a 16-byte K chunk holds 8
// pixels, the row pitch is P = D + 8 (a multiple of 8) so a tap offset o = ky*P + kx = 8a + kx never carries into the
// chunk index, and the phases kx = sx + sg are covered by X shifts sx in {0,1} (stacked along M) and G shifts sg in
// {0,2,4} (stacked along N; even shifts keep half2 pixel pairs aligned): ONE M=128 x N=192 MMA per 16 pixels accumulates
// all five taps of the kernel row (the sixth combination and the lo x lo quadrant are discarded).  Operands are scaled
// by per-tensor powers of two (the sum runs over all images, so one scale per tensor) and split into fp16 hi + lo.
//     lanes   0..31  X_0 hi    32..63  X_0 lo    64..95  X_1 hi    96..127 X_1 lo
//     columns 64*c + [0,32) G_{2c} hi,  64*c + [32,64) G_{2c} lo,   c = 0,1,2
// =====================================================================================================================
constexpr int HG_BP = 64;                    // pixels per item (4 K-steps of 16)
constexpr int HG_CH = HG_BP / 8;             // 8 pixel chunks of 8
constexpr int HG_GROUPS = 3;                 // loader groups = smem stages; group g fills stage g with every third item
constexpr int HG_GWARPS = 4;                 // warps per group: 2 build X (8 channel chunks x 8 pixel chunks), 2 build G
constexpr int HG_LOADERS = HG_GROUPS * HG_GWARPS * 32;
constexpr int HG_THREADS = HG_LOADERS + 32;  // 13 warps: 128 registers per thread are available (16-warp granularity)
constexpr int HG_A_ROWS = 129;               // 128 + 1 (== 1 mod 8)
constexpr int HG_B_ROWS = 193;               // 3 copies x (32 hi + 32 lo) + 1 (== 1 mod 8)
constexpr int HG_A_BYTES = HG_CH * HG_A_ROWS * 16;
constexpr int HG_B_BYTES = HG_CH * HG_B_ROWS * 16;
constexpr int HG_STAGE_BYTES = HG_A_BYTES + HG_B_BYTES;

// per-tensor largest magnitude (fp32 bit pattern) of a chunked activation tensor
__global__ void __launch_bounds__(256) wg_absmax_kernel(ConvInput in, int D, uint32_t* __restrict__ out) {
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
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// largest of n per-image maxima (every thread computes the same value; n is a few hundred words, L1/L2 resident)
__device__ __forceinline__ uint32_t wg_tensor_max(const uint32_t* __restrict__ m, int n) {
  uint32_t v = 0u;
  for (int i = (int)(threadIdx.x & 31); i < n; i += 32) v = max(v, __ldg(m + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// two scaled floats -> packed fp16 pair hi and the pair lo = rn(value - hi)
__device__ __forceinline__ void hg_pair(float a, float b, float sc, uint32_t& hi, uint32_t& lo) { f16_split2(a * sc, b * sc, hi, lo); }
// (a.hi16, b.lo16): the pixel pair shifted by one pixel
__device__ __forceinline__ uint32_t hg_shift1(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5432); }

__global__ void __launch_bounds__(HG_THREADS, 1)
conv5x5_wgrad_f16_kernel(ConvInput in, int N, int D, const float4* __restrict__ dpre, int g_planes, int g_chunk0, int g_chunks,
                         int cout_total, int co0, const uint32_t* __restrict__ xmax, int xmax_n,
                         const uint32_t* __restrict__ gmax, int gmax_n, float* __restrict__ partial) {
  constexpr int NCO = 32;
  const int P = D + 8;                                     // multiple of 8
  const int QN = (D + 4) * P;                              // padded-linear input pixels per image
  const int NBLK = (QN + HG_BP - 1) / HG_BP;
  const int ky = blockIdx.y;
  const int n_items = N * NBLK;
  const int cin_chunks = in.total_chunks;
  const int cin_pad = cin_chunks * 4;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;   // provably warp-uniform roles
  const int a0 = (ky * P) >> 3;                            // chunk offset of this kernel row
  const bool have = n_items > (int)blockIdx.x;
  const float sx_scale = __uint_as_float(f16_scale_bits(wg_tensor_max(xmax, xmax_n)));     // one scale per TENSOR: the sum runs
  const float sg_scale = __uint_as_float(f16_scale_bits(wg_tensor_max(gmax, gmax_n)));     // over all images

  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[HG_GROUPS], empty_bar[HG_GROUPS], acc_bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int MMA_WARP = HG_LOADERS / 32;

  if (tid == 0) {
    for (int s = 0; s < HG_GROUPS; ++s) { mbar_init(&full_bar[s], HG_GWARPS * 32); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc<256>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  float4 bias_acc = make_float4(0.f, 0.f, 0.f, 0.f);

  if (warp < MMA_WARP) {
    // ===================== loaders: group g fills stage g with every third item =====================
    const int grp = warp / HG_GWARPS, wl = warp % HG_GWARPS;
    const int t64 = (wl & 1) * 32 + lane;                  // 64 lane tasks per operand: 8 channel chunks x 8 pixel chunks
    const int cc = t64 >> 3, j = t64 & 7;
    const uint32_t magic = (uint32_t)((0x100000000ull + (uint32_t)P - 1) / (uint32_t)P);
    uint8_t* st = smem_raw + (size_t)grp * HG_STAGE_BYTES;
    const int first = blockIdx.x + grp * gridDim.x, stride = HG_GROUPS * gridDim.x;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wl < 2) {
      const bool act = cc < cin_chunks;                    // X channel chunk cc, pixel chunk j
      uint4* row = reinterpret_cast<uint4*>(st) + (size_t)j * HG_A_ROWS + 4 * cc;
      auto load_x = [&](int item, float4* xv) {            // pixels q .. q+8 of the padded-linear input space
#pragma unroll
        for (int e = 0; e < 9; ++e) xv[e] = zero4;
        if (!act) return;
        const int n = item / NBLK, q = (item - n * NBLK) * HG_BP + 8 * j;
        const float4* src = tc_src_chunk_ptr(in, cc, n, D);
        const int r = (int)__umulhi((uint32_t)q, magic);
        const int iy = r - 2, ix = q - r * P - 2;          // q..q+7 share a row (q and P are multiples of 8)
        if (q < QN && (unsigned)iy < (unsigned)D) {
          const float4* rp = src + (size_t)iy * D + ix;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if ((unsigned)(ix + e) < (unsigned)D) xv[e] = __ldg(rp + e);
        }
        int ix8 = ix + 8, iy8 = iy;
        if (ix8 >= P - 2) { ix8 -= P; ++iy8; }
        if (q + 8 < QN && (unsigned)iy8 < (unsigned)D && (unsigned)ix8 < (unsigned)D) xv[8] = __ldg(src + (size_t)iy8 * D + ix8);
      };
      float4 xv[9], xn[9];
      if (first < n_items) load_x(first, xv);
      int k = 0;
      for (int item = first; item < n_items; item += stride, ++k) {
        if (item + stride < n_items) load_x(item + stride, xn);
        if (k >= 1) mbar_wait(&empty_bar[grp], (k - 1) & 1);
        if (act) {
#define OP3_HG_X_CHANNEL(C, R)                                                                                   \
          {                                                                                                      \
            uint32_t h01, l01, h23, l23, h45, l45, h67, l67, h8, l8;                                             \
            hg_pair(xv[0].C, xv[1].C, sx_scale, h01, l01);                                                       \
            hg_pair(xv[2].C, xv[3].C, sx_scale, h23, l23);                                                       \
            hg_pair(xv[4].C, xv[5].C, sx_scale, h45, l45);                                                       \
            hg_pair(xv[6].C, xv[7].C, sx_scale, h67, l67);                                                       \
            hg_pair(xv[8].C, 0.f, sx_scale, h8, l8);                                                             \
            row[R] = make_uint4(h01, h23, h45, h67);                                                             \
            row[32 + R] = make_uint4(l01, l23, l45, l67);                                                        \
            row[64 + R] = make_uint4(hg_shift1(h01, h23), hg_shift1(h23, h45), hg_shift1(h45, h67), hg_shift1(h67, h8)); \
            row[96 + R] = make_uint4(hg_shift1(l01, l23), hg_shift1(l23, l45), hg_shift1(l45, l67), hg_shift1(l67, l8)); \
          }
          OP3_HG_X_CHANNEL(x, 0)
          OP3_HG_X_CHANNEL(y, 1)
          OP3_HG_X_CHANNEL(z, 2)
          OP3_HG_X_CHANNEL(w, 3)
#undef OP3_HG_X_CHANNEL
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[grp]);
#pragma unroll
        for (int e = 0; e < 9; ++e) xv[e] = xn[e];
      }
    } else {
      const bool act = cc < g_chunks;                      // G channel chunk cc, pixel chunk j
      uint4* row = reinterpret_cast<uint4*>(st + HG_A_BYTES) + (size_t)j * HG_B_ROWS + 4 * cc;
      // pixels p0-4 .. p0+7 of the output-gradient space (chunk j - a0 of this block, plus the 4 pixels before it)
      auto load_g = [&](int item, float4* gv) {
#pragma unroll
        for (int e = 0; e < 12; ++e) gv[e] = zero4;
        if (!act) return;
        const int n = item / NBLK, Q0 = (item - n * NBLK) * HG_BP;
        const float4* src = dpre + ((size_t)n * g_planes + g_chunk0 + cc) * D * D;
        const int p0 = Q0 + 8 * (j - a0);
        if (p0 >= 0) {
          const int y = (int)__umulhi((uint32_t)p0, magic), x = p0 - y * P;
          if (y < D) {
            const float4* rp = src + (size_t)y * D + x;
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (x + e < D) gv[4 + e] = __ldg(rp + e);
          }
        }
        const int pm = p0 - 8;                             // the chunk before: only its last 4 pixels are needed
        if (pm >= 0) {
          const int y = (int)__umulhi((uint32_t)pm, magic), x = pm - y * P;
          if (y < D) {
            const float4* rp = src + (size_t)y * D + x;
#pragma unroll
            for (int e = 4; e < 8; ++e)
              if (x + e < D) gv[e - 4] = __ldg(rp + e);
          }
        }
      };
      float4 gv[12], gn[12];
      if (first < n_items) load_g(first, gv);
      int k = 0;
      for (int item = first; item < n_items; item += stride, ++k) {
        if (item + stride < n_items) load_g(item + stride, gn);      // in flight while this item is converted and stored
        if (k >= 1) mbar_wait(&empty_bar[grp], (k - 1) & 1);
        if (act) {
#define OP3_HG_G_CHANNEL(C, R)                                                                                   \
          {                                                                                                      \
            uint32_t hm2, lm2, hm1, lm1, h0, l0, h1, l1, h2, l2, h3, l3;                                         \
            hg_pair(gv[0].C, gv[1].C, sg_scale, hm2, lm2);                                                       \
            hg_pair(gv[2].C, gv[3].C, sg_scale, hm1, lm1);                                                       \
            hg_pair(gv[4].C, gv[5].C, sg_scale, h0, l0);                                                         \
            hg_pair(gv[6].C, gv[7].C, sg_scale, h1, l1);                                                         \
            hg_pair(gv[8].C, gv[9].C, sg_scale, h2, l2);                                                         \
            hg_pair(gv[10].C, gv[11].C, sg_scale, h3, l3);                                                       \
            row[R] = make_uint4(h0, h1, h2, h3);              /* G_0: pixels p0 .. p0+7     */                    \
            row[32 + R] = make_uint4(l0, l1, l2, l3);                                                            \
            row[64 + R] = make_uint4(hm1, h0, h1, h2);        /* G_2: pixels p0-2 .. p0+5   */                    \
            row[96 + R] = make_uint4(lm1, l0, l1, l2);                                                           \
            row[128 + R] = make_uint4(hm2, hm1, h0, h1);      /* G_4: pixels p0-4 .. p0+3   */                    \
            row[160 + R] = make_uint4(lm2, lm1, l0, l1);                                                         \
          }
          OP3_HG_G_CHANNEL(x, 0)
          OP3_HG_G_CHANNEL(y, 1)
          OP3_HG_G_CHANNEL(z, 2)
          OP3_HG_G_CHANNEL(w, 3)
#undef OP3_HG_G_CHANNEL
          if (ky == 0) {                                   // a0 = 0: the chunks tile the pixel space exactly once
#pragma unroll
            for (int e = 4; e < 12; ++e) {
              bias_acc.x += gv[e].x; bias_acc.y += gv[e].y; bias_acc.z += gv[e].z; bias_acc.w += gv[e].w;
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[grp]);
#pragma unroll
        for (int e = 0; e < 12; ++e) gv[e] = gn[e];
      }
    }
  } else {
    // whole warp in uniform control flow, tcgen05 instructions on the elect.sync leader (uniform-register descriptors)
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    // ===================== MMA issuer: [X_0; X_1] x [G_0 | G_2 | G_4], K = 16 pixels =====================
    const uint32_t idesc = make_idesc_f16(128, 192);
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it % HG_GROUPS;
      mbar_wait(&full_bar[s], (it / HG_GROUPS) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem_raw + (size_t)s * HG_STAGE_BYTES);
      const uint64_t ad = make_smem_desc(st, HG_A_ROWS, 8);
      const uint64_t bd = make_smem_desc(st + HG_A_BYTES, HG_B_ROWS, 8);
      const uint32_t a_lo32 = (uint32_t)ad, a_hi32 = (uint32_t)(ad >> 32);
      const uint32_t b_lo32 = (uint32_t)bd, b_hi32 = (uint32_t)(bd >> 32);
#pragma unroll
      for (int ks = 0; ks < HG_CH / 2; ++ks)
        if (leader) mma_f16(tmem_u, a_lo32 + (uint32_t)(2 * ks * HG_A_ROWS), a_hi32, b_lo32 + (uint32_t)(2 * ks * HG_B_ROWS), b_hi32, idesc,
                (it > 0 || ks > 0) ? 1u : 0u);
      if (leader) tc_commit(&empty_bar[s]);
      __syncwarp();
    }
    if (leader) tc_commit(&acc_bar);
  }
  // ===================== epilogue =====================
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);         // [3 copies][2 shifts][32 ci][NCO]: the X_lo x G_hi sums
  float* s_bias = red + 3 * 2 * 32 * NCO;                  // [3 groups][8 chunks][4]
  float* slab = partial + (size_t)blockIdx.x * ((size_t)25 * cin_pad * cout_total + cout_total);
  const float inv = 1.f / (sx_scale * sg_scale);           // powers of two: exact
  if (warp < MMA_WARP && have) {
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
  }
  if (warp < MMA_WARP && (warp % HG_GWARPS) >= 2) {       // G warps: 8 lanes (pixel chunks) share a channel chunk
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      bias_acc.x += __shfl_xor_sync(0xffffffffu, bias_acc.x, o); bias_acc.y += __shfl_xor_sync(0xffffffffu, bias_acc.y, o);
      bias_acc.z += __shfl_xor_sync(0xffffffffu, bias_acc.z, o); bias_acc.w += __shfl_xor_sync(0xffffffffu, bias_acc.w, o);
    }
    if ((lane & 7) == 0) {
      const int cc = ((warp % HG_GWARPS) & 1) * 4 + (lane >> 3);
      float* dst = s_bias + ((warp / HG_GWARPS) * 8 + cc) * 4;
      dst[0] = bias_acc.x; dst[1] = bias_acc.y; dst[2] = bias_acc.z; dst[3] = bias_acc.w;
    }
  }
  if (warp == 1 || warp == 3) {          // lanes 32..63 / 96..127: X_lo rows, only their G_hi columns are used
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int c0 = 0; c0 < NCO; c0 += 16) {
        float v[16];
        const uint32_t t = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 64 + c0);
        if (have) { tmem_ld16(t, v); tmem_ld_wait(); }
#pragma unroll
        for (int i = 0; i < 16; ++i) red[((size_t)(c * 2 + (warp >> 1)) * 32 + lane) * NCO + c0 + i] = have ? v[i] : 0.f;
      }
  }
  __syncthreads();
  if (warp == 0 || warp == 2) {          // lanes 0..31 / 64..95: X_hi rows of tap kx = 2c / 2c + 1
    const int sub = warp >> 1;
    for (int c = 0; c < 3; ++c) {
      const int kx = 2 * c + sub;
      if (kx > 4) continue;
#pragma unroll
      for (int c0 = 0; c0 < NCO; c0 += 16) {
        float v[16], u[16];
        const uint32_t t = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 64 + c0);
        if (have) { tmem_ld16(t, v); tmem_ld16(t + NCO, u); tmem_ld_wait(); }
        if (lane < cin_pad) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int co = c0 + i;
            if (co < g_chunks * 4) {
              const float r = have ? v[i] + u[i] + red[((size_t)(c * 2 + sub) * 32 + lane) * NCO + co] : 0.f;
              slab[((size_t)(ky * 5 + kx) * cin_pad + lane) * cout_total + co0 + co] = r * inv;
            }
          }
        }
      }
    }
  }
  if (ky == 0 && tid < g_chunks * 4)
    slab[(size_t)25 * cin_pad * cout_total + co0 + tid] = s_bias[tid] + s_bias[32 + tid] + s_bias[64 + tid];
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<256>(tmem_base);
}


// =====================================================================================================================
// Ring-buffered fp16 variant: ONE CTA accumulates all five kernel rows, so every X / G element is loaded, scaled, split
// and staged ONCE instead of once per kernel row (the per-row kernel above spends its time in the loaders: 5x redundant
// conversions).  The G operand of kernel row ky for X chunk c is G chunk c - 9*ky... in general c - ky*P/8: the last
// 4*P/8 + 8 G chunks live in a shared-memory ring (57 slots of [G_0hi G_2hi G_4hi | G_0lo G_2lo G_4lo] rows); per 64-pixel
// block and kernel row the issuer runs 4 K-steps x 2 MMAs (B = the 96 hi rows, then the 96 lo rows, same accumulator
// columns: D += A*G_hi + A*G_lo), N = 96, accumulators at TMEM columns 96*ky (480 of 512).
// Each CTA owns a contiguous range of the global (image-major) block sequence; where the range enters an image it first
// refills the ring with the 5 blocks of G before its first block (steps with no MMAs).  Ring slots are numbered by a
// running per-CTA counter, so a slot is rewritten 7 steps after it was filled, two steps after its last reader (the same
// empty barrier that guards the two A stages).  The X block of a step is copied to TMEM (tcgen05.cp) and read from there.
// =====================================================================================================================
#ifdef OP3_TC_TRACE
__device__ unsigned long long g_ring_trace[256][8];   // issuer: operand wait, loop total, MMAs, steps; one G-loader warp: fence, convert, wait, stores
#endif
#ifndef OP3_RING_SMEM_A
#define OP3_RING_TMEM_A 1     /* X block staged in TMEM with tcgen05.cp (define OP3_RING_SMEM_A for the smem-A form) */
#endif
__host__ __device__ inline int ring_vb0(int P) { return (2 * P) / HG_BP; }                         // leading all-halo blocks of an item
__host__ __device__ inline int ring_live_blocks(int D, int P) { return ((D + 2) * P + HG_BP - 1) / HG_BP - ring_vb0(P); }
constexpr int RB_RING = 56;                  // G chunks in the ring (+ 1 guard slot = copy of slot 0 for wrapped K pairs)
constexpr int RB_PRE = 5;                    // refill steps per item: 5 blocks = 40 chunks >= 4*P/8 = 36 (D = 64)
constexpr int RB_LOADERS = 2 * 4 * 32;       // two groups (even / odd steps) of 2 X warps + 2 G warps
constexpr int RB_THREADS = RB_LOADERS + 32;
// Operand chunks are DENSE (128 / 192 rows of 16 bytes): every 8-row core matrix is one aligned 128-byte line, so the
// tensor core fetches it in one shared-memory wavefront (a padded row count would make most core matrices straddle two
// lines and double the operand-fetch time).  Store conflicts are avoided by the row order instead: channel 4*cc + r of a
// 32-channel block lives in row 8*r + cc, so the 8 lanes of a quarter warp (cc = 0..7, same pixel chunk) write 8
// consecutive rows = one 128-byte line per store instruction.
constexpr int RB_A_ROWS = 128, RB_B_ROWS = 192;
constexpr int RB_A_BYTES = HG_CH * RB_A_ROWS * 16;
constexpr int RB_SLOT_BYTES = RB_B_ROWS * 16;
constexpr int RB_RING_BYTES = (RB_RING + 1) * RB_SLOT_BYTES;
constexpr int RB_SMEM = RB_RING_BYTES + 2 * RB_A_BYTES;

__global__ void __launch_bounds__(RB_THREADS, 1)
conv5x5_wgrad_ring_kernel(ConvInput in, int N, int D, const float4* __restrict__ dpre, int g_planes,
                          int g_chunk0, int g_chunks, int cout_total, int co0, const uint32_t* __restrict__ xmax, int xmax_n,
                          const uint32_t* __restrict__ gmax, int gmax_n, float* __restrict__ partial, int W, int nstrips) {
  constexpr int NCO = 32;
  // An "image" below is one item: a whole image (nstrips == 1, W == D) or one vertical strip of W output columns (D = 128:
  // two strips of 64 -- the G ring holds 4 rows of pitch W + 8).  The X window of a strip reads its real neighbour columns.
  const int P = W + 8;                                     // multiple of 8
  const int QN = (D + 4) * P;                              // padded-linear input pixels per item
  // Blocks whose X pixels are all halo rows (zeros) do no work at all: the first VB0 blocks of an item (rows 0, 1) only feed
  // G into the ring, which the refill steps of the first live block do anyway, and the blocks behind row D + 1 are dropped.
  // D = 64: 73 live blocks (2 .. 74) of the 77 that cover the padded rows.
  const int VB0 = ring_vb0(P);
  const int NBLK = ring_live_blocks(D, P);                 // live blocks per item; block index vb = VB0 + live index
  // Each CTA owns a contiguous range [g0, g1) of the global block sequence (image-major), balanced to one block; an
  // "item" is one image's share of that range.  Only the first item of a CTA can start inside an image.
  const long long total_blocks = (long long)N * nstrips * NBLK;
  const int g0 = (int)(total_blocks * blockIdx.x / gridDim.x), g1 = (int)(total_blocks * (blockIdx.x + 1) / gridDim.x);
  const int item0 = g0 / NBLK;
  const int n_items = g1 > g0 ? (g1 + NBLK - 1) / NBLK : 0;   // one past the last image this CTA touches
  const int cin_chunks = in.total_chunks;
  const int cin_pad = cin_chunks * 4;
  // warp index broadcast by shuffle: the role branches are then provably warp-uniform (see the issuer below)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;
  const int achunks = P >> 3;                              // G chunk offset per kernel row
  const float sx_scale = __uint_as_float(f16_scale_bits(wg_tensor_max(xmax, xmax_n)));     // one scale per TENSOR: the sum runs
  const float sg_scale = __uint_as_float(f16_scale_bits(wg_tensor_max(gmax, gmax_n)));     // over all images

#ifdef OP3_TC_TRACE
  const unsigned long long t_kernel0 = clock64();
#endif
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* ring = smem_raw;
  uint8_t* a_stage = smem_raw + RB_RING_BYTES;
  __shared__ uint64_t full_bar[2], empty_bar[2], acc_bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int MMA_WARP = RB_LOADERS / 32;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], RB_LOADERS / 2); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  float4 bias_acc = make_float4(0.f, 0.f, 0.f, 0.f);

  // step sequence of this CTA: for item = blockIdx.x, += gridDim.x: vb = b0 - RB_PRE .. b1 - 1  (vb < b0: ring refill only)
  auto item_b0 = [&](int item) { int b = g0 - item * NBLK; return (b > 0 ? b : 0) + VB0; };
  auto item_b1 = [&](int item) { int e = g1 - item * NBLK; return (e < NBLK ? e : NBLK) + VB0; };
  auto advance = [&](int& item, int& vb) {                 // next step; item >= n_items afterwards means "no more steps"
    if (++vb >= item_b1(item)) {
      item += 1;
      vb = item_b0(item) - RB_PRE;
    }
  };

  if (warp < MMA_WARP) {
    // ===================== loaders: group g handles the steps t with t % 2 == g =====================
    const int grp = warp >> 2, wl = warp & 3;
    const int t64 = (wl & 1) * 32 + lane;                  // 64 lane tasks per operand: 8 pixel chunks x 8 channel chunks
    const int j = t64 >> 3, cc = t64 & 7;                  // a quarter warp shares the pixel chunk (see the row order above)
    const uint32_t magic = (uint32_t)((0x100000000ull + (uint32_t)P - 1) / (uint32_t)P);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int item = item0, vb = item_b0(item0) - RB_PRE, t = 0;
    if (grp == 1) { advance(item, vb); t = 1; }
    int nitem = item, nvb = vb;                            // the group's next step (t + 2)
    auto step2 = [&](int& it_, int& vb_) { advance(it_, vb_); if (it_ < n_items) advance(it_, vb_); };
    if (wl < 2) {
      const bool act = cc < cin_chunks;                    // X channel chunk cc, pixel chunk j
      auto load_x = [&](int item_, int vb_, float4* xv) {  // pixels q .. q+8 of the padded-linear input space
#pragma unroll
        for (int e = 0; e < 9; ++e) xv[e] = zero4;
        if (!act || vb_ < item_b0(item_)) return;
        const int n = item_ / nstrips, x0 = (item_ - n * nstrips) * W, q = vb_ * HG_BP + 8 * j;
        const float4* src = tc_src_chunk_ptr(in, cc, n, D);
        const int r = (int)__umulhi((uint32_t)q, magic);
        const int iy = r - 2, ix = q - r * P - 2 + x0;     // q..q+7 share a row (q and P are multiples of 8)
        if (q < QN && (unsigned)iy < (unsigned)D) {
          const float4* rp = src + (size_t)iy * D + ix;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if ((unsigned)(ix + e) < (unsigned)D) xv[e] = __ldg(rp + e);
        }
        int ix8 = ix + 8, iy8 = iy;
        if (ix8 - x0 >= P - 2) { ix8 -= P; ++iy8; }
        if (q + 8 < QN && (unsigned)iy8 < (unsigned)D && (unsigned)ix8 < (unsigned)D) xv[8] = __ldg(src + (size_t)iy8 * D + ix8);
      };
      float4 xv[9], xn[9];
      if (item < n_items) load_x(item, vb, xv);
      int k = 0;
      for (; item < n_items; t += 2, ++k) {
        nitem = item; nvb = vb;
        step2(nitem, nvb);
        if (nitem < n_items) load_x(nitem, nvb, xn);
        // Scale + hi/lo split IN PLACE before the stage is free: the packed (hi, lo) words of a pixel pair overwrite the
        // two fp32 values they came from, so nothing is added to the register budget (168 of 170) and between "stage empty"
        // and "stage full" only the shifted copies and the 16 stores remain.
        const bool fill = act && vb >= item_b0(item);
        uint32_t x8l[4];                                   // lo half of the ninth pixel (its hi half replaces xv[8])
        if (fill) {
#define OP3_HG_X_CONVERT(C, R)                                                                                   \
          {                                                                                                      \
            uint32_t hh, ll;                                                                                     \
            hg_pair(xv[0].C, xv[1].C, sx_scale, hh, ll); xv[0].C = __uint_as_float(hh); xv[1].C = __uint_as_float(ll); \
            hg_pair(xv[2].C, xv[3].C, sx_scale, hh, ll); xv[2].C = __uint_as_float(hh); xv[3].C = __uint_as_float(ll); \
            hg_pair(xv[4].C, xv[5].C, sx_scale, hh, ll); xv[4].C = __uint_as_float(hh); xv[5].C = __uint_as_float(ll); \
            hg_pair(xv[6].C, xv[7].C, sx_scale, hh, ll); xv[6].C = __uint_as_float(hh); xv[7].C = __uint_as_float(ll); \
            hg_pair(xv[8].C, 0.f, sx_scale, hh, ll); xv[8].C = __uint_as_float(hh); x8l[R] = ll;                    \
          }
          OP3_HG_X_CONVERT(x, 0)
          OP3_HG_X_CONVERT(y, 1)
          OP3_HG_X_CONVERT(z, 2)
          OP3_HG_X_CONVERT(w, 3)
#undef OP3_HG_X_CONVERT
        }
        if (k >= 1) mbar_wait_warp(&empty_bar[grp], (k - 1) & 1);
        if (fill) {
          uint4* row = reinterpret_cast<uint4*>(a_stage + (size_t)grp * RB_A_BYTES) + (size_t)j * RB_A_ROWS + cc;
#define OP3_HG_X_STORE(C, R)                                                                                     \
          {                                                                                                      \
            const uint32_t h01 = __float_as_uint(xv[0].C), l01 = __float_as_uint(xv[1].C), h23 = __float_as_uint(xv[2].C),   \
                           l23 = __float_as_uint(xv[3].C), h45 = __float_as_uint(xv[4].C), l45 = __float_as_uint(xv[5].C),   \
                           h67 = __float_as_uint(xv[6].C), l67 = __float_as_uint(xv[7].C), h8 = __float_as_uint(xv[8].C),    \
                           l8 = x8l[R];                                                                          \
            row[8 * R] = make_uint4(h01, h23, h45, h67);                                                         \
            row[32 + 8 * R] = make_uint4(l01, l23, l45, l67);                                                    \
            row[64 + 8 * R] = make_uint4(hg_shift1(h01, h23), hg_shift1(h23, h45), hg_shift1(h45, h67), hg_shift1(h67, h8)); \
            row[96 + 8 * R] = make_uint4(hg_shift1(l01, l23), hg_shift1(l23, l45), hg_shift1(l45, l67), hg_shift1(l67, l8)); \
          }
          OP3_HG_X_STORE(x, 0)
          OP3_HG_X_STORE(y, 1)
          OP3_HG_X_STORE(z, 2)
          OP3_HG_X_STORE(w, 3)
#undef OP3_HG_X_STORE
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[grp]);
#pragma unroll
        for (int e = 0; e < 9; ++e) xv[e] = xn[e];
        item = nitem; vb = nvb;
      }
    } else {
      const bool act = cc < g_chunks;                      // G channel chunk cc, pixel chunk j
      // pixels p0-4 .. p0+7 of the output-gradient space: G chunk 8*vb + j of the image and the 4 pixels before it
      auto load_g = [&](int item_, int vb_, float4* gv) {
#pragma unroll
        for (int e = 0; e < 12; ++e) gv[e] = zero4;
        if (!act) return;
        const int n = item_ / nstrips, x0 = (item_ - n * nstrips) * W;
        const float4* src = dpre + ((size_t)n * g_planes + g_chunk0 + cc) * D * D;
        const int p0 = vb_ * HG_BP + 8 * j;
        if (p0 >= 0) {
          const int y = (int)__umulhi((uint32_t)p0, magic), x = p0 - y * P;
          if (y < D && x < W) {                            // x and W are multiples of 8: all 8 pixels are valid or none
            const float4* rp = src + (size_t)y * D + x + x0;
#pragma unroll
            for (int e = 0; e < 8; ++e) gv[4 + e] = __ldg(rp + e);
          }
        }
        const int pm = p0 - 8;                             // the chunk before: only its last 4 pixels are needed
        if (pm >= 0) {
          const int y = (int)__umulhi((uint32_t)pm, magic), x = pm - y * P;
          if (y < D && x < W) {
            const float4* rp = src + (size_t)y * D + x + x0;
#pragma unroll
            for (int e = 4; e < 8; ++e) gv[e - 4] = __ldg(rp + e);
          }
        }
      };
      float4 gv[12], gn[12];
      if (item < n_items) load_g(item, vb, gv);
      int k = 0;
      for (; item < n_items; t += 2, ++k) {
        nitem = item; nvb = vb;
        step2(nitem, nvb);
        if (nitem < n_items) load_g(nitem, nvb, gn);
        // bias sums, then scale + hi/lo split IN PLACE before the ring slots are free: the packed (hi, lo) words of a pixel
        // pair overwrite the two fp32 values they came from (no extra registers); between "slots empty" and "slots full"
        // only the three shifted row sets are assembled and stored
#ifdef OP3_TC_TRACE
        const unsigned long long tc0 = clock64();
#endif
        if (act) {
          // refill chunks belong to the previous range's sum -- except the G chunks of the leading halo blocks, which only
          // ever pass through the refill of the range that starts the item
          const int b0i = item_b0(item);
          if (vb >= b0i || (b0i == VB0 && vb >= 0)) {
#pragma unroll
            for (int e = 4; e < 12; ++e) {
              bias_acc.x += gv[e].x; bias_acc.y += gv[e].y; bias_acc.z += gv[e].z; bias_acc.w += gv[e].w;
            }
          }
#define OP3_HG_G_CONVERT(C)                                                                                      \
          _Pragma("unroll") for (int e = 0; e < 12; e += 2) {                                                    \
            uint32_t hh, ll;                                                                                     \
            hg_pair(gv[e].C, gv[e + 1].C, sg_scale, hh, ll);                                                     \
            gv[e].C = __uint_as_float(hh); gv[e + 1].C = __uint_as_float(ll);                                    \
          }
          OP3_HG_G_CONVERT(x)
          OP3_HG_G_CONVERT(y)
          OP3_HG_G_CONVERT(z)
          OP3_HG_G_CONVERT(w)
#undef OP3_HG_G_CONVERT
        }
#ifdef OP3_TC_TRACE
        const unsigned long long tw0 = clock64();
#endif
        if (k >= 1) mbar_wait_warp(&empty_bar[grp], (k - 1) & 1);
#ifdef OP3_TC_TRACE
        const unsigned long long tw1 = clock64();
#endif
        if (act) {
          const int slot = (8 * t + j) % RB_RING;          // running slot counter: 8 new chunks per step
          uint4* row = reinterpret_cast<uint4*>(ring + (size_t)slot * RB_SLOT_BYTES) + cc;
          uint4* guard = reinterpret_cast<uint4*>(ring + (size_t)RB_RING * RB_SLOT_BYTES) + cc;
          const bool dup = slot == 0;                      // slot 0 is mirrored behind the ring for K pairs that wrap
#define OP3_HG_G_STORE(C, R)                                                                                     \
          {                                                                                                      \
            const uint32_t hm2 = __float_as_uint(gv[0].C), lm2 = __float_as_uint(gv[1].C), hm1 = __float_as_uint(gv[2].C),   \
                           lm1 = __float_as_uint(gv[3].C), h0 = __float_as_uint(gv[4].C), l0 = __float_as_uint(gv[5].C),     \
                           h1 = __float_as_uint(gv[6].C), l1 = __float_as_uint(gv[7].C), h2 = __float_as_uint(gv[8].C),      \
                           l2 = __float_as_uint(gv[9].C), h3 = __float_as_uint(gv[10].C), l3 = __float_as_uint(gv[11].C);    \
            const uint4 g0h = make_uint4(h0, h1, h2, h3), g0l = make_uint4(l0, l1, l2, l3);       /* p0   .. p0+7 */ \
            const uint4 g2h = make_uint4(hm1, h0, h1, h2), g2l = make_uint4(lm1, l0, l1, l2);     /* p0-2 .. p0+5 */ \
            const uint4 g4h = make_uint4(hm2, hm1, h0, h1), g4l = make_uint4(lm2, lm1, l0, l1);   /* p0-4 .. p0+3 */ \
            row[8 * R] = g0h; row[32 + 8 * R] = g2h; row[64 + 8 * R] = g4h;                                      \
            row[96 + 8 * R] = g0l; row[128 + 8 * R] = g2l; row[160 + 8 * R] = g4l;                               \
            if (dup) {                                                                                           \
              guard[8 * R] = g0h; guard[32 + 8 * R] = g2h; guard[64 + 8 * R] = g4h;                              \
              guard[96 + 8 * R] = g0l; guard[128 + 8 * R] = g2l; guard[160 + 8 * R] = g4l;                       \
            }                                                                                                    \
          }
          OP3_HG_G_STORE(x, 0)
          OP3_HG_G_STORE(y, 1)
          OP3_HG_G_STORE(z, 2)
          OP3_HG_G_STORE(w, 3)
#undef OP3_HG_G_STORE
        }
#ifdef OP3_TC_TRACE
        const unsigned long long tw2 = clock64();
#endif
        fence_proxy_async();
#ifdef OP3_TC_TRACE
        const unsigned long long tw3 = clock64();
#endif
        mbar_arrive(&full_bar[grp]);
#ifdef OP3_TC_TRACE
        if (tid == 64) {
          atomicAdd(&g_ring_trace[blockIdx.x][6], tw1 - tw0); atomicAdd(&g_ring_trace[blockIdx.x][7], tw2 - tw1);
          atomicAdd(&g_ring_trace[blockIdx.x][4], tw3 - tw2); atomicAdd(&g_ring_trace[blockIdx.x][5], tw0 - tc0);
        }
#endif
#pragma unroll
        for (int e = 0; e < 12; ++e) gv[e] = gn[e];
        item = nitem; vb = nvb;
      }
    }
  } else {
    // ===================== MMA issuer =====================
    // The WHOLE warp runs this loop in uniform control flow; only the tcgen05 instructions are predicated on the lane chosen
    // by elect.sync.  ptxas then keeps every descriptor word (loop counters + constants) in UNIFORM registers: UTCHMMA
    // preceded by one UIADD3.  Issued from inside an `if (lane == 0)` region, each MMA was wrapped in an ELECT /
    // R2UR.BROADCAST waterfall of ~12 instructions -- ~25 cycles per MMA, the whole gap between the isolated (49-56
    // cycles) and the in-situ (75) rate of these N=96 MMAs, independent of loaders, operand source or bookkeeping.
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = make_idesc_f16(128, 96);
    const uint64_t bd0 = make_smem_desc(smem_u32(ring), RB_B_ROWS, 8);
    const uint32_t b_lo0 = (uint32_t)bd0, b_hi32 = (uint32_t)(bd0 >> 32);
    const uint64_t ad0 = make_smem_desc(smem_u32(a_stage), RB_A_ROWS, 8);
    const uint32_t a_lo0 = (uint32_t)ad0, a_hi32 = (uint32_t)(ad0 >> 32);
    int item = item0, t = 0, c0 = 0;                       // c0 = (8 t) mod RING: slot of this block's first G chunk
    int cur_b0 = item_b0(item0), cur_b1 = item_b1(item0), vb = cur_b0 - RB_PRE;
    bool first = true;
    int koff[5];                                           // (-achunks * ky) mod RING
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) koff[ky] = (5 * RB_RING - achunks * ky) % RB_RING;
#ifdef OP3_TC_TRACE
    unsigned long long t_full = 0, n_mma = 0;
    const unsigned long long t_begin = clock64();
#endif
    for (; item < n_items; ++t) {
      const int s = t & 1;
      const bool real = vb >= cur_b0;
      const uint32_t a_lo32 = a_lo0 + (uint32_t)(s * (RB_A_BYTES / 16));
      uint32_t boff[5][HG_CH / 2];                         // wrapped ring slots of the 5 x 4 K-pair operands, 16-byte units
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        int slot = c0 + koff[ky];
        if (slot >= RB_RING) slot -= RB_RING;
#pragma unroll
        for (int ks = 0; ks < HG_CH / 2; ++ks) {
          boff[ky][ks] = b_lo0 + (uint32_t)(slot * RB_B_ROWS);
          slot += 2;
          if (slot >= RB_RING) slot -= RB_RING;
        }
      }
#ifdef OP3_TC_TRACE
      const unsigned long long tf0 = clock64();
#endif
      mbar_wait(&full_bar[s], (t >> 1) & 1);
#ifdef OP3_TC_TRACE
      t_full += clock64() - tf0;
      if (real) n_mma += 40;
#endif
      tc_fence_after();
      if (real) {
        // The X block goes to TENSOR MEMORY first (4 K-steps x 8 columns behind the 480 accumulator columns): its ten MMAs
        // per K-step then fetch only the 3 KB of G from shared memory instead of 4 + 3 KB.  tcgen05.cp and tcgen05.mma
        // execute in issue order, so the next block's copy cannot overtake this block's MMAs.
#ifdef OP3_RING_TMEM_A
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < HG_CH / 2; ++ks)
            tc_cp_a128(tmem_u + (uint32_t)(480 + 8 * ks), a_lo32 + (uint32_t)(2 * ks * RB_A_ROWS), a_hi32);
        }
#endif
        // consecutive MMAs go to DIFFERENT accumulators (kernel rows): back-to-back MMAs on one accumulator serialise
#pragma unroll
        for (int ks = 0; ks < HG_CH / 2; ++ks) {
#ifdef OP3_RING_TMEM_A
          const uint32_t a_tm = tmem_u + (uint32_t)(480 + 8 * ks);
#else
          const uint32_t aoff = a_lo32 + (uint32_t)(2 * ks * RB_A_ROWS);
#endif
#pragma unroll
          for (int hl = 0; hl < 2; ++hl)
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
              const uint32_t dcol = tmem_u + (uint32_t)(96 * ky);
              if (leader) {
#ifdef OP3_RING_TMEM_A
                if (ks == 0 && hl == 0) mma_f16_ts(dcol, a_tm, boff[ky][ks], b_hi32, idesc, first ? 0u : 1u);
                else mma_f16_ts_acc(dcol, a_tm, boff[ky][ks] + (hl ? 96u : 0u), b_hi32, idesc);
#else
                if (ks == 0 && hl == 0) mma_f16(dcol, aoff, a_hi32, boff[ky][ks], b_hi32, idesc, first ? 0u : 1u);
                else mma_f16_acc(dcol, aoff, a_hi32, boff[ky][ks] + (hl ? 96u : 0u), b_hi32, idesc);
#endif
              }
            }
        }
        first = false;
      }
      if (leader) tc_commit(&empty_bar[s]);
      __syncwarp();
      if (++vb >= cur_b1) {                                // next image of this CTA's range
        item += 1;
        cur_b0 = item_b0(item); cur_b1 = item_b1(item);
        vb = cur_b0 - RB_PRE;
      }
      c0 += 8;
      if (c0 >= RB_RING) c0 -= RB_RING;
    }
    if (leader) tc_commit(&acc_bar);
#ifdef OP3_TC_TRACE
    if (leader) {
      atomicAdd(&g_ring_trace[blockIdx.x][0], t_full); atomicAdd(&g_ring_trace[blockIdx.x][1], clock64() - t_begin);
      atomicAdd(&g_ring_trace[blockIdx.x][2], n_mma); atomicAdd(&g_ring_trace[blockIdx.x][3], (unsigned long long)t);
    }
#endif
  }
  // ===================== epilogue =====================
  __syncthreads();
  // TMEM lanes: quadrant 0 / 2 = X shift 0 / 1 hi rows, quadrant 1 / 3 = the matching lo rows; columns 96*ky + 32*c + pc.
  // (1) the lo quadrants go to shared memory (row pitch 481 floats: conflict-free), (2) the hi quadrants add them, scale
  // and leave the sums there, (3) all warps write the slab as coalesced 128-byte rows of 32 output channels.
  constexpr int RPITCH = 481;
  float* red = reinterpret_cast<float*>(smem_raw);         // [2 shifts][32 rows][RPITCH]
  float* s_bias = red + 2 * 32 * RPITCH;                   // [4 G warps][8 chunks][4]
  float* slab = partial + (size_t)blockIdx.x * ((size_t)25 * cin_pad * cout_total + cout_total);
  const float inv = 1.f / (sx_scale * sg_scale);           // powers of two: exact
  if (warp < MMA_WARP) {
    mbar_wait_warp(&acc_bar, 0);
    tc_fence_after();
  }
  if (warp < MMA_WARP && (warp & 3) >= 2) {                // G warps: 8 lanes (pixel chunks) share a channel chunk
#pragma unroll
    for (int o = 16; o >= 8; o >>= 1) {
      bias_acc.x += __shfl_xor_sync(0xffffffffu, bias_acc.x, o); bias_acc.y += __shfl_xor_sync(0xffffffffu, bias_acc.y, o);
      bias_acc.z += __shfl_xor_sync(0xffffffffu, bias_acc.z, o); bias_acc.w += __shfl_xor_sync(0xffffffffu, bias_acc.w, o);
    }
    if (lane < 8) {                                        // lane = channel chunk; one partial per G warp
      float* dst = s_bias + (((warp >> 2) * 2 + (warp & 1)) * 8 + lane) * 4;
      dst[0] = bias_acc.x; dst[1] = bias_acc.y; dst[2] = bias_acc.z; dst[3] = bias_acc.w;
    }
  }
  const int quad = warp & 3, chalf = warp >> 2;            // two warps per TMEM lane quadrant split the 480 columns
  if (warp < MMA_WARP && (quad & 1)) {
    float* dst = red + ((size_t)(quad >> 1) * 32 + lane) * RPITCH;
    for (int c0 = 240 * chalf; c0 < 240 * chalf + 240; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) dst[c0 + i] = v[i];
    }
  }
  __syncthreads();
  if (warp < MMA_WARP && !(quad & 1)) {
    float* dst = red + ((size_t)(quad >> 1) * 32 + lane) * RPITCH;
    for (int c0 = 240 * chalf; c0 < 240 * chalf + 240; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) dst[c0 + i] = (v[i] + dst[c0 + i]) * inv;
    }
  }
  __syncthreads();
  {
    const int co = lane, pc = 8 * (co & 3) + (co >> 2);    // column 8*r + cc holds output channel 4*cc + r
    if (co < g_chunks * 4)
      for (int r = warp; r < 25 * cin_pad; r += RB_THREADS / 32) {
        const int tap = r / cin_pad, ci = r - tap * cin_pad;
        const int ky = tap / 5, kx = tap - 5 * ky;
        const int row = 8 * (ci & 3) + (ci >> 2);          // row 8*r + cc holds input channel 4*cc + r
        slab[((size_t)tap * cin_pad + ci) * cout_total + co0 + co] =
            red[((size_t)(kx & 1) * 32 + row) * RPITCH + 96 * ky + 32 * (kx >> 1) + pc];
      }
  }
  if (tid < g_chunks * 4)
    slab[(size_t)25 * cin_pad * cout_total + co0 + tid] = (s_bias[tid] + s_bias[32 + tid]) + (s_bias[64 + tid] + s_bias[96 + tid]);
  tc_fence_before();
  __syncthreads();
#ifdef OP3_TC_TRACE
#endif
  if (warp == MMA_WARP) tmem_dealloc<512>(tmem_base);
}

int wg_sets() { return 29; }       // 5 kernel rows x 29 sets = 145 persistent CTAs on 148 SMs

template <int NCO, bool SPLIT3>
int launch_wgrad_tc(cudaStream_t st, int N, int D, const ConvInput& in, const float4* dpre, int g_planes, int g_chunk0,
                    int g_chunks, int cout_total, int co0, float* scratch) {
  using Cfg = WgCfg<NCO, SPLIT3>;
  size_t smem = 2 * (size_t)Cfg::STAGE_BYTES;
  size_t red = (size_t)(3 * 2 * 32 * NCO + 64) * sizeof(float);
  if (red > smem) smem = red;
  OP3_REQUIRE(smem <= 226 * 1024, "conv5x5_wgrad_tc: %zu bytes of shared memory", smem);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_wgrad_tc_kernel<NCO, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  }
  dim3 grid(wg_sets(), 5);
  conv5x5_wgrad_tc_kernel<NCO, SPLIT3><<<grid, WG_THREADS, smem, st>>>(in, N, D, dpre, g_planes, g_chunk0, g_chunks, cout_total,
                                                                      co0, scratch);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

int launch_wgrad_f16(cudaStream_t st, int N, int D, const ConvInput& in, const float4* dpre, int g_planes, int g_chunk0,
                     int g_chunks, int cout_total, int co0, const uint32_t* xmax, int xmax_n, const uint32_t* gmax,
                     int gmax_n, float* scratch) {
  size_t smem = HG_GROUPS * (size_t)HG_STAGE_BYTES;
  OP3_REQUIRE(smem <= 226 * 1024, "conv5x5_wgrad_f16: %zu bytes of shared memory", smem);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_wgrad_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  }
  dim3 grid(wg_sets(), 5);
  conv5x5_wgrad_f16_kernel<<<grid, HG_THREADS, smem, st>>>(in, N, D, dpre, g_planes, g_chunk0, g_chunks, cout_total, co0, xmax,
                                                          xmax_n, gmax, gmax_n, scratch);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

// returns the number of per-CTA partial slabs written (the grid size)
// Strip decomposition for the ring kernel: the refill of RB_PRE blocks must cover the 4-row halo of G chunks,
// 4 (W + 8) / 8 <= 8 RB_PRE, i.e. W <= 72: whole images up to D = 72, else the fewest equal strips (W a multiple of 8).
bool ring_strips(int D, int* W, int* nstrips) {
  if (D % 8) return false;
  for (int ns = 1; ns <= 16; ++ns) {
    if (D % ns) continue;
    const int w = D / ns;
    if (w % 8 == 0 && 4 * ((w + 8) / 8) <= 8 * RB_PRE) { *W = w; *nstrips = ns; return true; }
  }
  return false;
}

int launch_wgrad_ring(cudaStream_t st, int N, int D, const ConvInput& in, const float4* dpre, int g_planes, int g_chunk0,
                      int g_chunks, int cout_total, int co0, const uint32_t* xmax, int xmax_n, const uint32_t* gmax,
                      int gmax_n, float* scratch, int* grid_out) {
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_wgrad_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  }
  const int g_wg_sms = device_sm_count() > 148 ? 148 : device_sm_count();   // scratch is sized for 148 slabs
  int W = 0, nstrips = 0;
  OP3_REQUIRE(ring_strips(D, &W, &nstrips), "conv5x5_wgrad_ring: D=%d has no strip decomposition", D);
  const int P = W + 8, NBLK = ring_live_blocks(D, P);
  const long long total_blocks = (long long)N * nstrips * NBLK;
  const int grid = total_blocks < g_wg_sms ? (int)total_blocks : g_wg_sms;
  conv5x5_wgrad_ring_kernel<<<grid, RB_THREADS, RB_SMEM, st>>>(in, N, D, dpre, g_planes, g_chunk0, g_chunks, cout_total,
                                                              co0, xmax, xmax_n, gmax, gmax_n, scratch, W, nstrips);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  *grid_out = grid;
  return OP3_OK;
}

}  // namespace

#ifdef OP3_TC_TRACE
extern "C" int op3_debug_ring_trace(unsigned long long* out /*[256][4] host*/, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_ring_trace, sizeof(unsigned long long) * 256 * 8);
  if (reset) { static unsigned long long z[256][8]; cudaMemcpyToSymbol(g_ring_trace, z, sizeof(z)); }
  return 0;
}
#endif

size_t conv5x5_wgrad_tc_scratch_floats(int cin_pad, int cout) {
  return (size_t)wg_sets() * ((size_t)25 * cin_pad * cout + cout);
}

// dwt [25][cin_pad][cout] and dbias [cout] are ACCUMULATED (same contract as conv5x5_wgrad).
int conv5x5_wgrad_tc(cudaStream_t st, int precision, int N, int D, const ConvInput& in, int cout, const float4* dpre,
                     float* scratch, float* dwt, float* dbias, const uint32_t* dpre_maxbits) {
  if (N <= 0) return OP3_OK;
  OP3_REQUIRE(cout % 4 == 0 && cout <= 64 && in.total_chunks <= 8 && D % 4 == 0,
              "conv5x5_wgrad_tc: unsupported cin chunks %d / cout %d / D %d", in.total_chunks, cout, D);
  const int cin_real = in.real_cin ? in.real_cin : in.total_chunks * 4;
  ProfScope prof(st, KID_CONV_WGRAD, 2.0 * 25 * cin_real * cout * (double)N * D * D);
  const bool s3 = precision != OP3_PREC_TF32;       // the fp16 forward mode keeps 3xTF32 weight gradients
  const int planes = cout / 4;
  const int cin_pad = in.total_chunks * 4;
  const size_t slab = (size_t)25 * cin_pad * cout + cout;
  int n_slabs = wg_sets();
  const bool f16 = precision == OP3_PREC_F16X3 && (cout >= 32 || cout == 4) && D % 8 == 0;
  int ring_w = 0, ring_ns = 0;
  const bool use_ring = f16 && ring_strips(D, &ring_w, &ring_ns);
  const size_t zero_slabs = use_ring ? 148 : (size_t)wg_sets();
  // The per-row kernels leave padding entries of their slabs untouched (they must read as zero); the ring kernel writes
  // every entry of every slab it owns, so only the 8 operand-maxima words behind the slabs are cleared for it.
  if (use_ring) OP3_CHECK(cudaMemsetAsync(scratch + zero_slabs * slab, 0, 8 * sizeof(float), st));
  else OP3_CHECK(cudaMemsetAsync(scratch, 0, (zero_slabs * slab + 8) * sizeof(float), st));
  if (f16) {
    // per-tensor operand maxima live in the 8 words after the partial slabs (scratch: conv5x5_wgrad_scratch_floats)
    // operand maxima: per-image arrays recorded by the producers (n = N entries), else one word reduced here (n = 1)
    uint32_t* mx = reinterpret_cast<uint32_t*>(scratch + zero_slabs * slab);
    const uint32_t *xmax = in.maxbits, *gmax = dpre_maxbits;
    int xmax_n = N, gmax_n = N;
    if (xmax == nullptr) {
      wg_absmax_kernel<<<dim3(N, in.total_chunks), 256, 0, st>>>(in, D, mx);
      OP3_COUNT_LAUNCH();
      xmax = mx; xmax_n = 1;
    }
    if (gmax == nullptr) {
      ConvInput gin{};
      gin.src[0].ptr = dpre; gin.src[0].nchunks = planes; gin.src[0].div = 1;
      gin.nsrc = 1; gin.total_chunks = planes;
      wg_absmax_kernel<<<dim3(N, planes), 256, 0, st>>>(gin, D, mx + 4);
      OP3_COUNT_LAUNCH();
      gmax = mx + 4; gmax_n = 1;
    }
    OP3_LAUNCH_CHECK();
    if (use_ring) {                                      // ring kernel: its refill covers the 4-row halo of G chunks
      int grid = 0;
      for (int h = 0; h * 32 < cout; ++h)
        OP3_TRY(launch_wgrad_ring(st, N, D, in, dpre, planes, 8 * h, planes - 8 * h < 8 ? planes - 8 * h : 8, cout, 32 * h, xmax, xmax_n,
                                  gmax, gmax_n, scratch, &grid));
      n_slabs = grid;
    } else {
      for (int h = 0; h * 32 < cout; ++h)
        OP3_TRY(launch_wgrad_f16(st, N, D, in, dpre, planes, 8 * h, planes - 8 * h < 8 ? planes - 8 * h : 8, cout, 32 * h, xmax, xmax_n,
                                 gmax, gmax_n, scratch));
    }
  } else if (cout <= 16) {
    OP3_TRY((s3 ? launch_wgrad_tc<16, true>(st, N, D, in, dpre, planes, 0, planes, cout, 0, scratch)
                : launch_wgrad_tc<16, false>(st, N, D, in, dpre, planes, 0, planes, cout, 0, scratch)));
  } else {
    for (int h = 0; h * 32 < cout; ++h) {
      const int chunks = (cout - h * 32 >= 32) ? 8 : (cout - h * 32) / 4;
      OP3_TRY((s3 ? launch_wgrad_tc<32, true>(st, N, D, in, dpre, planes, 8 * h, chunks, cout, 32 * h, scratch)
                  : launch_wgrad_tc<32, false>(st, N, D, in, dpre, planes, 8 * h, chunks, cout, 32 * h, scratch)));
    }
  }
  const int n_w = 25 * cin_pad * cout;
  wgrad_tc_reduce_kernel<<<ceil_div(n_w + cout, 64), 256, 0, st>>>(scratch, n_slabs, n_w, cout, dwt, dbias);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

int conv5x5_wgrad_auto(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float4* dpre, float* scratch,
                       float* dwt, float* dbias, const uint32_t* dpre_maxbits) {
  const int prec = op3_get_precision();
  // the 4-channel layer stays on the FMA pipe in the TF32 modes (see conv5x5_auto); the fp16 ring kernel beats it (0.31 vs
  // 0.44 ms) because its cost does not depend on the number of output channels
  if (prec == OP3_PREC_FP32 || (cout == 4 && prec != OP3_PREC_F16X3)) return conv5x5_wgrad(st, N, D, in, cout, dpre, scratch, dwt, dbias);
  return conv5x5_wgrad_tc(st, prec, N, D, in, cout, dpre, scratch, dwt, dbias, dpre_maxbits);
}

}  // namespace op3

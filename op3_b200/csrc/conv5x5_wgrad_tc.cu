// tcgen05 weight gradient of the 5x5 convolution:
//     dW[tap][ci][co] = sum_{n,p} X[n][p + o_tap][ci] * G[n][p][co]        (p, o in padded-linear pixel space, pitch D+4)
// The contraction runs over PIXELS, so both operands must be K-major with pixels along K.  Measured on B200
// (tests/probes/probe_mnmajor.cu, probe_mn2.cu): kind::tf32 with MN-major (transposed) operands silently produces zeros
// for every swizzle mode, so the loaders transpose 4 pixels x 4 channels blocks in registers into
//     [pixel chunk of 4][channel row][4 pixels]            (16-byte rows, K-major, no swizzle).
// A 16-byte K chunk holds 4 pixels, so a tap offset o = 4a + phi can only be applied in whole chunks; the residual
// phase phi is absorbed by keeping FOUR copies of the G block, copy phi shifted by phi pixels.
//   * One CTA owns one kernel ROW ky (5 taps): its X window needs no halo beyond 4 pixels and 5 accumulators
//     (128 lanes x N columns each) fit TMEM.  grid = 5 x S persistent CTAs over (image, 64-pixel block) items.
//   * 3xTF32: A rows = [X_hi (32); X_lo (32); ones], B rows = [G_hi; G_lo]: ONE MMA yields all hi/lo cross products,
//     the quadrants are summed in the epilogue; the `ones` row gives the bias gradient for free.
//   * warps 0-7 transpose/round/split operands into smem (2 stages), warp 8 issues the MMAs, warps 0-3 run the epilogue.
#include "common.cuh"
#include "tc_common.cuh"

namespace op3 {

namespace {

constexpr int WG_BP = 64;                    // pixels per block (8 K-steps of 8)
constexpr int WG_CH = WG_BP / 4;             // 16 pixel chunks per block
constexpr int WG_GCH = WG_CH + 1;            // G copies need one extra chunk (tap offsets a and a+1)
constexpr int WG_LOADERS = 256;
constexpr int WG_THREADS = 288;
constexpr int WG_A_ROWS = 73;                // rows per A chunk (>= 65 used; == 1 mod 8 -> conflict-free 16-byte stores)

template <int NCO, bool SPLIT3>
struct WgCfg {
  static constexpr int NB = SPLIT3 ? 2 * NCO : NCO;          // B rows = accumulator columns per tap
  static constexpr int B_ROWS = NB + 1;                      // == 1 mod 8
  static constexpr int ONES_ROW = SPLIT3 ? 64 : 32;
  static constexpr int A_BYTES = WG_CH * WG_A_ROWS * 16 + 2048;          // + slack: M=128 reads run past the last chunk
  static constexpr int G_BYTES = WG_GCH * B_ROWS * 16;
  static constexpr int STAGE_BYTES = A_BYTES + 4 * G_BYTES;
  static constexpr int TMEM_COLS = 5 * NB <= 128 ? 128 : (5 * NB <= 256 ? 256 : 512);
};

__device__ __forceinline__ void split_store(float4* hi_dst, float4* lo_dst, float a, float b, float c, float d, bool split) {
  float4 h = make_float4(tf32_rn(a), tf32_rn(b), tf32_rn(c), tf32_rn(d));
  *hi_dst = h;
  if (split) *lo_dst = make_float4(tf32_rn(a - h.x), tf32_rn(b - h.y), tf32_rn(c - h.z), tf32_rn(d - h.w));
}

template <int NCO, bool SPLIT3>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv5x5_wgrad_tc_kernel(ConvInput in, int N, int D, const float4* __restrict__ dpre, int g_planes, int g_chunk0,
                        int g_chunks, int cout_total, int co0, float* __restrict__ partial) {
  using Cfg = WgCfg<NCO, SPLIT3>;
  constexpr int NB = Cfg::NB, B_ROWS = Cfg::B_ROWS;
  const int P = D + 4;
  const int QN = (D + 4) * P;                              // padded-linear input pixels per image
  const int NBLK = (QN + WG_BP - 1) / WG_BP;
  const int ky = blockIdx.y;
  const int n_items = N * NBLK;
  const int cin_chunks = in.total_chunks;
  const int cin_pad = cin_chunks * 4;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  // tap offsets o_k = ky*P + kx = 4*a_k + phi_k
  const int o0 = ky * P;
  const int a_max = (o0 + 4) >> 2;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[2], empty_bar[2], acc_bar;
  __shared__ uint32_t tmem_base_s;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], WG_LOADERS); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_s);
  // constant rows: the `ones` row of every A chunk (bias gradient), written once for both stages
  for (int e = tid; e < 2 * WG_CH; e += WG_THREADS) {
    int s = e / WG_CH, j = e % WG_CH;
    float4* a = reinterpret_cast<float4*>(smem_raw + (size_t)s * Cfg::STAGE_BYTES) + (size_t)j * WG_A_ROWS + Cfg::ONES_ROW;
    *a = make_float4(1.f, 1.f, 1.f, 1.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 8) {
    // ===================== loaders: transpose 4 pixels x 4 channels, round to tf32, split hi/lo =====================
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int n = item / NBLK, blk = item % NBLK;
      const int Q0 = blk * WG_BP;
      const int s = it & 1;
      if (it >= 2) mbar_wait(&empty_bar[s], ((it >> 1) - 1) & 1);
      uint8_t* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
      float4* A = reinterpret_cast<float4*>(st);
      // Warp w owns channel chunk w of X (lanes 0..15 = pixel chunks) and of G (lanes 0..17 = raw aligned chunks
      // rj = lane-1 in [-1,16]).  All global loads are issued first; the four phase copies of G are then assembled from
      // the lane's own pixels and its left neighbour's (warp shuffles) instead of four separate loads.
      float4 xv[4], gv[4];
      const bool do_x = warp < cin_chunks && lane < WG_CH;
      const bool do_g = warp < g_chunks && lane < WG_GCH + 1;
      if (do_x) {
        const float4* src = tc_src_chunk_ptr(in, warp, n, D);
        int q = Q0 + 4 * lane;
        int iy = q / P - 2, ix = q - (iy + 2) * P - 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          bool ok = q + e < QN && iy >= 0 && iy < D && ix >= 0 && ix < D;
          xv[e] = ok ? __ldg(src + (size_t)iy * D + ix) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (++ix >= P - 2) { ix -= P; ++iy; }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) gv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (do_g) {
        const float4* src = dpre + ((size_t)n * g_planes + g_chunk0 + warp) * D * D;
        const int p = Q0 - 4 * a_max + 4 * (lane - 1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int pe = p + e;
          const int y = pe >= 0 ? pe / P : 0;
          const int x = pe - y * P;
          if (pe >= 0 && y < D && x < D) gv[e] = __ldg(src + (size_t)y * D + x);
        }
      }
      if (do_x) {
        float4* row = A + (size_t)lane * WG_A_ROWS + 4 * warp;
        split_store(row + 0, row + 32 + 0, xv[0].x, xv[1].x, xv[2].x, xv[3].x, SPLIT3);
        split_store(row + 1, row + 32 + 1, xv[0].y, xv[1].y, xv[2].y, xv[3].y, SPLIT3);
        split_store(row + 2, row + 32 + 2, xv[0].z, xv[1].z, xv[2].z, xv[3].z, SPLIT3);
        split_store(row + 3, row + 32 + 3, xv[0].w, xv[1].w, xv[2].w, xv[3].w, SPLIT3);
      }
      if (warp < g_chunks) {                      // whole warp: shuffles must be convergent
        // left neighbour's pixels 1..3 (raw chunk rj-1), needed for phases 1..3
        float4 lv[4];
#pragma unroll
        for (int e = 1; e < 4; ++e) {
          lv[e].x = __shfl_up_sync(0xffffffffu, gv[e].x, 1); lv[e].y = __shfl_up_sync(0xffffffffu, gv[e].y, 1);
          lv[e].z = __shfl_up_sync(0xffffffffu, gv[e].z, 1); lv[e].w = __shfl_up_sync(0xffffffffu, gv[e].w, 1);
        }
        if (lane >= 1 && lane < WG_GCH + 1) {     // local chunk lj = lane-1 in [0,17)
          const int lj = lane - 1;
#pragma unroll
          for (int phi = 0; phi < 4; ++phi) {
            // copy phi, element e <- raw pixel (e - phi) of this chunk, or (e - phi + 4) of the left neighbour
            float4 w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = e >= phi ? gv[e - phi] : lv[e - phi + 4];
            float4* G = reinterpret_cast<float4*>(st + Cfg::A_BYTES + (size_t)phi * Cfg::G_BYTES);
            float4* row = G + (size_t)lj * B_ROWS + 4 * warp;
            split_store(row + 0, row + NCO + 0, w[0].x, w[1].x, w[2].x, w[3].x, SPLIT3);
            split_store(row + 1, row + NCO + 1, w[0].y, w[1].y, w[2].y, w[3].y, SPLIT3);
            split_store(row + 2, row + NCO + 2, w[0].z, w[1].z, w[2].z, w[3].z, SPLIT3);
            split_store(row + 3, row + NCO + 3, w[0].w, w[1].w, w[2].w, w[3].w, SPLIT3);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&full_bar[s]);
    }
  } else if (lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_tf32(128, NB);
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait(&full_bar[s], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem_raw + (size_t)s * Cfg::STAGE_BYTES);
      const uint64_t ad = make_smem_desc(st, WG_A_ROWS, 8);
      const uint32_t a_lo32 = (uint32_t)ad, a_hi32 = (uint32_t)(ad >> 32);
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int o = o0 + kx;
        const int phi = o & 3, da = a_max - (o >> 2);          // local G chunk = X chunk + da
        const uint64_t bd = make_smem_desc(st + Cfg::A_BYTES + phi * Cfg::G_BYTES + da * B_ROWS * 16, B_ROWS, 8);
        const uint32_t b_lo32 = (uint32_t)bd, b_hi32 = (uint32_t)(bd >> 32);
        const uint32_t dcol = tmem_base + (uint32_t)(kx * NB);
#pragma unroll
        for (int ks = 0; ks < WG_BP / 8; ++ks)
          mma_tf32(dcol, a_lo32 + (uint32_t)(2 * ks * WG_A_ROWS), a_hi32, b_lo32 + (uint32_t)(2 * ks * B_ROWS), b_hi32, idesc,
                   (it > 0 || ks > 0) ? 1u : 0u);
      }
      tc_commit(&empty_bar[s]);
    }
    tc_commit(&acc_bar);
  }
  // ===================== epilogue: quadrant sums -> per-CTA partial slab =====================
  __syncthreads();                      // loaders are done with smem; reuse it as the combine buffer
  float* red = reinterpret_cast<float*>(smem_raw);         // [5 taps][32 ci][NCO] contributions of the X_lo rows
  float* slab = partial + (size_t)blockIdx.x * ((size_t)25 * cin_pad * cout_total + cout_total);
  if (warp < 4) {
    if (n_items > (int)blockIdx.x) {
      mbar_wait(&acc_bar, 0);
      tc_fence_after();
    }
    const bool have = n_items > (int)blockIdx.x;
    // pass 1: warp 1 (lanes 32..63 = X_lo rows) parks its quadrant sums in smem
    if (SPLIT3 && warp == 1) {
      for (int kx = 0; kx < 5; ++kx)
#pragma unroll
        for (int c0 = 0; c0 < NCO; c0 += 16) {
          float v[16], u[16];
          const uint32_t t = tmem_base + ((uint32_t)32 << 16) + (uint32_t)(kx * NB + c0);
          if (have) { tmem_ld16(t, v); tmem_ld16(t + NCO, u); tmem_ld_wait(); }
#pragma unroll
          for (int i = 0; i < 16; ++i) red[((size_t)kx * 32 + lane) * NCO + c0 + i] = have ? v[i] + u[i] : 0.f;
        }
    }
  }
  __syncthreads();
  if (warp == 0) {
    const bool have = n_items > (int)blockIdx.x;
    for (int kx = 0; kx < 5; ++kx)
#pragma unroll
      for (int c0 = 0; c0 < NCO; c0 += 16) {
        float v[16], u[16];
        const uint32_t t = tmem_base + (uint32_t)(kx * NB + c0);
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
              if (SPLIT3) r += (have ? u[i] : 0.f) + red[((size_t)kx * 32 + lane) * NCO + co];
              slab[((size_t)(ky * 5 + kx) * cin_pad + lane) * cout_total + co0 + co] = r;
            }
          }
        }
      }
  }
  // bias gradient: the `ones` row (TMEM lane ONES_ROW) of tap (ky = 0, kx = 0); other CTAs contribute zero
  if (warp == Cfg::ONES_ROW / 32) {
    const bool have = n_items > (int)blockIdx.x && ky == 0;
#pragma unroll
    for (int c0 = 0; c0 < NCO; c0 += 16) {
      float v[16], u[16];
      const uint32_t t = tmem_base + ((uint32_t)Cfg::ONES_ROW << 16) + (uint32_t)c0;
      if (have) {
        tmem_ld16(t, v);
        if (SPLIT3) tmem_ld16(t + NCO, u);
        tmem_ld_wait();
      }
      if (lane == 0 && ky == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < g_chunks * 4)
            slab[(size_t)25 * cin_pad * cout_total + co0 + c0 + i] = have ? v[i] + (SPLIT3 ? u[i] : 0.f) : 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ partial, int S, int n_w, int n_b, float* __restrict__ dwt,
                                       float* __restrict__ dbias) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int tot = n_w + n_b;
  if (i >= tot) return;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += partial[(size_t)k * tot + i];
  if (i < n_w) dwt[i] += s;
  else if (dbias) dbias[i - n_w] += s;
}

int wg_sets() { return 29; }       // 5 kernel rows x 29 sets = 145 persistent CTAs on 148 SMs

template <int NCO, bool SPLIT3>
int launch_wgrad_tc(cudaStream_t st, int N, int D, const ConvInput& in, const float4* dpre, int g_planes, int g_chunk0,
                    int g_chunks, int cout_total, int co0, float* scratch) {
  using Cfg = WgCfg<NCO, SPLIT3>;
  size_t smem = 2 * (size_t)Cfg::STAGE_BYTES;
  size_t red = (size_t)5 * 32 * NCO * sizeof(float);
  if (red > smem) smem = red;
  OP3_REQUIRE(smem <= 226 * 1024, "conv5x5_wgrad_tc: %zu bytes of shared memory", smem);
  static bool attr_set = false;
  if (!attr_set) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_wgrad_tc_kernel<NCO, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  dim3 grid(wg_sets(), 5);
  conv5x5_wgrad_tc_kernel<NCO, SPLIT3><<<grid, WG_THREADS, smem, st>>>(in, N, D, dpre, g_planes, g_chunk0, g_chunks, cout_total,
                                                                      co0, scratch);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

}  // namespace

size_t conv5x5_wgrad_tc_scratch_floats(int cin_pad, int cout) {
  return (size_t)wg_sets() * ((size_t)25 * cin_pad * cout + cout);
}

// dwt [25][cin_pad][cout] and dbias [cout] are ACCUMULATED (same contract as conv5x5_wgrad).
int conv5x5_wgrad_tc(cudaStream_t st, int precision, int N, int D, const ConvInput& in, int cout, const float4* dpre,
                     float* scratch, float* dwt, float* dbias) {
  if (N <= 0) return OP3_OK;
  OP3_REQUIRE(cout % 4 == 0 && cout <= 64 && in.total_chunks <= 8, "conv5x5_wgrad_tc: unsupported cin chunks %d / cout %d",
              in.total_chunks, cout);
  const int cin_real = in.real_cin ? in.real_cin : in.total_chunks * 4;
  ProfScope prof(st, KID_CONV_WGRAD, 2.0 * 25 * cin_real * cout * (double)N * D * D);
  const bool s3 = precision == OP3_PREC_TF32X3;
  const int planes = cout / 4;
  // every CTA writes its full region of its slab; slabs of different kernel rows are disjoint, so no zero fill is needed,
  // but halves / rows that are never written must not leak old data: clear once.
  const int cin_pad = in.total_chunks * 4;
  const size_t slab = (size_t)25 * cin_pad * cout + cout;
  OP3_CHECK(cudaMemsetAsync(scratch, 0, (size_t)wg_sets() * slab * sizeof(float), st));
  if (cout <= 16) {
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
  wgrad_tc_reduce_kernel<<<ceil_div(n_w + cout, 256), 256, 0, st>>>(scratch, wg_sets(), n_w, cout, dwt, dbias);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}


int conv5x5_wgrad_auto(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float4* dpre, float* scratch,
                       float* dwt, float* dbias) {
  const int prec = op3_get_precision();
  if (prec == OP3_PREC_FP32) return conv5x5_wgrad(st, N, D, in, cout, dpre, scratch, dwt, dbias);
  return conv5x5_wgrad_tc(st, prec, N, D, in, cout, dpre, scratch, dwt, dbias);
}

}  // namespace op3

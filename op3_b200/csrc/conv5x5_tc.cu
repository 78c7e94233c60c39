// tcgen05 implicit-GEMM 5x5 convolution (forward; also dgrad through transformed weights) for sm_100a.
//
//   D[pixels, cout] += A[pixels, (tap, ci)] * W[(tap, ci), cout]      M = 128 pixels / MMA, N = cout, K = 8 (tf32)
//
// * The activation halo of a group of G 128-pixel tiles sits in shared memory as [k-chunk][padded-linear pixel][4 ch]
//   (16-byte rows, no swizzle).  In padded-linear pixel space (row pitch D+4) every tap is a pure shift, so the A
//   operand of tap (ky,kx) is the SAME buffer with the descriptor start address advanced by (ky*(D+4)+kx)*16 bytes
//   (verified on B200 by tests/probes/probe_tcgen05.cu): no im2col copy is ever made; the 4 pad columns of each
//   row produce garbage outputs that the epilogue drops.
// * Accumulators live in TMEM (one 128-lane x N-column block per tile); one elected thread issues all MMAs.
// * kind::tf32 truncates fp32 operands to their top 19 bits (measured).  PRECISION 3xTF32 ("parity mode") therefore
//   needs only ONE extra operand copy: a_lo = a - trunc(a), produced in shared memory by the loader threads, and
//   weights pre-split on the host side of the kernel (op3_prepare):  D += a*[w_hi;w_lo] (one MMA, N = 2*cout, the two
//   halves are summed in the epilogue) ; D[:, :cout] += a_lo*w_hi.   1xTF32 ("speed mode") issues a*w only.
// * Loader threads use cp.async (zero-fill for the conv padding / image borders), make their writes visible to the
//   async proxy (fence.proxy.async) and hand stages to the MMA thread through mbarriers; tcgen05.commit hands them back.
#include "common.cuh"
#include "tc_common.cuh"

namespace op3 {

namespace {

constexpr int TC_THREADS = 160;        // warps 0-3: loaders + epilogue (TMEM lane quadrants), warp 4: MMA issuer

template <int NOUT, bool SPLIT3>
struct TcCfg {
  static constexpr int GMAX = SPLIT3 ? 4 : 8;                    // max tiles per CTA (runtime G <= GMAX)
  static constexpr int NB = SPLIT3 ? 2 * NOUT : NOUT;            // B rows per tap slab
  static constexpr int ACC = NB;                                 // TMEM columns per tile
  static constexpr int NSTAGE = SPLIT3 ? 2 : (NOUT == 64 ? 2 : 3);
  static constexpr int TMEM_COLS = (GMAX * ACC <= 32) ? 32 : (GMAX * ACC <= 64) ? 64 : (GMAX * ACC <= 128) ? 128 : (GMAX * ACC <= 256) ? 256 : 512;
  static_assert(GMAX * ACC <= 512, "accumulators exceed TMEM");
};

template <int NOUT, bool SPLIT3>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv5x5_tc_kernel(ConvInput in, int D, int G, const float* __restrict__ wt, const float* __restrict__ bias, int cout_real,
                  float4* __restrict__ out, int out_planes, int out_chunk0, int out_chunks, int epi,
                  const float4* __restrict__ ref) {
  using Cfg = TcCfg<NOUT, SPLIT3>;
  constexpr int NB = Cfg::NB, ACC = Cfg::ACC;
  const int P = D + 4;                                   // padded row pitch
  const int PH = 128 * G + 4 * P + 4;                    // halo pixels of one group
  const int tiles_img = (D * P + 127) / 128;
  const int groups_img = (tiles_img + G - 1) / G;
  const int n = blockIdx.x / groups_img, grp = blockIdx.x % groups_img;
  const int tile0 = grp * G;
  const int ntile = min(G, tiles_img - tile0);
  const int q0 = tile0 * 128;                            // first padded-linear output pixel of this group
  const int ngroups = (in.total_chunks + 1) / 2;         // K groups of 8 channels
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

  // ---- shared memory carve-up: per stage [A raw | A lo (3x) | W slab]
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int a_bytes = 2 * PH * 16;
  const int w_bytes = 25 * 2 * NB * 16;
  const int stage_bytes = a_bytes * (SPLIT3 ? 2 : 1) + w_bytes;
  constexpr int NSTAGE = Cfg::NSTAGE;
  __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], accum_bar;
  __shared__ uint32_t tmem_base_s;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ===================== loaders =====================
    for (int cg = 0; cg < ngroups; ++cg) {
      const int s = cg % NSTAGE;
      if (cg >= NSTAGE) mbar_wait(&empty_bar[s], ((cg / NSTAGE) - 1) & 1);
      uint8_t* st = smem_raw + (size_t)s * stage_bytes;
      float4* a_raw = reinterpret_cast<float4*>(st);
      float4* a_lo = reinterpret_cast<float4*>(st + a_bytes);
      float4* w_s = reinterpret_cast<float4*>(st + a_bytes * (SPLIT3 ? 2 : 1));
      // weights: contiguous slab of this K group
      const float4* wsrc = reinterpret_cast<const float4*>(wt) + (size_t)cg * (25 * 2 * NB);
      for (int e = tid; e < 25 * 2 * NB; e += 128) cp_async16(&w_s[e], &wsrc[e], true);
      // activations: 2 k-chunks x PH halo pixels
      for (int kc = 0; kc < 2; ++kc) {
        const int c = cg * 2 + kc;
        const float4* src = c < in.total_chunks ? tc_src_chunk_ptr(in, c, n, D) : nullptr;
        for (int h = tid; h < PH; h += 128) {
          int q = q0 + h;
          int iy = q / P - 2, ix = q % P - 2;
          bool ok = src != nullptr && iy >= 0 && iy < D && ix >= 0 && ix < D;
          cp_async16(&a_raw[kc * PH + h], ok ? src + (size_t)iy * D + ix : reinterpret_cast<const float4*>(wt), ok);
        }
      }
      cp_async_commit();
      cp_async_wait<0>();
      // The tensor core TRUNCATES fp32 operands to tf32, which biases every product towards zero.  Each loader thread
      // therefore rounds the granules it loaded itself to nearest tf32 in place (a_hi) and, in 3xTF32 mode, also
      // stores a_lo = rn_tf32(a - a_hi); both are exactly representable, so the hardware truncation becomes a no-op.
      for (int kc = 0; kc < 2; ++kc)
        for (int h = tid; h < PH; h += 128) {
          float4 v = a_raw[kc * PH + h];
          float4 hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
          a_raw[kc * PH + h] = hi;
          if (SPLIT3)
            a_lo[kc * PH + h] = make_float4(tf32_rn(v.x - hi.x), tf32_rn(v.y - hi.y), tf32_rn(v.z - hi.z), tf32_rn(v.w - hi.w));
        }
      fence_proxy_async();
      mbar_arrive(&full_bar[s]);
    }
    // ===================== epilogue =====================
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    for (int t = 0; t < ntile; ++t) {
      const int p = q0 + 128 * t + warp * 32 + lane;
      const int y = p / P, x = p % P;
      const bool valid = y < D && x < D;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * ACC);
#pragma unroll
      for (int c0 = 0; c0 < NOUT; c0 += 16) {
        float v[16];
        tmem_ld16(trow + c0, v);
        if (SPLIT3) {
          float u[16];
          tmem_ld16(trow + NOUT + c0, u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += u[j];
        } else {
          tmem_ld_wait();
        }
        if (valid) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int c4 = c0 / 4 + j4;
            if (c4 >= out_chunks) continue;
            float4 r = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            if (bias) {
              const int cb = c4 * 4;
              r.x += cb < cout_real ? bias[cb] : 0.f;
              r.y += cb + 1 < cout_real ? bias[cb + 1] : 0.f;
              r.z += cb + 2 < cout_real ? bias[cb + 2] : 0.f;
              r.w += cb + 3 < cout_real ? bias[cb + 3] : 0.f;
            }
            const size_t o = (((size_t)n * out_planes + out_chunk0 + c4) * D + y) * D + x;
            if (epi == EPI_ELU) {
              r.x = elu_f(r.x); r.y = elu_f(r.y); r.z = elu_f(r.z); r.w = elu_f(r.w);
            } else if (epi == EPI_MUL_ELUGRAD) {
              float4 a = ref[o];
              r.x *= elu_grad_from_out(a.x); r.y *= elu_grad_from_out(a.y);
              r.z *= elu_grad_from_out(a.z); r.w *= elu_grad_from_out(a.w);
            }
            out[o] = r;
          }
        }
      }
    }
  } else if (lane == 0) {
    // ===================== MMA issuer (one thread) =====================
    const uint32_t idesc1 = make_idesc_tf32(128, NB);
    const uint32_t idesc2 = make_idesc_tf32(128, NOUT);
    for (int cg = 0; cg < ngroups; ++cg) {
      const int s = cg % NSTAGE;
      mbar_wait(&full_bar[s], (cg / NSTAGE) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem_raw + (size_t)s * stage_bytes);
      const uint64_t ad = make_smem_desc(st, PH, 8);
      const uint64_t ald = make_smem_desc(st + a_bytes, PH, 8);
      const uint64_t bd = make_smem_desc(st + a_bytes * (SPLIT3 ? 2 : 1), NB, 8);
      const uint32_t a_lo32 = (uint32_t)ad, a_hi32 = (uint32_t)(ad >> 32);
      const uint32_t al_lo32 = (uint32_t)ald, al_hi32 = (uint32_t)(ald >> 32);
      const uint32_t b_lo32 = (uint32_t)bd, b_hi32 = (uint32_t)(bd >> 32);
      for (int t = 0; t < ntile; ++t) {
        const uint32_t dcol = tmem_base + (uint32_t)(t * ACC);
#pragma unroll
        for (int tap = 0; tap < 25; ++tap) {
          const uint32_t aoff = (uint32_t)(128 * t + (tap / 5) * P + (tap % 5));   // 16-byte units
          const uint32_t boff = (uint32_t)(tap * 2 * NB);
          mma_tf32(dcol, a_lo32 + aoff, a_hi32, b_lo32 + boff, b_hi32, idesc1, (cg > 0 || tap > 0) ? 1u : 0u);
          if (SPLIT3) mma_tf32(dcol, al_lo32 + aoff, al_hi32, b_lo32 + boff, b_hi32, idesc2, 1u);
        }
      }
      if (cg + 1 < ngroups) tc_commit(&empty_bar[s]);
    }
    tc_commit(&accum_bar);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int NOUT, bool SPLIT3>
int launch_tc(cudaStream_t st, int N, int D, const ConvInput& in, const float* wt, const float* bias, int cout_real,
              float4* out, int out_planes, int out_chunk0, int out_chunks, int epi, const float4* ref) {
  using Cfg = TcCfg<NOUT, SPLIT3>;
  const int P = D + 4;
  int G = Cfg::GMAX;
  size_t smem = 0;
  for (; G >= 1; --G) {
    const int PH = 128 * G + 4 * P + 4;
    smem = (size_t)Cfg::NSTAGE * ((size_t)2 * PH * 16 * (SPLIT3 ? 2 : 1) + 25 * 2 * Cfg::NB * 16);
    if (smem <= 226 * 1024) break;
  }
  OP3_REQUIRE(G >= 1, "conv5x5_tc: D=%d does not fit in shared memory", D);
  static bool attr_set = false;
  if (!attr_set) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_tc_kernel<NOUT, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  const int tiles_img = (D * P + 127) / 128;
  const int groups_img = (tiles_img + G - 1) / G;
  conv5x5_tc_kernel<NOUT, SPLIT3><<<N * groups_img, TC_THREADS, smem, st>>>(in, D, G, wt, bias, cout_real, out, out_planes,
                                                                          out_chunk0, out_chunks, epi, ref);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

}  // namespace

int tc_nout(int cout) { return cout <= 16 ? 16 : (cout <= 32 ? 32 : 64); }

size_t conv5x5_tc_weight_floats(int cin_pad, int cout) {
  const int groups = (cin_pad / 4 + 1) / 2;
  return (size_t)groups * 25 * 2 * (2 * tc_nout(cout)) * 4;       // sized for the 3x layout (hi and lo rows)
}

// Weight slab layout consumed by the kernel (built by op3_prepare): [K group of 8 ci][tap 25][k-chunk 2][row NB][4 ci]
// with NB = nout rows (1xTF32: raw fp32) or 2*nout rows (3xTF32: trunc_tf32(w) rows, then w - trunc rows).  cout = 64 in
// 3xTF32 mode is stored and run as two independent halves of 32 output channels (TMEM / smem budget).
int conv5x5_tc_forward(cudaStream_t st, int precision, int N, int D, const ConvInput& in, int cout, const float* wt,
                       const float* bias, float4* out, int epi, const float4* elu_ref) {
  if (N <= 0) return OP3_OK;
  OP3_REQUIRE(cout % 4 == 0 && cout <= 64, "conv5x5_tc: unsupported cout=%d", cout);
  const int cin_real = in.real_cin ? in.real_cin : in.total_chunks * 4;
  ProfScope prof(st, KID_CONV_FWD, 2.0 * 25 * cin_real * (cout == 20 ? 17 : cout) * (double)N * D * D);
  const int nout = tc_nout(cout);
  const int planes = cout / 4;
  if (precision == OP3_PREC_TF32X3) {
    if (nout == 16) return launch_tc<16, true>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref);
    if (nout == 32) return launch_tc<32, true>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref);
    const size_t half = (size_t)((in.total_chunks + 1) / 2) * 25 * 2 * 64 * 4;
    OP3_TRY((launch_tc<32, true>(st, N, D, in, wt, bias, 32, out, planes, 0, 8, epi, elu_ref)));
    return launch_tc<32, true>(st, N, D, in, wt + half, bias ? bias + 32 : nullptr, cout - 32, out, planes, 8, planes - 8, epi, elu_ref);
  }
  if (nout == 16) return launch_tc<16, false>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref);
  if (nout == 32) return launch_tc<32, false>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref);
  return launch_tc<64, false>(st, N, D, in, wt, bias, cout, out, planes, 0, planes, epi, elu_ref);
}


// fp32 CUDA-core kernel in exact mode, tcgen05 kernel otherwise
int conv5x5_auto(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float* wt_ffma, const float* wt_tc,
                 const float* bias, float4* out, int epi, const float4* elu_ref) {
  const int prec = op3_get_precision();
  if (prec == OP3_PREC_FP32) return conv5x5_forward(st, N, D, in, cout, wt_ffma, bias, out, epi, elu_ref);
  return conv5x5_tc_forward(st, prec, N, D, in, cout, wt_tc, bias, out, epi, elu_ref);
}

}  // namespace op3

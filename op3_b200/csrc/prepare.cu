// Derived weights (op3_prepare) and folding of their gradients back to PyTorch layouts (op3_finalize_grads).
//
// Decoder layer 1 (conv_networks.py:326-335 feeds a spatially CONSTANT latent plus two coordinate planes into a
// 5x5/pad-2 conv) collapses to:  pre1[n,co,y,x] = b[co] + sum_c z[n,c]*Wsum[cls(y,x)][co][c] + cmap[co][y][x]
// where the valid-tap set only depends on one of 25 border classes (hard part 2).
#include "layout.cuh"

namespace op3 {

namespace {

__host__ __device__ inline bool tap_valid(int cls_1d, int k) {
  int lo = 2 - cls_1d; if (lo < 0) lo = 0;
  int hi = 6 - cls_1d; if (hi > 4) hi = 4;
  return k >= lo && k <= hi;
}

__device__ __forceinline__ double lin_coord(int i, int D) {  // np.linspace(-1, 1, D)[i]
  if (i == D - 1) return 1.0;
  return -1.0 + (double)i * (2.0 / (double)(D - 1));
}

// W [cout][cin][5][5] -> wf [25][cin_pad][cout] and wd [25][cout][cin_pad] (flipped taps), mode 1 = refinement L1 perm
__global__ void prep_conv_kernel(const float* __restrict__ W, int cout, int cin, int cin_pad, int mode,
                                 float* __restrict__ wf, float* __restrict__ wd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int tot = 25 * cin_pad * cout;
  if (i >= tot) return;
  int co = i % cout, cip = (i / cout) % cin_pad, tap = i / (cout * cin_pad);
  int ci = mode == 1 ? ref_l1_channel(cip) : cip;
  float v = (ci >= 0 && ci < cin) ? W[((size_t)co * cin + ci) * 25 + tap] : 0.f;
  wf[i] = v;
  wd[((size_t)(24 - tap) * cout + co) * cin_pad + cip] = v;
}

__global__ void prep_wsum_kernel(const float* __restrict__ W1, const float* __restrict__ b1, int R,
                                 float* __restrict__ wsum, float* __restrict__ biasT) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int tot = NCLS * 32 * R;
  if (i < NCLS * 32) biasT[i] = b1[i % 32];
  if (i >= tot) return;
  int c = i % R, co = (i / R) % 32, cls = i / (R * 32);
  int cy = cls / 5, cx = cls % 5;
  const float* w = W1 + ((size_t)co * (R + 2) + c) * 25;
  float s = 0.f;
  for (int ky = 0; ky < 5; ++ky)
    for (int kx = 0; kx < 5; ++kx)
      if (tap_valid(cy, ky) && tap_valid(cx, kx)) s += w[ky * 5 + kx];
  wsum[i] = s;
}

__global__ void prep_cmap_kernel(const float* __restrict__ W1, int R, int D, float* __restrict__ cmap,
                                 float* __restrict__ coords) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < D * D) {
    int y = i / D, x = i % D;
    reinterpret_cast<float4*>(coords)[i] = make_float4((float)lin_coord(x, D), (float)lin_coord(y, D), 0.f, 0.f);
  }
  if (i >= 32 * D * D) return;
  int x = i % D, y = (i / D) % D, co = i / (D * D);
  const float* wx = W1 + ((size_t)co * (R + 2) + R) * 25;
  const float* wy = wx + 25;
  float s = 0.f;
  for (int ky = 0; ky < 5; ++ky) {
    int yy = y + ky - 2;
    if (yy < 0 || yy >= D) continue;
    float cyv = (float)lin_coord(yy, D);
    for (int kx = 0; kx < 5; ++kx) {
      int xx = x + kx - 2;
      if (xx < 0 || xx >= D) continue;
      s = fmaf(wx[ky * 5 + kx], (float)lin_coord(xx, D), s);
      s = fmaf(wy[ky * 5 + kx], cyv, s);
    }
  }
  cmap[(((size_t)(co / 4) * D + y) * D + x) * 4 + (co % 4)] = s;
}

// tcgen05 weight slab: element (g, tap, kc, row, e) of [groups][25][2][NB][4]; GEMM input channel = g*8+kc*4+e.
//   transposed = 0 (forward): value = W[co][ci][tap];  transposed = 1 (dgrad): value = W[ci_g][co_g][24-tap] where the
//   GEMM input channels are the layer's outputs.  mode 1 applies the refinement-L1 channel permutation to the layer's
//   INPUT channel index.  split3: rows [0,nout) = rn_tf32(w), rows [nout,2nout) = rn_tf32(w - rn_tf32(w)).
__global__ void prep_tc_kernel(const float* __restrict__ W, int cout_src, int cin_src, int mode, int transposed, int groups,
                               int nout, int split3, int co_base, float* __restrict__ slab) {
  const int NB = split3 ? 2 * nout : nout;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int tot = groups * 25 * 2 * NB * 4;
  if (i >= tot) return;
  int e = i % 4, row = (i / 4) % NB, kc = (i / (4 * NB)) % 2, tap = (i / (8 * NB)) % 25, g = i / (200 * NB);
  int k_in = g * 8 + kc * 4 + e;                 // GEMM K channel
  int n_out = (row % nout) + co_base;            // GEMM N channel
  float w = 0.f;
  if (!transposed) {
    int ci = mode == 1 ? ref_l1_channel(k_in) : k_in;
    if (ci >= 0 && ci < cin_src && n_out < cout_src) w = W[((size_t)n_out * cin_src + ci) * 25 + tap];
  } else {
    int ci = mode == 1 ? ref_l1_channel(n_out) : n_out;   // layer input channel = GEMM output
    if (ci >= 0 && ci < cin_src && k_in < cout_src) w = W[((size_t)k_in * cin_src + ci) * 25 + (24 - tap)];
  }
  // the tensor core truncates to tf32: store round-to-nearest values so that truncation is a no-op (see conv5x5_tc.cu)
  float hi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xFFFFE000u);
  if (split3) {
    float lo = w - hi;
    w = row < nout ? hi : __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u);
  } else {
    w = hi;
  }
  slab[i] = w;
}

// fp16 mode: scale[0] = power of two that brings max|W| into [2^13, 2^14) (see f16_scale_bits); one block
__global__ void __launch_bounds__(1024) prep_wscale_kernel(const float* __restrict__ W, int n, float* __restrict__ scale) {
  __shared__ float s_m[32];
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) m = fmaxf(m, fabsf(W[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 32; ++i) m = fmaxf(m, s_m[i]);
    scale[0] = __uint_as_float(f16_scale_bits(__float_as_uint(m)));
    scale[1] = scale[2] = scale[3] = 0.f;
  }
}

// fp16 slab: [K group of 16 ci][tap 25][k-chunk 2][row 2*nout][8 halves]; rows [0,nout) = rn_f16(s w), rows [nout,2nout) =
// rn_f16(s w - hi); same channel mapping as prep_tc_kernel.
__global__ void prep_tc_f16_kernel(const float* __restrict__ W, int cout_src, int cin_src, int mode, int transposed,
                                   int groups16, int nout, int co_base, const float* __restrict__ scale,
                                   __half* __restrict__ slab) {
  const int NB = 2 * nout;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int tot = groups16 * 25 * 2 * NB * 8;
  if (i >= tot) return;
  int e = i % 8, row = (i / 8) % NB, kc = (i / (8 * NB)) % 2, tap = (i / (16 * NB)) % 25, g = i / (400 * NB);
  int k_in = g * 16 + kc * 8 + e;                // GEMM K channel
  int n_out = (row % nout) + co_base;            // GEMM N channel
  float w = 0.f;
  if (!transposed) {
    int ci = mode == 1 ? ref_l1_channel(k_in) : k_in;
    if (ci >= 0 && ci < cin_src && n_out < cout_src) w = W[((size_t)n_out * cin_src + ci) * 25 + tap];
  } else {
    int ci = mode == 1 ? ref_l1_channel(n_out) : n_out;   // layer input channel = GEMM output
    if (ci >= 0 && ci < cin_src && k_in < cout_src) w = W[((size_t)k_in * cin_src + ci) * 25 + (24 - tap)];
  }
  const float ws = w * scale[0];
  const __half hi = __float2half_rn(ws);
  slab[i] = row < nout ? hi : __float2half_rn(ws - __half2float(hi));
}

// ---- gradient folding ------------------------------------------------------------------------------------
__global__ void fin_conv_kernel(const float* __restrict__ dwf, const float* __restrict__ dbf, int cout, int cin,
                                int cin_pad, int mode, float* __restrict__ dW, float* __restrict__ db) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cout && db) db[i] += dbf[i];
  int tot = 25 * cin_pad * cout;
  if (i >= tot) return;
  int co = i % cout, cip = (i / cout) % cin_pad, tap = i / (cout * cin_pad);
  int ci = mode == 1 ? ref_l1_channel(cip) : cip;
  if (ci >= 0 && ci < cin) dW[((size_t)co * cin + ci) * 25 + tap] += dwf[i];
}

__global__ void fin_wsum_kernel(const float* __restrict__ dwsum, const float* __restrict__ dbT, int R,
                                float* __restrict__ dW1, float* __restrict__ db1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 32) {
    float s = 0.f;
    for (int cls = 0; cls < NCLS; ++cls) s += dbT[cls * 32 + i];
    db1[i] += s;
  }
  int tot = 32 * R * 25;
  if (i >= tot) return;
  int tap = i % 25, c = (i / 25) % R, co = i / (25 * R);
  int ky = tap / 5, kx = tap % 5;
  float s = 0.f;
  for (int cls = 0; cls < NCLS; ++cls)
    if (tap_valid(cls / 5, ky) && tap_valid(cls % 5, kx)) s += dwsum[((size_t)cls * 32 + co) * R + c];
  dW1[((size_t)co * (R + 2) + c) * 25 + tap] += s;
}

// one block per (co, coord channel, tap): deterministic tree reduction over D*D pixels
__global__ void __launch_bounds__(256) fin_cmap_kernel(const float* __restrict__ dcmap, int R, int D,
                                                       float* __restrict__ dW1) {
  __shared__ float red[256];
  int b = blockIdx.x;
  int tap = b % 25, cc = (b / 25) % 2, co = b / 50;
  int ky = tap / 5, kx = tap % 5;
  float s = 0.f;
  for (int p = threadIdx.x; p < D * D; p += 256) {
    int y = p / D, x = p % D;
    int yy = y + ky - 2, xx = x + kx - 2;
    if (yy < 0 || yy >= D || xx < 0 || xx >= D) continue;
    float cv = cc == 0 ? (float)lin_coord(xx, D) : (float)lin_coord(yy, D);
    s = fmaf(dcmap[(((size_t)(co / 4) * D + y) * D + x) * 4 + (co % 4)], cv, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dW1[((size_t)co * (R + 2) + R + cc) * 25 + tap] += red[0];
}

}  // namespace
}  // namespace op3

using namespace op3;

extern "C" size_t op3_prepared_floats(const Op3Dims* d) {
  if (validate_dims(d) != OP3_OK) return 0;
  return prep_layout(d).total;
}

extern "C" int op3_prepare(const Op3Dims* d, const Op3Params* p, float* prep, void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(p && prep, "op3_prepare: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  PrepLayout L = prep_layout(d);
  const int R = d->Rd + d->Rs, D = d->D;
  // a standalone DecoderNetwork / RefinementNetwork only owns its own weights: NULL groups are skipped
  const bool has_dec = p->dec_conv_w[0] != nullptr, has_ref = p->ref_conv_w[0] != nullptr;
  OP3_REQUIRE(has_dec || has_ref, "op3_prepare: neither decoder nor refinement weights given");
  const int dec_cout[4] = {32, 32, 32, 4};
  if (has_dec) {
    prep_wsum_kernel<<<ceil_div(NCLS * 32 * R, 256), 256, 0, st>>>(p->dec_conv_w[0], p->dec_conv_b[0], R, prep + L.wsum,
                                                                  prep + L.biasT);
    OP3_COUNT_LAUNCH();
    prep_cmap_kernel<<<ceil_div(32 * D * D, 256), 256, 0, st>>>(p->dec_conv_w[0], R, D, prep + L.cmap, prep + L.coords);
    OP3_COUNT_LAUNCH();
  }
  for (int i = 0; i < 4 && has_dec; ++i) {
    prep_conv_kernel<<<ceil_div(25 * 32 * dec_cout[i], 256), 256, 0, st>>>(p->dec_conv_w[i + 1], dec_cout[i], 32, 32, 0,
                                                                           prep + L.dec_wf[i], prep + L.dec_wd[i]);
    OP3_COUNT_LAUNCH();
    OP3_CHECK(cudaMemcpyAsync(prep + L.dec_bf[i], p->dec_conv_b[i + 1], dec_cout[i] * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
  }
  const int ref_cin[3] = {17, 32, 32}, ref_pad[3] = {REF_CIN_PAD, 32, 32}, ref_cout[3] = {32, 32, d->Rs};
  for (int i = 0; i < 3 && has_ref; ++i) {
    prep_conv_kernel<<<ceil_div(25 * ref_pad[i] * ref_cout[i], 256), 256, 0, st>>>(
        p->ref_conv_w[i], ref_cout[i], ref_cin[i], ref_pad[i], i == 0 ? 1 : 0, prep + L.ref_wf[i], prep + L.ref_wd[i]);
    OP3_COUNT_LAUNCH();
    OP3_CHECK(cudaMemcpyAsync(prep + L.ref_bf[i], p->ref_conv_b[i], ref_cout[i] * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
  }
  // tcgen05 slabs for the current precision mode
  const int prec = op3_get_precision();
  if (prec != OP3_PREC_FP32) {
    const int s3 = prec == OP3_PREC_TF32X3;
    auto build = [&](const float* W, int cout_src, int cin_src, int mode, int transposed, int gemm_cin_pad, int gemm_cout,
                     float* slab) {
      if (prec == OP3_PREC_F16X3) {
        const int nout_full = tc_nout(gemm_cout);
        // narrow layers get 32-row slabs (zero rows beyond cout) when the quint-tap kernel will run them (conv5x5_tc_forward)
        const bool widen = nout_full == 16 && tc_quint_fits(d->D, gemm_cin_pad / 4);
        const int halves = nout_full == 64 ? 2 : 1, nout = (halves == 2 || widen) ? 32 : nout_full;
        const int groups16 = (gemm_cin_pad / 4 + 3) / 4;
        const size_t per = conv5x5_tc_f16_slab_floats(gemm_cin_pad, nout);
        const int tot = groups16 * 25 * 2 * (2 * nout) * 8;
        for (int h = 0; h < halves; ++h) {
          float* base = slab + (size_t)h * per;
          prep_wscale_kernel<<<1, 1024, 0, st>>>(W, cout_src * cin_src * 25, base + per - 4);
          OP3_COUNT_LAUNCH();
          prep_tc_f16_kernel<<<ceil_div(tot, 256), 256, 0, st>>>(W, cout_src, cin_src, mode, transposed, groups16, nout, 32 * h,
                                                                base + per - 4, reinterpret_cast<__half*>(base));
          OP3_COUNT_LAUNCH();
        }
        return;
      }
      const int groups = (gemm_cin_pad / 4 + 1) / 2;
      const int nout_full = tc_nout(gemm_cout);
      const int halves = (s3 && nout_full == 64) ? 2 : 1;          // see conv5x5_tc_forward
      const int nout = halves == 2 ? 32 : nout_full;
      const int NB = s3 ? 2 * nout : nout;
      const int per = groups * 25 * 2 * NB * 4;
      for (int h = 0; h < halves; ++h) {
        prep_tc_kernel<<<ceil_div(per, 256), 256, 0, st>>>(W, cout_src, cin_src, mode, transposed, groups, nout, s3, 32 * h,
                                                          slab + (size_t)h * per);
        OP3_COUNT_LAUNCH();
      }
    };
    for (int i = 0; i < 4 && has_dec; ++i) {
      build(p->dec_conv_w[i + 1], dec_cout[i], 32, 0, 0, 32, dec_cout[i], prep + L.dec_tf[i]);
      build(p->dec_conv_w[i + 1], dec_cout[i], 32, 0, 1, dec_cout[i], 32, prep + L.dec_td[i]);
    }
    for (int i = 0; i < 3 && has_ref; ++i) {
      build(p->ref_conv_w[i], ref_cout[i], ref_cin[i], i == 0 ? 1 : 0, 0, ref_pad[i], ref_cout[i], prep + L.ref_tf[i]);
      build(p->ref_conv_w[i], ref_cout[i], ref_cin[i], i == 0 ? 1 : 0, 1, ref_cout[i], ref_pad[i], prep + L.ref_td[i]);
    }
  }
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

extern "C" int op3_finalize_grads(const Op3Dims* d, const float* pg, const Op3Params* g, void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(pg && g, "op3_finalize_grads: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  PrepLayout L = prep_layout(d);
  const int R = d->Rd + d->Rs, D = d->D;
  fin_wsum_kernel<<<ceil_div(32 * R * 25, 256), 256, 0, st>>>(pg + L.wsum, pg + L.biasT, R, g->dec_conv_w[0],
                                                             g->dec_conv_b[0]);
  OP3_COUNT_LAUNCH();
  fin_cmap_kernel<<<32 * 2 * 25, 256, 0, st>>>(pg + L.cmap, R, D, g->dec_conv_w[0]);
  OP3_COUNT_LAUNCH();
  const int dec_cout[4] = {32, 32, 32, 4};
  for (int i = 0; i < 4; ++i) {
    fin_conv_kernel<<<ceil_div(25 * 32 * dec_cout[i], 256), 256, 0, st>>>(pg + L.dec_wf[i], pg + L.dec_bf[i], dec_cout[i],
                                                                          32, 32, 0, g->dec_conv_w[i + 1],
                                                                          g->dec_conv_b[i + 1]);
    OP3_COUNT_LAUNCH();
  }
  const int ref_cin[3] = {17, 32, 32}, ref_pad[3] = {REF_CIN_PAD, 32, 32}, ref_cout[3] = {32, 32, d->Rs};
  for (int i = 0; i < 3; ++i) {
    fin_conv_kernel<<<ceil_div(25 * ref_pad[i] * ref_cout[i], 256), 256, 0, st>>>(
        pg + L.ref_wf[i], pg + L.ref_bf[i], ref_cout[i], ref_cin[i], ref_pad[i], i == 0 ? 1 : 0, g->ref_conv_w[i],
        g->ref_conv_b[i]);
    OP3_COUNT_LAUNCH();
  }
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

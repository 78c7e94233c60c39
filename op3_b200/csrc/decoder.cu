// K-slot spatial-broadcast decoder: replaces DecoderNetwork.forward (op3/torch/op3_modules/decoder_network.py:83-88)
// and BroadcastCNN.forward/apply_forward (op3/torch/conv_networks.py:326-350) plus their backward.
//   layer 1 : collapsed (class-sum GEMM + coordinate map, never materialising the (N,130,D,D) broadcast input)
//   layer 2-5: 5x5 convs over float4 channel-chunk activations, ELU fused in the epilogue
#include "layout.cuh"

namespace op3 {

namespace {

__device__ __forceinline__ int border_class(int v, int D) { return v < 2 ? v : (v >= D - 2 ? 4 - (D - 1 - v) : 2); }

// a1[n][c4][y][x] = ELU(T[n][cls(y,x)*32 + c4*4 ..] + cmap[c4][y][x])
// One thread owns one float4 position of the coordinate map and streams EXP_NG images through it (the position's border
// class and map value are computed once).  maxbits (optional): per-image largest magnitude of a1 for the fp16 conv mode
// (zeroed by the caller).
// FAST (fp16 conv mode): ELU through ex2.approx like the tensor-core conv epilogues of that mode (conv5x5_tc.cu, elu_tc) --
// with expm1f the kernel is bound by its ~120 instructions per float4 (45 us for 134 MB), not by the store bandwidth.
constexpr int EXP_NG = 8;
template <bool FAST>
__global__ void __launch_bounds__(256) dec_l1_expand_kernel(const float* __restrict__ T, const float4* __restrict__ cmap,
                                                            int N, int D, float4* __restrict__ a1, uint32_t* __restrict__ maxbits) {
  const int n0 = blockIdx.y * EXP_NG;
  const int i = blockIdx.x * 256 + threadIdx.x;  // over 8*D*D
  const int tot = 8 * D * D;
  const bool live = i < tot;
  __shared__ float s_mx[EXP_NG][8];
  int cls = 0, c4 = 0;
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const int x = i % D, y = (i / D) % D;
    c4 = i / (D * D);
    cls = border_class(y, D) * 5 + border_class(x, D);
    m = cmap[i];
  }
#pragma unroll
  for (int g = 0; g < EXP_NG; ++g) {
    const int n = n0 + g;
    float mx = 0.f;
    if (n >= N) { if ((threadIdx.x & 31) == 0) s_mx[g][threadIdx.x >> 5] = 0.f; continue; }
    if (live) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(T + (size_t)n * (NCLS * 32) + cls * 32 + c4 * 4));
      auto elu = [](float x) { return FAST ? (x > 0.f ? x : __expf(x) - 1.f) : elu_f(x); };
      const float4 v = make_float4(elu(t.x + m.x), elu(t.y + m.y), elu(t.z + m.z), elu(t.w + m.w));
      a1[(size_t)n * tot + i] = v;
      mx = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    if (maxbits != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if ((threadIdx.x & 31) == 0) s_mx[g][threadIdx.x >> 5] = mx;
    }
  }
  if (maxbits != nullptr) {                                // one conditional atomic per block and image
    __syncthreads();
    const int g = threadIdx.x;
    if (g < EXP_NG && n0 + g < N) {
      float mx = s_mx[g][0];
#pragma unroll
      for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_mx[g][w]);
      if (__float_as_uint(mx) > __ldcg(&maxbits[n0 + g])) atomicMax(&maxbits[n0 + g], __float_as_uint(mx));
    }
  }
}

// dT[n][cls*32 + co] = sum over pixels of class cls of g1[n][co][y][x]; one CTA per (chunk, n); deterministic.
__global__ void __launch_bounds__(256) dec_l1_class_reduce_kernel(const float4* __restrict__ g1, int D,
                                                                  float* __restrict__ dT) {
  __shared__ float4 wacc[8][NCLS];
  const int c4 = blockIdx.x, n = blockIdx.y;
  const int lane = threadIdx.x % 32, w = threadIdx.x / 32;
  for (int i = lane; i < NCLS; i += 32) wacc[w][i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  const float4* src = g1 + ((size_t)n * 8 + c4) * D * D;
  // interior columns: per-lane partial sums run over all rows of one row class and are reduced across the warp once per class
  // (a warp meets at most three classes), not once per row -- the kernel was bound by those 20 shuffles per row
  float4 mid = make_float4(0.f, 0.f, 0.f, 0.f);
  int mid_cy = -1;
  auto flush = [&]() {
    if (mid_cy >= 0) {
      mid.x = warp_sum(mid.x); mid.y = warp_sum(mid.y); mid.z = warp_sum(mid.z); mid.w = warp_sum(mid.w);
      if (lane == 0) {
        float4 a = wacc[w][mid_cy * 5 + 2];
        a.x += mid.x; a.y += mid.y; a.z += mid.z; a.w += mid.w;
        wacc[w][mid_cy * 5 + 2] = a;
      }
      __syncwarp();
    }
    mid = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto add = [&](int cy, int x, const float4& v) {
    const int cx = border_class(x, D);
    if (cx == 2) {
      mid.x += v.x; mid.y += v.y; mid.z += v.z; mid.w += v.w;
    } else {  // exactly one lane owns each border column of a row
      float4 a = wacc[w][cy * 5 + cx];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      wacc[w][cy * 5 + cx] = a;
    }
  };
  if (D == 64) {           // the warp's eight rows x two float4 per lane: all sixteen loads in flight before the first use
    float4 v[8][2];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      v[r][0] = src[(size_t)(w + 8 * r) * 64 + lane];
      v[r][1] = src[(size_t)(w + 8 * r) * 64 + lane + 32];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int cy = border_class(w + 8 * r, 64);
      if (cy != mid_cy) { flush(); mid_cy = cy; }
      add(cy, lane, v[r][0]);
      add(cy, lane + 32, v[r][1]);
    }
  } else {
    for (int y = w; y < D; y += 8) {
      const int cy = border_class(y, D);
      if (cy != mid_cy) { flush(); mid_cy = cy; }
      for (int x = lane; x < D; x += 32) add(cy, x, src[(size_t)y * D + x]);
    }
  }
  flush();
  __syncthreads();
  if (threadIdx.x < NCLS) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < 8; ++k) {
      float4 a = wacc[k][threadIdx.x];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
    *reinterpret_cast<float4*>(dT + (size_t)n * (NCLS * 32) + threadIdx.x * 32 + c4 * 4) = s;
  }
}

// dcmap[i] += sum_n g1[n][i]   (i over 8*D*D float4).  A block covers 64 positions; its four 64-thread groups sum a quarter
// of the images each (4x the blocks and a 4x shorter serial chain), partial sums are added in a fixed order.
__global__ void __launch_bounds__(256) dec_l1_dcmap_kernel(const float4* __restrict__ g1, int N, int D,
                                                           float4* __restrict__ dcmap) {
  __shared__ float4 part[3][64];
  const int grp = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int i = blockIdx.x * 64 + t;
  const int tot = 8 * D * D;
  const int nq = (N + 3) / 4, nb = grp * nq, ne = min(N, nb + nq);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < tot) {
#pragma unroll 4
    for (int n = nb; n < ne; ++n) {
      float4 v = g1[(size_t)n * tot + i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  if (grp > 0) part[grp - 1][t] = s;
  __syncthreads();
  if (grp == 0 && i < tot) {
#pragma unroll
    for (int g = 0; g < 3; ++g) { float4 p = part[g][t]; s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w; }
    float4 o = dcmap[i];
    o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
    dcmap[i] = o;
  }
}

struct DecActs {
  float* T;        // [N][800]
  float4* a[4];    // a1..a4: [N][8][D][D] float4
};
inline DecActs dec_acts(const Op3Dims* d, int N, float* base) {
  DecActs A;
  size_t plane = (size_t)N * 32 * d->D * d->D;
  A.T = base;
  size_t o = ((size_t)N * NCLS * 32 + 3) & ~(size_t)3;
  for (int i = 0; i < 4; ++i) A.a[i] = reinterpret_cast<float4*>(base + o + i * plane);
  return A;
}

inline ConvInput single_src(const float4* p, int chunks) {
  ConvInput in{};
  in.src[0].ptr = p; in.src[0].nchunks = chunks; in.src[0].div = 1;
  in.nsrc = 1; in.total_chunks = chunks;
  return in;
}

}  // namespace
}  // namespace op3

using namespace op3;

extern "C" size_t op3_decoder_acts_floats(const Op3Dims* d, int N) {
  if (validate_dims(d) != OP3_OK) return 0;
  return (((size_t)N * NCLS * 32 + 3) & ~(size_t)3) + 4 * (size_t)N * 32 * d->D * d->D + tc_scratch_floats(N);
}

extern "C" size_t op3_decoder_scratch_floats(const Op3Dims* d, int N) {
  if (validate_dims(d) != OP3_OK) return 0;
  // four gradient planes: each layer's dPre keeps its own buffer, so a weight-gradient kernel that runs late on the
  // weight-gradient lane (op3_set_wgrad_overlap) never sees its operand overwritten by a later dgrad
  return 4 * (size_t)N * 32 * d->D * d->D + (((size_t)N * NCLS * 32 + 3) & ~(size_t)3) +
         conv5x5_wgrad_scratch_floats(32, 32) + tc_scratch_floats(N);
}

extern "C" int op3_decoder_fwd(const Op3Dims* d, int N, const float* prep, const float* z, float* acts, float* out,
                               void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(prep && z && acts && out && N > 0, "op3_decoder_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  PrepLayout L = prep_layout(d);
  const int R = d->Rd + d->Rs, D = d->D;
  DecActs A = dec_acts(d, N, acts);
  float* tcs = acts + (((size_t)N * NCLS * 32 + 3) & ~(size_t)3) + 4 * (size_t)N * 32 * D * D;   // fp16-mode operand maxima
  // T[n][cls*32+co] = b[co] + sum_c z[n][c] * Wsum[cls*32+co][c]
  OP3_TRY(gemm(st, N, NCLS * 32, R, z, R, 0, prep + L.wsum, R, 1, A.T, NCLS * 32, prep + L.biasT, ACT_NONE, 0));
  // fp16 mode: slots 1..4 of tcs keep the per-image maxima of a1..a4 (written by their producers, reused by backward)
  const bool f16 = op3_get_precision() == OP3_PREC_F16X3;
  if (f16) OP3_CHECK(cudaMemsetAsync(tcs, 0, tc_scratch_floats(N) * sizeof(float), st));
  if (f16)
    dec_l1_expand_kernel<true><<<dim3(ceil_div(8 * D * D, 256), ceil_div(N, EXP_NG)), 256, 0, st>>>(
        A.T, reinterpret_cast<const float4*>(prep + L.cmap), N, D, A.a[0], tc_slot(tcs, N, 1));
  else
    dec_l1_expand_kernel<false><<<dim3(ceil_div(8 * D * D, 256), ceil_div(N, EXP_NG)), 256, 0, st>>>(
        A.T, reinterpret_cast<const float4*>(prep + L.cmap), N, D, A.a[0], nullptr);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  // layers 2-4 are the implicit-GEMM hot spot (M = N*D*D pixels, N = 32, K = 800); layer 5 has 4 outputs
  for (int i = 0; i < 3; ++i) {
    ConvInput in = single_src(A.a[i], 8);
    if (f16) { in.maxbits = tc_slot(tcs, N, 1 + i); in.out_maxbits = tc_slot(tcs, N, 2 + i); }
    OP3_TRY(conv5x5_auto(st, N, D, in, 32, prep + L.dec_wf[i], prep + L.dec_tf[i], prep + L.dec_bf[i], A.a[i + 1], EPI_ELU,
                         nullptr, tcs));
  }
  ConvInput in5 = single_src(A.a[3], 8);
  if (f16) in5.maxbits = tc_slot(tcs, N, 4);               // recorded by layer 4's epilogue
  OP3_TRY(conv5x5_auto(st, N, D, in5, 4, prep + L.dec_wf[3], prep + L.dec_tf[3], prep + L.dec_bf[3],
                       reinterpret_cast<float4*>(out), EPI_NONE, nullptr, tcs));
  return OP3_OK;
}

extern "C" int op3_decoder_bwd(const Op3Dims* d, int N, const float* prep, const float* z, const float* acts,
                               const float* d_out, float* scratch, float* dz, int accumulate_dz, float* pg,
                               void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(prep && z && acts && d_out && scratch && dz && N > 0, "op3_decoder_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  PrepLayout L = prep_layout(d);
  const int R = d->Rd + d->Rs, D = d->D;
  DecActs A = dec_acts(d, N, const_cast<float*>(acts));
  const size_t plane = (size_t)N * 32 * D * D;
  float4* gbuf[4];
  for (int i = 0; i < 4; ++i) gbuf[i] = reinterpret_cast<float4*>(scratch + i * plane);
  float* dT = scratch + 4 * plane;
  float* wscr = dT + (((size_t)N * NCLS * 32 + 3) & ~(size_t)3);
  float* tcs = wscr + conv5x5_wgrad_scratch_floats(32, 32);

  // fp16 mode: the forward pass left the maxima of a1..a4 in slots 1..4 of the acts' scratch; the dgrad kernels record
  // the maxima of the gradients they write in slots 1..4 of this call's scratch
  const bool f16 = op3_get_precision() == OP3_PREC_F16X3;
  const float* atcs = acts + (((size_t)N * NCLS * 32 + 3) & ~(size_t)3) + 4 * (size_t)N * 32 * D * D;
  if (f16) OP3_CHECK(cudaMemsetAsync(tcs, 0, tc_scratch_floats(N) * sizeof(float), st));
  // layer 5 (32 -> 4, no activation): g5 = d_out
  const float4* g = reinterpret_cast<const float4*>(d_out);
  const uint32_t* gmax = nullptr;  // maxima of g (NULL: not recorded)
  int gch = 1;  // chunks of g
  for (int i = 3; i >= 0; --i) {  // conv layer i+2, input a[i], output gradient g (dPre of that layer)
    const int cout = i == 3 ? 4 : 32;
    float4* gn = gbuf[i];
    // dgrad: conv over g with flipped/transposed weights, times ELU'(a[i]) -> dPre of the layer below
    ConvInput gin = single_src(g, gch);
    if (f16) { gin.maxbits = gmax; gin.out_maxbits = tc_slot(tcs, N, 1 + i); }
    OP3_TRY(conv5x5_auto(st, N, D, gin, 32, prep + L.dec_wd[i], prep + L.dec_td[i], nullptr, gn, EPI_MUL_ELUGRAD, A.a[i], tcs));
    if (pg) {
      ConvInput xin = single_src(A.a[i], 8);
      if (f16) xin.maxbits = tc_slot(atcs, N, 1 + i);
      // d_out carries no recorded maxima: the dgrad call above reduced them into slot 0 of tcs (conv5x5_tc_forward), and no
      // later call of this function touches that slot -- the weight gradient reuses them instead of reducing d_out again
      const uint32_t* wmax = (f16 && gmax == nullptr) ? tc_slot(tcs, N, 0) : gmax;
      cudaStream_t sw = st;
      OP3_TRY(wgrad_stream(st, &sw));
      OP3_TRY(conv5x5_wgrad_auto(sw, N, D, xin, cout, g, wscr, pg + L.dec_wf[i], pg + L.dec_bf[i], wmax));
    }
    g = gn;
    gmax = f16 ? gin.out_maxbits : nullptr;
    gch = 8;
  }
  // layer 1: g = dPre1 [N][8][D][D]
  dim3 grid(8, N);
  dec_l1_class_reduce_kernel<<<grid, 256, 0, st>>>(g, D, dT);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  OP3_TRY(gemm(st, N, R, NCLS * 32, dT, NCLS * 32, 0, prep + L.wsum, R, 0, dz, R, nullptr, ACT_NONE, accumulate_dz ? 1 : 0));
  if (pg) {
    // d Wsum[cls*32+co][c] += sum_n dT[n][.] z[n][c] ; d biasT += column sums of dT ; d cmap += sum_n g
    OP3_TRY(gemm_wgrad(st, NCLS * 32, R, N, dT, NCLS * 32, z, R, pg + L.wsum, R, pg + L.biasT));
    dec_l1_dcmap_kernel<<<ceil_div(8 * D * D, 64), 256, 0, st>>>(g, N, D, reinterpret_cast<float4*>(pg + L.cmap));
    OP3_COUNT_LAUNCH();
    OP3_LAUNCH_CHECK();
  }
  return OP3_OK;
}

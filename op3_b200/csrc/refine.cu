// Refinement network + latent-state helpers.
// Replaces RefinementNetwork.forward (op3/torch/op3_modules/refinement_network.py:189-219: conv x3 + ELU, AvgPool2d(D),
// FC+ELU, LSTM cell, two linear heads), the LayerNorm of the lambda gradients and extra_input assembly of
// OP3Model.refine (op3_model.py:232-249), get_samples / _softplus_to_std (:130-153) and their backward.
#include "layout.cuh"
#include "fused_mlp.cuh"

namespace op3 {

namespace {

constexpr int LSTM = OP3_LSTM, HID = OP3_HID;

__device__ __forceinline__ float sp_std(float x) { return sqrtf(logf(1.f + expf(fminf(x, 80.f))) + 1e-5f); }
// d std / d x for std = sqrt(softplus(min(x,80)) + 1e-5)
__device__ __forceinline__ float sp_std_grad(float x, float stdv) { return x < 80.f ? sigmoid_f(x) / (2.f * stdv) : 0.f; }

__global__ void sample_kernel(int N, int R, int Rd, int Rs, const float* __restrict__ l1, const float* __restrict__ l2,
                              const float* __restrict__ eps, const float* __restrict__ deter_src, float* __restrict__ z) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (deter_src && i < N * Rd) z[(size_t)(i / Rd) * R + (i % Rd)] = deter_src[(size_t)(i / Rd) * R + (i % Rd)];
  if (i >= N * Rs) return;
  int n = i / Rs, j = i % Rs;
  z[(size_t)n * R + Rd + j] = eps[i] * sp_std(l2[i]) + l1[i];
}

// see op3_state_bwd in op3_b200.h (no __restrict__: dprior may alias dl when the prior IS the posterior)
__global__ void state_bwd_kernel(int N, int R, int Rd, int Rs, const float* l1, const float* l2, const float* eps,
                                 const float* prior1, const float* prior2, const float* dz, float w_kl, float* dl1,
                                 float* dl2, float* dprior1, float* dprior2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Rs) return;
  int n = i / Rs, j = i % Rs;
  float s2 = sp_std(l2[i]);
  float ds2 = sp_std_grad(l2[i], s2);
  float dsmp = dz ? dz[(size_t)n * R + Rd + j] : 0.f;
  float g1 = dsmp, g2 = dsmp * eps[i] * ds2;
  if (w_kl != 0.f) {
    float p2 = prior2[i];
    float s1 = sp_std(p2);
    float dm = l1[i] - prior1[i];
    float inv1 = 1.f / (s1 * s1);
    g1 += w_kl * dm * inv1;
    g2 += w_kl * (-1.f / s2 + s2 * inv1) * ds2;
    if (dprior1) {
      float a = dprior1[i] + w_kl * (-dm * inv1);
      dprior1[i] = a;
      float b = dprior2[i] + w_kl * (1.f / s1 - (s2 * s2 + dm * dm) * inv1 / s1) * sp_std_grad(p2, s1);
      dprior2[i] = b;
    }
  }
  dl1[i] += g1;
  dl2[i] += g2;
}

// One warp per slot row.  Builds columns [128, 448) of the LSTM input X: LN(g_l1) | LN(g_l2) | l1 | l2 | samples
// (op3_model.py:232-248) where g_l = d total / d lambda (decoder path through samples + beta/B * dKL).
__global__ void __launch_bounds__(128) refine_extras_kernel(int N, int R, int Rd, int Rs, Op3RefineIO io, float w_kl,
                                                            const float* __restrict__ sc1, const float* __restrict__ ce1,
                                                            const float* __restrict__ sc2, const float* __restrict__ ce2,
                                                            float* __restrict__ X, int ldx, float* __restrict__ xn) {
  const int lane = threadIdx.x % 32;
  const int n = blockIdx.x * 4 + threadIdx.x / 32;
  if (n >= N) return;
  float g1[2], g2[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    int j = lane + 32 * t;
    size_t i = (size_t)n * Rs + j;
    float l2v = io.l2[i], l1v = io.l1[i];
    float s2 = sp_std(l2v), ds2 = sp_std_grad(l2v, s2);
    float dsmp = io.dz[(size_t)n * R + Rd + j];
    float a = dsmp, b = dsmp * io.eps[i] * ds2;
    if (!io.prior_is_post && w_kl != 0.f) {
      float s1 = sp_std(io.prior2[i]);
      float dm = l1v - io.prior1[i];
      a += w_kl * dm / (s1 * s1);
      b += w_kl * (-1.f / s2 + s2 / (s1 * s1)) * ds2;
    }
    g1[t] = a; g2[t] = b;
    X[(size_t)n * ldx + HID + 2 * Rs + j] = l1v;
    X[(size_t)n * ldx + HID + 3 * Rs + j] = l2v;
    X[(size_t)n * ldx + HID + 4 * Rs + j] = io.z[(size_t)n * R + Rd + j];
  }
  // LayerNorm over Rs=64 (modules.py:19-46): mean, unbiased std, (x-mean)/(std+1e-6)*scale+center
  float m1 = warp_sum(g1[0] + g1[1]) / 64.f, m2 = warp_sum(g2[0] + g2[1]) / 64.f;
  float v1 = warp_sum((g1[0] - m1) * (g1[0] - m1) + (g1[1] - m1) * (g1[1] - m1)) / 63.f;
  float v2 = warp_sum((g2[0] - m2) * (g2[0] - m2) + (g2[1] - m2) * (g2[1] - m2)) / 63.f;
  float i1 = 1.f / (sqrtf(v1) + 1e-6f), i2 = 1.f / (sqrtf(v2) + 1e-6f);
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    int j = lane + 32 * t;
    float a = (g1[t] - m1) * i1, b = (g2[t] - m2) * i2;
    xn[(size_t)n * 2 * Rs + j] = a;
    xn[(size_t)n * 2 * Rs + Rs + j] = b;
    X[(size_t)n * ldx + HID + j] = a * sc1[j] + ce1[j];
    X[(size_t)n * ldx + HID + Rs + j] = b * sc2[j] + ce2[j];
  }
}

// pooled[n][co] = mean_{y,x} r3[n][co/4][y][x][co%4]; one CTA per (chunk, n); fixed-order tree
__global__ void __launch_bounds__(256) avgpool_kernel(const float4* __restrict__ r3, int chunks, int DD,
                                                      float* __restrict__ pooled) {
  __shared__ float4 red[256];
  const int c4 = blockIdx.x, n = blockIdx.y;
  const float4* src = r3 + ((size_t)n * chunks + c4) * DD;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = threadIdx.x; p < DD; p += 256) {
    float4 v = src[p];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      float4 a = red[threadIdx.x], b = red[threadIdx.x + o];
      red[threadIdx.x] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float inv = 1.f / (float)DD;
    float4 a = red[0];
    *reinterpret_cast<float4*>(pooled + (size_t)n * chunks * 4 + c4 * 4) = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  }
}

// g3[n][c4][p] = dpooled[n][c4*4..] / DD * ELU'(r3)
// One thread owns one pixel of one image and streams the channel chunks through it.  maxbits (optional): per-image largest
// magnitude of g3 for the fp16 conv mode (zeroed by the caller): one conditional atomic per block.
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dpooled, const float4* __restrict__ r3,
                                                          int chunks, int DD, float4* __restrict__ g3,
                                                          uint32_t* __restrict__ maxbits) {
  const int n = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const float inv = 1.f / (float)DD;
  float mx = 0.f;
  if (p < DD) {
#pragma unroll 4
    for (int c4 = 0; c4 < chunks; ++c4) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(dpooled + (size_t)n * chunks * 4 + c4 * 4));
      const size_t o = ((size_t)n * chunks + c4) * DD + p;
      const float4 a = r3[o];
      const float4 v = make_float4(d.x * inv * elu_grad_from_out(a.x), d.y * inv * elu_grad_from_out(a.y),
                                   d.z * inv * elu_grad_from_out(a.z), d.w * inv * elu_grad_from_out(a.w));
      g3[o] = v;
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  }
  if (maxbits != nullptr) {
    __shared__ float s_mx[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_mx[w]);
      const uint32_t b = __float_as_uint(mx);
      if (b > __ldcg(&maxbits[n])) atomicMax(&maxbits[n], b);
    }
  }
}

__global__ void lstm_fwd_kernel(int N, const float* __restrict__ gates, const float* __restrict__ c, float* __restrict__ h2,
                                float* __restrict__ c2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * LSTM) return;
  int n = i / LSTM, j = i % LSTM;
  const float* g = gates + (size_t)n * 4 * LSTM;
  float gi = sigmoid_f(g[j]), gf = sigmoid_f(g[LSTM + j]), gg = tanhf(g[2 * LSTM + j]), go = sigmoid_f(g[3 * LSTM + j]);
  float cn = gf * c[i] + gi * gg;
  c2[i] = cn;
  h2[i] = go * tanhf(cn);
}

// dgates from (dh2, dc2); also dc_prev (+=)
__global__ void lstm_bwd_kernel(int N, const float* __restrict__ gates, const float* __restrict__ c,
                                const float* __restrict__ dh2, const float* __restrict__ dc2, float* __restrict__ dgates,
                                float* __restrict__ dc_prev) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * LSTM) return;
  int n = i / LSTM, j = i % LSTM;
  const float* g = gates + (size_t)n * 4 * LSTM;
  float gi = sigmoid_f(g[j]), gf = sigmoid_f(g[LSTM + j]), gg = tanhf(g[2 * LSTM + j]), go = sigmoid_f(g[3 * LSTM + j]);
  float cp = c[i];
  float cn = gf * cp + gi * gg;
  float tc = tanhf(cn);
  float dh = dh2 ? dh2[i] : 0.f;
  float dcn = (dc2 ? dc2[i] : 0.f) + dh * go * (1.f - tc * tc);
  float* dg = dgates + (size_t)n * 4 * LSTM;
  dg[j] = dcn * gg * gi * (1.f - gi);
  dg[LSTM + j] = dcn * cp * gf * (1.f - gf);
  dg[2 * LSTM + j] = dcn * gi * (1.f - gg * gg);
  dg[3 * LSTM + j] = dh * tc * go * (1.f - go);
  if (dc_prev) dc_prev[i] += dcn * gf;
}

// out[n] += sum_m X[m][n] * Y[m][n]   (Y == nullptr -> plain column sum); deterministic
__global__ void __launch_bounds__(256) colsum_prod_kernel(int M, int N, const float* __restrict__ X, int ldx,
                                                          const float* __restrict__ Y, int ldy, float* __restrict__ out) {
  __shared__ float part[8][33];
  int lane = threadIdx.x % 32, w = threadIdx.x / 32;
  int n = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (n < N)
    for (int m = w; m < M; m += 8) s += X[(size_t)m * ldx + n] * (Y ? Y[(size_t)m * ldy + n] : 1.f);
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][lane];
    out[n] += t;
  }
}

// dst[m][0..n) (+)= src[m][0..n)  with row strides
__global__ void add_cols_kernel(int M, int N, const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  int m = i / N, n = i % N;
  dst[(size_t)m * ldd + n] += src[(size_t)m * lds + n];
}

struct RefActs {
  float4 *r1, *r2, *r3;
  float *pooled, *X, *gates, *xn;
};
inline RefActs ref_acts(const Op3Dims* d, float* base) {
  RefActs A;
  const size_t N = (size_t)d->B * d->K, DD = (size_t)d->D * d->D;
  size_t o = 0;
  A.r1 = reinterpret_cast<float4*>(base + o); o += N * 32 * DD;
  A.r2 = reinterpret_cast<float4*>(base + o); o += N * 32 * DD;
  A.r3 = reinterpret_cast<float4*>(base + o); o += N * d->Rs * DD;
  A.pooled = base + o; o += N * d->Rs;
  A.X = base + o; o += N * (HID + 5 * d->Rs);
  A.gates = base + o; o += N * 4 * LSTM;
  A.xn = base + o; o += N * 2 * d->Rs;
  return A;
}
inline size_t ref_acts_floats(const Op3Dims* d) {
  const size_t N = (size_t)d->B * d->K, DD = (size_t)d->D * d->D;
  return N * (64 + d->Rs) * DD + N * (d->Rs + HID + 5 * d->Rs + 4 * LSTM + 2 * d->Rs);
}
inline size_t ref_acts_total(const Op3Dims* d) { return ((ref_acts_floats(d) + 3) & ~(size_t)3) + tc_scratch_floats(d->B * d->K); }
inline float* ref_acts_tcs(const Op3Dims* d, float* base) { return base + ((ref_acts_floats(d) + 3) & ~(size_t)3); }

inline ConvInput refine_l1_input(const Op3Dims* d, const Op3RefineIO* io, const float* prep, const PrepLayout& L) {
  ConvInput in{};
  in.src[0] = {reinterpret_cast<const float4*>(io->dec_out), 1, 1};
  in.src[1] = {reinterpret_cast<const float4*>(io->mix1), 1, 1};
  in.src[2] = {reinterpret_cast<const float4*>(io->mix2), 1, 1};
  in.src[3] = {reinterpret_cast<const float4*>(io->smp), 1, d->K};
  in.src[4] = {reinterpret_cast<const float4*>(prep + L.coords), 1, 0};
  in.nsrc = 5; in.total_chunks = 5; in.real_cin = 17;
  return in;
}
inline ConvInput one_src(const float4* p, int chunks) {
  ConvInput in{};
  in.src[0] = {p, chunks, 1};
  in.nsrc = 1; in.total_chunks = chunks;
  return in;
}


// ===================================================================================================================
// Fused head of the refinement network (refinement_network.py:203-219 + the extra_input assembly of op3_model.py:232-249):
// one CTA owns a tile of slot rows and runs FC+ELU, the LayerNorm'd lambda gradients, both LSTM gate products, the cell and
// the two linear heads with the row tile in shared memory (fused_mlp.cuh).
// ===================================================================================================================
using namespace fused;
constexpr int XW_ = HID + 5 * 64;        // 448 columns of the LSTM input (Rs = 64)
constexpr int LPO = PITCH(64), LXW = PITCH(XW_), LHH = PITCH(128), LG = PITCH(512);   // panel pitches (fused_mlp.cuh)
constexpr int REF_FWD_SMEM_FLOATS = GEMM_SMEM_FLOATS + 16 * (LPO + LXW + LHH + LG + LHH);
constexpr int REF_BWD_SMEM_FLOATS = GEMM_SMEM_FLOATS + 16 * (LPO + LHH + LG + LXW);

// extra == nullptr: columns [128, 448) of X are built from the latent state (op3_refine_fwd); else copied from extra [N][320]
__global__ void __launch_bounds__(FT, 1) ref_head_fwd_kernel(int N, int R, int Rd, const __grid_constant__ Op3Params p,
                                                          const __grid_constant__ Op3RefineIO io, float w_kl,
                                                          const float* __restrict__ extra, const float* __restrict__ pooled,
                                                          float* __restrict__ X, float* __restrict__ gates,
                                                          float* __restrict__ xn, int rpc, int tc) {
  extern __shared__ __align__(16) float sm[];
  float* wt = sm;
  float* PPOOL = wt + GEMM_SMEM_FLOATS;   // [16][LPO]
  float* PX = PPOOL + 16 * LPO;           // [16][LXW]
  float* PHP = PX + 16 * LXW;             // [16][LHH]  previous h
  float* PG = PHP + 16 * LHH;             // [16][LG]   gate pre-activations
  float* PH2 = PG + 16 * LG;              // [16][LHH]  new h
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int Rs = 64;
  const int row0 = blockIdx.x * rpc, rows = min(rpc, N - row0);
  zero_panel(PPOOL, REF_FWD_SMEM_FLOATS - GEMM_SMEM_FLOATS);   // rows the GEMMs compute but nobody owns must stay finite
  __syncthreads();
  load_panel(PPOOL, LPO, pooled + (size_t)row0 * Rs, Rs, rows, Rs);
  load_panel(PHP, LHH, io.h + (size_t)row0 * LSTM, LSTM, rows, LSTM);
  if (extra != nullptr) {
    for (int e = tid; e < rows * 5 * Rs; e += FT) {
      const int r = e / (5 * Rs), c = e - r * 5 * Rs;
      const float v = extra[(size_t)(row0 + r) * 5 * Rs + c];
      PX[r * LXW + HID + c] = v;
      X[(size_t)(row0 + r) * XW_ + HID + c] = v;
    }
  } else {
    const float *sc1 = p.ln1d_scale[0], *ce1 = p.ln1d_center[0], *sc2 = p.ln1d_scale[1], *ce2 = p.ln1d_center[1];
    for (int r = warp; r < rows; r += FT / 32) {             // one warp per slot row: see refine_extras_kernel
      const int n = row0 + r;
      float g1[2], g2[2];
      float* xr = PX + r * LXW + HID;
      float* xg = X + (size_t)n * XW_ + HID;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        const size_t i = (size_t)n * Rs + j;
        const float l2v = io.l2[i], l1v = io.l1[i];
        const float s2 = sp_std(l2v), ds2 = sp_std_grad(l2v, s2);
        const float dsmp = io.dz[(size_t)n * R + Rd + j];
        float a = dsmp, b = dsmp * io.eps[i] * ds2;
        if (!io.prior_is_post && w_kl != 0.f) {
          const float s1 = sp_std(io.prior2[i]);
          const float dm = l1v - io.prior1[i];
          a += w_kl * dm / (s1 * s1);
          b += w_kl * (-1.f / s2 + s2 / (s1 * s1)) * ds2;
        }
        g1[t] = a; g2[t] = b;
        const float zs = io.z[(size_t)n * R + Rd + j];
        xr[2 * Rs + j] = l1v; xr[3 * Rs + j] = l2v; xr[4 * Rs + j] = zs;
        xg[2 * Rs + j] = l1v; xg[3 * Rs + j] = l2v; xg[4 * Rs + j] = zs;
      }
      const float m1 = warp_sum(g1[0] + g1[1]) / 64.f, m2 = warp_sum(g2[0] + g2[1]) / 64.f;
      const float v1 = warp_sum((g1[0] - m1) * (g1[0] - m1) + (g1[1] - m1) * (g1[1] - m1)) / 63.f;
      const float v2 = warp_sum((g2[0] - m2) * (g2[0] - m2) + (g2[1] - m2) * (g2[1] - m2)) / 63.f;
      const float i1 = 1.f / (sqrtf(v1) + 1e-6f), i2 = 1.f / (sqrtf(v2) + 1e-6f);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        const float a = (g1[t] - m1) * i1, b = (g2[t] - m2) * i2;
        xn[(size_t)n * 2 * Rs + j] = a;
        xn[(size_t)n * 2 * Rs + Rs + j] = b;
        const float ya = a * sc1[j] + ce1[j], yb = b * sc2[j] + ce2[j];
        xr[j] = ya; xr[Rs + j] = yb;
        xg[j] = ya; xg[Rs + j] = yb;
      }
    }
  }
  __syncthreads();
  cta_gemm<false>(tc, PPOOL, LPO, rows, p.ref_fc.w, Rs, Rs, HID, wt, epi_fwd(p.ref_fc.b, ACT_ELU, PX, LXW, X + (size_t)row0 * XW_, XW_));
  cta_gemm<false>(tc, PX, LXW, rows, p.lstm_w_ih, XW_, XW_, 4 * LSTM, wt, epi_fwd(p.lstm_b_ih, ACT_NONE, PG, LG, nullptr, 0));
  {
    Epi e = epi_fwd(p.lstm_b_hh, ACT_NONE, PG, LG, gates + (size_t)row0 * 4 * LSTM, 4 * LSTM);
    e.so_acc = 1;
    cta_gemm<false>(tc, PHP, LHH, rows, p.lstm_w_hh, LSTM, LSTM, 4 * LSTM, wt, e);
  }
  for (int e = tid; e < rows * LSTM; e += FT) {               // LSTM cell, gate order i,f,g,o
    const int r = e / LSTM, j = e - r * LSTM;
    const float* g = PG + r * LG;
    const float gi = sigmoid_f(g[j]), gf = sigmoid_f(g[LSTM + j]), gg = tanhf(g[2 * LSTM + j]), go = sigmoid_f(g[3 * LSTM + j]);
    const size_t i = (size_t)(row0 + r) * LSTM + j;
    const float cn = gf * io.c[i] + gi * gg;
    const float hn = go * tanhf(cn);
    io.c_out[i] = cn;
    io.h_out[i] = hn;
    PH2[r * LHH + j] = hn;
  }
  __syncthreads();
  cta_gemm<false>(tc, PH2, LHH, rows, p.ref_last_fc.w, LSTM, LSTM, Rs, wt,
                  epi_fwd(p.ref_last_fc.b, ACT_NONE, nullptr, 0, io.l1_out + (size_t)row0 * Rs, Rs));
  cta_gemm<false>(tc, PH2, LHH, rows, p.ref_last_fc2.w, LSTM, LSTM, Rs, wt,
                  epi_fwd(p.ref_last_fc2.b, ACT_NONE, nullptr, 0, io.l2_out + (size_t)row0 * Rs, Rs));
}

// Backward of the head for a row tile.  Writes dgates and dX (columns [0,128) already multiplied by ELU' of the FC output:
// the operands of the batched weight-gradient launch), dpool; adds the direct paths into dl1, dl2, dz_acc, dh, dc.
__global__ void __launch_bounds__(FT, 1) ref_head_bwd_kernel(int N, int R, int Rd, const __grid_constant__ Op3Params p,
                                                          const float* __restrict__ gates, const float* __restrict__ Xa,
                                                          const float* __restrict__ c_prev, const float* __restrict__ d_l1n,
                                                          const float* __restrict__ d_l2n, const float* __restrict__ d_h,
                                                          const float* __restrict__ d_c, float* __restrict__ dgates,
                                                          float* __restrict__ dX, float* __restrict__ dpool, float* dl1,
                                                          float* dl2, float* dz_acc, float* dh, float* dc, int rpc, int tc) {
  extern __shared__ __align__(16) float sm[];
  float* wt = sm;
  float* PU = wt + GEMM_SMEM_FLOATS;      // [16][LPO]
  float* PDH2 = PU + 16 * LPO;            // [16][LHH]
  float* PDG = PDH2 + 16 * LHH;           // [16][LG]
  float* PDX = PDG + 16 * LG;             // [16][LXW]
  const int tid = threadIdx.x;
  constexpr int Rs = 64;
  const int row0 = blockIdx.x * rpc, rows = min(rpc, N - row0);
  zero_panel(PU, REF_BWD_SMEM_FLOATS - GEMM_SMEM_FLOATS);      // rows the GEMMs compute but nobody owns must stay finite
  __syncthreads();
  for (int e = tid; e < rows * LSTM; e += FT) PDH2[(e / LSTM) * LHH + (e % LSTM)] = d_h ? d_h[(size_t)row0 * LSTM + e] : 0.f;
  for (int hd = 0; hd < 2; ++hd) {
    const float* up = hd == 0 ? d_l1n : d_l2n;
    if (up == nullptr) continue;
    __syncthreads();
    load_panel(PU, LPO, up + (size_t)row0 * Rs, Rs, rows, Rs);
    __syncthreads();
    cta_gemm<true>(tc, PU, LPO, rows, hd == 0 ? p.ref_last_fc.w : p.ref_last_fc2.w, LSTM, Rs, LSTM, wt,
                   epi_bwd(nullptr, 0, PDH2, LHH, 1, nullptr, 0, 0));
  }
  __syncthreads();
  for (int e = tid; e < rows * LSTM; e += FT) {               // LSTM cell backward
    const int r = e / LSTM, j = e - r * LSTM;
    const size_t i = (size_t)(row0 + r) * LSTM + j;
    const float* g = gates + (size_t)(row0 + r) * 4 * LSTM;
    const float gi = sigmoid_f(g[j]), gf = sigmoid_f(g[LSTM + j]), gg = tanhf(g[2 * LSTM + j]), go = sigmoid_f(g[3 * LSTM + j]);
    const float cp = c_prev[i];
    const float cn = gf * cp + gi * gg;
    const float tc = tanhf(cn);
    const float dhv = PDH2[r * LHH + j];
    const float dcn = (d_c ? d_c[i] : 0.f) + dhv * go * (1.f - tc * tc);
    const float a0 = dcn * gg * gi * (1.f - gi), a1 = dcn * cp * gf * (1.f - gf), a2 = dcn * gi * (1.f - gg * gg),
                a3 = dhv * tc * go * (1.f - go);
    float* pg = PDG + r * LG;
    float* dg = dgates + (size_t)(row0 + r) * 4 * LSTM;
    pg[j] = a0; pg[LSTM + j] = a1; pg[2 * LSTM + j] = a2; pg[3 * LSTM + j] = a3;
    dg[j] = a0; dg[LSTM + j] = a1; dg[2 * LSTM + j] = a2; dg[3 * LSTM + j] = a3;
    dc[i] += dcn * gf;
  }
  __syncthreads();
  {
    Epi e = epi_bwd(Xa + (size_t)row0 * XW_, XW_, PDX, LXW, 0, dX + (size_t)row0 * XW_, XW_, 0);
    e.gy_cols = HID;                      // ELU' of the FC output for its 128 columns only
    cta_gemm<true>(tc, PDG, LG, rows, p.lstm_w_ih, XW_, 4 * LSTM, XW_, wt, e);
  }
  for (int e = tid; e < rows * 3 * Rs; e += FT) {   // direct paths into the previous state's lambdas / samples (op3_model.py:244-248)
    const int r = e / (3 * Rs), c = e - r * 3 * Rs;
    const float v = PDX[r * LXW + HID + 2 * Rs + c];
    if (c < Rs) dl1[(size_t)(row0 + r) * Rs + c] += v;
    else if (c < 2 * Rs) dl2[(size_t)(row0 + r) * Rs + c - Rs] += v;
    else dz_acc[(size_t)(row0 + r) * R + Rd + c - 2 * Rs] += v;
  }
  cta_gemm<true>(tc, PDG, LG, rows, p.lstm_w_hh, LSTM, 4 * LSTM, LSTM, wt,
                 epi_bwd(nullptr, 0, nullptr, 0, 0, dh + (size_t)row0 * LSTM, LSTM, 1));
  cta_gemm<true>(tc, PDX, LXW, rows, p.ref_fc.w, Rs, HID, Rs, wt, epi_bwd(nullptr, 0, nullptr, 0, 0, dpool + (size_t)row0 * Rs, Rs, 0));
}

// LayerNorm (modules.py:19-46) parameter gradients of the two lambda-gradient inputs in one launch:
// scale[j][n] += sum_m dX[m][128 + 64 j + n] * xn[m][64 j + n]; center[j][n] += sum_m dX[m][128 + 64 j + n].  Deterministic.
__global__ void __launch_bounds__(256) ln1d_grad_kernel(int M, const float* __restrict__ dX, const float* __restrict__ xn,
                                                        float* s0, float* c0, float* s1, float* c1) {
  __shared__ float part[2][8][33];
  const int lane = threadIdx.x % 32, w = threadIdx.x / 32;
  const int j = blockIdx.x >> 1, n = (blockIdx.x & 1) * 32 + lane;
  float a = 0.f, b = 0.f;
  for (int m = w; m < M; m += 8) {
    const float d = dX[(size_t)m * XW_ + HID + j * 64 + n];
    a += d * xn[(size_t)m * 128 + j * 64 + n];
    b += d;
  }
  part[0][w][lane] = a; part[1][w][lane] = b;
  __syncthreads();
  if (w < 2) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[w][i][lane];
    float* out = w == 0 ? (j == 0 ? s0 : s1) : (j == 0 ? c0 : c1);
    out[n] += t;
  }
}

inline bool ref_head_fusable(const Op3Params* p) {
  return aligned16(p->ref_fc.w) && aligned16(p->lstm_w_ih) && aligned16(p->lstm_w_hh) && aligned16(p->ref_last_fc.w) &&
         aligned16(p->ref_last_fc2.w);
}
inline int head_rows_per_cta(int N) {
  int r = ((ceil_div(N, device_sm_count()) + 3) / 4) * 4;
  return r < 4 ? 4 : (r > RTM ? RTM : r);
}
inline double ref_head_flops(int N) { return 2.0 * N * (64.0 * HID + (double)XW_ * 512 + 128.0 * 512 + 2 * 128.0 * 64); }

int ref_head_fwd_fused(cudaStream_t st, const Op3Dims* d, const Op3Params* p, const Op3RefineIO& io, const float* extra,
                       const RefActs& A) {
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    OP3_CHECK(cudaFuncSetAttribute(ref_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   REF_FWD_SMEM_FLOATS * (int)sizeof(float)));
  const int N = d->B * d->K, rpc = head_rows_per_cta(N);
  ProfScope prof(st, KID_GEMM, ref_head_flops(N));
  ref_head_fwd_kernel<<<ceil_div(N, rpc), FT, REF_FWD_SMEM_FLOATS * sizeof(float), st>>>(
      N, d->Rd + d->Rs, d->Rd, *p, io, d->beta / (float)d->B, extra, A.pooled, A.X, A.gates, A.xn, rpc,
      op3_get_precision() != OP3_PREC_FP32);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

}  // namespace
}  // namespace op3

using namespace op3;

extern "C" int op3_sample_fwd(const Op3Dims* d, int N, const float* l1, const float* l2, const float* eps,
                              const float* deter_src, float* z, void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(l1 && l2 && eps && z && N > 0, "op3_sample_fwd: bad argument");
  int tot = N * (d->Rs > d->Rd ? d->Rs : d->Rd);
  sample_kernel<<<ceil_div(tot, 256), 256, 0, (cudaStream_t)stream>>>(N, d->Rd + d->Rs, d->Rd, d->Rs, l1, l2, eps, deter_src, z);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

extern "C" int op3_state_bwd(const Op3Dims* d, int N, const float* l1, const float* l2, const float* eps,
                             const float* prior1, const float* prior2, const float* dz, float w_kl, float* dl1,
                             float* dl2, float* dprior1, float* dprior2, void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(l1 && l2 && eps && dl1 && dl2 && N > 0, "op3_state_bwd: bad argument");
  OP3_REQUIRE(w_kl == 0.f || (prior1 && prior2), "op3_state_bwd: prior required when w_kl != 0");
  OP3_REQUIRE((dprior1 == nullptr) == (dprior2 == nullptr), "op3_state_bwd: dprior1/dprior2 must both be given");
  state_bwd_kernel<<<ceil_div(N * d->Rs, 256), 256, 0, (cudaStream_t)stream>>>(N, d->Rd + d->Rs, d->Rd, d->Rs, l1, l2, eps,
                                                                            prior1, prior2, dz, w_kl, dl1, dl2, dprior1,
                                                                            dprior2);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

extern "C" size_t op3_refine_acts_floats(const Op3Dims* d) {
  if (validate_dims(d) != OP3_OK) return 0;
  return ref_acts_total(d);
}

extern "C" size_t op3_refine_scratch_floats(const Op3Dims* d) {
  if (validate_dims(d) != OP3_OK) return 0;
  const size_t N = (size_t)d->B * d->K, DD = (size_t)d->D * d->D;
  size_t w = conv5x5_wgrad_scratch_floats(32, d->Rs);
  size_t w2 = conv5x5_wgrad_scratch_floats(32, 32);
  if (w2 > w) w = w2;
  return N * (d->Rs + 64) * DD + N * (4 * LSTM + HID + 5 * d->Rs + LSTM + HID + d->Rs) + ((w + 3) & ~(size_t)3) +
         tc_scratch_floats((int)N);
}

extern "C" int op3_refine_fwd(const Op3Dims* d, const Op3Params* p, const float* prep, const Op3RefineIO* io,
                              void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(p && prep && io && io->dec_out && io->mix1 && io->mix2 && io->smp && io->dz && io->z && io->l1 && io->l2 &&
                  io->eps && io->prior1 && io->prior2 && io->h && io->c && io->acts && io->l1_out && io->l2_out && io->h_out && io->c_out,
              "op3_refine_fwd: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = d->B * d->K, D = d->D, DD = D * D, Rs = d->Rs, R = d->Rd + d->Rs;
  const int XW = HID + 5 * Rs;
  PrepLayout L = prep_layout(d);
  RefActs A = ref_acts(d, io->acts);
  float* tcs = ref_acts_tcs(d, io->acts);
  // conv stack 17(20) -> 32 -> 32 -> Rs, ELU each (refinement_network.py:195)
  {
    // fp16 mode: slots 1..3 of tcs keep the per-image maxima of r1..r3 (recorded by their producers, reused by backward)
    const bool f16 = op3_get_precision() == OP3_PREC_F16X3;
    if (f16) OP3_CHECK(cudaMemsetAsync(tcs, 0, tc_scratch_floats(N) * sizeof(float), st));
    ConvInput i1 = refine_l1_input(d, io, prep, L), i2 = one_src(A.r1, 8), i3 = one_src(A.r2, 8);
    if (f16) {
      i1.out_maxbits = tc_slot(tcs, N, 1);
      i2.maxbits = tc_slot(tcs, N, 1); i2.out_maxbits = tc_slot(tcs, N, 2);
      i3.maxbits = tc_slot(tcs, N, 2); i3.out_maxbits = tc_slot(tcs, N, 3);
    }
    OP3_TRY(conv5x5_auto(st, N, D, i1, 32, prep + L.ref_wf[0], prep + L.ref_tf[0], prep + L.ref_bf[0], A.r1, EPI_ELU, nullptr, tcs));
    OP3_TRY(conv5x5_auto(st, N, D, i2, 32, prep + L.ref_wf[1], prep + L.ref_tf[1], prep + L.ref_bf[1], A.r2, EPI_ELU, nullptr, tcs));
    OP3_TRY(conv5x5_auto(st, N, D, i3, Rs, prep + L.ref_wf[2], prep + L.ref_tf[2], prep + L.ref_bf[2], A.r3, EPI_ELU, nullptr, tcs));
  }
  avgpool_kernel<<<dim3(Rs / 4, N), 256, 0, st>>>(A.r3, Rs / 4, DD, A.pooled);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  if (fused_mlp_enabled() && ref_head_fusable(p)) return ref_head_fwd_fused(st, d, p, *io, nullptr, A);
  // X = [ ELU(fc(pooled)) | LN(g_l1) | LN(g_l2) | l1 | l2 | samples ]
  OP3_TRY(gemm(st, N, HID, Rs, A.pooled, Rs, 0, p->ref_fc.w, Rs, 1, A.X, XW, p->ref_fc.b, ACT_ELU, 0));
  refine_extras_kernel<<<ceil_div(N, 4), 128, 0, st>>>(N, R, d->Rd, Rs, *io, d->beta / (float)d->B, p->ln1d_scale[0],
                                                     p->ln1d_center[0], p->ln1d_scale[1], p->ln1d_center[1], A.X, XW, A.xn);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  // LSTM cell, gate order i,f,g,o
  OP3_TRY(gemm(st, N, 4 * LSTM, XW, A.X, XW, 0, p->lstm_w_ih, XW, 1, A.gates, 4 * LSTM, p->lstm_b_ih, ACT_NONE, 0));
  OP3_TRY(gemm(st, N, 4 * LSTM, LSTM, io->h, LSTM, 0, p->lstm_w_hh, LSTM, 1, A.gates, 4 * LSTM, p->lstm_b_hh, ACT_NONE, 1));
  lstm_fwd_kernel<<<ceil_div(N * LSTM, 256), 256, 0, st>>>(N, A.gates, io->c, io->h_out, io->c_out);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  OP3_TRY(gemm(st, N, Rs, LSTM, io->h_out, LSTM, 0, p->ref_last_fc.w, LSTM, 1, io->l1_out, Rs, p->ref_last_fc.b, ACT_NONE, 0));
  OP3_TRY(gemm(st, N, Rs, LSTM, io->h_out, LSTM, 0, p->ref_last_fc2.w, LSTM, 1, io->l2_out, Rs, p->ref_last_fc2.b, ACT_NONE, 0));
  return OP3_OK;
}

// Standalone RefinementNetwork.forward (refinement_network.py:189-219): the caller supplies the 17 conv input channels
// already packed as 5 float4 chunks in this library's channel order (see ref_l1_channel) and the 5*Rs extra inputs.
extern "C" int op3_refine_net_fwd(const Op3Dims* d, const Op3Params* p, const float* prep, const float* in_chunks,
                                  const float* extra, const float* h, const float* c, float* acts, float* l1_out,
                                  float* l2_out, float* h_out, float* c_out, void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(p && prep && in_chunks && extra && h && c && acts && l1_out && l2_out && h_out && c_out,
              "op3_refine_net_fwd: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = d->B * d->K, D = d->D, DD = D * D, Rs = d->Rs;
  const int XW = HID + 5 * Rs;
  PrepLayout L = prep_layout(d);
  RefActs A = ref_acts(d, acts);
  float* tcs = ref_acts_tcs(d, acts);
  ConvInput in = one_src(reinterpret_cast<const float4*>(in_chunks), 5);
  in.real_cin = 17;
  {
    // fp16 mode: slots 1..3 of tcs keep the per-image maxima of r1..r3 (recorded by their producers, reused by backward)
    const bool f16 = op3_get_precision() == OP3_PREC_F16X3;
    if (f16) OP3_CHECK(cudaMemsetAsync(tcs, 0, tc_scratch_floats(N) * sizeof(float), st));
    ConvInput i1 = in, i2 = one_src(A.r1, 8), i3 = one_src(A.r2, 8);
    if (f16) {
      i1.out_maxbits = tc_slot(tcs, N, 1);
      i2.maxbits = tc_slot(tcs, N, 1); i2.out_maxbits = tc_slot(tcs, N, 2);
      i3.maxbits = tc_slot(tcs, N, 2); i3.out_maxbits = tc_slot(tcs, N, 3);
    }
    OP3_TRY(conv5x5_auto(st, N, D, i1, 32, prep + L.ref_wf[0], prep + L.ref_tf[0], prep + L.ref_bf[0], A.r1, EPI_ELU, nullptr, tcs));
    OP3_TRY(conv5x5_auto(st, N, D, i2, 32, prep + L.ref_wf[1], prep + L.ref_tf[1], prep + L.ref_bf[1], A.r2, EPI_ELU, nullptr, tcs));
    OP3_TRY(conv5x5_auto(st, N, D, i3, Rs, prep + L.ref_wf[2], prep + L.ref_tf[2], prep + L.ref_bf[2], A.r3, EPI_ELU, nullptr, tcs));
  }
  avgpool_kernel<<<dim3(Rs / 4, N), 256, 0, st>>>(A.r3, Rs / 4, DD, A.pooled);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  if (fused_mlp_enabled() && ref_head_fusable(p) && aligned16(h)) {
    Op3RefineIO io{};
    io.h = h; io.c = c; io.l1_out = l1_out; io.l2_out = l2_out; io.h_out = h_out; io.c_out = c_out;
    return ref_head_fwd_fused(st, d, p, io, extra, A);
  }
  OP3_TRY(gemm(st, N, HID, Rs, A.pooled, Rs, 0, p->ref_fc.w, Rs, 1, A.X, XW, p->ref_fc.b, ACT_ELU, 0));
  OP3_CHECK(cudaMemcpy2DAsync(A.X + HID, (size_t)XW * sizeof(float), extra, (size_t)5 * Rs * sizeof(float),
                              (size_t)5 * Rs * sizeof(float), N, cudaMemcpyDeviceToDevice, st));
  OP3_TRY(gemm(st, N, 4 * LSTM, XW, A.X, XW, 0, p->lstm_w_ih, XW, 1, A.gates, 4 * LSTM, p->lstm_b_ih, ACT_NONE, 0));
  OP3_TRY(gemm(st, N, 4 * LSTM, LSTM, h, LSTM, 0, p->lstm_w_hh, LSTM, 1, A.gates, 4 * LSTM, p->lstm_b_hh, ACT_NONE, 1));
  lstm_fwd_kernel<<<ceil_div(N * LSTM, 256), 256, 0, st>>>(N, A.gates, c, h_out, c_out);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  OP3_TRY(gemm(st, N, Rs, LSTM, h_out, LSTM, 0, p->ref_last_fc.w, LSTM, 1, l1_out, Rs, p->ref_last_fc.b, ACT_NONE, 0));
  OP3_TRY(gemm(st, N, Rs, LSTM, h_out, LSTM, 0, p->ref_last_fc2.w, LSTM, 1, l2_out, Rs, p->ref_last_fc2.b, ACT_NONE, 0));
  return OP3_OK;
}

extern "C" int op3_refine_bwd(const Op3Dims* d, const Op3Params* p, const float* prep, const Op3RefineIO* io,
                              const float* d_l1n, const float* d_l2n, const float* d_h, const float* d_c, float* d_in,
                              float* dl1, float* dl2,
                              float* dz_acc, float* dh, float* dc, const Op3Params* g, float* pg, float* scratch,
                              void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(p && prep && io && io->acts && io->h && io->c && io->h_out && d_in && dl1 && dl2 && dz_acc && dh && dc && g &&
                  pg && scratch,
              "op3_refine_bwd: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = d->B * d->K, D = d->D, DD = D * D, Rs = d->Rs, R = d->Rd + d->Rs;
  const int XW = HID + 5 * Rs;
  PrepLayout L = prep_layout(d);
  RefActs A = ref_acts(d, io->acts);
  // scratch carve-up
  size_t o = 0;
  float4* g3 = reinterpret_cast<float4*>(scratch + o); o += (size_t)N * Rs * DD;
  float4* g2 = reinterpret_cast<float4*>(scratch + o); o += (size_t)N * 32 * DD;
  float4* g1 = reinterpret_cast<float4*>(scratch + o); o += (size_t)N * 32 * DD;
  float* dgates = scratch + o; o += (size_t)N * 4 * LSTM;
  float* dX = scratch + o; o += (size_t)N * XW;
  float* dh2 = scratch + o; o += (size_t)N * LSTM;
  float* dfc = scratch + o; o += (size_t)N * HID;   // (unused separately: dX[:, :HID] is used in place)
  float* dpool = scratch + o; o += (size_t)N * Rs;
  float* wscr = scratch + o;
  {
    size_t w = conv5x5_wgrad_scratch_floats(32, Rs), w2 = conv5x5_wgrad_scratch_floats(32, 32);
    o += ((w > w2 ? w : w2) + 3) & ~(size_t)3;
  }
  float* tcs = scratch + o;               // fp16-mode operand maxima
  (void)dfc;

  if (fused_mlp_enabled() && ref_head_fusable(p) && (!d_l1n || aligned16(d_l1n)) && (!d_l2n || aligned16(d_l2n))) {
    static PerDeviceOnce attr_once;
    if (attr_once.first())
      OP3_CHECK(cudaFuncSetAttribute(ref_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     REF_BWD_SMEM_FLOATS * (int)sizeof(float)));
    const int rpc = head_rows_per_cta(N);
    {
      ProfScope prof(st, KID_GEMM, 2.0 * ref_head_flops(N));
      ref_head_bwd_kernel<<<ceil_div(N, rpc), FT, REF_BWD_SMEM_FLOATS * sizeof(float), st>>>(
          N, R, d->Rd, *p, A.gates, A.X, io->c, d_l1n, d_l2n, d_h, d_c, dgates, dX, dpool, dl1, dl2, dz_acc, dh, dc, rpc,
          op3_get_precision() != OP3_PREC_FP32);
      OP3_COUNT_LAUNCH();
      OP3_LAUNCH_CHECK();
    }
    WgradItem it[5];
    int n = 0;
    if (d_l1n) it[n++] = {d_l1n, io->h_out, g->ref_last_fc.w, g->ref_last_fc.b, Rs, LSTM, N, Rs, LSTM, LSTM};
    if (d_l2n) it[n++] = {d_l2n, io->h_out, g->ref_last_fc2.w, g->ref_last_fc2.b, Rs, LSTM, N, Rs, LSTM, LSTM};
    it[n++] = {dgates, A.X, g->lstm_w_ih, g->lstm_b_ih, 4 * LSTM, XW, N, 4 * LSTM, XW, XW};
    it[n++] = {dgates, io->h, g->lstm_w_hh, g->lstm_b_hh, 4 * LSTM, LSTM, N, 4 * LSTM, LSTM, LSTM};
    it[n++] = {dX, A.pooled, g->ref_fc.w, g->ref_fc.b, HID, Rs, N, XW, Rs, Rs};
    OP3_TRY(gemm_wgrad_batch(st, it, n));
    ln1d_grad_kernel<<<4, 256, 0, st>>>(N, dX, A.xn, g->ln1d_scale[0], g->ln1d_center[0], g->ln1d_scale[1], g->ln1d_center[1]);
    OP3_COUNT_LAUNCH();
    OP3_LAUNCH_CHECK();
  } else {
  // heads: lam = h2 W^T + b
  OP3_CHECK(cudaMemsetAsync(dh2, 0, (size_t)N * LSTM * sizeof(float), st));
  if (d_h) OP3_TRY(axpy(st, (long long)N * LSTM, d_h, dh2));
  if (d_l1n) {
    OP3_TRY(gemm(st, N, LSTM, Rs, d_l1n, Rs, 0, p->ref_last_fc.w, LSTM, 0, dh2, LSTM, nullptr, ACT_NONE, 1));
    OP3_TRY(gemm_wgrad(st, Rs, LSTM, N, d_l1n, Rs, io->h_out, LSTM, g->ref_last_fc.w, LSTM, g->ref_last_fc.b));
  }
  if (d_l2n) {
    OP3_TRY(gemm(st, N, LSTM, Rs, d_l2n, Rs, 0, p->ref_last_fc2.w, LSTM, 0, dh2, LSTM, nullptr, ACT_NONE, 1));
    OP3_TRY(gemm_wgrad(st, Rs, LSTM, N, d_l2n, Rs, io->h_out, LSTM, g->ref_last_fc2.w, LSTM, g->ref_last_fc2.b));
  }
  // LSTM cell
  lstm_bwd_kernel<<<ceil_div(N * LSTM, 256), 256, 0, st>>>(N, A.gates, io->c, dh2, d_c, dgates, dc);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  OP3_TRY(gemm(st, N, XW, 4 * LSTM, dgates, 4 * LSTM, 0, p->lstm_w_ih, XW, 0, dX, XW, nullptr, ACT_NONE, 0));
  OP3_TRY(gemm(st, N, LSTM, 4 * LSTM, dgates, 4 * LSTM, 0, p->lstm_w_hh, LSTM, 0, dh, LSTM, nullptr, ACT_NONE, 1));
  OP3_TRY(gemm_wgrad(st, 4 * LSTM, XW, N, dgates, 4 * LSTM, A.X, XW, g->lstm_w_ih, XW, g->lstm_b_ih));
  OP3_TRY(gemm_wgrad(st, 4 * LSTM, LSTM, N, dgates, 4 * LSTM, io->h, LSTM, g->lstm_w_hh, LSTM, g->lstm_b_hh));
  // extras: LN1d parameter gradients and the direct paths into the previous state's lambdas / samples
  for (int j = 0; j < 2; ++j) {
    colsum_prod_kernel<<<ceil_div(Rs, 32), 256, 0, st>>>(N, Rs, dX + HID + j * Rs, XW, A.xn + j * Rs, 2 * Rs, g->ln1d_scale[j]);
    OP3_COUNT_LAUNCH();
    colsum_prod_kernel<<<ceil_div(Rs, 32), 256, 0, st>>>(N, Rs, dX + HID + j * Rs, XW, nullptr, 0, g->ln1d_center[j]);
    OP3_COUNT_LAUNCH();
  }
  add_cols_kernel<<<ceil_div(N * Rs, 256), 256, 0, st>>>(N, Rs, dX + HID + 2 * Rs, XW, dl1, Rs);
  OP3_COUNT_LAUNCH();
  add_cols_kernel<<<ceil_div(N * Rs, 256), 256, 0, st>>>(N, Rs, dX + HID + 3 * Rs, XW, dl2, Rs);
  OP3_COUNT_LAUNCH();
  add_cols_kernel<<<ceil_div(N * Rs, 256), 256, 0, st>>>(N, Rs, dX + HID + 4 * Rs, XW, dz_acc + d->Rd, R);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  // FC (+ELU) on the pooled features
  OP3_TRY(act_backward(st, N, HID, dX, XW, A.X, XW, ACT_ELU));
  OP3_TRY(gemm(st, N, Rs, HID, dX, XW, 0, p->ref_fc.w, Rs, 0, dpool, Rs, nullptr, ACT_NONE, 0));
  OP3_TRY(gemm_wgrad(st, HID, Rs, N, dX, XW, A.pooled, Rs, g->ref_fc.w, Rs, g->ref_fc.b));
  }
  // conv stack
  // fp16 mode: maxima of r1, r2 come from the forward pass (acts), those of g3, g2, g1 are recorded in slots 1..3 here
  const bool f16 = op3_get_precision() == OP3_PREC_F16X3;
  const float* atcs = ref_acts_tcs(d, io->acts);
  if (f16) OP3_CHECK(cudaMemsetAsync(tcs, 0, tc_scratch_floats(N) * sizeof(float), st));
  avgpool_bwd_kernel<<<dim3(ceil_div(DD, 256), N), 256, 0, st>>>(dpool, A.r3, Rs / 4, DD, g3,
                                                                         f16 ? tc_slot(tcs, N, 1) : nullptr);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  ConvInput x2 = one_src(A.r2, 8), x1 = one_src(A.r1, 8), gi3 = one_src(g3, Rs / 4), gi2 = one_src(g2, 8), gi1 = one_src(g1, 8);
  const uint32_t *m3 = nullptr, *m2 = nullptr, *m1 = nullptr;
  if (f16) {
    x2.maxbits = tc_slot(atcs, N, 2); x1.maxbits = tc_slot(atcs, N, 1);
    m3 = tc_slot(tcs, N, 1); m2 = tc_slot(tcs, N, 2); m1 = tc_slot(tcs, N, 3);
    gi3.maxbits = m3; gi3.out_maxbits = tc_slot(tcs, N, 2);
    gi2.maxbits = m2; gi2.out_maxbits = tc_slot(tcs, N, 3);
    gi1.maxbits = m1;
  }
  // weight gradients go to the weight-gradient lane (common.cuh: `st` itself unless op3_set_wgrad_overlap(1)); g3, g2, g1 are
  // separate buffers and nothing on `st` writes what they read
  cudaStream_t sw = st;
  OP3_TRY(wgrad_stream(st, &sw));
  OP3_TRY(conv5x5_wgrad_auto(sw, N, D, x2, Rs, g3, wscr, pg + L.ref_wf[2], pg + L.ref_bf[2], m3));
  OP3_TRY(conv5x5_auto(st, N, D, gi3, 32, prep + L.ref_wd[2], prep + L.ref_td[2], nullptr, g2, EPI_MUL_ELUGRAD, A.r2, tcs));
  OP3_TRY(wgrad_stream(st, &sw));
  OP3_TRY(conv5x5_wgrad_auto(sw, N, D, x1, 32, g2, wscr, pg + L.ref_wf[1], pg + L.ref_bf[1], m2));
  OP3_TRY(conv5x5_auto(st, N, D, gi2, 32, prep + L.ref_wd[1], prep + L.ref_td[1], nullptr, g1, EPI_MUL_ELUGRAD, A.r1, tcs));
  OP3_TRY(wgrad_stream(st, &sw));
  ConvInput x0 = refine_l1_input(d, io, prep, L);
  if (f16) x0.maxbits = tc_slot(atcs, N, 0);     // reduced by the forward pass's first conv call into slot 0 of the acts' scratch
  OP3_TRY(conv5x5_wgrad_auto(sw, N, D, x0, 32, g1, wscr, pg + L.ref_wf[0], pg + L.ref_bf[0], m1));
  OP3_TRY(conv5x5_auto(st, N, D, gi1, REF_CIN_PAD, prep + L.ref_wd[0], prep + L.ref_td[0], nullptr,
                       reinterpret_cast<float4*>(d_in), EPI_NONE, nullptr, tcs));
  return OP3_OK;
}

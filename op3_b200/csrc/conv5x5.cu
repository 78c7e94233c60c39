// 5x5 / stride 1 / pad 2 convolution over float4 channel-chunk ("C4") activations: [n][c/4][y][x] float4.
// fp32 CUDA-core ("exact") path: forward (also used as dgrad with flipped/transposed weights) and wgrad.
// Reference ops replaced: nn.Conv2d in BroadcastCNN.apply_forward (op3/torch/conv_networks.py:343-350) and
// RefinementNetwork._apply_forward (op3/torch/op3_modules/refinement_network.py:224-231) and their autograd
// backward.  The conv input may be gathered from several tensors (ConvInput) so that the 17-channel
// refinement input (op3_model.py:219-229) is never materialised.
#include "common.cuh"

namespace op3 {

namespace {

constexpr int TR = 8;            // output rows per CTA (forward)
constexpr int TW = 64;           // output cols per CTA
constexpr int HW = TW + 4;       // halo width
constexpr int FWD_IN_F4 = (TR + 4) * HW;

__device__ __forceinline__ const float4* src_chunk_ptr(const ConvInput& in, int c, int n, int D) {
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

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int COUT, int COUT_T>
__global__ void __launch_bounds__(128 * (COUT / COUT_T))
conv5x5_fwd_kernel(ConvInput in, int D, const float* __restrict__ wt, const float* __restrict__ bias,
                   float4* __restrict__ out, int epi, const float4* __restrict__ ref) {
  constexpr int NGRP = COUT / COUT_T;
  constexpr int NT = 128 * NGRP;
  constexpr int W_F4 = 25 * COUT;  // float4 count of one weight chunk slab: 25 taps * 4 ci * COUT / 4
  extern __shared__ float4 smem[];
  float4* s_in[2] = {smem, smem + FWD_IN_F4 + W_F4};
  float4* s_w[2] = {smem + FWD_IN_F4, smem + 2 * FWD_IN_F4 + W_F4};

  const int tid = threadIdx.x;
  const int g = tid / 128, pq = tid % 128, r = pq / 16, q = pq % 16;
  const int tiles_x = (D + TW - 1) / TW;
  const int tx0 = (blockIdx.x % tiles_x) * TW, ty0 = (blockIdx.x / tiles_x) * TR;
  const int n = blockIdx.y;
  const int cin_pad = in.total_chunks * 4;

  float acc[4][COUT_T];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < COUT_T; ++c) acc[j][c] = bias ? bias[g * COUT_T + c] : 0.f;

  auto load_chunk = [&](int c, int stage) {
    const float4* src = src_chunk_ptr(in, c, n, D);
    for (int e = tid; e < FWD_IN_F4; e += NT) {
      int yy = e / HW, xx = e % HW;
      int gy = ty0 + yy - 2, gx = tx0 + xx - 2;
      bool ok = gy >= 0 && gy < D && gx >= 0 && gx < D;
      cp_async16(&s_in[stage][e], ok ? src + (size_t)gy * D + gx : src, ok);
    }
    // weights: for each tap, 4*COUT contiguous floats at wt[(tap*cin_pad + c*4)*COUT]
    for (int e = tid; e < W_F4; e += NT) {
      int tap = e / COUT, o = e % COUT;  // o indexes float4 within the 4*COUT floats of this tap
      cp_async16(&s_w[stage][e], wt + ((size_t)tap * cin_pad + c * 4) * COUT + o * 4, true);
    }
    cp_async_commit();
  };

  const int nch = in.total_chunks;
  load_chunk(0, 0);
  for (int c = 0; c < nch; ++c) {
    int stage = c & 1;
    if (c + 1 < nch) {
      load_chunk(c + 1, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float4* tin = s_in[stage];
    const float* tw = reinterpret_cast<const float*>(s_w[stage]);
#pragma unroll 1
    for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = tin[(r + ky) * HW + q + 16 * j + kx];
        const float* wp = tw + (size_t)((ky * 5 + kx) * 4) * COUT + g * COUT_T;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
#pragma unroll
          for (int cg = 0; cg < COUT_T / 4; ++cg) {
            float4 w = *reinterpret_cast<const float4*>(wp + ci * COUT + cg * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a = ci == 0 ? v[j].x : (ci == 1 ? v[j].y : (ci == 2 ? v[j].z : v[j].w));
              acc[j][cg * 4 + 0] = fmaf(a, w.x, acc[j][cg * 4 + 0]);
              acc[j][cg * 4 + 1] = fmaf(a, w.y, acc[j][cg * 4 + 1]);
              acc[j][cg * 4 + 2] = fmaf(a, w.z, acc[j][cg * 4 + 2]);
              acc[j][cg * 4 + 3] = fmaf(a, w.w, acc[j][cg * 4 + 3]);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  const int gy = ty0 + r;
  if (gy >= D) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int gx = tx0 + q + 16 * j;
    if (gx >= D) continue;
#pragma unroll
    for (int cg = 0; cg < COUT_T / 4; ++cg) {
      size_t o = (((size_t)n * (COUT / 4) + g * (COUT_T / 4) + cg) * D + gy) * D + gx;
      float4 v = make_float4(acc[j][cg * 4], acc[j][cg * 4 + 1], acc[j][cg * 4 + 2], acc[j][cg * 4 + 3]);
      if (epi == EPI_ELU) {
        v.x = elu_f(v.x); v.y = elu_f(v.y); v.z = elu_f(v.z); v.w = elu_f(v.w);
      } else if (epi == EPI_MUL_ELUGRAD) {
        float4 a = ref[o];
        v.x *= elu_grad_from_out(a.x); v.y *= elu_grad_from_out(a.y);
        v.z *= elu_grad_from_out(a.z); v.w *= elu_grad_from_out(a.w);
      }
      out[o] = v;
    }
  }
}

template <int COUT, int COUT_T>
int launch_fwd(cudaStream_t st, int N, int D, const ConvInput& in, const float* wt, const float* bias, float4* out,
               int epi, const float4* ref) {
  constexpr int NT = 128 * (COUT / COUT_T);
  size_t smem = 2 * (size_t)(FWD_IN_F4 + 25 * COUT) * sizeof(float4);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_fwd_kernel<COUT, COUT_T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  }
  int tiles = ceil_div(D, TW) * ceil_div(D, TR);
  dim3 grid(tiles, N);
  conv5x5_fwd_kernel<COUT, COUT_T><<<grid, NT, smem, st>>>(in, D, wt, bias, out, epi, ref);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

// ------------------------------------------------------------------------------------------------
// wgrad: dW[tap][ci][co] = sum_{n,y,x} in[n][ci][y+ky-2][x+kx-2] * dpre[n][co][y][x]
// grid = (tile sets S, input chunks); every CTA keeps its 25*4*COUT partial in registers over a strided
// set of (n, 4-row band) tiles, then a second kernel reduces over S in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int WR = 4;                        // rows per wgrad tile
constexpr int WG_IN_F4 = (WR + 4) * HW;      // halo tile of one input chunk

template <int COUT, int COG, int PG>
__global__ void __launch_bounds__(20 * (COUT / COG) * PG)
conv5x5_wgrad_kernel(ConvInput in, int N, int D, const float4* __restrict__ dpre, float* __restrict__ partial) {
  constexpr int NCG = COUT / COG;
  constexpr int NT = 20 * NCG * PG;
  constexpr int DP_F4 = (COUT / 4) * WR * TW;
  constexpr int ROWS_PER_PG = WR / PG;
  static_assert(WR % PG == 0, "pixel groups must divide the tile rows");
  extern __shared__ float4 smem[];
  float4* s_in[2] = {smem, smem + WG_IN_F4 + DP_F4};
  float4* s_dp[2] = {smem + WG_IN_F4, smem + 2 * WG_IN_F4 + DP_F4};

  const int tid = threadIdx.x;
  const int pg = tid / (20 * NCG);
  const int rem = tid % (20 * NCG);
  const int cog = rem / 20, ky = (rem % 20) / 4, ci = rem % 4;
  const int c = blockIdx.y;  // input chunk
  const int tiles_x = (D + TW - 1) / TW, tiles_y = D / WR;
  const int tiles_per_img = tiles_x * tiles_y;
  const long long total_tiles = (long long)N * tiles_per_img;

  float acc[5][COG];
  float bsum[COG];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int o = 0; o < COG; ++o) acc[k][o] = 0.f;
#pragma unroll
  for (int o = 0; o < COG; ++o) bsum[o] = 0.f;

  auto load_tile = [&](long long t, int stage) {
    int n = (int)(t / tiles_per_img);
    int ti = (int)(t % tiles_per_img);
    int tx0 = (ti % tiles_x) * TW, ty0 = (ti / tiles_x) * WR;
    const float4* src = src_chunk_ptr(in, c, n, D);
    for (int e = tid; e < WG_IN_F4; e += NT) {
      int yy = e / HW, xx = e % HW;
      int gy = ty0 + yy - 2, gx = tx0 + xx - 2;
      bool ok = gy >= 0 && gy < D && gx >= 0 && gx < D;
      cp_async16(&s_in[stage][e], ok ? src + (size_t)gy * D + gx : src, ok);
    }
    for (int e = tid; e < DP_F4; e += NT) {
      int ch = e / (WR * TW), p = e % (WR * TW);
      int yy = p / TW, xx = p % TW;
      int gy = ty0 + yy, gx = tx0 + xx;
      bool ok = gy < D && gx < D;
      const float4* s = dpre + (((size_t)n * (COUT / 4) + ch) * D + (ok ? gy : 0)) * D + (ok ? gx : 0);
      cp_async16(&s_dp[stage][e], s, ok);
    }
    cp_async_commit();
  };

  long long t = blockIdx.x;
  int stage = 0;
  if (t < total_tiles) load_tile(t, 0);
  for (; t < total_tiles; t += gridDim.x, stage ^= 1) {
    long long tn = t + gridDim.x;
    if (tn < total_tiles) {
      load_tile(tn, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* tin = reinterpret_cast<const float*>(s_in[stage]);
    const float4* tdp = s_dp[stage];
#pragma unroll 1
    for (int rr = 0; rr < ROWS_PER_PG; ++rr) {
      const int y = pg * ROWS_PER_PG + rr;
      const float* row = tin + (size_t)((y + ky) * HW) * 4 + ci;
      float iv[5];
#pragma unroll
      for (int k = 0; k < 4; ++k) iv[k + 1] = row[k * 4];
#pragma unroll 4
      for (int x = 0; x < TW; ++x) {
#pragma unroll
        for (int k = 0; k < 4; ++k) iv[k] = iv[k + 1];
        iv[4] = row[(x + 4) * 4];
        float dp[COG];
#pragma unroll
        for (int o4 = 0; o4 < COG / 4; ++o4) {
          float4 d = tdp[((cog * (COG / 4) + o4) * WR + y) * TW + x];
          dp[o4 * 4] = d.x; dp[o4 * 4 + 1] = d.y; dp[o4 * 4 + 2] = d.z; dp[o4 * 4 + 3] = d.w;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
          for (int o = 0; o < COG; ++o) acc[k][o] = fmaf(iv[k], dp[o], acc[k][o]);
        if (c == 0 && ky == 0 && ci == 0) {
#pragma unroll
          for (int o = 0; o < COG; ++o) bsum[o] += dp[o];
        }
      }
    }
    __syncthreads();
  }

  // cross pixel-group reduction through smem, then one partial slab per CTA
  float* red = reinterpret_cast<float*>(smem);  // [PG][20*NCG][5*COG + COG]
  constexpr int PER_T = 6 * COG;
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int o = 0; o < COG; ++o) red[(size_t)(pg * 20 * NCG + rem) * PER_T + k * COG + o] = acc[k][o];
#pragma unroll
  for (int o = 0; o < COG; ++o) red[(size_t)(pg * 20 * NCG + rem) * PER_T + 5 * COG + o] = bsum[o];
  __syncthreads();
  if (pg == 0) {
    const int cin_pad = in.total_chunks * 4;
    float* slab = partial + (size_t)blockIdx.x * (25 * cin_pad * COUT + COUT);
    for (int e = 0; e < PER_T; ++e) {
      float s = 0.f;
#pragma unroll
      for (int p = 0; p < PG; ++p) s += red[(size_t)(p * 20 * NCG + rem) * PER_T + e];
      if (e < 5 * COG) {
        int kx = e / COG, o = e % COG;
        slab[((size_t)(ky * 5 + kx) * cin_pad + c * 4 + ci) * COUT + cog * COG + o] = s;
      } else if (c == 0 && ky == 0 && ci == 0) {
        slab[(size_t)25 * cin_pad * COUT + cog * COG + (e - 5 * COG)] = s;
      }
    }
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int S, int n_w, int n_b,
                                    float* __restrict__ dwt, float* __restrict__ dbias) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int tot = n_w + n_b;
  if (i >= tot) return;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += partial[(size_t)k * tot + i];
  if (i < n_w) dwt[i] += s;
  else if (dbias) dbias[i - n_w] += s;
}

// ------------------------------------------------------------------------------------------------
// wgrad v2 (cout >= 32): a thread owns ALL 25 taps of one input channel x 4 output channels (100 accumulators) and
// walks one tile row with a 5x5 sliding register window: per pixel 5 LDS.32 + 1 LDS.128 feed 100 FMAs, the window
// loads are conflict-free (lanes = 4 consecutive channels x broadcast) -> FMA-pipe bound instead of smem bound.
// ------------------------------------------------------------------------------------------------
template <int COUT, int WR2>
__global__ void __launch_bounds__(COUT * WR2)
conv5x5_wgrad2_kernel(ConvInput in, int N, int D, const float4* __restrict__ dpre, float* __restrict__ partial) {
  constexpr int NSUB = COUT / 4;            // float4 groups of output channels
  constexpr int GT = 4 * NSUB;              // threads per pixel group (= COUT)
  constexpr int NT = GT * WR2;              // one pixel group per tile row
  constexpr int IN_F4 = (WR2 + 4) * HW;
  constexpr int DP_F4 = NSUB * WR2 * TW;
  extern __shared__ float4 smem[];
  float4* s_in[2] = {smem, smem + IN_F4 + DP_F4};
  float4* s_dp[2] = {smem + IN_F4, smem + 2 * IN_F4 + DP_F4};

  const int tid = threadIdx.x;
  const int pg = tid / GT, rem = tid % GT, sub = rem / 4, ci = rem % 4;
  const int c = blockIdx.y;
  const int tiles_x = (D + TW - 1) / TW, tiles_y = D / WR2;
  const int tiles_per_img = tiles_x * tiles_y;
  const long long total_tiles = (long long)N * tiles_per_img;

  float acc[25][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 25; ++t)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[t][o] = 0.f;

  auto load_tile = [&](long long t, int stage) {
    int n = (int)(t / tiles_per_img);
    int ti = (int)(t % tiles_per_img);
    int tx0 = (ti % tiles_x) * TW, ty0 = (ti / tiles_x) * WR2;
    const float4* src = src_chunk_ptr(in, c, n, D);
    for (int e = tid; e < IN_F4; e += NT) {
      int yy = e / HW, xx = e % HW;
      int gy = ty0 + yy - 2, gx = tx0 + xx - 2;
      bool ok = gy >= 0 && gy < D && gx >= 0 && gx < D;
      cp_async16(&s_in[stage][e], ok ? src + (size_t)gy * D + gx : src, ok);
    }
    for (int e = tid; e < DP_F4; e += NT) {
      int ch = e / (WR2 * TW), p = e % (WR2 * TW);
      int gy = ty0 + p / TW, gx = tx0 + p % TW;
      bool ok = gy < D && gx < D;
      const float4* sp = dpre + (((size_t)n * NSUB + ch) * D + (ok ? gy : 0)) * D + (ok ? gx : 0);
      cp_async16(&s_dp[stage][e], sp, ok);
    }
    cp_async_commit();
  };

  long long t = blockIdx.x;
  int stage = 0;
  if (t < total_tiles) load_tile(t, 0);
  for (; t < total_tiles; t += gridDim.x, stage ^= 1) {
    long long tn = t + gridDim.x;
    if (tn < total_tiles) {
      load_tile(tn, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* tin = reinterpret_cast<const float*>(s_in[stage]) + ci;
    const float4* tdp = s_dp[stage] + (size_t)(sub * WR2 + pg) * TW;
    float w[5][5];                           // w[ky][slot]: input window, slot (x + kx) % 5 holds column x + kx
#pragma unroll
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
      for (int k = 0; k < 4; ++k) w[ky][k] = tin[(size_t)((pg + ky) * HW + k) * 4];
#pragma unroll 1
    for (int x0 = 0; x0 < TW; x0 += 5) {
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int x = x0 + u;
        if (x < TW) {
#pragma unroll
          for (int ky = 0; ky < 5; ++ky) w[ky][(u + 4) % 5] = tin[(size_t)((pg + ky) * HW + x + 4) * 4];
          const float4 d = tdp[x];
#pragma unroll
          for (int ky = 0; ky < 5; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
              const float a = w[ky][(u + kx) % 5];
              acc[ky * 5 + kx][0] = fmaf(a, d.x, acc[ky * 5 + kx][0]);
              acc[ky * 5 + kx][1] = fmaf(a, d.y, acc[ky * 5 + kx][1]);
              acc[ky * 5 + kx][2] = fmaf(a, d.z, acc[ky * 5 + kx][2]);
              acc[ky * 5 + kx][3] = fmaf(a, d.w, acc[ky * 5 + kx][3]);
            }
          if (c == 0 && ci == 0) { bsum[0] += d.x; bsum[1] += d.y; bsum[2] += d.z; bsum[3] += d.w; }
        }
      }
    }
    __syncthreads();
  }

  // cross pixel-group reduction through smem (fixed order), one partial slab per CTA
  float* red = reinterpret_cast<float*>(smem);  // [WR2][GT][104]
  constexpr int PER_T = 104;
#pragma unroll
  for (int tp = 0; tp < 25; ++tp)
#pragma unroll
    for (int o = 0; o < 4; ++o) red[(size_t)(pg * GT + rem) * PER_T + tp * 4 + o] = acc[tp][o];
#pragma unroll
  for (int o = 0; o < 4; ++o) red[(size_t)(pg * GT + rem) * PER_T + 100 + o] = bsum[o];
  __syncthreads();
  const int cin_pad = in.total_chunks * 4;
  float* slab = partial + (size_t)blockIdx.x * (25 * cin_pad * COUT + COUT);
  for (int e = pg; e < PER_T; e += WR2) {      // the WR2 groups share the 104 outputs of each (ci, sub)
    float sacc = 0.f;
#pragma unroll
    for (int p = 0; p < WR2; ++p) sacc += red[(size_t)(p * GT + rem) * PER_T + e];
    if (e < 100) {
      int tp = e / 4, o = e % 4;
      slab[((size_t)tp * cin_pad + c * 4 + ci) * COUT + sub * 4 + o] = sacc;
    } else if (c == 0 && ci == 0) {
      slab[(size_t)25 * cin_pad * COUT + sub * 4 + (e - 100)] = sacc;
    }
  }
}

int wgrad_sets(int chunks) { int s = 296 / chunks; return s < 1 ? 1 : s; }

template <int COUT, int COG, int PG>
int launch_wgrad(cudaStream_t st, int N, int D, const ConvInput& in, const float4* dpre, float* scratch, float* dwt,
                 float* dbias) {
  constexpr int NT = 20 * (COUT / COG) * PG;
  constexpr int DP_F4 = (COUT / 4) * WR * TW;
  size_t smem = 2 * (size_t)(WG_IN_F4 + DP_F4) * sizeof(float4);
  size_t red = (size_t)NT * 6 * COG * sizeof(float);
  if (red > smem) smem = red;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_wgrad_kernel<COUT, COG, PG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  }
  int chunks = in.total_chunks;
  long long total_tiles = (long long)N * ceil_div(D, TW) * (D / WR);
  int S = wgrad_sets(chunks);
  if (S > total_tiles) S = (int)total_tiles;
  dim3 grid(S, chunks);
  conv5x5_wgrad_kernel<COUT, COG, PG><<<grid, NT, smem, st>>>(in, N, D, dpre, scratch);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  int n_w = 25 * chunks * 4 * COUT, n_b = COUT;
  wgrad_reduce_kernel<<<ceil_div(n_w + n_b, 256), 256, 0, st>>>(scratch, S, n_w, n_b, dwt, dbias);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

template <int COUT, int WR2>
int launch_wgrad2(cudaStream_t st, int N, int D, const ConvInput& in, const float4* dpre, float* scratch, float* dwt,
                  float* dbias) {
  constexpr int NT = COUT * WR2;
  constexpr int IN_F4 = (WR2 + 4) * HW, DP_F4 = (COUT / 4) * WR2 * TW;
  size_t smem = 2 * (size_t)(IN_F4 + DP_F4) * sizeof(float4);
  size_t red = (size_t)NT * 104 * sizeof(float);
  if (red > smem) smem = red;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    OP3_CHECK(cudaFuncSetAttribute(conv5x5_wgrad2_kernel<COUT, WR2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int chunks = in.total_chunks;
  long long total_tiles = (long long)N * ceil_div(D, TW) * (D / WR2);
  int S = wgrad_sets(chunks);
  if (S > total_tiles) S = (int)total_tiles;
  dim3 grid(S, chunks);
  conv5x5_wgrad2_kernel<COUT, WR2><<<grid, NT, smem, st>>>(in, N, D, dpre, scratch);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  int n_w = 25 * chunks * 4 * COUT, n_b = COUT;
  wgrad_reduce_kernel<<<ceil_div(n_w + n_b, 256), 256, 0, st>>>(scratch, S, n_w, n_b, dwt, dbias);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

}  // namespace

int conv5x5_forward(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float* wt, const float* bias,
                    float4* out, int epi, const float4* elu_ref) {
  OP3_REQUIRE(D % TR == 0, "conv5x5: D=%d must be a multiple of %d", D, TR);
  if (N <= 0) return OP3_OK;
  const int cin_real = in.real_cin ? in.real_cin : in.total_chunks * 4;
  ProfScope prof(st, KID_CONV_FWD, 2.0 * 25 * cin_real * (cout == 20 ? 17 : cout) * (double)N * D * D);
  switch (cout) {
    case 4:  return launch_fwd<4, 4>(st, N, D, in, wt, bias, out, epi, elu_ref);
    case 20: return launch_fwd<20, 20>(st, N, D, in, wt, bias, out, epi, elu_ref);
    case 32: return launch_fwd<32, 16>(st, N, D, in, wt, bias, out, epi, elu_ref);
    case 64: return launch_fwd<64, 16>(st, N, D, in, wt, bias, out, epi, elu_ref);
    default: break;
  }
  set_error("conv5x5_forward: unsupported cout=%d", cout);
  return OP3_ERR_ARG;
}

size_t conv5x5_wgrad_scratch_floats(int cin_pad, int cout) {
  // also covers the tcgen05 kernels (conv5x5_wgrad_tc.cu): up to 148 per-CTA partial slabs + 8 words of operand maxima
  const size_t slab = (size_t)25 * cin_pad * cout + cout;
  const size_t a = (size_t)wgrad_sets(cin_pad / 4) * slab, b = 148 * slab + 8;
  return a > b ? a : b;
}

int conv5x5_wgrad(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float4* dpre, float* scratch,
                  float* dwt, float* dbias) {
  OP3_REQUIRE(D % 8 == 0, "conv5x5_wgrad: D=%d must be a multiple of 8", D);
  if (N <= 0) return OP3_OK;
  const int cin_real = in.real_cin ? in.real_cin : in.total_chunks * 4;
  ProfScope prof(st, KID_CONV_WGRAD, 2.0 * 25 * cin_real * cout * (double)N * D * D);
  switch (cout) {
    case 4:  return launch_wgrad<4, 4, 4>(st, N, D, in, dpre, scratch, dwt, dbias);
    case 32: return launch_wgrad2<32, 8>(st, N, D, in, dpre, scratch, dwt, dbias);
    case 64: return launch_wgrad2<64, 4>(st, N, D, in, dpre, scratch, dwt, dbias);
    default: break;
  }
  set_error("conv5x5_wgrad: unsupported cout=%d", cout);
  return OP3_ERR_ARG;
}

}  // namespace op3

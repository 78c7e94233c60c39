// Small-GEMM + elementwise helpers (fp32 CUDA-core path) used by the refinement head, the pairwise dynamics network
//  and the collapsed decoder layer 1.  These are < 1 % of the step FLOPs; shapes are
// M = B*K or B*K*(K-1) rows, widths 1..800.
#include "common.cuh"

namespace op3 {

namespace {

constexpr int BM = 32, BN = 32, BK = 16, KG = 4;

// These GEMMs are tiny (M = B*K or B*K*(K-1) rows, widths <= 800, K <= 800) and latency bound: a 64x64 tiling gives 2..50
// CTAs whose serial K loop pays one global-load latency per slab.  So: 32x32 output tiles (4x the CTAs) and the K range
// split over the four 64-thread groups of the CTA (4x shorter serial chain), each group register-prefetching its next
// 16-deep slab while it multiplies the current one; the four partial tiles are added in a fixed order through shared memory.
template <int TA, int TB>
__device__ __forceinline__ void gemm_tile(int M, int N, int K, const float* __restrict__ A, int lda,
                                          const float* __restrict__ B, int ldb, float* __restrict__ C,
                                          int ldc, const float* __restrict__ bias, int act, int accumulate,
                                          float* __restrict__ colsum_out, int tile_x, int tile_y) {
  static_assert(BM == BN, "square tiles");
  __shared__ float sm[2 * KG * BK * (BM + 4)];
  float (*As)[BK][BM + 4] = reinterpret_cast<float (*)[BK][BM + 4]>(sm);
  float (*Bs)[BK][BN + 4] = reinterpret_cast<float (*)[BK][BN + 4]>(sm + KG * BK * (BM + 4));
  const int tid = threadIdx.x;
  const int kg = tid >> 6, t = tid & 63;
  const int tx = t & 7, ty = t >> 3;
  const int m0 = tile_y * BM, n0 = tile_x * BN;
  const int Kq = ((K + KG * BK - 1) / (KG * BK)) * BK;      // K range of one group (multiple of BK)
  const int kbeg = kg * Kq, kend = min(K, kbeg + Kq);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[8], rb[8];
  // TA == 1 only (weight gradients dW = dY^T X): the first column of tiles also accumulates sum_k A[k][m] = the bias gradient
  const bool do_cs = TA == 1 && colsum_out != nullptr && tile_x == 0 && tx == 0;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};

  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int e = t + r * 64;
      int mm, kk;
      if (TA == 0) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int gm = m0 + mm, gk = k0 + kk;
      ra[r] = (gm < M && gk < kend) ? (TA == 0 ? A[(size_t)gm * lda + gk] : A[(size_t)gk * lda + gm]) : 0.f;
      int nn;
      if (TB == 0) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      int gn = n0 + nn;
      gk = k0 + kk;
      rb[r] = (gn < N && gk < kend) ? (TB == 0 ? B[(size_t)gk * ldb + gn] : B[(size_t)gn * ldb + gk]) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int e = t + r * 64;
      int mm, kk;
      if (TA == 0) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      As[kg][kk][mm] = ra[r];
      int nn;
      if (TB == 0) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      Bs[kg][kk][nn] = rb[r];
    }
  };

  fetch(kbeg);
  for (int it = 0; it < Kq / BK; ++it) {                   // same trip count for every group: barriers stay uniform
    const int k0 = kbeg + it * BK;
    stash();
    __syncthreads();
    if (it + 1 < Kq / BK) fetch(k0 + BK);                  // in flight while this slab is multiplied
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kg][kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kg][kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (do_cs) {
#pragma unroll
        for (int i = 0; i < 4; ++i) cs[i] += a[i];
      }
    }
    __syncthreads();
  }
  // partial tiles of groups 1..3 -> shared memory (aliasing the operand buffers), summed by group 0 in a fixed order
  float (*part)[BM][BN + 1] = reinterpret_cast<float (*)[BM][BN + 1]>(sm);
  static_assert(sizeof(sm) >= (3 * BM * (BN + 1) + KG * BM) * sizeof(float), "reduction buffer does not fit");
  float (*cs_part)[BM] = reinterpret_cast<float (*)[BM]>(sm + 3 * BM * (BN + 1));
  if (do_cs) {
#pragma unroll
    for (int i = 0; i < 4; ++i) cs_part[kg][ty * 4 + i] = cs[i];
  }
  if (kg > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[kg - 1][ty * 4 + i][tx * 4 + j] = acc[i][j];
  }
  __syncthreads();
  if (kg == 0 && do_cs) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int mm = ty * 4 + i;
      if (m0 + mm < M) colsum_out[m0 + mm] += (cs_part[0][mm] + cs_part[1][mm]) + (cs_part[2][mm] + cs_part[3][mm]);
    }
  }
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int gm = m0 + ty * 4 + i;
      if (gm >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int gn = n0 + tx * 4 + j;
        if (gn >= N) continue;
        float v = ((acc[i][j] + part[0][ty * 4 + i][tx * 4 + j]) + part[1][ty * 4 + i][tx * 4 + j]) + part[2][ty * 4 + i][tx * 4 + j];
        if (bias) v += bias[gn];
        if (accumulate) v += C[(size_t)gm * ldc + gn];
        if (act == ACT_ELU) v = elu_f(v);
        else if (act == ACT_SIGMOID) v = sigmoid_f(v);
        else if (act == ACT_RELU) v = fmaxf(v, 0.f);
        else if (act == ACT_TANH) v = tanhf(v);
        C[(size_t)gm * ldc + gn] = v;
      }
    }
  }
}

template <int TA, int TB>
__global__ void __launch_bounds__(256) gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                   const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                   int ldc, const float* __restrict__ bias, int act, int accumulate,
                                                   float* __restrict__ colsum_out) {
  gemm_tile<TA, TB>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, colsum_out, blockIdx.x, blockIdx.y);
}

// All weight + bias gradients of a fused MLP chain in ONE launch: item i is gW[out][in] += dY^T X, gB[out] += colsum(dY)
// over `rows` rows; its tiles are the blocks [tile0[i], tile0[i+1]).
struct WgradBatch {
  WgradItem it[WGRAD_BATCH_MAX];
  int tile0[WGRAD_BATCH_MAX + 1];
  int n;
};
__global__ void __launch_bounds__(256) gemm_wgrad_batch_kernel(const __grid_constant__ WgradBatch b) {
  int i = 0;
  while (i + 1 < b.n && (int)blockIdx.x >= b.tile0[i + 1]) ++i;
  const WgradItem& w = b.it[i];
  const int t = blockIdx.x - b.tile0[i];
  const int tiles_x = (w.in + BN - 1) / BN;
  gemm_tile<1, 0>(w.out, w.in, w.rows, w.dY, w.lddy, w.X, w.ldx, w.gW, w.ldgw, nullptr, ACT_NONE, 1, w.gB, t % tiles_x, t / tiles_x);
}

__global__ void act_backward_kernel(int M, int N, float* __restrict__ dY, int ldd, const float* __restrict__ Y,
                                    int ldy, int act) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  int m = (int)(i / N), n = (int)(i % N);
  float y = Y[(size_t)m * ldy + n];
  float g = act == ACT_ELU ? elu_grad_from_out(y) : (act == ACT_SIGMOID ? y * (1.f - y) : 1.f);
  dY[(size_t)m * ldd + n] *= g;
}

// one block per 32 columns; 8 warps stride over rows; deterministic order
__global__ void __launch_bounds__(256) colsum_kernel(int M, int N, const float* __restrict__ X, int ldx,
                                                     float* __restrict__ out, int accumulate) {
  __shared__ float part[8][33];
  int lane = threadIdx.x % 32, w = threadIdx.x / 32;
  int n = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (n < N)
    for (int m = w; m < M; m += 8) s += X[(size_t)m * ldx + n];
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][lane];
    out[n] = accumulate ? out[n] + t : t;
  }
}

__global__ void axpy_kernel(long long n, const float* __restrict__ x, float* __restrict__ y) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

}  // namespace

int gemm(cudaStream_t st, int M, int N, int K, const float* A, int lda, int transA, const float* B, int ldb,
         int transB, float* C, int ldc, const float* bias, int act, int accumulate) {
  if (M <= 0 || N <= 0) return OP3_OK;
  ProfScope prof(st, KID_GEMM, 2.0 * M * N * K);
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  if (transA == 0 && transB == 0) gemm_kernel<0, 0><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, nullptr);
  else if (transA == 0 && transB == 1) gemm_kernel<0, 1><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, nullptr);
  else if (transA == 1 && transB == 0) gemm_kernel<1, 0><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, nullptr);
  else gemm_kernel<1, 1><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, nullptr);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

// gW[out][in] += dY^T X and gB[out] += column sums of dY in ONE launch (dY: [M][out] rows, X: [M][in] rows)
int gemm_wgrad(cudaStream_t st, int out, int in, int M, const float* dY, int lddy, const float* X, int ldx, float* gW,
               int ldgw, float* gB) {
  if (out <= 0 || in <= 0) return OP3_OK;
  ProfScope prof(st, KID_GEMM, 2.0 * out * in * M);
  dim3 grid(ceil_div(in, BN), ceil_div(out, BM));
  gemm_kernel<1, 0><<<grid, 256, 0, st>>>(out, in, M, dY, lddy, X, ldx, gW, ldgw, nullptr, ACT_NONE, 1, gB);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

int gemm_wgrad_batch(cudaStream_t st, const WgradItem* items, int n) {
  if (n <= 0) return OP3_OK;
  OP3_REQUIRE(n <= WGRAD_BATCH_MAX, "gemm_wgrad_batch: %d items > %d", n, WGRAD_BATCH_MAX);
  WgradBatch b;
  double work = 0.0;
  int tiles = 0, m = 0;
  for (int i = 0; i < n; ++i) {
    if (items[i].out <= 0 || items[i].in <= 0 || items[i].rows <= 0) continue;
    b.it[m] = items[i];
    b.tile0[m] = tiles;
    tiles += ceil_div(items[i].in, BN) * ceil_div(items[i].out, BM);
    work += 2.0 * items[i].out * items[i].in * items[i].rows;
    ++m;
  }
  b.tile0[m] = tiles;
  b.n = m;
  if (tiles == 0) return OP3_OK;
  ProfScope prof(st, KID_GEMM, work);
  gemm_wgrad_batch_kernel<<<tiles, 256, 0, st>>>(b);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

int act_backward(cudaStream_t st, int M, int N, float* dY, int ldd, const float* Y, int ldy, int act) {
  if (act == ACT_NONE || M <= 0 || N <= 0) return OP3_OK;
  long long n = (long long)M * N;
  act_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(M, N, dY, ldd, Y, ldy, act);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

int colsum(cudaStream_t st, int M, int N, const float* X, int ldx, float* out, int accumulate) {
  if (N <= 0) return OP3_OK;
  colsum_kernel<<<ceil_div(N, 32), 256, 0, st>>>(M, N, X, ldx, out, accumulate);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

int axpy(cudaStream_t st, long long n, const float* x, float* y) {
  if (n <= 0) return OP3_OK;
  axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, x, y);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

int fill_zero(cudaStream_t st, void* p, size_t bytes) {
  if (bytes == 0) return OP3_OK;
  OP3_CHECK(cudaMemsetAsync(p, 0, bytes, st));
  return OP3_OK;
}

}  // namespace op3

// y[M][out] = act(x[M][in] W^T + b): one nn.Linear + activation of networks.py:63-74 (Mlp.forward)
extern "C" int op3_linear_fwd(int M, int in, int out, const float* x, const float* w, const float* b, int act, float* y,
                              void* stream) {
  OP3_REQUIRE(M >= 0 && in > 0 && out > 0 && x && w && y, "op3_linear_fwd: bad arguments");
  OP3_REQUIRE(act >= op3::ACT_NONE && act <= op3::ACT_TANH, "op3_linear_fwd: activation %d not in 0..4", act);
  return op3::gemm((cudaStream_t)stream, M, out, in, x, in, 0, w, in, 1, y, out, b, act, 0);
}

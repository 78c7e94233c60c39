// Fused per-pixel mask-softmax over K + Gaussian-mixture log-likelihood + KL + refinement-input kernel family.
// Replaces: F.softmax(mask_logits, dim=1) (op3_model.py:306); _gaussian_prob (:123-126); get_loss (:313-327);
// kl_divergence_prior_post (:137-145); and in refine (:209-229) posterior_mask, leave_out_ll, the autograd.grad
// outputs x_hat_grad / mask_grad (closed forms) and the four LayerNorm2D
// (op3/torch/modules.py:48-69: per-sample mean / UNBIASED std over (C,H,W), (x-mean)/(std+1e-5), per-channel affine).
//
// HBM-bound.  One thread owns one pixel of one sample and keeps its K slot float4s in registers (coalesced 16-byte
// loads, slot-major planes), LN statistics are reduced warp-shuffle -> smem -> per-tile fp64 partials (fixed order,
// run-to-run deterministic) with sums centred on the value at pixel 0 to avoid cancellation.
#include <cooperative_groups.h>
#include "layout.cuh"

namespace cg = cooperative_groups;

namespace op3 {

namespace {

constexpr int MIX_THREADS = 256;
constexpr int STATS_PPT = 2;  // pixels per thread in the statistics pass
constexpr float MIX_INV_2S2 = 1.f / 0.06f;     // 1 / (channels * 2 sigma^2), sigma = 0.1 (op3_model.py:79,126)
constexpr float MIX_INV_NORM = 1.f / 0.24805f;  // 1 / (sqrt(sigma^6) * 248.05)
constexpr float MIX_INV_S2 = 1.f / 0.03f;       // d(-S/0.06)/dc = -(c - x)/0.03

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int KMAX>
struct Pixel {
  float4 v[KMAX];   // (r,g,b,logit) per slot
  float m[KMAX];    // softmax mask
  float p[KMAX];    // gaussian colour likelihood
  float x0, x1, x2; // target pixel
  float P;          // mixture likelihood
  float sp;         // sum_k p_k
  float gP;         // d clog / d P = (-1/B)/(P+1e-12)
  float rsp;        // 1 / (sum_k p_k + 1e-8)

  __device__ __forceinline__ void load(const float* dec, const float* img, int b, int K, int DD, int pix) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) v[k] = ldg4(dec + ((size_t)(b * K + k) * DD + pix) * 4);
    if (img) { x0 = __ldg(img + pix); x1 = __ldg(img + DD + pix); x2 = __ldg(img + 2 * DD + pix); }
    else { x0 = x1 = x2 = 0.f; }
  }
  __device__ __forceinline__ void compute(int K, float invB) {
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) mx = fmaxf(mx, v[k].w);
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) { m[k] = expf(v[k].w - mx); se += m[k]; }
    const float rse = 1.f / se;                        // one reciprocal per pixel; products are within 1 ulp of the quotients
    P = 0.f; sp = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        m[k] = m[k] * rse;
        float d0 = v[k].x - x0, d1 = v[k].y - x1, d2 = v[k].z - x2;
        float S = d0 * d0 + d1 * d1 + d2 * d2;
        p[k] = expf(-S * MIX_INV_2S2) * MIX_INV_NORM;  // op3_model.py:126 (ch*2*sigma^2 = 0.06, sqrt(sigma^6)*248.05)
        P += m[k] * p[k];
        sp += p[k];
      }
    gP = (-invB) / (P + 1e-12f);
    rsp = 1.f / (sp + 1e-8f);
  }
  __device__ __forceinline__ float mask_grad(int k) const { return gP * p[k]; }
  __device__ __forceinline__ float leave_out(int k) const { return P - m[k] * p[k]; }
  __device__ __forceinline__ float post_mask(int k) const { return p[k] * rsp; }
  // x_hat_grad = d clog / d colors = -gP * m * p * (c - x) / 0.03
  __device__ __forceinline__ float xh_coef(int k) const { return -gP * m[k] * p[k] * MIX_INV_S2; }
};

// stats buffer layout (in floats; the fp64 regions are 8-byte aligned because every count below is even)
struct StatsLayout {
  int tiles, nq;
  size_t ln;        // [B][K][3][2] + [B][2] floats: (mean, 1/(std+eps)) per LayerNorm2D group
  size_t partial;   // [B][tiles][nq] doubles: raw sum / sum of squares per group, -log P sum, squared recon error
  size_t sample;    // [B][3] doubles: per-sample clog, kl, squared error (+ 2 floats of counter at the end)
  size_t counter;   // [B + 1] unsigned: tiles done per sample, samples done (last-block-done reduction of the loss scalars)
  size_t gpart;     // [B*blocks][12] doubles (LN2D parameter-gradient partials of the backward pass)
  size_t total;
};
__host__ __device__ inline StatsLayout stats_layout(int B, int K, int D) {
  StatsLayout L;
  L.tiles = (D * D + MIX_THREADS * STATS_PPT - 1) / (MIX_THREADS * STATS_PPT);
  L.nq = 6 * K + 4;
  size_t o = 0;
  L.ln = o; o += (size_t)B * K * 6 + (size_t)B * 2; o = (o + 1) & ~(size_t)1;
  L.partial = o; o += 2 * (size_t)B * L.tiles * L.nq;
  L.sample = o; o += 2 * (size_t)B * 3;
  L.counter = o; o += (size_t)(B + 2) & ~(size_t)1;
  L.gpart = o; o += 2 * (size_t)B * ((D * D + MIX_THREADS - 1) / MIX_THREADS) * 12;
  L.total = o;
  return L;
}

__device__ __forceinline__ float softplus_std(float x) { return sqrtf(logf(1.f + expf(fminf(x, 80.f))) + 1e-5f); }

// ---------------------------------------------------------------------------------------------------------
// pass A: masks / colours / final_recon outputs + LN statistics + loss partials.
// Every warp centres its sums on the values of its FIRST pixel (fp32, no cancellation), reduces them with shuffles,
// converts them to raw fp64 sums (sum x = S1 + n s, sum x^2 = S2 + 2 s S1 + n s^2) and the block adds those in a fixed order.
// ---------------------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(MIX_THREADS)
mix_stats_kernel(Op3MixtureIO io, const float* __restrict__ mse_image, long long mse_bstride, int B, int K, int D,
                 int want_stats, int Rs, float beta) {
  const int DD = D * D;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  const float invB = 1.f / (float)B;
  const float* img = io.image ? io.image + (size_t)b * io.image_bstride : nullptr;
  StatsLayout L = stats_layout(B, K, D);
  __shared__ double s_red[MIX_THREADS / 32][6 * 16 + 4];
  __shared__ bool s_last;

  Pixel<KMAX> px;
  float acc[6 * KMAX], shift[3 * KMAX];
  float accP = 0.f, accP2 = 0.f, shiftP = 0.f, accLog = 0.f, accSq = 0.f;
  int npix = 0;
#pragma unroll
  for (int i = 0; i < 6 * KMAX; ++i) acc[i] = 0.f;

#pragma unroll 1
  for (int it = 0; it < STATS_PPT; ++it) {
    int pix = (tile * STATS_PPT + it) * MIX_THREADS + tid;
    if (pix >= DD) break;
    px.load(io.dec_out, img, b, K, DD, pix);
    px.compute(K, invB);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        if (io.mask) io.mask[(size_t)(b * K + k) * DD + pix] = px.m[k];
        if (io.colors) {
          float* c = io.colors + (size_t)(b * K + k) * 3 * DD + pix;
          c[0] = px.v[k].x; c[DD] = px.v[k].y; c[2 * DD] = px.v[k].z;
        }
        r0 += px.v[k].x * px.m[k]; r1 += px.v[k].y * px.m[k]; r2 += px.v[k].z * px.m[k];
        if (want_stats) {
          float mg = px.mask_grad(k), lo = px.leave_out(k), cf = px.xh_coef(k);
          float x0 = cf * (px.v[k].x - px.x0), x1 = cf * (px.v[k].y - px.x1), x2 = cf * (px.v[k].z - px.x2);
          if (npix == 0) {             // warp-uniform branch: DD % 32 == 0, a warp's 32 pixels are all valid or all not
            shift[3 * k] = __shfl_sync(0xffffffffu, mg, 0);
            shift[3 * k + 1] = __shfl_sync(0xffffffffu, lo, 0);
            shift[3 * k + 2] = __shfl_sync(0xffffffffu, x0, 0);
          }
          float d = mg - shift[3 * k];
          acc[6 * k] += d; acc[6 * k + 1] += d * d;
          d = lo - shift[3 * k + 1];
          acc[6 * k + 2] += d; acc[6 * k + 3] += d * d;
          float sh = shift[3 * k + 2];
          float e0 = x0 - sh, e1 = x1 - sh, e2 = x2 - sh;
          acc[6 * k + 4] += e0 + e1 + e2; acc[6 * k + 5] += e0 * e0 + e1 * e1 + e2 * e2;
        }
      }
    if (io.final_recon) {
      float* fr = io.final_recon + (size_t)b * 3 * DD + pix;
      fr[0] = r0; fr[DD] = r1; fr[2 * DD] = r2;
    }
    if (want_stats) {
      if (npix == 0) shiftP = __shfl_sync(0xffffffffu, px.P, 0);
      float d = px.P - shiftP;
      accP += d; accP2 += d * d;
      accLog += -logf(px.P + 1e-12f);
      if (mse_image) {
        const float* mi = mse_image + (size_t)b * mse_bstride;
        float q0 = r0 - __ldg(mi + pix), q1 = r1 - __ldg(mi + DD + pix), q2 = r2 - __ldg(mi + 2 * DD + pix);
        accSq += q0 * q0 + q1 * q1 + q2 * q2;
      }
    }
    ++npix;
  }
  if (!want_stats) return;

  // centred fp32 sums -> warp shuffle (fp32: 64 centred values) -> lane 0 converts to RAW fp64 sums -> smem -> per-(sample,
  // tile) partial.  All lanes of a warp share lane 0's shift, so the centred sums add directly.
  const double n1 = 32.0 * npix, n3 = 96.0 * npix;
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) {
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        float S1 = warp_sum(acc[6 * k + 2 * g]), S2 = warp_sum(acc[6 * k + 2 * g + 1]);
        if (lane == 0) {
          double sft = npix ? (double)shift[3 * k + g] : 0.0, nn = g == 2 ? n3 : n1;
          s_red[warp][6 * k + 2 * g] = (double)S1 + nn * sft;
          s_red[warp][6 * k + 2 * g + 1] = (double)S2 + 2.0 * sft * (double)S1 + nn * sft * sft;
        }
      }
    }
  {
    float S1 = warp_sum(accP), S2 = warp_sum(accP2), S3 = warp_sum(accLog), S4 = warp_sum(accSq);
    if (lane == 0) {
      double sft = npix ? (double)shiftP : 0.0;
      s_red[warp][6 * K] = (double)S1 + n1 * sft;
      s_red[warp][6 * K + 1] = (double)S2 + 2.0 * sft * (double)S1 + n1 * sft * sft;
      s_red[warp][6 * K + 2] = (double)S3;
      s_red[warp][6 * K + 3] = (double)S4;
    }
  }
  __syncthreads();
  if (tid < L.nq) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < MIX_THREADS / 32; ++w) s += s_red[w][tid];
    reinterpret_cast<double*>(io.stats + L.partial)[((size_t)b * L.tiles + tile) * L.nq + tid] = s;
  }

  // ---- loss scalars (get_loss, op3_model.py:313-327): the LAST tile of a sample sums that sample's partials (fixed
  // order) and its KL; the last sample to finish sums the per-sample values (fixed order).  Counters are zeroed by
  // the host (memset node) before the launch.
  unsigned* counters = reinterpret_cast<unsigned*>(io.stats + L.counter);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&counters[b], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const double* part = reinterpret_cast<const double*>(io.stats + L.partial);
  double* samp = reinterpret_cast<double*>(io.stats + L.sample);
  double kl = 0.0;
  for (int i = tid; i < K * Rs; i += MIX_THREADS) {
    size_t o = (size_t)b * K * Rs + i;
    float sq_ = softplus_std(io.post_l2[o]), spv = softplus_std(io.prior_l2[o]);
    float dm = io.post_l1[o] - io.prior_l1[o];
    kl += (double)(logf(spv / sq_) + (sq_ * sq_ + dm * dm) / (2.f * spv * spv) - 0.5f);
  }
  kl = warp_sum_d(kl);
  __syncthreads();                       // s_red is reused
  if (lane == 0) s_red[warp][0] = kl;
  __syncthreads();
  if (tid == 0) {
    double klv = 0.0, cl = 0.0, sq = 0.0;
#pragma unroll
    for (int w = 0; w < MIX_THREADS / 32; ++w) klv += s_red[w][0];
    for (int t = 0; t < L.tiles; ++t) {
      cl += __ldcg(part + ((size_t)b * L.tiles + t) * L.nq + 6 * K + 2);
      sq += __ldcg(part + ((size_t)b * L.tiles + t) * L.nq + 6 * K + 3);
    }
    samp[b * 3] = cl; samp[b * 3 + 1] = klv; samp[b * 3 + 2] = sq;
    __threadfence();
    if (atomicAdd(&counters[B], 1u) == (unsigned)B - 1) {
      __threadfence();
      cl = 0.0; klv = 0.0; sq = 0.0;
      for (int i = 0; i < B; ++i) { cl += __ldcg(samp + i * 3); klv += __ldcg(samp + i * 3 + 1); sq += __ldcg(samp + i * 3 + 2); }
      float clog = (float)(cl / B), klf = (float)(klv / B);
      io.losses[0] = clog;
      io.losses[1] = klf;
      io.losses[2] = clog + beta * klf;
      io.losses[3] = mse_image ? (float)(sq / ((double)B * 3.0 * D * D)) : 0.f;
    }
  }
}

// LN statistics of sample b, group q in [0, 3K] (q = 3K: the pixel-likelihood group): (mean, 1/(std + 1e-5)), unbiased std
__device__ __forceinline__ float2 ln_stat_from_partials(const double* part, const StatsLayout& L, int b, int K, int q, int DD) {
  int k = q / 3, g = q % 3;
  int qi = q == 3 * K ? 6 * K : 6 * k + 2 * g;
  double s = 0.0, s2 = 0.0;
  for (int t = 0; t < L.tiles; ++t) {
    s += part[((size_t)b * L.tiles + t) * L.nq + qi];
    s2 += part[((size_t)b * L.tiles + t) * L.nq + qi + 1];
  }
  double n = (q != 3 * K && g == 2) ? 3.0 * DD : (double)DD;
  double mean = s / n;
  double var = (s2 - s * s / n) / (n - 1.0);
  if (var < 0.0) var = 0.0;
  return make_float2((float)mean, 1.f / ((float)sqrt(var) + 1e-5f));
}

// ---------------------------------------------------------------------------------------------------------
// pass B: refinement-input chunks (LayerNorm2D applied) + the gradient seed d total / d dec_out.
// Each CTA derives the LayerNorm statistics of its sample from the fp64 partials (3K+1 tiny reductions).
// ---------------------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(MIX_THREADS)
mix_emit_kernel(Op3MixtureIO io, const float* __restrict__ g0, const float* __restrict__ g1, const float* __restrict__ g2,
                const float* __restrict__ g3, const float* __restrict__ b0, const float* __restrict__ b1,
                const float* __restrict__ b2, const float* __restrict__ b3, int B, int K, int D) {
  const int DD = D * D;
  const int b = blockIdx.y;
  const int pix = blockIdx.x * MIX_THREADS + threadIdx.x;
  const float invB = 1.f / (float)B;
  StatsLayout L = stats_layout(B, K, D);
  __shared__ float2 s_ln[3 * 16 + 1];
  if (threadIdx.x < 3 * K + 1) {
    float2 st = ln_stat_from_partials(reinterpret_cast<const double*>(io.stats + L.partial), L, b, K, threadIdx.x, DD);
    s_ln[threadIdx.x] = st;
    if (blockIdx.x == 0) {     // keep them for the backward pass
      int q = threadIdx.x;
      size_t o = q == 3 * K ? L.ln + (size_t)B * K * 6 + (size_t)b * 2 : L.ln + ((size_t)(b * K + q / 3) * 3 + q % 3) * 2;
      io.stats[o] = st.x; io.stats[o + 1] = st.y;
    }
  }
  __syncthreads();
  if (pix >= DD) return;
  const float* img = io.image + (size_t)b * io.image_bstride;
  Pixel<KMAX> px;
  px.load(io.dec_out, img, b, K, DD, pix);
  px.compute(K, invB);
  const float ga0 = __ldg(g0), be0 = __ldg(b0), ga1 = __ldg(g1), be1 = __ldg(b1), ga2 = __ldg(g2), be2 = __ldg(b2);
  const float ga3x = __ldg(g3), ga3y = __ldg(g3 + 1), ga3z = __ldg(g3 + 2);
  const float be3x = __ldg(b3), be3y = __ldg(b3 + 1), be3z = __ldg(b3 + 2);
  const float gPP = px.gP * px.P;  // sum_j m_j * mask_grad_j
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) {
      const float2 l0 = s_ln[3 * k], l1 = s_ln[3 * k + 1], l2 = s_ln[3 * k + 2];
      size_t o = ((size_t)(b * K + k) * DD + pix);
      float mg = px.mask_grad(k), lo = px.leave_out(k), cf = px.xh_coef(k);
      float xh0 = cf * (px.v[k].x - px.x0), xh1 = cf * (px.v[k].y - px.x1), xh2 = cf * (px.v[k].z - px.x2);
      float4 o1 = make_float4(px.m[k], px.post_mask(k), (mg - l0.x) * l0.y * ga0 + be0, (lo - l1.x) * l1.y * ga2 + be2);
      float4 o2 = make_float4((xh0 - l2.x) * l2.y * ga3x + be3x, (xh1 - l2.x) * l2.y * ga3y + be3y,
                              (xh2 - l2.x) * l2.y * ga3z + be3z, 0.f);
      reinterpret_cast<float4*>(io.mix1)[o] = o1;
      reinterpret_cast<float4*>(io.mix2)[o] = o2;
      if (io.seed) reinterpret_cast<float4*>(io.seed)[o] = make_float4(xh0, xh1, xh2, px.m[k] * (mg - gPP));
    }
  const float2 lP = s_ln[3 * K];
  reinterpret_cast<float4*>(io.smp)[(size_t)b * DD + pix] =
      make_float4(px.x0, px.x1, px.x2, (px.P - lP.x) * lP.y * ga1 + be1);
}

// ---------------------------------------------------------------------------------------------------------
// Single-launch forward (statistics + losses + refinement inputs) for the sizes whose decode fits in shared memory.
// One thread-block CLUSTER of MIXF_CL CTAs owns one sample: every CTA stages its DD/MIXF_CL pixels of the K slot planes
// and of the target frame in shared memory with 1-D bulk copies (one elected thread, one mbarrier), so each input byte
// crosses HBM exactly once.  Pass 1 reads the staged pixels, writes masks / colours / final_recon and accumulates the
// LayerNorm2D sums (same warp-centred fp32 -> raw fp64 scheme as mix_stats_kernel); the per-CTA fp64 partials are
// exchanged through distributed shared memory (fixed rank order => deterministic, every CTA derives identical
// statistics); pass 2 re-reads the staged pixels from shared memory and writes the normalised refinement inputs and
// the gradient seed.  The loss scalars are reduced by rank 0 of each cluster; the last sample (self-resetting counter
// `sync`, zero before the first call, left zero by every call) sums the samples in index order.
// ---------------------------------------------------------------------------------------------------------
#ifndef OP3_MIXF_CL
#define OP3_MIXF_CL 8
#endif
#ifndef OP3_MIXF_THREADS
#define OP3_MIXF_THREADS 128
#endif
constexpr int MIXF_CL = OP3_MIXF_CL;
constexpr int MIXF_THREADS = OP3_MIXF_THREADS;
constexpr int MIXF_WARPS = MIXF_THREADS / 32;
constexpr int MIXF_NQMAX = 6 * 16 + 4;

constexpr int MIXF_MAXIT = 16;          // pixels per thread (= staging sub-chunks, one mbarrier each)
struct MixfSmem {                       // fixed part of the dynamic shared memory; the pixel staging area follows
  unsigned long long mbar[MIXF_MAXIT];
  double part[MIXF_NQMAX];              // this CTA's raw fp64 sums (read by the other CTAs of the cluster)
  double red[MIXF_WARPS][MIXF_NQMAX];
  float2 ln[3 * 16 + 1];
  int last;
};
__host__ __device__ inline size_t mixf_stage_offset() { return (sizeof(MixfSmem) + 127) & ~(size_t)127; }
// staging per pixel: K float4 (r,g,b,logit -> mask after pass 1) + 3 floats of the target + K floats of p_k
__host__ __device__ inline size_t mixf_smem_bytes(int K, int px) { return mixf_stage_offset() + (size_t)px * (20 * K + 12); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

#ifdef OP3_MIX_TRACE
__device__ unsigned long long g_mix_trace[16];      // per phase: summed cycles of thread 0 of every CTA; [15] = CTAs
#define MIXT(i) do { if (threadIdx.x == 0) { const long long _t = clock64(); atomicAdd(&g_mix_trace[i], (unsigned long long)(_t - mt_prev)); mt_prev = _t; } } while (0)
#else
#define MIXT(i) do { } while (0)
#endif
// EXACT: K == KMAX, so every `k < K` slot predicate folds away at compile time (~10 % of the instructions at K = 4)
template <int KMAX, bool EXACT>
__global__ void __cluster_dims__(MIXF_CL, 1, 1) __launch_bounds__(MIXF_THREADS)
mix_fused_kernel(Op3MixtureIO io, const float* __restrict__ mse_image, long long mse_bstride,
                 const float* __restrict__ g0, const float* __restrict__ g1, const float* __restrict__ g2,
                 const float* __restrict__ g3, const float* __restrict__ b0, const float* __restrict__ b1,
                 const float* __restrict__ b2, const float* __restrict__ b3, int B, int K_in, int D, int emit, int Rs,
                 float beta, unsigned* __restrict__ sync) {
  const int K = EXACT ? KMAX : K_in;
#ifdef OP3_MIX_TRACE
  long long mt_prev = clock64();
  if (threadIdx.x == 0) atomicAdd(&g_mix_trace[15], 1ull);
#endif
  extern __shared__ __align__(128) unsigned char mixf_raw[];
  MixfSmem& sm = *reinterpret_cast<MixfSmem*>(mixf_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int DD = D * D, PX = DD / MIXF_CL;
  const int b = blockIdx.y, rank = blockIdx.x;                 // cluster = the MIXF_CL CTAs of one sample
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  const int pix0 = rank * PX;
  const float invB = 1.f / (float)B;
  const int nq = 6 * K + 4;
  StatsLayout L = stats_layout(B, K, D);
  float4* s_dec = reinterpret_cast<float4*>(mixf_raw + mixf_stage_offset());          // [K][PX]
  float* s_img = reinterpret_cast<float*>(s_dec + (size_t)K * PX);                    // [3][PX]

  float* s_p = s_img + 3 * PX;                                                        // [K][PX]
  const int npix = PX / MIXF_THREADS;
  if (tid == 0) {
    for (int it = 0; it < npix; ++it) mbar_init(reinterpret_cast<uint64_t*>(&sm.mbar[it]), 1);
    fence_mbar_init();
  }
  __syncthreads();                                             // mbarriers initialised before anyone arms / polls them
  // One sub-chunk of MIXF_THREADS pixels per pass-1 iteration (compute starts when the first 1/npix of the bytes landed).  The
  // 3 + K bulk copies of every sub-chunk are issued by DIFFERENT lanes of warp 0 (one thread issuing all 28 of them took
  // 4800 cycles before pass 1 could start), the byte counts are armed by warp 1.
  if (warp == 1 && lane < npix) {
    const uint32_t sub = (uint32_t)MIXF_THREADS * (16u * K + 12u);
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n"
                 :: "r"(smem_u32(&sm.mbar[lane])), "r"(sub) : "memory");
  }
  if (warp == 0) {
    const float* img = io.image + (size_t)b * io.image_bstride + pix0;
    const int per = 3 + K;
    for (int e = lane; e < npix * per; e += 32) {
      const int it = e / per, c = e - it * per, o = it * MIXF_THREADS;
      unsigned long long* bar = &sm.mbar[it];
      if (c < 3) bulk_g2s(s_img + c * PX + o, img + (size_t)c * DD + o, MIXF_THREADS * 4, bar);
      else bulk_g2s(s_dec + (size_t)(c - 3) * PX + o, io.dec_out + ((size_t)(b * K + (c - 3)) * DD + pix0 + o) * 4, MIXF_THREADS * 16, bar);
    }
  }
  MIXT(0);                                                     // prologue: barrier init, copy issue, KL

  // ---------------- pass 1: outputs + statistics
  Pixel<KMAX> px;
  float acc[6 * KMAX], shift[3 * KMAX];
  float accP = 0.f, accP2 = 0.f, shiftP = 0.f, accLog = 0.f, accSq = 0.f;
#pragma unroll
  for (int i = 0; i < 6 * KMAX; ++i) acc[i] = 0.f;
#pragma unroll 1
  for (int it = 0; it < npix; ++it) {
    const int lp = it * MIXF_THREADS + tid, pix = pix0 + lp;
    mbar_wait(reinterpret_cast<uint64_t*>(&sm.mbar[it]), 0);
    if (it == 0) MIXT(1);                                      // first staged sub-chunk has landed
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) px.v[k] = s_dec[(size_t)k * PX + lp];
    px.x0 = s_img[lp]; px.x1 = s_img[PX + lp]; px.x2 = s_img[2 * PX + lp];
    px.compute(K, invB);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        if (emit) {                                              // pass 2 reuses the mask and the colour likelihood
          reinterpret_cast<float*>(s_dec + (size_t)k * PX + lp)[3] = px.m[k];
          s_p[(size_t)k * PX + lp] = px.p[k];
        }
        if (io.mask) io.mask[(size_t)(b * K + k) * DD + pix] = px.m[k];
        if (io.colors) {
          float* c = io.colors + (size_t)(b * K + k) * 3 * DD + pix;
          c[0] = px.v[k].x; c[DD] = px.v[k].y; c[2 * DD] = px.v[k].z;
        }
        r0 += px.v[k].x * px.m[k]; r1 += px.v[k].y * px.m[k]; r2 += px.v[k].z * px.m[k];
        if (emit) {
          float mg = px.mask_grad(k), lo = px.leave_out(k), cf = px.xh_coef(k);
          float x0 = cf * (px.v[k].x - px.x0), x1 = cf * (px.v[k].y - px.x1), x2 = cf * (px.v[k].z - px.x2);
          if (it == 0) {
            shift[3 * k] = __shfl_sync(0xffffffffu, mg, 0);
            shift[3 * k + 1] = __shfl_sync(0xffffffffu, lo, 0);
            shift[3 * k + 2] = __shfl_sync(0xffffffffu, x0, 0);
          }
          float d = mg - shift[3 * k];
          acc[6 * k] += d; acc[6 * k + 1] += d * d;
          d = lo - shift[3 * k + 1];
          acc[6 * k + 2] += d; acc[6 * k + 3] += d * d;
          float sh = shift[3 * k + 2];
          float e0 = x0 - sh, e1 = x1 - sh, e2 = x2 - sh;
          acc[6 * k + 4] += e0 + e1 + e2; acc[6 * k + 5] += e0 * e0 + e1 * e1 + e2 * e2;
        }
      }
    if (io.final_recon) {
      float* fr = io.final_recon + (size_t)b * 3 * DD + pix;
      fr[0] = r0; fr[DD] = r1; fr[2 * DD] = r2;
    }
    if (emit) {
      if (it == 0) shiftP = __shfl_sync(0xffffffffu, px.P, 0);
      float d = px.P - shiftP;
      accP += d; accP2 += d * d;
    }
    accLog += -logf(px.P + 1e-12f);
    if (mse_image) {
      const float* mi = mse_image + (size_t)b * mse_bstride;
      float q0 = r0 - __ldg(mi + pix), q1 = r1 - __ldg(mi + DD + pix), q2 = r2 - __ldg(mi + 2 * DD + pix);
      accSq += q0 * q0 + q1 * q1 + q2 * q2;
    }
  }
  MIXT(2);                                                     // pass 1
  const double n1 = 32.0 * npix, n3 = 96.0 * npix;
  if (emit) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          float S1 = warp_sum(acc[6 * k + 2 * g]), S2 = warp_sum(acc[6 * k + 2 * g + 1]);
          if (lane == 0) {
            double sft = (double)shift[3 * k + g], nn = g == 2 ? n3 : n1;
            sm.red[warp][6 * k + 2 * g] = (double)S1 + nn * sft;
            sm.red[warp][6 * k + 2 * g + 1] = (double)S2 + 2.0 * sft * (double)S1 + nn * sft * sft;
          }
        }
      }
  }
  {
    float S1 = warp_sum(accP), S2 = warp_sum(accP2), S3 = warp_sum(accLog), S4 = warp_sum(accSq);
    if (lane == 0) {
      double sft = (double)shiftP;
      sm.red[warp][6 * K] = (double)S1 + n1 * sft;
      sm.red[warp][6 * K + 1] = (double)S2 + 2.0 * sft * (double)S1 + n1 * sft * sft;
      sm.red[warp][6 * K + 2] = (double)S3;
      sm.red[warp][6 * K + 3] = (double)S4;
    }
  }
  __syncthreads();
  if (tid < nq && (emit || tid >= 6 * K)) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < MIXF_WARPS; ++w) s += sm.red[w][tid];
    sm.part[tid] = s;
  }
  MIXT(3);                                                     // warp / CTA reduction of the statistics
  cluster.barrier_arrive();                                    // this CTA's partials are written ...
  // KL of this sample (rank 0 only), between arrive and wait: the other seven CTAs of the cluster do not wait for it
  double kl = 0.0;
  if (rank == 0) {
    for (int i = tid; i < K * Rs; i += MIXF_THREADS) {
      size_t o = (size_t)b * K * Rs + i;
      float sq_ = softplus_std(io.post_l2[o]), spv = softplus_std(io.prior_l2[o]);
      float dm = io.post_l1[o] - io.prior_l1[o];
      kl += (double)(logf(spv / sq_) + (sq_ * sq_ + dm * dm) / (2.f * spv * spv) - 0.5f);
    }
    kl = warp_sum_d(kl);
  }
  cluster.barrier_wait();                                      // ... and every CTA's partials are visible cluster-wide
  MIXT(4);                                                     // cluster barrier

  // ---------------- cluster-wide statistics (fixed rank order), loss scalars
  if (emit && tid < 3 * K + 1) {
    const int q = tid, k = q / 3, g = q % 3;
    const int qi = q == 3 * K ? 6 * K : 6 * k + 2 * g;
    // all 2 x MIXF_CL remote reads in flight at once (a distributed-shared-memory load costs ~700 cycles; one rank per loop
    // iteration put 8 of them back to back on every CTA's critical path), then added in rank order: still deterministic
    double ps[MIXF_CL], ps2[MIXF_CL];
#pragma unroll
    for (int r = 0; r < MIXF_CL; ++r) {
      const double* rp = cluster.map_shared_rank(sm.part, r);
      ps[r] = rp[qi]; ps2[r] = rp[qi + 1];
    }
    double s = 0.0, s2 = 0.0;
#pragma unroll
    for (int r = 0; r < MIXF_CL; ++r) { s += ps[r]; s2 += ps2[r]; }
    const double n = (q != 3 * K && g == 2) ? 3.0 * DD : (double)DD;
    const double mean = s / n;
    double var = (s2 - s * s / n) / (n - 1.0);
    if (var < 0.0) var = 0.0;
    const float2 st = make_float2((float)mean, 1.f / ((float)sqrt(var) + 1e-5f));
    sm.ln[q] = st;
    if (rank == 0) {                                           // kept for the backward pass
      size_t o = q == 3 * K ? L.ln + (size_t)B * K * 6 + (size_t)b * 2 : L.ln + ((size_t)(b * K + k) * 3 + g) * 2;
      io.stats[o] = st.x; io.stats[o + 1] = st.y;
    }
  }
  if (rank == 0) {
    if (lane == 0) sm.red[warp][0] = kl;                       // red[][] is free again (all reads precede cluster.sync)
    __syncthreads();
    if (tid == 0) {
      double klv = 0.0, cl = 0.0, sq = 0.0;
#pragma unroll
      for (int w = 0; w < MIXF_WARPS; ++w) klv += sm.red[w][0];
      double pc[MIXF_CL], pq[MIXF_CL];
#pragma unroll
      for (int r = 0; r < MIXF_CL; ++r) {
        const double* rp = cluster.map_shared_rank(sm.part, r);
        pc[r] = rp[6 * K + 2]; pq[r] = rp[6 * K + 3];
      }
#pragma unroll
      for (int r = 0; r < MIXF_CL; ++r) { cl += pc[r]; sq += pq[r]; }
      double* samp = reinterpret_cast<double*>(io.stats + L.sample);
      samp[b * 3] = cl; samp[b * 3 + 1] = klv; samp[b * 3 + 2] = sq;
      __threadfence();
      if (atomicAdd(sync, 1u) == (unsigned)B - 1) {
        __threadfence();
        cl = 0.0; klv = 0.0; sq = 0.0;
        for (int i = 0; i < B; ++i) { cl += __ldcg(samp + i * 3); klv += __ldcg(samp + i * 3 + 1); sq += __ldcg(samp + i * 3 + 2); }
        const float clog = (float)(cl / B), klf = (float)(klv / B);
        io.losses[0] = clog;
        io.losses[1] = klf;
        io.losses[2] = clog + beta * klf;
        io.losses[3] = mse_image ? (float)(sq / ((double)B * 3.0 * D * D)) : 0.f;
        *sync = 0u;                                            // ready for the next call on this stream
      }
    }
  }
  cluster.barrier_arrive();                                    // this CTA no longer reads remote shared memory
  MIXT(5);                                                     // cluster-wide statistics, loss scalars
  if (emit) {
    __syncthreads();                                           // sm.ln
    MIXT(6);
    // ---------------- pass 2: refinement inputs + gradient seed from the staged pixels
    const float ga0 = __ldg(g0), be0 = __ldg(b0), ga1 = __ldg(g1), be1 = __ldg(b1), ga2 = __ldg(g2), be2 = __ldg(b2);
    const float ga3x = __ldg(g3), ga3y = __ldg(g3 + 1), ga3z = __ldg(g3 + 2);
    const float be3x = __ldg(b3), be3y = __ldg(b3 + 1), be3z = __ldg(b3 + 2);
    const float2 lP = sm.ln[3 * K];
#pragma unroll 1
    for (int it = 0; it < npix; ++it) {
      const int lp = it * MIXF_THREADS + tid, pix = pix0 + lp;
      const float x0 = s_img[lp], x1 = s_img[PX + lp], x2 = s_img[2 * PX + lp];
      float4 v[KMAX];                                            // (r, g, b, mask)
      float pk[KMAX];
      float P = 0.f, sp = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) {
          v[k] = s_dec[(size_t)k * PX + lp];
          pk[k] = s_p[(size_t)k * PX + lp];
          P += v[k].w * pk[k];
          sp += pk[k];
        }
      const float gP = (-invB) / (P + 1e-12f), rsp = 1.f / (sp + 1e-8f);
      const float gPP = gP * P;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) {
          const float2 l0 = sm.ln[3 * k], l1 = sm.ln[3 * k + 1], l2 = sm.ln[3 * k + 2];
          const size_t o = ((size_t)(b * K + k) * DD + pix);
          const float mk = v[k].w, mp = mk * pk[k];
          const float mg = gP * pk[k], lo = P - mp, cf = -gP * mk * pk[k] * MIX_INV_S2;
          const float xh0 = cf * (v[k].x - x0), xh1 = cf * (v[k].y - x1), xh2 = cf * (v[k].z - x2);
          reinterpret_cast<float4*>(io.mix1)[o] =
              make_float4(mk, pk[k] * rsp, (mg - l0.x) * l0.y * ga0 + be0, (lo - l1.x) * l1.y * ga2 + be2);
          reinterpret_cast<float4*>(io.mix2)[o] = make_float4((xh0 - l2.x) * l2.y * ga3x + be3x, (xh1 - l2.x) * l2.y * ga3y + be3y,
                                                              (xh2 - l2.x) * l2.y * ga3z + be3z, 0.f);
          if (io.seed) reinterpret_cast<float4*>(io.seed)[o] = make_float4(xh0, xh1, xh2, mk * (mg - gPP));
        }
      reinterpret_cast<float4*>(io.smp)[(size_t)b * DD + pix] = make_float4(x0, x1, x2, (P - lP.x) * lP.y * ga1 + be1);
    }
  }
  MIXT(7);                                                     // pass 2
  cluster.barrier_wait();                                      // nobody exits while a peer may still read its partials
  MIXT(8);
}

// ---------------------------------------------------------------------------------------------------------
// backward w.r.t. dec_out (+ LayerNorm2D gamma/beta gradient partials)
// ---------------------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(MIX_THREADS)
mix_bwd_kernel(Op3MixtureIO io, float w_clog, const float4* __restrict__ d_in, float4* __restrict__ d_dec,
               double* __restrict__ gpart, int B, int K, int D) {
  const int DD = D * D;
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  const int pix = blockIdx.x * MIX_THREADS + tid;
  const bool live = pix < DD;
  const float invB = 1.f / (float)B;
  StatsLayout L = stats_layout(B, K, D);
  float lg[12];  // dgamma0,dbeta0, dgamma1,dbeta1, dgamma2,dbeta2, dgamma3 rgb, dbeta3 rgb
#pragma unroll
  for (int i = 0; i < 12; ++i) lg[i] = 0.f;
  if (live) {
    const float* img = io.image + (size_t)b * io.image_bstride;
    Pixel<KMAX> px;
    px.load(io.dec_out, img, b, K, DD, pix);
    px.compute(K, invB);
    float gm[KMAX], dpm[KMAX];
    float4 u0[KMAX];
    float sum_mg = 0.f, sum_dpm_p = 0.f;
    const float* lnP = io.stats + L.ln + (size_t)B * K * 6 + (size_t)b * 2;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        float4 u1 = make_float4(0.f, 0.f, 0.f, 0.f), u2 = u1, u3 = u1;
        u0[k] = u1;
        if (d_in) {
          size_t base = (size_t)(b * K + k) * 5 * DD + pix;
          u0[k] = d_in[base]; u1 = d_in[base + DD]; u2 = d_in[base + 2 * (size_t)DD]; u3 = d_in[base + 3 * (size_t)DD];
          const float* ln = io.stats + L.ln + (size_t)(b * K + k) * 6;
          float cf = px.xh_coef(k);
          lg[0] += u1.z * (px.mask_grad(k) - ln[0]) * ln[1]; lg[1] += u1.z;
          lg[2] += u3.w * (px.P - lnP[0]) * lnP[1];          lg[3] += u3.w;
          lg[4] += u1.w * (px.leave_out(k) - ln[2]) * ln[3]; lg[5] += u1.w;
          lg[6] += u2.x * (cf * (px.v[k].x - px.x0) - ln[4]) * ln[5]; lg[9] += u2.x;
          lg[7] += u2.y * (cf * (px.v[k].y - px.x1) - ln[4]) * ln[5]; lg[10] += u2.y;
          lg[8] += u2.z * (cf * (px.v[k].z - px.x2) - ln[4]) * ln[5]; lg[11] += u2.z;
        }
        gm[k] = u1.x + w_clog * px.mask_grad(k);
        dpm[k] = u1.y;
        sum_mg += px.m[k] * gm[k];
        sum_dpm_p += dpm[k] * px.p[k];
      }
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        float dp = (dpm[k] - sum_dpm_p * px.rsp) * px.rsp;     // through posterior_mask = p/(sum p + 1e-8)
        float cc = w_clog * px.xh_coef(k) - dp * px.p[k] * MIX_INV_S2;
        float4 o;
        o.x = u0[k].x + cc * (px.v[k].x - px.x0);
        o.y = u0[k].y + cc * (px.v[k].y - px.x1);
        o.z = u0[k].z + cc * (px.v[k].z - px.x2);
        o.w = u0[k].w + px.m[k] * (gm[k] - sum_mg);
        d_dec[(size_t)(b * K + k) * DD + pix] = o;
      }
  }
  if (d_in && gpart) {
    __shared__ float s_red[MIX_THREADS / 32][12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      float s = warp_sum(lg[i]);
      if (lane == 0) s_red[warp][i] = s;
    }
    __syncthreads();
    if (tid < 12) {
      double s = 0.0;
      for (int w = 0; w < MIX_THREADS / 32; ++w) s += (double)s_red[w][tid];
      gpart[((size_t)b * gridDim.x + blockIdx.x) * 12 + tid] = s;
    }
  }
}

// one warp per quantity (12 warps); lanes stride over the per-CTA partials, fixed shuffle tree => deterministic
__global__ void __launch_bounds__(384) mix_bwd_reduce_kernel(const double* __restrict__ gpart, int nblocks, Op3Params grads) {
  const int i = threadIdx.x / 32, lane = threadIdx.x % 32;
  double s = 0.0;
  for (int k = lane; k < nblocks; k += 32) s += gpart[(size_t)k * 12 + i];
  s = warp_sum_d(s);
  if (lane != 0) return;
  float v = (float)s;
  switch (i) {
    case 0: grads.ln2d_gamma[0][0] += v; break;
    case 1: grads.ln2d_beta[0][0] += v; break;
    case 2: grads.ln2d_gamma[1][0] += v; break;
    case 3: grads.ln2d_beta[1][0] += v; break;
    case 4: grads.ln2d_gamma[2][0] += v; break;
    case 5: grads.ln2d_beta[2][0] += v; break;
    case 6: case 7: case 8: grads.ln2d_gamma[3][i - 6] += v; break;
    default: grads.ln2d_beta[3][i - 9] += v; break;
  }
}

}  // namespace
}  // namespace op3

using namespace op3;

#ifdef OP3_MIX_TRACE
extern "C" int op3_debug_mix_trace(unsigned long long* out /*[16] host*/, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_mix_trace, sizeof(unsigned long long) * 16);
  if (reset) { static unsigned long long z[16]; cudaMemcpyToSymbol(g_mix_trace, z, sizeof(z)); }
  return 0;
}
#endif

extern "C" size_t op3_mixture_stats_floats(const Op3Dims* d) {
  if (validate_dims(d) != OP3_OK) return 0;
  return stats_layout(d->B, d->K, d->D).total;
}

// ---------------------------------------------------------------------------------------------------------
// Standalone OP3Model.get_loss (op3_model.py:313-327) on caller-provided NCHW colours and post-softmax masks.
// ---------------------------------------------------------------------------------------------------------
namespace op3 {
namespace {

__global__ void __launch_bounds__(MIX_THREADS) get_loss_pixels_kernel(int K, int DD, const float* __restrict__ colors,
                                                                      const float* __restrict__ mask,
                                                                      const float* __restrict__ image,
                                                                      float* __restrict__ color_probs,
                                                                      float* __restrict__ pixel_ll, double* __restrict__ partial) {
  const int b = blockIdx.y, pix = blockIdx.x * MIX_THREADS + threadIdx.x;
  __shared__ double s_red[MIX_THREADS];
  double nl = 0.0;
  if (pix < DD) {
    const float* img = image + (size_t)b * 3 * DD;
    const float x0 = __ldg(img + pix), x1 = __ldg(img + DD + pix), x2 = __ldg(img + 2 * DD + pix);
    float P = 0.f;
    for (int k = 0; k < K; ++k) {
      const float* c = colors + (size_t)(b * K + k) * 3 * DD;
      const float d0 = __ldg(c + pix) - x0, d1 = __ldg(c + DD + pix) - x1, d2 = __ldg(c + 2 * DD + pix) - x2;
      const float p = expf(-(d0 * d0 + d1 * d1 + d2 * d2) / 0.06f) / 0.24805f;
      color_probs[(size_t)(b * K + k) * DD + pix] = p;
      P += __ldg(mask + (size_t)(b * K + k) * DD + pix) * p;
    }
    pixel_ll[(size_t)b * DD + pix] = P;
    nl = -(double)logf(P + 1e-12f);
  }
  s_red[threadIdx.x] = nl;
  __syncthreads();
  for (int o = MIX_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)b * gridDim.x + blockIdx.x] = s_red[0];
}

__global__ void __launch_bounds__(256) get_loss_final_kernel(int B, int n_kl, int n_part, const double* __restrict__ partial,
                                                             const float* __restrict__ post_l1, const float* __restrict__ post_l2,
                                                             const float* __restrict__ prior_l1,
                                                             const float* __restrict__ prior_l2, float beta,
                                                             float* __restrict__ losses) {
  __shared__ double s_kl[256], s_cl[256];
  const int tid = threadIdx.x;
  double kl = 0.0, cl = 0.0;
  for (int i = tid; i < n_kl; i += 256) {
    const float sq_ = softplus_std(post_l2[i]), spv = softplus_std(prior_l2[i]);
    const float dm = post_l1[i] - prior_l1[i];
    kl += (double)(logf(spv / sq_) + (sq_ * sq_ + dm * dm) / (2.f * spv * spv) - 0.5f);
  }
  for (int i = tid; i < n_part; i += 256) cl += partial[i];
  s_kl[tid] = kl; s_cl[tid] = cl;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s_kl[tid] += s_kl[tid + o]; s_cl[tid] += s_cl[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    const float klf = (float)(s_kl[0] / B), clog = (float)(s_cl[0] / B);
    losses[0] = klf;
    losses[1] = clog;
    losses[2] = clog + beta * klf;
  }
}

}  // namespace
}  // namespace op3

extern "C" size_t op3_get_loss_scratch_floats(const Op3Dims* d) {
  if (validate_dims(d) != OP3_OK) return 0;
  return 2 * (size_t)d->B * ((d->D * d->D + MIX_THREADS - 1) / MIX_THREADS);
}

extern "C" int op3_get_loss(const Op3Dims* d, const float* colors, const float* mask, const float* image,
                            const float* post_l1, const float* post_l2, const float* prior_l1, const float* prior_l2,
                            float* color_probs, float* pixel_ll, float* losses, float* scratch, void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(colors && mask && image && post_l1 && post_l2 && prior_l1 && prior_l2 && color_probs && pixel_ll && losses && scratch,
              "op3_get_loss: NULL argument");
  OP3_REQUIRE(((uintptr_t)scratch & 7) == 0, "op3_get_loss: scratch must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int DD = d->D * d->D, tiles = (DD + MIX_THREADS - 1) / MIX_THREADS;
  ProfScope prof(st, KID_MIXTURE, (double)d->B * DD * (6.0 * d->K + 4.0) * 4.0);
  double* partial = reinterpret_cast<double*>(scratch);
  get_loss_pixels_kernel<<<dim3(tiles, d->B), MIX_THREADS, 0, st>>>(d->K, DD, colors, mask, image, color_probs, pixel_ll, partial);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  get_loss_final_kernel<<<1, 256, 0, st>>>(d->B, d->B * d->K * d->Rs, d->B * tiles, partial, post_l1, post_l2, prior_l1, prior_l2,
                                           d->beta, losses);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

extern "C" size_t op3_mixture_bwd_scratch_floats(const Op3Dims* d) { (void)d; return 0; }

// `want_loss`: bit 0 = compute LN statistics + losses (needed for loss and/or refine inputs); bit 1 = losses[3] is the
// reconstruction mse against io->mse_image.
extern "C" int op3_mixture_fwd(const Op3Dims* d, const Op3Params* p, const Op3MixtureIO* io, int want_loss, void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(io && io->dec_out, "op3_mixture_fwd: NULL io");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, K = d->K, D = d->D;
  const bool emit = io->mix1 != nullptr;
  const int want_stats = (want_loss & 1) || emit;
  if (want_stats)
    OP3_REQUIRE(io->image && io->stats && io->losses && io->post_l1 && io->post_l2 && io->prior_l1 && io->prior_l2,
                "op3_mixture_fwd: statistics requested but image/stats/losses/lambda pointers missing");
  if (emit) OP3_REQUIRE(p && io->mix2 && io->smp, "op3_mixture_fwd: refine inputs requested but mix2/smp/params missing");
  StatsLayout L = stats_layout(B, K, D);
  // algorithmic bytes: full forward (11K+4)*4 per pixel-sample, outputs-only (8K+3... see DESIGN.md)
  ProfScope prof(st, KID_MIXTURE, (emit ? (11.0 * K + 4) : (8.0 * K + 6)) * 4.0 * B * D * D);
  // single-launch cluster kernel whenever one sample's decode fits the shared memory of MIXF_CL CTAs
  const int DD = D * D, PX = DD / MIXF_CL;
  const size_t fused_smem = mixf_smem_bytes(K, PX);
  if (want_stats && io->sync && DD % (MIXF_CL * MIXF_THREADS) == 0 && PX / MIXF_THREADS <= MIXF_MAXIT && fused_smem <= 200 * 1024 &&
      ((uintptr_t)io->image & 15) == 0 && (io->image_bstride & 3) == 0 && ((uintptr_t)io->dec_out & 15) == 0) {
    dim3 gf(MIXF_CL, B);
    const float* nul = nullptr;
#define OP3_MIX_FUSED(KM, EX)                                                                                          \
  do {                                                                                                                 \
    static PerDeviceOnce attr_once;                                                                                    \
    if (attr_once.first()) {                                                                                           \
      OP3_CHECK(cudaFuncSetAttribute(mix_fused_kernel<KM, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    }                                                                                                                  \
    mix_fused_kernel<KM, EX><<<gf, MIXF_THREADS, fused_smem, st>>>(                                                    \
        *io, io->mse_image, io->mse_image_bstride, emit ? p->ln2d_gamma[0] : nul, emit ? p->ln2d_gamma[1] : nul,       \
        emit ? p->ln2d_gamma[2] : nul, emit ? p->ln2d_gamma[3] : nul, emit ? p->ln2d_beta[0] : nul,                    \
        emit ? p->ln2d_beta[1] : nul, emit ? p->ln2d_beta[2] : nul, emit ? p->ln2d_beta[3] : nul, B, K, D,             \
        emit ? 1 : 0, d->Rs, d->beta, io->sync);                                                                       \
  } while (0)
    if (K == 4) OP3_MIX_FUSED(4, true); else if (K < 4) OP3_MIX_FUSED(4, false);
    else if (K == 8) OP3_MIX_FUSED(8, true); else if (K < 8) OP3_MIX_FUSED(8, false);
    else OP3_MIX_FUSED(16, false);
#undef OP3_MIX_FUSED
    OP3_COUNT_LAUNCH();
    OP3_LAUNCH_CHECK();
    return OP3_OK;
  }
  dim3 gs(L.tiles, B);
  if (want_stats) OP3_CHECK(cudaMemsetAsync(io->stats + L.counter, 0, (size_t)(B + 1) * sizeof(unsigned), st));
#define OP3_MIX_STATS(KM) \
  mix_stats_kernel<KM><<<gs, MIX_THREADS, 0, st>>>(*io, io->mse_image, io->mse_image_bstride, B, K, D, want_stats, d->Rs, d->beta)
  if (K <= 4) OP3_MIX_STATS(4); else if (K <= 8) OP3_MIX_STATS(8); else OP3_MIX_STATS(16);
#undef OP3_MIX_STATS
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  if (!want_stats) return OP3_OK;
  if (emit) {
    dim3 ge(ceil_div(D * D, MIX_THREADS), B);
#define OP3_MIX_EMIT(KM)                                                                                         \
  mix_emit_kernel<KM><<<ge, MIX_THREADS, 0, st>>>(*io, p->ln2d_gamma[0], p->ln2d_gamma[1], p->ln2d_gamma[2],    \
                                                  p->ln2d_gamma[3], p->ln2d_beta[0], p->ln2d_beta[1],           \
                                                  p->ln2d_beta[2], p->ln2d_beta[3], B, K, D)
    if (K <= 4) OP3_MIX_EMIT(4); else if (K <= 8) OP3_MIX_EMIT(8); else OP3_MIX_EMIT(16);
#undef OP3_MIX_EMIT
    OP3_COUNT_LAUNCH();
    OP3_LAUNCH_CHECK();
  }
  return OP3_OK;
}

extern "C" int op3_mixture_bwd(const Op3Dims* d, const Op3MixtureIO* io, float w_clog, const float* d_in,
                               float* d_dec_out, const Op3Params* grads, float* scratch, void* stream) {
  (void)scratch;
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(io && io->dec_out && io->image && io->stats && d_dec_out, "op3_mixture_bwd: NULL argument");
  OP3_REQUIRE(!d_in || grads, "op3_mixture_bwd: grads required when d_in is given");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, K = d->K, D = d->D;
  StatsLayout L = stats_layout(B, K, D);
  ProfScope prof(st, KID_MIXTURE, (14.0 * K + 3) * 4.0 * B * D * D);
  dim3 g(ceil_div(D * D, MIX_THREADS), B);
  double* gpart = reinterpret_cast<double*>(io->stats + L.gpart);
#define OP3_MIX_BWD(KM)                                                                                  \
  mix_bwd_kernel<KM><<<g, MIX_THREADS, 0, st>>>(*io, w_clog, reinterpret_cast<const float4*>(d_in),       \
                                                reinterpret_cast<float4*>(d_dec_out), gpart, B, K, D)
  if (K <= 4) OP3_MIX_BWD(4); else if (K <= 8) OP3_MIX_BWD(8); else OP3_MIX_BWD(16);
#undef OP3_MIX_BWD
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  if (d_in) {
    mix_bwd_reduce_kernel<<<1, 384, 0, st>>>(gpart, (int)(g.x * B), *grads);
    OP3_COUNT_LAUNCH();
    OP3_LAUNCH_CHECK();
  }
  return OP3_OK;
}

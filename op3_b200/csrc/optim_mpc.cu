// Trainer-step and MPC helper kernels.
//  * op3_adam_clip replaces torch.nn.utils.clip_grad_norm_(params, 5.0) + optim.Adam.step (op3/torch/op3_modules/
//    op3_trainer.py:136,188-189) with one pass over flat parameter / gradient / moment buffers.
//  * op3_mpc_cost replaces Cost.image_mse + min over K (exps/stack_exps/mpc_stack.py:103-127,160-184,225-228).
#include "common.cuh"

namespace op3 {
namespace {

constexpr int NORM_BLOCKS = 512;

__global__ void __launch_bounds__(256) sqnorm_kernel(long long n, const float* __restrict__ g, float scale,
                                                     double* __restrict__ partial) {
  __shared__ double red[256];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float v = g[i] * scale;
    s += (double)v * (double)v;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float scale, float lr,
                                                   float b1, float b2, float eps, float bc1, float bc2_sqrt, float max_norm,
                                                   const double* __restrict__ partial, int nparts, float* __restrict__ norm_out) {
  __shared__ float s_coef;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nparts; ++i) t += partial[i];
    float total = (float)sqrt(t);
    float coef = max_norm / (total + 1e-6f);
    s_coef = coef < 1.f ? coef : 1.f;
    if (blockIdx.x == 0) norm_out[0] = total;
  }
  __syncthreads();
  const float coef = s_coef * scale;
  const float step = lr / bc1;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float gi = g[i] * coef;
    float mi = m[i] + (gi - m[i]) * (1.f - b1);          // torch: exp_avg.lerp_(grad, 1 - beta1)
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step * (mi / denom);
  }
}

// sums[(a*Kp+k)*G + g] = distance of goal g and prediction (a,k) over n_el elements, before the mean; one CTA per (a,k);
// fixed-order reduction.  CMP 0: sum of squared differences (Cost.image_mse / Cost.mse, mpc_stack.py:103-127).
// CMP 1: 'psuedo_intersect' (mpc_stack.py:84-92): 1 - |{g > .01, p > .01, (g-p)^2 < .08}| / (|{g > .01}| + |{p > .01}|).
template <int GMAX, int CMP>
__global__ void __launch_bounds__(256) mpc_sqerr_kernel(int G, int n_el, const float* __restrict__ pred,
                                                        const float* __restrict__ goal, double* __restrict__ sums) {
  __shared__ double red[256];
  const float* pp = pred + (size_t)blockIdx.x * n_el;
  float acc[GMAX], cg[GMAX];
  float cp = 0.f;
#pragma unroll
  for (int g = 0; g < GMAX; ++g) { acc[g] = 0.f; cg[g] = 0.f; }
  for (int i = threadIdx.x; i < n_el; i += 256) {
    float pv = pp[i];
    if (CMP == 1) cp += pv > 0.01f ? 1.f : 0.f;
#pragma unroll
    for (int g = 0; g < GMAX; ++g)
      if (g < G) {
        const float gv = __ldg(goal + (size_t)g * n_el + i);
        const float dlt = gv - pv;
        if (CMP == 0) {
          acc[g] = fmaf(dlt, dlt, acc[g]);
        } else {
          const bool m1 = gv > 0.01f;
          cg[g] += m1 ? 1.f : 0.f;
          acc[g] += (m1 && pv > 0.01f && dlt * dlt < 0.08f) ? 1.f : 0.f;
        }
      }
  }
  auto block_sum = [&](double mine) {
    red[threadIdx.x] = mine;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  const double pcount = CMP == 1 ? block_sum((double)cp) : 0.0;
  for (int g = 0; g < G; ++g) {
    double mine = 0.0, mg = 0.0;
#pragma unroll
    for (int q = 0; q < GMAX; ++q)
      if (q == g) { mine = (double)acc[q]; mg = (double)cg[q]; }
    const double a = block_sum(mine);
    if (CMP == 0) {
      if (threadIdx.x == 0) sums[(size_t)blockIdx.x * G + g] = a;
    } else {
      const double gcount = block_sum(mg);
      // counts are small integers: the float division below is the reference's (0/0 = NaN when both images are empty)
      if (threadIdx.x == 0) sums[(size_t)blockIdx.x * G + g] = (double)(1.f - (float)a / ((float)gcount + (float)pcount));
    }
  }
}

// min over the Kp predictions of post(dist): post 0 = raw, 1 = -exp(-dist) ('negative_exp', mpc_stack.py:94-100)
__global__ void mpc_min_kernel(int Na, int Kp, int G, double inv_n, int post, const double* __restrict__ sums,
                               float* __restrict__ cost, int* __restrict__ argk) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Na * G) return;
  int a = i % Na, g = i / Na;
  float best = 0.f;
  int bk = 0;
  for (int k = 0; k < Kp; ++k) {
    float c = (float)(sums[((size_t)a * Kp + k) * G + g] * inv_n);
    if (post == 1) c = -expf(-c);
    if (k == 0 || c < best) { best = c; bk = k; }   // ties -> lowest k, like torch.min
  }
  cost[(size_t)g * Na + a] = best;
  if (argk) argk[(size_t)g * Na + a] = bk;
}

__global__ void sub_images_kernel(long long n_slots, int DD, const float* __restrict__ colors, const float* __restrict__ mask,
                                  float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots * 3 * DD) return;
  long long s = i / (3LL * DD);
  int pix = (int)(i % DD);
  out[i] = colors[i] * mask[s * DD + pix];
}

// CEM refit: one thread per action coordinate walks the F elite rows in rank order (fp64, fixed order => the refit is
// bit-reproducible and identical on every rank of a sharded rollout).
__global__ void cem_refit_kernel(int TA, int F, const float* __restrict__ actions, const long long* __restrict__ order,
                                 float min_std, float* __restrict__ mean, float* __restrict__ stdv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= TA) return;
  double s = 0.0;
  for (int f = 0; f < F; ++f) s += (double)actions[(size_t)order[f] * TA + j];
  const double m = s / F;
  double v = 0.0;
  for (int f = 0; f < F; ++f) {
    const double dlt = (double)actions[(size_t)order[f] * TA + j] - m;
    v += dlt * dlt;
  }
  mean[j] = (float)m;
  stdv[j] = (float)(sqrt(v / F) + (double)min_std);          // numpy .std(): population std (ddof = 0)
}

__global__ void cem_resample_kernel(long long n, int TA, const float* __restrict__ mean, const float* __restrict__ stdv,
                                    const float* __restrict__ noise, const float* __restrict__ lo,
                                    const float* __restrict__ hi, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = (int)(i % TA);
  float v = fmaf(stdv[j], noise[i], mean[j]);
  if (lo) v = fmaxf(v, lo[j]);
  if (hi) v = fminf(v, hi[j]);
  out[i] = v;
}

// uint8 frames -> planar fp32 in [0,1] (op3_trainer.py:148-153 `/ 255.0`; load_dataset's NHWC -> NCHW moveaxis,
// train_op3.py:27-28,46-47).  One thread per 4 output floats of one colour plane; 4-byte loads in the planar layout.
__global__ void __launch_bounds__(256) frames_u8_kernel(long long n4, int DD, int channels_last,
                                                        const unsigned char* __restrict__ in, float4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // index of a float4 of the planar output
  if (i >= n4) return;
  float4 v;
  if (!channels_last) {
    const uchar4 u = reinterpret_cast<const uchar4*>(in)[i];
    v = make_float4(u.x / 255.f, u.y / 255.f, u.z / 255.f, u.w / 255.f);
  } else {
    const long long e = i * 4, plane = e / DD;             // plane = frame*3 + c
    const int pix = (int)(e % DD), c = (int)(plane % 3);
    const unsigned char* src = in + ((plane / 3) * DD + pix) * 3 + c;
    v = make_float4(src[0] / 255.f, src[3] / 255.f, src[6] / 255.f, src[9] / 255.f);
  }
  out[i] = v;
}

}  // namespace
}  // namespace op3

using namespace op3;

extern "C" int op3_prepare_frames(long long n_frames, int D, int channels_last, const unsigned char* in, float* out,
                                  void* stream) {
  OP3_REQUIRE(n_frames > 0 && D > 0 && D % 2 == 0 && in && out, "op3_prepare_frames: bad argument (D must be even)");
  OP3_REQUIRE(((uintptr_t)in & 3) == 0 && ((uintptr_t)out & 15) == 0, "op3_prepare_frames: in must be 4-byte, out 16-byte aligned");
  const long long n4 = n_frames * 3 * D * D / 4;
  frames_u8_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n4, D * D, channels_last, in,
                                                                                    reinterpret_cast<float4*>(out));
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

extern "C" int op3_cem_refit(int Na, int TA, int F, const float* actions, const long long* order, float min_std,
                             float* mean, float* stdv, void* stream) {
  OP3_REQUIRE(Na > 0 && TA > 0 && F > 0 && F <= Na && actions && order && mean && stdv, "op3_cem_refit: bad argument");
  cem_refit_kernel<<<ceil_div(TA, 64), 64, 0, (cudaStream_t)stream>>>(TA, F, actions, order, min_std, mean, stdv);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

extern "C" int op3_cem_resample(int Na, int TA, const float* mean, const float* stdv, const float* noise, const float* lo,
                                const float* hi, float* out, void* stream) {
  OP3_REQUIRE(Na > 0 && TA > 0 && mean && stdv && noise && out, "op3_cem_resample: bad argument");
  const long long n = (long long)Na * TA;
  cem_resample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, TA, mean, stdv, noise, lo, hi, out);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

extern "C" int op3_adam_clip(long long n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int step,
                             float lr, float beta1, float beta2, float eps, float max_norm, float grad_scale,
                             float* scratch, void* stream) {
  OP3_REQUIRE(n > 0 && params && grads && exp_avg && exp_avg_sq && scratch && step >= 1, "op3_adam_clip: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = reinterpret_cast<double*>(scratch + 2);  // scratch[0] = norm; doubles 8-byte aligned at +2 floats
  int blocks = (int)((n + 255) / 256);
  if (blocks > NORM_BLOCKS) blocks = NORM_BLOCKS;
  sqnorm_kernel<<<blocks, 256, 0, st>>>(n, grads, grad_scale, partial);
  OP3_COUNT_LAUNCH();
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  int ab = (int)((n + 255) / 256);
  if (ab > 1184) ab = 1184;
  adam_kernel<<<ab, 256, 0, st>>>(n, params, grads, exp_avg, exp_avg_sq, grad_scale, lr, beta1, beta2, eps, bc1, bc2s,
                                  max_norm, partial, blocks, scratch);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

extern "C" int op3_mpc_cost_ex(int Na, int Kp, int G, int n_el, int compare, int post, const float* pred, const float* goal,
                               float* cost, int* argk, void* stream) {
  OP3_REQUIRE(Na > 0 && Kp > 0 && G > 0 && G <= 16 && n_el > 0 && pred && goal && cost, "op3_mpc_cost: bad argument (G<=16)");
  OP3_REQUIRE((compare == 0 || compare == 1) && (post == 0 || post == 1), "op3_mpc_cost: compare %d / post %d not in {0,1}", compare, post);
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = nullptr;
  OP3_CHECK(cudaMallocAsync((void**)&sums, (size_t)Na * Kp * G * sizeof(double), st));  // stream-ordered temp, freed below
  if (compare == 0) {
    if (G <= 4) mpc_sqerr_kernel<4, 0><<<Na * Kp, 256, 0, st>>>(G, n_el, pred, goal, sums);
    else mpc_sqerr_kernel<16, 0><<<Na * Kp, 256, 0, st>>>(G, n_el, pred, goal, sums);
  } else {
    if (G <= 4) mpc_sqerr_kernel<4, 1><<<Na * Kp, 256, 0, st>>>(G, n_el, pred, goal, sums);
    else mpc_sqerr_kernel<16, 1><<<Na * Kp, 256, 0, st>>>(G, n_el, pred, goal, sums);
  }
  OP3_COUNT_LAUNCH();
  mpc_min_kernel<<<ceil_div(Na * G, 256), 256, 0, st>>>(Na, Kp, G, compare == 0 ? 1.0 / (double)n_el : 1.0, post, sums, cost, argk);
  OP3_COUNT_LAUNCH();
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(sums, st);
  OP3_CHECK(e);
  return OP3_OK;
}

extern "C" int op3_mpc_cost(int Na, int Kp, int G, int D, const float* pred, const float* goal, float* cost, int* argk,
                            void* stream) {
  OP3_REQUIRE(D > 0, "op3_mpc_cost: bad argument");
  return op3_mpc_cost_ex(Na, Kp, G, 3 * D * D, 0, 0, pred, goal, cost, argk, stream);
}

extern "C" int op3_sub_images(long long n_slots, int D, const float* colors, const float* mask, float* out, void* stream) {
  OP3_REQUIRE(n_slots > 0 && D > 0 && colors && mask && out, "op3_sub_images: bad argument");
  long long tot = n_slots * 3 * D * D;
  sub_images_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_slots, D * D, colors, mask, out);
  OP3_COUNT_LAUNCH();
  OP3_LAUNCH_CHECK();
  return OP3_OK;
}

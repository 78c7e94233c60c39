// Shared device/host helpers for the op3_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/op3_b200.h"

namespace op3 {

// ---- error plumbing (C ABI never throws; see op3_last_error) ---------------------------------
void set_error(const char* fmt, ...);
#define OP3_CHECK(expr)                                                                      \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      op3::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return OP3_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)
#define OP3_LAUNCH_CHECK() OP3_CHECK(cudaGetLastError())
#define OP3_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      op3::set_error(__VA_ARGS__);      \
      return OP3_ERR_ARG;               \
    }                                   \
  } while (0)
#define OP3_TRY(call)            \
  do {                           \
    int _rc = (call);            \
    if (_rc != OP3_OK) return _rc; \
  } while (0)

// ---- launch accounting (bench.py reports gpu_launches from this) ------------------------------
extern unsigned long long g_launch_count;
#define OP3_COUNT_LAUNCH() (++op3::g_launch_count)

// ---- optional per-kernel-family timing (op3_profile_enable / op3_profile_collect), CUDA events on the launch stream
enum { KID_CONV_FWD = 0, KID_CONV_WGRAD = 1, KID_MIXTURE = 2, KID_GEMM = 3, KID_COUNT = 4 };
extern bool g_profile_on;
void profile_begin(cudaStream_t st, int kid, double work);
void profile_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st; bool on;
  ProfScope(cudaStream_t s, int kid, double work) : st(s), on(g_profile_on) { if (on) profile_begin(st, kid, work); }
  ~ProfScope() { if (on) profile_end(st); }
};

// fused row-tile MLP chains (fused_mlp.cuh) for the dynamics network and the refinement head; op3_set_fused_mlp(0) selects the
// layer-by-layer launches (kept for A/B tests and for parameter buffers whose rows are not 16-byte aligned)
bool fused_mlp_enabled();
// quint-tap conv kernel as CTA pairs (tcgen05 cta_group::2); op3_set_tcq_pair(0) selects the single-CTA form
bool tcq_pair_enabled();

// weight-gradient lane (api.cu): the stream a conv weight-gradient launch of a backward call goes to -- `st` itself unless
// op3_set_wgrad_overlap(1), then the per-device side stream, ordered after everything issued to `st` so far
int wgrad_stream(cudaStream_t st, cudaStream_t* out);
bool wgrad_overlap_enabled();

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }
// derivative of ELU(alpha=1) expressed through its OUTPUT a: a>0 -> 1, else a+1 (= e^x)
__device__ __forceinline__ float elu_grad_from_out(float a) { return a > 0.f ? 1.f : a + 1.f; }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- per-device host-side caches: cudaFuncSetAttribute and the SM count belong to the CURRENT device, not the process
constexpr int OP3_MAX_DEVICES = 64;
struct PerDeviceOnce {          // `static PerDeviceOnce once; if (once.first()) cudaFuncSetAttribute(...)`
  bool done[OP3_MAX_DEVICES] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= OP3_MAX_DEVICES) return true;
    const bool f = !done[dev];
    done[dev] = true;
    return f;
  }
};
inline int device_sm_count() {  // SMs of the current device (cached per device)
  static int sms[OP3_MAX_DEVICES] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= OP3_MAX_DEVICES) return 148;
  if (sms[dev] == 0 && cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms[dev] = 148;
  return sms[dev];
}

// ---- internal kernels shared between translation units ----------------------------------------
// gemm.cu ------------------------------------------------------------------------------------
enum { ACT_NONE = 0, ACT_ELU = 1, ACT_SIGMOID = 2, ACT_RELU = 3, ACT_TANH = 4 };
// C[M,N] = act( opA(A)[M,K] * opB(B)[K,N] + bias[N] ) (+ C if accumulate)
//   transA=0: A is row-major [M,K] (lda); transA=1: A is stored [K,M] (lda)
//   transB=0: B is stored [K,N] (ldb);    transB=1: B is stored [N,K] (ldb)  (nn.Linear weight)
int gemm(cudaStream_t st, int M, int N, int K, const float* A, int lda, int transA, const float* B, int ldb,
         int transB, float* C, int ldc, const float* bias, int act, int accumulate);
// weight + bias gradient of y = x W^T + b in one launch: gW[out,in] += dY^T X, gB[out] += sum_m dY[m,:]
int gemm_wgrad(cudaStream_t st, int out, int in, int M, const float* dY, int lddy, const float* X, int ldx, float* gW,
               int ldgw, float* gB);
// several gemm_wgrad problems in one launch (fused MLP chains: fused_mlp.cu)
constexpr int WGRAD_BATCH_MAX = 28;
struct WgradItem { const float* dY; const float* X; float* gW; float* gB; int out, in, rows, lddy, ldx, ldgw; };
int gemm_wgrad_batch(cudaStream_t st, const WgradItem* items, int n);
// dPre[m,n] = dY[m,n] * act'(Y[m,n]) in place on dY (Y is the activation OUTPUT)
int act_backward(cudaStream_t st, int M, int N, float* dY, int ldd, const float* Y, int ldy, int act);
// out[n] (+)= sum_m X[m,n]
int colsum(cudaStream_t st, int M, int N, const float* X, int ldx, float* out, int accumulate);
// y (+)= x  (flat)
int axpy(cudaStream_t st, long long n, const float* x, float* y);
int fill_zero(cudaStream_t st, void* p, size_t bytes);

// conv5x5.cu ---------------------------------------------------------------------------------
struct ConvSrc {          // one group of float4 channel-chunks of the (virtual) conv input
  const float4* ptr;      // [n_idx][nchunks][D][D] float4
  int nchunks;
  int div;                // n_idx = (div > 0) ? n / div : 0   (div=1 per slot-image, K per sample, 0 constant)
  int pitch;              // chunks per image in memory when only `nchunks` of them are read (0: = nchunks)
};
struct ConvInput {
  ConvSrc src[5];
  int nsrc;
  int total_chunks;
  int real_cin;           // channels that carry data (0 = 4*total_chunks); only used for FLOP accounting
  // fp16 mode only (see conv5x5_tc.cu): per-image largest magnitudes (fp32 bit patterns).  `maxbits` != NULL: already
  // known for this input (written by its producer), the consumer skips its own reduction pass.  `out_maxbits` != NULL:
  // the tensor-core kernel atomically maxes the magnitudes of the tensor it writes into it (the caller zeroes it).
  const uint32_t* maxbits;
  uint32_t* out_maxbits;
};
enum { EPI_NONE = 0, EPI_ELU = 1, EPI_MUL_ELUGRAD = 2,
       EPI_ACCUM = 8, EPI_MAX2 = 16 };   // flags of the quint-tap kernel's second K pass (conv5x5_tc.cu)
// out[n][cout/4][D][D] = epi( conv5x5_pad2(in, wt) + bias ); wt layout [25][cin_pad][cout] (cout fastest)
int conv5x5_forward(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float* wt, const float* bias,
                    float4* out, int epi, const float4* elu_ref);
// dW[25][cin_pad][cout] and db[cout] partial sums over all (n, pixels); scratch sized by conv5x5_wgrad_scratch_floats
size_t conv5x5_wgrad_scratch_floats(int cin_pad, int cout);
int conv5x5_wgrad(cudaStream_t st, int N, int D, const ConvInput& in, int cout, const float4* dpre, float* scratch,
                  float* dwt /*[25][cin_pad][cout], accumulated*/, float* dbias /*[cout], accumulated*/);

}  // namespace op3

// Pairwise K x K slot-interaction dynamics: replaces PhysicsNetwork.forward (op3/torch/op3_modules/physics_network.py:94-141),
// Mlp.forward (op3/torch/networks.py:63-74) and get_all_attention_values (physics_network.py:156-187), plus backward.
// Eleven one-hidden-layer MLPs chained over M = B*K (per slot) or B*K*(K-1) (per ordered pair) rows.
#include "layout.cuh"
#include "fused_mlp.cuh"

namespace op3 {

namespace {

struct MlpDims { int in, hid, out, out_act; };

int mlp_fwd(cudaStream_t st, int M, const float* x, int ldx, const Op3Mlp& w, MlpDims s, float* h, float* y, int ldy) {
  OP3_TRY(gemm(st, M, s.hid, s.in, x, ldx, 0, w.fc0.w, s.in, 1, h, s.hid, w.fc0.b, ACT_ELU, 0));
  OP3_TRY(gemm(st, M, s.out, s.hid, h, s.hid, 0, w.last.w, s.hid, 1, y, ldy, w.last.b, s.out_act, 0));
  return OP3_OK;
}

// dy is turned into dPre in place when the output activation is not the identity.
int mlp_bwd(cudaStream_t st, int M, const float* x, int ldx, const Op3Mlp& w, const Op3Mlp& gw, MlpDims s, const float* h,
            const float* y, int ldy, float* dy, int lddy, float* dh, float* dx, int lddx, int acc_dx) {
  OP3_TRY(act_backward(st, M, s.out, dy, lddy, y, ldy, s.out_act));
  OP3_TRY(gemm_wgrad(st, s.out, s.hid, M, dy, lddy, h, s.hid, gw.last.w, s.hid, gw.last.b));
  OP3_TRY(gemm(st, M, s.hid, s.out, dy, lddy, 0, w.last.w, s.hid, 0, dh, s.hid, nullptr, ACT_NONE, 0));
  OP3_TRY(act_backward(st, M, s.hid, dh, s.hid, h, s.hid, ACT_ELU));
  OP3_TRY(gemm_wgrad(st, s.hid, s.in, M, dh, s.hid, x, ldx, gw.fc0.w, s.in, gw.fc0.b));
  if (dx) OP3_TRY(gemm(st, M, s.in, s.hid, dh, s.hid, 0, w.fc0.w, s.in, 0, dx, lddx, nullptr, ACT_NONE, acc_dx));
  return OP3_OK;
}

// arep[n][0..A) = actions[n / K][0..A)
__global__ void rep_actions_kernel(int N, int K, int A, const float* __restrict__ actions, float* __restrict__ arep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * A) return;
  int n = i / A, a = i % A;
  arep[i] = actions[(size_t)(n / K) * A + a];
}

// enc[n][c] = eff[n][c] * att[n]
__global__ void gate_kernel(int N, int R, const float* __restrict__ eff, const float* __restrict__ att,
                            float* __restrict__ enc, int lde) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * R) return;
  int n = i / R, c = i % R;
  enc[(size_t)n * lde + c] = eff[i] * att[n];
}

// d_eff = d_enc * att ; d_att[n] = sum_c d_enc[n][c]*eff[n][c]   (one warp per row)
__global__ void __launch_bounds__(128) gate_bwd_kernel(int N, int R, const float* __restrict__ d_enc, int ldd,
                                                       const float* __restrict__ eff, const float* __restrict__ att,
                                                       float* __restrict__ d_eff, float* __restrict__ d_att) {
  int lane = threadIdx.x % 32, n = blockIdx.x * 4 + threadIdx.x / 32;
  if (n >= N) return;
  float a = att[n], s = 0.f;
  for (int c = lane; c < R; c += 32) {
    float de = d_enc[(size_t)n * ldd + c];
    d_eff[(size_t)n * R + c] = de * a;
    s += de * eff[(size_t)n * R + c];
  }
  s = warp_sum(s);
  if (lane == 0) d_att[n] = s;
}

// pairs[(b*K+i)*(K-1)+jj] = [enc[b,i] | enc[b,j]],  j = jj < i ? jj : jj+1   (physics_network.py:114-121)
__global__ void pairs_kernel(int B, int K, int R, const float* __restrict__ enc, int lde, float* __restrict__ pairs) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (long long)B * K * (K - 1) * 2 * R;
  if (i >= tot) return;
  int c = (int)(i % (2 * R));
  long long row = i / (2 * R);
  int jj = (int)(row % (K - 1));
  int n = (int)(row / (K - 1));
  int b = n / K, ii = n % K;
  int j = jj < ii ? jj : jj + 1;
  int src = c < R ? (b * K + ii) : (b * K + j);
  pairs[i] = enc[(size_t)src * lde + (c % R)];
}

// d_enc[b,i][c] += sum_jj d_pairs[(b,i),jj][c] + sum_{i' != i} d_pairs[(b,i'), pos(i)][R + c]
__global__ void pairs_bwd_kernel(int B, int K, int R, const float* __restrict__ d_pairs, float* __restrict__ d_enc, int ldd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K * R) return;
  int c = i % R, n = i / R;
  int b = n / K, ii = n % K;
  float s = 0.f;
  for (int jj = 0; jj < K - 1; ++jj) s += d_pairs[((size_t)n * (K - 1) + jj) * 2 * R + c];
  for (int ip = 0; ip < K; ++ip) {
    if (ip == ii) continue;
    int pos = ii < ip ? ii : ii - 1;
    s += d_pairs[((size_t)(b * K + ip) * (K - 1) + pos) * 2 * R + R + c];
  }
  d_enc[(size_t)n * ldd + c] += s;
}

// total[n][c] = sum_jj ef[(n,jj)][c] * pa[(n,jj)]
__global__ void effect_sum_kernel(int N, int K, int E, const float* __restrict__ ef, const float* __restrict__ pa,
                                  float* __restrict__ total, int ldt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * E) return;
  int n = i / E, c = i % E;
  float s = 0.f;
  for (int jj = 0; jj < K - 1; ++jj) {
    size_t r = (size_t)n * (K - 1) + jj;
    s += ef[r * E + c] * pa[r];
  }
  total[(size_t)n * ldt + c] = s;
}

// d_ef[(n,jj)][c] = d_total[n][c]*pa ; d_pa[(n,jj)] = sum_c d_total[n][c]*ef    (one warp per pair row, E = 32)
__global__ void __launch_bounds__(128) effect_sum_bwd_kernel(int NP, int K, int E, const float* __restrict__ d_total, int ldt,
                                                             const float* __restrict__ ef, const float* __restrict__ pa,
                                                             float* __restrict__ d_ef, float* __restrict__ d_pa) {
  int lane = threadIdx.x % 32, r = blockIdx.x * 4 + threadIdx.x / 32;
  if (r >= NP) return;
  int n = r / (K - 1);
  float a = pa[r], s = 0.f;
  for (int c = lane; c < E; c += 32) {
    float dt = d_total[(size_t)n * ldt + c];
    d_ef[(size_t)r * E + c] = dt * a;
    s += dt * ef[(size_t)r * E + c];
  }
  s = warp_sum(s);
  if (lane == 0) d_pa[r] = s;
}

__global__ void zero_cols_kernel(int M, int N, float* __restrict__ x, int ldx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  x[(size_t)(i / N) * ldx + (i % N)] = 0.f;
}

// deltas[n] = sum_c |enc[n][c] - enc_flat[n][c]|   (physics_network.py:172)
__global__ void __launch_bounds__(128) deltas_kernel(int N, int R, const float* __restrict__ enc, int lde,
                                                     const float* __restrict__ encf, int ldf, float* __restrict__ out) {
  int lane = threadIdx.x % 32, n = blockIdx.x * 4 + threadIdx.x / 32;
  if (n >= N) return;
  float s = 0.f;
  for (int c = lane; c < R; c += 32) s += fabsf(enc[(size_t)n * lde + c] - encf[(size_t)n * ldf + c]);
  s = warp_sum(s);
  if (lane == 0) out[n] = s;
}

struct DynActs {
  float *ie_h, *SA, *arep, *ae_h, *eff_h, *eff, *att_h, *att, *SE;
  float *pairs, *pw_h, *inter, *ef_h, *ef, *pa_h, *pa;
  float *mg_h, *agg, *det_h, *l1_h, *l2_h;
  size_t total;
};
inline DynActs dyn_acts(const Op3Dims* d, float* base) {
  DynActs A;
  const size_t N = (size_t)d->B * d->K, NP = N * (d->K - 1), R = d->Rd + d->Rs, E = OP3_AENC;
  size_t o = 0;
  auto take = [&](size_t n) { float* r = base ? base + o : nullptr; o += (n + 3) & ~(size_t)3; return r; };
  A.ie_h = take(N * R); A.SA = take(N * (R + E)); A.arep = take(N * 4); A.ae_h = take(N * R);
  A.eff_h = take(N * R); A.eff = take(N * R); A.att_h = take(N * R); A.att = take(N); A.SE = take(N * (R + E));
  A.pairs = take(NP * 2 * R); A.pw_h = take(NP * 2 * R); A.inter = take(NP * R); A.ef_h = take(NP * R);
  A.ef = take(NP * E); A.pa_h = take(NP * R); A.pa = take(NP);
  A.mg_h = take(N * R); A.agg = take(N * R); A.det_h = take(N * R); A.l1_h = take(N * R); A.l2_h = take(N * R);
  A.total = o;
  return A;
}


// ===================================================================================================================
// Fused path: ONE CTA owns S whole samples (S*K <= 16 slot rows and their S*K*(K-1) ordered pair rows, walked in chunks of
// 16) and runs the entire chain of physics_network.py:94-141 with the activations in shared memory (fused_mlp.cuh).
// Every intermediate the backward pass / the attention probes need is also written to `acts` in the DynActs layout.
// ===================================================================================================================
using namespace fused;

// panel pitches (floats): width + 4, see PITCH() in fused_mlp.cuh; sized for R = 128
constexpr int LR = PITCH(128), L2R = PITCH(256), LSE = PITCH(160), L32 = PITCH(32), L64 = PITCH(64);
constexpr int DYN_FWD_SMEM_FLOATS = GEMM_SMEM_FLOATS + 16 * (LR + L2R + LSE + LSE + LR + L32 + L2R + LR + L32 + 2 + LR);
constexpr int DYN_BWD_SMEM_FLOATS = GEMM_SMEM_FLOATS + 16 * (L64 + L2R + LR + LR + LSE + L32 + L32 + LR + L2R + LSE + LR);

__global__ void __launch_bounds__(FT, 1) dyn_fwd_fused_kernel(const Op3Dims d, const __grid_constant__ Op3Params p, const DynActs A,
                                                           const float* __restrict__ z, const float* __restrict__ actions,
                                                           float* __restrict__ out_z, float* __restrict__ l1_out,
                                                           float* __restrict__ l2_out, int S, int tc) {
  extern __shared__ __align__(16) float sm[];
  float* wt = sm;
  float* PZ = wt + GEMM_SMEM_FLOATS;     // [16][LR]   z rows
  float* PH = PZ + 16 * LR;              // [16][L2R]  hidden layer of the MLP in flight
  float* PSA = PH + 16 * L2R;            // [16][LSE]  [inertia enc | action enc]
  float* PSE = PSA + 16 * LSE;           // [16][LSE]  [gated enc | total pair effect]
  float* PEFF = PSE + 16 * LSE;          // [16][LR]
  float* PA = PEFF + 16 * LR;            // [16][L32]  actions (zero padded), then the action attention in column 0
  float* PP = PA + 16 * L32;             // [16][L2R]  pair inputs of the chunk
  float* PI = PP + 16 * L2R;             // [16][LR]   pair encodings
  float* PEF = PI + 16 * LR;             // [16][L32]  pair effects
  float* PPA = PEF + 16 * L32;           // [32]       pair attentions
  float* PAGG = PPA + 32;                // [16][LR]
  const int tid = threadIdx.x;
#ifdef OP3_FUSED_TRACE
  const long long tk0 = clock64();
#endif
  const int K = d.K, R = d.Rd + d.Rs, E = OP3_AENC, Rs = d.Rs, Rd = d.Rd, ldS = R + E, Ad = d.A;
  const int s0 = blockIdx.x * S, ns = min(S, d.B - s0);
  const int row0 = s0 * K, rows = ns * K;
  const size_t r0 = (size_t)row0;
  // one MLP: hidden layer (ELU) into PH + acts, output layer into a panel + acts
  auto mlp = [&](const float* X, int ldx, int nrows, const Op3Mlp& m, int in, int hid, int out, int out_act, float* gh, float* so,
                 int ldso, float* go, int ldgo) {
    cta_gemm<false>(tc, X, ldx, nrows, m.fc0.w, in, in, hid, wt, epi_fwd(m.fc0.b, ACT_ELU, PH, L2R, gh, hid));
    cta_gemm<false>(tc, PH, L2R, nrows, m.last.w, hid, hid, out, wt, epi_fwd(m.last.b, out_act, so, ldso, go, ldgo));
  };
  zero_panel(PZ, DYN_FWD_SMEM_FLOATS - GEMM_SMEM_FLOATS);      // rows the GEMMs compute but nobody owns must stay finite
  __syncthreads();
  load_panel(PZ, LR, z + r0 * R, R, rows, R);
  __syncthreads();
  if (Ad > 0) {
    mlp(PZ, LR, rows, p.dyn_inertia, R, R, R, ACT_ELU, A.ie_h + r0 * R, PSA, LSE, A.SA + r0 * ldS, ldS);
    for (int e = tid; e < rows * Ad; e += FT) {               // physics_network.py:104-106: every slot sees its sample's action
      const int r = e / Ad, a = e - r * Ad;
      const float v = actions[(size_t)(s0 + r / K) * Ad + a];
      PA[r * L32 + a] = v;
      A.arep[(r0 + r) * Ad + a] = v;
    }
    __syncthreads();
    mlp(PA, L32, rows, p.dyn_action_enc, Ad, R, E, ACT_ELU, A.ae_h + r0 * R, PSA + R, LSE, A.SA + r0 * ldS + R, ldS);
    mlp(PSA, LSE, rows, p.dyn_action_effect, ldS, R, R, ACT_ELU, A.eff_h + r0 * R, PEFF, LR, A.eff + r0 * R, R);
    mlp(PSA, LSE, rows, p.dyn_action_att, ldS, R, 1, ACT_SIGMOID, A.att_h + r0 * R, PA, L32, A.att + r0, 1);
    for (int e = tid; e < rows * R; e += FT) {                // enc = effect * attention (physics_network.py:111)
      const int r = e / R, c = e - r * R;
      const float v = PEFF[r * LR + c] * PA[r * L32];
      PSE[r * LSE + c] = v;
      A.SE[(r0 + r) * ldS + c] = v;
    }
  } else {
    mlp(PZ, LR, rows, p.dyn_inertia, R, R, R, ACT_ELU, A.ie_h + r0 * R, PSE, LSE, A.SE + r0 * ldS, ldS);
  }
  for (int e = tid; e < rows * E; e += FT) PSE[(e / E) * LSE + R + (e % E)] = 0.f;
  __syncthreads();
  if (K > 1) {
    const int npr = rows * (K - 1);
    const size_t pr0 = r0 * (K - 1);
    const int q4 = (2 * R) >> 2;
    for (int c0 = 0; c0 < npr; c0 += RTM) {
      const int pc = min(RTM, npr - c0);
      for (int e = tid; e < pc * q4; e += FT) {               // [enc_i | enc_j], j != i in increasing j (physics_network.py:114-121)
        const int lp = e / q4, c = (e - lp * q4) * 4;
        const int gp = c0 + lp, nl = gp / (K - 1), jj = gp - nl * (K - 1);
        const int ii = nl % K, j = jj < ii ? jj : jj + 1;
        const int src = c < R ? nl : nl - ii + j;
        const float4 v = *reinterpret_cast<const float4*>(PSE + src * LSE + (c < R ? c : c - R));
        *reinterpret_cast<float4*>(PP + lp * L2R + c) = v;
        *reinterpret_cast<float4*>(A.pairs + (pr0 + gp) * 2 * R + c) = v;
      }
      __syncthreads();
      const size_t g0 = pr0 + c0;
      mlp(PP, L2R, pc, p.dyn_pairwise, 2 * R, 2 * R, R, ACT_ELU, A.pw_h + g0 * 2 * R, PI, LR, A.inter + g0 * R, R);
      mlp(PI, LR, pc, p.dyn_inter_effect, R, R, E, ACT_ELU, A.ef_h + g0 * R, PEF, L32, A.ef + g0 * E, E);
      mlp(PI, LR, pc, p.dyn_inter_att, R, R, 1, ACT_SIGMOID, A.pa_h + g0 * R, PPA, 1, A.pa + g0, 1);
      for (int e = tid; e < rows * E; e += FT) {              // total effect += effect * attention, pairs in increasing j
        const int nl = e / E, c = e - nl * E;
        const int lo = max(nl * (K - 1), c0) - c0, hi = min((nl + 1) * (K - 1), c0 + pc) - c0;
        float s = PSE[nl * LSE + R + c];
        for (int lp = lo; lp < hi; ++lp) s += PEF[lp * L32 + c] * PPA[lp];
        PSE[nl * LSE + R + c] = s;
      }
      __syncthreads();
    }
  }
  for (int e = tid; e < rows * E; e += FT) A.SE[(r0 + e / E) * ldS + R + (e % E)] = PSE[(e / E) * LSE + R + (e % E)];
  mlp(PSE, LSE, rows, p.dyn_merge, ldS, R, R, ACT_ELU, A.mg_h + r0 * R, PAGG, LR, A.agg + r0 * R, R);
  if (Rd > 0) mlp(PAGG, LR, rows, p.dyn_det_out, R, R, Rd, ACT_NONE, A.det_h + r0 * R, nullptr, 0, out_z + r0 * R, R);
  mlp(PAGG, LR, rows, p.dyn_lam1_out, R, R, Rs, ACT_NONE, A.l1_h + r0 * R, nullptr, 0, l1_out + r0 * Rs, Rs);
  mlp(PAGG, LR, rows, p.dyn_lam2_out, R, R, Rs, ACT_NONE, A.l2_h + r0 * R, nullptr, 0, l2_out + r0 * Rs, Rs);
#ifdef OP3_FUSED_TRACE
  FTRACE(0, clock64() - tk0);
#endif
}

// Backward of the chain above for the same samples: every pre-activation gradient is written to the gradient twin G of its
// activation buffer (the batched weight-gradient launch that follows contracts them over ALL rows), dz is written.
__global__ void __launch_bounds__(FT, 1) dyn_bwd_fused_kernel(const Op3Dims d, const __grid_constant__ Op3Params p, const DynActs A,
                                                           const DynActs G, const float* __restrict__ d_out_z,
                                                           const float* __restrict__ d_l1, const float* __restrict__ d_l2,
                                                           float* __restrict__ dz, int S, int tc) {
  extern __shared__ __align__(16) float sm[];
  float* wt = sm;
  float* PU = wt + GEMM_SMEM_FLOATS;     // [16][L64]  upstream gradient of the head in flight
  float* PH = PU + 16 * L64;             // [16][L2R]  hidden-layer pre-activation gradient
  float* PDAGG = PH + 16 * L2R;          // [16][LR]
  float* PX = PDAGG + 16 * LR;           // [16][LR]   pre-activation gradient feeding the next layer down
  float* PDSE = PX + 16 * LR;            // [16][LSE]
  float* PDEF = PDSE + 16 * LSE;         // [16][L32]
  float* PD1 = PDEF + 16 * L32;          // [16][L32]  scalar attention gradients in column 0, zero elsewhere
  float* PDI = PD1 + 16 * L32;           // [16][LR]
  float* PDP = PDI + 16 * LR;            // [16][L2R]
  float* PDSA = PDP + 16 * L2R;          // [16][LSE]
  float* PDEFF = PDSA + 16 * LSE;        // [16][LR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = d.K, R = d.Rd + d.Rs, E = OP3_AENC, Rs = d.Rs, Rd = d.Rd, ldS = R + E, Ad = d.A;
  const int s0 = blockIdx.x * S, ns = min(S, d.B - s0);
  const int row0 = s0 * K, rows = ns * K;
  const size_t r0 = (size_t)row0;
  // one MLP backwards from the pre-activation gradient dY of its output layer (panel, `out` columns):
  // dPre_hidden = (dY W_last) * ELU'(h) -> PH + twin gh; then dX = dPre_hidden W_fc0 -> panel so (set / accumulate) and/or global go
  auto mlp = [&](const float* dY, int ldy, int nrows, const Op3Mlp& m, int in, int hid, int out, const float* h, float* gh,
                 float* so, int ldso, int so_acc, float* go, int ldgo) {
    cta_gemm<true>(tc, dY, ldy, nrows, m.last.w, hid, out, hid, wt, epi_bwd(h, hid, PH, L2R, 0, gh, hid, 0));
    if (so != nullptr || go != nullptr)
      cta_gemm<true>(tc, PH, L2R, nrows, m.fc0.w, in, hid, in, wt, epi_bwd(nullptr, 0, so, ldso, so_acc, go, ldgo, 0));
  };
  zero_panel(PU, DYN_BWD_SMEM_FLOATS - GEMM_SMEM_FLOATS);      // rows the GEMMs compute but nobody owns must stay finite
  __syncthreads();
  bool first = true;
  for (int hd = (Rd > 0 ? 0 : 1); hd < 3; ++hd) {
    const Op3Mlp& m = hd == 0 ? p.dyn_det_out : (hd == 1 ? p.dyn_lam1_out : p.dyn_lam2_out);
    const float* up = hd == 0 ? d_out_z : (hd == 1 ? d_l1 : d_l2);
    const int ldu = hd == 0 ? R : Rs, wu = hd == 0 ? Rd : Rs;
    const float* hs = hd == 0 ? A.det_h : (hd == 1 ? A.l1_h : A.l2_h);
    float* gh = hd == 0 ? G.det_h : (hd == 1 ? G.l1_h : G.l2_h);
    load_panel(PU, L64, up + r0 * ldu, ldu, rows, wu);
    __syncthreads();
    mlp(PU, L64, rows, m, R, R, wu, hs + r0 * R, gh + r0 * R, PDAGG, LR, first ? 0 : 1, nullptr, 0);
    first = false;
  }
  for (int e = tid; e < rows * R; e += FT) {
    const int r = e / R, c = e - r * R;
    const size_t gi = (r0 + r) * R + c;
    const float v = PDAGG[r * LR + c] * elu_grad_from_out(A.agg[gi]);
    PX[r * LR + c] = v;
    G.agg[gi] = v;
  }
  __syncthreads();
  mlp(PX, LR, rows, p.dyn_merge, ldS, R, R, A.mg_h + r0 * R, G.mg_h + r0 * R, PDSE, LSE, 0, nullptr, 0);
  if (K > 1) {
    const int npr = rows * (K - 1);
    const size_t pr0 = r0 * (K - 1);
    for (int c0 = 0; c0 < npr; c0 += RTM) {
      const int pc = min(RTM, npr - c0);
      const size_t g0 = pr0 + c0;
      for (int lp = warp; lp < pc; lp += FT / 32) {           // total = sum_j effect * attention: one warp per pair row (E = 32)
        const int nl = (c0 + lp) / (K - 1);
        const float dt = PDSE[nl * LSE + R + lane];
        const float pa = A.pa[g0 + lp], ef = A.ef[(g0 + lp) * E + lane];
        const float de = dt * pa * elu_grad_from_out(ef);
        PDEF[lp * L32 + lane] = de;
        G.ef[(g0 + lp) * E + lane] = de;
        const float s = warp_sum(dt * ef);
        if (lane == 0) {
          const float dp = s * pa * (1.f - pa);
          PD1[lp * L32] = dp;
          G.pa[g0 + lp] = dp;
        }
      }
      __syncthreads();
      mlp(PDEF, L32, pc, p.dyn_inter_effect, R, R, E, A.ef_h + g0 * R, G.ef_h + g0 * R, PDI, LR, 0, nullptr, 0);
      mlp(PD1, L32, pc, p.dyn_inter_att, R, R, 1, A.pa_h + g0 * R, G.pa_h + g0 * R, PDI, LR, 1, nullptr, 0);
      for (int e = tid; e < pc * R; e += FT) {
        const int r = e / R, c = e - r * R;
        const size_t gi = (g0 + r) * R + c;
        const float v = PDI[r * LR + c] * elu_grad_from_out(A.inter[gi]);
        PX[r * LR + c] = v;
        G.inter[gi] = v;
      }
      __syncthreads();
      mlp(PX, LR, pc, p.dyn_pairwise, 2 * R, 2 * R, R, A.pw_h + g0 * 2 * R, G.pw_h + g0 * 2 * R, PDP, L2R, 0, nullptr, 0);
      for (int e = tid; e < rows * R; e += FT) {              // gather of the pair-input gradients, pair rows in increasing order
        const int nl = e / R, c = e - nl * R;
        const int smp = nl / K, ii = nl - smp * K;
        float s = 0.f;
        for (int lp = 0; lp < pc; ++lp) {
          const int gp = c0 + lp, ni = gp / (K - 1);
          if (ni / K != smp) continue;
          const int jj = gp - ni * (K - 1), i2 = ni - smp * K, j = jj < i2 ? jj : jj + 1;
          if (ni == nl) s += PDP[lp * L2R + c];
          if (j == ii) s += PDP[lp * L2R + R + c];
        }
        PDSE[nl * LSE + c] += s;
      }
      __syncthreads();
    }
  }
  const float* dy_in;                                        // pre-activation gradient of the inertia encoder's output layer
  int ld_in;
  if (Ad > 0) {
    for (int r = warp; r < rows; r += FT / 32) {              // enc = effect * attention
      const float att = A.att[row0 + r];
      float s = 0.f;
      for (int c = lane; c < R; c += 32) {
        const size_t gi = (r0 + r) * R + c;
        const float de = PDSE[r * LSE + c], eff = A.eff[gi];
        const float v = de * att * elu_grad_from_out(eff);
        PDEFF[r * LR + c] = v;
        G.eff[gi] = v;
        s += de * eff;
      }
      s = warp_sum(s);
      if (lane == 0) {
        const float da = s * att * (1.f - att);
        PD1[r * L32] = da;
        G.att[row0 + r] = da;
      }
    }
    __syncthreads();
    mlp(PDEFF, LR, rows, p.dyn_action_effect, ldS, R, R, A.eff_h + r0 * R, G.eff_h + r0 * R, PDSA, LSE, 0, nullptr, 0);
    mlp(PD1, L32, rows, p.dyn_action_att, ldS, R, 1, A.att_h + r0 * R, G.att_h + r0 * R, PDSA, LSE, 1, nullptr, 0);
    for (int e = tid; e < rows * ldS; e += FT) {
      const int r = e / ldS, c = e - r * ldS;
      const size_t gi = (r0 + r) * ldS + c;
      const float v = PDSA[r * LSE + c] * elu_grad_from_out(A.SA[gi]);
      PDSA[r * LSE + c] = v;
      G.SA[gi] = v;
    }
    __syncthreads();
    mlp(PDSA + R, LSE, rows, p.dyn_action_enc, Ad, R, E, A.ae_h + r0 * R, G.ae_h + r0 * R, nullptr, 0, 0, nullptr, 0);
    dy_in = PDSA; ld_in = LSE;
  } else {
    for (int e = tid; e < rows * R; e += FT) {
      const int r = e / R, c = e - r * R;
      const size_t gi = (r0 + r) * ldS + c;
      const float v = PDSE[r * LSE + c] * elu_grad_from_out(A.SE[gi]);
      PX[r * LR + c] = v;
      G.SE[gi] = v;
    }
    __syncthreads();
    dy_in = PX; ld_in = LR;
  }
  mlp(dy_in, ld_in, rows, p.dyn_inertia, R, R, R, A.ie_h + r0 * R, G.ie_h + r0 * R, nullptr, 0, 0, dz + r0 * R, R);
}

inline bool mlp_aligned(const Op3Mlp& m) { return aligned16(m.fc0.w) && aligned16(m.last.w); }
// the fused kernels need 16-byte aligned weight rows (cp.async) -- true for the engine's flat parameter buffer
inline bool dyn_fusable(const Op3Dims* d, const Op3Params* p) {
  bool ok = mlp_aligned(p->dyn_inertia) && mlp_aligned(p->dyn_pairwise) && mlp_aligned(p->dyn_inter_effect) &&
            mlp_aligned(p->dyn_inter_att) && mlp_aligned(p->dyn_merge) && mlp_aligned(p->dyn_lam1_out) && mlp_aligned(p->dyn_lam2_out);
  if (d->Rd > 0) ok = ok && mlp_aligned(p->dyn_det_out);
  if (d->A > 0) ok = ok && mlp_aligned(p->dyn_action_enc) && mlp_aligned(p->dyn_action_effect) && mlp_aligned(p->dyn_action_att);
  return ok;
}
inline int dyn_samples_per_cta(const Op3Dims* d) {
  const int smax = RTM / d->K > 0 ? RTM / d->K : 1;
  int S = ceil_div(d->B, device_sm_count());
  return S < 1 ? 1 : (S > smax ? smax : S);
}
inline double dyn_flops(const Op3Dims* d) {
  const double R = d->Rd + d->Rs, E = OP3_AENC, N = (double)d->B * d->K, NP = N * (d->K - 1);
  double per_slot = 2 * R * R + (d->A > 0 ? d->A * R + R * E + 2 * ((R + E) * R) + R * R + R : 0) + (R + E) * R + R * R +
                    (d->Rd > 0 ? R * R + R * d->Rd : 0) + 2 * (R * R + R * d->Rs);
  double per_pair = 4 * R * R + 2 * R * R + 2 * R * R + R * E + R;
  return 2.0 * (N * per_slot + NP * per_pair);
}

}  // namespace
}  // namespace op3

using namespace op3;

#ifdef OP3_FUSED_TRACE
extern "C" int op3_debug_fused_trace(unsigned long long* out /*[8] host*/, int reset) {
  if (out) cudaMemcpyFromSymbol(out, fused::g_fused_trace, sizeof(unsigned long long) * 8);
  if (reset) { static unsigned long long z[8]; cudaMemcpyToSymbol(fused::g_fused_trace, z, sizeof(z)); }
  return 0;
}
#endif

extern "C" size_t op3_dynamics_acts_floats(const Op3Dims* d) {
  if (validate_dims(d) != OP3_OK) return 0;
  return dyn_acts(d, nullptr).total;
}

// backward scratch mirrors the activation layout (one gradient buffer per activation) plus hidden-gradient scratch
extern "C" size_t op3_dynamics_scratch_floats(const Op3Dims* d) {
  if (validate_dims(d) != OP3_OK) return 0;
  const size_t N = (size_t)d->B * d->K, NP = N * (d->K > 1 ? d->K - 1 : 1), R = d->Rd + d->Rs;
  return dyn_acts(d, nullptr).total + NP * 2 * R + N * R + 64;
}

extern "C" int op3_dynamics_fwd(const Op3Dims* d, const Op3Params* p, const float* z, const float* actions, float* acts,
                                float* out_z, float* l1_out, float* l2_out, void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(p && z && acts && l1_out && l2_out, "op3_dynamics_fwd: NULL argument");
  OP3_REQUIRE(d->Rd == 0 || out_z, "op3_dynamics_fwd: out_z required when Rd > 0");
  OP3_REQUIRE((d->A > 0) == (actions != nullptr), "op3_dynamics_fwd: actions must be given iff A > 0 (A=%d)", d->A);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, K = d->K, N = B * K, NP = N * (K - 1), R = d->Rd + d->Rs, E = OP3_AENC, Rs = d->Rs;
  DynActs A = dyn_acts(d, acts);
  const int ldS = R + E;
  if (fused_mlp_enabled() && dyn_fusable(d, p) && aligned16(z)) {
    static PerDeviceOnce attr_once;
    if (attr_once.first())
      OP3_CHECK(cudaFuncSetAttribute(dyn_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     DYN_FWD_SMEM_FLOATS * (int)sizeof(float)));
    const int S = dyn_samples_per_cta(d);
    ProfScope prof(st, KID_GEMM, dyn_flops(d));
    dyn_fwd_fused_kernel<<<ceil_div(B, S), FT, DYN_FWD_SMEM_FLOATS * sizeof(float), st>>>(
        *d, *p, A, z, actions, out_z, l1_out, l2_out, S, op3_get_precision() != OP3_PREC_FP32);
    OP3_COUNT_LAUNCH();
    OP3_LAUNCH_CHECK();
    return OP3_OK;
  }
  if (d->A > 0) {
    OP3_TRY(mlp_fwd(st, N, z, R, p->dyn_inertia, {R, R, R, ACT_ELU}, A.ie_h, A.SA, ldS));
    rep_actions_kernel<<<ceil_div(N * d->A, 256), 256, 0, st>>>(N, K, d->A, actions, A.arep);
    OP3_COUNT_LAUNCH();
    OP3_TRY(mlp_fwd(st, N, A.arep, d->A, p->dyn_action_enc, {d->A, R, E, ACT_ELU}, A.ae_h, A.SA + R, ldS));
    OP3_TRY(mlp_fwd(st, N, A.SA, ldS, p->dyn_action_effect, {R + E, R, R, ACT_ELU}, A.eff_h, A.eff, R));
    OP3_TRY(mlp_fwd(st, N, A.SA, ldS, p->dyn_action_att, {R + E, R, 1, ACT_SIGMOID}, A.att_h, A.att, 1));
    gate_kernel<<<ceil_div(N * R, 256), 256, 0, st>>>(N, R, A.eff, A.att, A.SE, ldS);
    OP3_COUNT_LAUNCH();
  } else {
    OP3_TRY(mlp_fwd(st, N, z, R, p->dyn_inertia, {R, R, R, ACT_ELU}, A.ie_h, A.SE, ldS));
  }
  if (K > 1) {
    long long tot = (long long)NP * 2 * R;
    pairs_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(B, K, R, A.SE, ldS, A.pairs);
    OP3_COUNT_LAUNCH();
    OP3_TRY(mlp_fwd(st, NP, A.pairs, 2 * R, p->dyn_pairwise, {2 * R, 2 * R, R, ACT_ELU}, A.pw_h, A.inter, R));
    OP3_TRY(mlp_fwd(st, NP, A.inter, R, p->dyn_inter_effect, {R, R, E, ACT_ELU}, A.ef_h, A.ef, E));
    OP3_TRY(mlp_fwd(st, NP, A.inter, R, p->dyn_inter_att, {R, R, 1, ACT_SIGMOID}, A.pa_h, A.pa, 1));
    effect_sum_kernel<<<ceil_div(N * E, 256), 256, 0, st>>>(N, K, E, A.ef, A.pa, A.SE + R, ldS);
    OP3_COUNT_LAUNCH();
  } else {
    zero_cols_kernel<<<ceil_div(N * E, 256), 256, 0, st>>>(N, E, A.SE + R, ldS);
    OP3_COUNT_LAUNCH();
  }
  OP3_LAUNCH_CHECK();
  OP3_TRY(mlp_fwd(st, N, A.SE, ldS, p->dyn_merge, {R + E, R, R, ACT_ELU}, A.mg_h, A.agg, R));
  if (d->Rd > 0) OP3_TRY(mlp_fwd(st, N, A.agg, R, p->dyn_det_out, {R, R, d->Rd, ACT_NONE}, A.det_h, out_z, R));
  OP3_TRY(mlp_fwd(st, N, A.agg, R, p->dyn_lam1_out, {R, R, Rs, ACT_NONE}, A.l1_h, l1_out, Rs));
  OP3_TRY(mlp_fwd(st, N, A.agg, R, p->dyn_lam2_out, {R, R, Rs, ACT_NONE}, A.l2_h, l2_out, Rs));
  return OP3_OK;
}

extern "C" int op3_dynamics_probes(const Op3Dims* d, const float* acts, float* act_att, float* pair_att, float* deltas,
                                   void* stream) {
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(acts, "op3_dynamics_probes: NULL acts");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = d->B * d->K, NP = N * (d->K - 1), R = d->Rd + d->Rs, E = OP3_AENC;
  DynActs A = dyn_acts(d, const_cast<float*>(acts));
  if (act_att && d->A > 0) OP3_CHECK(cudaMemcpyAsync(act_att, A.att, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (pair_att && d->K > 1) OP3_CHECK(cudaMemcpyAsync(pair_att, A.pa, (size_t)NP * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (deltas) {
    if (d->A > 0) {
      deltas_kernel<<<ceil_div(N, 4), 128, 0, st>>>(N, R, A.SE, R + E, A.SA, R + E, deltas);
      OP3_COUNT_LAUNCH();
      OP3_LAUNCH_CHECK();
    } else {
      OP3_CHECK(cudaMemsetAsync(deltas, 0, (size_t)N * sizeof(float), st));
    }
  }
  return OP3_OK;
}

extern "C" int op3_dynamics_bwd(const Op3Dims* d, const Op3Params* p, const float* z, const float* actions,
                                const float* acts, const float* d_out_z, const float* d_l1, const float* d_l2, float* dz,
                                const Op3Params* g,
                                float* scratch, void* stream) {
  (void)actions;
  OP3_TRY(validate_dims(d));
  OP3_REQUIRE(p && z && acts && d_l1 && d_l2 && dz && g && scratch, "op3_dynamics_bwd: NULL argument");
  OP3_REQUIRE(d->Rd == 0 || d_out_z, "op3_dynamics_bwd: d_out_z required when Rd > 0");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, K = d->K, N = B * K, NP = N * (K - 1), R = d->Rd + d->Rs, E = OP3_AENC, Rs = d->Rs;
  const int ldS = R + E;
  DynActs A = dyn_acts(d, const_cast<float*>(acts));
  DynActs G = dyn_acts(d, scratch);  // gradient twin of every activation buffer
  float* dh = scratch + G.total;     // hidden-gradient scratch, up to [NP][2R]
  if (fused_mlp_enabled() && dyn_fusable(d, p) && aligned16(d_l1) && aligned16(d_l2) && (d->Rd == 0 || aligned16(d_out_z))) {
    static PerDeviceOnce attr_once;
    if (attr_once.first())
      OP3_CHECK(cudaFuncSetAttribute(dyn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     DYN_BWD_SMEM_FLOATS * (int)sizeof(float)));
    const int S = dyn_samples_per_cta(d);
    {
      ProfScope prof(st, KID_GEMM, 2.0 * dyn_flops(d));
      dyn_bwd_fused_kernel<<<ceil_div(B, S), FT, DYN_BWD_SMEM_FLOATS * sizeof(float), st>>>(
          *d, *p, A, G, d_out_z, d_l1, d_l2, dz, S, op3_get_precision() != OP3_PREC_FP32);
      OP3_COUNT_LAUNCH();
      OP3_LAUNCH_CHECK();
    }
    // every weight and bias gradient of the eleven MLPs in one launch: gW += dPre^T X over all rows
    WgradItem it[22];
    int n = 0;
    auto add = [&](const Op3Mlp& gw, const float* dy1, int ldy1, const float* h, int hid, int out, const float* dy0, const float* x,
                   int ldx, int in, int rows) {
      it[n++] = {dy1, h, gw.last.w, gw.last.b, out, hid, rows, ldy1, hid, hid};
      it[n++] = {dy0, x, gw.fc0.w, gw.fc0.b, hid, in, rows, hid, ldx, in};
    };
    if (d->Rd > 0) add(g->dyn_det_out, d_out_z, R, A.det_h, R, d->Rd, G.det_h, A.agg, R, R, N);
    add(g->dyn_lam1_out, d_l1, Rs, A.l1_h, R, Rs, G.l1_h, A.agg, R, R, N);
    add(g->dyn_lam2_out, d_l2, Rs, A.l2_h, R, Rs, G.l2_h, A.agg, R, R, N);
    add(g->dyn_merge, G.agg, R, A.mg_h, R, R, G.mg_h, A.SE, ldS, ldS, N);
    if (K > 1) {
      add(g->dyn_inter_effect, G.ef, E, A.ef_h, R, E, G.ef_h, A.inter, R, R, NP);
      add(g->dyn_inter_att, G.pa, 1, A.pa_h, R, 1, G.pa_h, A.inter, R, R, NP);
      add(g->dyn_pairwise, G.inter, R, A.pw_h, 2 * R, R, G.pw_h, A.pairs, 2 * R, 2 * R, NP);
    }
    if (d->A > 0) {
      add(g->dyn_action_effect, G.eff, R, A.eff_h, R, R, G.eff_h, A.SA, ldS, ldS, N);
      add(g->dyn_action_att, G.att, 1, A.att_h, R, 1, G.att_h, A.SA, ldS, ldS, N);
      add(g->dyn_action_enc, G.SA + R, ldS, A.ae_h, R, E, G.ae_h, A.arep, d->A, d->A, N);
      add(g->dyn_inertia, G.SA, ldS, A.ie_h, R, R, G.ie_h, z, R, R, N);
    } else {
      add(g->dyn_inertia, G.SE, ldS, A.ie_h, R, R, G.ie_h, z, R, R, N);
    }
    return gemm_wgrad_batch(st, it, n);
  }
  // outputs (identity activations: upstream gradients are used in place, they are not modified)
  bool first = true;
  if (d->Rd > 0) {
    OP3_TRY(mlp_bwd(st, N, A.agg, R, p->dyn_det_out, g->dyn_det_out, {R, R, d->Rd, ACT_NONE}, A.det_h, nullptr, 0,
                    const_cast<float*>(d_out_z), R, dh, G.agg, R, 0));
    first = false;
  }
  OP3_TRY(mlp_bwd(st, N, A.agg, R, p->dyn_lam1_out, g->dyn_lam1_out, {R, R, Rs, ACT_NONE}, A.l1_h, nullptr, 0,
                  const_cast<float*>(d_l1), Rs, dh, G.agg, R, first ? 0 : 1));
  OP3_TRY(mlp_bwd(st, N, A.agg, R, p->dyn_lam2_out, g->dyn_lam2_out, {R, R, Rs, ACT_NONE}, A.l2_h, nullptr, 0,
                  const_cast<float*>(d_l2), Rs, dh, G.agg, R, 1));
  // merge: x = SE = [enc | total_effect]
  OP3_TRY(mlp_bwd(st, N, A.SE, ldS, p->dyn_merge, g->dyn_merge, {R + E, R, R, ACT_ELU}, A.mg_h, A.agg, R, G.agg, R, dh,
                  G.SE, ldS, 0));
  if (K > 1) {
    effect_sum_bwd_kernel<<<ceil_div(NP, 4), 128, 0, st>>>(NP, K, E, G.SE + R, ldS, A.ef, A.pa, G.ef, G.pa);
    OP3_COUNT_LAUNCH();
    OP3_LAUNCH_CHECK();
    OP3_TRY(mlp_bwd(st, NP, A.inter, R, p->dyn_inter_effect, g->dyn_inter_effect, {R, R, E, ACT_ELU}, A.ef_h, A.ef, E, G.ef, E,
                    dh, G.inter, R, 0));
    OP3_TRY(mlp_bwd(st, NP, A.inter, R, p->dyn_inter_att, g->dyn_inter_att, {R, R, 1, ACT_SIGMOID}, A.pa_h, A.pa, 1, G.pa, 1,
                    dh, G.inter, R, 1));
    OP3_TRY(mlp_bwd(st, NP, A.pairs, 2 * R, p->dyn_pairwise, g->dyn_pairwise, {2 * R, 2 * R, R, ACT_ELU}, A.pw_h, A.inter, R,
                    G.inter, R, dh, G.pairs, 2 * R, 0));
    pairs_bwd_kernel<<<ceil_div(N * R, 256), 256, 0, st>>>(B, K, R, G.pairs, G.SE, ldS);
    OP3_COUNT_LAUNCH();
    OP3_LAUNCH_CHECK();
  }
  if (d->A > 0) {
    gate_bwd_kernel<<<ceil_div(N, 4), 128, 0, st>>>(N, R, G.SE, ldS, A.eff, A.att, G.eff, G.att);
    OP3_COUNT_LAUNCH();
    OP3_LAUNCH_CHECK();
    OP3_TRY(mlp_bwd(st, N, A.SA, ldS, p->dyn_action_effect, g->dyn_action_effect, {R + E, R, R, ACT_ELU}, A.eff_h, A.eff, R,
                    G.eff, R, dh, G.SA, ldS, 0));
    OP3_TRY(mlp_bwd(st, N, A.SA, ldS, p->dyn_action_att, g->dyn_action_att, {R + E, R, 1, ACT_SIGMOID}, A.att_h, A.att, 1,
                    G.att, 1, dh, G.SA, ldS, 1));
    OP3_TRY(mlp_bwd(st, N, A.arep, d->A, p->dyn_action_enc, g->dyn_action_enc, {d->A, R, E, ACT_ELU}, A.ae_h, A.SA + R, ldS,
                    G.SA + R, ldS, dh, nullptr, 0, 0));
    OP3_TRY(mlp_bwd(st, N, z, R, p->dyn_inertia, g->dyn_inertia, {R, R, R, ACT_ELU}, A.ie_h, A.SA, ldS, G.SA, ldS, dh, dz, R, 0));
  } else {
    OP3_TRY(mlp_bwd(st, N, z, R, p->dyn_inertia, g->dyn_inertia, {R, R, R, ACT_ELU}, A.ie_h, A.SE, ldS, G.SE, ldS, dh, dz, R, 0));
  }
  return OP3_OK;
}

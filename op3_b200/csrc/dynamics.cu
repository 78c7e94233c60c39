// Pairwise K x K slot-interaction dynamics: replaces PhysicsNetwork.forward (op3/torch/op3_modules/physics_network.py:94-141),
// Mlp.forward (op3/torch/networks.py:63-74) and get_all_attention_values (physics_network.py:156-187), plus backward.
// Eleven one-hidden-layer MLPs chained over M = B*K (per slot) or B*K*(K-1) (per ordered pair) rows.
#include "layout.cuh"

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

}  // namespace
}  // namespace op3

using namespace op3;

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

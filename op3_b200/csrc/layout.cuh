// Layout of the "prepared" (derived-weight) buffer and of its gradient twin.  See op3_prepare in op3_b200.h.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace op3 {

constexpr int NCLS = 25;          // border classes of a 5x5/pad-2 conv over a spatially constant input
constexpr int REF_CIN_PAD = 20;   // refinement conv-1 input: 5 float4 chunks (dec_out, mix1, mix2, smp, coords)

struct PrepLayout {
  size_t wsum;       // [25*32][R]        decoder L1 class-summed weights      (grad twin: d wsum)
  size_t biasT;      // [25*32]           decoder L1 bias replicated per class (grad twin: column sums of dT)
  size_t cmap;       // [8][D][D] float4  decoder L1 coordinate map            (grad twin: d cmap)
  size_t dec_wf[4];  // decoder layers 2..5 forward  [25][32][cout]            (grad twin: d wf)
  size_t dec_bf[4];  // biases of layers 2..5                                   (grad twin: d bias)
  size_t dec_wd[4];  // decoder layers 2..5 dgrad    [25][cout][32]
  size_t ref_wf[3];  // refinement layers forward [25][20|32|32][32|32|Rs]     (grad twin: d wf)
  size_t ref_bf[3];  // refinement conv biases                                  (grad twin: d bias)
  size_t ref_wd[3];  // refinement dgrad [25][32][20], [25][32][32], [25][Rs][32]
  size_t coords;     // [D][D] float4 (x, y, 0, 0)
  // tcgen05 weight slabs (conv5x5_tc.cu): [K group][tap][k-chunk][row][4], forward and dgrad orientation
  size_t dec_tf[4], dec_td[4], ref_tf[3], ref_td[3];
  size_t total;
};

inline PrepLayout prep_layout(const Op3Dims* d) {
  PrepLayout L;
  const int R = d->Rd + d->Rs, D = d->D;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };  // keep 16-byte alignment
  L.wsum = take((size_t)NCLS * 32 * R);
  L.biasT = take(NCLS * 32);
  L.cmap = take((size_t)32 * D * D);
  const int dec_cout[4] = {32, 32, 32, 4};
  for (int i = 0; i < 4; ++i) L.dec_wf[i] = take((size_t)25 * 32 * dec_cout[i]);
  for (int i = 0; i < 4; ++i) L.dec_bf[i] = take(dec_cout[i]);
  for (int i = 0; i < 4; ++i) L.dec_wd[i] = take((size_t)25 * dec_cout[i] * 32);
  const int ref_cin[3] = {REF_CIN_PAD, 32, 32};
  const int ref_cout[3] = {32, 32, d->Rs};
  for (int i = 0; i < 3; ++i) L.ref_wf[i] = take((size_t)25 * ref_cin[i] * ref_cout[i]);
  for (int i = 0; i < 3; ++i) L.ref_bf[i] = take(ref_cout[i]);
  for (int i = 0; i < 3; ++i) L.ref_wd[i] = take((size_t)25 * ref_cout[i] * ref_cin[i]);
  L.coords = take((size_t)4 * D * D);
  for (int i = 0; i < 4; ++i) L.dec_tf[i] = take(conv5x5_tc_weight_floats(32, dec_cout[i]));
  for (int i = 0; i < 4; ++i) L.dec_td[i] = take(conv5x5_tc_weight_floats(dec_cout[i], 32));
  for (int i = 0; i < 3; ++i) L.ref_tf[i] = take(conv5x5_tc_weight_floats(ref_cin[i], ref_cout[i]));
  for (int i = 0; i < 3; ++i) L.ref_td[i] = take(conv5x5_tc_weight_floats(ref_cout[i], ref_cin[i]));
  L.total = o;
  return L;
}

// Channel order of the refinement conv input as this library stores it (5 float4 chunks) -> reference channel
// index of `a` in op3_model.py:219-229 (+ coords 15,16 of refinement_network.py:192-193); -1 = zero padding.
__host__ __device__ inline int ref_l1_channel(int cip) {
  switch (cip) {
    case 0: return 3;  case 1: return 4;  case 2: return 5;  case 3: return 7;     // dec_out: colors rgb, mask_logits
    case 4: return 6;  case 5: return 8;  case 6: return 9;  case 7: return 11;    // mix1: mask, posterior_mask, LN(mask_grad), LN(leave_out)
    case 8: return 12; case 9: return 13; case 10: return 14;                      // mix2: LN(x_hat_grad) rgb
    case 12: return 0; case 13: return 1; case 14: return 2; case 15: return 10;   // smp: image rgb, LN(pixel ll)
    case 16: return 15; case 17: return 16;                                        // coords x, y
    default: return -1;
  }
}

// internal entry points implemented in decoder.cu / refine.cu and used across translation units
inline int validate_dims(const Op3Dims* d) {
  if (!d) { set_error("dims is NULL"); return OP3_ERR_ARG; }
  if (d->Rs != 64) { set_error("only sto_repsize=64 is built (got %d)", d->Rs); return OP3_ERR_ARG; }
  if (d->Rd != 64 && d->Rd != 0) { set_error("det_repsize must be 64 or 0 (got %d)", d->Rd); return OP3_ERR_ARG; }
  if (d->A != 0 && d->A != 4) { set_error("action size must be 0 or 4 (got %d)", d->A); return OP3_ERR_ARG; }
  if (d->D % 8 != 0 || d->D < 8) { set_error("image side D=%d must be a positive multiple of 8", d->D); return OP3_ERR_ARG; }
  if (d->K < 1 || d->K > 16) { set_error("K=%d out of range [1,16]", d->K); return OP3_ERR_ARG; }
  if (d->B < 1) { set_error("B=%d must be >= 1", d->B); return OP3_ERR_ARG; }
  return OP3_OK;
}

}  // namespace op3

/* op3_b200 -- C ABI of the sm_100a kernels behind the OP3 iterative amortized-inference step.
 *
 * The reference (jcoreyes/OP3) is pure Python: it has NO FFI.  Its "operator interface" for this path is the
 * nn.Module surface of op3/torch/op3_modules/{op3_model,decoder_network,refinement_network,physics_network}.py
 * and op3_trainer.py.  Each entry point below names the reference call it replaces (file:line relative to the
 * reference checkout).  The Python mirror of those modules (op3_b200/torch/...) binds these symbols with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 data owned by the caller (PyTorch); the library allocates nothing
 *    persistent and keeps no pointers across calls; all work is enqueued on `stream` (a cudaStream_t) with no
 *    host synchronisation;
 *  - every function returns 0 (OP3_OK) or a negative error code; the message is available from
 *    op3_last_error() (thread local);
 *  - "f4 planes" = activations stored as [n][c/4][D][D] float4 (4 channels interleaved per pixel);
 *  - N = B*K slot images; R = Rd+Rs latent size; latents are rows [deter(Rd) | samples(Rs)].
 */
#ifndef OP3_B200_H_
#define OP3_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OP3_OK 0
#define OP3_ERR_ARG (-1)
#define OP3_ERR_CUDA (-2)

#define OP3_LSTM 128    /* refinement_network.py:41  lstm_size        */
#define OP3_HID 128     /* refinement_network.py:38  hidden_sizes[0]  */
#define OP3_AENC 32     /* physics_network.py:18     action_enc_size  */
#define OP3_CH 32       /* decoder/refinement conv width (decoder_network.py:19, refinement_network.py:37) */

typedef struct Op3Dims {
  int B;       /* local batch                                                        */
  int K;       /* slots                                                              */
  int D;       /* image side (64 or 128; multiple of 8)                               */
  int Rd;      /* deterministic latent size (64 or 0)                                */
  int Rs;      /* stochastic latent size (64)                                        */
  int A;       /* action size seen by the dynamics net: 0 or 4                       */
  float beta;  /* KL weight (op3_model.py:80)                                        */
} Op3Dims;

/* One linear layer: weight [out][in] row major (nn.Linear), bias [out]. */
typedef struct Op3Linear { float* w; float* b; } Op3Linear;
/* networks.py:18 Mlp with one hidden layer: fcs.0 then last_fc. */
typedef struct Op3Mlp { Op3Linear fc0; Op3Linear last; } Op3Mlp;

/* Pointers to every tensor of the reference state_dict, PyTorch layouts.
 * The same struct type, pointing into the gradient buffers, receives parameter gradients (accumulated). */
typedef struct Op3Params {
  float* inital_deter_state;  /* [Rd]  (sic, op3_model.py:92) */
  float* initial_lambdas1;    /* [Rs] */
  float* initial_lambdas2;    /* [Rs] */
  float* ref_conv_w[3];       /* [32][17][5][5], [32][32][5][5], [Rs][32][5][5] */
  float* ref_conv_b[3];
  Op3Linear ref_fc;           /* [128][Rs]                       refinement_net.fc_layers.0 */
  float* lstm_w_ih;           /* [512][128+5*Rs]                 gate order i,f,g,o */
  float* lstm_w_hh;           /* [512][128] */
  float* lstm_b_ih;           /* [512] */
  float* lstm_b_hh;           /* [512] */
  Op3Linear ref_last_fc;      /* [Rs][128] -> lambdas1 */
  Op3Linear ref_last_fc2;     /* [Rs][128] -> lambdas2 */
  Op3Mlp dyn_inertia;         /* R -> R -> R */
  Op3Mlp dyn_action_enc;      /* A -> R -> 32           (A > 0 only) */
  Op3Mlp dyn_action_effect;   /* R+32 -> R -> R         (A > 0 only) */
  Op3Mlp dyn_action_att;      /* R+32 -> R -> 1         (A > 0 only) */
  Op3Mlp dyn_pairwise;        /* 2R -> 2R -> R */
  Op3Mlp dyn_inter_effect;    /* R -> R -> 32 */
  Op3Mlp dyn_inter_att;       /* R -> R -> 1 */
  Op3Mlp dyn_merge;           /* R+32 -> R -> R */
  Op3Mlp dyn_det_out;         /* R -> R -> Rd           (Rd > 0 only) */
  Op3Mlp dyn_lam1_out;        /* R -> R -> Rs */
  Op3Mlp dyn_lam2_out;        /* R -> R -> Rs */
  float* dec_conv_w[5];       /* [32][R+2][5][5], [32][32][5][5] x3, [4][32][5][5] */
  float* dec_conv_b[5];
  float* ln2d_gamma[4];       /* sizes 1,1,1,3: mask_grad, P, leave_out, x_hat_grad (op3_model.py:83-84) */
  float* ln2d_beta[4];
  float* ln1d_scale[2];       /* [Rs] each: lambdas_grad_1, lambdas_grad_2 (op3_model.py:85-86) */
  float* ln1d_center[2];
} Op3Params;

const char* op3_version(void);
const char* op3_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long op3_launch_count(void);

/* Arithmetic of the 5x5 convolutions (decoder layers 2-5, refinement conv stack, forward and dgrad):
 *   0 = fp32 CUDA-core kernels (exact mode, bit-faithful fp32 products),
 *   1 = tcgen05 3xTF32 operand splitting (fp32-grade; parity mode),
 *   2 = tcgen05 1xTF32 (speed mode; operands rounded to 11-bit significands),
 *   3 = tcgen05 fp16 hi/lo operand split with per-image power-of-two scaling for forward/dgrad (fp32-grade like mode 1
 *       at half the operand bytes per FLOP); weight gradients use mode 1.
 * Process-wide; call op3_prepare again after changing it (the derived weights depend on the mode). */
int op3_set_precision(int mode);
int op3_get_precision(void);

/* The dynamics network and the head of the refinement network run as CTA-resident chains of layers (one launch forward, one
 * backward + one batched weight-gradient launch); 0 selects the layer-by-layer launches instead (A/B tests).  Default 1. */
int op3_set_fused_mlp(int on);
/* 1: the quint-tap convolution kernel of the fp16 mode runs as clusters of two CTAs sharing the weight operand of one
 * tcgen05.mma.cta_group::2 stream (bit-identical results; its MMA stream is 27 % faster but the kernel is not, yet: its
 * loaders and epilogue are issue-bound).  Default 0 = the single-CTA form. */
int op3_set_tcq_pair(int on);
/* Weight-gradient lane.  1: the convolution weight-gradient kernels of op3_decoder_bwd / op3_refine_bwd are issued to a
 * per-device side stream, ordered after the work already issued to the call's stream (event edge; under stream capture they
 * become a parallel branch of the graph), so they fill the SMs that the small kernels of the dgrad chain leave idle.  The
 * caller then owns the join: op3_wgrad_join(stream) makes `stream` wait for the lane, and must be called before any buffer
 * handed to those calls (acts, d_out, scratch, pg) is overwritten, freed or read back, before op3_finalize_grads, and before a
 * capture ends.  Gradients are bit-identical to the serial order (one FIFO lane per device).  Default 0: everything on the
 * call's stream, no join needed.  The reference has no counterpart (autograd runs op3_model.py's backward on one stream). */
int op3_set_wgrad_overlap(int on);
int op3_wgrad_join(void* stream);

/* Optional timing of kernel families with CUDA events on the launch stream (used by bench.py for the roofline
 * numbers).  Families: 0 = conv5x5 forward/dgrad, 1 = conv5x5 wgrad, 2 = mixture kernels, 3 = small GEMMs.
 * collect() synchronises on the recorded events and returns, per family, the summed milliseconds, the summed
 * algorithmic work (FLOPs, or bytes for family 2) and the number of timed calls; it also resets the recorder. */
#define OP3_PROFILE_FAMILIES 4
int op3_profile_enable(int on);
int op3_profile_collect(double* ms, double* work, long long* launches);

/* ---- derived weights -------------------------------------------------------------------------------------
 * Re-laid-out / algebraically collapsed copies of the conv weights (tap-major layouts for forward and dgrad, the
 * 25 border-class sums + coordinate map of decoder layer 1).  Must be refreshed after every optimizer step.
 * The same layout is used for the gradient accumulators `prep_grad` that op3_*_bwd add into; op3_finalize_grads
 * folds them back into PyTorch-layout gradients. */
size_t op3_prepared_floats(const Op3Dims* d);
int op3_prepare(const Op3Dims* d, const Op3Params* p, float* prepared, void* stream);
int op3_finalize_grads(const Op3Dims* d, const float* prep_grad, const Op3Params* grads, void* stream);

/* ---- decoder: replaces DecoderNetwork.forward (decoder_network.py:83-88) + BroadcastCNN.forward
 * (conv_networks.py:326-350) and their backward.  z [N][R] -> out [N][D][D] float4 = (r,g,b,mask_logit).
 * `acts` (op3_decoder_acts_floats) keeps the saved activations for the backward passes. */
size_t op3_decoder_acts_floats(const Op3Dims* d, int N);
size_t op3_decoder_scratch_floats(const Op3Dims* d, int N);
int op3_decoder_fwd(const Op3Dims* d, int N, const float* prepared, const float* z, float* acts, float* out,
                    void* stream);
/* d_out [N][D][D] float4 -> dz [N][R] (added to dz if accumulate_dz); if prep_grad != NULL also accumulates weight
 * gradients.
 * (the dz-only form is the torch.autograd.grad call of OP3Model.refine, op3_model.py:212-214) */
int op3_decoder_bwd(const Op3Dims* d, int N, const float* prepared, const float* z, const float* acts,
                    const float* d_out, float* scratch, float* dz, int accumulate_dz, float* prep_grad, void* stream);

/* ---- mixture: replaces F.softmax over K (op3_model.py:306), OP3Model.get_loss (:313-327), _gaussian_prob
 * (:123-126), kl_divergence_prior_post (:137-145) and the closed-form, detached gradient inputs + LayerNorm2D of
 * OP3Model.refine (:209-229).  See DESIGN.md for the byte model. */
typedef struct Op3MixtureIO {
  const float* dec_out;     /* [N][D][D] float4 (r,g,b,logit)                                         in  */
  const float* image;       /* target frame, planar [3][D][D] per sample (NULL: outputs only, no loss) in  */
  long long image_bstride;  /* floats between consecutive samples of `image`                              */
  const float* mse_image;   /* frame the reconstruction error is measured against (op3_model.py:412), or NULL */
  long long mse_image_bstride;
  const float* post_l1;     /* [B][K][Rs] posterior lambdas1 / lambdas2                               in  */
  const float* post_l2;
  const float* prior_l1;    /* [B][K][Rs]                                                             in  */
  const float* prior_l2;
  float* mask;              /* [B][K][D][D] softmax masks, planar (NULL = skip)                       out */
  float* colors;            /* [B][K][3][D][D] planar colours (NULL = skip)                           out */
  float* final_recon;       /* [B][3][D][D] sum_k colors*mask (NULL = skip)                           out */
  float* stats;             /* op3_mixture_stats_floats(): LN statistics, partials, losses            out */
  float* losses;            /* [4]: clog, kl, total (unweighted, op3_model.py:321-326), recon mse       out */
  float* mix1;              /* [N][D][D] float4 (mask, posterior_mask, LN(mask_grad), LN(leave_out))  out, NULL = no refine inputs */
  float* mix2;              /* [N][D][D] float4 (LN(x_hat_grad) rgb, 0)                               out */
  float* smp;               /* [B][D][D] float4 (image rgb, LN(pixel likelihood))                     out */
  float* seed;              /* [N][D][D] float4 d total / d dec_out (x_hat_grad rgb, dlogit)          out */
  unsigned* sync;           /* optional: ONE device word, zero before the first call and left zero by every call, private
                               to the calling stream.  When given (and the decode of one sample fits the shared memory of
                               an 8-CTA cluster) the forward is ONE launch: statistics, losses and refinement inputs from
                               pixels staged once in shared memory.  NULL: statistics + emit kernels (two passes).   */
} Op3MixtureIO;
size_t op3_mixture_stats_floats(const Op3Dims* d);
int op3_mixture_fwd(const Op3Dims* d, const Op3Params* p, const Op3MixtureIO* io, int want_loss, void* stream);
/* Backward of the above w.r.t. dec_out.  w_clog = d(total)/d(clog of this decode) (0 if this decode carries no
 * loss).  d_in = gradient of the refinement conv input chunks [N][5][D][D] float4 in the chunk order
 * (dec_out, mix1, mix2, smp, coords) or NULL.  LayerNorm2D gamma/beta gradients are accumulated into `grads`. */
int op3_mixture_bwd(const Op3Dims* d, const Op3MixtureIO* io, float w_clog, const float* d_in, float* d_dec_out,
                    const Op3Params* grads, float* scratch, void* stream);
size_t op3_mixture_bwd_scratch_floats(const Op3Dims* d);

/* ---- latent state helpers: get_samples / _softplus_to_std (op3_model.py:130-153) and the KL gradient. */
/* z[:, Rd:] = eps * std(l2) + l1  (z has row stride R); if deter_src != NULL also z[:, :Rd] = deter_src[:, :Rd]
 * (refine keeps the deterministic part of the state, op3_model.py:257) */
int op3_sample_fwd(const Op3Dims* d, int N, const float* l1, const float* l2, const float* eps, const float* deter_src,
                   float* z, void* stream);
/* Folds d(samples) and the KL term into lambda gradients:
 *   dl1 += dz[:,Rd:] + w_kl * dKL/dl1 ; dl2 += dz[:,Rd:]*eps*dstd/dl2 + w_kl * dKL/dl2 ;
 *   dprior1/2 += w_kl * dKL/dprior (if non-NULL).  dz may be NULL. */
int op3_state_bwd(const Op3Dims* d, int N, const float* l1, const float* l2, const float* eps, const float* prior1,
                  const float* prior2, const float* dz, float w_kl, float* dl1, float* dl2, float* dprior1,
                  float* dprior2, void* stream);

/* ---- refinement network: replaces RefinementNetwork.forward (refinement_network.py:189-219) plus the
 * LayerNorm (modules.py:19-46) of the lambda gradients and the assembly of extra_input (op3_model.py:232-249). */
typedef struct Op3RefineIO {
  const float* dec_out;   /* conv input chunk 0: [N][D][D] float4                 */
  const float* mix1;      /* chunk 1                                              */
  const float* mix2;      /* chunk 2                                              */
  const float* smp;       /* chunk 3: [B][D][D] float4                            */
  const float* dz;        /* [N][R] decoder latent gradient of the unweighted loss */
  const float* z;         /* [N][R] current latent rows (samples in cols Rd..)     */
  const float* l1;        /* [N][Rs] posterior lambdas1 / lambdas2                 */
  const float* l2;
  const float* eps;       /* [N][Rs] noise that produced z's samples               */
  const float* prior1;    /* [N][Rs]                                               */
  const float* prior2;
  int prior_is_post;      /* prior tensors alias the posterior (state fresh out of a dynamics step) */
  const float* h;         /* [N][128] LSTM h, c                                    */
  const float* c;
  float* acts;            /* op3_refine_acts_floats()                              */
  float* l1_out;          /* [N][Rs] new lambdas1                                  */
  float* l2_out;          /* [N][Rs] new lambdas2                                  */
  float* h_out;           /* [N][128]                                              */
  float* c_out;
} Op3RefineIO;
size_t op3_refine_acts_floats(const Op3Dims* d);
size_t op3_refine_scratch_floats(const Op3Dims* d);
int op3_refine_fwd(const Op3Dims* d, const Op3Params* p, const float* prepared, const Op3RefineIO* io, void* stream);
/* Standalone RefinementNetwork.forward(input, hidden1, hidden2, extra_input) (refinement_network.py:189-219).
 * in_chunks [N][5][D][D] float4: the 15 input channels + 2 coordinate planes packed in this library's chunk order
 * (chunk 0: colors rgb, mask_logits; 1: mask, posterior_mask, LN(mask_grad), LN(leave_out); 2: LN(x_hat_grad) rgb, 0;
 * 3: image rgb, LN(pixel ll); 4: x, y, 0, 0); extra [N][5*Rs]; h, c [N][128]. */
int op3_refine_net_fwd(const Op3Dims* d, const Op3Params* p, const float* prepared, const float* in_chunks,
                       const float* extra, const float* h, const float* c, float* acts, float* l1_out, float* l2_out,
                       float* h_out, float* c_out, void* stream);
/* upstream: d_l1_new, d_l2_new [N][Rs], d_h, d_c [N][128] (any may be NULL = zero).  Outputs: d_in [N][5][D][D] float4
 * (conv input chunk gradients), and ADDS the direct paths into dl1, dl2 [N][Rs], dz_samples (into dz_acc[:,Rd:],
 * row stride R), dh, dc [N][128].  Weight gradients are accumulated into grads / prep_grad. */
int op3_refine_bwd(const Op3Dims* d, const Op3Params* p, const float* prepared, const Op3RefineIO* io,
                   const float* d_l1_new, const float* d_l2_new, const float* d_h, const float* d_c, float* d_in,
                   float* dl1, float* dl2,
                   float* dz_acc, float* dh, float* dc, const Op3Params* grads, float* prep_grad, float* scratch,
                   void* stream);

/* ---- dynamics: replaces PhysicsNetwork.forward (physics_network.py:94-141) and get_all_attention_values
 * (:156-187).  z [N][R], actions [B][A] (already reduced to the 4 used columns) or NULL.
 * out_z: deter written to cols [0,Rd) of rows with stride R (NULL if Rd == 0); l1_out, l2_out [N][Rs]. */
size_t op3_dynamics_acts_floats(const Op3Dims* d);
int op3_dynamics_fwd(const Op3Dims* d, const Op3Params* p, const float* z, const float* actions, float* acts,
                     float* out_z, float* l1_out, float* l2_out, void* stream);
/* probes [N][3 + (K-1)]: state-action attention, |enc - enc_flat| sum, pair attentions (read from acts) */
int op3_dynamics_probes(const Op3Dims* d, const float* acts, float* act_att, float* pair_att, float* deltas,
                        void* stream);
int op3_dynamics_bwd(const Op3Dims* d, const Op3Params* p, const float* z, const float* actions, const float* acts,
                     const float* d_out_z /* [N][R], cols [0,Rd) used */, const float* d_l1 /* [N][Rs] */,
                     const float* d_l2 /* [N][Rs] */, float* dz /* [N][R] written */, const Op3Params* grads,
                     float* scratch, void* stream);
size_t op3_dynamics_scratch_floats(const Op3Dims* d);

/* ---- trainer step: clip_grad_norm_(5.0) + Adam (op3_trainer.py:136,188-189), fused over flat buffers.
 * state[0] receives the pre-clip gradient norm; `scratch` >= 1024 floats. */
int op3_adam_clip(long long n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int step,
                  float lr, float beta1, float beta2, float eps, float max_norm, float grad_scale, float* scratch,
                  void* stream);

/* ---- MPC cost: Cost.image_mse + min over K (mpc_stack.py:160-184, 225-228).
 * pred [Na][Kp][3][D][D] planar, goal [G][3][D][D] -> cost [G][Na] = min_k mean((goal_g - pred_a,k)^2),
 * argk [G][Na] (int).  Fixed reduction order: results are bit-reproducible run to run. */
int op3_mpc_cost(int Na, int Kp, int G, int D, const float* pred, const float* goal, float* cost, int* argk,
                 void* stream);
/* General form (mpc_stack.py:60-100): rows of n_el elements (3*D*D for sub-images, R for core_type 'latent');
 * compare 0 = 'mse', 1 = 'psuedo_intersect'; post 0 = 'raw', 1 = 'negative_exp' (applied before the min over Kp). */
int op3_mpc_cost_ex(int Na, int Kp, int G, int n_el, int compare, int post, const float* pred, const float* goal, float* cost,
                    int* argk, void* stream);
/* Standalone OP3Model.get_loss (op3_model.py:313-327): colors [B][K][3][D][D], mask [B][K][1][D][D] (post-softmax),
 * image [B][3][D][D], lambdas [B*K][Rs] -> color_probs [B][K][D][D], pixel_ll [B][D][D] (pixel_complete_log_likelihood),
 * losses[3] = (kle_loss, complete_log_likelihood, total_loss).  scratch: op3_get_loss_scratch_floats, 8-byte aligned. */
size_t op3_get_loss_scratch_floats(const Op3Dims* d);
int op3_get_loss(const Op3Dims* d, const float* colors, const float* mask, const float* image, const float* post_l1,
                 const float* post_l2, const float* prior_l1, const float* prior_l2, float* color_probs, float* pixel_ll,
                 float* losses, float* scratch, void* stream);

/* CEM refit of the action distribution on the F best candidates (mpc_stack.py:283-288, mpc_pickplace.py:376-389):
 * mean[j], std[j] = numpy mean / population std (ddof 0) of actions[order[0..F)][j], std += min_std (0.05 in pickplace).
 * actions [Na][TA] (TA = T*A flattened), order = candidate indices sorted by cost (int64, as torch.sort returns). */
int op3_cem_refit(int Na, int TA, int F, const float* actions, const long long* order, float min_std, float* mean,
                  float* std, void* stream);
/* out[a][j] = clip(mean[j] + std[j]*noise[a][j], lo[j], hi[j]) -- env.sample_multiple_action_gaussian
 * (op3/envs/blocks/mujoco/block_pick_and_place.py:390-407) with caller-provided standard-normal noise; lo/hi NULL = no clip */
int op3_cem_resample(int Na, int TA, const float* mean, const float* std, const float* noise, const float* lo,
                     const float* hi, float* out, void* stream);

/* Data feed: uint8 frames -> planar fp32 / 255 on the device (OP3Trainer.prepare_tensors, op3_trainer.py:148-153, fused
 * with load_dataset's channels-last -> channels-first move, train_op3.py:27-28,46-47).  in: [n_frames][3][D][D]
 * (channels_last = 0) or [n_frames][D][D][3] (channels_last = 1) bytes; out: [n_frames][3][D][D] floats. */
int op3_prepare_frames(long long n_frames, int D, int channels_last, const unsigned char* in, float* out, void* stream);

/* One layer of Mlp.forward (op3/torch/networks.py:63-74): y[M][out] = act(x[M][in] W^T + b), W [out][in] (nn.Linear).
 * act: 0 identity, 1 ELU, 2 sigmoid, 3 ReLU, 4 tanh. */
int op3_linear_fwd(int M, int in, int out, const float* x, const float* w, const float* b, int act, float* y, void* stream);

/* sub_images[a][k][c] = colors*mask (op3_model.py:612) from planar colours / masks */
int op3_sub_images(long long n_slots, int D, const float* colors, const float* mask, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OP3_B200_H_ */

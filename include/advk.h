/* advk.h -- C ABI of libadvchain_b200.so (hand-written sm_100a CUDA kernels for the advchain
 * chained adversarial-augmentation hot path).
 *
 * The reference (cherise215/advchain) has NO FFI / plugin boundary: its device work is issued
 * as PyTorch ATen calls from Python (`F.grid_sample`, `F.affine_grid`, `F.interpolate`,
 * `F.conv_transposeNd`, depthwise `nn.ConvNd`, `.inverse()`).  Each entry
 * point below therefore cites the reference *call site(s)* whose ATen work it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a maintainer
 * would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless marked "host"; fp32 everywhere
 *   - every launch function returns 0 on success, <0 on error (advk_last_error() has the text);
 *     it never allocates, never synchronises, and enqueues on `stream` (a cudaStream_t)
 *   - tensors are contiguous, batch-major, torch layout  N x C x [D x] H x W
 *   - x <-> W (last axis), y <-> H, z <-> D, exactly like torch's grid_sample / the
 *     reference's base grid (adv_morph.py:14-55)
 *   - "field" = dense sampling grid in normalised [-1,1] coordinates stored INTERLEAVED per voxel:
 *     float2 (x,y) for d==2, float4 (x,y,z,0) for d==3; N*S elements (S = D*H*W)
 *   - gradient outputs documented as "accumulated" must be zero-initialised by the caller
 */
#ifndef ADVK_H
#define ADVK_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADVK_ABI_VERSION 3

enum { ADVK_OK = 0, ADVK_ERR_ARG = -1, ADVK_ERR_UNSUPPORTED = -2, ADVK_ERR_CUDA = -3 };
enum { ADVK_PAD_ZEROS = 0, ADVK_PAD_BORDER = 1, ADVK_PAD_REFLECTION = 2 };
enum { ADVK_INTERP_LINEAR = 0, ADVK_INTERP_NEAREST = 1,
       ADVK_INTERP_BICUBIC = 2 /* single-stage 2-D warps only (F.grid_sample mode="bicubic") */ };

/* PGD update modes (advk_pgd_update) */
enum {
  ADVK_UPD_L2_ASCENT = 0,  /* p += step * g/(||g||_2+1e-20)   adv_noise.py:51-64, adv_bias.py:139-148, adv_morph.py:501-516 */
  ADVK_UPD_SIGN_ASCENT = 1,/* p += step * sign(g)             adv_affine.py:182-198 */
  ADVK_UPD_L2_POWER = 2,   /* p  = g/(||g||_2+1e-20)          power iteration branches of the above */
  ADVK_UPD_SIGN_POWER = 3  /* p  = sign(g) */
};

typedef struct advk_geom {
  int d;        /* spatial dims: 2 or 3 */
  int N;        /* batch */
  int D, H, W;  /* torch spatial order; D must be 1 when d == 2 */
} advk_geom;

/* adv_affine.py:73-105 config; 2-D uses rot[0] only. */
typedef struct advk_affine_cfg {
  int d;
  float rot[3];   /* rot (2-D) | rot_x, rot_y, rot_z */
  float scale[3]; /* scale_x, scale_y, scale_z */
  float shift[3]; /* shift_x, shift_y, shift_z */
} advk_affine_cfg;

/* adv_morph.py:247-258, 391-421: low-res velocity lattice + the 1-D factor of the separable,
 * sum-normalised Gaussian (9 taps for sigma=1). */
typedef struct advk_morph_cfg {
  int lr[3];       /* low-res lattice in torch order (Dl,Hl,Wl); Dl = 1 for d == 2 */
  int ktaps;       /* odd, <= 15 */
  float gauss[15]; /* taps[0..ktaps-1], sum 1 */
} advk_morph_cfg;

/* adv_bias.py:202-335 geometry. `A[ax]` are dense (low[ax] x n_cp[ax]) row-major matrices
 * (device pointers) that fold conv_transpose(kernel, stride, padding) + crop along one axis:
 * low = A_D (x) A_H (x) A_W applied to the control points. */
typedef struct advk_bias_cfg {
  int n_cp[3];        /* control-point lattice, torch order (D,H,W); 1 for the unused axis */
  int low[3];         /* low-res field size, torch order */
  const float* A[3];  /* device pointers */
  float up_scale[3];  /* ATen source-index scale per axis for the linear upsample (align_corners=False) */
  int upsample;       /* 0: low-res field is already full-res */
  int use_log;        /* bias = exp(field) (1) or 1 + field (0)          adv_bias.py:331-334 */
  float magnitude;    /* clip to [1-m, 1+m]                               adv_bias.py:337-356 */
} advk_bias_cfg;

int advk_abi_version(void);
const char* advk_last_error(void);                 /* thread-local, valid until the next call */
int advk_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes); /* host ptrs */

/* ---- launch accounting / live kernel timing (measurement support; no reference counterpart) --
 * advk_launch_count: number of kernels this library launched (kid < 0: all kernels) since the
 *   last reset.  advk_prof_configure(kid, capacity): bracket every launch of kernel `kid`
 *   (-1: every kernel, -2: off) with a pair of CUDA events recorded on the launching stream, up to
 *   `capacity` records.  advk_prof_collect synchronises on the recorded events, writes
 *   (kernel id, milliseconds) per record to the host arrays, resets the record list and returns
 *   the number of records (<0 on error).  advk_prof_group_runs(1): when ONE kernel is profiled, a run of
 *   back-to-back launches of it on one stream is bracketed by one event pair (returns the previous
 *   setting); advk_prof_collect_runs additionally returns the launches each record covers. */
int advk_kernel_count(void);
const char* advk_kernel_name(int kid);
unsigned long long advk_launch_count(int kid, int reset);
int advk_prof_configure(int kid, int capacity);
int advk_prof_collect(int* kernel_ids, float* ms, int max_records);  /* host ptrs */
int advk_prof_group_runs(int enable);
/* programmatic dependent launch of the library's kernels (default on; ADVK_PDL=0); returns the previous setting */
int advk_set_pdl(int enable);
int advk_prof_collect_runs(int* kernel_ids, float* ms, int* launches, int max_records);  /* host ptrs */

/* ---- AdvAffine: parameters -> matrices -------------------------------------------------
 * replaces gen_batch_affine_matrix (adv_affine.py:210-273: Hardtanh, cos/sin, stack, matmul)
 * and get_inverse_matrix (adv_affine.py:316-324: batched .inverse()).
 * param: N x (5|9); theta / theta_inv: N x d x (d+1). `pscale` multiplies param first
 * (xi in power-iteration training, adv_affine.py:136-138; else 1). */
int advk_affine_theta_fwd(const advk_affine_cfg* cfg, const float* param, float pscale, int N,
                          float* theta, float* theta_inv, void* stream);
/* g_theta_inv may be NULL. g_param is written (not accumulated). */
int advk_affine_theta_bwd(const advk_affine_cfg* cfg, const float* param, float pscale, int N,
                          const float* g_theta, const float* g_theta_inv, float* g_param,
                          void* stream);

/* ---- warps ------------------------------------------------------------------------------
 * advk_warp_affine_*: F.affine_grid + F.grid_sample (adv_affine.py:297-313) with the grid
 *   generated on the fly.
 * advk_warp_field_*:  F.grid_sample with a dense field (adv_morph.py:546-557); the field is
 *   clamped to [-1,1] on load (adv_morph.py:304, 489-490) and g_field is masked accordingly.
 * src/out: N x C x S.  pad_values: NULL, or N per-sample constants v: sample (src - v) with
 * zero padding and add v back ('lowest' / numeric padding, adv_affine.py:299-311,
 * adv_morph.py:542-554).
 * bwd: g_src (nullable) is ACCUMULATED with atomics; g_theta (nullable, N x d x (d+1)) is
 * ACCUMULATED; g_field is written. */
int advk_warp_affine_fwd(const advk_geom* g, int C, const float* src, const float* theta,
                         int pad_mode, int interp, const float* pad_values, float* out,
                         void* stream);
int advk_warp_affine_bwd(const advk_geom* g, int C, const float* g_out, const float* src,
                         const float* theta, int pad_mode, int interp, const float* pad_values,
                         float* g_src, float* g_theta, void* stream);
int advk_warp_field_fwd(const advk_geom* g, int C, const float* src, const void* field,
                        int pad_mode, int interp, const float* pad_values, float* out,
                        void* stream);
int advk_warp_field_bwd(const advk_geom* g, int C, const float* g_out, const float* src,
                        const void* field, int pad_mode, int interp, const float* pad_values,
                        float* g_src, void* g_field, void* stream);

/* ---- AdvMorph: velocity -> deformation field -------------------------------------------
 * replaces DemonsCompose (adv_morph.py:454-491): depthwise Gaussian conv (low-res), F.interpolate,
 * vectorFieldExponentiation2D/3D (:116-177: n grid_sample self-compositions, border padding),
 * compose-with-base grid_sample, full-res depthwise Gaussian conv, clamp.
 * v: N x d x lr (channel 0 = x displacement, like the reference).  `scale` = +-epsilon (or +-xi).
 *
 * advk_morph_unorm2: sum over the whole batch of |upsampled smoothed velocity|^2, for the 3-D
 *   step-count rule of adv_morph.py:159-162 (the host picks nb_steps). out: 1 float, overwritten.
 * levels: (nb_steps+2) fields of N*S elements each: phi_0 .. phi_n, kept for the backward, plus one
 *   scratch field used by the forward.
 * field_out: the UNCLAMPED composed field (consumers clamp on load); N*S elements.
 * u_lr: scratch N x d x lr floats.
 * norm2_out (nullable): 1 float, receives the same sum as advk_morph_unorm2 as a by-product of the
 *   build (graph-captured loops check the step-count rule after the fact, advk_morph_steps_check). */
int advk_morph_unorm2(const advk_geom* g, const advk_morph_cfg* cfg, const float* v, float scale,
                      float* u_lr, float* out_norm2, void* stream);
int advk_morph_field_fwd(const advk_geom* g, const advk_morph_cfg* cfg, const float* v,
                         float scale, int nb_steps, float* u_lr, void* levels, void* field_out,
                         float* norm2_out, void* stream);
/* scratch: 5 fields of N*S elements (4 are used); lr_scratch: advk_morph_lr_scratch_floats() floats.
 * g_v: N x d x lr, written. g_field: gradient w.r.t. field_out, i.e. w.r.t. the UNCLAMPED field: a
 * consumer that clamps the field on load (every warp of this library; torch.clamp in user code) has already
 * zeroed it outside the clamp range (adv_morph.py:489-490). */
size_t advk_morph_lr_scratch_floats(const advk_geom* g, const advk_morph_cfg* cfg);
int advk_morph_field_bwd(const advk_geom* g, const advk_morph_cfg* cfg, float scale, int nb_steps,
                         const void* levels, const void* field_out, const void* g_field,
                         void* scratch, float* lr_scratch, float* g_v, void* stream);
/* Kernel variants of the field build (A/B timing and tests), a bit mask; 0 = defaults.
 *   bit 0: the plain predecessors of the lean squaring-step kernels (predicated forward step; one RED per
 *          corner, memset nodes).  Default: no validity predicates -- an outside corner has weight 0 under
 *          border padding and is redirected to an inside one --, x hand-off by weight before the RED, the
 *          adjoint zeroes the ping-pong buffer it consumed.
 *   bit 1: the two-launch predecessor (xy tile kernel + z column kernel) of the 3-D full-resolution Gaussian.
 *          Default: ONE launch per direction, plane tiles staged by TMA (cp.async.bulk.tensor, out-of-bounds
 *          zero fill = the convolution's zero padding) on an mbarrier ring; forward, the last squaring step
 *          writes the Gaussian's input itself.
 *   bit 2: the lean adjoint zeroes the scatter target it consumed itself (two targets).  Default: three targets
 *          in rotation, zeroed by memsets on a library-owned side stream beside the next launch (event fork /
 *          join: capturable, no host synchronisation).
 *   bit 3: force the adjoint on 32 x 8 tiles (one warp per row) with a second hand-off along y through shared
 *          memory: 2 corner REDs + the Jacobian RED per voxel instead of 4 + 1.  Default: used where the tiles are
 *          at least 97 % full (W, H multiples of 32, 8 or large), the lean linear kernel elsewhere.
 *   bit 5: force the lean linear adjoint.
 * Environment ADVK_SSB_MODE.  Results agree up to fp32 summation order.  A negative mask only queries;
 * returns the previous mask. */
int advk_morph_tune(int ssb_mode_mask);

/* ---- AdvNoise / AdvBias: intensity stage ------------------------------------------------
 * advk_bias_lowfield_*: conv_transposeNd + crop (adv_bias.py:293-307) folded into per-axis
 *   matrices (2-D: A[0] = NULL, n_cp[0] = low[0] = 1). cp: N x n_cp; low: N x low.
 *   cp_scale: xi in power-iteration training else 1.
 * advk_intensity_fwd: out = stage2(stage1(x)) with
 *     noise:  x + noise_scale*delta                               (adv_noise.py:79-90)
 *     bias:   x * clip(exp|1+ (upsample(low)))                    (adv_bias.py:313-356, 186)
 *   `order`: 0 noise only, 1 bias only, 2 noise then bias, 3 bias then noise.
 *   use_ignore/ignore_value: voxels with |x - ignore| < 1e-8 keep `ignore` (adv_noise.py:85-88,
 *   adv_bias.py:176-182; applied after each stage like the reference).
 *   bias_out (nullable): N x 1 x S clipped bias field (the reference's `bias_field` attribute).
 * advk_intensity_bwd: g_x (nullable, written), g_delta (nullable, written),
 *   g_up (nullable, written): N x S gradient w.r.t. the upsampled pre-exp field.
 * advk_bias_upsample_adjoint: g_up (N x S) -> g_low (N x low), the adjoint of the
 *   align_corners=False linear upsample; scratch: advk_bias_scratch_floats() floats. */
int advk_bias_lowfield_fwd(const advk_bias_cfg* cfg, int N, const float* cp, float cp_scale,
                           float* low, void* stream);
int advk_bias_lowfield_bwd(const advk_bias_cfg* cfg, int N, const float* g_low, float cp_scale,
                           float* g_cp, void* stream);
int advk_intensity_fwd(const advk_geom* g, int C, int order, const float* x, const float* delta,
                       float noise_scale, const float* low, const advk_bias_cfg* bias,
                       int use_ignore, float ignore_value, float* out, float* bias_out,
                       void* stream);
int advk_intensity_bwd(const advk_geom* g, int C, int order, const float* g_out, const float* x,
                       const float* delta, float noise_scale, const float* low,
                       const advk_bias_cfg* bias, int use_ignore, float ignore_value, float* g_x,
                       float* g_delta, float* g_up, void* stream);
size_t advk_bias_scratch_floats(const advk_geom* g, const advk_bias_cfg* bias);
int advk_bias_upsample_adjoint(const advk_geom* g, const advk_bias_cfg* bias, const float* g_up,
                               float* scratch, float* g_low, void* stream);

/* ---- PGD parameter update ---------------------------------------------------------------
 * replaces unit_normalize (adv_transformation_base.py:129-156) + the optimize_parameters
 * bodies cited at the ADVK_UPD_* enum.  sumsq: N doubles scratch (overwritten). `param` is
 * updated in place; for the *_POWER modes `grad` may alias `param` (rescale_parameters,
 * adv_noise.py:92-94, adv_morph.py:518-522). */
int advk_pgd_update(float* param, const float* grad, float step, int mode, int N,
                    size_t per_sample, double* sumsq, void* stream);
/* Same, with the loop's NaN/Inf guard (adv_compose_solver.py:345-346) evaluated on the device:
 * `guard` points to the step's loss (1 float); when it is NaN or Inf the parameters are left
 * unchanged.  This removes the step's only host synchronisation, so a whole PGD iteration can be
 * captured in a CUDA graph. */
int advk_pgd_update_guarded(float* param, const float* grad, float step, int mode, int N,
                            size_t per_sample, double* sumsq, const float* guard, void* stream);
/* Device-side check of the 3-D step-count rule (adv_morph.py:159-162) for graph-captured loops
 * that run with a fixed nb_steps: *violations is incremented when the rule applied to *norm2
 * (from advk_morph_unorm2) gives a different count. */
int advk_morph_steps_check(const float* norm2, int nb_steps, int min_steps, int* violations,
                           void* stream);
/* Early verdict of a graph-captured iteration (the reference reads its NaN guard and its 3-D step count on
 * the host in the middle of every iteration, adv_compose_solver.py:345, adv_morph.py:159-162; here the host
 * needs ONE word per call, and needs it before the iteration has finished so that the next call can be enqueued
 * behind the running one).  Increments *seq (device, 1 word) and stores (*seq << 8) | (*violations & 0xff) as one
 * 32-bit word to `host_word`, which must be pinned host memory that the device can address (cudaHostAlloc /
 * torch pin_memory under unified addressing).  The host polls that word; no stream synchronisation.
 * norm2 / host_norm2 (both nullable): *norm2 (device; the sum advk_morph_field_fwd left in norm2_out, i.e. what the
 * step rule was checked against) is stored to the pinned float host_norm2 before the word, so a host that sees the
 * word also sees the norm: a multi-iteration loop predicts the next iteration's count from it. */
int advk_publish_verdict(const int* violations, unsigned* seq, unsigned* host_word, const float* norm2,
                         float* host_norm2, void* stream);

/* ---- fused chain apply ---------------------------------------------------------------------
 * ONE launch applies a whole chain of transforms to an N x C x S tensor, ONE launch applies its
 * adjoint.  Replaces the per-transform loops of ComposeAdversarialTransformSolver.forward
 * (adv_compose_solver.py:148-176: t.forward per transform + the if_norm_image clamp),
 * predict_forward (:184-197), backward / predict_backward (:199-219) and the valid-region mask
 * predict_backward(predict_forward(ones)) != 0 (:321-325), i.e. the F.grid_sample /
 * F.affine_grid / elementwise ATen sequences of adv_noise.py:79-90, adv_bias.py:152-188,
 * adv_morph.py:524-558 and adv_affine.py:289-314 -- and autograd's backward of all of them.
 *
 * A chain is a program of 1..ADVK_CHAIN_MAX_STAGES stages in application order.  Each stage is
 *   INTENSITY   (noise and/or bias, `intensity_order` as in advk_intensity_fwd),
 *   WARP_FIELD  (grid_sample with a dense field, clamped to [-1,1] on load), or
 *   WARP_AFFINE (affine grid generated on the fly from theta: N x d x (d+1)).
 * The stages run as phases of one persistent cooperative grid separated by grid barriers (a
 * resampling needs the complete output of the stage before it); intermediates go to `stash`.
 * do_clamp: clamp(out, lo, hi) after the last stage (gradient passes on the closed interval).
 * want_mask: a one-channel N x S mask rides along the WARP stages: mask_src (NULL = ones) is
 *   resampled by every WARP stage with that stage's padding; the last one writes mask_out,
 *   as (value != 0) when binarize_mask.
 * The prediction warp-back is the same call with the reversed geometric stages
 * (theta_inv, field of -v), C = K classes, mask_src = the image chain's mask_out.
 *
 * advk_chain_workspace_floats: sizes of `stash` (fwd -> bwd) and `scratch` (bwd only).
 * bwd: per-stage gradient outputs are the g_* pointers of the descriptor (NULL = not wanted):
 *   g_delta N x C x S written; g_up N x S written (-> advk_bias_upsample_adjoint);
 *   g_field N x S field elements written; g_theta N x d x (d+1) ACCUMULATED (caller zeroes).
 *   g_src (nullable): gradient w.r.t. the chain input, N x C x S, written (16-byte aligned).
 * Stages before the first one with a requested gradient are skipped. */
#define ADVK_CHAIN_MAX_STAGES 6
enum { ADVK_STAGE_INTENSITY = 0, ADVK_STAGE_WARP_FIELD = 1, ADVK_STAGE_WARP_AFFINE = 2 };

typedef struct advk_chain_stage {
  int kind;
  /* INTENSITY */
  int intensity_order;
  float noise_scale;
  int use_ignore;
  float ignore_value;
  const advk_bias_cfg* bias; /* host pointer; NULL unless the stage has a bias */
  const float* delta;        /* N x C x S */
  const float* low;          /* low-res bias field from advk_bias_lowfield_fwd */
  /* WARP_* */
  const void* field;
  const float* theta;
  int pad_mode;
  int interp;
  const float* pad_values;   /* NULL or N per-sample constants, as in advk_warp_*_fwd */
  /* gradient outputs, used by advk_chain_apply_bwd only */
  float* g_delta;
  float* g_up;
  void* g_field;
  float* g_theta;
} advk_chain_stage;

typedef struct advk_chain_desc {
  advk_geom g;
  int C;
  int n_stages;
  advk_chain_stage stages[ADVK_CHAIN_MAX_STAGES];
  int do_clamp;
  float clamp_lo, clamp_hi;
  int want_mask;
  int binarize_mask;
} advk_chain_desc;

/* 1: one cooperative launch per chain pass (default when the device supports it and the
 * environment variable ADVK_CHAIN_COOP is not "0"); 0: one plain launch per stage.  Returns the
 * previous setting.  Results are identical; exists for A/B timing. */
int advk_chain_set_cooperative(int enable);
/* 1 (default; environment ADVK_CHAIN_PACK): chains made of warp stages only with C % 4 == 0 (the
 * K-class prediction path) keep their intermediates and scatter targets channel-packed, one float4
 * per voxel per 4 channels; 0: planar everywhere.  Returns the previous setting.  Results agree up
 * to fp32 summation order. */
int advk_chain_set_packed(int enable);
/* 1 (default; environment ADVK_CHAIN_LEAN): chains whose warp stages all use zeros padding, linear
 * interpolation and no pad values run the compile-time specialised per-stage kernels
 * (csrc/advk_chain_lean.cuh); 0: the generic stage executor for everything (A/B timing, tests).  Returns the
 * previous setting.  Results agree up to fp32 summation order.  The setting must not change between a
 * forward call and the backward call that consumes its stash. */
int advk_chain_set_lean(int enable);
int advk_chain_workspace_floats(const advk_chain_desc* d, size_t* stash_floats, size_t* scratch_floats);
int advk_chain_apply_fwd(const advk_chain_desc* d, const float* src, const float* mask_src,
                         float* stash, float* out, float* mask_out, void* stream);
int advk_chain_apply_bwd(const advk_chain_desc* d, const float* g_out, const float* src,
                         const float* stash, float* scratch, float* g_src, void* stream);

/* ---- consistency loss -------------------------------------------
 * replaces calc_segmentation_consistency (common/loss.py:8-87) for scales=[0] and divergence
 * types 'mse' (:55-64, incl. the second division by N*S), 'contour' (:65-79 -> contour_loss
 * :102-220, incl. the 3-D gy := gx kernel re-use) and 'kl' (:49-54 -> kl_divergence :223-249); a
 * weight of 0 disables a term.
 * output, reference: N x K x S logits (reference is used as probabilities when is_gt != 0);
 * mask: N x S (one channel, broadcast over K) or NULL.  scratch: advk_loss_scratch_floats()
 * floats, 8-byte aligned; it carries the softmax / edge maps from fwd to bwd.  loss: 1 float
 * (device).  bwd: upstream = device pointer to dL/dloss (NULL = 1); g_output N x K x S, written. */
size_t advk_loss_scratch_floats(const advk_geom* g, int K);
/* 1 (default; environment ADVK_LOSS_FUSED): the contour term and its adjoint run as ONE kernel (3-D: z-marching) in
 * the forward call (the Sobel responses stay in shared memory, the adjoint s_c is left in `scratch` for the
 * backward call); 0: the two-kernel predecessor (responses written by fwd, adjoint applied by bwd).  Same
 * results.  A negative value only queries; returns the previous setting.  Must not change between a forward
 * call and the backward call that consumes its scratch. */
int advk_loss_tune(int fused);
int advk_consistency_loss_fwd(const advk_geom* g, int K, const float* output, const float* reference,
                              const float* mask, float w_mse, float w_contour, float w_kl, int is_gt,
                              float* scratch, float* loss, void* stream);
int advk_consistency_loss_bwd(const advk_geom* g, int K, const float* mask, float w_mse,
                              float w_contour, float w_kl, int is_gt, const float* scratch,
                              const float* upstream, float* g_output, void* stream);

/* ---- solver glue ------------------------------------------------------------------------
 * advk_clamp: out = clamp(x, lo, hi)   (solver.forward if_norm_image, adv_compose_solver.py:167-175)
 * advk_clamp_bwd: g_x = g_out * [lo <= x <= hi]
 * advk_nonzero_mask: m = (x != 0) ? 1 : 0  in place (adv_compose_solver.py:325) */
int advk_clamp(const float* x, float lo, float hi, float* out, size_t n, void* stream);
int advk_clamp_bwd(const float* g_out, const float* x, float lo, float hi, float* g_x, size_t n,
                   void* stream);
int advk_nonzero_mask(float* x, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADVK_H */

/*
 * fullbatch_b200 -- C ABI of the B200 (sm_100a) kernels behind the full-batch gradient-regularised step.
 *
 * The reference (JonasGeiping/fullbatchtraining) has no native/FFI layer: its hot path is Python calling torch ops.
 * This header introduces the boundary UNDER the three Python call sites the reference exposes for the path
 * (SURVEY.md 8b).  Each entry point names the reference code whose device work it replaces (file:line relative to the
 * reference tree).  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a cudaError_t value or an FB_ERR_* code otherwise; fb_last_error() has text;
 *   - all device work is enqueued on `stream` (a cudaStream_t passed as void*), no host synchronisation, no allocation;
 *   - pointers are device pointers unless the name says host; sizes are elements unless the name says bytes;
 *   - activations are NHWC; "hi"/"lo" are the two bf16 planes of a split fp32 value (x ~= hi + lo), lo may be NULL
 *     (plain-bf16 mode); conv outputs and activation gradients are fp32 NHWC; output gradients fed to the tensor
 *     cores (dY) are plain bf16;
 *   - unsupported shapes return FB_ERR_UNSUPPORTED: there is NO fallback path.
 *
 * Microbatch groups.  The reference walks its microbatches one by one (training.py:148-173); here ONE launch of every
 * kernel serves `ng` consecutive microbatches ("groups", up to FB_MAX_GROUPS): the batch dimension of every activation
 * is ng * (images per microbatch), BatchNorm statistics, loss, gradient norms and weight gradients stay PER GROUP, and
 * in the finite-difference pass every group g uses its own perturbed parameters theta + eps_g * v_g.  Per-group
 * arguments are addressed as base + g * <...>_gstride.  Results are independent of ng bit for bit: every reduction has
 * an order that only depends on the problem of ONE group.
 *
 * Gradient layout.  Conv weight gradients live in the flat gradient buffers in the layout the tensor-core kernels
 * produce, [co][tap][ci] ("native"), at the flat offset of the parameter; everything else is in parameter order.
 * All flat-buffer arithmetic (norms, FD combine, running mean) is elementwise and layout agnostic; fb_flat_relayout
 * converts to / from the reference's OIHW order (fullbatch/training/utils.py:34) at the boundary, once per step.
 */
#ifndef FULLBATCH_B200_H
#define FULLBATCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_ERR_BAD_ARG 1001
#define FB_ERR_UNSUPPORTED 1002
#define FB_ERR_DRIVER 1003

#define FB_TMAP_BYTES 128 /* sizeof(CUtensorMap) */
#define FB_MAX_TAPS 9
#define FB_MAX_A_MAPS 8
#define FB_MAX_B_MAPS 2
#define FB_MAX_WGRAD_TAPS 9
#define FB_MAX_GROUPS 16

int fb_version(void);
/* copies the calling thread's last error text into buf (NUL terminated); returns its length */
int fb_last_error(char* buf, size_t n);

/* ---- TMA descriptors (host side; written into caller-owned 128-byte, 64-byte aligned blobs) -------------------- */

/* 4-D bf16 view (c, w, h, n) of an NHWC activation plane (or of one stride-2 phase of it): extents and element strides
 * of w/h/n (c is contiguous); box = box_c x box_w x box_h x box_n, SWIZZLE_128B (box_c must be 64), OOB -> 0. */
int fb_tmap_encode_act4d(void* host_blob, const void* base, int c, int w, int h, int n, int64_t stride_w,
                         int64_t stride_h, int64_t stride_n, int box_c, int box_w, int box_h, int box_n);
/* 2-D bf16 row-major matrix [rows][k] with row stride `ld` elements; box = box_k(64) x box_rows, SWIZZLE_128B. */
int fb_tmap_encode_mat2d(void* host_blob, const void* base, int k, int rows, int64_t ld, int box_k, int box_rows);

/* ---- convolutions as tcgen05 implicit GEMMs -------------------------------------------------------------------- */

/* One filter tap of the implicit GEMM: the A box is fetched from activation phase map `phase` shifted by (dh, dw)
 * pixels; the tap's weights start at column b_k0 of the weight matrix and run over `cblocks` 64-channel blocks. */
typedef struct {
  int8_t phase, dh, dw, pad;
  int32_t b_k0;
} fb_tap;

/* out[pixel, 0..n_total) (+)= sum_taps sum_cblocks  A_tap[pixel tile, 64] * B_tap[n_tile rows, 64]^T
 * Used for: conv forward (resnets.py:69-73,206-210,285-291 -> cuDNN fprop), dgrad of stride-1 and stride-2 convs
 * (autograd, training.py:82 / modules.py:230), 1x1 shortcut convs and the im2col'ed stem.
 * The 128-pixel M tile is the TMA box tile_w x tile_h x tile_n of the OUTPUT pixel grid (width == tile_w).
 * a_planes / b_planes = 2: operands are bf16 hi+lo pairs, accumulated as hi*hi + hi*lo + lo*hi (a=2,b=2) or
 * a*b_hi + a*b_lo (a=1,b=2).  A maps are indexed [phase * a_planes + plane], B maps [plane].
 * Groups: grid_n = ng * mg_imgs images; group g reads weight rows [g * b_group_rows + n, ...) (b_group_rows = 0: all
 * groups share the weights; the B maps then cover ng_max * b_group_rows rows). */
typedef struct {
  const void* host_a_maps; /* n_phases * a_planes x 128 B */
  const void* host_b_maps; /* b_planes x 128 B */
  int32_t n_phases, a_planes, b_planes;
  int32_t n_taps, cblocks;
  fb_tap taps[FB_MAX_TAPS];
  int32_t tile_w, tile_h, tile_n; /* product must be 128 */
  int32_t grid_h, grid_n;         /* rows and images of the output pixel grid (all groups) */
  int32_t n_total, n_tile;        /* GEMM N (output channels) and the per-CTA N tile: 64, 128 or 256 */
  float* out;                     /* fp32; element (n,h,w,c) at out + n*out_sn + h*out_sh + w*out_sw + c */
  int64_t out_sn, out_sh, out_sw;
  int32_t accumulate; /* 0: overwrite, 1: out += result */
  /* optional tap groups (0 = one group of all n_taps): tap group t covers taps [tap0, tap0 + n_taps) and writes to
   * out + out_off; every tap group runs over the same pixel grid.  One launch then serves the four output phases of a
   * stride-2 dgrad (1 + 2 + 2 + 4 taps).  Not combinable with the BatchNorm statistics. */
  int32_t n_tapgroups;
  struct {
    int32_t tap0, n_taps;
    int64_t out_off;
  } tapgroups[4];
  /* microbatch groups */
  int32_t mg_imgs;      /* images per group (0: one group of grid_n images) */
  int32_t ng;           /* number of groups in this launch, grid_n == ng * mg_imgs (0 -> 1) */
  int32_t b_group_rows; /* weight-matrix rows per group, 0 = shared */
  int32_t reverse;      /* walk the tiles back to front (start where the producer of the operands ended: L2 reuse) */
  /* optional train-mode BatchNorm statistics of `out` (resnets.py:71 / torch.nn.BatchNorm2d), fused into the epilogue:
   * per-CTA column sums / sums of squares go to stats_ws[g][row][2][n_total] (rows = fb_conv_stats_rows), and the LAST
   * CTA to finish a (group, N tile) reduces them in a fixed order: bn_mean / bn_rstd [ng][n_total] (biased variance,
   * rstd = 1/sqrt(var + bn_eps)) and, if bn_batch != NULL, bn_batch[g][0|1][n_total] = batch mean / UNBIASED variance
   * for the running-stat EMA (fb_bn_ema_multi).  tickets: ng_max * (n_total / n_tile) uint32, zero on first use
   * (self-resetting).  NULL stats_ws: off. */
  float* stats_ws;
  uint32_t* tickets;
  float *bn_mean, *bn_rstd, *bn_batch;
  float bn_eps;
  /* CTA pairs (thread-block clusters of 2, tcgen05 cta_group::2: M = 256 tiles over two SMs, each CTA fetches half of
   * every weight tile).  1: the B maps were encoded with boxes of n_tile / 2 rows and the launch uses pairs; only valid
   * where fb_conv_pair_ok(...) says so (a function of ONE group's problem: results stay independent of ng).  A pair
   * accumulates all operand-plane products in one accumulator: same products, another fp32 order than single-CTA tiles.
   * 2: clusters of two independent CTAs that fetch half of every weight tile each and multicast it to both (same maps
   * and conditions as 1; bit-identical to 0). */
  int32_t cta_pair;
  /* 1: haloed A boxes for 3x3 / stride-1 problems whose tiles are whole rows of one image (tile_n == 1).  The A maps
   * were encoded with boxes of tile_h + 2 rows, the taps come in triples (dh = -1, 0, 1 at one dw): one box fetch per
   * (dw, channel block) serves three taps (the K order becomes dw-major; still a function of one group's problem). */
  int32_t halo;
  /* Tile schedule = partial-statistics rows per group (fb_conv_stats_rows; a function of ONE group's problem and of
   * these two constants of the caller, never of ng).  policy_groups: groups per launch the schedule is tuned for
   * (<= 0: 1).  sched_k_iters: K iterations (taps x channel blocks) of one tile for launches that collect statistics
   * -- fewer, longer super-tiles amortise the statistics flush; 0: one row per CTA of a wave (dgrad). */
  int32_t policy_groups;
  int32_t sched_k_iters;
} fb_conv_gemm_args;
/* partial rows per (group, N tile) written to stats_ws for m_tiles_per_group x n_tiles tiles per group */
int fb_conv_stats_rows(int m_tiles_per_group, int n_tiles, int policy_groups, int sched_k_iters);
/* 1 if a problem with this many 128-pixel tiles per group and N tiles can run as CTA pairs (FB_CTA2=0 disables) */
int fb_conv_pair_ok(int m_tiles_per_group, int n_tiles, int policy_groups, int sched_k_iters);
int fb_conv_gemm(const fb_conv_gemm_args* args, void* stream);

typedef struct {
  int8_t phase, dh, dw;
  int8_t k_index; /* filter position this tap's gradient belongs to: column block k_index*cin (kh*k + kw) */
} fb_wgrad_tap;

/* Weight gradient (autograd wgrad, training.py:82 / modules.py:230), per group g and split s over the group's
 * 128-pixel blocks:  out[g][s][co][tap*cin + ci] = sum_{pixels of (g, s)} dY[pixel, co] * X[pixel + tap shift, ci]
 * at out + g*out_gstride + s*out_sstride + co*(n_taps*cin).  With splits == 1 `out` can be the flat gradient buffer
 * itself (native layout, no reduction pass); otherwise the splits are summed by fb_reduce_multi in a fixed order.
 * X maps are indexed [phase * planes + plane]. */
typedef struct {
  const void* host_dy_map; /* 1 x 128 B */
  const void* host_x_maps; /* n_x_maps x 128 B */
  int32_t n_x_maps, planes;
  int32_t n_taps, cblocks;
  fb_wgrad_tap taps[FB_MAX_WGRAD_TAPS];
  int32_t slots_per_cta; /* accumulators (tap, ci-block pairs) per CTA: slots_per_cta * planes <= 8 */
  int32_t cout, cin;     /* cin = 64*cblocks; row length = n_taps*cin */
  int32_t tile_w, tile_h, tile_n;
  int32_t grid_h, grid_n; /* dY pixel grid (all groups) */
  int32_t splits;         /* split-K over the 128-pixel blocks of ONE group */
  float* out;
  int64_t out_gstride, out_sstride;
  /* halo != 0 (3x3 / stride 1, tile_n == 1, tile_w * 128 B a multiple of 1024): taps are ordered in triples that share
   * dw (dh = -1, 0, +1), slots_per_cta == 3, and the X maps have boxes of tile_h + 2 rows: per pixel block and triple
   * ONE haloed X box is fetched and the three row shifts are aligned views of it. */
  int32_t halo;
  int32_t mg_imgs, ng; /* images per group / groups in this launch (0, 0: one group of grid_n images) */
} fb_wgrad_args;
int fb_conv_wgrad(const fb_wgrad_args* args, void* stream);

/* dst[g][r][c] = sum_{s < splits} src[g][s][r][c] in the fixed order s = 0, 1, ... for a TABLE of matrices (all conv
 * layers of the network in one launch): the deterministic split-K reduction of fb_conv_wgrad.  Entries live in device
 * memory; entry i owns blocks [block_start, block_start + n_blocks), n_blocks = ceil(rows*cols / (256 * vec)) with
 * vec = 4 if cols, src_ld, dst_ld and both bases are multiples of 4 elements, else 1.  dst = dst_base + dst_off. */
typedef struct {
  const float* src;
  int64_t src_gstride, src_sstride; /* elements */
  int64_t dst_off, dst_ld;
  int32_t splits, rows, cols, src_ld;
  int32_t block_start, n_blocks;
  int32_t vec, pad;
} fb_reduce_entry;
int fb_reduce_multi(const fb_reduce_entry* table_dev, int n_entries, int total_blocks, float* dst_base,
                    int64_t dst_gstride, int ng, void* stream);

/* fp32 OIHW conv weight -> bf16 hi/lo GEMM operands: wf[co][tap][ci] (forward / wgrad order) and wd[ci][tap][co]
 * (dgrad); lo pointers may be NULL.  wd_* may be NULL (first layer needs no dgrad).  ld_f / ld_d: row strides. */
int fb_weight_prep(const float* w_oihw, int cout, int cin, int taps, void* wf_hi, void* wf_lo, int64_t ld_f,
                   void* wd_hi, void* wd_lo, int64_t ld_d, void* stream);

/* All conv weights of the network in one launch, optionally at the PERTURBED point of every group
 * (GradRegularizer._forward_differences, modules.py:215-226: theta' = theta + eps_n * v, v = bs*g [+ acc*pre_grads];
 * central differences modules.py:279-286 use scale = +-0.5):
 *     w'_g = theta + (scale * eps_n[g]) * (bs * grad[g] + acc * pre)          eps_n[g] = scal[eps_base + g]
 * with grad / pre in the NATIVE gradient layout (so theta' itself is never materialised for conv weights).
 * grad == NULL: plain theta (ng must be 1).  Group g writes operand set g (base + g * w?_gstride elements).
 * Table entries live in device memory; entry i owns the blocks [block_start, block_start + n_blocks) with
 * n_blocks = (cout/32)*(cin/32) (or any count >= 1 for the stem, cin < 32, whose native layout is OIHW). */
typedef struct {
  int64_t w_offset; /* element offset of the weight inside theta / the gradient buffers */
  int32_t cout, cin, taps;
  int32_t block_start, n_blocks;
  int32_t pad;
  void *wf_hi, *wf_lo, *wd_hi, *wd_lo; /* lo / wd pointers may be NULL */
  int64_t ld_f, ld_d;
  int64_t wf_gstride, wd_gstride; /* elements between the operand sets of consecutive groups */
} fb_wprep_entry;
int fb_weight_prep_multi(const float* theta, const fb_wprep_entry* table_dev, int n_entries, int total_blocks,
                         const float* grad, int64_t grad_gstride, const float* pre, float bs, float acc, float scale,
                         const float* scal, int eps_base, int ng, void* stream);

/* ---- bandwidth-bound layer kernels ----------------------------------------------------------------------------- */

/* x [n,3,32,32] fp32 NCHW (optionally gathered through perm[first + i]) -> 3x3/pad-1 patches [n*1024][64] bf16 hi/lo,
 * column = ci*9 + kh*3 + kw (27 used).  If first_dev != NULL the first sample is first + *first_dev * cursor_stride
 * (device-side microbatch cursor).  labels_out[i] = labels[perm ? perm[first+i] : first+i]. */
int fb_stem_im2col(const float* x, const int64_t* labels, const int64_t* perm, const int32_t* first_dev, int64_t first,
                   int cursor_stride, int n, void* patches_hi, void* patches_lo, int64_t* labels_out, void* stream);

/* Same from a device-resident uint8 HWC dataset [N][32][32][3] with the reference's CIFAR training augmentation applied
 * on the fly (config/data/CIFAR10.yaml:22-26 via torchvision, data_preparation.py:173-200): RandomCrop(32, padding 4)
 * -> RandomHorizontalFlip -> ToTensor -> Normalize(mean, std).  aug (device, int8[.][4], may be NULL = no augmentation)
 * holds (dx, dy, flip, 0) per POSITION of the epoch order: crop offsets 0..8 in the zero-padded 40x40 image.
 * mean3 / std3 are host pointers. */
int fb_stem_im2col_u8aug(const uint8_t* x_hwc, const int64_t* labels, const int64_t* perm, const int32_t* first_dev,
                         int64_t first, int cursor_stride, int n, const int8_t* aug, const float* mean3,
                         const float* std3, void* patches_hi, void* patches_lo, int64_t* labels_out, void* stream);

/* Train-mode BatchNorm statistics over y[P][C] (resnets.py:71 / torch.nn.BatchNorm2d): mean, rstd = 1/sqrt(var+eps)
 * (biased var) and the running-stat EMA with unbiased variance.  Stand-alone building block (the engine takes the
 * statistics from the conv epilogue).  ws: >= 2*C*1024 floats of scratch. */
int fb_bn_stats(const float* y, int64_t P, int C, float* ws, float* mean, float* rstd, float* running_mean,
                float* running_var, float momentum, float eps, void* stream);

/* out = [relu]( bn(y) + [bn2(y2)] + [res] ) written as bf16 hi/lo planes (BasicBlock.forward resnets.py:214-230),
 * for ng groups of P pixels: group g uses mean/rstd[g][C] and gamma/beta + g*param_gstride (the perturbed BatchNorm
 * parameters of group g in pass 2; 0 = shared).  One streaming pass, no grid synchronisation.  The second normalised
 * branch and the identity residual exclude each other (a block has a downsample path or an identity shortcut). */
typedef struct {
  const float *y, *mean, *rstd, *gamma, *beta;
  const float *y2, *mean2, *rstd2, *gamma2, *beta2; /* optional second normalised branch (downsample), NULL if none */
  const void *res_hi, *res_lo;                      /* optional identity residual (bf16 planes) */
  int32_t relu;
  int64_t P; /* pixels per group */
  int32_t C;
  void *out_hi, *out_lo;
  int32_t ng; /* 0 -> 1 */
  int64_t param_gstride;
  int32_t reverse;
  uint8_t* mask_out; /* optional ReLU mask as BITS for fb_bn_bwd: byte e/8, bit e%8 of element e = [out > 0] */
} fb_bn_apply_args;
int fb_bn_apply(const fb_bn_apply_args* args, void* stream);

/* BatchNorm(+ReLU) backward for ng groups of P pixels.  dz = (dA + dA2) * [mask > 0]; the mask is either the bit plane
 * written by fb_bn_apply (mask_bits: 1/16 of the bytes of a bf16 plane -- the kernels are bandwidth bound) or the bf16
 * activation plane itself (mask_hi); both NULL: no ReLU (dA is taken as dz).  dA2 NULL: no second addend.  Two launches
 * without grid synchronisation: a column reduction whose last block per group finalises (fixed order), then a streaming
 * apply.  Writes dgamma / dbeta (+ g*grad_gstride), dY as bf16 (tensor-core operand) and optionally dz as fp32
 * (`dz_out`: the gradient of the shortcut branch that ends in the same activation).  dz_out may alias dA (in place:
 * every element is read and written by one thread).  With two addends dz_out is stored by the REDUCE launch and read
 * back by the apply launch instead of dA, dA2 and the mask (one fp32 read per element less; same bits).
 * gamma + g*param_gstride.
 * ws: >= 16 + ng*(2*C*fb_bn_bwd_chunks + 2*C) floats, the first 64 bytes ZERO on first use (self-resetting tickets). */
typedef struct {
  const float* dA;
  const float* dA2;
  const void* mask_hi;
  const float *y, *mean, *rstd, *gamma;
  int64_t P; /* pixels per group */
  int32_t C;
  float* ws;
  float *dgamma, *dbeta;
  void* dy_bf16;
  float* dz_out;
  int32_t ng; /* 0 -> 1 */
  int64_t param_gstride, grad_gstride;
  int32_t policy_groups; /* the reduction is cut into ~2*148/policy_groups chunks per group (0 -> ng) */
  int32_t reverse;       /* reduce back to front, apply front to back (0: the other way round) */
  const uint8_t* mask_bits;
} fb_bn_bwd_args;
int fb_bn_bwd(const fb_bn_bwd_args* args, void* stream);
int fb_bn_bwd_chunks(int64_t P, int C, int policy_groups);

/* Running-stat EMA of all BatchNorm layers in one launch, in the REFERENCE's order: for every group g (= microbatch,
 * loader order) and every pass p < n_passes (pass 1, FD pass 2, ...: training.py:159, modules.py:227-230):
 *     running = (1 - momentum) * running + momentum * batch[p][g]
 * batch = fb_conv_gemm_args.bn_batch of pass p: batch + p*pass_stride + g*2*C, [0] mean, [1] unbiased variance. */
typedef struct {
  float *running_mean, *running_var;
  const float* batch;
  int64_t pass_stride;
  int32_t C, c_start; /* c_start: first thread index of this entry (prefix sum of C) */
} fb_bn_ema_entry;
int fb_bn_ema_multi(const fb_bn_ema_entry* table_dev, int n_entries, int total_channels, int n_passes, int ng,
                    float momentum, void* stream);

/* AvgPool2d(2) on bf16 hi/lo planes (downsample 'C', resnets.py:147-152) and its backward (dX = up(dP)/4). */
int fb_avgpool2_fwd(const void* in_hi, const void* in_lo, int n, int h, int w, int c, void* out_hi, void* out_lo,
                    void* stream);
int fb_avgpool2_bwd(const float* dP, int n, int h, int w, int c, float* dX, int accumulate, void* stream);

/* Global average pool + Linear + LabelSmoothCrossEntropyLoss + accuracy + their backward
 * (resnets.py:106-107,183-186; modules.py:96-101; training.py:79-80) for ng groups of n images.  Group g uses
 * fc_w / fc_b + g*param_gstride, writes its mean loss to scal[loss_base + g], its correct count to
 * scal[correct_base + g], d(fc.weight) / d(fc.bias) + g*grad_gstride, and dA [ng*n][hw][c] fp32.
 * ws >= ng*n*(c+32). */
int fb_head_fwd_bwd(const void* a_hi, const void* a_lo, int n, int hw, int c, const float* fc_w, const float* fc_b,
                    const int64_t* labels, int classes, float smoothing, float* ws, float* scal, int loss_base,
                    int correct_base, float* d_fcw, float* d_fcb, float* dA, int ng, int64_t param_gstride,
                    int64_t grad_gstride, void* stream);

/* ---- flat-buffer (multi-tensor) kernels: GradRegularizer._forward_differences + running mean --------------------- */

/* Per group g < ng: s_g = sum (a*x[g] + b*y)^2 over n fp32 elements (y may be NULL; a = 1, b = 0: plain |x|^2),
 * deterministic two-stage reduction (training.py:162, modules.py:217-223).  scal[slot_base + g] = s_g;
 * norms_out[*cursor + g] = s_g if norms_out != NULL (grad_norms[k], training.py:162);
 * eps_mode 1: scal[eps_base + g] = eps / sqrt(bs*bs*s_g)   (modules.py:223 with v = bs*g)
 * eps_mode 2: scal[eps_base + g] = eps / sqrt(s_g)         (s_g already is |v|^2).   ws >= ng*1024 doubles. */
int fb_flat_sqnorm(const float* x, int64_t x_gstride, const float* y, float a, float b, int64_t n, int ng, double* ws,
                   float* scal, int slot_base, float* norms_out, const int32_t* cursor, int eps_mode, float bs,
                   float eps, int eps_base, void* stream);

/* theta_p[g][i] = theta[i] + (scale*eps_n[g]) * (bs*grad[g][i] + acc*pre[i]) on a TABLE of index ranges (the
 * parameters that are not conv weights: BatchNorm weight / bias, fc -- conv weights are perturbed inside
 * fb_weight_prep_multi).  ranges_dev: n_ranges x (offset, length, first thread index) int64 triples. */
int fb_perturb_ranges(const float* theta, const float* grad, int64_t grad_gstride, const float* pre,
                      const int64_t* ranges_dev, int n_ranges, int64_t total, float bs, float acc, float scale,
                      const float* scal, int eps_base, float* theta_p, int64_t theta_p_gstride, int ng, void* stream);

/* For g = 0..ng-1 in order (= loader order):
 *   g_reg = grad[g] + cf * (g_plus[g] - g_minus[g]) / eps_n[g]   (modules.py:232-240; g_minus == NULL: grad[g], i.e.
 *           forward differences; central differences modules.py:292-299 pass both); cf = scal[cf_slot]
 *   if write_g: grad[g] <- g_reg;   if avg: avg += (g_reg - avg) / (*cursor + g + 1)   (training.py:45-47,168) */
int fb_fd_combine(float* grad, const float* g_plus, const float* g_minus, int64_t gstride, float* avg, int64_t n,
                  int ng, const float* scal, int eps_base, int cf_slot, const int32_t* cursor, int write_g,
                  void* stream);
/* For g in order: [if clip > 0 and sqrt(scal[norm_base+g]) > clip: grad[g] *= clip/(norm+1e-6), scal[clipped_slot] += 1
 * (hyp.batch_clip, training/utils.py:4-19, training.py:166-168)]; avg += (grad[g] - avg)/(*cursor + g + 1)
 * (GradRegularizer._pass, modules.py:177-178 + training.py:45-47). */
int fb_mean_accumulate(float* grad, int64_t gstride, float* avg, int64_t n, int ng, const int32_t* cursor, float* scal,
                       int norm_base, float clip, int clipped_slot, void* stream);
/* End of a group launch: *cursor += cursor_step (the number of microbatches until this lane's next launch: ng for one
 * lane, lanes * G when launches alternate between lanes); totals[loss_slot] += scal[loss_base + g],
 * totals[correct_slot] += scal[correct_base + g] for g in order (training.py:172-173).  `totals` is shared by all lanes
 * (may be `scal` itself): called in loader order, the loss is summed microbatch by microbatch whatever ng and the lane
 * count are. */
int fb_group_finish(int32_t* cursor, int ng, int cursor_step, const float* scal, float* totals, int loss_slot,
                    int correct_slot, int loss_base, int correct_base, void* stream);
/* x *= alpha over n elements (rank-weighting before the all-reduce, training/utils.py:31-41) */
int fb_flat_scale(float* x, int64_t n, float alpha, void* stream);

/* Flat buffer between the reference's parameter order with OIHW conv weights (training/utils.py:34) and the native
 * gradient layout ([co][tap][ci] for 3x3 convs; identical otherwise).  to_native = 0: native -> OIHW.  table_dev:
 * n_entries x (offset, cout, cin, taps, first block) int64 quintuples of the 3x3 convs; src != dst; everything else is
 * copied. */
int fb_flat_relayout(const float* src, float* dst, int64_t n, const int64_t* table_dev, int n_entries,
                     int total_blocks, int to_native, void* stream);

/* The step right after the path as one sweep: clip by the global L2 norm (coef from
 * scal[norm_slot] = |g|^2, training.py:198-211; clip <= 0 disables), torch.optim.SGD update with weight decay,
 * momentum, dampening, Nesterov (optimizers.py:25-28, same operation order as torch/optim/sgd.py) and
 * scal[param_norm_slot] = sum theta_new^2 (training.py:92).  ws >= 1024 doubles. */
int fb_sgd_step(float* theta, float* grad, float* momentum_buf, int64_t n, float* scal, int norm_slot, float clip,
                float lr, float momentum, float dampening, float weight_decay, int nesterov, int first_step,
                int write_clipped_grad, double* ws, int param_norm_slot, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FULLBATCH_B200_H */

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
 * a*b_hi + a*b_lo (a=1,b=2).  A maps are indexed [phase * a_planes + plane], B maps [plane]. */
typedef struct {
  const void* host_a_maps; /* n_phases * a_planes x 128 B */
  const void* host_b_maps; /* b_planes x 128 B */
  int32_t n_phases, a_planes, b_planes;
  int32_t n_taps, cblocks;
  fb_tap taps[FB_MAX_TAPS];
  int32_t tile_w, tile_h, tile_n; /* product must be 128 */
  int32_t grid_h, grid_n;         /* rows and images of the output pixel grid */
  int32_t n_total, n_tile;        /* GEMM N (output channels) and the per-CTA N tile: 64, 128 or 256 */
  float* out;                     /* fp32; element (n,h,w,c) at out + n*out_sn + h*out_sh + w*out_sw + c */
  int64_t out_sn, out_sh, out_sw;
  int32_t accumulate; /* 0: overwrite, 1: out += result */
  /* optional BatchNorm statistics fused into the epilogue: stats_out[row][0|1][n_total] receives per-CTA column sums
   * and sums of squares of `out`, rows = fb_conv_stats_rows(...); feed them to fb_bn_fwd_fused.  NULL: off. */
  float* stats_out;
  /* optional tap groups (0 = one group of all n_taps): group g covers taps [tap0, tap0 + n_taps) and writes to
   * out + out_off; every group runs over the same pixel grid.  One launch then serves the four output phases of a
   * stride-2 dgrad (1 + 2 + 2 + 4 taps).  Not combinable with stats_out. */
  int32_t n_groups;
  struct {
    int32_t tap0, n_taps;
    int64_t out_off;
  } groups[4];
  /* optional BatchNorm-BACKWARD statistics of the stored tensor (a dgrad whose output is the upstream gradient dA of a
   * BatchNorm + ReLU): with bwd_y != NULL, stats_out receives per-CTA sums of dA*m and dA*m*xhat instead, where
   * m = (bwd_mask > 0) (bf16 post-activation plane, may be NULL) and xhat = (bwd_y - bwd_mean) * bwd_rstd; bwd_y and
   * bwd_mask have the layout of `out`.  Feed them to fb_bn_bwd_fused (fb_bn_bwd_args.stats): its first pass over
   * dA / y / mask and one grid barrier disappear.  Requires a single producer of dA (accumulate == 0, one tap group). */
  const float* bwd_y;
  const void* bwd_mask;
  const float* bwd_mean;
  const float* bwd_rstd;
} fb_conv_gemm_args;
/* number of partial rows written to stats_out for a problem of m_tiles x (n_total / n_tile) tiles */
int fb_conv_stats_rows(int m_tiles, int n_tiles);
int fb_conv_gemm(const fb_conv_gemm_args* args, void* stream);

/* 3x3 / stride 1 / pad 1 convolution (forward or dgrad) with haloed A boxes: per column shift dw one box of
 * (2*TH + 2) rows is fetched and the three row shifts are aligned views of it; a CTA tile is 256 pixels (two M=128
 * halves sharing every weight tile).  Requires W in {8,16,32,64,128}, H % (256/W) == 0.  A maps: [plane], box =
 * 64 x W x (256/W + 2) x 1; B maps: [plane], box = 64 x n_tile.  b_k0[dw+1][dh+1] = first weight column of the tap
 * that reads input pixel (h+dh, w+dw).  Same output conventions as fb_conv_gemm. */
typedef struct {
  const void* host_a_maps;
  const void* host_b_maps;
  int32_t a_planes, b_planes;
  int32_t b_k0[3][3];
  int32_t cblocks;
  int32_t w, h, n;
  int32_t n_total, n_tile; /* n_tile: 64 or 128 */
  float* out;
  int64_t out_sn, out_sh, out_sw;
  int32_t accumulate;
  float* stats_out; /* as in fb_conv_gemm_args; rows = fb_conv_stats_rows(m_tiles, n_total / n_tile) */
  /* Tile geometry (0 = default).  imgs == 1: a tile is halves * (128 / w) consecutive rows of one image
   * (m_tiles = n * h / (halves * 128 / w)), a_maps dims (C, W, H, N), box (64, w, halves*128/w + 2, 1).
   * imgs > 1 (imgs * w * h == 128, small maps): a 128-pixel half is `imgs` whole images whose rows are interleaved
   * in shared memory ([h][img][w]) so that the three row shifts stay 1024-byte aligned views;
   * m_tiles = n / (halves * imgs), a_maps dims (C, W, N, H), box (64, w, imgs, h + 2). */
  int32_t imgs;   /* default 1 */
  int32_t halves; /* 128-pixel halves per CTA tile sharing each weight tile: 1 or 2 (default 2) */
} fb_conv3x3_args;
int fb_conv3x3(const fb_conv3x3_args* args, void* stream);

typedef struct {
  int8_t phase, dh, dw;
  int8_t k_index; /* filter position this tap's gradient belongs to: partial column block k_index*cin (kh*k + kw) */
} fb_wgrad_tap;

/* Weight gradient: partial[split][co][tap*cin + ci] = sum_{pixels of split} dY[pixel, co] * X[pixel + tap shift, ci]
 * (autograd wgrad, training.py:82 / modules.py:230).  X maps are indexed [phase * planes + plane]. */
typedef struct {
  const void* host_dy_map; /* 1 x 128 B */
  const void* host_x_maps; /* n_x_maps x 128 B */
  int32_t n_x_maps, planes;
  int32_t n_taps, cblocks;
  fb_wgrad_tap taps[FB_MAX_WGRAD_TAPS];
  int32_t slots_per_cta; /* accumulators (tap, ci-block pairs) per CTA: slots_per_cta * planes <= 8 */
  int32_t cout, cin;     /* cin = 64*cblocks; row length of partial = n_taps*cin */
  int32_t tile_w, tile_h, tile_n;
  int32_t grid_h, grid_n; /* dY pixel grid */
  int32_t splits;         /* split-K over 128-pixel blocks */
  float* partial;         /* [splits][cout][n_taps*cin] fp32 */
  /* halo != 0 (3x3 / stride 1, tile_n == 1, tile_w * 128 B a multiple of 1024): taps are ordered in triples that share
   * dw (dh = -1, 0, +1), slots_per_cta == 3, and the X maps have boxes of tile_h + 2 rows: per pixel block and triple
   * ONE haloed X box is fetched and the three row shifts are aligned views of it (a third less operand traffic for a
   * kernel that is bound by it). */
  int32_t halo;
} fb_wgrad_args;
int fb_conv_wgrad(const fb_wgrad_args* args, void* stream);

/* Sum split-K partials in a fixed order and scatter to the reference layout (OIHW, training/utils.py:34 order).
 * mode 0: partial columns are tap*cin_stored + ci  -> g[co][ci][tap];  mode 1: columns already ci*taps + tap. */
int fb_wgrad_finalize(const float* partial, int splits, int cout, int cin, int taps, int cin_stored, int mode,
                      float* g_oihw, void* stream);

/* fp32 OIHW conv weight -> bf16 hi/lo GEMM operands: wf[co][tap][ci] (forward / wgrad order) and wd[ci][tap][co]
 * (dgrad); lo pointers may be NULL.  wd_* may be NULL (first layer needs no dgrad).  ld_f / ld_d: row strides. */
int fb_weight_prep(const float* w_oihw, int cout, int cin, int taps, void* wf_hi, void* wf_lo, int64_t ld_f,
                   void* wd_hi, void* wd_lo, int64_t ld_d, void* stream);

/* All conv weights of the network in one launch.  Table entries live in device memory; entry i owns the blocks
 * [block_start, block_start + n_blocks) with n_blocks = (cout/32)*(cin/32) (or any count >= 1 for the stem, cin < 32). */
typedef struct {
  int64_t w_offset; /* element offset of the OIHW weight inside theta */
  int32_t cout, cin, taps;
  int32_t block_start, n_blocks;
  int32_t pad;
  void *wf_hi, *wf_lo, *wd_hi, *wd_lo; /* lo / wd pointers may be NULL */
  int64_t ld_f, ld_d;
} fb_wprep_entry;
int fb_weight_prep_multi(const float* theta, const fb_wprep_entry* table_dev, int n_entries, int total_blocks,
                         void* stream);

/* ---- bandwidth-bound layer kernels ----------------------------------------------------------------------------- */

/* x [n,3,32,32] fp32 NCHW (optionally gathered through perm[first + i]) -> 3x3/pad-1 patches [n*1024][64] bf16 hi/lo,
 * column = ci*9 + kh*3 + kw (27 used).  `first` is read from *first_dev if first_dev != NULL (device-side microbatch
 * cursor).  labels_out[i] = labels[perm ? perm[first+i] : first+i]. */
int fb_stem_im2col(const float* x, const int64_t* labels, const int64_t* perm, const int32_t* first_dev, int64_t first,
                   int n, void* patches_hi, void* patches_lo, int64_t* labels_out, void* stream);

/* Same from a device-resident uint8 HWC dataset [N][32][32][3] with the reference's CIFAR training augmentation applied
 * on the fly (config/data/CIFAR10.yaml:22-26 via torchvision, data_preparation.py:173-200): RandomCrop(32, padding 4)
 * -> RandomHorizontalFlip -> ToTensor -> Normalize(mean, std).  aug (device, int8[.][4], may be NULL = no augmentation)
 * holds (dx, dy, flip, 0) per POSITION of the epoch order: crop offsets 0..8 in the zero-padded 40x40 image.
 * mean3 / std3 are host pointers. */
int fb_stem_im2col_u8aug(const uint8_t* x_hwc, const int64_t* labels, const int64_t* perm, const int32_t* first_dev,
                         int64_t first, int n, const int8_t* aug, const float* mean3, const float* std3,
                         void* patches_hi, void* patches_lo, int64_t* labels_out, void* stream);

/* Train-mode BatchNorm statistics over y[P][C] (resnets.py:71 / torch.nn.BatchNorm2d): mean, rstd = 1/sqrt(var+eps)
 * (biased var) and the running-stat EMA with unbiased variance.  ws: >= 2*C*1024 floats of scratch. */
int fb_bn_stats(const float* y, int64_t P, int C, float* ws, float* mean, float* rstd, float* running_mean,
                float* running_var, float momentum, float eps, void* stream);

/* out = [relu]( bn(y) + [bn2(y2)] + [res] ) written as bf16 hi/lo planes (BasicBlock.forward resnets.py:214-230). */
typedef struct {
  const float *y, *mean, *rstd, *gamma, *beta;
  const float *y2, *mean2, *rstd2, *gamma2, *beta2; /* optional second normalised branch (downsample), NULL if none */
  const void *res_hi, *res_lo;                      /* optional identity residual (bf16 planes) */
  int32_t relu;
  int64_t P;
  int32_t C;
  void *out_hi, *out_lo;
} fb_bn_apply_args;
int fb_bn_apply(const fb_bn_apply_args* args, void* stream);

/* BatchNorm(+ReLU) backward.  dz = (dA + dA2) * [mask_hi > 0] (mask_hi NULL: no ReLU; dA2 NULL: no second addend).  Writes dgamma/dbeta (fp32, C each),
 * dY as bf16 (tensor-core operand), and optionally dz as fp32 (`dz_out`, the identity-branch gradient; if
 * dz_accumulate != 0 it is added to dz_out instead of overwriting).  ws: >= 2*C*1024 floats of scratch. */
typedef struct {
  const float* dA;
  const float* dA2; /* optional second addend of the incoming gradient (shortcut branch), NULL if none */
  const void* mask_hi;
  const float *y, *mean, *rstd, *gamma;
  int64_t P;
  int32_t C;
  float* ws;
  float *dgamma, *dbeta;
  void* dy_bf16;
  float* dz_out;
  int32_t dz_accumulate;
  /* fb_bn_bwd_fused only: per-CTA partial sums [stats_rows][2][C] of dA*m and dA*m*xhat written by the dgrad that
   * produced dA (fb_conv_gemm_args.bwd_y); NULL: the kernel reduces them itself. */
  const float* stats;
  int32_t stats_rows;
} fb_bn_bwd_args;
int fb_bn_bwd(const fb_bn_bwd_args* args, void* stream);

/* Fused variants (one persistent launch with two grid-wide barriers: statistics -> finalize -> apply; the second pass
 * over the tensors is served by L2).  fb_bn_fwd_fused = fb_bn_stats (for y, and for y2 if given) + fb_bn_apply: it
 * WRITES args->mean / args->rstd (and mean2_out / rstd2_out).  fb_bn_bwd_fused = fb_bn_bwd without dz_accumulate.
 * ws: >= 2*C*1024 floats, the first 16 bytes ZERO on first use (self-resetting barrier counters). */
int fb_bn_fwd_fused(const fb_bn_apply_args* args, float* mean2_out, float* rstd2_out, float* running_mean,
                    float* running_var, float* running_mean2, float* running_var2, float momentum, float eps, float* ws,
                    const float* stats, int stats_rows, const float* stats2, int stats_rows2, void* stream);
/* stats / stats2 (optional): per-CTA column statistics [rows][2][C] written by the producing convolution
 * (fb_conv_gemm_args.stats_out); when given, the kernel skips its own statistics pass and one grid barrier. */
int fb_bn_bwd_fused(const fb_bn_bwd_args* args, void* stream);

/* AvgPool2d(2) on bf16 hi/lo planes (downsample 'C', resnets.py:147-152) and its backward (dX = up(dP)/4). */
int fb_avgpool2_fwd(const void* in_hi, const void* in_lo, int n, int h, int w, int c, void* out_hi, void* out_lo,
                    void* stream);
int fb_avgpool2_bwd(const float* dP, int n, int h, int w, int c, float* dX, int accumulate, void* stream);

/* Global average pool + Linear + LabelSmoothCrossEntropyLoss + accuracy + their backward
 * (resnets.py:106-107,183-186; modules.py:96-101; training.py:79-80).  Adds mean loss to scal[loss_slot] and the
 * correct count to scal[correct_slot]; writes d(fc.weight), d(fc.bias) and dA [n][hw][c] fp32.  ws >= n*(c+32). */
int fb_head_fwd_bwd(const void* a_hi, const void* a_lo, int n, int hw, int c, const float* fc_w, const float* fc_b,
                    const int64_t* labels, int classes, float smoothing, float* ws, float* scal, int loss_slot,
                    int correct_slot, float* d_fcw, float* d_fcb, float* dA, void* stream);

/* ---- flat-buffer (multi-tensor) kernels: GradRegularizer._forward_differences + running mean --------------------- */

/* scal[slot] = sum x^2 over n fp32 elements, deterministic two-stage reduction (training.py:162, modules.py:223);
 * if norms_out != NULL the value is also stored in norms_out[*cursor] (grad_norms[k], training.py:162).
 * ws >= 1024 doubles. */
int fb_flat_sqnorm(const float* x, int64_t n, double* ws, float* scal, int slot, float* norms_out,
                   const int32_t* cursor, void* stream);

/* eps_n = eps / sqrt(sum (bs*g)^2) from scal[sq_slot]; theta_p = theta + eps_n * (bs * g) (modules.py:217-226);
 * stores eps_n in scal[eps_slot]. */
int fb_fd_perturb(const float* theta, const float* g, int64_t n, float block_strength, float eps, float* scal,
                  int sq_slot, int eps_slot, float* theta_p, void* stream);

/* g_reg = g + cf * (g2 - g) / eps_n (modules.py:232-240), cf = scal[cf_slot] if cf_slot >= 0 (device-resident lr/4, so
 * that a captured CUDA graph survives learning-rate changes) else the `cf` argument; if avg != NULL:
 * avg += (g_reg - avg) / (count0 + *cursor + 1) (training.py:45-47,168); if write_g: g <- g_reg. */
int fb_fd_combine(float* g, const float* g2, float* avg, int64_t n, const float* scal, int eps_slot, float cf,
                  int cf_slot, const int32_t* cursor, int32_t count0, int write_g, void* stream);
/* avg += (g - avg) / (count0 + *cursor + 1) without regulariser (GradRegularizer._pass, modules.py:177-178) */
int fb_mean_accumulate(const float* g, float* avg, int64_t n, const int32_t* cursor, int32_t count0, void* stream);
/* *cursor += delta  (device-side microbatch cursor, so that a captured CUDA graph can be replayed per microbatch) */
int fb_cursor_add(int32_t* cursor, int32_t delta, void* stream);
/* x *= alpha over n elements (rank-weighting before the all-reduce, training/utils.py:31-41) */
int fb_flat_scale(float* x, int64_t n, float alpha, void* stream);

/* ---- generalised variants for the rest of hyp.grad_reg / hyp.batch_clip ---------------------- */
/* scal[slot] = sum (a*x + b*y)^2 (y may be NULL): |v|^2 for v = bs*g + acc*pre_grads (modules.py:217-223) */
int fb_flat_sqnorm_axpby(const float* x, const float* y, float a, float b, int64_t n, double* ws, float* scal, int slot,
                         void* stream);
/* eps_n = eps / sqrt(scal[vsq_slot]); theta_p = theta + scale*eps_n*(bs*g + acc*pre) (pre may be NULL).  scale = 1:
 * forward differences with acc_strength (modules.py:217-226); scale = +-0.5: central differences (modules.py:279-286) */
int fb_fd_perturb_ex(const float* theta, const float* g, const float* pre, int64_t n, float block_strength,
                     float acc_strength, float eps, float scale, float* scal, int vsq_slot, int eps_slot, float* theta_p,
                     void* stream);
/* g_reg = g + cf*(g_plus - g_minus)/eps_n (central differences, modules.py:292-299); avg / cursor / write_g as in
 * fb_fd_combine */
int fb_fd_combine_ex(float* g, const float* g_plus, const float* g_minus, float* avg, int64_t n, const float* scal,
                     int eps_slot, float cf, int cf_slot, const int32_t* cursor, int32_t count0, int write_g,
                     void* stream);
/* hyp.batch_clip (training/utils.py:4-19, training.py:166-168): g *= clip/(|g|+1e-6) if |g| = sqrt(scal[norm_slot]) >
 * clip, scal[clipped_slot] += 1 in that case, then avg += (g - avg)/(count0 + *cursor + 1) */
int fb_mean_accumulate_clip(float* g, float* avg, int64_t n, const int32_t* cursor, int32_t count0, float* scal,
                            int norm_slot, float clip, int clipped_slot, void* stream);

/* The step right after the path as one sweep: clip by the global L2 norm (coef from
 * scal[norm_slot] = |g|^2, training.py:198-211; clip <= 0 disables), torch.optim.SGD update with weight decay,
 * momentum, dampening, Nesterov (optimizers.py:25-28, same operation order as torch/optim/sgd.py) and
 * scal[param_norm_slot] = sum theta_new^2 (training.py:92).  ws >= 1024 doubles. */
int fb_sgd_step(float* theta, float* grad, float* momentum_buf, int64_t n, float* scal, int norm_slot, float clip,
                float lr, float momentum, float dampening, float weight_decay, int nesterov, int first_step,
                int write_clipped_grad, double* ws, int param_norm_slot, void* stream);

/* Development aid: with FB_KERNEL_DEBUG=1 fb_conv3x3 accumulates per-role wait cycles of CTA 0 in 32 device counters;
 * this call synchronises the device, copies them to host32[32] and optionally clears them. */
int fb_debug_counters(long long* host32, int clear);

#ifdef __cplusplus
}
#endif
#endif /* FULLBATCH_B200_H */

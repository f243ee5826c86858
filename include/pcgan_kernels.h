/*
 * pcgan_kernels.h — C ABI of libpcgan_kernels.so (B200 / sm_100a).
 *
 * This is the whole drop-in boundary of the wsgan_emb hot path: plain pointers,
 * sizes and POD structs; no torch types.  The Python host side
 * (pcgan_b200/_lib.py) binds it with ctypes.  Every entry point names the piece
 * of the reference it replaces (paths relative to phymhan/pc-gan).
 *
 * Conventions
 *   - every function returns 0 on success or a negative pcgan_status; the text
 *     of the last failure on the calling thread is pcgan_last_error()
 *   - the caller owns all device memory (inputs, outputs, statistics); the
 *     library allocates no device memory and keeps no pointer past return,
 *     except TMA descriptors cached inside an igemm plan (re-encoded when the
 *     pointers change)
 *   - every launch goes to the stream handed in; no call synchronises; all
 *     calls are CUDA-graph capturable
 *   - the current CUDA device must be the one that owns the pointers
 *   - activations are NHWC bf16 in *physically padded* buffers
 *     [N][H+2p][W+2p][C]; the halo (zeros or a reflection of the interior) is
 *     written by the kernel that produces the buffer, so padding never costs a
 *     pass of its own and every convolution is a "valid" one
 */
#ifndef PCGAN_KERNELS_H_
#define PCGAN_KERNELS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCGAN_ABI_VERSION 21
#define PCGAN_MAX_TAPS 64

typedef void* pcgan_stream_t; /* a cudaStream_t */

typedef enum {
  PCGAN_OK = 0,
  PCGAN_ERR_INVALID = -1,     /* bad argument / inconsistent descriptor      */
  PCGAN_ERR_UNSUPPORTED = -2, /* shape outside what the kernels implement    */
  PCGAN_ERR_CUDA = -3         /* CUDA runtime / driver error (text in last_error) */
} pcgan_status;

typedef enum { PCGAN_ACT_NONE = 0, PCGAN_ACT_RELU = 1, PCGAN_ACT_LRELU = 2, PCGAN_ACT_TANH = 3, PCGAN_ACT_SIGMOID = 4 } pcgan_act;
typedef enum { PCGAN_DT_BF16 = 0, PCGAN_DT_F32 = 1 } pcgan_dtype;
typedef enum { PCGAN_HALO_ZERO = 0, PCGAN_HALO_REFLECT = 1 } pcgan_halo;
typedef enum { PCGAN_IGEMM_KMAJOR = 0, PCGAN_IGEMM_WGRAD = 1 } pcgan_igemm_kind;
typedef enum { PCGAN_STATS_NONE = 0, PCGAN_STATS_ON = 1 } pcgan_stats_mode;
typedef enum { PCGAN_LOSS_BCE = 0, PCGAN_LOSS_MSE = 1, PCGAN_LOSS_L1 = 2, PCGAN_LOSS_ELO_NLL = 3, PCGAN_LOSS_ELO_NLL_SCORE = 4 } pcgan_loss_kind;

int pcgan_abi_version(void);
const char* pcgan_last_error(void);
/* sizeof() of an ABI struct by name (-1 if unknown): lets a binding check its mirror of the layout. */
int64_t pcgan_sizeof(const char* struct_name);

/* ------------------------------------------------------------------------- *
 * Implicit-GEMM convolution engine (tcgen05 + TMEM + TMA)
 *
 * Replaces every nn.Conv2d / nn.ConvTranspose2d forward, data-gradient and
 * weight-gradient on the path: models/networks.py:578-605 (ResnetGenerator),
 * :621-648 (ResnetBlock), :747-775 (NLayerDiscriminator), :1014-1027
 * (SiameseFeature head), models/resnet.py:20-28,134 (ResNet-18 trunk), which
 * the reference dispatches to cuDNN / oneDNN through ATen.
 *
 * One kernel family, described by data: the host planner (pcgan_b200/plan.py)
 * turns a convolution into
 *   A : a <=5-D TMA view of a padded NHWC activation buffer (dim 0 = channels)
 *   B : a <=5-D TMA view of a packed bf16 weight matrix [rows][K]
 *   a tap table: per filter tap, the coordinate offsets of the A box
 *   an output map: how a row of the 128-row tile becomes an output address
 * KMAJOR  D[M=pixels][N=Cout] = sum_taps A_tap[M][Cin] * B_tap[N][Cin]^T
 *         (forward and data-gradient; transposed convs and strided dgrads run
 *          one launch per sub-pixel phase)
 * WGRAD   D[M=Cout][N=Cin]    = sum_pixels dY[pix][M]^T * X_tap[pix][N]
 *         (both operands MN-major in shared memory; split-K with fp32 atomics)
 * ------------------------------------------------------------------------- */
typedef struct {
  uint64_t dims[5];    /* extent per dimension, dim 0 innermost (contiguous) */
  uint64_t strides[5]; /* bytes; strides[0] is ignored (elements are packed) */
  uint32_t box[5];     /* TMA box; box[0] must be 64 (one 128-byte swizzle row) */
} pcgan_tmap;

typedef struct {
  int32_t lo, hi;  /* component valid iff lo <= c < hi                   */
  int64_t stride;  /* output elements per unit of (c - lo)              */
} pcgan_comp;

typedef struct {
  int32_t kind;    /* pcgan_igemm_kind */
  int32_t block_n; /* UMMA N: 16..256, multiple of 16 (WGRAD: multiple of 64) */
  pcgan_tmap a, b;

  /* Tiles.  KMAJOR: one tile = one A box = 128 (or fewer) output pixels.
   * WGRAD: one "tile" = one K block of 64 pixels (A = dY box, B = X box).
   * The tile index is mixed-radix over t_count (digit 0 fastest); the box
   * coordinate of outer dim d (d = 0..3 <-> tensor dims 1..4) is
   *   A: a_base[d] + sum_j t_j*a_step[j][d]   (+ tap_off[tap][d] for KMAJOR)
   *   B: b_base[d] + sum_j t_j*b_step[j][d] + tap_off[tap][d]   (WGRAD only;
   *      KMAJOR B is the weight matrix: coords (k, n_tile*block_n, 0,0,0)). */
  int32_t t_count[4];
  int32_t a_base[4], a_step[4][4];
  int32_t b_base[4], b_step[4][4];
  int32_t n_tiles; /* tiles along N (KMAJOR: Cout/block_n; WGRAD: Cin blocks) */
  int32_t m_tiles; /* WGRAD: tiles of 128 along Cout; KMAJOR: ignored      */
  int32_t ksplit;  /* WGRAD: split of the pixel loop across CTAs; KMAJOR: 1 */

  /* K loop = num_taps x cchunks chunks of 64 elements.
   * KMAJOR: A dim-0 coordinate = tap_c0[tap] + 64*cc, B k-coordinate =
   * tap_bk[tap] + 64*cc.  WGRAD: B outer coords get tap_off, B dim-0
   * coordinate = tap_c0[tap] + nt*block_n, output column base = tap_bk[tap]. */
  int32_t num_taps, cchunks;
  int32_t tap_off[PCGAN_MAX_TAPS][4];
  int32_t tap_c0[PCGAN_MAX_TAPS];
  int32_t tap_bk[PCGAN_MAX_TAPS];

  /* Epilogue.  KMAJOR: row r of a tile has box-local index (i1..i4) over
   * a.box[1..4]; g_d = e_base[d] + sum_j t_j*e_step[j][d] + i_d is split into
   * c0 = g / p1, c1 = (g % p1) / p2, c2 = g % p2 (p1 == 0: c0 = 0 and the
   * remainder is g; p2 == 0: c1 = remainder, c2 = 0); the row is stored iff
   * every component is inside [lo,hi) and lands at
   * out + sum (c - lo)*stride + channel*out_cstride. */
  int32_t e_base[4], e_step[4][4], e_p1[4], e_p2[4];
  pcgan_comp e_comp[4][3];
  int32_t out_dtype;   /* pcgan_dtype (WGRAD: always f32, accumulated with atomics) */
  int32_t act;         /* pcgan_act applied after bias                     */
  float act_slope;
  int32_t n_valid;     /* real output channels (columns >= n_valid are dropped)  */
  int64_t out_cstride; /* elements between channels: 1 = NHWC, H*W = NCHW  */
  int32_t stats_mode;  /* per-channel sum / sum-of-squares of (acc + bias) of the
                          stored rows, atomically added to stats[group][n_valid][2] */
  int32_t stats_dim, stats_comp; /* group index = that component of row 0 of the tile; stats_dim < 0: group 0 */
  /* WGRAD only */
  int32_t m_valid;     /* Cout                                             */
  int32_t wg_ncols;    /* valid columns per tap (Cin or packed row width)  */
  int64_t ldo;         /* output row pitch in elements                     */
  /* 1: launch as thread-block clusters of two CTAs that take two M tiles of the same N tile (tap, K split) and issue
   * ONE tcgen05.mma.cta_group::2 of M = 256 per K step: each CTA fetches its own A tile and half of the B tile, so the
   * shared-memory operand traffic of a 128x256 tile drops from 48 to 32 KB per K chunk per SM (shared-memory bandwidth,
   * not the tensor pipe, bounds these tiles), and drains its own 128 accumulator rows.
   * KMAJOR: any shape (an odd M-tile count recomputes and drops one tile); WGRAD: m_tiles and block_n/64 must be even. */
  int32_t pair;
  /* KMAJOR shift-sum epilogue (shift_taps > 0): for convolutions with very few output channels (generator head 64 -> 3,
   * data gradients towards 3/4-channel images) the horizontal filter taps move from K to N: the accumulator's columns
   * are shift_taps groups of shift_cpad channels and the stored value of tile row i, channel c is
   *     sum_j acc[i + j][j*shift_cpad + c]        (rows i >= a_rows - (shift_taps - 1) store nothing; tiles overlap),
   * so each activation row is fetched once per filter ROW instead of once per filter TAP.  n_valid <= shift_cpad <= 8,
   * shift_taps*shift_cpad <= 32 = block_n, no statistics. */
  int32_t shift_taps, shift_cpad;
  /* KMAJOR windowed A (a_window == 8): for 8-channel inputs (image stems, gradients of 3-channel heads) whose filter row
   * of up to 8 taps x 8 channels is one K chunk.  A is the plain tensor (a.dims[0] == a.box[0] == 8, 16-byte pixels);
   * a.box[1] = rows + 7 consecutive pixels of one image row (or of the flattened grid) are fetched ONCE per filter row and
   * row m of the tile reads the 64 elements starting at pixel m, through an overlapping shared-memory descriptor, instead
   * of fetching every pixel 8 times.  Packed weights as for the overlapping-stride form: [cout][tap][8 pixels x 8]. */
  int32_t a_window;
  /* WGRAD filter rows in N (wg_box_dim = 1..4, 0 = off): the nb = block_n/64 boxes of N tile nt are not 64-channel slices of
   * one pixel box but the same channels (tap_c0) at coordinate nt*nb + j along B tensor dim wg_box_dim (a filter-row
   * dimension with the image-row stride, added to the view by the planner; boxes beyond its extent read zeros).  With
   * num_taps = 1 the M operand is fetched once per N tile instead of once per filter row. */
  int32_t wg_box_dim;
  /* 0: bf16 operands (tcgen05.mma.kind::f16), K chunks of 64 elements.  1: fp32 operands multiplied as TF32
   * (tcgen05.mma.kind::tf32, 10-bit mantissa products, fp32 accumulate): both tensor maps are fp32, a K chunk is 32
   * elements (still 128 bytes), WGRAD boxes are 64 pixels x 32 channels, box[0] of both maps is 32; no pairing, no
   * windowed A.  Everything else (tap tables, tile maps, epilogue) is unchanged. */
  int32_t tf32;
  /* statistics group = (sample index selected by stats_dim / stats_comp) / stats_div; 0 or 1: the sample itself.
   * stats_div = samples per group batches several BatchNorm batches (passes of a network) into one launch. */
  int32_t stats_div;
} pcgan_igemm_desc;

typedef struct pcgan_igemm_plan pcgan_igemm_plan;

/* Validates the descriptor (host only, no GPU needed) and copies it. */
int pcgan_igemm_plan_create(const pcgan_igemm_desc* desc, pcgan_igemm_plan** plan);
void pcgan_igemm_plan_destroy(pcgan_igemm_plan* plan);
/* Host-only self-test of the magic-number divisions the kernel uses for its per-tile coordinate arithmetic (the same
 * formula evaluated on the host against x / d): returns the number of mismatches over `iters` random divisors. */
int64_t pcgan_selftest_fastdiv(uint32_t seed, int32_t iters);
/* a, b: operand base pointers (16-byte aligned); out: output base; bias: f32
 * [n_valid] or NULL; stats: f32 [groups][n_valid][2] or NULL. */
int pcgan_igemm_run(pcgan_igemm_plan* plan, const void* a, const void* b, void* out,
                    const float* bias, float* stats, pcgan_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Layout / packing kernels
 * ------------------------------------------------------------------------- */
/* dst[i] = idx[i] >= 0 ? bf16(src[idx[i]]) : 0.  Packs OIHW fp32 master weights
 * (nn.Conv2d.weight, IOHW for ConvTranspose2d: networks.py:595) into the
 * [rows][K] bf16 operand of a plan. */
int pcgan_gather_cast_bf16(const float* src, const int32_t* idx, void* dst_bf16, int64_t n, pcgan_stream_t stream);
/* dst[idx[i]] (+)= src[i] for idx[i] >= 0: packed fp32 weight gradient -> .grad in OIHW. */
/* packed fp32 operand of a TF32 plan: dst[i] = idx[i] >= 0 ? round_to_nearest_tf32(src[idx[i]]) : 0 */
int pcgan_gather_tf32(const float* src, const int32_t* idx, float* dst, int64_t n, pcgan_stream_t stream);
int pcgan_scatter_f32(const float* src, const int32_t* idx, float* dst, int64_t n, int32_t accumulate, pcgan_stream_t stream);

/* The same two operations for many tensors in ONE launch: `items` is a device array of `count` descriptors (a network's
 * whole set of packed operands after an optimizer step / of packed weight gradients after a backward pass);
 * max_n = the largest items[i].n (sizes the grid). */
typedef struct {
  const float* src;   /* gather: fp32 master weight;  scatter: packed fp32 gradient            */
  const int32_t* idx; /* n entries: index into the master-layout tensor, or -1                 */
  void* dst;          /* gather: packed bf16 operand; scatter: fp32 master-layout .grad        */
  int64_t n;
} pcgan_batch_item;
int pcgan_gather_cast_bf16_batched(const pcgan_batch_item* items, int32_t count, int64_t max_n, pcgan_stream_t stream);
int pcgan_scatter_f32_batched(const pcgan_batch_item* items, int32_t count, int64_t max_n, int32_t accumulate, pcgan_stream_t stream);

/* NCHW fp32 image [N][Cs][H][W] (+ optional per-sample scalar z[N] appended as
 * channel Cs: networks.py:610-611, :780-782 torch.cat((input, z_img), 1)),
 * optionally bilinearly resized to Ho x Wo with align_corners=True
 * (util/util.py:111-117) and multiplied by act'(t) of a second NCHW tensor t holding the
 * activation's OUTPUT (mul_kind PCGAN_ACT_TANH: 1 - t*t, PCGAN_ACT_SIGMOID: t*(1-t); the
 * backward of nn.Tanh / nn.Sigmoid), -> padded NHWC bf16 [N][Ho+2p][Wo+2p][Cd].
 * Under reflect halo the z plane stays constant; under zero halo it is 0 in the
 * halo, exactly as ReflectionPad2d / Conv2d(padding=1) see it. */
typedef struct {
  const float* src; const float* z; const float* mul_out; int32_t mul_kind;
  void* dst;
  int32_t n, cs, h, w;     /* source geometry                                  */
  int32_t ho, wo;          /* destination interior (== h,w when not resizing)  */
  int32_t cd, pad, halo;   /* destination channels (multiple of 8), pad, mode  */
  int64_t dst_n_stride;    /* elements between samples (0: packed)             */
} pcgan_pack_args;
int pcgan_pack_nchw(const pcgan_pack_args* a, pcgan_stream_t stream);

/* Bilinear resize with align_corners=True of NCHW fp32 planes [planes][h][w] -> [planes][ho][wo]
 * (util.upsample2d, util/util.py:111-117) and its adjoint (dst[planes][h][w] = sum of the
 * resized-grid gradients weighted by their interpolation coefficients). */
int pcgan_resize_nchw_fwd(const float* src, float* dst, int64_t planes, int32_t h, int32_t w, int32_t ho, int32_t wo, pcgan_stream_t stream);
int pcgan_resize_nchw_bwd(const float* gdst, float* gsrc, int64_t planes, int32_t h, int32_t w, int32_t ho, int32_t wo, pcgan_stream_t stream);

/* Adjoint of the bilinear resize above, reading an NHWC bf16 gradient
 * [N][Hs+2p][Ws+2p][C] (interior only) and writing / accumulating NCHW fp32
 * [N][Cd][H][W] (first Cd channels). */
typedef struct {
  const void* g; float* dst;
  int32_t n, c, hs, ws, pad; /* gradient geometry (resized size)        */
  int32_t cd, h, w;          /* destination                              */
  int32_t accumulate;
  float scale;
} pcgan_unpack_args;
int pcgan_unpack_resize_bwd(const pcgan_unpack_args* a, pcgan_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Normalisation (+ activation, residual) — nn.InstanceNorm2d(affine=False,
 * track_running_stats=True) and nn.BatchNorm2d in training mode
 * (networks.py:22-34; used at :580-581,:588-589,:601-602,:633,:646,:759,:768,
 * resnet.py:47-65,136; networks.py:1021).
 * ------------------------------------------------------------------------- */
/* stats[g][c][2] = (sum, sum of squares) over `count` elements ->
 * mean/rstd [g][c], scale = gamma*rstd, shift = beta - mean*scale, and the
 * running-stat EMA (momentum, unbiased variance; for instance norm the batch
 * mean of the per-instance statistics).  gamma/beta/running_* may be NULL. */
typedef struct {
  const float* stats; int32_t groups, c; float count, eps, momentum;
  const float* gamma; const float* beta;
  float* mean; float* rstd; float* scale; float* shift;
  float* running_mean; float* running_var;
  /* Channel dropout in front of a BatchNorm (nn.Dropout2d before bn, resnet.py:58-65): the convolution emits
   * per-sample statistics stats[in_groups = N][c][2]; with drop_mask[N][c] (0 or 1/(1-p)) the batch statistics of the
   * masked tensor are sum_n m*S1 and sum_n m*m*S2.  in_groups == 0: stats has `groups` groups (no combination). */
  const float* drop_mask; int32_t in_groups;
} pcgan_norm_finalize_args;
int pcgan_norm_finalize(const pcgan_norm_finalize_args* a, pcgan_stream_t stream);

/* y = act(scale*x + shift [+ res_scale*res + res_shift]) written into the
 * interior AND halo of a padded NHWC bf16 buffer.  x: NHWC bf16 with its own
 * pad (interior read).  scale/shift are [groups][C] with groups = N (instance)
 * or 1 (batch), or NULL (identity).  Optional channel-dropout mask[N][C] (f32, already scaled by
 * 1/(1-p)) multiplies x first (nn.Dropout2d, resnet.py:58-65). */
typedef struct {
  const void* x; int32_t x_pad;
  const void* res; int32_t res_pad;
  void* y; int32_t y_pad; int32_t y_halo;
  int32_t n, h, w, c;
  const float* scale; const float* shift; int32_t groups;
  const float* res_scale; const float* res_shift; int32_t res_groups;
  const float* drop_mask;
  int32_t act; float act_slope;
  const float* post_mask; /* optional [N][C]: multiplies the activation's output (nn.Dropout2d between BatchNorm and
                             LeakyReLU, networks.py:1021-1023: lrelu(m*v) == m*lrelu(v) for m >= 0) */
  /* Fused finalize (stats != NULL; scale / shift above are then ignored): every block derives the scale / shift of its
   * channels from the raw statistics [groups][c][2] exactly as pcgan_norm_finalize does, and the first block of each
   * group stores mean / rstd / scale / shift [groups][c] for the backward pass.  The running statistics are updated by
   * pcgan_norm_running_batched. */
  const float* stats; float count, eps;
  const float* gamma; const float* beta;
  float* mean_out; float* rstd_out; float* scale_out; float* shift_out;
  /* The output may be a channel slice of a wider buffer (the skip concatenations of UnetGenerator,
   * networks.py:722-733): y holds y_c >= c channels per pixel and this launch writes channels [y_c0, y_c0 + c).
   * y_c == 0: a buffer of exactly c channels.  Zero halo only. */
  int32_t y_c, y_c0;
} pcgan_norm_apply_args;
int pcgan_norm_apply(const pcgan_norm_apply_args* a, pcgan_stream_t stream);

/* Running-statistics EMA (and num_batches_tracked += 1) of every normalisation layer of a network pass in ONE launch:
 * for each item, running_mean = (1-momentum)*running_mean + momentum*mean over groups of the group means, running_var
 * likewise with the unbiased variance — what pcgan_norm_finalize does per layer (nn.InstanceNorm2d with
 * track_running_stats=True / nn.BatchNorm2d in training mode).  `items` is a device array. */
typedef struct {
  const float* stats;      /* [groups][c][2] */
  float* running_mean; float* running_var;   /* [c], may be NULL */
  int64_t* num_batches_tracked;              /* scalar, may be NULL */
  int32_t groups, c;
  float count, momentum;
  int32_t sequential;   /* 0: the groups are the samples of one pass (InstanceNorm: batch mean of the instance statistics);
                           1: the groups are successive BatchNorm batches, one momentum step each, in order */
  int32_t reserved_;
} pcgan_running_item;
int pcgan_norm_running_batched(const pcgan_running_item* items, int32_t count, int32_t max_c, pcgan_stream_t stream);

/* out[N][H][W][C] (pad out_pad, zero halo kept) = fold(gpad) [+ add]: folds the gradient of a padded
 * buffer (reflect: halo gradients are added onto their mirror pixels; zero: halo
 * dropped) back onto the interior — the adjoint of the halo write above
 * (nn.ReflectionPad2d backward). */
typedef struct {
  const void* gpad; int32_t g_pad; int32_t halo;
  const void* add; int32_t add_pad;
  void* out; int32_t out_pad;
  int32_t n, h, w, c;
} pcgan_fold_args;
int pcgan_halo_fold(const pcgan_fold_args* a, pcgan_stream_t stream);
/* The same reflect fold IN PLACE on the padded gradient gpad: the interior pixels within g_pad of a border receive
 * the halo values that ReflectionPad2d copied from them (add / out are ignored; the halo itself is left as it is).
 * Afterwards the interior of gpad is the folded gradient: consumers read it with a dropped halo (norm backward
 * dy_fold = 1, pcgan_halo_fold with PCGAN_HALO_ZERO).  Touches 2*g_pad rows and columns only. */
int pcgan_halo_accumulate(const pcgan_fold_args* a, pcgan_stream_t stream);

/* Backward of y = act(scale*x + shift (+res...)): with g = dy * act'(y),
 * pass 1 accumulates sums[g][c][2] = (sum g, sum g*xhat); pass 2 writes
 * dx = scale*(g - sum_g/count - xhat*sum_gxhat/count) (count == 0: plain
 * dx = scale*g, i.e. no normalisation) into a zero-haloed padded buffer and, if
 * asked, g itself (gradient of the residual branch). y is recomputed from x. */
typedef struct {
  const void* dy; int32_t dy_pad;       /* NHWC bf16 gradient wrt y (interior) */
  const void* x; int32_t x_pad;         /* saved pre-norm activations          */
  const void* res; int32_t res_pad;     /* residual input (to recompute y)     */
  const float* mean; const float* rstd; const float* scale; const float* shift; int32_t groups;
  const float* res_scale; const float* res_shift; int32_t res_groups;
  const float* drop_mask;
  int32_t act; float act_slope;
  int32_t n, h, w, c;
  float count;                          /* elements per statistic; 0 = no norm  */
  float* sums;                          /* [groups][c][2], zeroed by the caller */
  void* dx; int32_t dx_pad;             /* pass 2 outputs                       */
  void* dres; int32_t dres_pad;         /* optional: g (masked dy)              */
  int32_t dy_fold;                      /* 0: dy is read at its interior; 2: dy is the gradient of a reflect-padded
                                           buffer (pad dy_pad) whose halo is folded onto the mirror pixels while
                                           reading (nn.ReflectionPad2d backward fused into this pass); 1: halo dropped */
  const float* post_mask;               /* optional [N][C]: the forward multiplied the activation output by it */
  int32_t affine;                       /* 0: scale == rstd and shift == -mean*rstd exactly (no gamma / beta), so
                                           xhat is the pre-activation itself (InstanceNorm2d(affine=False)); 1: general */
} pcgan_norm_bwd_args;
int pcgan_norm_bwd_reduce(const pcgan_norm_bwd_args* a, pcgan_stream_t stream);
int pcgan_norm_bwd_apply(const pcgan_norm_bwd_args* a, pcgan_stream_t stream);
/* Both passes in ONE launch for the lean InstanceNorm path (affine == 0, no masks, dx only): a thread-block cluster
 * per sample keeps dy and x resident in shared memory and reduces the per-channel sums through distributed shared
 * memory (3 tensor passes over HBM instead of 5).  `sums` is not written.  _supported() tells (1 / 0) whether the
 * arguments qualify (image rows split evenly over the cluster and fit its shared memory); otherwise use the two passes. */
int pcgan_norm_bwd_fused_supported(const pcgan_norm_bwd_args* a);
int pcgan_norm_bwd_fused(const pcgan_norm_bwd_args* a, pcgan_stream_t stream);
int pcgan_norm_bwd_fused_active_clusters(void);   /* diagnostic: resident clusters (cudaOccupancyMaxActiveClusters), -1 before the first launch */

/* 3x3 stride-2 max pooling on padded NHWC bf16, pooling padding pool_pad = 1 (nn.MaxPool2d(3, 2, 1), resnet.py:137) or
 * 0 (nn.MaxPool2d(3, 2), the AlexNet feature extractor, networks.py:1224-1236); idx keeps the window position (0..8)
 * of the maximum for the backward pass. */
typedef struct {
  const void* x; int32_t x_pad; void* y; int32_t y_pad; uint8_t* idx;
  int32_t n, h, w, c; /* input geometry; output is floor((h + 2*pool_pad - 3) / 2) + 1 squared */
  int32_t pool_pad;
} pcgan_maxpool_args;
int pcgan_maxpool3x3s2_fwd(const pcgan_maxpool_args* a, pcgan_stream_t stream);
/* dx (input geometry, pad x_pad, zero halo) from dy (output geometry, pad y_pad). */
int pcgan_maxpool3x3s2_bwd(const void* dy, int32_t dy_pad, const uint8_t* idx, void* dx, int32_t dx_pad,
                           int32_t n, int32_t h, int32_t w, int32_t c, int32_t pool_pad, pcgan_stream_t stream);

/* Backward of an activation fused into a convolution epilogue, from the sign of the stored output y:
 * dx = y > 0 ? dy : slope * dy (nn.ReLU: slope 0; networks.py:1223-1235).  Any channel count that is a multiple of 8. */
int pcgan_act_bwd(const void* dy, int32_t dy_pad, const void* y, int32_t y_pad, void* dx, int32_t dx_pad,
                  int32_t n, int32_t h, int32_t w, int32_t c, float slope, pcgan_stream_t stream);
/* to_f32 = 1: interior of a padded NHWC bf16 buffer -> contiguous NHWC fp32 (the feature map a torch loss consumes);
 * to_f32 = 0: the reverse (its gradient). */
int pcgan_nhwc_cast(const void* src, void* dst, int32_t pad, int32_t n, int32_t h, int32_t w, int32_t c, int32_t to_f32,
                    pcgan_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Input pipeline (data/base_dataset.py:24-64, data/wsgan_emb_dataset.py:36-49):
 * transforms.Resize([load, load], BICUBIC) -> RandomCrop(fine) -> RandomHorizontalFlip
 * -> ToTensor -> Normalize(0.5, 0.5) of n decoded RGB images (uint8 HWC in device
 * memory, any sizes) into dst fp32 [n][3][fine][fine] in one launch.  The resize is
 * PIL's: antialiased bicubic (a = -0.5, support scaled by the down-scaling factor), the
 * horizontal pass rounded to 8 bits before the vertical pass.  The random draws (crop
 * origin in the resized image, flip) are the caller's.
 * ------------------------------------------------------------------------- */
typedef struct {
  const uint8_t* src; int32_t h, w;     /* decoded image, row-major RGB */
  int32_t crop_y, crop_x;               /* crop origin in the load x load image */
  int32_t flip;
  int32_t reserved_;
} pcgan_image_item;
typedef struct {
  const pcgan_image_item* items;        /* device array [n] */
  float* dst; int32_t n, load, fine;
} pcgan_augment_args;
int pcgan_augment(const pcgan_augment_args* a, pcgan_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Losses: vectorised, coalesced reductions (GANLoss networks.py:386-420 ->
 * nn.BCELoss with log clamped at -100 / nn.MSELoss; nn.L1Loss and nn.MSELoss in
 * wsgan_emb_model.py:149,402,430; BinaryNLLLoss networks.py:473-482).
 *   BCE     mean(-(t*max(log p,-100) + (1-t)*max(log(1-p),-100)))
 *   MSE     mean((p-t)^2)          L1  mean(|p-t|)
 *   ELO_NLL mean(-(t*log(p+1e-20) + (1-t)*log(1-p+1e-20)))
 *   ELO_NLL_SCORE the same with p = sigmoid(input): the input is the rating difference (siamese.py:674-676), the
 *           gradient is taken with respect to it
 * target: either a full tensor t[n] (per_sample = 0) or one value per sample
 * t[n / per_sample].  loss is atomically accumulated (*loss += weight*mean).
 * grad (optional) = weight * dmean/dp, same shape as p. */
typedef struct {
  int32_t kind; const float* p; const float* target; int64_t n; int64_t per_sample;
  float weight; const float* weight_dev; /* optional device scalar multiplied into weight */
  float* loss; float* grad;
} pcgan_loss_args;
int pcgan_loss(const pcgan_loss_args* a, pcgan_stream_t stream);

/* Fused Adam step over one flat fp32 parameter (torch.optim.Adam semantics,
 * wsgan_emb_model.py:153-154): step count and lr are read from device memory
 * so the launch is graph-capturable. */
int pcgan_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, float beta1, float beta2,
               float eps, const float* step, pcgan_stream_t stream);

/* Multi-tensor form of pcgan_adam: one launch updates every tensor of a parameter group (the two Adam steps of
 * wsgan_emb_model.py:451-461 walk 48 + 13 tensors).  `step` (device scalar) holds the number of updates already applied:
 * the update uses step + 1 for the bias corrections and a second, one-thread launch behind it stores step + 1. */
typedef struct {
  float* p; const float* g; float* m; float* v;
  int64_t n;
} pcgan_adam_item;
int pcgan_adam_batched(const pcgan_adam_item* items, int32_t count, int64_t max_n, const float* lr, double beta1, double beta2,
                       double eps, float* step, pcgan_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCGAN_KERNELS_H_ */

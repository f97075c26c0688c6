/*
 * bm2f_msda.h — C ABI of the B200-native multi-scale deformable attention library
 * (libbm2f_msda.so, built from bm2f_b200/csrc/ for sm_100a only).
 *
 * This is the drop-in boundary for the one path this repository replaces.  Each entry
 * point names the reference interface it stands in for.  Paths are relative to
 * /root/reference/mask2former/modeling/pixel_decoder/ops/.
 *
 *   reference (pybind module "MultiScaleDeformableAttention", src/vision.cpp:18-21)
 *     ms_deform_attn_forward  (src/ms_deform_attn.h:25-44  -> src/cuda/ms_deform_attn_cuda.cu:25-85)
 *     ms_deform_attn_backward (src/ms_deform_attn.h:46-66  -> src/cuda/ms_deform_attn_cuda.cu:88-158)
 *   and, below them, the launchers ms_deformable_im2col_cuda / ms_deformable_col2im_cuda
 *   (src/cuda/ms_deform_im2col_cuda.cuh:928-959, 961-1332).
 *
 * Conventions
 *   - Plain pointers and sizes; no torch / ATen types.  All pointers are DEVICE pointers
 *     on the current CUDA device unless the name ends in _host.
 *   - Tensors are dense and contiguous (the reference asserts the same, .cu:33-43):
 *       value        (N, S, M, D)         element type = dtype
 *       spatial_shapes (L, 2) int64 = (H_l, W_l);  level_start_index (L) int64 — both read
 *                      ON THE DEVICE (the reference dereferences them in-kernel, .cuh:279-282);
 *                      the library never copies them to the host.
 *       sampling_loc (N, Lq, M, L, P, 2)  last dim = (x, y) normalised to [0,1]
 *       attn_weight  (N, Lq, M, L, P)
 *       output / grad_output (N, Lq, M*D)
 *     For BM2F_DTYPE_F32 / F64 every floating tensor has that type.  For BM2F_DTYPE_BF16 the
 *     value / output / grad_output tensors are bf16 while sampling_loc, attn_weight, their
 *     gradients AND grad_value are float32: a bf16 location has +-0.5 px error at W=256, and
 *     ~44 bf16 atomic adds per grad_value element lose 3e-2 relative (measured), so grad_value
 *     accumulates in fp32 and the caller casts.  The reference has no 16-bit path at all
 *     (.cu:69,139).
 *   - Callee never allocates result buffers: the caller (the torch shim) owns them, so the
 *     framework's caching allocator and stream semantics hold.  `output` need not be
 *     zeroed; `grad_value` is zero-filled by the library (reference: at::zeros, .cu:59,126);
 *     grad_sampling_loc and grad_attn_weight are fully overwritten.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream).  No internal synchronisation: re-entrant across streams and devices.  The library's
 *     static state is: a launch counter (atomic), a thread-local error string, cached device attributes, the
 *     process-wide default tuning (bm2f_msda_set_default_tuning: set it before concurrent use, it is read without a
 *     lock), the diagnostics pointer of bm2f_msda_debug_phase_profile, and — host-buffer entry only — one workspace
 *     per device ordinal, each behind its own lock (calls for different devices run concurrently; two calls for the
 *     same device serialise).
 *   - Return value: 0 on success; a negative BM2F_ERR_* code otherwise, in which case
 *     bm2f_msda_last_error() returns a message for the calling thread.  Unlike the
 *     reference (which only printf's launch failures, .cuh:953-957,1326-1330) kernel-launch
 *     errors are returned.
 *   - There is no CPU implementation and no fallback: with no usable sm_100 device the
 *     calls fail with BM2F_ERR_CUDA.
 */
#ifndef BM2F_MSDA_H_
#define BM2F_MSDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BM2F_MSDA_ABI_VERSION 1

typedef enum {
    BM2F_DTYPE_F32 = 0,
    BM2F_DTYPE_F64 = 1,
    BM2F_DTYPE_BF16 = 2
} bm2f_dtype_t;

typedef enum {
    BM2F_OK = 0,
    BM2F_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, bad dtype)     */
    BM2F_ERR_UNSUPPORTED = -2, /* shape outside what the kernels cover (e.g. L > 16)            */
    BM2F_ERR_CUDA = -3,        /* CUDA runtime / driver error, message has cudaGetErrorString   */
    BM2F_ERR_IM2COL_STEP = -4  /* batch % min(batch, im2col_step) != 0  (.cu:55-57)             */
} bm2f_status_t;

/* Tunables of the D=32 fast path; zero-initialise for defaults.  Exposed so that bench.py and
 * the tests can sweep / pin every variant through the same ABI. */
typedef struct {
    int vec;        /* floats per lane per corner: 4 (8 lanes x 128-bit, default), 2, 1, or 8 =
                       256-bit forward gathers (fp32 forward only; backward then uses 4)          */
    int staging;    /* 0 = default, 1 = TMA tensor-map staging of loc/weights, 2 = direct loads */
    int strip_w;    /* queries per strip row (stage): 8, 16 or 32; 0 = default (32)              */
    int rows;       /* strip rows per job, 0 = auto from problem size                            */
    int ctas_per_sm;/* persistent CTAs per SM: 1 or 2; 0 = default (1)                            */
    int force_generic; /* 1 = always use the any-D kernel                                        */
    int order;      /* 0 = default; 1 = 1-D query order (ignore level geometry)                  */
    int merge;      /* backward: 1 = merge equal-pixel corners in-warp before the REDs; 0/2 = off
                       (default: measured slower on B200)                          */
    int geo;        /* forward kernel: 0 = default: geometry warps with 16-byte records (28 consumer + 3
                       geometry warps; float32, TMA staging, strip 32), else the consumer-lane kernel;
                       2 = consumer-lane geometry kernel msda_fwd_fast_kernel (the round-1 default); 11 = geometry warps
                       with 32-byte lean records; 1 = with predicated records (16 consumer warps); 3 = + 256-bit
                       gathers (L = 3).  All equal bit for bit, except that 0 may differ by one rounding of a bilinear
                       weight (2^-25) for points whose footprint crosses the left border of a level   */
    int bwd;        /* backward kernel: 0 = default (anchor-sorted when it applies — float32, D = 32, M = 8, P = 4, L <= 4,
                       num_query == spatial_size, order == 0 — and batch * num_query >= 32768; per-corner otherwise),
                       1 = per-corner vector REDs (msda_bwd_fast_kernel), 2 = anchor-sorted, error if it does not apply
                        */
    int bwd_margin; /* anchor-sorted backward: window margin around a query tile in pixels of the sampled level; points
                       further out take the slow per-corner path inside the kernel.  0 = default (6)                  */
    int bwd_lanes;  /* anchor-sorted backward: lanes per sampling point, 4 (x 8 channels, default) or 8 (x 4 channels)  */
    int reserved[4]; /* reserved[0]: anchor-sorted backward A/B variant for L = 3 (msda_bwd_sorted.cu: pick) */
} bm2f_msda_tuning_t;

/* Library / ABI identification. */
int bm2f_msda_abi_version(void);
const char *bm2f_msda_build_info(void);

/* Message of the last failing call made by this thread ("" if none). */
const char *bm2f_msda_last_error(void);

/* Number of kernels this library has launched since load (all threads).  bench.py reports
 * the difference over its timed region as `gpu_launches`. */
uint64_t bm2f_msda_launch_count(void);

/* Diagnostics (tools/bwd_phases.py): when device_buffer != NULL, every anchor-sorted backward launched afterwards makes
 * thread 0 of each CTA record the cycles it spends per phase — 8 int64 per CTA: wait for loc/attn, phase 1, scan,
 * scatter, wait for grad_output, phase 2, phase 3, sorted points — into device_buffer[blockIdx.x * 8 ...] (room for
 * 2 * SM count CTAs).  NULL switches it off.  Process-wide, not synchronised: a measurement aid, not an API. */
void bm2f_msda_debug_phase_profile(void *device_buffer);

/* Process-wide default tuning used when a call passes tuning == NULL. */
void bm2f_msda_set_default_tuning(const bm2f_msda_tuning_t *tuning);

/* Precondition shared with the reference host wrapper (.cu:53-57):
 * step = min(batch, im2col_step); batch % step must be 0.  Returns BM2F_OK or
 * BM2F_ERR_IM2COL_STEP. */
int bm2f_msda_check_im2col_step(int batch, int im2col_step);

/*
 * Forward.  Replaces ms_deform_attn_cuda_forward + ms_deformable_im2col_cuda
 * (.cu:25-85, .cuh:928-959):
 *   output[b,q,m,:] = sum_{l,p} attn[b,q,m,l,p] *
 *                     bilinear_zero_pad(value[b, level l, m, :], x*W_l - 0.5, y*H_l - 0.5)
 */
int bm2f_msda_forward(const void *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index, const void *sampling_loc,
                      const void *attn_weight, void *output,
                      int batch, int spatial_size, int num_heads, int channels,
                      int num_levels, int num_query, int num_point,
                      int dtype, const bm2f_msda_tuning_t *tuning, void *stream);

/*
 * Backward.  Replaces ms_deform_attn_cuda_backward + ms_deformable_col2im_cuda and its six
 * kernel variants (.cu:88-158, .cuh:961-1332).
 */
int bm2f_msda_backward(const void *value, const int64_t *spatial_shapes,
                       const int64_t *level_start_index, const void *sampling_loc,
                       const void *attn_weight, const void *grad_output,
                       void *grad_value, void *grad_sampling_loc, void *grad_attn_weight,
                       int batch, int spatial_size, int num_heads, int channels,
                       int num_levels, int num_query, int num_point,
                       int dtype, const bm2f_msda_tuning_t *tuning, void *stream);

/*
 * Fused variants: the sampling kernels also do the element-wise prologue of MSDeformAttn.forward
 * (ops/modules/ms_deform_attn.py:101-109) in registers —
 *     attention_weights = softmax(attn_logits over the L*P points of a (query, head))
 *     sampling_loc      = reference_points[:, :, None, :, None, :] + sampling_offsets / (W_l, H_l)
 * so the normalised weights and the locations are never materialised, and the backward returns the
 * gradients of the raw Linear outputs (softmax backward and the 1/(W,H) scaling included).
 *   reference_points (N, Lq, L, 2) float32, (x, y) in [0,1]   — the 2-d branch of the module; they get no
 *   gradient (Mask2Former builds them from the level shapes, msdeformattn.py:141-153).
 *   reference_points == NULL (needs num_query == spatial_size): exactly those encoder reference points with valid
 *   ratios 1 — every level gets the query pixel's centre ((x + 0.5) / W_q, (y + 0.5) / H_q) — computed in the kernel
 *   from the query index, bit-identical to the tensor torch builds, and without a dependent global load in front of
 *   every gather (forward 1.42 -> ~1.0 ms at cfg 2).
 *   sampling_offsets (N, Lq, M, L, P, 2) float32 in pixels;  attn_logits (N, Lq, M, L, P) float32.
 * Shapes outside D=32, M=8, P=4, L<=4 (f32) / L=3 (bf16) return BM2F_ERR_UNSUPPORTED: compose the
 * prologue yourself and call bm2f_msda_forward/backward.  bm2f_msda_fused_supported() tells in advance.
 */
int bm2f_msda_fused_supported(int num_heads, int channels, int num_levels, int num_point, int dtype);

int bm2f_msda_fused_forward(const void *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const void *reference_points,
                            const void *sampling_offsets, const void *attn_logits, void *output,
                            int batch, int spatial_size, int num_heads, int channels,
                            int num_levels, int num_query, int num_point,
                            int dtype, const bm2f_msda_tuning_t *tuning, void *stream);

int bm2f_msda_fused_backward(const void *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const void *reference_points,
                             const void *sampling_offsets, const void *attn_logits,
                             const void *grad_output, void *grad_value,
                             void *grad_sampling_offsets, void *grad_attn_logits,
                             int batch, int spatial_size, int num_heads, int channels,
                             int num_levels, int num_query, int num_point,
                             int dtype, const bm2f_msda_tuning_t *tuning, void *stream);

/*
 * Packed variants: sampling offsets and attention logits are the column blocks of ONE matrix
 * offsets_logits (batch * num_query, num_heads * num_levels * num_point * 3) — row = [M*L*P*2 offsets | M*L*P logits] —
 * i.e. the output of a single projection with the weights of `sampling_offsets` and `attention_weights` stacked
 * (ops/modules/ms_deform_attn.py:101-102 as one 256 -> 288 GEMM); grad_offsets_logits has the same layout, so the input
 * gradient of both projections is one GEMM as well.  Same kernels and bit-identical results as the unpacked entry points
 * (the matrices are read through TMA tensor maps whose row stride is the packed width; tuning.staging must be 0 or 1).
 */
int bm2f_msda_fused_forward_packed(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                   const void *reference_points, const void *offsets_logits, void *output, int batch,
                                   int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                                   int num_point, int dtype, const bm2f_msda_tuning_t *tuning, void *stream);
int bm2f_msda_fused_backward_packed(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                    const void *reference_points, const void *offsets_logits, const void *grad_output,
                                    void *grad_value, void *grad_offsets_logits, int batch, int spatial_size, int num_heads,
                                    int channels, int num_levels, int num_query, int num_point, int dtype,
                                    const bm2f_msda_tuning_t *tuning, void *stream);

/*
 * Projection GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), for the four nn.Linear layers of
 * MSDeformAttn (ops/modules/ms_deform_attn.py:59-62; applied at :98, :101, :102, :124):
 *     y[rows, out_features] = x[rows, in_features] * weight[out_features, in_features]^T + bias
 * float32 in / out; output width a multiple of 256 or one of {288, 192, 96}, reduction length a multiple of 32
 * (forward: in_features = 256 = d_model; bm2f_linear_backward_input: grad_x = grad_y * weight, width 256).
 * split = 3: three-term TF32 split with fp32 accumulation (fp32-grade result, what the fp32 reference
 * module needs); split = 1: single TF32 pass; nothing else is accepted.  `workspace` =
 * bm2f_linear_workspace_bytes() of device memory for the hi/lo halves of the weight (rewritten on every call).
 * bias may be NULL.
 */
size_t bm2f_linear_workspace_bytes(int out_features, int in_features);

/* Kernel variants of the GEMMs kept for A/B measurements (same results).  Process-wide, read without a
 * lock: a measurement tool sets it, production code never does.  NULL restores the defaults (all zero). */
typedef struct {
    int variant;      /* forward / grad_x GEMM: 0 = default (persistent kernel; single TF32 pass: activation tile by TMA),
                         1 = one tile per CTA, 2 = coalesced-store epilogue instead of the TMA store, 3 = clusters of two
                         CTAs with TMA-multicast weights, 4 / 5 / 6 = more activation k-blocks in flight (8 producer warps
                         x 5 / 4 x 4 / 8 x 4; with split = 1, 5 = register-staged activations), 7 = CTA pairs issuing
                         tcgen05.mma.cta_group::2 */
    int dw_row_cap;   /* weight-gradient GEMM: rows reduced per CTA capped at 256 * dw_row_cap (shortens the fp32
                         accumulation chain in TMEM); 0 = no cap */
    int reserved[6];
} bm2f_linear_tuning_t;
int bm2f_linear_set_tuning(const bm2f_linear_tuning_t *tuning);
int bm2f_linear_forward(const void *x, const void *weight, const void *bias, void *y, void *workspace,
                        int rows, int out_features, int in_features, int split, void *stream);
int bm2f_linear_backward_input(const void *grad_y, const void *weight, void *grad_x, void *workspace,
                               int rows, int out_features, int in_features, int split, void *stream);
/* grad_weight[out, in] = grad_y^T x and grad_bias[out] = column sums of grad_y (may be NULL); both are
 * zeroed and then accumulated with red.global.add by row-chunk CTAs (summation order is free).
 * in_features must be a multiple of 256. */
int bm2f_linear_backward_weight(const void *grad_y, const void *x, void *grad_weight, void *grad_bias,
                                int rows, int out_features, int in_features, int split, void *stream);

/* FFN of the encoder layer (reference: msdeformattn.py:116-120, linear2(relu(linear1(src)))):
 *   bm2f_linear_relu_forward:            y = max(x W^T + b, 0), ReLU in the GEMM epilogue
 *   bm2f_linear_backward_input_masked:   grad_x = (grad_y W) where mask > 0 else 0 — with mask = the hidden
 *     activation this is linear2's input gradient with linear1's ReLU backward applied in the epilogue, so the
 *     masked gradient is written once and feeds linear1's backward GEMMs directly.
 * Output width a multiple of 256, or 192 / 96. */
int bm2f_linear_relu_forward(const void *x, const void *weight, const void *bias, void *y, void *workspace,
                             int rows, int out_features, int in_features, int split, void *stream);
int bm2f_linear_backward_input_masked(const void *grad_y, const void *weight, const void *mask, void *grad_x,
                                      void *workspace, int rows, int out_features, int in_features, int split,
                                      void *stream);
/* grad_x = grad_y * weight + addend, the sum done in the GEMM epilogue (gradient branches meeting at one tensor:
 * q = src + pos feeds two projections, src feeds a third).  addend (rows, in_features) may alias grad_x. */
int bm2f_linear_backward_input_accumulate(const void *grad_y, const void *weight, const void *addend, void *grad_x,
                                          void *workspace, int rows, int out_features, int in_features, int split,
                                          void *stream);

/*
 * Fused residual-add + LayerNorm of the encoder layer (reference: msdeformattn.py:115-131,
 * `src = src + dropout(src2); src = norm(src)` with dropout = 0), 256 channels, float32:
 *   forward : z = x + residual (kept for backward), y = LayerNorm(z) * gamma + beta, per-row mean / rstd
 *   backward: grad_z (= gradient of both x and residual), grad_gamma, grad_beta (zeroed, then accumulated)
 */
int bm2f_add_layernorm_forward(const void *x, const void *residual, const void *gamma, const void *beta,
                               float eps, void *z, void *y, void *mean, void *rstd, int rows, int channels,
                               void *stream);
int bm2f_add_layernorm_backward(const void *grad_y, const void *z, const void *mean, const void *rstd,
                                const void *gamma, void *grad_z, void *grad_gamma, void *grad_beta, int rows,
                                int channels, void *stream);

/* In-place x[row, :] = 0 where row_mask[row] != 0 (bool / uint8 per row), float32 rows of `channels` values:
 * `value.masked_fill(input_padding_mask[..., None], 0)` (ops/modules/ms_deform_attn.py:99-100) and its backward,
 * writing only the masked rows instead of making a full pass. */
int bm2f_zero_masked_rows(void *x, const void *row_mask, int rows, int channels, void *stream);

/*
 * Pixel-decoder glue either side of the encoder, token-major float32 rows of 256 channels.
 *
 * bm2f_transpose_batched: in (batch, rows, cols) -> out (batch, cols, rows).  NCHW backbone feature
 *   (N, C, H*W) -> token rows (N, H*W, C) for the 1x1 `input_proj` conv, which then is bm2f_linear_forward
 *   (msdeformattn.py:214-227, 321), and the reverse for its input gradient.
 *
 * bm2f_groupnorm_tokens_forward: nn.GroupNorm(32, 256) (msdeformattn.py:217, 224) of y (batch, tokens, 256),
 *   statistics over (tokens x 8 channels) per (image, group); row t of image n is written at
 *   out + n * out_batch_stride + t * 256 (floats), i.e. straight into this level's slice of the concatenated
 *   (N, S, 256) encoder input (replaces flatten(2).transpose(1, 2) + torch.cat, msdeformattn.py:73-80).
 *   mean / rstd: (batch, 32) float32, kept for backward.  workspace: batch * 32 * 2 doubles.
 * bm2f_groupnorm_tokens_backward: grad_out rows at grad_out + n * grad_batch_stride + t * 256;
 *   grad_y (batch, tokens, 256) contiguous; grad_gamma / grad_beta (256) are zeroed, then accumulated.
 *   workspace: batch * 32 * 2 doubles + batch * 32 * 2 floats (bm2f_groupnorm_tokens_workspace_bytes).
 *
 * bm2f_sine_position_embedding: PositionEmbeddingSine(num_pos_feats, temperature, normalize, scale) for an
 *   all-False mask (transformer_decoder/position_encoding.py:29-52; msdeformattn.py:322 always passes none):
 *   out (height * width, 2 * num_pos_feats) token-major, identical for every image of the batch.
 */
int bm2f_transpose_batched(const void *in, void *out, int batch, int rows, int cols, void *stream);
size_t bm2f_groupnorm_tokens_workspace_bytes(int batch);
int bm2f_groupnorm_tokens_forward(const void *y, const void *gamma, const void *beta, float eps, void *out,
                                  int64_t out_batch_stride, void *mean, void *rstd, void *workspace, int batch,
                                  int tokens, int channels, int groups, void *stream);
int bm2f_groupnorm_tokens_backward(const void *grad_out, int64_t grad_batch_stride, const void *y, const void *mean,
                                   const void *rstd, const void *gamma, void *grad_y, void *grad_gamma,
                                   void *grad_beta, void *workspace, int batch, int tokens, int channels, int groups,
                                   void *stream);
int bm2f_sine_position_embedding(void *out, int height, int width, int num_pos_feats, float temperature, float scale,
                                 int normalize, void *stream);

/*
 * FPN tail of the pixel decoder, token-major float32 rows of 256 channels; replaces the reference's
 *     cur_fpn = lateral_conv(x); y = cur_fpn + F.interpolate(out[-1], size=cur_fpn.shape[-2:], mode="bilinear",
 *     align_corners=False); y = output_conv(y); mask_features(y)              (msdeformattn.py:341-358)
 * where lateral_conv = 1x1 conv + GroupNorm (bm2f_linear_forward + bm2f_groupnorm_tokens_stats), output_conv = 3x3 conv
 * + GroupNorm + ReLU, mask_features = 1x1 conv with bias (bm2f_linear_forward).
 *
 * "haloed" image = (batch, height + 2, width + 2, 256) rows whose border rows are zero: with the padding in memory a
 * filter tap of the 3x3 convolution is a constant row shift, so its activation tiles are plain 2-D TMA boxes.
 *
 * bm2f_conv3x3_forward: y (batch, height, width, 256) dense = conv2d(x, weight (256, 256, 3, 3), padding = 1, no bias)
 *   on tcgen05; x_halo haloed.  split as for bm2f_linear_forward (1 = single TF32 pass, what cuDNN does for the
 *   reference's nn.Conv2d under torch's default flags; 3 = tf32x3).  workspace: bm2f_conv3x3_workspace_bytes().
 * bm2f_conv3x3_backward_input: grad_x dense from the haloed gradient of y (the same kernel with flipped taps).
 * bm2f_conv3x3_backward_weight: grad_weight (256, 256, 3, 3) from the haloed gradient and the haloed input (split 1:
 *   both operands reach the tensor core through TMA as MN-major tiles; split 3: transposing-producer kernel, one launch
 *   over the 9 taps).
 * bm2f_groupnorm_tokens_stats: mean / rstd (batch, 32) of y (batch, tokens, 256); workspace as for
 *   bm2f_groupnorm_tokens_forward.
 * bm2f_fpn_merge_forward: y_halo = GroupNorm(lateral; mean, rstd, gamma, beta) + bilinear upsample (align_corners =
 *   False, torch's operation order) of enc (image n at enc + n * enc_batch_stride floats, enc_height x enc_width rows of
 *   256) to height x width; writes the zero halo as well.
 * bm2f_fpn_upsample_backward: grad_enc = adjoint of that upsample applied to grad_y (dense), as a gather: deterministic,
 *   every grad_enc row is written exactly once.
 * bm2f_groupnorm_relu_tokens_apply: out = relu(GroupNorm(y)).
 * bm2f_groupnorm_relu_tokens_backward: gradient of relu(GroupNorm(y)) with respect to y, written as a haloed image
 *   (the conv gradients read it); grad_gamma / grad_beta (256) are zeroed, then accumulated.  workspace:
 *   bm2f_groupnorm_tokens_workspace_bytes(batch).
 */
size_t bm2f_conv3x3_workspace_bytes(int out_channels, int in_channels);
/* A/B only (process-wide, like bm2f_linear_set_tuning): 1 (default) = the single-pass kernel shares every weight k-block
 * between two row tiles, 0 = one row tile per weight k-block.  Same results. */
int bm2f_conv3x3_set_variant(int pair_tiles);
int bm2f_conv3x3_forward(const void *x_halo, const void *weight, void *y, void *workspace, int batch, int height, int width,
                         int out_channels, int in_channels, int split, void *stream);
int bm2f_conv3x3_backward_input(const void *grad_halo, const void *weight, void *grad_x, void *workspace, int batch,
                                int height, int width, int out_channels, int in_channels, int split, void *stream);
int bm2f_conv3x3_backward_weight(const void *grad_halo, const void *x_halo, void *grad_weight, void *workspace, int batch,
                                 int height, int width, int out_channels, int in_channels, int split, void *stream);
int bm2f_groupnorm_tokens_stats(const void *y, float eps, void *mean, void *rstd, void *workspace, int batch, int tokens,
                                int channels, int groups, void *stream);
int bm2f_fpn_merge_forward(const void *lateral, const void *mean, const void *rstd, const void *gamma, const void *beta,
                           const void *enc, int64_t enc_batch_stride, void *y_halo, int batch, int height, int width,
                           int enc_height, int enc_width, int channels, void *stream);
int bm2f_fpn_upsample_backward(const void *grad_y, void *grad_enc, int64_t grad_enc_batch_stride, int batch, int height,
                               int width, int enc_height, int enc_width, int channels, void *stream);
int bm2f_groupnorm_relu_tokens_apply(const void *y, const void *mean, const void *rstd, const void *gamma, const void *beta,
                                     void *out, int batch, int tokens, int channels, int groups, void *stream);
int bm2f_groupnorm_relu_tokens_backward(const void *grad_out, const void *y, const void *mean, const void *rstd,
                                        const void *gamma, const void *beta, void *grad_halo, void *grad_gamma,
                                        void *grad_beta, void *workspace, int batch, int height, int width, int channels,
                                        int groups, void *stream);

/*
 * Host-buffer convenience used for end-to-end measurement and by non-torch callers:
 * every pointer is a HOST pointer (pinned memory gives full PCIe rate).  The library copies
 * the inputs to a device workspace it owns for the duration of the call, runs forward and,
 * when grad_output_host != NULL, backward, copies the results back and synchronises.
 * The batch is pipelined in image chunks over three streams / workspace slots (uploads of a chunk first, then
 * both kernels, then its downloads), so the H2D engine, the SMs and the D2H engine work on different chunks.
 * Any of the result pointers may be NULL to skip that copy-back.  The workspace is kept between calls (sized by the
 * largest call so far); bm2f_msda_release_host_workspace() frees it and the streams.
 */
int bm2f_msda_release_host_workspace(void);
int bm2f_msda_forward_backward_host(const void *value_host, const int64_t *spatial_shapes_host,
                                    const int64_t *level_start_index_host,
                                    const void *sampling_loc_host, const void *attn_weight_host,
                                    const void *grad_output_host, void *output_host,
                                    void *grad_value_host, void *grad_sampling_loc_host,
                                    void *grad_attn_weight_host,
                                    int batch, int spatial_size, int num_heads, int channels,
                                    int num_levels, int num_query, int num_point,
                                    int dtype, const bm2f_msda_tuning_t *tuning);

#ifdef __cplusplus
}
#endif
#endif /* BM2F_MSDA_H_ */

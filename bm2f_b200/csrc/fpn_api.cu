// C ABI of the FPN tail of the pixel decoder (fpn_kernels.cuh; reference msdeformattn.py:341-358):
// 3 x 3 convolution on tcgen05 over zero-haloed token images, the GroupNorm / upsample / merge kernels around it.
#include "api_common.cuh"
#include "glue_kernels.cuh"
#include "fpn_kernels.cuh"

using namespace bm2f;
using namespace bm2f::host;

namespace {
constexpr int kC = kFpnC;
int g_conv_pair_tiles = 1;      // A/B: bm2f_conv3x3_set_variant(0) = one row tile per weight k-block

int check_device(int *sms)
{
    int cc = 0;
    const int rc = device_info(sms, &cc);
    if (rc) return rc;
    if (cc < 10) return fail(BM2F_ERR_CUDA, "this library contains sm_100a code only; device has cc %d.x", cc);
    return BM2F_OK;
}

int check_image(int batch, int height, int width, int out_channels, int in_channels)
{
    if (batch <= 0 || height <= 0 || width <= 0) return fail(BM2F_ERR_INVALID, "conv3x3: batch / height / width must be positive");
    if (out_channels != kC || in_channels != kC)
        return fail(BM2F_ERR_UNSUPPORTED, "conv3x3: instantiated for 256 -> 256 channels (got %d -> %d)", in_channels, out_channels);
    const long long rows = static_cast<long long>(batch) * (height + 2) * (width + 2);
    if (rows > 0x7fffffffLL - 4096) return fail(BM2F_ERR_UNSUPPORTED, "conv3x3: image too large (%lld haloed rows)", rows);
    return BM2F_OK;
}

// y (dense) = conv(x_halo) with GEMM weights w_hi / w_lo (256, 9 * 256) already prepared in the workspace
int launch_conv_gemm(const float *x_halo, const float *w_hi, const float *w_lo, float *y, int batch, int height, int width,
                     int split, int sms, cudaStream_t st)
{
    LinearParams p{};
    p.x = x_halo; p.y = y; p.bias = nullptr;
    p.M = batch * (height + 2) * (width + 2);
    p.N = kC; p.K = 9 * kC; p.slices = 1; p.split = split;
    p.ldx = kC; p.conv_w = width; p.conv_h = height;
    CUtensorMap mh, ml, mx;
    int rc;
    if ((rc = make_map(&mh, w_hi, kC, 9 * kC, kC, kGemmBlockK, true))) return rc;
    if ((rc = make_map(&ml, w_lo, kC, 9 * kC, kC, kGemmBlockK, true))) return rc;
    if ((rc = make_map(&mx, x_halo, p.M, kC, kGemmBlockM, kGemmBlockK, true))) return rc;
    const int tiles = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    int grid = tiles < sms ? tiles : sms;
    if (split == 1 && g_conv_pair_tiles) {
        // two row tiles per weight k-block (RT = 2)
        constexpr auto kern = &linear_tf32x3_persistent_kernel<256, 1, kGemmProducerWarps, 3, false, true, true, 2>;
        constexpr int smem = linear_conv2_smem_bytes<256>();
        if ((rc = ensure_dynamic_smem<kern>(smem, "cudaFuncSetAttribute(conv3x3 smem)"))) return rc;
        const int pairs = (tiles + 1) / 2;
        grid = pairs < sms ? pairs : sms;
        kern<<<grid, kGemmThreadsPersistent, smem, st>>>(p, mh, mh, mx, mx);
    } else if (split == 1) {
        constexpr auto kern = &linear_tf32x3_persistent_kernel<256, 1, kGemmProducerWarps, 3, false, true, true>;
        constexpr int smem = linear_xtma_smem_bytes<256>();
        if ((rc = ensure_dynamic_smem<kern>(smem, "cudaFuncSetAttribute(conv3x3 smem)"))) return rc;
        kern<<<grid, kGemmThreadsPersistent, smem, st>>>(p, mh, mh, mx, mx);
    } else {
        constexpr auto kern = &linear_tf32x3_persistent_kernel<256, 1, kGemmProducerWarps, 3, false, false, true>;
        constexpr int smem = linear_persistent_smem_bytes<256>();
        if ((rc = ensure_dynamic_smem<kern>(smem, "cudaFuncSetAttribute(conv3x3 tf32x3 smem)"))) return rc;
        kern<<<grid, kGemmThreadsPersistent, smem, st>>>(p, mh, ml, mx, mx);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch conv3x3 kernel");
    count_launch(1);
    return BM2F_OK;
}

int conv_common(const void *x_halo, const void *weight, void *y, void *workspace, int batch, int height, int width,
                int out_channels, int in_channels, int split, void *stream, int mode)
{
    if (!x_halo || !weight || !y || !workspace) return fail(BM2F_ERR_INVALID, "null pointer");
    int rc = check_image(batch, height, width, out_channels, in_channels);
    if (rc) return rc;
    if (split != 1 && split != 3) return fail(BM2F_ERR_INVALID, "split must be 3 (tf32x3) or 1 (single TF32 pass)");
    if (!aligned16(x_halo) || !aligned16(y) || !aligned16(weight) || !aligned16(workspace))
        return fail(BM2F_ERR_UNSUPPORTED, "conv3x3: tensors must be 16-byte aligned");
    int sms = 0;
    if ((rc = check_device(&sms))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *w_hi = static_cast<float *>(workspace);
    float *w_lo = w_hi + static_cast<size_t>(kC) * 9 * kC;
    const int n = kC * kC * 9;
    conv3x3_weight_prep_kernel<<<(n + 255) / 256, 256, 0, st>>>(static_cast<const float *>(weight), w_hi, w_lo, kC, kC, mode);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch conv3x3_weight_prep_kernel");
    count_launch(1);
    return launch_conv_gemm(static_cast<const float *>(x_halo), w_hi, w_lo, static_cast<float *>(y), batch, height, width,
                            split, sms, st);
}
}  // namespace

extern "C" {

int bm2f_conv3x3_set_variant(int pair_tiles)
{
    g_conv_pair_tiles = pair_tiles ? 1 : 0;
    return BM2F_OK;
}

size_t bm2f_conv3x3_workspace_bytes(int out_channels, int in_channels)
{
    if (out_channels <= 0 || in_channels <= 0) return 0;
    return static_cast<size_t>(2) * out_channels * in_channels * 9 * sizeof(float);
}

int bm2f_conv3x3_forward(const void *x_halo, const void *weight, void *y, void *workspace, int batch, int height, int width,
                         int out_channels, int in_channels, int split, void *stream)
{
    return conv_common(x_halo, weight, y, workspace, batch, height, width, out_channels, in_channels, split, stream, 0);
}

int bm2f_conv3x3_backward_input(const void *grad_halo, const void *weight, void *grad_x, void *workspace, int batch,
                                int height, int width, int out_channels, int in_channels, int split, void *stream)
{
    // grad_x[p, c] = sum_t sum_o grad[p - shift(t), o] * w[o, c, t]: the same convolution with flipped taps and the
    // channel roles exchanged
    return conv_common(grad_halo, weight, grad_x, workspace, batch, height, width, in_channels, out_channels, split, stream, 1);
}

int bm2f_conv3x3_backward_weight(const void *grad_halo, const void *x_halo, void *grad_weight, void *workspace, int batch,
                                 int height, int width, int out_channels, int in_channels, int split, void *stream)
{
    if (!grad_halo || !x_halo || !grad_weight || !workspace) return fail(BM2F_ERR_INVALID, "null pointer");
    int rc = check_image(batch, height, width, out_channels, in_channels);
    if (rc) return rc;
    if (split != 1 && split != 3) return fail(BM2F_ERR_INVALID, "split must be 3 (tf32x3) or 1 (single TF32 pass)");
    if (!aligned16(grad_halo) || !aligned16(x_halo) || !aligned16(grad_weight) || !aligned16(workspace))
        return fail(BM2F_ERR_UNSUPPORTED, "conv3x3: tensors must be 16-byte aligned");
    int sms = 0;
    if ((rc = check_device(&sms))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rows = batch * (height + 2) * (width + 2);
    float *dwk = static_cast<float *>(workspace);                          // (256, 9 * 256) in the GEMM layout
    cudaError_t e = cudaMemsetAsync(dwk, 0, static_cast<size_t>(kC) * 9 * kC * 4, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(conv3x3 grad_weight)");
    int chunks = sms / 9;
    if (chunks < 1) chunks = 1;
    int rows_per_chunk = ((rows + chunks - 1) / chunks + 31) / 32 * 32;
    chunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
    if (split == 1) {
        DwTmaParams p{};
        p.dw = dwk; p.ld_dw = 9 * kC; p.units = 9; p.rows = rows; p.rows_per_chunk = rows_per_chunk; p.conv_wp = width + 2;
        CUtensorMap mg, mx;
        if ((rc = make_map(&mg, static_cast<const float *>(grad_halo), rows, kC, 32, 32, 2))) return rc;
        if ((rc = make_map(&mx, static_cast<const float *>(x_halo), rows, kC, 32, 32, 2))) return rc;
        constexpr int smem = linear_dw_tma_smem_bytes();
        if ((rc = ensure_dynamic_smem<&linear_dw_tma_kernel>(smem, "cudaFuncSetAttribute(dW TMA smem)"))) return rc;
        linear_dw_tma_kernel<<<dim3(9, chunks), kDwTmaThreads, smem, st>>>(p, mg, mx);
    } else {
        // two 128-feature halves x 9 taps x row chunks on the transposing-producer kernel
        chunks = sms / 18;
        if (chunks < 1) chunks = 1;
        rows_per_chunk = ((rows + chunks - 1) / chunks + 31) / 32 * 32;
        chunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
        LinearDwParams p{};
        p.g = static_cast<const float *>(grad_halo); p.x = static_cast<const float *>(x_halo);
        p.dw = dwk; p.db = nullptr;
        p.M = rows; p.N = kC; p.ldx = kC; p.rows_per_chunk = rows_per_chunk; p.split = 3;
        p.conv_wp = width + 2; p.ld_dw = 9 * kC;
        constexpr int smem = linear_dw_smem_bytes();
        if ((rc = ensure_dynamic_smem<&linear_dw_tf32x3_kernel>(smem, "cudaFuncSetAttribute(dW smem)"))) return rc;
        linear_dw_tf32x3_kernel<<<dim3(2, chunks, 9), kDwThreads, smem, st>>>(p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch conv3x3 weight-gradient kernel");
    const int n = kC * kC * 9;
    conv3x3_weight_grad_unpack_kernel<<<(n + 255) / 256, 256, 0, st>>>(dwk, static_cast<float *>(grad_weight), kC, kC);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch conv3x3_weight_grad_unpack_kernel");
    count_launch(2);
    return BM2F_OK;
}

int bm2f_fpn_merge_forward(const void *lateral, const void *mean, const void *rstd, const void *gamma, const void *beta,
                           const void *enc, int64_t enc_batch_stride, void *y_halo, int batch, int height, int width,
                           int enc_height, int enc_width, int channels, void *stream)
{
    if (!lateral || !mean || !rstd || !gamma || !beta || !enc || !y_halo) return fail(BM2F_ERR_INVALID, "null pointer");
    if (channels != kC) return fail(BM2F_ERR_UNSUPPORTED, "fpn merge: 256 channels only (got %d)", channels);
    if (batch <= 0 || height <= 0 || width <= 0 || enc_height <= 0 || enc_width <= 0)
        return fail(BM2F_ERR_INVALID, "fpn merge: non-positive dimension");
    if (!aligned16(lateral) || !aligned16(enc) || !aligned16(y_halo) || !aligned16(gamma) || !aligned16(beta) ||
        enc_batch_stride % 4 != 0)
        return fail(BM2F_ERR_UNSUPPORTED, "fpn merge: tensors must be 16-byte aligned");
    int sms = 0;
    int rc = check_device(&sms);
    if (rc) return rc;
    const long long rows = static_cast<long long>(batch) * (height + 2) * (width + 2);
    const int grid = static_cast<int>(std::min<long long>((rows + 7) / 8, static_cast<long long>(sms) * 16));
    // torch: scale = input_size / output_size in float (area_pixel_compute_scale, align_corners = False)
    const float sh = static_cast<float>(enc_height) / static_cast<float>(height);
    const float sw = static_cast<float>(enc_width) / static_cast<float>(width);
    fpn_merge_forward_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float *>(lateral), static_cast<const float *>(mean), static_cast<const float *>(rstd),
        static_cast<const float *>(gamma), static_cast<const float *>(beta), static_cast<const float *>(enc),
        static_cast<long long>(enc_batch_stride), static_cast<float *>(y_halo), batch, height, width, enc_height, enc_width,
        sh, sw);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch fpn_merge_forward_kernel");
    count_launch(1);
    return BM2F_OK;
}

int bm2f_fpn_upsample_backward(const void *grad_y, void *grad_enc, int64_t grad_enc_batch_stride, int batch, int height,
                               int width, int enc_height, int enc_width, int channels, void *stream)
{
    if (!grad_y || !grad_enc) return fail(BM2F_ERR_INVALID, "null pointer");
    if (channels != kC) return fail(BM2F_ERR_UNSUPPORTED, "fpn upsample backward: 256 channels only (got %d)", channels);
    if (batch <= 0 || height <= 0 || width <= 0 || enc_height <= 0 || enc_width <= 0)
        return fail(BM2F_ERR_INVALID, "fpn upsample backward: non-positive dimension");
    if (!aligned16(grad_y) || !aligned16(grad_enc) || grad_enc_batch_stride % 4 != 0)
        return fail(BM2F_ERR_UNSUPPORTED, "fpn upsample backward: tensors must be 16-byte aligned");
    int sms = 0;
    int rc = check_device(&sms);
    if (rc) return rc;
    const long long pix = static_cast<long long>(batch) * enc_height * enc_width;
    const int grid = static_cast<int>(std::min<long long>((pix + 7) / 8, static_cast<long long>(sms) * 16));
    const float sh = static_cast<float>(enc_height) / static_cast<float>(height);
    const float sw = static_cast<float>(enc_width) / static_cast<float>(width);
    fpn_upsample_backward_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float *>(grad_y), static_cast<float *>(grad_enc), static_cast<long long>(grad_enc_batch_stride),
        batch, height, width, enc_height, enc_width, sh, sw);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch fpn_upsample_backward_kernel");
    count_launch(1);
    return BM2F_OK;
}

int bm2f_groupnorm_tokens_stats(const void *y, float eps, void *mean, void *rstd, void *workspace, int batch, int tokens,
                                int channels, int groups, void *stream)
{
    if (!y || !mean || !rstd || !workspace) return fail(BM2F_ERR_INVALID, "null pointer");
    if (channels != kC || groups != kFpnGroups)
        return fail(BM2F_ERR_UNSUPPORTED, "token GroupNorm: 256 channels in 32 groups only (got %d / %d)", channels, groups);
    if (batch <= 0 || tokens <= 0) return fail(BM2F_ERR_INVALID, "token GroupNorm: batch / tokens must be positive");
    int sms = 0;
    int rc = check_device(&sms);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *sums = static_cast<double *>(workspace);
    cudaError_t e = cudaMemsetAsync(sums, 0, static_cast<size_t>(batch) * groups * 2 * sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(GroupNorm sums)");
    int gx = (sms * 8 + batch - 1) / batch;
    gx = std::max(1, std::min(gx, (tokens + 7) / 8));
    groupnorm_tokens_stats_kernel<<<dim3(gx, batch), 256, 0, st>>>(static_cast<const float *>(y), sums, tokens);
    const int n_stats = batch * groups;
    groupnorm_tokens_finalize_kernel<<<(n_stats + 127) / 128, 128, 0, st>>>(
        sums, static_cast<float *>(mean), static_cast<float *>(rstd), n_stats,
        static_cast<double>(tokens) * (channels / groups), eps, 1);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch token GroupNorm statistics");
    count_launch(2);
    return BM2F_OK;
}

int bm2f_groupnorm_relu_tokens_apply(const void *y, const void *mean, const void *rstd, const void *gamma, const void *beta,
                                     void *out, int batch, int tokens, int channels, int groups, void *stream)
{
    if (!y || !mean || !rstd || !gamma || !beta || !out) return fail(BM2F_ERR_INVALID, "null pointer");
    if (channels != kC || groups != kFpnGroups)
        return fail(BM2F_ERR_UNSUPPORTED, "token GroupNorm: 256 channels in 32 groups only (got %d / %d)", channels, groups);
    if (batch <= 0 || tokens <= 0) return fail(BM2F_ERR_INVALID, "token GroupNorm: batch / tokens must be positive");
    int sms = 0;
    int rc = check_device(&sms);
    if (rc) return rc;
    int gx = (sms * 8 + batch - 1) / batch;
    gx = std::max(1, std::min(gx, (tokens + 7) / 8));
    groupnorm_relu_apply_kernel<<<dim3(gx, batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float *>(y), static_cast<const float *>(mean), static_cast<const float *>(rstd),
        static_cast<const float *>(gamma), static_cast<const float *>(beta), static_cast<float *>(out), tokens);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch groupnorm_relu_apply_kernel");
    count_launch(1);
    return BM2F_OK;
}

int bm2f_groupnorm_relu_tokens_backward(const void *grad_out, const void *y, const void *mean, const void *rstd,
                                        const void *gamma, const void *beta, void *grad_halo, void *grad_gamma,
                                        void *grad_beta, void *workspace, int batch, int height, int width, int channels,
                                        int groups, void *stream)
{
    if (!grad_out || !y || !mean || !rstd || !gamma || !beta || !grad_halo || !grad_gamma || !grad_beta || !workspace)
        return fail(BM2F_ERR_INVALID, "null pointer");
    if (channels != kC || groups != kFpnGroups)
        return fail(BM2F_ERR_UNSUPPORTED, "token GroupNorm: 256 channels in 32 groups only (got %d / %d)", channels, groups);
    if (batch <= 0 || height <= 0 || width <= 0) return fail(BM2F_ERR_INVALID, "token GroupNorm: non-positive dimension");
    int sms = 0;
    int rc = check_device(&sms);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tokens = height * width;
    double *sums = static_cast<double *>(workspace);
    float *c1 = reinterpret_cast<float *>(sums + static_cast<size_t>(batch) * groups * 2);
    float *c2 = c1 + static_cast<size_t>(batch) * groups;
    cudaError_t e = cudaMemsetAsync(sums, 0, static_cast<size_t>(batch) * groups * 2 * sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(grad_gamma, 0, channels * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(grad_beta, 0, channels * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(GroupNorm backward sums)");
    int gx = (sms * 4 + batch - 1) / batch;
    gx = std::max(1, std::min(gx, (tokens + 7) / 8));
    groupnorm_relu_bwd_stats_kernel<<<dim3(gx, batch), 256, 0, st>>>(
        static_cast<const float *>(grad_out), static_cast<const float *>(y), static_cast<const float *>(mean),
        static_cast<const float *>(rstd), static_cast<const float *>(gamma), static_cast<const float *>(beta), sums,
        static_cast<float *>(grad_gamma), static_cast<float *>(grad_beta), tokens);
    const int n_stats = batch * groups;
    groupnorm_tokens_finalize_kernel<<<(n_stats + 127) / 128, 128, 0, st>>>(
        sums, c1, c2, n_stats, static_cast<double>(tokens) * (channels / groups), 0.f, 0);
    const long long rows = static_cast<long long>(batch) * (height + 2) * (width + 2);
    const int grid = static_cast<int>(std::min<long long>((rows + 7) / 8, static_cast<long long>(sms) * 16));
    groupnorm_relu_bwd_apply_kernel<<<grid, 256, 0, st>>>(
        static_cast<const float *>(grad_out), static_cast<const float *>(y), static_cast<const float *>(mean),
        static_cast<const float *>(rstd), c1, c2, static_cast<const float *>(gamma), static_cast<const float *>(beta),
        static_cast<float *>(grad_halo), batch, height, width);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch token GroupNorm + ReLU backward");
    count_launch(3);
    return BM2F_OK;
}

}  // extern "C"

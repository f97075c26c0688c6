// C ABI of the encoder-layer and pixel-decoder glue kernels (ln_kernels.cuh, glue_kernels.cuh).
#include "api_common.cuh"
#include "ln_kernels.cuh"
#include "glue_kernels.cuh"

using namespace bm2f;
using namespace bm2f::host;

extern "C" {

// ---------------------------------------------------------------------------------------------
// fused residual-add + LayerNorm (ln_kernels.cuh)
// ---------------------------------------------------------------------------------------------
int bm2f_add_layernorm_forward(const void *x, const void *residual, const void *gamma, const void *beta, float eps,
                               void *z, void *y, void *mean, void *rstd, int rows, int channels, void *stream)
{
    if (!x || !residual || !gamma || !beta || !z || !y || !mean || !rstd) return fail(BM2F_ERR_INVALID, "null pointer");
    if (rows <= 0) return fail(BM2F_ERR_INVALID, "rows must be positive");
    if (channels != kLnC) return fail(BM2F_ERR_UNSUPPORTED, "add+LayerNorm is built for %d channels (got %d)", kLnC, channels);
    if (!aligned16(x) || !aligned16(residual) || !aligned16(gamma) || !aligned16(beta) || !aligned16(z) || !aligned16(y))
        return fail(BM2F_ERR_UNSUPPORTED, "add+LayerNorm: tensors must be 16-byte aligned");
    int sms = 0, cc = 0;
    int rc = device_info(&sms, &cc);
    if (rc) return rc;
    const int blocks = (rows + 7) / 8 < sms * 8 ? (rows + 7) / 8 : sms * 8;
    add_layernorm_fwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float *>(x), static_cast<const float *>(residual), static_cast<const float *>(gamma),
        static_cast<const float *>(beta), eps, static_cast<float *>(z), static_cast<float *>(y),
        static_cast<float *>(mean), static_cast<float *>(rstd), rows);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch add_layernorm_fwd_kernel");
    count_launch(1);
    return BM2F_OK;
}

int bm2f_add_layernorm_backward(const void *grad_y, const void *z, const void *mean, const void *rstd, const void *gamma,
                                void *grad_z, void *grad_gamma, void *grad_beta, int rows, int channels, void *stream)
{
    if (!grad_y || !z || !mean || !rstd || !gamma || !grad_z || !grad_gamma || !grad_beta)
        return fail(BM2F_ERR_INVALID, "null pointer");
    if (rows <= 0) return fail(BM2F_ERR_INVALID, "rows must be positive");
    if (channels != kLnC) return fail(BM2F_ERR_UNSUPPORTED, "add+LayerNorm is built for %d channels (got %d)", kLnC, channels);
    if (!aligned16(grad_y) || !aligned16(z) || !aligned16(gamma) || !aligned16(grad_z))
        return fail(BM2F_ERR_UNSUPPORTED, "add+LayerNorm: tensors must be 16-byte aligned");
    int sms = 0, cc = 0;
    int rc = device_info(&sms, &cc);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(grad_gamma, 0, kLnC * 4, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(grad_beta, 0, kLnC * 4, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_gamma / grad_beta)");
    const int blocks = (rows + 7) / 8 < sms * 4 ? (rows + 7) / 8 : sms * 4;
    add_layernorm_bwd_kernel<<<blocks, 256, 0, st>>>(static_cast<const float *>(grad_y), static_cast<const float *>(z),
                                                     static_cast<const float *>(mean), static_cast<const float *>(rstd),
                                                     static_cast<const float *>(gamma), static_cast<float *>(grad_z),
                                                     static_cast<float *>(grad_gamma), static_cast<float *>(grad_beta), rows);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch add_layernorm_bwd_kernel");
    count_launch(1);
    return BM2F_OK;
}

int bm2f_zero_masked_rows(void *x, const void *row_mask, int rows, int channels, void *stream)
{
    if (!x || !row_mask) return fail(BM2F_ERR_INVALID, "null pointer");
    if (rows <= 0 || channels <= 0 || channels % 4 != 0 || !aligned16(x))
        return fail(BM2F_ERR_UNSUPPORTED, "zero_masked_rows: rows > 0, channels %% 4 == 0, 16-byte aligned rows");
    int sms = 0, cc = 0;
    int rc = device_info(&sms, &cc);
    if (rc) return rc;
    const int blocks = (rows + 7) / 8 < sms * 8 ? (rows + 7) / 8 : sms * 8;
    zero_masked_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float *>(x), static_cast<const unsigned char *>(row_mask), rows, channels);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch zero_masked_rows_kernel");
    count_launch(1);
    return BM2F_OK;
}

// ---------------------------------------------------------------------------------------------
// pixel-decoder glue (glue_kernels.cuh)
// ---------------------------------------------------------------------------------------------
int bm2f_transpose_batched(const void *in, void *out, int batch, int rows, int cols, void *stream)
{
    if (!in || !out) return fail(BM2F_ERR_INVALID, "null pointer");
    if (batch <= 0 || rows <= 0 || cols <= 0) return fail(BM2F_ERR_INVALID, "batch / rows / cols must be positive");
    if (batch > 65535 || (rows + 31) / 32 > 65535) return fail(BM2F_ERR_UNSUPPORTED, "transpose: batch or rows too large");
    const dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
    transpose_batched_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float *>(in),
                                                                                  static_cast<float *>(out), rows, cols);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch transpose_batched_kernel");
    count_launch(1);
    return BM2F_OK;
}

size_t bm2f_groupnorm_tokens_workspace_bytes(int batch)
{
    return static_cast<size_t>(batch > 0 ? batch : 0) * kGnGroups * (2 * sizeof(double) + 2 * sizeof(float));
}

namespace {
int gn_check(int batch, int tokens, int channels, int groups)
{
    if (batch <= 0 || tokens <= 0) return fail(BM2F_ERR_INVALID, "batch / tokens must be positive");
    if (channels != kGnC || groups != kGnGroups)
        return fail(BM2F_ERR_UNSUPPORTED, "token GroupNorm is built for %d channels in %d groups (got %d / %d)", kGnC,
                    kGnGroups, channels, groups);
    if (batch > 65535) return fail(BM2F_ERR_UNSUPPORTED, "token GroupNorm: batch too large");
    return BM2F_OK;
}
int gn_chunks(int batch, int tokens, int sms)
{
    int chunks = (sms * 4 + batch - 1) / batch;          // ~4 CTAs per SM over the whole batch
    const int max_chunks = (tokens + 7) / 8;             // at least one row per warp
    if (chunks > max_chunks) chunks = max_chunks;
    return chunks < 1 ? 1 : chunks;
}
}  // namespace

int bm2f_groupnorm_tokens_forward(const void *y, const void *gamma, const void *beta, float eps, void *out,
                                  int64_t out_batch_stride, void *mean, void *rstd, void *workspace, int batch, int tokens,
                                  int channels, int groups, void *stream)
{
    if (!y || !gamma || !beta || !out || !mean || !rstd || !workspace) return fail(BM2F_ERR_INVALID, "null pointer");
    int rc = gn_check(batch, tokens, channels, groups);
    if (rc) return rc;
    if (!aligned16(y) || !aligned16(out) || !aligned16(gamma) || !aligned16(beta) || out_batch_stride % 4 != 0)
        return fail(BM2F_ERR_UNSUPPORTED, "token GroupNorm: tensors must be 16-byte aligned");
    int sms = 0, cc = 0;
    if ((rc = device_info(&sms, &cc))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *sums = static_cast<double *>(workspace);
    const int n_stats = batch * kGnGroups;
    cudaError_t e = cudaMemsetAsync(sums, 0, static_cast<size_t>(n_stats) * 2 * sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(GroupNorm sums)");
    const dim3 grid(gn_chunks(batch, tokens, sms), batch);
    groupnorm_tokens_stats_kernel<<<grid, 256, 0, st>>>(static_cast<const float *>(y), sums, tokens);
    groupnorm_tokens_finalize_kernel<<<(n_stats + 127) / 128, 128, 0, st>>>(
        sums, static_cast<float *>(mean), static_cast<float *>(rstd), n_stats, static_cast<double>(tokens) * (kGnC / kGnGroups),
        eps, 1);
    groupnorm_tokens_apply_kernel<<<grid, 256, 0, st>>>(static_cast<const float *>(y), static_cast<const float *>(mean),
                                                        static_cast<const float *>(rstd), static_cast<const float *>(gamma),
                                                        static_cast<const float *>(beta), static_cast<float *>(out),
                                                        static_cast<long long>(out_batch_stride), tokens);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch groupnorm_tokens kernels");
    count_launch(3);
    return BM2F_OK;
}

int bm2f_groupnorm_tokens_backward(const void *grad_out, int64_t grad_batch_stride, const void *y, const void *mean,
                                   const void *rstd, const void *gamma, void *grad_y, void *grad_gamma, void *grad_beta,
                                   void *workspace, int batch, int tokens, int channels, int groups, void *stream)
{
    if (!grad_out || !y || !mean || !rstd || !gamma || !grad_y || !grad_gamma || !grad_beta || !workspace)
        return fail(BM2F_ERR_INVALID, "null pointer");
    int rc = gn_check(batch, tokens, channels, groups);
    if (rc) return rc;
    if (!aligned16(grad_out) || !aligned16(y) || !aligned16(grad_y) || !aligned16(gamma) || grad_batch_stride % 4 != 0)
        return fail(BM2F_ERR_UNSUPPORTED, "token GroupNorm: tensors must be 16-byte aligned");
    int sms = 0, cc = 0;
    if ((rc = device_info(&sms, &cc))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_stats = batch * kGnGroups;
    double *sums = static_cast<double *>(workspace);
    float *c1 = reinterpret_cast<float *>(sums + static_cast<size_t>(n_stats) * 2);
    float *c2 = c1 + n_stats;
    cudaError_t e = cudaMemsetAsync(sums, 0, static_cast<size_t>(n_stats) * 2 * sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(grad_gamma, 0, kGnC * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(grad_beta, 0, kGnC * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(GroupNorm backward sums)");
    const dim3 grid(gn_chunks(batch, tokens, sms), batch);
    groupnorm_tokens_bwd_stats_kernel<<<grid, 256, 0, st>>>(
        static_cast<const float *>(grad_out), static_cast<long long>(grad_batch_stride), static_cast<const float *>(y),
        static_cast<const float *>(mean), static_cast<const float *>(rstd), static_cast<const float *>(gamma), sums,
        static_cast<float *>(grad_gamma), static_cast<float *>(grad_beta), tokens);
    groupnorm_tokens_finalize_kernel<<<(n_stats + 127) / 128, 128, 0, st>>>(
        sums, c1, c2, n_stats, static_cast<double>(tokens) * (kGnC / kGnGroups), 0.f, 0);
    groupnorm_tokens_bwd_apply_kernel<<<grid, 256, 0, st>>>(
        static_cast<const float *>(grad_out), static_cast<long long>(grad_batch_stride), static_cast<const float *>(y),
        static_cast<const float *>(mean), static_cast<const float *>(rstd), c1, c2, static_cast<const float *>(gamma),
        static_cast<float *>(grad_y), tokens);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch groupnorm_tokens backward kernels");
    count_launch(3);
    return BM2F_OK;
}

int bm2f_sine_position_embedding(void *out, int height, int width, int num_pos_feats, float temperature, float scale,
                                 int normalize, void *stream)
{
    if (!out) return fail(BM2F_ERR_INVALID, "null pointer");
    if (height <= 0 || width <= 0 || num_pos_feats <= 0) return fail(BM2F_ERR_INVALID, "height / width / num_pos_feats must be positive");
    int sms = 0, cc = 0;
    const int rc = device_info(&sms, &cc);
    if (rc) return rc;
    const size_t total = static_cast<size_t>(height) * width * 2 * num_pos_feats;
    size_t blocks = (total + 255) / 256;
    if (blocks > static_cast<size_t>(sms) * 8) blocks = static_cast<size_t>(sms) * 8;
    sine_pos_embed_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float *>(out), height, width, num_pos_feats, temperature, scale, normalize);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch sine_pos_embed_kernel");
    count_launch(1);
    return BM2F_OK;
}

}  // extern "C"

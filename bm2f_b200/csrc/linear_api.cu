// C ABI of the tcgen05 projection GEMMs (linear_tf32x3.cuh): argument checks, tensor maps, launches.
#include "api_common.cuh"
#include "linear_tf32x3.cuh"

using namespace bm2f;
using namespace bm2f::host;

namespace {
template <int NT, int NH>
int launch_linear(const LinearParams &p, const float *w_hi, const float *w_lo, cudaStream_t st)
{
    constexpr int N = NT * NH;
    CUtensorMap mh, ml;
    int rc;
    if ((rc = make_map(&mh, w_hi, N, p.K, NT, kGemmBlockK, true))) return rc;
    if ((rc = make_map(&ml, w_lo, N, p.K, NT, kGemmBlockK, true))) return rc;
    CUtensorMap my;
    if ((rc = make_map(&my, p.y, p.M, N, kGemmBlockM, 32, true))) return rc;
    constexpr int smem = linear_smem_bytes<NT, NH>();
    if ((rc = ensure_dynamic_smem<&linear_tf32x3_kernel<NT, NH>>(smem, "cudaFuncSetAttribute(linear smem)"))) return rc;
    const int grid = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    linear_tf32x3_kernel<NT, NH><<<grid, kGemmThreads, smem, st>>>(p, mh, ml, my);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch linear_tf32x3_kernel");
    count_launch(1);
    return BM2F_OK;
}
template <int NT, int CL = 1, int PW = kGemmProducerWarps, int XD = 3>
int launch_linear_persistent(const LinearParams &p, const float *w_hi, const float *w_lo, int sms, cudaStream_t st)
{
    CUtensorMap mh, ml, my;
    int rc;
    // CL > 1: every CTA of the cluster fetches NT / CL weight rows per k-block and multicasts them
    if ((rc = make_map(&mh, w_hi, p.N, p.K, NT / CL, kGemmBlockK, true))) return rc;
    if ((rc = make_map(&ml, w_lo, p.N, p.K, NT / CL, kGemmBlockK, true))) return rc;
    if ((rc = make_map(&my, p.y, p.M, p.N, 32, 32, true))) return rc;      // one 32 x 32 box per epilogue warp
    constexpr int smem = linear_persistent_smem_bytes<NT>();
    if ((rc = ensure_dynamic_smem<&linear_tf32x3_persistent_kernel<NT, CL, PW, XD>>(smem, "cudaFuncSetAttribute(persistent linear smem)")))
        return rc;
    const int row_tiles = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int tiles = ((row_tiles + CL - 1) / CL) * p.slices * CL;          // CTAs that have work
    int grid = tiles < sms ? tiles : sms;
    grid -= grid % CL;
    if (CL == 1) {
        linear_tf32x3_persistent_kernel<NT, 1, PW, XD><<<grid, gemm_threads_persistent(PW), smem, st>>>(p, mh, ml, my, my);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(gemm_threads_persistent(PW));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const cudaError_t le = cudaLaunchKernelEx(&cfg, linear_tf32x3_persistent_kernel<NT, CL, PW, XD>, p, mh, ml, my, my);
        if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(linear_tf32x3_persistent_kernel, cluster)");
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch linear_tf32x3_persistent_kernel");
    count_launch(1);
    return BM2F_OK;
}
// single TF32 pass with the activation tile TMA-loaded into the MMA stage (no producer warps, four stages)
template <int NT>
int launch_linear_xtma(const LinearParams &p, const float *w_hi, int sms, cudaStream_t st)
{
    CUtensorMap mh, my, mx;
    int rc;
    if ((rc = make_map(&mh, w_hi, p.N, p.K, NT, kGemmBlockK, true))) return rc;
    if ((rc = make_map(&my, p.y, p.M, p.N, 32, 32, true))) return rc;
    if ((rc = make_map(&mx, p.x, p.M, p.K, kGemmBlockM, kGemmBlockK, true))) return rc;
    constexpr int smem = linear_xtma_smem_bytes<NT>();
    if ((rc = ensure_dynamic_smem<&linear_tf32x3_persistent_kernel<NT, 1, kGemmProducerWarps, 3, false, true>>(
             smem, "cudaFuncSetAttribute(xtma linear smem)")))
        return rc;
    const int tiles = ((p.M + kGemmBlockM - 1) / kGemmBlockM) * p.slices;
    const int grid = tiles < sms ? tiles : sms;
    linear_tf32x3_persistent_kernel<NT, 1, kGemmProducerWarps, 3, false, true>
        <<<grid, kGemmThreadsPersistent, smem, st>>>(p, mh, mh, my, mx);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch linear_tf32x3_persistent_kernel (TMA activations)");
    count_launch(1);
    return BM2F_OK;
}

template <int NT>
int launch_linear_pair(const LinearParams &p, const float *w_hi, const float *w_lo, int sms, cudaStream_t st)
{
    CUtensorMap mh, ml, my;
    int rc;
    if ((rc = make_map(&mh, w_hi, p.N, p.K, NT / 2, kGemmBlockK, true))) return rc;      // each CTA stages half the rows
    if ((rc = make_map(&ml, w_lo, p.N, p.K, NT / 2, kGemmBlockK, true))) return rc;
    if ((rc = make_map(&my, p.y, p.M, p.N, 32, 32, true))) return rc;
    constexpr int smem = linear_pair_smem_bytes<NT>();
    if ((rc = ensure_dynamic_smem<&linear_tf32x3_persistent_kernel<NT, 2, kGemmProducerWarps, 3, true>>(smem, "cudaFuncSetAttribute(pair linear smem)"))) return rc;
    const int row_tiles = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int tiles = ((row_tiles + 1) / 2) * p.slices * 2;
    int grid = tiles < sms ? tiles : sms;
    grid -= grid % 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreadsPersistent);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, linear_tf32x3_persistent_kernel<NT, 2, kGemmProducerWarps, 3, true>, p, mh, ml, my, my);
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(linear_tf32x3_persistent_kernel, CTA pair)");
    count_launch(1);
    return BM2F_OK;
}
}  // namespace
extern "C" {

size_t bm2f_linear_workspace_bytes(int out_features, int in_features)
{
    return static_cast<size_t>(2) * out_features * in_features * sizeof(float);
}

namespace {
bm2f_linear_tuning_t g_linear_tuning = {};      // A/B variants only (bm2f_linear_set_tuning)
}  // namespace

int bm2f_linear_set_tuning(const bm2f_linear_tuning_t *tuning)
{
    if (tuning) g_linear_tuning = *tuning;
    else memset(&g_linear_tuning, 0, sizeof(g_linear_tuning));
    return BM2F_OK;
}

namespace {
// y[rows, n_out] = x[rows, k_red] * w'[n_out, k_red]^T (+ bias); w' = weight or its transpose
int linear_common(const void *x, const void *weight, const void *bias, void *y, void *workspace, int rows, int n_out,
                  int k_red, int transpose_weight, int split, void *stream, int relu = 0, const void *mask = nullptr /* output mask */,
                  const void *addend = nullptr)
{
    if (!x || !weight || !y || !workspace) return fail(BM2F_ERR_INVALID, "null pointer");
    if (rows <= 0) return fail(BM2F_ERR_INVALID, "rows must be positive");
    if (k_red <= 0 || k_red % kGemmBlockK != 0 || k_red > kGemmKMax)
        return fail(BM2F_ERR_UNSUPPORTED, "tcgen05 projection GEMM needs a reduction length that is a multiple of %d "
                    "and <= %d (got %d)", kGemmBlockK, kGemmKMax, k_red);
    // kernel variant (bm2f_linear_tuning_t.variant, A/B only): 0 persistent kernel (double-buffered TMEM accumulator),
    // 1 one tile per CTA, 2 coalesced-store epilogue instead of the TMA store, 3 clusters of two CTAs with TMA-multicast
    // weights, 4 / 5 / 6 more activation bytes in flight (8 producer warps x 5 k-blocks / 4 x 4 / 8 x 4), 7 CTA pairs
    // issuing tcgen05.mma.cta_group::2 (M = 256 per pair).  split 1 (single TF32 pass) loads the activation tile by TMA
    // straight into the MMA stage by default; variant 5 selects the register-staged activation path for comparison
    const int variant = g_linear_tuning.variant;
    if (variant < 0 || variant > 8) return fail(BM2F_ERR_INVALID, "bm2f_linear_tuning_t.variant %d out of range (0..8)", variant);
    const int xvar = variant >= 4 ? variant - 3 : 0;
    const bool cluster2 = variant == 3, stg_epilogue = variant == 2, one_tile = variant == 1;
    if (split != 1 && split != 3) return fail(BM2F_ERR_INVALID, "split must be 3 (tf32x3) or 1 (single TF32 pass)");
    if (!aligned16(x) || !aligned16(y) || !aligned16(weight) || !aligned16(workspace) || (bias && !aligned16(bias)))
        return fail(BM2F_ERR_UNSUPPORTED, "linear: tensors must be 16-byte aligned");
    int sms = 0, cc = 0;
    int rc = device_info(&sms, &cc);
    if (rc) return rc;
    if (cc < 10) return fail(BM2F_ERR_CUDA, "this library contains sm_100a code only; device has cc %d.x", cc);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *w_hi = static_cast<float *>(workspace);
    float *w_lo = w_hi + static_cast<size_t>(n_out) * k_red;
    const int n = n_out * k_red;
    split_tf32_kernel<<<(n + 255) / 256, 256, 0, st>>>(static_cast<const float *>(weight), w_hi, w_lo, n, n_out, k_red,
                                                       transpose_weight);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch split_tf32_kernel");
    count_launch(1);
    LinearParams p{};
    p.x = static_cast<const float *>(x); p.bias = static_cast<const float *>(bias); p.y = static_cast<float *>(y);
    p.M = rows; p.N = n_out; p.K = k_red; p.slices = 1; p.relu = relu; p.out_mask = static_cast<const float *>(mask); p.store_mode = stg_epilogue ? 1 : 0;
    p.addend = static_cast<const float *>(addend);
    if (addend && !aligned16(addend)) return fail(BM2F_ERR_UNSUPPORTED, "linear: addend must be 16-byte aligned");
    p.split = split;
    if (mask && !aligned16(mask)) return fail(BM2F_ERR_UNSUPPORTED, "linear: mask must be 16-byte aligned");
    // single TF32 pass: the activation tile goes through TMA like the weights (no split to compute): default for split = 1
    const bool xtma = split == 1 && !one_tile && !stg_epilogue && !cluster2 && (xvar == 0 || xvar == 5);
    if (!one_tile) {
        if (n_out % 256 == 0) {      // 256-wide column slices (1024-wide FFN layer = 4 slices sharing the row tile)
            p.slices = n_out / 256;
            if (xtma) return launch_linear_xtma<256>(p, w_hi, sms, st);
            if (xvar == 4) return launch_linear_pair<256>(p, w_hi, w_lo, sms, st);
            if (xvar == 1) return launch_linear_persistent<256, 1, 8, 5>(p, w_hi, w_lo, sms, st);
            if (xvar == 2) return launch_linear_persistent<256, 1, 4, 4>(p, w_hi, w_lo, sms, st);
            if (xvar == 3) return launch_linear_persistent<256, 1, 8, 4>(p, w_hi, w_lo, sms, st);
            return cluster2 ? launch_linear_persistent<256, 2>(p, w_hi, w_lo, sms, st)
                            : launch_linear_persistent<256, 1>(p, w_hi, w_lo, sms, st);
        }
        switch (n_out) {
        case 192: if (xtma) return launch_linear_xtma<192>(p, w_hi, sms, st);
                  if (xvar == 4) return launch_linear_pair<192>(p, w_hi, w_lo, sms, st);
                  return cluster2 ? launch_linear_persistent<192, 2>(p, w_hi, w_lo, sms, st)
                                  : launch_linear_persistent<192, 1>(p, w_hi, w_lo, sms, st);
        case 96: if (xtma) return launch_linear_xtma<96>(p, w_hi, sms, st);
                 if (xvar == 4) return launch_linear_pair<96>(p, w_hi, w_lo, sms, st);
                 return cluster2 ? launch_linear_persistent<96, 2>(p, w_hi, w_lo, sms, st)
                                 : launch_linear_persistent<96, 1>(p, w_hi, w_lo, sms, st);
        case 288:            // offsets || logits: two 144-column slices of the same row tile run on neighbouring CTAs (the
                             // activation tile is read from HBM once, from L2 the second time); 144 = 4.5 x 32 columns:
                             // coalesced-store epilogue with a column predicate (a 32-wide TMA store box would spill
                             // into the other slice)
            if (!relu && !mask && !addend && xvar == 0 && !cluster2) {
                p.slices = 2;
                p.store_mode = 1;
                return launch_linear_persistent<144, 1>(p, w_hi, w_lo, sms, st);
            }
            break;
        default: break;
        }
    }
    if (relu || mask || addend)
        return fail(BM2F_ERR_UNSUPPORTED, "linear: relu / mask / addend need the persistent kernel (width %% 256 == 0, 192 or 96)");
    switch (n_out) {
    case 256: return launch_linear<256, 1>(p, w_hi, w_lo, st);
    case 288: return launch_linear<144, 2>(p, w_hi, w_lo, st);
    case 192: return launch_linear<192, 1>(p, w_hi, w_lo, st);
    case 96: return launch_linear<96, 1>(p, w_hi, w_lo, st);
    default:
        return fail(BM2F_ERR_UNSUPPORTED, "linear: output width %d not instantiated (256, 288, 192, 96)", n_out);
    }
}
}  // namespace

int bm2f_linear_forward(const void *x, const void *weight, const void *bias, void *y, void *workspace, int rows,
                        int out_features, int in_features, int split, void *stream)
{
    return linear_common(x, weight, bias, y, workspace, rows, out_features, in_features, 0, split, stream);
}

int bm2f_linear_relu_forward(const void *x, const void *weight, const void *bias, void *y, void *workspace, int rows,
                             int out_features, int in_features, int split, void *stream)
{
    return linear_common(x, weight, bias, y, workspace, rows, out_features, in_features, 0, split, stream, 1);
}

namespace {
int linear_dw_common(const void *grad_y, const void *x, void *grad_weight, void *grad_bias, int rows,
                     int out_features, int in_features, int split, void *stream)
{
    if (!grad_y || !x || !grad_weight) return fail(BM2F_ERR_INVALID, "null pointer");
    if (rows <= 0 || out_features <= 0) return fail(BM2F_ERR_INVALID, "rows / out_features must be positive");
    if (in_features <= 0 || in_features % 256 != 0)
        return fail(BM2F_ERR_UNSUPPORTED, "weight-gradient GEMM needs in_features to be a multiple of 256 (got %d)",
                    in_features);
    if (out_features > 8192 || in_features > 8192) return fail(BM2F_ERR_UNSUPPORTED, "layer too large");
    // A/B knob: cap the rows one CTA reduces at 256 * c, i.e. shorten the TMEM accumulation chain
    if (g_linear_tuning.dw_row_cap < 0) return fail(BM2F_ERR_INVALID, "bm2f_linear_tuning_t.dw_row_cap must be >= 0");
    const int row_cap = g_linear_tuning.dw_row_cap * 256;
    if (split != 1 && split != 3) return fail(BM2F_ERR_INVALID, "split must be 3 (tf32x3) or 1 (single TF32 pass)");
    int sms = 0, cc = 0;
    int rc = device_info(&sms, &cc);
    if (rc) return rc;
    if (cc < 10) return fail(BM2F_ERR_CUDA, "this library contains sm_100a code only; device has cc %d.x", cc);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(grad_weight, 0, static_cast<size_t>(out_features) * in_features * 4, st);
    if (e == cudaSuccess && grad_bias) e = cudaMemsetAsync(grad_bias, 0, static_cast<size_t>(out_features) * 4, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_weight / grad_bias)");
    if (split == 1 && out_features % 256 == 0 && rows < (1 << 30) && aligned16(grad_y) && aligned16(x)) {
        // single TF32 pass: both operands reach the tensor core through TMA as MN-major tiles (linear_dw_tma_kernel)
        const int units = in_features / 256, nsl = out_features / 256;
        int chunks = sms / (units * nsl);
        if (chunks < 1) chunks = 1;
        int rows_per_chunk = ((rows + chunks - 1) / chunks + 31) / 32 * 32;
        chunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
        DwTmaParams p{};
        p.dw = static_cast<float *>(grad_weight); p.ld_dw = in_features; p.units = units; p.rows = rows;
        p.rows_per_chunk = rows_per_chunk; p.conv_wp = 0;
        CUtensorMap mg, mx;
        if ((rc = make_map(&mg, static_cast<const float *>(grad_y), rows, out_features, 32, 32, 2))) return rc;
        if ((rc = make_map(&mx, static_cast<const float *>(x), rows, in_features, 32, 32, 2))) return rc;
        constexpr int smem = linear_dw_tma_smem_bytes();
        if ((rc = ensure_dynamic_smem<&linear_dw_tma_kernel>(smem, "cudaFuncSetAttribute(dW TMA smem)"))) return rc;
        linear_dw_tma_kernel<<<dim3(units * nsl, chunks), kDwTmaThreads, smem, st>>>(p, mg, mx);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "launch linear_dw_tma_kernel");
        count_launch(1);
        if (grad_bias) {
            int ctas = sms * 2;
            const int rows_per_cta = (rows + ctas - 1) / ctas;
            ctas = (rows + rows_per_cta - 1) / rows_per_cta;
            column_sum_kernel<<<dim3(ctas, (out_features + 255) / 256), 256, 0, st>>>(
                static_cast<const float *>(grad_y), static_cast<float *>(grad_bias), rows, out_features, rows_per_cta);
            e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(e, "launch column_sum_kernel");
            count_launch(1);
        }
        return BM2F_OK;
    }
    const int n_tiles = (out_features + 127) / 128;
    const int k_slices = in_features / 256;
    int chunks = sms / (n_tiles * k_slices);
    if (chunks < 1) chunks = 1;
    if (row_cap > 0 && (rows + chunks - 1) / chunks > row_cap) chunks = (rows + row_cap - 1) / row_cap;
    int rows_per_chunk = ((rows + chunks - 1) / chunks + 31) / 32 * 32;
    chunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
    LinearDwParams p{};
    p.g = static_cast<const float *>(grad_y); p.x = static_cast<const float *>(x);
    p.dw = static_cast<float *>(grad_weight); p.db = static_cast<float *>(grad_bias);
    p.M = rows; p.N = out_features; p.ldx = in_features; p.rows_per_chunk = rows_per_chunk; p.split = split;
    constexpr int smem = linear_dw_smem_bytes();
    if ((rc = ensure_dynamic_smem<&linear_dw_tf32x3_kernel>(smem, "cudaFuncSetAttribute(dW smem)"))) return rc;
    linear_dw_tf32x3_kernel<<<dim3(n_tiles, chunks, k_slices), kDwThreads, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch linear_dw_tf32x3_kernel");
    count_launch(1);
    return BM2F_OK;
}
}  // namespace

int bm2f_linear_backward_weight(const void *grad_y, const void *x, void *grad_weight, void *grad_bias, int rows,
                                int out_features, int in_features, int split, void *stream)
{
    return linear_dw_common(grad_y, x, grad_weight, grad_bias, rows, out_features, in_features, split, stream);
}

int bm2f_linear_backward_input(const void *grad_y, const void *weight, void *grad_x, void *workspace, int rows,
                               int out_features, int in_features, int split, void *stream)
{
    // grad_x[rows, in] = grad_y[rows, out] * weight[out, in]: a GEMM over k = out with w' = weight^T (in, out)
    return linear_common(grad_y, weight, nullptr, grad_x, workspace, rows, in_features, out_features, 1, split, stream);
}

int bm2f_linear_backward_input_accumulate(const void *grad_y, const void *weight, const void *addend, void *grad_x,
                                          void *workspace, int rows, int out_features, int in_features, int split,
                                          void *stream)
{
    // grad_x = grad_y * weight + addend in the GEMM epilogue: gradient branches that meet at one tensor are summed
    // without a separate element-wise pass.  addend == grad_x accumulates in place.
    if (!addend) return fail(BM2F_ERR_INVALID, "null addend");
    return linear_common(grad_y, weight, nullptr, grad_x, workspace, rows, in_features, out_features, 1, split, stream, 0,
                         nullptr, addend);
}

int bm2f_linear_backward_input_masked(const void *grad_y, const void *weight, const void *mask, void *grad_x,
                                      void *workspace, int rows, int out_features, int in_features, int split,
                                      void *stream)
{
    // grad_x = (grad_y * weight) where mask > 0, else 0: the ReLU backward of the layer that produced this layer's
    // input is applied in the GEMM epilogue, so the masked gradient is produced in one pass
    if (!mask) return fail(BM2F_ERR_INVALID, "null mask");
    return linear_common(grad_y, weight, nullptr, grad_x, workspace, rows, in_features, out_features, 1, split, stream, 0,
                         mask);
}

}  // extern "C"

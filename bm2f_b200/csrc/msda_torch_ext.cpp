// Python extension module "MultiScaleDeformableAttention": the reference's binding surface
// (ops/src/vision.cpp:18-21) on top of the C ABI in include/bm2f_msda.h.
//
//   ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
//                          im2col_step) -> Tensor[N, Lq, M*D]
//   ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
//                           grad_output, im2col_step) -> [grad_value, grad_sampling_loc, grad_attn_weight]
//
// Same positional signatures, same preconditions and error type (RuntimeError) as
// ms_deform_attn_cuda_forward/backward (ops/src/cuda/ms_deform_attn_cuda.cu:25-158) and the
// CPU/CUDA dispatch in ops/src/ms_deform_attn.h:25-66.  This file is plumbing only: it checks
// tensors, allocates the results with the caller's options (so torch's caching allocator and
// stream semantics hold), picks the current stream and device, and forwards raw pointers.
// PyTorch appears nowhere below this file.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <vector>

#include "../../include/bm2f_msda.h"

namespace {

int dtype_code(const at::Tensor &value)
{
    switch (value.scalar_type()) {
    case at::kFloat: return BM2F_DTYPE_F32;
    case at::kDouble: return BM2F_DTYPE_F64;
    case at::kBFloat16: return BM2F_DTYPE_BF16;
    default:
        TORCH_CHECK(false, "ms_deform_attn: unsupported dtype ", value.scalar_type(),
                    " (float32, float64 and bfloat16 are implemented)");
    }
    return -1;
}

struct Checked {
    at::Tensor loc, attn;  // possibly converted (bf16 value path keeps float32 locations/weights)
    int N, S, M, D, L, Lq, P, dtype;
};

Checked check_inputs(const at::Tensor &value, const at::Tensor &spatial_shapes, const at::Tensor &level_start_index,
                     const at::Tensor &sampling_loc, const at::Tensor &attn_weight, int im2col_step)
{
    // reference: ms_deform_attn.h:40-43 / 62-65
    TORCH_CHECK(value.is_cuda(), "Not implemented on the CPU");
    // reference: ms_deform_attn_cuda.cu:33-43
    TORCH_CHECK(value.is_contiguous(), "value tensor has to be contiguous");
    TORCH_CHECK(spatial_shapes.is_contiguous(), "spatial_shapes tensor has to be contiguous");
    TORCH_CHECK(level_start_index.is_contiguous(), "level_start_index tensor has to be contiguous");
    TORCH_CHECK(sampling_loc.is_contiguous(), "sampling_loc tensor has to be contiguous");
    TORCH_CHECK(attn_weight.is_contiguous(), "attn_weight tensor has to be contiguous");
    TORCH_CHECK(spatial_shapes.is_cuda(), "spatial_shapes must be a CUDA tensor");
    TORCH_CHECK(level_start_index.is_cuda(), "level_start_index must be a CUDA tensor");
    TORCH_CHECK(sampling_loc.is_cuda(), "sampling_loc must be a CUDA tensor");
    TORCH_CHECK(attn_weight.is_cuda(), "attn_weight must be a CUDA tensor");
    // the reference reinterprets both tables as int64 (.cu:72-73); make that explicit
    TORCH_CHECK(spatial_shapes.scalar_type() == at::kLong, "spatial_shapes must be int64");
    TORCH_CHECK(level_start_index.scalar_type() == at::kLong, "level_start_index must be int64");
    TORCH_CHECK(value.dim() == 4, "value must be (N, S, M, D)");
    TORCH_CHECK(sampling_loc.dim() == 6 && sampling_loc.size(5) == 2, "sampling_loc must be (N, Lq, M, L, P, 2)");
    TORCH_CHECK(attn_weight.dim() == 5, "attn_weight must be (N, Lq, M, L, P)");
    TORCH_CHECK(spatial_shapes.dim() == 2 && spatial_shapes.size(1) == 2, "spatial_shapes must be (L, 2)");

    Checked c;
    c.N = static_cast<int>(value.size(0));
    c.S = static_cast<int>(value.size(1));
    c.M = static_cast<int>(value.size(2));
    c.D = static_cast<int>(value.size(3));
    c.L = static_cast<int>(spatial_shapes.size(0));
    c.Lq = static_cast<int>(sampling_loc.size(1));
    c.P = static_cast<int>(sampling_loc.size(4));
    TORCH_CHECK(level_start_index.numel() == c.L, "level_start_index must have one entry per level");
    TORCH_CHECK(sampling_loc.size(0) == c.N && sampling_loc.size(2) == c.M && sampling_loc.size(3) == c.L,
                "sampling_loc shape does not match value / spatial_shapes");
    TORCH_CHECK(attn_weight.size(0) == c.N && attn_weight.size(1) == c.Lq && attn_weight.size(2) == c.M &&
                    attn_weight.size(3) == c.L && attn_weight.size(4) == c.P,
                "attn_weight shape does not match sampling_loc");
    // reference: .cu:53-57
    TORCH_CHECK(bm2f_msda_check_im2col_step(c.N, im2col_step) == BM2F_OK, bm2f_msda_last_error());

    c.dtype = dtype_code(value);
    if (c.dtype == BM2F_DTYPE_BF16) {
        // bf16 values, fp32 sampling locations / weights (see include/bm2f_msda.h)
        c.loc = sampling_loc.scalar_type() == at::kFloat ? sampling_loc : sampling_loc.to(at::kFloat);
        c.attn = attn_weight.scalar_type() == at::kFloat ? attn_weight : attn_weight.to(at::kFloat);
    } else {
        TORCH_CHECK(sampling_loc.scalar_type() == value.scalar_type() &&
                        attn_weight.scalar_type() == value.scalar_type(),
                    "sampling_loc / attn_weight must have the dtype of value");
        c.loc = sampling_loc;
        c.attn = attn_weight;
    }
    return c;
}

at::Tensor ms_deform_attn_forward(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                  const at::Tensor &level_start_index, const at::Tensor &sampling_loc,
                                  const at::Tensor &attn_weight, const int im2col_step)
{
    const Checked c = check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step);
    const c10::cuda::CUDAGuard guard(value.device());
    auto output = at::empty({c.N, c.Lq, c.M * c.D}, value.options());  // every element is written
    const int rc = bm2f_msda_forward(value.data_ptr(), spatial_shapes.data_ptr<int64_t>(),
                                     level_start_index.data_ptr<int64_t>(), c.loc.data_ptr(), c.attn.data_ptr(),
                                     output.data_ptr(), c.N, c.S, c.M, c.D, c.L, c.Lq, c.P, c.dtype, nullptr,
                                     at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "ms_deform_attn_forward: ", bm2f_msda_last_error());
    return output;
}

std::vector<at::Tensor> ms_deform_attn_backward(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                                const at::Tensor &level_start_index, const at::Tensor &sampling_loc,
                                                const at::Tensor &attn_weight, const at::Tensor &grad_output,
                                                const int im2col_step)
{
    const Checked c = check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step);
    TORCH_CHECK(grad_output.is_contiguous(), "grad_output tensor has to be contiguous");
    TORCH_CHECK(grad_output.is_cuda(), "grad_output must be a CUDA tensor");
    TORCH_CHECK(grad_output.scalar_type() == value.scalar_type(), "grad_output must have the dtype of value");
    TORCH_CHECK(grad_output.numel() == static_cast<int64_t>(c.N) * c.Lq * c.M * c.D,
                "grad_output must be (N, Lq, M*D)");
    const c10::cuda::CUDAGuard guard(value.device());
    // zero-filled by the library; bf16 values accumulate their gradient in fp32 (include/bm2f_msda.h)
    auto grad_value = c.dtype == BM2F_DTYPE_BF16 ? at::empty(value.sizes(), value.options().dtype(at::kFloat))
                                                 : at::empty_like(value);
    auto grad_loc = at::empty_like(c.loc);
    auto grad_attn = at::empty_like(c.attn);
    const int rc = bm2f_msda_backward(value.data_ptr(), spatial_shapes.data_ptr<int64_t>(),
                                      level_start_index.data_ptr<int64_t>(), c.loc.data_ptr(), c.attn.data_ptr(),
                                      grad_output.data_ptr(), grad_value.data_ptr(), grad_loc.data_ptr(),
                                      grad_attn.data_ptr(), c.N, c.S, c.M, c.D, c.L, c.Lq, c.P, c.dtype, nullptr,
                                      at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "ms_deform_attn_backward: ", bm2f_msda_last_error());
    if (grad_value.scalar_type() != value.scalar_type()) grad_value = grad_value.to(value.scalar_type());
    if (grad_loc.scalar_type() != sampling_loc.scalar_type()) grad_loc = grad_loc.to(sampling_loc.scalar_type());
    if (grad_attn.scalar_type() != attn_weight.scalar_type()) grad_attn = grad_attn.to(attn_weight.scalar_type());
    return {grad_value, grad_loc, grad_attn};
}


// ---- fused prologue variants (not part of the reference surface; used by bm2f_b200.ops.modules) ----
struct FusedChecked {
    int N, S, M, D, L, Lq, P, dtype;
};

// reference_points: (N,Lq,L,2) tensor, or None = the encoder's pixel-centre reference points, computed in the kernel
FusedChecked check_fused(const at::Tensor &value, const at::Tensor &spatial_shapes, const at::Tensor &level_start_index,
                         const c10::optional<at::Tensor> &reference_points_opt, const at::Tensor &sampling_offsets,
                         const at::Tensor &attn_logits)
{
    TORCH_CHECK(value.is_cuda(), "Not implemented on the CPU");
    for (const at::Tensor *t : {&value, &spatial_shapes, &level_start_index, &sampling_offsets, &attn_logits}) {
        TORCH_CHECK(t->is_cuda(), "all tensors must be CUDA tensors");
        TORCH_CHECK(t->is_contiguous(), "all tensors have to be contiguous");
    }
    TORCH_CHECK(spatial_shapes.scalar_type() == at::kLong && level_start_index.scalar_type() == at::kLong,
                "spatial_shapes / level_start_index must be int64");
    TORCH_CHECK(value.dim() == 4 && sampling_offsets.dim() == 6 && sampling_offsets.size(5) == 2,
                "value must be (N,S,M,D), sampling_offsets (N,Lq,M,L,P,2)");
    TORCH_CHECK(sampling_offsets.scalar_type() == at::kFloat && attn_logits.scalar_type() == at::kFloat,
                "sampling_offsets / attn_logits must be float32");
    FusedChecked c;
    c.N = static_cast<int>(value.size(0)); c.S = static_cast<int>(value.size(1));
    c.M = static_cast<int>(value.size(2)); c.D = static_cast<int>(value.size(3));
    c.L = static_cast<int>(spatial_shapes.size(0)); c.Lq = static_cast<int>(sampling_offsets.size(1));
    c.P = static_cast<int>(sampling_offsets.size(4));
    TORCH_CHECK(sampling_offsets.size(0) == c.N && sampling_offsets.size(2) == c.M && sampling_offsets.size(3) == c.L,
                "sampling_offsets shape does not match value / spatial_shapes");
    TORCH_CHECK(attn_logits.numel() == static_cast<int64_t>(c.N) * c.Lq * c.M * c.L * c.P,
                "attn_logits must be (N,Lq,M,L*P)");
    if (reference_points_opt.has_value()) {
        const at::Tensor &reference_points = *reference_points_opt;
        TORCH_CHECK(reference_points.is_cuda() && reference_points.is_contiguous() &&
                        reference_points.scalar_type() == at::kFloat,
                    "reference_points must be a contiguous float32 CUDA tensor");
        TORCH_CHECK(reference_points.dim() == 4 && reference_points.size(3) == 2,
                    "fused path takes 2-d reference points (N,Lq,L,2)");
        TORCH_CHECK(reference_points.size(0) == c.N && reference_points.size(1) == c.Lq && reference_points.size(2) == c.L,
                    "reference_points must be (N,Lq,L,2)");
    } else {
        TORCH_CHECK(c.Lq == c.S, "reference_points=None (pixel-centre reference points) needs Lq == S");
    }
    c.dtype = dtype_code(value);
    return c;
}

bool ms_deform_attn_fused_supported(int64_t num_heads, int64_t channels, int64_t num_levels, int64_t num_point,
                                    bool bf16)
{
    return bm2f_msda_fused_supported(static_cast<int>(num_heads), static_cast<int>(channels),
                                     static_cast<int>(num_levels), static_cast<int>(num_point),
                                     bf16 ? BM2F_DTYPE_BF16 : BM2F_DTYPE_F32) != 0;
}

at::Tensor ms_deform_attn_fused_forward(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                        const at::Tensor &level_start_index,
                                        const c10::optional<at::Tensor> &reference_points,
                                        const at::Tensor &sampling_offsets, const at::Tensor &attn_logits)
{
    const FusedChecked c = check_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                       attn_logits);
    const c10::cuda::CUDAGuard guard(value.device());
    auto output = at::empty({c.N, c.Lq, c.M * c.D}, value.options());
    const int rc = bm2f_msda_fused_forward(value.data_ptr(), spatial_shapes.data_ptr<int64_t>(),
                                           level_start_index.data_ptr<int64_t>(),
                                           reference_points.has_value() ? reference_points->data_ptr() : nullptr,
                                           sampling_offsets.data_ptr(), attn_logits.data_ptr(), output.data_ptr(),
                                           c.N, c.S, c.M, c.D, c.L, c.Lq, c.P, c.dtype, nullptr,
                                           at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "ms_deform_attn_fused_forward: ", bm2f_msda_last_error());
    return output;
}

std::vector<at::Tensor> ms_deform_attn_fused_backward(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                                      const at::Tensor &level_start_index,
                                                      const c10::optional<at::Tensor> &reference_points,
                                                      const at::Tensor &sampling_offsets,
                                                      const at::Tensor &attn_logits, const at::Tensor &grad_output)
{
    const FusedChecked c = check_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                       attn_logits);
    TORCH_CHECK(grad_output.is_cuda() && grad_output.is_contiguous(), "grad_output must be a contiguous CUDA tensor");
    TORCH_CHECK(grad_output.scalar_type() == value.scalar_type(), "grad_output must have the dtype of value");
    const c10::cuda::CUDAGuard guard(value.device());
    auto grad_value = at::empty(value.sizes(), value.options().dtype(at::kFloat));
    auto grad_off = at::empty_like(sampling_offsets);
    auto grad_logits = at::empty_like(attn_logits);
    const int rc = bm2f_msda_fused_backward(value.data_ptr(), spatial_shapes.data_ptr<int64_t>(),
                                            level_start_index.data_ptr<int64_t>(),
                                           reference_points.has_value() ? reference_points->data_ptr() : nullptr,
                                            sampling_offsets.data_ptr(), attn_logits.data_ptr(), grad_output.data_ptr(),
                                            grad_value.data_ptr(), grad_off.data_ptr(), grad_logits.data_ptr(), c.N,
                                            c.S, c.M, c.D, c.L, c.Lq, c.P, c.dtype, nullptr,
                                            at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "ms_deform_attn_fused_backward: ", bm2f_msda_last_error());
    if (grad_value.scalar_type() != value.scalar_type()) grad_value = grad_value.to(value.scalar_type());
    return {grad_value, grad_off, grad_logits};
}


// ---- packed fused op: offsets_logits (N, Lq, M*L*P*3) = [offsets | logits] per row, the output of ONE projection ----
namespace {
FusedChecked check_fused_packed(const at::Tensor &value, const at::Tensor &spatial_shapes, const at::Tensor &level_start_index,
                                const c10::optional<at::Tensor> &reference_points_opt, const at::Tensor &oa, int64_t num_point)
{
    TORCH_CHECK(value.is_cuda(), "Not implemented on the CPU");
    for (const at::Tensor *t : {&value, &spatial_shapes, &level_start_index, &oa}) {
        TORCH_CHECK(t->is_cuda(), "all tensors must be CUDA tensors");
        TORCH_CHECK(t->is_contiguous(), "all tensors have to be contiguous");
    }
    TORCH_CHECK(spatial_shapes.scalar_type() == at::kLong && level_start_index.scalar_type() == at::kLong,
                "spatial_shapes / level_start_index must be int64");
    TORCH_CHECK(value.dim() == 4 && oa.dim() == 3 && oa.scalar_type() == at::kFloat,
                "value must be (N,S,M,D), offsets_logits float32 (N,Lq,M*L*P*3)");
    FusedChecked c;
    c.N = static_cast<int>(value.size(0)); c.S = static_cast<int>(value.size(1));
    c.M = static_cast<int>(value.size(2)); c.D = static_cast<int>(value.size(3));
    c.L = static_cast<int>(spatial_shapes.size(0)); c.Lq = static_cast<int>(oa.size(1));
    c.P = static_cast<int>(num_point);
    TORCH_CHECK(oa.size(0) == c.N && oa.size(2) == static_cast<int64_t>(c.M) * c.L * c.P * 3,
                "offsets_logits must be (N,Lq,M*L*P*3)");
    if (reference_points_opt.has_value()) {
        const at::Tensor &r = *reference_points_opt;
        TORCH_CHECK(r.is_cuda() && r.is_contiguous() && r.scalar_type() == at::kFloat && r.dim() == 4 && r.size(3) == 2 &&
                        r.size(0) == c.N && r.size(1) == c.Lq && r.size(2) == c.L,
                    "reference_points must be a contiguous float32 CUDA tensor (N,Lq,L,2)");
    } else {
        TORCH_CHECK(c.Lq == c.S, "reference_points=None (pixel-centre reference points) needs Lq == S");
    }
    c.dtype = dtype_code(value);
    return c;
}
}  // namespace

at::Tensor ms_deform_attn_fused_forward_packed(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                               const at::Tensor &level_start_index,
                                               const c10::optional<at::Tensor> &reference_points,
                                               const at::Tensor &offsets_logits, int64_t num_point)
{
    const FusedChecked c = check_fused_packed(value, spatial_shapes, level_start_index, reference_points, offsets_logits, num_point);
    const c10::cuda::CUDAGuard guard(value.device());
    auto output = at::empty({c.N, c.Lq, c.M * c.D}, value.options());
    const int rc = bm2f_msda_fused_forward_packed(value.data_ptr(), spatial_shapes.data_ptr<int64_t>(),
                                                  level_start_index.data_ptr<int64_t>(),
                                                  reference_points.has_value() ? reference_points->data_ptr() : nullptr,
                                                  offsets_logits.data_ptr(), output.data_ptr(), c.N, c.S, c.M, c.D, c.L, c.Lq,
                                                  c.P, c.dtype, nullptr, at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "ms_deform_attn_fused_forward_packed: ", bm2f_msda_last_error());
    return output;
}

std::vector<at::Tensor> ms_deform_attn_fused_backward_packed(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                                             const at::Tensor &level_start_index,
                                                             const c10::optional<at::Tensor> &reference_points,
                                                             const at::Tensor &offsets_logits, int64_t num_point,
                                                             const at::Tensor &grad_output)
{
    const FusedChecked c = check_fused_packed(value, spatial_shapes, level_start_index, reference_points, offsets_logits, num_point);
    TORCH_CHECK(grad_output.is_cuda() && grad_output.is_contiguous(), "grad_output must be a contiguous CUDA tensor");
    TORCH_CHECK(grad_output.scalar_type() == value.scalar_type(), "grad_output must have the dtype of value");
    const c10::cuda::CUDAGuard guard(value.device());
    auto grad_value = at::empty(value.sizes(), value.options().dtype(at::kFloat));
    auto grad_oa = at::empty_like(offsets_logits);
    const int rc = bm2f_msda_fused_backward_packed(value.data_ptr(), spatial_shapes.data_ptr<int64_t>(),
                                                   level_start_index.data_ptr<int64_t>(),
                                                   reference_points.has_value() ? reference_points->data_ptr() : nullptr,
                                                   offsets_logits.data_ptr(), grad_output.data_ptr(), grad_value.data_ptr(),
                                                   grad_oa.data_ptr(), c.N, c.S, c.M, c.D, c.L, c.Lq, c.P, c.dtype, nullptr,
                                                   at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "ms_deform_attn_fused_backward_packed: ", bm2f_msda_last_error());
    if (grad_value.scalar_type() != value.scalar_type()) grad_value = grad_value.to(value.scalar_type());
    return {grad_value, grad_oa};
}

// ---- tcgen05 projection GEMM (forward of nn.Linear, fp32) -------------------------------------------------
bool linear_tf32x3_supported(int64_t in_features, int64_t out_features)
{
    const bool width_ok = out_features % 256 == 0 || out_features == 288 || out_features == 192 || out_features == 96;
    return in_features % 256 == 0 && in_features <= 4096 && width_ok && out_features <= 4096;
}

// (grad_weight, grad_bias) = (grad_y^T x, column sums of grad_y) on the tensor cores
std::vector<at::Tensor> linear_tf32x3_backward_weight(const at::Tensor &grad_y, const at::Tensor &x, int64_t split,
                                                      bool with_bias)
{
    TORCH_CHECK(grad_y.is_cuda() && x.is_cuda(), "linear_tf32x3_backward_weight: CUDA tensors only");
    TORCH_CHECK(grad_y.scalar_type() == at::kFloat && x.scalar_type() == at::kFloat, "float32 only");
    const c10::cuda::CUDAGuard guard(x.device());
    auto g = grad_y.contiguous();
    auto xc = x.contiguous();
    const int64_t k = xc.size(-1);
    TORCH_CHECK(k % 256 == 0, "in_features must be a multiple of 256");
    const int64_t rows = xc.numel() / k;
    const int64_t n = g.size(-1);
    TORCH_CHECK(g.numel() == rows * n, "grad_y / x row count mismatch");
    auto gw = at::empty({n, k}, xc.options());
    auto gb = with_bias ? at::empty({n}, xc.options()) : at::Tensor();
    const int rc = bm2f_linear_backward_weight(g.data_ptr(), xc.data_ptr(), gw.data_ptr(),
                                               with_bias ? gb.data_ptr() : nullptr, static_cast<int>(rows),
                                               static_cast<int>(n), static_cast<int>(k), static_cast<int>(split),
                                               at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "linear_tf32x3_backward_weight: ", bm2f_msda_last_error());
    return {gw, gb};
}

// grad_x = grad_y @ weight on the tensor cores (reduction over out_features, output width in_features = 256)
at::Tensor linear_tf32x3_backward_input(const at::Tensor &grad_y, const at::Tensor &weight, int64_t split)
{
    TORCH_CHECK(grad_y.is_cuda() && weight.is_cuda(), "linear_tf32x3_backward_input: CUDA tensors only");
    TORCH_CHECK(grad_y.scalar_type() == at::kFloat && weight.scalar_type() == at::kFloat, "float32 only");
    TORCH_CHECK(weight.dim() == 2 && grad_y.size(-1) == weight.size(0), "shape mismatch");
    TORCH_CHECK(linear_tf32x3_supported(weight.size(1), weight.size(0)), "unsupported layer shape");
    const c10::cuda::CUDAGuard guard(grad_y.device());
    auto g = grad_y.contiguous();
    auto wc = weight.contiguous();
    const int64_t rows = g.numel() / g.size(-1);
    auto sizes = g.sizes().vec();
    sizes.back() = wc.size(1);
    auto gx = at::empty(sizes, g.options());
    auto ws = at::empty({static_cast<int64_t>(bm2f_linear_workspace_bytes(wc.size(0), wc.size(1)) / 4)}, g.options());
    const int rc = bm2f_linear_backward_input(g.data_ptr(), wc.data_ptr(), gx.data_ptr(), ws.data_ptr(),
                                              static_cast<int>(rows), static_cast<int>(wc.size(0)),
                                              static_cast<int>(wc.size(1)), static_cast<int>(split),
                                              at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "linear_tf32x3_backward_input: ", bm2f_msda_last_error());
    return gx;
}

// grad_x = grad_y @ weight + addend (epilogue accumulation); `out` may be the addend itself (in place) or a new tensor
at::Tensor linear_tf32x3_backward_input_accumulate(const at::Tensor &grad_y, const at::Tensor &weight,
                                                   const at::Tensor &addend, bool in_place, int64_t split)
{
    TORCH_CHECK(grad_y.is_cuda() && weight.is_cuda() && addend.is_cuda(), "CUDA tensors only");
    TORCH_CHECK(grad_y.scalar_type() == at::kFloat && weight.scalar_type() == at::kFloat && addend.scalar_type() == at::kFloat,
                "float32 only");
    TORCH_CHECK(weight.dim() == 2 && grad_y.size(-1) == weight.size(0) && addend.size(-1) == weight.size(1), "shape mismatch");
    TORCH_CHECK(linear_tf32x3_supported(weight.size(1), weight.size(0)), "unsupported layer shape");
    TORCH_CHECK(addend.is_contiguous(), "addend must be contiguous");
    const c10::cuda::CUDAGuard guard(grad_y.device());
    auto g = grad_y.contiguous();
    auto wc = weight.contiguous();
    const int64_t rows = g.numel() / g.size(-1);
    TORCH_CHECK(addend.numel() == rows * wc.size(1), "addend must have one row per grad_y row");
    auto out = in_place ? addend : at::empty_like(addend);
    auto ws = at::empty({static_cast<int64_t>(bm2f_linear_workspace_bytes(wc.size(0), wc.size(1)) / 4)}, g.options());
    const int rc = bm2f_linear_backward_input_accumulate(g.data_ptr(), wc.data_ptr(), addend.data_ptr(), out.data_ptr(),
                                                         ws.data_ptr(), static_cast<int>(rows), static_cast<int>(wc.size(0)),
                                                         static_cast<int>(wc.size(1)), static_cast<int>(split),
                                                         at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "linear_tf32x3_backward_input_accumulate: ", bm2f_msda_last_error());
    return out;
}

at::Tensor linear_tf32x3(const at::Tensor &x, const at::Tensor &weight, const c10::optional<at::Tensor> &bias,
                         int64_t split)
{
    TORCH_CHECK(x.is_cuda() && weight.is_cuda(), "linear_tf32x3: CUDA tensors only (no CPU path)");
    TORCH_CHECK(x.scalar_type() == at::kFloat && weight.scalar_type() == at::kFloat, "linear_tf32x3: float32 only");
    TORCH_CHECK(weight.dim() == 2 && x.size(-1) == weight.size(1), "linear_tf32x3: shape mismatch");
    TORCH_CHECK(linear_tf32x3_supported(weight.size(1), weight.size(0)), "linear_tf32x3: unsupported layer shape ",
                weight.size(1), " -> ", weight.size(0));
    const c10::cuda::CUDAGuard guard(x.device());
    auto xc = x.contiguous();
    auto wc = weight.contiguous();
    const int64_t rows = xc.numel() / xc.size(-1);
    auto sizes = xc.sizes().vec();
    sizes.back() = wc.size(0);
    auto y = at::empty(sizes, xc.options());
    auto ws = at::empty({static_cast<int64_t>(bm2f_linear_workspace_bytes(wc.size(0), wc.size(1)) / 4)}, xc.options());
    const void *b = nullptr;
    at::Tensor bc;
    if (bias.has_value() && bias->defined()) {
        bc = bias->contiguous();
        TORCH_CHECK(bc.is_cuda() && bc.scalar_type() == at::kFloat && bc.numel() == wc.size(0), "linear_tf32x3: bad bias");
        b = bc.data_ptr();
    }
    const int rc = bm2f_linear_forward(xc.data_ptr(), wc.data_ptr(), b, y.data_ptr(), ws.data_ptr(),
                                       static_cast<int>(rows), static_cast<int>(wc.size(0)),
                                       static_cast<int>(wc.size(1)), static_cast<int>(split),
                                       at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "linear_tf32x3: ", bm2f_msda_last_error());
    return y;
}


// ---- FFN layer with fused ReLU and fused residual + LayerNorm (encoder layer, msdeformattn.py:92-131) -----
at::Tensor linear_relu_tf32x3(const at::Tensor &x, const at::Tensor &weight, const at::Tensor &bias, int64_t split)
{
    TORCH_CHECK(x.is_cuda() && weight.is_cuda() && bias.is_cuda(), "linear_relu_tf32x3: CUDA tensors only");
    TORCH_CHECK(x.scalar_type() == at::kFloat && weight.scalar_type() == at::kFloat, "float32 only");
    TORCH_CHECK(weight.dim() == 2 && x.size(-1) == weight.size(1), "shape mismatch");
    const c10::cuda::CUDAGuard guard(x.device());
    auto xc = x.contiguous();
    auto wc = weight.contiguous();
    auto bc = bias.contiguous();
    const int64_t rows = xc.numel() / xc.size(-1);
    auto sizes = xc.sizes().vec();
    sizes.back() = wc.size(0);
    auto y = at::empty(sizes, xc.options());
    auto ws = at::empty({static_cast<int64_t>(bm2f_linear_workspace_bytes(wc.size(0), wc.size(1)) / 4)}, xc.options());
    const int rc = bm2f_linear_relu_forward(xc.data_ptr(), wc.data_ptr(), bc.data_ptr(), y.data_ptr(), ws.data_ptr(),
                                            static_cast<int>(rows), static_cast<int>(wc.size(0)),
                                            static_cast<int>(wc.size(1)), static_cast<int>(split),
                                            at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "linear_relu_tf32x3: ", bm2f_msda_last_error());
    return y;
}

// backward of y = linear2(relu(linear1(x))) given h = relu(linear1(x)):
// returns (grad_x, grad_W1, grad_b1, grad_W2, grad_b2); every GEMM on tcgen05, the ReLU mask applied in an epilogue
std::vector<at::Tensor> ffn_tf32x3_backward(const at::Tensor &grad_y, const at::Tensor &x, const at::Tensor &h,
                                            const at::Tensor &w1, const at::Tensor &w2, int64_t split)
{
    TORCH_CHECK(grad_y.is_cuda() && x.is_cuda() && h.is_cuda() && w1.is_cuda() && w2.is_cuda(), "CUDA tensors only");
    const c10::cuda::CUDAGuard guard(x.device());
    auto g = grad_y.contiguous();
    auto xc = x.contiguous();
    auto hc = h.contiguous();
    auto w1c = w1.contiguous();
    auto w2c = w2.contiguous();
    const int d = static_cast<int>(w1c.size(1)), f = static_cast<int>(w1c.size(0));      // d_model, d_ffn
    TORCH_CHECK(w2c.size(0) == d && w2c.size(1) == f, "linear2 must be (d_model, d_ffn)");
    const int64_t rows = xc.numel() / d;
    auto stream = at::cuda::getCurrentCUDAStream().stream();
    auto ws = at::empty({static_cast<int64_t>(bm2f_linear_workspace_bytes(f, d) / 4)}, xc.options());
    auto g1 = at::empty_like(hc);                                   // grad of the hidden activation, ReLU-masked
    int rc = bm2f_linear_backward_input_masked(g.data_ptr(), w2c.data_ptr(), hc.data_ptr(), g1.data_ptr(), ws.data_ptr(),
                                               static_cast<int>(rows), d, f, static_cast<int>(split), stream);
    TORCH_CHECK(rc == BM2F_OK, "ffn backward (masked grad_h): ", bm2f_msda_last_error());
    auto gw2 = at::empty_like(w2c);
    auto gb2 = at::empty({d}, xc.options());
    rc = bm2f_linear_backward_weight(g.data_ptr(), hc.data_ptr(), gw2.data_ptr(), gb2.data_ptr(), static_cast<int>(rows),
                                     d, f, static_cast<int>(split), stream);
    TORCH_CHECK(rc == BM2F_OK, "ffn backward (grad_W2): ", bm2f_msda_last_error());
    auto gx = at::empty_like(xc);
    rc = bm2f_linear_backward_input(g1.data_ptr(), w1c.data_ptr(), gx.data_ptr(), ws.data_ptr(), static_cast<int>(rows), f,
                                    d, static_cast<int>(split), stream);
    TORCH_CHECK(rc == BM2F_OK, "ffn backward (grad_x): ", bm2f_msda_last_error());
    auto gw1 = at::empty_like(w1c);
    auto gb1 = at::empty({f}, xc.options());
    rc = bm2f_linear_backward_weight(g1.data_ptr(), xc.data_ptr(), gw1.data_ptr(), gb1.data_ptr(), static_cast<int>(rows),
                                     f, d, static_cast<int>(split), stream);
    TORCH_CHECK(rc == BM2F_OK, "ffn backward (grad_W1): ", bm2f_msda_last_error());
    return {gx, gw1, gb1, gw2, gb2};
}

// in place: x[row] = 0 where mask[row]; x (..., C) float32 contiguous, mask (...) bool
void zero_masked_rows_(at::Tensor x, const at::Tensor &mask)
{
    TORCH_CHECK(x.is_cuda() && mask.is_cuda(), "zero_masked_rows_: CUDA tensors only");
    TORCH_CHECK(x.scalar_type() == at::kFloat && x.is_contiguous(), "zero_masked_rows_: contiguous float32 tensor");
    TORCH_CHECK(mask.scalar_type() == at::kBool, "zero_masked_rows_: bool mask");
    const int64_t c = x.size(-1);
    auto m = mask.contiguous();
    TORCH_CHECK(m.numel() * c == x.numel(), "zero_masked_rows_: mask must have one entry per row");
    const c10::cuda::CUDAGuard guard(x.device());
    const int rc = bm2f_zero_masked_rows(x.data_ptr(), m.data_ptr(), static_cast<int>(m.numel()), static_cast<int>(c),
                                         at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "zero_masked_rows_: ", bm2f_msda_last_error());
}

std::vector<at::Tensor> add_layernorm_forward(const at::Tensor &x, const at::Tensor &residual, const at::Tensor &gamma,
                                              const at::Tensor &beta, double eps)
{
    TORCH_CHECK(x.is_cuda() && residual.is_cuda(), "add_layernorm: CUDA tensors only (no CPU path)");
    TORCH_CHECK(x.scalar_type() == at::kFloat && x.size(-1) == 256 && x.sizes() == residual.sizes(),
                "add_layernorm: float32, 256 channels, equal shapes");
    const c10::cuda::CUDAGuard guard(x.device());
    auto xc = x.contiguous();
    auto rc_ = residual.contiguous();
    const int64_t rows = xc.numel() / 256;
    auto z = at::empty_like(xc);
    auto y = at::empty_like(xc);
    auto mean = at::empty({rows}, xc.options());
    auto rstd = at::empty({rows}, xc.options());
    const int rc = bm2f_add_layernorm_forward(xc.data_ptr(), rc_.data_ptr(), gamma.contiguous().data_ptr(),
                                              beta.contiguous().data_ptr(), static_cast<float>(eps), z.data_ptr(),
                                              y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), static_cast<int>(rows), 256,
                                              at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "add_layernorm_forward: ", bm2f_msda_last_error());
    return {y, z, mean, rstd};
}

std::vector<at::Tensor> add_layernorm_backward(const at::Tensor &grad_y, const at::Tensor &z, const at::Tensor &mean,
                                               const at::Tensor &rstd, const at::Tensor &gamma)
{
    const c10::cuda::CUDAGuard guard(z.device());
    auto g = grad_y.contiguous();
    const int64_t rows = z.numel() / 256;
    auto dz = at::empty_like(z);
    auto dgamma = at::empty({256}, z.options());
    auto dbeta = at::empty({256}, z.options());
    const int rc = bm2f_add_layernorm_backward(g.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                               gamma.contiguous().data_ptr(), dz.data_ptr(), dgamma.data_ptr(),
                                               dbeta.data_ptr(), static_cast<int>(rows), 256,
                                               at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "add_layernorm_backward: ", bm2f_msda_last_error());
    return {dz, dgamma, dbeta};
}

// ---- pixel-decoder glue (msdeformattn.py:214-227, 316-325; position_encoding.py:29-52) -------------------------------
at::Tensor transpose_batched(const at::Tensor &x)
{
    TORCH_CHECK(x.is_cuda(), "transpose_batched: CUDA tensors only (no CPU path)");
    TORCH_CHECK(x.scalar_type() == at::kFloat && x.dim() == 3, "transpose_batched: float32 (batch, rows, cols)");
    const c10::cuda::CUDAGuard guard(x.device());
    auto xc = x.contiguous();
    auto out = at::empty({xc.size(0), xc.size(2), xc.size(1)}, xc.options());
    const int rc = bm2f_transpose_batched(xc.data_ptr(), out.data_ptr(), static_cast<int>(xc.size(0)),
                                          static_cast<int>(xc.size(1)), static_cast<int>(xc.size(2)),
                                          at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "transpose_batched: ", bm2f_msda_last_error());
    return out;
}

// y (batch, tokens, 256) -> out[:, token_offset : token_offset + tokens, :] of the (batch, S, 256) encoder input
std::vector<at::Tensor> groupnorm_tokens_forward(const at::Tensor &y, const at::Tensor &gamma, const at::Tensor &beta,
                                                 double eps, at::Tensor &out, int64_t token_offset)
{
    TORCH_CHECK(y.is_cuda() && out.is_cuda(), "groupnorm_tokens: CUDA tensors only (no CPU path)");
    TORCH_CHECK(y.scalar_type() == at::kFloat && y.dim() == 3 && y.size(2) == 256 && y.is_contiguous(),
                "groupnorm_tokens: y must be contiguous float32 (batch, tokens, 256)");
    TORCH_CHECK(out.scalar_type() == at::kFloat && out.dim() == 3 && out.size(2) == 256 && out.is_contiguous() &&
                    out.size(0) == y.size(0) && token_offset >= 0 && token_offset + y.size(1) <= out.size(1),
                "groupnorm_tokens: out must be contiguous float32 (batch, S, 256) holding the level's token range");
    const c10::cuda::CUDAGuard guard(y.device());
    const int batch = static_cast<int>(y.size(0)), tokens = static_cast<int>(y.size(1));
    auto mean = at::empty({batch, 32}, y.options());
    auto rstd = at::empty({batch, 32}, y.options());
    auto ws = at::empty({static_cast<int64_t>(bm2f_groupnorm_tokens_workspace_bytes(batch))}, y.options().dtype(at::kByte));
    const int rc = bm2f_groupnorm_tokens_forward(
        y.data_ptr(), gamma.contiguous().data_ptr(), beta.contiguous().data_ptr(), static_cast<float>(eps),
        out.data_ptr<float>() + token_offset * 256, out.size(1) * 256, mean.data_ptr(), rstd.data_ptr(), ws.data_ptr(), batch,
        tokens, 256, 32, at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "groupnorm_tokens_forward: ", bm2f_msda_last_error());
    return {mean, rstd};
}

std::vector<at::Tensor> groupnorm_tokens_backward(const at::Tensor &grad_out, int64_t token_offset, const at::Tensor &y,
                                                  const at::Tensor &mean, const at::Tensor &rstd, const at::Tensor &gamma)
{
    TORCH_CHECK(grad_out.is_cuda() && grad_out.scalar_type() == at::kFloat && grad_out.dim() == 3 &&
                    grad_out.size(2) == 256 && grad_out.is_contiguous() && grad_out.size(0) == y.size(0) &&
                    token_offset >= 0 && token_offset + y.size(1) <= grad_out.size(1),
                "groupnorm_tokens_backward: grad_out must be contiguous float32 (batch, S, 256)");
    const c10::cuda::CUDAGuard guard(y.device());
    const int batch = static_cast<int>(y.size(0)), tokens = static_cast<int>(y.size(1));
    auto dy = at::empty_like(y);
    auto dgamma = at::empty({256}, y.options());
    auto dbeta = at::empty({256}, y.options());
    auto ws = at::empty({static_cast<int64_t>(bm2f_groupnorm_tokens_workspace_bytes(batch))}, y.options().dtype(at::kByte));
    const int rc = bm2f_groupnorm_tokens_backward(
        grad_out.data_ptr<float>() + token_offset * 256, grad_out.size(1) * 256, y.data_ptr(), mean.data_ptr(),
        rstd.data_ptr(), gamma.contiguous().data_ptr(), dy.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(),
        batch, tokens, 256, 32, at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "groupnorm_tokens_backward: ", bm2f_msda_last_error());
    return {dy, dgamma, dbeta};
}

// ---- FPN tail (bm2f_conv3x3_*, bm2f_fpn_*, bm2f_groupnorm_relu_tokens_*; msdeformattn.py:341-358) ---------------
namespace {
void check_tokens4(const at::Tensor &t, const char *what, int64_t halo)
{
    TORCH_CHECK(t.is_cuda(), what, ": CUDA tensors only (no CPU path)");
    TORCH_CHECK(t.scalar_type() == at::kFloat && t.dim() == 4 && t.size(3) == 256 && t.is_contiguous(), what,
                ": contiguous float32 (batch, height", halo ? " + 2" : "", ", width", halo ? " + 2" : "", ", 256) expected");
}
at::Tensor conv_workspace(const at::Tensor &like)
{
    return at::empty({static_cast<int64_t>(bm2f_conv3x3_workspace_bytes(256, 256))}, like.options().dtype(at::kByte));
}
}  // namespace

// x_halo (batch, H + 2, W + 2, 256) zero-haloed -> (batch, H, W, 256); weight (256, 256, 3, 3)
at::Tensor conv3x3_tokens_forward(const at::Tensor &x_halo, const at::Tensor &weight, int64_t split)
{
    check_tokens4(x_halo, "conv3x3_tokens_forward", 1);
    TORCH_CHECK(weight.is_cuda() && weight.scalar_type() == at::kFloat && weight.dim() == 4 && weight.size(0) == 256 &&
                    weight.size(1) == 256 && weight.size(2) == 3 && weight.size(3) == 3,
                "conv3x3_tokens_forward: weight must be float32 (256, 256, 3, 3)");
    const c10::cuda::CUDAGuard guard(x_halo.device());
    const int batch = static_cast<int>(x_halo.size(0)), h = static_cast<int>(x_halo.size(1)) - 2, w = static_cast<int>(x_halo.size(2)) - 2;
    TORCH_CHECK(h > 0 && w > 0, "conv3x3_tokens_forward: empty image");
    auto y = at::empty({batch, h, w, 256}, x_halo.options());
    auto ws = conv_workspace(x_halo);
    const int rc = bm2f_conv3x3_forward(x_halo.data_ptr(), weight.contiguous().data_ptr(), y.data_ptr(), ws.data_ptr(), batch, h,
                                        w, 256, 256, static_cast<int>(split), at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "conv3x3_tokens_forward: ", bm2f_msda_last_error());
    return y;
}

at::Tensor conv3x3_tokens_backward_input(const at::Tensor &grad_halo, const at::Tensor &weight, int64_t split)
{
    check_tokens4(grad_halo, "conv3x3_tokens_backward_input", 1);
    const c10::cuda::CUDAGuard guard(grad_halo.device());
    const int batch = static_cast<int>(grad_halo.size(0)), h = static_cast<int>(grad_halo.size(1)) - 2, w = static_cast<int>(grad_halo.size(2)) - 2;
    auto gx = at::empty({batch, h, w, 256}, grad_halo.options());
    auto ws = conv_workspace(grad_halo);
    const int rc = bm2f_conv3x3_backward_input(grad_halo.data_ptr(), weight.contiguous().data_ptr(), gx.data_ptr(), ws.data_ptr(),
                                               batch, h, w, 256, 256, static_cast<int>(split),
                                               at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "conv3x3_tokens_backward_input: ", bm2f_msda_last_error());
    return gx;
}

at::Tensor conv3x3_tokens_backward_weight(const at::Tensor &grad_halo, const at::Tensor &x_halo, int64_t split)
{
    check_tokens4(grad_halo, "conv3x3_tokens_backward_weight", 1);
    check_tokens4(x_halo, "conv3x3_tokens_backward_weight", 1);
    TORCH_CHECK(grad_halo.sizes() == x_halo.sizes(), "conv3x3_tokens_backward_weight: shape mismatch");
    const c10::cuda::CUDAGuard guard(grad_halo.device());
    const int batch = static_cast<int>(x_halo.size(0)), h = static_cast<int>(x_halo.size(1)) - 2, w = static_cast<int>(x_halo.size(2)) - 2;
    auto gw = at::empty({256, 256, 3, 3}, x_halo.options());
    auto ws = conv_workspace(x_halo);
    const int rc = bm2f_conv3x3_backward_weight(grad_halo.data_ptr(), x_halo.data_ptr(), gw.data_ptr(), ws.data_ptr(), batch, h, w,
                                                256, 256, static_cast<int>(split), at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "conv3x3_tokens_backward_weight: ", bm2f_msda_last_error());
    return gw;
}

// mean, rstd (batch, 32) of y (batch, tokens, 256)
std::vector<at::Tensor> groupnorm_tokens_stats(const at::Tensor &y, double eps)
{
    TORCH_CHECK(y.is_cuda(), "groupnorm_tokens_stats: CUDA tensors only (no CPU path)");
    TORCH_CHECK(y.scalar_type() == at::kFloat && y.dim() >= 3 && y.size(-1) == 256 && y.is_contiguous(),
                "groupnorm_tokens_stats: y must be contiguous float32 (batch, ..., 256)");
    const c10::cuda::CUDAGuard guard(y.device());
    const int batch = static_cast<int>(y.size(0)), tokens = static_cast<int>(y.numel() / y.size(0) / 256);
    auto mean = at::empty({batch, 32}, y.options());
    auto rstd = at::empty({batch, 32}, y.options());
    auto ws = at::empty({static_cast<int64_t>(bm2f_groupnorm_tokens_workspace_bytes(batch))}, y.options().dtype(at::kByte));
    const int rc = bm2f_groupnorm_tokens_stats(y.data_ptr(), static_cast<float>(eps), mean.data_ptr(), rstd.data_ptr(),
                                               ws.data_ptr(), batch, tokens, 256, 32, at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "groupnorm_tokens_stats: ", bm2f_msda_last_error());
    return {mean, rstd};
}

// y_halo (batch, H + 2, W + 2, 256) = GroupNorm(lateral (batch, H, W, 256)) + bilinear upsample of enc_level
// (batch, enc_h * enc_w, 256): rows contiguous, the batch stride is free (a level's slice of the (batch, S, 256) encoder
// output is passed as a view, no copy)
at::Tensor fpn_merge_forward(const at::Tensor &lateral, const at::Tensor &mean, const at::Tensor &rstd,
                             const at::Tensor &gamma, const at::Tensor &beta, const at::Tensor &enc_level, int64_t enc_h,
                             int64_t enc_w)
{
    check_tokens4(lateral, "fpn_merge_forward", 0);
    TORCH_CHECK(enc_level.is_cuda() && enc_level.scalar_type() == at::kFloat && enc_level.dim() == 3 &&
                    enc_level.size(2) == 256 && enc_level.stride(2) == 1 && enc_level.stride(1) == 256 &&
                    enc_level.size(0) == lateral.size(0) && enc_level.size(1) == enc_h * enc_w,
                "fpn_merge_forward: enc_level must be float32 (batch, enc_h * enc_w, 256) with contiguous rows");
    const c10::cuda::CUDAGuard guard(lateral.device());
    const int batch = static_cast<int>(lateral.size(0)), h = static_cast<int>(lateral.size(1)), w = static_cast<int>(lateral.size(2));
    auto y = at::empty({batch, h + 2, w + 2, 256}, lateral.options());
    const int rc = bm2f_fpn_merge_forward(lateral.data_ptr(), mean.contiguous().data_ptr(), rstd.contiguous().data_ptr(),
                                          gamma.contiguous().data_ptr(), beta.contiguous().data_ptr(), enc_level.data_ptr(),
                                          batch > 1 ? enc_level.stride(0) : enc_h * enc_w * 256, y.data_ptr(), batch, h, w,
                                          static_cast<int>(enc_h), static_cast<int>(enc_w), 256,
                                          at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "fpn_merge_forward: ", bm2f_msda_last_error());
    return y;
}

// adjoint of the upsample: grad_y (batch, H, W, 256) -> (batch, enc_h * enc_w, 256)
at::Tensor fpn_upsample_backward(const at::Tensor &grad_y, int64_t enc_h, int64_t enc_w)
{
    check_tokens4(grad_y, "fpn_upsample_backward", 0);
    const c10::cuda::CUDAGuard guard(grad_y.device());
    const int batch = static_cast<int>(grad_y.size(0)), h = static_cast<int>(grad_y.size(1)), w = static_cast<int>(grad_y.size(2));
    auto ge = at::empty({batch, enc_h * enc_w, 256}, grad_y.options());
    const int rc = bm2f_fpn_upsample_backward(grad_y.data_ptr(), ge.data_ptr(), enc_h * enc_w * 256, batch, h, w,
                                              static_cast<int>(enc_h), static_cast<int>(enc_w), 256,
                                              at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "fpn_upsample_backward: ", bm2f_msda_last_error());
    return ge;
}

at::Tensor groupnorm_relu_tokens_apply(const at::Tensor &y, const at::Tensor &mean, const at::Tensor &rstd,
                                       const at::Tensor &gamma, const at::Tensor &beta)
{
    check_tokens4(y, "groupnorm_relu_tokens_apply", 0);
    const c10::cuda::CUDAGuard guard(y.device());
    auto out = at::empty_like(y);
    const int rc = bm2f_groupnorm_relu_tokens_apply(y.data_ptr(), mean.contiguous().data_ptr(), rstd.contiguous().data_ptr(),
                                                    gamma.contiguous().data_ptr(), beta.contiguous().data_ptr(), out.data_ptr(),
                                                    static_cast<int>(y.size(0)), static_cast<int>(y.size(1) * y.size(2)), 256, 32,
                                                    at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "groupnorm_relu_tokens_apply: ", bm2f_msda_last_error());
    return out;
}

// returns grad_halo (batch, H + 2, W + 2, 256), grad_gamma, grad_beta
std::vector<at::Tensor> groupnorm_relu_tokens_backward(const at::Tensor &grad_out, const at::Tensor &y, const at::Tensor &mean,
                                                       const at::Tensor &rstd, const at::Tensor &gamma, const at::Tensor &beta)
{
    check_tokens4(y, "groupnorm_relu_tokens_backward", 0);
    check_tokens4(grad_out, "groupnorm_relu_tokens_backward", 0);
    TORCH_CHECK(grad_out.sizes() == y.sizes(), "groupnorm_relu_tokens_backward: shape mismatch");
    const c10::cuda::CUDAGuard guard(y.device());
    const int batch = static_cast<int>(y.size(0)), h = static_cast<int>(y.size(1)), w = static_cast<int>(y.size(2));
    auto gh = at::empty({batch, h + 2, w + 2, 256}, y.options());
    auto dgamma = at::empty({256}, y.options());
    auto dbeta = at::empty({256}, y.options());
    auto ws = at::empty({static_cast<int64_t>(bm2f_groupnorm_tokens_workspace_bytes(batch))}, y.options().dtype(at::kByte));
    const int rc = bm2f_groupnorm_relu_tokens_backward(grad_out.data_ptr(), y.data_ptr(), mean.contiguous().data_ptr(),
                                                       rstd.contiguous().data_ptr(), gamma.contiguous().data_ptr(),
                                                       beta.contiguous().data_ptr(), gh.data_ptr(), dgamma.data_ptr(),
                                                       dbeta.data_ptr(), ws.data_ptr(), batch, h, w, 256, 32,
                                                       at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "groupnorm_relu_tokens_backward: ", bm2f_msda_last_error());
    return {gh, dgamma, dbeta};
}

at::Tensor sine_position_embedding(const at::Tensor &like, int64_t height, int64_t width, int64_t num_pos_feats,
                                   double temperature, double scale, bool normalize)
{
    TORCH_CHECK(like.is_cuda(), "sine_position_embedding: CUDA device only (no CPU path)");
    const c10::cuda::CUDAGuard guard(like.device());
    auto out = at::empty({height * width, 2 * num_pos_feats}, like.options().dtype(at::kFloat));
    const int rc = bm2f_sine_position_embedding(out.data_ptr(), static_cast<int>(height), static_cast<int>(width),
                                                static_cast<int>(num_pos_feats), static_cast<float>(temperature),
                                                static_cast<float>(scale), normalize ? 1 : 0,
                                                at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == BM2F_OK, "sine_position_embedding: ", bm2f_msda_last_error());
    return out;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.doc() = "B200-native multi-scale deformable attention (drop-in for the Deformable-DETR op)";
    m.def("ms_deform_attn_forward", &ms_deform_attn_forward, "ms_deform_attn_forward");
    m.def("ms_deform_attn_backward", &ms_deform_attn_backward, "ms_deform_attn_backward");
    m.def("ms_deform_attn_fused_supported", &ms_deform_attn_fused_supported);
    m.def("ms_deform_attn_fused_forward", &ms_deform_attn_fused_forward, "fused softmax + location prologue + sampling");
    m.def("ms_deform_attn_fused_backward", &ms_deform_attn_fused_backward, "backward of the fused op");
    m.def("ms_deform_attn_fused_forward_packed", &ms_deform_attn_fused_forward_packed,
          "fused op on packed offsets||logits (N, Lq, M*L*P*3): the output of one 256 -> 288 projection");
    m.def("ms_deform_attn_fused_backward_packed", &ms_deform_attn_fused_backward_packed);
    m.def("linear_tf32x3", &linear_tf32x3, "tcgen05 projection GEMM (y = x W^T + b), split=3: tf32x3, 1: tf32");
    m.def("linear_tf32x3_supported", &linear_tf32x3_supported);
    m.def("linear_set_tuning", [](int variant, int dw_row_cap) {
        bm2f_linear_tuning_t t = {};
        t.variant = variant;
        t.dw_row_cap = dw_row_cap;
        bm2f_linear_set_tuning(&t);
    }, "A/B kernel variants of the tcgen05 GEMMs (bm2f_linear_tuning_t); (0, 0) = defaults", pybind11::arg("variant") = 0,
          pybind11::arg("dw_row_cap") = 0);
    m.def("linear_tf32x3_backward_weight", &linear_tf32x3_backward_weight, "grad_W = grad_y^T x, grad_b = sum grad_y (tcgen05)");
    m.def("linear_tf32x3_backward_input", &linear_tf32x3_backward_input, "grad_x = grad_y @ weight (tcgen05)");
    m.def("linear_tf32x3_backward_input_accumulate", &linear_tf32x3_backward_input_accumulate,
          "grad_x = grad_y @ weight + addend, summed in the GEMM epilogue");
    m.def("linear_relu_tf32x3", &linear_relu_tf32x3, "y = relu(x W^T + b) on tcgen05");
    m.def("ffn_tf32x3_backward", &ffn_tf32x3_backward, "backward of linear2(relu(linear1(x)))");
    m.def("zero_masked_rows_", &zero_masked_rows_, "in-place masked_fill(mask[..., None], 0) touching only masked rows");
    m.def("add_layernorm_forward", &add_layernorm_forward, "z = x + r, y = LayerNorm(z)");
    m.def("add_layernorm_backward", &add_layernorm_backward);
    m.def("transpose_batched", &transpose_batched, "(B, R, C) -> (B, C, R)");
    m.def("groupnorm_tokens_forward", &groupnorm_tokens_forward, "GroupNorm(32, 256) on token rows, written into the encoder input");
    m.def("groupnorm_tokens_backward", &groupnorm_tokens_backward);
    m.def("conv3x3_tokens_forward", &conv3x3_tokens_forward, "3x3 conv (256 -> 256, padding 1) of a zero-haloed token image on tcgen05");
    m.def("conv3x3_set_variant", [](int pair_tiles) { bm2f_conv3x3_set_variant(pair_tiles); },
          "A/B: 1 (default) two row tiles per weight k-block, 0 one");
    m.def("conv3x3_tokens_backward_input", &conv3x3_tokens_backward_input);
    m.def("conv3x3_tokens_backward_weight", &conv3x3_tokens_backward_weight);
    m.def("groupnorm_tokens_stats", &groupnorm_tokens_stats);
    m.def("fpn_merge_forward", &fpn_merge_forward, "haloed y = GroupNorm(lateral) + bilinear upsample of an encoder level");
    m.def("fpn_upsample_backward", &fpn_upsample_backward);
    m.def("groupnorm_relu_tokens_apply", &groupnorm_relu_tokens_apply);
    m.def("groupnorm_relu_tokens_backward", &groupnorm_relu_tokens_backward);
    m.def("sine_position_embedding", &sine_position_embedding, "PositionEmbeddingSine for an all-False mask, token-major");
    m.def("abi_version", []() { return bm2f_msda_abi_version(); });
    m.def("build_info", []() { return std::string(bm2f_msda_build_info()); });
    m.def("launch_count", []() { return bm2f_msda_launch_count(); });
}

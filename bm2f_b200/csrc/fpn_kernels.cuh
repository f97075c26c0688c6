// FPN tail of the pixel decoder on token-major rows (reference:
// /root/reference/mask2former/modeling/pixel_decoder/msdeformattn.py:341-358):
//
//     cur_fpn = lateral_conv(x)                                   1x1 conv (no bias) + GroupNorm(32, 256)
//     y       = cur_fpn + F.interpolate(out[-1], size=cur_fpn.shape[-2:], mode="bilinear", align_corners=False)
//     y       = output_conv(y)                                    3x3 conv (no bias) + GroupNorm + ReLU
//     mask_features(y)                                            1x1 conv with bias
//
// Everything is kept as (image, pixel, 256-channel) rows.  The 3x3 convolution reads its input from a ZERO-HALOED image
// (batch, H + 2, W + 2, 256): with the halo in memory a filter tap is a constant row shift, so the activation tile of a
// tap is a plain 2-D TMA box (linear_tf32x3_persistent_kernel<.., CONV>) and padding needs no predicates.  The kernels
// here produce / consume that layout:
//
//   fpn_merge_forward_kernel     haloed y = GroupNorm(lateral) + bilinear upsample of the encoder's finest level
//   fpn_upsample_backward_kernel adjoint of the upsample as a gather (deterministic, no atomics)
//   groupnorm_relu_apply_kernel  out = relu(GroupNorm(y))
//   groupnorm_relu_bwd_*         its backward; grad rows are written into a zero-haloed image for the conv gradients
//   conv_dw_tma_kernel           grad_W of the 3x3 conv, single TF32 pass: both operands are pixel-major, i.e. MN-major
//                                for an MMA that reduces over pixels, and go from global memory to the tensor core through
//                                TMA alone (no transposing producers)
#pragma once

#include "linear_tf32x3.cuh"

namespace bm2f {

constexpr int kFpnC = 256;
constexpr int kFpnGroups = 32;

// ---- bilinear source index, torch's area_pixel_compute_source_index for align_corners = False ----------
struct BilinearTap {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ BilinearTap bilinear_tap(float scale, int dst, int in_size)
{
    float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    BilinearTap t;
    t.i0 = static_cast<int>(src);
    if (t.i0 > in_size - 1) t.i0 = in_size - 1;
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l1 = src - static_cast<float>(t.i0);
    t.l0 = 1.f - t.l1;
    return t;
}

// y_halo[n, Y + 1, X + 1, :] = (lat[n, Y, X, :] - mean) * rstd * gamma + beta + bilinear(enc[n], Y, X); halo rows = 0.
// One warp per haloed row, lane = 8 channels.  enc rows at enc + n * enc_batch_stride + (i * w + j) * 256.
__global__ void __launch_bounds__(256) fpn_merge_forward_kernel(const float *__restrict__ lat, const float *__restrict__ mean,
                                                                const float *__restrict__ rstd,
                                                                const float *__restrict__ gamma,
                                                                const float *__restrict__ beta,
                                                                const float *__restrict__ enc, long long enc_batch_stride,
                                                                float *__restrict__ y_halo, int batch, int H, int W, int h,
                                                                int w, float scale_h, float scale_w)
{
    const int lane = threadIdx.x & 31;
    const long long rows = static_cast<long long>(batch) * (H + 2) * (W + 2);
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane), g1 = __ldg(reinterpret_cast<const float4 *>(gamma) + 32 + lane);
    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(beta) + lane), b1 = __ldg(reinterpret_cast<const float4 *>(beta) + 32 + lane);
    for (long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); r < rows; r += static_cast<long long>(gridDim.x) * 8) {
        const int n = static_cast<int>(r / ((H + 2) * (W + 2)));
        const int rem = static_cast<int>(r - static_cast<long long>(n) * (H + 2) * (W + 2));
        const int yp = rem / (W + 2), xp = rem - yp * (W + 2);
        float4 *o = reinterpret_cast<float4 *>(y_halo + r * kFpnC);
        if (yp < 1 || yp > H || xp < 1 || xp > W) {
            o[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            o[32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const int Y = yp - 1, X = xp - 1;
        const float m0 = __ldg(mean + n * kFpnGroups + (lane >> 1)), r0 = __ldg(rstd + n * kFpnGroups + (lane >> 1));
        const float m1 = __ldg(mean + n * kFpnGroups + 16 + (lane >> 1)), r1 = __ldg(rstd + n * kFpnGroups + 16 + (lane >> 1));
        const float4 *lrow = reinterpret_cast<const float4 *>(lat + (static_cast<size_t>(n) * H * W + static_cast<size_t>(Y) * W + X) * kFpnC);
        const float4 a = __ldg(lrow + lane), b = __ldg(lrow + 32 + lane);
        const BilinearTap ty = bilinear_tap(scale_h, Y, h), tx = bilinear_tap(scale_w, X, w);
        const float *eimg = enc + static_cast<size_t>(n) * enc_batch_stride;
        const float4 *e00 = reinterpret_cast<const float4 *>(eimg + (static_cast<size_t>(ty.i0) * w + tx.i0) * kFpnC);
        const float4 *e01 = reinterpret_cast<const float4 *>(eimg + (static_cast<size_t>(ty.i0) * w + tx.i1) * kFpnC);
        const float4 *e10 = reinterpret_cast<const float4 *>(eimg + (static_cast<size_t>(ty.i1) * w + tx.i0) * kFpnC);
        const float4 *e11 = reinterpret_cast<const float4 *>(eimg + (static_cast<size_t>(ty.i1) * w + tx.i1) * kFpnC);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int k = half * 32 + lane;
            const float4 v00 = __ldg(e00 + k), v01 = __ldg(e01 + k), v10 = __ldg(e10 + k), v11 = __ldg(e11 + k);
            const float4 v = half ? b : a, g = half ? g1 : g0, bb = half ? b1 : b0;
            const float m = half ? m1 : m0, rs = half ? r1 : r0;
            float4 out;
            // same operation order as torch's upsample_bilinear2d: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)
            out.x = ((v.x - m) * rs * g.x + bb.x) + (ty.l0 * (tx.l0 * v00.x + tx.l1 * v01.x) + ty.l1 * (tx.l0 * v10.x + tx.l1 * v11.x));
            out.y = ((v.y - m) * rs * g.y + bb.y) + (ty.l0 * (tx.l0 * v00.y + tx.l1 * v01.y) + ty.l1 * (tx.l0 * v10.y + tx.l1 * v11.y));
            out.z = ((v.z - m) * rs * g.z + bb.z) + (ty.l0 * (tx.l0 * v00.z + tx.l1 * v01.z) + ty.l1 * (tx.l0 * v10.z + tx.l1 * v11.z));
            out.w = ((v.w - m) * rs * g.w + bb.w) + (ty.l0 * (tx.l0 * v00.w + tx.l1 * v01.w) + ty.l1 * (tx.l0 * v10.w + tx.l1 * v11.w));
            o[k] = out;
        }
    }
}

// grad_enc[n, i, j, :] = sum over the fine pixels (Y, X) whose bilinear footprint contains (i, j) of weight * grad_y[n, Y, X, :]
// (grad_y dense (batch, H, W, 256)).  One warp per coarse pixel; the candidate fine rows / columns are the ones whose
// source index can reach i resp. j, their weights are recomputed exactly as in the forward kernel.
__global__ void __launch_bounds__(256) fpn_upsample_backward_kernel(const float *__restrict__ grad_y,
                                                                    float *__restrict__ grad_enc,
                                                                    long long grad_enc_batch_stride, int batch, int H,
                                                                    int W, int h, int w, float scale_h, float scale_w)
{
    const int lane = threadIdx.x & 31;
    const long long pix = static_cast<long long>(batch) * h * w;
    for (long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); r < pix; r += static_cast<long long>(gridDim.x) * 8) {
        const int n = static_cast<int>(r / (h * w));
        const int rem = static_cast<int>(r - static_cast<long long>(n) * h * w);
        const int i = rem / w, j = rem - i * w;
        // fine rows whose (clamped) source index lies in (i - 1, i + 1): Y in ((i - 0.5) / s - 0.5, (i + 1.5) / s - 0.5)
        int y_lo = static_cast<int>(floorf((static_cast<float>(i) - 0.5f) / scale_h - 0.5f)) - 1;
        int y_hi = static_cast<int>(ceilf((static_cast<float>(i) + 1.5f) / scale_h - 0.5f)) + 1;
        int x_lo = static_cast<int>(floorf((static_cast<float>(j) - 0.5f) / scale_w - 0.5f)) - 1;
        int x_hi = static_cast<int>(ceilf((static_cast<float>(j) + 1.5f) / scale_w - 0.5f)) + 1;
        if (i == 0) y_lo = 0;              // sources clamped to 0 all land on row 0
        if (j == 0) x_lo = 0;
        if (i == h - 1) y_hi = H - 1;      // and the ones past the last source row on row h - 1
        if (j == w - 1) x_hi = W - 1;
        y_lo = max(y_lo, 0); x_lo = max(x_lo, 0); y_hi = min(y_hi, H - 1); x_hi = min(x_hi, W - 1);
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *gimg = grad_y + static_cast<size_t>(n) * H * W * kFpnC;
        for (int Y = y_lo; Y <= y_hi; ++Y) {
            const BilinearTap ty = bilinear_tap(scale_h, Y, h);
            const float wy = (ty.i0 == i ? ty.l0 : 0.f) + (ty.i1 == i ? ty.l1 : 0.f);
            if (wy == 0.f) continue;
            for (int X = x_lo; X <= x_hi; ++X) {
                const BilinearTap tx = bilinear_tap(scale_w, X, w);
                const float wx = (tx.i0 == j ? tx.l0 : 0.f) + (tx.i1 == j ? tx.l1 : 0.f);
                if (wx == 0.f) continue;
                const float wgt = wy * wx;
                const float4 *g = reinterpret_cast<const float4 *>(gimg + (static_cast<size_t>(Y) * W + X) * kFpnC);
                const float4 a = __ldg(g + lane), b = __ldg(g + 32 + lane);
                acc0.x += wgt * a.x; acc0.y += wgt * a.y; acc0.z += wgt * a.z; acc0.w += wgt * a.w;
                acc1.x += wgt * b.x; acc1.y += wgt * b.y; acc1.z += wgt * b.z; acc1.w += wgt * b.w;
            }
        }
        float4 *o = reinterpret_cast<float4 *>(grad_enc + static_cast<size_t>(n) * grad_enc_batch_stride + static_cast<size_t>(rem) * kFpnC);
        o[lane] = acc0;
        o[32 + lane] = acc1;
    }
}

// out = relu((y - mean) * rstd * gamma + beta), dense rows
__global__ void __launch_bounds__(256) groupnorm_relu_apply_kernel(const float *__restrict__ y, const float *__restrict__ mean,
                                                                   const float *__restrict__ rstd,
                                                                   const float *__restrict__ gamma,
                                                                   const float *__restrict__ beta, float *__restrict__ out,
                                                                   int tokens)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.y;
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane), g1 = __ldg(reinterpret_cast<const float4 *>(gamma) + 32 + lane);
    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(beta) + lane), b1 = __ldg(reinterpret_cast<const float4 *>(beta) + 32 + lane);
    const float m0 = __ldg(mean + n * kFpnGroups + (lane >> 1)), r0 = __ldg(rstd + n * kFpnGroups + (lane >> 1));
    const float m1 = __ldg(mean + n * kFpnGroups + 16 + (lane >> 1)), r1 = __ldg(rstd + n * kFpnGroups + 16 + (lane >> 1));
    const float *img = y + static_cast<size_t>(n) * tokens * kFpnC;
    float *dst = out + static_cast<size_t>(n) * tokens * kFpnC;
    for (int t = blockIdx.x * 8 + warp; t < tokens; t += gridDim.x * 8) {
        const float4 *row = reinterpret_cast<const float4 *>(img + static_cast<size_t>(t) * kFpnC);
        const float4 a = __ldg(row + lane), b = __ldg(row + 32 + lane);
        float4 *o = reinterpret_cast<float4 *>(dst + static_cast<size_t>(t) * kFpnC);
        o[lane] = make_float4(fmaxf((a.x - m0) * r0 * g0.x + b0.x, 0.f), fmaxf((a.y - m0) * r0 * g0.y + b0.y, 0.f),
                              fmaxf((a.z - m0) * r0 * g0.z + b0.z, 0.f), fmaxf((a.w - m0) * r0 * g0.w + b0.w, 0.f));
        o[32 + lane] = make_float4(fmaxf((b.x - m1) * r1 * g1.x + b1.x, 0.f), fmaxf((b.y - m1) * r1 * g1.y + b1.y, 0.f),
                                   fmaxf((b.z - m1) * r1 * g1.z + b1.z, 0.f), fmaxf((b.w - m1) * r1 * g1.w + b1.w, 0.f));
    }
}

// backward pass 1 of relu(GroupNorm(y)): with dy' = grad_out where the forward output was positive, else 0:
// sums[(n * 32 + g) * 2 + {0, 1}] += sum dy' * gamma, sum dy' * gamma * xhat; dgamma[c] += sum dy' * xhat, dbeta[c] += sum dy'
__global__ void __launch_bounds__(256) groupnorm_relu_bwd_stats_kernel(const float *__restrict__ grad_out,
                                                                       const float *__restrict__ y,
                                                                       const float *__restrict__ mean,
                                                                       const float *__restrict__ rstd,
                                                                       const float *__restrict__ gamma,
                                                                       const float *__restrict__ beta,
                                                                       double *__restrict__ sums, float *__restrict__ dgamma,
                                                                       float *__restrict__ dbeta, int tokens)
{
    __shared__ float s_grp[8][32][2];
    __shared__ float s_ch[8][256][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.y;
    float gam[8], bet[8];
    {
        const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane), g1 = __ldg(reinterpret_cast<const float4 *>(gamma) + 32 + lane);
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(beta) + lane), b1 = __ldg(reinterpret_cast<const float4 *>(beta) + 32 + lane);
        gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
        bet[0] = b0.x; bet[1] = b0.y; bet[2] = b0.z; bet[3] = b0.w; bet[4] = b1.x; bet[5] = b1.y; bet[6] = b1.z; bet[7] = b1.w;
    }
    const float m0 = __ldg(mean + n * kFpnGroups + (lane >> 1)), r0 = __ldg(rstd + n * kFpnGroups + (lane >> 1));
    const float m1 = __ldg(mean + n * kFpnGroups + 16 + (lane >> 1)), r1 = __ldg(rstd + n * kFpnGroups + 16 + (lane >> 1));
    const float *img = y + static_cast<size_t>(n) * tokens * kFpnC;
    const float *gimg = grad_out + static_cast<size_t>(n) * tokens * kFpnC;
    float dg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, db[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float sa0 = 0.f, sb0 = 0.f, sa1 = 0.f, sb1 = 0.f;      // per lane: group lane/2 (first half) and 16 + lane/2 (second)
    for (int t = blockIdx.x * 8 + warp; t < tokens; t += gridDim.x * 8) {
        const float4 *row = reinterpret_cast<const float4 *>(img + static_cast<size_t>(t) * kFpnC);
        const float4 *grow = reinterpret_cast<const float4 *>(gimg + static_cast<size_t>(t) * kFpnC);
        const float4 a = __ldg(row + lane), b = __ldg(row + 32 + lane);
        const float4 da = __ldg(grow + lane), dc = __ldg(grow + 32 + lane);
        const float yv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const float dv[8] = {da.x, da.y, da.z, da.w, dc.x, dc.y, dc.z, dc.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float m = c < 4 ? m0 : m1, rs = c < 4 ? r0 : r1;
            const float xhat = (yv[c] - m) * rs;
            const float d = (xhat * gam[c] + bet[c] > 0.f) ? dv[c] : 0.f;
            dg[c] += d * xhat;
            db[c] += d;
            const float dgm = d * gam[c];
            if (c < 4) { sa0 += dgm; sb0 += dgm * xhat; }
            else { sa1 += dgm; sb1 += dgm * xhat; }
        }
    }
    // lanes 2g, 2g + 1 hold the two halves of group g (first 128 channels) / 16 + g (second 128 channels)
    sa0 += __shfl_xor_sync(0xffffffffu, sa0, 1); sb0 += __shfl_xor_sync(0xffffffffu, sb0, 1);
    sa1 += __shfl_xor_sync(0xffffffffu, sa1, 1); sb1 += __shfl_xor_sync(0xffffffffu, sb1, 1);
    if ((lane & 1) == 0) {
        s_grp[warp][lane >> 1][0] = sa0; s_grp[warp][lane >> 1][1] = sb0;
        s_grp[warp][16 + (lane >> 1)][0] = sa1; s_grp[warp][16 + (lane >> 1)][1] = sb1;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int ch = (c < 4 ? 0 : 128) + lane * 4 + (c & 3);
        s_ch[warp][ch][0] = dg[c];
        s_ch[warp][ch][1] = db[c];
    }
    __syncthreads();
    {
        const int ch = threadIdx.x;
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) { a += s_ch[w8][ch][0]; b += s_ch[w8][ch][1]; }
        atomicAdd(dgamma + ch, a);
        atomicAdd(dbeta + ch, b);
    }
    if (threadIdx.x < 64) {
        const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
        double acc = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) acc += static_cast<double>(s_grp[w8][g][which]);
        atomicAdd(sums + (static_cast<size_t>(n) * kFpnGroups + g) * 2 + which, acc);
    }
}

// backward pass 2: grad_y = rstd * (dy' * gamma - c1 - xhat * c2), written into the zero-haloed image
// grad_halo (batch, H + 2, W + 2, 256) (one warp per haloed row; halo rows are written as zeros)
__global__ void __launch_bounds__(256) groupnorm_relu_bwd_apply_kernel(const float *__restrict__ grad_out,
                                                                       const float *__restrict__ y,
                                                                       const float *__restrict__ mean,
                                                                       const float *__restrict__ rstd,
                                                                       const float *__restrict__ c1,
                                                                       const float *__restrict__ c2,
                                                                       const float *__restrict__ gamma,
                                                                       const float *__restrict__ beta,
                                                                       float *__restrict__ grad_halo, int batch, int H, int W)
{
    const int lane = threadIdx.x & 31;
    const long long rows = static_cast<long long>(batch) * (H + 2) * (W + 2);
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane), g1 = __ldg(reinterpret_cast<const float4 *>(gamma) + 32 + lane);
    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(beta) + lane), b1 = __ldg(reinterpret_cast<const float4 *>(beta) + 32 + lane);
    for (long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); r < rows; r += static_cast<long long>(gridDim.x) * 8) {
        const int n = static_cast<int>(r / ((H + 2) * (W + 2)));
        const int rem = static_cast<int>(r - static_cast<long long>(n) * (H + 2) * (W + 2));
        const int yp = rem / (W + 2), xp = rem - yp * (W + 2);
        float4 *o = reinterpret_cast<float4 *>(grad_halo + r * kFpnC);
        if (yp < 1 || yp > H || xp < 1 || xp > W) {
            o[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            o[32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const size_t t = static_cast<size_t>(n) * H * W + static_cast<size_t>(yp - 1) * W + (xp - 1);
        const int ga = n * kFpnGroups + (lane >> 1), gb = ga + 16;
        const float m0 = __ldg(mean + ga), r0 = __ldg(rstd + ga), p0 = __ldg(c1 + ga), q0 = __ldg(c2 + ga);
        const float m1 = __ldg(mean + gb), r1 = __ldg(rstd + gb), p1 = __ldg(c1 + gb), q1 = __ldg(c2 + gb);
        const float4 *row = reinterpret_cast<const float4 *>(y + t * kFpnC);
        const float4 *grow = reinterpret_cast<const float4 *>(grad_out + t * kFpnC);
        const float4 a = __ldg(row + lane), b = __ldg(row + 32 + lane);
        const float4 da = __ldg(grow + lane), dc = __ldg(grow + 32 + lane);
        auto one = [](float yv, float dv, float m, float rs, float g, float bb, float p, float q) {
            const float xhat = (yv - m) * rs;
            const float d = (xhat * g + bb > 0.f) ? dv : 0.f;
            return rs * (d * g - p - xhat * q);
        };
        o[lane] = make_float4(one(a.x, da.x, m0, r0, g0.x, b0.x, p0, q0), one(a.y, da.y, m0, r0, g0.y, b0.y, p0, q0),
                              one(a.z, da.z, m0, r0, g0.z, b0.z, p0, q0), one(a.w, da.w, m0, r0, g0.w, b0.w, p0, q0));
        o[32 + lane] = make_float4(one(b.x, dc.x, m1, r1, g1.x, b1.x, p1, q1), one(b.y, dc.y, m1, r1, g1.y, b1.y, p1, q1),
                                   one(b.z, dc.z, m1, r1, g1.z, b1.z, p1, q1), one(b.w, dc.w, m1, r1, g1.w, b1.w, p1, q1));
    }
}

// conv weight (O, C, 3, 3) -> K-major GEMM weights, exact TF32 split.
//   mode 0 (forward):        out[o][t * C + c] = w[o][c][t]              (O rows of 9 C)
//   mode 1 (input gradient): out[c][t * O + o] = w[o][c][8 - t]          (C rows of 9 O: flipped taps, transposed channels)
__global__ void conv3x3_weight_prep_kernel(const float *__restrict__ w, float *__restrict__ hi, float *__restrict__ lo, int O,
                                           int C, int mode)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= O * C * 9) return;
    float x;
    if (mode == 0) {
        const int o = i / (9 * C), r = i - o * 9 * C, t = r / C, c = r - t * C;
        x = w[(static_cast<size_t>(o) * C + c) * 9 + t];
    } else {
        const int c = i / (9 * O), r = i - c * 9 * O, t = r / O, o = r - t * O;
        x = w[(static_cast<size_t>(o) * C + c) * 9 + (8 - t)];
    }
    const float h = tf32_hi(x);
    hi[i] = h;
    lo[i] = x - h;
}

// dw_k (O, 9 C) in the forward GEMM layout -> grad of the (O, C, 3, 3) weight
__global__ void conv3x3_weight_grad_unpack_kernel(const float *__restrict__ dwk, float *__restrict__ dw, int O, int C)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= O * C * 9) return;
    const int o = i / (9 * C), r = i - o * 9 * C, c = r / 9, t = r - c * 9;
    dw[i] = dwk[static_cast<size_t>(o) * 9 * C + t * C + c];
}

}  // namespace bm2f

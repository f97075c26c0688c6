// Pixel-decoder glue around the encoder (reference: msdeformattn.py:214-227, 316-325 and
// transformer_decoder/position_encoding.py:29-52), token-major throughout:
//
//   transpose_batched      (B, R, C) -> (B, C, R): NCHW backbone feature <-> (tokens, channels) rows, 32 x 32 smem tiles
//   groupnorm_tokens_*     GroupNorm(32, 256) of the 1x1-conv output, computed on token-major rows and written straight
//                          into this level's slice of the concatenated (N, S, 256) encoder input (replaces
//                          GroupNorm on NCHW + flatten(2).transpose(1, 2) + torch.cat, msdeformattn.py:66-82)
//   sine_pos_embed         PositionEmbeddingSine(128, normalize=True) for an all-False mask: batch-independent, so it
//                          is produced once per level shape as (H*W, 256) rows instead of N NCHW copies
//
// Statistics: per-CTA fp32 partial sums (<= a few thousand elements each) are combined in fp64 atomics, and
// var = E[y^2] - mean^2 is evaluated in fp64.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ln_kernels.cuh"

namespace bm2f {

constexpr int kGnC = 256;       // conv_dim
constexpr int kGnGroups = 32;   // nn.GroupNorm(32, conv_dim): 8 channels = 2 float4 per group

// in (B, R, C) -> out (B, C, R)
static __global__ void __launch_bounds__(256) transpose_batched_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                                int R, int C)
{
    __shared__ float tile[32][33];
    const size_t base = static_cast<size_t>(blockIdx.z) * R * C;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int r = r0 + ty + i, c = c0 + tx;
        if (r < R && c < C) tile[ty + i][tx] = __ldg(in + base + static_cast<size_t>(r) * C + c);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;
        if (r < R && c < C) out[base + static_cast<size_t>(c) * R + r] = tile[tx][ty + i];
    }
}

// sums[(n * 32 + g) * 2 + {0, 1}] += sum y, sum y^2 over this CTA's rows of image n.  grid (chunks, batch).
static __global__ void __launch_bounds__(256) groupnorm_tokens_stats_kernel(const float *__restrict__ y, double *__restrict__ sums,
                                                                     int tokens)
{
    __shared__ float s_part[8][64][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.y;
    const float *img = y + static_cast<size_t>(n) * tokens * kGnC;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
    for (int t = blockIdx.x * 8 + warp; t < tokens; t += gridDim.x * 8) {
        const float4 *row = reinterpret_cast<const float4 *>(img + static_cast<size_t>(t) * kGnC);
        const float4 a = __ldg(row + lane), b = __ldg(row + 32 + lane);
        s0 += (a.x + a.y) + (a.z + a.w);
        q0 += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
        s1 += (b.x + b.y) + (b.z + b.w);
        q1 += (b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w);
    }
    s_part[warp][lane][0] = s0;
    s_part[warp][lane][1] = q0;
    s_part[warp][32 + lane][0] = s1;
    s_part[warp][32 + lane][1] = q1;
    __syncthreads();
    if (threadIdx.x < 64) {   // thread = (group, which): two float4 slots per group
        const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += static_cast<double>(s_part[w][2 * g][which]) + static_cast<double>(s_part[w][2 * g + 1][which]);
        atomicAdd(sums + (static_cast<size_t>(n) * kGnGroups + g) * 2 + which, acc);
    }
}

// forward: mean / rstd;  backward: c1 = sum(dy*gamma)/cnt, c2 = sum(dy*gamma*xhat)/cnt
static __global__ void groupnorm_tokens_finalize_kernel(const double *__restrict__ sums, float *__restrict__ a, float *__restrict__ b,
                                                 int n_stats, double count, float eps, int forward)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_stats) return;
    const double s = sums[2 * i], q = sums[2 * i + 1];
    if (forward) {
        const double m = s / count;
        double v = q / count - m * m;
        if (v < 0.0) v = 0.0;
        a[i] = static_cast<float>(m);
        b[i] = static_cast<float>(1.0 / sqrt(v + static_cast<double>(eps)));
    } else {
        a[i] = static_cast<float>(s / count);
        b[i] = static_cast<float>(q / count);
    }
}

// out[n, t, :] = (y[n, t, :] - mean[n, g]) * rstd[n, g] * gamma + beta, out rows at out + n * out_batch_stride + t * 256
static __global__ void __launch_bounds__(256) groupnorm_tokens_apply_kernel(const float *__restrict__ y, const float *__restrict__ mean,
                                                                     const float *__restrict__ rstd,
                                                                     const float *__restrict__ gamma,
                                                                     const float *__restrict__ beta, float *__restrict__ out,
                                                                     long long out_batch_stride, int tokens)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.y;
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane);
    const float4 g1 = __ldg(reinterpret_cast<const float4 *>(gamma) + 32 + lane);
    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(beta) + lane);
    const float4 b1 = __ldg(reinterpret_cast<const float4 *>(beta) + 32 + lane);
    const float m0 = __ldg(mean + n * kGnGroups + (lane >> 1)), r0 = __ldg(rstd + n * kGnGroups + (lane >> 1));
    const float m1 = __ldg(mean + n * kGnGroups + 16 + (lane >> 1)), r1 = __ldg(rstd + n * kGnGroups + 16 + (lane >> 1));
    const float *img = y + static_cast<size_t>(n) * tokens * kGnC;
    float *dst = out + static_cast<size_t>(n) * out_batch_stride;
    for (int t = blockIdx.x * 8 + warp; t < tokens; t += gridDim.x * 8) {
        const float4 *row = reinterpret_cast<const float4 *>(img + static_cast<size_t>(t) * kGnC);
        const float4 a = __ldg(row + lane), b = __ldg(row + 32 + lane);
        float4 *o = reinterpret_cast<float4 *>(dst + static_cast<size_t>(t) * kGnC);
        o[lane] = make_float4((a.x - m0) * r0 * g0.x + b0.x, (a.y - m0) * r0 * g0.y + b0.y,
                              (a.z - m0) * r0 * g0.z + b0.z, (a.w - m0) * r0 * g0.w + b0.w);
        o[32 + lane] = make_float4((b.x - m1) * r1 * g1.x + b1.x, (b.y - m1) * r1 * g1.y + b1.y,
                                   (b.z - m1) * r1 * g1.z + b1.z, (b.w - m1) * r1 * g1.w + b1.w);
    }
}

// backward pass 1: per (n, g) sums of dy*gamma and dy*gamma*xhat (fp64 atomics), per-channel dgamma / dbeta (fp32 atomics)
static __global__ void __launch_bounds__(256) groupnorm_tokens_bwd_stats_kernel(const float *__restrict__ grad_out,
                                                                         long long grad_batch_stride,
                                                                         const float *__restrict__ y,
                                                                         const float *__restrict__ mean,
                                                                         const float *__restrict__ rstd,
                                                                         const float *__restrict__ gamma,
                                                                         double *__restrict__ sums, float *__restrict__ dgamma,
                                                                         float *__restrict__ dbeta, int tokens)
{
    __shared__ float s_dg[8][kGnC], s_db[8][kGnC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.y;
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane);
    const float4 g1 = __ldg(reinterpret_cast<const float4 *>(gamma) + 32 + lane);
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float m0 = __ldg(mean + n * kGnGroups + (lane >> 1)), r0 = __ldg(rstd + n * kGnGroups + (lane >> 1));
    const float m1 = __ldg(mean + n * kGnGroups + 16 + (lane >> 1)), r1 = __ldg(rstd + n * kGnGroups + 16 + (lane >> 1));
    const float *img = y + static_cast<size_t>(n) * tokens * kGnC;
    const float *gimg = grad_out + static_cast<size_t>(n) * grad_batch_stride;
    float dg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, db[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float sa0 = 0.f, sb0 = 0.f, sa1 = 0.f, sb1 = 0.f;
    for (int t = blockIdx.x * 8 + warp; t < tokens; t += gridDim.x * 8) {
        const float4 *row = reinterpret_cast<const float4 *>(img + static_cast<size_t>(t) * kGnC);
        const float4 *grow = reinterpret_cast<const float4 *>(gimg + static_cast<size_t>(t) * kGnC);
        const float4 a = __ldg(row + lane), b = __ldg(row + 32 + lane);
        const float4 da = __ldg(grow + lane), dc = __ldg(grow + 32 + lane);
        const float xh[8] = {(a.x - m0) * r0, (a.y - m0) * r0, (a.z - m0) * r0, (a.w - m0) * r0,
                             (b.x - m1) * r1, (b.y - m1) * r1, (b.z - m1) * r1, (b.w - m1) * r1};
        const float dy[8] = {da.x, da.y, da.z, da.w, dc.x, dc.y, dc.z, dc.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dg[i] += dy[i] * xh[i];
            db[i] += dy[i];
            const float g = dy[i] * gm[i];
            if (i < 4) { sa0 += g; sb0 += g * xh[i]; } else { sa1 += g; sb1 += g * xh[i]; }
        }
    }
    // group sums: lanes 2g and 2g+1 hold the two float4 slots of group g (first half) and 16+g (second half)
    sa0 += __shfl_xor_sync(0xffffffffu, sa0, 1);
    sb0 += __shfl_xor_sync(0xffffffffu, sb0, 1);
    sa1 += __shfl_xor_sync(0xffffffffu, sa1, 1);
    sb1 += __shfl_xor_sync(0xffffffffu, sb1, 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s_dg[warp][lane * 4 + i] = dg[i];
        s_db[warp][lane * 4 + i] = db[i];
        s_dg[warp][128 + lane * 4 + i] = dg[4 + i];
        s_db[warp][128 + lane * 4 + i] = db[4 + i];
    }
    __shared__ float s_grp[8][kGnGroups][2];
    if ((lane & 1) == 0) {
        s_grp[warp][lane >> 1][0] = sa0;
        s_grp[warp][lane >> 1][1] = sb0;
        s_grp[warp][16 + (lane >> 1)][0] = sa1;
        s_grp[warp][16 + (lane >> 1)][1] = sb1;
    }
    __syncthreads();
    const int f = threadIdx.x;
    float ga = 0.f, gb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        ga += s_dg[w][f];
        gb += s_db[w][f];
    }
    atomicAdd(dgamma + f, ga);
    atomicAdd(dbeta + f, gb);
    if (threadIdx.x < 64) {
        const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += static_cast<double>(s_grp[w][g][which]);
        atomicAdd(sums + (static_cast<size_t>(n) * kGnGroups + g) * 2 + which, acc);
    }
}

// backward pass 2: grad_y = rstd * (dy*gamma - c1 - xhat * c2)
static __global__ void __launch_bounds__(256) groupnorm_tokens_bwd_apply_kernel(const float *__restrict__ grad_out,
                                                                         long long grad_batch_stride,
                                                                         const float *__restrict__ y,
                                                                         const float *__restrict__ mean,
                                                                         const float *__restrict__ rstd,
                                                                         const float *__restrict__ c1,
                                                                         const float *__restrict__ c2,
                                                                         const float *__restrict__ gamma,
                                                                         float *__restrict__ grad_y, int tokens)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.y;
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane);
    const float4 g1 = __ldg(reinterpret_cast<const float4 *>(gamma) + 32 + lane);
    const int ga = n * kGnGroups + (lane >> 1), gb = ga + 16;
    const float m0 = __ldg(mean + ga), r0 = __ldg(rstd + ga), p0 = __ldg(c1 + ga), q0 = __ldg(c2 + ga);
    const float m1 = __ldg(mean + gb), r1 = __ldg(rstd + gb), p1 = __ldg(c1 + gb), q1 = __ldg(c2 + gb);
    const float *img = y + static_cast<size_t>(n) * tokens * kGnC;
    const float *gimg = grad_out + static_cast<size_t>(n) * grad_batch_stride;
    float *dst = grad_y + static_cast<size_t>(n) * tokens * kGnC;
    for (int t = blockIdx.x * 8 + warp; t < tokens; t += gridDim.x * 8) {
        const float4 *row = reinterpret_cast<const float4 *>(img + static_cast<size_t>(t) * kGnC);
        const float4 *grow = reinterpret_cast<const float4 *>(gimg + static_cast<size_t>(t) * kGnC);
        const float4 a = __ldg(row + lane), b = __ldg(row + 32 + lane);
        const float4 da = __ldg(grow + lane), dc = __ldg(grow + 32 + lane);
        float4 *o = reinterpret_cast<float4 *>(dst + static_cast<size_t>(t) * kGnC);
        o[lane] = make_float4(r0 * (da.x * g0.x - p0 - (a.x - m0) * r0 * q0), r0 * (da.y * g0.y - p0 - (a.y - m0) * r0 * q0),
                              r0 * (da.z * g0.z - p0 - (a.z - m0) * r0 * q0), r0 * (da.w * g0.w - p0 - (a.w - m0) * r0 * q0));
        o[32 + lane] = make_float4(r1 * (dc.x * g1.x - p1 - (b.x - m1) * r1 * q1), r1 * (dc.y * g1.y - p1 - (b.y - m1) * r1 * q1),
                                   r1 * (dc.z * g1.z - p1 - (b.z - m1) * r1 * q1), r1 * (dc.w * g1.w - p1 - (b.w - m1) * r1 * q1));
    }
}

// pos[t, c], t = i * W + j: c < F: y part, c >= F: x part (position_encoding.py:33-52 with mask = None, normalize = True):
//   e = (i + 1) / (H + 1e-6) * scale   (resp. (j + 1) / (W + 1e-6));  v = e / temperature^(2 * (k / 2) / F), k = c mod F;
//   sin for even k, cos for odd k.  Same fp32 operation order as torch (true divisions, no reciprocal multiplies).
static __global__ void sine_pos_embed_kernel(float *__restrict__ out, int H, int W, int F, float temperature, float scale,
                                      int normalize)
{
    const int C = 2 * F;
    const size_t total = static_cast<size_t>(H) * W * C;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        const int t = static_cast<int>(idx / C);
        const int i = t / W, j = t - i * W;
        const bool ypart = c < F;
        const int k = ypart ? c : c - F;
        float e = static_cast<float>((ypart ? i : j) + 1);
        if (normalize) {
            const float last = static_cast<float>(ypart ? H : W);
            e = e / (last + 1e-6f) * scale;
        }
        const float v = e / powf(temperature, static_cast<float>(2 * (k / 2)) / static_cast<float>(F));
        out[idx] = (k & 1) ? cosf(v) : sinf(v);
    }
}

}  // namespace bm2f

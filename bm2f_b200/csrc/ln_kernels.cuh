// Fused residual-add + LayerNorm for the encoder layer (reference: msdeformattn.py:125-126 and 115-120:
// `src = src + dropout(src2); src = norm(src)` with dropout 0 — all configs that use this decoder set
// DROPOUT 0.0), d_model = 256.
//
//   forward : z = x + r;  y = (z - mean(z)) * rstd(z) * gamma + beta       writes z (LN input, kept for backward),
//             y, mean, rstd                                                 one pass: read 2, write 2 tensors
//   backward: dz = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma, xhat = (z - mean) * rstd
//             dgamma += sum_rows dy * xhat, dbeta += sum_rows dy;  dx = dr = dz
//
// One warp per row: a lane owns 8 of the 256 features (two float4), statistics by xor-shuffles; HBM-bound.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bm2f {

constexpr int kLnC = 256;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static __global__ void __launch_bounds__(256) add_layernorm_fwd_kernel(const float *__restrict__ x, const float *__restrict__ r,
                                                                const float *__restrict__ gamma,
                                                                const float *__restrict__ beta, float eps,
                                                                float *__restrict__ z, float *__restrict__ y,
                                                                float *__restrict__ mean, float *__restrict__ rstd,
                                                                int rows)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane);
    const float4 g1 = __ldg(reinterpret_cast<const float4 *>(gamma) + 32 + lane);
    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(beta) + lane);
    const float4 b1 = __ldg(reinterpret_cast<const float4 *>(beta) + 32 + lane);
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
        const float4 *xr = reinterpret_cast<const float4 *>(x + static_cast<size_t>(row) * kLnC);
        const float4 *rr = reinterpret_cast<const float4 *>(r + static_cast<size_t>(row) * kLnC);
        float4 a = __ldg(xr + lane), c = __ldg(xr + 32 + lane);
        const float4 ra = __ldg(rr + lane), rc = __ldg(rr + 32 + lane);
        a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
        c.x += rc.x; c.y += rc.y; c.z += rc.z; c.w += rc.w;
        const float mu = warp_sum(a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w) * (1.f / kLnC);
        const float d0 = a.x - mu, d1 = a.y - mu, d2 = a.z - mu, d3 = a.w - mu;
        const float d4 = c.x - mu, d5 = c.y - mu, d6 = c.z - mu, d7 = c.w - mu;
        const float var = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3 + d4 * d4 + d5 * d5 + d6 * d6 + d7 * d7) *
                          (1.f / kLnC);
        const float rs = rsqrtf(var + eps);
        float4 *zr = reinterpret_cast<float4 *>(z + static_cast<size_t>(row) * kLnC);
        float4 *yr = reinterpret_cast<float4 *>(y + static_cast<size_t>(row) * kLnC);
        zr[lane] = a;
        zr[32 + lane] = c;
        yr[lane] = make_float4(d0 * rs * g0.x + b0.x, d1 * rs * g0.y + b0.y, d2 * rs * g0.z + b0.z, d3 * rs * g0.w + b0.w);
        yr[32 + lane] = make_float4(d4 * rs * g1.x + b1.x, d5 * rs * g1.y + b1.y, d6 * rs * g1.z + b1.z, d7 * rs * g1.w + b1.w);
        if (lane == 0) {
            mean[row] = mu;
            rstd[row] = rs;
        }
    }
}

static __global__ void __launch_bounds__(256) add_layernorm_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ z,
                                                                const float *__restrict__ mean,
                                                                const float *__restrict__ rstd,
                                                                const float *__restrict__ gamma, float *__restrict__ dz,
                                                                float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                                int rows)
{
    __shared__ float s_dg[8][kLnC], s_db[8][kLnC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane);
    const float4 g1 = __ldg(reinterpret_cast<const float4 *>(gamma) + 32 + lane);
    float dg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, db[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
        const float4 *dr = reinterpret_cast<const float4 *>(dy + static_cast<size_t>(row) * kLnC);
        const float4 *zr = reinterpret_cast<const float4 *>(z + static_cast<size_t>(row) * kLnC);
        const float4 da = __ldg(dr + lane), dc = __ldg(dr + 32 + lane);
        const float4 za = __ldg(zr + lane), zc = __ldg(zr + 32 + lane);
        const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
        const float xh[8] = {(za.x - mu) * rs, (za.y - mu) * rs, (za.z - mu) * rs, (za.w - mu) * rs,
                             (zc.x - mu) * rs, (zc.y - mu) * rs, (zc.z - mu) * rs, (zc.w - mu) * rs};
        const float dyv[8] = {da.x, da.y, da.z, da.w, dc.x, dc.y, dc.z, dc.w};
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float g[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            g[i] = dyv[i] * gm[i];
            s1 += g[i];
            s2 += g[i] * xh[i];
            dg[i] += dyv[i] * xh[i];
            db[i] += dyv[i];
        }
        const float c1 = warp_sum(s1) * (1.f / kLnC), c2 = warp_sum(s2) * (1.f / kLnC);
        float4 *out = reinterpret_cast<float4 *>(dz + static_cast<size_t>(row) * kLnC);
        out[lane] = make_float4(rs * (g[0] - c1 - xh[0] * c2), rs * (g[1] - c1 - xh[1] * c2),
                                rs * (g[2] - c1 - xh[2] * c2), rs * (g[3] - c1 - xh[3] * c2));
        out[32 + lane] = make_float4(rs * (g[4] - c1 - xh[4] * c2), rs * (g[5] - c1 - xh[5] * c2),
                                     rs * (g[6] - c1 - xh[6] * c2), rs * (g[7] - c1 - xh[7] * c2));
    }
    // per-feature partial sums of this block -> one atomic per feature per block
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s_dg[warp][lane * 4 + i] = dg[i];
        s_db[warp][lane * 4 + i] = db[i];
        s_dg[warp][128 + lane * 4 + i] = dg[4 + i];
        s_db[warp][128 + lane * 4 + i] = db[4 + i];
    }
    __syncthreads();
    const int f = threadIdx.x;   // 256 threads <-> 256 features
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        a += s_dg[w][f];
        b += s_db[w][f];
    }
    atomicAdd(dgamma + f, a);
    atomicAdd(dbeta + f, b);
}


// value.masked_fill(padding_mask[..., None], 0) without touching unmasked rows (reference:
// ops/modules/ms_deform_attn.py:99-100; Mask2Former always passes an all-False mask, msdeformattn.py:62): one warp per
// row reads the mask byte and only masked rows are written.  Used in place on tensors the caller owns.
static __global__ void __launch_bounds__(256) zero_masked_rows_kernel(float *__restrict__ x, const unsigned char *__restrict__ mask,
                                                               int rows, int channels)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
        if (mask[row]) {
            float4 *r = reinterpret_cast<float4 *>(x + static_cast<size_t>(row) * channels);
            for (int c = lane; c < channels / 4; c += 32) r[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

}  // namespace bm2f

// Any-shape kernels (any D / L / P, float / double): coverage path for everything the D = 32
// fast path does not take (the reference's own test sweeps D in {2,30,32,64,71,1025,2048,3096}
// in float64, ops/test.py:84-89).  One warp per (image, query, head); lanes stride the channels.
#pragma once

#include "msda_common.cuh"

namespace bm2f {

struct GenericParams {
    const void *value;
    const int64_t *shapes;
    const int64_t *start;
    const void *loc;
    const void *attn;
    const void *grad_out;
    void *out;
    void *grad_value;
    void *grad_loc;
    void *grad_attn;
    int N, S, M, D, L, Lq, P;
};

template <typename T>
struct Taps {
    long long pix[4];
    T wgt[4], dx[4], dy[4];
    int n;
};

// Same rule set as make_footprint(), in the tensor's own precision (float or double).
template <typename T>
__device__ __forceinline__ Taps<T> resolve_taps(T loc_x, T loc_y, int H, int W)
{
    Taps<T> t;
    t.n = 0;
    const T h_im = loc_y * static_cast<T>(H) - static_cast<T>(0.5);
    const T w_im = loc_x * static_cast<T>(W) - static_cast<T>(0.5);
    if (!(h_im > static_cast<T>(-1) && w_im > static_cast<T>(-1) && h_im < static_cast<T>(H) &&
          w_im < static_cast<T>(W)))
        return t;
    const T fy = floor(h_im), fx = floor(w_im);
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const T lh = h_im - fy, lw = w_im - fx;
    const T hh = static_cast<T>(1) - lh, hw = static_cast<T>(1) - lw;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dy = k >> 1, dx = k & 1;
        const int y = y0 + dy, x = x0 + dx;
        if (y < 0 || y > H - 1 || x < 0 || x > W - 1) continue;
        const T wy = dy ? lh : hh, wx = dx ? lw : hw;
        t.pix[t.n] = static_cast<long long>(y) * W + x;
        t.wgt[t.n] = wy * wx;
        t.dx[t.n] = dx ? wy : -wy;
        t.dy[t.n] = dy ? wx : -wx;
        ++t.n;
    }
    return t;
}

template <typename T>
__global__ void __launch_bounds__(256) msda_fwd_generic_kernel(const GenericParams p)
{
    const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long total = static_cast<long long>(p.N) * p.Lq * p.M;
    if (wid >= total) return;
    const int m = static_cast<int>(wid % p.M);
    const long long bq = wid / p.M;
    const int b = static_cast<int>(bq / p.Lq);
    const T *value = static_cast<const T *>(p.value);
    const T *loc = static_cast<const T *>(p.loc) + wid * p.L * p.P * 2;
    const T *attn = static_cast<const T *>(p.attn) + wid * p.L * p.P;
    T *out = static_cast<T *>(p.out) + wid * p.D;
    const long long row = static_cast<long long>(p.M) * p.D;

    for (int c0 = 0; c0 < p.D; c0 += 32) {
        const int c = c0 + lane;
        T acc = 0;
        for (int l = 0; l < p.L; ++l) {
            const bool lvl_ok = level_in_bounds(p.start[l], p.shapes[2 * l], p.shapes[2 * l + 1], p.S);
            const int H = lvl_ok ? static_cast<int>(p.shapes[2 * l]) : 0, W = lvl_ok ? static_cast<int>(p.shapes[2 * l + 1]) : 0;
            const T *vl = value + (static_cast<long long>(b) * p.S + (lvl_ok ? p.start[l] : 0)) * row + static_cast<long long>(m) * p.D;
            for (int pt = 0; pt < p.P; ++pt) {
                const int i = l * p.P + pt;
                const Taps<T> t = resolve_taps<T>(loc[2 * i], loc[2 * i + 1], H, W);
                const T a = attn[i];
                if (c < p.D)
                    for (int k = 0; k < t.n; ++k) acc += t.wgt[k] * a * vl[t.pix[k] * row + c];
            }
        }
        if (c < p.D) out[c] = acc;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) msda_bwd_generic_kernel(const GenericParams p)
{
    const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long total = static_cast<long long>(p.N) * p.Lq * p.M;
    if (wid >= total) return;
    const int m = static_cast<int>(wid % p.M);
    const long long bq = wid / p.M;
    const int b = static_cast<int>(bq / p.Lq);
    const T *value = static_cast<const T *>(p.value);
    T *grad_value = static_cast<T *>(p.grad_value);
    const T *loc = static_cast<const T *>(p.loc) + wid * p.L * p.P * 2;
    const T *attn = static_cast<const T *>(p.attn) + wid * p.L * p.P;
    const T *go = static_cast<const T *>(p.grad_out) + wid * p.D;
    T *grad_loc = static_cast<T *>(p.grad_loc) + wid * p.L * p.P * 2;
    T *grad_attn = static_cast<T *>(p.grad_attn) + wid * p.L * p.P;
    const long long row = static_cast<long long>(p.M) * p.D;

    for (int l = 0; l < p.L; ++l) {
        const bool lvl_ok = level_in_bounds(p.start[l], p.shapes[2 * l], p.shapes[2 * l + 1], p.S);
        const int H = lvl_ok ? static_cast<int>(p.shapes[2 * l]) : 0, W = lvl_ok ? static_cast<int>(p.shapes[2 * l + 1]) : 0;
        const long long base = (static_cast<long long>(b) * p.S + (lvl_ok ? p.start[l] : 0)) * row + static_cast<long long>(m) * p.D;
        for (int pt = 0; pt < p.P; ++pt) {
            const int i = l * p.P + pt;
            const Taps<T> t = resolve_taps<T>(loc[2 * i], loc[2 * i + 1], H, W);
            const T a = attn[i];
            T s_val = 0, s_dx = 0, s_dy = 0;
            for (int c = lane; c < p.D; c += 32) {
                const T g = go[c];
                for (int k = 0; k < t.n; ++k) {
                    const long long e = base + t.pix[k] * row + c;
                    const T d = g * value[e];
                    atomicAdd(grad_value + e, t.wgt[k] * a * g);
                    s_val += t.wgt[k] * d;
                    s_dx += t.dx[k] * d;
                    s_dy += t.dy[k] * d;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s_val += __shfl_xor_sync(0xffffffffu, s_val, o);
                s_dx += __shfl_xor_sync(0xffffffffu, s_dx, o);
                s_dy += __shfl_xor_sync(0xffffffffu, s_dy, o);
            }
            if (lane == 0) {
                grad_attn[i] = s_val;
                grad_loc[2 * i] = static_cast<T>(W) * a * s_dx;
                grad_loc[2 * i + 1] = static_cast<T>(H) * a * s_dy;
            }
        }
    }
}

}  // namespace bm2f

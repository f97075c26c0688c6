// Shared device helpers for the sm_100a multi-scale deformable attention kernels.
//
// Behaviour contract restated from the reference kernels
// (/root/reference/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh):
//   * pixel coordinates  h_im = y*H - 0.5, w_im = x*W - 0.5                     (.cuh:290-291)
//   * a point contributes only if -1 < h_im < H and -1 < w_im < W               (.cuh:293)
//   * each of the four corners is dropped on its own when outside the level    (.cuh:60-83)
//   * weights hh*hw, hh*lw, lh*hw, lh*lw with lh = h_im - floor(h_im)           (.cuh:43-50,85)
//   * grad_loc is scaled by W (x) and H (y), grad_attn = <grad_out, sample>     (.cuh:160-163)
// None of the reference's code is reused; the mapping of work to threads is different
// (see DESIGN.md).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bm2f {

constexpr int kMaxLevels = 16;

// ------------------------------------------------------------------------------------------
// mbarrier / TMA (cp.async.bulk.tensor) wrappers — raw PTX, sm_90+/sm_100a.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// generic-proxy writes to shared memory become ordered with later async-proxy (TMA) accesses of the same bytes
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 2-D tiled TMA load: box -> shared memory, completion counted on `bar` in bytes.
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// request a 2-D box into L2 only (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap *map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ------------------------------------------------------------------------------------------
// Vector loads of VEC consecutive channels, returned as float.  Read-only path (ld.global.nc).
// red_add: vector reductions into grad_value.  They never alias anything the kernels read, so
// the asm carries no memory clobber and corner loads may be scheduled across them.
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
struct VecIO;

template <>
struct VecIO<float, 4> {
    __device__ static __forceinline__ void load(const float *p, float (&v)[4])
    {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ static __forceinline__ void store(float *p, const float (&v)[4])
    {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    // 128-bit vector reduction: REDG.E.ADD.F32x4 (one 16-byte atomic per lane, 8 lanes = 1 line)
    __device__ static __forceinline__ void red_add(float *p, const float (&v)[4])
    {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]),
                     "f"(v[2]), "f"(v[3]));
    }
};
template <>
struct VecIO<float, 2> {
    __device__ static __forceinline__ void load(const float *p, float (&v)[2])
    {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        v[0] = t.x; v[1] = t.y;
    }
    __device__ static __forceinline__ void store(float *p, const float (&v)[2])
    {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    }
    __device__ static __forceinline__ void red_add(float *p, const float (&v)[2])
    {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]));
    }
};
template <>
struct VecIO<float, 1> {
    __device__ static __forceinline__ void load(const float *p, float (&v)[1]) { v[0] = __ldg(p); }
    __device__ static __forceinline__ void store(float *p, const float (&v)[1]) { *p = v[0]; }
    __device__ static __forceinline__ void red_add(float *p, const float (&v)[1])
    {
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]));
    }
};


// 256-bit predicated read-only load (LDG.E.ENL2.256.CONSTANT on sm_100a): 8 consecutive floats.
// v must be zero-initialised by the caller; nothing is fetched when pred is false.
__device__ __forceinline__ void ldg256_pred(const float *p, bool pred, float (&v)[8])
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "setp.ne.b32 P1, %9, 0;\n"
        "@P1 ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
        "}\n"
        : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7])
        : "l"(p), "r"(static_cast<int>(pred)));
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}

template <>
struct VecIO<__nv_bfloat16, 4> {
    __device__ static __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[4])
    {
        const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
        v[0] = bf16lo(t.x); v[1] = bf16hi(t.x); v[2] = bf16lo(t.y); v[3] = bf16hi(t.y);
    }
    __device__ static __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[4])
    {
        *reinterpret_cast<uint2 *>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
    // REDG.E.ADD.BF16x2 vector form (v2.bf16x2 = 4 channels)
    __device__ static __forceinline__ void red_add(__nv_bfloat16 *p, const float (&v)[4])
    {
        asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1, %2};" ::"l"(p), "r"(pack_bf16x2(v[0], v[1])),
                     "r"(pack_bf16x2(v[2], v[3])));
    }
};
template <>
struct VecIO<__nv_bfloat16, 2> {
    __device__ static __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[2])
    {
        const uint32_t t = __ldg(reinterpret_cast<const uint32_t *>(p));
        v[0] = bf16lo(t); v[1] = bf16hi(t);
    }
    __device__ static __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[2])
    {
        *reinterpret_cast<uint32_t *>(p) = pack_bf16x2(v[0], v[1]);
    }
    __device__ static __forceinline__ void red_add(__nv_bfloat16 *p, const float (&v)[2])
    {
        asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(p), "r"(pack_bf16x2(v[0], v[1])));
    }
};

// A level whose slice [start, start + H*W) does not lie inside the S pixels of `value` (a level table that disagrees
// with the tensor; the reference asserts the sum on the host, ops/modules/ms_deform_attn.py:96) is treated as EMPTY by
// every kernel: H = W = 0 drops all its corners, so nothing is read or written outside value / grad_value.
__device__ __forceinline__ bool level_in_bounds(long long start, long long H, long long W, long long S)
{
    return H >= 0 && W >= 0 && start >= 0 && H <= S && W <= S && start + H * W <= S;
}

// ------------------------------------------------------------------------------------------
// One bilinear footprint: the four corner weights / validity of a sampling point.
// ------------------------------------------------------------------------------------------
struct Footprint {
    float lh, lw, hh, hw;  // fractional parts and complements
    int y0, x0;            // top-left corner (may be -1)
    bool ok[4];            // corner (dy,dx) = (k>>1, k&1) inside the level AND point in range
    bool in_range;         // the reference's -1 < h_im < H && -1 < w_im < W test
};

__device__ __forceinline__ Footprint make_footprint(float loc_x, float loc_y, int H, int W, float Hf, float Wf)
{
    Footprint f;
    const float h_im = fmaf(loc_y, Hf, -0.5f);
    const float w_im = fmaf(loc_x, Wf, -0.5f);
    f.in_range = (h_im > -1.f) && (w_im > -1.f) && (h_im < Hf) && (w_im < Wf);
    const float fy = floorf(h_im), fx = floorf(w_im);
    f.y0 = static_cast<int>(fy);
    f.x0 = static_cast<int>(fx);
    f.lh = h_im - fy;
    f.lw = w_im - fx;
    f.hh = 1.f - f.lh;
    f.hw = 1.f - f.lw;
    const bool y0ok = f.in_range && (f.y0 >= 0);
    const bool y1ok = f.in_range && (f.y0 + 1 <= H - 1);
    const bool x0ok = f.x0 >= 0;
    const bool x1ok = f.x0 + 1 <= W - 1;
    f.ok[0] = y0ok && x0ok;
    f.ok[1] = y0ok && x1ok;
    f.ok[2] = y1ok && x0ok;
    f.ok[3] = y1ok && x1ok;
    return f;
}

}  // namespace bm2f

// C ABI of libbm2f_msda.so (see include/bm2f_msda.h): argument checks, variant selection,
// TMA tensor-map encoding and kernel launches.  No torch types here; the torch-facing shim
// (msda_torch_ext.cpp) and ctypes callers sit on top of this file.
//
// Replaces the reference host wrappers ms_deform_attn_cuda_forward/backward and the launchers
// ms_deformable_im2col_cuda / ms_deformable_col2im_cuda
// (/root/reference/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_attn_cuda.cu:25-158,
//  ms_deform_im2col_cuda.cuh:928-1332).
#include "../../include/bm2f_msda.h"
#include "api_common.cuh"

#include <type_traits>

#include "msda_fast.cuh"
#include "msda_generic.cuh"

#ifndef BM2F_BWD_SORTED_DEFAULT
#define BM2F_BWD_SORTED_DEFAULT 1     // anchor-sorted backward for large encoder shapes (8 % faster at cfg 2, profiles/r02_*)
#endif

namespace bm2f {
namespace host {
// msda_bwd_sorted.cu
bool bwd_sorted_eligible(const Dims &d, int dtype, const bm2f_msda_tuning_t &t);
int run_bwd_sorted(FastParams p, const Dims &d, const bm2f_msda_tuning_t &t, bool fused, cudaStream_t st);
}  // namespace host
}  // namespace bm2f

namespace {

using namespace bm2f;
using namespace bm2f::host;

// Backward kernel choice (tuning.bwd): 2 = anchor-sorted or fail, 1 = per-corner, 0 = default.
int choose_bwd_sorted(const Dims &d, int dtype, const bm2f_msda_tuning_t &t, bool *sorted)
{
    const bool ok = bwd_sorted_eligible(d, dtype, t);
    if (t.bwd == 2 && !ok)
        return fail(BM2F_ERR_UNSUPPORTED, "anchor-sorted backward needs float32, D=32, M=8, P=4, L<=4, num_query == spatial_size "
                    "and order == 0 (got D=%d M=%d P=%d L=%d Lq=%d S=%d dtype=%d)", d.D, d.M, d.P, d.L, d.Lq, d.S, dtype);
    // default: only where it was measured faster — enough (image, query) rows to fill the machine with whole chunks:
    // cfg 2 at 2 / 3 / 4 images (43 008+ rows) 0.338 / 0.490 / 0.649 ms vs 0.369 / 0.540 / 0.713 per-corner; a single
    // 512^2 image (5 376 rows) is launch / tail bound and stays on the per-corner kernel (0.065 vs 0.100 ms)
    const bool big = static_cast<long long>(d.N) * d.Lq >= 32768;
    *sorted = ok && t.bwd != 1 && (t.bwd == 2 || (BM2F_BWD_SORTED_DEFAULT && big));
    return BM2F_OK;
}

// ---- fast-path variant table ----------------------------------------------------------------
constexpr int kNWarp = 16;

template <typename T, int VEC, int L_, int SW, bool TMA, int CPS>
cudaError_t launch_fast(bool bwd, bool merge, bool wide, const FastParams &p, const CUtensorMap &ml, const CUtensorMap &mw,
                        int grid, cudaStream_t st)
{
    constexpr int G = (SW < kNWarp) ? kNWarp / SW : 1;
    constexpr int threads = (kNWarp + (TMA ? 1 : 0)) * 32;
    if (bwd && merge && VEC == 4)
        msda_bwd_fast_kernel<T, VEC, L_, 4, SW, kNWarp, G, TMA, CPS, true><<<grid, threads, 0, st>>>(p, ml, mw);
    else if (bwd)
        msda_bwd_fast_kernel<T, VEC, L_, 4, SW, kNWarp, G, TMA, CPS, false><<<grid, threads, 0, st>>>(p, ml, mw);
    else if (wide && std::is_same<T, float>::value && VEC == 4)
        msda_fwd_fast256_kernel<L_, 4, SW, kNWarp, G, TMA, CPS><<<grid, threads, 0, st>>>(p, ml, mw);
    else
        msda_fwd_fast_kernel<T, VEC, L_, 4, SW, kNWarp, G, TMA, CPS><<<grid, threads, 0, st>>>(p, ml, mw);
    return cudaGetLastError();
}

// Fused-prologue variants (softmax + location arithmetic in the kernel): default tile shape only.
template <typename T, int L_, bool TMA>
cudaError_t launch_fused(bool bwd, const FastParams &p, const CUtensorMap &ml, const CUtensorMap &mw, int grid,
                         cudaStream_t st)
{
    constexpr int SW = 32, CPS = 1, G = 1;
    constexpr int threads = (kNWarp + (TMA ? 2 : 0)) * 32;      // + TMA producer warp + prologue warp
    if (bwd)
        msda_bwd_fast_kernel<T, 4, L_, 4, SW, kNWarp, G, TMA, CPS, false, true><<<grid, threads, 0, st>>>(p, ml, mw);
    else
        msda_fwd_fast_kernel<T, 4, L_, 4, SW, kNWarp, G, TMA, CPS, true><<<grid, threads, 0, st>>>(p, ml, mw);
    return cudaGetLastError();
}

// Forward with geometry warps (msda_fwd_geo_kernel): 16 consumer warps + TMA warp + kGeoWarps geometry warps.
constexpr int kGeoWarps = 2, kGeoStages = 4;
template <typename T, int L_, bool FUSED, bool WIDE = false, int CPS = 1, int NWC = kNWarp, int LEAN = 0, int NGEO = kGeoWarps>
int launch_geo(const FastParams &p, const CUtensorMap &ml, const CUtensorMap &mw, int grid, cudaStream_t st)
{
    constexpr int SW = 32;
    constexpr int threads = (NWC + 1 + NGEO) * 32;
    static_assert(threads <= 1024, "consumer + TMA + geometry warps in one CTA");
    constexpr int smem = GeoRing<L_, 4, SW, kGeoStages>::kBytes;
    int rc = ensure_dynamic_smem<&msda_fwd_geo_kernel<T, L_, 4, SW, NWC, NGEO, kGeoStages, FUSED, WIDE, CPS, LEAN>>(
        smem, "cudaFuncSetAttribute(geometry-warp forward smem)");
    if (rc) return rc;
    msda_fwd_geo_kernel<T, L_, 4, SW, NWC, NGEO, kGeoStages, FUSED, WIDE, CPS, LEAN><<<grid, threads, smem, st>>>(p, ml, mw);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch msda_fwd_geo_kernel");
    count_launch(1);
    return BM2F_OK;
}

struct FastChoice {
    int vec, sw, tma, cps, merge, wide;
};

#define BM2F_CASE(T, VEC, L_, SW, TMA, CPS)                                              \
    if (c.vec == VEC && c.sw == SW && c.tma == TMA && c.cps == CPS) {                    \
        *found = true;                                                                   \
        return launch_fast<T, VEC, L_, SW, (TMA != 0), CPS>(bwd, c.merge != 0, c.wide != 0, p, ml, mw, grid, st);   \
    }

// Variant grid for the Mask2Former shape (L = 3, P = 4, fp32): the product library carries the 128-bit kernels for
// strips of 16 / 32 queries; the rest of the round-1 sweep grid is compiled with -DBM2F_SWEEP only.
cudaError_t dispatch_f32_l3(const FastChoice &c, bool bwd, const FastParams &p, const CUtensorMap &ml,
                            const CUtensorMap &mw, int grid, cudaStream_t st, bool *found)
{
#define BM2F_ROW(VEC, SW)        \
    BM2F_CASE(float, VEC, 3, SW, 0, 1) \
    BM2F_CASE(float, VEC, 3, SW, 0, 2) \
    BM2F_CASE(float, VEC, 3, SW, 1, 1) \
    BM2F_CASE(float, VEC, 3, SW, 1, 2)
    BM2F_ROW(4, 16) BM2F_ROW(4, 32)
#ifdef BM2F_SWEEP       // tools/sweep.py grid (BM2F_SWEEP=1 python -m bm2f_b200.build): narrower loads, 8-query strips
    BM2F_ROW(4, 8)
    BM2F_ROW(2, 8) BM2F_ROW(2, 16) BM2F_ROW(2, 32)
    BM2F_ROW(1, 8) BM2F_ROW(1, 16) BM2F_ROW(1, 32)
#endif
#undef BM2F_ROW
    return cudaSuccess;
}

// Other level counts and bf16: the default shape of each staging mode only.
template <typename T, int L_>
cudaError_t dispatch_default(const FastChoice &c, bool bwd, const FastParams &p, const CUtensorMap &ml,
                             const CUtensorMap &mw, int grid, cudaStream_t st, bool *found)
{
    BM2F_CASE(T, 4, L_, 32, 0, 1)
    BM2F_CASE(T, 4, L_, 32, 1, 1)
    return cudaSuccess;
}
#undef BM2F_CASE

// true when the D = 32 fast kernels can take the problem
bool fast_eligible(const Dims &d, int dtype, const bm2f_msda_tuning_t &t, const void *value, const void *io,
                   const void *loc, const void *attn)
{
    if (t.force_generic) return false;
    if (d.D != 32 || d.P != 4 || d.M != kHeads) return false;
    if (dtype == BM2F_DTYPE_F32) {
        if (d.L < 1 || d.L > 4) return false;
    } else if (dtype == BM2F_DTYPE_BF16) {
        if (d.L != 3) return false;
    } else {
        return false;
    }
    // 32-bit element offsets inside one image and 32-bit query rows
    if (static_cast<long long>(d.S) * d.M * d.D >= (1ll << 31)) return false;
    if (static_cast<long long>(d.N) * d.Lq >= (1ll << 31)) return false;
    if (!aligned16(value) || !aligned16(io) || !aligned16(loc) || !aligned16(attn)) return false;
    return true;
}

int run_fast(bool bwd, FastParams p, const Dims &d, int dtype, const bm2f_msda_tuning_t &t, cudaStream_t st,
             bool fused = false)
{
    int sms = 0, cc = 0;
    int rc = device_info(&sms, &cc);
    if (rc) return rc;
    if (cc < 10) return fail(BM2F_ERR_CUDA, "this library contains sm_100a code only; device has cc %d.x", cc);

    FastChoice c;
    c.vec = t.vec ? t.vec : 4;
    // vec = 8: 256-bit forward gathers (fp32); the backward scatter has no 256-bit RED and stays at vec = 4
    c.wide = (c.vec == 8 && dtype == BM2F_DTYPE_F32);
    if (c.vec == 8) c.vec = 4;
    // defaults = winner of the round-1 sweep on B200 (profiles/r01_sweep_cfg2.txt)
    c.sw = t.strip_w ? t.strip_w : 32;
    c.tma = t.staging ? (t.staging == 1) : 1;
    c.cps = t.ctas_per_sm ? t.ctas_per_sm : 1;
    // in-warp merging of equal-pixel corners removes 17 % of the REDs but its match/shuffle chain costs
    // more than it saves on B200 (2.76 ms -> 3.53 ms, profiles/r01_sweep_cfg2_b.txt): off unless asked for
    c.merge = (t.merge == 1);
    const bool sweepable = (dtype == BM2F_DTYPE_F32 && d.L == 3);
    if (!sweepable || fused) { c.vec = 4; c.sw = 32; c.cps = 1; c.merge = 0; c.wide = 0; }

    const int grid_max = sms * c.cps;
    int rows = t.rows;
    if (rows <= 0) {
        rows = 32;
        auto jobs = [&](int r) {
            return static_cast<long long>(d.N) * d.M * ((d.Lq + c.sw * r - 1) / (c.sw * r));
        };
        while (rows > 1 && jobs(rows) < 16ll * grid_max) rows >>= 1;   // small problems: more, shorter jobs
    }
    p.rows = rows;
    p.order = t.order;
    // upper bound on the job count (level geometry can only add edge strips); the kernel
    // recomputes the exact number from the device-resident shape table
    const long long est = static_cast<long long>(d.N) * d.M * ((d.Lq + c.sw * rows - 1) / (c.sw * rows) + 2 * d.L);
    const int grid = static_cast<int>(est < grid_max ? est : grid_max);

    CUtensorMap ml, mw;
    memset(&ml, 0, sizeof(ml));
    memset(&mw, 0, sizeof(mw));
    if (c.tma) {
        const int LP = d.L * d.P;
        const uint64_t rows_total = static_cast<uint64_t>(d.N) * d.Lq;
        if ((rc = make_map(&ml, p.loc, rows_total, static_cast<uint64_t>(d.M) * LP * 2, c.sw, LP * 2, 0, p.ld_packed))) return rc;
        if ((rc = make_map(&mw, p.attn, rows_total, static_cast<uint64_t>(d.M) * LP, c.sw, LP, 0, p.ld_packed))) return rc;
    }

    // DEFAULT forward (float32, TMA staging, strip 32, one CTA per SM, plain entry point): geometry warps with lean records
    // (msda_fwd_geo_kernel<.., LEAN>): 28 consumer warps + TMA warp + 3 geometry warps = one full CTA.  The consumers
    // gather unconditionally (dropped corners carry weight 0 and an in-image offset) and are dealt queries round-robin
    // across stages.  Measured at cfg 2 x 16: 0.89 ms vs 0.97 ms for msda_fwd_fast_kernel (tuning.geo = 2), identical bits.
    // DEFAULT forward (float32, TMA staging, strip 32, one CTA per SM, plain entry point): geometry warps with 16-byte
    // records (msda_fwd_geo_kernel<.., LEAN = 2>): 28 consumer warps + TMA warp + 3 geometry warps = one full CTA.  The
    // consumers gather unconditionally from a clamped anchor (dropped rows / columns carry weight 0), one shared load per
    // point, and are dealt queries round-robin across stages.  Measured at cfg 2 x 16: 0.84 ms vs 0.89 ms with 32-byte lean
    // records (tuning.geo = 11) and 0.97 ms for the consumer-lane kernel msda_fwd_fast_kernel (tuning.geo = 2).
    if (!bwd && (t.geo == 0 || t.geo == 17) && c.tma && c.sw == 32 && c.cps == 1 && !c.wide && dtype == BM2F_DTYPE_F32 && !fused) {
        switch (d.L) {
        case 1: return launch_geo<float, 1, false, false, 1, 28, 2, 3>(p, ml, mw, grid, st);
        case 2: return launch_geo<float, 2, false, false, 1, 28, 2, 3>(p, ml, mw, grid, st);
        case 3: return launch_geo<float, 3, false, false, 1, 28, 2, 3>(p, ml, mw, grid, st);
        case 4: return launch_geo<float, 4, false, false, 1, 28, 2, 3>(p, ml, mw, grid, st);
        default: break;
        }
    }
    // fused entry point on the same kernel (A/B, tuning.geo = 21 / 22 / 23: 28 + 3, 26 + 5, 24 + 7 warps): the geometry
    // warps also do the softmax and the location arithmetic
    if (!bwd && fused && t.geo >= 21 && t.geo <= 23 && c.tma && dtype == BM2F_DTYPE_F32 && d.L == 3) {
        if (t.geo == 21) return launch_geo<float, 3, true, false, 1, 28, 2, 3>(p, ml, mw, grid, st);
        if (t.geo == 22) return launch_geo<float, 3, true, false, 1, 26, 2, 5>(p, ml, mw, grid, st);
        return launch_geo<float, 3, true, false, 1, 24, 2, 7>(p, ml, mw, grid, st);
    }
    // the same kernel on bf16 values (cfg 3: Swin-L, bf16 forward): 64-byte corner lines, fp32 records and accumulation
    if (!bwd && (t.geo == 0 || t.geo == 17) && c.tma && c.sw == 32 && c.cps == 1 && !c.wide && dtype == BM2F_DTYPE_BF16 && !fused && d.L == 3)
        return launch_geo<__nv_bfloat16, 3, false, false, 1, 28, 2, 3>(p, ml, mw, grid, st);
    if (!bwd && t.geo == 11 && c.tma && c.sw == 32 && c.cps == 1 && !c.wide && dtype == BM2F_DTYPE_F32 && !fused && d.L == 3)
        return launch_geo<float, 3, false, false, 1, 28, 1, 3>(p, ml, mw, grid, st);      // 32-byte lean records (A/B)
    // geometry-warp forward variants kept for A/B (cfg shape, L = 3): tuning.geo = 1 with two CTAs per SM
    if (!bwd && t.geo == 1 && c.tma && c.sw == 32 && c.cps == 2 && dtype == BM2F_DTYPE_F32 && d.L == 3 && !fused)
        return launch_geo<float, 3, false, false, 2>(p, ml, mw, grid, st);
#ifdef BM2F_SWEEP
    // predicated records with 20 / 24 / 28 consumer warps (5 / 6 / 7); lean records with other warp splits (8 .. 16)
    if (!bwd && t.geo >= 5 && t.geo <= 16 && c.tma && c.sw == 32 && c.cps == 1 && dtype == BM2F_DTYPE_F32 && d.L == 3 && !fused) {
        switch (t.geo) {
        case 5: return launch_geo<float, 3, false, false, 1, 20>(p, ml, mw, grid, st);
        case 6: return launch_geo<float, 3, false, false, 1, 24>(p, ml, mw, grid, st);
        case 7: return launch_geo<float, 3, false, false, 1, 28>(p, ml, mw, grid, st);
        case 8: return launch_geo<float, 3, false, false, 1, 16, 1>(p, ml, mw, grid, st);
        case 9: return launch_geo<float, 3, false, false, 1, 24, 1>(p, ml, mw, grid, st);
        case 10: return launch_geo<float, 3, false, false, 1, 28, 1>(p, ml, mw, grid, st);
        case 12: return launch_geo<float, 3, false, false, 1, 26, 1, 5>(p, ml, mw, grid, st);
        case 13: return launch_geo<float, 3, false, false, 1, 24, 1, 6>(p, ml, mw, grid, st);
        case 14: return launch_geo<float, 3, false, false, 1, 24, 1, 4>(p, ml, mw, grid, st);
        case 15: return launch_geo<float, 3, false, false, 1, 20, 1, 8>(p, ml, mw, grid, st);
        case 16: return launch_geo<float, 3, false, false, 1, 16, 1, 4>(p, ml, mw, grid, st);
        default: break;
        }
    }
#endif
    if (!bwd && t.geo == 3 && c.tma && c.sw == 32 && c.cps == 1 && dtype == BM2F_DTYPE_F32 && d.L == 3 && !fused)
        return launch_geo<float, 3, false, true>(p, ml, mw, grid, st);      // + 256-bit gathers (cfg shape only)
    if (!bwd && t.geo == 1 && c.tma && c.sw == 32 && c.cps == 1 && !c.wide && dtype == BM2F_DTYPE_F32) {
        switch (d.L) {
        case 1: return fused ? launch_geo<float, 1, true>(p, ml, mw, grid, st) : launch_geo<float, 1, false>(p, ml, mw, grid, st);
        case 2: return fused ? launch_geo<float, 2, true>(p, ml, mw, grid, st) : launch_geo<float, 2, false>(p, ml, mw, grid, st);
        case 3: return fused ? launch_geo<float, 3, true>(p, ml, mw, grid, st) : launch_geo<float, 3, false>(p, ml, mw, grid, st);
        case 4: return fused ? launch_geo<float, 4, true>(p, ml, mw, grid, st) : launch_geo<float, 4, false>(p, ml, mw, grid, st);
        }
    }

    bool found = false;
    cudaError_t e = cudaSuccess;
    if (fused) {
        found = true;
#define BM2F_FUSED(T, L_) (c.tma ? launch_fused<T, L_, true>(bwd, p, ml, mw, grid, st) : launch_fused<T, L_, false>(bwd, p, ml, mw, grid, st))
        if (dtype == BM2F_DTYPE_F32) {
            switch (d.L) {
            case 1: e = BM2F_FUSED(float, 1); break;
            case 2: e = BM2F_FUSED(float, 2); break;
            case 3: e = BM2F_FUSED(float, 3); break;
            case 4: e = BM2F_FUSED(float, 4); break;
            }
        } else {
            e = BM2F_FUSED(__nv_bfloat16, 3);
        }
#undef BM2F_FUSED
    } else if (dtype == BM2F_DTYPE_F32) {
        switch (d.L) {
        case 1: e = dispatch_default<float, 1>(c, bwd, p, ml, mw, grid, st, &found); break;
        case 2: e = dispatch_default<float, 2>(c, bwd, p, ml, mw, grid, st, &found); break;
        case 3: e = dispatch_f32_l3(c, bwd, p, ml, mw, grid, st, &found); break;
        case 4: e = dispatch_default<float, 4>(c, bwd, p, ml, mw, grid, st, &found); break;
        }
    } else {
        e = dispatch_default<__nv_bfloat16, 3>(c, bwd, p, ml, mw, grid, st, &found);
    }
    if (!found)
        return fail(BM2F_ERR_UNSUPPORTED, "no fast kernel variant vec=%d strip_w=%d staging=%s ctas_per_sm=%d", c.vec,
                    c.sw, c.tma ? "tma" : "direct", c.cps);
    if (e != cudaSuccess) return cuda_fail(e, bwd ? "launch msda_bwd_fast_kernel" : "launch msda_fwd_fast_kernel");
    count_launch(1);
    return BM2F_OK;
}

int run_generic(bool bwd, const GenericParams &p, int dtype, cudaStream_t st)
{
    if (dtype == BM2F_DTYPE_BF16)
        return fail(BM2F_ERR_UNSUPPORTED,
                    "bf16 is implemented for D=32, L=3, P=4 only (got D=%d L=%d P=%d)", p.D, p.L, p.P);
    const long long warps = static_cast<long long>(p.N) * p.Lq * p.M;
    const long long blocks = (warps + 7) / 8;
    if (blocks >= (1ll << 31)) return fail(BM2F_ERR_UNSUPPORTED, "problem too large for the generic kernel");
    const int grid = static_cast<int>(blocks);
    if (dtype == BM2F_DTYPE_F32) {
        if (bwd) msda_bwd_generic_kernel<float><<<grid, 256, 0, st>>>(p);
        else msda_fwd_generic_kernel<float><<<grid, 256, 0, st>>>(p);
    } else {
        if (bwd) msda_bwd_generic_kernel<double><<<grid, 256, 0, st>>>(p);
        else msda_fwd_generic_kernel<double><<<grid, 256, 0, st>>>(p);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, bwd ? "launch msda_bwd_generic_kernel" : "launch msda_fwd_generic_kernel");
    count_launch(1);
    return BM2F_OK;
}

}  // namespace

extern "C" {

int bm2f_msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                      const void *sampling_loc, const void *attn_weight, void *output, int batch, int spatial_size,
                      int num_heads, int channels, int num_levels, int num_query, int num_point, int dtype,
                      const bm2f_msda_tuning_t *tuning, void *stream)
{
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    int rc = check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, d, dtype);
    if (rc) return rc;
    if (!output) return fail(BM2F_ERR_INVALID, "null output pointer");
    if (num_levels > kMaxLevels) return fail(BM2F_ERR_UNSUPPORTED, "num_levels %d > %d", num_levels, kMaxLevels);
    const bm2f_msda_tuning_t t = resolve_tuning(tuning);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    if (fast_eligible(d, dtype, t, value, output, sampling_loc, attn_weight)) {
        FastParams p{};
        p.value = value; p.shapes = spatial_shapes; p.start = level_start_index;
        p.loc = static_cast<const float *>(sampling_loc); p.attn = static_cast<const float *>(attn_weight);
        p.out = output;
        p.N = d.N; p.S = d.S; p.M = d.M; p.Lq = d.Lq;
        return run_fast(false, p, d, dtype, t, st);
    }
    GenericParams g{};
    g.value = value; g.shapes = spatial_shapes; g.start = level_start_index; g.loc = sampling_loc;
    g.attn = attn_weight; g.out = output;
    g.N = d.N; g.S = d.S; g.M = d.M; g.D = d.D; g.L = d.L; g.Lq = d.Lq; g.P = d.P;
    return run_generic(false, g, dtype, st);
}

int bm2f_msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_loc, const void *attn_weight, const void *grad_output, void *grad_value,
                       void *grad_sampling_loc, void *grad_attn_weight, int batch, int spatial_size, int num_heads,
                       int channels, int num_levels, int num_query, int num_point, int dtype,
                       const bm2f_msda_tuning_t *tuning, void *stream)
{
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    int rc = check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, d, dtype);
    if (rc) return rc;
    if (!grad_output || !grad_value || !grad_sampling_loc || !grad_attn_weight)
        return fail(BM2F_ERR_INVALID, "null gradient pointer");
    if (num_levels > kMaxLevels) return fail(BM2F_ERR_UNSUPPORTED, "num_levels %d > %d", num_levels, kMaxLevels);
    const bm2f_msda_tuning_t t = resolve_tuning(tuning);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    // grad_value accumulates through atomics: zero it first (reference: at::zeros_like, .cu:126)
    // bf16 values still accumulate grad_value in float32 (see include/bm2f_msda.h)
    const size_t gv_bytes = static_cast<size_t>(d.N) * d.S * d.M * d.D * (dtype == BM2F_DTYPE_F64 ? 8 : 4);
    cudaError_t e = cudaMemsetAsync(grad_value, 0, gv_bytes, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_value)");

    if (fast_eligible(d, dtype, t, value, grad_value, sampling_loc, attn_weight) && aligned16(grad_output) &&
        aligned16(grad_sampling_loc) && aligned16(grad_attn_weight)) {
        FastParams p{};
        p.value = value; p.shapes = spatial_shapes; p.start = level_start_index;
        p.loc = static_cast<const float *>(sampling_loc); p.attn = static_cast<const float *>(attn_weight);
        p.grad_out = grad_output; p.grad_value = grad_value;
        p.grad_loc = static_cast<float *>(grad_sampling_loc); p.grad_attn = static_cast<float *>(grad_attn_weight);
        p.N = d.N; p.S = d.S; p.M = d.M; p.Lq = d.Lq;
        bool sorted = false;
        if ((rc = choose_bwd_sorted(d, dtype, t, &sorted))) return rc;
        if (sorted) return run_bwd_sorted(p, d, t, false, st);
        return run_fast(true, p, d, dtype, t, st);
    }
    if (t.bwd == 2) return fail(BM2F_ERR_UNSUPPORTED, "anchor-sorted backward: unsupported shape or alignment");
    GenericParams g{};
    g.value = value; g.shapes = spatial_shapes; g.start = level_start_index; g.loc = sampling_loc;
    g.attn = attn_weight; g.grad_out = grad_output; g.grad_value = grad_value; g.grad_loc = grad_sampling_loc;
    g.grad_attn = grad_attn_weight;
    g.N = d.N; g.S = d.S; g.M = d.M; g.D = d.D; g.L = d.L; g.Lq = d.Lq; g.P = d.P;
    return run_generic(true, g, dtype, st);
}


int bm2f_msda_fused_supported(int num_heads, int channels, int num_levels, int num_point, int dtype)
{
    if (channels != 32 || num_point != 4 || num_heads != kHeads) return 0;
    if (dtype == BM2F_DTYPE_F32) return num_levels >= 1 && num_levels <= 4;
    if (dtype == BM2F_DTYPE_BF16) return num_levels == 3;
    return 0;
}

static int fused_forward_impl(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                              const void *reference_points, const void *sampling_offsets, const void *attn_logits,
                              void *output, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                              int num_query, int num_point, int dtype, const bm2f_msda_tuning_t *tuning, void *stream,
                              int ld_packed)
{
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    int rc = check_common(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, d, dtype);
    if (rc) return rc;
    if (!output) return fail(BM2F_ERR_INVALID, "null output pointer");
    if (!reference_points && num_query != spatial_size)
        return fail(BM2F_ERR_INVALID, "reference_points == NULL (pixel-centre reference points) needs num_query == spatial_size");
    const bm2f_msda_tuning_t t = resolve_tuning(tuning);
    if (!bm2f_msda_fused_supported(num_heads, channels, num_levels, num_point, dtype) ||
        !fast_eligible(d, dtype, t, value, output, sampling_offsets, attn_logits) ||
        (reinterpret_cast<uintptr_t>(reference_points) & 7u))
        return fail(BM2F_ERR_UNSUPPORTED,
                    "fused path covers D=32, M=8, P=4, L<=4 (f32) / L=3 (bf16), 16-byte aligned tensors "
                    "(got M=%d D=%d L=%d P=%d dtype=%d); use bm2f_msda_forward",
                    num_heads, channels, num_levels, num_point, dtype);
    FastParams p{};
    p.value = value; p.shapes = spatial_shapes; p.start = level_start_index;
    p.loc = static_cast<const float *>(sampling_offsets); p.attn = static_cast<const float *>(attn_logits);
    p.ref = static_cast<const float *>(reference_points);
    p.out = output;
    p.N = d.N; p.S = d.S; p.M = d.M; p.Lq = d.Lq;
    p.ld_packed = ld_packed;
    if (ld_packed && t.staging != 0 && t.staging != 1)
        return fail(BM2F_ERR_UNSUPPORTED, "packed offsets||logits are read through TMA tensor maps only (tuning.staging = %d)", t.staging);
    return run_fast(false, p, d, dtype, t, static_cast<cudaStream_t>(stream), true);
}

int bm2f_msda_fused_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const void *reference_points, const void *sampling_offsets, const void *attn_logits,
                            void *output, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                            int num_query, int num_point, int dtype, const bm2f_msda_tuning_t *tuning, void *stream)
{
    return fused_forward_impl(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits, output,
                              batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, dtype, tuning,
                              stream, 0);
}

int bm2f_msda_fused_forward_packed(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                   const void *reference_points, const void *offsets_logits, void *output, int batch,
                                   int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                                   int num_point, int dtype, const bm2f_msda_tuning_t *tuning, void *stream)
{
    // row r of offsets_logits: [M * L * P * 2 offsets | M * L * P logits], as written by ONE projection GEMM
    if (!offsets_logits) return fail(BM2F_ERR_INVALID, "null tensor pointer");
    const int lp = num_levels * num_point;
    const float *oa = static_cast<const float *>(offsets_logits);
    return fused_forward_impl(value, spatial_shapes, level_start_index, reference_points, oa, oa + static_cast<size_t>(num_heads) * lp * 2,
                              output, batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, dtype, tuning,
                              stream, num_heads * lp * 3);
}

static int fused_backward_impl(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                               const void *reference_points, const void *sampling_offsets, const void *attn_logits,
                               const void *grad_output, void *grad_value, void *grad_sampling_offsets,
                               void *grad_attn_logits, int batch, int spatial_size, int num_heads, int channels,
                               int num_levels, int num_query, int num_point, int dtype,
                               const bm2f_msda_tuning_t *tuning, void *stream, int ld_packed)
{
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    int rc = check_common(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, d, dtype);
    if (rc) return rc;
    if (!grad_output || !grad_value || !grad_sampling_offsets || !grad_attn_logits)
        return fail(BM2F_ERR_INVALID, "null gradient pointer");
    if (!reference_points && num_query != spatial_size)
        return fail(BM2F_ERR_INVALID, "reference_points == NULL (pixel-centre reference points) needs num_query == spatial_size");
    const bm2f_msda_tuning_t t = resolve_tuning(tuning);
    if (!bm2f_msda_fused_supported(num_heads, channels, num_levels, num_point, dtype) ||
        !fast_eligible(d, dtype, t, value, grad_value, sampling_offsets, attn_logits) || !aligned16(grad_output) ||
        !aligned16(grad_sampling_offsets) || !aligned16(grad_attn_logits) ||
        (reinterpret_cast<uintptr_t>(reference_points) & 7u))
        return fail(BM2F_ERR_UNSUPPORTED, "fused path: unsupported shape or alignment; use bm2f_msda_backward");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t gv_bytes = static_cast<size_t>(d.N) * d.S * d.M * d.D * 4;
    cudaError_t e = cudaMemsetAsync(grad_value, 0, gv_bytes, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_value)");
    FastParams p{};
    p.value = value; p.shapes = spatial_shapes; p.start = level_start_index;
    p.loc = static_cast<const float *>(sampling_offsets); p.attn = static_cast<const float *>(attn_logits);
    p.ref = static_cast<const float *>(reference_points);
    p.grad_out = grad_output; p.grad_value = grad_value;
    p.grad_loc = static_cast<float *>(grad_sampling_offsets); p.grad_attn = static_cast<float *>(grad_attn_logits);
    p.N = d.N; p.S = d.S; p.M = d.M; p.Lq = d.Lq;
    p.ld_packed = ld_packed;
    if (ld_packed && t.staging != 0 && t.staging != 1)
        return fail(BM2F_ERR_UNSUPPORTED, "packed offsets||logits are read through TMA tensor maps only (tuning.staging = %d)", t.staging);
    bool sorted = false;
    if ((rc = choose_bwd_sorted(d, dtype, t, &sorted))) return rc;
    if (sorted) return run_bwd_sorted(p, d, t, true, st);
    return run_fast(true, p, d, dtype, t, st, true);
}

int bm2f_msda_fused_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                             const void *reference_points, const void *sampling_offsets, const void *attn_logits,
                             const void *grad_output, void *grad_value, void *grad_sampling_offsets,
                             void *grad_attn_logits, int batch, int spatial_size, int num_heads, int channels,
                             int num_levels, int num_query, int num_point, int dtype,
                             const bm2f_msda_tuning_t *tuning, void *stream)
{
    return fused_backward_impl(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits,
                               grad_output, grad_value, grad_sampling_offsets, grad_attn_logits, batch, spatial_size, num_heads,
                               channels, num_levels, num_query, num_point, dtype, tuning, stream, 0);
}

int bm2f_msda_fused_backward_packed(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                    const void *reference_points, const void *offsets_logits, const void *grad_output,
                                    void *grad_value, void *grad_offsets_logits, int batch, int spatial_size, int num_heads,
                                    int channels, int num_levels, int num_query, int num_point, int dtype,
                                    const bm2f_msda_tuning_t *tuning, void *stream)
{
    if (!offsets_logits || !grad_offsets_logits) return fail(BM2F_ERR_INVALID, "null tensor pointer");
    const size_t split_at = static_cast<size_t>(num_heads) * num_levels * num_point * 2;
    const float *oa = static_cast<const float *>(offsets_logits);
    float *goa = static_cast<float *>(grad_offsets_logits);
    return fused_backward_impl(value, spatial_shapes, level_start_index, reference_points, oa, oa + split_at, grad_output,
                               grad_value, goa, goa + split_at, batch, spatial_size, num_heads, channels, num_levels, num_query,
                               num_point, dtype, tuning, stream, num_heads * num_levels * num_point * 3);
}

}  // extern "C"

// C ABI of libbm2f_msda.so (see include/bm2f_msda.h): argument checks, variant selection,
// TMA tensor-map encoding and kernel launches.  No torch types here; the torch-facing shim
// (msda_torch_ext.cpp) and ctypes callers sit on top of this file.
//
// Replaces the reference host wrappers ms_deform_attn_cuda_forward/backward and the launchers
// ms_deformable_im2col_cuda / ms_deformable_col2im_cuda
// (/root/reference/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_attn_cuda.cu:25-158,
//  ms_deform_im2col_cuda.cuh:928-1332).
#include "../../include/bm2f_msda.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "msda_fast.cuh"
#include "msda_generic.cuh"
#include "linear_tf32x3.cuh"
#include "ln_kernels.cuh"
#include "glue_kernels.cuh"

namespace {

using namespace bm2f;

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
bm2f_msda_tuning_t g_default_tuning = {};
std::mutex g_tuning_mu;

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    return fail(BM2F_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

// ---- device facts, cached per device (immutable once written) ----------------------------
struct DevInfo {
    std::atomic<int> ready{0};
    int sms = 0;
    int cc_major = 0;
};
DevInfo g_dev[64];

int device_info(int *sms, int *cc_major)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) return fail(BM2F_ERR_UNSUPPORTED, "device ordinal %d out of range", dev);
    DevInfo &d = g_dev[dev];
    if (!d.ready.load(std::memory_order_acquire)) {
        int s = 0, maj = 0;
        if ((e = cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
            return cuda_fail(e, "cudaDeviceGetAttribute(SM count)");
        if ((e = cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess)
            return cuda_fail(e, "cudaDeviceGetAttribute(cc major)");
        d.sms = s;
        d.cc_major = maj;
        d.ready.store(1, std::memory_order_release);
    }
    *sms = d.sms;
    *cc_major = d.cc_major;
    return BM2F_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) function attribute: remember it per
// (kernel, device ordinal).  A race only repeats the same idempotent call.
template <auto Kernel>
int ensure_dynamic_smem(int bytes, const char *what)
{
    static std::atomic<uint64_t> done{0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) return fail(BM2F_ERR_UNSUPPORTED, "device ordinal %d out of range", dev);
    if ((done.load(std::memory_order_acquire) >> dev) & 1ull) return BM2F_OK;
    e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return cuda_fail(e, what);
    done.fetch_or(1ull << dev, std::memory_order_release);
    return BM2F_OK;
}

// ---- TMA tensor maps ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 2-D fp32 matrix (rows x cols, row-major) -> boxes of (box_rows x box_cols).
int make_map(CUtensorMap *map, const float *base, uint64_t rows, uint64_t cols, uint32_t box_rows,
             uint32_t box_cols, bool swizzle128 = false)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(BM2F_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {cols * sizeof(float)};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(BM2F_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
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
template <typename T, int L_, bool FUSED, bool WIDE = false, int CPS = 1>
int launch_geo(const FastParams &p, const CUtensorMap &ml, const CUtensorMap &mw, int grid, cudaStream_t st)
{
    constexpr int SW = 32;
    constexpr int threads = (kNWarp + 1 + kGeoWarps) * 32;
    constexpr int smem = GeoRing<L_, 4, SW, kGeoStages>::kBytes;
    int rc = ensure_dynamic_smem<&msda_fwd_geo_kernel<T, L_, 4, SW, kNWarp, kGeoWarps, kGeoStages, FUSED, WIDE, CPS>>(
        smem, "cudaFuncSetAttribute(geometry-warp forward smem)");
    if (rc) return rc;
    msda_fwd_geo_kernel<T, L_, 4, SW, kNWarp, kGeoWarps, kGeoStages, FUSED, WIDE, CPS><<<grid, threads, smem, st>>>(p, ml, mw);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch msda_fwd_geo_kernel");
    g_launches.fetch_add(1, std::memory_order_relaxed);
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

// Full sweep grid for the Mask2Former shape (L = 3, P = 4, fp32).
cudaError_t dispatch_f32_l3(const FastChoice &c, bool bwd, const FastParams &p, const CUtensorMap &ml,
                            const CUtensorMap &mw, int grid, cudaStream_t st, bool *found)
{
#define BM2F_ROW(VEC, SW)        \
    BM2F_CASE(float, VEC, 3, SW, 0, 1) \
    BM2F_CASE(float, VEC, 3, SW, 0, 2) \
    BM2F_CASE(float, VEC, 3, SW, 1, 1) \
    BM2F_CASE(float, VEC, 3, SW, 1, 2)
    BM2F_ROW(4, 8) BM2F_ROW(4, 16) BM2F_ROW(4, 32)
    BM2F_ROW(2, 8) BM2F_ROW(2, 16) BM2F_ROW(2, 32)
    BM2F_ROW(1, 8) BM2F_ROW(1, 16) BM2F_ROW(1, 32)
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

bm2f_msda_tuning_t resolve_tuning(const bm2f_msda_tuning_t *t)
{
    bm2f_msda_tuning_t r;
    if (t) {
        r = *t;
    } else {
        std::lock_guard<std::mutex> lk(g_tuning_mu);
        r = g_default_tuning;
    }
    return r;
}

struct Dims {
    int N, S, M, D, L, Lq, P;
};

int check_common(const void *value, const int64_t *shapes, const int64_t *start, const void *loc, const void *attn,
                 const Dims &d, int dtype)
{
    if (!value || !shapes || !start || !loc || !attn) return fail(BM2F_ERR_INVALID, "null tensor pointer");
    if (d.N <= 0 || d.S <= 0 || d.M <= 0 || d.D <= 0 || d.L <= 0 || d.Lq <= 0 || d.P <= 0)
        return fail(BM2F_ERR_INVALID, "non-positive dimension (N=%d S=%d M=%d D=%d L=%d Lq=%d P=%d)", d.N, d.S, d.M,
                    d.D, d.L, d.Lq, d.P);
    if (dtype != BM2F_DTYPE_F32 && dtype != BM2F_DTYPE_F64 && dtype != BM2F_DTYPE_BF16)
        return fail(BM2F_ERR_INVALID, "unknown dtype %d", dtype);
    return BM2F_OK;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
        if ((rc = make_map(&ml, p.loc, rows_total, static_cast<uint64_t>(d.M) * LP * 2, c.sw, LP * 2))) return rc;
        if ((rc = make_map(&mw, p.attn, rows_total, static_cast<uint64_t>(d.M) * LP, c.sw, LP))) return rc;
    }

    // geometry-warp forward (float32, TMA staging, strip 32): opt-in through tuning.geo = 1
    if (!bwd && t.geo == 1 && c.tma && c.sw == 32 && c.cps == 2 && dtype == BM2F_DTYPE_F32 && d.L == 3 && !fused)
        return launch_geo<float, 3, false, false, 2>(p, ml, mw, grid, st);  // two CTAs per SM (cfg shape only)
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return BM2F_OK;
}

size_t elem_size(int dtype) { return dtype == BM2F_DTYPE_F64 ? 8 : (dtype == BM2F_DTYPE_BF16 ? 2 : 4); }
size_t loc_elem_size(int dtype) { return dtype == BM2F_DTYPE_F64 ? 8 : 4; }

}  // namespace

extern "C" {

int bm2f_msda_abi_version(void) { return BM2F_MSDA_ABI_VERSION; }

const char *bm2f_msda_build_info(void)
{
    return "bm2f_msda sm_100a fast(D=32,P=4,L<=4: f32; L=3: bf16) + generic(f32,f64); built " __DATE__ " " __TIME__;
}

const char *bm2f_msda_last_error(void) { return g_err; }

uint64_t bm2f_msda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void bm2f_msda_set_default_tuning(const bm2f_msda_tuning_t *tuning)
{
    std::lock_guard<std::mutex> lk(g_tuning_mu);
    if (tuning) g_default_tuning = *tuning;
    else memset(&g_default_tuning, 0, sizeof(g_default_tuning));
}

int bm2f_msda_check_im2col_step(int batch, int im2col_step)
{
    if (batch <= 0 || im2col_step <= 0)
        return fail(BM2F_ERR_IM2COL_STEP, "batch(%d) and im2col_step(%d) must be positive", batch, im2col_step);
    const int step = batch < im2col_step ? batch : im2col_step;
    if (batch % step != 0)
        return fail(BM2F_ERR_IM2COL_STEP, "batch(%d) must divide im2col_step(%d)", batch, step);
    return BM2F_OK;
}

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
        return run_fast(true, p, d, dtype, t, st);
    }
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

int bm2f_msda_fused_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const void *reference_points, const void *sampling_offsets, const void *attn_logits,
                            void *output, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                            int num_query, int num_point, int dtype, const bm2f_msda_tuning_t *tuning, void *stream)
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
    return run_fast(false, p, d, dtype, t, static_cast<cudaStream_t>(stream), true);
}

int bm2f_msda_fused_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                             const void *reference_points, const void *sampling_offsets, const void *attn_logits,
                             const void *grad_output, void *grad_value, void *grad_sampling_offsets,
                             void *grad_attn_logits, int batch, int spatial_size, int num_heads, int channels,
                             int num_levels, int num_query, int num_point, int dtype,
                             const bm2f_msda_tuning_t *tuning, void *stream)
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
    return run_fast(true, p, d, dtype, t, st, true);
}

// ---------------------------------------------------------------------------------------------
// tcgen05 projection GEMM (linear_tf32x3.cuh)
// ---------------------------------------------------------------------------------------------
}  // extern "C"
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return BM2F_OK;
}
}  // namespace
extern "C" {

size_t bm2f_linear_workspace_bytes(int out_features, int in_features)
{
    return static_cast<size_t>(2) * out_features * in_features * sizeof(float);
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
    // split 3 / 1: persistent kernel (double-buffered TMEM accumulator); +10: one tile per CTA (kept for A/B);
    // +20: persistent kernel with coalesced-store epilogue instead of the TMA store (A/B);
    // +30: persistent kernel in clusters of two CTAs with TMA-multicast weights (A/B)
    // +40 / +50 / +60: more activation bytes in flight (8 producer warps x 5 k-blocks / 4 x 4 / 8 x 4) (A/B)
    // +70: CTA pairs issuing tcgen05.mma.cta_group::2 (M = 256 per pair)
    // split 1 (single TF32 pass) loads the activation tile by TMA straight into the MMA stage by default (= 1 + 80);
    // 1 + 50 selects the register-staged activation path of the three-term kernel for comparison
    int xvar = 0;
    if (split >= 40) { xvar = split / 10 - 3; split -= (xvar + 3) * 10; }
    const bool cluster2 = split >= 30;
    if (cluster2) split -= 30;
    const bool stg_epilogue = split >= 20;
    if (stg_epilogue) split -= 20;
    const bool one_tile = split >= 10;
    if (one_tile) split -= 10;
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
        default: break;      // 288 = 2 x 144 columns does not fit two accumulators: one-tile kernel
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
    // split + 100 * c (A/B knob): cap the rows one CTA reduces at 256 * c, i.e. shorten the TMEM accumulation chain
    int row_cap = 0;
    if (split >= 100) { row_cap = (split / 100) * 256; split %= 100; }
    if (split != 1 && split != 3) return fail(BM2F_ERR_INVALID, "split must be 3 (tf32x3) or 1 (single TF32 pass)");
    int sms = 0, cc = 0;
    int rc = device_info(&sms, &cc);
    if (rc) return rc;
    if (cc < 10) return fail(BM2F_ERR_CUDA, "this library contains sm_100a code only; device has cc %d.x", cc);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(grad_weight, 0, static_cast<size_t>(out_features) * in_features * 4, st);
    if (e == cudaSuccess && grad_bias) e = cudaMemsetAsync(grad_bias, 0, static_cast<size_t>(out_features) * 4, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_weight / grad_bias)");
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(3, std::memory_order_relaxed);
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
    g_launches.fetch_add(3, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return BM2F_OK;
}

// ---------------------------------------------------------------------------------------------
// Host-buffer entry: chunked, double-buffered H2D -> kernels -> D2H.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kHostSlots = 3;   // chunks in flight: one uploading, one computing / downloading, one draining
struct HostPath {
    std::mutex mu;
    int dev = -1;
    cudaStream_t streams[kHostSlots] = {};
    void *ws[kHostSlots] = {};
    size_t ws_bytes = 0;
    int64_t *tabs = nullptr;  // shapes (2L) + start (L)
} g_host;

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
}  // namespace

int bm2f_msda_release_host_workspace(void)
{
    std::lock_guard<std::mutex> lk(g_host.mu);
    if (g_host.dev < 0) return BM2F_OK;
    int cur = 0;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if ((e = cudaSetDevice(g_host.dev)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    for (int i = 0; i < kHostSlots; ++i) {
        if (g_host.streams[i]) { cudaStreamSynchronize(g_host.streams[i]); cudaStreamDestroy(g_host.streams[i]); }
        if (g_host.ws[i]) cudaFree(g_host.ws[i]);
        g_host.streams[i] = nullptr;
        g_host.ws[i] = nullptr;
    }
    if (g_host.tabs) cudaFree(g_host.tabs);
    g_host.tabs = nullptr;
    g_host.ws_bytes = 0;
    g_host.dev = -1;
    cudaSetDevice(cur);
    return BM2F_OK;
}

int bm2f_msda_forward_backward_host(const void *value_host, const int64_t *spatial_shapes_host,
                                    const int64_t *level_start_index_host, const void *sampling_loc_host,
                                    const void *attn_weight_host, const void *grad_output_host, void *output_host,
                                    void *grad_value_host, void *grad_sampling_loc_host, void *grad_attn_weight_host,
                                    int batch, int spatial_size, int num_heads, int channels, int num_levels,
                                    int num_query, int num_point, int dtype, const bm2f_msda_tuning_t *tuning)
{
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    int rc = check_common(value_host, spatial_shapes_host, level_start_index_host, sampling_loc_host,
                          attn_weight_host, d, dtype);
    if (rc) return rc;
    if (num_levels > kMaxLevels) return fail(BM2F_ERR_UNSUPPORTED, "num_levels %d > %d", num_levels, kMaxLevels);
    const bool bwd = grad_output_host != nullptr;
    const size_t e = elem_size(dtype), el = loc_elem_size(dtype);
    const size_t v_img = static_cast<size_t>(d.S) * d.M * d.D * e;
    const size_t gv_img = static_cast<size_t>(d.S) * d.M * d.D * (dtype == BM2F_DTYPE_F64 ? 8 : 4);
    const size_t o_img = static_cast<size_t>(d.Lq) * d.M * d.D * e;
    const size_t l_img = static_cast<size_t>(d.Lq) * d.M * d.L * d.P * 2 * el;
    const size_t a_img = static_cast<size_t>(d.Lq) * d.M * d.L * d.P * el;

    // images per chunk: up to 16 chunks per call keep the pipeline's fill / drain (one chunk's upload at the start, one
    // chunk's download at the end are not overlapped) at ~1/16 of the call
    int chunk = d.N >= 16 ? d.N / 16 : 1;
    const size_t slot_bytes = static_cast<size_t>(chunk) *
                              (align256(v_img) + align256(l_img) + align256(a_img) + align256(o_img) +
                               (bwd ? align256(o_img) + align256(gv_img) + align256(l_img) + align256(a_img) : 0)) +
                              4096;

    std::lock_guard<std::mutex> lk(g_host.mu);
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaGetDevice");
    if (g_host.dev != dev || g_host.ws_bytes < slot_bytes) {
        for (int i = 0; i < kHostSlots; ++i) {
            if (g_host.ws[i]) cudaFree(g_host.ws[i]);
            g_host.ws[i] = nullptr;
            if (!g_host.streams[i] || g_host.dev != dev) {
                if ((ce = cudaStreamCreateWithFlags(&g_host.streams[i], cudaStreamNonBlocking)) != cudaSuccess)
                    return cuda_fail(ce, "cudaStreamCreate");
            }
            if ((ce = cudaMalloc(&g_host.ws[i], slot_bytes)) != cudaSuccess) {
                g_host.ws_bytes = 0;
                return cuda_fail(ce, "cudaMalloc(host-path workspace)");
            }
        }
        if (!g_host.tabs || g_host.dev != dev) {
            if ((ce = cudaMalloc(reinterpret_cast<void **>(&g_host.tabs), sizeof(int64_t) * 3 * kMaxLevels)) !=
                cudaSuccess)
                return cuda_fail(ce, "cudaMalloc(level tables)");
        }
        g_host.ws_bytes = slot_bytes;
        g_host.dev = dev;
    }
    cudaStream_t s0 = g_host.streams[0];
    if ((ce = cudaMemcpyAsync(g_host.tabs, spatial_shapes_host, sizeof(int64_t) * 2 * d.L, cudaMemcpyHostToDevice,
                              s0)) != cudaSuccess)
        return cuda_fail(ce, "H2D spatial_shapes");
    if ((ce = cudaMemcpyAsync(g_host.tabs + 2 * kMaxLevels, level_start_index_host, sizeof(int64_t) * d.L,
                              cudaMemcpyHostToDevice, s0)) != cudaSuccess)
        return cuda_fail(ce, "H2D level_start_index");
    if ((ce = cudaStreamSynchronize(s0)) != cudaSuccess) return cuda_fail(ce, "sync level tables");
    const int64_t *d_shapes = g_host.tabs, *d_start = g_host.tabs + 2 * kMaxLevels;

    auto hp = [](const void *base, size_t off) { return static_cast<const char *>(base) + off; };
    auto hpw = [](void *base, size_t off) { return static_cast<char *>(base) + off; };

    // Per chunk, on its slot's stream: ALL uploads first (grad_output included), then both kernels, then all downloads.
    // Copy engines serve requests in issue order, so an upload queued between a chunk's kernels and its downloads would
    // hold back the next chunk's uploads (head-of-line blocking) — with this order the H2D engine, the SMs and the D2H
    // engine each work on a different chunk.
    int slot = 0;
    for (int b0 = 0; b0 < d.N; b0 += chunk, slot = (slot + 1) % kHostSlots) {
        const int nb = (d.N - b0 < chunk) ? d.N - b0 : chunk;
        cudaStream_t st = g_host.streams[slot];
        char *w = static_cast<char *>(g_host.ws[slot]);
        auto take = [&](size_t per_img) { char *r = w; w += static_cast<size_t>(chunk) * align256(per_img); return r; };
        char *dv = take(v_img), *dl = take(l_img), *da = take(a_img), *dout = take(o_img);
        char *dgo = nullptr, *dgv = nullptr, *dgl = nullptr, *dga = nullptr;
        if (bwd) { dgo = take(o_img); dgv = take(gv_img); dgl = take(l_img); dga = take(a_img); }

#define BM2F_CP(dst, src, bytes, kind, what)                                                           \
    if ((ce = cudaMemcpyAsync(dst, src, bytes, kind, st)) != cudaSuccess) return cuda_fail(ce, what);
        BM2F_CP(dv, hp(value_host, b0 * v_img), nb * v_img, cudaMemcpyHostToDevice, "H2D value")
        BM2F_CP(dl, hp(sampling_loc_host, b0 * l_img), nb * l_img, cudaMemcpyHostToDevice, "H2D sampling_loc")
        BM2F_CP(da, hp(attn_weight_host, b0 * a_img), nb * a_img, cudaMemcpyHostToDevice, "H2D attn_weight")
        if (bwd) BM2F_CP(dgo, hp(grad_output_host, b0 * o_img), nb * o_img, cudaMemcpyHostToDevice, "H2D grad_output")
        rc = bm2f_msda_forward(dv, d_shapes, d_start, dl, da, dout, nb, d.S, d.M, d.D, d.L, d.Lq, d.P, dtype, tuning,
                               st);
        if (rc) return rc;
        if (bwd) {
            rc = bm2f_msda_backward(dv, d_shapes, d_start, dl, da, dgo, dgv, dgl, dga, nb, d.S, d.M, d.D, d.L, d.Lq,
                                    d.P, dtype, tuning, st);
            if (rc) return rc;
        }
        if (output_host) BM2F_CP(hpw(output_host, b0 * o_img), dout, nb * o_img, cudaMemcpyDeviceToHost, "D2H output")
        if (bwd) {
            if (grad_value_host)
                BM2F_CP(hpw(grad_value_host, b0 * gv_img), dgv, nb * gv_img, cudaMemcpyDeviceToHost, "D2H grad_value")
            if (grad_sampling_loc_host)
                BM2F_CP(hpw(grad_sampling_loc_host, b0 * l_img), dgl, nb * l_img, cudaMemcpyDeviceToHost,
                        "D2H grad_sampling_loc")
            if (grad_attn_weight_host)
                BM2F_CP(hpw(grad_attn_weight_host, b0 * a_img), dga, nb * a_img, cudaMemcpyDeviceToHost,
                        "D2H grad_attn_weight")
        }
#undef BM2F_CP
    }
    for (int i = 0; i < kHostSlots; ++i)
        if ((ce = cudaStreamSynchronize(g_host.streams[i])) != cudaSuccess) return cuda_fail(ce, "host-path sync");
    return BM2F_OK;
}

}  // extern "C"

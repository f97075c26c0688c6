// Shared host state of libbm2f_msda.so (see api_common.cuh) and the identification entry points of the C ABI.
#include "api_common.cuh"

#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace bm2f {
namespace host {
namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
bm2f_msda_tuning_t g_default_tuning = {};
std::mutex g_tuning_mu;

// ---- device facts, cached per device (immutable once written) ----------------------------
struct DevInfo {
    std::atomic<int> ready{0};
    int sms = 0;
    int cc_major = 0;
};
DevInfo g_dev[64];

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
}  // namespace

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

const char *last_error() { return g_err; }
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

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

// 2-D fp32 matrix (rows x cols, row-major) -> boxes of (box_rows x box_cols).
int make_map(CUtensorMap *map, const float *base, uint64_t rows, uint64_t cols, uint32_t box_rows,
             uint32_t box_cols, int swizzle, uint64_t ld_elems)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(BM2F_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {(ld_elems ? ld_elems : cols) * sizeof(float)};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                       : swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(BM2F_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return BM2F_OK;
}

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

void set_default_tuning(const bm2f_msda_tuning_t *tuning)
{
    std::lock_guard<std::mutex> lk(g_tuning_mu);
    if (tuning) g_default_tuning = *tuning;
    else memset(&g_default_tuning, 0, sizeof(g_default_tuning));
}

}  // namespace host
}  // namespace bm2f

using namespace bm2f::host;

extern "C" {

int bm2f_msda_abi_version(void) { return BM2F_MSDA_ABI_VERSION; }

const char *bm2f_msda_build_info(void)
{
    return "bm2f_msda sm_100a fast(D=32,P=4,L<=4: f32; L=3: bf16) + generic(f32,f64); built " __DATE__ " " __TIME__
#ifdef BM2F_SWEEP
           " [sweep variants]"
#endif
        ;
}

const char *bm2f_msda_last_error(void) { return last_error(); }

uint64_t bm2f_msda_launch_count(void) { return launch_count(); }

void bm2f_msda_set_default_tuning(const bm2f_msda_tuning_t *tuning) { set_default_tuning(tuning); }

int bm2f_msda_check_im2col_step(int batch, int im2col_step)
{
    if (batch <= 0 || im2col_step <= 0)
        return fail(BM2F_ERR_IM2COL_STEP, "batch(%d) and im2col_step(%d) must be positive", batch, im2col_step);
    const int step = batch < im2col_step ? batch : im2col_step;
    if (batch % step != 0)
        return fail(BM2F_ERR_IM2COL_STEP, "batch(%d) must divide im2col_step(%d)", batch, step);
    return BM2F_OK;
}

}  // extern "C"

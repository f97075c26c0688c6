// Host-side helpers shared by the translation units of libbm2f_msda.so: error reporting, the launch counter,
// per-device facts, TMA tensor-map encoding.  Nothing here is exported; the C ABI is include/bm2f_msda.h.
#pragma once
#include "../../include/bm2f_msda.h"

#include <atomic>
#include <cstdint>
#include <cstring>

#include <cuda.h>
#include <cuda_runtime.h>

namespace bm2f {
namespace host {

int fail(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
const char *last_error();
void count_launch(uint64_t n = 1);
uint64_t launch_count();

// SM count and compute-capability major of the current device (cached per ordinal)
int device_info(int *sms, int *cc_major);

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

// 2-D fp32 matrix (rows x cols, row-major) -> boxes of (box_rows x box_cols).
// swizzle: 0 none, 1 (true) SWIZZLE_128B, 2 SWIZZLE_128B_ATOM_32B (32-byte swizzle units: MN-major TF32 MMA operands)
// ld_elems: row stride in elements when the matrix is a column block of a wider one (0 = cols)
int make_map(CUtensorMap *map, const float *base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
             int swizzle = 0, uint64_t ld_elems = 0);

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct Dims {
    int N, S, M, D, L, Lq, P;
};
int check_common(const void *value, const int64_t *shapes, const int64_t *start, const void *loc, const void *attn,
                 const Dims &d, int dtype);
inline size_t elem_size(int dtype) { return dtype == BM2F_DTYPE_F64 ? 8 : (dtype == BM2F_DTYPE_BF16 ? 2 : 4); }
inline size_t loc_elem_size(int dtype) { return dtype == BM2F_DTYPE_F64 ? 8 : 4; }

bm2f_msda_tuning_t resolve_tuning(const bm2f_msda_tuning_t *t);

}  // namespace host
}  // namespace bm2f

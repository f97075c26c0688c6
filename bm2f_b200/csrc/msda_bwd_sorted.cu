// Launcher of the anchor-sorted backward (msda_bwd_sorted.cuh).  Called from bm2f_msda_backward /
// bm2f_msda_fused_backward (msda_api.cu) when the problem qualifies: float32, D = 32, M = 8, P = 4, L <= 4 and the
// queries are the pixels of the levels (Lq == S, encoder self-attention).
#include "api_common.cuh"
#include "msda_bwd_sorted.cuh"
#include "msda_bwd_pixel.cuh"
#include "msda_bwd_owner.cuh"

#include <atomic>

namespace bm2f {
namespace host {

namespace {
constexpr int kCells = 8192;      // 16-bit anchor cells per chunk (16 KB)

std::atomic<long long *> g_prof{nullptr};     // diagnostics: bm2f_msda_debug_phase_profile

template <int L_, int RMAX, int NWARP, int LPP, bool FUSED, int CPS, int PIPE>
int launch(const FastParams &p, int marg, const CUtensorMap &ml, const CUtensorMap &mw, const CUtensorMap &mg, int sms,
           long long est_jobs, cudaStream_t st)
{
    constexpr int smem = SortedSmem<L_, RMAX, kCells>::kBytes;
    int rc = ensure_dynamic_smem<&msda_bwd_sorted_kernel<L_, RMAX, NWARP, LPP, FUSED, kCells, CPS, PIPE>>(
        smem, "cudaFuncSetAttribute(sorted backward smem)");
    if (rc) return rc;
    const long long grid_max = static_cast<long long>(sms) * CPS;
    const int grid = static_cast<int>(est_jobs < grid_max ? est_jobs : grid_max);
    msda_bwd_sorted_kernel<L_, RMAX, NWARP, LPP, FUSED, kCells, CPS, PIPE>
        <<<grid, NWARP * 32, smem, st>>>(p, marg, g_prof.load(std::memory_order_relaxed), ml, mw, mg);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch msda_bwd_sorted_kernel");
    count_launch(1);
    return BM2F_OK;
}

template <int L_, int RMAX, int NWARP, bool FUSED, int CPS, bool F2>
int launch_pixel(const FastParams &p, int marg, const CUtensorMap &ml, const CUtensorMap &mw, const CUtensorMap &mg, int sms,
                 long long est_jobs, cudaStream_t st)
{
    constexpr int smem = PixelSmem<L_, RMAX, kCells>::kBytes;
    int rc = ensure_dynamic_smem<&msda_bwd_pixel_kernel<L_, RMAX, NWARP, FUSED, kCells, CPS, F2>>(
        smem, "cudaFuncSetAttribute(pixel-owner backward smem)");
    if (rc) return rc;
    const long long grid_max = static_cast<long long>(sms) * CPS;
    const int grid = static_cast<int>(est_jobs < grid_max ? est_jobs : grid_max);
    msda_bwd_pixel_kernel<L_, RMAX, NWARP, FUSED, kCells, CPS, F2>
        <<<grid, NWARP * 32, smem, st>>>(p, marg, g_prof.load(std::memory_order_relaxed), ml, mw, mg);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch msda_bwd_pixel_kernel");
    count_launch(1);
    return BM2F_OK;
}

template <int L_, int RMAX, int NWARP, bool FUSED, int CPS, int CELLS>
int launch_owner(const FastParams &p, int marg, const CUtensorMap &ml, const CUtensorMap &mw, const CUtensorMap &mg, int sms,
                 long long est_jobs, cudaStream_t st)
{
    constexpr int smem = OwnerSmem<L_, RMAX, CELLS, NWARP>::kBytes;
    int rc = ensure_dynamic_smem<&msda_bwd_owner_kernel<L_, RMAX, NWARP, FUSED, CELLS, CPS>>(
        smem, "cudaFuncSetAttribute(owner backward smem)");
    if (rc) return rc;
    const long long grid_max = static_cast<long long>(sms) * CPS;
    const int grid = static_cast<int>(est_jobs < grid_max ? est_jobs : grid_max);
    msda_bwd_owner_kernel<L_, RMAX, NWARP, FUSED, CELLS, CPS>
        <<<grid, NWARP * 32, smem, st>>>(p, marg, g_prof.load(std::memory_order_relaxed), ml, mw, mg);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "launch msda_bwd_owner_kernel");
    count_launch(1);
    return BM2F_OK;
}

template <int L_, int RMAX, bool FUSED>
int pick(const FastParams &p, int marg, int lanes, int variant, const CUtensorMap &ml, const CUtensorMap &mw,
         const CUtensorMap &mg, int sms, long long est_jobs, cudaStream_t st)
{
    // default: two CTAs of 8 warps per SM — while one sits in the short single-threaded / barrier phases the other computes
#define BM2F_V(NWARP, LPP, CPS, PIPE, R) return launch<L_, R, NWARP, LPP, FUSED, CPS, PIPE>(p, marg, ml, mw, mg, sms, est_jobs, st)
    if constexpr (L_ == 3 && !FUSED) {          // A/B variants (tuning.reserved[0]) for the benchmark shape only
        if (variant == 10) return launch_owner<L_, 5, 8, FUSED, 2, 4096>(p, marg, ml, mw, mg, sms, est_jobs, st);
        if (variant == 11) return launch_owner<L_, 4, 8, FUSED, 2, 4096>(p, marg, ml, mw, mg, sms, est_jobs, st);
        if (variant == 12) return launch_owner<L_, 3, 8, FUSED, 2, 4096>(p, marg, ml, mw, mg, sms, est_jobs, st);
        if (variant == 6) return launch_pixel<L_, 5, 8, FUSED, 2, false>(p, marg, ml, mw, mg, sms, est_jobs, st);
        if (variant == 7) return launch_pixel<L_, 8, 16, FUSED, 1, false>(p, marg, ml, mw, mg, sms, est_jobs, st);
        if (variant == 8) return launch_pixel<L_, 5, 8, FUSED, 2, true>(p, marg, ml, mw, mg, sms, est_jobs, st);
        if (variant == 9) return launch_pixel<L_, 8, 16, FUSED, 1, true>(p, marg, ml, mw, mg, sms, est_jobs, st);
        switch (variant * 10 + lanes) {
        case 18: BM2F_V(8, 8, 2, true, RMAX);
        case 14: BM2F_V(8, 4, 2, true, RMAX);
        case 28: BM2F_V(16, 8, 1, true, 8);
        case 24: BM2F_V(16, 4, 1, true, 8);
        case 38: BM2F_V(6, 8, 2, true, RMAX);
        case 34: BM2F_V(6, 4, 2, true, RMAX);
        case 48: BM2F_V(12, 8, 1, true, 9);
        case 44: BM2F_V(12, 4, 1, true, 9);
        case 58: BM2F_V(16, 8, 1, false, 8);
        case 54: BM2F_V(16, 4, 1, false, 8);
        case 134: BM2F_V(8, 4, 2, 2, RMAX);      // 13..16: phase 2 by cell ownership (PIPE = 2 / 3)
        case 138: BM2F_V(8, 8, 2, 2, RMAX);
        case 144: BM2F_V(8, 4, 2, 3, RMAX);
        case 154: BM2F_V(16, 4, 1, 2, 8);
        case 164: BM2F_V(16, 4, 1, 3, 8);
        case 174: BM2F_V(8, 4, 2, 0, 4);        // 17: 128-query chunks (78 KB per CTA, more of the SM's memory left to L1): 2.70 ms
        default: break;
        }
    }
    if (lanes == 4) BM2F_V(8, 4, 2, false, RMAX);
    BM2F_V(8, 8, 2, false, RMAX);
#undef BM2F_V
}
}  // namespace

bool bwd_sorted_eligible(const Dims &d, int dtype, const bm2f_msda_tuning_t &t)
{
    return dtype == BM2F_DTYPE_F32 && d.D == 32 && d.M == kHeads && d.P == 4 && d.L >= 1 && d.L <= 4 && d.Lq == d.S &&
           t.order == 0 && !t.force_generic;
}

int run_bwd_sorted(FastParams p, const Dims &d, const bm2f_msda_tuning_t &t, bool fused, cudaStream_t st)
{
    int sms = 0, cc = 0;
    int rc = device_info(&sms, &cc);
    if (rc) return rc;
    if (cc < 10) return fail(BM2F_ERR_CUDA, "this library contains sm_100a code only; device has cc %d.x", cc);
    const int marg = t.bwd_margin > 0 ? t.bwd_margin : 6;
    if (marg > 64) return fail(BM2F_ERR_INVALID, "bwd_margin %d out of range (1..64)", marg);
    const int lanes = t.bwd_lanes == 8 ? 8 : 4;       // 4 lanes x 8 channels per point measured faster (2.53 vs 2.92 ms, cfg 2)
    p.order = t.order;
    p.rows = 0;
    const int LP = d.L * d.P;
    const uint64_t rows_total = static_cast<uint64_t>(d.N) * d.Lq;
    CUtensorMap ml, mw, mg;
    if ((rc = make_map(&ml, p.loc, rows_total, static_cast<uint64_t>(d.M) * LP * 2, 32, LP * 2, 0, p.ld_packed))) return rc;
    if ((rc = make_map(&mw, p.attn, rows_total, static_cast<uint64_t>(d.M) * LP, 32, LP, 0, p.ld_packed))) return rc;
    if ((rc = make_map(&mg, static_cast<const float *>(p.grad_out), rows_total, static_cast<uint64_t>(d.M) * d.D, 32, d.D)))
        return rc;
    // lower bound on the chunk count (tiles hold at most 32 x RMAX queries); the kernel enumerates the exact tiles
    // from the device-resident shape table
    const long long est_jobs = static_cast<long long>(d.N) * d.M * ((d.Lq + 32 * 8 - 1) / (32 * 8));
    const int variant = t.reserved[0];
#define BM2F_SORTED(L_, RMAX)                                                                                         \
    return fused ? pick<L_, RMAX, true>(p, marg, lanes, variant, ml, mw, mg, sms, est_jobs, st)                       \
                 : pick<L_, RMAX, false>(p, marg, lanes, variant, ml, mw, mg, sms, est_jobs, st)
    switch (d.L) {
    case 1: BM2F_SORTED(1, 6);
    case 2: BM2F_SORTED(2, 6);
    case 3: BM2F_SORTED(3, 6);
    case 4: BM2F_SORTED(4, 4);
    }
#undef BM2F_SORTED
    return fail(BM2F_ERR_UNSUPPORTED, "sorted backward: L = %d", d.L);
}

void set_phase_profile(long long *dev) { g_prof.store(dev, std::memory_order_relaxed); }

}  // namespace host
}  // namespace bm2f

extern "C" void bm2f_msda_debug_phase_profile(void *device_buffer)
{
    bm2f::host::set_phase_profile(static_cast<long long *>(device_buffer));
}

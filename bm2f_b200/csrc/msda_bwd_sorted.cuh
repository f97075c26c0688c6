// Backward of multi-scale deformable attention with on-chip aggregation of grad_value (sm_100a, D = 32, P = 4, fp32).
//
// Why: the per-corner kernel (msda_bwd_fast_kernel) issues one 128-byte L2 reduction per bilinear corner — 48 per
// (query, head), 7.7x the op's algorithmic bytes — and sits on the L2 reduction rate (profiles/r01_*).  In encoder
// self-attention neighbouring queries sample neighbouring pixels, so most of those reductions hit the same lines.
// This kernel sorts the sampling points of a tile of queries by the pixel they land on and lets ONE lane group
// accumulate everything that lands on a pixel quad in registers before it touches L2:
//
//   chunk   = (image b, head m, tile of TW x TH queries of one level), TW <= 32, TH <= RMAX.  Its sampling
//             locations / attention weights (TMA boxes of 32 queries x L*P) and grad_output rows (32 x 32 floats)
//             are staged in shared memory.
//   phase 1 = one THREAD per sampling point: bilinear footprint computed once (the per-corner kernel repeats it in
//             the 8 channel lanes), record {attn, lh, lw, anchor pixel | corner mask | level} written to shared
//             memory; the point is counted into the cell of its anchor (top-left pixel of the footprint) inside a
//             per-level window around the tile (ATOMS.ADD returns its rank in the cell).
//   scan    = exclusive prefix sum over the window cells; phase 1c scatters the point ids -> points sorted by anchor.
//   phase 2 = the sorted list is cut into equal ranges, one per lane group (LPP lanes x 32/LPP channels).  A group
//             walks its range: when the anchor changes it flushes four pixel accumulators with vector REDs and loads
//             the four value lines of the new anchor; per point it reads the query's grad_output row from shared
//             memory ONCE (not once per corner), updates the four accumulators, forms the three dot products of
//             grad_attn / grad_loc, reduces them over the group's lanes and writes them back into the record.
//   phase 3 = one thread per query writes grad_sampling_loc / grad_attn_weight (fused mode: softmax backward).
//
// Points whose anchor falls outside the window (sampling offsets larger than the margin, or queries that are not the
// pixels of the levels) are handled on the spot by their warp, one corner line at a time — correct for any input,
// fast for the locality deformable attention has.
//
// Reference semantics: ms_deform_attn_col2im_bilinear and ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1
// (/root/reference/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:92-164, 306-408).
#pragma once

#include "msda_fast.cuh"

namespace bm2f {

struct SortedTabs {
    int H[4], W[4], start[4];
    int TW[4], TH[4], ntx[4], nty[4], job_base[4];
    int nql;           // query levels walked (1 in flat mode)
    int jobs_per_bm;
    int flat;          // 1: queries are not the pixels of the levels -> 1-D tiles, no windows
};

struct SortedWin {     // per chunk: anchor window of every sampled level
    int x0[4], y0[4], ww[4], wh[4], base[4];
    int ncells;
};

struct SortedJob {
    int b, m, k, x0, y0, ncols, nrows, qstart, Wq;   // tile of level k at (x0, y0); queries qstart + y * Wq + x
};

template <int L_, int RMAX, int CELLS_MAX>
struct SortedSmem {
    static constexpr int LP = L_ * 4;
    static constexpr int QMAX = 32 * RMAX;
    static constexpr int NPT = QMAX * LP;
    static constexpr int kLoc = 0;                                 // RMAX boxes of 32 x LP float2
    static constexpr int kAttn = kLoc + QMAX * LP * 8;             // RMAX boxes of 32 x LP float
    static constexpr int kGo = kAttn + QMAX * LP * 4;              // RMAX boxes of 32 x 32 float
    static constexpr int kRec = kGo + QMAX * 128;                  // NPT float4 records
    static constexpr int kPerm = kRec + NPT * 16;                  // NPT uint16 point ids, sorted by anchor
    static constexpr int kCnt = kPerm + ((NPT * 2 + 15) / 16) * 16;  // CELLS_MAX 16-bit counts -> offsets, two per word
    static constexpr int kBytes = kCnt + CELLS_MAX * 2;
};

template <int L_, int RMAX, int CELLS_MAX>
__device__ __forceinline__ int sorted_cells_upper(const SortedTabs &t, int k, int tw, int th, int marg)
{
    int cells = 0;
    for (int l = 0; l < L_; ++l) {
        int ww = (tw * t.W[l] + t.W[k] - 1) / t.W[k] + 2 * marg + 3;
        int wh = (th * t.H[l] + t.H[k] - 1) / t.H[k] + 2 * marg + 3;
        ww = min(ww, t.W[l] + 1);
        wh = min(wh, t.H[l] + 1);
        cells += ww * wh;
    }
    return cells;
}

template <int L_, int RMAX, int CELLS_MAX>
__device__ __forceinline__ void sorted_build_tabs(SortedTabs &t, const FastParams &p, int marg)
{
    int total = 0;
    for (int l = 0; l < L_; ++l) {
        t.H[l] = static_cast<int>(p.shapes[2 * l]);
        t.W[l] = static_cast<int>(p.shapes[2 * l + 1]);
        t.start[l] = static_cast<int>(p.start[l]);
        if (!level_in_bounds(p.start[l], p.shapes[2 * l], p.shapes[2 * l + 1], p.S)) { t.H[l] = 0; t.W[l] = 0; t.start[l] = 0; }
        total += t.H[l] * t.W[l];
    }
    int jobs = 0;
    if (p.order == 0 && total == p.Lq) {
        t.flat = 0;
        t.nql = L_;
        for (int k = 0; k < L_; ++k) {
            const int nx = (t.W[k] + 31) / 32;
            int tw = (t.W[k] + nx - 1) / nx, th = min(RMAX, t.H[k]);
            // the anchor windows of all sampled levels must fit the cell table: shrink the tile until they do
            while (sorted_cells_upper<L_, RMAX, CELLS_MAX>(t, k, tw, th, marg) > CELLS_MAX && (th > 1 || tw > 1)) {
                if (th > 1) th = (th + 1) / 2;
                else tw = (tw + 1) / 2;
            }
            t.TW[k] = tw;
            t.TH[k] = th;
            t.ntx[k] = (t.W[k] + tw - 1) / tw;
            t.nty[k] = (t.H[k] + th - 1) / th;
            t.job_base[k] = jobs;
            jobs += t.ntx[k] * t.nty[k];
        }
    } else {
        t.flat = 1;
        t.nql = 1;
        t.TW[0] = 32;
        t.TH[0] = RMAX;
        t.ntx[0] = 1;
        t.nty[0] = (p.Lq + 32 * RMAX - 1) / (32 * RMAX);
        t.job_base[0] = 0;
        jobs = t.nty[0];
    }
    t.jobs_per_bm = jobs;
}

__device__ __forceinline__ SortedJob sorted_decode_job(const SortedTabs &t, const FastParams &p, int j)
{
    SortedJob o;
    const int per_b = kHeads * t.jobs_per_bm;
    o.b = j / per_b;
    const int r = j - o.b * per_b;
    const int jj = r / kHeads;
    o.m = r - jj * kHeads;
    int k = 0;
    while (k + 1 < t.nql && jj >= t.job_base[k + 1]) ++k;
    o.k = k;
    const int tt = jj - t.job_base[k];
    const int ty = tt / t.ntx[k], tx = tt - ty * t.ntx[k];
    o.x0 = tx * t.TW[k];
    o.y0 = ty * t.TH[k];
    if (t.flat) {
        o.Wq = 32;
        o.qstart = 0;
        const int rows_total = (p.Lq + 31) / 32;
        o.nrows = min(t.TH[0], rows_total - o.y0);
        o.ncols = 32;          // the last row may be shorter: see sorted_row_cols
    } else {
        o.Wq = t.W[k];
        o.qstart = t.start[k];
        o.ncols = min(t.TW[k], t.W[k] - o.x0);
        o.nrows = min(t.TH[k], t.H[k] - o.y0);
    }
    return o;
}

// number of valid queries in tile row r
__device__ __forceinline__ int sorted_row_cols(const SortedTabs &t, const FastParams &p, const SortedJob &job, int r)
{
    if (!t.flat) return job.ncols;
    return min(32, p.Lq - (job.y0 + r) * 32);
}

template <int L_>
__device__ __forceinline__ void sorted_window(SortedWin &w, const SortedTabs &t, const SortedJob &job, int marg, int cells_max)
{
    int base = 0;
    for (int l = 0; l < L_; ++l) {
        int ww = 0, wh = 0, lox = 0, loy = 0;
        if (!t.flat) {
            const int Wk = t.W[job.k], Hk = t.H[job.k];
            lox = max((job.x0 * t.W[l]) / Wk - marg - 1, -1);
            loy = max((job.y0 * t.H[l]) / Hk - marg - 1, -1);
            const int hix = min(((job.x0 + job.ncols) * t.W[l] + Wk - 1) / Wk + marg, t.W[l] - 1);
            const int hiy = min(((job.y0 + job.nrows) * t.H[l] + Hk - 1) / Hk + marg, t.H[l] - 1);
            ww = max(hix - lox + 1, 0);
            wh = max(hiy - loy + 1, 0);
            if (ww > 0 && base + ww * wh > cells_max) wh = max((cells_max - base) / ww, 0);   // cannot happen with the tile search
        }
        w.x0[l] = lox; w.y0[l] = loy; w.ww[l] = ww; w.wh[l] = wh; w.base[l] = base;
        base += ww * wh;
    }
    w.ncells = base;
}

// Cell table: CELLS_MAX 16-bit counters packed two per 32-bit word (shared-memory atomics are 32-bit).  Cell c lives
// in half c / NW of word c % NW (NW = CELLS_MAX / 2), so neighbouring cells — which neighbouring queries hit in the
// same warp instruction — are different words in different banks.  Counts and offsets stay below 65536 (a chunk has at
// most QMAX * L * P points), so a carry never crosses the halves.
template <int NW>
__device__ __forceinline__ uint32_t cell_count_rank(uint32_t *cnt, int cell)
{
    const int half = cell >= NW;
    const uint32_t old = atomicAdd(&cnt[cell & (NW - 1)], half ? 0x10000u : 1u);
    return half ? old >> 16 : old & 0xffffu;
}
template <int NW>
__device__ __forceinline__ uint32_t cell_offset(const uint32_t *cnt, int cell)
{
    const uint32_t w = cnt[cell & (NW - 1)];
    return cell >= NW ? w >> 16 : w & 0xffffu;
}

// Exclusive prefix sum over cells [0, n) in place (cell order: low halves of words 0..NW-1, then high halves);
// returns the total.  One round per half: every thread owns NW / NT consecutive words.
template <int NT, int NW>
__device__ __forceinline__ int sorted_block_scan(uint32_t *cnt, int n, uint32_t *warp_sums, int tid)
{
    constexpr int NWARPS = NT / 32, PER = ((NW + NT - 1) / NT + 3) / 4 * 4;      // whole uint4s per thread
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t carry = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int nh = min(n - half * NW, NW);
        if (nh <= 0) break;                        // uniform
        const int sh = half * 16;
        const int idx = tid * PER;
        uint32_t w[PER];
#pragma unroll
        for (int i = 0; i < PER; i += 4) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (idx + i < nh) v = *reinterpret_cast<const uint4 *>(cnt + idx + i);
            w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) s += (w[i] >> sh) & 0xffffu;
        uint32_t incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < NWARPS; ++k) {
            const uint32_t ws = warp_sums[k];
            if (k < warp) woff += ws;
            tot += ws;
        }
        uint32_t run = carry + woff + incl - s;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const uint32_t c = (w[i] >> sh) & 0xffffu;
            w[i] = (w[i] & ~(0xffffu << sh)) | (run << sh);
            run += c;
        }
#pragma unroll
        for (int i = 0; i < PER; i += 4)
            if (idx + i < nh) *reinterpret_cast<uint4 *>(cnt + idx + i) = make_uint4(w[i], w[i + 1], w[i + 2], w[i + 3]);
        carry += tot;
        __syncthreads();
    }
    return static_cast<int>(carry);
}

// packed fp32 pairs: FFMA2 (fma.rn.f32x2) does two FMAs per issue slot on sm_100a
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

constexpr int kSortedSkip = -1;       // record already final ({ga, gx, gy, a}); not in the sorted list

// PIPE: 0 lane groups walk equal ranges of the sorted list, 1 same with the next anchor's lines prefetched,
// 2 lane group g owns the cells g, g + G, g + 2 G, ... (whole anchor runs; every group gets the same mix of coarse- and
// fine-level cells, so the groups finish together), 3 same with the next own cell's lines prefetched
template <int L_, int RMAX, int NWARP, int LPP, bool FUSED, int CELLS_MAX, int CPS, int PIPE>
__global__ void __launch_bounds__(NWARP * 32, CPS)
msda_bwd_sorted_kernel(const FastParams p, const int marg, long long *prof_out, const __grid_constant__ CUtensorMap tm_loc,
                       const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_go)
{
    constexpr int P_ = 4, LP = L_ * P_, NT = NWARP * 32, D = 32, MD = kHeads * D;
    using SM = SortedSmem<L_, RMAX, CELLS_MAX>;
    constexpr int QMAX = SM::QMAX;
    constexpr int PTS_PER_LEVEL = QMAX * P_;
    static_assert(PTS_PER_LEVEL % NT == 0, "points of a level must divide evenly over the CTA");
    constexpr int ITERS = PTS_PER_LEVEL / NT;
    constexpr int CH = D / LPP, NV = CH / 4, GPW = 32 / LPP, G = NWARP * GPW;
    static_assert(CH % 4 == 0, "whole float4s per lane");
    constexpr int NW = CELLS_MAX / 2;
    static_assert((NW & (NW - 1)) == 0, "cell words: a power of two");
    static_assert(SM::NPT < 65536, "16-bit counts / offsets / point ids");

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ SortedTabs tabs;
    __shared__ SortedWin win;
    __shared__ alignas(8) uint64_t bar_la, bar_go;
    __shared__ uint32_t warp_sums[NWARP];

    float2 *s_loc = reinterpret_cast<float2 *>(smem + SM::kLoc);
    float *s_attn = reinterpret_cast<float *>(smem + SM::kAttn);
    float *s_go = reinterpret_cast<float *>(smem + SM::kGo);
    float4 *s_rec = reinterpret_cast<float4 *>(smem + SM::kRec);
    uint16_t *s_perm = reinterpret_cast<uint16_t *>(smem + SM::kPerm);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem + SM::kCnt);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        sorted_build_tabs<L_, RMAX, CELLS_MAX>(tabs, p, marg);
        mbar_init(&bar_la, 1);
        mbar_init(&bar_go, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tm_loc);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_go);
    }
    __syncthreads();

    int H[L_], W[L_], st[L_];
    float Hf[L_], Wf[L_];
#pragma unroll
    for (int l = 0; l < L_; ++l) {
        H[l] = tabs.H[l]; W[l] = tabs.W[l]; st[l] = tabs.start[l];
        Hf[l] = static_cast<float>(H[l]); Wf[l] = static_cast<float>(W[l]);
    }
    const int total_jobs = p.N * kHeads * tabs.jobs_per_bm;

    auto issue_la = [&](const SortedJob &job) {      // thread 0
        mbar_arrive_expect_tx(&bar_la, static_cast<uint32_t>(job.nrows) * (32 * LP * 12));
        for (int r = 0; r < job.nrows; ++r) {
            const int row = job.b * p.Lq + job.qstart + (job.y0 + r) * job.Wq + job.x0;
            tma_load_2d(reinterpret_cast<unsigned char *>(s_loc) + r * (32 * LP * 8), &tm_loc, job.m * LP * 2, row, &bar_la);
            tma_load_2d(reinterpret_cast<unsigned char *>(s_attn) + r * (32 * LP * 4), &tm_w, job.m * LP, row, &bar_la);
        }
    };
    auto issue_go = [&](const SortedJob &job) {      // thread 0
        mbar_arrive_expect_tx(&bar_go, static_cast<uint32_t>(job.nrows) * (32 * D * 4));
        for (int r = 0; r < job.nrows; ++r) {
            const int row = job.b * p.Lq + job.qstart + (job.y0 + r) * job.Wq + job.x0;
            tma_load_2d(reinterpret_cast<unsigned char *>(s_go) + r * (32 * D * 4), &tm_go, job.m * D, row, &bar_go);
        }
    };

    if (tid == 0 && static_cast<int>(blockIdx.x) < total_jobs) {
        const SortedJob j0 = sorted_decode_job(tabs, p, blockIdx.x);
        issue_la(j0);
        issue_go(j0);
    }

    const float *value = static_cast<const float *>(p.value);
    const float *grad_out_unused = nullptr;
    (void)grad_out_unused;
    float *grad_value = static_cast<float *>(p.grad_value);

    // optional phase profile (tools/bwd_phases.py): cycles thread 0 spends in each phase, summed over this CTA's chunks
    const bool prof = prof_out != nullptr && tid == 0;
    long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = prof ? clock64() : 0;

    uint32_t phase = 0;
    for (int j = blockIdx.x; j < total_jobs; j += gridDim.x, phase ^= 1) {
        const SortedJob job = sorted_decode_job(tabs, p, j);
        const bool have_next = j + static_cast<int>(gridDim.x) < total_jobs;
        // ---- A: window of this chunk, clear the cell counters ----
        if (tid == 0) sorted_window<L_>(win, tabs, job, marg, CELLS_MAX);
        for (int i = tid * 4; i < NW; i += NT * 4) *reinterpret_cast<uint4 *>(s_cnt + i) = make_uint4(0, 0, 0, 0);
        __syncthreads();
        int wx0[L_], wy0[L_], ww[L_], wh[L_], wbase[L_];
#pragma unroll
        for (int l = 0; l < L_; ++l) {
            wx0[l] = win.x0[l]; wy0[l] = win.y0[l]; ww[l] = win.ww[l]; wh[l] = win.wh[l]; wbase[l] = win.base[l];
        }
        const int ncells = win.ncells;
        const size_t img = (static_cast<size_t>(job.b) * p.S * kHeads + job.m) * D;
        const float *vbm = value + img;
        float *gbm = grad_value + img;

        mbar_wait(&bar_la, phase);
        if (prof) { const long long c = clock64(); pt_[0] += c - t_prev; t_prev = c; }     // window + clear + wait loc/attn

        // ---- phase 0 (fused): softmax over each query's logits, loc = ref + offset / (W, H); lane = query ----
        if constexpr (FUSED) {
            static_assert(!FUSED || P_ == 4, "");
            for (int qi = tid; qi < QMAX; qi += NT) {
                const int r = qi >> 5, xi = qi & 31;
                if (r >= job.nrows || xi >= sorted_row_cols(tabs, p, job, r)) continue;
                float4 *sw = reinterpret_cast<float4 *>(s_attn + qi * LP);
                float4 w[L_];
#pragma unroll
                for (int l = 0; l < L_; ++l) w[l] = sw[l];
                float mx = fmaxf(fmaxf(w[0].x, w[0].y), fmaxf(w[0].z, w[0].w));
#pragma unroll
                for (int l = 1; l < L_; ++l) mx = fmaxf(mx, fmaxf(fmaxf(w[l].x, w[l].y), fmaxf(w[l].z, w[l].w)));
                float sum = 0.f;
#pragma unroll
                for (int l = 0; l < L_; ++l) {
                    w[l].x = __expf(w[l].x - mx); w[l].y = __expf(w[l].y - mx);
                    w[l].z = __expf(w[l].z - mx); w[l].w = __expf(w[l].w - mx);
                    sum += (w[l].x + w[l].y) + (w[l].z + w[l].w);
                }
                const float inv = __frcp_rn(sum);
#pragma unroll
                for (int l = 0; l < L_; ++l)
                    sw[l] = make_float4(w[l].x * inv, w[l].y * inv, w[l].z * inv, w[l].w * inv);
                const int q = job.qstart + (job.y0 + r) * job.Wq + job.x0 + xi;
                float4 *sl = reinterpret_cast<float4 *>(s_loc + qi * LP);
                const float2 *ref_q = p.ref ? reinterpret_cast<const float2 *>(p.ref) + (static_cast<size_t>(job.b) * p.Lq + q) * L_
                                            : nullptr;
                float2 centre = make_float2(0.f, 0.f);
                if (!p.ref) {
                    // encoder reference points, valid ratios 1 (msdeformattn.py:141-153): the query pixel's centre
                    float wq = Wf[0], hq = Hf[0];
#pragma unroll
                    for (int k = 1; k < L_; ++k)
                        if (job.k == k) { wq = Wf[k]; hq = Hf[k]; }
                    centre = make_float2((static_cast<float>(job.x0 + xi) + 0.5f) / wq,
                                         (static_cast<float>(job.y0 + r) + 0.5f) / hq);
                }
#pragma unroll
                for (int l = 0; l < L_; ++l) {
                    const float2 rr = ref_q ? __ldg(ref_q + l) : centre;
                    const float rw = 1.f / Wf[l], rh = 1.f / Hf[l];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 o = sl[2 * l + h];
                        o.x = fmaf(o.x, rw, rr.x); o.y = fmaf(o.y, rh, rr.y);
                        o.z = fmaf(o.z, rw, rr.x); o.w = fmaf(o.w, rh, rr.y);
                        sl[2 * l + h] = o;
                    }
                }
            }
            fence_proxy_async_smem();      // these buffers are later overwritten by TMA (async proxy)
            __syncthreads();
        }

        // ---- phase 1: one thread per sampling point ----
        uint32_t cr[L_ * ITERS];           // (cell << 16 | rank) of this thread's sorted points, 0xffffffff otherwise
        bool go_ready = false;
#pragma unroll
        for (int l = 0; l < L_; ++l) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const int idx = it * NT + tid;
                const int qi = idx >> 2, pp = idx & 3;
                const int pt = qi * LP + l * P_ + pp;
                const int r = qi >> 5, xi = qi & 31;
                const bool valid = r < job.nrows && xi < sorted_row_cols(tabs, p, job, r);
                const float2 xy = s_loc[pt];
                const float a = s_attn[pt];
                const Footprint f = make_footprint(xy.x, xy.y, H[l], W[l], Hf[l], Wf[l]);
                const int mask = (f.ok[0] ? 1 : 0) | (f.ok[1] ? 2 : 0) | (f.ok[2] ? 4 : 0) | (f.ok[3] ? 8 : 0);
                const bool live = valid && mask != 0;
                const int pix = st[l] + f.y0 * W[l] + f.x0;           // anchor pixel (may lie one row / column outside)
                const int cx = f.x0 - wx0[l], cy = f.y0 - wy0[l];
                const bool inside = live && static_cast<unsigned>(cx) < static_cast<unsigned>(ww[l]) &&
                                    static_cast<unsigned>(cy) < static_cast<unsigned>(wh[l]);
                uint32_t c = 0xffffffffu;
                float4 rec = make_float4(0.f, 0.f, 0.f, a);             // final form {ga, gx, gy, a}: skipped points
                if (inside) {
                    const int cell = wbase[l] + cy * ww[l] + cx;
                    const uint32_t rank = cell_count_rank<NW>(s_cnt, cell);
                    c = (static_cast<uint32_t>(cell) << 16) | rank;
                    rec = make_float4(a, f.lh, f.lw, __int_as_float((pix + W[l] + 1) | (mask << 24) | (l << 28)));
                }
                // points outside the window: this warp processes them now, FOUR at a time — lane group g (8 lanes x 4
                // channels, the per-corner kernel's shape: 128-bit gathers, vector REDs, 3-step shuffle reductions)
                // takes the g-th pending point; the group's first lane writes the point's final record
                unsigned fb = __ballot_sync(0xffffffffu, live && !inside);
                const bool mine_fallback = live && !inside;
                if (fb) {
                    if (!go_ready) { mbar_wait(&bar_go, phase); go_ready = true; }
                    const int lg4 = lane >> 3, sub4 = lane & 7;
                    while (fb) {
                        // the lg4-th set bit of fb (or none)
                        unsigned rest = fb;
                        int src = -1;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int bpos = rest ? __ffs(rest) - 1 : -1;
                            if (g == lg4) src = bpos;
                            rest &= rest - 1;
                        }
                        fb = rest;
                        const int s_ = src < 0 ? 0 : src;
                        const float a_ = __shfl_sync(0xffffffffu, a, s_);
                        const float lh = __shfl_sync(0xffffffffu, f.lh, s_), lw = __shfl_sync(0xffffffffu, f.lw, s_);
                        const int pix_ = __shfl_sync(0xffffffffu, pix, s_);
                        const int mask_all = __shfl_sync(0xffffffffu, mask, s_);
                        const int mask_ = src < 0 ? 0 : mask_all;
                        const int pt_ = __shfl_sync(0xffffffffu, pt, s_);
                        const float hh = 1.f - lh, hw = 1.f - lw;
                        const float4 g4 = *reinterpret_cast<const float4 *>(s_go + (pt_ / LP) * D + sub4 * 4);
                        const float cw[4] = {hh * hw, hh * lw, lh * hw, lh * lw};
                        const int poff[4] = {0, 1, W[l], W[l] + 1};
                        float t[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            t[k] = 0.f;
                            if ((mask_ >> k) & 1) {
                                const long long e = static_cast<long long>(pix_ + poff[k]) * MD + sub4 * 4;
                                const float4 v4 = __ldg(reinterpret_cast<const float4 *>(vbm + e));
                                t[k] = (g4.x * v4.x + g4.y * v4.y) + (g4.z * v4.z + g4.w * v4.w);
                                const float wk = a_ * cw[k];
                                const float rv[4] = {wk * g4.x, wk * g4.y, wk * g4.z, wk * g4.w};
                                VecIO<float, 4>::red_add(gbm + e, rv);
                            }
                        }
                        float pa = hh * (hw * t[0] + lw * t[1]) + lh * (hw * t[2] + lw * t[3]);
                        float px = hh * (t[1] - t[0]) + lh * (t[3] - t[2]);
                        float py = hw * (t[2] - t[0]) + lw * (t[3] - t[1]);
#pragma unroll
                        for (int o = 4; o > 0; o >>= 1) {
                            pa += __shfl_xor_sync(0xffffffffu, pa, o);
                            px += __shfl_xor_sync(0xffffffffu, px, o);
                            py += __shfl_xor_sync(0xffffffffu, py, o);
                        }
                        if (src >= 0 && sub4 == 0) s_rec[pt_] = make_float4(pa, a_ * px, a_ * py, a_);
                    }
                }
                if (!mine_fallback) s_rec[pt] = rec;
                cr[l * ITERS + it] = c;
            }
        }
        __syncthreads();
        if (prof) { const long long c = clock64(); pt_[1] += c - t_prev; t_prev = c; }     // (prologue +) phase 1

        // ---- scan + scatter: points sorted by anchor cell ----
        const int nsorted = sorted_block_scan<NT, NW>(s_cnt, ncells, warp_sums, tid);
        if (prof) { const long long c = clock64(); pt_[2] += c - t_prev; t_prev = c; }     // scan
#pragma unroll
        for (int l = 0; l < L_; ++l) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t c = cr[l * ITERS + it];
                if (c != 0xffffffffu) {
                    const int idx = it * NT + tid;
                    const int pt = (idx >> 2) * LP + l * P_ + (idx & 3);
                    s_perm[cell_offset<NW>(s_cnt, c >> 16) + (c & 0xffffu)] = static_cast<uint16_t>(pt);
                }
            }
        }
        __syncthreads();
        if (prof) { const long long c = clock64(); pt_[3] += c - t_prev; t_prev = c; }     // scatter
        // loc / attn buffers are free: prefetch the next chunk's
        SortedJob next{};
        if (tid == 0 && have_next) {
            next = sorted_decode_job(tabs, p, j + gridDim.x);
            issue_la(next);
        }

        // ---- phase 2: lane groups walk equal ranges of the sorted points ----
        if (!go_ready) mbar_wait(&bar_go, phase);
        if (prof) { const long long c = clock64(); pt_[4] += c - t_prev; t_prev = c; }
        {
            const int lg = lane / LPP, sub = lane % LPP;
            const unsigned gmask = (LPP == 32 ? 0xffffffffu : ((1u << LPP) - 1u)) << (lg * LPP);
            const int g = warp * GPW + lg;
            int i = static_cast<int>((static_cast<long long>(nsorted) * g) / G);
            const int iend = static_cast<int>((static_cast<long long>(nsorted) * (g + 1)) / G);
            // channel ownership: the lane's n-th float4 is channels n * 4 * LPP + sub * 4 .. + 3, so that ONE load / RED
            // instruction of the group covers LPP * 16 contiguous bytes = whole 32-byte sectors (with 8 consecutive channels
            // per lane every instruction touched all four sectors of the line and filled half of each)
            constexpr int CSTR = 4 * LPP;
            const float *vb = vbm + sub * 4;
            float *gb = gbm + sub * 4;
            int cur = kSortedSkip, cmask = 0, cW = 0;
            long long e00 = 0;
            float acc[4][CH], v[4][CH];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int c = 0; c < CH; ++c) { acc[k][c] = 0.f; v[k][c] = 0.f; }

            auto flush = [&]() {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((cmask >> k) & 1) {
                        float *dst = gb + e00 + ((k & 1) ? MD : 0) + ((k >> 1) ? static_cast<long long>(cW) * MD : 0);
#pragma unroll
                        for (int n = 0; n < NV; ++n) {
                            const float rv[4] = {acc[k][4 * n], acc[k][4 * n + 1], acc[k][4 * n + 2], acc[k][4 * n + 3]};
                            VecIO<float, 4>::red_add(dst + CSTR * n, rv);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CH; ++c) acc[k][c] = 0.f;
                }
            };
            // anchor of a record -> corner mask, level width, element offset of the anchor pixel
            auto decode = [&](int packed, int &mask, int &wl, long long &e) {
                mask = (packed >> 24) & 15;
                const int lvl = (packed >> 28) & 3;
                wl = W[0];
#pragma unroll
                for (int k = 1; k < L_; ++k)
                    if (lvl == k) wl = W[k];
                e = static_cast<long long>((packed & 0xffffff) - (wl + 1)) * MD;
            };
            auto load_lines = [&](float (&dst)[4][CH], int mask, int wl, long long e) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float *srcp = vb + e + ((k & 1) ? MD : 0) + ((k >> 1) ? static_cast<long long>(wl) * MD : 0);
#pragma unroll
                    for (int n = 0; n < NV; ++n) {
                        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if ((mask >> k) & 1) t4 = __ldg(reinterpret_cast<const float4 *>(srcp + CSTR * n));
                        dst[k][4 * n] = t4.x; dst[k][4 * n + 1] = t4.y; dst[k][4 * n + 2] = t4.z; dst[k][4 * n + 3] = t4.w;
                    }
                }
            };
            // one point: grad_output row of its query from shared memory, four accumulators, three dot products
            auto process = [&](int pt, const float4 rec) {
                const int qi = pt / LP;
                float go[CH];
#pragma unroll
                for (int n = 0; n < NV; ++n) {
                    const float4 t4 = *reinterpret_cast<const float4 *>(s_go + qi * D + sub * 4 + CSTR * n);
                    go[4 * n] = t4.x; go[4 * n + 1] = t4.y; go[4 * n + 2] = t4.z; go[4 * n + 3] = t4.w;
                }
                const float a = rec.x, lh = rec.y, lw = rec.z;
                const float hh = 1.f - lh, hw = 1.f - lw;
                float t[4];
#ifdef BM2F_SORTED_F2      // A/B: measured neutral (2.55 vs 2.54 ms: the kernel is latency-bound, not issue-bound)
                // packed pairs: the 64 FMAs of a point (4 dot products + 4 accumulator updates over CH channels) issue as
                // 32 FFMA2; a dot product becomes two interleaved partial sums
                unsigned long long go2[CH / 2];
#pragma unroll
                for (int c = 0; c < CH / 2; ++c) go2[c] = pack2(go[2 * c], go[2 * c + 1]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    unsigned long long t2 = pack2(0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < CH / 2; ++c) t2 = ffma2(go2[c], pack2(v[k][2 * c], v[k][2 * c + 1]), t2);
                    float lo, hi;
                    unpack2(t2, lo, hi);
                    t[k] = lo + hi;
                }
#else
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    t[k] = 0.f;
#pragma unroll
                    for (int c = 0; c < CH; ++c) t[k] = fmaf(go[c], v[k][c], t[k]);
                }
#endif
                float pa = hh * (hw * t[0] + lw * t[1]) + lh * (hw * t[2] + lw * t[3]);
                float px = hh * (t[1] - t[0]) + lh * (t[3] - t[2]);
                float py = hw * (t[2] - t[0]) + lw * (t[3] - t[1]);
#pragma unroll
                for (int o = LPP / 2; o > 0; o >>= 1) {
                    pa += __shfl_xor_sync(gmask, pa, o);
                    px += __shfl_xor_sync(gmask, px, o);
                    py += __shfl_xor_sync(gmask, py, o);
                }
                const float wy0 = a * hh, wy1 = a * lh;
                const float cw[4] = {wy0 * hw, wy0 * lw, wy1 * hw, wy1 * lw};
#ifdef BM2F_SORTED_F2      // A/B: measured neutral (2.55 vs 2.54 ms: the kernel is latency-bound, not issue-bound)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned long long cw2 = pack2(cw[k], cw[k]);
#pragma unroll
                    for (int c = 0; c < CH / 2; ++c) {
                        const unsigned long long a2 = ffma2(cw2, go2[c], pack2(acc[k][2 * c], acc[k][2 * c + 1]));
                        unpack2(a2, acc[k][2 * c], acc[k][2 * c + 1]);
                    }
                }
#else
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int c = 0; c < CH; ++c) acc[k][c] = fmaf(cw[k], go[c], acc[k][c]);
#endif
                if (sub == 0) s_rec[pt] = make_float4(pa, a * px, a * py, a);
            };

            if constexpr (PIPE == 2 || PIPE == 3) {
                // cell-strided ownership: the exclusive prefix sums left in the cell table by the scan give every cell's
                // run [off, end) of the sorted list; all points of a run share one anchor (and its corner mask)
                auto run_of = [&](int cell, int &off, int &end) {
                    off = static_cast<int>(cell_offset<NW>(s_cnt, cell));
                    end = cell + 1 < ncells ? static_cast<int>(cell_offset<NW>(s_cnt, cell + 1)) : nsorted;
                };
                if constexpr (PIPE == 2) {
                    for (int cell = g; cell < ncells; cell += G) {
                        int off, end;
                        run_of(cell, off, end);
                        if (end <= off) continue;
                        cur = __float_as_int(s_rec[s_perm[off]].w);
                        decode(cur, cmask, cW, e00);
                        load_lines(v, cmask, cW, e00);
                        for (int k = off; k < end; ++k) {
                            const int pt = s_perm[k];
                            process(pt, s_rec[pt]);
                        }
                        flush();
                    }
                } else {
                    // next non-empty own cell at or after `cell`
                    auto next_run = [&](int cell, int &off, int &end) {
                        for (; cell < ncells; cell += G) {
                            run_of(cell, off, end);
                            if (end > off) return cell;
                        }
                        return ncells;
                    };
                    int off, end;
                    int cell = next_run(g, off, end);
                    if (cell < ncells) {
                        cur = __float_as_int(s_rec[s_perm[off]].w);
                        decode(cur, cmask, cW, e00);
                        load_lines(v, cmask, cW, e00);
                    }
                    while (cell < ncells) {
                        int noff = 0, nend = 0;
                        const int ncell = next_run(cell + G, noff, nend);
                        float vn[4][CH];
                        int nmask = 0, nW = 0, npacked = kSortedSkip;
                        long long ne = 0;
                        if (ncell < ncells) {                 // request the next run's lines before working on this one
                            npacked = __float_as_int(s_rec[s_perm[noff]].w);
                            decode(npacked, nmask, nW, ne);
                            load_lines(vn, nmask, nW, ne);
                        }
                        for (int k = off; k < end; ++k) {
                            const int pt = s_perm[k];
                            process(pt, s_rec[pt]);
                        }
                        flush();
                        if (ncell < ncells) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
#pragma unroll
                                for (int c = 0; c < CH; ++c) v[k][c] = vn[k][c];
                            cur = npacked; cmask = nmask; cW = nW; e00 = ne;
                        }
                        cell = ncell; off = noff; end = nend;
                    }
                }
            } else if constexpr (PIPE == 1) {
                // the value lines of the NEXT anchor are requested while the current point is processed: the global
                // (L2) latency of an anchor change is overlapped instead of stalling the group at every change
                if (i < iend) {
                    int pt = s_perm[i];
                    float4 rec = s_rec[pt];
                    cur = __float_as_int(rec.w);
                    decode(cur, cmask, cW, e00);
                    load_lines(v, cmask, cW, e00);
                    for (; i < iend; ++i) {
                        const bool has_n = i + 1 < iend;
                        int pt_n = 0;
                        float4 rec_n = make_float4(0.f, 0.f, 0.f, 0.f);
                        int packed_n = kSortedSkip;
                        if (has_n) {
                            pt_n = s_perm[i + 1];
                            rec_n = s_rec[pt_n];
                            packed_n = __float_as_int(rec_n.w);
                        }
                        const bool trans = packed_n != cur;
                        float vn[4][CH];
                        int nmask = 0, nW = 0;
                        long long ne = 0;
                        if (trans && has_n) {
                            decode(packed_n, nmask, nW, ne);
                            load_lines(vn, nmask, nW, ne);
                        }
                        process(pt, rec);
                        if (trans) {
                            flush();
                            if (has_n) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
#pragma unroll
                                    for (int c = 0; c < CH; ++c) v[k][c] = vn[k][c];
                                cur = packed_n; cmask = nmask; cW = nW; e00 = ne;
                            }
                        }
                        pt = pt_n;
                        rec = rec_n;
                    }
                }
            } else {
                for (; i < iend; ++i) {
                    const int pt = s_perm[i];
                    const float4 rec = s_rec[pt];
                    const int packed = __float_as_int(rec.w);
                    if (packed != cur) {
                        if (cur != kSortedSkip) flush();
                        cur = packed;
                        decode(cur, cmask, cW, e00);
                        load_lines(v, cmask, cW, e00);
                    }
                    process(pt, rec);
                }
                if (cur != kSortedSkip) flush();
            }
        }
        __syncthreads();
        if (prof) { const long long c = clock64(); pt_[5] += c - t_prev; t_prev = c; pt_[7] += nsorted; }   // phase 2
        // grad_output rows are free: prefetch the next chunk's
        if (tid == 0 && have_next) issue_go(next);

        // ---- phase 3: one thread per query writes the gradients of its L*P points ----
        for (int qi = tid; qi < QMAX; qi += NT) {
            const int r = qi >> 5, xi = qi & 31;
            if (r >= job.nrows || xi >= sorted_row_cols(tabs, p, job, r)) continue;
            const int q = job.qstart + (job.y0 + r) * job.Wq + job.x0 + xi;
            const size_t qm = (static_cast<size_t>(job.b) * p.Lq + q) * kHeads + job.m;
            float4 rc[LP];
#pragma unroll
            for (int k = 0; k < LP; ++k) rc[k] = s_rec[qi * LP + k];
            // packed fused layout: gradients of offsets / logits are column blocks of one (N * Lq, ld_packed) matrix
            const size_t row = static_cast<size_t>(job.b) * p.Lq + q;
            float4 *ga4 = reinterpret_cast<float4 *>(p.grad_attn + (p.ld_packed ? row * p.ld_packed + job.m * LP : qm * LP));
            float4 *gl4 = reinterpret_cast<float4 *>(p.grad_loc + (p.ld_packed ? row * p.ld_packed + job.m * LP * 2 : qm * LP * 2));
            if constexpr (FUSED) {
                // softmax backward: grad_logit_i = a_i * (ga_i - sum_j a_j ga_j); d loc / d offset = 1 / (W, H) cancels
                float dot = 0.f;
#pragma unroll
                for (int k = 0; k < LP; ++k) dot = fmaf(rc[k].w, rc[k].x, dot);
#pragma unroll
                for (int l = 0; l < L_; ++l) {
                    ga4[l] = make_float4(rc[4 * l].w * (rc[4 * l].x - dot), rc[4 * l + 1].w * (rc[4 * l + 1].x - dot),
                                         rc[4 * l + 2].w * (rc[4 * l + 2].x - dot), rc[4 * l + 3].w * (rc[4 * l + 3].x - dot));
                    gl4[2 * l] = make_float4(rc[4 * l].y, rc[4 * l].z, rc[4 * l + 1].y, rc[4 * l + 1].z);
                    gl4[2 * l + 1] = make_float4(rc[4 * l + 2].y, rc[4 * l + 2].z, rc[4 * l + 3].y, rc[4 * l + 3].z);
                }
            } else {
#pragma unroll
                for (int l = 0; l < L_; ++l) {
                    ga4[l] = make_float4(rc[4 * l].x, rc[4 * l + 1].x, rc[4 * l + 2].x, rc[4 * l + 3].x);
                    gl4[2 * l] = make_float4(Wf[l] * rc[4 * l].y, Hf[l] * rc[4 * l].z, Wf[l] * rc[4 * l + 1].y, Hf[l] * rc[4 * l + 1].z);
                    gl4[2 * l + 1] = make_float4(Wf[l] * rc[4 * l + 2].y, Hf[l] * rc[4 * l + 2].z, Wf[l] * rc[4 * l + 3].y,
                                                 Hf[l] * rc[4 * l + 3].z);
                }
            }
        }
        if (prof) { const long long c = clock64(); pt_[6] += c - t_prev; t_prev = c; }     // phase 3 (thread 0's share)
        // the next iteration's barrier (after the counters are cleared) orders these record reads before phase 1 rewrites them
    }
    if (prof)
        for (int i = 0; i < 8; ++i) prof_out[blockIdx.x * 8 + i] = pt_[i];
}

}  // namespace bm2f

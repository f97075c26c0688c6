// Backward of multi-scale deformable attention, pixel-owner formulation (sm_100a, D = 32, P = 4, fp32).
//
// The per-corner kernel (msda_bwd_fast_kernel) issues one 128-byte L2 reduction per bilinear corner: 48 per
// (query, head), ~44 landing on every grad_value line, and it sits on the L2 reduction rate.  Here the corners of a
// tile of queries are sorted by the pixel they land on, and ONE THREAD owns a pixel: it keeps the pixel's 32 value
// channels and 32 gradient accumulators in registers, walks the pixel's corner list, and touches L2 once per pixel
// and chunk — no cross-lane traffic at all in the inner loop.
//
//   chunk   = (image b, head m, tile of TW x TH queries of one level), TW <= 32, TH <= RMAX; sampling locations /
//             attention weights (TMA boxes of 32 queries x L*P) and grad_output rows (32 x 32 floats) in shared memory.
//   phase 1 = one thread per sampling point: bilinear footprint (computed once, not in 8 channel lanes), the four
//             corner weights attn * bilinear written to w[point][corner]; every valid corner is counted into the cell
//             of its pixel inside a per-level window around the tile (shared-memory atomic returns its rank).
//   scan    = exclusive prefix sum over the cells; scatter: corner ids sorted by pixel.
//   phase 2 = warps take blocks of 32 consecutive cells, lane = pixel.  Per pixel: eight 128-bit loads of its value
//             line; per corner record: the query's grad_output row from shared memory (eight LDS.128, bank-conflict
//             free because lane L walks the row's 16-byte chunks in the order (s + L) mod 8 and keeps its registers
//             rotated the same way), 32 FMAs into the dot product <grad_out, value> and 32 FMAs into the accumulators;
//             the dot product replaces the weight in w[point][corner].  Then eight vector REDs flush the pixel.
//   phase 3 = one thread per point combines its four corner dot products into grad_attn / grad_loc terms; one thread
//             per query writes them (fused mode: softmax backward).
//
// Points whose footprint leaves the window (offsets beyond the margin, or queries that are not the pixels of the
// levels) are handled on the spot by their warp, one corner line at a time: correct for any input, fast for the
// locality deformable attention has.
//
// Reference semantics: ms_deform_attn_col2im_bilinear and ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1
// (/root/reference/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:92-164, 306-408).
#pragma once

#include "msda_bwd_sorted.cuh"

namespace bm2f {

struct PixelWin {      // per chunk and sampled level: anchor window (top-left corners) and the pixel window it implies
    int ax0[4], ay0[4], aw[4], ah[4];      // anchors (x0, y0) with ax0 <= x0 < ax0 + aw take the sorted path
    int px0[4], py0[4], pw[4], ph[4];      // pixels those anchors can touch, clipped to the level
    int base[4];                           // first cell of the level's pixel window
    int ncells;
};

template <int L_, int RMAX, int CELLS_MAX>
struct PixelSmem {
    static constexpr int LP = L_ * 4;
    static constexpr int QMAX = 32 * RMAX;
    static constexpr int NPT = QMAX * LP;
    static constexpr int NREC = NPT * 4;
    static constexpr int kLoc = 0;                                 // RMAX boxes of 32 x LP float2
    static constexpr int kAttn = kLoc + QMAX * LP * 8;             // RMAX boxes of 32 x LP float
    static constexpr int kGo = kAttn + QMAX * LP * 4;              // RMAX boxes of 32 x 32 float
    static constexpr int kW = kGo + QMAX * 128;                    // NPT float4: corner weights -> dot products -> point gradients
    static constexpr int kPerm = kW + NPT * 16;                    // NREC uint16 corner ids, sorted by pixel
    static constexpr int kCnt = kPerm + ((NREC * 2 + 15) / 16) * 16;   // CELLS_MAX 16-bit counts -> offsets, two per word
    static constexpr int kBytes = kCnt + CELLS_MAX * 2;
};

template <int L_>
__device__ __forceinline__ int pixel_cells_upper(const SortedTabs &t, int k, int tw, int th, int marg)
{
    int cells = 0;
    for (int l = 0; l < L_; ++l) {
        int ww = (tw * t.W[l] + t.W[k] - 1) / t.W[k] + 2 * marg + 4;
        int wh = (th * t.H[l] + t.H[k] - 1) / t.H[k] + 2 * marg + 4;
        ww = min(ww, t.W[l]);
        wh = min(wh, t.H[l]);
        cells += ww * wh;
    }
    return cells;
}

template <int L_, int RMAX, int CELLS_MAX>
__device__ __forceinline__ void pixel_build_tabs(SortedTabs &t, const FastParams &p, int marg)
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
            const int nx = max((t.W[k] + 31) / 32, 1);
            int tw = max((t.W[k] + nx - 1) / nx, 1), th = max(min(RMAX, t.H[k]), 1);
            // the pixel windows of all sampled levels must fit the cell table: shrink the tile until they do
            while (pixel_cells_upper<L_>(t, k, tw, th, marg) > CELLS_MAX && (th > 1 || tw > 1)) {
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

template <int L_>
__device__ __forceinline__ void pixel_window(PixelWin &w, const SortedTabs &t, const SortedJob &job, int marg, int cells_max)
{
    int base = 0;
    for (int l = 0; l < L_; ++l) {
        int aw = 0, ah = 0, lox = 0, loy = 0, px0 = 0, py0 = 0, pw = 0, ph = 0;
        if (!t.flat && t.W[l] > 0 && t.H[l] > 0) {
            const int Wk = t.W[job.k], Hk = t.H[job.k];
            lox = max((job.x0 * t.W[l]) / Wk - marg - 1, -1);
            loy = max((job.y0 * t.H[l]) / Hk - marg - 1, -1);
            const int hix = min(((job.x0 + job.ncols) * t.W[l] + Wk - 1) / Wk + marg, t.W[l] - 1);
            int hiy = min(((job.y0 + job.nrows) * t.H[l] + Hk - 1) / Hk + marg, t.H[l] - 1);
            aw = max(hix - lox + 1, 0);
            px0 = max(lox, 0);
            pw = min(hix + 1, t.W[l] - 1) - px0 + 1;
            py0 = max(loy, 0);
            // rows that do not fit the cell table any more are cut off (cannot happen with the tile search of build_tabs)
            while (hiy >= loy && base + pw * (min(hiy + 1, t.H[l] - 1) - py0 + 1) > cells_max) --hiy;
            ah = max(hiy - loy + 1, 0);
            ph = ah > 0 ? min(hiy + 1, t.H[l] - 1) - py0 + 1 : 0;
            if (aw <= 0 || ah <= 0) { aw = 0; ah = 0; pw = 0; ph = 0; }
        }
        w.ax0[l] = lox; w.ay0[l] = loy; w.aw[l] = aw; w.ah[l] = ah;
        w.px0[l] = px0; w.py0[l] = py0; w.pw[l] = pw; w.ph[l] = ph; w.base[l] = base;
        base += pw * ph;
    }
    w.ncells = base;
}

template <int L_, int RMAX, int NWARP, bool FUSED, int CELLS_MAX, int CPS, bool F2 = true>
__global__ void __launch_bounds__(NWARP * 32, CPS)
msda_bwd_pixel_kernel(const FastParams p, const int marg, long long *prof_out, const __grid_constant__ CUtensorMap tm_loc,
                      const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_go)
{
    constexpr int P_ = 4, LP = L_ * P_, NT = NWARP * 32, D = 32, MD = kHeads * D;
    using SM = PixelSmem<L_, RMAX, CELLS_MAX>;
    constexpr int QMAX = SM::QMAX;
    constexpr int PTS_PER_LEVEL = QMAX * P_;
    constexpr int ITERS = (PTS_PER_LEVEL + NT - 1) / NT;
    constexpr int NW = CELLS_MAX / 2;
    static_assert((NW & (NW - 1)) == 0, "cell words: a power of two");
    static_assert(SM::NREC < 65536, "16-bit counts / offsets / corner ids");

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ SortedTabs tabs;
    __shared__ PixelWin win;
    __shared__ alignas(8) uint64_t bar_la, bar_go;
    __shared__ uint32_t warp_sums[NWARP];
    __shared__ int s_next;

    float2 *s_loc = reinterpret_cast<float2 *>(smem + SM::kLoc);
    float *s_attn = reinterpret_cast<float *>(smem + SM::kAttn);
    float *s_go = reinterpret_cast<float *>(smem + SM::kGo);
    float4 *s_w4 = reinterpret_cast<float4 *>(smem + SM::kW);
    float *s_w = reinterpret_cast<float *>(smem + SM::kW);
    uint16_t *s_perm = reinterpret_cast<uint16_t *>(smem + SM::kPerm);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem + SM::kCnt);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        pixel_build_tabs<L_, RMAX, CELLS_MAX>(tabs, p, marg);
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
    float *grad_value = static_cast<float *>(p.grad_value);

    // optional phase profile (tools/bwd_phases.py): cycles thread 0 spends in each phase, summed over this CTA's chunks
    const bool prof = prof_out != nullptr && tid == 0;
    long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = prof ? clock64() : 0;
#define BM2F_PROF(i) if (prof) { const long long c_ = clock64(); pt_[i] += c_ - t_prev; t_prev = c_; }

    // footprint of sampling point (qi, l, pp) of this chunk; `live`: contributes at all; `inside`: sorted path
    struct Pt {
        Footprint f;
        float a;
        int mask, pix;
        bool live, inside;
    };

    uint32_t phase = 0;
    for (int j = blockIdx.x; j < total_jobs; j += gridDim.x, phase ^= 1) {
        const SortedJob job = sorted_decode_job(tabs, p, j);
        const bool have_next = j + static_cast<int>(gridDim.x) < total_jobs;
        // ---- A: windows of this chunk, clear the cell counters ----
        if (tid == 0) { pixel_window<L_>(win, tabs, job, marg, CELLS_MAX); s_next = 0; }
        for (int i = tid * 4; i < NW; i += NT * 4) *reinterpret_cast<uint4 *>(s_cnt + i) = make_uint4(0, 0, 0, 0);
        __syncthreads();
        int ax0[L_], ay0[L_], aw[L_], ah[L_], px0[L_], py0[L_], pw[L_], pbase[L_];
#pragma unroll
        for (int l = 0; l < L_; ++l) {
            ax0[l] = win.ax0[l]; ay0[l] = win.ay0[l]; aw[l] = win.aw[l]; ah[l] = win.ah[l];
            px0[l] = win.px0[l]; py0[l] = win.py0[l]; pw[l] = win.pw[l]; pbase[l] = win.base[l];
        }
        const int ncells = win.ncells;
        const size_t img = (static_cast<size_t>(job.b) * p.S * kHeads + job.m) * D;
        const float *vbm = value + img;
        float *gbm = grad_value + img;

        auto point = [&](int l, int qi, int pp) {
            Pt o;
            const int pt = qi * LP + l * P_ + pp;
            const int r = qi >> 5, xi = qi & 31;
            const bool valid = qi < QMAX && r < job.nrows && xi < sorted_row_cols(tabs, p, job, r);
            const float2 xy = s_loc[valid ? pt : 0];
            o.a = s_attn[valid ? pt : 0];
            o.f = make_footprint(xy.x, xy.y, H[l], W[l], Hf[l], Wf[l]);
            o.mask = (o.f.ok[0] ? 1 : 0) | (o.f.ok[1] ? 2 : 0) | (o.f.ok[2] ? 4 : 0) | (o.f.ok[3] ? 8 : 0);
            o.live = valid && o.mask != 0;
            o.pix = st[l] + o.f.y0 * W[l] + o.f.x0;                 // anchor pixel (may lie one row / column outside)
            o.inside = o.live && static_cast<unsigned>(o.f.x0 - ax0[l]) < static_cast<unsigned>(aw[l]) &&
                       static_cast<unsigned>(o.f.y0 - ay0[l]) < static_cast<unsigned>(ah[l]);
            return o;
        };

        mbar_wait(&bar_la, phase);
        BM2F_PROF(0)

        // ---- phase 0 (fused): softmax over each query's logits, loc = ref + offset / (W, H); lane = query ----
        if constexpr (FUSED) {
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
        // per point carried to the scatter: first cell (of corner 0) | mask, and the four 16-bit ranks
        uint32_t c_cell[L_ * ITERS], c_r01[L_ * ITERS], c_r23[L_ * ITERS];
        bool go_ready = false;
#pragma unroll
        for (int l = 0; l < L_; ++l) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const int idx = it * NT + tid;
                const int qi = idx >> 2, pp = idx & 3;
                const bool in_loop = idx < PTS_PER_LEVEL;
                const Pt o = point(l, in_loop ? qi : QMAX, pp);
                const int pt = qi * LP + l * P_ + pp;
                const float hh = o.f.hh, hw = o.f.hw, lh = o.f.lh, lw = o.f.lw;
                uint32_t cc = 0xffffffffu, r01 = 0, r23 = 0;
                float4 rec = make_float4(0.f, 0.f, 0.f, o.a);             // final form {ga, gx, gy, a}: skipped points
                if (o.inside) {
                    // cell of corner 0 (may be outside the pixel window when the anchor column / row is -1: such corners are masked)
                    const int c0 = pbase[l] + (o.f.y0 - py0[l]) * pw[l] + (o.f.x0 - px0[l]);
                    uint32_t rk[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((o.mask >> k) & 1) rk[k] = cell_count_rank<NW>(s_cnt, c0 + (k & 1) + (k >> 1) * pw[l]);
                    cc = (static_cast<uint32_t>(c0 + pw[l] + 1) << 4) | static_cast<uint32_t>(o.mask);   // biased: c0 >= -(pw + 1)
                    r01 = rk[0] | (rk[1] << 16);
                    r23 = rk[2] | (rk[3] << 16);
                    const float wy0 = o.a * hh, wy1 = o.a * lh;
                    rec = make_float4(o.f.ok[0] ? wy0 * hw : 0.f, o.f.ok[1] ? wy0 * lw : 0.f, o.f.ok[2] ? wy1 * hw : 0.f,
                                      o.f.ok[3] ? wy1 * lw : 0.f);
                }
                // points outside the window: this warp processes them now, one corner line (32 channels) at a time
                unsigned fb = __ballot_sync(0xffffffffu, o.live && !o.inside);
                if (fb) {
                    if (!go_ready) { mbar_wait(&bar_go, phase); go_ready = true; }
                    float ga = 0.f, gx = 0.f, gy = 0.f;
                    while (fb) {
                        const int src = __ffs(fb) - 1;
                        fb &= fb - 1;
                        const float a_ = __shfl_sync(0xffffffffu, o.a, src);
                        const float lh_ = __shfl_sync(0xffffffffu, lh, src), lw_ = __shfl_sync(0xffffffffu, lw, src);
                        const int pix_ = __shfl_sync(0xffffffffu, o.pix, src);
                        const int mask_ = __shfl_sync(0xffffffffu, o.mask, src);
                        const int qi_ = __shfl_sync(0xffffffffu, qi, src);
                        const float hh_ = 1.f - lh_, hw_ = 1.f - lw_;
                        const float g = s_go[qi_ * D + lane];
                        const float cw[4] = {hh_ * hw_, hh_ * lw_, lh_ * hw_, lh_ * lw_};
                        const int poff[4] = {0, 1, W[l], W[l] + 1};
                        float t[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            t[k] = 0.f;
                            if ((mask_ >> k) & 1) {
                                const long long e = static_cast<long long>(pix_ + poff[k]) * MD + lane;
                                t[k] = g * __ldg(vbm + e);
                                const float rv[1] = {a_ * cw[k] * g};
                                VecIO<float, 1>::red_add(gbm + e, rv);
                            }
                        }
                        float pa = hh_ * (hw_ * t[0] + lw_ * t[1]) + lh_ * (hw_ * t[2] + lw_ * t[3]);
                        float px = hh_ * (t[1] - t[0]) + lh_ * (t[3] - t[2]);
                        float py = hw_ * (t[2] - t[0]) + lw_ * (t[3] - t[1]);
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1) {
                            pa += __shfl_xor_sync(0xffffffffu, pa, o2);
                            px += __shfl_xor_sync(0xffffffffu, px, o2);
                            py += __shfl_xor_sync(0xffffffffu, py, o2);
                        }
                        if (lane == src) { ga = pa; gx = a_ * px; gy = a_ * py; }
                    }
                    if (o.live && !o.inside) rec = make_float4(ga, gx, gy, o.a);
                }
                if (in_loop) s_w4[pt] = rec;
                c_cell[l * ITERS + it] = cc;
                c_r01[l * ITERS + it] = r01;
                c_r23[l * ITERS + it] = r23;
            }
        }
        __syncthreads();
        BM2F_PROF(1)

        // ---- scan + scatter: corner ids sorted by pixel cell ----
        const int nsorted = sorted_block_scan<NT, NW>(s_cnt, ncells, warp_sums, tid);
        BM2F_PROF(2)
#pragma unroll
        for (int l = 0; l < L_; ++l) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t cc = c_cell[l * ITERS + it];
                if (cc != 0xffffffffu) {
                    const int idx = it * NT + tid;
                    const int pt = (idx >> 2) * LP + l * P_ + (idx & 3);
                    const int c0 = static_cast<int>(cc >> 4) - (pw[l] + 1);
                    const uint32_t rk[4] = {c_r01[l * ITERS + it] & 0xffffu, c_r01[l * ITERS + it] >> 16,
                                            c_r23[l * ITERS + it] & 0xffffu, c_r23[l * ITERS + it] >> 16};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((cc >> k) & 1)
                            s_perm[cell_offset<NW>(s_cnt, c0 + (k & 1) + (k >> 1) * pw[l]) + rk[k]] =
                                static_cast<uint16_t>(pt * 4 + k);
                }
            }
        }
        __syncthreads();
        BM2F_PROF(3)

        // ---- phase 2: a thread owns a pixel ----
        if (!go_ready) mbar_wait(&bar_go, phase);
        BM2F_PROF(4)
        for (;;) {
            int blk = 0;
            if (lane == 0) blk = atomicAdd(&s_next, 1);
            blk = __shfl_sync(0xffffffffu, blk, 0);
            if (blk * 32 >= ncells) break;
            const int c = blk * 32 + lane;
            int beg = nsorted, end = nsorted;
            if (c < ncells) {
                beg = static_cast<int>(cell_offset<NW>(s_cnt, c));
                end = c + 1 < ncells ? static_cast<int>(cell_offset<NW>(s_cnt, c + 1)) : nsorted;
            }
            const int cnt = end - beg;
            const int maxc = __reduce_max_sync(0xffffffffu, cnt);
            if (maxc == 0) continue;
            // cell -> level, pixel -> element offset of the pixel's 32 channels for this head
            int l = 0;
#pragma unroll
            for (int k = 1; k < L_; ++k) l = (c >= pbase[k]) ? k : l;
            int lw_ = pw[0], lx0 = px0[0], ly0 = py0[0], lb = pbase[0], lW = W[0], lst = st[0];
#pragma unroll
            for (int k = 1; k < L_; ++k)
                if (l == k) { lw_ = pw[k]; lx0 = px0[k]; ly0 = py0[k]; lb = pbase[k]; lW = W[k]; lst = st[k]; }
            const int rel = c - lb;
            const int cy = rel / max(lw_, 1), cx = rel - cy * lw_;
            const long long e = static_cast<long long>(lst + (ly0 + cy) * lW + lx0 + cx) * MD;
            // slot s of lane L holds channels 4 * ((s + L) & 7) ..+3: the shared-memory reads of a grad_output row then
            // touch eight different 16-byte columns within every quarter warp (no bank conflicts for any row mix)
            if constexpr (F2) {
                unsigned long long v2[8][2], a2[8][2];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cnt > 0) t4 = __ldg(reinterpret_cast<const float4 *>(vbm + e) + ((s + lane) & 7));
                    v2[s][0] = pack2(t4.x, t4.y); v2[s][1] = pack2(t4.z, t4.w);
                    a2[s][0] = 0ull; a2[s][1] = 0ull;
                }
                for (int jj = 0; jj < maxc; ++jj) {
                    if (jj < cnt) {
                        const int rid = s_perm[beg + jj];
                        const float wgt = s_w[rid];
                        const unsigned long long w2 = pack2(wgt, wgt);
                        const float *gor = s_go + (rid / (4 * LP)) * D;
                        unsigned long long t0 = 0ull, t1 = 0ull;
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const ulonglong2 g = *reinterpret_cast<const ulonglong2 *>(gor + 4 * ((s + lane) & 7));
                            t0 = ffma2(g.x, v2[s][0], t0);
                            t1 = ffma2(g.y, v2[s][1], t1);
                            a2[s][0] = ffma2(w2, g.x, a2[s][0]);
                            a2[s][1] = ffma2(w2, g.y, a2[s][1]);
                        }
                        float ta, tb, tc, td;
                        unpack2(t0, ta, tb);
                        unpack2(t1, tc, td);
                        s_w[rid] = (ta + tb) + (tc + td);
                    }
                }
                if (cnt > 0) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        float r[4];
                        unpack2(a2[s][0], r[0], r[1]);
                        unpack2(a2[s][1], r[2], r[3]);
                        VecIO<float, 4>::red_add(gbm + e + 4 * ((s + lane) & 7), r);
                    }
                }
            } else {
            float v[8][4], acc[8][4];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cnt > 0) t4 = __ldg(reinterpret_cast<const float4 *>(vbm + e) + ((s + lane) & 7));
                v[s][0] = t4.x; v[s][1] = t4.y; v[s][2] = t4.z; v[s][3] = t4.w;
                acc[s][0] = 0.f; acc[s][1] = 0.f; acc[s][2] = 0.f; acc[s][3] = 0.f;
            }
            for (int jj = 0; jj < maxc; ++jj) {
                if (jj < cnt) {
                    const int rid = s_perm[beg + jj];
                    const float wgt = s_w[rid];
                    const float *gor = s_go + (rid / (4 * LP)) * D;
                    float t = 0.f;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const float4 g = *reinterpret_cast<const float4 *>(gor + 4 * ((s + lane) & 7));
                        t = fmaf(g.x, v[s][0], t); t = fmaf(g.y, v[s][1], t);
                        t = fmaf(g.z, v[s][2], t); t = fmaf(g.w, v[s][3], t);
                        acc[s][0] = fmaf(wgt, g.x, acc[s][0]); acc[s][1] = fmaf(wgt, g.y, acc[s][1]);
                        acc[s][2] = fmaf(wgt, g.z, acc[s][2]); acc[s][3] = fmaf(wgt, g.w, acc[s][3]);
                    }
                    s_w[rid] = t;
                }
            }
            if (cnt > 0) {
#pragma unroll
                for (int s = 0; s < 8; ++s) VecIO<float, 4>::red_add(gbm + e + 4 * ((s + lane) & 7), acc[s]);
            }
            }
        }
        __syncthreads();
        BM2F_PROF(5)
        if (prof) pt_[7] += nsorted;
        // grad_output rows are free: prefetch the next chunk's
        SortedJob next{};
        if (tid == 0 && have_next) {
            next = sorted_decode_job(tabs, p, j + gridDim.x);
            issue_go(next);
        }

        // ---- phase 3a: one thread per point turns its four corner dot products into {grad_attn, d/dx, d/dy, attn} ----
#pragma unroll
        for (int l = 0; l < L_; ++l) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const int idx = it * NT + tid;
                const int qi = idx >> 2, pp = idx & 3;
                const Pt o = point(l, idx < PTS_PER_LEVEL ? qi : QMAX, pp);
                if (o.inside) {                       // everything else was written in its final form by phase 1
                    const int pt = qi * LP + l * P_ + pp;
                    const float4 t = s_w4[pt];
                    const float hh = o.f.hh, hw = o.f.hw, lh = o.f.lh, lw = o.f.lw;
                    const float pa = hh * (hw * t.x + lw * t.y) + lh * (hw * t.z + lw * t.w);
                    const float px = hh * (t.y - t.x) + lh * (t.w - t.z);
                    const float py = hw * (t.z - t.x) + lw * (t.w - t.y);
                    s_w4[pt] = make_float4(pa, o.a * px, o.a * py, o.a);
                }
            }
        }
        __syncthreads();
        // loc / attn buffers are free: prefetch the next chunk's
        if (tid == 0 && have_next) issue_la(next);

        // ---- phase 3b: one thread per query writes the gradients of its L*P points ----
        for (int qi = tid; qi < QMAX; qi += NT) {
            const int r = qi >> 5, xi = qi & 31;
            if (r >= job.nrows || xi >= sorted_row_cols(tabs, p, job, r)) continue;
            const int q = job.qstart + (job.y0 + r) * job.Wq + job.x0 + xi;
            const size_t qm = (static_cast<size_t>(job.b) * p.Lq + q) * kHeads + job.m;
            float4 rc[LP];
#pragma unroll
            for (int k = 0; k < LP; ++k) rc[k] = s_w4[qi * LP + k];
            float4 *ga4 = reinterpret_cast<float4 *>(p.grad_attn + qm * LP);
            float4 *gl4 = reinterpret_cast<float4 *>(p.grad_loc + qm * LP * 2);
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
        BM2F_PROF(6)
        // the next iteration's barrier (after the counters are cleared) orders these reads before phase 1 rewrites the records
    }
#undef BM2F_PROF
    if (prof)
        for (int i = 0; i < 8; ++i) prof_out[blockIdx.x * 8 + i] = pt_[i];
}

}  // namespace bm2f

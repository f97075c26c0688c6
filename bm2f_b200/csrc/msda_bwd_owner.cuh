// Backward of multi-scale deformable attention, pixel-owner formulation with balanced lanes (sm_100a, D = 32, P = 4, fp32).
//
// The per-corner kernel (msda_bwd_fast_kernel) issues one 128-byte L2 reduction per bilinear corner — 48 per (query, head),
// ~44 landing on every grad_value line — and sits on the L2 reduction rate.  Here ONE THREAD owns a pixel of grad_value: it
// keeps the pixel's 32 value channels and 32 gradient accumulators in registers, walks the list of sampling points whose
// footprint touches the pixel, and touches L2 once per pixel and chunk.  What differs from msda_bwd_pixel_kernel (the first
// pixel-owner version, 3.3 ms: 30 % lane utilisation in the inner loop, per-lane line loads that cost the L1 one wavefront per
// lane, 48 shared-memory atomics per (query, head)):
//
//   * points are sorted by their ANCHOR (top-left pixel of the footprint): one shared-memory atomic per point instead of
//     one per corner.  A pixel's list is the concatenation of the lists of the four anchors it can be a corner of — two
//     contiguous ranges of the sorted array, because anchor cells are numbered row-major;
//   * pixels are then counting-sorted by the length of their list, and a warp takes 32 pixels of (almost) equal length:
//     the inner loop runs with all lanes busy;
//   * the 32 value lines of a block are loaded by the warp cooperatively (8 lanes x 16 B = one line per quarter warp, four
//     L1 wavefronts per instruction instead of 32) and transposed through a small shared-memory stage; the accumulators
//     go back the same way and leave as full-line vector reductions (REDG.E.ADD.F32x4, 8 lanes = one 128-B line);
//   * the sorted array holds one 16-byte record per point — pixel coordinates of the sample, attention weight, packed
//     (grad_output row offset | point id) — read with ONE sequential LDS.128 per list entry; the corner weight is the
//     bilinear "hat" a * (1 - |px - w_im|) * (1 - |py - h_im|), the same product the reference forms from lh / lw, and
//     the dot product goes to slot (py & 1, px & 1) of the point: the four corners of a footprint have four different
//     parities, so no per-entry corner bookkeeping is left in the inner loop.
//
//   chunk   = (image b, head m, tile of TW x TH queries of one level); locations / attention weights (TMA boxes of
//             32 queries x L*P) and grad_output rows (32 x 32 floats) staged in shared memory.
//   phase 1 = one thread per sampling point: footprint, anchor cell counted (ATOMS returns the rank).
//   scan    = exclusive prefix sum over the anchor cells; scatter: point records sorted by anchor (they overwrite the
//             staged locations / weights, which are dead by then).
//   pixels  = one thread per pixel of the window: list length from six cell offsets, histogram by length, scan, scatter.
//   phase 2 = warps take blocks of 32 pixels, lane = pixel: per list entry one record, eight conflict-free LDS.128 of the
//             query's grad_output row, 16 FFMA2 into <grad_out, value> and 16 FFMA2 into the accumulators.
//   phase 3 = one thread per sorted point folds its four corner dot products into grad_attn / grad_loc terms; one thread
//             per (query, level) writes them (fused mode: softmax backward).
//
// Points whose anchor lies outside the window (offsets beyond the margin) are handled on the spot by their warp, one
// corner line at a time: correct for any input, fast for the locality deformable attention has.
//
// Reference semantics: ms_deform_attn_col2im_bilinear and ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1
// (/root/reference/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:92-164, 306-408).
#pragma once

#include "msda_bwd_pixel.cuh"

namespace bm2f {

struct OwnerWin {      // per chunk and sampled level
    int ax0[4], ay0[4], aw[4], ah[4];      // anchors (x0, y0) with ax0 <= x0 < ax0 + aw, ay0 <= y0 < ay0 + ah are sorted
    int abase[4];                          // first anchor cell of the level (row-major, aw cells per row)
    int pbase[4];                          // first pixel cell: pixel grid (aw + 1) x (ah + 1) with origin (ax0, ay0)
    uint32_t pmagic[4];                    // ceil(2^32 / (aw + 1)): row of a pixel cell by one multiply-high
    int ncells, npix;
};

template <int L_, int RMAX, int CELLS_MAX, int NWARP>
struct OwnerSmem {
    static constexpr int LP = L_ * 4;
    static constexpr int QMAX = 32 * RMAX;
    static constexpr int NPT = QMAX * LP;
    static constexpr int kStageWarp = 2048;                        // 16 lines of 128 B per warp
    // [0, NPT * 16): TMA-staged locations (NPT float2) and weights (NPT float) until phase 1 has read them, then the
    // sorted point records (NPT float4)
    static constexpr int kLoc = 0;
    static constexpr int kAttn = kLoc + QMAX * LP * 8;
    static constexpr int kRec = 0;
    static constexpr int kGo = NPT * 16;                           // RMAX boxes of 32 x 32 float
    static constexpr int kDot = kGo + QMAX * 128;                  // NPT float4: corner dot products -> point gradients
    static constexpr int kCnt = kDot + NPT * 16;                   // CELLS_MAX 16-bit counts -> offsets, two per word
    static constexpr int kPix = kCnt + CELLS_MAX * 2;              // CELLS_MAX uint16 pixel cells sorted by list length
    static constexpr int kStage = kPix + CELLS_MAX * 2;            // per-warp transpose stage
    static constexpr int kBytes = kStage + NWARP * kStageWarp;
};

// shared-space accesses by 32-bit shared address (no generic-address arithmetic in the inner loop)
__device__ __forceinline__ float4 lds128(uint32_t sa)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa));
    return v;
}
__device__ __forceinline__ ulonglong2 lds128_u64(uint32_t sa)
{
    ulonglong2 v;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(sa));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t sa, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sa), "f"(v) : "memory");
}

template <int L_>
__device__ __forceinline__ int owner_cells_upper(const SortedTabs &t, int k, int tw, int th, int marg)
{
    int cells = 0;
    for (int l = 0; l < L_; ++l) {
        int ww = (tw * t.W[l] + t.W[k] - 1) / t.W[k] + 2 * marg + 3;      // anchors: span + 2 * marg + 2, pixels: + 1
        int wh = (th * t.H[l] + t.H[k] - 1) / t.H[k] + 2 * marg + 3;
        ww = min(ww, t.W[l] + 2);
        wh = min(wh, t.H[l] + 2);
        cells += ww * wh;
    }
    return cells;
}

template <int L_, int RMAX, int CELLS_MAX>
__device__ __forceinline__ void owner_build_tabs(SortedTabs &t, const FastParams &p, int marg)
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
            // the pixel windows of all sampled levels (+ one sentinel cell) must fit the cell tables: shrink the tile until they do
            while (owner_cells_upper<L_>(t, k, tw, th, marg) > CELLS_MAX - 1 && (th > 1 || tw > 1)) {
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
__device__ __forceinline__ void owner_window(OwnerWin &w, const SortedTabs &t, const SortedJob &job, int marg, int cells_max)
{
    int abase = 0, pbase = 0;
    for (int l = 0; l < L_; ++l) {
        int aw = 0, ah = 0, lox = 0, loy = 0;
        if (!t.flat && t.W[l] > 0 && t.H[l] > 0) {
            const int Wk = t.W[job.k], Hk = t.H[job.k];
            lox = max((job.x0 * t.W[l]) / Wk - marg - 1, -1);
            loy = max((job.y0 * t.H[l]) / Hk - marg - 1, -1);
            const int hix = min(((job.x0 + job.ncols) * t.W[l] + Wk - 1) / Wk + marg, t.W[l] - 1);
            int hiy = min(((job.y0 + job.nrows) * t.H[l] + Hk - 1) / Hk + marg, t.H[l] - 1);
            aw = max(hix - lox + 1, 0);
            // rows that do not fit the cell table any more are cut off (cannot happen with the tile search of build_tabs)
            while (hiy >= loy && pbase + (aw + 1) * (hiy - loy + 2) > cells_max - 1) --hiy;
            ah = max(hiy - loy + 1, 0);
            if (aw <= 0 || ah <= 0) { aw = 0; ah = 0; }
        }
        w.ax0[l] = lox; w.ay0[l] = loy; w.aw[l] = aw; w.ah[l] = ah;
        w.abase[l] = abase; w.pbase[l] = pbase;
        w.pmagic[l] = static_cast<uint32_t>((0x100000000ull + aw) / static_cast<unsigned>(aw + 1));
        abase += aw * ah;
        pbase += aw > 0 ? (aw + 1) * (ah + 1) : 0;
    }
    w.ncells = abase;
    w.npix = pbase;
}

template <int L_, int RMAX, int NWARP, bool FUSED, int CELLS_MAX, int CPS>
__global__ void __launch_bounds__(NWARP * 32, CPS)
msda_bwd_owner_kernel(const FastParams p, const int marg, long long *prof_out, const __grid_constant__ CUtensorMap tm_loc,
                      const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_go)
{
    constexpr int P_ = 4, LP = L_ * P_, NT = NWARP * 32, D = 32, MD = kHeads * D;
    using SM = OwnerSmem<L_, RMAX, CELLS_MAX, NWARP>;
    constexpr int QMAX = SM::QMAX;
    constexpr int PTS_PER_LEVEL = QMAX * P_;
    constexpr int ITERS = (PTS_PER_LEVEL + NT - 1) / NT;
    constexpr int NW = CELLS_MAX / 2;
    constexpr int NBIN = 64;
    static_assert((NW & (NW - 1)) == 0, "cell words: a power of two");
    static_assert(SM::NPT < 65536 && CELLS_MAX <= 65536 && QMAX * 128 < 65536, "16-bit counts / offsets / ids / row offsets");

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ SortedTabs tabs;
    __shared__ OwnerWin win;
    __shared__ alignas(8) uint64_t bar_la, bar_go;
    __shared__ uint32_t warp_sums[NWARP];
    __shared__ uint32_t s_bin[NBIN];
    __shared__ int s_next, s_npos;

    float2 *s_loc = reinterpret_cast<float2 *>(smem + SM::kLoc);
    float *s_attn = reinterpret_cast<float *>(smem + SM::kAttn);
    float4 *s_rec = reinterpret_cast<float4 *>(smem + SM::kRec);
    float *s_go = reinterpret_cast<float *>(smem + SM::kGo);
    float4 *s_dot4 = reinterpret_cast<float4 *>(smem + SM::kDot);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem + SM::kCnt);
    uint16_t *s_pix = reinterpret_cast<uint16_t *>(smem + SM::kPix);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float4 *stg = reinterpret_cast<float4 *>(smem + SM::kStage + warp * SM::kStageWarp);
    if (tid == 0) {
        owner_build_tabs<L_, RMAX, CELLS_MAX>(tabs, p, marg);
        mbar_init(&bar_la, 1);
        mbar_init(&bar_go, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tm_loc);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_go);
    }
    __syncthreads();

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

    // shared-space byte addresses used by the hand-scheduled inner loop
    const uint32_t rec_sa = smem_u32(smem + SM::kRec), dot_sa = smem_u32(smem + SM::kDot);
    // slot s of a lane holds channels 4 * (s ^ (lane & 7)) ..+3 of ITS pixel: the reads of a grad_output row then touch
    // eight different 16-byte columns within every quarter warp (no bank conflicts for any mix of rows), and the address
    // of slot s is (row + 16 * (lane & 7)) ^ (16 * s): one LOP3 with an immediate, no per-slot register (rows are
    // 128-byte aligned in the shared window)
    const uint32_t go_lane = smem_u32(smem + SM::kGo) + 16u * (lane & 7);

    // optional phase profile (tools/bwd_phases.py): cycles thread 0 spends in each phase, summed over this CTA's chunks
    const bool prof = prof_out != nullptr && tid == 0;
    long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = prof ? clock64() : 0;
#define BM2F_PROF(i) if (prof) { const long long c_ = clock64(); pt_[i] += c_ - t_prev; t_prev = c_; }

    uint32_t phase = 0;
    for (int j = blockIdx.x; j < total_jobs; j += gridDim.x, phase ^= 1) {
        const SortedJob job = sorted_decode_job(tabs, p, j);
        const bool have_next = j + static_cast<int>(gridDim.x) < total_jobs;
        // ---- A: windows of this chunk, clear the counters ----
        if (tid == 0) {
            owner_window<L_>(win, tabs, job, marg, CELLS_MAX);
            s_next = 0;
            if (have_next) {
                // the next chunk's boxes are requested into L2 now; the TMA loads issued after this chunk's phases then hit L2
                const SortedJob nj = sorted_decode_job(tabs, p, j + gridDim.x);
                for (int r = 0; r < nj.nrows; ++r) {
                    const int row = nj.b * p.Lq + nj.qstart + (nj.y0 + r) * nj.Wq + nj.x0;
                    tma_prefetch_l2_2d(&tm_loc, nj.m * LP * 2, row);
                    tma_prefetch_l2_2d(&tm_w, nj.m * LP, row);
                    tma_prefetch_l2_2d(&tm_go, nj.m * D, row);
                }
            }
        }
        for (int i = tid * 4; i < NW; i += NT * 4) *reinterpret_cast<uint4 *>(s_cnt + i) = make_uint4(0, 0, 0, 0);
        if (tid < NBIN) s_bin[tid] = 0;
        __syncthreads();
        const int ncells = win.ncells;
        const size_t img = (static_cast<size_t>(job.b) * p.S * kHeads + job.m) * D;
        const float *vbm = value + img;
        float *gbm = grad_value + img;

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
                    const float wq = static_cast<float>(tabs.W[job.k]), hq = static_cast<float>(tabs.H[job.k]);
                    centre = make_float2((static_cast<float>(job.x0 + xi) + 0.5f) / wq,
                                         (static_cast<float>(job.y0 + r) + 0.5f) / hq);
                }
#pragma unroll
                for (int l = 0; l < L_; ++l) {
                    const float2 rr = ref_q ? __ldg(ref_q + l) : centre;
                    const float rw = 1.f / static_cast<float>(tabs.W[l]), rh = 1.f / static_cast<float>(tabs.H[l]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 o = sl[2 * l + h];
                        o.x = fmaf(o.x, rw, rr.x); o.y = fmaf(o.y, rh, rr.y);
                        o.z = fmaf(o.z, rw, rr.x); o.w = fmaf(o.w, rh, rr.y);
                        sl[2 * l + h] = o;
                    }
                }
            }
            __syncthreads();
        }

        // ---- phase 1: one thread per sampling point ----
        // Real loops over (level, pass) with the level's parameters read from shared memory: the code of a pass exists
        // once (an unrolled version is 190 KB of SASS and spends a fifth of its cycles on instruction-cache misses).
        // A point inside its window leaves {w_im, h_im, attn, anchor cell << 16 | rank} in its dot-product slot for the
        // scatter pass below; every other point gets its final gradients {ga, gx, gy, attn} there.  One bit per point (a
        // ballot word per warp and pass, kept in the idle transpose stage) tells the two apart.
        uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + SM::kStage);
        bool go_ready = false;
#pragma unroll 1
        for (int l = 0; l < L_; ++l) {
            const int Hl = tabs.H[l], Wl = tabs.W[l], stl = tabs.start[l];
            const float Hfl = static_cast<float>(Hl), Wfl = static_cast<float>(Wl);
            const int axl = win.ax0[l], ayl = win.ay0[l], awl = win.aw[l], ahl = win.ah[l], abl = win.abase[l];
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
                const int idx = it * NT + tid;
                const int qi = idx >> 2, pp = idx & 3;
                const bool in_loop = idx < PTS_PER_LEVEL;
                const int pt = qi * LP + l * P_ + pp;
                const int r = qi >> 5, xi = qi & 31;
                const bool valid = in_loop && r < job.nrows && xi < sorted_row_cols(tabs, p, job, r);
                const float2 xy = s_loc[in_loop ? pt : 0];
                const float a = s_attn[in_loop ? pt : 0];
                const Footprint f = make_footprint(xy.x, xy.y, Hl, Wl, Hfl, Wfl);
                const int mask = (f.ok[0] ? 1 : 0) | (f.ok[1] ? 2 : 0) | (f.ok[2] ? 4 : 0) | (f.ok[3] ? 8 : 0);
                const bool live = valid && mask != 0;
                const int cx = f.x0 - axl, cy = f.y0 - ayl;
                const bool inside = live && static_cast<unsigned>(cx) < static_cast<unsigned>(awl) &&
                                    static_cast<unsigned>(cy) < static_cast<unsigned>(ahl);
                float4 rec = make_float4(0.f, 0.f, 0.f, a);              // final form {ga, gx, gy, a}: skipped points
                if (inside) {
                    const int cell = abl + cy * awl + cx;
                    const uint32_t rank = cell_count_rank<NW>(s_cnt, cell);
                    // the sample in pixel units, exactly as make_footprint forms it (lw = w_im - floor(w_im))
                    rec = make_float4(fmaf(xy.x, Wfl, -0.5f), fmaf(xy.y, Hfl, -0.5f), a,
                                      __uint_as_float((static_cast<uint32_t>(cell) << 16) | rank));
                }
                const unsigned inb = __ballot_sync(0xffffffffu, inside);
                if (lane == 0) s_mask[(l * ITERS + it) * NWARP + warp] = inb;
                // points outside the window: this warp processes them now, one corner line (32 channels) at a time
                unsigned fb = __ballot_sync(0xffffffffu, live && !inside);
                if (fb) {
                    if (!go_ready) { mbar_wait(&bar_go, phase); go_ready = true; }
                    const int pix = stl + f.y0 * Wl + f.x0;
                    float ga = 0.f, gx = 0.f, gy = 0.f;
                    while (fb) {
                        const int src = __ffs(fb) - 1;
                        fb &= fb - 1;
                        const float a_ = __shfl_sync(0xffffffffu, a, src);
                        const float lh_ = __shfl_sync(0xffffffffu, f.lh, src), lw_ = __shfl_sync(0xffffffffu, f.lw, src);
                        const int pix_ = __shfl_sync(0xffffffffu, pix, src);
                        const int mask_ = __shfl_sync(0xffffffffu, mask, src);
                        const int qi_ = __shfl_sync(0xffffffffu, qi, src);
                        const float hh_ = 1.f - lh_, hw_ = 1.f - lw_;
                        const float g = s_go[qi_ * D + lane];
                        const float cw[4] = {hh_ * hw_, hh_ * lw_, lh_ * hw_, lh_ * lw_};
                        const int poff[4] = {0, 1, Wl, Wl + 1};
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
                    if (live && !inside) rec = make_float4(ga, gx, gy, a);
                }
                if (in_loop) s_dot4[pt] = rec;
            }
        }
        __syncthreads();           // every location / weight has been read: the record array may overwrite them
        BM2F_PROF(1)

        // ---- scan over the anchor cells (+ one empty sentinel cell: its offset is the total), scatter the records ----
        const int nsorted = sorted_block_scan<NT, NW>(s_cnt, ncells + 1, warp_sums, tid);
        BM2F_PROF(2)
#pragma unroll 1
        for (int li = 0; li < L_ * ITERS; ++li) {
            const int l = li / ITERS, it = li - l * ITERS;
            const unsigned inb = s_mask[li * NWARP + warp];
            if ((inb >> lane) & 1u) {
                const int idx = it * NT + tid;
                const int qi = idx >> 2;
                const int pt = qi * LP + l * P_ + (idx & 3);
                const float4 stash = s_dot4[pt];
                const uint32_t c = __float_as_uint(stash.w);
                // {w_im, h_im, attention weight, byte offset of the query's grad_output row << 16 | point id}
                s_rec[cell_offset<NW>(s_cnt, c >> 16) + (c & 0xffffu)] =
                    make_float4(stash.x, stash.y, stash.z,
                                __uint_as_float((static_cast<uint32_t>(qi * (D * 4)) << 16) | static_cast<uint32_t>(pt)));
                s_dot4[pt] = make_float4(0.f, 0.f, 0.f, 0.f);            // corner dot products, filled by the pixel owners
            }
        }

        // ---- pixels: list = two contiguous ranges of the sorted points (anchor rows py - 1 and py); counting sort by length:
        //      the length is computed twice (histogram pass, placement pass) instead of being carried in registers ----
        auto list_length = [&](int rel, int gw, int gsz, int axl, int ayl, int awl, int ahl, int abl, int Wl, int Hl, uint32_t mg) {
            const int jy = static_cast<int>(__umulhi(static_cast<uint32_t>(rel), mg));
            const int ix = rel - jy * gw;
            const int px_ = axl + ix, py_ = ayl + jy;
            const bool ok = rel < gsz && px_ >= 0 && px_ < Wl && py_ >= 0 && py_ < Hl;
            const int xl = max(ix - 1, 0), xr = min(ix + 1, awl);
            int n = 0;
            if (ok && jy >= 1) {
                const int rb = abl + (jy - 1) * awl;
                n += static_cast<int>(cell_offset<NW>(s_cnt, rb + xr)) - static_cast<int>(cell_offset<NW>(s_cnt, rb + xl));
            }
            if (ok && jy < ahl) {
                const int rb = abl + jy * awl;
                n += static_cast<int>(cell_offset<NW>(s_cnt, rb + xr)) - static_cast<int>(cell_offset<NW>(s_cnt, rb + xl));
            }
            return n;
        };
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
            for (int l = 0; l < L_; ++l) {
                const int awl = win.aw[l], ahl = win.ah[l], axl = win.ax0[l], ayl = win.ay0[l], abl = win.abase[l];
                const int pbl = win.pbase[l], Wl = tabs.W[l], Hl = tabs.H[l];
                const uint32_t mg = win.pmagic[l];
                const int gw = awl + 1, gsz = awl > 0 ? gw * (ahl + 1) : 0;
#pragma unroll 1
                for (int rel = tid; rel < gsz; rel += NT) {
                    const int n = list_length(rel, gw, gsz, axl, ayl, awl, ahl, abl, Wl, Hl, mg);
                    if (n > 0) {
                        const uint32_t slot = atomicAdd(&s_bin[min(n, NBIN - 1)], 1u);
                        if (pass) s_pix[slot] = static_cast<uint16_t>(pbl + rel);
                    }
                }
            }
            __syncthreads();
            if (pass == 0) {
                BM2F_PROF(3)
                if (warp == 0) {
                    // longest lists first: exclusive prefix over the bins in descending order (two bins per lane)
                    const uint32_t c0 = s_bin[NBIN - 1 - 2 * lane], c1 = s_bin[NBIN - 2 - 2 * lane];
                    uint32_t incl = c0 + c1;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += u;
                    }
                    const uint32_t excl = incl - (c0 + c1);
                    s_bin[NBIN - 1 - 2 * lane] = excl;
                    s_bin[NBIN - 2 - 2 * lane] = excl + c0;
                    if (lane == 31) s_npos = static_cast<int>(incl - c1);      // bin 0 is never counted (c1 of lane 31 = 0)
                }
                __syncthreads();
            }
        }
        const int npos = s_npos;

        // ---- phase 2: a thread owns a pixel; a warp takes 32 pixels with lists of (almost) equal length ----
        if (!go_ready) mbar_wait(&bar_go, phase);
        BM2F_PROF(4)
        for (;;) {
            int blk = 0;
            if (lane == 0) blk = atomicAdd(&s_next, 1);
            blk = __shfl_sync(0xffffffffu, blk, 0);
            if (blk * 32 >= npos) break;
            const bool have = blk * 32 + lane < npos;
            // this lane's pixel: level, coordinates, element offset, the two ranges of the record array
            int n = 0, nA = 0;
            uint32_t ptrA = 0, ptrB = 0, dslot = 0;
            float pxf = 0.f, pyf = 0.f;
            int e = -1;                                  // element offset of the pixel's 32 channels within (image, head)
            if (have) {
                const int pc = s_pix[blk * 32 + lane];
                int l = 0;
#pragma unroll
                for (int k = 1; k < L_; ++k) l = (pc >= win.pbase[k]) ? k : l;
                const int aw_ = win.aw[l], ah_ = win.ah[l], ax_ = win.ax0[l], ay_ = win.ay0[l], ab_ = win.abase[l], pb_ = win.pbase[l];
                const int lW = tabs.W[l], lst = tabs.start[l];
                const uint32_t mg = win.pmagic[l];
                const int rel = pc - pb_;
                const int jy = static_cast<int>(__umulhi(static_cast<uint32_t>(rel), mg));
                const int ix = rel - jy * (aw_ + 1);
                const int px = ax_ + ix, py = ay_ + jy;
                const int xl = max(ix - 1, 0), xr = min(ix + 1, aw_);
                int begA = 0, endA = 0, begB = 0, endB = 0;
                if (jy >= 1) {
                    const int rb = ab_ + (jy - 1) * aw_;
                    begA = static_cast<int>(cell_offset<NW>(s_cnt, rb + xl));
                    endA = static_cast<int>(cell_offset<NW>(s_cnt, rb + xr));
                }
                if (jy < ah_) {
                    const int rb = ab_ + jy * aw_;
                    begB = static_cast<int>(cell_offset<NW>(s_cnt, rb + xl));
                    endB = static_cast<int>(cell_offset<NW>(s_cnt, rb + xr));
                }
                nA = endA - begA;
                n = nA + (endB - begB);
                ptrA = rec_sa + 16u * static_cast<uint32_t>(begA);
                ptrB = rec_sa + 16u * static_cast<uint32_t>(begB);
                if (nA == 0) ptrA = ptrB;
                pxf = static_cast<float>(px);
                pyf = static_cast<float>(py);
                dslot = dot_sa + 4u * static_cast<uint32_t>((py & 1) * 2 + (px & 1));
                e = (lst + py * lW + px) * MD;
            }
            const int maxn = __reduce_max_sync(0xffffffffu, n);
            // the block's 32 value lines: coalesced loads (a quarter warp = one 128-B line), transposed through the stage
            unsigned long long v2[8][2], a2[8][2];
            {
                float4 tmp[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int el = __shfl_sync(0xffffffffu, e, 4 * i + (lane >> 3));
                    tmp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (el >= 0) tmp[i] = __ldg(reinterpret_cast<const float4 *>(vbm + el) + (lane & 7));
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) stg[(4 * i + (lane >> 3)) * 8 + (lane & 7)] = tmp[4 * h + i];
                    __syncwarp();
                    if ((lane >> 4) == h) {
#pragma unroll
                        for (int s_ = 0; s_ < 8; ++s_) {
                            const float4 t4 = stg[(lane & 15) * 8 + (s_ ^ (lane & 7))];
                            v2[s_][0] = pack2(t4.x, t4.y); v2[s_][1] = pack2(t4.z, t4.w);
                        }
                    }
                    __syncwarp();
                }
            }
#pragma unroll
            for (int s_ = 0; s_ < 8; ++s_) { a2[s_][0] = 0ull; a2[s_][1] = 0ull; }

            // list walk: one record per entry, read one entry ahead of its use.  The lists of a block have (almost) the
            // same length: the first minn entries run without any predication, the few left over with it.
            const int minn = __reduce_min_sync(0xffffffffu, n);
            uint32_t ptr = ptrA;
            float4 rec_n = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n > 0) rec_n = lds128(ptr);
            auto entry = [&](const float4 rec) {
                // bilinear weight of this pixel in the sample's footprint (|d| <= 1 by construction of the lists)
                const float wx = 1.f - fabsf(pxf - rec.x);
                const float wy = 1.f - fabsf(pyf - rec.y);
                const float wgt = rec.z * wy * wx;
                const uint32_t pk = __float_as_uint(rec.w);
                const uint32_t row = go_lane + (pk >> 16);
                const unsigned long long w2 = pack2(wgt, wgt);
                ulonglong2 g[8];
#pragma unroll
                for (int s_ = 0; s_ < 8; ++s_) g[s_] = lds128_u64(row ^ (16u * s_));
                unsigned long long t0 = 0ull, t1 = 0ull;
#pragma unroll
                for (int s_ = 0; s_ < 8; ++s_) {
                    t0 = ffma2(g[s_].x, v2[s_][0], t0);
                    t1 = ffma2(g[s_].y, v2[s_][1], t1);
                    a2[s_][0] = ffma2(w2, g[s_].x, a2[s_][0]);
                    a2[s_][1] = ffma2(w2, g[s_].y, a2[s_][1]);
                }
                float ta, tb, tc, td;
                unpack2(t0, ta, tb);
                unpack2(t1, tc, td);
                sts32(dslot + 16u * (pk & 0xffffu), (ta + tb) + (tc + td));
            };
            int jj = 0;
#pragma unroll 1
            for (; jj < minn; ++jj) {
                const float4 rec = rec_n;
                ptr = (jj + 1 == nA) ? ptrB : ptr + 16u;
                if (jj + 1 < n) rec_n = lds128(ptr);
                entry(rec);
            }
#pragma unroll 1
            for (; jj < maxn; ++jj) {
                const float4 rec = rec_n;
                const bool act = jj < n;
                ptr = (jj + 1 == nA) ? ptrB : ptr + 16u;
                if (jj + 1 < n) rec_n = lds128(ptr);
                if (act) entry(rec);
            }
            // accumulators -> stage (own line, rotated) -> full-line vector reductions
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if ((lane >> 4) == h) {
#pragma unroll
                    for (int s_ = 0; s_ < 8; ++s_) {
                        float4 t4;
                        unpack2(a2[s_][0], t4.x, t4.y);
                        unpack2(a2[s_][1], t4.z, t4.w);
                        stg[(lane & 15) * 8 + (s_ ^ (lane & 7))] = t4;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ll = 4 * i + (lane >> 3);
                    const int el = __shfl_sync(0xffffffffu, e, 16 * h + ll);
                    const float4 t4 = stg[ll * 8 + (lane & 7)];
                    if (el >= 0) {
                        const float r[4] = {t4.x, t4.y, t4.z, t4.w};
                        VecIO<float, 4>::red_add(gbm + el + 4 * (lane & 7), r);
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        BM2F_PROF(5)
        if (prof) pt_[7] += npos;
        // grad_output rows are free: prefetch the next chunk's
        SortedJob next{};
        if (tid == 0 && have_next) {
            next = sorted_decode_job(tabs, p, j + gridDim.x);
            issue_go(next);
        }

        // ---- phase 3a: one thread per sorted point turns its corner dot products into {grad_attn, d/dx, d/dy, attn} ----
        for (int i = tid; i < nsorted; i += NT) {
            const float4 rec = s_rec[i];
            const int pt = static_cast<int>(__float_as_uint(rec.w) & 0xffffu);
            const float4 t = s_dot4[pt];                // slot (py & 1) * 2 + (px & 1) of the corner's pixel
            const float fx = floorf(rec.x), fy = floorf(rec.y);
            const float lw = rec.x - fx, lh = rec.y - fy, hw = 1.f - lw, hh = 1.f - lh;
            const bool ox = (static_cast<int>(fx) & 1) != 0, oy = (static_cast<int>(fy) & 1) != 0;
            // corner (dy, dx) sits in slot ((y0 + dy) & 1, (x0 + dx) & 1)
            const float r0x = ox ? t.y : t.x, r0y = ox ? t.x : t.y;      // slots of parity row 0: columns x0, x0 + 1
            const float r1x = ox ? t.w : t.z, r1y = ox ? t.z : t.w;      // slots of parity row 1
            const float t00 = oy ? r1x : r0x, t01 = oy ? r1y : r0y;      // corners of row y0
            const float t10 = oy ? r0x : r1x, t11 = oy ? r0y : r1y;      // corners of row y0 + 1
            const float a = rec.z;
            const float pa = hh * (hw * t00 + lw * t01) + lh * (hw * t10 + lw * t11);
            const float dx = hh * (t01 - t00) + lh * (t11 - t10);
            const float dy = hw * (t10 - t00) + lw * (t11 - t01);
            s_dot4[pt] = make_float4(pa, a * dx, a * dy, a);
        }
        fence_proxy_async_smem();      // the record / location region was written by the generic proxy and is a TMA target again
        __syncthreads();
        // the record region is free: load the next chunk's locations / weights
        if (tid == 0 && have_next) issue_la(next);

        // ---- phase 3b: one thread per (query, level) writes the gradients of its P points: three consecutive lanes write the
        //      48 / 96 contiguous bytes of a (query, head) — a warp instruction touches ~11 lines instead of 32 ----
        for (int t_ = tid; t_ < QMAX * L_; t_ += NT) {
            const int qi = t_ / L_, l = t_ - qi * L_;
            const int r = qi >> 5, xi = qi & 31;
            if (r >= job.nrows || xi >= sorted_row_cols(tabs, p, job, r)) continue;
            const int q = job.qstart + (job.y0 + r) * job.Wq + job.x0 + xi;
            const size_t qm = (static_cast<size_t>(job.b) * p.Lq + q) * kHeads + job.m;
            float4 rc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) rc[k] = s_dot4[qi * LP + 4 * l + k];
            float4 *ga4 = reinterpret_cast<float4 *>(p.grad_attn + qm * LP);
            float4 *gl4 = reinterpret_cast<float4 *>(p.grad_loc + qm * LP * 2);
            if constexpr (FUSED) {
                // softmax backward: grad_logit_i = a_i * (ga_i - sum_j a_j ga_j); d loc / d offset = 1 / (W, H) cancels
                float dot = 0.f;
#pragma unroll
                for (int k = 0; k < LP; ++k) {
                    const float4 o = s_dot4[qi * LP + k];
                    dot = fmaf(o.w, o.x, dot);
                }
                ga4[l] = make_float4(rc[0].w * (rc[0].x - dot), rc[1].w * (rc[1].x - dot), rc[2].w * (rc[2].x - dot),
                                     rc[3].w * (rc[3].x - dot));
                gl4[2 * l] = make_float4(rc[0].y, rc[0].z, rc[1].y, rc[1].z);
                gl4[2 * l + 1] = make_float4(rc[2].y, rc[2].z, rc[3].y, rc[3].z);
            } else {
                const float wl = static_cast<float>(tabs.W[l]), hl = static_cast<float>(tabs.H[l]);
                ga4[l] = make_float4(rc[0].x, rc[1].x, rc[2].x, rc[3].x);
                gl4[2 * l] = make_float4(wl * rc[0].y, hl * rc[0].z, wl * rc[1].y, hl * rc[1].z);
                gl4[2 * l + 1] = make_float4(wl * rc[2].y, hl * rc[2].z, wl * rc[3].y, hl * rc[3].z);
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

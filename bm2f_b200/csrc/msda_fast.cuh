// D = 32 fast path: persistent, strip-walking forward / backward sampling kernels for sm_100a.
//
// Work decomposition:
//   job   = (image b, head m, query-level k, strip c, row chunk r): a column strip of SW
//           consecutive queries walked down `rows` image rows.  One head per job keeps the L1
//           working set to the bilinear footprints of ONE head around a slowly moving window,
//           so corner lines fetched for one row are re-used by the next rows from L1.
//   stage = one strip row: SW consecutive queries of head m.  Their sampling locations
//           (LP x 2 floats) and attention weights (LP floats) are a strided 2-D box of the
//           (N*Lq, M*LP*{2,1}) matrices: ONE TMA tensor-map copy each, landed in a shared-memory
//           ring by a producer warp and consumed through full/empty mbarriers.
//   warp  = one (query, head) at a time.  The 32 lanes are LG groups of LPC lanes; a group owns
//           one sampling point, its LPC lanes span the 32 head channels with VEC-wide loads
//           (VEC = 4: 8 lanes x 128 bit = one 128-byte line per corner, LG = 4 points in flight).
//
// When the queries are not the pixels of the levels (Lq != sum H*W) the same kernels walk the
// queries in plain 1-D order.
#pragma once
#include <type_traits>

#include "msda_common.cuh"

namespace bm2f {

constexpr int kHeads = 8;   // the fast path is compiled for M = 8 (d_model 256 = 8 x 32)

struct FastParams {
    const void *value;           // (N,S,M,32)
    const int64_t *shapes;       // (L,2) device
    const int64_t *start;        // (L)   device
    const float *loc;            // (N,Lq,M,L,P,2)        fused mode: raw sampling offsets in pixels
    const float *attn;           // (N,Lq,M,L,P)          fused mode: raw logits
    const float *ref;            // (N,Lq,L,2) reference points, fused mode only (loc = ref + offset / (W,H));
                                 // nullptr: pixel centres of the query's own level, computed in the kernel (Lq == S)
    const void *grad_out;        // (N,Lq,M,32)            bwd only
    void *out;                   // (N,Lq,M,32)            fwd only
    void *grad_value;            // (N,S,M,32) float32      bwd only
    float *grad_loc;             // like loc               bwd only
    float *grad_attn;            // like attn              bwd only
    int N, S, M, Lq;
    int rows;                    // strip rows per job
    int order;                   // 1 = force 1-D query order
    // packed fused layout (bm2f_msda_fused_*_packed): offsets and logits (and their gradients) are column blocks of ONE
    // (N * Lq, ld_packed) matrix — the output of a single offsets||logits projection.  0 = separate contiguous tensors.
    // The kernels read both through TMA tensor maps (row stride = ld_packed); only the gradient stores index memory.
    int ld_packed;
};

struct QLevel {
    int qstart, qend, Hq, Wq, nstrips, nchunks, job_base;   // queries [qstart, qend) walked as an Hq x Wq grid
};

template <int L_>
struct Tabs {
    int H[L_], W[L_], start[L_];
    QLevel ql[L_];
    int nql;
    int jobs_per_bm;
};

// ring depth: 8 stages unless that would exceed the 48 KB static shared-memory budget
constexpr int ring_stages(int sw, int lp) { return sw * lp * 12 * 8 <= 40 * 1024 ? 8 : 4; }

template <int L_, int SW>
__device__ __forceinline__ void build_tabs(Tabs<L_> &t, const FastParams &p)
{
    // one thread; L_ <= 16 so this is a few dozen instructions
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
        // queries are the pixels of the levels (encoder self-attention): walk level geometry
        int q0 = 0;
        for (int l = 0; l < L_; ++l) {
            QLevel &q = t.ql[l];
            const int npix = t.H[l] * t.W[l];
            q.qstart = q0;
            q.qend = q0 + npix;
            // column strips only pay off when they tile the level width without much waste; otherwise this
            // level is walked in raster order, SW consecutive pixels per stage (e.g. W = 20 or 40 with SW = 32)
            const int padded = (t.W[l] + SW - 1) / SW * SW;
            if ((padded - t.W[l]) * 8 <= t.W[l]) {
                q.Hq = t.H[l];
                q.Wq = t.W[l];
            } else {
                q.Wq = SW;
                q.Hq = (npix + SW - 1) / SW;
            }
            q.nstrips = (q.Wq + SW - 1) / SW;
            q.nchunks = (q.Hq + p.rows - 1) / p.rows;
            q.job_base = jobs;
            jobs += q.nstrips * q.nchunks;
            q0 += npix;
        }
        t.nql = L_;
    } else {
        QLevel &q = t.ql[0];
        q.qstart = 0;
        q.qend = p.Lq;
        q.Wq = SW;
        q.Hq = (p.Lq + SW - 1) / SW;
        q.nstrips = 1;
        q.nchunks = (q.Hq + p.rows - 1) / p.rows;
        q.job_base = 0;
        jobs = q.nchunks;
        t.nql = 1;
    }
    t.jobs_per_bm = jobs;
}

// Enumerates the stages of this CTA in a fixed order; producer and consumers call it with the
// same arguments so they agree on the stage sequence without exchanging descriptors.
template <int L_, int SW, typename F>
__device__ __forceinline__ void for_each_stage(const Tabs<L_> &t, const FastParams &p, F &&f)
{
    const int per_b = kHeads * t.jobs_per_bm;
    const int total_jobs = p.N * per_b;
    int s = 0;
    for (int j = blockIdx.x; j < total_jobs; j += gridDim.x) {
        const int b = j / per_b;
        const int r = j - b * per_b;
        const int jj = r / kHeads;
        const int m = r - jj * kHeads;
        int k = 0;
        while (k + 1 < t.nql && jj >= t.ql[k + 1].job_base) ++k;
        const QLevel ql = t.ql[k];
        const int tt = jj - ql.job_base;
        const int chunk = tt / ql.nstrips;
        const int strip = tt - chunk * ql.nstrips;
        const int y_end = min(ql.Hq, (chunk + 1) * p.rows);
        const int x0 = strip * SW;
        const int nq_row = min(SW, ql.Wq - x0);
        // geometry of the stage for the analytic reference points: level k, pixel row y, first column x0 when this
        // level is walked as its true H x W grid; y = -1 when it is walked in raster order (or queries are no pyramid)
        const bool grid = t.nql == L_ && ql.Wq == t.W[k] && ql.qstart == t.start[k];
        for (int y = chunk * p.rows; y < y_end; ++y, ++s) {
            const int q_base = ql.qstart + y * ql.Wq + x0;
            f(s, b, m, q_base, min(nq_row, ql.qend - q_base), k, grid ? y : -1, x0);
        }
    }
}

template <int SW, int LP>
struct StageLayout {
    static constexpr int kLocBytes = ((SW * LP * 8 + 127) / 128) * 128;
    static constexpr int kWBytes = ((SW * LP * 4 + 127) / 128) * 128;
    static constexpr int kBytes = kLocBytes + kWBytes;
    static constexpr int kTxBytes = SW * LP * 12;  // what the two TMA boxes deliver
};

// Shared-memory state of the loc/weight ring + its producer.
template <int L_, int P_, int SW, int WPG>
struct Ring {
    using Lay = StageLayout<SW, L_ * P_>;
    static constexpr int kStages = ring_stages(SW, L_ * P_);
    alignas(128) unsigned char data[kStages * Lay::kBytes];
    alignas(8) uint64_t full[kStages];
    alignas(8) uint64_t empty[kStages];
    alignas(8) uint64_t ready[kStages];     // fused mode: prologue warp -> consumers

    __device__ __forceinline__ void init()
    {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], WPG);
            mbar_init(&ready[i], 1);
        }
        fence_mbar_init();
    }
    __device__ __forceinline__ const float2 *loc(int slot, int qi) const
    {
        return reinterpret_cast<const float2 *>(data + slot * Lay::kBytes) + qi * (L_ * P_);
    }
    __device__ __forceinline__ const float *w(int slot, int qi) const
    {
        return reinterpret_cast<const float *>(data + slot * Lay::kBytes + Lay::kLocBytes) + qi * (L_ * P_);
    }
    __device__ __forceinline__ float2 *loc_rw(int slot, int qi)
    {
        return reinterpret_cast<float2 *>(data + slot * Lay::kBytes) + qi * (L_ * P_);
    }
    __device__ __forceinline__ float *w_rw(int slot, int qi)
    {
        return reinterpret_cast<float *>(data + slot * Lay::kBytes + Lay::kLocBytes) + qi * (L_ * P_);
    }
    // Producer (one lane): keep the ring full with the CTA's stage sequence.
    __device__ __forceinline__ void produce(const Tabs<L_> &t, const FastParams &p, const CUtensorMap *tm_loc,
                                            const CUtensorMap *tm_w)
    {
        constexpr int LP = L_ * P_;
        for_each_stage<L_, SW>(t, p, [&](int s, int b, int m, int q_base, int, int, int, int) {
            const int slot = s % kStages;
            mbar_wait(&empty[slot], ((s / kStages) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[slot], Lay::kTxBytes);
            unsigned char *dst = data + slot * Lay::kBytes;
            const int row = b * p.Lq + q_base;
            tma_load_2d(dst, tm_loc, m * LP * 2, row, &full[slot]);
            tma_load_2d(dst + Lay::kLocBytes, tm_w, m * LP, row, &full[slot]);
        });
    }
};


// Fused prologue of MSDeformAttn.forward (reference: ops/modules/ms_deform_attn.py:101-109), done in
// registers per (query, head): attention weights = softmax over the L*P logits, sampling locations =
// reference point + offset / (W_l, H_l).  Each lane holds NIT points (one per level iteration); the LG
// lane groups hold the other points, so max / sum are folded across groups with xor-shuffles.
// Reference points (fused mode).  ref_q != nullptr: the caller's (N, Lq, L, 2) tensor — a dependent global load per
// (query, head) in front of every gather address.  ref_q == nullptr: the encoder's own reference points with valid
// ratios 1 (msdeformattn.py:141-153): every level gets the query pixel's centre ((x + 0.5) / W_q, (y + 0.5) / H_q),
// computed here from the query index with the same fp32 division torch does (linspace values i + 0.5 are exact), so
// the result is bit-identical and no load sits on the critical path.
template <int L_>
__device__ __forceinline__ float2 pixel_centre_ref(int q, const int (&st)[L_], const int (&W)[L_], const float (&Wf)[L_],
                                                   const float (&Hf)[L_])
{
    int s0 = st[0], w = W[0];
    float wf = Wf[0], hf = Hf[0];
#pragma unroll
    for (int k = 1; k < L_; ++k)
        if (q >= st[k]) { s0 = st[k]; w = W[k]; wf = Wf[k]; hf = Hf[k]; }   // level starts ascend
    const int idx = q - s0;
    const int y = idx / w;
    const int x = idx - y * w;
    return make_float2((static_cast<float>(x) + 0.5f) / wf, (static_cast<float>(y) + 0.5f) / hf);
}

// Stage-level part of the analytic reference point: row term and the level's width, selected once per stage.
template <int L_>
struct StageRef {
    float wf, ry;      // level width; (y + 0.5) / H   (valid when row >= 0)
    int row;
    __device__ __forceinline__ StageRef(int kq, int yq, const float (&Wf)[L_], const float (&Hf)[L_])
    {
        float hf = Hf[0];
        wf = Wf[0];
#pragma unroll
        for (int k = 1; k < L_; ++k)
            if (kq == k) { wf = Wf[k]; hf = Hf[k]; }
        row = yq;
        ry = (static_cast<float>(yq) + 0.5f) / hf;
    }
    // reference point of query q = q_base + qi, column x0q + qi of the stage's row
    __device__ __forceinline__ float2 centre(int q, int x, const int (&st)[L_], const int (&W)[L_], const float (&Wf)[L_],
                                             const float (&Hf)[L_]) const
    {
        if (row >= 0) return make_float2((static_cast<float>(x) + 0.5f) / wf, ry);
        return pixel_centre_ref<L_>(q, st, W, Wf, Hf);
    }
};

// Locations: loc = ref + offset / (W_l, H_l).  No cross-lane dependency, so the gather addresses can be formed at once.
template <int NIT, int LG, int L_, int P_>
__device__ __forceinline__ void fused_locations(float (&xs)[NIT], float (&ys)[NIT], const float2 *ref_q, const float2 centre,
                                                const float (&rW)[L_], const float (&rH)[L_])
{
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int l = (it * LG) / P_;
        const float2 r = ref_q ? __ldg(ref_q + l) : centre;
        xs[it] = fmaf(xs[it], rW[l], r.x);
        ys[it] = fmaf(ys[it], rH[l], r.y);
    }
}

// Softmax over the L*P logits of a (query, head): each lane holds NIT of them, the LG lane groups the others.
template <int NIT, int LPC>
__device__ __forceinline__ void fused_softmax(float (&ws)[NIT])
{
    float mx = ws[0];
#pragma unroll
    for (int it = 1; it < NIT; ++it) mx = fmaxf(mx, ws[it]);
#pragma unroll
    for (int o = LPC; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        ws[it] = __expf(ws[it] - mx);
        sum += ws[it];
    }
#pragma unroll
    for (int o = LPC; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = __frcp_rn(sum);
#pragma unroll
    for (int it = 0; it < NIT; ++it) ws[it] *= inv;
}

template <int NIT, int LPC, int LG, int L_, int P_>
__device__ __forceinline__ void fused_prologue(float (&xs)[NIT], float (&ys)[NIT], float (&ws)[NIT],
                                               const float2 *ref_q, const float2 centre, const float (&rW)[L_],
                                               const float (&rH)[L_])
{
    fused_softmax<NIT, LPC>(ws);
    fused_locations<NIT, LG, L_, P_>(xs, ys, ref_q, centre, rW, rH);
}

// Prologue warp (fused mode with TMA staging).  The consumer lanes are organised as 4 point groups x 8 channel lanes, so
// a prologue done by the consumers is computed 8 times over (every lane of a group repeats its point's softmax and
// location arithmetic: +36 % warp instructions, +50 % forward time, measured).  Here ONE extra warp converts a whole
// stage in shared memory, lane = query: softmax over the query's L*P logits in registers, loc = ref + offset / (W, H),
// written back in place; the consumers then run exactly the unfused code on the staged data.  ~150 warp instructions
// per stage instead of ~2000, off the consumers' critical path (the ring keeps the prologue several stages ahead).
template <int L_, int P_, int SW, typename RingT>
__device__ __forceinline__ void prologue_role(RingT &ring, const Tabs<L_> &tabs, const FastParams &p, int lane)
{
    constexpr int LP = L_ * P_, kStages = RingT::kStages;
    static_assert(P_ == 4, "one float4 of logits / two float4 of offsets per level");
    int W[L_], st[L_];
    float Hf[L_], Wf[L_], rW[L_], rH[L_];
#pragma unroll
    for (int l = 0; l < L_; ++l) {
        W[l] = tabs.W[l]; st[l] = tabs.start[l];
        Hf[l] = static_cast<float>(tabs.H[l]); Wf[l] = static_cast<float>(W[l]);
        rW[l] = 1.f / Wf[l]; rH[l] = 1.f / Hf[l];
    }
    for_each_stage<L_, SW>(tabs, p, [&](int s, int b, int, int q_base, int nq, int kq, int yq, int x0q) {
        const int slot = s % kStages;
        mbar_wait(&ring.full[slot], (s / kStages) & 1);
        const StageRef<L_> sref(kq, yq, Wf, Hf);
        for (int qi = lane; qi < nq; qi += 32) {
            const int q = q_base + qi;
            float4 *sw = reinterpret_cast<float4 *>(ring.w_rw(slot, qi));
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
            float4 *sl = reinterpret_cast<float4 *>(ring.loc_rw(slot, qi));      // (x, y) pairs: two points per float4
            const float2 *ref_q = p.ref ? reinterpret_cast<const float2 *>(p.ref) + (static_cast<size_t>(b) * p.Lq + q) * L_
                                        : nullptr;
            const float2 centre = p.ref ? make_float2(0.f, 0.f) : sref.centre(q, x0q + qi, st, W, Wf, Hf);
#pragma unroll
            for (int l = 0; l < L_; ++l) {
                const float2 r = ref_q ? __ldg(ref_q + l) : centre;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 o = sl[2 * l + h];
                    o.x = fmaf(o.x, rW[l], r.x); o.y = fmaf(o.y, rH[l], r.y);
                    o.z = fmaf(o.z, rW[l], r.x); o.w = fmaf(o.w, rH[l], r.y);
                    sl[2 * l + h] = o;
                }
            }
        }
        fence_proxy_async_smem();      // the slot is later overwritten by TMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.ready[slot]);
    });
}

// ------------------------------------------------------------------------------------------
// Forward
// ------------------------------------------------------------------------------------------
template <typename T, int VEC, int L_, int P_, int SW, int NWARP, int G, bool TMA, int CPS, bool FUSED = false>
__global__ void __launch_bounds__((NWARP + (TMA ? 1 : 0) + (TMA && FUSED ? 1 : 0)) * 32, CPS)
msda_fwd_fast_kernel(const FastParams p, const __grid_constant__ CUtensorMap tm_loc,
                     const __grid_constant__ CUtensorMap tm_w)
{
    constexpr int D = 32, LP = L_ * P_, LPC = D / VEC, LG = 32 / LPC, NIT = LP / LG, WPG = NWARP / G;
    static_assert(P_ % LG == 0, "a lane-group iteration must stay inside one level");
    static_assert(NWARP % G == 0, "");

    __shared__ Tabs<L_> tabs;
    using RingT = Ring<L_, P_, TMA ? SW : 1, WPG>;
    constexpr int kStages = RingT::kStages;
    __shared__ RingT ring;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        build_tabs<L_, SW>(tabs, p);
        if (TMA) {
            ring.init();
            tma_prefetch_desc(&tm_loc);
            tma_prefetch_desc(&tm_w);
        }
    }
    __syncthreads();

    if (TMA && warp == NWARP) {
        if (lane == 0) ring.produce(tabs, p, &tm_loc, &tm_w);
        return;
    }
    constexpr bool PRO = TMA && FUSED;      // prologue done by a dedicated warp on the staged data
    if (PRO && warp == NWARP + 1) {
        prologue_role<L_, P_, SW>(ring, tabs, p, lane);
        return;
    }

    int H[L_], W[L_], st[L_];
    float Hf[L_], Wf[L_];
#pragma unroll
    for (int l = 0; l < L_; ++l) {
        H[l] = tabs.H[l]; W[l] = tabs.W[l]; st[l] = tabs.start[l];
        Hf[l] = static_cast<float>(H[l]); Wf[l] = static_cast<float>(W[l]);
    }
    float rW[L_], rH[L_];
#pragma unroll
    for (int l = 0; l < L_; ++l) { rW[l] = 1.f / Wf[l]; rH[l] = 1.f / Hf[l]; }
    (void)rW; (void)rH;
    const int g = warp / WPG, wi = warp % WPG;
    const int lg = lane / LPC, sub = lane % LPC;
    constexpr int MD = kHeads * D;  // elements between horizontally adjacent pixels (immediate offset)
    const T *value = static_cast<const T *>(p.value);
    T *out = static_cast<T *>(p.out);

    for_each_stage<L_, SW>(tabs, p, [&](int s, int b, int m, int q_base, int nq, int kq, int yq, int x0q) {
        if (G > 1 && (s % G) != g) return;
        const int slot = s % kStages;
        if (TMA) mbar_wait(PRO ? &ring.ready[slot] : &ring.full[slot], (s / kStages) & 1);
        const T *vlane = value + (static_cast<size_t>(b) * p.S * kHeads + m) * D + sub * VEC;
        const StageRef<L_> sref(kq, yq, Wf, Hf);
        (void)sref;
        for (int qi = wi; qi < nq; qi += WPG) {
            const int q = q_base + qi;
            float xs[NIT], ys[NIT], ws[NIT];
            if (TMA) {
                const float2 *sl = ring.loc(slot, qi);
                const float *sw = ring.w(slot, qi);
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const float2 t = sl[it * LG + lg];
                    xs[it] = t.x; ys[it] = t.y; ws[it] = sw[it * LG + lg];
                }
            } else {
                const size_t k0 = ((static_cast<size_t>(b) * p.Lq + q) * kHeads + m) * LP;
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const float2 t = __ldg(reinterpret_cast<const float2 *>(p.loc) + k0 + it * LG + lg);
                    xs[it] = t.x; ys[it] = t.y; ws[it] = __ldg(p.attn + k0 + it * LG + lg);
                }
            }
            if constexpr (FUSED && !PRO)
                fused_locations<NIT, LG, L_, P_>(
                    xs, ys,
                    p.ref ? reinterpret_cast<const float2 *>(p.ref) + (static_cast<size_t>(b) * p.Lq + q) * L_ : nullptr,
                    p.ref ? make_float2(0.f, 0.f) : sref.centre(q, x0q + qi, st, W, Wf, Hf), rW, rH);
            float acc[VEC];
#pragma unroll
            for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int l = (it * LG) / P_;
                const Footprint f = make_footprint(xs[it], ys[it], H[l], W[l], Hf[l], Wf[l]);
                const T *row0 = vlane + static_cast<long long>(st[l] + f.y0 * W[l] + f.x0) * MD;
                const T *row1 = row0 + W[l] * MD;
                const T *corner[4] = {row0, row0 + MD, row1, row1 + MD};
                const float wx[2] = {f.hw, f.lw};
                float v[4][VEC];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int c = 0; c < VEC; ++c) v[k][c] = 0.f;
                    if (f.ok[k]) VecIO<T, VEC>::load(corner[k], v[k]);
                }
                // fused mode: the softmax (two shuffle rounds, exps, a reciprocal) runs while the first gathers are in
                // flight instead of in front of every address computation
                if constexpr (FUSED && !PRO)
                    if (it == 0) fused_softmax<NIT, LPC>(ws);
                const float wy[2] = {ws[it] * f.hh, ws[it] * f.lh};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float cw = wy[k >> 1] * wx[k & 1];   // v == 0 for dropped corners
#pragma unroll
                    for (int c = 0; c < VEC; ++c) acc[c] = fmaf(cw, v[k][c], acc[c]);
                }
            }
            // fold the LG point groups
#pragma unroll
            for (int o = LPC; o < 32; o <<= 1)
#pragma unroll
                for (int c = 0; c < VEC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
            if (lg == 0)
                VecIO<T, VEC>::store(out + ((static_cast<size_t>(b) * p.Lq + q) * kHeads + m) * D + sub * VEC, acc);
        }
        if (TMA) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[slot]);
        }
    });
}


// ------------------------------------------------------------------------------------------
// Forward with geometry warps.  In msda_fwd_fast_kernel the 8 channel lanes of a point group all repeat the point's
// scalar geometry — pixel coordinates, floor, bilinear weights, range tests, corner addresses: ~45 of the ~65
// instructions of an iteration — before they can issue their gathers.  Here NGEO extra warps do that once per point
// (lane = (query, point) of the TMA-staged stage) and leave a 32-byte record per point in a second shared-memory ring:
// four element offsets (-1 = corner dropped) and the four combined weights attn * bilinear.  A consumer iteration is
// then two 128-bit shared loads, four address adds, four gathers and sixteen FMAs, in the same order and with the same
// operands as the fast kernel, so the output is bit-identical.
//   TMA warp -> raw ring (loc, attn) -> geometry warps -> record ring -> 16 consumer warps
// ------------------------------------------------------------------------------------------
template <int L_, int P_, int SW, int RSTAGES>
struct GeoRing {
    static constexpr int kPoints = SW * L_ * P_;
    static constexpr int kStageBytes = kPoints * 32;
    static constexpr int kBytes = RSTAGES * kStageBytes;
    __device__ static __forceinline__ int4 *offs(unsigned char *base, int rslot)
    {
        return reinterpret_cast<int4 *>(base + rslot * kStageBytes);
    }
    __device__ static __forceinline__ float4 *cws(unsigned char *base, int rslot)
    {
        return reinterpret_cast<float4 *>(base + rslot * kStageBytes + kPoints * 16);
    }
};

// WIDE (fp32): the consumers gather with 256-bit loads — 4 lanes x 32 B per corner, 8 corners (two points) per warp
// instruction, the access shape with the 1.6x higher measured L1 line rate (profiles/r01_microbench.txt).  The partial
// sums are partitioned differently across lanes, so the result equals the default kernel to rounding, not bitwise.
// LEAN: a dropped corner (outside the level, or the point out of range) is recorded with weight 0 and the offset of an
// in-image pixel instead of -1, so the consumers issue their four gathers unconditionally: no predicates, no zero
// initialisation of the destination registers — the loop is two shared loads, four address computations, four gathers,
// sixteen FMAs.  0 * v adds nothing, so the result equals the predicated kernels bit for bit as long as `value` is finite.
// LEAN = 2: 16-byte records {anchor offset | x-drop flags, a * wy0, a * wy1, lw}: ONE 128-bit shared load per point
// instead of two (an LDS.128 costs four L1 data-pipe wavefronts even when the 8 lanes of a group read the same 16 bytes,
// and that pipe is what bounds this kernel), the other three corner addresses follow from the anchor.  The anchor is
// clamped into [0, H - 2] x [0, W - 2] so that all four addresses lie inside the level; a border point's surviving row /
// column moves to the other slot (weights exact in y; in x the surviving weight is stored as hw and read back as
// 1 - hw, which differs from lw by <= 2^-25 when lw < 0.5).  Single-row / single-column levels use a zero stride.
template <typename T, int L_, int P_, int SW, int NWARP, int NGEO, int RSTAGES, bool FUSED = false, bool WIDE = false,
          int CPS = 1, int LEAN = 0>
__global__ void __launch_bounds__((NWARP + 1 + NGEO) * 32, CPS)
msda_fwd_geo_kernel(const FastParams p, const __grid_constant__ CUtensorMap tm_loc, const __grid_constant__ CUtensorMap tm_w)
{
    constexpr int D = 32, VEC = 4, LP = L_ * P_, LPC = D / VEC, LG = 32 / LPC, NIT = LP / LG;
    static_assert(P_ % LG == 0, "a lane-group iteration must stay inside one level");
    __shared__ Tabs<L_> tabs;
    using RingT = Ring<L_, P_, SW, NGEO>;          // raw ring: released by the geometry warps
    using Geo = GeoRing<L_, P_, SW, RSTAGES>;
    constexpr int kStages = RingT::kStages;
    __shared__ RingT ring;
    __shared__ alignas(8) uint64_t rec_full[RSTAGES], rec_empty[RSTAGES];
    extern __shared__ __align__(128) unsigned char geo_smem[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        build_tabs<L_, SW>(tabs, p);
        ring.init();
        for (int i = 0; i < RSTAGES; ++i) {
            mbar_init(&rec_full[i], NGEO);
            mbar_init(&rec_empty[i], NWARP);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tm_loc);
        tma_prefetch_desc(&tm_w);
    }
    __syncthreads();

    if (warp == NWARP) {                                   // ---- TMA producer ----
        if (lane == 0) ring.produce(tabs, p, &tm_loc, &tm_w);
        return;
    }
    int H[L_], W[L_], st[L_];
    float Hf[L_], Wf[L_];
#pragma unroll
    for (int l = 0; l < L_; ++l) {
        H[l] = tabs.H[l]; W[l] = tabs.W[l]; st[l] = tabs.start[l];
        Hf[l] = static_cast<float>(H[l]); Wf[l] = static_cast<float>(W[l]);
    }
    constexpr int MD = kHeads * D;
    // LEAN = 2: a level with a single row or column (never in Mask2Former's pyramids) has no 2 x 2 block to clamp the
    // anchor into; the whole launch then uses the 32-byte lean records (uniform decision from the device-resident table)
    bool degenerate = false;
#pragma unroll
    for (int l = 0; l < L_; ++l) degenerate = degenerate || H[l] < 2 || W[l] < 2;
    (void)degenerate;

    if (warp > NWARP) {                                    // ---- geometry warps ----
        const int gw = warp - NWARP - 1;
        float rW[L_], rH[L_];
#pragma unroll
        for (int l = 0; l < L_; ++l) { rW[l] = 1.f / Wf[l]; rH[l] = 1.f / Hf[l]; }
        (void)rW; (void)rH;
        // GMODE = record format of this launch (see the consumers): separate loop bodies keep the two-point unrolling tight
        auto produce = [&](auto mode_c) {
            constexpr int GMODE = decltype(mode_c)::value;
            for_each_stage<L_, SW>(tabs, p, [&](int s, int b, int, int q_base, int nq, int kq, int yq, int x0q) {
                const int slot = s % kStages, rslot = s % RSTAGES;
                mbar_wait(&ring.full[slot], (s / kStages) & 1);
                if constexpr (FUSED) {
                    // softmax over each query's logits, lane = query (this warp's share of the stage's queries)
                    for (int qi = gw * 32 + lane; qi < nq; qi += NGEO * 32) {
                        float4 *sw = reinterpret_cast<float4 *>(ring.w_rw(slot, qi));
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
                    }
                    // every geometry warp normalises the queries it reads below only if the split matches; with NGEO > 1 the
                    // point loop crosses queries of the other warp, so all geometry warps meet here first
                    if (NGEO > 1) asm volatile("bar.sync 2, %0;" ::"n"(NGEO * 32) : "memory");
                    else __syncwarp();
                }
                mbar_wait(&rec_empty[rslot], ((s / RSTAGES) & 1) ^ 1);
                int4 *ro = Geo::offs(geo_smem, rslot);
                float4 *rc = Geo::cws(geo_smem, rslot);
                const StageRef<L_> sref(kq, yq, Wf, Hf);
                (void)sref;
                const int npts = nq * LP;
                // a geometry warp is one dependent chain per point (~100 instructions): two points per trip give it the
                // instruction-level parallelism to keep up with the consumers (measured: the record ring, not L1, paced the
                // geometry-warp variants)
    #pragma unroll 2
                for (int idx = gw * 32 + lane; idx < npts; idx += NGEO * 32) {
                    const int qi = idx / LP, j = idx - qi * LP;
                    const float2 xy = ring.loc(slot, qi)[j];
                    const float a = ring.w(slot, qi)[j];
                    int l = 0;
    #pragma unroll
                    for (int k = 1; k < L_; ++k) l = (j >= k * P_) ? k : l;
                    int Hl = H[0], Wl = W[0], stl = st[0];
                    float Hfl = Hf[0], Wfl = Wf[0], rWl = rW[0], rHl = rH[0];
    #pragma unroll
                    for (int k = 1; k < L_; ++k)
                        if (l == k) { Hl = H[k]; Wl = W[k]; stl = st[k]; Hfl = Hf[k]; Wfl = Wf[k]; rWl = rW[k]; rHl = rH[k]; }
                    (void)rWl; (void)rHl;
                    float x = xy.x, y = xy.y;
                    if constexpr (FUSED) {
                        const int q = q_base + qi;
                        float2 r;
                        if (p.ref) r = __ldg(reinterpret_cast<const float2 *>(p.ref) + (static_cast<size_t>(b) * p.Lq + q) * L_ + l);
                        else r = sref.centre(q, x0q + qi, st, W, Wf, Hf);
                        x = fmaf(x, rWl, r.x);
                        y = fmaf(y, rHl, r.y);
                    }
                    const Footprint f = make_footprint(x, y, Hl, Wl, Hfl, Wfl);
                    const int e00 = (stl + f.y0 * Wl + f.x0) * MD;
                    const float wy0 = a * f.hh, wy1 = a * f.lh;
                    if constexpr (GMODE == 2) {
                        // separable weights: rows (a * hh, a * lh) masked by the row tests, columns (hw, lw) by the column tests
                        const bool y0ok = f.in_range && f.y0 >= 0, y1ok = f.in_range && f.y0 + 1 <= Hl - 1;
                        float r0 = y0ok ? wy0 : 0.f, r1 = y1ok ? wy1 : 0.f;
                        int ay = f.y0, ax = f.x0;
                        if (!f.in_range) { ay = 0; ax = 0; }
                        if (ay < 0) { ay = 0; r0 = r1; r1 = 0.f; }                        // row y = 0 is the footprint's lower row
                        else if (ay > Hl - 2) { ay = Hl - 2; r1 = r0; r0 = 0.f; }         // row y = H - 1 is its upper row
                        float lwf = f.in_range ? f.lw : 0.f;      // a non-finite location has a NaN fraction: keep it out of 0 * w
                        int flags = 0;
                        if (ax < 0) { ax = 0; lwf = f.hw; flags = 2; }                    // column 0 carries lw: stored as hw, read as 1 - hw
                        else if (ax > Wl - 2) { ax = Wl - 2; lwf = f.hw; flags = 1; }     // column W - 1 carries hw in slot 1
                        const int e = (stl + ay * Wl + ax) * MD;
                        rc[idx] = make_float4(__int_as_float(e | flags), r0, r1, lwf);
                    } else if constexpr (GMODE == 1) {
                        const int safe = (stl + max(min(max(f.y0, 0), Hl - 1), 0) * Wl + max(min(max(f.x0, 0), Wl - 1), 0)) * MD;   // (empty level: element 0)
                        ro[idx] = make_int4(f.ok[0] ? e00 : safe, f.ok[1] ? e00 + MD : safe, f.ok[2] ? e00 + Wl * MD : safe,
                                            f.ok[3] ? e00 + Wl * MD + MD : safe);
                        rc[idx] = make_float4(f.ok[0] ? wy0 * f.hw : 0.f, f.ok[1] ? wy0 * f.lw : 0.f, f.ok[2] ? wy1 * f.hw : 0.f,
                                              f.ok[3] ? wy1 * f.lw : 0.f);
                    } else {
                    ro[idx] = make_int4(f.ok[0] ? e00 : -1, f.ok[1] ? e00 + MD : -1, f.ok[2] ? e00 + Wl * MD : -1,
                                        f.ok[3] ? e00 + Wl * MD + MD : -1);
                    rc[idx] = make_float4(wy0 * f.hw, wy0 * f.lw, wy1 * f.hw, wy1 * f.lw);
                    }
                }
                if constexpr (FUSED) fence_proxy_async_smem();   // the raw slot (softmax written in place) goes back to TMA
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&ring.empty[slot]);
                    mbar_arrive(&rec_full[rslot]);
                }
            });
        };
        if (LEAN == 2 && degenerate) produce(std::integral_constant<int, 1>{});
        else produce(std::integral_constant<int, LEAN>{});
        return;
    }

    // ---- consumers ----
    const int wi = warp;
    const int lg = lane / LPC, sub = lane % LPC;
    const T *value = static_cast<const T *>(p.value);
    T *out = static_cast<T *>(p.out);
    if constexpr (WIDE) {
        static_assert(std::is_same<T, float>::value && LP % 2 == 0, "256-bit gathers: fp32, points in pairs");
        const int s8 = lane >> 2, sub8 = lane & 3;     // corner slot of the instruction, 8-channel group
        const int kk = s8 & 3, pp = s8 >> 2;           // corner of the point, point of the pair
        for_each_stage<L_, SW>(tabs, p, [&](int s, int b, int m, int q_base, int nq, int, int, int) {
            const int rslot = s % RSTAGES;
            mbar_wait(&rec_full[rslot], (s / RSTAGES) & 1);
            const int *ro = reinterpret_cast<const int *>(Geo::offs(geo_smem, rslot));
            const float *rc = reinterpret_cast<const float *>(Geo::cws(geo_smem, rslot));
            const float *vlane = static_cast<const float *>(p.value) + (static_cast<size_t>(b) * p.S * kHeads + m) * D + sub8 * 8;
            for (int qi = wi; qi < nq; qi += NWARP) {
                const int q = q_base + qi;
                float acc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
                for (int it = 0; it < LP / 2; ++it) {
                    const int e = (qi * LP + it * 2 + pp) * 4 + kk;
                    const int off = ro[e];
                    const float cw = rc[e];
                    float v[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = 0.f;
                    ldg256_pred(vlane + off, off >= 0, v);
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c] = fmaf(cw, v[c], acc[c]);
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
                if (s8 == 0) {
                    float4 *dst = reinterpret_cast<float4 *>(static_cast<float *>(p.out) +
                                                             ((static_cast<size_t>(b) * p.Lq + q) * kHeads + m) * D + sub8 * 8);
                    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&rec_empty[rslot]);
        });
        return;
    }
    // MODE = record format of this launch (LEAN, or 1 when LEAN = 2 meets a degenerate level): separate loop bodies
    auto consume = [&](auto mode_c) {
        constexpr int MODE = decltype(mode_c)::value;
        for_each_stage<L_, SW>(tabs, p, [&](int s, int b, int m, int q_base, int nq, int, int, int) {
            const int rslot = s % RSTAGES;
            mbar_wait(&rec_full[rslot], (s / RSTAGES) & 1);
            const int4 *ro = Geo::offs(geo_smem, rslot);
            const float4 *rc = Geo::cws(geo_smem, rslot);
            const T *vlane = value + (static_cast<size_t>(b) * p.S * kHeads + m) * D + sub * VEC;
            // queries are dealt to the consumer warps round-robin ACROSS stages (flat index s * SW + qi), so that a warp count
            // that does not divide the stage (24, 28 warps for 32 queries) still keeps every warp equally busy
            int first = wi - (s * SW) % NWARP;
            if (first < 0) first += NWARP;
            for (int qi = first; qi < nq; qi += NWARP) {
                const int q = q_base + qi;
                float acc[VEC];
    #pragma unroll
                for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
    #pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    int off[4];
                    float cw[4];
                    if constexpr (MODE == 2) {
                        static_assert(MODE != 2 || LG == P_, "one iteration = the four points of one level");
                        const float4 r = rc[qi * LP + it * LG + lg];
                        const int o = __float_as_int(r.x);
                        const float hwf = 1.f - r.w;
                        const float wx0 = (o & 1) ? 0.f : hwf, wx1 = (o & 2) ? 0.f : r.w;
                        cw[0] = r.y * wx0; cw[1] = r.y * wx1; cw[2] = r.z * wx0; cw[3] = r.z * wx1;
                        off[0] = o & ~3;
                        off[1] = off[0] + MD;
                        off[2] = off[0] + W[it] * MD;
                        off[3] = off[2] + MD;
                    } else {
                        const int4 o = ro[qi * LP + it * LG + lg];
                        const float4 cw4 = rc[qi * LP + it * LG + lg];
                        off[0] = o.x; off[1] = o.y; off[2] = o.z; off[3] = o.w;
                        cw[0] = cw4.x; cw[1] = cw4.y; cw[2] = cw4.z; cw[3] = cw4.w;
                    }
                    float v[4][VEC];
    #pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if constexpr (MODE >= 1) {
                            VecIO<T, VEC>::load(vlane + off[k], v[k]);
                        } else {
    #pragma unroll
                            for (int c = 0; c < VEC; ++c) v[k][c] = 0.f;
                            if (off[k] >= 0) VecIO<T, VEC>::load(vlane + off[k], v[k]);
                        }
                    }
    #pragma unroll
                    for (int k = 0; k < 4; ++k)
    #pragma unroll
                        for (int c = 0; c < VEC; ++c) acc[c] = fmaf(cw[k], v[k][c], acc[c]);
                }
    #pragma unroll
                for (int o = LPC; o < 32; o <<= 1)
    #pragma unroll
                    for (int c = 0; c < VEC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
                if (lg == 0)
                    VecIO<T, VEC>::store(out + ((static_cast<size_t>(b) * p.Lq + q) * kHeads + m) * D + sub * VEC, acc);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&rec_empty[rslot]);
        });
    };
    if (LEAN == 2 && degenerate) consume(std::integral_constant<int, 1>{});
    else consume(std::integral_constant<int, LEAN>{});
}


// ------------------------------------------------------------------------------------------
// Forward, 256-bit variant (fp32 only): 4 lanes x 32 B per corner.  The 32 lanes are 8 groups =
// 4 points x 2 footprint rows; a lane loads the two horizontally adjacent corners of its row with
// two LDG.256, so a warp instruction moves 8 lines.  Measured L1 line rate of this access shape is
// 1.6x the 8 x 16 B shape (profiles/r01_microbench.txt).
// ------------------------------------------------------------------------------------------
template <int L_, int P_, int SW, int NWARP, int G, bool TMA, int CPS>
__global__ void __launch_bounds__((NWARP + (TMA ? 1 : 0)) * 32, CPS)
msda_fwd_fast256_kernel(const FastParams p, const __grid_constant__ CUtensorMap tm_loc,
                        const __grid_constant__ CUtensorMap tm_w)
{
    constexpr int D = 32, LP = L_ * P_, WPG = NWARP / G;
    static_assert(P_ == 4, "one lane-group quartet per level");
    using RingT = Ring<L_, P_, TMA ? SW : 1, WPG>;
    constexpr int kStages = RingT::kStages;
    __shared__ Tabs<L_> tabs;
    __shared__ RingT ring;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        build_tabs<L_, SW>(tabs, p);
        if (TMA) {
            ring.init();
            tma_prefetch_desc(&tm_loc);
            tma_prefetch_desc(&tm_w);
        }
    }
    __syncthreads();
    if (TMA && warp == NWARP) {
        if (lane == 0) ring.produce(tabs, p, &tm_loc, &tm_w);
        return;
    }
    int H[L_], W[L_], st[L_];
    float Hf[L_], Wf[L_];
#pragma unroll
    for (int l = 0; l < L_; ++l) {
        H[l] = tabs.H[l]; W[l] = tabs.W[l]; st[l] = tabs.start[l];
        Hf[l] = static_cast<float>(H[l]); Wf[l] = static_cast<float>(W[l]);
    }
    const int g = warp / WPG, wi = warp % WPG;
    const int grp = lane >> 2, sub = lane & 3;
    const int pt = grp & 3, dy = grp >> 2;
    constexpr int MD = kHeads * D;
    const float *value = static_cast<const float *>(p.value);
    float *out = static_cast<float *>(p.out);

    for_each_stage<L_, SW>(tabs, p, [&](int s, int b, int m, int q_base, int nq, int kq, int yq, int x0q) {
        if (G > 1 && (s % G) != g) return;
        const int slot = s % kStages;
        if (TMA) mbar_wait(&ring.full[slot], (s / kStages) & 1);
        const float *vlane = value + (static_cast<size_t>(b) * p.S * kHeads + m) * D + sub * 8;
        for (int qi = wi; qi < nq; qi += WPG) {
            const int q = q_base + qi;
            float xs[L_], ys[L_], ws[L_];
            if (TMA) {
                const float2 *sl = ring.loc(slot, qi);
                const float *sw = ring.w(slot, qi);
#pragma unroll
                for (int l = 0; l < L_; ++l) {
                    const float2 t = sl[l * 4 + pt];
                    xs[l] = t.x; ys[l] = t.y; ws[l] = sw[l * 4 + pt];
                }
            } else {
                const size_t k0 = ((static_cast<size_t>(b) * p.Lq + q) * kHeads + m) * LP;
#pragma unroll
                for (int l = 0; l < L_; ++l) {
                    const float2 t = __ldg(reinterpret_cast<const float2 *>(p.loc) + k0 + l * 4 + pt);
                    xs[l] = t.x; ys[l] = t.y; ws[l] = __ldg(p.attn + k0 + l * 4 + pt);
                }
            }
            float acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
            for (int l = 0; l < L_; ++l) {
                const Footprint f = make_footprint(xs[l], ys[l], H[l], W[l], Hf[l], Wf[l]);
                const bool ok0 = dy ? f.ok[2] : f.ok[0], ok1 = dy ? f.ok[3] : f.ok[1];
                const float wrow = ws[l] * (dy ? f.lh : f.hh);
                const float *c0 = vlane + static_cast<long long>(st[l] + (f.y0 + dy) * W[l] + f.x0) * MD;
                float v0[8], v1[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) { v0[c] = 0.f; v1[c] = 0.f; }
                ldg256_pred(c0, ok0, v0);
                ldg256_pred(c0 + MD, ok1, v1);
                const float w0 = wrow * f.hw, w1 = wrow * f.lw;
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = fmaf(w1, v1[c], fmaf(w0, v0[c], acc[c]));
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
            if (grp == 0) {
                float4 *o4 = reinterpret_cast<float4 *>(out + ((static_cast<size_t>(b) * p.Lq + q) * kHeads + m) * D + sub * 8);
                o4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                o4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
        }
        if (TMA) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[slot]);
        }
    });
}

// ------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------
template <typename T, int VEC, int L_, int P_, int SW, int NWARP, int G, bool TMA, int CPS, bool MERGE,
          bool FUSED = false>
__global__ void __launch_bounds__((NWARP + (TMA ? 1 : 0) + (TMA && FUSED ? 1 : 0)) * 32, CPS)
msda_bwd_fast_kernel(const FastParams p, const __grid_constant__ CUtensorMap tm_loc,
                     const __grid_constant__ CUtensorMap tm_w)
{
    constexpr int D = 32, LP = L_ * P_, LPC = D / VEC, LG = 32 / LPC, NIT = LP / LG, WPG = NWARP / G;
    static_assert(P_ % LG == 0, "a lane-group iteration must stay inside one level");

    __shared__ Tabs<L_> tabs;
    using RingT = Ring<L_, P_, TMA ? SW : 1, WPG>;
    constexpr int kStages = RingT::kStages;
    __shared__ RingT ring;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        build_tabs<L_, SW>(tabs, p);
        if (TMA) {
            ring.init();
            tma_prefetch_desc(&tm_loc);
            tma_prefetch_desc(&tm_w);
        }
    }
    __syncthreads();

    if (TMA && warp == NWARP) {
        if (lane == 0) ring.produce(tabs, p, &tm_loc, &tm_w);
        return;
    }
    constexpr bool PRO = TMA && FUSED;      // prologue done by a dedicated warp on the staged data
    if (PRO && warp == NWARP + 1) {
        prologue_role<L_, P_, SW>(ring, tabs, p, lane);
        return;
    }

    int H[L_], W[L_], st[L_];
    float Hf[L_], Wf[L_];
#pragma unroll
    for (int l = 0; l < L_; ++l) {
        H[l] = tabs.H[l]; W[l] = tabs.W[l]; st[l] = tabs.start[l];
        Hf[l] = static_cast<float>(H[l]); Wf[l] = static_cast<float>(W[l]);
    }
    float rW[L_], rH[L_];
#pragma unroll
    for (int l = 0; l < L_; ++l) { rW[l] = 1.f / Wf[l]; rH[l] = 1.f / Hf[l]; }
    (void)rW; (void)rH;
    const int g = warp / WPG, wi = warp % WPG;
    const int lg = lane / LPC, sub = lane % LPC;
    constexpr int MD = kHeads * D;  // elements between horizontally adjacent pixels (immediate offset)
    const T *value = static_cast<const T *>(p.value);
    const T *grad_out = static_cast<const T *>(p.grad_out);
    float *grad_value = static_cast<float *>(p.grad_value);   // fp32 accumulation for every value dtype

    for_each_stage<L_, SW>(tabs, p, [&](int s, int b, int m, int q_base, int nq, int kq, int yq, int x0q) {
        if (G > 1 && (s % G) != g) return;
        const int slot = s % kStages;
        if (TMA) mbar_wait(PRO ? &ring.ready[slot] : &ring.full[slot], (s / kStages) & 1);
        const size_t img = (static_cast<size_t>(b) * p.S * kHeads + m) * D + sub * VEC;
        const T *vlane = value + img;
        float *gvlane = grad_value + img;
        const StageRef<L_> sref(kq, yq, Wf, Hf);
        (void)sref;
        for (int qi = wi; qi < nq; qi += WPG) {
            const int q = q_base + qi;
            const size_t qm = (static_cast<size_t>(b) * p.Lq + q) * kHeads + m;
            float go[VEC];
            VecIO<T, VEC>::load(grad_out + qm * D + sub * VEC, go);
            float xs[NIT], ys[NIT], ws[NIT];
            if (TMA) {
                const float2 *sl = ring.loc(slot, qi);
                const float *sw = ring.w(slot, qi);
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const float2 t = sl[it * LG + lg];
                    xs[it] = t.x; ys[it] = t.y; ws[it] = sw[it * LG + lg];
                }
            } else {
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const float2 t = __ldg(reinterpret_cast<const float2 *>(p.loc) + qm * LP + it * LG + lg);
                    xs[it] = t.x; ys[it] = t.y; ws[it] = __ldg(p.attn + qm * LP + it * LG + lg);
                }
            }
            if constexpr (FUSED && !PRO)
                fused_locations<NIT, LG, L_, P_>(
                    xs, ys,
                    p.ref ? reinterpret_cast<const float2 *>(p.ref) + (static_cast<size_t>(b) * p.Lq + q) * L_ : nullptr,
                    p.ref ? make_float2(0.f, 0.f) : sref.centre(q, x0q + qi, st, W, Wf, Hf), rW, rH);
            float ga_keep[NIT], gx_keep[NIT], gy_keep[NIT];   // fused mode: written after the softmax backward
            (void)ga_keep; (void)gx_keep; (void)gy_keep;
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int l = (it * LG) / P_;
                const Footprint f = make_footprint(xs[it], ys[it], H[l], W[l], Hf[l], Wf[l]);
                const long long e00 = static_cast<long long>(st[l] + f.y0 * W[l] + f.x0) * MD;
                const long long eoff[4] = {e00, e00 + MD, e00 + W[l] * MD, e00 + W[l] * MD + MD};
                const float wy[2] = {f.hh, f.lh};
                const float wx[2] = {f.hw, f.lw};
                float v[4][VEC];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int c = 0; c < VEC; ++c) v[k][c] = 0.f;
                    if (f.ok[k]) VecIO<T, VEC>::load(vlane + eoff[k], v[k]);
                }
                if constexpr (FUSED && !PRO)
                    if (it == 0) fused_softmax<NIT, LPC>(ws);      // overlaps the first gathers (see forward)
                // scatter grad_value; dot products of grad_out with the four corners
                float t[4];
                if constexpr (MERGE && VEC == 4) {
                    // Warp-aggregated scatter: the 16 corners of this iteration (4 points x 4 corners) all
                    // carry the same grad_out vector, so corners that land on the same pixel are merged by
                    // adding their scalar weights; one RED per distinct pixel.  Lanes sub<4 of every group
                    // stand for corner k=sub; match.any finds equal pixels, two pointer-jumping rounds sum
                    // up to 4 chained duplicates, every 4th chain member issues.
                    const int kk = sub & 3;
                    const bool okk = kk == 0 ? f.ok[0] : (kk == 1 ? f.ok[1] : (kk == 2 ? f.ok[2] : f.ok[3]));
                    const bool part = (sub < 4) && okk;
                    const float wyk = (kk >> 1) ? f.lh : f.hh, wxk = (kk & 1) ? f.lw : f.hw;
                    const int pix = (f.y0 + (kk >> 1)) * W[l] + f.x0 + (kk & 1);
                    const int key = part ? pix : -1 - lane;
                    float sum = part ? ws[it] * wyk * wxk : 0.f;
                    const unsigned mask = __match_any_sync(0xffffffffu, key);
                    const unsigned above = mask & ~((2u << lane) - 1u);
                    int nxt = above ? __ffs(above) - 1 : -1;
#pragma unroll
                    for (int round = 0; round < 2; ++round) {
                        const int src = nxt < 0 ? lane : nxt;
                        const float sv = __shfl_sync(0xffffffffu, sum, src);
                        const int nv = __shfl_sync(0xffffffffu, nxt, src);
                        sum += nxt < 0 ? 0.f : sv;
                        nxt = nxt < 0 ? -1 : nv;
                    }
                    const int pos = __popc(mask & ((1u << lane) - 1u));
                    const float merged = (part && (pos & 3) == 0) ? sum : 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float wk = __shfl_sync(0xffffffffu, merged, (lane & 24) | k);
                        if (wk != 0.f) {
                            float r[VEC];
#pragma unroll
                            for (int c = 0; c < VEC; ++c) r[c] = wk * go[c];
                            VecIO<float, VEC>::red_add(gvlane + eoff[k], r);
                        }
                        t[k] = 0.f;
#pragma unroll
                        for (int c = 0; c < VEC; ++c) t[k] = fmaf(go[c], v[k][c], t[k]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (f.ok[k]) {
                            const float cwk = ws[it] * wy[k >> 1] * wx[k & 1];
                            float r[VEC];
#pragma unroll
                            for (int c = 0; c < VEC; ++c) r[c] = cwk * go[c];
                            VecIO<float, VEC>::red_add(gvlane + eoff[k], r);
                        }
                        t[k] = 0.f;
#pragma unroll
                        for (int c = 0; c < VEC; ++c) t[k] = fmaf(go[c], v[k][c], t[k]);
                    }
                }
                // <grad_out, sample>, d/dw_im, d/dh_im restricted to this lane's channels
                float pa = f.hh * (f.hw * t[0] + f.lw * t[1]) + f.lh * (f.hw * t[2] + f.lw * t[3]);
                float px = f.hh * (t[1] - t[0]) + f.lh * (t[3] - t[2]);
                float py = f.hw * (t[2] - t[0]) + f.lw * (t[3] - t[1]);
#pragma unroll
                for (int o = LPC / 2; o > 0; o >>= 1) {
                    pa += __shfl_xor_sync(0xffffffffu, pa, o);
                    px += __shfl_xor_sync(0xffffffffu, px, o);
                    py += __shfl_xor_sync(0xffffffffu, py, o);
                }
                if constexpr (FUSED) {
                    // d loc / d offset = 1 / (W, H): the W, H factors of grad_loc cancel
                    ga_keep[it] = f.in_range ? pa : 0.f;
                    gx_keep[it] = f.in_range ? ws[it] * px : 0.f;
                    gy_keep[it] = f.in_range ? ws[it] * py : 0.f;
                } else if (sub == 0) {
                    // skipped points leave zeros (reference outputs are zero-initialised)
                    const size_t k = qm * LP + it * LG + lg;
                    p.grad_attn[k] = f.in_range ? pa : 0.f;
                    reinterpret_cast<float2 *>(p.grad_loc)[k] =
                        f.in_range ? make_float2(Wf[l] * ws[it] * px, Hf[l] * ws[it] * py) : make_float2(0.f, 0.f);
                }
            }
            if constexpr (FUSED) {
                // softmax backward: grad_logit_i = a_i * (ga_i - sum_j a_j ga_j)
                float dot = 0.f;
#pragma unroll
                for (int it = 0; it < NIT; ++it) dot = fmaf(ws[it], ga_keep[it], dot);
#pragma unroll
                for (int o = LPC; o < 32; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (sub == 0) {
#pragma unroll
                    // contiguous: element (row * M + m) * LP + j of grad_attn / float2 of grad_loc; packed: row * ld (+ m * LP + j)
                    const size_t row = static_cast<size_t>(b) * p.Lq + q;
                    const size_t ka = p.ld_packed ? row * p.ld_packed + m * LP : qm * LP;
                    const size_t kl = p.ld_packed ? row * (p.ld_packed / 2) + m * LP : qm * LP;
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        p.grad_attn[ka + it * LG + lg] = ws[it] * (ga_keep[it] - dot);
                        reinterpret_cast<float2 *>(p.grad_loc)[kl + it * LG + lg] = make_float2(gx_keep[it], gy_keep[it]);
                    }
                }
            }
        }
        if (TMA) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[slot]);
        }
    });
}

}  // namespace bm2f

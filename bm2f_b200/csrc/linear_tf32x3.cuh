// tcgen05 GEMM for the MSDeformAttn projections (value_proj, sampling_offsets||attention_weights,
// output_proj — reference: nn.Linear x4 in ops/modules/ms_deform_attn.py:59-62,98-102,124):
//
//     Y[M, N] = X[M, K] * W[N, K]^T + bias[N]          K = 256 (d_model), N in {96..288}, fp32 in / out
//
// fp32 parity matters (the reference runs these layers in fp32, msdeformattn.py:314-320), so the product
// is computed as a 3-term TF32 split on the 5th-generation tensor cores:
//     x = x_hi + x_lo,  w = w_hi + w_lo    (hi = top 19 bits, exactly representable in TF32)
//     Y ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo                (the dropped x_lo*w_lo term is ~2^-22 relative)
// with fp32 accumulation in tensor memory (TMEM).
//
// One CTA per 128-row tile of X, 6 warps:
//   warp 0      TMEM allocation; lane 0 = TMA producer of the W_hi / W_lo k-blocks (tensor maps, 128B swizzle)
//   warp 1      lane 0 = MMA issuer: tcgen05.mma.cta_group::1.kind::tf32, M=128, N=tile width, K=8 per instruction
//   warps 2..5  X producers: coalesced 128-bit global loads, split into hi / lo, written to shared memory in
//               the UMMA K-major SWIZZLE_128B layout; after the main loop the same warps are the epilogue:
//               tcgen05.ld (TMEM -> registers), + bias, swizzled staging tile, TMA store (cp.async.bulk.tensor).
// Two shared-memory stages of one k-block (32 floats = one 128-byte swizzle row) each, full/empty mbarriers;
// tcgen05.commit releases a stage when the MMAs that read it have retired.
#pragma once

#include "msda_common.cuh"

namespace bm2f {

constexpr int kGemmKMax = 4096;     // reduction length is a runtime multiple of 32 (256 for the projections, 1024 in the FFN)
constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 32;     // floats per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int kGemmStages = 2;
constexpr int kGemmThreads = 192;

struct LinearParams {
    const float *x;      // (M, K)
    const float *bias;   // (N) or nullptr
    float *y;            // (M, N)
    int M, N, K;        // N = total output width (row stride of y and rows of the weight)
    int slices;         // persistent kernel: N / NT column slices, tile = (row tile, slice)
    int relu;           // epilogue: y = max(y, 0)
    int store_mode;     // persistent kernel epilogue: 0 = per-warp TMA store, 1 = coalesced 128-bit stores through smem
    const float *out_mask;  // optional (M, N): epilogue zeroes y where out_mask <= 0 (ReLU backward of the previous layer)
    const float *addend;    // optional (M, N): epilogue adds it (gradient accumulation; may alias y: a tile is read, then
                            // written, by the one CTA that owns it)
    int split;           // 3 = tf32x3 (fp32-grade), 1 = single TF32 pass
    // CONV instantiation (3 x 3 convolution over token rows, conv3x3 in fpn_api.cu): x is the zero-haloed image
    // (batch, conv_h + 2, conv_w + 2, ldx) and M counts its rows; K = 9 * ldx, k-block kb belongs to tap kb / (ldx / 32)
    // whose activation row is the tile row shifted by (tap / 3 - 1) * (conv_w + 2) + (tap % 3 - 1); y is the dense
    // (batch, conv_h, conv_w, N) image: halo rows of a tile are computed and dropped
    int ldx, conv_w, conv_h;
};


// x = hi + lo with hi exactly representable in TF32 (10 explicit mantissa bits), rounded to nearest so that the
// low parts have no sign bias (matters for the 10^5-term weight-gradient reductions)
__device__ __forceinline__ float tf32_hi(float x)
{
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));      // one instruction: round to nearest, low 13 bits zero
    return __uint_as_float(u);
}

// ---- tcgen05 / TMEM wrappers -------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 128-bit store to a shared-memory address (state space known: STS.128 instead of a generic ST.E.128)
__device__ __forceinline__ void sts128(uint32_t smem_addr, const float4 &v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t smem_addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr) : "memory");
    return v;
}

// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// same, delivered to the barrier at this shared-memory offset in every CTA of `cta_mask` (thread-block cluster)
__device__ __forceinline__ void umma_commit_multicast(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
// 2-D tiled TMA load whose box lands at the same shared-memory offset of every CTA in `cta_mask`; each destination
// CTA's barrier (same offset) receives the byte count
__device__ __forceinline__ void tma_load_2d_multicast(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar,
                                                      uint16_t cta_mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread i <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, one 128-byte row per matrix row, 8-row atoms of
// 1024 bytes (stride byte offset).  Field layout = cute::UMMA::SmemDescriptor (sm_100).
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);          // [0,14)  start address >> 4
    d |= static_cast<uint64_t>(0) << 16;                             // [16,30) leading byte offset (unused, 1 atom in K)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                     // [32,46) stride byte offset >> 4
    d |= static_cast<uint64_t>(1) << 46;                             // [46,48) descriptor version (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                             // [61,64) layout: SWIZZLE_128B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulator, TF32 A/B, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n)
{
    return (1u << 4)                                   // c_format = F32
           | (2u << 7)                                 // a_format = TF32
           | (2u << 10)                                // b_format = TF32
           | (0u << 15) | (0u << 16)                   // A, B K-major
           | (static_cast<uint32_t>(n >> 3) << 17)     // N / 8
           | (static_cast<uint32_t>(m >> 4) << 24);    // M / 16
}

// mbarrier wait with a poll budget: a descriptor / barrier bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t i = 0; i < (1u << 24); ++i) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

// NT = tile width in columns: the whole N of the layer (N <= 256) or half of it (N = 288 -> 2 x 144).
// NH = number of such tiles per CTA (1 or 2); TMEM columns used = NH * NT.
template <int NT, int NH>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_tf32x3_kernel(const LinearParams p, const __grid_constant__ CUtensorMap tm_whi,
                     const __grid_constant__ CUtensorMap tm_wlo, const __grid_constant__ CUtensorMap tm_y)
{
    constexpr int N = NT * NH;
    constexpr int kXBytes = kGemmBlockM * kGemmBlockK * 4;              // 16 KB
    constexpr int kWBytes = N * kGemmBlockK * 4;                        // N x 128 B
    constexpr int kStageBytes = 2 * kXBytes + 2 * kWBytes;
    const int kKBlocks = p.K / kGemmBlockK;                             // 8 for the forward layers
    constexpr uint32_t kTmemCols = (N <= 32) ? 32 : (N <= 64) ? 64 : (N <= 128) ? 128 : (N <= 256) ? 256 : 512;
    static_assert(NT % 16 == 0 && NT <= 256, "UMMA N for M = 128");
    static_assert((NT * 128) % 1024 == 0, "second tile must start on a swizzle atom");

    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment; the launch allocates 1 KB of slack for this round-up
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[kGemmStages], empty_bar[kGemmStages], acc_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kGemmBlockM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmStages; ++s) {
            mbar_init(&full_bar[s], 128 + 1);   // 128 X-producer threads + the TMA producer's expect_tx arrive
            mbar_init(&empty_bar[s], 1);        // tcgen05.commit
        }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tm_whi);
        tma_prefetch_desc(&tm_wlo);
        tma_prefetch_desc(&tm_y);
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    auto stage_ptr = [&](int s) { return smem + s * kStageBytes; };

    if (warp == 0) {
        // ---------------- TMA producer: W_hi and W_lo k-blocks ----------------
        if (lane == 0) {
            for (int kb = 0; kb < kKBlocks; ++kb) {
                const int s = kb % kGemmStages;
                mbar_wait_bounded(&empty_bar[s], ((kb / kGemmStages) & 1) ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], p.split == 3 ? 2 * kWBytes : kWBytes);
                unsigned char *w_hi = stage_ptr(s) + 2 * kXBytes;
                unsigned char *w_lo = w_hi + kWBytes;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    tma_load_2d(w_hi + h * NT * 128, &tm_whi, kb * kGemmBlockK, h * NT, &full_bar[s]);
                    if (p.split == 3) tma_load_2d(w_lo + h * NT * 128, &tm_wlo, kb * kGemmBlockK, h * NT, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kGemmBlockM, NT);
            for (int kb = 0; kb < kKBlocks; ++kb) {
                const int s = kb % kGemmStages;
                mbar_wait_bounded(&full_bar[s], (kb / kGemmStages) & 1);
                tc_fence_after();
                const uint32_t x_hi = smem_u32(stage_ptr(s));
                const uint32_t x_lo = x_hi + kXBytes;
                const uint32_t w_hi = x_hi + 2 * kXBytes;
                const uint32_t w_lo = w_hi + kWBytes;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const uint32_t d = tmem_base + h * NT;
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 8; ++k) {
                        const uint32_t koff = k * 32;     // 8 floats along K inside the 128-byte swizzle row
                        const uint64_t a_hi = umma_desc_k128(x_hi + koff);
                        const uint64_t b_hi = umma_desc_k128(w_hi + h * NT * 128 + koff);
                        umma_tf32(d, a_hi, b_hi, idesc, (kb | k) ? 1u : 0u);
                        if (p.split == 3) {
                            const uint64_t a_lo = umma_desc_k128(x_lo + koff);
                            const uint64_t b_lo = umma_desc_k128(w_lo + h * NT * 128 + koff);
                            umma_tf32(d, a_lo, b_hi, idesc, 1u);
                            umma_tf32(d, a_hi, b_lo, idesc, 1u);
                        }
                    }
                }
                umma_commit(&empty_bar[s]);            // stage reusable once these MMAs retire
            }
            umma_commit(&acc_bar);                     // accumulator complete
        }
    } else {
        // ---------------- X producers (warps 2..5), then epilogue ----------------
        const int t = threadIdx.x - 64;                // 0..127
        const int c16 = t & 7;                         // 16-byte chunk inside the 128-byte row
        const int rsub = t >> 3;                       // 0..15
        auto load_x = [&](int kb, float4 (&v)[8]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int gr = row0 + rsub + 16 * j;
                v[j] = gr < p.M ? __ldg(reinterpret_cast<const float4 *>(p.x + static_cast<size_t>(gr) * p.K +
                                                                           kb * kGemmBlockK) + c16)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        float4 cur[8], nxt[8];
        load_x(0, cur);
        for (int kb = 0; kb < kKBlocks; ++kb) {
            const int s = kb % kGemmStages;
            if (kb + 1 < kKBlocks) load_x(kb + 1, nxt);        // next k-block's loads fly while this one is converted
            mbar_wait_bounded(&empty_bar[s], ((kb / kGemmStages) & 1) ^ 1);
            unsigned char *x_hi = stage_ptr(s);
            unsigned char *x_lo = x_hi + kXBytes;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = rsub + 16 * j;
                const uint32_t off = r * 128 + ((c16 ^ (r & 7)) << 4);      // SWIZZLE_128B
                float4 hi, lo;
                hi.x = tf32_hi(cur[j].x); lo.x = cur[j].x - hi.x;
                hi.y = tf32_hi(cur[j].y); lo.y = cur[j].y - hi.y;
                hi.z = tf32_hi(cur[j].z); lo.z = cur[j].z - hi.z;
                hi.w = tf32_hi(cur[j].w); lo.w = cur[j].w - hi.w;
                *reinterpret_cast<float4 *>(x_hi + off) = hi;
                *reinterpret_cast<float4 *>(x_lo + off) = lo;
            }
            fence_async_smem();                        // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(&full_bar[s]);
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
        }
        // epilogue: thread <-> one output row; warp w may touch TMEM lanes 32*(w%4) .. +31.  Each 32-column
        // chunk goes TMEM -> registers (+bias) -> a SWIZZLE_128B staging tile in the (now idle) stage memory
        // -> one TMA store of a 128 x 32 box; rows past M are clipped by the tensor map.
        mbar_wait_bounded(&acc_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int r = q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0, it = 0; c0 < N; c0 += 32, ++it) {
            float acc[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, acc);
            unsigned char *stg = smem + (it & 1) * kXBytes;             // two 16 KB staging tiles
            if (it >= 2) {                                               // tile written two chunks ago must be drained
                if (t == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                float4 o = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
                if (p.bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + c));
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                *reinterpret_cast<float4 *>(stg + r * 128 + ((((c >> 2) ^ (r & 7))) << 4)) = o;
            }
            fence_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (t == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                 reinterpret_cast<uint64_t>(&tm_y)),
                             "r"(smem_u32(stg)), "r"(c0), "r"(row0)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}


// ------------------------------------------------------------------------------------------------------
// Persistent variant (N <= 256): one CTA per SM loops over its row tiles; the accumulator is double-buffered in
// TMEM so the epilogue of tile i (warps 6..9: TMEM -> staging tile -> TMA store) overlaps the MMAs of tile
// i+1, and the shared-memory stage ring / X register prefetch run across tile boundaries.
// Warps: 0 TMA (weights), 1 MMA issuer, 2..5 activation producers, 6..9 epilogue.
// ------------------------------------------------------------------------------------------------------
constexpr int kGemmProducerWarps = 4;   // activation producers (8 measured slower at N = 256: the weight stream from L2 is the
                                        // contended resource there; faster only for the 96-wide layer)
constexpr int kGemmThreadsPersistent = (2 + kGemmProducerWarps + 4) * 32;
constexpr int gemm_threads_persistent(int producer_warps) { return (2 + producer_warps + 4) * 32; }

// ------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2).  With both operands in shared memory a single-CTA TF32 MMA (M = 128, N = 256,
// K = 8: 128 cycles) reads A 4 KB + B 8 KB = 96 B/cycle of the SM's 128 B/cycle shared-memory bandwidth; together with the
// TMA weight writes (42 B/cycle) and the producers' hi/lo stores (21 B/cycle) the k-loop needs 159 B/cycle and runs at
// ~1.0 us per k-block instead of the tensor pipe's 0.8 us (measured: 0.87-1.23 us; more loads in flight, more producer
// warps and multicast weights all changed nothing).  In a CTA pair the M = 256 MMA reads each CTA's own 128 rows of A and
// only HALF of B from each CTA (64 B/cycle), every CTA stages only its half of the weight k-block (21 B/cycle), and
// the smaller stage buys a third pipeline stage: 106 B/cycle, tensor-bound.
//   leader (cluster rank 0): issues tcgen05.mma.cta_group::2 for both CTAs once BOTH stages are full; commits are
//       multicast to the stage / accumulator barriers of both CTAs.
//   peer (rank 1): same TMA / producer / epilogue roles on its own row tile; its idle MMA warp relays "stage full" to the
//       leader with a remote mbarrier arrive, its epilogue threads hand accumulators back the same way.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_map_shared(uint32_t cta_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier that receives arrivals from the peer CTA (cluster-scope acquire), bounded like mbar_wait_bounded
__device__ __forceinline__ void mbar_wait_cluster_bounded(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t i = 0; i < (1u << 24); ++i) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *smem_slot, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

constexpr int kPairStages = 3;
constexpr int kXtmaStages = 4;

// CL = 2: thread-block cluster of two CTAs working on neighbouring row tiles in lockstep.  Each CTA fetches HALF of every
// weight k-block and TMA-multicasts it to both, so the weight stream out of L2 (the contended resource of this kernel:
// 512 KB per 128-row tile) is halved; a stage is released to both TMA producers by multicast tcgen05.commit arrivals.
// PW activation-producer warps keep XD k-blocks of X in flight in registers: (XD - 1) * 16 KB outstanding per SM.
// XTMA (single TF32 pass only): the activation tile needs no hi/lo split, so it is TMA-loaded straight into the MMA
// stage like the weights (the tensor core ignores the low 13 mantissa bits); the producer warps idle and the freed
// shared memory holds four stages.
// RT = 2 (XTMA + CONV only): two row tiles share every weight k-block — both accumulators are filled side by side, the
// weight stream out of L2 and the weight writes into shared memory halve per MMA (the 3 x 3 convolution re-reads 2.4 MB
// of weights per tile and is bound by exactly that); the epilogue of a tile pair is then not overlapped with MMAs (4 us
// after a 39 us main loop of 72 k-blocks).
template <int NT, int CL = 1, int PW = kGemmProducerWarps, int XD = 3, bool PAIR = false, bool XTMA = false, bool CONV = false,
          int RT = 1>
__global__ void __launch_bounds__(gemm_threads_persistent(PW), 1)
linear_tf32x3_persistent_kernel(const LinearParams p, const __grid_constant__ CUtensorMap tm_whi,
                                const __grid_constant__ CUtensorMap tm_wlo, const __grid_constant__ CUtensorMap tm_y,
                                const __grid_constant__ CUtensorMap tm_x)
{
    constexpr int N = NT;
    constexpr int kXBytes = kGemmBlockM * kGemmBlockK * 4;
    static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
    static_assert(!XTMA || (CL == 1 && !PAIR), "XTMA: single CTAs");
    static_assert(!CONV || (CL == 1 && !PAIR), "CONV: single CTAs");
    static_assert(RT == 1 || (RT == 2 && XTMA && CONV), "two row tiles per weight block: TMA-loaded convolution only");
    constexpr int kStages = RT == 2 ? 3 : XTMA ? kXtmaStages : PAIR ? kPairStages : kGemmStages;
    constexpr int kWBytes = (PAIR ? N / 2 : N) * kGemmBlockK * 4;   // PAIR: this CTA's half of the weight k-block
    constexpr int kSlots = XTMA ? 1 : 2;                            // hi (+ lo) tiles per operand and stage
    constexpr int kStageBytes = RT * kSlots * kXBytes + kSlots * kWBytes;
    constexpr uint32_t kTmemCols = (2 * N <= 64) ? 64 : (2 * N <= 128) ? 128 : (2 * N <= 256) ? 256 : 512;
    static_assert(NT % 16 == 0 && NT <= 256, "UMMA N for M = 128, two accumulators in 512 TMEM columns");

    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *staging = smem + kStages * kStageBytes;            // 2 x 16 KB epilogue tiles
    __shared__ uint64_t full_bar[kStages], empty_bar[kStages], acc_full[2], acc_empty[2];
    __shared__ uint64_t peer_full[kStages], peer_acc_empty[2];      // PAIR, leader only: arrivals from the peer CTA
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kKBlocks = p.K / kGemmBlockK;
    const int slices = p.slices;
    // the cluster (CL CTAs) walks virtual tiles = (group of CL row tiles, slice); CTA `crank` owns row tile group * CL + crank
    const int crank = CL > 1 ? static_cast<int>(blockIdx.x) % CL : 0;
    const int cid = static_cast<int>(blockIdx.x) / CL, ncl = static_cast<int>(gridDim.x) / CL;
    const int row_tiles = (p.M + RT * kGemmBlockM - 1) / (RT * kGemmBlockM);      // RT = 2: pairs of row tiles
    const int num_tiles = ((row_tiles + CL - 1) / CL) * slices;   // tile = row_tile_group * slices + slice
    // (row tile, slice) of a tile index advance incrementally: an integer division per tile / k-block in the
    // single-threaded TMA and MMA roles costs ~200 cycles of latency each and made this kernel 1.7x slower (measured)
    const int step_rt = ncl / slices, step_sl = ncl % slices;
    struct TileIter {
        int tile, rt, sl;
    };
    auto first_tile = [&]() { return TileIter{cid, cid / slices, cid % slices}; };
    auto next_tile = [&](TileIter &t) {
        t.tile += ncl;
        t.rt += step_rt;
        t.sl += step_sl;
        if (t.sl >= slices) { t.sl -= slices; ++t.rt; }
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], XTMA ? 1 : PW * 32 + 1);
            mbar_init(&empty_bar[s], PAIR ? 1 : CL);   // a commit from every CTA of the cluster / the leader's multicast commit
            if (PAIR) mbar_init(&peer_full[s], 1);     // relay thread of the peer CTA
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);        // tcgen05.commit after the last k-block of a tile
            mbar_init(&acc_empty[b], 128);     // every epilogue thread, after its last TMEM read of the tile
            if (PAIR) mbar_init(&peer_acc_empty[b], 128);   // the peer's epilogue threads (remote arrives)
        }
        fence_mbar_init();
        tma_prefetch_desc(&tm_whi);
        tma_prefetch_desc(&tm_wlo);
        tma_prefetch_desc(&tm_y);
        if (XTMA) tma_prefetch_desc(&tm_x);
    }
    if (warp == 0) {
        if (PAIR) tmem_alloc_pair(&tmem_base_slot, kTmemCols);
        else tmem_alloc(&tmem_base_slot, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();        // peers' barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    auto stage_ptr = [&](int s) { return smem + s * kStageBytes; };

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer (W_hi, W_lo) ----
            int g = 0;
            for (TileIter ti = first_tile(); ti.tile < num_tiles; next_tile(ti)) {
                const int wrow = ti.sl * NT;
                int kc = 0, tdx = -1, tdy = -1;          // CONV: channel block and tap offset of this k-block
                for (int kb = 0; kb < kKBlocks; ++kb, ++g) {
                    const int s = g % kStages;
                    mbar_wait_bounded(&empty_bar[s], ((g / kStages) & 1) ^ 1);
                    uint64_t *tbar = &full_bar[s];
                    mbar_arrive_expect_tx(tbar, XTMA ? kWBytes + RT * kXBytes : p.split == 3 ? 2 * kWBytes : kWBytes);
                    if (XTMA) {    // rows past M (CONV: also before row 0) are zero-filled by the TMA unit
                        if (CONV) {
#pragma unroll
                            for (int t = 0; t < RT; ++t)
                                tma_load_2d(stage_ptr(s) + t * kXBytes, &tm_x, kc * kGemmBlockK,
                                            (ti.rt * RT + t) * kGemmBlockM + tdy * (p.conv_w + 2) + tdx, tbar);
                            if (++kc == p.ldx / kGemmBlockK) {
                                kc = 0;
                                if (++tdx == 2) { tdx = -1; ++tdy; }
                            }
                        } else {
                            tma_load_2d(stage_ptr(s), &tm_x, kb * kGemmBlockK, (ti.rt * CL + crank) * kGemmBlockM, tbar);
                        }
                    }
                    unsigned char *w_hi = stage_ptr(s) + RT * kSlots * kXBytes;
                    if (CL == 1) {
                        tma_load_2d(w_hi, &tm_whi, kb * kGemmBlockK, wrow, tbar);
                        if (p.split == 3) tma_load_2d(w_hi + kWBytes, &tm_wlo, kb * kGemmBlockK, wrow, tbar);
                    } else if (PAIR) {
                        // this CTA's half of the weight rows, into its own shared memory only
                        constexpr int kShare = NT / 2;
                        tma_load_2d(w_hi, &tm_whi, kb * kGemmBlockK, wrow + crank * kShare, tbar);
                        if (p.split == 3)
                            tma_load_2d(w_hi + kWBytes, &tm_wlo, kb * kGemmBlockK, wrow + crank * kShare, tbar);
                    } else {
                        // this CTA's share of the weight rows, delivered to every CTA of the cluster
                        constexpr int kShare = NT / CL;
                        constexpr uint16_t kMask = (1u << CL) - 1;
                        unsigned char *dst = w_hi + crank * kShare * 128;
                        tma_load_2d_multicast(dst, &tm_whi, kb * kGemmBlockK, wrow + crank * kShare, tbar, kMask);
                        if (p.split == 3)
                            tma_load_2d_multicast(dst + kWBytes, &tm_wlo, kb * kGemmBlockK, wrow + crank * kShare,
                                                  tbar, kMask);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if constexpr (PAIR) {
            if (lane == 0 && crank == 0) {                     // ---- MMA issuer (leader CTA, for the pair) ----
                constexpr uint32_t idesc = umma_idesc_tf32(2 * kGemmBlockM, NT);
                int g = 0, i = 0;
                for (int tile = cid; tile < num_tiles; tile += ncl, ++i) {
                    const int b = i & 1;
                    mbar_wait_bounded(&acc_empty[b], ((i >> 1) & 1) ^ 1);               // own epilogue has drained it
                    mbar_wait_cluster_bounded(&peer_acc_empty[b], ((i >> 1) & 1) ^ 1);  // and the peer's
                    tc_fence_after();
                    const uint32_t d = tmem_base + b * NT;
                    for (int kb = 0; kb < kKBlocks; ++kb, ++g) {
                        const int s = g % kStages;
                        mbar_wait_bounded(&full_bar[s], (g / kStages) & 1);
                        mbar_wait_cluster_bounded(&peer_full[s], (g / kStages) & 1);
                        tc_fence_after();
                        const uint32_t x_hi = smem_u32(stage_ptr(s));
                        const uint32_t x_lo = x_hi + kXBytes;
                        const uint32_t w_hi = x_hi + kSlots * kXBytes;
                        const uint32_t w_lo = w_hi + kWBytes;
    #pragma unroll
                        for (int k = 0; k < kGemmBlockK / 8; ++k) {
                            const uint32_t koff = k * 32;
                            const uint64_t a_hi = umma_desc_k128(x_hi + koff);
                            const uint64_t b_hi = umma_desc_k128(w_hi + koff);
                            umma_tf32_pair(d, a_hi, b_hi, idesc, (kb | k) ? 1u : 0u);
                            if (p.split == 3) {
                                umma_tf32_pair(d, umma_desc_k128(x_lo + koff), b_hi, idesc, 1u);
                                umma_tf32_pair(d, a_hi, umma_desc_k128(w_lo + koff), idesc, 1u);
                            }
                        }
                        umma_commit_pair(&empty_bar[s], 3);     // stage s is free again in both CTAs
                    }
                    umma_commit_pair(&acc_full[b], 3);           // both halves of the accumulator are complete
                }
            } else if (lane == 0) {                              // ---- peer CTA: relay "my stage is full" to the leader ----
                int g = 0;
                for (int tile = cid; tile < num_tiles; tile += ncl) {
                    for (int kb = 0; kb < kKBlocks; ++kb, ++g) {
                        const int s = g % kStages;
                        mbar_wait_bounded(&full_bar[s], (g / kStages) & 1);
                        mbar_arrive_remote(cluster_map_shared(smem_u32(&peer_full[s]), 0));
                    }
                }
            }
        } else
        if (lane == 0) {                                   // ---- MMA issuer ----
            constexpr uint32_t idesc = umma_idesc_tf32(kGemmBlockM, NT);
            int g = 0, i = 0;
            for (int tile = cid; tile < num_tiles; tile += ncl, ++i) {
                const int b = RT == 2 ? 0 : i & 1;
                if (RT == 2) {                                              // both accumulators belong to this tile pair
                    mbar_wait_bounded(&acc_empty[0], (i & 1) ^ 1);
                    mbar_wait_bounded(&acc_empty[1], (i & 1) ^ 1);
                } else {
                    mbar_wait_bounded(&acc_empty[b], ((i >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                }
                tc_fence_after();
                const uint32_t d = tmem_base + b * NT;
                for (int kb = 0; kb < kKBlocks; ++kb, ++g) {
                    const int s = g % kStages;
                    mbar_wait_bounded(&full_bar[s], (g / kStages) & 1);
                    tc_fence_after();
                    const uint32_t x_hi = smem_u32(stage_ptr(s));
                    const uint32_t x_lo = x_hi + kXBytes;
                    const uint32_t w_hi = x_hi + RT * kSlots * kXBytes;
                    const uint32_t w_lo = w_hi + kWBytes;
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 8; ++k) {
                        const uint32_t koff = k * 32;
                        const uint64_t a_hi = umma_desc_k128(x_hi + koff);
                        const uint64_t b_hi = umma_desc_k128(w_hi + koff);
                        umma_tf32(d, a_hi, b_hi, idesc, (kb | k) ? 1u : 0u);
                        if (RT == 2) umma_tf32(d + NT, umma_desc_k128(x_hi + kXBytes + koff), b_hi, idesc, (kb | k) ? 1u : 0u);
                        if (p.split == 3) {
                            umma_tf32(d, umma_desc_k128(x_lo + koff), b_hi, idesc, 1u);
                            umma_tf32(d, a_hi, umma_desc_k128(w_lo + koff), idesc, 1u);
                        }
                    }
                    if (CL == 1) umma_commit(&empty_bar[s]);
                    else umma_commit_multicast(&empty_bar[s], (1u << CL) - 1);
                }
                umma_commit(&acc_full[b]);
                if (RT == 2) umma_commit(&acc_full[1]);
            }
        }
    } else if (warp < 2 + PW) {
        if constexpr (!XTMA) {
        // ---- X producers: global -> registers -> hi/lo -> swizzled shared memory ----
        constexpr int kRowsPerPass = PW * 4;          // 8 threads per 128-byte row segment
        constexpr int kPasses = kGemmBlockM / kRowsPerPass;
        const int t = threadIdx.x - 64;
        const int c16 = t & 7, rsub = t >> 3;
        auto load_x = [&](int rt, int kb, float4 (&v)[kPasses]) {
            int rbase = rt * kGemmBlockM + rsub;
            size_t cbase = static_cast<size_t>(kb) * kGemmBlockK;
            const int ld = CONV ? p.ldx : p.K;
            if (CONV) {            // k-block -> (tap, channel block); the tap shifts the row inside the zero-haloed image
                const int cpk = p.ldx / kGemmBlockK, tap = kb / cpk;
                cbase = static_cast<size_t>(kb - tap * cpk) * kGemmBlockK;
                rbase += (tap / 3 - 1) * (p.conv_w + 2) + (tap % 3 - 1);
            }
#pragma unroll
            for (int j = 0; j < kPasses; ++j) {
                const int gr = rbase + kRowsPerPass * j;
                v[j] = (gr < p.M && (!CONV || gr >= 0))
                           ? __ldg(reinterpret_cast<const float4 *>(p.x + static_cast<size_t>(gr) * ld + cbase) + c16)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // kXDepth k-blocks of X are in flight per thread (registers): with one block in flight the kernel is bound by
        // global-load latency (16 KB per SM outstanding), measured 2x slower on the 1024-wide FFN shapes
        constexpr int kXDepth = XD;
        float4 buf[kXDepth][kPasses];
        TileIter lt = first_tile(), ct = first_tile();   // load iterator (kXDepth - 1 k-blocks ahead), convert iterator
        int lkb = 0, kb = 0, g = 0;
        auto advance = [&](TileIter &tl, int &k) {
            if (++k == kKBlocks) { k = 0; next_tile(tl); }
        };
#pragma unroll
        for (int d = 0; d < kXDepth - 1; ++d) {
            if (lt.tile < num_tiles) load_x(lt.rt * CL + crank, lkb, buf[d]);
            advance(lt, lkb);
        }
        while (ct.tile < num_tiles) {
#pragma unroll
            for (int u = 0; u < kXDepth; ++u) {
                if (ct.tile >= num_tiles) break;
                if (lt.tile < num_tiles) load_x(lt.rt * CL + crank, lkb, buf[(u + kXDepth - 1) % kXDepth]);
                advance(lt, lkb);
                const int s = g % kStages;
                mbar_wait_bounded(&empty_bar[s], ((g / kStages) & 1) ^ 1);
                const uint32_t x_hi_s = smem_u32(stage_ptr(s));
                const uint32_t x_lo_s = x_hi_s + kXBytes;
#pragma unroll
                for (int j = 0; j < kPasses; ++j) {
                    const int r = rsub + kRowsPerPass * j;
                    const uint32_t off = r * 128 + ((c16 ^ (r & 7)) << 4);
                    const float4 v = buf[u][j];
                    float4 hi, lo;
                    hi.x = tf32_hi(v.x); lo.x = v.x - hi.x;
                    hi.y = tf32_hi(v.y); lo.y = v.y - hi.y;
                    hi.z = tf32_hi(v.z); lo.z = v.z - hi.z;
                    hi.w = tf32_hi(v.w); lo.w = v.w - hi.w;
                    sts128(x_hi_s + off, hi);
                    sts128(x_lo_s + off, lo);
                }
                fence_async_smem();
                mbar_arrive(&full_bar[s]);
                advance(ct, kb);
                ++g;
            }
        }
        }
    } else {
        // ---- epilogue (last 4 warps; warp % 4 selects the TMEM lane quarter) ----
        const int q = warp & 3;
        int i = 0, chunk = 0;
        for (TileIter ti = first_tile(); ti.tile < num_tiles; next_tile(ti), ++i) {
          for (int sub = 0; sub < RT; ++sub) {                  // RT = 2: the two row tiles of the pair, one accumulator each
            const int b = RT == 2 ? sub : i & 1;
            const int rt_own = RT == 2 ? ti.rt * 2 + sub : ti.rt * CL + crank;      // this CTA's row tile of the group
            // CONV: the 8 rows this lane stores per 32-column chunk, as rows of the dense image (-1: halo or past the end)
            int conv_row[8];
            if (CONV) {
                const int wp = p.conv_w + 2, hp = p.conv_h + 2;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int pr = rt_own * kGemmBlockM + q * 32 + it * 4 + (lane >> 3);
                    const int img = pr / (hp * wp), rem = pr - img * (hp * wp);
                    const int yp = rem / wp, xp = rem - yp * wp;
                    conv_row[it] = (pr < p.M && yp >= 1 && yp <= p.conv_h && xp >= 1 && xp <= p.conv_w)
                                       ? (img * p.conv_h + yp - 1) * p.conv_w + xp - 1 : -1;
                }
            }
            mbar_wait_bounded(&acc_full[b], RT == 2 ? i & 1 : (i >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 32, ++chunk) {
                float acc[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * NT + c0, acc);
                if (c0 + 32 >= N) {                      // last TMEM read of this tile: hand the accumulator back
                    tc_fence_before();
                    if (!PAIR || crank == 0) mbar_arrive(&acc_empty[b]);
                    else mbar_arrive_remote(cluster_map_shared(smem_u32(&peer_acc_empty[b]), 0));
                }
                // every epilogue warp owns a private pair of 32-row x 128-byte staging tiles and issues its own
                // TMA store (box 32 x 32): no CTA-level barrier in the epilogue, four independent store pipelines
                unsigned char *stg = staging + q * (2 * 4096) + (chunk & 1) * 4096;
                const uint32_t stg_s = smem_u32(stg);
                const int grow = rt_own * kGemmBlockM + q * 32 + lane;     // this thread's output row
                if (chunk >= 2 && p.store_mode == 0 && !CONV) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    float4 o = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
                    if (p.bias && (NT % 32 == 0 || c0 + c < NT)) {      // NT = 144: the last chunk has 16 columns
                        const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.bias + ti.sl * NT + c0 + c));
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    }
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (p.addend && grow < p.M) {
                        const float4 ad = *reinterpret_cast<const float4 *>(      // plain load: addend may alias y
                            p.addend + static_cast<size_t>(grow) * p.N + ti.sl * NT + c0 + c);
                        o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
                    }
                    if (p.out_mask && grow < p.M) {
                        const float4 mk = __ldg(reinterpret_cast<const float4 *>(
                            p.out_mask + static_cast<size_t>(grow) * p.N + ti.sl * NT + c0 + c));
                        o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f;
                        o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
                    }
                    sts128(stg_s + lane * 128 + ((((c >> 2) ^ (lane & 7))) << 4), o);
                }
                if (CONV) {
                    __syncwarp();
                    const int ch = lane & 7;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = it * 4 + (lane >> 3);
                        const float4 v = *reinterpret_cast<const float4 *>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
                        if (conv_row[it] >= 0)
                            *reinterpret_cast<float4 *>(p.y + static_cast<size_t>(conv_row[it]) * p.N + ti.sl * NT + c0 + ch * 4) = v;
                    }
                    __syncwarp();
                } else if (p.store_mode == 0) {
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                         reinterpret_cast<uint64_t>(&tm_y)),
                                     "r"(smem_u32(stg)), "r"(ti.sl * NT + c0), "r"(rt_own * kGemmBlockM + q * 32)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else {
                    // transposed read-back: 8 lanes cover one row's 128 bytes, every store instruction writes 4 full lines
                    __syncwarp();
                    const int ch = lane & 7;
                    float *ybase = p.y + static_cast<size_t>(rt_own * kGemmBlockM + q * 32) * p.N + ti.sl * NT + c0 + ch * 4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = it * 4 + (lane >> 3);
                        const float4 v = *reinterpret_cast<const float4 *>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
                        if (rt_own * kGemmBlockM + q * 32 + row < p.M && (NT % 32 == 0 || c0 + ch * 4 < NT))
                            *reinterpret_cast<float4 *>(ybase + static_cast<size_t>(row) * p.N) = v;
                    }
                    __syncwarp();
                }
            }
          }
        }
        if (lane == 0 && p.store_mode == 0 && !CONV) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();        // no CTA leaves while a peer may still signal its barriers
    if (warp == 0) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int NT>
constexpr int linear_pair_smem_bytes()
{
    return kPairStages * (2 * kGemmBlockM * kGemmBlockK * 4 + 2 * (NT / 2) * kGemmBlockK * 4) + 2 * kGemmBlockM * kGemmBlockK * 4 + 1024;
}

template <int NT>
constexpr int linear_xtma_smem_bytes()
{
    return kXtmaStages * (kGemmBlockM * kGemmBlockK * 4 + NT * kGemmBlockK * 4) + 2 * kGemmBlockM * kGemmBlockK * 4 + 1024;
}

// RT = 2 convolution: 3 stages of (2 activation tiles + one weight k-block) + the epilogue staging tiles
template <int NT>
constexpr int linear_conv2_smem_bytes()
{
    return 3 * (2 * kGemmBlockM * kGemmBlockK * 4 + NT * kGemmBlockK * 4) + 2 * kGemmBlockM * kGemmBlockK * 4 + 1024;
}

template <int NT>
constexpr int linear_persistent_smem_bytes()
{
    return kGemmStages * (2 * kGemmBlockM * kGemmBlockK * 4 + 2 * NT * kGemmBlockK * 4) + 2 * kGemmBlockM * kGemmBlockK * 4 + 1024;
}

// ------------------------------------------------------------------------------------------------------
// Weight / bias gradient of the projections on tcgen05:
//     dW[n, k] = sum_r G[r, n] * X[r, k]        db[n] = sum_r G[r, n]         (r over ~10^5 rows)
// The reduction index r is the ROW of both inputs, so producer threads transpose on the fly: four rows of one
// feature become one 16-byte chunk of that feature's K-major (K = r) SWIZZLE_128B row — lane = feature makes
// both the global loads (32 x 4 B contiguous) and the swizzled 128-bit shared stores conflict-free.
//   A = G^T tile (128 features n x 32 rows), B = [X^T ; ones ; 0] tile (256 features k + 1 row of ones x 32 rows):
//   column 256 of the accumulator is the bias gradient.  TF32 3-term split as in the forward kernel.
// grid = (n tiles) x (row chunks): every CTA accumulates its chunk in TMEM and adds the 128 x 257 partial result
// to dW / db with red.global.add (outputs must be zeroed by the caller).
// ------------------------------------------------------------------------------------------------------
constexpr int kDwThreads = 320;        // warp 0: TMEM + constants, warp 1: MMA issuer, warps 2..9: transposing producers
constexpr int kDwNB = 272;             // 256 features + ones row, padded to a multiple of 16

struct LinearDwParams {
    const float *g;      // (M, N)  grad_y
    const float *x;      // (M, ldx); blockIdx.z selects a 256-feature slice of it
    float *dw;           // (N, ldx), accumulated into
    float *db;           // (N) or nullptr, accumulated into
    int M, N, ldx;
    int rows_per_chunk;  // multiple of 32
    int split;
    // 3 x 3 convolution weight gradient (tf32x3 path; the single-pass path is conv_dw_tma_kernel): conv_wp = W + 2 > 0,
    // g and x are zero-haloed (rows, 256) images, blockIdx.z = filter tap: x rows are shifted by the tap and the tap's
    // 256 columns of dw (N, ld_dw = 9 * 256) are written
    int conv_wp, ld_dw;
};

static __global__ void __launch_bounds__(kDwThreads, 1) linear_dw_tf32x3_kernel(const LinearDwParams p)
{
    constexpr int kABytes = 128 * 128;                 // 128 features x 128 B
    constexpr int kBBytes = kDwNB * 128;               // 272 features x 128 B
    constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    const int k0 = p.conv_wp ? 0 : blockIdx.z * 256;   // feature slice of x handled by this CTA
    const int dw_c0 = blockIdx.z * 256;                // its columns of dw
    const int ld_dw = p.conv_wp ? p.ld_dw : p.ldx;
    const int tz = static_cast<int>(blockIdx.z);
    const int xshift = p.conv_wp ? (tz / 3 - 1) * p.conv_wp + (tz % 3 - 1) : 0;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[kGemmStages], empty_bar[kGemmStages], acc_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * 128;
    const int r_begin = blockIdx.y * p.rows_per_chunk;
    const int r_end = min(p.M, r_begin + p.rows_per_chunk);
    const int kblocks = r_end > r_begin ? (r_end - r_begin + 31) / 32 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmStages; ++s) {
            mbar_init(&full_bar[s], 256);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base_slot, 512);
        // constant rows of B: row 256 = ones (hi only), rows 257..271 = 0, in both stages
        for (int s = 0; s < kGemmStages; ++s) {
            unsigned char *b_hi = smem + s * kStageBytes + 2 * kABytes;
            unsigned char *b_lo = b_hi + kBBytes;
            for (int e = lane; e < 16 * 8; e += 32) {              // 16 rows x 8 chunks
                const int row = 256 + e / 8, ch = e % 8;
                const float v = (row == 256) ? 1.f : 0.f;
                const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
                *reinterpret_cast<float4 *>(b_hi + off) = make_float4(v, v, v, v);
                *reinterpret_cast<float4 *>(b_lo + off) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    auto stage_ptr = [&](int s) { return smem + s * kStageBytes; };

    if (warp == 1) {
        if (lane == 0 && kblocks > 0) {
            // UMMA N is limited to 256: features 0..255 in one instruction, the ones/zero rows 256..271 (bias
            // gradient column) in a second N = 16 instruction on the same A operand
            constexpr uint32_t idesc = umma_idesc_tf32(128, 256);
            constexpr uint32_t idesc_b = umma_idesc_tf32(128, kDwNB - 256);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kGemmStages;
                mbar_wait_bounded(&full_bar[s], (kb / kGemmStages) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(stage_ptr(s));
                const uint32_t a_lo = a_hi + kABytes;
                const uint32_t b_hi = a_hi + 2 * kABytes;
                const uint32_t b_lo = b_hi + kBBytes;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t koff = k * 32;
                    const uint32_t acc = (kb | k) ? 1u : 0u;
                    const uint64_t da = umma_desc_k128(a_hi + koff), dbd = umma_desc_k128(b_hi + koff);
                    const uint64_t dones = umma_desc_k128(b_hi + 256 * 128 + koff);
                    umma_tf32(tmem_base, da, dbd, idesc, acc);
                    umma_tf32(tmem_base + 256, da, dones, idesc_b, acc);
                    if (p.split == 3) {
                        const uint64_t dal = umma_desc_k128(a_lo + koff);
                        umma_tf32(tmem_base, dal, dbd, idesc, 1u);
                        umma_tf32(tmem_base + 256, dal, dones, idesc_b, 1u);
                        umma_tf32(tmem_base, da, umma_desc_k128(b_lo + koff), idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&acc_bar);
        }
    } else if (warp >= 2) {
        // ---- transposing producers: 256 threads; a thread owns 3 features (one of G, two of X) x 4 row quads ----
        // item (a, b): feature slot a in {A: n0 + fl, B: k0 + fl, B: k0 + 128 + fl}, rows 4*(qh + 2b) .. +3 of the k-block.
        // lane <-> consecutive features: global loads are 128-byte coalesced, the swizzled 16-byte stores of a
        // quarter-warp hit 8 different bank groups.
        const int t = threadIdx.x - 64;
        const int fl = t & 127, qh = t >> 7;
        const bool a_ok = n0 + fl < p.N;
        const float *src[3] = {p.g + n0 + fl, p.x + k0 + fl, p.x + k0 + 128 + fl};
        const int ld[3] = {p.N, p.ldx, p.ldx};
        const uint32_t tile_off[3] = {static_cast<uint32_t>(fl) * 128u, 2u * kABytes + static_cast<uint32_t>(fl) * 128u,
                                      2u * kABytes + static_cast<uint32_t>(fl + 128) * 128u};
        const uint32_t lo_off[3] = {kABytes, kBBytes, kBBytes};
        auto load_items = [&](int kb, float (&v)[3][4][4]) {
            const int rb = r_begin + kb * 32 + qh * 4;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = rb + b * 8 + i;
                        const int rs = a > 0 ? r + xshift : r;        // conv: x row of this tap (halo rows are zero)
                        v[a][b][i] = ((a > 0 || a_ok) && r < r_end && rs >= 0 && rs < p.M)
                                         ? __ldg(src[a] + static_cast<size_t>(rs) * ld[a]) : 0.f;
                    }
        };
        float cur[3][4][4], nxt[3][4][4];
        if (kblocks > 0) load_items(0, cur);
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % kGemmStages;
            if (kb + 1 < kblocks) load_items(kb + 1, nxt);
            mbar_wait_bounded(&empty_bar[s], ((kb / kGemmStages) & 1) ^ 1);
            unsigned char *a_hi = stage_ptr(s);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int rq = qh + 2 * b;                                    // 16-byte chunk = row quad
                    unsigned char *hi_p = a_hi + tile_off[a] + ((rq ^ (fl & 7)) << 4);
                    float4 hi, lo;
                    hi.x = tf32_hi(cur[a][b][0]); lo.x = cur[a][b][0] - hi.x;
                    hi.y = tf32_hi(cur[a][b][1]); lo.y = cur[a][b][1] - hi.y;
                    hi.z = tf32_hi(cur[a][b][2]); lo.z = cur[a][b][2] - hi.z;
                    hi.w = tf32_hi(cur[a][b][3]); lo.w = cur[a][b][3] - hi.w;
                    *reinterpret_cast<float4 *>(hi_p) = hi;
                    *reinterpret_cast<float4 *>(hi_p + lo_off[a]) = lo;
                }
            fence_async_smem();
            mbar_arrive(&full_bar[s]);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int i = 0; i < 4; ++i) cur[a][b][i] = nxt[a][b][i];
        }
        // ---- epilogue (warps 2..5): TMEM lane = feature n, columns = k (0..255) and the bias column 256 ----
        if (warp < 6 && kblocks > 0) {
            mbar_wait_bounded(&acc_bar, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int n = n0 + q * 32 + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < 288; c0 += 32) {
                float acc[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, acc);
                if (n < p.N) {
                    if (c0 < 256) {
                        float *dst = p.dw + static_cast<size_t>(n) * ld_dw + dw_c0 + c0;
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float r4[4] = {acc[c], acc[c + 1], acc[c + 2], acc[c + 3]};
                            VecIO<float, 4>::red_add(dst + c, r4);
                        }
                    } else if (p.db && blockIdx.z == 0) {
                        const float r1[1] = {acc[0]};
                        VecIO<float, 1>::red_add(p.db + n, r1);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------
// Weight gradient with TMA-fed MN-major operands, single TF32 pass (linear layers and the 3x3 convolution):
//     dW[n, unit * 256 + c] = sum over rows r of G[r, n] * X[r + shift(unit), x_col0(unit) + c]
// linear layer: unit = 256-column slice of X (shift 0); 3x3 convolution: unit = filter tap, G and X zero-haloed images,
// shift = the tap's row shift, x_col0 = 0.  The reduction index is the ROW of both operands, i.e. both are MN-major for
// the MMA (features contiguous, reduction strided).  tcgen05 takes MN-major TF32 operands (instruction-descriptor bits
// 15 / 16), so the tiles go global -> shared by TMA (2-D boxes of 32 rows x 32 features) and shared -> tensor core by
// descriptor: no transposing producer threads (linear_dw_tf32x3_kernel spends its time there).  For 32-bit MN-major
// operands the only swizzled layout the MMA accepts is SWIZZLE_128B with 32-byte swizzle units
// (cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B = TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): canonical
// ((32, n), (4, k)) with feature-chunk stride LBO = 4096 B (one box) and 4-row stride SBO = 512 B; one K = 8 MMA reads
// two such 4-row atoms.
// grid = (units x 256-row slices of dW, row chunks); CTA = a 256 x 256 block of dW: two M = 128 halves x N = 256 -> 512
// TMEM columns.  Warp 0 TMA, warp 1 MMA, warps 2..5 epilogue (red.global.add of the partial sums; dW zeroed by the caller).
// ------------------------------------------------------------------------------------------------------
constexpr int kDwTmaStages = 3;
constexpr int kDwTmaThreads = 192;
constexpr int kDwTmaTile = 8 * 32 * 128;      // 8 feature chunks x 32 rows x 128 B = 32 KB per operand

struct DwTmaParams {
    float *dw;           // (N, ld_dw), accumulated into
    int ld_dw;
    int units;           // X column slices (linear) or filter taps (conv)
    int rows;
    int rows_per_chunk;  // multiple of 32
    int conv_wp;         // > 0: unit = tap of a 3x3 filter over zero-haloed images of row width conv_wp
};

__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);          // [0,14)  start address >> 4
    d |= static_cast<uint64_t>(4096 >> 4) << 16;                     // [16,30) leading byte offset: next 32-feature chunk
    d |= static_cast<uint64_t>(512 >> 4) << 32;                      // [32,46) stride byte offset: next 4 reduction rows
    d |= static_cast<uint64_t>(1) << 46;                             // [46,48) descriptor version (sm_100)
    d |= static_cast<uint64_t>(1) << 61;                             // [61,64) layout: SWIZZLE_128B_BASE32B
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int m, int n)
{
    return umma_idesc_tf32(m, n) | (1u << 15) | (1u << 16);          // A and B MN-major
}

static __global__ void __launch_bounds__(kDwTmaThreads, 1)
linear_dw_tma_kernel(const DwTmaParams p, const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x)
{
    constexpr int kStageBytes = 2 * kDwTmaTile;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[kDwTmaStages], empty_bar[kDwTmaStages], acc_bar;
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nslice = static_cast<int>(blockIdx.x) / p.units, unit = static_cast<int>(blockIdx.x) - nslice * p.units;
    const int shift = p.conv_wp ? (unit / 3 - 1) * p.conv_wp + (unit % 3 - 1) : 0;
    const int x_col0 = p.conv_wp ? 0 : unit * 256, g_col0 = nslice * 256;
    const int r_begin = blockIdx.y * p.rows_per_chunk;
    const int r_end = min(p.rows, r_begin + p.rows_per_chunk);
    const int kblocks = r_end > r_begin ? (r_end - r_begin + 31) / 32 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kDwTmaStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tm_g);
        tma_prefetch_desc(&tm_x);
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kDwTmaStages;
                mbar_wait_bounded(&empty_bar[s], ((kb / kDwTmaStages) & 1) ^ 1);
                uint64_t *bar = &full_bar[s];
                mbar_arrive_expect_tx(bar, kStageBytes);
                unsigned char *g_t = smem + s * kStageBytes, *x_t = g_t + kDwTmaTile;
                // chunks are multiples of 32 rows, so a box never reaches into the next chunk; past the last row of the
                // tensors (and before the first, for a negative tap shift) the TMA unit fills zeros
                const int r = r_begin + kb * 32;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    tma_load_2d(g_t + ch * 4096, &tm_g, g_col0 + ch * 32, r, bar);
                    tma_load_2d(x_t + ch * 4096, &tm_x, x_col0 + ch * 32, r + shift, bar);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && kblocks > 0) {
            constexpr uint32_t idesc = umma_idesc_tf32_mn(128, 256);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kDwTmaStages;
                mbar_wait_bounded(&full_bar[s], (kb / kDwTmaStages) & 1);
                tc_fence_after();
                const uint32_t g_t = smem_u32(smem + s * kStageBytes), x_t = g_t + kDwTmaTile;
#pragma unroll
                for (int k = 0; k < 4; ++k) {               // 8 reduction rows per MMA
                    const uint64_t db = umma_desc_mn128(x_t + k * 1024);
                    umma_tf32(tmem_base, umma_desc_mn128(g_t + k * 1024), db, idesc, (kb | k) ? 1u : 0u);
                    umma_tf32(tmem_base + 256, umma_desc_mn128(g_t + 4 * 4096 + k * 1024), db, idesc, (kb | k) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&acc_bar);
        }
    } else if (kblocks > 0) {
        // ---- epilogue: TMEM lane = output feature n (two halves), column = input feature c ----
        mbar_wait_bounded(&acc_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int n = g_col0 + half * 128 + q * 32 + lane;
            float *dst = p.dw + static_cast<size_t>(n) * p.ld_dw + unit * 256;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                float acc[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 256 + c0, acc);
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float r4[4] = {acc[c], acc[c + 1], acc[c + 2], acc[c + 3]};
                    VecIO<float, 4>::red_add(dst + c0 + c, r4);
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

constexpr int linear_dw_tma_smem_bytes() { return kDwTmaStages * 2 * kDwTmaTile + 1024; }

// db[c] += sum over this CTA's rows of g[r, c]  (bias gradient next to linear_dw_tma_kernel; db zeroed by the caller)
static __global__ void __launch_bounds__(256) column_sum_kernel(const float *__restrict__ g, float *__restrict__ db, int rows,
                                                                int cols, int rows_per_cta)
{
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    for (int c = blockIdx.y * 256 + threadIdx.x; c < cols; c += gridDim.y * 256) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int r = r0;
        for (; r + 3 < r1; r += 4) {
            a0 += __ldg(g + static_cast<size_t>(r) * cols + c);
            a1 += __ldg(g + static_cast<size_t>(r + 1) * cols + c);
            a2 += __ldg(g + static_cast<size_t>(r + 2) * cols + c);
            a3 += __ldg(g + static_cast<size_t>(r + 3) * cols + c);
        }
        for (; r < r1; ++r) a0 += __ldg(g + static_cast<size_t>(r) * cols + c);
        atomicAdd(db + c, (a0 + a1) + (a2 + a3));
    }
}

constexpr int linear_dw_smem_bytes() { return kGemmStages * (2 * 128 * 128 + 2 * kDwNB * 128) + 1024; }

template <int NT, int NH>
constexpr int linear_smem_bytes()
{
    return kGemmStages * (2 * kGemmBlockM * kGemmBlockK * 4 + 2 * NT * NH * kGemmBlockK * 4) + 1024;
}

// W (N, K) -> W_hi, W_lo (exact TF32 split); tiny, run once per weight version
// transpose != 0: w is (cols, rows) row-major and hi / lo receive its transpose (rows, cols)
static __global__ void split_tf32_kernel(const float *__restrict__ w, float *__restrict__ hi, float *__restrict__ lo, int n,
                                  int rows, int cols, int transpose)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float x = transpose ? w[(i % cols) * rows + i / cols] : w[i];
        const float h = tf32_hi(x);
        hi[i] = h;
        lo[i] = x - h;
    }
}

}  // namespace bm2f

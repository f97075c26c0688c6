// Stand-alone microbenchmarks that establish the machine limits the sampling kernels are judged
// against (SURVEY.md §8d asks for a measured L2-gather peak; MEASURED_PEAKS.json only has HBM copy):
//
//   gather  — every warp instruction fetches whole 128-byte lines at pseudo-random line indices
//             inside a buffer of a given size (22/44/88/352 MB = value of 1/2/4/16 COCO images; 96 KB
//             = L1-resident window).  Shapes: 32 lanes x 4 B (one line / instruction, the
//             reference's access shape), 8 lanes x 16 B (four lines / instruction, the fast path's
//             shape), 4 lanes x 32 B (eight lines / instruction, 256-bit loads).
//   red     — vector / scalar fp32 reductions (red.global.add) to pseudo-random lines: 32 x scalar
//             (one line / instruction, the reference's atomicAdd shape) and 8 x v4 (four lines).
//
// Prints one line per measurement:  name  buffer_MB  GB/s  (useful bytes = lines x 128 B).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o msda_microbench msda_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                        \
        }                                                                                   \
    } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// MODE 0: 32 lanes x 4 B -> 1 line / instr.   MODE 1: 8 lanes x 16 B -> 4 lines / instr.
// MODE 2: 4 lanes x 32 B -> 8 lines / instr (ld.global.nc.v8.f32).
// `window` > 0 confines each CTA's lines to a private window of that many lines (L1-resident test).
template <int MODE>
__global__ void __launch_bounds__(512) gather_kernel(const float *__restrict__ buf, uint32_t nlines, int iters,
                                                     uint32_t window, float *sink)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr int LPL = MODE == 0 ? 32 : (MODE == 1 ? 8 : 4);   // lanes per line
    const int grp = lane / LPL, sub = lane % LPL;
    const uint32_t wbase = window ? (uint32_t)(((uint64_t)blockIdx.x * window) % (nlines - window)) : 0u;
    const uint32_t range = window ? window : nlines;
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        const uint32_t h = mix((gwarp * 977u + i) * 64u + grp);
        const uint32_t line = wbase + (uint32_t)(((uint64_t)h * range) >> 32);
        const float *p = buf + (size_t)line * 32;
        if (MODE == 0) {
            acc += __ldg(p + sub);
        } else if (MODE == 1) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p) + sub);
            acc += v.x + v.y + v.z + v.w;
        } else {
            float a, b, c, d, e, f, g, hh;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(hh)
                         : "l"(p + sub * 8));
            acc += a + b + c + d + e + f + g + hh;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

// MODE 0: 32 lanes x scalar red (1 line / instr).  MODE 1: 8 lanes x red.v4 (4 lines / instr).
template <int MODE>
__global__ void __launch_bounds__(512) red_kernel(float *buf, uint32_t nlines, int iters, uint32_t window)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr int LPL = MODE == 0 ? 32 : 8;
    const int grp = lane / LPL, sub = lane % LPL;
    const uint32_t wbase = window ? (uint32_t)(((uint64_t)blockIdx.x * window) % (nlines - window)) : 0u;
    const uint32_t range = window ? window : nlines;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        const uint32_t h = mix((gwarp * 977u + i) * 64u + grp);
        const uint32_t line = wbase + (uint32_t)(((uint64_t)h * range) >> 32);
        float *p = buf + (size_t)line * 32;
        if (MODE == 0) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + sub), "f"(1.0f));
        } else {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p + sub * 4), "f"(1.0f));
        }
    }
}


// TMA bulk reduction: every warp stages CHUNK bytes of addends in shared memory and lets the
// TMA engine add them into global memory (cp.reduce.async.bulk ... .add.f32), bypassing the
// LSU RED path.  CHUNK = 128 (one line) or 512 (four consecutive lines).
template <int CHUNK>
__global__ void __launch_bounds__(512) tma_red_kernel(float *buf, uint32_t nlines, int iters, uint32_t window)
{
    __shared__ __align__(128) float stage[16][2][CHUNK / 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t wbase = window ? (uint32_t)(((uint64_t)blockIdx.x * window) % (nlines - window)) : 0u;
    const uint32_t range = (window ? window : nlines) - CHUNK / 128;
    for (int i = 0; i < iters; ++i) {
        float *s = stage[warp][i & 1];
        if (i >= 2) {   // the buffer written two iterations ago must have been read by the TMA engine
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
        for (int k = lane; k < CHUNK / 4; k += 32) s[k] = 1.0f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const uint32_t h = mix((gwarp * 977u + i) * 64u);
            const uint32_t line = wbase + (uint32_t)(((uint64_t)h * range) >> 32);
            float *g = buf + (size_t)line * 32;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g),
                         "r"((uint32_t)__cvta_generic_to_shared(s)), "n"(CHUNK)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Hybrid: every warp iteration issues one red.v4 (4 random lines through the LSU/L1TEX path) AND one
// 512-byte TMA bulk reduce (4 consecutive lines through the TMA engine).  If the two paths have separate
// per-SM limits and L2 has headroom, lines/s exceeds either path alone.
__global__ void __launch_bounds__(512) hybrid_red_kernel(float *buf, uint32_t nlines, int iters)
{
    __shared__ __align__(128) float stage[16][2][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int grp = lane >> 3, sub = lane & 7;
    for (int i = 0; i < iters; ++i) {
        float *s = stage[warp][i & 1];
        if (i >= 2) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
        reinterpret_cast<float4 *>(s)[lane] = make_float4(1.f, 1.f, 1.f, 1.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        const uint32_t h = mix((gwarp * 977u + i) * 64u + grp);
        const uint32_t line = (uint32_t)(((uint64_t)h * (nlines - 4)) >> 32);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(buf + (size_t)line * 32 + sub * 4), "f"(1.0f));
        if (lane == 0) {
            const uint32_t h2 = mix((gwarp * 977u + i) * 64u + 17u);
            const uint32_t l2 = (uint32_t)(((uint64_t)h2 * (nlines - 4)) >> 32);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(
                             buf + (size_t)l2 * 32),
                         "r"((uint32_t)__cvta_generic_to_shared(s))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float time_ms(F launch, int reps)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch();
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}


// Shared-memory integer reductions in the backward's access shape (fixed-point accumulation of the 32 x 32
// level on chip): the accumulator of one (image, head) is 1024 pixels x 32 int32 words = 128 KB.  A warp instruction
// = 4 point groups x 8 lanes; a group adds one corner line (32 words) at a pseudo-random pixel with 4 x
// red.shared.add.u32 per lane (there is no vector form).  SWZ rotates a pixel's channel groups by 8 words per
// (pixel & 3) so that the 4 groups of an instruction do not land on the same banks by construction.
template <int SWZ>
__global__ void __launch_bounds__(512) smem_red_kernel(int iters, uint32_t *sink)
{
    extern __shared__ uint32_t acc[];
    for (int i = threadIdx.x; i < 1024 * 32; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, grp = lane >> 3, l8 = lane & 7;
    uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 4 + grp + 12345u;
    for (int it = 0; it < iters; ++it) {
        seed = mix(seed + it);
        const uint32_t pix = seed & 1023u;
        const uint32_t rot = SWZ ? (pix & 3u) * 8u : 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t w = pix * 32u + ((l8 * 4u + c + rot) & 31u);
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(acc + w))),
                         "r"(seed >> 8)
                         : "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = acc[blockIdx.x & 1023];
}

// Shared-memory gather: every warp instruction reads four pseudo-random 128-byte rows of a 96 KB shared-memory window
// with 8 lanes x 16 B each (LDS.128) — the rate a value window staged in shared memory would be read at.
__global__ void __launch_bounds__(512) smem_gather_kernel(int iters, float *sink)
{
    extern __shared__ __align__(128) float win[];
    constexpr int kRows = 768;
    for (int i = threadIdx.x; i < kRows * 32; i += blockDim.x) win[i] = 1.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        const uint32_t h = mix((gwarp * 977u + i) * 64u + grp);
        const uint32_t row = (uint32_t)(((uint64_t)h * kRows) >> 32);
        const float4 v = *reinterpret_cast<const float4 *>(win + row * 32 + sub * 4);
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) *sink = acc;
}

// --quick: the few peaks bench.py divides by, measured on the GPU and in the run that reports them; one JSON line.
// L1-resident gather of lines that are `stride_lines` x 128 B apart (the value tensor's (N, S, M, D) layout puts the lines
// of ONE head 8 lines = 1024 B apart): does the tag stage keep its rate when a CTA's lines use every 8th line only?
__global__ void __launch_bounds__(512) strided_gather_kernel(const float *__restrict__ buf, int iters, int stride_lines,
                                                             uint32_t rows_log2, float *sink)
{
    const int sub = threadIdx.x & 7;
    const float *base = buf + (size_t)blockIdx.x * ((size_t)32 << rows_log2) * stride_lines;
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        h = h * 1664525u + 1013904223u;
        const uint32_t row = h >> (32 - rows_log2);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(base + (size_t)row * 32 * stride_lines) + sub);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

template <bool SHARED>
__global__ void __launch_bounds__(512) lean_gather_kernel(const float *__restrict__ buf, int iters, float *sink)
{
    extern __shared__ __align__(128) float win[];
    constexpr uint32_t kRows = 512;                 // 64 KB window
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    const float *base = buf + (size_t)blockIdx.x * kRows * 32;
    if (SHARED) {
        for (int i = threadIdx.x; i < (int)kRows * 32; i += blockDim.x) win[i] = base[i];
        __syncthreads();
    }
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        h = h * 1664525u + 1013904223u;
        const uint32_t row = h >> 23;
        float4 v;
        if (SHARED) v = *reinterpret_cast<const float4 *>(win + row * 32 + sub * 4);
        else v = __ldg(reinterpret_cast<const float4 *>(base + row * 32) + sub);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
    (void)grp;
}

// lean L2-resident gather / reduction: same LCG loop, lines drawn from the whole buffer
template <bool RED>
__global__ void __launch_bounds__(512) lean_l2_kernel(float *__restrict__ buf, uint32_t nlines, int iters, float *sink)
{
    const int sub = threadIdx.x & 7;
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        h = h * 1664525u + 1013904223u;
        const uint32_t line = (uint32_t)(((uint64_t)h * nlines) >> 32);
        float *p = buf + (size_t)line * 32 + sub * 4;
        if (RED) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(0.f) : "memory");
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (!RED && acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

int lean_main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float *buf, *sink;
    CK(cudaMalloc(&buf, (size_t)sms * 4 * 65536));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, (size_t)sms * 4 * 65536));
    CK(cudaFuncSetAttribute(lean_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int iters = 8192;
    for (int ctas = 1; ctas <= 2; ++ctas)
        for (int threads = 256; threads <= 512; threads *= 2) {
            const double lines = (double)sms * ctas * (threads / 32) * iters * 4;
            const float t_l1 = time_ms([&] { lean_gather_kernel<false><<<sms * ctas, threads>>>(buf, iters, sink); }, 5);
            const float t_sm = time_ms([&] { lean_gather_kernel<true><<<sms * ctas, threads, 65536>>>(buf, iters, sink); }, 5);
            printf("lean gather 8x16B, 64 KB window, %d CTA(s)/SM x %d threads:  L1 (ld.global.nc) %8.1f GB/s = %.3f lines/clk/SM"
                   "   shared (ld.shared) %8.1f GB/s = %.3f lines/clk/SM\n", ctas, threads, lines * 128 / t_l1 * 1e-6,
                   lines / (t_l1 * 1e-3) / sms / (prop.clockRate * 1e3), lines * 128 / t_sm * 1e-6,
                   lines / (t_sm * 1e-3) / sms / (prop.clockRate * 1e3));
        }
    float *big;
    CK(cudaMalloc(&big, 352ull << 20));
    CK(cudaMemset(big, 0, 352ull << 20));
    // one head's lines: 512 (64 KB) or 256 (32 KB) or 128 (16 KB) lines per CTA at a stride of 1 / 2 / 4 / 8 lines
    for (uint32_t rl : {9u, 8u, 7u})
        for (int stride : {1, 2, 4, 8}) {
            const double lines = (double)sms * 16 * iters * 4;
            const float tt = time_ms([&] { strided_gather_kernel<<<sms, 512>>>(big, iters, stride, rl, sink); }, 5);
            printf("L1 gather 8x16B, %3u lines per CTA at a stride of %d lines (%4d B), 1 CTA/SM x 512 threads: %8.1f GB/s = %.3f lines/clk/SM\n",
                   1u << rl, stride, stride * 128, lines * 128 / tt * 1e-6, lines / (tt * 1e-3) / sms / (prop.clockRate * 1e3));
        }
    for (double mb : {44.0, 352.0}) {
        const uint32_t nl = (uint32_t)(mb * 1048576.0 / 128);
        for (int ctas = 1; ctas <= 4; ctas *= 2) {
            const int it2 = 1024;
            const double lines = (double)sms * ctas * 16 * it2 * 4;
            const float tg = time_ms([&] { lean_l2_kernel<false><<<sms * ctas, 512>>>(big, nl, it2, sink); }, 5);
            const float tr = time_ms([&] { lean_l2_kernel<true><<<sms * ctas, 512>>>(big, nl, it2, sink); }, 5);
            printf("lean 8x16B over %5.0f MB, %d CTA(s)/SM x 512 threads:  gather %8.1f GB/s   red.v4.f32 %8.1f GB/s\n", mb, ctas,
                   lines * 128 / tg * 1e-6, lines * 128 / tr * 1e-6);
        }
    }
    return 0;
}

// --quick: the few peaks bench.py divides by, measured on the GPU and in the run that reports them; one JSON line.
// Lean loops (an LCG step, one load / reduction of 8 lanes x 16 B x 4 lines, four adds): the round-1 kernels spent ~18
// instructions per load on index hashing and were issue-bound at roughly half of what L1 delivers.
int quick_main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const size_t bytes_buf = 512ull << 20;
    float *buf, *sink;
    CK(cudaMalloc(&buf, bytes_buf));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, bytes_buf));
    const uint32_t n44 = (uint32_t)(44.0 * 1048576.0 / 128);
    CK(cudaFuncSetAttribute(lean_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int it1 = 8192, it2 = 1024;
    const double l1_lines = (double)sms * 2 * 16 * it1 * 4, l2_lines = (double)sms * 4 * 16 * it2 * 4;
    const double gl1 = l1_lines * 128 / time_ms([&] { lean_gather_kernel<false><<<sms * 2, 512>>>(buf, it1, sink); }, 5) * 1e-6;
    const double lds = l1_lines * 128 / time_ms([&] { lean_gather_kernel<true><<<sms * 2, 512, 65536>>>(buf, it1, sink); }, 5) * 1e-6;
    const double gl2 = l2_lines * 128 / time_ms([&] { lean_l2_kernel<false><<<sms * 4, 512>>>(buf, n44, it2, sink); }, 5) * 1e-6;
    const double red = l2_lines * 128 / time_ms([&] { lean_l2_kernel<true><<<sms * 4, 512>>>(buf, n44, it2, sink); }, 5) * 1e-6;
    const size_t n = bytes_buf / 2;
    const double copy = 2.0 * n / time_ms([&] { CK(cudaMemcpyAsync(buf, (char *)buf + n, n, cudaMemcpyDeviceToDevice)); }, 5) * 1e-6;
    printf("{\"device\": \"%s\", \"sms\": %d, \"unit\": \"GB/s of 128-B lines\", \"gather_8x16B_l2_44MB\": %.1f, "
           "\"gather_8x16B_l1_64KB\": %.1f, \"red_8xv4f32_l2_44MB\": %.1f, \"lds_8x16B_smem_64KB\": %.1f, \"d2d_copy_rw\": %.1f}\n",
           prop.name, sms, gl2, gl1, red, lds, copy);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 1 && !strcmp(argv[1], "--quick")) return quick_main();
    if (argc > 1 && !strcmp(argv[1], "--lean")) return lean_main();
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("# device %s  SMs %d  L2 %.1f MB  smem/SM %zu KB\n", prop.name, prop.multiProcessorCount,
           prop.l2CacheSize / 1048576.0, prop.sharedMemPerMultiprocessor / 1024);
    const size_t max_bytes = 704ull << 20;
    float *buf, *sink;
    CK(cudaMalloc(&buf, max_bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, max_bytes));
    const int sms = prop.multiProcessorCount;
    const int threads = 512;
    const int grid = sms * 4;                      // 64 warps / SM
    const int warps = grid * threads / 32;
    const int iters = 512;

    const double mbs[] = {22, 44, 88, 352, 704};
    const char *gname[] = {"gather_32x4B_1line", "gather_8x16B_4lines", "gather_4x32B_8lines"};
    const int lines_per_instr[] = {1, 4, 8};
    for (int mode = 0; mode < 3; ++mode) {
        for (double mb : mbs) {
            const uint32_t nlines = (uint32_t)(mb * 1048576.0 / 128);
            auto launch = [&] {
                if (mode == 0) gather_kernel<0><<<grid, threads>>>(buf, nlines, iters, 0, sink);
                if (mode == 1) gather_kernel<1><<<grid, threads>>>(buf, nlines, iters, 0, sink);
                if (mode == 2) gather_kernel<2><<<grid, threads>>>(buf, nlines, iters, 0, sink);
            };
            const float ms = time_ms(launch, 5);
            const double bytes = (double)warps * iters * lines_per_instr[mode] * 128.0;
            printf("%-24s %7.1f MB  %9.1f GB/s\n", gname[mode], mb, bytes / ms * 1e-6);
        }
        // L1-resident: each CTA re-reads a private 96 KB window (768 lines); 1 CTA per SM
        {
            const uint32_t nlines = (uint32_t)(352.0 * 1048576.0 / 128);
            const int g1 = sms, it1 = 4096;
            auto launch = [&] {
                if (mode == 0) gather_kernel<0><<<g1, threads>>>(buf, nlines, it1, 768, sink);
                if (mode == 1) gather_kernel<1><<<g1, threads>>>(buf, nlines, it1, 768, sink);
                if (mode == 2) gather_kernel<2><<<g1, threads>>>(buf, nlines, it1, 768, sink);
            };
            const float ms = time_ms(launch, 5);
            const double bytes = (double)(g1 * threads / 32) * it1 * lines_per_instr[mode] * 128.0;
            printf("%-24s %7s     %9.1f GB/s  (L1-resident 96 KB window per SM)\n", gname[mode], "L1", bytes / ms * 1e-6);
        }
    }
    const char *rname[] = {"red_32xf32_1line", "red_8xv4f32_4lines"};
    const int rl[] = {1, 4};
    for (int mode = 0; mode < 2; ++mode) {
        for (double mb : mbs) {
            const uint32_t nlines = (uint32_t)(mb * 1048576.0 / 128);
            const int it = 128;
            auto launch = [&] {
                if (mode == 0) red_kernel<0><<<grid, threads>>>(buf, nlines, it, 0);
                if (mode == 1) red_kernel<1><<<grid, threads>>>(buf, nlines, it, 0);
            };
            const float ms = time_ms(launch, 5);
            const double bytes = (double)warps * it * rl[mode] * 128.0;
            printf("%-24s %7.1f MB  %9.1f GB/s\n", rname[mode], mb, bytes / ms * 1e-6);
        }
        {   // localised: each CTA reduces into a private 96 KB window (what a strip-walking job does)
            const uint32_t nlines = (uint32_t)(352.0 * 1048576.0 / 128);
            const int it = 256;
            auto launch = [&] {
                if (mode == 0) red_kernel<0><<<grid, threads>>>(buf, nlines, it, 768);
                if (mode == 1) red_kernel<1><<<grid, threads>>>(buf, nlines, it, 768);
            };
            const float ms = time_ms(launch, 5);
            const double bytes = (double)warps * it * rl[mode] * 128.0;
            printf("%-24s %7s     %9.1f GB/s  (96 KB window per CTA)\n", rname[mode], "win", bytes / ms * 1e-6);
        }
    }

    // ---- is the vector-RED rate an SM-side or an L2-side limit?  scale SMs and warps per SM ----
    {
        const uint32_t nlines = (uint32_t)(44.0 * 1048576.0 / 128);
        const int it = 256;
        const int grids[] = {sms / 4, sms / 2, sms, sms * 2, sms * 4};
        for (int g : grids) {
            for (int th : {128, 512}) {
                auto launch = [&] { red_kernel<1><<<g, th>>>(buf, nlines, it, 0); };
                const float ms = time_ms(launch, 3);
                const double bytes = (double)(g * th / 32) * it * 4 * 128.0;
                printf("red_v4_scaling  ctas %4d x %3d thr  %9.1f GB/s\n", g, th, bytes / ms * 1e-6);
            }
        }
        for (int g : grids) {
            auto launch = [&] { gather_kernel<1><<<g, 512>>>(buf, nlines, 512, 0, sink); };
            const float ms = time_ms(launch, 3);
            const double bytes = (double)(g * 512 / 32) * 512 * 4 * 128.0;
            printf("gather_v4_scaling ctas %4d x 512 thr  %9.1f GB/s\n", g, bytes / ms * 1e-6);
        }
    }
    // ---- TMA bulk reduce (cp.reduce.async.bulk .add.f32) ----
    for (double mb : {44.0, 352.0}) {
        const uint32_t nlines = (uint32_t)(mb * 1048576.0 / 128);
        const int it = 256;
        {
            auto launch = [&] { tma_red_kernel<128><<<grid, threads>>>(buf, nlines, it, 0); };
            const float ms = time_ms(launch, 3);
            printf("%-24s %7.1f MB  %9.1f GB/s\n", "tma_reduce_128B", mb, (double)warps * it * 128.0 / ms * 1e-6);
        }
        {
            auto launch = [&] { tma_red_kernel<512><<<grid, threads>>>(buf, nlines, it, 0); };
            const float ms = time_ms(launch, 3);
            printf("%-24s %7.1f MB  %9.1f GB/s\n", "tma_reduce_512B", mb, (double)warps * it * 512.0 / ms * 1e-6);
        }
    }
    {   // TMA reduce and LSU vector REDs at the same time (two streams): do the two paths add up?
        const uint32_t nlines = (uint32_t)(44.0 * 1048576.0 / 128);
        cudaStream_t s1, s2;
        CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
        const int it = 256;
        auto launch = [&] {
            tma_red_kernel<512><<<grid / 2, threads, 0, s1>>>(buf, nlines, it, 0);
            red_kernel<1><<<grid / 2, threads, 0, s2>>>(buf, nlines, it, 0);
            CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2));
        };
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        const double bytes = (double)(warps / 2) * it * 512.0 * 2;
        printf("%-24s %7.1f MB  %9.1f GB/s  (half grid TMA-reduce + half grid red.v4 concurrently)\n", "tma+lsu_red", 44.0,
               bytes / ms * 1e-6);
    }

    {   // hybrid LSU red.v4 + TMA bulk reduce from the same warps
        const uint32_t nlines = (uint32_t)(44.0 * 1048576.0 / 128);
        const int it = 256;
        for (int g : {sms, sms * 2, sms * 4}) {
            auto launch = [&] { hybrid_red_kernel<<<g, threads>>>(buf, nlines, it); };
            const float ms = time_ms(launch, 3);
            printf("hybrid_red_v4+tma512  ctas %4d  %9.1f GB/s\n", g, (double)(g * threads / 32) * it * 8 * 128.0 / ms * 1e-6);
        }
    }
    // shared-memory integer reductions, backward access shape (one 128-KB accumulator per CTA)
    {
        const int it = 4000, threads = 512;
        CK(cudaFuncSetAttribute(smem_red_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        CK(cudaFuncSetAttribute(smem_red_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        uint32_t *isink;
        CK(cudaMalloc(&isink, 4096));
        for (int swz = 0; swz < 2; ++swz) {
            auto launch = [&] {
                if (swz) smem_red_kernel<1><<<sms, threads, 128 * 1024>>>(it, isink);
                else smem_red_kernel<0><<<sms, threads, 128 * 1024>>>(it, isink);
                CK(cudaGetLastError());
            };
            const float ms = time_ms(launch, 3);
            const double lines = (double)sms * (threads / 8) * it;      // corner lines (32 words = 128 B) added
            printf("smem_red_u32 %s  %9.1f GB/s of 128-B corner lines = %.2f cycles per line per SM at 1.9 GHz\n",
                   swz ? "swizzled" : "plain   ", lines * 128.0 / ms * 1e-6, 1.9e9 * ms * 1e-3 * sms / lines);
        }
        CK(cudaFree(isink));
    }
    // plain streaming copy for reference (same denominator as MEASURED_PEAKS.json hbm_gbs)
    {
        const size_t n = 512ull << 20;
        float *dst;
        CK(cudaMalloc(&dst, n));
        auto launch = [&] { CK(cudaMemcpyAsync(dst, buf, n, cudaMemcpyDeviceToDevice)); };
        const float ms = time_ms(launch, 5);
        printf("%-24s %7.1f MB  %9.1f GB/s  (read+write)\n", "d2d_copy", n / 1048576.0, 2.0 * n / ms * 1e-6);
        CK(cudaFree(dst));
    }
    CK(cudaFree(buf));
    CK(cudaFree(sink));
    return 0;
}

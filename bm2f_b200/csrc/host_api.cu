// Host-buffer entry of the C ABI: chunked, pipelined H2D -> kernels -> D2H around bm2f_msda_forward / backward.
#include "api_common.cuh"
#include "msda_common.cuh"

#include <mutex>

using namespace bm2f;
using namespace bm2f::host;

extern "C" {

// ---------------------------------------------------------------------------------------------
// Host-buffer entry: chunked, double-buffered H2D -> kernels -> D2H.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kHostSlots = 3;   // chunks in flight: one uploading, one computing / downloading, one draining
constexpr int kMaxDevices = 64;
// One state per device ordinal, each behind its own lock: two host threads driving two GPUs do not serialise, and
// nothing of one device is ever freed or re-created because another device called in.
struct HostPath {
    std::mutex mu;
    bool live = false;
    cudaStream_t streams[kHostSlots] = {};
    void *ws[kHostSlots] = {};
    size_t ws_bytes = 0;
    int64_t *tabs = nullptr;  // shapes (2L) + start (L)
} g_host[kMaxDevices];

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// current device must be the state's device; the state's lock is held by the caller
void release_locked(HostPath &h)
{
    for (int i = 0; i < kHostSlots; ++i) {
        if (h.streams[i]) { cudaStreamSynchronize(h.streams[i]); cudaStreamDestroy(h.streams[i]); }
        if (h.ws[i]) cudaFree(h.ws[i]);
        h.streams[i] = nullptr;
        h.ws[i] = nullptr;
    }
    if (h.tabs) cudaFree(h.tabs);
    h.tabs = nullptr;
    h.ws_bytes = 0;
    h.live = false;
}

// asynchronous copies into / out of caller-owned host buffers must not outlive an error return
int drain_and_return(HostPath &h, int rc)
{
    for (int i = 0; i < kHostSlots; ++i)
        if (h.streams[i]) cudaStreamSynchronize(h.streams[i]);
    return rc;
}
}  // namespace

int bm2f_msda_release_host_workspace(void)
{
    int cur = 0;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    for (int dev = 0; dev < kMaxDevices; ++dev) {
        HostPath &h = g_host[dev];
        std::lock_guard<std::mutex> lk(h.mu);
        if (!h.live) continue;
        if ((e = cudaSetDevice(dev)) != cudaSuccess) { cudaSetDevice(cur); return cuda_fail(e, "cudaSetDevice"); }
        release_locked(h);
    }
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

    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaGetDevice");
    if (dev < 0 || dev >= kMaxDevices) return fail(BM2F_ERR_UNSUPPORTED, "device ordinal %d >= %d", dev, kMaxDevices);
    HostPath &h = g_host[dev];
    std::lock_guard<std::mutex> lk(h.mu);
    if (!h.live || h.ws_bytes < slot_bytes) {
        for (int i = 0; i < kHostSlots; ++i) {
            if (h.ws[i]) { cudaStreamSynchronize(h.streams[i]); cudaFree(h.ws[i]); }
            h.ws[i] = nullptr;
            if (!h.streams[i] && (ce = cudaStreamCreateWithFlags(&h.streams[i], cudaStreamNonBlocking)) != cudaSuccess)
                return cuda_fail(ce, "cudaStreamCreate");
            if ((ce = cudaMalloc(&h.ws[i], slot_bytes)) != cudaSuccess) {
                h.ws_bytes = 0;
                return cuda_fail(ce, "cudaMalloc(host-path workspace)");
            }
        }
        if (!h.tabs && (ce = cudaMalloc(reinterpret_cast<void **>(&h.tabs), sizeof(int64_t) * 3 * kMaxLevels)) != cudaSuccess)
            return cuda_fail(ce, "cudaMalloc(level tables)");
        h.ws_bytes = slot_bytes;
        h.live = true;
    }
    cudaStream_t s0 = h.streams[0];
    if ((ce = cudaMemcpyAsync(h.tabs, spatial_shapes_host, sizeof(int64_t) * 2 * d.L, cudaMemcpyHostToDevice,
                              s0)) != cudaSuccess)
        return cuda_fail(ce, "H2D spatial_shapes");
    if ((ce = cudaMemcpyAsync(h.tabs + 2 * kMaxLevels, level_start_index_host, sizeof(int64_t) * d.L,
                              cudaMemcpyHostToDevice, s0)) != cudaSuccess)
        return cuda_fail(ce, "H2D level_start_index");
    if ((ce = cudaStreamSynchronize(s0)) != cudaSuccess) return cuda_fail(ce, "sync level tables");
    const int64_t *d_shapes = h.tabs, *d_start = h.tabs + 2 * kMaxLevels;

    auto hp = [](const void *base, size_t off) { return static_cast<const char *>(base) + off; };
    auto hpw = [](void *base, size_t off) { return static_cast<char *>(base) + off; };

    // Per chunk, on its slot's stream: ALL uploads first (grad_output included), then both kernels, then all downloads.
    // Copy engines serve requests in issue order, so an upload queued between a chunk's kernels and its downloads would
    // hold back the next chunk's uploads (head-of-line blocking) — with this order the H2D engine, the SMs and the D2H
    // engine each work on a different chunk.
    int slot = 0;
    for (int b0 = 0; b0 < d.N; b0 += chunk, slot = (slot + 1) % kHostSlots) {
        const int nb = (d.N - b0 < chunk) ? d.N - b0 : chunk;
        cudaStream_t st = h.streams[slot];
        char *w = static_cast<char *>(h.ws[slot]);
        auto take = [&](size_t per_img) { char *r = w; w += static_cast<size_t>(chunk) * align256(per_img); return r; };
        char *dv = take(v_img), *dl = take(l_img), *da = take(a_img), *dout = take(o_img);
        char *dgo = nullptr, *dgv = nullptr, *dgl = nullptr, *dga = nullptr;
        if (bwd) { dgo = take(o_img); dgv = take(gv_img); dgl = take(l_img); dga = take(a_img); }

#define BM2F_CP(dst, src, bytes, kind, what)                                                           \
    if ((ce = cudaMemcpyAsync(dst, src, bytes, kind, st)) != cudaSuccess) return drain_and_return(h, cuda_fail(ce, what));
        BM2F_CP(dv, hp(value_host, b0 * v_img), nb * v_img, cudaMemcpyHostToDevice, "H2D value")
        BM2F_CP(dl, hp(sampling_loc_host, b0 * l_img), nb * l_img, cudaMemcpyHostToDevice, "H2D sampling_loc")
        BM2F_CP(da, hp(attn_weight_host, b0 * a_img), nb * a_img, cudaMemcpyHostToDevice, "H2D attn_weight")
        if (bwd) BM2F_CP(dgo, hp(grad_output_host, b0 * o_img), nb * o_img, cudaMemcpyHostToDevice, "H2D grad_output")
        rc = bm2f_msda_forward(dv, d_shapes, d_start, dl, da, dout, nb, d.S, d.M, d.D, d.L, d.Lq, d.P, dtype, tuning,
                               st);
        if (rc) return drain_and_return(h, rc);
        if (bwd) {
            rc = bm2f_msda_backward(dv, d_shapes, d_start, dl, da, dgo, dgv, dgl, dga, nb, d.S, d.M, d.D, d.L, d.Lq,
                                    d.P, dtype, tuning, st);
            if (rc) return drain_and_return(h, rc);
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
        if ((ce = cudaStreamSynchronize(h.streams[i])) != cudaSuccess) return drain_and_return(h, cuda_fail(ce, "host-path sync"));
    return BM2F_OK;
}

}  // extern "C"

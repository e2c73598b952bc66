// ctx.cu -- context life cycle, error strings, staging of host is_mat buffers.
#include "common.cuh"
#include "hostpool.h"

#include <algorithm>
#include <cstdlib>
#include <thread>
#include <sched.h>

namespace is {

int fail(is_ctx* ctx, int status, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return status;
}

static cudaEvent_t pool_event(is_ctx* ctx) {
    if (!ctx->kpool.empty()) { cudaEvent_t e = ctx->kpool.back(); ctx->kpool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

cudaEvent_t ktiming_begin(is_ctx* ctx, const char* name) {
    is_ctx::KRec r;
    r.name = name;
    r.e0 = pool_event(ctx);
    r.e1 = pool_event(ctx);
    r.bytes = ctx->next_bytes;
    ctx->next_bytes = 0;
    cudaEventRecord(r.e0, ctx->stream);
    ctx->krecs.push_back(r);
    return r.e1;
}

constexpr size_t BLOCK_CACHE_LIMIT = (size_t)6 << 30;   // per context

int DevBuf::alloc(is_ctx* c, size_t n) {
    release();
    ctx = c;
    if (n == 0) n = 16;
    n = align_up(n, 256);
    auto it = c->block_cache.lower_bound(n);
    if (it != c->block_cache.end() && it->first <= n + n / 4 + 4096) {   // close enough in size: reuse
        p = it->second;
        bytes = it->first;
        c->block_cache_bytes -= bytes;
        c->block_cache.erase(it);
        return IS_OK;
    }
    cudaError_t e = cudaMallocAsync(&p, n, c->stream);
    if (e != cudaSuccess) {
        // Out of memory: give back the blocks parked in this context and in its idle children (a child is only ever used from the
        // thread that runs this call or from workers this call has joined), let the pool return what it holds to the device
        // (the release threshold is "never" in steady state), and try once more.
        cudaGetLastError();
        auto flush = [](is_ctx* x) {
            for (auto& kv : x->block_cache) cudaFreeAsync(kv.second, x->stream);
            x->block_cache.clear();
            x->block_cache_bytes = 0;
            cudaStreamSynchronize(x->stream);
        };
        flush(c);
        for (is_ctx* ch : c->children) if (ch) flush(ch);
        if (c->pool) cudaMemPoolTrimTo(c->pool, 0);
        cudaGetLastError();
        e = cudaMallocAsync(&p, n, c->stream);
    }
    if (e != cudaSuccess) {
        p = nullptr;
        return fail(c, e == cudaErrorMemoryAllocation ? IS_ERR_NO_MEM : IS_ERR_CUDA, "cudaMallocAsync(%zu): %s", n,
                    cudaGetErrorString(e));
    }
    bytes = n;
    return IS_OK;
}

void DevBuf::release() {
    if (p && ctx) {
        if (ctx->block_cache_bytes + bytes <= BLOCK_CACHE_LIMIT) {
            ctx->block_cache.emplace(bytes, p);
            ctx->block_cache_bytes += bytes;
        } else {
            cudaFreeAsync(p, ctx->stream);
        }
    }
    p = nullptr;
    bytes = 0;
}

// Worker contexts: own stream + staging, sharing the parent's device and memory pool.  Indices 0..7 serve the
// concurrent seam pairs, SIDE_PYRAMID / SIDE_COPY the pipeline's side streams.
int child_ctx(is_ctx* parent, size_t k, is_ctx** out) {
    while (parent->children.size() <= k) {
        is_ctx* c = new is_ctx();
        c->device = parent->device;
        // the seam workers' short kernels go ahead of the bulk work of the side streams
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const int prio = parent->children.size() < (size_t)SIDE_PYRAMID ? hi : lo;
        if (cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, prio) != cudaSuccess) { delete c; return fail(parent, IS_ERR_CUDA, "cudaStreamCreate failed"); }
        c->stream = c->own_stream;
        c->pool = parent->pool;
        for (int e = 0; e < 5; ++e) cudaEventCreateWithFlags(&c->ev[e], cudaEventDisableTiming);
        parent->children.push_back(c);
    }
    *out = parent->children[k];
    return IS_OK;
}

// `to` waits for everything queued on `from` so far.  The events come from a fixed ring: an event that has been recorded and
// waited on may be re-recorded (the wait already captured the earlier record), so a long-running job never grows the pool.
int stream_after(is_ctx* ctx, cudaStream_t to, cudaStream_t from) {
    constexpr size_t RING = 64;
    if (ctx->sync_events.size() < RING) {
        cudaEvent_t e = nullptr;
        IS_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->sync_events.push_back(e);
        ctx->sync_next = ctx->sync_events.size() - 1;
    }
    cudaEvent_t e = ctx->sync_events[ctx->sync_next % ctx->sync_events.size()];
    ctx->sync_next = (ctx->sync_next + 1) % RING;
    IS_CUDA(ctx, cudaEventRecord(e, from));
    IS_CUDA(ctx, cudaStreamWaitEvent(to, e, 0));
    return IS_OK;
}

// account a worker's launches / timing records / error to the parent
void merge_child(is_ctx* parent, is_ctx* child) {
    parent->launches += child->launches;
    child->launches = 0;
    for (auto& r : child->krecs) parent->krecs.push_back(r);
    child->krecs.clear();
    if (!child->last_error.empty() && parent->last_error.empty()) parent->last_error = child->last_error;
}

int check_mat(is_ctx* ctx, const is_mat* m, const char* name) {
    if (!m) return fail(ctx, IS_ERR_BAD_ARG, "%s: null is_mat", name);
    if (!m->data || m->rows <= 0 || m->cols <= 0 || m->channels <= 0 || depth_bytes(m->depth) == 0)
        return fail(ctx, IS_ERR_BAD_ARG, "%s: empty or malformed is_mat (rows=%d cols=%d ch=%d depth=%d)", name, m->rows,
                    m->cols, m->channels, m->depth);
    if (m->step < (size_t)m->cols * m->channels * depth_bytes(m->depth))
        return fail(ctx, IS_ERR_BAD_ARG, "%s: step %zu smaller than a row", name, m->step);
    if (m->device >= 0 && m->device != ctx->device)
        return fail(ctx, IS_ERR_BAD_ARG, "%s: buffer lives on device %d, context on device %d", name, m->device, ctx->device);
    return IS_OK;
}

int alloc_mat(is_ctx* ctx, int rows, int cols, int channels, int depth, DevMat* out, size_t align) {
    out->rows = rows; out->cols = cols; out->channels = channels; out->depth = depth;
    out->step = align_up((size_t)cols * channels * depth_bytes(depth), align);
    IS_TRY(out->owned.alloc(ctx, out->step * (size_t)rows + 256));   // tail padding: vector loads may over-read a row end
    out->data = out->owned.p;
    out->host = nullptr;
    return IS_OK;
}

int stage_in(is_ctx* ctx, const is_mat* m, DevMat* out) {
    if (m->device >= 0) {
        out->data = m->data; out->rows = m->rows; out->cols = m->cols; out->channels = m->channels;
        out->depth = m->depth; out->step = m->step; out->host = nullptr;
        return IS_OK;
    }
    IS_TRY(alloc_mat(ctx, m->rows, m->cols, m->channels, m->depth, out));
    IS_CUDA(ctx, cudaMemcpy2DAsync(out->data, out->step, m->data, m->step, out->row_bytes(), m->rows,
                                   cudaMemcpyHostToDevice, ctx->stream));
    return IS_OK;
}

int stage_out(is_ctx* ctx, is_mat* m, DevMat* out, bool preload) {
    if (m->device >= 0) {
        out->data = m->data; out->rows = m->rows; out->cols = m->cols; out->channels = m->channels;
        out->depth = m->depth; out->step = m->step; out->host = nullptr;
        return IS_OK;
    }
    IS_TRY(alloc_mat(ctx, m->rows, m->cols, m->channels, m->depth, out));
    out->host = m;
    if (preload)
        IS_CUDA(ctx, cudaMemcpy2DAsync(out->data, out->step, m->data, m->step, out->row_bytes(), m->rows,
                                       cudaMemcpyHostToDevice, ctx->stream));
    return IS_OK;
}

int commit(is_ctx* ctx, DevMat* m) {
    if (!m->host) return IS_OK;
    IS_CUDA(ctx, cudaMemcpy2DAsync(m->host->data, m->host->step, m->data, m->step, m->row_bytes(), m->rows,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IS_OK;
}

int ensure_pinned(is_ctx* ctx, size_t bytes) {
    if (ctx->pinned_bytes >= bytes) return IS_OK;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    size_t n = align_up(bytes, 1 << 20);
    IS_CUDA(ctx, cudaMallocHost(&ctx->pinned, n));
    ctx->pinned_bytes = n;
    return IS_OK;
}

int pinned_alloc(is_ctx* ctx, size_t bytes, void** out) {
    bytes = align_up(bytes, 256);
    if (ctx->pinned_bytes < bytes || !ctx->pinned) {
        IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        IS_TRY(ensure_pinned(ctx, bytes > (size_t)(8 << 20) ? bytes : (size_t)(8 << 20)));
        ctx->pinned_off = 0;
    }
    if (ctx->pinned_off + bytes > ctx->pinned_bytes) {
        IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // earlier uploads have drained
        ctx->pinned_off = 0;
    }
    *out = (char*)ctx->pinned + ctx->pinned_off;
    ctx->pinned_off += bytes;
    return IS_OK;
}

// Small control transfers do not go through the copy engines: a panorama's bulk copies (hundreds of MB, one DMA queue per
// direction) would hold them up for milliseconds when several panoramas are in flight on one device -- the seam stage of one
// panorama waited for the download of another.  The pinned staging buffers are mapped into the device's address space
// (unified addressing), so a kernel moves the bytes with ordinary loads and stores over the host link instead.
__global__ void __launch_bounds__(256) k_copy_small(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, size_t bytes) {
    const size_t n16 = bytes >> 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = t; i < n16; i += stride) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (size_t i = (n16 << 4) + t; i < bytes; i += stride) dst[i] = src[i];
}

static bool sm_copy_ok(const void* a, const void* b) {
    static const bool off = getenv("IS_COPY_ENGINE_SMALL") != nullptr;            // tuning knob: control transfers as cudaMemcpyAsync again
    return !off && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

int copy_small(is_ctx* ctx, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    if (!sm_copy_ok(dst, src)) {
        IS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, kind, ctx->stream));
        return IS_OK;
    }
    const int blocks = (int)std::min<size_t>(296, (bytes / 16 + 255) / 256 + 1);
    k_copy_small<<<blocks, 256, 0, ctx->stream>>>(static_cast<const unsigned char*>(src), static_cast<unsigned char*>(dst), bytes);
    ctx->launches++;
    IS_CUDA(ctx, cudaGetLastError());
    return IS_OK;
}

int upload(is_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return IS_OK;
    void* stage = nullptr;
    IS_TRY(pinned_alloc(ctx, bytes, &stage));
    std::memcpy(stage, src, bytes);
    return copy_small(ctx, dst, stage, bytes, cudaMemcpyHostToDevice);
}

int download(is_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return IS_OK;
    if (ctx->pinned_dl_bytes < bytes) {
        if (ctx->pinned_dl) { IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->pinned_dl); }
        ctx->pinned_dl = nullptr;
        ctx->pinned_dl_bytes = 0;
        const size_t n = align_up(bytes > (size_t)(4 << 20) ? bytes : (size_t)(4 << 20), 1 << 20);
        IS_CUDA(ctx, cudaMallocHost(&ctx->pinned_dl, n));
        ctx->pinned_dl_bytes = n;
    }
    IS_TRY(copy_small(ctx, ctx->pinned_dl, src, bytes, cudaMemcpyDeviceToHost));
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(dst, ctx->pinned_dl, bytes);
    return IS_OK;
}

HostPool* host_pool(is_ctx* ctx) {
    if (!ctx->hpool) {
        // the cores this process may run on, shared with the other ranks of the box (one process per GPU: torchrun exports
        // LOCAL_WORLD_SIZE) -- eight ranks with eight spinning threads each on a 16-core host only get in each other's way
        unsigned hc = std::thread::hardware_concurrency();
        {
            cpu_set_t set;
            CPU_ZERO(&set);
            if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) hc = (unsigned)CPU_COUNT(&set);
        }
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) { const int lws = atoi(e); if (lws > 1) hc = std::max(1u, hc / (unsigned)lws); }
        size_t w = hc > 2 ? std::min<size_t>(hc - 1, 7) : (hc == 2 ? 1 : 0);
        if (const char* e = getenv("IS_HOST_THREADS")) w = (size_t)std::max(0, atoi(e) - 1);
        ctx->hpool = new HostPool(w);
    }
    return ctx->hpool;
}

int download_view(is_ctx* ctx, const void* src, size_t bytes, const void** view) {
    if (ctx->pinned_dl_bytes < bytes || !ctx->pinned_dl) {
        if (ctx->pinned_dl) { IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->pinned_dl); }
        ctx->pinned_dl = nullptr;
        ctx->pinned_dl_bytes = 0;
        const size_t n = align_up(bytes > (size_t)(4 << 20) ? bytes : (size_t)(4 << 20), 1 << 20);
        IS_CUDA(ctx, cudaMallocHost(&ctx->pinned_dl, n));
        ctx->pinned_dl_bytes = n;
    }
    if (bytes) IS_TRY(copy_small(ctx, ctx->pinned_dl, src, bytes, cudaMemcpyDeviceToHost));
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *view = ctx->pinned_dl;
    return IS_OK;
}

// pitched download (rows of `width` bytes) through the pinned bounce buffer; dst is packed with dpitch
int download2d(is_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) {
    if (width == 0 || height == 0) return IS_OK;
    const size_t bytes = width * height;
    if (ctx->pinned_dl_bytes < bytes) {
        if (ctx->pinned_dl) { IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->pinned_dl); }
        ctx->pinned_dl = nullptr;
        ctx->pinned_dl_bytes = 0;
        const size_t n = align_up(bytes > (size_t)(4 << 20) ? bytes : (size_t)(4 << 20), 1 << 20);
        IS_CUDA(ctx, cudaMallocHost(&ctx->pinned_dl, n));
        ctx->pinned_dl_bytes = n;
    }
    IS_CUDA(ctx, cudaMemcpy2DAsync(ctx->pinned_dl, width, src, spitch, width, height, cudaMemcpyDeviceToHost, ctx->stream));
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t r = 0; r < height; ++r) std::memcpy((char*)dst + r * dpitch, (const char*)ctx->pinned_dl + r * width, width);
    return IS_OK;
}

}  // namespace is

extern "C" {

const char* is_version(void) { return "imagestitch_b200 0.1 (sm_100a)"; }

const char* is_status_string(int status) {
    switch (status) {
        case IS_OK: return "ok";
        case IS_ERR_NO_MEM: return "out of memory";
        case IS_ERR_BAD_ARG: return "bad argument";
        case IS_ERR_UNSUPPORTED: return "not implemented";
        case IS_ERR_ASSERT: return "assertion failed";
        case IS_ERR_CUDA: return "CUDA error";
        case IS_ERR_INTERNAL: return "internal error";
        default: return status > 0 ? "ok (informational)" : "unknown error";
    }
}

int is_ctx_create(int device, is_ctx** out) {
    if (!out) return IS_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return IS_ERR_CUDA;   // no CPU fallback: fail loudly
    if (device < 0 || device >= count) return IS_ERR_BAD_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return IS_ERR_CUDA;
    is_ctx* c = new is_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return IS_ERR_CUDA; }
    c->stream = c->own_stream;
    if (cudaDeviceGetDefaultMemPool(&c->pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;   // keep freed workspace cached in the pool between calls
        cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    for (int i = 0; i < 5; ++i) cudaEventCreate(&c->ev[i]);
    *out = c;
    return IS_OK;
}

int is_ctx_destroy(is_ctx* ctx) {
    if (!ctx) return IS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    delete ctx->hpool;
    ctx->hpool = nullptr;
    for (is_ctx* c : ctx->children) is_ctx_destroy(c);
    ctx->children.clear();
    for (int i = 0; i < 5; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (auto& r : ctx->krecs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : ctx->kpool) cudaEventDestroy(e);
    for (auto e : ctx->sync_events) cudaEventDestroy(e);
    for (auto& kv : ctx->block_cache) cudaFreeAsync(kv.second, ctx->stream);
    ctx->block_cache.clear();
    cudaStreamSynchronize(ctx->stream);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pinned_dl) cudaFreeHost(ctx->pinned_dl);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return IS_OK;
}

int is_ctx_synchronize(is_ctx* ctx) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IS_OK;
}

int is_ctx_set_stream(is_ctx* ctx, void* stream) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)stream;     // NULL is a stream too: the legacy default stream (torch's default)
    return IS_OK;
}

int is_ctx_reset_stream(is_ctx* ctx) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = ctx->own_stream;
    return IS_OK;
}

int is_ctx_kernel_timing(is_ctx* ctx, int enable) {
    if (!ctx) return IS_ERR_BAD_ARG;
    ctx->ktiming = enable != 0;
    return IS_OK;
}

// JSON array [{"name": ..., "launches": n, "ms": total, "bytes": algorithmic bytes}, ...] of the launches recorded
// since the last report; synchronises the stream.  Returns the length needed (including the NUL).
int is_ctx_kernel_timing_report(is_ctx* ctx, char* buf, size_t cap) {
    if (!ctx) return IS_ERR_BAD_ARG;
    IS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    struct Agg { std::string name; int n; double ms; double bytes; };
    std::vector<Agg> agg;
    for (auto& r : ctx->krecs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        std::string nm = r.name;
        size_t i = 0;
        for (; i < agg.size(); ++i) if (agg[i].name == nm) break;
        if (i == agg.size()) agg.push_back(Agg{nm, 0, 0., 0.});
        agg[i].n++; agg[i].ms += ms; agg[i].bytes += r.bytes;
        ctx->kpool.push_back(r.e0);
        ctx->kpool.push_back(r.e1);
    }
    ctx->krecs.clear();
    std::string out = "[";
    for (size_t i = 0; i < agg.size(); ++i) {
        char line[512];
        std::string nm;
        for (char c : agg[i].name) if (c != '"' && c != '\\') nm += c;
        snprintf(line, sizeof(line), "%s{\"name\": \"%s\", \"launches\": %d, \"ms\": %.6f, \"bytes\": %.0f}", i ? ", " : "", nm.c_str(),
                 agg[i].n, agg[i].ms, agg[i].bytes);
        out += line;
    }
    out += "]";
    if (buf && cap) {
        size_t n = std::min(cap - 1, out.size());
        std::memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int)out.size() + 1;
}

int is_ctx_seam_speculation(const is_ctx* ctx) { return ctx ? ctx->seam_speculation_accepted : -1; }
int is_ctx_seam_path(const is_ctx* ctx) { return ctx ? ctx->seam_path : -1; }
int is_ctx_seam_waves(const is_ctx* ctx) { return ctx ? ctx->seam_waves : -1; }

int is_ctx_clear_plan_cache(is_ctx* ctx) {
    if (!ctx) return IS_ERR_BAD_ARG;
    ctx->plan_cache.clear();
    return IS_OK;
}

const char* is_ctx_last_error(const is_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
void* is_ctx_stream(is_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t is_ctx_kernel_launches(const is_ctx* ctx) { return ctx ? ctx->launches : 0; }
int is_ctx_device(const is_ctx* ctx) { return ctx ? ctx->device : -1; }

}  // extern "C"

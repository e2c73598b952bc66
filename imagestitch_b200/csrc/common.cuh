// common.cuh -- context, error plumbing and device-buffer helpers shared by the kernels' host code.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/imagestitch.h"

namespace is { class HostPool; }

struct is_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;       // stream in use (own_stream or an adopted one)
    cudaStream_t own_stream = nullptr;
    cudaMemPool_t pool = nullptr;
    uint64_t launches = 0;
    std::string last_error;
    // pinned staging for small host<->device control transfers
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    size_t pinned_off = 0;
    void* pinned_dl = nullptr;           // bounce buffer for downloads
    size_t pinned_dl_bytes = 0;
    float timings[4] = {0, 0, 0, 0};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // optional per-launch CUDA-event timing (is_ctx_kernel_timing): one (start, stop) pair per launch
    bool ktiming = false;
    struct KRec { const char* name; cudaEvent_t e0, e1; double bytes; };
    std::vector<KRec> krecs;
    std::vector<cudaEvent_t> kpool;
    double next_bytes = 0;   // algorithmic bytes the next launch accounts for (set by the host code, optional)
    // memo of warp_plan results (camera -> ROI): detectResultRoi is a few hundred microseconds of host libm per image
    std::vector<unsigned char> plan_cache;
    // worker contexts (own stream + staging) for the concurrent seam pairs; owned by this context
    std::vector<is_ctx*> children;
    // Free blocks of this context's stream-ordered allocations, by size.  A step allocates the same few dozen sizes over and
    // over; reusing them here skips the driver's allocator (a lock shared with every other stream) in steady state.  Safe
    // because a block is only handed out again on the stream it was released on.
    std::multimap<size_t, void*> block_cache;
    size_t block_cache_bytes = 0;
    // events for stream_after(): a fixed ring, reused round-robin
    std::vector<cudaEvent_t> sync_events;
    size_t sync_next = 0;
    std::vector<double> last_gains;       // exposure gains of the last is_pipeline_run (is_pipeline_last_gains)
    int seam_speculation_accepted = -1;   // last is_seam_dp_find: 1 concurrent result accepted, 0 fell back, -1 not attempted
    int seam_waves = 0;                   // last is_seam_dp_find: waves of the batched path
    int seam_path = 0;                    // last is_seam_dp_find: 2 batched path, 1 per-pair concurrent path, 0 sequential loop
    is::HostPool* hpool = nullptr;        // host threads for the per-pair control work of the seam stage (hostpool.h)
    // Work for a side stream that should start BEHIND the first kernels of the seam stage (toggles + special points: short, on the
    // critical path, and five times slower when the pyramid kernels run beside them): the seam stage calls it once, right after
    // it has queued those kernels; whoever set it runs it afterwards if it is still there.
    std::function<int()> deferred_side_work;
};

namespace is {

int fail(is_ctx* ctx, int status, const char* fmt, ...);
cudaEvent_t ktiming_begin(is_ctx* ctx, const char* name);   // records the start event, returns the stop event

#define IS_CUDA(ctx, expr)                                                                            \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return is::fail((ctx), IS_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                            cudaGetErrorString(_e));                                                  \
    } while (0)

#define IS_TRY(expr)                   \
    do {                               \
        int _s = (expr);               \
        if (_s != IS_OK) return _s;    \
    } while (0)

#define IS_REQUIRE(ctx, cond, status, msg)                                       \
    do {                                                                         \
        if (!(cond)) return is::fail((ctx), (status), "%s (%s)", (msg), #cond);  \
    } while (0)

// Every kernel launch of the library goes through this macro: it counts the launch (bench.py's
// gpu_launches) and surfaces launch-configuration errors immediately.
#define IS_LAUNCH(ctx, kernel, grid, block, smem, ...)                                  \
    do {                                                                                \
        cudaEvent_t _k1 = nullptr;                                                      \
        if ((ctx)->ktiming) _k1 = is::ktiming_begin((ctx), #kernel);                    \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                \
        if (_k1) cudaEventRecord(_k1, (ctx)->stream);                                   \
        (ctx)->launches++;                                                              \
        IS_CUDA((ctx), cudaGetLastError());                                             \
    } while (0)

inline int depth_bytes(int depth) {
    switch (depth) {
        case IS_8U: return 1;
        case IS_16S: return 2;
        case IS_32S: return 4;
        case IS_32F: return 4;
        default: return 0;
    }
}

inline size_t elem_bytes(const is_mat& m) { return (size_t)depth_bytes(m.depth) * m.channels; }

inline int div_up(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Stream-ordered device allocation owned by a context.
struct DevBuf {
    is_ctx* ctx = nullptr;
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            ctx = o.ctx; p = o.p; bytes = o.bytes;
            o.p = nullptr; o.bytes = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    int alloc(is_ctx* c, size_t n);
    void release();
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// A device-resident view of an is_mat.  Host mats are staged into a pitched device buffer (and
// copied back on commit() when they are outputs); device mats are used in place.
struct DevMat {
    void* data = nullptr;
    int rows = 0, cols = 0, channels = 0, depth = 0;
    size_t step = 0;
    DevBuf owned;          // set when staged
    const is_mat* host = nullptr;
    template <typename T> T* ptr() const { return reinterpret_cast<T*>(data); }
    size_t row_bytes() const { return (size_t)cols * channels * depth_bytes(depth); }
};

int stage_in(is_ctx* ctx, const is_mat* m, DevMat* out);                 // read access
int stage_out(is_ctx* ctx, is_mat* m, DevMat* out, bool preload);        // write access (preload: in-out)
int commit(is_ctx* ctx, DevMat* m);                                      // copy back to the host mat if staged
int alloc_mat(is_ctx* ctx, int rows, int cols, int channels, int depth, DevMat* out, size_t align = 256);

int ensure_pinned(is_ctx* ctx, size_t bytes);
// bump allocation out of the pinned staging buffer for async uploads; wraps around behind a stream sync
int pinned_alloc(is_ctx* ctx, size_t bytes, void** out);
// small control transfers through the pinned staging buffer (synchronous with the stream)
int upload(is_ctx* ctx, void* dst, const void* src, size_t bytes);
// one side in device memory, the other in pinned (mapped) host memory: moved by a kernel, not by a copy engine (see ctx.cu)
int copy_small(is_ctx* ctx, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
int download(is_ctx* ctx, void* dst, const void* src, size_t bytes);
int download2d(is_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height);
// download into the pinned bounce buffer and hand out a view of it (valid until the next download on this context)
int download_view(is_ctx* ctx, const void* src, size_t bytes, const void** view);

int check_mat(is_ctx* ctx, const is_mat* m, const char* name);
HostPool* host_pool(is_ctx* ctx);      // the context's host thread pool (created on first use)

enum { SIDE_PYRAMID = 8, SIDE_COPY = 9 };
int child_ctx(is_ctx* parent, size_t k, is_ctx** out);
int stream_after(is_ctx* ctx, cudaStream_t to, cudaStream_t from);
void merge_child(is_ctx* parent, is_ctx* child);

}  // namespace is

// tests/emu/orb_emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the marked region of imagestitch_b200/csrc/orb.cu: the ORB kernels AND the driver that strings them together
// (orb_find_core: level table, resize coefficients, the host-side selections between the launches), with "device memory" = host
// memory and every launch stepped through thread by thread.
#include "cuda_host_emul.h"

#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../include/imagestitch.h"

struct is_ctx { int unused; };
inline int __float2int_rn_(float v) { return __float2int_rn(v); }
inline float __int2float_rn(int v) { return (float)v; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p += v; return o; }
static inline int div_up(int a, int b) { return (a + b - 1) / b; }
#define IS_TRY(expr) do { int _s = (expr); if (_s != IS_OK) return _s; } while (0)

namespace is {
struct OrbBuf { void* p = nullptr; std::vector<uint8_t> store; };
static int orb_alloc(is_ctx*, OrbBuf* b, size_t bytes) { b->store.assign(bytes + 64, 0xcd); b->p = b->store.data(); return IS_OK; }
static int orb_h2d(is_ctx*, void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); return IS_OK; }
static int orb_d2h(is_ctx*, void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); return IS_OK; }
static int orb_d2h_view(is_ctx*, const void* src, size_t, const void** view) { *view = src; return IS_OK; }
static int orb_zero(is_ctx*, void* dst, size_t bytes) { std::memset(dst, 0, bytes); return IS_OK; }
static int orb_dump(is_ctx*, const char* name, const void* p, size_t bytes) {
    const char* dir = getenv("IS_ORB_DUMP");
    if (!dir) return IS_OK;
    const std::string path = std::string(dir) + "/" + name + ".bin";
    if (FILE* f = std::fopen(path.c_str(), "wb")) { std::fwrite(p, 1, bytes, f); std::fclose(f); }
    return IS_OK;
}
// the host pool's place: one OS thread per index (IS_EMU_SERIAL=1: a plain loop), so that the sections the product runs on its
// pool are exercised concurrently here as well
template <typename F> static void orb_parallel_for(is_ctx*, size_t n, F&& fn) {
    if (getenv("IS_EMU_SERIAL")) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::vector<std::thread> th;
    for (size_t i = 0; i < n; ++i) th.emplace_back([&fn, i] { fn(i); });
    for (auto& t : th) t.join();
}
#define ORB_LAUNCH(ctx, kernel, grid, block, ...) emu_launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })
#include "orb_region.inc"
}
using namespace is;

// kps: 6 floats per key point (x, y, size, angle, response, octave); desc: 32 bytes each; returns the count or a negative status
extern "C" int emu_orb_find(const uint8_t* img, int rows, int cols, int ch, size_t step, int grid_w, int grid_h, int nfeatures, float scale_factor, int nlevels,
                            float* kps, uint8_t* desc, int cap) {
    is_ctx ctx;
    OrbParams P{nfeatures, scale_factor, nlevels, grid_w, grid_h, 31, 31, 20};
    std::vector<KeyPt> k;
    std::vector<uint8_t> d;
    const int rc = orb_find_core(&ctx, img, step, rows, cols, ch, P, &k, &d);
    if (rc != IS_OK) return rc;
    const int n = std::min(cap, (int)k.size());
    for (int i = 0; i < n; ++i) {
        float* o = kps + 6 * (size_t)i;
        o[0] = k[(size_t)i].x; o[1] = k[(size_t)i].y; o[2] = k[(size_t)i].size; o[3] = k[(size_t)i].angle; o[4] = k[(size_t)i].response; o[5] = (float)k[(size_t)i].octave;
    }
    if (n) std::memcpy(desc, d.data(), (size_t)n * 32);
    return (int)k.size();
}

// tests/emu/cuda_host_emul_mt.h -- TEST INFRASTRUCTURE ONLY.
// Multi-threaded host emulation of ONE thread block at a time: every CUDA thread of the block is an OS thread, so kernels
// with shared memory, __syncthreads and spin-waits on mbarriers run as written.  Blocks of a grid run one after the other.
//   __shared__ x;                 -> static storage (the test's extraction turns `extern __shared__ ... name[];` into a
//                                    pointer to emu_dynamic_smem first)
//   __syncthreads()               -> a reusable barrier over the block's threads
//   mbarrier / bulk copies        -> tests/emu/tma.cuh (same names as imagestitch_b200/csrc/tma.cuh): the copy happens at
//                                    issue time, the transaction count and phase bookkeeping follow the PTX semantics
//   __shfl_{xor,up,down,}_sync    -> (full mask) exchange through a per-warp slot array between two per-warp barriers
// Not emulated: votes, clusters, 2-D tensor maps.
#pragma once

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __align__(n) alignas(n)
#define __launch_bounds__(...)

struct EmuDim3 {
    unsigned x = 1, y = 1, z = 1;
    EmuDim3() {}
    EmuDim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef EmuDim3 dim3;
static thread_local EmuDim3 threadIdx;
static EmuDim3 blockIdx, blockDim, gridDim;

struct alignas(16) float4 { float x, y, z, w; };

inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
template <typename T> inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;

alignas(128) static unsigned char emu_dynamic_smem[256 * 1024];
static std::unique_ptr<std::barrier<>> emu_block_barrier;
static std::unique_ptr<std::barrier<>> emu_warp_barrier[32];
static uint32_t emu_shfl_slot[32][32];
inline void __syncthreads() { emu_block_barrier->arrive_and_wait(); }

template <typename T> inline T __shfl_xor_sync(unsigned /*full mask*/, T v, int lane_mask) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    std::memcpy(&emu_shfl_slot[w][l], &v, 4);
    emu_warp_barrier[w]->arrive_and_wait();
    T r;
    std::memcpy(&r, &emu_shfl_slot[w][l ^ (unsigned)lane_mask], 4);
    emu_warp_barrier[w]->arrive_and_wait();
    return r;
}

template <typename T, typename Pick> inline T emu_shfl(T v, Pick pick) {          // pick(lane) -> source lane (or lane itself: keep own value)
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    std::memcpy(&emu_shfl_slot[w][l], &v, 4);
    emu_warp_barrier[w]->arrive_and_wait();
    T r;
    std::memcpy(&r, &emu_shfl_slot[w][pick(l)], 4);
    emu_warp_barrier[w]->arrive_and_wait();
    return r;
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d) { return emu_shfl(v, [d](unsigned l) { return l >= d ? l - d : l; }); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d) { return emu_shfl(v, [d](unsigned l) { return l + d < 32 ? l + d : l; }); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, [src](unsigned) { return (unsigned)src & 31; }); }

inline int __float2int_rn(float v) {                      // cvt.rni.s32.f32: round half even, saturate, NaN -> 0
    if (v != v) return 0;
    if (v >= 2147483648.f) return 2147483647;
    if (v <= -2147483648.f) return -2147483647 - 1;
    return (int)std::nearbyintf(v);
}
// byte dot products, funnel shift and byte permute as PTX defines them (unsigned forms; prmt without the sign-replicating modes)
inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
    return c;
}
inline uint32_t __dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { return c + (a & 0xffffu) * (b & 255u) + (a >> 16) * ((b >> 8) & 255u); }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u)); }
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 255u) << (8 * i);
    return r;
}

inline int __float2int_rz(float v) {                        // cvt.rzi.s32.f32: truncate, saturate, NaN -> 0
    if (v != v) return 0;
    if (v >= 2147483648.f) return 2147483647;
    if (v <= -2147483648.f) return -2147483647 - 1;
    return (int)v;
}

// kernels without barriers or shuffles: one host thread steps through the grid
template <typename F> inline void emu_launch(dim3 grid, dim3 block, F&& kernel_call) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tz = 0; tz < block.z; ++tz) for (unsigned ty = 0; ty < block.y; ++ty) for (unsigned tx = 0; tx < block.x; ++tx) {
            blockIdx = EmuDim3(bx, by, bz); threadIdx = EmuDim3(tx, ty, tz);
            kernel_call();
        }
}

// one block after the other; the threads of a block concurrently (1-D blocks are enough for the kernels covered)
template <typename F> inline void emu_launch_mt(unsigned grid_x, unsigned block_x, F&& kernel_call) {
    gridDim = EmuDim3(grid_x); blockDim = EmuDim3(block_x);
    for (unsigned b = 0; b < grid_x; ++b) {
        blockIdx = EmuDim3(b);
        emu_block_barrier = std::make_unique<std::barrier<>>((std::ptrdiff_t)block_x);
        for (unsigned w = 0; w * 32 < block_x; ++w) emu_warp_barrier[w] = std::make_unique<std::barrier<>>((std::ptrdiff_t)std::min(32u, block_x - w * 32));
        std::vector<std::thread> th;
        th.reserve(block_x);
        for (unsigned t = 0; t < block_x; ++t)
            th.emplace_back([&, t] {
                threadIdx = EmuDim3(t);
                kernel_call();
                emu_block_barrier->arrive_and_drop();          // a thread that returned early no longer takes part in barriers
                emu_warp_barrier[t >> 5]->arrive_and_drop();
            });
        for (auto& x : th) x.join();
    }
}

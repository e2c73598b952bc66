// tests/emu/cuda_host_emul_mt.h -- TEST INFRASTRUCTURE ONLY.
// Multi-threaded host emulation of ONE thread block at a time: every CUDA thread of the block is an OS thread, so kernels
// with shared memory, __syncthreads and spin-waits on mbarriers run as written.  Blocks of a grid run one after the other.
//   __shared__ x;                 -> static storage (the test's extraction turns `extern __shared__ ... name[];` into a
//                                    pointer to emu_dynamic_smem first)
//   __syncthreads()               -> a reusable barrier over the block's threads
//   mbarrier / bulk copies        -> tests/emu/tma.cuh (same names as imagestitch_b200/csrc/tma.cuh): the copy happens at
//                                    issue time, the transaction count and phase bookkeeping follow the PTX semantics
// Not emulated: warp shuffles / votes, clusters, 2-D tensor maps.
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
inline void __syncthreads() { emu_block_barrier->arrive_and_wait(); }

// one block after the other; the threads of a block concurrently (1-D blocks are enough for the kernels covered)
template <typename F> inline void emu_launch_mt(unsigned grid_x, unsigned block_x, F&& kernel_call) {
    gridDim = EmuDim3(grid_x); blockDim = EmuDim3(block_x);
    for (unsigned b = 0; b < grid_x; ++b) {
        blockIdx = EmuDim3(b);
        emu_block_barrier = std::make_unique<std::barrier<>>((std::ptrdiff_t)block_x);
        std::vector<std::thread> th;
        th.reserve(block_x);
        for (unsigned t = 0; t < block_x; ++t)
            th.emplace_back([&, t] {
                threadIdx = EmuDim3(t);
                kernel_call();
                emu_block_barrier->arrive_and_drop();          // a thread that returned early no longer takes part in barriers
            });
        for (auto& x : th) x.join();
    }
}

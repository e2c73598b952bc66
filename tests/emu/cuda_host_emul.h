// tests/emu/cuda_host_emul.h -- TEST INFRASTRUCTURE ONLY.
// Enough of the CUDA device vocabulary to compile simple one-thread-per-element kernels of imagestitch_b200/csrc for the
// HOST (g++ -ffp-contract=off): qualifiers vanish, the *_rn intrinsics are the IEEE operations they name, threadIdx /
// blockIdx / blockDim are plain variables that emu_launch() steps through.  Kernels with shared memory, barriers, warp
// shuffles or TMA are out of its reach -- it covers per-thread arithmetic and indexing, which is what it is for.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __restrict__

struct EmuDim3 {
    unsigned x = 1, y = 1, z = 1;
    EmuDim3() {}
    EmuDim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef EmuDim3 dim3;
static EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __float2int_rn(float v) {                      // cvt.rni.s32.f32: round half even, saturate, NaN -> 0
    if (v != v) return 0;
    if (v >= 2147483648.f) return 2147483647;
    if (v <= -2147483648.f) return -2147483647 - 1;
    return (int)std::nearbyintf(v);
}
template <typename T> inline T __ldg(const T* p) { return *p; }
#define __launch_bounds__(...)
using std::max;
using std::min;

template <typename F> inline void emu_launch(dim3 grid, dim3 block, F&& kernel_call) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tz = 0; tz < block.z; ++tz) for (unsigned ty = 0; ty < block.y; ++ty) for (unsigned tx = 0; tx < block.x; ++tx) {
            blockIdx = EmuDim3(bx, by, bz); threadIdx = EmuDim3(tx, ty, tz);
            kernel_call();
        }
}

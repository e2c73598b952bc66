// tests/emu/tma.cuh -- TEST INFRASTRUCTURE ONLY: host stand-ins for imagestitch_b200/csrc/tma.cuh (mbarrier + 1-D bulk copy).
// mbarrier state in the 64-bit word: bits 0..15 pending arrivals, bit 16 phase parity, bits 17..31 expected arrivals, bits 32..63
// outstanding transaction bytes.  A phase completes when pending == 0 and tx == 0: parity flips, pending = expected.
#pragma once

#include "cuda_host_emul_mt.h"

namespace is {

inline void emu_mbar_update(uint64_t* bar, int arrive, int64_t tx_delta) {
    std::atomic_ref<uint64_t> a(*bar);
    uint64_t old = a.load(std::memory_order_acquire), neu;
    do {
        uint64_t pending = old & 0xffff, parity = (old >> 16) & 1, expected = (old >> 17) & 0x7fff;
        int64_t tx = (int64_t)(old >> 32) + tx_delta;
        pending -= (uint64_t)arrive;
        if (pending == 0 && tx == 0) { parity ^= 1; pending = expected; }
        neu = pending | (parity << 16) | (expected << 17) | ((uint64_t)tx << 32);
    } while (!a.compare_exchange_weak(old, neu, std::memory_order_acq_rel));
}
inline void mbar_init(uint64_t* bar, int count) {
    std::atomic_ref<uint64_t>(*bar).store((uint64_t)count | ((uint64_t)count << 17), std::memory_order_release);
}
inline void mbar_fence_init() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_mbar_update(bar, 1, (int64_t)bytes); }   // arrive.expect_tx
inline void mbar_wait(uint64_t* bar, uint32_t parity) {                                                   // try_wait.parity loop
    std::atomic_ref<uint64_t> a(*bar);
    while (((a.load(std::memory_order_acquire) >> 16) & 1) == parity) std::this_thread::yield();
}
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {                          // cp.async.bulk + complete_tx
    std::memcpy(dst, src, bytes);
    emu_mbar_update(bar, 0, -(int64_t)bytes);
}

}  // namespace is

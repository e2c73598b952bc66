// hostpool.h -- a small persistent pool of host threads for the per-pair control work of the seam stage (run merging,
// contour lists, the order-dependent walk of updateLabelsUsingSeam): a few hundred microseconds per pair that would otherwise
// add up on the calling thread while the GPU waits.  Workers spin briefly between the phases of one call, then sleep.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace is {

class HostPool {
public:
    explicit HostPool(size_t workers) {
        for (size_t t = 0; t < workers; ++t) th_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    size_t workers() const { return th_.size(); }

    // fn(k) for k in [0, n); the caller takes part and returns when all are done
    void run(size_t n, const std::function<void(size_t)>& fn) {
        if (n == 0) return;
        if (n == 1 || th_.empty()) { for (size_t k = 0; k < n; ++k) fn(k); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            pending_.store(n, std::memory_order_relaxed);
            fn_ = &fn;
            n_ = n;
            next_ = 0;
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        work();
        // wait for the stragglers (short tasks: spin)
        while (pending_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
        std::lock_guard<std::mutex> lk(m_);
        fn_ = nullptr;
    }

private:
    void work() {
        for (;;) {
            const std::function<void(size_t)>* f;
            size_t k;
            {
                std::lock_guard<std::mutex> lk(m_);      // tasks take ~100 us each: claiming them under the lock costs nothing
                if (!fn_ || next_ >= n_) return;
                k = next_++;
                f = fn_;
            }
            (*f)(k);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            // spin for a while (the phases of one call follow each other within a millisecond or two of device work), then sleep
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(1500)) {
                if (gen_.load(std::memory_order_acquire) != seen) { got = true; break; }
                std::this_thread::yield();
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
            }
            {
                std::lock_guard<std::mutex> lk(m_);
                seen = gen_.load(std::memory_order_acquire);
                if (stop_) return;
                if (!fn_) continue;
            }
            work();
        }
    }

    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t n_ = 0;
    size_t next_ = 0;
    std::atomic<size_t> pending_{0};
    std::atomic<uint64_t> gen_{0};
    bool stop_ = false;
};

}  // namespace is

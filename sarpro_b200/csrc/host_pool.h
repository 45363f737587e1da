// host_pool.h -- the library's host worker threads (threshold tables of the general f32 path, host-side narrowing of f32
// rasters on their way to the device). Header-only; one pool per process.
#pragma once
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>

namespace sarpro {

// ---- worker pool ------------------------------------------------------------------------------
// The host work of a call sits on its critical path between two kernels or ahead of an upload (threshold tables: 2 x (4,095 +
// 65,535) independent boundaries per two-operation call; narrowing: 100 MB row chunks). Spawning 16 threads per table cost
// more than the table itself; the workers are therefore created once per process and woken per job. Chunks are pulled from
// an atomic counter, the calling thread works too. One job at a time (contexts on other threads queue on `gate`).
class WorkerPool {
public:
    static WorkerPool& get() {
        static std::once_flag once;
        std::call_once(once, [] { pthread_atfork(nullptr, nullptr, [] { inst_ = nullptr; }); });
        WorkerPool* p = inst_.load(std::memory_order_acquire);
        if (!p) {
            std::lock_guard<std::mutex> lk(make_mu_);
            p = inst_.load(std::memory_order_acquire);
            if (!p) { p = new WorkerPool(); inst_.store(p, std::memory_order_release); } // never destroyed: workers may outlive static destructors
        }
        return *p;
    }
    unsigned width() const { return n_workers_ + 1; }

    void run(uint32_t n, uint32_t chunk, const std::function<void(uint32_t, uint32_t)>& f) {
        if (n == 0) return;
        if (n_workers_ == 0 || n <= chunk) { f(0, n); return; }
        std::lock_guard<std::mutex> job(gate_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &f; n_ = n; chunk_ = chunk;
            next_.store(0, std::memory_order_relaxed);
            pending_ = n_workers_;
            ++generation_;
        }
        cv_.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    WorkerPool() {
        // the CPUs this process may run on (a rank of a multi-process job is often confined to a subset), at most 16
        unsigned hw = std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) hw = (unsigned)CPU_COUNT(&set);
        n_workers_ = hw > 1 ? std::min(16u, hw) - 1 : 0;
        for (unsigned i = 0; i < n_workers_; ++i) std::thread([this] { loop(); }).detach();
    }
    void drain() {
        for (;;) {
            const uint32_t a = next_.fetch_add(chunk_, std::memory_order_relaxed);
            if (a >= n_) break;
            (*fn_)(a, std::min(n_, a + chunk_));
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
            }
            drain();
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    static inline std::atomic<WorkerPool*> inst_{nullptr};
    static inline std::mutex make_mu_;
    std::mutex gate_, mu_;
    std::condition_variable cv_, done_;
    const std::function<void(uint32_t, uint32_t)>* fn_ = nullptr;
    uint32_t n_ = 0, chunk_ = 1;
    std::atomic<uint32_t> next_{0};
    unsigned pending_ = 0, n_workers_ = 0;
    uint64_t generation_ = 0;
};

template <typename F>
void parallel_for(uint32_t n, F&& f) {
    if (n < 2048) { f(0, n); return; }
    WorkerPool& pool = WorkerPool::get();
    // ~4 chunks per thread: the searches near the clip ends are cheaper than the ones inside the window
    const uint32_t chunk = std::max<uint32_t>(256, (n + 4 * pool.width() - 1) / (4 * pool.width()));
    pool.run(n, chunk, std::function<void(uint32_t, uint32_t)>(std::forward<F>(f)));
}

} // namespace sarpro

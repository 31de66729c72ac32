// A small fork-join pool for the host side of the engine: the per-chain work of one MCMC step (drawing a move,
// compiling the proposal blob, committing / rolling back) is independent across chains, so a step of B chains is a
// parallel-for over B items followed by ONE kernel launch.  Steps are ~100 us apart, so the workers spin (SCICONE_SPIN_US,
// default 1000 us) before they sleep, and the calling thread takes part in the loop.
//
// No reference counterpart: the reference is one chain per process (global RNG singleton).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace scicone {

class ForkJoinPool {
public:
    typedef void (*Fn)(int index, void* ctx);

    static ForkJoinPool& instance() {
        static ForkJoinPool pool;
        return pool;
    }
    int n_threads() const { return (int)workers_.size() + 1; }

    // fn(i, ctx) for i in [0, n); returns when every item is done.  Not re-entrant (one caller at a time; a call
    // from inside a worker runs inline).
    void run(int n, Fn fn, void* ctx) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty() || in_worker_ || busy_.exchange(true)) {
            for (int i = 0; i < n; ++i) fn(i, ctx);
            return;
        }
        fn_ = fn;
        ctx_ = ctx;
        n_ = n;
        next_.store(0, std::memory_order_relaxed);
        pending_.store(n, std::memory_order_release);  // publishes fn_/ctx_/n_/next_ to a worker that sees it
        {
            std::lock_guard<std::mutex> lk(mu_);
            generation_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        work();
        while (pending_.load(std::memory_order_acquire) > 0) cpu_relax();
        // no worker may still be inside work() of this generation when the next one is published
        while (active_.load(std::memory_order_acquire) > 0) cpu_relax();
        busy_.store(false);
    }

private:
    ForkJoinPool() {
        int n = 0;
        if (const char* e = std::getenv("SCICONE_THREADS")) n = std::atoi(e);
        if (n <= 0) {
            n = (int)std::thread::hardware_concurrency();
            if (n > 16) n = 16;
        }
        if (n < 1) n = 1;
        if (const char* e = std::getenv("SCICONE_SPIN_US")) spin_us_ = std::max(0, std::atoi(e));
        for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~ForkJoinPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            generation_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (std::thread& t : workers_) t.join();
    }
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    void work() {
        for (;;) {
            const int i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_) break;
            fn_(i, ctx_);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    void loop() {
        in_worker_ = true;
        unsigned long long seen = 0;
        for (;;) {
            // spin, then sleep until the next generation (waking 15 sleepers costs more than a whole MCMC step)
            const auto t0 = std::chrono::steady_clock::now();
            bool fresh = false;
            while (!(fresh = generation_.load(std::memory_order_acquire) != seen)) {
                cpu_relax();
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) break;
            }
            if (!fresh) {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_.load(std::memory_order_acquire) != seen; });
            }
            if (stop_) return;
            active_.fetch_add(1, std::memory_order_acq_rel);
            seen = generation_.load(std::memory_order_acquire);
            if (pending_.load(std::memory_order_acquire) > 0) work();
            active_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }

    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::atomic<unsigned long long> generation_{0};
    std::atomic<int> next_{0}, pending_{0}, active_{0};
    std::atomic<bool> busy_{false};
    bool stop_ = false;
    int spin_us_ = 1000;
    Fn fn_ = nullptr;
    void* ctx_ = nullptr;
    int n_ = 0;
    static thread_local bool in_worker_;
};
inline thread_local bool ForkJoinPool::in_worker_ = false;

}  // namespace scicone

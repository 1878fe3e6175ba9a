// expand_pool.h — the host threads that expand the compact wire format (wire.cuh) into the caller's tensors.  Plain C++
// (no CUDA), so that tests/test_expand_pool.py can stress it on the CPU (thread sanitizer included).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

// Host threads that expand the compact wire format (wire.cuh) into the caller's tensors: persistent, woken per job, the
// calling thread works along.  One pool per process; fl_host_threads sets its size before first use.
class ExpandPool {
  public:
    void set_threads(int n) {
        std::lock_guard<std::mutex> lk(mu_);
        if (th_.empty()) want_ = n < 1 ? 1 : (n > 64 ? 64 : n);
    }
    int threads() {
        std::lock_guard<std::mutex> lk(mu_);
        return want_ ? want_ : default_threads();
    }
    // Every job is its own object (function copy, counters), handed to the workers as a shared pointer under the lock: a
    // worker that wakes up late finds either its job exhausted or the next job complete with that job's own function — it
    // can never pair one job's function with another job's counters (which an earlier version could, once in many thousand
    // calls: a null function pointer under a fresh index).
    void parallel_for(int n, const std::function<void(int)> &fn) {
        if (n <= 0) return;
        auto j = std::make_shared<Job>();
        j->fn = fn; j->n = n; j->left.store(n);
        {
            std::unique_lock<std::mutex> lk(mu_);
            if (th_.empty()) start_locked();
            job_ = j; gen_++;
        }
        cv_.notify_all();
        run(*j);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return j->left.load() == 0; });
        if (job_ == j) job_.reset();
    }
    ~ExpandPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }

  private:
    static int default_threads() {
        unsigned hc = std::thread::hardware_concurrency();
        int n = hc ? (int)hc : 4;
        if (const char *s = getenv("LOCAL_WORLD_SIZE")) { const int w = atoi(s); if (w > 1) n = n / w > 2 ? n / w : 2; }
        return n > 32 ? 32 : n;
    }
    void start_locked() {
        if (!want_) want_ = default_threads();
        for (int k = 1; k < want_; k++) th_.emplace_back([this] { loop(); });
    }
    struct Job {
        std::function<void(int)> fn;
        int n = 0;
        std::atomic<int> next{0}, left{0};
    };
    void run(Job &j) {
        for (;;) {
            const int k = j.next.fetch_add(1);
            if (k >= j.n) break;
            j.fn(k);
            if (j.left.fetch_sub(1) == 1) {          // the last item: wake the caller (under the lock: no lost wake-up)
                std::lock_guard<std::mutex> lk(mu_);
                done_cv_.notify_all();
            }
        }
    }
    void loop() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            std::shared_ptr<Job> j = job_;
            lk.unlock();
            if (j) run(*j);
            j.reset();
            lk.lock();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> th_;
    std::shared_ptr<Job> job_;
    int want_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

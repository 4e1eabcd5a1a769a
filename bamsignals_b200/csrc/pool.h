// Minimal worker pool for the host fetch pipeline (BGZF block scan, inflate, record walk, result scatter).
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace bsg {

class Pool {
public:
    explicit Pool(int n) {
        if (n < 1) n = 1;
        for (int i = 0; i < n; ++i) threads_.emplace_back([this, i] { run(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return int(threads_.size()); }

    // task(worker_index)
    void submit(std::function<void(int)> f) {
        {
            std::lock_guard<std::mutex> g(m_);
            q_.push_back(std::move(f));
        }
        cv_.notify_one();
    }
    // background work: only runs when no regular task is waiting
    void submit_low(std::function<void(int)> f) {
        {
            std::lock_guard<std::mutex> g(m_);
            low_.push_back(std::move(f));
        }
        cv_.notify_one();
    }

    // Run f(i, worker) for i in [0,n) on the pool and wait (the caller does not participate).
    void parallel_for(int64_t n, int64_t grain, const std::function<void(int64_t, int64_t, int)>& f) {
        if (n <= 0) return;
        if (grain < 1) grain = 1;
        int64_t chunks = (n + grain - 1) / grain;
        int64_t left = chunks;   // guarded by dm (decrement + notify under the lock: the waiter owns these objects)
        std::mutex dm;
        std::condition_variable dcv;
        for (int64_t c = 0; c < chunks; ++c) {
            submit([&, c](int w) {
                f(c * grain, std::min(n, (c + 1) * grain), w);
                std::lock_guard<std::mutex> g(dm);
                if (--left == 0) dcv.notify_all();
            });
        }
        std::unique_lock<std::mutex> lk(dm);
        dcv.wait(lk, [&] { return left == 0; });
    }

private:
    void run(int idx) {
        for (;;) {
            std::function<void(int)> f;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty() || !low_.empty(); });
                if (!q_.empty()) { f = std::move(q_.front()); q_.pop_front(); }
                else if (!low_.empty()) { f = std::move(low_.front()); low_.pop_front(); }
                else return;
            }
            f(idx);
        }
    }
    std::vector<std::thread> threads_;
    std::deque<std::function<void(int)>> q_, low_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

}  // namespace bsg

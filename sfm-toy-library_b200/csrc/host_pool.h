// host_pool.h -- a small persistent pool of host threads owned by the context.  The ORB stage (orb.cu) uses it for the work that has to
// stay on the host: staging caller images into pinned memory while earlier images are already on their way to the GPU, and the
// per-(image, level) KeyPointsFilter::retainBest selections (std::nth_element, order-defining -- see orb.cu).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

class HostPool {
public:
    explicit HostPool(int n_threads) {
        for (int i = 0; i < n_threads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; gen_++; }
        cv_work_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size() + 1; }
    // run fn(0..n-1), the calling thread takes part; returns when every task is done
    void parallel_for(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty()) { for (int i = 0; i < n; i++) fn(i); return; }
        auto job = std::make_shared<Job>();          // a worker that wakes up late keeps ITS job (already drained), never a newer counter
        job->fn = &fn; job->n = n; job->pending = n;
        { std::lock_guard<std::mutex> lk(m_); job_ = job; gen_++; }
        cv_work_.notify_all();
        drain(*job);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return job->pending == 0; });
        job_.reset();
    }
    static int default_threads() {
        if (const char* e = getenv("SFMB200_HOST_THREADS")) { int v = atoi(e); if (v >= 1) return v > 64 ? 64 : v; }
        unsigned hc = std::thread::hardware_concurrency();
        return hc == 0 ? 4 : (hc > 16 ? 16 : (int)hc);
    }

private:
    struct Job { const std::function<void(int)>* fn = nullptr; int n = 0; int pending = 0; std::atomic<int> next{0}; };
    void drain(Job& j) {
        for (;;) {
            const int i = j.next.fetch_add(1);
            if (i >= j.n) break;
            (*j.fn)(i);                               // fn outlives the job: parallel_for waits for pending == 0
            std::lock_guard<std::mutex> lk(m_);
            if (--j.pending == 0) cv_done_.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                job = job_;
            }
            if (job) drain(*job);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::shared_ptr<Job> job_;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

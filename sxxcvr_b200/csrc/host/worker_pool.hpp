// worker_pool.hpp -- a handful of parked threads that split a loop over [0, n) with their owner.
// The group device uses it to spread the per-front-end stream bookkeeping (a microsecond of
// ALSA calls per front-end per period) over the host's cores while the GPU does the one
// conversion for all of them.  Plain C++; tested on the CPU under ThreadSanitizer.
#pragma once

#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace sxhost {

class WorkerPool {
public:
    explicit WorkerPool(unsigned helpers) : helpers_(helpers) {}
    ~WorkerPool()
    {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            quit_ = true;
        }
        wake_.notify_all();
        for (std::thread &t : threads_)
            t.join();
    }
    WorkerPool(const WorkerPool &) = delete;
    WorkerPool &operator=(const WorkerPool &) = delete;

    unsigned helpers() const { return helpers_; }

    // body(begin, end) for disjoint ranges covering [0, n); returns when all are done.  Ranges
    // below `grain` items are not worth a wake-up: the owner runs everything itself.
    void run(size_t n, size_t grain, const std::function<void(size_t, size_t)> &body)
    {
        if (n == 0)
            return;
        const size_t parts = std::min<size_t>(helpers_ + 1, (n + grain - 1) / (grain ? grain : 1));
        if (parts <= 1) {
            body(0, n);
            return;
        }
        start_threads();
        {
            std::lock_guard<std::mutex> lock(mutex_);
            body_ = &body;
            n_ = n;
            parts_ = parts;
            pending_ = unsigned(parts - 1);
            generation_++;
        }
        wake_.notify_all();
        part(0);
        std::unique_lock<std::mutex> lock(mutex_);
        done_.wait(lock, [this] { return pending_ == 0; });
    }

private:
    void part(size_t k) const
    {
        const size_t per = (n_ + parts_ - 1) / parts_;
        const size_t lo = k * per, hi = std::min(n_, lo + per);
        if (lo < hi)
            (*body_)(lo, hi);
    }
    void start_threads()
    {
        if (!threads_.empty())
            return;
        for (unsigned h = 0; h < helpers_; h++)
            threads_.emplace_back([this, h] { main(h + 1); });
    }
    void main(size_t k)
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lock(mutex_);
        for (;;) {
            wake_.wait(lock, [&] { return quit_ || generation_ != seen; });
            if (quit_)
                return;
            seen = generation_;
            if (k >= parts_)
                continue; // fewer parts than threads this time
            lock.unlock();
            part(k);
            lock.lock();
            if (--pending_ == 0)
                done_.notify_one();
        }
    }

    const unsigned helpers_;
    std::vector<std::thread> threads_;
    std::mutex mutex_;
    std::condition_variable wake_, done_;
    bool quit_ = false;
    uint64_t generation_ = 0;
    unsigned pending_ = 0;
    const std::function<void(size_t, size_t)> *body_ = nullptr;
    size_t n_ = 0, parts_ = 1;
};

} // namespace sxhost

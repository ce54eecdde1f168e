// par_copy.hpp -- memcpy split over a few threads, for the bounce copies of pageable callers.
//
// The *_host entry points move a pageable caller buffer (a numpy array, a std::vector: what the
// reference's readStream/writeStream callers pass as buffs[0], SoapySX.cpp:868-875, :969-976)
// through pinned staging, because the copy engines cannot read pageable memory.  One thread
// copies at 5-10 GB/s, a fifth of what the PCIe link behind it carries; a handful of threads
// close most of that gap.  Plain C++, no CUDA: tests/test_host_logic.py builds and runs it on
// the CPU, under ThreadSanitizer where the toolchain has it.
#pragma once

#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace sxhost {

class ParallelCopier {
public:
    // Copies below this size are done by the caller alone: waking helpers costs tens of
    // microseconds, a single thread moves 1 MiB in about a hundred.
    static constexpr size_t kMinParallelBytes = size_t(2) << 20;

    // `helpers` threads are started on first use and parked between copies.
    explicit ParallelCopier(unsigned helpers) : helpers_(helpers) {}

    ~ParallelCopier()
    {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            quit_ = true;
        }
        wake_.notify_all();
        for (std::thread &t : threads_)
            t.join();
    }

    ParallelCopier(const ParallelCopier &) = delete;
    ParallelCopier &operator=(const ParallelCopier &) = delete;

    unsigned helpers() const { return helpers_; }

    // memcpy(dst, src, bytes); returns when every byte has been copied.  One copy at a time
    // (the caller serialises: the host pipeline holds the context's host mutex).
    void copy(void *dst, const void *src, size_t bytes)
    {
        if (helpers_ == 0 || bytes < kMinParallelBytes) {
            std::memcpy(dst, src, bytes);
            return;
        }
        start_threads();
        const unsigned parts = helpers_ + 1;
        // slices are multiples of 4 KiB so that no two threads share a page of the destination
        const size_t slice = ((bytes + parts - 1) / parts + 4095) & ~size_t(4095); // parts * slice >= bytes
        {
            std::lock_guard<std::mutex> lock(mutex_);
            dst_ = static_cast<char *>(dst);
            src_ = static_cast<const char *>(src);
            bytes_ = bytes;
            slice_ = slice;
            pending_ = helpers_;
            generation_++;
        }
        wake_.notify_all();
        copy_slice(0); // the caller takes the first slice
        std::unique_lock<std::mutex> lock(mutex_);
        done_.wait(lock, [this] { return pending_ == 0; });
    }

private:
    void copy_slice(unsigned part) const
    {
        const size_t lo = size_t(part) * slice_;
        if (lo >= bytes_)
            return;
        std::memcpy(dst_ + lo, src_ + lo, bytes_ - lo < slice_ ? bytes_ - lo : slice_);
    }

    void start_threads()
    {
        if (!threads_.empty())
            return;
        threads_.reserve(helpers_);
        for (unsigned h = 0; h < helpers_; h++)
            threads_.emplace_back([this, h] { helper_main(h + 1); });
    }

    void helper_main(unsigned part)
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lock(mutex_);
        for (;;) {
            wake_.wait(lock, [&] { return quit_ || generation_ != seen; });
            if (quit_)
                return;
            seen = generation_;
            lock.unlock();
            copy_slice(part);
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
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, slice_ = 0;
};

} // namespace sxhost

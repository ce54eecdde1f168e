// par_copy.hpp -- memcpy split over a few threads, for the bounce copies of pageable callers.
//
// The *_host entry points move a pageable caller buffer (a numpy array, a std::vector: what the
// reference's readStream/writeStream callers pass as buffs[0], SoapySX.cpp:868-875, :969-976)
// through pinned staging, because the copy engines cannot read pageable memory.  One thread
// copies at 5-10 GB/s, a fifth of what the PCIe link behind it carries; a handful of threads
// close most of that gap.  Plain C++, no CUDA: tests/test_host_logic.py builds and runs it on
// the CPU, under ThreadSanitizer where the toolchain has it.
#pragma once

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define SXHOST_HAVE_NT_STORES 1
#endif

#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace sxhost {

// memcpy whose stores bypass the cache (x86: MOVNTDQ through the write-combining buffers).  A
// bounce copy writes a destination nobody reads soon -- the pinned staging the copy engine will
// fetch by DMA, or a caller buffer far larger than the caches -- so allocating its lines costs a
// read of the destination from DRAM for nothing: 3 bytes of memory traffic per byte copied
// instead of 2.  On a host whose memory bandwidth is what limits the pageable path, that is
// the difference.  Falls back to memcpy for small or oddly aligned pieces and off x86.
inline void stream_copy(void *dst, const void *src, size_t bytes)
{
#if defined(SXHOST_HAVE_NT_STORES)
    if (bytes >= 4096) {
        char *d = static_cast<char *>(dst);
        const char *s = static_cast<const char *>(src);
        const size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;
        if (head) {
            std::memcpy(d, s, head);
            d += head, s += head, bytes -= head;
        }
        const size_t lines = bytes / 64;
        for (size_t i = 0; i < lines; i++) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 32));
            const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(d), a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + 48), e);
            d += 64, s += 64;
        }
        _mm_sfence(); // the streamed lines are globally visible before anyone is told the copy is done
        std::memcpy(d, s, bytes - lines * 64);
        return;
    }
#endif
    std::memcpy(dst, src, bytes);
}

class ParallelCopier {
public:
    // Copies below this size are done by the caller alone: handing out pieces costs a few
    // microseconds, a single thread moves 1 MiB in about a hundred.
    static constexpr size_t kMinParallelBytes = size_t(2) << 20;
    // A helper that has finished its piece keeps looking for the next copy for this long before
    // it parks on the condition variable: the chunks of one *_host call follow each other within
    // tens of microseconds, and waking a parked thread costs about as much as copying a chunk.
    static constexpr unsigned kSpinsBeforeParking = 20000; // ~100 us of polling

    // `helpers` threads are started on first use and parked between copies.  `streaming`: use
    // cache-bypassing stores (stream_copy) for every piece.
    explicit ParallelCopier(unsigned helpers, bool streaming = false) : helpers_(helpers), streaming_(streaming) {}

    ~ParallelCopier()
    {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            quit_.store(true, std::memory_order_release);
        }
        wake_.notify_all();
        for (std::thread &t : threads_)
            t.join();
    }

    ParallelCopier(const ParallelCopier &) = delete;
    ParallelCopier &operator=(const ParallelCopier &) = delete;

    unsigned helpers() const { return helpers_; }
    bool streaming() const { return streaming_; }

    // memcpy(dst, src, bytes); returns when every byte has been copied.  One copy at a time
    // (the caller serialises: each pipeline side owns its copier).
    void copy(void *dst, const void *src, size_t bytes)
    {
        if (helpers_ == 0 || bytes < kMinParallelBytes) {
            move(dst, src, bytes);
            return;
        }
        start_threads();
        const unsigned parts = helpers_ + 1;
        // slices are multiples of 4 KiB so that no two threads share a page of the destination
        const size_t slice = ((bytes + parts - 1) / parts + 4095) & ~size_t(4095); // parts * slice >= bytes
        dst_ = static_cast<char *>(dst);
        src_ = static_cast<const char *>(src);
        bytes_ = bytes;
        slice_ = slice;
        pending_.store(helpers_, std::memory_order_relaxed);
        {
            // Published under the lock so that a helper about to park cannot miss it.
            std::lock_guard<std::mutex> lock(mutex_);
            generation_.fetch_add(1, std::memory_order_release);
        }
        wake_.notify_all();
        copy_slice(0); // the caller takes the first slice
        unsigned spins = 0;
        while (pending_.load(std::memory_order_acquire) != 0)
            if (++spins > 2000)
                std::this_thread::yield();
    }

private:
    void copy_slice(unsigned part) const
    {
        const size_t lo = size_t(part) * slice_;
        if (lo >= bytes_)
            return;
        move(dst_ + lo, src_ + lo, bytes_ - lo < slice_ ? bytes_ - lo : slice_);
    }

    void move(void *dst, const void *src, size_t bytes) const
    {
        if (streaming_)
            stream_copy(dst, src, bytes);
        else
            std::memcpy(dst, src, bytes);
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
        for (;;) {
            // Look for the next copy: poll for a while, then park.
            unsigned spins = 0;
            while (generation_.load(std::memory_order_acquire) == seen && !quit_.load(std::memory_order_acquire)) {
                if (++spins < kSpinsBeforeParking)
                    continue;
                std::unique_lock<std::mutex> lock(mutex_);
                wake_.wait(lock, [&] {
                    return quit_.load(std::memory_order_acquire) || generation_.load(std::memory_order_acquire) != seen;
                });
            }
            if (quit_.load(std::memory_order_acquire))
                return;
            seen = generation_.load(std::memory_order_acquire);
            copy_slice(part);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }

    const unsigned helpers_;
    const bool streaming_;
    std::vector<std::thread> threads_;
    std::mutex mutex_;
    std::condition_variable wake_;
    std::atomic<bool> quit_{false};
    std::atomic<uint64_t> generation_{0};
    std::atomic<unsigned> pending_{0};
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, slice_ = 0;
};

} // namespace sxhost

#include <atomic>
#include <functional>

namespace sxhost {

// One parked thread that runs one job at a time beside its owner: the *_host pipeline hands it
// "wait for each chunk's device-to-host copy, then bounce the chunk out to the caller's pageable
// buffer", so that those copies overlap the inbound bounce copies and the CUDA calls the owner
// is making for later chunks.  start() returns at once; finish() returns when the job has.
class Sidekick {
public:
    Sidekick() = default;
    ~Sidekick()
    {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            quit_.store(true, std::memory_order_release);
        }
        wake_.notify_all();
        if (thread_.joinable())
            thread_.join();
    }
    Sidekick(const Sidekick &) = delete;
    Sidekick &operator=(const Sidekick &) = delete;

    void start(std::function<void()> job)
    {
        if (!thread_.joinable())
            thread_ = std::thread([this] { main(); });
        job_ = std::move(job);
        {
            std::lock_guard<std::mutex> lock(mutex_); // so that a thread about to park cannot miss it
            started_.fetch_add(1, std::memory_order_release);
        }
        wake_.notify_all();
    }

    void finish()
    {
        unsigned spins = 0;
        while (finished_.load(std::memory_order_acquire) != started_.load(std::memory_order_relaxed))
            if (++spins > 2000)
                std::this_thread::yield();
    }

private:
    void main()
    {
        uint64_t seen = 0;
        for (;;) {
            // Calls follow each other closely while a stream is running: poll a while, then park.
            unsigned spins = 0;
            while (started_.load(std::memory_order_acquire) == seen && !quit_.load(std::memory_order_acquire)) {
                if (++spins < ParallelCopier::kSpinsBeforeParking)
                    continue;
                std::unique_lock<std::mutex> lock(mutex_);
                wake_.wait(lock, [&] {
                    return quit_.load(std::memory_order_acquire) || started_.load(std::memory_order_acquire) != seen;
                });
            }
            if (quit_.load(std::memory_order_acquire))
                return;
            seen = started_.load(std::memory_order_acquire);
            job_();
            finished_.store(seen, std::memory_order_release);
        }
    }

    std::thread thread_;
    std::mutex mutex_;
    std::condition_variable wake_;
    std::function<void()> job_;
    std::atomic<bool> quit_{false};
    std::atomic<uint64_t> started_{0}, finished_{0};
};

// Progress counter shared by two threads of a pipeline: one publishes "chunks 0..n-1 are done",
// the other waits for a given chunk.  Waiting spins briefly (a chunk takes tens of microseconds)
// and then yields.
class Progress {
public:
    void reset() { value_.store(0, std::memory_order_relaxed); }
    void publish(uint64_t n) { value_.store(n, std::memory_order_release); }
    uint64_t get() const { return value_.load(std::memory_order_acquire); }
    // Returns false if `abort` became non-zero while waiting.
    bool wait_for(uint64_t n, const std::atomic<int> &abort) const
    {
        unsigned spins = 0;
        while (value_.load(std::memory_order_acquire) < n) {
            if (abort.load(std::memory_order_relaxed))
                return false;
            if (++spins > 2000)
                std::this_thread::yield();
        }
        return true;
    }

private:
    std::atomic<uint64_t> value_{0};
};

// Chunk schedule of the copy-engine pipeline: small chunks first and last, so that the time
// during which only one direction of the link is busy (the first chunk's inbound copy, the last
// chunk's outbound copy) is short, and large chunks in between, where per-chunk overhead counts.
// Sizes double from c_min up to c_max and mirror at the end; every boundary is a multiple of 64
// frames so that 16-byte alignment carries over from the block start for every frame width
// (the up to 63 frames by which the length exceeds a multiple of 64 ride on the last chunk,
// which may therefore be that much longer than c_max).
struct ChunkSpan {
    size_t first, frames;
};

inline std::vector<ChunkSpan> plan_chunks(size_t length, size_t c_min, size_t c_max)
{
    auto round64 = [](size_t v) { return (v + 63) & ~size_t(63); };
    if (c_max < 64)
        c_max = 64;
    c_max = round64(c_max);
    c_min = (c_min == 0 || c_min > c_max) ? c_max : round64(c_min);
    std::vector<size_t> front, back;
    const size_t ragged = length % 64;
    size_t rem = length - ragged, c = c_min;
    while (c < c_max && rem >= 4 * c) {
        front.push_back(c);
        back.push_back(c);
        rem -= 2 * c;
        c *= 2;
    }
    if (c > c_max)
        c = c_max;
    std::vector<ChunkSpan> out;
    size_t at = 0;
    for (size_t f : front) {
        out.push_back({at, f});
        at += f;
    }
    if (rem > 0) {
        const size_t n_mid = (rem + c - 1) / c;
        const size_t each = round64((rem + n_mid - 1) / n_mid);
        size_t left = rem;
        while (left > 0) {
            const size_t take = left < each ? left : each;
            out.push_back({at, take});
            at += take;
            left -= take;
        }
    }
    for (size_t i = back.size(); i-- > 0;) {
        out.push_back({at, back[i]});
        at += back[i];
    }
    if (ragged) {
        if (out.empty())
            out.push_back({0, ragged});
        else
            out.back().frames += ragged;
    }
    return out;
}

} // namespace sxhost

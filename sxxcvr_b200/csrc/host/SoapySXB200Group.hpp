// SoapySXB200Group -- N SX1255 front-ends served as one device: one conversion per period for
// all of them.
//
// A SoapySX device converts one period (256 frames, 2 KiB, SoapySX.cpp:451) per readStream or
// writeStream call; on a GPU that call is a launch-and-wait of ~15 us for 2 KiB of work, six
// times what the reference's CPU loop needs.  The way to serve many front-ends from one B200 is
// not N such devices but this: every member keeps its own PCM pair, frame counters, timestamps
// and error codes -- the bookkeeping of readStream / writeStream is applied to each member by
// exactly the code the single device uses (stream_ops.hpp), spread over a few host threads --
// while their blocks sit side by side in one pinned staging buffer and are converted by ONE
// kernel: RX for all members, TX for all members, or, for the repeater pattern
// (example/linear_repeater.py:50-71), RX and TX of all members fused in one launch.
//
// Layout: member i's block is frames [i * numElems, (i + 1) * numElems) of every buffer
// (the I2S staging buffers inside, the caller's CF32 buffer outside).
#pragma once

#include <SoapySDR/Types.hpp>

#include <cstddef>
#include <memory>
#include <string>
#include <vector>

#include "SoapySXB200.hpp"
#include "stream_ops.hpp"
#include "worker_pool.hpp"

namespace sxhost {

class SoapySXB200Group {
public:
    // args: gpu=N, clock=32e6|38.4e6, period=frames (default 256), threshold=|z| that keys the PA
    // (default 1e-3), threads=host threads for the bookkeeping (default: half the hardware's).
    SoapySXB200Group(size_t members, const SoapySDR::Kwargs &args);
    ~SoapySXB200Group();
    SoapySXB200Group(const SoapySXB200Group &) = delete;
    SoapySXB200Group &operator=(const SoapySXB200Group &) = delete;

    size_t size() const { return members_.size(); }
    size_t period() const { return period_; }
    void setSampleRate(double rate); // one of the SX1255's rates (listSampleRates of the single device)
    double getSampleRate() const { return sample_rate_; }
    int activate();   // both directions of every member (activateStream, SoapySX.cpp:803-830)
    int deactivate(); // deactivateStream of both: PCMs stopped, counters back to 0 (:832-859)

    // readStream(numElems) on every member, one conversion.  cf32: [members][numElems] CF32 in
    // host (pageable or pinned) or device memory.  Per member: ret (frames, 0, or a negative
    // SOAPY_SDR_* code), flags, timeNs.  A member that delivered fewer than numElems frames
    // leaves the rest of its block unspecified.  Returns 0, or SOAPY_SDR_STREAM_ERROR if the
    // conversion itself failed.
    int readAll(void *cf32, size_t numElems, int *rets, int *flags, long long *timeNs, long timeoutUs);
    // writeStream(numElems, flags[i], timeNs[i]) on every member, one conversion.
    int writeAll(const void *cf32, size_t numElems, const int *flags, const long long *timeNs, int *rets,
                 long timeoutUs);
    // The repeater iteration: readStream(numElems) on every member, then
    // writeStream(numElems, HAS_TIME, that member's rx time + offset_ns) of the block just read,
    // RX and TX conversions of all members fused in one launch.  cf32 (may be null): where the
    // CF32 blocks are left, as the application would see them between the two calls.
    int repeatAll(void *cf32, size_t numElems, long long offset_ns, int *rx_rets, int *tx_rets,
                  long long *rx_timeNs, long timeoutUs);

    snd_pcm_t *pcm(size_t member, bool capture) const;
    sxgpu_ctx *gpu() const { return gpu_; }

private:
    struct Member {
        Endpoint rx, tx;
        Member() : rx("hw:CARD=SX1255,DEV=1", SND_PCM_STREAM_CAPTURE), tx("hw:CARD=SX1255,DEV=0", SND_PCM_STREAM_PLAYBACK) {}
    };
    void reserve(size_t numElems);

    sxgpu_ctx *gpu_ = nullptr;
    double master_clock_ = 38.4e6, sample_rate_ = 0;
    size_t period_ = 256;
    float tx_threshold2_ = 0.0f;
    std::vector<std::unique_ptr<Member>> members_;
    std::unique_ptr<WorkerPool> pool_;
    void *stage_rx_ = nullptr, *stage_tx_ = nullptr, *stage_cf_ = nullptr; // pinned, [members][capacity_]
    void *dev_in_ = nullptr, *dev_cf_ = nullptr, *dev_out_ = nullptr;       // device, large groups only
    size_t capacity_ = 0;
    std::vector<long long> rx_time_scratch_;
    std::vector<TxOutcome> tx_plan_; // per member, between the two halves of a write
};

} // namespace sxhost

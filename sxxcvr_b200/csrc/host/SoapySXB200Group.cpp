// SoapySXB200Group.cpp -- see SoapySXB200Group.hpp.  Host-side only: every sample conversion is
// a call into the sxgpu C ABI; there is no CPU conversion path.
#include "SoapySXB200Group.hpp"

#include <SoapySDR/Logger.hpp>

#include "stream_ops.hpp"
#include "sxgpu.h"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace sxhost {

namespace {
std::string arg(const SoapySDR::Kwargs &args, const char *key, const std::string &fallback)
{
    auto it = args.find(key);
    return it == args.end() ? fallback : it->second;
}
// A member's bookkeeping is ~1 us of ALSA calls: below this many members per thread a wake-up
// costs more than it saves.
constexpr size_t kMembersPerThread = 32;
// Up to this many frames in all members' blocks together the fused kernel works straight on the
// pinned staging buffers; above, the copy engines carry the blocks to device memory and back
// (crossover of the *_host entry points, profiles/r01_sweep_host_path.json).
constexpr size_t kZeroCopyFrames = size_t(1) << 17;
} // namespace

SoapySXB200Group::SoapySXB200Group(size_t members, const SoapySDR::Kwargs &args)
{
    if (members == 0)
        throw std::runtime_error("SoapySXB200Group: a group needs at least one member");
    const char *local_rank = std::getenv("LOCAL_RANK");
    const int ordinal = std::atoi(arg(args, "gpu", local_rank ? local_rank : "0").c_str());
    int rc = sxgpu_init(ordinal, &gpu_);
    if (rc != SXGPU_OK)
        throw std::runtime_error(std::string("SoapySXB200Group: cannot use GPU ") + std::to_string(ordinal) + ": " +
                                 sxgpu_strerror(rc) + " (there is no CPU fallback for the sample path)");
    try {
        master_clock_ = std::stod(arg(args, "clock", "38.4e6"));
        sample_rate_ = master_clock_ / 256.0;
        const float threshold = std::stof(arg(args, "threshold", "1e-3"));
        tx_threshold2_ = threshold * threshold;
        const unsigned long requested_period = std::stoul(arg(args, "period", "0"));
        unsigned threads = unsigned(std::stoul(arg(args, "threads", "0")));
        if (threads == 0)
            threads = std::max(1u, std::thread::hardware_concurrency() / 2);
        pool_ = std::make_unique<WorkerPool>(threads - 1);

        members_.reserve(members);
        for (size_t i = 0; i < members; i++) {
            auto m = std::make_unique<Member>();
            m->rx.open();
            m->tx.open();
            // setupStream of both directions in the default stream mode (SoapySX.cpp:740-794) ...
            m->rx.mode = m->tx.mode = Endpoint::Mode::Normal;
            m->rx.configure(requested_period);
            m->tx.configure(requested_period);
            m->rx.configured = m->tx.configured = true;
            // ... which links the two PCMs: one sample clock, one counter origin (:784-788)
            if (snd_pcm_link(m->rx.pcm, m->tx.pcm) < 0)
                throw std::runtime_error("ALSA error");
            members_.push_back(std::move(m));
        }
        period_ = members_[0]->rx.ring.period;
        tx_plan_.resize(members);
    } catch (...) {
        members_.clear();
        sxgpu_destroy(gpu_);
        throw;
    }
}

SoapySXB200Group::~SoapySXB200Group()
{
    members_.clear();
    for (void *p : {stage_rx_, stage_tx_, stage_cf_})
        if (p)
            sxgpu_free_host(gpu_, p);
    for (void *p : {dev_in_, dev_cf_, dev_out_})
        if (p)
            sxgpu_free(gpu_, p);
    sxgpu_destroy(gpu_);
}

void SoapySXB200Group::setSampleRate(double rate)
{
    if (!(rate > 0))
        throw std::runtime_error("Sample rate must be positive");
    const double divider = std::round(master_clock_ / rate);
    for (unsigned div : {1536u, 768u, 512u, 256u, 128u, 64u}) { // reference table, SoapySX.cpp:196-208
        if (double(div) == divider) {
            sample_rate_ = master_clock_ / divider;
            return;
        }
    }
    throw std::runtime_error("Unsupported sample rate");
}

int SoapySXB200Group::activate()
{
    int worst = 0;
    for (auto &m : members_) {
        for (Endpoint *ep : {&m->rx, &m->tx}) {
            if (ep->active) {
                worst = SOAPY_SDR_STREAM_ERROR;
                continue;
            }
            ep->active = true;
            if (snd_pcm_state(ep->pcm) == SND_PCM_STATE_PREPARED && snd_pcm_start(ep->pcm) < 0)
                worst = SOAPY_SDR_STREAM_ERROR;
        }
    }
    return worst;
}

int SoapySXB200Group::deactivate()
{
    int worst = 0;
    for (auto &m : members_) {
        if (!m->rx.active || !m->tx.active)
            worst = SOAPY_SDR_STREAM_ERROR;
        m->rx.active = m->tx.active = false;
        if (m->rx.reset() < 0 || m->tx.reset() < 0)
            worst = SOAPY_SDR_STREAM_ERROR;
    }
    return worst;
}

snd_pcm_t *SoapySXB200Group::pcm(size_t member, bool capture) const
{
    if (member >= members_.size())
        return nullptr;
    return capture ? members_[member]->rx.pcm : members_[member]->tx.pcm;
}

void SoapySXB200Group::reserve(size_t numElems)
{
    if (numElems <= capacity_)
        return;
    for (void **p : {&stage_rx_, &stage_tx_, &stage_cf_}) {
        if (*p)
            sxgpu_free_host(gpu_, *p);
        *p = nullptr;
    }
    for (void **p : {&dev_in_, &dev_cf_, &dev_out_}) {
        if (*p)
            sxgpu_free(gpu_, *p);
        *p = nullptr;
    }
    capacity_ = 0;
    const size_t bytes = members_.size() * numElems * 8;
    if (sxgpu_malloc_host(gpu_, &stage_rx_, bytes) != SXGPU_OK || sxgpu_malloc_host(gpu_, &stage_tx_, bytes) != SXGPU_OK ||
        sxgpu_malloc_host(gpu_, &stage_cf_, bytes) != SXGPU_OK)
        throw std::runtime_error(std::string("pinned staging allocation failed: ") + sxgpu_last_error(gpu_));
    if (members_.size() * numElems > kZeroCopyFrames &&
        (sxgpu_malloc(gpu_, &dev_in_, bytes) != SXGPU_OK || sxgpu_malloc(gpu_, &dev_cf_, bytes) != SXGPU_OK ||
         sxgpu_malloc(gpu_, &dev_out_, bytes) != SXGPU_OK))
        throw std::runtime_error(std::string("device staging allocation failed: ") + sxgpu_last_error(gpu_));
    capacity_ = numElems;
}

int SoapySXB200Group::readAll(void *cf32, size_t numElems, int *rets, int *flags, long long *timeNs, long timeoutUs)
{
    reserve(numElems);
    char *stage = static_cast<char *>(stage_rx_);
    pool_->run(members_.size(), kMembersPerThread, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            Endpoint &ep = members_[i]->rx;
            std::scoped_lock lock(ep.mutex);
            const RxOutcome rx = rx_before_convert(ep, sample_rate_, numElems, timeoutUs,
                                                   [&](size_t) { return stage + i * numElems * 8; });
            rets[i] = rx.ret;
            flags[i] = rx.flags;
            if (rx.time_valid)
                timeNs[i] = rx.time_ns;
        }
    });
    // One conversion for every member's block: staging [members][numElems] -> cf32 [members][numElems].
    if (sxgpu_convert_rx_buffer_host(gpu_, stage_rx_, 0, cf32, 0, members_.size() * numElems) != SXGPU_OK) {
        SoapySDR_logf(SOAPY_SDR_ERROR, "group rx GPU conversion failed: %s", sxgpu_last_error(gpu_));
        return SOAPY_SDR_STREAM_ERROR;
    }
    return 0;
}

int SoapySXB200Group::writeAll(const void *cf32, size_t numElems, const int *flags, const long long *timeNs, int *rets,
                               long timeoutUs)
{
    reserve(numElems);
    pool_->run(members_.size(), kMembersPerThread, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            Endpoint &ep = members_[i]->tx;
            std::scoped_lock lock(ep.mutex);
            tx_plan_[i] = tx_before_convert(ep, sample_rate_, numElems, flags ? flags[i] : 0, timeNs ? timeNs[i] : 0,
                                            timeoutUs);
        }
    });
    if (sxgpu_convert_tx_buffer_host(gpu_, cf32, 0, stage_tx_, 0, members_.size() * numElems, tx_threshold2_) != SXGPU_OK) {
        SoapySDR_logf(SOAPY_SDR_ERROR, "group tx GPU conversion failed: %s", sxgpu_last_error(gpu_));
        return SOAPY_SDR_STREAM_ERROR;
    }
    const char *stage = static_cast<const char *>(stage_tx_);
    pool_->run(members_.size(), kMembersPerThread, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            Endpoint &ep = members_[i]->tx;
            std::scoped_lock lock(ep.mutex);
            rets[i] = tx_plan_[i].convert ? tx_after_convert(ep, stage + i * numElems * 8, tx_plan_[i].length)
                                          : tx_plan_[i].ret;
        }
    });
    return 0;
}

int SoapySXB200Group::repeatAll(void *cf32, size_t numElems, long long offset_ns, int *rx_rets, int *tx_rets,
                                long long *rx_timeNs, long timeoutUs)
{
    reserve(numElems);
    char *in = static_cast<char *>(stage_rx_);
    const size_t total = members_.size() * numElems;
    std::vector<long long> &t_rx = rx_time_scratch_;
    t_rx.assign(members_.size(), 0);

    // 1. The read half of the bookkeeping: every member's frames land side by side in pinned staging.
    pool_->run(members_.size(), kMembersPerThread, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            Endpoint &ep = members_[i]->rx;
            std::scoped_lock lock(ep.mutex);
            const RxOutcome rx = rx_before_convert(ep, sample_rate_, numElems, timeoutUs,
                                                   [&](size_t) { return in + i * numElems * 8; });
            rx_rets[i] = rx.ret;
            t_rx[i] = rx.time_ns;
            if (rx_timeNs && rx.time_valid)
                rx_timeNs[i] = rx.time_ns;
        }
    });

    // 2. RX and TX conversions of every member in ONE launch, queued without waiting.  Small groups:
    // the kernel reads and writes the pinned staging buffers across PCIe itself.  Large groups:
    // one copy in, the fused kernel on device buffers, one copy out (the copy engines move 8 MiB in
    // a fifth of the time the SMs need to fetch it across the link).
    int rc;
    if (total <= kZeroCopyFrames) {
        rc = sxgpu_convert_loopback(gpu_, stage_rx_, stage_cf_, stage_tx_, total, tx_threshold2_, nullptr);
    } else {
        rc = sxgpu_memcpy_h2d(gpu_, dev_in_, stage_rx_, total * 8, nullptr);
        if (rc == SXGPU_OK)
            rc = sxgpu_convert_loopback(gpu_, dev_in_, cf32 ? dev_cf_ : nullptr, dev_out_, total, tx_threshold2_, nullptr);
        if (rc == SXGPU_OK)
            rc = sxgpu_memcpy_d2h(gpu_, stage_tx_, dev_out_, total * 8, nullptr);
        if (rc == SXGPU_OK && cf32)
            rc = sxgpu_memcpy_d2h(gpu_, stage_cf_, dev_cf_, total * 8, nullptr);
    }

    // 3. While the GPU works: where each member's timed write lands (this read's timestamp +
    // offset), the forward over the gap, the trim -- none of which needs the samples.  A member
    // whose read did not deliver a whole block writes nothing (the application would skip its
    // write, linear_repeater.py:59-61).
    pool_->run(members_.size(), kMembersPerThread, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            tx_plan_[i] = TxOutcome();
            if (rx_rets[i] != int(numElems))
                continue;
            Endpoint &ep = members_[i]->tx;
            std::scoped_lock lock(ep.mutex);
            tx_plan_[i] = tx_before_convert(ep, sample_rate_, numElems, SOAPY_SDR_HAS_TIME, t_rx[i] + offset_ns, timeoutUs);
        }
    });

    if (rc != SXGPU_OK || sxgpu_stream_sync(gpu_, nullptr) != SXGPU_OK) {
        SoapySDR_logf(SOAPY_SDR_ERROR, "group repeat GPU conversion failed: %s", sxgpu_last_error(gpu_));
        return SOAPY_SDR_STREAM_ERROR;
    }
    // 4. Hand every member's I2S frames to its PCM -- and its CF32 block to the caller, by the same
    // threads (one memcpy of the whole buffer on the calling thread was a third of the iteration at
    // 4096 members).
    const char *out = static_cast<const char *>(stage_tx_);
    const char *cf_stage = static_cast<const char *>(stage_cf_);
    pool_->run(members_.size(), kMembersPerThread, [&](size_t lo, size_t hi) {
        if (cf32)
            std::memcpy(static_cast<char *>(cf32) + lo * numElems * 8, cf_stage + lo * numElems * 8, (hi - lo) * numElems * 8);
        for (size_t i = lo; i < hi; i++) {
            Endpoint &ep = members_[i]->tx;
            std::scoped_lock lock(ep.mutex);
            tx_rets[i] = tx_plan_[i].convert ? tx_after_convert(ep, out + i * numElems * 8, tx_plan_[i].length)
                                             : tx_plan_[i].ret;
        }
    });
    return 0;
}

} // namespace sxhost

// ---------------------------------------------------------------------------------------
// Flat C view (sxg_*), for callers without C++: tests and bench.py drive it through ctypes.
// ---------------------------------------------------------------------------------------
namespace {
thread_local std::string g_group_error;
template <typename F> int group_guarded(F &&body)
{
    try {
        return body();
    } catch (const std::exception &e) {
        g_group_error = e.what();
    } catch (...) {
        g_group_error = "unknown exception";
    }
    return -1000;
}
} // namespace

extern "C" {

const char *sxg_last_error(void) { return g_group_error.c_str(); }

int sxg_create(size_t members, const char *args, sxhost::SoapySXB200Group **out)
{
    *out = nullptr;
    return group_guarded([&] {
        *out = new sxhost::SoapySXB200Group(members, SoapySDR::KwargsFromString(args ? args : ""));
        return 0;
    });
}
int sxg_destroy(sxhost::SoapySXB200Group *g)
{
    return group_guarded([&] {
        delete g;
        return 0;
    });
}
size_t sxg_size(sxhost::SoapySXB200Group *g) { return g->size(); }
size_t sxg_period(sxhost::SoapySXB200Group *g) { return g->period(); }
int sxg_set_sample_rate(sxhost::SoapySXB200Group *g, double rate)
{
    return group_guarded([&] {
        g->setSampleRate(rate);
        return 0;
    });
}
int sxg_activate(sxhost::SoapySXB200Group *g) { return group_guarded([&] { return g->activate(); }); }
int sxg_deactivate(sxhost::SoapySXB200Group *g) { return group_guarded([&] { return g->deactivate(); }); }
snd_pcm_t *sxg_pcm(sxhost::SoapySXB200Group *g, size_t member, int capture) { return g->pcm(member, capture != 0); }
int sxg_read_all(sxhost::SoapySXB200Group *g, void *cf32, size_t numElems, int *rets, int *flags, long long *timeNs,
                 long timeoutUs)
{
    return group_guarded([&] { return g->readAll(cf32, numElems, rets, flags, timeNs, timeoutUs); });
}
int sxg_write_all(sxhost::SoapySXB200Group *g, const void *cf32, size_t numElems, const int *flags,
                  const long long *timeNs, int *rets, long timeoutUs)
{
    return group_guarded([&] { return g->writeAll(cf32, numElems, flags, timeNs, rets, timeoutUs); });
}
int sxg_repeat_all(sxhost::SoapySXB200Group *g, void *cf32, size_t numElems, long long offset_ns, int *rx_rets,
                   int *tx_rets, long long *rx_timeNs, long timeoutUs)
{
    return group_guarded([&] { return g->repeatAll(cf32, numElems, offset_ns, rx_rets, tx_rets, rx_timeNs, timeoutUs); });
}
// `iters` repeater iterations timed natively (steady clock); returns 0 and the seconds, or the
// first unexpected per-member return value.
int sxg_bench_repeat(sxhost::SoapySXB200Group *g, size_t numElems, long long offset_ns, int iters, double *seconds)
{
    return group_guarded([&] {
        std::vector<int> rx(g->size()), tx(g->size());
        std::vector<long long> t(g->size());
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < iters; k++) {
            int rc = g->repeatAll(nullptr, numElems, offset_ns, rx.data(), tx.data(), t.data(), 1000000);
            if (rc != 0)
                return rc;
            for (size_t i = 0; i < g->size(); i++)
                if (rx[i] != int(numElems) || tx[i] != int(numElems))
                    return rx[i] != int(numElems) ? (rx[i] < 0 ? rx[i] : -2000) : (tx[i] < 0 ? tx[i] : -2001);
        }
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return 0;
    });
}

} // extern "C"

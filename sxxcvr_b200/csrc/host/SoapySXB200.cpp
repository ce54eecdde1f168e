// SoapySXB200.cpp -- see SoapySXB200.hpp.  Host-side only: every sample conversion in this
// file is a call into the sxgpu C ABI; there is no CPU conversion path.
#include "SoapySXB200.hpp"

#include <SoapySDR/Formats.h>
#include <SoapySDR/Logger.hpp>
#include <SoapySDR/Registry.hpp>
#include <SoapySDR/Time.hpp>

#include "sxgpu.h"
#include "stream_ops.hpp"

#include <algorithm>
#include <cerrno>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <stdexcept>

namespace sxhost {

// Blocking reads of this many frames and more run their conversion under the read itself (see
// readStream); the read then proceeds in pieces of kGatedReadPiece frames.
constexpr size_t kGatedReadFrames = size_t(1) << 22, kGatedReadPiece = size_t(1) << 20;

namespace {

// SX1255 sample-rate dividers that work over I2S with 32-bit slots (reference table,
// SoapySX.cpp:196-208; the 24- and 16-bit entries are commented out there as broken).
const unsigned kRateDividers[] = {1536, 768, 512, 256, 128, 64};

std::string kwarg(const SoapySDR::Kwargs &args, const char *key, const std::string &fallback)
{
    auto it = args.find(key);
    return it == args.end() ? fallback : it->second;
}

int check_alsa(int ret, const char *what)
{
    if (ret < 0)
        SoapySDR_logf(SOAPY_SDR_ERROR, "ALSA error in %s: %s", what, snd_strerror(ret));
    return ret;
}

} // namespace

// ---------------------------------------------------------------------------------------
// Endpoint
// ---------------------------------------------------------------------------------------
Endpoint::~Endpoint()
{
    if (pcm)
        snd_pcm_close(pcm);
}

void Endpoint::open()
{
    if (check_alsa(snd_pcm_open(&pcm, pcm_name, direction, 0), "snd_pcm_open") < 0) {
        pcm = nullptr;
        throw std::runtime_error("Error opening ALSA device");
    }
}

// Stops the stream, rewinds the frame counter and forgets queued frames (reference :419-432).
int Endpoint::reset()
{
    snd_pcm_drop(pcm); // fails harmlessly when already stopped
    int ret = snd_pcm_prepare(pcm);
    if (ret < 0)
        return ret;
    position = 0;
    return snd_pcm_reset(pcm);
}

// S32_LE, two interleaved channels, the largest ring the I2S DMA allows for the requested
// period; software parameters depend on the stream mode (reference :434-517).
void Endpoint::configure(unsigned long requested_period)
{
    if (!pcm)
        return;
    ring = sxplan::geometry_for_period(requested_period);

    snd_pcm_hw_params_t *hw = nullptr;
    snd_pcm_sw_params_t *sw = nullptr;
    auto fail = [&]() {
        if (hw)
            snd_pcm_hw_params_free(hw);
        if (sw)
            snd_pcm_sw_params_free(sw);
        throw std::runtime_error("Error configuring ALSA device");
    };
#define SX_ALSA(call)                                                                          \
    if (check_alsa((call), #call) < 0)                                                         \
    fail()

    unsigned periods = 0;
    SX_ALSA(snd_pcm_hw_params_malloc(&hw));
    SX_ALSA(snd_pcm_hw_params_any(pcm, hw));
    SX_ALSA(snd_pcm_hw_params_set_access(pcm, hw, SND_PCM_ACCESS_RW_INTERLEAVED));
    SX_ALSA(snd_pcm_hw_params_set_format(pcm, hw, SND_PCM_FORMAT_S32_LE));
    // The I2S clock comes from the SX1255, not from ALSA; the rate given here is a dummy.
    SX_ALSA(snd_pcm_hw_params_set_rate(pcm, hw, 192000, 0));
    SX_ALSA(snd_pcm_hw_params_set_channels(pcm, hw, 2));
    SX_ALSA(snd_pcm_hw_params_set_buffer_size_near(pcm, hw, &ring.buffer));
    SX_ALSA(snd_pcm_hw_params_set_period_size_near(pcm, hw, &ring.period, 0));
    SX_ALSA(snd_pcm_hw_params_get_periods(hw, &periods, 0));
    SX_ALSA(snd_pcm_hw_params(pcm, hw));
    snd_pcm_hw_params_free(hw);
    hw = nullptr;
    SoapySDR_logf(SOAPY_SDR_DEBUG, "I2S ring: %lu frames in %u periods of %lu", ring.buffer,
                  periods, ring.period);

    snd_pcm_uframes_t boundary = 0;
    SX_ALSA(snd_pcm_sw_params_malloc(&sw));
    SX_ALSA(snd_pcm_sw_params_current(pcm, sw));
    SX_ALSA(snd_pcm_sw_params_get_boundary(sw, &boundary));
    if (mode == Mode::Normal) {
        // Never stop on xrun; gaps in playback are played as silence.
        SX_ALSA(snd_pcm_sw_params_set_stop_threshold(pcm, sw, boundary));
        SX_ALSA(snd_pcm_sw_params_set_silence_threshold(pcm, sw, 0));
        SX_ALSA(snd_pcm_sw_params_set_silence_size(pcm, sw, boundary));
    } else {
        // Plain ALSA behaviour: an xrun stops the (linked) streams.
        SX_ALSA(snd_pcm_sw_params_set_stop_threshold(pcm, sw, ring.buffer));
        SX_ALSA(snd_pcm_sw_params_set_silence_threshold(pcm, sw, 0));
        SX_ALSA(snd_pcm_sw_params_set_silence_size(pcm, sw, 0));
    }
    SX_ALSA(snd_pcm_sw_params(pcm, sw));
    snd_pcm_sw_params_free(sw);
    sw = nullptr;

    SX_ALSA(reset());
#undef SX_ALSA
}

// ---------------------------------------------------------------------------------------
// PinnedFrames
// ---------------------------------------------------------------------------------------
PinnedFrames::~PinnedFrames()
{
    if (data_)
        sxgpu_free_host(gpu_, data_);
}

void PinnedFrames::reserve(size_t frames)
{
    if (frames <= capacity_)
        return;
    // Contents never need to survive: both directions fill the staging buffer afresh on
    // every call (reference :944-948, :1087-1093).
    if (data_)
        sxgpu_free_host(gpu_, data_);
    data_ = nullptr;
    capacity_ = 0;
    if (sxgpu_malloc_host(gpu_, &data_, frames * sizeof(uint64_t)) != SXGPU_OK)
        throw std::runtime_error(std::string("pinned staging allocation failed: ") +
                                 sxgpu_last_error(gpu_));
    capacity_ = frames;
}

// ---------------------------------------------------------------------------------------
// Construction
// ---------------------------------------------------------------------------------------
SoapySXB200::SoapySXB200(const SoapySDR::Kwargs &args)
    : rx_("hw:CARD=SX1255,DEV=1", SND_PCM_STREAM_CAPTURE),
      tx_("hw:CARD=SX1255,DEV=0", SND_PCM_STREAM_PLAYBACK)
{
    SoapySDR_logf(SOAPY_SDR_INFO, "Initializing SoapySX (B200 stream path)");

    // Which GPU: device argument gpu=N, else LOCAL_RANK (one process per GPU), else 0.
    const char *local_rank = std::getenv("LOCAL_RANK");
    gpu_ordinal_ = std::atoi(kwarg(args, "gpu", local_rank ? local_rank : "0").c_str());
    int rc = sxgpu_init(gpu_ordinal_, &gpu_);
    if (rc != SXGPU_OK)
        throw std::runtime_error(std::string("SoapySXB200: cannot use GPU ") +
                                 std::to_string(gpu_ordinal_) + ": " + sxgpu_strerror(rc) +
                                 " (there is no CPU fallback for the sample path)");

    // With no SX1255 to probe, start from what the reference concludes when its clock
    // detection is inconclusive: 38.4 MHz (SoapySX.cpp:656-659).  clock=32e6 selects the
    // other board variant.  The initial rate is masterClock/256 (:662).
    cs16_enabled_ = kwarg(args, "cs16", "0") == "1";
    // Period-sized blocks are converted by a resident kernel that is rung through a doorbell in
    // pinned memory -- no launch, no stream sync per call (csrc/sx_resident.cuh).  By default for
    // calls of up to 1024 frames, where it saves a third of the call (the reference's period is
    // 256, SoapySX.cpp:451); lowlatency=1 extends it to 4096 frames, lowlatency=0 turns it off
    // (e.g. when the GPU is time-sliced between processes: the kernel holds its slot for up to
    // 20 ms at a time while a stream is calling in).
    const std::string lowlatency = kwarg(args, "lowlatency", "");
    if (lowlatency != "0")
        sxgpu_set_option(gpu_, "resident_max_frames", lowlatency == "1" ? 4096 : 1024);
    // sxgpu.<option>=<integer>: any tuning option of the C ABI (include/sxgpu.h), e.g.
    // sxgpu.bounce_threads=2 when several devices share the host's cores.
    for (const auto &kv : args) {
        if (kv.first.rfind("sxgpu.", 0) != 0)
            continue;
        if (sxgpu_set_option(gpu_, kv.first.c_str() + 6, std::atoll(kv.second.c_str())) != SXGPU_OK) {
            const std::string why = sxgpu_last_error(gpu_);
            sxgpu_destroy(gpu_);
            throw std::runtime_error("SoapySXB200: bad device argument " + kv.first + ": " + why);
        }
    }
    master_clock_ = std::stod(kwarg(args, "clock", "38.4e6"));
    sample_rate_ = master_clock_ / 256.0;
    antenna_[SOAPY_SDR_RX] = "RX";
    antenna_[SOAPY_SDR_TX] = "TX";
    setFrequency(SOAPY_SDR_RX, 0, 433.92e6, SoapySDR::Kwargs());
    setFrequency(SOAPY_SDR_TX, 0, 433.92e6, SoapySDR::Kwargs());

    overlap_reads_ = kwarg(args, "overlap", "1") != "0";
    stage_rx_ = std::make_unique<PinnedFrames>(gpu_);
    stage_tx_ = std::make_unique<PinnedFrames>(gpu_);
    try {
        // Same initial staging size as the reference (:704-707).
        stage_rx_->reserve(8192);
        stage_tx_->reserve(8192);
        rx_.open();
        tx_.open();
    } catch (...) {
        stage_rx_.reset(); // pinned memory goes back before the context that owns it
        stage_tx_.reset();
        sxgpu_destroy(gpu_);
        throw;
    }
}

SoapySXB200::~SoapySXB200()
{
    SoapySDR_logf(SOAPY_SDR_INFO, "Uninitializing SoapySX (B200 stream path)");
    unpin_all(rx_);
    unpin_all(tx_);
    rx_helper_.reset(); // its thread is idle between calls; gone before the context it calls into
    stage_rx_.reset();  // pinned memory goes back before the context that owns it
    stage_tx_.reset();
    sxgpu_destroy(gpu_);
}

SoapySDR::Kwargs SoapySXB200::getHardwareInfo() const
{
    SoapySDR::Kwargs info;
    info["soapysx_tag"] = "sxxcvr-b200";
    info["soapysx_commit"] = "abi" + std::to_string(sxgpu_abi_version());
    info["hardware_version"] = "unknown"; // no HAT EEPROM on a GPU box (reference :1582-1587)
    sxgpu_info gi;
    if (sxgpu_device_info(gpu_, &gi) == SXGPU_OK) {
        info["gpu"] = gi.name;
        info["gpu_ordinal"] = std::to_string(gi.device);
        info["gpu_sm_count"] = std::to_string(gi.sm_count);
    }
    return info;
}

// ---------------------------------------------------------------------------------------
// Stream formats
// ---------------------------------------------------------------------------------------
std::vector<std::string> SoapySXB200::getStreamFormats(const int, const size_t) const
{
    // Reference :1610-1616 offers CF32 only.  CS16 is an extension with no reference behaviour
    // (DESIGN.md), listed only when the device was opened with cs16=1.
    if (cs16_enabled_)
        return std::vector<std::string>{SOAPY_SDR_CF32, SOAPY_SDR_CS16};
    return std::vector<std::string>{SOAPY_SDR_CF32};
}

std::string SoapySXB200::getNativeStreamFormat(const int, const size_t, double &fullScale) const
{
    fullScale = 1.0; // reference :1597-1608
    return SOAPY_SDR_CF32;
}

// ---------------------------------------------------------------------------------------
// Stream lifecycle
// ---------------------------------------------------------------------------------------
SoapySDR::Stream *SoapySXB200::setupStream(const int direction, const std::string &format,
                                           const std::vector<size_t> &,
                                           const SoapySDR::Kwargs &args)
{
    std::scoped_lock lock(rx_.mutex, tx_.mutex);

    const bool want_cs16 = cs16_enabled_ && format == SOAPY_SDR_CS16;
    if (format != SOAPY_SDR_CF32 && !want_cs16)
        throw std::runtime_error("Only CF32 format is currently supported");
    if (snd_pcm_state(rx_.pcm) == SND_PCM_STATE_RUNNING ||
        snd_pcm_state(tx_.pcm) == SND_PCM_STATE_RUNNING)
        throw std::runtime_error("Streams can be setup only if none of the streams are running");

    Endpoint &ep = (direction == SOAPY_SDR_RX) ? rx_ : tx_;
    if (ep.configured)
        throw std::runtime_error("Stream has been setup already");

    if (ep.is_tx()) {
        // The PA is keyed by the samples themselves: a frame whose |z| reaches `threshold`
        // carries the TX-enable bits.  Stored squared, in single precision (reference :766-774).
        float threshold = args.count("threshold") ? std::stof(args.at("threshold")) : 1.0e-3f;
        tx_threshold2_ = threshold * threshold;
    }

    ep.cs16 = want_cs16;
    ep.pin_caller_buffers = kwarg(args, "pin", "0") == "1";
    ep.mode = (kwarg(args, "link", "") == "1") ? Endpoint::Mode::Linked : Endpoint::Mode::Normal;
    ep.configure(args.count("period") ? std::stoul(args.at("period")) : 0);
    ep.configured = true;

    // Once both directions exist they share one sample clock: same start, same counter origin.
    if (!linked_ && rx_.configured && tx_.configured) {
        SoapySDR_logf(SOAPY_SDR_DEBUG, "Linking RX and TX PCMs");
        if (check_alsa(snd_pcm_link(rx_.pcm, tx_.pcm), "snd_pcm_link") < 0)
            throw std::runtime_error("ALSA error");
        linked_ = true;
    }
    return reinterpret_cast<SoapySDR::Stream *>(&ep);
}

void SoapySXB200::closeStream(SoapySDR::Stream *stream)
{
    Endpoint *ep = endpoint_of(stream);
    std::scoped_lock lock(ep->mutex);
    ep->configured = false;
    unpin_all(*ep);
}

// Stream argument pin=1 (ours; the reference has nothing like it).  SDR applications reuse one
// or two sample buffers for the life of a stream, and those are pageable memory (a numpy array,
// a std::vector), which the GPU can only reach through a CPU bounce copy.  With pin=1 the
// driver page-locks each distinct (buffer, size) it is handed, once, so that later calls move
// it by DMA.  The caller must keep such a buffer alive until closeStream: page-locking follows
// the virtual address, and freed-then-reused memory would still look pinned.  Off by default.
void SoapySXB200::pin_if_asked(Endpoint &ep, const void *buffer, size_t bytes)
{
    if (!ep.pin_caller_buffers || bytes == 0)
        return;
    for (const auto &known : ep.pinned)
        if (known.first == buffer && known.second >= bytes)
            return;
    for (auto &known : ep.pinned) {
        if (known.first == buffer) { // same buffer, now used with a larger size: pin again, larger
            sxgpu_host_unregister(gpu_, const_cast<void *>(buffer));
            known.second = 0;
        }
    }
    if (sxgpu_host_register(gpu_, const_cast<void *>(buffer), bytes) == SXGPU_OK) {
        ep.pinned.erase(std::remove_if(ep.pinned.begin(), ep.pinned.end(),
                                       [&](const std::pair<const void *, size_t> &k) { return k.first == buffer; }),
                        ep.pinned.end());
        ep.pinned.emplace_back(buffer, bytes);
    } else {
        // Already pinned by the caller, overlapping another registration, or not pinnable: the
        // conversion still works (in place if it is pinned, through the bounce copy if not).
        SoapySDR_logf(SOAPY_SDR_DEBUG, "pin=1: buffer %p not registered (%s)", buffer, sxgpu_last_error(gpu_));
    }
}

void SoapySXB200::unpin_all(Endpoint &ep)
{
    for (const auto &known : ep.pinned)
        if (known.second)
            sxgpu_host_unregister(gpu_, const_cast<void *>(known.first));
    ep.pinned.clear();
}

size_t SoapySXB200::getStreamMTU(SoapySDR::Stream *stream) const
{
    Endpoint *ep = endpoint_of(stream);
    std::scoped_lock lock(ep->mutex);
    return ep->ring.period;
}

int SoapySXB200::activateStream(SoapySDR::Stream *stream, const int, const long long, const size_t)
{
    std::scoped_lock lock(rx_.mutex, tx_.mutex);
    Endpoint *ep = endpoint_of(stream);
    if (ep->active) {
        SoapySDR_logf(SOAPY_SDR_ERROR, "Stream was already activated");
        return SOAPY_SDR_STREAM_ERROR;
    }
    ep->active = true;
    // Linked-mode streams start on the first TX write instead (reference :36-39, :821-825).
    if (ep->mode == Endpoint::Mode::Normal && snd_pcm_state(ep->pcm) == SND_PCM_STATE_PREPARED) {
        if (check_alsa(snd_pcm_start(ep->pcm), "snd_pcm_start") < 0)
            return SOAPY_SDR_STREAM_ERROR;
    }
    return 0;
}

int SoapySXB200::deactivateStream(SoapySDR::Stream *stream, const int, const long long)
{
    std::scoped_lock lock(rx_.mutex, tx_.mutex);
    Endpoint *ep = endpoint_of(stream);
    if (!ep->active) {
        SoapySDR_logf(SOAPY_SDR_ERROR, "Stream was already deactivated");
        return SOAPY_SDR_STREAM_ERROR;
    }
    ep->active = false;
    if (!rx_.active && !tx_.active) {
        SoapySDR_logf(SOAPY_SDR_INFO, "Stopping and resetting streams");
        if (check_alsa(rx_.reset(), "rx reset") < 0 || check_alsa(tx_.reset(), "tx reset") < 0)
            return SOAPY_SDR_STREAM_ERROR;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// RX
// ---------------------------------------------------------------------------------------
int SoapySXB200::readStream(SoapySDR::Stream *stream, void *const *buffs, const size_t numElems,
                            int &flags, long long &timeNs, const long timeoutUs)
{
    Endpoint &ep = *endpoint_of(stream);
    std::scoped_lock lock(ep.mutex);

    flags = 0;
    if (ep.is_tx())
        throw std::runtime_error("Wrong direction");

    // A long blocking read: the stand-in (like the kernel behind a real snd_pcm_readi) copies ring
    // after ring into the staging buffer, and nothing makes the conversion of the first pieces wait
    // for the last.  The conversion is started on a helper thread before the read and gated by how
    // far the read has come; this thread reads piece by piece and says so.
    if (overlap_reads_ && !ep.cs16 && numElems >= kGatedReadFrames && timeoutUs > 0 && ep.active) {
        pin_if_asked(ep, buffs[0], numElems * 8);
        stage_rx_->reserve(numElems);
        if (!rx_helper_)
            rx_helper_ = std::make_unique<Sidekick>();
        rx_ready_.store(0, std::memory_order_relaxed);
        int conv_rc = SXGPU_OK;
        size_t converted = 0;
        void *const stage = stage_rx_->data();
        void *const dest = buffs[0];
        rx_helper_->start([&, stage, dest] {
            conv_rc = sxgpu_convert_rx_buffer_host_gated(gpu_, stage, 0, dest, 0, numElems,
                                                         reinterpret_cast<const volatile uint64_t *>(&rx_ready_), &converted);
        });
        struct Release { // whatever happens to the read, the helper is told where the block ends and waited for
            SoapySXB200 *self;
            uint64_t got = 0;
            ~Release()
            {
                self->rx_ready_.store(got | (uint64_t(1) << 63), std::memory_order_release);
                self->rx_helper_->finish();
            }
        };
        RxOutcome rx;
        {
            Release release{this};
            rx = rx_before_convert(
                ep, sample_rate_, numElems, timeoutUs, [stage](size_t) { return stage; }, kGatedReadPiece,
                [this](size_t first, size_t frames) { rx_ready_.store(first + frames, std::memory_order_release); });
            release.got = rx.ret > 0 ? uint64_t(rx.ret) : 0;
        }
        if (rx.time_valid)
            timeNs = rx.time_ns;
        flags |= rx.flags;
        if (rx.ret <= 0)
            return rx.ret;
        if (conv_rc != SXGPU_OK || converted != size_t(rx.ret)) {
            SoapySDR_logf(SOAPY_SDR_ERROR, "rx GPU conversion failed: %s (%s)", sxgpu_strerror(conv_rc), sxgpu_last_error(gpu_));
            return SOAPY_SDR_STREAM_ERROR;
        }
        return rx.ret;
    }

    // Everything up to the conversion (csrc/host/stream_ops.hpp): pending frames, overrun skip,
    // non-blocking trim, snd_pcm_readi into pinned staging, timestamp, frame counter.
    const RxOutcome rx = rx_before_convert(ep, sample_rate_, numElems, timeoutUs, [this](size_t frames) {
        stage_rx_->reserve(frames);
        return stage_rx_->data();
    });
    if (rx.time_valid)
        timeNs = rx.time_ns;
    flags |= rx.flags;
    if (rx.ret <= 0)
        return rx.ret;
    const int got = rx.ret;

    // I2S words -> CF32 on the GPU, straight out of pinned staging into the caller's buffer.
    pin_if_asked(ep, buffs[0], numElems * (ep.cs16 ? 4 : 8));
    int rc = ep.cs16
                 ? sxgpu_convert_rx_buffer_cs16_host(gpu_, stage_rx_->data(), 0, buffs[0], 0, size_t(got))
                 : sxgpu_convert_rx_buffer_host(gpu_, stage_rx_->data(), 0, buffs[0], 0, size_t(got));
    if (rc != SXGPU_OK) {
        SoapySDR_logf(SOAPY_SDR_ERROR, "rx GPU conversion failed: %s (%s)", sxgpu_strerror(rc),
                      sxgpu_last_error(gpu_));
        return SOAPY_SDR_STREAM_ERROR;
    }
    return int(got);
}

// ---------------------------------------------------------------------------------------
// TX
// ---------------------------------------------------------------------------------------
int SoapySXB200::writeStream(SoapySDR::Stream *stream, const void *const *buffs,
                             const size_t numElems, int &flags, const long long timeNs,
                             const long timeoutUs)
{
    Endpoint &ep = *endpoint_of(stream);
    std::scoped_lock lock(ep.mutex);

    if (!ep.is_tx())
        throw std::runtime_error("Wrong direction");

    // Everything up to the conversion (csrc/host/stream_ops.hpp): where the block lands, late
    // bursts, the forward over the gap, the non-blocking trim.
    const TxOutcome tx = tx_before_convert(ep, sample_rate_, numElems, flags, timeNs, timeoutUs);
    if (!tx.convert)
        return tx.ret;
    const unsigned long length = tx.length;

    // CF32 -> I2S words on the GPU, from the caller's buffer into pinned staging.
    pin_if_asked(ep, buffs[0], numElems * (ep.cs16 ? 4 : 8));
    stage_tx_->reserve(length);
    int rc = ep.cs16 ? sxgpu_convert_tx_buffer_cs16_host(gpu_, buffs[0], 0, stage_tx_->data(), 0, length,
                                                         tx_threshold2_)
                     : sxgpu_convert_tx_buffer_host(gpu_, buffs[0], 0, stage_tx_->data(), 0, length,
                                                    tx_threshold2_);
    if (rc != SXGPU_OK) {
        SoapySDR_logf(SOAPY_SDR_ERROR, "tx GPU conversion failed: %s (%s)", sxgpu_strerror(rc),
                      sxgpu_last_error(gpu_));
        return SOAPY_SDR_STREAM_ERROR;
    }

    return tx_after_convert(ep, stage_tx_->data(), length);
}

// ---------------------------------------------------------------------------------------
// Hardware time: the frame being played right now, read off the TX side so that a TX
// thread never contends with the RX thread's lock (reference :1107-1139).
// ---------------------------------------------------------------------------------------
long long SoapySXB200::getHardwareTime(const std::string &what) const
{
    if (!what.empty())
        throw std::runtime_error("Unsupported time");
    std::scoped_lock lock(tx_.mutex);
    snd_pcm_sframes_t room = 0, queued = 0;
    if (snd_pcm_avail_delay(tx_.pcm, &room, &queued) < 0)
        throw std::runtime_error("ALSA error");
    return SoapySDR::ticksToTimeNs(tx_.position - int64_t(queued), sample_rate_);
}

// ---------------------------------------------------------------------------------------
// Sample rate: state only, but the SAME set of legal rates, because the rate feeds every
// timestamp (reference :1145-1219).
// ---------------------------------------------------------------------------------------
std::vector<double> SoapySXB200::listSampleRates(const int, const size_t) const
{
    std::vector<double> rates;
    for (unsigned div : kRateDividers)
        rates.push_back(master_clock_ / double(div));
    return rates;
}

SoapySDR::RangeList SoapySXB200::getSampleRateRange(const int direction, const size_t channel) const
{
    SoapySDR::RangeList ranges;
    for (double r : listSampleRates(direction, channel))
        ranges.push_back(SoapySDR::Range(r, r, 0));
    return ranges;
}

void SoapySXB200::setSampleRate(const int, const size_t, const double rate)
{
    std::scoped_lock lock(settings_mutex_);
    if (!(rate > 0))
        throw std::runtime_error("Sample rate must be positive");
    const double divider = std::round(master_clock_ / rate);
    for (unsigned div : kRateDividers) {
        if (double(div) == divider) {
            sample_rate_ = master_clock_ / divider;
            return;
        }
    }
    throw std::runtime_error("Unsupported sample rate");
}

double SoapySXB200::getSampleRate(const int, const size_t) const
{
    std::scoped_lock lock(settings_mutex_);
    return sample_rate_;
}

// ---------------------------------------------------------------------------------------
// RF settings: quantised like the SX1255 registers would (so get-after-set agrees with the
// reference), stored in plain members.
// ---------------------------------------------------------------------------------------
void SoapySXB200::setFrequency(const int direction, const size_t, const double frequency,
                               const SoapySDR::Kwargs &)
{
    std::scoped_lock lock(settings_mutex_);
    // 24-bit synthesiser word in steps of masterClock / 2^20 (reference :1236-1239).
    const double step = master_clock_ / double(1L << 20);
    const double top = step * double((1L << 24) - 1);
    const double clamped = std::min(std::max(frequency, 0.0), top);
    frequency_word_[direction == SOAPY_SDR_RX ? SOAPY_SDR_RX : SOAPY_SDR_TX] =
        uint32_t(int(std::round(clamped / step)));
}

double SoapySXB200::getFrequency(const int direction, const size_t) const
{
    std::scoped_lock lock(settings_mutex_);
    const double step = master_clock_ / double(1L << 20);
    return step * double(frequency_word_[direction == SOAPY_SDR_RX ? SOAPY_SDR_RX : SOAPY_SDR_TX]);
}

std::vector<std::string> SoapySXB200::listGains(const int direction, const size_t) const
{
    if (direction == SOAPY_SDR_RX)
        return {"LNA", "PGA"};
    return {"DAC", "MIXER"};
}

SoapySDR::Range SoapySXB200::getGainRange(const int direction, const size_t,
                                          const std::string &name) const
{
    // Reference :1291-1306.
    if (direction == SOAPY_SDR_RX) {
        if (name == "LNA") return SoapySDR::Range(0.0, 48.0, 6.0);
        if (name == "PGA") return SoapySDR::Range(0.0, 30.0, 2.0);
    } else {
        if (name == "DAC") return SoapySDR::Range(0.0, 9.0, 3.0);
        if (name == "MIXER") return SoapySDR::Range(0.0, 30.0, 2.0);
    }
    return SoapySDR::Range(0, 0, 0);
}

void SoapySXB200::setGain(const int direction, const size_t channel, const std::string &name,
                          const double value)
{
    std::scoped_lock lock(settings_mutex_);
    const auto names = listGains(direction, channel);
    for (size_t i = 0; i < names.size(); i++) {
        if (names[i] != name)
            continue;
        SoapySDR::Range r = getGainRange(direction, channel, name);
        double steps = std::round((std::min(std::max(value, r.minimum()), r.maximum()) - r.minimum()) / r.step());
        // The SX1255 LNA register only has every other 6 dB step below 36 dB, so odd steps
        // up to 6 read back one lower (register encoding at reference :1319-1327, :1354-1356).
        if (name == "LNA" && steps <= 6)
            steps -= std::fmod(steps, 2.0);
        gain_[direction == SOAPY_SDR_RX ? 1 : 0][i] = r.minimum() + r.step() * steps;
    }
}

double SoapySXB200::getGain(const int direction, const size_t channel, const std::string &name) const
{
    std::scoped_lock lock(settings_mutex_);
    const auto names = listGains(direction, channel);
    for (size_t i = 0; i < names.size(); i++)
        if (names[i] == name)
            return gain_[direction == SOAPY_SDR_RX ? 1 : 0][i];
    return 0.0;
}

void SoapySXB200::setGain(const int direction, const size_t channel, const double value)
{
    std::scoped_lock lock(settings_mutex_);
    // Same distribution rule as the reference (:1370-1394): park the fine-stepped stage near
    // a target and let the coarse stage cover the range.
    if (direction == SOAPY_SDR_RX) {
        setGain(direction, channel, "LNA", value - 12.0);
        setGain(direction, channel, "PGA", value - getGain(direction, channel, "LNA"));
    } else {
        setGain(direction, channel, "DAC", value - 26.0);
        setGain(direction, channel, "MIXER", value - getGain(direction, channel, "DAC"));
    }
}

std::vector<std::string> SoapySXB200::listAntennas(const int direction, const size_t) const
{
    if (direction == SOAPY_SDR_RX)
        return {"RX", "LB"};
    return {"TX", "NONE"};
}

void SoapySXB200::setAntenna(const int direction, const size_t, const std::string &name)
{
    std::scoped_lock lock(settings_mutex_);
    if (direction == SOAPY_SDR_RX) {
        if (name == "RX" || name == "LB" || name == "DLB")
            antenna_[SOAPY_SDR_RX] = name;
    } else if (name == "TX" || name == "NONE") {
        antenna_[SOAPY_SDR_TX] = name;
    }
}

std::string SoapySXB200::getAntenna(const int direction, const size_t) const
{
    std::scoped_lock lock(settings_mutex_);
    return antenna_[direction == SOAPY_SDR_RX ? SOAPY_SDR_RX : SOAPY_SDR_TX];
}

void SoapySXB200::writeSetting(const std::string &key, const std::string &value)
{
    std::scoped_lock lock(settings_mutex_);
    if (key == "PA" && (value == "ON" || value == "OFF" || value == "AUTO"))
        pa_mode_ = value; // reference :1472-1493 drives two GPIO lines here
    else if (key.rfind("sxgpu.", 0) == 0) // a tuning option of the C ABI; unknown keys are ignored like any other
        sxgpu_set_option(gpu_, key.c_str() + 6, std::atoll(value.c_str()));
}

std::string SoapySXB200::readSetting(const std::string &key) const
{
    std::scoped_lock lock(settings_mutex_);
    if (key.rfind("sxgpu.", 0) == 0) { // counters ("sxgpu.launches", ...) and options of the C ABI
        uint64_t counter = 0;
        int64_t option = 0;
        if (sxgpu_get_counter(gpu_, key.c_str() + 6, &counter) == SXGPU_OK)
            return std::to_string(counter);
        if (sxgpu_get_option(gpu_, key.c_str() + 6, &option) == SXGPU_OK)
            return std::to_string(option);
        return "";
    }
    return key == "PA" ? pa_mode_ : "";
}

} // namespace sxhost

// ---------------------------------------------------------------------------------------
// Registration: the same probe result as the reference (SoapySX.cpp:1629-1656).
// ---------------------------------------------------------------------------------------
static SoapySDR::KwargsList findSXB200(const SoapySDR::Kwargs &)
{
    SoapySDR::Kwargs found;
    found["label"] = "sx";
    found["driver"] = "sx";
    return SoapySDR::KwargsList{found};
}

static SoapySDR::Device *makeSXB200(const SoapySDR::Kwargs &args)
{
    return new sxhost::SoapySXB200(args);
}

static SoapySDR::Registry registerSXB200("sx", &findSXB200, &makeSXB200, SOAPY_SDR_ABI_VERSION);

// ---------------------------------------------------------------------------------------
// The bookkeeping rules, exported flat so they can be unit-tested without a GPU.
// ---------------------------------------------------------------------------------------
extern "C" {

void sxplan_geometry(unsigned long requested_period, unsigned long *period, unsigned long *buffer)
{
    sxplan::Geometry g = sxplan::geometry_for_period(requested_period);
    *period = g.period;
    *buffer = g.buffer;
}

unsigned long sxplan_overrun_skip(long pending, unsigned long buffer, unsigned long period)
{
    sxplan::Geometry g = {period, buffer};
    return sxplan::overrun_skip(pending, g);
}

unsigned long sxplan_trim_nonblocking(unsigned long wanted, long available, long timeoutUs)
{
    return sxplan::trim_nonblocking(wanted, available, timeoutUs);
}

int64_t sxplan_clock_after_forward(int64_t clock, int64_t position, int64_t target, int64_t ring, int64_t period)
{
    return sxplan::clock_after_forward(clock, position, target, ring, period);
}

void sxplan_place_tx_block(int64_t position, long queued, int has_time, int64_t time_ticks,
                           unsigned long period, int *discard, int64_t *write_position,
                           int64_t *underrun_jump)
{
    sxplan::TxPlacement p = sxplan::place_tx_block(position, queued, has_time != 0, time_ticks, period);
    *discard = p.discard ? 1 : 0;
    *write_position = p.write_position;
    *underrun_jump = p.underrun_jump;
}

} // extern "C"

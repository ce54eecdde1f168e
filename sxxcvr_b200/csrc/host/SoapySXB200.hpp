// SoapySXB200 -- the "driver=sx" SoapySDR device whose stream path runs on a B200.
//
// It keeps the plugin surface of the reference driver (class SoapySX, SoapySX.cpp:524-1624)
// for the stream path -- setupStream / closeStream / activateStream / deactivateStream /
// getStreamMTU / readStream / writeStream / getHardwareTime / getStreamFormats /
// getNativeStreamFormat, the CF32 format, SOAPY_SDR_HAS_TIME timestamps, the `threshold`,
// `link` and `period` stream arguments and the driver=sx probe -- with the same return
// values, flags, timestamps, exceptions and sample-counter bookkeeping.  What changes is
// where the samples are converted: instead of two scalar loops on the calling thread
// (convert_rx_buffer / convert_tx_buffer, :103-137) the block goes through the sxgpu C ABI
// (include/sxgpu.h) to sm_100a kernels, and the two grow-only staging vectors (:552-557)
// become pinned host memory that the GPU reads and writes directly.
//
// The SX1255 control plane (registers, SPI, GPIO) is out of scope: there is no chip on a GPU
// box.  Sample rate, frequency, gain and antenna are kept as plain state so that the
// reference's example scripts' setup calls succeed and timestamps use the right rate.
#pragma once

#include <SoapySDR/Device.hpp>

#include <alsa/asoundlib.h>

#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include <atomic>

#include "par_copy.hpp"
#include "stream_plan.hpp"

struct sxgpu_ctx;

namespace sxhost {

// One direction of the I2S link: the PCM handle plus the driver-side stream state
// (the reference keeps the same facts in AlsaPcm, SoapySX.cpp:369-393).
struct Endpoint {
    enum class Mode { Normal, Linked }; // reference enum stream_mode, :31-44

    const char *pcm_name;
    snd_pcm_stream_t direction;
    snd_pcm_t *pcm = nullptr;
    mutable std::mutex mutex;
    Mode mode = Mode::Normal;
    bool cs16 = false; // EXTENSION: stream format CS16 instead of CF32 (device argument cs16=1)
    bool pin_caller_buffers = false; // stream argument pin=1, see SoapySXB200::pin_if_asked
    std::vector<std::pair<const void *, size_t>> pinned; // caller buffers this stream page-locked
    bool configured = false;
    bool active = false;
    int64_t position = 0; // frames read / written / skipped since the last reset
    sxplan::Geometry ring = {0, 0};

    Endpoint(const char *name, snd_pcm_stream_t dir) : pcm_name(name), direction(dir) {}
    ~Endpoint();
    bool is_tx() const { return direction == SND_PCM_STREAM_PLAYBACK; }
    void open();
    void configure(unsigned long requested_period);
    int reset();
};

// Grow-only pinned host staging for one direction: element = one 8-byte I2S frame.
class PinnedFrames {
public:
    explicit PinnedFrames(sxgpu_ctx *gpu) : gpu_(gpu) {}
    ~PinnedFrames();
    PinnedFrames(const PinnedFrames &) = delete;
    PinnedFrames &operator=(const PinnedFrames &) = delete;
    void reserve(size_t frames);
    void *data() { return data_; }
    size_t capacity() const { return capacity_; }

private:
    sxgpu_ctx *gpu_;
    void *data_ = nullptr;
    size_t capacity_ = 0;
};

class SoapySXB200 : public SoapySDR::Device {
public:
    explicit SoapySXB200(const SoapySDR::Kwargs &args);
    ~SoapySXB200() override;

    // identification
    std::string getDriverKey() const override { return "sx"; }
    std::string getHardwareKey() const override { return "sx"; }
    SoapySDR::Kwargs getHardwareInfo() const override;
    size_t getNumChannels(const int) const override { return 1; }

    // streams
    std::vector<std::string> getStreamFormats(const int direction, const size_t channel) const override;
    std::string getNativeStreamFormat(const int direction, const size_t channel,
                                      double &fullScale) const override;
    SoapySDR::Stream *setupStream(const int direction, const std::string &format,
                                  const std::vector<size_t> &channels,
                                  const SoapySDR::Kwargs &args) override;
    void closeStream(SoapySDR::Stream *stream) override;
    size_t getStreamMTU(SoapySDR::Stream *stream) const override;
    int activateStream(SoapySDR::Stream *stream, const int flags, const long long timeNs,
                       const size_t numElems) override;
    int deactivateStream(SoapySDR::Stream *stream, const int flags, const long long timeNs) override;
    int readStream(SoapySDR::Stream *stream, void *const *buffs, const size_t numElems, int &flags,
                   long long &timeNs, const long timeoutUs) override;
    int writeStream(SoapySDR::Stream *stream, const void *const *buffs, const size_t numElems,
                    int &flags, const long long timeNs, const long timeoutUs) override;

    // time
    bool hasHardwareTime(const std::string &what) const override { return what.empty(); }
    long long getHardwareTime(const std::string &what) const override;

    // sample rate (state only)
    std::vector<double> listSampleRates(const int direction, const size_t channel) const override;
    SoapySDR::RangeList getSampleRateRange(const int direction, const size_t channel) const override;
    void setSampleRate(const int direction, const size_t channel, const double rate) override;
    double getSampleRate(const int direction, const size_t channel) const override;

    // RF settings (state only; no SX1255 behind them)
    void setFrequency(const int direction, const size_t channel, const double frequency,
                      const SoapySDR::Kwargs &args) override;
    double getFrequency(const int direction, const size_t channel) const override;
    std::vector<std::string> listGains(const int direction, const size_t channel) const override;
    SoapySDR::Range getGainRange(const int direction, const size_t channel,
                                 const std::string &name) const override;
    void setGain(const int direction, const size_t channel, const double value) override;
    void setGain(const int direction, const size_t channel, const std::string &name,
                 const double value) override;
    double getGain(const int direction, const size_t channel, const std::string &name) const override;
    std::vector<std::string> listAntennas(const int direction, const size_t channel) const override;
    void setAntenna(const int direction, const size_t channel, const std::string &name) override;
    std::string getAntenna(const int direction, const size_t channel) const override;
    void writeSetting(const std::string &key, const std::string &value) override;
    std::string readSetting(const std::string &key) const override;

    sxgpu_ctx *gpu() const { return gpu_; }

private:
    Endpoint *endpoint_of(SoapySDR::Stream *stream) const
    {
        return reinterpret_cast<Endpoint *>(stream);
    }
    void pin_if_asked(Endpoint &ep, const void *buffer, size_t bytes);
    void unpin_all(Endpoint &ep);

    sxgpu_ctx *gpu_ = nullptr;
    int gpu_ordinal_ = 0;
    bool cs16_enabled_ = false; // EXTENSION, off by default: the reference offers CF32 only

    double master_clock_;
    double sample_rate_;
    mutable std::recursive_mutex settings_mutex_;
    uint32_t frequency_word_[2] = {0, 0}; // indexed by direction
    // [0] = TX {DAC, MIXER}, [1] = RX {LNA, PGA}.  Start where the reference's register defaults
    // read back (init_registers, SoapySX.cpp:146-176: 0x08 = 0x2E -> DAC 6 dB, MIXER 28 dB;
    // 0x0C = 0x3F -> LNA 48 dB, PGA 30 dB), so getGain before any setGain agrees too.
    double gain_[2][2] = {{6.0, 28.0}, {48.0, 30.0}};
    std::string antenna_[2];
    std::string pa_mode_ = "AUTO";

    mutable Endpoint rx_;
    mutable Endpoint tx_;
    float tx_threshold2_ = 0.0f;
    bool linked_ = false;

    // Long blocking reads (kGatedReadFrames and more): the conversion runs on a helper thread while
    // this one is still reading, gated by how far the read has come (sxgpu_convert_rx_buffer_host_gated).
    std::unique_ptr<Sidekick> rx_helper_;
    alignas(64) std::atomic<uint64_t> rx_ready_{0};
    bool overlap_reads_ = true; // device argument overlap=0 turns it off
    std::unique_ptr<PinnedFrames> stage_rx_;
    std::unique_ptr<PinnedFrames> stage_tx_;
};

} // namespace sxhost

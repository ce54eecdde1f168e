// sxh_* : a flat C view of a SoapySDR "driver=sx" device, for callers without C++ or SWIG
// (the parity tests and bench.py drive it through ctypes).  It talks only to the public
// SoapySDR::Device interface, so the same file is linked into the product module
// (libsxsoapy.so, CUDA-backed SoapySXB200) and into the CPU oracle build of the unmodified
// reference (oracle/_ref/libsx_ref.so); a test runs one script against both and compares.
// C++ exceptions are turned into SXH_THREW + sxh_last_error().
#include <SoapySDR/Device.hpp>
#include <SoapySDR/Logger.hpp>
#include <SoapySDR/Time.hpp>

#include <alsa/asoundlib.h>

#include <chrono>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#define SXH_THREW (-1000)

struct sxh_device {
    SoapySDR::Device *dev = nullptr;
    snd_pcm_t *capture = nullptr;
    snd_pcm_t *playback = nullptr;
};

namespace {
thread_local std::string t_error;
thread_local std::string t_text;

template <typename F> int guarded(F &&body)
{
    try {
        return body();
    } catch (const std::exception &e) {
        t_error = e.what();
    } catch (...) {
        t_error = "unknown exception";
    }
    return SXH_THREW;
}
}

extern "C" {

const char *sxh_last_error(void) { return t_error.c_str(); }

void sxh_set_log_level(int level) { SoapySDR_setLogLevel(SoapySDRLogLevel(level)); }

// "key=value, ..." per result, results separated by ';'.
const char *sxh_enumerate(const char *args)
{
    t_text.clear();
    guarded([&] {
        for (const auto &kw : SoapySDR::Device::enumerate(std::string(args ? args : ""))) {
            if (!t_text.empty())
                t_text += ";";
            t_text += SoapySDR::KwargsToString(kw);
        }
        return 0;
    });
    return t_text.c_str();
}

int sxh_make(const char *args, sxh_device **out)
{
    *out = nullptr;
    return guarded([&] {
        // The device opens its capture and playback PCMs in its constructor
        // (reference SoapySX.cpp:698-699, :720-721); pick them out of the stub's registry.
        size_t before = sx_alsa_pcm_count();
        SoapySDR::Device *dev = SoapySDR::Device::make(std::string(args ? args : ""));
        sxh_device *h = new sxh_device();
        h->dev = dev;
        for (size_t i = before; i < sx_alsa_pcm_count(); i++) {
            snd_pcm_t *pcm = sx_alsa_pcm_at(i);
            (sx_alsa_pcm_is_capture(pcm) ? h->capture : h->playback) = pcm;
        }
        *out = h;
        return 0;
    });
}

int sxh_unmake(sxh_device *h)
{
    return guarded([&] {
        SoapySDR::Device::unmake(h->dev);
        delete h;
        return 0;
    });
}

snd_pcm_t *sxh_pcm(sxh_device *h, int capture) { return capture ? h->capture : h->playback; }

void *sxh_setup_stream(sxh_device *h, int direction, const char *format, const char *args)
{
    void *stream = nullptr;
    guarded([&] {
        stream = h->dev->setupStream(direction, format, std::vector<size_t>{0},
                                     SoapySDR::KwargsFromString(args ? args : ""));
        return 0;
    });
    return stream;
}

int sxh_close_stream(sxh_device *h, void *stream)
{
    return guarded([&] {
        h->dev->closeStream(static_cast<SoapySDR::Stream *>(stream));
        return 0;
    });
}

int sxh_activate(sxh_device *h, void *stream, int flags, long long timeNs, size_t numElems)
{
    return guarded([&] {
        return h->dev->activateStream(static_cast<SoapySDR::Stream *>(stream), flags, timeNs,
                                      numElems);
    });
}

int sxh_deactivate(sxh_device *h, void *stream, int flags, long long timeNs)
{
    return guarded([&] {
        return h->dev->deactivateStream(static_cast<SoapySDR::Stream *>(stream), flags, timeNs);
    });
}

long sxh_mtu(sxh_device *h, void *stream)
{
    long mtu = -1;
    guarded([&] {
        mtu = long(h->dev->getStreamMTU(static_cast<SoapySDR::Stream *>(stream)));
        return 0;
    });
    return mtu;
}

int sxh_read(sxh_device *h, void *stream, void *buf, size_t numElems, int *flags,
             long long *timeNs, long timeoutUs)
{
    return guarded([&] {
        void *buffs[1] = {buf};
        return h->dev->readStream(static_cast<SoapySDR::Stream *>(stream), buffs, numElems,
                                  *flags, *timeNs, timeoutUs);
    });
}

int sxh_write(sxh_device *h, void *stream, const void *buf, size_t numElems, int *flags,
              long long timeNs, long timeoutUs)
{
    return guarded([&] {
        const void *buffs[1] = {buf};
        return h->dev->writeStream(static_cast<SoapySDR::Stream *>(stream), buffs, numElems,
                                   *flags, timeNs, timeoutUs);
    });
}

int sxh_hardware_time(sxh_device *h, const char *what, long long *timeNs)
{
    return guarded([&] {
        *timeNs = h->dev->getHardwareTime(what ? what : "");
        return 0;
    });
}

int sxh_has_hardware_time(sxh_device *h, const char *what)
{
    return guarded([&] { return h->dev->hasHardwareTime(what ? what : "") ? 1 : 0; });
}

int sxh_set_sample_rate(sxh_device *h, int direction, double rate)
{
    return guarded([&] {
        h->dev->setSampleRate(direction, 0, rate);
        return 0;
    });
}

double sxh_get_sample_rate(sxh_device *h, int direction)
{
    double rate = -1.0;
    guarded([&] {
        rate = h->dev->getSampleRate(direction, 0);
        return 0;
    });
    return rate;
}

int sxh_list_sample_rates(sxh_device *h, int direction, double *out, int capacity)
{
    return guarded([&] {
        std::vector<double> rates = h->dev->listSampleRates(direction, 0);
        for (size_t i = 0; i < rates.size() && int(i) < capacity; i++)
            out[i] = rates[i];
        return int(rates.size());
    });
}

int sxh_num_channels(sxh_device *h, int direction)
{
    return guarded([&] { return int(h->dev->getNumChannels(direction)); });
}

// Comma-joined list of stream formats.
const char *sxh_stream_formats(sxh_device *h, int direction)
{
    t_text.clear();
    guarded([&] {
        for (const auto &f : h->dev->getStreamFormats(direction, 0)) {
            if (!t_text.empty())
                t_text += ",";
            t_text += f;
        }
        return 0;
    });
    return t_text.c_str();
}

const char *sxh_native_format(sxh_device *h, int direction, double *fullScale)
{
    t_text.clear();
    guarded([&] {
        t_text = h->dev->getNativeStreamFormat(direction, 0, *fullScale);
        return 0;
    });
    return t_text.c_str();
}

const char *sxh_driver_key(sxh_device *h)
{
    t_text.clear();
    guarded([&] {
        t_text = h->dev->getDriverKey();
        return 0;
    });
    return t_text.c_str();
}

const char *sxh_hardware_key(sxh_device *h)
{
    t_text.clear();
    guarded([&] {
        t_text = h->dev->getHardwareKey();
        return 0;
    });
    return t_text.c_str();
}

const char *sxh_hardware_info(sxh_device *h)
{
    t_text.clear();
    guarded([&] {
        t_text = SoapySDR::KwargsToString(h->dev->getHardwareInfo());
        return 0;
    });
    return t_text.c_str();
}

int sxh_set_frequency(sxh_device *h, int direction, double frequency)
{
    return guarded([&] {
        h->dev->setFrequency(direction, 0, frequency);
        return 0;
    });
}

double sxh_get_frequency(sxh_device *h, int direction)
{
    double f = -1.0;
    guarded([&] {
        f = h->dev->getFrequency(direction, 0);
        return 0;
    });
    return f;
}

int sxh_set_gain(sxh_device *h, int direction, double gain)
{
    return guarded([&] {
        h->dev->setGain(direction, 0, gain);
        return 0;
    });
}

// name == NULL or "" selects the overall gain.
int sxh_set_gain_element(sxh_device *h, int direction, const char *name, double gain)
{
    return guarded([&] {
        if (name && *name)
            h->dev->setGain(direction, 0, name, gain);
        else
            h->dev->setGain(direction, 0, gain);
        return 0;
    });
}

int sxh_get_gain(sxh_device *h, int direction, const char *name, double *gain)
{
    return guarded([&] {
        *gain = (name && *name) ? h->dev->getGain(direction, 0, name) : h->dev->getGain(direction, 0);
        return 0;
    });
}

// out[0..2] = minimum, maximum, step
int sxh_gain_range(sxh_device *h, int direction, const char *name, double *out)
{
    return guarded([&] {
        SoapySDR::Range r = (name && *name) ? h->dev->getGainRange(direction, 0, name)
                                            : h->dev->getGainRange(direction, 0);
        out[0] = r.minimum(), out[1] = r.maximum(), out[2] = r.step();
        return 0;
    });
}

static const char *joined(const std::vector<std::string> &items)
{
    t_text.clear();
    for (const auto &i : items) {
        if (!t_text.empty())
            t_text += ",";
        t_text += i;
    }
    return t_text.c_str();
}

const char *sxh_list_gains(sxh_device *h, int direction)
{
    t_text.clear();
    guarded([&] {
        joined(h->dev->listGains(direction, 0));
        return 0;
    });
    return t_text.c_str();
}

const char *sxh_list_antennas(sxh_device *h, int direction)
{
    t_text.clear();
    guarded([&] {
        joined(h->dev->listAntennas(direction, 0));
        return 0;
    });
    return t_text.c_str();
}

int sxh_set_antenna(sxh_device *h, int direction, const char *name)
{
    return guarded([&] {
        h->dev->setAntenna(direction, 0, name);
        return 0;
    });
}

const char *sxh_get_antenna(sxh_device *h, int direction)
{
    t_text.clear();
    guarded([&] {
        t_text = h->dev->getAntenna(direction, 0);
        return 0;
    });
    return t_text.c_str();
}

int sxh_read_registers(sxh_device *h, const char *name, unsigned addr, size_t length, unsigned *out)
{
    return guarded([&] {
        std::vector<unsigned> v = h->dev->readRegisters(name ? name : "", addr, length);
        for (size_t i = 0; i < v.size() && i < length; i++)
            out[i] = v[i];
        return int(v.size());
    });
}

int sxh_write_registers(sxh_device *h, const char *name, unsigned addr, const unsigned *values, size_t length)
{
    return guarded([&] {
        h->dev->writeRegisters(name ? name : "", addr, std::vector<unsigned>(values, values + length));
        return 0;
    });
}

const char *sxh_read_setting(sxh_device *h, const char *key)
{
    t_text.clear();
    guarded([&] {
        t_text = h->dev->readSetting(key);
        return 0;
    });
    return t_text.c_str();
}

int sxh_write_setting(sxh_device *h, const char *key, const char *value)
{
    return guarded([&] {
        h->dev->writeSetting(key, value);
        return 0;
    });
}

// ---- timing loops ------------------------------------------------------------------------------
// The same loops, compiled into both libraries, so that the product and the unmodified reference
// are timed by identical code with no interpreter between the calls.  Each returns 0 and the
// elapsed seconds of `iters` iterations (steady clock), or the first unexpected return value.
// per_iter_us (may be NULL) receives every iteration's duration in microseconds.

// readStream(n) then writeStream(n, HAS_TIME, that block's time + latency_ns): the repeater
// iteration of example/linear_repeater.py:50-71 with an identity process().
int sxh_bench_pairs(sxh_device *h, void *rx, void *tx, void *buf, size_t n, int iters,
                    long long latency_ns, double *seconds, float *per_iter_us)
{
    return guarded([&] {
        using clock = std::chrono::steady_clock;
        void *rbuffs[1] = {buf};
        const void *wbuffs[1] = {buf};
        auto *rs = static_cast<SoapySDR::Stream *>(rx);
        auto *ts = static_cast<SoapySDR::Stream *>(tx);
        const auto t0 = clock::now();
        auto last = t0;
        for (int i = 0; i < iters; i++) {
            int flags = 0;
            long long t = 0;
            int r = h->dev->readStream(rs, rbuffs, n, flags, t, 1000000);
            if (r != int(n))
                return r < 0 ? r : -2000;
            int wflags = SOAPY_SDR_HAS_TIME;
            int w = h->dev->writeStream(ts, wbuffs, n, wflags, t + latency_ns, 1000000);
            if (w != int(n))
                return w < 0 ? w : -2001;
            if (per_iter_us) {
                const auto now = clock::now();
                per_iter_us[i] = std::chrono::duration<float, std::micro>(now - last).count();
                last = now;
            }
        }
        *seconds = std::chrono::duration<double>(clock::now() - t0).count();
        return 0;
    });
}

// readStream(n) only.
int sxh_bench_reads(sxh_device *h, void *rx, void *buf, size_t n, int iters, double *seconds)
{
    return guarded([&] {
        using clock = std::chrono::steady_clock;
        void *rbuffs[1] = {buf};
        auto *rs = static_cast<SoapySDR::Stream *>(rx);
        const auto t0 = clock::now();
        for (int i = 0; i < iters; i++) {
            int flags = 0;
            long long t = 0;
            int r = h->dev->readStream(rs, rbuffs, n, flags, t, 1000000);
            if (r != int(n))
                return r < 0 ? r : -2000;
        }
        *seconds = std::chrono::duration<double>(clock::now() - t0).count();
        return 0;
    });
}

// Untimed blocking writeStream(n) only (example/tx_test.py:47-53).
int sxh_bench_writes(sxh_device *h, void *tx, const void *buf, size_t n, int iters, double *seconds)
{
    return guarded([&] {
        using clock = std::chrono::steady_clock;
        const void *wbuffs[1] = {buf};
        auto *ts = static_cast<SoapySDR::Stream *>(tx);
        const auto t0 = clock::now();
        for (int i = 0; i < iters; i++) {
            int flags = 0;
            int w = h->dev->writeStream(ts, wbuffs, n, flags, 0, 1000000);
            if (w != int(n))
                return w < 0 ? w : -2001;
        }
        *seconds = std::chrono::duration<double>(clock::now() - t0).count();
        return 0;
    });
}

long long sxh_ticks_to_time_ns(long long ticks, double rate)
{
    return SoapySDR::ticksToTimeNs(ticks, rate);
}

long long sxh_time_ns_to_ticks(long long timeNs, double rate)
{
    return SoapySDR::timeNsToTicks(timeNs, rate);
}

} // extern "C"

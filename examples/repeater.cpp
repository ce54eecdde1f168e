// repeater.cpp -- full-duplex loop through the driver=sx device, C++ only.
//
// The call pattern of the reference's example/linear_repeater.py:50-71 (read a period, process
// it, write it back timestamped a fixed latency later), written against the SoapySDR C++ API
// exactly as an application on the Raspberry Pi would, but served by the B200 stream path.
// "Processing" is a gain of 0.5.  At the end the program checks, through the ALSA stand-in,
// that every transmitted block landed exactly `latency` frames after the block it answers,
// and prints the per-iteration time.
//
//   usage: sx_repeater [blocks] [period] [rate] [device-args, e.g. "clock=32e6, gpu=1"]
#include <SoapySDR/Device.hpp>
#include <SoapySDR/Formats.h>
#include <SoapySDR/Logger.hpp>

#include <alsa/asoundlib.h>

#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv)
{
    const int blocks = argc > 1 ? std::atoi(argv[1]) : 200;
    const size_t period = argc > 2 ? size_t(std::atol(argv[2])) : 256;
    const double rate = argc > 3 ? std::atof(argv[3]) : 75000.0;
    const long long latency_frames = 3 * (long long)period;
    SoapySDR::setLogLevel(SOAPY_SDR_WARNING);

    const size_t pcm_before = sx_alsa_pcm_count();
    SoapySDR::Kwargs device_args = SoapySDR::KwargsFromString(argc > 4 ? argv[4] : "");
    device_args["driver"] = "sx";
    SoapySDR::Device *sdr = SoapySDR::Device::make(device_args);
    snd_pcm_t *playback = nullptr;
    for (size_t i = pcm_before; i < sx_alsa_pcm_count(); i++)
        if (!sx_alsa_pcm_is_capture(sx_alsa_pcm_at(i)))
            playback = sx_alsa_pcm_at(i);

    sdr->setSampleRate(SOAPY_SDR_RX, 0, rate);
    sdr->setSampleRate(SOAPY_SDR_TX, 0, rate);
    sdr->setFrequency(SOAPY_SDR_RX, 0, 432.55e6);
    sdr->setFrequency(SOAPY_SDR_TX, 0, 434.55e6);
    SoapySDR::Kwargs period_arg{{"period", std::to_string(period)}};
    SoapySDR::Kwargs tx_args = period_arg;
    tx_args["threshold"] = "0"; // keep the transmitter keyed
    SoapySDR::Stream *rx = sdr->setupStream(SOAPY_SDR_RX, SOAPY_SDR_CF32, {0}, period_arg);
    SoapySDR::Stream *tx = sdr->setupStream(SOAPY_SDR_TX, SOAPY_SDR_CF32, {0}, tx_args);
    sdr->activateStream(rx);
    sdr->activateStream(tx);

    const long long latency_ns = std::llround(double(latency_frames) * 1e9 / rate);
    std::vector<std::complex<float>> buf(period);
    void *buffs[1] = {buf.data()};
    int failures = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int b = 0; b < blocks; b++) {
        int flags = 0;
        long long time_ns = 0;
        int got = sdr->readStream(rx, buffs, period, flags, time_ns);
        if (got != int(period) || !(flags & SOAPY_SDR_HAS_TIME)) {
            std::fprintf(stderr, "read %d: ret %d flags %d\n", b, got, flags);
            failures++;
            continue;
        }
        for (auto &z : buf)
            z *= 0.5f;
        int tx_flags = SOAPY_SDR_HAS_TIME;
        int sent = sdr->writeStream(tx, buffs, period, tx_flags, time_ns + latency_ns);
        if (sent != int(period)) {
            std::fprintf(stderr, "write %d: ret %d\n", b, sent);
            failures++;
        }
    }
    double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / blocks;

    // Constant latency: block b was captured at frames [b*period, (b+1)*period) and must have been
    // played at [b*period + latency, ...): silence before, data from there on, no holes.
    for (long long p = 0; p < latency_frames; p++)
        failures += sx_alsa_sink_written(playback, p) ? 1 : 0;
    for (long long p = latency_frames; p < latency_frames + (long long)blocks * (long long)period; p++)
        failures += sx_alsa_sink_written(playback, p) ? 0 : 1;
    failures += sx_alsa_appl_ptr(playback) == latency_frames + (long long)blocks * (long long)period ? 0 : 1;

    sdr->deactivateStream(rx);
    sdr->deactivateStream(tx);
    sdr->closeStream(rx);
    sdr->closeStream(tx);
    SoapySDR::Device::unmake(sdr);
    std::printf("%s: %d blocks of %zu frames at %.0f Hz, TX %lld frames after RX, %.1f us per read+write\n",
                failures ? "FAILED" : "OK", blocks, period, rate, latency_frames, us);
    return failures ? 1 : 0;
}

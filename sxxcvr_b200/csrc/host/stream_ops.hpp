// stream_ops.hpp -- the host-side halves of readStream / writeStream around the conversion.
//
// A stream call is: bookkeeping against the PCM (what the reference does in SoapySX.cpp:897-953
// and :989-1088), the conversion (:957, :1090 -- here a call into the sxgpu C ABI), and for TX
// the hand-over to the PCM (:1093-1097).  The bookkeeping halves live here so that the single
// device (SoapySXB200) and the group device (SoapySXB200Group: N front-ends, ONE conversion per
// period for all of them) apply literally the same rules.  Decisions come from stream_plan.hpp;
// this file only talks to ALSA.  Log lines and return codes are the reference's.
#pragma once

#include <SoapySDR/Constants.h>
#include <SoapySDR/Errors.h>
#include <SoapySDR/Logger.hpp>
#include <SoapySDR/Time.hpp>

#include <alsa/asoundlib.h>

#include <algorithm>
#include <cerrno>
#include <climits>
#include <cstdint>

#include "SoapySXB200.hpp"
#include "stream_plan.hpp"

namespace sxhost {

// ALSA error -> SoapySDR stream error (reference :339-360): -EPIPE is an xrun, named by
// direction; anything else is a generic stream error.
inline int stream_error_from_alsa(const Endpoint &ep, long alsa_error)
{
    if (alsa_error == -EPIPE)
        return ep.is_tx() ? SOAPY_SDR_UNDERFLOW : SOAPY_SDR_OVERFLOW;
    return SOAPY_SDR_STREAM_ERROR;
}

// readStream up to the conversion.  `staging(frames)` returns where snd_pcm_readi may put
// `frames` I2S frames.  Result: ret < 0 an error code, ret == 0 nothing to convert (inactive
// stream, or a non-blocking call with nothing pending), ret > 0 that many frames are in the
// staging buffer, stamped with time_ns / flags.  The caller holds ep.mutex.
struct RxOutcome {
    int ret = 0;
    int flags = 0;
    long long time_ns = 0;
    bool time_valid = false; // time_ns was set (the reference leaves timeNs untouched otherwise)
};

// `piece` > 0: a transfer longer than that is read piece by piece -- what snd_pcm_readi does inside
// anyway, a ring (at most 65536 frames) at a time -- and on_piece(first, frames) is told about
// every piece as it lands, so that its conversion need not wait for the last one.  A piece that
// fails after others have landed ends the call with the frames read so far, as a failing
// snd_pcm_readi does; the error resurfaces on the next call.
template <class Staging, class OnPiece>
inline RxOutcome rx_before_convert(Endpoint &ep, double sample_rate, size_t numElems, long timeoutUs, Staging &&staging,
                                   size_t piece, OnPiece &&on_piece)
{
    RxOutcome out;
    if (!ep.active)
        return out; // an inactive capture PCM would block forever (reference :887-894)

    snd_pcm_sframes_t pending = 0, delay = 0;
    int ret = snd_pcm_avail_delay(ep.pcm, &pending, &delay);
    if (ret < 0) {
        SoapySDR_logf(SOAPY_SDR_ERROR, "rx snd_pcm_avail_delay: %d", ret);
        out.ret = stream_error_from_alsa(ep, ret);
        return out;
    }

    if (unsigned long skip = sxplan::overrun_skip(pending, ep.ring)) {
        snd_pcm_sframes_t skipped = snd_pcm_forward(ep.pcm, skip);
        if (skipped < 0) {
            SoapySDR_logf(SOAPY_SDR_ERROR, "rx snd_pcm_forward: %ld", long(skipped));
            out.ret = stream_error_from_alsa(ep, skipped);
            return out;
        }
        ep.position += skipped;
        pending -= skipped;
        SoapySDR_logf(SOAPY_SDR_WARNING, "RX buffer overrun. Skipped %ld samples", long(skipped));
    }

    unsigned long length = (unsigned long)std::min(numElems, (size_t)ULONG_MAX);
    length = sxplan::trim_nonblocking(length, pending, timeoutUs);
    if (length == 0)
        return out;

    snd_pcm_sframes_t got = 0;
    if (piece == 0 || length <= piece) {
        got = snd_pcm_readi(ep.pcm, staging(size_t(length)), length);
        if (got < 0) {
            out.ret = stream_error_from_alsa(ep, got);
            return out;
        }
        if (got > 0)
            on_piece(size_t(0), size_t(got));
    } else {
        char *base = static_cast<char *>(staging(size_t(length)));
        while ((unsigned long)got < length) {
            const unsigned long want = std::min<unsigned long>(piece, length - (unsigned long)got);
            const snd_pcm_sframes_t r = snd_pcm_readi(ep.pcm, base + size_t(got) * 8, want);
            if (r < 0) {
                if (got == 0) {
                    out.ret = stream_error_from_alsa(ep, r);
                    return out;
                }
                break;
            }
            if (r > 0)
                on_piece(size_t(got), size_t(r));
            got += r;
            if ((unsigned long)r < want)
                break;
        }
    }

    // The block's timestamp is the counter value of its first frame.
    out.time_ns = SoapySDR::ticksToTimeNs(ep.position, sample_rate);
    out.time_valid = true;
    out.flags |= SOAPY_SDR_HAS_TIME;
    ep.position += got;
    out.ret = int(got);
    return out;
}

template <class Staging>
inline RxOutcome rx_before_convert(Endpoint &ep, double sample_rate, size_t numElems, long timeoutUs, Staging &&staging)
{
    return rx_before_convert(ep, sample_rate, numElems, timeoutUs, staging, size_t(0), [](size_t, size_t) {});
}

// writeStream up to the conversion: where the block lands, the forward over the gap, the
// non-blocking trim.  Result: convert == false means the call is over and returns `ret`
// (inactive: 0; late timed burst: numElems, nothing written; an error code; nothing fits: 0);
// convert == true means `length` frames are to be converted and handed to tx_after_convert.
struct TxOutcome {
    bool convert = false;
    int ret = 0;
    unsigned long length = 0;
};

inline TxOutcome tx_before_convert(Endpoint &ep, double sample_rate, size_t numElems, int flags, long long timeNs,
                                   long timeoutUs)
{
    TxOutcome out;
    if (!ep.active)
        return out;

    snd_pcm_sframes_t room = 0, queued = 0;
    int ret = snd_pcm_avail_delay(ep.pcm, &room, &queued);
    if (ret < 0) {
        SoapySDR_logf(SOAPY_SDR_ERROR, "tx snd_pcm_avail_delay: %d", ret);
        out.ret = stream_error_from_alsa(ep, ret);
        return out;
    }

    unsigned long length = (unsigned long)std::min(numElems, (size_t)ULONG_MAX);

    const bool timed = (flags & SOAPY_SDR_HAS_TIME) != 0;
    const int64_t time_ticks = timed ? SoapySDR::timeNsToTicks(timeNs, sample_rate) : 0;
    sxplan::TxPlacement where = sxplan::place_tx_block(ep.position, queued, timed, time_ticks, ep.ring.period);
    if (where.discard) {
        // Late bursts are dropped whole and reported as sent, as most SDR drivers do
        // (reference :1013-1023).
        SoapySDR_logf(SOAPY_SDR_WARNING, "Discarding %lu TX samples timed in the past", length);
        out.ret = int(length);
        return out;
    }
    if (where.underrun_jump > 0)
        SoapySDR_logf(SOAPY_SDR_WARNING, "TX buffer underrun. Forwarding TX stream by %lld samples",
                      (long long)where.underrun_jump);

    // Move the ring's write pointer up to the block's position; what is skipped plays as
    // silence.  When the ring cannot take the whole gap yet, take what fits and wait.
    int64_t gap = where.write_position - ep.position;
    while (gap > 0) {
        snd_pcm_sframes_t step = (snd_pcm_sframes_t)std::min(gap, (int64_t)LONG_MAX);
        snd_pcm_sframes_t fits = snd_pcm_forwardable(ep.pcm);
        if (fits < 0) {
            SoapySDR_logf(SOAPY_SDR_ERROR, "tx snd_pcm_forwardable: %ld", long(fits));
            out.ret = stream_error_from_alsa(ep, fits);
            return out;
        }
        snd_pcm_sframes_t moved;
        if (step < fits) {
            moved = snd_pcm_forward(ep.pcm, step);
        } else {
            moved = snd_pcm_forward(ep.pcm, fits);
            snd_pcm_wait(ep.pcm, -10001);
        }
        if (moved < 0) {
            SoapySDR_logf(SOAPY_SDR_ERROR, "tx snd_pcm_forward: %ld", long(moved));
            out.ret = stream_error_from_alsa(ep, moved);
            return out;
        }
        ep.position += moved;
        gap -= moved;
        room -= moved;
    }

    length = sxplan::trim_nonblocking(length, room, timeoutUs);
    if (length == 0)
        return out;
    out.convert = true;
    out.length = length;
    return out;
}

// writeStream after the conversion: hand the I2S frames to the PCM (reference :1093-1097).
inline int tx_after_convert(Endpoint &ep, const void *staging, unsigned long length)
{
    snd_pcm_sframes_t sent = snd_pcm_writei(ep.pcm, staging, length);
    if (sent < 0)
        return stream_error_from_alsa(ep, sent);
    ep.position += sent;
    return int(sent);
}

} // namespace sxhost

// stream_plan.hpp -- the sample-counter bookkeeping of the SoapySX stream path as pure
// functions: no ALSA, no CUDA, no locks.  The device (SoapySXB200.cpp) asks these what to do
// and then does it; tests ask them directly (exported as sxplan_* in SoapySXB200.cpp).
//
// Every rule here restates observable behaviour of the reference's readStream/writeStream
// (SoapySX.cpp:868-1105); the line each one comes from is cited.  "Position" is the stream's
// running frame counter since the last reset (AlsaPcm::position, :378).
#pragma once

#include <cstdint>

// The same rules run per stream on the GPU (csrc/sx_bank.cuh), so they are plain functions
// with no library calls.
#if defined(__CUDACC__)
#define SXPLAN_HD __host__ __device__ inline
#else
#define SXPLAN_HD inline
#endif

namespace sxplan {

// I2S ring geometry.  Period defaults to 256 frames, both are capped at 65 536 (the most
// the Pi's I2S DMA accepts), and the buffer is the largest whole number of periods (:451,
// :464-466).
struct Geometry {
    unsigned long period;
    unsigned long buffer;
};

SXPLAN_HD Geometry geometry_for_period(unsigned long requested_period)
{
    const unsigned long limit = 65536;
    Geometry g;
    g.period = requested_period != 0 ? requested_period : 256ul;
    if (g.period > limit)
        g.period = limit;
    g.buffer = limit / g.period * g.period;
    return g;
}

// Capture overrun (:910-915): more frames pending than the ring holds means the oldest were
// overwritten.  Skip them in whole periods, plus two periods of margin, so period-aligned
// readers stay aligned.  Returns 0 when nothing was lost.
SXPLAN_HD unsigned long overrun_skip(long pending, const Geometry &g)
{
    if (pending <= long(g.buffer))
        return 0;
    unsigned long lost = (unsigned long)pending - g.buffer;
    return (lost / g.period + 2) * g.period;
}

// A call with timeoutUs <= 0 must not block (:934-942, :1076-1085): trim the transfer to
// what the ring can take or give right now.
SXPLAN_HD unsigned long trim_nonblocking(unsigned long wanted, long available, long timeoutUs)
{
    if (timeoutUs > 0)
        return wanted;
    if (available <= 0)
        return 0;
    return (unsigned long)available < wanted ? (unsigned long)available : wanted;
}

// Where a TX block lands (:1000-1038).
struct TxPlacement {
    bool discard;           // timed block already in the past: report it written, write nothing
    int64_t write_position; // counter value of the block's first frame
    int64_t underrun_jump;  // frames an untimed stream was moved ahead after an underrun
};

// `queued` is ALSA's playback delay: frames written but not yet played (negative after an
// underrun), so position - queued is the frame being played now (:1000).
SXPLAN_HD TxPlacement place_tx_block(int64_t position, long queued, bool has_time,
                                  int64_t time_ticks, unsigned long period)
{
    const int64_t now_playing = position - int64_t(queued);
    TxPlacement p = {false, position, 0};
    if (has_time) {
        p.write_position = time_ticks; // :1012
        p.discard = now_playing > time_ticks; // :1017-1023
        return p;
    }
    const int64_t late = now_playing - position; // :1032
    if (late > 0) {
        p.underrun_jump = (late / int64_t(period) + 2) * int64_t(period); // :1034
        p.write_position = position + p.underrun_jump;
    }
    return p;
}

// Virtual-clock model of the forward loop (:1043-1073) for a stream whose hardware is simulated
// (the stream bank): the write pointer is forwarded from `position` to `target` (> position)
// through a ring of `ring` frames that drains as `clock` advances; room = clock + ring - pointer.
// The reference forwards what fits and, while the gap is not closed, waits until a period of
// room is free (snd_pcm_wait with avail_min = period).  Step by step: if the gap is below the
// room it is closed at once.  Otherwise the pointer takes all the room (leaving min(room, 0)),
// the wait brings the room to exactly one period, and from then on every turn moves one period
// and waits one period until less than a period of gap is left, which then fits.  The clock
// the loop ends with, without running it:
SXPLAN_HD int64_t clock_after_forward(int64_t clock, int64_t position, int64_t target, int64_t ring, int64_t period)
{
    const int64_t gap = target - position;
    if (gap <= 0)
        return clock;
    const int64_t room = clock + ring - position;
    const int64_t fits = room > 0 ? room : 0;
    if (gap < fits)
        return clock;
    clock += period - (room < 0 ? room : 0);     // first wait: room becomes exactly one period
    clock += (gap - fits) / period * period;     // one more period of waiting per whole period of gap left
    return clock;
}

} // namespace sxplan

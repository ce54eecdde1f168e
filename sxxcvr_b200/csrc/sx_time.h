// Sample counter <-> nanosecond timestamp, usable from host and device code.
//
// The reference converts through SoapySDR::ticksToTimeNs / timeNsToTicks (SoapySX.cpp:564,
// :570), an external dependency it does not pin (SoapySX/CMakeLists.txt:45).  This restates
// upstream's lib/TimeC.cpp: whole seconds in integers, the sub-second remainder -- and the
// fractional part of a non-integer rate such as 32 MHz / 1536 -- in double, one llround.
// On the device every double operation is an explicit round-to-nearest intrinsic so that no
// multiply-add is contracted and the result is bit-identical to the host's.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SX_TIME_HD __host__ __device__ inline
#else
#define SX_TIME_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define SX_DMUL(a, b) __dmul_rn((a), (b))
#define SX_DADD(a, b) __dadd_rn((a), (b))
#define SX_DSUB(a, b) __dsub_rn((a), (b))
#define SX_DDIV(a, b) __ddiv_rn((a), (b))
#else
#define SX_DMUL(a, b) ((a) * (b))
#define SX_DADD(a, b) ((a) + (b))
#define SX_DSUB(a, b) ((a) - (b))
#define SX_DDIV(a, b) ((a) / (b))
#endif

SX_TIME_HD long long sx_ticks_to_time_ns(long long ticks, double rate)
{
    const long long whole_rate = (long long)rate;
    const long long seconds = ticks / whole_rate;
    const long long leftover = ticks - seconds * whole_rate;
    const double drift = SX_DMUL((double)seconds, SX_DSUB(rate, (double)whole_rate));
    const double sub_second_ns = SX_DDIV(SX_DMUL(SX_DSUB((double)leftover, drift), 1000000000.0), rate);
    return seconds * 1000000000LL + llround(sub_second_ns);
}

SX_TIME_HD long long sx_time_ns_to_ticks(long long time_ns, double rate)
{
    const long long whole_rate = (long long)rate;
    const long long seconds = time_ns / 1000000000LL;
    const long long leftover_ns = time_ns - seconds * 1000000000LL;
    const double drift = SX_DMUL((double)seconds, SX_DSUB(rate, (double)whole_rate));
    const double sub_second_ticks = SX_DADD(drift, SX_DDIV(SX_DMUL((double)leftover_ns, rate), 1000000000.0));
    return seconds * whole_rate + llround(sub_second_ticks);
}

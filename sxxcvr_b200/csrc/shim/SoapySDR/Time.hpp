// Minimal SoapySDR-compatible tick/time conversion (shim).  Reference call sites:
// SoapySX.cpp:564 (timeNsToTicks), :570 (ticksToTimeNs).
#pragma once
#include <SoapySDR/Time.h>
namespace SoapySDR {
static inline long long ticksToTimeNs(const long long ticks, const double rate)
{
    return SoapySDR_ticksToTimeNs(ticks, rate);
}
static inline long long timeNsToTicks(const long long timeNs, const double rate)
{
    return SoapySDR_timeNsToTicks(timeNs, rate);
}
}

// Minimal SoapySDR-compatible tick/time conversion, C side (shim).
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
long long SoapySDR_ticksToTimeNs(const long long ticks, const double rate);
long long SoapySDR_timeNsToTicks(const long long timeNs, const double rate);
#ifdef __cplusplus
}
#endif

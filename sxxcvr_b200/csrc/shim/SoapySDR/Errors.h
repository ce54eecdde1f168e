// Minimal SoapySDR-compatible stream error codes (shim).  Used at reference
// SoapySX.cpp:339-360, :817, :845.
#pragma once
#define SOAPY_SDR_TIMEOUT (-1)
#define SOAPY_SDR_STREAM_ERROR (-2)
#define SOAPY_SDR_CORRUPTION (-3)
#define SOAPY_SDR_OVERFLOW (-4)
#define SOAPY_SDR_NOT_SUPPORTED (-5)
#define SOAPY_SDR_TIME_ERROR (-6)
#define SOAPY_SDR_UNDERFLOW (-7)
#ifdef __cplusplus
extern "C" {
#endif
const char *SoapySDR_errToStr(int errorCode);
#ifdef __cplusplus
}
#endif

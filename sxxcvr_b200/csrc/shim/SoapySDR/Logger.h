// Minimal SoapySDR-compatible logger, C side (shim).  The reference logs through
// SoapySDR_logf at 32 sites (e.g. SoapySX.cpp:904, :996 on every read/write).
#pragma once
#include <stdarg.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef enum {
    SOAPY_SDR_FATAL = 1,
    SOAPY_SDR_CRITICAL = 2,
    SOAPY_SDR_ERROR = 3,
    SOAPY_SDR_WARNING = 4,
    SOAPY_SDR_NOTICE = 5,
    SOAPY_SDR_INFO = 6,
    SOAPY_SDR_DEBUG = 7,
    SOAPY_SDR_TRACE = 8,
    SOAPY_SDR_SSI = 9
} SoapySDRLogLevel;
typedef void (*SoapySDRLogHandler)(const SoapySDRLogLevel logLevel, const char *message);
void SoapySDR_log(const SoapySDRLogLevel logLevel, const char *message);
void SoapySDR_vlogf(const SoapySDRLogLevel logLevel, const char *format, va_list argList);
void SoapySDR_logf(const SoapySDRLogLevel logLevel, const char *format, ...);
void SoapySDR_registerLogHandler(const SoapySDRLogHandler handler);
void SoapySDR_setLogLevel(const SoapySDRLogLevel logLevel);
SoapySDRLogLevel SoapySDR_getLogLevel(void);
#ifdef __cplusplus
}
#endif

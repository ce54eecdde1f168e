// Minimal SoapySDR-compatible stream format strings (shim).
#pragma once
#include <stddef.h>
#define SOAPY_SDR_CF64 "CF64"
#define SOAPY_SDR_CF32 "CF32"
#define SOAPY_SDR_CS32 "CS32"
#define SOAPY_SDR_CS16 "CS16"
#define SOAPY_SDR_CS8 "CS8"
#ifdef __cplusplus
extern "C" {
#endif
size_t SoapySDR_formatToSize(const char *format);
#ifdef __cplusplus
}
#endif

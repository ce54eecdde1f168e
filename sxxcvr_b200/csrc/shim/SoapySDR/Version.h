// Minimal SoapySDR-compatible version macros (shim).
#pragma once
#define SOAPY_SDR_API_VERSION 0x00080000
#define SOAPY_SDR_ABI_VERSION "0.8"

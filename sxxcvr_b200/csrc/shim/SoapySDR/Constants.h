// Minimal SoapySDR-compatible constants (shim; upstream SoapySDR 0.8 is not installed in
// this image).  Values follow upstream include/SoapySDR/Constants.h so a build against a
// real SoapySDR is source compatible.  Used at reference SoapySX.cpp:760, :951, :1009.
#pragma once
#define SOAPY_SDR_TX 0
#define SOAPY_SDR_RX 1
#define SOAPY_SDR_END_BURST (1 << 1)
#define SOAPY_SDR_HAS_TIME (1 << 2)
#define SOAPY_SDR_END_ABRUPT (1 << 3)
#define SOAPY_SDR_ONE_PACKET (1 << 4)
#define SOAPY_SDR_MORE_FRAGMENTS (1 << 5)
#define SOAPY_SDR_WAIT_TRIGGER (1 << 6)

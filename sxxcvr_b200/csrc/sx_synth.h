// Synthetic I2S capture frames.  Frame k of a capture stream is a pure function of
// (seed, k), so the host-side ALSA stub and the device-side generator kernel produce the
// same "ADC output" without any data crossing PCIe.  I = left slot = low word,
// Q = right slot = high word (slot order: reference dts/sx1255_raspberrypi.dts:58-59,
// SoapySX.cpp:126-135).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SX_HD __host__ __device__ __forceinline__
#else
#define SX_HD static inline
#endif

#define SX_SYNTH_DEFAULT_SEED 0x53581255ull

SX_HD uint64_t sx_synth_frame(uint64_t seed, uint64_t k)
{
    uint64_t z = seed + (k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

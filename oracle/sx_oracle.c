/*
 * sx_oracle.c -- CPU oracle for the SoapySX IQ sample path.  TEST INFRASTRUCTURE ONLY.
 * See sx_oracle.h for the rules on who may use this and how its parity is pinned.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile).  Contraction is
 * switched off explicitly because the reference's threshold test (SoapySX.cpp:132) is an
 * un-fused multiply-add on its x86-64 Release build (no -march, SoapySX/CMakeLists.txt:10-13).
 */
#include "sx_oracle.h"

#include <math.h>
#include <string.h>

/* float -> int32, truncating toward zero, with the saturating / NaN->0 behaviour of
 * ARM fcvtzs and CUDA cvt.rzi.s32.f32.  Stays inside defined C for every input. */
static int32_t trunc_sat_s32(float p)
{
    if (p != p)
        return 0;
    if (p >= 2147483648.0f)
        return INT32_MAX;
    if (p <= -2147483648.0f)
        return INT32_MIN;
    return (int32_t)p;
}

/* SoapySX.cpp:103-112: dest[i] = 2^-31 * (float)src[i] over 2*length words. */
void sxo_convert_rx_buffer(const void *src, size_t src_offset,
                           void *dest, size_t dest_offset, size_t length)
{
    const int32_t *in = (const int32_t *)src + 2 * src_offset;
    float *out = (float *)dest + 2 * dest_offset;
    const float k = 4.656612873077392578125e-10f; /* 2^-31, bits 0x30000000 (:107) */
    size_t nwords = 2 * length;
    for (size_t w = 0; w < nwords; w++)
        out[w] = k * (float)in[w];
}

/* std::min(f, 1.0f) then std::max(., -1.0f) exactly as libstdc++ expands them
 * (SoapySX.cpp:124): both return their FIRST argument when the comparison is false,
 * so a NaN input stays NaN. */
static float clamp_unit(float f)
{
    float c = (1.0f < f) ? 1.0f : f;
    return (c < -1.0f) ? -1.0f : c;
}

/* SoapySX.cpp:116-137. */
void sxo_convert_tx_buffer(const void *src, size_t src_offset,
                           void *dest, size_t dest_offset, size_t length,
                           float tx_threshold2)
{
    const float *in = (const float *)src + 2 * src_offset;
    int32_t *out = (int32_t *)dest + 2 * dest_offset;
    const float k = 2147483648.0f; /* (float)0x7FFFFFFF rounds to 2^31 (:120) */
    for (size_t n = 0; n < length; n++) {
        float fi = in[2 * n], fq = in[2 * n + 1];
        uint32_t vi = (uint32_t)trunc_sat_s32(k * clamp_unit(fi));
        uint32_t vq = (uint32_t)trunc_sat_s32(k * clamp_unit(fq));
        vi &= 0xFFFFFFFCu; /* :130 */
        vq &= 0xFFFFFFFCu; /* :131 */
        /* :132 -- on the UN-clamped inputs; two roundings for the products, one for the sum */
        float ii = fi * fi;
        float qq = fq * fq;
        float mag2 = ii + qq;
        if (mag2 >= tx_threshold2)
            vi |= 3u; /* :133 */
        out[2 * n] = (int32_t)vi;
        out[2 * n + 1] = (int32_t)vq;
    }
}

/* Extension (no reference): S32 I2S word -> CS16 by dropping the low 16 bits.
 * Arithmetic shift = floor; full scale 32768. */
void sxo_convert_rx_buffer_cs16(const void *src, size_t src_offset,
                                void *dest, size_t dest_offset, size_t length)
{
    const int32_t *in = (const int32_t *)src + 2 * src_offset;
    int16_t *out = (int16_t *)dest + 2 * dest_offset;
    for (size_t w = 0; w < 2 * length; w++) {
        uint32_t u = (uint32_t)in[w];
        out[w] = (int16_t)(uint16_t)(u >> 16);
    }
}

/* Extension (no reference): CS16 -> S32 I2S word.  word = s << 16 (the two reserved low
 * bits are already 0); TX-enable uses the same un-fused |z|^2 >= thr2 test on
 * f = s * 2^-15. */
void sxo_convert_tx_buffer_cs16(const void *src, size_t src_offset,
                                void *dest, size_t dest_offset, size_t length,
                                float tx_threshold2)
{
    const int16_t *in = (const int16_t *)src + 2 * src_offset;
    int32_t *out = (int32_t *)dest + 2 * dest_offset;
    for (size_t n = 0; n < length; n++) {
        int16_t si = in[2 * n], sq = in[2 * n + 1];
        uint32_t vi = (uint32_t)(uint16_t)si << 16;
        uint32_t vq = (uint32_t)(uint16_t)sq << 16;
        float fi = (float)si * 3.0517578125e-05f, fq = (float)sq * 3.0517578125e-05f;
        float ii = fi * fi;
        float qq = fq * fq;
        float mag2 = ii + qq;
        if (mag2 >= tx_threshold2)
            vi |= 3u;
        out[2 * n] = (int32_t)vi;
        out[2 * n + 1] = (int32_t)vq;
    }
}

/* Extension (no reference): S16_LE I2S frames <-> CF32. */
void sxo_convert_rx_buffer_s16(const void *src, size_t src_offset,
                               void *dest, size_t dest_offset, size_t length)
{
    const int16_t *in = (const int16_t *)src + 2 * src_offset;
    float *out = (float *)dest + 2 * dest_offset;
    for (size_t w = 0; w < 2 * length; w++)
        out[w] = 3.0517578125e-05f * (float)in[w]; /* 2^-15, exact */
}

static uint16_t trunc_sat_s16(float p)
{
    if (p != p)
        return 0;
    if (p >= 32768.0f)
        return 0x7FFF;
    if (p <= -32768.0f)
        return 0x8000;
    return (uint16_t)(int16_t)(int32_t)p;
}

void sxo_convert_tx_buffer_s16(const void *src, size_t src_offset,
                               void *dest, size_t dest_offset, size_t length,
                               float tx_threshold2)
{
    const float *in = (const float *)src + 2 * src_offset;
    uint16_t *out = (uint16_t *)dest + 2 * dest_offset;
    for (size_t n = 0; n < length; n++) {
        float fi = in[2 * n], fq = in[2 * n + 1];
        uint16_t vi = trunc_sat_s16(32768.0f * fi) & 0xFFFCu;
        uint16_t vq = trunc_sat_s16(32768.0f * fq) & 0xFFFCu;
        float ii = fi * fi;
        float qq = fq * fq;
        float mag2 = ii + qq;
        if (mag2 >= tx_threshold2)
            vi |= 3u;
        out[2 * n] = vi;
        out[2 * n + 1] = vq;
    }
}

/* SoapySDR lib/TimeC.cpp (external, unpinned): whole seconds in integers, the
 * remainder in double, llround. */
long long sxo_ticks_to_time_ns(long long ticks, double rate)
{
    const long long ratell = (long long)rate;
    const long long full = ticks / ratell;
    const long long err = ticks - full * ratell;
    const double part = (double)full * (rate - (double)ratell);
    const double frac = (((double)err - part) * 1000000000.0) / rate;
    return full * 1000000000LL + llround(frac);
}

long long sxo_time_ns_to_ticks(long long time_ns, double rate)
{
    const long long ratell = (long long)rate;
    const long long full = time_ns / 1000000000LL;
    const long long err = time_ns - full * 1000000000LL;
    const double part = (double)full * (rate - (double)ratell);
    const double frac = part + ((double)err * rate) / 1000000000.0;
    return full * ratell + llround(frac);
}

/* SoapySX.cpp:451, :464-466. */
void sxo_alsa_sizes(unsigned long period_arg, unsigned long *period, unsigned long *buffer)
{
    const unsigned long cap = 65536;
    unsigned long p = period_arg > 0 ? period_arg : 256;
    if (p > cap)
        p = cap;
    *period = p;
    *buffer = cap / p * p;
}

/* SoapySX.cpp:910-915. */
unsigned long sxo_rx_overrun_skip(long avail, unsigned long buffer, unsigned long period)
{
    if (avail <= (long)buffer)
        return 0;
    unsigned long overwritten = (unsigned long)avail - buffer;
    return (overwritten / period + 2) * period;
}

/* SoapySX.cpp:1032-1035. */
int64_t sxo_tx_underrun_forward(int64_t playback_position, int64_t write_position,
                                unsigned long period)
{
    int64_t diff = playback_position - write_position;
    if (diff <= 0)
        return 0;
    return (diff / (int64_t)period + 2) * (int64_t)period;
}

void sxo_stats_words(const uint32_t *words, size_t nwords, uint64_t base_index, sxo_stats *out)
{
    sxo_stats s;
    memset(&s, 0, sizeof s);
    for (size_t i = 0; i < nwords; i++) {
        uint64_t w = words[i];
        uint64_t idx = base_index + i;
        s.sum += w;
        s.wsum += w * (2 * idx + 1);
        s.x ^= w;
        if (!(idx & 1) && (w & 2))
            s.tx_on++;
        uint32_t top = (uint32_t)w & 0xFFFFFFFCu;
        if (top == 0x7FFFFFFCu || top == 0x80000000u)
            s.rail++;
    }
    s.count = nwords;
    *out = s;
}

static uint64_t splitmix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* frame k: I = low word, Q = high word of splitmix64(seed + (k+1)*golden). */
void sxo_synth_frames(int32_t *dst, uint64_t first_frame, size_t nframes, uint64_t seed)
{
    for (size_t n = 0; n < nframes; n++) {
        uint64_t k = first_frame + n;
        uint64_t z = splitmix64(seed + (k + 1) * 0x9E3779B97F4A7C15ull);
        dst[2 * n] = (int32_t)(uint32_t)z;
        dst[2 * n + 1] = (int32_t)(uint32_t)(z >> 32);
    }
}

/*
 * sx_oracle.h -- CPU oracle for the SoapySX IQ sample path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's hot-path arithmetic
 * (tejeez/sxxcvr, SoapySX/SoapySX.cpp).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * sxxcvr_b200/ links, imports or calls it: the product path is CUDA only.
 *
 * Parity pinning: the reference ships no golden vectors or automated tests for this
 * path (SoapySX/test/README.md:1-4).  The pin is the reference source itself, compiled
 * unmodified into oracle/_ref/libsx_ref.so (see oracle/Makefile) and compared against
 * this restatement word-for-word in tests/test_oracle.py, plus the known-answer vectors
 * under tests/golden/ that were generated from that build (tests/golden/make_golden.py).
 *
 * Semantics where the reference is undefined C++ (float->int32 of a value >= 2^31 or NaN,
 * SoapySX.cpp:124-125): this oracle implements the behaviour of the reference's only
 * deployment target (ARM fcvtzs/vcvt: saturate, NaN -> 0), which is also what CUDA's
 * cvt.rzi.s32.f32 does.  The x86 build of the reference differs there and is compared
 * only on the defined domain.  See DESIGN.md "Parity policy".
 */
#ifndef SX_ORACLE_H
#define SX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SoapySX.cpp:103-112.  Offsets and length are in frames (1 frame = I word + Q word). */
void sxo_convert_rx_buffer(const void *src, size_t src_offset,
                           void *dest, size_t dest_offset, size_t length);

/* SoapySX.cpp:116-137. */
void sxo_convert_tx_buffer(const void *src, size_t src_offset,
                           void *dest, size_t dest_offset, size_t length,
                           float tx_threshold2);

/* Extensions with NO reference implementation (SoapySX.cpp:752-753 rejects everything
 * but CF32).  Specified in DESIGN.md; the oracle here is the specification. */
void sxo_convert_rx_buffer_cs16(const void *src, size_t src_offset,
                                void *dest, size_t dest_offset, size_t length);
void sxo_convert_tx_buffer_cs16(const void *src, size_t src_offset,
                                void *dest, size_t dest_offset, size_t length,
                                float tx_threshold2);

/* Extension with NO reference implementation: 16-bit I2S slots.  The reference's sample-rate
 * table lists the 16-bit SX1255 modes only as commented-out entries that "did not work"
 * (SoapySX.cpp:200-207) and hard-wires SND_PCM_FORMAT_S32_LE (:474).  Specification (ours):
 * an S16 frame is [I:int16][Q:int16]; RX: f = s * 2^-15; TX: v = trunc(2^15 * f) saturated to
 * int16, low two bits cleared, TX-enable = both low bits of I, same un-fused threshold test. */
void sxo_convert_rx_buffer_s16(const void *src, size_t src_offset,
                               void *dest, size_t dest_offset, size_t length);
void sxo_convert_tx_buffer_s16(const void *src, size_t src_offset,
                               void *dest, size_t dest_offset, size_t length,
                               float tx_threshold2);

/* SoapySDR::ticksToTimeNs / timeNsToTicks as used at SoapySX.cpp:564,570.
 * SoapySDR is an external, unpinned dependency (SoapySX/CMakeLists.txt:45); this is a
 * restatement of its published algorithm (lib/TimeC.cpp).  Parity unpinned upstream. */
long long sxo_ticks_to_time_ns(long long ticks, double rate);
long long sxo_time_ns_to_ticks(long long time_ns, double rate);

/* AlsaPcm::configure buffer sizing, SoapySX.cpp:451,464-466. */
void sxo_alsa_sizes(unsigned long period_arg, unsigned long *period, unsigned long *buffer);
/* RX overrun skip, SoapySX.cpp:910-915.  Returns frames to skip (0 if no overrun). */
unsigned long sxo_rx_overrun_skip(long avail, unsigned long buffer, unsigned long period);
/* Untimed TX underrun forward, SoapySX.cpp:1032-1035.  Returns frames added. */
int64_t sxo_tx_underrun_forward(int64_t playback_position, int64_t write_position,
                                unsigned long period);

/* Order-sensitive checksum over 32-bit words (ours; used for full-size parity and
 * the multi-GPU statistics gather).  All fields are sums mod 2^64, so any partition of
 * the word range can be reduced in any order. */
typedef struct {
    uint64_t sum;      /* sum of words */
    uint64_t wsum;     /* sum of word * (2*(base+i)+1) */
    uint64_t x;        /* xor of words (low 32 bits) */
    uint64_t count;    /* number of words */
    uint64_t tx_on;    /* even-index words with bit 1 set (TX-enable flag, :126-133) */
    uint64_t rail;     /* words whose upper 30 bits sit on a rail: 0x7FFFFFFC or 0x80000000 */
} sxo_stats;
void sxo_stats_words(const uint32_t *words, size_t nwords, uint64_t base_index, sxo_stats *out);

/* Synthetic I2S capture frames: frame k of a stream is a pure function of (seed, k).
 * Same function as the device-side generator; it stands in for the SX1255 ADC. */
void sxo_synth_frames(int32_t *dst, uint64_t first_frame, size_t nframes, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif

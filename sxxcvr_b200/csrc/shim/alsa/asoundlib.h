/*
 * Stub of <alsa/asoundlib.h>: the subset of the libasound PCM API that the SoapySX stream
 * path uses (reference SoapySX.cpp:397-504, :755-756, :786, :822-823, :898, :916, :948,
 * :990, :1049-1062, :1093, :1129), backed by a deterministic in-process model of the
 * SX1255 I2S capture/playback pair instead of hardware.  There is no SX1255 and no
 * libasound on a GPU box; this is the "stubbed ALSA source/sink fed with synthetic frames".
 *
 * Model (alsa_stub.cpp):
 *  - every PCM has an application pointer (frames read/written/forwarded since prepare)
 *    and a hardware pointer (frames captured/played since start), both 64-bit, no wrap;
 *  - linked PCMs share a virtual sample clock and start/stop together;
 *  - the clock only moves when a blocking call has to wait (it advances by exactly the
 *    deficit) or when a test calls sx_alsa_advance();
 *  - capture frame k is a pure function of (seed, k) or comes from a caller-supplied table;
 *  - playback frames land in a position-indexed timeline; unwritten positions are silence.
 * Type names, enum values and signatures follow alsa-lib so driver code is source
 * compatible with the real header.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct _snd_pcm snd_pcm_t;
typedef struct _snd_pcm_hw_params snd_pcm_hw_params_t;
typedef struct _snd_pcm_sw_params snd_pcm_sw_params_t;
typedef unsigned long snd_pcm_uframes_t;
typedef long snd_pcm_sframes_t;

typedef enum _snd_pcm_stream {
    SND_PCM_STREAM_PLAYBACK = 0,
    SND_PCM_STREAM_CAPTURE,
} snd_pcm_stream_t;

typedef enum _snd_pcm_access {
    SND_PCM_ACCESS_MMAP_INTERLEAVED = 0,
    SND_PCM_ACCESS_MMAP_NONINTERLEAVED,
    SND_PCM_ACCESS_MMAP_COMPLEX,
    SND_PCM_ACCESS_RW_INTERLEAVED,
    SND_PCM_ACCESS_RW_NONINTERLEAVED,
} snd_pcm_access_t;

typedef enum _snd_pcm_format {
    SND_PCM_FORMAT_UNKNOWN = -1,
    SND_PCM_FORMAT_S8 = 0,
    SND_PCM_FORMAT_U8,
    SND_PCM_FORMAT_S16_LE,
    SND_PCM_FORMAT_S16_BE,
    SND_PCM_FORMAT_U16_LE,
    SND_PCM_FORMAT_U16_BE,
    SND_PCM_FORMAT_S24_LE,
    SND_PCM_FORMAT_S24_BE,
    SND_PCM_FORMAT_U24_LE,
    SND_PCM_FORMAT_U24_BE,
    SND_PCM_FORMAT_S32_LE,
    SND_PCM_FORMAT_S32_BE,
} snd_pcm_format_t;

typedef enum _snd_pcm_state {
    SND_PCM_STATE_OPEN = 0,
    SND_PCM_STATE_SETUP,
    SND_PCM_STATE_PREPARED,
    SND_PCM_STATE_RUNNING,
    SND_PCM_STATE_XRUN,
    SND_PCM_STATE_DRAINING,
    SND_PCM_STATE_PAUSED,
    SND_PCM_STATE_SUSPENDED,
    SND_PCM_STATE_DISCONNECTED,
} snd_pcm_state_t;

const char *snd_strerror(int errnum);

int snd_pcm_open(snd_pcm_t **pcm, const char *name, snd_pcm_stream_t stream, int mode);
int snd_pcm_close(snd_pcm_t *pcm);
int snd_pcm_prepare(snd_pcm_t *pcm);
int snd_pcm_reset(snd_pcm_t *pcm);
int snd_pcm_start(snd_pcm_t *pcm);
int snd_pcm_drop(snd_pcm_t *pcm);
snd_pcm_state_t snd_pcm_state(snd_pcm_t *pcm);
int snd_pcm_link(snd_pcm_t *pcm1, snd_pcm_t *pcm2);
int snd_pcm_wait(snd_pcm_t *pcm, int timeout);
int snd_pcm_avail_delay(snd_pcm_t *pcm, snd_pcm_sframes_t *availp, snd_pcm_sframes_t *delayp);
snd_pcm_sframes_t snd_pcm_forwardable(snd_pcm_t *pcm);
snd_pcm_sframes_t snd_pcm_forward(snd_pcm_t *pcm, snd_pcm_uframes_t frames);
snd_pcm_sframes_t snd_pcm_readi(snd_pcm_t *pcm, void *buffer, snd_pcm_uframes_t size);
snd_pcm_sframes_t snd_pcm_writei(snd_pcm_t *pcm, const void *buffer, snd_pcm_uframes_t size);

int snd_pcm_hw_params_malloc(snd_pcm_hw_params_t **ptr);
void snd_pcm_hw_params_free(snd_pcm_hw_params_t *obj);
int snd_pcm_hw_params_any(snd_pcm_t *pcm, snd_pcm_hw_params_t *params);
int snd_pcm_hw_params_set_access(snd_pcm_t *pcm, snd_pcm_hw_params_t *params,
                                 snd_pcm_access_t access);
int snd_pcm_hw_params_set_format(snd_pcm_t *pcm, snd_pcm_hw_params_t *params,
                                 snd_pcm_format_t val);
int snd_pcm_hw_params_set_rate(snd_pcm_t *pcm, snd_pcm_hw_params_t *params, unsigned int val,
                               int dir);
int snd_pcm_hw_params_set_channels(snd_pcm_t *pcm, snd_pcm_hw_params_t *params,
                                   unsigned int val);
int snd_pcm_hw_params_set_buffer_size_near(snd_pcm_t *pcm, snd_pcm_hw_params_t *params,
                                           snd_pcm_uframes_t *val);
int snd_pcm_hw_params_set_period_size_near(snd_pcm_t *pcm, snd_pcm_hw_params_t *params,
                                           snd_pcm_uframes_t *val, int *dir);
int snd_pcm_hw_params_get_periods(const snd_pcm_hw_params_t *params, unsigned int *val,
                                  int *dir);
int snd_pcm_hw_params(snd_pcm_t *pcm, snd_pcm_hw_params_t *params);

int snd_pcm_sw_params_malloc(snd_pcm_sw_params_t **ptr);
void snd_pcm_sw_params_free(snd_pcm_sw_params_t *obj);
int snd_pcm_sw_params_current(snd_pcm_t *pcm, snd_pcm_sw_params_t *params);
int snd_pcm_sw_params_get_boundary(const snd_pcm_sw_params_t *params, snd_pcm_uframes_t *val);
int snd_pcm_sw_params_set_stop_threshold(snd_pcm_t *pcm, snd_pcm_sw_params_t *params,
                                         snd_pcm_uframes_t val);
int snd_pcm_sw_params_set_silence_threshold(snd_pcm_t *pcm, snd_pcm_sw_params_t *params,
                                            snd_pcm_uframes_t val);
int snd_pcm_sw_params_set_silence_size(snd_pcm_t *pcm, snd_pcm_sw_params_t *params,
                                       snd_pcm_uframes_t val);
int snd_pcm_sw_params(snd_pcm_t *pcm, snd_pcm_sw_params_t *params);

/* ------------------------------------------------------------------------------------
 * Stub control surface (not part of alsa-lib): what a test or bench uses in place of the
 * physical HAT.
 * ---------------------------------------------------------------------------------- */

/* Registry of open PCMs, in snd_pcm_open order. */
size_t sx_alsa_pcm_count(void);
snd_pcm_t *sx_alsa_pcm_at(size_t index);
int sx_alsa_pcm_is_capture(snd_pcm_t *pcm);

/* Virtual sample clock: advance every RUNNING PCM linked with `pcm` by `frames`. */
void sx_alsa_advance(snd_pcm_t *pcm, int64_t frames);
/* 1 (default): a blocking call that lacks frames/space advances the clock by the deficit.
 * 0: the call returns what it can right now (short transfer, possibly 0). */
void sx_alsa_set_free_run(snd_pcm_t *pcm, int free_run);
/* Cap frames moved per readi/writei call (0 = no cap): models short transfers. */
void sx_alsa_set_max_transfer(snd_pcm_t *pcm, snd_pcm_uframes_t frames);

/* Capture source.  Default: frame k = sx_synth_frame(seed, k) with seed 0x53581255. */
void sx_alsa_set_capture_seed(snd_pcm_t *pcm, uint64_t seed);
/* Capture frames from a table instead: frame k = table[k % nframes] (copied). */
void sx_alsa_set_capture_table(snd_pcm_t *pcm, const int32_t *frames, size_t nframes);

/* Playback sink.  Frames are stored by absolute position up to `max_frames`
 * (default 1<<22); frames beyond the limit are counted but not kept. */
void sx_alsa_set_sink_limit(snd_pcm_t *pcm, size_t max_frames);
/* Copy `nframes` frames starting at `position` out of the playback timeline
 * (silence where nothing was written).  Returns frames copied. */
size_t sx_alsa_sink_read(snd_pcm_t *pcm, int64_t position, size_t nframes, int32_t *out);
void sx_alsa_sink_clear(snd_pcm_t *pcm);
/* 1 if a frame was written at `position`, 0 if it is silence. */
int sx_alsa_sink_written(snd_pcm_t *pcm, int64_t position);

/* Pointers as the "hardware" sees them. */
int64_t sx_alsa_hw_ptr(snd_pcm_t *pcm);
int64_t sx_alsa_appl_ptr(snd_pcm_t *pcm);
uint64_t sx_alsa_frames_transferred(snd_pcm_t *pcm);
snd_pcm_uframes_t sx_alsa_buffer_size(snd_pcm_t *pcm);
snd_pcm_uframes_t sx_alsa_period_size(snd_pcm_t *pcm);

/* One-shot fault injection: the `skip`-th next call of `op` returns `err` (negative errno). */
typedef enum {
    SX_ALSA_OP_AVAIL_DELAY = 0,
    SX_ALSA_OP_READI,
    SX_ALSA_OP_WRITEI,
    SX_ALSA_OP_FORWARD,
    SX_ALSA_OP_FORWARDABLE,
    SX_ALSA_OP_START,
    SX_ALSA_OP_COUNT_
} sx_alsa_op;
void sx_alsa_inject_error(snd_pcm_t *pcm, sx_alsa_op op, int err, unsigned skip);

#ifdef __cplusplus
}
#endif

// Deterministic in-process model of the SX1255 I2S capture/playback PCM pair, exposed
// through the libasound subset declared in shim/alsa/asoundlib.h.  See that header for
// the model.  This replaces the hardware I/O side of the reference driver
// (AlsaPcm, SoapySX.cpp:369-518); it never converts samples.
#include <alsa/asoundlib.h>

#include "../sx_synth.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <climits>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {

// Linked PCMs share one of these (snd_pcm_link, reference SoapySX.cpp:786).  The group is also
// the unit of locking: PCMs that share a clock are serialised against each other, PCMs of
// different devices are not -- an RX thread and a TX thread of one device contend exactly as
// two threads on one sound card would, and N devices on N threads do not contend at all.
struct ClockGroup {
    std::recursive_mutex mutex;
    std::vector<snd_pcm_t *> members;
};

struct Fault {
    int err = 0;
    unsigned skip = 0;
    bool armed = false;
};

const snd_pcm_uframes_t kBoundary = snd_pcm_uframes_t(1) << 62;
const snd_pcm_uframes_t kMaxBuffer = 65536; // observed Pi limit, reference :464
const unsigned kWaitWatchdog = 1000;

} // namespace

struct _snd_pcm_hw_params {
    snd_pcm_uframes_t buffer_size = kMaxBuffer;
    snd_pcm_uframes_t period_size = 256;
    snd_pcm_format_t format = SND_PCM_FORMAT_S32_LE;
    unsigned channels = 2;
    unsigned rate = 192000;
    snd_pcm_access_t access = SND_PCM_ACCESS_RW_INTERLEAVED;
};

struct _snd_pcm_sw_params {
    snd_pcm_uframes_t boundary = kBoundary;
    snd_pcm_uframes_t start_threshold = 1;
    snd_pcm_uframes_t stop_threshold = kMaxBuffer;
    snd_pcm_uframes_t silence_threshold = 0;
    snd_pcm_uframes_t silence_size = 0;
    snd_pcm_uframes_t avail_min = 256;
};

struct _snd_pcm {
    std::string name;
    snd_pcm_stream_t dir = SND_PCM_STREAM_PLAYBACK;
    snd_pcm_state_t state = SND_PCM_STATE_OPEN;
    _snd_pcm_hw_params hw;
    _snd_pcm_sw_params sw;
    std::shared_ptr<ClockGroup> group;           // ownership
    std::atomic<ClockGroup *> group_now{nullptr}; // what Guard reads: the same object, without a lock

    int64_t appl_ptr = 0; // frames the application has consumed / produced / forwarded
    int64_t hw_ptr = 0;   // frames the "hardware" has captured / played

    bool free_run = true;
    unsigned fruitless_waits = 0; // consecutive snd_pcm_wait calls that timed out (watchdog, see snd_pcm_wait)
    snd_pcm_uframes_t max_transfer = 0;
    uint64_t transferred = 0;

    // capture source
    uint64_t seed = SX_SYNTH_DEFAULT_SEED;
    std::vector<uint64_t> table;

    // playback sink
    size_t sink_limit = size_t(1) << 22;
    std::vector<uint64_t> timeline;
    std::vector<uint8_t> written;
    uint64_t dropped_beyond_limit = 0;

    Fault faults[SX_ALSA_OP_COUNT_];

    bool capture() const { return dir == SND_PCM_STREAM_CAPTURE; }
    // Frames ready to read (capture) or free to write (playback).
    int64_t avail() const
    {
        return capture() ? hw_ptr - appl_ptr : hw_ptr + int64_t(hw.buffer_size) - appl_ptr;
    }
    // Frames between the application and the hardware: queued (playback) or pending (capture).
    int64_t delay() const { return capture() ? hw_ptr - appl_ptr : appl_ptr - hw_ptr; }
};

namespace {

std::mutex g_registry;   // guards g_open
std::mutex g_link_mutex; // one snd_pcm_link at a time (it needs two groups)
std::vector<snd_pcm_t *> g_open;
// Groups emptied by snd_pcm_link are parked here instead of being freed, so that a thread which
// read a PCM's old group pointer just before the link can still lock it, notice, and retry.
std::vector<std::shared_ptr<ClockGroup>> g_retired_groups; // guarded by g_link_mutex

// Holds the lock of the clock group `pcm` belongs to.  No process-wide lock is touched on this
// path (thousands of PCMs on a handful of threads would all meet there): the PCM's group is read
// from an atomic pointer, locked, and checked again -- the pointer only changes inside
// snd_pcm_link, which holds both groups' locks while it re-points the members, so a group that is
// still the PCM's group after its lock was taken stays so until the lock is released.
class Guard {
public:
    explicit Guard(snd_pcm_t *pcm)
    {
        for (;;) {
            group_ = pcm->group_now.load(std::memory_order_acquire);
            group_->mutex.lock();
            if (pcm->group_now.load(std::memory_order_acquire) == group_)
                return;
            group_->mutex.unlock();
        }
    }
    ~Guard() { group_->mutex.unlock(); }
    Guard(const Guard &) = delete;
    Guard &operator=(const Guard &) = delete;

private:
    ClockGroup *group_;
};

bool takeFault(snd_pcm_t *pcm, sx_alsa_op op, int *err)
{
    Fault &f = pcm->faults[op];
    if (!f.armed)
        return false;
    if (f.skip > 0) {
        f.skip--;
        return false;
    }
    f.armed = false;
    *err = f.err;
    return true;
}

// A stream whose stop threshold is below the boundary stops on xrun and takes its linked
// partner with it (STREAM_MODE_LINK, reference :29-44, :498).
void checkXrun(snd_pcm_t *pcm)
{
    if (pcm->state != SND_PCM_STATE_RUNNING)
        return;
    if (pcm->sw.stop_threshold >= pcm->sw.boundary)
        return;
    if (pcm->avail() >= int64_t(pcm->sw.stop_threshold)) {
        for (snd_pcm_t *m : pcm->group->members)
            if (m->state == SND_PCM_STATE_RUNNING)
                m->state = SND_PCM_STATE_XRUN;
    }
}

void advanceGroup(snd_pcm_t *pcm, int64_t frames)
{
    if (frames <= 0)
        return;
    for (snd_pcm_t *m : pcm->group->members)
        if (m->state == SND_PCM_STATE_RUNNING)
            m->hw_ptr += frames;
    // Evaluate xruns only after every member has moved, so the result does not depend
    // on member order.
    for (snd_pcm_t *m : pcm->group->members)
        checkXrun(m);
}

int startGroup(snd_pcm_t *pcm)
{
    if (pcm->state != SND_PCM_STATE_PREPARED)
        return -EBADFD;
    for (snd_pcm_t *m : pcm->group->members)
        if (m->state == SND_PCM_STATE_PREPARED)
            m->state = SND_PCM_STATE_RUNNING;
    return 0;
}

// Frames first .. first + n - 1 of the capture stream.  A table is a periodic signal: whole
// runs are copied, as snd_pcm_readi copies out of the DMA ring.
void captureFrames(const snd_pcm_t *pcm, int64_t first, uint64_t *out, size_t n)
{
    if (pcm->table.empty()) {
        for (size_t i = 0; i < n; i++)
            out[i] = sx_synth_frame(pcm->seed, uint64_t(first) + i);
        return;
    }
    const size_t size = pcm->table.size();
    size_t at = size_t(uint64_t(first) % size);
    while (n > 0) {
        const size_t run = std::min(n, size - at);
        std::memcpy(out, &pcm->table[at], run * sizeof(uint64_t));
        out += run;
        n -= run;
        at = 0;
    }
}

// Frames [position, position + n) go into the timeline; what falls outside [0, sink_limit) is
// only counted.  One copy for the part that is kept, as the DMA ring of a sound card would take it.
void sinkStore(snd_pcm_t *pcm, int64_t position, const uint64_t *frames, size_t n)
{
    const int64_t lo = std::max<int64_t>(position, 0);
    const int64_t hi = std::min<int64_t>(position + int64_t(n), int64_t(std::min<size_t>(pcm->sink_limit, size_t(INT64_MAX))));
    const size_t kept = hi > lo ? size_t(hi - lo) : 0;
    pcm->dropped_beyond_limit += n - kept;
    if (kept == 0)
        return;
    if (size_t(hi) > pcm->timeline.size()) {
        size_t grown = std::min(pcm->sink_limit, std::max(size_t(hi), pcm->timeline.size() * 2));
        pcm->timeline.resize(grown, 0);
        pcm->written.resize(grown, 0);
    }
    std::memcpy(&pcm->timeline[size_t(lo)], frames + (lo - position), kept * sizeof(uint64_t));
    std::memset(&pcm->written[size_t(lo)], 1, kept);
}

} // namespace

extern "C" {

const char *snd_strerror(int errnum)
{
    return std::strerror(errnum < 0 ? -errnum : errnum);
}

int snd_pcm_open(snd_pcm_t **out, const char *name, snd_pcm_stream_t stream, int)
{
    snd_pcm_t *pcm = new _snd_pcm();
    pcm->name = name ? name : "";
    pcm->dir = stream;
    pcm->group = std::make_shared<ClockGroup>();
    pcm->group->members.push_back(pcm);
    pcm->group_now.store(pcm->group.get(), std::memory_order_release);
    std::lock_guard<std::mutex> r(g_registry);
    g_open.push_back(pcm);
    *out = pcm;
    return 0;
}

int snd_pcm_close(snd_pcm_t *pcm)
{
    std::shared_ptr<ClockGroup> keep; // the group may lose its last owner here: not while it is locked
    {
        Guard lock(pcm);
        keep = pcm->group;
        std::lock_guard<std::mutex> r(g_registry);
        auto &members = pcm->group->members;
        members.erase(std::remove(members.begin(), members.end(), pcm), members.end());
        g_open.erase(std::remove(g_open.begin(), g_open.end(), pcm), g_open.end());
    }
    delete pcm;
    return 0;
}

// Like the kernel's snd_pcm_action_nonatomic, prepare and reset act on the whole linked
// group.  The playback timeline is kept (tests read it after deactivation); use
// sx_alsa_sink_clear() to wipe it.
int snd_pcm_prepare(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    if (pcm->state == SND_PCM_STATE_OPEN)
        return -EBADFD;
    for (snd_pcm_t *m : pcm->group->members)
        if (m->state == SND_PCM_STATE_RUNNING)
            return -EBUSY;
    for (snd_pcm_t *m : pcm->group->members) {
        if (m->state == SND_PCM_STATE_OPEN)
            continue;
        m->state = SND_PCM_STATE_PREPARED;
        m->appl_ptr = 0;
        m->hw_ptr = 0;
    }
    return 0;
}

// appl_ptr = hw_ptr: forget queued / pending frames.
int snd_pcm_reset(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    if (pcm->state != SND_PCM_STATE_RUNNING && pcm->state != SND_PCM_STATE_PREPARED)
        return -EBADFD;
    for (snd_pcm_t *m : pcm->group->members)
        if (m->state == SND_PCM_STATE_RUNNING || m->state == SND_PCM_STATE_PREPARED)
            m->appl_ptr = m->hw_ptr;
    return 0;
}

int snd_pcm_start(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    int err;
    if (takeFault(pcm, SX_ALSA_OP_START, &err))
        return err;
    return startGroup(pcm);
}

// Stops this stream and everything linked to it; pending frames are dropped.
int snd_pcm_drop(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    if (pcm->state == SND_PCM_STATE_OPEN)
        return -EBADFD;
    for (snd_pcm_t *m : pcm->group->members)
        if (m->state == SND_PCM_STATE_RUNNING || m->state == SND_PCM_STATE_XRUN ||
            m->state == SND_PCM_STATE_PREPARED)
            m->state = SND_PCM_STATE_SETUP;
    return 0;
}

snd_pcm_state_t snd_pcm_state(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    return pcm->state;
}

int snd_pcm_link(snd_pcm_t *a, snd_pcm_t *b)
{
    std::lock_guard<std::mutex> one_link(g_link_mutex);
    Guard lock_a(a);
    if (a->group == b->group)
        return -EALREADY;
    Guard lock_b(b);
    std::shared_ptr<ClockGroup> old = b->group;
    for (snd_pcm_t *m : old->members) {
        m->group = a->group;
        m->group_now.store(a->group.get(), std::memory_order_release);
        a->group->members.push_back(m);
    }
    old->members.clear();
    g_retired_groups.push_back(old); // lock_b still holds its mutex: it must outlive this call
    return 0;
}

// Waits until avail >= avail_min.  In free-run mode the wait is the clock advancing.
//
// Watchdog: when the clock cannot move (free-run off, or the stream is not running) the wait
// times out, and a driver that forwards-and-waits in a loop (reference SoapySX.cpp:1045-1073)
// would spin for ever -- on real hardware it would sit in 10 s waits instead.  After
// kWaitWatchdog fruitless waits in a row the next snd_pcm_forwardable reports -EIO, so the loop
// ends with a stream error and no test can hang.
int snd_pcm_wait(snd_pcm_t *pcm, int)
{
    Guard lock(pcm);
    if (pcm->state == SND_PCM_STATE_XRUN)
        return -EPIPE;
    int64_t need = int64_t(pcm->sw.avail_min) - pcm->avail();
    if (need > 0) {
        if (!pcm->free_run || pcm->state != SND_PCM_STATE_RUNNING) {
            pcm->fruitless_waits++;
            return 0; // timed out
        }
        advanceGroup(pcm, need);
        if (pcm->state == SND_PCM_STATE_XRUN)
            return -EPIPE;
    }
    pcm->fruitless_waits = 0;
    return 1;
}

int snd_pcm_avail_delay(snd_pcm_t *pcm, snd_pcm_sframes_t *availp, snd_pcm_sframes_t *delayp)
{
    Guard lock(pcm);
    int err;
    if (takeFault(pcm, SX_ALSA_OP_AVAIL_DELAY, &err))
        return err;
    if (pcm->state == SND_PCM_STATE_XRUN)
        return -EPIPE;
    if (pcm->state == SND_PCM_STATE_OPEN)
        return -EBADFD;
    *availp = snd_pcm_sframes_t(pcm->avail());
    *delayp = snd_pcm_sframes_t(pcm->delay());
    return 0;
}

snd_pcm_sframes_t snd_pcm_forwardable(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    int err;
    if (takeFault(pcm, SX_ALSA_OP_FORWARDABLE, &err))
        return err;
    if (pcm->fruitless_waits >= kWaitWatchdog) {
        pcm->fruitless_waits = 0;
        return -EIO;
    }
    if (pcm->state == SND_PCM_STATE_XRUN)
        return -EPIPE;
    return snd_pcm_sframes_t(std::max<int64_t>(pcm->avail(), 0));
}

snd_pcm_sframes_t snd_pcm_forward(snd_pcm_t *pcm, snd_pcm_uframes_t frames)
{
    Guard lock(pcm);
    int err;
    if (takeFault(pcm, SX_ALSA_OP_FORWARD, &err))
        return err;
    if (pcm->state == SND_PCM_STATE_XRUN)
        return -EPIPE;
    int64_t room = std::max<int64_t>(pcm->avail(), 0);
    int64_t moved = std::min<int64_t>(room, int64_t(std::min<snd_pcm_uframes_t>(frames, LONG_MAX)));
    pcm->appl_ptr += moved;
    return snd_pcm_sframes_t(moved);
}

snd_pcm_sframes_t snd_pcm_readi(snd_pcm_t *pcm, void *buffer, snd_pcm_uframes_t size)
{
    Guard lock(pcm);
    int err;
    if (takeFault(pcm, SX_ALSA_OP_READI, &err))
        return err;
    if (!pcm->capture())
        return -EBADFD;
    if (pcm->state == SND_PCM_STATE_PREPARED && size >= pcm->sw.start_threshold)
        startGroup(pcm);
    if (pcm->state == SND_PCM_STATE_XRUN)
        return -EPIPE;
    if (pcm->state != SND_PCM_STATE_RUNNING)
        return -EBADFD;

    int64_t want = int64_t(std::min<snd_pcm_uframes_t>(size, LONG_MAX));
    if (pcm->max_transfer > 0)
        want = std::min<int64_t>(want, int64_t(pcm->max_transfer));
    int64_t have = std::max<int64_t>(pcm->avail(), 0);
    if (have < want && pcm->free_run) {
        advanceGroup(pcm, want - have);
        if (pcm->state == SND_PCM_STATE_XRUN)
            return -EPIPE;
        have = std::max<int64_t>(pcm->avail(), 0);
    }
    int64_t n = std::min(want, have);
    captureFrames(pcm, pcm->appl_ptr, static_cast<uint64_t *>(buffer), size_t(n));
    pcm->appl_ptr += n;
    pcm->transferred += uint64_t(n);
    return snd_pcm_sframes_t(n);
}

snd_pcm_sframes_t snd_pcm_writei(snd_pcm_t *pcm, const void *buffer, snd_pcm_uframes_t size)
{
    Guard lock(pcm);
    int err;
    if (takeFault(pcm, SX_ALSA_OP_WRITEI, &err))
        return err;
    if (pcm->capture())
        return -EBADFD;
    if (pcm->state == SND_PCM_STATE_XRUN)
        return -EPIPE;
    if (pcm->state != SND_PCM_STATE_RUNNING && pcm->state != SND_PCM_STATE_PREPARED)
        return -EBADFD;

    int64_t want = int64_t(std::min<snd_pcm_uframes_t>(size, LONG_MAX));
    if (pcm->max_transfer > 0)
        want = std::min<int64_t>(want, int64_t(pcm->max_transfer));
    const uint64_t *in = static_cast<const uint64_t *>(buffer);
    int64_t done = 0;
    while (done < want) {
        int64_t room = std::max<int64_t>(pcm->avail(), 0);
        int64_t n = std::min(want - done, room);
        if (n > 0) {
            sinkStore(pcm, pcm->appl_ptr, in + done, size_t(n));
            pcm->appl_ptr += n;
            done += n;
            // Default start threshold is 1 frame: the first write starts the stream
            // (and its linked capture partner), reference :39, test_linked_streams.py:36.
            if (pcm->state == SND_PCM_STATE_PREPARED &&
                pcm->delay() >= int64_t(pcm->sw.start_threshold))
                startGroup(pcm);
        }
        if (done == want)
            break;
        if (!pcm->free_run || pcm->state != SND_PCM_STATE_RUNNING)
            break; // short write
        advanceGroup(pcm, want - done - std::max<int64_t>(pcm->avail(), 0));
        if (pcm->state == SND_PCM_STATE_XRUN)
            return done > 0 ? snd_pcm_sframes_t(done) : -EPIPE;
    }
    pcm->transferred += uint64_t(done);
    return snd_pcm_sframes_t(done);
}

// ---- hw params ------------------------------------------------------------------------
int snd_pcm_hw_params_malloc(snd_pcm_hw_params_t **ptr)
{
    *ptr = new _snd_pcm_hw_params();
    return 0;
}
void snd_pcm_hw_params_free(snd_pcm_hw_params_t *obj) { delete obj; }
int snd_pcm_hw_params_any(snd_pcm_t *, snd_pcm_hw_params_t *params)
{
    *params = _snd_pcm_hw_params();
    return 0;
}
int snd_pcm_hw_params_set_access(snd_pcm_t *, snd_pcm_hw_params_t *params, snd_pcm_access_t access)
{
    if (access != SND_PCM_ACCESS_RW_INTERLEAVED)
        return -EINVAL;
    params->access = access;
    return 0;
}
int snd_pcm_hw_params_set_format(snd_pcm_t *, snd_pcm_hw_params_t *params, snd_pcm_format_t val)
{
    // The I2S link carries two 32-bit slots (dts/sx1255_raspberrypi.dts:58-59).
    if (val != SND_PCM_FORMAT_S32_LE)
        return -EINVAL;
    params->format = val;
    return 0;
}
int snd_pcm_hw_params_set_rate(snd_pcm_t *, snd_pcm_hw_params_t *params, unsigned int val, int)
{
    params->rate = val;
    return 0;
}
int snd_pcm_hw_params_set_channels(snd_pcm_t *, snd_pcm_hw_params_t *params, unsigned int val)
{
    if (val != 2)
        return -EINVAL;
    params->channels = val;
    return 0;
}
int snd_pcm_hw_params_set_buffer_size_near(snd_pcm_t *, snd_pcm_hw_params_t *params,
                                           snd_pcm_uframes_t *val)
{
    *val = std::max<snd_pcm_uframes_t>(std::min(*val, kMaxBuffer), 2);
    params->buffer_size = *val;
    return 0;
}
int snd_pcm_hw_params_set_period_size_near(snd_pcm_t *, snd_pcm_hw_params_t *params,
                                           snd_pcm_uframes_t *val, int *)
{
    *val = std::max<snd_pcm_uframes_t>(std::min(*val, params->buffer_size), 1);
    params->period_size = *val;
    return 0;
}
int snd_pcm_hw_params_get_periods(const snd_pcm_hw_params_t *params, unsigned int *val, int *)
{
    *val = unsigned(params->buffer_size / params->period_size);
    return 0;
}
int snd_pcm_hw_params(snd_pcm_t *pcm, snd_pcm_hw_params_t *params)
{
    Guard lock(pcm);
    if (pcm->state == SND_PCM_STATE_RUNNING)
        return -EBUSY;
    pcm->hw = *params;
    pcm->sw.avail_min = params->period_size;
    pcm->sw.stop_threshold = params->buffer_size;
    // alsa-lib prepares the stream as part of installing hw params.
    pcm->state = SND_PCM_STATE_PREPARED;
    pcm->appl_ptr = 0;
    pcm->hw_ptr = 0;
    return 0;
}

// ---- sw params ------------------------------------------------------------------------
int snd_pcm_sw_params_malloc(snd_pcm_sw_params_t **ptr)
{
    *ptr = new _snd_pcm_sw_params();
    return 0;
}
void snd_pcm_sw_params_free(snd_pcm_sw_params_t *obj) { delete obj; }
int snd_pcm_sw_params_current(snd_pcm_t *pcm, snd_pcm_sw_params_t *params)
{
    Guard lock(pcm);
    *params = pcm->sw;
    return 0;
}
int snd_pcm_sw_params_get_boundary(const snd_pcm_sw_params_t *params, snd_pcm_uframes_t *val)
{
    *val = params->boundary;
    return 0;
}
int snd_pcm_sw_params_set_stop_threshold(snd_pcm_t *, snd_pcm_sw_params_t *params,
                                         snd_pcm_uframes_t val)
{
    params->stop_threshold = val;
    return 0;
}
int snd_pcm_sw_params_set_silence_threshold(snd_pcm_t *, snd_pcm_sw_params_t *params,
                                            snd_pcm_uframes_t val)
{
    params->silence_threshold = val;
    return 0;
}
int snd_pcm_sw_params_set_silence_size(snd_pcm_t *, snd_pcm_sw_params_t *params,
                                       snd_pcm_uframes_t val)
{
    params->silence_size = val;
    return 0;
}
int snd_pcm_sw_params(snd_pcm_t *pcm, snd_pcm_sw_params_t *params)
{
    Guard lock(pcm);
    pcm->sw = *params;
    return 0;
}

// ---- stub control surface ---------------------------------------------------------------
size_t sx_alsa_pcm_count(void)
{
    std::lock_guard<std::mutex> r(g_registry);
    return g_open.size();
}
snd_pcm_t *sx_alsa_pcm_at(size_t index)
{
    std::lock_guard<std::mutex> r(g_registry);
    return index < g_open.size() ? g_open[index] : nullptr;
}
int sx_alsa_pcm_is_capture(snd_pcm_t *pcm) { return pcm->capture() ? 1 : 0; }

void sx_alsa_advance(snd_pcm_t *pcm, int64_t frames)
{
    Guard lock(pcm);
    advanceGroup(pcm, frames);
}
void sx_alsa_set_free_run(snd_pcm_t *pcm, int free_run)
{
    Guard lock(pcm);
    pcm->free_run = free_run != 0;
}
void sx_alsa_set_max_transfer(snd_pcm_t *pcm, snd_pcm_uframes_t frames)
{
    Guard lock(pcm);
    pcm->max_transfer = frames;
}
void sx_alsa_set_capture_seed(snd_pcm_t *pcm, uint64_t seed)
{
    Guard lock(pcm);
    pcm->seed = seed;
    pcm->table.clear();
}
void sx_alsa_set_capture_table(snd_pcm_t *pcm, const int32_t *frames, size_t nframes)
{
    Guard lock(pcm);
    pcm->table.resize(nframes);
    std::memcpy(pcm->table.data(), frames, nframes * sizeof(uint64_t));
}
void sx_alsa_set_sink_limit(snd_pcm_t *pcm, size_t max_frames)
{
    Guard lock(pcm);
    pcm->sink_limit = max_frames;
}
size_t sx_alsa_sink_read(snd_pcm_t *pcm, int64_t position, size_t nframes, int32_t *out)
{
    Guard lock(pcm);
    uint64_t *o = reinterpret_cast<uint64_t *>(out);
    std::memset(o, 0, nframes * sizeof(uint64_t));
    const int64_t lo = std::max<int64_t>(position, 0);
    const int64_t hi = std::min<int64_t>(position + int64_t(nframes), int64_t(pcm->timeline.size()));
    if (hi > lo)
        std::memcpy(o + (lo - position), &pcm->timeline[size_t(lo)], size_t(hi - lo) * sizeof(uint64_t));
    return nframes;
}
void sx_alsa_sink_clear(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    pcm->timeline.clear();
    pcm->written.clear();
    pcm->dropped_beyond_limit = 0;
}
int sx_alsa_sink_written(snd_pcm_t *pcm, int64_t position)
{
    Guard lock(pcm);
    return (position >= 0 && size_t(position) < pcm->written.size()) ? pcm->written[size_t(position)] : 0;
}
int64_t sx_alsa_hw_ptr(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    return pcm->hw_ptr;
}
int64_t sx_alsa_appl_ptr(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    return pcm->appl_ptr;
}
uint64_t sx_alsa_frames_transferred(snd_pcm_t *pcm)
{
    Guard lock(pcm);
    return pcm->transferred;
}
snd_pcm_uframes_t sx_alsa_buffer_size(snd_pcm_t *pcm) { return pcm->hw.buffer_size; }
snd_pcm_uframes_t sx_alsa_period_size(snd_pcm_t *pcm) { return pcm->hw.period_size; }

void sx_alsa_inject_error(snd_pcm_t *pcm, sx_alsa_op op, int err, unsigned skip)
{
    Guard lock(pcm);
    if (op < 0 || op >= SX_ALSA_OP_COUNT_)
        return;
    pcm->faults[op].err = err;
    pcm->faults[op].skip = skip;
    pcm->faults[op].armed = true;
}

} // extern "C"

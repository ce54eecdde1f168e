// sx_resident.cuh -- a resident converter for period-sized blocks.
//
// The reference's native unit of work is one ALSA period (256 frames = 2 KiB, SoapySX.cpp:451)
// per readStream/writeStream call.  For such a block a kernel launch plus a stream
// synchronisation (~14 us) dwarfs the conversion.  This kernel stays resident on one SM and
// is driven through a mailbox in pinned, device-mapped host memory:
//
//   host:   fill the descriptor, then store request = ++seq           (x86 stores stay ordered)
//   device: thread 0 polls `request` with ld.cv (never from a stale cache line); on a new value
//           the CTA converts the block with 8-byte-per-frame ld.cv / st.wt accesses straight
//           from and to the host buffers across PCIe, fences to system scope, and thread 0
//           stores done = seq
//   host:   polls `done`.
//
// The kernel can never pin the GPU.  It leaves (a) after kResidentIdleNs without work, (b) at
// once when the host sets `quit` -- the library does so before every device-wide
// synchronisation of its own (sxgpu_destroy, sxgpu_bank_destroy, growing the staging ring) --
// and (c) after kResidentLifeNs in total even when it is busy, so that a device-wide
// synchronisation the APPLICATION issues (cudaDeviceSynchronize, cudaFree, another library's)
// waits for at most that long while a stream calls in every period.  The host relaunches it on
// the next small call.  `alive` tells the host whether a kernel is listening; the exit
// handshake is: clear `alive`, fence, poll once more and serve a request that raced in.
#pragma once

#include "sx_kernels.cuh"

namespace sx {

struct alignas(64) Mailbox {
    // Written by the host: three 16-byte pieces of one cache line, fetched by the device with
    // three independent loads (one PCIe round trip, not three in a row).  The loads may be
    // served at different moments, so every piece carries the request's sequence number (or its
    // low 15 bits) in the word the host stores LAST within that piece: a piece whose tag is the
    // new number holds the new request's fields (x86 stores become visible in program order), and
    // a request is taken only when all three tags agree.  No second look is needed.
    unsigned long long request; // [piece A] sequence number of the latest request (stored last of all)
    const char *src;            // [piece A] device-visible address of the source (stored before `request`)
    char *dst;                  // [piece B]
    unsigned long long packed;  // [piece B] bits 0..15: frames; bit 16: op (0 = RX S32->CF32, 1 = TX CF32->S32);
                                //           bits 17..31: low 15 bits of the sequence number;
                                //           bits 32..63: tx_threshold2 as float bits (stored after dst)
    unsigned long long quit;    // [piece C] non-zero: leave now (set by the host around device-wide syncs)
    unsigned long long request2; // [piece C] the sequence number again
    unsigned long long pad0[2];
    // written by the device (own cache line)
    unsigned long long done; // sequence number of the last request served
    unsigned long long served;
    int alive;
    int pad1[11];
};

constexpr unsigned long long kResidentIdleNs = 2000000ull;  // 2 ms without a request
constexpr unsigned long long kResidentLifeNs = 20000000ull; // 20 ms in total, busy or not

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.global.wt.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ Pack<2> ld_sys_frame(const void *p)
{
    Pack<2> r;
    asm volatile("ld.global.cv.v2.b32 {%0,%1}, [%2];" : "=r"(r.w[0]), "=r"(r.w[1]) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_sys_frame(void *p, const Pack<2> &v)
{
    asm volatile("st.global.wt.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]) : "memory");
}
struct Request {
    unsigned long long seq;
    const char *src;
    char *dst;
    unsigned int nframes;
    unsigned int thr2_bits;
    int op;
    unsigned long long quit;
};
// The record as three 16-byte loads.  r.seq is the new sequence number only if every piece
// belongs to it; otherwise it is reported as `last_seen` (nothing new yet, look again).
__device__ __forceinline__ Request ld_request(const Mailbox *box, unsigned long long last_seen)
{
    unsigned long long a, b, c, d, q, a2;
    asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(box) : "memory");
    asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];"
                 : "=l"(c), "=l"(d)
                 : "l"(reinterpret_cast<const char *>(box) + 16)
                 : "memory");
    asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];"
                 : "=l"(q), "=l"(a2)
                 : "l"(reinterpret_cast<const char *>(box) + 32)
                 : "memory");
    Request r;
    r.quit = q;
    r.src = reinterpret_cast<const char *>(b);
    r.dst = reinterpret_cast<char *>(c);
    r.nframes = unsigned(d) & 0xFFFFu;
    r.op = int((unsigned(d) >> 16) & 1u);
    r.thr2_bits = unsigned(d >> 32);
    const bool whole = a == a2 && ((unsigned(d) >> 17) & 0x7FFFu) == unsigned(a & 0x7FFFu);
    r.seq = whole ? a : last_seen;
    return r;
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <class Op>
__device__ __forceinline__ void resident_convert(const char *src, char *dst, unsigned long long n, float thr2)
{
    // Both sides are 8-byte frames; four independent loads per thread before the first use.
    constexpr int U = 4;
    for (unsigned long long base = 0; base < n; base += (unsigned long long)blockDim.x * U) {
        Pack<2> in[U], out[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            unsigned long long f = base + threadIdx.x + (unsigned long long)j * blockDim.x;
            if (f < n)
                in[j] = ld_sys_frame(src + f * 8);
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
            unsigned long long f = base + threadIdx.x + (unsigned long long)j * blockDim.x;
            if (f < n) {
                Op::template apply<1>(in[j], out[j], thr2);
                st_sys_frame(dst + f * 8, out[j]);
            }
        }
    }
}

__global__ void __launch_bounds__(256) resident_kernel(Mailbox *box, unsigned long long last_seen)
{
    __shared__ int s_found, s_leave;
    __shared__ unsigned long long s_seq, s_n;
    __shared__ const char *s_src;
    __shared__ char *s_dst;
    __shared__ float s_thr2;
    __shared__ int s_op;

    unsigned long long idle_since = 0, born = 0; // thread 0 only
    if (threadIdx.x == 0) {
        box->alive = 1;
        __threadfence_system();
        idle_since = born = global_timer_ns();
    }
    for (;;) {
        if (threadIdx.x == 0) {
            // Poll the doorbell.  On idle timeout: clear `alive` (so the host knows nobody will
            // answer any more), then look exactly once more for a request that raced in.
            bool leaving = false, found = false;
            Request req;
            for (;;) {
                req = ld_request(box, last_seen);
                if (req.seq != last_seen) {
                    found = true;
                    break;
                }
                if (leaving)
                    break;
                const unsigned long long now = global_timer_ns();
                if (req.quit || now - idle_since > kResidentIdleNs || now - born > kResidentLifeNs) {
                    box->alive = 0;
                    __threadfence_system();
                    leaving = true;
                }
            }
            if (found) {
                s_src = req.src;
                s_dst = req.dst;
                s_n = req.nframes;
                s_thr2 = __uint_as_float(req.thr2_bits);
                s_op = req.op;
                s_seq = req.seq;
            }
            s_found = found ? 1 : 0;
            s_leave = leaving ? 1 : 0;
        }
        __syncthreads();
        const int found = s_found, leave = s_leave;
        if (!found)
            return; // idle timeout and nothing raced in
        const unsigned long long seq = s_seq;
        if (s_op == 0)
            resident_convert<RxCf32>(s_src, s_dst, s_n, s_thr2);
        else
            resident_convert<TxCf32>(s_src, s_dst, s_n, s_thr2);
        __threadfence_system(); // every thread's stores reach the host before the flag does
        __syncthreads();        // ... and every thread is done with the shared descriptor
        if (threadIdx.x == 0) {
            st_sys_u64(&box->done, seq);
            __threadfence_system();
            last_seen = seq;
            idle_since = global_timer_ns();
        }
        if (leave)
            return; // `alive` is already clear; the host relaunches for the next block
        // (a kernel past its lifetime leaves through the poll loop above: it clears `alive`,
        // looks once more, and goes)
    }
}

} // namespace sx

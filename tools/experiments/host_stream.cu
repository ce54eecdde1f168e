// host_stream.cu -- mid-size host-buffer calls (2^20 .. 2^24 frames, both sides pinned): can
// anything beat the library's chunk pipeline?
//
// The library's *_host pipeline queues, per chunk, a host-to-device copy, an event, a kernel (which
// up to 2^23 frames writes its output straight into pinned host memory) and another event, with a
// quarter of the block per chunk: 29-31 GB/s each way at 2^20 frames, 38 at 2^22, against 46 at
// 2^26.  Schemes timed here, wall clock per synchronous call:
//   pipeline_K_chunks            the library's scheme (kernel writes host memory)
//   pipeline_3_stages_K_chunks   the same with the output copied back by the copy engine
//   streamed_write_value_*       ONE kernel launched first; the copy engine lands chunk after chunk,
//                                each followed by cuStreamWriteValue32 on a counter the kernel's CTAs
//                                wait for; the kernel writes host memory (or, "_to_device", device
//                                memory: that variant shows the counter mechanism costs nothing)
//   streamed_both_copy_engines_* the resident kernel between two copy-engine streams: finished
//                                chunks announced to the device-to-host stream by a counter that
//                                cuStreamWaitValue32 waits for; everything queued up front
//   copies_only_up_then_down_K   no kernel at all: chunk c copied down after chunk c came up
// Result (profiles/r02_host_stream_limits.json): the copies alone take what the library's pipeline
// takes.  The ceiling is the platform's per-copy cost (6-8 us per queued copy), not the schedule.
// Standalone; prints one JSON object per line.  NOT a product path.
#include "sx_kernels.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace sx;

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));             \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

typedef int (*WriteValue32)(cudaStream_t, unsigned long long, unsigned int, unsigned int); // CUresult cuStreamWriteValue32

struct StreamedArgs {
    const char *stage;
    char *dst;
    uint64_t total, chunk;
    uint32_t nchunks, base;
    const uint32_t *flag;
    uint32_t *error;
    float thr2;
    uint32_t sleep_ns;
    uint32_t *arrivals; // [nchunks] CTAs that have finished chunk c (zeroed by the host)
    uint32_t *done;     // raised to base + c + 1 when every CTA has
};

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <class Op> __global__ void __launch_bounds__(256) streamed_kernel(StreamedArgs a)
{
    __shared__ int s_ok;
    const uint32_t tid = blockIdx.x * 256 + threadIdx.x, nthreads = gridDim.x * 256;
    BlockDesc d = {a.stage, a.dst, a.total, a.thr2, 0};
    for (uint32_t c = 0; c < a.nchunks; c++) {
        if (threadIdx.x == 0) {
            const uint32_t want = a.base + c + 1;
            const unsigned long long t0 = globaltimer();
            int ok = 1;
            while (int32_t(ld_acquire(a.flag) - want) < 0) {
                if (globaltimer() - t0 > 2000000000ull) {
                    ok = 0;
                    break;
                }
                if (a.sleep_ns)
                    __nanosleep(a.sleep_ns);
            }
            s_ok = ok;
        }
        __syncthreads();
        if (!s_ok) {
            if (threadIdx.x == 0)
                atomicExch(a.error, 1u);
            return;
        }
        const uint64_t lo = uint64_t(c) * a.chunk;
        const uint64_t hi = lo + a.chunk < a.total ? lo + a.chunk : a.total;
        convert_span<Op>(d, lo, hi, tid, nthreads);
        __syncthreads();
        if (a.done && threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&a.arrivals[c], 1u) == gridDim.x - 1) {
                a.arrivals[c] = 0;
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.done), "r"(a.base + c + 1) : "memory");
            }
        }
    }
}

template <class Op> __global__ void __launch_bounds__(256) chunk_kernel(const char *src, char *dst, uint64_t n, float thr2)
{
    BlockDesc d = {src, dst, n, thr2, 0};
    convert_span<Op>(d, 0, n, blockIdx.x * 256 + threadIdx.x, gridDim.x * 256);
}

static double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main()
{
    WriteValue32 write_value = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", reinterpret_cast<void **>(&write_value), cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        write_value = nullptr;
    WriteValue32 wait_value = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", reinterpret_cast<void **>(&wait_value), cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        wait_value = nullptr;
    printf("{\"cuStreamWriteValue32\": %s, \"cuStreamWaitValue32\": %s}\n", write_value ? "true" : "false", wait_value ? "true" : "false");

    const uint64_t maxn = uint64_t(1) << 24;
    char *h_src, *h_dst, *d_stage;
    uint32_t *d_flag, *d_err, *h_seq;
    CK(cudaHostAlloc(&h_src, maxn * 8, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_dst, maxn * 8, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_seq, 65536 * 4, cudaHostAllocDefault));
    CK(cudaMalloc(&d_stage, maxn * 8));
    char *d_out;
    CK(cudaMalloc(&d_out, maxn * 8));
    uint32_t *d_arrivals, *d_done;
    CK(cudaMalloc(&d_arrivals, 4096 * 4));
    CK(cudaMemset(d_arrivals, 0, 4096 * 4));
    CK(cudaMalloc(&d_done, 4));
    CK(cudaMemset(d_done, 0, 4));
    cudaStream_t s_d2h;
    CK(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
    CK(cudaMalloc(&d_flag, 4));
    CK(cudaMalloc(&d_err, 4));
    CK(cudaMemset(d_flag, 0, 4));
    CK(cudaMemset(d_err, 0, 4));
    for (uint64_t i = 0; i < maxn * 2; i++)
        reinterpret_cast<int32_t *>(h_src)[i] = int32_t(i * 2654435761u);
    cudaStream_t s_copy, s_comp;
    CK(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
    cudaEvent_t ev[64];
    for (auto &e : ev)
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint32_t seq = 0;

    auto check = [&](uint64_t n) {
        for (uint64_t i : {uint64_t(0), n - 1, n / 2, 2 * n - 1}) {
            const float want = 4.656612873077392578125e-10f * float(reinterpret_cast<int32_t *>(h_src)[i]);
            if (reinterpret_cast<float *>(h_dst)[i] != want) {
                fprintf(stderr, "MISMATCH at word %llu of %llu frames\n", (unsigned long long)i, (unsigned long long)n);
                exit(2);
            }
        }
    };

    for (int lg : {20, 22, 24}) {
        const uint64_t n = uint64_t(1) << lg;
        const int reps = lg <= 20 ? 200 : lg <= 22 ? 60 : 20;
        // ---- the library's scheme: K chunks, copy -> event -> kernel
        for (int K : {4, 8}) {
            const uint64_t cf = n / K;
            auto call = [&] {
                for (int c = 0; c < K; c++) {
                    CK(cudaMemcpyAsync(d_stage + c * cf * 8, h_src + c * cf * 8, cf * 8, cudaMemcpyHostToDevice, s_copy));
                    CK(cudaEventRecord(ev[c], s_copy));
                    CK(cudaStreamWaitEvent(s_comp, ev[c], 0));
                    chunk_kernel<RxCf32><<<sms, 256, 0, s_comp>>>(d_stage + c * cf * 8, h_dst + c * cf * 8, cf, 0.0f);
                }
                CK(cudaStreamSynchronize(s_comp));
            };
            for (int i = 0; i < 3; i++)
                call();
            memset(h_dst, 0, 64);
            const double t0 = now_us();
            for (int i = 0; i < reps; i++)
                call();
            const double us = (now_us() - t0) / reps;
            check(n);
            printf("{\"frames\": %llu, \"scheme\": \"pipeline_%d_chunks\", \"us\": %.1f, \"gbs_each_way\": %.2f}\n", (unsigned long long)n, K,
                   us, 8.0 * n / us / 1e3);
        }
        // ---- one resident kernel, chunks announced by stream memory operations (or 4-byte copies)
        for (int flag_kind = 0; flag_kind < 0; flag_kind++) {
            if (flag_kind == 0 && !write_value)
                continue;
            for (uint64_t cf : {uint64_t(1) << 16, uint64_t(1) << 18}) {
                if (cf * 2 > n)
                    continue;
                const uint32_t nchunks = uint32_t((n + cf - 1) / cf);
                struct V { int grid; uint32_t sleep; bool dev; };
                for (V v : {V{sms, 0, false}, V{sms, 500, false}, V{32, 0, false}, V{sms, 0, true}, V{32, 500, true}}) {
                    const int grid_ctas = v.grid, grid_mul = v.grid;
                    const uint32_t sleep_ns = v.sleep;
                    const bool to_device = v.dev;
                    auto call = [&] {
                        StreamedArgs a = {d_stage, to_device ? d_out : h_dst, n, cf, nchunks, seq, d_flag, d_err, 0.0f, sleep_ns, nullptr, nullptr};
                        streamed_kernel<RxCf32><<<grid_ctas, 256, 0, s_comp>>>(a);
                        for (uint32_t c = 0; c < nchunks; c++) {
                            const uint64_t lo = c * cf, len = (lo + cf < n ? cf : n - lo);
                            CK(cudaMemcpyAsync(d_stage + lo * 8, h_src + lo * 8, len * 8, cudaMemcpyHostToDevice, s_copy));
                            if (flag_kind == 0) {
                                if (write_value(s_copy, (unsigned long long)d_flag, seq + c + 1, 0) != 0) {
                                    fprintf(stderr, "cuStreamWriteValue32 failed\n");
                                    exit(3);
                                }
                            } else {
                                h_seq[(seq + c) & 65535] = seq + c + 1;
                                CK(cudaMemcpyAsync(d_flag, &h_seq[(seq + c) & 65535], 4, cudaMemcpyHostToDevice, s_copy));
                            }
                        }
                        CK(cudaStreamSynchronize(s_comp));
                        seq += nchunks;
                    };
                    for (int i = 0; i < 3; i++)
                        call();
                    memset(h_dst, 0, 64);
                    const double t0 = now_us();
                    for (int i = 0; i < reps; i++)
                        call();
                    const double us = (now_us() - t0) / reps;
                    if (!to_device)
                        check(n);
                    uint32_t err = 0;
                    CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
                    printf("{\"frames\": %llu, \"scheme\": \"streamed_%s_chunk%llu_grid%d_sleep%u%s\", \"us\": %.1f, \"gbs_each_way\": %.2f, \"error\": %u}\n",
                           (unsigned long long)n, flag_kind == 0 ? "write_value" : "memcpy_flag", (unsigned long long)cf, grid_mul, sleep_ns,
                           to_device ? "_to_device" : "", us,
                           8.0 * n / us / 1e3, err);
                    fflush(stdout);
                }
            }
        }
        // ---- resident kernel between two copy-engine streams: chunks announced to the kernel by
        // cuStreamWriteValue32, finished chunks announced to the device-to-host stream by a counter the
        // kernel raises and cuStreamWaitValue32 waits for.  Everything is queued up front.
        if (write_value && wait_value) {
            for (uint64_t cf : {uint64_t(1) << 16, uint64_t(1) << 17, uint64_t(1) << 18, uint64_t(1) << 19}) {
                if (cf * 2 > n)
                    continue;
                const uint32_t nchunks = uint32_t((n + cf - 1) / cf);
                for (int grid_ctas : {sms, 32}) {
                    auto call = [&] {
                        StreamedArgs a = {d_stage, d_out, n, cf, nchunks, seq, d_flag, d_err, 0.0f, 0, d_arrivals, d_done};
                        streamed_kernel<RxCf32><<<grid_ctas, 256, 0, s_comp>>>(a);
                        for (uint32_t c = 0; c < nchunks; c++) {
                            const uint64_t lo = c * cf, len = (lo + cf < n ? cf : n - lo);
                            CK(cudaMemcpyAsync(d_stage + lo * 8, h_src + lo * 8, len * 8, cudaMemcpyHostToDevice, s_copy));
                            if (write_value(s_copy, (unsigned long long)d_flag, seq + c + 1, 0) != 0)
                                exit(3);
                            if (wait_value(s_d2h, (unsigned long long)d_done, seq + c + 1, 0 /* CU_STREAM_WAIT_VALUE_GEQ */) != 0)
                                exit(4);
                            CK(cudaMemcpyAsync(h_dst + lo * 8, d_out + lo * 8, len * 8, cudaMemcpyDeviceToHost, s_d2h));
                        }
                        CK(cudaStreamSynchronize(s_d2h));
                        CK(cudaStreamSynchronize(s_comp));
                        seq += nchunks;
                    };
                    for (int i = 0; i < 3; i++)
                        call();
                    memset(h_dst, 0, 64);
                    const double t0 = now_us();
                    for (int i = 0; i < reps; i++)
                        call();
                    const double us = (now_us() - t0) / reps;
                    check(n);
                    printf("{\"frames\": %llu, \"scheme\": \"streamed_both_copy_engines_chunk%llu_grid%d\", \"us\": %.1f, \"gbs_each_way\": %.2f}\n",
                           (unsigned long long)n, (unsigned long long)cf, grid_ctas, us, 8.0 * n / us / 1e3);
                    fflush(stdout);
                }
            }
        }
        // ---- three-stage pipeline with events (the library's copy-engine output mode)
        for (int K : {4, 8, 16}) {
            const uint64_t cf = n / K;
            auto call = [&] {
                for (int c = 0; c < K; c++) {
                    CK(cudaMemcpyAsync(d_stage + c * cf * 8, h_src + c * cf * 8, cf * 8, cudaMemcpyHostToDevice, s_copy));
                    CK(cudaEventRecord(ev[c], s_copy));
                    CK(cudaStreamWaitEvent(s_comp, ev[c], 0));
                    chunk_kernel<RxCf32><<<sms, 256, 0, s_comp>>>(d_stage + c * cf * 8, d_out + c * cf * 8, cf, 0.0f);
                    CK(cudaEventRecord(ev[32 + c], s_comp));
                    CK(cudaStreamWaitEvent(s_d2h, ev[32 + c], 0));
                    CK(cudaMemcpyAsync(h_dst + c * cf * 8, d_out + c * cf * 8, cf * 8, cudaMemcpyDeviceToHost, s_d2h));
                }
                CK(cudaStreamSynchronize(s_d2h));
            };
            for (int i = 0; i < 3; i++)
                call();
            memset(h_dst, 0, 64);
            const double t0 = now_us();
            for (int i = 0; i < reps; i++)
                call();
            const double us = (now_us() - t0) / reps;
            check(n);
            printf("{\"frames\": %llu, \"scheme\": \"pipeline_3_stages_%d_chunks\", \"us\": %.1f, \"gbs_each_way\": %.2f}\n", (unsigned long long)n, K,
                   us, 8.0 * n / us / 1e3);
        }
        // ---- both copy engines, no kernel at all: K chunks up, K chunks down, chunk c down after chunk c up
        for (int K : {1, 4, 16}) {
            const uint64_t cf = n / K;
            auto call = [&] {
                for (int c = 0; c < K; c++) {
                    CK(cudaMemcpyAsync(d_stage + c * cf * 8, h_src + c * cf * 8, cf * 8, cudaMemcpyHostToDevice, s_copy));
                    CK(cudaEventRecord(ev[c], s_copy));
                    CK(cudaStreamWaitEvent(s_d2h, ev[c], 0));
                    CK(cudaMemcpyAsync(h_dst + c * cf * 8, d_stage + c * cf * 8, cf * 8, cudaMemcpyDeviceToHost, s_d2h));
                }
                CK(cudaStreamSynchronize(s_d2h));
            };
            for (int i = 0; i < 3; i++)
                call();
            const double t0 = now_us();
            for (int i = 0; i < reps; i++)
                call();
            const double us = (now_us() - t0) / reps;
            printf("{\"frames\": %llu, \"scheme\": \"copies_only_up_then_down_%d_chunks\", \"us\": %.1f, \"gbs_each_way\": %.2f}\n", (unsigned long long)n, K,
                   us, 8.0 * n / us / 1e3);
        }
        // ---- the floor: the host-to-device copy alone, and a copy each way at once
        {
            const double t0 = now_us();
            for (int i = 0; i < reps; i++) {
                CK(cudaMemcpyAsync(d_stage, h_src, n * 8, cudaMemcpyHostToDevice, s_copy));
                CK(cudaStreamSynchronize(s_copy));
            }
            const double us = (now_us() - t0) / reps;
            printf("{\"frames\": %llu, \"scheme\": \"h2d_copy_alone\", \"us\": %.1f, \"gbs_each_way\": %.2f}\n", (unsigned long long)n, us,
                   8.0 * n / us / 1e3);
        }
    }
    return 0;
}

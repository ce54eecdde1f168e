// pcie_staging.cu -- does write-combined pinned memory move faster over the link than ordinary
// pinned memory, one way and with both directions loaded?
//
// Why: the host-buffer leg of the sample path is PCIe-bound (46-47 GB/s each way with both
// directions loaded, 55-57 one way).  The RX staging buffer is written only by the CPU (the
// I2S frames arriving) and read only by the GPU, which is the textbook case for
// cudaHostAllocWriteCombined.  Copy engines only, no kernels.
//     nvcc -O3 -std=c++17 -o tools/experiments/bin/pcie_staging tools/experiments/pcie_staging.cu
// Prints one JSON object.  NOT a product path and not part of any judged number.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                      \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)

int main()
{
    const size_t bytes = size_t(512) << 20, chunk = size_t(16) << 20;
    char *h_plain, *h_wc, *h_out, *d_in, *d_out;
    CK(cudaHostAlloc(reinterpret_cast<void **>(&h_plain), bytes, cudaHostAllocDefault));
    CK(cudaHostAlloc(reinterpret_cast<void **>(&h_wc), bytes, cudaHostAllocWriteCombined));
    CK(cudaHostAlloc(reinterpret_cast<void **>(&h_out), bytes, cudaHostAllocDefault));
    CK(cudaMalloc(reinterpret_cast<void **>(&d_in), bytes));
    CK(cudaMalloc(reinterpret_cast<void **>(&d_out), bytes));
    std::memset(h_plain, 1, bytes);
    std::memset(h_wc, 1, bytes); // CPU writes to write-combined memory are fine; reads are what is slow
    CK(cudaMemset(d_out, 2, bytes));
    cudaStream_t up, down;
    CK(cudaStreamCreate(&up));
    CK(cudaStreamCreate(&down));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));

    auto run = [&](const char *name, const char *h_src, bool both, bool last) {
        std::vector<float> ms;
        for (int rep = 0; rep < 8; rep++) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a, up));
            CK(cudaStreamWaitEvent(down, a, 0));
            for (size_t off = 0; off < bytes; off += chunk) { // chunked like the library's pipeline
                CK(cudaMemcpyAsync(d_in + off, h_src + off, chunk, cudaMemcpyHostToDevice, up));
                if (both)
                    CK(cudaMemcpyAsync(h_out + off, d_out + off, chunk, cudaMemcpyDeviceToHost, down));
            }
            CK(cudaEventRecord(b, down));
            CK(cudaStreamWaitEvent(up, b, 0));
            CK(cudaEventRecord(b, up));
            CK(cudaEventSynchronize(b));
            float t;
            CK(cudaEventElapsedTime(&t, a, b));
            if (rep >= 2)
                ms.push_back(t);
        }
        std::sort(ms.begin(), ms.end());
        std::printf("  \"%s\": {\"best_gbs_each_way\": %.1f, \"median_gbs_each_way\": %.1f}%s\n", name,
                    bytes / ms.front() / 1e6, bytes / ms[ms.size() / 2] / 1e6, last ? "" : ",");
    };
    std::printf("{\n  \"bytes_each_way\": %zu, \"chunk_bytes\": %zu,\n", bytes, chunk);
    run("h2d_only_plain_pinned", h_plain, false, false);
    run("h2d_only_write_combined", h_wc, false, false);
    run("h2d_plus_d2h_plain_pinned", h_plain, true, false);
    run("h2d_plus_d2h_write_combined_source", h_wc, true, true);
    std::printf("}\n");
    return 0;
}

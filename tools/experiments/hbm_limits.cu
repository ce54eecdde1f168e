// hbm_limits.cu -- what the HBM of this part gives a read-only, a write-only and a copy stream
// through the same bulk-async (TMA) machinery the converters use, next to cudaMemcpy.
//
// Why: the converters run at the device-copy figure (6.5-6.7 TB/s, about 80 % of the
// specified 8 TB/s).  Before looking for more, split that figure: how fast can the part read
// alone and write alone, and does a copy that alternates long read-only and write-only phases
// (fewer DRAM bus turnarounds) beat one that interleaves them tile by tile?
//
// Standalone: uses only the PTX helpers of sx_kernels.cuh, none of the library.
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sxxcvr_b200/csrc \
//          -o gpurun_out/hbm_limits tools/experiments/hbm_limits.cu && gpurun_out/hbm_limits
// Prints one JSON object.  NOT a product path and not yet part of any judged number.
#include "sx_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace sx;

constexpr uint32_t kTile = 16384; // bytes per bulk copy, as the converters' 2048-frame tiles

struct Args {
    const char *src;
    char *dst;
    uint64_t ntiles;
};

// Loads only: a ring of STAGES tiles kept in flight by thread 0, nothing consumes them.
template <int STAGES> __global__ void read_only_kernel(Args a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + size_t(STAGES) * kTile);
    if (threadIdx.x != 0)
        return;
    for (int s = 0; s < STAGES; s++)
        bulk::mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint64_t pol = bulk::make_policy(0);
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint64_t mine = first >= a.ntiles ? 0 : (a.ntiles - first + stride - 1) / stride;
    for (uint64_t i = 0; i < uint64_t(STAGES) && i < mine; i++) {
        bulk::mbar_expect_tx(&full[i], kTile);
        bulk::load_g2s(smem + i * kTile, a.src + (first + i * stride) * kTile, kTile, &full[i], pol);
    }
    for (uint64_t i = 0; i < mine; i++) {
        const int s = int(i % STAGES);
        bulk::mbar_wait(&full[s], uint32_t(i / STAGES) & 1u);
        const uint64_t nxt = i + STAGES;
        if (nxt < mine) {
            bulk::mbar_expect_tx(&full[s], kTile);
            bulk::load_g2s(smem + size_t(s) * kTile, a.src + (first + nxt * stride) * kTile, kTile, &full[s], pol);
        }
    }
}

// Stores only: the same tile of shared memory is stored over and over, at most DEPTH bulk
// groups outstanding.
template <int DEPTH> __global__ void write_only_kernel(Args a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    for (uint32_t i = threadIdx.x; i < kTile / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(smem)[i] = i;
    bulk::fence_async_smem();
    __syncthreads();
    if (threadIdx.x != 0)
        return;
    const uint64_t pol = bulk::make_policy(0);
    for (uint64_t t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        bulk::store_s2g(a.dst + t * kTile, smem, kTile, pol);
        bulk::commit_group();
        bulk::wait_group_read<DEPTH - 1>();
    }
    bulk::wait_group_all();
}

// Copy, interleaved: tile i is stored from the buffer it was loaded into as soon as it has
// arrived; the buffer of tile i-1 is refilled once its store has read it.
template <int STAGES> __global__ void copy_interleaved_kernel(Args a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + size_t(STAGES) * kTile);
    if (threadIdx.x != 0)
        return;
    for (int s = 0; s < STAGES; s++)
        bulk::mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint64_t pol = bulk::make_policy(0);
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint64_t mine = first >= a.ntiles ? 0 : (a.ntiles - first + stride - 1) / stride;
    for (uint64_t i = 0; i < uint64_t(STAGES) && i < mine; i++) {
        bulk::mbar_expect_tx(&full[i], kTile);
        bulk::load_g2s(smem + i * kTile, a.src + (first + i * stride) * kTile, kTile, &full[i], pol);
    }
    for (uint64_t i = 0; i < mine; i++) {
        const int s = int(i % STAGES);
        bulk::mbar_wait(&full[s], uint32_t(i / STAGES) & 1u);
        bulk::store_s2g(a.dst + (first + i * stride) * kTile, smem + size_t(s) * kTile, kTile, pol);
        bulk::commit_group();
        if (i >= 1) { // the store of tile i-1 has read its buffer once at most one group is pending
            bulk::wait_group_read<1>();
            const uint64_t nxt = i - 1 + STAGES;
            if (nxt < mine) {
                const int ps = int((i - 1) % STAGES);
                bulk::mbar_expect_tx(&full[ps], kTile);
                bulk::load_g2s(smem + size_t(ps) * kTile, a.src + (first + nxt * stride) * kTile, kTile, &full[ps], pol);
            }
        }
    }
    bulk::wait_group_all();
}

// Copy, phased: load PHASE tiles, wait for all of them, store all of them, wait until the
// stores have read shared memory, repeat.  Each CTA alternates between pure reading and pure
// writing; with SYNC the whole grid does so in step (cooperative launch not needed: the phases
// are aligned only by starting together and doing equal work).
template <int PHASE> __global__ void copy_phased_kernel(Args a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + size_t(PHASE) * kTile);
    if (threadIdx.x != 0)
        return;
    bulk::mbar_init(&full[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint64_t pol = bulk::make_policy(0);
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint64_t mine = first >= a.ntiles ? 0 : (a.ntiles - first + stride - 1) / stride;
    uint32_t parity = 0;
    for (uint64_t base = 0; base < mine; base += PHASE) {
        const uint32_t n = uint32_t(mine - base < uint64_t(PHASE) ? mine - base : uint64_t(PHASE));
        bulk::mbar_expect_tx(&full[0], n * kTile);
        for (uint32_t j = 0; j < n; j++)
            bulk::load_g2s(smem + size_t(j) * kTile, a.src + (first + (base + j) * stride) * kTile, kTile, &full[0], pol);
        bulk::mbar_wait(&full[0], parity);
        parity ^= 1u;
        for (uint32_t j = 0; j < n; j++)
            bulk::store_s2g(a.dst + (first + (base + j) * stride) * kTile, smem + size_t(j) * kTile, kTile, pol);
        bulk::commit_group();
        bulk::wait_group_read<0>();
    }
    bulk::wait_group_all();
}

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                      \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)

template <class F> static void timed(const char *name, double bytes, F launch, bool last = false)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; i++)
        launch();
    CK(cudaDeviceSynchronize());
    std::vector<float> ms;
    for (int i = 0; i < 20; i++) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float t;
        CK(cudaEventElapsedTime(&t, a, b));
        ms.push_back(t);
    }
    CK(cudaGetLastError());
    std::sort(ms.begin(), ms.end());
    std::printf("  \"%s\": {\"best_gbs\": %.1f, \"median_gbs\": %.1f}%s\n", name, bytes / ms.front() / 1e6,
                bytes / ms[ms.size() / 2] / 1e6, last ? "" : ",");
}

template <class K> static void allow_smem(K kernel, size_t bytes)
{
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const uint64_t bytes = uint64_t(1) << 30; // 1 GiB each way, far beyond the 126 MB L2
    char *src, *dst;
    CK(cudaMalloc(&src, bytes));
    CK(cudaMalloc(&dst, bytes));
    CK(cudaMemset(src, 1, bytes));
    CK(cudaMemset(dst, 2, bytes));
    Args a = {src, dst, bytes / kTile};
    const int grid = prop.multiProcessorCount;
    std::printf("{\n  \"device\": \"%s\", \"sms\": %d, \"bytes_each_way\": %llu, \"tile_bytes\": %u,\n", prop.name, grid,
                (unsigned long long)bytes, kTile);

    timed("cudaMemcpy_d2d_read_plus_write", 2.0 * bytes, [&] { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice)); });
    timed("cudaMemset_write_only", 1.0 * bytes, [&] { CK(cudaMemsetAsync(dst, 3, bytes)); });

    allow_smem(read_only_kernel<4>, 4 * kTile + 64);
    allow_smem(read_only_kernel<8>, 8 * kTile + 64);
    allow_smem(read_only_kernel<12>, 12 * kTile + 128);
    timed("bulk_read_only_4_stages", 1.0 * bytes, [&] { read_only_kernel<4><<<grid, 32, 4 * kTile + 64>>>(a); });
    timed("bulk_read_only_8_stages", 1.0 * bytes, [&] { read_only_kernel<8><<<grid, 32, 8 * kTile + 64>>>(a); });
    timed("bulk_read_only_12_stages", 1.0 * bytes, [&] { read_only_kernel<12><<<grid, 32, 12 * kTile + 128>>>(a); });

    timed("bulk_write_only_depth_2", 1.0 * bytes, [&] { write_only_kernel<2><<<grid, 128, kTile>>>(a); });
    timed("bulk_write_only_depth_4", 1.0 * bytes, [&] { write_only_kernel<4><<<grid, 128, kTile>>>(a); });
    timed("bulk_write_only_depth_8", 1.0 * bytes, [&] { write_only_kernel<8><<<grid, 128, kTile>>>(a); });

    allow_smem(copy_interleaved_kernel<4>, 4 * kTile + 64);
    allow_smem(copy_interleaved_kernel<8>, 8 * kTile + 64);
    timed("bulk_copy_interleaved_4_stages", 2.0 * bytes, [&] { copy_interleaved_kernel<4><<<grid, 32, 4 * kTile + 64>>>(a); });
    timed("bulk_copy_interleaved_8_stages", 2.0 * bytes, [&] { copy_interleaved_kernel<8><<<grid, 32, 8 * kTile + 64>>>(a); });

    allow_smem(copy_phased_kernel<4>, 4 * kTile + 64);
    allow_smem(copy_phased_kernel<8>, 8 * kTile + 64);
    allow_smem(copy_phased_kernel<12>, 12 * kTile + 64);
    timed("bulk_copy_phased_4_tiles", 2.0 * bytes, [&] { copy_phased_kernel<4><<<grid, 32, 4 * kTile + 64>>>(a); });
    timed("bulk_copy_phased_8_tiles", 2.0 * bytes, [&] { copy_phased_kernel<8><<<grid, 32, 8 * kTile + 64>>>(a); });
    timed("bulk_copy_phased_12_tiles", 2.0 * bytes, [&] { copy_phased_kernel<12><<<grid, 32, 12 * kTile + 64>>>(a); }, true);
    std::printf("}\n");

    // the copies must have copied
    std::vector<char> probe(4096);
    CK(cudaMemcpy(probe.data(), dst + bytes - 4096, 4096, cudaMemcpyDeviceToHost));
    for (char c : probe)
        if (c != 1) {
            std::fprintf(stderr, "copy check failed\n");
            return 1;
        }
    return 0;
}

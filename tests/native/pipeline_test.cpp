// Exercises the host pipeline pieces of csrc/host/par_copy.hpp on the CPU, without CUDA:
//  * plan_chunks: every schedule covers [0, length) exactly once, in order, on 64-frame
//    boundaries, with no chunk above the cap and the ramp mirrored at both ends;
//  * Sidekick + Progress: the two-thread producer/retirer handshake of convert_host, with a
//    4-slot ring, where the "device" is a memcpy -- every byte must arrive, and an aborted run
//    must release both threads.
// Built and run by tests/test_host_logic.py, under ThreadSanitizer when available.
#include "host/par_copy.hpp"
#include "host/worker_pool.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace sxhost;

static int check_plan(size_t length, size_t c_min, size_t c_max)
{
    const std::vector<ChunkSpan> plan = plan_chunks(length, c_min, c_max);
    const size_t cap = ((c_max < 64 ? 64 : c_max) + 63) & ~size_t(63);
    size_t at = 0;
    for (size_t i = 0; i < plan.size(); i++) {
        if (plan[i].first != at || plan[i].frames == 0 || plan[i].frames > cap + (i + 1 == plan.size() ? 63 : 0)) {
            std::fprintf(stderr, "plan(%zu,%zu,%zu): chunk %zu = [%zu,+%zu) at %zu\n", length, c_min, c_max, i,
                         plan[i].first, plan[i].frames, at);
            return 1;
        }
        if (plan[i].first % 64 != 0) {
            std::fprintf(stderr, "plan(%zu,%zu,%zu): boundary %zu not on 64 frames\n", length, c_min, c_max, plan[i].first);
            return 1;
        }
        at += plan[i].frames;
    }
    if (at != length || (length == 0) != plan.empty()) {
        std::fprintf(stderr, "plan(%zu,%zu,%zu) covers %zu\n", length, c_min, c_max, at);
        return 1;
    }
    return 0;
}

int main(int argc, char **argv)
{
    const unsigned helpers = argc > 1 ? unsigned(std::atoi(argv[1])) : 2;
    int plans = 0;
    const size_t lengths[] = {0, 1, 63, 64, 65, 1000, 65536, 65537, (1u << 18) + 1, (1u << 20) + 17, (1u << 23) - 5,
                              size_t(1) << 27};
    const size_t mins[] = {0, 1, 64, 1 << 12, 1 << 16, 1 << 22};
    const size_t maxs[] = {1, 64, 1024, (1 << 19) + 5, 1 << 22};
    for (size_t l : lengths)
        for (size_t a : mins)
            for (size_t b : maxs) {
                if (check_plan(l, a, b))
                    return 1;
                plans++;
            }

    { // the ramp: small first and last chunks, doubling towards the middle, mirrored
        const std::vector<ChunkSpan> p = plan_chunks((size_t(1) << 20) + 17, 1 << 12, 1 << 17);
        if (p.size() < 9 || p.front().frames != 4096 || p.back().frames != 4096 + 17 || p[1].frames != 8192 ||
            p[p.size() - 2].frames != 8192 || p[2].frames != 16384) {
            std::fprintf(stderr, "ramp: unexpected schedule (%zu chunks, first %zu, last %zu)\n", p.size(),
                         p.front().frames, p.back().frames);
            return 1;
        }
        const std::vector<ChunkSpan> flat = plan_chunks(size_t(1) << 20, 0, 1 << 17);
        if (flat.size() != 8 || flat.front().frames != (1 << 17)) {
            std::fprintf(stderr, "uniform schedule: %zu chunks\n", flat.size());
            return 1;
        }
    }

    // The handshake: `issued` counts chunks the owner has queued, `retired` chunks the sidekick
    // has copied out; slot i % K may be refilled once chunk i - K has retired.
    constexpr int K = 4;
    Sidekick sidekick;
    Progress issued, retired;
    std::atomic<int> abort_flag{0};
    ParallelCopier copier_in(helpers), copier_out(helpers);
    int runs = 0;
    for (size_t length : {size_t(1) << 20, (size_t(3) << 20) + 12345, size_t(9) << 20}) {
        for (int fail_at : {-1, 5}) {
            const std::vector<ChunkSpan> chunks = plan_chunks(length, 1 << 12, 1 << 17);
            std::vector<unsigned char> src(length), dst(length, 0), slots[K];
            for (auto &s : slots)
                s.resize((size_t(1) << 17) + 64);
            for (size_t i = 0; i < length; i++)
                src[i] = (unsigned char)(i * 2654435761u >> 24);
            issued.reset();
            retired.reset();
            abort_flag.store(0);
            sidekick.start([&] {
                for (size_t i = 0; i < chunks.size(); i++) {
                    if (!issued.wait_for(i + 1, abort_flag))
                        return;
                    copier_out.copy(dst.data() + chunks[i].first, slots[i % K].data(), chunks[i].frames);
                    retired.publish(i + 1);
                }
            });
            bool failed = false;
            for (size_t i = 0; i < chunks.size() && !failed; i++) {
                if (i >= K && !retired.wait_for(i - K + 1, abort_flag))
                    break;
                if (int(i) == fail_at) { // an error on the owner's side: tell the sidekick and stop
                    abort_flag.store(1);
                    failed = true;
                    break;
                }
                copier_in.copy(slots[i % K].data(), src.data() + chunks[i].first, chunks[i].frames);
                issued.publish(i + 1);
            }
            sidekick.finish();
            if (!failed && dst != src) {
                std::fprintf(stderr, "pipeline: %zu bytes did not arrive intact\n", length);
                return 1;
            }
            runs++;
        }
    }
    // WorkerPool: disjoint ranges covering [0, n), every item exactly once, many rounds, ranges
    // below the grain run on the owner alone.
    {
        WorkerPool pool(helpers);
        for (size_t n : {size_t(0), size_t(1), size_t(31), size_t(32), size_t(33), size_t(1000), size_t(4096)}) {
            for (int round = 0; round < 20; round++) {
                std::vector<int> hits(n, 0);
                pool.run(n, 32, [&](size_t lo, size_t hi) {
                    for (size_t i = lo; i < hi; i++)
                        hits[i]++;
                });
                for (size_t i = 0; i < n; i++)
                    if (hits[i] != 1) {
                        std::fprintf(stderr, "pool: item %zu of %zu ran %d times\n", i, n, hits[i]);
                        return 1;
                    }
            }
        }
    }
    std::printf("pipeline ok: %d plans, %d runs, %u helpers\n", plans, runs, helpers);
    return 0;
}

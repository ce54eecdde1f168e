// Exercises sxhost::ParallelCopier (csrc/host/par_copy.hpp) on the CPU: every size class, odd
// sizes and offsets, guard bytes on both sides of the destination, many copies through one pool.
// Built and run by tests/test_host_logic.py, under ThreadSanitizer when available.
#include "host/par_copy.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

static uint64_t state = 0x53581255ull;
static uint64_t next()
{
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    return state;
}

int main(int argc, char **argv)
{
    const unsigned helpers = argc > 1 ? unsigned(std::atoi(argv[1])) : 3;
    const int rounds = argc > 2 ? std::atoi(argv[2]) : 3;
    const size_t M = sxhost::ParallelCopier::kMinParallelBytes;
    const size_t sizes[] = {0, 1, 4095, M - 1, M, M + 1, 3 * M + 12345, 4096 * (helpers + 1) * 700 + helpers,
                            8 * M + 7};
    size_t largest = 0;
    for (size_t n : sizes)
        largest = n > largest ? n : largest;
    std::vector<unsigned char> src(largest + 64), dst(largest + 128);
    for (auto &b : src)
        b = (unsigned char)next();

    sxhost::ParallelCopier copier(helpers);
    int checked = 0;
    for (int round = 0; round < rounds; round++) {
        for (size_t n : sizes) {
            const size_t so = next() % 64, d_o = 32 + next() % 64;
            for (auto &b : dst)
                b = 0xA5;
            copier.copy(dst.data() + d_o, src.data() + so, n);
            for (size_t i = 0; i < dst.size(); i++) {
                const unsigned char want = (i >= d_o && i < d_o + n) ? src[so + (i - d_o)] : 0xA5;
                if (dst[i] != want) {
                    std::fprintf(stderr, "mismatch: size %zu offset %zu byte %zu\n", n, d_o, i);
                    return 1;
                }
            }
            checked++;
        }
    }
    std::printf("par_copy ok: %d copies, %u helpers\n", checked, copier.helpers());
    return 0;
}

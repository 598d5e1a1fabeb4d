// Host-only check of kinematic-icp_b200/csrc/kicp_stager.hpp (test infrastructure): many staging jobs of random sizes, granules and
// helper counts; after wait_prefix(upto) the first `upto` bytes must already be in place, after the last wait everything.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "kicp_stager.hpp"

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 300;
    std::mt19937_64 rng(12345);
    long long bytes_total = 0;
    for (int helpers : {0, 1, 3, 7}) {
        kicp::Stager st(helpers);
        for (int r = 0; r < rounds; ++r) {
            const size_t bytes = (size_t)(rng() % (r % 10 == 0 ? (8u << 20) : (300u << 10))) + (r % 7 == 0 ? 0 : 1);
            const size_t gran = (size_t)1 << (10 + rng() % 8);  // 1 KB .. 128 KB
            std::vector<unsigned char> src(bytes + 64), dst(bytes + 64, 0xEE);
            for (auto &b : src) b = (unsigned char)rng();
            auto job = st.start(src.data(), dst.data(), bytes, gran);
            size_t upto = 0;
            while (true) {
                upto = std::min(bytes, upto + (size_t)(rng() % (1u << 20)) + 1);
                st.wait_prefix(*job, upto);
                if (memcmp(src.data(), dst.data(), upto) != 0) {
                    printf("FAIL prefix: helpers %d round %d bytes %zu gran %zu upto %zu\n", helpers, r, bytes, gran, upto);
                    return 1;
                }
                if (upto >= bytes) break;
            }
            for (size_t k = bytes; k < bytes + 64; ++k)
                if (dst[k] != 0xEE) {
                    printf("FAIL overrun: helpers %d round %d\n", helpers, r);
                    return 1;
                }
            bytes_total += (long long)bytes;
        }
    }
    printf("OK %lld bytes staged\n", bytes_total);
    return 0;
}

// copy_pool_stress.cpp -- TEST ONLY: hammers CopyPool (goofy_b200/csrc/copy_pool.h) with jobs of varying size and
// pauses long enough for the workers to fall asleep, checking every copy.  Built and run by tests/test_host_logic.py
// (plain and, when the toolchain has it, under ThreadSanitizer).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../goofy_b200/csrc/copy_pool.h"

int main(int argc, char** argv)
{
    const int jobs = argc > 1 ? atoi(argv[1]) : 2000;
    const size_t N = 8u << 20;
    std::vector<uint8_t> src(N), dst(N);
    for (size_t i = 0; i < N; ++i) src[i] = (uint8_t)(i * 131u + 7u);
    unsigned long long checks = 0;
    for (int it = 0; it < jobs; ++it) {
        const size_t rows = 64 + (size_t)(it * 37) % 960, rowBytes = 4096, pitch = 8192;
        std::memset(dst.data(), 0, rows * rowBytes);
        CopyPool::get().copy2d(dst.data(), rowBytes, src.data(), pitch, rowBytes, rows);
        for (size_t r = 0; r < rows; r += 17) {
            if (std::memcmp(dst.data() + r * rowBytes, src.data() + r * pitch, rowBytes)) {
                std::printf("MISMATCH job %d row %zu\n", it, r);
                return 1;
            }
            ++checks;
        }
        if (it % 3 == 0) {   // the alpha-stripping job (pack2d) through the same handshake: 1024 pixels per row
            const size_t pixels = rowBytes / 4u;
            std::memset(dst.data(), 0, rows * pixels * 3u);
            CopyPool::get().pack2d(dst.data(), pixels * 3u, src.data(), pitch, pixels, rows, it % 2 == 0);
            for (size_t r = 0; r < rows; r += 13) {
                const uint8_t* d = dst.data() + r * pixels * 3u;
                const uint8_t* q = src.data() + r * pitch;
                for (size_t x = 0; x < pixels; ++x)
                    if (d[3 * x] != q[4 * x] || d[3 * x + 1] != q[4 * x + 1] || d[3 * x + 2] != q[4 * x + 2]) {
                        std::printf("PACK MISMATCH job %d row %zu pixel %zu\n", it, r, x);
                        return 1;
                    }
                ++checks;
            }
        }
        if (it % 5 == 0) std::this_thread::sleep_for(std::chrono::microseconds(it % 300));  // lets the workers fall asleep
        if (it % 7 == 0) CopyPool::get().copy1d(dst.data(), src.data(), 300000 + (size_t)it);
    }
    std::printf("ok %d jobs %llu row checks\n", jobs, checks);
    return 0;
}

// rgb_pack_host.cpp -- TEST ONLY: the alpha-stripping row packers of goofy_b200/csrc/rgb_pack.h (scalar and SSSE3)
// against a byte loop, on widths that are multiples of 4 (the float-reference flavour's contract) and of 16, at every
// source / destination misalignment the staging code can produce.  Built and run by tests/test_host_logic.py.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../goofy_b200/csrc/rgb_pack.h"

int main()
{
    std::vector<uint8_t> src(4 * 4096 + 64), want(3 * 4096 + 64), got(3 * 4096 + 64);
    unsigned seed = 12345u;
    for (auto& b : src) { seed = seed * 1664525u + 1013904223u; b = (uint8_t)(seed >> 24); }
    unsigned long long checks = 0;
    for (size_t pixels = 4; pixels <= 4096; pixels += (pixels < 128 ? 4 : 332)) {
        for (size_t so = 0; so < 32; so += 4) {
            for (size_t d0 = 0; d0 < 16; d0 += 4) {
                const uint8_t* s = src.data() + so;
                for (size_t x = 0; x < pixels; ++x)
                    for (int c = 0; c < 3; ++c) want[d0 + 3 * x + c] = s[4 * x + c];
                for (int variant = 0; variant < 4; ++variant) {
                    std::fill(got.begin(), got.end(), (uint8_t)0xEE);
                    if (variant == 0) gbpack::pack_row_scalar(got.data() + d0, s, pixels);
                    else if (variant == 1) gbpack::pack_row(got.data() + d0, s, pixels, false);
                    else if (variant == 2) gbpack::pack_row(got.data() + d0, s, pixels, true);
                    else gbpack::pack_rows(got.data() + d0, 0, s, 0, pixels, 0, 1, true);
                    for (size_t i = 0; i < got.size(); ++i) {
                        const bool inside = i >= d0 && i < d0 + 3 * pixels;
                        if (got[i] != (inside ? want[i] : (uint8_t)0xEE)) {
                            std::printf("MISMATCH variant %d pixels %zu src+%zu dst+%zu byte %zu\n", variant, pixels, so, d0, i);
                            return 1;
                        }
                    }
                    ++checks;
                }
            }
        }
    }
#ifdef GB_PACK_X86
    std::printf("ok %llu rows, ssse3 %d\n", checks, (int)gbpack::have_ssse3());
#else
    std::printf("ok %llu rows, portable\n", checks);
#endif
    return 0;
}

// hybrid_choice_host.cpp -- TEST ONLY: HybridChoice (goofy_b200/csrc/hybrid_choice.h) driven through the three situations
// measured on the GPU boxes (profiles/r02_rgb24_sessions.md).  Built and run by tests/test_host_logic.py.
#include <cstdio>

#include "../goofy_b200/csrc/hybrid_choice.h"

struct Host {
    double plain, packing;   // input bytes per second a call achieves either way
};

// runs `calls` calls, returns how many packed; `firstCold`: the very first call is three times slower
static int run(HybridChoice& c, const Host& h, int calls, bool firstCold = true, int neighbours = -1)
{
    int packed = 0;
    for (int i = 0; i < calls; ++i) {
        const bool p = c.next(neighbours);
        packed += p;
        double rate = p ? h.packing : h.plain;
        if (firstCold && c.calls == 1u) rate /= 3.0;
        c.record(p, rate);
    }
    return packed;
}

#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) {                                                   \
            std::printf("FAILED line %d: %s\n", __LINE__, #cond);        \
            return 1;                                                    \
        }                                                                \
    } while (0)

int main()
{
    {   // one process, link-bound host: plain 53.6 GB/s, packing 61 GB/s -> packs, except the first two calls and one probe in sixteen
        HybridChoice c(48e9);
        const Host h{53.6e9, 61e9};
        CHECK(run(c, h, 2) == 0);
        CHECK(run(c, h, 62, false) == 62 - 4);   // calls 15, 31, 47, 63 probe the plain pipeline
        CHECK(c.ratePlain > 53e9 && c.ratePlain < 54e9);   // the cold first call never entered the mean
    }
    {   // one process per GPU on a shared host: plain uploads reach 35 GB/s -> never packs, whatever packing would measure
        HybridChoice c(48e9);
        CHECK(run(c, Host{35e9, 40e9}, 100) == 0);
        // ... the other ranks finish: plain calls reach link rate again, packing resumes
        CHECK(run(c, Host{53.6e9, 61e9}, 40, false) >= 30);
    }
    {   // link-bound, but packing is slower on this host (few cores): tries it once, then only probes
        HybridChoice c(48e9);
        const int packed = run(c, Host{53.6e9, 45e9}, 66);
        CHECK(packed >= 1 && packed <= 1 + 4);
    }
    {   // the gate can be lowered (GOOFY_B200_HYBRID_MIN_LINK_GBS) for slower links
        HybridChoice c(20e9);
        CHECK(run(c, Host{26e9, 30e9}, 34) >= 28);
    }
    {   // neighbours counted (NVML): with any, never; alone, even a slow link packs (the gate is for the "cannot tell" case)
        HybridChoice c((double)kHybridMinPlainGBs * 1e9);
        CHECK(run(c, Host{51.5e9, 60e9}, 40, true, 1) == 0);
        CHECK(run(c, Host{51.5e9, 60e9}, 40, false, 3) == 0);
        HybridChoice d((double)kHybridMinPlainGBs * 1e9);
        CHECK(run(d, Host{47e9, 60e9}, 34, true, 0) >= 28);
        HybridChoice e((double)kHybridMinPlainGBs * 1e9);
        CHECK(run(e, Host{47e9, 60e9}, 34, true, -1) == 0);
    }
    std::printf("ok\n");
    return 0;
}

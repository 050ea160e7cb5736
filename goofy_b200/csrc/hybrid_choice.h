// hybrid_choice.h -- when does the host path's hybrid scheduler (host_pipeline.cuh: run_hybrid) pack?  Plain C++ (no
// CUDA), so tests/hybrid_choice_host.cpp can drive it through the situations measured on the GPU boxes.
#pragma once
#include <cstdint>

namespace {

// A B200 link (PCIe Gen5 x16) carries a plain host call at 53-54 GB/s of input when nothing else holds it back.
constexpr int kHybridMinPlainGBs = 48;

// One per calling thread.  next() decides whether the coming call on a large pinned image packs (alpha-stripped strips
// from the back of the image) or runs as plain DMA; record() takes the input rate the call then achieved.
// `neighbours` = GPUs of this box that run somebody else's compute process (host_neighbours.cuh), -1 = cannot tell.
//   - the first two calls are plain; the first one is cold (page tables, clocks) and not recorded
//   - no packing with neighbours: their uploads share the host's memory system with this process's packing, and the
//     cost falls on THEM, so comparing one's own calls cannot see it (two ranks: -12 %, four: -12...-22 %)
//   - when the neighbours cannot be counted: no packing while the plain calls stay below `minPlain` bytes per second
//     (something other than the link bounds this process's uploads: four ranks on one host get 27-45 GB/s each)
//   - otherwise the faster of the two running means; every sixteenth call runs the other way to keep both current
struct HybridChoice {
    double minPlain;
    double ratePacking = 0.0, ratePlain = 0.0;   // input bytes per second, running means; 0 = not measured yet
    uint32_t calls = 0;
    explicit HybridChoice(double minPlainBytesPerSecond) : minPlain(minPlainBytesPerSecond) {}
    bool next(int neighbours = -1)   // true: this call packs
    {
        const uint32_t n = calls++;
        if (n < 2u) return false;
        if (neighbours > 0) return false;
        if (neighbours < 0 && ratePlain < minPlain) return false;
        if (ratePacking == 0.0) return true;
        const bool preferred = ratePacking > ratePlain;
        return (n & 15u) == 15u ? !preferred : preferred;
    }
    void record(bool packed, double bytesPerSecond)
    {
        if (calls <= 1u) return;
        double& r = packed ? ratePacking : ratePlain;
        r = r == 0.0 ? bytesPerSecond : 0.5 * r + 0.5 * bytesPerSecond;
    }
};

}  // namespace

// Test shim: include/goofy_png.h behind a C entry point so tests/test_png_loader.py can compare it with PIL.
#include "../include/goofy_png.h"

extern "C" int goofy_png_probe(const char* path, int requireShape, unsigned int* w, unsigned int* h, unsigned char* out, unsigned long long capacity,
                               char* error, int errorCapacity)
{
    goofy::png::Image im = goofy::png::load(path, requireShape != 0);
    if (!im.error.empty()) {
        std::snprintf(error, (size_t)errorCapacity, "%s", im.error.c_str());
        return 1;
    }
    *w = im.width;
    *h = im.height;
    const unsigned long long bytes = (unsigned long long)im.width * im.height * 4;
    const int aligned = ((uintptr_t)im.rgba % 64) == 0;
    if (out && bytes <= capacity) std::memcpy(out, im.rgba, bytes);
    goofy::png::freeImage(im);
    return aligned ? 0 : 2;
}

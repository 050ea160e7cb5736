// goofy_tc_reference.h -- drop-in for the reference's Src/goofy_tc_reference.h (namespace goofyRef),
// backed by the float-reference flavour of the B200 kernels.
//
// goofyRef::compressDXT1 / compressETC1 keep the reference's signatures (Src/goofy_tc_reference.h:7-8)
// and produce byte-identical output to Src/goofy_tc_reference.cpp:794-850 for the tight stride the
// reference harness uses.  (With a padded stride the reference advances block rows by width*16 bytes
// instead of stride*4 -- :818, :847 -- and reads the wrong rows; here `stride` is honoured.)
#ifndef GOOFY_TC_REFERENCE_B200_H
#define GOOFY_TC_REFERENCE_B200_H

#include "goofy_b200.h"

namespace goofyRef {

inline int compressDXT1(unsigned char* result, const unsigned char* input, unsigned int width, unsigned int height, unsigned int stride)
{
    return goofy_b200_compress_dxt1_floatref(result, input, width, height, stride);
}

inline int compressETC1(unsigned char* result, const unsigned char* input, unsigned int width, unsigned int height, unsigned int stride)
{
    return goofy_b200_compress_etc1_floatref(result, input, width, height, stride);
}

}  // namespace goofyRef

#endif

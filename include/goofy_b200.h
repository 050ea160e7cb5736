/*
 * goofy_b200.h -- C ABI of libgoofy_b200.so, the B200 (sm_100a) implementation of Goofy's
 * DXT1/BC1 and ETC1s block encoders.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  The C++
 * wrappers in include/goofy_tc.h (namespace goofy, same signatures as the reference) and
 * the Python mirror in goofy_b200/api.py both sit on top of exactly these entry points.
 * There is NO CPU fallback: every entry point returns a GOOFY_B200_E_* code if CUDA or a
 * B200-class device is unavailable.
 *
 * Reference interface each entry point replaces (paths into the reference repository):
 *   goofy_b200_compress_dxt1  <- goofy::compressDXT1   GoofyTC/goofy_tc.h:11, defined :1497-1526
 *   goofy_b200_compress_etc1  <- goofy::compressETC1   GoofyTC/goofy_tc.h:12, defined :1528-1557
 *   (function-pointer shape)  <- CompressFunc_t         Src/main.cpp:644
 * The *_device / *_batch / *_dual / *_sharded entry points are the batched device-resident
 * variants BASELINE.json's north_star adds; the per-block arithmetic they run is the same
 * (goofySimdEncode<>, GoofyTC/goofy_tc.h:1069-1494).
 *
 * Image contract (identical to the reference, GoofyTC/goofy_tc.h:1497-1523 and :175-178):
 *   - input is RGBA8, `stride` bytes between rows (may exceed width*4), alpha ignored;
 *   - width % 16 == 0 (else -1), height % 4 == 0 (else -2), checked in that order;
 *   - rows 16-byte aligned: input % 16 == 0 and stride % 16 == 0;
 *   - output is (height/4)*(width/4) blocks of 8 bytes, row-major, width*height/2 bytes;
 *   - width == 0 or height == 0 returns 0 and writes nothing.
 */
#ifndef GOOFY_B200_H
#define GOOFY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOOFY_B200_ABI_VERSION 3   /* 2: GoofyB200Image gained dst2, GOOFY_B200_BOTH; 3: packed-RGB input, host-path knobs */

/* codec selectors */
#define GOOFY_B200_DXT1 0
#define GOOFY_B200_ETC1 1
/* Both codecs from ONE read of every pixel (5 B/px of HBM traffic instead of 9): DXT1 blocks to `dst`, ETC1s blocks
 * to `dst2`.  Accepted by the batch entry points whose descriptors carry two result pointers
 * (goofy_b200_encode_batch_device, _batch_sharded, _host_batch); the uniform entry points have
 * goofy_b200_encode_dual_device / _dual_host / _dual_sharded_host for it. */
#define GOOFY_B200_BOTH 2
/* Second flavour: bit-exact with the reference's float "idea" encoder goofyRef::compressDXT1/ETC1
 * (Src/goofy_tc_reference.cpp:794-850) instead of with the SSE2 path.  The two flavours differ
 * by design (rounding, tie-break, minimum range, table thresholds).  Accepted by the host,
 * device and uniform-batch entry points; width only needs to be a multiple of 4 (:796). */
#define GOOFY_B200_DXT1_FLOATREF 16
#define GOOFY_B200_ETC1_FLOATREF 17

/* return codes: 0, -1, -2 are the reference's; the rest are new and never collide with them */
#define GOOFY_B200_OK 0
#define GOOFY_B200_E_WIDTH (-1)      /* width % 16 != 0            (goofy_tc.h:1500-1503) */
#define GOOFY_B200_E_HEIGHT (-2)     /* height % 4 != 0            (goofy_tc.h:1505-1508) */
#define GOOFY_B200_E_NULL (-3)       /* null pointer with non-empty image */
#define GOOFY_B200_E_ALIGN (-4)      /* input/stride not 16-byte, output not 8-byte aligned */
#define GOOFY_B200_E_STRIDE (-5)     /* stride < width*4 */
#define GOOFY_B200_E_CODEC (-6)      /* unknown codec selector */
#define GOOFY_B200_E_DEVICE (-7)     /* no CUDA device / bad device index / wrong device for pointer */
#define GOOFY_B200_E_ARGS (-8)       /* other invalid argument (batch size, pitch, ...) */
#define GOOFY_B200_E_CUDA_BASE (-100) /* CUDA failure: code = -100 - cudaError_t */

/* One image of a batch.  All pointers are device pointers on `device`
 * (device < 0: the calling thread's current device). */
typedef struct GoofyB200Image {
    const void* src; /* RGBA8, 16-byte aligned */
    void* dst;       /* width*height/2 bytes, 8-byte aligned (GOOFY_B200_BOTH: the DXT1 blocks) */
    uint32_t width;
    uint32_t height;
    uint32_t stride; /* bytes */
    int32_t device;
    void* dst2;      /* GOOFY_B200_BOTH only: the ETC1s blocks, width*height/2 bytes, 8-byte aligned; ignored otherwise */
} GoofyB200Image;

int goofy_b200_abi_version(void);
int goofy_b200_device_count(void);
/* Static description of a return code (never NULL). */
const char* goofy_b200_error_string(int code);
/* Kernels launched by this library in the calling process so far (all threads, all devices). */
uint64_t goofy_b200_kernel_launches(void);
/* Name of the encode kernel the CALLING THREAD launched last through this library ("" before the first launch), e.g.
 * "encode_direct_kernel<dxt1>" -- what a call actually ran, for benchmarks and profiles.  Static storage. */
const char* goofy_b200_last_launch_kernel(void);
/* Sets of host-path scratch (streams, device strips, pinned staging strips, descriptor arena) created in this process
 * so far.  A host thread leases one set while it lives and returns it to a pool when it exits, so this number tracks
 * the largest number of threads that were inside the host-pointer / batch entry points at the same time, not the
 * number of threads that ever called. */
uint64_t goofy_b200_host_scratch_sets(void);

/* Image load layer used by the uniform device entry points (process-wide):
 *   AUTO    the library picks per shape (default)
 *   DIRECT  one thread per block, four coalesced 128-bit global loads, persistent CTAs walking down the image
 *   TMA     2D tensor-map tiles staged in shared memory by persistent CTAs
 * Both produce identical bytes.  set returns the previous setting (or GOOFY_B200_E_ARGS). */
#define GOOFY_B200_LOAD_AUTO 0
#define GOOFY_B200_LOAD_DIRECT 1
#define GOOFY_B200_LOAD_TMA 2
#define GOOFY_B200_LOAD_ONESHOT 3 /* DIRECT with one-shot CTAs instead of persistent row-walking CTAs */
#define GOOFY_B200_LOAD_ASYNC 4   /* row-walking CTAs prefetching the next block row with cp.async into shared memory */
int goofy_b200_set_load_path(int path);
int goofy_b200_get_load_path(void);

/* Host path: drop the alpha byte while staging (the encoders ignore it; 4 of the 4.5 bytes per pixel that cross PCIe
 * are input).  Process-wide; set returns the previous setting (or GOOFY_B200_E_ARGS).
 *   OFF     every pixel crosses the link as RGBA
 *   AUTO    (default) pageable input is staged as packed RGB -- the staging copy is made anyway; large pinned input is
 *           split between plain DMA and alpha-stripped strips according to how fast the host's cores pack -- while no
 *           other GPU of the box runs somebody else's compute process (goofy_b200_host_neighbours: one process per
 *           GPU on a shared host is bound by host memory, not by the link, and packing then makes everybody slower) and
 *           while calls that pack measure faster than calls that do not
 *   ALWAYS  every strip of a large pinned image is alpha-stripped first (experiments)
 *   PAGEABLE  pageable input only; pinned input always goes plain DMA.  For launchers that run ONE PROCESS PER GPU on a
 *           shared host where NVML is not available to AUTO: the packing of one rank loads the host memory all ranks
 *           upload from, which no rank can see in its own measurements (2 ranks: +6 % on one box, -12 % on another;
 *           4 ranks -12...-22 %)
 * The bytes produced are identical in every mode.  Environment: GOOFY_B200_HOST_RGB=0|1|2|3 (initial mode),
 * GOOFY_B200_HOST_THREADS=n (host threads per staging job, the caller included; default min(8, cores / 2)). */
#define GOOFY_B200_HOST_RGB_OFF 0
#define GOOFY_B200_HOST_RGB_AUTO 1
#define GOOFY_B200_HOST_RGB_ALWAYS 2
#define GOOFY_B200_HOST_RGB_PAGEABLE 3
int goofy_b200_set_host_rgb_staging(int mode);
int goofy_b200_get_host_rgb_staging(void);
/* Host threads that work on one staging job (copy or alpha strip), the calling thread included. */
int goofy_b200_host_threads(void);
/* GPUs of this box that run somebody else's compute process, as AUTO sees them (NVML, all GPUs whatever
 * CUDA_VISIBLE_DEVICES says; this process's own contexts do not count); -1 = cannot tell (no NVML).  With neighbours AUTO
 * never alpha-strips pinned input: their uploads share the host's memory system with the packing. */
int goofy_b200_host_neighbours(void);
/* What the host path sent over the link so far in this process (all threads): bytes host -> device (copy engine and
 * zero-copy kernel reads); strips of large pinned images sent as they were / alpha-stripped first; calls on large pinned
 * images that ran with packing / as plain DMA (AUTO measures both and uses the faster).  Any pointer may be NULL. */
void goofy_b200_host_link_stats(uint64_t* bytes_uploaded, uint64_t* raw_strips, uint64_t* packed_strips,
                                uint64_t* packing_calls, uint64_t* plain_calls);

/* ---- drop-in host-pointer API: same arguments, order and return codes as the reference ---- */
int goofy_b200_compress_dxt1(unsigned char* result, const unsigned char* input, unsigned int width,
                             unsigned int height, unsigned int stride);
int goofy_b200_compress_etc1(unsigned char* result, const unsigned char* input, unsigned int width,
                             unsigned int height, unsigned int stride);
/* goofyRef::compressDXT1 / compressETC1 (Src/goofy_tc_reference.h:7-8) -- the float-reference flavour. */
int goofy_b200_compress_dxt1_floatref(unsigned char* result, const unsigned char* input, unsigned int width,
                                      unsigned int height, unsigned int stride);
int goofy_b200_compress_etc1_floatref(unsigned char* result, const unsigned char* input, unsigned int width,
                                      unsigned int height, unsigned int stride);
/* Same, codec chosen at run time.  Host buffers may be pageable or pinned; pinned buffers
 * (cudaHostAlloc / cudaHostRegister) are copied without a staging hop.  Uses the calling
 * thread's current device and synchronises before returning. */
int goofy_b200_encode_host(int codec, void* result, const void* input, uint32_t width, uint32_t height,
                           uint32_t stride);

/* Both codecs from ONE upload of a host image (the dual-output kernel behind the host pipeline): 4 B/px cross the link
 * instead of 8 for two single-codec calls.  Same argument checks and codes as the single-codec host calls. */
int goofy_b200_encode_dual_host(void* result_dxt1, void* result_etc1, const void* input, uint32_t width, uint32_t height,
                                uint32_t stride);

/* A host image of packed RGB8 pixels (3 bytes per pixel, rows `stride` >= width*3 bytes apart; input and stride 4-byte
 * aligned): 3 instead of 4 input bytes per pixel cross PCIe with no host-side work at all -- for callers whose decoder
 * delivers RGB (the reference's own loader widens to RGBA and sets alpha to 0xFF, Src/main.cpp:328-335, only because
 * its encoders want 16-byte pixels groups).  Same bytes as the RGBA calls on the same pixels.  codec: GOOFY_B200_DXT1,
 * _ETC1, _BOTH (result2 = the ETC1s blocks, ignored otherwise) or a float-reference flavour. */
int goofy_b200_encode_rgb24_host(int codec, void* result, void* result2, const void* input, uint32_t width, uint32_t height,
                                 uint32_t stride);

/* n host images (src / dst / dst2 are HOST pointers here, `device` is ignored; codec may be GOOFY_B200_BOTH) through one pipeline on the calling thread's
 * current device: the copies and kernels of neighbouring images overlap, which a loop of single-image calls (each of
 * which waits for its own result) cannot do -- the host-side batch for the reference harness's per-image loop
 * (Src/main.cpp:646-743).  Every image is validated before anything starts; returns when all results are in place. */
int goofy_b200_encode_host_batch(int codec, const GoofyB200Image* images, uint32_t n_images);

/* ---- device-resident API: pointers are device memory on the current device; asynchronous
 *      on `stream` (a cudaStream_t, NULL = default stream); no allocation, no sync ---- */
int goofy_b200_encode_device(int codec, void* d_result, const void* d_input, uint32_t width, uint32_t height,
                             uint32_t stride, void* stream);

/* Relaxed shapes (beyond the reference, which rejects them): any width and height >= 1, stride and input only
 * 4-byte aligned.  Blocks overhanging the right / bottom edge replicate the last column / row, so the result
 * holds ceil(width/4) * ceil(height/4) blocks.  On shapes the strict entry points accept, the bytes are identical.
 * Meant for odd-sized mip levels; it uses scalar loads and is not the bandwidth-optimal path. */
int goofy_b200_encode_relaxed_device(int codec, void* d_result, const void* d_input, uint32_t width, uint32_t height,
                                     uint32_t stride, void* stream);

/* Packed RGB8 input: 3 bytes per pixel, NO alpha byte (the encoders ignore alpha, GoofyTC/goofy_tc.h:297), rows
 * `stride` >= width*3 bytes apart; d_input, stride and input_image_pitch only need 4-byte alignment.  The bytes
 * produced are those the RGBA entry points produce for the same pixels with any alpha.  3.5 instead of 4.5 bytes of
 * memory traffic per pixel, for callers whose pixels are RGB to begin with (decoded PNG / JPEG) and for the host path,
 * which drops the alpha byte while staging (goofy_b200_set_host_rgb_staging).  codec: GOOFY_B200_DXT1, _ETC1, _BOTH
 * (d_result2 = the ETC1s blocks, ignored otherwise) or a float-reference flavour.  n_images at fixed pitches
 * (pitches ignored for n_images == 1). */
int goofy_b200_encode_rgb24_device(int codec, void* d_result, void* d_result2, const void* d_input, uint32_t width,
                                   uint32_t height, uint32_t stride, uint64_t input_image_pitch,
                                   uint64_t result_image_pitch, uint32_t n_images, void* stream);

/* n images of one shape laid out at fixed pitches (bytes) from d_input / d_result. */
int goofy_b200_encode_batch_uniform_device(int codec, void* d_result, const void* d_input, uint32_t width,
                                           uint32_t height, uint32_t stride, uint64_t input_image_pitch,
                                           uint64_t result_image_pitch, uint32_t n_images, void* stream);

/* Both codecs from one read of the input (5 B/px instead of 9 B/px of HBM traffic). */
int goofy_b200_encode_dual_device(void* d_result_dxt1, void* d_result_etc1, const void* d_input, uint32_t width,
                                  uint32_t height, uint32_t stride, uint64_t input_image_pitch,
                                  uint64_t result_image_pitch, uint32_t n_images, void* stream);

/* n images of arbitrary shapes on the current device (descs is a HOST array; descs[i].device
 * must be < 0 or the current device).  One launch per call; the descriptor table is copied
 * to the device on `stream`.  codec: GOOFY_B200_DXT1, _ETC1, _BOTH (descs[i].dst2 holds the ETC1s blocks) or one
 * of the float-reference flavours (widths then only need to be multiples of 4). */
int goofy_b200_encode_batch_device(int codec, const GoofyB200Image* descs, uint32_t n_images, void* stream);

/* ---- the step after the encoder (reference: Src/main.cpp:561-613 decode, :403-469 MSE/PSNR) ---- */
/* Decode width*height/2 bytes of BC1 / ETC1 blocks into RGBA8 (alpha 255; BC1 3-colour
 * "transparent" index writes 0,0,0,0).  width and height multiples of 4. */
int goofy_b200_decode_device(int codec, void* d_rgba, const void* d_blocks, uint32_t width, uint32_t height,
                             uint32_t stride, void* stream);
/* Decode and compare with the source pixels without writing the decoded image: ADDS the sum of
 * squared errors of the R, G and B channels to d_sse_rgb[0..2] (device uint64, caller zeroes).
 * psnrRGB of the reference harness = 10*log10(768^2 / ((sse[0]+sse[1]+sse[2]) / pixels)). */
int goofy_b200_block_sse_device(int codec, const void* d_blocks, const void* d_rgba, uint32_t width, uint32_t height,
                                uint32_t stride, uint64_t* d_sse_rgb, void* stream);

/* ---- multi-GPU shard scheduler (no collectives: blocks are independent) ---- */
/* Images may live on different devices (descs[i].device >= 0 required).  Work is grouped per
 * device, launched from one host thread per device, and all devices are synchronised
 * before returning. */
int goofy_b200_encode_batch_sharded(int codec, const GoofyB200Image* descs, uint32_t n_images);
/* One host image split into n_gpus horizontal strips of whole block rows; strip g is
 * encoded on device g through the host path.  n_gpus <= 0 means all visible devices. */
int goofy_b200_encode_sharded_host(int codec, void* result, const void* input, uint32_t width, uint32_t height,
                                   uint32_t stride, int n_gpus);
/* Same partition, both codecs from one upload of every strip (goofy_b200_encode_dual_host per device). */
int goofy_b200_encode_dual_sharded_host(void* result_dxt1, void* result_etc1, const void* input, uint32_t width,
                                        uint32_t height, uint32_t stride, int n_gpus);
/* Strip partition used by the scheduler: block rows [first, first+count) for shard `shard`. */
void goofy_b200_strip_partition(uint32_t height, int n_shards, int shard, uint32_t* first_block_row,
                                uint32_t* block_row_count);

#ifdef __cplusplus
}
#endif
#endif /* GOOFY_B200_H */

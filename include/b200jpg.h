/*
 * b200jpg.h -- the C ABI of the B200-native baseline-JPEG decode path.
 *
 * This is the drop-in boundary beneath the reference's C++ interface (include/interface/jpeg.hpp restates
 * thorfdbg/libjpeg's interface/jpeg.hpp:79-250): plain pointers and sizes, no C++ and no torch types.
 * The C++ `class JPEG` shim (libjpeg_b200/csrc/jpeg_shim.cpp) calls exactly these entry points; the batch
 * entry points are what the throughput benchmark and the Python host layer bind (ctypes).
 *
 * What each entry point replaces in the reference (file:line relative to the reference tree):
 *   b200jpg_parse            Decoder::ParseHeaderIncremental            codestream/decoder.cpp:77
 *                            Tables::ParseTablesIncremental             codestream/tables.cpp:1003-1420
 *                            Frame::ParseMarker / Scan::ParseMarker     marker/frame.cpp:111, marker/scan.cpp:163
 *                            EntropyParser::ParseRestartMarker (index)  codestream/entropyparser.cpp:117-136
 *   b200jpg_batch_decode     SequentialScan::ParseMCU / DecodeBlock     codestream/sequentialscan.cpp:381,678
 *     (entropy stage)        HuffmanDecoder::Get, BitStream<false>      coding/huffmandecoder.hpp:103, io/bitstream.cpp:56
 *     (reconstruction stage) BlockBitmapRequester::ReconstructRegion    control/blockbitmaprequester.cpp:1249
 *                            IDCT::InverseTransformBlock                dct/idct.cpp:226
 *                            Upsampler<sx,sy>::UpsampleRegion           upsampling/upsampler.cpp:83
 *                            YCbCrTrafo::YCbCr2RGB                      colortrafo/ycbcrtrafo.cpp:679
 *   b200jpg_last_error       JPEG::LastError                            interface/jpeg.cpp:962
 *
 * Error codes are the reference's (interface/parameters.hpp:1156-1228).
 * There is NO CPU fallback: every decode entry point fails with B200JPG_ERR_NO_DEVICE when no CUDA device
 * is usable.
 */
#ifndef B200JPG_H
#define B200JPG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef B200JPG_API
#define B200JPG_API __attribute__((visibility("default")))
#endif

#define B200JPG_OK 0
#define B200JPG_ERR_INVALID_PARAMETER (-1024)
#define B200JPG_ERR_UNEXPECTED_EOF (-1025)
#define B200JPG_ERR_OBJECT_DOESNT_EXIST (-1031)
#define B200JPG_ERR_NOT_IMPLEMENTED (-1034)
#define B200JPG_ERR_MALFORMED_STREAM (-1038)
#define B200JPG_ERR_OUT_OF_MEMORY (-2048)
#define B200JPG_ERR_NO_DEVICE (-8193) /* in the reference's USER_ERROR range (-8192 and below) */
#define B200JPG_ERR_CUDA (-8194)

#define B200JPG_MAX_COMPONENTS 4
#define B200JPG_MAX_SCANS 16

/* Geometry of one parsed codestream (host only, no GPU needed). */
typedef struct b200jpg_frame_info {
    uint32_t width, height;
    uint8_t ncomp, precision, frame_type /* 0 = SOF0, 1 = SOF1, 2 = SOF2 (progressive) */, ycbcr /* 1: YCbCr->RGB applies */;
    uint8_t comp_id[B200JPG_MAX_COMPONENTS];
    uint8_t hs[B200JPG_MAX_COMPONENTS], vs[B200JPG_MAX_COMPONENTS];     /* sampling factors of the SOF */
    uint8_t subx[B200JPG_MAX_COMPONENTS], suby[B200JPG_MAX_COMPONENTS]; /* hmax/hs, vmax/vs (reference's SubX/SubY) */
    uint8_t tq[B200JPG_MAX_COMPONENTS];
    uint32_t mcu_cols, mcu_rows;                                       /* interleaved MCU grid */
    uint32_t blocks_w[B200JPG_MAX_COMPONENTS], blocks_h[B200JPG_MAX_COMPONENTS]; /* MCU-padded block grid */
    uint32_t nscans;
    uint32_t restart_interval;  /* DRI in effect at the first scan, MCUs; 0 = none */
    uint32_t n_intervals;       /* restart intervals over all scans */
    uint64_t ecs_bytes;         /* entropy coded bytes over all scans, as stored (stuffing + RSTn included) */
    uint64_t stored_blocks;     /* sum_c ceil(ceil(W/subx)/8)*ceil(ceil(H/suby)/8): the reference's block store */
} b200jpg_frame_info;

/* --- host-side parse, no device ------------------------------------------------------------------ */
/* Parses the marker segments of one codestream. Returns B200JPG_OK or a negative error code. */
B200JPG_API int b200jpg_parse(const uint8_t *data, size_t len, b200jpg_frame_info *info);

/* Builds the device-independent Huffman / quantisation table blob of scan `scan` of one codestream on the host
 * (replaces HuffmanTemplate::BuildDecoder coding/huffmantemplate.cpp:802-874 and IDCT::DefineQuant dct/idct.cpp:98-108).
 * Returns the blob size (also when dst is NULL or too small), 0 on error. This blob is what rank 0 broadcasts. */
B200JPG_API uint64_t b200jpg_build_tables(const uint8_t *data, size_t len, int scan, uint8_t *dst, uint64_t capacity);

/* --- device context -------------------------------------------------------------------------------- */
typedef struct b200jpg_ctx b200jpg_ctx;

/* Creates a decode context on CUDA device `device` (current device if < 0). */
B200JPG_API int b200jpg_create(int device, b200jpg_ctx **ctx);
B200JPG_API void b200jpg_destroy(b200jpg_ctx *ctx);
/* Frees the device / pinned buffers the context keeps for reuse by later batches (buffers of live batches are untouched). */
B200JPG_API void b200jpg_trim(b200jpg_ctx *ctx);
/* Message and code of the last failure on this context (code 0 / "" when none). ctx may be NULL for
 * failures of b200jpg_create / b200jpg_parse on the calling thread. */
B200JPG_API int b200jpg_last_error(b200jpg_ctx *ctx, const char **message);

/* --- batches --------------------------------------------------------------------------------------- */
typedef struct b200jpg_batch b200jpg_batch;

/* Host stage: parses `n` codestreams (host pointers), builds the restart-interval index and the table
 * sets, and packs everything the kernels need into pinned staging memory.  Output layout: frame i is
 * written at out + out_offset(i) as interleaved 8-bit pixels, ncomp bytes per pixel, row pitch
 * width*ncomp (the layout a BitMapHook client with BytesPerPixel = depth gets from the reference).
 * Frames that fail to parse make the call fail unless `tolerate_bad` is non-zero, in which case they are
 * skipped and reported through b200jpg_batch_frame_status. */
B200JPG_API int b200jpg_batch_create(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, int tolerate_bad,
                         b200jpg_batch **batch);
/* The same with request flags -- what a RectangleRequest carries besides the rectangle (codestream/rectanglerequest.cpp:
 * 93-165): B200JPG_FLAG_NO_COLOR_TRANSFORM = JPGTAG_MATRIX_LTRAFO set to JPGFLAG_MATRIX_COLORTRANSFORMATION_NONE: the
 * components are upsampled and delivered as they are (YCbCr frames come out as Y, Cb, Cr). */
#define B200JPG_FLAG_NO_COLOR_TRANSFORM 1u
/* B200JPG_FLAG_NO_UPSAMPLE = JPGTAG_DECODER_UPSAMPLE false (control/bitmapctrl.cpp:273-293, BlockBitmapRequester::
 * ReconstructUnsampled control/blockbitmaprequester.cpp:1013-1074): no upsampling and no colour transformation (the reference
 * refuses one without the other); frame i is written as plane after plane, component c holding ceil(width / subx[c]) x
 * ceil(height / suby[c]) samples, rows tightly packed.  Samples are bytes, or native-endian 16-bit for 12-bit frames. */
#define B200JPG_FLAG_NO_UPSAMPLE 2u
B200JPG_API int b200jpg_batch_create_ex(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, int tolerate_bad,
                            unsigned flags, b200jpg_batch **batch);
/* Waits for the work last enqueued for this batch (its buffers return to the context's pool for reuse by the
 * next batch; b200jpg_destroy frees the pool). The caller must have consumed `out_dev` before reusing it. */
B200JPG_API void b200jpg_batch_destroy(b200jpg_batch *batch);

B200JPG_API int b200jpg_batch_frame_info(const b200jpg_batch *batch, int i, b200jpg_frame_info *info);
/* Byte offset / size of frame i inside the batch output buffer; total size with i = -1. */
B200JPG_API uint64_t b200jpg_batch_out_offset(const b200jpg_batch *batch, int i);
B200JPG_API uint64_t b200jpg_batch_out_bytes(const b200jpg_batch *batch, int i);
/* Algorithmic-traffic accounting for the roofline: ECS bytes and stored blocks of the batch. */
B200JPG_API uint64_t b200jpg_batch_ecs_bytes(const b200jpg_batch *batch);
B200JPG_API uint64_t b200jpg_batch_stored_blocks(const b200jpg_batch *batch);
B200JPG_API uint64_t b200jpg_batch_h2d_bytes(const b200jpg_batch *batch);

/* Replace the Huffman/quantisation table blob of this batch by an externally supplied device-independent
 * blob (multi-GPU: rank 0 exports, NCCL broadcasts, other ranks import). export returns the size. */
B200JPG_API uint64_t b200jpg_batch_export_tables(const b200jpg_batch *batch, uint8_t *dst, uint64_t capacity);
B200JPG_API int b200jpg_batch_import_tables(b200jpg_batch *batch, const uint8_t *src, uint64_t size);

/* H2D of the packed codestream bytes + descriptors on `stream` (a cudaStream_t, 0 = default). */
B200JPG_API int b200jpg_batch_upload(b200jpg_batch *batch, void *stream);

/* Re-runs only the device-side restart index that b200jpg_batch_upload builds behind its copy (EntropyParser::
 * ParseRestartMarker bookkeeping, codestream/entropyparser.cpp:117-136) -- idempotent; lets a caller time it. */
B200JPG_API int b200jpg_batch_reindex(b200jpg_batch *batch, void *stream);

/* The hot path: entropy decode kernel(s) then reconstruction kernel(s), asynchronous on `stream`.
 * `out_dev` is a DEVICE pointer to b200jpg_batch_out_bytes(batch,-1) bytes. */
B200JPG_API int b200jpg_batch_decode(b200jpg_batch *batch, uint8_t *out_dev, void *stream);
/* Only one of the two stages (profiling / roofline measurements / parity of the intermediate). */
B200JPG_API int b200jpg_batch_decode_entropy(b200jpg_batch *batch, void *stream);
B200JPG_API int b200jpg_batch_reconstruct(b200jpg_batch *batch, uint8_t *out_dev, void *stream);

/* After the stream has been synchronised: per-frame status (0 or a negative error code). Performs a small
 * D2H copy of the error words on first use. */
B200JPG_API int b200jpg_batch_frame_status(b200jpg_batch *batch, int i);

/* Debug / parity access: copies the dequantised coefficient plane of component c of frame i to the host
 * as int16 [blocks_h][blocks_w][64] in raster order inside the block. */
B200JPG_API int b200jpg_batch_read_coefficients(b200jpg_batch *batch, int i, int c, int16_t *dst, uint64_t capacity_elems);

/* Number of kernel launches the last b200jpg_batch_decode* call issued. */
B200JPG_API int b200jpg_batch_last_launch_count(const b200jpg_batch *batch);
/* CUDA-event timing of the two stages of the last decode call when timing is enabled (ms); timing makes
 * the call record events on `stream`, it does not synchronise. Read after synchronising. */
B200JPG_API void b200jpg_batch_enable_timing(b200jpg_batch *batch, int on);
B200JPG_API int b200jpg_batch_last_timing(b200jpg_batch *batch, float *entropy_ms, float *reconstruct_ms);
/* Share of entropy_ms spent in the unstuffing pre-pass (stage a0), ms; negative when not available. */
B200JPG_API float b200jpg_batch_last_unstuff_ms(b200jpg_batch *batch);

/* Measured int32 issue rate of the device in Gop/s (integer multiply-add counted as 2 ops):
 * multiply-add only, add/logic only, and a 1:1 mix. Denominator of the reconstruction kernel's roofline. */
B200JPG_API int b200jpg_microbench_int32(int device, float *imad_gops, float *alu_gops, float *mix_gops);

/* Self-test of the host logic behind restart-less scans (tests only, no device): replays the synchronisation rounds of the
 * device kernel for scan 0 of one codestream on the host and checks the resulting work items against a front-to-back walk.
 * Returns B200JPG_OK, a parser error, or -1..-4 for an inconsistent partition. */
B200JPG_API int b200jpg_selftest_restartless(const uint8_t *data, size_t len, uint32_t *rounds, uint32_t *n_segments);
/* Host-only self-test of the decoder-table cache of b200jpg_batch_create (frames of one source share their tables; the tables of a
 * scan are built once per distinct set of Huffman specifications, quantisation tables and scan parameters): every scan of the
 * given codestreams through the cache, twice, against a fresh HuffmanTemplate::BuildDecoder-equivalent build
 * (coding/huffmantemplate.cpp:802-874).  Returns the number of scans compared, or a negative value on a mismatch. */
B200JPG_API int b200jpg_selftest_table_cache(const uint8_t *const *frames, const size_t *lens, int n);

/* One-call convenience used by the C++ JPEG shim: host codestreams in, HOST pixels out (upload, decode,
 * download, synchronise).  `out_host` receives b200jpg_batch_out_bytes(batch,-1) bytes. */
B200JPG_API int b200jpg_decode_to_host(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, uint8_t *out_host,
                           uint64_t out_capacity);
B200JPG_API int b200jpg_decode_to_host_ex(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, uint8_t *out_host,
                              uint64_t out_capacity, unsigned flags);

/* The same, but the pixels STAY ON THE DEVICE (consumers that work on the GPU skip the PCIe copy of 24.9 MB per 4K frame):
 * *out_dev receives a device buffer (frame i at the 256-byte aligned running offset, frame 0 at 0; *out_bytes its size), to be
 * released with b200jpg_device_free.  b200jpg_device_copy_rect copies a rectangle of rows between two device bitmaps (what the
 * JPEG::DisplayRectangle shim does for clients whose BitMapHook hands out device pointers, JPGTAG_B200_DEVICE_BITMAPS). */
B200JPG_API int b200jpg_decode_to_device_ex(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, unsigned flags,
                                uint8_t **out_dev, uint64_t *out_bytes);
B200JPG_API void b200jpg_device_free(b200jpg_ctx *ctx, uint8_t *p);
B200JPG_API int b200jpg_device_copy_rect(b200jpg_ctx *ctx, uint8_t *dst_dev, int64_t dst_pitch, const uint8_t *src_dev, int64_t src_pitch,
                             uint64_t width_bytes, uint64_t rows);

#ifdef __cplusplus
}
#endif
#endif

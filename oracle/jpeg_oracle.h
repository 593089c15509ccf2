/*
 * oracle/jpeg_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's (thorfdbg/libjpeg) Huffman decode path -- sequential
 * (SOF0/SOF1) and, as groundwork for SURVEY 8f2, progressive (SOF2: first passes in
 * codestream/sequentialscan.cpp, refinement passes in codestream/refinementscan.cpp):
 * marker parse -> Huffman decode -> dequant + integer IDCT -> centred chroma upsampling (factors 1..4) ->
 * YCbCr->RGB -> 8-bit store.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use it, and only as the checker.  The product (libjpeg_b200/) never links or calls it.
 *
 * Parity is PINNED: tests/test_oracle.py compares this restatement byte for byte with the
 * unmodified reference built by oracle/Makefile (oracle/_ref/refharness) and with the committed golden
 * fixtures under tests/golden/ that were produced by that reference build.
 */
#ifndef B200JPG_ORACLE_H
#define B200JPG_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: the reference's numeric values (interface/parameters.hpp:1156-1228) */
#define JPGO_OK 0
#define JPGO_ERR_INVALID_PARAMETER (-1024)
#define JPGO_ERR_UNEXPECTED_EOF (-1025)
#define JPGO_ERR_NOT_IMPLEMENTED (-1034)
#define JPGO_ERR_MALFORMED_STREAM (-1038)
#define JPGO_ERR_OUT_OF_MEMORY (-2048)

#define JPGO_MAX_COMP 4
#define JPGO_MAX_SCANS 16

typedef struct {
    int ns;                      /* components in scan */
    int comp[JPGO_MAX_COMP];     /* frame component index, SOS order */
    int td[JPGO_MAX_COMP], ta[JPGO_MAX_COMP];
    int restart_interval;        /* DRI in effect at this SOS (MCUs), 0 = none */
    size_t ecs_offset;           /* first byte after the SOS header */
    size_t ecs_end;              /* offset of the first non-RST marker after the ECS */
    int mcu_cols, mcu_rows;      /* MCU grid of THIS scan (block grid for ns==1) */
    int ss, se, ah, al;          /* spectral selection and successive approximation (0, 63, 0, point transform for sequential) */
} jpgo_scan;

typedef struct {
    int width, height, ncomp, precision;
    int frame_type;              /* 0 = SOF0 baseline, 1 = SOF1 extended sequential, 2 = SOF2 progressive */
    int cid[JPGO_MAX_COMP], hs[JPGO_MAX_COMP], vs[JPGO_MAX_COMP], tq[JPGO_MAX_COMP];
    int hmax, vmax;
    int subx[JPGO_MAX_COMP], suby[JPGO_MAX_COMP];
    int mcu_cols, mcu_rows;      /* interleaved MCU grid of the frame */
    int bw[JPGO_MAX_COMP], bh[JPGO_MAX_COMP];   /* MCU-padded block grid per component */
    int sbw[JPGO_MAX_COMP], sbh[JPGO_MAX_COMP]; /* stored grid of the reference (blockbuffer.cpp:212-265) */
    int ycbcr;                   /* 1: YCbCr->RGB applies (tables.cpp:2023-2030) */
    int nscans;
    jpgo_scan scan[JPGO_MAX_SCANS];
    uint16_t quant[4][64];       /* raster order (quantization.cpp:501-526) */
    int quant_defined[4];
} jpgo_info;

/* Parse all markers up to EOI; fills info. Returns JPGO_OK or a negative reference error code. */
int jpgo_read_info(const uint8_t *data, size_t len, jpgo_info *info);

/* Entropy-decode every scan. planes[c] must hold bw[c]*bh[c]*64 int32 (raster order inside the block,
 * QUANTIZED values exactly as the reference stores them in QuantizedRow blocks).  Zero-initialised here. */
int jpgo_decode_coefficients(const uint8_t *data, size_t len, const jpgo_info *info, int32_t *const planes[]);

/* Reconstruct pixels from quantized coefficient planes: out = interleaved 8-bit, ncomp bytes per pixel,
 * row pitch width*ncomp (what the reference writes through BitMapHook with BytesPerPixel = depth). */
int jpgo_reconstruct(const jpgo_info *info, int32_t *const planes[], uint8_t *out);

/* Whole path. out must hold width*height*ncomp bytes; 8-bit frames only. */
int jpgo_decode(const uint8_t *data, size_t len, uint8_t *out, size_t out_capacity, jpgo_info *info_out);
/* The same into native-endian 16-bit samples (what a CTYP_UWORD client bitmap receives): 8- and 12-bit frames. */
int jpgo_decode16(const uint8_t *data, size_t len, uint16_t *out, size_t capacity_in_samples, jpgo_info *info_out);
/* 1 when the stream carries JPEG XT boxes beyond the file type (merging specification, residual codestream, ...): jpgo_decode then
 * merges the residual layer (the 8-bit integer profile) or reports NOT_IMPLEMENTED */
int jpgo_has_xt_layer(const uint8_t *data, size_t len);
int jpgo_reconstruct16(const jpgo_info *info, int32_t *const planes[], uint16_t *out);
/* JPGTAG_DECODER_UPSAMPLE = false: the components as planes at their own resolution, 16-bit samples, no colour transformation */
int jpgo_reconstruct_planes16(const jpgo_info *info, int32_t *const planes[], uint16_t *out);
int jpgo_decode_planes16(const uint8_t *data, size_t len, uint16_t *out, size_t capacity_in_samples, jpgo_info *info_out);

/* building blocks, exported so the tests can hit them directly */
void jpgo_idct_block(int32_t *target, const int32_t *source, const uint16_t *delta_raster, int32_t dcoffset);
int jpgo_build_huffman(const uint8_t bits[16], const uint8_t *vals, int nvals, uint8_t sym1[256], uint8_t len1[256],
                       uint8_t *sym2 /* [256][256] */, uint8_t *len2 /* [256][256] */, uint8_t has2[256]);
extern const int jpgo_scan_order[64];

#ifdef __cplusplus
}
#endif
#endif

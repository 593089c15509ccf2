// internal.hpp -- shared declarations of the B200 baseline-JPEG decode path (host parser <-> CUDA stages).
//
// Data layout in HBM (one "batch" = n codestreams decoded by the same launches):
//   bytes      : the codestreams, each copied whole at a 16-byte aligned offset and followed by FF D9 + zero
//                padding, so a bit reader that runs off a damaged segment always meets a marker.
//   clean      : the entropy coded segments with byte stuffing removed, one 16-byte aligned run per restart
//                interval, stored as big-endian 32-bit words and followed by >= 32 zero bytes (stage a0 output).
//   coef       : int16, one 128-byte block per 8x8 DCT block, DEQUANTISED (coefficient * delta, the << 4
//                preshift of dct/idct.cpp:105 is applied by the reconstruction kernels), raster order inside
//                the block; per component a plane [blocks_h][blocks_w] over the MCU-padded grid.
//   samples    : int32, IDCT output (4 fractional bits, level shifted) of the SUBSAMPLED components only,
//                plane [8*blocks_h][8*blocks_w] -- the role of the reference's upsampler line buffers
//                (upsampling/upsamplerbase.cpp:300-327), whole-frame instead of a sliding window.
//   out        : interleaved 8-bit pixels, caller-provided.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "b200jpg.h"

namespace b200jpg {

// ---- host-side parse results -------------------------------------------------------------------------
struct HuffSpec {  // DHT payload, coding/huffmantemplate.cpp:878-904
    bool defined = false;
    uint8_t bits[16] = {0};
    uint8_t vals[256] = {0};
    int nvals = 0;
};

struct ScanInfo {  // one SOS + its entropy coded segment
    int ns = 0;
    int comp[4] = {0, 0, 0, 0};  // frame component index, SOS order
    int td[4] = {0, 0, 0, 0}, ta[4] = {0, 0, 0, 0};
    int lowbit = 0;              // point transform Al
    // progressive frames (SOF2): spectral selection Ss..Se and the bit position Ah of the previous pass of this band
    // (0: first pass); sequential scans have 0, 63, 0
    bool progressive = false;
    int ss = 0, se = 63, ah = 0;
    uint32_t dri = 0;            // MCUs per restart interval, 0 = none
    size_t ecs_off = 0, ecs_end = 0;
    uint32_t mcu_cols = 0, mcu_rows = 0;
    // byte offset (within the codestream) of the first ECS byte of every restart interval;
    // SIZE_MAX marks an interval the stream does not contain (zero-filled like an invalid segment,
    // codestream/sequentialscan.cpp:415-419)
    std::vector<size_t> interval_off;
    std::vector<size_t> interval_end;  // offset of the marker (or end of data) that terminates the interval
    // true: the host did not walk the entropy coded segment; ecs_end is the offset of the closing EOI and the restart
    // index above is a placeholder that restart_index_kernel fills in on the device
    bool device_index = false;
    // the data ends inside this scan's last restart interval (no marker behind it): the decoder reads zero bits there and
    // must not report the overrun (the reference does not, io/bitstream.cpp:103-105)
    bool eof_tail = false;
    // a sequential scan without restart markers that is long enough to be worth cutting up: decoded through synchronisation
    // points (specsync.hpp) instead of by a single lane
    bool spec = false;
    HuffSpec dc[4], ac[4];       // tables in effect at this SOS
    uint16_t quant[4][64];       // zig-zag order as transmitted, in effect at this SOS
    bool quant_defined[4] = {false, false, false, false};
};

// JPEG XT (ISO/IEC 18477) residual layer of a frame, SURVEY 8f3: what the APP11 boxes in front of the first scan say
// (boxes/box.cpp:95-205, boxes/mergingspecbox.cpp, boxes/outputconversionbox.cpp:92-127, boxes/colortrafobox.cpp:55-78).
// Covered: the 8-bit integer profile the reference encoder writes with `-r -q -Q` -- an ordinary DCT codestream in the RESI
// box, merged per pixel (YCbCrTrafo::YCbCr2RGB colortrafo/ycbcrtrafo.cpp:747-880) with default (identity) tables, clamped
// output.  Anything else that is present makes the frame NOT_IMPLEMENTED; it is never decoded as if the boxes were absent.
struct XtLayer {
    bool present = false;      // a merging specification and a residual codestream were found and are covered
    bool l_ycbcr = false;      // base image: YCbCr -> RGB before the merge
    bool r_ycbcr = false;      // residual image: YCbCr -> RGB before the merge
    std::vector<uint8_t> resi;  // the residual codestream, reassembled from its APP11 pieces
};

struct ParsedFrame {
    b200jpg_frame_info info{};
    std::vector<ScanInfo> scans;
    XtLayer xt;
};

// Returns 0 or a negative reference error code; `err` receives a message. With `device_index` a scan that carries all
// components of the frame and uses restart markers is not walked on the host (ScanInfo::device_index).
int parse_codestream(const uint8_t *data, size_t len, ParsedFrame &out, std::string &err, bool device_index = false);

// Restart bookkeeping with the reference's resynchronisation (codestream/entropyparser.cpp:117-199) over the marker
// sequence of one entropy coded segment; off / end are sized to the number of restart intervals (parse.cpp).
int resolve_restart_sequence(const std::vector<size_t> &rst_at, const std::vector<uint8_t> &rst_id, size_t ecs_off, size_t ecs_end,
                             bool ended_by_eof, std::vector<size_t> &off, std::vector<size_t> &end);

extern const uint8_t kZigZagToRaster[64];  // dct/dct.cpp:57-73

// ---- table sets: what the entropy kernel needs besides the bytes ------------------------------------
// Serialised, device independent (this blob is what rank 0 broadcasts over NCCL):
//   uint32 magic, uint32 total_bytes, uint32 lut_words, uint32 flags
//   uint16 lut_off[8]        word offset of the first-level LUT of DC0..3, AC0..3 (0xFFFF = undefined)
//   uint32 qz[4][160][2]     per quantisation table, index = zig-zag position k: {delta << lowbit, byte offset of the
//                            raster position inside a block}; entries 64..95 multiply by 2^16 (a coefficient whose run
//                            leaves the block overflows the int16 store: out-of-sync error; a ZRL that leaves it carries
//                            amplitude 0 and just ends the block), entries 96..159 are where "the block ends" symbols
//                            land (no error); both kinds point at the staging block's pad slot
//   uint32 lut[lut_words]    per table: 2^kLutL1Bits first-level entries, then 2^(16-kLutL1Bits) per second-level table
// LUT entry: [4:0] s = value bits that follow the code, [9:5] code length (0: pointer to second-level table
// [17:10], all other fields 0; 31: unused code, coding/huffmandecoder.hpp:87), [25:19] step of the zig-zag index
// (AC: run + 1, ZRL 16, EOB and error entries kQzBlockEnds), [30:26] total = length + s, [31] decoding this entry is
// an error (unused code, DC category > 15, AC symbol that baseline does not define).
constexpr uint32_t kTableMagic = 0x4a54424du;  // "MBTJ"
constexpr int kLutL1Bits = 11;  // 2 KB entries per table; at q75 fewer than 0.4 % of the AC codes are longer
constexpr int kQzEntries = 160;    // per quantisation table, see above
constexpr int kQzBlockEnds = 96;   // zig-zag step of symbols that end the block: lands in the entries 96..159
constexpr int kTableHeaderBytes = 16 + 16 + 4 * kQzEntries * 2 * 4;

struct TableSet {
    std::vector<uint8_t> blob;
    uint32_t lut_words() const;
};

// Builds the blob for one scan (two-level decoder tables, coding/huffmantemplate.cpp:802-874).
int build_table_set(const ScanInfo &scan, TableSet &out, std::string &err);

// ---- device-side descriptors -------------------------------------------------------------------------
struct ClassScan {        // one scan of one frame inside a scan class
    uint64_t coef_base[4];  // element (int16) offset of the plane of scan component c
    uint32_t frame;         // index into the batch
    uint32_t pad;
};

// One scan whose restart index is built on the device. Offsets are bytes inside the batch's input buffer; the three
// arrays are this scan's slices of its class's interval arrays.
struct IndexScan {
    uint64_t ecs_off, ecs_end;                 // first entropy coded byte, offset of the closing EOI
    uint64_t off_arr, end_arr, clean_arr;      // uint64[n_intervals] each
    uint64_t clean_base;                       // start of the scan's region in the unstuffed buffer (16-byte aligned)
    uint32_t n_intervals, frame;
};
constexpr uint32_t kCleanSlackPerInterval = 80;  // device-indexed scans: clean_off[k] = base + (off[k] - ecs_off) + 80 k, 16-aligned
int launch_restart_index(const IndexScan *scans_dev, uint32_t n_scans, uint8_t *input_dev, uint32_t *index_status, void *stream);

// bit 63 of an interval_end entry / bit 31 of an interval_len entry: the data (not a marker) ends this interval -- the
// decoder reads zero bits behind it like the reference's bit reader at EOF and does not report the overrun
constexpr uint64_t kIntervalEofFlag = 1ull << 63;
constexpr uint32_t kIntervalLenEofFlag = 1u << 31;
// bit 30 of an interval_len entry: the stream does not contain this interval (its MCUs stay cleared). An interval that is
// present but EMPTY (two restart markers back to back) is decoded -- from zero bits, like the reference does
constexpr uint32_t kIntervalLenAbsent = 1u << 30;
constexpr uint32_t kIntervalLenMask = (1u << 30) - 1u;

struct SpecSegment;
struct SpecLog;

struct ScanClassParams {  // uniform over a launch of the entropy kernel
    int ns;
    int mw[4], mh[4];       // blocks per MCU of scan component c (1,1 for single component scans)
    int bw[4];              // plane pitch in blocks of scan component c
    int dc_slot[4], ac_slot[4], q_slot[4];
    uint32_t mcu_cols, total_mcus, dri /* MCUs per interval, >= 1 */, intervals_per_scan;
    uint32_t n_scans;       // scans in this class
    uint32_t lut_words;
    // progressive scans (SOF2): handled by progressive_scan_kernel instead of entropy_decode_kernel
    int progressive, ss, se, ah, al;
    int ordinal;            // position of the scan inside its frame: progressive classes are launched in this order
    // restart-less sequential scans decoded through synchronisation points (specsync.hpp): the scan is ONE interval to the
    // unstuffing kernel and segs_per_scan work items (SpecSegment) to the decoder
    int indexed;
    uint32_t segs_per_scan;
};

// One progressive frame for the dequantisation pass that follows its last scan: the progressive kernels keep
// quantised levels in the coefficient store, stage b expects dequantised coefficients.
struct ProgFrame {
    uint64_t coef_base[4];   // int16 element offsets of the component planes
    uint32_t n_blocks[4];    // blocks per plane (MCU-padded grid), 0 = component absent
    uint16_t q_raster[4][64];  // quantiser of the component in raster order
    uint32_t frame, pad;
};

// Component-fused progressive decoding (progfused_sm100.cu): one launch runs a restart interval through ALL scans of a group
// -- the interleaved DC scans of a frame, or the AC scans of one component -- in file order.
constexpr int kPfMaxScans = 6;
struct PfScan {                    // one scan of the group, uniform over the launch
    const uint8_t *tables;         // device copy of the scan's table-set blob
    const uint64_t *clean_off;     // [n_frames * intervals] offsets of the unstuffed intervals of this scan
    const uint32_t *interval_len;  // [..] unstuffed lengths (flags: kIntervalLenAbsent / kIntervalLenEofFlag)
    int ss, se, ah, al;
    int dc_slot[4], ac_slot, q_slot;
    uint32_t lut_words;
    int lut_share;                 // AC kernel: the first scan of the launch whose decoder tables are the same words (itself if none)
};
struct PfLaunch {
    int n_scans;
    PfScan scan[kPfMaxScans];
    int ns;                        // components in the group's scans (DC group: all of the frame; AC group: 1)
    int mw[4], mh[4], bw[4];       // blocks per MCU and plane pitch of scan component c
    uint32_t ac_cols[4], ac_rows[4];  // DC group: the block grid the component's own (single-component) scans cover
    uint16_t dc_quant[4];          // quantiser of coefficient 0 of scan component c
    uint32_t mcu_cols, total_mcus, dri, intervals, n_frames;
    const ClassScan *frames;       // [n_frames] coefficient plane of every scan component + frame index
    const uint8_t *clean;
    int16_t *coef;
    int16_t *dcplane;              // one quantised DC level per block, indexed like the coefficient store / 64
    uint32_t *frame_status;
};
int launch_pf_dc(const PfLaunch &L, void *stream);
int launch_pf_ac(const PfLaunch &L, void *stream);

struct FrameRecon {       // per frame, for the reconstruction kernels
    uint64_t coef_base[4];    // int16 element offsets
    uint64_t sample_base[4];  // element offsets into the sample planes (subsampled components only)
    uint64_t out_base;        // byte offset into the output buffer
    uint32_t width, height;
    uint32_t bw[4], bh[4];
    uint32_t ncomp, ycbcr, subx, suby;  // subx/suby of the subsampled (chroma) components
    uint32_t cw, ch;                    // true subsampled size ceil(W/subx), ceil(H/suby)
    uint32_t status_idx, precision;     // sample precision of the frame (8, or 12: generic reconstruction, 16-bit samples out)
    uint8_t csx[4], csy[4];             // generic reconstruction: subsampling factors of every component (1..4)
    // JPEG XT (generic reconstruction): bit 0 = merge with the residual image whose int32 sample planes start at
    // res_sample_base (block grid width res_bw, factors res_csx / res_csy), bit 1 / 2 = YCbCr -> RGB on the base / residual side
    uint32_t xt, xt_pad;
    uint64_t res_sample_base[4];
    uint32_t res_bw[4];
    uint8_t res_csx[4], res_csy[4];
};

// kernel launchers (huffman_sm100.cu, recon_sm100.cu). All asynchronous on `stream`.
struct EntropyLaunch {
    ScanClassParams p;
    const uint8_t *bytes;           // packed codestreams
    const uint64_t *interval_off;   // [n_scans * intervals_per_scan] first ECS byte, ~0ull = absent
    const uint64_t *interval_end;   // [..] offset of the marker that ends the interval
    const uint64_t *clean_off;      // [..] 16-byte aligned offset of the interval's unstuffed bytes in `clean`
    uint8_t *clean;                 // unstuffed, big-endian-word entropy coded data (written by a0, read by a1)
    uint32_t *interval_len;         // [..] unstuffed length in bytes (written by a0)
    const ClassScan *scans;         // [n_scans]
    const uint8_t *tables;          // device copy of the table-set blob
    int16_t *coef;
    uint32_t *frame_status;         // [n_frames]
    // sequential scans: {count, interval indices ...} of the intervals whose decoder read past the marker that ends them;
    // overrun_verdict_kernel replays exactly those with the reference's bit-reader bookkeeping to decide whether the
    // reference would have thrown (io/bitstream.hpp:168-208, bitstream.cpp:56-118)
    uint32_t *overrun_list;
    // indexed classes (p.indexed): work items written by spec_sync_kernel and its scratch, [n_scans * segs_per_scan] each
    struct SpecSegment *spec_segments;
    unsigned long long *spec_exits, *spec_entries;
    uint32_t *spec_counts;
    int32_t *spec_dc_sums;  // [.. * 4]
    struct SpecLog *spec_logs;
};
int launch_unstuff(const EntropyLaunch &l, void *stream);
int launch_entropy(const EntropyLaunch &l, void *stream);
int launch_overrun_verdict(const EntropyLaunch &l, void *stream);
int launch_spec_sync(const EntropyLaunch &l, void *stream);  // specsync_sm100.cu
// progressive_sm100.cu: one scan class of progressive frames; quantised levels -> dequantised coefficients afterwards
int launch_progressive_scan(const EntropyLaunch &l, void *stream);
int launch_progressive_dequant(const ProgFrame *frames_dev, uint32_t n_frames, uint32_t max_blocks, int16_t *coef, uint32_t *frame_status,
                               void *stream);

struct ReconLaunch {
    const FrameRecon *frames;  // device
    uint32_t n_frames;
    uint32_t max_bw0, max_bh0;     // largest luma block grid in the group
    uint32_t max_bwc, max_bhc;     // largest chroma block grid in the group
    uint32_t ncomp, subx, suby;    // uniform over the group
    bool generic;                  // formats outside the tuned kernels: any component count / factors, per-frame parameters
    bool planes_out;               // B200JPG_FLAG_NO_UPSAMPLE: the components leave as planes at their own resolution
    int generic_phase;             // generic groups: 0 = IDCT then reconstruction, 1 = IDCT only, 2 = reconstruction only
    const int16_t *coef;
    int16_t *samples16;            // chroma sample planes every frame goes through
    int32_t *samples32;            // the same planes for the exact pass over frames flagged `narrow`
    uint32_t *narrow_flags;        // [frames in batch], zeroed per decode: a chroma sample does not fit int16
    uint32_t *wide_flags;          // [frames in batch], zeroed per decode: a chroma sample leaves the 32-bit colour range
    uint32_t *narrow_list;         // [1 + frames in group] scratch: {count, group-local indices of the flagged frames}
    uint8_t *out;
};
int launch_recon(const ReconLaunch &l, void *stream, int *launches);

}  // namespace b200jpg

// huffman_sm100.cu -- stage (a): sequential-scan Huffman decode + dequantisation on sm_100a.
//
// Replaces the reference's entropy path for SOF0/SOF1 Huffman scans:
//   SequentialScan::ParseMCU / DecodeBlock / Restart   codestream/sequentialscan.cpp:381-428, 678-773, 266-274
//   EntropyParser::BeginReadMCU                        codestream/entropyparser.hpp:147-160
//   HuffmanDecoder::Get (two-level 8+8 bit lookup)     coding/huffmandecoder.hpp:103-124
//   BitStream<false>::Fill / Get / PeekWord / SkipBits io/bitstream.cpp:56-118, io/bitstream.hpp:168-208
//   dequantisation multiplier                          dct/idct.cpp:98-108 (the << 4 is left to stage b)
//
// Mapping.  The restart interval is the unit of work (each one restarts the bit reader byte-aligned and
// resets the DC predictors, so intervals are independent).  A warp decodes 32 restart intervals, one per
// lane, in lock step block by block: every lane owns its private bit window (64-bit register pair) fed from a
// 64-byte shared-memory ring that cp.async (LDGSTS, 16 bytes at a time, three chunks ahead of the reader)
// keeps filled, so HBM latency never sits on the decode chain; the Huffman tables of the scan live in shared
// memory as combined (total bits | code length | symbol) entries, and each lane
// scatters its coefficients de-zigzagged and dequantised into a private 128-byte shared-memory block that is
// flushed to HBM as eight 16-byte vector stores -- explicit zeros included, so the coefficient store needs
// no memset and every 128-byte block line is written exactly once.  Warp votes keep the per-symbol loop
// convergent.  All intervals of all frames that share scan geometry and tables form one launch.
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.hpp"

namespace b200jpg {
namespace {

constexpr int kThreads = 128;
constexpr int kStageStride = 72;  // int16 per lane: 64 + 8 pad -> 144-byte stride, conflict-free 16-byte accesses
constexpr unsigned kFull = 0xffffffffu;

// error codes written to frame_status (reference's numeric values, interface/parameters.hpp:1156-1228)
constexpr uint32_t kErrMalformed = 1038u;      // -(-1038)
constexpr uint32_t kErrUnexpectedEof = 1025u;  // -(-1025)

struct BitWindow {
    const uint8_t *base;
    uint32_t *ring;     // this lane's 16-word (64-byte) ring in shared memory: word (pos >> 2) & 15
    uint64_t pos;       // next unread byte
    uint64_t w;         // MSB-aligned window
    uint32_t chunk;     // 16-byte chunk index the reader is in; chunks chunk .. chunk+3 are requested
    int n;              // bits in w (real + virtual)
    int vbits;          // virtual zero bits appended after a marker was met (io/bitstream.cpp:96-101)
    bool stopped;
};

__device__ __forceinline__ void cp_async16(uint32_t *smem_dst, const uint8_t *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ring_request(BitWindow &b, uint32_t chunk) {
    cp_async16(b.ring + ((chunk & 3u) << 2), b.base + ((uint64_t)chunk << 4));
    cp_async_commit();
}

// Opens the window at byte offset `off`: four chunks in flight, the first two landed.
__device__ __forceinline__ void open_window(BitWindow &b, uint64_t off) {
    b.pos = off;
    b.chunk = (uint32_t)(off >> 4);
    b.w = 0;
    b.n = 0;
    b.vbits = 0;
    b.stopped = false;
#pragma unroll
    for (uint32_t i = 0; i < 4; i++) ring_request(b, b.chunk + i);
    cp_async_wait<2>();
}

__device__ __forceinline__ uint32_t ring_byte(const BitWindow &b, uint64_t pos) {
    return (b.ring[(uint32_t)(pos >> 2) & 15u] >> ((uint32_t)(pos & 3) * 8u)) & 0xffu;
}

// Guarantees n >= 32 (a Huffman code of <= 16 bits plus <= 15 value bits always fits).
__device__ __forceinline__ void refill(BitWindow &b) {
    if (b.n > 32) return;
    if (!b.stopped) {
        uint32_t c = (uint32_t)(b.pos >> 4);
        if (c != b.chunk) {  // entered the next chunk: request the one three ahead, chunks c and c+1 must have landed
            b.chunk = c;
            ring_request(b, c + 3);
            cp_async_wait<2>();
        }
        uint32_t wq = (uint32_t)(b.pos >> 2);
        uint32_t lo = b.ring[wq & 15u], hi = b.ring[(wq + 1) & 15u];
        uint32_t raw = __funnelshift_r(lo, hi, (uint32_t)(b.pos & 3) * 8u);  // bytes pos..pos+3, little endian
        uint32_t ff = ((~raw) - 0x01010101u) & raw & 0x80808080u;             // any byte == 0xFF ?
        if (ff == 0) {
            uint32_t be = __byte_perm(raw, 0, 0x0123);
            b.w |= (uint64_t)be << (32 - b.n);
            b.n += 32;
            b.pos += 4;
            return;
        }
        // rare: a 0xFF among the next four bytes -> byte stuffing or a marker (io/bitstream.cpp:63-101)
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            uint32_t v = ring_byte(b, b.pos);
            if (v == 0xffu) {
                if (ring_byte(b, b.pos + 1) != 0u) {
                    b.stopped = true;  // marker: stay in front of it, feed zeros from now on
                    break;
                }
                b.pos += 2;
            } else {
                b.pos += 1;
            }
            b.w |= (uint64_t)v << (56 - b.n);
            b.n += 8;
        }
        if (!b.stopped || b.n > 32) return;
    }
    // in front of a marker: the reference appends zero bits (io/bitstream.cpp:96-101); count them so the
    // end-of-interval check can tell whether any of them was actually consumed
    b.n += 32;
    b.vbits += 32;
}

__device__ __forceinline__ void consume(BitWindow &b, int bits) {
    b.w <<= bits;
    b.n -= bits;
}

template <bool kLutShared>
__device__ __forceinline__ uint32_t lut_at(const uint32_t *lut, uint32_t idx) {
    if (kLutShared) return lut[idx];
    return __ldg(lut + idx);
}

// Huffman symbol at the head of the window: returns (total bits << 16) | (len << 8) | symbol, len == 0xff for
// an unused code; total = code length + value bits that follow.
template <bool kLutShared>
__device__ __forceinline__ uint32_t huff_peek(const uint32_t *lut, uint32_t off, const BitWindow &b) {
    uint32_t peek = (uint32_t)(b.w >> 48);
    uint32_t e = lut_at<kLutShared>(lut, off + (peek >> 8));
    if ((e & 0xff00u) == 0) e = lut_at<kLutShared>(lut, off + 256u * (e & 0xffu) + (peek & 0xffu));
    return e;
}

// `s` value bits that follow a code of `len` bits, sign-extended as in sequentialscan.cpp:692-696 / 757-762
__device__ __forceinline__ int value_bits(const BitWindow &b, int len, int s) {
    uint32_t hi = (uint32_t)(b.w >> 32);
    uint32_t v = ((hi << len) >> 1) >> (31 - s);  // s == 0 -> 0
    uint32_t thresh = (1u << s) >> 1;
    return (int)v - ((v < thresh) ? (int)((1u << s) - 1u) : 0);
}

template <bool kLutShared>
__global__ void __launch_bounds__(kThreads)
entropy_decode_kernel(ScanClassParams p, const uint8_t *__restrict__ bytes, const uint64_t *__restrict__ interval_off,
                      const ClassScan *__restrict__ scans, const uint8_t *__restrict__ tables, int16_t *__restrict__ coef,
                      uint32_t *__restrict__ frame_status) {
    extern __shared__ __align__(16) uint8_t smem[];
    // layout: [stage: kThreads * kStageStride int16][ring: kThreads * 16 uint32][qz: 4*64 uint32][lut: lut_words uint32 (if shared)]
    int16_t *stage_all = reinterpret_cast<int16_t *>(smem);
    uint32_t *ring_all = reinterpret_cast<uint32_t *>(smem + kThreads * kStageStride * 2);
    uint32_t *qz = ring_all + kThreads * 16;
    uint32_t *lut_s = qz + 4 * 64;

    const uint32_t *g_qz = reinterpret_cast<const uint32_t *>(tables + 32);
    const uint32_t *g_lut = reinterpret_cast<const uint32_t *>(tables + kTableHeaderBytes);
    for (int i = threadIdx.x; i < 4 * 64; i += kThreads) qz[i] = g_qz[i];
    if (kLutShared) {
        for (uint32_t i = threadIdx.x; i < p.lut_words; i += kThreads) lut_s[i] = g_lut[i];
    }
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *st = reinterpret_cast<uint4 *>(stage_all + threadIdx.x * kStageStride);
#pragma unroll
        for (int i = 0; i < 8; i++) st[i] = z;
    }
    __syncthreads();
    const uint32_t *lut = kLutShared ? lut_s : g_lut;
    const uint16_t *lut_off = reinterpret_cast<const uint16_t *>(tables + 16);

    int16_t *stage = stage_all + threadIdx.x * kStageStride;
    const uint64_t total_intervals = (uint64_t)p.n_scans * p.intervals_per_scan;
    const uint64_t g = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool lane_valid = g < total_intervals;

    // per-lane interval
    uint32_t j = 0, iv = 0;
    if (lane_valid) {
        j = (uint32_t)(g / p.intervals_per_scan);
        iv = (uint32_t)(g % p.intervals_per_scan);
    }
    uint64_t off = lane_valid ? interval_off[g] : ~0ull;
    const uint32_t mcu0 = iv * p.dri;
    uint32_t nmcu = 0;
    if (lane_valid) nmcu = (p.total_mcus - mcu0 < p.dri) ? (p.total_mcus - mcu0) : p.dri;
    uint32_t mx = mcu0 % p.mcu_cols, my = mcu0 / p.mcu_cols;

    uint64_t plane[4];
    uint32_t frame = 0;
    {
        const ClassScan &cs = scans[lane_valid ? j : 0];
#pragma unroll
        for (int c = 0; c < 4; c++) plane[c] = cs.coef_base[c];
        frame = cs.frame;
    }

    BitWindow b;
    b.base = bytes;
    b.ring = ring_all + threadIdx.x * 16;
    b.pos = 0;
    b.chunk = 0;
    b.w = 0;
    b.n = 0;
    b.vbits = 0;
    b.stopped = false;
    if (lane_valid && off != ~0ull) open_window(b, off);
    bool decoding = lane_valid && off != ~0ull;  // absent interval: blocks stay zero (sequentialscan.cpp:415-419)
    uint32_t err = 0;
    int pred[4] = {0, 0, 0, 0};

    // table offsets (uniform)
    uint32_t dc_off[4], ac_off[4], q_off[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        dc_off[c] = (c < p.ns) ? lut_off[p.dc_slot[c]] : 0;
        ac_off[c] = (c < p.ns) ? lut_off[4 + p.ac_slot[c]] : 0;
        q_off[c] = (c < p.ns) ? 64u * p.q_slot[c] : 0;
    }

    for (uint32_t mi = 0; mi < p.dri; mi++) {
        const bool has_mcu = mi < nmcu;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c >= p.ns) break;
            for (int y = 0; y < p.mh[c]; y++) {
                for (int x = 0; x < p.mw[c]; x++) {
                    bool busy = has_mcu && decoding;
                    int k = 1;
                    // ---- DC: sequentialscan.cpp:682-701
                    if (busy) {
                        refill(b);
                        uint32_t e = huff_peek<kLutShared>(lut, dc_off[c], b);
                        int len = (int)((e >> 8) & 0xffu), s = (int)(e & 0xffu);
                        if (len > 16 || s > 15) {
                            err = kErrMalformed;
                            busy = false;
                            decoding = false;
                        } else {
                            int diff = value_bits(b, len, s);
                            consume(b, (int)(e >> 16));
                            pred[c] += diff;
                            int v = pred[c] * (int)(qz[q_off[c]] >> 8);
                            if (v != (int)(int16_t)v) err = kErrMalformed;  // does not fit the int16 store
                            stage[0] = (int16_t)v;
                        }
                    }
                    // ---- AC: sequentialscan.cpp:704-771, one symbol per warp-convergent iteration
                    while (__any_sync(kFull, busy)) {
                        if (busy) {
                            refill(b);
                            uint32_t e = huff_peek<kLutShared>(lut, ac_off[c], b);
                            int len = (int)((e >> 8) & 0xffu), rs = (int)(e & 0xffu);
                            int r = rs >> 4, s = rs & 15;
                            if (len > 16) {
                                err = kErrMalformed;
                                busy = false;
                                decoding = false;
                            } else if (s == 0) {
                                consume(b, len);
                                if (r == 15) {
                                    k += 16;  // ZRL; the reference re-tests k <= 63 and silently ends the block
                                    busy = (k <= 63);
                                } else if (r == 0) {
                                    busy = false;  // EOB
                                } else {
                                    err = kErrMalformed;  // sequentialscan.cpp:750-752
                                    busy = false;
                                    decoding = false;
                                }
                            } else {
                                k += r;
                                int diff = value_bits(b, len, s);
                                consume(b, (int)(e >> 16));
                                if (k >= 64) {
                                    err = kErrMalformed;  // :764-766
                                    busy = false;
                                    decoding = false;
                                } else {
                                    uint32_t qe = qz[q_off[c] + k];
                                    int v = diff * (int)(qe >> 8);
                                    if (v != (int)(int16_t)v) err = kErrMalformed;
                                    stage[qe & 0xffu] = (int16_t)v;
                                    k++;
                                    busy = (k <= 63);
                                }
                            }
                        }
                    }
                    // ---- flush the block (zeros included) and clear the staging block
                    if (has_mcu) {
                        uint32_t bx = mx * p.mw[c] + x, by = my * p.mh[c] + y;
                        uint4 *dst = reinterpret_cast<uint4 *>(coef + plane[c] + ((uint64_t)by * p.bw[c] + bx) * 64u);
                        uint4 *st = reinterpret_cast<uint4 *>(stage);
                        uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            uint4 v = st[i];
                            st[i] = z;
                            dst[i] = v;
                        }
                    }
                }
            }
        }
        if (++mx == p.mcu_cols) {
            mx = 0;
            my++;
        }
    }
    // a valid stream never consumes bits beyond the marker that ends its interval
    if (decoding && err == 0 && b.vbits > b.n) err = kErrUnexpectedEof;
    if (err) atomicMax(frame_status + frame, err);
}

}  // namespace

int launch_entropy(const EntropyLaunch &l, void *stream) {
    const uint64_t total = (uint64_t)l.p.n_scans * l.p.intervals_per_scan;
    if (total == 0) return 0;
    const uint32_t grid = (uint32_t)((total + kThreads - 1) / kThreads);
    size_t base_smem = (size_t)kThreads * kStageStride * 2 + (size_t)kThreads * 64 + 4 * 64 * 4;
    size_t lut_bytes = (size_t)l.p.lut_words * 4;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (base_smem + lut_bytes <= 160 * 1024) {
        size_t smem = base_smem + lut_bytes;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(entropy_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        entropy_decode_kernel<true><<<grid, kThreads, smem, s>>>(l.p, l.bytes, l.interval_off, l.scans, l.tables, l.coef, l.frame_status);
    } else {
        entropy_decode_kernel<false><<<grid, kThreads, base_smem, s>>>(l.p, l.bytes, l.interval_off, l.scans, l.tables, l.coef, l.frame_status);
    }
    e = cudaGetLastError();
    return (int)e;
}

}  // namespace b200jpg

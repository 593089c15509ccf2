// recon_sm100.cu -- stage (b): integer IDCT + centred-bilinear chroma upsampling + YCbCr->RGB + 8-bit store.
//
// Replaces the reference's reconstruction path (bit-exact, all int32 like the reference's LONG):
//   BlockBitmapRequester::ReconstructRegion / PullQData / PushReconstructedData / ReconstructUnsampled
//                                                     control/blockbitmaprequester.cpp:1249-1272,1079-1112,1151-1224,1013-1074
//   IDCT<4,LONG,false,false>::InverseTransformBlock   dct/idct.cpp:226-339 (constants idct.cpp:65-78, idct.hpp:70-77)
//   UpsamplerBase::DefineRegion (edge replication)    upsampling/upsamplerbase.cpp:300-327
//   Upsampler<sx,sy>::UpsampleRegion                  upsampling/upsampler.cpp:83-112
//   VerticalFilterCore<1|2>, HorizontalFilterCore<1|2> upsampling/upsampler.cpp:114-168, 270-307 (incl. the in-place
//                                                     read-after-write of out[1], :301-302, which is part of the contract)
//   YCbCrTrafo<UBYTE,3,ClampFlag,YCbCr,Zero>::YCbCr2RGB colortrafo/ycbcrtrafo.cpp:679-1008 (:842-850, :922-935)
//
// Mapping.  One thread owns one 8x8 block.
//  b1 idct_planes_kernel: blocks of the non-luma components -> int32 sample planes (the whole-frame equivalent of
//     the reference's upsampler line buffers); both 1-D passes in registers.
//  b2 reconstruct_kernel: a warp owns 32 horizontally adjacent luma blocks (256 x 8 pixels), a CTA four such rows.
//     The warp's 4 KB of coefficients come in with coalesced 16-byte loads and are handed to their lanes through
//     shared memory; the row pass and the column pass of the luma IDCT run as two compact loops over a shared-memory
//     tile laid out [coefficient][thread] (bank-conflict free, thread-private columns); a third loop walks the eight
//     output lines: chroma window from the planes (clamped addressing = the reference's edge replication at the true
//     subsampled size), vertical + horizontal filter cores, colour transform, and each lane stores its 24 bytes of
//     interleaved RGB per line.  The kernel is co-limited by instruction issue and by the L1 data pipe, so global
//     accesses are shaped for few L1 wavefronts; small loop bodies keep it inside the instruction cache (the fully
//     unrolled register-resident version stalled on instruction fetch).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "internal.hpp"

namespace b200jpg {
namespace {

constexpr int kThreadsB = 128;
constexpr int kWide = 65535;  // |sample| above this may overflow the 32-bit colour arithmetic -> 64-bit path

#define WMUL(a, k) ((int)((unsigned)(a) * (unsigned)(int)(k)))
#define WADD(a, b) ((int)((unsigned)(a) + (unsigned)(b)))
#define WSUB(a, b) ((int)((unsigned)(a) - (unsigned)(b)))

// One 8-point pass of dct/idct.cpp:237-287 (rows, round = 256, shift = 9) / :291-334 (columns, 2048, 12).
// Constants are WORD(x * 512 + 0.5) of the reference's TO_FIX table.
template <int kRound, int kShift>
__device__ __forceinline__ void idct8(int &v0, int &v1, int &v2, int &v3, int &v4, int &v5, int &v6, int &v7) {
    int z1 = WMUL(WADD(v2, v6), 277);
    int tmp2 = WADD(z1, WMUL(v6, -946));
    int tmp3 = WADD(z1, WMUL(v2, 392));
    int tmp0 = (int)((unsigned)WADD(v0, v4) << 9);
    int tmp1 = (int)((unsigned)WSUB(v0, v4) << 9);
    int tmp10 = WADD(tmp0, tmp3), tmp13 = WSUB(tmp0, tmp3);
    int tmp11 = WADD(tmp1, tmp2), tmp12 = WSUB(tmp1, tmp2);
    int t0 = v7, t1 = v5, t2 = v3, t3 = v1;
    int y1 = WADD(t0, t3), y2 = WADD(t1, t2), y3 = WADD(t0, t2), y4 = WADD(t1, t3);
    int z5 = WMUL(WADD(y3, y4), 602);
    t0 = WMUL(t0, 153);
    t1 = WMUL(t1, 1051);
    t2 = WMUL(t2, 1573);
    t3 = WMUL(t3, 769);
    y1 = WMUL(y1, -461);
    y2 = WMUL(y2, -1312);
    y3 = WADD(WMUL(y3, -1004), z5);
    y4 = WADD(WMUL(y4, -200), z5);
    t0 = WADD(t0, WADD(y1, y3));
    t1 = WADD(t1, WADD(y2, y4));
    t2 = WADD(t2, WADD(y2, y3));
    t3 = WADD(t3, WADD(y1, y4));
    v0 = WADD(WADD(tmp10, t3), kRound) >> kShift;
    v7 = WADD(WSUB(tmp10, t3), kRound) >> kShift;
    v1 = WADD(WADD(tmp11, t2), kRound) >> kShift;
    v6 = WADD(WSUB(tmp11, t2), kRound) >> kShift;
    v2 = WADD(WADD(tmp12, t1), kRound) >> kShift;
    v5 = WADD(WSUB(tmp12, t1), kRound) >> kShift;
    v3 = WADD(WADD(tmp13, t0), kRound) >> kShift;
    v4 = WADD(WSUB(tmp13, t0), kRound) >> kShift;
}

// eight dequantised int16 coefficients (one block row) -> ints carrying the << 4 preshift of dct/idct.cpp:105
__device__ __forceinline__ void unpack_row(const uint4 q, int (&v)[8]) {
    v[0] = (int)(short)(q.x & 0xffffu) << 4;
    v[1] = (int)(short)(q.x >> 16) << 4;
    v[2] = (int)(short)(q.y & 0xffffu) << 4;
    v[3] = (int)(short)(q.y >> 16) << 4;
    v[4] = (int)(short)(q.z & 0xffffu) << 4;
    v[5] = (int)(short)(q.z >> 16) << 4;
    v[6] = (int)(short)(q.w & 0xffffu) << 4;
    v[7] = (int)(short)(q.w >> 16) << 4;
}

// ---- instruction-selection helpers ---------------------------------------------------------------------------------------
// Both integer pipes of an SM sub-partition (ALU: IADD3 / SHF / LOP3 / LEA / PRMT; FMA-heavy: IMAD, IMAD.IADD, IMAD.MOV, VIADD)
// take one warp instruction every other cycle (tools/opbench.cu, profiles/), so the kernel runs at the pace of the fuller one
// -- r01: FMA-heavy 79 % busy, ALU 57 %. ptxas likes "x * 3 + y" as IMAD followed by VIADD for the rounding constant: two
// FMA-heavy instructions per filter tap. Spelled as below it emits IADD3 (a + b + r) and one IMAD (b * 2 + t): one on each pipe.
#ifndef B200JPG_FUSED_PLAIN
__device__ __forceinline__ int add3(int a, int b, int c) {
    int d;
    asm("{ .reg .s32 x; add.s32 x, %1, %2; add.s32 %0, x, %3; }" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <int C>
__device__ __forceinline__ int add2c(int a, int b) {  // a + b + C
    int d;
    asm("{ .reg .s32 x; add.s32 x, %1, %2; add.s32 %0, x, %3; }" : "=r"(d) : "r"(a), "r"(b), "n"(C));
    return d;
}
template <int C>
__device__ __forceinline__ int sub2c(int a, int b) {  // a - b + C
    int d;
    asm("{ .reg .s32 x; sub.s32 x, %1, %2; add.s32 %0, x, %3; }" : "=r"(d) : "r"(a), "r"(b), "n"(C));
    return d;
}
// (a + 3 b + R) >> 2: upsampling/upsampler.cpp:136-168, 283-307
template <int R>
__device__ __forceinline__ int tap(int a, int b) {
    int d;
    asm("{ .reg .s32 x, y; add.s32 x, %1, %2; add.s32 x, x, %3; shl.b32 y, %2, 1; add.s32 x, x, y; shr.s32 %0, x, 2; }" : "=r"(d) : "r"(a), "r"(b), "n"(R));
    return d;
}
#else
__device__ __forceinline__ int add3(int a, int b, int c) { return WADD(WADD(a, b), c); }
template <int C>
__device__ __forceinline__ int add2c(int a, int b) { return WADD(WADD(a, b), C); }
template <int C>
__device__ __forceinline__ int sub2c(int a, int b) { return WADD(WSUB(a, b), C); }
template <int R>
__device__ __forceinline__ int tap(int a, int b) { return WADD(WADD(a, WMUL(3, b)), R) >> 2; }
#endif

// HorizontalFilterCore<2> (upsampler.cpp:283-307) on w[0..5] with the taps above
__device__ __forceinline__ void hfilter2t(const int (&w)[6], int (&o)[8]) {
    o[7] = tap<1>(w[5], w[4]);
    o[6] = tap<2>(w[3], w[4]);
    o[5] = tap<1>(w[4], w[3]);
    o[4] = tap<2>(w[2], w[3]);
    o[3] = tap<1>(w[3], w[2]);
    o[2] = tap<2>(w[1], w[2]);
    o[1] = tap<1>(o[2], w[1]);  // reads the freshly written out[2] (upsampler.cpp:301-302)
    o[0] = tap<2>(w[0], w[1]);
}

// One 8-point pass like idct8, with the rounding constant riding in the three-input adds of the outputs
template <int kRound, int kShift>
__device__ __forceinline__ void idct8t(int &v0, int &v1, int &v2, int &v3, int &v4, int &v5, int &v6, int &v7) {
    const int z1 = WMUL(WADD(v2, v6), 277);
    const int tmp2 = WADD(z1, WMUL(v6, -946));
    const int tmp3 = WADD(z1, WMUL(v2, 392));
    const int a04 = WADD(v0, v4), s04 = WSUB(v0, v4);
    const int tmp10 = WADD(WMUL(a04, 512), tmp3), tmp13 = WSUB(WMUL(a04, 512), tmp3);
    const int tmp11 = WADD(WMUL(s04, 512), tmp2), tmp12 = WSUB(WMUL(s04, 512), tmp2);
    const int y1 = WADD(v7, v1), y2 = WADD(v5, v3), y3 = WADD(v7, v3), y4 = WADD(v5, v1);
    const int z5 = WMUL(WADD(y3, y4), 602);
    const int p1 = WMUL(y1, -461), p2 = WMUL(y2, -1312);
    const int p3 = WADD(WMUL(y3, -1004), z5), p4 = WADD(WMUL(y4, -200), z5);
    const int t0 = WADD(WADD(WMUL(v7, 153), p1), p3);
    const int t1 = WADD(WADD(WMUL(v5, 1051), p2), p4);
    const int t2 = WADD(WADD(WMUL(v3, 1573), p2), p3);
    const int t3 = WADD(WADD(WMUL(v1, 769), p1), p4);
    v0 = add2c<kRound>(tmp10, t3) >> kShift;
    v7 = sub2c<kRound>(tmp10, t3) >> kShift;
    v1 = add2c<kRound>(tmp11, t2) >> kShift;
    v6 = sub2c<kRound>(tmp11, t2) >> kShift;
    v2 = add2c<kRound>(tmp12, t1) >> kShift;
    v5 = sub2c<kRound>(tmp12, t1) >> kShift;
    v3 = add2c<kRound>(tmp13, t0) >> kShift;
    v4 = sub2c<kRound>(tmp13, t0) >> kShift;
}

// eight dequantised int16 coefficients -> ints WITHOUT the << 4 preshift of dct/idct.cpp:105. The row pass is linear in front
// of its rounding shift, so with inputs 16 times smaller ((16 x + 256) >> 9) == ((x + 16) >> 5) for every integer x; the
// reference's int32 arithmetic cannot wrap in this pass for coefficients that fit int16 (|16 x| <= 16 * 3825 * (32768 + 1024)
// < 2^31), so nothing is lost by never forming 16 x.
__device__ __forceinline__ void unpack_row_raw(const uint4 q, int (&v)[8]) {
    v[0] = (int)(short)(q.x & 0xffffu);
    v[1] = (int)q.x >> 16;
    v[2] = (int)(short)(q.y & 0xffffu);
    v[3] = (int)q.y >> 16;
    v[4] = (int)(short)(q.z & 0xffffu);
    v[5] = (int)q.z >> 16;
    v[6] = (int)(short)(q.w & 0xffffu);
    v[7] = (int)q.w >> 16;
}

// ---- b1: non-luma components -> sample planes ---------------------------------------------------------
// Samples of real images fit 16 bits with room to spare, so the planes that every frame goes through are int16
// (half the HBM traffic of this memory-bound kernel). A frame in which any sample does not fit is flagged `narrow`
// and redone by the same kernels instantiated for int32 planes (kListed: they walk the list of flagged frames
// instead of the whole group, so the exact pass costs three tiny launches when nothing is flagged).
constexpr int kWideSlots = 4;  // grid extent of the exact pass in the frame dimension; each slot strides over the list

template <typename T>
__device__ __forceinline__ void idct_planes_block(const FrameRecon &f, int c, uint32_t t, const int16_t *__restrict__ coef,
                                                  T *__restrict__ samples, uint32_t *__restrict__ flags) {
    const uint32_t bw = f.bw[c], bh = f.bh[c];
    if (t >= bw * bh) return;
    const uint32_t bx = t % bw, by = t / bw;
    int s[8][8];
    const uint4 *src = reinterpret_cast<const uint4 *>(coef + f.coef_base[c] + (uint64_t)t * 64u);
#pragma unroll
    for (int r = 0; r < 8; r++) unpack_row_raw(__ldg(src + r), s[r]);
    // dcoffset << 3 (idct.cpp:233,244); the << 4 preshift is folded into the row pass' rounding. 12-bit frames only come
    // through the int32 planes (tables.cpp:1877-1891: the same LONG IDCT, level shift 1 << 11)
    s[0][0] = WADD(s[0][0], sizeof(T) == 4 ? (int)(8u << (f.precision - 1u)) : (128 << 3));
#pragma unroll
    for (int r = 0; r < 8; r++) idct8t<16, 5>(s[r][0], s[r][1], s[r][2], s[r][3], s[r][4], s[r][5], s[r][6], s[r][7]);
    int mx = 0, mn = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        idct8t<2048, 12>(s[0][k], s[1][k], s[2][k], s[3][k], s[4][k], s[5][k], s[6][k], s[7][k]);
#pragma unroll
        for (int r = 0; r < 8; r += 2) {
            mx = __vimax3_s32(mx, s[r][k], s[r + 1][k]);
            mn = __vimin3_s32(mn, s[r][k], s[r + 1][k]);
        }
    }
    const uint32_t pitch = 8u * bw;
    T *dst = samples + f.sample_base[c] + (uint64_t)(8u * by) * pitch + 8u * bx;
    if (sizeof(T) == 2) {
        if (mx > 32767 || mn < -32768) atomicOr(flags + f.status_idx, 1u);  // `narrow`: the frame needs the int32 planes
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint4 w;
            w.x = ((uint32_t)s[r][0] & 0xffffu) | ((uint32_t)s[r][1] << 16);
            w.y = ((uint32_t)s[r][2] & 0xffffu) | ((uint32_t)s[r][3] << 16);
            w.z = ((uint32_t)s[r][4] & 0xffffu) | ((uint32_t)s[r][5] << 16);
            w.w = ((uint32_t)s[r][6] & 0xffffu) | ((uint32_t)s[r][7] << 16);
            *reinterpret_cast<uint4 *>(dst + (uint64_t)r * pitch) = w;
        }
    } else {
        if (mx > kWide || mn < -kWide) atomicOr(flags + f.status_idx, 1u);  // `wide`: colour needs 64 bits
#pragma unroll
        for (int r = 0; r < 8; r++) {
            int4 *d = reinterpret_cast<int4 *>(dst + (uint64_t)r * pitch);
            d[0] = make_int4(s[r][0], s[r][1], s[r][2], s[r][3]);
            d[1] = make_int4(s[r][4], s[r][5], s[r][6], s[r][7]);
        }
    }
}

// grid (blocks / 128, frames or kWideSlots, components - 1); list = {count, frame indices...} for kListed
template <typename T, bool kListed>
__global__ void __launch_bounds__(kThreadsB)
idct_planes_kernel(const FrameRecon *__restrict__ frames, const int16_t *__restrict__ coef, T *__restrict__ samples, uint32_t *__restrict__ flags,
                   const uint32_t *__restrict__ list, int first_comp) {
    const int c = first_comp + blockIdx.z;
    const uint32_t t = blockIdx.x * kThreadsB + threadIdx.x;
    if (!kListed) {
        const FrameRecon &f = frames[blockIdx.y];
        if (c < (int)f.ncomp) idct_planes_block<T>(f, c, t, coef, samples, flags);
    } else {
        const uint32_t n = list[0];
        for (uint32_t k = blockIdx.y; k < n; k += gridDim.y) {
            const FrameRecon &f = frames[list[1 + k]];
            if (c < (int)f.ncomp) idct_planes_block<T>(f, c, t, coef, samples, flags);
        }
    }
}

// the frames of a group whose `narrow` flag is set -> list = {count, group-local indices}
__global__ void __launch_bounds__(256)
narrow_list_kernel(const FrameRecon *__restrict__ frames, uint32_t n_frames, const uint32_t *__restrict__ narrow_flags, uint32_t *__restrict__ list) {
    if (threadIdx.x == 0) list[0] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_frames; i += blockDim.x)
        if (narrow_flags[frames[i].status_idx]) list[1 + atomicAdd(list, 1u)] = i;
}

// ---- b2 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// NW samples of line y of a sample plane starting at column x0. Away from the frame edge this is a plain run;
// at the edge the clamped addressing reproduces dest[-1] = dest[0], dest[width] = dest[width-1]
// (upsamplerbase.cpp:322-323) and the duplicated first / last line (upsampler.cpp:100-106).
template <int NW, typename T>
__device__ __forceinline__ void load_row(const T *__restrict__ plane, uint32_t pitch, int y, int x0, int cw, int ch, bool interior, int (&v)[NW]) {
    if (interior) {
        const T *row = plane + (uint64_t)y * pitch + x0;
#pragma unroll
        for (int j = 0; j < NW; j++) v[j] = __ldg(row + j);
    } else {
        const T *row = plane + (uint64_t)clampi(y, 0, ch - 1) * pitch;
#pragma unroll
        for (int j = 0; j < NW; j++) v[j] = __ldg(row + clampi(x0 + j, 0, cw - 1));
    }
}

// The same window for a warp whose 32 blocks all lie inside the plane horizontally: a lane fetches its own 8/SX samples
// with 16-byte loads (4 instead of 24 L1 wavefronts per warp instruction) and, for SX == 2, the two neighbouring columns
// xl, xr (already clamped = the edge replication above) with one load each; the line is clamped here.
template <int NW, int SX, typename T>
__device__ __forceinline__ void load_row_own(const T *__restrict__ plane, uint32_t pitch, int y, int ch, uint32_t bx, int xl, int xr, int (&v)[NW]) {
    const T *row = plane + (uint64_t)clampi(y, 0, ch - 1) * pitch;
    if (sizeof(T) == 2) {
        if (SX == 2) {
            const uint2 a = __ldg(reinterpret_cast<const uint2 *>(row) + bx);
            v[0] = __ldg(row + xl), v[NW - 1] = __ldg(row + xr);
            v[1] = (int)(short)(a.x & 0xffffu), v[2] = (int)a.x >> 16, v[3] = (int)(short)(a.y & 0xffffu), v[4] = (int)a.y >> 16;
        } else {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(row) + bx);
            v[0] = (int)(short)(a.x & 0xffffu), v[1] = (int)a.x >> 16, v[2] = (int)(short)(a.y & 0xffffu), v[3] = (int)a.y >> 16;
            v[NW - 4] = (int)(short)(a.z & 0xffffu), v[NW - 3] = (int)a.z >> 16, v[NW - 2] = (int)(short)(a.w & 0xffffu), v[NW - 1] = (int)a.w >> 16;
        }
    } else if (SX == 2) {
        const int4 a = __ldg(reinterpret_cast<const int4 *>(row) + bx);
        v[0] = __ldg(row + xl), v[1] = a.x, v[2] = a.y, v[3] = a.z, v[4] = a.w, v[NW - 1] = __ldg(row + xr);
    } else {
        const int4 a = __ldg(reinterpret_cast<const int4 *>(row) + 2 * bx), b = __ldg(reinterpret_cast<const int4 *>(row) + 2 * bx + 1);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[NW - 4] = b.x, v[NW - 3] = b.y, v[NW - 2] = b.z, v[NW - 1] = b.w;
    }
}

// HorizontalFilterCore<2> on a window w[0..5] (w[j] = sample at subsampled x0 - 1 + j): upsampler.cpp:283-307.
__device__ __forceinline__ void hfilter2(const int (&w)[6], int (&o)[8]) {
    o[7] = WADD(WADD(w[5], WMUL(3, w[4])), 1) >> 2;
    o[6] = WADD(WADD(w[3], WMUL(3, w[4])), 2) >> 2;
    o[5] = WADD(WADD(w[4], WMUL(3, w[3])), 1) >> 2;
    o[4] = WADD(WADD(w[2], WMUL(3, w[3])), 2) >> 2;
    o[3] = WADD(WADD(w[3], WMUL(3, w[2])), 1) >> 2;
    o[2] = WADD(WADD(w[1], WMUL(3, w[2])), 2) >> 2;
    o[1] = WADD(WADD(o[2], WMUL(3, w[1])), 1) >> 2;  // reads the freshly written out[2] (upsampler.cpp:301-302)
    o[0] = WADD(WADD(w[0], WMUL(3, w[1])), 2) >> 2;
}

__device__ __forceinline__ int sat_u8_64(long long v) { return (int)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
// four values -> four bytes, each saturated to 0..255 (CLAMP(255, v)), p0 in the low byte: two I2IP instead of four
// conversions plus the shifts and ORs
__device__ __forceinline__ uint32_t pack_sat4(int p0, int p1, int p2, int p3) {
    uint32_t hi, w;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(p3), "r"(p2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(p1), "r"(p0), "r"(hi));
    return w;
}

// ycbcrtrafo.cpp:842-850 with the matrix of colortransformerfactory.cpp:136-138 (13 fractional bits) and
// FIX_COLOR_TO_INT (tools/numerics.hpp:65).  The reference multiplies in 64 bits; 32 bits give the same result
// while |y|, |cb|, |cr| <= 65535 -- blocks that violate that (damaged streams only) are flagged `wide`.
// The results are NOT yet clamped (pack_sat4 does that), except in the 64-bit mode.
template <int MODE>  // 0: 32-bit YCbCr, 1: 64-bit YCbCr (samples outside the guarded range), 2: identity
__device__ __forceinline__ void to_rgb(int y, int cbv, int crv, int &r, int &g, int &b) {
    if (MODE == 2) {  // COLOR_TO_INT (tools/numerics.hpp:69)
        r = WADD(y, 8) >> 4;
        g = WADD(cbv, 8) >> 4;
        b = WADD(crv, 8) >> 4;
        return;
    }
    const int cb = WSUB(cbv, 128 << 4), cr = WSUB(crv, 128 << 4);
    if (MODE == 0) {
        const int yy = y * 8192 + 65536;
        r = (yy + cr * 11485) >> 17;
        g = (yy - cb * 2819 - cr * 5850) >> 17;
        b = (yy + cb * 14516) >> 17;
    } else {
        const long long Y = y, CB = cb, CR = cr;
        r = sat_u8_64((Y * 8192 + CR * 11485 + 65536) >> 17);
        g = sat_u8_64((Y * 8192 - CB * 2819 - CR * 5850 + 65536) >> 17);
        b = sat_u8_64((Y * 8192 + CB * 14516 + 65536) >> 17);
    }
}

template <int NC, int SX, int SY, typename T>
__device__ __forceinline__ void reconstruct_tile(const FrameRecon &f, int *ys, const int16_t *__restrict__ coef, const T *__restrict__ samples,
                                                 const uint32_t *__restrict__ wide_flags, uint8_t *__restrict__ out) {
    const uint32_t W = f.width, H = f.height;
    const uint32_t vbw = (W + 7) >> 3, vbh = (H + 7) >> 3;  // blocks that carry visible pixels
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bx0 = blockIdx.x * 32, bx = bx0 + lane, by = blockIdx.y * (kThreadsB / 32) + warp;
    if (by >= vbh) return;  // whole warp
    const bool valid = bx < vbw;
    const int X = 8 * (int)bx, Y = 8 * (int)by;
    int *my = ys + threadIdx.x;

    // ---- luma coefficients: the warp's 32 blocks are 4 KB of contiguous HBM. Eight coalesced 16-byte loads per lane
    // (four full lines per warp instruction instead of 32 partial ones) park them in the upper half of the warp's tile
    // columns, 16-byte piece (block, row) at row segment 32 + 4*row + block/8, slot (block%8) ^ row: conflict-free for
    // these writes and for the per-lane reads below. The row pass overwrites a segment only after its piece was consumed.
    {
        const uint32_t wseg = (uint32_t)__cvta_generic_to_shared(ys) + 128u * warp;  // bytes; row segment j at + 512*j
        const uint4 *src = reinterpret_cast<const uint4 *>(coef + f.coef_base[0] + ((uint64_t)by * f.bw[0] + bx0) * 64u);
        const uint32_t prow = lane & 7u, sub = lane >> 3;
        uint4 pc[8];
#pragma unroll
        for (uint32_t i = 0; i < 8; i++) {
            const bool have = bx0 + 4u * i + sub < f.bw[0];  // past the end of the block row: nothing (those lanes store nothing)
            pc[i] = have ? __ldg(src + i * 32u + lane) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (uint32_t i = 0; i < 8; i++) {
            const uint32_t blk = 4u * i + sub;
            const uint32_t dst = wseg + 512u * (32u + 4u * prow + (blk >> 3)) + 16u * ((blk & 7u) ^ prow);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pc[i].x), "r"(pc[i].y), "r"(pc[i].z), "r"(pc[i].w) : "memory");
        }
        __syncwarp();
        // ---- row pass (dct/idct.cpp:237-287): the next row's piece is read before this row's results are written
        const uint32_t mine = wseg + 512u * (32u + (lane >> 3));
        auto piece = [&](int r) {
            uint4 q;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                         : "r"(mine + 2048u * (uint32_t)r + 16u * ((lane & 7u) ^ (uint32_t)r))
                         : "memory");
            return q;
        };
        uint4 q = piece(0);
#pragma unroll 1
        for (int r = 0; r < 8; r++) {
            const uint4 qn = piece((r < 7) ? r + 1 : r);
            int v[8];
            unpack_row_raw(q, v);
            if (r == 0) v[0] = WADD(v[0], 128 << 3);  // dcoffset << 3 (idct.cpp:233,244), preshift folded into the rounding (unpack_row_raw)
            idct8t<16, 5>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; k++) my[(8 * r + k) * kThreadsB] = v[k];
            q = qn;
        }
    }
    // ---- column pass (:291-334) + range guard for the 32-bit colour arithmetic
    int mx = 0, mn = 0;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        int v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = my[(8 * r + k) * kThreadsB];
        idct8t<2048, 12>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
#pragma unroll
        for (int r = 0; r < 8; r++) my[(8 * r + k) * kThreadsB] = v[r];
        mx = __vimax3_s32(__vimax3_s32(mx, v[0], v[1]), v[2], v[3]);
        mx = __vimax3_s32(__vimax3_s32(mx, v[4], v[5]), v[6], v[7]);
        mn = __vimin3_s32(__vimin3_s32(mn, v[0], v[1]), v[2], v[3]);
        mn = __vimin3_s32(__vimin3_s32(mn, v[4], v[5]), v[6], v[7]);
    }

    const int xmax = (X + 7 < (int)W) ? 7 : (int)((W - 1) & 7);
    const int ymax = (Y + 7 < (int)H) ? 7 : (int)((H - 1) & 7);
    const uint32_t opitch = W * NC;
    uint8_t *obase = out + f.out_base + (uint64_t)Y * opitch + (uint64_t)X * NC;

    if (NC == 1) {  // identity, COLOR_TO_INT (numerics.hpp:69) + clamp
        if (!valid) return;
#pragma unroll 1
        for (int r = 0; r <= ymax; r++) {
            uint8_t *o = obase + (uint64_t)r * opitch;
            int px[8];
#pragma unroll
            for (int x = 0; x < 8; x++) px[x] = WADD(my[(8 * r + x) * kThreadsB], 8) >> 4;
            const uint32_t w0 = pack_sat4(px[0], px[1], px[2], px[3]), w1 = pack_sat4(px[4], px[5], px[6], px[7]);
            if (xmax == 7 && ((reinterpret_cast<uintptr_t>(o) & 7u) == 0)) {
                *reinterpret_cast<uint2 *>(o) = make_uint2(w0, w1);
            } else {
#pragma unroll
                for (int x = 0; x < 8; x++)
                    if (x <= xmax) o[x] = (uint8_t)(((x < 4) ? w0 : w1) >> (8 * (x & 3)));
            }
        }
        return;
    }

    constexpr int NW = (SX == 2) ? 6 : 8;
    const int cw = (int)f.cw, ch = (int)f.ch;
    const uint32_t cpitch = 8u * f.bw[1];
    const T *p1 = samples + f.sample_base[1];
    const T *p2 = samples + f.sample_base[2];
    const int cx0 = X / SX - ((SX == 2) ? 1 : 0);  // window column 0 (upsampler.cpp:87,108-109)
    const int cy0 = Y / SY;
    const bool ycbcr = f.ycbcr != 0;
    // int16 planes: chroma is inside the guarded range by construction, only this block's luma can leave it
    const bool wide = (sizeof(T) == 4 && wide_flags[f.status_idx] != 0u) || mx > kWide || mn < -kWide;
    // the whole window lies inside the plane: no clamping needed
    const bool interior = valid && cx0 >= 0 && cx0 + NW <= cw && cy0 - 1 >= 0 && cy0 + ((SY == 2) ? 5 : 8) <= ch;
    // warp-uniform: every block of the warp is visible and every lane's own samples lie inside the plane
    constexpr int OWN = 8 / SX;
    const bool fast = (bx0 + 32 <= vbw) && ((int)((bx0 + 32) * OWN) <= cw);
    const int xl = (cx0 > 0) ? cx0 : 0, xr = (cx0 + NW - 1 < cw) ? cx0 + NW - 1 : cw - 1;
    auto fetch = [&](int y, int (&d1)[NW], int (&d2)[NW]) {
        if (fast) {
            load_row_own<NW, SX, T>(p1, cpitch, y, ch, bx, xl, xr, d1);
            load_row_own<NW, SX, T>(p2, cpitch, y, ch, bx, xl, xr, d2);
        } else if (valid) {
            load_row<NW, T>(p1, cpitch, y, cx0, cw, ch, interior, d1);
            load_row<NW, T>(p2, cpitch, y, cx0, cw, ch, interior, d2);
        }
    };

    // rolling lines: SY == 2 keeps top/cur/bot (upsampler.cpp:92-106), SY == 1 only cur
    int top1[NW], cur1[NW], bot1[NW], top2[NW], cur2[NW], bot2[NW];
    if (SY == 2) {
        fetch(cy0 - 1, top1, top2);
        fetch(cy0 + 1, bot1, bot2);
    }
    fetch(cy0, cur1, cur2);

    auto lines = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        // one output line; for SY == 2 the line parity is a compile-time constant (the loop below is unrolled by two),
        // so the choice of the neighbour line and of the rounding needs no selects
        auto one_line = [&](int r, auto odd_tag) {
            constexpr bool odd = decltype(odd_tag)::value;
            uint32_t wd[6];
            if (valid) {
                int px[24];
                int v1[NW], v2[NW];
                if (SY == 2) {  // VerticalFilterCore<2>, upsampler.cpp:136-168: even lines lean on top, odd lines on bot
                    constexpr int ra = odd ? 1 : 2, rb = odd ? 2 : 1;  // rounding of even / odd window columns
#pragma unroll
                    for (int j = 0; j < NW; j += 2) {
                        v1[j] = tap<ra>(odd ? bot1[j] : top1[j], cur1[j]), v1[j + 1] = tap<rb>(odd ? bot1[j + 1] : top1[j + 1], cur1[j + 1]);
                        v2[j] = tap<ra>(odd ? bot2[j] : top2[j], cur2[j]), v2[j + 1] = tap<rb>(odd ? bot2[j + 1] : top2[j + 1], cur2[j + 1]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NW; j++) {
                        v1[j] = cur1[j];
                        v2[j] = cur2[j];
                    }
                }
                int c1[8], c2[8];
                if (SX == 2) {
                    int w1[6], w2[6];
#pragma unroll
                    for (int j = 0; j < 6; j++) {
                        w1[j] = v1[j];
                        w2[j] = v2[j];
                    }
                    hfilter2t(w1, c1);
                    hfilter2t(w2, c2);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        c1[j] = v1[j];
                        c2[j] = v2[j];
                    }
                }
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    const int yv = my[(8 * r + x) * kThreadsB];
                    int R, G, B;
                    to_rgb<MODE>(yv, c1[x], c2[x], R, G, B);
                    px[3 * x] = R;
                    px[3 * x + 1] = G;
                    px[3 * x + 2] = B;
                }
#pragma unroll
                for (int k = 0; k < 6; k++) wd[k] = pack_sat4(px[4 * k], px[4 * k + 1], px[4 * k + 2], px[4 * k + 3]);
            }
            if (valid) {
                uint8_t *o = obase + (uint64_t)r * opitch;
                if (xmax == 7 && ((reinterpret_cast<uintptr_t>(o) & 7u) == 0)) {
                    uint2 *o2 = reinterpret_cast<uint2 *>(o);
#pragma unroll
                    for (int k = 0; k < 3; k++) o2[k] = make_uint2(wd[2 * k], wd[2 * k + 1]);
                } else {
#pragma unroll
                    for (int x = 0; x < 8; x++) {
                        if (x <= xmax) {
#pragma unroll
                            for (int i = 3 * x; i < 3 * x + 3; i++) o[i] = (uint8_t)(wd[i >> 2] >> (8 * (i & 3)));
                        }
                    }
                }
            }
        };
        if (SY == 2) {
#pragma unroll 1
            for (int r = 0; r < 8; r += 2) {
                if (r > ymax) break;  // uniform over the warp (same block row)
                one_line(r, std::false_type());
                if (r + 1 <= ymax) one_line(r + 1, std::true_type());
                // advance the line window after every odd output line (upsampler.cpp:160-165)
                if (valid) {
#pragma unroll
                    for (int j = 0; j < NW; j++) {
                        top1[j] = cur1[j];
                        cur1[j] = bot1[j];
                        top2[j] = cur2[j];
                        cur2[j] = bot2[j];
                    }
                }
                if (r < 6) fetch(cy0 + (r >> 1) + 2, bot1, bot2);
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < 8; r++) {
                if (r > ymax) break;
                one_line(r, std::false_type());
                if (r < 7) fetch(cy0 + r + 1, cur1, cur2);
            }
        }
    };
    // one instantiation per colour mode keeps every per-pixel branch out of the line loop
    if (!ycbcr) lines(std::integral_constant<int, 2>());
    else if (wide) lines(std::integral_constant<int, 1>());
    else lines(std::integral_constant<int, 0>());
}

// grid (luma block columns / 32, luma block rows / 4, frames or kWideSlots)
// five CTAs per SM: ptxas fits the kernel into 86 registers without spilling, and 20 resident warps hide the load latency
// better than 16 (six CTAs = 80 registers is slower again)
#ifndef B200JPG_B2_MINCTAS
#define B200JPG_B2_MINCTAS 5
#endif
template <int NC, int SX, int SY, typename T, bool kListed>
__global__ void __launch_bounds__(kThreadsB, B200JPG_B2_MINCTAS)
reconstruct_kernel(const FrameRecon *__restrict__ frames, const int16_t *__restrict__ coef, const T *__restrict__ samples,
                   const uint32_t *__restrict__ wide_flags, const uint32_t *__restrict__ list, uint8_t *__restrict__ out) {
    __shared__ int ys[64 * kThreadsB];  // [coefficient][thread]; warps only ever touch their own 32 columns
    if (!kListed) {
        reconstruct_tile<NC, SX, SY, T>(frames[blockIdx.z], ys, coef, samples, wide_flags, out);
    } else {
        const uint32_t n = list[0];
        for (uint32_t k = blockIdx.z; k < n; k += gridDim.z) {
            reconstruct_tile<NC, SX, SY, T>(frames[list[1 + k]], ys, coef, samples, wide_flags, out);
            __syncwarp();
        }
    }
}


// =====================================================================================================
// fused reconstruction for 4:2:0 frames: chroma IDCT + luma IDCT + upsampling + colour + store in ONE kernel
// =====================================================================================================
// A CTA is one warp. It owns a strip of 32 luma block columns (256 pixels) and walks a segment of the frame top to bottom,
// two luma block rows per chroma block row:
//     [C(r0-1)]  C(r0)   L(2 r0)  C(r0+1)  L(2 r0 + 1)  L(2 r0 + 2)  C(r0+2)  L(2 r0 + 3) ...
// C(k): the 16 Cb + 16 Cr blocks under the strip -- one per lane, the same two compact IDCT loops as a luma row -- go as
// int16 samples into a two-slot ring in shared memory (slot k & 1: eight sample rows of 128 columns per component plus the
// one column either side that the upsampling window of the strip's first / last luma block needs).  Those two halo columns
// belong to the neighbouring strips' blocks; only that one column of them is computed: per (component, side) eight single
// outputs of the row pass -- one dot product per lane, the row pass is linear before its rounding shift -- and one column
// pass.  L(j): exactly the luma row of reconstruct_kernel, with the chroma window read from the ring instead of from planes
// in HBM.  Nothing but coefficients is read from and nothing but pixels is written to HBM: the sample planes, the kernel that
// wrote them and their 16.5 MB per 4K frame round trip are gone.  A segment recomputes one chroma block row of its upper
// neighbour (1 / seg_rows of the chroma work); the host picks seg_rows so that the grid fills the chip.
// Frames whose chroma samples do not fit int16 are flagged `narrow` and redone by the int32 instance (list-driven), as before.
// ring: [slot 2][component 2][row 8][128 columns] of T, followed by the halo columns [slot][component][row][left, right]
template <typename T>
struct RingOf {
    static constexpr int kRowBytes = 128 * (int)sizeof(T);
    static constexpr int kCompBytes = 8 * kRowBytes;
    static constexpr int kSlotBytes = 2 * kCompBytes;
    static constexpr int kMainBytes = 2 * kSlotBytes;
    static constexpr int kHaloRowBytes = 2 * (int)sizeof(T);
    static constexpr int kHaloCompBytes = 8 * kHaloRowBytes;
    static constexpr int kHaloSlotBytes = 2 * kHaloCompBytes;
    static constexpr int kBytes = kMainBytes + 2 * kHaloSlotBytes;
};

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lds_s16(uint32_t a) {
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts16(uint32_t a, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((short)v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, int v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// one ring sample (T = short / int) at a shared-space byte address
template <typename T>
__device__ __forceinline__ int ring_load(uint32_t a) {
    return sizeof(T) == 2 ? lds_s16(a) : (int)lds32(a);
}
template <typename T>
__device__ __forceinline__ void ring_store(uint32_t a, int v) {
    if (sizeof(T) == 2) sts16(a, v);
    else sts32(a, v);
}

// The warp's 32 blocks of coefficients come in as 4 KB of 16-byte pieces, eight per lane (piece i * 32 + lane): a luma row is
// one contiguous run, a chroma row two runs of 2 KB (Cb blocks 0..15, Cr blocks 16..31). tile_fetch only ISSUES the loads --
// the strip walk calls it for the NEXT unit before it computes the current one, so HBM latency hides behind a whole unit of
// arithmetic --, tile_idct parks the pieces in the upper half of the tile (piece (block, row) at row segment 32 + 4 row +
// block / 8, slot (block % 8) ^ row: conflict-free both ways), hands every lane its own block and runs the row pass and the
// column pass (dct/idct.cpp:237-334) as two compact loops over the tile [sample][lane].
template <typename Have>
__device__ __forceinline__ void tile_fetch(uint4 (&pc)[8], const uint4 *__restrict__ src_a, const uint4 *__restrict__ src_b, Have have) {
    const uint32_t lane = threadIdx.x & 31, sub = lane >> 3;
#pragma unroll
    for (uint32_t i = 0; i < 8; i++) {
        const uint4 *p = (i < 4) ? src_a + (i * 32u + lane) : src_b + ((i - 4u) * 32u + lane);
        pc[i] = have(4u * i + sub) ? __ldg(p) : make_uint4(0u, 0u, 0u, 0u);
    }
}

__device__ __forceinline__ void tile_stash(int *ys, const uint4 (&pc)[8]) {
    const uint32_t lane = threadIdx.x & 31, prow = lane & 7u, sub = lane >> 3;
    const uint32_t wseg = (uint32_t)__cvta_generic_to_shared(ys);
#pragma unroll
    for (uint32_t i = 0; i < 8; i++) {
        const uint32_t blk = 4u * i + sub;
        const uint32_t dst = wseg + 128u * (32u + 4u * prow + (blk >> 3)) + 16u * ((blk & 7u) ^ prow);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pc[i].x), "r"(pc[i].y), "r"(pc[i].z), "r"(pc[i].w) : "memory");
    }
    __syncwarp();
}

// Returns this lane's block's min / max sample in mn / mx. `sink(k, v)` receives column k of the block (v[r] = sample of row r):
// a luma row puts it back into the tile for the line loop, a chroma row sends it straight to the sample ring.
template <typename Sink>
__device__ __forceinline__ void tile_idct(int *ys, int &mn, int &mx, Sink sink) {
    const uint32_t lane = threadIdx.x & 31;
    int *my = ys + lane;
    const uint32_t wseg = (uint32_t)__cvta_generic_to_shared(ys);
    const uint32_t mine = wseg + 128u * (32u + (lane >> 3));
    auto piece = [&](int r) { return lds128(mine + 512u * (uint32_t)r + 16u * ((lane & 7u) ^ (uint32_t)r)); };
    uint4 q = piece(0);
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
        const uint4 qn = piece((r < 7) ? r + 1 : r);
        int v[8];
        unpack_row_raw(q, v);
        if (r == 0) v[0] = WADD(v[0], 128 << 3);  // dcoffset << 3 (idct.cpp:233,244), the << 4 preshift left out like in the inputs
        idct8t<16, 5>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) my[(8 * r + k) * 32] = v[k];
        q = qn;
    }
    mx = 0, mn = 0;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        int v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = my[(8 * r + k) * 32];
        idct8t<2048, 12>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        sink(k, v);
        mx = __vimax3_s32(__vimax3_s32(mx, v[0], v[1]), v[2], v[3]);
        mx = __vimax3_s32(__vimax3_s32(mx, v[4], v[5]), v[6], v[7]);
        mn = __vimin3_s32(__vimin3_s32(mn, v[0], v[1]), v[2], v[3]);
        mn = __vimin3_s32(__vimin3_s32(mn, v[4], v[5]), v[6], v[7]);
    }
}

template <typename T, bool kListed>
__device__ __forceinline__ void reconstruct420_strip(const FrameRecon &f, int *ys, uint32_t ring, const int16_t *__restrict__ coef,
                                                     uint32_t *__restrict__ narrow_flags, uint8_t *__restrict__ out, uint32_t seg_rows) {
    using R = RingOf<T>;
    const uint32_t W = f.width, H = f.height;
    const uint32_t vbw = (W + 7) >> 3, vbh = (H + 7) >> 3;  // luma blocks that carry visible pixels
    const int cw = (int)f.cw, ch = (int)f.ch;
    const uint32_t vch = ((uint32_t)ch + 7u) >> 3;          // chroma block rows that carry real lines
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sx = blockIdx.x, bx0 = sx * 32u, bx = bx0 + lane;
    const uint32_t r0 = blockIdx.y * seg_rows;
    const uint32_t rows_all = (vbh + 1u) >> 1;              // chroma block rows that have luma rows to serve
    if (r0 >= rows_all || bx0 >= vbw) return;
    const uint32_t r1 = (r0 + seg_rows < rows_all) ? r0 + seg_rows : rows_all;
    const uint32_t bwc = f.bw[1];
    int *my = ys + lane;
    uint32_t wide_slot[2] = {0u, 0u};  // warp-uniform: a chroma sample of that ring slot lies outside the 32-bit colour range

    // ---- C(k): chroma block row k -> ring slot k & 1
    auto chroma_row = [&](uint32_t k) {
        const uint32_t slot = k & 1u;
        const uint32_t comp = lane >> 4, bl = lane & 15u;  // this lane's block: Cb / Cr, column 16 sx + bl
        const uint32_t cbx0 = 16u * sx;
        int mn, mx;
        // samples -> ring, one column of the block per step of the column pass (two lanes share a bank: comp stride and the
        // 16-byte block pitch put lanes bl and bl + 8 on the same one)
        const uint32_t base = ring + slot * R::kSlotBytes + comp * R::kCompBytes + (8u * bl) * (uint32_t)sizeof(T);
        tile_idct(ys, mn, mx, [&](int k, const int (&v)[8]) {
#pragma unroll
            for (int r = 0; r < 8; r++) ring_store<T>(base + (uint32_t)r * R::kRowBytes + (uint32_t)k * (uint32_t)sizeof(T), v[r]);
        });
        // ---- halo columns: lane (h, r) = (lane >> 3, lane & 7); h & 1 = side (0: column 7 of the block to the left, 1: column 0
        // of the block to the right), h >> 1 = component. Output 0 / 7 of the row pass over coefficient row r is one dot
        // product (the butterfly is linear in front of its rounding shift):
        //   out0 = 512 v0 + 710 v1 + 669 v2 + 602 v3 + 512 v4 + 402 v5 + 277 v6 + 141 v7,  out7 = the same with the odd terms negated
        {
            const uint32_t h = lane >> 3, r = lane & 7u, side = h & 1u, hc = h >> 1;
            const int nb = side ? (int)(cbx0 + 16u) : (int)cbx0 - 1;
            const bool exists = nb >= 0 && nb < (int)bwc;
            int v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (exists) unpack_row_raw(__ldg(reinterpret_cast<const uint4 *>(coef + f.coef_base[1 + hc] + ((uint64_t)k * bwc + (uint32_t)nb) * 64u) + r), v);
            if (r == 0) v[0] = WADD(v[0], 128 << 3);
            const int even = WADD(WADD(WMUL(v[0], 512), WMUL(v[2], 669)), WADD(WMUL(v[4], 512), WMUL(v[6], 277)));
            const int odd = WADD(WADD(WMUL(v[1], 710), WMUL(v[3], 602)), WADD(WMUL(v[5], 402), WMUL(v[7], 141)));
            const int inter = WADD(side ? WADD(even, odd) : WSUB(even, odd), 16) >> 5;  // row pass without the << 4 preshift, see unpack_row_raw
            int c[8];
#pragma unroll
            for (int i = 0; i < 8; i++) c[i] = __shfl_sync(0xffffffffu, inter, (int)((lane & ~7u) + (uint32_t)i));
            idct8t<2048, 12>(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]);
            int mine = c[0];
#pragma unroll
            for (int i = 1; i < 8; i++) mine = (r == (uint32_t)i) ? c[i] : mine;
            ring_store<T>(ring + R::kMainBytes + slot * R::kHaloSlotBytes + hc * R::kHaloCompBytes + r * R::kHaloRowBytes + side * (uint32_t)sizeof(T), mine);
            if (exists) {
                mx = mine > mx ? mine : mx;
                mn = mine < mn ? mine : mn;
            }
        }
        if (sizeof(T) == 2) {
            if (mx > 32767 || mn < -32768) atomicOr(narrow_flags + f.status_idx, 1u);  // `narrow`: the frame needs the int32 ring
        } else {
            wide_slot[slot] = __any_sync(0xffffffffu, mx > kWide || mn < -kWide) ? 1u : 0u;
        }
        __syncwarp();
    };

    // ---- L(by): one row of 32 luma blocks -> pixels
    const uint32_t opitch = W * 3u;
    const bool valid = bx < vbw;
    const int X = 8 * (int)bx;
    const int xmax = (X + 7 < (int)W) ? 7 : (int)((W - 1) & 7);
    const int cx0 = X / 2 - 1;
    // this lane's six window columns lie inside the plane: own four by one 8-byte (int16) load, the neighbours by one load each
    const bool fastx = valid && cx0 >= 0 && cx0 + 6 <= cw;
    const uint32_t own_off = (4u * lane) * (uint32_t)sizeof(T);
    const uint32_t left_off = (4u * lane - 1u) * (uint32_t)sizeof(T), right_off = (4u * lane + 4u) * (uint32_t)sizeof(T);
    auto luma_row = [&](uint32_t by) {
        const int Y = 8 * (int)by;
        int mn, mx;
        tile_idct(ys, mn, mx, [&](int k, const int (&v)[8]) {
#pragma unroll
            for (int r = 0; r < 8; r++) my[(8 * r + k) * 32] = v[r];
        });
        if (!valid) return;
        const int ymax = (Y + 7 < (int)H) ? 7 : (int)((H - 1) & 7);
        uint8_t *obase = out + f.out_base + (uint64_t)Y * opitch + (uint64_t)X * 3u;
        const int cy0 = Y / 2;
        const bool ycbcr = f.ycbcr != 0;
        const bool wide = (sizeof(T) == 4 && (wide_slot[0] | wide_slot[1]) != 0u) || mx > kWide || mn < -kWide;
        // chroma line y (clamped to the plane = the duplicated first / last line, upsampler.cpp:100-106) of both components
        auto fetch = [&](int y, int (&d1)[6], int (&d2)[6]) {
            const int yc = clampi(y, 0, ch - 1);
            const uint32_t slot = ((uint32_t)yc >> 3) & 1u, rr = (uint32_t)yc & 7u;
            const uint32_t rowa = ring + slot * R::kSlotBytes + rr * R::kRowBytes;
            const uint32_t haloa = ring + R::kMainBytes + slot * R::kHaloSlotBytes + rr * R::kHaloRowBytes;  // {left, right} of the row
            if (fastx) {
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    int (&d)[6] = c ? d2 : d1;
                    const uint32_t a = rowa + (uint32_t)c * R::kCompBytes, ha = haloa + (uint32_t)c * R::kHaloCompBytes;
                    if (sizeof(T) == 2) {
                        const uint2 o = lds64(a + own_off);
                        d[1] = (int)(short)(o.x & 0xffffu), d[2] = (int)o.x >> 16, d[3] = (int)(short)(o.y & 0xffffu), d[4] = (int)o.y >> 16;
                    } else {
                        const uint4 o = lds128(a + own_off);
                        d[1] = (int)o.x, d[2] = (int)o.y, d[3] = (int)o.z, d[4] = (int)o.w;
                    }
                    d[0] = ring_load<T>(lane == 0 ? ha : a + left_off);
                    d[5] = ring_load<T>(lane == 31 ? ha + (uint32_t)sizeof(T) : a + right_off);
                }
            } else {
                // at the frame edge: dest[-1] = dest[0], dest[width] = dest[width-1] at the TRUE subsampled width
                // (upsamplerbase.cpp:322-323) = clamped addressing; a clamped column is a strip column or one of its halos
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    const int local = clampi(cx0 + j, 0, cw - 1) - 128 * (int)sx;
                    const bool in_halo = local < 0 || local > 127;
                    const uint32_t a = in_halo ? haloa + (local < 0 ? 0u : (uint32_t)sizeof(T)) : rowa + (uint32_t)local * (uint32_t)sizeof(T);
                    d1[j] = ring_load<T>(a);
                    d2[j] = ring_load<T>(a + (in_halo ? R::kHaloCompBytes : R::kCompBytes));
                }
            }
        };
        // four chroma lines live at a time: output lines r, r+1 lean on (A, B, C) = (top, cur, bot), lines r+2, r+3 on (B, C, D)
        // (upsampler.cpp:92-106,160-165); the loop below runs twice, so the line window is renamed once instead of rotated four times
        int A1[6], B1[6], C1[6], D1[6], A2[6], B2[6], C2[6], D2[6];
        fetch(cy0 - 1, A1, A2);
        fetch(cy0, B1, B2);
        fetch(cy0 + 1, C1, C2);
        fetch(cy0 + 2, D1, D2);
        auto lines = [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            // one output line: `n` = the neighbour line it leans on (top for even lines, bot for odd ones), `c` = the current line
            auto one_line = [&](int r, auto odd_tag, const int (&n1)[6], const int (&c1w)[6], const int (&n2)[6], const int (&c2w)[6]) {
                constexpr bool odd = decltype(odd_tag)::value;
                constexpr int ra = odd ? 1 : 2, rb = odd ? 2 : 1;  // rounding of even / odd window columns (upsampler.cpp:136-168)
                int w1[6], w2[6];
#pragma unroll
                for (int j = 0; j < 6; j += 2) {
                    w1[j] = tap<ra>(n1[j], c1w[j]), w1[j + 1] = tap<rb>(n1[j + 1], c1w[j + 1]);
                    w2[j] = tap<ra>(n2[j], c2w[j]), w2[j + 1] = tap<rb>(n2[j + 1], c2w[j + 1]);
                }
                int c1[8], c2[8];
                hfilter2t(w1, c1);
                hfilter2t(w2, c2);
                int px[24];
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    int Rr, G, B;
                    to_rgb<MODE>(my[(8 * r + x) * 32], c1[x], c2[x], Rr, G, B);
                    px[3 * x] = Rr, px[3 * x + 1] = G, px[3 * x + 2] = B;
                }
                uint32_t wd[6];
#pragma unroll
                for (int k = 0; k < 6; k++) wd[k] = pack_sat4(px[4 * k], px[4 * k + 1], px[4 * k + 2], px[4 * k + 3]);
                uint8_t *o = obase + (uint64_t)r * opitch;
                if (xmax == 7 && ((reinterpret_cast<uintptr_t>(o) & 7u) == 0)) {
                    uint2 *o2 = reinterpret_cast<uint2 *>(o);
#pragma unroll
                    for (int k = 0; k < 3; k++) o2[k] = make_uint2(wd[2 * k], wd[2 * k + 1]);
                } else {
#pragma unroll
                    for (int x = 0; x < 8; x++) {
                        if (x <= xmax) {
#pragma unroll
                            for (int i = 3 * x; i < 3 * x + 3; i++) o[i] = (uint8_t)(wd[i >> 2] >> (8 * (i & 3)));
                        }
                    }
                }
            };
#pragma unroll 1
            for (int r = 0; r < 8; r += 4) {
                if (r > ymax) break;
                one_line(r, std::false_type(), A1, B1, A2, B2);
                if (r + 1 <= ymax) one_line(r + 1, std::true_type(), C1, B1, C2, B2);
                if (r + 2 <= ymax) one_line(r + 2, std::false_type(), B1, C1, B2, C2);
                if (r + 3 <= ymax) one_line(r + 3, std::true_type(), D1, C1, D2, C2);
                if (r == 0) {
#pragma unroll
                    for (int j = 0; j < 6; j++) {
                        A1[j] = C1[j], B1[j] = D1[j];
                        A2[j] = C2[j], B2[j] = D2[j];
                    }
                    fetch(cy0 + 3, C1, C2);
                    fetch(cy0 + 4, D1, D2);
                }
            }
        };
        if (!ycbcr) lines(std::integral_constant<int, 2>());
        else if (wide) lines(std::integral_constant<int, 1>());
        else lines(std::integral_constant<int, 0>());
    };

    // ---- the walk: [C(r0-1)] C(r0) { L(2k) [C(k+1)] [L(2k+1)] } for k = r0 .. r1-1, the coefficients of unit n+1 in flight
    // (32 registers per lane) while unit n is computed
    uint32_t kk = r0;
    int phase = 0;
    auto next_unit = [&](int &kind, uint32_t &idx) {  // kind 0: none, 1: chroma block row idx, 2: luma block row idx
        for (;;) {
            switch (phase) {
            case 0:
                phase = 1;
                if (r0 > 0) {
                    kind = 1, idx = r0 - 1;
                    return;
                }
                break;
            case 1:
                phase = 2;
                kind = 1, idx = r0;
                return;
            case 2:
                if (kk >= r1) {
                    phase = 5;
                    break;
                }
                phase = 3;
                kind = 2, idx = 2u * kk;
                return;
            case 3:
                phase = 4;
                if (kk + 1u < vch) {
                    kind = 1, idx = kk + 1u;
                    return;
                }
                break;
            case 4: {
                const uint32_t odd = 2u * kk + 1u;
                phase = 2;
                kk++;
                if (odd < vbh) {
                    kind = 2, idx = odd;
                    return;
                }
                break;
            }
            default:
                kind = 0, idx = 0;
                return;
            }
        }
    };
    uint4 pc[8];
    auto fetch_unit = [&](int kind, uint32_t idx) {
        if (kind == 1) {
            const uint32_t cbx0 = 16u * sx;
            const uint4 *src_a = reinterpret_cast<const uint4 *>(coef + f.coef_base[1] + ((uint64_t)idx * bwc + cbx0) * 64u);
            const uint4 *src_b = reinterpret_cast<const uint4 *>(coef + f.coef_base[2] + ((uint64_t)idx * bwc + cbx0) * 64u);
            tile_fetch(pc, src_a, src_b, [&](uint32_t blk) { return cbx0 + (blk & 15u) < bwc; });
        } else if (kind == 2) {
            const uint4 *src = reinterpret_cast<const uint4 *>(coef + f.coef_base[0] + ((uint64_t)idx * f.bw[0] + bx0) * 64u);
            tile_fetch(pc, src, src + 128, [&](uint32_t blk) { return bx0 + blk < f.bw[0]; });
        }
    };
    int kind, nkind;
    uint32_t idx, nidx;
    next_unit(kind, idx);
    fetch_unit(kind, idx);
    while (kind) {
        next_unit(nkind, nidx);
        __syncwarp();  // every lane is done with the tile (and, before a chroma unit, with the ring slot it replaces)
        tile_stash(ys, pc);
        fetch_unit(nkind, nidx);
        if (kind == 1) chroma_row(idx);
        else luma_row(idx);
        kind = nkind, idx = nidx;
    }
}

// twelve one-warp CTAs per SM: 168 registers hold the prefetched coefficients of the next unit and a four-line chroma window
// without spilling (13 and 10 CTAs per SM run at the same speed: the kernel is bound by instruction issue, not by occupancy)
#ifndef B200JPG_FUSED_CTAS
#define B200JPG_FUSED_CTAS 12
#endif
// grid (strips of 32 luma block columns, segments of seg_rows chroma block rows, frames or kWideSlots); one warp per CTA
template <typename T, bool kListed>
__global__ void __launch_bounds__(32, (sizeof(T) == 2) ? B200JPG_FUSED_CTAS : 8)
reconstruct420_kernel(const FrameRecon *__restrict__ frames, const int16_t *__restrict__ coef, uint32_t *__restrict__ narrow_flags,
                      const uint32_t *__restrict__ list, uint8_t *__restrict__ out, uint32_t seg_rows) {
    __shared__ int ys[64 * 32];
    __shared__ __align__(16) uint8_t ring_mem[RingOf<T>::kBytes];
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(ring_mem);
    if (!kListed) {
        reconstruct420_strip<T, kListed>(frames[blockIdx.z], ys, ring, coef, narrow_flags, out, seg_rows);
    } else {
        const uint32_t n = list[0];
        for (uint32_t k = blockIdx.z; k < n; k += gridDim.z) {
            reconstruct420_strip<T, kListed>(frames[list[1 + k]], ys, ring, coef, narrow_flags, out, seg_rows);
            __syncwarp();
        }
    }
}


// =====================================================================================================
// generic reconstruction: any number of components (1..4), any subsampling factors (1..4, per component)
// =====================================================================================================
// The formats outside the tuned kernels -- chroma factors 3 and 4, components with different factors, a subsampled first
// component, two or four components -- are rare, so this path is written for exactness, not for speed: every
// component goes through idct_planes_kernel into an int32 sample plane, then one thread per 8x8 OUTPUT block follows
// Upsampler<sx,sy>::UpsampleRegion to the letter (upsampling/upsampler.cpp:83-112, VerticalFilterCore<1..4> :114-268,
// HorizontalFilterCore<1..4> :270-386 -- including the in-place stores of the horizontal cores, whose order is part of the
// result) and the colour transformation in 64 bits (ycbcrtrafo.cpp:842-850 / identity numerics.hpp:69).
__device__ __forceinline__ int gmix(int a, int wa, int b, int wb, int rnd, int sh) {
    return (int)((unsigned)wa * (unsigned)a + (unsigned)wb * (unsigned)b + (unsigned)rnd) >> sh;
}

// one 8x8 block of component samples at output position (X, Y), upsampled by (sx, sy); plane = IDCT output, pitch in samples,
// (w, h) = true subsampled size; edge replication dest[-1] = dest[0], dest[w] = dest[w-1] (upsamplerbase.cpp:322-323) by clamping
__device__ void generic_upsample_block(const int32_t *__restrict__ plane, uint32_t pitch, int w, int h, int sx, int sy, int X, int Y, int *out) {
    const int y = Y / sy;
    const int x0 = X / sx - ((sx > 1) ? 1 : 0);
    int top = (y > 0) ? y - 1 : y, cur = y, bot = (y + 1 < h) ? y + 1 : y;
    int ymod = Y % sy;
    const int xmod = X % sx;
    auto at = [&](int x, int yy) { return plane[(size_t)yy * pitch + (size_t)clampi(x, 0, w - 1)]; };
    for (int row = 0; row < 8; row++) {
        int *o = out + 8 * row;
        int advance = 0;
        if (sy == 1) {
            for (int j = 0; j < 8; j++) o[j] = at(x0 + j, cur);
            if (cur + 1 < h) cur++;
        } else if (sy == 2) {
            const int nb = (ymod == 0) ? top : bot;
            for (int j = 0; j < 8; j++) o[j] = gmix(at(x0 + j, nb), 1, at(x0 + j, cur), 3, ((j & 1) == ymod) ? 2 : 1, 2);
            advance = (ymod == 1);
        } else if (sy == 3) {
            if (ymod == 1) {
                for (int j = 0; j < 8; j++) o[j] = at(x0 + j, cur);
            } else {
                const int nb = (ymod == 0) ? top : bot;
                for (int j = 0; j < 8; j++) o[j] = gmix(at(x0 + j, nb), 1, at(x0 + j, cur), 3, (((j & 1) == 0) == (ymod == 0)) ? 2 : 1, 2);
            }
            advance = (ymod == 2);
        } else {
            const int nb = (ymod < 2) ? top : bot;
            const bool far = (ymod == 0 || ymod == 3);  // far from the sample line: weights 3:5, else 1:7
            for (int j = 0; j < 8; j++) {
                int rnd;
                if (ymod == 0 || ymod == 2 || ymod == 3) rnd = (j & 1) ? 3 : 4;
                else rnd = (j & 1) ? 4 : 3;
                o[j] = gmix(at(x0 + j, nb), far ? 3 : 1, at(x0 + j, cur), far ? 5 : 7, rnd, 3);
            }
            advance = (ymod == 3);
        }
        if (sy > 1) {
            if (advance) {
                ymod = 0;
                top = cur;
                cur = bot;
                if (bot + 1 < h) bot++;
            } else {
                ymod++;
            }
        }
        // horizontal, in place, in the reference's store order (src[i] = o[i + 1])
        if (sx == 2) {
            int *src = o + 1, t;
            o[7] = gmix(src[4], 1, src[3], 3, 1, 2);
            o[6] = gmix(src[2], 1, src[3], 3, 2, 2);
            o[5] = gmix(src[3], 1, src[2], 3, 1, 2);
            o[4] = gmix(src[1], 1, src[2], 3, 2, 2);
            o[3] = gmix(src[2], 1, src[1], 3, 1, 2);
            o[2] = gmix(src[0], 1, src[1], 3, 2, 2);
            t = src[0];
            o[1] = gmix(src[1], 1, t, 3, 1, 2);  // src[1] is the NEW o[2]
            o[0] = gmix(src[-1], 1, t, 3, 2, 2);
        } else if (sx == 3) {
            int *src = o + 1, t;
            if (xmod == 0) {
                o[7] = src[2];
                o[6] = gmix(src[1], 1, src[2], 3, 2, 2);
                o[5] = gmix(src[2], 1, src[1], 3, 1, 2);
                o[4] = src[1];
                o[3] = gmix(src[0], 1, src[1], 3, 2, 2);
                o[2] = gmix(src[1], 1, src[0], 3, 1, 2);
                o[0] = gmix(src[-1], 1, src[0], 3, 2, 2);
                o[1] = src[0];
            } else if (xmod == 1) {
                o[7] = gmix(src[3], 1, src[2], 3, 1, 2);
                o[6] = src[2];
                o[5] = gmix(src[1], 1, src[2], 3, 2, 2);
                o[4] = gmix(src[2], 1, src[1], 3, 1, 2);
                o[3] = src[1];
                t = src[0];
                o[2] = gmix(t, 1, src[1], 3, 2, 2);
                o[1] = gmix(src[1], 1, t, 3, 1, 2);
                o[0] = t;
            } else {
                o[7] = gmix(src[2], 1, src[3], 3, 2, 2);
                o[6] = gmix(src[3], 1, src[2], 3, 1, 2);
                o[5] = src[2];
                o[4] = gmix(src[1], 1, src[2], 3, 2, 2);
                o[3] = gmix(src[2], 1, src[1], 3, 1, 2);
                o[2] = src[1];
                t = src[0];
                o[1] = gmix(t, 1, src[1], 3, 2, 2);
                o[0] = gmix(src[1], 1, t, 3, 1, 2);
            }
        } else if (sx == 4) {
            int *src = o + 1, t;
            o[7] = gmix(src[2], 3, src[1], 5, 1, 3);
            o[6] = gmix(src[2], 1, src[1], 7, 2, 3);
            o[5] = gmix(src[0], 1, src[1], 7, 1, 3);
            o[4] = gmix(src[0], 3, src[1], 5, 2, 3);
            t = src[0];
            o[3] = gmix(src[1], 3, t, 5, 1, 3);
            o[2] = gmix(src[1], 1, t, 7, 2, 3);
            o[1] = gmix(src[-1], 1, t, 7, 1, 3);
            o[0] = gmix(src[-1], 3, t, 5, 2, 3);
        }
    }
}

// grid (output block columns / 64, output block rows, frames)
__global__ void __launch_bounds__(64)
generic_reconstruct_kernel(const FrameRecon *__restrict__ frames, const int32_t *__restrict__ samples, uint8_t *__restrict__ out) {
    const FrameRecon &f = frames[blockIdx.z];
    const uint32_t W = f.width, H = f.height, nc = f.ncomp;
    const uint32_t bx = blockIdx.x * 64 + threadIdx.x, by = blockIdx.y;
    if (bx >= (W + 7) / 8 || by >= (H + 7) / 8) return;
    const int X = 8 * (int)bx, Y = 8 * (int)by;
    // 64 upsampled samples per subsampled component (local memory: this path is not tuned); a component at full resolution is
    // read where it lies -- inside the image its sample plane IS the upsampled image
    int buf[4][64];
    const int32_t *direct[4] = {nullptr, nullptr, nullptr, nullptr};
    for (uint32_t c = 0; c < nc; c++) {
        const int sx = f.csx[c], sy = f.csy[c];
        const int w = (int)((W + sx - 1) / sx), h = (int)((H + sy - 1) / sy);
        if (sx == 1 && sy == 1) direct[c] = samples + f.sample_base[c] + (uint64_t)Y * (8u * f.bw[c]) + (uint64_t)X;
        else generic_upsample_block(samples + f.sample_base[c], 8u * f.bw[c], w, h, sx, sy, X, Y, buf[c]);
    }
    auto base_at = [&](uint32_t c, int x, int y) -> int { return direct[c] ? direct[c][(uint64_t)y * (8u * f.bw[c]) + (uint32_t)x] : buf[c][8 * y + x]; };
    const int xmax = (X + 7 < (int)W) ? 7 : (int)((W - 1) & 7), ymax = (Y + 7 < (int)H) ? 7 : (int)((H - 1) & 7);
    if (f.xt & 1u) {
        // ---- JPEG XT: base image + residual image, YCbCrTrafo<..., Residual | Extended | ClampFlag, ...>::YCbCr2RGB
        // colortrafo/ycbcrtrafo.cpp:747-880 with the tables of the 8-bit integer profile (colortransformerfactory.cpp:300-352,
        // 425-520; ParametricToneMappingBox::ScaledTableOf boxes/parametrictonemappingbox.cpp:387-426): Q = identity over the
        // pre-shifted range 0..4095, R2 = floor(i / 16 + 0.5), L = identity over 0..255, C = identity; a lookup clamps its index
        int rbuf[3][64];
        const int32_t *rdirect[3] = {nullptr, nullptr, nullptr};
        for (uint32_t c = 0; c < nc; c++) {
            const int sx = f.res_csx[c], sy = f.res_csy[c];
            const int w = (int)((W + sx - 1) / sx), h = (int)((H + sy - 1) / sy);
            if (sx == 1 && sy == 1) rdirect[c] = samples + f.res_sample_base[c] + (uint64_t)Y * (8u * f.res_bw[c]) + (uint64_t)X;
            else generic_upsample_block(samples + f.res_sample_base[c], 8u * f.res_bw[c], w, h, sx, sy, X, Y, rbuf[c]);
        }
        auto res_at = [&](uint32_t c, int x, int y) -> int { return rdirect[c] ? rdirect[c][(uint64_t)y * (8u * f.res_bw[c]) + (uint32_t)x] : rbuf[c][8 * y + x]; };
        auto clamp_to = [](long long v, long long mx) { return v < 0 ? 0ll : (v > mx ? mx : v); };
        for (int y = 0; y <= ymax; y++) {
            uint8_t *o = out + f.out_base + ((uint64_t)(Y + y) * W + (uint64_t)X) * nc;
            for (int x = 0; x <= xmax; x++) {
                const int i = 8 * y + x;
                long long res[3] = {128, 128, 128}, v[3] = {0, 0, 0};
                if (nc == 3 && (f.xt & 4u)) {  // the residual: Q table, R transformation (FIX_COLOR_TO_INTCOLOR), R2 table
                    const long long yv = clamp_to(res_at(0, x, y), 4095), cb = clamp_to(res_at(1, x, y), 4095) - (128 << 4), cr = clamp_to(res_at(2, x, y), 4095) - (128 << 4);
                    res[0] = (yv * 8192 + cr * 11485 + 4096) >> 13;
                    res[1] = (yv * 8192 - cb * 2819 - cr * 5850 + 4096) >> 13;
                    res[2] = (yv * 8192 + cb * 14516 + 4096) >> 13;
                } else {
                    for (uint32_t c = 0; c < nc; c++) res[c] = clamp_to(res_at(c, x, y), 4095);
                }
                if (nc == 3 && (f.xt & 2u)) {  // the base image: L transformation (FIX_COLOR_TO_INT), L table
                    const long long yv = base_at(0, x, y), cb = (long long)base_at(1, x, y) - (128 << 4), cr = (long long)base_at(2, x, y) - (128 << 4);
                    v[0] = (yv * 8192 + cr * 11485 + 65536) >> 17;
                    v[1] = (yv * 8192 - cb * 2819 - cr * 5850 + 65536) >> 17;
                    v[2] = (yv * 8192 + cb * 14516 + 65536) >> 17;
                } else {
                    for (uint32_t c = 0; c < nc; c++) v[c] = ((long long)base_at(c, x, y) + 8) >> 4;
                }
                for (uint32_t c = 0; c < nc; c++)  // merge and clamp (:863-878, :935-947)
                    o[nc * x + c] = (uint8_t)clamp_to(clamp_to(v[c], 255) + ((clamp_to(res[c], 4095) + 8) >> 4) - 128, 255);
            }
        }
        return;
    }
    // level shift, chroma offset and clamp scale with the precision; 12-bit frames leave as native-endian 16-bit samples
    // (what the reference writes into CTYP_UWORD bitmaps)
    const bool deep = f.precision > 8;
    const long long maxval = (1ll << f.precision) - 1, coff = 16ll << (f.precision - 1u);
    for (int y = 0; y <= ymax; y++) {
        const uint64_t at = ((uint64_t)(Y + y) * W + (uint64_t)X) * nc;
        uint8_t *o8 = out + f.out_base + at;
        uint16_t *o16 = reinterpret_cast<uint16_t *>(out + f.out_base) + at;
        for (int x = 0; x <= xmax; x++) {
            long long v[4];
            if (nc == 3 && f.ycbcr) {
                const long long yv = base_at(0, x, y), cb = (long long)base_at(1, x, y) - coff, cr = (long long)base_at(2, x, y) - coff;
                v[0] = (yv * 8192 + cr * 11485 + 65536) >> 17;
                v[1] = (yv * 8192 - cb * 2819 - cr * 5850 + 65536) >> 17;
                v[2] = (yv * 8192 + cb * 14516 + 65536) >> 17;
            } else {
                for (uint32_t c = 0; c < nc; c++) v[c] = ((long long)base_at(c, x, y) + 8) >> 4;
            }
            for (uint32_t c = 0; c < nc; c++) {
                const long long w = v[c] < 0 ? 0 : (v[c] > maxval ? maxval : v[c]);
                if (deep) o16[nc * x + c] = (uint16_t)w;
                else o8[nc * x + c] = (uint8_t)w;
            }
        }
    }
}

// B200JPG_FLAG_NO_UPSAMPLE: component blockIdx.z of frame blockIdx.y, one thread per sample of its true (subsampled) size:
// COLOR_TO_INT and the clamp of the identity transformation (tools/numerics.hpp:69), nothing else
__global__ void __launch_bounds__(256)
planes_out_kernel(const FrameRecon *__restrict__ frames, const int32_t *__restrict__ samples, uint8_t *__restrict__ out) {
    const FrameRecon &f = frames[blockIdx.y];
    const uint32_t c = blockIdx.z;
    if (c >= f.ncomp) return;
    uint64_t before = 0;  // samples of the planes in front of this one
    for (uint32_t k = 0; k < c; k++) before += (uint64_t)((f.width + f.csx[k] - 1) / f.csx[k]) * ((f.height + f.csy[k] - 1) / f.csy[k]);
    const uint32_t w = (f.width + f.csx[c] - 1) / f.csx[c], h = (f.height + f.csy[c] - 1) / f.csy[c];
    const int maxval = (1 << f.precision) - 1;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < (uint64_t)w * h; i += (uint64_t)gridDim.x * 256) {
        const uint32_t x = (uint32_t)(i % w), y = (uint32_t)(i / w);
        int v = (samples[f.sample_base[c] + (uint64_t)y * (8u * f.bw[c]) + x] + 8) >> 4;
        v = v < 0 ? 0 : (v > maxval ? maxval : v);
        if (f.precision > 8) reinterpret_cast<uint16_t *>(out + f.out_base)[before + i] = (uint16_t)v;
        else out[f.out_base + before + i] = (uint8_t)v;
    }
}

}  // namespace

template <typename T, bool kListed>
static void launch_b2(const ReconLaunch &l, dim3 grid, const T *samples, cudaStream_t s) {
    if (l.ncomp == 1) {
        reconstruct_kernel<1, 1, 1, T, kListed><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, samples, l.wide_flags, l.narrow_list, l.out);
    } else if (l.subx == 2 && l.suby == 2) {
        reconstruct_kernel<3, 2, 2, T, kListed><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, samples, l.wide_flags, l.narrow_list, l.out);
    } else if (l.subx == 2 && l.suby == 1) {
        reconstruct_kernel<3, 2, 1, T, kListed><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, samples, l.wide_flags, l.narrow_list, l.out);
    } else if (l.subx == 1 && l.suby == 2) {
        reconstruct_kernel<3, 1, 2, T, kListed><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, samples, l.wide_flags, l.narrow_list, l.out);
    } else {
        reconstruct_kernel<3, 1, 1, T, kListed><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, samples, l.wide_flags, l.narrow_list, l.out);
    }
}

// seg_rows: chroma block rows per CTA of the fused kernel -- as long as possible (a segment recomputes one chroma block row
// of its upper neighbour) while the grid still holds several waves of CTAs
static uint32_t fused_seg_rows(const ReconLaunch &l, uint32_t strips, uint32_t rows_all) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (uint64_t)sms * B200JPG_FUSED_CTAS * 6u;  // six waves
    uint32_t seg = 16;
    while (seg > 1 && (uint64_t)strips * ((rows_all + seg - 1) / seg) * l.n_frames < want) seg >>= 1;
    return seg;
}

static int launch_recon_fused420(const ReconLaunch &l, cudaStream_t s, int *launches) {
    const uint32_t strips = (l.max_bw0 + 31) / 32, rows_all = (l.max_bh0 + 1) / 2;
    const uint32_t seg = fused_seg_rows(l, strips, rows_all);
    const uint32_t gy = (rows_all + seg - 1) / seg;
    reconstruct420_kernel<int16_t, false><<<dim3(strips, gy, l.n_frames), 32, 0, s>>>(l.frames, l.coef, l.narrow_flags, nullptr, l.out, seg);
    // the exact pass over the frames flagged `narrow` (none for real images: two near-empty launches)
    narrow_list_kernel<<<1, 256, 0, s>>>(l.frames, l.n_frames, l.narrow_flags, l.narrow_list);
    reconstruct420_kernel<int32_t, true><<<dim3(strips, gy, kWideSlots), 32, 0, s>>>(l.frames, l.coef, l.narrow_flags, l.narrow_list, l.out, seg);
    if (launches) *launches = 3;
    return (int)cudaGetLastError();
}

static int launch_recon_generic(const ReconLaunch &l, cudaStream_t s, int *launches) {
    uint32_t mb = 0;
    mb = l.max_bwc * l.max_bhc;  // the largest component block grid of the group (set by the host for generic groups)
    const uint32_t cblocks = (mb + kThreadsB - 1) / kThreadsB;
    int n = 0;
    if (l.generic_phase != 2) {
        idct_planes_kernel<int32_t, false><<<dim3(cblocks, l.n_frames, l.ncomp), kThreadsB, 0, s>>>(l.frames, l.coef, l.samples32, l.wide_flags, nullptr, 0);
        n++;
    }
    if (l.generic_phase == 1) {
        if (launches) *launches = n;
        return (int)cudaGetLastError();
    }
    n++;
    if (l.planes_out) {
        const uint32_t gx = std::min<uint32_t>((l.max_bw0 * l.max_bh0 * 64u + 255u) / 256u, 4096u);
        planes_out_kernel<<<dim3(gx, l.n_frames, l.ncomp), 256, 0, s>>>(l.frames, l.samples32, l.out);
    } else {
        generic_reconstruct_kernel<<<dim3((l.max_bw0 + 63) / 64, l.max_bh0, l.n_frames), 64, 0, s>>>(l.frames, l.samples32, l.out);
    }
    if (launches) *launches = n;
    return (int)cudaGetLastError();
}

int launch_recon(const ReconLaunch &l, void *stream, int *launches) {
    cudaStream_t s = (cudaStream_t)stream;
    int n = 0;
    if (l.generic) return launch_recon_generic(l, s, launches);
    // B200JPG_FUSED=1 selects the single-kernel reconstruction of 4:2:0 frames (no sample planes: 51 instead of 66 MB of DRAM
    // traffic per 4K frame and 25 MB less memory per frame, but 19 % more time: the stage is bound by instruction issue, not by
    // HBM, and the fused kernel issues more -- profiles/README.md); the default is the two-kernel path through int16 planes.
    const char *fused = getenv("B200JPG_FUSED");
    if (l.ncomp == 3 && l.subx == 2 && l.suby == 2 && fused && fused[0] == '1') return launch_recon_fused420(l, s, launches);
    const uint32_t cblocks = (l.max_bwc * l.max_bhc + kThreadsB - 1) / kThreadsB;
    const uint32_t gx = (l.max_bw0 + 31) / 32, gy = (l.max_bh0 + kThreadsB / 32 - 1) / (kThreadsB / 32);
    // every frame through the int16 planes, B200JPG_RECON_CHUNK frames at a time (0: all at once): in chunks the planes b1 writes
    // are still in L2 when b2 reads them
    uint32_t chunk = l.n_frames;
    if (const char *e = getenv("B200JPG_RECON_CHUNK")) {
        const long v = atol(e);
        if (v > 0 && (uint32_t)v < chunk) chunk = (uint32_t)v;
    }
    for (uint32_t f0 = 0; f0 < l.n_frames; f0 += chunk) {
        const uint32_t nf = (l.n_frames - f0 < chunk) ? (l.n_frames - f0) : chunk;
        ReconLaunch c = l;
        c.frames = l.frames + f0;
        if (l.ncomp > 1) {
            idct_planes_kernel<int16_t, false><<<dim3(cblocks, nf, l.ncomp - 1), kThreadsB, 0, s>>>(c.frames, l.coef, l.samples16, l.narrow_flags, nullptr, 1);
            n++;
        }
        launch_b2<int16_t, false>(c, dim3(gx, gy, nf), l.samples16, s);
        n++;
    }
    // the exact pass over the frames that were flagged `narrow` (none for real images: three near-empty launches)
    if (l.ncomp > 1) {
        narrow_list_kernel<<<1, 256, 0, s>>>(l.frames, l.n_frames, l.narrow_flags, l.narrow_list);
        idct_planes_kernel<int32_t, true><<<dim3(cblocks, kWideSlots, l.ncomp - 1), kThreadsB, 0, s>>>(l.frames, l.coef, l.samples32, l.wide_flags, l.narrow_list, 1);
        launch_b2<int32_t, true>(l, dim3(gx, gy, kWideSlots), l.samples32, s);
        n += 3;
    }
    if (launches) *launches = n;
    return (int)cudaGetLastError();
}

}  // namespace b200jpg

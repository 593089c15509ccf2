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
// Mapping.  One thread owns one 8x8 block: the 64 coefficients live in registers through both 1-D passes, so the
// transform needs no shared memory, no shuffles and no synchronisation.  Kernel b1 transforms the blocks of the
// non-luma components into int32 sample planes (the whole-frame equivalent of the reference's upsampler line
// buffers).  Kernel b2 transforms one luma block per thread, pulls the matching chroma window from the planes
// (clamped addressing = the reference's edge replication at the true subsampled size), runs the vertical and
// horizontal filter cores and the colour transform in registers and stores 8 x 24 bytes of interleaved RGB.
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.hpp"

namespace b200jpg {
namespace {

constexpr int kThreadsB = 128;

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

// Loads one dequantised int16 block (128 bytes) and leaves the 64 reconstructed samples (4 fractional bits,
// level shift included) in s[row][col].
__device__ __forceinline__ void idct_block(const int16_t *__restrict__ blk, int (&s)[8][8]) {
    const uint4 *src = reinterpret_cast<const uint4 *>(blk);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        uint4 q = __ldg(src + r);
        // coefficient * delta is stored; the multiplier of dct/idct.cpp:105 is delta << 4
        s[r][0] = (int)(short)(q.x & 0xffffu) << 4;
        s[r][1] = (int)(short)(q.x >> 16) << 4;
        s[r][2] = (int)(short)(q.y & 0xffffu) << 4;
        s[r][3] = (int)(short)(q.y >> 16) << 4;
        s[r][4] = (int)(short)(q.z & 0xffffu) << 4;
        s[r][5] = (int)(short)(q.z >> 16) << 4;
        s[r][6] = (int)(short)(q.w & 0xffffu) << 4;
        s[r][7] = (int)(short)(q.w >> 16) << 4;
    }
    s[0][0] = WADD(s[0][0], 128 << 7);  // dcoffset << (preshift + 3), idct.cpp:233,244
#pragma unroll
    for (int r = 0; r < 8; r++) idct8<256, 9>(s[r][0], s[r][1], s[r][2], s[r][3], s[r][4], s[r][5], s[r][6], s[r][7]);
#pragma unroll
    for (int c = 0; c < 8; c++) idct8<2048, 12>(s[0][c], s[1][c], s[2][c], s[3][c], s[4][c], s[5][c], s[6][c], s[7][c]);
}

// ---- b1: non-luma components -> sample planes ---------------------------------------------------------
__global__ void __launch_bounds__(kThreadsB)
idct_planes_kernel(const FrameRecon *__restrict__ frames, const int16_t *__restrict__ coef, int32_t *__restrict__ samples) {
    const FrameRecon &f = frames[blockIdx.y];
    const int c = 1 + blockIdx.z;
    if (c >= (int)f.ncomp) return;
    const uint32_t bw = f.bw[c], bh = f.bh[c];
    const uint32_t t = blockIdx.x * kThreadsB + threadIdx.x;
    if (t >= bw * bh) return;
    const uint32_t bx = t % bw, by = t / bw;
    int s[8][8];
    idct_block(coef + f.coef_base[c] + (uint64_t)t * 64u, s);
    const uint32_t pitch = 8u * bw;
    int32_t *dst = samples + f.sample_base[c] + (uint64_t)(8u * by) * pitch + 8u * bx;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int4 *d = reinterpret_cast<int4 *>(dst + (uint64_t)r * pitch);
        d[0] = make_int4(s[r][0], s[r][1], s[r][2], s[r][3]);
        d[1] = make_int4(s[r][4], s[r][5], s[r][6], s[r][7]);
    }
}

// ---- b2 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Window of the sample plane of one component for an 8x8 output block: NW columns starting at column x0
// (clamped to [0, cw-1]: dest[-1] = dest[0], dest[width] = dest[width-1], upsamplerbase.cpp:322-323).
template <int NW>
__device__ __forceinline__ void load_row(const int32_t *__restrict__ plane, uint32_t pitch, int y, int x0, int cw, int ch, int (&v)[NW]) {
    const int32_t *row = plane + (uint64_t)clampi(y, 0, ch - 1) * pitch;
#pragma unroll
    for (int j = 0; j < NW; j++) v[j] = __ldg(row + clampi(x0 + j, 0, cw - 1));
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

__device__ __forceinline__ uint32_t clamp255(int v) { return (uint32_t)min(max(v, 0), 255); }
__device__ __forceinline__ uint32_t clamp255_64(long long v) { return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// ycbcrtrafo.cpp:842-850 with the matrix of colortransformerfactory.cpp:136-138 (13 fractional bits) and
// FIX_COLOR_TO_INT (tools/numerics.hpp:65); the reference multiplies in 64 bits, which only matters for samples
// far outside the 8-bit range (damaged streams): those take the 64-bit branch.
__device__ __forceinline__ void ycc_to_rgb(int y, int cbv, int crv, uint32_t &r, uint32_t &g, uint32_t &b) {
    int cb = WSUB(cbv, 128 << 4), cr = WSUB(crv, 128 << 4);
    if ((((unsigned)(y + 32768) | (unsigned)(cb + 32768) | (unsigned)(cr + 32768)) >> 16) == 0) {
        int yy = y * 8192 + 65536;
        r = clamp255((yy + cr * 11485) >> 17);
        g = clamp255((yy - cb * 2819 - cr * 5850) >> 17);
        b = clamp255((yy + cb * 14516) >> 17);
    } else {
        long long Y = y, CB = (long long)cbv - (128 << 4), CR = (long long)crv - (128 << 4);
        r = clamp255_64((Y * 8192 + CR * 11485 + 65536) >> 17);
        g = clamp255_64((Y * 8192 - CB * 2819 - CR * 5850 + 65536) >> 17);
        b = clamp255_64((Y * 8192 + CB * 14516 + 65536) >> 17);
    }
}

template <int NC, int SX, int SY>
__global__ void __launch_bounds__(kThreadsB)
reconstruct_kernel(const FrameRecon *__restrict__ frames, const int16_t *__restrict__ coef, const int32_t *__restrict__ samples,
                   uint8_t *__restrict__ out) {
    const FrameRecon &f = frames[blockIdx.y];
    const uint32_t W = f.width, H = f.height;
    const uint32_t vbw = (W + 7) >> 3, vbh = (H + 7) >> 3;  // blocks that carry visible pixels
    const uint32_t t = blockIdx.x * kThreadsB + threadIdx.x;
    if (t >= vbw * vbh) return;
    const uint32_t bx = t % vbw, by = t / vbw;
    const int X = 8 * bx, Y = 8 * by;

    int s[8][8];
    idct_block(coef + f.coef_base[0] + ((uint64_t)by * f.bw[0] + bx) * 64u, s);

    const int xmax = (X + 7 < (int)W) ? 7 : (int)((W - 1) & 7);
    const int ymax = (Y + 7 < (int)H) ? 7 : (int)((H - 1) & 7);
    const uint32_t opitch = W * NC;
    uint8_t *obase = out + f.out_base + (uint64_t)Y * opitch + (uint64_t)X * NC;

    if (NC == 1) {  // identity, COLOR_TO_INT (numerics.hpp:69) + clamp
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (r > ymax) break;
            uint8_t *o = obase + (uint64_t)r * opitch;
#pragma unroll
            for (int x = 0; x < 8; x++)
                if (x <= xmax) o[x] = (uint8_t)clamp255(WADD(s[r][x], 8) >> 4);
        }
        return;
    }

    constexpr int NW = (SX == 2) ? 6 : 8;
    const int cw = (int)f.cw, ch = (int)f.ch;
    const uint32_t cpitch = 8u * f.bw[1];
    const int32_t *p1 = samples + f.sample_base[1];
    const int32_t *p2 = samples + f.sample_base[2];
    const int cx0 = X / SX - ((SX == 2) ? 1 : 0);  // window column 0 (upsampler.cpp:87,108-109)
    const int cy0 = Y / SY;
    const bool ycbcr = f.ycbcr != 0;

    // rolling rows: SY == 2 keeps top/cur/bot (upsampler.cpp:92-106), SY == 1 only cur
    int top1[NW], cur1[NW], bot1[NW], top2[NW], cur2[NW], bot2[NW];
    if (SY == 2) {
        load_row<NW>(p1, cpitch, cy0 - 1, cx0, cw, ch, top1);
        load_row<NW>(p2, cpitch, cy0 - 1, cx0, cw, ch, top2);
        load_row<NW>(p1, cpitch, cy0 + 1, cx0, cw, ch, bot1);
        load_row<NW>(p2, cpitch, cy0 + 1, cx0, cw, ch, bot2);
    }
    load_row<NW>(p1, cpitch, cy0, cx0, cw, ch, cur1);
    load_row<NW>(p2, cpitch, cy0, cx0, cw, ch, cur2);

#pragma unroll
    for (int r = 0; r < 8; r++) {
        int v1[NW], v2[NW];
        if (SY == 2) {  // VerticalFilterCore<2>, upsampler.cpp:136-168
            if ((r & 1) == 0) {
#pragma unroll
                for (int j = 0; j < NW; j++) {
                    v1[j] = WADD(WADD(top1[j], WMUL(3, cur1[j])), (j & 1) ? 1 : 2) >> 2;
                    v2[j] = WADD(WADD(top2[j], WMUL(3, cur2[j])), (j & 1) ? 1 : 2) >> 2;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NW; j++) {
                    v1[j] = WADD(WADD(bot1[j], WMUL(3, cur1[j])), (j & 1) ? 2 : 1) >> 2;
                    v2[j] = WADD(WADD(bot2[j], WMUL(3, cur2[j])), (j & 1) ? 2 : 1) >> 2;
                }
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
            hfilter2(w1, c1);
            hfilter2(w2, c2);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                c1[j] = v1[j];
                c2[j] = v2[j];
            }
        }
        // colour + store of output row r
        uint32_t px[24];
#pragma unroll
        for (int x = 0; x < 8; x++) {
            uint32_t R, G, B;
            if (ycbcr) {
                ycc_to_rgb(s[r][x], c1[x], c2[x], R, G, B);
            } else {
                R = clamp255(WADD(s[r][x], 8) >> 4);
                G = clamp255(WADD(c1[x], 8) >> 4);
                B = clamp255(WADD(c2[x], 8) >> 4);
            }
            px[3 * x] = R;
            px[3 * x + 1] = G;
            px[3 * x + 2] = B;
        }
        if (r <= ymax) {
            uint8_t *o = obase + (uint64_t)r * opitch;
            if (xmax == 7 && ((reinterpret_cast<uintptr_t>(o) & 7u) == 0)) {
                uint2 *o2 = reinterpret_cast<uint2 *>(o);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    uint32_t lo = px[8 * k] | (px[8 * k + 1] << 8) | (px[8 * k + 2] << 16) | (px[8 * k + 3] << 24);
                    uint32_t hi = px[8 * k + 4] | (px[8 * k + 5] << 8) | (px[8 * k + 6] << 16) | (px[8 * k + 7] << 24);
                    o2[k] = make_uint2(lo, hi);
                }
            } else {
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    if (x <= xmax) {
                        o[3 * x] = (uint8_t)px[3 * x];
                        o[3 * x + 1] = (uint8_t)px[3 * x + 1];
                        o[3 * x + 2] = (uint8_t)px[3 * x + 2];
                    }
                }
            }
        }
        // advance the line window after every odd output line (upsampler.cpp:160-165) / every line for SY == 1
        if (SY == 2) {
            if (r & 1) {
#pragma unroll
                for (int j = 0; j < NW; j++) {
                    top1[j] = cur1[j];
                    cur1[j] = bot1[j];
                    top2[j] = cur2[j];
                    cur2[j] = bot2[j];
                }
                if (r < 7) {
                    load_row<NW>(p1, cpitch, cy0 + (r >> 1) + 2, cx0, cw, ch, bot1);
                    load_row<NW>(p2, cpitch, cy0 + (r >> 1) + 2, cx0, cw, ch, bot2);
                }
            }
        } else if (r < 7) {
            load_row<NW>(p1, cpitch, cy0 + r + 1, cx0, cw, ch, cur1);
            load_row<NW>(p2, cpitch, cy0 + r + 1, cx0, cw, ch, cur2);
        }
    }
}

}  // namespace

int launch_recon(const ReconLaunch &l, void *stream, int *launches) {
    cudaStream_t s = (cudaStream_t)stream;
    int n = 0;
    if (l.ncomp > 1) {
        uint32_t blocks = l.max_bwc * l.max_bhc;
        dim3 grid((blocks + kThreadsB - 1) / kThreadsB, l.n_frames, l.ncomp - 1);
        idct_planes_kernel<<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, l.samples);
        n++;
    }
    uint32_t vblocks = l.max_bw0 * l.max_bh0;
    dim3 grid((vblocks + kThreadsB - 1) / kThreadsB, l.n_frames, 1);
    if (l.ncomp == 1) {
        reconstruct_kernel<1, 1, 1><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, l.samples, l.out);
    } else if (l.subx == 2 && l.suby == 2) {
        reconstruct_kernel<3, 2, 2><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, l.samples, l.out);
    } else if (l.subx == 2 && l.suby == 1) {
        reconstruct_kernel<3, 2, 1><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, l.samples, l.out);
    } else if (l.subx == 1 && l.suby == 2) {
        reconstruct_kernel<3, 1, 2><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, l.samples, l.out);
    } else {
        reconstruct_kernel<3, 1, 1><<<grid, kThreadsB, 0, s>>>(l.frames, l.coef, l.samples, l.out);
    }
    n++;
    if (launches) *launches = n;
    return (int)cudaGetLastError();
}

}  // namespace b200jpg

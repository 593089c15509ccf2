// abi.cpp -- implementation of the C ABI declared in include/b200jpg.h (compiled by nvcc as host C++).
//
// Host side of the decode path: parses the codestreams (parse.cpp), groups scans that share geometry and
// tables into launch classes, lays the batch out in HBM (see internal.hpp) and drives the two CUDA stages.
// There is no CPU decode path in here: without a CUDA device every decode entry point fails.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges show up in Nsight Systems timelines, cost nothing without a tool attached

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "internal.hpp"
#include "specsync.hpp"

namespace {
struct NvtxRange {  // one range per stage of the path
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace

using namespace b200jpg;

namespace {

thread_local std::string g_tls_error;
thread_local int g_tls_code = 0;

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

}  // namespace

static void set_thread_error(int code, const std::string &msg) {
    g_tls_code = code;
    g_tls_error = msg;
}

// Device / pinned-host buffers are recycled through the context: a streaming caller creates and destroys one batch per
// chunk of frames, and cudaMalloc / cudaHostAlloc per chunk would dominate the host side of the pipeline.
struct PoolBuf {
    void *p;
    size_t bytes;
    int kind;  // 0 = pinned host, 1 = device
};

struct b200jpg_ctx {
    int device = 0;
    std::string error;
    int code = 0;
    std::mutex pool_mutex;
    std::vector<PoolBuf> pool;
    void *get(int kind, size_t bytes, cudaError_t *err) {
        *err = cudaSuccess;
        if (bytes == 0) bytes = 256;
        {
            std::lock_guard<std::mutex> lock(pool_mutex);
            int best = -1;
            for (size_t i = 0; i < pool.size(); i++)
                if (pool[i].kind == kind && pool[i].bytes >= bytes && pool[i].bytes <= bytes + bytes / 2 + (1u << 20) &&
                    (best < 0 || pool[i].bytes < pool[(size_t)best].bytes))
                    best = (int)i;
            if (best >= 0) {
                void *p = pool[(size_t)best].p;
                pool.erase(pool.begin() + best);
                return p;
            }
        }
        void *p = nullptr;
        *err = kind ? cudaMalloc(&p, bytes) : cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
        if (*err != cudaSuccess) {  // make room and retry once
            trim();
            *err = kind ? cudaMalloc(&p, bytes) : cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
        }
        return *err == cudaSuccess ? p : nullptr;
    }
    void put(int kind, void *p, size_t bytes) {
        if (!p) return;
        if (bytes == 0) bytes = 256;
        std::lock_guard<std::mutex> lock(pool_mutex);
        pool.push_back(PoolBuf{p, bytes, kind});
    }
    void trim() {
        std::lock_guard<std::mutex> lock(pool_mutex);
        for (auto &b : pool) {
            if (b.kind) cudaFree(b.p);
            else cudaFreeHost(b.p);
        }
        pool.clear();
    }
    int fail(int c, const std::string &m) {
        {
            std::lock_guard<std::mutex> lock(pool_mutex);  // contexts are shared between threads (the JPEG shim)
            code = c;
            error = m;
        }
        set_thread_error(c, m);  // b200jpg_last_error(NULL, ...) on the failing thread always sees its own failure
        return c;
    }
    int fail_cuda(cudaError_t e, const char *what) {
        return fail(B200JPG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    }
};

namespace {

struct ScanClass {
    ScanClassParams p{};
    int table_set = 0;
    std::vector<ClassScan> scans;
    std::vector<uint64_t> interval_off;  // absolute offsets into the packed byte buffer
    std::vector<uint64_t> interval_end;
    std::vector<uint64_t> clean_off;     // offsets into the unstuffed buffer
    std::vector<uint8_t> scan_on_device; // per scan: its restart index is built by restart_index_kernel
    std::vector<uint64_t> scan_ecs_off, scan_ecs_end;  // per scan, absolute
    uint64_t interval_base = 0;          // index of this class's first interval in d_interval_len
    uint64_t spec_base = 0;              // indexed classes: first work item of this class in the spec arrays
    bool fused = false;                  // progressive class decoded by the component-fused kernels (its frames form a PfGroupHost)
    int frame_comp0 = 0;                 // frame component of the scan's first component
    // offsets (bytes) of the device copies inside the input buffer
    uint64_t dev_scans = 0, dev_intervals = 0, dev_interval_end = 0, dev_clean_off = 0;
};

struct ReconGroup {
    uint32_t ncomp = 0, subx = 1, suby = 1;
    bool generic = false;
    std::vector<FrameRecon> frames;
    uint32_t max_bw0 = 0, max_bh0 = 0, max_bwc = 0, max_bhc = 0;
    uint64_t dev_frames = 0;
};

// progressive frames whose scan script lets a component be decoded in one go (progfused_sm100.cu): the classes of its scans
struct PfGroupHost {
    std::vector<int> dc;       // interleaved DC scans, file order
    std::vector<int> ac[4];    // single-component AC scans per frame component, file order
    int comp_of_ac[4] = {0, 0, 0, 0};
};

// the tuned reconstruction kernels cover grey frames and three-component frames whose first component is not subsampled and
// whose two other components share factors of 1 or 2; everything else goes through the generic kernels
static bool frame_is_generic(const b200jpg_frame_info &fi, unsigned flags = 0) {
    if (flags & B200JPG_FLAG_NO_UPSAMPLE) return true;  // planes of every component, then one sample -> one output value
    if (fi.precision != 8) return true;  // 12-bit frames: int32 planes, 16-bit samples out
    if (fi.ncomp == 1) return fi.subx[0] != 1 || fi.suby[0] != 1;
    if (fi.ncomp != 3) return true;
    return fi.subx[0] != 1 || fi.suby[0] != 1 || fi.subx[1] != fi.subx[2] || fi.suby[1] != fi.suby[2] || fi.subx[1] > 2 || fi.suby[1] > 2;
}

struct ClassKey {
    int v[40];
    bool operator<(const ClassKey &o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};

}  // namespace

struct b200jpg_batch {
    b200jpg_ctx *ctx = nullptr;
    int n = 0;
    unsigned flags = 0;  // B200JPG_FLAG_*
    std::vector<ParsedFrame> frames;
    std::vector<int> parse_status;
    int n_user = 0;                       // frames of the caller; [n_user, n) are residual codestreams of JPEG XT frames
    std::vector<int> xt_child, xt_parent;  // [n_user] index of a frame's residual frame or -1; [n] the reverse
    std::vector<uint64_t> out_off, out_bytes;
    uint64_t out_total = 0;
    std::vector<TableSet> table_sets;
    std::vector<uint64_t> dev_tables;  // offset of each table blob inside the input buffer
    std::vector<ScanClass> classes;
    std::vector<ReconGroup> groups;
    std::vector<IndexScan> index_scans;  // scans whose restart index is built on the device, at upload
    uint64_t dev_index_scans = 0;
    std::vector<PfGroupHost> pf_groups;  // progressive frames decoded component by component
    std::vector<uint16_t> pf_dc_quant;   // per group: quantiser of coefficient 0 per frame component (4 each)
    int16_t *d_dcplane = nullptr;        // fused progressive groups: one DC level per block
    std::vector<ProgFrame> prog_frames;  // progressive frames of the scan-by-scan path: dequantised after their last scan
    uint64_t dev_prog_frames = 0;
    uint32_t prog_max_blocks = 0;
    uint64_t ecs_bytes = 0, stored_blocks = 0;
    bool clear_coef = false;  // some component of some frame is coded by no scan: its plane must read as zeros

    // staging / device memory
    uint8_t *h_input = nullptr;  // pinned
    uint64_t input_bytes = 0;    // codestream bytes + descriptors
    uint64_t bytes_region = 0;   // size of the codestream region at the start of the input buffer
    uint8_t *d_input = nullptr;
    int16_t *d_coef = nullptr;
    uint64_t coef_elems = 0;
    int32_t *d_samples = nullptr;     // int32 planes: exact pass of frames flagged `narrow`
    int16_t *d_samples16 = nullptr;   // int16 planes: every frame
    uint64_t sample_elems = 0;
    uint8_t *d_clean = nullptr;
    uint64_t clean_bytes = 0;
    uint32_t *d_interval_len = nullptr;
    uint32_t *d_overrun = nullptr;    // per class {count, interval indices}: class ci starts at interval_base + ci
    uint8_t *d_spec = nullptr;        // restart-less scans: segments, exits, entries, counts, DC sums of all indexed classes
    size_t sz_spec = 0;
    uint64_t n_spec = 0;              // work items over all indexed classes
    size_t sz_overrun = 0;
    uint64_t n_intervals = 0;
    uint32_t *d_status = nullptr;
    std::vector<uint32_t> h_status;
    bool status_fetched = false;
    bool uploaded = false;

    size_t sz_status = 0, sz_ilen = 0;
    cudaEvent_t ev_last = nullptr;  // recorded after the last work enqueued for this batch: destroy waits for it
    int last_launches = 0;
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start, after a0, after a1, after b
};

extern "C" {

// Frames of one source share their tables: the decoder tables of a scan are built once per distinct set of inputs of
// build_table_set -- the scan's Huffman specifications, quantisation tables, point transform and spectral selection -- and copied
// for the other scans that have the same (b200jpg_selftest_table_cache checks the cache against fresh builds on the host).
struct TableCache {
    std::map<std::string, TableSet> built;
};

static std::string table_inputs_key(const ScanInfo &sc) {
    std::string k;
    k.reserve(2048);
    auto put = [&](const void *p, size_t n) { k.append(reinterpret_cast<const char *>(p), n); };
    const int head[6] = {sc.progressive ? 1 : 0, sc.ss, sc.se, sc.ah, sc.lowbit, sc.ns};
    put(head, sizeof(head));
    put(sc.td, sizeof(sc.td));
    put(sc.ta, sizeof(sc.ta));
    for (int pass = 0; pass < 2; pass++)
        for (int t = 0; t < 4; t++) {
            const HuffSpec &h = pass ? sc.ac[t] : sc.dc[t];
            const uint8_t d = h.defined ? 1 : 0;
            put(&d, 1);
            if (!h.defined) continue;
            put(h.bits, 16);
            put(&h.nvals, sizeof(h.nvals));
            put(h.vals, (size_t)std::min(std::max(h.nvals, 0), 256));
        }
    for (int t = 0; t < 4; t++) {
        const uint8_t d = sc.quant_defined[t] ? 1 : 0;
        put(&d, 1);
        if (d) put(sc.quant[t], sizeof(sc.quant[t]));
    }
    return k;
}

static int build_table_set_cached(const ScanInfo &sc, TableSet &out, std::string &err, TableCache &cache) {
    std::string key = table_inputs_key(sc);
    auto it = cache.built.find(key);
    if (it != cache.built.end()) {
        out = it->second;
        return B200JPG_OK;
    }
    const int rc = build_table_set(sc, out, err);
    if (rc == B200JPG_OK) cache.built.emplace(std::move(key), out);
    return rc;
}

// JPEG XT: the residual codestream of `base` (its RESI box) as a frame of its own; what of it the path covers
static int parse_residual(const ParsedFrame &base, ParsedFrame &rf, std::string &err, bool device_index) {
    int rst;
    try {
        rst = parse_codestream(base.xt.resi.data(), base.xt.resi.size(), rf, err, device_index);
    } catch (...) {
        rst = B200JPG_ERR_MALFORMED_STREAM;
        err = "could not be parsed";
    }
    const b200jpg_frame_info &bi = base.info, &ri = rf.info;
    if (rst == 0 && (rf.xt.present || ri.width != bi.width || ri.height != bi.height || ri.ncomp != bi.ncomp || ri.precision != 8)) {
        rst = B200JPG_ERR_NOT_IMPLEMENTED;
        err = "outside the JPEG XT profile of the B200 path (8-bit DCT residual of the frame's size)";
    }
    if (rst != 0) err = "residual codestream: " + err;
    return rst;
}

int b200jpg_parse(const uint8_t *data, size_t len, b200jpg_frame_info *info) {
    ParsedFrame pf;
    std::string err;
    int rc;
    try {
        rc = parse_codestream(data, len, pf, err);
        if (rc == 0 && pf.xt.present) {
            ParsedFrame rf;
            rc = parse_residual(pf, rf, err, false);
        }
    } catch (const std::bad_alloc &) {
        rc = B200JPG_ERR_OUT_OF_MEMORY;
        err = "out of memory while parsing the codestream";
    }
    g_tls_code = rc;
    g_tls_error = err;
    if (info) *info = pf.info;
    return rc;
}

uint64_t b200jpg_build_tables(const uint8_t *data, size_t len, int scan, uint8_t *dst, uint64_t capacity) {
    ParsedFrame pf;
    std::string err;
    TableSet ts;
    int rc;
    try {
        rc = parse_codestream(data, len, pf, err);
        if (rc == 0 && (scan < 0 || scan >= (int)pf.scans.size())) {
            rc = B200JPG_ERR_INVALID_PARAMETER;
            err = "scan index out of range";
        }
        if (rc == 0) rc = build_table_set(pf.scans[scan], ts, err);
    } catch (const std::bad_alloc &) {
        rc = B200JPG_ERR_OUT_OF_MEMORY;
        err = "out of memory while building the decoder tables";
    }
    g_tls_code = rc;
    g_tls_error = err;
    if (rc) return 0;
    if (dst && capacity >= ts.blob.size()) memcpy(dst, ts.blob.data(), ts.blob.size());
    return ts.blob.size();
}

int b200jpg_create(int device, b200jpg_ctx **out) {
    if (!out) return B200JPG_ERR_INVALID_PARAMETER;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_tls_code = B200JPG_ERR_NO_DEVICE;
        g_tls_error = "no CUDA device available: the B200 JPEG decode path has no CPU fallback";
        return B200JPG_ERR_NO_DEVICE;
    }
    if (device < 0) {
        e = cudaGetDevice(&device);
        if (e != cudaSuccess) device = 0;
    }
    if (device >= count) {
        g_tls_code = B200JPG_ERR_INVALID_PARAMETER;
        g_tls_error = "CUDA device ordinal out of range";
        return B200JPG_ERR_INVALID_PARAMETER;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess || prop.major < 10) {
        g_tls_code = B200JPG_ERR_NO_DEVICE;
        g_tls_error = "device is not a Blackwell (sm_100a) GPU: the kernels are built for sm_100a only";
        return B200JPG_ERR_NO_DEVICE;
    }
    auto *c = new b200jpg_ctx();
    c->device = device;
    *out = c;
    return B200JPG_OK;
}

void b200jpg_destroy(b200jpg_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->trim();
    delete ctx;
}

void b200jpg_trim(b200jpg_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->trim();
}

int b200jpg_last_error(b200jpg_ctx *ctx, const char **message) {
    if (ctx) {
        if (message) *message = ctx->error.c_str();
        return ctx->code;
    }
    if (message) *message = g_tls_error.c_str();
    return g_tls_code;
}

void b200jpg_batch_destroy(b200jpg_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    if (b->ev_last) {  // buffers go back to the pool: nothing of this batch may still be in flight
        cudaEventSynchronize(b->ev_last);
        cudaEventDestroy(b->ev_last);
    }
    b->ctx->put(0, b->h_input, b->input_bytes);
    b->ctx->put(1, b->d_input, b->input_bytes);
    b->ctx->put(1, b->d_coef, b->coef_elems * sizeof(int16_t));
    b->ctx->put(1, b->d_samples, b->sample_elems * sizeof(int32_t));
    b->ctx->put(1, b->d_samples16, b->sample_elems * sizeof(int16_t));
    b->ctx->put(1, b->d_clean, b->clean_bytes);
    b->ctx->put(1, b->d_interval_len, b->sz_ilen);
    b->ctx->put(1, b->d_overrun, b->sz_overrun);
    if (b->d_spec) b->ctx->put(1, b->d_spec, b->sz_spec);
    if (b->d_dcplane) b->ctx->put(1, b->d_dcplane, b->coef_elems / 64 * sizeof(int16_t));
    b->ctx->put(1, b->d_status, b->sz_status);
    for (auto &e : b->ev)
        if (e) cudaEventDestroy(e);
    delete b;
}

static int batch_create_impl(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, int tolerate_bad, unsigned flags,
                             b200jpg_batch **out);

int b200jpg_batch_create(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, int tolerate_bad,
                         b200jpg_batch **out) {
    return b200jpg_batch_create_ex(ctx, frames, lens, n, tolerate_bad, 0u, out);
}

int b200jpg_batch_create_ex(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, int tolerate_bad, unsigned flags,
                            b200jpg_batch **out) {
    if (!ctx) return B200JPG_ERR_INVALID_PARAMETER;
    try {  // no exception crosses the C ABI
        return batch_create_impl(ctx, frames, lens, n, tolerate_bad, flags, out);
    } catch (const std::bad_alloc &) {
        if (out) *out = nullptr;
        return ctx->fail(B200JPG_ERR_OUT_OF_MEMORY, "out of host memory while preparing the batch");
    } catch (...) {
        if (out) *out = nullptr;
        return ctx->fail(B200JPG_ERR_INVALID_PARAMETER, "unexpected failure while preparing the batch");
    }
}

static int batch_create_impl(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, int tolerate_bad, unsigned flags,
                             b200jpg_batch **out) {
    NvtxRange nvtx("b200jpg batch_create: parse + pack");
    if (!out || !frames || !lens || n <= 0) return ctx->fail(B200JPG_ERR_INVALID_PARAMETER, "invalid batch arguments");
    *out = nullptr;
    cudaError_t ce = cudaSetDevice(ctx->device);
    if (ce != cudaSuccess) return ctx->fail_cuda(ce, "cudaSetDevice");

    std::unique_ptr<b200jpg_batch, void (*)(b200jpg_batch *)> bp(new b200jpg_batch(), b200jpg_batch_destroy);
    b200jpg_batch *b = bp.get();
    b->ctx = ctx;
    b->n = b->n_user = n;
    b->flags = flags;
    b->frames.resize(n);
    b->parse_status.assign(n, 0);
    std::vector<std::string> errs(n);
    std::vector<const uint8_t *> fp(frames, frames + n);  // codestreams of the batch: the caller's, then the residual
    std::vector<size_t> fl(lens, lens + n);               // codestreams of JPEG XT frames (internal frames behind them)

    // ---- parse (host threads). Interleaved scans with restart markers get their restart index on the device (8f1);
    // B200JPG_HOST_INDEX=1 keeps the memchr pass over every entropy coded byte on the host for all of them.
    const bool device_index = getenv("B200JPG_HOST_INDEX") == nullptr;
    {
        unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        nt = std::min<unsigned>(nt, (unsigned)n);
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; t++)
            pool.emplace_back([&, t]() {
                for (int i = (int)t; i < n; i += (int)nt) {
                    try {  // an exception must not leave the thread (std::terminate): it fails this frame only
                        b->parse_status[i] = parse_codestream(fp[i], fl[i], b->frames[i], errs[i], device_index);
                    } catch (const std::bad_alloc &) {
                        b->parse_status[i] = B200JPG_ERR_OUT_OF_MEMORY;
                        errs[i] = "out of memory while parsing the codestream";
                    } catch (...) {
                        b->parse_status[i] = B200JPG_ERR_MALFORMED_STREAM;
                        errs[i] = "codestream could not be parsed";
                    }
                }
            });
        for (auto &th : pool) th.join();
    }
    // ---- JPEG XT: the residual codestream of a frame is one more frame of the batch -- same entropy kernels, same
    // IDCT -- that owns no output; the parent's reconstruction merges the two (generic_reconstruct_kernel)
    const int n_user = n;
    b->xt_child.assign(n_user, -1);
    for (int i = 0; i < n_user; i++) {
        if (b->parse_status[i] != 0 || !b->frames[i].xt.present) continue;
        if (flags & (B200JPG_FLAG_NO_COLOR_TRANSFORM | B200JPG_FLAG_NO_UPSAMPLE)) {
            b->parse_status[i] = B200JPG_ERR_NOT_IMPLEMENTED;
            errs[i] = "JPEG XT frames are only reconstructed with upsampling and colour transformation";
            continue;
        }
        ParsedFrame rf;
        std::string rerr;
        const int rst = parse_residual(b->frames[i], rf, rerr, device_index);
        if (rst != 0) {
            b->parse_status[i] = rst;
            errs[i] = rerr;
            continue;
        }
        b->xt_child[i] = (int)b->frames.size();
        fp.push_back(b->frames[i].xt.resi.data());  // (the vectors' buffers stay where they are when b->frames grows: moved, not copied)
        fl.push_back(b->frames[i].xt.resi.size());
        b->frames.push_back(std::move(rf));
        b->parse_status.push_back(0);
        errs.emplace_back();
    }
    for (int i = 0; i < n_user; i++)  // push_back may have moved the frames: take the addresses again
        if (b->xt_child[i] >= 0) fp[b->xt_child[i]] = b->frames[i].xt.resi.data();
    n = (int)b->frames.size();
    b->n = n;
    b->xt_parent.assign(n, -1);
    for (int i = 0; i < n_user; i++)
        if (b->xt_child[i] >= 0) b->xt_parent[b->xt_child[i]] = i;
    std::vector<std::vector<TableSet>> frame_tables(n);
    TableCache table_cache;
    for (int i = 0; i < n; i++) {
        ParsedFrame &pf = b->frames[i];
        int &st = b->parse_status[i];
        if (st == 0) {  // what the kernels cover
            const b200jpg_frame_info &fi = pf.info;
            bool factors_ok = fi.ncomp >= 1 && fi.ncomp <= 4;
            for (int c = 0; c < fi.ncomp; c++) factors_ok = factors_ok && fi.subx[c] >= 1 && fi.subx[c] <= 4 && fi.suby[c] >= 1 && fi.suby[c] <= 4;
            if (!factors_ok) {  // the reference's upsampler cores exist for factors 1..4 (upsampling/upsamplerbase.cpp:CreateUpsampler)
                st = B200JPG_ERR_NOT_IMPLEMENTED;
                errs[i] = "only frames of one to four components with subsampling factors of one to four are supported";
            } else {
                // A sequential frame codes a component at most once. A component no scan codes (the stream ends early: the
                // reference just stops at the EOI / the end of the data, Frame::ParseTrailer marker/frame.cpp:1062-1092)
                // keeps zero coefficients: the coefficient store is cleared for such batches.
                int seen[4] = {0, 0, 0, 0};
                for (auto &sc : pf.scans)
                    for (int k = 0; k < sc.ns; k++) seen[sc.comp[k]]++;
                for (int c = 0; c < fi.ncomp; c++) {
                    if (fi.frame_type != 2 && seen[c] > 1) {
                        st = B200JPG_ERR_NOT_IMPLEMENTED;
                        errs[i] = "sequential frame codes a component more than once, not supported by the B200 path";
                    }
                    if (seen[c] == 0) b->clear_coef = true;
                }
            }
        }
        // all table sets of the frame before any of its scans joins a launch class: a frame that fails here must not leave
        // scans behind that decode bytes nobody uploads
        if (st == 0) {
            frame_tables[i].resize(pf.scans.size());
            for (size_t si = 0; si < pf.scans.size() && st == 0; si++) {
                try {
                    st = build_table_set_cached(pf.scans[si], frame_tables[i][si], errs[i], table_cache);
                } catch (const std::bad_alloc &) {
                    st = B200JPG_ERR_OUT_OF_MEMORY;
                    errs[i] = "out of memory while building the decoder tables";
                }
            }
        }
        if (st != 0 && b->xt_parent[i] >= 0) {  // a residual codestream the kernels do not cover fails its frame
            const int par = b->xt_parent[i];
            if (b->parse_status[par] == 0) b->parse_status[par] = st, errs[par] = "residual codestream: " + errs[i];
            b->xt_child[par] = -1;
            if (!tolerate_bad) return ctx->fail(st, "frame " + std::to_string(par) + ": " + errs[par]);
        }
        if (st != 0 && !tolerate_bad) return ctx->fail(st, "frame " + std::to_string(i) + ": " + errs[i]);
    }
    for (int i = 0; i < n; i++)  // ... and a frame that failed takes its residual codestream with it
        if (b->xt_parent[i] >= 0 && b->parse_status[b->xt_parent[i]] != 0 && b->parse_status[i] == 0) b->parse_status[i] = b->parse_status[b->xt_parent[i]];

    // ---- layout of the packed input buffer: [codestreams][table blobs][per class: scans, intervals][per group: frames]
    std::vector<uint64_t> byte_off(n, 0);
    uint64_t cur = 0;
    for (int i = 0; i < n; i++) {
        byte_off[i] = cur;
        if (b->parse_status[i] == 0) cur = align_up(cur + fl[i] + 32, 16);
    }
    b->bytes_region = align_up(cur + 64, 256);

    // coefficient / sample / output layout
    b->out_off.assign(n, 0);
    b->out_bytes.assign(n, 0);
    std::vector<std::array<uint64_t, 4>> coef_base(n), sample_base(n);
    uint64_t coef_cur = 0, sample_cur = 0, out_cur = 0;
    for (int i = 0; i < n; i++) {
        if (b->parse_status[i] != 0) continue;
        const b200jpg_frame_info &fi = b->frames[i].info;
        for (int c = 0; c < fi.ncomp; c++) {
            coef_base[i][c] = coef_cur;
            coef_cur += (uint64_t)fi.blocks_w[c] * fi.blocks_h[c] * 64;
            if (c > 0 || frame_is_generic(fi, flags) || b->xt_parent[i] >= 0 || (i < n_user && b->xt_child[i] >= 0)) {  // generic reconstruction keeps a sample plane of every component
                sample_base[i][c] = sample_cur;
                sample_cur += (uint64_t)fi.blocks_w[c] * fi.blocks_h[c] * 64;
            }
        }
        b->out_off[i] = out_cur;
        uint64_t samples = (uint64_t)fi.width * fi.height * fi.ncomp;
        if (flags & B200JPG_FLAG_NO_UPSAMPLE) {  // plane after plane, every component at its own resolution
            samples = 0;
            for (int c = 0; c < fi.ncomp; c++)
                samples += (uint64_t)((fi.width + fi.subx[c] - 1) / fi.subx[c]) * ((fi.height + fi.suby[c] - 1) / fi.suby[c]);
        }
        b->out_bytes[i] = (b->xt_parent[i] >= 0) ? 0 : samples * (fi.precision > 8 ? 2u : 1u);  // a residual codestream has no pixels of its own
        out_cur = align_up(out_cur + b->out_bytes[i], 256);
        b->ecs_bytes += fi.ecs_bytes;
        b->stored_blocks += fi.stored_blocks;
    }
    b->out_total = out_cur;
    b->coef_elems = coef_cur;
    b->sample_elems = sample_cur;

    // ---- table sets (deduplicated by content) and scan classes
    std::map<std::vector<uint8_t>, int> table_index;
    std::map<ClassKey, int> class_index;
    for (int i = 0; i < n; i++) {
        if (b->parse_status[i] != 0) continue;
        ParsedFrame &pf = b->frames[i];
        const b200jpg_frame_info &fi = pf.info;
        for (size_t si = 0; si < pf.scans.size(); si++) {
            auto &sc = pf.scans[si];
            TableSet &ts = frame_tables[i][si];
            int ti;
            auto it = table_index.find(ts.blob);
            if (it == table_index.end()) {
                ti = (int)b->table_sets.size();
                table_index.emplace(ts.blob, ti);
                b->table_sets.push_back(std::move(ts));
            } else {
                ti = it->second;
            }
            ScanClassParams p{};
            p.ns = sc.ns;
            for (int k = 0; k < sc.ns; k++) {
                int ci = sc.comp[k];
                p.mw[k] = (sc.ns > 1) ? fi.hs[ci] : 1;
                p.mh[k] = (sc.ns > 1) ? fi.vs[ci] : 1;
                p.bw[k] = (int)fi.blocks_w[ci];
                p.dc_slot[k] = sc.td[k];
                p.ac_slot[k] = sc.ta[k];
                p.q_slot[k] = fi.tq[ci];
            }
            p.mcu_cols = sc.mcu_cols;
            p.total_mcus = sc.mcu_cols * sc.mcu_rows;
            p.dri = sc.dri ? sc.dri : p.total_mcus;
            p.intervals_per_scan = (uint32_t)sc.interval_off.size();
            p.lut_words = b->table_sets[ti].lut_words();
            p.progressive = sc.progressive ? 1 : 0;
            p.ss = sc.ss;
            p.se = sc.se;
            p.ah = sc.ah;
            p.al = sc.lowbit;
            p.ordinal = sc.progressive ? (int)si : 0;
            p.indexed = sc.spec ? 1 : 0;
            ClassKey key{};
            int kk = 0;
            key.v[kk++] = p.ns;
            for (int k = 0; k < 4; k++) {
                key.v[kk++] = p.mw[k];
                key.v[kk++] = p.mh[k];
                key.v[kk++] = p.bw[k];
                key.v[kk++] = p.dc_slot[k];
                key.v[kk++] = p.ac_slot[k];
                key.v[kk++] = p.q_slot[k];
            }
            key.v[kk++] = (int)p.mcu_cols;
            key.v[kk++] = (int)p.total_mcus;
            key.v[kk++] = (int)p.dri;
            key.v[kk++] = (int)p.intervals_per_scan;
            key.v[kk++] = ti;
            key.v[kk++] = p.progressive;
            key.v[kk++] = p.ss;
            key.v[kk++] = p.se;
            key.v[kk++] = p.ah;
            key.v[kk++] = p.al;
            key.v[kk++] = p.ordinal;
            key.v[kk++] = p.indexed;
            int cidx;
            auto cit = class_index.find(key);
            if (cit == class_index.end()) {
                cidx = (int)b->classes.size();
                class_index.emplace(key, cidx);
                ScanClass cl;
                cl.p = p;
                cl.table_set = ti;
                b->classes.push_back(std::move(cl));
            } else {
                cidx = cit->second;
            }
            ScanClass &cl = b->classes[cidx];
            cl.frame_comp0 = sc.comp[0];
            ClassScan cs{};
            for (int k = 0; k < sc.ns; k++) cs.coef_base[k] = coef_base[i][sc.comp[k]];
            cs.frame = (uint32_t)i;
            cl.scans.push_back(cs);
            cl.scan_on_device.push_back(sc.device_index ? 1 : 0);
            cl.scan_ecs_off.push_back(byte_off[i] + (uint64_t)sc.ecs_off);
            cl.scan_ecs_end.push_back(byte_off[i] + (uint64_t)sc.ecs_end);
            if (sc.spec) {  // work items per scan: what the longest scan of the class needs (unstuffing only shrinks the data)
                const uint32_t need = (uint32_t)(((uint64_t)(sc.ecs_end - sc.ecs_off) * 8u + kSpecSeqBits - 1u) / kSpecSeqBits) + 1u;
                cl.p.segs_per_scan = std::max(cl.p.segs_per_scan, need);
            }
            for (size_t k = 0; k < sc.interval_off.size(); k++) {
                size_t off = sc.interval_off[k], end = sc.interval_end[k];
                cl.interval_off.push_back(off == SIZE_MAX ? ~0ull : byte_off[i] + (uint64_t)off);
                uint64_t e = off == SIZE_MAX ? 0ull : byte_off[i] + (uint64_t)end;
                // the data ends inside the scan's last interval: flagged for the decoder (kIntervalEofFlag)
                if (sc.eof_tail && off != SIZE_MAX && k + 1 == sc.interval_off.size()) e |= kIntervalEofFlag;
                cl.interval_end.push_back(e);
            }
        }
    }
    // the scans of a progressive frame build on each other: launch order = scan order (sequential classes: ordinal 0)
    std::stable_sort(b->classes.begin(), b->classes.end(), [](const ScanClass &x, const ScanClass &y) { return x.p.ordinal < y.p.ordinal; });
    // ---- component-fused progressive groups: classes that hold exactly the same frames in the same order and whose scans are
    // either interleaved DC scans of all components or single-component AC scans with equal restart intervals per group
    std::vector<uint8_t> frame_fused(n, 0);
    if (getenv("B200JPG_NO_PFUSE") == nullptr) {
        std::map<std::vector<uint32_t>, std::vector<int>> by_frames;
        for (size_t ci = 0; ci < b->classes.size(); ci++) {
            const ScanClass &cl = b->classes[ci];
            if (!cl.p.progressive) continue;
            std::vector<uint32_t> ids;
            for (auto &cs : cl.scans) ids.push_back(cs.frame);
            by_frames[ids].push_back((int)ci);
        }
        for (auto &kv : by_frames) {
            const std::vector<uint32_t> &ids = kv.first;
            std::vector<int> cls = kv.second;  // ascending ordinal (the classes are sorted)
            bool ok = !ids.empty();
            for (uint32_t id : ids) ok = ok && b->frames[id].scans.size() == cls.size();
            PfGroupHost grp;
            const b200jpg_frame_info &fi0 = b->frames[ids[0]].info;
            for (int ci : cls) {
                const ScanClassParams &p = b->classes[(size_t)ci].p;
                if (!ok) break;
                if (p.ss == 0) {
                    ok = p.ns == (int)fi0.ncomp && (grp.dc.empty() ? p.ah == 0 : p.ah != 0);
                    if (ok && !grp.dc.empty()) {
                        const ScanClassParams &q = b->classes[(size_t)grp.dc[0]].p;
                        ok = q.dri == p.dri && q.intervals_per_scan == p.intervals_per_scan && q.total_mcus == p.total_mcus;
                    }
                    grp.dc.push_back(ci);
                } else {
                    const int c = b->classes[(size_t)ci].frame_comp0;
                    ok = p.ns == 1 && c < 4;
                    if (ok && !grp.ac[c].empty()) {
                        const ScanClassParams &q = b->classes[(size_t)grp.ac[c][0]].p;
                        ok = q.dri == p.dri && q.intervals_per_scan == p.intervals_per_scan && q.total_mcus == p.total_mcus && q.mcu_cols == p.mcu_cols;
                    }
                    if (ok) grp.ac[c].push_back(ci);
                }
            }
            ok = ok && !grp.dc.empty() && grp.dc.size() <= (size_t)kPfMaxScans;
            for (int c = 0; c < 4; c++) ok = ok && grp.ac[c].size() <= (size_t)kPfMaxScans;
            if (!ok) continue;
            for (int ci : cls) b->classes[(size_t)ci].fused = true;
            for (uint32_t id : ids) frame_fused[id] = 1;
            const ParsedFrame &pf0 = b->frames[ids[0]];
            for (int c = 0; c < 4; c++)
                b->pf_dc_quant.push_back(c < fi0.ncomp && pf0.scans[0].quant_defined[fi0.tq[c]] ? pf0.scans[0].quant[fi0.tq[c]][0] : 0);
            b->pf_groups.push_back(std::move(grp));
        }
    }
    for (int i = 0; i < n; i++) {
        if (b->parse_status[i] != 0 || b->frames[i].info.frame_type != 2 || frame_fused[i]) continue;
        const ParsedFrame &pf = b->frames[i];
        ProgFrame f{};
        f.frame = (uint32_t)i;
        for (int c = 0; c < pf.info.ncomp; c++) {
            f.coef_base[c] = coef_base[i][c];
            f.n_blocks[c] = pf.info.blocks_w[c] * pf.info.blocks_h[c];
            b->prog_max_blocks = std::max(b->prog_max_blocks, f.n_blocks[c]);
            const ScanInfo &last = pf.scans.back();  // quantisers in effect at the end of the frame
            for (int k = 0; k < 64; k++) f.q_raster[c][kZigZagToRaster[k]] = last.quant_defined[pf.info.tq[c]] ? last.quant[pf.info.tq[c]][k] : 0;
        }
        b->prog_frames.push_back(f);
    }

    // unstuffed-buffer layout. Host-indexed scans: every interval gets its source length rounded up to 16 bytes + 48 bytes
    // of zero tail. Device-indexed scans: one region of ECS length + kCleanSlackPerInterval per interval, inside which the
    // index kernel places interval k at (its source offset) + kCleanSlackPerInterval * k -- no lengths are known here.
    {
        uint64_t ccur = 0, ibase = 0;
        for (size_t ci = 0; ci < b->classes.size(); ci++) {
            auto &cl = b->classes[ci];
            cl.p.n_scans = (uint32_t)cl.scans.size();
            cl.interval_base = ibase;
            cl.clean_off.resize(cl.interval_off.size());
            const size_t nint = cl.p.intervals_per_scan;
            for (size_t j = 0; j < cl.scans.size(); j++) {
                if (cl.scan_on_device[j]) {
                    IndexScan is{};
                    is.ecs_off = cl.scan_ecs_off[j];
                    is.ecs_end = cl.scan_ecs_end[j];
                    is.off_arr = (uint64_t)ci;  // class index for now: becomes the array offset once the layout is known
                    is.end_arr = (uint64_t)(j * nint);
                    is.clean_base = ccur;
                    is.n_intervals = (uint32_t)nint;
                    is.frame = cl.scans[j].frame;
                    b->index_scans.push_back(is);
                    for (size_t k = 0; k < nint; k++) cl.clean_off[j * nint + k] = ccur;
                    ccur += align_up((is.ecs_end - is.ecs_off) + (uint64_t)kCleanSlackPerInterval * nint + 64, 16);
                } else {
                    for (size_t k = j * nint; k < (j + 1) * nint; k++) {
                        cl.clean_off[k] = ccur;
                        uint64_t len = cl.interval_off[k] == ~0ull ? 0 : (cl.interval_end[k] & ~kIntervalEofFlag) - cl.interval_off[k];
                        ccur += align_up(len, 16) + 48;
                    }
                }
            }
            ibase += cl.interval_off.size();
            if (cl.p.indexed) {
                cl.spec_base = b->n_spec;
                b->n_spec += (uint64_t)cl.scans.size() * cl.p.segs_per_scan;
            }
        }
        b->clean_bytes = ccur + 512;  // slack: the ring prefetch runs up to 64 bytes ahead of the reader
        b->n_intervals = ibase;
    }

    // ---- reconstruction groups
    for (int i = 0; i < n; i++) {
        if (b->parse_status[i] != 0) continue;
        const b200jpg_frame_info &fi = b->frames[i].info;
        const int xt_res = (i < n_user) ? b->xt_child[i] : -1;  // JPEG XT: this frame's residual frame
        const bool generic = frame_is_generic(fi, flags) || xt_res >= 0 || b->xt_parent[i] >= 0;
        uint32_t sx = fi.ncomp > 1 ? fi.subx[1] : 1, sy = fi.ncomp > 1 ? fi.suby[1] : 1;
        if (generic) sx = sy = 0;  // one group per component count: the generic kernels read the factors per frame
        ReconGroup *g = nullptr;
        for (auto &gg : b->groups)
            if (gg.ncomp == fi.ncomp && gg.subx == sx && gg.suby == sy && gg.generic == generic) g = &gg;
        if (!g) {
            b->groups.emplace_back();
            g = &b->groups.back();
            g->ncomp = fi.ncomp;
            g->subx = sx;
            g->suby = sy;
            g->generic = generic;
        }
        FrameRecon fr{};
        for (int c = 0; c < fi.ncomp; c++) {
            fr.coef_base[c] = coef_base[i][c];
            fr.sample_base[c] = sample_base[i][c];
            fr.bw[c] = fi.blocks_w[c];
            fr.bh[c] = fi.blocks_h[c];
            fr.csx[c] = fi.subx[c];
            fr.csy[c] = fi.suby[c];
        }
        fr.out_base = b->out_off[i];
        fr.width = fi.width;
        fr.height = fi.height;
        if (b->xt_parent[i] >= 0) fr.width = fr.height = 0;  // a residual codestream: planes only, its frame's reconstruction reads them
        if (xt_res >= 0) {
            const b200jpg_frame_info &ri = b->frames[xt_res].info;
            fr.xt = 1u | (b->frames[i].xt.l_ycbcr ? 2u : 0u) | (b->frames[i].xt.r_ycbcr ? 4u : 0u);
            for (int c = 0; c < ri.ncomp; c++) {
                fr.res_sample_base[c] = sample_base[xt_res][c];
                fr.res_bw[c] = ri.blocks_w[c];
                fr.res_csx[c] = ri.subx[c];
                fr.res_csy[c] = ri.suby[c];
            }
        }
        fr.ncomp = fi.ncomp;
        fr.ycbcr = (flags & B200JPG_FLAG_NO_COLOR_TRANSFORM) ? 0 : fi.ycbcr;  // JPGTAG_MATRIX_LTRAFO = ..._NONE (rectanglerequest.cpp:150-152)
        fr.subx = sx;
        fr.suby = sy;
        fr.cw = sx ? (fi.width + sx - 1) / sx : fi.width;  // (generic groups carry the factors per component: csx / csy)
        fr.ch = sy ? (fi.height + sy - 1) / sy : fi.height;
        fr.status_idx = (uint32_t)i;
        fr.precision = fi.precision;
        g->frames.push_back(fr);
        g->max_bw0 = std::max(g->max_bw0, (fi.width + 7) / 8);
        g->max_bh0 = std::max(g->max_bh0, (fi.height + 7) / 8);
        if (generic) {  // the largest block grid of any component: the extent of the generic IDCT launch
            for (int c = 0; c < fi.ncomp; c++)
                if ((uint64_t)fi.blocks_w[c] * fi.blocks_h[c] > (uint64_t)g->max_bwc * g->max_bhc) g->max_bwc = fi.blocks_w[c], g->max_bhc = fi.blocks_h[c];
        } else if (fi.ncomp > 1) {
            g->max_bwc = std::max(g->max_bwc, fi.blocks_w[1]);
            g->max_bhc = std::max(g->max_bhc, fi.blocks_h[1]);
        }
    }

    // ---- descriptor region
    cur = b->bytes_region;
    b->dev_tables.resize(b->table_sets.size());
    for (size_t t = 0; t < b->table_sets.size(); t++) {
        b->dev_tables[t] = cur;
        cur = align_up(cur + b->table_sets[t].blob.size(), 256);
    }
    for (auto &cl : b->classes) {
        cl.dev_scans = cur;
        cur = align_up(cur + cl.scans.size() * sizeof(ClassScan), 256);
        cl.dev_intervals = cur;
        cur = align_up(cur + cl.interval_off.size() * sizeof(uint64_t), 256);
        cl.dev_interval_end = cur;
        cur = align_up(cur + cl.interval_end.size() * sizeof(uint64_t), 256);
        cl.dev_clean_off = cur;
        cur = align_up(cur + cl.clean_off.size() * sizeof(uint64_t), 256);
    }
    for (auto &g : b->groups) {
        g.dev_frames = cur;
        cur = align_up(cur + g.frames.size() * sizeof(FrameRecon), 256);
    }
    b->dev_prog_frames = cur;
    cur = align_up(cur + b->prog_frames.size() * sizeof(ProgFrame), 256);
    b->dev_index_scans = cur;
    cur = align_up(cur + b->index_scans.size() * sizeof(IndexScan), 256);
    for (auto &is : b->index_scans) {  // class index + first interval -> offsets of the scan's slices
        const ScanClass &cl = b->classes[(size_t)is.off_arr];
        const uint64_t first = is.end_arr * sizeof(uint64_t);
        is.off_arr = cl.dev_intervals + first;
        is.end_arr = cl.dev_interval_end + first;
        is.clean_arr = cl.dev_clean_off + first;
    }
    b->input_bytes = cur;

    // ---- pinned staging + device memory
    b->h_input = (uint8_t *)ctx->get(0, b->input_bytes, &ce);
    if (ce != cudaSuccess) return ctx->fail(B200JPG_ERR_OUT_OF_MEMORY, std::string("pinned staging allocation failed: ") + cudaGetErrorString(ce));
    {
        unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 8u);
        nt = std::min<unsigned>(nt, (unsigned)n);
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; t++)
            pool.emplace_back([&, t]() {
                for (int i = (int)t; i < n; i += (int)nt) {
                    if (b->parse_status[i] != 0) continue;
                    uint8_t *dst = b->h_input + byte_off[i];
                    memcpy(dst, fp[i], fl[i]);
                    // sentinel: a damaged segment must still meet a marker before it leaves its codestream
                    uint64_t pad_end = align_up(byte_off[i] + fl[i] + 32, 16);
                    memset(dst + fl[i], 0, pad_end - (byte_off[i] + fl[i]));
                    dst[fl[i]] = 0xff;
                    dst[fl[i] + 1] = 0xd9;
                }
            });
        for (auto &th : pool) th.join();
    }
    // tail of the codestream region (after the last frame) is zero + a final sentinel
    {
        uint64_t last_end = 0;
        for (int i = 0; i < n; i++)
            if (b->parse_status[i] == 0) last_end = align_up(byte_off[i] + fl[i] + 32, 16);
        memset(b->h_input + last_end, 0, b->bytes_region - last_end);
        b->h_input[last_end] = 0xff;
        b->h_input[last_end + 1] = 0xd9;
    }
    for (size_t t = 0; t < b->table_sets.size(); t++)
        memcpy(b->h_input + b->dev_tables[t], b->table_sets[t].blob.data(), b->table_sets[t].blob.size());
    for (auto &cl : b->classes) {
        memcpy(b->h_input + cl.dev_scans, cl.scans.data(), cl.scans.size() * sizeof(ClassScan));
        memcpy(b->h_input + cl.dev_intervals, cl.interval_off.data(), cl.interval_off.size() * sizeof(uint64_t));
        memcpy(b->h_input + cl.dev_interval_end, cl.interval_end.data(), cl.interval_end.size() * sizeof(uint64_t));
        memcpy(b->h_input + cl.dev_clean_off, cl.clean_off.data(), cl.clean_off.size() * sizeof(uint64_t));
    }
    for (auto &g : b->groups) memcpy(b->h_input + g.dev_frames, g.frames.data(), g.frames.size() * sizeof(FrameRecon));
    if (!b->index_scans.empty()) memcpy(b->h_input + b->dev_index_scans, b->index_scans.data(), b->index_scans.size() * sizeof(IndexScan));
    if (!b->prog_frames.empty()) memcpy(b->h_input + b->dev_prog_frames, b->prog_frames.data(), b->prog_frames.size() * sizeof(ProgFrame));

    b->sz_ilen = sizeof(uint32_t) * (size_t)std::max<uint64_t>(b->n_intervals, 1);
    b->sz_status = sizeof(uint32_t) * (5 * (size_t)n + 1);  // status words, wide flags, narrow flags, narrow list, index status
    b->d_input = (uint8_t *)ctx->get(1, b->input_bytes, &ce);
    if (ce == cudaSuccess && b->coef_elems) b->d_coef = (int16_t *)ctx->get(1, b->coef_elems * sizeof(int16_t), &ce);
    // sample planes only for the reconstruction groups that go through them (with B200JPG_FUSED=1, 4:2:0 frames are reconstructed
    // by the fused kernel: coefficients -> pixels, chroma samples live in shared memory)
    {
        const char *fused = getenv("B200JPG_FUSED");
        const bool use_fused = fused && fused[0] == '1';
        bool need_planes = false;
        for (auto &g : b->groups) need_planes = need_planes || g.generic || (g.ncomp > 1 && !(use_fused && g.ncomp == 3 && g.subx == 2 && g.suby == 2));
        if (!need_planes) b->sample_elems = 0;
    }
    if (ce == cudaSuccess && b->sample_elems) b->d_samples = (int32_t *)ctx->get(1, b->sample_elems * sizeof(int32_t), &ce);
    if (ce == cudaSuccess && b->sample_elems) b->d_samples16 = (int16_t *)ctx->get(1, b->sample_elems * sizeof(int16_t), &ce);
    if (ce == cudaSuccess) b->d_clean = (uint8_t *)ctx->get(1, b->clean_bytes, &ce);
    if (ce == cudaSuccess) b->d_interval_len = (uint32_t *)ctx->get(1, b->sz_ilen, &ce);
    b->sz_overrun = sizeof(uint32_t) * (size_t)(b->n_intervals + b->classes.size() + 1);
    if (ce == cudaSuccess) b->d_overrun = (uint32_t *)ctx->get(1, b->sz_overrun, &ce);
    if (ce == cudaSuccess) b->d_status = (uint32_t *)ctx->get(1, b->sz_status, &ce);
    if (ce == cudaSuccess && !b->pf_groups.empty()) b->d_dcplane = (int16_t *)ctx->get(1, b->coef_elems / 64 * sizeof(int16_t), &ce);
    if (ce == cudaSuccess && b->n_spec) {
        b->sz_spec = (size_t)align_up(b->n_spec, 4) * (sizeof(SpecSegment) + 8 + 8 + 4 + 16 + sizeof(SpecLog)) + 256;
        b->d_spec = (uint8_t *)ctx->get(1, b->sz_spec, &ce);
    }
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&b->ev_last, cudaEventDisableTiming);
    if (ce != cudaSuccess) return ctx->fail(B200JPG_ERR_OUT_OF_MEMORY, std::string("device allocation failed: ") + cudaGetErrorString(ce));
    b->h_status.assign(n, 0);
    *out = bp.release();
    return B200JPG_OK;
}

int b200jpg_batch_frame_info(const b200jpg_batch *b, int i, b200jpg_frame_info *info) {
    if (!b || i < 0 || i >= b->n_user || !info) return B200JPG_ERR_INVALID_PARAMETER;
    *info = b->frames[i].info;
    return b->parse_status[i];
}

uint64_t b200jpg_batch_out_offset(const b200jpg_batch *b, int i) { return (b && i >= 0 && i < b->n_user) ? b->out_off[i] : 0; }
uint64_t b200jpg_batch_out_bytes(const b200jpg_batch *b, int i) {
    if (!b) return 0;
    if (i < 0) return b->out_total;
    return i < b->n_user ? b->out_bytes[i] : 0;
}
uint64_t b200jpg_batch_ecs_bytes(const b200jpg_batch *b) { return b ? b->ecs_bytes : 0; }
uint64_t b200jpg_batch_stored_blocks(const b200jpg_batch *b) { return b ? b->stored_blocks : 0; }
uint64_t b200jpg_batch_h2d_bytes(const b200jpg_batch *b) { return b ? b->input_bytes : 0; }

uint64_t b200jpg_batch_export_tables(const b200jpg_batch *b, uint8_t *dst, uint64_t capacity) {
    if (!b || b->table_sets.size() != 1) return 0;
    const auto &blob = b->table_sets[0].blob;
    if (dst && capacity >= blob.size()) memcpy(dst, blob.data(), blob.size());
    return blob.size();
}

int b200jpg_batch_import_tables(b200jpg_batch *b, const uint8_t *src, uint64_t size) {
    if (!b || !src) return B200JPG_ERR_INVALID_PARAMETER;
    if (b->table_sets.size() != 1) return b->ctx->fail(B200JPG_ERR_INVALID_PARAMETER, "table import needs a batch with exactly one table set");
    auto &blob = b->table_sets[0].blob;
    uint32_t hdr[4];
    if (size < 16) return b->ctx->fail(B200JPG_ERR_INVALID_PARAMETER, "table blob too small");
    memcpy(hdr, src, 16);
    if (hdr[0] != kTableMagic || hdr[1] != size || size != blob.size())
        return b->ctx->fail(B200JPG_ERR_INVALID_PARAMETER, "table blob does not match this batch (different table geometry)");
    memcpy(blob.data(), src, size);
    memcpy(b->h_input + b->dev_tables[0], src, size);
    b->uploaded = false;
    return B200JPG_OK;
}

// the restart index of device-indexed scans (restart_index_kernel): depends only on the uploaded bytes
static int run_restart_index(b200jpg_batch *b, void *stream) {
    uint32_t *index_status = b->d_status + 4 * (size_t)b->n + 1;
    cudaError_t e = cudaMemsetAsync(index_status, 0, sizeof(uint32_t) * (size_t)b->n, (cudaStream_t)stream);
    if (e != cudaSuccess) return b->ctx->fail_cuda(e, "index status reset");
    if (!b->index_scans.empty()) {
        int rc = launch_restart_index(reinterpret_cast<const IndexScan *>(b->d_input + b->dev_index_scans), (uint32_t)b->index_scans.size(),
                                      b->d_input, index_status, stream);
        if (rc != 0) return b->ctx->fail_cuda((cudaError_t)rc, "restart index kernel launch");
    }
    cudaEventRecord(b->ev_last, (cudaStream_t)stream);
    return B200JPG_OK;
}

int b200jpg_batch_upload(b200jpg_batch *b, void *stream) {
    NvtxRange nvtx("b200jpg upload: H2D + restart index");
    if (!b) return B200JPG_ERR_INVALID_PARAMETER;
    cudaSetDevice(b->ctx->device);
    cudaError_t e = cudaMemcpyAsync(b->d_input, b->h_input, b->input_bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return b->ctx->fail_cuda(e, "upload");
    int rc = run_restart_index(b, stream);  // built once per upload, right behind the copy
    if (rc) return rc;
    b->uploaded = true;
    return B200JPG_OK;
}

int b200jpg_batch_reindex(b200jpg_batch *b, void *stream) {
    if (!b) return B200JPG_ERR_INVALID_PARAMETER;
    if (!b->uploaded) return b->ctx->fail(B200JPG_ERR_OBJECT_DOESNT_EXIST, "batch has not been uploaded");
    cudaSetDevice(b->ctx->device);
    return run_restart_index(b, stream);
}

static int run_entropy(b200jpg_batch *b, void *stream) {
    NvtxRange nvtx("b200jpg entropy: unstuff + Huffman decode");
    // the status words start from what the restart index found at upload (out-of-sequence restart markers)
    cudaError_t e = cudaMemcpyAsync(b->d_status, b->d_status + 4 * (size_t)b->n + 1, sizeof(uint32_t) * (size_t)b->n, cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream);
    if (e != cudaSuccess) return b->ctx->fail_cuda(e, "status reset");
    if (!b->prog_frames.empty() || b->clear_coef) {  // progressive scans accumulate into the coefficient store: it starts from zero
        e = cudaMemsetAsync(b->d_coef, 0, b->coef_elems * sizeof(int16_t), (cudaStream_t)stream);
        if (e != cudaSuccess) return b->ctx->fail_cuda(e, "coefficient store reset");
    }
    for (size_t ci = 0; ci < b->classes.size(); ci++) {  // the overrun lists start empty
        e = cudaMemsetAsync(b->d_overrun + b->classes[ci].interval_base + ci, 0, sizeof(uint32_t), (cudaStream_t)stream);
        if (e != cudaSuccess) return b->ctx->fail_cuda(e, "overrun list reset");
    }
    for (int pass = 0; pass < 2; pass++) {  // a0 for every class, then a1 for every class
        if (pass == 1 && b->timing && b->ev[3]) cudaEventRecord(b->ev[3], (cudaStream_t)stream);
        for (size_t ci = 0; ci < b->classes.size(); ci++) {
            auto &cl = b->classes[ci];
            EntropyLaunch l{};
            l.overrun_list = b->d_overrun + cl.interval_base + ci;
            l.p = cl.p;
            l.bytes = b->d_input;
            l.interval_off = reinterpret_cast<const uint64_t *>(b->d_input + cl.dev_intervals);
            l.interval_end = reinterpret_cast<const uint64_t *>(b->d_input + cl.dev_interval_end);
            l.clean_off = reinterpret_cast<const uint64_t *>(b->d_input + cl.dev_clean_off);
            l.clean = b->d_clean;
            l.interval_len = b->d_interval_len + cl.interval_base;
            l.scans = reinterpret_cast<const ClassScan *>(b->d_input + cl.dev_scans);
            l.tables = b->d_input + b->dev_tables[cl.table_set];
            l.coef = b->d_coef;
            l.frame_status = b->d_status;
            if (pass == 1 && cl.fused) continue;  // decoded by the component-fused launches below
            if (cl.p.indexed) {  // slices of the spec arrays: [segments | exits | entries | counts | dc sums], each over all work items
                const uint64_t n = align_up(b->n_spec, 4), o = cl.spec_base;  // (every slice starts 16-byte aligned)
                uint8_t *q = b->d_spec;
                l.spec_segments = reinterpret_cast<SpecSegment *>(q) + o;
                q += n * sizeof(SpecSegment);
                l.spec_exits = reinterpret_cast<unsigned long long *>(q) + o;
                q += n * 8;
                l.spec_entries = reinterpret_cast<unsigned long long *>(q) + o;
                q += n * 8;
                l.spec_counts = reinterpret_cast<uint32_t *>(q) + o;
                q += n * 4;
                l.spec_dc_sums = reinterpret_cast<int32_t *>(q) + 4 * o;
                q += n * 16;
                l.spec_logs = reinterpret_cast<SpecLog *>(q) + o;
                if (pass == 1) {
                    int rs = launch_spec_sync(l, stream);
                    if (rs != 0) return b->ctx->fail_cuda((cudaError_t)rs, "synchronisation kernel launch");
                    b->last_launches++;
                }
            }
            int rc = pass == 0 ? launch_unstuff(l, stream) : (cl.p.progressive ? launch_progressive_scan(l, stream) : launch_entropy(l, stream));
            if (rc != 0) return b->ctx->fail_cuda((cudaError_t)rc, pass == 0 ? "unstuff kernel launch" : "entropy kernel launch");
            b->last_launches++;
            if (pass == 1 && !cl.p.progressive && !cl.p.indexed) {
                rc = launch_overrun_verdict(l, stream);
                if (rc != 0) return b->ctx->fail_cuda((cudaError_t)rc, "overrun verdict kernel launch");
                b->last_launches++;
            }
        }
    }
    // ---- progressive frames decoded component by component: the DC scans into the side plane, then every component's AC scans
    for (size_t gi = 0; gi < b->pf_groups.size(); gi++) {
        const PfGroupHost &grp = b->pf_groups[gi];
        auto fill_scan = [&](PfScan &ps, const ScanClass &cl) {
            ps.tables = b->d_input + b->dev_tables[cl.table_set];
            ps.clean_off = reinterpret_cast<const uint64_t *>(b->d_input + cl.dev_clean_off);
            ps.interval_len = b->d_interval_len + cl.interval_base;
            ps.ss = cl.p.ss, ps.se = cl.p.se, ps.ah = cl.p.ah, ps.al = cl.p.al;
            for (int k = 0; k < 4; k++) ps.dc_slot[k] = cl.p.dc_slot[k];
            ps.ac_slot = cl.p.ac_slot[0];
            ps.q_slot = cl.p.q_slot[0];
            ps.lut_words = cl.p.lut_words;
        };
        auto fill_geometry = [&](PfLaunch &L, const ScanClass &cl) {
            L.ns = cl.p.ns;
            for (int k = 0; k < 4; k++) L.mw[k] = cl.p.mw[k], L.mh[k] = cl.p.mh[k], L.bw[k] = cl.p.bw[k];
            L.mcu_cols = cl.p.mcu_cols, L.total_mcus = cl.p.total_mcus, L.dri = cl.p.dri, L.intervals = cl.p.intervals_per_scan;
            L.n_frames = (uint32_t)cl.scans.size();
            L.frames = reinterpret_cast<const ClassScan *>(b->d_input + cl.dev_scans);
            L.clean = b->d_clean, L.coef = b->d_coef, L.dcplane = b->d_dcplane, L.frame_status = b->d_status;
        };
        {
            PfLaunch L{};
            L.n_scans = (int)grp.dc.size();
            for (size_t s = 0; s < grp.dc.size(); s++) fill_scan(L.scan[s], b->classes[(size_t)grp.dc[s]]);
            const ScanClass &dc0 = b->classes[(size_t)grp.dc[0]];
            fill_geometry(L, dc0);
            // scan component k of the DC scans is frame component comp[k] of the SOS: its quantiser and the block grid that
            // component's own AC scans cover (a component without AC scans is completed by the DC kernel alone)
            const ParsedFrame &pf0 = b->frames[dc0.scans[0].frame];
            for (int k = 0; k < L.ns; k++) {
                const int ci = pf0.scans[(size_t)dc0.p.ordinal].comp[k];
                L.ac_cols[k] = L.ac_rows[k] = 0;
                L.dc_quant[k] = b->pf_dc_quant[4 * gi + (size_t)ci];
                if (!grp.ac[ci].empty()) {
                    const ScanClassParams &ap = b->classes[(size_t)grp.ac[ci][0]].p;
                    L.ac_cols[k] = ap.mcu_cols;
                    L.ac_rows[k] = ap.total_mcus / ap.mcu_cols;
                }
            }
            int rc = launch_pf_dc(L, stream);
            if (rc != 0) return b->ctx->fail_cuda((cudaError_t)rc, "progressive DC kernel launch");
            b->last_launches++;
        }
        for (int c = 0; c < 4; c++) {
            if (grp.ac[c].empty()) continue;
            PfLaunch L{};
            L.n_scans = (int)grp.ac[c].size();
            for (size_t s = 0; s < grp.ac[c].size(); s++) {
                fill_scan(L.scan[s], b->classes[(size_t)grp.ac[c][s]]);
                // scans that decode with the same Huffman tables (the usual case: one AC table per component, or the
                // standard tables) share one copy in shared memory; the quantiser pairs differ with the point transform
                L.scan[s].lut_share = (int)s;
                const std::vector<uint8_t> &mine = b->table_sets[(size_t)b->classes[(size_t)grp.ac[c][s]].table_set].blob;
                for (size_t t = 0; t < s; t++) {
                    const std::vector<uint8_t> &other = b->table_sets[(size_t)b->classes[(size_t)grp.ac[c][t]].table_set].blob;
                    if (L.scan[t].lut_share == (int)t && L.scan[t].lut_words == L.scan[s].lut_words && L.scan[t].ac_slot == L.scan[s].ac_slot &&
                        mine.size() == other.size() && memcmp(mine.data() + 16, other.data() + 16, 16) == 0 &&
                        memcmp(mine.data() + kTableHeaderBytes, other.data() + kTableHeaderBytes, (size_t)L.scan[s].lut_words * 4) == 0) {
                        L.scan[s].lut_share = (int)t;
                        break;
                    }
                }
            }
            fill_geometry(L, b->classes[(size_t)grp.ac[c][0]]);
            L.dc_quant[0] = b->pf_dc_quant[4 * gi + (size_t)c];
            int rc = launch_pf_ac(L, stream);
            if (rc != 0) return b->ctx->fail_cuda((cudaError_t)rc, "progressive AC kernel launch");
            b->last_launches++;
        }
    }
    if (!b->prog_frames.empty()) {  // quantised levels -> the dequantised coefficients stage b expects
        int rc = launch_progressive_dequant(reinterpret_cast<const ProgFrame *>(b->d_input + b->dev_prog_frames), (uint32_t)b->prog_frames.size(),
                                            b->prog_max_blocks, b->d_coef, b->d_status, stream);
        if (rc != 0) return b->ctx->fail_cuda((cudaError_t)rc, "progressive dequantisation kernel launch");
        b->last_launches++;
    }
    cudaEventRecord(b->ev_last, (cudaStream_t)stream);
    return B200JPG_OK;
}

static int run_recon(b200jpg_batch *b, uint8_t *out_dev, void *stream) {
    NvtxRange nvtx("b200jpg reconstruction: IDCT + upsampling + colour");
    cudaError_t me = cudaMemsetAsync(b->d_status + b->n, 0, sizeof(uint32_t) * 2 * (size_t)b->n, (cudaStream_t)stream);
    if (me != cudaSuccess) return b->ctx->fail_cuda(me, "flag reset");
    for (auto &g : b->groups) {
        ReconLaunch l{};
        l.frames = reinterpret_cast<const FrameRecon *>(b->d_input + g.dev_frames);
        l.n_frames = (uint32_t)g.frames.size();
        l.max_bw0 = g.max_bw0;
        l.max_bh0 = g.max_bh0;
        l.max_bwc = g.max_bwc;
        l.max_bhc = g.max_bhc;
        l.ncomp = g.ncomp;
        l.subx = g.subx;
        l.suby = g.suby;
        l.generic = g.generic;
        l.planes_out = (b->flags & B200JPG_FLAG_NO_UPSAMPLE) != 0;
        l.coef = b->d_coef;
        l.samples16 = b->d_samples16;
        l.samples32 = b->d_samples;
        l.wide_flags = b->d_status + b->n;
        l.narrow_flags = b->d_status + 2 * (size_t)b->n;
        l.narrow_list = b->d_status + 3 * (size_t)b->n;
        l.out = out_dev;
        int launches = 0;
        // grid.y carries the frame index: at most 65535 frames per launch. A generic group that is cut runs the IDCT of ALL its
        // frames first: a JPEG XT frame reads the planes of its residual frame, which may sit in another part
        const bool cut = l.n_frames > 65535u;
        for (int phase = (l.generic && cut) ? 1 : 0; phase <= ((l.generic && cut) ? 2 : 0); phase++) {
            for (uint32_t first = 0; first < l.n_frames; first += 65535u) {
                ReconLaunch part = l;
                part.frames = l.frames + first;
                part.n_frames = std::min<uint32_t>(65535u, l.n_frames - first);
                part.generic_phase = phase;
                int rc = launch_recon(part, stream, &launches);
                if (rc != 0) return b->ctx->fail_cuda((cudaError_t)rc, "reconstruction kernel launch");
                b->last_launches += launches;
            }
        }
    }
    cudaEventRecord(b->ev_last, (cudaStream_t)stream);
    return B200JPG_OK;
}

static int ensure_events(b200jpg_batch *b) {
    for (auto &e : b->ev)
        if (!e && cudaEventCreate(&e) != cudaSuccess) return b->ctx->fail(B200JPG_ERR_CUDA, "cudaEventCreate failed");
    return 0;
}

int b200jpg_batch_decode(b200jpg_batch *b, uint8_t *out_dev, void *stream) {
    if (!b || !out_dev) return B200JPG_ERR_INVALID_PARAMETER;
    if (!b->uploaded) return b->ctx->fail(B200JPG_ERR_OBJECT_DOESNT_EXIST, "batch has not been uploaded");
    cudaSetDevice(b->ctx->device);
    b->last_launches = 0;
    b->status_fetched = false;
    if (b->timing) {
        if (ensure_events(b)) return B200JPG_ERR_CUDA;
        cudaEventRecord(b->ev[0], (cudaStream_t)stream);
    }
    int rc = run_entropy(b, stream);
    if (rc) return rc;
    if (b->timing) cudaEventRecord(b->ev[1], (cudaStream_t)stream);
    rc = run_recon(b, out_dev, stream);
    if (rc) return rc;
    if (b->timing) cudaEventRecord(b->ev[2], (cudaStream_t)stream);
    return B200JPG_OK;
}

int b200jpg_batch_decode_entropy(b200jpg_batch *b, void *stream) {
    if (!b) return B200JPG_ERR_INVALID_PARAMETER;
    if (!b->uploaded) return b->ctx->fail(B200JPG_ERR_OBJECT_DOESNT_EXIST, "batch has not been uploaded");
    cudaSetDevice(b->ctx->device);
    b->last_launches = 0;
    b->status_fetched = false;
    return run_entropy(b, stream);
}

int b200jpg_batch_reconstruct(b200jpg_batch *b, uint8_t *out_dev, void *stream) {
    if (!b || !out_dev) return B200JPG_ERR_INVALID_PARAMETER;
    if (!b->uploaded) return b->ctx->fail(B200JPG_ERR_OBJECT_DOESNT_EXIST, "batch has not been uploaded");
    cudaSetDevice(b->ctx->device);
    b->last_launches = 0;
    return run_recon(b, out_dev, stream);
}

int b200jpg_batch_frame_status(b200jpg_batch *b, int i) {
    if (!b || i < 0 || i >= b->n_user) return B200JPG_ERR_INVALID_PARAMETER;
    if (b->parse_status[i] != 0) return b->parse_status[i];
    if (!b->status_fetched) {
        cudaSetDevice(b->ctx->device);
        cudaError_t e = cudaMemcpy(b->h_status.data(), b->d_status, sizeof(uint32_t) * (size_t)b->n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return b->ctx->fail_cuda(e, "status download");
        b->status_fetched = true;
    }
    if (b->h_status[i] == 0 && b->xt_child[i] >= 0) return -(int)b->h_status[b->xt_child[i]];  // JPEG XT: the residual codestream's errors are the frame's
    return -(int)b->h_status[i];
}

int b200jpg_batch_read_coefficients(b200jpg_batch *b, int i, int c, int16_t *dst, uint64_t capacity) {
    if (!b || i < 0 || i >= b->n || !dst) return B200JPG_ERR_INVALID_PARAMETER;
    if (b->parse_status[i] != 0) return b->parse_status[i];
    const b200jpg_frame_info &fi = b->frames[i].info;
    if (c < 0 || c >= fi.ncomp) return B200JPG_ERR_INVALID_PARAMETER;
    uint64_t elems = (uint64_t)fi.blocks_w[c] * fi.blocks_h[c] * 64;
    if (capacity < elems) return B200JPG_ERR_INVALID_PARAMETER;
    // recompute the plane base exactly as batch_create did
    uint64_t base = 0;
    for (int k = 0; k < b->n; k++) {
        if (b->parse_status[k] != 0) continue;
        const b200jpg_frame_info &fk = b->frames[k].info;
        for (int cc = 0; cc < fk.ncomp; cc++) {
            if (k == i && cc == c) goto found;
            base += (uint64_t)fk.blocks_w[cc] * fk.blocks_h[cc] * 64;
        }
    }
found:
    cudaSetDevice(b->ctx->device);
    cudaError_t e = cudaMemcpy(dst, b->d_coef + base, elems * sizeof(int16_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return b->ctx->fail_cuda(e, "coefficient download");
    return B200JPG_OK;
}

int b200jpg_batch_last_launch_count(const b200jpg_batch *b) { return b ? b->last_launches : 0; }

void b200jpg_batch_enable_timing(b200jpg_batch *b, int on) {
    if (b) b->timing = on != 0;
}

float b200jpg_batch_last_unstuff_ms(b200jpg_batch *b) {
    float a = 0;
    if (!b || !b->ev[0] || !b->ev[3]) return -1.f;
    if (cudaEventElapsedTime(&a, b->ev[0], b->ev[3]) != cudaSuccess) return -1.f;
    return a;
}

int b200jpg_batch_last_timing(b200jpg_batch *b, float *entropy_ms, float *reconstruct_ms) {
    if (!b || !b->ev[0]) return B200JPG_ERR_INVALID_PARAMETER;
    float a = 0, c = 0;
    if (cudaEventElapsedTime(&a, b->ev[0], b->ev[1]) != cudaSuccess) return B200JPG_ERR_CUDA;
    if (cudaEventElapsedTime(&c, b->ev[1], b->ev[2]) != cudaSuccess) return B200JPG_ERR_CUDA;
    if (entropy_ms) *entropy_ms = a;
    if (reconstruct_ms) *reconstruct_ms = c;
    return B200JPG_OK;
}

// Host replay of the synchronisation rounds of spec_sync_kernel on scan 0 of one codestream, checked against a plain
// front-to-back walk of the same stream (tests only: it validates the round logic, prefix sums and clipping that the device
// kernel shares through specsync.hpp; nothing in the decode path calls it).
int b200jpg_selftest_restartless(const uint8_t *data, size_t len, uint32_t *rounds, uint32_t *n_segments) {
    ParsedFrame pf;
    std::string err;
    int rc = parse_codestream(data, len, pf, err);
    if (rc) return rc;
    const ScanInfo &sc = pf.scans[0];
    if (sc.progressive || sc.interval_off.size() != 1) return B200JPG_ERR_INVALID_PARAMETER;
    TableSet ts;
    rc = build_table_set(sc, ts, err);
    if (rc) return rc;
    // unstuff (io/bitstream.cpp:56-118) into big-endian words
    std::vector<uint8_t> bytes;
    for (size_t i = sc.ecs_off; i < sc.ecs_end; i++) {
        if (data[i] == 0xff) {
            if (i + 1 < sc.ecs_end && data[i + 1] == 0x00) {
                bytes.push_back(0xff);
                i++;
                continue;
            }
            break;  // a marker
        }
        bytes.push_back(data[i]);
    }
    const uint32_t len_bytes = (uint32_t)bytes.size();
    bytes.resize((bytes.size() + 3) / 4 * 4 + 64, 0);
    std::vector<uint32_t> words(bytes.size() / 4);
    for (size_t i = 0; i < words.size(); i++)
        words[i] = ((uint32_t)bytes[4 * i] << 24) | ((uint32_t)bytes[4 * i + 1] << 16) | ((uint32_t)bytes[4 * i + 2] << 8) | bytes[4 * i + 3];
    const uint16_t *lut_off = reinterpret_cast<const uint16_t *>(ts.blob.data() + 16);
    SpecScan ss{};
    ss.lut = reinterpret_cast<const uint32_t *>(ts.blob.data() + kTableHeaderBytes);
    const b200jpg_frame_info &fi = pf.info;
    for (int c = 0; c < sc.ns; c++) {
        ss.dc_tab[c] = lut_off[sc.td[c]];
        ss.ac_tab[c] = lut_off[4 + sc.ta[c]];
        const int nb = sc.ns > 1 ? fi.hs[sc.comp[c]] * fi.vs[sc.comp[c]] : 1;
        for (int k = 0; k < nb; k++) ss.comp_of_block[ss.blocks_per_mcu++] = (uint8_t)c;
    }
    const uint32_t total_mcus = sc.mcu_cols * sc.mcu_rows;
    std::vector<SpecSegment> segs;
    extern unsigned long long g_spec_replay_bits;
    g_spec_replay_bits = 0;
    const int r = spec_sync_host_replay(ss, words.data(), len_bytes, total_mcus, segs);
    if (getenv("B200JPG_SPEC_STATS"))
        fprintf(stderr, "restart-less replay: %u rounds, %.2f x the stream walked\n", (unsigned)r, (double)g_spec_replay_bits / (8.0 * len_bytes + 1));
    if (rounds) *rounds = (uint32_t)r;
    if (n_segments) *n_segments = (uint32_t)segs.size();
    // the truth: one walk from the first bit, block by block
    const uint32_t total_blocks = total_mcus * ss.blocks_per_mcu, nwords = (len_bytes + 3) / 4;
    std::vector<uint32_t> start_bit(total_blocks + 1);
    std::vector<int32_t> pred_at((size_t)(total_blocks + 1) * 4, 0);
    SpecState st{0, 0};
    int32_t pred[4] = {0, 0, 0, 0};
    uint32_t walked = 0;
    for (; walked < total_blocks && st.bit < len_bytes * 8u; walked++) {
        start_bit[walked] = st.bit;
        for (int c = 0; c < 4; c++) pred_at[(size_t)walked * 4 + c] = pred[c];
        const SpecResult one = spec_decode(ss, words.data(), nwords, len_bytes * 8u, st, st.bit + 1);
        for (int c = 0; c < 4; c++) pred[c] += one.dc_sum[c];
        st = one.exit;
    }
    uint32_t covered = 0;
    for (const SpecSegment &g : segs) {
        if (g.n_blocks == 0) continue;
        if (g.first_block != covered) return -1;                       // the segments tile the blocks in order
        if (g.first_block < walked) {
            if (g.bit != start_bit[g.first_block]) return -2;          // and start where their first block starts
            for (int c = 0; c < sc.ns; c++)
                if (g.pred[c] != pred_at[(size_t)g.first_block * 4 + c]) return -3;  // with the predictors of that place
        }
        covered += g.n_blocks;
    }
    if (covered != total_blocks) return -4;
    return B200JPG_OK;
}

int b200jpg_selftest_table_cache(const uint8_t *const *frames, const size_t *lens, int n) {
    if (!frames || !lens || n <= 0) return B200JPG_ERR_INVALID_PARAMETER;
    TableCache cache;
    int scans_checked = 0;
    for (int pass = 0; pass < 2; pass++)  // twice: the second time every scan comes out of the cache
        for (int i = 0; i < n; i++) {
            ParsedFrame pf;
            std::string err;
            if (parse_codestream(frames[i], lens[i], pf, err) != 0) continue;
            for (const ScanInfo &sc : pf.scans) {
                TableSet cached, fresh;
                std::string e1, e2;
                const int r1 = build_table_set_cached(sc, cached, e1, cache), r2 = build_table_set(sc, fresh, e2);
                if (r1 != r2) return -1;
                if (r1 == 0 && cached.blob != fresh.blob) return -2;
                scans_checked++;
            }
        }
    return scans_checked;
}

int b200jpg_decode_to_host(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, uint8_t *out_host,
                           uint64_t out_capacity) {
    return b200jpg_decode_to_host_ex(ctx, frames, lens, n, out_host, out_capacity, 0u);
}

int b200jpg_decode_to_host_ex(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, uint8_t *out_host,
                              uint64_t out_capacity, unsigned flags) {
    if (!ctx) return B200JPG_ERR_INVALID_PARAMETER;
    b200jpg_batch *b = nullptr;
    int rc = b200jpg_batch_create_ex(ctx, frames, lens, n, 0, flags, &b);
    if (rc) return rc;
    uint8_t *d_out = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t e = cudaSuccess;
    if (out_capacity < b->out_total) {
        rc = ctx->fail(B200JPG_ERR_INVALID_PARAMETER, "output buffer too small");
        goto done;
    }
    e = cudaMalloc((void **)&d_out, b->out_total);
    if (e != cudaSuccess) {
        rc = ctx->fail(B200JPG_ERR_OUT_OF_MEMORY, "device allocation of the output failed");
        goto done;
    }
    e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        rc = ctx->fail_cuda(e, "cudaStreamCreate");
        goto done;
    }
    rc = b200jpg_batch_upload(b, s);
    if (!rc) rc = b200jpg_batch_decode(b, d_out, s);
    if (!rc) {
        e = cudaMemcpyAsync(out_host, d_out, b->out_total, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = ctx->fail_cuda(e, "download");
    }
    if (!rc) {
        for (int i = 0; i < n && !rc; i++) {
            int st = b200jpg_batch_frame_status(b, i);
            if (st) rc = ctx->fail(st, "frame " + std::to_string(i) + ": invalid stream, found invalid huffman code in entropy coded segment");
        }
    }
done:
    if (s) cudaStreamDestroy(s);
    if (d_out) cudaFree(d_out);
    b200jpg_batch_destroy(b);
    return rc;
}

int b200jpg_decode_to_device_ex(b200jpg_ctx *ctx, const uint8_t *const *frames, const size_t *lens, int n, unsigned flags, uint8_t **out_dev,
                                uint64_t *out_bytes) {
    if (!ctx || !out_dev) return B200JPG_ERR_INVALID_PARAMETER;
    *out_dev = nullptr;
    b200jpg_batch *b = nullptr;
    int rc = b200jpg_batch_create_ex(ctx, frames, lens, n, 0, flags, &b);
    if (rc) return rc;
    uint8_t *d_out = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_out, b->out_total ? b->out_total : 1);
    if (e != cudaSuccess) {
        rc = ctx->fail(B200JPG_ERR_OUT_OF_MEMORY, "device allocation of the output failed");
        goto done;
    }
    e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        rc = ctx->fail_cuda(e, "cudaStreamCreate");
        goto done;
    }
    rc = b200jpg_batch_upload(b, s);
    if (!rc) rc = b200jpg_batch_decode(b, d_out, s);
    if (!rc) {
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = ctx->fail_cuda(e, "decode");
    }
    for (int i = 0; i < n && !rc; i++) {
        int st = b200jpg_batch_frame_status(b, i);
        if (st) rc = ctx->fail(st, "frame " + std::to_string(i) + ": invalid stream, found invalid huffman code in entropy coded segment");
    }
    if (!rc) {
        if (out_bytes) *out_bytes = b->out_total;
        *out_dev = d_out;
        d_out = nullptr;
    }
done:
    if (s) cudaStreamDestroy(s);
    if (d_out) cudaFree(d_out);
    b200jpg_batch_destroy(b);
    return rc;
}

void b200jpg_device_free(b200jpg_ctx *ctx, uint8_t *p) {
    if (!p) return;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFree(p);
}

int b200jpg_device_copy_rect(b200jpg_ctx *ctx, uint8_t *dst_dev, int64_t dst_pitch, const uint8_t *src_dev, int64_t src_pitch, uint64_t width_bytes,
                             uint64_t rows) {
    if (!ctx || !dst_dev || !src_dev || dst_pitch < 0 || src_pitch < 0) return B200JPG_ERR_INVALID_PARAMETER;
    if (width_bytes == 0 || rows == 0) return B200JPG_OK;
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMemcpy2D(dst_dev, (size_t)dst_pitch, src_dev, (size_t)src_pitch, (size_t)width_bytes, (size_t)rows, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) return ctx->fail_cuda(e, "device rectangle copy");
    return B200JPG_OK;
}

}  // extern "C"

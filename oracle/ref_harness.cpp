/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Drives the UNMODIFIED reference (thorfdbg/libjpeg, built by oracle/Makefile into _ref/libjpegref.a)
 * through its public interface only -- JPEG::Construct / Read / GetInformation / DisplayRectangle /
 * Destruct with a memory I/O hook and a BitMapHook -- in the access pattern of the reference's own
 * client (8-row stripes from the top, cmd/reconstruct.cpp:312-342).  Uses:
 *   refharness decode <in.jpg> <out.raw> [stripe]     ground-truth pixels (interleaved, depth bytes/pixel)
 *   refharness bench  <in.jpg> <iters> <procs>        CPU baseline: frames/s over `procs` forked workers
 * It is the ground truth for oracle/jpeg_oracle.c and for the CUDA path, and the "reference" CPU arm
 * of bench.py.  Nothing under libjpeg_b200/ links it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

#include "interface/hooks.hpp"
#include "interface/jpeg.hpp"
#include "interface/parameters.hpp"
#include "interface/tagitem.hpp"
#include "tools/traits.hpp"

struct MemStream {
    const unsigned char *data;
    size_t size, pos;
};

static JPG_LONG MemIOHook(struct JPG_Hook *hook, struct JPG_TagItem *tags) {
    MemStream *ms = (MemStream *)hook->hk_pData;
    switch (tags->GetTagData(JPGTAG_FIO_ACTION)) {
    case JPGFLAG_ACTION_READ: {
        unsigned char *buf = (unsigned char *)tags->GetTagPtr(JPGTAG_FIO_BUFFER);
        size_t want = (size_t)tags->GetTagData(JPGTAG_FIO_SIZE);
        size_t left = ms->size - ms->pos;
        if (want > left) want = left;
        memcpy(buf, ms->data + ms->pos, want);
        ms->pos += want;
        return (JPG_LONG)want;
    }
    case JPGFLAG_ACTION_QUERY:
        return 0;
    default:
        return -1; /* no seeking, no writing */
    }
}

struct Canvas {
    unsigned char *mem;
    unsigned width, height, depth;
    unsigned bytes; /* per sample: 1 for precision <= 8 (CTYP_UBYTE), 2 above (CTYP_UWORD) */
};

static JPG_LONG BitmapHookFn(struct JPG_Hook *hook, struct JPG_TagItem *tags) {
    Canvas *cv = (Canvas *)hook->hk_pData;
    if (tags->GetTagData(JPGTAG_BIO_ACTION) == JPGFLAG_BIO_REQUEST) {
        unsigned comp = (unsigned)tags->GetTagData(JPGTAG_BIO_COMPONENT);
        unsigned maxy = (unsigned)tags->GetTagData(JPGTAG_BIO_MAXY);
        /* whole-frame canvas; height rounded up so that the partial last block row is written (the
         * reference reconstructs BIO_HEIGHT >> 3 block rows, control/blockbitmaprequester.cpp:1240) */
        tags->SetTagPtr(JPGTAG_BIO_MEMORY, cv->mem + comp * cv->bytes);
        tags->SetTagData(JPGTAG_BIO_WIDTH, cv->width);
        tags->SetTagData(JPGTAG_BIO_HEIGHT, ((maxy + 8) & ~7u));
        tags->SetTagData(JPGTAG_BIO_BYTESPERROW, cv->width * cv->depth * cv->bytes);
        tags->SetTagData(JPGTAG_BIO_BYTESPERPIXEL, cv->depth * cv->bytes);
        tags->SetTagData(JPGTAG_BIO_PIXELTYPE, cv->bytes == 2 ? CTYP_UWORD : CTYP_UBYTE);
    }
    return 0;
}

static unsigned char *slurp(const char *path, size_t *size) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *d = (unsigned char *)malloc(n + 16);
    if (fread(d, 1, n, f) != (size_t)n) {
        fclose(f);
        free(d);
        return NULL;
    }
    fclose(f);
    *size = (size_t)n;
    return d;
}

static double now(void) {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

/* one full decode through the public API; returns 0 on success. *out is (re)allocated as needed. */
static int decode_once(const unsigned char *data, size_t size, unsigned stripe, Canvas *cv, double *t_read,
                       double *t_disp, int *errcode) {
    MemStream ms = {data, size, 0};
    struct JPG_Hook iohook(MemIOHook, &ms);
    class JPEG *jpeg = JPEG::Construct(NULL);
    int ok = 0;
    if (!jpeg) return -1;
    struct JPG_TagItem rtags[] = {JPG_PointerTag(JPGTAG_HOOK_IOHOOK, &iohook),
                                  JPG_PointerTag(JPGTAG_HOOK_IOSTREAM, &ms),
                                  JPG_ValueTag(JPGTAG_HOOK_BUFFERSIZE, 1 << 20), JPG_EndTag};
    double t0 = now();
    if (jpeg->Read(rtags)) {
        double t1 = now();
        struct JPG_TagItem itags[] = {JPG_ValueTag(JPGTAG_IMAGE_WIDTH, 0), JPG_ValueTag(JPGTAG_IMAGE_HEIGHT, 0),
                                      JPG_ValueTag(JPGTAG_IMAGE_DEPTH, 0), JPG_ValueTag(JPGTAG_IMAGE_PRECISION, 0),
                                      JPG_EndTag};
        if (jpeg->GetInformation(itags)) {
            unsigned w = itags->GetTagData(JPGTAG_IMAGE_WIDTH), h = itags->GetTagData(JPGTAG_IMAGE_HEIGHT);
            unsigned d = itags->GetTagData(JPGTAG_IMAGE_DEPTH);
            unsigned bytes = (itags->GetTagData(JPGTAG_IMAGE_PRECISION) > 8) ? 2 : 1;
            if (!cv->mem || cv->width != w || cv->height != h || cv->depth != d || cv->bytes != bytes) {
                free(cv->mem);
                cv->mem = (unsigned char *)calloc((size_t)w * ((h + 7) & ~7u) * d * bytes, 1);
                cv->width = w;
                cv->height = h;
                cv->depth = d;
                cv->bytes = bytes;
            }
            struct JPG_Hook bmhook(BitmapHookFn, cv);
            struct JPG_TagItem dtags[] = {JPG_PointerTag(JPGTAG_BIH_HOOK, &bmhook), JPG_ValueTag(JPGTAG_DECODER_MINY, 0),
                                          JPG_ValueTag(JPGTAG_DECODER_MAXY, 0), JPG_ValueTag(JPGTAG_DECODER_UPSAMPLE, 1),
                                          JPG_EndTag};
            unsigned y = 0;
            ok = 1;
            if (stripe == 0) stripe = h;
            while (y < h && ok) {
                unsigned last = y + stripe;
                if (last > h) last = h;
                dtags[1].ti_Data.ti_lData = y;
                dtags[2].ti_Data.ti_lData = last - 1;
                ok = jpeg->DisplayRectangle(dtags);
                y = last;
            }
            double t2 = now();
            if (t_read) *t_read += t1 - t0;
            if (t_disp) *t_disp += t2 - t1;
        }
    }
    if (!ok && errcode) {
        const char *msg;
        *errcode = jpeg->LastError(msg);
        fprintf(stderr, "reference error %d: %s\n", *errcode, msg ? msg : "?");
    }
    JPEG::Destruct(jpeg);
    return ok ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc >= 4 && !strcmp(argv[1], "decode")) {
        size_t size;
        unsigned char *data = slurp(argv[2], &size);
        unsigned stripe = (argc > 4) ? (unsigned)atoi(argv[4]) : 8;
        Canvas cv = {NULL, 0, 0, 0, 1};
        int err = 0;
        if (!data) return 2;
        if (decode_once(data, size, stripe, &cv, NULL, NULL, &err)) {
            printf("ERROR %d\n", err);
            return 1;
        }
        FILE *o = fopen(argv[3], "wb");
        fwrite(cv.mem, 1, (size_t)cv.width * cv.height * cv.depth * cv.bytes, o);
        fclose(o);
        if (cv.bytes == 2) printf("%u %u %u 16\n", cv.width, cv.height, cv.depth); /* native-endian 16-bit samples */
        else printf("%u %u %u\n", cv.width, cv.height, cv.depth);
        return 0;
    }
    if (argc >= 5 && !strcmp(argv[1], "bench")) {
        /* files: comma separated list, each worker cycles over it */
        int iters = atoi(argv[3]), procs = atoi(argv[4]);
        char *list = strdup(argv[2]);
        unsigned char *datas[256];
        size_t sizes[256];
        int nfiles = 0;
        for (char *tok = strtok(list, ","); tok && nfiles < 256; tok = strtok(NULL, ",")) {
            datas[nfiles] = slurp(tok, &sizes[nfiles]);
            if (!datas[nfiles]) {
                fprintf(stderr, "cannot read %s\n", tok);
                return 2;
            }
            nfiles++;
        }
        int fds[2];
        if (pipe(fds)) return 2;
        double t0 = now();
        for (int p = 0; p < procs; p++) {
            if (fork() == 0) {
                Canvas cv = {NULL, 0, 0, 0, 1};
                double tr = 0, td = 0;
                int fail = 0;
                for (int i = 0; i < iters; i++) fail |= decode_once(datas[(p + i) % nfiles], sizes[(p + i) % nfiles], 8, &cv, &tr, &td, NULL);
                double rec[3] = {tr, td, (double)fail};
                if (write(fds[1], rec, sizeof(rec)) != (ssize_t)sizeof(rec)) _exit(3);
                _exit(0);
            }
        }
        double tr = 0, td = 0;
        int fail = 0;
        for (int p = 0; p < procs; p++) {
            double rec[3];
            if (read(fds[0], rec, sizeof(rec)) != (ssize_t)sizeof(rec)) return 3;
            tr += rec[0];
            td += rec[1];
            fail |= (int)rec[2];
        }
        while (wait(NULL) > 0) {
        }
        double wall = now() - t0;
        long frames = (long)iters * procs;
        printf("{\"frames\": %ld, \"procs\": %d, \"wall_s\": %.6f, \"fps\": %.3f, \"read_ms_per_frame\": %.3f, "
               "\"display_ms_per_frame\": %.3f, \"failed\": %d}\n",
               frames, procs, wall, frames / wall, 1e3 * tr / frames, 1e3 * td / frames, fail);
        return fail;
    }
    fprintf(stderr, "usage: refharness decode in.jpg out.raw [stripe] | bench a.jpg[,b.jpg..] iters procs\n");
    return 2;
}

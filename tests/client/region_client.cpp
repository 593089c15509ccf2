// region_client.cpp -- a client of the thorfdbg/libjpeg interface that exercises the corners of DisplayRectangle the plain
// stripe client does not: horizontal crops (JPGTAG_DECODER_MINX / MAXX, codestream/rectanglerequest.cpp:93-127), planar client
// bitmaps (BytesPerPixel = 1, one plane per component) and a BitMapHook that reports an error (cmd/bitmaphook.cpp, interface/
// bitmaphook.cpp:196-198).  Written only against interface/*.hpp: it compiles against the REFERENCE's headers + library (that
// is how tests/golden/regions.npz was made, make_regions.py) and against include/ + libb200jpg.so.
//   usage: region_client in.jpg out.raw crop <minx> <maxx> | planar | hookerr <n>
// The canvas (width x round8(height) x depth, pre-filled with 0x5A) is decoded in 8-row stripes from the top, the reference's
// own access pattern, and written to out.raw: interleaved for crop / hookerr, plane after plane for planar.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "interface/hooks.hpp"
#include "interface/jpeg.hpp"
#include "interface/parameters.hpp"
#include "interface/tagitem.hpp"
#ifndef CTYP_UBYTE
#define CTYP_UBYTE 1
#endif

struct Canvas {
    unsigned char *mem;
    unsigned width, height, rows, depth;
    int planar, fail_at, requests;
};

static JPG_LONG FileHook(struct JPG_Hook *hook, struct JPG_TagItem *tags) {
    FILE *in = (FILE *)hook->hk_pData;
    if (tags->GetTagData(JPGTAG_FIO_ACTION) == JPGFLAG_ACTION_READ)
        return (JPG_LONG)fread(tags->GetTagPtr(JPGTAG_FIO_BUFFER), 1, (size_t)tags->GetTagData(JPGTAG_FIO_SIZE), in);
    return tags->GetTagData(JPGTAG_FIO_ACTION) == JPGFLAG_ACTION_QUERY ? 0 : -1;
}

static JPG_LONG CanvasHook(struct JPG_Hook *hook, struct JPG_TagItem *tags) {
    Canvas *cv = (Canvas *)hook->hk_pData;
    if (tags->GetTagData(JPGTAG_BIO_ACTION) != JPGFLAG_BIO_REQUEST) return 0;
    JPG_LONG comp = tags->GetTagData(JPGTAG_BIO_COMPONENT);
    if (++cv->requests == cv->fail_at) return -4711;
    if (cv->planar) {
        tags->SetTagPtr(JPGTAG_BIO_MEMORY, cv->mem + (size_t)comp * cv->width * cv->rows);
        tags->SetTagData(JPGTAG_BIO_BYTESPERROW, cv->width);
        tags->SetTagData(JPGTAG_BIO_BYTESPERPIXEL, 1);
    } else {
        tags->SetTagPtr(JPGTAG_BIO_MEMORY, cv->mem + comp);
        tags->SetTagData(JPGTAG_BIO_BYTESPERROW, cv->width * cv->depth);
        tags->SetTagData(JPGTAG_BIO_BYTESPERPIXEL, cv->depth);
    }
    tags->SetTagData(JPGTAG_BIO_WIDTH, cv->width);
    tags->SetTagData(JPGTAG_BIO_HEIGHT, cv->rows);
    tags->SetTagData(JPGTAG_BIO_PIXELTYPE, CTYP_UBYTE);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    FILE *in = fopen(argv[1], "rb");
    if (!in) return 2;
    const char *mode = argv[3];
    JPG_LONG minx = -1, maxx = -1;
    Canvas cv;
    memset(&cv, 0, sizeof(cv));
    if (!strcmp(mode, "crop") && argc >= 6) minx = atoi(argv[4]), maxx = atoi(argv[5]);
    else if (!strcmp(mode, "planar")) cv.planar = 1;
    else if (!strcmp(mode, "hookerr") && argc >= 5) cv.fail_at = atoi(argv[4]);
    else return 2;
    int rc = 1;
    struct JPG_Hook filehook(FileHook, in);
    class JPEG *jpeg = JPEG::Construct(NULL);
    if (!jpeg) return 2;
    struct JPG_TagItem rtags[] = {JPG_PointerTag(JPGTAG_HOOK_IOHOOK, &filehook), JPG_PointerTag(JPGTAG_HOOK_IOSTREAM, in), JPG_EndTag};
    struct JPG_TagItem itags[] = {JPG_ValueTag(JPGTAG_IMAGE_WIDTH, 0), JPG_ValueTag(JPGTAG_IMAGE_HEIGHT, 0), JPG_ValueTag(JPGTAG_IMAGE_DEPTH, 0),
                                  JPG_EndTag};
    if (jpeg->Read(rtags) && jpeg->GetInformation(itags)) {
        cv.width = (unsigned)itags->GetTagData(JPGTAG_IMAGE_WIDTH);
        cv.height = (unsigned)itags->GetTagData(JPGTAG_IMAGE_HEIGHT);
        cv.depth = (unsigned)itags->GetTagData(JPGTAG_IMAGE_DEPTH);
        cv.rows = (cv.height + 7) & ~7u;
        const size_t bytes = (size_t)cv.width * cv.rows * cv.depth;
        cv.mem = (unsigned char *)malloc(bytes);
        memset(cv.mem, 0x5A, bytes);
        struct JPG_Hook bmhook(CanvasHook, &cv);
        struct JPG_TagItem dtags[] = {JPG_PointerTag(JPGTAG_BIH_HOOK, &bmhook), JPG_ValueTag(JPGTAG_DECODER_MINY, 0), JPG_ValueTag(JPGTAG_DECODER_MAXY, 7),
                                      JPG_ValueTag(minx >= 0 ? JPGTAG_DECODER_MINX : JPGTAG_TAG_IGNORE, minx),
                                      JPG_ValueTag(maxx >= 0 ? JPGTAG_DECODER_MAXX : JPGTAG_TAG_IGNORE, maxx), JPG_EndTag};
        unsigned y = 0;
        int ok = 1;
        while (y < cv.height && ok) {
            unsigned last = y + 8 > cv.height ? cv.height : y + 8;
            dtags[1].ti_Data.ti_lData = (JPG_LONG)y;
            dtags[2].ti_Data.ti_lData = (JPG_LONG)last - 1;
            ok = jpeg->DisplayRectangle(dtags);
            y = last;
        }
        const char *msg = NULL;
        JPG_LONG code = ok ? 0 : jpeg->LastError(msg);
        printf("%u %u %u rows=%u ok=%d error=%d requests=%d\n", cv.width, cv.height, cv.depth, cv.rows, ok, (int)code, cv.requests);
        FILE *out = fopen(argv[2], "wb");
        if (out) {
            fwrite(cv.mem, 1, bytes, out);
            fclose(out);
            rc = 0;
        }
        free(cv.mem);
    } else {
        const char *msg = NULL;
        JPG_LONG code = jpeg->LastError(msg);
        fprintf(stderr, "decode failed: error %d - %s\n", (int)code, msg ? msg : "?");
    }
    JPEG::Destruct(jpeg);
    fclose(in);
    return rc;
}

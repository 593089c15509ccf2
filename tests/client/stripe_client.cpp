// stripe_client.cpp -- an ordinary client of the thorfdbg/libjpeg interface, written only against
// interface/{jpeg,tagitem,hooks,parameters}.hpp.  It compiles unchanged against the REFERENCE's headers + library and
// against this repository's include/ + libb200jpg.so: the drop-in acceptance test.
//
// It follows the access pattern of the reference's own demo client (cmd/reconstruct.cpp:312-342 and
// cmd/bitmaphook.cpp:102-255): file I/O hook, GetInformation, then DisplayRectangle in 8-row stripes into ONE reused
// 8-row buffer whose base pointer is anchored at canvas row 0 (mem - miny * stride), BIO_HEIGHT = miny + 8.
//   usage: stripe_client in.jpg out.pnm
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

struct Stripe {
    unsigned char *mem;  // 8 rows
    unsigned width, height, depth;
    FILE *out;
    int requests, releases;
};

static JPG_LONG FileHook(struct JPG_Hook *hook, struct JPG_TagItem *tags) {
    FILE *in = (FILE *)hook->hk_pData;
    switch (tags->GetTagData(JPGTAG_FIO_ACTION)) {
    case JPGFLAG_ACTION_READ: {
        void *buffer = tags->GetTagPtr(JPGTAG_FIO_BUFFER);
        JPG_LONG size = tags->GetTagData(JPGTAG_FIO_SIZE);
        return (JPG_LONG)fread(buffer, 1, (size_t)size, in);
    }
    case JPGFLAG_ACTION_SEEK:
        return -1;
    case JPGFLAG_ACTION_QUERY:
        return 0;
    }
    return -1;
}

static JPG_LONG StripeHook(struct JPG_Hook *hook, struct JPG_TagItem *tags) {
    Stripe *st = (Stripe *)hook->hk_pData;
    JPG_LONG comp = tags->GetTagData(JPGTAG_BIO_COMPONENT);
    JPG_LONG miny = tags->GetTagData(JPGTAG_BIO_MINY), maxy = tags->GetTagData(JPGTAG_BIO_MAXY);
    if (comp < 0 || comp >= (JPG_LONG)st->depth) return -1;
    switch (tags->GetTagData(JPGTAG_BIO_ACTION)) {
    case JPGFLAG_BIO_REQUEST: {
        const JPG_LONG stride = (JPG_LONG)(st->width * st->depth);
        st->requests++;
        tags->SetTagPtr(JPGTAG_BIO_MEMORY, st->mem + comp - (long)miny * stride);  // canvas-anchored
        tags->SetTagData(JPGTAG_BIO_WIDTH, st->width);
        tags->SetTagData(JPGTAG_BIO_HEIGHT, 8 + miny);
        tags->SetTagData(JPGTAG_BIO_BYTESPERROW, stride);
        tags->SetTagData(JPGTAG_BIO_BYTESPERPIXEL, st->depth);
        tags->SetTagData(JPGTAG_BIO_PIXELTYPE, CTYP_UBYTE);
        break;
    }
    case JPGFLAG_BIO_RELEASE:
        st->releases++;
        if (comp == (JPG_LONG)st->depth - 1)  // last component of the stripe: flush its rows
            fwrite(st->mem, 1, (size_t)st->width * st->depth * (size_t)(maxy - miny + 1), st->out);
        break;
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s in.jpg out.pnm\n", argv[0]);
        return 2;
    }
    FILE *in = fopen(argv[1], "rb");
    if (!in) return 2;
    int rc = 1;
    struct JPG_Hook filehook(FileHook, in);
    class JPEG *jpeg = JPEG::Construct(NULL);
    if (jpeg) {
        struct JPG_TagItem rtags[] = {JPG_PointerTag(JPGTAG_HOOK_IOHOOK, &filehook), JPG_PointerTag(JPGTAG_HOOK_IOSTREAM, in), JPG_EndTag};
        if (jpeg->Read(rtags)) {
            unsigned char subx[4], suby[4];
            struct JPG_TagItem itags[] = {JPG_ValueTag(JPGTAG_IMAGE_WIDTH, 0),      JPG_ValueTag(JPGTAG_IMAGE_HEIGHT, 0),
                                          JPG_ValueTag(JPGTAG_IMAGE_DEPTH, 0),      JPG_ValueTag(JPGTAG_IMAGE_PRECISION, 0),
                                          JPG_PointerTag(JPGTAG_IMAGE_SUBX, subx),  JPG_PointerTag(JPGTAG_IMAGE_SUBY, suby),
                                          JPG_ValueTag(JPGTAG_IMAGE_SUBLENGTH, 4),  JPG_EndTag};
            if (jpeg->GetInformation(itags)) {
                Stripe st;
                st.width = (unsigned)itags->GetTagData(JPGTAG_IMAGE_WIDTH);
                st.height = (unsigned)itags->GetTagData(JPGTAG_IMAGE_HEIGHT);
                st.depth = (unsigned)itags->GetTagData(JPGTAG_IMAGE_DEPTH);
                st.requests = st.releases = 0;
                st.mem = (unsigned char *)calloc((size_t)st.width * st.depth, 8);
                st.out = fopen(argv[2], "wb");
                if (st.mem && st.out) {
                    fprintf(st.out, "P%c\n%u %u\n255\n", st.depth == 1 ? '5' : '6', st.width, st.height);
                    struct JPG_Hook bmhook(StripeHook, &st);
                    struct JPG_TagItem dtags[] = {JPG_PointerTag(JPGTAG_BIH_HOOK, &bmhook), JPG_ValueTag(JPGTAG_DECODER_MINY, 0),
                                                  JPG_ValueTag(JPGTAG_DECODER_MAXY, 7), JPG_ValueTag(JPGTAG_DECODER_UPSAMPLE, 1),
                                                  JPG_EndTag};
                    unsigned y = 0;
                    int ok = 1;
                    while (y < st.height && ok) {
                        unsigned last = y + 8 > st.height ? st.height : y + 8;
                        dtags[1].ti_Data.ti_lData = (JPG_LONG)y;
                        dtags[2].ti_Data.ti_lData = (JPG_LONG)last - 1;
                        ok = jpeg->DisplayRectangle(dtags);
                        y = last;
                    }
                    if (ok) rc = 0;
                    printf("%u %u %u requests=%d releases=%d subx=%d,%d suby=%d,%d\n", st.width, st.height, st.depth, st.requests, st.releases, subx[0],
                           st.depth > 1 ? subx[1] : 0, suby[0], st.depth > 1 ? suby[1] : 0);
                }
                if (st.out) fclose(st.out);
                free(st.mem);
            }
        }
        if (rc) {
            const char *msg = NULL;
            JPG_LONG code = jpeg->LastError(msg);
            fprintf(stderr, "decode failed: error %d - %s\n", (int)code, msg ? msg : "?");
        }
        JPEG::Destruct(jpeg);
    }
    fclose(in);
    return rc;
}

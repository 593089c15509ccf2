// device_client.cpp -- a client of the thorfdbg/libjpeg interface whose bitmap lives in CUDA DEVICE memory: the BitMapHook hands out
// a device pointer and the request carries JPGTAG_B200_DEVICE_BITMAPS, so the decoded pixels never cross PCIe (GPU
// consumers).  8-row stripes from the top like cmd/reconstruct.cpp:312-342; the canvas is copied back at the end only to be
// written to out.raw for the comparison with the reference's pixels.
//   usage: device_client in.jpg out.raw          (interleaved canvas, width x round8(height) x depth bytes, pre-filled with 0x5A)
#include <cuda_runtime.h>
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
    unsigned char *dev;
    unsigned width, rows, depth;
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
    tags->SetTagPtr(JPGTAG_BIO_MEMORY, cv->dev + comp);  // a DEVICE pointer
    tags->SetTagData(JPGTAG_BIO_BYTESPERROW, cv->width * cv->depth);
    tags->SetTagData(JPGTAG_BIO_BYTESPERPIXEL, cv->depth);
    tags->SetTagData(JPGTAG_BIO_WIDTH, cv->width);
    tags->SetTagData(JPGTAG_BIO_HEIGHT, cv->rows);
    tags->SetTagData(JPGTAG_BIO_PIXELTYPE, CTYP_UBYTE);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *in = fopen(argv[1], "rb");
    if (!in) return 2;
    struct JPG_Hook fhook(FileHook, in);
    class JPEG *jpeg = JPEG::Construct(NULL);
    if (!jpeg) return 3;
    struct JPG_TagItem rtags[] = {JPG_PointerTag(JPGTAG_HOOK_IOHOOK, &fhook), JPG_PointerTag(JPGTAG_HOOK_IOSTREAM, in), JPG_EndTag};
    if (!jpeg->Read(rtags)) {
        const char *e = 0;
        fprintf(stderr, "read failed %ld\n", (long)jpeg->LastError(e));
        return 4;
    }
    struct JPG_TagItem itags[] = {JPG_ValueTag(JPGTAG_IMAGE_WIDTH, 0), JPG_ValueTag(JPGTAG_IMAGE_HEIGHT, 0), JPG_ValueTag(JPGTAG_IMAGE_DEPTH, 0), JPG_EndTag};
    if (!jpeg->GetInformation(itags)) return 5;
    Canvas cv;
    cv.width = (unsigned)itags->GetTagData(JPGTAG_IMAGE_WIDTH);
    const unsigned height = (unsigned)itags->GetTagData(JPGTAG_IMAGE_HEIGHT);
    cv.depth = (unsigned)itags->GetTagData(JPGTAG_IMAGE_DEPTH);
    cv.rows = (height + 7) & ~7u;
    const size_t bytes = (size_t)cv.width * cv.rows * cv.depth;
    if (cudaMalloc((void **)&cv.dev, bytes) != cudaSuccess || cudaMemset(cv.dev, 0x5A, bytes) != cudaSuccess) return 6;
    struct JPG_Hook bhook(CanvasHook, &cv);
    int ok = 1;
    for (unsigned y = 0; y < height && ok; y += 8) {
        struct JPG_TagItem dtags[] = {JPG_PointerTag(JPGTAG_BIH_HOOK, &bhook), JPG_ValueTag(JPGTAG_DECODER_MINY, (JPG_LONG)y),
                                      JPG_ValueTag(JPGTAG_DECODER_MAXY, (JPG_LONG)(y + 7 < height ? y + 7 : height - 1)),
                                      JPG_ValueTag(JPGTAG_B200_DEVICE_BITMAPS, 1), JPG_EndTag};
        ok = jpeg->DisplayRectangle(dtags) != 0;
    }
    const char *e = 0;
    const long code = ok ? 0 : (long)jpeg->LastError(e);
    unsigned char *host = (unsigned char *)malloc(bytes);
    if (!host || cudaMemcpy(host, cv.dev, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return 7;
    FILE *out = fopen(argv[2], "wb");
    if (!out) return 2;
    fwrite(host, 1, bytes, out);
    fclose(out);
    printf("%u %u %u ok=%d err=%ld\n", cv.width, height, cv.depth, ok, code);
    JPEG::Destruct(jpeg);
    cudaFree(cv.dev);
    fclose(in);
    return 0;
}

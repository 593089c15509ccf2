// tagitem_check.cpp -- prints the results of JPG_TagItem operations; built against the reference and against this
// repository's implementation, the two transcripts must be identical (tests/test_shim_cpu.py).
#include <stdio.h>

#include "interface/parameters.hpp"
#include "interface/tagitem.hpp"

static void dump(const char *what, const struct JPG_TagItem *t) {
    printf("%s:", what);
    while (t) {
        printf(" %08x=%d", (unsigned)t->ti_Tag, (int)t->ti_Data.ti_lData);
        t = t->NextTagItem();
    }
    printf("\n");
}

int main() {
    struct JPG_TagItem tail[] = {JPG_ValueTag(JPGTAG_IMAGE_DEPTH, 3), JPG_ValueTag(JPGTAG_TAG_IGNORE, 99), JPG_ValueTag(JPGTAG_IMAGE_PRECISION, 8),
                                 JPG_EndTag};
    struct JPG_TagItem list[] = {JPG_ValueTag(JPGTAG_IMAGE_WIDTH, 640),
                                 JPG_ValueTag(JPGTAG_TAG_SKIP, 1),
                                 JPG_ValueTag(JPGTAG_IMAGE_QUALITY, 11),  // skipped
                                 JPG_ValueTag(JPGTAG_IMAGE_HEIGHT, 480),
                                 JPG_ValueTag(JPGTAG_TAG_IGNORE, 5),
                                 JPG_ValueTag(JPGTAG_IMAGE_WIDTH, 641),  // second match: never found
                                 JPG_Continue(tail)};
    dump("walk", list);
    printf("get width %d height %d depth %d prec %d quality(default 7) %d\n", (int)list->GetTagData(JPGTAG_IMAGE_WIDTH),
           (int)list->GetTagData(JPGTAG_IMAGE_HEIGHT), (int)list->GetTagData(JPGTAG_IMAGE_DEPTH), (int)list->GetTagData(JPGTAG_IMAGE_PRECISION),
           (int)list->GetTagData(JPGTAG_IMAGE_QUALITY, 7));
    list->SetTagData(JPGTAG_IMAGE_DEPTH, 4);
    list->SetTagData(JPGTAG_IMAGE_QUALITY, 50);  // absent: silent no-op
    printf("after set depth %d quality %d\n", (int)list->GetTagData(JPGTAG_IMAGE_DEPTH), (int)list->GetTagData(JPGTAG_IMAGE_QUALITY, -1));
    printf("find from item 3: %s\n", list[3].FindTagItem(JPGTAG_IMAGE_WIDTH) == &list[5] ? "second width" : "other");
    printf("ptr default %s\n", list->GetTagPtr(JPGTAG_BIO_MEMORY, (JPG_APTR)list) == (JPG_APTR)list ? "default" : "found");

    struct JPG_TagItem defaults[] = {JPG_ValueTag(JPGTAG_IMAGE_WIDTH, 1), JPG_ValueTag(JPGTAG_IMAGE_QUALITY, 75), JPG_ValueTag(JPGTAG_IMAGE_ERRORBOUND, 2),
                                     JPG_EndTag};
    struct JPG_TagItem drop[] = {JPG_ValueTag(JPGTAG_IMAGE_ERRORBOUND, 0), JPG_EndTag};
    struct JPG_TagItem target[16];
    JPG_LONG n = JPG_TagItem::FilterTags(NULL, list, defaults, drop);
    JPG_LONG m = JPG_TagItem::FilterTags(target, list, defaults, drop);
    printf("filter count %d %d\n", (int)n, (int)m);
    dump("filtered", target);

    struct JPG_TagItem a[] = {JPG_ValueTag(JPGTAG_IMAGE_WIDTH, 1), JPG_EndTag};
    struct JPG_TagItem b[] = {JPG_ValueTag(JPGTAG_IMAGE_HEIGHT, 2), JPG_EndTag};
    struct JPG_TagItem *patched = a->TagOn(b);
    printf("tagon patched index %d\n", (int)(patched - a));
    dump("joined", a);

    struct JPG_TagItem sets[] = {JPG_ValueTag(JPGTAG_IMAGE_WIDTH, 1), JPG_ValueTag(JPGTAG_IMAGE_HEIGHT, 2), JPG_ValueTag(JPGTAG_IMAGE_DEPTH, 3), JPG_EndTag};
    sets[1].SetTagSet();
    sets->ClearTagSets();
    printf("sets: %08x %08x %08x\n", (unsigned)sets[0].ti_Tag, (unsigned)sets[1].ti_Tag, (unsigned)sets[2].ti_Tag);
    printf("sizeof item %d\n", (int)sizeof(struct JPG_TagItem));
    return 0;
}

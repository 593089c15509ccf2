// tagitem.hpp -- tag lists, the variable-argument mechanism of the JPEG interface
// (reference: interface/tagitem.hpp:77-199, semantics interface/tagitem.cpp:63-324).
//
// A tag list is an array of (id, value) items. Ids with bit 31 set are user tags; the four ids below steer the
// traversal: DONE ends a list, IGNORE skips one item, MORE continues at ti_pPtr, SKIP jumps over 1 + ti_lData items.
#ifndef B200JPG_INTERFACE_TAGITEM_HPP
#define B200JPG_INTERFACE_TAGITEM_HPP

#include "jpgtypes.hpp"

typedef JPG_ULONG JPG_Tag;

#define JPGTAG_TAG_DONE (0L)
#define JPGTAG_TAG_END (0L)
#define JPGTAG_TAG_IGNORE (1L)
#define JPGTAG_TAG_MORE (2L)
#define JPGTAG_TAG_SKIP (3L)
#define JPGTAG_TAG_USER (((JPG_ULONG)1) << 31)
#define JPGTAG_SET (((JPG_ULONG)1) << 30)

#ifdef __cplusplus
#define JPG_PointerTag(id, ptr) JPG_TagItem(id, (JPG_APTR)(ptr))
#define JPG_ValueTag(id, v) JPG_TagItem(id, (JPG_LONG)(v))
#define JPG_FloatTag(id, f) JPG_TagItem(id, (JPG_FLOAT)(f))
#define JPG_Continue(tag) JPG_TagItem(JPGTAG_TAG_MORE, const_cast<struct JPG_TagItem *>(tag))
#define JPG_EndTag JPG_TagItem(JPGTAG_TAG_DONE)
#endif

struct JPG_EXPORT JPG_TagItem {
    JPG_Tag ti_Tag;
    union JPG_EXPORT TagContents {
        JPG_LONG ti_lData;
        JPG_FLOAT ti_fData;
        JPG_APTR ti_pPtr;
#ifdef __cplusplus
        TagContents(JPG_LONG v) : ti_pPtr(0) { ti_lData = v; }
        TagContents(JPG_FLOAT v) : ti_pPtr(0) { ti_fData = v; }
        TagContents(JPG_APTR v) : ti_pPtr(v) {}
        TagContents(void) {}
#endif
    } ti_Data;

#ifdef __cplusplus
    JPG_TagItem(JPG_Tag tag, JPG_LONG data) : ti_Tag(tag), ti_Data(data) {}
    JPG_TagItem(JPG_Tag tag, JPG_FLOAT data) : ti_Tag(tag), ti_Data(data) {}
    JPG_TagItem(JPG_Tag tag, JPG_APTR ptr = 0) : ti_Tag(tag), ti_Data(ptr) {}
    JPG_TagItem(void) {}

    // Next user tag after this one, following MORE / SKIP / IGNORE; NULL at the end of the list.
    struct JPG_TagItem *NextTagItem(void);
    const struct JPG_TagItem *NextTagItem(void) const { return const_cast<struct JPG_TagItem *>(this)->NextTagItem(); }
    // First item with the given id, starting the search at this item.
    struct JPG_TagItem *FindTagItem(JPG_Tag id);
    const struct JPG_TagItem *FindTagItem(JPG_Tag id) const { return const_cast<struct JPG_TagItem *>(this)->FindTagItem(id); }
    // Turns the terminating DONE into a MORE that continues at `add`; returns the patched item.
    struct JPG_TagItem *TagOn(struct JPG_TagItem *add);
    const struct JPG_TagItem *Continue(const struct JPG_TagItem *add) {
        ti_Data.ti_pPtr = const_cast<struct JPG_TagItem *>(add);
        return this;
    }
    // Value of the first item with this id, or the default.
    JPG_LONG GetTagData(JPG_Tag id, JPG_LONG defdata = 0) const;
    JPG_FLOAT GetTagFloat(JPG_Tag id, JPG_FLOAT defdata = 0.0) const;
    JPG_APTR GetTagPtr(JPG_Tag id, JPG_APTR defptr = 0) const;
    // Overwrite the first item with this id; silently nothing when absent (the hook protocol relies on it).
    void SetTagData(JPG_Tag id, JPG_LONG data);
    void SetTagFloat(JPG_Tag id, JPG_FLOAT data);
    void SetTagPtr(JPG_Tag id, JPG_APTR ptr);
    void SetTagSet(void);
    void ClearTagSets(void);
    // Copies the user tags of `source`, then those of `defaults` that are neither in `source` nor in `drop`,
    // into `target` (may be NULL to only count). Returns the item count including the terminator.
    static JPG_LONG FilterTags(struct JPG_TagItem *target, const struct JPG_TagItem *source, const struct JPG_TagItem *defaults,
                               const struct JPG_TagItem *drop);
#endif
};

#endif

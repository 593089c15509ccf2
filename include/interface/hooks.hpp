// hooks.hpp -- call-back hooks of the JPEG interface (reference: interface/hooks.hpp:116-173).
// A hook is an entry point plus one pointer of client data; the library calls hk_Entry(hook, taglist).
// The I/O hook (JPGTAG_HOOK_IOHOOK) and the bitmap hook (JPGTAG_BIH_HOOK) both return a JPG_LONG; a negative
// return value makes the library call fail with that code.
#ifndef B200JPG_INTERFACE_HOOKS_HPP
#define B200JPG_INTERFACE_HOOKS_HPP

#include "jpgtypes.hpp"
#include "tagitem.hpp"

struct JPG_EXPORT JPG_Hook {
#ifdef __cplusplus
    typedef JPG_LONG(LongHookFunction)(struct JPG_Hook *, struct JPG_TagItem *tag);
    typedef JPG_APTR(APtrHookFunction)(struct JPG_Hook *, struct JPG_TagItem *tag);
#endif
    union JPG_EXPORT HookCallOut {
        JPG_LONG (*hk_pLongEntry)(struct JPG_Hook *, struct JPG_TagItem *tag);
        JPG_APTR (*hk_pAPtrEntry)(struct JPG_Hook *, struct JPG_TagItem *tag);
#ifdef __cplusplus
        HookCallOut(LongHookFunction *hook) : hk_pLongEntry(hook) {}
        HookCallOut(APtrHookFunction *hook) : hk_pAPtrEntry(hook) {}
        HookCallOut(void) : hk_pLongEntry(0) {}
#endif
    } hk_Entry, hk_SubEntry;  // hk_SubEntry is never called by the library; free for the client
    JPG_APTR hk_pData;        // client data

#ifdef __cplusplus
    JPG_Hook(LongHookFunction *hook = 0, JPG_APTR data = 0) : hk_Entry(hook), hk_pData(data) {}
    JPG_Hook(APtrHookFunction *hook, JPG_APTR data = 0) : hk_Entry(hook), hk_pData(data) {}
    JPG_LONG CallLong(struct JPG_TagItem *tag) { return (*hk_Entry.hk_pLongEntry)(this, tag); }
    JPG_APTR CallAPtr(struct JPG_TagItem *tag) { return (*hk_Entry.hk_pAPtrEntry)(this, tag); }
#endif
};

#endif

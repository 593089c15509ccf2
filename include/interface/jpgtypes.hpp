// jpgtypes.hpp -- scalar types of the JPEG interface (reference: interface/jpgtypes.hpp:63-110).
// Restated for an LP64 / C++11 toolchain: fixed-width integers instead of autoconf-probed ones.
#ifndef B200JPG_INTERFACE_JPGTYPES_HPP
#define B200JPG_INTERFACE_JPGTYPES_HPP

#include <stdint.h>

typedef int32_t JPG_LONG;    // 32-bit signed
typedef uint32_t JPG_ULONG;  // 32-bit unsigned
typedef float JPG_FLOAT;
typedef void *JPG_APTR;      // generic pointer carried in tag items
typedef const void *JPG_CPTR;

#define JPG_TRUE (1)
#define JPG_FALSE (0)

#define JPG_MIN_LONG ((JPG_LONG)(-0x7fffffffL - 1))
#define JPG_MAX_LONG ((JPG_LONG)(0x7fffffffL))
#define JPG_MIN_ULONG ((JPG_ULONG)(0x00000000UL))
#define JPG_MAX_ULONG ((JPG_ULONG)(0xffffffffUL))

#ifndef JPG_EXPORT
#define JPG_EXPORT __attribute__((visibility("default")))
#endif

#endif

/*
 * oracle/ref_config/autoconfig.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Hand-written platform description (Linux, x86-64, gcc >= 9) that stands in
 * for the header the reference's autoconf step would generate.  It lets
 * oracle/Makefile compile the UNMODIFIED reference sources where they lie
 * under /root/reference with plain g++ (-DUSE_AUTOCONF -Ioracle/ref_config),
 * without running the reference's own configure/make machinery.
 * Every macro below is a feature test the reference's config.h / std/ wrappers
 * ask for (config.h:33-83, std/*.hpp); values are facts about this toolchain.
 */
#ifndef B200JPG_ORACLE_AUTOCONFIG_H
#define B200JPG_ORACLE_AUTOCONFIG_H

#define HAS_ALIGNED 1
#define HAS_ATTRIBUTES 1
#define HAS_CONST_CAST 1
#define HAS_REINTERPRET_CAST 1
#define HAS_INT8_T 1
#define HAS_INT16_T 1
#define HAS_INT32_T 1
#define HAS_INT64_T 1
#define HAS_UINT8_T 1
#define HAS_UINT16_T 1
#define HAS_UINT32_T 1
#define HAS_UINT64_T 1
#define HAS_INT_TEMPLATE_ARGS 1
#define HAS_LOCAL_TEMPLATES 1
#define HAS_MAY_ALIAS 1
#define HAS_MEMBER_INIT 1
#define HAS_PRIVATE_ACCESS 1
#define HAS_PTRDIFF_T 1
#define HAS_SC_CLK_TCK 1
#define HAS_STDERR_FILENO 1
#define HAS_STDIN_FILENO 1
#define HAS_STDOUT_FILENO 1
#define HAS_STRUCT_TIMEVAL 1
#define HAS_STRUCT_TMS 1
#define HAS_TEMPLATE_CASTS 1
#define HAS_VISIBILITY_DEFAULT 1
#define HAS_VISIBILITY_HIDDEN 1
#define HAS_VISIBILITY_INTERNAL 1
#define HAS__NULL_TYPE 1
#define HAVE_ALWAYS_INLINE 1
#define HAVE_ASSERT_H 1
#define HAVE_BUILTIN_EXPECT 1
#define HAVE_BUILTIN_MEMCPY 1
#define HAVE_BUILTIN_MEMMOVE 1
#define HAVE_BUILTIN_MEMSET 1
#define HAVE_CLOCK 1
#define HAVE_CLOSE 1
#define HAVE_CTYPE_H 1
#define HAVE_ERRNO_H 1
#define HAVE_FCNTL_H 1
#define HAVE_FREE 1
#define HAVE_FSTAT 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_INTTYPES_H 1
#define HAVE_ISSPACE 1
#define HAVE_LONGJMP 1
#define HAVE_LONG_LONG 1
#define HAVE_LSEEK 1
#define HAVE_MALLOC 1
#define HAVE_MATH_H 1
#define HAVE_MEMCHR 1
#define HAVE_MEMMOVE 1
#define HAVE_MEMORY_H 1
#define HAVE_MEMSET 1
#define HAVE_NORETURN 1
#define HAVE_OPEN 1
#define HAVE_READ 1
#define HAVE_RENAME 1
#define HAVE_RESTRICTED_PTRS 1
#define HAVE_SETJMP_H 1
#define HAVE_SIGNAL_H 1
#define HAVE_SNPRINTF 1
#define HAVE_STDARG_H 1
#define HAVE_STDDEF_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDIO_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRCHR 1
#define HAVE_STRERROR 1
#define HAVE_STRINGS_H 1
#define HAVE_STRING_H 1
#define HAVE_STRRCHR 1
#define HAVE_STRTOD 1
#define HAVE_STRTOL 1
#define HAVE_SYSCONF 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TIMES_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_TIME 1
#define HAVE_TIMES 1
#define HAVE_TIME_H 1
#define HAVE_UNISTD_H 1
#define HAVE_UNSIGNED_LONG_LONG 1
#define HAVE_VPRINTF 1
#define HAVE_VSNPRINTF 1
#define HAVE_WRITE 1
#define JPG_LIL_ENDIAN 1
#define NATURAL_ALIGNMENT 8
#define SIZEOF_CHAR 1
#define SIZEOF_SHORT 2
#define SIZEOF_INT 4
#define SIZEOF_LONG 8
#define SIZEOF_LONG_LONG 8
#define SIZEOF_VOID_P 8
#define STDC_HEADERS 1
#define TIME_WITH_SYS_TIME 1
#define PACKAGE_NAME "libjpeg-ref-oracle"
#define PACKAGE_VERSION "oracle"

#endif

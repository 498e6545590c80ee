/* gd_utilities.h -- small string/number helpers of the host layer.
 * Same names and behaviour as the reference's utilities.h:23-36 (re-typed). */
#ifndef gd_utilities_H
#define gd_utilities_H

#include <inttypes.h>
#include <stddef.h>
#include <float.h>

typedef int32_t  s32;
typedef uint32_t u32;
typedef int64_t  s64;
typedef uint64_t u64;

#define u32Max ((u32) -1)

#ifndef true
#define true  1
#define false 0
#endif

#ifdef __GNUC__
#define arg_dont_complain(arg) arg __attribute__ ((unused))
#else
#define arg_dont_complain(arg) arg
#endif

char*  copy_string            (const char* s);
int    strcmp_prefix          (const char* str1, const char* str2);
int    strcmp_suffix          (const char* str1, const char* str2);
int    string_to_int          (const char* s);
int    string_to_u32          (const char* s);
int    string_to_unitized_int (const char* s, int byThousands);
double string_to_double       (const char* s);
int    try_string_to_double   (const char* s, double* v);
char*  skip_whitespace        (char* s);
char*  skip_darkspace         (char* s);
char*  ucommatize             (const u64 v);
void   safe_strncpy           (char* dest, const char* src, size_t n);

#endif

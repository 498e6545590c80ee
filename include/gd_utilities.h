/* gd_utilities.h -- small string/number helpers the host layer offers to operators.
 *
 * Operators written against the reference include its utilities.h and call these helpers by
 * name, so the names, argument order and behaviour follow the reference (utilities.h:23-36,
 * utilities.c); the declarations are grouped by what they do and each cites the definition it
 * mirrors.  Implemented in genodsp_b200/host/gd_utilities.c. */
#ifndef gd_utilities_H
#define gd_utilities_H

#include <stddef.h>
#include <stdint.h>
#include <inttypes.h>
#include <float.h>

/* ---- fixed-width integer shorthands used throughout the operator interface ------------------- */
typedef uint32_t u32;
typedef int32_t  s32;
typedef uint64_t u64;
typedef int64_t  s64;
#define u32Max ((u32) -1)

#ifndef true
#define false 0
#define true  1
#endif

/* marks a parameter an apply/parse function must accept but does not use */
#if defined(__GNUC__)
#  define arg_dont_complain(arg) arg __attribute__ ((unused))
#else
#  define arg_dont_complain(arg) arg
#endif

/* ---- text -> number: every one of these exits with a message on malformed input ---------------- */
int string_to_int (const char* s);                                   /* utilities.c:135      */
int string_to_u32 (const char* s);                                   /* utilities.c:180      */
/* integer with an optional K, M or G suffix: powers of 1000 (byThousands) or of 1024 */
int string_to_unitized_int (const char* s, int byThousands);         /* utilities.c:236      */
double string_to_double (const char* s);                             /* utilities.c:334      */
/* the same conversion without the exit: returns false and leaves *v alone */
int try_string_to_double (const char* s, double* v);                 /* utilities.c:375      */

/* ---- number -> text ---------------------------------------------------------------------------- */
/* "1,234,567"; the text lives in one of five static buffers used in turn */
char* ucommatize (const u64 v);                                      /* utilities.c:501      */

/* ---- strings ----------------------------------------------------------------------------------- */
/* heap copy; the caller frees */
char* copy_string (const char* s);                                   /* utilities.c:31       */
/* strcmp of str2 against the head (tail) of str1 cut to str2's length: 0 when str2 is a prefix (suffix) of str1 */
int strcmp_prefix (const char* str1, const char* str2);              /* utilities.c:66       */
int strcmp_suffix (const char* str1, const char* str2);              /* utilities.c:93       */
/* first character that is not (that is) white space, or the terminating zero */
char* skip_whitespace (char* s);                                     /* utilities.c:427      */
char* skip_darkspace (char* s);                                      /* utilities.c:430      */
/* strncpy that always terminates */
void safe_strncpy (char* dest, const char* src, size_t n);           /* utilities.c:544      */

#endif

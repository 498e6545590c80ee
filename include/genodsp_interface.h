/* genodsp_interface.h -- operator registration interface of the B200 build.
 *
 * Source-compatible restatement (re-typed, not copied) of the reference's
 * plugin boundary, rsharris/genodsp genodsp_interface.h:
 *   valtype / string helpers            :20-26
 *   spec, chromsOfInterest, chromsSorted :37-57
 *   the five-function operator group    :76-93
 *   dspop, dspinfo, table macros        :101-125
 *   global options, enums                :139-153
 *   services the core exports           :167-190
 * An operator written against the reference header compiles against this one
 * unchanged.  The one semantic difference: `valVector` (and the `v` handed to an
 * apply function) is a DEVICE pointer into the genome buffer; operators reach
 * the kernels through include/gdsp_b200.h and the gd_device services below.
 */
#ifndef genodsp_interface_H
#define genodsp_interface_H

#include <stdio.h>
#include "gd_utilities.h"

#ifdef globals_owner
#define global
#else
#define global extern
#endif

/* ---- per-base values ---------------------------------------------------- */

typedef double valtype;
#define string_to_valtype(s)       ((valtype) string_to_double(s))
#define try_string_to_valtype(s,v) try_string_to_double(s,(valtype*)v)
#define valtypeFmt     "%f"
#define valtypeFmtPrec "%.*f"
#define valtypeMax     DBL_MAX
#define valtypePuny    DBL_MIN

/* ---- chromosome table ----------------------------------------------------
 * chromsOfInterest: linked list in input order (output uses this order)
 * chromsSorted:     NULL-terminated array, longest first (operators use this) */

typedef struct spec
	{
	struct spec* next;
	char*        chrom;
	int          flag;
	u32          start;       /* uninteresting bases before the vector starts   */
	u32          length;      /* number of entries in valVector (never zero)    */
	valtype*     valVector;   /* DEVICE pointer: this chromosome's cells        */
	} spec;

#ifdef globals_owner
global spec*  chromsOfInterest = NULL;
global spec** chromsSorted     = NULL;
#else
global spec*  chromsOfInterest;
global spec** chromsSorted;
#endif

/* ---- operators: short / usage / parse / free / apply ---------------------- */

#define opfuncargs_short (char*,int,FILE*,char*)
#define opfuncargs_usage (char*,FILE*,char*)
#define opfuncargs_parse (char*,int,char**)
#define opfuncargs_free  (struct dspop*)
#define opfuncargs_apply (struct dspop*,char*,u32,valtype*)

typedef void          (*opfunc_short) opfuncargs_short;
typedef void          (*opfunc_usage) opfuncargs_usage;
typedef struct dspop* (*opfunc_parse) opfuncargs_parse;
typedef void          (*opfunc_free)  opfuncargs_free;
typedef void          (*opfunc_apply) opfuncargs_apply;

#define dspprototypes(funcName) \
void          funcName##_short opfuncargs_short; \
void          funcName##_usage opfuncargs_usage; \
struct dspop* funcName##_parse opfuncargs_parse; \
void          funcName##_free  opfuncargs_free;  \
void          funcName##_apply opfuncargs_apply;

/* control record header; every operator's private record starts with one.
 * parse() must set atRandom: 0 = applied to one chromosome vector at a time,
 * 1 = called once for the whole genome with v == NULL. */
typedef struct dspop
	{
	struct dspop* next;
	char*         name;
	opfunc_apply  funcApply;
	opfunc_free   funcFree;
	int           atRandom;
	} dspop;

typedef struct dspinfo
	{
	char*        name;
	opfunc_short funcShort;
	opfunc_usage funcUsage;
	opfunc_parse funcParse;
	opfunc_free  funcFree;
	opfunc_apply funcApply;
	} dspinfo;

#define dspinforecord(name,funcName) \
	{ name, funcName##_short, funcName##_usage, funcName##_parse, funcName##_free, funcName##_apply }
#define dspinfoalias(name) \
	{ name, NULL, NULL, NULL, NULL, NULL }

/* ---- miscellany ------------------------------------------------------------- */

#ifndef M_PI
#define M_PI 3.14159265358979323846264
#endif

#ifdef globals_owner
global int trackOperations     = 0;
global int reportComments      = 0;
global u32 reportInputProgress = 0;
#else
global int trackOperations;
global int reportComments;
global u32 reportInputProgress;
#endif

#define uncovered_NA   -1
#define uncovered_show 1
#define uncovered_hide 0

#define ri_overlapSum 0
#define ri_overlapMin 1
#define ri_overlapMax 2

/* ---- services exported by the core ------------------------------------------ */

void     chastise               (const char* format, ...);
spec*    find_chromosome_spec   (char* chrom);
void     read_intervals         (FILE* f, int valCol, int originOne,
                                 int overlapOp, int clear, valtype missingVal);
int      read_interval          (FILE* f, char* buffer, int bufferLen, int valCol,
                                 char** chrom, u32* start, u32* end, valtype* val);
void     report_intervals       (FILE* f, int precision, int noOutputValues,
                                 int collapseRuns, int showUncovered, int originOne);
void     read_all_chromosomes   (char* filename);
void     write_all_chromosomes  (char* filename);
valtype* get_scratch_vector     (void);     /* DEVICE buffers of maxLength entries */
s32*     get_scratch_ints       (void);
void     release_scratch_vector (valtype* v);
void     release_scratch_ints   (s32* v);
void     set_named_global       (char* name, valtype val);
valtype  get_named_global       (char* name, valtype defaultVal);
int      named_global_exists    (char* name, valtype* val);
void     report_named_globals   (FILE* f, char* indent);
void     tracking_report        (const char* format, ...);
int      valtype_ascending      (const void* v1, const void* v2);

#endif /* genodsp_interface_H */

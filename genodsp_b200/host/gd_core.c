/* gd_core.c -- command line, chromosome table, named variables, interval text
 * I/O and the pipeline executor of the B200 build of genodsp.
 *
 * Host-side counterpart of the reference's genodsp.c: same command grammar
 * (parse_options :284-629, process_operator_options :634-723), same chromosome
 * file format (:728-814), same interval line grammar (read_interval :1384-1534),
 * same output text (report_intervals :1561-1691), same named-variable service
 * (:2056-2202) and progress protocol (:2219-2238).  What differs is WHERE the
 * per-base work happens: the accumulate loops, every operator body and the
 * run-length scan run on the GPU through include/gdsp_b200.h, over one packed
 * genome buffer, and consecutive pointwise operators are fused into one launch.
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <float.h>
#include <unistd.h>
#include <sys/stat.h>

#define globals_owner
#include "genodsp_interface.h"
#include "gd_device.h"
#include "gd_ops.h"

char* programName = "genodsp";
#define programVersion "0.0.10-b200.1"

/* ---- command line state ---------------------------------------------------- */

static dspop* pipeline = NULL;
static dspop* tailOp   = NULL;

static int valColumn      = 4-1;
static int noOutputValues = false;
static int valPrecision   = 0;
static int collapseRuns   = true;
static int showUncovered  = uncovered_hide;
static int clipToLength   = false;
static int originOne      = false;
static int inhibitOutput  = false;
static int dbgInput = false, dbgPipe = false, dbgGlobals = false;

#define specialPipeChar '='

/* ---- usage / chastise -------------------------------------------------------- */

static opfunc_usage chastiseUsage     = NULL;
static char*        chastiseUsageName = NULL;

static void usage (char* message)
	{
	static const char* text[] = {
	"  --chromosomes=<filename>  (required) read chromosome names and lengths from",
	"                            a file",
	"  --value=<col>             input intervals contain a value in the specified",
	"                            column;  by default we assume this is in column 4",
	"  --novalue                 input intervals have no value (value given is 1)",
	"  --nooutputvalue           don't write value with output intervals",
	"  --precision=<number>      number of digits to round output values to",
	"                            (by default, output is rounded to integers)",
	"  --nocollapse              in output, don't collapse runs of identical values",
	"                            to intervals",
	"  --uncovered:hide          don't output intervals that have no coverage",
	"                            (this is the default)",
	"  --uncovered:show          in output, include intervals that have no coverage",
	"  --uncovered:NA            in output, mark uncovered intervals as NA",
	"  --cliptochromosome        clip interals to chromosome length",
	"                            (default is to report such intervals as errors)",
	"  --origin=one              input/output intervals are origin-one, closed",
	"  --origin=zero             input/output intervals are origin-zero, half-open",
	"                            (this is the default)",
	"  --nooutput                don't output the resulting intervals/values",
	"                            (by default these are written to stdout)",
	"  --window=<length>         (W=) size of window",
	"                            (for operators that have a window size)",
	"  --help[=<operator>]       get detail about a particular operator",
	"  ?                         list available operators with brief descriptions",
	"  ?<operator>               same as --help=<operator>",
	"  --report=comments         copy comments from the input to stderr. Comments",
	"                            are lines beginning with a \"#\". This can be",
	"                            helpful in tracking progress during a long run.",
	"  --progress=input:<n>      report processing of every nth input line",
	"  --progress=operations     report each operation as it begins",
	"  --version                 report the program version and quit",
	"",
	"Note that if input intervals overlap, their values are summed.",
	"",
	"Input is usually piped in on stdin. However, if the first operator is \"input\"",
	"stdin is ignored.",
	"", NULL };
	if (message != NULL) fprintf (stderr, "%s\n", message);
	fprintf (stderr, "usage: [cat <file>] | %s --chromosomes=<filename> [options] [operations]\n\n", programName);
	for (int i = 0; text[i] != NULL; i++) fprintf (stderr, "%s\n", text[i]);
	fprintf (stderr, "For a list of available operations, do \"genodsp ?\".\n");
	fprintf (stderr, "For more detailed descriptions of the operations, do \"genodsp --help\".\n");
	exit (EXIT_FAILURE);
	}

void chastise (const char* format, ...)
	{
	va_list args;
	va_start (args, format);
	if (format != NULL) vfprintf (stderr, format, args);
	va_end (args);
	if (chastiseUsage != NULL)
		{
		(*chastiseUsage) (chastiseUsageName, stderr, "  ");
		exit (EXIT_FAILURE);
		}
	usage (NULL);
	}

static void usage_operations (void)
	{
	fprintf (stderr, "Operations (general form is %c <operator> [arguments]):\n", specialPipeChar);
	for (u32 ix = 0; ix < dspTableLen; ix++)
		if (dspTable[ix].funcShort != NULL) (*dspTable[ix].funcShort) (dspTable[ix].name, 12, stderr, "  ");
	exit (EXIT_FAILURE);
	}

/* name -> table row; an alias resolves to the nearest real row above it */
static dspinfo* find_operator (const char* name)
	{
	dspinfo* real = NULL;
	for (u32 ix = 0; ix < dspTableLen; ix++)
		{
		if (dspTable[ix].funcShort != NULL) real = &dspTable[ix];
		if (strcmp (name, dspTable[ix].name) == 0) return real;
		}
	return NULL;
	}

/* ---- named variables ------------------------------------------------------------ */

typedef struct namedglobal { struct namedglobal* next;  char* name;  valtype v; } namedglobal;
static namedglobal* namedGlobalHead = NULL;

static namedglobal* find_global (const char* name)
	{
	for (namedglobal* ng = namedGlobalHead; ng != NULL; ng = ng->next)
		if (strcmp (name, ng->name) == 0) return ng;
	return NULL;
	}

void set_named_global (char* name, valtype val)
	{
	if (dbgGlobals) fprintf (stderr, "set_named_global(%s," valtypeFmt ")\n", name, val);
	namedglobal* ng = find_global (name);
	if (ng == NULL)
		{
		ng = (namedglobal*) malloc (sizeof (namedglobal));
		if (ng == NULL) { fprintf (stderr, "failed to allocate named global \"%s\"\n", name);  exit (EXIT_FAILURE); }
		ng->name = copy_string (name);
		ng->next = namedGlobalHead;             /* newest first, as the reference lists them */
		namedGlobalHead = ng;
		}
	ng->v = val;
	}

valtype get_named_global (char* name, valtype defaultVal)
	{
	namedglobal* ng = find_global (name);
	if (dbgGlobals)
		{
		if (ng == NULL) fprintf (stderr, "get_named_global(%s) = " valtypeFmt " (default)\n", name, defaultVal);
		else            fprintf (stderr, "get_named_global(%s) = " valtypeFmt "\n", name, ng->v);
		}
	return (ng == NULL) ? defaultVal : ng->v;
	}

int named_global_exists (char* name, valtype* v)
	{
	namedglobal* ng = find_global (name);
	if (dbgGlobals)
		{
		if (ng == NULL) fprintf (stderr, "named_global_exists(%s) =  (not found)\n", name);
		else            fprintf (stderr, "named_global_exists(%s) = " valtypeFmt "\n", name, ng->v);
		}
	if (ng == NULL) return false;
	if (v != NULL) *v = ng->v;
	return true;
	}

void report_named_globals (FILE* f, char* indent)
	{
	int nameW = 1;
	if (indent == NULL) indent = "";
	for (namedglobal* ng = namedGlobalHead; ng != NULL; ng = ng->next)
		if ((int) strlen (ng->name) > nameW) nameW = (int) strlen (ng->name);
	if (nameW > 20) nameW = 20;
	for (namedglobal* ng = namedGlobalHead; ng != NULL; ng = ng->next)
		fprintf (f, "%s%*s = " valtypeFmt "\n", indent, nameW, ng->name, ng->v);
	}

static void free_named_globals (void)
	{
	namedglobal* next;
	for (namedglobal* ng = namedGlobalHead; ng != NULL; ng = next)
		{ next = ng->next;  free (ng->name);  free (ng); }
	namedGlobalHead = NULL;
	}

/* ---- progress line ---------------------------------------------------------------- */

void tracking_report (const char* format, ...)
	{
	static char line[1001];
	static int  prevLen = 0;
	va_list args;
	line[0] = 0;
	va_start (args, format);
	if (format != NULL) vsnprintf (line, sizeof (line), format, args);
	va_end (args);
	int len = (int) strlen (line);
	int newline = (len > 0) && (line[len-1] == '\n');
	if (newline) line[--len] = 0;
	fprintf (stderr, "%s", line);
	if (prevLen > len) fprintf (stderr, "%*s", prevLen - len, "");
	if (newline) { fprintf (stderr, "\n");  prevLen = 0; }
	        else { fprintf (stderr, "\r");  prevLen = len; }
	}

int valtype_ascending (const void* a, const void* b)
	{
	const valtype x = *(const valtype*) a, y = *(const valtype*) b;
	return (x > y) - (x < y);
	}

/* ---- chromosome table -------------------------------------------------------------- */

static int add_chromosome_spec (char* name, u32 chromStart, u32 chromLength)
	{
	spec* tail = NULL;
	if (chromLength == 0) return true;
	for (spec* s = chromsOfInterest; s != NULL; s = s->next)
		{
		if (strcmp (name, s->chrom) == 0) return false;
		tail = s;
		}
	spec* n = (spec*) malloc (sizeof (spec));
	if (n == NULL) { fprintf (stderr, "failed to allocate spec for \"%s\", %d bytes\n", name, (int) sizeof (spec));  exit (EXIT_FAILURE); }
	n->next = NULL;  n->chrom = copy_string (name);  n->flag = 0;
	n->start = chromStart;  n->length = chromLength;  n->valVector = NULL;
	if (tail == NULL) chromsOfInterest = n; else tail->next = n;
	return true;
	}

spec* find_chromosome_spec (char* chrom)
	{
	for (spec* s = chromsOfInterest; s != NULL; s = s->next)
		if (strcmp (chrom, s->chrom) == 0) return s;
	return NULL;
	}

static void read_chromosome_lengths (char* filename)
	{
	char line[1001];
	u32  lineNumber = 0;
	int  missingEol = false;
	FILE* f = fopen (filename, "rt");
	if (f == NULL) { fprintf (stderr, "can't open \"%s\" for reading\n", filename);  exit (EXIT_FAILURE); }
	while (fgets (line, sizeof (line), f) != NULL)
		{
		lineNumber++;
		if (missingEol)
			{ fprintf (stderr, "problem at line %u, line is longer than internal buffer\n", lineNumber - 1);  exit (EXIT_FAILURE); }
		int len = (int) strlen (line);
		if (len != 0) missingEol = (line[len-1] != '\n');
		char* scan = skip_whitespace (line);
		if (*scan == 0 || *scan == '#') continue;
		char* chrom = line;
		char* mark = skip_darkspace (line);
		scan = skip_whitespace (mark);
		if (*mark != 0) *mark = 0;
		if (*scan == 0)
			{ fprintf (stderr, "problem at line %u, line contains no chromosome length\n", lineNumber);  exit (EXIT_FAILURE); }
		char* field = scan;
		mark = skip_darkspace (scan);
		if (*mark != 0) *mark = 0;
		u32 chromLength = string_to_u32 (field);
		if (!add_chromosome_spec (chrom, 0, chromLength))
			{
			fprintf (stderr, "problem at line %u, chromosome \"%s\" appears more than once\n", lineNumber, chrom);
			exit (EXIT_FAILURE);
			}
		}
	fclose (f);
	}

static int spec_length_descending (const void* a, const void* b)
	{
	const u32 x = (*(spec* const*) a)->length, y = (*(spec* const*) b)->length;
	return (x < y) - (x > y);
	}

static void sort_chromosomes_by_length (void)
	{
	int n = 0;
	for (spec* s = chromsOfInterest; s != NULL; s = s->next) n++;
	chromsSorted = (spec**) malloc ((n + 1) * sizeof (spec*));
	if (chromsSorted == NULL) { fprintf (stderr, "failed to allocate sorted chromosome list\n");  exit (EXIT_FAILURE); }
	int i = 0;
	for (spec* s = chromsOfInterest; s != NULL; s = s->next) chromsSorted[i++] = s;
	chromsSorted[i] = NULL;
	qsort (chromsSorted, n, sizeof (spec*), spec_length_descending);   /* same libc qsort, same comparator */
	}

/* ---- scratch vectors (device) -------------------------------------------------------- */

typedef struct scratch { struct scratch* next;  int inUse;  void* vec; } scratch;
static scratch* scratchD = NULL;
static scratch* scratchI = NULL;

static void* scratch_get (scratch** head, size_t bytes)
	{
	for (scratch* s = *head; s != NULL; s = s->next)
		if (!s->inUse) { s->inUse = true;  return s->vec; }
	scratch* s = (scratch*) malloc (sizeof (scratch));
	s->next = *head;  *head = s;  s->inUse = true;
	gd_check (gdsp_malloc (gd.ctx, bytes, &s->vec), "scratch");
	return s->vec;
	}

static void scratch_release (scratch* head, void* v)
	{ for (scratch* s = head; s != NULL; s = s->next) if (s->vec == v) { s->inUse = false;  break; } }

valtype* get_scratch_vector (void) { return (valtype*) scratch_get (&scratchD, (size_t) gd.maxLength * sizeof (valtype)); }
s32*     get_scratch_ints   (void) { return (s32*)     scratch_get (&scratchI, (size_t) gd.maxLength * sizeof (s32)); }
void release_scratch_vector (valtype* v) { scratch_release (scratchD, v); }
void release_scratch_ints   (s32* v)     { scratch_release (scratchI, v); }

static void free_scratch_vectors (void)
	{
	scratch* next;
	for (scratch* s = scratchD; s != NULL; s = next) { next = s->next;  gdsp_free (gd.ctx, s->vec);  free (s); }
	for (scratch* s = scratchI; s != NULL; s = next) { next = s->next;  gdsp_free (gd.ctx, s->vec);  free (s); }
	scratchD = scratchI = NULL;
	}

/* ---- interval reader --------------------------------------------------------------------
 * read_interval keeps the reference's line grammar and messages (genodsp.c:1384-1534): lines of at
 * most 1000 characters, "track " lines, blank lines and '#' comments skipped, first column must
 * not start with a blank, start/end unsigned integers, value taken from column valCol.  The
 * numeric fields go through a fast decimal parser and fall back to the sscanf-based helpers for
 * anything unusual, so acceptance and error texts are unchanged. */

static void phase_mark (const char* what);

static u64 riLineNumber = 0;         /* never reset, like the reference's static counter */
static int riMissingEol = false;

/* fgets with the same contract (at most len-1 characters, stops after a newline, NULL at end of
 * file with nothing read) fed from 1 MB blocks: the locked per-call stdio path costs more than the
 * parsing of a 25-character line. */
#define LR_CAP (1u << 20)
static struct { FILE* f;  char* buf;  size_t pos, len;  int eof;  const char* mem;  size_t memLen; } lr;

/* feed the line reader from memory: the next gd_fgets(…, f) calls return the lines of [p, p+n) and then NULL */
static void gd_fgets_from_memory (FILE* f, const char* p, size_t n)
	{
	if (lr.buf == NULL) lr.buf = (char*) malloc (LR_CAP);
	lr.f = f;  lr.mem = p;  lr.memLen = n;  lr.pos = lr.len = 0;  lr.eof = false;
	}

static char* gd_fgets (char* dst, int dstLen, FILE* f)
	{
	if (lr.f != f)
		{
		lr.f = f;  lr.pos = lr.len = 0;  lr.eof = false;  lr.mem = NULL;  lr.memLen = 0;
		if (lr.buf == NULL) lr.buf = (char*) malloc (LR_CAP);
		}
	int n = 0;
	while (n < dstLen - 1)
		{
		if (lr.pos == lr.len)
			{
			if (lr.eof) break;
			if (lr.mem != NULL)
				{
				lr.len = (lr.memLen < LR_CAP) ? lr.memLen : LR_CAP;
				memcpy (lr.buf, lr.mem, lr.len);  lr.mem += lr.len;  lr.memLen -= lr.len;
				}
			else lr.len = fread (lr.buf, 1, LR_CAP, f);
			lr.pos = 0;
			if (lr.len == 0) { lr.eof = true;  break; }
			}
		size_t avail = lr.len - lr.pos, room = (size_t) (dstLen - 1 - n);
		size_t take = (avail < room) ? avail : room;
		char* nl = (char*) memchr (lr.buf + lr.pos, '\n', take);
		if (nl != NULL) take = (size_t) (nl - (lr.buf + lr.pos)) + 1;
		memcpy (dst + n, lr.buf + lr.pos, take);
		n += (int) take;  lr.pos += take;
		if (nl != NULL) break;
		}
	if (n == 0) { lr.f = NULL;  return NULL; }          /* the FILE may be closed and its address reused */
	dst[n] = 0;
	return dst;
	}

/* same classification as isspace() in the C locale, without the call */
static inline int is_ws (char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
static inline char* ws_skip   (char* s) { while (*s != 0 &&  is_ws (*s)) s++;  return s; }
static inline char* dark_skip (char* s) { while (*s != 0 && !is_ws (*s)) s++;  return s; }

static inline int fast_u32 (const char* s, u32* out)
	{
	u64 v = 0;
	const char* p = s;
	if (*p < '0' || *p > '9') return false;
	while (*p >= '0' && *p <= '9') { v = v * 10 + (u64) (*p - '0');  if (v > 0xffffffffull) return false;  p++; }
	if (*p != 0) return false;
	*out = (u32) v;
	return true;
	}

static inline valtype parse_value (const char* s)
	{
	char c = s[0];
	if ((c >= '0' && c <= '9') || ((c == '-' || c == '+' || c == '.') && s[1] >= '0' && s[1] <= '9'))
		{
		char* endp;
		double v = strtod (s, &endp);
		if (*endp == 0) return v;
		}
	return string_to_valtype (s);
	}

int read_interval (FILE* f, char* buffer, int bufferLen, int valCol,
                   char** _chrom, u32* _start, u32* _end, valtype* _val)
	{
	char *scan, *mark, *field;
	u32  start, end;
	valtype val;

	while (true)
		{
		if (gd_fgets (buffer, bufferLen, f) == NULL) return false;
		riLineNumber++;
		if (riMissingEol)
			{
			fprintf (stderr, "problem at line %s, line is longer than internal buffer\n", ucommatize (riLineNumber - 1));
			exit (EXIT_FAILURE);
			}
		int len = (int) strlen (buffer);
		if (len != 0) riMissingEol = (buffer[len-1] != '\n');
		if (dbgInput) fprintf (stderr, "input = \"%s\"\n", buffer);
		if (buffer[0] == 't' && strcmp_prefix (buffer, "track ") == 0) continue;

		int progressNow = (reportInputProgress != 0)
		               && (riLineNumber == 1 || riLineNumber % reportInputProgress == 0);
		scan = ws_skip (buffer);
		if (*scan == 0)
			{
			if (progressNow) fprintf (stderr, "progress: input line %s\n", ucommatize (riLineNumber));
			continue;
			}
		if (*scan == '#')
			{
			if (reportComments)   fprintf (stderr, "input line %s: %s", ucommatize (riLineNumber), scan);
			else if (progressNow) fprintf (stderr, "progress: input line %s\n", ucommatize (riLineNumber));
			continue;
			}
		if (progressNow) fprintf (stderr, "progress: input line %s\n", ucommatize (riLineNumber));
		break;
		}

	char* chrom = scan = buffer;
	if (*scan == ' ')
		{
		fprintf (stderr, "problem at line %s, line contains no chromosome or begins with whitespace\n", ucommatize (riLineNumber));
		exit (EXIT_FAILURE);
		}
	mark = dark_skip (scan);  scan = ws_skip (mark);  if (*mark != 0) *mark = 0;
	if (*scan == 0)
		{
		fprintf (stderr, "problem at line %s, line contains no interval start\n"
		                 "(expected \"chromosome start end ...\", but there are fewer than 2 fields)\n", ucommatize (riLineNumber));
		exit (EXIT_FAILURE);
		}
	field = scan;  mark = dark_skip (scan);  scan = ws_skip (mark);  if (*mark != 0) *mark = 0;
	if (!fast_u32 (field, &start)) start = (u32) string_to_u32 (field);
	if (*scan == 0)
		{
		fprintf (stderr, "problem at line %s, line contains no interval end\n"
		                 "(expected \"chromosome start end ...\", but there are fewer than 3 fields)\n", ucommatize (riLineNumber));
		exit (EXIT_FAILURE);
		}
	field = scan;  mark = dark_skip (scan);  scan = ws_skip (mark);  if (*mark != 0) *mark = 0;
	if (!fast_u32 (field, &end)) end = (u32) string_to_u32 (field);

	if (valCol == -1 || _val == NULL) val = 1.0;
	else
		{
		for (int col = 3; col <= valCol; col++)
			{
			if (*scan == 0)
				{
				fprintf (stderr, "problem at line %s, line contains no interval value\n"
				                 "(expected \"chromosome start end value\", but there are fewer than 4 fields)\n", ucommatize (riLineNumber));
				exit (EXIT_FAILURE);
				}
			field = scan;  mark = dark_skip (scan);  scan = ws_skip (mark);
			}
		if (*mark != 0) *mark = 0;
		val = parse_value (field);
		}
	if (_chrom != NULL) *_chrom = chrom;
	if (_start != NULL) *_start = start;
	if (_end   != NULL) *_end   = end;
	if (_val   != NULL) *_val   = val;
	return true;
	}

/* ---- parallel tokenizer ---------------------------------------------------------------------
 * The interval text is read in 32 MB blocks cut at a newline; a block is split into sub-chunks at
 * newlines and every sub-chunk is tokenized by its own thread into a private SoA batch, while the main
 * thread already reads the next block.  A worker only accepts what the reference's read_interval accepts
 * without a second look -- "chrom ws digits ws digits [ws ...]" lines of at most 1000 characters, "track "
 * lines, blank lines, '#' comments -- and flags its sub-chunk as soon as it meets anything else (a field
 * the fast parser does not take, a line that starts with a blank, an interval beyond its chromosome ...).
 * A flagged sub-chunk is re-read by the sequential read_interval from memory, with the line counter where
 * it belongs: acceptance, messages, line numbers and exit status are the reference's by construction
 * (genodsp.c:1384-1534).  --progress=input, --report=comments, --debug=input and --progress=operations keep
 * the one-line-at-a-time path. */
#include <pthread.h>

#define PR_BLOCK   (32u << 20)
#define PR_MAXTHR  32

typedef struct prchunk
	{
	const char* p;  size_t n;          /* whole lines (the stream's last line may lack its newline)   */
	int     valCol, origin;
	u32*    seg;  u32* start;  u32* end;  double* val;
	u64     cnt, cap, lines;
	int     flagged, allInt;
	double  sumAbs;
	} prchunk;

static int ingest_interval (ivlist* l, spec* cs, int segIx, char* chrom, u32 start, u32 end, valtype val, u32 o,
                            int* allInt, double* sumAbs, int quiet);

static void* pr_worker (void* arg)
	{
	prchunk* c = (prchunk*) arg;
	const char* p = c->p;  const char* stop = c->p + c->n;
	char   prevName[1001];  size_t prevLen = (size_t) -1;
	spec*  cs = NULL;  int segIx = -1;
	const u32 o = (u32) c->origin;
	c->cnt = 0;  c->lines = 0;  c->flagged = false;  c->allInt = true;  c->sumAbs = 0.0;
	while (p < stop)
		{
		const char* nl = (const char*) memchr (p, '\n', (size_t) (stop - p));
		const char* e  = (nl != NULL) ? nl : stop;                 /* line text is [p, e)                  */
		const char* next = (nl != NULL) ? nl + 1 : stop;
		if ((size_t) (next - p) > 1000) { c->flagged = true;  return NULL; }     /* fgets would split it         */
		c->lines++;
		const char* q = p;
		if (e - p >= 6 && memcmp (p, "track ", 6) == 0) { p = next;  continue; }
		while (q < e && is_ws (*q)) q++;
		if (q == e || *q == '#') { p = next;  continue; }
		if (q != p) { c->flagged = true;  return NULL; }           /* begins with whitespace               */
		const char* nameEnd = q;
		while (nameEnd < e && !is_ws (*nameEnd)) nameEnd++;
		const size_t nameLen = (size_t) (nameEnd - q);
		if (memchr (q, 0, (size_t) (e - q)) != NULL) { c->flagged = true;  return NULL; }   /* NUL inside the line */
		u64 f[2];
		const char* r = nameEnd;
		for (int k = 0; k < 2; k++)
			{
			while (r < e && is_ws (*r)) r++;
			if (r == e || *r < '0' || *r > '9') { c->flagged = true;  return NULL; }
			u64 v = 0;
			while (r < e && *r >= '0' && *r <= '9') { v = v * 10 + (u64) (*r - '0');  if (v > 0xffffffffull) { c->flagged = true;  return NULL; }  r++; }
			if (r < e && !is_ws (*r)) { c->flagged = true;  return NULL; }
			f[k] = v;
			}
		valtype val = 1.0;
		if (c->valCol != -1)
			{
			const char* fs = NULL;  const char* fe = NULL;
			for (int col = 3; col <= c->valCol; col++)
				{
				while (r < e && is_ws (*r)) r++;
				if (r == e) { c->flagged = true;  return NULL; }
				fs = r;
				while (r < e && !is_ws (*r)) r++;
				fe = r;
				}
			char tmp[64];
			const size_t fl = (size_t) (fe - fs);
			const char c0 = fs[0];
			if (fl >= sizeof (tmp)
			 || !((c0 >= '0' && c0 <= '9') || ((c0 == '-' || c0 == '+' || c0 == '.') && fl > 1 && fs[1] >= '0' && fs[1] <= '9')))
				{ c->flagged = true;  return NULL; }
			memcpy (tmp, fs, fl);  tmp[fl] = 0;
			char* endp;
			val = strtod (tmp, &endp);
			if (*endp != 0) { c->flagged = true;  return NULL; }
			}
		if (nameLen != prevLen || memcmp (q, prevName, nameLen) != 0)
			{
			memcpy (prevName, q, nameLen);  prevName[nameLen] = 0;  prevLen = nameLen;
			cs = find_chromosome_spec (prevName);
			segIx = (cs != NULL) ? gd_sorted_index (cs) : -1;
			}
		p = next;
		if (cs == NULL) continue;
		ivlist one;                                                 /* a view on this chunk's arrays          */
		one.n = c->cnt;  one.cap = c->cap;  one.seg = c->seg;  one.start = c->start;  one.end = c->end;  one.val = c->val;
		if (!ingest_interval (&one, cs, segIx, prevName, (u32) f[0], (u32) f[1], val, o, &c->allInt, &c->sumAbs, true))
			{ c->flagged = true;  return NULL; }
		c->cnt = one.n;
		}
	return NULL;
	}

/* what read_intervals does with one parsed line (genodsp.c:1245-1330 up to the per-base loop): origin shift,
 * clipping, the "beyond the end" check, then the interval joins the batch.  quiet: return false instead of
 * printing the message and exiting (the parallel tokenizer then replays the line sequentially). */
static int ingest_interval (ivlist* l, spec* cs, int segIx, char* chrom, u32 start, u32 end, valtype val, u32 o,
                            int* allInt, double* sumAbs, int quiet)
	{
	start -= o;
	u32 a = start, b = end;
	if (clipToLength)
		{
		if (start > cs->start + cs->length) a = start = cs->start + cs->length;
		if (end   > cs->start + cs->length) b = end   = cs->start + cs->length;
		}
	if (cs->start == 0)
		{
		if (end > cs->length)
			{
			if (quiet) return false;
			fprintf (stderr, "%s %d %d is beyond the end of the chromosome (L=%d)\n", chrom, start, end, cs->length);
			exit (EXIT_FAILURE);
			}
		}
	else
		{
		if (end <= cs->start) return true;
		b = end - cs->start;
		a = (start <= cs->start) ? 0 : start - cs->start;
		if (a >= cs->length) return true;
		if (b >= cs->length) b = cs->length;
		}
	if (a >= b) return true;                  /* the reference's loop body would not execute */
	if (val != floor (val) || fabs (val) > 1e6) *allInt = false;
	*sumAbs += fabs (val);
	if (quiet)
		{
		if (l->n == l->cap) return false;     /* cannot happen: cap = bytes / 6 + 16 */
		l->seg[l->n] = (u32) segIx;  l->start[l->n] = a;  l->end[l->n] = b;  l->val[l->n] = val;  l->n++;
		}
	else ivlist_push (l, (u32) segIx, a, b, val);
	return true;
	}

/* the one-line-at-a-time loop (also the replay of a flagged sub-chunk, fed from memory) */
static void read_intervals_sequential (FILE* f, int valCol, u32 o, ivlist* l, int* allInt, double* sumAbs)
	{
	char   line[1001], prevChrom[1001];
	char*  chrom;
	spec*  cs = NULL;
	int    segIx = -1;
	u32    start, end;
	valtype val;
	prevChrom[0] = 0;
	while (read_interval (f, line, sizeof (line), valCol, &chrom, &start, &end, &val))
		{
		if (strcmp (chrom, prevChrom) != 0)
			{
			cs = find_chromosome_spec (chrom);
			segIx = (cs != NULL) ? gd_sorted_index (cs) : -1;
			safe_strncpy (prevChrom, chrom, sizeof (prevChrom) - 1);
			}
		if (cs == NULL) continue;
		if (trackOperations && !cs->flag) { tracking_report ("input(%s)\n", chrom);  cs->flag = true; }
		ingest_interval (l, cs, segIx, chrom, start, end, val, o, allInt, sumAbs, false);
		}
	}

static void ivlist_append (ivlist* l, const prchunk* c)
	{
	if (l->n + c->cnt > l->cap)
		{
		u64 nc = l->cap ? l->cap : (1u << 16);
		while (nc < l->n + c->cnt) nc *= 2;
		l->seg   = (u32*) realloc (l->seg,   nc * sizeof (u32));
		l->start = (u32*) realloc (l->start, nc * sizeof (u32));
		l->end   = (u32*) realloc (l->end,   nc * sizeof (u32));
		l->val   = (double*) realloc (l->val, nc * sizeof (double));
		if (!l->seg || !l->start || !l->end || !l->val)
			{ fprintf (stderr, "out of memory holding %llu intervals\n", (unsigned long long) nc);  exit (EXIT_FAILURE); }
		l->cap = nc;
		}
	memcpy (l->seg + l->n,   c->seg,   c->cnt * sizeof (u32));
	memcpy (l->start + l->n, c->start, c->cnt * sizeof (u32));
	memcpy (l->end + l->n,   c->end,   c->cnt * sizeof (u32));
	memcpy (l->val + l->n,   c->val,   c->cnt * sizeof (double));
	l->n += c->cnt;
	}

static int pr_thread_count (void)
	{
	const char* e = getenv ("GENODSP_THREADS");
	long n = (e != NULL) ? atol (e) : sysconf (_SC_NPROCESSORS_ONLN);
	if (n < 1) n = 1;
	if (n > PR_MAXTHR) n = PR_MAXTHR;
	return (int) n;
	}

static size_t pr_fill (FILE* f, char* buf, size_t have, int* eof)
	{
	while (!*eof && have < PR_BLOCK)
		{
		size_t r = fread (buf + have, 1, PR_BLOCK - have, f);
		if (r == 0) { *eof = true;  break; }
		have += r;
		}
	return have;
	}

static void pr_merge (FILE* f, prchunk* ch, int used, int valCol, u32 o, ivlist* l, int* allInt, double* sumAbs)
	{
	for (int t = 0; t < used; t++)
		{
		if (!ch[t].flagged)
			{
			ivlist_append (l, &ch[t]);
			riLineNumber += ch[t].lines;
			if (!ch[t].allInt) *allInt = false;
			*sumAbs += ch[t].sumAbs;
			}
		else
			{
			gd_fgets_from_memory (f, ch[t].p, ch[t].n);
			read_intervals_sequential (f, valCol, o, l, allInt, sumAbs);
			}
		}
	}

static void read_intervals_parallel (FILE* f, int valCol, u32 o, ivlist* l, int* allInt, double* sumAbs, int nthr)
	{
	/* three text blocks in rotation: one being tokenized, one being merged (a flagged sub-chunk is replayed from
	 * its text), one being filled; two sets of per-thread batches: the workers fill one while the main thread
	 * merges the other into the interval list */
	char* blk[3];
	for (int b = 0; b < 3; b++) blk[b] = (char*) malloc (PR_BLOCK + 1024);
	static prchunk ch[2][PR_MAXTHR];
	pthread_t th[PR_MAXTHR];  int started[PR_MAXTHR];
	const u64 cap = PR_BLOCK / 6 / (u64) nthr + 4096;          /* shares are equal up to one line; a 6-byte line is the shortest */
	for (int s2 = 0; s2 < 2; s2++)
		for (int t = 0; t < nthr; t++)
			{
			prchunk* c = &ch[s2][t];
			memset (c, 0, sizeof (prchunk));
			c->cap = cap;  c->valCol = valCol;  c->origin = (int) o;
			c->seg = (u32*) malloc (cap * 4);  c->start = (u32*) malloc (cap * 4);  c->end = (u32*) malloc (cap * 4);
			c->val = (double*) malloc (cap * 8);
			if (!c->seg || !c->start || !c->end || !c->val) { fprintf (stderr, "out of memory for the input buffers\n");  exit (EXIT_FAILURE); }
			}
	if (!blk[0] || !blk[1] || !blk[2]) { fprintf (stderr, "out of memory for the input buffers\n");  exit (EXIT_FAILURE); }
	/* a regular file tells its size: reserve the list once instead of doubling it */
	{
	struct stat st;
	if (fstat (fileno (f), &st) == 0 && S_ISREG (st.st_mode) && st.st_size > 0)
		{
		prchunk none;  memset (&none, 0, sizeof (none));
		const u64 want = (u64) st.st_size / 16 + 1024;
		if (want > l->cap)
			{
			l->seg = (u32*) realloc (l->seg, want * 4);  l->start = (u32*) realloc (l->start, want * 4);
			l->end = (u32*) realloc (l->end, want * 4);  l->val = (double*) realloc (l->val, want * 8);
			if (!l->seg || !l->start || !l->end || !l->val) { fprintf (stderr, "out of memory holding %llu intervals\n", (unsigned long long) want);  exit (EXIT_FAILURE); }
			l->cap = want;
			}
		}
	}
	int cur = 0, set = 0, eof = false, pendingUsed = 0;
	size_t got = 0;
	for (;;)
		{
		got = pr_fill (f, blk[cur], got, &eof);
		if (got == 0) break;
		size_t whole = got;                                   /* bytes that form whole lines */
		if (!eof) while (whole > 0 && blk[cur][whole-1] != '\n') whole--;
		if (whole == 0)
			{
			/* not one newline in 32 MB: the sequential reader reports the over-long line */
			pr_merge (f, ch[set ^ 1], pendingUsed, valCol, o, l, allInt, sumAbs);  pendingUsed = 0;
			gd_fgets_from_memory (f, blk[cur], got);
			read_intervals_sequential (f, valCol, o, l, allInt, sumAbs);
			got = 0;
			continue;
			}
		int used = 0;
		size_t at = 0;
		for (int t = 0; t < nthr && at < whole; t++)          /* sub-chunks cut at newlines */
			{
			size_t end = (t == nthr - 1) ? whole : at + (whole - at) / (size_t) (nthr - t);
			while (end < whole && blk[cur][end-1] != '\n') end++;
			ch[set][t].p = blk[cur] + at;  ch[set][t].n = end - at;
			at = end;  used = t + 1;
			}
		for (int t = 0; t < used; t++)
			{
			started[t] = (pthread_create (&th[t], NULL, pr_worker, &ch[set][t]) == 0);
			if (!started[t]) pr_worker (&ch[set][t]);
			}
		/* while they tokenize: merge the previous block's batches, move the partial last line to the next text
		 * block and read on behind it */
		pr_merge (f, ch[set ^ 1], pendingUsed, valCol, o, l, allInt, sumAbs);
		const int nxt = (cur + 1) % 3;
		size_t nextGot = got - whole;
		memcpy (blk[nxt], blk[cur] + whole, nextGot);
		nextGot = pr_fill (f, blk[nxt], nextGot, &eof);
		for (int t = 0; t < used; t++) if (started[t]) pthread_join (th[t], NULL);
		pendingUsed = used;  set ^= 1;
		cur = nxt;  got = nextGot;
		}
	pr_merge (f, ch[set ^ 1], pendingUsed, valCol, o, l, allInt, sumAbs);
	for (int s2 = 0; s2 < 2; s2++)
		for (int t = 0; t < nthr; t++) { free (ch[s2][t].seg);  free (ch[s2][t].start);  free (ch[s2][t].end);  free (ch[s2][t].val); }
	for (int b = 0; b < 3; b++) free (blk[b]);
	}

/* read_intervals: text -> SoA batch on the host -> accumulation on the GPU.
 * Replaces the per-base loops of genodsp.c:1307-1330.  Validation (origin shift, clipping,
 * "beyond the end of the chromosome") happens here with the reference's messages. */
void read_intervals (FILE* f, int valCol, int originOne_, int overlapOp, int clear, valtype missingVal)
	{
	u32    o = originOne_ ? 1 : 0;
	ivlist l;

	ivlist_init (&l);
	if (trackOperations) for (int i = 0; i < gd.nchrom; i++) chromsSorted[i]->flag = false;
	int allInt = true;
	double sumAbs = 0.0;

	const int nthr = pr_thread_count ();
	if (nthr > 1 && !trackOperations && !dbgInput && reportInputProgress == 0 && !reportComments && !riMissingEol)
		read_intervals_parallel (f, valCol, o, &l, &allInt, &sumAbs, nthr);
	else
		read_intervals_sequential (f, valCol, o, &l, &allInt, &sumAbs);
	if (getenv ("GENODSP_PARSE_ONLY") != NULL)
		{
		/* tokenizer self-check (tests/test_cpu_suite.py): what was parsed, without touching the device */
		u64 h = 1469598103934665603ull;
		for (u64 i = 0; i < l.n; i++)
			{
			u64 bits;  memcpy (&bits, &l.val[i], 8);
			u64 w[4] = { l.seg[i], l.start[i], l.end[i], bits };
			for (int k = 0; k < 4; k++) { h ^= w[k];  h *= 1099511628211ull; }
			}
		printf ("intervals=%llu hash=%016llx allInt=%d lines=%llu\n", (unsigned long long) l.n, (unsigned long long) h,
		        allInt, (unsigned long long) riLineNumber);
		exit (EXIT_SUCCESS);
		}
	phase_mark ("text parse");
	gd_device_wait ();                        /* the device was opening while the text was parsed */
	phase_mark ("device wait");
	if (trackOperations) tracking_report ("input(--done--)\n");

	if (overlapOp != ri_overlapSum)
		gd_input_minmax (&l, overlapOp, clear, missingVal);      /* gd_ops_io.c */
	else
		{
		/* clear: every vector starts at missingVal, the first interval touching a cell replaces it
		 * (genodsp.c:1327).  For missingVal==0 that is a plain sum; otherwise the uncovered cells are
		 * set to missingVal afterwards through the union of the intervals. */
		int mode = (valCol == -1 || (allInt && sumAbs < 2.0e9)) ? GDSP_ACC_I32 : GDSP_ACC_F64;
		/* real (non-integer) values: every cell must see the reference's additions in file order.  So must
		 * any input with a non-zero missingVal: a running sum that comes back to missingVal reads as "not
		 * yet covered" and the next interval replaces it (--missing=3 --novalue: depth 4 gives 1) */
		if ((mode == GDSP_ACC_F64 || (clear && missingVal != 0.0))
		 && gd_apply_intervals_exact (&l, clear ? GD_EXACT_CLEAR : GD_EXACT_ADD, missingVal, "input"))
			{ ivlist_free (&l);  return; }
		void* work = gd_work (gdsp_accumulate_work_bytes (gd.genome, gd.cells, mode));
		gd_check (gdsp_accumulate_host (gd.ctx, gd.genome, gd.sig, gd.cells, work, l.seg, l.start, l.end,
		                                (valCol == -1) ? NULL : l.val, l.n, mode, clear ? 0 : 1), "input");
		if (clear && missingVal != 0.0)
			{
			ivlist_union (&l);
			gd_set_outside (&l, missingVal);                      /* gd_ops_files.c */
			}
		}
	ivlist_free (&l);
	}

/* ---- output ---------------------------------------------------------------------------------
 * report_intervals: the GPU detects the runs (gdsp_runs), the host formats them.  Chromosomes are
 * written in input order, one chromosome at a time so the run buffers stay small. */

static char*  outBuf = NULL;
static size_t outLen = 0;
#define OUT_CAP (1u << 22)

static inline void out_flush (FILE* f)
	{ if (outLen) { fwrite (outBuf, 1, outLen, f);  outLen = 0; } }

static inline void out_u32 (u32 v)
	{
	char tmp[12];  int n = 0;
	do { tmp[n++] = (char) ('0' + v % 10);  v /= 10; } while (v);
	while (n) outBuf[outLen++] = tmp[--n];
	}

static FILE* outFile = NULL;             /* where out_value sends a value too long for the buffer */

static inline void out_value (int precision, valtype v)
	{
	/* integers with precision 0 are by far the common case; everything else goes through printf
	 * so that rounding (half-even on the binary value) is glibc's, as in the reference */
	if (precision == 0 && v == floor (v) && fabs (v) < 4.0e9 && !(v == 0 && signbit (v)))
		{
		if (v < 0) { outBuf[outLen++] = '-';  v = -v; }
		out_u32 ((u32) v);
		}
	else
		{
		/* room left in the buffer: at least 1024 - (name + two numbers) here.  A value that does not fit
		 * (--precision in the hundreds, DBL_MAX with many decimals) is flushed around: never advance by more
		 * than was written */
		const size_t room = OUT_CAP + 1024 - outLen - 2;
		int n = snprintf (outBuf + outLen, room, valtypeFmtPrec, precision, v);
		if (n >= 0 && (size_t) n < room) outLen += (size_t) n;
		else
			{
			fwrite (outBuf, 1, outLen, outFile);  outLen = 0;
			fprintf (outFile, valtypeFmtPrec, precision, v);
			}
		}
	}

static void out_line (FILE* f, const char* chrom, size_t chromLen, u32 s, u32 e, int kind, int precision, valtype v)
	{
	if (outLen + chromLen + 512 > OUT_CAP) out_flush (f);
	outFile = f;
	memcpy (outBuf + outLen, chrom, chromLen);  outLen += chromLen;
	outBuf[outLen++] = '\t';  out_u32 (s);
	outBuf[outLen++] = '\t';  out_u32 (e);
	if (kind == 1) { outBuf[outLen++] = '\t';  out_value (precision, v); }
	else if (kind == 2) { memcpy (outBuf + outLen, "\tNA", 3);  outLen += 3; }
	outBuf[outLen++] = '\n';
	}

void report_intervals (FILE* f, int precision, int noOutVals, int collapse, int showUncov, int originOne_)
	{
	u32 o = originOne_ ? 1 : 0;
	if (outBuf == NULL) outBuf = (char*) malloc (OUT_CAP + 1024);
	outLen = 0;
	u64 cap = 1u << 20;
	u32 *hS = NULL, *hE = NULL;  valtype* hV = NULL;  u64 hCap = 0;
	void *dS = NULL, *dE = NULL, *dV = NULL;  u64 dCap = 0;
	void *dText = NULL, *hText = NULL;  size_t dTextCap = 0, hTextCap = 0;
	u64 hostFirst = 0;

	for (spec* cs = chromsOfInterest; cs != NULL; cs = cs->next)
		{
		if (trackOperations) tracking_report ("output(%s)\n", cs->chrom);
		int ix = gd_sorted_index (cs);
		u64 nRuns = 0, first[2];
		while (true)
			{
			if (dCap < cap)
				{
				if (dS) { gdsp_free (gd.ctx, dS);  gdsp_free (gd.ctx, dE);  gdsp_free (gd.ctx, dV); }
				gd_check (gdsp_malloc (gd.ctx, cap * 4, &dS), "output");
				gd_check (gdsp_malloc (gd.ctx, cap * 4, &dE), "output");
				gd_check (gdsp_malloc (gd.ctx, cap * 8, &dV), "output");
				dCap = cap;
				}
			int st = gdsp_runs (gd.ctx, gd.single[ix], gd.sig, collapse, showUncov,
			                    (u32*) dS, (u32*) dE, (double*) dV, dCap, &nRuns, first);
			if (st == GDSP_ERR_CAPACITY) { cap = nRuns + 1024;  continue; }
			gd_check (st, "output");
			break;
			}
		/* device formatter (gdsp_format_runs): the text of the chromosome is produced on the GPU chunk by
		 * chunk and only copied out and written here.  NA gap lines and values the device declines
		 * (NaN, infinities, >= 2^63) take the host loop below. */
		if (nRuns && showUncov != uncovered_NA && !getenv ("GENODSP_HOST_FORMAT"))
			{
			const u64 CHUNK = 4u << 20;
			int declined = false;
			u64 done = 0;
			for (; done < nRuns && !declined; )
				{
				u64 m = (nRuns - done < CHUNK) ? nRuns - done : CHUNK;
				size_t need = gdsp_format_runs_max_bytes (m, cs->chrom);
				if (dTextCap < need)
					{
					if (dText) gdsp_free (gd.ctx, dText);
					gd_check (gdsp_malloc (gd.ctx, need, &dText), "output");
					dTextCap = need;
					}
				u64 bytes = 0;
				gd_check (gdsp_format_runs (gd.ctx, (u32*) dS + done, (u32*) dE + done, (double*) dV + done, m, cs->chrom,
				                            cs->start + o, cs->start, !noOutVals, precision, (char*) dText, dTextCap,
				                            &bytes, &declined), "output");
				if (declined) break;
				if (hTextCap < bytes)
					{
					if (hText) gdsp_free_host (hText);
					hTextCap = bytes + (bytes >> 2) + (1u << 20);
					gd_check (gdsp_malloc_host (hTextCap, &hText), "output");
					}
				gd_check (gdsp_d2h (gd.ctx, hText, dText, bytes), "output");
				out_flush (f);
				fwrite (hText, 1, bytes, f);
				done += m;
				}
			if (!declined) continue;
			/* the host formats what is left of this chromosome */
			hostFirst = done;
			}
		else hostFirst = 0;
		if (hCap < nRuns)
			{
			free (hS);  free (hE);  free (hV);
			hCap = nRuns + 1024;
			hS = (u32*) malloc (hCap * 4);  hE = (u32*) malloc (hCap * 4);  hV = (valtype*) malloc (hCap * 8);
			}
		if (nRuns)
			{
			gd_check (gdsp_d2h (gd.ctx, hS, dS, nRuns * 4), "output");
			gd_check (gdsp_d2h (gd.ctx, hE, dE, nRuns * 4), "output");
			gd_check (gdsp_d2h (gd.ctx, hV, dV, nRuns * 8), "output");
			}
		size_t cl = strlen (cs->chrom);
		u32 prevEnd = 0;
		if (hostFirst > 0 && hostFirst <= nRuns) prevEnd = cs->start + hE[hostFirst - 1];
		for (u64 r = hostFirst; r < nRuns; r++)
			{
			u32 s = cs->start + hS[r], e = cs->start + hE[r];
			if (showUncov == uncovered_NA && s != prevEnd) out_line (f, cs->chrom, cl, prevEnd + o, s, 2, 0, 0.0);
			out_line (f, cs->chrom, cl, s + o, e, noOutVals ? 0 : 1, precision, hV[r]);
			prevEnd = e;
			}
		if (showUncov == uncovered_NA && cs->start + cs->length != prevEnd)
			out_line (f, cs->chrom, cl, prevEnd + o, cs->start + cs->length, 2, 0, 0.0);
		}
	out_flush (f);
	if (trackOperations) tracking_report ("output(--done--)\n");
	if (dS) { gdsp_free (gd.ctx, dS);  gdsp_free (gd.ctx, dE);  gdsp_free (gd.ctx, dV); }
	if (dText) gdsp_free (gd.ctx, dText);
	if (hText) gdsp_free_host (hText);
	free (hS);  free (hE);  free (hV);
	}

void read_all_chromosomes (char* filename)
	{
	FILE* f = fopen (filename, "rt");
	if (f == NULL) { fprintf (stderr, "can't open \"%s\" for reading\n", filename);  exit (EXIT_FAILURE); }
	if (trackOperations) fprintf (stderr, "read_all(%s)\n", filename);
	int save = trackOperations;  trackOperations = false;
	read_intervals (f, 4-1, false, ri_overlapSum, true, 0.0);
	trackOperations = save;
	fclose (f);
	}

void write_all_chromosomes (char* filename)
	{
	FILE* f = fopen (filename, "wt");
	if (f == NULL) { fprintf (stderr, "can't open \"%s\" for writing\n", filename);  exit (EXIT_FAILURE); }
	if (trackOperations) fprintf (stderr, "write_all(%s)\n", filename);
	int save = trackOperations;  trackOperations = false;
	report_intervals (f, 10, false, true, false, false);
	trackOperations = save;
	fclose (f);
	}

/* ---- option parsing ----------------------------------------------------------------------- */

static int process_operator_options (int argc, char** argv)
	{
	char* arg = argv[0];
	int   consumed = 0;
	char* dspName;

	if (arg[0] == specialPipeChar && arg[1] == 0) { argv++;  argc--;  consumed++;  dspName = argv[0]; }
	else dspName = skip_whitespace (arg + 1);
	if (dbgPipe) fprintf (stderr, "  dspName=\"%s\"\n", dspName);

	dspinfo* info = find_operator (dspName);
	if (info == NULL) chastise ("\"%s\" is not a known operation\n", dspName);
	argv++;  argc--;  consumed++;

	int nargs = 0;
	while (nargs < argc && argv[nargs][0] != specialPipeChar) nargs++;
	consumed += nargs;

	chastiseUsage = info->funcUsage;  chastiseUsageName = info->name;
	dspop* op = (*info->funcParse) (info->name, nargs, argv);
	chastiseUsage = NULL;  chastiseUsageName = NULL;

	op->name      = copy_string (info->name);
	op->funcApply = info->funcApply;
	op->funcFree  = info->funcFree;
	op->next      = NULL;
	if (tailOp == NULL) pipeline = op; else tailOp->next = op;
	tailOp = op;
	if (dbgPipe) fprintf (stderr, "  argsConsumed=%d\n", consumed);
	return consumed;
	}

static void help_for (char* name)
	{
	dspinfo* info = find_operator (name);
	if (info == NULL) { fprintf (stderr, "\"%s\" is not a known operation\n", name);  exit (EXIT_FAILURE); }
	fprintf (stderr, "=== %s ===\n", info->name);
	(*info->funcUsage) (info->name, stderr, "  ");
	exit (EXIT_SUCCESS);
	}

static void parse_options (int _argc, char** _argv)
	{
	int    argc = _argc - 1;
	char** argv = _argv + 1;
	char*  chromsFilename = NULL;

	if (argc == 0) chastise (NULL);
	while (argc > 0)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');
		if (argVal != NULL) argVal++;

		if (arg[0] == specialPipeChar)
			{
			if (argc == 1 && arg[1] == 0)
				chastise ("%c at end of command line, with no operation\n", specialPipeChar);
			int used = process_operator_options (argc, argv);
			argv += used - 1;  argc -= used - 1;
			}
		else if (strcmp_prefix (arg, "--chromosomes=") == 0 || strcmp_prefix (arg, "--chroms=") == 0)
			chromsFilename = argVal;
		else if (strcmp (arg, "--novalue") == 0 || strcmp (arg, "--novalues") == 0 || strcmp (arg, "--value=none") == 0)
			{ valColumn = -1;  set_named_global ("valColumn", (valtype) valColumn); }
		else if (strcmp_prefix (arg, "--value=") == 0)
			{
			valColumn = string_to_int (argVal) - 1;
			if (valColumn == -1) chastise ("value column can't be 0 (\"%s\")\n", arg);
			if (valColumn < 0)   chastise ("value column can't be negative (\"%s\")\n", arg);
			if (valColumn < 3)   chastise ("value column can't be 1, 2 or 3 (\"%s\")\n", arg);
			set_named_global ("valColumn", (valtype) valColumn);
			}
		else if (strcmp (arg, "--nooutputvalue") == 0 || strcmp (arg, "--nooutputvalues") == 0)
			{ noOutputValues = true;  set_named_global ("noOutputValues", (valtype) noOutputValues); }
		else if (strcmp_prefix (arg, "--precision=") == 0)
			{
			valPrecision = string_to_int (argVal);
			if (valPrecision < 0) chastise ("precision can't be negative (\"%s\")\n", arg);
			set_named_global ("valPrecision", (valtype) valPrecision);
			}
		else if (strcmp (arg, "--nocollapse") == 0)
			{ collapseRuns = false;  set_named_global ("collapseRuns", (valtype) collapseRuns); }
		else if (strcmp (arg, "--uncovered:hide") == 0 || strcmp (arg, "--hide:uncovered") == 0)
			{ showUncovered = uncovered_hide;  set_named_global ("showUncovered", (valtype) showUncovered); }
		else if (strcmp (arg, "--uncovered:show") == 0 || strcmp (arg, "--show:uncovered") == 0)
			{ showUncovered = uncovered_show;  set_named_global ("showUncovered", (valtype) showUncovered); }
		else if (strcmp (arg, "--uncovered:NA") == 0 || strcmp (arg, "--uncovered:mark") == 0
		      || strcmp (arg, "--mark:uncovered") == 0 || strcmp (arg, "--markgaps") == 0)
			{ showUncovered = uncovered_NA;  set_named_global ("showUncovered", (valtype) showUncovered); }
		else if (strcmp (arg, "--cliptochromosome") == 0 || strcmp (arg, "--cliptochrom") == 0
		      || strcmp (arg, "--cliptolength") == 0 || strcmp (arg, "--clip") == 0)
			clipToLength = true;
		else if (strcmp (arg, "--origin=one") == 0 || strcmp (arg, "--origin=1") == 0)
			{ originOne = true;  set_named_global ("originOne", (valtype) originOne); }
		else if (strcmp (arg, "--origin=zero") == 0 || strcmp (arg, "--origin=0") == 0)
			{ originOne = false;  set_named_global ("originOne", (valtype) originOne); }
		else if (strcmp (arg, "--nooutput") == 0)
			inhibitOutput = true;
		else if (strcmp_prefix (arg, "--window=") == 0 || strcmp_prefix (arg, "W=") == 0 || strcmp_prefix (arg, "--W=") == 0)
			{
			int w = string_to_unitized_int (argVal, true);
			if (w == 0) chastise ("window size can't be zero (\"%s\")\n", arg);
			if (w < 0)  chastise ("window size can't be negative (\"%s\")\n", arg);
			set_named_global ("windowSize", (valtype) w);
			}
		else if (strcmp (arg, "?") == 0)
			usage_operations ();
		else if (strcmp_prefix (arg, "?=") == 0 || strcmp_prefix (arg, "--help=") == 0)
			{
			if (strcmp (argVal, "*") == 0) goto help_all;
			help_for (argVal);
			}
		else if (strcmp_prefix (arg, "?") == 0)
			help_for (arg + 1);
		else if (strcmp (arg, "--help") == 0)
			{
		help_all:
			for (u32 ix = 0; ix < dspTableLen; ix++)
				{
				if (dspTable[ix].funcShort == NULL) continue;
				fprintf (stderr, "=== %s ===\n", dspTable[ix].name);
				(*dspTable[ix].funcUsage) (dspTable[ix].name, stderr, "  ");
				}
			exit (EXIT_SUCCESS);
			}
		else if (strcmp (arg, "--report=comments") == 0 || strcmp (arg, "--report:comments") == 0)
			reportComments = true;
		else if (strcmp_prefix (arg, "--progress=input:") == 0 || strcmp_prefix (arg, "--progress:input=") == 0
		      || strcmp_prefix (arg, "--progress:input:") == 0)
			{
			if (strcmp_prefix (argVal, "input:") == 0) argVal = strchr (arg, ':') + 1;
			reportInputProgress = (u32) string_to_unitized_int (argVal, true);
			}
		else if (strcmp (arg, "--progress=operations") == 0 || strcmp (arg, "--progress:operations") == 0
		      || strcmp (arg, "--debug=operations") == 0)
			trackOperations = true;
		else if (strcmp (arg, "--version") == 0)
			{
			fprintf (stderr, "%s (version %s, sm_100a build of genodsp 0.0.10 released 20220616)\n", programName, programVersion);
			exit (EXIT_SUCCESS);
			}
		else if (strcmp (arg, "--debug=input") == 0)   dbgInput = true;
		else if (strcmp (arg, "--debug=pipe") == 0)    dbgPipe = true;
		else if (strcmp (arg, "--debug=globals") == 0) dbgGlobals = true;
		else if (strcmp_prefix (arg, "--") == 0)
			chastise ("Can't understand \"%s\"\n", arg);
		else
			{
			/* <chromosome>:<length> or <chromosome>:<start>:<end> (origin zero, half open) */
			char* c1 = strchr (arg, ':');
			if (c1 == NULL)
				{
				fprintf (stderr, "\"%s\" contains no chromosome length\n"
				                 "(expected \"chromosome:length\" or \"chromosome:start:end\")\n", arg);
				exit (EXIT_FAILURE);
				}
			char* c2 = strchr (c1 + 1, ':');
			u32 cStart = 0, cLen;
			*(c1++) = 0;
			if (c2 == NULL) cLen = (u32) string_to_u32 (c1);
			else { *(c2++) = 0;  cStart = (u32) string_to_u32 (c1);  cLen = (u32) string_to_u32 (c2) - cStart; }
			if (!add_chromosome_spec (arg, cStart, cLen)) chastise ("can't specify %s more than once\n", arg);
			}
		argv++;  argc--;
		}
	if (chromsFilename != NULL) read_chromosome_lengths (chromsFilename);
	if (chromsOfInterest == NULL) chastise ("gotta give me some chromosome names\n");
	}

/* ---- executor ------------------------------------------------------------------------------
 * The reference runs every operator as its own pass, chromosome by chromosome
 * (genodsp.c:900-936).  Here each operator covers the whole packed genome in one launch, and a run
 * of consecutive pointwise operators (gd_ops.h: gd_pointwise_descriptor) becomes ONE kernel. */

dspop* gd_pipeline_head (void) { return pipeline; }
int    gd_output_inhibited (void) { return inhibitOutput; }

/* run the operators [first, stop): consecutive pointwise ones as a single fused launch */
/* NVTX ranges around every operator of the pipeline (and around a fused run of pointwise operators), so that a
 * timeline (nsys) or an ncu report shows which kernels belong to which command-line operator.  Header-only
 * (nvtx3): without a profiler attached a push/pop is a pointer test.  Built when the CUDA headers are present. */
#ifdef GDSP_NVTX
#include <nvtx3/nvToolsExt.h>
#define GD_RANGE_PUSH(name) nvtxRangePushA (name)
#define GD_RANGE_POP()      nvtxRangePop ()
#else
#define GD_RANGE_PUSH(name) ((void) 0)
#define GD_RANGE_POP()      ((void) 0)
#endif

static void exec_range (dspop* first, dspop* stop)
	{
	dspop* op = first;
	while (op != stop)
		{
		gdsp_pw_op prog[GDSP_MAX_POINTWISE];
		gd_pw_resources res[GDSP_MAX_POINTWISE];
		int n = 0, k = 0;
		dspop* scan = op;
		while (scan != stop && k < GDSP_MAX_POINTWISE && gd_is_pointwise (scan))
			{
			int got = gd_pointwise_descriptor (scan, &prog[n], &res[n]);     /* 0 when the operator is a no-op */
			n += got;  k++;
			scan = scan->next;
			}
		if (scan != op)
			{
			GD_RANGE_PUSH ("pointwise (fused)");
			if (n > 0) gd_check (gdsp_pointwise (gd.ctx, gd.genome, gd.sig, gd.sig, prog, n), "pointwise");
			GD_RANGE_POP ();
			for (int i = 0; i < n; i++) gd_pw_release (&res[i]);
			op = scan;
			continue;
			}
		GD_RANGE_PUSH (op->name);
		if (op->atRandom || gd_is_genome_capable (op))
			(*op->funcApply) (op, "*", gd.maxLength, NULL);
		else
			/* an operator written against the reference contract: one chromosome vector at a time */
			for (int i = 0; i < gd.nchrom; i++)
				(*op->funcApply) (op, chromsSorted[i]->chrom, chromsSorted[i]->length, chromsSorted[i]->valVector);
		GD_RANGE_POP ();
		op = op->next;
		}
	}

static void run_pipeline (void)
	{
	if (!trackOperations) { exec_range (pipeline, NULL);  return; }

	/* --progress=operations: reproduce the reference's trace (genodsp.c:900-936): a group of
	 * per-chromosome operators is listed chromosome by chromosome, variables are resolved (and
	 * announced) when an operator first meets a chromosome, whole-genome operators print op(*) */
	dspop* first = pipeline;
	while (first != NULL)
		{
		dspop* stop = first;
		while (stop != NULL && !stop->atRandom) stop = stop->next;
		if (stop != first)
			{
			for (int i = 0; i < gd.nchrom; i++)
				for (dspop* op = first; op != stop; op = op->next)
					{
					fprintf (stderr, "%s(%s)\n", op->name, chromsSorted[i]->chrom);
					if (i == 0) gd_resolve_variables (op);
					}
			exec_range (first, stop);
			}
		if (stop == NULL) break;
		tracking_report ("%s(*)\n", stop->name);
		exec_range (stop, stop->next);
		first = stop->next;
		}
	}

/* GENODSP_TIMING=1: wall clock of the program's phases on stderr (not part of the reference's output) */
#include <time.h>
static double phase_clock (void)
	{ struct timespec t;  clock_gettime (CLOCK_MONOTONIC, &t);  return t.tv_sec + 1e-9 * t.tv_nsec; }
static double phaseT0;
static void phase_mark (const char* what)
	{
	static int on = -1;
	if (on < 0) on = (getenv ("GENODSP_TIMING") != NULL);
	if (!on) return;
	double now = phase_clock ();
	if (what != NULL) fprintf (stderr, "[timing] %-12s %.3f s\n", what, now - phaseT0);
	phaseT0 = now;
	}

int main (int argc, char** argv)
	{
	phase_mark (NULL);
	set_named_global ("valColumn",     (valtype) valColumn);
	set_named_global ("valPrecision",  (valtype) valPrecision);
	set_named_global ("collapseRuns",  (valtype) collapseRuns);
	set_named_global ("showUncovered", (valtype) showUncovered);
	set_named_global ("originOne",     (valtype) originOne);

	parse_options (argc, argv);
	sort_chromosomes_by_length ();
	for (int i = 0; chromsSorted[i] != NULL; i++)
		if (chromsSorted[i]->length == 0)
			{ fprintf (stderr, "no length was specified for %s\n", chromsSorted[i]->chrom);  return EXIT_FAILURE; }

	if (trackOperations)
		for (int i = 0; chromsSorted[i] != NULL; i++)
			tracking_report ("allocate(%s / %s bytes)\n", chromsSorted[i]->chrom, ucommatize (chromsSorted[i]->length));
	phase_mark ("parse args");
	gd_device_open ();                        /* CUDA initialisation continues on a helper thread ... */
	if (trackOperations) tracking_report ("allocate(--done--)\n");
	phase_mark ("device start");

	if (pipeline == NULL || strcmp (pipeline->name, "input") != 0)
		read_intervals (stdin, valColumn, originOne, ri_overlapSum, false, 0.0);
	gd_device_wait ();                        /* ... and is joined after the text has been parsed */
	if (getenv ("GENODSP_TIMING") != NULL) gdsp_sync (gd.ctx);
	phase_mark ("input");

	run_pipeline ();
	if (getenv ("GENODSP_TIMING") != NULL) gdsp_sync (gd.ctx);
	phase_mark ("operators");

	if (!inhibitOutput)
		report_intervals (stdout, valPrecision, noOutputValues, collapseRuns, showUncovered, originOne);
	phase_mark ("output");

	gdsp_sync (gd.ctx);
	free_scratch_vectors ();
	free_named_globals ();
	dspop* next;
	for (dspop* op = pipeline; op != NULL; op = next)
		{
		next = op->next;
		if (op->name != NULL) { free (op->name);  op->name = NULL; }
		(*op->funcFree) (op);
		}
	gd_device_close ();
	spec* nextSpec;
	for (spec* s = chromsOfInterest; s != NULL; s = nextSpec)
		{ nextSpec = s->next;  free (s->chrom);  free (s); }
	chromsOfInterest = NULL;
	free (chromsSorted);  chromsSorted = NULL;
	return EXIT_SUCCESS;
	}

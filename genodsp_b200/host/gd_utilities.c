/* gd_utilities.c -- string/number helpers (behaviour of the reference's
 * utilities.c: K/M/G suffixes :236-309, "inf"/"1/inf" literals :334-370,
 * comma grouping :501-541).  Error texts match the reference so that scripts
 * that parse stderr keep working. */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <ctype.h>
#include <limits.h>
#include "gd_utilities.h"

static void die (const char* fmt, const char* s)
	{
	fprintf (stderr, fmt, s);
	exit (EXIT_FAILURE);
	}

char* copy_string (const char* s)
	{
	if (s == NULL) return NULL;
	size_t n = strlen (s) + 1;
	char* t = (char*) malloc (n);
	if (t == NULL)
		{
		fprintf (stderr, "failed to allocate %lld bytes to copy \"%s\"\n", (long long) n, s);
		exit (EXIT_FAILURE);
		}
	memcpy (t, s, n);
	return t;
	}

int strcmp_prefix (const char* str1, const char* str2)
	{ return strncmp (str1, str2, strlen (str2)); }

int strcmp_suffix (const char* str1, const char* str2)
	{
	size_t n1 = strlen (str1), n2 = strlen (str2);
	return (n2 <= n1) ? strcmp (str1 + n1 - n2, str2) : strcmp (str1, str2);
	}

static const char* skip_blanks (const char* s)
	{
	while (*s == ' ' || *s == '\t' || *s == '\n') s++;
	return s;
	}

int string_to_int (const char* s)
	{
	const char* t = skip_blanks (s);
	int v;  char extra;
	if (*t == 0) die ("an empty string is not an integer\n%s", "");
	if (sscanf (t, "%d%c", &v, &extra) != 1) die ("\"%s\" is not an integer\n", s);
	if ((v < 0 && *t != '-') || (v > 0 && *t == '-'))
		die ("\"%s\" is outside the range of a signed integer\n", s);
	return v;
	}

int string_to_u32 (const char* s)
	{
	const char* t = skip_blanks (s);
	u32 v;  char extra;
	if (*t == 0) die ("an empty string is not an unsigned integer\n%s", "");
	if (*t == '-' || sscanf (t, "%u%c", &v, &extra) != 1) die ("\"%s\" is not an unsigned integer\n", s);
	return (int) v;
	}

/* integer with optional K/M/G suffix; fractional values ("1.5M") allowed */
int string_to_unitized_int (const char* s, int byThousands)
	{
	char  tmp[20];
	const char* parse = s;
	int   len = (int) strlen (s), mult = 1, v = 0;
	float vf;  char extra;

	if (len < (int) sizeof (tmp))
		{
		strcpy (tmp, s);
		parse = tmp;
		if (len > 0)
			{
			switch (tmp[len-1])
				{
				case 'K': case 'k': mult = byThousands ? 1000       : 1024;               break;
				case 'M': case 'm': mult = byThousands ? 1000000    : 1024 * 1024;        break;
				case 'G': case 'g': mult = byThousands ? 1000000000 : 1024 * 1024 * 1024; break;
				}
			if (mult != 1) tmp[len-1] = 0;
			}
		}
	if (sscanf (parse, "%d%c", &v, &extra) == 1)
		{
		if (mult != 1)
			{
			if ((v > 0 && v > INT_MAX / mult) || (v < 0 && -v > INT_MAX / mult))
				die ("\"%s\" is out of range for an integer\n", s);
			v *= mult;
			}
		return v;
		}
	if (sscanf (parse, "%f%c", &vf, &extra) != 1) die ("\"%s\" is not an integer\n", s);
	if ((vf > 0 && vf * mult > INT_MAX) || (vf < 0 && -vf * mult > INT_MAX))
		die ("\"%s\" is out of range for an integer\n", s);
	return (int) ((vf * mult) + .5);
	}

int try_string_to_double (const char* s, double* out)
	{
	const char* t = skip_blanks (s);
	double v;  char extra;
	if (*t == 0) return false;
	if      (strcmp (s, "inf")    == 0 || strcmp (s, "+inf")   == 0) v =  DBL_MAX;
	else if (strcmp (s, "-inf")   == 0)                              v = -DBL_MAX;
	else if (strcmp (s, "1/inf")  == 0 || strcmp (s, "+1/inf") == 0) v =  DBL_MIN;
	else if (strcmp (s, "-1/inf") == 0)                              v = -DBL_MIN;
	else if (sscanf (s, "%lf%c", &v, &extra) != 1) return false;
	if (out != NULL) *out = v;
	return true;
	}

double string_to_double (const char* s)
	{
	double v;
	if (*skip_blanks (s) == 0) die ("an empty string is not a number\n%s", "");
	if (!try_string_to_double (s, &v)) die ("\"%s\" is not a number\n", s);
	return v;
	}

char* skip_whitespace (char* s) { while (*s != 0 &&  isspace ((unsigned char) *s)) s++;  return s; }
char* skip_darkspace  (char* s) { while (*s != 0 && !isspace ((unsigned char) *s)) s++;  return s; }

/* decimal with thousands separators; five rotating static buffers like the reference */
char* ucommatize (const u64 v)
	{
	static char buf[5][52];
	static int  which = 4;
	char digits[32];
	which = (which + 1) % 5;
	char* out = buf[which];
	int n = snprintf (digits, sizeof (digits), "%jd", (intmax_t) v);
	int lead = n % 3, o = 0;
	for (int i = 0; i < n; i++)
		{
		if (i > 0 && (i - lead) % 3 == 0 && digits[i-1] != '-') out[o++] = ',';
		out[o++] = digits[i];
		}
	out[o] = 0;
	return out;
	}

void safe_strncpy (char* dest, const char* src, size_t n)
	{
	if (n == 0) return;
	size_t k = 0;
	for (; k + 1 < n && src[k] != 0; k++) dest[k] = src[k];
	for (; k < n; k++) dest[k] = 0;
	}

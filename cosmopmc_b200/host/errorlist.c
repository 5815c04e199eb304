/* errorlist.c -- error stack + allocation helpers of the pmclib host API
 * (include/pmctools/errorlist.h). */
#include "pmctools/errorlist.h"
#include <stdarg.h>

error *newError(int errV, const char *where, const char *text, error *prev)
{
   error *e = (error *)malloc(sizeof(error));
   if (!e) { fprintf(stderr, "errorlist: out of memory\n"); exit(-1); }
   e->errValue = errV;
   snprintf(e->errWhere, WHR_SZ, "%s", where ? where : "");
   snprintf(e->errText, TXT_SZ, "%s", text ? text : "");
   e->next = prev;
   return e;
}

error *newErrorVA(int errV, const char *where, const char *fmt, error *prev, ...)
{
   char txt[TXT_SZ];
   va_list ap;
   va_start(ap, prev);
   vsnprintf(txt, TXT_SZ, fmt, ap);
   va_end(ap);
   return newError(errV, where, txt, prev);
}

int _isError(const error *err) { return err != NULL; }

/* value of the originating (deepest non-forward) error */
int getErrorValue(const error *err)
{
   int v = noErr;
   for (; err; err = err->next)
      if (err->errValue != forwardErr) v = err->errValue;
   return v;
}

void printError(FILE *F, const error *err)
{
   for (; err; err = err->next) {
      if (err->errValue == forwardErr) fprintf(F, "  forwarded at %s\n", err->errWhere);
      else fprintf(F, "Error %d at %s: %s\n", err->errValue, err->errWhere, err->errText);
   }
}

void stringError(char *str, const error *err)
{
   str[0] = 0;
   for (; err; err = err->next)
      if (err->errValue != forwardErr) {
         snprintf(str, TXT_SZ, "Error %d at %s: %s", err->errValue, err->errWhere, err->errText);
      }
}

void purgeError(error **err)
{
   while (err && *err) { error *n = (*err)->next; free(*err); *err = n; }
}

void endError(error **err) { purgeError(err); }
error *unmanagedError(void) { return NULL; }

void *malloc_err(size_t sz, error **err)
{
   void *p = malloc(sz ? sz : 1);
   if (!p) *err = addErrorVA(io_alloc, "Cannot allocate %zu bytes", *err, __LINE__, sz);
   return p;
}

void *calloc_err(size_t n, size_t sz, error **err)
{
   void *p = calloc(n ? n : 1, sz ? sz : 1);
   if (!p) *err = addErrorVA(io_alloc, "Cannot allocate %zu x %zu bytes", *err, __LINE__, n, sz);
   return p;
}

void *realloc_err(void *q, size_t sz, error **err)
{
   void *p = realloc(q, sz ? sz : 1);
   if (!p) *err = addErrorVA(io_alloc, "Cannot reallocate %zu bytes", *err, __LINE__, sz);
   return p;
}

FILE *fopen_err(const char *name, const char *mode, error **err)
{
   FILE *F = fopen(name, mode);
   if (!F) *err = addErrorVA(io_file, "Cannot open file '%s' (mode %s)", *err, __LINE__, name, mode);
   return F;
}

/* pmctools/errorlist.h -- error stack of the pmclib host API, as used by every
 * in-scope reference source (257 forwardError uses; semantics
 * Manual/manual.tex:2815-2884).  Convention: `error **err` is the last
 * argument; a callee appends to the list and returns a dummy value; the
 * caller tests with forwardError / quitOnError.  Part of the B200-native
 * replacement for pmclib. */
#ifndef PMCTOOLS_ERRORLIST_H
#define PMCTOOLS_ERRORLIST_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WHR_SZ 256
#define TXT_SZ 2048

typedef struct _pmc_error {
  int  errValue;
  char errWhere[WHR_SZ];
  char errText[TXT_SZ];
  struct _pmc_error *next;
} error;

#define noErr     0
#define forwardErr (-123456789)

error *newError(int errV, const char *where, const char *text, error *prev);
error *newErrorVA(int errV, const char *where, const char *fmt, error *prev, ...);
int   _isError(const error *err);
int   getErrorValue(const error *err);
void  printError(FILE *F, const error *err);
void  stringError(char *str, const error *err);
void  purgeError(error **err);
void  endError(error **err);
error *unmanagedError(void);

#define _DEBUGHERE_(fmt, ...) fprintf(stderr, "%s:%d " fmt "\n", __FILE__, __LINE__, __VA_ARGS__)
#define fprintfDEBUG(F, ...) fprintf(F, __VA_ARGS__)

#define _PMC_STR2(x) #x
#define _PMC_STR(x) _PMC_STR2(x)
#define _PMC_WHERE(li) __FILE__ ":" _PMC_STR(li)

#define isError(err) (_isError(err))
#define addError(errV, txt, prev, li) newError((errV), _PMC_WHERE(li), (txt), (prev))
#define addErrorVA(errV, fmt, prev, li, ...) newErrorVA((errV), _PMC_WHERE(li), (fmt), (prev), __VA_ARGS__)
#define topError(errV, txt, li) newError((errV), _PMC_WHERE(li), (txt), NULL)

/* The macros are plain brace blocks (not do-while): the reference uses them both
 * with and without a trailing semicolon (wrappers/src/param.c:1345). */
#define forwardError(err, li, ret) \
  { if (_isError(err)) { (err) = newError(forwardErr, _PMC_WHERE(li), "", (err)); return ret; } }
#define forwardErrorNoReturn(err, li) \
  { if (_isError(err)) { (err) = newError(forwardErr, _PMC_WHERE(li), "", (err)); } }
#define testErrorRet(test, errV, txt, err, li, ret) \
  { if (test) { (err) = newError((errV), _PMC_WHERE(li), (txt), (err)); return ret; } }
#define testErrorRetVA(test, errV, fmt, err, li, ret, ...) \
  { if (test) { (err) = newErrorVA((errV), _PMC_WHERE(li), (fmt), (err), __VA_ARGS__); return ret; } }
#define testError(test, errV, txt, err, li) testErrorRet(test, errV, txt, err, li, )
#define exitOnError(err, F) \
  { if (_isError(err)) { printError((F), (err)); exit(getErrorValue(err)); } }
#define quitOnError(err, li, F) \
  { if (_isError(err)) { (err) = newError(forwardErr, _PMC_WHERE(li), "", (err)); printError((F), (err)); exit(getErrorValue(err)); } }
#define quitOnErrorStr(err, li, F, str) \
  { if (_isError(err)) { fprintf((F), "%s\n", (str)); quitOnError(err, li, F); } }
/* print and drop the error, continue (manual.tex:507-520) */
#define ParameterErrorVerb(err, param, quiet, ndim) \
  { if (_isError(err)) { if (!(quiet)) { int i_; fprintf(stderr, "Error at parameter ("); \
         for (i_ = 0; i_ < (int)(ndim); i_++) fprintf(stderr, "%g ", (param)[i_]); fprintf(stderr, "): "); \
         printError(stderr, (err)); } purgeError(&(err)); } }

/* allocation / file helpers (pmctools/io.h in upstream) */
void *malloc_err(size_t sz, error **err);
void *calloc_err(size_t n, size_t sz, error **err);
void *realloc_err(void *p, size_t sz, error **err);
FILE *fopen_err(const char *name, const char *mode, error **err);

/* error-code bases used by the in-scope reference files */
#define io_base      (-200)
#define io_alloc     (-1 + io_base)
#define io_file      (-2 + io_base)
#define io_eof       (-3 + io_base)
#define mv_base      (-300)
#define mv_allocate  (-1 + mv_base)
#define mv_serialize (-2 + mv_base)
#define mv_outOfBound (-3 + mv_base)
#define mv_dimension (-4 + mv_base)
#define mv_cholesky  (-5 + mv_base)
#define mv_negative  (-6 + mv_base)
#define mv_negWeight (-7 + mv_base)
#define mv_file      (-8 + mv_base)
#define pmc_base     (-6000)
#define pmc_allocate (-1 + pmc_base)
#define pmc_serialize (-2 + pmc_base)
#define pmc_outOfBound (-3 + pmc_base)
#define pmc_badComm  (-4 + pmc_base)
#define pmc_negWeight (-5 + pmc_base)
#define pmc_cholesky (-6 + pmc_base)
#define pmc_negative (-7 + pmc_base)
#define pmc_undef    (-8 + pmc_base)
#define pmc_file     (-9 + pmc_base)
#define pmc_dimension (-10 + pmc_base)
#define pmc_type     (-11 + pmc_base)
#define pmc_negHatCl (-12 + pmc_base)
#define pmc_infnan   (-13 + pmc_base)
#define pmc_incompat (-14 + pmc_base)
#define pmc_nosamplep (-15 + pmc_base)
#define pmc_sort     (-16 + pmc_base)
#define pmc_infinite (-17 + pmc_base)
#define pmc_isLog    (-18 + pmc_base)
#define pmc_tooManySteps (-19 + pmc_base)
#define math_base         (-500)
#define math_negative     (-1 + math_base)
#define math_singularValue (-2 + math_base)
#define math_tooManySteps (-3 + math_base)
#define math_underflow    (-4 + math_base)
#define math_infnan       (-5 + math_base)
#define math_wrongValue   (-6 + math_base)
#define math_alloc        (-7 + math_base)
#define math_interpoloutofrange (-8 + math_base)
#define math_interpol2small (-9 + math_base)
#define math_interpol2big (-10 + math_base)
#define math_stackTooSmall (-11 + math_base)
#define math_overflow     (-12 + math_base)
#define math_unknown      (-13 + math_base)

#define tls_base     (-700)
#define pb_base      (-6100)
#define pb_allocate  (-1 + pb_base)
#define pb_outOfBound (-2 + pb_base)

#ifdef __cplusplus
}
#endif
#endif

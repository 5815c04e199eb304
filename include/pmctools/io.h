/* pmctools/io.h -- small I/O helpers of pmclib used by the reference's tools/
 * and wrappers/ (tools/src/config.c:117, wrappers/src/sn.c:80,
 * wrappers/src/param.c:17, exec/cosmo_pmc.c:745). */
#ifndef PMCTOOLS_IO_H
#define PMCTOOLS_IO_H
#include <stdio.h>
#include <time.h>
#include <sys/times.h>
#include <sys/stat.h>
#include <unistd.h>
#include "errorlist.h"
#ifdef __cplusplus
extern "C" {
#endif
unsigned int numberoflines(const char *name, error **err);
unsigned int numberoflines_comments(const char *name, unsigned int *ncomment, error **err);
void chomp(char *line);                                  /* strip the trailing newline */
void print_parameter(FILE *where, size_t npar, const double *params);
time_t start_time(FILE *FOUT);
void end_time(time_t t_start, FILE *FOUT);
void read_double(char **str, double *x, error **err);
#ifdef __cplusplus
}
#endif
#endif

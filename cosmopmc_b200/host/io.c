/* io.c -- pmctools/io.h helpers. */
#include "pmctools/io.h"
#include <string.h>

unsigned int numberoflines_comments(const char *name, unsigned int *ncomment, error **err)
{
   FILE *F = fopen_err(name, "r", err);
   forwardError(*err, __LINE__, 0);
   char line[16384];
   unsigned int n = 0, nc = 0;
   while (fgets(line, sizeof(line), F)) {
      char *s = line;
      while (*s == ' ' || *s == '\t') s++;
      if (*s == '#') nc++;
      else if (*s != '\n' && *s != 0) n++;
   }
   fclose(F);
   if (ncomment) *ncomment = nc;
   return n;
}

unsigned int numberoflines(const char *name, error **err)
{
   unsigned int nc, n = numberoflines_comments(name, &nc, err);
   forwardError(*err, __LINE__, 0);
   return n + nc;
}

void chomp(char *line)
{
   size_t n = strlen(line);
   while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
}

void print_parameter(FILE *where, size_t npar, const double *params)
{
   for (size_t i = 0; i < npar; i++) fprintf(where, "% .5f ", params[i]);
   fprintf(where, "\n");
}

time_t start_time(FILE *FOUT)
{
   time_t t = time(NULL);
   if (FOUT) fprintf(FOUT, "Started at %s", ctime(&t));
   return t;
}

void end_time(time_t t_start, FILE *FOUT)
{
   time_t t = time(NULL);
   if (FOUT) {
      double dt = difftime(t, t_start);
      fprintf(FOUT, "Ended at %s", ctime(&t));
      fprintf(FOUT, "Computation time %.0fs (= %02d:%02d:%02d)\n", dt, (int)(dt / 3600), ((int)dt % 3600) / 60, (int)dt % 60);
   }
}

void read_double(char **str, double *x, error **err)
{
   char *end;
   *x = strtod(*str, &end);
   testErrorRet(end == *str, io_eof, "Cannot read a double", *err, __LINE__, );
   *str = end;
}

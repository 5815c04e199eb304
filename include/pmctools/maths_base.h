/* pmctools/maths_base.h -- constants and tiny macros (pmclib maths_base.h). */
#ifndef PMCTOOLS_MATHS_BASE_H
#define PMCTOOLS_MATHS_BASE_H
#include <math.h>
#define pi     3.14159265358979323846
#define pi_sqr 9.86960440108935861883
#define twopi  6.28318530717958647693
#define ln2pi  1.83787706640934548356
#define arcmin 2.90888208665721580e-4
#define arcsec 4.84813681109535993e-6
#define ABS(a) ((a) < 0 ? -(a) : (a))
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#define SIGN(a, b) ((b) >= 0.0 ? fabs(a) : -fabs(a))
#define DSQR(a) ((a) * (a))
#define dsqr(a) ((a) * (a))
#define DCUB(a) ((a) * (a) * (a))
#define ISQR(a) ((a) * (a))
/* scale factor at the onset of acceleration used by the de_conservative prior volume
 * (wrappers/src/param.c:1091) */
#define a_acc 0.66666666666666663
/* 1-, 2-, 3-sigma confidence levels (exec/exec_helper.c:63-119) */
#define conf_68 0.6827
#define conf_90 0.9
#define conf_95 0.9545
#define conf_99 0.9973
#endif

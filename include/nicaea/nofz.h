/* nicaea/nofz.h -- opaque stub (redshift distributions are out of scope). */
#ifndef NICAEA_NOFZ_H
#define NICAEA_NOFZ_H
typedef int nzmode_t;
typedef struct redshift_stub { int Nzbin; } redshift_t;
#endif

/* pmclib/parabox.h -- axis-aligned parameter box (flat prior support):
 * parabox_from_config, wrappers/src/param.c:915-926; passed to
 * simulate_mix_mvdens at exec/cosmo_pmc.c:320. */
#ifndef PMCLIB_PARABOX_H
#define PMCLIB_PARABOX_H
#include "pmctools/errorlist.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { int ndim; double *min, *max; int *set; } parabox;
parabox *init_parabox(int ndim, error **err);
void add_slab(parabox *pb, int idim, double sinf, double ssup, error **err);
void free_parabox(parabox **pb);
int  isinBox(const parabox *pb, const double *pos, error **err);
#ifdef __cplusplus
}
#endif
#endif

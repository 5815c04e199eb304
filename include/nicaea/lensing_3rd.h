/* nicaea/lensing_3rd.h -- opaque stub (third-order lensing is out of scope). */
#ifndef NICAEA_LENSING_3RD_H
#define NICAEA_LENSING_3RD_H
#include "nicaea/lensing.h"
typedef struct cosmo_3rd_stub { cosmo_lens *lens; } cosmo_3rd;
#endif

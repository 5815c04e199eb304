/* pmclib/tools.h -- misc. pmclib helpers referenced by the reference's headers. */
#ifndef PMCLIB_TOOLS_H
#define PMCLIB_TOOLS_H
#include "pmctools/errorlist.h"
#include "pmctools/maths.h"
#include "pmctools/mvdens.h"
#include "pmclib/pmc.h"
#define tls_cosmo_par (-17 + tls_base)
#define tls_file      (-2 + tls_base)
#define tls_overflow  (-3 + tls_base)
#endif

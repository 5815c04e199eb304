/* out_of_scope_stubs.c -- link stubs for the likelihood modules that the
 * reference's plug-in registry names (wrappers/src/wrappers.c:9-31) but that are
 * outside the PMC hot path of this repository:
 * weak lensing, n(z), halo model / galaxy clustering, photo-z cross-correlation,
 * topology.  Selecting one of them in a config file raises wr_undef-style errors. */
#include "param.h"
#include "all_wrappers.h"

#define STUB(name)                                                                                 \
   functions_wrapper_t *name(error **err)                                                          \
   {                                                                                               \
      *err = addError(mk_undef, #name ": this data type is outside the B200 PMC hot path", *err, __LINE__); \
      return NULL;                                                                                 \
   }
STUB(init_functions_Lensing)
STUB(init_functions_Nz)
STUB(init_functions_topo)

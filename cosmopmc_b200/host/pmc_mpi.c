/* pmc_mpi.c -- pmclib/pmc_mpi.h for nproc == 1: the scatter/gather of the
 * reference's master/worker scheme degenerates to identities (the B200-native
 * design shards by Philox counter and needs no sample traffic at all). */
#include "pmclib/pmc_mpi.h"

int send_simulation(pmc_simu *psim, int nproc, error **err)
{
   testErrorRetVA(nproc != 1, pmc_badComm, "This build runs one process per GPU (nproc = %d)", *err, __LINE__, 0, nproc);
   return (int)psim->nsamples;
}
void receive_simulation(pmc_simu *psim, int nproc, int myid, error **err)
{
   (void)psim; (void)nproc; (void)myid;
   *err = addError(pmc_badComm, "receive_simulation: no worker ranks in the one-process-per-GPU design", *err, __LINE__);
}
void send_importance_weight(int myid, int nproc, pmc_simu *psim, size_t nok) { (void)myid; (void)nproc; (void)psim; (void)nok; }
size_t receive_importance_weight(pmc_simu *psim, int nproc, size_t master_nok, int master_samples, error **err)
{
   (void)psim; (void)nproc; (void)master_samples; (void)err;
   return master_nok;
}
void send_mix_mvdens(mix_mvdens *m, int nproc, error **err) { (void)m; (void)nproc; (void)err; }
mix_mvdens *receive_mix_mvdens(int myid, int nproc, error **err)
{
   (void)myid; (void)nproc;
   *err = addError(pmc_badComm, "receive_mix_mvdens: no worker ranks in the one-process-per-GPU design", *err, __LINE__);
   return NULL;
}

/* pmclib/pmc_mpi.h -- the MPI scatter/gather helpers of pmclib
 * (exec/cosmo_pmc.c:326-336,355-362,374,387).  The B200-native design has no
 * scatter/gather (every rank draws its own shard; one NCCL all-gather of the EM
 * statistics, include/pmcb200.h), so with nproc == 1 these are identities. */
#ifndef PMCLIB_PMC_MPI_H
#define PMCLIB_PMC_MPI_H
#include <mpi.h>   /* pmclib's pmc_mpi.h pulls MPI in (exec/importance_sample.c:15,176-178) */
#include "pmclib/pmc.h"
#ifdef __cplusplus
extern "C" {
#endif
int  send_simulation(pmc_simu *psim, int nproc, error **err);            /* returns the master's share */
void receive_simulation(pmc_simu *psim, int nproc, int myid, error **err);
void send_importance_weight(int myid, int nproc, pmc_simu *psim, size_t nok);
size_t receive_importance_weight(pmc_simu *psim, int nproc, size_t master_nok, int master_samples, error **err);
void send_mix_mvdens(mix_mvdens *m, int nproc, error **err);
mix_mvdens *receive_mix_mvdens(int myid, int nproc, error **err);
#ifdef __cplusplus
}
#endif
#endif

/* mpi.h -- single-process stub of the four MPI calls made by CosmoPMC's PMC driver
 * (exec/cosmo_pmc.c:580-582,752).  In the B200-native design there is no MPI
 * scatter/gather: one process drives one GPU (or torch.distributed / NCCL launches
 * one process per GPU), so rank = 0 and size = 1 here. */
#ifndef PMCB200_MPI_STUB_H
#define PMCB200_MPI_STUB_H
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = 1; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
/* point-to-point calls of exec/go_fishing.c:179-185,397-408: unreachable with size == 1 */
typedef int MPI_Datatype;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)dest; (void)tag; (void)c; return 1; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s; return 1; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
#endif

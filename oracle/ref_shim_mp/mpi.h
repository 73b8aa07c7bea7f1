/*
 * mpi.h -- multi-process MPI stand-in (TEST INFRASTRUCTURE, not product code).
 *
 * Same purpose and same entry points as oracle/ref_shim/mpi.h (the one-rank stand-in), but for N
 * processes on one host: mpi_multi.c carries messages over Unix-domain sockets (full mesh, eager /
 * buffered sends, a progress engine inside every blocking call) and builds the collectives on
 * them.  It lets oracle/Makefile compile the UNMODIFIED reference into
 * oracle/_ref/libsaena_ref_mp.so and run its real multi-rank code paths -- row partitioning,
 * local/remote split, float halo, repartition / shrink, distributed setup -- in this image, which
 * has no MPI: the multi-rank parity oracle and the multi-core CPU baseline (oracle/mprun.py).
 *
 * Differences from the one-rank header: a datatype handle carries its element kind next to its
 * size ((kind << 24) | bytes), because reductions over several contributions must know what they add.
 */
#ifndef SAENA_B200_ORACLE_MPI_MULTI_H
#define SAENA_B200_ORACLE_MPI_MULTI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;   /* (kind << 24) | extent in bytes */
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Group;
typedef int MPI_Info;
typedef int MPI_Errhandler;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef struct mpi_serial_file *MPI_File;
typedef struct mpi_serial_win *MPI_Win;

typedef struct MPI_Status {
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
    int count_bytes;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_GROUP_NULL 0
#define MPI_GROUP_EMPTY 1
#define MPI_REQUEST_NULL 0
#define MPI_INFO_NULL 0
#define MPI_DATATYPE_NULL 0
#define MPI_OP_NULL 0
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_SOURCE (-2)
#define MPI_ANY_TAG (-1)
#define MPI_PROC_NULL (-3)
#define MPI_IN_PLACE ((void *)-1)
#define MPI_BOTTOM ((void *)0)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MAX_ERROR_STRING 256
#define MPI_TAG_UB 1
#define MPI_KEYVAL_INVALID 0
#define MPI_ERRORS_RETURN 1
#define MPI_ERRORS_ARE_FATAL 0
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

/* datatypes: (kind << 24) | sizeof */
#define SBMPI_K_DERIVED 0
#define SBMPI_K_I8 1
#define SBMPI_K_U8 2
#define SBMPI_K_I16 3
#define SBMPI_K_U16 4
#define SBMPI_K_I32 5
#define SBMPI_K_U32 6
#define SBMPI_K_I64 7
#define SBMPI_K_U64 8
#define SBMPI_K_F32 9
#define SBMPI_K_F64 10
#define SBMPI_K_F128 11
#define SBMPI_K_C128 12
#define SBMPI_K_P_FLOAT_INT 13
#define SBMPI_K_P_DOUBLE_INT 14
#define SBMPI_K_P_LONG_INT 15
#define SBMPI_K_P_SHORT_INT 16
#define SBMPI_K_P_2INT 17
#define SBMPI_K_P_LDOUBLE_INT 18
#define SBMPI_DT(kind, bytes) (((kind) << 24) | (bytes))
#define SBMPI_DT_SIZE(dt) ((dt) & 0xFFFFFF)
#define SBMPI_DT_KIND(dt) (((dt) >> 24) & 0x7F)
#define MPI_CHAR SBMPI_DT(SBMPI_K_I8, 1)
#define MPI_SIGNED_CHAR SBMPI_DT(SBMPI_K_I8, 1)
#define MPI_UNSIGNED_CHAR SBMPI_DT(SBMPI_K_U8, 1)
#define MPI_BYTE SBMPI_DT(SBMPI_K_U8, 1)
#define MPI_CXX_BOOL SBMPI_DT(SBMPI_K_U8, 1)
#define MPI_C_BOOL SBMPI_DT(SBMPI_K_U8, 1)
#define MPI_SHORT SBMPI_DT(SBMPI_K_I16, 2)
#define MPI_UNSIGNED_SHORT SBMPI_DT(SBMPI_K_U16, 2)
#define MPI_INT SBMPI_DT(SBMPI_K_I32, 4)
#define MPI_UNSIGNED SBMPI_DT(SBMPI_K_U32, 4)
#define MPI_FLOAT SBMPI_DT(SBMPI_K_F32, 4)
#define MPI_LONG SBMPI_DT(SBMPI_K_I64, 8)
#define MPI_UNSIGNED_LONG SBMPI_DT(SBMPI_K_U64, 8)
#define MPI_LONG_LONG SBMPI_DT(SBMPI_K_I64, 8)
#define MPI_LONG_LONG_INT SBMPI_DT(SBMPI_K_I64, 8)
#define MPI_UNSIGNED_LONG_LONG SBMPI_DT(SBMPI_K_U64, 8)
#define MPI_DOUBLE SBMPI_DT(SBMPI_K_F64, 8)
#define MPI_LONG_DOUBLE SBMPI_DT(SBMPI_K_F128, 16)
#define MPI_FLOAT_INT SBMPI_DT(SBMPI_K_P_FLOAT_INT, 8)
#define MPI_DOUBLE_INT SBMPI_DT(SBMPI_K_P_DOUBLE_INT, 16)
#define MPI_LONG_INT SBMPI_DT(SBMPI_K_P_LONG_INT, 16)
#define MPI_SHORT_INT SBMPI_DT(SBMPI_K_P_SHORT_INT, 8)
#define MPI_2INT SBMPI_DT(SBMPI_K_P_2INT, 8)
#define MPI_LONG_DOUBLE_INT SBMPI_DT(SBMPI_K_P_LDOUBLE_INT, 32)
#define MPI_C_DOUBLE_COMPLEX SBMPI_DT(SBMPI_K_C128, 16)
#define MPI_DOUBLE_COMPLEX SBMPI_DT(SBMPI_K_C128, 16)

/* ops; handles >= 100 are MPI_Op_create'd user functions */
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_PROD 4
#define MPI_LOR 5
#define MPI_LAND 6
#define MPI_BOR 7
#define MPI_BAND 8
#define MPI_MAXLOC 9
#define MPI_MINLOC 10

#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4
#define MPI_MODE_RDWR 8
#define MPI_MODE_NOPRECEDE 1
#define MPI_MODE_NOSTORE 2
#define MPI_MODE_NOSUCCEED 4
#define MPI_MODE_NOPUT 8

typedef void(MPI_User_function)(void *, void *, int *, MPI_Datatype *);

int MPI_Init(int *argc, char ***argv);
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Initialized(int *flag);
int MPI_Finalize(void);
int MPI_Finalized(int *flag);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);
int MPI_Pcontrol(const int level, ...);
int MPI_Get_processor_name(char *name, int *len);

int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *out);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *out);
int MPI_Comm_create_group(MPI_Comm comm, MPI_Group group, int tag, MPI_Comm *out);
int MPI_Comm_set_errhandler(MPI_Comm comm, MPI_Errhandler eh);
int MPI_Group_incl(MPI_Group group, int n, const int ranks[], MPI_Group *out);
int MPI_Group_free(MPI_Group *group);
int MPI_Attr_get(MPI_Comm comm, int keyval, void *attr, int *flag);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Scan(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Exscan(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, MPI_Comm comm);
int MPI_Allgatherv(const void *s, int sc, MPI_Datatype st, void *r, const int *rc, const int *displs,
                   MPI_Datatype rt, MPI_Comm comm);
int MPI_Gather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Gatherv(const void *s, int sc, MPI_Datatype st, void *r, const int *rc, const int *displs,
                MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Scatterv(const void *s, const int *sc, const int *displs, MPI_Datatype st, void *r, int rc,
                 MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Alltoall(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, MPI_Comm comm);
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, const int *rc,
                  const int *rd, MPI_Datatype rt, MPI_Comm comm);

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Bsend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Issend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Sendrecv(const void *s, int sc, MPI_Datatype st, int dest, int stag, void *r, int rc, MPI_Datatype rt,
                 int src, int rtag, MPI_Comm comm, MPI_Status *status);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
int MPI_Waitall(int n, MPI_Request reqs[], MPI_Status sts[]);
int MPI_Waitany(int n, MPI_Request reqs[], int *index, MPI_Status *st);
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *st);
int MPI_Testall(int n, MPI_Request reqs[], int *flag, MPI_Status sts[]);
int MPI_Probe(int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Iprobe(int src, int tag, MPI_Comm comm, int *flag, MPI_Status *st);
int MPI_Request_free(MPI_Request *req);
int MPI_Cancel(MPI_Request *req);
int MPI_Get_count(const MPI_Status *st, MPI_Datatype dt, int *count);

int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *newt);
int MPI_Type_create_struct(int n, const int bl[], const MPI_Aint disp[], const MPI_Datatype types[],
                           MPI_Datatype *newt);
int MPI_Type_commit(MPI_Datatype *dt);
int MPI_Type_free(MPI_Datatype *dt);
int MPI_Type_size(MPI_Datatype dt, int *size);
int MPI_Op_create(MPI_User_function *fn, int commute, MPI_Op *op);
int MPI_Op_free(MPI_Op *op);

int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr);
int MPI_Free_mem(void *base);
int MPI_Win_create(void *base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win *win);
int MPI_Win_fence(int assert_, MPI_Win win);
int MPI_Win_free(MPI_Win *win);
int MPI_Put(const void *origin, int ocount, MPI_Datatype odt, int target, MPI_Aint tdisp, int tcount,
            MPI_Datatype tdt, MPI_Win win);

int MPI_File_open(MPI_Comm comm, const char *name, int amode, MPI_Info info, MPI_File *fh);
int MPI_File_read_at(MPI_File fh, MPI_Offset off, void *buf, int count, MPI_Datatype dt, MPI_Status *st);
int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void *buf, int count, MPI_Datatype dt, MPI_Status *st);
int MPI_File_close(MPI_File *fh);

#ifdef __cplusplus
}
#endif
#endif

/* Single-rank MPI stand-in used ONLY to compile the reference's CPU classes into
 * oracle/_ref/libthunder_ref.so (test infrastructure; never linked into the product).
 *
 * The container has no MPI.  The reference's hot-path classes (Projector, Reconstructor,
 * Particle, the logDataVSPrior free functions) use MPI only for (a) rank bookkeeping in
 * class Parallel (reference include/Parallel.h) and (b) in-place all-reduces inside
 * Reconstructor::allReduceF/T/O (reference src/Reconstructor.cpp:2350-2520).  With one
 * rank per hemisphere an in-place all-reduce is the identity, which is what this header
 * implements.  Point-to-point calls abort: nothing on the oracle path may reach them.
 */
#ifndef THB_ORACLE_MPI_STUB_H
#define THB_ORACLE_MPI_STUB_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Group;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_IN_PLACE ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_MAX_PROCESSOR_NAME 256

enum { MPI_CHAR = 1, MPI_BYTE, MPI_INT, MPI_LONG, MPI_UNSIGNED, MPI_UNSIGNED_LONG, MPI_FLOAT,
       MPI_DOUBLE, MPI_COMPLEX, MPI_DOUBLE_COMPLEX, MPI_C_BOOL, MPI_LONG_LONG };
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN, MPI_LAND, MPI_LOR };

static inline int thb_mpi_type_size_(MPI_Datatype t)
{
    switch (t) {
        case MPI_CHAR: case MPI_BYTE: case MPI_C_BOOL: return 1;
        case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
        case MPI_LONG: case MPI_UNSIGNED_LONG: case MPI_DOUBLE: case MPI_COMPLEX: case MPI_LONG_LONG: return 8;
        case MPI_DOUBLE_COMPLEX: return 16;
        default: return 0;
    }
}

static inline void thb_mpi_unreachable_(const char* what)
{
    fprintf(stderr, "oracle mpi stub: %s is not available in the single-rank oracle build\n", what);
    abort();
}

static inline int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* n) { (void)c; *n = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_group(MPI_Comm c, MPI_Group* g) { (void)c; *g = 1; return MPI_SUCCESS; }
static inline int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* out)
{ (void)g; (void)n; (void)ranks; *out = 1; return MPI_SUCCESS; }
static inline int MPI_Group_free(MPI_Group* g) { *g = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* out)
{ (void)c; (void)g; *out = MPI_COMM_SELF; return MPI_SUCCESS; }
static inline int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return MPI_SUCCESS; }
static inline int MPI_Type_size(MPI_Datatype t, int* size) { *size = thb_mpi_type_size_(t); return MPI_SUCCESS; }
static inline int MPI_Get_count(const MPI_Status* s, MPI_Datatype t, int* count)
{ (void)t; *count = s ? s->count_ : 0; return MPI_SUCCESS; }
static inline int MPI_Get_processor_name(char* name, int* len)
{ strcpy(name, "oracle"); *len = 6; return MPI_SUCCESS; }

static inline int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
    (void)op; (void)c;
    if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf)
        memcpy(recvbuf, sendbuf, (size_t)count * (size_t)thb_mpi_type_size_(t));
    return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void)root; return MPI_Allreduce(sendbuf, recvbuf, count, t, op, c); }
static inline int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm c)
{ (void)buf; (void)count; (void)t; (void)root; (void)c; return MPI_SUCCESS; }
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; thb_mpi_unreachable_("MPI_Send"); return 1; }
static inline int MPI_Ssend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; thb_mpi_unreachable_("MPI_Ssend"); return 1; }
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; thb_mpi_unreachable_("MPI_Recv"); return 1; }

#ifdef __cplusplus
}
#endif

#endif /* THB_ORACLE_MPI_STUB_H */

// TEST INFRASTRUCTURE (oracle/_ref build of the reference's geometry::Domain only).
// A thread-backed stand-in for the handful of MPI calls that the reference's net:: layer, octree::DistributedStore
// (one-sided windows) and geometry::Domain make: every "rank" is a thread of one process, communicators are
// small shared objects, collectives meet at a barrier, windows are plain memory that peers memcpy from/to.
// Only what that path executes is implemented (fake_mpi.cc); the rest of the names exist so that the
// unmodified reference sources compile, and abort if reached.
#pragma once
#include <cstddef>
#include <cstdint>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Win;
typedef int MPI_Info;
typedef int MPI_Group;
typedef int MPI_Request;
typedef int MPI_Errhandler;
typedef int MPI_File;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef long long MPI_Count;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };

#define MPI_SUCCESS 0
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_WIN_NULL 0
#define MPI_INFO_NULL 0
#define MPI_GROUP_NULL 0
#define MPI_GROUP_EMPTY (-1)
#define MPI_REQUEST_NULL 0
#define MPI_DATATYPE_NULL 0
#define MPI_ERRORS_RETURN 1
#define MPI_ERRORS_ARE_FATAL 2
#define MPI_STATUS_IGNORE ((MPI_Status*)nullptr)
#define MPI_STATUSES_IGNORE ((MPI_Status*)nullptr)
#define MPI_IN_PLACE ((void*)-1)
#define MPI_BOTTOM ((void*)0)
#define MPI_UNWEIGHTED ((int*)nullptr)
#define MPI_MAX_ERROR_STRING 256
#define MPI_IDENT 0
#define MPI_CONGRUENT 1
#define MPI_UNEQUAL 3
#define MPI_COMM_TYPE_SHARED 1
#define MPI_LOCK_SHARED 1
#define MPI_LOCK_EXCLUSIVE 2
#define MPI_MODE_NOCHECK 1
#define MPI_MODE_NOPRECEDE 2
#define MPI_MODE_NOSUCCEED 4
#define MPI_MODE_RDONLY 2
#define MPI_MODE_WRONLY 4
#define MPI_MODE_CREATE 1
#define MPI_MODE_EXCL 64
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_UNDEFINED (-32766)
#define MPI_THREAD_SINGLE 0

// datatypes: kind << 24 | bytes for the predefined ones (kind 1 signed, 2 unsigned, 3 floating, 4 opaque);
// handles >= 1 << 28 index the table of derived types
#define FAKEMPI_DT(kind, bytes) (((kind) << 24) | (bytes))
#define MPI_CHAR FAKEMPI_DT(1, 1)
#define MPI_SIGNED_CHAR FAKEMPI_DT(1, 1)
#define MPI_SHORT FAKEMPI_DT(1, 2)
#define MPI_INT FAKEMPI_DT(1, 4)
#define MPI_LONG FAKEMPI_DT(1, 8)
#define MPI_LONG_LONG FAKEMPI_DT(1, 8)
#define MPI_INT8_T FAKEMPI_DT(1, 1)
#define MPI_INT16_T FAKEMPI_DT(1, 2)
#define MPI_INT32_T FAKEMPI_DT(1, 4)
#define MPI_INT64_T FAKEMPI_DT(1, 8)
#define MPI_UNSIGNED_CHAR FAKEMPI_DT(2, 1)
#define MPI_UNSIGNED_SHORT FAKEMPI_DT(2, 2)
#define MPI_UNSIGNED FAKEMPI_DT(2, 4)
#define MPI_UNSIGNED_LONG FAKEMPI_DT(2, 8)
#define MPI_UNSIGNED_LONG_LONG FAKEMPI_DT(2, 8)
#define MPI_UINT8_T FAKEMPI_DT(2, 1)
#define MPI_UINT16_T FAKEMPI_DT(2, 2)
#define MPI_UINT32_T FAKEMPI_DT(2, 4)
#define MPI_UINT64_T FAKEMPI_DT(2, 8)
#define MPI_FLOAT FAKEMPI_DT(3, 4)
#define MPI_DOUBLE FAKEMPI_DT(3, 8)
#define MPI_BYTE FAKEMPI_DT(4, 1)
#define MPI_AINT FAKEMPI_DT(1, 8)

#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3
#define MPI_LOR 4
#define MPI_LAND 5
#define MPI_BOR 6
#define MPI_BAND 7
#define MPI_PROD 8

extern "C" {
// ---- implemented (fake_mpi.cc)
int MPI_Initialized(int* flag);
int MPI_Finalized(int* flag);
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Comm_dup(MPI_Comm, MPI_Comm*);
int MPI_Comm_free(MPI_Comm*);
int MPI_Comm_compare(MPI_Comm, MPI_Comm, int*);
int MPI_Comm_split(MPI_Comm, int color, int key, MPI_Comm*);
int MPI_Comm_split_type(MPI_Comm, int type, int key, MPI_Info, MPI_Comm*);
int MPI_Comm_set_errhandler(MPI_Comm, MPI_Errhandler);
int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void*, int, MPI_Datatype, int root, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int root, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);
int MPI_Gather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int root, MPI_Comm);
int MPI_Gatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, int root, MPI_Comm);
int MPI_Alltoall(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Send(const void*, int, MPI_Datatype, int dest, int tag, MPI_Comm);
int MPI_Ssend(const void*, int, MPI_Datatype, int dest, int tag, MPI_Comm);
int MPI_Recv(void*, int, MPI_Datatype, int src, int tag, MPI_Comm, MPI_Status*);
int MPI_Isend(const void*, int, MPI_Datatype, int dest, int tag, MPI_Comm, MPI_Request*);
int MPI_Irecv(void*, int, MPI_Datatype, int src, int tag, MPI_Comm, MPI_Request*);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Waitall(int, MPI_Request*, MPI_Status*);
int MPI_Type_create_struct(int, const int*, const MPI_Aint*, const MPI_Datatype*, MPI_Datatype*);
int MPI_Type_commit(MPI_Datatype*);
int MPI_Type_free(MPI_Datatype*);
int MPI_Type_size(MPI_Datatype, int*);
int MPI_Get_address(const void*, MPI_Aint*);
int MPI_Info_create(MPI_Info*);
int MPI_Info_set(MPI_Info, const char*, const char*);
int MPI_Info_free(MPI_Info*);
int MPI_Win_allocate(MPI_Aint bytes, int disp_unit, MPI_Info, MPI_Comm, void* baseptr, MPI_Win*);
int MPI_Win_free(MPI_Win*);
int MPI_Win_fence(int, MPI_Win);
int MPI_Win_lock(int, int, int, MPI_Win);
int MPI_Win_unlock(int, MPI_Win);
int MPI_Win_set_errhandler(MPI_Win, MPI_Errhandler);
int MPI_Get(void*, int, MPI_Datatype, int rank, MPI_Aint disp, int, MPI_Datatype, MPI_Win);
int MPI_Put(const void*, int, MPI_Datatype, int rank, MPI_Aint disp, int, MPI_Datatype, MPI_Win);
int MPI_Error_string(int, char*, int*);
int MPI_Abort(MPI_Comm, int);
double MPI_Wtime(void);
// ---- named by the sources, never reached on this path (abort)
int MPI_Scan(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Scatter(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Comm_group(MPI_Comm, MPI_Group*);
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm*);
int MPI_Group_free(MPI_Group*);
int MPI_Group_rank(MPI_Group, int*);
int MPI_Group_size(MPI_Group, int*);
int MPI_Group_incl(MPI_Group, int, const int*, MPI_Group*);
int MPI_Group_excl(MPI_Group, int, const int*, MPI_Group*);
int MPI_Group_translate_ranks(MPI_Group, int, const int*, MPI_Group, int*);
int MPI_Dist_graph_create_adjacent(MPI_Comm, int, const int*, const int*, int, const int*, const int*, MPI_Info, int, MPI_Comm*);
int MPI_Dist_graph_neighbors_count(MPI_Comm, int*, int*, int*);
int MPI_Dist_graph_neighbors(MPI_Comm, int, int*, int*, int, int*, int*);
int MPI_Neighbor_allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Neighbor_allgatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);
int MPI_Ineighbor_alltoall(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm, MPI_Request*);
int MPI_Ineighbor_alltoallv(const void*, const int*, const int*, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm, MPI_Request*);
// ---- the harness: a rank is a thread
void fakempi_run(int nranks, void (*body)(int rank, void* arg), void* arg);
}

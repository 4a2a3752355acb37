/* oracle/stub/mpi/mpi.h -- TEST INFRASTRUCTURE ONLY (oracle): a one-rank MPI, just enough to compile and run the
 * reference's vendored Zoltan (src/zoltan, version 3.901) inside oracle/_ref so that the partitions the host mirror's
 * own RCB produces can be compared with Zoltan's (src/Partition/ZoltanGeom.cpp:139-244). There is no MPI in this
 * image. Every collective is the identity on one rank; point-to-point calls between different ranks cannot occur
 * and abort. Not part of the product. */
#ifndef ORACLE_STUB_MPI_H
#define ORACLE_STUB_MPI_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
#define MPI_VERSION 3
#define MPI_SUBVERSION 1
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Group;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count; } MPI_Status;
typedef void (MPI_User_function)( void*, void*, int*, MPI_Datatype* );
#define MPI_SUCCESS 0
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_REQUEST_NULL 0
#define MPI_IN_PLACE ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_MAX_ERROR_STRING 256
#define MPI_MAX_PROCESSOR_NAME 256
/* predefined datatypes: handle = size in bytes + 1000*kind (sizes recovered by MPI_Type_size) */
#define MPI_CHAR 1001
#define MPI_BYTE 2001
#define MPI_SHORT 3002
#define MPI_INT 4004
#define MPI_UNSIGNED 5004
#define MPI_LONG 6008
#define MPI_UNSIGNED_LONG 7008
#define MPI_LONG_LONG 8008
#define MPI_LONG_LONG_INT 8008
#define MPI_UNSIGNED_LONG_LONG 9008
#define MPI_FLOAT 10004
#define MPI_DOUBLE 11008
#define MPI_2INT 12008
#define MPI_FLOAT_INT 13008
#define MPI_DOUBLE_INT 14016
#define MPI_UB 15000
#define MPI_LB 16000
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LOR 4
#define MPI_BOR 5
#define MPI_MAXLOC 6
#define MPI_MINLOC 7
#define MPI_LAND 8
#define MPI_BAND 9
int MPI_Init( int*, char*** );
int MPI_Initialized( int* );
int MPI_Finalize( void );
int MPI_Abort( MPI_Comm, int );
int MPI_Comm_rank( MPI_Comm, int* );
int MPI_Comm_size( MPI_Comm, int* );
int MPI_Comm_dup( MPI_Comm, MPI_Comm* );
int MPI_Comm_split( MPI_Comm, int, int, MPI_Comm* );
int MPI_Comm_free( MPI_Comm* );
int MPI_Comm_group( MPI_Comm, MPI_Group* );
int MPI_Comm_create( MPI_Comm, MPI_Group, MPI_Comm* );
int MPI_Group_incl( MPI_Group, int, const int*, MPI_Group* );
int MPI_Group_excl( MPI_Group, int, const int*, MPI_Group* );
int MPI_Group_free( MPI_Group* );
int MPI_Barrier( MPI_Comm );
int MPI_Bcast( void*, int, MPI_Datatype, int, MPI_Comm );
int MPI_Allreduce( const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm );
int MPI_Reduce( const void*, void*, int, MPI_Datatype, MPI_Op, int, MPI_Comm );
int MPI_Scan( const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm );
int MPI_Reduce_scatter( const void*, void*, const int*, MPI_Datatype, MPI_Op, MPI_Comm );
int MPI_Allgather( const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm );
int MPI_Allgatherv( const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm );
int MPI_Gather( const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm );
int MPI_Gatherv( const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, int, MPI_Comm );
int MPI_Scatter( const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm );
int MPI_Alltoall( const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm );
int MPI_Alltoallv( const void*, const int*, const int*, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm );
int MPI_Send( const void*, int, MPI_Datatype, int, int, MPI_Comm );
int MPI_Rsend( const void*, int, MPI_Datatype, int, int, MPI_Comm );
int MPI_Isend( const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request* );
int MPI_Recv( void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status* );
int MPI_Irecv( void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request* );
int MPI_Sendrecv( const void*, int, MPI_Datatype, int, int, void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status* );
int MPI_Wait( MPI_Request*, MPI_Status* );
int MPI_Waitall( int, MPI_Request*, MPI_Status* );
int MPI_Waitany( int, MPI_Request*, int*, MPI_Status* );
int MPI_Waitsome( int, MPI_Request*, int*, int*, MPI_Status* );
int MPI_Op_create( MPI_User_function*, int, MPI_Op* );
int MPI_Op_free( MPI_Op* );
int MPI_Type_size( MPI_Datatype, int* );
int MPI_Type_contiguous( int, MPI_Datatype, MPI_Datatype* );
int MPI_Type_struct( int, int*, MPI_Aint*, MPI_Datatype*, MPI_Datatype* );
int MPI_Type_create_struct( int, const int*, const MPI_Aint*, const MPI_Datatype*, MPI_Datatype* );
int MPI_Type_create_resized( MPI_Datatype, MPI_Aint, MPI_Aint, MPI_Datatype* );
int MPI_Type_commit( MPI_Datatype* );
int MPI_Type_free( MPI_Datatype* );
int MPI_Address( void*, MPI_Aint* );
int MPI_Get_address( const void*, MPI_Aint* );
int MPI_Get_processor_name( char*, int* );
int MPI_Error_string( int, char*, int* );
double MPI_Wtime( void );
double MPI_Wtick( void );
#ifdef __cplusplus
}
#endif
#endif

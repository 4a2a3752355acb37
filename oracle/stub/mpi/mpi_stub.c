/* oracle/stub/mpi/mpi_stub.c -- TEST INFRASTRUCTURE ONLY: the one-rank MPI behind oracle/stub/mpi/mpi.h. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "mpi.h"

/* derived datatypes: handle 100000 + index into a size table (extent = size: the structs Zoltan describes are
 * sent as whole C structs, MPI_UB gives their extent) */
static int g_tsize[4096]; static int g_nt = 0;
static int tsize( MPI_Datatype t ) {
  if (t >= 100000) return g_tsize[t-100000];
  if (t == MPI_UB || t == MPI_LB) return 0;
  return t % 1000;
}
static int newtype( int size ) { if (g_nt >= 4096) { fprintf( stderr, "mpi stub: too many datatypes\n" ); abort(); } g_tsize[g_nt] = size; return 100000 + g_nt++; }
static void nope( const char* what ) { fprintf( stderr, "mpi stub: %s between different ranks cannot happen on one rank\n", what ); abort(); }
static void copy( const void* s, void* r, int n, MPI_Datatype t ) { if (s != MPI_IN_PLACE && s != r && n > 0) memmove( r, s, (size_t)n*(size_t)tsize(t) ); }

int MPI_Init( int* a, char*** b ) { (void)a; (void)b; return 0; }
int MPI_Initialized( int* f ) { *f = 1; return 0; }
int MPI_Finalize( void ) { return 0; }
int MPI_Abort( MPI_Comm c, int e ) { (void)c; fprintf( stderr, "MPI_Abort(%d)\n", e ); abort(); }
int MPI_Comm_rank( MPI_Comm c, int* r ) { (void)c; *r = 0; return 0; }
int MPI_Comm_size( MPI_Comm c, int* s ) { (void)c; *s = 1; return 0; }
int MPI_Comm_dup( MPI_Comm c, MPI_Comm* n ) { *n = c; return 0; }
int MPI_Comm_split( MPI_Comm c, int color, int key, MPI_Comm* n ) { (void)key; *n = color == MPI_UNDEFINED ? MPI_COMM_NULL : c; return 0; }
int MPI_Comm_free( MPI_Comm* c ) { *c = MPI_COMM_NULL; return 0; }
int MPI_Comm_group( MPI_Comm c, MPI_Group* g ) { (void)c; *g = 1; return 0; }
int MPI_Comm_create( MPI_Comm c, MPI_Group g, MPI_Comm* n ) { *n = g ? c : MPI_COMM_NULL; return 0; }
int MPI_Group_incl( MPI_Group g, int n, const int* r, MPI_Group* o ) { (void)g; (void)r; *o = n > 0; return 0; }
int MPI_Group_excl( MPI_Group g, int n, const int* r, MPI_Group* o ) { (void)g; (void)r; *o = n == 0; return 0; }
int MPI_Group_free( MPI_Group* g ) { *g = 0; return 0; }
int MPI_Barrier( MPI_Comm c ) { (void)c; return 0; }
int MPI_Bcast( void* b, int n, MPI_Datatype t, int root, MPI_Comm c ) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
int MPI_Allreduce( const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c ) { (void)o; (void)c; copy( s, r, n, t ); return 0; }
int MPI_Reduce( const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c ) { (void)o; (void)root; (void)c; copy( s, r, n, t ); return 0; }
int MPI_Scan( const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c ) { (void)o; (void)c; copy( s, r, n, t ); return 0; }
int MPI_Reduce_scatter( const void* s, void* r, const int* cnt, MPI_Datatype t, MPI_Op o, MPI_Comm c ) { (void)o; (void)c; copy( s, r, cnt[0], t ); return 0; }
int MPI_Allgather( const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, MPI_Comm c ) { (void)rn; (void)rt; (void)c; copy( s, r, n, t ); return 0; }
int MPI_Allgatherv( const void* s, int n, MPI_Datatype t, void* r, const int* rn, const int* d, MPI_Datatype rt, MPI_Comm c ) {
  (void)rn; (void)c; if (s != MPI_IN_PLACE) memmove( (char*)r + (size_t)d[0]*(size_t)tsize(rt), s, (size_t)n*(size_t)tsize(t) ); return 0; }
int MPI_Gather( const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c ) { (void)rn; (void)rt; (void)root; (void)c; copy( s, r, n, t ); return 0; }
int MPI_Gatherv( const void* s, int n, MPI_Datatype t, void* r, const int* rn, const int* d, MPI_Datatype rt, int root, MPI_Comm c ) {
  (void)rn; (void)root; (void)c; if (s != MPI_IN_PLACE) memmove( (char*)r + (size_t)d[0]*(size_t)tsize(rt), s, (size_t)n*(size_t)tsize(t) ); return 0; }
int MPI_Scatter( const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c ) { (void)rn; (void)rt; (void)root; (void)c; copy( s, r, n, t ); return 0; }
int MPI_Alltoall( const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt, MPI_Comm c ) { (void)rn; (void)rt; (void)c; copy( s, r, n, t ); return 0; }
int MPI_Alltoallv( const void* s, const int* sn, const int* sd, MPI_Datatype t, void* r, const int* rn, const int* rd, MPI_Datatype rt, MPI_Comm c ) {
  (void)rn; (void)c; memmove( (char*)r + (size_t)rd[0]*(size_t)tsize(rt), (const char*)s + (size_t)sd[0]*(size_t)tsize(t), (size_t)sn[0]*(size_t)tsize(t) ); return 0; }
/* point to point on one rank = messages to oneself (Zoltan's communication plans send their "self message"
 * sizes through MPI, Utilities/Communication/comm_invert_map.c:170-181): posted receives and a queue of messages
 * sent before their receive was posted, matched by tag in order */
typedef struct { void* buf; size_t bytes; int tag, done, used; } Posted;
typedef struct { void* data; size_t bytes; int tag, used; } Queued;
static Posted g_post[65536]; static int g_npost = 0;
static Queued g_queue[65536]; static int g_nq = 0;
static void deliver( Posted* p, const void* data, size_t bytes ) {
  if (bytes > p->bytes) { fprintf( stderr, "mpi stub: message longer than the receive buffer\n" ); abort(); }
  if (bytes) memcpy( p->buf, data, bytes );
  p->done = 1;
}
static int self_send( const void* b, int n, MPI_Datatype t, int d, int tag ) {
  if (d != 0) nope( "a send" );
  size_t bytes = (size_t)n*(size_t)tsize(t);
  for (int i=0; i<g_npost; ++i)
    if (g_post[i].used && !g_post[i].done && (g_post[i].tag == tag || g_post[i].tag == MPI_ANY_TAG)) { g_post[i].tag = tag; deliver( &g_post[i], b, bytes ); return 0; }
  if (g_nq >= 65536) { fprintf( stderr, "mpi stub: message queue full\n" ); abort(); }
  Queued* q = &g_queue[g_nq++];
  q->data = malloc( bytes ? bytes : 1 ); if (bytes) memcpy( q->data, b, bytes ); q->bytes = bytes; q->tag = tag; q->used = 1;
  return 0;
}
static int post_recv( void* b, int n, MPI_Datatype t, int src, int tag ) {
  if (src != 0 && src != MPI_ANY_SOURCE) nope( "a receive" );
  int slot = -1;
  for (int i=0; i<g_npost; ++i) if (!g_post[i].used) { slot = i; break; }
  if (slot < 0) { if (g_npost >= 65536) { fprintf( stderr, "mpi stub: too many posted receives\n" ); abort(); } slot = g_npost++; }
  Posted* p = &g_post[slot];
  p->buf = b; p->bytes = (size_t)n*(size_t)tsize(t); p->tag = tag; p->done = 0; p->used = 1;
  for (int i=0; i<g_nq; ++i)
    if (g_queue[i].used && (tag == MPI_ANY_TAG || g_queue[i].tag == tag)) {
      p->tag = g_queue[i].tag; deliver( p, g_queue[i].data, g_queue[i].bytes ); free( g_queue[i].data ); g_queue[i].used = 0; break; }
  while (g_nq > 0 && !g_queue[g_nq-1].used) --g_nq;
  return slot;
}
static void finish( int slot, MPI_Status* st ) {
  Posted* p = &g_post[slot];
  if (!p->used || !p->done) { fprintf( stderr, "mpi stub: waiting for a message nobody sent (one rank)\n" ); abort(); }
  if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = p->tag; st->MPI_ERROR = 0; st->count = (int)p->bytes; }
  p->used = 0;
}
int MPI_Send( const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c ) { (void)c; return self_send( b, n, t, d, tag ); }
int MPI_Rsend( const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c ) { (void)c; return self_send( b, n, t, d, tag ); }
int MPI_Isend( const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* q ) { (void)c; *q = MPI_REQUEST_NULL; return self_send( b, n, t, d, tag ); }
int MPI_Recv( void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st ) { (void)c; finish( post_recv( b, n, t, s, tag ), st ); return 0; }
int MPI_Irecv( void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request* q ) { (void)c; *q = post_recv( b, n, t, s, tag ) + 1; return 0; }
int MPI_Sendrecv( const void* sb, int sn, MPI_Datatype st, int d, int stag, void* rb, int rn, MPI_Datatype rt, int s, int rtag, MPI_Comm c, MPI_Status* status ) {
  (void)c; int slot = post_recv( rb, rn, rt, s, rtag ); self_send( sb, sn, st, d, stag ); finish( slot, status ); return 0; }
int MPI_Wait( MPI_Request* q, MPI_Status* s ) { if (*q != MPI_REQUEST_NULL) finish( *q - 1, s ); *q = MPI_REQUEST_NULL; return 0; }
int MPI_Waitall( int n, MPI_Request* q, MPI_Status* s ) { for (int i=0; i<n; ++i) MPI_Wait( q+i, s ? s+i : 0 ); return 0; }
int MPI_Waitany( int n, MPI_Request* q, int* idx, MPI_Status* s ) {
  for (int i=0; i<n; ++i) if (q[i] != MPI_REQUEST_NULL && g_post[q[i]-1].done) { *idx = i; return MPI_Wait( q+i, s ); }
  for (int i=0; i<n; ++i) if (q[i] != MPI_REQUEST_NULL) { *idx = i; return MPI_Wait( q+i, s ); }      /* aborts: nobody will send */
  *idx = MPI_UNDEFINED; return 0; }
int MPI_Waitsome( int n, MPI_Request* q, int* out, int* idx, MPI_Status* s ) {
  int k = 0;
  for (int i=0; i<n; ++i) if (q[i] != MPI_REQUEST_NULL && g_post[q[i]-1].done) { MPI_Wait( q+i, s ? s+k : 0 ); idx[k++] = i; }
  *out = k ? k : MPI_UNDEFINED; return 0; }
int MPI_Op_create( MPI_User_function* f, int commute, MPI_Op* o ) { (void)f; (void)commute; *o = 100; return 0; }
int MPI_Op_free( MPI_Op* o ) { *o = 0; return 0; }
int MPI_Type_size( MPI_Datatype t, int* s ) { *s = tsize( t ); return 0; }
int MPI_Type_contiguous( int n, MPI_Datatype t, MPI_Datatype* o ) { *o = newtype( n*tsize(t) ); return 0; }
int MPI_Type_create_struct( int n, const int* len, const MPI_Aint* disp, const MPI_Datatype* types, MPI_Datatype* o ) {
  long ext = 0;          /* extent: the MPI_UB marker if present, else the end of the last member */
  for (int i=0; i<n; ++i) { long e = (long)disp[i] + (long)len[i]*tsize( types[i] ); if (types[i] == MPI_UB) { ext = (long)disp[i]; break; } if (e > ext) ext = e; }
  *o = newtype( (int)ext ); return 0; }
int MPI_Type_struct( int n, int* len, MPI_Aint* disp, MPI_Datatype* types, MPI_Datatype* o ) { return MPI_Type_create_struct( n, len, disp, types, o ); }
int MPI_Type_create_resized( MPI_Datatype t, MPI_Aint lb, MPI_Aint extent, MPI_Datatype* o ) { (void)t; (void)lb; *o = newtype( (int)extent ); return 0; }
int MPI_Type_commit( MPI_Datatype* t ) { (void)t; return 0; }
int MPI_Type_free( MPI_Datatype* t ) { *t = 0; return 0; }
int MPI_Address( void* p, MPI_Aint* a ) { *a = (MPI_Aint)p; return 0; }
int MPI_Get_address( const void* p, MPI_Aint* a ) { *a = (MPI_Aint)p; return 0; }
int MPI_Get_processor_name( char* n, int* l ) { strcpy( n, "oracle" ); *l = 6; return 0; }
int MPI_Error_string( int e, char* s, int* l ) { *l = snprintf( s, MPI_MAX_ERROR_STRING, "mpi stub error %d", e ); return 0; }
double MPI_Wtime( void ) { struct timespec t; clock_gettime( CLOCK_MONOTONIC, &t ); return (double)t.tv_sec + 1e-9*(double)t.tv_nsec; }
double MPI_Wtick( void ) { return 1e-9; }

/* oracle/zoltan_geom.c -- TEST INFRASTRUCTURE ONLY (oracle/_ref): the reference's call into Zoltan for its
 * geometric partitioners, inciter::geomPartMesh (src/Partition/ZoltanGeom.cpp:139-244), restated in C around the
 * reference's own vendored Zoltan 3.901 (src/zoltan, compiled where it lies by oracle/Makefile against the
 * one-rank MPI of oracle/stub/mpi): the same parameters (LB_APPROACH PARTITION, OBJ_WEIGHT_DIM 0, RETURN_LISTS
 * PART, AVERAGE_CUTS 1, NUM_GLOBAL_PARTS) and the same four query functions over the element centroids. Used by
 * tests/test_oracle_zoltan.py to pin the host mirror's own RCB (xyst_b200/host/mesh.cpp) to Zoltan's. */
#include <stdio.h>
#include <stdlib.h>
#include "zoltan.h"

typedef struct { int n; const double *x, *y, *z; } Cent;

static int num_obj( void* data, int* ierr ) { *ierr = ZOLTAN_OK; return ((Cent*)data)->n; }                 /* :53-58 */
static void obj_list( void* data, int sg, int sl, ZOLTAN_ID_PTR gid, ZOLTAN_ID_PTR lid, int wd, float* w, int* ierr ) {  /* :60-72 */
  (void)sg; (void)sl; (void)wd; (void)w;
  Cent* c = (Cent*)data; *ierr = ZOLTAN_OK;
  for (int i=0; i<c->n; ++i) { gid[i] = (ZOLTAN_ID_TYPE)i; lid[i] = (ZOLTAN_ID_TYPE)i; }
}
static int num_geom( void* data, int* ierr ) { (void)data; *ierr = ZOLTAN_OK; return 3; }                      /* :74-78 */
static void geom_list( void* data, int sg, int sl, int n, ZOLTAN_ID_PTR gid, ZOLTAN_ID_PTR lid, int nd, double* g, int* ierr ) {  /* :80-99 */
  (void)gid; (void)lid;
  Cent* c = (Cent*)data;
  if (sg != 1 || sl != 1 || nd != 3) { *ierr = ZOLTAN_FATAL; return; }
  *ierr = ZOLTAN_OK;
  for (int i=0; i<n; ++i) { g[3*i] = c->x[i]; g[3*i+1] = c->y[i]; g[3*i+2] = c->z[i]; }
}

/* element centroids (already computed by the caller as ZoltanGeom.cpp:101-138 does) -> part of every element */
int orc_zoltan_geom( const char* alg, int nelem, const double* cx, const double* cy, const double* cz, int npart, int* part )
{
  float ver;
  Cent c = { nelem, cx, cy, cz };
  int changes, ng, nl, nimp, nexp, *impp, *imptp, *expp, *exptp;
  ZOLTAN_ID_PTR impg, impl, expg, expl;
  char np[32];
  if (Zoltan_Initialize( 0, NULL, &ver ) != ZOLTAN_OK) return 1;
  struct Zoltan_Struct* zz = Zoltan_Create( MPI_COMM_WORLD );
  if (!zz) return 2;
  snprintf( np, sizeof np, "%d", npart );
  Zoltan_Set_Param( zz, "DEBUG_LEVEL", "0" );
  Zoltan_Set_Param( zz, "LB_METHOD", alg );
  Zoltan_Set_Param( zz, "LB_APPROACH", "PARTITION" );
  Zoltan_Set_Param( zz, "NUM_GID_ENTRIES", "1" );
  Zoltan_Set_Param( zz, "NUM_LID_ENTRIES", "1" );
  Zoltan_Set_Param( zz, "OBJ_WEIGHT_DIM", "0" );
  Zoltan_Set_Param( zz, "RETURN_LISTS", "PART" );
  Zoltan_Set_Param( zz, "RCB_OUTPUT_LEVEL", "0" );
  Zoltan_Set_Param( zz, "AVERAGE_CUTS", "1" );
  Zoltan_Set_Param( zz, "NUM_GLOBAL_PARTS", np );
  Zoltan_Set_Num_Obj_Fn( zz, num_obj, &c );
  Zoltan_Set_Obj_List_Fn( zz, obj_list, &c );
  Zoltan_Set_Num_Geom_Fn( zz, num_geom, &c );
  Zoltan_Set_Geom_Multi_Fn( zz, geom_list, &c );
  int rc = Zoltan_LB_Partition( zz, &changes, &ng, &nl, &nimp, &impg, &impl, &impp, &imptp, &nexp, &expg, &expl, &expp, &exptp );
  if (rc != ZOLTAN_OK || nexp != nelem) { Zoltan_Destroy( &zz ); return 3; }
  for (int p=0; p<nexp; ++p) part[ expl[p] ] = exptp[p];                                                     /* :227-229 */
  Zoltan_LB_Free_Part( &impg, &impl, &impp, &imptp );
  Zoltan_LB_Free_Part( &expg, &expl, &expp, &exptp );
  Zoltan_Destroy( &zz );
  return 0;
}

// xyst_b200.cu -- B200 (sm_100a) implementation of the RieCG hot path behind the
// C ABI of include/xyst_b200.h. No CPU fallback: every compute entry needs a device.
//
// Design (DESIGN.md has the full account):
//  * The reference walks superedges and scatter-adds into nodes
//    (src/Physics/Riemann.cpp:266-326, :696-759). Here the accumulation is turned
//    around into a deterministic NODE GATHER: superedges are flattened to unique
//    edges, each node owns a sliced-ELL (32 nodes per slice = one warp) incidence
//    list, and one thread sums its node's contributions in a fixed order. No atomics,
//    no zero-fill, bit-reproducible from run to run.
//  * Everything nodal is structure-of-arrays (U, primitives W, coordinates X,
//    gradients G: one array per component, stride NP), and EDGES live in slots ordered
//    like their owner nodes: the j-th edge owned by node p (its lower endpoint) sits at
//    slice_base(p/32) + j*32 + p%32. A warp of the flux kernel therefore works on 32
//    consecutive owner nodes, and the k-th neighbour of consecutive nodes is, on a
//    locality-ordered mesh, a run of consecutive nodes too -- endpoint loads, normal
//    loads and flux stores coalesce into a few 128-byte lines per request instead of
//    one line per lane (the first, array-of-structures version of these kernels was
//    bound by L1 wavefronts, see profiles/).
//  * Gradients: one gather kernel reads neighbour primitives and the edge normals,
//    adds the boundary-face part and divides by the nodal volume (fuses riemann::grad
//    + RieCG::rhs :936-939).
//  * Fluxes: one thread per edge slot does MUSCL + Rusanov/HLLC once and stores 5
//    doubles; a second gather kernel sums them per node, adds boundary and source terms
//    and applies the Runge-Kutta update in the same pass (fuses advdom/advbnd/src +
//    RieCG::solve :1011-1021), also refreshing the primitive variables.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <array>
#include <map>
#include <algorithm>
#include <numeric>
#include <stdexcept>
#include "xyst_b200.h"

namespace {

thread_local std::string g_err;

int fail( const std::string& m ) { g_err = m; return 1; }

#define CK( call ) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  throw std::runtime_error( std::string( #call ) + ": " + cudaGetErrorString( e_ ) ); } while (0)

#define API_BEGIN try {
#define API_END } catch (std::exception& e) { return fail( e.what() ); } return 0;

// tuning knobs (compile-time; defaults chosen from the ncu measurements under profiles/)
#ifndef FLUX_THREADS
#define FLUX_THREADS 128
#endif
#ifndef FLUX_MINB
#define FLUX_MINB 6
#endif
#ifndef NODE_THREADS
#define NODE_THREADS 256
#endif
#ifndef GRAD_MINB
#define GRAD_MINB 4
#endif
#ifndef RHS_MINB
#define RHS_MINB 4
#endif
#ifndef GRAD_UNROLL
#define GRAD_UNROLL 14     // interior nodes of a Kuhn-split box have 14 edges: one full batch
#endif
#ifndef ZAL_UNROLL
#define ZAL_UNROLL 2
#endif
#ifndef RHS_UNROLL
#define RHS_UNROLL 14
#endif

constexpr int kGradUnroll = GRAD_UNROLL, kRhsUnroll = RHS_UNROLL, kZalUnroll = ZAL_UNROLL;
constexpr int NC = 5;          // flow components handled by the kernels

// ---------------------------------------------------------------------------------
// NCCL through dlopen: the process normally has torch's libnccl.so.2 loaded already
// ---------------------------------------------------------------------------------
struct Nccl {
  void* h = nullptr;
  typedef struct { char internal[128]; } UniqueId;
  int (*GetUniqueId)( UniqueId* ) = nullptr;
  int (*CommInitRank)( void**, int, UniqueId, int ) = nullptr;
  int (*CommDestroy)( void* ) = nullptr;
  int (*Send)( const void*, size_t, int, int, void*, cudaStream_t ) = nullptr;
  int (*Recv)( void*, size_t, int, int, void*, cudaStream_t ) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)( const void*, void*, size_t, int, int, void*, cudaStream_t ) = nullptr;
  const char* (*GetErrorString)( int ) = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (auto n : names) { h = dlopen( n, RTLD_NOW | RTLD_GLOBAL ); if (h) break; }
    if (!h) return false;
    auto S = [&]( const char* s ){ return dlsym( h, s ); };
    GetUniqueId = reinterpret_cast< decltype(GetUniqueId) >( S( "ncclGetUniqueId" ) );
    CommInitRank = reinterpret_cast< decltype(CommInitRank) >( S( "ncclCommInitRank" ) );
    CommDestroy = reinterpret_cast< decltype(CommDestroy) >( S( "ncclCommDestroy" ) );
    Send = reinterpret_cast< decltype(Send) >( S( "ncclSend" ) );
    Recv = reinterpret_cast< decltype(Recv) >( S( "ncclRecv" ) );
    GroupStart = reinterpret_cast< decltype(GroupStart) >( S( "ncclGroupStart" ) );
    GroupEnd = reinterpret_cast< decltype(GroupEnd) >( S( "ncclGroupEnd" ) );
    AllReduce = reinterpret_cast< decltype(AllReduce) >( S( "ncclAllReduce" ) );
    GetErrorString = reinterpret_cast< decltype(GetErrorString) >( S( "ncclGetErrorString" ) );
    return GetUniqueId && CommInitRank && Send && Recv && GroupStart && GroupEnd && AllReduce;
  }
};
Nccl g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MIN = 3;   // nccl.h enums (stable ABI)
#define NK( call ) do { int r_ = (call); if (r_ != 0) throw std::runtime_error( \
  std::string( #call ) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString( r_ ) : "nccl error") ); } while (0)

template< class T > struct DevBuf {
  T* p = nullptr; size_t n = 0;
  void alloc( size_t m ) { release(); n = m; if (m) CK( cudaMalloc( &p, m*sizeof(T) ) ); }
  void upload( const std::vector< T >& h, cudaStream_t s ) {
    alloc( h.size() );
    if (!h.empty()) { CK( cudaMemcpyAsync( p, h.data(), h.size()*sizeof(T), cudaMemcpyHostToDevice, s ) );
                      CK( cudaStreamSynchronize( s ) ); }
  }
  void release() { if (p) cudaFree( p ); p = nullptr; n = 0; }
  ~DevBuf() { release(); }
};

struct Prof {
  std::vector< std::pair< cudaEvent_t, cudaEvent_t > > ev;
  double ms = 0.0; uint64_t n = 0;
};

} // namespace

struct xyst_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, comm_stream = nullptr, aux_stream = nullptr;
  bool own_stream = false;
  xyst_params prm{};
  size_t npoin = 0, NP = 0, nedge = 0, nslot = 0, ntri = 0, nslice = 0, nent = 0;
  uint64_t launches = 0;
  // nodal state, structure of arrays with stride NP (npoin rounded up to 32)
  DevBuf< double > U, Un, W, X, G, vol, v;   // [5][NP] [5][NP] [5][NP] [3][NP] [15][NP]
  DevBuf< double > R, stage, S;              // reference layout [npoin][5]: rhs out, copy staging, source
  int src_mask = 0;
  // edge slots (owner-slice order): endpoints (-1 = padding), normals [3][nslot], fluxes [5][nslot]
  DevBuf< int > ep, eq;
  DevBuf< double > D, F;
  DevBuf< double2 > D2;                  // normals' (x,y) as one 16-byte pair per slot (rows 0,1 of D)
  DevBuf< int2 > inc_eq;                 // (inc_e, inc_q) as one 8-byte pair per entry
  // sliced-ELL node incidence: signed (slot+1) and neighbour node
  DevBuf< long long > sl_base;           // [nslice+1] entry offsets
  DevBuf< int > inc_e, inc_q;
  // boundary faces
  DevBuf< int > tri;                     // [ntri][3]
  DevBuf< unsigned char > besym;         // [ntri][3]
  DevBuf< int > bslot;                   // [npoin] boundary-node slot or -1
  DevBuf< int > bn_node, bn_off, bn_face;// boundary nodes, CSR of (face*4+k)
  DevBuf< double > Gb, Rb;               // [nbn][15], [nbn][5]
  DevBuf< double > fn;                   // [ntri][3] face normals cross(ba,ca)/12 (constant)
  size_t nbn = 0;
  // BCs: union node list with per-node records
  DevBuf< int > bc_node, bc_dir, bc_symoff, bc_faroff, bc_pre;
  DevBuf< int > dir_mask; DevBuf< double > dir_val, sym_n, far_n, pre_val;
  size_t nbc = 0, ndir = 0;
  double far_r = 0, far_p = 0, far_u[3] = {0,0,0};
  // reductions
  DevBuf< double > red; double* red_host = nullptr;
  // halo
  void* comm = nullptr; int nranks = 1, rank = 0;
  std::vector< int > neigh; std::vector< size_t > neigh_off;
  DevBuf< int > sh_node;                 // unique shared nodes
  std::vector< int > sh_node_h; DevBuf< unsigned char > sh_flag;   // host copy; [npoin] 1 = shared (built on first use)
  DevBuf< int > sh_send;                 // [nsend] index into unique list, per neighbour segment
  DevBuf< int > sh_roff, sh_ridx;        // CSR unique node -> positions in recv buffer
  DevBuf< double > sh_part, sh_sendbuf, sh_recvbuf;
  size_t nsh = 0, nsend = 0;
  bool rb_pending = false;              // Rb of the current state already in flight on aux_stream
  // LaxCG: W holds (p,u,v,w,T); Wn = these at time level n. Steady state: local time steps.
  bool lax = false, steady = false;
  double rgas = 0.0, kvinf = 0.0;
  DevBuf< double > Wn, dtp;              // [5][NP], [NP]
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_e = nullptr;
  // ZalCG: integrals stride, FCT parameters, P/Q (C) and low-order solution
  int dstride = 3;
  int maxdeg = 0;                        // largest number of edges at a node
  xyst_zalcg_params zal{ 1, 0, 0, 0, 1.0 };
  DevBuf< double > zP, zQ, zUL;          // [10][NP] [10][NP] [5][NP]
  DevBuf< int > bcof;                    // [npoin] slot in the BC node list or -1
  // KozCG: tetrahedra (sorted by lowest node), node -> (tet, local index) sliced ELL, per-tet buffers
  size_t ntet = 0, knent = 0;
  DevBuf< int > ktet;                    // [4][ntet]
  DevBuf< long long > kbase;             // [nslice+1]
  DevBuf< int > kinc;                    // tet*4+a, -1 = padding
  DevBuf< double > kT, kSc;              // [40][ntet] per-tet contributions; [ntet][5] centroid source
  bool ksrc = false;
  // ChoCG: velocity (3 rotating buffers: time level n, current, next), pressure, divergence,
  // gradients of the CG solution / pressure / velocity, momentum flux; BC lists
  bool cho = false;
  xyst_chocg_params chp{};
  DevBuf< double > cUa, cUb, cUc, cP, cDiv, cSg, cPg, cVg, cFl, cR, cS;
  double *cU = nullptr, *cUn = nullptr, *cUx = nullptr;   // current, time level n, scratch (point into cUa/b/c)
  DevBuf< int > cb_dnode, cb_dmask, cb_snode, cb_soff, cb_nnode;
  DevBuf< double > cb_dval, cb_snorm;
  size_t cb_nd = 0, cb_ns = 0, cb_nn = 0;
  // pressure solve BCs (matrix-free: masked rows/columns), Neumann part, rhs override
  DevBuf< unsigned char > cg_bc;
  DevBuf< double > cg_bcval, cg_neu, cg_rhs0, cg_bcsmall;
  DevBuf< int > cg_bcnode; std::vector< size_t > cg_bcnodes_h;
  bool cg_hasbc = false, cg_hasneu = false, cg_hasrhs0 = false;
  // linear solver: sliced-ELL matrix over scalar rows + CG vectors
  size_t cg_nrow = 0, cg_ncomp = 1, cg_nslice = 0, cg_nent = 0;
  DevBuf< long long > cg_base; DevBuf< int > cg_col; DevBuf< double > cg_val, cg_diag;
  DevBuf< double > cg_x, cg_b, cg_r, cg_p, cg_q, cg_z, cg_d, cg_mask, cg_cnt, cg_scal;
  double cg_normb = 0.0; bool cg_converged = false, cg_finished = false;
  // profiling
  bool prof_on = false;
  std::map< std::string, Prof > prof;
};

namespace {

// ---------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------
struct DParams { double gamma, stab2coef; int flux, stab2, exact; double rgas, kvinf; };
// what the node kernels need to convert between variable sets; rgas > 0 selects LaxCG
struct Mode { double gamma, rgas, kvinf; };

// Primitive variables from conserved ones, Riemann.cpp:211-227
__device__ __forceinline__ void primitive( const double u[NC], double w[NC] ) {
  w[0] = u[0];
  w[1] = u[1] / w[0];
  w[2] = u[2] / w[0];
  w[3] = u[3] / w[0];
  w[4] = u[4] / w[0] - 0.5*(w[1]*w[1] + w[2]*w[2] + w[3]*w[3]);
}

// LaxCG::primitive, LaxCG.cpp:115-137: (r,ru,rv,rw,rE) -> (p,u,v,w,T)
__device__ __forceinline__ void lax_primitive( const double u[NC], double w[NC], double gamma, double rgas ) {
  double r = u[0];
  double uu = u[1]/r, vv = u[2]/r, ww = u[3]/r;
  double p = (u[4] - 0.5*r*(uu*uu + vv*vv + ww*ww)) * (gamma-1.0);
  w[0] = p; w[1] = uu; w[2] = vv; w[3] = ww; w[4] = p/r/rgas;
}
// LaxCG::conservative, LaxCG.cpp:139-164
__device__ __forceinline__ void lax_conservative( const double w[NC], double u[NC], double gamma, double rgas ) {
  double p = w[0], uu = w[1], vv = w[2], ww = w[3], T = w[4];
  double r = p/T/rgas;
  u[0] = r; u[1] = r*uu; u[2] = r*vv; u[3] = r*ww;
  u[4] = p/(gamma-1.0) + 0.5*r*(uu*uu + vv*vv + ww*ww);
}
__device__ __forceinline__ void primitive_of( const double u[NC], double w[NC], const Mode& M ) {
  if (M.rgas > 0.0) lax_primitive( u, w, M.gamma, M.rgas ); else primitive( u, w );
}
// lax::refvel, Lax.cpp:344-360
__device__ __forceinline__ double lax_refvel( double r, double p, double v, double gamma, double kvinf ) {
  return fmin( sqrt( gamma * p / r ), fmax( v, kvinf ) );
}

// 8-byte asynchronous global->shared copy (LDGSTS): in flight without holding a register
__device__ __forceinline__ void cp_async8( double* smem_dst, const double* gsrc )
{
  unsigned d = (unsigned)__cvta_generic_to_shared( smem_dst );
  asm volatile( "cp.async.ca.shared.global [%0], [%1], 8;" :: "r"( d ), "l"( gsrc ) : "memory" );
}
__device__ __forceinline__ void cp_async_commit() { asm volatile( "cp.async.commit_group;" ::: "memory" ); }
// flat index of tk::Fields G(p,i), i = c*3+j, in the structure-of-arrays gradient storage
__host__ __device__ __forceinline__ size_t gidx( int i, size_t p, size_t NP ) { return (size_t)i*NP + p; }
template< int N > __device__ __forceinline__ void cp_async_wait() { asm volatile( "cp.async.wait_group %0;" :: "n"( N ) : "memory" ); }

// Primitive variables and coordinates of a node as four 16-byte pairs, pair k of node p at
// WX[k*NP+p]: (w0,w1) (w2,w3) (w4,x) (y,z) -- what both edge sweeps gather from an edge's other
// end: 3 loads (gradient) or 4 (flux) instead of 5 or 8, still coalesced over consecutive nodes.
// x,y,z are written once at upload; writers of W leave them alone.
__device__ __forceinline__ void load_w( const double2* __restrict__ WX, size_t NP, size_t p, double w[NC] ) {
  double2 a = __ldg( WX + p ), b = __ldg( WX + NP + p ), c = __ldg( WX + 2*NP + p );
  w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x;
}
__device__ __forceinline__ void load_wx( const double2* __restrict__ WX, size_t NP, size_t p, double w[NC], double x[3] ) {
  double2 a = __ldg( WX + p ), b = __ldg( WX + NP + p ), c = __ldg( WX + 2*NP + p ), d = __ldg( WX + 3*NP + p );
  w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; x[0] = c.y; x[1] = d.x; x[2] = d.y;
}
__device__ __forceinline__ void store_w( double* __restrict__ W, size_t NP, size_t p, const double w[NC] ) {
  double2* WX = reinterpret_cast< double2* >( W );
  WX[p] = make_double2( w[0], w[1] );
  WX[NP+p] = make_double2( w[2], w[3] );
  W[(2*NP+p)*2] = w[4];
}
__device__ __forceinline__ double get_w( const double* __restrict__ W, size_t NP, int c, size_t p ) {
  return W[((size_t)(c>>1)*NP + p)*2 + (size_t)(c&1)];
}
// Edge fluxes as (f0,f1) (f2,f3) pairs and f4 per slot
__device__ __forceinline__ void store_f( double* __restrict__ F, size_t nslot, size_t e, const double f[NC] ) {
  double2* F2 = reinterpret_cast< double2* >( F );
  F2[e] = make_double2( f[0], f[1] );
  F2[nslot+e] = make_double2( f[2], f[3] );
  F[4*nslot+e] = f[4];
}
__device__ __forceinline__ void load_f( const double* __restrict__ F, size_t nslot, size_t e, double f[NC] ) {
  const double2* F2 = reinterpret_cast< const double2* >( F );
  double2 a = __ldg( F2 + e ), b = __ldg( F2 + nslot + e );
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = __ldg( F + 4*nslot + e );
}

// reference layout [node][comp] -> SoA state + primitives
__global__ void k_set_state( size_t n, size_t NP, const double* __restrict__ A,
                             double* __restrict__ U, double* __restrict__ W, Mode M )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  double u[NC], w[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) u[c] = A[p*NC+c];
  primitive_of( u, w, M );
  #pragma unroll
  for (int c=0; c<NC; ++c) U[c*NP+p] = u[c];
  store_w( W, NP, p, w );
}

__global__ void k_get_state( size_t n, size_t NP, const double* __restrict__ U, double* __restrict__ A )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  #pragma unroll
  for (int c=0; c<NC; ++c) A[p*NC+c] = U[c*NP+p];
}

// ---------------------------------------------------------------------------------
// boundary-face contributions, gathered per boundary node
//   gradient part: Riemann.cpp:334-360 (incl. the direction-indexed g[j]*n[j] form)
//   flux part    : Riemann.cpp:798-871
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void face_normal( const double* __restrict__ X, size_t NP, const int N[3], double n[3] ) {
  double a[3], b[3], c[3];
  #pragma unroll
  for (int j=0; j<3; ++j) { a[j] = X[j*NP+N[0]]; b[j] = X[j*NP+N[1]]; c[j] = X[j*NP+N[2]]; }
  double ba[3] = { b[0]-a[0], b[1]-a[1], b[2]-a[2] }, ca[3] = { c[0]-a[0], c[1]-a[1], c[2]-a[2] };
  n[0] = (ba[1]*ca[2] - ca[1]*ba[2]) / 12.0;
  n[1] = (ba[2]*ca[0] - ca[2]*ba[0]) / 12.0;
  n[2] = (ba[0]*ca[1] - ca[0]*ba[1]) / 12.0;
}

// the normals depend on coordinates only: computed once at upload with this arithmetic
__global__ void k_face_normals( int ntri, size_t NP, const int* __restrict__ tri, const double* __restrict__ X,
                                double* __restrict__ fn )
{
  int f = blockIdx.x*blockDim.x + threadIdx.x;
  if (f >= ntri) return;
  int N[3] = { tri[f*3+0], tri[f*3+1], tri[f*3+2] };
  double n[3];
  face_normal( X, NP, N, n );
  fn[(size_t)f*3+0] = n[0]; fn[(size_t)f*3+1] = n[1]; fn[(size_t)f*3+2] = n[2];
}

__global__ void k_bnd_grad( int nbn, size_t NP, const int* __restrict__ bn_off, const int* __restrict__ bn_face,
                            const int* __restrict__ tri, const double* __restrict__ fn,
                            const double* __restrict__ W, double* __restrict__ Gb )
{
  int b = blockIdx.x*blockDim.x + threadIdx.x;
  if (b >= nbn) return;
  double acc[15];
  #pragma unroll
  for (int i=0; i<15; ++i) acc[i] = 0.0;
  for (int i=bn_off[b]; i<bn_off[b+1]; ++i) {
    int f = bn_face[i] >> 2;
    int N[3] = { tri[f*3+0], tri[f*3+1], tri[f*3+2] };
    double n[3] = { fn[(size_t)f*3+0], fn[(size_t)f*3+1], fn[(size_t)f*3+2] };
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double u0 = get_w( W, NP, c, N[0] ), u1 = get_w( W, NP, c, N[1] ), u2 = get_w( W, NP, c, N[2] );
      double uab = (u0 + u1)/4.0;
      double ubc = (u1 + u2)/4.0;
      double uca = (u2 + u0)/4.0;
      double g[3] = { uab + uca + u0, uab + ubc + u1, ubc + uca + u2 };
      #pragma unroll
      for (int j=0; j<3; ++j) acc[c*3+j] += g[j] * n[j];
    }
  }
  #pragma unroll
  for (int i=0; i<15; ++i) Gb[(size_t)b*15+i] = acc[i];
}

__global__ void k_bnd_rhs( int nbn, size_t NP, const int* __restrict__ bn_off, const int* __restrict__ bn_face,
                           const int* __restrict__ tri, const unsigned char* __restrict__ besym,
                           const double* __restrict__ fn, const double* __restrict__ U,
                           double* __restrict__ Rb, double gamma, const double* __restrict__ W, double rgas )
{
  int b = blockIdx.x*blockDim.x + threadIdx.x;
  if (b >= nbn) return;
  if (rgas > 0.0) {                       // lax::advbnd, Lax.cpp:840-950, on (p,u,v,w,T)
    double acc[NC] = { 0, 0, 0, 0, 0 };
    for (int i=bn_off[b]; i<bn_off[b+1]; ++i) {
      int f = bn_face[i] >> 2, k = bn_face[i] & 3;
      int N[3] = { tri[f*3+0], tri[f*3+1], tri[f*3+2] };
      double n[3] = { fn[(size_t)f*3+0], fn[(size_t)f*3+1], fn[(size_t)f*3+2] };
      double fl[NC][3];
      #pragma unroll
      for (int m=0; m<3; ++m) {
        double pr = get_w( W, NP, 0, N[m] ), uu = get_w( W, NP, 1, N[m] ), vv = get_w( W, NP, 2, N[m] ), ww = get_w( W, NP, 3, N[m] ), T = get_w( W, NP, 4, N[m] );
        double rA = pr/T/rgas;
        double ruA = uu * rA, rvA = vv * rA, rwA = ww * rA;
        double reA = pr/(gamma-1.0) + 0.5*(ruA*ruA + rvA*rvA + rwA*rwA)/rA;
        double vn = besym[f*3+m] ? 0.0 : (n[0]*uu + n[1]*vv + n[2]*ww);
        fl[0][m] = rA*vn;
        fl[1][m] = ruA*vn + pr*n[0];
        fl[2][m] = rvA*vn + pr*n[1];
        fl[3][m] = rwA*vn + pr*n[2];
        fl[4][m] = (reA + pr)*vn;
      }
      #pragma unroll
      for (int c=0; c<NC; ++c) {
        double fab = (fl[c][0] + fl[c][1])/4.0;
        double fbc = (fl[c][1] + fl[c][2])/4.0;
        double fca = (fl[c][2] + fl[c][0])/4.0;
        double add = k == 0 ? fab + fca + fl[c][0] : (k == 1 ? fab + fbc + fl[c][1] : fbc + fca + fl[c][2]);
        acc[c] += add;
      }
    }
    #pragma unroll
    for (int c=0; c<NC; ++c) Rb[(size_t)b*NC+c] = acc[c];
    return;
  }
  double acc[NC] = { 0, 0, 0, 0, 0 };
  for (int i=bn_off[b]; i<bn_off[b+1]; ++i) {
    int f = bn_face[i] >> 2, k = bn_face[i] & 3;
    int N[3] = { tri[f*3+0], tri[f*3+1], tri[f*3+2] };
    double n[3] = { fn[(size_t)f*3+0], fn[(size_t)f*3+1], fn[(size_t)f*3+2] };
    double fl[NC][3];
    #pragma unroll
    for (int m=0; m<3; ++m) {
      double r = U[N[m]], ru = U[NP+N[m]], rv = U[2*NP+N[m]], rw = U[3*NP+N[m]], re = U[4*NP+N[m]];
      double p = (re - 0.5*(ru*ru + rv*rv + rw*rw)/r) * (gamma-1.0);
      double vn = besym[f*3+m] ? 0.0 : (n[0]*ru + n[1]*rv + n[2]*rw)/r;
      fl[0][m] = r*vn;
      fl[1][m] = ru*vn + p*n[0];
      fl[2][m] = rv*vn + p*n[1];
      fl[3][m] = rw*vn + p*n[2];
      fl[4][m] = (re + p)*vn;
    }
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double fab = (fl[c][0] + fl[c][1])/4.0;
      double fbc = (fl[c][1] + fl[c][2])/4.0;
      double fca = (fl[c][2] + fl[c][0])/4.0;
      double add = k == 0 ? fab + fca + fl[c][0] : (k == 1 ? fab + fbc + fl[c][1] : fbc + fca + fl[c][2]);
      acc[c] += add;
    }
  }
  #pragma unroll
  for (int c=0; c<NC; ++c) Rb[(size_t)b*NC+c] = acc[c];
}

// ---------------------------------------------------------------------------------
// gradient gather: one warp per 32-node slice, one thread per node
//   G(p) = [ sum_edges -/+ d*(w_q + w_p)  +  boundary part ] / vol(p)
// entry = signed edge slot: +(slot+1) if p is the edge's second node (receives +f),
// -(slot+1) if it is the first (receives -f), 0 = padding (multiplier 0 on slot 0)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void grad_sum( size_t p, int lane, long long base, int kmax,
    const int2* __restrict__ inc_eq, const double2* __restrict__ D2, const double* __restrict__ D,
    size_t nslot, const double* __restrict__ W, size_t NP, double acc[15] )
{
  const double2* WX = reinterpret_cast< const double2* >( W );
  double wp[NC];
  load_w( WX, NP, p, wp );
  #pragma unroll
  for (int i=0; i<15; ++i) acc[i] = 0.0;
  #pragma unroll kGradUnroll
  for (int k=0; k<kmax; ++k) {
    long long i = base + (long long)k*32 + lane;
    int2 eq = __ldg( inc_eq + i );
    int se = eq.x, q = eq.y;
    double sg = se > 0 ? 1.0 : (se < 0 ? -1.0 : 0.0);
    size_t sl = se == 0 ? 0 : (size_t)(abs(se)-1);
    double2 d01 = __ldg( D2 + sl );
    double d0 = sg * d01.x, d1 = sg * d01.y, d2 = sg * __ldg( D + 2*nslot + sl );
    double wq[NC];
    load_w( WX, NP, (size_t)q, wq );
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double s = wq[c] + wp[c];
      acc[c*3+0] += d0 * s;
      acc[c*3+1] += d1 * s;
      acc[c*3+2] += d2 * s;
    }
  }
}

__global__ void __launch_bounds__(NODE_THREADS, GRAD_MINB)
k_grad_node( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq,
             const double2* __restrict__ D2, const double* __restrict__ D, size_t nslot,
             const double* __restrict__ W, const int* __restrict__ bslot, const double* __restrict__ Gb,
             const double* __restrict__ vol, double* __restrict__ G, int defer_bnd )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[15];
  grad_sum( p, lane, base, kmax, inc_eq, D2, D, nslot, W, NP, acc );
  int b = bslot[p];
  if (b >= 0) {
    if (defer_bnd) {          // boundary part and division follow in k_grad_bfix (same operation order)
      #pragma unroll
      for (int i=0; i<15; ++i) G[i*NP+p] = acc[i];
      return;
    }
    #pragma unroll
    for (int i=0; i<15; ++i) acc[i] += Gb[(size_t)b*15+i];
  }
  double vp = vol[p];
  #pragma unroll
  for (int i=0; i<15; ++i) G[i*NP+p] = acc[i]/vp;
}

// boundary nodes: G = (domain sum + boundary sum) / vol, once both are known
__global__ void k_grad_bfix( int nbn, size_t NP, const int* __restrict__ bn_node, const double* __restrict__ Gb,
                             const double* __restrict__ vol, double* __restrict__ G )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nbn*15) return;
  size_t b = i / 15, k = i % 15;
  size_t p = bn_node[b];
  size_t g = gidx( (int)k, p, NP );
  G[g] = (G[g] + Gb[b*15+k]) / vol[p];
}

// partial (un-normalised) gradient sums of the shared nodes, for the halo exchange
__global__ void k_grad_shared( int nsh, size_t NP, const int* __restrict__ sh_node,
             const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq,
             const double2* __restrict__ D2, const double* __restrict__ D, size_t nslot,
             const double* __restrict__ W, const int* __restrict__ bslot,
             const double* __restrict__ Gb, double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  size_t slice = p >> 5; int lane = p & 31;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[15];
  grad_sum( p, lane, base, kmax, inc_eq, D2, D, nslot, W, NP, acc );
  int b = bslot[p];
  if (b >= 0) for (int j=0; j<15; ++j) acc[j] += Gb[(size_t)b*15+j];
  for (int j=0; j<15; ++j) part[(size_t)i*15+j] = acc[j];
}

__global__ void k_pack( int nsend, int w, const int* __restrict__ sh_send,
                        const double* __restrict__ part, double* __restrict__ sendbuf )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nsend*w) return;
  size_t s = i / w, c = i % w;
  sendbuf[i] = part[(size_t)sh_send[s]*w + c];
}

__global__ void k_grad_finish( int nsh, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
             const int* __restrict__ ridx, const double* __restrict__ part,
             const double* __restrict__ recvbuf, const double* __restrict__ vol, double* __restrict__ G )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  double vp = vol[p];
  for (int j=0; j<15; ++j) {
    double a = part[(size_t)i*15+j];
    for (int r=roff[i]; r<roff[i+1]; ++r) a += recvbuf[(size_t)ridx[r]*15+j];
    G[gidx( j, p, NP )] = a / vp;
  }
}

// ---------------------------------------------------------------------------------
// edge fluxes: MUSCL (Riemann.cpp:34-143) + Rusanov (:369-478) or HLLC (:480-650)
// ---------------------------------------------------------------------------------
#define MUSCL_EPS 1.0e-9
#define MUSCL_K (1.0/3.0)
#ifndef MUSCL_V2
#define MUSCL_V2 1        // 0: the first form of the two-reciprocal limiter (A/B timing)
#endif

// van Leer limited extrapolation increments for one component.
// exact: the reference's expression tree (8 divisions). fast: with a = d2+eps, b = d1+eps
// the two limiter values are phi(a/b) = 2a/(a+b) and phi(b/a) = 2b/(a+b) when a and b
// have the same sign and 0 otherwise, i.e. one reciprocal per side.
// 1/x to ~1 ulp without the IEEE division's slow path: hardware seed (MUFU.RCP64H, ~20
// bits) + two Newton steps. Only used by the non-"exact" limiter form; x is a sum of two
// same-signed numbers of magnitude >= 1e-9 whenever the result is used.
__device__ __forceinline__ double fast_rcp( double x )
{
  double y;
  asm( "rcp.approx.ftz.f64 %0, %1;" : "=d"( y ) : "d"( x ) );
#if MUSCL_V2
  // one cubically convergent step: y (1 + e + e^2), e = 1 - x y  (seed error 2^-20 -> 2^-60)
  double e = fma( -x, y, 1.0 );
  double t = fma( e, e, e );
  return fma( y, t, y );
#else
  double e = fma( -x, y, 1.0 );
  y = fma( y, e, y );
  e = fma( -x, y, 1.0 );
  y = fma( y, e, y );
  return y;
#endif
}

template< bool EXACT >
__device__ __forceinline__ void vanleer( double d1, double d2, double d3, double& incL, double& incR )
{
  if (EXACT) {
    double rcL = (d2 + MUSCL_EPS) / (d1 + MUSCL_EPS);
    double rcR = (d2 + MUSCL_EPS) / (d3 + MUSCL_EPS);
    double rLinv = (d1 + MUSCL_EPS) / (d2 + MUSCL_EPS);
    double rRinv = (d3 + MUSCL_EPS) / (d2 + MUSCL_EPS);
    double phiL = (fabs(rcL) + rcL) / (fabs(rcL) + 1.0);
    double phiR = (fabs(rcR) + rcR) / (fabs(rcR) + 1.0);
    double phi_L_inv = (fabs(rLinv) + rLinv) / (fabs(rLinv) + 1.0);
    double phi_R_inv = (fabs(rRinv) + rRinv) / (fabs(rRinv) + 1.0);
    incL = 0.25*(d1*(1.0-MUSCL_K)*phiL + d2*(1.0+MUSCL_K)*phi_L_inv);
    incR = 0.25*(d3*(1.0-MUSCL_K)*phiR + d2*(1.0+MUSCL_K)*phi_R_inv);
  } else {
#if MUSCL_V2
    // phi(a/b) = 2a/(a+b), phi(b/a) = 2b/(a+b) for same-signed a, b (else 0), so that
    // inc = 0.25 [ d1 (1-k) 2a + d2 (1+k) 2b ] / (a+b): one reciprocal, five multiply-adds per side
    const double c1 = 0.5*(1.0-MUSCL_K), c2 = 0.5*(1.0+MUSCL_K);
    double a = d2 + MUSCL_EPS, bL = d1 + MUSCL_EPS, bR = d3 + MUSCL_EPS;
    double t2 = c2*d2;
    double vL = fma( c1*d1, a, t2*bL ) * fast_rcp( a + bL );
    double vR = fma( c1*d3, a, t2*bR ) * fast_rcp( a + bR );
    // same strict sign <=> positive product (|a|, |b| are >= ~1e-9 unless a difference hits -1e-9
    // to the last bit, so the product cannot underflow in practice); integer sign-bit tests cost
    // 5 % more kernel time, the kernel being limited by instruction issue
    incL = a*bL > 0.0 ? vL : 0.0;
    incR = a*bR > 0.0 ? vR : 0.0;
#else
    double a = d2 + MUSCL_EPS, bL = d1 + MUSCL_EPS, bR = d3 + MUSCL_EPS;
    bool sL = (a > 0.0 && bL > 0.0) || (a < 0.0 && bL < 0.0);
    bool sR = (a > 0.0 && bR > 0.0) || (a < 0.0 && bR < 0.0);
    double iL = 2.0 * fast_rcp( a + bL ), iR = 2.0 * fast_rcp( a + bR );
    double phiL = sL ? a*iL : 0.0, phi_L_inv = sL ? bL*iL : 0.0;
    double phiR = sR ? a*iR : 0.0, phi_R_inv = sR ? bR*iR : 0.0;
    incL = 0.25*(d1*(1.0-MUSCL_K)*phiL + d2*(1.0+MUSCL_K)*phi_L_inv);
    incR = 0.25*(d3*(1.0-MUSCL_K)*phiR + d2*(1.0+MUSCL_K)*phi_R_inv);
#endif
  }
}

// gp/gq: the 15 gradient components of the two end nodes, element i at gp[i*gsp], gq[i*gsq]
template< bool EXACT >
__device__ __forceinline__ void muscl( const double* gp, int gsp, const double* gq, int gsq,
                                       const double vw[3], double l[NC], double r[NC] )
{
  double ls[NC], rs[NC], d1[NC], d3[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    ls[c] = l[c]; rs[c] = r[c];
    double g1 = gp[(c*3+0)*gsp]*vw[0] + gp[(c*3+1)*gsp]*vw[1] + gp[(c*3+2)*gsp]*vw[2];
    double g2 = gq[(c*3+0)*gsq]*vw[0] + gq[(c*3+1)*gsq]*vw[1] + gq[(c*3+2)*gsq]*vw[2];
    double delta2 = r[c] - l[c];
    d1[c] = 2.0 * g1 - delta2;
    d3[c] = 2.0 * g2 - delta2;
    double incL, incR;
    vanleer< EXACT >( d1[c], delta2, d3[c], incL, incR );
    l[c] += incL;
    r[c] -= incR;
  }
  // first order where density or internal energy could turn negative (:129-130)
  if (ls[0] < d1[0] || ls[4] < d1[4]) {
    #pragma unroll
    for (int c=0; c<NC; ++c) l[c] = ls[c];
  }
  if (rs[0] < -d3[0] || rs[4] < -d3[4]) {
    #pragma unroll
    for (int c=0; c<NC; ++c) r[c] = rs[c];
  }
}

__device__ __forceinline__ void rusanov( double l[NC], double r[NC], const double n[3],
                                         const DParams& P, double f[NC] )
{
  double g = P.gamma;
  double pL = (l[0]*l[4]) * (g-1.0);
  double pR = (r[0]*r[4]) * (g-1.0);
  const double gg1 = g*(g-1.0), eL = l[4], eR = r[4];
  double nx = n[0], ny = n[1], nz = n[2];
  double vnL = l[1]*nx + l[2]*ny + l[3]*nz;
  double vnR = r[1]*nx + r[2]*ny + r[3]*nz;
  l[4] = (l[4] + 0.5*(l[1]*l[1] + l[2]*l[2] + l[3]*l[3])) * l[0];
  l[1] *= l[0]; l[2] *= l[0]; l[3] *= l[0];
  r[4] = (r[4] + 0.5*(r[1]*r[1] + r[2]*r[2] + r[3]*r[3])) * r[0];
  r[1] *= r[0]; r[2] *= r[0]; r[3] *= r[0];
  double len = sqrt( nx*nx + ny*ny + nz*nz );
#if MUSCL_V2
  double sl, sr;
  if (P.exact) { sl = fabs(vnL) + sqrt( g * pL / l[0] )*len; sr = fabs(vnR) + sqrt( g * pR / r[0] )*len; }
  else { sl = fabs(vnL) + sqrt( gg1 * eL )*len; sr = fabs(vnR) + sqrt( gg1 * eR )*len; }   // g p / rho = g (g-1) e
#else
  double sl = fabs(vnL) + sqrt( g * pL / l[0] )*len;
  double sr = fabs(vnR) + sqrt( g * pR / r[0] )*len;
#endif
  double fw = fmax( sl, sr );
  f[0] = l[0]*vnL + r[0]*vnR + fw*(r[0] - l[0]);
  f[1] = l[1]*vnL + r[1]*vnR + (pL + pR)*nx + fw*(r[1] - l[1]);
  f[2] = l[2]*vnL + r[2]*vnR + (pL + pR)*ny + fw*(r[2] - l[2]);
  f[3] = l[3]*vnL + r[3]*vnR + (pL + pR)*nz + fw*(r[3] - l[3]);
  f[4] = (l[4] + pL)*vnL + (r[4] + pR)*vnR + fw*(r[4] - l[4]);
  if (P.stab2) {
    double fws = P.stab2coef * fw;
    #pragma unroll
    for (int c=0; c<NC; ++c) f[c] -= fws*(l[c] - r[c]);
  }
}

__device__ __forceinline__ void hllc( double l[NC], double r[NC], const double n[3],
                                      const DParams& P, double f[NC] )
{
  double g = P.gamma;
  double nx = -n[0], ny = -n[1], nz = -n[2];
  double len = sqrt( nx*nx + ny*ny + nz*nz );
  nx /= len; ny /= len; nz /= len;
  double qL = l[1]*nx + l[2]*ny + l[3]*nz;
  double qR = r[1]*nx + r[2]*ny + r[3]*nz;
  double pL = (l[0]*l[4]) * (g-1.0);
  double pR = (r[0]*r[4]) * (g-1.0);
  l[4] = (l[4] + 0.5*(l[1]*l[1] + l[2]*l[2] + l[3]*l[3])) * l[0];
  l[1] *= l[0]; l[2] *= l[0]; l[3] *= l[0];
  r[4] = (r[4] + 0.5*(r[1]*r[1] + r[2]*r[2] + r[3]*r[3])) * r[0];
  r[1] *= r[0]; r[2] *= r[0]; r[3] *= r[0];
  double cL = sqrt( g * pL / l[0] );
  double cR = sqrt( g * pR / r[0] );
  double sL = fmin( qL - cL, qR - cR );
  double sR = fmax( qL + cL, qR + cR );
  double tL = sL - qL;
  double tR = sR - qR;
  double sM = (r[0]*qR*tR - l[0]*qL*tL + pL - pR) / (r[0]*tR - l[0]*tL);
  double pS = pL - l[0]*tL*(qL - sM);
  double uL[NC], uR[NC];
  double s = sL - sM;
  uL[0] = tL*l[0]/s;
  uL[1] = (tL*l[1] + (pS-pL)*nx)/s;
  uL[2] = (tL*l[2] + (pS-pL)*ny)/s;
  uL[3] = (tL*l[3] + (pS-pL)*nz)/s;
  uL[4] = (tL*l[4] - pL*qL + pS*sM)/s;
  s = sR - sM;
  uR[0] = tR*r[0]/s;
  uR[1] = (tR*r[1] + (pS-pR)*nx)/s;
  uR[2] = (tR*r[2] + (pS-pR)*ny)/s;
  uR[3] = (tR*r[3] + (pS-pR)*nz)/s;
  uR[4] = (tR*r[4] - pR*qR + pS*sM)/s;
  double L2 = -2.0*len;
  nx *= L2; ny *= L2; nz *= L2;
  if (sL > 0.0) {
    double qL2 = qL * L2;
    f[0] = l[0]*qL2;
    f[1] = l[1]*qL2 + pL*nx;
    f[2] = l[2]*qL2 + pL*ny;
    f[3] = l[3]*qL2 + pL*nz;
    f[4] = (l[4] + pL)*qL2;
  } else if (sL <= 0.0 && sM > 0.0) {
    double qL2 = qL * L2, sL2 = sL * L2;
    f[0] = l[0]*qL2 + sL2*(uL[0] - l[0]);
    f[1] = l[1]*qL2 + pL*nx + sL2*(uL[1] - l[1]);
    f[2] = l[2]*qL2 + pL*ny + sL2*(uL[2] - l[2]);
    f[3] = l[3]*qL2 + pL*nz + sL2*(uL[3] - l[3]);
    f[4] = (l[4] + pL)*qL2 + sL2*(uL[4] - l[4]);
  } else if (sM <= 0.0 && sR >= 0.0) {
    double qR2 = qR * L2, sR2 = sR * L2;
    f[0] = r[0]*qR2 + sR2*(uR[0] - r[0]);
    f[1] = r[1]*qR2 + pR*nx + sR2*(uR[1] - r[1]);
    f[2] = r[2]*qR2 + pR*ny + sR2*(uR[2] - r[2]);
    f[3] = r[3]*qR2 + pR*nz + sR2*(uR[3] - r[3]);
    f[4] = (r[4] + pR)*qR2 + sR2*(uR[4] - r[4]);
  } else {
    double qR2 = qR * L2;
    f[0] = r[0]*qR2;
    f[1] = r[1]*qR2 + pR*nx;
    f[2] = r[2]*qR2 + pR*ny;
    f[3] = r[3]*qR2 + pR*nz;
    f[4] = (r[4] + pR)*qR2;
  }
  if (P.stab2) {
    double sl = fabs(qL) + cL, sr = fabs(qR) + cR;
    double fws = P.stab2coef * fmax(sl,sr) * len;
    #pragma unroll
    for (int c=0; c<NC; ++c) f[c] -= fws * (l[c] - r[c]);
  }
}

// lax::sigvel, Lax.cpp:362-388: signal velocities of the preconditioned system
__device__ __forceinline__ void lax_sigvel( double p, double T, double v, double vn, const DParams& P,
                                            double& vpri, double& cpri )
{
  double g = P.gamma, rgas = P.rgas;
  double cp = g*rgas/(g-1.0);
  double r = p/T/rgas;
  double rp = r/p;
  double rt = -r/T;
  double vr = lax_refvel( r, p, v, g, P.kvinf );
  double vr2 = vr*vr;
  double beta = rp + rt/r/cp;
  double alpha = 0.5*(1.0 - beta*vr2);
  vpri = vn*(1.0 - alpha);
  cpri = sqrt( alpha*alpha*vn*vn + vr2 );
}

// edge-end state (p,u,v,w,T) -> conserved, Lax.cpp:438-451
__device__ __forceinline__ void lax_edge_conserved( double l[NC], double pL, const DParams& P ) {
  l[0] = pL/l[4]/P.rgas;
  l[1] *= l[0]; l[2] *= l[0]; l[3] *= l[0];
  l[4] = pL/(P.gamma-1.0) + 0.5*(l[1]*l[1] + l[2]*l[2] + l[3]*l[3])/l[0];
}

// lax::rusanov, Lax.cpp:390-511
__device__ __forceinline__ void lax_rusanov( double l[NC], double r[NC], const double n[3],
                                             const DParams& P, double f[NC] )
{
  double nx = n[0], ny = n[1], nz = n[2];
  double vnL = l[1]*nx + l[2]*ny + l[3]*nz;
  double vnR = r[1]*nx + r[2]*ny + r[3]*nz;
  double pL = l[0], pR = r[0];
  double len = sqrt( nx*nx + ny*ny + nz*nz );
  double vpL, cpL, vpR, cpR;
  lax_sigvel( l[0], l[4], sqrt( l[1]*l[1] + l[2]*l[2] + l[3]*l[3] ), vnL, P, vpL, cpL );
  lax_sigvel( r[0], r[4], sqrt( r[1]*r[1] + r[2]*r[2] + r[3]*r[3] ), vnR, P, vpR, cpR );
  lax_edge_conserved( l, pL, P );
  lax_edge_conserved( r, pR, P );
  double sp = fmax( fabs(vpL-cpL), fmax( fabs(vpR-cpR), fmax( fabs(vpL+cpL), fabs(vpR+cpR) ) ) );
  double fw = fmax( -sp, sp ) * len;
  f[0] = l[0]*vnL + r[0]*vnR + fw*(r[0] - l[0]);
  f[1] = l[1]*vnL + r[1]*vnR + (pL + pR)*nx + fw*(r[1] - l[1]);
  f[2] = l[2]*vnL + r[2]*vnR + (pL + pR)*ny + fw*(r[2] - l[2]);
  f[3] = l[3]*vnL + r[3]*vnR + (pL + pR)*nz + fw*(r[3] - l[3]);
  f[4] = (l[4] + pL)*vnL + (r[4] + pR)*vnR + fw*(r[4] - l[4]);
  if (P.stab2) {
    double fws = P.stab2coef * fw;
    #pragma unroll
    for (int c=0; c<NC; ++c) f[c] -= fws*(l[c] - r[c]);
  }
}

// lax::hllc, Lax.cpp:513-723 (wave speed option 3: symmetric +-sp; no artificial viscosity)
__device__ __forceinline__ void lax_hllc( double l[NC], double r[NC], const double n[3],
                                          const DParams& P, double f[NC] )
{
  double nx = -n[0], ny = -n[1], nz = -n[2];
  double len = sqrt( nx*nx + ny*ny + nz*nz );
  nx /= len; ny /= len; nz /= len;
  double qL = l[1]*nx + l[2]*ny + l[3]*nz;
  double qR = r[1]*nx + r[2]*ny + r[3]*nz;
  double pL = l[0], pR = r[0];
  double vpL, cpL, vpR, cpR;
  lax_sigvel( l[0], l[4], sqrt( l[1]*l[1] + l[2]*l[2] + l[3]*l[3] ), qL*len, P, vpL, cpL );
  lax_sigvel( r[0], r[4], sqrt( r[1]*r[1] + r[2]*r[2] + r[3]*r[3] ), qR*len, P, vpR, cpR );
  lax_edge_conserved( l, pL, P );
  lax_edge_conserved( r, pR, P );
  double sp = fmax( fabs(vpL-cpL), fmax( fabs(vpR-cpR), fmax( fabs(vpL+cpL), fabs(vpR+cpR) ) ) );
  double sL = -sp, sR = +sp;
  double tL = sL - qL;
  double tR = sR - qR;
  double sM = (r[0]*qR*tR - l[0]*qL*tL + pL - pR) / (r[0]*tR - l[0]*tL);
  double pS = pL - l[0]*tL*(qL - sM);
  double uL[NC], uR[NC];
  double s = sL - sM;
  uL[0] = tL*l[0]/s;
  uL[1] = (tL*l[1] + (pS-pL)*nx)/s;
  uL[2] = (tL*l[2] + (pS-pL)*ny)/s;
  uL[3] = (tL*l[3] + (pS-pL)*nz)/s;
  uL[4] = (tL*l[4] - pL*qL + pS*sM)/s;
  s = sR - sM;
  uR[0] = tR*r[0]/s;
  uR[1] = (tR*r[1] + (pS-pR)*nx)/s;
  uR[2] = (tR*r[2] + (pS-pR)*ny)/s;
  uR[3] = (tR*r[3] + (pS-pR)*nz)/s;
  uR[4] = (tR*r[4] - pR*qR + pS*sM)/s;
  double L2 = -2.0*len;
  nx *= L2; ny *= L2; nz *= L2;
  if (sL > 0.0) {
    double qL2 = qL * L2;
    f[0] = l[0]*qL2;
    f[1] = l[1]*qL2 + pL*nx;
    f[2] = l[2]*qL2 + pL*ny;
    f[3] = l[3]*qL2 + pL*nz;
    f[4] = (l[4] + pL)*qL2;
  } else if (sL <= 0.0 && sM > 0.0) {
    double qL2 = qL * L2, sL2 = sL * L2;
    f[0] = l[0]*qL2 + sL2*(uL[0] - l[0]);
    f[1] = l[1]*qL2 + pL*nx + sL2*(uL[1] - l[1]);
    f[2] = l[2]*qL2 + pL*ny + sL2*(uL[2] - l[2]);
    f[3] = l[3]*qL2 + pL*nz + sL2*(uL[3] - l[3]);
    f[4] = (l[4] + pL)*qL2 + sL2*(uL[4] - l[4]);
  } else if (sM <= 0.0 && sR >= 0.0) {
    double qR2 = qR * L2, sR2 = sR * L2;
    f[0] = r[0]*qR2 + sR2*(uR[0] - r[0]);
    f[1] = r[1]*qR2 + pR*nx + sR2*(uR[1] - r[1]);
    f[2] = r[2]*qR2 + pR*ny + sR2*(uR[2] - r[2]);
    f[3] = r[3]*qR2 + pR*nz + sR2*(uR[3] - r[3]);
    f[4] = (r[4] + pR)*qR2 + sR2*(uR[4] - r[4]);
  } else {
    double qR2 = qR * L2;
    f[0] = r[0]*qR2;
    f[1] = r[1]*qR2 + pR*nx;
    f[2] = r[2]*qR2 + pR*ny;
    f[3] = r[3]*qR2 + pR*nz;
    f[4] = (r[4] + pR)*qR2;
  }
}

// one thread per edge slot; a warp covers the j-th owned edge of 32 consecutive nodes.
// The 30 gradient values of the two end nodes are fetched with cp.async straight into a
// per-thread column of shared memory: the copies need no registers while in flight, so
// every thread has its whole working set (46 doubles) outstanding at once and the kernel
// still fits enough warps per SM to cover the latency; the limiter then reads its
// operands from shared memory as it goes.
template< bool EXACT, int FLUX >
__global__ void __launch_bounds__(FLUX_THREADS, FLUX_MINB)
k_flux_edge( size_t nslot, size_t NP, const int* __restrict__ ep, const int* __restrict__ eq,
             const double* __restrict__ D, const double* __restrict__ W, const double* __restrict__ X,
             const double* __restrict__ G, double* __restrict__ F, DParams P, size_t e0, size_t e1 )
{
  __shared__ double sg[30*FLUX_THREADS];
  size_t e = e0 + blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= e1) return;
  int pi = ep[e];
  if (pi < 0) return;                      // padding slot
  size_t p = pi, q = eq[e];
  double* gp = sg + threadIdx.x;
  double* gq = gp + 15*FLUX_THREADS;
  #pragma unroll
  for (int i=0; i<15; ++i) { cp_async8( gp + i*FLUX_THREADS, G + i*NP + p ); cp_async8( gq + i*FLUX_THREADS, G + i*NP + q ); }
  cp_async_commit();
  double n[3] = { D[e], D[nslot+e], D[2*nslot+e] };
  double l[NC], r[NC], vw[3], xp[3];
  const double2* WX = reinterpret_cast< const double2* >( W );
  load_wx( WX, NP, p, l, xp );
  load_wx( WX, NP, q, r, vw );
  #pragma unroll
  for (int j=0; j<3; ++j) vw[j] -= xp[j];
  cp_async_wait< 0 >();
  muscl< EXACT >( gp, FLUX_THREADS, gq, FLUX_THREADS, vw, l, r );
  double f[NC];
  if (FLUX == 0) rusanov( l, r, n, P, f ); else if (FLUX == 1) hllc( l, r, n, P, f );
  else if (FLUX == 2) lax_rusanov( l, r, n, P, f ); else lax_hllc( l, r, n, P, f );
  store_f( F, nslot, e, f );
}

// ---------------------------------------------------------------------------------
// flux gather per node (+ boundary + source) and, fused, the RK stage update
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void rhs_sum( size_t p, int lane, long long base, int kmax,
    const int* __restrict__ inc_e, const double* __restrict__ F, size_t nslot,
    const int* __restrict__ bslot, const double* __restrict__ Rb, const double* __restrict__ S,
    int src_mask, const double* __restrict__ v, double acc[NC] )
{
  #pragma unroll
  for (int c=0; c<NC; ++c) acc[c] = 0.0;
  // sg*f is exact, so this equals the add/subtract of the reference's scatter
  #pragma unroll kRhsUnroll
  for (int k=0; k<kmax; ++k) {
    int se = __ldg( inc_e + base + (long long)k*32 + lane );
    double sg = se > 0 ? 1.0 : (se < 0 ? -1.0 : 0.0);
    size_t sl = se == 0 ? 0 : (size_t)(abs(se)-1);
    double f[NC];
    load_f( F, nslot, sl, f );
    #pragma unroll
    for (int c=0; c<NC; ++c) acc[c] = fma( sg, f[c], acc[c] );
  }
  int b = bslot[p];
  if (b >= 0) {
    #pragma unroll
    for (int c=0; c<NC; ++c) acc[c] += Rb[(size_t)b*NC+c];
  }
  if (src_mask) {
    double vp = v[p];
    #pragma unroll
    for (int c=0; c<NC; ++c) if (src_mask & (1<<c)) acc[c] -= S[p*NC+c] * vp;
  }
}

// RK stage update of one node from its summed rhs.
//   RieCG::solve, RieCG.cpp:1011-1021:  u = un - rk dt rhs / vol   (dt = local dtp if steady)
//   LaxCG::solve, LaxCG.cpp:1150-1176:  (p,u,v,w,T) = (..)_n + P^-1 (-rk dt rhs / vol), then
//   back to conserved variables; W keeps what LaxCG::primitive gives for the next stage.
struct StageArgs { double rk, dt; const double* dtp; int stage; Mode M; };

template< bool LAX >
__device__ __forceinline__ void node_update( size_t p, size_t NP, const double acc[NC], double vp,
    const double* __restrict__ Un, double* __restrict__ U, double* __restrict__ W,
    double* __restrict__ Wn, double* __restrict__ UnOut, const StageArgs& A )
{
  double dtl = A.dtp ? A.dtp[p] : A.dt;
  double u[NC], w[NC];
  if (LAX) {
    double g = A.M.gamma, rgas = A.M.rgas;
    double wn[NC];
    #pragma unroll
    for (int c=0; c<NC; ++c) w[c] = get_w( W, NP, c, p );
    if (A.stage == 0) {
      #pragma unroll
      for (int c=0; c<NC; ++c) { wn[c] = w[c]; Wn[c*NP+p] = w[c]; }
    } else {
      #pragma unroll
      for (int c=0; c<NC; ++c) wn[c] = Wn[c*NP+p];
    }
    double R = -A.rk * dtl / vp;
    // inverse of the time-derivative preconditioning matrix, LaxCG::precond :166-226
    double pr = w[0], uu = w[1], vv = w[2], ww = w[3], T = w[4];
    double r = pr/T/rgas;
    double cp = g*rgas/(g-1.0);
    double k = uu*uu + vv*vv + ww*ww;
    double vr = lax_refvel( r, pr, sqrt(k), g, A.M.kvinf );
    double vr2 = vr*vr;
    double rt = -r/T;
    double H = cp*T + k/2.0;
    double theta = 1.0/vr2 - rt/r/cp;
    double coef = r*cp*theta + rt;
    double q[NC] = { R*acc[0], R*acc[1], R*acc[2], R*acc[3], R*acc[4] };
    double wnew[NC];
    wnew[0] = wn[0] + (rt*(H - k) + r*cp)/coef*q[0] + rt*uu/coef*q[1] + rt*vv/coef*q[2] + rt*ww/coef*q[3] + (-rt/coef)*q[4];
    wnew[1] = wn[1] + (-uu/r)*q[0] + 1.0/r*q[1] + 0.0*q[2] + 0.0*q[3] + 0.0*q[4];
    wnew[2] = wn[2] + (-vv/r)*q[0] + 0.0*q[1] + 1.0/r*q[2] + 0.0*q[3] + 0.0*q[4];
    wnew[3] = wn[3] + (-ww/r)*q[0] + 0.0*q[1] + 0.0*q[2] + 1.0/r*q[3] + 0.0*q[4];
    wnew[4] = wn[4] + (-(theta*(H - k) - 1.0)/coef)*q[0] + (-theta*uu/coef)*q[1] + (-theta*vv/coef)*q[2]
                    + (-theta*ww/coef)*q[3] + theta/coef*q[4];
    lax_conservative( wnew, u, g, rgas );
    lax_primitive( u, w, g, rgas );
    #pragma unroll
    for (int c=0; c<NC; ++c) U[c*NP+p] = u[c];
    store_w( W, NP, p, w );
    if (A.stage == 2) {                   // conservative( m_un ) for the diagnostics, :1196
      double un[NC];
      lax_conservative( wn, un, g, rgas );
      #pragma unroll
      for (int c=0; c<NC; ++c) UnOut[c*NP+p] = un[c];
    }
  } else {
    double rkdt = A.rk * dtl;
    #pragma unroll
    for (int c=0; c<NC; ++c) { u[c] = Un[c*NP+p] - rkdt * acc[c] / vp; U[c*NP+p] = u[c]; }
    primitive( u, w );
    store_w( W, NP, p, w );
  }
}

template< bool FUSED, bool LAX >
__global__ void __launch_bounds__(NODE_THREADS, RHS_MINB)
k_rhs_node( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int* __restrict__ inc_e,
            const double* __restrict__ F, size_t nslot, const int* __restrict__ bslot,
            const double* __restrict__ Rb, const double* __restrict__ S, int src_mask,
            const double* __restrict__ v, const double* __restrict__ vol, const double* __restrict__ Un,
            StageArgs A, double* __restrict__ U, double* __restrict__ W, double* __restrict__ R,
            double* __restrict__ Wn, double* __restrict__ UnOut, size_t slice0, size_t slice1,
            const unsigned char* __restrict__ skip )
{
  size_t slice = slice0 + ((blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (slice >= slice1 || p >= npoin) return;
  // nodes shared with other partitions are updated by k_rhs_finish from the complete sums (and,
  // for LaxCG, from the still unmodified primitives of this stage)
  if (FUSED && skip && skip[p]) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[NC];
  rhs_sum( p, lane, base, kmax, inc_e, F, nslot, bslot, Rb, S, src_mask, v, acc );
  if (FUSED) {
    node_update< LAX >( p, NP, acc, vol[p], Un, U, W, Wn, UnOut, A );
  } else {
    #pragma unroll
    for (int c=0; c<NC; ++c) R[p*NC+c] = acc[c];
  }
}

__global__ void k_rhs_shared( int nsh, const int* __restrict__ sh_node,
            const long long* __restrict__ sl_base, const int* __restrict__ inc_e,
            const double* __restrict__ F, size_t nslot, const int* __restrict__ bslot,
            const double* __restrict__ Rb, const double* __restrict__ S, int src_mask,
            const double* __restrict__ v, double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  size_t slice = p >> 5; int lane = p & 31;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[NC];
  rhs_sum( p, lane, base, kmax, inc_e, F, nslot, bslot, Rb, S, src_mask, v, acc );
  for (int c=0; c<NC; ++c) part[(size_t)i*NC+c] = acc[c];
}

template< bool FUSED >
__global__ void k_rhs_finish( int nsh, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
            const int* __restrict__ ridx, const double* __restrict__ part,
            const double* __restrict__ recvbuf, const double* __restrict__ vol,
            const double* __restrict__ Un, StageArgs A, double* __restrict__ U,
            double* __restrict__ W, double* __restrict__ R, double* __restrict__ Wn, double* __restrict__ UnOut )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  double acc[NC];
  for (int c=0; c<NC; ++c) {
    double a = part[(size_t)i*NC+c];
    for (int r=roff[i]; r<roff[i+1]; ++r) a += recvbuf[(size_t)ridx[r]*NC+c];
    acc[c] = a;
  }
  if (FUSED) {
    if (A.M.rgas > 0.0) node_update< true >( p, NP, acc, vol[p], Un, U, W, Wn, UnOut, A );
    else node_update< false >( p, NP, acc, vol[p], Un, U, W, Wn, UnOut, A );
  } else {
    for (int c=0; c<NC; ++c) R[p*NC+c] = acc[c];
  }
}

// unfused RK update from a materialised R (drop-in for RieCG::solve :1016-1021)
__global__ void k_update( size_t npoin, size_t NP, const double* __restrict__ R, const double* __restrict__ vol,
                          const double* __restrict__ Un, StageArgs A, double* __restrict__ U,
                          double* __restrict__ W, double* __restrict__ Wn, double* __restrict__ UnOut )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  double acc[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) acc[c] = R[p*NC+c];
  if (A.M.rgas > 0.0) node_update< true >( p, NP, acc, vol[p], Un, U, W, Wn, UnOut, A );
  else node_update< false >( p, NP, acc, vol[p], Un, U, W, Wn, UnOut, A );
}

// ---------------------------------------------------------------------------------
// boundary conditions, BC.cpp:29-241, one thread per BC node applying dirbc, symbc,
// farbc, prebc in the reference's order, then refreshing the primitive variables
// ---------------------------------------------------------------------------------
struct FarState { double r, p, u, v, w; };

__global__ void k_bc( int nbc, size_t NP, const int* __restrict__ node, const int* __restrict__ dir,
                      const int* __restrict__ dir_mask, const double* __restrict__ dir_val,
                      const int* __restrict__ symoff, const double* __restrict__ sym_n,
                      const int* __restrict__ faroff, const double* __restrict__ far_n, FarState fs,
                      const int* __restrict__ pre, const double* __restrict__ pre_val,
                      double gamma, double* __restrict__ U, double* __restrict__ W, Mode M )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nbc) return;
  size_t p = node[i];
  double u[NC];
  for (int c=0; c<NC; ++c) u[c] = U[c*NP+p];
  int d = dir[i];
  if (d >= 0) for (int c=0; c<NC; ++c) if (dir_mask[d*NC+c] == 1) u[c] = dir_val[d*NC+c];
  for (int s=symoff[i]; s<symoff[i+1]; ++s) {                 // symbc, BC.cpp:110-136
    const double* n = sym_n + (size_t)s*3;
    double vn = u[1]*n[0] + u[2]*n[1] + u[3]*n[2];
    u[1] -= vn * n[0];
    u[2] -= vn * n[1];
    u[3] -= vn * n[2];
  }
  for (int s=faroff[i]; s<faroff[i+1]; ++s) {                 // farbc, BC.cpp:152-220
    const double* n = far_n + (size_t)s*3;
    double vn = fs.u*n[0] + fs.v*n[1] + fs.w*n[2];
    double a = sqrt( gamma * fs.p / fs.r );
    double M = vn / a;
    if (M <= -1.0) {
      u[0] = fs.r; u[1] = fs.r*fs.u; u[2] = fs.r*fs.v; u[3] = fs.r*fs.w;
      u[4] = fs.p/(gamma-1.0) + 0.5*fs.r*(fs.u*fs.u + fs.v*fs.v + fs.w*fs.w);
    } else if (M > -1.0 && M < 0.0) {
      double pr = (u[4] - 0.5*(u[1]*u[1] + u[2]*u[2] + u[3]*u[3])/u[0]) * (gamma-1.0);
      u[0] = fs.r; u[1] = fs.r*fs.u; u[2] = fs.r*fs.v; u[3] = fs.r*fs.w;
      u[4] = pr/(gamma-1.0) + 0.5*fs.r*(fs.u*fs.u + fs.v*fs.v + fs.w*fs.w);
    } else if (M >= 0.0 && M < 1.0) {
      double uu = u[1]/u[0], vv = u[2]/u[0], ww = u[3]/u[0];
      u[4] = fs.p/(gamma-1.0) + 0.5*u[0]*(uu*uu + vv*vv + ww*ww);
    }
  }
  int pb = pre[i];
  if (pb >= 0) {                                              // prebc, BC.cpp:222-241
    u[0] = pre_val[pb*2+0];
    double uu = u[1]/u[0], vv = u[2]/u[0], ww = u[3]/u[0];
    u[4] = pre_val[pb*2+1]/(gamma-1.0) + 0.5*u[0]*(uu*uu + vv*vv + ww*ww);
  }
  double w[NC];
  for (int c=0; c<NC; ++c) U[c*NP+p] = u[c];
  primitive_of( u, w, M );
  store_w( W, NP, p, w );
}

// ---------------------------------------------------------------------------------
// reductions: time step (RieCG.cpp:827-839) and diagnostics (NodeDiagnostics.cpp:85-118)
// two-pass, fixed tree => deterministic
// ---------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 1184;   // 8 x 148 SMs
constexpr int RED_THREADS = 256;

template< int NV, bool MIN >
__device__ __forceinline__ void block_reduce( double v[NV], double* __restrict__ out )
{
  __shared__ double sm[NV][RED_THREADS/32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  #pragma unroll
  for (int i=0; i<NV; ++i) {
    double a = v[i];
    #pragma unroll
    for (int o=16; o>0; o>>=1) { double b = __shfl_xor_sync( 0xffffffffu, a, o ); a = MIN ? fmin(a,b) : a + b; }
    if (lane == 0) sm[i][w] = a;
  }
  __syncthreads();
  if (w == 0) {
    #pragma unroll
    for (int i=0; i<NV; ++i) {
      double a = lane < RED_THREADS/32 ? sm[i][lane] : (MIN ? 1.7976931348623157e308 : 0.0);
      #pragma unroll
      for (int o=16; o>0; o>>=1) { double b = __shfl_xor_sync( 0xffffffffu, a, o ); a = MIN ? fmin(a,b) : a + b; }
      if (lane == 0) out[(size_t)blockIdx.x*NV+i] = a;
    }
  }
}

__global__ void __launch_bounds__(RED_THREADS)
k_dt( size_t npoin, size_t NP, const double* __restrict__ U, const double* __restrict__ vol, double gamma,
      double* __restrict__ part, Mode M, double cfl, double* __restrict__ dtp )
{
  double m[1] = { 1.7976931348623157e308 };
  const size_t stride = (size_t)gridDim.x*blockDim.x;
  for (size_t p0 = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p0 < npoin; p0 += 4*stride) {
    double a[4][6];
    #pragma unroll
    for (int k=0; k<4; ++k) {               // 24 independent loads in flight per thread
      size_t p = min( p0 + k*stride, npoin-1 );
      #pragma unroll
      for (int c=0; c<NC; ++c) a[k][c] = U[c*NP+p];
      a[k][5] = vol[p];
    }
    #pragma unroll
    for (int k=0; k<4; ++k) {
      double r = a[k][0], u = a[k][1]/r, v = a[k][2]/r, w = a[k][3]/r;
      double L = cbrt( a[k][5] );
      double e;
      if (M.rgas > 0.0) {                  // LaxCG::charvel, LaxCG.cpp:228-259
        double cp = gamma*M.rgas/(gamma-1.0);
        double kk = u*u + v*v + w*w;
        double ei = a[k][4]/r - kk/2.0;
        double pr = (r*ei) * (gamma-1.0);
        double T = pr/r/M.rgas;
        double rp = r/pr;
        double rt = -r/T;
        double vel = sqrt( kk );
        double vr = lax_refvel( r, pr, vel, gamma, M.kvinf );
        double vr2 = vr*vr;
        double beta = rp + rt/r/cp;
        double alpha = 0.5*(1.0 - beta*vr2);
        double vpri = vel*(1.0 - alpha);
        double cpri = sqrt( alpha*alpha*kk + vr2 );
        e = L / fmax( fabs(vpri) + cpri, 1.0e-8 );
      } else {
        double pr = (a[k][4] - 0.5*r*(u*u + v*v + w*w)) * (gamma-1.0);
        double c = sqrt( gamma * fmax(pr,0.0) / r );
        double vel = sqrt( u*u + v*v + w*w );
        e = L / fmax( vel+c, 1.0e-8 );
      }
      if (dtp && p0 + k*stride < npoin) dtp[p0 + k*stride] = e * cfl;   // local time step (steady)
      m[0] = fmin( m[0], e );
    }
  }
  block_reduce< 1, true >( m, part );
}

template< int NV, bool MIN >
__global__ void __launch_bounds__(RED_THREADS)
k_reduce_final( int nblocks, const double* __restrict__ part, double* __restrict__ out )
{
  double a[NV];
  #pragma unroll
  for (int i=0; i<NV; ++i) a[i] = MIN ? 1.7976931348623157e308 : 0.0;
  for (int b=threadIdx.x; b<nblocks; b+=blockDim.x) {
    #pragma unroll
    for (int i=0; i<NV; ++i) { double x = part[(size_t)b*NV+i]; a[i] = MIN ? fmin(a[i],x) : a[i] + x; }
  }
  block_reduce< NV, MIN >( a, out );
}

constexpr int NDIAG = 4*NC+1;
__global__ void __launch_bounds__(RED_THREADS)
k_diag( size_t npoin, size_t NP, const double* __restrict__ U, const double* __restrict__ Un,
        const double* __restrict__ v, const double* __restrict__ an, double* __restrict__ part )
{
  double a[NDIAG];
  #pragma unroll
  for (int i=0; i<NDIAG; ++i) a[i] = 0.0;
  for (size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p < npoin; p += (size_t)gridDim.x*blockDim.x) {
    double vp = v[p], u[NC];
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      u[c] = U[c*NP+p];
      double d = u[c] - Un[c*NP+p];
      a[c] += u[c]*u[c]*vp;
      a[NC+c] += d*d*vp;
    }
    a[2*NC] += u[4]*vp;
    if (an) {
      double w[NC];
      primitive( u, w );
      #pragma unroll
      for (int c=0; c<NC; ++c) {
        double du = w[c] - an[p*NC+c];
        a[2*NC+1+c] += du*du*vp;
        a[3*NC+1+c] += fabs(du)*vp;
      }
    }
  }
  block_reduce< NDIAG, false >( a, part );
}

// ---------------------------------------------------------------------------------
// ZalCG: Taylor-Galerkin two-step edge flux (Zalesak.cpp:31-200, no source term) and the
// flux-corrected-transport passes of ZalCG.cpp:1056-1607 as node gathers
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_zal_flux_edge( size_t nslot, size_t NP, const int* __restrict__ ep, const int* __restrict__ eq,
                 const double* __restrict__ D, const double* __restrict__ U, const double* __restrict__ X,
                 double dt, DParams P, double* __restrict__ F )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= nslot) return;
  int pi = ep[e];
  if (pi < 0) return;
  size_t p = pi, q = eq[e];
  double g = P.gamma;
  double dx = X[p] - X[q], dy = X[NP+p] - X[NP+q], dz = X[2*NP+p] - X[2*NP+q];
  double dl = dx*dx + dy*dy + dz*dz;
  dx /= dl; dy /= dl; dz /= dl;
  double rL = U[p], ruL = U[NP+p], rvL = U[2*NP+p], rwL = U[3*NP+p], reL = U[4*NP+p];
  double pL = (reL - 0.5*(ruL*ruL + rvL*rvL + rwL*rwL)/rL) * (g-1.0);
  double dnL = (ruL*dx + rvL*dy + rwL*dz)/rL;
  double rR = U[q], ruR = U[NP+q], rvR = U[2*NP+q], rwR = U[3*NP+q], reR = U[4*NP+q];
  double pR = (reR - 0.5*(ruR*ruR + rvR*rvR + rwR*rwR)/rR) * (g-1.0);
  double dnR = (ruR*dx + rvR*dy + rwR*dz)/rR;
  double nx = D[e], ny = D[nslot+e], nz = D[2*nslot+e];
  double dp = pL - pR;
  double rh  = 0.5*(rL + rR - dt*(rL*dnL - rR*dnR));
  double ruh = 0.5*(ruL + ruR - dt*(ruL*dnL - ruR*dnR + dp*dx));
  double rvh = 0.5*(rvL + rvR - dt*(rvL*dnL - rvR*dnR + dp*dy));
  double rwh = 0.5*(rwL + rwR - dt*(rwL*dnL - rwR*dnR + dp*dz));
  double reh = 0.5*(reL + reR - dt*((reL+pL)*dnL - (reR+pR)*dnR));
  double ph = (reh - 0.5*(ruh*ruh + rvh*rvh + rwh*rwh)/rh) * (g-1.0);
  double vn = (ruh*nx + rvh*ny + rwh*nz)/rh;
  double f[NC];
  f[0] = 2.0*rh*vn;
  f[1] = 2.0*(ruh*vn + ph*nx);
  f[2] = 2.0*(rvh*vn + ph*ny);
  f[3] = 2.0*(rwh*vn + ph*nz);
  f[4] = 2.0*(reh + ph)*vn;
  if (P.stab2) {
    double vnL = (ruL*nx + rvL*ny + rwL*nz)/rL;
    double vnR = (ruR*nx + rvR*ny + rwR*nz)/rR;
    double len = sqrt( nx*nx + ny*ny + nz*nz );
    double cL = sqrt( g * fmax(pL,0.0) / fmax(rL,1.0e-8) );
    double cR = sqrt( g * fmax(pR,0.0) / fmax(rR,1.0e-8) );
    double fw = P.stab2coef * fmax( fabs(vnL) + cL*len, fabs(vnR) + cR*len );
    f[0] -= fw*(rL - rR); f[1] -= fw*(ruL - ruR); f[2] -= fw*(rvL - rvR);
    f[3] -= fw*(rwL - rwR); f[4] -= fw*(reL - reR);
  }
  store_f( F, nslot, e, f );
}

// pass 1 (aec + first half of alw): R = sum +-F + boundary; P+/- from the antidiffusive edge
// contributions aec = -dif*ctau*(u_first - u_second) (ZalCG.cpp:1071-1115), symmetry BC on P
// (:1117-1133), then P /= vol and the low-order solution ul = u - dt R/vol - P+ - P- (:1195-1204)
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_zal_node1( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot,
             const double* __restrict__ F, const double* __restrict__ U, const int* __restrict__ bslot,
             const double* __restrict__ Rb, const int* __restrict__ bcof, const int* __restrict__ symoff,
             const double* __restrict__ sym_n, const double* __restrict__ vol, double dt, double ctau, int fct,
             double* __restrict__ P, double* __restrict__ UL, double* __restrict__ R )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double up[NC], r[NC], pp[NC], pn[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { up[c] = U[c*NP+p]; r[c] = 0.0; pp[c] = 0.0; pn[c] = 0.0; }
  // padding entries (se = 0) point at the node itself and at slot 0 with weight 0: no branch
  #pragma unroll kZalUnroll
  for (int k=0; k<kmax; ++k) {
    long long i = base + (long long)k*32 + lane;
    int2 eq = __ldg( inc_eq + i );
    const int se = eq.x;
    size_t q = (size_t)eq.y;
    size_t sl = se == 0 ? 0 : (size_t)(abs(se)-1);
    double dif = se == 0 ? 0.0 : __ldg( D + 3*nslot + sl );
    double fl[NC];
    load_f( F, nslot, sl, fl );
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double f = se == 0 ? 0.0 : fl[c];
      double uq = __ldg( U + c*NP + q );
      if (se < 0) {                       // this node is the edge's first node
        r[c] -= f;
        double aec = -dif * ctau * (up[c] - uq);
        if (aec > 0.0) pn[c] -= aec; else pp[c] -= aec;
      } else {                            // second node
        r[c] += f;
        double aec = -dif * ctau * (uq - up[c]);
        if (aec > 0.0) pp[c] += aec; else pn[c] += aec;
      }
    }
  }
  int b = bslot[p];
  if (b >= 0) {
    #pragma unroll
    for (int c=0; c<NC; ++c) r[c] += Rb[(size_t)b*NC+c];
  }
  if (!fct) {
    #pragma unroll
    for (int c=0; c<NC; ++c) R[p*NC+c] = r[c];
    return;
  }
  int bc = bcof[p];
  if (bc >= 0)
    for (int s=symoff[bc]; s<symoff[bc+1]; ++s) {
      const double* n = sym_n + (size_t)s*3;
      double rvnp = pp[1]*n[0] + pp[2]*n[1] + pp[3]*n[2];
      double rvnn = pn[1]*n[0] + pn[2]*n[1] + pn[3]*n[2];
      pp[1] -= rvnp * n[0]; pn[1] -= rvnn * n[0];
      pp[2] -= rvnp * n[1]; pn[2] -= rvnn * n[1];
      pp[3] -= rvnp * n[2]; pn[3] -= rvnn * n[2];
    }
  double vp = vol[p];
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    pp[c] /= vp; pn[c] /= vp;
    P[(2*c)*NP+p] = pp[c]; P[(2*c+1)*NP+p] = pn[c];
    UL[c*NP+p] = up[c] - dt*r[c]/vp - pp[c] - pn[c];
    R[p*NC+c] = r[c];
  }
}

// pass 2 (second half of alw + first half of lim): allowed bounds Q+/- over the edge
// neighbours (:1206-1290), Q -= ul, limit coefficients C+/- (:1361-1380) -> Q
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_zal_node2( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq, const double* __restrict__ U, const double* __restrict__ UL,
             const double* __restrict__ P, int clip, double* __restrict__ Q )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double hp[NC], lp[NC], qa[NC], qb[NC], ulp[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    ulp[c] = UL[c*NP+p]; double u = U[c*NP+p];
    hp[c] = clip ? ulp[c] : fmax( ulp[c], u );
    lp[c] = clip ? ulp[c] : fmin( ulp[c], u );
    qa[c] = -1.7976931348623157e308; qb[c] = 1.7976931348623157e308;
  }
  #pragma unroll kZalUnroll
  for (int k=0; k<kmax; ++k) {                 // padding entries point at the node itself: no effect on the bounds
    long long i = base + (long long)k*32 + lane;
    size_t q = (size_t)__ldg( inc_eq + i ).y;
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double ulq = __ldg( UL + c*NP + q );
      double hq = ulq, lq = ulq;
      if (!clip) { double uq = __ldg( U + c*NP + q ); hq = fmax( ulq, uq ); lq = fmin( ulq, uq ); }
      qa[c] = fmax( qa[c], fmax( hp[c], hq ) );
      qb[c] = fmin( qb[c], fmin( lp[c], lq ) );
    }
  }
  const double eps = 2.220446049250313e-16;
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    double a = qa[c] - ulp[c], b = qb[c] - ulp[c];
    double pa = P[(2*c)*NP+p], pb = P[(2*c+1)*NP+p];
    Q[(2*c)*NP+p]   = pa <  eps ? 0.0 : fmin( 1.0, a/pa );
    Q[(2*c+1)*NP+p] = pb > -eps ? 0.0 : fmin( 1.0, b/pb );
  }
}

// pass 3 (second half of lim + solve): limited antidiffusive contributions (:1382-1481) and
// u = ul + a/vol (:1552-1557)
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_zal_node3( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot,
             const double* __restrict__ U, const double* __restrict__ UL, const double* __restrict__ Q,
             const double* __restrict__ vol, double ctau, int sysmask, double* __restrict__ Unew,
             double* __restrict__ W )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double up[NC], cpa[NC], cpb[NC], a[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { up[c] = U[c*NP+p]; cpa[c] = Q[(2*c)*NP+p]; cpb[c] = Q[(2*c+1)*NP+p]; a[c] = 0.0; }
  #pragma unroll kZalUnroll
  for (int k=0; k<kmax; ++k) {                 // padding entries: dif = 0, neighbour = the node itself
    long long i = base + (long long)k*32 + lane;
    int2 eq = __ldg( inc_eq + i );
    const int se = eq.x;
    size_t q = (size_t)eq.y;
    double dif = se == 0 ? 0.0 : __ldg( D + 3*nslot + (size_t)(abs(se)-1) );
    double aec[NC], coef[NC];
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double uq = __ldg( U + c*NP + q );
      double cqa = __ldg( Q + (2*c)*NP + q ), cqb = __ldg( Q + (2*c+1)*NP + q );
      if (se < 0) {      // first = this node, second = q
        aec[c] = -dif * ctau * (up[c] - uq);
        coef[c] = fmin( aec[c] < 0.0 ? cpa[c] : cpb[c], aec[c] > 0.0 ? cqa : cqb );
      } else {           // first = q, second = this node
        aec[c] = -dif * ctau * (uq - up[c]);
        coef[c] = fmin( aec[c] < 0.0 ? cqa : cqb, aec[c] > 0.0 ? cpa[c] : cpb[c] );
      }
    }
    double cs = 1.0;
    #pragma unroll
    for (int c=0; c<NC; ++c) if (sysmask & (1<<c)) cs = fmin( cs, coef[c] );
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      if (sysmask & (1<<c)) coef[c] = cs;
      double v = aec[c] * coef[c];
      if (se < 0) a[c] -= v; else a[c] += v;
    }
  }
  double vp = vol[p], u[NC], w[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { u[c] = UL[c*NP+p] + a[c]/vp; Unew[c*NP+p] = u[c]; }
  primitive( u, w );
  store_w( W, NP, p, w );
}

// fct = false: u = u - dt R/vol (ZalCG.cpp:1560-1567)
__global__ void k_zal_nofct( size_t npoin, size_t NP, const double* __restrict__ R, const double* __restrict__ vol,
                             const double* __restrict__ U, double dt, double* __restrict__ Unew, double* __restrict__ W )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  double vp = vol[p], u[NC], w[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { u[c] = U[c*NP+p] - dt*R[p*NC+c]/vp; Unew[c*NP+p] = u[c]; }
  primitive( u, w );
  store_w( W, NP, p, w );
}

// ---------------------------------------------------------------------------------
// KozCG: element-based Taylor-Galerkin (Kozak.cpp:29-180) and FCT (KozCG.cpp:774-1197).
// Each pass = one kernel over tetrahedra writing per-tet, per-local-node values, and one
// node gather over the incident tetrahedra (fixed order, no atomics).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double koz_geom( const double* __restrict__ X, size_t NP, const int N[4], double grad[4][3] )
{
  double x[4][3];
  #pragma unroll
  for (int a=0; a<4; ++a) { x[a][0] = X[N[a]]; x[a][1] = X[NP+N[a]]; x[a][2] = X[2*NP+N[a]]; }
  double ba[3] = { x[1][0]-x[0][0], x[1][1]-x[0][1], x[1][2]-x[0][2] },
         ca[3] = { x[2][0]-x[0][0], x[2][1]-x[0][1], x[2][2]-x[0][2] },
         da[3] = { x[3][0]-x[0][0], x[3][1]-x[0][1], x[3][2]-x[0][2] };
  grad[1][0] = ca[1]*da[2] - da[1]*ca[2]; grad[1][1] = ca[2]*da[0] - da[2]*ca[0]; grad[1][2] = ca[0]*da[1] - da[0]*ca[1];
  grad[2][0] = da[1]*ba[2] - ba[1]*da[2]; grad[2][1] = da[2]*ba[0] - ba[2]*da[0]; grad[2][2] = da[0]*ba[1] - ba[0]*da[1];
  grad[3][0] = ba[1]*ca[2] - ca[1]*ba[2]; grad[3][1] = ba[2]*ca[0] - ca[2]*ba[0]; grad[3][2] = ba[0]*ca[1] - ca[0]*ba[1];
  #pragma unroll
  for (int i=0; i<3; ++i) grad[0][i] = -grad[1][i]-grad[2][i]-grad[3][i];
  return ba[0]*grad[1][0] + ba[1]*grad[1][1] + ba[2]*grad[1][2];          // triple(ba,ca,da)
}

// pass 1: rhs contributions T[a*5+c] and antidiffusive element contributions T[20+c*4+a]
__global__ void __launch_bounds__(128)
k_koz_elem1( size_t ntet, size_t NP, const int* __restrict__ tet, const double* __restrict__ U,
             const double* __restrict__ X, const double* __restrict__ S, const double* __restrict__ Sc,
             double dt, double gamma, double ctau, int fct, double* __restrict__ T )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= ntet) return;
  int N[4] = { tet[e], tet[ntet+e], tet[2*ntet+e], tet[3*ntet+e] };
  double grad[4][3];
  double J = koz_geom( X, NP, N, grad );
  double u[4][NC], p[4];
  #pragma unroll
  for (int a=0; a<4; ++a) {
    #pragma unroll
    for (int c=0; c<NC; ++c) u[a][c] = U[c*NP+N[a]];
    p[a] = (u[a][4] - 0.5*(u[a][1]*u[a][1] + u[a][2]*u[a][2] + u[a][3]*u[a][3])/u[a][0]) * (gamma-1.0);
  }
  double ue[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) ue[c] = (u[0][c] + u[1][c] + u[2][c] + u[3][c])/4.0;
  double coef = dt/J/2.0;
  #pragma unroll
  for (int j=0; j<3; ++j)
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double cg = coef * grad[a][j];
      double uj = u[a][j+1] / u[a][0];
      ue[0] -= cg * u[a][j+1];
      ue[1] -= cg * u[a][1] * uj;
      ue[2] -= cg * u[a][2] * uj;
      ue[3] -= cg * u[a][3] * uj;
      ue[j+1] -= cg * p[a];
      ue[4] -= cg * (u[a][4] + p[a]) * uj;
    }
  if (S) {
    coef = dt/8.0;
    #pragma unroll
    for (int a=0; a<4; ++a)
      #pragma unroll
      for (int c=0; c<NC; ++c) ue[c] += coef * S[(size_t)N[a]*NC+c];
  }
  double pr = (ue[4] - 0.5*(ue[1]*ue[1] + ue[2]*ue[2] + ue[3]*ue[3])/ue[0]) * (gamma-1.0);
  double R[4][NC];
  #pragma unroll
  for (int a=0; a<4; ++a)
    #pragma unroll
    for (int c=0; c<NC; ++c) R[a][c] = 0.0;
  coef = 1.0/6.0;
  #pragma unroll
  for (int j=0; j<3; ++j) {
    double uj = ue[j+1] / ue[0];
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double cg = coef * grad[a][j];
      R[a][0] += cg * ue[j+1];
      R[a][1] += cg * ue[1] * uj;
      R[a][2] += cg * ue[2] * uj;
      R[a][3] += cg * ue[3] * uj;
      R[a][j+1] += cg * pr;
      R[a][4] += cg * (ue[4] + pr) * uj;
    }
  }
  if (S) {
    coef = J/24.0;
    #pragma unroll
    for (int a=0; a<4; ++a)
      #pragma unroll
      for (int c=0; c<NC; ++c) R[a][c] += coef * Sc[e*NC+c];
  }
  #pragma unroll
  for (int a=0; a<4; ++a)
    #pragma unroll
    for (int c=0; c<NC; ++c) T[(size_t)(a*NC+c)*ntet+e] = R[a][c];
  if (fct) {
    #pragma unroll
    for (int c=0; c<NC; ++c)
      #pragma unroll
      for (int a=0; a<4; ++a) {
        double aec = 0.0;
        #pragma unroll
        for (int b=0; b<4; ++b) { double m = J/120.0 * ((a == b) ? 3.0 : -1.0); aec += m * ctau * u[b][c]; }
        T[(size_t)(20+c*4+a)*ntet+e] = aec;
      }
  }
}

// node pass 1: R, P+/-, symmetry BC on P, low-order solution ul = u + dt R/vol - P+ - P-
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_koz_node1( size_t npoin, size_t NP, size_t ntet, const long long* __restrict__ kbase, const int* __restrict__ kinc,
             const double* __restrict__ T, const double* __restrict__ U, const int* __restrict__ bcof,
             const int* __restrict__ symoff, const double* __restrict__ sym_n, const double* __restrict__ vol,
             double dt, int fct, double* __restrict__ P, double* __restrict__ UL, double* __restrict__ R )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = kbase[slice];
  int kmax = (int)((kbase[slice+1] - base) >> 5);
  double r[NC], pp[NC], pn[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { r[c] = 0.0; pp[c] = 0.0; pn[c] = 0.0; }
  for (int k=0; k<kmax; ++k) {
    int ta = __ldg( kinc + base + (long long)k*32 + lane );
    if (ta < 0) continue;
    size_t e = (size_t)(ta >> 2); int a = ta & 3;
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      r[c] += __ldg( T + (size_t)(a*NC+c)*ntet + e );
      if (fct) { double aec = __ldg( T + (size_t)(20+c*4+a)*ntet + e ); pp[c] += fmax( 0.0, aec ); pn[c] += fmin( 0.0, aec ); }
    }
  }
  #pragma unroll
  for (int c=0; c<NC; ++c) R[p*NC+c] = r[c];
  if (!fct) return;
  int bc = bcof[p];
  if (bc >= 0)
    for (int s=symoff[bc]; s<symoff[bc+1]; ++s) {
      const double* n = sym_n + (size_t)s*3;
      double rvnp = pp[1]*n[0] + pp[2]*n[1] + pp[3]*n[2];
      double rvnn = pn[1]*n[0] + pn[2]*n[1] + pn[3]*n[2];
      pp[1] -= rvnp * n[0]; pn[1] -= rvnn * n[0];
      pp[2] -= rvnp * n[1]; pn[2] -= rvnn * n[1];
      pp[3] -= rvnp * n[2]; pn[3] -= rvnn * n[2];
    }
  double vp = vol[p];
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    pp[c] /= vp; pn[c] /= vp;
    P[(2*c)*NP+p] = pp[c]; P[(2*c+1)*NP+p] = pn[c];
    UL[c*NP+p] = U[c*NP+p] + dt*r[c]/vp - pp[c] - pn[c];
  }
}

// pass 2: per-tet allowed bounds over its 4 nodes -> T[c*2], T[c*2+1]
__global__ void __launch_bounds__(128)
k_koz_elem2( size_t ntet, size_t NP, const int* __restrict__ tet, const double* __restrict__ U,
             const double* __restrict__ UL, int clip, double* __restrict__ T )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= ntet) return;
  int N[4] = { tet[e], tet[ntet+e], tet[2*ntet+e], tet[3*ntet+e] };
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    double alwp = -1.7976931348623157e308, alwn = 1.7976931348623157e308;
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double ul = UL[c*NP+N[a]];
      if (clip) { alwp = fmax( alwp, ul ); alwn = fmin( alwn, ul ); }
      else { double u = U[c*NP+N[a]]; alwp = fmax( alwp, fmax( ul, u ) ); alwn = fmin( alwn, fmin( ul, u ) ); }
    }
    T[(size_t)(2*c)*ntet+e] = alwp; T[(size_t)(2*c+1)*ntet+e] = alwn;
  }
}

// node pass 2: Q+/- = max/min over incident tets, minus ul, limit coefficients C+/-
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_koz_node2( size_t npoin, size_t NP, size_t ntet, const long long* __restrict__ kbase, const int* __restrict__ kinc,
             const double* __restrict__ T, const double* __restrict__ UL, const double* __restrict__ P,
             double* __restrict__ Q )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = kbase[slice];
  int kmax = (int)((kbase[slice+1] - base) >> 5);
  double qa[NC], qb[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { qa[c] = -1.7976931348623157e308; qb[c] = 1.7976931348623157e308; }
  for (int k=0; k<kmax; ++k) {
    int ta = __ldg( kinc + base + (long long)k*32 + lane );
    if (ta < 0) continue;
    size_t e = (size_t)(ta >> 2);
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      qa[c] = fmax( qa[c], __ldg( T + (size_t)(2*c)*ntet + e ) );
      qb[c] = fmin( qb[c], __ldg( T + (size_t)(2*c+1)*ntet + e ) );
    }
  }
  const double eps = 2.220446049250313e-16;
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    double ul = UL[c*NP+p];
    double a = qa[c] - ul, b = qb[c] - ul;
    double pa = P[(2*c)*NP+p], pb = P[(2*c+1)*NP+p];
    Q[(2*c)*NP+p]   = pa <  eps ? 0.0 : fmin( 1.0, a/pa );
    Q[(2*c+1)*NP+p] = pb > -eps ? 0.0 : fmin( 1.0, b/pb );
  }
}

// pass 3: limited antidiffusive element contributions coef[c]*aec[c][a] -> T[a*5+c]
__global__ void __launch_bounds__(128)
k_koz_elem3( size_t ntet, size_t NP, const int* __restrict__ tet, const double* __restrict__ U,
             const double* __restrict__ X, const double* __restrict__ Q, double ctau, int sysmask,
             double* __restrict__ T )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= ntet) return;
  int N[4] = { tet[e], tet[ntet+e], tet[2*ntet+e], tet[3*ntet+e] };
  double grad[4][3];
  double J = koz_geom( X, NP, N, grad );
  double coef[NC], aec[NC][4];
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    double u[4];
    #pragma unroll
    for (int a=0; a<4; ++a) u[a] = U[c*NP+N[a]];
    coef[c] = 1.0;
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double v = 0.0;
      #pragma unroll
      for (int b=0; b<4; ++b) { double m = J/120.0 * ((a == b) ? 3.0 : -1.0); v += m * ctau * u[b]; }
      aec[c][a] = v;
      coef[c] = fmin( coef[c], v > 0.0 ? Q[(2*c)*NP+N[a]] : Q[(2*c+1)*NP+N[a]] );
    }
  }
  double cs = 1.0;
  #pragma unroll
  for (int c=0; c<NC; ++c) if (sysmask & (1<<c)) cs = fmin( cs, coef[c] );
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    if (sysmask & (1<<c)) coef[c] = cs;
    #pragma unroll
    for (int a=0; a<4; ++a) T[(size_t)(a*NC+c)*ntet+e] = coef[c] * aec[c][a];
  }
}

// node pass 3: a = sum of limited contributions, u = ul + a/vol (KozCG.cpp:1140-1146)
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_koz_node3( size_t npoin, size_t NP, size_t ntet, const long long* __restrict__ kbase, const int* __restrict__ kinc,
             const double* __restrict__ T, const double* __restrict__ UL, const double* __restrict__ vol,
             double* __restrict__ Unew, double* __restrict__ W )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = kbase[slice];
  int kmax = (int)((kbase[slice+1] - base) >> 5);
  double a_[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) a_[c] = 0.0;
  for (int k=0; k<kmax; ++k) {
    int ta = __ldg( kinc + base + (long long)k*32 + lane );
    if (ta < 0) continue;
    size_t e = (size_t)(ta >> 2); int a = ta & 3;
    #pragma unroll
    for (int c=0; c<NC; ++c) a_[c] += __ldg( T + (size_t)(a*NC+c)*ntet + e );
  }
  double vp = vol[p], u[NC], w[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { u[c] = UL[c*NP+p] + a_[c]/vp; Unew[c*NP+p] = u[c]; }
  primitive( u, w );
  store_w( W, NP, p, w );
}

// fct = false: u = u + dt R/vol (KozCG.cpp:1150-1157)
__global__ void k_koz_nofct( size_t npoin, size_t NP, const double* __restrict__ R, const double* __restrict__ vol,
                             const double* __restrict__ U, double dt, double* __restrict__ Unew, double* __restrict__ W )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  double vp = vol[p], u[NC], w[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { u[c] = U[c*NP+p] + dt*R[p*NC+c]/vp; Unew[c*NP+p] = u[c]; }
  primitive( u, w );
  store_w( W, NP, p, w );
}

// ---------------------------------------------------------------------------------
// linear solver: CSR::mult (CSR.cpp:154-172) and the vector operations of
// ConjugateGradients.cpp:584-823. Matrix in sliced ELL over scalar rows (one warp per 32
// rows, entry k of lane l at base + 32k + l: coalesced), dots by fixed two-pass trees.
// device scalars: [0] rho [1] rho0 [2] alpha [3] beta [4] normr2 [5] finished
//                 [8..] partial sums handed to the all-reduce
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_spmv( size_t nrow, const long long* __restrict__ base, const int* __restrict__ col,
        const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t r = slice*32 + lane;
  if (r >= nrow) return;
  long long b = base[slice];
  int kmax = (int)((base[slice+1] - b) >> 5);
  double acc = 0.0;
  #pragma unroll 4
  for (int k=0; k<kmax; ++k) {
    long long i = b + (long long)k*32 + lane;
    acc += __ldg( val + i ) * __ldg( x + __ldg( col + i ) );
  }
  y[r] = acc;
}

// The same product with Dirichlet rows/columns applied on the fly: what tk::CSR::dirichlet
// (CSR.cpp:106-152) leaves in the matrix -- BC columns zero, BC rows the unit row (diag =
// 1/count over the sharing partitions) -- without touching (or having to restore,
// ConjugateGradients.cpp:809) the stored values.
__global__ void __launch_bounds__(256)
k_spmv_bc( size_t nrow, const long long* __restrict__ base, const int* __restrict__ col,
           const double* __restrict__ val, const unsigned char* __restrict__ bc,
           const double* __restrict__ cnt, const double* __restrict__ x, double* __restrict__ y )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t r = slice*32 + lane;
  if (r >= nrow) return;
  if (bc[r]) { y[r] = (1.0 / cnt[r]) * x[r]; return; }
  long long b = base[slice];
  int kmax = (int)((base[slice+1] - b) >> 5);
  double acc = 0.0;
  #pragma unroll 4
  for (int k=0; k<kmax; ++k) {
    long long i = b + (long long)k*32 + lane;
    int cl = __ldg( col + i );
    double a = bc[cl] ? 0.0 : __ldg( val + i );
    acc += a * __ldg( x + cl );
  }
  y[r] = acc;
}

// rhs with BCs (ConjugateGradients::apply/r :451-556): b += neumann; b -= A(:,bc) val; b(bc) = val
// (own part r of the column sums: shared rows are summed over the partitions by the caller)
__global__ void __launch_bounds__(256)
k_cg_bc_colsum( size_t nrow, const long long* __restrict__ base, const int* __restrict__ col,
                const double* __restrict__ val, const unsigned char* __restrict__ bc,
                const double* __restrict__ bcval, double* __restrict__ rsum )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t r = slice*32 + lane;
  if (r >= nrow) return;
  long long b = base[slice];
  int kmax = (int)((base[slice+1] - b) >> 5);
  double acc = 0.0;
  for (int k=kmax-1; k>=0; --k) {            // descending column id, the order BCs are applied in (:481)
    long long i = b + (long long)k*32 + lane;
    int cl = __ldg( col + i );
    if (bc[cl]) acc += __ldg( val + i ) * bcval[cl];
  }
  rsum[r] = acc;
}
__global__ void k_cg_bc_rhs( size_t nrow, const unsigned char* __restrict__ bc, const double* __restrict__ bcval,
                             const double* __restrict__ neu, const double* __restrict__ rsum, double* __restrict__ b )
{
  size_t r = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  double v = b[r];
  if (neu) v += neu[r];
  v -= rsum[r];
  b[r] = bc[r] ? bcval[r] : v;
}
__global__ void k_cg_bc_diag( size_t nrow, const unsigned char* __restrict__ bc, const double* __restrict__ cnt,
                              double* __restrict__ d )
{
  size_t r = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  if (bc[r]) d[r] = 1.0 / cnt[r];
}

// p = z + beta p    (ConjugateGradients::next :584-599)
__global__ void k_cg_p( size_t n, const double* __restrict__ scal, const double* __restrict__ z, double* __restrict__ p )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i] = z[i] + scal[3] * p[i];
}

// partial masked dot products (ConjugateGradients::dot :128-151): NV pairs at once
template< int NV >
__global__ void __launch_bounds__(RED_THREADS)
k_cg_dot( size_t n, const double* __restrict__ mask, const double* __restrict__ a0, const double* __restrict__ b0,
          const double* __restrict__ a1, const double* __restrict__ b1, double* __restrict__ part )
{
  double s[NV];
  #pragma unroll
  for (int k=0; k<NV; ++k) s[k] = 0.0;
  for (size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x*blockDim.x) {
    double m = mask[i];
    s[0] += m * (a0[i] * b0[i]);
    if (NV > 1) s[1] += m * (a1[i] * b1[i]);
  }
  block_reduce< NV, false >( s, part );
}

// r -= alpha q ; z = r/d ; x += alpha p ; partial (r,z) and (r,r)   (pq :671-704, rz :717-727)
__global__ void __launch_bounds__(RED_THREADS)
k_cg_update( size_t n, const double* __restrict__ scal, const double* __restrict__ mask,
             const double* __restrict__ q, const double* __restrict__ d, const double* __restrict__ p,
             double* __restrict__ r, double* __restrict__ z, double* __restrict__ x, double* __restrict__ part )
{
  double s[2] = { 0.0, 0.0 };
  double alpha = scal[2];
  for (size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x*blockDim.x) {
    double ri = r[i] - alpha * q[i];
    double zi = ri / d[i];
    r[i] = ri; z[i] = zi;
    x[i] += alpha * p[i];
    double m = mask[i];
    s[0] += m * (ri * zi);
    s[1] += m * (ri * ri);
  }
  block_reduce< 2, false >( s, part );
}

// scalar bookkeeping after a reduction; mode 0: alpha = rho/(p,q) (pq :671-690);
// mode 1: rho0 = rho, rho = (r,z), beta = rho/rho0, normr2 = (r,r)  (next :586-588, rz :719)
__global__ void k_cg_scalars( int mode, double* __restrict__ scal )
{
  if (threadIdx.x || blockIdx.x) return;
  if (mode == 0) {
    double d = scal[8];
    if (fabs(d) < 2.220446049250313e-16) { scal[5] = 1.0; scal[2] = 0.0; } else scal[2] = scal[0] / d;
  } else {
    scal[1] = scal[0];
    scal[0] = scal[8];
    scal[3] = scal[0] / scal[1];
    scal[4] = scal[9];
  }
}

// shared rows: sum of the sharers (halo) then optionally divide by the count (x :772-785)
__global__ void k_cg_shared_get( int nsh, int w, const int* __restrict__ sh_node, const double* __restrict__ v, double* __restrict__ part )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nsh*w) return;
  part[i] = v[(size_t)sh_node[i/w]*w + i%w];
}
__global__ void k_cg_shared_put( int nsh, int w, const int* __restrict__ sh_node, const int* __restrict__ roff,
                                 const int* __restrict__ ridx, const double* __restrict__ part,
                                 const double* __restrict__ recvbuf, const double* __restrict__ cnt, int average,
                                 double* __restrict__ v )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nsh*w) return;
  size_t s = i / w, c = i % w;
  double a = part[i];
  for (int r=roff[s]; r<roff[s+1]; ++r) a += recvbuf[(size_t)ridx[r]*w + c];
  size_t row = (size_t)sh_node[s]*w + c;
  v[row] = average ? a / cnt[row] : a;
}
__global__ void k_cg_inv( size_t n, const double* __restrict__ cnt, double* __restrict__ d )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i < n) d[i] = 1.0 / cnt[i];
}
__global__ void k_cg_div( size_t n, const double* __restrict__ r, const double* __restrict__ d, double* __restrict__ z )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i < n) z[i] = r[i] / d[i];
}
__global__ void k_cg_resid( size_t n, const double* __restrict__ b, double* __restrict__ r, double* __restrict__ p )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i < n) { double v = r[i] * -1.0 + b[i]; r[i] = v; p[i] = v; }     // initres :293-298
}

#define XYST_CHOCG_KERNELS
#include "chocg.cuh"
#undef XYST_CHOCG_KERNELS

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
inline unsigned nblk( size_t n, int t ) { return (unsigned)((n + (size_t)t - 1) / (size_t)t); }

struct ProfScope {
  xyst_ctx* c; Prof* pr = nullptr; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope( xyst_ctx* ctx, const char* name ) : c( ctx ) {
    if (!c->prof_on) return;
    pr = &c->prof[name];
    CK( cudaEventCreate( &a ) ); CK( cudaEventCreate( &b ) );
    CK( cudaEventRecord( a, c->stream ) );
  }
  ~ProfScope() {
    if (!pr) return;
    cudaEventRecord( b, c->stream );
    pr->ev.emplace_back( a, b );
  }
};

DParams dparams( const xyst_ctx* c ) {
  return DParams{ c->prm.gamma, c->prm.stab2coef, c->prm.flux, c->prm.stab2, c->prm.exact_muscl, c->rgas, c->kvinf };
}
Mode mode( const xyst_ctx* c ) { return Mode{ c->prm.gamma, c->lax ? c->rgas : 0.0, c->kvinf }; }

void need_mesh( xyst_ctx* c ) { if (!c->npoin) throw std::runtime_error( "no mesh uploaded" ); }

// ---- halo exchange of per-shared-node partial sums (width w doubles) ----------------
// on_comm = false: the partial sums were produced on the compute stream, pack there and hand over
// to the side stream. on_comm = true: the caller produced them on the side stream itself (the
// RieCG sweeps: nothing of the exchange sits in front of the full-mesh gather on the compute stream).
void exchange( xyst_ctx* c, int w, bool on_comm = false )
{
  auto ps = on_comm ? c->comm_stream : c->stream;
  k_pack<<< nblk( c->nsend*(size_t)w, 256 ), 256, 0, ps >>>( (int)c->nsend, w, c->sh_send.p,
    c->sh_part.p, c->sh_sendbuf.p ); ++c->launches;
  if (!on_comm) {
    CK( cudaEventRecord( c->ev_a, c->stream ) );
    CK( cudaStreamWaitEvent( c->comm_stream, c->ev_a, 0 ) );
  }
  NK( g_nccl.GroupStart() );
  for (size_t i=0; i<c->neigh.size(); ++i) {
    size_t off = c->neigh_off[i]*(size_t)w, cnt = (c->neigh_off[i+1]-c->neigh_off[i])*(size_t)w;
    NK( g_nccl.Send( c->sh_sendbuf.p + off, cnt, NCCL_FLOAT64, c->neigh[i], c->comm, c->comm_stream ) );
    NK( g_nccl.Recv( c->sh_recvbuf.p + off, cnt, NCCL_FLOAT64, c->neigh[i], c->comm, c->comm_stream ) );
  }
  NK( g_nccl.GroupEnd() );
  CK( cudaEventRecord( c->ev_b, c->comm_stream ) );
}
void exchange_wait( xyst_ctx* c ) { CK( cudaStreamWaitEvent( c->stream, c->ev_b, 0 ) ); }

void do_grad( xyst_ctx* c )
{
  need_mesh( c );
  auto s = c->stream;
  bool halo = c->nsh > 0 && c->comm;
  // Both boundary kernels depend only on the state at stage start, so they run on a side stream
  // under the gradient gather; boundary nodes are finished afterwards. With several partitions
  // the partial sums of the shared nodes, their packing and the NCCL send/recv run on the
  // communication stream as well, so the compute stream goes straight into the full-mesh gather.
  bool overlap = c->nbn > 0;
  if (overlap || halo) CK( cudaEventRecord( c->ev_c, s ) );
  if (overlap) {
    CK( cudaStreamWaitEvent( c->aux_stream, c->ev_c, 0 ) );
    k_bnd_grad<<< nblk( c->nbn, 128 ), 128, 0, c->aux_stream >>>( (int)c->nbn, c->NP, c->bn_off.p, c->bn_face.p,
      c->tri.p, c->fn.p, c->W.p, c->Gb.p ); ++c->launches;
    CK( cudaEventRecord( c->ev_d, c->aux_stream ) );
    k_bnd_rhs<<< nblk( c->nbn, 128 ), 128, 0, c->aux_stream >>>( (int)c->nbn, c->NP, c->bn_off.p, c->bn_face.p,
      c->tri.p, c->besym.p, c->fn.p, c->U.p, c->Rb.p, c->prm.gamma, c->W.p, c->rgas ); ++c->launches;
    CK( cudaEventRecord( c->ev_e, c->aux_stream ) );
    c->rb_pending = true;
  }
  if (halo) {
    auto cs = c->comm_stream;
    CK( cudaStreamWaitEvent( cs, c->ev_c, 0 ) );
    if (overlap) CK( cudaStreamWaitEvent( cs, c->ev_d, 0 ) );        // needs the boundary part Gb
    k_grad_shared<<< nblk( c->nsh, 128 ), 128, 0, cs >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sl_base.p,
      c->inc_eq.p, c->D2.p, c->D.p, c->nslot, c->W.p, c->bslot.p, c->Gb.p, c->sh_part.p ); ++c->launches;
    exchange( c, 15, true );
  }
  {
    ProfScope ps( c, "grad" );
    k_grad_node<<< nblk( c->nslice*32, NODE_THREADS ), NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->sl_base.p,
      c->inc_eq.p, c->D2.p, c->D.p, c->nslot, c->W.p, c->bslot.p, c->Gb.p, c->vol.p, c->G.p, overlap ? 1 : 0 );
    ++c->launches;
  }
  if (overlap) {
    CK( cudaStreamWaitEvent( s, c->ev_d, 0 ) );
    k_grad_bfix<<< nblk( c->nbn*15, 256 ), 256, 0, s >>>( (int)c->nbn, c->NP, c->bn_node.p, c->Gb.p, c->vol.p, c->G.p ); ++c->launches;
  }
  if (halo) {
    exchange_wait( c );
    k_grad_finish<<< nblk( c->nsh, 128 ), 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p,
      c->sh_ridx.p, c->sh_part.p, c->sh_recvbuf.p, c->vol.p, c->G.p ); ++c->launches;
  }
  CK( cudaGetLastError() );
}

void do_flux( xyst_ctx* c, size_t e0 = 0, size_t e1 = ~(size_t)0 )
{
  auto s = c->stream;
  auto P = dparams( c );
  ProfScope ps( c, "flux" );
  e1 = std::min( e1, c->nslot );
  if (e1 <= e0) return;
  unsigned g = nblk( e1 - e0, FLUX_THREADS );
  #define LAUNCH_FLUX( EX, FL ) k_flux_edge< EX, FL ><<< g, FLUX_THREADS, 0, s >>>( c->nslot, c->NP, c->ep.p, \
      c->eq.p, c->D.p, c->W.p, c->X.p, c->G.p, c->F.p, P, e0, e1 )
  int fl = P.flux + (c->lax ? 2 : 0);
  if (P.exact) { if (fl == 0) LAUNCH_FLUX( true, 0 ); else if (fl == 1) LAUNCH_FLUX( true, 1 );
                 else if (fl == 2) LAUNCH_FLUX( true, 2 ); else LAUNCH_FLUX( true, 3 ); }
  else         { if (fl == 0) LAUNCH_FLUX( false, 0 ); else if (fl == 1) LAUNCH_FLUX( false, 1 );
                 else if (fl == 2) LAUNCH_FLUX( false, 2 ); else LAUNCH_FLUX( false, 3 ); }
  #undef LAUNCH_FLUX
  ++c->launches;
}

static const double rkcoef[3] = { 1.0/3.0, 1.0/2.0, 1.0 };   // RieCG.cpp:41

// the node gather over slices [s0,s1) on stream st
void launch_rhs_node( xyst_ctx* c, bool fused, const StageArgs& A, const double* Un, double* Uout,
                      size_t s0, size_t s1, cudaStream_t st )
{
  if (s1 <= s0) return;
  unsigned g = nblk( (s1-s0)*32, NODE_THREADS );
  const unsigned char* skip = nullptr;
  if (c->nsh > 0 && c->comm) {
    if (c->sh_flag.n != c->npoin) {
      std::vector< unsigned char > f( c->npoin, 0 );
      for (auto i : c->sh_node_h) f[(size_t)i] = 1;
      c->sh_flag.upload( f, st );
    }
    skip = c->sh_flag.p;
  }
  if (fused && c->lax)
    k_rhs_node< true, true ><<< g, NODE_THREADS, 0, st >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->F.p, c->nslot,
      c->bslot.p, c->Rb.p, c->S.p, c->src_mask, c->v.p, c->vol.p, Un, A, Uout, c->W.p, c->R.p, c->Wn.p, c->Un.p, s0, s1, skip );
  else if (fused)
    k_rhs_node< true, false ><<< g, NODE_THREADS, 0, st >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->F.p, c->nslot,
      c->bslot.p, c->Rb.p, c->S.p, c->src_mask, c->v.p, c->vol.p, Un, A, Uout, c->W.p, c->R.p, c->Wn.p, c->Un.p, s0, s1, skip );
  else
    k_rhs_node< false, false ><<< g, NODE_THREADS, 0, st >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->F.p, c->nslot,
      c->bslot.p, c->Rb.p, c->S.p, c->src_mask, c->v.p, c->vol.p, Un, A, Uout, c->W.p, c->R.p, c->Wn.p, c->Un.p, s0, s1, skip );
  ++c->launches;
}

// nodal gather of the rhs; fused = apply the RK update in the same pass
// Uin: conserved state the fluxes were computed from; Un: state at time level n;
// Uout: where the updated state goes (may alias Uin)
void do_rhs_nodes( xyst_ctx* c, bool fused, int stage, double dt, const double* Uin, const double* Un, double* Uout )
{
  auto s = c->stream;
  bool halo = c->nsh > 0 && c->comm;
  if (c->rb_pending) {                   // computed on the side stream during this stage's do_grad
    CK( cudaStreamWaitEvent( s, c->ev_e, 0 ) );
    if (halo) CK( cudaStreamWaitEvent( c->comm_stream, c->ev_e, 0 ) );
    c->rb_pending = false;
  } else if (c->nbn) {
    k_bnd_rhs<<< nblk( c->nbn, 128 ), 128, 0, s >>>( (int)c->nbn, c->NP, c->bn_off.p, c->bn_face.p,
      c->tri.p, c->besym.p, c->fn.p, Uin, c->Rb.p, c->prm.gamma, c->W.p, c->rgas ); ++c->launches;
  }
  StageArgs A{ rkcoef[stage], dt, c->steady ? c->dtp.p : nullptr, stage, mode( c ) };
  if (halo) {                            // fluxes (and Rb) are ready on the compute stream here
    auto cs = c->comm_stream;
    CK( cudaEventRecord( c->ev_a, s ) );
    CK( cudaStreamWaitEvent( cs, c->ev_a, 0 ) );
    k_rhs_shared<<< nblk( c->nsh, 128 ), 128, 0, cs >>>( (int)c->nsh, c->sh_node.p, c->sl_base.p,
      c->inc_e.p, c->F.p, c->nslot, c->bslot.p, c->Rb.p, c->S.p, c->src_mask, c->v.p, c->sh_part.p ); ++c->launches;
    exchange( c, NC, true );
  }
  {
    ProfScope ps( c, fused ? "update" : "rhsnode" );
    launch_rhs_node( c, fused, A, Un, Uout, 0, c->nslice, s );
  }
  if (halo) {
    exchange_wait( c );
    unsigned g = nblk( c->nsh, 128 );
    if (fused)
      k_rhs_finish< true ><<< g, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p,
        c->sh_part.p, c->sh_recvbuf.p, c->vol.p, Un, A, Uout, c->W.p, c->R.p, c->Wn.p, c->Un.p );
    else
      k_rhs_finish< false ><<< g, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p,
        c->sh_part.p, c->sh_recvbuf.p, c->vol.p, Un, A, Uout, c->W.p, c->R.p, c->Wn.p, c->Un.p );
    ++c->launches;
  }
  CK( cudaGetLastError() );
}

void do_bc( xyst_ctx* c )
{
  if (!c->nbc) return;
  FarState fs{ c->far_r, c->far_p, c->far_u[0], c->far_u[1], c->far_u[2] };
  k_bc<<< nblk( c->nbc, 128 ), 128, 0, c->stream >>>( (int)c->nbc, c->NP, c->bc_node.p, c->bc_dir.p,
    c->dir_mask.p, c->dir_val.p, c->bc_symoff.p, c->sym_n.p, c->bc_faroff.p, c->far_n.p, fs,
    c->bc_pre.p, c->pre_val.p, c->prm.gamma, c->U.p, c->W.p, mode( c ) ); ++c->launches;
  CK( cudaGetLastError() );
}


void save_un( xyst_ctx* c ) {           // RieCG.cpp:1011  m_un = m_u
  CK( cudaMemcpyAsync( c->Un.p, c->U.p, c->NP*NC*sizeof(double), cudaMemcpyDeviceToDevice, c->stream ) );
}

} // namespace

// =================================================================================
// C ABI
// =================================================================================
extern "C" {

const char* xyst_last_error(void) { return g_err.c_str(); }

int xyst_device_count(void) { int n = 0; if (cudaGetDeviceCount( &n ) != cudaSuccess) return 0; return n; }

int xyst_ctx_create( int device, const xyst_params* params, xyst_ctx** out )
{
  API_BEGIN
  if (!params || !out) throw std::runtime_error( "null argument" );
  if (params->ncomp != NC) throw std::runtime_error( "only ncomp = 5 (Euler system) is supported" );
  if (params->flux != 0 && params->flux != 1) throw std::runtime_error( "Flux not configured" );
  int n = 0;
  if (cudaGetDeviceCount( &n ) != cudaSuccess || n == 0)
    throw std::runtime_error( "no CUDA device: the B200 path has no CPU fallback" );
  if (device < 0 || device >= n) throw std::runtime_error( "invalid device ordinal" );
  CK( cudaSetDevice( device ) );
  auto c = new xyst_ctx;
  c->device = device; c->prm = *params;
  CK( cudaStreamCreateWithFlags( &c->stream, cudaStreamNonBlocking ) ); c->own_stream = true;
  // side streams at the highest priority: their small kernels (boundary terms, shared-node sums,
  // packing, NCCL send/recv) get the thread-block slots the full-mesh gather frees first
  // instead of queueing behind all of its blocks
  { int lo = 0, hi = 0;
    CK( cudaDeviceGetStreamPriorityRange( &lo, &hi ) );
    CK( cudaStreamCreateWithPriority( &c->comm_stream, cudaStreamNonBlocking, hi ) );
    CK( cudaStreamCreateWithPriority( &c->aux_stream, cudaStreamNonBlocking, hi ) ); }
  CK( cudaEventCreateWithFlags( &c->ev_c, cudaEventDisableTiming ) );
  CK( cudaEventCreateWithFlags( &c->ev_d, cudaEventDisableTiming ) );
  CK( cudaEventCreateWithFlags( &c->ev_e, cudaEventDisableTiming ) );
  CK( cudaEventCreateWithFlags( &c->ev_a, cudaEventDisableTiming ) );
  CK( cudaEventCreateWithFlags( &c->ev_b, cudaEventDisableTiming ) );
  c->red.alloc( (size_t)RED_BLOCKS*NDIAG + NDIAG );
  CK( cudaMallocHost( &c->red_host, NDIAG*sizeof(double) ) );
  *out = c;
  API_END
}

int xyst_ctx_set_stream( xyst_ctx* c, void* st )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (c->own_stream && c->stream) { CK( cudaStreamSynchronize( c->stream ) ); CK( cudaStreamDestroy( c->stream ) ); }
  c->stream = static_cast< cudaStream_t >( st ); c->own_stream = false;
  API_END
}

int xyst_ctx_destroy( xyst_ctx* c )
{
  API_BEGIN
  if (!c) return 0;
  cudaSetDevice( c->device );
  cudaDeviceSynchronize();
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy( c->comm );
  for (auto& [k,p] : c->prof) for (auto& e : p.ev) { cudaEventDestroy( e.first ); cudaEventDestroy( e.second ); }
  if (c->own_stream && c->stream) cudaStreamDestroy( c->stream );
  if (c->comm_stream) cudaStreamDestroy( c->comm_stream );
  if (c->aux_stream) cudaStreamDestroy( c->aux_stream );
  for (auto e : { c->ev_c, c->ev_d, c->ev_e }) if (e) cudaEventDestroy( e );
  if (c->ev_a) cudaEventDestroy( c->ev_a );
  if (c->ev_b) cudaEventDestroy( c->ev_b );
  if (c->red_host) cudaFreeHost( c->red_host );
  delete c;
  API_END
}

int xyst_sync( xyst_ctx* c ) { API_BEGIN CK( cudaSetDevice( c->device ) ); CK( cudaStreamSynchronize( c->stream ) ); API_END }

static int mesh_upload_impl( xyst_ctx* c, size_t npoin, const double* x, const double* y, const double* z,
                      const size_t nsup[3], const size_t* const dsupedge[3],
                      const double* const dsupint[3], size_t ntri, const size_t* triinpoel,
                      const uint8_t* besym, const double* vol, const double* v, size_t stride )
{
  API_BEGIN
  c->dstride = (int)stride;
  CK( cudaSetDevice( c->device ) );
  if (npoin == 0 || npoin > 0x7fffffffULL) throw std::runtime_error( "npoin out of range" );
  // --- flatten superedges to edges (orientation and normals as given) ---------------
  // tk::lpoed / tk::lpoet, src/Mesh/DerivedData.hpp:40-44
  static const int lpoed[6][2] = { {0,1}, {1,2}, {2,0}, {0,3}, {1,3}, {2,3} };
  static const int lpoet[3][2] = { {0,1}, {1,2}, {2,0} };
  size_t ne = nsup[0]*6 + nsup[1]*3 + nsup[2];
  if (ne > 0x7ffffff0ULL) throw std::runtime_error( "too many edges for 32-bit edge ids" );
  struct E { int p, q; double d[5]; };
  const int drows = stride > 4 ? 5 : 4;
  std::vector< E > edges; edges.reserve( ne );
  auto chk = [&]( size_t id ){ if (id >= npoin) throw std::runtime_error( "node id out of range in superedge" ); return (int)id; };
  for (size_t e=0; e<nsup[0]; ++e)
    for (int k=0; k<6; ++k) {
      E ed; ed.p = chk( dsupedge[0][e*4+lpoed[k][0]] ); ed.q = chk( dsupedge[0][e*4+lpoed[k][1]] );
      ed.d[3] = ed.d[4] = 0.0; for (size_t j=0; j<stride; ++j) ed.d[j] = dsupint[0][(e*6+k)*stride+j];
      edges.push_back( ed );
    }
  for (size_t e=0; e<nsup[1]; ++e)
    for (int k=0; k<3; ++k) {
      E ed; ed.p = chk( dsupedge[1][e*3+lpoet[k][0]] ); ed.q = chk( dsupedge[1][e*3+lpoet[k][1]] );
      ed.d[3] = ed.d[4] = 0.0; for (size_t j=0; j<stride; ++j) ed.d[j] = dsupint[1][(e*3+k)*stride+j];
      edges.push_back( ed );
    }
  for (size_t e=0; e<nsup[2]; ++e) {
    E ed; ed.p = chk( dsupedge[2][e*2+0] ); ed.q = chk( dsupedge[2][e*2+1] );
    ed.d[3] = ed.d[4] = 0.0; for (size_t j=0; j<stride; ++j) ed.d[j] = dsupint[2][e*stride+j];
    edges.push_back( ed );
  }
  // --- edge slots in owner order ----------------------------------------------------
  // owner = lower endpoint; its edges sorted by the other endpoint. The j-th edge owned
  // by node o lives in slot ebase[o/32] + j*32 + o%32 (padding slots have ep = -1).
  std::vector< int > perm( ne );
  std::iota( perm.begin(), perm.end(), 0 );
  std::sort( perm.begin(), perm.end(), [&]( int a, int b ){
    int la = std::min( edges[a].p, edges[a].q ), lb = std::min( edges[b].p, edges[b].q );
    if (la != lb) return la < lb;
    int ha = std::max( edges[a].p, edges[a].q ), hb = std::max( edges[b].p, edges[b].q );
    if (ha != hb) return ha < hb;
    return a < b; } );
  size_t nslice = (npoin + 31) / 32;
  std::vector< int > udeg( npoin, 0 );
  for (size_t i=0; i<ne; ++i) ++udeg[ std::min( edges[i].p, edges[i].q ) ];
  std::vector< long long > ebase( nslice+1, 0 );
  for (size_t s=0; s<nslice; ++s) {
    int km = 0;
    for (size_t p=s*32; p<std::min( npoin, s*32+32 ); ++p) km = std::max( km, udeg[p] );
    ebase[s+1] = ebase[s] + (long long)km*32;
  }
  size_t nslot = (size_t)ebase[nslice];
  if (nslot > 0x7ffffff0ULL) throw std::runtime_error( "too many edge slots for 32-bit ids" );
  std::vector< int > ep( nslot, -1 ), eq( nslot, -1 ), slot_of( ne );
  std::vector< double > ed( (size_t)drows*nslot, 0.0 );
  { std::vector< int > fillu( npoin, 0 );
    for (size_t i=0; i<ne; ++i) {
      const auto& e = edges[perm[i]];
      int o = std::min( e.p, e.q );
      size_t sl = (size_t)ebase[o/32] + (size_t)fillu[o]*32 + (size_t)(o%32);
      ++fillu[o];
      ep[sl] = e.p; eq[sl] = e.q; slot_of[i] = (int)sl;
      for (int j=0; j<drows; ++j) ed[j*nslot+sl] = e.d[j];
    } }
  // --- sliced-ELL incidence: node -> (signed slot, neighbour) -------------------------
  // reference scatter: G(p) -= f, G(q) += f with f = d*(u_q+u_p)  (Riemann.cpp:321-323).
  // Per node: its owned edges first (ascending other endpoint), then the edges owned by
  // lower neighbours (ascending neighbour) -- the fixed summation order of the gathers.
  std::vector< int > deg( npoin, 0 );
  for (size_t i=0; i<ne; ++i) { ++deg[edges[i].p]; ++deg[edges[i].q]; }
  std::vector< long long > base( nslice+1, 0 );
  for (size_t s=0; s<nslice; ++s) {
    int km = 0;
    for (size_t p=s*32; p<std::min( npoin, s*32+32 ); ++p) km = std::max( km, deg[p] );
    base[s+1] = base[s] + (long long)km*32;
  }
  size_t nent = (size_t)base[nslice];
  c->maxdeg = 0; for (size_t p=0; p<npoin; ++p) c->maxdeg = std::max( c->maxdeg, deg[p] );
  std::vector< int > inc_e( nent, 0 ), inc_q( nent, 0 ), fill( npoin, 0 );
  for (size_t sl=0; sl<nslice; ++sl)            // padding: the node itself (last node for the tail slice)
    for (size_t j=(size_t)base[sl]; j<(size_t)base[sl+1]; ++j) inc_q[j] = (int)std::min( npoin-1, sl*32 + (j - (size_t)base[sl])%32 );
  auto addinc = [&]( int node, int other, int signedslot ) {
    size_t slot = (size_t)base[node/32] + (size_t)fill[node]*32 + (size_t)(node%32);
    ++fill[node];
    inc_e[slot] = signedslot; inc_q[slot] = other;
  };
  for (size_t i=0; i<ne; ++i) {                 // owned edges (sorted by owner, other)
    const auto& e = edges[perm[i]];
    int o = std::min( e.p, e.q ), h = std::max( e.p, e.q );
    addinc( o, h, o == e.q ? slot_of[i]+1 : -(slot_of[i]+1) );
  }
  for (size_t i=0; i<ne; ++i) {                 // edges owned by lower neighbours: arrive ascending in owner
    const auto& e = edges[perm[i]];
    int o = std::min( e.p, e.q ), h = std::max( e.p, e.q );
    addinc( h, o, h == e.q ? slot_of[i]+1 : -(slot_of[i]+1) );
  }
  // --- boundary faces: node -> (face, local index) CSR -------------------------------
  std::vector< int > tri( ntri*3 );
  for (size_t i=0; i<ntri*3; ++i) tri[i] = chk( triinpoel[i] );
  std::vector< int > bslot( npoin, -1 ), bn_node;
  for (size_t i=0; i<ntri*3; ++i) if (bslot[tri[i]] < 0) { bslot[tri[i]] = 0; }
  for (size_t p=0; p<npoin; ++p) if (bslot[p] == 0) { bslot[p] = (int)bn_node.size(); bn_node.push_back( (int)p ); }
  size_t nbn = bn_node.size();
  std::vector< int > bn_off( nbn+1, 0 ), bn_face( ntri*3 );
  for (size_t i=0; i<ntri*3; ++i) ++bn_off[ bslot[tri[i]]+1 ];
  for (size_t b=0; b<nbn; ++b) bn_off[b+1] += bn_off[b];
  { std::vector< int > f( bn_off.begin(), bn_off.end()-1 );
    for (size_t t=0; t<ntri; ++t) for (int k=0; k<3; ++k) bn_face[ f[ bslot[tri[t*3+k]] ]++ ] = (int)(t*4) + k; }
  // --- upload ------------------------------------------------------------------------
  auto s = c->stream;
  size_t NP = nslice*32;
  c->npoin = npoin; c->NP = NP; c->nedge = ne; c->nslot = nslot; c->ntri = ntri; c->nslice = nslice;
  c->nent = nent; c->nbn = nbn;
  c->ep.upload( ep, s ); c->eq.upload( eq, s ); c->D.upload( ed, s );
  { std::vector< double2 > d2( nslot );
    for (size_t i=0; i<nslot; ++i) d2[i] = make_double2( ed[i], ed[nslot+i] );
    c->D2.upload( d2, s );
    std::vector< int2 > iq( nent );
    for (size_t i=0; i<nent; ++i) iq[i] = make_int2( inc_e[i], inc_q[i] );
    c->inc_eq.upload( iq, s ); }
  c->sl_base.upload( base, s ); c->inc_e.upload( inc_e, s ); c->inc_q.upload( inc_q, s );
  c->tri.upload( tri, s );
  c->besym.upload( ntri ? std::vector< unsigned char >( besym, besym + ntri*3 ) : std::vector< unsigned char >(), s );
  c->bslot.upload( bslot, s ); c->bn_node.upload( bn_node, s ); c->bn_off.upload( bn_off, s );
  c->bn_face.upload( bn_face, s );
  c->Gb.alloc( nbn*15 ); c->Rb.alloc( nbn*NC );
  c->bcof.upload( std::vector< int >( npoin, -1 ), s );
  c->zP.release(); c->zQ.release(); c->zUL.release();
  { std::vector< double > pv( NP, 1.0 ), pw( NP, 1.0 ), px( 3*NP, 0.0 );
    for (size_t p=0; p<npoin; ++p) { pv[p] = vol[p]; pw[p] = v[p]; px[p] = x[p]; px[NP+p] = y[p]; px[2*NP+p] = z[p]; }
    c->vol.upload( pv, s ); c->v.upload( pw, s ); c->X.upload( px, s ); }
  c->fn.alloc( std::max< size_t >( ntri, 1 )*3 );
  if (ntri) { k_face_normals<<< nblk( ntri, 128 ), 128, 0, s >>>( (int)ntri, NP, c->tri.p, c->X.p, c->fn.p ); ++c->launches; CK( cudaGetLastError() ); }
  if (c->cho) {            // the projection solver keeps its own state (chocg.cuh)
    for (auto* b : { &c->U, &c->Un, &c->W, &c->G, &c->R, &c->stage, &c->F }) b->release();
    c->S.release(); c->src_mask = 0;
    CK( cudaStreamSynchronize( s ) );
    return 0;
  }
  c->U.alloc( NP*NC ); c->Un.alloc( NP*NC ); c->G.alloc( NP*15 );
  c->R.alloc( npoin*NC ); c->stage.alloc( npoin*NC ); c->F.alloc( std::max< size_t >( nslot, 1 )*NC );
  { std::vector< double > one( NP*NC, 1.0 );    // a harmless state until xyst_state_set
    c->U.upload( one, s ); c->Un.upload( one, s );
    // primitives + coordinates as pairs (w0,w1) (w2,w3) (w4,x) (y,z), see load_wx
    std::vector< double > wx( NP*8, 1.0 );
    for (size_t p=0; p<npoin; ++p) { wx[(2*NP+p)*2+1] = x[p]; wx[(3*NP+p)*2] = y[p]; wx[(3*NP+p)*2+1] = z[p]; }
    c->W.upload( wx, s ); }
  CK( cudaMemsetAsync( c->G.p, 0, NP*15*sizeof(double), s ) );
  CK( cudaMemsetAsync( c->F.p, 0, std::max< size_t >( nslot, 1 )*NC*sizeof(double), s ) );
  c->S.release(); c->src_mask = 0;
  CK( cudaStreamSynchronize( s ) );
  API_END
}

int xyst_mesh_upload( xyst_ctx* c, size_t npoin, const double* x, const double* y, const double* z,
                      const size_t nsup[3], const size_t* const dsupedge[3],
                      const double* const dsupint[3], size_t ntri, const size_t* triinpoel,
                      const uint8_t* besym, const double* vol, const double* v )
{ return mesh_upload_impl( c, npoin, x, y, z, nsup, dsupedge, dsupint, ntri, triinpoel, besym, vol, v, 3 ); }

int xyst_zalcg_mesh_upload( xyst_ctx* c, size_t npoin, const double* x, const double* y, const double* z,
                            const size_t nsup[3], const size_t* const dsupedge[3],
                            const double* const dsupint[3], size_t ntri, const size_t* triinpoel,
                            const uint8_t* besym, const double* vol, const double* v )
{ return mesh_upload_impl( c, npoin, x, y, z, nsup, dsupedge, dsupint, ntri, triinpoel, besym, vol, v, 4 ); }

int xyst_zalcg_config( xyst_ctx* c, const xyst_zalcg_params* p )
{
  API_BEGIN
  if (!p) throw std::runtime_error( "null argument" );
  c->zal = *p;
  API_END
}

namespace {
void zal_need( xyst_ctx* c ) {
  need_mesh( c );
  if (c->dstride != 4) throw std::runtime_error( "ZalCG needs stride-4 superedge integrals: use xyst_zalcg_mesh_upload" );
  if (c->nsh > 0 && c->comm) throw std::runtime_error( "ZalCG on several partitions is not implemented yet" );
  if (!c->zP.p) { c->zP.alloc( c->NP*10 ); c->zQ.alloc( c->NP*10 ); c->zUL.alloc( c->NP*NC ); }
}
void zal_flux_and_bnd( xyst_ctx* c, double dt )
{
  auto s = c->stream;
  if (c->nbn) { k_bnd_rhs<<< nblk( c->nbn, 128 ), 128, 0, s >>>( (int)c->nbn, c->NP, c->bn_off.p, c->bn_face.p,
                  c->tri.p, c->besym.p, c->fn.p, c->U.p, c->Rb.p, c->prm.gamma, c->W.p, c->rgas ); ++c->launches; }
  { ProfScope ps( c, "zalflux" );
    k_zal_flux_edge<<< nblk( c->nslot, 128 ), 128, 0, s >>>( c->nslot, c->NP, c->ep.p, c->eq.p, c->D.p, c->U.p, c->X.p,
      dt, dparams( c ), c->F.p ); ++c->launches; }
}
void zal_node1( xyst_ctx* c, double dt, int fct )
{
  k_zal_node1<<< nblk( c->nslice*32, NODE_THREADS ), NODE_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->sl_base.p,
    c->inc_eq.p, c->D.p, c->nslot, c->F.p, c->U.p, c->bslot.p, c->Rb.p, c->bcof.p, c->bc_symoff.p,
    c->sym_n.p, c->vol.p, dt, c->zal.fctdif, fct, c->zP.p, c->zUL.p, c->R.p ); ++c->launches;
}
}

int xyst_zalcg_rhs( xyst_ctx* c, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  zal_need( c );
  zal_flux_and_bnd( c, dt );
  zal_node1( c, dt, 0 );
  CK( cudaGetLastError() );
  API_END
}

int xyst_zalcg_step( xyst_ctx* c, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  zal_need( c );
  auto s = c->stream;
  unsigned g = nblk( c->nslice*32, NODE_THREADS );
  zal_flux_and_bnd( c, dt );
  if (c->zal.fct) {
    zal_node1( c, dt, 1 );
    k_zal_node2<<< g, NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->sl_base.p, c->inc_eq.p, c->U.p, c->zUL.p,
      c->zP.p, c->zal.fctclip, c->zQ.p ); ++c->launches;
    // new state into the other buffer, then swap: Un keeps the old state for the diagnostics
    k_zal_node3<<< g, NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->sl_base.p, c->inc_eq.p, c->D.p, c->nslot,
      c->U.p, c->zUL.p, c->zQ.p, c->vol.p, c->zal.fctdif, c->zal.fctsys_mask, c->Un.p, c->W.p ); ++c->launches;
  } else {
    zal_node1( c, dt, 0 );
    k_zal_nofct<<< nblk( c->npoin, 256 ), 256, 0, s >>>( c->npoin, c->NP, c->R.p, c->vol.p, c->U.p, dt, c->Un.p, c->W.p ); ++c->launches;
  }
  std::swap( c->U.p, c->Un.p );
  do_bc( c );
  CK( cudaGetLastError() );
  API_END
}

int xyst_bc_upload( xyst_ctx* c, size_t ndir, const size_t* dirbcmasks, const double* dirvals,
                    size_t nsym, const size_t* symbcnodes, const double* symbcnorms,
                    size_t nfar, const size_t* farbcnodes, const double* farbcnorms,
                    double far_density, double far_pressure, const double far_velocity[3],
                    size_t npre, const size_t* prebcnodes, const double* prebcvals )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  // union of BC nodes; per node: dirichlet slot, symmetry entries, farfield entries,
  // pressure slot -- entries keep the list order of the reference (a node repeats once
  // per side set it has a normal in, RieCG.cpp:583-596)
  std::map< int, int > slot;
  std::vector< int > nodes;
  auto add = [&]( size_t p ){ if (p >= c->npoin) throw std::runtime_error( "BC node id out of range" );
    auto it = slot.find( (int)p ); if (it == slot.end()) { slot[(int)p] = (int)nodes.size(); nodes.push_back( (int)p ); } };
  for (size_t i=0; i<ndir; ++i) add( dirbcmasks[i*(NC+1)] );
  for (size_t i=0; i<nsym; ++i) add( symbcnodes[i] );
  for (size_t i=0; i<nfar; ++i) add( farbcnodes[i] );
  for (size_t i=0; i<npre; ++i) add( prebcnodes[i] );
  // sort the union by node id for locality, rebuild slots
  std::sort( nodes.begin(), nodes.end() );
  for (size_t i=0; i<nodes.size(); ++i) slot[nodes[i]] = (int)i;
  size_t nbc = nodes.size();
  std::vector< int > dir( nbc, -1 ), pre( nbc, -1 ), symoff( nbc+1, 0 ), faroff( nbc+1, 0 );
  std::vector< int > dmask( ndir*NC );
  std::vector< double > dval( ndir*NC, 0.0 ), symn( nsym*3 ), farn( nfar*3 ), preval( npre*2 );
  for (size_t i=0; i<ndir; ++i) {
    dir[ slot[(int)dirbcmasks[i*(NC+1)]] ] = (int)i;
    for (int k=0; k<NC; ++k) { dmask[i*NC+k] = (int)dirbcmasks[i*(NC+1)+1+k]; if (dirvals) dval[i*NC+k] = dirvals[i*NC+k]; }
  }
  for (size_t i=0; i<nsym; ++i) ++symoff[ slot[(int)symbcnodes[i]]+1 ];
  for (size_t i=0; i<nfar; ++i) ++faroff[ slot[(int)farbcnodes[i]]+1 ];
  for (size_t i=0; i<nbc; ++i) { symoff[i+1] += symoff[i]; faroff[i+1] += faroff[i]; }
  { std::vector< int > f( symoff.begin(), symoff.end()-1 );
    for (size_t i=0; i<nsym; ++i) { int k = f[ slot[(int)symbcnodes[i]] ]++; for (int j=0; j<3; ++j) symn[(size_t)k*3+j] = symbcnorms[i*3+j]; } }
  { std::vector< int > f( faroff.begin(), faroff.end()-1 );
    for (size_t i=0; i<nfar; ++i) { int k = f[ slot[(int)farbcnodes[i]] ]++; for (int j=0; j<3; ++j) farn[(size_t)k*3+j] = farbcnorms[i*3+j]; } }
  for (size_t i=0; i<npre; ++i) { pre[ slot[(int)prebcnodes[i]] ] = (int)i; preval[i*2] = prebcvals[i*2]; preval[i*2+1] = prebcvals[i*2+1]; }
  auto s = c->stream;
  c->nbc = nbc; c->ndir = ndir;
  { std::vector< int > bcof( c->npoin, -1 ); for (size_t i=0; i<nbc; ++i) bcof[nodes[i]] = (int)i; c->bcof.upload( bcof, s ); }
  c->bc_node.upload( nodes, s ); c->bc_dir.upload( dir, s ); c->bc_pre.upload( pre, s );
  c->bc_symoff.upload( symoff, s ); c->bc_faroff.upload( faroff, s );
  c->dir_mask.upload( dmask, s ); c->dir_val.upload( dval, s );
  c->sym_n.upload( symn, s ); c->far_n.upload( farn, s ); c->pre_val.upload( preval, s );
  c->far_r = far_density; c->far_p = far_pressure;
  if (far_velocity) for (int j=0; j<3; ++j) c->far_u[j] = far_velocity[j];
  API_END
}

int xyst_dirbc_values( xyst_ctx* c, const double* dirvals )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (c->ndir) { CK( cudaMemcpyAsync( c->dir_val.p, dirvals, c->ndir*NC*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
                 CK( cudaStreamSynchronize( c->stream ) ); }
  API_END
}

int xyst_src_upload( xyst_ctx* c, const double* S )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (!S) { c->S.release(); c->src_mask = 0; return 0; }
  int mask = 0;
  for (size_t p=0; p<c->npoin; ++p) for (int k=0; k<NC; ++k) if (S[p*NC+k] != 0.0) mask |= 1<<k;
  c->S.upload( std::vector< double >( S, S + c->npoin*NC ), c->stream );
  c->src_mask = mask;
  API_END
}

int xyst_state_set( xyst_ctx* c, const double* U )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (c->rb_pending) { CK( cudaStreamSynchronize( c->aux_stream ) ); c->rb_pending = false; }
  CK( cudaMemcpyAsync( c->stage.p, U, c->npoin*NC*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  k_set_state<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, c->stage.p, c->U.p, c->W.p, mode( c ) );
  ++c->launches;
  CK( cudaGetLastError() );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

int xyst_state_get( xyst_ctx* c, double* U )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  k_get_state<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, c->U.p, c->stage.p );
  ++c->launches;
  CK( cudaGetLastError() );
  CK( cudaMemcpyAsync( U, c->stage.p, c->npoin*NC*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

int xyst_riecg_grad( xyst_ctx* c ) { API_BEGIN CK( cudaSetDevice( c->device ) ); do_grad( c ); API_END }

int xyst_grad_get( xyst_ctx* c, double* G )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  std::vector< double > h( c->NP*15 );
  CK( cudaMemcpyAsync( h.data(), c->G.p, h.size()*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  for (size_t p=0; p<c->npoin; ++p) for (int i=0; i<15; ++i) G[p*15+(size_t)i] = h[gidx( i, p, c->NP )];
  API_END
}

int xyst_riecg_rhs( xyst_ctx* c )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  do_flux( c );
  do_rhs_nodes( c, false, 0, 0.0, c->U.p, c->Un.p, c->U.p );
  API_END
}

int xyst_rhs_get( xyst_ctx* c, double* R )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  CK( cudaMemcpyAsync( R, c->R.p, c->npoin*NC*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

int xyst_rk_update( xyst_ctx* c, int stage, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (stage < 0 || stage > 2) throw std::runtime_error( "stage must be 0, 1 or 2" );
  if (stage == 0 && !c->lax) save_un( c );
  StageArgs A{ rkcoef[stage], dt, c->steady ? c->dtp.p : nullptr, stage, mode( c ) };
  k_update<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, c->R.p, c->vol.p, c->Un.p,
    A, c->U.p, c->W.p, c->Wn.p, c->Un.p ); ++c->launches;
  CK( cudaGetLastError() );
  API_END
}

int xyst_apply_bc( xyst_ctx* c ) { API_BEGIN CK( cudaSetDevice( c->device ) ); need_mesh( c ); do_bc( c ); API_END }

int xyst_dt_min( xyst_ctx* c, double cfl, double* dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( c->npoin, RED_THREADS ) );
  k_dt<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->U.p, c->vol.p, c->prm.gamma, c->red.p, mode( c ), cfl,
    c->steady ? c->dtp.p : nullptr );
  k_reduce_final< 1, true ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, c->red.p + (size_t)RED_BLOCKS*NDIAG );
  c->launches += 2;
  CK( cudaMemcpyAsync( c->red_host, c->red.p + (size_t)RED_BLOCKS*NDIAG, sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  *dt = c->red_host[0] * cfl;
  API_END
}

int xyst_riecg_stage( xyst_ctx* c, int stage, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (stage < 0 || stage > 2) throw std::runtime_error( "stage must be 0, 1 or 2" );
  do_grad( c );
  do_flux( c );
  if (c->lax)           // time level n is kept in (p,u,v,w,T) form (Wn); Un is refreshed at stage 2
    do_rhs_nodes( c, true, stage, dt, c->U.p, c->Un.p, c->U.p );
  else if (stage == 0) { // un = u (RieCG.cpp:1011) without a copy: write the new state into the
                        // other buffer and swap the roles of the two
    do_rhs_nodes( c, true, stage, dt, c->U.p, c->U.p, c->Un.p );
    std::swap( c->U.p, c->Un.p );
  } else
    do_rhs_nodes( c, true, stage, dt, c->U.p, c->Un.p, c->U.p );
  do_bc( c );
  API_END
}

int xyst_riecg_step( xyst_ctx* c, double dt )
{
  for (int s=0; s<3; ++s) if (int r = xyst_riecg_stage( c, s, dt )) return r;
  return 0;
}

int xyst_diag( xyst_ctx* c, const double* an, double* out )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  DevBuf< double > dan;
  if (an) dan.upload( std::vector< double >( an, an + c->npoin*NC ), c->stream );
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( c->npoin, RED_THREADS ) );
  k_diag<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->U.p, c->Un.p, c->v.p, dan.p, c->red.p );
  k_reduce_final< NDIAG, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, c->red.p + (size_t)RED_BLOCKS*NDIAG );
  c->launches += 2;
  CK( cudaMemcpyAsync( c->red_host, c->red.p + (size_t)RED_BLOCKS*NDIAG, NDIAG*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  for (int i=0; i<NDIAG; ++i) out[i] = c->red_host[i];
  API_END
}

// ---- multi-GPU ----------------------------------------------------------------------
int xyst_comm_unique_id( void* id128 )
{
  API_BEGIN
  if (!g_nccl.load()) throw std::runtime_error( "cannot load libnccl.so.2" );
  Nccl::UniqueId id;
  NK( g_nccl.GetUniqueId( &id ) );
  std::memcpy( id128, &id, 128 );
  API_END
}

int xyst_comm_init( xyst_ctx* c, int nranks, int rank, const void* id128 )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!g_nccl.load()) throw std::runtime_error( "cannot load libnccl.so.2" );
  Nccl::UniqueId id;
  std::memcpy( &id, id128, 128 );
  NK( g_nccl.CommInitRank( &c->comm, nranks, id, rank ) );
  c->nranks = nranks; c->rank = rank;
  API_END
}

int xyst_halo_upload( xyst_ctx* c, int nneigh, const int* neigh_rank, const size_t* neigh_off,
                      const size_t* shared )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  c->neigh.assign( neigh_rank, neigh_rank + nneigh );
  c->neigh_off.assign( neigh_off, neigh_off + nneigh + 1 );
  size_t nsend = nneigh ? neigh_off[nneigh] : 0;
  std::vector< int > uniq;
  for (size_t i=0; i<nsend; ++i) {   // may be called before the mesh upload (volume exchange)
    if (c->npoin && shared[i] >= c->npoin) throw std::runtime_error( "shared node id out of range" );
    uniq.push_back( (int)shared[i] ); }
  std::sort( uniq.begin(), uniq.end() );
  uniq.erase( std::unique( uniq.begin(), uniq.end() ), uniq.end() );
  std::vector< int > send( nsend ), roff( uniq.size()+1, 0 ), ridx( nsend );
  for (size_t i=0; i<nsend; ++i) {
    send[i] = (int)( std::lower_bound( uniq.begin(), uniq.end(), (int)shared[i] ) - uniq.begin() );
    ++roff[ send[i]+1 ];
  }
  for (size_t i=0; i<uniq.size(); ++i) roff[i+1] += roff[i];
  { std::vector< int > f( roff.begin(), roff.end()-1 );
    for (size_t i=0; i<nsend; ++i) ridx[ f[send[i]]++ ] = (int)i; }   // ascending recv position = fixed neighbour order
  auto s = c->stream;
  c->nsh = uniq.size(); c->nsend = nsend; c->sh_node_h = uniq; c->sh_flag.release();
  c->sh_node.upload( uniq, s ); c->sh_send.upload( send, s ); c->sh_roff.upload( roff, s ); c->sh_ridx.upload( ridx, s );
  c->sh_part.alloc( uniq.size()*15 ); c->sh_sendbuf.alloc( nsend*15 ); c->sh_recvbuf.alloc( nsend*15 );
  API_END
}

__global__ void k_halo_add( int nsh, int w, const int* __restrict__ roff, const int* __restrict__ ridx,
                            const double* __restrict__ recvbuf, double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  for (int j=0; j<w; ++j) {
    double a = part[(size_t)i*w+j];
    for (int r=roff[i]; r<roff[i+1]; ++r) a += recvbuf[(size_t)ridx[r]*w+j];
    part[(size_t)i*w+j] = a;
  }
}

int xyst_halo_sum( xyst_ctx* c, int w, double* vals )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!c->nsh || !c->comm) return 0;
  if (w < 1 || w > 15) throw std::runtime_error( "halo_sum: width must be 1..15" );
  CK( cudaMemcpyAsync( c->sh_part.p, vals, c->nsh*(size_t)w*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  exchange( c, w );
  exchange_wait( c );
  k_halo_add<<< nblk( c->nsh, 128 ), 128, 0, c->stream >>>( (int)c->nsh, w, c->sh_roff.p, c->sh_ridx.p,
    c->sh_recvbuf.p, c->sh_part.p ); ++c->launches;
  CK( cudaMemcpyAsync( vals, c->sh_part.p, c->nsh*(size_t)w*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

static int allreduce( xyst_ctx* c, double* v, int n, int op )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!c->comm) return 0;                       // single partition: nothing to do
  if (n > NDIAG) throw std::runtime_error( "allreduce: too many values" );
  double* d = c->red.p + (size_t)RED_BLOCKS*NDIAG;
  CK( cudaMemcpyAsync( d, v, n*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  NK( g_nccl.AllReduce( d, d, (size_t)n, NCCL_FLOAT64, op, c->comm, c->stream ) );
  CK( cudaMemcpyAsync( c->red_host, d, n*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  for (int i=0; i<n; ++i) v[i] = c->red_host[i];
  API_END
}
int xyst_allreduce_min( xyst_ctx* c, double* v, int n ) { return allreduce( c, v, n, NCCL_MIN ); }
int xyst_allreduce_sum( xyst_ctx* c, double* v, int n ) { return allreduce( c, v, n, NCCL_SUM ); }

// ---- LaxCG / steady state ---------------------------------------------------------------
int xyst_laxcg_config( xyst_ctx* c, const xyst_laxcg_params* p )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (!p) throw std::runtime_error( "null argument" );
  if (!(p->rgas > 0.0)) throw std::runtime_error( "spec_gas_const must be positive" );
  c->lax = true; c->rgas = p->rgas;
  c->kvinf = p->turkel * std::sqrt( p->velinf[0]*p->velinf[0] + p->velinf[1]*p->velinf[1] + p->velinf[2]*p->velinf[2] );
  c->Wn.alloc( c->NP*NC );
  CK( cudaMemsetAsync( c->Wn.p, 0, c->NP*NC*sizeof(double), c->stream ) );
  API_END
}

int xyst_steady( xyst_ctx* c, int on )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  c->steady = on != 0;
  if (c->steady) { c->dtp.alloc( c->NP ); CK( cudaMemsetAsync( c->dtp.p, 0, c->NP*sizeof(double), c->stream ) ); }
  API_END
}

// ---- KozCG -----------------------------------------------------------------------------
int xyst_kozcg_mesh_upload( xyst_ctx* c, size_t npoin, const double* x, const double* y, const double* z,
                            size_t ntet, const size_t* inpoel, const double* vol, const double* v,
                            const double* Sn, const double* Sc )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (npoin == 0 || npoin > 0x7fffffffULL || ntet == 0 || ntet > 0x1fffffffULL) throw std::runtime_error( "size out of range" );
  // nodal part exactly as for the edge-based solvers, with no superedges and no boundary faces
  const size_t nsup[3] = { 0, 0, 0 }; const size_t* se[3] = { nullptr, nullptr, nullptr }; const double* si[3] = { nullptr, nullptr, nullptr };
  if (int r = mesh_upload_impl( c, npoin, x, y, z, nsup, se, si, 0, nullptr, nullptr, vol, v, 3 )) return r;
  // tetrahedra sorted by their lowest node (locality); the order only fixes summation order
  std::vector< int > perm( ntet );
  std::iota( perm.begin(), perm.end(), 0 );
  auto lo = [&]( int e ){ const auto N = inpoel + (size_t)e*4; return std::min( std::min( N[0], N[1] ), std::min( N[2], N[3] ) ); };
  std::stable_sort( perm.begin(), perm.end(), [&]( int a, int b ){ return lo(a) < lo(b); } );
  std::vector< int > tet( 4*ntet ), deg( npoin, 0 );
  for (size_t i=0; i<ntet; ++i)
    for (size_t a=0; a<4; ++a) {
      size_t n = inpoel[(size_t)perm[i]*4+a];
      if (n >= npoin) throw std::runtime_error( "node id out of range in inpoel" );
      tet[a*ntet+i] = (int)n; ++deg[n];
    }
  size_t nslice = c->nslice;
  std::vector< long long > base( nslice+1, 0 );
  for (size_t s=0; s<nslice; ++s) {
    int km = 0;
    for (size_t p=s*32; p<std::min( npoin, s*32+32 ); ++p) km = std::max( km, deg[p] );
    base[s+1] = base[s] + (long long)km*32;
  }
  std::vector< int > inc( (size_t)base[nslice], -1 ), fill( npoin, 0 );
  for (size_t i=0; i<ntet; ++i)
    for (size_t a=0; a<4; ++a) {
      int n = tet[a*ntet+i];
      inc[ (size_t)base[n/32] + (size_t)fill[n]*32 + (size_t)(n%32) ] = (int)(i*4+a);
      ++fill[n];
    }
  auto s = c->stream;
  c->ntet = ntet; c->knent = inc.size();
  c->ktet.upload( tet, s ); c->kbase.upload( base, s ); c->kinc.upload( inc, s );
  c->kT.alloc( 40*ntet );
  c->ksrc = Sn && Sc;
  if (c->ksrc) {
    c->S.upload( std::vector< double >( Sn, Sn + npoin*NC ), s );
    std::vector< double > sc( ntet*NC );
    for (size_t i=0; i<ntet; ++i) for (size_t k=0; k<NC; ++k) sc[i*NC+k] = Sc[(size_t)perm[i]*NC+k];
    c->kSc.upload( sc, s );
  }
  c->zP.alloc( c->NP*10 ); c->zQ.alloc( c->NP*10 ); c->zUL.alloc( c->NP*NC );
  API_END
}

namespace {
void koz_need( xyst_ctx* c ) {
  need_mesh( c );
  if (!c->ntet) throw std::runtime_error( "KozCG needs xyst_kozcg_mesh_upload" );
  if (c->nsh > 0 && c->comm) throw std::runtime_error( "KozCG on several partitions is not implemented yet" );
}
void koz_pass1( xyst_ctx* c, double dt, int fct )
{
  auto s = c->stream;
  { ProfScope ps( c, "kozelem" );
    k_koz_elem1<<< nblk( c->ntet, 128 ), 128, 0, s >>>( c->ntet, c->NP, c->ktet.p, c->U.p, c->X.p,
      c->ksrc ? c->S.p : nullptr, c->kSc.p, dt, c->prm.gamma, c->zal.fctdif, fct, c->kT.p ); ++c->launches; }
  k_koz_node1<<< nblk( c->nslice*32, NODE_THREADS ), NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->ntet, c->kbase.p, c->kinc.p,
    c->kT.p, c->U.p, c->bcof.p, c->bc_symoff.p, c->sym_n.p, c->vol.p, dt, fct, c->zP.p, c->zUL.p, c->R.p ); ++c->launches;
}
}

int xyst_kozcg_rhs( xyst_ctx* c, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  koz_need( c );
  koz_pass1( c, dt, 0 );
  CK( cudaGetLastError() );
  API_END
}

int xyst_kozcg_step( xyst_ctx* c, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  koz_need( c );
  auto s = c->stream;
  unsigned gn = nblk( c->nslice*32, NODE_THREADS ), ge = nblk( c->ntet, 128 );
  if (c->zal.fct) {
    koz_pass1( c, dt, 1 );
    k_koz_elem2<<< ge, 128, 0, s >>>( c->ntet, c->NP, c->ktet.p, c->U.p, c->zUL.p, c->zal.fctclip, c->kT.p ); ++c->launches;
    k_koz_node2<<< gn, NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->ntet, c->kbase.p, c->kinc.p, c->kT.p, c->zUL.p, c->zP.p, c->zQ.p ); ++c->launches;
    k_koz_elem3<<< ge, 128, 0, s >>>( c->ntet, c->NP, c->ktet.p, c->U.p, c->X.p, c->zQ.p, c->zal.fctdif, c->zal.fctsys_mask, c->kT.p ); ++c->launches;
    k_koz_node3<<< gn, NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->ntet, c->kbase.p, c->kinc.p, c->kT.p, c->zUL.p, c->vol.p, c->Un.p, c->W.p ); ++c->launches;
  } else {
    koz_pass1( c, dt, 0 );
    k_koz_nofct<<< nblk( c->npoin, 256 ), 256, 0, s >>>( c->npoin, c->NP, c->R.p, c->vol.p, c->U.p, dt, c->Un.p, c->W.p ); ++c->launches;
  }
  std::swap( c->U.p, c->Un.p );
  do_bc( c );
  CK( cudaGetLastError() );
  API_END
}

// ---- linear solver ------------------------------------------------------------------
static void cg_halo( xyst_ctx* c, double* v, int average )
{
  if (!(c->nsh > 0 && c->comm)) return;
  int w = (int)c->cg_ncomp;
  size_t n = c->nsh*(size_t)w;
  k_cg_shared_get<<< nblk( n, 256 ), 256, 0, c->stream >>>( (int)c->nsh, w, c->sh_node.p, v, c->sh_part.p ); ++c->launches;
  exchange( c, w );
  exchange_wait( c );
  k_cg_shared_put<<< nblk( n, 256 ), 256, 0, c->stream >>>( (int)c->nsh, w, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p,
    c->sh_part.p, c->sh_recvbuf.p, c->cg_cnt.p, average, v ); ++c->launches;
}

// local partial sums -> scal[8..8+nv), all-reduced over the communicator
static void cg_reduce( xyst_ctx* c, int nv, int nb )
{
  double* out = c->cg_scal.p + 8;
  if (nv == 1) k_reduce_final< 1, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, out );
  else         k_reduce_final< 2, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, out );
  ++c->launches;
  if (c->comm) NK( g_nccl.AllReduce( out, out, (size_t)nv, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream ) );
}

// y = A x with the Dirichlet conditions of the current solve, if any
static void cg_spmv( xyst_ctx* c, const double* x, double* y )
{
  unsigned g = nblk( c->cg_nslice*32, 256 );
  if (c->cg_hasbc)
    k_spmv_bc<<< g, 256, 0, c->stream >>>( c->cg_nrow, c->cg_base.p, c->cg_col.p, c->cg_val.p, c->cg_bc.p, c->cg_cnt.p, x, y );
  else
    k_spmv<<< g, 256, 0, c->stream >>>( c->cg_nrow, c->cg_base.p, c->cg_col.p, c->cg_val.p, x, y );
  ++c->launches;
}

// ConjugateGradients::setup :105-126 with x and b already on the device (cg_x, cg_b)
static void cg_setup_dev( xyst_ctx* c, int pc )
{
  size_t n = c->cg_nrow;
  auto s = c->stream;
  CK( cudaMemsetAsync( c->cg_scal.p, 0, 16*sizeof(double), s ) );
  c->cg_converged = false; c->cg_finished = false;
  // residual(): r = A x (own) summed over sharers; pc(): q = 1/count or diag(A), summed
  cg_spmv( c, c->cg_x.p, c->cg_r.p );
  if (pc == 0) { k_cg_inv<<< nblk( n, 256 ), 256, 0, s >>>( n, c->cg_cnt.p, c->cg_d.p ); ++c->launches; }
  else if (pc == 1) {
    CK( cudaMemcpyAsync( c->cg_d.p, c->cg_diag.p, n*sizeof(double), cudaMemcpyDeviceToDevice, s ) );
    if (c->cg_hasbc) { k_cg_bc_diag<<< nblk( n, 256 ), 256, 0, s >>>( n, c->cg_bc.p, c->cg_cnt.p, c->cg_d.p ); ++c->launches; }
  }
  else throw std::runtime_error( "unknown preconditioner" );
  cg_halo( c, c->cg_r.p, 0 );
  cg_halo( c, c->cg_d.p, 0 );
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( n, RED_THREADS ) );
  // normb = sqrt((b,b))
  k_cg_dot< 1 ><<< nb, RED_THREADS, 0, s >>>( n, c->cg_mask.p, c->cg_b.p, c->cg_b.p, nullptr, nullptr, c->red.p ); ++c->launches;
  cg_reduce( c, 1, nb );
  CK( cudaMemcpyAsync( c->red_host, c->cg_scal.p + 8, sizeof(double), cudaMemcpyDeviceToHost, s ) );
  CK( cudaStreamSynchronize( s ) );
  c->cg_normb = std::sqrt( c->red_host[0] );
  // initres(): r = b - r, p = r, z = r/d, rho = (r,z)
  k_cg_resid<<< nblk( n, 256 ), 256, 0, s >>>( n, c->cg_b.p, c->cg_r.p, c->cg_p.p ); ++c->launches;
  k_cg_div<<< nblk( n, 256 ), 256, 0, s >>>( n, c->cg_r.p, c->cg_d.p, c->cg_z.p ); ++c->launches;
  k_cg_dot< 1 ><<< nb, RED_THREADS, 0, s >>>( n, c->cg_mask.p, c->cg_r.p, c->cg_z.p, nullptr, nullptr, c->red.p ); ++c->launches;
  cg_reduce( c, 1, nb );
  CK( cudaMemcpyAsync( c->cg_scal.p, c->cg_scal.p + 8, sizeof(double), cudaMemcpyDeviceToDevice, s ) );   // rho
  CK( cudaGetLastError() );
  CK( cudaStreamSynchronize( s ) );
}

int xyst_csr_upload( xyst_ctx* c, size_t nrow, size_t ncomp, const size_t* ia, const size_t* ja, const double* a )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!nrow || !ncomp || nrow % ncomp) throw std::runtime_error( "csr_upload: bad sizes" );
  if (nrow > 0x7fffffffULL) throw std::runtime_error( "csr_upload: too many rows" );
  size_t nslice = (nrow + 31) / 32;
  std::vector< long long > base( nslice+1, 0 );
  for (size_t s=0; s<nslice; ++s) {
    size_t km = 0;
    for (size_t r=s*32; r<std::min( nrow, s*32+32 ); ++r) km = std::max( km, ia[r+1]-ia[r] );
    base[s+1] = base[s] + (long long)km*32;
  }
  size_t nent = (size_t)base[nslice];
  std::vector< int > col( nent ); std::vector< double > val( nent, 0.0 ), diag( nrow, 0.0 );
  for (size_t s=0; s<nslice; ++s)
    for (size_t j=(size_t)base[s]; j<(size_t)base[s+1]; ++j) col[j] = (int)std::min( nrow-1, s*32 + (j-(size_t)base[s])%32 );
  for (size_t r=0; r<nrow; ++r)
    for (size_t j=ia[r]-1, k=0; j<ia[r+1]-1; ++j, ++k) {
      if (ja[j] < 1 || ja[j] > nrow) throw std::runtime_error( "csr_upload: column out of range" );
      size_t slot = (size_t)base[r/32] + k*32 + r%32;
      col[slot] = (int)(ja[j]-1); val[slot] = a[j];
      if (ja[j]-1 == r) diag[r] = a[j];
    }
  auto s = c->stream;
  c->cg_nrow = nrow; c->cg_ncomp = ncomp; c->cg_nslice = nslice; c->cg_nent = nent;
  c->cg_base.upload( base, s ); c->cg_col.upload( col, s ); c->cg_val.upload( val, s ); c->cg_diag.upload( diag, s );
  for (auto* v : { &c->cg_x, &c->cg_b, &c->cg_r, &c->cg_p, &c->cg_q, &c->cg_z, &c->cg_d, &c->cg_mask, &c->cg_cnt }) v->alloc( nrow );
  c->cg_scal.alloc( 16 );
  // one partition until xyst_cg_setup says otherwise: every row counted once; x = 0
  c->cg_mask.upload( std::vector< double >( nrow, 1.0 ), s ); c->cg_cnt.upload( std::vector< double >( nrow, 1.0 ), s );
  CK( cudaMemsetAsync( c->cg_x.p, 0, nrow*sizeof(double), s ) );
  c->cg_hasbc = false;
  if (c->nsh) {           // halo buffers wide enough for ncomp values per shared node
    size_t w = std::max< size_t >( 15, ncomp );
    c->sh_part.alloc( c->nsh*w ); c->sh_sendbuf.alloc( c->nsend*w ); c->sh_recvbuf.alloc( c->nsend*w );
  }
  API_END
}

int xyst_csr_mult( xyst_ctx* c, const double* x, double* r )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!c->cg_nrow) throw std::runtime_error( "no matrix uploaded" );
  size_t n = c->cg_nrow;
  CK( cudaMemcpyAsync( c->cg_p.p, x, n*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  k_spmv<<< nblk( c->cg_nslice*32, 256 ), 256, 0, c->stream >>>( n, c->cg_base.p, c->cg_col.p, c->cg_val.p, c->cg_p.p, c->cg_q.p ); ++c->launches;
  CK( cudaGetLastError() );
  CK( cudaMemcpyAsync( r, c->cg_q.p, n*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

int xyst_cg_setup( xyst_ctx* c, const double* x, const double* b, int pc, const uint8_t* slave,
                   const double* count, double* normb )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!c->cg_nrow) throw std::runtime_error( "no matrix uploaded" );
  size_t n = c->cg_nrow, nc = c->cg_ncomp, np = n/nc;
  auto s = c->stream;
  std::vector< double > mask( n, 1.0 ), cnt( n, 1.0 );
  for (size_t i=0; i<np; ++i) for (size_t k=0; k<nc; ++k) {
    if (slave && slave[i]) mask[i*nc+k] = 0.0;
    if (count) cnt[i*nc+k] = count[i];
  }
  c->cg_mask.upload( mask, s ); c->cg_cnt.upload( cnt, s );
  CK( cudaMemcpyAsync( c->cg_x.p, x, n*sizeof(double), cudaMemcpyHostToDevice, s ) );
  CK( cudaMemcpyAsync( c->cg_b.p, b, n*sizeof(double), cudaMemcpyHostToDevice, s ) );
  c->cg_hasbc = false;
  cg_setup_dev( c, pc );
  if (normb) *normb = c->cg_normb;
  API_END
}

int xyst_cg_solve( xyst_ctx* c, size_t maxit, double tol, size_t* it_out, double* normr_out )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!c->cg_nrow) throw std::runtime_error( "no matrix uploaded" );
  size_t n = c->cg_nrow, it = 0;
  auto s = c->stream;
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( n, RED_THREADS ) );
  double normr = 0.0;
  if (!c->cg_converged) {
    // beta = 0 on the first pass (next :586)
    CK( cudaMemsetAsync( c->cg_scal.p + 3, 0, sizeof(double), s ) );
    for (;;) {
      k_cg_p<<< nblk( n, 256 ), 256, 0, s >>>( n, c->cg_scal.p, c->cg_z.p, c->cg_p.p ); ++c->launches;
      { ProfScope ps( c, "spmv" );
        cg_spmv( c, c->cg_p.p, c->cg_q.p ); }
      cg_halo( c, c->cg_q.p, 0 );
      k_cg_dot< 1 ><<< nb, RED_THREADS, 0, s >>>( n, c->cg_mask.p, c->cg_p.p, c->cg_q.p, nullptr, nullptr, c->red.p ); ++c->launches;
      cg_reduce( c, 1, nb );
      k_cg_scalars<<< 1, 32, 0, s >>>( 0, c->cg_scal.p ); ++c->launches;
      k_cg_update<<< nb, RED_THREADS, 0, s >>>( n, c->cg_scal.p, c->cg_mask.p, c->cg_q.p, c->cg_d.p, c->cg_p.p,
        c->cg_r.p, c->cg_z.p, c->cg_x.p, c->red.p ); ++c->launches;
      cg_reduce( c, 2, nb );
      k_cg_scalars<<< 1, 32, 0, s >>>( 1, c->cg_scal.p ); ++c->launches;
      cg_halo( c, c->cg_x.p, 1 );
      CK( cudaMemcpyAsync( c->red_host, c->cg_scal.p + 4, 2*sizeof(double), cudaMemcpyDeviceToHost, s ) );
      CK( cudaStreamSynchronize( s ) );
      ++it;
      double nbn = c->cg_normb > 1.0e-14 ? c->cg_normb : 1.0;
      normr = std::sqrt( c->red_host[0] );
      bool finished = c->red_host[1] != 0.0;
      if (finished || normr < tol*nbn || it >= maxit) { c->cg_converged = !(normr > tol*nbn); break; }
    }
    CK( cudaGetLastError() );
  }
  if (it_out) *it_out = it;
  if (normr_out) *normr_out = normr;
  API_END
}

int xyst_cg_get_x( xyst_ctx* c, double* x )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!c->cg_nrow) throw std::runtime_error( "no matrix uploaded" );
  CK( cudaMemcpyAsync( x, c->cg_x.p, c->cg_nrow*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

#define XYST_CHOCG_API
#include "chocg.cuh"
#undef XYST_CHOCG_API

uint64_t xyst_launch_count( const xyst_ctx* c ) { return c ? c->launches : 0; }
uint64_t xyst_nedge( const xyst_ctx* c ) { return c ? c->nedge : 0; }

int xyst_kernel_time( xyst_ctx* c, const char* kernel, int reset, double* ms, uint64_t* launches )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  c->prof_on = true;
  auto& p = c->prof[ kernel ];
  CK( cudaStreamSynchronize( c->stream ) );
  for (auto& e : p.ev) {
    float t = 0; CK( cudaEventElapsedTime( &t, e.first, e.second ) );
    p.ms += t; ++p.n;
    cudaEventDestroy( e.first ); cudaEventDestroy( e.second );
  }
  p.ev.clear();
  if (ms) *ms = p.ms;
  if (launches) *launches = p.n;
  if (reset) { p.ms = 0.0; p.n = 0; }
  API_END
}

} // extern "C"

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
#include "layout.hpp"

namespace {

thread_local std::string g_err;

int fail( const std::string& m ) { g_err = m; return 1; }

#define CK( call ) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  throw std::runtime_error( std::string( #call ) + ": " + cudaGetErrorString( e_ ) ); } while (0)

#define API_BEGIN try {
#define API_END } catch (std::exception& e) { return fail( e.what() ); } return 0;

// tuning knobs (compile-time; defaults chosen from the ncu measurements under profiles/)
#ifndef NODE_THREADS
#define NODE_THREADS 256
#endif
#ifndef GRAD_THREADS
#define GRAD_THREADS 128   // 128 registers per thread: the fully unrolled incidence loop without spills
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
  DevBuf() = default;
  DevBuf( const DevBuf& ) = delete;
  DevBuf& operator=( const DevBuf& ) = delete;
  DevBuf( DevBuf&& o ) noexcept : p( o.p ), n( o.n ) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=( DevBuf&& o ) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  void alloc( size_t m ) { release(); n = m; if (m) CK( cudaMalloc( &p, m*sizeof(T) ) ); }
  void upload( const std::vector< T >& h, cudaStream_t s ) {
    alloc( h.size() );
    if (!h.empty()) { CK( cudaMemcpyAsync( p, h.data(), h.size()*sizeof(T), cudaMemcpyHostToDevice, s ) );
                      CK( cudaStreamSynchronize( s ) ); }
  }
  void release() { if (p) cudaFree( p ); p = nullptr; n = 0; }
  ~DevBuf() { release(); }
};

// per-kernel timing on request (xyst_kernel_time): events come from a pool created when a name is
// first asked for, so a timed loop does not create events; names nobody asked for record nothing
struct Prof {
  std::vector< std::pair< cudaEvent_t, cudaEvent_t > > ev;   // pool
  size_t used = 0;
  double ms = 0.0; uint64_t n = 0;
  bool on = false;
};

} // namespace

// One linear solver (the selected one is the base of the context, so that c->cg_* names it)
struct CgState {
  // solve BCs (matrix-free: masked rows/columns), Neumann part, rhs override
  DevBuf< unsigned char > cg_bc;
  DevBuf< double > cg_bcval, cg_neu, cg_rhs0, cg_bcsmall;
  DevBuf< int > cg_bcnode; std::vector< size_t > cg_bcnodes_h;
  bool cg_hasbc = false, cg_hasneu = false, cg_hasrhs0 = false;
  // sliced-ELL matrix over scalar rows + CG vectors
  size_t cg_nrow = 0, cg_ncomp = 1, cg_nslice = 0, cg_nent = 0;
  DevBuf< long long > cg_base; DevBuf< int > cg_col; DevBuf< double > cg_val, cg_diag;
  DevBuf< double > cg_x, cg_b, cg_r, cg_p, cg_q, cg_z, cg_d, cg_mask, cg_cnt, cg_scal;
  double cg_normb = 0.0; bool cg_converged = false, cg_finished = false;
};

struct xyst_ctx : CgState {
  int device = 0;
  cudaStream_t stream = nullptr, comm_stream = nullptr, aux_stream = nullptr;
  bool own_stream = false;
  xyst_params prm{};
  size_t npoin = 0, NP = 0, nedge = 0, nslot = 0, ntri = 0, nslice = 0, nent = 0;
  uint64_t launches = 0;
  // nodal state, structure of arrays with stride NP (npoin rounded up to 32)
  DevBuf< double > U, Un, W, X, G, vol, v;   // [5][NP] [5][NP] [5][NP] [3][NP] [15][NP]
  DevBuf< double > cvol;                     // cbrt(vol), the length scale of the time step
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
  std::vector< int > sh_old_h;           // the unique shared nodes in the caller's numbering (ascending)
  std::vector< double > sh_cnt_h;        // per shared node: partitions contributing to it (own included)
  std::vector< uint8_t > sh_slave_h;     // per shared node: counted by a higher rank in global dot products
  std::vector< int > sh_node_h; DevBuf< unsigned char > sh_flag;   // host copy (library numbering); [npoin] 1 = shared (built on first use)
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
  DevBuf< double > kUE, ksUL, ksP, ksQ;  // transported scalars in KozCG / ZalCG: element half-step flow [4][ntet], per scalar ul, P+/-, Q+/-
  DevBuf< double > zSn, zSe;             // ZalCG source term: at the nodes [npoin][5], at the edge midpoints [5][nslot]
  bool zsrc = false;
  bool koz_frozen = false;               // KozCG::m_freezeflow > 1: only the scalars advance
  bool ksrc = false;
  std::vector< int > kperm;              // device tet order -> caller's tet index
  // ChoCG: velocity (3 rotating buffers: time level n, current, next), pressure, divergence,
  // gradients of the CG solution / pressure / velocity, momentum flux; BC lists
  bool cho = false;
  xyst_chocg_params chp{};
  DevBuf< double > cUa, cUb, cUc, cP, cDiv, cSg, cPg, cVg, cFl, cR, cS;
  double *cU = nullptr, *cUn = nullptr, *cUx = nullptr;   // current, time level n, scratch (point into cUa/b/c)
  int cns = 0; double cdif = 0.0;          // transported scalars of ChoCG/LohCG: rows after the velocity rows; diffusivity
  DevBuf< int > cpin; size_t ncpin = 0; double cpin_val = 1.0;   // point-source nodes (problems::point_src)
  DevBuf< int > cb_dnode, cb_dmask, cb_snode, cb_soff, cb_nnode;
  DevBuf< double > cb_dval, cb_snorm;
  size_t cb_nd = 0, cb_ns = 0, cb_nn = 0;
  // LohCG (lohcg.cuh): (p,u,v,w) in cUa/b/c as [4][NP] with cU/cUn/cUx pointing at the velocity rows,
  // sound speed, gradients of all unknowns, 4-component Dirichlet list, pressure Dirichlet list
  bool loh = false;
  double loh_s = 1.0;
  DevBuf< double > lG, lb_dval, lp_val;
  DevBuf< int > lb_dnode, lb_dmask, lp_node;
  size_t lb_nd = 0, lp_n = 0;
  // the linear solvers not selected at the moment (xyst_cg_select): [0] pressure, [1] momentum
  CgState cg_other[2]; int cg_sel = 0;
  // internal node order (locality.hpp): empty = the caller's order
  std::vector< int > new2old_h, old2new_h; DevBuf< int > new2old;
  int tile_nodes = 256;
  // owner-slot view of the edges for the thread-per-owner kernels: slot base per slice, and per slot
  // the edge's other end | orientation bit (31: the owner is the edge's SECOND node), -1 = padding
  DevBuf< long long > ebase; DevBuf< int > eo;
  // transported scalars (riecg_scalar.cuh): ncomp = 5 + ns
  int ncomp = NC, ns = 0;
  DevBuf< double > sU, sUn, sG, sF, EV, sGb, sRb, sS, sdir_val;   // [ns][NP] x2, [3ns][NP], [ns][nslot], [3][nslot], ...
  DevBuf< int > sdir_mask, spin;         // scalar Dirichlet masks; nodes of the point source
  size_t nspin = 0; double spin_val = 1.0;
  bool s_src = false;
  // owner's share of the nodal flux sums (k_flux_own) and the incoming-edge lists of k_update_in
  DevBuf< double > Racc; DevBuf< long long > in_base; DevBuf< int > in_e;
  bool gradp_attr = false;
  int grad_mode = 1, grad_waves = 1;     // 1: persistent gradient kernel with incidence prefetch, 0: one warp per slice
  // profiling
  bool prof_on = false;
  std::map< std::string, Prof > prof;
};

namespace {

#include "riecg_kernels.cuh"
#include "riecg_own.cuh"
#include "riecg_scalar.cuh"
#include "zalcg_kernels.cuh"
#include "kozcg_kernels.cuh"
#include "cg_kernels.cuh"
#define XYST_CHOCG_KERNELS
#include "chocg.cuh"
#undef XYST_CHOCG_KERNELS
#define XYST_LOHCG_KERNELS
#include "lohcg.cuh"
#undef XYST_LOHCG_KERNELS

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
inline unsigned nblk( size_t n, int t ) { return (unsigned)((n + (size_t)t - 1) / (size_t)t); }

struct ProfScope {
  xyst_ctx* c; Prof* pr = nullptr; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope( xyst_ctx* ctx, const char* name ) : c( ctx ) {
    if (!c->prof_on) return;
    auto it = c->prof.find( name );
    if (it == c->prof.end() || !it->second.on) return;
    pr = &it->second;
    if (pr->used == pr->ev.size()) {          // pool exhausted between two queries: grow (rare)
      cudaEvent_t x, y; CK( cudaEventCreate( &x ) ); CK( cudaEventCreate( &y ) );
      pr->ev.emplace_back( x, y );
    }
    a = pr->ev[pr->used].first; b = pr->ev[pr->used].second;
    CK( cudaEventRecord( a, c->stream ) );
  }
  ~ProfScope() {
    if (!pr) return;
    cudaEventRecord( b, c->stream );
    ++pr->used;
  }
};

DParams dparams( const xyst_ctx* c ) {
  return DParams{ c->prm.gamma, c->prm.stab2coef, c->prm.flux, c->prm.stab2, c->prm.exact_muscl, c->rgas, c->kvinf };
}
Mode mode( const xyst_ctx* c ) { return Mode{ c->prm.gamma, c->lax ? c->rgas : 0.0, c->kvinf }; }

// caller's node id -> internal id (identity unless the mesh upload re-ordered the nodes)
inline bool reordered( const xyst_ctx* c ) { return !c->old2new_h.empty(); }
inline size_t to_new( const xyst_ctx* c, size_t old, const char* what ) {
  if (old >= c->npoin) throw std::runtime_error( std::string( what ) + " out of range" );
  return reordered( c ) ? (size_t)c->old2new_h[old] : old;
}
// nodal array [npoin][w] in the caller's order -> internal order
inline std::vector< double > rows_to_new( const xyst_ctx* c, const double* a, size_t w ) {
  std::vector< double > r( c->npoin*w );
  if (!reordered( c )) { std::copy( a, a + c->npoin*w, r.begin() ); return r; }
  for (size_t i=0; i<c->npoin; ++i) { const double* src = a + (size_t)c->new2old_h[i]*w; for (size_t k=0; k<w; ++k) r[i*w+k] = src[k]; }
  return r;
}

// device list of the shared nodes in the library's numbering (the mesh may arrive after the halo)
void refresh_shared( xyst_ctx* c ) {
  c->sh_node_h = c->sh_old_h;
  if (reordered( c )) for (auto& i : c->sh_node_h) i = (int)to_new( c, (size_t)i, "shared node id" );
  c->sh_flag.release();
  c->sh_node.upload( c->sh_node_h, c->stream );
}

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
    size_t smem = (size_t)2*(size_t)c->maxdeg*GRAD_THREADS*sizeof(int2);
    if (c->grad_mode == 1 && c->maxdeg > 0 && smem <= 96*1024) {    // persistent form with the incidence prefetch
      if (!c->gradp_attr) { CK( cudaFuncSetAttribute( k_grad_node_p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem ) ); c->gradp_attr = true; }
      int nsm = 148; cudaDeviceGetAttribute( &nsm, cudaDevAttrMultiProcessorCount, c->device );
      unsigned g = (unsigned)std::min< size_t >( nblk( c->nslice*32, GRAD_THREADS ), (size_t)nsm*GRAD_MINB*c->grad_waves );
      k_grad_node_p<<< g, GRAD_THREADS, smem, s >>>( c->npoin, c->NP, c->nslice, c->maxdeg, c->sl_base.p,
        c->inc_eq.p, c->D2.p, c->D.p, c->nslot, c->W.p, c->bslot.p, c->Gb.p, c->vol.p, c->G.p, overlap ? 1 : 0 );
    } else
    k_grad_node<<< nblk( c->nslice*32, GRAD_THREADS ), GRAD_THREADS, 0, s >>>( c->npoin, c->NP, c->sl_base.p,
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
  // the owner's own share was summed by the flux kernel: gather the incoming edges only
  #define UPD_IN( FU, LX ) k_update_in< FU, LX ><<< g, NODE_THREADS, 0, st >>>( c->npoin, c->NP, c->in_base.p, c->in_e.p, \
      c->Racc.p, c->F.p, c->nslot, c->bslot.p, c->Rb.p, c->S.p, c->src_mask, c->v.p, c->vol.p, Un, A, Uout, c->W.p, c->R.p, \
      c->Wn.p, c->Un.p, skip, c->ncomp, s0, s1 )
  if (fused && c->lax) UPD_IN( true, true ); else if (fused) UPD_IN( true, false ); else UPD_IN( false, false );
  #undef UPD_IN
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
        c->sh_part.p, c->sh_recvbuf.p, c->vol.p, Un, A, Uout, c->W.p, c->W.p, c->R.p, c->Wn.p, c->Un.p );
    else
      k_rhs_finish< false ><<< g, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p,
        c->sh_part.p, c->sh_recvbuf.p, c->vol.p, Un, A, Uout, c->W.p, c->W.p, c->R.p, c->Wn.p, c->Un.p );
    ++c->launches;
  }
  CK( cudaGetLastError() );
}

// thread-per-owner flux kernel: per-edge fluxes F and the owners' own shares Racc
void launch_flux( xyst_ctx* c, size_t s0, size_t s1, cudaStream_t s )
{
  if (s1 <= s0) return;
  auto P = dparams( c );
  unsigned g = nblk( (s1-s0)*32, OWN_THREADS );
  #define LAUNCH_OWN( EX, FL ) k_flux_own< EX, FL ><<< g, OWN_THREADS, 0, s >>>( s0, s1, c->NP, c->nslot, \
      c->ebase.p, c->eo.p, c->D.p, c->W.p, c->G.p, c->F.p, c->Racc.p, P, c->ns ? c->EV.p : nullptr )
  int fl = P.flux + (c->lax ? 2 : 0);
  if (P.exact) { if (fl == 0) LAUNCH_OWN( true, 0 ); else if (fl == 1) LAUNCH_OWN( true, 1 );
                 else if (fl == 2) LAUNCH_OWN( true, 2 ); else LAUNCH_OWN( true, 3 ); }
  else         { if (fl == 0) LAUNCH_OWN( false, 0 ); else if (fl == 1) LAUNCH_OWN( false, 1 );
                 else if (fl == 2) LAUNCH_OWN( false, 2 ); else LAUNCH_OWN( false, 3 ); }
  #undef LAUNCH_OWN
  ++c->launches;
}
void do_flux( xyst_ctx* c )
{
  ProfScope ps( c, "flux" );
  launch_flux( c, 0, c->nslice, c->stream );
}

// ---- transported scalars (riecg_scalar.cuh) -------------------------------------------
void scal_need( xyst_ctx* c ) {
  if (c->lax) throw std::runtime_error( "transported scalars are not implemented for LaxCG" );
}
// boundary parts + gradients of the scalars; needs the state at stage start
void scal_grad( xyst_ctx* c )
{
  scal_need( c );
  auto s = c->stream;
  if (c->nbn) { k_scal_bnd<<< nblk( c->nbn, 128 ), 128, 0, s >>>( (int)c->nbn, c->ns, c->NP, c->bn_off.p, c->bn_face.p, c->tri.p,
                  c->besym.p, c->fn.p, c->U.p, c->sU.p, c->sGb.p, c->sRb.p ); ++c->launches; }
  k_scal_grad<<< nblk( c->nslice*32, 128 ), 128, 0, s >>>( c->npoin, c->ns, c->NP, c->sl_base.p, c->inc_eq.p, c->D.p, c->nslot,
    c->sU.p, c->bslot.p, c->sGb.p, c->vol.p, c->sG.p ); ++c->launches;
  CK( cudaGetLastError() );
  if (c->nsh > 0 && c->comm)              // several partitions (RieCG::comgrad): each part is already over the full nodal volume
    for (int k=0; k<c->ns; ++k) {
      double* G = c->sG.p + (size_t)3*k*c->NP;
      unsigned g = nblk( c->nsh*(size_t)3, 256 );
      k_soa_shared_get<<< g, 256, 0, s >>>( (int)c->nsh, 3, c->NP, c->sh_node.p, G, c->sh_part.p ); ++c->launches;
      exchange( c, 3 ); exchange_wait( c );
      k_soa_shared_put<<< g, 256, 0, s >>>( (int)c->nsh, 3, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p,
        c->sh_part.p, c->sh_recvbuf.p, G ); ++c->launches;
    }
}
// scalar edge fluxes (after the flow's k_flux_own, which leaves EV) and nodal sums / update
void scal_flux_nodes( xyst_ctx* c, bool fused, int stage, double dt )
{
  auto s = c->stream;
  if (c->prm.exact_muscl) k_scal_flux< true ><<< nblk( c->nslot, 128 ), 128, 0, s >>>( c->nslot, c->ns, c->NP, c->ep.p, c->eq.p, c->X.p, c->EV.p, c->sU.p, c->sG.p, c->sF.p );
  else k_scal_flux< false ><<< nblk( c->nslot, 128 ), 128, 0, s >>>( c->nslot, c->ns, c->NP, c->ep.p, c->eq.p, c->X.p, c->EV.p, c->sU.p, c->sG.p, c->sF.p );
  ++c->launches;
  unsigned g = nblk( c->nslice*32, 128 );
  const double* S = c->s_src ? c->sS.p : nullptr;
  if (fused) {
    // stage 0: un = u without a copy, as for the flow variables
    const double* un = stage == 0 ? c->sU.p : c->sUn.p;
    double* out = stage == 0 ? c->sUn.p : c->sU.p;
    k_scal_node< true ><<< g, 128, 0, s >>>( c->npoin, c->ns, c->NP, c->sl_base.p, c->inc_e.p, c->sF.p, c->nslot, c->bslot.p,
      c->sRb.p, S, c->v.p, c->vol.p, un, rkcoef[stage]*dt, c->steady ? c->dtp.p : nullptr, rkcoef[stage], out, c->R.p, c->ncomp );
    if (c->nsh > 0 && c->comm) {          // several partitions (RieCG::comrhs): the shared nodes from the complete sums
      unsigned gs = nblk( c->nsh, 128 );
      k_scal_sh<<< gs, 128, 0, s >>>( (int)c->nsh, c->ns, c->sh_node.p, c->sl_base.p, c->inc_e.p, c->sF.p, c->nslot, c->bslot.p,
        c->sRb.p, S, c->v.p, c->sh_part.p ); ++c->launches;
      exchange( c, c->ns ); exchange_wait( c );
      k_scal_shfin< true ><<< gs, 128, 0, s >>>( (int)c->nsh, c->ns, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p,
        c->sh_recvbuf.p, c->vol.p, un, rkcoef[stage]*dt, c->steady ? c->dtp.p : nullptr, rkcoef[stage], out, c->R.p, c->ncomp ); ++c->launches;
    }
    if (stage == 0) std::swap( c->sU.p, c->sUn.p );
  } else {
    k_scal_node< false ><<< g, 128, 0, s >>>( c->npoin, c->ns, c->NP, c->sl_base.p, c->inc_e.p, c->sF.p, c->nslot, c->bslot.p,
      c->sRb.p, S, c->v.p, c->vol.p, c->sUn.p, 0.0, nullptr, 0.0, c->sU.p, c->R.p, c->ncomp );
    if (c->nsh > 0 && c->comm) {
      unsigned gs = nblk( c->nsh, 128 );
      k_scal_sh<<< gs, 128, 0, s >>>( (int)c->nsh, c->ns, c->sh_node.p, c->sl_base.p, c->inc_e.p, c->sF.p, c->nslot, c->bslot.p,
        c->sRb.p, S, c->v.p, c->sh_part.p ); ++c->launches;
      exchange( c, c->ns ); exchange_wait( c );
      k_scal_shfin< false ><<< gs, 128, 0, s >>>( (int)c->nsh, c->ns, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p,
        c->sh_recvbuf.p, c->vol.p, c->sUn.p, 0.0, nullptr, 0.0, c->sU.p, c->R.p, c->ncomp ); ++c->launches;
    }
  }
  ++c->launches;
  CK( cudaGetLastError() );
}

void do_bc( xyst_ctx* c, bool flow = true )
{
  if (c->ns && c->nspin) {        // the point source acts on the updated solution, then the BCs (RieCG.cpp:1023-1028)
    k_scal_pin<<< nblk( c->nspin, 128 ), 128, 0, c->stream >>>( (int)c->nspin, c->NP, c->spin.p, c->spin_val, c->sU.p ); ++c->launches;
  }
  if (c->ns && c->nbc && c->ndir) {
    k_scal_bc<<< nblk( c->nbc, 128 ), 128, 0, c->stream >>>( (int)c->nbc, c->ns, c->NP, c->bc_node.p, c->bc_dir.p,
      c->sdir_mask.p, c->sdir_val.p, c->sU.p ); ++c->launches;
  }
  if (!c->nbc || !flow) return;
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
  if (params->ncomp < NC || params->ncomp > NC+8) throw std::runtime_error( "ncomp must be 5 (Euler system) + at most 8 transported scalars" );
  if (params->flux != 0 && params->flux != 1) throw std::runtime_error( "Flux not configured" );
  int n = 0;
  if (cudaGetDeviceCount( &n ) != cudaSuccess || n == 0)
    throw std::runtime_error( "no CUDA device: the B200 path has no CPU fallback" );
  if (device < 0 || device >= n) throw std::runtime_error( "invalid device ordinal" );
  CK( cudaSetDevice( device ) );
  auto c = new xyst_ctx;
  c->device = device; c->prm = *params; c->ncomp = params->ncomp; c->ns = params->ncomp - NC;
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
                      const uint8_t* besym, const double* vol, const double* v, size_t stride,
                      bool allow_reorder = false )
{
  API_BEGIN
  c->dstride = (int)stride;
  CK( cudaSetDevice( c->device ) );
  if (npoin == 0 || npoin > 0x7fffffffULL) throw std::runtime_error( "npoin out of range" );
  // --- layout (layout.hpp): internal node order, edge slots, incidence, tiles ---------------
  // Everything below works in the library's own numbering; ids and nodal arrays crossing the
  // ABI are mapped at the entry points (to_new / rows_to_new). XYST_REORDER=0 keeps the caller's.
  layout::Options opt;
  // The caller's order (the reference's own locality renumbering, RieCG.cpp:82-100) is kept unless
  // XYST_REORDER=1 asks for the library's tile order (locality.hpp): on the benchmark boxes the
  // reference order coalesces the neighbour gathers better.
  opt.reorder = false;
  { const char* e = getenv( "XYST_REORDER" ); if (e && e[0] == '1') opt.reorder = allow_reorder;
    e = getenv( "XYST_TILE" ); if (e && atoi( e ) >= 32) c->tile_nodes = std::min( 256, atoi( e ) / 32 * 32 );
    e = getenv( "XYST_GRAD_MODE" ); if (e) c->grad_mode = atoi( e );
    e = getenv( "XYST_GRAD_WAVES" ); if (e && atoi( e ) > 0) c->grad_waves = atoi( e ); }
  opt.tile_nodes = (size_t)c->tile_nodes;
  for (size_t i=0; i<ntri*3; ++i) if (triinpoel[i] >= npoin) throw std::runtime_error( "node id out of range in superedge" );
  layout::Mesh M = layout::build( npoin, x, y, z, nsup, dsupedge, dsupint, stride, opt );
  c->new2old_h = M.new2old; c->old2new_h = M.old2new; c->new2old.release();
  std::vector< double > px, py, pz, pvol, pvv;
  if (!M.new2old.empty()) {
    const auto& n2o = M.new2old;
    px.resize( npoin ); py.resize( npoin ); pz.resize( npoin ); pvol.resize( npoin ); pvv.resize( npoin );
    for (size_t i=0; i<npoin; ++i) { size_t o = (size_t)n2o[i]; px[i] = x[o]; py[i] = y[o]; pz[i] = z[o]; pvol[i] = vol[o]; pvv[i] = v[o]; }
    x = px.data(); y = py.data(); z = pz.data(); vol = pvol.data(); v = pvv.data();
    c->new2old.upload( c->new2old_h, c->stream );
  }
  const size_t ne = M.ne, nslice = M.nslice, nslot = M.nslot, nent = M.nent;
  const auto &ep = M.ep, &eq = M.eq; const auto& ed = M.ed; const auto &ebase = M.ebase, &base = M.base;
  const auto &inc_e = M.inc_e, &inc_q = M.inc_q;
  c->maxdeg = M.maxdeg;
  // --- boundary faces: node -> (face, local index) CSR -------------------------------
  std::vector< int > tri( ntri*3 );
  for (size_t i=0; i<ntri*3; ++i) tri[i] = (int)M.to_new( triinpoel[i] );
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
  c->eo.upload( M.eo, s ); c->ebase.upload( ebase, s );
  c->in_base.upload( M.in_base, s ); c->in_e.upload( M.in_e, s );
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
    c->vol.upload( pv, s ); c->v.upload( pw, s ); c->X.upload( px, s );
    c->cvol.alloc( NP );
    k_cbrt<<< nblk( NP, 256 ), 256, 0, s >>>( NP, c->vol.p, c->cvol.p ); ++c->launches; CK( cudaGetLastError() ); }
  c->fn.alloc( std::max< size_t >( ntri, 1 )*3 );
  if (ntri) { k_face_normals<<< nblk( ntri, 128 ), 128, 0, s >>>( (int)ntri, NP, c->tri.p, c->X.p, c->fn.p ); ++c->launches; CK( cudaGetLastError() ); }
  if (c->nsh) refresh_shared( c );
  if (c->cho) {            // the projection solver keeps its own state (chocg.cuh)
    for (auto* b : { &c->U, &c->Un, &c->W, &c->G, &c->R, &c->stage, &c->F }) b->release();
    c->S.release(); c->src_mask = 0;
    CK( cudaStreamSynchronize( s ) );
    return 0;
  }
  c->U.alloc( NP*NC ); c->Un.alloc( NP*NC ); c->G.alloc( NP*2*NGP ); c->Racc.alloc( NP*NC );
  c->R.alloc( npoin*(size_t)c->ncomp ); c->stage.alloc( npoin*(size_t)c->ncomp ); c->F.alloc( std::max< size_t >( nslot, 1 )*NC );
  if (c->ns) {
    if (stride != 3 && stride != 4) throw std::runtime_error( "transported scalars are implemented for RieCG, ZalCG and KozCG only" );
    size_t ns = (size_t)c->ns;
    c->sU.upload( std::vector< double >( ns*NP, 0.0 ), s ); c->sUn.upload( std::vector< double >( ns*NP, 0.0 ), s );
    c->sG.upload( std::vector< double >( 3*ns*NP, 0.0 ), s );
    c->sF.alloc( ns*std::max< size_t >( nslot, 1 ) ); c->EV.alloc( 3*std::max< size_t >( nslot, 1 ) );
    c->sGb.alloc( std::max< size_t >( nbn, 1 )*3*ns ); c->sRb.alloc( std::max< size_t >( nbn, 1 )*ns );
    c->sS.release(); c->s_src = false;
  }
  { std::vector< double > one( NP*NC, 1.0 );    // a harmless state until xyst_state_set
    c->U.upload( one, s ); c->Un.upload( one, s );
    // primitives + coordinates as pairs (w0,w1) (w2,w3) (w4,x) (y,z), see load_wx
    std::vector< double > wx( NP*8, 1.0 );
    for (size_t p=0; p<npoin; ++p) { wx[(2*NP+p)*2+1] = x[p]; wx[(3*NP+p)*2] = y[p]; wx[(3*NP+p)*2+1] = z[p]; }
    c->W.upload( wx, s ); }
  CK( cudaMemsetAsync( c->G.p, 0, NP*2*NGP*sizeof(double), s ) );
  CK( cudaMemsetAsync( c->F.p, 0, std::max< size_t >( nslot, 1 )*NC*sizeof(double), s ) );
  c->S.release(); c->src_mask = 0;
  CK( cudaStreamSynchronize( s ) );
  API_END
}

int xyst_mesh_upload( xyst_ctx* c, size_t npoin, const double* x, const double* y, const double* z,
                      const size_t nsup[3], const size_t* const dsupedge[3],
                      const double* const dsupint[3], size_t ntri, const size_t* triinpoel,
                      const uint8_t* besym, const double* vol, const double* v )
{ return mesh_upload_impl( c, npoin, x, y, z, nsup, dsupedge, dsupint, ntri, triinpoel, besym, vol, v, 3, !c->cho ); }

int xyst_zalcg_mesh_upload( xyst_ctx* c, size_t npoin, const double* x, const double* y, const double* z,
                            const size_t nsup[3], const size_t* const dsupedge[3],
                            const double* const dsupint[3], size_t ntri, const size_t* triinpoel,
                            const uint8_t* besym, const double* vol, const double* v )
{ return mesh_upload_impl( c, npoin, x, y, z, nsup, dsupedge, dsupint, ntri, triinpoel, besym, vol, v, 4, !c->cho ); }

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
  if (!c->zP.p) { c->zP.alloc( c->NP*10 ); c->zQ.alloc( c->NP*10 ); c->zUL.alloc( c->NP*NC ); c->zsrc = false; c->koz_frozen = false; }
  if (c->ns && c->ksUL.n != (size_t)c->ns*c->NP) {
    size_t ns = (size_t)c->ns;
    c->ksUL.alloc( ns*c->NP ); c->ksP.alloc( 2*ns*c->NP ); c->ksQ.alloc( 2*ns*c->NP );
  }
}
bool zal_halo( const xyst_ctx* c ) { return c->nsh > 0 && c->comm; }
void zal_flux_and_bnd( xyst_ctx* c, double dt )
{
  auto s = c->stream;
  if (c->nbn) { k_bnd_rhs<<< nblk( c->nbn, 128 ), 128, 0, s >>>( (int)c->nbn, c->NP, c->bn_off.p, c->bn_face.p,
                  c->tri.p, c->besym.p, c->fn.p, c->U.p, c->Rb.p, c->prm.gamma, c->W.p, c->rgas ); ++c->launches; }
  if (c->ns && c->nbn) { k_scal_bnd<<< nblk( c->nbn, 128 ), 128, 0, s >>>( (int)c->nbn, c->ns, c->NP, c->bn_off.p, c->bn_face.p,
                  c->tri.p, c->besym.p, c->fn.p, c->U.p, c->sU.p, c->sGb.p, c->sRb.p ); ++c->launches; }
  { ProfScope ps( c, "zalflux" );
    k_zal_flux_edge<<< nblk( c->nslot, 128 ), 128, 0, s >>>( c->nslot, c->NP, c->ep.p, c->eq.p, c->D.p, c->U.p, c->X.p,
      dt, c->steady ? c->dtp.p : nullptr, dparams( c ), c->F.p, c->zsrc ? c->zSn.p : nullptr, c->ns, c->sU.p, c->sF.p ); ++c->launches; }
}
// transported scalars of ZalCG, one at a time through the three node passes; results in sUn
void zal_scalars( xyst_ctx* c, double dt, int fct )
{
  auto s = c->stream;
  size_t NP = c->NP;
  unsigned g = nblk( c->nslice*32, NODE_THREADS );
  const double* dtp = c->steady ? c->dtp.p : nullptr;
  for (int k=0; k<c->ns; ++k) {
    const double* sv = c->sU.p + (size_t)k*NP; const double* sf = c->sF.p + (size_t)k*c->nslot;
    double *sn = c->sUn.p + (size_t)k*NP, *ul = c->ksUL.p + (size_t)k*NP, *P = c->ksP.p + (size_t)2*k*NP, *Q = c->ksQ.p + (size_t)2*k*NP;
    // with several partitions each pass is followed by: own sums of the shared nodes -> exchange -> finish
    auto shared = [&]( int mode, int w ) {
      if (!zal_halo( c )) return;
      unsigned gs = nblk( c->nsh, 128 );
      k_zal_ssh<<< gs, 128, 0, s >>>( mode, (int)c->nsh, NP, c->sh_node.p, c->sl_base.p, c->inc_eq.p, c->D.p, c->nslot, sf, sv,
        ul, Q, c->bslot.p, c->sRb.p, c->ns, k, c->zal.fctdif, c->zal.fctclip, c->sh_part.p ); ++c->launches;
      exchange( c, w ); exchange_wait( c );
      k_fct_sfin<<< gs, 128, 0, s >>>( mode, (int)c->nsh, NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p,
        c->sh_recvbuf.p, sv, c->vol.p, -dt, dtp, fct, P, fct ? ul : sn, Q, sn ); ++c->launches;
    };
    k_zal_snode1<<< g, NODE_THREADS, 0, s >>>( c->npoin, NP, c->sl_base.p, c->inc_eq.p, c->D.p, c->nslot, sf, sv, c->bslot.p,
      c->sRb.p, c->ns, k, c->vol.p, dt, dtp, c->zal.fctdif, fct, P, fct ? ul : sn ); ++c->launches;
    shared( 1, 3 );
    if (!fct) continue;
    k_zal_snode2<<< g, NODE_THREADS, 0, s >>>( c->npoin, NP, c->sl_base.p, c->inc_eq.p, sv, ul, P, c->zal.fctclip, Q ); ++c->launches;
    shared( 2, 2 );
    k_zal_snode3<<< g, NODE_THREADS, 0, s >>>( c->npoin, NP, c->sl_base.p, c->inc_eq.p, c->D.p, c->nslot, sv, ul, Q, c->vol.p,
      c->zal.fctdif, sn ); ++c->launches;
    shared( 3, 1 );
  }
}
// With several partitions every pass is followed by the exchange of the shared nodes' own sums
// (ZalCG::comrhs+comaec: sums, comalw: max/min, comlim: sums; ZalCG.cpp:1023-1053,1139-1148,1297-1333,
// 1490-1499) and a kernel that finishes those nodes from the complete values.
void zal_node1( xyst_ctx* c, double dt, int fct )
{
  auto s = c->stream;
  k_zal_node1<<< nblk( c->nslice*32, NODE_THREADS ), NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->sl_base.p,
    c->inc_eq.p, c->D.p, c->nslot, c->F.p, c->U.p, c->bslot.p, c->Rb.p, c->bcof.p, c->bc_symoff.p,
    c->sym_n.p, c->vol.p, dt, c->steady ? c->dtp.p : nullptr, c->zal.fctdif, fct, c->zP.p, c->zUL.p, c->R.p,
    c->zsrc ? c->zSe.p : nullptr ); ++c->launches;
  if (!zal_halo( c )) return;
  unsigned g = nblk( c->nsh, 128 );
  k_zal_sh1<<< g, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sl_base.p, c->inc_eq.p, c->D.p, c->nslot, c->F.p,
    c->U.p, c->bslot.p, c->Rb.p, c->zal.fctdif, c->bcof.p, c->bc_symoff.p, c->sym_n.p, c->sh_part.p,
    c->zsrc ? c->zSe.p : nullptr ); ++c->launches;
  exchange( c, 15 ); exchange_wait( c );
  k_zal_fin1<<< g, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p, c->sh_recvbuf.p,
    c->U.p, c->vol.p, dt, c->steady ? c->dtp.p : nullptr, fct, c->zP.p, c->zUL.p, c->R.p ); ++c->launches;
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
  if (c->ns && (c->zal.fctsys_mask >> NC)) throw std::runtime_error( "ZalCG: fctsys over transported scalars is not implemented" );
  const bool frozen = c->koz_frozen;            // ZalCG::m_freezeflow > 1 (ZalCG.cpp:948-952, 1549, 1577-1584)
  zal_flux_and_bnd( c, dt );
  if (c->ns) zal_scalars( c, dt, c->zal.fct ? 1 : 0 );
  if (frozen) {                                 // only the scalars advance; the flow of time level n stays
    CK( cudaMemcpyAsync( c->Un.p, c->U.p, c->NP*NC*sizeof(double), cudaMemcpyDeviceToDevice, s ) );
    std::swap( c->sU.p, c->sUn.p );
    do_bc( c, false );
    CK( cudaGetLastError() );
    return 0;
  }
  if (c->zal.fct) {
    zal_node1( c, dt, 1 );
    k_zal_node2<<< g, NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->sl_base.p, c->inc_eq.p, c->U.p, c->zUL.p,
      c->zP.p, c->zal.fctclip, c->zQ.p ); ++c->launches;
    if (zal_halo( c )) {
      unsigned gs = nblk( c->nsh, 128 );
      k_zal_sh2<<< gs, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sl_base.p, c->inc_eq.p, c->U.p, c->zUL.p,
        c->zal.fctclip, c->sh_part.p ); ++c->launches;
      exchange( c, 10 ); exchange_wait( c );
      k_zal_fin2<<< gs, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p,
        c->sh_recvbuf.p, c->zUL.p, c->zP.p, c->zQ.p ); ++c->launches;
    }
    // new state into the other buffer, then swap: Un keeps the old state for the diagnostics
    k_zal_node3<<< g, NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->sl_base.p, c->inc_eq.p, c->D.p, c->nslot,
      c->U.p, c->zUL.p, c->zQ.p, c->vol.p, c->zal.fctdif, c->zal.fctsys_mask, c->Un.p, c->W.p ); ++c->launches;
    if (zal_halo( c )) {
      unsigned gs = nblk( c->nsh, 128 );
      k_zal_sh3<<< gs, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sl_base.p, c->inc_eq.p, c->D.p, c->nslot,
        c->U.p, c->zQ.p, c->zal.fctdif, c->zal.fctsys_mask, c->sh_part.p ); ++c->launches;
      exchange( c, NC ); exchange_wait( c );
      k_zal_fin3<<< gs, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p,
        c->sh_recvbuf.p, c->zUL.p, c->vol.p, c->Un.p, c->W.p ); ++c->launches;
    }
  } else {
    zal_node1( c, dt, 0 );
    k_zal_nofct<<< nblk( c->npoin, 256 ), 256, 0, s >>>( c->npoin, c->NP, c->R.p, c->vol.p, c->U.p, dt, c->steady ? c->dtp.p : nullptr, c->Un.p, c->W.p ); ++c->launches;
  }
  std::swap( c->U.p, c->Un.p );
  if (c->ns) std::swap( c->sU.p, c->sUn.p );
  do_bc( c );
  CK( cudaGetLastError() );
  API_END
}

// frozen flow in ZalCG (tag::freezeflow; ZalCG::dt :948-952, solve :1549,1577-1584): as xyst_kozcg_freeze
int xyst_zalcg_freeze( xyst_ctx* c, int on )
{
  API_BEGIN
  zal_need( c );
  if (on && !c->ns) throw std::runtime_error( "xyst_zalcg_freeze: no transported scalar in this context" );
  c->koz_frozen = on != 0;
  API_END
}

// End nodes of the device's edge slots (caller's numbering; slots of padding: SIZE_MAX both), for callers that
// evaluate something per edge (the ZalCG source term at the edge midpoints). nslot from xyst_nslot.
size_t xyst_nslot( xyst_ctx* c ) { return c ? c->nslot : 0; }
int xyst_edge_list( xyst_ctx* c, size_t* p, size_t* q )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  std::vector< int > ep( c->nslot ), eq( c->nslot );
  CK( cudaMemcpyAsync( ep.data(), c->ep.p, c->nslot*sizeof(int), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaMemcpyAsync( eq.data(), c->eq.p, c->nslot*sizeof(int), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  for (size_t i=0; i<c->nslot; ++i) {
    if (ep[i] < 0) { p[i] = q[i] = (size_t)-1; continue; }
    p[i] = reordered( c ) ? (size_t)c->new2old_h[ (size_t)ep[i] ] : (size_t)ep[i];
    q[i] = reordered( c ) ? (size_t)c->new2old_h[ (size_t)eq[i] ] : (size_t)eq[i];
  }
  API_END
}

// ZalCG source term (problems::SRC through zalesak::rhs, Zalesak.cpp:118-128,152-163): values at the nodes at t
// [npoin][ncomp] and at the midpoints of the edge slots at t + dt/2 [nslot][ncomp] (xyst_edge_list; padding
// slots ignored); NULL, NULL switches it off. Source columns of transported scalars must be zero.
int xyst_zalcg_src( xyst_ctx* c, const double* Sn, const double* Se )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  zal_need( c );
  if (!Sn && !Se) { c->zsrc = false; return 0; }
  if (!Sn || !Se) throw std::runtime_error( "xyst_zalcg_src: null argument" );
  size_t m = (size_t)c->ncomp, np = c->npoin, nslot = c->nslot;
  std::vector< double > sn( np*NC ), se( std::max< size_t >( nslot, 1 )*NC, 0.0 );
  for (size_t i=0; i<np; ++i) {
    size_t j = reordered( c ) ? to_new( c, i, "node id" ) : i;
    for (size_t k=0; k<m; ++k) {
      if (k < NC) sn[j*NC+k] = Sn[i*m+k];
      else if (Sn[i*m+k] != 0.0) throw std::runtime_error( "ZalCG: a source term of a transported scalar is not implemented" ); }
  }
  for (size_t i=0; i<nslot; ++i) for (size_t k=0; k<m; ++k) {
    if (k < NC) se[k*nslot+i] = Se[i*m+k];
    else if (Se[i*m+k] != 0.0 && std::isfinite( Se[i*m+k] )) throw std::runtime_error( "ZalCG: a source term of a transported scalar is not implemented" ); }
  c->zSn.upload( sn, c->stream ); c->zSe.upload( se, c->stream );
  CK( cudaStreamSynchronize( c->stream ) );
  c->zsrc = true;
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
  const size_t NCA = (size_t)c->ncomp, NSC = (size_t)c->ns;      // all components, transported scalars
  // union of BC nodes; per node: dirichlet slot, symmetry entries, farfield entries,
  // pressure slot -- entries keep the list order of the reference (a node repeats once
  // per side set it has a normal in, RieCG.cpp:583-596)
  if (!dirvals)                      // a mask of 1 without a value would impose density 0
    for (size_t i=0; i<ndir; ++i) for (size_t k=0; k<NCA; ++k)
      if (dirbcmasks[i*(NCA+1)+1+k] == 1) throw std::runtime_error( "xyst_bc_upload: Dirichlet masks set but no values given" );
  std::map< int, int > slot;
  std::vector< int > nodes;
  std::vector< size_t > mdir, msym, mfar, mpre;
  if (reordered( c )) {            // the lists as the library numbers the nodes
    mdir.assign( dirbcmasks, dirbcmasks + ndir*(NCA+1) );
    for (size_t i=0; i<ndir; ++i) mdir[i*(NCA+1)] = to_new( c, dirbcmasks[i*(NCA+1)], "BC node id" );
    msym.resize( nsym ); for (size_t i=0; i<nsym; ++i) msym[i] = to_new( c, symbcnodes[i], "BC node id" );
    mfar.resize( nfar ); for (size_t i=0; i<nfar; ++i) mfar[i] = to_new( c, farbcnodes[i], "BC node id" );
    mpre.resize( npre ); for (size_t i=0; i<npre; ++i) mpre[i] = to_new( c, prebcnodes[i], "BC node id" );
    dirbcmasks = mdir.data(); symbcnodes = msym.data(); farbcnodes = mfar.data(); prebcnodes = mpre.data();
  }
  auto add = [&]( size_t p ){ if (p >= c->npoin) throw std::runtime_error( "BC node id out of range" );
    auto it = slot.find( (int)p ); if (it == slot.end()) { slot[(int)p] = (int)nodes.size(); nodes.push_back( (int)p ); } };
  for (size_t i=0; i<ndir; ++i) add( dirbcmasks[i*(NCA+1)] );
  for (size_t i=0; i<nsym; ++i) add( symbcnodes[i] );
  for (size_t i=0; i<nfar; ++i) add( farbcnodes[i] );
  for (size_t i=0; i<npre; ++i) add( prebcnodes[i] );
  // sort the union by node id for locality, rebuild slots
  std::sort( nodes.begin(), nodes.end() );
  for (size_t i=0; i<nodes.size(); ++i) slot[nodes[i]] = (int)i;
  size_t nbc = nodes.size();
  std::vector< int > dir( nbc, -1 ), pre( nbc, -1 ), symoff( nbc+1, 0 ), faroff( nbc+1, 0 );
  std::vector< int > dmask( ndir*NC ), smask( ndir*NSC );
  std::vector< double > sval( ndir*NSC, 0.0 );
  std::vector< double > dval( ndir*NC, 0.0 ), symn( nsym*3 ), farn( nfar*3 ), preval( npre*2 );
  for (size_t i=0; i<ndir; ++i) {
    dir[ slot[(int)dirbcmasks[i*(NCA+1)]] ] = (int)i;
    for (int k=0; k<NC; ++k) { dmask[i*NC+k] = (int)dirbcmasks[i*(NCA+1)+1+(size_t)k]; if (dirvals) dval[i*NC+k] = dirvals[i*NCA+(size_t)k]; }
    for (size_t k=0; k<NSC; ++k) { smask[i*NSC+k] = (int)dirbcmasks[i*(NCA+1)+1+NC+k]; if (dirvals) sval[i*NSC+k] = dirvals[i*NCA+NC+k]; }
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
  if (NSC) { c->sdir_mask.upload( smask, s ); c->sdir_val.upload( sval, s ); }
  c->sym_n.upload( symn, s ); c->far_n.upload( farn, s ); c->pre_val.upload( preval, s );
  c->far_r = far_density; c->far_p = far_pressure;
  if (far_velocity) for (int j=0; j<3; ++j) c->far_u[j] = far_velocity[j];
  API_END
}

int xyst_dirbc_values( xyst_ctx* c, const double* dirvals )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (c->ndir && !dirvals) throw std::runtime_error( "xyst_dirbc_values: null values with Dirichlet nodes present" );
  if (c->ndir && !c->ns) { CK( cudaMemcpyAsync( c->dir_val.p, dirvals, c->ndir*NC*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
                 CK( cudaStreamSynchronize( c->stream ) ); }
  else if (c->ndir) {               // rows of 5 + ns values: flow and scalar parts go to their own arrays
    const size_t w = (size_t)c->ncomp, ns = (size_t)c->ns;
    std::vector< double > f( c->ndir*NC ), sc( c->ndir*ns );
    for (size_t i=0; i<c->ndir; ++i) { for (int k=0; k<NC; ++k) f[i*NC+k] = dirvals[i*w+(size_t)k]; for (size_t k=0; k<ns; ++k) sc[i*ns+k] = dirvals[i*w+NC+k]; }
    CK( cudaMemcpyAsync( c->dir_val.p, f.data(), f.size()*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
    CK( cudaMemcpyAsync( c->sdir_val.p, sc.data(), sc.size()*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
    CK( cudaStreamSynchronize( c->stream ) );
  }
  API_END
}

int xyst_src_upload( xyst_ctx* c, const double* S )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (!S) { c->S.release(); c->src_mask = 0; c->sS.release(); c->s_src = false; return 0; }
  const size_t w = (size_t)c->ncomp, ns = (size_t)c->ns;
  auto rows = rows_to_new( c, S, w );
  // flow components [npoin][5] and transported scalars [npoin][ns] separately
  std::vector< double > f( c->npoin*NC ), sc( c->npoin*ns );
  int mask = 0; bool any = false;
  for (size_t p=0; p<c->npoin; ++p) {
    for (int k=0; k<NC; ++k) { f[p*NC+k] = rows[p*w+(size_t)k]; if (f[p*NC+k] != 0.0) mask |= 1<<k; }
    for (size_t k=0; k<ns; ++k) { sc[p*ns+k] = rows[p*w+NC+k]; if (sc[p*ns+k] != 0.0) any = true; }
  }
  c->S.upload( f, c->stream );
  c->src_mask = mask;
  if (ns) { c->sS.upload( sc, c->stream ); c->s_src = any; }
  API_END
}

int xyst_scalar_pin( xyst_ctx* c, size_t n, const size_t* nodes, double value )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (!c->ns) throw std::runtime_error( "xyst_scalar_pin: no transported scalar in this context" );
  std::vector< int > h( n );
  for (size_t i=0; i<n; ++i) h[i] = (int)to_new( c, nodes[i], "point-source node id" );
  c->spin.upload( h, c->stream ); c->nspin = n; c->spin_val = value;
  API_END
}

int xyst_state_set( xyst_ctx* c, const double* U )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (c->rb_pending) { CK( cudaStreamSynchronize( c->aux_stream ) ); c->rb_pending = false; }
  CK( cudaMemcpyAsync( c->stage.p, U, c->npoin*(size_t)c->ncomp*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  k_set_state<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, c->stage.p, c->new2old.p, c->U.p, c->W.p, mode( c ),
    c->ncomp, c->sU.p );
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
  k_get_state<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, c->U.p, c->new2old.p, c->stage.p, c->ncomp, c->sU.p );
  ++c->launches;
  CK( cudaGetLastError() );
  CK( cudaMemcpyAsync( U, c->stage.p, c->npoin*(size_t)c->ncomp*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

int xyst_riecg_grad( xyst_ctx* c ) { API_BEGIN CK( cudaSetDevice( c->device ) ); do_grad( c ); if (c->ns) scal_grad( c ); API_END }

int xyst_grad_get( xyst_ctx* c, double* G )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  std::vector< double > h( c->NP*2*NGP ), hs( c->NP*3*(size_t)c->ns );
  CK( cudaMemcpyAsync( h.data(), c->G.p, h.size()*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  if (c->ns) CK( cudaMemcpyAsync( hs.data(), c->sG.p, hs.size()*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  const size_t w = 3*(size_t)c->ncomp;
  for (size_t p=0; p<c->npoin; ++p) {
    size_t o = reordered( c ) ? (size_t)c->new2old_h[p] : p;
    for (int i=0; i<15; ++i) G[o*w+(size_t)i] = h[gidx( i, p, c->NP )];
    for (size_t i=0; i<3*(size_t)c->ns; ++i) G[o*w+15+i] = hs[i*c->NP+p];
  }
  API_END
}

// Gradients given by the caller (already divided by the nodal volumes), npoin x 15: the reference's
// riemann::rhs takes G as an argument
int xyst_grad_set( xyst_ctx* c, const double* G )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (!G) throw std::runtime_error( "null argument" );
  if (c->ns) throw std::runtime_error( "xyst_grad_set: transported scalars are not supported here" );
  std::vector< double > h( c->NP*2*NGP, 0.0 );
  for (size_t p=0; p<c->npoin; ++p) {
    size_t o = reordered( c ) ? (size_t)c->new2old_h[p] : p;
    for (int i=0; i<15; ++i) h[gidx( i, p, c->NP )] = G[o*15+(size_t)i];
  }
  CK( cudaMemcpyAsync( c->G.p, h.data(), h.size()*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

// New boundary symmetry flags (RieCG::m_besym) / own nodal volumes V() for an uploaded mesh
int xyst_besym_upload( xyst_ctx* c, const uint8_t* besym )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (c->ntri && !besym) throw std::runtime_error( "null argument" );
  if (c->rb_pending) { CK( cudaStreamSynchronize( c->aux_stream ) ); c->rb_pending = false; }   // boundary fluxes of the old flags
  if (c->ntri) { CK( cudaMemcpyAsync( c->besym.p, besym, c->ntri*3, cudaMemcpyHostToDevice, c->stream ) );
                 CK( cudaStreamSynchronize( c->stream ) ); }
  API_END
}

int xyst_v_upload( xyst_ctx* c, const double* v )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (!v) throw std::runtime_error( "null argument" );
  std::vector< double > h( c->NP, 1.0 );
  for (size_t p=0; p<c->npoin; ++p) h[p] = v[ reordered( c ) ? (size_t)c->new2old_h[p] : p ];
  CK( cudaMemcpyAsync( c->v.p, h.data(), h.size()*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

int xyst_riecg_rhs( xyst_ctx* c )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  do_flux( c );
  if (c->ns) scal_flux_nodes( c, false, 0, 0.0 );
  do_rhs_nodes( c, false, 0, 0.0, c->U.p, c->Un.p, c->U.p );
  API_END
}

int xyst_rhs_get( xyst_ctx* c, double* R )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  const size_t w = (size_t)c->ncomp;
  if (!reordered( c )) {
    CK( cudaMemcpyAsync( R, c->R.p, c->npoin*w*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
    CK( cudaStreamSynchronize( c->stream ) );
  } else {
    std::vector< double > h( c->npoin*w );
    CK( cudaMemcpyAsync( h.data(), c->R.p, h.size()*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
    CK( cudaStreamSynchronize( c->stream ) );
    for (size_t p=0; p<c->npoin; ++p) for (size_t k=0; k<w; ++k) R[(size_t)c->new2old_h[p]*w+k] = h[p*w+k];
  }
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
    A, c->U.p, c->W.p, c->Wn.p, c->Un.p, c->ncomp ); ++c->launches;
  if (c->ns) {
    if (stage == 0) CK( cudaMemcpyAsync( c->sUn.p, c->sU.p, c->NP*(size_t)c->ns*sizeof(double), cudaMemcpyDeviceToDevice, c->stream ) );
    k_scal_update<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, c->ns, c->NP, c->R.p, c->ncomp, c->vol.p,
      c->sUn.p, rkcoef[stage]*dt, c->steady ? c->dtp.p : nullptr, rkcoef[stage], c->sU.p ); ++c->launches;
  }
  CK( cudaGetLastError() );
  API_END
}

int xyst_apply_bc( xyst_ctx* c ) { API_BEGIN CK( cudaSetDevice( c->device ) ); need_mesh( c ); do_bc( c ); API_END }

static int dt_min_impl( xyst_ctx* c, double cfl, double* dt, bool all )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( c->npoin, RED_THREADS ) );
  k_dt<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->U.p, c->cvol.p, c->prm.gamma, c->red.p, mode( c ), cfl,
    c->steady ? c->dtp.p : nullptr );
  double* d = c->red.p + (size_t)RED_BLOCKS*NDIAG;
  k_reduce_final< 1, true ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, d );
  c->launches += 2;
  // over all partitions: the all-reduce works on the device value, one copy back and one wait per step
  if (all && c->comm) NK( g_nccl.AllReduce( d, d, 1, NCCL_FLOAT64, NCCL_MIN, c->comm, c->stream ) );
  CK( cudaMemcpyAsync( c->red_host, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  *dt = c->red_host[0] * cfl;
  API_END
}
int xyst_dt_min( xyst_ctx* c, double cfl, double* dt ) { return dt_min_impl( c, cfl, dt, false ); }
int xyst_dt_min_all( xyst_ctx* c, double cfl, double* dt ) { return dt_min_impl( c, cfl, dt, true ); }

int xyst_riecg_stage( xyst_ctx* c, int stage, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  need_mesh( c );
  if (stage < 0 || stage > 2) throw std::runtime_error( "stage must be 0, 1 or 2" );
  do_grad( c );
  if (c->ns) scal_grad( c );
  do_flux( c );
  if (c->ns) scal_flux_nodes( c, true, stage, dt );
  auto nodes = [&]( const double* Un, double* Uout ) { do_rhs_nodes( c, true, stage, dt, c->U.p, Un, Uout ); };
  if (c->lax)           // time level n is kept in (p,u,v,w,T) form (Wn); Un is refreshed at stage 2
    nodes( c->Un.p, c->U.p );
  else if (stage == 0) { // un = u (RieCG.cpp:1011) without a copy: write the new state into the
                        // other buffer and swap the roles of the two
    nodes( c->U.p, c->Un.p );
    std::swap( c->U.p, c->Un.p );
  } else
    nodes( c->Un.p, c->U.p );
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
  const size_t w = (size_t)c->ncomp, ns = (size_t)c->ns;
  DevBuf< double > dan, dans;
  if (an) {
    auto rows = rows_to_new( c, an, w );
    if (!ns) dan.upload( rows, c->stream );
    else {
      std::vector< double > f( c->npoin*NC ), sc( c->npoin*ns );
      for (size_t p=0; p<c->npoin; ++p) { for (int k=0; k<NC; ++k) f[p*NC+k] = rows[p*w+(size_t)k]; for (size_t k=0; k<ns; ++k) sc[p*ns+k] = rows[p*w+NC+k]; }
      dan.upload( f, c->stream ); dans.upload( sc, c->stream );
    }
  }
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( c->npoin, RED_THREADS ) );
  double* res = c->red.p + (size_t)RED_BLOCKS*NDIAG;
  k_diag<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->U.p, c->Un.p, c->v.p, dan.p, c->red.p );
  k_reduce_final< NDIAG, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, res );
  c->launches += 2;
  CK( cudaMemcpyAsync( c->red_host, res, NDIAG*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  // layout of out: [0,w) sum u^2 v, [w,2w) sum (u-un)^2 v, [2w] sum u_4 v, [2w+1,3w+1) L2 error sums, [3w+1,4w+1) L1
  for (size_t i=0; i<4*w+1; ++i) out[i] = 0.0;
  for (int k=0; k<NC; ++k) {
    out[(size_t)k] = c->red_host[k]; out[w+(size_t)k] = c->red_host[NC+k];
    out[2*w+1+(size_t)k] = c->red_host[2*NC+1+k]; out[3*w+1+(size_t)k] = c->red_host[3*NC+1+k];
  }
  out[2*w] = c->red_host[2*NC];
  for (size_t k=0; k<ns; ++k) {
    k_scal_diag<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, (int)k, (int)ns, c->NP, c->sU.p, c->sUn.p, c->v.p, dans.p, c->red.p );
    k_reduce_final< 4, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, res );
    c->launches += 2;
    CK( cudaMemcpyAsync( c->red_host, res, 4*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
    CK( cudaStreamSynchronize( c->stream ) );
    out[NC+k] = c->red_host[0]; out[w+NC+k] = c->red_host[1]; out[2*w+1+NC+k] = c->red_host[2]; out[3*w+1+NC+k] = c->red_host[3];
  }
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
  // per shared node: number of partitions contributing (tk::count, Reorder.cpp:379-390) and whether a
  // sharer with a higher rank counts it in the dot products (tk::slave, :392-416)
  c->sh_cnt_h.assign( uniq.size(), 1.0 ); c->sh_slave_h.assign( uniq.size(), 0 );
  for (int k=0; k<nneigh; ++k)
    for (size_t i=neigh_off[k]; i<neigh_off[k+1]; ++i) {
      c->sh_cnt_h[ send[i] ] += 1.0;
      if (neigh_rank[k] > c->rank) c->sh_slave_h[ send[i] ] = 1;
    }
  auto s = c->stream;
  c->nsh = uniq.size(); c->nsend = nsend; c->sh_old_h = uniq;
  refresh_shared( c );
  c->sh_send.upload( send, s ); c->sh_roff.upload( roff, s ); c->sh_ridx.upload( ridx, s );
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
  double* d = c->red.p + (size_t)RED_BLOCKS*NDIAG;
  for (int o=0; o<n; o+=NDIAG) {                // in pieces of the scratch size (diagnostics with transported scalars)
    int m = std::min( NDIAG, n-o );
    CK( cudaMemcpyAsync( d, v+o, m*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
    NK( g_nccl.AllReduce( d, d, (size_t)m, NCCL_FLOAT64, op, c->comm, c->stream ) );
    CK( cudaMemcpyAsync( c->red_host, d, m*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
    CK( cudaStreamSynchronize( c->stream ) );
    for (int i=0; i<m; ++i) v[o+i] = c->red_host[i];
  }
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
// flow columns of the nodal [npoin][ncomp] and centroid [ntet][ncomp] source values (centroids in the device's
// element order); the transported scalars' source columns must be zero (kozak::rhs adds s[c] for every c:
// a non-zero scalar source is not implemented)
static void koz_src_split( xyst_ctx* c, size_t npoin, size_t ntet, const std::vector< int >& perm, const double* Sn,
                           const double* Sc, std::vector< double >& sn, std::vector< double >& sc )
{
  size_t m = (size_t)c->ncomp;
  sn.resize( npoin*NC ); sc.resize( ntet*NC );
  for (size_t i=0; i<npoin; ++i) for (size_t k=0; k<m; ++k) {
    if (k < NC) sn[i*NC+k] = Sn[i*m+k];
    else if (Sn[i*m+k] != 0.0) throw std::runtime_error( "KozCG: a source term of a transported scalar is not implemented" ); }
  for (size_t i=0; i<ntet; ++i) for (size_t k=0; k<m; ++k) {
    if (k < NC) sc[i*NC+k] = Sc[(size_t)perm[i]*m+k];
    else if (Sc[(size_t)perm[i]*m+k] != 0.0) throw std::runtime_error( "KozCG: a source term of a transported scalar is not implemented" ); }
}

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
  c->kperm = perm;
  c->ksrc = Sn && Sc;
  if (c->ksrc) {
    std::vector< double > sn, sc;
    koz_src_split( c, npoin, ntet, perm, Sn, Sc, sn, sc );
    c->S.upload( sn, s ); c->kSc.upload( sc, s );
  }
  c->zP.alloc( c->NP*10 ); c->zQ.alloc( c->NP*10 ); c->zUL.alloc( c->NP*NC );
  c->koz_frozen = false;
  if (c->ns) {
    size_t ns = (size_t)c->ns;
    c->kUE.alloc( 4*ntet ); c->ksUL.alloc( ns*c->NP ); c->ksP.alloc( 2*ns*c->NP ); c->ksQ.alloc( 2*ns*c->NP );
  }
  API_END
}

// new source term values (time-dependent problems: nodes at t, centroids at t + dt/2,
// Kozak.cpp:97-108,160-171)
int xyst_kozcg_src( xyst_ctx* c, const double* Sn, const double* Sc )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!c->ntet || c->kperm.size() != c->ntet) throw std::runtime_error( "no KozCG mesh uploaded" );
  if (!Sn || !Sc) throw std::runtime_error( "xyst_kozcg_src: null argument" );
  auto s = c->stream;
  size_t ntet = c->ntet;
  if (c->S.n != c->npoin*NC) c->S.alloc( c->npoin*NC );
  if (c->kSc.n != ntet*NC) c->kSc.alloc( ntet*NC );
  std::vector< double > sn, sc;
  koz_src_split( c, c->npoin, ntet, c->kperm, Sn, Sc, sn, sc );
  CK( cudaMemcpyAsync( c->S.p, sn.data(), c->npoin*NC*sizeof(double), cudaMemcpyHostToDevice, s ) );
  CK( cudaMemcpyAsync( c->kSc.p, sc.data(), ntet*NC*sizeof(double), cudaMemcpyHostToDevice, s ) );
  CK( cudaStreamSynchronize( s ) );
  c->ksrc = true;
  API_END
}

namespace {
void koz_need( xyst_ctx* c ) {
  need_mesh( c );
  if (!c->ntet) throw std::runtime_error( "KozCG needs xyst_kozcg_mesh_upload" );
}
void koz_pass1( xyst_ctx* c, double dt, int fct )
{
  auto s = c->stream;
  { ProfScope ps( c, "kozelem" );
    k_koz_elem1<<< nblk( c->ntet, 128 ), 128, 0, s >>>( c->ntet, c->NP, c->ktet.p, c->U.p, c->X.p,
      c->ksrc ? c->S.p : nullptr, c->kSc.p, dt, c->prm.gamma, c->zal.fctdif, fct, c->kT.p, c->ns ? c->kUE.p : nullptr ); ++c->launches; }
  k_koz_node1<<< nblk( c->nslice*32, NODE_THREADS ), NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->ntet, c->kbase.p, c->kinc.p,
    c->kT.p, c->U.p, c->bcof.p, c->bc_symoff.p, c->sym_n.p, c->vol.p, dt, fct, c->zP.p, c->zUL.p, c->R.p ); ++c->launches;
  if (zal_halo( c )) {          // comrhs + comaec
    unsigned g = nblk( c->nsh, 128 );
    k_koz_sh<<< g, 128, 0, s >>>( 1, (int)c->nsh, c->NP, c->ntet, c->sh_node.p, c->kbase.p, c->kinc.p, c->kT.p,
      c->bcof.p, c->bc_symoff.p, c->sym_n.p, c->sh_part.p ); ++c->launches;
    exchange( c, 15 ); exchange_wait( c );
    k_koz_fin1<<< g, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p, c->sh_recvbuf.p,
      c->U.p, c->vol.p, dt, fct, c->zP.p, c->zUL.p, c->R.p ); ++c->launches;
  }
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

// transported scalars of KozCG, one at a time through the same passes as the flow (kozak::rhs scalar rows,
// KozCG::aec/alw/lim/solve per component); results in sUn
static void koz_scalars( xyst_ctx* c, double dt, int fct )
{
  auto s = c->stream;
  size_t NP = c->NP, ntet = c->ntet;
  unsigned gn = nblk( c->nslice*32, NODE_THREADS ), ge = nblk( ntet, 128 );
  for (int k=0; k<c->ns; ++k) {
    const double* sv = c->sU.p + (size_t)k*NP;
    double *sn = c->sUn.p + (size_t)k*NP, *ul = c->ksUL.p + (size_t)k*NP, *P = c->ksP.p + (size_t)2*k*NP, *Q = c->ksQ.p + (size_t)2*k*NP;
    auto shared = [&]( int mode, int w ) {          // several partitions: own sums of the shared nodes -> exchange -> finish
      if (!zal_halo( c )) return;
      unsigned gs = nblk( c->nsh, 128 );
      k_koz_ssh<<< gs, 128, 0, s >>>( mode, (int)c->nsh, ntet, c->sh_node.p, c->kbase.p, c->kinc.p, c->kT.p, c->sh_part.p ); ++c->launches;
      exchange( c, w ); exchange_wait( c );
      k_fct_sfin<<< gs, 128, 0, s >>>( mode, (int)c->nsh, NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p,
        c->sh_recvbuf.p, sv, c->vol.p, dt, nullptr, fct, P, fct ? ul : sn, Q, sn ); ++c->launches;
    };
    k_koz_selem1<<< ge, 128, 0, s >>>( ntet, NP, c->ktet.p, c->U.p, sv, c->X.p, c->kUE.p, dt, c->zal.fctdif, fct, c->kT.p ); ++c->launches;
    k_koz_snode1<<< gn, NODE_THREADS, 0, s >>>( c->npoin, NP, ntet, c->kbase.p, c->kinc.p, c->kT.p, sv, c->vol.p, dt, fct, P, fct ? ul : sn ); ++c->launches;
    shared( 1, 3 );
    if (!fct) continue;
    k_koz_elem2< 1 ><<< ge, 128, 0, s >>>( ntet, NP, c->ktet.p, sv, ul, c->zal.fctclip, c->kT.p ); ++c->launches;
    k_koz_node2< 1 ><<< gn, NODE_THREADS, 0, s >>>( c->npoin, NP, ntet, c->kbase.p, c->kinc.p, c->kT.p, ul, P, Q ); ++c->launches;
    shared( 2, 2 );
    k_koz_elem3< 1 ><<< ge, 128, 0, s >>>( ntet, NP, c->ktet.p, sv, c->X.p, Q, c->zal.fctdif, 0, c->kT.p ); ++c->launches;
    k_koz_node3< 1, false ><<< gn, NODE_THREADS, 0, s >>>( c->npoin, NP, ntet, c->kbase.p, c->kinc.p, c->kT.p, ul, c->vol.p, sn, nullptr ); ++c->launches;
    shared( 3, 1 );
  }
}

// frozen flow (KozCG::dt :669-674, solve :1140-1176): only the transported scalars advance
int xyst_kozcg_freeze( xyst_ctx* c, int on )
{
  API_BEGIN
  koz_need( c );
  if (on && !c->ns) throw std::runtime_error( "xyst_kozcg_freeze: no transported scalar in this context" );
  c->koz_frozen = on != 0;
  API_END
}

int xyst_kozcg_step( xyst_ctx* c, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  koz_need( c );
  auto s = c->stream;
  unsigned gn = nblk( c->nslice*32, NODE_THREADS ), ge = nblk( c->ntet, 128 );
  if (c->ns && (c->zal.fctsys_mask >> NC)) throw std::runtime_error( "KozCG: fctsys over transported scalars is not implemented" );
  const bool frozen = c->koz_frozen;
  if (c->zal.fct) {
    koz_pass1( c, dt, 1 );
    if (c->ns) koz_scalars( c, dt, 1 );              // (the per-tet buffer is free again after the flow's node pass 1)
    if (!frozen) {
    k_koz_elem2< NC ><<< ge, 128, 0, s >>>( c->ntet, c->NP, c->ktet.p, c->U.p, c->zUL.p, c->zal.fctclip, c->kT.p ); ++c->launches;
    k_koz_node2< NC ><<< gn, NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->ntet, c->kbase.p, c->kinc.p, c->kT.p, c->zUL.p, c->zP.p, c->zQ.p ); ++c->launches;
    if (zal_halo( c )) {        // comalw: max / min over the sharers, then the limit coefficients as in ZalCG
      unsigned gs = nblk( c->nsh, 128 );
      k_koz_sh<<< gs, 128, 0, s >>>( 2, (int)c->nsh, c->NP, c->ntet, c->sh_node.p, c->kbase.p, c->kinc.p, c->kT.p,
        c->bcof.p, c->bc_symoff.p, c->sym_n.p, c->sh_part.p ); ++c->launches;
      exchange( c, 10 ); exchange_wait( c );
      k_zal_fin2<<< gs, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p,
        c->sh_recvbuf.p, c->zUL.p, c->zP.p, c->zQ.p ); ++c->launches;
    }
    k_koz_elem3< NC ><<< ge, 128, 0, s >>>( c->ntet, c->NP, c->ktet.p, c->U.p, c->X.p, c->zQ.p, c->zal.fctdif, c->zal.fctsys_mask, c->kT.p ); ++c->launches;
    k_koz_node3< NC, true ><<< gn, NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->ntet, c->kbase.p, c->kinc.p, c->kT.p, c->zUL.p, c->vol.p, c->Un.p, c->W.p ); ++c->launches;
    if (zal_halo( c )) {        // comlim
      unsigned gs = nblk( c->nsh, 128 );
      k_koz_sh<<< gs, 128, 0, s >>>( 3, (int)c->nsh, c->NP, c->ntet, c->sh_node.p, c->kbase.p, c->kinc.p, c->kT.p,
        c->bcof.p, c->bc_symoff.p, c->sym_n.p, c->sh_part.p ); ++c->launches;
      exchange( c, NC ); exchange_wait( c );
      k_zal_fin3<<< gs, 128, 0, s >>>( (int)c->nsh, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p, c->sh_part.p,
        c->sh_recvbuf.p, c->zUL.p, c->vol.p, c->Un.p, c->W.p ); ++c->launches;
    }
    }
  } else {
    koz_pass1( c, dt, 0 );
    if (c->ns) koz_scalars( c, dt, 0 );
    if (!frozen) { k_koz_nofct<<< nblk( c->npoin, 256 ), 256, 0, s >>>( c->npoin, c->NP, c->R.p, c->vol.p, c->U.p, dt, c->Un.p, c->W.p ); ++c->launches; }
  }
  if (frozen)          // the flow of time level n stays (and is its own previous state for the diagnostics)
    CK( cudaMemcpyAsync( c->Un.p, c->U.p, c->NP*NC*sizeof(double), cudaMemcpyDeviceToDevice, s ) );
  else
    std::swap( c->U.p, c->Un.p );
  if (c->ns) std::swap( c->sU.p, c->sUn.p );
  do_bc( c, !frozen );
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
  // (after the device-side stop flag is set the partial sums are stale: k_cg_scalars ignores them)
  double* out = c->cg_scal.p + 8;
  if (nv == 1) k_reduce_final< 1, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, out );
  else         k_reduce_final< 2, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, out );
  ++c->launches;
  if (c->comm) NK( g_nccl.AllReduce( out, out, (size_t)nv, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream ) );
}

// y = A x with the Dirichlet conditions of the current solve, if any
static void cg_spmv( xyst_ctx* c, const double* x, double* y, const double* done = nullptr )
{
  unsigned g = nblk( c->cg_nslice*32, 256 );
  if (c->cg_hasbc)
    k_spmv_bc<<< g, 256, 0, c->stream >>>( c->cg_nrow, c->cg_base.p, c->cg_col.p, c->cg_val.p, c->cg_bc.p, c->cg_cnt.p, x, y, done );
  else
    k_spmv<<< g, 256, 0, c->stream >>>( c->cg_nrow, c->cg_base.p, c->cg_col.p, c->cg_val.p, x, y, done );
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
  k_cg_dot< 1 ><<< nb, RED_THREADS, 0, s >>>( n, c->cg_mask.p, c->cg_b.p, c->cg_b.p, nullptr, nullptr, c->red.p, nullptr ); ++c->launches;
  cg_reduce( c, 1, nb );
  CK( cudaMemcpyAsync( c->red_host, c->cg_scal.p + 8, sizeof(double), cudaMemcpyDeviceToHost, s ) );
  CK( cudaStreamSynchronize( s ) );
  c->cg_normb = std::sqrt( c->red_host[0] );
  // initres(): r = b - r, p = r, z = r/d, rho = (r,z)
  k_cg_resid<<< nblk( n, 256 ), 256, 0, s >>>( n, c->cg_b.p, c->cg_r.p, c->cg_p.p ); ++c->launches;
  k_cg_div<<< nblk( n, 256 ), 256, 0, s >>>( n, c->cg_r.p, c->cg_d.p, c->cg_z.p ); ++c->launches;
  k_cg_dot< 1 ><<< nb, RED_THREADS, 0, s >>>( n, c->cg_mask.p, c->cg_r.p, c->cg_z.p, nullptr, nullptr, c->red.p, nullptr ); ++c->launches;
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
  { std::vector< double > mask( nrow, 1.0 ), cnt( nrow, 1.0 );
    if (c->nsh && c->comm && c->sh_cnt_h.size() == c->nsh)     // partitioned mesh with one row block per node
      for (size_t i=0; i<c->nsh; ++i) {
        size_t p = (size_t)c->sh_old_h[i];
        if ((p+1)*ncomp > nrow) continue;
        for (size_t k=0; k<ncomp; ++k) { cnt[p*ncomp+k] = c->sh_cnt_h[i]; if (c->sh_slave_h[i]) mask[p*ncomp+k] = 0.0; }
      }
    c->cg_mask.upload( mask, s ); c->cg_cnt.upload( cnt, s ); }
  CK( cudaMemsetAsync( c->cg_x.p, 0, nrow*sizeof(double), s ) );
  c->cg_hasbc = false;
  if (c->nsh) {           // halo buffers wide enough for ncomp values per shared node
    size_t w = std::max< size_t >( 15, ncomp );
    c->sh_part.alloc( c->nsh*w ); c->sh_sendbuf.alloc( c->nsend*w ); c->sh_recvbuf.alloc( c->nsend*w );
  }
  API_END
}

// the context holds two linear solvers (ChoCG with theta > 0: ChoCG::m_cgpre and m_cgmom,
// ChoCG.cpp:126-140); all xyst_csr_* / xyst_cg_* entries act on the selected one
int xyst_cg_select( xyst_ctx* c, int which )
{
  API_BEGIN
  if (which < 0 || which > 1) throw std::runtime_error( "xyst_cg_select: which must be 0 (pressure) or 1 (momentum)" );
  if (which != c->cg_sel) {
    CgState& cur = *c;
    c->cg_other[c->cg_sel] = std::move( cur );
    cur = std::move( c->cg_other[which] );
    c->cg_sel = which;
  }
  API_END
}

// new values for the matrix of xyst_csr_upload (same ia/ja); the solution vector is kept
int xyst_csr_update( xyst_ctx* c, const size_t* ia, const double* a )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!c->cg_nrow) throw std::runtime_error( "no matrix uploaded" );
  size_t nrow = c->cg_nrow, nslice = c->cg_nslice;
  std::vector< long long > base( nslice+1, 0 );
  for (size_t s=0; s<nslice; ++s) {
    size_t km = 0;
    for (size_t r=s*32; r<std::min( nrow, s*32+32 ); ++r) km = std::max( km, ia[r+1]-ia[r] );
    base[s+1] = base[s] + (long long)km*32;
  }
  if ((size_t)base[nslice] != c->cg_nent) throw std::runtime_error( "csr_update: structure differs from the uploaded one" );
  std::vector< double > val( c->cg_nent, 0.0 ), diag( nrow, 0.0 );
  #pragma omp parallel for schedule(static)
  for (size_t r=0; r<nrow; ++r)
    for (size_t j=ia[r]-1, k=0; j<ia[r+1]-1; ++j, ++k) val[ (size_t)base[r/32] + k*32 + r%32 ] = a[j];
  CK( cudaMemcpyAsync( c->cg_val.p, val.data(), val.size()*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  k_cg_getdiag<<< nblk( c->cg_nslice*32, 256 ), 256, 0, c->stream >>>( nrow, c->cg_base.p, c->cg_col.p, c->cg_val.p, c->cg_diag.p ); ++c->launches;
  CK( cudaGetLastError() );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

int xyst_csr_mult( xyst_ctx* c, const double* x, double* r )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  if (!c->cg_nrow) throw std::runtime_error( "no matrix uploaded" );
  size_t n = c->cg_nrow;
  CK( cudaMemcpyAsync( c->cg_p.p, x, n*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  k_spmv<<< nblk( c->cg_nslice*32, 256 ), 256, 0, c->stream >>>( n, c->cg_base.p, c->cg_col.p, c->cg_val.p, c->cg_p.p, c->cg_q.p, nullptr ); ++c->launches;
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
    // beta = 0 on the first pass (next :586); done flag and iteration counter
    CK( cudaMemsetAsync( c->cg_scal.p + 3, 0, sizeof(double), s ) );
    CK( cudaMemsetAsync( c->cg_scal.p + 6, 0, 2*sizeof(double), s ) );
    const double nbn = c->cg_normb > 1.0e-14 ? c->cg_normb : 1.0;
    const double* done = c->cg_scal.p + 6;
    // The stop test of ConjugateGradients::x (:787-823) runs on the device (k_cg_scalars) and turns the
    // kernels of later iterations into no-ops, so one partition reads it back only every few
    // iterations; with a communicator every iteration is read (the collectives cannot be skipped).
    const size_t batch = c->comm ? 1 : 8;
    for (bool stop=false; !stop; ) {
      for (size_t b=0; b<batch && it+b<std::max< size_t >( maxit, 1 ); ++b) {
        k_cg_p<<< nblk( n, 256 ), 256, 0, s >>>( n, c->cg_scal.p, c->cg_z.p, c->cg_p.p ); ++c->launches;
        { ProfScope ps( c, "spmv" );
          cg_spmv( c, c->cg_p.p, c->cg_q.p, done ); }
        cg_halo( c, c->cg_q.p, 0 );
        k_cg_dot< 1 ><<< nb, RED_THREADS, 0, s >>>( n, c->cg_mask.p, c->cg_p.p, c->cg_q.p, nullptr, nullptr, c->red.p, done ); ++c->launches;
        cg_reduce( c, 1, nb );
        k_cg_scalars<<< 1, 32, 0, s >>>( 0, c->cg_scal.p, 0.0, 0.0 ); ++c->launches;
        k_cg_update<<< nb, RED_THREADS, 0, s >>>( n, c->cg_scal.p, c->cg_mask.p, c->cg_q.p, c->cg_d.p, c->cg_p.p,
          c->cg_r.p, c->cg_z.p, c->cg_x.p, c->red.p ); ++c->launches;
        cg_reduce( c, 2, nb );
        k_cg_scalars<<< 1, 32, 0, s >>>( 1, c->cg_scal.p, tol*nbn, (double)maxit ); ++c->launches;
        cg_halo( c, c->cg_x.p, 1 );
      }
      CK( cudaMemcpyAsync( c->red_host, c->cg_scal.p + 4, 4*sizeof(double), cudaMemcpyDeviceToHost, s ) );
      CK( cudaStreamSynchronize( s ) );
      it = (size_t)c->red_host[3];
      normr = std::sqrt( c->red_host[0] );
      stop = c->red_host[2] != 0.0 || it >= maxit;
      if (stop) c->cg_converged = !(normr > tol*nbn);
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
#define XYST_LOHCG_API
#include "lohcg.cuh"
#undef XYST_LOHCG_API

uint64_t xyst_launch_count( const xyst_ctx* c ) { return c ? c->launches : 0; }
uint64_t xyst_nedge( const xyst_ctx* c ) { return c ? c->nedge : 0; }

int xyst_kernel_time( xyst_ctx* c, const char* kernel, int reset, double* ms, uint64_t* launches )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  c->prof_on = true;
  auto& p = c->prof[ kernel ];
  if (!p.on) {                                  // first request for this name: switch it on, fill its pool
    p.on = true;
    for (int i=0; i<512; ++i) { cudaEvent_t x, y; CK( cudaEventCreate( &x ) ); CK( cudaEventCreate( &y ) ); p.ev.emplace_back( x, y ); }
  }
  CK( cudaStreamSynchronize( c->stream ) );
  for (size_t i=0; i<p.used; ++i) {
    float t = 0; CK( cudaEventElapsedTime( &t, p.ev[i].first, p.ev[i].second ) );
    p.ms += t; ++p.n;
  }
  p.used = 0;
  if (ms) *ms = p.ms;
  if (launches) *launches = p.n;
  if (reset) { p.ms = 0.0; p.n = 0; }
  API_END
}

} // extern "C"

// xyst_b200/csrc/riecg_tile.cuh -- thread-per-owner edge kernels of RieCG / LaxCG
// Part of the single translation unit xyst_b200.cu (included inside its anonymous namespace).
//
// One thread per OWNER node (the lower endpoint of an edge in the library's node order): the owner's
// 23 values (5 primitives, 3 coordinates, 15 gradient components) stay in registers over its ~7
// owned edges, the other end arrives with 12 16-byte gathers. An edge is always evaluated
// owner-first. Where the reference's orientation (src/Inciter/RieCG.cpp:659-714) has the owner as
// the edge's SECOND node, the limiter's epsilon (src/Physics/Riemann.cpp:92-95) and the edge normal
// are negated: with d1' = -d3, d2' = -d2, d3' = -d1 and eps' = -eps every quotient of the van Leer
// function is the reference's (numerator and denominator both negated), so the two reconstructed
// states are exactly the reference's, swapped, and the Riemann flux is the reference's with the
// opposite sign. Hence always: owner -= f', other end += f'.
//
//  k_flux_own    : drop-in for k_flux_edge (writes the per-edge fluxes F for a separate gather)
//  k_stage_tile  : one thread block per TILE of consecutive nodes (locality.hpp): edge fluxes,
//                  nodal sums and the Runge-Kutta update in one pass; no per-edge array in HBM.
//                  Fluxes whose receiver lives in the tile travel through shared memory, in the
//                  receiver's own column (k-th incoming edge of local node t at [k*TN + t]), so the
//                  sum per node runs in a fixed order: own edges ascending, then incoming edges
//                  ascending. Edges owned by a node of ANOTHER tile ("foreign") are evaluated a second
//                  time here, by the same device function on the same operands in the same order --
//                  bitwise the owner tile's value, so the scheme stays conservative.

#ifndef OWN_THREADS
#define OWN_THREADS 128
#endif
#ifndef OWN_MINB
#define OWN_MINB 4
#endif
#ifndef OWN_UNROLL
#define OWN_UNROLL 1
#endif
#ifndef TILE_MINB
#define TILE_MINB 2
#endif
constexpr int kOwnUnroll = OWN_UNROLL;

__device__ __forceinline__ void load_g( const double2* __restrict__ G2, size_t NP, size_t p, double g[15] ) {
  double2 a[NGP];
  #pragma unroll
  for (int k=0; k<NGP; ++k) a[k] = __ldg( G2 + (size_t)k*NP + p );
  #pragma unroll
  for (int k=0; k<7; ++k) { g[2*k] = a[k].x; g[2*k+1] = a[k].y; }
  g[14] = a[7].x;
}

// flux f' of the edge (owner -> other), owner-first. s = +1: the owner is the reference's first
// node, -1: its second. n = the stored (reference-oriented) normal.
template< bool EXACT, int FLUX >
__device__ __forceinline__ void edge_flux_owner( const double wo[NC], const double xo[3], const double go[15],
    const double wq[NC], const double xq[3], const double gq[15], double s, const double nref[3],
    const DParams& P, double f[NC] )
{
  double l[NC], r[NC], vw[3], n[3];
  #pragma unroll
  for (int c=0; c<NC; ++c) { l[c] = wo[c]; r[c] = wq[c]; }
  #pragma unroll
  for (int j=0; j<3; ++j) { vw[j] = xq[j] - xo[j]; n[j] = s * nref[j]; }
  muscl< EXACT >( go, 1, gq, 1, vw, l, r, s * MUSCL_EPS );
  if (FLUX == 0) rusanov( l, r, n, P, f ); else if (FLUX == 1) hllc( l, r, n, P, f );
  else if (FLUX == 2) lax_rusanov( l, r, n, P, f ); else lax_hllc( l, r, n, P, f );
}

template< bool EXACT, int FLUX >
__global__ void __launch_bounds__(OWN_THREADS, OWN_MINB)
k_flux_own( size_t nslice, size_t NP, size_t nslot, const long long* __restrict__ ebase, const int* __restrict__ eo,
            const double* __restrict__ D, const double* __restrict__ W, const double* __restrict__ G,
            double* __restrict__ F, DParams P )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (slice >= nslice) return;
  size_t p = slice*32 + lane;
  const double2* WX = reinterpret_cast< const double2* >( W );
  const double2* G2 = reinterpret_cast< const double2* >( G );
  const double2* D2 = reinterpret_cast< const double2* >( D );   // rows 0,1 are NOT interleaved: use D directly
  (void)D2;
  long long b0 = ebase[slice];
  int kmax = (int)((ebase[slice+1] - b0) >> 5);
  double wo[NC], xo[3], go[15];
  load_wx( WX, NP, p, wo, xo );
  load_g( G2, NP, p, go );
  #pragma unroll kOwnUnroll
  for (int j=0; j<kmax; ++j) {
    size_t sl = (size_t)b0 + (size_t)j*32 + lane;
    int e = __ldg( eo + sl );
    bool valid = e != -1;
    double s = e < 0 ? -1.0 : 1.0;
    size_t q = valid ? (size_t)(e & 0x7fffffff) : p;
    double n[3] = { __ldg( D + sl ), __ldg( D + nslot + sl ), __ldg( D + 2*nslot + sl ) };
    double wq[NC], xq[3], gq[15];
    load_wx( WX, NP, q, wq, xq );
    load_g( G2, NP, q, gq );
    double f[NC];
    edge_flux_owner< EXACT, FLUX >( wo, xo, go, wq, xq, gq, s, n, P, f );
    if (valid) {
      #pragma unroll
      for (int c=0; c<NC; ++c) f[c] *= s;       // F holds the reference-oriented flux for k_rhs_node
      store_f( F, nslot, sl, f );
    }
  }
}

// ---------------------------------------------------------------------------------
// fused stage kernel
// ---------------------------------------------------------------------------------
struct TileArgs {
  size_t npoin, NP, nslot;
  const int* tile_list;            // tiles this launch works on (blockIdx.x -> tile), or null = identity
  const int* tile_sl;              // [ntile+1] first slice of each tile
  const int* foff;                 // [ntile+1] offsets into the foreign-edge lists
  const int* fa; const int* fsl; const unsigned short* fdst;   // foreign edges: owner, global slot, F_s index
  const long long* ebase; const int* eo; const unsigned short* els;   // owned slots: other end|orientation, F_s index or 0xffff
  const unsigned char* indeg;      // [NP] number of incoming (not owned) edges
  const double* D;                 // [3][nslot] reference-oriented normals
  const double* W; const double* G;
  int fstride;                     // doubles per component plane of F_s
  // nodal tail: boundary + source, update
  const int* bslot; const double* Rb; const double* S; int src_mask; const double* v; const double* vol;
  const double* Un; double* U; double* Wout; double* R; double* Wn; double* UnOut;
  const int* shidx; double* part;  // shared nodes (several partitions): index into part, or -1
  StageArgs A; DParams P;
};

template< bool EXACT, int FLUX, bool FUSED, bool LAX >
__global__ void __launch_bounds__(256, TILE_MINB)
k_stage_tile( TileArgs T )
{
  extern __shared__ double Fs[];
  const int tile = T.tile_list ? T.tile_list[blockIdx.x] : (int)blockIdx.x;
  const int s0 = T.tile_sl[tile], s1 = T.tile_sl[tile+1];
  const int tn = (s1 - s0)*32;                       // nodes of this tile
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t NP = T.NP, nslot = T.nslot;
  const double2* WX = reinterpret_cast< const double2* >( T.W );
  const double2* G2 = reinterpret_cast< const double2* >( T.G );
  const bool active = tid < tn;
  const size_t p = (size_t)s0*32 + (active ? tid : 0);
  double acc[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) acc[c] = 0.0;
  // ---- owned edges ----
  if (active) {
    size_t slice = (size_t)s0 + (tid >> 5);
    long long b0 = T.ebase[slice];
    int kmax = (int)((T.ebase[slice+1] - b0) >> 5);
    double wo[NC], xo[3], go[15];
    load_wx( WX, NP, p, wo, xo );
    load_g( G2, NP, p, go );
    #pragma unroll kOwnUnroll
    for (int j=0; j<kmax; ++j) {
      size_t sl = (size_t)b0 + (size_t)j*32 + lane;
      int e = __ldg( T.eo + sl );
      unsigned dst = __ldg( T.els + sl );
      bool valid = e != -1;
      double s = e < 0 ? -1.0 : 1.0;
      size_t q = valid ? (size_t)(e & 0x7fffffff) : p;
      double n[3] = { __ldg( T.D + sl ), __ldg( T.D + nslot + sl ), __ldg( T.D + 2*nslot + sl ) };
      double wq[NC], xq[3], gq[15];
      load_wx( WX, NP, q, wq, xq );
      load_g( G2, NP, q, gq );
      double f[NC];
      edge_flux_owner< EXACT, FLUX >( wo, xo, go, wq, xq, gq, s, n, T.P, f );
      if (valid) {
        #pragma unroll
        for (int c=0; c<NC; ++c) acc[c] -= f[c];
        if (dst != 0xffffu) {
          #pragma unroll
          for (int c=0; c<NC; ++c) Fs[c*T.fstride + dst] = f[c];
        }
      }
    }
  }
  // ---- foreign edges: owned by a node of another tile, received here ----
  for (int i = T.foff[tile] + tid; i < T.foff[tile+1]; i += blockDim.x) {
    size_t a = (size_t)__ldg( T.fa + i ), sl = (size_t)__ldg( T.fsl + i );
    unsigned dst = __ldg( T.fdst + i );
    int e = __ldg( T.eo + sl );
    double s = e < 0 ? -1.0 : 1.0;
    size_t q = (size_t)(e & 0x7fffffff);
    double n[3] = { __ldg( T.D + sl ), __ldg( T.D + nslot + sl ), __ldg( T.D + 2*nslot + sl ) };
    double wo[NC], xo[3], go[15], wq[NC], xq[3], gq[15];
    load_wx( WX, NP, a, wo, xo );
    load_g( G2, NP, a, go );
    load_wx( WX, NP, q, wq, xq );
    load_g( G2, NP, q, gq );
    double f[NC];
    edge_flux_owner< EXACT, FLUX >( wo, xo, go, wq, xq, gq, s, n, T.P, f );
    #pragma unroll
    for (int c=0; c<NC; ++c) Fs[c*T.fstride + dst] = f[c];
  }
  __syncthreads();
  if (!active || p >= T.npoin) return;
  // ---- incoming fluxes, ascending owner ----
  { int kin = T.indeg[p];
    for (int k=0; k<kin; ++k) {
      #pragma unroll
      for (int c=0; c<NC; ++c) acc[c] += Fs[c*T.fstride + k*tn + tid];
    } }
  // ---- boundary faces, source (as rhs_sum) ----
  int b = T.bslot[p];
  if (b >= 0) {
    #pragma unroll
    for (int c=0; c<NC; ++c) acc[c] += T.Rb[(size_t)b*NC+c];
  }
  if (T.src_mask) {
    double vp = T.v[p];
    #pragma unroll
    for (int c=0; c<NC; ++c) if (T.src_mask & (1<<c)) acc[c] -= T.S[p*NC+c] * vp;
  }
  int sh = T.shidx ? T.shidx[p] : -1;
  if (sh >= 0) {                    // shared with other partitions: finished by k_rhs_finish
    #pragma unroll
    for (int c=0; c<NC; ++c) T.part[(size_t)sh*NC+c] = acc[c];
    if (FUSED) return;
  }
  if (FUSED) {
    node_update< LAX >( p, NP, acc, T.vol[p], T.Un, T.U, T.W, T.Wout, T.Wn, T.UnOut, T.A );
  } else {
    #pragma unroll
    for (int c=0; c<NC; ++c) T.R[p*NC+c] = acc[c];
  }
}

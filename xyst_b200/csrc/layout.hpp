// xyst_b200/csrc/layout.hpp -- device data layout of one mesh partition, computed on the host.
// Pure host code without CUDA types, so that tests/test_layout.py can build and check it on a CPU.
//
// Input: the reference's superedge groups (RieCG::m_dsupedge/m_dsupint, src/Inciter/RieCG.cpp:620-736)
// in the caller's node numbering. Output, all in the library's own node order (locality.hpp):
//  * unique oriented edges in OWNER SLOTS: owner = lower endpoint; the j-th edge owned by node o sits
//    in slot ebase[o/32] + 32 j + o%32 (ep/eq = its two nodes as the reference orients it, -1 = padding;
//    eo = other end | bit 31 if the owner is the edge's second node);
//  * the sliced-ELL node incidence of the gather kernels: per node its owned edges (ascending other
//    end), then the edges owned by lower neighbours (ascending owner) -- the fixed summation order;
//  * the INCOMING-edge lists (edges owned by lower neighbours) of k_update_in, which completes the
//    nodal sums the owner-thread flux kernel starts.
#pragma once
#include <vector>
#include <array>
#include <algorithm>
#include <numeric>
#include <stdexcept>
#include <cstddef>
#include <cstdint>
#include <cmath>
#include "locality.hpp"

namespace layout {

struct Options {
  bool reorder = false;        // re-order the nodes into tiles (locality.hpp)
  size_t tile_nodes = 256;     // nodes per tile of the library's order (multiple of 32)
};

struct Mesh {
  size_t npoin = 0, ne = 0, nslice = 0, nslot = 0, nent = 0;
  int drows = 4, maxdeg = 0;
  std::vector< int > new2old, old2new;             // empty: the caller's order is kept
  std::vector< long long > ebase, base;            // [nslice+1] slot / incidence-entry offsets
  std::vector< int > ep, eq, eo;                   // [nslot]
  std::vector< double > ed;                        // [drows][nslot] edge integrals (normal, + extra terms)
  std::vector< int > inc_e, inc_q;                 // [nent] signed slot+1 (0 = padding), neighbour
  std::vector< long long > in_base;                // [nslice+1] offsets of the INCOMING-edge lists
  std::vector< int > in_e;                         // signed slot+1 of the edges owned by lower neighbours
  size_t to_new( size_t old ) const { return old2new.empty() ? old : (size_t)old2new[old]; }
};

// tk::lpoed / tk::lpoet, src/Mesh/DerivedData.hpp:40-44
static const int lpoed[6][2] = { {0,1}, {1,2}, {2,0}, {0,3}, {1,3}, {2,3} };
static const int lpoet[3][2] = { {0,1}, {1,2}, {2,0} };

inline Mesh build( size_t npoin, const double* x, const double* y, const double* z,
                   const size_t nsup[3], const size_t* const dsupedge[3], const double* const dsupint[3],
                   size_t stride, const Options& opt )
{
  Mesh M;
  M.npoin = npoin;
  if (npoin == 0 || npoin > 0x7fffffffULL) throw std::runtime_error( "npoin out of range" );
  static const size_t nn[3] = { 4, 3, 2 };
  for (int k=0; k<3; ++k) for (size_t i=0; i<nsup[k]*nn[k]; ++i)
    if (dsupedge[k][i] >= npoin) throw std::runtime_error( "node id out of range in superedge" );
  if (opt.reorder) {
    M.new2old = locality::tile_order( npoin, x, y, z, opt.tile_nodes );
    M.old2new.resize( npoin );
    for (size_t i=0; i<npoin; ++i) M.old2new[ (size_t)M.new2old[i] ] = (int)i;
  }
  // --- flatten superedges to edges (orientation and integrals as given) ----------------
  size_t ne = nsup[0]*6 + nsup[1]*3 + nsup[2];
  if (ne > 0x7ffffff0ULL) throw std::runtime_error( "too many edges for 32-bit edge ids" );
  M.ne = ne;
  const int drows = stride > 4 ? 5 : 4;
  M.drows = drows;
  std::vector< int > P( ne ), Q( ne );                    // end nodes, library numbering
  auto integ = [&]( size_t i ) -> const double* {         // integrals of flattened edge i
    if (i < nsup[0]*6) return dsupint[0] + i*stride;
    i -= nsup[0]*6;
    if (i < nsup[1]*3) return dsupint[1] + i*stride;
    return dsupint[2] + (i - nsup[1]*3)*stride;
  };
  #pragma omp parallel for schedule(static)
  for (size_t e=0; e<nsup[0]; ++e)
    for (int k=0; k<6; ++k) {
      P[e*6+k] = (int)M.to_new( dsupedge[0][e*4+lpoed[k][0]] ); Q[e*6+k] = (int)M.to_new( dsupedge[0][e*4+lpoed[k][1]] ); }
  { size_t o = nsup[0]*6;
    #pragma omp parallel for schedule(static)
    for (size_t e=0; e<nsup[1]; ++e)
      for (int k=0; k<3; ++k) {
        P[o+e*3+k] = (int)M.to_new( dsupedge[1][e*3+lpoet[k][0]] ); Q[o+e*3+k] = (int)M.to_new( dsupedge[1][e*3+lpoet[k][1]] ); }
    o += nsup[1]*3;
    #pragma omp parallel for schedule(static)
    for (size_t e=0; e<nsup[2]; ++e) { P[o+e] = (int)M.to_new( dsupedge[2][e*2] ); Q[o+e] = (int)M.to_new( dsupedge[2][e*2+1] ); } }
  // --- edge slots in owner order ----------------------------------------------------------
  // edges sorted by (owner, other end, input position)
  struct Key { uint64_t k; int i; };
  std::vector< Key > perm( ne );
  #pragma omp parallel for schedule(static)
  for (size_t i=0; i<ne; ++i) {
    uint64_t lo = (uint64_t)std::min( P[i], Q[i] ), hi = (uint64_t)std::max( P[i], Q[i] );
    perm[i] = Key{ (lo << 32) | hi, (int)i };
  }
  std::sort( perm.begin(), perm.end(), []( const Key& a, const Key& b ){ return a.k != b.k ? a.k < b.k : a.i < b.i; } );
  size_t nslice = (npoin + 31) / 32;
  M.nslice = nslice;
  std::vector< int > udeg( npoin, 0 ), deg( npoin, 0 );
  for (size_t i=0; i<ne; ++i) { ++udeg[ std::min( P[i], Q[i] ) ]; ++deg[P[i]]; ++deg[Q[i]]; }
  M.ebase.assign( nslice+1, 0 ); M.base.assign( nslice+1, 0 );
  for (size_t s=0; s<nslice; ++s) {
    int km = 0, kd = 0;
    for (size_t p=s*32; p<std::min( npoin, s*32+32 ); ++p) { km = std::max( km, udeg[p] ); kd = std::max( kd, deg[p] ); }
    M.ebase[s+1] = M.ebase[s] + (long long)km*32;
    M.base[s+1] = M.base[s] + (long long)kd*32;
  }
  size_t nslot = (size_t)M.ebase[nslice];
  if (nslot > 0x7ffffff0ULL) throw std::runtime_error( "too many edge slots for 32-bit ids" );
  M.nslot = nslot;
  M.ep.assign( nslot, -1 ); M.eq.assign( nslot, -1 ); M.eo.assign( nslot, -1 );
  M.ed.assign( (size_t)drows*nslot, 0.0 );
  std::vector< int > slot_of( ne );                       // by sorted position
  { std::vector< int > fillu( npoin, 0 );
    for (size_t i=0; i<ne; ++i) {
      size_t e = (size_t)perm[i].i;
      int o = std::min( P[e], Q[e] );
      size_t sl = (size_t)M.ebase[(size_t)o/32] + (size_t)fillu[o]*32 + (size_t)(o%32);
      ++fillu[o];
      M.ep[sl] = P[e]; M.eq[sl] = Q[e]; slot_of[i] = (int)sl;
      M.eo[sl] = P[e] < Q[e] ? Q[e] : (int)( (unsigned)P[e] | 0x80000000u );
      const double* d = integ( e );
      for (size_t j=0; j<stride; ++j) M.ed[j*nslot+sl] = d[j];
      // RieCG/LaxCG (3 integrals per edge): the free 4th row holds the normal's length, which the
      // Riemann solvers would otherwise recompute per edge and stage (Riemann.cpp:412)
      if (stride == 3) M.ed[3*nslot+sl] = std::sqrt( d[0]*d[0] + d[1]*d[1] + d[2]*d[2] );
    } }
  // --- sliced-ELL incidence: node -> (signed slot, neighbour) -------------------------------
  // reference scatter: G(p) -= f, G(q) += f with f = d*(u_q+u_p)  (Riemann.cpp:321-323): the
  // entry is +(slot+1) for the edge's second node, -(slot+1) for its first, 0 = padding.
  size_t nent = (size_t)M.base[nslice];
  M.nent = nent;
  M.maxdeg = 0; for (size_t p=0; p<npoin; ++p) M.maxdeg = std::max( M.maxdeg, deg[p] );
  M.inc_e.assign( nent, 0 ); M.inc_q.assign( nent, 0 );
  for (size_t sl=0; sl<nslice; ++sl)            // padding: the node itself (last node for the tail slice)
    for (size_t j=(size_t)M.base[sl]; j<(size_t)M.base[sl+1]; ++j)
      M.inc_q[j] = (int)std::min( npoin-1, sl*32 + (j - (size_t)M.base[sl])%32 );
  { std::vector< int > fill( npoin, 0 );
    auto addinc = [&]( int node, int other, int signedslot ) {
      size_t slot = (size_t)M.base[(size_t)node/32] + (size_t)fill[node]*32 + (size_t)(node%32);
      ++fill[node];
      M.inc_e[slot] = signedslot; M.inc_q[slot] = other;
    };
    for (size_t i=0; i<ne; ++i) {                 // owned edges (sorted by owner, other)
      size_t e = (size_t)perm[i].i;
      int o = std::min( P[e], Q[e] ), h = std::max( P[e], Q[e] );
      addinc( o, h, o == Q[e] ? slot_of[i]+1 : -(slot_of[i]+1) );
    }
    for (size_t i=0; i<ne; ++i) {                 // edges owned by lower neighbours: arrive ascending in owner
      size_t e = (size_t)perm[i].i;
      int o = std::min( P[e], Q[e] ), h = std::max( P[e], Q[e] );
      addinc( h, o, h == Q[e] ? slot_of[i]+1 : -(slot_of[i]+1) );
    } }
  // --- incoming edges only (k_update_in): sliced ELL, ascending owner ---------------------------
  { M.in_base.assign( nslice+1, 0 );
    for (size_t s=0; s<nslice; ++s) {
      int km = 0;
      for (size_t p=s*32; p<std::min( npoin, s*32+32 ); ++p) km = std::max( km, deg[p] - udeg[p] );
      M.in_base[s+1] = M.in_base[s] + (long long)km*32;
    }
    M.in_e.assign( (size_t)M.in_base[nslice], 0 );
    std::vector< int > fill( npoin, 0 );
    for (size_t i=0; i<ne; ++i) {
      size_t e = (size_t)perm[i].i;
      int h = std::max( P[e], Q[e] );
      size_t pos = (size_t)M.in_base[(size_t)h/32] + (size_t)fill[h]*32 + (size_t)(h%32);
      ++fill[h];
      M.in_e[pos] = h == Q[e] ? slot_of[i]+1 : -(slot_of[i]+1);
    } }
  return M;
}

} // namespace layout

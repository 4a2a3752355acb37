// tests/layout_check.cpp -- CPU check of the device data layout (xyst_b200/csrc/layout.hpp).
// Test infrastructure: built by tests/test_layout.py with g++, no CUDA. It replays what the fused
// stage kernel (riecg_tile.cuh) does with the tile structures, with the edge's slot id standing in
// for its flux, and verifies that every node receives exactly its incident edges, in the order of
// the gather kernels' incidence lists.
#include <cstdio>
#include <cstring>
#include <string>
#include <set>
#include "../xyst_b200/csrc/layout.hpp"

extern "C" int layout_check( size_t npoin, const double* x, const double* y, const double* z,
                             const size_t nsup[3], const size_t* const dsupedge[3], const double* const dsupint[3],
                             size_t stride, int reorder, size_t tile_nodes, size_t cap,
                             size_t* stats /* [10] */, char* msg, size_t msglen )
{
  auto fail = [&]( const std::string& m ){ std::snprintf( msg, msglen, "%s", m.c_str() ); return 1; };
  try {
    layout::Options opt; opt.reorder = reorder != 0; opt.tiles = true; opt.tile_nodes = tile_nodes; opt.cap = cap;
    layout::Mesh M = layout::build( npoin, x, y, z, nsup, dsupedge, dsupint, stride, opt );
    // 1. permutation
    if (opt.reorder) {
      if (M.new2old.size() != npoin) return fail( "perm size" );
      std::vector< char > seen( npoin, 0 );
      for (auto o : M.new2old) { if (o < 0 || (size_t)o >= npoin || seen[(size_t)o]) return fail( "not a permutation" ); seen[(size_t)o] = 1; }
      for (size_t i=0; i<npoin; ++i) if ((size_t)M.old2new[(size_t)M.new2old[i]] != i) return fail( "old2new is not the inverse" );
    }
    // 2. every input edge sits in exactly one slot with its orientation and integrals
    size_t nvalid = 0;
    for (size_t sl=0; sl<M.nslot; ++sl) if (M.ep[sl] >= 0) ++nvalid;
    if (nvalid != M.ne) return fail( "number of filled slots != number of edges" );
    {
      size_t i = 0;
      auto find = [&]( size_t a, size_t b, const double* d ) -> bool {
        int p = (int)M.to_new( a ), q = (int)M.to_new( b );
        int o = std::min( p, q );
        size_t slice = (size_t)o/32;
        for (long long s=M.ebase[slice] + o%32; s<M.ebase[slice+1]; s+=32)
          if (M.ep[(size_t)s] == p && M.eq[(size_t)s] == q) {
            for (size_t j=0; j<stride; ++j) if (M.ed[j*M.nslot+(size_t)s] != d[j]) return false;
            int eo = M.eo[(size_t)s];
            if ((eo & 0x7fffffff) != std::max( p, q )) return false;
            if ((eo < 0) != (o == q)) return false;
            return true;
          }
        return false;
      };
      for (size_t e=0; e<nsup[0]; ++e) for (int k=0; k<6; ++k, ++i)
        if (!find( dsupedge[0][e*4+layout::lpoed[k][0]], dsupedge[0][e*4+layout::lpoed[k][1]], dsupint[0]+(e*6+k)*stride )) return fail( "tet-superedge edge not found" );
      for (size_t e=0; e<nsup[1]; ++e) for (int k=0; k<3; ++k, ++i)
        if (!find( dsupedge[1][e*3+layout::lpoet[k][0]], dsupedge[1][e*3+layout::lpoet[k][1]], dsupint[1]+(e*3+k)*stride )) return fail( "triangle-superedge edge not found" );
      for (size_t e=0; e<nsup[2]; ++e, ++i)
        if (!find( dsupedge[2][e*2], dsupedge[2][e*2+1], dsupint[2]+e*stride )) return fail( "edge not found" );
    }
    // 3. replay of the tile kernel
    size_t nforeign = M.fa.size(), maxtn = 0;
    std::vector< long long > Fs( (size_t)std::max( M.fstride, 1 ) );
    for (size_t t=0; t<M.ntile; ++t) {
      int s0 = M.tile_sl[t], s1 = M.tile_sl[t+1], tn = (s1-s0)*32;
      if (tn > 256 || tn <= 0) return fail( "tile size" );
      maxtn = std::max( maxtn, (size_t)tn );
      std::fill( Fs.begin(), Fs.end(), -1 );
      std::vector< std::vector< long long > > own( (size_t)tn );
      for (int tid=0; tid<tn; ++tid) {
        size_t p = (size_t)s0*32 + (size_t)tid, slice = p/32;
        int kmax = (int)((M.ebase[slice+1]-M.ebase[slice]) >> 5);
        for (int j=0; j<kmax; ++j) {
          size_t sl = (size_t)M.ebase[slice] + (size_t)j*32 + p%32;
          int e = M.eo[sl];
          if (e == -1) continue;
          if (std::min( M.ep[sl], M.eq[sl] ) != (int)p) return fail( "slot not owned by its thread" );
          own[(size_t)tid].push_back( (long long)sl );
          unsigned dst = M.els[sl];
          size_t q = (size_t)(e & 0x7fffffff);
          bool intile = q >= (size_t)s0*32 && q < (size_t)s1*32;
          if (intile != (dst != 0xffffu)) return fail( "els does not match the receiver's tile" );
          if (dst != 0xffffu) {
            if ((int)dst >= M.fstride) return fail( "els beyond fstride" );
            if (Fs[dst] != -1) return fail( "two fluxes in one shared-memory position" );
            if (dst % (unsigned)tn != q - (size_t)s0*32) return fail( "els column is not the receiver" );
            Fs[dst] = (long long)sl;
          }
        }
      }
      for (int i=M.foff[t]; i<M.foff[t+1]; ++i) {
        size_t a = (size_t)M.fa[(size_t)i], sl = (size_t)M.fsl[(size_t)i]; unsigned dst = M.fdst[(size_t)i];
        int e = M.eo[sl];
        if (e == -1) return fail( "foreign edge on a padding slot" );
        size_t q = (size_t)(e & 0x7fffffff);
        if ((size_t)std::min( M.ep[sl], M.eq[sl] ) != a) return fail( "foreign owner mismatch" );
        if (a >= (size_t)s0*32 && a < (size_t)s1*32) return fail( "foreign owner inside the tile" );
        if (!(q >= (size_t)s0*32 && q < (size_t)s1*32)) return fail( "foreign receiver outside the tile" );
        if ((int)dst >= M.fstride || Fs[dst] != -1) return fail( "foreign position taken" );
        if (dst % (unsigned)tn != q - (size_t)s0*32) return fail( "foreign column is not the receiver" );
        Fs[dst] = (long long)sl;
      }
      for (int tid=0; tid<tn; ++tid) {
        size_t p = (size_t)s0*32 + (size_t)tid;
        if (p >= npoin) continue;
        std::vector< long long > got = own[(size_t)tid];
        for (int k=0; k<(int)M.indeg[p]; ++k) {
          long long v = Fs[(size_t)k*(size_t)tn + (size_t)tid];
          if (v < 0) return fail( "missing incoming flux" );
          if (std::max( M.ep[(size_t)v], M.eq[(size_t)v] ) != (int)p) return fail( "incoming flux of another node" );
          got.push_back( v );
        }
        // the gather kernels' incidence list of this node: same edges, same order
        size_t slice = p/32;
        std::vector< long long > ref;
        for (long long s=M.base[slice] + (long long)(p%32); s<M.base[slice+1]; s+=32) {
          int se = M.inc_e[(size_t)s];
          if (se != 0) ref.push_back( (long long)(std::abs( se )-1) );
        }
        if (got != ref) return fail( "tile sums differ from the incidence list (edges or order)" );
      }
    }
    // coalescing of the owner kernels' gathers: 32-byte sectors and 128-byte lines one warp-wide 16-byte
    // load of the other ends touches (ideal: 16 and 4), averaged over all (slice, j) with >= 1 valid lane
    { double sec = 0, lin = 0, cnt = 0;
      for (size_t sl=0; sl<M.nslice; ++sl) {
        int kmax = (int)((M.ebase[sl+1]-M.ebase[sl]) >> 5);
        for (int j=0; j<kmax; ++j) {
          std::set< long long > S, L; int nv = 0;
          for (int lane=0; lane<32; ++lane) {
            int e = M.eo[(size_t)M.ebase[sl] + (size_t)j*32 + (size_t)lane];
            if (e == -1) continue;
            long long q = e & 0x7fffffff; ++nv;
            S.insert( q/2 ); L.insert( q/8 );
          }
          if (nv) { sec += (double)S.size()*32.0/nv; lin += (double)L.size()*32.0/nv; cnt += 1; }
        }
      }
      stats[8] = (size_t)(1000.0*sec/cnt); stats[9] = (size_t)(1000.0*lin/cnt); }
    stats[0] = M.ne; stats[1] = M.nslot; stats[2] = M.ntile; stats[3] = nforeign; stats[4] = (size_t)M.fstride;
    stats[5] = maxtn; stats[6] = M.nent; stats[7] = (size_t)M.maxdeg;
  } catch (std::exception& e) { return fail( e.what() ); }
  return 0;
}

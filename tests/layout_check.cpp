// tests/layout_check.cpp -- CPU check of the device data layout (xyst_b200/csrc/layout.hpp).
// Test infrastructure: built by tests/test_layout.py with g++, no CUDA. It replays what the owner-thread
// flux kernel and k_update_in (riecg_own.cuh) do with the slot and incoming-edge structures, with the
// edge's slot id standing in for its flux, and verifies that every node receives exactly its incident
// edges, in the order of the gather kernels' incidence lists.
#include <cstdio>
#include <cstring>
#include <string>
#include <set>
#include "../xyst_b200/csrc/layout.hpp"

extern "C" int layout_check( size_t npoin, const double* x, const double* y, const double* z,
                             const size_t nsup[3], const size_t* const dsupedge[3], const double* const dsupint[3],
                             size_t stride, int reorder, size_t tile_nodes,
                             size_t* stats /* [6] */, char* msg, size_t msglen )
{
  auto fail = [&]( const std::string& m ){ std::snprintf( msg, msglen, "%s", m.c_str() ); return 1; };
  try {
    layout::Options opt; opt.reorder = reorder != 0; opt.tile_nodes = tile_nodes;
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
    // 3. replay of the flux kernel + k_update_in with the slot id standing in for the flux: every node
    //    must receive exactly its incident edges, own ones first (ascending other end), then the
    //    incoming ones (ascending owner) -- the order of the gather kernels' incidence lists
    for (size_t p=0; p<npoin; ++p) {
      size_t slice = p/32;
      std::vector< long long > got, ref;
      for (long long s=M.ebase[slice] + (long long)(p%32); s<M.ebase[slice+1]; s+=32) {
        if (M.eo[(size_t)s] == -1) continue;
        if (std::min( M.ep[(size_t)s], M.eq[(size_t)s] ) != (int)p) return fail( "slot not owned by its thread" );
        got.push_back( s );
      }
      for (long long s=M.in_base[slice] + (long long)(p%32); s<M.in_base[slice+1]; s+=32) {
        int se = M.in_e[(size_t)s];
        if (se == 0) continue;
        long long sl = std::abs( se )-1;
        if (std::max( M.ep[(size_t)sl], M.eq[(size_t)sl] ) != (int)p) return fail( "incoming edge of another node" );
        if ((se > 0) != (M.eq[(size_t)sl] == (int)p)) return fail( "incoming edge: wrong sign" );
        got.push_back( sl );
      }
      for (long long s=M.base[slice] + (long long)(p%32); s<M.base[slice+1]; s+=32) {
        int se = M.inc_e[(size_t)s];
        if (se != 0) ref.push_back( (long long)(std::abs( se )-1) );
      }
      if (got != ref) return fail( "owner + incoming lists differ from the incidence list (edges or order)" );
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
      stats[4] = (size_t)(1000.0*sec/cnt); stats[5] = (size_t)(1000.0*lin/cnt); }
    stats[0] = M.ne; stats[1] = M.nslot; stats[2] = M.nent; stats[3] = (size_t)M.maxdeg;
  } catch (std::exception& e) { return fail( e.what() ); }
  return 0;
}

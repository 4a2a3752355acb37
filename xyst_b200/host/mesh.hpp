// xyst_b200/host/mesh.hpp -- host-side mesh data for the B200 RieCG path.
//
// Scalable (sort/CSR based, no hash maps) counterparts of the reference's mesh
// preparation that feeds the hot path:
//   structured box generator    (BASELINE.json: synthetic Kuhn 6-tet box meshes)
//   global2local                src/Mesh/Reorder.cpp:279-306
//   points surrounding points   src/Mesh/DerivedData.cpp:132-223 (as unique-edge CSR)
//   coordinate bisection        stands in for Zoltan RCB, src/Partition/ZoltanGeom.cpp:139-244
// Results on small meshes are compared bit for bit with the oracle in tests/.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <map>
#include <vector>

namespace xyst {

using real = double;
using Coords = std::array< std::vector< real >, 3 >;

//! A tetrahedron mesh (or a partition of one) with its side sets, global node ids
struct TetMesh {
  Coords coord;                                         //!< node coordinates, indexed by global id
                                                        //!< (or by position in `gid` if `gid` non-empty)
  std::vector< std::size_t > gid;                       //!< optional: global ids of the stored nodes (sorted)
  std::vector< std::size_t > ginpoel;                   //!< 4 global node ids per tet
  std::map< int, std::vector< std::size_t > > sidetri;  //!< side set id -> boundary triangles (3 global ids each)
};

//! Box [0,Lx]x[0,Ly]x[0,Lz] of nx*ny*nz hexahedra, each split into 6 tetrahedra along the
//! (0,0,0)-(1,1,1) diagonal. Node id (k(ny+1)+j)(nx+1)+i. Side sets 1..6 = x-,x+,y-,y+,z-,z+.
//! Only the hexes [i0,i1)x[j0,j1)x[k0,k1) are generated (a partition of the full box);
//! node ids and coordinates always refer to the FULL box.
TetMesh boxMesh( std::size_t nx, std::size_t ny, std::size_t nz, real Lx, real Ly, real Lz,
                 std::size_t i0, std::size_t i1, std::size_t j0, std::size_t j1,
                 std::size_t k0, std::size_t k1 );

//! Recursive coordinate bisection of tet centroids into nparts (power of two) parts,
//! cutting the longest extent at the median; deterministic. Returns part id per tet.
std::vector< int > rcb( const Coords& coord, const std::vector< std::size_t >& ginpoel, int nparts );
//! the same for part = "rib" (Zoltan's recursive inertial bisection, serial_rib + inertial3d restated)
std::vector< int > rib( const Coords& coord, const std::vector< std::size_t >& ginpoel, int nparts );

//! Hex-index ranges of the `part`-th of nparts (1,2,4,8) RCB parts of a uniform box: what
//! rcb() yields for boxMesh (verified in tests), without building the full mesh.
std::array< std::size_t, 6 > boxPartRange( std::size_t nx, std::size_t ny, std::size_t nz,
                                           int nparts, int part );

//! Unique undirected edges of a tet mesh as CSR over the LOWER local node id:
//! edges of node p are hi[ off[p] .. off[p+1] ), ascending. Edge id = position in hi.
struct EdgeCSR {
  std::vector< std::size_t > off;
  std::vector< std::uint32_t > hi;
  std::size_t nedge() const { return hi.size(); }
  //! edge id of (a,b), a != b; must exist
  std::size_t find( std::size_t a, std::size_t b ) const {
    if (a > b) { auto t = a; a = b; b = t; }
    std::size_t lo = off[a], up = off[a+1];
    while (lo < up) { auto m = (lo+up)/2; if (hi[m] < b) lo = m+1; else up = m; }
    return lo;
  }
};
EdgeCSR uniqueEdges( const std::vector< std::size_t >& inpoel, std::size_t npoin );

//! Points surrounding points (both directions, ascending per point) from the edge CSR
void psupFromEdges( const EdgeCSR& e, std::size_t npoin,
                    std::vector< std::size_t >& off, std::vector< std::uint32_t >& nbr );

} // xyst::

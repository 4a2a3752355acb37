// xyst_b200/csrc/locality.hpp -- node order used INSIDE the device library (host code).
//
// The reference renumbers the nodes of a partition for locality with a graph traversal
// (src/Inciter/RieCG.cpp:82-100); callers of the C ABI keep that numbering, and every array
// crossing the ABI is in it. Inside the library the nodes are re-ordered once more, for the GPU:
// a k-d bisection of the node coordinates into TILES of exactly `tile` consecutive nodes (the
// last one may be short), nodes inside a tile in lexicographic (z,y,x) order. One thread block
// works on one tile; its nodes' neighbours are then the tile itself plus a thin halo, so that
// the edge kernels find their operands in L1/shared memory, and 32 consecutive nodes (one warp)
// reach runs of consecutive neighbours (coalesced 16-byte gathers).
//
// The order is monotone on a structured box: a node with componentwise larger coordinates gets a
// larger id, so every node owns exactly the edges to its "upper" neighbours.
#pragma once
#include <vector>
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>

namespace locality {

struct Coords { const double *x, *y, *z; };

inline bool lex_less( const Coords& X, int a, int b, int axis )
{
  // primary key: coordinate along `axis`; ties: (z,y,x) lexicographic, then the caller's id
  const double* k[3] = { X.x, X.y, X.z };
  if (k[axis][a] != k[axis][b]) return k[axis][a] < k[axis][b];
  if (X.z[a] != X.z[b]) return X.z[a] < X.z[b];
  if (X.y[a] != X.y[b]) return X.y[a] < X.y[b];
  if (X.x[a] != X.x[b]) return X.x[a] < X.x[b];
  return a < b;
}

inline void bisect( const Coords& X, int* idx, size_t n, size_t tile, const double w[3] )
{
  while (n > tile) {
    // axis of the largest weighted extent
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    const double* k[3] = { X.x, X.y, X.z };
    for (size_t i=0; i<n; ++i)
      for (int j=0; j<3; ++j) { double v = k[j][idx[i]]; lo[j] = std::min( lo[j], v ); hi[j] = std::max( hi[j], v ); }
    int axis = 0; double best = -1.0;
    for (int j=0; j<3; ++j) { double e = (hi[j]-lo[j])*w[j]; if (e > best) { best = e; axis = j; } }
    size_t ntile = (n + tile - 1) / tile;
    size_t nleft = (ntile/2) * tile;          // whole tiles on the left
    std::nth_element( idx, idx + nleft, idx + n, [&]( int a, int b ){ return lex_less( X, a, b, axis ); } );
    // recurse into the smaller half, loop on the larger (bounded stack depth)
    if (nleft <= n - nleft) { bisect( X, idx, nleft, tile, w ); idx += nleft; n -= nleft; }
    else { bisect( X, idx + nleft, n - nleft, tile, w ); n = nleft; }
  }
  std::sort( idx, idx + n, [&]( int a, int b ){ return lex_less( X, a, b, 2 ); } );
}

// new2old[i] = caller's id of the node that gets internal id i
inline std::vector< int > tile_order( size_t npoin, const double* x, const double* y, const double* z, size_t tile )
{
  std::vector< int > idx( npoin );
  for (size_t i=0; i<npoin; ++i) idx[i] = (int)i;
  Coords X{ x, y, z };
  // x runs fastest inside a tile: prefer tiles about twice as long in x as in y and z
  double w[3] = { 0.5, 1.0, 1.0 };
  if (const char* e = getenv( "XYST_TILE_WX" )) w[0] = atof( e );
  bisect( X, idx.data(), npoin, tile, w );
  return idx;
}

} // namespace locality

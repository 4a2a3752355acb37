// xyst_b200/host/mesh.cpp -- see mesh.hpp
#include "mesh.hpp"
#include <algorithm>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <cmath>

namespace xyst {

namespace {
// local edges of a tetrahedron, tk::lpoed (src/Mesh/DerivedData.hpp:40-41)
const int lpoed[6][2] = { {0,1}, {1,2}, {2,0}, {0,3}, {1,3}, {2,3} };

// The six Kuhn tetrahedra of the unit hex as corner offsets (dx,dy,dz), each a monotone
// path from corner 000 to corner 111; node order chosen so that the Jacobian
// triple(b-a, c-a, d-a) is positive (the reference requires J>0,
// src/Inciter/Discretization.cpp:100-101,:629).
struct Kuhn {
  int c[6][4][3];
  Kuhn() {
    const int perm[6][3] = { {0,1,2}, {0,2,1}, {1,0,2}, {1,2,0}, {2,0,1}, {2,1,0} };
    for (int t=0; t<6; ++t) {
      int v[4][3] = { {0,0,0}, {0,0,0}, {0,0,0}, {1,1,1} };
      v[1][perm[t][0]] = 1;
      v[2][perm[t][0]] = 1; v[2][perm[t][1]] = 1;
      double ba[3], ca[3], da[3];
      for (int d=0; d<3; ++d) { ba[d] = v[1][d]-v[0][d]; ca[d] = v[2][d]-v[0][d]; da[d] = v[3][d]-v[0][d]; }
      double J = ba[0]*(ca[1]*da[2]-da[1]*ca[2]) + ba[1]*(ca[2]*da[0]-da[2]*ca[0]) + ba[2]*(ca[0]*da[1]-da[0]*ca[1]);
      if (J < 0) for (int d=0; d<3; ++d) std::swap( v[1][d], v[2][d] );
      for (int k=0; k<4; ++k) for (int d=0; d<3; ++d) c[t][k][d] = v[k][d];
    }
  }
};
const Kuhn kuhn;
}

TetMesh boxMesh( std::size_t nx, std::size_t ny, std::size_t nz, real Lx, real Ly, real Lz,
                 std::size_t i0, std::size_t i1, std::size_t j0, std::size_t j1,
                 std::size_t k0, std::size_t k1 )
{
  if (i1 > nx || j1 > ny || k1 > nz || i0 >= i1 || j0 >= j1 || k0 >= k1)
    throw std::runtime_error( "boxMesh: invalid hex range" );
  TetMesh m;
  const std::size_t px = nx+1, py = ny+1;
  auto id = [&]( std::size_t i, std::size_t j, std::size_t k ){ return (k*py + j)*px + i; };
  const bool full = i0 == 0 && j0 == 0 && k0 == 0 && i1 == nx && j1 == ny && k1 == nz;
  // nodes of this range; coordinates are those of the full box (i*Lx/nx, ...)
  const std::size_t mx = i1-i0+1, my = j1-j0+1, mz = k1-k0+1;
  for (auto& c : m.coord) c.resize( mx*my*mz );
  if (!full) m.gid.resize( mx*my*mz );
  #pragma omp parallel for schedule(static)
  for (std::size_t k=k0; k<=k1; ++k)
    for (std::size_t j=j0; j<=j1; ++j)
      for (std::size_t i=i0; i<=i1; ++i) {
        std::size_t l = ((k-k0)*my + (j-j0))*mx + (i-i0);
        m.coord[0][l] = Lx * static_cast< real >( i ) / static_cast< real >( nx );
        m.coord[1][l] = Ly * static_cast< real >( j ) / static_cast< real >( ny );
        m.coord[2][l] = Lz * static_cast< real >( k ) / static_cast< real >( nz );
        if (!full) m.gid[l] = id( i, j, k );      // ascending in l since ranges are boxes
      }
  // tetrahedra
  const std::size_t nhex = (i1-i0)*(j1-j0)*(k1-k0);
  m.ginpoel.resize( nhex*24 );
  #pragma omp parallel for schedule(static)
  for (std::size_t k=k0; k<k1; ++k)
    for (std::size_t j=j0; j<j1; ++j)
      for (std::size_t i=i0; i<i1; ++i) {
        std::size_t h = ((k-k0)*(j1-j0) + (j-j0))*(i1-i0) + (i-i0);
        for (int t=0; t<6; ++t)
          for (int v=0; v<4; ++v)
            m.ginpoel[h*24 + static_cast<std::size_t>(t)*4 + static_cast<std::size_t>(v)] =
              id( i+static_cast<std::size_t>(kuhn.c[t][v][0]), j+static_cast<std::size_t>(kuhn.c[t][v][1]),
                  k+static_cast<std::size_t>(kuhn.c[t][v][2]) );
      }
  // side sets: tet faces lying on a face of the FULL box, outward orientation as the
  // reference derives it from the tet (src/Inciter/Partitioner.cpp:575-577)
  const int tf[4][3] = { {0,2,1}, {0,1,3}, {0,3,2}, {1,2,3} };
  auto addfaces = [&]( int set, int dim, int side, std::size_t a0, std::size_t a1, std::size_t b0, std::size_t b1,
                       std::size_t fixed ) {
    auto& out = m.sidetri[set];
    for (std::size_t b=b0; b<b1; ++b)
      for (std::size_t a=a0; a<a1; ++a) {
        std::size_t i, j, k;
        if (dim == 0) { i = fixed; j = a; k = b; } else if (dim == 1) { i = a; j = fixed; k = b; } else { i = a; j = b; k = fixed; }
        for (int t=0; t<6; ++t)
          for (int f=0; f<4; ++f) {
            bool on = true;
            for (int v=0; v<3; ++v) if (kuhn.c[t][tf[f][v]][dim] != side) on = false;
            if (!on) continue;
            for (int v=0; v<3; ++v) {
              const int* c = kuhn.c[t][tf[f][v]];
              out.push_back( id( i+static_cast<std::size_t>(c[0]), j+static_cast<std::size_t>(c[1]), k+static_cast<std::size_t>(c[2]) ) );
            }
          }
      }
    if (out.empty()) m.sidetri.erase( set );
  };
  if (i0 == 0)  addfaces( 1, 0, 0, j0, j1, k0, k1, 0 );
  if (i1 == nx) addfaces( 2, 0, 1, j0, j1, k0, k1, nx-1 );
  if (j0 == 0)  addfaces( 3, 1, 0, i0, i1, k0, k1, 0 );
  if (j1 == ny) addfaces( 4, 1, 1, i0, i1, k0, k1, ny-1 );
  if (k0 == 0)  addfaces( 5, 2, 0, i0, i1, j0, j1, 0 );
  if (k1 == nz) addfaces( 6, 2, 1, i0, i1, j0, j1, nz-1 );
  return m;
}

// ---- recursive coordinate bisection as the reference gets it from Zoltan ---------------------------------
// inciter::geomPartMesh (src/Partition/ZoltanGeom.cpp:139-244) hands the element centroids to Zoltan 3.901
// (vendored under src/zoltan) with LB_METHOD RCB, LB_APPROACH PARTITION, no weights, AVERAGE_CUTS 1. What follows
// restates the algorithm Zoltan runs for that: serial_rcb (src/zoltan/src/rcb/rcb.c:1516-1860) with cut_dimension
// (:1380-1413), Zoltan_Divide_Parts (ha/divide_machine.c), Zoltan_RB_find_median (par/par_median.c:88-560) and
// Zoltan_RB_Average_Cut (par/par_average.c), for unit weights on one process. tests/test_oracle_zoltan.py compares
// it element by element with the reference's own Zoltan sources compiled into oracle/_ref.
namespace {

struct RcbBox { real lo[3], hi[3]; };

// Zoltan_RB_find_median for unit weights: marks every dot 0 (lower set) or 1; returns the cut after averaging
real zoltanMedian( const std::vector< real >& dots, real fractionlo, real valuemin, real valuemax, real weight,
                   std::vector< int >& dotmark, real& wlo, real& whi )
{
  const real TINY = 1.0e-6, DMAX = std::numeric_limits< real >::max();
  const auto dotnum = dots.size();
  std::vector< std::size_t > dotlist( dotnum );
  std::iota( dotlist.begin(), dotlist.end(), 0 );
  auto numlist = dotnum;
  const real tolerance = 1.0 * (0.5 + TINY);
  const real targetlo = fractionlo * weight, targethi = weight - targetlo;
  real weightlo = 0.0, weighthi = 0.0, tmp_half = 0.0;
  std::size_t indexlo = 0, indexhi = 0;
  while (true) {
    if (weight != 0.0)
      tmp_half = valuemin + (targetlo - weightlo) / (weight - weightlo - weighthi) * (valuemax - valuemin);
    else
      tmp_half = 0.5 * (valuemin + valuemax);
    real totallo = 0.0, totalhi = 0.0, valuelo = -DMAX, valuehi = DMAX, wtlo = 0.0, wthi = 0.0;
    int countlo = 0, counthi = 0;
    for (std::size_t j=0; j<numlist; ++j) {
      auto i = dotlist[j];
      if (dots[i] <= tmp_half) {
        totallo += 1.0; dotmark[i] = 0;
        if (dots[i] > valuelo) { valuelo = dots[i]; wtlo = 1.0; countlo = 1; indexlo = i; }
        else if (dots[i] == valuelo) { wtlo += 1.0; ++countlo; }
      } else {
        totalhi += 1.0; dotmark[i] = 1;
        if (dots[i] < valuehi) { valuehi = dots[i]; wthi = 1.0; counthi = 1; indexhi = i; }
        else if (dots[i] == valuehi) { wthi += 1.0; ++counthi; }
      }
    }
    int markactive;
    if (weightlo + totallo < targetlo) {                 // lower half too small
      weightlo += totallo;
      tmp_half = valuehi;
      if (counthi == 1) {
        if (weightlo + wthi < targetlo) dotmark[indexhi] = 0;
        else {
          if (weightlo + wthi - targetlo < targetlo - weightlo) { dotmark[indexhi] = 0; weightlo += wthi; }
          weighthi = weight - weightlo;
          break;
        }
      } else {
        bool done = false;
        real wtok = wthi;                                // one process: all dots at valuehi are mine
        if (weightlo + wthi >= targetlo) {
          real wtmax = targetlo - weightlo;
          if (wtok > wtmax) wtok = wtok - (wtok - wtmax);
          done = true;
        }
        real wtsum = 0.0;
        for (std::size_t j=0; j<numlist; ++j) {
          auto i = dotlist[j];
          if (dots[i] == valuehi && (wtsum + 1.0 - wtok < wtok - wtsum)) { dotmark[i] = 0; wtsum += 1.0; }
        }
        if (done) { weightlo += wtsum; weighthi = weight - weightlo; break; }
      }
      weightlo += wthi;
      if (targetlo - weightlo <= tolerance) { weighthi = weight - weightlo; break; }
      valuemin = valuehi;
      markactive = 1;
    }
    else if (weighthi + totalhi < targethi) {            // upper half too small
      weighthi += totalhi;
      tmp_half = valuelo;
      if (countlo == 1) {
        if (weighthi + wtlo < targethi) dotmark[indexlo] = 1;
        else {
          if (weighthi + wtlo - targethi < targethi - weighthi) { dotmark[indexlo] = 1; weighthi += wtlo; }
          weightlo = weight - weighthi;
          break;
        }
      } else {
        bool done = false;
        real wtok = wtlo;
        if (weighthi + wtlo >= targethi) {
          real wtmax = targethi - weighthi;
          if (wtok > wtmax) wtok = wtok - (wtok - wtmax);
          done = true;
        }
        real wtsum = 0.0;
        for (std::size_t j=0; j<numlist; ++j) {
          auto i = dotlist[j];
          if (dots[i] == valuelo && (wtsum + 1.0 - wtok < wtok - wtsum)) { dotmark[i] = 1; wtsum += 1.0; }
        }
        if (done) { weighthi += wtsum; weightlo = weight - weighthi; break; }
      }
      weighthi += wtlo;
      if (targethi - weighthi <= tolerance) { weightlo = weight - weighthi; break; }
      valuemax = valuelo;
      markactive = 0;
    }
    else { weightlo += totallo; weighthi += totalhi; break; }
    std::size_t k = 0;
    for (std::size_t j=0; j<numlist; ++j) { auto i = dotlist[j]; if (dotmark[i] == markactive) dotlist[k++] = i; }
    numlist = k;
  }
  // Zoltan_RB_Average_Cut: halfway between the closest dots of the two sets
  real v0 = -DMAX, v1 = DMAX;
  for (std::size_t i=0; i<dotnum; ++i) { if (dotmark[i] == 0) { if (dots[i] > v0) v0 = dots[i]; } else if (dots[i] < v1) v1 = dots[i]; }
  wlo = weightlo; whi = weighthi;
  return 0.5 * (v0 + v1);
}

void zoltanSerialRcb( const std::array< std::vector< real >, 3 >& cen, RcbBox box, real weight, std::size_t* dindx,
                      std::size_t* tmpdindx, std::size_t dotnum, int num_parts, int partlower, std::vector< int >& part )
{
  if (num_parts == 1) { for (std::size_t i=0; i<dotnum; ++i) part[dindx[i]] = partlower; return; }
  // Zoltan_Divide_Parts with uniform part sizes
  int partmid = partlower + (num_parts - 1)/2 + 1;
  real fractionlo = 0.0, sum = 0.0;
  for (int i=0; i<num_parts; ++i) { if (partlower + i < partmid) fractionlo += 1.0; sum += 1.0; }
  fractionlo /= sum;
  // cut_dimension: the longest side of the (cut, not recomputed) box; ties go to the lower dimension
  int dim = 0;
  if (box.hi[1] - box.lo[1] > box.hi[0] - box.lo[0]) dim = 1;
  if (dim == 0 && box.hi[2] - box.lo[2] > box.hi[0] - box.lo[0]) dim = 2;
  if (dim == 1 && box.hi[2] - box.lo[2] > box.hi[1] - box.lo[1]) dim = 2;
  std::vector< real > coord( dotnum );
  for (std::size_t i=0; i<dotnum; ++i) coord[i] = cen[static_cast< std::size_t >( dim )][dindx[i]];
  std::vector< int > dotmark( dotnum, 0 );
  real wlo, whi;
  auto D = static_cast< std::size_t >( dim );
  real valuehalf = zoltanMedian( coord, fractionlo, box.lo[D], box.hi[D], weight, dotmark, wlo, whi );
  // set 0 dots in order at the front, set 1 dots from the back (rcb.c:1797-1803)
  std::size_t set0 = 0, set1 = dotnum;
  for (std::size_t i=0; i<dotnum; ++i) { if (dotmark[i] == 0) tmpdindx[set0++] = dindx[i]; else tmpdindx[--set1] = dindx[i]; }
  std::copy( tmpdindx, tmpdindx + dotnum, dindx );
  int nlo = partmid - partlower;
  if (nlo > 0 && set1 != 0) { auto b = box; b.hi[D] = valuehalf;
    zoltanSerialRcb( cen, b, wlo, dindx, tmpdindx, set0, nlo, partlower, part ); }
  int nhi = partlower + num_parts - partmid;
  if (nhi > 0 && set0 != dotnum) { auto b = box; b.lo[D] = valuehalf;
    zoltanSerialRcb( cen, b, whi, dindx + set1, tmpdindx + set1, dotnum - set1, nhi, partmid, part ); }
}

// ---- recursive inertial bisection (part = "rib"): Zoltan's serial_rib (src/zoltan/src/rcb/rib.c:1060-1215) with the
// direction of Zoltan_RIB_inertial3d (rcb/inertial3d.c:69-218): centre of mass, inertia tensor, eigenvector of its
// largest eigenvalue (roots of the characteristic cubic :221-313, eigenvector by pivoted elimination :329-455), the dots
// projected on it, then the same median search, set ordering and recursion as RCB.

// largest root of the characteristic polynomial of a symmetric 3x3 matrix (Zoltan_evals3)
real ribLargestEigenvalue( const real H[3][3] )
{
  real xmax = 0.0;
  for (int i=0; i<3; ++i) for (int j=i; j<3; ++j) xmax = std::max( xmax, std::abs( H[i][j] ) );
  real m[3][3];
  for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) m[i][j] = xmax != 0.0 ? H[i][j] / xmax : H[i][j];
  const real a1 = -(m[0][0] + m[1][1] + m[2][2]);
  const real a2 = (m[0][0]*m[1][1] - m[0][1]*m[1][0]) + (m[0][0]*m[2][2] - m[0][2]*m[2][0]) + (m[1][1]*m[2][2] - m[1][2]*m[2][1]);
  const real det = m[0][0]*m[1][1]*m[2][2] + m[0][1]*m[1][2]*m[2][0] + m[0][2]*m[1][0]*m[2][1]
                 - m[0][1]*m[1][0]*m[2][2] - m[0][0]*m[1][2]*m[2][1] - m[0][2]*m[1][1]*m[2][0];
  const real a3 = -det;
  auto sgn = []( real x ){ return x >= 0 ? 1.0 : -1.0; };
  real r1, r2, r3;
  if (a3 == 0) {                                     // one root is zero: the quadratic
    r1 = 0;
    real q = -.5 * (a1 + sgn(a1) * std::sqrt( std::max( 0.0, a1*a1 - 4*a2 ) ));
    r2 = q; r3 = a2 / q;
  } else {
    real q = (a1*a1 - 3*a2) / 9;
    real r = (2*a1*a1*a1 - 9*a1*a2 + 27*a3) / 54;
    real q3 = q*q*q, rr = r*r;
    const real tol = 1.0e-6, HALFPI = 1.570796327, TWOPI = 6.283185307;       // (the constants as Zoltan has them)
    if (q3 < rr && std::abs( q3 - rr ) < tol * (std::abs( q3 ) + std::abs( rr ))) q3 = rr;
    if (q3 >= rr) {                                  // three real roots
      real theta;
      if (r == 0) theta = HALFPI;
      else { q3 = std::sqrt( q3 ); if (q3 < std::abs( r )) q3 = std::abs( r ); theta = std::acos( r / q3 ); }
      q = -2 * std::sqrt( q );
      r1 = q * std::cos( theta / 3 ) - a1 / 3;
      r2 = q * std::cos( (theta + TWOPI) / 3 ) - a1 / 3;
      r3 = q * std::cos( (theta + 2 * TWOPI) / 3 ) - a1 / 3;
    } else {                                         // one real root
      real theta = std::sqrt( rr - q3 ) + std::abs( r );
      theta = std::pow( theta, 1.0 / 3.0 );
      r1 = r2 = r3 = -sgn(r) * (theta + q / theta) - a1 / 3;
    }
  }
  r1 *= xmax; r2 *= xmax; r3 *= xmax;
  return std::max( std::max( r1, r2 ), r3 );
}

// eigenvector of A for the eigenvalue ev by fully pivoted elimination (Zoltan_eigenvec3), normalised
void ribEigenvector( const real A[3][3], real ev, real evec[3] )
{
  real m[3][3];
  for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) m[i][j] = A[i][j];
  for (int i=0; i<3; ++i) m[i][i] -= ev;
  int ind[3] = { 0, 1, 2 }, imax = 0, jmax = 0;
  real xmax = 0.0;
  for (int i=0; i<3; ++i) for (int j=i; j<3; ++j) if (std::abs( m[i][j] ) > xmax) { imax = i; jmax = j; xmax = std::abs( m[i][j] ); }
  if (xmax == 0.0) { evec[0] = 1.0; evec[1] = evec[2] = 0.0; }
  else {
    for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) m[i][j] /= xmax;
    if (imax != 0) for (int j=0; j<3; ++j) std::swap( m[0][j], m[imax][j] );
    if (jmax != 0) { for (int i=0; i<3; ++i) std::swap( m[i][0], m[i][jmax] ); ind[0] = jmax; ind[jmax] = 0; }
    for (int i=1; i<3; ++i) for (int j=1; j<3; ++j) m[i][j] = m[0][0] * m[i][j] - m[i][0] * m[0][j];
    xmax = 0.0;
    for (int i=1; i<3; ++i) for (int j=i; j<3; ++j) if (std::abs( m[i][j] ) > xmax) { imax = i; jmax = j; xmax = std::abs( m[i][j] ); }
    real ex, ey, ez;
    if (xmax < 1.0e-6) { ey = 1.0; ex = ez = 0; }                              // two-fold degenerate
    else {
      if (imax != 1) for (int j=0; j<3; ++j) std::swap( m[1][j], m[imax][j] );
      if (jmax != 1) { for (int i=0; i<3; ++i) std::swap( m[i][1], m[i][2] ); std::swap( ind[1], ind[2] ); }
      ez = m[0][0] * m[1][1];
      ey = -m[1][2] * m[0][0];
      ex = m[0][1] * m[1][2] - m[0][2] * m[1][1];
    }
    evec[ind[0]] = ex; evec[ind[1]] = ey; evec[ind[2]] = ez;
  }
  real norm = std::sqrt( evec[0]*evec[0] + evec[1]*evec[1] + evec[2]*evec[2] );
  for (int i=0; i<3; ++i) evec[i] /= norm;
}

void zoltanSerialRib( const std::array< std::vector< real >, 3 >& cen, real weight, std::size_t* dindx, std::size_t* tmpdindx,
                      std::size_t dotnum, int num_parts, int partlower, std::vector< int >& part )
{
  if (num_parts == 1) { for (std::size_t i=0; i<dotnum; ++i) part[dindx[i]] = partlower; return; }
  int partmid = partlower + (num_parts - 1)/2 + 1;
  real fractionlo = 0.0, sum = 0.0;
  for (int i=0; i<num_parts; ++i) { if (partlower + i < partmid) fractionlo += 1.0; sum += 1.0; }
  fractionlo /= sum;
  // centre of mass, inertia tensor, principal direction, projections (unit weights)
  real cm[3] = { 0.0, 0.0, 0.0 };
  for (std::size_t j=0; j<dotnum; ++j) { auto i = dindx[j]; cm[0] += cen[0][i]; cm[1] += cen[1][i]; cm[2] += cen[2][i]; }
  const real wsum = static_cast< real >( dotnum );
  for (auto& c : cm) c /= wsum;
  real xx = 0, yy = 0, zz = 0, xy = 0, xz = 0, yz = 0;
  for (std::size_t j=0; j<dotnum; ++j) { auto i = dindx[j];
    real dx = cen[0][i] - cm[0], dy = cen[1][i] - cm[1], dz = cen[2][i] - cm[2];
    xx += dx*dx; yy += dy*dy; zz += dz*dz; xy += dx*dy; xz += dx*dz; yz += dy*dz; }
  real T[3][3] = { { xx, xy, xz }, { xy, yy, yz }, { xz, yz, zz } };
  real evec[3];
  ribEigenvector( T, ribLargestEigenvalue( T ), evec );
  std::vector< real > value( dotnum );
  real vlo = std::numeric_limits< real >::max(), vhi = -std::numeric_limits< real >::max();
  for (std::size_t j=0; j<dotnum; ++j) { auto i = dindx[j];
    value[j] = (cen[0][i] - cm[0])*evec[0] + (cen[1][i] - cm[1])*evec[1] + (cen[2][i] - cm[2])*evec[2];
    vlo = std::min( vlo, value[j] ); vhi = std::max( vhi, value[j] ); }
  std::vector< int > dotmark( dotnum, 0 );
  real wlo, whi;
  zoltanMedian( value, fractionlo, vlo, vhi, weight, dotmark, wlo, whi );
  std::size_t set0 = 0, set1 = dotnum;
  for (std::size_t i=0; i<dotnum; ++i) { if (dotmark[i] == 0) tmpdindx[set0++] = dindx[i]; else tmpdindx[--set1] = dindx[i]; }
  std::copy( tmpdindx, tmpdindx + dotnum, dindx );
  int nlo = partmid - partlower;
  if (nlo > 0 && set1 != 0) zoltanSerialRib( cen, wlo, dindx, tmpdindx, set0, nlo, partlower, part );
  int nhi = partlower + num_parts - partmid;
  if (nhi > 0 && set0 != dotnum) zoltanSerialRib( cen, whi, dindx + set1, tmpdindx + set1, dotnum - set0, nhi, partmid, part );
}

std::array< std::vector< real >, 3 > centroidsOf( const Coords& coord, const std::vector< std::size_t >& ginpoel )
{
  std::size_t nel = ginpoel.size()/4;
  std::array< std::vector< real >, 3 > cen;
  for (auto& c : cen) c.resize( nel );
  for (std::size_t e=0; e<nel; ++e)
    for (std::size_t d=0; d<3; ++d) {
      const auto N = ginpoel.data() + e*4;
      cen[d][e] = (coord[d][N[0]] + coord[d][N[1]] + coord[d][N[2]] + coord[d][N[3]]) / 4.0;     // ZoltanGeom.cpp:133-135
    }
  return cen;
}

} // namespace

std::vector< int > rib( const Coords& coord, const std::vector< std::size_t >& ginpoel, int nparts )
{
  if (nparts < 1) throw std::runtime_error( "rib: nparts must be positive" );
  auto cen = centroidsOf( coord, ginpoel );
  std::size_t nel = ginpoel.size()/4;
  std::vector< int > part( nel, 0 );
  if (nel == 0 || nparts == 1) return part;
  std::vector< std::size_t > dindx( 2*nel );
  std::iota( dindx.begin(), dindx.begin() + static_cast< std::ptrdiff_t >( nel ), 0 );
  zoltanSerialRib( cen, static_cast< real >( nel ), dindx.data(), dindx.data() + nel, nel, nparts, 0, part );
  return part;
}

std::vector< int > rcb( const Coords& coord, const std::vector< std::size_t >& ginpoel, int nparts )
{
  if (nparts < 1) throw std::runtime_error( "rcb: nparts must be positive" );
  std::size_t nel = ginpoel.size()/4;
  auto cen = centroidsOf( coord, ginpoel );
  std::vector< int > part( nel, 0 );
  if (nel == 0 || nparts == 1) return part;
  RcbBox box;
  for (std::size_t d=0; d<3; ++d) {
    box.lo[d] = *std::min_element( cen[d].begin(), cen[d].end() );
    box.hi[d] = *std::max_element( cen[d].begin(), cen[d].end() );
  }
  std::vector< std::size_t > dindx( 2*nel );
  std::iota( dindx.begin(), dindx.begin() + static_cast< std::ptrdiff_t >( nel ), 0 );
  zoltanSerialRcb( cen, box, static_cast< real >( nel ), dindx.data(), dindx.data() + nel, nel, nparts, 0, part );
  return part;
}

std::array< std::size_t, 6 > boxPartRange( std::size_t nx, std::size_t ny, std::size_t nz, int nparts, int part )
{
  std::size_t lo[3] = { 0, 0, 0 }, hi[3] = { nx, ny, nz };
  int np = nparts, p = part;
  while (np > 1) {
    int dim = 0;
    for (int d=1; d<3; ++d) if (hi[d]-lo[d] > hi[dim]-lo[dim]) dim = d;
    std::size_t ext = hi[dim]-lo[dim];
    if (ext % 2) throw std::runtime_error( "boxPartRange: extent not divisible by 2" );
    std::size_t mid = lo[dim] + ext/2;
    if (p < np/2) hi[dim] = mid; else { lo[dim] = mid; p -= np/2; }
    np /= 2;
  }
  return {{ lo[0], hi[0], lo[1], hi[1], lo[2], hi[2] }};
}

EdgeCSR uniqueEdges( const std::vector< std::size_t >& inpoel, std::size_t npoin )
{
  const std::size_t nel = inpoel.size()/4;
  std::vector< std::size_t > cnt( npoin+1, 0 );
  for (std::size_t e=0; e<nel; ++e) {
    const auto N = inpoel.data() + e*4;
    for (const auto& pq : lpoed) ++cnt[ std::min( N[pq[0]], N[pq[1]] ) + 1 ];
  }
  for (std::size_t p=0; p<npoin; ++p) cnt[p+1] += cnt[p];
  std::vector< std::uint32_t > tmp( cnt[npoin] );
  {
    std::vector< std::size_t > fill( cnt.begin(), cnt.end()-1 );
    for (std::size_t e=0; e<nel; ++e) {
      const auto N = inpoel.data() + e*4;
      for (const auto& pq : lpoed) {
        auto a = N[pq[0]], b = N[pq[1]];
        tmp[ fill[ std::min(a,b) ]++ ] = static_cast< std::uint32_t >( std::max(a,b) );
      }
    }
  }
  EdgeCSR out;
  out.off.assign( npoin+1, 0 );
  #pragma omp parallel for schedule(dynamic,4096)
  for (std::size_t p=0; p<npoin; ++p) {
    auto b = tmp.begin()+static_cast<std::ptrdiff_t>(cnt[p]), e = tmp.begin()+static_cast<std::ptrdiff_t>(cnt[p+1]);
    std::sort( b, e );
    out.off[p+1] = static_cast< std::size_t >( std::unique( b, e ) - b );
  }
  for (std::size_t p=0; p<npoin; ++p) out.off[p+1] += out.off[p];
  out.hi.resize( out.off[npoin] );
  #pragma omp parallel for schedule(dynamic,4096)
  for (std::size_t p=0; p<npoin; ++p)
    std::copy( tmp.begin()+static_cast<std::ptrdiff_t>(cnt[p]),
               tmp.begin()+static_cast<std::ptrdiff_t>(cnt[p] + (out.off[p+1]-out.off[p])),
               out.hi.begin()+static_cast<std::ptrdiff_t>(out.off[p]) );
  return out;
}

void psupFromEdges( const EdgeCSR& e, std::size_t npoin,
                    std::vector< std::size_t >& off, std::vector< std::uint32_t >& nbr )
{
  off.assign( npoin+1, 0 );
  for (std::size_t p=0; p<npoin; ++p) {
    off[p+1] += e.off[p+1]-e.off[p];
    for (auto i=e.off[p]; i<e.off[p+1]; ++i) ++off[ e.hi[i]+1 ];
  }
  for (std::size_t p=0; p<npoin; ++p) off[p+1] += off[p];
  nbr.resize( off[npoin] );
  std::vector< std::size_t > fill( off.begin(), off.end()-1 );
  // lower neighbours arrive in ascending p, then the node's own higher neighbours
  // (ascending): every list ends up sorted ascending without a sort
  for (std::size_t p=0; p<npoin; ++p)
    for (auto i=e.off[p]; i<e.off[p+1]; ++i) nbr[ fill[ e.hi[i] ]++ ] = static_cast< std::uint32_t >( p );
  for (std::size_t p=0; p<npoin; ++p)
    for (auto i=e.off[p]; i<e.off[p+1]; ++i) nbr[ fill[p]++ ] = e.hi[i];
}

} // xyst::

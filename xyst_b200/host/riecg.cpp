// xyst_b200/host/riecg.cpp -- see riecg.hpp
#include "riecg.hpp"
#include "problems.hpp"
#include "siphash.hpp"
#include "refhashset.hpp"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <numeric>
#include <stdexcept>

namespace xyst {

namespace {
// node-ordering conventions of the reference, src/Mesh/DerivedData.hpp:36-44
const int lpofa[4][3] = { {1,2,3}, {2,0,3}, {3,0,1}, {0,2,1} };
const int lpoed[6][2] = { {0,1}, {1,2}, {2,0}, {0,3}, {1,3}, {2,3} };
const int lpoet[3][2] = { {0,1}, {1,2}, {2,0} };
// faces of a tet as categorised for side sets, src/Inciter/Partitioner.cpp:575-577
const int catfa[4][3] = { {0,2,1}, {0,1,3}, {0,3,2}, {1,2,3} };

double now() {
  return std::chrono::duration< double >( std::chrono::steady_clock::now().time_since_epoch() ).count();
}
void ck( int rc ) { if (rc) throw std::runtime_error( xyst_last_error() ); }

inline void cross( const real a[3], const real b[3], real r[3] ) {
  r[0] = a[1]*b[2] - b[1]*a[2]; r[1] = a[2]*b[0] - b[2]*a[0]; r[2] = a[0]*b[1] - b[0]*a[1];
}
}

// ---------------------------------------------------------------------------------------
// Discretization
// ---------------------------------------------------------------------------------------
Discretization::Discretization( const TetMesh& chunk, const Config& cfg )
  : m_cfg( cfg ), m_t( cfg.t0 ), m_dt( cfg.dt ), m_dtn( cfg.dt )
{
  const auto& g = chunk.ginpoel;
  if (g.empty()) throw std::runtime_error( "No elements assigned to Discretization" );
  // global2local (src/Mesh/Reorder.cpp:279-306): gid = sorted unique global ids
  std::size_t gmin = g[0], gmax = g[0];
  for (auto n : g) { gmin = std::min( gmin, n ); gmax = std::max( gmax, n ); }
  m_gmin = gmin;
  m_g2l.assign( gmax-gmin+1, 0 );
  for (auto n : g) m_g2l[n-gmin] = 1;
  std::size_t np = 0;
  for (auto& f : m_g2l) if (f) f = static_cast< std::uint32_t >( ++np );
  m_gid.resize( np );
  for (std::size_t i=0; i<m_g2l.size(); ++i) if (m_g2l[i]) m_gid[m_g2l[i]-1] = i+gmin;
  m_inpoel.resize( g.size() );
  #pragma omp parallel for schedule(static)
  for (std::size_t i=0; i<g.size(); ++i) m_inpoel[i] = m_g2l[g[i]-gmin]-1;
  // coordinates by local id (Discretization::setCoord :537-558)
  for (std::size_t d=0; d<3; ++d) m_coord[d].resize( np );
  if (chunk.gid.empty()) {
    for (std::size_t d=0; d<3; ++d) for (std::size_t i=0; i<np; ++i) m_coord[d][i] = chunk.coord[d][ m_gid[i] ];
  } else {
    // chunk.coord is indexed like chunk.gid (sorted)
    for (std::size_t i=0; i<np; ++i) {
      auto it = std::lower_bound( chunk.gid.begin(), chunk.gid.end(), m_gid[i] );
      if (it == chunk.gid.end() || *it != m_gid[i]) throw std::runtime_error( "node without coordinates" );
      auto k = static_cast< std::size_t >( it - chunk.gid.begin() );
      for (std::size_t d=0; d<3; ++d) m_coord[d][i] = chunk.coord[d][k];
    }
  }
  m_vol.assign( np, 0.0 );
  m_v.assign( np, 0.0 );
}

std::size_t Discretization::lid( std::size_t g ) const {
  if (g < m_gmin || g-m_gmin >= m_g2l.size() || !m_g2l[g-m_gmin]) throw std::runtime_error( "global id not on this partition" );
  return m_g2l[g-m_gmin]-1;
}

void Discretization::vol()
{
  const auto& x = m_coord[0]; const auto& y = m_coord[1]; const auto& z = m_coord[2];
  std::fill( m_vol.begin(), m_vol.end(), 0.0 );
  for (std::size_t e=0; e<m_inpoel.size()/4; ++e) {        // tet order = summation order of the reference
    const auto N = m_inpoel.data() + e*4;
    real ba[3] = { x[N[1]]-x[N[0]], y[N[1]]-y[N[0]], z[N[1]]-z[N[0]] },
         ca[3] = { x[N[2]]-x[N[0]], y[N[2]]-y[N[0]], z[N[2]]-z[N[0]] },
         da[3] = { x[N[3]]-x[N[0]], y[N[3]]-y[N[0]], z[N[3]]-z[N[0]] }, c[3];
    cross( ca, da, c );
    const auto J = (ba[0]*c[0] + ba[1]*c[1] + ba[2]*c[2]) / 24.0;
    if (!(J > 0)) throw std::runtime_error( "Element Jacobian non-positive" );
    for (std::size_t j=0; j<4; ++j) m_vol[N[j]] += J;
  }
  m_v = m_vol;
}

void Discretization::remap( const std::vector< std::size_t >& newid )
{
  #pragma omp parallel for schedule(static)
  for (std::size_t i=0; i<m_inpoel.size(); ++i) m_inpoel[i] = newid[m_inpoel[i]];
  for (auto& l : m_g2l) if (l) l = static_cast< std::uint32_t >( newid[l-1]+1 );
  auto permute = [&]( auto& a ){ auto b = a; for (std::size_t o=0; o<a.size(); ++o) b[newid[o]] = a[o]; a = std::move(b); };
  permute( m_gid ); permute( m_vol ); permute( m_v );
  permute( m_coord[0] ); permute( m_coord[1] ); permute( m_coord[2] );
}

void Discretization::setdt( real newdt ) {
  m_dtn = m_dt; m_dt = newdt;
  if (m_t + m_dt > m_cfg.term) m_dt = m_cfg.term - m_t;
}
void Discretization::next() { ++m_it; m_t += m_dt; }
bool Discretization::finished() const {
  auto eps = std::numeric_limits< real >::epsilon();
  return std::abs( m_t - m_cfg.term ) < eps || m_it >= m_cfg.nstep || (m_res > 0.0 && m_res < m_cfg.residual);
}

std::vector< std::size_t > Discretization::sharedNodes() const {
  std::vector< std::size_t > s;
  for (const auto& [r,g] : m_nodeCommMap) for (auto i : g) s.push_back( lid(i) );
  std::sort( s.begin(), s.end() );
  s.erase( std::unique( s.begin(), s.end() ), s.end() );
  return s;
}

// ---------------------------------------------------------------------------------------
// RieCG
// ---------------------------------------------------------------------------------------
RieCG::RieCG( Discretization& disc, const TetMesh& chunk, const Config& cfg )
  : m_disc( disc ), m_cfg( cfg ), m_sidetri( chunk.sidetri )
{
  if (cfg.solver != "riecg" && cfg.solver != "zalcg" && cfg.solver != "kozcg" && cfg.solver != "laxcg" && cfg.solver != "chocg" && cfg.solver != "lohcg")
    throw std::runtime_error( "Unknown solver: " + cfg.solver );
  m_zal = cfg.solver == "zalcg"; m_stride = m_zal ? 4 : 3; m_koz = cfg.solver == "kozcg"; m_lax = cfg.solver == "laxcg";
  m_cho = cfg.solver == "chocg"; if (m_cho) m_stride = 5;      // ChoCG::domint, ChoCG.cpp:399-446
  m_loh = cfg.solver == "lohcg"; if (m_loh) m_stride = 4;      // LohCG::domint, LohCG.cpp:407-453
  if (m_loh) { if (cfg.ncomp < 4u || cfg.ncomp > 8u) throw std::runtime_error( "LohCG: ncomp must be 4 (p,u,v,w) + at most 4 transported scalars" ); }
  else if (m_cho ? (cfg.ncomp < 3u || cfg.ncomp > 7u) : (cfg.ncomp < 5u || cfg.ncomp > 13u))
    throw std::runtime_error( m_cho ? "ChoCG: ncomp must be 3 (velocity) + at most 4 transported scalars" : "ncomp must be 5 (+ at most 8 transported scalars)" );
  if (cfg.ncomp > 5u && !m_cho && !m_loh && cfg.solver != "riecg" && cfg.solver != "kozcg" && cfg.solver != "zalcg")
    throw std::runtime_error( "transported scalars are implemented for RieCG, ZalCG, KozCG, ChoCG and LohCG only" );
  // Transporter::matchsets as the reference executes it (Transporter.cpp:125-187 with the
  // short-circuit at :347-348): with at least one side set named in the configuration the
  // faces of ALL side sets of the mesh keep their boundary integrals; with none, no face does.
  bool any = !cfg.bc_sym.empty() || !cfg.bc_far.empty() || !cfg.bc_pre.empty();
  for (const auto& d : cfg.bc_dir) if (!d.empty()) any = true;
  for (const auto& d : cfg.p_bc_dir) if (!d.empty()) any = true;
  if (!cfg.p_bc_sym.empty() || !cfg.bc_noslip.empty()) any = true;
  if (!any) m_sidetri.clear();
}

RieCG::~RieCG() { if (m_ctx) xyst_ctx_destroy( m_ctx ); }

void RieCG::attach( int device, int nranks, int rank, const void* ncclid )
{
  xyst_params p{};
  p.ncomp = m_cho || m_loh ? 5 : static_cast< int >( m_cfg.ncomp );      // (the context's Euler kernels stay unused for ChoCG)
  if (m_cfg.flux == "rusanov" || m_cho || m_loh) p.flux = 0; else if (m_cfg.flux == "hllc") p.flux = 1;
  else throw std::runtime_error( "Flux not configured" );       // Riemann.cpp:676-681
  p.stab2 = m_cfg.stab2; p.stab2coef = m_cfg.stab2coef; p.gamma = m_cfg.gamma;
  p.exact_muscl = m_cfg.exact_muscl;
  ck( xyst_ctx_create( device, &p, &m_ctx ) );
  m_nranks = nranks; m_rank = rank;
  if (nranks > 1) ck( xyst_comm_init( m_ctx, nranks, rank, ncclid ) );
  if (!m_halosum)
    m_halosum = [this]( int w, std::vector< real >& vals ){
      uploadHalo();
      ck( xyst_halo_sum( m_ctx, w, vals.data() ) ); };
  if (!m_allreduce) m_nccl_reduce = true;       // the library's own NCCL reductions (no caller-supplied hooks)
  if (!m_allreduce)
    m_allreduce = [this]( int op, std::vector< real >& v ){
      ck( op == 0 ? xyst_allreduce_sum( m_ctx, v.data(), static_cast< int >( v.size() ) )
                  : xyst_allreduce_min( m_ctx, v.data(), static_cast< int >( v.size() ) ) ); };
}

//! shared-node lists for the device (same order on both sides: ascending global id)
void RieCG::uploadHalo()
{
  if (m_haloup || m_nranks < 2 || !m_ctx) return;
  std::vector< int > nr; std::vector< std::size_t > off{ 0 }, sh;
  for (const auto& [r,g] : m_disc.NodeCommMap()) {
    nr.push_back( r );
    for (auto i : g) sh.push_back( m_disc.lid(i) );
    off.push_back( sh.size() );
  }
  ck( xyst_halo_upload( m_ctx, static_cast< int >( nr.size() ), nr.data(), off.data(), sh.data() ) );
  m_haloup = true;
}

//! Mesh-locality renumbering, RieCG.cpp:82-100: first touch order of (p, neighbours of p)
void RieCG::renumber()
{
  auto np = m_disc.Gid().size();
  auto edges = uniqueEdges( m_disc.Inpoel(), np );
  std::vector< std::size_t > off; std::vector< std::uint32_t > nbr;
  psupFromEdges( edges, np, off, nbr );
  const auto none = std::numeric_limits< std::size_t >::max();
  std::vector< std::size_t > map( np, none );
  std::size_t n = 0;
  for (std::size_t p=0; p<np; ++p) {
    if (map[p] == none) map[p] = n++;
    for (auto i=off[p]; i<off[p+1]; ++i) { auto q = nbr[i]; if (map[q] == none) map[q] = n++; }
  }
  if (n != np) throw std::runtime_error( "Mesh-locality reorder map size mismatch" );
  m_disc.remap( map );
  for (auto& t : m_triinpoel) t = map[t];
}

//! Boundary faces of this partition: for every tet, in order, those of its faces that lie
//! on a side set (Partitioner.cpp:539-626), oriented as the tet has them.
void RieCG::boundaryFaces( const TetMesh& )
{
  struct Key { std::array< std::size_t, 3 > n; int set; };
  std::vector< Key > keys;
  for (const auto& [s,tri] : m_sidetri)
    for (std::size_t f=0; f<tri.size()/3; ++f) {
      Key k{ {{ tri[f*3], tri[f*3+1], tri[f*3+2] }}, s };
      std::sort( k.n.begin(), k.n.end() );
      keys.push_back( k );
    }
  // a face listed in several sets belongs to the last one assigned (highest id)
  std::sort( keys.begin(), keys.end(), []( const Key& a, const Key& b ){ return a.n < b.n || (a.n == b.n && a.set < b.set); } );
  const auto& gid = m_disc.Gid();
  const auto& inpoel = m_disc.Inpoel();
  // flag nodes on side sets to skip interior tets quickly
  std::vector< std::uint8_t > onb( gid.size(), 0 );
  for (const auto& k : keys) for (auto g : k.n) {
    try { onb[ m_disc.lid(g) ] = 1; } catch (...) {}
  }
  std::map< int, std::vector< std::size_t > > bconn;     // set -> local node triples
  for (std::size_t e=0; e<inpoel.size()/4; ++e) {
    const auto N = inpoel.data() + e*4;
    if (onb[N[0]] + onb[N[1]] + onb[N[2]] + onb[N[3]] < 3) continue;
    for (const auto& f : catfa) {
      if (!(onb[N[f[0]]] && onb[N[f[1]]] && onb[N[f[2]]])) continue;
      std::array< std::size_t, 3 > key{{ gid[N[f[0]]], gid[N[f[1]]], gid[N[f[2]]] }};
      std::sort( key.begin(), key.end() );
      auto it = std::upper_bound( keys.begin(), keys.end(), key, []( const std::array< std::size_t, 3 >& a, const Key& b ){ return a < b.n; } );
      if (it == keys.begin() || (it-1)->n != key) continue;
      auto& s = bconn[ (it-1)->set ];
      s.push_back( N[f[0]] ); s.push_back( N[f[1]] ); s.push_back( N[f[2]] );
    }
  }
  m_bface.clear(); m_triinpoel.clear();
  std::size_t nf = 0;
  for (const auto& [s,t] : bconn) {
    auto& b = m_bface[s];
    for (std::size_t i=0; i<t.size()/3; ++i) b.push_back( nf++ );
    m_triinpoel.insert( m_triinpoel.end(), t.begin(), t.end() );
  }
}

//! Domain edge integrals, RieCG.cpp:339-382: contributions are added per edge in tet order
//! (bit-identical to the reference's hash-map accumulation); d is 3 doubles per edge,
//! oriented from the lower to the higher GLOBAL id.
void RieCG::domint( const EdgeCSR& edges, std::vector< real >& d ) const
{
  const auto& inpoel = m_disc.Inpoel(); const auto& gid = m_disc.Gid();
  const auto& x = m_disc.Coord()[0]; const auto& y = m_disc.Coord()[1]; const auto& z = m_disc.Coord()[2];
  const auto st = m_stride;
  d.assign( edges.nedge()*st, 0.0 );
  for (std::size_t e=0; e<inpoel.size()/4; ++e) {
    const auto N = inpoel.data() + e*4;
    real ba[3] = { x[N[1]]-x[N[0]], y[N[1]]-y[N[0]], z[N[1]]-z[N[0]] },
         ca[3] = { x[N[2]]-x[N[0]], y[N[2]]-y[N[0]], z[N[2]]-z[N[0]] },
         da[3] = { x[N[3]]-x[N[0]], y[N[3]]-y[N[0]], z[N[3]]-z[N[0]] };
    real g[4][3];
    cross( ca, da, g[1] ); cross( da, ba, g[2] ); cross( ba, ca, g[3] );
    for (std::size_t i=0; i<3; ++i) g[0][i] = -g[1][i]-g[2][i]-g[3][i];
    real cx[3]; cross( ca, da, cx );
    const auto J120 = (ba[0]*cx[0] + ba[1]*cx[1] + ba[2]*cx[2]) / 120.0;     // ZalCG.cpp:375,386
    for (const auto& pq : lpoed) {
      auto p = pq[0], q = pq[1];
      real sig = gid[N[p]] > gid[N[q]] ? -1.0 : 1.0;
      auto n = d.data() + edges.find( N[p], N[q] )*st;
      n[0] += sig * (g[p][0] - g[q][0]) / 48.0;
      n[1] += sig * (g[p][1] - g[q][1]) / 48.0;
      n[2] += sig * (g[p][2] - g[q][2]) / 48.0;
      if (st == 4 && !m_loh) n[3] += J120;
      if (m_loh) {                                                       // LohCG.cpp:449
        auto J = ba[0]*cx[0] + ba[1]*cx[1] + ba[2]*cx[2];
        n[3] += (g[p][0]*g[q][0] + g[p][1]*g[q][1] + g[p][2]*g[q][2]) / J / 6.0;
      }
      if (st == 5) {                                                     // ChoCG.cpp:441-442
        auto J = ba[0]*cx[0] + ba[1]*cx[1] + ba[2]*cx[2];
        n[3] += J / 120.0;
        n[4] += (g[p][0]*g[q][0] + g[p][1]*g[q][1] + g[p][2]*g[q][2]) / J / 6.0;
      }
    }
  }
}

//! Superedge groups, RieCG.cpp:620-736. Tetrahedra: greedy pass in element order.
//! Triangles: the reference walks an unordered set of the faces not covered by a
//! tetrahedron superedge in hash-iteration order; because edge orientation inside a
//! triangle differs from that of a single edge and the limiter is orientation dependent at
//! 1e-9 (see siphash.hpp), the same container type, hash and insert/erase sequence are
//! used here so that the same triangles come out with the same node order. Leftover edges:
//! low -> high global id as in the reference (their order does not matter to the device).
void RieCG::domsuped( const EdgeCSR& edges, const std::vector< real >& d )
{
  const auto& inpoel = m_disc.Inpoel(); const auto& gid = m_disc.Gid();
  for (auto& a : m_dsupedge) a.clear();
  for (auto& a : m_dsupint) a.clear();
  std::vector< std::uint8_t > claimed( edges.nedge(), 0 );
  auto ntet = inpoel.size()/4;
  // reforder: 0 = walk the faces in element order (same edges and integrals, other triangles;
  // ~10x faster at 10^8 tets), anything else = the reference's hash order. No size-dependent default.
  const bool reforder = m_cfg.reforder != 0;
  // hashes of all tet faces in parallel, then the (inherently sequential) set operations
  std::vector< std::uint64_t > fh( reforder ? ntet*4 : 0 );
  if (reforder) {
  #pragma omp parallel for schedule(static)
  for (std::size_t e=0; e<ntet; ++e) {
    const auto N = inpoel.data() + e*4;
    for (std::size_t k=0; k<4; ++k) fh[e*4+k] = IdHash<3>()( {{ N[lpofa[k][0]], N[lpofa[k][1]], N[lpofa[k][2]] }} );
  }
  }
  RefOrderFaceSet untri( reforder ? ntet*2 + ntet/8 : 0 );
  if (reforder)
  for (std::size_t e=0; e<ntet; ++e) {
    const auto N = inpoel.data() + e*4;
    for (std::size_t k=0; k<4; ++k) untri.insert( {{ N[lpofa[k][0]], N[lpofa[k][1]], N[lpofa[k][2]] }}, fh[e*4+k] );
  }
  std::vector< std::uint64_t >().swap( fh );
  for (std::size_t e=0; e<ntet; ++e) {
    const auto N = inpoel.data() + e*4;
    std::size_t id[6]; bool all = true;
    for (int k=0; k<6; ++k) { id[k] = edges.find( N[lpoed[k][0]], N[lpoed[k][1]] ); if (claimed[id[k]]) { all = false; break; } }
    if (!all) continue;
    for (int k=0; k<4; ++k) m_dsupedge[0].push_back( N[k] );
    if (reforder) for (const auto& f : lpofa) untri.erase( {{ N[f[0]], N[f[1]], N[f[2]] }} );
    for (int k=0; k<6; ++k) {
      real sig = gid[N[lpoed[k][0]]] < gid[N[lpoed[k][1]]] ? 1.0 : -1.0;
      for (std::size_t j=0; j<3; ++j) m_dsupint[0].push_back( sig * d[id[k]*m_stride+j] );
      for (std::size_t j=3; j<m_stride; ++j) m_dsupint[0].push_back( d[id[k]*m_stride+j] );
      claimed[id[k]] = 1;
    }
  }
  auto tryface = [&]( const RefOrderFaceSet::Face& T ) {
    std::size_t id[3]; bool all = true;
    for (int k=0; k<3; ++k) { id[k] = edges.find( T[static_cast<std::size_t>(lpoet[k][0])], T[static_cast<std::size_t>(lpoet[k][1])] ); if (claimed[id[k]]) { all = false; break; } }
    if (!all) return;
    for (std::size_t k=0; k<3; ++k) m_dsupedge[1].push_back( T[k] );
    for (int k=0; k<3; ++k) {
      real sig = gid[T[static_cast<std::size_t>(lpoet[k][0])]] < gid[T[static_cast<std::size_t>(lpoet[k][1])]] ? 1.0 : -1.0;
      for (std::size_t j=0; j<3; ++j) m_dsupint[1].push_back( sig * d[id[k]*m_stride+j] );
      for (std::size_t j=3; j<m_stride; ++j) m_dsupint[1].push_back( d[id[k]*m_stride+j] );
      claimed[id[k]] = 1;
    }
  };
  if (reforder) untri.forEach( tryface );
  else
    for (std::size_t e=0; e<ntet; ++e) {        // element order: a face is met first at its first tet
      const auto N = inpoel.data() + e*4;
      for (const auto& f : lpofa) tryface( {{ N[f[0]], N[f[1]], N[f[2]] }} );
    }
  for (std::size_t p=0; p+1<edges.off.size(); ++p)
    for (auto i=edges.off[p]; i<edges.off[p+1]; ++i) {
      if (claimed[i]) continue;
      std::size_t a = p, b = edges.hi[i];
      if (gid[a] > gid[b]) std::swap( a, b );           // low gid -> high gid (:706-714)
      m_dsupedge[2].push_back( a ); m_dsupedge[2].push_back( b );
      for (std::size_t j=0; j<m_stride; ++j) m_dsupint[2].push_back( d[i*m_stride+j] );
    }
}

//! Boundary point normals, RieCG.cpp:281-337 (inverse-distance-squared weighted face normals)
void RieCG::bndint()
{
  const auto& x = m_disc.Coord()[0]; const auto& y = m_disc.Coord()[1]; const auto& z = m_disc.Coord()[2];
  m_bnorm.clear();
  for (const auto& [ setid, faceids ] : m_bface)
    for (auto f : faceids) {
      const auto N = m_triinpoel.data() + f*3;
      real ba[3] = { x[N[1]]-x[N[0]], y[N[1]]-y[N[0]], z[N[1]]-z[N[0]] },
           ca[3] = { x[N[2]]-x[N[0]], y[N[2]]-y[N[0]], z[N[2]]-z[N[0]] }, n[3];
      cross( ba, ca, n );
      auto A2 = std::sqrt( n[0]*n[0] + n[1]*n[1] + n[2]*n[2] );
      n[0] /= A2; n[1] /= A2; n[2] /= A2;
      const real centroid[3] = { (x[N[0]] + x[N[1]] + x[N[2]]) / 3.0,
                                 (y[N[0]] + y[N[1]] + y[N[2]]) / 3.0,
                                 (z[N[0]] + z[N[1]] + z[N[2]]) / 3.0 };
      for (const auto& ij : lpoet) {
        auto p = N[ ij[0] ];
        real r = 1.0 / ( (centroid[0] - x[p]) * (centroid[0] - x[p]) +
                         (centroid[1] - y[p]) * (centroid[1] - y[p]) +
                         (centroid[2] - z[p]) * (centroid[2] - z[p]) );
        auto& bpn = m_bnorm[setid][p];
        bpn[0] += r * n[0]; bpn[1] += r * n[1]; bpn[2] += r * n[2]; bpn[3] += r;
      }
    }
}

//! RieCG::setupBC :109-245 and the list-building half of streamable() :527-618
void RieCG::setupBC()
{
  auto ncomp = m_cfg.ncomp;
  auto facenodes = [&]( int s, std::set< std::size_t >& out ) {
    auto k = m_bface.find( s );
    if (k != m_bface.end()) for (auto f : k->second) for (int j=0; j<3; ++j) out.insert( m_triinpoel[f*3+static_cast<std::size_t>(j)] );
  };
  if (m_cho || m_loh) choSetupBC();          // LohCG::setupDirBC :215-296 is ChoCG's
  // Dirichlet: node -> mask (0 -> 1 overwrite only)
  std::map< std::size_t, std::vector< int > > dirbcset;
  if (!m_cho && !m_loh) for (const auto& mask : m_cfg.bc_dir) {
    if (mask.size() != ncomp+1) throw std::runtime_error( "Incorrect Dirichlet BC mask ncomp" );
    std::set< std::size_t > nodes;
    facenodes( mask[0], nodes );
    for (auto p : nodes) {
      auto& m = dirbcset[p];
      if (m.empty()) m.resize( ncomp, 0 );
      for (std::size_t c=0; c<ncomp; ++c) if (!m[c]) m[c] = mask[c+1];
    }
  }
  if (!m_cho && !m_loh) m_dirbcmasks.clear();
  for (const auto& [p,mask] : dirbcset) { m_dirbcmasks.push_back( p ); for (auto m : mask) m_dirbcmasks.push_back( static_cast< std::size_t >( m ) ); }
  // pressure BC
  m_prebcnodes.clear(); m_prebcvals.clear();
  for (std::size_t i=0; i<m_cfg.bc_pre.size(); ++i) {
    std::set< std::size_t > nodes;
    facenodes( m_cfg.bc_pre[i], nodes );
    for (auto p : nodes) { m_prebcnodes.push_back( p ); m_prebcvals.push_back( m_cfg.pre_density[i] ); m_prebcvals.push_back( m_cfg.pre_pressure[i] ); }
  }
  // symmetry / farfield node sets; farfield wins (:243-244)
  m_symbcnodeset.clear(); m_farbcnodeset.clear();
  for (auto s : m_cfg.bc_sym) facenodes( s, m_symbcnodeset );
  for (auto s : m_cfg.bc_far) facenodes( s, m_farbcnodeset );
  for (auto i : m_farbcnodeset) m_symbcnodeset.erase( i );
  // boundary-element symmetry flags (:534-538)
  m_besym.resize( m_triinpoel.size() );
  for (std::size_t i=0; i<m_triinpoel.size(); ++i) m_besym[i] = static_cast< std::uint8_t >( m_symbcnodeset.count( m_triinpoel[i] ) );
  // streamable symmetry / farfield lists: one entry per (node, side set with a normal there) (:581-615)
  auto stream = [&]( const std::set< std::size_t >& nodes, const std::vector< int >& sets,
                     std::vector< std::size_t >& outn, std::vector< real >& outv ) {
    outn.clear(); outv.clear();
    for (auto p : nodes)
      for (auto s : sets) {
        auto m = m_bnorm.find( s );
        if (m == m_bnorm.end()) continue;
        auto r = m->second.find( p );
        if (r == m->second.end()) continue;
        outn.push_back( p );
        outv.push_back( r->second[0] ); outv.push_back( r->second[1] ); outv.push_back( r->second[2] );
      }
  };
  stream( m_symbcnodeset, m_cfg.bc_sym, m_symbcnodes, m_symbcnorms );
  stream( m_farbcnodeset, m_cfg.bc_far, m_farbcnodes, m_farbcnorms );
}

void RieCG::prepare()
{
  auto t0 = now();
  boundaryFaces( TetMesh{} );
  m_disc.vol();                      // the reference computes volumes before renumbering ...
  m_ownvol = 0.0;                    // ... and sums the mesh volume then (Discretization.cpp:719-721)
  for (auto v : m_disc.V()) m_ownvol += v;
  timings.push_back( now()-t0 ); t0 = now();
  if (!m_zal && !m_koz) renumber();  // ZalCG/KozCG keep the global2local order (ZalCG.cpp:82-92)
  timings.push_back( now()-t0 ); t0 = now();
  if (m_koz) {                       // KozCG::feop (KozCG.cpp:246-276): boundary integrals only
    timings.push_back( 0.0 ); timings.push_back( 0.0 );
    bndint();
    timings.push_back( now()-t0 );
    return;
  }
  auto np = m_disc.Gid().size();
  auto edges = uniqueEdges( m_disc.Inpoel(), np );
  std::vector< real > d;
  domint( edges, d );
  timings.push_back( now()-t0 ); t0 = now();
  domsuped( edges, d );
  timings.push_back( now()-t0 ); t0 = now();
  bndint();
  timings.push_back( now()-t0 );
}

void RieCG::hostSetup()
{
  if (m_nranks > 1 && (!m_halosum || !m_allreduce))
    throw std::runtime_error( "RieCG::hostSetup: no communication attached" );
  auto t0 = now();
  auto np = m_disc.Gid().size();
  auto shared = m_disc.sharedNodes();
  // --- comvol / totalvol (Discretization.cpp:663-725) ---
  real tv = m_ownvol;
  if (!shared.empty()) {
    std::vector< real > vals( shared.size() );
    for (std::size_t i=0; i<shared.size(); ++i) vals[i] = m_disc.m_v[shared[i]];
    m_halosum( 1, vals );
    for (std::size_t i=0; i<shared.size(); ++i) m_disc.m_vol[shared[i]] = vals[i];
  }
  if (m_nranks > 1) { std::vector< real > t{ tv }; m_allreduce( 0, t ); tv = t[0]; }
  m_disc.MeshVol() = tv;
  // --- comnorm (RieCG.cpp:264-277,384-407) + bnorm() normalisation (:504-513) ---
  if (m_nranks > 1) {
    // every rank walks the same list of side sets: all sets of the (global) configuration
    std::set< int > sets;
    for (auto s : m_cfg.bc_sym) sets.insert( s );
    for (auto s : m_cfg.bc_far) sets.insert( s );
    for (auto s : sets) {
      std::vector< real > vals( shared.size()*4, 0.0 );
      auto m = m_bnorm.find( s );
      if (m != m_bnorm.end())
        for (std::size_t i=0; i<shared.size(); ++i) {
          auto r = m->second.find( shared[i] );
          if (r != m->second.end()) for (std::size_t k=0; k<4; ++k) vals[i*4+k] = r->second[k];
        }
      m_halosum( 4, vals );
      for (std::size_t i=0; i<shared.size(); ++i)
        if (vals[i*4+3] != 0.0) { auto& n = m_bnorm[s][shared[i]]; for (std::size_t k=0; k<4; ++k) n[k] = vals[i*4+k]; }
    }
  }
  for (auto& [s,b] : m_bnorm) for (auto& [p,n] : b) { n[0] /= n[3]; n[1] /= n[3]; n[2] /= n[3]; }
  setupBC();
  m_bnorm.clear();
  // --- initial conditions (problems::initialize, Problems.cpp:1134-1167) ---
  const auto& co = m_disc.Coord();
  auto ncomp = m_cfg.ncomp;
  if (m_u0.size() != np*ncomp) {
    auto ic = problems::IC( m_cfg );
    m_u0.resize( np*ncomp );
    #pragma omp parallel for schedule(static)
    for (std::size_t i=0; i<np; ++i) {
      auto s = ic( co[0][i], co[1][i], co[2][i], m_disc.T() );
      for (std::size_t c=0; c<ncomp; ++c) m_u0[i*ncomp+c] = s[c];
    }
  }
  m_timedep = problems::timeDependent( m_cfg );
  if (m_timedep && (m_lax || m_cfg.steady))
    throw std::runtime_error( "time-dependent problems are not hooked up for LaxCG and steady-state runs" );
  evalDirvals( m_disc.T() );
  evalSrc( m_disc.T() );
  if (m_cho || m_loh) choPrelhs();           // LohCG::prelhs :140-181 is ChoCG's
  m_hostready = true;
  timings.push_back( now()-t0 );
}

void RieCG::evalDirvals( real t )
{
  const auto& co = m_disc.Coord();
  auto ncomp = m_cfg.ncomp;
  m_dirvals.clear();
  if (m_dirbcmasks.empty()) return;
  auto ic = problems::IC( m_cfg );
  auto nd = m_dirbcmasks.size()/(ncomp+1);
  m_dirvals.resize( nd*ncomp );
  #pragma omp parallel for schedule(static)
  for (std::size_t i=0; i<nd; ++i) {
    auto p = m_dirbcmasks[i*(ncomp+1)];
    auto s = ic( co[0][p], co[1][p], co[2][p], t );
    for (std::size_t c=0; c<ncomp; ++c) m_dirvals[i*ncomp+c] = s[c];
  }
}

void RieCG::evalSrc( real t )
{
  const auto& co = m_disc.Coord();
  auto ncomp = m_cfg.ncomp;
  auto np = m_disc.Gid().size();
  m_src.clear();
  auto src = problems::SRC( m_cfg );
  if (!src) return;
  m_src.resize( np*ncomp );
  #pragma omp parallel for schedule(static)
  for (std::size_t i=0; i<np; ++i) {
    auto s = src( co[0][i], co[1][i], co[2][i], t );
    for (std::size_t c=0; c<ncomp; ++c) m_src[i*ncomp+c] = s[c];
  }
}

//! problems::SRC at the tet centroids at time t (kozak::rhs, Kozak.cpp:160-171)
void RieCG::evalSrcCentroids( real t, std::vector< real >& sc )
{
  const auto& co = m_disc.Coord();
  const auto& inpoel = m_disc.Inpoel();
  auto ncomp = m_cfg.ncomp;
  sc.clear();
  auto src = problems::SRC( m_cfg );
  if (!src) return;
  sc.resize( inpoel.size()/4*ncomp );
  #pragma omp parallel for schedule(static)
  for (std::size_t e=0; e<inpoel.size()/4; ++e) {
    const auto N = inpoel.data() + e*4;
    auto xe = (co[0][N[0]] + co[0][N[1]] + co[0][N[2]] + co[0][N[3]]) / 4.0;
    auto ye = (co[1][N[0]] + co[1][N[1]] + co[1][N[2]] + co[1][N[3]]) / 4.0;
    auto ze = (co[2][N[0]] + co[2][N[1]] + co[2][N[2]] + co[2][N[3]]) / 4.0;
    auto v = src( xe, ye, ze, t );
    for (std::size_t c=0; c<ncomp; ++c) sc[e*ncomp+c] = v[c];
  }
}

void RieCG::setup()
{
  if (!m_ctx) throw std::runtime_error( "RieCG::setup: attach() a device first" );
  if (!m_hostready) hostSetup();
  auto t0 = now();
  if (m_cho) { choSetup(); timings.push_back( now()-t0 ); timings.push_back( 0.0 ); return; }
  if (m_loh) { lohSetup(); timings.push_back( now()-t0 ); timings.push_back( 0.0 ); return; }
  uploadHalo();
  auto np = m_disc.Gid().size();
  auto ncomp = m_cfg.ncomp;
  // --- device upload ---
  const auto& co = m_disc.Coord();
  std::size_t nsup[3] = { m_dsupedge[0].size()/4, m_dsupedge[1].size()/3, m_dsupedge[2].size()/2 };
  const std::size_t* se[3] = { m_dsupedge[0].data(), m_dsupedge[1].data(), m_dsupedge[2].data() };
  const real* si[3] = { m_dsupint[0].data(), m_dsupint[1].data(), m_dsupint[2].data() };
  if (m_koz) {
    // kozak::rhs source terms (Kozak.cpp:97-108,160-171): nodes and tet centroids
    std::vector< real > sc;
    const auto& inpoel = m_disc.Inpoel();
    evalSrcCentroids( m_disc.T(), sc );
    ck( xyst_kozcg_mesh_upload( m_ctx, np, co[0].data(), co[1].data(), co[2].data(), inpoel.size()/4, inpoel.data(),
                                m_disc.Vol().data(), m_disc.V().data(),
                                sc.empty() ? nullptr : m_src.data(), sc.empty() ? nullptr : sc.data() ) );
  } else
  ck( (m_zal ? xyst_zalcg_mesh_upload : xyst_mesh_upload)( m_ctx, np, co[0].data(), co[1].data(), co[2].data(),
                        nsup, se, si, m_triinpoel.size()/3, m_triinpoel.data(), m_besym.data(),
                        m_disc.Vol().data(), m_disc.V().data() ) );
  if (m_lax) {                                     // LaxCG.cpp:115-259
    xyst_laxcg_params lp{ m_cfg.rgas, m_cfg.turkel, { m_cfg.velinf[0], m_cfg.velinf[1], m_cfg.velinf[2] } };
    ck( xyst_laxcg_config( m_ctx, &lp ) );
  }
  if (m_cfg.steady) {
    if (m_koz) throw std::runtime_error( "steady state is a RieCG/LaxCG/ZalCG option" );
    ck( xyst_steady( m_ctx, 1 ) );
  }
  if (m_zal || m_koz) {
    xyst_zalcg_params zp{};
    zp.fct = m_cfg.fct; zp.fctclip = m_cfg.fctclip; zp.fctdif = m_cfg.fctdif;
    for (auto c : m_cfg.fctsys) zp.fctsys_mask |= 1 << (c-1);
    ck( xyst_zalcg_config( m_ctx, &zp ) );
  }
  timings.push_back( now()-t0 ); t0 = now();
  ck( xyst_bc_upload( m_ctx, m_dirbcmasks.size()/(ncomp+1), m_dirbcmasks.data(), m_dirvals.data(),
                      m_symbcnodes.size(), m_symbcnodes.data(), m_symbcnorms.data(),
                      m_farbcnodes.size(), m_farbcnodes.data(), m_farbcnorms.data(),
                      m_cfg.far_density, m_cfg.far_pressure, m_cfg.far_velocity.data(),
                      m_prebcnodes.size(), m_prebcnodes.data(), m_prebcvals.data() ) );
  if (!m_src.empty() && !m_koz && !m_zal) ck( xyst_src_upload( m_ctx, m_src.data() ) );      // (ZalCG, KozCG: per step, with the edge / centroid part)
  ck( xyst_state_set( m_ctx, m_u0.data() ) );
  BC();                                            // RieCG::merge :754
  ck( xyst_sync( m_ctx ) );
  timings.push_back( now()-t0 );
}

real RieCG::dt()
{
  real mindt;
  auto eps = std::numeric_limits< real >::epsilon();
  bool reduced = false;
  if (std::abs( m_cfg.dt ) > eps) mindt = m_cfg.dt;
  else if (m_nranks > 1 && m_nccl_reduce) { ck( xyst_dt_min_all( m_ctx, m_cfg.cfl, &mindt ) ); reduced = true; }   // contribute(min_double) :850
  else ck( xyst_dt_min( m_ctx, m_cfg.cfl, &mindt ) );
  if ((m_koz || m_zal) && !(std::abs( m_cfg.dt ) > eps)) {      // KozCG::dt :669-674, ZalCG::dt :948-952: frozen flow,
    if (m_disc.T() > m_cfg.freezetime && m_cfg.freezeflow > 1.0 && m_freezeflow <= 1.0) {      // the scalars advance with freezeflow x dt
      m_freezeflow = m_cfg.freezeflow;
      ck( (m_koz ? xyst_kozcg_freeze : xyst_zalcg_freeze)( m_ctx, 1 ) );
    }
    mindt *= m_freezeflow;
  }
  if (m_nranks > 1 && !reduced) { std::vector< real > t{ mindt }; m_allreduce( 1, t ); mindt = t[0]; }
  return mindt;
}

void RieCG::advance( real newdt )
{
  auto eps = std::numeric_limits< real >::epsilon();
  if (newdt < eps) m_finished = true;
  if (m_stage == 0) m_disc.setdt( newdt );
}

void RieCG::grad() { ck( xyst_riecg_grad( m_ctx ) ); }
void RieCG::rhs() { ck( xyst_riecg_rhs( m_ctx ) ); }
void RieCG::BC() { ck( xyst_apply_bc( m_ctx ) ); }
void RieCG::solve()
{
  ck( xyst_rk_update( m_ctx, m_stage, m_disc.Dt() ) );
  BC();
  m_stage = (m_stage + 1) % 3;
}

bool RieCG::step( std::vector< real >* diagrow )
{
  if (m_cho) return choStep( diagrow );
  if (m_loh) return lohStep( diagrow );
  if (m_finished) return false;
  // problems::point_src: active for all stages of a step that starts at or after the release time
  if (m_cfg.problem == "point_src" && m_cfg.ncomp > 5 && m_cfg.src_radius >= 0.0 && !m_pinned &&
      !(m_disc.T() < m_cfg.src_release_time)) {
    const auto& co = m_disc.Coord();
    std::vector< std::size_t > nodes;
    for (std::size_t i=0; i<co[0].size(); ++i) {
      auto rx = m_cfg.src_location[0] - co[0][i], ry = m_cfg.src_location[1] - co[1][i], rz = m_cfg.src_location[2] - co[2][i];
      if (rx*rx + ry*ry + rz*rz < m_cfg.src_radius*m_cfg.src_radius) nodes.push_back( i );
    }
    ck( xyst_scalar_pin( m_ctx, nodes.size(), nodes.data(), 1.0 ) );
    m_pinned = true;
  }
  advance( dt() );
  if (m_zal) {                                                  // ZalCG.cpp:973-1607
    if (problems::SRC( m_cfg ) && (m_timedep || m_zedge[0].empty())) {
      // zalesak::rhs source term: at the nodes at t, at the edge midpoints at t + dt/2 (Zalesak.cpp:118-128,152-163)
      if (m_zedge[0].empty()) {
        auto ns = xyst_nslot( m_ctx );
        m_zedge[0].resize( ns ); m_zedge[1].resize( ns );
        ck( xyst_edge_list( m_ctx, m_zedge[0].data(), m_zedge[1].data() ) );
      }
      evalSrc( m_disc.T() );
      auto src = problems::SRC( m_cfg );
      const auto& co = m_disc.Coord();
      const auto ncomp = m_cfg.ncomp, ns = m_zedge[0].size();
      const auto te = m_disc.T() + m_disc.Dt()/2.0;
      std::vector< real > se( ns*ncomp, 0.0 );
      #pragma omp parallel for schedule(static)
      for (std::size_t e=0; e<ns; ++e) {
        auto p = m_zedge[0][e], q = m_zedge[1][e];
        if (p == static_cast< std::size_t >( -1 )) continue;
        auto sv = src( (co[0][p] + co[0][q])/2.0, (co[1][p] + co[1][q])/2.0, (co[2][p] + co[2][q])/2.0, te );
        for (std::size_t c=0; c<ncomp; ++c) se[e*ncomp+c] = sv[c];
      }
      ck( xyst_zalcg_src( m_ctx, m_src.data(), se.data() ) );
    }
    if (m_timedep) {                                            // BC( m_a, d->T() + d->Dt() ) :1575
      evalDirvals( m_disc.T() + m_disc.Dt() );
      if (!m_dirvals.empty()) ck( xyst_dirbc_values( m_ctx, m_dirvals.data() ) );
    }
    ck( xyst_zalcg_step( m_ctx, m_disc.Dt() ) );
  }
  else if (m_koz) {                                            // KozCG.cpp:691-1197
    if (m_timedep) {
      // sources at the nodes at t and at the centroids at t + dt/2 (Kozak.cpp:104,163), Dirichlet
      // values at t + dt (KozCG::solve -> BC)
      std::vector< real > sc;
      evalSrc( m_disc.T() );
      evalSrcCentroids( m_disc.T() + m_disc.Dt()/2.0, sc );
      if (!m_src.empty()) ck( xyst_kozcg_src( m_ctx, m_src.data(), sc.data() ) );
      evalDirvals( m_disc.T() + m_disc.Dt() );
      if (!m_dirvals.empty()) ck( xyst_dirbc_values( m_ctx, m_dirvals.data() ) );
    }
    ck( xyst_kozcg_step( m_ctx, m_disc.Dt() ) );
  }
  else if (m_timedep) {
    // source at the time level of the step for all stages (RieCG.cpp:949), Dirichlet values at the
    // stage time t + rk dt (:1028): refreshed on the host, three separate stage calls
    static const real rk[3] = { 1.0/3.0, 1.0/2.0, 1.0 };
    evalSrc( m_disc.T() );
    if (!m_src.empty()) ck( xyst_src_upload( m_ctx, m_src.data() ) );
    for (int s=0; s<3; ++s) {
      evalDirvals( m_disc.T() + rk[s]*m_disc.Dt() );
      if (!m_dirvals.empty()) ck( xyst_dirbc_values( m_ctx, m_dirvals.data() ) );
      ck( xyst_riecg_stage( m_ctx, s, m_disc.Dt() ) );
    }
  }
  else ck( xyst_riecg_step( m_ctx, m_disc.Dt() ) );
  if (diagrow && (m_disc.It()+1) % m_cfg.diag_iter == 0) *diagrow = diagnostics();
  else {
    if (diagrow) diagrow->clear();
    if (m_cfg.steady) m_disc.residual( 1.0 );      // no diagnostics this step: evalres( 1.0 ), RieCG.cpp:1055
  }
  m_disc.next();
  if (m_disc.finished()) m_finished = true;
  return !m_finished;
}

std::vector< real > RieCG::diagnostics()
{
  auto ncomp = m_cfg.ncomp;
  auto sol = problems::SOL( m_cfg );
  std::vector< real > an;
  if (sol) {
    const auto& co = m_disc.Coord();
    auto np = co[0].size();
    an.resize( np*ncomp );
    for (std::size_t i=0; i<np; ++i) {
      auto s = sol( co[0][i], co[1][i], co[2][i], m_disc.T()+m_disc.Dt() );
      s[1] /= s[0]; s[2] /= s[0]; s[3] /= s[0];
      s[4] = s[4] / s[0] - 0.5*(s[1]*s[1] + s[2]*s[2] + s[3]*s[3]);
      for (std::size_t c=0; c<ncomp; ++c) an[i*ncomp+c] = s[c];
    }
  }
  std::vector< real > d( 4*ncomp+1, 0.0 );
  ck( xyst_diag( m_ctx, sol ? an.data() : nullptr, d.data() ) );
  if (m_nranks > 1) m_allreduce( 0, d );
  auto mv = m_disc.MeshVol();
  std::vector< real > row{ static_cast< real >( m_disc.It()+1 ), m_disc.T()+m_disc.Dt(), m_disc.Dt() };
  for (std::size_t i=0; i<ncomp; ++i) row.push_back( std::sqrt( d[i] / mv ) );
  for (std::size_t i=0; i<ncomp; ++i) row.push_back( std::sqrt( d[ncomp+i] / mv ) );
  row.push_back( d[2*ncomp] );
  if (m_cfg.steady) m_disc.residual( row[ 3 + ncomp + m_cfg.rescomp - 1 ] );   // evalres, RieCG.cpp:1062-1075
  if (sol) {
    for (std::size_t i=0; i<ncomp; ++i) row.push_back( std::sqrt( d[2*ncomp+1+i] / mv ) );
    for (std::size_t i=0; i<ncomp; ++i) row.push_back( d[3*ncomp+1+i] / mv );
  }
  return row;
}

std::vector< real > RieCG::solution()
{
  if (m_cho) return choGet( "u", m_cfg.ncomp );
  if (m_loh) { std::vector< real > r( m_disc.Gid().size()*m_cfg.ncomp ); ck( xyst_lohcg_get_u( m_ctx, r.data() ) ); return r; }
  std::vector< real > u( m_disc.Gid().size()*m_cfg.ncomp );
  ck( xyst_state_get( m_ctx, u.data() ) );
  return u;
}

void RieCG::setSolution( const std::vector< real >& u ) { ck( m_loh ? xyst_lohcg_set_u( m_ctx, u.data() ) : m_cho ? xyst_chocg_set_u( m_ctx, u.data() ) : xyst_state_set( m_ctx, u.data() ) ); }

} // xyst::

// xyst_b200/host/chocg.cpp -- the ChoCG members of the host mirror (solver = "chocg"):
// projection method for constant-density flow, src/Inciter/ChoCG.cpp + chocg.ci, on one or several
// partitions (the device library sums the shared nodes' parts after every operator; here: the global
// reductions and the union of the solvers' Dirichlet rows over the partitions sharing a node).
// Setup pieces the other solvers do not have: Dirichlet BCs with values and pressure BCs
// (ChoCG::setupDirBC :210-300), no-slip nodes (:655-682), the pressure Poisson matrix
// (ChoCG::prelhs :146-188 on tk::CSR, src/LinearSolver/CSR.cpp:19-84). The time step sequences
// the device entry points (xyst_chocg_*, xyst_cg_solve) the way the chare's SDAG code does;
// every nodal and edge loop runs on the device.
#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>
#include "riecg.hpp"
#include "problems.hpp"

namespace xyst {

namespace {
void ck( int rc ) { if (rc) throw std::runtime_error( xyst_last_error() ); }
const real rkcoef[4][4] = { { 1.0, 0, 0, 0 }, { 1.0/2.0, 1.0, 0, 0 }, { 1.0/3.0, 1.0/2.0, 1.0, 0 },
                            { 1.0/4.0, 1.0/3.0, 1.0/2.0, 1.0 } };           // ChoCG.cpp:43-48
inline void cross( const real a[3], const real b[3], real r[3] ) {
  r[0] = a[1]*b[2] - b[1]*a[2]; r[1] = a[2]*b[0] - b[2]*a[0]; r[2] = a[0]*b[1] - b[0]*a[1];
}
}

//! Dirichlet masks and values of the velocity and of the pressure, no-slip nodes
void RieCG::choSetupBC()
{
  auto facenodes = [&]( int s, std::set< std::size_t >& out ) {
    auto k = m_bface.find( s );
    if (k != m_bface.end()) for (auto f : k->second) for (std::size_t j=0; j<3; ++j) out.insert( m_triinpoel[f*3+j] );
  };
  auto dirbc = [&]( const std::vector< std::vector< int > >& cfgmask, const std::vector< std::vector< real > >& cfgval,
                    std::size_t ncomp, std::vector< std::size_t >& mask, std::vector< real >& val ) {
    std::map< int, std::vector< real > > dirval;
    for (const auto& s : cfgval) if (!s.empty()) dirval[ static_cast< int >( s[0] ) ].assign( s.begin()+1, s.end() );
    std::map< std::size_t, std::pair< std::vector< int >, std::vector< real > > > dirbcset;
    for (const auto& vec : cfgmask) {
      if (vec.size() != ncomp+1) throw std::runtime_error( "Incorrect Dirichlet BC mask ncomp" );
      std::set< std::size_t > nodes;
      facenodes( vec[0], nodes );
      std::vector< real > v( ncomp, 0.0 );
      auto m = dirval.find( vec[0] );
      if (m != dirval.end()) { if (m->second.size() != ncomp) throw std::runtime_error( "Incorrect Dirichlet BC val ncomp" ); v = m->second; }
      for (auto p : nodes) {
        auto& mv = dirbcset[p];
        mv.second = v;
        if (mv.first.empty()) mv.first.resize( ncomp, 0 );
        for (std::size_t c=0; c<ncomp; ++c) if (!mv.first[c]) mv.first[c] = vec[c+1];
      }
    }
    mask.clear(); val.clear();
    for (const auto& [p,mv] : dirbcset) {
      mask.push_back( p ); for (auto m : mv.first) mask.push_back( static_cast< std::size_t >( m ) );
      val.push_back( static_cast< real >( p ) ); val.insert( val.end(), mv.second.begin(), mv.second.end() );
    }
  };
  dirbc( m_cfg.bc_dir, m_cfg.bc_dirval, m_cfg.ncomp, m_dirbcmasks, m_dirbcval );
  dirbc( m_cfg.p_bc_dir, m_cfg.p_bc_dirval, 1, m_dirbcmaskp, m_dirbcvalp );
  std::set< std::size_t > ns;
  for (auto s : m_cfg.bc_noslip) facenodes( s, ns );
  m_noslipbcnodes.assign( ns.begin(), ns.end() );
}

//! Pressure Poisson matrix: tk::CSR structure from the points surrounding points (full rows,
//! ascending columns, 1-based), A(a,b) -= grad N_a . grad N_b / (6 J) per tetrahedron
void RieCG::choPrelhs()
{
  const auto& inpoel = m_disc.Inpoel();
  auto np = m_disc.Gid().size();
  auto edges = uniqueEdges( inpoel, np );
  std::vector< std::size_t > off; std::vector< std::uint32_t > nbr;
  psupFromEdges( edges, np, off, nbr );
  auto& ia = m_plhs_ia; auto& ja = m_plhs_ja; auto& a = m_plhs_a;
  ia.assign( np+1, 1 );
  for (std::size_t i=0; i<np; ++i) ia[i+1] = ia[i] + 1 + (off[i+1] - off[i]);
  ja.resize( ia[np]-1 ); a.assign( ja.size(), 0.0 );
  #pragma omp parallel for schedule(static)
  for (std::size_t i=0; i<np; ++i) {
    auto j = ia[i]-1;
    bool self = false;
    for (auto k=off[i]; k<off[i+1]; ++k) {
      if (!self && nbr[k] > i) { ja[j++] = i+1; self = true; }
      ja[j++] = static_cast< std::size_t >( nbr[k] ) + 1;
    }
    if (!self) ja[j++] = i+1;
  }
  // Row-parallel assembly: every row visits its surrounding tetrahedra in ascending element order,
  // i.e. each matrix entry receives its contributions in the order of the reference's element loop
  // (bitwise the same sums), without write conflicts between rows.
  const auto ntet = inpoel.size()/4;
  std::vector< std::size_t > eoff( np+1, 0 );
  for (std::size_t i=0; i<inpoel.size(); ++i) ++eoff[ inpoel[i]+1 ];
  for (std::size_t i=0; i<np; ++i) eoff[i+1] += eoff[i];
  std::vector< std::uint64_t > esup( inpoel.size() );           // tet*4 + local node
  { std::vector< std::size_t > fill( eoff.begin(), eoff.end()-1 );
    for (std::size_t e=0; e<ntet; ++e) for (std::size_t k=0; k<4; ++k) esup[ fill[ inpoel[e*4+k] ]++ ] = e*4+k; }
  const auto& X = m_disc.Coord()[0]; const auto& Y = m_disc.Coord()[1]; const auto& Z = m_disc.Coord()[2];
  #pragma omp parallel for schedule(dynamic,1024)
  for (std::size_t row=0; row<np; ++row) {
    auto rb = ja.begin() + static_cast< std::ptrdiff_t >( ia[row]-1 ), re = ja.begin() + static_cast< std::ptrdiff_t >( ia[row+1]-1 );
    for (auto i=eoff[row]; i<eoff[row+1]; ++i) {
      const auto e = esup[i] >> 2; const auto p = static_cast< std::size_t >( esup[i] & 3 );
      const auto N = inpoel.data() + e*4;
      real ba[3] = { X[N[1]]-X[N[0]], Y[N[1]]-Y[N[0]], Z[N[1]]-Z[N[0]] },
           ca[3] = { X[N[2]]-X[N[0]], Y[N[2]]-Y[N[0]], Z[N[2]]-Z[N[0]] },
           da[3] = { X[N[3]]-X[N[0]], Y[N[3]]-Y[N[0]], Z[N[3]]-Z[N[0]] };
      real grad[4][3];
      cross( ca, da, grad[1] ); cross( da, ba, grad[2] ); cross( ba, ca, grad[3] );
      const auto J = (ba[0]*grad[1][0] + ba[1]*grad[1][1] + ba[2]*grad[1][2]) * 6.0;
      for (std::size_t k=0; k<3; ++k) grad[0][k] = -grad[1][k]-grad[2][k]-grad[3][k];
      for (std::size_t q=0; q<4; ++q) {
        auto it = std::lower_bound( rb, re, N[q]+1 );
        a[ static_cast< std::size_t >( it - ja.begin() ) ] -=
          (grad[p][0]*grad[q][0] + grad[p][1]*grad[q][1] + grad[p][2]*grad[q][2]) / J;
      }
    }
  }
}

//! Pressure BCs and problem functions of ChoCG::pinit :1047-1122 (= LohCG::pinit :1140-1215):
//! Dirichlet node -> value, Neumann vector, rhs override
void RieCG::choPressureSetup()
{
  auto np = m_disc.Gid().size();
  const auto& co = m_disc.Coord();
  const auto& x = co[0]; const auto& y = co[1]; const auto& z = co[2];
  auto pic = problems::PRESSURE_IC( m_cfg );
  m_pbc.clear();
  for (std::size_t i=0; i<m_dirbcmaskp.size()/2; ++i) {
    auto p = m_dirbcmaskp[i*2]; auto mask = m_dirbcmaskp[i*2+1];
    if (mask == 1) m_pbc[p] = pic( x[p], y[p], z[p] );
    else if (mask == 2 && !m_dirbcvalp.empty()) m_pbc[p] = m_dirbcvalp[i*2+1];
  }
  if (m_cfg.p_hydrostat != ~0ULL) {
    const auto& gid = m_disc.Gid();
    for (std::size_t p=0; p<np; ++p)
      if (gid[p] == m_cfg.p_hydrostat) { if (!m_pbc.count( p )) m_pbc[p] = pic( x[p], y[p], z[p] ); break; }
  }
  // several partitions: a shared node may lie on a pressure-BC face of only some of its partitions; every
  // sharer must treat its row as a Dirichlet row (ConjugateGradients::init :391-417 sends the BCs of the
  // shared nodes to the fellow chares, apply :446-449 merges them)
  if (m_nranks > 1) {
    auto shared = m_disc.sharedNodes();
    std::vector< real > vals( shared.size()*2, 0.0 );
    for (std::size_t i=0; i<shared.size(); ++i) {
      auto k = m_pbc.find( shared[i] );
      if (k != m_pbc.end()) { vals[i*2] = 1.0; vals[i*2+1] = k->second; }
    }
    m_halosum( 2, vals );
    for (std::size_t i=0; i<shared.size(); ++i)
      if (vals[i*2] > 0.5 && !m_pbc.count( shared[i] )) m_pbc[ shared[i] ] = vals[i*2+1] / vals[i*2];
  }
  m_neubc.clear();
  if (auto pg = problems::PRESSURE_GRAD( m_cfg )) {
    std::vector< std::uint8_t > besym( m_triinpoel.size()/3, 0 );
    for (auto s : m_cfg.p_bc_sym) { auto k = m_bface.find( s ); if (k != m_bface.end()) for (auto f : k->second) besym[f] = 1; }
    m_neubc.assign( np, 0.0 );
    for (std::size_t e=0; e<m_triinpoel.size()/3; ++e)
      if (besym[e]) {
        const auto N = m_triinpoel.data() + e*3;
        real a[3] = { x[N[1]]-x[N[0]], y[N[1]]-y[N[0]], z[N[1]]-z[N[0]] },
             b[3] = { x[N[2]]-x[N[0]], y[N[2]]-y[N[0]], z[N[2]]-z[N[0]] };
        real n[3] = { (a[1]*b[2] - a[2]*b[1]) / 6.0, (a[2]*b[0] - a[0]*b[2]) / 6.0, (a[0]*b[1] - a[1]*b[0]) / 6.0 };
        for (std::size_t k=0; k<3; ++k) { auto g = pg( x[N[k]], y[N[k]], z[N[k]] ); m_neubc[ N[k] ] -= n[0]*g[0] + n[1]*g[1] + n[2]*g[2]; }
      }
  }
  m_prhs.clear();
  if (auto pr = problems::PRESSURE_RHS( m_cfg )) {
    m_prhs.resize( np );
    const auto& vol = m_disc.Vol();
    for (std::size_t i=0; i<np; ++i) m_prhs[i] = pr( x[i], y[i], z[i] ) * vol[i];
  }
}

//! Dirichlet rows of the momentum solve (ChoCG::solve :1580-1599): masked Dirichlet components and
//! no-slip nodes, value 0; scalar row = node*3 + component
void RieCG::choMomRows()
{
  std::set< std::size_t > rows;
  const auto nm = m_cfg.ncomp + 1;
  for (std::size_t i=0; i<m_dirbcmasks.size()/nm; ++i)
    for (std::size_t c=0; c<3; ++c) if (m_dirbcmasks[i*nm+1+c]) rows.insert( m_dirbcmasks[i*nm]*3+c );
  for (auto p : m_noslipbcnodes) for (std::size_t c=0; c<3; ++c) rows.insert( p*3+c );
  if (m_nranks > 1) {                     // union over the partitions sharing a node, as for the pressure rows
    auto shared = m_disc.sharedNodes();
    std::vector< real > vals( shared.size()*3, 0.0 );
    for (std::size_t i=0; i<shared.size(); ++i)
      for (std::size_t c=0; c<3; ++c) if (rows.count( shared[i]*3+c )) vals[i*3+c] = 1.0;
    m_halosum( 3, vals );
    for (std::size_t i=0; i<shared.size(); ++i)
      for (std::size_t c=0; c<3; ++c) if (vals[i*3+c] > 0.5) rows.insert( shared[i]*3+c );
  }
  m_mbcrows.assign( rows.begin(), rows.end() );
}

//! Momentum matrix of the semi-implicit solve, ChoCG::lhs :1433-1477: per tetrahedron
//! A(a,b,c) -= J/dt/120 (2 if a == b else 1) + theta mu grad N_a . grad N_b / (6J), the same for the
//! three components, in tk::CSR( ncomp, psup ) block form (momlhs :190-208): scalar row i*3+c holds
//! the entries of node row i at columns j*3+c. Row-parallel assembly in the reference's element order
//! (bitwise the same sums); new values go to the device's momentum solver every step (dt changes).
void RieCG::choLhs()
{
  const auto& inpoel = m_disc.Inpoel();
  auto np = m_disc.Gid().size();
  const auto& ia = m_plhs_ia; const auto& ja = m_plhs_ja;       // scalar structure = the pressure matrix's
  if (m_eoff.empty()) {
    const auto ntet = inpoel.size()/4;
    m_eoff.assign( np+1, 0 );
    for (std::size_t i=0; i<inpoel.size(); ++i) ++m_eoff[ inpoel[i]+1 ];
    for (std::size_t i=0; i<np; ++i) m_eoff[i+1] += m_eoff[i];
    m_esup.resize( inpoel.size() );
    std::vector< std::size_t > fill( m_eoff.begin(), m_eoff.end()-1 );
    for (std::size_t e=0; e<ntet; ++e) for (std::size_t k=0; k<4; ++k) m_esup[ fill[ inpoel[e*4+k] ]++ ] = e*4+k;
    m_mlhs_ia.assign( np*3+1, 1 );
    for (std::size_t i=0; i<np; ++i) for (std::size_t c=0; c<3; ++c) m_mlhs_ia[i*3+c+1] = m_mlhs_ia[i*3+c] + (ia[i+1]-ia[i]);
    choMomRows();
  }
  const auto& X = m_disc.Coord()[0]; const auto& Y = m_disc.Coord()[1]; const auto& Z = m_disc.Coord()[2];
  const auto dt = m_disc.Dt(); const auto theta = m_cfg.theta, mu = m_cfg.mu;
  auto& a = m_mlhs_a;
  a.assign( m_mlhs_ia[np*3]-1, 0.0 );
  #pragma omp parallel for schedule(dynamic,1024)
  for (std::size_t row=0; row<np; ++row) {
    auto rb = ja.begin() + static_cast< std::ptrdiff_t >( ia[row]-1 ), re = ja.begin() + static_cast< std::ptrdiff_t >( ia[row+1]-1 );
    const auto rnz = ia[row+1]-ia[row];
    auto out = a.data() + (m_mlhs_ia[row*3]-1);                   // component 0 of this node row
    for (auto i=m_eoff[row]; i<m_eoff[row+1]; ++i) {
      const auto e = m_esup[i] >> 2; const auto p = static_cast< std::size_t >( m_esup[i] & 3 );
      const auto N = inpoel.data() + e*4;
      real ba[3] = { X[N[1]]-X[N[0]], Y[N[1]]-Y[N[0]], Z[N[1]]-Z[N[0]] },
           ca[3] = { X[N[2]]-X[N[0]], Y[N[2]]-Y[N[0]], Z[N[2]]-Z[N[0]] },
           da[3] = { X[N[3]]-X[N[0]], Y[N[3]]-Y[N[0]], Z[N[3]]-Z[N[0]] };
      real grad[4][3];
      cross( ca, da, grad[1] ); cross( da, ba, grad[2] ); cross( ba, ca, grad[3] );
      const auto J = ba[0]*grad[1][0] + ba[1]*grad[1][1] + ba[2]*grad[1][2];        // J = 6V
      for (std::size_t k=0; k<3; ++k) grad[0][k] = -grad[1][k]-grad[2][k]-grad[3][k];
      for (std::size_t q=0; q<4; ++q) {
        auto v = J/dt/120.0 * ((p == q) ? 2.0 : 1.0);
        v += theta * mu * (grad[p][0]*grad[q][0] + grad[p][1]*grad[q][1] + grad[p][2]*grad[q][2]) / J / 6.0;
        auto it = std::lower_bound( rb, re, N[q]+1 );
        out[ it - rb ] -= v;
      }
    }
    for (std::size_t c=1; c<3; ++c) std::copy( out, out + rnz, out + c*rnz );
  }
  ck( xyst_cg_select( m_ctx, 1 ) );
  if (!m_mlhsup) {
    std::vector< std::size_t > jab( a.size() );
    for (std::size_t i=0; i<np; ++i) for (std::size_t c=0; c<3; ++c) {
      auto o = m_mlhs_ia[i*3+c]-1;
      for (auto j=ia[i]-1; j<ia[i+1]-1; ++j) jab[ o + (j-(ia[i]-1)) ] = (ja[j]-1)*3 + c + 1;
    }
    ck( xyst_csr_upload( m_ctx, np*3, 3, m_mlhs_ia.data(), jab.data(), a.data() ) );
    m_mlhsup = true;
  } else
    ck( xyst_csr_update( m_ctx, m_mlhs_ia.data(), a.data() ) );
  ck( xyst_cg_select( m_ctx, 0 ) );
}

//! Values of physics::dirbc (BC.cpp:29-72) at time t for the nodes of m_dirbcmasks: mask 1 = the initial
//! condition evaluated at the BC time, mask 2 = the configured value; [node][ncomp]
void RieCG::choDirvals( real t, std::vector< real >& dv ) const
{
  const auto ncomp = m_cfg.ncomp;
  const auto& co = m_disc.Coord();
  auto nd = m_dirbcmasks.size()/(ncomp+1);
  dv.assign( nd*ncomp, 0.0 );
  auto ic = problems::IC( m_cfg );
  for (std::size_t i=0; i<nd; ++i) {
    auto p = m_dirbcmasks[i*(ncomp+1)];
    auto u = ic( co[0][p], co[1][p], co[2][p], t );
    for (std::size_t c=0; c<ncomp; ++c) {
      auto mask = m_dirbcmasks[i*(ncomp+1)+1+c];
      if (mask == 1) dv[i*ncomp+c] = u[c];
      else if (mask == 2 && !m_dirbcval.empty()) dv[i*ncomp+c] = m_dirbcval[i*(ncomp+1)+1+c];
    }
  }
}

//! Dirichlet values of the next BC application on the device (time-dependent initial conditions)
void RieCG::choBCtime( real t )
{
  if (m_dirbcmasks.empty()) return;
  std::vector< real > dv;
  choDirvals( t, dv );
  ck( xyst_chocg_dirbc_values( m_ctx, dv.data() ) );
}

//! Device upload and the start-up sequence of ChoCG::merge :816-837 onwards: make the initial
//! velocity divergence-free and compute the initial pressure
void RieCG::choSetup()
{
  if (m_cfg.rk < 1 || m_cfg.rk > 4) throw std::runtime_error( "ChoCG: rk must be 1..4" );
  auto np = m_disc.Gid().size();
  const auto& co = m_disc.Coord();
  const auto& x = co[0]; const auto& y = co[1]; const auto& z = co[2];
  std::size_t nsup[3] = { m_dsupedge[0].size()/4, m_dsupedge[1].size()/3, m_dsupedge[2].size()/2 };
  const std::size_t* se[3] = { m_dsupedge[0].data(), m_dsupedge[1].data(), m_dsupedge[2].data() };
  const real* si[3] = { m_dsupint[0].data(), m_dsupint[1].data(), m_dsupint[2].data() };
  xyst_chocg_params prm{};
  // an unknown flux only matters once chorin::rhs is called (Chorin.cpp:1030-1035): see choStep
  if (m_cfg.flux == "damp4") prm.flux = 1; else prm.flux = 0;
  prm.stab = m_cfg.stab; prm.stab2 = m_cfg.stab2; prm.stab2coef = m_cfg.stab2coef; prm.mu = m_cfg.mu;
  ck( xyst_chocg_mesh_upload( m_ctx, np, x.data(), y.data(), z.data(), nsup, se, si, m_triinpoel.size()/3,
                              m_triinpoel.data(), m_disc.Vol().data(), m_disc.V().data(), &prm ) );
  const auto ncomp = m_cfg.ncomp;
  if (ncomp < 3 || ncomp > 7) throw std::runtime_error( "ChoCG: three velocity components + up to four transported scalars" );
  if (ncomp > 3) {
    if (m_cfg.theta > std::numeric_limits< real >::epsilon())
      throw std::runtime_error( "ChoCG: the semi-implicit momentum solve with transported scalars is not implemented" );
    ck( xyst_chocg_scalars( m_ctx, static_cast< int >( ncomp-3 ), m_cfg.dif ) );
  }
  // physics::dirbc (BC.cpp:29-72): mask 1 = value of the initial condition, 2 = configured value
  auto nd = m_dirbcmasks.size()/(ncomp+1);
  std::vector< std::size_t > dn( nd ); std::vector< int > dm( nd*ncomp ); std::vector< real > dv;
  for (std::size_t i=0; i<nd; ++i) {
    dn[i] = m_dirbcmasks[i*(ncomp+1)];
    for (std::size_t c=0; c<ncomp; ++c) {
      auto mask = static_cast< int >( m_dirbcmasks[i*(ncomp+1)+1+c] );
      if (mask == 2 && m_dirbcval.empty()) mask = 0;
      if (mask != 1 && mask != 2) mask = 0;
      dm[i*ncomp+c] = mask;
    }
  }
  choDirvals( m_disc.T(), dv );
  ck( xyst_chocg_bc_upload( m_ctx, nd, dn.data(), dm.data(), dv.data(), m_symbcnodes.size(), m_symbcnodes.data(),
                            m_symbcnorms.data(), m_noslipbcnodes.size(), m_noslipbcnodes.data() ) );
  ck( xyst_csr_upload( m_ctx, np, 1, m_plhs_ia.data(), m_plhs_ja.data(), m_plhs_a.data() ) );
  choPressureSetup();
  m_psol.clear();
  if (auto ps = problems::PRESSURE_SOL( m_cfg )) { m_psol.resize( np ); for (std::size_t i=0; i<np; ++i) m_psol[i] = ps( x[i], y[i], z[i] ); }
  if (!m_src.empty()) ck( xyst_chocg_src( m_ctx, m_src.data() ) );
  ck( xyst_chocg_set_u( m_ctx, m_u0.data() ) );
  ck( xyst_chocg_apply_bc( m_ctx ) );
  m_np = 0; m_initial = true;
  ck( xyst_chocg_div( m_ctx, 0, m_disc.Dt(), m_np > 1 ) );
  choPinit(); choPsolve();
  ck( xyst_chocg_grad( m_ctx, 0 ) );
  choPsolved( nullptr );
  ck( xyst_sync( m_ctx ) );
}

void RieCG::choPinit()
{
  std::vector< std::size_t > nodes; std::vector< real > vals;
  for (const auto& [p,v] : m_pbc) { nodes.push_back( p ); vals.push_back( m_np > 1 ? 0.0 : v ); }
  ck( xyst_chocg_pinit( m_ctx, m_np > 1 ? m_disc.Dt() : 1.0, nodes.size(), nodes.data(), vals.data(),
                        m_neubc.empty() ? nullptr : m_neubc.data(), m_prhs.empty() ? nullptr : m_prhs.data(),
                        m_cfg.p_pc == "jacobi" ? 1 : 0 ) );
}

void RieCG::choPsolve()
{
  real normr = 0.0;
  ck( xyst_cg_solve( m_ctx, m_cfg.p_iter, m_cfg.p_tol, &m_pit, &normr ) );
}

void RieCG::choPsolved( std::vector< real >* diagrow )
{
  if (m_np != 1) {
    if (m_timedep) choBCtime( m_disc.T() + m_disc.Dt() );          // BC( m_u, d->T() + d->Dt() ) :1215
    ck( xyst_chocg_project( m_ctx, m_np > 1 ? m_disc.Dt() : 1.0 ) );
  }
  if (m_initial) {
    if (m_cfg.nstep == 1) {                  // test first Poisson solve only (:1223-1229)
      ck( xyst_chocg_pressure_update( m_ctx, 0 ) );
      m_lastdiag = choDiag();
      m_finished = true;
    } else if (++m_np < 2) {
      ck( xyst_chocg_vgrad( m_ctx ) );
      ck( xyst_chocg_flux( m_ctx ) );
      ck( xyst_chocg_div( m_ctx, 1, m_disc.Dt(), m_np > 1 ) );
      choPinit(); choPsolve();               // m_np == 1: no gradient of the solution needed (:1138-1140)
      choPsolved( diagrow );
    } else {
      ck( xyst_chocg_pressure_update( m_ctx, 0 ) );
      ck( xyst_chocg_grad( m_ctx, 1 ) );
      m_initial = false;
    }
  } else {
    ck( xyst_chocg_pressure_update( m_ctx, 1 ) );
    ck( xyst_chocg_grad( m_ctx, 1 ) );
    auto row = choDiag();
    if (diagrow) *diagrow = row;
  }
}

bool RieCG::choStep( std::vector< real >* diagrow )
{
  if (diagrow) diagrow->clear();
  if (m_finished) {                          // the nstep = 1 run finished during setup: hand out its row
    if (diagrow && !m_lastdiag.empty()) { *diagrow = m_lastdiag; m_lastdiag.clear(); }
    return false;
  }
  if (m_cfg.flux != "damp2" && m_cfg.flux != "damp4") throw std::runtime_error( "Flux not correctly configured" );
  auto eps = std::numeric_limits< real >::epsilon();
  real mindt;
  if (std::abs( m_cfg.dt ) > eps) mindt = m_cfg.dt;
  else {
    ck( xyst_chocg_dt_min( m_ctx, m_cfg.cfl, m_cfg.dif, &mindt ) );
    if (m_disc.T() > m_cfg.freezetime) m_freezeflow = m_cfg.freezeflow;        // :1396-1399
    mindt *= m_freezeflow;
    if (m_nranks > 1) { std::vector< real > t{ mindt }; m_allreduce( 1, t ); mindt = t[0]; }   // contribute(min_double) :1407-1410
  }
  if (mindt < eps) m_finished = true;
  m_disc.setdt( mindt );
  const bool implicit = m_cfg.theta > eps;
  // frozen flow (solve :1550-1552,1564-1570): the velocity of before the update comes back once pred() has
  // returned -- after the stage's BCs and velocity gradient and, at the last stage of a serial run, after div()
  const bool frozen = m_freezeflow > 1.0;
  if (frozen && implicit) throw std::runtime_error( "ChoCG: freezeflow with the semi-implicit momentum solve is not implemented" );
  if (implicit) choLhs();                    // advance :1414-1431
  // problems::point_src (ChoCG::pred :1655-1657): active for all stages of a step that starts at or after
  // the release time
  if (m_cfg.problem == "point_src" && m_cfg.ncomp > 3 && m_cfg.src_radius >= 0.0 && !m_pinned &&
      !(m_disc.T() < m_cfg.src_release_time)) {
    const auto& co = m_disc.Coord();
    std::vector< std::size_t > nodes;
    for (std::size_t i=0; i<co[0].size(); ++i) {
      auto rx = m_cfg.src_location[0] - co[0][i], ry = m_cfg.src_location[1] - co[1][i], rz = m_cfg.src_location[2] - co[2][i];
      if (rx*rx + ry*ry + rz*rz < m_cfg.src_radius*m_cfg.src_radius) nodes.push_back( i );
    }
    ck( xyst_chocg_pin( m_ctx, nodes.size(), nodes.data(), 1.0 ) );
    m_pinned = true;
  }
  if (m_timedep) {                           // chorin::rhs( ..., d->T(), ... ) :1489: the source at the step's time level
    evalSrc( m_disc.T() );
    if (!m_src.empty()) ck( xyst_chocg_src( m_ctx, m_src.data() ) );
  }
  for (std::uint64_t s=0; s<m_cfg.rk; ++s)
    if (!implicit || s+1 < m_cfg.rk) {       // solve :1555-1572
      if (m_timedep) choBCtime( m_disc.T() + rkcoef[m_cfg.rk-1][s] * m_disc.Dt() );    // pred :1660
      ck( xyst_chocg_stage( m_ctx, static_cast< int >( s ), rkcoef[m_cfg.rk-1][s], m_disc.Dt() ) );
      if (frozen && s+1 < m_cfg.rk) ck( xyst_chocg_restore_velocity( m_ctx ) );
    } else {                                   // semi-implicit momentum solve at the last stage, :1574-1645
      ck( xyst_chocg_rhs( m_ctx ) );
      ck( xyst_cg_select( m_ctx, 1 ) );
      ck( xyst_chocg_minit( m_ctx, m_mbcrows.size(), m_mbcrows.data(), m_cfg.mom_pc == "jacobi" ? 1 : 0 ) );
      real normr = 0.0;
      ck( xyst_cg_solve( m_ctx, m_cfg.mom_iter, m_cfg.mom_tol, &m_mit, &normr ) );
      ck( xyst_chocg_mupdate( m_ctx, static_cast< int >( s ) ) );
      ck( xyst_cg_select( m_ctx, 0 ) );
    }
  ck( xyst_chocg_div( m_ctx, 0, m_disc.Dt(), m_np > 1 ) );
  if (frozen) ck( xyst_chocg_restore_velocity( m_ctx ) );
  choPinit(); choPsolve();
  ck( xyst_chocg_grad( m_ctx, 0 ) );
  choPsolved( diagrow );
  if (m_disc.finished()) m_finished = true;
  return !m_finished;
}

//! ChoCG::diag :1697-1714 (next() first) + NodeDiagnostics::precompute :147-268 +
//! Transporter::prediagnostics :1528-1608. Returns an empty row on non-diagnostics steps.
std::vector< real > RieCG::choDiag()
{
  m_disc.next();
  if ((m_disc.It()+1) % m_cfg.diag_iter) return {};
  const auto& co = m_disc.Coord();
  auto np = co[0].size();
  const auto nc = m_cfg.ncomp;
  std::vector< real > anu;
  bool psol = !m_psol.empty();
  auto sol = problems::SOL( m_cfg );
  if (sol && !psol) {
    anu.resize( np*nc );
    for (std::size_t i=0; i<np; ++i) { auto s = sol( co[0][i], co[1][i], co[2][i], m_disc.T()+m_disc.Dt() );
      for (std::size_t c=0; c<nc; ++c) anu[i*nc+c] = s[c]; }
  }
  real d[32];
  ck( xyst_chocg_diag( m_ctx, psol ? m_psol.data() : nullptr, anu.empty() ? nullptr : anu.data(), d ) );
  if (m_nranks > 1) for (std::size_t o=0; o<16+4*(nc-3); o+=16) {
    std::vector< real > t( d+o, d+std::min< std::size_t >( o+16, 16+4*(nc-3) ) ); m_allreduce( 0, t ); std::copy( t.begin(), t.end(), d+o ); }
  // sums of component c (0 = pressure, 1..3 velocity, 4.. scalars): [0] L2 solution [1] L2 increment [2] L2 error [3] L1 error
  auto sum = [&]( std::size_t k, std::size_t c ) -> real {
    if (c > 3) return d[16 + 4*(c-4) + k];
    return k == 0 ? d[c] : k == 1 ? d[4+c] : k == 2 ? d[9+c] : d[12+c]; };
  std::size_t ncomp = psol ? 0 : nc;
  auto mv = m_disc.MeshVol();
  std::vector< real > row{ static_cast< real >( m_disc.It() ), m_disc.T(), m_disc.Dt() };
  for (std::size_t i=0; i<=ncomp; ++i) row.push_back( std::sqrt( sum( 0, i ) / mv ) );
  for (std::size_t i=0; i<=ncomp; ++i) row.push_back( std::sqrt( sum( 1, i ) / mv ) );
  if (psol) { row.push_back( std::sqrt( d[8] / mv ) ); row.push_back( d[9] / mv ); }
  if (!anu.empty()) {
    for (std::size_t i=1; i<=ncomp; ++i) row.push_back( std::sqrt( sum( 2, i ) / mv ) );
    for (std::size_t i=1; i<=ncomp; ++i) row.push_back( sum( 3, i ) / mv );
  }
  return row;
}

std::vector< real > RieCG::choGet( const char* what, std::size_t width )
{
  std::vector< real > r( m_disc.Gid().size()*width );
  ck( xyst_chocg_get( m_ctx, what, r.data() ) );
  return r;
}

} // xyst::

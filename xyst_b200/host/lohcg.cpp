// xyst_b200/host/lohcg.cpp -- the LohCG members of the host mirror (solver = "lohcg"):
// artificial-compressibility solver for constant-density flow, src/Inciter/LohCG.cpp + lohcg.ci,
// on one or several partitions. Unknowns (p,u,v,w). Setup shares ChoCG's pieces (Dirichlet masks with values,
// pressure BCs, no-slip nodes, the pressure Poisson matrix: LohCG::setupDirBC :215-296 and
// prelhs :140-181 are the ChoCG members of the same names); the start-up makes the initial
// velocity divergence-free with two pressure solves (LohCG::merge :909-931 onwards) and the time
// step is explicit Runge-Kutta on all four unknowns. Every nodal and edge loop runs on the device.
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
                            { 1.0/4.0, 1.0/3.0, 1.0/2.0, 1.0 } };           // LohCG.cpp:42-47
}

void RieCG::lohSetup()
{
  if (m_cfg.rk < 1 || m_cfg.rk > 4) throw std::runtime_error( "LohCG: rk must be 1..4" );
  auto np = m_disc.Gid().size();
  const auto& co = m_disc.Coord();
  const auto& x = co[0]; const auto& y = co[1]; const auto& z = co[2];
  std::size_t nsup[3] = { m_dsupedge[0].size()/4, m_dsupedge[1].size()/3, m_dsupedge[2].size()/2 };
  const std::size_t* se[3] = { m_dsupedge[0].data(), m_dsupedge[1].data(), m_dsupedge[2].data() };
  const real* si[3] = { m_dsupint[0].data(), m_dsupint[1].data(), m_dsupint[2].data() };
  xyst_lohcg_params prm{};
  // an unknown flux only matters once lohner::rhs is called (Lohner.cpp:1116-1121): see lohStep
  prm.flux = m_cfg.flux == "damp4" ? 1 : 0;
  prm.stab = m_cfg.stab; prm.stab2 = m_cfg.stab2; prm.stab2coef = m_cfg.stab2coef; prm.mu = m_cfg.mu;
  prm.soundspeed = m_cfg.soundspeed;
  ck( xyst_lohcg_mesh_upload( m_ctx, np, x.data(), y.data(), z.data(), nsup, se, si, m_triinpoel.size()/3,
                              m_triinpoel.data(), m_disc.Vol().data(), m_disc.V().data(), &prm ) );
  const auto ncomp = m_cfg.ncomp;
  if (ncomp > 4) ck( xyst_lohcg_scalars( m_ctx, static_cast< int >( ncomp-4 ), m_cfg.dif ) );
  // physics::dirbc (BC.cpp:29-72): mask 1 = value of the initial condition at the BC time (LohCG::merge :923:
  // t + dt), 2 = configured value
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
  choDirvals( m_disc.T() + m_disc.Dt(), dv );
  // physics::dirbcp (BC.cpp:74-108)
  std::vector< std::size_t > pn; std::vector< real > pv;
  auto pic = problems::PRESSURE_IC( m_cfg );
  for (std::size_t i=0; i<m_dirbcmaskp.size()/2; ++i) {
    auto p = m_dirbcmaskp[i*2]; auto mask = m_dirbcmaskp[i*2+1];
    if (mask == 1) { pn.push_back( p ); pv.push_back( pic( x[p], y[p], z[p] ) ); }
    else if (mask == 2 && !m_dirbcvalp.empty()) { pn.push_back( p ); pv.push_back( m_dirbcvalp[i*2+1] ); }
  }
  ck( xyst_lohcg_bc_upload( m_ctx, nd, dn.data(), dm.data(), dv.data(), pn.size(), pn.data(), pv.data(),
                            m_symbcnodes.size(), m_symbcnodes.data(), m_symbcnorms.data(),
                            m_noslipbcnodes.size(), m_noslipbcnodes.data() ) );
  ck( xyst_csr_upload( m_ctx, np, 1, m_plhs_ia.data(), m_plhs_ja.data(), m_plhs_a.data() ) );
  choPressureSetup();
  if (!m_src.empty()) ck( xyst_lohcg_src( m_ctx, m_src.data() ) );
  ck( xyst_lohcg_set_u( m_ctx, m_u0.data() ) );
  // LohCG::merge :917-931: velocity BCs, divergence of the velocity, first pressure solve
  ck( xyst_lohcg_apply_bc( m_ctx, 0 ) );
  m_np = 0;
  ck( xyst_chocg_div( m_ctx, 0, 0.0, 0 ) );
  lohPinit();
  { real normr = 0.0; ck( xyst_cg_solve( m_ctx, m_cfg.p_iter, m_cfg.p_tol, &m_pit, &normr ) ); }
  ck( xyst_chocg_grad( m_ctx, 0 ) );
  lohPsolved();
  ck( xyst_sync( m_ctx ) );
}

void RieCG::lohPinit()
{
  std::vector< std::size_t > nodes; std::vector< real > vals;
  for (const auto& [p,v] : m_pbc) { nodes.push_back( p ); vals.push_back( v ); }
  ck( xyst_chocg_pinit( m_ctx, 1.0, nodes.size(), nodes.data(), vals.data(),
                        m_neubc.empty() ? nullptr : m_neubc.data(), m_prhs.empty() ? nullptr : m_prhs.data(),
                        m_cfg.p_pc == "jacobi" ? 1 : 0 ) );
}

void RieCG::lohPsolved()
{
  if (m_np != 1) {
    if (m_timedep) choBCtime( m_disc.T() + m_disc.Dt() );          // LohCG::psolved :1311
    ck( xyst_lohcg_project( m_ctx ) );
  }
  if (m_cfg.nstep == 1) {                    // test first Poisson solve only (:1330-1337)
    ck( xyst_lohcg_pressure_set( m_ctx ) );
    m_lastdiag = lohDiag();
    m_finished = true;
  } else if (++m_np < 2) {
    // second solve: divergence of the momentum flux -> initial pressure (:1339-1352)
    ck( xyst_chocg_vgrad( m_ctx ) );
    ck( xyst_chocg_flux( m_ctx ) );
    ck( xyst_chocg_div( m_ctx, 1, 0.0, 0 ) );
    lohPinit();
    { real normr = 0.0; ck( xyst_cg_solve( m_ctx, m_cfg.p_iter, m_cfg.p_tol, &m_pit, &normr ) ); }
    lohPsolved();                            // m_np == 1: no gradient of the solution needed (:1232-1234)
  } else
    ck( xyst_lohcg_pressure_set( m_ctx ) );
}

bool RieCG::lohStep( std::vector< real >* diagrow )
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
    ck( xyst_lohcg_dt_min( m_ctx, m_cfg.cfl, m_cfg.dif, &mindt ) );
    if (m_nranks > 1) { std::vector< real > t{ mindt }; m_allreduce( 1, t ); mindt = t[0]; }   // contribute(min_double) :1445-1446
  }
  if (mindt < eps) m_finished = true;
  m_disc.setdt( mindt );
  // problems::point_src (LohCG::solve :1615-1617): active for all stages of a step that starts at or after the
  // release time
  if (m_cfg.problem == "point_src" && m_cfg.ncomp > 4 && m_cfg.src_radius >= 0.0 && !m_pinned &&
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
  if (m_timedep) {                           // lohner::rhs( ..., d->T(), ... ) :1546: the source at the step's time level
    evalSrc( m_disc.T() );
    if (!m_src.empty()) ck( xyst_lohcg_src( m_ctx, m_src.data() ) );
  }
  for (std::uint64_t s=0; s<m_cfg.rk; ++s) {
    if (m_timedep) choBCtime( m_disc.T() + rkcoef[m_cfg.rk-1][s] * m_disc.Dt() );      // solve :1620, solved :1640
    ck( xyst_lohcg_stage( m_ctx, static_cast< int >( s ), rkcoef[m_cfg.rk-1][s], m_disc.Dt() ) );
  }
  auto row = lohDiag();
  if (diagrow) *diagrow = row;
  if (m_disc.finished()) m_finished = true;
  return !m_finished;
}

//! LohCG::diag :1366-1382 (next() first) + NodeDiagnostics::accompute :270-372 +
//! Transporter::acdiagnostics :1621-1700. Returns an empty row on non-diagnostics steps.
std::vector< real > RieCG::lohDiag()
{
  m_disc.next();
  if ((m_disc.It()+1) % m_cfg.diag_iter) return {};
  const auto& co = m_disc.Coord();
  auto np = co[0].size();
  const auto nc = m_cfg.ncomp;
  std::vector< real > an;
  if (auto sol = problems::SOL( m_cfg )) {
    an.resize( np*nc );
    for (std::size_t i=0; i<np; ++i) { auto s = sol( co[0][i], co[1][i], co[2][i], m_disc.T()+m_disc.Dt() );
      for (std::size_t c=0; c<nc; ++c) an[i*nc+c] = s[c]; }
  }
  real d[32];
  ck( xyst_lohcg_diag( m_ctx, an.empty() ? nullptr : an.data(), d ) );
  if (m_nranks > 1) for (std::size_t o=0; o<16+4*(nc-4); o+=16) {
    std::vector< real > t( d+o, d+std::min< std::size_t >( o+16, 16+4*(nc-4) ) ); m_allreduce( 0, t ); std::copy( t.begin(), t.end(), d+o ); }
  // sums of component c (0 = p, 1..3 velocity, 4.. scalars): [0] L2 solution [1] L2 increment [2] L2 error [3] L1 error
  auto sum = [&]( std::size_t k, std::size_t c ) -> real { return c > 3 ? d[16 + 4*(c-4) + k] : d[4*k + c]; };
  auto mv = m_disc.MeshVol();
  std::vector< real > row{ static_cast< real >( m_disc.It() ), m_disc.T(), m_disc.Dt() };
  for (std::size_t i=0; i<nc; ++i) row.push_back( std::sqrt( sum( 0, i ) / mv ) );
  for (std::size_t i=0; i<nc; ++i) row.push_back( std::sqrt( sum( 1, i ) / mv ) );
  if (!an.empty()) {
    for (std::size_t i=1; i<nc; ++i) row.push_back( std::sqrt( sum( 2, i ) / mv ) );
    for (std::size_t i=1; i<nc; ++i) row.push_back( sum( 3, i ) / mv );
  }
  return row;
}

} // xyst::

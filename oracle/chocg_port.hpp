// oracle/chocg_port.hpp -- TEST INFRASTRUCTURE ONLY (never linked by the product).
//
// Serial restatement of the ChoCG solver chare of the reference (projection method for
// constant-density flow, src/Inciter/ChoCG.cpp) on top of the shared setup pipeline of
// driver.hpp, the Chorin edge operators (physics_port.hpp or, with -DORACLE_REF, the
// reference's own Chorin.cpp) and the conjugate gradients restatement (cg_port.hpp).
// Explicit (theta = 0) or semi-implicit (theta > 0: consistent-mass + theta-weighted viscous matrix,
// block CG on the three velocity components at the last RK stage) momentum update, velocity
// components only (ncomp = 3).
// Each member cites the ChoCG.cpp lines it follows; the SDAG control flow
// (src/Inciter/chocg.ci) is unrolled into plain calls, chares are visited in index order.
#pragma once
#include "driver.hpp"
#include "cg_port.hpp"

namespace orc {

class ChoRun : public Run {
  public:
    cg::Solver cgpre, cgmom;              // pressure solve; momentum solve (theta > 0, ChoCG.cpp:134-140)
    std::size_t mit = 0;                  // iterations of the last momentum solve
    int np = 0;                           // ChoCG::m_np
    bool initial = true;                  // Discretization::Initial()
    std::vector< real > rk;               // m_rkcoef, ChoCG.cpp:43-48
    std::size_t pit = 0;                  // iterations of the last pressure solve

    ChoRun( const MeshInput& in, const Cfg& c, const std::vector< std::size_t >& target, int nchare )
      : Run( in, c, target, nchare )
    {
      if (cfg.ncomp < 3) throw std::runtime_error( "oracle ChoCG: three velocity components (+ transported scalars)" );
      static const std::vector< std::vector< real > > rkcoef{ { 1.0 }, { 1.0/2.0, 1.0 }, { 1.0/3.0, 1.0/2.0, 1.0 },
                                                               { 1.0/4.0, 1.0/3.0, 1.0/2.0, 1.0 } };
      rk = rkcoef.at( cfg.rk - 1 );
      // ChoCG ctor :126-131 + prelhs :146-188: pressure Laplacian on the renumbered mesh
      for (auto& cp : ch) {
        auto& c_ = *cp;
        auto psup = be::genPsup( c_.inpoel, 4, be::genEsup( c_.inpoel, 4 ) );
        cg::CommMap cm;
        for (const auto& [k,n] : c_.nodeCommMap) cm[k] = n;
        auto k = cgpre.add( 1, psup, c_.gid, cm );
        if (cfg.theta > std::numeric_limits< real >::epsilon()) cgmom.add( cfg.ncomp, psup, c_.gid, cm );   // momlhs :190-208
        auto& A = cgpre.parts[k]->A;
        const auto& X = c_.coord[0]; const auto& Y = c_.coord[1]; const auto& Z = c_.coord[2];
        for (std::size_t e=0; e<c_.inpoel.size()/4; ++e) {
          const auto N = c_.inpoel.data() + e*4;
          real ba[3] = { X[N[1]]-X[N[0]], Y[N[1]]-Y[N[0]], Z[N[1]]-Z[N[0]] },
               ca[3] = { X[N[2]]-X[N[0]], Y[N[2]]-Y[N[0]], Z[N[2]]-Z[N[0]] },
               da[3] = { X[N[3]]-X[N[0]], Y[N[3]]-Y[N[0]], Z[N[3]]-Z[N[0]] };
          auto cross = []( const real a[3], const real b[3], real r[3] ){
            r[0] = a[1]*b[2] - b[1]*a[2]; r[1] = a[2]*b[0] - b[2]*a[0]; r[2] = a[0]*b[1] - b[0]*a[1]; };
          real grad[4][3];
          cross( ca, da, grad[1] ); cross( da, ba, grad[2] ); cross( ba, ca, grad[3] );
          const auto J = (ba[0]*grad[1][0] + ba[1]*grad[1][1] + ba[2]*grad[1][2]) * 6.0;
          for (std::size_t i=0; i<3; ++i) grad[0][i] = -grad[1][i]-grad[2][i]-grad[3][i];
          for (std::size_t a=0; a<4; ++a)
            for (std::size_t b=0; b<4; ++b)
              A( N[a], N[b] ) -= (grad[a][0]*grad[b][0] + grad[a][1]*grad[b][1] + grad[a][2]*grad[b][2]) / J;
        }
      }
      for (auto& cp : ch) u0.push_back( cp->u );
      // ChoCG::merge :816-837 onwards: make the initial velocity divergence free, initial pressure
      div_u();
      pinit(); psolve();
      sgrad(); psolved();
    }

    std::vector< be::Fields > u0;         // initial velocity, before the first projection

    //! nodal values of the problem functions and the assembled pressure matrix of one chare, for
    //! the parity tests of the device path (the host solver class evaluates these itself)
    std::vector< real > exported( std::size_t k, const std::string& n ) {
      auto& c_ = *ch.at( k );
      const auto& x = c_.coord[0]; const auto& y = c_.coord[1]; const auto& z = c_.coord[2];
      std::vector< real > r;
      if (n == "u0") { const auto& f = u0.at( k ); for (std::size_t i=0; i<f.nunk(); ++i) for (std::size_t c=0; c<f.nprop(); ++c) r.push_back( f(i,c) ); }
      else if (n == "p_ic") { auto ic = be::PRESSURE_IC(); for (std::size_t i=0; i<x.size(); ++i) r.push_back( ic( x[i], y[i], z[i] ) ); }
      else if (n == "p_sol") { if (auto f = be::PRESSURE_SOL()) for (std::size_t i=0; i<x.size(); ++i) r.push_back( f( x[i], y[i], z[i] ) ); }
      else if (n == "p_rhs") { if (auto f = be::PRESSURE_RHS()) for (std::size_t i=0; i<x.size(); ++i) r.push_back( f( x[i], y[i], z[i] ) * c_.vol[i] ); }
      else if (n == "u_sol") { if (auto f = be::SOL()) for (std::size_t i=0; i<x.size(); ++i) { auto s = f( x[i], y[i], z[i], t+dt ); r.insert( r.end(), s.begin(), s.end() ); } }
      else if (n == "neubc") {
        if (auto pg = be::PRESSURE_GRAD()) {
          std::vector< std::uint8_t > besym( c_.triinpoel.size(), 0 );
          for (auto s : cfg.p_bc_sym) { auto kk = c_.bface.find( s ); if (kk != c_.bface.end()) for (auto f : kk->second) besym[f] = 1; }
          r.assign( x.size(), 0.0 );
          for (std::size_t e=0; e<c_.triinpoel.size()/3; ++e)
            if (besym[e]) {
              const auto N = c_.triinpoel.data() + e*3;
              real nn[3]; port::crossdiv6( c_.coord, N, nn );
              for (std::size_t a=0; a<3; ++a) { auto g = pg( x[N[a]], y[N[a]], z[N[a]] ); r[ N[a] ] -= nn[0]*g[0] + nn[1]*g[1] + nn[2]*g[2]; }
            }
        }
      }
      else if (n == "hydrostat") { if (cfg.p_hydrostat != ~0ULL) { auto pi = c_.lid.find( cfg.p_hydrostat ); if (pi != c_.lid.end()) r.push_back( static_cast< real >( pi->second ) ); } }
      else if (n == "plhs_a") {
        auto& P = *cgpre.parts.at( k );
        const auto& ia = P.S.IA(); const auto& ja = P.S.JA();
        for (std::size_t row=0; row+1<ia.size(); ++row) for (std::size_t j=ia[row]-1; j<ia[row+1]-1; ++j) r.push_back( P.A( row, ja[j]-1 ) );
      }
      else if (n == "mlhs_a") {            // momentum matrix of the last step (restored after the solve)
        if (cgmom.parts.empty()) return r;
        auto& P = *cgmom.parts.at( k );
        const auto& ia = P.S.IA(); const auto& ja = P.S.JA();
        auto nc = P.A.Ncomp();
        for (std::size_t row=0; row+1<ia.size(); ++row) for (std::size_t j=ia[row]-1; j<ia[row+1]-1; ++j)
          r.push_back( P.A( row/nc, (ja[j]-1)/nc, row%nc ) );
      }
      else throw std::runtime_error( "oracle ChoCG: unknown export " + n );
      return r;
    }

    //! sum a nodal field over the chares sharing each node (com* entry methods)
    template< class Get > void sumShared( Get get ) {
      std::vector< std::unordered_map< std::size_t, std::vector< real > > > recv( ch.size() );
      for (std::size_t a=0; a<ch.size(); ++a) {
        auto& F = get( *ch[a] );
        for (const auto& [b,n] : ch[a]->nodeCommMap)
          for (auto g : n) {
            auto i = ch[a]->lid.at( g );
            auto& acc = recv[static_cast<std::size_t>(b)][g];
            if (acc.empty()) acc.assign( F.nprop(), 0.0 );
            for (std::size_t c=0; c<F.nprop(); ++c) acc[c] += F(i,c);
          }
      }
      for (std::size_t b=0; b<ch.size(); ++b) {
        auto& F = get( *ch[b] );
        for (const auto& [g,r] : recv[b]) { auto i = ch[b]->lid.at( g ); for (std::size_t c=0; c<r.size(); ++c) F(i,c) += r[c]; }
      }
    }
    void sumSharedScalar( std::vector< real > Chare::*v ) {
      std::vector< std::unordered_map< std::size_t, real > > recv( ch.size() );
      for (std::size_t a=0; a<ch.size(); ++a)
        for (const auto& [b,n] : ch[a]->nodeCommMap)
          for (auto g : n) recv[static_cast<std::size_t>(b)][g] += ((*ch[a]).*v)[ ch[a]->lid.at( g ) ];
      for (std::size_t b=0; b<ch.size(); ++b)
        for (const auto& [g,r] : recv[b]) ((*ch[b]).*v)[ ch[b]->lid.at( g ) ] += r;
    }
    //! ChoCG::fingrad :840-865 (the sum over chares has been done by sumShared)
    static void fingrad( Chare& c_, be::Fields& grad ) {
      for (std::size_t p=0; p<grad.nunk(); ++p) for (std::size_t c=0; c<grad.nprop(); ++c) grad(p,c) /= c_.vol[p];
    }

    //! ChoCG::div :868-900 of the velocity
    void div_u() {
      for (auto& cp : ch) { auto& c_ = *cp;
        std::fill( c_.div.begin(), c_.div.end(), 0.0 );
        be::chorin_div( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, dt, c_.pr, c_.pgrad, c_.u, c_.div, np > 1 ); }
      sumSharedScalar( &Chare::div );
    }
    //! ChoCG::div :868-900 of the momentum flux (m_np == 1)
    void div_flux() {
      for (auto& cp : ch) { auto& c_ = *cp;
        fingrad( c_, c_.mflux );
        be::symbc( c_.mflux, c_.symbcnodes, c_.symbcnorms, 0 );
        std::fill( c_.div.begin(), c_.div.end(), 0.0 );
        be::chorin_div( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, dt, c_.pr, c_.pgrad, c_.mflux, c_.div, np > 1 ); }
      sumSharedScalar( &Chare::div );
    }
    //! ChoCG::velgrad :923-946
    void velgrad() {
      for (auto& cp : ch) { auto& c_ = *cp; c_.grad.fill( 0.0 );
        be::chorin_vgrad( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.u, c_.grad ); }
      sumShared( []( Chare& c_ ) -> be::Fields& { return c_.grad; } );
    }
    //! ChoCG::flux :972-1000
    void flux() {
      for (auto& cp : ch) { auto& c_ = *cp;
        fingrad( c_, c_.grad );
        c_.mflux.fill( 0.0 );
        be::chorin_flux( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.u, c_.grad, c_.mflux ); }
      sumShared( []( Chare& c_ ) -> be::Fields& { return c_.mflux; } );
    }

    //! ChoCG::pinit :1025-1125
    void pinit() {
      std::vector< std::vector< real > > b( ch.size() );
      std::vector< cg::BCs > bcs( ch.size() );
      for (std::size_t k=0; k<ch.size(); ++k) {
        auto& c_ = *ch[k];
        const auto& x = c_.coord[0]; const auto& y = c_.coord[1]; const auto& z = c_.coord[2];
        if (np > 1) for (auto& d : c_.div) d /= dt;
        auto& dirbc = bcs[k].dirbc; auto& neubc = bcs[k].neubc;
        if (np < 3) {
          if (!cfg.p_bc_dir.empty()) {
            auto ic = be::PRESSURE_IC();
            for (std::size_t i=0; i<c_.dirbcmaskp.size()/2; ++i) {
              auto p = c_.dirbcmaskp[i*2+0];
              auto mask = c_.dirbcmaskp[i*2+1];
              if (mask == 1) { auto val = np > 1 ? 0.0 : ic( x[p], y[p], z[p] ); dirbc[p] = {{ { 1, val } }}; }
              else if (mask == 2 && !c_.dirbcvalp.empty()) { auto val = np > 1 ? 0.0 : c_.dirbcvalp[i*2+1]; dirbc[p] = {{ { 1, val } }}; }
            }
          }
          if (auto pg = be::PRESSURE_GRAD()) {
            std::vector< std::uint8_t > besym( c_.triinpoel.size(), 0 );
            for (auto s : cfg.p_bc_sym) { auto kk = c_.bface.find( s ); if (kk != c_.bface.end()) for (auto f : kk->second) besym[f] = 1; }
            neubc.assign( x.size(), 0.0 );
            for (std::size_t e=0; e<c_.triinpoel.size()/3; ++e)
              if (besym[e]) {
                const auto N = c_.triinpoel.data() + e*3;
                real n[3]; port::crossdiv6( c_.coord, N, n );
                for (std::size_t a=0; a<3; ++a) { auto g = pg( x[N[a]], y[N[a]], z[N[a]] ); neubc[ N[a] ] -= n[0]*g[0] + n[1]*g[1] + n[2]*g[2]; }
              }
          }
          if (cfg.p_hydrostat != ~0ULL) {
            auto pi = c_.lid.find( cfg.p_hydrostat );
            if (pi != c_.lid.end()) {
              auto p = pi->second;
              auto ic = be::PRESSURE_IC();
              auto val = np > 1 ? 0.0 : ic( x[p], y[p], z[p] );
              auto& bb = dirbc[p];
              if (bb.empty()) bb = {{ { 1, val } }};
            }
          }
          if (auto pr = be::PRESSURE_RHS())
            for (std::size_t i=0; i<x.size(); ++i) c_.div[i] = pr( x[i], y[i], z[i] ) * c_.vol[i];
        }
        b[k] = c_.div;
      }
      cgpre.init( b, bcs, np < 3, cfg.p_pc );
    }
    //! ChoCG::psolve :1127-1142
    void psolve() { cgpre.solve( cfg.p_iter, cfg.p_tol ); pit = cgpre.it; }
    //! ChoCG::sgrad :1145-1168
    void sgrad() {
      for (std::size_t k=0; k<ch.size(); ++k) { auto& c_ = *ch[k];
        c_.sgrad.fill( 0.0 );
        be::chorin_grad( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, cgpre.parts[k]->x, c_.sgrad ); }
      sumShared( []( Chare& c_ ) -> be::Fields& { return c_.sgrad; } );
    }
    //! ChoCG::pgrad :1256-1280 + finpgrad :1306-1321
    void pgrad() {
      for (auto& cp : ch) { auto& c_ = *cp; c_.pgrad.fill( 0.0 );
        be::chorin_grad( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.pr, c_.pgrad ); }
      sumShared( []( Chare& c_ ) -> be::Fields& { return c_.pgrad; } );
      for (auto& cp : ch) fingrad( *cp, cp->pgrad );
    }
    //! ChoCG::psolved :1194-1253
    void psolved() {
      if (np != 1) {
        auto pdt = np > 1 ? dt : 1.0;
        for (auto& cp : ch) { auto& c_ = *cp;
          fingrad( c_, c_.sgrad );
          for (std::size_t i=0; i<c_.u.nunk(); ++i) {
            c_.u(i,0) -= pdt * c_.sgrad(i,0);
            c_.u(i,1) -= pdt * c_.sgrad(i,1);
            c_.u(i,2) -= pdt * c_.sgrad(i,2);
          }
          c_.BC( t + dt ); }
      }
      if (initial) {
        if (cfg.nstep == 1) {              // test first Poisson solve only
          for (std::size_t k=0; k<ch.size(); ++k) ch[k]->pr = cgpre.parts[k]->x;
          chodiag();
          finished = true;
        } else if (++np < 2) {
          velgrad(); flux(); div_flux();
          pinit(); psolve();               // m_np == 1: straight to psolved (:1138-1140)
          psolved();
        } else {
          for (std::size_t k=0; k<ch.size(); ++k) ch[k]->pr = cgpre.parts[k]->x;
          pgrad();
          initial = false;                 // start :1324-1337
        }
      } else {
        for (std::size_t k=0; k<ch.size(); ++k) { auto& c_ = *ch[k]; const auto& x = cgpre.parts[k]->x;
          for (std::size_t i=0; i<c_.pr.size(); ++i) c_.pr[i] += x[i]; }
        pgrad();
        chodiag();
      }
    }

    //! ChoCG::lhs :1433-1477: consistent mass / dt + theta * viscous Laplacian, the same for every component
    void lhs() {
      if (cfg.theta < std::numeric_limits< real >::epsilon()) return;
      for (std::size_t k=0; k<ch.size(); ++k) {
        auto& c_ = *ch[k];
        auto& A = cgmom.parts[k]->A;
        A.zero();
        const auto& X = c_.coord[0]; const auto& Y = c_.coord[1]; const auto& Z = c_.coord[2];
        auto ncomp = c_.u.nprop();
        for (std::size_t e=0; e<c_.inpoel.size()/4; ++e) {
          const auto N = c_.inpoel.data() + e*4;
          real ba[3] = { X[N[1]]-X[N[0]], Y[N[1]]-Y[N[0]], Z[N[1]]-Z[N[0]] },
               ca[3] = { X[N[2]]-X[N[0]], Y[N[2]]-Y[N[0]], Z[N[2]]-Z[N[0]] },
               da[3] = { X[N[3]]-X[N[0]], Y[N[3]]-Y[N[0]], Z[N[3]]-Z[N[0]] };
          auto cross = []( const real a[3], const real b[3], real r[3] ){
            r[0] = a[1]*b[2] - b[1]*a[2]; r[1] = a[2]*b[0] - b[2]*a[0]; r[2] = a[0]*b[1] - b[0]*a[1]; };
          real grad[4][3];
          cross( ca, da, grad[1] ); cross( da, ba, grad[2] ); cross( ba, ca, grad[3] );
          const auto J = ba[0]*grad[1][0] + ba[1]*grad[1][1] + ba[2]*grad[1][2];      // J = 6V
          for (std::size_t i=0; i<3; ++i) grad[0][i] = -grad[1][i]-grad[2][i]-grad[3][i];
          for (std::size_t a=0; a<4; ++a)
            for (std::size_t b=0; b<4; ++b) {
              auto v = J/dt/120.0 * ((a == b) ? 2.0 : 1.0);
              v += cfg.theta * cfg.mu * (grad[a][0]*grad[b][0] + grad[a][1]*grad[b][1] + grad[a][2]*grad[b][2]) / J / 6.0;
              for (std::size_t c=0; c<ncomp; ++c) A( N[a], N[b], c ) -= v;
            }
        }
      }
    }
    //! the semi-implicit branch of ChoCG::solve :1574-1607 + msolve :1610-1623 + msolved :1625-1645
    void msolve() {
      std::vector< std::vector< real > > b( ch.size() );
      std::vector< cg::BCs > bcs( ch.size() );
      for (std::size_t k=0; k<ch.size(); ++k) {
        auto& c_ = *ch[k];
        auto ncomp = c_.u.nprop(), nmask = ncomp + 1;
        auto& dirbc = bcs[k].dirbc;
        if (np < 3) {
          for (std::size_t i=0; i<c_.dirbcmasks.size()/nmask; ++i) {
            auto& bc = dirbc[ c_.dirbcmasks[i*nmask+0] ];
            bc.resize( ncomp );
            for (std::size_t c=0; c<ncomp; ++c) bc[c] = { static_cast< int >( c_.dirbcmasks[i*nmask+1+c] ), 0.0 };
          }
          for (auto p : c_.noslipbcnodes) {
            auto& bc = dirbc[p];
            bc.resize( ncomp );
            for (std::size_t c=0; c<ncomp; ++c) bc[c] = { 1, 0.0 };
          }
        }
        b[k] = c_.rhs.vec();
      }
      cgmom.init( b, bcs, np < 3, cfg.mom_pc );
      cgmom.solve( cfg.mom_iter, cfg.mom_tol ); mit = cgmom.it;
      for (std::size_t k=0; k<ch.size(); ++k) {
        auto& c_ = *ch[k]; const auto& du = cgmom.parts[k]->x;
        auto ncomp = c_.u.nprop();
        for (std::size_t i=0; i<c_.u.nunk(); ++i)
          for (std::size_t c=0; c<ncomp; ++c) c_.u(i,c) = c_.un(i,c) + du[i*ncomp+c];
      }
    }

    //! ChoCG::dt :1356-1411 (local minimum)
    real chodt( const Chare& c_ ) {
      auto eps = std::numeric_limits< real >::epsilon();
      if (std::abs( cfg.dt ) > eps) return cfg.dt;
      real mindt = std::numeric_limits< real >::max();
      auto large = std::numeric_limits< real >::max();
      auto dif = std::max( cfg.mu, cfg.dif );
      for (std::size_t i=0; i<c_.u.nunk(); ++i) {
        auto u = c_.u(i,0), v = c_.u(i,1), w = c_.u(i,2);
        auto vel = std::sqrt( u*u + v*v + w*w );
        auto L = std::cbrt( c_.vol[i] );
        auto euler_dt = L / std::max( vel, 1.0e-8 );
        mindt = std::min( mindt, euler_dt );
        auto dif_dt = dif > eps ? L * L / dif : large;
        mindt = std::min( mindt, dif_dt );
      }
      mindt *= cfg.cfl;
      if (t > cfg.freezetime) freezeflow = cfg.freezeflow;               // :1396-1399
      return mindt * freezeflow;
    }
    real freezeflow = 1.0;                // ChoCG::m_freezeflow

    //! one time step: dt :1356, advance :1414, rhs :1479, solve :1529, pred :1647, corr :1671,
    //! div :868, pinit, psolve, sgrad, psolved, pgrad, diag :1697
    bool step() override {
      if (finished) return false;
      real mindt = std::numeric_limits< real >::max();
      for (auto& cp : ch) mindt = std::min( mindt, chodt( *cp ) );
      auto eps = std::numeric_limits< real >::epsilon();
      if (mindt < eps) finished = true;
      dtn = dt; dt = mindt;
      if (t + dt > cfg.term) dt = cfg.term - t;
      lhs();                                                               // advance :1414-1431
      const bool implicit = cfg.theta > eps;
      for (std::size_t stage=0; stage<rk.size(); ++stage) {
        for (auto& cp : ch) { auto& c_ = *cp;
          be::chorin_rhs( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.v, t, c_.pr, c_.u, c_.grad, c_.rhs ); }
        sumShared( []( Chare& c_ ) -> be::Fields& { return c_.rhs; } );
        for (auto& cp : ch) if (stage == 0) cp->un = cp->u;
        // frozen flow (:1550-1552,1564-1570): the velocity of before the update comes back once pred() has
        // returned. In a serial run pred() runs through corr() (and, at the last stage, div()) inline, so those
        // see the updated velocity; the restored one enters the next rhs and the projection.
        const bool frozen = freezeflow > 1.0 && (!implicit || stage+1 < rk.size());
        std::vector< be::Fields > ufrozen;
        if (frozen) for (auto& cp : ch) ufrozen.push_back( cp->u );
        if (!implicit || stage+1 < rk.size()) {                            // solve :1555-1572
          for (auto& cp : ch) { auto& c_ = *cp;
            auto sdt = rk[stage] * dt;
            for (std::size_t i=0; i<c_.u.nunk(); ++i)
              for (std::size_t c=0; c<c_.u.nprop(); ++c) c_.u(i,c) = c_.un(i,c) - sdt*c_.rhs(i,c)/c_.vol[i]; }
        } else msolve();
        for (auto& cp : ch) { be::phys_src( cp->coord, t, cp->u );         // pred :1647-1668
          cp->BC( t + rk[stage] * dt ); }
        if (cfg.flux == "damp4") { velgrad(); for (auto& cp : ch) fingrad( *cp, cp->grad ); }   // corr :1677
        if (stage+1 == rk.size()) div_u();
        if (frozen)
          for (std::size_t k=0; k<ch.size(); ++k)
            for (std::size_t i=0; i<ch[k]->u.nunk(); ++i)
              for (std::size_t c=0; c<3; ++c) ch[k]->u(i,c) = ufrozen[k](i,c);
      }
      pinit(); psolve();
      sgrad(); psolved();
      if (done()) finished = true;
      return !finished;
    }

    //! ChoCG::diag :1697-1714 (next() first) + NodeDiagnostics::precompute :147-268 +
    //! Transporter::prediagnostics :1528-1608
    void chodiag() {
      ++it; t += dt;
      if ((it+1) % cfg.diag_iter) return;
      auto psol = be::PRESSURE_SOL();
      auto sol = be::SOL();
      std::size_t ncomp = psol ? 0 : cfg.ncomp;
      std::vector< std::vector< real > > d( 4, std::vector< real >( ncomp+1, 0.0 ) );
      for (std::size_t k=0; k<ch.size(); ++k) {
        auto& c_ = *ch[k];
        const auto& p = c_.pr; const auto& dp = cgpre.parts[k]->x; const auto& u = c_.u; const auto& un = c_.un; const auto& v = c_.v;
        std::vector< std::vector< real > > diag( 4, std::vector< real >( ncomp+1, 0.0 ) );
        for (std::size_t i=0; i<u.nunk(); ++i) {
          diag[0][0] += p[i] * p[i] * v[i];
          for (std::size_t c=0; c<ncomp; ++c) diag[0][c+1] += u(i,c) * u(i,c) * v[i];
          diag[1][0] += dp[i] * dp[i] * v[i];
          for (std::size_t c=0; c<ncomp; ++c) diag[1][c+1] += (u(i,c)-un(i,c)) * (u(i,c)-un(i,c)) * v[i];
          if (psol) { auto pd = p[i] - psol( c_.coord[0][i], c_.coord[1][i], c_.coord[2][i] );
            diag[2][0] += pd * pd * v[i]; diag[3][0] += std::abs( pd ) * v[i]; }
          if (sol) { auto s = sol( c_.coord[0][i], c_.coord[1][i], c_.coord[2][i], t+dt );   // T()+Dt() after next(), :213
            for (std::size_t c=0; c<ncomp; ++c) { auto du = u(i,c) - s[c]; diag[2][c+1] += du * du * v[i]; diag[3][c+1] += std::abs( du ) * v[i]; } }
        }
        for (std::size_t a=0; a<4; ++a) for (std::size_t c=0; c<=ncomp; ++c) d[a][c] += diag[a][c];
      }
      std::vector< real > row{ static_cast< real >( it ), t, dt };
      for (std::size_t i=0; i<=ncomp; ++i) row.push_back( std::sqrt( d[0][i] / meshvol ) );
      for (std::size_t i=0; i<=ncomp; ++i) row.push_back( std::sqrt( d[1][i] / meshvol ) );
      if (psol) { row.push_back( std::sqrt( d[2][0] / meshvol ) ); row.push_back( d[3][0] / meshvol ); }
      if (sol) {
        for (std::size_t i=1; i<=ncomp; ++i) row.push_back( std::sqrt( d[2][i] / meshvol ) );
        for (std::size_t i=1; i<=ncomp; ++i) row.push_back( d[3][i] / meshvol );
      }
      diagrows.push_back( std::move(row) );
    }
};

} // orc::

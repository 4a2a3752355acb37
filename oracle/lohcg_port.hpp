// oracle/lohcg_port.hpp -- TEST INFRASTRUCTURE ONLY (never linked by the product).
//
// Serial restatement of the LohCG solver chare of the reference (artificial-compressibility
// solver for constant-density flow, src/Inciter/LohCG.cpp) on top of the shared setup pipeline of
// driver.hpp, the Lohner edge operators (physics_port.hpp or, with -DORACLE_REF, the reference's
// own Lohner.cpp) and the conjugate gradients restatement (cg_port.hpp) that makes the initial
// velocity divergence-free. Unknowns (p,u,v,w), ncomp = 4. Each member cites the LohCG.cpp lines
// it follows; the SDAG control flow (src/Inciter/lohcg.ci) is unrolled into plain calls, chares
// are visited in index order.
#pragma once
#include "chocg_port.hpp"

namespace orc {

class LohRun : public Run {
  public:
    cg::Solver cgpre;
    int np = 0;                           // LohCG::m_np
    std::vector< real > rk;               // m_rkcoef, LohCG.cpp:42-47
    std::size_t pit = 0;

    LohRun( const MeshInput& in, const Cfg& c, const std::vector< std::size_t >& target, int nchare )
      : Run( in, c, target, nchare )
    {
      if (cfg.ncomp < 4) throw std::runtime_error( "oracle LohCG: unknowns (p,u,v,w) (+ transported scalars)" );
      static const std::vector< std::vector< real > > rkcoef{ { 1.0 }, { 1.0/2.0, 1.0 }, { 1.0/3.0, 1.0/2.0, 1.0 },
                                                               { 1.0/4.0, 1.0/3.0, 1.0/2.0, 1.0 } };
      rk = rkcoef.at( cfg.rk - 1 );
      // LohCG ctor :96-104 + prelhs :140-181: pressure Laplacian on the renumbered mesh
      for (auto& cp : ch) {
        auto& c_ = *cp;
        auto psup = be::genPsup( c_.inpoel, 4, be::genEsup( c_.inpoel, 4 ) );
        cg::CommMap cm;
        for (const auto& [k,n] : c_.nodeCommMap) cm[k] = n;
        auto k = cgpre.add( 1, psup, c_.gid, cm );
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
      // LohCG::merge :909-931 onwards: BCs at t + dt, then make the initial velocity divergence free
      for (auto& cp : ch) cp->BCnoP( t + dt );
      div( false );
      pinit(); psolve();
      sgrad(); psolved();
    }

    void sumSharedScalar( std::vector< real > Chare::*v ) {
      std::vector< std::unordered_map< std::size_t, real > > recv( ch.size() );
      for (std::size_t a=0; a<ch.size(); ++a)
        for (const auto& [b,n] : ch[a]->nodeCommMap)
          for (auto g : n) recv[static_cast<std::size_t>(b)][g] += ((*ch[a]).*v)[ ch[a]->lid.at( g ) ];
      for (std::size_t b=0; b<ch.size(); ++b)
        for (const auto& [g,r] : recv[b]) ((*ch[b]).*v)[ ch[b]->lid.at( g ) ] += r;
    }
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
    static void fingrad( Chare& c_, be::Fields& grad ) {            // :942-967 (sum over chares done by sumShared)
      for (std::size_t p=0; p<grad.nunk(); ++p) for (std::size_t c=0; c<grad.nprop(); ++c) grad(p,c) /= c_.vol[p];
    }

    //! LohCG::div :970-1001 of the velocity (of_flux = false, pos 1) or of the momentum flux (pos 0)
    void div( bool of_flux ) {
      for (auto& cp : ch) { auto& c_ = *cp;
        if (np == 1) { fingrad( c_, c_.mflux ); be::symbc( c_.mflux, c_.symbcnodes, c_.symbcnorms, 0 ); }
        std::fill( c_.div.begin(), c_.div.end(), 0.0 );
        if (of_flux) be::lohner_div( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.mflux, c_.div, 0 );
        else be::lohner_div( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.u, c_.div, 1 ); }
      sumSharedScalar( &Chare::div );
    }
    //! LohCG::velgrad :1025-1048 + flux :1074-1101
    void velgrad() {
      for (auto& cp : ch) { auto& c_ = *cp; c_.vgrad.fill( 0.0 );
        be::lohner_vgrad( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.u, c_.vgrad ); }
      sumShared( []( Chare& c_ ) -> be::Fields& { return c_.vgrad; } );
    }
    void flux() {
      for (auto& cp : ch) { auto& c_ = *cp;
        fingrad( c_, c_.vgrad );
        c_.mflux.fill( 0.0 );
        be::lohner_flux( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.u, c_.vgrad, c_.mflux ); }
      sumShared( []( Chare& c_ ) -> be::Fields& { return c_.mflux; } );
    }

    //! LohCG::pinit :1127-1219 (BCs always applied)
    void pinit() {
      std::vector< std::vector< real > > b( ch.size() );
      std::vector< cg::BCs > bcs( ch.size() );
      for (std::size_t k=0; k<ch.size(); ++k) {
        auto& c_ = *ch[k];
        const auto& x = c_.coord[0]; const auto& y = c_.coord[1]; const auto& z = c_.coord[2];
        auto& dirbc = bcs[k].dirbc; auto& neubc = bcs[k].neubc;
        if (!cfg.p_bc_dir.empty()) {
          auto ic = be::PRESSURE_IC();
          for (std::size_t i=0; i<c_.dirbcmaskp.size()/2; ++i) {
            auto p = c_.dirbcmaskp[i*2+0];
            auto mask = c_.dirbcmaskp[i*2+1];
            if (mask == 1) dirbc[p] = {{ { 1, ic( x[p], y[p], z[p] ) } }};
            else if (mask == 2 && !c_.dirbcvalp.empty()) dirbc[p] = {{ { 1, c_.dirbcvalp[i*2+1] } }};
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
        b[k] = c_.div;
      }
      cgpre.init( b, bcs, true, cfg.p_pc );
    }
    void psolve() { cgpre.solve( cfg.p_iter, cfg.p_tol ); pit = cgpre.it; }      // :1222-1237
    //! LohCG::sgrad :1240-1263
    void sgrad() {
      for (std::size_t k=0; k<ch.size(); ++k) { auto& c_ = *ch[k];
        c_.sgrad.fill( 0.0 );
        be::lohner_grad( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, cgpre.parts[k]->x, c_.sgrad ); }
      sumShared( []( Chare& c_ ) -> be::Fields& { return c_.sgrad; } );
    }
    //! LohCG::psolved :1289-1323 + transferIC :1326-1363
    void psolved() {
      if (np != 1) {
        for (auto& cp : ch) { auto& c_ = *cp;
          fingrad( c_, c_.sgrad );
          for (std::size_t i=0; i<c_.u.nunk(); ++i) {
            c_.u(i,1) -= c_.sgrad(i,0);
            c_.u(i,2) -= c_.sgrad(i,1);
            c_.u(i,3) -= c_.sgrad(i,2);
          }
          c_.BCnoP( t + dt ); }
      }
      if (cfg.nstep == 1) {                // test first Poisson solve only
        for (std::size_t k=0; k<ch.size(); ++k) { auto& c_ = *ch[k]; const auto& x = cgpre.parts[k]->x;
          for (std::size_t i=0; i<c_.u.nunk(); ++i) c_.u(i,0) = x[i]; }
        lohdiag();
        finished = true;
      } else if (++np < 2) {
        velgrad(); flux(); div( true );
        pinit(); psolve();                 // m_np == 1: straight to psolved (:1232-1234)
        psolved();
      } else {
        for (std::size_t k=0; k<ch.size(); ++k) { auto& c_ = *ch[k]; const auto& x = cgpre.parts[k]->x;
          for (std::size_t i=0; i<c_.u.nunk(); ++i) c_.u(i,0) = x[i]; }
      }
    }

    //! LohCG::dt :1401-1449 (local minimum)
    real lohdt( const Chare& c_ ) const {
      auto eps = std::numeric_limits< real >::epsilon();
      if (std::abs( cfg.dt ) > eps) return cfg.dt;
      real mindt = std::numeric_limits< real >::max();
      auto large = std::numeric_limits< real >::max();
      auto c = cfg.soundspeed;
      auto dif = std::max( cfg.mu, cfg.dif );
      for (std::size_t i=0; i<c_.u.nunk(); ++i) {
        auto u = c_.u(i,1), v = c_.u(i,2), w = c_.u(i,3);
        auto vel = std::sqrt( u*u + v*v + w*w );
        auto L = std::cbrt( c_.vol[i] );
        auto euler_dt = L / std::max( vel+c, 1.0e-8 );
        mindt = std::min( mindt, euler_dt );
        auto visc_dt = dif > eps ? L * L / dif : large;
        mindt = std::min( mindt, visc_dt );
      }
      return mindt * cfg.cfl;
    }

    //! one time step: dt :1401, advance :1452, stage/grad :1470-1508, rhs :1534, solve :1587, solved :1634, diag :1366
    bool step() override {
      if (finished) return false;
      real mindt = std::numeric_limits< real >::max();
      for (auto& cp : ch) mindt = std::min( mindt, lohdt( *cp ) );
      auto eps = std::numeric_limits< real >::epsilon();
      if (mindt < eps) finished = true;
      dtn = dt; dt = mindt;
      if (t + dt > cfg.term) dt = cfg.term - t;
      const bool damp4 = cfg.flux == "damp4";
      for (std::size_t stage=0; stage<rk.size(); ++stage) {
        if (damp4) {
          for (auto& cp : ch) { auto& c_ = *cp; c_.grad.fill( 0.0 );
            be::lohner_gradall( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.u, c_.grad ); }
          sumShared( []( Chare& c_ ) -> be::Fields& { return c_.grad; } );
          for (auto& cp : ch) fingrad( *cp, cp->grad );
        }
        static const be::Fields nograd;
        for (auto& cp : ch) { auto& c_ = *cp;
          be::lohner_rhs( c_.dsupedge, c_.dsupint, c_.coord, c_.triinpoel, c_.v, t, c_.u, damp4 ? c_.grad : nograd, c_.rhs ); }
        sumShared( []( Chare& c_ ) -> be::Fields& { return c_.rhs; } );
        for (auto& cp : ch) { auto& c_ = *cp;
          if (stage == 0) c_.un = c_.u;
          auto sdt = rk[stage] * dt;
          for (std::size_t i=0; i<c_.u.nunk(); ++i)
            for (std::size_t c=0; c<c_.u.nprop(); ++c) c_.u(i,c) = c_.un(i,c) - sdt*c_.rhs(i,c)/c_.vol[i];
          be::phys_src( c_.coord, t, c_.u );                              // :1615-1617
          c_.BC( t + rk[stage] * dt ); }
      }
      lohdiag();
      if (done()) finished = true;
      return !finished;
    }

    //! LohCG::diag :1366-1382 (next() first) + NodeDiagnostics::accompute :270-372 +
    //! Transporter::acdiagnostics :1621-1700
    void lohdiag() {
      ++it; t += dt;
      if ((it+1) % cfg.diag_iter) return;
      auto sol = be::SOL();
      auto ncomp = cfg.ncomp;
      std::vector< std::vector< real > > d( 4, std::vector< real >( ncomp, 0.0 ) );
      for (auto& cp : ch) {
        auto& c_ = *cp;
        const auto& u = c_.u; const auto& un = c_.un; const auto& v = c_.v;
        std::vector< std::vector< real > > diag( 4, std::vector< real >( ncomp, 0.0 ) );
        for (std::size_t i=0; i<u.nunk(); ++i) {
          for (std::size_t c=0; c<ncomp; ++c) diag[0][c] += u(i,c) * u(i,c) * v[i];
          for (std::size_t c=0; c<ncomp; ++c) diag[1][c] += (u(i,c)-un(i,c)) * (u(i,c)-un(i,c)) * v[i];
          if (sol) { auto s = sol( c_.coord[0][i], c_.coord[1][i], c_.coord[2][i], t+dt );
            for (std::size_t c=1; c<ncomp && c<s.size(); ++c) { auto du = u(i,c) - s[c]; diag[2][c] += du * du * v[i]; diag[3][c] += std::abs( du ) * v[i]; } }
        }
        for (std::size_t a=0; a<4; ++a) for (std::size_t c=0; c<ncomp; ++c) d[a][c] += diag[a][c];
      }
      std::vector< real > row{ static_cast< real >( it ), t, dt };
      for (std::size_t i=0; i<ncomp; ++i) row.push_back( std::sqrt( d[0][i] / meshvol ) );
      for (std::size_t i=0; i<ncomp; ++i) row.push_back( std::sqrt( d[1][i] / meshvol ) );
      if (sol) {
        for (std::size_t i=1; i<ncomp; ++i) row.push_back( std::sqrt( d[2][i] / meshvol ) );
        for (std::size_t i=1; i<ncomp; ++i) row.push_back( d[3][i] / meshvol );
      }
      diagrows.push_back( std::move(row) );
    }
};

} // orc::

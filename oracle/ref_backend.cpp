// oracle/ref_backend.cpp -- TEST INFRASTRUCTURE ONLY; compiled only into
// oracle/_ref/liboracle_ref.so together with the reference's own sources.
// Defines the reference's global configuration object (normally filled by its
// Lua parser, src/Control/InciterConfig.cpp) and fills the fields the hot path
// reads from the oracle's plain Cfg struct.
#include "backend.hpp"
#include "EOS.hpp"

namespace inciter { ctr::Config g_cfg; }

namespace orc { namespace be {

void set_cfg( const Cfg& c )
{
  auto& g = inciter::g_cfg;
  g = inciter::ctr::Config();
  g.get< tag::problem >() = c.problem;
  g.get< tag::flux >() = c.flux;
  g.get< tag::solver >() = c.solver;
  g.get< tag::fct >() = c.fct;
  g.get< tag::fctdif >() = c.fctdif;
  g.get< tag::fctclip >() = c.fctclip;
  g.get< tag::fctsys >() = c.fctsys;
  g.get< tag::problem_ncomp >() = c.ncomp;
  g.get< tag::mat_spec_heat_ratio >() = c.gamma;
  g.get< tag::problem_p0 >() = c.p0;
  g.get< tag::problem_alpha >() = c.alpha;
  g.get< tag::problem_kappa >() = c.kappa;
  g.get< tag::soundspeed >() = c.soundspeed;
  g.get< tag::problem_r0 >() = c.r0;
  g.get< tag::problem_ce >() = c.ce;
  g.get< tag::problem_beta >() = std::vector< double >{ c.beta[0], c.beta[1], c.beta[2] };
  g.get< tag::cfl >() = c.cfl;
  g.get< tag::dt >() = c.dt;
  g.get< tag::t0 >() = c.t0;
  g.get< tag::term >() = c.term;
  g.get< tag::nstep >() = c.nstep;
  g.get< tag::stab2 >() = c.stab2;
  g.get< tag::stab2coef >() = c.stab2coef;
  g.get< tag::steady >() = c.steady;
  g.get< tag::bc_sym >() = c.bc_sym;
  g.get< tag::bc_dir >() = c.bc_dir;
  g.get< tag::bc_far, tag::sidesets >() = c.bc_far;
  g.get< tag::bc_far, tag::density >() = c.far_density;
  g.get< tag::bc_far, tag::pressure >() = c.far_pressure;
  g.get< tag::bc_far, tag::velocity >() =
    std::vector< double >{ c.far_velocity[0], c.far_velocity[1], c.far_velocity[2] };
  g.get< tag::bc_pre, tag::sidesets >() = c.bc_pre;
  g.get< tag::bc_pre, tag::density >() = c.pre_density;
  g.get< tag::bc_pre, tag::pressure >() = c.pre_pressure;
  g.get< tag::diag_iter >() = c.diag_iter;
  g.get< tag::mat_spec_gas_const >() = c.rgas;
  g.get< tag::turkel >() = c.turkel;
  g.get< tag::velinf >() = std::vector< double >{ c.velinf[0], c.velinf[1], c.velinf[2] };
  g.get< tag::mat_dyn_viscosity >() = c.mu;
  g.get< tag::mat_dyn_diffusivity >() = c.dif;
  g.get< tag::stab >() = c.stab;
  g.get< tag::rk >() = c.rk;
  g.get< tag::residual >() = c.residual;
  g.get< tag::rescomp >() = c.rescomp;
  if (c.solver == "chocg" || c.solver == "lohcg")
    g.get< tag::ic, tag::velocity >() = std::vector< double >{ c.ic_velocity[0], c.ic_velocity[1], c.ic_velocity[2] };
  else if (c.problem == "userdef" || c.problem == "point_src") {
    g.get< tag::ic, tag::density >() = c.ic_density;
    g.get< tag::ic, tag::pressure >() = c.ic_pressure;
    g.get< tag::ic, tag::velocity >() = std::vector< double >{ c.ic_velocity[0], c.ic_velocity[1], c.ic_velocity[2] };
  }
  if (c.src_radius >= 0.0) {
    auto& s = g.get< tag::problem_src >();
    s.get< tag::location >() = std::vector< double >{ c.src_location[0], c.src_location[1], c.src_location[2] };
    s.get< tag::radius >() = c.src_radius;
    s.get< tag::release_time >() = c.src_release_time;
  }
  port::set_cfg( c );
}

real eos_pressure( real re ) { return eos::pressure( re ); }
real eos_soundspeed( real r, real p ) { return eos::soundspeed( r, p ); }

}} // orc::be::

// xyst_b200/host/problems.hpp -- host-side problem definitions used by the RieCG mirror:
// initial conditions, analytic solutions and source terms of the configured problem
// (cf. src/Physics/Problems.cpp: sedov::ic :337, sod::ic :373, taylor_green::ic/src
// :410/:433, vortical_flow::ic/src :454/:480, dispatch IC() :1071, SOL() :1114, SRC() :1299) and the ideal-gas EOS
// (src/Physics/EOS.hpp:29-56). Evaluated once on the host; the device receives arrays.
#pragma once
#include <array>
#include <cmath>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include "riecg.hpp"

namespace xyst {
namespace problems {

// values of all unknowns of a node: 5 flow variables + up to 8 transported scalars (unused entries 0)
using State = std::array< real, 13 >;
using Fn = std::function< State( real, real, real, real ) >;

inline real totalenergy( real g, real r, real u, real v, real w, real p ) {
  return p / (g-1.0) + 0.5 * r * (u*u + v*v + w*w);
}

//! scalar of problems::slot_cyl (Problems.cpp:549-629): a cone, a hump and a slotted cylinder carried by the
//! solid-body rotation about (0.5,0.5); 0 elsewhere
inline real slot_cyl_scalar( real x, real y, real t ) {
  using std::sin; using std::cos; using std::sqrt;
  real s = 0.0;
  const real R0 = 0.15;
  auto axdist = []( real x0, real y0 ){ return sqrt( (x0-0.5)*(x0-0.5) + (y0-0.5)*(y0-0.5) ); };
  real r = axdist( 0.5, 0.25 );                    // cone
  real kx = 0.5 + r*sin( t ), ky = 0.5 - r*cos( t );
  r = axdist( 0.25, 0.5 );                         // hump
  real hx = 0.5 + r*sin( t-M_PI/2.0 ), hy = 0.5 - r*cos( t-M_PI/2.0 );
  r = axdist( 0.5, 0.75 );                         // slotted cylinder
  real cx = 0.5 + r*sin( t+M_PI ), cy = 0.5 - r*cos( t+M_PI );
  // corner points of the slot, rotated with the flow
  real ax = 0.525, ay = cy - r*cos( std::asin( 0.025/r ) ), bx = 0.525, by = 0.8, gx = 0.475, gy = 0.8;
  auto rotx = [t]( real px, real py ){ return 0.5 + cos(t)*(px-0.5) - sin(t)*(py-0.5); };
  auto roty = [t]( real px, real py ){ return 0.5 + sin(t)*(px-0.5) + cos(t)*(py-0.5); };
  real rax = rotx( ax, ay ), ray = roty( ax, ay ), rbx = rotx( bx, by ), rby = roty( bx, by ),
       rgx = rotx( gx, gy ), rgy = roty( gx, gy );
  real v1x = rbx-rax, v1y = rby-ray, v2x = rgx-rbx, v2y = rgy-rby;
  real v1 = sqrt( v1x*v1x + v1y*v1y ), v2 = sqrt( v2x*v2x + v2y*v2y );
  r = sqrt( (x-kx)*(x-kx) + (y-ky)*(y-ky) ) / R0;
  if (r < 1.0) s = 0.6*(1.0-r);
  r = sqrt( (x-hx)*(x-hx) + (y-hy)*(y-hy) ) / R0;
  if (r < 1.0) s = 0.2*(1.0 + cos( M_PI*std::min( r, 1.0 ) ));
  r = sqrt( (x-cx)*(x-cx) + (y-cy)*(y-cy) ) / R0;
  real d1 = (v1x*(y-ray) - (x-rax)*v1y) / v1;      // signed distances from the two slot sides
  real d2 = (v2x*(y-rby) - (x-rbx)*v2y) / v2;
  if (r < 1.0 && (d1 > 0.05 || d1 < 0.0 || d2 < 0.0)) s = 0.6;
  return s;
}

inline Fn IC( const Config& cfg ) {
  const real g = cfg.gamma;
  if (cfg.solver == "lohcg") {                  // unknowns (p,u,v,w): entries 0..3
    const auto& p = cfg.problem;
    if (p == "userdef") { const auto vel = cfg.ic_velocity;                    // userdef::ic :53-65
      return [vel]( real, real, real, real ) -> State { return {{ 0, vel[0], vel[1], vel[2], 0 }}; }; }
    if (p == "point_src") { const auto vel = cfg.ic_velocity;                  // userdef::ic :53-65 (+ scalar 0)
      return [vel]( real, real, real, real ) -> State { return {{ 0, vel[0], vel[1], vel[2], 0 }}; }; }
    if (p == "slot_cyl")                                                       // slot_cyl::ic :538-544: (p,u,v,w,s)
      return []( real x, real y, real, real t ) -> State { return {{ 0.0, 0.5 - y, x - 0.5, 0.0, slot_cyl_scalar( x, y, t ) }}; };
    if (p == "poiseuille")                                                     // poiseuille::ic :1017-1019
      return []( real, real, real, real ) -> State { return {{ 0, 0, 0, 0, 0 }}; };
    throw std::runtime_error( "problem type ic not hooked up: " + p );
  }
  if (cfg.solver == "chocg") {                  // velocity unknowns only: entries 0..2
    const auto& p = cfg.problem;
    if (p == "userdef") { const auto vel = cfg.ic_velocity;                    // userdef::ic :44-52
      return [vel]( real, real, real, real ) -> State { return {{ vel[0], vel[1], vel[2], 0, 0 }}; }; }
    if (p.find( "poisson" ) != std::string::npos)
      return []( real, real, real, real ) -> State { return {{ 0, 0, 0, 0, 0 }}; };
    if (p == "point_src") { const auto vel = cfg.ic_velocity;                  // userdef::ic :44-52 (+ scalar 0)
      return [vel]( real, real, real, real ) -> State { return {{ vel[0], vel[1], vel[2], 0, 0 }}; }; }
    if (p == "slot_cyl")                                                       // slot_cyl::ic :531-537: (u,v,w,s)
      return []( real x, real y, real, real t ) -> State { return {{ 0.5 - y, x - 0.5, 0.0, slot_cyl_scalar( x, y, t ), 0 }}; };
    if (p == "poiseuille") { const real mu = cfg.mu;                           // poiseuille::ic :999-1026
      return [mu]( real, real y, real, real ) -> State {
        auto dpdx = -0.12;
        return {{ -dpdx * y * (1.0 - y) / 2.0 / mu, 0, 0, 0, 0 }}; }; }
    throw std::runtime_error( "problem type ic not hooked up: " + p );
  }
  if (cfg.problem == "slot_cyl")                // slot_cyl::ic, Problems.cpp:513-634: solid-body rotation about
    return [g]( real x, real y, real, real t ) -> State {        // (0.5,0.5) carrying a cone, a hump and a slotted cylinder
      State u{};
      const real p0 = 1.0;
      u[0] = 1.0; u[1] = u[0] * (0.5 - y); u[2] = u[0] * (x - 0.5); u[3] = 0.0;
      u[4] = totalenergy( g, u[0], u[1]/u[0], u[2]/u[0], u[3]/u[0], p0 );
      u[5] = slot_cyl_scalar( x, y, t );
      return u; };
  if (cfg.problem == "sedov") {
    const real p0 = cfg.p0;
    return [g,p0]( real x, real y, real z, real ) -> State {
      auto eps = std::numeric_limits< real >::epsilon();
      real p = (std::abs(x) < eps && std::abs(y) < eps && std::abs(z) < eps) ? p0 : 0.67e-4;
      real r = 1.0, u = 0.0, v = 0.0, w = 0.0;
      return {{ r, r*u, r*v, r*w, totalenergy( g, r, u, v, w, p ) }}; };
  }
  if (cfg.problem == "sod")
    return [g]( real x, real, real, real ) -> State {
      real r, p;
      if (x < 0.5) { r = 1.0; p = 1.0; } else { r = 0.125; p = 0.1; }
      real u = 0.0, v = 0.0, w = 0.0;
      return {{ r, r*u, r*v, r*w, totalenergy( g, r, u, v, w, p ) }}; };
  if (cfg.problem == "taylor_green")
    return [g]( real x, real y, real, real ) -> State {
      real r = 1.0;
      real p = 10.0 + r/4.0*(std::cos(2.0*M_PI*x) + std::cos(2.0*M_PI*y));
      real u =  std::sin(M_PI*x) * std::cos(M_PI*y);
      real v = -std::cos(M_PI*x) * std::sin(M_PI*y);
      real w = 0.0;
      return {{ r, r*u, r*v, r*w, totalenergy( g, r, u, v, w, p ) }}; };
  if (cfg.problem == "vortical_flow") {         // vortical_flow::ic :454-478
    const real a = cfg.alpha, k = cfg.kappa, p0 = cfg.p0;
    return [g,a,k,p0]( real x, real y, real z, real ) -> State {
      real ru = a*x - k*y;
      real rv = k*x + a*y;
      real rw = -2.0*a*z;
      real rE = (ru*ru + rv*rv + rw*rw)/2.0 + (p0 - 2.0*a*a*z*z) / (g - 1.0);
      return {{ 1.0, ru, rv, rw, rE }}; };
  }
  if (cfg.problem == "nonlinear_energy_growth") {   // nonlinear_energy_growth::ic :126-160
    const real ce = cfg.ce, r0 = cfg.r0, a = cfg.alpha, k = cfg.kappa; const auto b = cfg.beta;
    return [ce,r0,a,k,b]( real x, real y, real z, real t ) -> State {
      auto hx = std::cos(b[0]*M_PI*x) * std::cos(b[1]*M_PI*y) * std::cos(b[2]*M_PI*z);
      auto r = r0 + std::exp(-a*t) * (1.0 - x*x - y*y - z*z);
      auto re = r * std::pow( -3.0*(ce + k*hx*hx*t), -1.0/3.0 );
      return {{ r, 0.0, 0.0, 0.0, re }}; };
  }
  if (cfg.problem == "rayleigh_taylor") {       // rayleigh_taylor::ic :225-258
    const real a = cfg.alpha, p0 = cfg.p0, r0 = cfg.r0, k = cfg.kappa; const auto b = cfg.beta;
    return [g,a,p0,r0,k,b]( real x, real y, real z, real t ) -> State {
      real gx = b[0]*x*x + b[1]*y*y + b[2]*z*z;
      real r = r0 - gx;
      real ft = std::cos(k*M_PI*t);
      real u = ft * z * std::sin(M_PI*x);
      real v = ft * z * std::cos(M_PI*y);
      real w = ft * ( -0.5*M_PI*z*z*(std::cos(M_PI*x) - std::sin(M_PI*y)) );
      return {{ r, r*u, r*v, r*w, totalenergy( g, r, u, v, w, p0 + a*gx ) }}; };
  }
  if (cfg.problem == "userdef" || cfg.problem == "point_src") {     // userdef::ic :28-115, density + velocity + pressure (scalars 0)
    const real r = cfg.ic_density, p = cfg.ic_pressure;
    const auto vel = cfg.ic_velocity;
    return [g,r,p,vel]( real, real, real, real ) -> State {
      real ru = r*vel[0], rv = r*vel[1], rw = r*vel[2];
      return {{ r, ru, rv, rw, totalenergy( g, r, ru/r, rv/r, rw/r, p ) }}; };
  }
  throw std::runtime_error( "problem type ic not hooked up: " + cfg.problem );
}

inline Fn SOL( const Config& cfg ) {
  const auto& p = cfg.problem;
  if (p == "userdef" || p == "sod" || p == "sedov" || p == "point_src") return {};
  return IC( cfg );
}

//! IC (= Dirichlet values, analytic solution) or source term change in time
inline bool timeDependent( const Config& cfg ) {
  return cfg.problem == "nonlinear_energy_growth" || cfg.problem == "rayleigh_taylor" || cfg.problem == "slot_cyl";
}

inline Fn SRC( const Config& cfg ) {
  if (cfg.problem == "slot_cyl" && cfg.solver == "chocg") {      // slot_cyl::src :655-658: (u,v,w,s) unknowns
    auto ic = IC( cfg );
    return [ic]( real x, real y, real z, real t ) -> State {
      auto u = ic( x, y, z, t ); State s{}; s[0] = -u[1]; s[1] = u[0]; return s; };
  }
  if (cfg.problem == "slot_cyl") {                 // slot_cyl::src :636-670: centripetal momentum source
    auto ic = IC( cfg );
    return [ic]( real x, real y, real z, real t ) -> State {
      auto u = ic( x, y, z, t ); State s{}; s[1] = -u[2]; s[2] = u[1]; return s; };
  }
  if (cfg.problem == "nonlinear_energy_growth") {   // nonlinear_energy_growth::src :162-219
    const real a = cfg.alpha, ce = cfg.ce, kappa = cfg.kappa, r0 = cfg.r0, g = cfg.gamma; const auto b = cfg.beta;
    return [a,ce,kappa,r0,g,b]( real x, real y, real z, real t ) -> State {
      using std::sin; using std::cos; using std::pow;
      auto gx = 1.0 - x*x - y*y - z*z;
      std::array< real, 3 > dg{{ -2.0*x, -2.0*y, -2.0*z }};
      auto h = cos(b[0]*M_PI*x) * cos(b[1]*M_PI*y) * cos(b[2]*M_PI*z);
      std::array< real, 3 >
        dh{{ -b[0]*M_PI*sin(b[0]*M_PI*x)*cos(b[1]*M_PI*y)*cos(b[2]*M_PI*z),
             -b[1]*M_PI*cos(b[0]*M_PI*x)*sin(b[1]*M_PI*y)*cos(b[2]*M_PI*z),
             -b[2]*M_PI*cos(b[0]*M_PI*x)*cos(b[1]*M_PI*y)*sin(b[2]*M_PI*z) }};
      auto ft = std::exp(-a*t);
      auto dfdt = -a*ft;
      auto rho = r0 + ft*gx;
      std::array< real, 3 > drdx{{ ft*dg[0], ft*dg[1], ft*dg[2] }};
      auto drdt = gx*dfdt;
      auto ie = pow( -3.0*(ce + kappa*h*h*t), -1.0/3.0 );
      std::array< real, 3 > dedx{{ 2.0 * pow(ie,4.0) * kappa * h * dh[0] * t,
                                   2.0 * pow(ie,4.0) * kappa * h * dh[1] * t,
                                   2.0 * pow(ie,4.0) * kappa * h * dh[2] * t }};
      const auto dedt = kappa * h * h * pow(ie,4.0);
      State s{{ 0, 0, 0, 0, 0 }};
      s[0] = drdt;
      s[1] = (g-1.0)*(rho*dedx[0] + ie*drdx[0]);
      s[2] = (g-1.0)*(rho*dedx[1] + ie*drdx[1]);
      s[3] = (g-1.0)*(rho*dedx[2] + ie*drdx[2]);
      s[4] = rho*dedt + ie*drdt;
      return s; };
  }
  if (cfg.problem == "rayleigh_taylor") {       // rayleigh_taylor::src :260-332
    const real a = cfg.alpha, k = cfg.kappa, p0 = cfg.p0, g = cfg.gamma; const auto b = cfg.beta;
    auto ic = IC( cfg );
    return [a,k,p0,g,b,ic]( real x, real y, real z, real t ) -> State {
      using std::sin; using std::cos;
      auto U = ic( x, y, z, t );
      auto rho = U[0];
      auto u = U[1]/U[0];
      auto v = U[2]/U[0];
      auto w = U[3]/U[0];
      auto E = U[4]/U[0];
      auto p = p0 + a*(b[0]*x*x + b[1]*y*y + b[2]*z*z);
      std::array< real, 3 > drdx{{ -2.0*b[0]*x, -2.0*b[1]*y, -2.0*b[2]*z }};
      std::array< real, 3 > dpdx{{ 2.0*a*b[0]*x, 2.0*a*b[1]*y, 2.0*a*b[2]*z }};
      real ft = cos(k*M_PI*t);
      std::array< real, 3 > dudx{{ ft*M_PI*z*cos(M_PI*x), 0.0, ft*sin(M_PI*x) }};
      std::array< real, 3 > dvdx{{ 0.0, -ft*M_PI*z*sin(M_PI*y), ft*cos(M_PI*y) }};
      std::array< real, 3 > dwdx{{ ft*M_PI*0.5*M_PI*z*z*sin(M_PI*x),
                                   ft*M_PI*0.5*M_PI*z*z*cos(M_PI*y),
                                  -ft*M_PI*z*(cos(M_PI*x) - sin(M_PI*y)) }};
      std::array< real, 3 > dedx{{
        dpdx[0]/rho/(g-1.0) - p/(g-1.0)/rho/rho*drdx[0]
        + u*dudx[0] + v*dvdx[0] + w*dwdx[0],
        dpdx[1]/rho/(g-1.0) - p/(g-1.0)/rho/rho*drdx[1]
        + u*dudx[1] + v*dvdx[1] + w*dwdx[1],
        dpdx[2]/rho/(g-1.0) - p/(g-1.0)/rho/rho*drdx[2]
        + u*dudx[2] + v*dvdx[2] + w*dwdx[2] }};
      auto dudt = -k*M_PI*sin(k*M_PI*t)*z*sin(M_PI*x);
      auto dvdt = -k*M_PI*sin(k*M_PI*t)*z*cos(M_PI*y);
      auto dwdt =  k*M_PI*sin(k*M_PI*t)/2*M_PI*z*z*(cos(M_PI*x) - sin(M_PI*y));
      auto dedt = u*dudt + v*dvdt + w*dwdt;
      State s{{ 0, 0, 0, 0, 0 }};
      s[0] = u*drdx[0] + v*drdx[1] + w*drdx[2];
      s[1] = rho*dudt+u*s[0]+dpdx[0] + U[1]*dudx[0]+U[2]*dudx[1]+U[3]*dudx[2];
      s[2] = rho*dvdt+v*s[0]+dpdx[1] + U[1]*dvdx[0]+U[2]*dvdx[1]+U[3]*dvdx[2];
      s[3] = rho*dwdt+w*s[0]+dpdx[2] + U[1]*dwdx[0]+U[2]*dwdx[1]+U[3]*dwdx[2];
      s[4] = rho*dedt + E*s[0] + U[1]*dedx[0]+U[2]*dedx[1]+U[3]*dedx[2]
           + u*dpdx[0]+v*dpdx[1]+w*dpdx[2];
      return s; };
  }
  if (cfg.problem == "vortical_flow") {         // vortical_flow::src :480-507
    const real a = cfg.alpha, k = cfg.kappa, g = cfg.gamma;
    auto ic = IC( cfg );
    return [a,k,g,ic]( real x, real y, real z, real ) -> State {
      auto u = ic( x, y, z, 0.0 );
      State s{{ 0, 0, 0, 0, 0 }};
      s[1] = a*u[1]/u[0] - k*u[2]/u[0];
      s[2] = k*u[1]/u[0] + a*u[2]/u[0];
      s[4] = (s[1]*u[1] + s[2]*u[2])/u[0] + 8.0*a*a*a*z*z/(g-1.0);
      return s; };
  }
  if (cfg.problem == "taylor_green")
    return []( real x, real y, real, real ) -> State {
      State s{{ 0, 0, 0, 0, 0 }};
      s[4] = 3.0*M_PI/8.0*( std::cos(3.0*M_PI*x)*std::cos(M_PI*y)
                          - std::cos(3.0*M_PI*y)*std::cos(M_PI*x) );
      return s; };
  return {};
}

// pressure problems of the projection solvers, Problems.cpp:841-997,1172-1262
using PFn = std::function< real( real, real, real ) >;
inline PFn PRESSURE_RHS( const Config& cfg ) {
  const auto& p = cfg.problem;
  if (p == "poisson_const") return []( real, real, real ){ return 6.0; };
  if (p == "poisson_sine") return []( real x, real y, real z ){ return -M_PI * M_PI * x * y * std::sin( M_PI * z ); };
  if (p == "poisson_sine3") return []( real x, real y, real z ){
    return -3.0 * M_PI * M_PI * std::sin(M_PI*x) * std::sin(M_PI*y) * std::sin(M_PI*z); };
  if (p == "poisson_neumann") return []( real x, real y, real ){ return -3.0 * std::cos(2.0*x) * std::exp(y); };
  return {};
}
inline PFn PRESSURE_IC( const Config& cfg ) {
  const auto& p = cfg.problem;
  if (p == "userdef" || p == "slot_cyl" || p == "sheardiff" || p == "poiseuille" || p == "point_src")
    return []( real, real, real ){ return 0.0; };
  if (p == "poisson_const") return []( real x, real y, real z ){ return x*x + y*y + z*z; };
  if (p == "poisson_sine") return []( real x, real y, real z ){ return x * y * std::sin( M_PI * z ); };
  if (p == "poisson_sine3") return []( real x, real y, real z ){ return std::sin(M_PI*x) * std::sin(M_PI*y) * std::sin(M_PI*z); };
  if (p == "poisson_neumann") return []( real x, real y, real ){ return std::cos(2.0*x) * std::exp(y); };
  throw std::runtime_error( "pressure ic not hooked up: " + p );
}
inline PFn PRESSURE_SOL( const Config& cfg ) {
  const auto& p = cfg.problem;
  if (p == "userdef" || p == "slot_cyl" || p == "poiseuille" || p == "point_src" || p == "sheardiff") return {};
  return PRESSURE_IC( cfg );
}
inline std::function< std::array< real, 3 >( real, real, real ) > PRESSURE_GRAD( const Config& cfg ) {
  if (cfg.problem == "poisson_neumann")
    return []( real x, real y, real ) -> std::array< real, 3 > {
      return {{ -2.0 * std::sin( 2.0 * x ) * std::exp( y ), std::cos(2.0*x) * std::exp(y), 0.0 }}; };
  return {};
}

} // problems::
} // xyst::

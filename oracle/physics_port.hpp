// oracle/physics_port.hpp -- TEST INFRASTRUCTURE ONLY (never linked by the product).
//
// Plain serial C++ restatement of the arithmetic of the reference's RieCG hot
// path, written to be bit-identical to the reference when compiled without FMA
// contraction (the reference's GNU Release build has none, SURVEY.md 8a'-14):
//   nodal gradients            src/Physics/Riemann.cpp:229-367
//   MUSCL + van Leer           src/Physics/Riemann.cpp:34-143 (flow), :145-209 (scalars)
//   Rusanov / HLLC edge flux   src/Physics/Riemann.cpp:369-478 / :480-650
//   domain/boundary/source rhs src/Physics/Riemann.cpp:652-946
//   ideal gas EOS              src/Physics/EOS.hpp:29-56
//   nodal BCs                  src/Physics/BC.cpp:29-241
//   problem ICs / sources      src/Physics/Problems.cpp:337-440,1071-1330
// It is pinned against the reference's own objects (oracle/_ref, built from
// /root/reference in place) by tests/test_oracle_ref.py and against the
// reference's regression goldens by tests/test_oracle_golden.py.
#pragma once
#include <vector>
#include <array>
#include <string>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <stdexcept>
#include <functional>
#include <unordered_set>
#include <algorithm>

namespace orc {

using real = double;

//! Run configuration: the subset of the reference's g_cfg read on this path
//! (src/Control/InciterConfig.hpp:161-413; defaults InciterConfig.cpp:1707-1757)
struct Cfg {
  std::string problem = "userdef";
  std::string flux = "rusanov";
  std::size_t ncomp = 5;
  real gamma = 1.4;             // mat_spec_heat_ratio
  real p0 = 0.0;                // problem_p0 (sedov, vortical_flow)
  real alpha = 0.0, kappa = 0.0; // problem_alpha, problem_kappa (manufactured solutions)
  real r0 = 0.0, ce = 0.0;       // problem_r0, problem_ce
  real soundspeed = 1.0;         // LohCG artificial speed of sound (tag::soundspeed)
  std::array< real, 3 > beta{{0,0,0}};   // problem_beta
  real cfl = 0.0;
  real dt = 0.0;                // constant dt if |dt|>eps
  real t0 = 0.0;
  real term = 1.0e+300;
  std::uint64_t nstep = ~0ULL;
  bool stab2 = false;
  real stab2coef = 0.2;
  bool steady = false;
  std::vector< int > bc_sym;
  std::vector< std::vector< int > > bc_dir;   // { setid, mask_0 .. mask_{ncomp-1} }
  std::vector< int > bc_far;
  real far_density = 0.0, far_pressure = 0.0;
  std::array< real, 3 > far_velocity{{0,0,0}};
  std::vector< std::vector< int > > bc_pre;   // side set groups
  std::vector< real > pre_density, pre_pressure;
  std::vector< int > fieldout_sets, integout_sets;
  std::uint64_t diag_iter = 1;
  // ZalCG (flux-corrected transport), defaults InciterConfig.cpp:1751-1757
  std::string solver = "riecg";
  bool fct = true;
  real fctdif = 1.0;
  bool fctclip = false;
  std::vector< std::uint64_t > fctsys;       // 1-based component ids limited as a system
  // LaxCG (time-derivative preconditioning), defaults InciterConfig.cpp:1409,1731-1739
  real rgas = 287.052874;                    // mat_spec_gas_const
  real turkel = 0.5;
  std::array< real, 3 > velinf{{1.0,1.0,1.0}};
  real residual = 0.0;
  std::uint64_t rescomp = 1;
  // ic block of user-defined problems (Problems.cpp:28-115)
  real ic_density = 0.0, ic_pressure = 0.0;
  std::array< real, 3 > ic_velocity{{0,0,0}};
  // ChoCG (projection method), defaults InciterConfig.cpp:1411-1412,1558-1564,1728-1748
  real mu = 0.0, dif = 0.0;                  // mat_dyn_viscosity, mat_dyn_diffusivity
  bool stab = true;
  std::uint64_t rk = 1;
  std::vector< int > bc_noslip;
  std::vector< std::vector< real > > bc_dirval;      // { setid, val_0 .. }
  std::uint64_t p_iter = 10;
  real p_tol = 1.0e-3;
  std::string p_pc = "none";
  std::vector< std::vector< int > > p_bc_dir;        // { setid, mask }
  std::vector< std::vector< real > > p_bc_dirval;    // { setid, val }
  std::vector< int > p_bc_sym;
  std::uint64_t p_hydrostat = ~0ULL;
  // problems::point_src (problem = { src = { location, radius, release_time } }); radius < 0 = not configured
  std::array< real, 3 > src_location{{ 0, 0, 0 }};
  real src_radius = -1.0, src_release_time = 0.0;
  // ZalCG/KozCG: freeze the flow after freezetime and advance the scalars with freezeflow x dt
  real freezeflow = 1.0, freezetime = 0.0;
  real fctfreeze = -1.0e300;     // ZalCG steady state: freeze the FCT limit coefficients once the residual is below this (default: never)
  // semi-implicit momentum solve of ChoCG (tag::theta, mom_iter, mom_tol, mom_pc)
  real theta = 0.0;
  std::uint64_t mom_iter = 10;
  real mom_tol = 1.0e-3;
  std::string mom_pc = "none";
};

//! Nodal field container with the reference's default layout [node][component]
//! (src/Base/Data.hpp:420-426 with FIELD_DATA_LAYOUT_AS_FIELD_MAJOR)
class PFields {
  public:
    PFields() : m_n(0), m_p(0) {}
    PFields( std::size_t n, std::size_t p ) : m_v( n*p, 0.0 ), m_n(n), m_p(p) {}
    real& operator()( std::size_t i, std::size_t c ) { return m_v[ i*m_p + c ]; }
    const real& operator()( std::size_t i, std::size_t c ) const { return m_v[ i*m_p + c ]; }
    std::vector< real > operator[]( std::size_t i ) const {
      return std::vector< real >( m_v.begin()+static_cast<long>(i*m_p),
                                  m_v.begin()+static_cast<long>((i+1)*m_p) ); }
    std::size_t nunk() const { return m_n; }
    std::size_t nprop() const { return m_p; }
    void fill( real a ) { std::fill( m_v.begin(), m_v.end(), a ); }
    std::vector< real >& vec() { return m_v; }
    const std::vector< real >& vec() const { return m_v; }
  private:
    std::vector< real > m_v;
    std::size_t m_n, m_p;
};

namespace port {

using Fields = PFields;
using Coords = std::array< std::vector< real >, 3 >;

inline Cfg& cfg() { static Cfg c; return c; }
inline void set_cfg( const Cfg& c ) { cfg() = c; }

// ---- EOS (EOS.hpp:29-56) ----------------------------------------------------
inline real eos_pressure( real re ) { auto g = cfg().gamma; return re * (g-1.0); }
inline real eos_soundspeed( real r, real p ) { auto g = cfg().gamma; return std::sqrt( g * p / r ); }
inline real eos_totalenergy( real r, real u, real v, real w, real p ) {
  auto g = cfg().gamma; return p / (g-1.0) + 0.5 * r * (u*u + v*v + w*w); }

// ---- problems (Problems.cpp) -------------------------------------------------
using ICFn = std::function< std::vector< real >( real, real, real, real ) >;

inline std::vector< real > ic_sedov( real x, real y, real z, real ) {      // :337-367
  auto eps = std::numeric_limits< real >::epsilon();
  real p;
  if (std::abs(x) < eps && std::abs(y) < eps && std::abs(z) < eps) p = cfg().p0;
  else p = 0.67e-4;
  real r = 1.0, u = 0.0, v = 0.0, w = 0.0;
  real rE = eos_totalenergy( r, u, v, w, p );
  return { r, r*u, r*v, r*w, rE };
}
inline std::vector< real > ic_sod( real x, real, real, real ) {            // :373-404
  real r, p;
  if (x < 0.5) { r = 1.0; p = 1.0; } else { r = 0.125; p = 0.1; }
  real u = 0.0, v = 0.0, w = 0.0;
  real rE = eos_totalenergy( r, u, v, w, p );
  return { r, r*u, r*v, r*w, rE };
}
inline std::vector< real > ic_taylor_green( real x, real y, real, real ) { // :410-431
  real r = 1.0;
  real p = 10.0 + r/4.0*(std::cos(2.0*M_PI*x) + std::cos(2.0*M_PI*y));
  real u =  std::sin(M_PI*x) * std::cos(M_PI*y);
  real v = -std::cos(M_PI*x) * std::sin(M_PI*y);
  real w = 0.0;
  auto rE = eos_totalenergy( r, u, v, w, p );
  return { r, r*u, r*v, r*w, rE };
}
inline std::vector< real > src_taylor_green( real x, real y, real, real ) { // :433-452
  std::vector< real > s( 5, 0.0 );
  s[4] = 3.0*M_PI/8.0*( std::cos(3.0*M_PI*x)*std::cos(M_PI*y)
                      - std::cos(3.0*M_PI*y)*std::cos(M_PI*x) );
  return s;
}

inline std::vector< real > ic_vortical_flow( real x, real y, real z, real ) { // :454-478
  auto a = cfg().alpha, k = cfg().kappa, p0 = cfg().p0, g = cfg().gamma;
  real ru = a*x - k*y;
  real rv = k*x + a*y;
  real rw = -2.0*a*z;
  real rE = (ru*ru + rv*rv + rw*rw)/2.0 + (p0 - 2.0*a*a*z*z) / (g - 1.0);
  return { 1.0, ru, rv, rw, rE };
}
inline std::vector< real > src_vortical_flow( real x, real y, real z, real ) { // :480-507
  auto a = cfg().alpha, k = cfg().kappa, g = cfg().gamma;
  auto u = ic_vortical_flow( x, y, z, 0.0 );
  std::vector< real > s( 5, 0.0 );
  s[1] = a*u[1]/u[0] - k*u[2]/u[0];
  s[2] = k*u[1]/u[0] + a*u[2]/u[0];
  s[4] = (s[1]*u[1] + s[2]*u[2])/u[0] + 8.0*a*a*a*z*z/(g-1.0);
  return s;
}
inline std::vector< real > ic_nleg( real x, real y, real z, real t ) {       // nonlinear_energy_growth::ic :126-160
  using std::cos;
  auto ce = cfg().ce, r0 = cfg().r0, a = cfg().alpha, k = cfg().kappa; const auto& b = cfg().beta;
  auto ec = [ ce, t ]( real kappa, real h, real p ) { return std::pow( -3.0*(ce + kappa*h*h*t), p ); };
  auto hx = cos(b[0]*M_PI*x) * cos(b[1]*M_PI*y) * cos(b[2]*M_PI*z);
  auto r = r0 + std::exp(-a*t) * (1.0 - x*x - y*y - z*z);
  auto re = r * ec(k,hx,-1.0/3.0);
  return { r, 0.0, 0.0, 0.0, re };
}
inline std::vector< real > src_nleg( real x, real y, real z, real t ) {      // :162-219
  using std::sin; using std::cos; using std::pow;
  auto a = cfg().alpha; const auto& b = cfg().beta; auto ce = cfg().ce, kappa = cfg().kappa, r0 = cfg().r0, g = cfg().gamma;
  auto gx = 1.0 - x*x - y*y - z*z;
  std::array< real, 3 > dg{ -2.0*x, -2.0*y, -2.0*z };
  auto h = cos(b[0]*M_PI*x) * cos(b[1]*M_PI*y) * cos(b[2]*M_PI*z);
  std::array< real, 3 >
    dh{ -b[0]*M_PI*sin(b[0]*M_PI*x)*cos(b[1]*M_PI*y)*cos(b[2]*M_PI*z),
        -b[1]*M_PI*cos(b[0]*M_PI*x)*sin(b[1]*M_PI*y)*cos(b[2]*M_PI*z),
        -b[2]*M_PI*cos(b[0]*M_PI*x)*cos(b[1]*M_PI*y)*sin(b[2]*M_PI*z) };
  auto ft = std::exp(-a*t);
  auto dfdt = -a*ft;
  auto rho = r0 + ft*gx;
  std::array< real, 3 > drdx{ ft*dg[0], ft*dg[1], ft*dg[2] };
  auto drdt = gx*dfdt;
  auto ie = pow( -3.0*(ce + kappa*h*h*t), -1.0/3.0 );
  std::array< real, 3 > dedx{ 2.0 * pow(ie,4.0) * kappa * h * dh[0] * t,
                              2.0 * pow(ie,4.0) * kappa * h * dh[1] * t,
                              2.0 * pow(ie,4.0) * kappa * h * dh[2] * t };
  const auto dedt = kappa * h * h * pow(ie,4.0);
  std::vector< real > s( 5, 0.0 );
  s[0] = drdt;
  s[1] = (g-1.0)*(rho*dedx[0] + ie*drdx[0]);
  s[2] = (g-1.0)*(rho*dedx[1] + ie*drdx[1]);
  s[3] = (g-1.0)*(rho*dedx[2] + ie*drdx[2]);
  s[4] = rho*dedt + ie*drdt;
  return s;
}
inline std::vector< real > ic_rayleigh_taylor( real x, real y, real z, real t ) {   // rayleigh_taylor::ic :225-258
  using std::sin; using std::cos;
  auto a = cfg().alpha; const auto& b = cfg().beta; auto p0 = cfg().p0, r0 = cfg().r0, k = cfg().kappa;
  real gx = b[0]*x*x + b[1]*y*y + b[2]*z*z;
  real r = r0 - gx;
  real ft = cos(k*M_PI*t);
  real u = ft * z * sin(M_PI*x);
  real v = ft * z * cos(M_PI*y);
  real w = ft * ( -0.5*M_PI*z*z*(cos(M_PI*x) - sin(M_PI*y)) );
  real rE = eos_totalenergy( r, u, v, w, p0 + a*gx );
  return { r, r*u, r*v, r*w, rE };
}
inline std::vector< real > src_rayleigh_taylor( real x, real y, real z, real t ) {  // :260-332
  using std::sin; using std::cos;
  auto a = cfg().alpha; const auto& b = cfg().beta; auto k = cfg().kappa, p0 = cfg().p0, g = cfg().gamma;
  auto U = ic_rayleigh_taylor( x, y, z, t );
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
  std::vector< real > s( 5, 0.0 );
  s[0] = u*drdx[0] + v*drdx[1] + w*drdx[2];
  s[1] = rho*dudt+u*s[0]+dpdx[0] + U[1]*dudx[0]+U[2]*dudx[1]+U[3]*dudx[2];
  s[2] = rho*dvdt+v*s[0]+dpdx[1] + U[1]*dvdx[0]+U[2]*dvdx[1]+U[3]*dvdx[2];
  s[3] = rho*dwdt+w*s[0]+dpdx[2] + U[1]*dwdx[0]+U[2]*dwdx[1]+U[3]*dwdx[2];
  s[4] = rho*dedt + E*s[0] + U[1]*dedx[0]+U[2]*dedx[1]+U[3]*dedx[2]
       + u*dpdx[0]+v*dpdx[1]+w*dpdx[2];
  return s;
}
inline std::vector< real > ic_userdef( real, real, real, real ) {           // :28-115 (pressure given)
  std::vector< real > u( cfg().ncomp, 0.0 );
  u[0] = cfg().ic_density;
  u[1] = u[0] * cfg().ic_velocity[0];
  u[2] = u[0] * cfg().ic_velocity[1];
  u[3] = u[0] * cfg().ic_velocity[2];
  u[4] = eos_totalenergy( u[0], u[1]/u[0], u[2]/u[0], u[3]/u[0], cfg().ic_pressure );
  return u;
}

//! slot_cyl: solid-body rotation about (0.5,0.5) in the x-y plane carrying a cone, a cosine hump and a
//! slotted cylinder in the first scalar, Problems.cpp:511-672. The bodies start a quarter turn apart
//! at distance 0.25 from the axis and have radius 0.15; unknowns depend on the solver (6 for the
//! compressible solvers, 3 velocities + scalar for chocg, p + velocities + scalar for lohcg).
inline std::vector< real > ic_slot_cyl( real x, real y, real, real t ) {    // :513-634
  using std::sin; using std::cos; using std::sqrt;
  const auto& solver = cfg().solver;
  const bool cho = solver == "chocg", loh = solver == "lohcg";
  std::vector< real > u( cho ? 4 : loh ? 5 : 6, 0.0 );
  std::size_t sc = cho ? 3 : loh ? 4 : 5;
  if (cho) { u[0] = 0.5 - y; u[1] = x - 0.5; }
  else if (loh) { u[1] = 0.5 - y; u[2] = x - 0.5; }
  else {
    const real p0 = 1.0;
    u[0] = 1.0; u[1] = u[0] * (0.5 - y); u[2] = u[0] * (x - 0.5); u[3] = 0.0;
    u[4] = eos_totalenergy( u[0], u[1]/u[0], u[2]/u[0], u[3]/u[0], p0 );
  }
  const real R0 = 0.15;
  // distance of a body's initial centre (x0,y0) from the rotation axis
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
  if (r < 1.0) u[sc] = 0.6*(1.0-r);
  r = sqrt( (x-hx)*(x-hx) + (y-hy)*(y-hy) ) / R0;
  if (r < 1.0) u[sc] = 0.2*(1.0 + cos( M_PI*std::min( r, 1.0 ) ));
  r = sqrt( (x-cx)*(x-cx) + (y-cy)*(y-cy) ) / R0;
  // signed distances from the two slot sides
  real d1 = (v1x*(y-ray) - (x-rax)*v1y) / v1;
  real d2 = (v2x*(y-rby) - (x-rbx)*v2y) / v2;
  if (r < 1.0 && (d1 > 0.05 || d1 < 0.0 || d2 < 0.0)) u[sc] = 0.6;
  return u;
}
inline std::vector< real > src_slot_cyl( real x, real y, real z, real t ) { // :636-670: centripetal momentum source
  auto u = ic_slot_cyl( x, y, z, t );
  std::vector< real > s( u.size(), 0.0 );
  if (cfg().solver == "chocg") { s[0] = -u[1]; s[1] = u[0]; }
  else { s[1] = -u[2]; s[2] = u[1]; }
  return s;
}

inline std::vector< real > ic_poiseuille( real, real y, real, real ) {      // :999-1026 (chocg)
  auto dpdx = -0.12;
  auto u = -dpdx * y * (1.0 - y) / 2.0 / cfg().mu;
  return { u, 0.0, 0.0 };
}

inline ICFn IC() {                                                          // :1071-1108
  const auto& p = cfg().problem;
  if (cfg().solver == "lohcg") {             // unknowns (p,u,v,w)
    if (p == "userdef" || p == "point_src") return []( real, real, real, real ){      // :53-65
      std::vector< real > u( cfg().ncomp, 0.0 );
      u[1] = cfg().ic_velocity[0]; u[2] = cfg().ic_velocity[1]; u[3] = cfg().ic_velocity[2];
      return u; };
    if (p == "poiseuille") return []( real, real, real, real ){ return std::vector< real >{ 0, 0, 0, 0 }; };   // :1017-1019
    if (p == "slot_cyl") return ic_slot_cyl;
    throw std::runtime_error( "oracle port: problem type ic not hooked up for lohcg: " + p );
  }
  if (cfg().solver == "chocg") {             // velocity unknowns only
    if (p == "userdef" || p == "point_src") return []( real, real, real, real ){      // :44-52
      std::vector< real > u( cfg().ncomp, 0.0 );
      u[0] = cfg().ic_velocity[0]; u[1] = cfg().ic_velocity[1]; u[2] = cfg().ic_velocity[2];
      return u; };
    if (p.find("poisson") != std::string::npos) return []( real, real, real, real ){ return std::vector< real >{ 0, 0, 0 }; };
    if (p == "poiseuille") return ic_poiseuille;
  }
  if (p == "userdef" || p == "point_src") return ic_userdef;                // :1078-1079
  if (p == "sedov") return ic_sedov;
  if (p == "sod") return ic_sod;
  if (p == "taylor_green") return ic_taylor_green;
  if (p == "slot_cyl") return ic_slot_cyl;
  if (p == "vortical_flow") return ic_vortical_flow;
  if (p == "nonlinear_energy_growth") return ic_nleg;
  if (p == "rayleigh_taylor") return ic_rayleigh_taylor;
  throw std::runtime_error( "oracle port: problem type ic not hooked up: " + p );
}
inline ICFn SOL() {                                                         // :1114-1131
  const auto& p = cfg().problem;
  if (p == "userdef" || p == "sod" || p == "sedov" || p == "point_src") return {};
  return IC();
}

// pressure problems of the projection solvers, Problems.cpp:841-997,1172-1262
using PFn = std::function< real( real, real, real ) >;
inline PFn PRESSURE_RHS() {
  const auto& p = cfg().problem;
  if (p == "poisson_const") return []( real, real, real ){ return 6.0; };
  if (p == "poisson_sine") return []( real x, real y, real z ){ return -M_PI * M_PI * x * y * std::sin( M_PI * z ); };
  if (p == "poisson_sine3") return []( real x, real y, real z ){ using std::sin;
    return -3.0 * M_PI * M_PI * sin(M_PI*x) * sin(M_PI*y) * sin(M_PI*z); };
  if (p == "poisson_neumann") return []( real x, real y, real ){ return -3.0 * std::cos(2.0*x) * std::exp(y); };
  return {};
}
inline PFn PRESSURE_IC() {
  const auto& p = cfg().problem;
  if (p == "userdef" || p == "slot_cyl" || p == "sheardiff" || p == "poiseuille" || p == "point_src")
    return []( real, real, real ){ return 0.0; };
  if (p == "poisson_const") return []( real x, real y, real z ){ return x*x + y*y + z*z; };
  if (p == "poisson_sine") return []( real x, real y, real z ){ return x * y * std::sin( M_PI * z ); };
  if (p == "poisson_sine3") return []( real x, real y, real z ){ using std::sin; return sin(M_PI*x) * sin(M_PI*y) * sin(M_PI*z); };
  if (p == "poisson_neumann") return []( real x, real y, real ){ return std::cos(2.0*x) * std::exp(y); };
  throw std::runtime_error( "oracle port: pressure ic not hooked up: " + p );
}
inline PFn PRESSURE_SOL() {
  const auto& p = cfg().problem;
  if (p == "userdef" || p == "slot_cyl" || p == "poiseuille" || p == "point_src" || p == "sheardiff") return {};
  return PRESSURE_IC();
}
inline std::function< std::array< real, 3 >( real, real, real ) > PRESSURE_GRAD() {
  if (cfg().problem == "poisson_neumann")
    return []( real x, real y, real ) -> std::array< real, 3 > {
      return {{ -2.0 * std::sin( 2.0 * x ) * std::exp( y ), std::cos(2.0*x) * std::exp(y), 0.0 }}; };
  return {};
}
inline ICFn SRC() {                                                         // :1299-1325
  const auto& p = cfg().problem;
  if (p == "taylor_green") return src_taylor_green;
  if (p == "slot_cyl") return src_slot_cyl;
  if (p == "vortical_flow") return src_vortical_flow;
  if (p == "nonlinear_energy_growth") return src_nleg;
  if (p == "rayleigh_taylor") return src_rayleigh_taylor;
  return {};
}

//! problems::PHYS_SRC() -> point_src::src, Problems.cpp:764-823: a source that sets the first scalar to 1
//! inside a sphere from the release time on, applied directly to the solution (not a rhs term)
inline void phys_src( const Coords& coord, real t, Fields& U ) {
  if (cfg().problem != "point_src") return;
  if (U.nprop() == 5) return;
  if (cfg().src_radius < 0.0) return;
  if (t < cfg().src_release_time) return;
  std::size_t sc = cfg().solver == "chocg" ? 3 : cfg().solver == "lohcg" ? 4 : 5;
  const auto& s = cfg().src_location; auto sr = cfg().src_radius;
  for (std::size_t i=0; i<U.nunk(); ++i) {
    auto rx = s[0] - coord[0][i], ry = s[1] - coord[1][i], rz = s[2] - coord[2][i];
    if (rx*rx + ry*ry + rz*rz < sr*sr) U(i,sc) = 1.0;
  }
}

inline void initialize( const Coords& coord, Fields& U, real t ) {          // :1134-1167
  auto ic = IC();
  for (std::size_t i=0; i<coord[0].size(); ++i) {
    auto s = ic( coord[0][i], coord[1][i], coord[2][i], t );
    for (std::size_t c=0; c<s.size(); ++c) U(i,c) = s[c];
  }
}

// ---- BCs (BC.cpp) --------------------------------------------------------------
inline void dirbc( Fields& U, real t, const Coords& coord,
                   const std::vector< std::size_t >& dirbcmask,
                   const std::vector< double >& dirbcval = {} )             // :29-72
{
  auto ncomp = U.nprop();
  auto nmask = ncomp + 1;
  if (dirbcmask.empty()) return;
  auto ic = IC();
  for (std::size_t i=0; i<dirbcmask.size()/nmask; ++i) {
    auto p = dirbcmask[i*nmask+0];
    auto u = ic( coord[0][p], coord[1][p], coord[2][p], t );
    for (std::size_t c=0; c<ncomp; ++c) {
      auto mask = dirbcmask[i*nmask+1+c];
      if (mask == 1) U(p,c) = u[c];
      else if (mask == 2 && !dirbcval.empty()) U(p,c) = dirbcval[i*nmask+1+c];
    }
  }
}

//! pressure Dirichlet BC of LohCG, BC.cpp:74-108 (defined after PRESSURE_IC below)
inline void dirbcp( Fields& U, const Coords& coord, const std::vector< std::size_t >& dirbcmaskp,
                    const std::vector< double >& dirbcvalp );

inline void noslipbc( Fields& U, const std::vector< std::size_t >& nodes, std::size_t pos ) {  // :138-150
  for (auto p : nodes) U(p,pos+0) = U(p,pos+1) = U(p,pos+2) = 0.0;
}

inline void symbc( Fields& U, const std::vector< std::size_t >& nodes,
                   const std::vector< real >& norms, std::size_t pos )       // :110-136
{
  for (std::size_t i=0; i<nodes.size(); ++i) {
    auto p = nodes[i];
    auto n = norms.data() + i*3;
    auto& u = U(p,pos+0);
    auto& v = U(p,pos+1);
    auto& w = U(p,pos+2);
    auto vn = u*n[0] + v*n[1] + w*n[2];
    u -= vn * n[0];
    v -= vn * n[1];
    w -= vn * n[2];
  }
}

inline void farbc( Fields& U, const std::vector< std::size_t >& nodes,
                   const std::vector< real >& norms )                        // :152-220
{
  if (cfg().bc_far.empty()) return;
  real fr = cfg().far_density;
  real fu = cfg().far_velocity[0], fv = cfg().far_velocity[1], fw = cfg().far_velocity[2];
  real fp = cfg().far_pressure;
  for (std::size_t i=0; i<nodes.size(); ++i) {
    auto p = nodes[i];
    auto nx = norms[i*3+0], ny = norms[i*3+1], nz = norms[i*3+2];
    auto& r = U(p,0); auto& ru = U(p,1); auto& rv = U(p,2); auto& rw = U(p,3); auto& re = U(p,4);
    auto vn = fu*nx + fv*ny + fw*nz;
    auto a = eos_soundspeed( fr, fp );
    auto M = vn / a;
    if (M <= -1.0) {
      r = fr; ru = fr*fu; rv = fr*fv; rw = fr*fw;
      re = eos_totalenergy( fr, fu, fv, fw, fp );
    } else if (M > -1.0 && M < 0.0) {
      auto pr = eos_pressure( re - 0.5*(ru*ru + rv*rv + rw*rw)/r );
      r = fr; ru = fr*fu; rv = fr*fv; rw = fr*fw;
      re = eos_totalenergy( fr, fu, fv, fw, pr );
    } else if (M >= 0.0 && M < 1.0) {
      re = eos_totalenergy( r, ru/r, rv/r, rw/r, fp );
    }
  }
}

inline void prebc( Fields& U, const std::vector< std::size_t >& nodes,
                   const std::vector< real >& vals )                         // :222-241
{
  for (std::size_t i=0; i<nodes.size(); ++i) {
    auto p = nodes[i];
    U(p,0) = vals[i*2+0];
    U(p,4) = eos_totalenergy( U(p,0), U(p,1)/U(p,0), U(p,2)/U(p,0),
                              U(p,3)/U(p,0), vals[i*2+1] );
  }
}

// ---- Riemann.cpp -----------------------------------------------------------------
static const real muscl_eps = 1.0e-9;
static const real muscl_const = 1.0/3.0;

inline void primitive( std::size_t ncomp, std::size_t i, const Fields& U, real u[] ) // :211-227
{
  u[0] = U(i,0);
  u[1] = U(i,1) / u[0];
  u[2] = U(i,2) / u[0];
  u[3] = U(i,3) / u[0];
  u[4] = U(i,4) / u[0] - 0.5*(u[1]*u[1] + u[2]*u[2] + u[3]*u[3]);
  for (std::size_t c=5; c<ncomp; ++c) u[c] = U(i,c);
}

//! van Leer-limited MUSCL reconstruction of component range [c0,c1)      // :34-143,:145-209
inline void muscl_range( std::size_t p, std::size_t q, const Coords& coord, const Fields& G,
                         std::size_t c0, std::size_t c1, real l[], real r[],
                         real d1[], real d3[] )
{
  real vw[3] = { coord[0][q]-coord[0][p], coord[1][q]-coord[1][p], coord[2][q]-coord[2][p] };
  for (std::size_t c=c0; c<c1; ++c) {
    auto g1 = G(p,c*3+0)*vw[0] + G(p,c*3+1)*vw[1] + G(p,c*3+2)*vw[2];
    auto g2 = G(q,c*3+0)*vw[0] + G(q,c*3+1)*vw[1] + G(q,c*3+2)*vw[2];
    real delta2 = r[c] - l[c];
    real delta1 = 2.0 * g1 - delta2;
    real delta3 = 2.0 * g2 - delta2;
    auto rcL = (delta2 + muscl_eps) / (delta1 + muscl_eps);
    auto rcR = (delta2 + muscl_eps) / (delta3 + muscl_eps);
    auto rLinv = (delta1 + muscl_eps) / (delta2 + muscl_eps);
    auto rRinv = (delta3 + muscl_eps) / (delta2 + muscl_eps);
    auto phiL = (std::abs(rcL) + rcL) / (std::abs(rcL) + 1.0);
    auto phiR = (std::abs(rcR) + rcR) / (std::abs(rcR) + 1.0);
    auto phi_L_inv = (std::abs(rLinv) + rLinv) / (std::abs(rLinv) + 1.0);
    auto phi_R_inv = (std::abs(rRinv) + rRinv) / (std::abs(rRinv) + 1.0);
    l[c] += 0.25*(delta1*(1.0-muscl_const)*phiL + delta2*(1.0+muscl_const)*phi_L_inv);
    r[c] -= 0.25*(delta3*(1.0-muscl_const)*phiR + delta2*(1.0+muscl_const)*phi_R_inv);
    d1[c] = delta1; d3[c] = delta3;
  }
}

//! MUSCL for the five flow variables incl. first-order fallback            // :34-143
inline void muscl_flow( std::size_t p, std::size_t q, const Coords& coord, const Fields& G,
                        real l[], real r[] )
{
  real ls[5], rs[5], d1[5], d3[5];
  std::memcpy( ls, l, sizeof ls );
  std::memcpy( rs, r, sizeof rs );
  muscl_range( p, q, coord, G, 0, 5, l, r, d1, d3 );
  if (ls[0] < d1[0] || ls[4] < d1[4]) std::memcpy( l, ls, sizeof ls );
  if (rs[0] < -d3[0] || rs[4] < -d3[4]) std::memcpy( r, rs, sizeof rs );
}

using FluxFn = void(*)( const Coords&, const Fields&, const real[], std::size_t, std::size_t,
                        const real[], const real[], real[] );

inline void rusanov( const Coords& coord, const Fields& G, const real dsupint[],
                     std::size_t p, std::size_t q, const real L[], const real R[], real f[] ) // :369-478
{
  auto ncomp = G.nprop() / 3;
  std::vector< real > lv( L, L+ncomp ), rv( R, R+ncomp );
  real* l = lv.data(); real* r = rv.data();
  muscl_flow( p, q, coord, G, l, r );
  auto pL = eos_pressure( l[0]*l[4] );
  auto pR = eos_pressure( r[0]*r[4] );
  auto nx = dsupint[0], ny = dsupint[1], nz = dsupint[2];
  auto vnL = l[1]*nx + l[2]*ny + l[3]*nz;       // symL = symR = 0 always (:707-753)
  auto vnR = r[1]*nx + r[2]*ny + r[3]*nz;
  l[4] = (l[4] + 0.5*(l[1]*l[1] + l[2]*l[2] + l[3]*l[3])) * l[0];
  l[1] *= l[0]; l[2] *= l[0]; l[3] *= l[0];
  r[4] = (r[4] + 0.5*(r[1]*r[1] + r[2]*r[2] + r[3]*r[3])) * r[0];
  r[1] *= r[0]; r[2] *= r[0]; r[3] *= r[0];
  auto len = std::sqrt( nx*nx + ny*ny + nz*nz );
  auto sl = std::abs(vnL) + eos_soundspeed(l[0],pL)*len;
  auto sr = std::abs(vnR) + eos_soundspeed(r[0],pR)*len;
  auto fw = std::max( sl, sr );
  f[0] = l[0]*vnL + r[0]*vnR + fw*(r[0] - l[0]);
  f[1] = l[1]*vnL + r[1]*vnR + (pL + pR)*nx + fw*(r[1] - l[1]);
  f[2] = l[2]*vnL + r[2]*vnR + (pL + pR)*ny + fw*(r[2] - l[2]);
  f[3] = l[3]*vnL + r[3]*vnR + (pL + pR)*nz + fw*(r[3] - l[3]);
  f[4] = (l[4] + pL)*vnL + (r[4] + pR)*vnR + fw*(r[4] - l[4]);
  if (cfg().stab2) {
    auto fws = cfg().stab2coef * fw;
    f[0] -= fws*(l[0] - r[0]);
    f[1] -= fws*(l[1] - r[1]);
    f[2] -= fws*(l[2] - r[2]);
    f[3] -= fws*(l[3] - r[3]);
    f[4] -= fws*(l[4] - r[4]);
  }
  if (ncomp == 5) return;
  std::vector< real > d1( ncomp ), d3( ncomp );
  muscl_range( p, q, coord, G, 5, ncomp, l, r, d1.data(), d3.data() );
  auto sw = std::max( std::abs(vnL), std::abs(vnR) );
  for (std::size_t c=5; c<ncomp; ++c) f[c] = l[c]*vnL + r[c]*vnR + sw*(r[c] - l[c]);
}

inline void hllc( const Coords& coord, const Fields& G, const real dsupint[],
                  std::size_t p, std::size_t q, const real L[], const real R[], real f[] )   // :480-650
{
  auto ncomp = G.nprop() / 3;
  std::vector< real > lv( L, L+ncomp ), rv( R, R+ncomp );
  real* l = lv.data(); real* r = rv.data();
  muscl_flow( p, q, coord, G, l, r );
  auto nx = -dsupint[0], ny = -dsupint[1], nz = -dsupint[2];
  auto len = std::sqrt( nx*nx + ny*ny + nz*nz );
  nx /= len; ny /= len; nz /= len;
  auto qL = l[1]*nx + l[2]*ny + l[3]*nz;
  auto qR = r[1]*nx + r[2]*ny + r[3]*nz;
  auto pL = eos_pressure( l[0]*l[4] );
  auto pR = eos_pressure( r[0]*r[4] );
  l[4] = (l[4] + 0.5*(l[1]*l[1] + l[2]*l[2] + l[3]*l[3])) * l[0];
  l[1] *= l[0]; l[2] *= l[0]; l[3] *= l[0];
  r[4] = (r[4] + 0.5*(r[1]*r[1] + r[2]*r[2] + r[3]*r[3])) * r[0];
  r[1] *= r[0]; r[2] *= r[0]; r[3] *= r[0];
  auto cL = eos_soundspeed(l[0],pL);
  auto cR = eos_soundspeed(r[0],pR);
  auto sL = std::fmin( qL - cL, qR - cR );
  auto sR = std::fmax( qL + cL, qR + cR );
  auto tL = sL - qL;
  auto tR = sR - qR;
  auto sM = (r[0]*qR*tR - l[0]*qL*tL + pL - pR) / (r[0]*tR - l[0]*tL);
  auto pS = pL - l[0]*tL*(qL - sM);
  real uL[5], uR[5];
  auto s = sL - sM;
  uL[0] = tL*l[0]/s;
  uL[1] = (tL*l[1] + (pS-pL)*nx)/s;
  uL[2] = (tL*l[2] + (pS-pL)*ny)/s;
  uL[3] = (tL*l[3] + (pS-pL)*nz)/s;
  uL[4] = (tL*l[4] - pL*qL + pS*sM)/s;
  s = sR - sM;
  uR[0] = tR*r[0]/s;
  uR[1] = (tR*r[1] + (pS-pR)*nx)/s;
  uR[2] = (tR*r[2] + (pS-pR)*ny)/s;
  uR[3] = (tR*r[3] + (pS-pR)*nz)/s;
  uR[4] = (tR*r[4] - pR*qR + pS*sM)/s;
  auto L2 = -2.0*len;
  nx *= L2; ny *= L2; nz *= L2;
  if (sL > 0.0) {
    auto qL2 = qL * L2;
    f[0] = l[0]*qL2;
    f[1] = l[1]*qL2 + pL*nx;
    f[2] = l[2]*qL2 + pL*ny;
    f[3] = l[3]*qL2 + pL*nz;
    f[4] = (l[4] + pL)*qL2;
  } else if (sL <= 0.0 && sM > 0.0) {
    auto qL2 = qL * L2;
    auto sL2 = sL * L2;
    f[0] = l[0]*qL2 + sL2*(uL[0] - l[0]);
    f[1] = l[1]*qL2 + pL*nx + sL2*(uL[1] - l[1]);
    f[2] = l[2]*qL2 + pL*ny + sL2*(uL[2] - l[2]);
    f[3] = l[3]*qL2 + pL*nz + sL2*(uL[3] - l[3]);
    f[4] = (l[4] + pL)*qL2 + sL2*(uL[4] - l[4]);
  } else if (sM <= 0.0 && sR >= 0.0) {
    auto qR2 = qR * L2;
    auto sR2 = sR * L2;
    f[0] = r[0]*qR2 + sR2*(uR[0] - r[0]);
    f[1] = r[1]*qR2 + pR*nx + sR2*(uR[1] - r[1]);
    f[2] = r[2]*qR2 + pR*ny + sR2*(uR[2] - r[2]);
    f[3] = r[3]*qR2 + pR*nz + sR2*(uR[3] - r[3]);
    f[4] = (r[4] + pR)*qR2 + sR2*(uR[4] - r[4]);
  } else {
    auto qR2 = qR * L2;
    f[0] = r[0]*qR2;
    f[1] = r[1]*qR2 + pR*nx;
    f[2] = r[2]*qR2 + pR*ny;
    f[3] = r[3]*qR2 + pR*nz;
    f[4] = (r[4] + pR)*qR2;
  }
  if (cfg().stab2) {
    auto sl = std::abs(qL) + cL;
    auto sr = std::abs(qR) + cR;
    auto fws = cfg().stab2coef * std::max(sl,sr) * len;
    f[0] -= fws * (l[0] - r[0]);
    f[1] -= fws * (l[1] - r[1]);
    f[2] -= fws * (l[2] - r[2]);
    f[3] -= fws * (l[3] - r[3]);
    f[4] -= fws * (l[4] - r[4]);
  }
  if (ncomp == 5) return;
  std::vector< real > d1( ncomp ), d3( ncomp );
  muscl_range( p, q, coord, G, 5, ncomp, l, r, d1.data(), d3.data() );
  auto sw = std::max( std::abs(qL), std::abs(qR) ) * len;
  for (std::size_t c=5; c<ncomp; ++c) f[c] = (l[c]*qL + r[c]*qR)*len + sw*(r[c] - l[c]);
}

//! Nodal gradients of primitive variables, weak (un-normalised) form       // :229-367
//! loader of the nodal variables the edge loops work on
struct LoadPrimitive { void operator()( std::size_t ncomp, std::size_t i, const Fields& U, real u[] ) const { primitive( ncomp, i, U, u ); } };
struct LoadAsIs { void operator()( std::size_t ncomp, std::size_t i, const Fields& U, real u[] ) const { for (std::size_t c=0; c<ncomp; ++c) u[c] = U(i,c); } };

template< class Load >
inline void grad_impl( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                  const std::array< std::vector< real >, 3 >& dsupint,
                  const Coords& coord,
                  const std::vector< std::size_t >& triinpoel,
                  const Fields& U, Fields& G, Load primitive )
{
  auto ncomp = U.nprop();
  G.fill( 0.0 );
  std::vector< real > ub( 4*ncomp );
  real* u[4] = { ub.data(), ub.data()+ncomp, ub.data()+2*ncomp, ub.data()+3*ncomp };

  for (std::size_t e=0; e<dsupedge[0].size()/4; ++e) {                      // tets :266-289
    const auto N = dsupedge[0].data() + e*4;
    for (int k=0; k<4; ++k) primitive( ncomp, N[k], U, u[k] );
    const auto d = dsupint[0].data();
    for (std::size_t c=0; c<ncomp; ++c)
      for (std::size_t j=0; j<3; ++j) {
        real f[6];
        f[0] = d[(e*6+0)*3+j] * (u[1][c] + u[0][c]);
        f[1] = d[(e*6+1)*3+j] * (u[2][c] + u[1][c]);
        f[2] = d[(e*6+2)*3+j] * (u[0][c] + u[2][c]);
        f[3] = d[(e*6+3)*3+j] * (u[3][c] + u[0][c]);
        f[4] = d[(e*6+4)*3+j] * (u[3][c] + u[1][c]);
        f[5] = d[(e*6+5)*3+j] * (u[3][c] + u[2][c]);
        G(N[0],c*3+j) = G(N[0],c*3+j) - f[0] + f[2] - f[3];
        G(N[1],c*3+j) = G(N[1],c*3+j) + f[0] - f[1] - f[4];
        G(N[2],c*3+j) = G(N[2],c*3+j) + f[1] - f[2] - f[5];
        G(N[3],c*3+j) = G(N[3],c*3+j) + f[3] + f[4] + f[5];
      }
  }
  for (std::size_t e=0; e<dsupedge[1].size()/3; ++e) {                      // triangles :291-310
    const auto N = dsupedge[1].data() + e*3;
    for (int k=0; k<3; ++k) primitive( ncomp, N[k], U, u[k] );
    const auto d = dsupint[1].data();
    for (std::size_t c=0; c<ncomp; ++c)
      for (std::size_t j=0; j<3; ++j) {
        real f[3];
        f[0] = d[(e*3+0)*3+j] * (u[1][c] + u[0][c]);
        f[1] = d[(e*3+1)*3+j] * (u[2][c] + u[1][c]);
        f[2] = d[(e*3+2)*3+j] * (u[0][c] + u[2][c]);
        G(N[0],c*3+j) = G(N[0],c*3+j) - f[0] + f[2];
        G(N[1],c*3+j) = G(N[1],c*3+j) + f[0] - f[1];
        G(N[2],c*3+j) = G(N[2],c*3+j) + f[1] - f[2];
      }
  }
  for (std::size_t e=0; e<dsupedge[2].size()/2; ++e) {                      // edges :312-326
    const auto N = dsupedge[2].data() + e*2;
    for (int k=0; k<2; ++k) primitive( ncomp, N[k], U, u[k] );
    const auto d = dsupint[2].data() + e*3;
    for (std::size_t c=0; c<ncomp; ++c)
      for (std::size_t j=0; j<3; ++j) {
        real f = d[j] * (u[1][c] + u[0][c]);
        G(N[0],c*3+j) -= f;
        G(N[1],c*3+j) += f;
      }
  }
  const auto& x = coord[0]; const auto& y = coord[1]; const auto& z = coord[2];
  for (std::size_t e=0; e<triinpoel.size()/3; ++e) {                        // boundary :334-360
    const auto N = triinpoel.data() + e*3;
    for (int k=0; k<3; ++k) primitive( ncomp, N[k], U, u[k] );
    real ba[3] = { x[N[1]]-x[N[0]], y[N[1]]-y[N[0]], z[N[1]]-z[N[0]] },
         ca[3] = { x[N[2]]-x[N[0]], y[N[2]]-y[N[0]], z[N[2]]-z[N[0]] };
    real n[3] = { ba[1]*ca[2] - ca[1]*ba[2], ba[2]*ca[0] - ca[2]*ba[0], ba[0]*ca[1] - ca[0]*ba[1] };
    n[0] /= 12.0; n[1] /= 12.0; n[2] /= 12.0;
    for (std::size_t c=0; c<ncomp; ++c) {
      auto uab = (u[0][c] + u[1][c])/4.0;
      auto ubc = (u[1][c] + u[2][c])/4.0;
      auto uca = (u[2][c] + u[0][c])/4.0;
      real g[] = { uab + uca + u[0][c], uab + ubc + u[1][c], ubc + uca + u[2][c] };
      for (std::size_t j=0; j<3; ++j) {      // g indexed by direction j: reference quirk :351-357
        G(N[0],c*3+j) += g[j] * n[j];
        G(N[1],c*3+j) += g[j] * n[j];
        G(N[2],c*3+j) += g[j] * n[j];
      }
    }
  }
}

//! riemann::grad, Riemann.cpp:229-367
inline void grad( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                  const std::array< std::vector< real >, 3 >& dsupint,
                  const Coords& coord,
                  const std::vector< std::size_t >& triinpoel,
                  const Fields& U, Fields& G )
{ grad_impl( dsupedge, dsupint, coord, triinpoel, U, G, LoadPrimitive() ); }

template< class Load >
inline void advdom_impl( const Coords& coord,
                    const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                    const std::array< std::vector< real >, 3 >& dsupint,
                    const Fields& G, const Fields& U, Fields& R, FluxFn flux, Load primitive )
{
  auto ncomp = U.nprop();
  std::vector< real > ub( 4*ncomp ), fb( 6*ncomp );
  real* u[4] = { ub.data(), ub.data()+ncomp, ub.data()+2*ncomp, ub.data()+3*ncomp };
  real* f[6]; for (int k=0; k<6; ++k) f[k] = fb.data() + static_cast<std::size_t>(k)*ncomp;

  for (std::size_t e=0; e<dsupedge[0].size()/4; ++e) {
    const auto N = dsupedge[0].data() + e*4;
    for (int k=0; k<4; ++k) primitive( ncomp, N[k], U, u[k] );
    const auto d = dsupint[0].data();
    flux( coord, G, d+(e*6+0)*3, N[0], N[1], u[0], u[1], f[0] );
    flux( coord, G, d+(e*6+1)*3, N[1], N[2], u[1], u[2], f[1] );
    flux( coord, G, d+(e*6+2)*3, N[2], N[0], u[2], u[0], f[2] );
    flux( coord, G, d+(e*6+3)*3, N[0], N[3], u[0], u[3], f[3] );
    flux( coord, G, d+(e*6+4)*3, N[1], N[3], u[1], u[3], f[4] );
    flux( coord, G, d+(e*6+5)*3, N[2], N[3], u[2], u[3], f[5] );
    for (std::size_t c=0; c<ncomp; ++c) {
      R(N[0],c) = R(N[0],c) - f[0][c] + f[2][c] - f[3][c];
      R(N[1],c) = R(N[1],c) + f[0][c] - f[1][c] - f[4][c];
      R(N[2],c) = R(N[2],c) + f[1][c] - f[2][c] - f[5][c];
      R(N[3],c) = R(N[3],c) + f[3][c] + f[4][c] + f[5][c];
    }
  }
  for (std::size_t e=0; e<dsupedge[1].size()/3; ++e) {
    const auto N = dsupedge[1].data() + e*3;
    for (int k=0; k<3; ++k) primitive( ncomp, N[k], U, u[k] );
    const auto d = dsupint[1].data();
    flux( coord, G, d+(e*3+0)*3, N[0], N[1], u[0], u[1], f[0] );
    flux( coord, G, d+(e*3+1)*3, N[1], N[2], u[1], u[2], f[1] );
    flux( coord, G, d+(e*3+2)*3, N[2], N[0], u[2], u[0], f[2] );
    for (std::size_t c=0; c<ncomp; ++c) {
      R(N[0],c) = R(N[0],c) - f[0][c] + f[2][c];
      R(N[1],c) = R(N[1],c) + f[0][c] - f[1][c];
      R(N[2],c) = R(N[2],c) + f[1][c] - f[2][c];
    }
  }
  for (std::size_t e=0; e<dsupedge[2].size()/2; ++e) {
    const auto N = dsupedge[2].data() + e*2;
    for (int k=0; k<2; ++k) primitive( ncomp, N[k], U, u[k] );
    const auto d = dsupint[2].data();
    flux( coord, G, d+e*3, N[0], N[1], u[0], u[1], f[0] );
    for (std::size_t c=0; c<ncomp; ++c) {
      R(N[0],c) -= f[0][c];
      R(N[1],c) += f[0][c];
    }
  }
}

inline void advdom( const Coords& coord,
                    const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                    const std::array< std::vector< real >, 3 >& dsupint,
                    const Fields& G, const Fields& U, Fields& R )             // :652-766
{
  FluxFn flux;
  if (cfg().flux == "rusanov") flux = rusanov;
  else if (cfg().flux == "hllc") flux = hllc;
  else throw std::runtime_error( "oracle port: Flux not configured" );
  advdom_impl( coord, dsupedge, dsupint, G, U, R, flux, LoadPrimitive() );
}

inline void advbnd( const std::vector< std::size_t >& triinpoel, const Coords& coord,
                    const std::vector< std::uint8_t >& besym, const Fields& U, Fields& R ) // :768-878
{
  auto ncomp = U.nprop();
  const auto& x = coord[0]; const auto& y = coord[1]; const auto& z = coord[2];
  std::vector< real > fb( ncomp*3 );
  auto f = [&]( std::size_t c, std::size_t k ) -> real& { return fb[c*3+k]; };
  for (std::size_t e=0; e<triinpoel.size()/3; ++e) {
    const auto N = triinpoel.data() + e*3;
    real ba[3] = { x[N[1]]-x[N[0]], y[N[1]]-y[N[0]], z[N[1]]-z[N[0]] },
         ca[3] = { x[N[2]]-x[N[0]], y[N[2]]-y[N[0]], z[N[2]]-z[N[0]] };
    real nx = ba[1]*ca[2] - ca[1]*ba[2], ny = ba[2]*ca[0] - ca[2]*ba[0], nz = ba[0]*ca[1] - ca[0]*ba[1];
    nx /= 12.0; ny /= 12.0; nz /= 12.0;
    const auto sym = besym.data() + e*3;
    for (std::size_t k=0; k<3; ++k) {
      auto r = U(N[k],0), ru = U(N[k],1), rv = U(N[k],2), rw = U(N[k],3), re = U(N[k],4);
      real p = eos_pressure( re - 0.5*(ru*ru + rv*rv + rw*rw)/r );
      real vn = sym[k] ? 0.0 : (nx*ru + ny*rv + nz*rw)/r;
      f(0,k) = r*vn;
      f(1,k) = ru*vn + p*nx;
      f(2,k) = rv*vn + p*ny;
      f(3,k) = rw*vn + p*nz;
      f(4,k) = (re + p)*vn;
      for (std::size_t c=5; c<ncomp; ++c) f(c,k) = U(N[k],c)*vn;
    }
    for (std::size_t c=0; c<ncomp; ++c) {
      auto fab = (f(c,0) + f(c,1))/4.0;
      auto fbc = (f(c,1) + f(c,2))/4.0;
      auto fca = (f(c,2) + f(c,0))/4.0;
      R(N[0],c) += fab + fca + f(c,0);
      R(N[1],c) += fab + fbc + f(c,1);
      R(N[2],c) += fbc + fca + f(c,2);
    }
  }
}

inline void src( const Coords& coord, const std::vector< real >& v, real t,
                 const std::vector< real >& tp, Fields& R )                    // :880-907
{
  auto s_ = SRC();
  if (!s_) return;
  for (std::size_t p=0; p<R.nunk(); ++p) {
    if (cfg().steady) t = tp[p];
    auto s = s_( coord[0][p], coord[1][p], coord[2][p], t );
    for (std::size_t c=0; c<s.size(); ++c) R(p,c) -= s[c] * v[p];
  }
}

inline void rhs( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                 const std::array< std::vector< real >, 3 >& dsupint,
                 const Coords& coord,
                 const std::vector< std::size_t >& triinpoel,
                 const std::vector< std::uint8_t >& besym,
                 const Fields& G, const Fields& U, const std::vector< real >& v,
                 real t, const std::vector< real >& tp, Fields& R )            // :909-946
{
  R.fill( 0.0 );
  advdom( coord, dsupedge, dsupint, G, U, R );
  advbnd( triinpoel, coord, besym, U, R );
  src( coord, v, t, tp, R );
}

// ---- Lax.cpp: time-derivative preconditioned edge fluxes for LaxCG ------------------------
// nodal unknowns are (p,u,v,w,T) throughout (LaxCG::primitive, LaxCG.cpp:115-137)

//! lax::grad, Lax.cpp:216-343: same edge/boundary formulas, applied to the unknowns as they are
inline void lax_grad( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                      const std::array< std::vector< real >, 3 >& dsupint,
                      const Coords& coord, const std::vector< std::size_t >& triinpoel,
                      const Fields& U, Fields& G )
{ grad_impl( dsupedge, dsupint, coord, triinpoel, U, G, LoadAsIs() ); }

inline real lax_refvel( real r, real p, real v ) {                          // :344-360
  const auto& vi = cfg().velinf;
  auto vinf = std::sqrt( vi[0]*vi[0] + vi[1]*vi[1] + vi[2]*vi[2] );
  return std::min( eos_soundspeed( r, p ), std::max( v, cfg().turkel*vinf ) );
}

inline void lax_sigvel( real p, real T, real v, real vn, real& vpri, real& cpri ) {  // :362-388
  auto g = cfg().gamma, rgas = cfg().rgas;
  auto cp = g*rgas/(g-1.0);
  auto r = p/T/rgas;
  auto rp = r/p;
  auto rt = -r/T;
  auto vr = lax_refvel( r, p, v );
  auto vr2 = vr*vr;
  auto beta = rp + rt/r/cp;
  auto alpha = 0.5*(1.0 - beta*vr2);
  vpri = vn*(1.0 - alpha);
  cpri = std::sqrt( alpha*alpha*vn*vn + vr2 );
}

inline real length3( real a, real b, real c ) { return std::sqrt( a*a + b*b + c*c ); }

//! (p,u,v,w,T) -> (r,ru,rv,rw,rE) of an edge-end state, :438-451
inline void lax_conserved( real l[], real pL ) {
  auto g = cfg().gamma, rgas = cfg().rgas;
  l[0] = pL/l[4]/rgas;
  l[1] *= l[0]; l[2] *= l[0]; l[3] *= l[0];
  l[4] = pL/(g-1.0) + 0.5*(l[1]*l[1] + l[2]*l[2] + l[3]*l[3])/l[0];
}

inline void lax_rusanov( const Coords& coord, const Fields& G, const real dsupint[],
                         std::size_t p, std::size_t q, const real L[], const real R[], real f[] ) // :390-511
{
  auto ncomp = G.nprop() / 3;
  std::vector< real > lv( L, L+ncomp ), rv( R, R+ncomp );
  real* l = lv.data(); real* r = rv.data();
  muscl_flow( p, q, coord, G, l, r );        // same limiter, fallback tests pressure and temperature
  auto nx = dsupint[0], ny = dsupint[1], nz = dsupint[2];
  auto vnL = l[1]*nx + l[2]*ny + l[3]*nz;
  auto vnR = r[1]*nx + r[2]*ny + r[3]*nz;
  auto pL = l[0], pR = r[0];
  auto len = length3( nx, ny, nz );
  real vpL, cpL, vpR, cpR;
  lax_sigvel( l[0], l[4], length3(l[1],l[2],l[3]), vnL, vpL, cpL );
  lax_sigvel( r[0], r[4], length3(r[1],r[2],r[3]), vnR, vpR, cpR );
  lax_conserved( l, pL );
  lax_conserved( r, pR );
  using std::abs; using std::max;
  auto sp = max(abs(vpL-cpL),max(abs(vpR-cpR),max(abs(vpL+cpL),abs(vpR+cpR))));
  auto sL = -sp, sR = +sp;
  auto fw = std::max( sL, sR ) * len;
  f[0] = l[0]*vnL + r[0]*vnR + fw*(r[0] - l[0]);
  f[1] = l[1]*vnL + r[1]*vnR + (pL + pR)*nx + fw*(r[1] - l[1]);
  f[2] = l[2]*vnL + r[2]*vnR + (pL + pR)*ny + fw*(r[2] - l[2]);
  f[3] = l[3]*vnL + r[3]*vnR + (pL + pR)*nz + fw*(r[3] - l[3]);
  f[4] = (l[4] + pL)*vnL + (r[4] + pR)*vnR + fw*(r[4] - l[4]);
  if (cfg().stab2) {
    auto fws = cfg().stab2coef * fw;
    for (int c=0; c<5; ++c) f[c] -= fws*(l[c] - r[c]);
  }
  if (ncomp == 5) return;
  std::vector< real > d1( ncomp ), d3( ncomp );
  muscl_range( p, q, coord, G, 5, ncomp, l, r, d1.data(), d3.data() );
  auto sw = std::max( std::abs(vnL), std::abs(vnR) );
  for (std::size_t c=5; c<ncomp; ++c) f[c] = l[c]*vnL + r[c]*vnR + sw*(r[c] - l[c]);
}

inline void lax_hllc( const Coords& coord, const Fields& G, const real dsupint[],
                      std::size_t p, std::size_t q, const real L[], const real R[], real f[] )   // :513-723
{
  auto ncomp = G.nprop() / 3;
  std::vector< real > lv( L, L+ncomp ), rv( R, R+ncomp );
  real* l = lv.data(); real* r = rv.data();
  muscl_flow( p, q, coord, G, l, r );
  auto nx = -dsupint[0], ny = -dsupint[1], nz = -dsupint[2];
  auto len = length3( nx, ny, nz );
  nx /= len; ny /= len; nz /= len;
  auto qL = l[1]*nx + l[2]*ny + l[3]*nz;
  auto qR = r[1]*nx + r[2]*ny + r[3]*nz;
  auto pL = l[0], pR = r[0];
  real vpL, cpL, vpR, cpR;
  lax_sigvel( l[0], l[4], length3(l[1],l[2],l[3]), qL*len, vpL, cpL );
  lax_sigvel( r[0], r[4], length3(r[1],r[2],r[3]), qR*len, vpR, cpR );
  lax_conserved( l, pL );
  lax_conserved( r, pR );
  using std::abs; using std::max;
  auto sp = max(abs(vpL-cpL),max(abs(vpR-cpR),max(abs(vpL+cpL),abs(vpR+cpR))));
  auto sL = -sp, sR = +sp;
  auto tL = sL - qL;
  auto tR = sR - qR;
  auto sM = (r[0]*qR*tR - l[0]*qL*tL + pL - pR) / (r[0]*tR - l[0]*tL);
  auto pS = pL - l[0]*tL*(qL - sM);
  real uL[5], uR[5];
  auto s = sL - sM;
  uL[0] = tL*l[0]/s;
  uL[1] = (tL*l[1] + (pS-pL)*nx)/s;
  uL[2] = (tL*l[2] + (pS-pL)*ny)/s;
  uL[3] = (tL*l[3] + (pS-pL)*nz)/s;
  uL[4] = (tL*l[4] - pL*qL + pS*sM)/s;
  s = sR - sM;
  uR[0] = tR*r[0]/s;
  uR[1] = (tR*r[1] + (pS-pR)*nx)/s;
  uR[2] = (tR*r[2] + (pS-pR)*ny)/s;
  uR[3] = (tR*r[3] + (pS-pR)*nz)/s;
  uR[4] = (tR*r[4] - pR*qR + pS*sM)/s;
  auto L2 = -2.0*len;
  nx *= L2; ny *= L2; nz *= L2;
  if (sL > 0.0) {
    auto qL2 = qL * L2;
    f[0] = l[0]*qL2;
    f[1] = l[1]*qL2 + pL*nx;
    f[2] = l[2]*qL2 + pL*ny;
    f[3] = l[3]*qL2 + pL*nz;
    f[4] = (l[4] + pL)*qL2;
  }
  else if (sL <= 0.0 && sM > 0.0) {
    auto qL2 = qL * L2;
    auto sL2 = sL * L2;
    f[0] = l[0]*qL2 + sL2*(uL[0] - l[0]);
    f[1] = l[1]*qL2 + pL*nx + sL2*(uL[1] - l[1]);
    f[2] = l[2]*qL2 + pL*ny + sL2*(uL[2] - l[2]);
    f[3] = l[3]*qL2 + pL*nz + sL2*(uL[3] - l[3]);
    f[4] = (l[4] + pL)*qL2 + sL2*(uL[4] - l[4]);
  }
  else if (sM <= 0.0 && sR >= 0.0) {
    auto qR2 = qR * L2;
    auto sR2 = sR * L2;
    f[0] = r[0]*qR2 + sR2*(uR[0] - r[0]);
    f[1] = r[1]*qR2 + pR*nx + sR2*(uR[1] - r[1]);
    f[2] = r[2]*qR2 + pR*ny + sR2*(uR[2] - r[2]);
    f[3] = r[3]*qR2 + pR*nz + sR2*(uR[3] - r[3]);
    f[4] = (r[4] + pR)*qR2 + sR2*(uR[4] - r[4]);
  }
  else {
    auto qR2 = qR * L2;
    f[0] = r[0]*qR2;
    f[1] = r[1]*qR2 + pR*nx;
    f[2] = r[2]*qR2 + pR*ny;
    f[3] = r[3]*qR2 + pR*nz;
    f[4] = (r[4] + pR)*qR2;
  }
  if (ncomp == 5) return;
  // the reference reconstructs the scalars but leaves their hllc fluxes unset (:707-717)
  std::vector< real > d1( ncomp ), d3( ncomp );
  muscl_range( p, q, coord, G, 5, ncomp, l, r, d1.data(), d3.data() );
}

inline void lax_advbnd( const std::vector< std::size_t >& triinpoel, const Coords& coord,
                        const std::vector< std::uint8_t >& besym, const Fields& U, Fields& R ) // :840-950
{
  auto ncomp = U.nprop();
  auto g = cfg().gamma, rgas = cfg().rgas;
  const auto& x = coord[0]; const auto& y = coord[1]; const auto& z = coord[2];
  std::vector< real > fb( ncomp*3 );
  auto f = [&]( std::size_t c, std::size_t k ) -> real& { return fb[c*3+k]; };
  for (std::size_t e=0; e<triinpoel.size()/3; ++e) {
    const auto N = triinpoel.data() + e*3;
    real ba[3] = { x[N[1]]-x[N[0]], y[N[1]]-y[N[0]], z[N[1]]-z[N[0]] },
         ca[3] = { x[N[2]]-x[N[0]], y[N[2]]-y[N[0]], z[N[2]]-z[N[0]] };
    real nx = ba[1]*ca[2] - ca[1]*ba[2], ny = ba[2]*ca[0] - ca[2]*ba[0], nz = ba[0]*ca[1] - ca[0]*ba[1];
    nx /= 12.0; ny /= 12.0; nz /= 12.0;
    const auto sym = besym.data() + e*3;
    for (std::size_t k=0; k<3; ++k) {
      auto rA  = U(N[k],0)/U(N[k],4)/rgas;
      auto ruA = U(N[k],1) * rA;
      auto rvA = U(N[k],2) * rA;
      auto rwA = U(N[k],3) * rA;
      auto reA = U(N[k],0)/(g-1.0) + 0.5*(ruA*ruA + rvA*rvA + rwA*rwA)/rA;
      real vn = sym[k] ? 0.0 : (nx*U(N[k],1) + ny*U(N[k],2) + nz*U(N[k],3));
      f(0,k) = rA*vn;
      f(1,k) = ruA*vn + U(N[k],0)*nx;
      f(2,k) = rvA*vn + U(N[k],0)*ny;
      f(3,k) = rwA*vn + U(N[k],0)*nz;
      f(4,k) = (reA + U(N[k],0))*vn;
      for (std::size_t c=5; c<ncomp; ++c) f(c,k) = U(N[k],c)*vn;
    }
    for (std::size_t c=0; c<ncomp; ++c) {
      auto fab = (f(c,0) + f(c,1))/4.0;
      auto fbc = (f(c,1) + f(c,2))/4.0;
      auto fca = (f(c,2) + f(c,0))/4.0;
      R(N[0],c) += fab + fca + f(c,0);
      R(N[1],c) += fab + fbc + f(c,1);
      R(N[2],c) += fbc + fca + f(c,2);
    }
  }
}

//! lax::rhs, Lax.cpp:981-1018
inline void lax_rhs( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                     const std::array< std::vector< real >, 3 >& dsupint,
                     const Coords& coord, const std::vector< std::size_t >& triinpoel,
                     const std::vector< std::uint8_t >& besym,
                     const Fields& G, const Fields& U, const std::vector< real >& v,
                     real t, const std::vector< real >& tp, Fields& R )
{
  R.fill( 0.0 );
  FluxFn flux;
  if (cfg().flux == "rusanov") flux = lax_rusanov;
  else if (cfg().flux == "hllc") flux = lax_hllc;
  else throw std::runtime_error( "oracle port: Flux not configured" );
  advdom_impl( coord, dsupedge, dsupint, G, U, R, flux, LoadAsIs() );
  lax_advbnd( triinpoel, coord, besym, U, R );
  src( coord, v, t, tp, R );
}

// ---- Chorin.cpp: edge operators of the projection solver ChoCG --------------------------
// superedge integrals have stride 5: normal(3), J/120, grad_p.grad_q/(6J) (ChoCG.cpp:399-446)

//! visit all domain edges of the superedge groups in the reference's order and scatter
//! pattern: fn( p, q, integrals ) returns nothing, add( node, sign ) applies +-f
template< class EdgeFn >
inline void chorin_foredge( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                            const std::array< std::vector< real >, 3 >& dsupint, EdgeFn fn )
{
  for (std::size_t e=0; e<dsupedge[0].size()/4; ++e) {
    const auto N = dsupedge[0].data() + e*4;
    const auto d = dsupint[0].data();
    fn( 0, e, N, d );
  }
  for (std::size_t e=0; e<dsupedge[1].size()/3; ++e) {
    const auto N = dsupedge[1].data() + e*3;
    const auto d = dsupint[1].data();
    fn( 1, e, N, d );
  }
  for (std::size_t e=0; e<dsupedge[2].size()/2; ++e) {
    const auto N = dsupedge[2].data() + e*2;
    const auto d = dsupint[2].data();
    fn( 2, e, N, d );
  }
}

inline void crossdiv6( const Coords& coord, const std::size_t N[3], real n[3] ) {
  const auto& x = coord[0]; const auto& y = coord[1]; const auto& z = coord[2];
  real a[3] = { x[N[1]]-x[N[0]], y[N[1]]-y[N[0]], z[N[1]]-z[N[0]] },
       b[3] = { x[N[2]]-x[N[0]], y[N[2]]-y[N[0]], z[N[2]]-z[N[0]] };
  n[0] = (a[1]*b[2] - a[2]*b[1]) / 6.0;      // tk::crossdiv, Vector.hpp
  n[1] = (a[2]*b[0] - a[0]*b[2]) / 6.0;
  n[2] = (a[0]*b[1] - a[1]*b[0]) / 6.0;
}

//! edge divergence with optional pressure stabilisation, Chorin.cpp:34-83
inline real chorin_div_edge( const Coords& coord, const real d[], real dt, const std::vector< real >& P,
                             const Fields& G, const Fields& U, std::size_t p, std::size_t q, bool stab,
                             std::size_t pos = 0 )
{
  real div = d[0] * (U(p,pos+0) + U(q,pos+0)) + d[1] * (U(p,pos+1) + U(q,pos+1)) + d[2] * (U(p,pos+2) + U(q,pos+2));
  if (!stab) return div;
  auto dx = coord[0][p] - coord[0][q];
  auto dy = coord[1][p] - coord[1][q];
  auto dz = coord[2][p] - coord[2][q];
  auto dl = std::sqrt( dx*dx + dy*dy + dz*dz );
  auto p2 = P[q] - P[p];
  auto D = std::sqrt( d[0]*d[0] + d[1]*d[1] + d[2]*d[2] );
  auto dpx = G(p,0) + G(q,0);
  auto dpy = G(p,1) + G(q,1);
  auto dpz = G(p,2) + G(q,2);
  auto p4 = 0.5 * (dx*dpx + dy*dpy + dz*dpz);
  div += D*dt/dl*(p2 + p4);
  return div;
}

//! chorin::div, Chorin.cpp:85-209 (accumulates into D); with S = 4 integrals per edge and the velocity
//! at components pos.. it is lohner::div, Lohner.cpp:35-159
inline void chorin_div( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                        const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                        const std::vector< std::size_t >& triinpoel, real dt, const std::vector< real >& P,
                        const Fields& G, const Fields& U, std::vector< real >& D, bool stab,
                        std::size_t S = 5, std::size_t pos = 0 )
{
  auto ed = [&]( const real* d, std::size_t p, std::size_t q ){ return chorin_div_edge( coord, d, dt, P, G, U, p, q, stab, pos ); };
  chorin_foredge( dsupedge, dsupint, [&]( int kind, std::size_t e, const std::size_t* N, const real* d ){
    if (kind == 0) {
      real f[] = { ed( d+(e*6+0)*S, N[0], N[1] ), ed( d+(e*6+1)*S, N[1], N[2] ), ed( d+(e*6+2)*S, N[2], N[0] ),
                   ed( d+(e*6+3)*S, N[0], N[3] ), ed( d+(e*6+4)*S, N[1], N[3] ), ed( d+(e*6+5)*S, N[2], N[3] ) };
      D[N[0]] = D[N[0]] - f[0] + f[2] - f[3];
      D[N[1]] = D[N[1]] + f[0] - f[1] - f[4];
      D[N[2]] = D[N[2]] + f[1] - f[2] - f[5];
      D[N[3]] = D[N[3]] + f[3] + f[4] + f[5];
    } else if (kind == 1) {
      real f[] = { ed( d+(e*3+0)*S, N[0], N[1] ), ed( d+(e*3+1)*S, N[1], N[2] ), ed( d+(e*3+2)*S, N[2], N[0] ) };
      D[N[0]] = D[N[0]] - f[0] + f[2];
      D[N[1]] = D[N[1]] + f[0] - f[1];
      D[N[2]] = D[N[2]] + f[1] - f[2];
    } else {
      real f = ed( d+e*S, N[0], N[1] );
      D[N[0]] -= f;
      D[N[1]] += f;
    }
  } );
  for (std::size_t e=0; e<triinpoel.size()/3; ++e) {
    const auto N = triinpoel.data() + e*3;
    real n[3]; crossdiv6( coord, N, n );
    auto uxA = U(N[0],pos+0), uyA = U(N[0],pos+1), uzA = U(N[0],pos+2);
    auto uxB = U(N[1],pos+0), uyB = U(N[1],pos+1), uzB = U(N[1],pos+2);
    auto uxC = U(N[2],pos+0), uyC = U(N[2],pos+1), uzC = U(N[2],pos+2);
    auto ux = (6.0*uxA + uxB + uxC)/8.0;
    auto uy = (6.0*uyA + uyB + uyC)/8.0;
    auto uz = (6.0*uzA + uzB + uzC)/8.0;
    D[N[0]] += ux*n[0] + uy*n[1] + uz*n[2];
    ux = (uxA + 6.0*uxB + uxC)/8.0;
    uy = (uyA + 6.0*uyB + uyC)/8.0;
    uz = (uzA + 6.0*uzB + uzC)/8.0;
    D[N[1]] += ux*n[0] + uy*n[1] + uz*n[2];
    ux = (uxA + uxB + 6.0*uxC)/8.0;
    uy = (uyA + uyB + 6.0*uyC)/8.0;
    uz = (uzA + uzB + 6.0*uzC)/8.0;
    D[N[2]] += ux*n[0] + uy*n[1] + uz*n[2];
  }
}

//! gradient of ncomp nodal scalars get(node,i) into G(node,i*3+j): common body of
//! chorin::vgrad (:211-334) and chorin::grad (:336-449); accumulates
template< class Get >
inline void chorin_grad_impl( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                              const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                              const std::vector< std::size_t >& triinpoel, std::size_t ncomp, Get U, Fields& G,
                              std::size_t S = 5 )
{
  chorin_foredge( dsupedge, dsupint, [&]( int kind, std::size_t e, const std::size_t* N, const real* d ){
    for (std::size_t i=0; i<ncomp; ++i) {
      auto i3 = i*3;
      if (kind == 0) {
        real u[] = { U(N[0],i), U(N[1],i), U(N[2],i), U(N[3],i) };
        for (std::size_t j=0; j<3; ++j) {
          real f[] = { d[(e*6+0)*S+j] * (u[1] + u[0]), d[(e*6+1)*S+j] * (u[2] + u[1]), d[(e*6+2)*S+j] * (u[0] + u[2]),
                       d[(e*6+3)*S+j] * (u[3] + u[0]), d[(e*6+4)*S+j] * (u[3] + u[1]), d[(e*6+5)*S+j] * (u[3] + u[2]) };
          G(N[0],i3+j) = G(N[0],i3+j) - f[0] + f[2] - f[3];
          G(N[1],i3+j) = G(N[1],i3+j) + f[0] - f[1] - f[4];
          G(N[2],i3+j) = G(N[2],i3+j) + f[1] - f[2] - f[5];
          G(N[3],i3+j) = G(N[3],i3+j) + f[3] + f[4] + f[5];
        }
      } else if (kind == 1) {
        real u[] = { U(N[0],i), U(N[1],i), U(N[2],i) };
        for (std::size_t j=0; j<3; ++j) {
          real f[] = { d[(e*3+0)*S+j] * (u[1] + u[0]), d[(e*3+1)*S+j] * (u[2] + u[1]), d[(e*3+2)*S+j] * (u[0] + u[2]) };
          G(N[0],i3+j) = G(N[0],i3+j) - f[0] + f[2];
          G(N[1],i3+j) = G(N[1],i3+j) + f[0] - f[1];
          G(N[2],i3+j) = G(N[2],i3+j) + f[1] - f[2];
        }
      } else {
        real u[] = { U(N[0],i), U(N[1],i) };
        for (std::size_t j=0; j<3; ++j) {
          real f = d[e*S+j] * (u[1] + u[0]);
          G(N[0],i3+j) -= f;
          G(N[1],i3+j) += f;
        }
      }
    }
  } );
  for (std::size_t e=0; e<triinpoel.size()/3; ++e) {
    const auto N = triinpoel.data() + e*3;
    real n[3]; crossdiv6( coord, N, n );
    for (std::size_t i=0; i<ncomp; ++i) {
      real u[] = { U(N[0],i), U(N[1],i), U(N[2],i) };
      auto i3 = i*3;
      auto f = (6.0*u[0] + u[1] + u[2])/8.0;
      G(N[0],i3+0) += f * n[0]; G(N[0],i3+1) += f * n[1]; G(N[0],i3+2) += f * n[2];
      f = (u[0] + 6.0*u[1] + u[2])/8.0;
      G(N[1],i3+0) += f * n[0]; G(N[1],i3+1) += f * n[1]; G(N[1],i3+2) += f * n[2];
      f = (u[0] + u[1] + 6.0*u[2])/8.0;
      G(N[2],i3+0) += f * n[0]; G(N[2],i3+1) += f * n[1]; G(N[2],i3+2) += f * n[2];
    }
  }
}

inline void chorin_vgrad( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                          const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                          const std::vector< std::size_t >& triinpoel, const Fields& U, Fields& G )
{ chorin_grad_impl( dsupedge, dsupint, coord, triinpoel, U.nprop(),
                    [&]( std::size_t p, std::size_t i ){ return U(p,i); }, G ); }

inline void chorin_grad( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                         const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                         const std::vector< std::size_t >& triinpoel, const std::vector< real >& U, Fields& G )
{ chorin_grad_impl( dsupedge, dsupint, coord, triinpoel, 1,
                    [&]( std::size_t p, std::size_t ){ return U[p]; }, G ); }

//! momentum flux of an edge (:451-480) and of a point (:482-509)
inline real chorin_flux2( const Fields& U, const Fields& G, std::size_t i, std::size_t j, std::size_t p, std::size_t q,
                          std::size_t pos = 0 ) {
  auto inv = U(p,i+pos)*U(p,j+pos) + U(q,i+pos)*U(q,j+pos);
  auto eps = std::numeric_limits< real >::epsilon();
  auto mu = cfg().mu;
  if (mu < eps) return -inv;
  auto vis = G(p,i*3+j) + G(p,j*3+i) + G(q,i*3+j) + G(q,j*3+i);
  if (i == j) vis -= 2.0/3.0 * ( G(p,0) + G(p,4) + G(p,8) + G(q,0) + G(q,4) + G(q,8) );
  return mu*vis - inv;
}
inline real chorin_flux1( const Fields& U, const Fields& G, std::size_t i, std::size_t j, std::size_t p,
                          std::size_t pos = 0 ) {
  auto inv = U(p,i+pos)*U(p,j+pos);
  auto eps = std::numeric_limits< real >::epsilon();
  auto mu = cfg().mu;
  if (mu < eps) return -inv;
  auto vis = G(p,i*3+j) + G(p,j*3+i);
  if (i == j) vis -= 2.0/3.0 * ( G(p,0) + G(p,4) + G(p,8) );
  return mu*vis - inv;
}

//! chorin::flux, Chorin.cpp:511-638 (accumulates into F)
inline void chorin_flux( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                         const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                         const std::vector< std::size_t >& triinpoel, const Fields& U, const Fields& G, Fields& F,
                         std::size_t S = 5, std::size_t pos = 0 )
{
  chorin_foredge( dsupedge, dsupint, [&]( int kind, std::size_t e, const std::size_t* N, const real* d ){
    for (std::size_t i=0; i<3; ++i)
      for (std::size_t j=0; j<3; ++j) {
        if (kind == 0) {
          real f[] = { d[(e*6+0)*S+j] * chorin_flux2(U,G,i,j,N[1],N[0],pos), d[(e*6+1)*S+j] * chorin_flux2(U,G,i,j,N[2],N[1],pos),
                       d[(e*6+2)*S+j] * chorin_flux2(U,G,i,j,N[0],N[2],pos), d[(e*6+3)*S+j] * chorin_flux2(U,G,i,j,N[3],N[0],pos),
                       d[(e*6+4)*S+j] * chorin_flux2(U,G,i,j,N[3],N[1],pos), d[(e*6+5)*S+j] * chorin_flux2(U,G,i,j,N[3],N[2],pos) };
          F(N[0],i) = F(N[0],i) - f[0] + f[2] - f[3];
          F(N[1],i) = F(N[1],i) + f[0] - f[1] - f[4];
          F(N[2],i) = F(N[2],i) + f[1] - f[2] - f[5];
          F(N[3],i) = F(N[3],i) + f[3] + f[4] + f[5];
        } else if (kind == 1) {
          real f[] = { d[(e*3+0)*S+j] * chorin_flux2(U,G,i,j,N[1],N[0],pos), d[(e*3+1)*S+j] * chorin_flux2(U,G,i,j,N[2],N[1],pos),
                       d[(e*3+2)*S+j] * chorin_flux2(U,G,i,j,N[0],N[2],pos) };
          F(N[0],i) = F(N[0],i) - f[0] + f[2];
          F(N[1],i) = F(N[1],i) + f[0] - f[1];
          F(N[2],i) = F(N[2],i) + f[1] - f[2];
        } else {
          real f = d[e*S+j] * chorin_flux2(U,G,i,j,N[1],N[0],pos);
          F(N[0],i) -= f;
          F(N[1],i) += f;
        }
      }
  } );
  for (std::size_t e=0; e<triinpoel.size()/3; ++e) {
    const auto N = triinpoel.data() + e*3;
    real n[3]; crossdiv6( coord, N, n );
    for (std::size_t i=0; i<3; ++i) {
      auto fxA = chorin_flux1(U,G,i,0,N[0],pos), fyA = chorin_flux1(U,G,i,1,N[0],pos), fzA = chorin_flux1(U,G,i,2,N[0],pos);
      auto fxB = chorin_flux1(U,G,i,0,N[1],pos), fyB = chorin_flux1(U,G,i,1,N[1],pos), fzB = chorin_flux1(U,G,i,2,N[1],pos);
      auto fxC = chorin_flux1(U,G,i,0,N[2],pos), fyC = chorin_flux1(U,G,i,1,N[2],pos), fzC = chorin_flux1(U,G,i,2,N[2],pos);
      auto fx = (6.0*fxA + fxB + fxC)/8.0;
      auto fy = (6.0*fyA + fyB + fyC)/8.0;
      auto fz = (6.0*fzA + fzB + fzC)/8.0;
      F(N[0],i) += fx*n[0] + fy*n[1] + fz*n[2];
      fx = (fxA + 6.0*fxB + fxC)/8.0;
      fy = (fyA + 6.0*fyB + fyC)/8.0;
      fz = (fzA + 6.0*fzB + fzC)/8.0;
      F(N[1],i) += fx*n[0] + fy*n[1] + fz*n[2];
      fx = (fxA + fxB + 6.0*fxC)/8.0;
      fy = (fyA + fyB + 6.0*fyC)/8.0;
      fz = (fzA + fzB + 6.0*fzC)/8.0;
      F(N[2],i) += fx*n[0] + fy*n[1] + fz*n[2];
    }
  }
}

//! advection edge flux with second-order damping, Chorin.cpp:640-709 (velocity components)
inline void chorin_adv_damp2( const real supint[], const Fields& U, const Fields&, const std::vector< real >& P,
                              const Coords&, std::size_t p, std::size_t q, real f[] )
{
  auto nx = supint[0], ny = supint[1], nz = supint[2];
  auto uL = U(p,0), vL = U(p,1), wL = U(p,2);
  auto vnL = uL*nx + vL*ny + wL*nz;
  auto uR = U(q,0), vR = U(q,1), wR = U(q,2);
  auto vnR = uR*nx + vR*ny + wR*nz;
  real aw = 0.0;
  if (cfg().stab) aw = std::abs( vnL + vnR ) / 2.0;
  if (cfg().stab2) aw += cfg().stab2coef * std::max( std::abs(vnL), std::abs(vnR) );
  auto v = supint[4] * cfg().mu;
  auto pf = P[p] + P[q];
  f[0] = uL*vnL + uR*vnR + pf*nx + (aw-v)*(uR-uL);
  f[1] = vL*vnL + vR*vnR + pf*ny + (aw-v)*(vR-vL);
  f[2] = wL*vnL + wR*vnR + pf*nz + (aw-v)*(wR-wL);
  auto ncomp = U.nprop();
  if (ncomp == 3) return;
  auto d = supint[4] * cfg().dif;
  for (std::size_t c=3; c<ncomp; ++c) f[c] = U(p,c)*vnL + U(q,c)*vnR + (aw-d)*(U(q,c)-U(p,c));
}

//! advection edge flux with fourth-order damping (limited reconstruction), Chorin.cpp:711-829
inline void chorin_adv_damp4( const real supint[], const Fields& U, const Fields& G, const std::vector< real >& P,
                              const Coords& coord, std::size_t p, std::size_t q, real f[] )
{
  auto dx = coord[0][q] - coord[0][p];
  auto dy = coord[1][q] - coord[1][p];
  auto dz = coord[2][q] - coord[2][p];
  auto ncomp = U.nprop();
  std::vector< real > uL( ncomp ), uR( ncomp );
  for (std::size_t i=0; i<ncomp; ++i) { uL[i] = U(p,i); uR[i] = U(q,i); }
  for (std::size_t c=0; c<ncomp; ++c) {
    auto g1 = G(p,c*3+0)*dx + G(p,c*3+1)*dy + G(p,c*3+2)*dz;
    auto g2 = G(q,c*3+0)*dx + G(q,c*3+1)*dy + G(q,c*3+2)*dz;
    auto delta2 = uR[c] - uL[c];
    auto delta1 = 2.0 * g1 - delta2;
    auto delta3 = 2.0 * g2 - delta2;
    auto rL = (delta2 + muscl_eps) / (delta1 + muscl_eps);
    auto rR = (delta2 + muscl_eps) / (delta3 + muscl_eps);
    auto rLinv = (delta1 + muscl_eps) / (delta2 + muscl_eps);
    auto rRinv = (delta3 + muscl_eps) / (delta2 + muscl_eps);
    auto phiL = (std::abs(rL) + rL) / (std::abs(rL) + 1.0);
    auto phiR = (std::abs(rR) + rR) / (std::abs(rR) + 1.0);
    auto phi_L_inv = (std::abs(rLinv) + rLinv) / (std::abs(rLinv) + 1.0);
    auto phi_R_inv = (std::abs(rRinv) + rRinv) / (std::abs(rRinv) + 1.0);
    uL[c] += 0.25*(delta1*(1.0-muscl_const)*phiL + delta2*(1.0+muscl_const)*phi_L_inv);
    uR[c] -= 0.25*(delta3*(1.0-muscl_const)*phiR + delta2*(1.0+muscl_const)*phi_R_inv);
  }
  auto nx = supint[0], ny = supint[1], nz = supint[2];
  auto vnL = uL[0]*nx + uL[1]*ny + uL[2]*nz;
  auto vnR = uR[0]*nx + uR[1]*ny + uR[2]*nz;
  real aw = 0.0;
  if (cfg().stab) aw = std::abs( vnL + vnR ) / 2.0;
  if (cfg().stab2) aw += cfg().stab2coef * std::max( std::abs(vnL), std::abs(vnR) );
  auto v = supint[4] * cfg().mu;
  auto pf = P[p] + P[q];
  f[0] = uL[0]*vnL + uR[0]*vnR + pf*nx + aw*(uR[0]-uL[0]) - v*(U(q,0)-U(p,0));
  f[1] = uL[1]*vnL + uR[1]*vnR + pf*ny + aw*(uR[1]-uL[1]) - v*(U(q,1)-U(p,1));
  f[2] = uL[2]*vnL + uR[2]*vnR + pf*nz + aw*(uR[2]-uL[2]) - v*(U(q,2)-U(p,2));
  if (ncomp == 3) return;
  auto d = supint[4] * cfg().dif;
  for (std::size_t c=3; c<ncomp; ++c) f[c] = uL[c]*vnL + uR[c]*vnR + aw*(uR[c]-uL[c]) - d*(U(q,c)-U(p,c));
}

//! chorin::rhs = adv (:831-982) + src (:984-1007), Chorin.cpp:1010-1044
inline void chorin_rhs( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                        const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                        const std::vector< std::size_t >& triinpoel, const std::vector< real >& v, real t,
                        const std::vector< real >& P, const Fields& U, const Fields& G, Fields& R )
{
  R.fill( 0.0 );
  auto ncomp = U.nprop();
  auto adv = cfg().flux == "damp2" ? chorin_adv_damp2 : chorin_adv_damp4;
  if (cfg().flux != "damp2" && cfg().flux != "damp4") throw std::runtime_error( "oracle port: Flux not correctly configured" );
  std::vector< real > fb( 6*ncomp );
  real* f[6]; for (int k=0; k<6; ++k) f[k] = fb.data() + static_cast<std::size_t>(k)*ncomp;
  chorin_foredge( dsupedge, dsupint, [&]( int kind, std::size_t e, const std::size_t* N, const real* d ){
    if (kind == 0) {
      adv( d+(e*6+0)*5, U, G, P, coord, N[0], N[1], f[0] );
      adv( d+(e*6+1)*5, U, G, P, coord, N[1], N[2], f[1] );
      adv( d+(e*6+2)*5, U, G, P, coord, N[2], N[0], f[2] );
      adv( d+(e*6+3)*5, U, G, P, coord, N[0], N[3], f[3] );
      adv( d+(e*6+4)*5, U, G, P, coord, N[1], N[3], f[4] );
      adv( d+(e*6+5)*5, U, G, P, coord, N[2], N[3], f[5] );
      for (std::size_t c=0; c<ncomp; ++c) {
        R(N[0],c) = R(N[0],c) - f[0][c] + f[2][c] - f[3][c];
        R(N[1],c) = R(N[1],c) + f[0][c] - f[1][c] - f[4][c];
        R(N[2],c) = R(N[2],c) + f[1][c] - f[2][c] - f[5][c];
        R(N[3],c) = R(N[3],c) + f[3][c] + f[4][c] + f[5][c];
      }
    } else if (kind == 1) {
      adv( d+(e*3+0)*5, U, G, P, coord, N[0], N[1], f[0] );
      adv( d+(e*3+1)*5, U, G, P, coord, N[1], N[2], f[1] );
      adv( d+(e*3+2)*5, U, G, P, coord, N[2], N[0], f[2] );
      for (std::size_t c=0; c<ncomp; ++c) {
        R(N[0],c) = R(N[0],c) - f[0][c] + f[2][c];
        R(N[1],c) = R(N[1],c) + f[0][c] - f[1][c];
        R(N[2],c) = R(N[2],c) + f[1][c] - f[2][c];
      }
    } else {
      adv( d+e*5, U, G, P, coord, N[0], N[1], f[0] );
      for (std::size_t c=0; c<ncomp; ++c) { R(N[0],c) -= f[0][c]; R(N[1],c) += f[0][c]; }
    }
  } );
  std::vector< real > fl( ncomp*3 );
  auto F = [&]( std::size_t c, std::size_t k ) -> real& { return fl[c*3+k]; };
  for (std::size_t e=0; e<triinpoel.size()/3; ++e) {
    const auto N = triinpoel.data() + e*3;
    real n[3]; crossdiv6( coord, N, n );
    for (std::size_t k=0; k<3; ++k) {
      auto u = U(N[k],0), vv = U(N[k],1), w = U(N[k],2);
      auto p = P[N[k]];
      auto vn = n[0]*u + n[1]*vv + n[2]*w;
      F(0,k) = u*vn + p*n[0];
      F(1,k) = vv*vn + p*n[1];
      F(2,k) = w*vn + p*n[2];
      for (std::size_t c=3; c<ncomp; ++c) F(c,k) = U(N[k],c)*vn;
    }
    for (std::size_t c=0; c<ncomp; ++c) {
      R(N[0],c) += (6.0*F(c,0) + F(c,1) + F(c,2))/8.0;
      R(N[1],c) += (F(c,0) + 6.0*F(c,1) + F(c,2))/8.0;
      R(N[2],c) += (F(c,0) + F(c,1) + 6.0*F(c,2))/8.0;
    }
  }
  if (auto s_ = SRC())
    for (std::size_t p=0; p<R.nunk(); ++p) {
      auto s = s_( coord[0][p], coord[1][p], coord[2][p], t );
      for (std::size_t c=0; c<s.size(); ++c) R(p,c) -= s[c] * v[p];
    }
}

inline void dirbcp( Fields& U, const Coords& coord, const std::vector< std::size_t >& dirbcmaskp,
                    const std::vector< double >& dirbcvalp )
{
  auto ic = PRESSURE_IC();
  for (std::size_t i=0; i<dirbcmaskp.size()/2; ++i) {
    auto p = dirbcmaskp[i*2+0];
    auto mask = dirbcmaskp[i*2+1];
    if (mask == 1) U(p,0) = ic( coord[0][p], coord[1][p], coord[2][p] );
    else if (mask == 2 && !dirbcvalp.empty()) U(p,0) = dirbcvalp[i*2+1];
  }
}

// ---- Lohner.cpp: edge operators of LohCG (artificial compressibility; unknowns p,u,v,w[,c..]) ----
// superedge integrals have stride 4: normal(3), grad_p.grad_q/(6J) (LohCG.cpp:407-453). div, grad,
// vgrad and flux are the Chorin operators on these integrals with the velocity at components 1..3.
inline void lohner_div( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                        const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                        const std::vector< std::size_t >& triinpoel, const Fields& U, std::vector< real >& D,
                        std::size_t pos )                                   // Lohner.cpp:35-159
{ static const std::vector< real > nop; static const Fields nof;
  chorin_div( dsupedge, dsupint, coord, triinpoel, 0.0, nop, nof, U, D, false, 4, pos ); }
inline void lohner_grad( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                         const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                         const std::vector< std::size_t >& triinpoel, const std::vector< real >& U, Fields& G )
{ chorin_grad_impl( dsupedge, dsupint, coord, triinpoel, 1,                 // :161-277
                    [&]( std::size_t p, std::size_t ){ return U[p]; }, G, 4 ); }
inline void lohner_vgrad( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                          const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                          const std::vector< std::size_t >& triinpoel, const Fields& U, Fields& G )
{ chorin_grad_impl( dsupedge, dsupint, coord, triinpoel, 3,                 // :279-400
                    [&]( std::size_t p, std::size_t i ){ return U(p,i+1); }, G, 4 ); }
inline void lohner_flux( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                         const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                         const std::vector< std::size_t >& triinpoel, const Fields& U, const Fields& G, Fields& F )
{ chorin_flux( dsupedge, dsupint, coord, triinpoel, U, G, F, 4, 1 ); }      // :402-589
inline void lohner_gradall( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                            const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                            const std::vector< std::size_t >& triinpoel, const Fields& U, Fields& G )
{ if (G.nprop() == 0) return;                                               // :591-722
  chorin_grad_impl( dsupedge, dsupint, coord, triinpoel, U.nprop(),
                    [&]( std::size_t p, std::size_t i ){ return U(p,i); }, G, 4 ); }

//! pressure + momentum (+ scalar) edge flux with second-order damping, Lohner.cpp:724-795
inline void lohner_adv_damp2( const real supint[], const Fields& U, const Fields&, const Coords&,
                              std::size_t p, std::size_t q, real f[] )
{
  auto ncomp = U.nprop();
  auto nx = supint[0], ny = supint[1], nz = supint[2];
  auto pL = U(p,0), uL = U(p,1), vL = U(p,2), wL = U(p,3);
  auto vnL = uL*nx + vL*ny + wL*nz;
  auto pR = U(q,0), uR = U(q,1), vR = U(q,2), wR = U(q,3);
  auto vnR = uR*nx + vR*ny + wR*nz;
  auto s = cfg().soundspeed;
  auto s2 = s*s;
  auto v = supint[3] * cfg().mu;
  real aw = 0.0;
  if (cfg().stab) aw = std::abs( vnL + vnR ) / 2.0;
  if (cfg().stab2) {
    auto len = length3( nx, ny, nz );
    auto sl = std::abs(vnL) + s*len;
    auto sr = std::abs(vnR) + s*len;
    aw += cfg().stab2coef * std::max(sl,sr);
  }
  auto pf = pL + pR;
  f[0] = (vnL + vnR + aw*(pR - pL))*s2;
  f[1] = uL*vnL + uR*vnR + pf*nx + (aw-v)*(uR - uL);
  f[2] = vL*vnL + vR*vnR + pf*ny + (aw-v)*(vR - vL);
  f[3] = wL*vnL + wR*vnR + pf*nz + (aw-v)*(wR - wL);
  auto d = supint[3] * cfg().dif;
  for (std::size_t c=4; c<ncomp; ++c) f[c] = U(p,c)*vnL + U(q,c)*vnR + (aw-d)*(U(q,c) - U(p,c));
}

//! the same with fourth-order damping (limited reconstruction of all unknowns), Lohner.cpp:797-914
inline void lohner_adv_damp4( const real supint[], const Fields& U, const Fields& G, const Coords& coord,
                              std::size_t p, std::size_t q, real f[] )
{
  const auto ncomp = U.nprop();
  auto dx = coord[0][q] - coord[0][p];
  auto dy = coord[1][q] - coord[1][p];
  auto dz = coord[2][q] - coord[2][p];
  std::vector< real > uL( ncomp ), uR( ncomp );
  for (std::size_t i=0; i<ncomp; ++i) { uL[i] = U(p,i); uR[i] = U(q,i); }
  for (std::size_t c=0; c<ncomp; ++c) {
    auto g1 = G(p,c*3+0)*dx + G(p,c*3+1)*dy + G(p,c*3+2)*dz;
    auto g2 = G(q,c*3+0)*dx + G(q,c*3+1)*dy + G(q,c*3+2)*dz;
    auto delta2 = uR[c] - uL[c];
    auto delta1 = 2.0 * g1 - delta2;
    auto delta3 = 2.0 * g2 - delta2;
    auto rL = (delta2 + muscl_eps) / (delta1 + muscl_eps);
    auto rR = (delta2 + muscl_eps) / (delta3 + muscl_eps);
    auto rLinv = (delta1 + muscl_eps) / (delta2 + muscl_eps);
    auto rRinv = (delta3 + muscl_eps) / (delta2 + muscl_eps);
    auto phiL = (std::abs(rL) + rL) / (std::abs(rL) + 1.0);
    auto phiR = (std::abs(rR) + rR) / (std::abs(rR) + 1.0);
    auto phi_L_inv = (std::abs(rLinv) + rLinv) / (std::abs(rLinv) + 1.0);
    auto phi_R_inv = (std::abs(rRinv) + rRinv) / (std::abs(rRinv) + 1.0);
    uL[c] += 0.25*(delta1*(1.0-muscl_const)*phiL + delta2*(1.0+muscl_const)*phi_L_inv);
    uR[c] -= 0.25*(delta3*(1.0-muscl_const)*phiR + delta2*(1.0+muscl_const)*phi_R_inv);
  }
  auto nx = supint[0], ny = supint[1], nz = supint[2];
  auto vnL = uL[1]*nx + uL[2]*ny + uL[3]*nz;
  auto vnR = uR[1]*nx + uR[2]*ny + uR[3]*nz;
  auto s = cfg().soundspeed;
  auto s2 = s*s;
  auto v = supint[3] * cfg().mu;
  real aw = 0.0;
  if (cfg().stab) aw = std::abs( vnL + vnR ) / 2.0;
  if (cfg().stab2) {
    auto len = length3( nx, ny, nz );
    auto sl = std::abs(vnL) + s*len;
    auto sr = std::abs(vnR) + s*len;
    aw += cfg().stab2coef * std::max(sl,sr);
  }
  auto pf = uL[0] + uR[0];
  f[0] = (vnL + vnR + aw*(uR[0]-uL[0]))*s2;
  f[1] = uL[1]*vnL + uR[1]*vnR + pf*nx + aw*(uR[1]-uL[1]) - v*(U(q,1)-U(p,1));
  f[2] = uL[2]*vnL + uR[2]*vnR + pf*ny + aw*(uR[2]-uL[2]) - v*(U(q,2)-U(p,2));
  f[3] = uL[3]*vnL + uR[3]*vnR + pf*nz + aw*(uR[3]-uL[3]) - v*(U(q,3)-U(p,3));
  auto d = supint[3] * cfg().dif;
  for (std::size_t c=4; c<ncomp; ++c) f[c] = uL[c]*vnL + uR[c]*vnR + aw*(uR[c]-uL[c]) - d*(U(q,c)-U(p,c));
}

//! lohner::rhs = adv (:916-1071) + src (:1073-1097), Lohner.cpp:1099-1130
inline void lohner_rhs( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                        const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                        const std::vector< std::size_t >& triinpoel, const std::vector< real >& v, real t,
                        const Fields& U, const Fields& G, Fields& R )
{
  R.fill( 0.0 );
  auto ncomp = U.nprop();
  if (cfg().flux != "damp2" && cfg().flux != "damp4") throw std::runtime_error( "oracle port: Flux not correctly configured" );
  auto adv = cfg().flux == "damp2" ? lohner_adv_damp2 : lohner_adv_damp4;
  std::vector< real > fb( 6*ncomp );
  real* f[6]; for (int k=0; k<6; ++k) f[k] = fb.data() + static_cast<std::size_t>(k)*ncomp;
  chorin_foredge( dsupedge, dsupint, [&]( int kind, std::size_t e, const std::size_t* N, const real* d ){
    if (kind == 0) {
      adv( d+(e*6+0)*4, U, G, coord, N[0], N[1], f[0] );
      adv( d+(e*6+1)*4, U, G, coord, N[1], N[2], f[1] );
      adv( d+(e*6+2)*4, U, G, coord, N[2], N[0], f[2] );
      adv( d+(e*6+3)*4, U, G, coord, N[0], N[3], f[3] );
      adv( d+(e*6+4)*4, U, G, coord, N[1], N[3], f[4] );
      adv( d+(e*6+5)*4, U, G, coord, N[2], N[3], f[5] );
      for (std::size_t c=0; c<ncomp; ++c) {
        R(N[0],c) = R(N[0],c) - f[0][c] + f[2][c] - f[3][c];
        R(N[1],c) = R(N[1],c) + f[0][c] - f[1][c] - f[4][c];
        R(N[2],c) = R(N[2],c) + f[1][c] - f[2][c] - f[5][c];
        R(N[3],c) = R(N[3],c) + f[3][c] + f[4][c] + f[5][c];
      }
    } else if (kind == 1) {
      adv( d+(e*3+0)*4, U, G, coord, N[0], N[1], f[0] );
      adv( d+(e*3+1)*4, U, G, coord, N[1], N[2], f[1] );
      adv( d+(e*3+2)*4, U, G, coord, N[2], N[0], f[2] );
      for (std::size_t c=0; c<ncomp; ++c) {
        R(N[0],c) = R(N[0],c) - f[0][c] + f[2][c];
        R(N[1],c) = R(N[1],c) + f[0][c] - f[1][c];
        R(N[2],c) = R(N[2],c) + f[1][c] - f[2][c];
      }
    } else {
      adv( d+e*4, U, G, coord, N[0], N[1], f[0] );
      for (std::size_t c=0; c<ncomp; ++c) { R(N[0],c) -= f[0][c]; R(N[1],c) += f[0][c]; }
    }
  } );
  auto s = cfg().soundspeed;
  auto s2 = s * s;
  std::vector< real > fl( ncomp*3 );
  auto F = [&]( std::size_t c, std::size_t k ) -> real& { return fl[c*3+k]; };
  for (std::size_t e=0; e<triinpoel.size()/3; ++e) {
    const auto N = triinpoel.data() + e*3;
    real n[3]; crossdiv6( coord, N, n );
    for (std::size_t k=0; k<3; ++k) {
      auto p = U(N[k],0), u = U(N[k],1), vv = U(N[k],2), w = U(N[k],3);
      auto vn = n[0]*u + n[1]*vv + n[2]*w;
      F(0,k) = vn * s2;
      F(1,k) = u*vn + p*n[0];
      F(2,k) = vv*vn + p*n[1];
      F(3,k) = w*vn + p*n[2];
      for (std::size_t c=4; c<ncomp; ++c) F(c,k) = U(N[k],c)*vn;
    }
    for (std::size_t c=0; c<ncomp; ++c) {
      R(N[0],c) += (6.0*F(c,0) + F(c,1) + F(c,2))/8.0;
      R(N[1],c) += (F(c,0) + 6.0*F(c,1) + F(c,2))/8.0;
      R(N[2],c) += (F(c,0) + F(c,1) + 6.0*F(c,2))/8.0;
    }
  }
  if (auto s_ = SRC())
    for (std::size_t p=0; p<R.nunk(); ++p) {
      auto sv = s_( coord[0][p], coord[1][p], coord[2][p], t );
      for (std::size_t c=0; c<sv.size(); ++c) R(p,c) -= sv[c] * v[p];
    }
}

// ---- Zalesak.cpp: Taylor-Galerkin two-step edge flux for ZalCG --------------------------
//! edge flux, Zalesak.cpp:31-200 (problems without source term; the f[ncomp..2ncomp) source
//! half is only produced when problems::SRC() is set)
inline void zal_advedge( const real supint[], const Fields& U, const Coords& coord, real t, real dt,
                         const std::vector< real >& tp, const std::vector< real >& dtp,
                         std::size_t p, std::size_t q, real f[], const ICFn& src )
{
  const auto steady = cfg().steady;
  const auto ncomp = U.nprop();
  const auto& x = coord[0]; const auto& y = coord[1]; const auto& z = coord[2];
  auto dx = x[p] - x[q], dy = y[p] - y[q], dz = z[p] - z[q];
  auto dl = dx*dx + dy*dy + dz*dz;
  dx /= dl; dy /= dl; dz /= dl;
  auto rL = U(p,0), ruL = U(p,1), rvL = U(p,2), rwL = U(p,3), reL = U(p,4);
  auto pL = eos_pressure( reL - 0.5*(ruL*ruL + rvL*rvL + rwL*rwL)/rL );
  auto dnL = (ruL*dx + rvL*dy + rwL*dz)/rL;
  auto rR = U(q,0), ruR = U(q,1), rvR = U(q,2), rwR = U(q,3), reR = U(q,4);
  auto pR = eos_pressure( reR - 0.5*(ruR*ruR + rvR*rvR + rwR*rwR)/rR );
  auto dnR = (ruR*dx + rvR*dy + rwR*dz)/rR;
  auto nx = supint[0], ny = supint[1], nz = supint[2];
  std::vector< real > ue( ncomp );
  if (steady) dt = (dtp[p] + dtp[q])/2.0;                                   // Zalesak.cpp:107
  auto dp = pL - pR;
  ue[0] = 0.5*(rL + rR - dt*(rL*dnL - rR*dnR));
  ue[1] = 0.5*(ruL + ruR - dt*(ruL*dnL - ruR*dnR + dp*dx));
  ue[2] = 0.5*(rvL + rvR - dt*(rvL*dnL - rvR*dnR + dp*dy));
  ue[3] = 0.5*(rwL + rwR - dt*(rwL*dnL - rwR*dnR + dp*dz));
  ue[4] = 0.5*(reL + reR - dt*((reL+pL)*dnL - (reR+pR)*dnR));
  for (std::size_t c=5; c<ncomp; ++c) ue[c] = 0.5*(U(p,c) + U(q,c) - dt*(U(p,c)*dnL - U(q,c)*dnR));
  if (src) {
    if (steady) t = (tp[p] + tp[q])/2.0;                                    // :125
    auto coef = dt/4.0;
    auto sL = src( x[p], y[p], z[p], t );
    auto sR = src( x[q], y[q], z[q], t );
    for (std::size_t c=0; c<ncomp; ++c) ue[c] += coef*(sL[c] + sR[c]);
  }
  auto rh = ue[0], ruh = ue[1], rvh = ue[2], rwh = ue[3], reh = ue[4];
  auto ph = eos_pressure( reh - 0.5*(ruh*ruh + rvh*rvh + rwh*rwh)/rh );
  auto vn = (ruh*nx + rvh*ny + rwh*nz)/rh;
  f[0] = 2.0*rh*vn;
  f[1] = 2.0*(ruh*vn + ph*nx);
  f[2] = 2.0*(rvh*vn + ph*ny);
  f[3] = 2.0*(rwh*vn + ph*nz);
  f[4] = 2.0*(reh + ph)*vn;
  for (std::size_t c=5; c<ncomp; ++c) f[c] = 2.0*ue[c]*vn;
  if (src) {
    auto coef = -5.0/3.0*supint[3];
    auto se = src( (x[p] + x[q])/2.0, (y[p] + y[q])/2.0, (z[p] + z[q])/2.0, t+dt/2.0 );
    for (std::size_t c=0; c<ncomp; ++c) f[ncomp+c] = coef*se[c];
  }
  if (!cfg().stab2) return;
  auto stab2coef = cfg().stab2coef;
  auto vnL = (ruL*nx + rvL*ny + rwL*nz)/rL;
  auto vnR = (ruR*nx + rvR*ny + rwR*nz)/rR;
  auto len = std::sqrt( nx*nx + ny*ny + nz*nz );
  auto cL = eos_soundspeed( std::max(rL,1.0e-8), std::max(pL,0.0) );
  auto cR = eos_soundspeed( std::max(rR,1.0e-8), std::max(pR,0.0) );
  auto sl = std::abs(vnL) + cL*len;
  auto sr = std::abs(vnR) + cR*len;
  auto fw = stab2coef * std::max( sl, sr );
  f[0] -= fw*(rL - rR); f[1] -= fw*(ruL - ruR); f[2] -= fw*(rvL - rvR);
  f[3] -= fw*(rwL - rwR); f[4] -= fw*(reL - reR);
  for (std::size_t c=5; c<ncomp; ++c) f[c] -= fw*(U(p,c) - U(q,c));
}

//! zalesak::rhs, Zalesak.cpp:202-457 (superedge integrals have stride 4)
inline void zal_rhs( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                     const std::array< std::vector< real >, 3 >& dsupint, const Coords& coord,
                     const std::vector< std::size_t >& triinpoel, const std::vector< std::uint8_t >& besym,
                     real t, real dt, const std::vector< real >& tp, const std::vector< real >& dtp,
                     const Fields& U, Fields& R )
{
  auto ncomp = U.nprop();
  auto src = SRC();
  R.fill( 0.0 );
  std::vector< real > fb( 6*ncomp*2 );
  real* f[6]; for (int k=0; k<6; ++k) f[k] = fb.data() + static_cast<std::size_t>(k)*ncomp*2;
  for (std::size_t e=0; e<dsupedge[0].size()/4; ++e) {
    const auto N = dsupedge[0].data() + e*4;
    const auto d = dsupint[0].data();
    zal_advedge( d+(e*6+0)*4, U, coord, t, dt, tp, dtp, N[0], N[1], f[0], src );
    zal_advedge( d+(e*6+1)*4, U, coord, t, dt, tp, dtp, N[1], N[2], f[1], src );
    zal_advedge( d+(e*6+2)*4, U, coord, t, dt, tp, dtp, N[2], N[0], f[2], src );
    zal_advedge( d+(e*6+3)*4, U, coord, t, dt, tp, dtp, N[0], N[3], f[3], src );
    zal_advedge( d+(e*6+4)*4, U, coord, t, dt, tp, dtp, N[1], N[3], f[4], src );
    zal_advedge( d+(e*6+5)*4, U, coord, t, dt, tp, dtp, N[2], N[3], f[5], src );
    for (std::size_t c=0; c<ncomp; ++c) {
      R(N[0],c) = R(N[0],c) - f[0][c] + f[2][c] - f[3][c];
      R(N[1],c) = R(N[1],c) + f[0][c] - f[1][c] - f[4][c];
      R(N[2],c) = R(N[2],c) + f[1][c] - f[2][c] - f[5][c];
      R(N[3],c) = R(N[3],c) + f[3][c] + f[4][c] + f[5][c];
      if (src) {
        auto nc = ncomp + c;
        R(N[0],c) += f[0][nc] + f[2][nc] + f[3][nc];
        R(N[1],c) += f[0][nc] + f[1][nc] + f[4][nc];
        R(N[2],c) += f[1][nc] + f[2][nc] + f[5][nc];
        R(N[3],c) += f[3][nc] + f[4][nc] + f[5][nc];
      }
    }
  }
  for (std::size_t e=0; e<dsupedge[1].size()/3; ++e) {
    const auto N = dsupedge[1].data() + e*3;
    const auto d = dsupint[1].data();
    zal_advedge( d+(e*3+0)*4, U, coord, t, dt, tp, dtp, N[0], N[1], f[0], src );
    zal_advedge( d+(e*3+1)*4, U, coord, t, dt, tp, dtp, N[1], N[2], f[1], src );
    zal_advedge( d+(e*3+2)*4, U, coord, t, dt, tp, dtp, N[2], N[0], f[2], src );
    for (std::size_t c=0; c<ncomp; ++c) {
      R(N[0],c) = R(N[0],c) - f[0][c] + f[2][c];
      R(N[1],c) = R(N[1],c) + f[0][c] - f[1][c];
      R(N[2],c) = R(N[2],c) + f[1][c] - f[2][c];
      if (src) {
        auto nc = ncomp + c;
        R(N[0],c) += f[0][nc] + f[2][nc];
        R(N[1],c) += f[0][nc] + f[1][nc];
        R(N[2],c) += f[1][nc] + f[2][nc];
      }
    }
  }
  for (std::size_t e=0; e<dsupedge[2].size()/2; ++e) {
    const auto N = dsupedge[2].data() + e*2;
    const auto d = dsupint[2].data();
    zal_advedge( d+e*4, U, coord, t, dt, tp, dtp, N[0], N[1], f[0], src );
    for (std::size_t c=0; c<ncomp; ++c) {
      R(N[0],c) -= f[0][c];
      R(N[1],c) += f[0][c];
      if (src) { auto nc = ncomp + c; R(N[0],c) += f[0][nc]; R(N[1],c) += f[0][nc]; }
    }
  }
  advbnd( triinpoel, coord, besym, U, R );      // Zalesak.cpp:309-419, same form as Riemann.cpp:768-878
}

// ---- Kozak.cpp: element-based Taylor-Galerkin rhs for KozCG --------------------------------
inline void koz_rhs( const std::vector< std::size_t >& inpoel, const Coords& coord, real t, real dt,
                     const Fields& U, Fields& R )                               // Kozak.cpp:29-180
{
  R.fill( 0.0 );
  const auto ncomp = U.nprop();
  const auto& x = coord[0]; const auto& y = coord[1]; const auto& z = coord[2];
  auto src = SRC();
  std::vector< real > ue( ncomp );
  for (std::size_t e=0; e<inpoel.size()/4; ++e) {
    const auto N = inpoel.data() + e*4;
    real ba[3] = { x[N[1]]-x[N[0]], y[N[1]]-y[N[0]], z[N[1]]-z[N[0]] },
         ca[3] = { x[N[2]]-x[N[0]], y[N[2]]-y[N[0]], z[N[2]]-z[N[0]] },
         da[3] = { x[N[3]]-x[N[0]], y[N[3]]-y[N[0]], z[N[3]]-z[N[0]] };
    auto cross = []( const real a[3], const real b[3], real r[3] ){
      r[0] = a[1]*b[2] - b[1]*a[2]; r[1] = a[2]*b[0] - b[2]*a[0]; r[2] = a[0]*b[1] - b[0]*a[1]; };
    real grad[4][3], cx[3];
    cross( ca, da, cx );
    const auto J = ba[0]*cx[0] + ba[1]*cx[1] + ba[2]*cx[2];
    cross( ca, da, grad[1] ); cross( da, ba, grad[2] ); cross( ba, ca, grad[3] );
    for (std::size_t i=0; i<3; ++i) grad[0][i] = -grad[1][i]-grad[2][i]-grad[3][i];
    real p[4];
    for (std::size_t a=0; a<4; ++a) {
      auto r = U(N[a],0), ru = U(N[a],1), rv = U(N[a],2), rw = U(N[a],3);
      p[a] = eos_pressure( U(N[a],4) - 0.5*(ru*ru + rv*rv + rw*rw)/r );
    }
    for (std::size_t c=0; c<ncomp; ++c) ue[c] = (U(N[0],c) + U(N[1],c) + U(N[2],c) + U(N[3],c))/4.0;
    auto coef = dt/J/2.0;
    for (std::size_t j=0; j<3; ++j)
      for (std::size_t a=0; a<4; ++a) {
        auto cg = coef * grad[a][j];
        auto uj = U(N[a],j+1) / U(N[a],0);
        ue[0] -= cg * U(N[a],j+1);
        ue[1] -= cg * U(N[a],1) * uj;
        ue[2] -= cg * U(N[a],2) * uj;
        ue[3] -= cg * U(N[a],3) * uj;
        ue[j+1] -= cg * p[a];
        ue[4] -= cg * (U(N[a],4) + p[a]) * uj;
        for (std::size_t c=5; c<ncomp; ++c) ue[c] -= cg * U(N[a],c) * uj;
      }
    if (src) {
      coef = dt/8.0;
      for (std::size_t a=0; a<4; ++a) {
        auto s = src( x[N[a]], y[N[a]], z[N[a]], t );
        for (std::size_t c=0; c<ncomp; ++c) ue[c] += coef * s[c];
      }
    }
    auto r = ue[0], ru = ue[1], rv = ue[2], rw = ue[3];
    auto pr = eos_pressure( ue[4] - 0.5*(ru*ru + rv*rv + rw*rw)/r );
    coef = 1.0/6.0;
    for (std::size_t j=0; j<3; ++j) {
      auto uj = ue[j+1] / ue[0];
      for (std::size_t a=0; a<4; ++a) {
        auto cg = coef * grad[a][j];
        R(N[a],0) += cg * ue[j+1];
        R(N[a],1) += cg * ue[1] * uj;
        R(N[a],2) += cg * ue[2] * uj;
        R(N[a],3) += cg * ue[3] * uj;
        R(N[a],j+1) += cg * pr;
        R(N[a],4) += cg * (ue[4] + pr) * uj;
        for (std::size_t c=5; c<ncomp; ++c) R(N[a],c) += cg * ue[c] * uj;
      }
    }
    if (src) {
      auto se = src( (x[N[0]] + x[N[1]] + x[N[2]] + x[N[3]])/4.0, (y[N[0]] + y[N[1]] + y[N[2]] + y[N[3]])/4.0,
                     (z[N[0]] + z[N[1]] + z[N[2]] + z[N[3]])/4.0, t+dt/2.0 );
      coef = J/24.0;
      for (std::size_t a=0; a<4; ++a) for (std::size_t c=0; c<ncomp; ++c) R(N[a],c) += coef * se[c];
    }
  }
}

// ---- Mesh/DerivedData.cpp ----------------------------------------------------------
using LinkedList = std::pair< std::vector< std::size_t >, std::vector< std::size_t > >;

inline LinkedList genEsup( const std::vector< std::size_t >& inpoel, std::size_t nnpe ) // :49-130
{
  auto npoin = *std::max_element( inpoel.begin(), inpoel.end() ) + 1;
  std::vector< std::size_t > esup2( npoin+1, 0 );
  for (auto n : inpoel) ++esup2[ n+1 ];
  for (std::size_t i=1; i<npoin+1; ++i) esup2[i] += esup2[i-1];
  std::vector< std::size_t > esup1( esup2[npoin]+1 );
  std::size_t e = 0;
  for (auto n : inpoel) { auto j = esup2[n]+1; esup2[n] = j; esup1[j] = e/nnpe; ++e; }
  for (auto i=npoin; i>0; --i) esup2[i] = esup2[i-1];
  esup2[0] = 0;
  return { std::move(esup1), std::move(esup2) };
}

inline LinkedList genPsup( const std::vector< std::size_t >& inpoel, std::size_t nnpe,
                           const LinkedList& esup )                            // :132-215
{
  auto npoin = *std::max_element( inpoel.begin(), inpoel.end() ) + 1;
  const auto& esup1 = esup.first; const auto& esup2 = esup.second;
  std::vector< std::size_t > psup2( npoin+1 ), psup1( 1, 0 );
  std::vector< std::size_t > lpoin( npoin, 0 );
  psup2[0] = 0;
  std::size_t j = 0;
  for (std::size_t p=0; p<npoin; ++p) {
    for (std::size_t i=esup2[p]+1; i<=esup2[p+1]; ++i)
      for (std::size_t n=0; n<nnpe; ++n) {
        auto q = inpoel[ esup1[i]*nnpe + n ];
        if (q != p && lpoin[q] != p+1) { ++j; psup1.push_back( q ); lpoin[q] = p+1; }
      }
    psup2[p+1] = j;
  }
  for (std::size_t p=0; p<npoin; ++p)          // neighbour ids ascending per point :216-220
    std::sort( psup1.begin() + static_cast<std::ptrdiff_t>(psup2[p]+1),
               psup1.begin() + static_cast<std::ptrdiff_t>(psup2[p+1]+1) );
  return { std::move(psup1), std::move(psup2) };
}

} // port::
} // orc::

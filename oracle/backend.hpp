// oracle/backend.hpp -- TEST INFRASTRUCTURE ONLY (never linked by the product).
//
// Selects what the serial driver restatement (driver.hpp) calls for the
// Charm++-free parts of the path:
//   default      : the plain C++ restatement in physics_port.hpp
//   -DORACLE_REF : the reference's own, unmodified translation units compiled in
//                  place from /root/reference/src (oracle/Makefile target `ref`,
//                  output oracle/_ref/liboracle_ref.so). Hash containers are then
//                  the reference's tk::UnsMesh types as well.
#pragma once
#include "siphash.hpp"
#include "physics_port.hpp"

#ifdef ORACLE_REF
  #include "Fields.hpp"
  #include "UnsMesh.hpp"
  #include "DerivedData.hpp"
  #include "Riemann.hpp"
  #include "Zalesak.hpp"
  #include "Kozak.hpp"
  #include "Lax.hpp"
  #include "Chorin.hpp"
  #include "Lohner.hpp"
  #include "BC.hpp"
  #include "Problems.hpp"
  #include "InciterConfig.hpp"
  namespace inciter { extern ctr::Config g_cfg; }
#endif

namespace orc {
namespace be {

using Coords = std::array< std::vector< real >, 3 >;

#ifdef ORACLE_REF

using Fields = tk::Fields;
template< std::size_t N > using Hash = tk::UnsMesh::Hash< N >;
template< std::size_t N > using Eq = tk::UnsMesh::Eq< N >;
inline const char* name() { return "reference"; }

void set_cfg( const Cfg& c );   // ref_backend.cpp: fills inciter::g_cfg

inline void grad( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                  const std::array< std::vector< real >, 3 >& dsupint,
                  const Coords& coord, const std::vector< std::size_t >& triinpoel,
                  const Fields& U, Fields& G )
{ riemann::grad( dsupedge, dsupint, coord, triinpoel, U, G ); }

inline void rhs( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                 const std::array< std::vector< real >, 3 >& dsupint,
                 const Coords& coord, const std::vector< std::size_t >& triinpoel,
                 const std::vector< std::uint8_t >& besym, const Fields& G, const Fields& U,
                 const std::vector< real >& v, real t, const std::vector< real >& tp, Fields& R )
{ riemann::rhs( dsupedge, dsupint, coord, triinpoel, besym, G, U, v, t, tp, R ); }

inline void zal_rhs( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                     const std::array< std::vector< real >, 3 >& dsupint,
                     const Coords& coord, const std::vector< std::size_t >& triinpoel,
                     const std::vector< std::uint8_t >& besym, real t, real dt,
                     const std::vector< real >& tp, const std::vector< real >& dtp, const Fields& U, Fields& R )
{ zalesak::rhs( dsupedge, dsupint, coord, triinpoel, besym, t, dt, tp, dtp, U, R ); }

inline void koz_rhs( const std::vector< std::size_t >& inpoel, const Coords& coord, real t, real dt,
                     const Fields& U, Fields& R )
{ std::vector< real > tp, dtp; kozak::rhs( inpoel, coord, t, dt, tp, dtp, U, R ); }

inline void lax_grad( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                      const std::array< std::vector< real >, 3 >& dsupint,
                      const Coords& coord, const std::vector< std::size_t >& triinpoel,
                      const Fields& U, Fields& G )
{ lax::grad( dsupedge, dsupint, coord, triinpoel, U, G ); }

inline void lax_rhs( const std::array< std::vector< std::size_t >, 3 >& dsupedge,
                     const std::array< std::vector< real >, 3 >& dsupint,
                     const Coords& coord, const std::vector< std::size_t >& triinpoel,
                     const std::vector< std::uint8_t >& besym, const Fields& G, const Fields& U,
                     const std::vector< real >& v, real t, const std::vector< real >& tp, Fields& R )
{ lax::rhs( dsupedge, dsupint, coord, triinpoel, besym, G, U, v, t, tp, R ); }

inline real lax_refvel( real r, real p, real v ) { return lax::refvel( r, p, v ); }

inline void initialize( const Coords& coord, Fields& U, real t )
{ problems::initialize( coord, U, t, 0, {} ); }

inline void dirbc( Fields& U, real t, const Coords& coord, const std::vector< std::size_t >& m,
                   const std::vector< double >& val = {} )
{ physics::dirbc( 0, U, t, coord, {}, m, val ); }
inline void noslipbc( Fields& U, const std::vector< std::size_t >& n, std::size_t pos ) { physics::noslipbc( U, n, pos ); }
inline void phys_src( const Coords& coord, real t, Fields& U ) { if (auto s = problems::PHYS_SRC()) s( coord, t, U ); }

using SupEdge = std::array< std::vector< std::size_t >, 3 >;
using SupInt = std::array< std::vector< real >, 3 >;
inline void chorin_div( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                        real dt, const std::vector< real >& P, const Fields& G, const Fields& U,
                        std::vector< real >& D, bool stab ) { chorin::div( e, d, coord, tri, dt, P, G, U, D, stab ); }
inline void chorin_vgrad( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                          const Fields& U, Fields& G ) { chorin::vgrad( e, d, coord, tri, U, G ); }
inline void chorin_grad( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                         const std::vector< real >& U, Fields& G ) { chorin::grad( e, d, coord, tri, U, G ); }
inline void chorin_flux( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                         const Fields& U, const Fields& G, Fields& F ) { chorin::flux( e, d, coord, tri, U, G, F ); }
inline void chorin_rhs( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                        const std::vector< real >& v, real t, const std::vector< real >& P, const Fields& U,
                        const Fields& G, Fields& R ) { chorin::rhs( e, d, coord, tri, v, t, P, U, G, R ); }
inline void lohner_div( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                        const Fields& U, std::vector< real >& D, std::size_t pos ) { lohner::div( e, d, coord, tri, U, D, pos ); }
inline void lohner_grad( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                         const std::vector< real >& U, Fields& G ) { lohner::grad( e, d, coord, tri, U, G ); }
inline void lohner_vgrad( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                          const Fields& U, Fields& G ) { lohner::vgrad( e, d, coord, tri, U, G ); }
inline void lohner_flux( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                         const Fields& U, const Fields& G, Fields& F ) { lohner::flux( e, d, coord, tri, U, G, F ); }
inline void lohner_gradall( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                            const Fields& U, Fields& G ) { lohner::grad( e, d, coord, tri, U, G ); }
inline void lohner_rhs( const SupEdge& e, const SupInt& d, const Coords& coord, const std::vector< std::size_t >& tri,
                        const std::vector< real >& v, real t, const Fields& U, const Fields& G, Fields& R )
{ lohner::rhs( e, d, coord, tri, v, t, U, G, R ); }
inline void dirbcp( Fields& U, const Coords& coord, const std::vector< std::size_t >& m, const std::vector< double >& v )
{ physics::dirbcp( 0, U, coord, m, v ); }
inline port::PFn PRESSURE_RHS() { return problems::PRESSURE_RHS(); }
inline port::PFn PRESSURE_IC() { auto f = problems::PRESSURE_IC(); return [f]( real x, real y, real z ){ return f( x, y, z, 0 ); }; }
inline port::PFn PRESSURE_SOL() { auto f = problems::PRESSURE_SOL(); if (!f) return {};
  return [f]( real x, real y, real z ){ return f( x, y, z, 0 ); }; }
inline std::function< std::array< real, 3 >( real, real, real ) > PRESSURE_GRAD() { return problems::PRESSURE_GRAD(); }
inline void symbc( Fields& U, const std::vector< std::size_t >& n, const std::vector< real >& nn,
                   std::size_t pos ) { physics::symbc( U, n, nn, pos ); }
inline void farbc( Fields& U, const std::vector< std::size_t >& n, const std::vector< real >& nn )
{ physics::farbc( U, n, nn ); }
inline void prebc( Fields& U, const std::vector< std::size_t >& n, const std::vector< real >& v )
{ physics::prebc( U, n, v ); }

inline port::ICFn SOL() {
  auto s = problems::SOL();
  if (!s) return {};
  return [s]( real x, real y, real z, real t ){ return s( x, y, z, t, 0 ); };
}
real eos_pressure( real re );
real eos_soundspeed( real r, real p );

using LinkedList = std::pair< std::vector< std::size_t >, std::vector< std::size_t > >;
inline LinkedList genEsup( const std::vector< std::size_t >& inpoel, std::size_t nnpe )
{ return tk::genEsup( inpoel, nnpe ); }
inline LinkedList genPsup( const std::vector< std::size_t >& inpoel, std::size_t nnpe,
                           const LinkedList& esup ) { return tk::genPsup( inpoel, nnpe, esup ); }

#else

using Fields = PFields;
template< std::size_t N > using Hash = IdHash< N >;
template< std::size_t N > using Eq = IdEq< N >;
inline const char* name() { return "port"; }

inline void set_cfg( const Cfg& c ) { port::set_cfg( c ); }
using port::grad;
using port::rhs;
using port::zal_rhs;
using port::koz_rhs;
using port::lax_grad;
using port::lax_rhs;
using port::lax_refvel;
using port::initialize;
using port::dirbc;
using port::noslipbc;
using port::phys_src;
using port::chorin_div;
using port::chorin_vgrad;
using port::chorin_grad;
using port::chorin_flux;
using port::chorin_rhs;
using port::lohner_div;
using port::lohner_grad;
using port::lohner_vgrad;
using port::lohner_flux;
using port::lohner_gradall;
using port::lohner_rhs;
using port::dirbcp;
using port::PRESSURE_RHS;
using port::PRESSURE_IC;
using port::PRESSURE_SOL;
using port::PRESSURE_GRAD;
using port::symbc;
using port::farbc;
using port::prebc;
using port::SOL;
inline real eos_pressure( real re ) { return port::eos_pressure( re ); }
inline real eos_soundspeed( real r, real p ) { return port::eos_soundspeed( r, p ); }
using port::LinkedList;
using port::genEsup;
using port::genPsup;

#endif

} // be::
} // orc::

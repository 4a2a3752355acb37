// include/xyst_shim.hpp -- drop-in wrappers with the reference's own signatures over the C ABI.
//
// A maintainer of the reference builds this header inside the reference tree (it includes the
// reference's Fields.hpp / InciterConfig.hpp / Problems.hpp) and replaces, in the solver chares,
//     riemann::grad( ... )   by   xyst_shim::riemann::grad( ... )      (Inciter/RieCG.cpp:879)
//     riemann::rhs( ... )    by   xyst_shim::riemann::rhs( ... )       (Inciter/RieCG.cpp:948)
//     zalesak::rhs( ... )    by   xyst_shim::zalesak::rhs( ... )       (Inciter/ZalCG.cpp:999)
// Argument lists, ownership and results are those of src/Physics/Riemann.hpp:19-39 and
// src/Physics/Zalesak.hpp:19-30: caller-owned std::vector / tk::Fields, outputs overwritten with this
// chare's partial sums, configuration read from inciter::g_cfg, errors thrown as exceptions. Device
// state (one xyst_ctx per mesh chunk) is created on first use and found again by the address of the
// chunk's superedge array -- the chare owns that array for its lifetime.
//
// tests/test_gpu_shim.py calls these wrappers with real tk::Fields next to the reference's own
// functions (both compiled into oracle/_ref/liboracle_ref.so) and compares the results.
#pragma once
#include <array>
#include <vector>
#include <map>
#include <memory>
#include <string>
#include <stdexcept>
#include <cstdint>
#include "xyst_b200.h"
#include "Fields.hpp"
#include "InciterConfig.hpp"
#include "Problems.hpp"

namespace inciter { extern ctr::Config g_cfg; }

namespace xyst_shim {

using SupEdge = std::array< std::vector< std::size_t >, 3 >;
using SupInt = std::array< std::vector< tk::real >, 3 >;
using Coord = std::array< std::vector< tk::real >, 3 >;

inline void ck( int rc ) { if (rc) throw std::runtime_error( std::string( "xyst_b200: " ) + xyst_last_error() ); }

//! Device context of one mesh chunk, with what has been uploaded so far
struct Chunk {
  xyst_ctx* ctx = nullptr;
  std::size_t stride = 0, npoin = 0;
  std::string flux; double gamma = 0, stab2coef = 0; bool stab2 = false;
  const void* besym = nullptr; const void* v = nullptr;
  ~Chunk() { if (ctx) xyst_ctx_destroy( ctx ); }
};

inline std::map< const void*, std::unique_ptr< Chunk > >& chunks() { static std::map< const void*, std::unique_ptr< Chunk > > m; return m; }
//! Forget a chunk (a chare calls this from its destructor or after mesh refinement)
inline void release( const SupEdge& dsupedge ) { chunks().erase( dsupedge.data() ); }

inline Chunk& chunk( const SupEdge& dsupedge, const SupInt& dsupint, const Coord& coord,
                     const std::vector< std::size_t >& triinpoel, const std::vector< std::uint8_t >* besym,
                     const std::vector< tk::real >* v, std::size_t stride, std::size_t ncomp )
{
  using namespace inciter;
  auto& slot = chunks()[ dsupedge.data() ];
  const auto& flux = g_cfg.get< tag::flux >();
  auto gamma = g_cfg.get< tag::mat_spec_heat_ratio >();
  bool stab2 = g_cfg.get< tag::stab2 >(); auto stab2coef = g_cfg.get< tag::stab2coef >();
  auto npoin = coord[0].size();
  if (slot && (slot->stride != stride || slot->npoin != npoin || slot->flux != flux || slot->gamma != gamma ||
               slot->stab2 != stab2 || slot->stab2coef != stab2coef)) slot.reset();
  if (!slot) {
    slot.reset( new Chunk );
    xyst_params p{};
    p.ncomp = static_cast< int32_t >( ncomp );
    if (flux == "rusanov") p.flux = 0; else if (flux == "hllc") p.flux = 1; else throw std::runtime_error( "Flux not configured" );
    p.stab2 = stab2; p.exact_muscl = 1; p.gamma = gamma; p.stab2coef = stab2coef;
    int dev = 0;
    ck( xyst_ctx_create( dev, &p, &slot->ctx ) );
    const std::size_t nsup[3] = { dsupedge[0].size()/4, dsupedge[1].size()/3, dsupedge[2].size()/2 };
    const std::size_t* se[3] = { dsupedge[0].data(), dsupedge[1].data(), dsupedge[2].data() };
    const double* si[3] = { dsupint[0].data(), dsupint[1].data(), dsupint[2].data() };
    // riemann::grad returns the weak sums; the division by the nodal volume is the caller's
    // (RieCG.cpp:936-939): the library's fused "/ vol" runs with unit volumes
    std::vector< double > one( npoin, 1.0 );
    std::vector< std::uint8_t > nob( triinpoel.size(), 0 );
    auto up = stride == 4 ? xyst_zalcg_mesh_upload : xyst_mesh_upload;
    ck( up( slot->ctx, npoin, coord[0].data(), coord[1].data(), coord[2].data(), nsup, se, si,
            triinpoel.size()/3, triinpoel.data(), besym ? besym->data() : nob.data(), one.data(),
            v ? v->data() : one.data() ) );
    slot->stride = stride; slot->npoin = npoin; slot->flux = flux; slot->gamma = gamma; slot->stab2 = stab2;
    slot->stab2coef = stab2coef; slot->besym = besym ? besym->data() : nullptr; slot->v = v ? v->data() : nullptr;
  }
  // boundary symmetry flags and own nodal volumes arrive with the rhs call only
  if (besym && slot->besym != besym->data()) { ck( xyst_besym_upload( slot->ctx, besym->data() ) ); slot->besym = besym->data(); }
  if (v && slot->v != v->data()) { ck( xyst_v_upload( slot->ctx, v->data() ) ); slot->v = v->data(); }
  return *slot;
}

//! problems::SRC() evaluated as riemann::src does (Riemann.cpp:880-907), handed to the library
inline void source( Chunk& k, const Coord& coord, tk::real t, const std::vector< tk::real >& tp, std::size_t ncomp )
{
  auto src = problems::SRC();
  if (!src) { ck( xyst_src_upload( k.ctx, nullptr ) ); return; }
  std::vector< double > S( k.npoin*ncomp, 0.0 );
  for (std::size_t p=0; p<k.npoin; ++p) {
    if (inciter::g_cfg.get< tag::steady >()) t = tp[p];
    auto s = src( coord[0][p], coord[1][p], coord[2][p], t, /*meshid=*/0 );
    for (std::size_t c=0; c<s.size() && c<ncomp; ++c) S[p*ncomp+c] = s[c];
  }
  ck( xyst_src_upload( k.ctx, S.data() ) );
}

namespace riemann {

//! src/Physics/Riemann.hpp:19-26
inline void grad( const SupEdge& dsupedge, const SupInt& dsupint, const Coord& coord,
                  const std::vector< std::size_t >& triinpoel, const tk::Fields& U, tk::Fields& G )
{
  auto& k = chunk( dsupedge, dsupint, coord, triinpoel, nullptr, nullptr, 3, U.nprop() );
  ck( xyst_state_set( k.ctx, U.vec().data() ) );
  ck( xyst_riecg_grad( k.ctx ) );
  ck( xyst_grad_get( k.ctx, G.vec().data() ) );
}

//! src/Physics/Riemann.hpp:28-39
inline void rhs( const SupEdge& dsupedge, const SupInt& dsupint, const Coord& coord,
                 const std::vector< std::size_t >& triinpoel, const std::vector< std::uint8_t >& besym,
                 const tk::Fields& G, const tk::Fields& U, const std::vector< tk::real >& v,
                 tk::real t, const std::vector< tk::real >& tp, tk::Fields& R )
{
  auto& k = chunk( dsupedge, dsupint, coord, triinpoel, &besym, &v, 3, U.nprop() );
  source( k, coord, t, tp, U.nprop() );
  ck( xyst_state_set( k.ctx, U.vec().data() ) );
  ck( xyst_grad_set( k.ctx, G.vec().data() ) );
  ck( xyst_riecg_rhs( k.ctx ) );
  ck( xyst_rhs_get( k.ctx, R.vec().data() ) );
}

} // riemann::

namespace zalesak {

//! src/Physics/Zalesak.hpp:19-30 (time-accurate runs; dt = the global time step)
inline void rhs( const SupEdge& dsupedge, const SupInt& dsupint, const Coord& coord,
                 const std::vector< std::size_t >& triinpoel, const std::vector< std::uint8_t >& besym,
                 tk::real /*t*/, tk::real dt, const std::vector< tk::real >& /*tp*/,
                 const std::vector< tk::real >& /*dtp*/, const tk::Fields& U, tk::Fields& R )
{
  if (inciter::g_cfg.get< tag::steady >()) throw std::runtime_error( "xyst_shim::zalesak::rhs: local time stepping goes through xyst_steady()" );
  auto& k = chunk( dsupedge, dsupint, coord, triinpoel, &besym, nullptr, 4, U.nprop() );
  ck( xyst_state_set( k.ctx, U.vec().data() ) );
  ck( xyst_zalcg_rhs( k.ctx, dt ) );
  ck( xyst_rhs_get( k.ctx, R.vec().data() ) );
}

} // zalesak::

} // xyst_shim::

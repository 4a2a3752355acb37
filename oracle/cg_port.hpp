// oracle/cg_port.hpp -- TEST INFRASTRUCTURE ONLY (never linked by the product).
//
// Serial restatement of the reference's linear-solver hot path used by the ChoCG/LohCG
// pressure projection:
//   tk::CSR            src/LinearSolver/CSR.cpp:19-172 (ctor from psup, operator(), dirichlet, mult)
//   ConjugateGradients src/LinearSolver/ConjugateGradients.cpp:105-823 (setup, dot, pc, initres,
//                      next, qAp, q, pq, rz, x) -- the Charm++ chare array becomes a vector of
//                      partitions stepped in lock-step with their shared-row sums / averages
//                      done through memory
//   tk::count / slave  src/Mesh/Reorder.cpp:379-402
// With -DORACLE_REF the matrix class is the reference's own tk::CSR (CSR.cpp compiled in place).
// Pinned by the reference's unit-test known answers (tests/unit/LinearSolver/TestCSR.cpp:297-364,
// TestConjugateGradients.cpp:216-217,290-291,457-458,531-532) in tests/test_oracle_cg.py.
#pragma once
#include <vector>
#include <cstdlib>
#include <map>
#include <unordered_map>
#include <unordered_set>
#include <cmath>
#include <string>
#include <stdexcept>
#include <limits>
#include <algorithm>
#include <memory>
#ifdef ORACLE_REF
  #include "CSR.hpp"
#endif

namespace orc {
namespace cg {

using real = double;
using Psup = std::pair< std::vector< std::size_t >, std::vector< std::size_t > >;
using CommMap = std::unordered_map< int, std::unordered_set< std::size_t > >;

//! Block CSR with 1-based ia/ja, full rows, columns ascending (CSR.cpp:19-84)
class PCSR {
  public:
    PCSR() = default;
    PCSR( std::size_t nc, const Psup& psup ) : ncomp( nc ), rnz( psup.second.size()-1 ), ia( rnz.size()*nc+1 ) {
      const auto& psup1 = psup.first; const auto& psup2 = psup.second;
      std::size_t nnz = 0;
      ia[0] = 1;
      for (std::size_t i=0; i<psup2.size()-1; ++i) {
        rnz[i] = 1 + (psup2[i+1] - psup2[i]);
        nnz += rnz[i] * ncomp;
        for (std::size_t k=0; k<ncomp; ++k) ia[i*ncomp+k+1] = ia[i*ncomp+k] + rnz[i];
      }
      a.resize( nnz, 0.0 ); ja.resize( nnz );
      for (std::size_t i=0; i<rnz.size(); ++i)
        for (std::size_t k=0; k<ncomp; ++k) {
          auto itmp = i*ncomp+k;
          ja[ia[itmp]-1] = itmp+1;
          for (std::size_t n=1, j=psup2[i]+1; j<=psup2[i+1]; ++j) ja[ia[itmp]-1+(n++)] = psup1[j]*ncomp+k+1;
          std::sort( ja.begin()+static_cast<long>(ia[itmp]-1), ja.begin()+static_cast<long>(ia[itmp+1]-1) );   // :72-78
        }
    }
    real& operator()( std::size_t row, std::size_t col, std::size_t pos=0 ) {          // :86-104
      auto rncomp = row * ncomp;
      for (std::size_t j=ia[rncomp+pos]-1; j<ia[rncomp+pos+1]-1; ++j) if (col*ncomp+pos+1 == ja[j]) return a[j];
      throw std::runtime_error( "Sparse matrix index not found" );
    }
    void dirichlet( std::size_t i, real val, std::vector< real >& b, const std::vector< std::size_t >& gid,
                    const CommMap& nodecommap, std::size_t pos ) {                      // :106-152
      auto incomp = i * ncomp;
      for (std::size_t r=0; r<rnz.size()*ncomp; ++r)
        for (std::size_t j=ia[r]-1; j<ia[r+1]-1; ++j)
          if (incomp+pos+1 == ja[j]) { b[r] += a[j] * val; a[j] = 0.0; break; }
      real cnt = 1.0;
      if (!nodecommap.empty()) for (const auto& s : nodecommap) if (s.second.count( gid[i] )) cnt += 1.0;
      auto diag = nodecommap.empty() ? 1.0 : 1.0/cnt;
      for (std::size_t j=ia[incomp+pos]-1; j<ia[incomp+pos+1]-1; ++j) a[j] = (incomp+pos+1 == ja[j]) ? diag : 0.0;
    }
    void mult( const std::vector< real >& x, std::vector< real >& r ) const {           // :154-172
      std::fill( r.begin(), r.end(), 0.0 );
      for (std::size_t i=0; i<rnz.size()*ncomp; ++i)
        for (std::size_t j=ia[i]-1; j<ia[i+1]-1; ++j) r[i] += a[j] * x[ja[j]-1];
    }
    void zero() { std::fill( a.begin(), a.end(), 0.0 ); }                                // CSR.hpp:63
    std::size_t Ncomp() const { return ncomp; }
    std::size_t rsize() const { return rnz.size()*ncomp; }
    const std::vector< std::size_t >& IA() const { return ia; }
    const std::vector< std::size_t >& JA() const { return ja; }
    const std::vector< real >& Avals() const { return a; }
  private:
    std::size_t ncomp = 1;
    std::vector< std::size_t > rnz, ia, ja;
    std::vector< real > a;
};

#ifdef ORACLE_REF
//! the reference's own matrix class behind the same small interface
class RCSR {
  public:
    RCSR() : m( 1, Psup{ {0}, {0,0} } ) {}
    RCSR( std::size_t nc, const Psup& psup ) : m( nc, psup ), nrow( (psup.second.size()-1)*nc ) {}
    real& operator()( std::size_t row, std::size_t col, std::size_t pos=0 ) { return m( row, col, pos ); }
    void dirichlet( std::size_t i, real val, std::vector< real >& b, const std::vector< std::size_t >& gid,
                    const CommMap& c, std::size_t pos ) { m.dirichlet( i, val, b, gid, c, pos ); }
    void mult( const std::vector< real >& x, std::vector< real >& r ) const { m.mult( x, r ); }
    void zero() { m.zero(); }
    std::size_t Ncomp() const { return m.Ncomp(); }
    std::size_t rsize() const { return nrow; }
    //! structure and values through the reference's own writer-free accessors: rebuilt with
    //! the port class (identical structure, checked in tests) and filled via operator()
    tk::CSR m;
    std::size_t nrow = 0;
};
using Matrix = RCSR;
inline const char* matrix_backend() { return "reference"; }
#else
using Matrix = PCSR;
inline const char* matrix_backend() { return "port"; }
#endif

inline real count( const CommMap& map, std::size_t node ) {                               // Reorder.cpp:379-386
  real c = 1.0;
  for (const auto& s : map) if (s.second.count( node )) c += 1.0;
  return c;
}
inline bool slave( const CommMap& map, std::size_t node, int chare ) {                    // Reorder.cpp:393-402
  for (const auto& s : map) if (s.first < chare && s.second.count( node )) return true;
  return false;
}

//! One partition of the CG chare array
struct Part {
  Matrix A;
  PCSR S;                                  // structure twin (ia/ja) for export, same ctor input
  std::vector< real > x, b, r, p, q, z, d;
  std::vector< std::size_t > gid;
  std::unordered_map< std::size_t, std::size_t > lid;
  CommMap nodeCommMap;
  int index = 0;
  // boundary conditions of the current solve (ConjugateGradients::m_dirbc, m_An)
  std::map< std::size_t, std::vector< std::pair< int, real > > > dirbc;
  Matrix An;
};

//! Dirichlet (local node -> {set?, value} per component) and Neumann data of one partition
struct BCs {
  std::unordered_map< std::size_t, std::vector< std::pair< int, real > > > dirbc;
  std::vector< real > neubc;
};

class Solver {
  public:
    std::vector< std::unique_ptr< Part > > parts;
    std::string pc = "none";
    real normb = 0, rho = 0, rho0 = 0, alpha = 0, normr = 0;
    std::size_t it = 0;
    bool converged = false, finished = false;

    std::size_t add( std::size_t ncomp, const Psup& psup, const std::vector< std::size_t >& gid, const CommMap& cm ) {
      auto p = std::make_unique< Part >();
      p->A = Matrix( ncomp, psup ); p->S = PCSR( ncomp, psup );
      auto n = gid.size()*ncomp;
      for (auto* v : { &p->x, &p->b, &p->r, &p->p, &p->q, &p->z, &p->d }) v->assign( n, 0.0 );
      p->gid = gid; p->nodeCommMap = cm; p->index = static_cast< int >( parts.size() );
      for (std::size_t i=0; i<gid.size(); ++i) p->lid[gid[i]] = i;
      parts.push_back( std::move(p) );
      return parts.size()-1;
    }

    //! ConjugateGradients::dot :128-151 summed over all partitions (contribute(sum_double))
    //! ORACLE_DOT_REVERSE=1 (tests only): the same sum taken from the last node to the first -- a
    //! rounding-level perturbation of every dot product, to measure how far the iteration carries it
    real dot( std::vector< real > Part::*a, std::vector< real > Part::*b ) const {
      static const bool reverse = [](){ const char* e = std::getenv( "ORACLE_DOT_REVERSE" ); return e && e[0] == '1'; }();
      real D = 0.0;
      for (const auto& pp : parts) {
        const auto& P = *pp; auto ncomp = P.A.Ncomp();
        real d = 0.0;
        auto n = (P.*a).size()/ncomp;
        for (std::size_t k=0; k<n; ++k) {
          auto i = reverse ? n-1-k : k;
          if (!slave( P.nodeCommMap, P.gid[i], P.index ))
            for (std::size_t c=0; c<ncomp; ++c) d += (P.*a)[i*ncomp+c] * (P.*b)[i*ncomp+c];
        }
        D += d;
      }
      return D;
    }

    //! sum a nodal vector over the partitions sharing each node (comres/comq/comd)
    void halosum( std::vector< real > Part::*v ) {
      std::vector< std::unordered_map< std::size_t, std::vector< real > > > recv( parts.size() );
      for (auto& pp : parts) {
        auto& P = *pp; auto ncomp = P.A.Ncomp();
        for (const auto& [c,n] : P.nodeCommMap)
          for (auto g : n) {
            auto i = P.lid.at( g );
            auto& acc = recv[static_cast<std::size_t>(c)][g];
            if (acc.empty()) acc.assign( ncomp, 0.0 );
            for (std::size_t k=0; k<ncomp; ++k) acc[k] += (P.*v)[i*ncomp+k];
          }
      }
      for (std::size_t c=0; c<parts.size(); ++c) {
        auto& P = *parts[c]; auto ncomp = P.A.Ncomp();
        for (const auto& [g,val] : recv[c]) { auto i = P.lid.at( g ); for (std::size_t k=0; k<ncomp; ++k) (P.*v)[i*ncomp+k] += val[k]; }
      }
    }

    bool applied = false;                  // m_apply: matrix is restored after the solve

    //! init :336-428, combc :430-448, apply :451-505, comr/r :508-556
    //! b: complete right hand sides per partition (empty: keep)
    void init( const std::vector< std::vector< real > >& b, const std::vector< BCs >& bcs, bool apply, const std::string& pc_ ) {
      pc = pc_;
      for (std::size_t k=0; k<parts.size(); ++k) if (k < b.size() && !b[k].empty()) parts[k]->b = b[k];
      applied = apply;
      if (!apply) { setup(); return; }
      for (std::size_t k=0; k<parts.size(); ++k) {
        auto& P = *parts[k];
        P.An = P.A;
        if (!bcs[k].neubc.empty()) P.q = bcs[k].neubc; else std::fill( P.q.begin(), P.q.end(), 0.0 );
        P.dirbc.clear();
        for (const auto& [i,bc] : bcs[k].dirbc) P.dirbc[i] = bc;
      }
      // combc: Dirichlet BCs and Neumann contributions at shared nodes go to the sharers
      std::vector< std::map< std::size_t, std::vector< std::pair< int, real > > > > dirbcc( parts.size() );
      std::vector< std::unordered_map< std::size_t, std::vector< real > > > qc( parts.size() );
      for (auto& pp : parts) {
        auto& P = *pp; auto ncomp = P.A.Ncomp();
        for (const auto& [c,n] : P.nodeCommMap)
          for (auto g : n) {
            auto i = P.lid.at( g );
            auto j = P.dirbc.find( i );
            if (j != P.dirbc.end()) dirbcc[static_cast<std::size_t>(c)][g] = j->second;
            auto& acc = qc[static_cast<std::size_t>(c)][g];
            if (acc.empty()) acc.assign( ncomp, 0.0 );
            for (std::size_t d=0; d<ncomp; ++d) acc[d] += P.q[i*ncomp+d];
          }
      }
      for (std::size_t k=0; k<parts.size(); ++k) {
        auto& P = *parts[k]; auto ncomp = P.A.Ncomp();
        for (const auto& [g,bc] : dirbcc[k]) P.dirbc[ P.lid.at(g) ] = bc;
        for (const auto& [g,q] : qc[k]) { auto i = P.lid.at( g ); for (std::size_t c=0; c<ncomp; ++c) P.q[i*ncomp+c] += q[c]; }
        for (std::size_t i=0; i<P.b.size(); ++i) P.b[i] += P.q[i];
        std::fill( P.r.begin(), P.r.end(), 0.0 );
        for (auto bi = P.dirbc.rbegin(); bi != P.dirbc.rend(); ++bi)
          for (std::size_t j=0; j<ncomp; ++j)
            if (bi->second[j].first) P.A.dirichlet( bi->first, bi->second[j].second, P.r, P.gid, P.nodeCommMap, j );
      }
      halosum( &Part::r );                                               // comr
      for (auto& pp : parts) {
        auto& P = *pp; auto ncomp = P.A.Ncomp();
        for (std::size_t i=0; i<P.b.size(); ++i) P.b[i] -= P.r[i];
        for (const auto& [i,bc] : P.dirbc)
          for (std::size_t j=0; j<ncomp; ++j) if (bc[j].first) P.b[i*ncomp+j] = bc[j].second;
      }
      setup();
    }

    //! setup :105-126, residual :164-190, pc :213-259, initres :280-319
    real setup() {
      converged = false; finished = false;
      for (auto& pp : parts) {
        auto& P = *pp; auto ncomp = P.A.Ncomp();
        P.A.mult( P.x, P.r );                                          // residual(): r = A x (own part)
        if (pc == "none") { for (std::size_t i=0; i<P.q.size()/ncomp; ++i) { auto c = count( P.nodeCommMap, P.gid[i] ); for (std::size_t k=0; k<ncomp; ++k) P.q[i*ncomp+k] = 1.0 / c; } }
        else if (pc == "jacobi") { for (std::size_t i=0; i<P.q.size()/ncomp; ++i) for (std::size_t k=0; k<ncomp; ++k) P.q[i*ncomp+k] = P.A( i, i, k ); }
        else throw std::runtime_error( "unknown preconditioner" );
      }
      halosum( &Part::r ); halosum( &Part::q );
      normb = std::sqrt( dot( &Part::b, &Part::b ) );
      for (auto& pp : parts) {
        auto& P = *pp;
        for (auto& rr : P.r) rr *= -1.0;
        for (std::size_t i=0; i<P.r.size(); ++i) P.r[i] += P.b[i];
        P.p = P.r;
        P.d = P.q;
        for (std::size_t i=0; i<P.z.size(); ++i) P.z[i] = P.r[i] / P.d[i];
      }
      rho = dot( &Part::r, &Part::z );
      return normb;
    }

    //! solve :558-582 + next/qAp/q/pq/rz/x :584-823
    real solve( std::size_t maxit, real tol ) {
      it = 0;
      real nr = std::sqrt( normr );
      if (converged) return nr;
      for (;;) {
        alpha = it == 0 ? 0.0 : rho/rho0;
        rho0 = rho;
        for (auto& pp : parts) { auto& P = *pp; for (std::size_t i=0; i<P.p.size(); ++i) P.p[i] = P.z[i] + alpha * P.p[i]; }
        for (auto& pp : parts) pp->A.mult( pp->p, pp->q );
        halosum( &Part::q );
        auto d = dot( &Part::p, &Part::q );
        const auto eps = std::numeric_limits< real >::epsilon();
        if (std::abs(d) < eps) { finished = true; alpha = 0.0; } else alpha = rho / d;
        for (auto& pp : parts) {
          auto& P = *pp;
          for (std::size_t i=0; i<P.r.size(); ++i) P.r[i] -= alpha * P.q[i];
          for (std::size_t i=0; i<P.z.size(); ++i) P.z[i] = P.r[i] / P.d[i];
        }
        auto rz = dot( &Part::r, &Part::z );
        normr = dot( &Part::r, &Part::r );
        rho = rz;
        for (auto& pp : parts) { auto& P = *pp; for (std::size_t i=0; i<P.x.size(); ++i) P.x[i] += alpha * P.p[i]; }
        // x :772-823: shared nodes: sum of all sharers' values divided by the count
        {
          std::vector< std::unordered_map< std::size_t, std::vector< real > > > recv( parts.size() );
          for (auto& pp : parts) { auto& P = *pp; auto ncomp = P.A.Ncomp();
            for (const auto& [c,n] : P.nodeCommMap) for (auto g : n) {
              auto i = P.lid.at( g ); auto& acc = recv[static_cast<std::size_t>(c)][g];
              if (acc.empty()) acc.assign( ncomp, 0.0 );
              for (std::size_t k=0; k<ncomp; ++k) acc[k] += P.x[i*ncomp+k]; } }
          for (std::size_t c=0; c<parts.size(); ++c) { auto& P = *parts[c]; auto ncomp = P.A.Ncomp();
            for (const auto& [g,val] : recv[c]) { auto i = P.lid.at( g );
              for (std::size_t k=0; k<ncomp; ++k) P.x[i*ncomp+k] += val[k];
              auto cnt = count( P.nodeCommMap, g );
              for (std::size_t k=0; k<ncomp; ++k) P.x[i*ncomp+k] /= cnt; } }
        }
        ++it;
        auto nb = normb > 1.0e-14 ? normb : 1.0;
        nr = std::sqrt( normr );
        if (finished || nr < tol*nb || it >= maxit) {
          converged = !(nr > tol*nb);
          if (applied) for (auto& pp : parts) pp->A = pp->An;          // restore, :809
          return nr;
        }
      }
    }
};

//! Laplacian A(a,b) += J/6 grad_a . grad_b, as in the reference's unit tests
//! (tests/unit/LinearSolver/TestConjugateGradients.cpp:170-195) for all ncomp positions
inline void laplacian( Matrix& A, const std::vector< std::size_t >& inpoel,
                       const std::vector< real >& X, const std::vector< real >& Y, const std::vector< real >& Z )
{
  for (std::size_t e=0; e<inpoel.size()/4; ++e) {
    const std::size_t N[4] = { inpoel[e*4+0], inpoel[e*4+1], inpoel[e*4+2], inpoel[e*4+3] };
    real ba[3] = { X[N[1]]-X[N[0]], Y[N[1]]-Y[N[0]], Z[N[1]]-Z[N[0]] },
         ca[3] = { X[N[2]]-X[N[0]], Y[N[2]]-Y[N[0]], Z[N[2]]-Z[N[0]] },
         da[3] = { X[N[3]]-X[N[0]], Y[N[3]]-Y[N[0]], Z[N[3]]-Z[N[0]] };
    auto cross = []( const real a[3], const real b[3], real r[3] ){ r[0] = a[1]*b[2] - b[1]*a[2]; r[1] = a[2]*b[0] - b[2]*a[0]; r[2] = a[0]*b[1] - b[0]*a[1]; };
    real c[3]; cross( ca, da, c );
    const auto J = ba[0]*c[0] + ba[1]*c[1] + ba[2]*c[2];
    real grad[4][3];
    auto crossdiv = [&]( const real a[3], const real b[3], real r[3] ){
      r[0] = (a[1]*b[2] - b[1]*a[2]) / J; r[1] = (a[2]*b[0] - b[2]*a[0]) / J; r[2] = (a[0]*b[1] - b[0]*a[1]) / J; };
    crossdiv( ca, da, grad[1] ); crossdiv( da, ba, grad[2] ); crossdiv( ba, ca, grad[3] );
    for (std::size_t i=0; i<3; ++i) grad[0][i] = -grad[1][i]-grad[2][i]-grad[3][i];
    for (std::size_t a=0; a<4; ++a)
      for (std::size_t k=0; k<3; ++k)
        for (std::size_t b=0; b<4; ++b)
          for (std::size_t pos=0; pos<A.Ncomp(); ++pos)
            A( N[a], N[b], pos ) += J/6 * grad[a][k] * grad[b][k];
  }
}

} // cg::
} // orc::

// oracle/capi.cpp -- TEST INFRASTRUCTURE ONLY (never linked by the product).
// C entry points (ctypes) around the serial oracle driver; see oracle/oracle.h.
#include <cstring>
#include <string>
#include "driver.hpp"
#include "cg_port.hpp"
#include "chocg_port.hpp"
#include "lohcg_port.hpp"
#include "oracle.h"

using namespace orc;

namespace {
thread_local std::string g_err;

struct Handle { MeshInput in; std::unique_ptr< Run > run; };

Cfg to_cfg( const orc_cfg* c ) {
  Cfg k;
  k.problem = c->problem; k.flux = c->flux; k.ncomp = static_cast< std::size_t >( c->ncomp );
  k.soundspeed = c->soundspeed;
  if (c->src_radius > 0.0) { k.src_location = {{ c->src_location[0], c->src_location[1], c->src_location[2] }};
    k.src_radius = c->src_radius; k.src_release_time = c->src_release_time; }
  if (c->freezeflow != 0.0) k.freezeflow = c->freezeflow;
  k.freezetime = c->freezetime;
  if (c->fctfreeze != 0.0) k.fctfreeze = c->fctfreeze;
  k.theta = c->theta; k.mom_iter = c->mom_iter ? c->mom_iter : 10; k.mom_tol = c->mom_tol; if (c->mom_pc[0]) k.mom_pc = c->mom_pc;
  k.alpha = c->alpha; k.kappa = c->kappa; k.r0 = c->r0; k.ce = c->ce; k.beta = {{ c->beta[0], c->beta[1], c->beta[2] }};
  k.gamma = c->gamma; k.p0 = c->p0; k.cfl = c->cfl; k.dt = c->dt; k.t0 = c->t0; k.term = c->term;
  k.nstep = c->nstep; k.stab2 = c->stab2 != 0; k.stab2coef = c->stab2coef; k.steady = c->steady != 0;
  k.diag_iter = c->diag_iter ? c->diag_iter : 1;
  for (int i=0; i<c->nsym; ++i) k.bc_sym.push_back( c->sym[i] );
  for (int i=0; i<c->ndir; ++i) {
    std::vector< int > m( k.ncomp+1 );
    for (std::size_t j=0; j<k.ncomp+1; ++j) m[j] = c->dir[i][j];
    k.bc_dir.push_back( m );
  }
  for (int i=0; i<c->nfar; ++i) k.bc_far.push_back( c->far_sets[i] );
  k.far_density = c->far_density; k.far_pressure = c->far_pressure;
  k.far_velocity = {{ c->far_velocity[0], c->far_velocity[1], c->far_velocity[2] }};
  for (int i=0; i<c->npre; ++i) {
    k.bc_pre.push_back( { c->pre_sets[i] } );
    k.pre_density.push_back( c->pre_density[i] );
    k.pre_pressure.push_back( c->pre_pressure[i] );
  }
  for (int i=0; i<c->nfieldout; ++i) k.fieldout_sets.push_back( c->fieldout_sets[i] );
  if (c->solver[0]) k.solver = c->solver;
  k.fct = c->fct != 0; k.fctclip = c->fctclip != 0; k.fctdif = c->fctdif;
  for (int i=0; i<c->nfctsys; ++i) k.fctsys.push_back( static_cast< std::uint64_t >( c->fctsys[i] ) );
  if (c->rgas != 0.0) k.rgas = c->rgas;
  k.turkel = c->turkel; k.velinf = {{ c->velinf[0], c->velinf[1], c->velinf[2] }};
  k.residual = c->residual; k.rescomp = c->rescomp ? c->rescomp : 1;
  k.ic_density = c->ic_density; k.ic_pressure = c->ic_pressure;
  k.ic_velocity = {{ c->ic_velocity[0], c->ic_velocity[1], c->ic_velocity[2] }};
  k.mu = c->mu; k.dif = c->dif; k.stab = c->stab != 0; k.rk = c->rk ? c->rk : 1;
  for (int i=0; i<c->nnoslip; ++i) k.bc_noslip.push_back( c->noslip[i] );
  for (int i=0; i<c->ndirval; ++i) { std::vector< real > v( k.ncomp+1 ); for (std::size_t j=0; j<k.ncomp+1; ++j) v[j] = c->dirval[i][j]; k.bc_dirval.push_back( v ); }
  k.p_iter = c->p_iter ? c->p_iter : 10; k.p_tol = c->p_tol; if (c->p_pc[0]) k.p_pc = c->p_pc;
  for (int i=0; i<c->np_dir; ++i) k.p_bc_dir.push_back( { c->p_dir[i][0], c->p_dir[i][1] } );
  for (int i=0; i<c->np_dirval; ++i) k.p_bc_dirval.push_back( { c->p_dirval[i][0], c->p_dirval[i][1] } );
  for (int i=0; i<c->np_sym; ++i) k.p_bc_sym.push_back( c->p_sym[i] );
  k.p_hydrostat = c->p_hydrostat_set ? c->p_hydrostat : ~0ULL;
  return k;
}

template< class T >
std::size_t put( const std::vector< T >& v, void* out, std::size_t cap ) {
  auto bytes = v.size()*sizeof(T);
  if (out && cap >= bytes && bytes) std::memcpy( out, v.data(), bytes );
  return bytes;
}
std::size_t putf( const be::Fields& f, void* out, std::size_t cap ) {
  std::vector< real > v( f.nunk()*f.nprop() );
  for (std::size_t i=0; i<f.nunk(); ++i) for (std::size_t c=0; c<f.nprop(); ++c) v[i*f.nprop()+c] = f(i,c);
  return put( v, out, cap );
}
}

#ifdef ORACLE_SHIM
#include "xyst_shim.hpp"
#endif

extern "C" {

const char* orc_backend() { return be::name(); }
const char* orc_last_error() { return g_err.c_str(); }

void* orc_create( std::size_t npoin, const double* x, const double* y, const double* z,
                  std::size_t ntet, const std::uint64_t* tets,
                  std::size_t ntri, const std::uint64_t* tris,
                  int nblocks, const int* block_type, const std::uint64_t* block_n,
                  int nsets, const int* set_id, const std::uint64_t* set_off,
                  const std::uint64_t* set_elem, const std::uint64_t* set_side,
                  const orc_cfg* cfg, int nchare, const std::uint64_t* target )
{
  try {
    auto h = std::make_unique< Handle >();
    auto& in = h->in;
    in.coord[0].assign( x, x+npoin ); in.coord[1].assign( y, y+npoin ); in.coord[2].assign( z, z+npoin );
    in.tets.assign( tets, tets+ntet*4 );
    if (ntri) in.tris.assign( tris, tris+ntri*3 );
    for (int b=0; b<nblocks; ++b) in.blocks.emplace_back( block_type[b], block_n[b] );
    for (int s=0; s<nsets; ++s) {
      in.ss_elem[ set_id[s] ].assign( set_elem+set_off[s], set_elem+set_off[s+1] );
      in.ss_side[ set_id[s] ].assign( set_side+set_off[s], set_side+set_off[s+1] );
    }
    std::vector< std::size_t > tg( ntet, 0 );
    if (target) for (std::size_t e=0; e<ntet; ++e) tg[e] = target[e];
    auto k = to_cfg( cfg );
    if (k.solver == "chocg") h->run.reset( new ChoRun( in, k, tg, nchare ) );
    else if (k.solver == "lohcg") h->run.reset( new LohRun( in, k, tg, nchare ) );
    else h->run.reset( new Run( in, k, tg, nchare ) );
    return h.release();
  } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

void orc_destroy( void* h ) { delete static_cast< Handle* >( h ); }

int orc_step( void* hv, int nsteps )
{
  auto h = static_cast< Handle* >( hv );
  try {
    be::set_cfg( h->run->cfg );
    int n = 0;
    while (n < nsteps && !h->run->finished) { h->run->step(); ++n; }
    return n;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

std::size_t orc_ndiag( void* hv ) { return static_cast< Handle* >( hv )->run->diagrows.size(); }

std::size_t orc_diagrow( void* hv, std::size_t i, double* out, std::size_t cap )
{
  const auto& r = static_cast< Handle* >( hv )->run->diagrows.at( i );
  if (out && cap >= r.size()) std::memcpy( out, r.data(), r.size()*sizeof(double) );
  return r.size();
}

double orc_scalar( void* hv, const char* name )
{
  auto& r = *static_cast< Handle* >( hv )->run;
  std::string n( name );
  if (n == "t") return r.t;
  if (n == "dt") return r.dt;
  if (n == "it") return static_cast< double >( r.it );
  if (n == "meshvol") return r.meshvol;
  if (n == "nchare") return static_cast< double >( r.ch.size() );
  if (n == "finished") return r.finished ? 1.0 : 0.0;
  if (n == "mit") { if (auto c = dynamic_cast< ChoRun* >( &r )) return static_cast< double >( c->mit ); }
  if (n == "pit") { if (auto c = dynamic_cast< ChoRun* >( &r )) return static_cast< double >( c->pit );
                    if (auto c = dynamic_cast< LohRun* >( &r )) return static_cast< double >( c->pit ); }
  return std::nan("");
}

std::size_t orc_get( void* hv, int chare, const char* name, void* out, std::size_t cap )
{
  auto h = static_cast< Handle* >( hv );
  auto& c = *h->run->ch.at( static_cast< std::size_t >( chare ) );
  std::string n( name );
  if (n == "gid") return put( c.gid, out, cap );
  if (n == "inpoel") return put( c.inpoel, out, cap );
  if (n == "pr") return put( c.pr, out, cap );
  if (n == "div") return put( c.div, out, cap );
  if (n == "sgrad") return putf( c.sgrad, out, cap );
  if (n == "pgrad") return putf( c.pgrad, out, cap );
  if (n == "mflux") return putf( c.mflux, out, cap );
  if (n == "dirbcmaskp") return put( c.dirbcmaskp, out, cap );
  if (n == "dirbcval") return put( c.dirbcval, out, cap );
  if (n == "dirbcvalp") return put( c.dirbcvalp, out, cap );
  if (n == "noslipbcnodes") return put( c.noslipbcnodes, out, cap );
  if (n == "dp") { if (auto cr = dynamic_cast< ChoRun* >( h->run.get() )) return put( cr->cgpre.parts[static_cast<std::size_t>(chare)]->x, out, cap ); }
  if (n == "u0" || n == "p_ic" || n == "p_sol" || n == "p_rhs" || n == "u_sol" || n == "neubc" || n == "hydrostat" || n == "plhs_a" || n == "mlhs_a") {
    if (auto cr = dynamic_cast< ChoRun* >( h->run.get() )) { be::set_cfg( cr->cfg ); return put( cr->exported( static_cast<std::size_t>(chare), n ), out, cap ); } }
  if (n == "plhs_ia") { if (auto cr = dynamic_cast< ChoRun* >( h->run.get() )) return put( cr->cgpre.parts[static_cast<std::size_t>(chare)]->S.IA(), out, cap ); }
  if (n == "plhs_ja") { if (auto cr = dynamic_cast< ChoRun* >( h->run.get() )) return put( cr->cgpre.parts[static_cast<std::size_t>(chare)]->S.JA(), out, cap ); }
  if (auto lr = dynamic_cast< LohRun* >( h->run.get() )) {          // LohCG: the assembled pressure matrix
    auto& P = *lr->cgpre.parts[static_cast<std::size_t>(chare)];
    if (n == "plhs_ia") return put( P.S.IA(), out, cap );
    if (n == "plhs_ja") return put( P.S.JA(), out, cap );
    if (n == "dp") return put( P.x, out, cap );
    if (n == "plhs_a") { std::vector< real > r; const auto& ia = P.S.IA(); const auto& ja = P.S.JA();
      for (std::size_t row=0; row+1<ia.size(); ++row) for (std::size_t j=ia[row]-1; j<ia[row+1]-1; ++j) r.push_back( P.A( row, ja[j]-1 ) );
      return put( r, out, cap ); }
  }
  if (n == "x") return put( c.coord[0], out, cap );
  if (n == "y") return put( c.coord[1], out, cap );
  if (n == "z") return put( c.coord[2], out, cap );
  if (n == "vol") return put( c.vol, out, cap );
  if (n == "v") return put( c.v, out, cap );
  if (n == "triinpoel") return put( c.triinpoel, out, cap );
  if (n == "besym") return put( c.besym, out, cap );
  if (n == "dsupedge0") return put( c.dsupedge[0], out, cap );
  if (n == "dsupedge1") return put( c.dsupedge[1], out, cap );
  if (n == "dsupedge2") return put( c.dsupedge[2], out, cap );
  if (n == "dsupint0") return put( c.dsupint[0], out, cap );
  if (n == "dsupint1") return put( c.dsupint[1], out, cap );
  if (n == "dsupint2") return put( c.dsupint[2], out, cap );
  if (n == "dirbcmasks") return put( c.dirbcmasks, out, cap );
  if (n == "symbcnodes") return put( c.symbcnodes, out, cap );
  if (n == "symbcnorms") return put( c.symbcnorms, out, cap );
  if (n == "farbcnodes") return put( c.farbcnodes, out, cap );
  if (n == "farbcnorms") return put( c.farbcnorms, out, cap );
  if (n == "prebcnodes") return put( c.prebcnodes, out, cap );
  if (n == "prebcvals") return put( c.prebcvals, out, cap );
  if (n == "u") return putf( c.u, out, cap );
  if (n == "un") return putf( c.un, out, cap );
  if (n == "rhs") return putf( c.rhs, out, cap );
  if (n == "grad") return putf( c.grad, out, cap );
  if (n == "p") return putf( c.p, out, cap );
  if (n == "q") return putf( c.q, out, cap );
  if (n == "a") return putf( c.a, out, cap );
  if (n == "bface") {       // flattened: setid, nfaces, face ids ...
    std::vector< std::uint64_t > f;
    for (const auto& [s,ids] : c.bface) { f.push_back( static_cast<std::uint64_t>(s) ); f.push_back( ids.size() ); for (auto i : ids) f.push_back( i ); }
    return put( f, out, cap );
  }
  if (n == "commmap") {     // flattened: neighbour, n, gids (sorted) ...
    std::vector< std::uint64_t > f;
    for (const auto& [b,g] : c.nodeCommMap) {
      f.push_back( static_cast<std::uint64_t>(b) ); f.push_back( g.size() );
      std::vector< std::size_t > s( g.begin(), g.end() ); std::sort( s.begin(), s.end() );
      for (auto i : s) f.push_back( i );
    }
    return put( f, out, cap );
  }
  g_err = "orc_get: unknown array " + n;
  return static_cast< std::size_t >( -1 );
}

int orc_set_u( void* hv, int chare, const double* u )
{
  auto& c = *static_cast< Handle* >( hv )->run->ch.at( static_cast< std::size_t >( chare ) );
  for (std::size_t i=0; i<c.u.nunk(); ++i) for (std::size_t k=0; k<c.u.nprop(); ++k) c.u(i,k) = u[i*c.u.nprop()+k];
  return 0;
}

// Replace one superedge group of a chare (ids + integrals as RieCG::m_dsupedge/m_dsupint hold them):
// lets a test run the reference's kernels on superedges built elsewhere, e.g. by the product's host
// mirror with its element-order triangle walk -- same edges and integrals, other grouping/orientation.
int orc_set_supedge( void* hv, int chare, int k, std::size_t nid, const std::uint64_t* ids,
                     std::size_t nint, const double* ints )
{
  if (k < 0 || k > 2) { g_err = "orc_set_supedge: group must be 0, 1 or 2"; return -1; }
  auto& c = *static_cast< Handle* >( hv )->run->ch.at( static_cast< std::size_t >( chare ) );
  c.dsupedge[static_cast<std::size_t>(k)].assign( ids, ids + nid );
  c.dsupint[static_cast<std::size_t>(k)].assign( ints, ints + nint );
  return 0;
}

int orc_kernel( void* hv, int chare, const char* what, int stage, double t, double dt )
{
  auto h = static_cast< Handle* >( hv );
  auto& c = *h->run->ch.at( static_cast< std::size_t >( chare ) );
  try {
    be::set_cfg( h->run->cfg );
    std::string w( what );
    if (w == "grad") c.grad_own();
    else if (w == "rhs") c.rhs_own( stage, t );
    else if (w == "solve") c.solve( stage, t, dt );
    else if (w == "bc") c.BC( t );
    else if (w == "mindt") { h->run->dt = c.mindt(); }
    else if (w == "zrhs") c.zrhs_own( t, dt );
    else if (w == "krhs") c.krhs_own( t, dt );
    else if (w == "lgrad") c.lgrad_own();
    else if (w == "lrhs") c.lrhs_own( stage, t );
    else if (w == "lsolve") c.lsolve( stage, t, dt );
    else if (w == "aec") c.aec_own();
    else if (w == "alw") c.alw_own( dt );
    else if (w == "lim") c.lim_own();
    else if (w == "zsolve") c.zsolve( t, dt );
#ifdef ORACLE_SHIM
    // the product's reference-signature wrappers (include/xyst_shim.hpp), called with the chare's own
    // containers next to the reference functions above: the compiled drop-in binding under test
    else if (w == "shim_grad") xyst_shim::riemann::grad( c.dsupedge, c.dsupint, c.coord, c.triinpoel, c.u, c.grad );
    else if (w == "shim_rhs") xyst_shim::riemann::rhs( c.dsupedge, c.dsupint, c.coord, c.triinpoel, c.besym, c.grad, c.u, c.v, t, c.tp, c.rhs );
    else if (w == "shim_zrhs") xyst_shim::zalesak::rhs( c.dsupedge, c.dsupint, c.coord, c.triinpoel, c.besym, t, dt, c.tp, c.dtp, c.u, c.rhs );
    else if (w == "shim_release") xyst_shim::release( c.dsupedge );
#endif
    else { g_err = "orc_kernel: unknown " + w; return -1; }
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// ---- linear solver (cg_port.hpp) --------------------------------------------------------
void* orc_cg_create( const char* pc ) {
  auto s = new cg::Solver; s->pc = pc; return s;
}
void orc_cg_destroy( void* h ) { delete static_cast< cg::Solver* >( h ); }
const char* orc_cg_backend() { return cg::matrix_backend(); }

int orc_cg_add( void* h, std::size_t npoin, std::size_t ncomp, std::size_t npsup1, const std::uint64_t* psup1,
                const std::uint64_t* psup2, const std::uint64_t* gid, int ncomm, const int* comm_rank,
                const std::uint64_t* comm_off, const std::uint64_t* comm_gid )
{
  try {
    cg::Psup ps;
    ps.first.assign( psup1, psup1+npsup1 ); ps.second.assign( psup2, psup2+npoin+1 );
    std::vector< std::size_t > g( gid, gid+npoin );
    cg::CommMap cm;
    for (int i=0; i<ncomm; ++i) cm[comm_rank[i]].insert( comm_gid+comm_off[i], comm_gid+comm_off[i+1] );
    return static_cast< int >( static_cast< cg::Solver* >( h )->add( ncomp, ps, g, cm ) );
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int orc_cg_laplacian( void* h, int part, std::size_t ntet, const std::uint64_t* inpoel,
                      const double* x, const double* y, const double* z )
{
  try {
    auto& P = *static_cast< cg::Solver* >( h )->parts.at( static_cast< std::size_t >( part ) );
    std::vector< std::size_t > inp( inpoel, inpoel+ntet*4 );
    auto np = P.gid.size();
    cg::laplacian( P.A, inp, std::vector< double >( x, x+np ), std::vector< double >( y, y+np ), std::vector< double >( z, z+np ) );
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int orc_cg_dirichlet( void* h, int part, std::size_t node, double val, std::size_t pos )
{
  try {
    auto& P = *static_cast< cg::Solver* >( h )->parts.at( static_cast< std::size_t >( part ) );
    P.A.dirichlet( node, val, P.b, P.gid, P.nodeCommMap, pos );
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int orc_cg_set( void* h, int part, const double* x, const double* b )
{
  auto& P = *static_cast< cg::Solver* >( h )->parts.at( static_cast< std::size_t >( part ) );
  if (x) P.x.assign( x, x+P.x.size() );
  if (b) P.b.assign( b, b+P.b.size() );
  return 0;
}

//! name: x b r p q z d (doubles), ia ja (uint64, 1-based), a (doubles, CSR order)
std::size_t orc_cg_get( void* h, int part, const char* name, void* out, std::size_t cap )
{
  auto& P = *static_cast< cg::Solver* >( h )->parts.at( static_cast< std::size_t >( part ) );
  std::string n( name );
  if (n == "x") return put( P.x, out, cap );
  if (n == "b") return put( P.b, out, cap );
  if (n == "r") return put( P.r, out, cap );
  if (n == "q") return put( P.q, out, cap );
  if (n == "d") return put( P.d, out, cap );
  if (n == "ia") return put( P.S.IA(), out, cap );
  if (n == "ja") return put( P.S.JA(), out, cap );
  if (n == "a") {
    const auto& ia = P.S.IA(); const auto& ja = P.S.JA(); auto nc = P.S.Ncomp();
    std::vector< double > a( ja.size() );
    for (std::size_t r=0; r+1<ia.size(); ++r)
      for (std::size_t j=ia[r]-1; j<ia[r+1]-1; ++j) a[j] = P.A( r/nc, (ja[j]-1)/nc, r%nc );
    return put( a, out, cap );
  }
  g_err = "orc_cg_get: unknown " + n; return static_cast< std::size_t >( -1 );
}

int orc_cg_mult( void* h, int part, const double* x, double* r )
{
  auto& P = *static_cast< cg::Solver* >( h )->parts.at( static_cast< std::size_t >( part ) );
  std::vector< double > xv( x, x+P.x.size() ), rv( P.x.size() );
  P.A.mult( xv, rv );
  std::copy( rv.begin(), rv.end(), r );
  return 0;
}

double orc_cg_setup( void* h ) {
  try { return static_cast< cg::Solver* >( h )->setup(); } catch (std::exception& e) { g_err = e.what(); return std::nan(""); }
}
double orc_cg_solve( void* h, std::size_t maxit, double tol, std::uint64_t* it ) {
  try { auto s = static_cast< cg::Solver* >( h ); auto r = s->solve( maxit, tol ); if (it) *it = s->it; return r; }
  catch (std::exception& e) { g_err = e.what(); return std::nan(""); }
}

std::uint64_t orc_siphash_ids( const std::uint64_t* ids, int n )
{
  if (n == 2) return IdHash<2>()( {{ ids[0], ids[1] }} );
  if (n == 3) return IdHash<3>()( {{ ids[0], ids[1], ids[2] }} );
  return 0;
}

} // extern "C"

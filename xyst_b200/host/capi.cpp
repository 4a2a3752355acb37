// xyst_b200/host/capi.cpp -- C entry points of include/xyst_host.h
#include <cstring>
#include <memory>
#include <string>
#include <algorithm>
#include <stdexcept>
#include <limits>
#include <cmath>
#include "riecg.hpp"
#include "refhashset.hpp"
#include "exodus.hpp"
#include "xyst_host.h"

using namespace xyst;

struct xyst_solver {
  Config cfg;
  TetMesh chunk;
  std::unique_ptr< Discretization > disc;
  std::unique_ptr< RieCG > riecg;
  std::unique_ptr< DiagWriter > diag;
};

namespace {
thread_local std::string g_err;
int fail( const std::string& m ) { g_err = m; return 1; }
#define API_BEGIN try {
#define API_END } catch (std::exception& e) { return fail( e.what() ); } return 0;

Config to_cfg( const xyst_host_cfg* c ) {
  if (!c) throw std::runtime_error( "null configuration" );
  // counts against the fixed extents of xyst_host_cfg's arrays, strings against their buffers
  auto cnt = [&]( int n, int cap, const char* what ) {
    if (n < 0 || n > cap) throw std::runtime_error( std::string( "xyst_host_cfg: " ) + what + " out of range" ); };
  cnt( c->nsym, 16, "nsym" ); cnt( c->ndir, 16, "ndir" ); cnt( c->nfar, 16, "nfar" ); cnt( c->npre, 16, "npre" );
  cnt( c->nnoslip, 16, "nnoslip" ); cnt( c->ndirval, 16, "ndirval" ); cnt( c->nfctsys, 8, "nfctsys" );
  cnt( c->np_dir, 16, "np_dir" ); cnt( c->np_dirval, 16, "np_dirval" ); cnt( c->np_sym, 16, "np_sym" );
  if (c->ncomp < 1 || c->ncomp + 1 > 12) throw std::runtime_error( "xyst_host_cfg: ncomp out of range (1..11)" );
  auto str = [&]( const char* s, std::size_t cap, const char* what ) {
    if (strnlen( s, cap ) == cap) throw std::runtime_error( std::string( "xyst_host_cfg: " ) + what + " is not NUL-terminated" ); };
  str( c->problem, sizeof c->problem, "problem" ); str( c->flux, sizeof c->flux, "flux" ); str( c->solver, sizeof c->solver, "solver" );
  str( c->p_pc, sizeof c->p_pc, "p_pc" ); str( c->mom_pc, sizeof c->mom_pc, "mom_pc" );
  Config k;
  k.problem = c->problem; k.flux = c->flux; k.ncomp = static_cast< std::size_t >( c->ncomp );
  k.alpha = c->alpha; k.kappa = c->kappa; k.r0 = c->r0; k.ce = c->ce; k.beta = {{ c->beta[0], c->beta[1], c->beta[2] }};
  k.gamma = c->gamma; k.p0 = c->p0; k.cfl = c->cfl; k.dt = c->dt; k.t0 = c->t0; k.term = c->term;
  k.nstep = c->nstep; k.diag_iter = c->diag_iter ? c->diag_iter : 1;
  k.stab2 = c->stab2 != 0; k.stab2coef = c->stab2coef; k.exact_muscl = c->exact_muscl != 0; k.reforder = c->reforder;
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
    k.bc_pre.push_back( c->pre_sets[i] );
    k.pre_density.push_back( c->pre_density[i] ); k.pre_pressure.push_back( c->pre_pressure[i] );
  }
  if (c->solver[0]) k.solver = c->solver;
  if (k.solver == "zalcg" || k.solver == "kozcg") { k.fct = c->fct != 0; k.fctclip = c->fctclip != 0; k.fctdif = c->fctdif;
    for (int i=0; i<c->nfctsys; ++i) k.fctsys.push_back( c->fctsys[i] ); }
  k.steady = c->steady != 0; k.residual = c->residual; k.rescomp = c->rescomp ? c->rescomp : 1;
  if (c->rgas != 0.0) k.rgas = c->rgas;
  k.turkel = c->turkel; k.velinf = {{ c->velinf[0], c->velinf[1], c->velinf[2] }};
  k.ic_density = c->ic_density; k.ic_pressure = c->ic_pressure;
  k.ic_velocity = {{ c->ic_velocity[0], c->ic_velocity[1], c->ic_velocity[2] }};
  if (c->soundspeed != 0.0) k.soundspeed = c->soundspeed;
  k.src_location = {{ c->src_location[0], c->src_location[1], c->src_location[2] }};
  k.src_radius = c->src_radius; k.src_release_time = c->src_release_time;
  if (c->freezeflow != 0.0) k.freezeflow = c->freezeflow;
  k.freezetime = c->freezetime;
  if (k.freezeflow > 1.0 && k.solver != "chocg" && k.solver != "kozcg" && k.solver != "zalcg")
    throw std::runtime_error( "freezeflow is implemented for ZalCG, KozCG and ChoCG only" );
  if (k.solver == "chocg" || k.solver == "lohcg") {
    k.mu = c->mu; k.dif = c->dif; k.stab = c->stab != 0; k.rk = c->rk ? c->rk : 1;
    for (int i=0; i<c->nnoslip; ++i) k.bc_noslip.push_back( c->noslip[i] );
    for (int i=0; i<c->ndirval; ++i) k.bc_dirval.emplace_back( c->dirval[i], c->dirval[i] + k.ncomp+1 );
    k.p_iter = c->p_iter; k.p_tol = c->p_tol; if (c->p_pc[0]) k.p_pc = c->p_pc;
    for (int i=0; i<c->np_dir; ++i) k.p_bc_dir.push_back( { c->p_dir[i][0], c->p_dir[i][1] } );
    for (int i=0; i<c->np_dirval; ++i) k.p_bc_dirval.push_back( { c->p_dirval[i][0], c->p_dirval[i][1] } );
    for (int i=0; i<c->np_sym; ++i) k.p_bc_sym.push_back( c->p_sym[i] );
    if (c->p_hydrostat_set) k.p_hydrostat = c->p_hydrostat;
    k.theta = c->theta; k.mom_iter = c->mom_iter ? c->mom_iter : 10; k.mom_tol = c->mom_tol; if (c->mom_pc[0]) k.mom_pc = c->mom_pc;
  }
  return k;
}

template< class T >
std::size_t put( const std::vector< T >& v, void* out, std::size_t cap ) {
  auto bytes = v.size()*sizeof(T);
  if (out && cap >= bytes && bytes) std::memcpy( out, v.data(), bytes );
  return bytes;
}

void finish( xyst_solver* s ) {
  s->disc.reset( new Discretization( s->chunk, s->cfg ) );
  s->riecg.reset( new RieCG( *s->disc, s->chunk, s->cfg ) );
}
}

extern "C" {

const char* xyst_host_last_error(void) { return g_err.c_str(); }

int xyst_solver_create_box( const xyst_host_cfg* cfg, size_t nx, size_t ny, size_t nz,
                            double Lx, double Ly, double Lz, int nparts, int part, xyst_solver** out )
{
  API_BEGIN
  auto s = std::make_unique< xyst_solver >();
  s->cfg = to_cfg( cfg );
  auto r = boxPartRange( nx, ny, nz, nparts, part );
  s->chunk = boxMesh( nx, ny, nz, Lx, Ly, Lz, r[0], r[1], r[2], r[3], r[4], r[5] );
  finish( s.get() );
  // shared nodes with every other partition: intersection of the node boxes, ascending gid
  const std::size_t px = nx+1, py = ny+1;
  for (int q=0; q<nparts; ++q) {
    if (q == part) continue;
    auto o = boxPartRange( nx, ny, nz, nparts, q );
    std::size_t lo[3], hi[3]; bool empty = false;
    for (int d=0; d<3; ++d) { lo[d] = std::max( r[2*d], o[2*d] ); hi[d] = std::min( r[2*d+1], o[2*d+1] ); if (lo[d] > hi[d]) empty = true; }
    if (empty) continue;
    auto& g = s->disc->NodeCommMap()[q];
    for (std::size_t k=lo[2]; k<=hi[2]; ++k) for (std::size_t j=lo[1]; j<=hi[1]; ++j) for (std::size_t i=lo[0]; i<=hi[0]; ++i)
      g.push_back( (k*py + j)*px + i );
  }
  *out = s.release();
  API_END
}

int xyst_solver_create_mesh( const xyst_host_cfg* cfg, size_t npoin, const double* x, const double* y,
                             const double* z, size_t ntet, const uint64_t* tets, int nsets,
                             const int* set_id, const uint64_t* set_off, const uint64_t* set_tri,
                             int nparts, int part, const int32_t* tetpart, xyst_solver** out )
{
  API_BEGIN
  auto s = std::make_unique< xyst_solver >();
  s->cfg = to_cfg( cfg );
  Coords co;
  co[0].assign( x, x+npoin ); co[1].assign( y, y+npoin ); co[2].assign( z, z+npoin );
  std::vector< std::size_t > g( tets, tets+ntet*4 );
  std::vector< int > tp;
  if (nparts > 1) { if (tetpart) tp.assign( tetpart, tetpart+ntet ); else tp = rcb( co, g, nparts ); }
  else tp.assign( ntet, 0 );
  // node -> partitions touching it (bit mask, nparts <= 64)
  if (nparts > 64) throw std::runtime_error( "at most 64 partitions" );
  std::vector< std::uint64_t > nodeparts( npoin, 0 );
  for (std::size_t e=0; e<ntet; ++e) for (int k=0; k<4; ++k) nodeparts[ g[e*4+static_cast<std::size_t>(k)] ] |= 1ULL << tp[e];
  auto& ch = s->chunk;
  ch.coord = co;                       // indexed by global id (gid left empty)
  for (std::size_t e=0; e<ntet; ++e) if (tp[e] == part) for (int k=0; k<4; ++k) ch.ginpoel.push_back( g[e*4+static_cast<std::size_t>(k)] );
  const std::uint64_t me = 1ULL << part;
  for (int i=0; i<nsets; ++i) {
    auto& t = ch.sidetri[ set_id[i] ];
    for (auto f=set_off[i]; f<set_off[i+1]; ++f) {
      bool mine = true;
      for (int k=0; k<3; ++k) if (!(nodeparts[ set_tri[f*3+static_cast<std::uint64_t>(k)] ] & me)) mine = false;
      if (mine) for (int k=0; k<3; ++k) t.push_back( set_tri[f*3+static_cast<std::uint64_t>(k)] );
    }
  }
  finish( s.get() );
  for (std::size_t p=0; p<npoin; ++p) {
    auto m = nodeparts[p];
    if (!(m & me) || m == me) continue;
    for (int q=0; q<nparts; ++q) if (q != part && (m & (1ULL << q))) s->disc->NodeCommMap()[q].push_back( p );
  }
  *out = s.release();
  API_END
}

int xyst_exo_read( const char* path, size_t* npoin, size_t* ntet, int* nsets, size_t* ntri,
                   double* x, double* y, double* z, uint64_t* tets, int32_t* set_id, uint64_t* set_off,
                   uint64_t* set_tri )
{
  API_BEGIN
  auto m = readExodus( path );
  std::size_t nt = 0;
  for (const auto& [id,t] : m.sidetri) nt += t.size()/3;
  if (npoin) *npoin = m.coord[0].size();
  if (ntet) *ntet = m.tets.size()/4;
  if (nsets) *nsets = static_cast< int >( m.sidetri.size() );
  if (ntri) *ntri = nt;
  if (x) std::copy( m.coord[0].begin(), m.coord[0].end(), x );
  if (y) std::copy( m.coord[1].begin(), m.coord[1].end(), y );
  if (z) std::copy( m.coord[2].begin(), m.coord[2].end(), z );
  if (tets) std::copy( m.tets.begin(), m.tets.end(), tets );
  if (set_id && set_off && set_tri) {
    std::size_t i = 0, k = 0; set_off[0] = 0;
    for (const auto& [id,t] : m.sidetri) {
      set_id[i] = id;
      for (auto n : t) set_tri[k++] = n;
      set_off[i+1] = set_off[i] + t.size()/3;
      ++i;
    }
  }
  API_END
}

int xyst_solver_create_exo( const xyst_host_cfg* cfg, const char* path, int nparts, int part, xyst_solver** out )
{
  std::vector< std::uint64_t > tets, off{ 0 }, tri;
  std::vector< int > ids;
  ExoMesh m;
  try {
    m = readExodus( path );
    tets.assign( m.tets.begin(), m.tets.end() );
    for (const auto& [id,t] : m.sidetri) { ids.push_back( id ); tri.insert( tri.end(), t.begin(), t.end() ); off.push_back( tri.size()/3 ); }
  } catch (std::exception& e) { return fail( e.what() ); }
  return xyst_solver_create_mesh( cfg, m.coord[0].size(), m.coord[0].data(), m.coord[1].data(), m.coord[2].data(),
                                  tets.size()/4, tets.data(), static_cast< int >( ids.size() ), ids.data(), off.data(),
                                  tri.data(), nparts, part, nullptr, out );
}

int xyst_solver_diag_file( xyst_solver* s, const char* path, int precision )
{
  API_BEGIN
  s->diag.reset( new DiagWriter( path, precision > 0 ? precision : 8, diagNames( s->cfg ) ) );
  API_END
}

int xyst_solver_destroy( xyst_solver* s ) { delete s; return 0; }

int xyst_solver_prepare( xyst_solver* s ) { API_BEGIN s->riecg->prepare(); API_END }

int xyst_solver_attach( xyst_solver* s, int device, int nranks, int rank, const void* id )
{ API_BEGIN s->riecg->attach( device, nranks, rank, id ); API_END }

int xyst_solver_set_comm( xyst_solver* s, xyst_comm_fn fn, void* user, int nranks, int rank )
{
  API_BEGIN
  s->riecg->setRanks( nranks, rank );
  s->riecg->setComm(
    [fn,user]( int w, std::vector< real >& v ){ fn( user, 0, w, v.size()/static_cast<std::size_t>(w), v.data() ); },
    [fn,user]( int op, std::vector< real >& v ){ fn( user, op == 0 ? 1 : 2, 1, v.size(), v.data() ); } );
  API_END
}

int xyst_solver_host_setup( xyst_solver* s ) { API_BEGIN s->riecg->hostSetup(); API_END }

int xyst_solver_set_u0( xyst_solver* s, const double* u0 )
{ API_BEGIN s->riecg->m_u0.assign( u0, u0 + s->disc->Gid().size()*s->cfg.ncomp ); API_END }

int xyst_solver_setup( xyst_solver* s ) { API_BEGIN s->riecg->setup(); API_END }

int xyst_solver_set_u( xyst_solver* s, const double* u )
{
  API_BEGIN
  s->riecg->setSolution( std::vector< real >( u, u + s->disc->Gid().size()*s->cfg.ncomp ) );
  API_END
}

int xyst_solver_step( xyst_solver* s, int nsteps, double* rows, size_t cap, size_t* nrows, size_t* ncols )
{
  API_BEGIN
  size_t nr = 0, nc = 0, used = 0;
  std::vector< real > row;
  for (int i=0; i<nsteps; ++i) {
    if (s->riecg->m_finished && !s->riecg->pendingDiag()) break;
    s->riecg->step( rows || s->diag ? &row : nullptr );
    if (s->diag && !row.empty()) s->diag->write( row );
    if (rows && !row.empty()) {
      nc = row.size();
      if (used + nc <= cap) { std::memcpy( rows+used, row.data(), nc*sizeof(double) ); used += nc; ++nr; }
    }
  }
  if (nrows) *nrows = nr;
  if (ncols) *ncols = nc;
  API_END
}

int xyst_solver_step_unfused( xyst_solver* s, int nsteps )
{
  API_BEGIN
  auto& r = *s->riecg;
  for (int i=0; i<nsteps; ++i) {
    if (r.m_finished) break;
    r.advance( r.dt() );
    for (int st=0; st<3; ++st) { r.grad(); r.rhs(); r.solve(); }
    r.Disc().next();
    if (r.Disc().finished()) r.m_finished = true;
  }
  API_END
}

double xyst_solver_scalar( xyst_solver* s, const char* name )
{
  std::string n( name );
  auto& d = *s->disc;
  if (n == "npoin") return static_cast< double >( d.Gid().size() );
  if (n == "ntet") return static_cast< double >( d.Inpoel().size()/4 );
  if (n == "nedge") return static_cast< double >( s->riecg->m_dsupedge[0].size()/4*6 + s->riecg->m_dsupedge[1].size() + s->riecg->m_dsupedge[2].size()/2 );
  if (n == "ntri") return static_cast< double >( s->riecg->m_triinpoel.size()/3 );
  if (n == "t") return d.T();
  if (n == "dt") return d.Dt();
  if (n == "it") return static_cast< double >( d.It() );
  if (n == "meshvol") return d.MeshVol();
  if (n == "finished") return s->riecg->m_finished ? 1.0 : 0.0;
  if (n == "pit") return static_cast< double >( s->riecg->m_pit );
  if (n == "mit") return static_cast< double >( s->riecg->m_mit );
  if (n == "nshared") return static_cast< double >( d.sharedNodes().size() );
  if (n.rfind( "timing", 0 ) == 0) { auto i = static_cast< std::size_t >( std::stoi( n.substr(6) ) ); return i < s->riecg->timings.size() ? s->riecg->timings[i] : -1.0; }
  return std::nan( "" );
}

size_t xyst_solver_get( xyst_solver* s, const char* name, void* out, size_t cap )
{
  try {
    std::string n( name );
    auto& d = *s->disc; auto& r = *s->riecg;
    if (n == "gid") return put( d.Gid(), out, cap );
    if (n == "inpoel") return put( d.Inpoel(), out, cap );
    if (n == "x") return put( d.Coord()[0], out, cap );
    if (n == "y") return put( d.Coord()[1], out, cap );
    if (n == "z") return put( d.Coord()[2], out, cap );
    if (n == "vol") return put( d.Vol(), out, cap );
    if (n == "v") return put( d.V(), out, cap );
    if (n == "triinpoel") return put( r.m_triinpoel, out, cap );
    if (n == "besym") return put( r.m_besym, out, cap );
    if (n == "dsupedge0") return put( r.m_dsupedge[0], out, cap );
    if (n == "dsupedge1") return put( r.m_dsupedge[1], out, cap );
    if (n == "dsupedge2") return put( r.m_dsupedge[2], out, cap );
    if (n == "dsupint0") return put( r.m_dsupint[0], out, cap );
    if (n == "dsupint1") return put( r.m_dsupint[1], out, cap );
    if (n == "dsupint2") return put( r.m_dsupint[2], out, cap );
    if (n == "dirbcmasks") return put( r.m_dirbcmasks, out, cap );
    if (n == "symbcnodes") return put( r.m_symbcnodes, out, cap );
    if (n == "symbcnorms") return put( r.m_symbcnorms, out, cap );
    if (n == "dirbcval") return put( r.m_dirbcval, out, cap );
    if (n == "dirbcmaskp") return put( r.m_dirbcmaskp, out, cap );
    if (n == "dirbcvalp") return put( r.m_dirbcvalp, out, cap );
    if (n == "noslipbcnodes") return put( r.m_noslipbcnodes, out, cap );
    if (n == "plhs_ia") return put( r.m_plhs_ia, out, cap );
    if (n == "plhs_ja") return put( r.m_plhs_ja, out, cap );
    if (n == "plhs_a") return put( r.m_plhs_a, out, cap );
    if (n == "mlhs_a") return put( r.m_mlhs_a, out, cap );
    if (n == "pr") return put( r.choGet( "pr", 1 ), out, cap );
    if (n == "dp") return put( r.choGet( "dp", 1 ), out, cap );
    if (n == "pgrad") return put( r.choGet( "pgrad", 3 ), out, cap );
    if (n == "u0") return put( r.m_u0, out, cap );
    if (n == "u") return put( r.solution(), out, cap );
    if (n == "pbc") return put( r.pressureBC(), out, cap );
    if (n == "mbcrows") return put( r.momentumBCRows(), out, cap );
    if (n == "shared") return put( d.sharedNodes(), out, cap );
    if (n == "bface") {
      std::vector< std::uint64_t > f;
      for (const auto& [st,ids] : r.m_bface) { f.push_back( static_cast<std::uint64_t>(st) ); f.push_back( ids.size() ); for (auto i : ids) f.push_back( i ); }
      return put( f, out, cap );
    }
    if (n == "commmap") {
      std::vector< std::uint64_t > f;
      for (const auto& [b,g] : d.NodeCommMap()) { f.push_back( static_cast<std::uint64_t>(b) ); f.push_back( g.size() ); for (auto i : g) f.push_back( i ); }
      return put( f, out, cap );
    }
    g_err = "xyst_solver_get: unknown array " + n;
  } catch (std::exception& e) { g_err = e.what(); }
  return static_cast< size_t >( -1 );
}

xyst_ctx* xyst_solver_ctx( xyst_solver* s ) { return s->riecg->ctx(); }

int xyst_box_counts( size_t nx, size_t ny, size_t nz, size_t* npoin, size_t* ntet, size_t* ntri )
{
  *npoin = (nx+1)*(ny+1)*(nz+1); *ntet = 6*nx*ny*nz; *ntri = 4*(nx*ny + ny*nz + nx*nz);
  return 0;
}

int xyst_box_mesh( size_t nx, size_t ny, size_t nz, double Lx, double Ly, double Lz,
                   double* x, double* y, double* z, uint64_t* tets,
                   int32_t set_id[6], uint64_t set_off[7], uint64_t* set_tri )
{
  API_BEGIN
  auto m = boxMesh( nx, ny, nz, Lx, Ly, Lz, 0, nx, 0, ny, 0, nz );
  std::copy( m.coord[0].begin(), m.coord[0].end(), x );
  std::copy( m.coord[1].begin(), m.coord[1].end(), y );
  std::copy( m.coord[2].begin(), m.coord[2].end(), z );
  std::copy( m.ginpoel.begin(), m.ginpoel.end(), tets );
  int i = 0; set_off[0] = 0; size_t k = 0;
  for (const auto& [s,t] : m.sidetri) {
    set_id[i] = s;
    for (auto n : t) set_tri[k++] = n;
    set_off[i+1] = set_off[i] + t.size()/3;
    ++i;
  }
  API_END
}

int xyst_chare_count( double virtualization, uint64_t load, int npe, uint64_t* chunksize, uint64_t* remainder, uint64_t* nchare )
{
  API_BEGIN
  auto eps = std::numeric_limits< double >::epsilon();
  if (!(virtualization > -eps && virtualization < 1.0+eps)) throw std::runtime_error( "Virtualization parameter must be between [0.0...1.0]" );
  if (npe <= 0) throw std::runtime_error( "Number of processing elements must be larger than zero" );
  if (!nchare) throw std::runtime_error( "null argument" );
  const auto n = static_cast< double >( load ) / npe;                          // LoadDistributor.cpp:71
  auto chunk = static_cast< uint64_t >( (1.0 - n) * virtualization + n );       // :74
  if (chunk == 0 || load < chunk) throw std::runtime_error( "Load must be larger than chunksize" );
  uint64_t nc = load / chunk;                                                    // :79
  uint64_t rem = load - nc * chunk;                                              // :82
  chunk += rem / nc;                                                             // :85
  rem = load - nc * chunk;                                                       // :88
  if (chunksize) *chunksize = chunk;
  if (remainder) *remainder = rem;
  *nchare = nc;
  API_END
}

int xyst_rcb( size_t npoin, const double* x, const double* y, const double* z, size_t ntet,
              const uint64_t* tets, int nparts, int32_t* part )
{
  API_BEGIN
  Coords co;
  co[0].assign( x, x+npoin ); co[1].assign( y, y+npoin ); co[2].assign( z, z+npoin );
  std::vector< std::size_t > g( tets, tets+ntet*4 );
  auto p = rcb( co, g, nparts );
  std::copy( p.begin(), p.end(), part );
  API_END
}

int xyst_rib( size_t npoin, const double* x, const double* y, const double* z, size_t ntet,
              const uint64_t* tets, int nparts, int32_t* part )
{
  API_BEGIN
  Coords co;
  co[0].assign( x, x+npoin ); co[1].assign( y, y+npoin ); co[2].assign( z, z+npoin );
  std::vector< std::size_t > g( tets, tets+ntet*4 );
  auto p = rib( co, g, nparts );
  std::copy( p.begin(), p.end(), part );
  API_END
}

// test hook: iteration order of the real std::unordered_set vs RefOrderFaceSet after
// inserting nface faces and erasing nerase of them; writes surviving faces (3 ids each)
int xyst_test_faceset_order( size_t nface, const uint64_t* faces, size_t nerase, const uint64_t* erase,
                             uint64_t* out_std, uint64_t* out_emu, size_t* nout )
{
  API_BEGIN
  using Face = std::array< std::size_t, 3 >;
  std::unordered_set< Face, IdHash<3>, IdEq<3> > a;
  RefOrderFaceSet b;
  for (size_t i=0; i<nface; ++i) { Face f{{ faces[i*3], faces[i*3+1], faces[i*3+2] }}; a.insert( f ); b.insert( f ); }
  for (size_t i=0; i<nerase; ++i) { Face f{{ erase[i*3], erase[i*3+1], erase[i*3+2] }}; a.erase( f ); b.erase( f ); }
  size_t k = 0;
  for (const auto& f : a) { for (auto n : f) out_std[k++] = n; }
  size_t m = 0;
  b.forEach( [&]( const Face& f ){ for (auto n : f) out_emu[m++] = n; } );
  if (k != m) throw std::runtime_error( "faceset size mismatch" );
  *nout = k/3;
  API_END
}

int xyst_box_part_range( size_t nx, size_t ny, size_t nz, int nparts, int part, uint64_t range[6] )
{
  API_BEGIN
  auto r = boxPartRange( nx, ny, nz, nparts, part );
  for (int i=0; i<6; ++i) range[i] = r[static_cast<std::size_t>(i)];
  API_END
}

} // extern "C"

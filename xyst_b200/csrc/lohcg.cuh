// LohCG device path: the artificial-compressibility solver for constant-density flow of the
// reference (src/Inciter/LohCG.cpp, src/Physics/Lohner.cpp). Unknowns (p,u,v,w), stored as four
// rows [4][NP] of three rotating buffers; superedge integrals of stride 4 (normal, grad_p.grad_q/(6J);
// LohCG::domint :407-453).
//
// lohner::div/grad/vgrad/flux are the Chorin operators on the velocity rows (Lohner.cpp:35-589 vs
// Chorin.cpp:85-638: same loops, four integrals per edge instead of five, velocity at components
// 1..3): the context's ChoCG velocity pointers (cU, cUn, cUx) point at ROW 1 of the LohCG state, so
// xyst_chocg_div/vgrad/flux/grad/project/pinit and the CG entries serve LohCG unchanged; the pressure
// is the row in front of them. What is LohCG's own is below: the gradient of all unknowns (damp4),
// lohner::rhs with the pressure equation (Lohner.cpp:724-1130) fused with the RK update
// (LohCG::solve :1587-1631), the pressure Dirichlet BC (physics::dirbcp), dt and diagnostics.
// Included twice by xyst_b200.cu like chocg.cuh (XYST_LOHCG_KERNELS / XYST_LOHCG_API).

#ifdef XYST_LOHCG_KERNELS

struct LohP { int stab, stab2; double stab2coef, mu, s; };

// advection edge flux of (p,u,v,w): second-order damping (Lohner.cpp:724-795) or fourth-order damping
// with the limited reconstruction of all four unknowns (:797-914)
template< bool DAMP4 >
__device__ __forceinline__ void loh_adv( const double d[3], double lap, const double ua[4], const double ub[4],
    const double ga[12], const double gb[12], const double dx[3], const LohP& C, double f[4] )
{
  double uL[4] = { ua[0], ua[1], ua[2], ua[3] }, uR[4] = { ub[0], ub[1], ub[2], ub[3] };
  if (DAMP4) {
    #pragma unroll
    for (int c=0; c<4; ++c) {
      double g1 = ga[c*3+0]*dx[0] + ga[c*3+1]*dx[1] + ga[c*3+2]*dx[2];
      double g2 = gb[c*3+0]*dx[0] + gb[c*3+1]*dx[1] + gb[c*3+2]*dx[2];
      double delta2 = uR[c] - uL[c];
      double delta1 = 2.0 * g1 - delta2;
      double delta3 = 2.0 * g2 - delta2;
      // van Leer limited increments in the one-reciprocal-per-side form of the RieCG kernels
      // (riecg_kernels.cuh, vanleer<false>: algebraically the reference's eight quotients, rounding
      // differs at 1e-16)
      double incL, incR;
      vanleer< false >( delta1, delta2, delta3, incL, incR );
      uL[c] += incL;
      uR[c] -= incR;
    }
  }
  double vnL = uL[1]*d[0] + uL[2]*d[1] + uL[3]*d[2];
  double vnR = uR[1]*d[0] + uR[2]*d[1] + uR[3]*d[2];
  double s2 = C.s*C.s;
  double v = lap * C.mu;
  double aw = 0.0;
  if (C.stab) aw = fabs( vnL + vnR ) / 2.0;
  if (C.stab2) {
    double len = sqrt( d[0]*d[0] + d[1]*d[1] + d[2]*d[2] );
    double sl = fabs(vnL) + C.s*len, sr = fabs(vnR) + C.s*len;
    aw += C.stab2coef * fmax( sl, sr );
  }
  double pf = uL[0] + uR[0];
  f[0] = (vnL + vnR + aw*(uR[0]-uL[0]))*s2;
  if (DAMP4) {
    #pragma unroll
    for (int c=1; c<4; ++c) f[c] = uL[c]*vnL + uR[c]*vnR + pf*d[c-1] + aw*(uR[c]-uL[c]) - v*(ub[c]-ua[c]);
  } else {
    #pragma unroll
    for (int c=1; c<4; ++c) f[c] = uL[c]*vnL + uR[c]*vnR + pf*d[c-1] + (aw-v)*(uR[c]-uL[c]);
  }
}

// lohner::rhs gathered per node; with Uout the update u = un - rk dt rhs / vol is applied in the same pass.
// U, Un, Uout, R: [4][NP] (p,u,v,w); G: [12][NP] gradients of the four unknowns (damp4)
template< bool DAMP4 >
__global__ void __launch_bounds__(NODE_THREADS)
k_loh_rhs( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int* __restrict__ inc_e,
           const int* __restrict__ inc_q, const double* __restrict__ D, size_t nslot,
           const double* __restrict__ U, const double* __restrict__ G, const double* __restrict__ X, LohP C,
           const int* __restrict__ bslot, const int* __restrict__ bn_off, const int* __restrict__ bn_face,
           const int* __restrict__ tri, const double* __restrict__ fn, const double* __restrict__ S,
           const double* __restrict__ v, const double* __restrict__ vol,
           const double* __restrict__ Un, double sdt, double* __restrict__ Uout, double* __restrict__ R )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[4] = { 0.0, 0.0, 0.0, 0.0 }, um[4], gm[12], xm[3];
  #pragma unroll
  for (int i=0; i<4; ++i) um[i] = U[i*NP+p];
  #pragma unroll
  for (int i=0; i<3; ++i) xm[i] = DAMP4 ? X[i*NP+p] : 0.0;
  #pragma unroll
  for (int i=0; i<12; ++i) gm[i] = DAMP4 ? G[i*NP+p] : 0.0;
  for (int k=0; k<kmax; ++k) {
    long long ii = base + (long long)k*32 + lane;
    int se = __ldg( inc_e + ii );
    if (se == 0) continue;
    int nb = __ldg( inc_q + ii );
    size_t sl = (size_t)(abs(se)-1);
    double d[3] = { __ldg( D + sl ), __ldg( D + nslot + sl ), __ldg( D + 2*nslot + sl ) };
    double lap = __ldg( D + 3*nslot + sl );
    double uo[4], go[12], xo[3];
    #pragma unroll
    for (int i=0; i<4; ++i) uo[i] = __ldg( U + i*NP + nb );
    #pragma unroll
    for (int i=0; i<3; ++i) xo[i] = DAMP4 ? __ldg( X + i*NP + nb ) : 0.0;
    #pragma unroll
    for (int i=0; i<12; ++i) go[i] = DAMP4 ? __ldg( G + i*NP + nb ) : 0.0;
    double f[4];
    const bool first = se < 0;             // this node is the edge's first node
    if (DAMP4) {
      // one evaluation for both orientations: the end states are ordered by selects, so that lanes
      // with different edge orientations do not diverge around the (expensive) limited reconstruction
      double ua[4], ub[4], ga[12], gb[12], dx[3];
      #pragma unroll
      for (int i=0; i<4; ++i) { ua[i] = first ? um[i] : uo[i]; ub[i] = first ? uo[i] : um[i]; }
      #pragma unroll
      for (int i=0; i<12; ++i) { ga[i] = first ? gm[i] : go[i]; gb[i] = first ? go[i] : gm[i]; }
      #pragma unroll
      for (int i=0; i<3; ++i) dx[i] = first ? xo[i]-xm[i] : xm[i]-xo[i];
      loh_adv< DAMP4 >( d, lap, ua, ub, ga, gb, dx, C, f );
      #pragma unroll
      for (int c=0; c<4; ++c) acc[c] = first ? acc[c] - f[c] : acc[c] + f[c];
    } else if (first) {
      double dx[3] = { 0.0, 0.0, 0.0 };
      loh_adv< DAMP4 >( d, lap, um, uo, gm, go, dx, C, f );
      #pragma unroll
      for (int c=0; c<4; ++c) acc[c] -= f[c];
    } else {
      double dx[3] = { 0.0, 0.0, 0.0 };
      loh_adv< DAMP4 >( d, lap, uo, um, go, gm, dx, C, f );
      #pragma unroll
      for (int c=0; c<4; ++c) acc[c] += f[c];
    }
  }
  int bs = bslot[p];
  if (bs >= 0) {
    double s2 = C.s*C.s;
    for (int i=bn_off[bs]; i<bn_off[bs+1]; ++i) {
      int f = bn_face[i] >> 2, kk = bn_face[i] & 3, N[3]; double n[3];
      cho_face( tri, fn, f, N, n );
      double fl[4][3];
      #pragma unroll
      for (int m=0; m<3; ++m) {
        double pr = U[N[m]], u = U[NP+N[m]], vv = U[2*NP+N[m]], w = U[3*NP+N[m]];
        double vn = n[0]*u + n[1]*vv + n[2]*w;
        fl[0][m] = vn * s2;
        fl[1][m] = u*vn + pr*n[0];
        fl[2][m] = vv*vn + pr*n[1];
        fl[3][m] = w*vn + pr*n[2];
      }
      #pragma unroll
      for (int c=0; c<4; ++c) acc[c] += cho_w8( fl[c][0], fl[c][1], fl[c][2], kk );
    }
  }
  if (S) {                                      // lohner::src, Lohner.cpp:1073-1087
    double vo = v[p];
    #pragma unroll
    for (int c=0; c<4; ++c) acc[c] -= S[c*NP+p] * vo;
  }
  if (R) {
    #pragma unroll
    for (int c=0; c<4; ++c) R[c*NP+p] = acc[c];
  }
  if (Uout) {
    double vp = vol[p];
    #pragma unroll
    for (int c=0; c<4; ++c) Uout[c*NP+p] = Un[c*NP+p] - sdt*acc[c]/vp;
  }
}

// physics::dirbc (BC.cpp:29-72) on all four unknowns; U: [4][NP]
__global__ void k_loh_dirbc( int nd, size_t NP, int m, const int* __restrict__ node, const int* __restrict__ mask,
                             const double* __restrict__ val, double* __restrict__ U )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nd) return;
  size_t p = node[i];
  for (int c=0; c<m; ++c) if (mask[i*m+c]) U[c*NP+p] = val[i*m+c];
}
// physics::dirbcp (BC.cpp:74-108): pressure Dirichlet values
__global__ void k_loh_pdir( int n, const int* __restrict__ node, const double* __restrict__ val, double* __restrict__ P )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) P[ node[i] ] = val[i];
}

// LohCG::dt :1401-1449; V: velocity rows [3][NP]
__global__ void __launch_bounds__(RED_THREADS)
k_loh_dt( size_t npoin, size_t NP, const double* __restrict__ V, const double* __restrict__ vol, double c, double dif,
          double* __restrict__ part )
{
  double m[1] = { 1.7976931348623157e308 };
  for (size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p < npoin; p += (size_t)gridDim.x*blockDim.x) {
    double u = V[p], v = V[NP+p], w = V[2*NP+p];
    double vel = sqrt( u*u + v*v + w*w );
    double L = cbrt( vol[p] );
    m[0] = fmin( m[0], L / fmax( vel+c, 1.0e-8 ) );
    if (dif > 2.220446049250313e-16) m[0] = fmin( m[0], L * L / dif );
  }
  block_reduce< 1, true >( m, part );
}

// NodeDiagnostics::accompute sums (NodeDiagnostics.cpp:270-372): [0..3] v u_c^2, [4..7] v (u-un)_c^2,
// with an analytic solution [8..11] L2 and [12..15] L1 error sums (component 0 unused)
__global__ void __launch_bounds__(RED_THREADS)
k_loh_diag( size_t npoin, size_t NP, const double* __restrict__ U, const double* __restrict__ Un,
            const double* __restrict__ v, const double* __restrict__ an, int anstride, double* __restrict__ part )
{
  double a[16];
  #pragma unroll
  for (int i=0; i<16; ++i) a[i] = 0.0;
  for (size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p < npoin; p += (size_t)gridDim.x*blockDim.x) {
    double vp = v[p];
    #pragma unroll
    for (int c=0; c<4; ++c) {
      double u = U[c*NP+p], du = u - Un[c*NP+p];
      a[c] += u*u*vp;
      a[4+c] += du*du*vp;
      if (an && c > 0) { double e = u - an[p*anstride+c]; a[8+c] += e*e*vp; a[12+c] += fabs(e)*vp; }
    }
  }
  block_reduce< 16, false >( a, part );
}

#endif // XYST_LOHCG_KERNELS

#ifdef XYST_LOHCG_API

namespace {
void loh_need( xyst_ctx* c ) {
  need_mesh( c );
  if (!c->loh) throw std::runtime_error( "LohCG needs stride-4 superedge integrals with the Laplacian term: use xyst_lohcg_mesh_upload" );
}
LohP lohp( const xyst_ctx* c ) { return LohP{ c->chp.stab, c->chp.stab2, c->chp.stab2coef, c->chp.mu, c->loh_s }; }
// [4][NP] state behind the velocity pointer of the ChoCG machinery
double* loh_state( double* vel, size_t NP ) { return vel - NP; }

int loh_rows( const xyst_ctx* c ) { return 4 + c->cns; }
void loh_rhs( xyst_ctx* c, const double* Un, double sdt, double* Uout, double* R ) {
  ProfScope ps( c, "loh_rhs" );
  auto g = cho_grid( c );
  size_t NP = c->NP;
  const double* U = loh_state( c->cU, NP );
  if ((cho_parts( c ) || c->cns) && !R) R = c->cR.p;            // the shared nodes' parts travel before they are used
  const double* S = c->cS.p;                                     // [4+ns][NP] or null
  if (c->chp.flux == 1) {
    { ProfScope pg( c, "loh_grad" );
      k_cho_grad< 4 ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, NP, c->sl_base.p, c->inc_e.p, c->inc_q.p,
        c->D.p, c->nslot, U, c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, c->vol.p, c->lG.p ); ++c->launches;
      for (int k=0; k<c->cns; ++k) {             // lohner::grad covers the transported scalars
        k_cho_grad< 1 ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, NP, c->sl_base.p, c->inc_e.p, c->inc_q.p,
          c->D.p, c->nslot, U + (4+k)*NP, c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, c->vol.p,
          c->lG.p + (size_t)(12+3*k)*NP ); ++c->launches;
      } }
    soa_halo( c, c->lG.p, 12 );                     // LohCG::comgrad, LohCG.cpp:1511-1533
    for (int k=0; k<c->cns; ++k) soa_halo( c, c->lG.p + (size_t)(12+3*k)*NP, 3 );
    k_loh_rhs< true ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
      c->nslot, U, c->lG.p, c->X.p, lohp( c ), c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, S, c->v.p, c->vol.p,
      Un, sdt, Uout, R );
  } else
    k_loh_rhs< false ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
      c->nslot, U, c->lG.p, c->X.p, lohp( c ), c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, S, c->v.p, c->vol.p,
      Un, sdt, Uout, R );
  ++c->launches;
  if (c->cns) {        // scalar rows: the Chorin scalar kernel on the velocity-relative rows (velocity rows 0..2, scalars 3..)
    ChoP P{ c->chp.stab, c->chp.stab2, c->chp.stab2coef, c->chp.mu, c->loh_s };
    const double* Sv = S ? S + NP : nullptr;
    if (c->chp.flux == 1)
      k_cho_srhs< true, 3 ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
        c->nslot, c->cU, c->lG.p + 3*NP, c->lG.p + 12*NP, c->X.p, P, c->cdif, c->cns, c->bslot.p, c->bn_off.p,
        c->bn_face.p, c->tri.p, c->fn.p, Sv, c->v.p, c->vol.p, Un ? Un + NP : nullptr, sdt, Uout ? Uout + NP : nullptr, R + NP );
    else
      k_cho_srhs< false, 3 ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
        c->nslot, c->cU, c->lG.p + 3*NP, c->lG.p + 12*NP, c->X.p, P, c->cdif, c->cns, c->bslot.p, c->bn_off.p,
        c->bn_face.p, c->tri.p, c->fn.p, Sv, c->v.p, c->vol.p, Un ? Un + NP : nullptr, sdt, Uout ? Uout + NP : nullptr, R + NP );
    ++c->launches;
  }
  CK( cudaGetLastError() );
  if (Uout) soa_halo_update( c, R, loh_rows( c ), Un, sdt, Uout ); else soa_halo( c, R, loh_rows( c ) );      // LohCG::comrhs, LohCG.cpp:1563-1585
}
// LohCG::solve/solved BCs: dirbc, dirbcp, symbc (pos 1), noslipbc (pos 1)
void loh_bc( xyst_ctx* c, bool pressure ) {
  auto s = c->stream;
  if (c->lb_nd) { k_loh_dirbc<<< nblk( c->lb_nd, 128 ), 128, 0, s >>>( (int)c->lb_nd, c->NP, loh_rows( c ), c->lb_dnode.p, c->lb_dmask.p, c->lb_dval.p, loh_state( c->cU, c->NP ) ); ++c->launches; }
  if (pressure && c->lp_n) { k_loh_pdir<<< nblk( c->lp_n, 128 ), 128, 0, s >>>( (int)c->lp_n, c->lp_node.p, c->lp_val.p, loh_state( c->cU, c->NP ) ); ++c->launches; }
  if (c->cb_ns) { k_cho_symbc<<< nblk( c->cb_ns, 128 ), 128, 0, s >>>( (int)c->cb_ns, c->NP, c->cb_snode.p, c->cb_soff.p, c->cb_snorm.p, c->cU ); ++c->launches; }
  if (c->cb_nn) { k_cho_noslip<<< nblk( c->cb_nn, 128 ), 128, 0, s >>>( (int)c->cb_nn, c->NP, c->cb_nnode.p, c->cU ); ++c->launches; }
  CK( cudaGetLastError() );
}
} // namespace

int xyst_lohcg_mesh_upload( xyst_ctx* c, size_t npoin, const double* x, const double* y, const double* z,
                            const size_t nsup[3], const size_t* const dsupedge[3],
                            const double* const dsupint[3], size_t ntri, const size_t* triinpoel,
                            const double* vol, const double* v, const xyst_lohcg_params* prm )
{
  if (!c || !prm) return fail( "null argument" );
  if (prm->flux != 0 && prm->flux != 1) return fail( "Flux not correctly configured" );
  c->cho = true; c->loh = true;
  std::vector< uint8_t > besym( ntri*3, 0 );
  if (int r = mesh_upload_impl( c, npoin, x, y, z, nsup, dsupedge, dsupint, ntri, triinpoel, besym.data(), vol, v, 4 )) { c->cho = c->loh = false; return r; }
  API_BEGIN
  c->chp = xyst_chocg_params{ prm->flux, prm->stab, prm->stab2, prm->stab2coef, prm->mu };
  c->loh_s = prm->soundspeed;
  size_t NP = c->NP;
  for (auto* b : { &c->cUa, &c->cUb, &c->cUc, &c->cR }) { b->alloc( 4*NP ); CK( cudaMemsetAsync( b->p, 0, 4*NP*sizeof(double), c->stream ) ); }
  for (auto* b : { &c->cSg, &c->cPg, &c->cFl }) { b->alloc( 3*NP ); CK( cudaMemsetAsync( b->p, 0, 3*NP*sizeof(double), c->stream ) ); }
  for (auto* b : { &c->cP, &c->cDiv }) { b->alloc( NP ); CK( cudaMemsetAsync( b->p, 0, NP*sizeof(double), c->stream ) ); }
  c->cVg.alloc( 9*NP ); CK( cudaMemsetAsync( c->cVg.p, 0, 9*NP*sizeof(double), c->stream ) );
  c->lG.alloc( 12*NP ); CK( cudaMemsetAsync( c->lG.p, 0, 12*NP*sizeof(double), c->stream ) );
  c->cS.release();
  c->cU = c->cUa.p + NP; c->cUn = c->cUb.p + NP; c->cUx = c->cUc.p + NP;     // velocity rows of the (p,u,v,w) buffers
  c->cns = 0; c->cdif = 0.0; c->ncpin = 0;
  c->cb_nd = c->cb_ns = c->cb_nn = 0; c->lp_n = c->lb_nd = 0;
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

int xyst_lohcg_bc_upload( xyst_ctx* c, size_t ndir, const size_t* dirnodes, const int* dirmask, const double* dirval,
                          size_t npdir, const size_t* pdirnodes, const double* pdirval,
                          size_t nsym, const size_t* symbcnodes, const double* symbcnorms,
                          size_t nnoslip, const size_t* noslipbcnodes )
{
  if (!c) return fail( "null argument" );
  if (!c->loh) return fail( "LohCG needs stride-4 superedge integrals with the Laplacian term: use xyst_lohcg_mesh_upload" );
  // symmetry and no-slip lists act on the velocity rows exactly as in ChoCG
  if (int r = xyst_chocg_bc_upload( c, 0, nullptr, nullptr, nullptr, nsym, symbcnodes, symbcnorms, nnoslip, noslipbcnodes )) return r;
  API_BEGIN
  auto s = c->stream;
  auto chk = [&]( size_t id ){ if (id >= c->npoin) throw std::runtime_error( "BC node id out of range" ); return (int)id; };
  size_t m = (size_t)loh_rows( c );
  std::vector< int > nd( ndir ), mk( ndir*m ); std::vector< double > vl( ndir*m );
  for (size_t i=0; i<ndir; ++i) { nd[i] = chk( dirnodes[i] );
    for (size_t k=0; k<m; ++k) { mk[i*m+k] = dirmask[i*m+k]; vl[i*m+k] = dirval ? dirval[i*m+k] : 0.0;
      if (mk[i*m+k] == 2 && !dirval) mk[i*m+k] = 0; } }
  c->lb_dnode.upload( nd, s ); c->lb_dmask.upload( mk, s ); c->lb_dval.upload( vl, s ); c->lb_nd = ndir;
  std::vector< int > pn( npdir ); std::vector< double > pv( npdir );
  for (size_t i=0; i<npdir; ++i) { pn[i] = chk( pdirnodes[i] ); pv[i] = pdirval[i]; }
  c->lp_node.upload( pn, s ); c->lp_val.upload( pv, s ); c->lp_n = npdir;
  API_END
}

int xyst_lohcg_set_u( xyst_ctx* c, const double* u ) { API_BEGIN CK( cudaSetDevice( c->device ) ); loh_need( c ); cho_set( c, u, loh_rows( c ), loh_state( c->cU, c->NP ) ); API_END }
int xyst_lohcg_get_u( xyst_ctx* c, double* u ) { API_BEGIN CK( cudaSetDevice( c->device ) ); loh_need( c ); cho_get( c, loh_state( c->cU, c->NP ), loh_rows( c ), u ); API_END }
int xyst_lohcg_get_rhs( xyst_ctx* c, double* r ) { API_BEGIN CK( cudaSetDevice( c->device ) ); loh_need( c ); cho_get( c, c->cR.p, loh_rows( c ), r ); API_END }

// Transported scalars after (p,u,v,w) (LohCG::m_u with problem_ncomp = 4 + ns): call after the mesh upload and
// before any state / BC upload. The state-like buffers get ns more rows, the gradient buffer 3 ns.
int xyst_lohcg_scalars( xyst_ctx* c, int ns, double diffusivity )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  loh_need( c );
  if (ns < 0 || ns > CHO_NSMAX) throw std::runtime_error( "xyst_lohcg_scalars: 0..4 transported scalars" );
  size_t NP = c->NP, m = 4 + (size_t)ns;
  for (auto* b : { &c->cUa, &c->cUb, &c->cUc, &c->cR }) { b->alloc( m*NP ); CK( cudaMemsetAsync( b->p, 0, m*NP*sizeof(double), c->stream ) ); }
  c->lG.alloc( 3*m*NP ); CK( cudaMemsetAsync( c->lG.p, 0, 3*m*NP*sizeof(double), c->stream ) );
  c->cS.release();
  c->cU = c->cUa.p + NP; c->cUn = c->cUb.p + NP; c->cUx = c->cUc.p + NP;
  c->cns = ns; c->cdif = diffusivity; c->ncpin = 0;
  c->lb_nd = 0;
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

// source term values at the nodes, [npoin][4+ns] (lohner::src, Lohner.cpp:1073-1087), or NULL for none
int xyst_lohcg_src( xyst_ctx* c, const double* S )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  loh_need( c );
  if (!S) { c->cS.release(); return 0; }
  size_t m = (size_t)loh_rows( c );
  if (c->cS.n != m*c->NP) { c->cS.alloc( m*c->NP ); CK( cudaMemsetAsync( c->cS.p, 0, m*c->NP*sizeof(double), c->stream ) ); }
  cho_set( c, S, (int)m, c->cS.p );
  API_END
}
int xyst_lohcg_apply_bc( xyst_ctx* c, int pressure ) { API_BEGIN CK( cudaSetDevice( c->device ) ); loh_need( c ); loh_bc( c, pressure != 0 ); API_END }

int xyst_lohcg_rhs( xyst_ctx* c )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  loh_need( c );
  loh_rhs( c, nullptr, 0.0, nullptr, c->cR.p );
  API_END
}

int xyst_lohcg_stage( xyst_ctx* c, int stage, double rkcoef_, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  loh_need( c );
  if (stage < 0) throw std::runtime_error( "stage must be >= 0" );
  size_t NP = c->NP;
  // un = u at stage 0 without a copy: the three (p,u,v,w) buffers rotate (velocity pointers carry them)
  if (stage == 0) { double* old_un = c->cUn; c->cUn = c->cU;
    loh_rhs( c, loh_state( c->cUn, NP ), rkcoef_*dt, loh_state( c->cUx, NP ), nullptr ); c->cU = c->cUx; c->cUx = old_un; }
  else { loh_rhs( c, loh_state( c->cUn, NP ), rkcoef_*dt, loh_state( c->cUx, NP ), nullptr ); std::swap( c->cU, c->cUx ); }
  cho_pin( c );                                     // problems::point_src, LohCG::solve :1615-1617
  loh_bc( c, true );
  API_END
}

// u -= grad(solution of the pressure solve), velocity BCs (LohCG::psolved :1300-1312)
int xyst_lohcg_project( xyst_ctx* c )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  loh_need( c );
  k_cho_project<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, 1.0, c->cSg.p, c->cU ); ++c->launches;
  loh_bc( c, false );
  API_END
}

// p = solution of the pressure solve (LohCG::transferIC :1332,1355)
int xyst_lohcg_pressure_set( xyst_ctx* c )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  loh_need( c ); cho_need_cg( c );
  k_cho_pupdate<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, 0, c->cg_x.p, loh_state( c->cU, c->NP ) ); ++c->launches;
  CK( cudaGetLastError() );
  API_END
}

int xyst_lohcg_dt_min( xyst_ctx* c, double cfl, double dif, double* dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  loh_need( c );
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( c->npoin, RED_THREADS ) );
  double* fin = c->red.p + (size_t)RED_BLOCKS*NDIAG;
  k_loh_dt<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->cU, c->vol.p, c->loh_s, std::max( c->chp.mu, dif ), c->red.p );
  k_reduce_final< 1, true ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, fin );
  c->launches += 2;
  CK( cudaMemcpyAsync( c->red_host, fin, sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  *dt = c->red_host[0] * cfl;
  API_END
}

int xyst_lohcg_diag( xyst_ctx* c, const double* an, double* out )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  loh_need( c );
  DevBuf< double > da;
  if (an) da.upload( std::vector< double >( an, an + c->npoin*(size_t)loh_rows( c ) ), c->stream );
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( c->npoin, RED_THREADS ) );
  double* fin = c->red.p + (size_t)RED_BLOCKS*NDIAG;
  k_loh_diag<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, loh_state( c->cU, c->NP ), loh_state( c->cUn, c->NP ),
    c->v.p, da.p, loh_rows( c ), c->red.p );
  k_reduce_final< 16, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, fin );
  c->launches += 2;
  CK( cudaMemcpyAsync( c->red_host, fin, 16*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  for (int i=0; i<16; ++i) out[i] = c->red_host[i];
  if (c->cns) {            // scalar rows: out[16+4k..] = L2 solution, L2 increment, L2 error, L1 error sums of scalar k
    k_cho_sdiag<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->cns, loh_rows( c ), c->cU, c->cUn, c->v.p, da.p, c->red.p );
    k_reduce_final< 4*CHO_NSMAX, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, fin );
    c->launches += 2;
    CK( cudaMemcpyAsync( c->red_host, fin, 4*CHO_NSMAX*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
    CK( cudaStreamSynchronize( c->stream ) );
    for (int i=0; i<4*c->cns; ++i) out[16+i] = c->red_host[i];
  }
  API_END
}

#endif // XYST_LOHCG_API

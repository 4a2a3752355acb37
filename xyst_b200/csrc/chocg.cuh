// ChoCG device path: the Chorin edge operators of the reference's projection solver for
// constant-density flow (src/Physics/Chorin.cpp:34-1044) and the nodal loops of
// src/Inciter/ChoCG.cpp that sit between them (fingrad, solve/pred update, BC, psolved, dt,
// diagnostics). Included twice by xyst_b200.cu: once for the kernels (XYST_CHOCG_KERNELS, inside
// the anonymous namespace) and once for the C ABI (XYST_CHOCG_API, inside extern "C").
//
// All edge sweeps are node gathers over the sliced-ELL incidence built by mesh_upload_impl:
// the operators are cheap per edge (a handful of FMAs), so each node recomputes the value of its
// incident edges from the two end states instead of reading a materialised per-edge flux. Both
// ends evaluate the identical expression on identical inputs (edge orientation p -> q as
// uploaded), so the two contributions cancel exactly as in the reference's scatter
// (conservation), and no atomics or zero-fill are needed. Boundary-face terms are gathered by
// the same thread through the boundary-node -> (face, local index) CSR.

#ifdef XYST_CHOCG_KERNELS

// edge end nodes in the uploaded orientation: entry se < 0 means this node is the edge's first
// node (receives -f), se > 0 its second (receives +f)
#define CHO_EDGE_LOOP_BEGIN \
  for (int k=0; k<kmax; ++k) { \
    long long ii = base + (long long)k*32 + lane; \
    int se = __ldg( inc_e + ii ); \
    if (se == 0) continue; \
    int nb = __ldg( inc_q + ii ); \
    size_t sl = (size_t)(abs(se)-1); \
    const bool first = se < 0; \
    const size_t a = first ? p : (size_t)nb, b = first ? (size_t)nb : p;
#define CHO_EDGE_LOOP_END }

// crossdiv(ba,ca,6) of a boundary face (Vector.hpp) from the stored cross/12: the factor 2 is exact
__device__ __forceinline__ void cho_face( const int* __restrict__ tri, const double* __restrict__ fn, int f,
                                          int N[3], double n[3] ) {
  N[0] = tri[f*3+0]; N[1] = tri[f*3+1]; N[2] = tri[f*3+2];
  n[0] = 2.0*fn[(size_t)f*3+0]; n[1] = 2.0*fn[(size_t)f*3+1]; n[2] = 2.0*fn[(size_t)f*3+2];
}
// the (6a+b+c)/8 boundary weighting of Chorin.cpp:186-205 for local face node k, terms in A,B,C order
__device__ __forceinline__ double cho_w8( double A, double B, double C, int k ) {
  return k == 0 ? (6.0*A + B + C)/8.0 : (k == 1 ? (A + 6.0*B + C)/8.0 : (A + B + 6.0*C)/8.0);
}

// chorin::div, Chorin.cpp:85-209 (+ the pressure stabilisation of the edge term, :34-83)
template< bool STAB >
__global__ void __launch_bounds__(NODE_THREADS)
k_cho_div( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int* __restrict__ inc_e,
           const int* __restrict__ inc_q, const double* __restrict__ D, size_t nslot,
           const double* __restrict__ V, const double* __restrict__ X, const double* __restrict__ P,
           const double* __restrict__ Pg, double dt, const int* __restrict__ bslot,
           const int* __restrict__ bn_off, const int* __restrict__ bn_face, const int* __restrict__ tri,
           const double* __restrict__ fn, double* __restrict__ out )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc = 0.0;
  CHO_EDGE_LOOP_BEGIN
    double d0 = __ldg( D + sl ), d1 = __ldg( D + nslot + sl ), d2 = __ldg( D + 2*nslot + sl );
    double f = d0 * (V[a] + V[b]) + d1 * (V[NP+a] + V[NP+b]) + d2 * (V[2*NP+a] + V[2*NP+b]);
    if (STAB) {
      double dx = X[a] - X[b], dy = X[NP+a] - X[NP+b], dz = X[2*NP+a] - X[2*NP+b];
      double dl = sqrt( dx*dx + dy*dy + dz*dz );
      double p2 = P[b] - P[a];
      double Dn = sqrt( d0*d0 + d1*d1 + d2*d2 );
      double dpx = Pg[a] + Pg[b], dpy = Pg[NP+a] + Pg[NP+b], dpz = Pg[2*NP+a] + Pg[2*NP+b];
      double p4 = 0.5 * (dx*dpx + dy*dpy + dz*dpz);
      f += Dn*dt/dl*(p2 + p4);
    }
    acc = first ? acc - f : acc + f;
  CHO_EDGE_LOOP_END
  int bs = bslot[p];
  if (bs >= 0)
    for (int i=bn_off[bs]; i<bn_off[bs+1]; ++i) {
      int f = bn_face[i] >> 2, kk = bn_face[i] & 3, N[3]; double n[3];
      cho_face( tri, fn, f, N, n );
      double ux = cho_w8( V[N[0]], V[N[1]], V[N[2]], kk );
      double uy = cho_w8( V[NP+N[0]], V[NP+N[1]], V[NP+N[2]], kk );
      double uz = cho_w8( V[2*NP+N[0]], V[2*NP+N[1]], V[2*NP+N[2]], kk );
      acc += ux*n[0] + uy*n[1] + uz*n[2];
    }
  out[p] = acc;
}

// chorin::grad (:336-449, M = 1) and chorin::vgrad (:211-334, M = 3) followed by the division
// by the nodal volume of ChoCG::fingrad (ChoCG.cpp:840-865) / finpgrad (:1306-1321)
template< int M >
__global__ void __launch_bounds__(NODE_THREADS)
k_cho_grad( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int* __restrict__ inc_e,
            const int* __restrict__ inc_q, const double* __restrict__ D, size_t nslot,
            const double* __restrict__ S, const int* __restrict__ bslot,
            const int* __restrict__ bn_off, const int* __restrict__ bn_face, const int* __restrict__ tri,
            const double* __restrict__ fn, const double* __restrict__ vol, double* __restrict__ G )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[M*3], up[M];
  #pragma unroll
  for (int i=0; i<M*3; ++i) acc[i] = 0.0;
  #pragma unroll
  for (int i=0; i<M; ++i) up[i] = S[i*NP+p];
  CHO_EDGE_LOOP_BEGIN
    (void)a; (void)b;
    double d[3] = { __ldg( D + sl ), __ldg( D + nslot + sl ), __ldg( D + 2*nslot + sl ) };
    double sg = first ? -1.0 : 1.0;
    #pragma unroll
    for (int i=0; i<M; ++i) {
      double s = __ldg( S + i*NP + nb );
      s = first ? s + up[i] : up[i] + s;          // u[second] + u[first]
      #pragma unroll
      for (int j=0; j<3; ++j) acc[i*3+j] += sg * (d[j] * s);
    }
  CHO_EDGE_LOOP_END
  int bs = bslot[p];
  if (bs >= 0)
    for (int i=bn_off[bs]; i<bn_off[bs+1]; ++i) {
      int f = bn_face[i] >> 2, kk = bn_face[i] & 3, N[3]; double n[3];
      cho_face( tri, fn, f, N, n );
      #pragma unroll
      for (int m=0; m<M; ++m) {
        double fv = cho_w8( S[m*NP+N[0]], S[m*NP+N[1]], S[m*NP+N[2]], kk );
        #pragma unroll
        for (int j=0; j<3; ++j) acc[m*3+j] += fv * n[j];
      }
    }
  double vp = vol[p];
  #pragma unroll
  for (int i=0; i<M*3; ++i) G[i*NP+p] = acc[i] / vp;
}

// momentum flux of an edge / of a point, Chorin.cpp:451-509
__device__ __forceinline__ double cho_flux2( const double up[3], const double gp[9], const double uq[3],
                                             const double gq[9], int i, int j, double mu, bool visc ) {
  double inv = up[i]*up[j] + uq[i]*uq[j];
  if (!visc) return -inv;
  double vis = gp[i*3+j] + gp[j*3+i] + gq[i*3+j] + gq[j*3+i];
  if (i == j) vis -= 2.0/3.0 * ( gp[0] + gp[4] + gp[8] + gq[0] + gq[4] + gq[8] );
  return mu*vis - inv;
}
__device__ __forceinline__ double cho_flux1( const double up[3], const double gp[9], int i, int j, double mu, bool visc ) {
  double inv = up[i]*up[j];
  if (!visc) return -inv;
  double vis = gp[i*3+j] + gp[j*3+i];
  if (i == j) vis -= 2.0/3.0 * ( gp[0] + gp[4] + gp[8] );
  return mu*vis - inv;
}

// chorin::flux, Chorin.cpp:511-638: weak (un-normalised) momentum flux
__global__ void __launch_bounds__(NODE_THREADS)
k_cho_flux( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int* __restrict__ inc_e,
            const int* __restrict__ inc_q, const double* __restrict__ D, size_t nslot,
            const double* __restrict__ U, const double* __restrict__ G, double mu,
            const int* __restrict__ bslot, const int* __restrict__ bn_off, const int* __restrict__ bn_face,
            const int* __restrict__ tri, const double* __restrict__ fn, double* __restrict__ F )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  const bool visc = !(mu < 2.220446049250313e-16);
  double acc[3] = { 0.0, 0.0, 0.0 }, um[3], gm[9];
  #pragma unroll
  for (int i=0; i<3; ++i) um[i] = U[i*NP+p];
  #pragma unroll
  for (int i=0; i<9; ++i) gm[i] = visc ? G[i*NP+p] : 0.0;
  CHO_EDGE_LOOP_BEGIN
    (void)a; (void)b;
    double d[3] = { __ldg( D + sl ), __ldg( D + nslot + sl ), __ldg( D + 2*nslot + sl ) };
    double uo[3], go[9];
    #pragma unroll
    for (int i=0; i<3; ++i) uo[i] = __ldg( U + i*NP + nb );
    #pragma unroll
    for (int i=0; i<9; ++i) go[i] = visc ? __ldg( G + i*NP + nb ) : 0.0;
    // flux( U, G, i, j, N[1], N[0] ): first argument the edge's second node
    const double* u2 = first ? uo : um; const double* g2 = first ? go : gm;
    const double* u1 = first ? um : uo; const double* g1 = first ? gm : go;
    #pragma unroll
    for (int i=0; i<3; ++i)
      #pragma unroll
      for (int j=0; j<3; ++j) {
        double f = d[j] * cho_flux2( u2, g2, u1, g1, i, j, mu, visc );
        acc[i] = first ? acc[i] - f : acc[i] + f;
      }
  CHO_EDGE_LOOP_END
  int bs = bslot[p];
  if (bs >= 0)
    for (int i=bn_off[bs]; i<bn_off[bs+1]; ++i) {
      int f = bn_face[i] >> 2, kk = bn_face[i] & 3, N[3]; double n[3];
      cho_face( tri, fn, f, N, n );
      double uu[3][3], gg[3][9];
      #pragma unroll
      for (int m=0; m<3; ++m) {
        #pragma unroll
        for (int c=0; c<3; ++c) uu[m][c] = U[c*NP+N[m]];
        #pragma unroll
        for (int c=0; c<9; ++c) gg[m][c] = visc ? G[c*NP+N[m]] : 0.0;
      }
      #pragma unroll
      for (int c=0; c<3; ++c) {
        double fl[3];
        #pragma unroll
        for (int j=0; j<3; ++j)
          fl[j] = cho_w8( cho_flux1( uu[0], gg[0], c, j, mu, visc ), cho_flux1( uu[1], gg[1], c, j, mu, visc ),
                          cho_flux1( uu[2], gg[2], c, j, mu, visc ), kk );
        acc[c] += fl[0]*n[0] + fl[1]*n[1] + fl[2]*n[2];
      }
    }
  #pragma unroll
  for (int i=0; i<3; ++i) F[i*NP+p] = acc[i];
}

struct ChoP { int stab, stab2; double stab2coef, mu, s; };      // s: LohCG's artificial sound speed (scalar kernel only)

// advection edge flux: second-order damping (Chorin.cpp:640-709) or fourth-order damping with
// the limited reconstruction (:711-829); velocity components only
template< bool DAMP4 >
__device__ __forceinline__ void cho_adv( const double d[3], double lap, const double ua[3], const double ub[3],
    const double ga[9], const double gb[9], const double dx[3], double pa, double pb, const ChoP& C, double f[3] )
{
  double uL[3] = { ua[0], ua[1], ua[2] }, uR[3] = { ub[0], ub[1], ub[2] };
  if (DAMP4) {
    #pragma unroll
    for (int c=0; c<3; ++c) {
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
  double vnL = uL[0]*d[0] + uL[1]*d[1] + uL[2]*d[2];
  double vnR = uR[0]*d[0] + uR[1]*d[1] + uR[2]*d[2];
  double aw = 0.0;
  if (C.stab) aw = fabs( vnL + vnR ) / 2.0;
  if (C.stab2) aw += C.stab2coef * fmax( fabs(vnL), fabs(vnR) );
  double v = lap * C.mu;
  double pf = pa + pb;
  if (DAMP4) {
    #pragma unroll
    for (int c=0; c<3; ++c) f[c] = uL[c]*vnL + uR[c]*vnR + pf*d[c] + aw*(uR[c]-uL[c]) - v*(ub[c]-ua[c]);
  } else {
    #pragma unroll
    for (int c=0; c<3; ++c) f[c] = uL[c]*vnL + uR[c]*vnR + pf*d[c] + (aw-v)*(uR[c]-uL[c]);
  }
}

// chorin::rhs (Chorin.cpp:831-1044) gathered per node; with Uout the explicit update of
// ChoCG::solve (ChoCG.cpp:1529-1545, u = un - rk dt rhs / vol) is applied in the same pass
template< bool DAMP4 >
__global__ void __launch_bounds__(NODE_THREADS)
k_cho_rhs( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int* __restrict__ inc_e,
           const int* __restrict__ inc_q, const double* __restrict__ D, size_t nslot,
           const double* __restrict__ U, const double* __restrict__ P, const double* __restrict__ G,
           const double* __restrict__ X, ChoP C, const int* __restrict__ bslot, const int* __restrict__ bn_off,
           const int* __restrict__ bn_face, const int* __restrict__ tri, const double* __restrict__ fn,
           const double* __restrict__ S, const double* __restrict__ v, const double* __restrict__ vol,
           const double* __restrict__ Un, double sdt, double* __restrict__ Uout, double* __restrict__ R )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[3] = { 0.0, 0.0, 0.0 }, um[3], gm[9], xm[3], pm = P[p];
  #pragma unroll
  for (int i=0; i<3; ++i) { um[i] = U[i*NP+p]; xm[i] = DAMP4 ? X[i*NP+p] : 0.0; }
  #pragma unroll
  for (int i=0; i<9; ++i) gm[i] = DAMP4 ? G[i*NP+p] : 0.0;
  CHO_EDGE_LOOP_BEGIN
    (void)a; (void)b;
    double d[3] = { __ldg( D + sl ), __ldg( D + nslot + sl ), __ldg( D + 2*nslot + sl ) };
    double lap = __ldg( D + 4*nslot + sl );
    double uo[3], go[9], xo[3], po = __ldg( P + nb );
    #pragma unroll
    for (int i=0; i<3; ++i) { uo[i] = __ldg( U + i*NP + nb ); xo[i] = DAMP4 ? __ldg( X + i*NP + nb ) : 0.0; }
    #pragma unroll
    for (int i=0; i<9; ++i) go[i] = DAMP4 ? __ldg( G + i*NP + nb ) : 0.0;
    double f[3];
    if (DAMP4) {
      // one evaluation for both orientations: the end states are ordered by selects, so that lanes
      // with different edge orientations do not diverge around the (expensive) limited reconstruction
      double ua[3], ub[3], ga[9], gb[9], dx[3];
      #pragma unroll
      for (int i=0; i<3; ++i) { ua[i] = first ? um[i] : uo[i]; ub[i] = first ? uo[i] : um[i]; }
      #pragma unroll
      for (int i=0; i<9; ++i) { ga[i] = first ? gm[i] : go[i]; gb[i] = first ? go[i] : gm[i]; }
      #pragma unroll
      for (int i=0; i<3; ++i) dx[i] = first ? xo[i]-xm[i] : xm[i]-xo[i];
      cho_adv< DAMP4 >( d, lap, ua, ub, ga, gb, dx, first ? pm : po, first ? po : pm, C, f );
      #pragma unroll
      for (int c=0; c<3; ++c) acc[c] = first ? acc[c] - f[c] : acc[c] + f[c];
    } else if (first) {
      double dx[3] = { xo[0]-xm[0], xo[1]-xm[1], xo[2]-xm[2] };
      cho_adv< DAMP4 >( d, lap, um, uo, gm, go, dx, pm, po, C, f );
      #pragma unroll
      for (int c=0; c<3; ++c) acc[c] -= f[c];
    } else {
      double dx[3] = { xm[0]-xo[0], xm[1]-xo[1], xm[2]-xo[2] };
      cho_adv< DAMP4 >( d, lap, uo, um, go, gm, dx, po, pm, C, f );
      #pragma unroll
      for (int c=0; c<3; ++c) acc[c] += f[c];
    }
  CHO_EDGE_LOOP_END
  int bs = bslot[p];
  if (bs >= 0)
    for (int i=bn_off[bs]; i<bn_off[bs+1]; ++i) {
      int f = bn_face[i] >> 2, kk = bn_face[i] & 3, N[3]; double n[3];
      cho_face( tri, fn, f, N, n );
      double fl[3][3];
      #pragma unroll
      for (int m=0; m<3; ++m) {
        double u = U[N[m]], vv = U[NP+N[m]], w = U[2*NP+N[m]], pr = P[N[m]];
        double vn = n[0]*u + n[1]*vv + n[2]*w;
        fl[0][m] = u*vn + pr*n[0];
        fl[1][m] = vv*vn + pr*n[1];
        fl[2][m] = w*vn + pr*n[2];
      }
      #pragma unroll
      for (int c=0; c<3; ++c) acc[c] += cho_w8( fl[c][0], fl[c][1], fl[c][2], kk );
    }
  if (S) {
    double vp = v[p];
    #pragma unroll
    for (int c=0; c<3; ++c) acc[c] -= S[c*NP+p] * vp;
  }
  if (R) {
    #pragma unroll
    for (int c=0; c<3; ++c) R[c*NP+p] = acc[c];
  }
  if (Uout) {
    double vp = vol[p];
    #pragma unroll
    for (int c=0; c<3; ++c) Uout[c*NP+p] = Un[c*NP+p] - sdt*acc[c]/vp;
  }
}

// ---- transported scalars (U rows 3.., G rows 9.. of the same buffers): the scalar rows of the
// advection flux (adv_damp2 Chorin.cpp:698-708, adv_damp4 :817-824), of the boundary integral
// (:947-981) and of the source (:1003-1007), and the explicit update of those rows. The normal
// velocities and the stabilisation speed are the flow flux's, recomputed here from the same inputs
// by the same expressions, so that the velocity kernel stays as it is.
constexpr int CHO_NSMAX = 4;
template< bool DAMP4, bool LOH >
__device__ __forceinline__ void cho_vn( const double d[3], const double ua[3], const double ub[3],
    const double ga[9], const double gb[9], const double dx[3], const ChoP& C, double& vnL, double& vnR, double& aw )
{
  double uL[3] = { ua[0], ua[1], ua[2] }, uR[3] = { ub[0], ub[1], ub[2] };
  if (DAMP4) {
    #pragma unroll
    for (int c=0; c<3; ++c) {
      double g1 = ga[c*3+0]*dx[0] + ga[c*3+1]*dx[1] + ga[c*3+2]*dx[2];
      double g2 = gb[c*3+0]*dx[0] + gb[c*3+1]*dx[1] + gb[c*3+2]*dx[2];
      double delta2 = uR[c] - uL[c];
      double delta1 = 2.0 * g1 - delta2;
      double delta3 = 2.0 * g2 - delta2;
      double incL, incR;
      vanleer< false >( delta1, delta2, delta3, incL, incR );
      uL[c] += incL;
      uR[c] -= incR;
    }
  }
  vnL = uL[0]*d[0] + uL[1]*d[1] + uL[2]*d[2];
  vnR = uR[0]*d[0] + uR[1]*d[1] + uR[2]*d[2];
  aw = 0.0;
  if (C.stab) aw = fabs( vnL + vnR ) / 2.0;
  if (C.stab2) {
    double sl = fabs(vnL), sr = fabs(vnR);
    if (LOH) {                                  // Lohner.cpp:772-777,888-893
      double len = sqrt( d[0]*d[0] + d[1]*d[1] + d[2]*d[2] );
      sl += C.s*len; sr += C.s*len;
    }
    aw += C.stab2coef * fmax( sl, sr );
  }
}

// U, Un, Uout, R, S: velocity rows 0..2, scalar rows 3..3+ns-1 (stride NP); G: rows 9+3k.. hold the
// gradient of scalar k (damp4). LAPROW: row of the Laplacian term in D (4: ChoCG, 3: LohCG; LohCG also has its
// own stab2 speed and the same scalar flux, Lohner.cpp:787-793,906-911, boundary term :1030-1056).
template< bool DAMP4, int LAPROW >
__global__ void __launch_bounds__(NODE_THREADS)
k_cho_srhs( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int* __restrict__ inc_e,
            const int* __restrict__ inc_q, const double* __restrict__ D, size_t nslot,
            const double* __restrict__ U, const double* __restrict__ G, const double* __restrict__ SG,
            const double* __restrict__ X, ChoP C, double dif, int ns,
            const int* __restrict__ bslot, const int* __restrict__ bn_off,
            const int* __restrict__ bn_face, const int* __restrict__ tri, const double* __restrict__ fn,
            const double* __restrict__ S, const double* __restrict__ v, const double* __restrict__ vol,
            const double* __restrict__ Un, double sdt, double* __restrict__ Uout, double* __restrict__ R )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[CHO_NSMAX], sm[CHO_NSMAX], um[3], gm[9], xm[3];
  #pragma unroll
  for (int k=0; k<CHO_NSMAX; ++k) { acc[k] = 0.0; sm[k] = k < ns ? U[(3+k)*NP+p] : 0.0; }
  #pragma unroll
  for (int i=0; i<3; ++i) { um[i] = U[i*NP+p]; xm[i] = DAMP4 ? X[i*NP+p] : 0.0; }
  #pragma unroll
  for (int i=0; i<9; ++i) gm[i] = DAMP4 ? G[i*NP+p] : 0.0;
  CHO_EDGE_LOOP_BEGIN
    (void)a; (void)b;
    double d[3] = { __ldg( D + sl ), __ldg( D + nslot + sl ), __ldg( D + 2*nslot + sl ) };
    double df = __ldg( D + LAPROW*nslot + sl ) * dif;
    double uo[3], go[9], xo[3];
    #pragma unroll
    for (int i=0; i<3; ++i) { uo[i] = __ldg( U + i*NP + nb ); xo[i] = DAMP4 ? __ldg( X + i*NP + nb ) : 0.0; }
    #pragma unroll
    for (int i=0; i<9; ++i) go[i] = DAMP4 ? __ldg( G + i*NP + nb ) : 0.0;
    // end states in the edge's orientation (a = first node)
    double ua[3], ub[3], ga[9], gb[9], dx[3], vnL, vnR, aw;
    #pragma unroll
    for (int i=0; i<3; ++i) { ua[i] = first ? um[i] : uo[i]; ub[i] = first ? uo[i] : um[i]; }
    #pragma unroll
    for (int i=0; i<9; ++i) { ga[i] = first ? gm[i] : go[i]; gb[i] = first ? go[i] : gm[i]; }
    #pragma unroll
    for (int i=0; i<3; ++i) dx[i] = first ? xo[i]-xm[i] : xm[i]-xo[i];
    cho_vn< DAMP4, LAPROW == 3 >( d, ua, ub, ga, gb, dx, C, vnL, vnR, aw );
    #pragma unroll
    for (int k=0; k<CHO_NSMAX; ++k) if (k < ns) {
      double so = __ldg( U + (3+k)*NP + nb );
      double sa = first ? sm[k] : so, sb = first ? so : sm[k];
      double f;
      if (DAMP4) {
        double gsm[3], gso[3];
        #pragma unroll
        for (int j=0; j<3; ++j) { gsm[j] = SG[(3*k+j)*NP+p]; gso[j] = __ldg( SG + (3*k+j)*NP + nb ); }
        double g1 = (first ? gsm[0] : gso[0])*dx[0] + (first ? gsm[1] : gso[1])*dx[1] + (first ? gsm[2] : gso[2])*dx[2];
        double g2 = (first ? gso[0] : gsm[0])*dx[0] + (first ? gso[1] : gsm[1])*dx[1] + (first ? gso[2] : gsm[2])*dx[2];
        double delta2 = sb - sa;
        double delta1 = 2.0 * g1 - delta2;
        double delta3 = 2.0 * g2 - delta2;
        double incL, incR;
        vanleer< false >( delta1, delta2, delta3, incL, incR );
        double sL = sa + incL, sR = sb - incR;
        f = sL*vnL + sR*vnR + aw*(sR-sL) - df*(sb-sa);
      } else
        f = sa*vnL + sb*vnR + (aw-df)*(sb-sa);
      acc[k] = first ? acc[k] - f : acc[k] + f;
    }
  CHO_EDGE_LOOP_END
  int bs = bslot[p];
  if (bs >= 0)
    for (int i=bn_off[bs]; i<bn_off[bs+1]; ++i) {
      int f = bn_face[i] >> 2, kk = bn_face[i] & 3, N[3]; double n[3], vn[3];
      cho_face( tri, fn, f, N, n );
      #pragma unroll
      for (int m=0; m<3; ++m) vn[m] = n[0]*U[N[m]] + n[1]*U[NP+N[m]] + n[2]*U[2*NP+N[m]];
      #pragma unroll
      for (int k=0; k<CHO_NSMAX; ++k) if (k < ns)
        acc[k] += cho_w8( U[(3+k)*NP+N[0]]*vn[0], U[(3+k)*NP+N[1]]*vn[1], U[(3+k)*NP+N[2]]*vn[2], kk );
    }
  double vp = vol[p], vo = v[p];
  #pragma unroll
  for (int k=0; k<CHO_NSMAX; ++k) if (k < ns) {
    double r = acc[k];
    if (S) r -= S[(3+k)*NP+p] * vo;
    if (R) R[(3+k)*NP+p] = r;
    if (Uout) Uout[(3+k)*NP+p] = Un[(3+k)*NP+p] - sdt*r/vp;
  }
}

// problems::point_src (Problems.cpp:764-823): the scalar of the listed nodes is held at a value
__global__ void k_cho_pin( int n, const int* __restrict__ node, double val, double* __restrict__ s )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) s[ node[i] ] = val;
}

// NodeDiagnostics::precompute sums of the scalar rows: per scalar k [4k..4k+3] = L2 of the solution,
// of the increment, and with an analytic solution the L2 and L1 error sums
__global__ void __launch_bounds__(RED_THREADS)
k_cho_sdiag( size_t npoin, size_t NP, int ns, int ncomp, const double* __restrict__ U, const double* __restrict__ Un,
             const double* __restrict__ v, const double* __restrict__ an, double* __restrict__ part )
{
  double a[4*CHO_NSMAX];
  #pragma unroll
  for (int i=0; i<4*CHO_NSMAX; ++i) a[i] = 0.0;
  for (size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p < npoin; p += (size_t)gridDim.x*blockDim.x) {
    double vp = v[p];
    #pragma unroll
    for (int k=0; k<CHO_NSMAX; ++k) if (k < ns) {
      double u = U[(3+k)*NP+p], du = u - Un[(3+k)*NP+p];
      a[4*k] += u*u*vp;
      a[4*k+1] += du*du*vp;
      if (an) { double e = u - an[p*ncomp+(ncomp-ns)+k]; a[4*k+2] += e*e*vp; a[4*k+3] += fabs(e)*vp; }
    }
  }
  block_reduce< 4*CHO_NSMAX, false >( a, part );
}

// ---- BCs, ChoCG::BC (ChoCG.cpp:1340-1353) -> physics::dirbc (BC.cpp:29-72), symbc (:110-136),
// noslipbc (:138-150), applied in this order by three launches on the same stream
__global__ void k_cho_dirbc( int nd, size_t NP, int m, const int* __restrict__ node, const int* __restrict__ mask,
                             const double* __restrict__ val, double* __restrict__ U )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nd) return;
  size_t p = node[i];
  for (int c=0; c<m; ++c) if (mask[i*m+c]) U[c*NP+p] = val[i*m+c];
}
// one thread per distinct node: its side sets' normals are applied one after the other
__global__ void k_cho_symbc( int ns, size_t NP, const int* __restrict__ node, const int* __restrict__ off,
                             const double* __restrict__ nrm, double* __restrict__ U )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= ns) return;
  size_t p = node[i];
  double u = U[p], v = U[NP+p], w = U[2*NP+p];
  for (int k=off[i]; k<off[i+1]; ++k) {
    double n0 = nrm[k*3+0], n1 = nrm[k*3+1], n2 = nrm[k*3+2];
    double vn = u*n0 + v*n1 + w*n2;
    u -= vn * n0; v -= vn * n1; w -= vn * n2;
  }
  U[p] = u; U[NP+p] = v; U[2*NP+p] = w;
}
__global__ void k_cho_noslip( int nn, size_t NP, const int* __restrict__ node, double* __restrict__ U )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nn) return;
  size_t p = node[i];
  U[p] = 0.0; U[NP+p] = 0.0; U[2*NP+p] = 0.0;
}

// F /= vol (ChoCG::div :879-881, fingrad of the momentum flux)
__global__ void k_cho_divvol( size_t npoin, size_t NP, int m, const double* __restrict__ vol, double* __restrict__ F )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  double vp = vol[p];
  for (int c=0; c<m; ++c) F[c*NP+p] /= vp;
}
// u -= pdt * sgrad (ChoCG::psolved :1204-1213)
__global__ void k_cho_project( size_t npoin, size_t NP, double pdt, const double* __restrict__ Sg, double* __restrict__ U )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  #pragma unroll
  for (int c=0; c<3; ++c) U[c*NP+p] -= pdt * Sg[c*NP+p];
}
// pr = x (:1241) or pr += x (:1249)
__global__ void k_cho_pupdate( size_t npoin, int increment, const double* __restrict__ x, double* __restrict__ P )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  P[p] = increment ? P[p] + x[p] : x[p];
}
// Poisson rhs: div / divisor (ChoCG::pinit :1044)
__global__ void k_cho_prhs( size_t npoin, double divisor, const double* __restrict__ div, double* __restrict__ b )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  b[p] = div[p] / divisor;
}

// ChoCG::dt :1356-1411
__global__ void __launch_bounds__(RED_THREADS)
k_cho_dt( size_t npoin, size_t NP, const double* __restrict__ U, const double* __restrict__ vol, double dif,
          double* __restrict__ part )
{
  double m[1] = { 1.7976931348623157e308 };
  for (size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p < npoin; p += (size_t)gridDim.x*blockDim.x) {
    double u = U[p], v = U[NP+p], w = U[2*NP+p];
    double vel = sqrt( u*u + v*v + w*w );
    double L = cbrt( vol[p] );
    m[0] = fmin( m[0], L / fmax( vel, 1.0e-8 ) );
    if (dif > 2.220446049250313e-16) m[0] = fmin( m[0], L * L / dif );
  }
  block_reduce< 1, true >( m, part );
}

// NodeDiagnostics::precompute sums, NodeDiagnostics.cpp:223-252
constexpr int NCHODIAG = 16;
__global__ void __launch_bounds__(RED_THREADS)
k_cho_diag( size_t npoin, size_t NP, const double* __restrict__ U, const double* __restrict__ Un,
            const double* __restrict__ P, const double* __restrict__ dp, const double* __restrict__ v,
            const double* __restrict__ anp, const double* __restrict__ anu, int anstride, double* __restrict__ part )
{
  double a[NCHODIAG];
  #pragma unroll
  for (int i=0; i<NCHODIAG; ++i) a[i] = 0.0;
  for (size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p < npoin; p += (size_t)gridDim.x*blockDim.x) {
    double vp = v[p], pr = P[p], d = dp[p];
    a[0] += pr*pr*vp;
    a[4] += d*d*vp;
    #pragma unroll
    for (int c=0; c<3; ++c) {
      double u = U[c*NP+p], du = u - Un[c*NP+p];
      a[1+c] += u*u*vp;
      a[5+c] += du*du*vp;
      if (anu) { double e = u - anu[p*anstride+c]; a[10+c] += e*e*vp; a[13+c] += fabs(e)*vp; }
    }
    if (anp) { double e = pr - anp[p]; a[8] += e*e*vp; a[9] += fabs(e)*vp; }
  }
  block_reduce< NCHODIAG, false >( a, part );
}

// right-hand side of the momentum solve: b[i*3+c] = R[c][i] (ChoCG::solve :1603, m_rhs.vec())
__global__ void k_cho_mrhs( size_t npoin, size_t NP, const double* __restrict__ R, double* __restrict__ b )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= npoin*3) return;
  size_t p = i / 3; int cc = (int)(i % 3);
  b[i] = R[cc*NP+p];
}
// u = un + du (ChoCG::msolved :1636-1641)
__global__ void k_cho_mupdate( size_t npoin, size_t NP, const double* __restrict__ Un, const double* __restrict__ du,
                               double* __restrict__ U )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= npoin*3) return;
  size_t p = i / 3; int cc = (int)(i % 3);
  U[cc*NP+p] = Un[cc*NP+p] + du[i];
}
__global__ void k_cg_bc_scatter( int nbc, const int* __restrict__ node, const double* __restrict__ v, double* __restrict__ bcval )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < nbc) bcval[ node[i] ] = v[i];
}

// ---- several partitions: the node gathers above are this partition's own sums at the nodes it
// shares (ChoCG::div/velgrad/flux/sgrad/pgrad/rhs send m_div[...] etc. of the chare-boundary
// nodes, ChoCG.cpp:889-898,936-945,989-998,1157-1167,1269-1279,1492-1501). The shared nodes'
// values of a [m][NP] field go to the exchange buffer, and come back as own + the sharers' in
// the fixed neighbour order (comdiv :903-920 and its siblings)
__global__ void k_soa_shared_get( int nsh, int m, size_t NP, const int* __restrict__ sh_node,
                                  const double* __restrict__ S, double* __restrict__ part )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nsh*m) return;
  part[i] = S[(i%m)*NP + (size_t)sh_node[i/m]];
}
__global__ void k_soa_shared_put( int nsh, int m, size_t NP, const int* __restrict__ sh_node,
                                  const int* __restrict__ roff, const int* __restrict__ ridx,
                                  const double* __restrict__ part, const double* __restrict__ recvbuf,
                                  double* __restrict__ S )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nsh*m) return;
  size_t s = i / m, k = i % m;
  double a = part[i];
  for (int r=roff[s]; r<roff[s+1]; ++r) a += recvbuf[(size_t)ridx[r]*m + k];
  S[k*NP + (size_t)sh_node[s]] = a;
}
// the same for the right-hand side of a Runge-Kutta stage, followed by the update of the shared
// nodes the gather kernel could not finish: u = un - rk dt (sum of the rhs parts) / vol
// (ChoCG::solve :1529-1545 after comrhs, LohCG::solve likewise)
__global__ void k_soa_shared_update( int nsh, int m, size_t NP, const int* __restrict__ sh_node,
                                     const int* __restrict__ roff, const int* __restrict__ ridx,
                                     const double* __restrict__ part, const double* __restrict__ recvbuf,
                                     const double* __restrict__ vol, const double* __restrict__ Un, double sdt,
                                     double* __restrict__ Uout, double* __restrict__ R )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nsh*m) return;
  size_t s = i / m, k = i % m;
  double a = part[i];
  for (int r=roff[s]; r<roff[s+1]; ++r) a += recvbuf[(size_t)ridx[r]*m + k];
  size_t p = (size_t)sh_node[s];
  R[k*NP+p] = a;
  Uout[k*NP+p] = Un[k*NP+p] - sdt*a/vol[p];
}

// reference layout [node][m] <-> structure of arrays with stride NP
__global__ void k_aos_to_soa( size_t n, size_t NP, int m, const double* __restrict__ A, double* __restrict__ S )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= n*m) return;
  size_t p = i / m; int c = (int)(i % m);
  S[c*NP+p] = A[i];
}
__global__ void k_soa_to_aos( size_t n, size_t NP, int m, const double* __restrict__ S, double* __restrict__ A )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= n*m) return;
  size_t p = i / m; int c = (int)(i % m);
  A[i] = S[c*NP+p];
}

#undef CHO_EDGE_LOOP_BEGIN
#undef CHO_EDGE_LOOP_END
#endif // XYST_CHOCG_KERNELS

#ifdef XYST_CHOCG_API

namespace {
void cho_need( xyst_ctx* c ) {
  need_mesh( c );
  if (!c->cho) throw std::runtime_error( "ChoCG needs stride-5 superedge integrals: use xyst_chocg_mesh_upload" );
}
// several partitions: every node gather is followed by the sum over the partitions sharing a node
bool cho_parts( const xyst_ctx* c ) { return c->nsh > 0 && c->comm; }
void soa_halo( xyst_ctx* c, double* S, int m ) {
  if (!cho_parts( c )) return;
  if ((size_t)m*c->nsh > c->sh_part.n) throw std::runtime_error( "shared-node buffer too small" );
  unsigned g = nblk( c->nsh*(size_t)m, 256 );
  k_soa_shared_get<<< g, 256, 0, c->stream >>>( (int)c->nsh, m, c->NP, c->sh_node.p, S, c->sh_part.p ); ++c->launches;
  exchange( c, m );
  exchange_wait( c );
  k_soa_shared_put<<< g, 256, 0, c->stream >>>( (int)c->nsh, m, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p,
    c->sh_part.p, c->sh_recvbuf.p, S ); ++c->launches;
  CK( cudaGetLastError() );
}
// the stage update of the shared nodes from the summed right-hand side parts (R, Un, Uout: [m][NP])
void soa_halo_update( xyst_ctx* c, double* R, int m, const double* Un, double sdt, double* Uout ) {
  if (!cho_parts( c )) return;
  unsigned g = nblk( c->nsh*(size_t)m, 256 );
  k_soa_shared_get<<< g, 256, 0, c->stream >>>( (int)c->nsh, m, c->NP, c->sh_node.p, R, c->sh_part.p ); ++c->launches;
  exchange( c, m );
  exchange_wait( c );
  k_soa_shared_update<<< g, 256, 0, c->stream >>>( (int)c->nsh, m, c->NP, c->sh_node.p, c->sh_roff.p, c->sh_ridx.p,
    c->sh_part.p, c->sh_recvbuf.p, c->vol.p, Un, sdt, Uout, R ); ++c->launches;
  CK( cudaGetLastError() );
}
void cho_need_cg( xyst_ctx* c ) {
  cho_need( c );
  if (c->cg_nrow != c->npoin) throw std::runtime_error( "ChoCG: upload the pressure Poisson matrix (one row per node) with xyst_csr_upload first" );
}
unsigned cho_grid( xyst_ctx* c ) { return nblk( c->nslice*32, NODE_THREADS ); }
// rows of the velocity-like buffers: 3 velocities + transported scalars (ChoCG::m_u.nprop())
int cho_rows( const xyst_ctx* c ) { return 3 + c->cns; }
ChoP chop( const xyst_ctx* c ) { return ChoP{ c->chp.stab, c->chp.stab2, c->chp.stab2coef, c->chp.mu }; }

void cho_set( xyst_ctx* c, const double* host, int m, double* soa ) {
  DevBuf< double > tmp; tmp.alloc( c->npoin*(size_t)m );
  CK( cudaMemcpyAsync( tmp.p, host, c->npoin*(size_t)m*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  k_aos_to_soa<<< nblk( c->npoin*(size_t)m, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, m, tmp.p, soa ); ++c->launches;
  CK( cudaGetLastError() );
  CK( cudaStreamSynchronize( c->stream ) );
}
void cho_get( xyst_ctx* c, const double* soa, int m, double* host ) {
  DevBuf< double > tmp; tmp.alloc( c->npoin*(size_t)m );
  k_soa_to_aos<<< nblk( c->npoin*(size_t)m, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, m, soa, tmp.p ); ++c->launches;
  CK( cudaGetLastError() );
  CK( cudaMemcpyAsync( host, tmp.p, c->npoin*(size_t)m*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
}
// ChoCG::BC on a velocity-like field: dirbc (only on the velocity itself), symbc, noslipbc
void cho_bc( xyst_ctx* c, double* U, bool dir, bool noslip ) {
  auto s = c->stream;
  if (dir && c->cb_nd) { k_cho_dirbc<<< nblk( c->cb_nd, 128 ), 128, 0, s >>>( (int)c->cb_nd, c->NP, cho_rows( c ), c->cb_dnode.p, c->cb_dmask.p, c->cb_dval.p, U ); ++c->launches; }
  if (c->cb_ns) { k_cho_symbc<<< nblk( c->cb_ns, 128 ), 128, 0, s >>>( (int)c->cb_ns, c->NP, c->cb_snode.p, c->cb_soff.p, c->cb_snorm.p, U ); ++c->launches; }
  if (noslip && c->cb_nn) { k_cho_noslip<<< nblk( c->cb_nn, 128 ), 128, 0, s >>>( (int)c->cb_nn, c->NP, c->cb_nnode.p, U ); ++c->launches; }
  CK( cudaGetLastError() );
}
void cho_vgrad( xyst_ctx* c ) {
  ProfScope ps( c, "cho_vgrad" );
  k_cho_grad< 3 ><<< cho_grid( c ), NODE_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p,
    c->D.p, c->nslot, c->cU, c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, c->vol.p, c->cVg.p ); ++c->launches;
  const int nsg = c->loh ? 0 : c->cns;              // (LohCG keeps its scalars' gradients in lG: loh_rhs)
  for (int k=0; k<nsg; ++k) {                       // chorin::vgrad covers the transported scalars (Chorin.cpp:230)
    k_cho_grad< 1 ><<< cho_grid( c ), NODE_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p,
      c->D.p, c->nslot, c->cU + (3+k)*c->NP, c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, c->vol.p,
      c->cVg.p + (size_t)(9+3*k)*c->NP ); ++c->launches;
  }
  CK( cudaGetLastError() );
  soa_halo( c, c->cVg.p, 9 );                       // comvgrad :949-970 (each part already over the full volume)
  for (int k=0; k<nsg; ++k) soa_halo( c, c->cVg.p + (size_t)(9+3*k)*c->NP, 3 );
}
void cho_rhs( xyst_ctx* c, const double* Un, double sdt, double* Uout, double* R ) {
  if (c->loh) throw std::runtime_error( "context holds a LohCG mesh: use xyst_lohcg_rhs / xyst_lohcg_stage" );
  ProfScope ps( c, "cho_rhs" );
  auto g = cho_grid( c );
  if (cho_parts( c ) && !R) R = c->cR.p;            // the shared nodes' parts travel before they are used
  if (c->chp.flux == 1)
    k_cho_rhs< true ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
      c->nslot, c->cU, c->cP.p, c->cVg.p, c->X.p, chop( c ), c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p,
      c->cS.p, c->v.p, c->vol.p, Un, sdt, Uout, R );
  else
    k_cho_rhs< false ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
      c->nslot, c->cU, c->cP.p, c->cVg.p, c->X.p, chop( c ), c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p,
      c->cS.p, c->v.p, c->vol.p, Un, sdt, Uout, R );
  ++c->launches;
  if (c->cns) {
    if (c->chp.flux == 1)
      k_cho_srhs< true, 4 ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
        c->nslot, c->cU, c->cVg.p, c->cVg.p + 9*c->NP, c->X.p, chop( c ), c->cdif, c->cns, c->bslot.p, c->bn_off.p,
        c->bn_face.p, c->tri.p, c->fn.p, c->cS.p, c->v.p, c->vol.p, Un, sdt, Uout, R );
    else
      k_cho_srhs< false, 4 ><<< g, NODE_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
        c->nslot, c->cU, c->cVg.p, c->cVg.p + 9*c->NP, c->X.p, chop( c ), c->cdif, c->cns, c->bslot.p, c->bn_off.p,
        c->bn_face.p, c->tri.p, c->fn.p, c->cS.p, c->v.p, c->vol.p, Un, sdt, Uout, R );
    ++c->launches;
  }
  CK( cudaGetLastError() );
  if (Uout) soa_halo_update( c, R, cho_rows( c ), Un, sdt, Uout ); else soa_halo( c, R, cho_rows( c ) );      // comrhs :1505-1527
}
// problems::point_src through ChoCG::pred :1655-1657 / LohCG::pred :1615-1617: first scalar row
void cho_pin( xyst_ctx* c ) {
  if (!c->ncpin) return;
  k_cho_pin<<< nblk( c->ncpin, 128 ), 128, 0, c->stream >>>( (int)c->ncpin, c->cpin.p, c->cpin_val, c->cU + 3*c->NP ); ++c->launches;
}
// ConjugateGradients::init :336-428 + apply :451-505 + r :508-556 for one partition: Dirichlet rows
// (scalar row ids of the selected solver) with values, optional Neumann vector, applied to cg_b and,
// on the fly, to the matrix
void cho_cg_bc( xyst_ctx* c, size_t nbc, const size_t* bcnodes, const double* bcvals, const double* neubc ) {
  auto s = c->stream;
  size_t n = c->cg_nrow;
  { // the BC node set rarely changes between solves: the row mask is rebuilt only when it does,
    // the values are scattered from an nbc-long upload
    std::vector< size_t > nodes( bcnodes, bcnodes + nbc );
    if (!c->cg_bc.p || c->cg_bc.n != n || nodes != c->cg_bcnodes_h) {
      std::vector< unsigned char > bc( n, 0 ); std::vector< int > nd( nbc );
      for (size_t i=0; i<nbc; ++i) { if (bcnodes[i] >= n) throw std::runtime_error( "pressure BC node out of range" );
        bc[ bcnodes[i] ] = 1; nd[i] = (int)bcnodes[i]; }
      c->cg_bc.upload( bc, s ); c->cg_bcnode.upload( nd, s ); c->cg_bcnodes_h = nodes;
      c->cg_bcval.alloc( n ); c->cg_bcsmall.alloc( std::max< size_t >( nbc, 1 ) );
      CK( cudaMemsetAsync( c->cg_bcval.p, 0, n*sizeof(double), s ) );
    }
    if (nbc) {
      std::vector< double > v( nbc, 0.0 );
      if (bcvals) std::copy( bcvals, bcvals + nbc, v.begin() );
      CK( cudaMemcpyAsync( c->cg_bcsmall.p, v.data(), nbc*sizeof(double), cudaMemcpyHostToDevice, s ) );
      CK( cudaStreamSynchronize( s ) );
      k_cg_bc_scatter<<< nblk( nbc, 256 ), 256, 0, s >>>( (int)nbc, c->cg_bcnode.p, c->cg_bcsmall.p, c->cg_bcval.p ); ++c->launches;
    } }
  c->cg_hasbc = true;
  const double* neu = nullptr;
  if (neubc) { c->cg_neu.upload( std::vector< double >( neubc, neubc+n ), s ); neu = c->cg_neu.p; }
  k_cg_bc_colsum<<< nblk( c->cg_nslice*32, 256 ), 256, 0, s >>>( n, c->cg_base.p, c->cg_col.p, c->cg_val.p, c->cg_bc.p,
    c->cg_bcval.p, c->cg_q.p ); ++c->launches;
  // several partitions: the Neumann vector and the column sums are this partition's parts
  // (ConjugateGradients::combc :407-428 sums qc, comr :508-525 sums rc at the shared rows)
  if (neu) cg_halo( c, c->cg_neu.p, 0 );
  cg_halo( c, c->cg_q.p, 0 );
  k_cg_bc_rhs<<< nblk( n, 256 ), 256, 0, s >>>( n, c->cg_bc.p, c->cg_bcval.p, neu, c->cg_q.p, c->cg_b.p ); ++c->launches;
  CK( cudaGetLastError() );
}
} // namespace

int xyst_chocg_mesh_upload( xyst_ctx* c, size_t npoin, const double* x, const double* y, const double* z,
                            const size_t nsup[3], const size_t* const dsupedge[3],
                            const double* const dsupint[3], size_t ntri, const size_t* triinpoel,
                            const double* vol, const double* v, const xyst_chocg_params* prm )
{
  if (!c || !prm) return fail( "null argument" );
  if (prm->flux != 0 && prm->flux != 1) return fail( "Flux not correctly configured" );
  c->cho = true;
  std::vector< uint8_t > besym( ntri*3, 0 );
  if (int r = mesh_upload_impl( c, npoin, x, y, z, nsup, dsupedge, dsupint, ntri, triinpoel, besym.data(), vol, v, 5 )) { c->cho = false; return r; }
  API_BEGIN
  c->chp = *prm;
  size_t NP = c->NP;
  for (auto* b : { &c->cUa, &c->cUb, &c->cUc, &c->cSg, &c->cPg, &c->cFl, &c->cR }) {
    b->alloc( 3*NP ); CK( cudaMemsetAsync( b->p, 0, 3*NP*sizeof(double), c->stream ) ); }
  for (auto* b : { &c->cP, &c->cDiv }) { b->alloc( NP ); CK( cudaMemsetAsync( b->p, 0, NP*sizeof(double), c->stream ) ); }
  c->cVg.alloc( 9*NP ); CK( cudaMemsetAsync( c->cVg.p, 0, 9*NP*sizeof(double), c->stream ) );
  c->cS.release();
  c->cU = c->cUa.p; c->cUn = c->cUb.p; c->cUx = c->cUc.p;
  c->cns = 0; c->cdif = 0.0; c->ncpin = 0;
  c->cb_nd = c->cb_ns = c->cb_nn = 0;
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

int xyst_chocg_bc_upload( xyst_ctx* c, size_t ndir, const size_t* dirnodes, const int* dirmask,
                          const double* dirval, size_t nsym, const size_t* symbcnodes,
                          const double* symbcnorms, size_t nnoslip, const size_t* noslipbcnodes )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  auto s = c->stream;
  auto chk = [&]( size_t id ){ if (id >= c->npoin) throw std::runtime_error( "BC node id out of range" ); return (int)id; };
  { size_t m = (size_t)cho_rows( c );
    std::vector< int > nd( ndir ), mk( ndir*m ); std::vector< double > vl( ndir*m );
    for (size_t i=0; i<ndir; ++i) { nd[i] = chk( dirnodes[i] );
      for (size_t k=0; k<m; ++k) { mk[i*m+k] = dirmask[i*m+k]; vl[i*m+k] = dirval ? dirval[i*m+k] : 0.0;
        if (mk[i*m+k] == 2 && !dirval) mk[i*m+k] = 0; } }             // BC.cpp:66: mask 2 needs a value list
    c->cb_dnode.upload( nd, s ); c->cb_dmask.upload( mk, s ); c->cb_dval.upload( vl, s ); c->cb_nd = ndir; }
  { // group the (node, normal) pairs by node, keeping the order in which a node's pairs arrive
    std::vector< int > first( c->npoin, -1 ), nodes, cnt;
    for (size_t i=0; i<nsym; ++i) { int p = chk( symbcnodes[i] );
      if (first[p] < 0) { first[p] = (int)nodes.size(); nodes.push_back( p ); cnt.push_back( 0 ); }
      ++cnt[ first[p] ]; }
    std::vector< int > off( nodes.size()+1, 0 );
    for (size_t i=0; i<nodes.size(); ++i) off[i+1] = off[i] + cnt[i];
    std::vector< int > fillp( off.begin(), off.end()-1 );
    std::vector< double > nr( nsym*3 );
    for (size_t i=0; i<nsym; ++i) { int k = fillp[ first[ symbcnodes[i] ] ]++;
      for (int j=0; j<3; ++j) nr[(size_t)k*3+j] = symbcnorms[i*3+j]; }
    c->cb_snode.upload( nodes, s ); c->cb_soff.upload( off, s ); c->cb_snorm.upload( nr, s ); c->cb_ns = nodes.size(); }
  { std::vector< int > nn( nnoslip );
    for (size_t i=0; i<nnoslip; ++i) nn[i] = chk( noslipbcnodes[i] );
    c->cb_nnode.upload( nn, s ); c->cb_nn = nnoslip; }
  API_END
}

int xyst_chocg_set_u( xyst_ctx* c, const double* u ) { API_BEGIN CK( cudaSetDevice( c->device ) ); cho_need( c ); cho_set( c, u, cho_rows( c ), c->cU ); API_END }
int xyst_chocg_get_u( xyst_ctx* c, double* u ) { API_BEGIN CK( cudaSetDevice( c->device ) ); cho_need( c ); cho_get( c, c->cU, cho_rows( c ), u ); API_END }

// Transported scalars next to the three velocity components (ChoCG::m_u with problem_ncomp > 3): call after
// the mesh upload and before any state / BC upload. The velocity-like buffers get ns more rows.
int xyst_chocg_scalars( xyst_ctx* c, int ns, double diffusivity )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  if (c->loh) throw std::runtime_error( "context holds a LohCG mesh: use xyst_lohcg_scalars" );
  if (ns < 0 || ns > CHO_NSMAX) throw std::runtime_error( "xyst_chocg_scalars: 0..4 transported scalars" );
  size_t NP = c->NP, m = 3 + (size_t)ns;
  for (auto* b : { &c->cUa, &c->cUb, &c->cUc, &c->cR }) { b->alloc( m*NP ); CK( cudaMemsetAsync( b->p, 0, m*NP*sizeof(double), c->stream ) ); }
  c->cVg.alloc( 3*m*NP ); CK( cudaMemsetAsync( c->cVg.p, 0, 3*m*NP*sizeof(double), c->stream ) );
  c->cS.release();
  c->cU = c->cUa.p; c->cUn = c->cUb.p; c->cUx = c->cUc.p;
  c->cns = ns; c->cdif = diffusivity; c->ncpin = 0;
  c->cb_nd = 0;
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

// new values for the Dirichlet nodes of the last xyst_chocg_bc_upload / xyst_lohcg_bc_upload
// (time-dependent physics::dirbc, BC.cpp:57-66): [ndir][rows]
int xyst_chocg_dirbc_values( xyst_ctx* c, const double* dirval )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  if (!dirval) throw std::runtime_error( "null argument" );
  if (c->loh) { if (c->lb_nd) CK( cudaMemcpyAsync( c->lb_dval.p, dirval, c->lb_nd*(size_t)(4+c->cns)*sizeof(double), cudaMemcpyHostToDevice, c->stream ) ); }
  else if (c->cb_nd) CK( cudaMemcpyAsync( c->cb_dval.p, dirval, c->cb_nd*(size_t)cho_rows( c )*sizeof(double), cudaMemcpyHostToDevice, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  API_END
}

// frozen flow (ChoCG::solve :1550-1552,1564-1570): the velocity rows of the time level n come back
// (the velocity of before the stage's update: with the flow frozen for the whole step that is un)
int xyst_chocg_restore_velocity( xyst_ctx* c )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  if (c->loh) throw std::runtime_error( "context holds a LohCG mesh" );
  CK( cudaMemcpyAsync( c->cU, c->cUn, 3*c->NP*sizeof(double), cudaMemcpyDeviceToDevice, c->stream ) );
  API_END
}

// problems::point_src: nodes whose (first) transported scalar is set to value after every stage's update
int xyst_chocg_pin( xyst_ctx* c, size_t n, const size_t* nodes, double value )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  if (!c->cns) throw std::runtime_error( "xyst_chocg_pin: no transported scalar in this context" );
  std::vector< int > h( n );
  for (size_t i=0; i<n; ++i) { if (nodes[i] >= c->npoin) throw std::runtime_error( "point-source node id out of range" ); h[i] = (int)nodes[i]; }
  c->cpin.upload( h, c->stream ); c->ncpin = n; c->cpin_val = value;
  API_END
}
int xyst_chocg_set_p( xyst_ctx* c, const double* p ) { API_BEGIN CK( cudaSetDevice( c->device ) ); cho_need( c ); cho_set( c, p, 1, c->cP.p ); API_END }

int xyst_chocg_get( xyst_ctx* c, const char* what, double* out )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  std::string w( what ? what : "" );
  if (w == "pr") cho_get( c, c->cP.p, 1, out );
  else if (w == "div") cho_get( c, c->cDiv.p, 1, out );
  else if (w == "sgrad") cho_get( c, c->cSg.p, 3, out );
  else if (w == "pgrad") cho_get( c, c->cPg.p, 3, out );
  else if (w == "flux") cho_get( c, c->cFl.p, 3, out );
  else if (w == "rhs") cho_get( c, c->cR.p, cho_rows( c ), out );
  else if (w == "vgrad") cho_get( c, c->cVg.p, 3*cho_rows( c ), out );
  else if (w == "un") cho_get( c, c->cUn, cho_rows( c ), out );
  else if (w == "u") cho_get( c, c->cU, cho_rows( c ), out );
  else if (w == "dp") { cho_need_cg( c ); CK( cudaMemcpyAsync( out, c->cg_x.p, c->npoin*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) ); CK( cudaStreamSynchronize( c->stream ) ); }
  else throw std::runtime_error( "xyst_chocg_get: unknown field " + w );
  API_END
}

int xyst_chocg_apply_bc( xyst_ctx* c ) { API_BEGIN CK( cudaSetDevice( c->device ) ); cho_need( c ); cho_bc( c, c->cU, true, true ); API_END }

int xyst_chocg_div( xyst_ctx* c, int which, double dt, int stab )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  auto s = c->stream;
  const double* V = c->cU;
  if (which == 1) {            // ChoCG::div :879-882: finish the momentum flux first
    k_cho_divvol<<< nblk( c->npoin, 256 ), 256, 0, s >>>( c->npoin, c->NP, 3, c->vol.p, c->cFl.p ); ++c->launches;
    cho_bc( c, c->cFl.p, false, false );
    V = c->cFl.p;
  } else if (which != 0) throw std::runtime_error( "xyst_chocg_div: which must be 0 (velocity) or 1 (momentum flux)" );
  ProfScope ps( c, "cho_div" );
  if (stab)
    k_cho_div< true ><<< cho_grid( c ), NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
      c->nslot, V, c->X.p, c->cP.p, c->cPg.p, dt, c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, c->cDiv.p );
  else
    k_cho_div< false ><<< cho_grid( c ), NODE_THREADS, 0, s >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
      c->nslot, V, c->X.p, c->cP.p, c->cPg.p, dt, c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, c->cDiv.p );
  ++c->launches;
  CK( cudaGetLastError() );
  soa_halo( c, c->cDiv.p, 1 );                      // comdiv :902-921
  API_END
}

int xyst_chocg_vgrad( xyst_ctx* c ) { API_BEGIN CK( cudaSetDevice( c->device ) ); cho_need( c ); cho_vgrad( c ); API_END }

int xyst_chocg_flux( xyst_ctx* c )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  ProfScope ps( c, "cho_flux" );
  k_cho_flux<<< cho_grid( c ), NODE_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p, c->D.p,
    c->nslot, c->cU, c->cVg.p, c->chp.mu, c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, c->cFl.p ); ++c->launches;
  CK( cudaGetLastError() );
  soa_halo( c, c->cFl.p, 3 );                       // comflux :1002-1023
  API_END
}

int xyst_chocg_grad( xyst_ctx* c, int which )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  const double* S; double* G;
  if (which == 0) { cho_need_cg( c ); S = c->cg_x.p; G = c->cSg.p; }
  else if (which == 1) { S = c->cP.p; G = c->cPg.p; }
  else throw std::runtime_error( "xyst_chocg_grad: which must be 0 (CG solution) or 1 (pressure)" );
  ProfScope ps( c, "cho_grad" );
  k_cho_grad< 1 ><<< cho_grid( c ), NODE_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->sl_base.p, c->inc_e.p, c->inc_q.p,
    c->D.p, c->nslot, S, c->bslot.p, c->bn_off.p, c->bn_face.p, c->tri.p, c->fn.p, c->vol.p, G ); ++c->launches;
  CK( cudaGetLastError() );
  soa_halo( c, G, 3 );                              // comsgrad :1171-1192, compgrad :1283-1304
  API_END
}

int xyst_chocg_src( xyst_ctx* c, const double* S )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  if (!S) { c->cS.release(); return 0; }
  size_t m = (size_t)cho_rows( c );
  if (c->cS.n != m*c->NP) { c->cS.alloc( m*c->NP ); CK( cudaMemsetAsync( c->cS.p, 0, m*c->NP*sizeof(double), c->stream ) ); }
  cho_set( c, S, (int)m, c->cS.p );
  API_END
}

int xyst_chocg_rhs( xyst_ctx* c )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  cho_rhs( c, nullptr, 0.0, nullptr, c->cR.p );
  API_END
}

int xyst_chocg_stage( xyst_ctx* c, int stage, double rkcoef_, double dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  if (stage < 0) throw std::runtime_error( "stage must be >= 0" );
  // un = u at stage 0 without a copy: the three velocity buffers rotate
  if (stage == 0) { double* old_un = c->cUn; c->cUn = c->cU; cho_rhs( c, c->cUn, rkcoef_*dt, c->cUx, nullptr ); c->cU = c->cUx; c->cUx = old_un; }
  else { cho_rhs( c, c->cUn, rkcoef_*dt, c->cUx, nullptr ); std::swap( c->cU, c->cUx ); }
  cho_pin( c );
  cho_bc( c, c->cU, true, true );
  if (c->chp.flux == 1) cho_vgrad( c );            // ChoCG::corr :1677
  API_END
}

int xyst_chocg_pinit( xyst_ctx* c, double divisor, size_t nbc, const size_t* bcnodes,
                      const double* bcvals, const double* neubc, const double* rhs0, int pc )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need_cg( c );
  auto s = c->stream;
  size_t n = c->npoin;
  if (rhs0) CK( cudaMemcpyAsync( c->cg_b.p, rhs0, n*sizeof(double), cudaMemcpyHostToDevice, s ) );
  else { k_cho_prhs<<< nblk( n, 256 ), 256, 0, s >>>( n, divisor, c->cDiv.p, c->cg_b.p ); ++c->launches; }
  cho_cg_bc( c, nbc, bcnodes, bcvals, neubc );
  cg_setup_dev( c, pc );
  API_END
}

// semi-implicit momentum solve (theta > 0) at the last RK stage, ChoCG::solve :1574-1607: with the
// momentum matrix selected (xyst_cg_select 1; block CSR with 3 scalar rows per node, ChoCG::lhs
// :1433-1477), b = rhs of the last xyst_chocg_rhs, Dirichlet rows (node*3+component) with value 0,
// initial guess = previous solution. Continue with xyst_cg_solve and xyst_chocg_mupdate.
int xyst_chocg_minit( xyst_ctx* c, size_t nbc, const size_t* bcrows, int pc )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  if (c->loh) throw std::runtime_error( "context holds a LohCG mesh" );
  if (c->cns) throw std::runtime_error( "ChoCG: the semi-implicit momentum solve with transported scalars is not implemented" );
  if (c->cg_nrow != c->npoin*3 || c->cg_ncomp != 3)
    throw std::runtime_error( "ChoCG: select the momentum matrix (3 scalar rows per node) with xyst_cg_select / xyst_csr_upload first" );
  k_cho_mrhs<<< nblk( c->npoin*3, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, c->cR.p, c->cg_b.p ); ++c->launches;
  cho_cg_bc( c, nbc, bcrows, nullptr, nullptr );
  cg_setup_dev( c, pc );
  API_END
}

// u = un + du, BC, and for the damp4 flux the velocity gradient (ChoCG::msolved :1625-1645 + pred);
// stage = index of this (the last) RK stage: at stage 0 un is the current velocity
int xyst_chocg_mupdate( xyst_ctx* c, int stage )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  if (c->cg_nrow != c->npoin*3) throw std::runtime_error( "ChoCG: the momentum solver is not selected" );
  if (stage < 0) throw std::runtime_error( "stage must be >= 0" );
  unsigned g = nblk( c->npoin*3, 256 );
  if (stage == 0) { double* old_un = c->cUn; c->cUn = c->cU;
    k_cho_mupdate<<< g, 256, 0, c->stream >>>( c->npoin, c->NP, c->cUn, c->cg_x.p, c->cUx ); c->cU = c->cUx; c->cUx = old_un; }
  else { k_cho_mupdate<<< g, 256, 0, c->stream >>>( c->npoin, c->NP, c->cUn, c->cg_x.p, c->cUx ); std::swap( c->cU, c->cUx ); }
  ++c->launches;
  cho_bc( c, c->cU, true, true );
  if (c->chp.flux == 1) cho_vgrad( c );
  API_END
}

int xyst_chocg_project( xyst_ctx* c, double pdt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  k_cho_project<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, c->NP, pdt, c->cSg.p, c->cU ); ++c->launches;
  cho_bc( c, c->cU, true, true );
  API_END
}

int xyst_chocg_pressure_update( xyst_ctx* c, int increment )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need_cg( c );
  k_cho_pupdate<<< nblk( c->npoin, 256 ), 256, 0, c->stream >>>( c->npoin, increment, c->cg_x.p, c->cP.p ); ++c->launches;
  CK( cudaGetLastError() );
  API_END
}

int xyst_chocg_dt_min( xyst_ctx* c, double cfl, double dif, double* dt )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need( c );
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( c->npoin, RED_THREADS ) );
  double* fin = c->red.p + (size_t)RED_BLOCKS*NDIAG;
  k_cho_dt<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->cU, c->vol.p, std::max( c->chp.mu, dif ), c->red.p );
  k_reduce_final< 1, true ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, fin );
  c->launches += 2;
  CK( cudaMemcpyAsync( c->red_host, fin, sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  *dt = c->red_host[0] * cfl;
  API_END
}

int xyst_chocg_diag( xyst_ctx* c, const double* an_p, const double* an_u, double* out )
{
  API_BEGIN
  CK( cudaSetDevice( c->device ) );
  cho_need_cg( c );
  static_assert( NCHODIAG <= NDIAG, "reduction scratch too small" );
  DevBuf< double > dp, du;
  if (an_p) dp.upload( std::vector< double >( an_p, an_p + c->npoin ), c->stream );
  if (an_u) du.upload( std::vector< double >( an_u, an_u + c->npoin*(size_t)cho_rows( c ) ), c->stream );
  int nb = (int)std::min< size_t >( RED_BLOCKS, nblk( c->npoin, RED_THREADS ) );
  double* fin = c->red.p + (size_t)RED_BLOCKS*NDIAG;
  k_cho_diag<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->cU, c->cUn, c->cP.p, c->cg_x.p, c->v.p, dp.p, du.p, cho_rows( c ), c->red.p );
  k_reduce_final< NCHODIAG, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, fin );
  c->launches += 2;
  CK( cudaMemcpyAsync( c->red_host, fin, NCHODIAG*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
  CK( cudaStreamSynchronize( c->stream ) );
  for (int i=0; i<NCHODIAG; ++i) out[i] = c->red_host[i];
  if (c->cns) {            // scalar rows: out[16+4k..] = L2 solution, L2 increment, L2 error, L1 error sums of scalar k
    k_cho_sdiag<<< nb, RED_THREADS, 0, c->stream >>>( c->npoin, c->NP, c->cns, cho_rows( c ), c->cU, c->cUn, c->v.p, du.p, c->red.p );
    k_reduce_final< 4*CHO_NSMAX, false ><<< 1, RED_THREADS, 0, c->stream >>>( nb, c->red.p, fin );
    c->launches += 2;
    CK( cudaMemcpyAsync( c->red_host, fin, 4*CHO_NSMAX*sizeof(double), cudaMemcpyDeviceToHost, c->stream ) );
    CK( cudaStreamSynchronize( c->stream ) );
    for (int i=0; i<4*c->cns; ++i) out[NCHODIAG+i] = c->red_host[i];
  }
  API_END
}

#endif // XYST_CHOCG_API

// xyst_b200/csrc/riecg_kernels.cuh -- RieCG / LaxCG device code: nodal state, boundary-face terms, gradient gather, MUSCL + Riemann edge fluxes, flux gather with the fused RK update, BCs, dt and diagnostics reductions
// Part of the single translation unit xyst_b200.cu (included inside its anonymous namespace).

// ---------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------
struct DParams { double gamma, stab2coef; int flux, stab2, exact; double rgas, kvinf; };
// what the node kernels need to convert between variable sets; rgas > 0 selects LaxCG
struct Mode { double gamma, rgas, kvinf; };

// Primitive variables from conserved ones, Riemann.cpp:211-227
// (one reciprocal of the density instead of the reference's four quotients: the same values up to
// one rounding each)
__device__ __forceinline__ void primitive( const double u[NC], double w[NC] ) {
  w[0] = u[0];
  double ir = 1.0 / w[0];
  w[1] = u[1] * ir;
  w[2] = u[2] * ir;
  w[3] = u[3] * ir;
  w[4] = u[4] * ir - 0.5*(w[1]*w[1] + w[2]*w[2] + w[3]*w[3]);
}

// LaxCG::primitive, LaxCG.cpp:115-137: (r,ru,rv,rw,rE) -> (p,u,v,w,T)
__device__ __forceinline__ void lax_primitive( const double u[NC], double w[NC], double gamma, double rgas ) {
  double r = u[0];
  double uu = u[1]/r, vv = u[2]/r, ww = u[3]/r;
  double p = (u[4] - 0.5*r*(uu*uu + vv*vv + ww*ww)) * (gamma-1.0);
  w[0] = p; w[1] = uu; w[2] = vv; w[3] = ww; w[4] = p/r/rgas;
}
// LaxCG::conservative, LaxCG.cpp:139-164
__device__ __forceinline__ void lax_conservative( const double w[NC], double u[NC], double gamma, double rgas ) {
  double p = w[0], uu = w[1], vv = w[2], ww = w[3], T = w[4];
  double r = p/T/rgas;
  u[0] = r; u[1] = r*uu; u[2] = r*vv; u[3] = r*ww;
  u[4] = p/(gamma-1.0) + 0.5*r*(uu*uu + vv*vv + ww*ww);
}
__device__ __forceinline__ void primitive_of( const double u[NC], double w[NC], const Mode& M ) {
  if (M.rgas > 0.0) lax_primitive( u, w, M.gamma, M.rgas ); else primitive( u, w );
}
// lax::refvel, Lax.cpp:344-360
__device__ __forceinline__ double lax_refvel( double r, double p, double v, double gamma, double kvinf ) {
  return fmin( sqrt( gamma * p / r ), fmax( v, kvinf ) );
}

__device__ __forceinline__ void cp_async_commit() { asm volatile( "cp.async.commit_group;" ::: "memory" ); }
// flat index of tk::Fields G(p,i), i = c*3+j, in the structure-of-arrays gradient storage
// Gradients are stored as eight 16-byte pairs per node, pair k of node p at G2[k*NP+p]:
// (g0,g1) ... (g12,g13) (g14,-) -- an edge's other end is gathered with 8 LDG.128.
constexpr int NGP = 8;
__host__ __device__ __forceinline__ size_t gidx( int i, size_t p, size_t NP ) { return ((size_t)(i>>1)*NP + p)*2 + (size_t)(i&1); }
__device__ __forceinline__ void store_g( double* __restrict__ G, size_t NP, size_t p, const double g[15] ) {
  double2* G2 = reinterpret_cast< double2* >( G );
  #pragma unroll
  for (int k=0; k<7; ++k) G2[(size_t)k*NP+p] = make_double2( g[2*k], g[2*k+1] );
  G[gidx( 14, p, NP )] = g[14];
}
template< int N > __device__ __forceinline__ void cp_async_wait() { asm volatile( "cp.async.wait_group %0;" :: "n"( N ) : "memory" ); }

// Primitive variables and coordinates of a node as four 16-byte pairs, pair k of node p at
// WX[k*NP+p]: (w0,w1) (w2,w3) (w4,x) (y,z) -- what both edge sweeps gather from an edge's other
// end: 3 loads (gradient) or 4 (flux) instead of 5 or 8, still coalesced over consecutive nodes.
// x,y,z are written once at upload; writers of W leave them alone.
__device__ __forceinline__ void load_w( const double2* __restrict__ WX, size_t NP, size_t p, double w[NC] ) {
  double2 a = __ldg( WX + p ), b = __ldg( WX + NP + p ), c = __ldg( WX + 2*NP + p );
  w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x;
}
__device__ __forceinline__ void load_wx( const double2* __restrict__ WX, size_t NP, size_t p, double w[NC], double x[3] ) {
  double2 a = __ldg( WX + p ), b = __ldg( WX + NP + p ), c = __ldg( WX + 2*NP + p ), d = __ldg( WX + 3*NP + p );
  w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; x[0] = c.y; x[1] = d.x; x[2] = d.y;
}
__device__ __forceinline__ void store_w( double* __restrict__ W, size_t NP, size_t p, const double w[NC] ) {
  double2* WX = reinterpret_cast< double2* >( W );
  WX[p] = make_double2( w[0], w[1] );
  WX[NP+p] = make_double2( w[2], w[3] );
  W[(2*NP+p)*2] = w[4];
}
__device__ __forceinline__ double get_w( const double* __restrict__ W, size_t NP, int c, size_t p ) {
  return W[((size_t)(c>>1)*NP + p)*2 + (size_t)(c&1)];
}
// Edge fluxes as (f0,f1) (f2,f3) pairs and f4 per slot
__device__ __forceinline__ void store_f( double* __restrict__ F, size_t nslot, size_t e, const double f[NC] ) {
  double2* F2 = reinterpret_cast< double2* >( F );
  F2[e] = make_double2( f[0], f[1] );
  F2[nslot+e] = make_double2( f[2], f[3] );
  F[4*nslot+e] = f[4];
}
__device__ __forceinline__ void load_f( const double* __restrict__ F, size_t nslot, size_t e, double f[NC] ) {
  const double2* F2 = reinterpret_cast< const double2* >( F );
  double2 a = __ldg( F2 + e ), b = __ldg( F2 + nslot + e );
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = __ldg( F + 4*nslot + e );
}

// reference layout [node][comp] -> SoA state + primitives
// (n2o: internal node id -> the caller's, or null if the two orders are the same)
// (ncomp = 5 + ns: the ns transported scalars of a node follow its flow variables in A and go to sU)
__global__ void k_set_state( size_t n, size_t NP, const double* __restrict__ A, const int* __restrict__ n2o,
                             double* __restrict__ U, double* __restrict__ W, Mode M, int ncomp, double* __restrict__ sU )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  double u[NC], w[NC];
  size_t o = n2o ? (size_t)n2o[p] : p;
  #pragma unroll
  for (int c=0; c<NC; ++c) u[c] = A[o*ncomp+c];
  for (int c=NC; c<ncomp; ++c) sU[(size_t)(c-NC)*NP+p] = A[o*ncomp+c];
  primitive_of( u, w, M );
  #pragma unroll
  for (int c=0; c<NC; ++c) U[c*NP+p] = u[c];
  store_w( W, NP, p, w );
}

__global__ void k_get_state( size_t n, size_t NP, const double* __restrict__ U, const int* __restrict__ n2o,
                             double* __restrict__ A, int ncomp, const double* __restrict__ sU )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  size_t o = n2o ? (size_t)n2o[p] : p;
  #pragma unroll
  for (int c=0; c<NC; ++c) A[o*ncomp+c] = U[c*NP+p];
  for (int c=NC; c<ncomp; ++c) A[o*ncomp+c] = sU[(size_t)(c-NC)*NP+p];
}

// ---------------------------------------------------------------------------------
// boundary-face contributions, gathered per boundary node
//   gradient part: Riemann.cpp:334-360 (incl. the direction-indexed g[j]*n[j] form)
//   flux part    : Riemann.cpp:798-871
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void face_normal( const double* __restrict__ X, size_t NP, const int N[3], double n[3] ) {
  double a[3], b[3], c[3];
  #pragma unroll
  for (int j=0; j<3; ++j) { a[j] = X[j*NP+N[0]]; b[j] = X[j*NP+N[1]]; c[j] = X[j*NP+N[2]]; }
  double ba[3] = { b[0]-a[0], b[1]-a[1], b[2]-a[2] }, ca[3] = { c[0]-a[0], c[1]-a[1], c[2]-a[2] };
  n[0] = (ba[1]*ca[2] - ca[1]*ba[2]) / 12.0;
  n[1] = (ba[2]*ca[0] - ca[2]*ba[0]) / 12.0;
  n[2] = (ba[0]*ca[1] - ca[0]*ba[1]) / 12.0;
}

// the normals depend on coordinates only: computed once at upload with this arithmetic
__global__ void k_face_normals( int ntri, size_t NP, const int* __restrict__ tri, const double* __restrict__ X,
                                double* __restrict__ fn )
{
  int f = blockIdx.x*blockDim.x + threadIdx.x;
  if (f >= ntri) return;
  int N[3] = { tri[f*3+0], tri[f*3+1], tri[f*3+2] };
  double n[3];
  face_normal( X, NP, N, n );
  fn[(size_t)f*3+0] = n[0]; fn[(size_t)f*3+1] = n[1]; fn[(size_t)f*3+2] = n[2];
}

__global__ void k_bnd_grad( int nbn, size_t NP, const int* __restrict__ bn_off, const int* __restrict__ bn_face,
                            const int* __restrict__ tri, const double* __restrict__ fn,
                            const double* __restrict__ W, double* __restrict__ Gb )
{
  int b = blockIdx.x*blockDim.x + threadIdx.x;
  if (b >= nbn) return;
  double acc[15];
  #pragma unroll
  for (int i=0; i<15; ++i) acc[i] = 0.0;
  for (int i=bn_off[b]; i<bn_off[b+1]; ++i) {
    int f = bn_face[i] >> 2;
    int N[3] = { tri[f*3+0], tri[f*3+1], tri[f*3+2] };
    double n[3] = { fn[(size_t)f*3+0], fn[(size_t)f*3+1], fn[(size_t)f*3+2] };
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double u0 = get_w( W, NP, c, N[0] ), u1 = get_w( W, NP, c, N[1] ), u2 = get_w( W, NP, c, N[2] );
      double uab = (u0 + u1)/4.0;
      double ubc = (u1 + u2)/4.0;
      double uca = (u2 + u0)/4.0;
      double g[3] = { uab + uca + u0, uab + ubc + u1, ubc + uca + u2 };
      #pragma unroll
      for (int j=0; j<3; ++j) acc[c*3+j] += g[j] * n[j];
    }
  }
  #pragma unroll
  for (int i=0; i<15; ++i) Gb[(size_t)b*15+i] = acc[i];
}

__global__ void k_bnd_rhs( int nbn, size_t NP, const int* __restrict__ bn_off, const int* __restrict__ bn_face,
                           const int* __restrict__ tri, const unsigned char* __restrict__ besym,
                           const double* __restrict__ fn, const double* __restrict__ U,
                           double* __restrict__ Rb, double gamma, const double* __restrict__ W, double rgas )
{
  int b = blockIdx.x*blockDim.x + threadIdx.x;
  if (b >= nbn) return;
  if (rgas > 0.0) {                       // lax::advbnd, Lax.cpp:840-950, on (p,u,v,w,T)
    double acc[NC] = { 0, 0, 0, 0, 0 };
    for (int i=bn_off[b]; i<bn_off[b+1]; ++i) {
      int f = bn_face[i] >> 2, k = bn_face[i] & 3;
      int N[3] = { tri[f*3+0], tri[f*3+1], tri[f*3+2] };
      double n[3] = { fn[(size_t)f*3+0], fn[(size_t)f*3+1], fn[(size_t)f*3+2] };
      double fl[NC][3];
      #pragma unroll
      for (int m=0; m<3; ++m) {
        double pr = get_w( W, NP, 0, N[m] ), uu = get_w( W, NP, 1, N[m] ), vv = get_w( W, NP, 2, N[m] ), ww = get_w( W, NP, 3, N[m] ), T = get_w( W, NP, 4, N[m] );
        double rA = pr/T/rgas;
        double ruA = uu * rA, rvA = vv * rA, rwA = ww * rA;
        double reA = pr/(gamma-1.0) + 0.5*(ruA*ruA + rvA*rvA + rwA*rwA)/rA;
        double vn = besym[f*3+m] ? 0.0 : (n[0]*uu + n[1]*vv + n[2]*ww);
        fl[0][m] = rA*vn;
        fl[1][m] = ruA*vn + pr*n[0];
        fl[2][m] = rvA*vn + pr*n[1];
        fl[3][m] = rwA*vn + pr*n[2];
        fl[4][m] = (reA + pr)*vn;
      }
      #pragma unroll
      for (int c=0; c<NC; ++c) {
        double fab = (fl[c][0] + fl[c][1])/4.0;
        double fbc = (fl[c][1] + fl[c][2])/4.0;
        double fca = (fl[c][2] + fl[c][0])/4.0;
        double add = k == 0 ? fab + fca + fl[c][0] : (k == 1 ? fab + fbc + fl[c][1] : fbc + fca + fl[c][2]);
        acc[c] += add;
      }
    }
    #pragma unroll
    for (int c=0; c<NC; ++c) Rb[(size_t)b*NC+c] = acc[c];
    return;
  }
  double acc[NC] = { 0, 0, 0, 0, 0 };
  for (int i=bn_off[b]; i<bn_off[b+1]; ++i) {
    int f = bn_face[i] >> 2, k = bn_face[i] & 3;
    int N[3] = { tri[f*3+0], tri[f*3+1], tri[f*3+2] };
    double n[3] = { fn[(size_t)f*3+0], fn[(size_t)f*3+1], fn[(size_t)f*3+2] };
    double fl[NC][3];
    #pragma unroll
    for (int m=0; m<3; ++m) {
      double r = U[N[m]], ru = U[NP+N[m]], rv = U[2*NP+N[m]], rw = U[3*NP+N[m]], re = U[4*NP+N[m]];
      double p = (re - 0.5*(ru*ru + rv*rv + rw*rw)/r) * (gamma-1.0);
      double vn = besym[f*3+m] ? 0.0 : (n[0]*ru + n[1]*rv + n[2]*rw)/r;
      fl[0][m] = r*vn;
      fl[1][m] = ru*vn + p*n[0];
      fl[2][m] = rv*vn + p*n[1];
      fl[3][m] = rw*vn + p*n[2];
      fl[4][m] = (re + p)*vn;
    }
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double fab = (fl[c][0] + fl[c][1])/4.0;
      double fbc = (fl[c][1] + fl[c][2])/4.0;
      double fca = (fl[c][2] + fl[c][0])/4.0;
      double add = k == 0 ? fab + fca + fl[c][0] : (k == 1 ? fab + fbc + fl[c][1] : fbc + fca + fl[c][2]);
      acc[c] += add;
    }
  }
  #pragma unroll
  for (int c=0; c<NC; ++c) Rb[(size_t)b*NC+c] = acc[c];
}

// ---------------------------------------------------------------------------------
// gradient gather: one warp per 32-node slice, one thread per node
//   G(p) = [ sum_edges -/+ d*(w_q + w_p)  +  boundary part ] / vol(p)
// entry = signed edge slot: +(slot+1) if p is the edge's second node (receives +f),
// -(slot+1) if it is the first (receives -f), 0 = padding (multiplier 0 on slot 0)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void grad_sum( size_t p, int lane, long long base, int kmax,
    const int2* __restrict__ inc_eq, const double2* __restrict__ D2, const double* __restrict__ D,
    size_t nslot, const double* __restrict__ W, size_t NP, double acc[15] )
{
  const double2* WX = reinterpret_cast< const double2* >( W );
  double wp[NC];
  load_w( WX, NP, p, wp );
  #pragma unroll
  for (int i=0; i<15; ++i) acc[i] = 0.0;
  #pragma unroll kGradUnroll
  for (int k=0; k<kmax; ++k) {
    long long i = base + (long long)k*32 + lane;
    int2 eq = __ldg( inc_eq + i );
    int se = eq.x, q = eq.y;
    double sg = se > 0 ? 1.0 : (se < 0 ? -1.0 : 0.0);
    size_t sl = se == 0 ? 0 : (size_t)(abs(se)-1);
    double2 d01 = __ldg( D2 + sl );
    double d0 = sg * d01.x, d1 = sg * d01.y, d2 = sg * __ldg( D + 2*nslot + sl );
    if (se == 0) d0 = d1 = d2 = 0.0;      // padding: the neighbour is the node itself, the normal must not be slot 0's
    double wq[NC];
    load_w( WX, NP, (size_t)q, wq );
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double s = wq[c] + wp[c];
      acc[c*3+0] += d0 * s;
      acc[c*3+1] += d1 * s;
      acc[c*3+2] += d2 * s;
    }
  }
}

__global__ void __launch_bounds__(GRAD_THREADS, GRAD_MINB)
k_grad_node( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq,
             const double2* __restrict__ D2, const double* __restrict__ D, size_t nslot,
             const double* __restrict__ W, const int* __restrict__ bslot, const double* __restrict__ Gb,
             const double* __restrict__ vol, double* __restrict__ G, int defer_bnd )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[15];
  grad_sum( p, lane, base, kmax, inc_eq, D2, D, nslot, W, NP, acc );
  int b = bslot[p];
  if (b >= 0) {
    if (defer_bnd) {          // boundary part and division follow in k_grad_bfix (same operation order)
      store_g( G, NP, p, acc );
      return;
    }
    #pragma unroll
    for (int i=0; i<15; ++i) acc[i] += Gb[(size_t)b*15+i];
  }
  // one quotient per node, then 15 products (the reference divides each component, RieCG.cpp:936-939:
  // the same values up to one rounding)
  double ivp = 1.0 / vol[p];
  #pragma unroll
  for (int i=0; i<15; ++i) acc[i] *= ivp;
  store_g( G, NP, p, acc );
}

// The same gather as a persistent kernel: every warp walks over slices s, s+S, s+2S, ... and fetches
// the NEXT slice's incidence entries into its own piece of shared memory (cp.async, 8 bytes per entry
// and lane) while it works on the current one -- the dependent chain incidence entry -> normal and
// neighbour state then costs one trip to memory per slice instead of two.
// Shared memory: [2][kcap][blockDim.x] int2.
__global__ void __launch_bounds__(GRAD_THREADS, GRAD_MINB)
k_grad_node_p( size_t npoin, size_t NP, size_t nslice, int kcap, const long long* __restrict__ sl_base,
               const int2* __restrict__ inc_eq, const double2* __restrict__ D2, const double* __restrict__ D,
               size_t nslot, const double* __restrict__ W, const int* __restrict__ bslot,
               const double* __restrict__ Gb, const double* __restrict__ vol, double* __restrict__ G, int defer_bnd )
{
  extern __shared__ int2 sinc[];
  const int lane = threadIdx.x & 31, nt = blockDim.x;
  const size_t wstride = (size_t)gridDim.x*(nt >> 5);
  size_t slice = (size_t)blockIdx.x*(nt >> 5) + (threadIdx.x >> 5);
  const double2* WX = reinterpret_cast< const double2* >( W );
  auto fetch = [&]( int buf, size_t s, long long& base, int& kmax ) {
    base = sl_base[s];
    kmax = (int)((sl_base[s+1] - base) >> 5);
    int2* dst = sinc + (size_t)buf*kcap*nt + threadIdx.x;
    for (int k=0; k<kmax; ++k) {
      unsigned d = (unsigned)__cvta_generic_to_shared( dst + (size_t)k*nt );
      asm volatile( "cp.async.ca.shared.global [%0], [%1], 8;" :: "r"( d ), "l"( inc_eq + base + (long long)k*32 + lane ) : "memory" );
    }
    cp_async_commit();
  };
  if (slice >= nslice) return;
  long long base, base_nx = 0; int kmax, kmax_nx = 0, buf = 0;
  fetch( 0, slice, base, kmax );
  for (; slice < nslice; slice += wstride, buf ^= 1) {
    const size_t nxt = slice + wstride;
    if (nxt < nslice) { fetch( buf ^ 1, nxt, base_nx, kmax_nx ); cp_async_wait< 1 >(); }
    else cp_async_wait< 0 >();
    const size_t p = slice*32 + lane;
    const int2* src = sinc + (size_t)buf*kcap*nt + threadIdx.x;
    double wp[NC], acc[15];
    load_w( WX, NP, p < npoin ? p : npoin-1, wp );
    #pragma unroll
    for (int i=0; i<15; ++i) acc[i] = 0.0;
    #pragma unroll kGradUnroll
    for (int k=0; k<kmax; ++k) {
      int2 eq = src[(size_t)k*nt];
      int se = eq.x, q = eq.y;
      double sg = se > 0 ? 1.0 : (se < 0 ? -1.0 : 0.0);
      size_t sl = se == 0 ? 0 : (size_t)(abs(se)-1);
      double2 d01 = __ldg( D2 + sl );
      double d0 = sg * d01.x, d1 = sg * d01.y, d2 = sg * __ldg( D + 2*nslot + sl );
      double wq[NC];
      load_w( WX, NP, (size_t)q, wq );
      #pragma unroll
      for (int c=0; c<NC; ++c) {
        double s = wq[c] + wp[c];
        acc[c*3+0] += d0 * s;
        acc[c*3+1] += d1 * s;
        acc[c*3+2] += d2 * s;
      }
    }
    if (p < npoin) {
      int b = bslot[p];
      if (b >= 0 && defer_bnd) store_g( G, NP, p, acc );     // finished by k_grad_bfix
      else {
        if (b >= 0) {
          #pragma unroll
          for (int i=0; i<15; ++i) acc[i] += Gb[(size_t)b*15+i];
        }
        double ivp = 1.0 / vol[p];
        #pragma unroll
        for (int i=0; i<15; ++i) acc[i] *= ivp;
        store_g( G, NP, p, acc );
      }
    }
    base = base_nx; kmax = kmax_nx;
  }
}

// boundary nodes: G = (domain sum + boundary sum) / vol, once both are known
__global__ void k_grad_bfix( int nbn, size_t NP, const int* __restrict__ bn_node, const double* __restrict__ Gb,
                             const double* __restrict__ vol, double* __restrict__ G )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nbn*15) return;
  size_t b = i / 15, k = i % 15;
  size_t p = bn_node[b];
  size_t g = gidx( (int)k, p, NP );
  G[g] = (G[g] + Gb[b*15+k]) * (1.0 / vol[p]);      // as k_grad_node: quotient first, then product
}

// partial (un-normalised) gradient sums of the shared nodes, for the halo exchange
__global__ void k_grad_shared( int nsh, size_t NP, const int* __restrict__ sh_node,
             const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq,
             const double2* __restrict__ D2, const double* __restrict__ D, size_t nslot,
             const double* __restrict__ W, const int* __restrict__ bslot,
             const double* __restrict__ Gb, double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  size_t slice = p >> 5; int lane = p & 31;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[15];
  grad_sum( p, lane, base, kmax, inc_eq, D2, D, nslot, W, NP, acc );
  int b = bslot[p];
  if (b >= 0) for (int j=0; j<15; ++j) acc[j] += Gb[(size_t)b*15+j];
  for (int j=0; j<15; ++j) part[(size_t)i*15+j] = acc[j];
}

__global__ void k_pack( int nsend, int w, const int* __restrict__ sh_send,
                        const double* __restrict__ part, double* __restrict__ sendbuf )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nsend*w) return;
  size_t s = i / w, c = i % w;
  sendbuf[i] = part[(size_t)sh_send[s]*w + c];
}

__global__ void k_grad_finish( int nsh, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
             const int* __restrict__ ridx, const double* __restrict__ part,
             const double* __restrict__ recvbuf, const double* __restrict__ vol, double* __restrict__ G )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  double ivp = 1.0 / vol[p];
  for (int j=0; j<15; ++j) {
    double a = part[(size_t)i*15+j];
    for (int r=roff[i]; r<roff[i+1]; ++r) a += recvbuf[(size_t)ridx[r]*15+j];
    G[gidx( j, p, NP )] = a * ivp;
  }
}

// ---------------------------------------------------------------------------------
// edge fluxes: MUSCL (Riemann.cpp:34-143) + Rusanov (:369-478) or HLLC (:480-650)
// ---------------------------------------------------------------------------------
#define MUSCL_EPS 1.0e-9
#define MUSCL_K (1.0/3.0)
#ifndef MUSCL_SIGN_INT
#define MUSCL_SIGN_INT 0
#endif
#ifndef MUSCL_V2
#define MUSCL_V2 1        // 0: the first form of the two-reciprocal limiter (A/B timing)
#endif

// van Leer limited extrapolation increments for one component.
// exact: the reference's expression tree (8 divisions). fast: with a = d2+eps, b = d1+eps
// the two limiter values are phi(a/b) = 2a/(a+b) and phi(b/a) = 2b/(a+b) when a and b
// have the same sign and 0 otherwise, i.e. one reciprocal per side.
// 1/x to ~1 ulp without the IEEE division's slow path: hardware seed (MUFU.RCP64H, ~20
// bits) + two Newton steps. Only used by the non-"exact" limiter form; x is a sum of two
// same-signed numbers of magnitude >= 1e-9 whenever the result is used.
__device__ __forceinline__ double fast_rcp( double x )
{
  double y;
  asm( "rcp.approx.ftz.f64 %0, %1;" : "=d"( y ) : "d"( x ) );
#if MUSCL_V2
  // one cubically convergent step: y (1 + e + e^2), e = 1 - x y  (seed error 2^-20 -> 2^-60)
  double e = fma( -x, y, 1.0 );
  double t = fma( e, e, e );
  return fma( y, t, y );
#else
  double e = fma( -x, y, 1.0 );
  y = fma( y, e, y );
  e = fma( -x, y, 1.0 );
  y = fma( y, e, y );
  return y;
#endif
}

template< bool EXACT >
__device__ __forceinline__ void vanleer( double d1, double d2, double d3, double& incL, double& incR,
                                         const double MUSCL_EPS_ = MUSCL_EPS )
{
  if (EXACT) {
    double rcL = (d2 + MUSCL_EPS_) / (d1 + MUSCL_EPS_);
    double rcR = (d2 + MUSCL_EPS_) / (d3 + MUSCL_EPS_);
    double rLinv = (d1 + MUSCL_EPS_) / (d2 + MUSCL_EPS_);
    double rRinv = (d3 + MUSCL_EPS_) / (d2 + MUSCL_EPS_);
    double phiL = (fabs(rcL) + rcL) / (fabs(rcL) + 1.0);
    double phiR = (fabs(rcR) + rcR) / (fabs(rcR) + 1.0);
    double phi_L_inv = (fabs(rLinv) + rLinv) / (fabs(rLinv) + 1.0);
    double phi_R_inv = (fabs(rRinv) + rRinv) / (fabs(rRinv) + 1.0);
    incL = 0.25*(d1*(1.0-MUSCL_K)*phiL + d2*(1.0+MUSCL_K)*phi_L_inv);
    incR = 0.25*(d3*(1.0-MUSCL_K)*phiR + d2*(1.0+MUSCL_K)*phi_R_inv);
  } else {
#if MUSCL_V2
    // phi(a/b) = 2a/(a+b), phi(b/a) = 2b/(a+b) for same-signed a, b (else 0), so that
    // inc = 0.25 [ d1 (1-k) 2a + d2 (1+k) 2b ] / (a+b): one reciprocal, five multiply-adds per side
    const double c1 = 0.5*(1.0-MUSCL_K), c2 = 0.5*(1.0+MUSCL_K);
    double a = d2 + MUSCL_EPS_, bL = d1 + MUSCL_EPS_, bR = d3 + MUSCL_EPS_;
    double t2 = c2*d2;
    double vL = fma( c1*d1, a, t2*bL ) * fast_rcp( a + bL );
    double vR = fma( c1*d3, a, t2*bR ) * fast_rcp( a + bR );
    // same strict sign <=> positive product (|a|, |b| are >= ~1e-9 unless a difference hits -1e-9
    // to the last bit, so the product cannot underflow in practice); integer sign-bit tests cost
    // 5 % more kernel time, the kernel being limited by instruction issue
#if MUSCL_SIGN_INT
    // the same test on the sign bits (integer pipe; differs only if a or b is exactly zero, where the
    // reference's quotients are 0/0 or x/0 anyway)
    int ha = __double2hiint( a );
    incL = (ha ^ __double2hiint( bL )) >= 0 ? vL : 0.0;
    incR = (ha ^ __double2hiint( bR )) >= 0 ? vR : 0.0;
#else
    incL = a*bL > 0.0 ? vL : 0.0;
    incR = a*bR > 0.0 ? vR : 0.0;
#endif
#else
    double a = d2 + MUSCL_EPS_, bL = d1 + MUSCL_EPS_, bR = d3 + MUSCL_EPS_;
    bool sL = (a > 0.0 && bL > 0.0) || (a < 0.0 && bL < 0.0);
    bool sR = (a > 0.0 && bR > 0.0) || (a < 0.0 && bR < 0.0);
    double iL = 2.0 * fast_rcp( a + bL ), iR = 2.0 * fast_rcp( a + bR );
    double phiL = sL ? a*iL : 0.0, phi_L_inv = sL ? bL*iL : 0.0;
    double phiR = sR ? a*iR : 0.0, phi_R_inv = sR ? bR*iR : 0.0;
    incL = 0.25*(d1*(1.0-MUSCL_K)*phiL + d2*(1.0+MUSCL_K)*phi_L_inv);
    incR = 0.25*(d3*(1.0-MUSCL_K)*phiR + d2*(1.0+MUSCL_K)*phi_R_inv);
#endif
  }
}

// gp/gq: the 15 gradient components of the two end nodes, element i at gp[i*gsp], gq[i*gsq]
template< bool EXACT >
__device__ __forceinline__ void muscl( const double* gp, int gsp, const double* gq, int gsq,
                                       const double vw[3], double l[NC], double r[NC],
                                       const double eps = MUSCL_EPS )
{
  double ls[NC], rs[NC], d1[NC], d3[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    ls[c] = l[c]; rs[c] = r[c];
    double g1 = gp[(c*3+0)*gsp]*vw[0] + gp[(c*3+1)*gsp]*vw[1] + gp[(c*3+2)*gsp]*vw[2];
    double g2 = gq[(c*3+0)*gsq]*vw[0] + gq[(c*3+1)*gsq]*vw[1] + gq[(c*3+2)*gsq]*vw[2];
    double delta2 = r[c] - l[c];
    d1[c] = 2.0 * g1 - delta2;
    d3[c] = 2.0 * g2 - delta2;
    double incL, incR;
    vanleer< EXACT >( d1[c], delta2, d3[c], incL, incR, eps );
    l[c] += incL;
    r[c] -= incR;
  }
  // first order where density or internal energy could turn negative (:129-130)
  if (ls[0] < d1[0] || ls[4] < d1[4]) {
    #pragma unroll
    for (int c=0; c<NC; ++c) l[c] = ls[c];
  }
  if (rs[0] < -d3[0] || rs[4] < -d3[4]) {
    #pragma unroll
    for (int c=0; c<NC; ++c) r[c] = rs[c];
  }
}

// (Rusanov, Riemann.cpp:369-478: rusanov_len in riecg_own.cuh, which takes the edge normal's length as an argument)

// ev (optional): what the scalar flux of the same edge needs (riecg_scalar.cuh), Riemann.cpp:636-643
__device__ __forceinline__ void hllc( double l[NC], double r[NC], const double n[3],
                                      const DParams& P, double f[NC], double* ev = nullptr )
{
  double g = P.gamma;
  double nx = -n[0], ny = -n[1], nz = -n[2];
  double len = sqrt( nx*nx + ny*ny + nz*nz );
  nx /= len; ny /= len; nz /= len;
  double qL = l[1]*nx + l[2]*ny + l[3]*nz;
  double qR = r[1]*nx + r[2]*ny + r[3]*nz;
  if (ev) { ev[0] = qL*len; ev[1] = qR*len; ev[2] = fmax( fabs(qL), fabs(qR) ) * len; }
  double pL = (l[0]*l[4]) * (g-1.0);
  double pR = (r[0]*r[4]) * (g-1.0);
  l[4] = (l[4] + 0.5*(l[1]*l[1] + l[2]*l[2] + l[3]*l[3])) * l[0];
  l[1] *= l[0]; l[2] *= l[0]; l[3] *= l[0];
  r[4] = (r[4] + 0.5*(r[1]*r[1] + r[2]*r[2] + r[3]*r[3])) * r[0];
  r[1] *= r[0]; r[2] *= r[0]; r[3] *= r[0];
  double cL = sqrt( g * pL / l[0] );
  double cR = sqrt( g * pR / r[0] );
  double sL = fmin( qL - cL, qR - cR );
  double sR = fmax( qL + cL, qR + cR );
  double tL = sL - qL;
  double tR = sR - qR;
  double sM = (r[0]*qR*tR - l[0]*qL*tL + pL - pR) / (r[0]*tR - l[0]*tL);
  double pS = pL - l[0]*tL*(qL - sM);
  double uL[NC], uR[NC];
  double s = sL - sM;
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
  double L2 = -2.0*len;
  nx *= L2; ny *= L2; nz *= L2;
  if (sL > 0.0) {
    double qL2 = qL * L2;
    f[0] = l[0]*qL2;
    f[1] = l[1]*qL2 + pL*nx;
    f[2] = l[2]*qL2 + pL*ny;
    f[3] = l[3]*qL2 + pL*nz;
    f[4] = (l[4] + pL)*qL2;
  } else if (sL <= 0.0 && sM > 0.0) {
    double qL2 = qL * L2, sL2 = sL * L2;
    f[0] = l[0]*qL2 + sL2*(uL[0] - l[0]);
    f[1] = l[1]*qL2 + pL*nx + sL2*(uL[1] - l[1]);
    f[2] = l[2]*qL2 + pL*ny + sL2*(uL[2] - l[2]);
    f[3] = l[3]*qL2 + pL*nz + sL2*(uL[3] - l[3]);
    f[4] = (l[4] + pL)*qL2 + sL2*(uL[4] - l[4]);
  } else if (sM <= 0.0 && sR >= 0.0) {
    double qR2 = qR * L2, sR2 = sR * L2;
    f[0] = r[0]*qR2 + sR2*(uR[0] - r[0]);
    f[1] = r[1]*qR2 + pR*nx + sR2*(uR[1] - r[1]);
    f[2] = r[2]*qR2 + pR*ny + sR2*(uR[2] - r[2]);
    f[3] = r[3]*qR2 + pR*nz + sR2*(uR[3] - r[3]);
    f[4] = (r[4] + pR)*qR2 + sR2*(uR[4] - r[4]);
  } else {
    double qR2 = qR * L2;
    f[0] = r[0]*qR2;
    f[1] = r[1]*qR2 + pR*nx;
    f[2] = r[2]*qR2 + pR*ny;
    f[3] = r[3]*qR2 + pR*nz;
    f[4] = (r[4] + pR)*qR2;
  }
  if (P.stab2) {
    double sl = fabs(qL) + cL, sr = fabs(qR) + cR;
    double fws = P.stab2coef * fmax(sl,sr) * len;
    #pragma unroll
    for (int c=0; c<NC; ++c) f[c] -= fws * (l[c] - r[c]);
  }
}

// lax::sigvel, Lax.cpp:362-388: signal velocities of the preconditioned system
__device__ __forceinline__ void lax_sigvel( double p, double T, double v, double vn, const DParams& P,
                                            double& vpri, double& cpri )
{
  double g = P.gamma, rgas = P.rgas;
  double cp = g*rgas/(g-1.0);
  double r = p/T/rgas;
  double rp = r/p;
  double rt = -r/T;
  double vr = lax_refvel( r, p, v, g, P.kvinf );
  double vr2 = vr*vr;
  double beta = rp + rt/r/cp;
  double alpha = 0.5*(1.0 - beta*vr2);
  vpri = vn*(1.0 - alpha);
  cpri = sqrt( alpha*alpha*vn*vn + vr2 );
}

// edge-end state (p,u,v,w,T) -> conserved, Lax.cpp:438-451
__device__ __forceinline__ void lax_edge_conserved( double l[NC], double pL, const DParams& P ) {
  l[0] = pL/l[4]/P.rgas;
  l[1] *= l[0]; l[2] *= l[0]; l[3] *= l[0];
  l[4] = pL/(P.gamma-1.0) + 0.5*(l[1]*l[1] + l[2]*l[2] + l[3]*l[3])/l[0];
}

// lax::rusanov, Lax.cpp:390-511
__device__ __forceinline__ void lax_rusanov( double l[NC], double r[NC], const double n[3],
                                             const DParams& P, double f[NC] )
{
  double nx = n[0], ny = n[1], nz = n[2];
  double vnL = l[1]*nx + l[2]*ny + l[3]*nz;
  double vnR = r[1]*nx + r[2]*ny + r[3]*nz;
  double pL = l[0], pR = r[0];
  double len = sqrt( nx*nx + ny*ny + nz*nz );
  double vpL, cpL, vpR, cpR;
  lax_sigvel( l[0], l[4], sqrt( l[1]*l[1] + l[2]*l[2] + l[3]*l[3] ), vnL, P, vpL, cpL );
  lax_sigvel( r[0], r[4], sqrt( r[1]*r[1] + r[2]*r[2] + r[3]*r[3] ), vnR, P, vpR, cpR );
  lax_edge_conserved( l, pL, P );
  lax_edge_conserved( r, pR, P );
  double sp = fmax( fabs(vpL-cpL), fmax( fabs(vpR-cpR), fmax( fabs(vpL+cpL), fabs(vpR+cpR) ) ) );
  double fw = fmax( -sp, sp ) * len;
  f[0] = l[0]*vnL + r[0]*vnR + fw*(r[0] - l[0]);
  f[1] = l[1]*vnL + r[1]*vnR + (pL + pR)*nx + fw*(r[1] - l[1]);
  f[2] = l[2]*vnL + r[2]*vnR + (pL + pR)*ny + fw*(r[2] - l[2]);
  f[3] = l[3]*vnL + r[3]*vnR + (pL + pR)*nz + fw*(r[3] - l[3]);
  f[4] = (l[4] + pL)*vnL + (r[4] + pR)*vnR + fw*(r[4] - l[4]);
  if (P.stab2) {
    double fws = P.stab2coef * fw;
    #pragma unroll
    for (int c=0; c<NC; ++c) f[c] -= fws*(l[c] - r[c]);
  }
}

// lax::hllc, Lax.cpp:513-723 (wave speed option 3: symmetric +-sp; no artificial viscosity)
__device__ __forceinline__ void lax_hllc( double l[NC], double r[NC], const double n[3],
                                          const DParams& P, double f[NC] )
{
  double nx = -n[0], ny = -n[1], nz = -n[2];
  double len = sqrt( nx*nx + ny*ny + nz*nz );
  nx /= len; ny /= len; nz /= len;
  double qL = l[1]*nx + l[2]*ny + l[3]*nz;
  double qR = r[1]*nx + r[2]*ny + r[3]*nz;
  double pL = l[0], pR = r[0];
  double vpL, cpL, vpR, cpR;
  lax_sigvel( l[0], l[4], sqrt( l[1]*l[1] + l[2]*l[2] + l[3]*l[3] ), qL*len, P, vpL, cpL );
  lax_sigvel( r[0], r[4], sqrt( r[1]*r[1] + r[2]*r[2] + r[3]*r[3] ), qR*len, P, vpR, cpR );
  lax_edge_conserved( l, pL, P );
  lax_edge_conserved( r, pR, P );
  double sp = fmax( fabs(vpL-cpL), fmax( fabs(vpR-cpR), fmax( fabs(vpL+cpL), fabs(vpR+cpR) ) ) );
  double sL = -sp, sR = +sp;
  double tL = sL - qL;
  double tR = sR - qR;
  double sM = (r[0]*qR*tR - l[0]*qL*tL + pL - pR) / (r[0]*tR - l[0]*tL);
  double pS = pL - l[0]*tL*(qL - sM);
  double uL[NC], uR[NC];
  double s = sL - sM;
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
  double L2 = -2.0*len;
  nx *= L2; ny *= L2; nz *= L2;
  if (sL > 0.0) {
    double qL2 = qL * L2;
    f[0] = l[0]*qL2;
    f[1] = l[1]*qL2 + pL*nx;
    f[2] = l[2]*qL2 + pL*ny;
    f[3] = l[3]*qL2 + pL*nz;
    f[4] = (l[4] + pL)*qL2;
  } else if (sL <= 0.0 && sM > 0.0) {
    double qL2 = qL * L2, sL2 = sL * L2;
    f[0] = l[0]*qL2 + sL2*(uL[0] - l[0]);
    f[1] = l[1]*qL2 + pL*nx + sL2*(uL[1] - l[1]);
    f[2] = l[2]*qL2 + pL*ny + sL2*(uL[2] - l[2]);
    f[3] = l[3]*qL2 + pL*nz + sL2*(uL[3] - l[3]);
    f[4] = (l[4] + pL)*qL2 + sL2*(uL[4] - l[4]);
  } else if (sM <= 0.0 && sR >= 0.0) {
    double qR2 = qR * L2, sR2 = sR * L2;
    f[0] = r[0]*qR2 + sR2*(uR[0] - r[0]);
    f[1] = r[1]*qR2 + pR*nx + sR2*(uR[1] - r[1]);
    f[2] = r[2]*qR2 + pR*ny + sR2*(uR[2] - r[2]);
    f[3] = r[3]*qR2 + pR*nz + sR2*(uR[3] - r[3]);
    f[4] = (r[4] + pR)*qR2 + sR2*(uR[4] - r[4]);
  } else {
    double qR2 = qR * L2;
    f[0] = r[0]*qR2;
    f[1] = r[1]*qR2 + pR*nx;
    f[2] = r[2]*qR2 + pR*ny;
    f[3] = r[3]*qR2 + pR*nz;
    f[4] = (r[4] + pR)*qR2;
  }
}

// ---------------------------------------------------------------------------------
// flux gather per node (+ boundary + source) and, fused, the RK stage update
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void rhs_sum( size_t p, int lane, long long base, int kmax,
    const int* __restrict__ inc_e, const double* __restrict__ F, size_t nslot,
    const int* __restrict__ bslot, const double* __restrict__ Rb, const double* __restrict__ S,
    int src_mask, const double* __restrict__ v, double acc[NC] )
{
  #pragma unroll
  for (int c=0; c<NC; ++c) acc[c] = 0.0;
  // sg*f is exact, so this equals the add/subtract of the reference's scatter
  #pragma unroll kRhsUnroll
  for (int k=0; k<kmax; ++k) {
    int se = __ldg( inc_e + base + (long long)k*32 + lane );
    double sg = se > 0 ? 1.0 : (se < 0 ? -1.0 : 0.0);
    size_t sl = se == 0 ? 0 : (size_t)(abs(se)-1);
    double f[NC];
    load_f( F, nslot, sl, f );
    #pragma unroll
    for (int c=0; c<NC; ++c) acc[c] = fma( sg, se == 0 ? 0.0 : f[c], acc[c] );   // (a padding entry must not pass on a NaN of slot 0)
  }
  int b = bslot[p];
  if (b >= 0) {
    #pragma unroll
    for (int c=0; c<NC; ++c) acc[c] += Rb[(size_t)b*NC+c];
  }
  if (src_mask) {
    double vp = v[p];
    #pragma unroll
    for (int c=0; c<NC; ++c) if (src_mask & (1<<c)) acc[c] -= S[p*NC+c] * vp;
  }
}

// RK stage update of one node from its summed rhs.
//   RieCG::solve, RieCG.cpp:1011-1021:  u = un - rk dt rhs / vol   (dt = local dtp if steady)
//   LaxCG::solve, LaxCG.cpp:1150-1176:  (p,u,v,w,T) = (..)_n + P^-1 (-rk dt rhs / vol), then
//   back to conserved variables; W keeps what LaxCG::primitive gives for the next stage.
struct StageArgs { double rk, dt; const double* dtp; int stage; Mode M; };

template< bool LAX >
__device__ __forceinline__ void node_update( size_t p, size_t NP, const double acc[NC], double vp,
    const double* __restrict__ Un, double* __restrict__ U, const double* Win, double* W,
    double* __restrict__ Wn, double* __restrict__ UnOut, const StageArgs& A )
{
  // Win: primitives of this stage; W: where the new ones go (the same array, or the other one of
  // two buffers when threads of other blocks may still read this node's old values)
  double dtl = A.dtp ? A.dtp[p] : A.dt;
  double u[NC], w[NC];
  if (LAX) {
    double g = A.M.gamma, rgas = A.M.rgas;
    double wn[NC];
    #pragma unroll
    for (int c=0; c<NC; ++c) w[c] = get_w( Win, NP, c, p );
    if (A.stage == 0) {
      #pragma unroll
      for (int c=0; c<NC; ++c) { wn[c] = w[c]; Wn[c*NP+p] = w[c]; }
    } else {
      #pragma unroll
      for (int c=0; c<NC; ++c) wn[c] = Wn[c*NP+p];
    }
    double R = -A.rk * dtl / vp;
    // inverse of the time-derivative preconditioning matrix, LaxCG::precond :166-226
    double pr = w[0], uu = w[1], vv = w[2], ww = w[3], T = w[4];
    double r = pr/T/rgas;
    double cp = g*rgas/(g-1.0);
    double k = uu*uu + vv*vv + ww*ww;
    double vr = lax_refvel( r, pr, sqrt(k), g, A.M.kvinf );
    double vr2 = vr*vr;
    double rt = -r/T;
    double H = cp*T + k/2.0;
    double theta = 1.0/vr2 - rt/r/cp;
    double coef = r*cp*theta + rt;
    double q[NC] = { R*acc[0], R*acc[1], R*acc[2], R*acc[3], R*acc[4] };
    double wnew[NC];
    wnew[0] = wn[0] + (rt*(H - k) + r*cp)/coef*q[0] + rt*uu/coef*q[1] + rt*vv/coef*q[2] + rt*ww/coef*q[3] + (-rt/coef)*q[4];
    wnew[1] = wn[1] + (-uu/r)*q[0] + 1.0/r*q[1] + 0.0*q[2] + 0.0*q[3] + 0.0*q[4];
    wnew[2] = wn[2] + (-vv/r)*q[0] + 0.0*q[1] + 1.0/r*q[2] + 0.0*q[3] + 0.0*q[4];
    wnew[3] = wn[3] + (-ww/r)*q[0] + 0.0*q[1] + 0.0*q[2] + 1.0/r*q[3] + 0.0*q[4];
    wnew[4] = wn[4] + (-(theta*(H - k) - 1.0)/coef)*q[0] + (-theta*uu/coef)*q[1] + (-theta*vv/coef)*q[2]
                    + (-theta*ww/coef)*q[3] + theta/coef*q[4];
    lax_conservative( wnew, u, g, rgas );
    lax_primitive( u, w, g, rgas );
    #pragma unroll
    for (int c=0; c<NC; ++c) U[c*NP+p] = u[c];
    store_w( W, NP, p, w );
    if (A.stage == 2) {                   // conservative( m_un ) for the diagnostics, :1196
      double un[NC];
      lax_conservative( wn, un, g, rgas );
      #pragma unroll
      for (int c=0; c<NC; ++c) UnOut[c*NP+p] = un[c];
    }
  } else {
    double rkdtv = A.rk * dtl / vp;      // rk dt R / vol with one quotient per node
    #pragma unroll
    for (int c=0; c<NC; ++c) { u[c] = Un[c*NP+p] - rkdtv * acc[c]; U[c*NP+p] = u[c]; }
    primitive( u, w );
    store_w( W, NP, p, w );
  }
}

__global__ void k_rhs_shared( int nsh, const int* __restrict__ sh_node,
            const long long* __restrict__ sl_base, const int* __restrict__ inc_e,
            const double* __restrict__ F, size_t nslot, const int* __restrict__ bslot,
            const double* __restrict__ Rb, const double* __restrict__ S, int src_mask,
            const double* __restrict__ v, double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  size_t slice = p >> 5; int lane = p & 31;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double acc[NC];
  rhs_sum( p, lane, base, kmax, inc_e, F, nslot, bslot, Rb, S, src_mask, v, acc );
  for (int c=0; c<NC; ++c) part[(size_t)i*NC+c] = acc[c];
}

template< bool FUSED >
__global__ void k_rhs_finish( int nsh, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
            const int* __restrict__ ridx, const double* __restrict__ part,
            const double* __restrict__ recvbuf, const double* __restrict__ vol,
            const double* __restrict__ Un, StageArgs A, double* __restrict__ U, const double* Win,
            double* W, double* __restrict__ R, double* __restrict__ Wn, double* __restrict__ UnOut )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  double acc[NC];
  for (int c=0; c<NC; ++c) {
    double a = part[(size_t)i*NC+c];
    for (int r=roff[i]; r<roff[i+1]; ++r) a += recvbuf[(size_t)ridx[r]*NC+c];
    acc[c] = a;
  }
  if (FUSED) {
    if (A.M.rgas > 0.0) node_update< true >( p, NP, acc, vol[p], Un, U, Win, W, Wn, UnOut, A );
    else node_update< false >( p, NP, acc, vol[p], Un, U, Win, W, Wn, UnOut, A );
  } else {
    for (int c=0; c<NC; ++c) R[p*NC+c] = acc[c];
  }
}

// unfused RK update from a materialised R (drop-in for RieCG::solve :1016-1021)
__global__ void k_update( size_t npoin, size_t NP, const double* __restrict__ R, const double* __restrict__ vol,
                          const double* __restrict__ Un, StageArgs A, double* __restrict__ U,
                          double* __restrict__ W, double* __restrict__ Wn, double* __restrict__ UnOut, int rstride )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  double acc[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) acc[c] = R[p*(size_t)rstride+c];
  if (A.M.rgas > 0.0) node_update< true >( p, NP, acc, vol[p], Un, U, W, W, Wn, UnOut, A );
  else node_update< false >( p, NP, acc, vol[p], Un, U, W, W, Wn, UnOut, A );
}

// ---------------------------------------------------------------------------------
// boundary conditions, BC.cpp:29-241, one thread per BC node applying dirbc, symbc,
// farbc, prebc in the reference's order, then refreshing the primitive variables
// ---------------------------------------------------------------------------------
struct FarState { double r, p, u, v, w; };

__global__ void k_bc( int nbc, size_t NP, const int* __restrict__ node, const int* __restrict__ dir,
                      const int* __restrict__ dir_mask, const double* __restrict__ dir_val,
                      const int* __restrict__ symoff, const double* __restrict__ sym_n,
                      const int* __restrict__ faroff, const double* __restrict__ far_n, FarState fs,
                      const int* __restrict__ pre, const double* __restrict__ pre_val,
                      double gamma, double* __restrict__ U, double* __restrict__ W, Mode M )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nbc) return;
  size_t p = node[i];
  double u[NC];
  for (int c=0; c<NC; ++c) u[c] = U[c*NP+p];
  int d = dir[i];
  if (d >= 0) for (int c=0; c<NC; ++c) if (dir_mask[d*NC+c] == 1) u[c] = dir_val[d*NC+c];
  for (int s=symoff[i]; s<symoff[i+1]; ++s) {                 // symbc, BC.cpp:110-136
    const double* n = sym_n + (size_t)s*3;
    double vn = u[1]*n[0] + u[2]*n[1] + u[3]*n[2];
    u[1] -= vn * n[0];
    u[2] -= vn * n[1];
    u[3] -= vn * n[2];
  }
  for (int s=faroff[i]; s<faroff[i+1]; ++s) {                 // farbc, BC.cpp:152-220
    const double* n = far_n + (size_t)s*3;
    double vn = fs.u*n[0] + fs.v*n[1] + fs.w*n[2];
    double a = sqrt( gamma * fs.p / fs.r );
    double M = vn / a;
    if (M <= -1.0) {
      u[0] = fs.r; u[1] = fs.r*fs.u; u[2] = fs.r*fs.v; u[3] = fs.r*fs.w;
      u[4] = fs.p/(gamma-1.0) + 0.5*fs.r*(fs.u*fs.u + fs.v*fs.v + fs.w*fs.w);
    } else if (M > -1.0 && M < 0.0) {
      double pr = (u[4] - 0.5*(u[1]*u[1] + u[2]*u[2] + u[3]*u[3])/u[0]) * (gamma-1.0);
      u[0] = fs.r; u[1] = fs.r*fs.u; u[2] = fs.r*fs.v; u[3] = fs.r*fs.w;
      u[4] = pr/(gamma-1.0) + 0.5*fs.r*(fs.u*fs.u + fs.v*fs.v + fs.w*fs.w);
    } else if (M >= 0.0 && M < 1.0) {
      double uu = u[1]/u[0], vv = u[2]/u[0], ww = u[3]/u[0];
      u[4] = fs.p/(gamma-1.0) + 0.5*u[0]*(uu*uu + vv*vv + ww*ww);
    }
  }
  int pb = pre[i];
  if (pb >= 0) {                                              // prebc, BC.cpp:222-241
    u[0] = pre_val[pb*2+0];
    double uu = u[1]/u[0], vv = u[2]/u[0], ww = u[3]/u[0];
    u[4] = pre_val[pb*2+1]/(gamma-1.0) + 0.5*u[0]*(uu*uu + vv*vv + ww*ww);
  }
  double w[NC];
  for (int c=0; c<NC; ++c) U[c*NP+p] = u[c];
  primitive_of( u, w, M );
  store_w( W, NP, p, w );
}

// ---------------------------------------------------------------------------------
// reductions: time step (RieCG.cpp:827-839) and diagnostics (NodeDiagnostics.cpp:85-118)
// two-pass, fixed tree => deterministic
// ---------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 1184;   // 8 x 148 SMs
constexpr int RED_THREADS = 256;

template< int NV, bool MIN >
__device__ __forceinline__ void block_reduce( double v[NV], double* __restrict__ out )
{
  __shared__ double sm[NV][RED_THREADS/32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  #pragma unroll
  for (int i=0; i<NV; ++i) {
    double a = v[i];
    #pragma unroll
    for (int o=16; o>0; o>>=1) { double b = __shfl_xor_sync( 0xffffffffu, a, o ); a = MIN ? fmin(a,b) : a + b; }
    if (lane == 0) sm[i][w] = a;
  }
  __syncthreads();
  if (w == 0) {
    #pragma unroll
    for (int i=0; i<NV; ++i) {
      double a = lane < RED_THREADS/32 ? sm[i][lane] : (MIN ? 1.7976931348623157e308 : 0.0);
      #pragma unroll
      for (int o=16; o>0; o>>=1) { double b = __shfl_xor_sync( 0xffffffffu, a, o ); a = MIN ? fmin(a,b) : a + b; }
      if (lane == 0) out[(size_t)blockIdx.x*NV+i] = a;
    }
  }
}

// characteristic length cbrt(vol) of RieCG::dt (RieCG.cpp:834): a constant of the mesh, computed once
__global__ void k_cbrt( size_t n, const double* __restrict__ vol, double* __restrict__ out )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p < n) out[p] = cbrt( vol[p] );
}

// (vol: the nodes' cbrt(vol) from k_cbrt)
__global__ void __launch_bounds__(RED_THREADS)
k_dt( size_t npoin, size_t NP, const double* __restrict__ U, const double* __restrict__ vol, double gamma,
      double* __restrict__ part, Mode M, double cfl, double* __restrict__ dtp )
{
  double m[1] = { 1.7976931348623157e308 };
  const size_t stride = (size_t)gridDim.x*blockDim.x;
  for (size_t p0 = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p0 < npoin; p0 += 4*stride) {
    double a[4][6];
    #pragma unroll
    for (int k=0; k<4; ++k) {               // 24 independent loads in flight per thread
      size_t p = min( p0 + k*stride, npoin-1 );
      #pragma unroll
      for (int c=0; c<NC; ++c) a[k][c] = U[c*NP+p];
      a[k][5] = vol[p];
    }
    #pragma unroll
    for (int k=0; k<4; ++k) {
      double r = a[k][0], u = a[k][1]/r, v = a[k][2]/r, w = a[k][3]/r;
      double L = a[k][5];
      double e;
      if (M.rgas > 0.0) {                  // LaxCG::charvel, LaxCG.cpp:228-259
        double cp = gamma*M.rgas/(gamma-1.0);
        double kk = u*u + v*v + w*w;
        double ei = a[k][4]/r - kk/2.0;
        double pr = (r*ei) * (gamma-1.0);
        double T = pr/r/M.rgas;
        double rp = r/pr;
        double rt = -r/T;
        double vel = sqrt( kk );
        double vr = lax_refvel( r, pr, vel, gamma, M.kvinf );
        double vr2 = vr*vr;
        double beta = rp + rt/r/cp;
        double alpha = 0.5*(1.0 - beta*vr2);
        double vpri = vel*(1.0 - alpha);
        double cpri = sqrt( alpha*alpha*kk + vr2 );
        e = L / fmax( fabs(vpri) + cpri, 1.0e-8 );
      } else {
        double pr = (a[k][4] - 0.5*r*(u*u + v*v + w*w)) * (gamma-1.0);
        double c = sqrt( gamma * fmax(pr,0.0) / r );
        double vel = sqrt( u*u + v*v + w*w );
        e = L / fmax( vel+c, 1.0e-8 );
      }
      if (dtp && p0 + k*stride < npoin) dtp[p0 + k*stride] = e * cfl;   // local time step (steady)
      m[0] = fmin( m[0], e );
    }
  }
  block_reduce< 1, true >( m, part );
}

template< int NV, bool MIN >
__global__ void __launch_bounds__(RED_THREADS)
k_reduce_final( int nblocks, const double* __restrict__ part, double* __restrict__ out )
{
  double a[NV];
  #pragma unroll
  for (int i=0; i<NV; ++i) a[i] = MIN ? 1.7976931348623157e308 : 0.0;
  for (int b=threadIdx.x; b<nblocks; b+=blockDim.x) {
    #pragma unroll
    for (int i=0; i<NV; ++i) { double x = part[(size_t)b*NV+i]; a[i] = MIN ? fmin(a[i],x) : a[i] + x; }
  }
  block_reduce< NV, MIN >( a, out );
}

constexpr int NDIAG = 4*NC+1;
__global__ void __launch_bounds__(RED_THREADS)
k_diag( size_t npoin, size_t NP, const double* __restrict__ U, const double* __restrict__ Un,
        const double* __restrict__ v, const double* __restrict__ an, double* __restrict__ part )
{
  double a[NDIAG];
  #pragma unroll
  for (int i=0; i<NDIAG; ++i) a[i] = 0.0;
  for (size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p < npoin; p += (size_t)gridDim.x*blockDim.x) {
    double vp = v[p], u[NC];
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      u[c] = U[c*NP+p];
      double d = u[c] - Un[c*NP+p];
      a[c] += u[c]*u[c]*vp;
      a[NC+c] += d*d*vp;
    }
    a[2*NC] += u[4]*vp;
    if (an) {
      double w[NC];
      primitive( u, w );
      #pragma unroll
      for (int c=0; c<NC; ++c) {
        double du = w[c] - an[p*NC+c];
        a[2*NC+1+c] += du*du*vp;
        a[3*NC+1+c] += fabs(du)*vp;
      }
    }
  }
  block_reduce< NDIAG, false >( a, part );
}

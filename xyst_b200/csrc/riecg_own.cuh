// xyst_b200/csrc/riecg_own.cuh -- thread-per-owner edge flux kernel of RieCG / LaxCG and the nodal gather + update
// Part of the single translation unit xyst_b200.cu (included inside its anonymous namespace).
//
// One thread per OWNER node (the lower endpoint of an edge in the library's node order): the owner's
// primitives and coordinates stay in registers over its ~7 owned edges (its 15 gradient components
// in the thread's own column of shared memory), the other end arrives with 12 16-byte gathers. An
// edge is always evaluated owner-first. Where the reference's orientation
// (src/Inciter/RieCG.cpp:659-714) has the owner as the edge's SECOND node, the limiter's epsilon
// (src/Physics/Riemann.cpp:92-95) and the edge normal are negated: with d1' = -d3, d2' = -d2,
// d3' = -d1 and eps' = -eps every quotient of the van Leer function is the reference's (numerator
// and denominator both negated), so the two reconstructed states are exactly the reference's,
// swapped, and the Riemann flux is the reference's with the opposite sign. Hence always:
// owner -= f', other end += f'.
//
//  k_flux_own   : per-edge fluxes F (reference-oriented) + the owner's own share of its nodal sum (Racc)
//  k_update_in  : Racc + the fluxes of the node's INCOMING edges + boundary + source, RK update
//
// Measured alternatives that were removed again: a fused per-tile kernel keeping
// fluxes in shared memory (with and without look-back between tiles), cp.async staging of the
// other end one edge ahead, the thread-per-edge kernel of round 1.

#ifndef OWN_THREADS
#define OWN_THREADS 128
#endif
#ifndef OWN_MINB
#define OWN_MINB 4
#endif
#ifndef OWN_UNROLL
#define OWN_UNROLL 1
#endif
#ifndef OWN_GSMEM
#define OWN_GSMEM 1        // 1: the owner's 15 gradient components live in shared memory, not in registers
#endif
constexpr int kOwnUnroll = OWN_UNROLL;

__device__ __forceinline__ void load_g( const double2* __restrict__ G2, size_t NP, size_t p, double g[15] ) {
  double2 a[NGP];
  #pragma unroll
  for (int k=0; k<NGP; ++k) a[k] = __ldg( G2 + (size_t)k*NP + p );
  #pragma unroll
  for (int k=0; k<7; ++k) { g[2*k] = a[k].x; g[2*k+1] = a[k].y; }
  g[14] = a[7].x;
}

// sqrt(x) for x >= 0 without the IEEE routine's special-case branches (which split the edge loop
// into many basic blocks): hardware seed (MUFU.RSQ64H, ~20 bits), two coupled Goldschmidt steps on
// (sqrt x, 1/(2 sqrt x)) -> ~1 ulp. x < 0 gives NaN like sqrt().
__device__ __forceinline__ double fast_sqrt( double x )
{
  double y;
  asm( "rsqrt.approx.ftz.f64 %0, %1;" : "=d"( y ) : "d"( x ) );
  double g = x*y, h = 0.5*y;
  double r = fma( -g, h, 0.5 );
  g = fma( g, r, g ); h = fma( h, r, h );
  r = fma( -g, g, x );
  g = fma( r, h, g );
  return x == 0.0 ? 0.0 : g;
}

// rusanov (riecg_kernels.cuh; src/Physics/Riemann.cpp:369-478) with the edge normal's length given
// (a per-edge constant kept next to the normal) and the limiter form as a template argument
template< bool EXACT >
__device__ __forceinline__ void rusanov_len( double l[NC], double r[NC], const double n[3], double len,
                                             const DParams& P, double f[NC], double* ev = nullptr )
{
  double g = P.gamma;
  double pL = (l[0]*l[4]) * (g-1.0);
  double pR = (r[0]*r[4]) * (g-1.0);
  const double gg1 = g*(g-1.0), eL = l[4], eR = r[4];
  double nx = n[0], ny = n[1], nz = n[2];
  double vnL = l[1]*nx + l[2]*ny + l[3]*nz;
  double vnR = r[1]*nx + r[2]*ny + r[3]*nz;
  if (ev) { ev[0] = vnL; ev[1] = vnR; ev[2] = fmax( fabs(vnL), fabs(vnR) ); }   // for the scalar flux, Riemann.cpp:462-470
  l[4] = (l[4] + 0.5*(l[1]*l[1] + l[2]*l[2] + l[3]*l[3])) * l[0];
  l[1] *= l[0]; l[2] *= l[0]; l[3] *= l[0];
  r[4] = (r[4] + 0.5*(r[1]*r[1] + r[2]*r[2] + r[3]*r[3])) * r[0];
  r[1] *= r[0]; r[2] *= r[0]; r[3] *= r[0];
  double sl, sr;
  if (EXACT) { sl = fabs(vnL) + sqrt( g * pL / l[0] )*len; sr = fabs(vnR) + sqrt( g * pR / r[0] )*len; }
  else { sl = fabs(vnL) + fast_sqrt( gg1 * eL )*len; sr = fabs(vnR) + fast_sqrt( gg1 * eR )*len; }   // g p / rho = g (g-1) e
  double fw = fmax( sl, sr );
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

// flux f' of the edge (owner -> other), owner-first. s = +1: the owner is the reference's first
// node, -1: its second. n = the stored (reference-oriented) normal.
template< bool EXACT, int FLUX >
__device__ __forceinline__ void edge_flux_owner( const double wo[NC], const double xo[3], const double go[15],
    const double wq[NC], const double xq[3], const double gq[15], double s, const double nref[4],
    const DParams& P, double f[NC], double* ev = nullptr )
{
  double l[NC], r[NC], vw[3], n[3];
  #pragma unroll
  for (int c=0; c<NC; ++c) { l[c] = wo[c]; r[c] = wq[c]; }
  #pragma unroll
  for (int j=0; j<3; ++j) { vw[j] = xq[j] - xo[j]; n[j] = s * nref[j]; }
  muscl< EXACT >( go, 1, gq, 1, vw, l, r, s * MUSCL_EPS );
  if (FLUX == 0) rusanov_len< EXACT >( l, r, n, nref[3], P, f, ev ); else if (FLUX == 1) hllc( l, r, n, P, f, ev );
  else if (FLUX == 2) lax_rusanov( l, r, n, P, f ); else lax_hllc( l, r, n, P, f );
}

template< bool EXACT, int FLUX >
__global__ void __launch_bounds__(OWN_THREADS, OWN_MINB)
k_flux_own( size_t slice0, size_t nslice, size_t NP, size_t nslot, const long long* __restrict__ ebase, const int* __restrict__ eo,
            const double* __restrict__ D, const double* __restrict__ W, const double* __restrict__ G,
            double* __restrict__ F, double* __restrict__ Racc, DParams P, double* __restrict__ EV )
{
  // EV (null without transported scalars): per edge, reference-oriented, the normal velocities of the
  // reconstructed states and the scalar dissipation speed (riecg_scalar.cuh)
  size_t slice = slice0 + ((blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5);     // slices [slice0, nslice)
  int lane = threadIdx.x & 31;
  if (slice >= nslice) return;
  size_t p = slice*32 + lane;
  const double2* WX = reinterpret_cast< const double2* >( W );
  const double2* G2 = reinterpret_cast< const double2* >( G );
  long long b0 = ebase[slice];
  int kmax = (int)((ebase[slice+1] - b0) >> 5);
  double acc[NC] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
#if OWN_GSMEM
  __shared__ double2 sgo[NGP*OWN_THREADS];          // the thread's own column: no conflicts, no barrier
#endif
  if (kmax > 0) {
    double wo[NC], xo[3], go[15];
    load_wx( WX, NP, p, wo, xo );
#if OWN_GSMEM
    #pragma unroll
    for (int k=0; k<NGP; ++k) sgo[k*OWN_THREADS + threadIdx.x] = __ldg( G2 + (size_t)k*NP + p );
#else
    load_g( G2, NP, p, go );
#endif
    // the next edge's other end and normal are fetched one iteration ahead, so that its operand
    // gathers can leave as soon as the iteration starts
    size_t sl = (size_t)b0 + lane;
    int e_nx = __ldg( eo + sl );
    #pragma unroll kOwnUnroll
    for (int j=0; j<kmax; ++j) {
#if OWN_GSMEM
      { const volatile double2* src = sgo + threadIdx.x;
        #pragma unroll
        for (int k=0; k<7; ++k) { go[2*k] = src[k*OWN_THREADS].x; go[2*k+1] = src[k*OWN_THREADS].y; }
        go[14] = src[7*OWN_THREADS].x; }
#endif
      const int e = e_nx;
      // (the normal is needed last, by the Riemann solver: its load hides behind the limiter)
      const double n[4] = { __ldg( D + sl ), __ldg( D + nslot + sl ), __ldg( D + 2*nslot + sl ), __ldg( D + 3*nslot + sl ) };
      const bool valid = e != -1;
      const double s = e < 0 ? -1.0 : 1.0;
      const size_t q = valid ? (size_t)(e & 0x7fffffff) : p;
      double wq[NC], xq[3], gq[15];
      load_wx( WX, NP, q, wq, xq );
      load_g( G2, NP, q, gq );
      if (j+1 < kmax) e_nx = __ldg( eo + sl + 32 );
      double f[NC], ev[3];
      edge_flux_owner< EXACT, FLUX >( wo, xo, go, wq, xq, gq, s, n, P, f, EV ? ev : nullptr );
      if (valid) {
        #pragma unroll
        for (int c=0; c<NC; ++c) { acc[c] -= f[c]; f[c] *= s; }     // F holds the reference-oriented flux
        store_f( F, nslot, sl, f );
        if (EV) {                    // owner-first (a,b) -> reference orientation: (a,b) or (-b,-a)
          EV[sl] = e < 0 ? -ev[1] : ev[0]; EV[nslot+sl] = e < 0 ? -ev[0] : ev[1]; EV[2*nslot+sl] = ev[2];
        }
      }
      sl += 32;
    }
  }
  // the owner's own share of its nodal sum; the receivers' shares are gathered by k_update_in
  #pragma unroll
  for (int c=0; c<NC; ++c) Racc[c*NP+p] = acc[c];
}

// Receiver side of k_flux_own fused with the RK update: Racc + the node's INCOMING edges (owned by
// lower neighbours, ascending; entries as in the incidence lists: +(slot+1) if the node is the edge's
// second node, -(slot+1) if its first, 0 = padding) + boundary + source, then node_update.
template< bool FUSED, bool LAX >
__global__ void __launch_bounds__(NODE_THREADS, RHS_MINB)
k_update_in( size_t npoin, size_t NP, const long long* __restrict__ in_base, const int* __restrict__ in_e,
             const double* __restrict__ Racc, const double* __restrict__ F, size_t nslot,
             const int* __restrict__ bslot, const double* __restrict__ Rb, const double* __restrict__ S, int src_mask,
             const double* __restrict__ v, const double* __restrict__ vol, const double* __restrict__ Un,
             StageArgs A, double* __restrict__ U, double* __restrict__ W, double* __restrict__ R,
             double* __restrict__ Wn, double* __restrict__ UnOut, const unsigned char* __restrict__ skip, int rstride,
             size_t slice0, size_t slice1 )
{
  size_t slice = slice0 + ((blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5);     // slices [slice0, slice1)
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (slice >= slice1 || p >= npoin) return;
  if (FUSED && skip && skip[p]) return;
  long long base = in_base[slice];
  int kmax = (int)((in_base[slice+1] - base) >> 5);
  double acc[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) acc[c] = Racc[c*NP+p];
  #pragma unroll 7
  for (int k=0; k<kmax; ++k) {
    int se = __ldg( in_e + base + (long long)k*32 + lane );
    double sg = se > 0 ? 1.0 : (se < 0 ? -1.0 : 0.0);
    size_t sl = se == 0 ? 0 : (size_t)(abs(se)-1);
    double f[NC];
    load_f( F, nslot, sl, f );
    #pragma unroll
    for (int c=0; c<NC; ++c) acc[c] = fma( sg, se == 0 ? 0.0 : f[c], acc[c] );
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
  if (FUSED) {
    node_update< LAX >( p, NP, acc, vol[p], Un, U, W, W, Wn, UnOut, A );
  } else {
    #pragma unroll
    for (int c=0; c<NC; ++c) R[p*(size_t)rstride+c] = acc[c];      // row stride = all components (5 + scalars)
  }
}


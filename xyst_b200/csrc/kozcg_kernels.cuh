// xyst_b200/csrc/kozcg_kernels.cuh -- KozCG device code: element-based Taylor-Galerkin + flux-corrected transport
// Part of the single translation unit xyst_b200.cu (included inside its anonymous namespace).

// ---------------------------------------------------------------------------------
// KozCG: element-based Taylor-Galerkin (Kozak.cpp:29-180) and FCT (KozCG.cpp:774-1197).
// Each pass = one kernel over tetrahedra writing per-tet, per-local-node values, and one
// node gather over the incident tetrahedra (fixed order, no atomics).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double koz_geom( const double* __restrict__ X, size_t NP, const int N[4], double grad[4][3] )
{
  double x[4][3];
  #pragma unroll
  for (int a=0; a<4; ++a) { x[a][0] = X[N[a]]; x[a][1] = X[NP+N[a]]; x[a][2] = X[2*NP+N[a]]; }
  double ba[3] = { x[1][0]-x[0][0], x[1][1]-x[0][1], x[1][2]-x[0][2] },
         ca[3] = { x[2][0]-x[0][0], x[2][1]-x[0][1], x[2][2]-x[0][2] },
         da[3] = { x[3][0]-x[0][0], x[3][1]-x[0][1], x[3][2]-x[0][2] };
  grad[1][0] = ca[1]*da[2] - da[1]*ca[2]; grad[1][1] = ca[2]*da[0] - da[2]*ca[0]; grad[1][2] = ca[0]*da[1] - da[0]*ca[1];
  grad[2][0] = da[1]*ba[2] - ba[1]*da[2]; grad[2][1] = da[2]*ba[0] - ba[2]*da[0]; grad[2][2] = da[0]*ba[1] - ba[0]*da[1];
  grad[3][0] = ba[1]*ca[2] - ca[1]*ba[2]; grad[3][1] = ba[2]*ca[0] - ca[2]*ba[0]; grad[3][2] = ba[0]*ca[1] - ca[0]*ba[1];
  #pragma unroll
  for (int i=0; i<3; ++i) grad[0][i] = -grad[1][i]-grad[2][i]-grad[3][i];
  return ba[0]*grad[1][0] + ba[1]*grad[1][1] + ba[2]*grad[1][2];          // triple(ba,ca,da)
}

// pass 1: rhs contributions T[a*5+c] and antidiffusive element contributions T[20+c*4+a]
__global__ void __launch_bounds__(128)
k_koz_elem1( size_t ntet, size_t NP, const int* __restrict__ tet, const double* __restrict__ U,
             const double* __restrict__ X, const double* __restrict__ S, const double* __restrict__ Sc,
             double dt, double gamma, double ctau, int fct, double* __restrict__ T, double* __restrict__ UE )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= ntet) return;
  int N[4] = { tet[e], tet[ntet+e], tet[2*ntet+e], tet[3*ntet+e] };
  double grad[4][3];
  double J = koz_geom( X, NP, N, grad );
  double u[4][NC], p[4];
  #pragma unroll
  for (int a=0; a<4; ++a) {
    #pragma unroll
    for (int c=0; c<NC; ++c) u[a][c] = U[c*NP+N[a]];
    p[a] = (u[a][4] - 0.5*(u[a][1]*u[a][1] + u[a][2]*u[a][2] + u[a][3]*u[a][3])/u[a][0]) * (gamma-1.0);
  }
  double ue[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) ue[c] = (u[0][c] + u[1][c] + u[2][c] + u[3][c])/4.0;
  double coef = dt/J/2.0;
  #pragma unroll
  for (int j=0; j<3; ++j)
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double cg = coef * grad[a][j];
      double uj = u[a][j+1] / u[a][0];
      ue[0] -= cg * u[a][j+1];
      ue[1] -= cg * u[a][1] * uj;
      ue[2] -= cg * u[a][2] * uj;
      ue[3] -= cg * u[a][3] * uj;
      ue[j+1] -= cg * p[a];
      ue[4] -= cg * (u[a][4] + p[a]) * uj;
    }
  if (S) {
    coef = dt/8.0;
    #pragma unroll
    for (int a=0; a<4; ++a)
      #pragma unroll
      for (int c=0; c<NC; ++c) ue[c] += coef * S[(size_t)N[a]*NC+c];
  }
  if (UE) {                       // half-step density and momentum of the element for the transported scalars
    #pragma unroll
    for (int c=0; c<4; ++c) UE[(size_t)c*ntet+e] = ue[c];
  }
  double pr = (ue[4] - 0.5*(ue[1]*ue[1] + ue[2]*ue[2] + ue[3]*ue[3])/ue[0]) * (gamma-1.0);
  double R[4][NC];
  #pragma unroll
  for (int a=0; a<4; ++a)
    #pragma unroll
    for (int c=0; c<NC; ++c) R[a][c] = 0.0;
  coef = 1.0/6.0;
  #pragma unroll
  for (int j=0; j<3; ++j) {
    double uj = ue[j+1] / ue[0];
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double cg = coef * grad[a][j];
      R[a][0] += cg * ue[j+1];
      R[a][1] += cg * ue[1] * uj;
      R[a][2] += cg * ue[2] * uj;
      R[a][3] += cg * ue[3] * uj;
      R[a][j+1] += cg * pr;
      R[a][4] += cg * (ue[4] + pr) * uj;
    }
  }
  if (S) {
    coef = J/24.0;
    #pragma unroll
    for (int a=0; a<4; ++a)
      #pragma unroll
      for (int c=0; c<NC; ++c) R[a][c] += coef * Sc[e*NC+c];
  }
  #pragma unroll
  for (int a=0; a<4; ++a)
    #pragma unroll
    for (int c=0; c<NC; ++c) T[(size_t)(a*NC+c)*ntet+e] = R[a][c];
  if (fct) {
    #pragma unroll
    for (int c=0; c<NC; ++c)
      #pragma unroll
      for (int a=0; a<4; ++a) {
        double aec = 0.0;
        #pragma unroll
        for (int b=0; b<4; ++b) { double m = J/120.0 * ((a == b) ? 3.0 : -1.0); aec += m * ctau * u[b][c]; }
        T[(size_t)(20+c*4+a)*ntet+e] = aec;
      }
  }
}

// node pass 1: R, P+/-, symmetry BC on P, low-order solution ul = u + dt R/vol - P+ - P-
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_koz_node1( size_t npoin, size_t NP, size_t ntet, const long long* __restrict__ kbase, const int* __restrict__ kinc,
             const double* __restrict__ T, const double* __restrict__ U, const int* __restrict__ bcof,
             const int* __restrict__ symoff, const double* __restrict__ sym_n, const double* __restrict__ vol,
             double dt, int fct, double* __restrict__ P, double* __restrict__ UL, double* __restrict__ R )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = kbase[slice];
  int kmax = (int)((kbase[slice+1] - base) >> 5);
  double r[NC], pp[NC], pn[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { r[c] = 0.0; pp[c] = 0.0; pn[c] = 0.0; }
  for (int k=0; k<kmax; ++k) {
    int ta = __ldg( kinc + base + (long long)k*32 + lane );
    if (ta < 0) continue;
    size_t e = (size_t)(ta >> 2); int a = ta & 3;
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      r[c] += __ldg( T + (size_t)(a*NC+c)*ntet + e );
      if (fct) { double aec = __ldg( T + (size_t)(20+c*4+a)*ntet + e ); pp[c] += fmax( 0.0, aec ); pn[c] += fmin( 0.0, aec ); }
    }
  }
  #pragma unroll
  for (int c=0; c<NC; ++c) R[p*NC+c] = r[c];
  if (!fct) return;
  int bc = bcof[p];
  if (bc >= 0)
    for (int s=symoff[bc]; s<symoff[bc+1]; ++s) {
      const double* n = sym_n + (size_t)s*3;
      double rvnp = pp[1]*n[0] + pp[2]*n[1] + pp[3]*n[2];
      double rvnn = pn[1]*n[0] + pn[2]*n[1] + pn[3]*n[2];
      pp[1] -= rvnp * n[0]; pn[1] -= rvnn * n[0];
      pp[2] -= rvnp * n[1]; pn[2] -= rvnn * n[1];
      pp[3] -= rvnp * n[2]; pn[3] -= rvnn * n[2];
    }
  // one quotient per node, then products (the reference divides each component: same values up to one rounding)
  double ivp = 1.0 / vol[p];
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    pp[c] *= ivp; pn[c] *= ivp;
    P[(2*c)*NP+p] = pp[c]; P[(2*c+1)*NP+p] = pn[c];
    UL[c*NP+p] = U[c*NP+p] + dt*r[c]*ivp - pp[c] - pn[c];
  }
}

// pass 2: per-tet allowed bounds over its 4 nodes -> T[c*2], T[c*2+1]
// (M components: the flow's NC, or one transported scalar at a time with U, UL, T pointing at its rows)
template< int M >
__global__ void __launch_bounds__(128)
k_koz_elem2( size_t ntet, size_t NP, const int* __restrict__ tet, const double* __restrict__ U,
             const double* __restrict__ UL, int clip, double* __restrict__ T )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= ntet) return;
  int N[4] = { tet[e], tet[ntet+e], tet[2*ntet+e], tet[3*ntet+e] };
  #pragma unroll
  for (int c=0; c<M; ++c) {
    double alwp = -1.7976931348623157e308, alwn = 1.7976931348623157e308;
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double ul = UL[c*NP+N[a]];
      if (clip) { alwp = fmax( alwp, ul ); alwn = fmin( alwn, ul ); }
      else { double u = U[c*NP+N[a]]; alwp = fmax( alwp, fmax( ul, u ) ); alwn = fmin( alwn, fmin( ul, u ) ); }
    }
    T[(size_t)(2*c)*ntet+e] = alwp; T[(size_t)(2*c+1)*ntet+e] = alwn;
  }
}

// node pass 2: Q+/- = max/min over incident tets, minus ul, limit coefficients C+/-
template< int M >
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_koz_node2( size_t npoin, size_t NP, size_t ntet, const long long* __restrict__ kbase, const int* __restrict__ kinc,
             const double* __restrict__ T, const double* __restrict__ UL, const double* __restrict__ P,
             double* __restrict__ Q )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = kbase[slice];
  int kmax = (int)((kbase[slice+1] - base) >> 5);
  double qa[M], qb[M];
  #pragma unroll
  for (int c=0; c<M; ++c) { qa[c] = -1.7976931348623157e308; qb[c] = 1.7976931348623157e308; }
  for (int k=0; k<kmax; ++k) {
    int ta = __ldg( kinc + base + (long long)k*32 + lane );
    if (ta < 0) continue;
    size_t e = (size_t)(ta >> 2);
    #pragma unroll
    for (int c=0; c<M; ++c) {
      qa[c] = fmax( qa[c], __ldg( T + (size_t)(2*c)*ntet + e ) );
      qb[c] = fmin( qb[c], __ldg( T + (size_t)(2*c+1)*ntet + e ) );
    }
  }
  const double eps = 2.220446049250313e-16;
  #pragma unroll
  for (int c=0; c<M; ++c) {
    double ul = UL[c*NP+p];
    double a = qa[c] - ul, b = qb[c] - ul;
    double pa = P[(2*c)*NP+p], pb = P[(2*c+1)*NP+p];
    Q[(2*c)*NP+p]   = pa <  eps ? 0.0 : fmin( 1.0, a/pa );
    Q[(2*c+1)*NP+p] = pb > -eps ? 0.0 : fmin( 1.0, b/pb );
  }
}

// pass 3: limited antidiffusive element contributions coef[c]*aec[c][a] -> T[a*M+c]
template< int M >
__global__ void __launch_bounds__(128)
k_koz_elem3( size_t ntet, size_t NP, const int* __restrict__ tet, const double* __restrict__ U,
             const double* __restrict__ X, const double* __restrict__ Q, double ctau, int sysmask,
             double* __restrict__ T )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= ntet) return;
  int N[4] = { tet[e], tet[ntet+e], tet[2*ntet+e], tet[3*ntet+e] };
  double grad[4][3];
  double J = koz_geom( X, NP, N, grad );
  double coef[M], aec[M][4];
  #pragma unroll
  for (int c=0; c<M; ++c) {
    double u[4];
    #pragma unroll
    for (int a=0; a<4; ++a) u[a] = U[c*NP+N[a]];
    coef[c] = 1.0;
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double v = 0.0;
      #pragma unroll
      for (int b=0; b<4; ++b) { double m = J/120.0 * ((a == b) ? 3.0 : -1.0); v += m * ctau * u[b]; }
      aec[c][a] = v;
      coef[c] = fmin( coef[c], v > 0.0 ? Q[(2*c)*NP+N[a]] : Q[(2*c+1)*NP+N[a]] );
    }
  }
  double cs = 1.0;
  #pragma unroll
  for (int c=0; c<M; ++c) if (sysmask & (1<<c)) cs = fmin( cs, coef[c] );
  #pragma unroll
  for (int c=0; c<M; ++c) {
    if (sysmask & (1<<c)) coef[c] = cs;
    #pragma unroll
    for (int a=0; a<4; ++a) T[(size_t)(a*M+c)*ntet+e] = coef[c] * aec[c][a];
  }
}

// node pass 3: a = sum of limited contributions, u = ul + a/vol (KozCG.cpp:1140-1146)
template< int M, bool FLOW >
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_koz_node3( size_t npoin, size_t NP, size_t ntet, const long long* __restrict__ kbase, const int* __restrict__ kinc,
             const double* __restrict__ T, const double* __restrict__ UL, const double* __restrict__ vol,
             double* __restrict__ Unew, double* __restrict__ W )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = kbase[slice];
  int kmax = (int)((kbase[slice+1] - base) >> 5);
  double a_[M];
  #pragma unroll
  for (int c=0; c<M; ++c) a_[c] = 0.0;
  for (int k=0; k<kmax; ++k) {
    int ta = __ldg( kinc + base + (long long)k*32 + lane );
    if (ta < 0) continue;
    size_t e = (size_t)(ta >> 2); int a = ta & 3;
    #pragma unroll
    for (int c=0; c<M; ++c) a_[c] += __ldg( T + (size_t)(a*M+c)*ntet + e );
  }
  double ivp = 1.0 / vol[p], u[M];
  #pragma unroll
  for (int c=0; c<M; ++c) { u[c] = UL[c*NP+p] + a_[c]*ivp; Unew[c*NP+p] = u[c]; }
  if (FLOW) {
    double uu[NC], w[NC];
    #pragma unroll
    for (int c=0; c<NC; ++c) uu[c] = u[c < M ? c : 0];
    primitive( uu, w );
    store_w( W, NP, p, w );
  }
}

// ---- transported scalars (kozak::rhs scalar rows, Kozak.cpp:84-86,133-135; FCT as for the flow, no symmetry
// BC on scalars): one scalar at a time. s: the scalar's nodal values [NP]; UE: the element's half-step density
// and momentum left by k_koz_elem1; T rows: [a] rhs contributions, [4+a] antidiffusive element contributions
__global__ void __launch_bounds__(128)
k_koz_selem1( size_t ntet, size_t NP, const int* __restrict__ tet, const double* __restrict__ U,
              const double* __restrict__ sv, const double* __restrict__ X, const double* __restrict__ UE,
              double dt, double ctau, int fct, double* __restrict__ T )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= ntet) return;
  int N[4] = { tet[e], tet[ntet+e], tet[2*ntet+e], tet[3*ntet+e] };
  double grad[4][3];
  double J = koz_geom( X, NP, N, grad );
  double u[4];
  #pragma unroll
  for (int a=0; a<4; ++a) u[a] = sv[N[a]];
  double ue = (u[0] + u[1] + u[2] + u[3])/4.0;
  double coef = dt/J/2.0;
  #pragma unroll
  for (int j=0; j<3; ++j)
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double cg = coef * grad[a][j];
      double uj = U[(j+1)*NP+N[a]] / U[N[a]];
      ue -= cg * u[a] * uj;
    }
  double R[4] = { 0.0, 0.0, 0.0, 0.0 };
  double r0 = UE[e];
  coef = 1.0/6.0;
  #pragma unroll
  for (int j=0; j<3; ++j) {
    double uj = UE[(size_t)(j+1)*ntet+e] / r0;
    #pragma unroll
    for (int a=0; a<4; ++a) R[a] += coef * grad[a][j] * ue * uj;
  }
  #pragma unroll
  for (int a=0; a<4; ++a) T[(size_t)a*ntet+e] = R[a];
  if (fct) {
    #pragma unroll
    for (int a=0; a<4; ++a) {
      double aec = 0.0;
      #pragma unroll
      for (int b=0; b<4; ++b) { double m = J/120.0 * ((a == b) ? 3.0 : -1.0); aec += m * ctau * u[b]; }
      T[(size_t)(4+a)*ntet+e] = aec;
    }
  }
}

// node pass 1 of one scalar: P+/-, low-order solution ul = s + dt r/vol - P+ - P-; without FCT s_new = s + dt r/vol
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_koz_snode1( size_t npoin, size_t NP, size_t ntet, const long long* __restrict__ kbase, const int* __restrict__ kinc,
              const double* __restrict__ T, const double* __restrict__ sv, const double* __restrict__ vol,
              double dt, int fct, double* __restrict__ P, double* __restrict__ UL )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = kbase[slice];
  int kmax = (int)((kbase[slice+1] - base) >> 5);
  double r = 0.0, pp = 0.0, pn = 0.0;
  for (int k=0; k<kmax; ++k) {
    int ta = __ldg( kinc + base + (long long)k*32 + lane );
    if (ta < 0) continue;
    size_t e = (size_t)(ta >> 2); int a = ta & 3;
    r += __ldg( T + (size_t)a*ntet + e );
    if (fct) { double aec = __ldg( T + (size_t)(4+a)*ntet + e ); pp += fmax( 0.0, aec ); pn += fmin( 0.0, aec ); }
  }
  double ivp = 1.0 / vol[p];
  if (!fct) { UL[p] = sv[p] + dt*r*ivp; return; }
  pp *= ivp; pn *= ivp;
  P[p] = pp; P[NP+p] = pn;
  UL[p] = sv[p] + dt*r*ivp - pp - pn;
}

// ---- several partitions (KozCG::comrhs/comaec :728,:839, comalw :944, comlim :1093) -------------------
// The full-mesh node passes above leave partial values at the nodes shared with other partitions. For
// those nodes the OWN sums of a pass are recomputed into the exchange buffer (mode 1: r, P+, P- with the
// symmetry BC on the own P as in the reference; mode 2: own bounds; mode 3: own limited sums), travel
// (sum, max/min, sum), and the nodes are finished from the complete values.
__global__ void k_koz_sh( int mode, int nsh, size_t NP, size_t ntet, const int* __restrict__ sh_node,
             const long long* __restrict__ kbase, const int* __restrict__ kinc, const double* __restrict__ T,
             const int* __restrict__ bcof, const int* __restrict__ symoff, const double* __restrict__ sym_n,
             double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  int lane = (int)(p & 31);
  long long base = kbase[p >> 5];
  int kmax = (int)((kbase[(p >> 5)+1] - base) >> 5);
  if (mode == 1) {
    double r[NC], pp[NC], pn[NC];
    for (int c=0; c<NC; ++c) { r[c] = 0.0; pp[c] = 0.0; pn[c] = 0.0; }
    for (int k=0; k<kmax; ++k) {
      int ta = __ldg( kinc + base + (long long)k*32 + lane );
      if (ta < 0) continue;
      size_t e = (size_t)(ta >> 2); int a = ta & 3;
      for (int c=0; c<NC; ++c) {
        r[c] += __ldg( T + (size_t)(a*NC+c)*ntet + e );
        double aec = __ldg( T + (size_t)(20+c*4+a)*ntet + e ); pp[c] += fmax( 0.0, aec ); pn[c] += fmin( 0.0, aec );
      }
    }
    zal_symp( p, pp, pn, bcof, symoff, sym_n );
    for (int c=0; c<NC; ++c) { part[(size_t)i*15+c] = r[c]; part[(size_t)i*15+5+c] = pp[c]; part[(size_t)i*15+10+c] = pn[c]; }
  } else if (mode == 2) {
    double qa[NC], qb[NC];
    for (int c=0; c<NC; ++c) { qa[c] = -1.7976931348623157e308; qb[c] = 1.7976931348623157e308; }
    for (int k=0; k<kmax; ++k) {
      int ta = __ldg( kinc + base + (long long)k*32 + lane );
      if (ta < 0) continue;
      size_t e = (size_t)(ta >> 2);
      for (int c=0; c<NC; ++c) {
        qa[c] = fmax( qa[c], __ldg( T + (size_t)(2*c)*ntet + e ) );
        qb[c] = fmin( qb[c], __ldg( T + (size_t)(2*c+1)*ntet + e ) );
      }
    }
    for (int c=0; c<NC; ++c) { part[(size_t)i*10+2*c] = qa[c]; part[(size_t)i*10+2*c+1] = qb[c]; }
  } else {
    double a_[NC];
    for (int c=0; c<NC; ++c) a_[c] = 0.0;
    for (int k=0; k<kmax; ++k) {
      int ta = __ldg( kinc + base + (long long)k*32 + lane );
      if (ta < 0) continue;
      size_t e = (size_t)(ta >> 2); int a = ta & 3;
      for (int c=0; c<NC; ++c) a_[c] += __ldg( T + (size_t)(a*NC+c)*ntet + e );
    }
    for (int c=0; c<NC; ++c) part[(size_t)i*NC+c] = a_[c];
  }
}

// finish of pass 1 at the shared nodes: R, P /= vol, ul = u + dt R/vol - P+ - P- (KozCG's sign, :897)
__global__ void k_koz_fin1( int nsh, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
             const int* __restrict__ ridx, const double* __restrict__ part, const double* __restrict__ recvbuf,
             const double* __restrict__ U, const double* __restrict__ vol, double dt, int fct,
             double* __restrict__ P, double* __restrict__ UL, double* __restrict__ R )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  double t[15];
  for (int j=0; j<15; ++j) {
    double a = part[(size_t)i*15+j];
    for (int r=roff[i]; r<roff[i+1]; ++r) a += recvbuf[(size_t)ridx[r]*15+j];
    t[j] = a;
  }
  for (int c=0; c<NC; ++c) R[p*NC+c] = t[c];
  if (!fct) return;
  double ivp = 1.0 / vol[p];
  for (int c=0; c<NC; ++c) {
    double pp = t[5+c]*ivp, pn = t[10+c]*ivp;
    P[(2*c)*NP+p] = pp; P[(2*c+1)*NP+p] = pn;
    UL[c*NP+p] = U[c*NP+p] + dt*t[c]*ivp - pp - pn;
  }
}

// fct = false: u = u + dt R/vol (KozCG.cpp:1150-1157)
__global__ void k_koz_nofct( size_t npoin, size_t NP, const double* __restrict__ R, const double* __restrict__ vol,
                             const double* __restrict__ U, double dt, double* __restrict__ Unew, double* __restrict__ W )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  double ivp = 1.0 / vol[p], u[NC], w[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { u[c] = U[c*NP+p] + dt*R[p*NC+c]*ivp; Unew[c*NP+p] = u[c]; }
  primitive( u, w );
  store_w( W, NP, p, w );
}

// one scalar on several partitions: the shared nodes' own sums of a pass from the per-tet buffer (T rows as
// written by k_koz_selem1 / k_koz_elem2<1> / k_koz_elem3<1>) -> part[i][3|2|1]; finished by k_fct_sfin
__global__ void k_koz_ssh( int mode, int nsh, size_t ntet, const int* __restrict__ sh_node,
              const long long* __restrict__ kbase, const int* __restrict__ kinc, const double* __restrict__ T,
              double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  int lane = (int)(p & 31);
  long long base = kbase[p >> 5];
  int kmax = (int)((kbase[(p >> 5)+1] - base) >> 5);
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  if (mode == 2) { v0 = -1.7976931348623157e308; v1 = 1.7976931348623157e308; }
  for (int k=0; k<kmax; ++k) {
    int ta = __ldg( kinc + base + (long long)k*32 + lane );
    if (ta < 0) continue;
    size_t e = (size_t)(ta >> 2); int a = ta & 3;
    if (mode == 1) {
      v0 += __ldg( T + (size_t)a*ntet + e );
      double aec = __ldg( T + (size_t)(4+a)*ntet + e ); v1 += fmax( 0.0, aec ); v2 += fmin( 0.0, aec );
    } else if (mode == 2) {
      v0 = fmax( v0, __ldg( T + e ) ); v1 = fmin( v1, __ldg( T + ntet + e ) );
    } else
      v0 += __ldg( T + (size_t)a*ntet + e );
  }
  if (mode == 1) { part[(size_t)i*3] = v0; part[(size_t)i*3+1] = v1; part[(size_t)i*3+2] = v2; }
  else if (mode == 2) { part[(size_t)i*2] = v0; part[(size_t)i*2+1] = v1; }
  else part[i] = v0;
}

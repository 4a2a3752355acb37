// xyst_b200/csrc/zalcg_kernels.cuh -- ZalCG device code: Taylor-Galerkin edge flux and the flux-corrected-transport node gathers
// Part of the single translation unit xyst_b200.cu (included inside its anonymous namespace).

// ---------------------------------------------------------------------------------
// ZalCG: Taylor-Galerkin two-step edge flux (Zalesak.cpp:31-200, no source term) and the
// flux-corrected-transport passes of ZalCG.cpp:1056-1607 as node gathers
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_zal_flux_edge( size_t nslot, size_t NP, const int* __restrict__ ep, const int* __restrict__ eq,
                 const double* __restrict__ D, const double* __restrict__ U, const double* __restrict__ X,
                 double dt, const double* __restrict__ dtp, DParams P, double* __restrict__ F,
                 const double* __restrict__ Sn, int ns, const double* __restrict__ sU, double* __restrict__ sF )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= nslot) return;
  int pi = ep[e];
  if (pi < 0) return;
  size_t p = pi, q = eq[e];
  if (dtp) dt = (dtp[p] + dtp[q])/2.0;      // steady state: local time step of the edge (Zalesak.cpp:107)
  double g = P.gamma;
  double dx = X[p] - X[q], dy = X[NP+p] - X[NP+q], dz = X[2*NP+p] - X[2*NP+q];
  double dl = dx*dx + dy*dy + dz*dz;
  dx /= dl; dy /= dl; dz /= dl;
  double rL = U[p], ruL = U[NP+p], rvL = U[2*NP+p], rwL = U[3*NP+p], reL = U[4*NP+p];
  double pL = (reL - 0.5*(ruL*ruL + rvL*rvL + rwL*rwL)/rL) * (g-1.0);
  double dnL = (ruL*dx + rvL*dy + rwL*dz)/rL;
  double rR = U[q], ruR = U[NP+q], rvR = U[2*NP+q], rwR = U[3*NP+q], reR = U[4*NP+q];
  double pR = (reR - 0.5*(ruR*ruR + rvR*rvR + rwR*rwR)/rR) * (g-1.0);
  double dnR = (ruR*dx + rvR*dy + rwR*dz)/rR;
  double nx = D[e], ny = D[nslot+e], nz = D[2*nslot+e];
  double dp = pL - pR;
  double rh  = 0.5*(rL + rR - dt*(rL*dnL - rR*dnR));
  double ruh = 0.5*(ruL + ruR - dt*(ruL*dnL - ruR*dnR + dp*dx));
  double rvh = 0.5*(rvL + rvR - dt*(rvL*dnL - rvR*dnR + dp*dy));
  double rwh = 0.5*(rwL + rwR - dt*(rwL*dnL - rwR*dnR + dp*dz));
  double reh = 0.5*(reL + reR - dt*((reL+pL)*dnL - (reR+pR)*dnR));
  if (Sn) {                                 // source at the end nodes into the half step (Zalesak.cpp:118-128)
    double coef = dt/4.0;
    rh  += coef*(Sn[p*NC+0] + Sn[q*NC+0]);
    ruh += coef*(Sn[p*NC+1] + Sn[q*NC+1]);
    rvh += coef*(Sn[p*NC+2] + Sn[q*NC+2]);
    rwh += coef*(Sn[p*NC+3] + Sn[q*NC+3]);
    reh += coef*(Sn[p*NC+4] + Sn[q*NC+4]);
  }
  double ph = (reh - 0.5*(ruh*ruh + rvh*rvh + rwh*rwh)/rh) * (g-1.0);
  double vn = (ruh*nx + rvh*ny + rwh*nz)/rh;
  double f[NC];
  f[0] = 2.0*rh*vn;
  f[1] = 2.0*(ruh*vn + ph*nx);
  f[2] = 2.0*(rvh*vn + ph*ny);
  f[3] = 2.0*(rwh*vn + ph*nz);
  f[4] = 2.0*(reh + ph)*vn;
  double fw = 0.0;
  if (P.stab2) {
    double vnL = (ruL*nx + rvL*ny + rwL*nz)/rL;
    double vnR = (ruR*nx + rvR*ny + rwR*nz)/rR;
    double len = sqrt( nx*nx + ny*ny + nz*nz );
    double cL = sqrt( g * fmax(pL,0.0) / fmax(rL,1.0e-8) );
    double cR = sqrt( g * fmax(pR,0.0) / fmax(rR,1.0e-8) );
    fw = P.stab2coef * fmax( fabs(vnL) + cL*len, fabs(vnR) + cR*len );
    f[0] -= fw*(rL - rR); f[1] -= fw*(ruL - ruR); f[2] -= fw*(rvL - rvR);
    f[3] -= fw*(rwL - rwR); f[4] -= fw*(reL - reR);
  }
  store_f( F, nslot, e, f );
  // transported scalars (:113-116,146-149,193-196; their source columns are zero): sF[k][e]
  for (int k=0; k<ns; ++k) {
    double sL = sU[(size_t)k*NP+p], sR = sU[(size_t)k*NP+q];
    double ue = 0.5*(sL + sR - dt*(sL*dnL - sR*dnR));
    double fs = 2.0*ue*vn;
    if (P.stab2) fs -= fw*(sL - sR);
    sF[(size_t)k*nslot+e] = fs;
  }
}

// pass 1 (aec + first half of alw): R = sum +-F + boundary; P+/- from the antidiffusive edge
// contributions aec = -dif*ctau*(u_first - u_second) (ZalCG.cpp:1071-1115), symmetry BC on P
// (:1117-1133), then P /= vol and the low-order solution ul = u - dt R/vol - P+ - P- (:1195-1204)
// raw sums of pass 1 at node p: r = sum +-F (+ boundary), P+/- before the symmetry BC and the division by vol
__device__ __forceinline__ void zal_sum1( size_t p, int lane, long long base, int kmax, size_t NP,
    const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot, const double* __restrict__ F,
    const double* __restrict__ U, const int* __restrict__ bslot, const double* __restrict__ Rb, double ctau,
    double up[NC], double r[NC], double pp[NC], double pn[NC], const double* __restrict__ Se = nullptr )
{
  #pragma unroll
  for (int c=0; c<NC; ++c) { up[c] = U[c*NP+p]; r[c] = 0.0; pp[c] = 0.0; pn[c] = 0.0; }
  // padding entries (se = 0) point at the node itself and at slot 0 with weight 0: no branch
  #pragma unroll kZalUnroll
  for (int k=0; k<kmax; ++k) {
    long long i = base + (long long)k*32 + lane;
    int2 eq = __ldg( inc_eq + i );
    const int se = eq.x;
    size_t q = (size_t)eq.y;
    size_t sl = se == 0 ? 0 : (size_t)(abs(se)-1);
    double dif = se == 0 ? 0.0 : __ldg( D + 3*nslot + sl );
    double fl[NC];
    load_f( F, nslot, sl, fl );
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double f = se == 0 ? 0.0 : fl[c];
      double uq = __ldg( U + c*NP + q );
      if (se < 0) {                       // this node is the edge's first node
        r[c] -= f;
        double aec = -dif * ctau * (up[c] - uq);
        if (aec > 0.0) pn[c] -= aec; else pp[c] -= aec;
      } else {                            // second node
        r[c] += f;
        double aec = -dif * ctau * (uq - up[c]);
        if (aec > 0.0) pp[c] += aec; else pn[c] += aec;
      }
      // source at the edge's midpoint, to both of its nodes (Zalesak.cpp:152-163, advdom :243-249,...)
      if (Se && se != 0) r[c] += (-5.0/3.0*dif) * __ldg( Se + (size_t)c*nslot + sl );
    }
  }
  int b = bslot[p];
  if (b >= 0) {
    #pragma unroll
    for (int c=0; c<NC; ++c) r[c] += Rb[(size_t)b*NC+c];
  }
}

// symmetry BC on the (own) antidiffusive sums P (:1117-1133)
__device__ __forceinline__ void zal_symp( size_t p, double pp[NC], double pn[NC], const int* __restrict__ bcof,
    const int* __restrict__ symoff, const double* __restrict__ sym_n )
{
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
}

// P /= vol, low-order solution ul = u - dt R/vol - P+ - P- (:1195-1204)
__device__ __forceinline__ void zal_fin1( size_t p, size_t NP, const double up[NC], const double r[NC], double pp[NC], double pn[NC],
    const double* __restrict__ vol, double dt, double* __restrict__ P, double* __restrict__ UL, double* __restrict__ R )
{
  // one quotient per node, then products (the reference divides each component: same values up to one rounding)
  double ivp = 1.0 / vol[p];
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    pp[c] *= ivp; pn[c] *= ivp;
    P[(2*c)*NP+p] = pp[c]; P[(2*c+1)*NP+p] = pn[c];
    UL[c*NP+p] = up[c] - dt*r[c]*ivp - pp[c] - pn[c];
    R[p*NC+c] = r[c];
  }
}

__global__ void __launch_bounds__(NODE_THREADS, 3)
k_zal_node1( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot,
             const double* __restrict__ F, const double* __restrict__ U, const int* __restrict__ bslot,
             const double* __restrict__ Rb, const int* __restrict__ bcof, const int* __restrict__ symoff,
             const double* __restrict__ sym_n, const double* __restrict__ vol, double dt, const double* __restrict__ dtp,
             double ctau, int fct, double* __restrict__ P, double* __restrict__ UL, double* __restrict__ R,
             const double* __restrict__ Se )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  if (dtp) dt = dtp[p];                     // steady state (ZalCG.cpp:1195)
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double up[NC], r[NC], pp[NC], pn[NC];
  zal_sum1( p, lane, base, kmax, NP, inc_eq, D, nslot, F, U, bslot, Rb, ctau, up, r, pp, pn, Se );
  if (!fct) {
    #pragma unroll
    for (int c=0; c<NC; ++c) R[p*NC+c] = r[c];
    return;
  }
  zal_symp( p, pp, pn, bcof, symoff, sym_n );
  zal_fin1( p, NP, up, r, pp, pn, vol, dt, P, UL, R );
}

// Several partitions (ZalCG::comrhs/comaec, ZalCG.cpp:1023-1053,1139-1148): the raw sums of the nodes
// shared with other partitions -> part[i][15] = (r, P+, P-), exchanged and summed, then finished
__global__ void k_zal_sh1( int nsh, size_t NP, const int* __restrict__ sh_node, const long long* __restrict__ sl_base,
             const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot, const double* __restrict__ F,
             const double* __restrict__ U, const int* __restrict__ bslot, const double* __restrict__ Rb, double ctau,
             const int* __restrict__ bcof, const int* __restrict__ symoff, const double* __restrict__ sym_n,
             double* __restrict__ part, const double* __restrict__ Se )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  long long base = sl_base[p >> 5];
  int kmax = (int)((sl_base[(p >> 5)+1] - base) >> 5);
  double up[NC], r[NC], pp[NC], pn[NC];
  zal_sum1( p, (int)(p & 31), base, kmax, NP, inc_eq, D, nslot, F, U, bslot, Rb, ctau, up, r, pp, pn, Se );
  zal_symp( p, pp, pn, bcof, symoff, sym_n );       // on the own sums, before they travel (as the reference)
  for (int c=0; c<NC; ++c) { part[(size_t)i*15+c] = r[c]; part[(size_t)i*15+5+c] = pp[c]; part[(size_t)i*15+10+c] = pn[c]; }
}

__global__ void k_zal_fin1( int nsh, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
             const int* __restrict__ ridx, const double* __restrict__ part, const double* __restrict__ recvbuf,
             const double* __restrict__ U, const double* __restrict__ vol, double dt, const double* __restrict__ dtp,
             int fct, double* __restrict__ P, double* __restrict__ UL, double* __restrict__ R )
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
  double up[NC];
  for (int c=0; c<NC; ++c) up[c] = U[c*NP+p];
  if (!fct) { for (int c=0; c<NC; ++c) R[p*NC+c] = t[c]; return; }
  if (dtp) dt = dtp[p];
  zal_fin1( p, NP, up, t, t+5, t+10, vol, dt, P, UL, R );
}

// pass 2 (second half of alw + first half of lim): allowed bounds Q+/- over the edge
// neighbours (:1206-1290), Q -= ul, limit coefficients C+/- (:1361-1380) -> Q
// raw allowed bounds of node p over its edge neighbours
__device__ __forceinline__ void zal_sum2( size_t p, int lane, long long base, int kmax, size_t NP,
    const int2* __restrict__ inc_eq, const double* __restrict__ U, const double* __restrict__ UL, int clip,
    double qa[NC], double qb[NC] )
{
  double hp[NC], lp[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    double ulp = UL[c*NP+p], u = U[c*NP+p];
    hp[c] = clip ? ulp : fmax( ulp, u );
    lp[c] = clip ? ulp : fmin( ulp, u );
    qa[c] = -1.7976931348623157e308; qb[c] = 1.7976931348623157e308;
  }
  #pragma unroll kZalUnroll
  for (int k=0; k<kmax; ++k) {                 // padding entries point at the node itself: no effect on the bounds
    long long i = base + (long long)k*32 + lane;
    size_t q = (size_t)__ldg( inc_eq + i ).y;
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double ulq = __ldg( UL + c*NP + q );
      double hq = ulq, lq = ulq;
      if (!clip) { double uq = __ldg( U + c*NP + q ); hq = fmax( ulq, uq ); lq = fmin( ulq, uq ); }
      qa[c] = fmax( qa[c], fmax( hp[c], hq ) );
      qb[c] = fmin( qb[c], fmin( lp[c], lq ) );
    }
  }
}

// Q -= ul, limit coefficients C+/- (:1361-1380) -> Q
__device__ __forceinline__ void zal_fin2( size_t p, size_t NP, const double qa[NC], const double qb[NC],
    const double* __restrict__ UL, const double* __restrict__ P, double* __restrict__ Q )
{
  const double eps = 2.220446049250313e-16;
  #pragma unroll
  for (int c=0; c<NC; ++c) {
    double ulp = UL[c*NP+p];
    double a = qa[c] - ulp, b = qb[c] - ulp;
    double pa = P[(2*c)*NP+p], pb = P[(2*c+1)*NP+p];
    Q[(2*c)*NP+p]   = pa <  eps ? 0.0 : fmin( 1.0, a/pa );
    Q[(2*c+1)*NP+p] = pb > -eps ? 0.0 : fmin( 1.0, b/pb );
  }
}

__global__ void __launch_bounds__(NODE_THREADS, 3)
k_zal_node2( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq, const double* __restrict__ U, const double* __restrict__ UL,
             const double* __restrict__ P, int clip, double* __restrict__ Q )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double qa[NC], qb[NC];
  zal_sum2( p, lane, base, kmax, NP, inc_eq, U, UL, clip, qa, qb );
  zal_fin2( p, NP, qa, qb, UL, P, Q );
}

// several partitions (ZalCG::comalw, ZalCG.cpp:1297-1333): own bounds of the shared nodes ->
// part[i][10] = (Q+ max, Q- min) per component, combined with max / min over the sharers
__global__ void k_zal_sh2( int nsh, size_t NP, const int* __restrict__ sh_node, const long long* __restrict__ sl_base,
             const int2* __restrict__ inc_eq, const double* __restrict__ U, const double* __restrict__ UL, int clip,
             double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  long long base = sl_base[p >> 5];
  int kmax = (int)((sl_base[(p >> 5)+1] - base) >> 5);
  double qa[NC], qb[NC];
  zal_sum2( p, (int)(p & 31), base, kmax, NP, inc_eq, U, UL, clip, qa, qb );
  for (int c=0; c<NC; ++c) { part[(size_t)i*10+2*c] = qa[c]; part[(size_t)i*10+2*c+1] = qb[c]; }
}

__global__ void k_zal_fin2( int nsh, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
             const int* __restrict__ ridx, const double* __restrict__ part, const double* __restrict__ recvbuf,
             const double* __restrict__ UL, const double* __restrict__ P, double* __restrict__ Q )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  double qa[NC], qb[NC];
  for (int c=0; c<NC; ++c) {
    double a = part[(size_t)i*10+2*c], b = part[(size_t)i*10+2*c+1];
    for (int r=roff[i]; r<roff[i+1]; ++r) { a = fmax( a, recvbuf[(size_t)ridx[r]*10+2*c] ); b = fmin( b, recvbuf[(size_t)ridx[r]*10+2*c+1] ); }
    qa[c] = a; qb[c] = b;
  }
  zal_fin2( p, NP, qa, qb, UL, P, Q );
}

// pass 3 (second half of lim + solve): limited antidiffusive contributions (:1382-1481) and
// u = ul + a/vol (:1552-1557)
// limited antidiffusive sums of node p
__device__ __forceinline__ void zal_sum3( size_t p, int lane, long long base, int kmax, size_t NP,
    const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot, const double* __restrict__ U,
    const double* __restrict__ Q, double ctau, int sysmask, double a[NC] )
{
  double up[NC], cpa[NC], cpb[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { up[c] = U[c*NP+p]; cpa[c] = Q[(2*c)*NP+p]; cpb[c] = Q[(2*c+1)*NP+p]; a[c] = 0.0; }
  #pragma unroll kZalUnroll
  for (int k=0; k<kmax; ++k) {                 // padding entries: dif = 0, neighbour = the node itself
    long long i = base + (long long)k*32 + lane;
    int2 eq = __ldg( inc_eq + i );
    const int se = eq.x;
    size_t q = (size_t)eq.y;
    double dif = se == 0 ? 0.0 : __ldg( D + 3*nslot + (size_t)(abs(se)-1) );
    double aec[NC], coef[NC];
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      double uq = __ldg( U + c*NP + q );
      double cqa = __ldg( Q + (2*c)*NP + q ), cqb = __ldg( Q + (2*c+1)*NP + q );
      if (se < 0) {      // first = this node, second = q
        aec[c] = -dif * ctau * (up[c] - uq);
        coef[c] = fmin( aec[c] < 0.0 ? cpa[c] : cpb[c], aec[c] > 0.0 ? cqa : cqb );
      } else {           // first = q, second = this node
        aec[c] = -dif * ctau * (uq - up[c]);
        coef[c] = fmin( aec[c] < 0.0 ? cqa : cqb, aec[c] > 0.0 ? cpa[c] : cpb[c] );
      }
    }
    double cs = 1.0;
    #pragma unroll
    for (int c=0; c<NC; ++c) if (sysmask & (1<<c)) cs = fmin( cs, coef[c] );
    #pragma unroll
    for (int c=0; c<NC; ++c) {
      if (sysmask & (1<<c)) coef[c] = cs;
      double v = aec[c] * coef[c];
      if (se < 0) a[c] -= v; else a[c] += v;
    }
  }
}

__device__ __forceinline__ void zal_fin3( size_t p, size_t NP, const double a[NC], const double* __restrict__ UL,
    const double* __restrict__ vol, double* __restrict__ Unew, double* __restrict__ W )
{
  double ivp = 1.0 / vol[p], u[NC], w[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { u[c] = UL[c*NP+p] + a[c]*ivp; Unew[c*NP+p] = u[c]; }
  primitive( u, w );
  store_w( W, NP, p, w );
}

__global__ void __launch_bounds__(NODE_THREADS, 3)
k_zal_node3( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot,
             const double* __restrict__ U, const double* __restrict__ UL, const double* __restrict__ Q,
             const double* __restrict__ vol, double ctau, int sysmask, double* __restrict__ Unew,
             double* __restrict__ W )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double a[NC];
  zal_sum3( p, lane, base, kmax, NP, inc_eq, D, nslot, U, Q, ctau, sysmask, a );
  zal_fin3( p, NP, a, UL, vol, Unew, W );
}

// several partitions (ZalCG::comlim, ZalCG.cpp:1490-1499): own limited sums of the shared nodes -> part[i][5]
__global__ void k_zal_sh3( int nsh, size_t NP, const int* __restrict__ sh_node, const long long* __restrict__ sl_base,
             const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot, const double* __restrict__ U,
             const double* __restrict__ Q, double ctau, int sysmask, double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  long long base = sl_base[p >> 5];
  int kmax = (int)((sl_base[(p >> 5)+1] - base) >> 5);
  double a[NC];
  zal_sum3( p, (int)(p & 31), base, kmax, NP, inc_eq, D, nslot, U, Q, ctau, sysmask, a );
  for (int c=0; c<NC; ++c) part[(size_t)i*NC+c] = a[c];
}

__global__ void k_zal_fin3( int nsh, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
             const int* __restrict__ ridx, const double* __restrict__ part, const double* __restrict__ recvbuf,
             const double* __restrict__ UL, const double* __restrict__ vol, double* __restrict__ Unew, double* __restrict__ W )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  double a[NC];
  for (int c=0; c<NC; ++c) {
    double t = part[(size_t)i*NC+c];
    for (int r=roff[i]; r<roff[i+1]; ++r) t += recvbuf[(size_t)ridx[r]*NC+c];
    a[c] = t;
  }
  zal_fin3( p, NP, a, UL, vol, Unew, W );
}

// fct = false: u = u - dt R/vol (ZalCG.cpp:1560-1567)
__global__ void k_zal_nofct( size_t npoin, size_t NP, const double* __restrict__ R, const double* __restrict__ vol,
                             const double* __restrict__ U, double dt, const double* __restrict__ dtp,
                             double* __restrict__ Unew, double* __restrict__ W )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  if (dtp) dt = dtp[p];                     // steady state (:1563)
  double ivp = 1.0 / vol[p], u[NC], w[NC];
  #pragma unroll
  for (int c=0; c<NC; ++c) { u[c] = U[c*NP+p] - dt*R[p*NC+c]*ivp; Unew[c*NP+p] = u[c]; }
  primitive( u, w );
  store_w( W, NP, p, w );
}


// ---- transported scalars in ZalCG: the three FCT node passes for ONE scalar (rows sv, UL, P[2], Q[2] of that
// ---- scalar; sF its edge fluxes). Same arithmetic as a flow component, no symmetry BC.
__global__ void __launch_bounds__(NODE_THREADS, 3)
k_zal_snode1( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq,
              const double* __restrict__ D, size_t nslot, const double* __restrict__ sF, const double* __restrict__ sv,
              const int* __restrict__ bslot, const double* __restrict__ sRb, int ns, int ks,
              const double* __restrict__ vol, double dt, const double* __restrict__ dtp, double ctau, int fct,
              double* __restrict__ P, double* __restrict__ UL )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  if (dtp) dt = dtp[p];
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double up = sv[p], r = 0.0, pp = 0.0, pn = 0.0;
  for (int k=0; k<kmax; ++k) {
    int2 eq = __ldg( inc_eq + base + (long long)k*32 + lane );
    const int se = eq.x;
    if (se == 0) continue;
    size_t q = (size_t)eq.y, sl = (size_t)(abs(se)-1);
    double dif = __ldg( D + 3*nslot + sl ), f = __ldg( sF + sl ), uq = __ldg( sv + q );
    if (se < 0) {
      r -= f;
      double aec = -dif * ctau * (up - uq);
      if (aec > 0.0) pn -= aec; else pp -= aec;
    } else {
      r += f;
      double aec = -dif * ctau * (uq - up);
      if (aec > 0.0) pp += aec; else pn += aec;
    }
  }
  int b = bslot[p];
  if (b >= 0) r += sRb[(size_t)b*ns + ks];
  double ivp = 1.0 / vol[p];
  if (!fct) { UL[p] = up - dt*r*ivp; return; }          // u = u - dt R/vol (ZalCG.cpp:1560-1567)
  pp *= ivp; pn *= ivp;
  P[p] = pp; P[NP+p] = pn;
  UL[p] = up - dt*r*ivp - pp - pn;
}

__global__ void __launch_bounds__(NODE_THREADS, 3)
k_zal_snode2( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq,
              const double* __restrict__ sv, const double* __restrict__ UL, const double* __restrict__ P, int clip,
              double* __restrict__ Q )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double ulp = UL[p], u = sv[p];
  double hp = clip ? ulp : fmax( ulp, u ), lp = clip ? ulp : fmin( ulp, u );
  double qa = -1.7976931348623157e308, qb = 1.7976931348623157e308;
  for (int k=0; k<kmax; ++k) {                 // padding entries point at the node itself: no effect on the bounds
    size_t q = (size_t)__ldg( inc_eq + base + (long long)k*32 + lane ).y;
    double ulq = __ldg( UL + q );
    double hq = ulq, lq = ulq;
    if (!clip) { double uq = __ldg( sv + q ); hq = fmax( ulq, uq ); lq = fmin( ulq, uq ); }
    qa = fmax( qa, fmax( hp, hq ) );
    qb = fmin( qb, fmin( lp, lq ) );
  }
  const double eps = 2.220446049250313e-16;
  double a = qa - ulp, b = qb - ulp;
  double pa = P[p], pb = P[NP+p];
  Q[p]    = pa <  eps ? 0.0 : fmin( 1.0, a/pa );
  Q[NP+p] = pb > -eps ? 0.0 : fmin( 1.0, b/pb );
}

__global__ void __launch_bounds__(NODE_THREADS, 3)
k_zal_snode3( size_t npoin, size_t NP, const long long* __restrict__ sl_base, const int2* __restrict__ inc_eq,
              const double* __restrict__ D, size_t nslot, const double* __restrict__ sv, const double* __restrict__ UL,
              const double* __restrict__ Q, const double* __restrict__ vol, double ctau, double* __restrict__ Unew )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  double up = sv[p], cpa = Q[p], cpb = Q[NP+p], a = 0.0;
  for (int k=0; k<kmax; ++k) {
    int2 eq = __ldg( inc_eq + base + (long long)k*32 + lane );
    const int se = eq.x;
    if (se == 0) continue;
    size_t q = (size_t)eq.y;
    double dif = __ldg( D + 3*nslot + (size_t)(abs(se)-1) );
    double uq = __ldg( sv + q ), cqa = __ldg( Q + q ), cqb = __ldg( Q + NP + q );
    if (se < 0) {      // first = this node, second = q
      double aec = -dif * ctau * (up - uq);
      double coef = fmin( aec < 0.0 ? cpa : cpb, aec > 0.0 ? cqa : cqb );
      a -= aec * coef;
    } else {           // first = q, second = this node
      double aec = -dif * ctau * (uq - up);
      double coef = fmin( aec < 0.0 ? cqa : cqb, aec > 0.0 ? cpa : cpb );
      a += aec * coef;
    }
  }
  Unew[p] = UL[p] + a / vol[p];
}

// ---- one scalar on several partitions (ZalCG::comrhs/comaec, comalw, comlim per component): the shared nodes'
// ---- OWN sums of a pass (mode 1: r, P+, P-; 2: bounds; 3: limited sum) -> part[i][3|2|1]; after the exchange
// ---- the nodes are finished from the complete values by k_fct_sfin (shared with KozCG)
__global__ void k_zal_ssh( int mode, int nsh, size_t NP, const int* __restrict__ sh_node, const long long* __restrict__ sl_base,
              const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot, const double* __restrict__ sF,
              const double* __restrict__ sv, const double* __restrict__ UL, const double* __restrict__ Q,
              const int* __restrict__ bslot, const double* __restrict__ sRb, int ns, int ks, double ctau, int clip,
              double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  int lane = (int)(p & 31);
  long long base = sl_base[p >> 5];
  int kmax = (int)((sl_base[(p >> 5)+1] - base) >> 5);
  double up = sv[p];
  if (mode == 1) {
    double r = 0.0, pp = 0.0, pn = 0.0;
    for (int k=0; k<kmax; ++k) {
      int2 eq = __ldg( inc_eq + base + (long long)k*32 + lane );
      const int se = eq.x;
      if (se == 0) continue;
      size_t q = (size_t)eq.y, sl = (size_t)(abs(se)-1);
      double dif = __ldg( D + 3*nslot + sl ), f = __ldg( sF + sl ), uq = __ldg( sv + q );
      if (se < 0) { r -= f; double aec = -dif * ctau * (up - uq); if (aec > 0.0) pn -= aec; else pp -= aec; }
      else        { r += f; double aec = -dif * ctau * (uq - up); if (aec > 0.0) pp += aec; else pn += aec; }
    }
    int b = bslot[p];
    if (b >= 0) r += sRb[(size_t)b*ns + ks];
    part[(size_t)i*3] = r; part[(size_t)i*3+1] = pp; part[(size_t)i*3+2] = pn;
  } else if (mode == 2) {
    double ulp = UL[p];
    double hp = clip ? ulp : fmax( ulp, up ), lp = clip ? ulp : fmin( ulp, up );
    double qa = -1.7976931348623157e308, qb = 1.7976931348623157e308;
    for (int k=0; k<kmax; ++k) {
      size_t q = (size_t)__ldg( inc_eq + base + (long long)k*32 + lane ).y;
      double ulq = __ldg( UL + q );
      double hq = ulq, lq = ulq;
      if (!clip) { double uq = __ldg( sv + q ); hq = fmax( ulq, uq ); lq = fmin( ulq, uq ); }
      qa = fmax( qa, fmax( hp, hq ) );
      qb = fmin( qb, fmin( lp, lq ) );
    }
    part[(size_t)i*2] = qa; part[(size_t)i*2+1] = qb;
  } else {
    double cpa = Q[p], cpb = Q[NP+p], a = 0.0;
    for (int k=0; k<kmax; ++k) {
      int2 eq = __ldg( inc_eq + base + (long long)k*32 + lane );
      const int se = eq.x;
      if (se == 0) continue;
      size_t q = (size_t)eq.y;
      double dif = __ldg( D + 3*nslot + (size_t)(abs(se)-1) );
      double uq = __ldg( sv + q ), cqa = __ldg( Q + q ), cqb = __ldg( Q + NP + q );
      if (se < 0) { double aec = -dif * ctau * (up - uq); a -= aec * fmin( aec < 0.0 ? cpa : cpb, aec > 0.0 ? cqa : cqb ); }
      else        { double aec = -dif * ctau * (uq - up); a += aec * fmin( aec < 0.0 ? cqa : cqb, aec > 0.0 ? cpa : cpb ); }
    }
    part[i] = a;
  }
}

// finish of a pass of one scalar at the shared nodes from own + received parts; sign = -1 (ZalCG: ul = u - dt r/vol)
// or +1 (KozCG: ul = u + dt r/vol). mode 1 writes P, UL (or, without FCT, the new value into UL); 2: Q; 3: Unew.
__global__ void k_fct_sfin( int mode, int nsh, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
              const int* __restrict__ ridx, const double* __restrict__ part, const double* __restrict__ recvbuf,
              const double* __restrict__ sv, const double* __restrict__ vol, double sdt, const double* __restrict__ dtp,
              int fct, double* __restrict__ P, double* __restrict__ UL, double* __restrict__ Q, double* __restrict__ Unew )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  if (mode == 1) {
    double t[3];
    for (int j=0; j<3; ++j) { double a = part[(size_t)i*3+j]; for (int r=roff[i]; r<roff[i+1]; ++r) a += recvbuf[(size_t)ridx[r]*3+j]; t[j] = a; }
    if (dtp) sdt = sdt < 0.0 ? -dtp[p] : dtp[p];
    double ivp = 1.0 / vol[p];
    if (!fct) { UL[p] = sv[p] + sdt*t[0]*ivp; return; }
    double pp = t[1]*ivp, pn = t[2]*ivp;
    P[p] = pp; P[NP+p] = pn;
    UL[p] = sv[p] + sdt*t[0]*ivp - pp - pn;
  } else if (mode == 2) {
    double a = part[(size_t)i*2], b = part[(size_t)i*2+1];
    for (int r=roff[i]; r<roff[i+1]; ++r) { a = fmax( a, recvbuf[(size_t)ridx[r]*2] ); b = fmin( b, recvbuf[(size_t)ridx[r]*2+1] ); }
    const double eps = 2.220446049250313e-16;
    double ulp = UL[p];
    a -= ulp; b -= ulp;
    double pa = P[p], pb = P[NP+p];
    Q[p]    = pa <  eps ? 0.0 : fmin( 1.0, a/pa );
    Q[NP+p] = pb > -eps ? 0.0 : fmin( 1.0, b/pb );
  } else {
    double a = part[i];
    for (int r=roff[i]; r<roff[i+1]; ++r) a += recvbuf[ridx[r]];
    Unew[p] = UL[p] + a / vol[p];
  }
}

// xyst_b200/csrc/riecg_scalar.cuh -- transported scalars of RieCG (ncomp = 5 + ns)
// Part of the single translation unit xyst_b200.cu (included inside its anonymous namespace).
//
// The reference carries passive scalars behind the five flow variables of tk::Fields: same gradient
// operator (src/Physics/Riemann.cpp:229-367, primitive of a scalar = its value, :211-227), their own
// MUSCL reconstruction without the positivity fallback (:145-209), and a scalar Riemann flux that
// upwinds with the normal velocities of the RECONSTRUCTED flow states of the same edge (rusanov
// :462-470, hllc :636-643), boundary flux u_c * vn (advbnd :840,852,864). The flow kernels stay
// specialised on five components; the scalars ride on top of them:
//   k_flux_own   leaves per edge, reference-oriented, the three numbers the scalar flux needs of the
//                flow: (cL, cR, sw) with  f_c = l_c cL + r_c cR + sw (r_c - l_c)
//   k_scal_bnd   boundary-face parts of gradient and flux per boundary node
//   k_scal_grad  nodal gradients of the scalars (gather over the incidence lists)
//   k_scal_flux  per edge: scalar MUSCL + flux
//   k_scal_node  nodal sums (+ boundary, source) and the RK update, or the rhs columns
// Plain structure-of-arrays with stride NP; these are not the benchmarked kernels.

__global__ void k_scal_bnd( int nbn, int ns, size_t NP, const int* __restrict__ bn_off, const int* __restrict__ bn_face,
                            const int* __restrict__ tri, const unsigned char* __restrict__ besym,
                            const double* __restrict__ fn, const double* __restrict__ U, const double* __restrict__ sU,
                            double* __restrict__ sGb, double* __restrict__ sRb )
{
  int b = blockIdx.x*blockDim.x + threadIdx.x;
  if (b >= nbn) return;
  for (int c=0; c<ns; ++c) {
    double g3[3] = { 0.0, 0.0, 0.0 }, r = 0.0;
    for (int i=bn_off[b]; i<bn_off[b+1]; ++i) {
      int f = bn_face[i] >> 2, k = bn_face[i] & 3;
      int N[3] = { tri[f*3+0], tri[f*3+1], tri[f*3+2] };
      double n[3] = { fn[(size_t)f*3+0], fn[(size_t)f*3+1], fn[(size_t)f*3+2] };
      double u[3], fl[3];
      for (int m=0; m<3; ++m) {
        u[m] = sU[(size_t)c*NP + N[m]];
        double rho = U[N[m]], ru = U[NP+N[m]], rv = U[2*NP+N[m]], rw = U[3*NP+N[m]];
        double vn = besym[f*3+m] ? 0.0 : (n[0]*ru + n[1]*rv + n[2]*rw)/rho;
        fl[m] = u[m]*vn;
      }
      // gradient part, Riemann.cpp:334-360 (g indexed by the direction j, as the reference has it)
      double uab = (u[0] + u[1])/4.0, ubc = (u[1] + u[2])/4.0, uca = (u[2] + u[0])/4.0;
      double g[3] = { uab + uca + u[0], uab + ubc + u[1], ubc + uca + u[2] };
      for (int j=0; j<3; ++j) g3[j] += g[j] * n[j];
      // flux part, Riemann.cpp:866-872
      double fab = (fl[0] + fl[1])/4.0, fbc = (fl[1] + fl[2])/4.0, fca = (fl[2] + fl[0])/4.0;
      r += k == 0 ? fab + fca + fl[0] : (k == 1 ? fab + fbc + fl[1] : fbc + fca + fl[2]);
    }
    for (int j=0; j<3; ++j) sGb[((size_t)b*ns + c)*3 + j] = g3[j];
    sRb[(size_t)b*ns + c] = r;
  }
}

__global__ void k_scal_grad( size_t npoin, int ns, size_t NP, const long long* __restrict__ sl_base,
                             const int2* __restrict__ inc_eq, const double* __restrict__ D, size_t nslot,
                             const double* __restrict__ sU, const int* __restrict__ bslot, const double* __restrict__ sGb,
                             const double* __restrict__ vol, double* __restrict__ sG )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  int b = bslot[p];
  double ivp = 1.0 / vol[p];
  for (int c=0; c<ns; ++c) {
    double up = sU[(size_t)c*NP + p], a[3] = { 0.0, 0.0, 0.0 };
    for (int k=0; k<kmax; ++k) {
      int2 eq = __ldg( inc_eq + base + (long long)k*32 + lane );
      if (eq.x == 0) continue;
      double sg = eq.x > 0 ? 1.0 : -1.0;
      size_t sl = (size_t)(abs(eq.x)-1);
      double s = sU[(size_t)c*NP + (size_t)eq.y] + up;
      for (int j=0; j<3; ++j) a[j] += sg * D[(size_t)j*nslot + sl] * s;
    }
    for (int j=0; j<3; ++j) {
      if (b >= 0) a[j] += sGb[((size_t)b*ns + c)*3 + j];
      sG[(size_t)(c*3+j)*NP + p] = a[j] * ivp;
    }
  }
}

template< bool EXACT >
__global__ void k_scal_flux( size_t nslot, int ns, size_t NP, const int* __restrict__ ep, const int* __restrict__ eq,
                             const double* __restrict__ X, const double* __restrict__ EV, const double* __restrict__ sU,
                             const double* __restrict__ sG, double* __restrict__ sF )
{
  size_t e = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (e >= nslot) return;
  int pi = ep[e];
  if (pi < 0) return;
  size_t p = pi, q = eq[e];
  double vw[3] = { X[q]-X[p], X[NP+q]-X[NP+p], X[2*NP+q]-X[2*NP+p] };
  double cL = EV[e], cR = EV[nslot+e], sw = EV[2*nslot+e];
  for (int c=0; c<ns; ++c) {
    double l = sU[(size_t)c*NP+p], r = sU[(size_t)c*NP+q];
    double g1 = sG[(size_t)(c*3+0)*NP+p]*vw[0] + sG[(size_t)(c*3+1)*NP+p]*vw[1] + sG[(size_t)(c*3+2)*NP+p]*vw[2];
    double g2 = sG[(size_t)(c*3+0)*NP+q]*vw[0] + sG[(size_t)(c*3+1)*NP+q]*vw[1] + sG[(size_t)(c*3+2)*NP+q]*vw[2];
    double d2 = r - l, d1 = 2.0*g1 - d2, d3 = 2.0*g2 - d2, incL, incR;
    vanleer< EXACT >( d1, d2, d3, incL, incR );
    l += incL; r -= incR;
    sF[(size_t)c*nslot + e] = l*cL + r*cR + sw*(r - l);
  }
}

template< bool FUSED >
__global__ void k_scal_node( size_t npoin, int ns, size_t NP, const long long* __restrict__ sl_base,
                             const int* __restrict__ inc_e, const double* __restrict__ sF, size_t nslot,
                             const int* __restrict__ bslot, const double* __restrict__ sRb, const double* __restrict__ sS,
                             const double* __restrict__ v, const double* __restrict__ vol, const double* __restrict__ sUn,
                             double rkdt, const double* __restrict__ dtp, double rk, double* __restrict__ sUo,
                             double* __restrict__ R, int ncomp )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t p = slice*32 + lane;
  if (p >= npoin) return;
  long long base = sl_base[slice];
  int kmax = (int)((sl_base[slice+1] - base) >> 5);
  int b = bslot[p];
  for (int c=0; c<ns; ++c) {
    double acc = 0.0;
    for (int k=0; k<kmax; ++k) {
      int se = __ldg( inc_e + base + (long long)k*32 + lane );
      if (se == 0) continue;
      double f = sF[(size_t)c*nslot + (size_t)(abs(se)-1)];
      acc = se > 0 ? acc + f : acc - f;
    }
    if (b >= 0) acc += sRb[(size_t)b*ns + c];
    if (sS) acc -= sS[p*(size_t)ns + c] * v[p];
    if (FUSED) {
      double f = (dtp ? rk*dtp[p] : rkdt) / vol[p];
      sUo[(size_t)c*NP + p] = sUn[(size_t)c*NP + p] - f*acc;
    } else
      R[p*(size_t)ncomp + 5 + c] = acc;
  }
}

// several partitions (RieCG::comrhs for the scalar columns): the own nodal sums of the shared nodes ->
// part[i][ns]; after the exchange the shared nodes are finished from the complete sums
__global__ void k_scal_sh( int nsh, int ns, const int* __restrict__ sh_node, const long long* __restrict__ sl_base,
                           const int* __restrict__ inc_e, const double* __restrict__ sF, size_t nslot,
                           const int* __restrict__ bslot, const double* __restrict__ sRb, const double* __restrict__ sS,
                           const double* __restrict__ v, double* __restrict__ part )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  int lane = (int)(p & 31);
  long long base = sl_base[p >> 5];
  int kmax = (int)((sl_base[(p >> 5)+1] - base) >> 5);
  int b = bslot[p];
  for (int c=0; c<ns; ++c) {
    double acc = 0.0;
    for (int k=0; k<kmax; ++k) {
      int se = __ldg( inc_e + base + (long long)k*32 + lane );
      if (se == 0) continue;
      double f = sF[(size_t)c*nslot + (size_t)(abs(se)-1)];
      acc = se > 0 ? acc + f : acc - f;
    }
    if (b >= 0) acc += sRb[(size_t)b*ns + c];
    if (sS) acc -= sS[p*(size_t)ns + c] * v[p];
    part[(size_t)i*ns + c] = acc;
  }
}
template< bool FUSED >
__global__ void k_scal_shfin( int nsh, int ns, size_t NP, const int* __restrict__ sh_node, const int* __restrict__ roff,
                              const int* __restrict__ ridx, const double* __restrict__ part, const double* __restrict__ recvbuf,
                              const double* __restrict__ vol, const double* __restrict__ sUn, double rkdt,
                              const double* __restrict__ dtp, double rk, double* __restrict__ sUo, double* __restrict__ R, int ncomp )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nsh) return;
  size_t p = sh_node[i];
  for (int c=0; c<ns; ++c) {
    double acc = part[(size_t)i*ns + c];
    for (int r=roff[i]; r<roff[i+1]; ++r) acc += recvbuf[(size_t)ridx[r]*ns + c];
    if (FUSED) {
      double f = (dtp ? rk*dtp[p] : rkdt) / vol[p];
      sUo[(size_t)c*NP + p] = sUn[(size_t)c*NP + p] - f*acc;
    } else
      R[p*(size_t)ncomp + 5 + c] = acc;
  }
}

// RieCG::solve for the scalar columns of a materialised R (xyst_rk_update)
__global__ void k_scal_update( size_t npoin, int ns, size_t NP, const double* __restrict__ R, int ncomp,
                               const double* __restrict__ vol, const double* __restrict__ sUn, double rkdt,
                               const double* __restrict__ dtp, double rk, double* __restrict__ sU )
{
  size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (p >= npoin) return;
  double f = (dtp ? rk*dtp[p] : rkdt) / vol[p];
  for (int c=0; c<ns; ++c) sU[(size_t)c*NP + p] = sUn[(size_t)c*NP + p] - f*R[p*(size_t)ncomp + 5 + c];
}

// problems::point_src (Problems.cpp:764-823; applied by RieCG::solve, RieCG.cpp:1023-1025): the first
// scalar is set to 1 at the listed nodes after every stage update, before the BCs
__global__ void k_scal_pin( int n, size_t NP, const int* __restrict__ node, double value, double* __restrict__ sU )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) sU[(size_t)node[i]] = value;
}

// physics::dirbc for the scalar components (BC.cpp:29-63)
__global__ void k_scal_bc( int nbc, int ns, size_t NP, const int* __restrict__ node, const int* __restrict__ dir,
                           const int* __restrict__ smask, const double* __restrict__ sval, double* __restrict__ sU )
{
  int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= nbc) return;
  int d = dir[i];
  if (d < 0) return;
  size_t p = node[i];
  for (int c=0; c<ns; ++c) if (smask[d*ns+c] == 1) sU[(size_t)c*NP + p] = sval[d*ns+c];
}

// NodeDiagnostics::rhocompute for the scalar components: per scalar (sum u^2 v, sum (u-un)^2 v,
// sum (u-an)^2 v, sum |u-an| v); an = analytic values [npoin][ns] or null
__global__ void __launch_bounds__(RED_THREADS)
k_scal_diag( size_t npoin, int c, int ns, size_t NP, const double* __restrict__ sU, const double* __restrict__ sUn,
             const double* __restrict__ v, const double* __restrict__ an, double* __restrict__ part )
{
  double a[4] = { 0.0, 0.0, 0.0, 0.0 };
  for (size_t p = blockIdx.x*(size_t)blockDim.x + threadIdx.x; p < npoin; p += (size_t)gridDim.x*blockDim.x) {
    double vp = v[p], u = sU[(size_t)c*NP + p], d = u - sUn[(size_t)c*NP + p];
    a[0] += u*u*vp; a[1] += d*d*vp;
    if (an) { double du = u - an[p*(size_t)ns + c]; a[2] += du*du*vp; a[3] += fabs(du)*vp; }
  }
  block_reduce< 4, false >( a, part );
}

// xyst_b200/csrc/cg_kernels.cuh -- linear solver device code: sliced-ELL SpMV (with on-the-fly Dirichlet rows/columns) and the fused conjugate-gradient vector kernels
// Part of the single translation unit xyst_b200.cu (included inside its anonymous namespace).

// ---------------------------------------------------------------------------------
// linear solver: CSR::mult (CSR.cpp:154-172) and the vector operations of
// ConjugateGradients.cpp:584-823. Matrix in sliced ELL over scalar rows (one warp per 32
// rows, entry k of lane l at base + 32k + l: coalesced), dots by fixed two-pass trees.
// device scalars: [0] rho [1] rho0 [2] alpha [3] beta [4] normr2 [5] finished
//                 [6] done (stop test of ConjugateGradients::x :787-823 evaluated on the device)
//                 [7] iterations done   [8..] partial sums handed to the all-reduce
// The kernels of the iteration take `done` (= scal + 6, or NULL outside the iteration) and return
// at once when it is set: the host may enqueue several iterations per read-back of the stop test
// and still stops exactly where the reference does.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_spmv( size_t nrow, const long long* __restrict__ base, const int* __restrict__ col,
        const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
        const double* __restrict__ done )
{
  if (done && *done != 0.0) return;
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t r = slice*32 + lane;
  if (r >= nrow) return;
  long long b = base[slice];
  int kmax = (int)((base[slice+1] - b) >> 5);
  double acc = 0.0;
  #pragma unroll 4
  for (int k=0; k<kmax; ++k) {
    long long i = b + (long long)k*32 + lane;
    acc += __ldg( val + i ) * __ldg( x + __ldg( col + i ) );
  }
  y[r] = acc;
}

// The same product with Dirichlet rows/columns applied on the fly: what tk::CSR::dirichlet
// (CSR.cpp:106-152) leaves in the matrix -- BC columns zero, BC rows the unit row (diag =
// 1/count over the sharing partitions) -- without touching (or having to restore,
// ConjugateGradients.cpp:809) the stored values.
__global__ void __launch_bounds__(256)
k_spmv_bc( size_t nrow, const long long* __restrict__ base, const int* __restrict__ col,
           const double* __restrict__ val, const unsigned char* __restrict__ bc,
           const double* __restrict__ cnt, const double* __restrict__ x, double* __restrict__ y,
           const double* __restrict__ done )
{
  if (done && *done != 0.0) return;
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t r = slice*32 + lane;
  if (r >= nrow) return;
  if (bc[r]) { y[r] = (1.0 / cnt[r]) * x[r]; return; }
  long long b = base[slice];
  int kmax = (int)((base[slice+1] - b) >> 5);
  double acc = 0.0;
  #pragma unroll 4
  for (int k=0; k<kmax; ++k) {
    long long i = b + (long long)k*32 + lane;
    int cl = __ldg( col + i );
    double a = bc[cl] ? 0.0 : __ldg( val + i );
    acc += a * __ldg( x + cl );
  }
  y[r] = acc;
}

// rhs with BCs (ConjugateGradients::apply/r :451-556): b += neumann; b -= A(:,bc) val; b(bc) = val
// (own part r of the column sums: shared rows are summed over the partitions by the caller)
__global__ void __launch_bounds__(256)
k_cg_bc_colsum( size_t nrow, const long long* __restrict__ base, const int* __restrict__ col,
                const double* __restrict__ val, const unsigned char* __restrict__ bc,
                const double* __restrict__ bcval, double* __restrict__ rsum )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t r = slice*32 + lane;
  if (r >= nrow) return;
  long long b = base[slice];
  int kmax = (int)((base[slice+1] - b) >> 5);
  double acc = 0.0;
  for (int k=kmax-1; k>=0; --k) {            // descending column id, the order BCs are applied in (:481)
    long long i = b + (long long)k*32 + lane;
    int cl = __ldg( col + i );
    if (bc[cl]) acc += __ldg( val + i ) * bcval[cl];
  }
  rsum[r] = acc;
}
// diagonal of the sliced-ELL matrix
__global__ void k_cg_getdiag( size_t nrow, const long long* __restrict__ base, const int* __restrict__ col,
                              const double* __restrict__ val, double* __restrict__ diag )
{
  size_t slice = (blockIdx.x*(size_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  size_t r = slice*32 + lane;
  if (r >= nrow) return;
  long long b = base[slice];
  int kmax = (int)((base[slice+1] - b) >> 5);
  double d = 0.0;
  for (int k=0; k<kmax; ++k) {
    long long i = b + (long long)k*32 + lane;
    if ((size_t)col[i] == r && val[i] != 0.0) d = val[i];
  }
  diag[r] = d;
}
__global__ void k_cg_bc_rhs( size_t nrow, const unsigned char* __restrict__ bc, const double* __restrict__ bcval,
                             const double* __restrict__ neu, const double* __restrict__ rsum, double* __restrict__ b )
{
  size_t r = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  double v = b[r];
  if (neu) v += neu[r];
  v -= rsum[r];
  b[r] = bc[r] ? bcval[r] : v;
}
__global__ void k_cg_bc_diag( size_t nrow, const unsigned char* __restrict__ bc, const double* __restrict__ cnt,
                              double* __restrict__ d )
{
  size_t r = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  if (bc[r]) d[r] = 1.0 / cnt[r];
}

// p = z + beta p    (ConjugateGradients::next :584-599)
__global__ void k_cg_p( size_t n, const double* __restrict__ scal, const double* __restrict__ z, double* __restrict__ p )
{
  if (scal[6] != 0.0) return;
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i] = z[i] + scal[3] * p[i];
}

// partial masked dot products (ConjugateGradients::dot :128-151): NV pairs at once
template< int NV >
__global__ void __launch_bounds__(RED_THREADS)
k_cg_dot( size_t n, const double* __restrict__ mask, const double* __restrict__ a0, const double* __restrict__ b0,
          const double* __restrict__ a1, const double* __restrict__ b1, double* __restrict__ part,
          const double* __restrict__ done )
{
  if (done && *done != 0.0) return;
  double s[NV];
  #pragma unroll
  for (int k=0; k<NV; ++k) s[k] = 0.0;
  for (size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x*blockDim.x) {
    double m = mask[i];
    s[0] += m * (a0[i] * b0[i]);
    if (NV > 1) s[1] += m * (a1[i] * b1[i]);
  }
  block_reduce< NV, false >( s, part );
}

// r -= alpha q ; z = r/d ; x += alpha p ; partial (r,z) and (r,r)   (pq :671-704, rz :717-727)
__global__ void __launch_bounds__(RED_THREADS)
k_cg_update( size_t n, const double* __restrict__ scal, const double* __restrict__ mask,
             const double* __restrict__ q, const double* __restrict__ d, const double* __restrict__ p,
             double* __restrict__ r, double* __restrict__ z, double* __restrict__ x, double* __restrict__ part )
{
  if (scal[6] != 0.0) return;
  double s[2] = { 0.0, 0.0 };
  double alpha = scal[2];
  for (size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x*blockDim.x) {
    double ri = r[i] - alpha * q[i];
    double zi = ri / d[i];
    r[i] = ri; z[i] = zi;
    x[i] += alpha * p[i];
    double m = mask[i];
    s[0] += m * (ri * zi);
    s[1] += m * (ri * ri);
  }
  block_reduce< 2, false >( s, part );
}

// scalar bookkeeping after a reduction; mode 0: alpha = rho/(p,q) (pq :671-690);
// mode 1: rho0 = rho, rho = (r,z), beta = rho/rho0, normr2 = (r,r)  (next :586-588, rz :719)
// and the stop test: finished, or ||r|| < thr (= tol * max-ed ||b||), or maxit iterations
__global__ void k_cg_scalars( int mode, double* __restrict__ scal, double thr, double maxit )
{
  if (threadIdx.x || blockIdx.x) return;
  if (scal[6] != 0.0) return;
  if (mode == 0) {
    double d = scal[8];
    if (fabs(d) < 2.220446049250313e-16) { scal[5] = 1.0; scal[2] = 0.0; } else scal[2] = scal[0] / d;
  } else {
    scal[1] = scal[0];
    scal[0] = scal[8];
    scal[3] = scal[0] / scal[1];
    scal[4] = scal[9];
    scal[7] += 1.0;
    if (scal[5] != 0.0 || sqrt( scal[4] ) < thr || scal[7] >= maxit) scal[6] = 1.0;
  }
}

// shared rows: sum of the sharers (halo) then optionally divide by the count (x :772-785)
__global__ void k_cg_shared_get( int nsh, int w, const int* __restrict__ sh_node, const double* __restrict__ v, double* __restrict__ part )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nsh*w) return;
  part[i] = v[(size_t)sh_node[i/w]*w + i%w];
}
__global__ void k_cg_shared_put( int nsh, int w, const int* __restrict__ sh_node, const int* __restrict__ roff,
                                 const int* __restrict__ ridx, const double* __restrict__ part,
                                 const double* __restrict__ recvbuf, const double* __restrict__ cnt, int average,
                                 double* __restrict__ v )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)nsh*w) return;
  size_t s = i / w, c = i % w;
  double a = part[i];
  for (int r=roff[s]; r<roff[s+1]; ++r) a += recvbuf[(size_t)ridx[r]*w + c];
  size_t row = (size_t)sh_node[s]*w + c;
  v[row] = average ? a / cnt[row] : a;
}
__global__ void k_cg_inv( size_t n, const double* __restrict__ cnt, double* __restrict__ d )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i < n) d[i] = 1.0 / cnt[i];
}
__global__ void k_cg_div( size_t n, const double* __restrict__ r, const double* __restrict__ d, double* __restrict__ z )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i < n) z[i] = r[i] / d[i];
}
__global__ void k_cg_resid( size_t n, const double* __restrict__ b, double* __restrict__ r, double* __restrict__ p )
{
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  if (i < n) { double v = r[i] * -1.0 + b[i]; r[i] = v; p[i] = v; }     // initres :293-298
}

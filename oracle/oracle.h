/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 * C interface of the CPU oracle (serial restatement of the reference's RieCG
 * path, oracle/driver.hpp). Only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline legs may load this library; the product never does. */
#ifndef XYST_ORACLE_H
#define XYST_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_cfg {
  char problem[32];
  char flux[16];
  int32_t ncomp;
  int32_t stab2;
  int32_t steady;
  int32_t nsym;
  int32_t sym[16];
  int32_t ndir;
  int32_t dir[16][12];        /* { setid, mask_0 .. mask_{ncomp-1} } */
  int32_t nfar;
  int32_t far_sets[16];
  int32_t npre;
  int32_t pre_sets[16];
  int32_t nfieldout;
  int32_t fieldout_sets[16];
  char solver[16];            /* "riecg" | "zalcg" | "kozcg" | "laxcg" */
  int32_t fct, fctclip, nfctsys;
  int32_t fctsys[8];
  double fctdif;
  uint64_t nstep;
  uint64_t diag_iter;
  double gamma, p0, cfl, dt, t0, term, stab2coef;
  double far_density, far_pressure, far_velocity[3];
  double pre_density[16], pre_pressure[16];
  /* LaxCG + steady state + user-defined initial conditions */
  double rgas, turkel, velinf[3], residual;
  uint64_t rescomp;
  double ic_density, ic_pressure, ic_velocity[3];
  /* ChoCG: viscosity/diffusivity, stabilisation, RK stages, noslip + valued Dirichlet BCs,
   * pressure solve (iterations, tolerance, preconditioner, BCs, hydrostat node) */
  double mu, dif;
  int32_t stab;
  uint64_t rk;
  int32_t nnoslip; int32_t noslip[16];
  int32_t ndirval; double dirval[16][12];
  uint64_t p_iter; double p_tol; char p_pc[16];
  int32_t np_dir; int32_t p_dir[16][2];
  int32_t np_dirval; double p_dirval[16][2];
  int32_t np_sym; int32_t p_sym[16];
  int32_t p_hydrostat_set; uint64_t p_hydrostat;
  /* manufactured-solution parameters (problem_alpha, problem_kappa; vortical_flow etc.) */
  double alpha, kappa;
  /* problem_r0, problem_ce, problem_beta (nonlinear_energy_growth, rayleigh_taylor) */
  double r0, ce, beta[3];
  double soundspeed;          /* LohCG artificial speed of sound */
  double src_location[3], src_radius, src_release_time;   /* problems::point_src; radius 0 = not configured */
  double freezeflow, freezetime;   /* ZalCG/KozCG scalar transport in a frozen flow; freezeflow 0 = 1.0 */
  /* ChoCG semi-implicit momentum solve: theta > 0 turns it on; iterations (0 = 10), tolerance, preconditioner */
  double theta; uint64_t mom_iter; double mom_tol; char mom_pc[16];
  double fctfreeze;           /* ZalCG steady state: freeze the FCT limit coefficients below this residual; 0 = never */
} orc_cfg;

const char* orc_backend(void);      /* "port" or "reference" */
const char* orc_last_error(void);
void* orc_create( size_t npoin, const double* x, const double* y, const double* z,
                  size_t ntet, const uint64_t* tets, size_t ntri, const uint64_t* tris,
                  int nblocks, const int* block_type, const uint64_t* block_n,
                  int nsets, const int* set_id, const uint64_t* set_off,
                  const uint64_t* set_elem, const uint64_t* set_side,
                  const orc_cfg* cfg, int nchare, const uint64_t* target );
void orc_destroy( void* h );
int orc_step( void* h, int nsteps );
size_t orc_ndiag( void* h );
size_t orc_diagrow( void* h, size_t i, double* out, size_t cap );
double orc_scalar( void* h, const char* name );
size_t orc_get( void* h, int chare, const char* name, void* out, size_t cap_bytes );
int orc_set_u( void* h, int chare, const double* u );
int orc_kernel( void* h, int chare, const char* what, int stage, double t, double dt );
uint64_t orc_siphash_ids( const uint64_t* ids, int n );

/* linear solver: tk::CSR + ConjugateGradients restatement (cg_port.hpp) */
void* orc_cg_create( const char* pc );
void orc_cg_destroy( void* h );
const char* orc_cg_backend(void);
int orc_cg_add( void* h, size_t npoin, size_t ncomp, size_t npsup1, const uint64_t* psup1,
                const uint64_t* psup2, const uint64_t* gid, int ncomm, const int* comm_rank,
                const uint64_t* comm_off, const uint64_t* comm_gid );
int orc_cg_laplacian( void* h, int part, size_t ntet, const uint64_t* inpoel,
                      const double* x, const double* y, const double* z );
int orc_cg_dirichlet( void* h, int part, size_t node, double val, size_t pos );
int orc_cg_set( void* h, int part, const double* x, const double* b );
size_t orc_cg_get( void* h, int part, const char* name, void* out, size_t cap );
int orc_cg_mult( void* h, int part, const double* x, double* r );
double orc_cg_setup( void* h );
double orc_cg_solve( void* h, size_t maxit, double tol, uint64_t* it );

#ifdef __cplusplus
}
#endif
#endif

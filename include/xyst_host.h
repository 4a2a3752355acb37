/* xyst_host.h -- C interface of the C++ host mirror (xyst_b200/host): the restated
 * Discretization + RieCG solver classes of the reference (src/Inciter/RieCG.cpp,
 * Discretization.cpp) driving the device C ABI of xyst_b200.h. Used by tests and
 * bench.py through ctypes; a C++ application links the classes directly.
 * Returns 0 on success; xyst_host_last_error() has the message otherwise. */
#ifndef XYST_HOST_H
#define XYST_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "xyst_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Control-file equivalent (the tags the RieCG path reads; src/Control/InciterConfig.hpp) */
typedef struct xyst_host_cfg {
  char problem[32];           /* "sod" | "sedov" | "taylor_green" | "vortical_flow" | "nonlinear_energy_growth" | "rayleigh_taylor" | "userdef" | (ChoCG) "poiseuille", "poisson_*" */
  char flux[16];              /* "rusanov" | "hllc" */
  int32_t ncomp;
  int32_t stab2;
  int32_t exact_muscl;
  int32_t reforder;           /* triangle superedge walk: 1 reference hash order, 0 element order, -1 auto */
  int32_t nsym;  int32_t sym[16];
  int32_t ndir;  int32_t dir[16][12];      /* { setid, mask_0 .. mask_{ncomp-1} } */
  int32_t nfar;  int32_t far_sets[16];
  int32_t npre;  int32_t pre_sets[16];
  uint64_t nstep;
  uint64_t diag_iter;
  double gamma, p0, cfl, dt, t0, term, stab2coef;
  double far_density, far_pressure, far_velocity[3];
  double pre_density[16], pre_pressure[16];
  char solver[16];            /* "riecg" (default when empty) | "zalcg" | "kozcg" | "laxcg" | "chocg" */
  int32_t fct, fctclip, nfctsys;
  int32_t fctsys[8];
  double fctdif;
  /* steady-state local time stepping, LaxCG preconditioning, user-defined initial conditions */
  int32_t steady;
  uint64_t rescomp;
  double residual;
  double rgas, turkel, velinf[3];      /* rgas = 0: reference default 287.052874 */
  double ic_density, ic_pressure, ic_velocity[3];
  /* ChoCG (solver "chocg", ncomp 3): flux "damp2" | "damp4", viscosity/diffusivity, stabilisation,
   * RK stages (1..4, 0 = 1), no-slip sets, Dirichlet values { setid, v_0..v_2 }, pressure solve
   * (iterations, tolerance, preconditioner "none" | "jacobi", Dirichlet { setid, mask } + values
   * { setid, value }, Neumann sets, hydrostat node (global id) if p_hydrostat_set) */
  double mu, dif;
  int32_t stab;
  int32_t nnoslip; int32_t noslip[16];
  uint64_t rk;
  int32_t ndirval; double dirval[16][12];
  uint64_t p_iter; double p_tol; char p_pc[16];
  int32_t np_dir; int32_t p_dir[16][2];
  int32_t np_dirval; double p_dirval[16][2];
  int32_t np_sym; int32_t p_sym[16];
  int32_t p_hydrostat_set; uint64_t p_hydrostat;
  double alpha, kappa;        /* problem_alpha, problem_kappa ("vortical_flow") */
  double r0, ce, beta[3];     /* problem_r0, problem_ce, problem_beta ("nonlinear_energy_growth", "rayleigh_taylor") */
  /* ChoCG semi-implicit momentum solve (theta > 0): CG iterations (0 = 10), tolerance, preconditioner */
  double theta; uint64_t mom_iter; double mom_tol; char mom_pc[16];
  double soundspeed;          /* LohCG (solver "lohcg", ncomp 4: p,u,v,w): artificial sound speed; 0 = reference default 1.0 */
  /* problem "point_src" (problems::point_src, src/Physics/Problems.cpp:764-823): the first transported scalar is
     set to 1 inside a sphere from the release time on, after every stage. src_radius < 0: no source. */
  double src_location[3], src_radius, src_release_time;
  /* frozen flow (tag::freezeflow / freezetime; ChoCG::dt :1396-1399, solve :1550-1570): once t > freezetime the
     time step is multiplied by freezeflow (> 1) and only the transported scalars advance. 0 = off (1.0). */
  double freezeflow, freezetime;
} xyst_host_cfg;

typedef struct xyst_solver xyst_solver;
/* Communication hooks for the setup/control path (default: NCCL through the device
 * context): op 0 = sum over the partitions sharing each unique shared node (n = number of
 * shared nodes, w values each), op 1 = all-reduce sum, op 2 = all-reduce min (n*w values) */
typedef void (*xyst_comm_fn)( void* user, int op, int w, size_t n, double* vals );

const char* xyst_host_last_error(void);

/* Partition `part` of `nparts` (1,2,4,8; coordinate bisection) of the structured box
 * [0,L]^3 of nx*ny*nz hexahedra split into 6 tets each; side sets 1..6 = x-,x+,y-,y+,z-,z+ */
int xyst_solver_create_box(const xyst_host_cfg* cfg, size_t nx, size_t ny, size_t nz,
                           double Lx, double Ly, double Lz, int nparts, int part,
                           xyst_solver** out);
/* Partition `part` of a general tet mesh given in full (global node ids): tetpart[e] is the
 * partition of tet e (NULL: recursive coordinate bisection into nparts). Side set s has the
 * boundary triangles set_tri[3*set_off[s] .. 3*set_off[s+1]). */
int xyst_solver_create_mesh(const xyst_host_cfg* cfg, size_t npoin, const double* x,
                            const double* y, const double* z, size_t ntet, const uint64_t* tets,
                            int nsets, const int* set_id, const uint64_t* set_off,
                            const uint64_t* set_tri, int nparts, int part, const int32_t* tetpart,
                            xyst_solver** out);
/* The same from an ExodusII mesh file (NetCDF classic CDF-1/CDF-2; 4-node tetrahedra, side sets on
 * tetrahedron faces or triangle-block elements), cf. src/IO/ExodusIIMeshReader.cpp */
int xyst_solver_create_exo(const xyst_host_cfg* cfg, const char* path, int nparts, int part,
                           xyst_solver** out);
/* Read such a file into arrays (call with NULL pointers first for the sizes): coordinates, tets,
 * side-set ids / offsets (in triangles) / triangles. */
int xyst_exo_read(const char* path, size_t* npoin, size_t* ntet, int* nsets, size_t* ntri,
                  double* x, double* y, double* z, uint64_t* tets, int32_t* set_id, uint64_t* set_off,
                  uint64_t* set_tri);
/* Write every diagnostics row of the following xyst_solver_step calls to a text file in the format of
 * the reference's diag file (src/IO/DiagWriter.cpp; scientific, given precision, 0 = 8 as the default) */
int xyst_solver_diag_file(xyst_solver* s, const char* path, int precision);
int xyst_solver_destroy(xyst_solver* s);

int xyst_solver_prepare(xyst_solver* s);     /* host only: renumber, volumes, edge integrals, superedges */
int xyst_solver_attach(xyst_solver* s, int device, int nranks, int rank, const void* ncclid128);
int xyst_solver_set_comm(xyst_solver* s, xyst_comm_fn fn, void* user, int nranks, int rank);
int xyst_solver_set_u0(xyst_solver* s, const double* u0);   /* user-defined IC, npoin x ncomp, local order */
int xyst_solver_host_setup(xyst_solver* s);  /* exchanges (volumes, normals), BC lists, ICs: no device */
int xyst_solver_set_u(xyst_solver* s, const double* u);    /* overwrite the device state (after setup) */
int xyst_solver_setup(xyst_solver* s);       /* host_setup if needed + device upload + BCs */
/* Advance up to nsteps time steps; diagnostics rows (ncols doubles each, layout of the
 * reference's diag file: it t dt L2(U)x5 L2(dU)x5 mE [L2err x5 L1err x5]) are appended to
 * rows (capacity cap doubles); *nrows, *ncols report what was written. */
int xyst_solver_step(xyst_solver* s, int nsteps, double* rows, size_t cap, size_t* nrows, size_t* ncols);
/* Unfused stepping through the reference-shaped members dt/advance/grad/rhs/solve. */
int xyst_solver_step_unfused(xyst_solver* s, int nsteps);

double xyst_solver_scalar(xyst_solver* s, const char* name); /* npoin ntet nedge t dt it meshvol finished nshared */
/* copy out an array by name (same names as the oracle's): returns bytes needed */
size_t xyst_solver_get(xyst_solver* s, const char* name, void* out, size_t cap_bytes);
xyst_ctx* xyst_solver_ctx(xyst_solver* s);

/* Stand-alone mesh helpers (tests): full box mesh and RCB */
int xyst_box_counts(size_t nx, size_t ny, size_t nz, size_t* npoin, size_t* ntet, size_t* ntri);
int xyst_box_mesh(size_t nx, size_t ny, size_t nz, double Lx, double Ly, double Lz,
                  double* x, double* y, double* z, uint64_t* tets,
                  int32_t set_id[6], uint64_t set_off[7], uint64_t* set_tri);
/* Number of partitions ("chares") the reference creates on npe processing elements with virtualization u in [0,1]
 * (command-line -u; 0 = one per PE): tk::linearLoadDistributor, src/Base/LoadDistributor.cpp:24-97, with load = the
 * number of mesh elements. Returns the count; chunksize/remainder as the reference computes them (may be NULL). */
int xyst_chare_count(double virtualization, uint64_t load, int npe, uint64_t* chunksize, uint64_t* remainder, uint64_t* nchare);
/* Element partition as the reference gets it from Zoltan's RCB (Partition/ZoltanGeom.cpp:139-244 with Zoltan 3.901's
 * serial_rcb / find_median restated; any number of parts): part[e] for every tetrahedron */
int xyst_rcb(size_t npoin, const double* x, const double* y, const double* z, size_t ntet,
             const uint64_t* tets, int nparts, int32_t* part);
/* ... and from Zoltan's RIB (part = "rib": serial_rib + Zoltan_RIB_inertial3d restated) */
int xyst_rib(size_t npoin, const double* x, const double* y, const double* z, size_t ntet,
             const uint64_t* tets, int nparts, int32_t* part);
int xyst_test_faceset_order(size_t nface, const uint64_t* faces, size_t nerase, const uint64_t* erase,
                            uint64_t* out_std, uint64_t* out_emu, size_t* nout);
int xyst_box_part_range(size_t nx, size_t ny, size_t nz, int nparts, int part, uint64_t range[6]);

#ifdef __cplusplus
}
#endif
#endif

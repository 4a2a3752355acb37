/* xyst_b200.h -- C ABI of the B200-native (sm_100a) RieCG hot path.
 *
 * The reference (jbakosi/xyst) has no plugin/FFI layer; its solver chares call
 * free functions in namespaces with caller-owned std::vector / tk::Fields
 * arguments and a hidden global configuration. This header is the
 * replacement seam: an opaque per-GPU context owns device-resident connectivity
 * and nodal state, and each entry point replaces one reference call site. Array
 * arguments use the reference's own host layouts (tk::Fields = row-major
 * [node][component] doubles, std::size_t ids) so a maintainer can bind them with
 * .data() pointers; see INTEGRATION.md for the stubs.
 *
 * All functions return 0 on success and a non-zero code on failure;
 * xyst_last_error() returns the message of the last failure on this thread (the
 * reference throws tk::Exception instead, src/Base/Exception.hpp:40-52). There is
 * no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef XYST_B200_H
#define XYST_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct xyst_ctx xyst_ctx;

/* Replaces the reads of inciter::g_cfg made inside the kernels
 * (src/Physics/Riemann.cpp:449,621,675; src/Physics/EOS.hpp:31,41,55). */
typedef struct xyst_params {
  int32_t ncomp;       /* scalar components per node; 5 = Euler system           */
  int32_t flux;        /* 0 = "rusanov" (Riemann.cpp:369), 1 = "hllc" (:480)      */
  int32_t stab2;       /* tag::stab2                                             */
  int32_t exact_muscl; /* 1: van Leer limiter with the reference's 8 divisions per
                          component (Riemann.cpp:92-99); 0: algebraically equal
                          form with 2 reciprocals per component                   */
  double gamma;        /* tag::mat_spec_heat_ratio                               */
  double stab2coef;    /* tag::stab2coef                                         */
} xyst_params;

const char* xyst_last_error(void);
int xyst_device_count(void);

/* Context: one per GPU / mesh partition (one reference chare). */
int xyst_ctx_create(int device, const xyst_params* params, xyst_ctx** out);
/* Launch on a caller-owned stream (a cudaStream_t), e.g. torch's current stream. */
int xyst_ctx_set_stream(xyst_ctx* ctx, void* cuda_stream);
int xyst_ctx_destroy(xyst_ctx* ctx);
int xyst_sync(xyst_ctx* ctx);

/* Device-resident mesh data of one partition. Arguments are the members the
 * reference's RieCG chare passes to riemann::grad/rhs (src/Inciter/RieCG.cpp:879,948):
 *   coord            Discretization::Coord()            3 arrays of npoin
 *   dsupedge[3]      RieCG::m_dsupedge  (RieCG.cpp:620-736): 4/3/2 node ids per tet/tri/edge
 *   dsupint[3]       RieCG::m_dsupint : 6x3 / 3x3 / 3 doubles per superedge
 *   triinpoel,besym  RieCG::m_triinpoel, m_besym (RieCG.cpp:534-538)
 *   vol, v           Discretization::Vol() (with neighbour contributions), V() (without)
 * nsup[k] is the number of superedges in group k. Superedges may come in any
 * order; the library re-sorts edges for locality and builds its own node
 * incidence structure (results do not depend on the grouping). */
int xyst_mesh_upload(xyst_ctx* ctx, size_t npoin,
                     const double* x, const double* y, const double* z,
                     const size_t nsup[3], const size_t* const dsupedge[3],
                     const double* const dsupint[3],
                     size_t ntri, const size_t* triinpoel, const uint8_t* besym,
                     const double* vol, const double* v);

/* Boundary-condition node lists of RieCG::BC (RieCG.cpp:764-785 -> BC.cpp:29-241):
 *   dirbcmasks  m_dirbcmasks: ndir x (1+ncomp) {node, mask_c...}; dirvals: ndir x ncomp
 *               values of problems::IC() at the node (evaluated by the host, which
 *               re-uploads them with xyst_dirbc_values() if the IC depends on time)
 *   symbcnodes/norms  m_symbcnodes, m_symbcnorms (a node may repeat, applied in order)
 *   farbcnodes/norms  m_farbcnodes, m_farbcnorms + far-field state
 *   prebcnodes/vals   m_prebcnodes, m_prebcvals {density, pressure} */
int xyst_bc_upload(xyst_ctx* ctx,
                   size_t ndir, const size_t* dirbcmasks, const double* dirvals,
                   size_t nsym, const size_t* symbcnodes, const double* symbcnorms,
                   size_t nfar, const size_t* farbcnodes, const double* farbcnorms,
                   double far_density, double far_pressure, const double far_velocity[3],
                   size_t npre, const size_t* prebcnodes, const double* prebcvals);
int xyst_dirbc_values(xyst_ctx* ctx, const double* dirvals);

/* Source term of riemann::src (Riemann.cpp:880-907): S = problems::SRC()(x,y,z,t)
 * per node, npoin x ncomp, NULL for none. R(p,c) -= S(p,c) * v[p]. */
int xyst_src_upload(xyst_ctx* ctx, const double* S);

/* problems::point_src (src/Physics/Problems.cpp:764-823, applied in RieCG::solve, RieCG.cpp:1023-1025): from now
 * on the first transported scalar is set to `value` at the listed nodes after every stage update, before
 * the BCs. n = 0 switches it off. Needs ncomp > 5. */
int xyst_scalar_pin(xyst_ctx* ctx, size_t n, const size_t* nodes, double value);

/* Nodal unknowns, tk::Fields layout npoin x ncomp (RieCG::m_u). */
int xyst_state_set(xyst_ctx* ctx, const double* U);
int xyst_state_get(xyst_ctx* ctx, double* U);

/* riemann::grad (Riemann.cpp:229-367) followed by the nodal normalisation of
 * RieCG::rhs (RieCG.cpp:936-939): G = (weak gradient sums) / vol.
 * xyst_grad_get returns G as npoin x 3*ncomp. */
int xyst_riecg_grad(xyst_ctx* ctx);
int xyst_grad_get(xyst_ctx* ctx, double* G);

/* riemann::rhs (Riemann.cpp:909-946): advdom + advbnd + src using the gradients
 * of the last xyst_riecg_grad. xyst_rhs_get returns R as npoin x ncomp. */
int xyst_riecg_rhs(xyst_ctx* ctx);
int xyst_rhs_get(xyst_ctx* ctx, double* R);
/* For callers that keep the reference's call structure (include/xyst_shim.hpp): gradients G as
 * riemann::rhs receives them (npoin x 3*ncomp, already divided by vol, RieCG.cpp:936-939), and
 * m_besym / V() of an uploaded mesh, which the reference passes to riemann::rhs only. */
int xyst_grad_set(xyst_ctx* ctx, const double* G);
int xyst_besym_upload(xyst_ctx* ctx, const uint8_t* besym);
int xyst_v_upload(xyst_ctx* ctx, const double* v);

/* RieCG::solve update (RieCG.cpp:1011-1021): stage 0 saves un=u; u = un - rk*dt*R/vol,
 * with R from the last xyst_riecg_rhs. */
int xyst_rk_update(xyst_ctx* ctx, int stage, double dt);

/* RieCG::BC (RieCG.cpp:764-785). */
int xyst_apply_bc(xyst_ctx* ctx);

/* RieCG::dt (RieCG.cpp:827-839): cfl * min_p cbrt(vol_p)/max(|v|+c,1e-8), this partition. */
int xyst_dt_min(xyst_ctx* ctx, double cfl, double* dt);
/* The same followed by the minimum over all partitions of the communicator (the reference's
 * contribute(min_double) to the host, RieCG.cpp:850), reduced on the device with NCCL: one wait per step. */
int xyst_dt_min_all(xyst_ctx* ctx, double cfl, double* dt);

/* One RK stage without materialising R: grad, edge fluxes, nodal gather fused with
 * the update, BCs. Equivalent to grad+rhs+rk_update+apply_bc. */
int xyst_riecg_stage(xyst_ctx* ctx, int stage, double dt);
/* Three stages (rkcoef = 1/3, 1/2, 1; RieCG.cpp:41). */
int xyst_riecg_step(xyst_ctx* ctx, double dt);

/* NodeDiagnostics::rhocompute sums (NodeDiagnostics.cpp:85-118) over this
 * partition: out[0..nc) = sum u^2 v, out[nc..2nc) = sum (u-un)^2 v, out[2nc] = sum u_4 v,
 * and, if `an` (npoin x ncomp analytic PRIMITIVE solution) is given,
 * out[2nc+1..3nc+1) = sum (prim-an)^2 v, out[3nc+1..4nc+1) = sum |prim-an| v.
 * `out` must hold 4*ncomp+1 doubles. */
int xyst_diag(xyst_ctx* ctx, const double* an, double* out);

/* Shared-node ("chare-boundary") partial-sum exchange, replacing comgrad/comrhs
 * (RieCG.cpp:881-916,955-990). neigh_rank[i] shares the nodes
 * shared[neigh_off[i]..neigh_off[i+1]) (local ids, SAME ORDER on both sides, i.e.
 * ordered by global id). xyst_comm_unique_id + xyst_comm_init create the NCCL
 * communicator (id bytes are broadcast by the caller, e.g. torch.distributed). */
int xyst_comm_unique_id(void* id128);
int xyst_comm_init(xyst_ctx* ctx, int nranks, int rank, const void* id128);
int xyst_halo_upload(xyst_ctx* ctx, int nneigh, const int* neigh_rank,
                     const size_t* neigh_off, const size_t* shared);
/* Setup-time helper: sum w (<= 15) doubles per UNIQUE shared node (ascending local id)
 * over all partitions sharing it, in place on a host array -- replaces comvol/comnorm
 * (src/Inciter/Discretization.cpp:663-702, src/Inciter/RieCG.cpp:264-277,384-407). */
int xyst_halo_sum(xyst_ctx* ctx, int w, double* vals);
/* NCCL all-reduce helpers for dt (min) and diagnostics (sum) over the communicator. */
int xyst_allreduce_min(xyst_ctx* ctx, double* v, int n);
int xyst_allreduce_sum(xyst_ctx* ctx, double* v, int n);

/* ---- ZalCG: Taylor-Galerkin edge flux + flux-corrected transport ------------------------
 * (src/Physics/Zalesak.cpp, src/Inciter/ZalCG.cpp:990-1607). Same context, state, BC and
 * dt/diagnostics entry points as RieCG; the superedge integrals have stride 4
 * (ZalCG::m_dsupint: normal + J/120, ZalCG.cpp:354-400). */
typedef struct xyst_zalcg_params {
  int32_t fct;          /* tag::fct (default true)                                  */
  int32_t fctclip;      /* tag::fctclip                                             */
  int32_t fctsys_mask;  /* bit c set <=> component c+1 listed in tag::fctsys         */
  int32_t pad_;
  double fctdif;        /* tag::fctdif                                              */
} xyst_zalcg_params;
int xyst_zalcg_config(xyst_ctx* ctx, const xyst_zalcg_params* p);
int xyst_zalcg_mesh_upload(xyst_ctx* ctx, size_t npoin,
                           const double* x, const double* y, const double* z,
                           const size_t nsup[3], const size_t* const dsupedge[3],
                           const double* const dsupint[3],
                           size_t ntri, const size_t* triinpoel, const uint8_t* besym,
                           const double* vol, const double* v);
/* zalesak::rhs (Zalesak.cpp:421-457) for problems without source term -> R (xyst_rhs_get) */
int xyst_zalcg_rhs(xyst_ctx* ctx, double dt);
/* one ZalCG time step: rhs, aec (P+/-), alw (low-order solution, Q+/-), lim (C+/-, limited
 * antidiffusive contributions), solve, BC (ZalCG.cpp:990-1607); un keeps the old state */
int xyst_zalcg_step(xyst_ctx* ctx, double dt);
/* Transported scalars in ZalCG (contexts with ncomp > 5; scalar rows of zalesak::rhs, Zalesak.cpp:113-116,
 * 146-149,193-196, boundary term :380-403, and the FCT passes per scalar; one partition) need nothing extra.
 * Source term of zalesak::rhs (Zalesak.cpp:118-128,152-163): the caller evaluates problems::SRC at the nodes
 * (time t) and at the midpoints of the device's edge slots (time t + dt/2): */
size_t xyst_nslot(xyst_ctx* ctx);
int xyst_edge_list(xyst_ctx* ctx, size_t* p /* [nslot] */, size_t* q /* [nslot]; SIZE_MAX = padding slot */);
int xyst_zalcg_src(xyst_ctx* ctx, const double* Sn /* [npoin][ncomp] */, const double* Se /* [nslot][ncomp] */);
/* frozen flow (tag::freezeflow; ZalCG::dt :948-952, solve :1549,1577-1584): only the scalars advance */
int xyst_zalcg_freeze(xyst_ctx* ctx, int on);

/* ---- LaxCG: time-derivative preconditioning for all Mach numbers ------------------------
 * (src/Physics/Lax.cpp:216-1018, src/Inciter/LaxCG.cpp:115-259,940-1214). Same superedge data
 * and entry points as RieCG: after xyst_laxcg_config the context keeps (p,u,v,w,T) next to the
 * conserved state and xyst_riecg_grad = lax::grad, xyst_riecg_rhs = lax::rhs, xyst_rk_update /
 * xyst_riecg_stage / xyst_riecg_step = LaxCG::solve (update preconditioned by LaxCG::precond),
 * xyst_dt_min = LaxCG::dt with LaxCG::charvel. Call after the mesh upload and before
 * xyst_state_set. rgas = mat_spec_gas_const, turkel and velinf as in the control file. */
typedef struct xyst_laxcg_params {
  double rgas;
  double turkel;
  double velinf[3];
} xyst_laxcg_params;
int xyst_laxcg_config(xyst_ctx* ctx, const xyst_laxcg_params* p);

/* steady = true (RieCG.cpp:812-825,1013, LaxCG.cpp:962-969,1153): local time stepping. xyst_dt_min
 * then also stores the nodal time steps dtp = cfl L/v, and the stage updates use them instead of
 * the dt argument. Sources are evaluated at the nodes once (time-independent). */
int xyst_steady(xyst_ctx* ctx, int on);

/* ---- KozCG: element-based Taylor-Galerkin + flux-corrected transport ---------------------
 * (src/Physics/Kozak.cpp:29-180, src/Inciter/KozCG.cpp:709-1197). Works on the tetrahedra
 * themselves: inpoel = Discretization::Inpoel(), 4 local node ids per tet. FCT parameters
 * through xyst_zalcg_config; state/BC/dt/diagnostics entry points as for RieCG.
 * Ssrc_nodes (npoin x ncomp) and Ssrc_cent (ntet x ncomp) are problems::SRC() evaluated at
 * the nodes and at the tet centroids (time-independent sources), or NULL. */
int xyst_kozcg_mesh_upload(xyst_ctx* ctx, size_t npoin, const double* x, const double* y, const double* z,
                           size_t ntet, const size_t* inpoel, const double* vol, const double* v,
                           const double* Ssrc_nodes, const double* Ssrc_cent);
/* new source values for time-dependent problems: nodes at t, centroids at t + dt/2 (Kozak.cpp:97-108,160-171) */
int xyst_kozcg_src(xyst_ctx* ctx, const double* Ssrc_nodes, const double* Ssrc_cent);
/* kozak::rhs -> R (xyst_rhs_get) */
int xyst_kozcg_rhs(xyst_ctx* ctx, double dt);
/* one KozCG time step: rhs, aec, alw, lim, solve, BC; un keeps the old state */
/* frozen flow (tag::freezeflow; KozCG::dt :669-674, solve :1140-1176): while on, xyst_kozcg_step advances only
 * the transported scalars (contexts with ncomp > 5: kozak::rhs scalar rows + FCT per scalar, one partition) */
int xyst_kozcg_freeze(xyst_ctx* ctx, int on);
int xyst_kozcg_step(xyst_ctx* ctx, double dt);

/* ---- ChoCG: projection method for constant-density flow -----------------------------------
 * Edge operators chorin::div/grad/vgrad/flux/rhs (src/Physics/Chorin.cpp:85-1044) on superedge
 * integrals of stride 5 (normal, J/120, grad_p.grad_q/(6J); ChoCG::domint, ChoCG.cpp:399-446) and
 * the solver steps of src/Inciter/ChoCG.cpp that call them. The state (velocity u, pressure, their
 * gradients, divergence) stays on the device; the pressure Poisson matrix (ChoCG::prelhs :146-188)
 * is given with xyst_csr_upload and solved with the xyst_cg_* entries below. With a communicator
 * and the shared-node lists (xyst_comm_init, xyst_halo_upload BEFORE the matrix upload) every
 * operator is followed by the sum of the sharing partitions' parts at the shared nodes
 * (ChoCG::comdiv/comvgrad/comflux/comsgrad/compgrad/comrhs, ChoCG.cpp:902-1527), the stage update
 * of those nodes is done from the summed rhs, and the linear solves are partitioned (row counts and
 * slave flags from the halo lists; Neumann and Dirichlet column-sum parts summed,
 * ConjugateGradients.cpp:407-428,508-525). The caller passes the UNION of the sharers' Dirichlet
 * rows for shared nodes (ConjugateGradients::init/apply :391-449); dt_min and diag return this
 * partition's values, to be reduced with xyst_allreduce_min/_sum. */
typedef struct xyst_chocg_params {
  int flux;          /* 0 = damp2, 1 = damp4 (Chorin.cpp:640-829) */
  int stab;          /* tag::stab */
  int stab2;         /* tag::stab2 */
  double stab2coef;
  double mu;         /* mat_dyn_viscosity */
} xyst_chocg_params;
int xyst_chocg_mesh_upload(xyst_ctx* ctx, size_t npoin, const double* x, const double* y, const double* z,
                           const size_t nsup[3], const size_t* const dsupedge[3],
                           const double* const dsupint[3], size_t ntri, const size_t* triinpoel,
                           const double* vol, const double* v, const xyst_chocg_params* prm);
/* ChoCG::BC (:1340-1353): physics::dirbc with values (dirmask/dirval: npoin-independent lists of
 * ndir entries { node, mask_0..2 } / { value_0..2 } with mask 1 or 2 = set to value), symbc in the
 * given order (a node may appear once per side set), noslipbc */
int xyst_chocg_bc_upload(xyst_ctx* ctx, size_t ndir, const size_t* dirnodes, const int* dirmask,
                         const double* dirval, size_t nsym, const size_t* symbcnodes,
                         const double* symbcnorms, size_t nnoslip, const size_t* noslipbcnodes);
/* Transported scalars next to the velocity (ChoCG::m_u with problem_ncomp = 3 + ns; the scalar rows of
 * chorin::vgrad / adv_damp2 / adv_damp4 / the boundary integral / src, Chorin.cpp:230,698-708,817-824,
 * 947-981,1003-1007): call after the mesh upload, before any state or BC upload. From then on u, un, rhs
 * rows have 3+ns entries, vgrad rows 3(3+ns), Dirichlet masks and values 3+ns per node, xyst_chocg_src
 * takes 3+ns columns, xyst_chocg_diag takes an_u with 3+ns columns and returns 16 + 4 ns sums
 * (per scalar: L2 solution, L2 increment, L2 error, L1 error). ns <= 4. */
int xyst_chocg_scalars(xyst_ctx* ctx, int ns, double diffusivity /* tag::mat_dyn_diffusivity */);
/* time-dependent Dirichlet values (physics::dirbc evaluates the IC at the BC time, BC.cpp:57-66):
 * new values [ndir][3+ns] for the nodes of the last bc_upload (LohCG: [ndir][4]) */
int xyst_chocg_dirbc_values(xyst_ctx* ctx, const double* dirval);
/* problems::point_src (Problems.cpp:764-823; ChoCG::pred :1655-1657): the first scalar of the listed
 * nodes is set to value after every stage update, before the BCs */
int xyst_chocg_pin(xyst_ctx* ctx, size_t n, const size_t* nodes, double value);
/* frozen flow (tag::freezeflow; ChoCG::solve :1550-1552,1564-1570): velocity rows = those of time level n */
int xyst_chocg_restore_velocity(xyst_ctx* ctx);
int xyst_chocg_set_u(xyst_ctx* ctx, const double* u /* [npoin][3 (+ns)] */);
int xyst_chocg_get_u(xyst_ctx* ctx, double* u);
int xyst_chocg_set_p(xyst_ctx* ctx, const double* p /* [npoin] */);
int xyst_chocg_get(xyst_ctx* ctx, const char* what, double* out);   /* "pr","div" [npoin]; "sgrad","pgrad","flux","rhs" [npoin][3]; "vgrad" [npoin][9]; "un" [npoin][3] */
int xyst_chocg_apply_bc(xyst_ctx* ctx);
/* chorin::div of the velocity (which = 0) or the momentum flux (1: ChoCG::div finishes the flux
 * first: /vol and symbc, :879-882); stab adds the pressure stabilisation (:58-80) */
int xyst_chocg_div(xyst_ctx* ctx, int which, double dt, int stab);
/* chorin::vgrad + ChoCG::fingrad (/vol); chorin::flux (weak, un-normalised as in ChoCG::flux) */
int xyst_chocg_vgrad(xyst_ctx* ctx);
int xyst_chocg_flux(xyst_ctx* ctx);
/* chorin::grad of the CG solution (which = 0 -> sgrad) or of the pressure (1 -> pgrad), /vol */
int xyst_chocg_grad(xyst_ctx* ctx, int which);
/* chorin::rhs -> rhs (xyst_chocg_get "rhs"); source term values at the nodes with xyst_chocg_src */
int xyst_chocg_src(xyst_ctx* ctx, const double* S /* [npoin][3] or NULL */);
int xyst_chocg_rhs(xyst_ctx* ctx);
/* one RK stage of ChoCG::solve + pred (:1529-1668): rhs, u = un - rk dt rhs/vol, BC, and for the
 * damp4 flux the velocity gradient of the new state */
int xyst_chocg_stage(xyst_ctx* ctx, int stage, double rkcoef, double dt);
/* pressure solve set-up, ChoCG::pinit (:1025-1125) + ConjugateGradients::init/apply (:336-556):
 * rhs = div / divisor (:1044; 1.0 = as is), or rhs0 if given (PRESSURE_RHS * vol, :1117-1122);
 * Dirichlet nodes/values and the Neumann vector are applied to the rhs and, on the fly, to the
 * matrix of xyst_csr_upload (which is left untouched, cf. the restore at :809); the initial guess
 * is the previous solution (:363). Continue with xyst_cg_solve. */
int xyst_chocg_pinit(xyst_ctx* ctx, double divisor, size_t nbc, const size_t* bcnodes,
                     const double* bcvals, const double* neubc, const double* rhs0, int pc);
/* Semi-implicit momentum solve (theta > 0) at the last RK stage, ChoCG::solve :1574-1607 + msolve +
 * msolved :1610-1645. With the momentum solver selected (xyst_cg_select 1) and its matrix (ChoCG::lhs
 * :1433-1477, block CSR with 3 scalar rows per node) given by xyst_csr_upload / xyst_csr_update:
 * minit: b = rhs of the last xyst_chocg_rhs, Dirichlet rows (node*3 + component) with value 0, initial
 * guess = previous solution; then xyst_cg_solve; mupdate: u = un + du, BC, velocity gradient for damp4
 * (stage = index of this RK stage: at stage 0 un is the current velocity). */
int xyst_chocg_minit(xyst_ctx* ctx, size_t nbc, const size_t* bcrows, int pc);
int xyst_chocg_mupdate(xyst_ctx* ctx, int stage);
/* u -= pdt * sgrad, then BC (ChoCG::psolved :1204-1217); pr = x or pr += x (:1241,1249) */
int xyst_chocg_project(xyst_ctx* ctx, double pdt);
int xyst_chocg_pressure_update(xyst_ctx* ctx, int increment);
/* ChoCG::dt (:1356-1411): min over nodes of L/|u| and L^2/max(mu,dif), times cfl */
int xyst_chocg_dt_min(xyst_ctx* ctx, double cfl, double dif, double* dt);
/* NodeDiagnostics::precompute sums (NodeDiagnostics.cpp:223-252), out[16]: [0..3] = sum v p^2,
 * v u_c^2; [4..7] = sum v dp^2, v (u-un)_c^2; with an_p (analytic pressure [npoin], or NULL):
 * [8] L2, [9] L1 sums; with an_u (analytic velocity [npoin][3], or NULL): [10..12] L2, [13..15] L1 */
int xyst_chocg_diag(xyst_ctx* ctx, const double* an_p, const double* an_u, double* out);

/* ---- LohCG: artificial-compressibility solver for constant-density flow -------------------
 * Unknowns (p,u,v,w). Edge operators lohner::div/grad/vgrad/flux/rhs (src/Physics/Lohner.cpp:35-1130)
 * on superedge integrals of stride 4 (normal, grad_p.grad_q/(6J); LohCG::domint, LohCG.cpp:407-453)
 * and the solver steps of src/Inciter/LohCG.cpp. div/grad/vgrad/flux are the Chorin operators on the
 * velocity: after xyst_lohcg_mesh_upload the entries xyst_chocg_div (stab 0), _vgrad, _flux, _grad(0),
 * _pinit (divisor 1), _get and xyst_cg_* serve this context too; the entries below are LohCG's own.
 * Several partitions as for ChoCG (LohCG::comgrad/comrhs, LohCG.cpp:1511-1585, besides the shared
 * Chorin entries' exchanges). */
typedef struct xyst_lohcg_params {
  int flux;          /* 0 = damp2, 1 = damp4 (Lohner.cpp:724-914) */
  int stab;          /* tag::stab */
  int stab2;         /* tag::stab2 */
  double stab2coef;
  double mu;         /* mat_dyn_viscosity */
  double soundspeed; /* tag::soundspeed */
} xyst_lohcg_params;
int xyst_lohcg_mesh_upload(xyst_ctx* ctx, size_t npoin, const double* x, const double* y, const double* z,
                           const size_t nsup[3], const size_t* const dsupedge[3],
                           const double* const dsupint[3], size_t ntri, const size_t* triinpoel,
                           const double* vol, const double* v, const xyst_lohcg_params* prm);
/* LohCG::solve BCs (:1617-1622): physics::dirbc on the four unknowns (ndir entries { node },
 * { mask_0..3 }, { value_0..3 }, mask 1 or 2 = set to value), physics::dirbcp (npdir pressure
 * Dirichlet nodes with values, BC.cpp:74-108), symbc and noslipbc on the velocity */
int xyst_lohcg_bc_upload(xyst_ctx* ctx, size_t ndir, const size_t* dirnodes, const int* dirmask,
                         const double* dirval, size_t npdir, const size_t* pdirnodes, const double* pdirval,
                         size_t nsym, const size_t* symbcnodes, const double* symbcnorms,
                         size_t nnoslip, const size_t* noslipbcnodes);
/* Transported scalars after (p,u,v,w) (LohCG::m_u with problem_ncomp = 4 + ns; scalar rows of lohner::grad /
 * adv_damp2 / adv_damp4 / the boundary integral / src, Lohner.cpp:787-793,906-911,1030-1056,1073-1087): call after
 * the mesh upload, before any state or BC upload; then u, rhs rows have 4+ns entries, Dirichlet masks and
 * values 4+ns per node, xyst_lohcg_diag takes `an` with 4+ns columns and returns 16 + 4 ns sums.
 * xyst_chocg_dirbc_values ([ndir][4+ns]) and xyst_chocg_pin serve this context too. */
int xyst_lohcg_scalars(xyst_ctx* ctx, int ns, double diffusivity);
/* source term values at the nodes [npoin][4+ns] (lohner::src), or NULL for none */
int xyst_lohcg_src(xyst_ctx* ctx, const double* S);
int xyst_lohcg_set_u(xyst_ctx* ctx, const double* u /* [npoin][4] */);
int xyst_lohcg_get_u(xyst_ctx* ctx, double* u);
int xyst_lohcg_get_rhs(xyst_ctx* ctx, double* rhs /* [npoin][4] */);
/* dirbc, dirbcp (if pressure != 0), symbc, noslipbc; LohCG::merge :917-921 applies them without dirbcp */
int xyst_lohcg_apply_bc(xyst_ctx* ctx, int pressure);
/* lohner::rhs (for damp4 preceded by the gradient of all unknowns, LohCG::grad :1485-1508 + fingrad) */
int xyst_lohcg_rhs(xyst_ctx* ctx);
/* one RK stage of LohCG::rhs + solve (:1534-1631): rhs, u = un - rk dt rhs/vol, BCs */
int xyst_lohcg_stage(xyst_ctx* ctx, int stage, double rkcoef, double dt);
/* u -= sgrad, velocity BCs (LohCG::psolved :1300-1312); p = CG solution (:1332,1355) */
int xyst_lohcg_project(xyst_ctx* ctx);
int xyst_lohcg_pressure_set(xyst_ctx* ctx);
/* LohCG::dt (:1401-1449): min over nodes of L/(|u|+c) and L^2/max(mu,dif), times cfl */
int xyst_lohcg_dt_min(xyst_ctx* ctx, double cfl, double dif, double* dt);
/* NodeDiagnostics::accompute sums (NodeDiagnostics.cpp:270-372), out[16]: [0..3] = sum v u_c^2,
 * [4..7] = sum v (u-un)_c^2; with an (analytic solution [npoin][4], or NULL): [9..11] L2, [13..15] L1 */
int xyst_lohcg_diag(xyst_ctx* ctx, const double* an, double* out);

/* ---- linear solver of the pressure projection (ChoCG/LohCG) -----------------------------
 * tk::CSR (src/LinearSolver/CSR.hpp:30-107): block CSR exactly as the reference stores it,
 * nrow = npoin*ncomp scalar rows, 1-based ia[nrow+1] / ja[nnz], values a[nnz] (after
 * CSR::dirichlet etc. were applied by the caller). */
int xyst_csr_upload(xyst_ctx* ctx, size_t nrow, size_t ncomp, const size_t* ia, const size_t* ja,
                    const double* a);
/* A context holds two linear solvers (ChoCG::m_cgpre and m_cgmom, ChoCG.cpp:126-140): which = 0
 * (pressure, selected initially) or 1 (momentum). Every xyst_csr_* / xyst_cg_* entry acts on the
 * selected one; the other keeps its matrix, vectors and BCs. */
int xyst_cg_select(xyst_ctx* ctx, int which);
/* New values a[nnz] for the uploaded matrix (same ia/ja, cf. CSR::zero + refill in ChoCG::lhs); the
 * solution vector (next initial guess) is kept. */
int xyst_csr_update(xyst_ctx* ctx, const size_t* ia, const double* a);
/* CSR::mult (CSR.cpp:154-172): r = A x, this partition's own contribution; host vectors. */
int xyst_csr_mult(xyst_ctx* ctx, const double* x, double* r);
/* ConjugateGradients::setup (ConjugateGradients.cpp:105-126 -> residual, pc, initres, normb,
 * rho): x = initial guess, b = right-hand side (complete on every partition); pc 0 = "none",
 * 1 = "jacobi"; slave[i] != 0 marks nodes counted by a lower partition (tk::slave, skipped in
 * dot products), count[i] = tk::count (1 + number of sharers); both NULL in serial.
 * Shared rows are summed/averaged over the lists of xyst_halo_upload. Returns ||b||. */
int xyst_cg_setup(xyst_ctx* ctx, const double* x, const double* b, int pc,
                  const uint8_t* slave, const double* count, double* normb);
/* ConjugateGradients::solve (:558-823): iterate until ||r|| < tol*max(||b||,->1) or maxit. */
int xyst_cg_solve(xyst_ctx* ctx, size_t maxit, double tol, size_t* it, double* normr);
int xyst_cg_get_x(xyst_ctx* ctx, double* x);

/* Counters: kernels launched by this context so far; edges held. */
uint64_t xyst_launch_count(const xyst_ctx* ctx);
uint64_t xyst_nedge(const xyst_ctx* ctx);
/* CUDA-event time (ms) accumulated inside flux-kernel launches, and their count,
 * since the last call with reset != 0 (bench.py's roofline figure). */
int xyst_kernel_time(xyst_ctx* ctx, const char* kernel, int reset, double* ms, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif

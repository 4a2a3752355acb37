// xyst_b200/host/riecg.hpp -- C++ host mirror of the reference's solver-side interface for
// the RieCG path: the same class and member names as src/Inciter/Discretization.{hpp,cpp}
// and src/Inciter/RieCG.{hpp,cpp}, minus Charm++. The members that hold the hot path
// (grad, rhs, solve, BC, dt, diagnostics) call the device C ABI (include/xyst_b200.h);
// the setup members build what the reference builds, with sort/CSR algorithms that scale
// to 10^8 tetrahedra instead of hash maps.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <vector>
#include "mesh.hpp"
#include "xyst_b200.h"

namespace xyst {

//! The fields of the reference's global inciter::g_cfg that this path reads
//! (src/Control/InciterConfig.hpp:161-413; defaults InciterConfig.cpp:1707-1757)
struct Config {
  std::string problem = "userdef";
  std::string flux = "rusanov";
  std::size_t ncomp = 5;
  real gamma = 1.4, p0 = 0.0, cfl = 0.0, dt = 0.0, t0 = 0.0, term = 1.0e+300;
  real alpha = 0.0, kappa = 0.0;             //!< problem_alpha, problem_kappa (manufactured solutions)
  real r0 = 0.0, ce = 0.0;                   //!< problem_r0, problem_ce
  std::array< real, 3 > beta{{ 0, 0, 0 }};   //!< problem_beta
  std::uint64_t nstep = ~0ULL, diag_iter = 1;
  bool stab2 = false;
  real stab2coef = 0.2;
  bool exact_muscl = false;                 //!< device option, see xyst_params
  //! Triangle superedges: 1 = walk the leftover faces in the reference's hash-set order
  //! (identical triangles, needed for parity beyond 1e-9, ~2 us per tet), 0 = walk them in
  //! element order (same edges and integrals, other triangles), -1 = 1 up to 4M tets
  int reforder = -1;
  //! "riecg" or "zalcg": the ZalCG variant (src/Inciter/ZalCG.cpp) shares the setup pipeline
  //! but keeps the global2local node order, carries 4 integrals per edge (normal + J/120)
  //! and advances with one Taylor-Galerkin + flux-corrected-transport stage per step
  std::string solver = "riecg";
  bool fct = true, fctclip = false;
  real fctdif = 1.0;
  std::vector< int > fctsys;                //!< 1-based components limited as a system
  //! steady = true: local time stepping towards a steady state, stop when the L2 residual of
  //! component rescomp falls below residual (Discretization.cpp:1251-1283)
  bool steady = false;
  real residual = 0.0;
  std::uint64_t rescomp = 1;
  //! LaxCG (solver = "laxcg"): gas constant, Turkel parameter, free-stream velocity
  real rgas = 287.052874, turkel = 0.5;
  std::array< real, 3 > velinf{{ 1.0, 1.0, 1.0 }};
  //! ic block of user-defined problems (density, pressure, velocity)
  real ic_density = 0.0, ic_pressure = 0.0;
  std::array< real, 3 > ic_velocity{{ 0, 0, 0 }};
  std::vector< int > bc_sym;
  std::vector< std::vector< int > > bc_dir; //!< { setid, mask_0 .. mask_{ncomp-1} }
  std::vector< int > bc_far;
  real far_density = 0.0, far_pressure = 0.0;
  std::array< real, 3 > far_velocity{{ 0, 0, 0 }};
  std::vector< int > bc_pre;
  std::vector< real > pre_density, pre_pressure;
  //! ChoCG (solver = "chocg", ncomp = 3 velocity unknowns): projection method for constant-
  //! density flow (src/Inciter/ChoCG.cpp). flux = "damp2" | "damp4"; number of RK stages;
  //! pressure solve: iterations, tolerance, preconditioner ("none" | "jacobi"), Dirichlet sets
  //! { setid, mask } with values { setid, value }, Neumann (symmetry) sets, hydrostat node (global id)
  real mu = 0.0, dif = 0.0;
  bool stab = true;
  std::uint64_t rk = 1;
  std::vector< int > bc_noslip;
  std::vector< std::vector< real > > bc_dirval;   //!< { setid, value_0 .. value_{ncomp-1} }
  std::uint64_t p_iter = 10;
  real p_tol = 1.0e-3;
  std::string p_pc = "none";
  std::vector< std::vector< int > > p_bc_dir;
  std::vector< std::vector< real > > p_bc_dirval;
  std::vector< int > p_bc_sym;
  std::uint64_t p_hydrostat = ~0ULL;
  //! ChoCG semi-implicit momentum solve (tag::theta > 0; momentum = { iter, tol, pc })
  real theta = 0.0;
  std::uint64_t mom_iter = 10;
  real mom_tol = 1.0e-3;
  std::string mom_pc = "none";
  //! LohCG (solver = "lohcg", ncomp = 4 unknowns p,u,v,w): artificial compressibility
  //! (src/Inciter/LohCG.cpp); shares the ChoCG keys above, plus the artificial sound speed
  real soundspeed = 1.0;
  std::array< real, 3 > src_location{{ 0, 0, 0 }}; real src_radius = -1.0, src_release_time = 0.0;   // problem "point_src"
  real freezeflow = 1.0, freezetime = 0.0;      // frozen flow: dt multiplier (> 1) once t > freezetime (ChoCG)
};

//! Sum, over all partitions sharing them, `w` doubles per unique shared node (ascending
//! local id); in place. Default: NCCL through the device context.
using HaloSum = std::function< void( int w, std::vector< real >& vals ) >;
//! All-reduce (op 0 = sum, 1 = min) over all partitions, in place. Default: NCCL.
using AllReduce = std::function< void( int op, std::vector< real >& vals ) >;

//! Mesh partition owner, cf. inciter::Discretization
class Discretization {
  public:
    explicit Discretization( const TetMesh& chunk, const Config& cfg );
    const std::vector< std::size_t >& Inpoel() const { return m_inpoel; }
    const std::vector< std::size_t >& Gid() const { return m_gid; }
    const Coords& Coord() const { return m_coord; }
    const std::vector< real >& Vol() const { return m_vol; }   //!< with neighbour contributions
    const std::vector< real >& V() const { return m_v; }       //!< own elements only
    std::map< int, std::vector< std::size_t > >& NodeCommMap() { return m_nodeCommMap; }
    std::size_t lid( std::size_t g ) const;
    real T() const { return m_t; }
    real Dt() const { return m_dt; }
    std::uint64_t It() const { return m_it; }
    real& MeshVol() { return m_meshvol; }
    void vol();                                          //!< Discretization.cpp:618-676 (own part)
    void remap( const std::vector< std::size_t >& newid );   //!< :560-606
    void setdt( real newdt );                            //!< :926-938
    void next();                                         //!< :941-983
    bool finished() const;                               //!< :1251-1262
    void residual( real r ) { m_res = r; }               //!< :1267-1283
    real m_res = 0.0;
    std::vector< std::size_t > sharedNodes() const;      //!< unique shared local ids, ascending
    std::vector< real > m_vol, m_v;
  private:
    const Config& m_cfg;
    std::vector< std::size_t > m_inpoel, m_gid;
    std::vector< std::uint32_t > m_g2l;                  //!< dense global->local (+1), 0 = absent
    std::size_t m_gmin = 0;
    Coords m_coord;
    std::map< int, std::vector< std::size_t > > m_nodeCommMap;  //!< neighbour -> shared GLOBAL ids, ascending
    real m_t, m_dt, m_dtn, m_meshvol = 0.0;
    std::uint64_t m_it = 0;
};

//! cf. inciter::RieCG
class RieCG {
  public:
    RieCG( Discretization& disc, const TetMesh& chunk, const Config& cfg );
    ~RieCG();
    //! create the device context and, for nranks>1, the NCCL communicator
    void attach( int device, int nranks, int rank, const void* ncclid );
    void setComm( HaloSum h, AllReduce a ) { m_halosum = std::move(h); m_allreduce = std::move(a); }
    void setRanks( int nranks, int rank ) { m_nranks = nranks; m_rank = rank; }
    //! host-only part of RieCG::RieCG + feop(): renumber, volumes, edge integrals, superedges
    void prepare();
    //! host part of setup: exchanges (volumes, boundary normals), BC lists, ICs
    void hostSetup();
    //! hostSetup() if not done + device upload + BC(t0)
    void setup();
    real dt();                       //!< RieCG.cpp:787-852 incl. the global min reduction
    void advance( real newdt );      //!< :854-869
    void grad();                     //!< :871-893 (+ comgrad, normalisation)
    void rhs();                      //!< :918-966 (+ comrhs)
    void solve();                    //!< :992-1057 update + BC for the current stage
    void BC();                       //!< :764-785
    //! one full time step through the fused device path; returns false when finished
    bool step( std::vector< real >* diagrow );
    //! NodeDiagnostics::rhocompute + Transporter::rhodiagnostics
    std::vector< real > diagnostics();
    std::vector< real > solution();  //!< m_u, npoin x ncomp
    void setSolution( const std::vector< real >& u );
    xyst_ctx* ctx() { return m_ctx; }
    Discretization& Disc() { return m_disc; }

    // streamable data, same names/meaning as the reference's members
    std::map< int, std::vector< std::size_t > > m_bface;
    std::vector< std::size_t > m_triinpoel;
    std::array< std::vector< std::size_t >, 3 > m_dsupedge;
    std::array< std::vector< real >, 3 > m_dsupint;
    std::vector< std::uint8_t > m_besym;
    std::vector< std::size_t > m_dirbcmasks, m_symbcnodes, m_farbcnodes, m_prebcnodes;
    std::vector< real > m_symbcnorms, m_farbcnorms, m_prebcvals;
    // ChoCG: Dirichlet values, pressure Dirichlet masks/values, no-slip nodes (ChoCG.hpp members of
    // the same names), the pressure Poisson matrix in the reference's CSR form (tk::CSR, 1-based)
    std::vector< real > m_dirbcval, m_dirbcvalp;
    std::vector< std::size_t > m_dirbcmaskp, m_noslipbcnodes;
    std::vector< std::size_t > m_plhs_ia, m_plhs_ja;
    std::vector< real > m_plhs_a;
    std::vector< real > m_mlhs_a;    //!< momentum matrix values of the last step (theta > 0), block CSR
    std::size_t m_pit = 0;           //!< iterations of the last pressure solve
    std::size_t m_mit = 0;           //!< iterations of the last momentum solve (theta > 0)
    std::vector< real > choGet( const char* what, std::size_t width );
    //! host-side views for the tests: pressure Dirichlet (local node, value) pairs and the momentum
    //! solve's Dirichlet rows after the union over the partitions sharing a node
    //! (computed at the first call: the union is a collective operation, the getter is called twice per array)
    std::vector< real > pressureBC() { if (!m_pbcview) { choPressureSetup(); m_pbcview = true; } std::vector< real > r;
      for (const auto& [p,v] : m_pbc) { r.push_back( static_cast< real >( p ) ); r.push_back( v ); } return r; }
    std::vector< std::size_t > momentumBCRows() { if (!m_mrowview) { choMomRows(); m_mrowview = true; } return m_mbcrows; }
    bool pendingDiag() const { return !m_lastdiag.empty(); }   //!< row of a run that ended during setup (nstep = 1)
    std::vector< real > m_u0;        //!< initial condition (npoin x ncomp)
    std::vector< double > timings;   //!< seconds spent in the setup phases (for reporting)
    bool m_finished = false;
  private:
    void renumber();                 //!< RieCG.cpp:82-100
    void boundaryFaces( const TetMesh& chunk );    //!< Partitioner.cpp:539-626 per partition
    void domint( const EdgeCSR& edges, std::vector< real >& d ) const;   //!< :339-382, ZalCG.cpp:354-400
    void domsuped( const EdgeCSR& edges, const std::vector< real >& d ); //!< :620-736
    void bndint();                   //!< :281-337
    void setupBC();                  //!< :109-245 (after normals are known)
    void uploadHalo();
    void evalDirvals( real t );      //!< physics::dirbc values = IC at the BC nodes at time t (BC.cpp:57-66)
    void evalSrcCentroids( real t, std::vector< real >& sc );   //!< problems::SRC at the tet centroids (kozak::rhs)
    void evalSrc( real t );          //!< problems::SRC at the nodes at time t (riemann::src, Riemann.cpp:880-907)
    bool m_pbcview = false, m_mrowview = false;
    bool m_pinned = false;           //!< point-source nodes handed to the device
    bool m_timedep = false;          //!< IC / source depend on time: BC values per stage, source per step
    bool m_haloup = false;
    Discretization& m_disc;
    const Config& m_cfg;
    std::map< int, std::vector< std::size_t > > m_sidetri;  //!< chunk side sets (global ids)
    std::map< int, std::map< std::size_t, std::array< real, 4 > > > m_bnorm;  //!< set -> local node -> normal
    std::set< std::size_t > m_symbcnodeset, m_farbcnodeset;
    xyst_ctx* m_ctx = nullptr;
    int m_nranks = 1, m_rank = 0;
    int m_stage = 0;
    HaloSum m_halosum;
    AllReduce m_allreduce;
    bool m_nccl_reduce = false;          // reductions by the device library itself (NCCL), not by caller-supplied hooks
    bool m_hostready = false;
    bool m_zal = false;
    bool m_lax = false;                    //!< LaxCG: same setup as RieCG, preconditioned update
    bool m_koz = false;                    //!< KozCG: element-based, no edge integrals
    bool m_cho = false;                    //!< ChoCG: stride-5 integrals, projection steps (chocg.cpp)
    int m_np = 0;                          //!< ChoCG::m_np
    real m_freezeflow = 1.0;               //!< ChoCG / KozCG / ZalCG::m_freezeflow
    std::array< std::vector< std::size_t >, 2 > m_zedge;   //!< end nodes of the device's edge slots (ZalCG source term)
    bool m_initial = true;                 //!< Discretization::Initial()
    std::map< std::size_t, real > m_pbc;   //!< pressure Dirichlet node -> value of the first solves
    std::vector< real > m_neubc, m_prhs, m_psol, m_lastdiag;
    void choSetupBC();                     //!< ChoCG::setupDirBC :210-300, streamable :655-682
    void choPrelhs();                      //!< ChoCG::prelhs :146-188 on tk::CSR( psup )
    void choLhs();                         //!< ChoCG::lhs :1433-1477 -> momentum matrix on the device
    std::vector< std::size_t > m_mlhs_ia, m_eoff, m_mbcrows;   //!< block-CSR row offsets; elements surrounding points; momentum BC rows
    std::vector< std::uint64_t > m_esup;
    bool m_mlhsup = false;
    void choPressureSetup();               //!< pressure BC values, Neumann vector, rhs override of pinit
    void choMomRows();                     //!< Dirichlet rows of the momentum solve (theta > 0)
    void choDirvals( real t, std::vector< real >& dv ) const;   //!< physics::dirbc values at time t
    void choBCtime( real t );              //!< ... handed to the device for the next BC application
    void choSetup();                       //!< device upload + ChoCG::merge :816-837 onwards
    bool choStep( std::vector< real >* diagrow );
    void choPinit();                       //!< :1025-1125
    void choPsolve();                      //!< :1127-1142
    void choPsolved( std::vector< real >* diagrow );   //!< :1194-1253
    std::vector< real > choDiag();         //!< NodeDiagnostics::precompute + Transporter::prediagnostics
    bool m_loh = false;                    //!< LohCG: stride-4 integrals with the Laplacian term (lohcg.cpp)
    void lohSetup();                       //!< device upload + LohCG::merge :909-931 onwards
    bool lohStep( std::vector< real >* diagrow );
    void lohPinit();                       //!< :1127-1219
    void lohPsolved();                     //!< :1289-1363
    std::vector< real > lohDiag();         //!< NodeDiagnostics::accompute + Transporter::acdiagnostics
    std::size_t m_stride = 3;
    real m_ownvol = 0.0;
    std::vector< real > m_dirvals, m_src;
};

} // xyst::

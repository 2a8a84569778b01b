/* hot_b200 — C ABI of the B200-native implicit-MPM hot path (drop-in for the HOT/Ziran operator surface).
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ / Eigen / torch types cross this boundary;
 *  - every call returns 0 on success or a negative code; hot_last_error(h) gives the message
 *    (the reference throws std::runtime_error from ZIRAN_ASSERT, Lib/Ziran/CS/Util/Debug.h:19-42);
 *  - scalar type is double (the reference binary hard-codes T=double, Projects/multigrid/main.cpp:12);
 *    3x3 matrices are column-major like Eigen (Lib/Ziran/CS/Util/Forward.h:10-13);
 *    vectors over grid DOFs are "TVStack" layout: n_nodes x (x,y,z) contiguous;
 *  - "host" pointers are caller-owned host buffers in ORIGINAL particle order (the order of the
 *    reference's particles.X.array); the library keeps its own device-resident state between calls;
 *  - one handle per GPU, one caller thread per handle (the reference is single-instance too);
 *  - there is NO CPU fallback: hot_create fails when no CUDA device is usable.
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef HOT_B200_H
#define HOT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hot_sim hot_sim; /* opaque; replaces MpmSimulationBase<double,3> + MpmGrid + force helper state */

/* ---- lifetime ---------------------------------------------------------------------------------- */
/* MpmSimulationBase ctor / MpmGrid ctor (Lib/MPM/MpmSimulationBase.cpp:60-118, Lib/MPM/MpmGrid.h:121-136).
 * device < 0 selects the current CUDA device. */
hot_sim* hot_create(double dx, double apic_rpic_ratio, double cfl, int device);
void hot_destroy(hot_sim* h);
const char* hot_last_error(hot_sim* h);
/* launch all work of this handle on an externally owned cudaStream_t (0 = legacy default stream) */
int hot_set_stream(hot_sim* h, void* cuda_stream);
int hot_synchronize(hot_sim* h);
/* number of CUDA kernels this handle has launched so far (bench.py's gpu_launches) */
long long hot_launch_count(hot_sim* h);

/* Per-kernel-class CUDA-event timing on the handle's stream (replaces ZIRAN_TIMER / ScopedTimer,
 * Lib/Ziran/CS/Util/Timer.h:34-58, and the per-level V-cycle table, MultigridPreconditioner.h:417-419).
 * enable: 0 off, 1 on, 2 on + reset.  hot_get_timings synchronises, fills up to n classes, returns the class count. */
int hot_timing(hot_sim* h, int enable);
int hot_get_timings(hot_sim* h, int n, double* ms_total, long long* counts);
const char* hot_timing_name(int kernel_class);

/* ---- SPGrid addressing (Lib/SPGrid/Core/SPGrid_Mask.h) -- device evaluation, host buffers ---------- */
/* Linear_Offset :150-166 */
int hot_linear_offset(hot_sim* h, long n, const int* ijk, unsigned long long* out);
/* LinearToCoord :176-189 */
int hot_linear_to_coord(hot_sim* h, long n, const unsigned long long* off, int* ijk);
/* Packed_Add :237-245 */
int hot_packed_add(hot_sim* h, long n, const unsigned long long* a, const unsigned long long* b, unsigned long long* out);
/* the same three operations for SPGrid_Mask of GridState<float,3> (64-byte record, data_bits 6, 4x4x4 pages: Lib/MPM/MpmGrid.h:14-34).
 * Addressing only: the compute kernels follow the reference binary's T = double (Projects/multigrid/main.cpp:12). */
int hot_linear_offset_f32(hot_sim* h, long n, const int* ijk, unsigned long long* out);
int hot_linear_to_coord_f32(hot_sim* h, long n, const unsigned long long* off, int* ijk);
int hot_packed_add_f32(hot_sim* h, long n, const unsigned long long* a, const unsigned long long* b, unsigned long long* out);

/* ---- particles --------------------------------------------------------------------------------- */
/* Upload of particles.X/V/mass ("P","V","m": Lib/Ziran/Math/Geometry/Particles.h:8-45), the APIC matrix C,
 * the F-based helper's F / element measure (Lib/MPM/Force/FBasedMpmForceHelper.h:20-47) and the
 * CorotatedIsotropic parameters mu, lambda (Lib/Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h:64-73).
 * AoS, original order. */
int hot_set_particles(hot_sim* h, long n, const double* X, const double* V, const double* mass, const double* C,
    const double* F, const double* vol, const double* mu, const double* lambda);
/* read back particle state in original order; any pointer may be NULL */
int hot_get_particles(hot_sim* h, double* X, double* V, double* C, double* F, double* gradV);
/* Pipelined exchange of the particle STATE (X, V, C, F; AoS, original order) with host buffers for a caller that keeps the
 * particles on the host between steps (the reference owns them in DataManager arrays, Lib/Ziran/CS/DataStructure/DataManager.h;
 * mass, volume and material parameters stay resident from hot_set_particles).  The transfers run on the library's own copy
 * streams: hot_upload_state_async of step k+1 may be issued as soon as hot_commit_state of step k has been called, and
 * hot_download_state_async returns at once, so both PCIe directions work while a step computes.
 *   upload:   the host arrays must stay valid until hot_commit_state (which makes the handle's stream wait for the copy and
 *             scatters the rows into the current particle order; the next call must be hot_sort_and_activate);
 *   download: the host arrays are complete after hot_wait_download. */
int hot_upload_state_async(hot_sim* h, const double* X, const double* V, const double* C, const double* F);
int hot_commit_state(hot_sim* h);
int hot_download_state_async(hot_sim* h, double* X, double* V, double* C, double* F);
int hot_wait_download(hot_sim* h);
long hot_num_particles(hot_sim* h);

/* ---- a5: MpmSimulationBase::sortParticlesAndPolluteGrid (Lib/MPM/MpmSimulationBase.cpp:1066-1137) ---- */
int hot_sort_and_activate(hot_sim* h);
long hot_num_groups(hot_sim* h);
long hot_num_pages(hot_sim* h);
/* particle_sorter, particle_order, particle_base_offset (MpmSimulationBase.h:98-112) */
int hot_get_sort(hot_sim* h, unsigned long long* sorter, int* order, unsigned long long* base_offset);
/* particle_group (first,last inclusive) and block_offset */
int hot_get_groups(hot_sim* h, int* first, int* last, unsigned long long* block_offset);
/* page_map->Get_Blocks() in first-Set order (Lib/SPGrid/Core/SPGrid_Page_Map.h:61-96) */
int hot_get_pages(hot_sim* h, unsigned long long* offsets);

/* ---- a6+a7: MpmSimulationBase::particlesToGrid (Lib/MPM/MpmSimulationBase.cpp:461-533, 611-656),
 *      MpmGrid::getNumNodes (Lib/MPM/MpmGrid.h:148-161) ------------------------------------------------ */
int hot_p2g(hot_sim* h, int* n_nodes);
int hot_num_nodes(hot_sim* h);
/* grid read-back in page-list order x element order: n_pages * elements_per_block entries */
int hot_get_grid(hot_sim* h, long long* idx, double* m, double* v);
/* node coordinate per DOF id (ImplicitSolver.h id2coord) */
int hot_get_id2coord(hot_sim* h, int* coord);
/* buildMassMatrix (Lib/MPM/MpmSimulationBase.cpp:817-826) */
int hot_get_mass_matrix(hot_sim* h, double* mass);

/* ---- a23: constructNewVelocityFromNewtonResult + gridToParticles + evolveStrain
 *      (Lib/MPM/MpmSimulationBase.cpp:891-901, 930-1042; Lib/MPM/Force/FBasedMpmForceHelper.cpp:100-114) --- */
int hot_set_dv(hot_sim* h, const double* dv);
int hot_g2p(hot_sim* h, double dt, int* flags /* [0] faster than dx, [1] faster than cfl*dx/2 */);

/* ---- plasticity: MpmSimulationBase::applyPlasticity (Lib/MPM/MpmSimulationBase.cpp:1044-1064), run by hot_g2p right after
 * evolveStrain like :1039-1041.  model 0 none; 1 VonMisesFixedCorotated::projectStrain (Lib/Ziran/Physics/PlasticityApplier.cpp:94-131),
 * params = {yield_stress}; 2 SnowPlasticity::projectStrain (:16-50), params = {psi, theta_c, theta_s, min_Jp, max_Jp}
 * (defaults 10, 2e-2, 7.5e-3, 0.6, 20: PlasticityApplier.h:61), hardening mu / lambda and Jp kept per particle on the device;
 * 3 Drucker-Prager, an EXTENSION (the reference ships no sand model; BASELINE's sand-column configuration names it): the Hencky-strain
 * return mapping of Klar et al. 2016 on the singular values at the same hook, params = {friction angle in degrees, cohesion (log-strain
 * shift, 0 for dry sand)}; parity is against the closed form in numpy (tests/test_oracle_plasticity.py), not against reference code. */
int hot_set_plasticity(hot_sim* h, int model, const double* params);
int hot_apply_plasticity(hot_sim* h);
int hot_get_plastic_state(hot_sim* h, double* Jp, double* mu, double* lambda);
int hot_set_plastic_state(hot_sim* h, const double* Jp);

/* ---- one object over the GPUs of a box (no counterpart in the reference, which is single-process: SURVEY 2a / 8e) ------
 * One process and one handle per GPU.  Every rank gives ITS particles to hot_set_particles (any distribution is correct; compact
 * slabs keep the seams small) and is locally a complete single-GPU object: own sort, own page list (its particles' pages + the
 * +1 neighbours of MpmSimulationBase.cpp:1104-1124), own DOF numbering.  Pages activated by two or more ranks are SHARED: after
 * every particle->grid scatter the partial sums of the shared pages travel between the sharing ranks only and are added in
 * ascending rank order on every sharer, so shared nodes carry bit-identical values on all their ranks and the gathers need no
 * communication.  DOF vectors are local (shared nodes replicated); dots / norms count a shared node on its lowest-ranked sharer
 * and all-reduce 1-3 scalars.  Node ids differ between ranks and from a single-GPU run: identify nodes by hot_get_id2coord.
 * Partitioned solver paths: matrix-free PN-PCG (-lsolver 2 --matfree) as is; the assembled-matrix / multigrid / L-BFGS path with
 * hot_set_ghost_ring (below).
 *
 * Transport, one of:
 *  - NCCL inside the library: rank 0 calls hot_comm_unique_id, the caller distributes the 128 bytes (MPI / torch.distributed / a
 *    file), every rank calls hot_comm_init_nccl.  The shared-page exchange is one grouped ncclSend / ncclRecv per neighbour on the
 *    handle's stream; NCCL is resolved with dlopen("libnccl.so.2") at that point, so the library loads without it.
 *  - the caller's callbacks (hot_set_partition): all pointers are device pointers, the calls are ordered after everything enqueued
 *    on the handle's stream and must be complete (or enqueued on that stream) when they return; non-zero = failure. */
typedef struct hot_transport {
    void* user;
    int (*all_reduce)(void* user, double* dev, long count, int op);                                  /* in place; op 0 sum, 1 max */
    int (*all_gather)(void* user, const void* dev_send, void* dev_recv, long bytes_per_rank);       /* recv = world x bytes_per_rank, by rank */
    int (*neighbor_exchange)(void* user, int n_peers, const int* peers, double* const* send, double* const* recv,
        const long* count);                                                                      /* count[j] doubles to AND from peers[j] */
} hot_transport;
int hot_comm_unique_id(void* id128);
int hot_comm_init_nccl(hot_sim* h, int rank, int world, const void* id128);
int hot_set_partition(hot_sim* h, int rank, int world, const hot_transport* transport);
/* {rank, world, neighbour ranks, shared local pages, pages exchanged per scatter (sum over neighbours), nodes this rank counts
 * in reductions, nodes of the whole object (the last two -1 before hot_p2g), local particles} */
int hot_get_partition(hot_sim* h, long* out8);
/* Constitutive model of the F-based force helper: 0 = CorotatedIsotropic (fixed corotated, the reference's model,
 * Lib/Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h), 1 = neo-Hookean, an EXTENSION in the same SvdBasedIsotropicHelper framework
 * (psi = mu/2 (|F|^2 - 3) - mu log J + lambda/2 log^2 J; the reference ships no such model, BASELINE's box-drop configuration names it).
 * Parity for model 1 is against the oracle's restatement, which is pinned by finite differences and numpy (tests/test_oracle_force.py). */
int hot_set_constitutive_model(hot_sim* h, int model);
/* Ghost ring for the assembled-matrix / multigrid path of a partitioned object (SquareMatrix rows reach two nodes around a node,
 * Projects/multigrid/ImplicitSolver.h:465-468): with on != 0 a rank also holds every page of the 27-neighbourhood of its shared
 * pages that some rank activates.  Needed by hot_build_matrix / hot_build_mg / hot_vcycle / the -lsolver 2 and 3 solves with a
 * matrix when world > 1; the transfers alone (hot_p2g / hot_g2p) and the matrix-free solver run without it (fewer pages to exchange).
 * Takes effect with the next hot_sort_and_activate. */
int hot_set_ghost_ring(hot_sim* h, int on);
/* host logic of the ghost ring, callable without a device (CPU tests): ghost pages of `rank` (ascending; out may be NULL to count),
 * and the authority rank of pages (the rank whose rows count: lowest rank that activates the page and the other half of its 4^3
 * Gauss-Seidel block; inputs as for hot_share_tables, the BASE lists without ghost pages) */
int hot_halo_pages(int rank, int world, int max_pages, const int* counts, const unsigned* all_pids, int* n_out, unsigned* out);
int hot_page_authority(int world, int max_pages, const int* counts, const unsigned* all_pids, int n, const unsigned* pids, int* auth);
/* how the shared pages travel after the last hot_sort_and_activate: 0 single rank, 1 grouped ncclSend / ncclRecv, 2 the caller's
 * callbacks, 3 peer memory (with hot_comm_init_nccl, when every rank can map every other rank's receive arena through cudaIpc:
 * the pack kernel stores the partial sums straight into the neighbours' HBM over NVLink and raises a flag there, the unpack kernel
 * waits on the flags - no collective kernel and no host work in between; HOT_XCHG=nccl keeps transport 1) */
int hot_get_transport(hot_sim* h);
/* The host logic behind the shared-page tables, callable without a device (CPU tests of the N > 1 path): from every rank's
 * ascending page-id list (all_pids: world rows of max_pages, counts[r] valid) builds this rank's neighbour list, exchange list
 * (x_slot: local slots, per neighbour in ascending page id) and, per shared local page, the contributions in ascending rank order
 * (sh_entry: index into the exchange list, -1 = own partial) + whether this rank is the page's lowest sharer (sh_owned).
 * Output arrays are caller-allocated for the worst case (see api.cu). */
int hot_share_tables(int rank, int world, int max_pages, const int* counts, const unsigned* all_pids, const int* slot_sorted, int* n_nbr,
    int* nbr_rank, long* nbr_off, long* nbr_cnt, int* n_x, int* x_slot, int* n_sh, int* sh_slot, int* sh_ptr, int* sh_entry, int* sh_owned);
/* synchronous raw copies on the handle's stream, for transport callbacks written without a CUDA binding */
int hot_memcpy_d2h(hot_sim* h, void* host, const void* dev, long bytes);
int hot_memcpy_h2d(hot_sim* h, void* dev, const void* host, long bytes);

/* ---- force model: the operator surface ImplicitSolverObjective drives (Projects/multigrid/ImplicitSolver.h) ------- */
/* simulation.dt / simulation.gravity (MpmSimulationBase.h:69-131) */
int hot_set_dt_gravity(hot_sim* h, double dt, const double* gravity3);
/* CorotatedIsotropic::project, the PSD clamp of the SVD-space Hessian blocks (CorotatedIsotropic.h:60,139-143) */
int hot_set_project(hot_sim* h, int project);
/* a8 output of buildInitialDvAndVnForNewton (Lib/MPM/MpmSimulationBase.cpp:1139-1184), evaluated by the host from its
 * collision objects: the CollisionNode table {node_id, P, R, Rinv, shouldRotate} (CollisionObject.h:16-45; 3x3
 * column-major, any of P/R/Rinv/slip may be NULL) and the collider velocity difference dv_bc (NULL = 0).
 * mode 0: project(v) = P v on every BC node (MultigridSimulation.h:104-125 default); mode 1: HOTSettings::boundaryType
 * == 1 && systemBCProject: rotate slip nodes with R, zero component 0 (slip) or the whole node (sticky).
 * Also sets the Newton initial guess dv = gravity*dt on free nodes and dv_bc on BC nodes (:1177-1180). */
int hot_set_bc(hot_sim* h, int mode, int n_bc, const int* node_id, const double* P, const double* R, const double* Rinv,
    const int* slip, const double* dv_bc);
int hot_get_dv(hot_sim* h, double* dv);
/* a8 on the device: AnalyticCollisionObject (Lib/Ziran/Math/Geometry/CollisionObject.h:46-120, .cpp:384-452) over an analytic level set
 * (AnalyticLevelSet.h:122-310), as plain data.  Object transform x = R s X + b with rates omega, dsdt, dbdt (the caller's
 * updateState(t) of MultigridSimulation.h:292-295 rewrites them between steps and calls hot_set_colliders again).  3x3 column-major. */
enum { HOT_COLLIDER_STICKY = 1, HOT_COLLIDER_SLIP = 2, HOT_COLLIDER_SEPARATE = 3, HOT_COLLIDER_GHOST = 4 };          /* COLLISION_OBJECT_TYPE */
enum { HOT_SHAPE_HALFSPACE = 0, HOT_SHAPE_SPHERE = 1, HOT_SHAPE_BOX = 2, HOT_SHAPE_CAPPED_CYLINDER = 3 };
typedef struct hot_collider {
    int type, shape;
    double friction;
    double p[8];        /* HalfSpace: origin[3], outward_normal[3] | Sphere: center[3], radius | AnalyticBox: half_edges[3] |
                           CappedCylinder (axis y): radius, height */
    double shape_R[9];  /* AnalyticBox / CappedCylinder own rigid transform: X_primitive = shape_R^-1 (X - shape_b).  From the reference's constructor
                           argument q = <w,x,y,z>: AnalyticBox rotates by q normalised; CappedCylinder by the quaternion <w = q[3], x = q[0], y = q[1],
                           z = q[2]> normalised (its 4-vector reaches Eigen::Quaternion in storage order, AnalyticLevelSet.h:248-251) - hot_b200_host.hpp
                           builds both that way */
    double shape_b[3];
    double R[9], s, b[3];            /* object transform */
    double omega[3], dsdt, dbdt[3];  /* and its rates */
} hot_collider;
int hot_set_colliders(hot_sim* h, int n, const hot_collider* objects);
/* buildInitialDvAndVnForNewton (Lib/MPM/MpmSimulationBase.cpp:1139-1184) with the device-resident objects: multiObjectCollision
 * (CollisionObject.cpp:108-149) per grid node, the CollisionNode table {node, P = I - K K^T, R, R^-1, shouldRotate} in node order and
 * the Newton initial guess (dv = gravity dt on free nodes, collider velocity difference on collision nodes), all on the device;
 * `mode` as in hot_set_bc.  Replaces the host evaluation + hot_set_bc; hot_get_bc reads the table back. */
int hot_build_bc(hot_sim* h, int mode, int* n_bc);
/* the current BC table (from hot_set_bc or hot_build_bc); any pointer may be NULL; returns the number of rows through *n_bc */
int hot_get_bc(hot_sim* h, int* n_bc, int* node_id, double* P, double* R, double* Rinv, int* slip);
/* CorotatedIsotropic<T,3> as an operator on n deformation gradients (Lib/Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h:78-230):
 * updateScratch (QR-SVD conventions of ImplicitQRSVD.h: det U = det V = +1, sigma sorted, sign on the last) + psi + firstPiola +
 * firstPiolaDifferential (dP for the given dF) + firstPiolaDerivative (dense 9x9 dPdF, column-major, index ij = i + 3 j), with or
 * without the PSD projection of --project.  3x3 blocks are column-major like Eigen's; every output pointer may be NULL. */
int hot_corotated_eval(hot_sim* h, long n, const double* F, double mu, double lambda, int project, const double* dF, double* psi,
    double* P, double* dP, double* dPdF, double* U, double* sigma, double* V);
/* FBasedMpmForceHelper::backupStrain / restoreStrain (Lib/MPM/Force/FBasedMpmForceHelper.cpp:25-44) */
int hot_backup_strain(hot_sim* h);
int hot_restore_strain(hot_sim* h);
/* ImplicitSolverObjective::updateState (ImplicitSolver.h:237-252): moveNodes(dv) (dv NULL = keep the device dv),
 * MpmForceBase::updatePositionBasedState (MpmForceBase.cpp:308-328); energy (nullable) = totalEnergy (:254-275) */
int hot_update_state(hot_sim* h, const double* dv, double* energy);
/* scratch_stress = vol P Fn^T and the trial F, original particle order (FBasedMpmForceHelper.cpp:72-97) */
int hot_get_stress(hot_sim* h, double* vPFnT, double* F);
/* FBasedMpmForceHelper::totalEnergy (Lib/MPM/Force/FBasedMpmForceHelper.cpp:116-136): sum_p vol psi(F_p) of the last updateState */
int hot_strain_energy(hot_sim* h, double* energy);
/* FBasedMpmForceHelper::Fn (FBasedMpmForceHelper.h:29): the strain saved by backupStrain, original particle order, 9 per particle */
int hot_get_strain_backup(hot_sim* h, double* Fn);
/* ImplicitSolverObjective::computeResidual (ImplicitSolver.h:128-155) */
int hot_compute_residual(hot_sim* h, double* residual);
/* objective.project (MultigridSimulation.h:104-125), in place */
int hot_project(hot_sim* h, double* v);
/* ImplicitSolverObjective::multiply with --matfree (ImplicitSolver.h:741-763): b = M x + dt^2 K x */
int hot_hessian_apply_mf(hot_sim* h, const double* x, double* b);
/* MpmSimulationBase::addScaledForces(scale, f) (Lib/MPM/MpmSimulationBase.cpp:829-833): f += scale * (elastic nodal force) */
int hot_add_scaled_forces(hot_sim* h, double scale, double* f);
/* MpmSimulationBase::addScaledForceDifferentials(scale, x, f) (:835-840): f += scale * df(x), df = -K x */
int hot_add_scaled_force_differentials(hot_sim* h, double scale, const double* x, double* f);
/* ImplicitSolverObjective::evaluatePerNodeCNTolerance (ImplicitSolver.h:667-696); tol may be NULL (kept on device) */
int hot_eval_cn_tolerance(hot_sim* h, double eps, double dt, double* tol);

/* ---- a15: assembled system (Projects/multigrid/ImplicitSolver.h:470-603) ------------------------------------------ */
/* buildMatrix<bcproject>: block rows of 125 slots, slot (dx+2)*25+(dy+2)*5+(dz+2) = neighbour at coord_i - d
 * (linearOffset :465-468), M + dt^2 K at the current trial state, BC-projected when bcproject (--bcproject). */
int hot_build_matrix(hot_sim* h, int bcproject);
/* entryCol (n x 125) / entryVal (n x 125 x 9, blocks column-major).  Structurally empty slots carry the row's own index
 * with a zero block (the reference pads them with column 0 / 1 and a zero block, :568-571: same operator). */
int hot_get_matrix(hot_sim* h, int* entryCol, double* entryVal);
/* buildDiagonal (:605-665): inverse diagonal blocks (Ainv 1) / entries (Ainv 0) of the matrix-free operator */
int hot_build_diagonal(hot_sim* h, int Ainv, double* diag_inv /* 9 per node, nullable */);

/* ---- a16-a20: Galerkin multigrid (Projects/multigrid/MultigridPreconditioner.h) ------------------------------------ */
/* MultigridBuilder::build :553-703 with HOTSettings {levelCnt, smoother, coarseSolver, Ainv, times, levelscale, topomega}
 * (= -mg_level -smoother -coarseSolver -Ainv -mg_times -mg_scale -mg_jomega; integer codes 0 Jacobi, 1 optimal Jacobi,
 * 2 PCG, 5 GS as in setup_logic :480-522; 6 Chebyshev / 7 IC are not provided) */
int hot_build_mg(hot_sim* h, int levels, int smoother, int coarse_solver, int Ainv, int times, int levelscale, double topomega);
int hot_mg_levels(hot_sim* h);
/* SquareMatrix::estimate2norm (Projects/multigrid/SquareMatrix.h:375-475): power iteration for the 2-norm of A_level, out = {lMax,
 * lMin = lMax / 30}; the Chebyshev smoother (-smoother 6, MultigridPreconditioner.h:227-264) uses them.  hot_build_mg runs it on the
 * levels where Chebyshev will be applied (:610-611, :682-683); the start vector is a fixed +-1 pattern (the reference seeds rand()
 * with the wall clock). */
int hot_estimate_2norm(hot_sim* h, int level, double* lmax_lmin);
int hot_get_level_dofs(hot_sim* h, int* dofs);
int hot_get_level_coords(hot_sim* h, int level, int* coord3);
/* number of structurally non-zero 3x3 blocks of A_level (the reference stores 125 per row regardless) */
int hot_level_nnz_blocks(hot_sim* h, int level, long long* nnzb);
/* kind 0: system matrix A_l (colsize 125, val 9 per entry); kind 1: prolongation P_l (colsize 8) and kind 2: restriction
 * R_l = P_l^T (colsize 32, 27 used) with ONE scalar weight per entry (the reference stores w * I3).  col / val nullable. */
int hot_get_level_matrix(hot_sim* h, int level, int kind, int* colsize, int* col, double* val);
/* SquareMatrix::diagonalVal and its inverse (diagonalBlock / diagonalEntry by Ainv), SquareMatrix.h:301-324 */
int hot_get_level_diagonal(hot_sim* h, int level, double* diagonalVal, double* diagonalInv);
/* markColors :582-605 as a sweep schedule: nodes in (colour, first-seen block, first-seen node) order, block b =
 * seq[block_start[b] .. block_start[b+1]), colour c = blocks [color_first_block[c], color_first_block[c+1]) */
int hot_get_gs_schedule(hot_sim* h, int level, int* n_blocks, int* color_first_block9, int* seq, int* block_start);
/* SquareMatrix::multiply (SquareMatrix.h:477-487) / SparseMatrix::multiply (SparseMatrixFast.h:60-73) */
int hot_spmv(hot_sim* h, int level, const double* x, double* b);
/* SparseMPMMatrix::transposeMultiply / multiply (MPMMultigridMatrix.h:63-70) between level and level+1 */
int hot_restrict(hot_sim* h, int level, const double* fine, double* coarse);
int hot_prolong(hot_sim* h, int level, const double* coarse, double* fine);
/* smoothFunc(u, r, du, dAu, A, iterations, tolerance) (:68-73); kind = the -smoother code; u, r updated in place;
 * initial_residual (nullable) = initialResiduals[level], the reference of the PCG stopping test (:197-209) */
int hot_smooth(hot_sim* h, int level, int kind, double* u, double* r, int iterations, double tolerance, const double* initial_residual);
/* MultigridOperator::operator() :362-421 */
int hot_vcycle(hot_sim* h, const double* in, double* out);
/* per-level [smooth, restrict, prolongate, merge] milliseconds of the last hot_vcycle (the table of :417-419), 10 x 4 */
int hot_vcycle_timing(hot_sim* h, double* ms40, int* coarse_cg_iters);
/* `reps` device-resident V-cycles on the right-hand side of the last hot_vcycle; total milliseconds (CUDA events) */
int hot_vcycle_bench(hot_sim* h, int reps, double* ms_total);
/* `reps` device-resident applications of one operator (measurement hook of bench.py): op 0 matrix-free Hessian apply,
 * 1 block SpMV on `level`, 2 updateState, 3 computeResidual, 4 one -smoother call on `level`, 5 hot_build_matrix,
 * 6 hot_build_mg, 7 one -coarseSolver call on `level` (right-hand side: the level's restricted initial residual of the
 * last hot_vcycle, so cg_smooth, MultigridPreconditioner.h:190-225, has to iterate; hot_vcycle_timing returns its count);
 * total milliseconds (CUDA events on the handle's stream) */
int hot_op_bench(hot_sim* h, int op, int level, int reps, double* ms_total);

/* ---- a21, a22, a24: solvers and the time step ----------------------------------------------------------------------- */
/* HOTSettings (Projects/multigrid/Configurations.h:18-42) + the solver limits MultigridSimulation / the objective hard-code
 * (MultigridSimulation.h:97-99: newton(objective, 1, 3), lbfgs(objective, 1, 10000); ImplicitSolver.h:78: cg(10000)).
 * Field names follow the command-line flags of Projects/multigrid/main.cpp:40-84. */
typedef struct hot_solver_options {
    int lsolver;        /* -lsolver: 1 = Newton + MINRES, 2 = Newton + inexact PCG (PN-PCG / PN-MGPCG), 3 = L-BFGS around the V-cycle (HOT) */
    int matfree;        /* --matfree: matrix-free apply + block-Jacobi (lsolver 2 only, README:13-15) */
    int project;        /* --project: PSD-project the particle Hessians */
    int bcproject;      /* --bcproject: BC-project the assembled system */
    int linesearch;     /* --linesearch */
    int usecn;          /* --usecn: characteristic-norm stopping test */
    int adaptive_h;     /* --adaptiveH: rebuild the Hessian approximation every 16 L-BFGS iterations */
    int mg_level;       /* -mg_level */
    int mg_times;       /* -mg_times */
    int mg_scale;       /* -mg_scale */
    int smoother;       /* -smoother */
    int coarse_solver;  /* -coarseSolver */
    int Ainv;           /* -Ainv */
    int max_newton_iterations; /* 3 */
    int max_lbfgs_iterations;  /* 10000 */
    int max_cg_iterations;     /* 10000 */
    double cneps;       /* -cneps */
    double topomega;    /* -mg_jomega */
} hot_solver_options;
/* fills the defaults of the HOT configuration (tog.sh:38): -lsolver 3 -Ainv 1 --project --linesearch --bcproject
 * -mg_level 3 -mg_times 1 -coarseSolver 2 -smoother 5 --usecn -cneps 1e-7 */
void hot_default_options(hot_solver_options* o);

#define HOT_LOG_CAP 256
/* telemetry of one backward-Euler solve: what the reference prints through ZIRAN_INFO (ImplicitSolver.h:186-207,
 * InexactConjugateGradient.h:73-80, LBFGS.h:337-349) */
typedef struct hot_solve_log {
    int iterations;              /* nonlinear iterations taken (Newton steps / L-BFGS iterations) */
    int converged;               /* shouldExitByCN returned true */
    int n_log;                   /* entries filled below (min(iterations + 1, HOT_LOG_CAP)) */
    int matrix_builds;           /* buildMatrix + hierarchy rebuilds */
    int total_linear_iterations; /* PCG iterations (lsolver 2) or V-cycles (lsolver 3) */
    int total_linesearch_probes; /* updateState calls made by lineSearch */
    double tolerance;            /* newton.tolerance = lbfgs.tolerance = cg.tolerance (MultigridSimulation.h:199-211) */
    double residual_norm[HOT_LOG_CAP]; /* ||r||_2 at the top of each nonlinear iteration */
    double scaled_norm[HOT_LOG_CAP];   /* sqrt(sum |r_i|^2 / tol_i^2 / n) with --usecn, else = residual_norm */
    double energy[HOT_LOG_CAP];        /* Ek (only with --linesearch) */
    int linear_iterations[HOT_LOG_CAP];/* PCG iterations of that Newton step */
} hot_solve_log;

/* InexactConjugateGradient::solve (Lib/Ziran/Math/Linear/InexactConjugateGradient.h:49-103) on A = assembled matrix
 * (matfree 0) or the matrix-free operator (matfree 1); preconditioner 0: none, 1: block/entry Jacobi of that operator,
 * 2: multigrid V-cycle.  x is in/out (must satisfy the BCs), returns the iteration count through *iters. */
int hot_pcg(hot_sim* h, const double* b, double* x, double tolerance, int max_iterations, int matfree, int preconditioner, int* iters);
/* MultigridSimulation::backwardEulerStep (Projects/multigrid/MultigridSimulation.h:188-233) minus
 * buildInitialDvAndVnForNewton, which the caller does through hot_set_bc: backupStrain, tolerances, Newton / L-BFGS solve
 * on the device-resident dv, restoreStrain.  The result stays in the device dv (hot_get_dv) for hot_g2p. */
int hot_backward_euler_step(hot_sim* h, const hot_solver_options* opt, hot_solve_log* log);
/* ImplicitSolverObjective::dv0 (ImplicitSolver.h:58): the last iterate the line search accepted.  With --linesearch the
 * reference exits with dv = dv0 + one more copy of the last step (x aliases simulation.dv, LBFGS.h:412-413 /
 * ExtendedNewtonsMethod.h:62 run after lineSearch already moved the nodes) and hot_backward_euler_step reproduces that;
 * without --linesearch dv0 == dv. */
int hot_get_dv0(hot_sim* h, double* dv0);

#ifdef __cplusplus
}
#endif
#endif

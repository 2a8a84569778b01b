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

/* ---- particles --------------------------------------------------------------------------------- */
/* Upload of particles.X/V/mass ("P","V","m": Lib/Ziran/Math/Geometry/Particles.h:8-45), the APIC matrix C,
 * the F-based helper's F / element measure (Lib/MPM/Force/FBasedMpmForceHelper.h:20-47) and the
 * CorotatedIsotropic parameters mu, lambda (Lib/Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h:64-73).
 * AoS, original order. */
int hot_set_particles(hot_sim* h, long n, const double* X, const double* V, const double* mass, const double* C,
    const double* F, const double* vol, const double* mu, const double* lambda);
/* read back particle state in original order; any pointer may be NULL */
int hot_get_particles(hot_sim* h, double* X, double* V, double* C, double* F, double* gradV);
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
/* FBasedMpmForceHelper::backupStrain / restoreStrain (Lib/MPM/Force/FBasedMpmForceHelper.cpp:25-44) */
int hot_backup_strain(hot_sim* h);
int hot_restore_strain(hot_sim* h);
/* ImplicitSolverObjective::updateState (ImplicitSolver.h:237-252): moveNodes(dv) (dv NULL = keep the device dv),
 * MpmForceBase::updatePositionBasedState (MpmForceBase.cpp:308-328); energy (nullable) = totalEnergy (:254-275) */
int hot_update_state(hot_sim* h, const double* dv, double* energy);
/* scratch_stress = vol P Fn^T and the trial F, original particle order (FBasedMpmForceHelper.cpp:72-97) */
int hot_get_stress(hot_sim* h, double* vPFnT, double* F);
/* ImplicitSolverObjective::computeResidual (ImplicitSolver.h:128-155) */
int hot_compute_residual(hot_sim* h, double* residual);
/* objective.project (MultigridSimulation.h:104-125), in place */
int hot_project(hot_sim* h, double* v);
/* ImplicitSolverObjective::multiply with --matfree (ImplicitSolver.h:741-763): b = M x + dt^2 K x */
int hot_hessian_apply_mf(hot_sim* h, const double* x, double* b);
/* ImplicitSolverObjective::evaluatePerNodeCNTolerance (ImplicitSolver.h:667-696); tol may be NULL (kept on device) */
int hot_eval_cn_tolerance(hot_sim* h, double eps, double dt, double* tol);

#ifdef __cplusplus
}
#endif
#endif

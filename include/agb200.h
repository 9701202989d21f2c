/* agb200.h — C ABI of the B200-native force path for AstroGenesis2.0.
 *
 * This library replaces, as a drop-in, the per-step force path that the reference keeps behind its
 * C++ class `Tree` (simulation/src/Physics/Tree/Tree.h:13-27) and that its driver calls in exactly
 * this order once at start-up and once per step (Physics/Simulation.cpp:121-139 and :276-285):
 *
 *     Tree* t = new Tree(sim);  t->buildTree();       // driver then reads t->root->radius
 *     t->calcVisualDensity();   t->calcGasDensity();  t->calculateForces();   delete t;
 *
 * The reference has no plugin/FFI layer of its own; the entry points below are what a binding of
 * that call surface needs (INTEGRATION.md shows the `Tree`-shaped C++ adaptor and the ctypes stub).
 * Plain pointers and sizes only; no C++/torch types; nothing throws across this boundary.  Every
 * function returns an agb_status; unlike the reference (which prints to std::cerr and continues,
 * e.g. Tree.cpp:70-74) problems are reported as codes.  One host thread per context.
 *
 * There is no CPU fallback: every entry point needs a CUDA device of compute capability 10.x and
 * fails with AGB_ERR_NO_DEVICE otherwise.
 */
#ifndef AGB200_H
#define AGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agb_ctx agb_ctx;

typedef enum {
    AGB_OK = 0,
    AGB_ERR_NO_DEVICE = 1,      /* no usable sm_100 device / CUDA runtime failure at create      */
    AGB_ERR_CUDA = 2,           /* a CUDA call or kernel failed; agb_last_error() has the text   */
    AGB_ERR_INVALID = 3,        /* bad argument or call order (e.g. forces before build_tree)    */
    AGB_ERR_DEPTH = 4,          /* two in-tree particles share all 63 octree levels (coincident  */
                                /* points: the reference recurses without bound, Node.cpp:618-666) */
    AGB_ERR_UNSUPPORTED = 5,    /* parameter range the parity path does not cover (see forces)   */
    AGB_ERR_NOMEM = 6
} agb_status;

typedef enum { AGB_MEM_HOST = 0, AGB_MEM_DEVICE = 1 } agb_memspace;

/* Structure-of-arrays view of the reference's particle record (Physics/Particle.h:18-57).
 * Required: x y z mass type.  Optional (NULL allowed): vx vy vz U (0), next_time (0), mu (0.58),
 * and the carried state rho P T h dUdt ax ay az (0) which the reference keeps inside Particle
 * between steps: orphan gas keeps rho/P/T (Node.cpp:749-793), dUdt is accumulated, never reset, by
 * the path (Node.cpp:167; TimeIntegration.cpp:37-40 resets it), and inactive particles keep acc
 * (Tree.cpp:75-80).  type: 1 = star, 2 = gas, 3 = dark matter. */
typedef struct {
    int64_t n;
    const double *x, *y, *z;
    const double *vx, *vy, *vz;
    const double *mass;
    const double *U;
    const double *next_time;        /* Particle::nextIntegrationTime */
    const double *mu;
    const uint8_t *type;
    const double *rho, *P, *T, *h, *dUdt;
    const double *ax, *ay, *az;
} agb_particles;

/* Results in the caller's particle order (the fields the path writes into Particle). NULL = skip. */
typedef struct {
    double *ax, *ay, *az;           /* Particle::acc           */
    double *dUdt;
    double *h, *rho, *P, *T;
    double *visualDensity;
} agb_results;

/* Field offsets (bytes) inside the caller's array-of-structs particle record, so the reference's
 * own `std::vector<Particle*>` (Simulation.h:70; sizeof(Particle) = 264) can be handed over as is.
 * Vectors are 3 consecutive doubles. A negative offset marks an absent optional field. */
typedef struct {
    int64_t position, velocity, acc, mass, type, U, next_time, mu, rho, P, T, h, dUdt, visualDensity;
} agb_aos_layout;

/* Whole-step counters of the last agb_forces() call. */
typedef struct {
    int64_t n_particles, n_in_tree, n_outliers, n_nodes, n_active;
    int64_t max_depth;
    int64_t edge_dropped;           /* particles that fail a cell-bounds test below the root (FP edge, Node.cpp:606-612); kept in tree, reported */
    int64_t interactions;           /* accepted (target, node) + (target, leaf) gravity pairs       */
    int64_t node_interactions, leaf_interactions, sph_interactions;
    int64_t node_visits;            /* calls of Node::calculateGravityForce the reference would make */
    int64_t mac_exact_fallbacks;    /* MAC / SPH-gate decisions taken on the exact FP64 slow path  */
    int64_t groups, gas_groups, gas_orphans;
    int64_t gas_ties_exact;         /* density-group decisions re-taken with the reference's own left-fold gasMass sums */
    int64_t gas_ties_unresolved;    /* ... that involved more than 8192 gas particles and kept the tree-order sums       */
    /* walk statistics: pop rounds, nodes popped, nodes straddling the opening radius of their warp, nodes opened by the
     * whole mask, 32-source tiles drained, rounds that touched the global-memory part of the stack */
    int64_t walk_rounds, walk_popped, walk_straddling, walk_opened, walk_tiles, walk_stack_spills;
    /* counter mode only: interaction-list entries whose acceptors span the warp / sit in one half / in one quarter of its
     * lanes, the acceptor bits of each class, and the entries imported from the far-field prepass (per group) */
    int64_t walk_ent_wide, walk_ent_half, walk_ent_quarter, walk_bits_wide, walk_bits_half, walk_bits_quarter, walk_ent_far;
    /* ... and the entries evaluated by each of the three pair loops of the mixed-precision walk (far + every target, far, near) */
    int64_t walk_ent_class0, walk_ent_class1, walk_ent_class2;
    int64_t sph_records;            /* mixed mode: 32-entry candidate records the walk wrote for the SPH pair kernel */
} agb_counters;

/* -------- lifetime: `new Tree(sim)` / `delete tree`, but persistent across steps (pooled memory) */
/* compat_cores = the reference's omp_get_max_threads(), which decides where bulk insertion hands
 * over to one-by-one insertion (Node.cpp:420) and hence which density groups see every particle
 * twice; <= 0 selects 1. */
int agb_create(agb_ctx** out, int device, int compat_cores);
int agb_destroy(agb_ctx* ctx);

/* -------- particle hand-over (replaces the path's direct reads of Simulation::particles)
 * Host arrays are uploaded asynchronously and must stay untouched until agb_build_tree() has returned (the reference
 * likewise reads its particles during buildTree); device arrays are read in place until the next hand-over. */
int agb_set_particles(agb_ctx* ctx, const agb_particles* p, int memspace);
int agb_set_particles_aos(agb_ctx* ctx, void* const* particles, int64_t n, const agb_aos_layout* layout);
/* Hand-over of arrays that are still being produced on the caller's own streams (the all-gather of a multi-GPU driver, a copy,
 * its integrator kernels): each event (cudaEvent_t, NULL = ready now) must have been recorded behind the producer of its group —
 * x y z mass type | next_time | vx vy vz U mu and the carried state — and the path reads a group only after its event.  With
 * agb_force_path the build, the densities and the gravity walk overlap the production of the last group; when ready_next_time and
 * ready_all are the same event there is no late group and only extent, keys and sort run ahead of it (no extra kernels). */
int agb_set_particles_staged(agb_ctx* ctx, const agb_particles* p, int memspace, void* ready_positions, void* ready_next_time, void* ready_all);

/* -------- the four calls of the reference's Tree (same order, same meaning) */
int agb_build_tree(agb_ctx* ctx, double* root_radius);                  /* Tree::buildTree, Tree.cpp:24-55; *root_radius = root->radius */
int agb_visual_density(agb_ctx* ctx, double visual_density_radius);     /* Tree::calcVisualDensity, Tree.cpp:152-176 */
int agb_gas_density(agb_ctx* ctx, double mass_in_h);                    /* Tree::calcGasDensity, Tree.cpp:119-150 */
int agb_forces(agb_ctx* ctx, double global_time, double e0, double theta);   /* Tree::calculateForces, Tree.cpp:57-83 */
/* Multi-GPU variant: walk only the `part`-th of `nparts` contiguous slices of the tree-ordered
 * targets (every GPU holds the same gathered particles and builds the same tree). */
int agb_forces_slice(agb_ctx* ctx, double global_time, double e0, double theta, int part, int nparts);
/* The four calls above in one, with a single host synchronisation at the end (a driver that does not need root->radius
 * between the calls: the reference reads it only at init, Simulation.cpp:123-126).  Same results as the separate calls;
 * *root_radius (optional) is the radius of this step's tree. */
int agb_force_path(agb_ctx* ctx, double visual_density_radius, double mass_in_h, double global_time, double e0, double theta,
                   int part, int nparts, double* root_radius);
/* The targets of slice (part, nparts) after agb_forces[_slice]: their number, and their results in compact form —
 * index[k] = position of the k-th target (tree order) in the caller's particle arrays, ax/ay/az/dUdt[k] its results.
 * What each GPU of a target-sharded run sends back instead of N-sized arrays: the slices of all parts are disjoint and
 * together hold every active particle exactly once (Tree.cpp:65-80 writes acc / dUdt of active particles only).
 * Output arrays need room for *count entries (query with index = NULL first, or size them for N); NULL = not wanted. */
int agb_get_slice_count(agb_ctx* ctx, int part, int nparts, int64_t* count);
int agb_get_slice_results(agb_ctx* ctx, int part, int nparts, uint32_t* index, double* ax, double* ay, double* az, double* dUdt, int memspace);
/* ... the same for every field the path writes into Particle: r names the wanted columns (NULL = skip), each with room for the
 * slice's *count entries; h / rho / P / T / visualDensity are those of the slice's targets (every GPU computes all densities). */
int agb_get_slice_results_all(agb_ctx* ctx, int part, int nparts, uint32_t* index, const agb_results* r, int memspace);
/* Registers HOST destinations (ideally pinned) for the compact results of slice (part, nparts): agb_force_path(.., part, nparts, ..)
 * then delivers them itself and returns when they have arrived — index and the density columns while the walk still runs (when
 * every particle is a force target, as in fixed-step runs), acc / dU/dt after it, piece by piece for a large slice (the walk of
 * piece k+1 hides the transfer of piece k; the counters then hold the sums over the pieces).  r = NULL unbinds. */
int agb_bind_slice_results(agb_ctx* ctx, int part, int nparts, uint32_t* index, const agb_results* r);

/* -------- device-resident driver loop (optional).  With particles handed over from HOST memory the
 * context owns device copies; these calls advance them in place exactly like the reference's loop, so nothing but the
 * scalar time crosses PCIe per step:
 *   agb_integrator_init(..)                       gas T from U (Simulation.cpp:108-112), nextIntegrationTime = 0
 *   build_tree / visual_density / gas_density / forces(global_time = 0)       initial forces (Simulation.cpp:120-139)
 *   agb_integrator_assign_all()                   power-of-two time steps from |acc| (Simulation.cpp:189-208)
 *   per step: agb_step_begin(&t)  ->  build_tree, visual_density, gas_density, forces(t)  ->  agb_step_end()
 * step_begin = re-binning of due particles, t = min nextIntegrationTime, first Kick + Drift of the active particles
 * (Simulation.cpp:213-272, TimeIntegration.cpp:10-26); step_end = Ueuler, Hubble rescale, second Kick, next += dt
 * (Simulation.cpp:296-341).  agb_get_state copies positions, velocities, U, next time and time step back (NULL = skip). */
int agb_integrator_init(agb_ctx* ctx, double eta, double min_time_step, double max_time_step, double H0, double e0);
int agb_integrator_assign_all(agb_ctx* ctx);
int agb_step_begin(agb_ctx* ctx, double* global_time);
int agb_step_end(agb_ctx* ctx);
int agb_get_state(agb_ctx* ctx, double* x, double* y, double* z, double* vx, double* vy, double* vz, double* U, double* next_time, double* time_step);
/* what the sub-grid hooks changed: particle types (gas that became stars) and Particle::sfr (the last conversion probability) */
int agb_get_subgrid_state(agb_ctx* ctx, uint8_t* type, double* sfr);

/* -------- results (replaces the path's direct writes into Particle) */
int agb_get_results(agb_ctx* ctx, const agb_results* r, int memspace);
/* Optional: register the destinations of agb_get_results in advance.  Each bound array is then sent on a second stream
 * as soon as its phase has finished — visualDensity after agb_visual_density, h / rho / P / T after agb_gas_density — so
 * those transfers overlap the tree walk; agb_get_results with the same pointers sends the rest (acc, dUdt) and waits for
 * all of it.  The arrays must stay valid (and untouched) until agb_get_results returns; r = NULL unbinds. */
int agb_bind_results(agb_ctx* ctx, const agb_results* r, int memspace);
int agb_get_results_aos(agb_ctx* ctx, void* const* particles, int64_t n, const agb_aos_layout* layout);
int agb_get_counters(agb_ctx* ctx, agb_counters* c);

/* -------- options */
typedef enum {
    AGB_OPT_TARGET_COUNTERS = 1,    /* 1: also record per-target visit / accept / SPH counts (parity tests) */
    AGB_OPT_COOLING = 3,            /* device-resident loop: 1 = free-free cooling of active gas in the second kick (Cooling.cpp:6-25;
                                       the reference ships the call commented out, Simulation.cpp:312-315).  Default 0. */
    AGB_OPT_STAR_FORMATION = 4,     /* device-resident loop: != 0 = stochastic gas -> star conversion (SFR.cpp:12-34; commented out in the
                                       reference, Simulation.cpp:316-320); the value seeds a counter-based generator keyed by
                                       (particle, time) in place of the reference's rand().  Default 0. */
    AGB_OPT_EXTENDED = 5,           /* 1 = extended-accuracy mode (NOT the reference's algorithm, parity unpinned): gravity with
                                       monopole + quadrupole moments, Newtonian with cubic-spline softening (length 2.8 e0), one interaction
                                       list per 32 targets, opening test (cell WIDTH) / distance < theta; SPH with a smoothing length per particle from (4 pi/3)(2h)^3 rho = massInH and
                                       neighbour loops for density, pressure, viscosity and dU/dt.  Same calls, FP64 throughout.  Default 0.
                                       (2 = the same without the quadrupole term: a validation aid.) */
    AGB_OPT_SLICE_PIECE = 6,        /* tuning: with agb_bind_slice_results, a slice of at least 2 x this many targets is walked in up to 4 pieces
                                       (exact sub-ranges: same bits) so that the acc / dU/dt of a piece leave while the next one walks.
                                       Default 2 000 000. */
    AGB_OPT_SLICE_DENSITIES = 7,    /* 1: a sliced agb_force_path (nparts > 1) produces visualDensity / h / rho / P / T only for the targets of ITS slice —
                                       what the slice getters hand back; the SPH terms of a target use its own h, rho, P only (Node.cpp:94,101,108), so
                                       acc and dU/dt are unchanged.  For callers that collect results per slice (one process per GPU); the full-length
                                       getters then hold zeros / the handed-over state outside the slice.  Needs every particle active (otherwise the step
                                       is redone with all densities).  Default 0. */
    AGB_OPT_PRECISION = 2           /* arithmetic of the pair forces: 0 = FP64 throughout (agrees with the reference to ~1e-14),
                                       1 = mixed (default): float-float displacements, FP32 law, FP64 accumulation; ~1e-7.
                                       The accepted (target, source) sets, SPH pair sets and densities are identical in both;
                                       the SPH pair algebra (kernel gradient, viscosity) runs in FP32 in mixed mode (~1e-7). */
} agb_option;
int agb_set_option(agb_ctx* ctx, int option, int64_t value);

/* -------- introspection used by the parity tests (tree topology, node table, per-target counters) */
/* caller order; leafdepth = -1 for particles outside the root cube; key = octant path root->leaf,
 * 3 bits per level, level l < 21 at key_hi >> (60-3l), else key_lo >> (60-3(l-21)). */
int agb_get_tree_particles(agb_ctx* ctx, int32_t* leafdepth, uint64_t* key_hi, uint64_t* key_lo);
int agb_get_node_count(agb_ctx* ctx, int64_t* n_internal);
int agb_get_nodes(agb_ctx* ctx, int32_t* depth, int64_t* count, int32_t* duplicated, uint64_t* key_hi, uint64_t* key_lo,
                  double* mass, double* comx, double* comy, double* comz, double* gas_mass, double* mvx, double* mvy, double* mvz);
int agb_get_target_counters(agb_ctx* ctx, int32_t* visits, int32_t* acc_nodes, int32_t* acc_leaves, int32_t* sph);

/* Device time (ms, CUDA events on the context's stream) of the last call of each phase:
 * [0] build_tree [1] visual_density [2] gas_density [3] forces (walk kernel only) [4] forces (whole call). */
int agb_get_phase_ms(agb_ctx* ctx, double ms[5]);
/* Finer device times of the last step (ms): [0] k_far [1] k_walk [2] k_sph; build_tree split into [3] root extent + keys
 * [4] radix sort [5] gather into tree order [6] lcp + scan + links [7] upward pass + finalize. */
int agb_get_kernel_ms(agb_ctx* ctx, double ms[8]);
/* The CUDA stream (cudaStream_t) all work of this context is issued on. */
int agb_get_stream(agb_ctx* ctx, void** stream);
/* Number of kernels this context launched since creation. */
int agb_get_launch_count(agb_ctx* ctx, int64_t* launches);

/* -------- several GPUs of one box behind one handle.  The reference's data-parallel axis is the loop
 * over force targets against one shared tree (Tree::calculateForces, Tree.cpp:65).  Every device receives the whole particle
 * set, builds the same tree and computes the same densities, walks its own slice of the tree-ordered targets, and the slices'
 * (index, acc, dU/dt) are exchanged over peer-to-peer copies, so every device ends with the complete result arrays — bit-identical
 * to a single-GPU run, for any number of devices.  One process, one host thread per device inside the calls, no NCCL.
 * devices = NULL selects 0 .. ndev-1.  Same call order and meaning as the single-context functions above; particle arrays are
 * HOST arrays; agb_multi_get_results fills caller-order host arrays (each device sends 1/ndev of the rows over its own link). */
typedef struct agb_multi agb_multi;
int agb_multi_create(agb_multi** out, const int* devices, int ndev, int compat_cores);
int agb_multi_destroy(agb_multi* m);
int agb_multi_device_count(agb_multi* m);
int agb_multi_context(agb_multi* m, int i, agb_ctx** ctx);              /* the i-th device's context (counters, timings) */
int agb_multi_set_option(agb_multi* m, int option, int64_t value);
int agb_multi_set_particles(agb_multi* m, const agb_particles* p);
int agb_multi_set_particles_aos(agb_multi* m, void* const* particles, int64_t n, const agb_aos_layout* layout);
int agb_multi_build_tree(agb_multi* m, double* root_radius);
int agb_multi_visual_density(agb_multi* m, double visual_density_radius);
int agb_multi_gas_density(agb_multi* m, double mass_in_h);
int agb_multi_forces(agb_multi* m, double global_time, double e0, double theta);
int agb_multi_force_path(agb_multi* m, double visual_density_radius, double mass_in_h, double global_time, double e0, double theta, double* root_radius);
int agb_multi_get_results(agb_multi* m, const agb_results* r);
int agb_multi_get_results_aos(agb_multi* m, void* const* particles, int64_t n, const agb_aos_layout* layout);
/* device-resident loop: the integrator kernels run replicated on every device, only the slices' results cross NVLink */
int agb_multi_integrator_init(agb_multi* m, double eta, double min_time_step, double max_time_step, double H0, double e0);
int agb_multi_integrator_assign_all(agb_multi* m);
int agb_multi_step_begin(agb_multi* m, double* global_time);
int agb_multi_step_end(agb_multi* m);
int agb_multi_get_state(agb_multi* m, double* x, double* y, double* z, double* vx, double* vy, double* vz, double* U, double* next_time, double* time_step);
int agb_multi_get_subgrid_state(agb_multi* m, uint8_t* type, double* sfr);
const char* agb_multi_last_error(agb_multi* m);

/* Roofline denominators measured on this device with tiny kernels (not part of the force path):
 * kind 0 = FP64 FMA throughput [TFLOP/s], 1 = FP32 FMA throughput [TFLOP/s], 2 = HBM copy bandwidth [GB/s]. */
int agb_microbench(agb_ctx* ctx, int kind, double* result);

const char* agb_strerror(int status);
const char* agb_last_error(agb_ctx* ctx);
const char* agb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AGB200_H */

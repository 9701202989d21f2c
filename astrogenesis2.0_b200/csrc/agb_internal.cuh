// agb_internal.cuh — device-side data model shared by the kernels of the B200 force path.
//
// HBM layout (all structure-of-arrays, 180 GB budget; ~330 B per particle incl. tree):
//   caller order   : x y z vx vy vz mass U next mu type      (inputs, bound or copied)
//                    ax ay az dUdt h rho P T vis              (carried state / results)
//   tree order     : key_hi key_lo perm                       (Morton-like 126-bit octant paths)
//                    src_pm[N+M]  double4 (x, y, z, mass)     unified "source" table: index < N is
//                    src_gv[N+M]  double4 (vx, vy, vz, gasM)  the i-th sorted particle, N+k is the
//                    src_flag[N+M]                            k-th internal octree node (COM, mVel)
//                    s_h s_rho s_P s_U s_mu s_next s_type leafdepth leafparent group
//   nodes          : child[M][8] (int32 source index, -1 = empty) depth first last parent
//                    mom_pm mom_gv (un-normalised moments of the upward pass) mark dup arrived
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/agb200.h"

#define AGB_MAX_LEVELS 63          /* three key words of 21 levels; the third exists only in 'deep' builds */
#define AGB_SHALLOW_LEVELS 42      /* key_hi + key_lo: what the default build can tell apart */
#define AGB_OUTLIER_BIT 0x8000000000000000ull
#define AGB_GAS_BIT 0x80000000u          /* sort payload: caller index (< 2^30) | "type == 2" */
#define AGB_IDX_MASK 0x7fffffffu

struct AgbDev {
    int64_t n = 0, cap = 0;
    // AGB_OPT_SLICE_DENSITIES: the per-particle outputs of the density passes are only produced for tree positions [dens_a0, dens_a1)
    // (a sliced step hands back nothing else; the SPH terms use the TARGET's h / rho / P only, Node.cpp:94,101,108)
    int64_t dens_a0 = 0, dens_a1 = INT64_MAX;
    int64_t ncap = 0;                  // capacity of the node arrays: one node per (first particle, depth) pair, so a tight pair
                                       // alone costs up to 41 nodes; grown on demand when a build reports more (agb_api.cu)
    int cores = 1;
    // caller-order inputs (owned copies unless `bound`)
    const double *x = nullptr, *y = nullptr, *z = nullptr, *vx = nullptr, *vy = nullptr, *vz = nullptr;
    const double *mass = nullptr, *U = nullptr, *next = nullptr, *mu = nullptr;
    const uint8_t* type = nullptr;
    // caller-order state / results (always owned)
    double *ax = nullptr, *ay = nullptr, *az = nullptr, *dUdt = nullptr, *h = nullptr, *rho = nullptr, *P = nullptr, *T = nullptr, *vis = nullptr;
    // keys + permutation, ping-pong for the radix sort
    uint64_t *khi[2] = {nullptr, nullptr}, *klo[2] = {nullptr, nullptr};
    uint32_t* perm[2] = {nullptr, nullptr};
    int cur = 0;                       // which ping-pong half holds the sorted result
    // tree-order particle data
    double4 *src_pm = nullptr, *src_gv = nullptr;
    uint8_t* src_flag = nullptr;
    double *s_h = nullptr, *s_rho = nullptr, *s_P = nullptr, *s_U = nullptr, *s_mu = nullptr, *s_next = nullptr, *s_T = nullptr;
    uint8_t* s_type = nullptr;
    int8_t* lcp = nullptr;             // common levels of sorted keys i, i+1
    int32_t *nodebase = nullptr, *nodecnt = nullptr;
    int32_t *leafparent = nullptr, *group = nullptr;
    int8_t* leafdepth = nullptr;
    // nodes (capacity cap)
    int32_t *child = nullptr, *nfirst = nullptr, *nlast = nullptr, *nparent = nullptr, *arrived = nullptr;
    int8_t* ndepth = nullptr;
    uint8_t *nmark = nullptr, *ndup = nullptr, *leafmark = nullptr;
    double4 *mom_pm = nullptr, *mom_gv = nullptr;
    int32_t* grouplist = nullptr;
    int32_t* lvl_list = nullptr;       // node ids by depth (level lists of the upward passes)
    int32_t* gasrank = nullptr;        // exclusive count of gas particles before tree position i
    // scratch
    double4* rec = nullptr;            // caller order: (x, y, z, mass) packed by the extent pass
    // 'deep' builds (agb_api.cu switches them on when a default build reports need_deep): all three key words of every particle,
    // three 8-pass sorts.  dk = caller-order key words (hi, lo, ex); kex = levels 42..62 in tree order
    bool deep = false;
    uint64_t* dk[3] = {nullptr, nullptr, nullptr};
    uint64_t* kex = nullptr;
    double* quad = nullptr;            // extended mode: traceless quadrupole of every node about its COM, (xx xy xz yy yz zz)
    unsigned int* ext_bar = nullptr;   // ... arrival counter of its level-synchronous pass
    double4* grec = nullptr;           // caller order, gas only: (vx, vy, vz, U), (mu, rho, P, T) packed before the gather
    uint32_t* blockhist = nullptr;     // radix sort: [256][nblocks]
    int32_t* scanblk = nullptr;
    // per-target counters (optional)
    int32_t *c_visits = nullptr, *c_accn = nullptr, *c_accl = nullptr, *c_sph = nullptr;
    // walk spill stack
    int2* spill = nullptr; int64_t spill_per_warp = 0; int spill_warps = 0;
    // far-field prepass output, per super-group of 256 targets
    int32_t *far_list = nullptr, *far_front = nullptr, *far_cnt = nullptr;
    int32_t* act_list = nullptr;       // tree positions of the active targets of the current forces call
    // mixed-precision SPH: records (32 x (source, accepting gas targets) + link) written by the walk for k_sph, last record per group
    int2* rec_ent = nullptr; int32_t *rec_next = nullptr, *rec_head = nullptr; int64_t rec_cap = 0;
};

// Device-resident scalars of one step (read back in a single copy when the host needs them).
struct AgbScalars {
    double partial_sum[1024], partial_sq[1024];
    double mean, stdev, limit, R;
    unsigned long long Rbits;
    int32_t n_in_tree, n_outliers, n_nodes, dup_keys, edge_dropped, max_depth;
    int32_t need_deep;                 // the two-word build met particles that share 42 levels, or > 4096 that share 21: rebuild with full keys
    int32_t n_groups, n_gas_groups, n_gas_orphans, n_active;
    unsigned int walk_next_group;
    unsigned long long cand_cursor;    // bump allocator of the SPH tile records
    unsigned long long c_interactions, c_node, c_leaf, c_sph, c_visits, c_exact, c_spill;
    int32_t bintotal[256];
    int32_t vis_level;
    int32_t walk_overflow, any_gas;
    int32_t next_uniform;              // every particle has the same nextIntegrationTime (fixed-step runs): the gather skips that column
    unsigned int grid_bar, grid_bar2;  // arrival counters of the level-synchronous upward passes (all / mVel only)
    int32_t lvl_cnt[64], lvl_cur[64];  // internal nodes per depth, fill cursors of the level lists
    int32_t node_overflow;             // the build needed more than ncap nodes: nothing past the capacity was written, the host grows and rebuilds
    int32_t n_gas_total, tie_exact, tie_unresolved, n_fold, n_long_runs, n_scan_tmp;
    unsigned long long gas_mmin, gas_mmax;   // bit patterns of the smallest / largest gas particle mass (equal: the exact tie sums need no sort)
    unsigned long long st_rounds, st_popped, st_mixed, st_open, st_drain;   // walk statistics (tuning)
    unsigned long long st_cls[10];     // counter mode: list entries / acceptor bits by lane span (any, one half, one quarter), far-list entries, entries per evaluation class
};

// Device-resident integrator state (agb_integrate.cu): mutable views of the owned particle copies + per-particle step
#define AGB_INT_BINS 128
struct AgbInt {
    double *x, *y, *z, *vx, *vy, *vz, *U, *next; const double* mu;
    double* timestep;
    double eta, e0, min_ts, max_ts;
    double scale_min; int k0; double scale_tab[AGB_INT_BINS];   // exp(H0 dt) per power-of-two bin 2^(k0 + j), from the host's libm
    // sub-grid hooks of the second kick (Simulation.cpp:311-320 calls them from there; commented out in the reference)
    int cooling, star_formation; unsigned long long seed;
    uint8_t* type; double* sfr;                                  // particle types (gas -> star) and Particle::sfr
    double sf_min, sf_tab[AGB_INT_BINS];                         // 1 - exp(-epsilon dt / t_star) per time-step bin, from the host's libm
};

// uniform deviate in [0, 1) of (seed, particle, time): the reference draws rand() / RAND_MAX inside an OpenMP loop (SFR.cpp:25),
// which is neither reproducible nor thread safe; a counter-based generator gives every (particle, step) its own number
__host__ __device__ __forceinline__ double agb_u01(unsigned long long seed, unsigned long long particle, double time)
{
    unsigned long long tb;
    memcpy(&tb, &time, 8);
    unsigned long long z = seed + particle * 0x9E3779B97F4A7C15ull + tb * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
int agb_launch_int_init(AgbDev& d, const AgbInt& I, cudaStream_t st);
int agb_launch_int_assign(AgbDev& d, const AgbInt& I, double gt, bool all, cudaStream_t st);
int agb_launch_int_min(AgbDev& d, const AgbInt& I, unsigned long long* out, cudaStream_t st);
int agb_launch_int_first(AgbDev& d, const AgbInt& I, double gt, cudaStream_t st);
int agb_launch_int_second(AgbDev& d, const AgbInt& I, double gt, cudaStream_t st);

// ---- host-callable launchers (each returns the number of kernels it launched) ----
int agb_launch_extent(const AgbDev& d, AgbScalars* s, cudaStream_t st, bool mass_late = false);   // mass_late: the masses are still on their way (agb_launch_fill_mass before the gather)
int agb_launch_fill_mass(const AgbDev& d, cudaStream_t st);
int agb_launch_keygen(AgbDev& d, AgbScalars* s, cudaStream_t st);
int agb_launch_sort(AgbDev& d, AgbScalars* s, cudaStream_t st);
int agb_launch_links(AgbDev& d, AgbScalars* s, cudaStream_t st, cudaEvent_t* ev = nullptr, bool late_gas = false);   // ev[0] after the gather, ev[1] after the links, before the upward pass
int agb_launch_late_gas(AgbDev& d, AgbScalars* s, cudaStream_t st);            // completes a late_gas build once velocities / U / mu have arrived (after the gas densities)
int agb_launch_gather_next(AgbDev& d, AgbScalars* s, cudaStream_t st);          // tree-order next_time of a late_gas build (arrives behind the masses)
int agb_launch_visual(AgbDev& d, AgbScalars* s, double radius, cudaStream_t st);
int agb_launch_gas_density(AgbDev& d, AgbScalars* s, double massInH, cudaStream_t st, bool late_pt = false);   // late_pt: h and rho only (P, T follow in agb_launch_late_gas)
// compact (index, acc, dUdt) of the active targets [a0, a1) in tree order (agb_get_slice_results)
void agb_slice_bounds(int64_t n_active, int part, int nparts, int64_t* a0, int64_t* a1);
int agb_launch_slice_results(const AgbDev& d, int64_t a0, int64_t a1, bool ident, uint32_t* index, double* const dst[9], cudaStream_t st);   // dst: ax ay az dUdt h rho P T vis
int agb_launch_walk(AgbDev& d, AgbScalars* s, double globalTime, double e0, double theta, int part, int nparts,
                    bool counters, bool any_gas, bool mixed, int sm_count, cudaStream_t st, cudaEvent_t* ev, int phase = 0);   // ev[0..4]: before k_far, before k_walk, after k_walk, after k_sph, before k_sph
// the step's force targets (globalTime == nextIntegrationTime) in tree order, compacted; resets the walk's counters
int agb_launch_active_list(AgbDev& d, AgbScalars* s, double globalTime, int sm_count, cudaStream_t st, bool mixed = false, double e0 = 0.0);
// extended-accuracy mode (agb_extended.cu): per-particle smoothing lengths and densities; quadrupole walk + neighbour-loop SPH forces
int agb_launch_gas_list(AgbDev& d, AgbScalars* s, cudaStream_t st);   // tree positions of the gas particles, compact, in d.nodecnt; count in s->n_gas_total
int agb_launch_extended_density(AgbDev& d, AgbScalars* s, double massInH, cudaStream_t st);
int agb_launch_extended_forces(AgbDev& d, AgbScalars* s, double globalTime, double e0, double theta, int part, int nparts, bool any_gas, bool use_quad, int sm_count, cudaStream_t st, cudaEvent_t* ev);
int agb_launch_dump_tree(AgbDev& d, AgbScalars* s, int32_t* leafdepth, uint64_t* khi, uint64_t* klo, cudaStream_t st);
int agb_launch_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* blk, int32_t* total_out, cudaStream_t st, const int32_t* skip_if_n = nullptr);
int agb_launch_microbench(int kind, int sm_count, cudaStream_t st, double* result);
// agb_multi.cu reaches into a context it drives (agb_api.cu)
void agb_ctx_internals(agb_ctx* c, AgbDev** d, cudaStream_t* st, int* device);
int agb_ctx_copy_particles_from(agb_ctx* c, agb_ctx* src);   // the host hand-over `src` received, repeated on c by peer-to-peer copies
void agb_ctx_join_uploads(agb_ctx* c);            // the compute stream waits for every upload of the last hand-over
int agb_launch_unpermute_counters(AgbDev& d, int32_t* v, int32_t* an, int32_t* al, int32_t* sp, cudaStream_t st);

// ---- small device helpers ----
__device__ __forceinline__ int agb_octant_at(uint64_t hi, uint64_t lo, uint64_t ex, int level)
{
    return level < 21 ? (int)((hi >> (60 - 3 * level)) & 7) : level < 42 ? (int)((lo >> (60 - 3 * (level - 21))) & 7) : (int)((ex >> (60 - 3 * (level - 42))) & 7);
}

// number of leading octree levels two keys share (0..63; 42 is the most a build without the third word can report)
__device__ __forceinline__ int agb_common_levels(uint64_t ahi, uint64_t alo, uint64_t aex, uint64_t bhi, uint64_t blo, uint64_t bex, bool deep)
{
    uint64_t xh = ahi ^ bhi;
    if (xh) return (__clzll((long long)xh) - 1) / 3;
    uint64_t xl = alo ^ blo;
    if (xl) return 21 + (__clzll((long long)xl) - 1) / 3;
    if (!deep) return AGB_SHALLOW_LEVELS;
    uint64_t xe = aex ^ bex;
    if (xe) return 42 + (__clzll((long long)xe) - 1) / 3;
    return AGB_MAX_LEVELS;
}

// agb_api.cu — the C ABI (include/agb200.h) over the device kernels: context, memory pool,
// particle hand-over, the four Tree calls and result read-back.  Host-side only.
#include "agb_internal.cuh"
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

int agb_walk_blocks(int sm_count);
int agb_walk_warps_per_block();
void agb_far_capacity(int* lcap, int* fcap, int* targets);
size_t agb_sort_scratch_words(int64_t cap);

struct agb_ctx {
    int device = 0, sm_count = 148;
    cudaStream_t st = nullptr, st_copy = nullptr;   // compute stream; upload stream for everything but x, y, z
    cudaStream_t st_zero = nullptr; cudaEvent_t ev_zero = nullptr;   // zero fills of the columns a hand-over leaves out (see agb_set_particles)
    bool zero_pending = false;                        // ... the compute stream has not been ordered behind them yet (device hand-over: they overlap extent, keys and sort)
    cudaEvent_t ev_in = nullptr;                      // uploads on st_copy complete
    cudaEvent_t ev_next = nullptr;                    // ... their first group (next_time and the carried acc / dUdt / h / rho): all the build, the densities and the gravity walk read
    cudaEvent_t ev_sync = nullptr;                    // compute stream reached the point of a new hand-over (orders st_copy after it)
    cudaEvent_t ev_pos = nullptr;                     // type, x, y, z are on the device (st): the other uploads start after them (they would share the link)
    cudaEvent_t ev_mass = nullptr; bool mass_late = false;   // host hand-over: the masses follow on the copy stream (needed from the gather on)
    bool in_pending = false, next_pending = false;
    cudaEvent_t xev[3] = {};                          // agb_set_particles_staged: the caller's "group is complete" events (positions+mass+type, next_time, the rest)
    cudaEvent_t evw[5] = {};                          // walk timing: before k_far, before k_walk, after k_walk, after k_sph, before k_sph
    cudaEvent_t ev[10] = {};
    cudaEvent_t tl[7] = {}; bool timeline = false;     // AGB_TIMELINE=1 (diagnostic): one step's hand-over / compute / delivery times on stderr
    cudaEvent_t evk[10] = {};                         // kernel-level timing: walk [0..3] = before k_far, k_walk, k_sph, after; build [4..9] = start, keys, sort, gather, links, end
    double kernel_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // k_far k_walk k_sph | extent+keys sort gather lcp+scan+links upward+finalize
    bool gas_hint_valid = false, gas_hint = false;   // agb_force_path: "the particle set holds gas" as of the last completed step
    AgbDev d;
    AgbScalars* s = nullptr;            // device
    AgbScalars hs;                      // host mirror (tail only is valid)
    double* in_d[10] = {};             // pooled device copies of caller-order inputs (x y z vx vy vz mass U next mu)
    uint8_t* in_type = nullptr;
    bool bound = false, have_particles = false, built = false, dens_done = false, forces_done = false;
    bool mixed = true;                  // AGB_OPT_PRECISION
    bool walked_mixed = false;          // the last walk used the FP32 pair law (mixed_in_range)
    bool ext_quad = true;               // ... with the quadrupole term (AGB_OPT_EXTENDED = 2: monopole only, for validation)
    bool slice_dens = false;            // AGB_OPT_SLICE_DENSITIES
    int64_t piece_targets = 2000000;    // AGB_OPT_SLICE_PIECE: bound slice results are pipelined in up to 4 pieces of at least this many targets
    bool extended = false;              // AGB_OPT_EXTENDED: quadrupoles + spline softening + per-particle-h SPH (agb_extended.cu)
    bool opt_cooling = false; unsigned long long opt_sf_seed = 0;   // AGB_OPT_COOLING, AGB_OPT_STAR_FORMATION
    double* sfr = nullptr;              // Particle::sfr (device-resident loop)
    bool target_counters = false, counters_valid = false, vis_timed = false, gas_timed = false, build_timed = false;
    double phase_ms[5] = {0, 0, 0, 0, 0};
    int64_t launches = 0;
    std::string err;
    // pinned staging for host particles
    double* stage = nullptr; size_t stage_bytes = 0;
    // device-resident integrator
    AgbInt I = {}; bool int_ready = false; double* timestep = nullptr; unsigned long long* d_min = nullptr; double int_time = 0.0;
    agb_counters last = {};
    // agb_bind_results: destinations registered in advance, streamed out as soon as their phase is done
    agb_results bres = {}; bool bres_on = false; int bres_space = AGB_MEM_HOST; bool bres_sent[9] = {};
    cudaEvent_t ev_out = nullptr;
    // agb_bind_slice_results: host destinations of one target slice's compact results, delivered by agb_force_path itself
    bool bs_on = false, bs_early = false, bs_late = false, bs_piped = false; int bs_part = 0, bs_nparts = 1; uint32_t* bs_index = nullptr; agb_results bs_res = {};
    double* bs_buf = nullptr; uint32_t* bs_idx = nullptr; int64_t bs_cap = 0;
};

namespace {

// NVTX ranges named like the reference's own phase log (Log::startProcess, Simulation.cpp:120-139 / :256-295), so a timeline
// of a run reads like its logs/processLog.csv.  Header-only NVTX: a no-op unless a profiler is attached.
struct Phase {
    explicit Phase(const char* name) { nvtxRangePushA(name); }
    ~Phase() { nvtxRangePop(); }
};

constexpr int64_t SPILL_PER_WARP = 8192;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            if (e_ == cudaErrorMemoryAllocation) { (void)cudaGetLastError(); return AGB_ERR_NOMEM; } \
            return AGB_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

template <class T> cudaError_t dalloc(T*& p, size_t count) { return cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T)); }
template <class T> void dfree(T*& p) { if (p) cudaFree((void*)p); p = nullptr; }
template <class T> struct DevTmp { T* p = nullptr; ~DevTmp() { if (p) cudaFree((void*)p); } };   // scoped device temporary

// Node-indexed arrays (and the scratch that doubles as per-node storage in the density pass).  Their capacity is separate
// from the particle capacity: the tree has one node per (first particle, depth) pair, so M ~ 0.48 N for galaxies but up
// to 41 N for sets of tight pairs (Sun / Earth / Moon: N = 3, M = 9).
void free_nodes(agb_ctx* c)
{
    AgbDev& d = c->d;
    dfree(d.src_pm); dfree(d.src_gv); dfree(d.src_flag);
    dfree(d.child); dfree(d.nfirst); dfree(d.nlast); dfree(d.nparent); dfree(d.arrived); dfree(d.ndepth);
    dfree(d.nmark); dfree(d.ndup); dfree(d.mom_pm); dfree(d.mom_gv); dfree(d.grouplist); dfree(d.lvl_list);
    dfree(d.klo[0]); dfree(d.nodebase); dfree(d.quad); dfree(d.ext_bar);
    d.ncap = 0;
}

void free_pool(agb_ctx* c)
{
    AgbDev& d = c->d;
    free_nodes(c);
    for (auto& q : c->in_d) dfree(q);
    dfree(c->in_type); dfree(c->timestep);
    dfree(d.ax); dfree(d.ay); dfree(d.az); dfree(d.dUdt); dfree(d.h); dfree(d.rho); dfree(d.P); dfree(d.T); dfree(d.vis);
    for (int i = 0; i < 2; i++) { dfree(d.khi[i]); dfree(d.klo[i]); dfree(d.perm[i]); }
    dfree(d.s_h); dfree(d.s_rho); dfree(d.s_P); dfree(d.s_U); dfree(d.s_mu); dfree(d.s_next); dfree(d.s_T); dfree(d.s_type);
    dfree(d.lcp); dfree(d.nodecnt); dfree(d.leafparent); dfree(d.group); dfree(d.leafdepth);
    dfree(d.leafmark); dfree(d.gasrank);
    dfree(d.rec); dfree(d.grec); dfree(d.blockhist); dfree(d.scanblk);
    for (auto& q : d.dk) dfree(q);
    dfree(d.kex);
    dfree(d.far_list); dfree(d.far_front); dfree(d.far_cnt); dfree(d.act_list);
    dfree(d.c_visits); dfree(d.c_accn); dfree(d.c_accl); dfree(d.c_sph);
    dfree(d.rec_ent); dfree(d.rec_next); dfree(d.rec_head); d.rec_cap = 0;
    d.cap = 0;
}

int ensure_nodes(agb_ctx* c, int64_t want)
{
    AgbDev& d = c->d;
    want = std::max<int64_t>(want, d.cap);
    if (want <= d.ncap) return AGB_OK;
    if (want + d.cap >= (1ll << 31)) { c->err = "more than 2^31 particles + nodes"; return AGB_ERR_NOMEM; }
    free_nodes(c);
    const size_t nc = (size_t)want, cap = (size_t)d.cap;
    CK(dalloc(d.src_pm, cap + nc)); CK(dalloc(d.src_gv, cap + nc)); CK(dalloc(d.src_flag, cap + nc));
    CK(dalloc(d.child, 8 * nc)); CK(dalloc(d.nfirst, nc)); CK(dalloc(d.nlast, nc)); CK(dalloc(d.nparent, nc)); CK(dalloc(d.arrived, nc)); CK(dalloc(d.ndepth, nc));
    CK(dalloc(d.nmark, nc)); CK(dalloc(d.ndup, nc)); CK(dalloc(d.mom_pm, nc)); CK(dalloc(d.mom_gv, nc)); CK(dalloc(d.grouplist, nc)); CK(dalloc(d.lvl_list, nc));
    CK(dalloc(d.klo[0], nc)); CK(dalloc(d.nodebase, nc));     // per particle in the build, per node (exact sums, fold list) in the density pass
    if (c->extended) { CK(dalloc(d.quad, 6 * nc)); CK(dalloc(d.ext_bar, 1)); }
    d.ncap = (int64_t)nc;
    return AGB_OK;
}

int ensure_deep(agb_ctx* c);

int ensure_pool(agb_ctx* c, int64_t n)
{
    AgbDev& d = c->d;
    if (n <= d.cap && d.ncap > 0) return AGB_OK;
    free_pool(c);
    const size_t cap = (size_t)n;
    d.x = d.y = d.z = d.vx = d.vy = d.vz = d.mass = d.U = d.next = d.mu = nullptr; d.type = nullptr;
    CK(dalloc(d.ax, cap)); CK(dalloc(d.ay, cap)); CK(dalloc(d.az, cap)); CK(dalloc(d.dUdt, cap)); CK(dalloc(d.h, cap));
    CK(dalloc(d.rho, cap)); CK(dalloc(d.P, cap)); CK(dalloc(d.T, cap)); CK(dalloc(d.vis, cap));
    for (int i = 0; i < 2; i++) { CK(dalloc(d.khi[i], cap)); CK(dalloc(d.perm[i], cap)); }
    CK(dalloc(d.klo[1], cap));
    CK(dalloc(d.s_h, cap)); CK(dalloc(d.s_rho, cap)); CK(dalloc(d.s_P, cap)); CK(dalloc(d.s_U, cap)); CK(dalloc(d.s_mu, cap));
    CK(dalloc(d.s_next, cap)); CK(dalloc(d.s_T, cap)); CK(dalloc(d.s_type, cap));
    CK(dalloc(d.lcp, cap)); CK(dalloc(d.nodecnt, cap)); CK(dalloc(d.leafparent, cap)); CK(dalloc(d.group, cap)); CK(dalloc(d.leafdepth, cap));
    CK(dalloc(d.leafmark, cap)); CK(dalloc(d.gasrank, cap + 1));
    CK(dalloc(d.rec, cap)); CK(dalloc(d.grec, 2 * cap));
    {
        int lcap, fcap, tg; agb_far_capacity(&lcap, &fcap, &tg);
        const size_t nsg = cap / (size_t)tg + 2;
        CK(dalloc(d.far_list, nsg * lcap)); CK(dalloc(d.far_front, nsg * fcap)); CK(dalloc(d.far_cnt, nsg * 3));
        CK(dalloc(d.act_list, cap));
    }
    for (auto& q : c->in_d) CK(dalloc(q, cap));
    CK(dalloc(c->timestep, cap));
    CK(dalloc(c->in_type, cap));
    CK(dalloc(d.blockhist, agb_sort_scratch_words((int64_t)cap)));               // digit totals, tickets and look-back status words of the 8 sort passes
    CK(dalloc(d.scanblk, (cap + 2047) / 2048 + 1));
    d.cap = (int64_t)cap;
    if (d.deep) { int rc = ensure_deep(c); if (rc) return rc; }
    return ensure_nodes(c, (int64_t)cap + 1024);
}

// Three-word keys (63 levels): switched on for good when a two-word build reports particles it cannot tell apart, or more than
// 4096 of them inside one 21-level cell (a root cube blown up by a few runaway particles: Tree.cpp:85-117 has no outlier guard)
int ensure_deep(agb_ctx* c)
{
    AgbDev& d = c->d;
    if (!d.kex) {
        const size_t cap = (size_t)d.cap;
        for (auto& q : d.dk) CK(dalloc(q, cap));
        CK(dalloc(d.kex, cap));
    }
    d.deep = true;                                              // only once the buffers exist
    return AGB_OK;
}

int ensure_counters(agb_ctx* c)
{
    AgbDev& d = c->d;
    if (d.c_visits) return AGB_OK;
    CK(dalloc(d.c_visits, (size_t)d.cap)); CK(dalloc(d.c_accn, (size_t)d.cap)); CK(dalloc(d.c_accl, (size_t)d.cap)); CK(dalloc(d.c_sph, (size_t)d.cap));
    return AGB_OK;
}

// Tile records of the mixed-precision SPH pass (260 B each): at most one per 32-source tile of a group with gas targets.
// The pool starts small; when a walk runs out of records k_sph does nothing, agb_forces grows the pool to what that walk
// asked for and walks again (once per run in practice: the pool is kept across steps).
int ensure_sph_records(agb_ctx* c, int64_t want)
{
    AgbDev& d = c->d;
    want = std::min<int64_t>(std::max<int64_t>(want, std::max<int64_t>(65536, d.cap / 4)), 0x7fffff00);   // C3 needs ~0.2 records per particle: no second walk on the first gas step
    if (!d.rec_head) CK(dalloc(d.rec_head, (size_t)d.cap / 32 + 16));
    if (d.rec_ent && d.rec_cap >= want) return AGB_OK;
    dfree(d.rec_ent); dfree(d.rec_next); d.rec_cap = 0;
    CK(dalloc(d.rec_ent, (size_t)want * 32)); CK(dalloc(d.rec_next, (size_t)want));
    d.rec_cap = want;
    return AGB_OK;
}

int fetch_scalars(agb_ctx* c)
{
    const size_t off = offsetof(AgbScalars, mean);
    CK(cudaMemcpyAsync((char*)&c->hs + off, (char*)c->s + off, sizeof(AgbScalars) - off, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return AGB_OK;
}

// copy (or zero / constant fill) one caller array into an owned device array
int put_array(agb_ctx* c, double* dst, const double* src, int64_t n, int memspace)
{
    if (!src) return AGB_OK;                            // zero-filled at the start of the hand-over (st_zero)
    CK(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), memspace == AGB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->st_copy));
    return AGB_OK;
}

// positions and masses go on the compute stream (the build starts with them: the extent pass packs (x, y, z, m) records);
// everything else is uploaded on a second stream that the build only joins before it permutes the particle data
// (k_gather), so it overlaps extent + keys + sort
int own_input(agb_ctx* c, const double*& slot, int which, const double* src, int64_t n)
{
    if (!src) { slot = nullptr; return AGB_OK; }
    CK(cudaMemcpyAsync(c->in_d[which], src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, which < 3 ? c->st : c->st_copy));
    slot = c->in_d[which];
    return AGB_OK;
}

// results table order: ax ay az dUdt h rho P T visualDensity
struct ResCol { double* dst; const double* src; };
void result_table(const agb_results& r, const AgbDev& d, ResCol cp[9])
{
    ResCol t[9] = {{r.ax, d.ax}, {r.ay, d.ay}, {r.az, d.az}, {r.dUdt, d.dUdt}, {r.h, d.h}, {r.rho, d.rho}, {r.P, d.P}, {r.T, d.T}, {r.visualDensity, d.vis}};
    for (int i = 0; i < 9; i++) cp[i] = t[i];
}

// send the bound destinations of columns [first, last] on the copy stream once the compute stream has produced them
int stream_out(agb_ctx* c, int first, int last)
{
    if (!c->bres_on || c->d.n == 0) return AGB_OK;
    ResCol cp[9];
    result_table(c->bres, c->d, cp);
    bool any = false;
    for (int i = first; i <= last; i++) any = any || cp[i].dst;
    if (!any) return AGB_OK;
    CK(cudaEventRecord(c->ev_out, c->st));
    CK(cudaStreamWaitEvent(c->st_copy, c->ev_out, 0));
    const cudaMemcpyKind k = c->bres_space == AGB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    for (int i = first; i <= last; i++)
        if (cp[i].dst) { CK(cudaMemcpyAsync(cp[i].dst, cp[i].src, (size_t)c->d.n * sizeof(double), k, c->st_copy)); c->bres_sent[i] = true; }
    return AGB_OK;
}

} // namespace

extern "C" {

const char* agb_version(void) { return "agb200 0.1 (sm_100a)"; }

const char* agb_strerror(int st)
{
    switch (st) {
    case AGB_OK: return "ok";
    case AGB_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
    case AGB_ERR_CUDA: return "CUDA error";
    case AGB_ERR_INVALID: return "invalid argument or call order";
    case AGB_ERR_DEPTH: return "coincident particles: octree deeper than 63 levels";
    case AGB_ERR_UNSUPPORTED: return "parameter range not covered by the parity path";
    case AGB_ERR_NOMEM: return "out of memory / traversal stack overflow";
    }
    return "unknown status";
}

const char* agb_last_error(agb_ctx* c) { return c ? c->err.c_str() : ""; }

static void destroy_handles(agb_ctx* c)
{
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->ev_out) cudaEventDestroy(c->ev_out);
    if (c->ev_sync) cudaEventDestroy(c->ev_sync);
    if (c->ev_next) cudaEventDestroy(c->ev_next);
    if (c->ev_pos) cudaEventDestroy(c->ev_pos);
    if (c->ev_mass) cudaEventDestroy(c->ev_mass);
    if (c->ev_zero) cudaEventDestroy(c->ev_zero);
    if (c->st_zero) cudaStreamDestroy(c->st_zero);
    for (auto& e : c->evw) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->evk) if (e) cudaEventDestroy(e);
    if (c->st_copy) cudaStreamDestroy(c->st_copy);
    if (c->st) cudaStreamDestroy(c->st);
    if (c->d.spill) cudaFree(c->d.spill);
    if (c->s) cudaFree(c->s);
    if (c->stage) cudaFreeHost(c->stage);
    if (c->d_min) cudaFree(c->d_min);
    if (c->sfr) cudaFree(c->sfr);
    if (c->bs_buf) cudaFree(c->bs_buf);
    if (c->bs_idx) cudaFree(c->bs_idx);
}

int agb_create(agb_ctx** out, int device, int compat_cores)
{
    if (!out) return AGB_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return AGB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return AGB_ERR_NO_DEVICE;
    if (prop.major != 10) return AGB_ERR_NO_DEVICE;         // the kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return AGB_ERR_NO_DEVICE;
    if (getenv("AGB200_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(getenv("AGB200_L2_FETCH")));   // tuning probe: 32 / 64 / 128 bytes
    agb_ctx* c = new agb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->d.cores = compat_cores > 0 ? compat_cores : 1;
    auto fail = [&](int rc) { destroy_handles(c); delete c; (void)cudaGetLastError(); return rc; };
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) return fail(AGB_ERR_NO_DEVICE);
    if (cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking) != cudaSuccess) return fail(AGB_ERR_NO_DEVICE);
    if (cudaStreamCreateWithFlags(&c->st_zero, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_zero, cudaEventDisableTiming) != cudaSuccess) return fail(AGB_ERR_NO_DEVICE);
    if (cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_sync, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_next, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_pos, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_mass, cudaEventDisableTiming) != cudaSuccess) return fail(AGB_ERR_NO_DEVICE);
    for (auto& e : c->evw) if (cudaEventCreate(&e) != cudaSuccess) return fail(AGB_ERR_NO_DEVICE);
    for (auto& e : c->ev) if (cudaEventCreate(&e) != cudaSuccess) return fail(AGB_ERR_NO_DEVICE);
    for (auto& e : c->evk) if (cudaEventCreate(&e) != cudaSuccess) return fail(AGB_ERR_NO_DEVICE);
    c->timeline = getenv("AGB_TIMELINE") != nullptr;
    if (c->timeline) for (auto& e : c->tl) if (cudaEventCreate(&e) != cudaSuccess) return fail(AGB_ERR_NO_DEVICE);
    if (cudaMalloc((void**)&c->s, sizeof(AgbScalars)) != cudaSuccess) return fail(AGB_ERR_NOMEM);
    cudaMemsetAsync(c->s, 0, sizeof(AgbScalars), c->st);
    memset(&c->hs, 0, sizeof(c->hs));
    c->d.spill_warps = agb_walk_blocks(c->sm_count) * agb_walk_warps_per_block();
    c->d.spill_per_warp = SPILL_PER_WARP;
    if (cudaMalloc((void**)&c->d.spill, (size_t)c->d.spill_warps * SPILL_PER_WARP * sizeof(int2)) != cudaSuccess) return fail(AGB_ERR_NOMEM);
    *out = c;
    return AGB_OK;
}

int agb_destroy(agb_ctx* c)
{
    if (!c) return AGB_ERR_INVALID;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st); cudaStreamSynchronize(c->st_copy); cudaStreamSynchronize(c->st_zero);
    free_pool(c);
    destroy_handles(c);
    delete c;
    return AGB_OK;
}

int agb_set_option(agb_ctx* c, int option, int64_t value)
{
    if (!c) return AGB_ERR_INVALID;
    if (option == AGB_OPT_TARGET_COUNTERS) { c->target_counters = value != 0; return AGB_OK; }
    if (option == AGB_OPT_PRECISION) { if (value != 0 && value != 1) return AGB_ERR_INVALID; c->mixed = value == 1; return AGB_OK; }
    if (option == AGB_OPT_COOLING) { c->opt_cooling = value != 0; return AGB_OK; }
    if (option == AGB_OPT_SLICE_PIECE) { if (value < 256) return AGB_ERR_INVALID; c->piece_targets = value; return AGB_OK; }
    if (option == AGB_OPT_SLICE_DENSITIES) { c->slice_dens = value != 0; return AGB_OK; }
    if (option == AGB_OPT_EXTENDED) {
        if (value != 0 && c->d.ncap > 0 && !c->d.quad) { CK(cudaSetDevice(c->device)); CK(dalloc(c->d.quad, 6 * (size_t)c->d.ncap)); CK(dalloc(c->d.ext_bar, 1)); }
        c->extended = value != 0; c->ext_quad = value != 2;
        return AGB_OK;
    }
    if (option == AGB_OPT_STAR_FORMATION) { c->opt_sf_seed = (unsigned long long)value; return AGB_OK; }
    return AGB_ERR_INVALID;
}

int agb_set_particles(agb_ctx* c, const agb_particles* p, int memspace)
{
    if (!c || !p || p->n < 0 || p->n >= (1ll << 30) || (p->n > 0 && (!p->x || !p->y || !p->z || !p->mass || !p->type))) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const int64_t n = p->n;
    int rc = ensure_pool(c, n);
    if (rc) return rc;
    AgbDev& d = c->d;
    d.n = n;
    d.dens_a0 = 0; d.dens_a1 = INT64_MAX;
    for (auto& e : c->xev) e = nullptr;
    c->mass_late = false;
    // Everything already queued on the compute stream (densities, walk, integrator kernels of the previous step) reads or
    // writes the buffers the copy stream is about to overwrite: order the copy stream after it.
    CK(cudaEventRecord(c->ev_sync, c->st));
    CK(cudaStreamWaitEvent(c->st_copy, c->ev_sync, 0));
    // Columns the caller leaves out (carried acc / dUdt / h / rho / P / T) and the visual density start at zero.  The fills are
    // kernels: queued behind the uploads they would have to wait until the persistent walk frees an SM, and the result copies
    // queued behind THEM on the copy stream with it (measured: 20 ms).  They run at once on a stream of their own, while the
    // positions are on the link, after everything of the last step on both streams.
    CK(cudaEventRecord(c->ev_zero, c->st_copy));
    CK(cudaStreamWaitEvent(c->st_zero, c->ev_zero, 0));
    {
        double* col[8] = {d.ax, d.ay, d.az, d.dUdt, d.h, d.rho, d.P, d.T};
        const double* given[8] = {p->ax, p->ay, p->az, p->dUdt, p->h, p->rho, p->P, p->T};
        for (int k = 0; k < 8; k++) if (!given[k] && n > 0) CK(cudaMemsetAsync(col[k], 0, (size_t)n * sizeof(double), c->st_zero));
        CK(cudaMemsetAsync(d.vis, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(double), c->st_zero));
    }
    CK(cudaEventRecord(c->ev_zero, c->st_zero));
    c->zero_pending = true;                                   // device hand-over: joined before the first kernel that touches those columns
    if (memspace == AGB_MEM_DEVICE) {
        // zero-copy: the caller's device arrays are read in place (they must stay valid until the next set_particles)
        d.x = p->x; d.y = p->y; d.z = p->z; d.vx = p->vx; d.vy = p->vy; d.vz = p->vz; d.mass = p->mass; d.U = p->U; d.next = p->next_time; d.mu = p->mu;
        d.type = p->type;
        c->bound = true;
    } else {
        if (c->timeline) cudaEventRecord(c->tl[0], c->st);
        CK(cudaMemcpyAsync(c->in_type, p->type, (size_t)n, cudaMemcpyHostToDevice, c->st));   // needed by the key pass
        if ((rc = own_input(c, d.x, 0, p->x, n)) || (rc = own_input(c, d.y, 1, p->y, n)) || (rc = own_input(c, d.z, 2, p->z, n))) return rc;
        // the build starts as soon as these have landed; everything else follows on the copy stream, in the order the path
        // needs it, and only AFTER them (two concurrent host-to-device streams would share the link and delay the positions)
        CK(cudaStreamWaitEvent(c->st, c->ev_zero, 0));      // long done; the copy stream follows through ev_pos
        c->zero_pending = false;
        CK(cudaEventRecord(c->ev_pos, c->st));
        if (c->timeline) cudaEventRecord(c->tl[1], c->st);
        CK(cudaStreamWaitEvent(c->st_copy, c->ev_pos, 0));
        if ((rc = own_input(c, d.mass, 6, p->mass, n))) return rc;             // the extent, key and sort passes run without the masses
        CK(cudaEventRecord(c->ev_mass, c->st_copy));
        if (c->timeline) cudaEventRecord(c->tl[2], c->st_copy);
        c->mass_late = true;
        if ((rc = own_input(c, d.next, 8, p->next_time, n))) return rc;
        d.type = c->in_type;
        c->bound = false;
    }
    // first group: what the gather, the densities and the gravity walk read or write (active flags, carried acc / dUdt / h / rho)
    if ((rc = put_array(c, d.ax, p->ax, n, memspace)) || (rc = put_array(c, d.ay, p->ay, n, memspace)) || (rc = put_array(c, d.az, p->az, n, memspace)) ||
        (rc = put_array(c, d.dUdt, p->dUdt, n, memspace)) || (rc = put_array(c, d.h, p->h, n, memspace)) || (rc = put_array(c, d.rho, p->rho, n, memspace))) return rc;
    CK(cudaEventRecord(c->ev_next, c->st_copy));
    if (c->timeline) cudaEventRecord(c->tl[3], c->st_copy);
    // second group: only the SPH pair pass (and P, T of the density groups) needs these
    if (memspace != AGB_MEM_DEVICE) {
        if ((rc = own_input(c, d.vx, 3, p->vx, n)) || (rc = own_input(c, d.vy, 4, p->vy, n)) || (rc = own_input(c, d.vz, 5, p->vz, n)) ||
            (rc = own_input(c, d.U, 7, p->U, n)) || (rc = own_input(c, d.mu, 9, p->mu, n))) return rc;
    }
    if ((rc = put_array(c, d.P, p->P, n, memspace)) || (rc = put_array(c, d.T, p->T, n, memspace))) return rc;
    CK(cudaEventRecord(c->ev_in, c->st_copy));
    if (c->timeline) cudaEventRecord(c->tl[4], c->st_copy);
    c->in_pending = true; c->next_pending = true;
    // Host arrays are read asynchronously (pinned memory makes that a true overlap): like the reference, which reads
    // Simulation::particles during buildTree, they must stay untouched until agb_build_tree has returned.
    c->have_particles = true; c->built = false; c->dens_done = false; c->forces_done = false; c->counters_valid = false;
    c->int_ready = false;
    for (bool& f : c->bres_sent) f = false;
    return AGB_OK;
}

int agb_set_particles_staged(agb_ctx* c, const agb_particles* p, int memspace, void* ready_positions, void* ready_next_time, void* ready_all)
{
    int rc = agb_set_particles(c, p, memspace);
    if (rc) return rc;
    // The caller's arrays are still being produced (an all-gather, a copy, its own kernels) on streams of its own: each group
    // is read only after the event the caller recorded behind its producer.  The compute stream needs the first group at once.
    c->xev[0] = (cudaEvent_t)ready_positions; c->xev[1] = (cudaEvent_t)ready_next_time; c->xev[2] = (cudaEvent_t)ready_all;
    if (c->xev[0]) CK(cudaStreamWaitEvent(c->st, c->xev[0], 0));
    return AGB_OK;
}

static inline double* aos_d(void* base, int64_t off) { return reinterpret_cast<double*>(static_cast<char*>(base) + off); }

// host loops over the caller's records (264 bytes apart in the reference: latency bound), split over a few threads
static void host_parallel(int64_t n, const std::function<void(int64_t, int64_t)>& f)
{
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), (int64_t)16, n / 65536}));
    if (nt <= 1) { f(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t] { f(n * t / nt, n * (t + 1) / nt); });
    for (auto& t : th) t.join();
}

int agb_set_particles_aos(agb_ctx* c, void* const* parts, int64_t n, const agb_aos_layout* L)
{
    if (!c || !parts || !L || n < 0 || n >= (1ll << 30) || L->position < 0 || L->mass < 0 || L->type < 0) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    // the pinned staging buffer may still be the source of an upload in flight (or the target of a download)
    CK(cudaStreamSynchronize(c->st_copy)); CK(cudaStreamSynchronize(c->st));
    // gather the array-of-structs records (the reference's Particle, 264 B each) into 21 SoA columns
    const size_t cols = 21, need = cols * (size_t)std::max<int64_t>(n, 1) * sizeof(double) + (size_t)n;
    if (need > c->stage_bytes) {
        if (c->stage) cudaFreeHost(c->stage);
        c->stage = nullptr; c->stage_bytes = 0;
        CK(cudaMallocHost((void**)&c->stage, need));
        c->stage_bytes = need;
    }
    double* col[21];
    for (size_t k = 0; k < cols; k++) col[k] = c->stage + k * (size_t)std::max<int64_t>(n, 1);
    uint8_t* typ = reinterpret_cast<uint8_t*>(c->stage + cols * (size_t)std::max<int64_t>(n, 1));
    auto vec = [&](int64_t off, int64_t i, int k0) {
        if (off < 0) { col[k0][i] = col[k0 + 1][i] = col[k0 + 2][i] = 0.0; return; }
        const double* v = aos_d(parts[i], off); col[k0][i] = v[0]; col[k0 + 1][i] = v[1]; col[k0 + 2][i] = v[2];
    };
    auto sca = [&](int64_t off, int64_t i, int k, double dflt) { col[k][i] = off < 0 ? dflt : *aos_d(parts[i], off); };
    for (int64_t i = 0; i < n; i++) if (!parts[i]) return AGB_ERR_INVALID;
    host_parallel(n, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i++) {
            vec(L->position, i, 0); vec(L->velocity, i, 3); vec(L->acc, i, 6);
            sca(L->mass, i, 9, 0); sca(L->U, i, 10, 0); sca(L->next_time, i, 11, 0); sca(L->mu, i, 12, 0.58);
            sca(L->rho, i, 13, 0); sca(L->P, i, 14, 0); sca(L->T, i, 15, 0); sca(L->h, i, 16, 0); sca(L->dUdt, i, 17, 0);
            typ[i] = *reinterpret_cast<const uint8_t*>(static_cast<const char*>(parts[i]) + L->type);
        }
    });
    agb_particles p;
    memset(&p, 0, sizeof(p));
    p.n = n;
    p.x = col[0]; p.y = col[1]; p.z = col[2]; p.vx = col[3]; p.vy = col[4]; p.vz = col[5]; p.ax = col[6]; p.ay = col[7]; p.az = col[8];
    p.mass = col[9]; p.U = col[10]; p.next_time = col[11]; p.mu = col[12]; p.rho = col[13]; p.P = col[14]; p.T = col[15]; p.h = col[16]; p.dUdt = col[17];
    p.type = typ;
    return agb_set_particles(c, &p, AGB_MEM_HOST);
}

// the kernels of Tree::buildTree on the context's stream (no synchronisation)
// the compute stream waits for every group of the last hand-over (the library's own uploads and the caller's staged events)
static void join_zero(agb_ctx* c)
{
    if (c->zero_pending) { cudaStreamWaitEvent(c->st, c->ev_zero, 0); c->zero_pending = false; }
}

static void join_uploads(agb_ctx* c)
{
    join_zero(c);
    if (!c->in_pending) return;
    cudaStreamWaitEvent(c->st, c->ev_in, 0);
    for (int k = 1; k < 3; k++) if (c->xev[k]) cudaStreamWaitEvent(c->st, c->xev[k], 0);
    c->in_pending = false; c->next_pending = false;
}

// late_gas (agb_force_path, host hand-over, mixed precision, gas): the build only waits for the first upload group; the
// tree's gas velocities are completed by launch_late_gas once the second group is there
static void launch_build(agb_ctx* c, bool late_gas = false)
{
    AgbDev& d = c->d;
    cudaEventRecord(c->evk[4], c->st);
    c->launches += agb_launch_extent(d, c->s, c->st, c->mass_late);
    c->launches += agb_launch_keygen(d, c->s, c->st);
    cudaEventRecord(c->evk[5], c->st);
    c->launches += agb_launch_sort(d, c->s, c->st);
    cudaEventRecord(c->evk[6], c->st);
    if (c->mass_late) { cudaStreamWaitEvent(c->st, c->ev_mass, 0); c->launches += agb_launch_fill_mass(d, c->st); c->mass_late = false; }
    join_zero(c);
    if (!late_gas && c->in_pending) join_uploads(c);
    c->launches += agb_launch_links(d, c->s, c->st, &c->evk[7], late_gas);
    if (late_gas) {
        // next_time (and the carried columns the densities and the walk write into) arrive behind the masses: nothing of the
        // build reads them, so the tree-order copy of next_time is made here, after the build, instead of stalling the gather
        if (c->next_pending) { cudaStreamWaitEvent(c->st, c->ev_next, 0); if (c->xev[1]) cudaStreamWaitEvent(c->st, c->xev[1], 0); c->next_pending = false; }
        c->launches += agb_launch_gather_next(d, c->s, c->st);
    }
    cudaEventRecord(c->evk[9], c->st);
    c->build_timed = true;
}

// more (first particle, depth) pairs than node slots: the step wrote nothing past the capacity; make room for a rebuild
static int grow_nodes(agb_ctx* c)
{
    const int64_t m = c->hs.n_nodes;
    return ensure_nodes(c, m + m / 8 + 1024);
}

int agb_build_tree(agb_ctx* c, double* root_radius)
{
    if (!c || !c->have_particles) return AGB_ERR_INVALID;
    Phase ph("build tree");
    CK(cudaSetDevice(c->device));
    AgbDev& d = c->d;
    c->built = false; c->dens_done = false; c->forces_done = false; c->counters_valid = false;
    if (d.n == 0) { if (root_radius) *root_radius = 0.0; memset((char*)&c->hs + offsetof(AgbScalars, mean), 0, sizeof(AgbScalars) - offsetof(AgbScalars, mean)); c->built = true; return AGB_OK; }
    for (int attempt = 0;; attempt++) {
        CK(cudaEventRecord(c->ev[0], c->st));
        launch_build(c);
        CK(cudaEventRecord(c->ev[1], c->st));
        CK(cudaGetLastError());
        int rc = fetch_scalars(c);
        if (rc) return rc;
        if (c->hs.need_deep && !d.deep) { if ((rc = ensure_deep(c))) return rc; attempt = -1; continue; }   // rebuild with three-word keys
        if (c->hs.n_nodes <= d.ncap) break;
        if (attempt >= 1) { c->err = "node table overflow"; return AGB_ERR_NOMEM; }
        if ((rc = grow_nodes(c))) return rc;
    }
    float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->phase_ms[0] = ms;
    c->hs.R = 0; memcpy(&c->hs.R, &c->hs.Rbits, 8);
    if (root_radius) *root_radius = c->hs.R;
    if (c->hs.dup_keys > 0) { c->err = "two particles share all 63 octree levels (coincident points)"; return AGB_ERR_DEPTH; }
    c->built = true;
    return AGB_OK;
}

int agb_visual_density(agb_ctx* c, double radius)
{
    if (!c || !c->built) return AGB_ERR_INVALID;
    Phase ph("Visual Density");
    CK(cudaSetDevice(c->device));
    if (c->d.n == 0) return AGB_OK;
    CK(cudaEventRecord(c->ev[2], c->st));
    c->launches += agb_launch_visual(c->d, c->s, radius, c->st);
    CK(cudaEventRecord(c->ev[3], c->st));
    c->vis_timed = true;
    CK(cudaGetLastError());
    return stream_out(c, 8, 8);
}

static int gas_density_impl(agb_ctx* c, double mass_in_h, bool late_pt);
int agb_gas_density(agb_ctx* c, double mass_in_h) { return gas_density_impl(c, mass_in_h, false); }

static int gas_density_impl(agb_ctx* c, double mass_in_h, bool late_pt)
{
    if (!c || !c->built) return AGB_ERR_INVALID;
    Phase ph("SPH density and update");
    CK(cudaSetDevice(c->device));
    c->dens_done = true;
    if (c->d.n == 0) return AGB_OK;
    CK(cudaEventRecord(c->ev[4], c->st));
    if (c->hs.any_gas && c->extended) c->launches += agb_launch_extended_density(c->d, c->s, mass_in_h, c->st);
    else if (c->hs.any_gas) c->launches += agb_launch_gas_density(c->d, c->s, mass_in_h, c->st, late_pt);
    CK(cudaEventRecord(c->ev[5], c->st));
    c->gas_timed = true;
    CK(cudaGetLastError());
    return stream_out(c, 4, late_pt ? 5 : 7);                // P and T follow after the late part of the build
}

static int forces_impl(agb_ctx* c, double global_time, double e0, double theta, int part, int nparts, bool late_gas);

// The FP32 pair law of the mixed mode works in units of R / 2^16 and keeps r^2 (r^2 + e0^2)^2 inside the FP32 range for
// separations down to ~1e-12 R and softening lengths between ~1e-10 R and ~50 R (agb_walk.cu).  A root cube blown up by runaway
// particles (or an exotic unit system) leaves that range: such steps are walked with FP64 pair arithmetic instead.
static bool mixed_in_range(const agb_ctx* c, double e0)
{
    if (!c->mixed) return false;
    double R = 0; memcpy(&R, &c->hs.Rbits, 8);
    if (!(R > 0.0)) return true;
    return e0 >= 1e-10 * R && e0 <= 50.0 * R && c->hs.max_depth <= 40;
}
int agb_forces_slice(agb_ctx* c, double global_time, double e0, double theta, int part, int nparts) { return forces_impl(c, global_time, e0, theta, part, nparts, false); }

// late_gas: the tree was built without the gas velocities / U / mu (launch_build(c, true)); they are folded in between the
// gravity walk and the SPH pair pass, by which time their upload has long finished
static int forces_impl(agb_ctx* c, double global_time, double e0, double theta, int part, int nparts, bool late_gas)
{
    if (!c || !c->built || nparts < 1 || part < 0 || part >= nparts) return AGB_ERR_INVALID;
    // For e0 <= 2.1474836e13 the reference's `abs(e)` (int abs(int), Node.cpp:302,354) can switch to the spline
    // softening length; that branch is not reproduced (dead for every SI configuration).
    if (!(e0 > 2.147483648e13)) { c->err = "e0 <= 2^31 * 1e4: the reference's int-abs softening branch is not covered"; return AGB_ERR_UNSUPPORTED; }
    Phase ph("Force Calculation");
    CK(cudaSetDevice(c->device));
    AgbDev& d = c->d;
    if (d.n == 0) { c->forces_done = true; return AGB_OK; }
    if (c->target_counters) { int rc = ensure_counters(c); if (rc) return rc; }
    const bool mixed = !c->extended && mixed_in_range(c, e0);  // with the scalars of the tree the walk will run on (call by call), or of the last step's (agb_force_path)
    c->walked_mixed = mixed;
    if (c->hs.any_gas && mixed) { int rc = ensure_sph_records(c, 0); if (rc) return rc; }
    // The targets are the ACTIVE particles in tree order; slice boundaries fall on multiples of 256 of them (the far-field
    // super-groups), so every warp owns the same 32 targets whatever the number of parts: results are bit-identical for
    // 1, 2, 4, 8 GPUs (same groups => same summation order).  The slicing itself happens on the device (agb_walk.cu).
    CK(cudaEventRecord(c->ev[6], c->st));
    // gas targets need h/rho/P: if the caller skipped gas_density they are orphans (h = 0) and get no SPH, like the reference
    const bool any_gas = c->hs.any_gas != 0;
    for (int attempt = 0;; attempt++) {
        if (c->extended)
            c->launches += agb_launch_extended_forces(d, c->s, global_time, e0, theta, part, nparts, any_gas, c->ext_quad, c->sm_count, c->st, c->evw);
        else if (late_gas) {
            c->launches += agb_launch_walk(d, c->s, global_time, e0, theta, part, nparts, c->target_counters, any_gas, mixed, c->sm_count, c->st, c->evw, 1);
            if (attempt == 0) {
                join_uploads(c);
                c->launches += agb_launch_late_gas(d, c->s, c->st);
                int rc = stream_out(c, 6, 7);
                if (rc) return rc;
            }
            c->launches += agb_launch_walk(d, c->s, global_time, e0, theta, part, nparts, c->target_counters, any_gas, mixed, c->sm_count, c->st, c->evw, 2);
        } else
            c->launches += agb_launch_walk(d, c->s, global_time, e0, theta, part, nparts, c->target_counters, any_gas, mixed, c->sm_count, c->st, c->evw, 0);
        CK(cudaEventRecord(c->ev[7], c->st));
        CK(cudaGetLastError());
        int rc = fetch_scalars(c);
        if (rc) return rc;
        if (c->hs.walk_overflow != 2 || attempt >= 2) break;
        // out of SPH tile records: nothing but acc was written (k_sph skipped itself); grow the pool and walk again
        rc = ensure_sph_records(c, (int64_t)(c->hs.cand_cursor + c->hs.cand_cursor / 4) + 1024);
        if (rc) return rc;
    }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->evw[0], c->evw[3]) == cudaSuccess) c->phase_ms[3] = ms;
    if (cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]) == cudaSuccess) c->phase_ms[4] = ms;
    for (int k = 0; k < 2; k++) if (cudaEventElapsedTime(&ms, c->evw[k], c->evw[k + 1]) == cudaSuccess) c->kernel_ms[k] = ms;
    if (cudaEventElapsedTime(&ms, c->evw[4], c->evw[3]) == cudaSuccess) c->kernel_ms[2] = ms;
    if (c->build_timed) for (int k = 0; k < 5; k++) if (cudaEventElapsedTime(&ms, c->evk[4 + k], c->evk[5 + k]) == cudaSuccess) c->kernel_ms[3 + k] = ms;
    if (c->vis_timed && cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]) == cudaSuccess) c->phase_ms[1] = ms;
    if (c->gas_timed && cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess) c->phase_ms[2] = ms;
    (void)cudaGetLastError();
    if (c->hs.walk_overflow == 3) { c->forces_done = false; return AGB_OK; }     // sliced densities, but some particles rest: nothing was walked, the caller redoes the step
    if (c->hs.walk_overflow) { c->err = c->hs.walk_overflow == 2 ? "SPH tile-record pool overflow" : "traversal stack overflow"; return AGB_ERR_NOMEM; }
    c->forces_done = true; c->counters_valid = c->target_counters;
    return AGB_OK;
}

int agb_forces(agb_ctx* c, double global_time, double e0, double theta) { return agb_forces_slice(c, global_time, e0, theta, 0, 1); }

// The four calls in one, with a single host synchronisation at the end.  Two things the separate calls learn from the
// device in between are taken from the previous step instead: the visual-density radius is the caller's (the reference
// fixes it at init, Simulation.cpp:126) and "the particle set holds gas" (which decides whether the density kernels and the
// SPH variant of the walk run) is verified after the fact; if it changed, the step is simply redone call by call.
// columns [first, last] (results table order: ax ay az dUdt h rho P T vis) of the bound slice: compact on `st`, copy out on `out`
// `off` = position of target a0 inside the bound slice (pieces of a slice are sent as they finish)
static int send_slice_columns(agb_ctx* c, int64_t a0, int64_t a1, int first, int last, bool with_index, cudaStream_t out, int64_t off = 0)
{
    const int64_t cnt = a1 - a0;
    if (cnt <= 0) return AGB_OK;
    double* host[9] = {c->bs_res.ax, c->bs_res.ay, c->bs_res.az, c->bs_res.dUdt, c->bs_res.h, c->bs_res.rho, c->bs_res.P, c->bs_res.T, c->bs_res.visualDensity};
    double* dev[9];
    bool any = with_index && c->bs_index;
    for (int k = 0; k < 9; k++) { dev[k] = (k >= first && k <= last && host[k]) ? c->bs_buf + (size_t)k * c->bs_cap + off : nullptr; any = any || dev[k]; }
    if (!any) return AGB_OK;
    c->launches += agb_launch_slice_results(c->d, a0, a1, true, (with_index && c->bs_index) ? c->bs_idx + off : nullptr, dev, c->st);
    if (out != c->st) { CK(cudaEventRecord(c->ev_out, c->st)); CK(cudaStreamWaitEvent(out, c->ev_out, 0)); }
    if (with_index && c->bs_index) CK(cudaMemcpyAsync(c->bs_index + off, c->bs_idx + off, (size_t)cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, out));
    for (int k = 0; k < 9; k++) if (dev[k]) CK(cudaMemcpyAsync(host[k] + off, dev[k], (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, out));
    return AGB_OK;
}

int agb_bind_slice_results(agb_ctx* c, int part, int nparts, uint32_t* index, const agb_results* r)
{
    if (!c || (r && (nparts < 1 || part < 0 || part >= nparts))) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st_copy));                 // nothing may still be flowing to the old destinations
    c->bs_on = r != nullptr;
    if (r) { c->bs_res = *r; c->bs_index = index; c->bs_part = part; c->bs_nparts = nparts; }
    return AGB_OK;
}

static int force_path_impl(agb_ctx* c, double visual_density_radius, double mass_in_h, double global_time, double e0, double theta, int part, int nparts, double* root_radius);

int agb_force_path(agb_ctx* c, double visual_density_radius, double mass_in_h, double global_time, double e0, double theta, int part, int nparts, double* root_radius)
{
    if (!c || !c->have_particles || nparts < 1 || part < 0 || part >= nparts) return AGB_ERR_INVALID;
    c->bs_early = false; c->bs_piped = false;
    int rc = force_path_impl(c, visual_density_radius, mass_in_h, global_time, e0, theta, part, nparts, root_radius);
    if (rc || !c->bs_on || part != c->bs_part || nparts != c->bs_nparts) return rc;
    // bound slice results: the density columns left during the walk when the step ran fused with every particle active (as the
    // last step had it); whatever is still missing goes now
    CK(cudaSetDevice(c->device));
    if (c->bs_early && c->hs.n_active == c->d.n) {
        int64_t a0 = 0, a1 = 0;
        agb_slice_bounds(c->d.n, part, nparts, &a0, &a1);
        if (!c->bs_piped) {
            if ((rc = send_slice_columns(c, a0, a1, 0, 3, false, c->st))) return rc;
            if (c->bs_late && (rc = send_slice_columns(c, a0, a1, 6, 7, false, c->st))) return rc;      // P, T exist only after the late part of the build
        }
        if (c->timeline) { cudaEventRecord(c->tl[5], c->st); cudaEventRecord(c->tl[6], c->st_copy); }
        CK(cudaStreamSynchronize(c->st)); CK(cudaStreamSynchronize(c->st_copy));
        if (c->timeline) {
            auto at = [&](cudaEvent_t e) { float ms = -1.f; if (cudaEventElapsedTime(&ms, c->ev[8], e) != cudaSuccess) { (void)cudaGetLastError(); ms = -1.f; } return ms; };
            fprintf(stderr, "agb timeline, bound slice %d/%d [ms after the build began]: build ..%.2f  visual density ..%.2f  gas density ..%.2f  last piece: walk %.2f (k_sph %.2f) ..%.2f |"
                            " delivered on the compute stream %.2f, on the copy stream %.2f\n",
                    part, nparts, at(c->ev[9]), at(c->ev[3]), at(c->ev[5]), at(c->evw[0]), at(c->evw[4]), at(c->evw[3]), at(c->tl[5]), at(c->tl[6]));
        }
        return AGB_OK;
    }
    CK(cudaStreamSynchronize(c->st_copy));
    return agb_get_slice_results_all(c, part, nparts, c->bs_index, &c->bs_res, AGB_MEM_HOST);
}

static int force_path_impl(agb_ctx* c, double visual_density_radius, double mass_in_h, double global_time, double e0, double theta, int part, int nparts, double* root_radius)
{
    auto stepwise = [&]() -> int {
        double R = 0.0;
        c->bs_early = false; c->bs_piped = false;               // whatever left early came from an abandoned attempt
        c->d.dens_a0 = 0; c->d.dens_a1 = INT64_MAX;
        int rc = agb_build_tree(c, &R);
        if (rc) return rc;
        if (root_radius) *root_radius = R;
        if ((rc = agb_visual_density(c, visual_density_radius))) return rc;
        if ((rc = agb_gas_density(c, mass_in_h))) return rc;
        if ((rc = agb_forces_slice(c, global_time, e0, theta, part, nparts))) return rc;
        c->gas_hint_valid = true; c->gas_hint = c->hs.any_gas != 0;
        return AGB_OK;
    };
    if (c->d.n == 0 || !c->gas_hint_valid) return stepwise();
    if (!(e0 > 2.147483648e13)) { c->err = "e0 <= 2^31 * 1e4: the reference's int-abs softening branch is not covered"; return AGB_ERR_UNSUPPORTED; }
    CK(cudaSetDevice(c->device));
    AgbDev& d = c->d;
    c->built = false; c->dens_done = false; c->forces_done = false; c->counters_valid = false;
    // Host hand-over still uploading (set_particles is asynchronous): in mixed precision only the SPH pair pass needs the gas
    // velocities / U / mu, so the build, the densities and the gravity walk start on the first upload group and the rest of
    // the transfer hides behind them.
    // (A staged hand-over whose next_time and last group share one event has no late group: the build then joins it before the
    // gather, which hides its production behind extent, keys and sort without the extra kernels of the late path.)
    const bool late_gas = c->in_pending && (!c->bound || (c->xev[2] && c->xev[2] != c->xev[1])) && !c->extended && mixed_in_range(c, e0) && c->gas_hint;
    CK(cudaEventRecord(c->ev[8], c->st));
    { Phase ph("build tree"); launch_build(c, late_gas); }
    CK(cudaEventRecord(c->ev[9], c->st));
    CK(cudaGetLastError());
    c->built = true;
    c->hs.any_gas = c->gas_hint ? 1 : 0;
    int rc;
    // AGB_OPT_SLICE_DENSITIES: density outputs for this slice's targets only — as long as every particle is a target (as in the
    // last step; verified for this one below), a slice of the targets is a range of tree positions
    const bool dens_sliced = c->slice_dens && nparts > 1 && c->hs.n_active == d.n && d.n > 0 && !c->extended;
    d.dens_a0 = 0; d.dens_a1 = INT64_MAX;
    if (dens_sliced) agb_slice_bounds(d.n, part, nparts, &d.dens_a0, &d.dens_a1);
    if ((rc = agb_visual_density(c, visual_density_radius))) return rc;
    if ((rc = gas_density_impl(c, mass_in_h, late_gas))) return rc;
    if (c->bs_on && part == c->bs_part && nparts == c->bs_nparts && c->hs.n_active == d.n && d.n > 0) {
        // bound slice results, every particle a target in the last step (verified for this one afterwards): index and density
        // columns of the slice leave on the copy stream while the walk runs
        int64_t a0 = 0, a1 = 0;
        agb_slice_bounds(d.n, part, nparts, &a0, &a1);
        if (a1 - a0 + 1 > c->bs_cap) {
            dfree(c->bs_buf); dfree(c->bs_idx); c->bs_cap = 0;
            const int64_t cap = a1 - a0 + 1024;
            CK(dalloc(c->bs_buf, 9 * (size_t)cap)); CK(dalloc(c->bs_idx, (size_t)cap));
            c->bs_cap = cap;
        }
        c->bs_late = late_gas;
        if ((rc = send_slice_columns(c, a0, a1, 4, late_gas ? 5 : 7, true, c->st_copy))) return rc;
        if ((rc = send_slice_columns(c, a0, a1, 8, 8, false, c->st_copy))) return rc;
        c->bs_early = true;
    }
    const int64_t my_targets = d.n / nparts;
    const int pieces = c->bs_early ? (int)std::min<int64_t>(4, std::max<int64_t>(1, my_targets / c->piece_targets)) : 1;
    if (pieces > 1) {
        // Bound slice results of a large slice: the slice is walked in `pieces` sub-slices (exact sub-ranges of it: same groups, same
        // bits) and the acc / dU/dt of each piece leave on the copy stream while the next piece walks.
        int64_t a0 = 0, a1 = 0;
        agb_slice_bounds(d.n, part, nparts, &a0, &a1);
        unsigned long long sums[7] = {0, 0, 0, 0, 0, 0, 0};
        bool sent_all = true;
        for (int k = 0; k < pieces; k++) {
            rc = forces_impl(c, global_time, e0, theta, part * pieces + k, nparts * pieces, late_gas && k == 0);
            if (rc) break;
            unsigned long long* f[7] = {&c->hs.c_node, &c->hs.c_leaf, &c->hs.c_sph, &c->hs.c_visits, &c->hs.c_exact, &c->hs.c_spill, &c->hs.c_interactions};
            for (int q = 0; q < 7; q++) sums[q] += *f[q];
            if (c->hs.n_active != d.n || c->hs.node_overflow || c->hs.need_deep) { sent_all = false; continue; }
            int64_t b0 = 0, b1 = 0;
            agb_slice_bounds(d.n, part * pieces + k, nparts * pieces, &b0, &b1);
            if ((rc = send_slice_columns(c, b0, b1, 0, 3, false, c->st_copy, b0 - a0))) return rc;
            if (c->bs_late && (rc = send_slice_columns(c, b0, b1, 6, 7, false, c->st_copy, b0 - a0))) return rc;
        }
        unsigned long long* f[7] = {&c->hs.c_node, &c->hs.c_leaf, &c->hs.c_sph, &c->hs.c_visits, &c->hs.c_exact, &c->hs.c_spill, &c->hs.c_interactions};
        for (int q = 0; q < 7; q++) *f[q] = sums[q];
        c->bs_piped = sent_all && rc == AGB_OK;
    } else
    rc = forces_impl(c, global_time, e0, theta, part, nparts, late_gas);     // ends with the step's only synchronisation
    float ms = 0; if (cudaEventElapsedTime(&ms, c->ev[8], c->ev[9]) == cudaSuccess) c->phase_ms[0] = ms;
    (void)cudaGetLastError();
    c->hs.R = 0; memcpy(&c->hs.R, &c->hs.Rbits, 8);
    if (root_radius) *root_radius = c->hs.R;
    d.dens_a0 = 0; d.dens_a1 = INT64_MAX;
    if (dens_sliced && c->hs.n_active != d.n) { c->built = false; c->forces_done = false; return stepwise(); }   // some particles rest: redo with all densities
    if (c->hs.need_deep && !d.deep) { c->built = false; c->forces_done = false; return stepwise(); }   // agb_build_tree switches to three-word keys
    if (c->walked_mixed && !mixed_in_range(c, e0)) { c->built = false; c->forces_done = false; return stepwise(); }   // this step's tree left the FP32 range: redo in FP64
    if (c->hs.n_nodes > d.ncap) {                                   // node table overflow: every kernel after the node count returned at once
        c->built = false; c->forces_done = false;
        if ((rc = grow_nodes(c))) return rc;
        return stepwise();
    }
    if (c->hs.dup_keys > 0) { c->built = false; c->forces_done = false; c->err = "two particles share all 63 octree levels (coincident points)"; return AGB_ERR_DEPTH; }
    if ((c->hs.any_gas != 0) != c->gas_hint) { c->gas_hint_valid = false; return stepwise(); }   // gas appeared / vanished: redo with the right kernels
    return rc;
}

int agb_get_counters(agb_ctx* c, agb_counters* o)
{
    if (!c || !o) return AGB_ERR_INVALID;
    const AgbScalars& h = c->hs;
    memset(o, 0, sizeof(*o));
    o->n_particles = c->d.n; o->n_in_tree = h.n_in_tree; o->n_outliers = h.n_outliers; o->n_nodes = h.n_nodes; o->n_active = h.n_active;
    o->max_depth = h.max_depth; o->edge_dropped = h.edge_dropped;
    o->node_interactions = (int64_t)h.c_node; o->leaf_interactions = (int64_t)h.c_leaf; o->interactions = (int64_t)(h.c_node + h.c_leaf);
    o->sph_interactions = (int64_t)h.c_sph; o->node_visits = c->counters_valid ? (int64_t)h.c_visits : -1; o->mac_exact_fallbacks = (int64_t)h.c_exact;
    o->groups = (c->d.n + 31) / 32; o->gas_groups = h.n_gas_groups; o->gas_orphans = h.n_gas_orphans; o->gas_ties_exact = h.tie_exact; o->gas_ties_unresolved = h.tie_unresolved;
    o->walk_rounds = (int64_t)h.st_rounds; o->walk_popped = (int64_t)h.st_popped; o->walk_straddling = (int64_t)h.st_mixed; o->walk_opened = (int64_t)h.st_open;
    o->walk_tiles = (int64_t)h.st_drain; o->walk_stack_spills = (int64_t)h.c_spill;
    o->walk_ent_wide = (int64_t)h.st_cls[0]; o->walk_ent_half = (int64_t)h.st_cls[1]; o->walk_ent_quarter = (int64_t)h.st_cls[2];
    o->walk_bits_wide = (int64_t)h.st_cls[3]; o->walk_bits_half = (int64_t)h.st_cls[4]; o->walk_bits_quarter = (int64_t)h.st_cls[5];
    o->walk_ent_far = (int64_t)h.st_cls[6];
    o->walk_ent_class0 = (int64_t)h.st_cls[7]; o->walk_ent_class1 = (int64_t)h.st_cls[8]; o->walk_ent_class2 = (int64_t)h.st_cls[9];
    o->sph_records = (int64_t)h.cand_cursor;
    return AGB_OK;
}

int agb_get_results(agb_ctx* c, const agb_results* r, int memspace)
{
    if (!c || !r || !c->have_particles) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const size_t b = (size_t)c->d.n * sizeof(double);
    const cudaMemcpyKind k = memspace == AGB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    join_uploads(c);
    ResCol cp[9], bp[9];
    result_table(*r, c->d, cp);
    result_table(c->bres, c->d, bp);
    bool streamed = false;
    for (int i = 0; i < 9; i++) {
        // a column already on its way to this very destination (agb_bind_results) is not sent twice
        if (c->bres_on && c->bres_sent[i] && cp[i].dst == bp[i].dst && memspace == c->bres_space) { streamed = streamed || cp[i].dst; continue; }
        if (cp[i].dst && b) CK(cudaMemcpyAsync(cp[i].dst, cp[i].src, b, k, c->st));
    }
    if (c->timeline) { cudaEventRecord(c->tl[5], c->st); cudaEventRecord(c->tl[6], c->st_copy); }
    CK(cudaStreamSynchronize(c->st));
    if (streamed) CK(cudaStreamSynchronize(c->st_copy));
    if (c->timeline && !c->bound) {
        CK(cudaStreamSynchronize(c->st_copy));
        auto at = [&](cudaEvent_t e) { float ms = -1.f; if (cudaEventElapsedTime(&ms, c->tl[0], e) != cudaSuccess) { (void)cudaGetLastError(); ms = -1.f; } return ms; };
        fprintf(stderr, "agb timeline [ms after the hand-over began]: positions %.2f  masses %.2f  first group %.2f  all uploads %.2f | build %.2f..%.2f  visual density ..%.2f  gas density %.2f..%.2f"
                        "  walk %.2f (k_walk %.2f, k_sph %.2f) ..%.2f | results on the compute stream %.2f, on the copy stream %.2f\n",
                at(c->tl[1]), at(c->tl[2]), at(c->tl[3]), at(c->tl[4]), at(c->ev[8]), at(c->ev[9]), at(c->ev[3]), at(c->ev[4]), at(c->ev[5]),
                at(c->evw[0]), at(c->evw[1]), at(c->evw[4]), at(c->evw[3]), at(c->tl[5]), at(c->tl[6]));
    }
    return AGB_OK;
}

int agb_bind_results(agb_ctx* c, const agb_results* r, int memspace)
{
    if (!c) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st_copy));                 // nothing may still be flowing to the old destinations
    c->bres_on = r != nullptr;
    if (r) { c->bres = *r; c->bres_space = memspace; }
    for (bool& f : c->bres_sent) f = false;
    return AGB_OK;
}

int agb_get_slice_count(agb_ctx* c, int part, int nparts, int64_t* count)
{
    if (!c || !count || !c->forces_done || nparts < 1 || part < 0 || part >= nparts) return AGB_ERR_INVALID;
    int64_t a0 = 0, a1 = 0;
    if (c->d.n > 0) agb_slice_bounds(c->hs.n_active, part, nparts, &a0, &a1);
    *count = a1 - a0;
    return AGB_OK;
}

int agb_get_slice_results_all(agb_ctx* c, int part, int nparts, uint32_t* index, const agb_results* r, int memspace)
{
    int64_t cnt = 0;
    if (!r || agb_get_slice_count(c, part, nparts, &cnt) != AGB_OK) return AGB_ERR_INVALID;
    if (cnt == 0) return AGB_OK;
    CK(cudaSetDevice(c->device));
    AgbDev& d = c->d;
    int64_t a0 = 0, a1 = 0;
    agb_slice_bounds(c->hs.n_active, part, nparts, &a0, &a1);
    const bool ident = c->hs.n_active == d.n;
    double* want[9] = {r->ax, r->ay, r->az, r->dUdt, r->h, r->rho, r->P, r->T, r->visualDensity};
    if (memspace == AGB_MEM_DEVICE) {
        c->launches += agb_launch_slice_results(d, a0, a1, ident, index, want, c->st);
        CK(cudaStreamSynchronize(c->st));
        return AGB_OK;
    }
    // host destination: compact on the device first, then D2H.  Scratch that is idle between the walk and the next build:
    // rec (4 doubles per particle), the idle halves of the sort's ping-pong (key_hi, key_lo) and the per-node moments.
    double* pool[9] = {reinterpret_cast<double*>(d.rec), reinterpret_cast<double*>(d.rec) + cnt, reinterpret_cast<double*>(d.rec) + 2 * cnt, reinterpret_cast<double*>(d.rec) + 3 * cnt,
                       reinterpret_cast<double*>(d.khi[d.cur ^ 1]), reinterpret_cast<double*>(d.klo[0]), reinterpret_cast<double*>(d.mom_pm), reinterpret_cast<double*>(d.mom_pm) + cnt,
                       reinterpret_cast<double*>(d.mom_gv)};
    double* dev[9];
    for (int k = 0; k < 9; k++) dev[k] = want[k] ? pool[k] : nullptr;
    uint32_t* idx = reinterpret_cast<uint32_t*>(d.nodecnt);
    c->launches += agb_launch_slice_results(d, a0, a1, ident, index ? idx : nullptr, dev, c->st);
    if (index) CK(cudaMemcpyAsync(index, idx, (size_t)cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st));
    for (int k = 0; k < 9; k++) if (want[k]) CK(cudaMemcpyAsync(want[k], dev[k], (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaGetLastError());
    return AGB_OK;
}

int agb_get_slice_results(agb_ctx* c, int part, int nparts, uint32_t* index, double* ax, double* ay, double* az, double* dUdt, int memspace)
{
    agb_results r;
    memset(&r, 0, sizeof(r));
    r.ax = ax; r.ay = ay; r.az = az; r.dUdt = dUdt;
    return agb_get_slice_results_all(c, part, nparts, index, &r, memspace);
}

int agb_get_results_aos(agb_ctx* c, void* const* parts, int64_t n, const agb_aos_layout* L)
{
    if (!c || !parts || !L || n != c->d.n) return AGB_ERR_INVALID;
    const size_t nn = (size_t)std::max<int64_t>(n, 1), need = 9 * nn * sizeof(double);
    if (need > c->stage_bytes) {
        if (c->stage) cudaFreeHost(c->stage);
        c->stage = nullptr; c->stage_bytes = 0;
        CK(cudaMallocHost((void**)&c->stage, need));
        c->stage_bytes = need;
    }
    double* col[9];
    for (int k = 0; k < 9; k++) col[k] = c->stage + (size_t)k * nn;
    agb_results r = {col[0], col[1], col[2], col[3], col[4], col[5], col[6], col[7], col[8]};
    int rc = agb_get_results(c, &r, AGB_MEM_HOST);
    if (rc) return rc;
    host_parallel(n, [&](int64_t a0, int64_t b0) {
        for (int64_t i = a0; i < b0; i++) {
            if (L->acc >= 0) { double* a = aos_d(parts[i], L->acc); a[0] = col[0][i]; a[1] = col[1][i]; a[2] = col[2][i]; }
            if (L->dUdt >= 0) *aos_d(parts[i], L->dUdt) = col[3][i];
            if (L->h >= 0) *aos_d(parts[i], L->h) = col[4][i];
            if (L->rho >= 0) *aos_d(parts[i], L->rho) = col[5][i];
            if (L->P >= 0) *aos_d(parts[i], L->P) = col[6][i];
            if (L->T >= 0) *aos_d(parts[i], L->T) = col[7][i];
            if (L->visualDensity >= 0) *aos_d(parts[i], L->visualDensity) = col[8][i];
        }
    });
    return AGB_OK;
}

int agb_get_tree_particles(agb_ctx* c, int32_t* leafdepth, uint64_t* key_hi, uint64_t* key_lo)
{
    if (!c || !c->built || !leafdepth || !key_hi || !key_lo) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const int64_t n = c->d.n;
    if (n == 0) return AGB_OK;
    DevTmp<int32_t> dl; DevTmp<uint64_t> dh, dlo;                       // freed on every return path
    CK(dalloc(dl.p, (size_t)n)); CK(dalloc(dh.p, (size_t)n)); CK(dalloc(dlo.p, (size_t)n));
    c->launches += agb_launch_dump_tree(c->d, c->s, dl.p, dh.p, dlo.p, c->st);
    CK(cudaMemcpyAsync(leafdepth, dl.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(key_hi, dh.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(key_lo, dlo.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return AGB_OK;
}

int agb_get_node_count(agb_ctx* c, int64_t* n_internal)
{
    if (!c || !c->built || !n_internal) return AGB_ERR_INVALID;
    *n_internal = c->hs.n_nodes;
    return AGB_OK;
}

int agb_get_nodes(agb_ctx* c, int32_t* depth, int64_t* count, int32_t* duplicated, uint64_t* key_hi, uint64_t* key_lo,
                  double* mass, double* comx, double* comy, double* comz, double* gas_mass, double* mvx, double* mvy, double* mvz)
{
    if (!c || !c->built) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const int64_t m = c->hs.n_nodes, n = c->d.n;
    if (m == 0) return AGB_OK;
    const AgbDev& d = c->d;
    std::vector<int8_t> dep((size_t)m); std::vector<int32_t> first((size_t)m), last((size_t)m); std::vector<uint8_t> dup((size_t)m);
    std::vector<double4> pm((size_t)m), gv((size_t)m); std::vector<uint64_t> khi((size_t)n), klo((size_t)n);
    CK(cudaMemcpyAsync(dep.data(), d.ndepth, (size_t)m, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(first.data(), d.nfirst, (size_t)m * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(last.data(), d.nlast, (size_t)m * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(dup.data(), d.ndup, (size_t)m, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(pm.data(), d.src_pm + n, (size_t)m * sizeof(double4), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(gv.data(), d.src_gv + n, (size_t)m * sizeof(double4), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(khi.data(), d.khi[d.cur], (size_t)n * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(klo.data(), d.klo[1], (size_t)n * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    for (int64_t k = 0; k < m; k++) {
        const int dp = dep[(size_t)k];
        uint64_t h = khi[(size_t)first[(size_t)k]], l = klo[(size_t)first[(size_t)k]];
        if (dp <= 0) { h = 0; l = 0; }
        else if (dp <= 21) { h &= ~0ull << (63 - 3 * dp); l = 0; }
        else if (dp < 42) { l &= ~0ull << (63 - 3 * (dp - 21)); }   // deeper nodes: all of key_lo belongs to the path
        if (depth) depth[k] = dp;
        if (count) count[k] = (int64_t)last[(size_t)k] - first[(size_t)k] + 1;
        if (duplicated) duplicated[k] = dup[(size_t)k];
        if (key_hi) key_hi[k] = h;
        if (key_lo) key_lo[k] = l;
        if (mass) mass[k] = pm[(size_t)k].w;
        if (comx) comx[k] = pm[(size_t)k].x;
        if (comy) comy[k] = pm[(size_t)k].y;
        if (comz) comz[k] = pm[(size_t)k].z;
        if (gas_mass) gas_mass[k] = gv[(size_t)k].w;
        if (mvx) mvx[k] = gv[(size_t)k].x;
        if (mvy) mvy[k] = gv[(size_t)k].y;
        if (mvz) mvz[k] = gv[(size_t)k].z;
    }
    return AGB_OK;
}

int agb_get_target_counters(agb_ctx* c, int32_t* visits, int32_t* acc_nodes, int32_t* acc_leaves, int32_t* sph)
{
    if (!c || !c->counters_valid || !visits || !acc_nodes || !acc_leaves || !sph) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const int64_t n = c->d.n;
    if (n == 0) return AGB_OK;
    DevTmp<int32_t> t;
    CK(dalloc(t.p, 4 * (size_t)n));
    int32_t* tmp = t.p;
    c->launches += agb_launch_unpermute_counters(c->d, tmp, tmp + n, tmp + 2 * n, tmp + 3 * n, c->st);
    CK(cudaMemcpyAsync(visits, tmp, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(acc_nodes, tmp + n, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(acc_leaves, tmp + 2 * n, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(sph, tmp + 3 * n, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return AGB_OK;
}

int agb_get_phase_ms(agb_ctx* c, double ms[5])
{
    if (!c || !ms) return AGB_ERR_INVALID;
    for (int i = 0; i < 5; i++) ms[i] = c->phase_ms[i];
    return AGB_OK;
}

int agb_get_kernel_ms(agb_ctx* c, double ms[8])
{
    if (!c || !ms) return AGB_ERR_INVALID;
    for (int i = 0; i < 8; i++) ms[i] = c->kernel_ms[i];
    return AGB_OK;
}

int agb_get_stream(agb_ctx* c, void** stream)
{
    if (!c || !stream) return AGB_ERR_INVALID;
    *stream = (void*)c->st;
    return AGB_OK;
}

// ---------------------------------------------------------------- device-resident driver loop
int agb_integrator_init(agb_ctx* c, double eta, double min_time_step, double max_time_step, double H0, double e0)
{
    if (!c || !c->have_particles || c->bound) return AGB_ERR_INVALID;          // needs the owned (host-uploaded) particle copies
    if (!(min_time_step > 0.0) || !(max_time_step >= min_time_step)) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    AgbDev& d = c->d;
    if (d.n > 0 && (!d.vx || !d.vy || !d.vz || !d.U || !d.next)) { c->err = "integrator needs vx, vy, vz, U and next_time arrays"; return AGB_ERR_INVALID; }
    join_uploads(c);
    AgbInt& I = c->I;
    I.x = c->in_d[0]; I.y = c->in_d[1]; I.z = c->in_d[2]; I.vx = c->in_d[3]; I.vy = c->in_d[4]; I.vz = c->in_d[5]; I.U = c->in_d[7]; I.next = c->in_d[8];
    I.mu = d.mu; I.timestep = c->timestep;
    I.eta = eta; I.e0 = e0; I.min_ts = min_time_step; I.max_ts = max_time_step;
    const double H0SI = (H0 * 1.0e3) / 3.08567758149137e22;                    // Math/Units.h, Simulation.cpp:329
    int kmin, kmax;
    frexp(min_time_step, &kmin); frexp(max_time_step, &kmax);
    I.k0 = kmin - 2;
    if (kmax - I.k0 >= AGB_INT_BINS) { c->err = "time-step range spans more than 2^126"; return AGB_ERR_UNSUPPORTED; }
    for (int j = 0; j < AGB_INT_BINS; j++) I.scale_tab[j] = exp(H0SI * ldexp(1.0, I.k0 + j));
    I.scale_min = exp(H0SI * min_time_step);
    // SFR.cpp:21-22: p = 1 - exp(-epsilon * timeStep / t_star), epsilon = 0.1, t_star = 1e15 s
    for (int j = 0; j < AGB_INT_BINS; j++) I.sf_tab[j] = 1 - exp(-0.1 * ldexp(1.0, I.k0 + j) / 1e15);
    I.sf_min = 1 - exp(-0.1 * min_time_step / 1e15);
    I.type = c->in_type;
    dfree(c->sfr);
    CK(dalloc(c->sfr, (size_t)d.n));
    CK(cudaMemsetAsync(c->sfr, 0, (size_t)std::max<int64_t>(d.n, 1) * sizeof(double), c->st));
    I.sfr = c->sfr;
    if (!c->d_min) CK(cudaMalloc((void**)&c->d_min, sizeof(unsigned long long)));
    if (d.n > 0) c->launches += agb_launch_int_init(d, I, c->st);
    c->int_time = 0.0; c->int_ready = true;
    CK(cudaGetLastError());
    return AGB_OK;
}

int agb_integrator_assign_all(agb_ctx* c)
{
    if (!c || !c->int_ready) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    if (c->d.n > 0) c->launches += agb_launch_int_assign(c->d, c->I, c->int_time, true, c->st);     // Simulation.cpp:189-208
    CK(cudaGetLastError());
    return AGB_OK;
}

int agb_step_begin(agb_ctx* c, double* global_time)
{
    if (!c || !c->int_ready || !global_time) return AGB_ERR_INVALID;
    Phase ph("first kick");
    CK(cudaSetDevice(c->device));
    AgbDev& d = c->d;
    if (d.n == 0) { *global_time = c->int_time; c->built = false; c->dens_done = false; c->forces_done = false; c->counters_valid = false; return AGB_OK; }
    c->launches += agb_launch_int_assign(d, c->I, c->int_time, false, c->st);         // Simulation.cpp:213-234
    c->launches += agb_launch_int_min(d, c->I, c->d_min, c->st);                      // :237-254
    unsigned long long bits = 0;
    CK(cudaMemcpyAsync(&bits, c->d_min, sizeof(bits), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    double gt; memcpy(&gt, &bits, 8);
    c->int_time = gt; *global_time = gt;
    c->launches += agb_launch_int_first(d, c->I, gt, c->st);                          // :258-272
    c->built = false; c->dens_done = false; c->forces_done = false; c->counters_valid = false;
    c->have_particles = true;
    CK(cudaGetLastError());
    return AGB_OK;
}

int agb_step_end(agb_ctx* c)
{
    if (!c || !c->int_ready || !c->forces_done) return AGB_ERR_INVALID;
    Phase ph("second kick");
    CK(cudaSetDevice(c->device));
    c->I.cooling = c->opt_cooling ? 1 : 0; c->I.star_formation = c->opt_sf_seed != 0 ? 1 : 0; c->I.seed = c->opt_sf_seed;
    if (c->d.n > 0) c->launches += agb_launch_int_second(c->d, c->I, c->int_time, c->st);             // :296-341
    CK(cudaGetLastError());
    return AGB_OK;
}

int agb_get_state(agb_ctx* c, double* x, double* y, double* z, double* vx, double* vy, double* vz, double* U, double* next_time, double* time_step)
{
    if (!c || !c->have_particles || c->bound) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const size_t b = (size_t)c->d.n * sizeof(double);
    struct { double* dst; const double* src; } cp[] = {{x, c->in_d[0]}, {y, c->in_d[1]}, {z, c->in_d[2]}, {vx, c->in_d[3]}, {vy, c->in_d[4]}, {vz, c->in_d[5]},
                                                       {U, c->in_d[7]}, {next_time, c->in_d[8]}, {time_step, c->timestep}};
    for (auto& e : cp) if (e.dst && b) CK(cudaMemcpyAsync(e.dst, e.src, b, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return AGB_OK;
}

int agb_get_subgrid_state(agb_ctx* c, uint8_t* type, double* sfr)
{
    if (!c || !c->int_ready) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    if (type && c->d.n) CK(cudaMemcpyAsync(type, c->in_type, (size_t)c->d.n, cudaMemcpyDeviceToHost, c->st));
    if (sfr && c->d.n) CK(cudaMemcpyAsync(sfr, c->sfr, (size_t)c->d.n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return AGB_OK;
}

int agb_microbench(agb_ctx* c, int kind, double* result)
{
    if (!c || !result || kind < 0 || kind > 2) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st));
    *result = 0;
    agb_launch_microbench(kind, c->sm_count, c->st, result);
    CK(cudaGetLastError());
    return AGB_OK;
}

int agb_get_launch_count(agb_ctx* c, int64_t* launches)
{
    if (!c || !launches) return AGB_ERR_INVALID;
    *launches = c->launches;
    return AGB_OK;
}

} // extern "C"

void agb_ctx_internals(agb_ctx* c, AgbDev** d, cudaStream_t* st, int* device)
{
    if (d) *d = &c->d;
    if (st) *st = c->st;
    if (device) *device = c->device;
}

// The hand-over `src` (another device) has just received from the host, repeated on `c` by peer-to-peer copies in the same
// three groups and on the same two streams as agb_set_particles, each group as soon as the source has it: several GPUs behind
// one handle read the host arrays once instead of once per device (agb_multi.cu).
int agb_ctx_copy_particles_from(agb_ctx* c, agb_ctx* src)
{
    if (!c || !src || !src->have_particles || src->bound) return AGB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const int64_t n = src->d.n;
    int rc = ensure_pool(c, n);
    if (rc) return rc;
    AgbDev& d = c->d;
    const AgbDev& sd = src->d;
    d.n = n;
    CK(cudaEventRecord(c->ev_sync, c->st));
    CK(cudaStreamWaitEvent(c->st_copy, c->ev_sync, 0));
    auto peer = [&](void* dst, const void* from, size_t bytes, cudaStream_t st) -> cudaError_t {
        return bytes ? cudaMemcpyPeerAsync(dst, c->device, from, src->device, bytes, st) : cudaSuccess;
    };
    const size_t b = (size_t)n * sizeof(double);
    auto input = [&](const double*& slot, int which, const double* from, cudaStream_t st) -> cudaError_t {
        if (!from) { slot = nullptr; return cudaSuccess; }
        slot = c->in_d[which];
        return peer(c->in_d[which], from, b, st);
    };
    // group 0: positions, masses, types (the build starts on them)
    CK(cudaStreamWaitEvent(c->st, src->ev_pos, 0));
    CK(input(d.x, 0, sd.x, c->st)); CK(input(d.y, 1, sd.y, c->st)); CK(input(d.z, 2, sd.z, c->st));
    CK(peer(c->in_type, src->in_type, (size_t)n, c->st));
    CK(cudaStreamWaitEvent(c->st, src->ev_mass, 0));           // the source receives its masses right behind the positions
    CK(input(d.mass, 6, sd.mass, c->st));
    d.type = c->in_type; c->bound = false;
    CK(cudaEventRecord(c->ev_pos, c->st));
    // group 1: next_time and the carried acc / dUdt / h / rho
    CK(cudaStreamWaitEvent(c->st_copy, src->ev_next, 0));
    CK(input(d.next, 8, sd.next, c->st_copy));
    CK(peer(d.ax, sd.ax, b, c->st_copy)); CK(peer(d.ay, sd.ay, b, c->st_copy)); CK(peer(d.az, sd.az, b, c->st_copy));
    CK(peer(d.dUdt, sd.dUdt, b, c->st_copy)); CK(peer(d.h, sd.h, b, c->st_copy)); CK(peer(d.rho, sd.rho, b, c->st_copy));
    CK(cudaMemsetAsync(d.vis, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(double), c->st_copy));
    CK(cudaEventRecord(c->ev_next, c->st_copy));
    // group 2: what only the SPH pair pass needs
    CK(cudaStreamWaitEvent(c->st_copy, src->ev_in, 0));
    CK(input(d.vx, 3, sd.vx, c->st_copy)); CK(input(d.vy, 4, sd.vy, c->st_copy)); CK(input(d.vz, 5, sd.vz, c->st_copy));
    CK(input(d.U, 7, sd.U, c->st_copy)); CK(input(d.mu, 9, sd.mu, c->st_copy));
    CK(peer(d.P, sd.P, b, c->st_copy)); CK(peer(d.T, sd.T, b, c->st_copy));
    CK(cudaEventRecord(c->ev_in, c->st_copy));
    c->in_pending = true; c->next_pending = true;
    c->have_particles = true; c->built = false; c->dens_done = false; c->forces_done = false; c->counters_valid = false;
    c->int_ready = false;
    for (bool& f : c->bres_sent) f = false;
    return AGB_OK;
}

void agb_ctx_join_uploads(agb_ctx* c)
{
    join_uploads(c);
}

// agb_multi.cu — several GPUs of one box behind one handle (the agb_multi_* entry points of include/agb200.h).
//
// The reference's data-parallel axis is the loop over force targets (Tree::calculateForces, Physics/Tree/Tree.cpp:65:
// `#pragma omp parallel for` over the particles against one shared tree).  Here every GPU receives the whole particle set,
// builds the same tree and computes the same densities (replicated: shipping a tree over NVLink costs more than building
// it from HBM), walks its own slice of the tree-ordered targets, and the slices' (index, acc, dU/dt) are exchanged over
// peer-to-peer copies so that every GPU ends up with the complete caller-order result arrays.  Because the 32-target
// groups and their summation order do not depend on the number of slices, the results are bit-identical to one GPU's.
//
// One host thread per device drives its agb_ctx (the C ABI's rule: one host thread per context); the calls of this file
// fan out, join, and return the first error.  No NCCL: one process owns all devices, cudaMemcpyPeerAsync is the collective.
#include "agb_internal.cuh"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

struct agb_multi {
    std::vector<agb_ctx*> ctx;
    std::vector<int> dev;
    int64_t n = 0;
    bool forces_done = false;
    bool serial = false;              // the same device listed twice (tests on a one-GPU box): its contexts take turns (their level-synchronous kernels must not share the SMs)
    // per device: compact results of its own slice (send) and room for one incoming slice (recv)
    std::vector<uint32_t*> s_idx, r_idx;
    std::vector<double*> s_val, r_val;
    std::vector<int64_t> s_cnt;
    int64_t buf_cap = 0;
    // pinned staging for array-of-structs hand-overs
    double* stage = nullptr; size_t stage_bytes = 0;
    std::string err;
};

namespace {

__global__ void k_apply_slice(const uint32_t* __restrict__ index, const double* __restrict__ val, int64_t cnt, int64_t stride,
                              double* __restrict__ ax, double* __restrict__ ay, double* __restrict__ az, double* __restrict__ dUdt)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    const uint32_t p = index[k];
    ax[p] = val[k]; ay[p] = val[stride + k]; az[p] = val[2 * stride + k]; dUdt[p] = val[3 * stride + k];
}

// run f(i) for every device on its own host thread; first non-OK status wins
int fan_out(agb_multi* m, const std::function<int(int)>& f)
{
    const int nd = (int)m->ctx.size();
    std::vector<int> rc((size_t)nd, AGB_OK);
    if (nd == 1 || m->serial) { for (int i = 0; i < nd; i++) rc[(size_t)i] = f(i); }
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nd; i++) th.emplace_back([&, i] { rc[(size_t)i] = f(i); });
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < nd; i++)
        if (rc[(size_t)i] != AGB_OK) { m->err = std::string("device ") + std::to_string(m->dev[(size_t)i]) + ": " + agb_last_error(m->ctx[(size_t)i]); return rc[(size_t)i]; }
    return AGB_OK;
}

void free_buffers(agb_multi* m)
{
    for (size_t i = 0; i < m->ctx.size(); i++) {
        cudaSetDevice(m->dev[i]);
        if (i < m->s_idx.size()) { cudaFree(m->s_idx[i]); cudaFree(m->r_idx[i]); cudaFree(m->s_val[i]); cudaFree(m->r_val[i]); }
    }
    m->s_idx.clear(); m->r_idx.clear(); m->s_val.clear(); m->r_val.clear();
    m->buf_cap = 0;
}

int ensure_buffers(agb_multi* m, int64_t n)
{
    const int nd = (int)m->ctx.size();
    const int64_t want = n / nd + 512 + 256;                       // slices end on multiples of 256 targets
    if (want <= m->buf_cap) return AGB_OK;
    free_buffers(m);
    m->s_idx.assign((size_t)nd, nullptr); m->r_idx.assign((size_t)nd, nullptr); m->s_val.assign((size_t)nd, nullptr); m->r_val.assign((size_t)nd, nullptr);
    for (int i = 0; i < nd; i++) {
        cudaSetDevice(m->dev[(size_t)i]);
        if (cudaMalloc((void**)&m->s_idx[(size_t)i], (size_t)want * 4) != cudaSuccess || cudaMalloc((void**)&m->r_idx[(size_t)i], (size_t)want * 4) != cudaSuccess ||
            cudaMalloc((void**)&m->s_val[(size_t)i], (size_t)want * 32) != cudaSuccess || cudaMalloc((void**)&m->r_val[(size_t)i], (size_t)want * 32) != cudaSuccess) {
            (void)cudaGetLastError(); m->err = "out of device memory (slice exchange buffers)"; return AGB_ERR_NOMEM;
        }
    }
    m->buf_cap = want;
    return AGB_OK;
}

// after the sliced walks: every device compacts its slice, then pulls the other devices' slices and applies them
int exchange_results(agb_multi* m)
{
    const int nd = (int)m->ctx.size();
    if (nd == 1 || m->n == 0) return AGB_OK;
    m->s_cnt.assign((size_t)nd, 0);
    int rc = fan_out(m, [&](int i) -> int {
        agb_ctx* c = m->ctx[(size_t)i];
        int64_t cnt = 0;
        int r = agb_get_slice_count(c, i, nd, &cnt);
        if (r) return r;
        if (cnt > m->buf_cap) return AGB_ERR_NOMEM;
        m->s_cnt[(size_t)i] = cnt;
        double* v = m->s_val[(size_t)i];
        // four columns at a fixed stride so that one peer copy moves them all
        return agb_get_slice_results(c, i, nd, m->s_idx[(size_t)i], v, v + m->buf_cap, v + 2 * m->buf_cap, v + 3 * m->buf_cap, AGB_MEM_DEVICE);   // synchronises the stream
    });
    if (rc) return rc;
    return fan_out(m, [&](int j) -> int {
        agb_ctx* c = m->ctx[(size_t)j];
        AgbDev* d = nullptr; cudaStream_t st = nullptr; int dev = 0;
        agb_ctx_internals(c, &d, &st, &dev);
        if (cudaSetDevice(dev) != cudaSuccess) return AGB_ERR_CUDA;
        for (int k = 1; k < nd; k++) {
            const int i = (j + k) % nd;                              // staggered: no two devices pull from the same peer at once
            const int64_t cnt = m->s_cnt[(size_t)i];
            if (cnt == 0) continue;
            if (cudaMemcpyPeerAsync(m->r_idx[(size_t)j], dev, m->s_idx[(size_t)i], m->dev[(size_t)i], (size_t)cnt * 4, st) != cudaSuccess ||
                cudaMemcpyPeerAsync(m->r_val[(size_t)j], dev, m->s_val[(size_t)i], m->dev[(size_t)i], (size_t)(3 * m->buf_cap + cnt) * 8, st) != cudaSuccess) {
                (void)cudaGetLastError();
                return AGB_ERR_CUDA;
            }
            k_apply_slice<<<(int)((cnt + 255) / 256), 256, 0, st>>>(m->r_idx[(size_t)j], m->r_val[(size_t)j], cnt, m->buf_cap, d->ax, d->ay, d->az, d->dUdt);
        }
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return AGB_ERR_CUDA;
        return AGB_OK;
    });
}

inline double* aos_d(void* base, int64_t off) { return reinterpret_cast<double*>(static_cast<char*>(base) + off); }

// host loops over the caller's records, split over a few threads (the records are 264 bytes apart: latency bound)
void host_parallel(int64_t n, const std::function<void(int64_t, int64_t)>& f)
{
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), (int64_t)16, n / 65536}));
    if (nt <= 1) { f(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t] { f(n * t / nt, n * (t + 1) / nt); });
    for (auto& t : th) t.join();
}

} // namespace

extern "C" {

int agb_multi_create(agb_multi** out, const int* devices, int ndev, int compat_cores)
{
    if (!out || ndev < 1 || ndev > 64) return AGB_ERR_INVALID;
    *out = nullptr;
    agb_multi* m = new agb_multi();
    for (int i = 0; i < ndev; i++) {
        agb_ctx* c = nullptr;
        const int dv = devices ? devices[i] : i;
        int rc = agb_create(&c, dv, compat_cores);
        if (rc) { for (auto* q : m->ctx) agb_destroy(q); delete m; return rc; }
        for (int q : m->dev) if (q == dv) m->serial = true;
        m->ctx.push_back(c); m->dev.push_back(dv);
    }
    // peer access for the slice exchange (an error here only means the copies are staged by the driver)
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < ndev; j++) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->dev[(size_t)i], m->dev[(size_t)j]);
            if (can) { cudaSetDevice(m->dev[(size_t)i]); cudaDeviceEnablePeerAccess(m->dev[(size_t)j], 0); (void)cudaGetLastError(); }
        }
    *out = m;
    return AGB_OK;
}

int agb_multi_destroy(agb_multi* m)
{
    if (!m) return AGB_OK;
    free_buffers(m);
    if (m->stage) cudaFreeHost(m->stage);
    for (auto* c : m->ctx) agb_destroy(c);
    delete m;
    return AGB_OK;
}

int agb_multi_device_count(agb_multi* m) { return m ? (int)m->ctx.size() : 0; }

int agb_multi_context(agb_multi* m, int i, agb_ctx** ctx)
{
    if (!m || !ctx || i < 0 || i >= (int)m->ctx.size()) return AGB_ERR_INVALID;
    *ctx = m->ctx[(size_t)i];
    return AGB_OK;
}

const char* agb_multi_last_error(agb_multi* m) { return m ? m->err.c_str() : ""; }

int agb_multi_set_option(agb_multi* m, int option, int64_t value)
{
    if (!m) return AGB_ERR_INVALID;
    if (option == AGB_OPT_SLICE_DENSITIES && value != 0) return AGB_ERR_INVALID;   // this handle hands back whole arrays: every device needs all densities
    for (auto* c : m->ctx) { int rc = agb_set_option(c, option, value); if (rc) return rc; }
    return AGB_OK;
}

int agb_multi_set_particles(agb_multi* m, const agb_particles* p)
{
    if (!m || !p) return AGB_ERR_INVALID;
    int rc = ensure_buffers(m, p->n);
    if (rc) return rc;
    m->n = p->n; m->forces_done = false;
    // The host arrays are read once: device 0 uploads them, the others repeat its hand-over with peer-to-peer copies over
    // NVLink, group by group as the data arrives (caller memory is usually pageable: N uploads of it would share the host's
    // staging bandwidth).  Without peer access, or on a one-GPU test box, every context uploads for itself.
    bool p2p = !m->serial;
    for (size_t i = 1; i < m->ctx.size() && p2p; i++) { int can = 0; cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[0]); p2p = can != 0; }
    if (!p2p || m->ctx.size() == 1) return fan_out(m, [&](int i) { return agb_set_particles(m->ctx[(size_t)i], p, AGB_MEM_HOST); });
    rc = agb_set_particles(m->ctx[0], p, AGB_MEM_HOST);
    if (rc) { m->err = agb_last_error(m->ctx[0]); return rc; }
    return fan_out(m, [&](int i) { return i == 0 ? (int)AGB_OK : agb_ctx_copy_particles_from(m->ctx[(size_t)i], m->ctx[0]); });
}

int agb_multi_set_particles_aos(agb_multi* m, void* const* parts, int64_t n, const agb_aos_layout* L)
{
    if (!m || !parts || !L || n < 0 || n >= (1ll << 30) || L->position < 0 || L->mass < 0 || L->type < 0) return AGB_ERR_INVALID;
    for (size_t i = 0; i < m->ctx.size(); i++) {                    // the staging buffer may still be the source of the previous hand-over
        AgbDev* d; cudaStream_t st; int dev;
        agb_ctx_internals(m->ctx[i], &d, &st, &dev);
        cudaSetDevice(dev); cudaDeviceSynchronize();
    }
    const size_t nn = (size_t)std::max<int64_t>(n, 1), cols = 18, need = cols * nn * sizeof(double) + nn;
    if (need > m->stage_bytes) {
        if (m->stage) cudaFreeHost(m->stage);
        m->stage = nullptr; m->stage_bytes = 0;
        if (cudaHostAlloc((void**)&m->stage, need, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); m->err = "out of pinned host memory"; return AGB_ERR_NOMEM; }
        m->stage_bytes = need;
    }
    double* col[18];
    for (size_t k = 0; k < cols; k++) col[k] = m->stage + k * nn;
    uint8_t* typ = reinterpret_cast<uint8_t*>(m->stage + cols * nn);
    std::atomic<bool> bad{false};
    host_parallel(n, [&](int64_t a, int64_t b) {
        auto vec = [&](int64_t off, int64_t i, int k0) {
            if (off < 0) { col[k0][i] = col[k0 + 1][i] = col[k0 + 2][i] = 0.0; return; }
            const double* v = aos_d(parts[i], off); col[k0][i] = v[0]; col[k0 + 1][i] = v[1]; col[k0 + 2][i] = v[2];
        };
        auto sca = [&](int64_t off, int64_t i, int k, double dflt) { col[k][i] = off < 0 ? dflt : *aos_d(parts[i], off); };
        for (int64_t i = a; i < b; i++) {
            if (!parts[i]) { bad = true; return; }
            vec(L->position, i, 0); vec(L->velocity, i, 3); vec(L->acc, i, 6);
            sca(L->mass, i, 9, 0); sca(L->U, i, 10, 0); sca(L->next_time, i, 11, 0); sca(L->mu, i, 12, 0.58);
            sca(L->rho, i, 13, 0); sca(L->P, i, 14, 0); sca(L->T, i, 15, 0); sca(L->h, i, 16, 0); sca(L->dUdt, i, 17, 0);
            typ[i] = *reinterpret_cast<const uint8_t*>(static_cast<const char*>(parts[i]) + L->type);
        }
    });
    if (bad) return AGB_ERR_INVALID;
    agb_particles p;
    memset(&p, 0, sizeof(p));
    p.n = n;
    p.x = col[0]; p.y = col[1]; p.z = col[2]; p.vx = col[3]; p.vy = col[4]; p.vz = col[5]; p.ax = col[6]; p.ay = col[7]; p.az = col[8];
    p.mass = col[9]; p.U = col[10]; p.next_time = col[11]; p.mu = col[12]; p.rho = col[13]; p.P = col[14]; p.T = col[15]; p.h = col[16]; p.dUdt = col[17];
    p.type = typ;
    return agb_multi_set_particles(m, &p);
}

int agb_multi_force_path(agb_multi* m, double visual_density_radius, double mass_in_h, double global_time, double e0, double theta, double* root_radius)
{
    if (!m) return AGB_ERR_INVALID;
    const int nd = (int)m->ctx.size();
    std::vector<double> R((size_t)nd, 0.0);
    int rc = fan_out(m, [&](int i) { return agb_force_path(m->ctx[(size_t)i], visual_density_radius, mass_in_h, global_time, e0, theta, i, nd, &R[(size_t)i]); });
    if (rc) return rc;
    if (root_radius) *root_radius = R[0];
    if ((rc = exchange_results(m))) return rc;
    m->forces_done = true;
    return AGB_OK;
}

// The four separate calls (a driver that needs root->radius before the densities, Simulation.cpp:121-139)
int agb_multi_build_tree(agb_multi* m, double* root_radius)
{
    if (!m) return AGB_ERR_INVALID;
    std::vector<double> R(m->ctx.size(), 0.0);
    int rc = fan_out(m, [&](int i) { return agb_build_tree(m->ctx[(size_t)i], &R[(size_t)i]); });
    if (!rc && root_radius) *root_radius = R[0];
    return rc;
}
int agb_multi_visual_density(agb_multi* m, double radius) { return m ? fan_out(m, [&](int i) { return agb_visual_density(m->ctx[(size_t)i], radius); }) : AGB_ERR_INVALID; }
int agb_multi_gas_density(agb_multi* m, double mass_in_h) { return m ? fan_out(m, [&](int i) { return agb_gas_density(m->ctx[(size_t)i], mass_in_h); }) : AGB_ERR_INVALID; }
int agb_multi_forces(agb_multi* m, double global_time, double e0, double theta)
{
    if (!m) return AGB_ERR_INVALID;
    const int nd = (int)m->ctx.size();
    int rc = fan_out(m, [&](int i) { return agb_forces_slice(m->ctx[(size_t)i], global_time, e0, theta, i, nd); });
    if (rc) return rc;
    if ((rc = exchange_results(m))) return rc;
    m->forces_done = true;
    return AGB_OK;
}

int agb_multi_get_results(agb_multi* m, const agb_results* r)
{
    if (!m || !r) return AGB_ERR_INVALID;
    const int nd = (int)m->ctx.size();
    const int64_t n = m->n;
    // every device holds the complete arrays: each sends its share of the rows over its own link
    return fan_out(m, [&](int i) -> int {
        const int64_t a = n * i / nd, b = n * (i + 1) / nd;
        AgbDev* d; cudaStream_t st; int dev;
        agb_ctx_internals(m->ctx[(size_t)i], &d, &st, &dev);
        if (cudaSetDevice(dev) != cudaSuccess) return AGB_ERR_CUDA;
        double* dst[9] = {r->ax, r->ay, r->az, r->dUdt, r->h, r->rho, r->P, r->T, r->visualDensity};
        const double* src[9] = {d->ax, d->ay, d->az, d->dUdt, d->h, d->rho, d->P, d->T, d->vis};
        agb_ctx_join_uploads(m->ctx[(size_t)i]);
        for (int k = 0; k < 9; k++)
            if (dst[k] && b > a) cudaMemcpyAsync(dst[k] + a, src[k] + a, (size_t)(b - a) * 8, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return AGB_ERR_CUDA;
        return AGB_OK;
    });
}

int agb_multi_get_results_aos(agb_multi* m, void* const* parts, int64_t n, const agb_aos_layout* L)
{
    if (!m || !parts || !L || n != m->n) return AGB_ERR_INVALID;
    const size_t nn = (size_t)std::max<int64_t>(n, 1), need = 9 * nn * sizeof(double);
    if (need > m->stage_bytes) {
        if (m->stage) cudaFreeHost(m->stage);
        m->stage = nullptr; m->stage_bytes = 0;
        if (cudaHostAlloc((void**)&m->stage, need, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return AGB_ERR_NOMEM; }
        m->stage_bytes = need;
    }
    double* col[9];
    for (int k = 0; k < 9; k++) col[k] = m->stage + (size_t)k * nn;
    agb_results r = {col[0], col[1], col[2], col[3], col[4], col[5], col[6], col[7], col[8]};
    int rc = agb_multi_get_results(m, &r);
    if (rc) return rc;
    host_parallel(n, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i++) {
            if (L->acc >= 0) { double* v = aos_d(parts[i], L->acc); v[0] = col[0][i]; v[1] = col[1][i]; v[2] = col[2][i]; }
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

// ---- device-resident loop: the integrator runs replicated (every device advances every particle: a few streaming kernels),
// so the only per-step exchange is the slices' (index, acc, dU/dt)
int agb_multi_integrator_init(agb_multi* m, double eta, double min_ts, double max_ts, double H0, double e0)
{ return m ? fan_out(m, [&](int i) { return agb_integrator_init(m->ctx[(size_t)i], eta, min_ts, max_ts, H0, e0); }) : AGB_ERR_INVALID; }
int agb_multi_integrator_assign_all(agb_multi* m) { return m ? fan_out(m, [&](int i) { return agb_integrator_assign_all(m->ctx[(size_t)i]); }) : AGB_ERR_INVALID; }
int agb_multi_step_begin(agb_multi* m, double* global_time)
{
    if (!m || !global_time) return AGB_ERR_INVALID;
    std::vector<double> t(m->ctx.size(), 0.0);
    int rc = fan_out(m, [&](int i) { return agb_step_begin(m->ctx[(size_t)i], &t[(size_t)i]); });
    if (rc) return rc;
    for (double v : t) if (v != t[0]) { m->err = "devices disagree on the next integration time"; return AGB_ERR_CUDA; }
    *global_time = t[0];
    return AGB_OK;
}
int agb_multi_step_end(agb_multi* m) { return m ? fan_out(m, [&](int i) { return agb_step_end(m->ctx[(size_t)i]); }) : AGB_ERR_INVALID; }
int agb_multi_get_state(agb_multi* m, double* x, double* y, double* z, double* vx, double* vy, double* vz, double* U, double* next_time, double* time_step)
{ return m ? agb_get_state(m->ctx[0], x, y, z, vx, vy, vz, U, next_time, time_step) : AGB_ERR_INVALID; }
int agb_multi_get_subgrid_state(agb_multi* m, uint8_t* type, double* sfr) { return m ? agb_get_subgrid_state(m->ctx[0], type, sfr) : AGB_ERR_INVALID; }

} // extern "C"

// agb_extended.cu — the extended-accuracy mode (AGB_OPT_EXTENDED): what BASELINE.json's north_star prose
// describes and the reference does NOT do.  On the same tree (agb_build.cu):
//   * gravity with monopole + traceless quadrupole moments, Newtonian with the cubic-spline softening of Springel, Yoshida &
//     White (2001) (the W2 kernel the reference carries as dead code, Math/kernel.cpp:41-56), one interaction list per group
//     of 32 tree-adjacent targets, opening criterion (cell width) / d_min < theta with d_min the distance from the group's
//     bounding box to the node's centre of mass (the textbook meaning of theta; the reference tests the HALF width per target);
//   * classic SPH: a smoothing length per gas particle from (4 pi / 3) (2 h)^3 rho(h) = massInH, rho = sum_j m_j W(r_ij, h_i)
//     over the neighbours within 2 h_i (tree range search), pressure + Monaghan–Gingold viscosity forces and dU/dt over the
//     same neighbours.
// There is no reference for this mode: parity is UNPINNED.  It is validated against direct summation (gravity) and a numpy /
// cKDTree restatement (SPH) in tests/test_gpu_extended.py and reported by bench.py --extended on separate lines.
// FP64 throughout.  No tensor cores (not a dense contraction).
#include "agb_internal.cuh"
#include <algorithm>

namespace {

constexpr int TPB = 256;
constexpr double kG = 6.67430e-11;         // Math/Constants.h:7
constexpr double kPI = 3.14159265358979323846;
constexpr double kGAMMA = 5.0 / 3.0, kKB = 1.38064852e-23, kPRTN = 1.6726219e-27;

struct Links { int4 a, b; };
__device__ __forceinline__ Links load_links(const int32_t* child, int k)
{
    Links L;
    L.a = reinterpret_cast<const int4*>(child)[2 * (size_t)k]; L.b = reinterpret_cast<const int4*>(child)[2 * (size_t)k + 1];
    return L;
}

__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int nblocks, unsigned int& epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += nblocks;
        __threadfence();
        atomicAdd(bar, 1u);
        while (*(volatile unsigned int*)bar < epoch) { }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void ldcg6(const double* p, double q[6])
{
    const double2 a = __ldcg(reinterpret_cast<const double2*>(p)), b = __ldcg(reinterpret_cast<const double2*>(p) + 1), c = __ldcg(reinterpret_cast<const double2*>(p) + 2);
    q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y; q[4] = c.x; q[5] = c.y;
}

// ------------------------------------------------------------------ quadrupole moments, level by level (deepest first)
// Q = sum_children [ Q_c + m_c (3 s s^T - |s|^2 I) ], s = COM_c - COM, stored as (xx, xy, xz, yy, yz, zz); a particle has Q_c = 0.
__global__ void __launch_bounds__(TPB) k_quad_levels(AgbDev d, const AgbScalars* __restrict__ s, const int32_t* __restrict__ list, double* __restrict__ quad, unsigned int* bar)
{
    const int nn = s->n_nodes;
    if (s->node_overflow || nn < 1) return;
    const int N = (int)d.n;
    unsigned int epoch = 0;
    int off_end = nn;
    int maxd = AGB_MAX_LEVELS;
    while (maxd > 0 && s->lvl_cnt[maxd] == 0) maxd--;
    for (int q = maxd + 1; q <= AGB_MAX_LEVELS; q++) off_end -= s->lvl_cnt[q];
    for (int dep = maxd; dep >= 0; dep--) {
        const int cnt = s->lvl_cnt[dep];
        const int beg = off_end - cnt;
        for (int idx = beg + blockIdx.x * TPB + threadIdx.x; idx < off_end; idx += gridDim.x * TPB) {
            const int k = list[idx];
            const Links L = load_links(d.child, k);
            const int ch[8] = {L.a.x, L.a.y, L.a.z, L.a.w, L.b.x, L.b.y, L.b.z, L.b.w};
            const double4 com = d.src_pm[N + k];
            double Q[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int o = 0; o < 8; o++) {
                const int c = ch[o];
                if (c < 0) continue;
                const double4 pm = d.src_pm[c];                         // particle, or child node (COM, mass): final since the build
                if (c >= N) { double qc[6]; ldcg6(quad + 6 * (size_t)(c - N), qc); for (int t = 0; t < 6; t++) Q[t] += qc[t]; }
                const double sx = pm.x - com.x, sy = pm.y - com.y, sz = pm.z - com.z, s2 = sx * sx + sy * sy + sz * sz, m = pm.w;
                Q[0] += m * (3.0 * sx * sx - s2); Q[1] += m * 3.0 * sx * sy; Q[2] += m * 3.0 * sx * sz;
                Q[3] += m * (3.0 * sy * sy - s2); Q[4] += m * 3.0 * sy * sz; Q[5] += m * (3.0 * sz * sz - s2);
            }
            double2* out = reinterpret_cast<double2*>(quad + 6 * (size_t)k);
            out[0] = make_double2(Q[0], Q[1]); out[1] = make_double2(Q[2], Q[3]); out[2] = make_double2(Q[4], Q[5]);
        }
        off_end = beg;
        if (cnt > 0 || dep == 0) grid_sync(bar, gridDim.x, epoch);
    }
}

// ------------------------------------------------------------------ gravity: one interaction list per group of 32 targets
constexpr int XWARPS = 4, XL = 1024, XS = 2048;      // 14.5 KB of shared memory per warp, 3 CTAs per SM
struct ExtWarp {
    int list[XL];
    int stack[XS];
    double4 spm[32];
    double sq[32][6];
};

__device__ __forceinline__ double warp_min(double v) { for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ double warp_max(double v) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    return v;
}

struct Slice { int64_t a0, a1; bool ident; };
__device__ __forceinline__ Slice target_slice(const AgbScalars* s, int64_t N, int part, int nparts)
{   // the same slices as agb_walk.cu: multiples of 256 active targets in tree order
    const int64_t na = s->n_active, nsg = (na + 255) / 256;
    Slice sl;
    sl.a0 = min(na, nsg * part / nparts * 256);
    sl.a1 = min(na, nsg * (part + 1) / nparts * 256);
    sl.ident = na == N;
    return sl;
}

// 1/sqrt(x) for positive normal x: MUFU.RSQ64H seed (~2^-22) + one cubically convergent step (~2^-60), as in agb_walk.cu
__device__ __forceinline__ double rsqrt_pos(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}

// spline-softened Newtonian attraction per unit mass and unit G: returns f with a = -f d  (Springel et al. 2001, eq. A1; Gadget-2)
__device__ __forceinline__ double soft_fac(double r2, double ri, double hs, double hs_inv3)
{
    if (r2 >= hs * hs) return ri * ri * ri;
    const double u = r2 * ri / hs;
    if (u < 0.5) return hs_inv3 * (10.666666666666666 + u * u * (32.0 * u - 38.4));
    return hs_inv3 * (21.333333333333332 - 48.0 * u + 38.4 * u * u - 10.666666666666666 * u * u * u - 0.06666666666666667 / (u * u * u));
}

__global__ void __launch_bounds__(XWARPS * 32, 3) k_walk_ext(AgbDev d, AgbScalars* s, const double* __restrict__ quad, double theta, double eps, int part, int nparts, int use_quad)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ExtWarp& sm = reinterpret_cast<ExtWarp*>(smem_raw)[warp];
    const unsigned lt = (1u << lane) - 1u;
    const int N = (int)d.n;
    if (s->node_overflow) return;
    const double R = __longlong_as_double((long long)s->Rbits);
    const int n_nodes = s->n_nodes, n_in_tree = s->n_in_tree;
    const double theta2 = theta * theta;
    const double hs = 2.8 * eps, hs_inv3 = hs > 0.0 ? 1.0 / (hs * hs * hs) : 0.0;
    const Slice sl = target_slice(s, d.n, part, nparts);
    const unsigned ngroups = (unsigned)((sl.a1 - sl.a0 + 31) / 32);
    unsigned long long tot_node = 0, tot_leaf = 0;
    for (;;) {
        unsigned g = 0;
        if (lane == 0) g = atomicAdd(&s->walk_next_group, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= ngroups) break;
        const int64_t idx = sl.a0 + (int64_t)g * 32 + lane;
        const bool active = idx < sl.a1;
        const int64_t t = !active ? -1 : sl.ident ? idx : (int64_t)d.act_list[idx];
        const double4 tp = active ? d.src_pm[t] : make_double4(0, 0, 0, 0);
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        const double lox = warp_min(active ? tp.x : inf), loy = warp_min(active ? tp.y : inf), loz = warp_min(active ? tp.z : inf);
        const double hix = warp_max(active ? tp.x : -inf), hiy = warp_max(active ? tp.y : -inf), hiz = warp_max(active ? tp.z : -inf);
        double ax = 0, ay = 0, az = 0;
        int sp = 0, lc = 0;
        if (n_nodes > 0) { if (lane == 0) sm.stack[0] = N; sp = 1; }
        else if (n_in_tree == 1) { if (lane == 0) sm.list[0] = 0; lc = 1; }
        __syncwarp();
        while (sp > 0 || lc > 0) {
            // ---- traversal: one node per lane against the group's box
            while (sp > 0 && lc <= XL - 32 * 9) {
                // A round pops up to 32 nodes and pushes up to 8 children each.  The rounds narrow as the stack fills so that a round
                // never leaves more than XS - 441 entries: from there even a plain depth-first descent (at most 7 entries per level,
                // 63 levels) fits.
                const int cnt = min(sp, max(1, min(32, (XS - 441 - sp) / 7)));
                sp -= cnt;
                const int node = lane < cnt ? sm.stack[sp + lane] : -1;
                __syncwarp();
                bool accept = false, open = false;
                Links L; L.a = make_int4(-1, -1, -1, -1); L.b = L.a;
                if (node >= 0) {
                    const double4 pm = d.src_pm[node];
                    if (pm.w != 0.0) {
                        const double rad = scalbn(R, -(int)d.ndepth[node - N]);
                        const double dx = fmax(0.0, fmax(lox - pm.x, pm.x - hix)), dy = fmax(0.0, fmax(loy - pm.y, pm.y - hiy)), dz = fmax(0.0, fmax(loz - pm.z, pm.z - hiz));
                        const double dmin2 = dx * dx + dy * dy + dz * dz;
                        accept = dmin2 * theta2 > 4.0 * rad * rad;           // cell width / distance < theta for every target of the group
                        open = !accept;
                        if (open) L = load_links(d.child, node - N);
                    }
                }
                const unsigned am = __ballot_sync(0xffffffffu, accept);
                if (accept) sm.list[lc + __popc(am & lt)] = node;
                lc += __popc(am);
                const unsigned om = __ballot_sync(0xffffffffu, open);
                if (om) {
                    const int ch[8] = {L.a.x, L.a.y, L.a.z, L.a.w, L.b.x, L.b.y, L.b.z, L.b.w};
                    int nl = 0, nn = 0;
                    if (open) {
#pragma unroll
                        for (int c = 0; c < 8; c++) { nl += (ch[c] >= 0 && ch[c] < N); nn += (ch[c] >= N); }
                    }
                    const int sc = warp_incl_scan(nl | (nn << 16), lane), st_ = __shfl_sync(0xffffffffu, sc, 31);
                    int pl = lc + (sc & 0xffff) - nl, pn = sp + (sc >> 16) - nn;
                    if (open) {
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            if (ch[c] >= N) { if (pn < XS) sm.stack[pn] = ch[c]; else s->walk_overflow = 1; pn++; }
                            else if (ch[c] >= 0) sm.list[pl++] = ch[c];
                        }
                    }
                    lc += st_ & 0xffff; sp = min(sp + (st_ >> 16), XS);
                }
                __syncwarp();
            }
            // ---- evaluation: full tiles of 32 sources (the rest too once the traversal is over)
            int done = 0;
            while (lc - done >= 32 || (sp == 0 && lc - done > 0)) {
                const int cnt = min(32, lc - done);
                int e = -1;
                if (lane < cnt) e = sm.list[done + lane];
                // nodes to the front of the tile, particles behind them: two branch-free loops
                const unsigned nodem = __ballot_sync(0xffffffffu, e >= N), leafm = __ballot_sync(0xffffffffu, e >= 0 && e < N);
                const int nnodes = use_quad ? __popc(nodem) : 0;
                if (lane < cnt) {
                    const bool nd = e >= N && use_quad;
                    const int pos = nd ? __popc(nodem & lt) : nnodes + __popc((use_quad ? leafm : (nodem | leafm)) & lt);
                    sm.spm[pos] = d.src_pm[e];
                    if (nd) {
                        double q[6];
                        ldcg6(quad + 6 * (size_t)(e - N), q);
#pragma unroll
                        for (int k = 0; k < 6; k++) sm.sq[pos][k] = q[k];
                    }
                }
                __syncwarp();
                if (active) {
#pragma unroll 2
                    for (int j = 0; j < nnodes; j++) {
                        const double4 q = sm.spm[j];
                        const double dx = tp.x - q.x, dy = tp.y - q.y, dz = tp.z - q.z;
                        const double r2 = dx * dx + dy * dy + dz * dz;
                        const double ri = r2 > 0.0 ? rsqrt_pos(r2) : 0.0;
                        const double f = q.w * soft_fac(r2, ri, hs, hs_inv3);
                        const double* Q = sm.sq[j];
                        const double qx = Q[0] * dx + Q[1] * dy + Q[2] * dz, qy = Q[1] * dx + Q[3] * dy + Q[4] * dz, qz = Q[2] * dx + Q[4] * dy + Q[5] * dz;
                        const double ri2 = ri * ri, ri5 = ri2 * ri2 * ri, dqd = dx * qx + dy * qy + dz * qz, c7 = 2.5 * dqd * ri5 * ri2;
                        ax += qx * ri5 - (c7 + f) * dx; ay += qy * ri5 - (c7 + f) * dy; az += qz * ri5 - (c7 + f) * dz;
                    }
#pragma unroll 4
                    for (int j = nnodes; j < cnt; j++) {
                        const double4 q = sm.spm[j];
                        const double dx = tp.x - q.x, dy = tp.y - q.y, dz = tp.z - q.z;
                        const double r2 = dx * dx + dy * dy + dz * dz;
                        // the target itself / a coincident source: d = 0 makes the term vanish as long as the factor stays finite
                        const double ri = r2 > 0.0 ? rsqrt_pos(r2) : 0.0;
                        const double f = q.w * soft_fac(r2, ri, hs, hs_inv3);
                        ax -= f * dx; ay -= f * dy; az -= f * dz;
                    }
                    const int nn_ = __popc(nodem);
                    tot_node += nn_; tot_leaf += cnt - nn_;
                }
                done += cnt;
                __syncwarp();
            }
            // keep what is left of the list at its front
            const int left = lc - done;
            int v = -1;
            if (lane < left) v = sm.list[done + lane];
            __syncwarp();
            if (lane < left) sm.list[lane] = v;
            lc = left;
            __syncwarp();
        }
        if (active) {
            const uint32_t p = d.perm[d.cur][t];
            const bool massless = tp.w == 0.0;
            d.ax[p] = massless ? 0.0 : kG * ax; d.ay[p] = massless ? 0.0 : kG * ay; d.az[p] = massless ? 0.0 : kG * az;
        }
    }
    for (int o = 16; o > 0; o >>= 1) { tot_node += __shfl_xor_sync(0xffffffffu, tot_node, o); tot_leaf += __shfl_xor_sync(0xffffffffu, tot_leaf, o); }
    if (lane == 0) { if (tot_node) atomicAdd(&s->c_node, tot_node); if (tot_leaf) atomicAdd(&s->c_leaf, tot_leaf); }
}

// ------------------------------------------------------------------ SPH with a smoothing length per particle
__device__ __forceinline__ double spline_w(double r, double h)
{   // Math/kernel.cpp:4-16
    const double a = 1.0 / (kPI * h * h * h), q = r / h;
    if (q < 1.0) return a * (1 - 1.5 * q * q + 0.75 * q * q * q);
    if (q < 2.0) { const double t = 2 - q; return a * 0.25 * (t * t * t); }
    return 0.0;
}
__device__ __forceinline__ double spline_dw(double r, double h)
{   // dW/dr, Math/kernel.cpp:18-39
    const double a = 1.0 / (kPI * h * h * h * h), q = r / h;
    if (q < 1.0) return a * (-3.0 * q + 2.25 * q * q);
    if (q < 2.0) { const double t = 2 - q; return a * (-0.75 * t * t); }
    return 0.0;
}

// squared distance from a point to the cell of node k (the depth-`dep` cell that holds the node's first particle)
__device__ __forceinline__ double cell_dist2(const AgbDev& d, double R, double invR, int k, double px, double py, double pz)
{
    const int dep = d.ndepth[k];
    const double w = scalbn(R, 1 - dep), iw = scalbn(invR, dep - 1);      // cell width 2 R / 2^dep and its inverse
    const double4 f = d.src_pm[d.nfirst[k]];
    const double cx = floor((f.x + R) * iw) * w - R, cy = floor((f.y + R) * iw) * w - R, cz = floor((f.z + R) * iw) * w - R;
    const double m = 1e-9 * w;                                             // rounding of the cell planes (and of 1 / R)
    const double dx = fmax(0.0, fmax(cx - m - px, px - (cx + w + m))), dy = fmax(0.0, fmax(cy - m - py, py - (cy + w + m))), dz = fmax(0.0, fmax(cz - m - pz, pz - (cz + w + m)));
    return dx * dx + dy * dy + dz * dz;
}

// sum_j m_j W(r_ij, h) over the gas particles within 2 h of (px, py, pz): depth-first over the cells that touch the sphere
// the smallest cell on the root path of tree position i that holds the whole ball of radius rad about (px, py, pz): searches start there
__device__ __forceinline__ int enclosing_node(const AgbDev& d, double R, double invR, int64_t i, double px, double py, double pz, double rad)
{
    int k = d.leafparent[i];
    while (k > 0) {
        const int dep = d.ndepth[k];
        const double w = scalbn(R, 1 - dep), iw = scalbn(invR, dep - 1);
        const double cx = floor((px + R) * iw) * w - R, cy = floor((py + R) * iw) * w - R, cz = floor((pz + R) * iw) * w - R, m = 1e-9 * w;
        if (px - rad > cx + m && px + rad < cx + w - m && py - rad > cy + m && py + rad < cy + w - m && pz - rad > cz + m && pz + rad < cz + w - m) break;
        k = d.nparent[k];
    }
    return max(k, 0);
}

__device__ double density_at(const AgbDev& d, const AgbScalars* s, double R, double invR, int64_t self, double px, double py, double pz, double h)
{
    const int N = (int)d.n;
    const double r2max = 4.0 * h * h;
    double rho = 0.0;
    int stack[96];
    int sp = 0;
    if (s->n_nodes > 0) stack[sp++] = enclosing_node(d, R, invR, self, px, py, pz, 2.0 * h);
    else if (s->n_in_tree == 1 && d.s_type[0] == 2) { const double4 q = d.src_pm[0]; const double r2 = (q.x - px) * (q.x - px) + (q.y - py) * (q.y - py) + (q.z - pz) * (q.z - pz); if (r2 < r2max) rho += q.w * spline_w(sqrt(r2), h); }
    while (sp > 0) {
        const int k = stack[--sp];
        if (!(d.src_gv[N + k].w > 0.0)) continue;                         // no gas below
        if (cell_dist2(d, R, invR, k, px, py, pz) >= r2max) continue;
        const Links L = load_links(d.child, k);
        const int ch[8] = {L.a.x, L.a.y, L.a.z, L.a.w, L.b.x, L.b.y, L.b.z, L.b.w};
#pragma unroll
        for (int o = 0; o < 8; o++) {
            const int c = ch[o];
            if (c < 0) continue;
            if (c >= N) { if (sp < 96) stack[sp++] = c - N; }
            else if (d.s_type[c] == 2) {
                const double4 q = d.src_pm[c];
                const double dx = q.x - px, dy = q.y - py, dz = q.z - pz, r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < r2max) rho += q.w * spline_w(sqrt(r2), h);
            }
        }
    }
    return rho;
}

// h_i from (4 pi / 3) (2 h)^3 rho(h) = massInH (monotone in h): bracket by doubling / halving, then bisection to 1e-10 relative
// one thread per GAS particle, taken from the compact tree-ordered list: the lanes of a warp search neighbouring regions
// (one thread per particle of any type left 2 of 32 lanes busy on a disk galaxy)
__global__ void __launch_bounds__(128) k_ext_density(AgbDev d, const AgbScalars* __restrict__ s, const int32_t* __restrict__ gas_list, double massInH)
{
    const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (r >= s->n_gas_total || s->node_overflow) return;
    const int64_t i = gas_list[r];
    const uint32_t p = d.perm[d.cur][i];
    if (i >= s->n_in_tree) { d.s_h[i] = 0.0; d.h[p] = 0.0; return; }       // outside the root cube: no neighbours, no SPH (like the reference's outliers)
    const double R = __longlong_as_double((long long)s->Rbits), invR = 1.0 / R;
    const double4 x = d.src_pm[i];
    auto F = [&](double h, double& rho) { rho = density_at(d, s, R, invR, i, x.x, x.y, x.z, h); return (4.0 * kPI / 3.0) * 8.0 * h * h * h * rho - massInH; };
    double rho = 0.0;
    // start at the particle's leaf cell (cheap evaluations first: the cost of one grows with h^3), or at the smoothing length
    // handed over with the particle
    double lo = scalbn(R, -(int)d.leafdepth[i]), hi = lo;
    if (d.h[p] > 0.0 && d.h[p] < 4.0 * R) lo = hi = d.h[p];
    double flo = F(lo, rho), fhi = flo;
    int guard = 0;
    if (flo > 0.0) { while (flo > 0.0 && guard++ < 400) { hi = lo; fhi = flo; lo *= 0.5; flo = F(lo, rho); } }
    else { while (fhi <= 0.0 && guard++ < 400) { lo = hi; flo = fhi; hi *= 2.0; fhi = F(hi, rho); if (hi > 4.0 * R) break; } }
    // Illinois variant of regula falsi inside the bracket (F is smooth and increasing), to 1e-9 relative
    int side = 0;
    for (int it = 0; it < 60 && hi - lo > 1e-9 * hi && fhi > 0.0 && flo <= 0.0; it++) {
        double mid = (lo * fhi - hi * flo) / (fhi - flo);
        if (!(mid > lo && mid < hi)) mid = 0.5 * (lo + hi);
        const double fm = F(mid, rho);
        if (fm > 0.0) { hi = mid; fhi = fm; if (side == 1) flo *= 0.5; side = 1; }
        else { lo = mid; flo = fm; if (side == -1) fhi *= 0.5; side = -1; }
    }
    const double h = fhi > 0.0 && flo <= 0.0 ? (lo * fhi - hi * flo) / (fhi - flo) : hi;
    rho = density_at(d, s, R, invR, i, x.x, x.y, x.z, h);
    const double U = d.s_U[i], mu = d.s_mu[i];
    const double P = (kGAMMA - 1.0) * U * rho, T = (kGAMMA - 1.0) * U * kPRTN * mu / kKB;
    d.s_h[i] = h; d.s_rho[i] = rho; d.s_P[i] = P; d.s_T[i] = T;
    d.h[p] = h; d.rho[p] = rho; d.P[p] = P; d.T[p] = T;
}

// pressure + viscosity over the neighbours within 2 h_i ("gather" form with the target's kernel):
//   a_i -= sum_j m_j (P_i/rho_i^2 + P_j/rho_j^2 + Pi_ij) grad_i W(r_ij, h_i),   dU_i/dt += 1/2 sum_j m_j (...) v_ij . grad_i W
//   Pi_ij = (-alpha c_ij mu_ij + beta mu_ij^2) / rho_ij for v_ij . r_ij < 0, mu_ij = h_ij v_ij.r_ij / (r^2 + 0.01 h_ij^2); alpha = 0.5, beta = 1
__global__ void __launch_bounds__(128) k_ext_sph_force(AgbDev d, const AgbScalars* __restrict__ s, int part, int nparts)
{
    const Slice sl = target_slice(s, d.n, part, nparts);
    const int64_t idx = sl.a0 + (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (idx >= sl.a1 || s->node_overflow) return;
    const int64_t i = sl.ident ? idx : (int64_t)d.act_list[idx];
    if (d.s_type[i] != 2 || i >= s->n_in_tree) return;
    const double hi_ = d.s_h[i], rhoi = d.s_rho[i];
    if (!(hi_ > 0.0) || !(rhoi > 0.0)) return;
    const int N = (int)d.n;
    const double R = __longlong_as_double((long long)s->Rbits), invR = 1.0 / R;
    const double4 xi = d.src_pm[i], vi = d.src_gv[i];
    const double Pi = d.s_P[i], ci = sqrt(kGAMMA * Pi / rhoi), pri = Pi / (rhoi * rhoi), r2max = 4.0 * hi_ * hi_;
    double ax = 0, ay = 0, az = 0, dU = 0;
    int stack[96];
    int sp = 0;
    if (s->n_nodes > 0) stack[sp++] = enclosing_node(d, R, invR, i, xi.x, xi.y, xi.z, 2.0 * hi_);
    while (sp > 0) {
        const int k = stack[--sp];
        if (!(d.src_gv[N + k].w > 0.0)) continue;
        if (cell_dist2(d, R, invR, k, xi.x, xi.y, xi.z) >= r2max) continue;
        const Links L = load_links(d.child, k);
        const int ch[8] = {L.a.x, L.a.y, L.a.z, L.a.w, L.b.x, L.b.y, L.b.z, L.b.w};
#pragma unroll
        for (int o = 0; o < 8; o++) {
            const int c = ch[o];
            if (c < 0) continue;
            if (c >= N) { if (sp < 96) stack[sp++] = c - N; continue; }
            if (c == (int)i || d.s_type[c] != 2) continue;
            const double4 xj = d.src_pm[c];
            const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z, r2 = dx * dx + dy * dy + dz * dz;
            if (!(r2 < r2max) || r2 == 0.0) continue;
            const double rhoj = d.s_rho[c], hj = d.s_h[c];
            if (!(rhoj > 0.0)) continue;
            const double4 vj = d.src_gv[c];
            const double Pj = d.s_P[c], r = sqrt(r2);
            const double vx = vi.x - vj.x, vy = vi.y - vj.y, vz = vi.z - vj.z, vr = vx * dx + vy * dy + vz * dz;
            double visc = 0.0;
            if (vr < 0.0) {
                const double hij = 0.5 * (hi_ + hj), cij = 0.5 * (ci + sqrt(kGAMMA * Pj / rhoj)), rhoij = 0.5 * (rhoi + rhoj);
                const double mu = hij * vr / (r2 + 0.01 * hij * hij);
                visc = (-0.5 * cij * mu + mu * mu) / rhoij;
            }
            const double term = xj.w * (pri + Pj / (rhoj * rhoj) + visc) * spline_dw(r, hi_) / r;    // times (dx, dy, dz) = grad_i W
            ax -= term * dx; ay -= term * dy; az -= term * dz;
            dU += 0.5 * term * vr;
        }
    }
    const uint32_t p = d.perm[d.cur][i];
    if (!(isnan(ax) || isnan(ay) || isnan(az))) { d.ax[p] += ax; d.ay[p] += ay; d.az[p] += az; }
    if (!isnan(dU)) d.dUdt[p] += dU;
}

} // namespace

static inline int nblk(int64_t n, int per) { return (int)((n + per - 1) / per); }

int agb_launch_extended_density(AgbDev& d, AgbScalars* s, double massInH, cudaStream_t st)
{
    int launches = agb_launch_gas_list(d, s, st);
    k_ext_density<<<nblk(d.n, 128), 128, 0, st>>>(d, s, d.nodecnt, massInH);     // (grid sized for "every particle is gas"; surplus blocks return at once)
    return launches + 1;
}

int agb_launch_extended_forces(AgbDev& d, AgbScalars* s, double globalTime, double e0, double theta, int part, int nparts, bool any_gas, bool use_quad, int sm_count, cudaStream_t st, cudaEvent_t* ev)
{
    int launches = agb_launch_active_list(d, s, globalTime, sm_count, st);
    if (d.n == 0) return launches;
    if (ev) cudaEventRecord(ev[0], st);
    const int nnb = nblk(d.ncap, TPB);
    static int occ = 0;
    if (!occ) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_quad_levels, TPB, 0); occ = std::max(1, std::min(occ, 8)); }
    cudaMemsetAsync(d.ext_bar, 0, sizeof(unsigned int), st);
    k_quad_levels<<<std::min(sm_count * occ, std::max(1, nnb)), TPB, 0, st>>>(d, s, d.lvl_list, d.quad, d.ext_bar); launches++;
    if (ev) cudaEventRecord(ev[1], st);
    const int smem = (int)sizeof(ExtWarp) * XWARPS;
    cudaFuncSetAttribute(k_walk_ext, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int64_t max_groups = (d.n / nparts + 256 + 31) / 32;
    k_walk_ext<<<(int)std::min<int64_t>((int64_t)sm_count * 3, (max_groups + XWARPS - 1) / XWARPS), XWARPS * 32, smem, st>>>(d, s, d.quad, theta, e0, part, nparts, use_quad ? 1 : 0); launches++;
    if (ev) { cudaEventRecord(ev[2], st); cudaEventRecord(ev[4], st); }
    if (any_gas) { k_ext_sph_force<<<nblk(d.n / nparts + 512, 128), 128, 0, st>>>(d, s, part, nparts); launches++; }
    if (ev) cudaEventRecord(ev[3], st);
    return launches;
}

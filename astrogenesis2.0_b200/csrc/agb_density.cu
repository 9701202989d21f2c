// agb_density.cu — visual density and node-group gas density on the device octree (sm_100a).
//
// Replaces Tree::calcVisualDensity + Node::calcVisualDensity (Physics/Tree/Tree.cpp:152-176,
// Physics/Tree/Node.cpp:832-877) and Tree::calcGasDensity + Node::calcGasDensity (Tree.cpp:119-150,
// Node.cpp:722-796, Math/kernel.cpp:4-16).
//
// The reference's density is NOT a neighbour search: every gas particle climbs from its leaf to the
// ancestor whose gasMass is closest to massInH, and every gas particle below that node gets
// h = 2*radius(node) and the same rho = sum_j m_j W(|x_j - COM_node|, h).  A later, higher group
// overwrites a lower one and an earlier higher group pre-empts lower ones, so the final state is
// order independent (SURVEY.md §8a "K5 restatement"): a particle takes the values of the TOPMOST
// node on its root path that is the first-stop node of some gas leaf; if there is none it is an
// orphan (h = 0, rho/P/T keep their previous values, no SPH force).  `childParticles` of a node is a
// contiguous range in tree order, so the sum is a segmented reduction; nodes where bulk insertion
// handed over to one-by-one insertion hold every particle twice (flag `ndup`), doubling rho.
//
// All kernels are HBM-bound streaming / pointer-chasing passes over O(N) data.
#include "agb_internal.cuh"
#include <algorithm>

namespace {

constexpr int TPB = 256;
constexpr double kPI = 3.14159265358979323846;   // Math/Constants.h:10
constexpr double kGAMMA = 5.0 / 3.0;             // Math/Constants.h:15
constexpr double kKB = 1.38064852e-23;           // Math/Constants.h:16
constexpr double kPRTN = 1.6726219e-27;          // Math/Constants.h:18

__device__ __forceinline__ double radius_at(double R, int depth) { return scalbn(R, -depth); }  // R/2/2/.. is exact

// ------------------------------------------------------------------ visual density
// Node::calcVisualDensity climbs while |rt - radius| > |rt - parent.radius|; radii are R*2^-k, so the
// test only depends on the level.  The root never computes (parent == nullptr, Node.cpp:834).
__global__ void __launch_bounds__(TPB) k_visual(AgbDev d, const uint32_t* __restrict__ perm, const AgbScalars* __restrict__ s, double rt)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x + d.dens_a0;
    if (i >= d.n || i >= d.dens_a1 || s->node_overflow) return;
    const uint32_t p = perm[i];
    double out = 0.0;                                          // Tree.cpp:156-161 zeroes every particle first
    if (i < s->n_in_tree) {
        const double R = __longlong_as_double((long long)s->Rbits);
        const int ld = d.leafdepth[i];
        int cur = ld;
        while (cur >= 1 && fabs(__dadd_rn(rt, -radius_at(R, cur))) > fabs(__dadd_rn(rt, -radius_at(R, cur - 1)))) cur--;
        if (cur >= 1) {
            double mass;
            if (cur == ld) mass = d.src_pm[i].w;
            else {
                int k = d.leafparent[i];                       // depth ld-1
                for (int j = ld - 1; j > cur; j--) k = d.nparent[k];
                mass = d.src_pm[d.n + k].w;
            }
            const double r = radius_at(R, cur);
            const double vol = __dmul_rn(__dmul_rn(r, r), r);
            if (!(vol == 0.0 || mass == 0.0)) {
                double dens = __ddiv_rn(mass, vol);
                if (!(dens == 0.0 || isinf(dens))) out = dens;
            }
        }
    }
    d.vis[p] = out;
}

// ------------------------------------------------------------------ gas density: mark first-stop nodes
__device__ __forceinline__ double gas_of(const AgbDev& d, int ref) { return d.src_gv[ref].w; }

// ---- exact gasMass of a node, as the reference accumulates it -------------------------------------
// With equal-mass gas particles and massInH a multiple of that mass, |M - g_node| == |M - g_parent|
// ties are common, and the reference breaks them by the rounding of ITS sums: `gasMass += p->mass` in
// list order = caller order restricted to the node (Node.cpp:477-480 bulk, :678-681 one-by-one).  The
// upward pass sums in octant order, which differs in the last bits.  So: (pass 0) climbs that meet a
// near-tie flag every ancestor that can still matter (gas <= 2 M) and park the particle; (fold) one
// block per flagged node recomputes the reference's left fold bit-exactly from the node's gas
// particles taken in caller order; (pass 1) parked particles climb again with those exact sums.
constexpr int FOLD_MAX = 8192;            // gas particles per node the fold kernel sorts in shared memory

__global__ void __launch_bounds__(TPB) k_gas_flags(AgbDev d, int32_t* __restrict__ flag)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < d.n) flag[i] = d.s_type[i] == 2;
}

__global__ void __launch_bounds__(TPB) k_gas_compact(AgbDev d, const uint32_t* __restrict__ perm, uint32_t* __restrict__ g_orig, double4* __restrict__ g_pm,
                                                       int32_t* __restrict__ g_tree, AgbScalars* s)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    const bool gas = i < d.n && d.s_type[i] == 2;
    unsigned long long mb = 0ull;
    if (gas) {
        const int r = d.gasrank[i];
        const double4 pm = d.src_pm[i];
        g_orig[r] = perm[i];
        g_pm[r] = pm;
        g_tree[r] = (int32_t)i;
        mb = (unsigned long long)__double_as_longlong(pm.w);       // masses are >= 0: bit order == value order
    }
    // smallest / largest gas mass of the step (one atomic pair per warp)
    unsigned long long lo = gas ? mb : ~0ull, hi = mb;
    for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0 && lo != ~0ull) {                   // equal masses: only the first few warps find anything to update
        if (lo < *(volatile unsigned long long*)&s->gas_mmin) atomicMin(&s->gas_mmin, lo);
        if (hi > *(volatile unsigned long long*)&s->gas_mmax) atomicMax(&s->gas_mmax, hi);
    }
}

struct GasFold {
    const int32_t* gasrank; const uint32_t* g_orig; const double4* g_pm; const int32_t* g_tree;
    uint8_t* nflag;            // per node: 0 = tree-order sum only, 1 = exact sum requested, 2 = exact sum ready, 3 = too large to fold
    double* nexact;            // per node: the reference's left-fold gasMass
    int32_t* foldlist;         // flagged nodes
    uint8_t* pending;          // per tree-order particle: parked in pass 0
};

__device__ __forceinline__ void node_gas_range(const AgbDev& d, const GasFold& F, const AgbScalars* s, int k, int* g0, int* g1)
{
    if (k == 0 && d.n >= (int64_t)d.cores * 100) { *g0 = 0; *g1 = s->n_gas_total; }     // the root sums every particle it is handed (Node.cpp:477-480)
    else { *g0 = F.gasrank[d.nfirst[k]]; *g1 = F.gasrank[d.nlast[k] + 1]; }
}

__device__ __forceinline__ void flag_node(const GasFold& F, AgbScalars* s, int k)
{
    // idempotent byte store + one list slot per node: the first writer (CAS on the aligned word holding the byte) appends
    unsigned int* word = reinterpret_cast<unsigned int*>(F.nflag + (k & ~3));
    const unsigned int sh = (k & 3) * 8;
    unsigned int old = *word;
    while (((old >> sh) & 0xffu) == 0u) {
        const unsigned int prev = atomicCAS(word, old, old | (1u << sh));
        if (prev == old) { F.foldlist[atomicAdd(&s->n_fold, 1)] = k; break; }
        old = prev;
    }
}

template <int PASS>
__global__ void __launch_bounds__(TPB) k_gas_mark(AgbDev d, AgbScalars* s, double M, GasFold F)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= s->n_in_tree || s->node_overflow) return;
    if (d.s_type[i] != 2) return;
    if (PASS == 1 && !F.pending[i]) return;
    const int N = (int)d.n;
    const double tau = 1e-9;
    // cur = -1 encodes "the particle's own leaf"
    int cur = -1;
    double g = d.src_gv[i].w;                                  // leaf gasMass = particle mass (Node.cpp:416)
    bool g_exact = true;
    while (true) {
        if (g == 0.0) return;                                  // Node.cpp:724
        int par = cur < 0 ? d.leafparent[i] : d.nparent[cur];
        if (par < 0) return;                                   // root: parent == nullptr (Node.cpp:727)
        double gp = gas_of(d, N + par);
        bool gp_exact = false;
        if (PASS == 1 && F.nflag[par] == 2) { gp = F.nexact[par]; gp_exact = true; }
        double d0 = fabs(__dadd_rn(M, -g)), d1 = fabs(__dadd_rn(M, -gp));
        if (fabs(g - M) <= tau * M || fabs(d0 - d1) <= tau * fmax(M, gp)) {
            // a decision of Node.cpp:737-751 hangs on the last bits of the two sums
            int a0, a1, b0, b1;
            if (cur < 0) { a0 = 0; a1 = 1; } else node_gas_range(d, F, s, cur, &a0, &a1);
            node_gas_range(d, F, s, par, &b0, &b1);
            if (a1 - a0 == b1 - b0) { gp = g; gp_exact = g_exact; d1 = d0; }   // same gas particles in the same order: the reference's sums are identical
            else if (PASS == 0) {
                // park: request exact sums for this node and for every ancestor that can still take part in a tie
                if (cur >= 0) flag_node(F, s, cur);
                for (int a = par; a >= 0; a = d.nparent[a]) { flag_node(F, s, a); if (gas_of(d, N + a) > 2.0 * M * (1.0 + 1e-6)) break; }
                F.pending[i] = 1;
                return;
            } else {
                atomicAdd((g_exact && gp_exact) ? &s->tie_exact : &s->tie_unresolved, 1);
            }
        }
        if (g < M && d0 > d1) { cur = par; g = gp; g_exact = gp_exact; continue; } // Node.cpp:737-746 (after the recursion the 2nd test is false)
        if (d0 < d1) {                                         // Node.cpp:749-751: compute here
            if (cur < 0) d.leafmark[i] = 1; else d.nmark[cur] = 1;
        }
        return;
    }
}

// one block per flagged node: gas particles -> shared memory, bitonic sort by caller index, serial left fold
__global__ void __launch_bounds__(TPB) k_gas_fold(AgbDev d, AgbScalars* s, GasFold F)
{
    extern __shared__ __align__(16) unsigned char fold_smem[];
    double* sm_m = reinterpret_cast<double*>(fold_smem);
    uint32_t* sm_i = reinterpret_cast<uint32_t*>(sm_m + FOLD_MAX);
    if (s->node_overflow) return;
    if (s->gas_mmin == s->gas_mmax) {
        // Every gas particle has the same mass m (the usual SPH initial conditions): the reference's left fold over a node's c gas
        // particles is m added c times, whatever their order.  One table of those sums per block, no sorting.
        if (threadIdx.x == 0) {
            const double m = __longlong_as_double((long long)s->gas_mmin);
            double sum = 0.0;
            sm_m[0] = 0.0;
            for (int j = 1; j <= FOLD_MAX; j++) { sum = __dadd_rn(sum, m); sm_m[j] = sum; }
        }
        __syncthreads();
        for (int q = blockIdx.x * TPB + threadIdx.x; q < s->n_fold; q += gridDim.x * TPB) {
            const int k = F.foldlist[q];
            int g0, g1;
            node_gas_range(d, F, s, k, &g0, &g1);
            const int cnt = g1 - g0;
            if (cnt > FOLD_MAX) { F.nflag[k] = 3; continue; }
            F.nexact[k] = sm_m[cnt];
            F.nflag[k] = 2;
        }
        return;
    }
    for (int q = blockIdx.x; q < s->n_fold; q += gridDim.x) {
        const int k = F.foldlist[q];
        int g0, g1;
        node_gas_range(d, F, s, k, &g0, &g1);
        const int cnt = g1 - g0;
        if (cnt > FOLD_MAX) { if (threadIdx.x == 0) F.nflag[k] = 3; continue; }
        int p2 = 1;
        while (p2 < cnt) p2 <<= 1;
        __syncthreads();
        for (int j = threadIdx.x; j < p2; j += TPB) {
            sm_i[j] = j < cnt ? F.g_orig[g0 + j] : 0xffffffffu;
            sm_m[j] = j < cnt ? F.g_pm[g0 + j].w : 0.0;
        }
        __syncthreads();
        for (int size = 2; size <= p2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int j = threadIdx.x; j < p2; j += TPB) {
                    const int partner = j ^ stride;
                    if (partner > j) {
                        const bool up = (j & size) == 0;
                        const uint32_t a = sm_i[j], b = sm_i[partner];
                        if ((a > b) == up) { sm_i[j] = b; sm_i[partner] = a; const double t = sm_m[j]; sm_m[j] = sm_m[partner]; sm_m[partner] = t; }
                    }
                }
                __syncthreads();
            }
        if (threadIdx.x == 0) {
            double sum = 0.0;
            for (int j = 0; j < cnt; j++) sum = __dadd_rn(sum, sm_m[j]);
            F.nexact[k] = sum;
            __threadfence();
            F.nflag[k] = 2;
        }
    }
}

// topmost marked node on the root path -> group id: N+k (node), i (own leaf) or -1 (orphan)
// LATE: U and mu have not arrived yet; P and T are written by k_gather_late (agb_build.cu)
template <bool LATE>
__global__ void __launch_bounds__(TPB) k_gas_group(AgbDev d, AgbScalars* s)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    bool orphan = false;
    if (i < s->n_in_tree && i >= d.dens_a0 && i < d.dens_a1 && !s->node_overflow && d.s_type[i] == 2) {
        int grp = d.leafmark[i] ? (int)i : -1;
        for (int k = d.leafparent[i]; k >= 0; k = d.nparent[k]) if (d.nmark[k]) grp = (int)d.n + k;
        d.group[i] = grp;
        orphan = grp < 0;
        if (grp == (int)i) {
            // the particle alone is its group: rho = m W(|x - COM_leaf|, 2 r_leaf), COM_leaf == x (Node.cpp:414)
            const double R = __longlong_as_double((long long)s->Rbits);
            const double h = __dmul_rn(radius_at(R, d.leafdepth[i]), 2.0);
            const double a = 1.0 / (kPI * h * h * h);
            const double rho = d.src_pm[i].w * a;
            d.s_h[i] = h; d.s_rho[i] = rho;
            if (!LATE) {
                d.s_P[i] = (kGAMMA - 1.0) * d.s_U[i] * rho;
                d.s_T[i] = (kGAMMA - 1.0) * d.s_U[i] * kPRTN * d.s_mu[i] / kKB;
            }
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, orphan);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s->n_gas_orphans, __popc(m));
}

// marked nodes with no marked ancestor are the groups that survive; compact them
__global__ void __launch_bounds__(TPB) k_gas_collect(AgbDev d, AgbScalars* s)
{
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= s->n_nodes || s->node_overflow || !d.nmark[k]) return;
    for (int a = d.nparent[k]; a >= 0; a = d.nparent[a]) if (d.nmark[a]) return;
    d.grouplist[atomicAdd(&s->n_gas_groups, 1)] = k;
}

__device__ __forceinline__ double spline_w(double r, double h)
{   // Math/kernel.cpp:4-16
    const double a = 1.0 / (kPI * h * h * h);
    const double q = r / h;
    if (q < 1.0) return a * (1 - 1.5 * q * q + 0.75 * q * q * q);
    if (q < 2.0) { double t = 2 - q; return a * 0.25 * (t * t * t); }
    return 0.0;
}

// one warp per surviving group: fixed-shape (lane-strided, then butterfly) sum => deterministic
template <bool LATE>
__global__ void __launch_bounds__(TPB) k_gas_sum(AgbDev d, const AgbScalars* __restrict__ s, GasFold F)
{
    const int lane = threadIdx.x & 31;
    const double R = __longlong_as_double((long long)s->Rbits);
    if (s->node_overflow) return;
    for (int g = (blockIdx.x * TPB + threadIdx.x) >> 5; g < s->n_gas_groups; g += (gridDim.x * TPB) >> 5) {
    const int k = d.grouplist[g];
    if (d.nfirst[k] >= d.dens_a1 || d.nlast[k] < d.dens_a0) continue;     // none of its particles is asked for
    const double h = __dmul_rn(radius_at(R, d.ndepth[k]), 2.0);          // Node.cpp:765
    const double4 com = d.src_pm[d.n + k];
    // the node's gas particles are a contiguous range of the compact gas list (tree order)
    const int g0 = F.gasrank[d.nfirst[k]], g1 = F.gasrank[d.nlast[k] + 1];
    double acc = 0.0;
    for (int r = g0 + lane; r < g1; r += 32) {
        const double4 pm = F.g_pm[r];
        const double dx = pm.x - com.x, dy = pm.y - com.y, dz = pm.z - com.z;
        acc += pm.w * spline_w(sqrt(dx * dx + dy * dy + dz * dz), h);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (d.ndup[k]) acc += acc;                                            // every particle listed twice (Node.cpp:518 + :615)
    for (int r = g0 + lane; r < g1; r += 32) {
        const int j = F.g_tree[r];
        d.s_h[j] = h; d.s_rho[j] = acc;
        if (!LATE) {
            d.s_P[j] = (kGAMMA - 1.0) * d.s_U[j] * acc;                       // Node.cpp:789
            d.s_T[j] = (kGAMMA - 1.0) * d.s_U[j] * kPRTN * d.s_mu[j] / kKB;   // Node.cpp:791
        }
    }
    }
}

// back to caller order (gas only, through the compact list).  LATE: h and the new rho only; orphans keep the rho they came with
template <bool LATE>
__global__ void __launch_bounds__(TPB) k_gas_scatter(AgbDev d, const AgbScalars* __restrict__ s, GasFold F)
{
    const int r = blockIdx.x * TPB + threadIdx.x;
    if (r >= s->n_gas_total) return;
    const int i = F.g_tree[r];
    if (i < d.dens_a0 || i >= d.dens_a1) return;
    const uint32_t p = F.g_orig[r];
    if (LATE) { d.h[p] = d.s_h[i]; if (i < s->n_in_tree && d.group[i] >= 0) d.rho[p] = d.s_rho[i]; }
    else { d.h[p] = d.s_h[i]; d.rho[p] = d.s_rho[i]; d.P[p] = d.s_P[i]; d.T[p] = d.s_T[i]; }
}

__global__ void k_gas_reset(AgbScalars* s)
{
    s->n_gas_groups = 0; s->n_gas_orphans = 0; s->tie_exact = 0; s->tie_unresolved = 0; s->n_fold = 0;
    s->gas_mmin = ~0ull; s->gas_mmax = 0ull;
}

} // namespace

static inline int nblk(int64_t n, int per) { return (int)((n + per - 1) / per); }

int agb_launch_visual(AgbDev& d, AgbScalars* s, double radius, cudaStream_t st)
{
    const int64_t cnt = std::min<int64_t>(d.n, d.dens_a1) - d.dens_a0;
    if (cnt > 0) k_visual<<<nblk(cnt, TPB), TPB, 0, st>>>(d, d.perm[d.cur], s, radius);
    return 1;
}

// tree positions of the gas particles, compact and in tree order, in d.nodecnt; their number in s->n_gas_total (device)
int agb_launch_gas_list(AgbDev& d, AgbScalars* s, cudaStream_t st)
{
    const int nb = nblk(d.n, TPB);
    k_gas_reset<<<1, 1, 0, st>>>(s);
    k_gas_flags<<<nb, TPB, 0, st>>>(d, d.nodecnt);
    int launches = agb_launch_scan_i32(d.nodecnt, d.gasrank, d.n, d.scanblk, &s->n_gas_total, st);
    cudaMemcpyAsync(d.gasrank + d.n, &s->n_gas_total, sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
    k_gas_compact<<<nb, TPB, 0, st>>>(d, d.perm[d.cur], d.perm[d.cur ^ 1], d.rec, d.nodecnt, s);
    return launches + 3;
}

int agb_launch_gas_density(AgbDev& d, AgbScalars* s, double massInH, cudaStream_t st, bool late_pt)
{
    const int nb = nblk(d.n, TPB);
    k_gas_reset<<<1, 1, 0, st>>>(s);
    // compact (caller index, mass) of the gas particles in tree order; scratch that is free after the build is reused:
    // flags -> nodecnt, caller indices -> the idle half of the sort's ping-pong permutation, masses -> rec
    uint32_t* g_orig = d.perm[d.cur ^ 1];
    double4* g_pm = d.rec;                                  // (x, y, z, m) of the gas particles, compact, tree order
    int32_t* g_tree = d.nodecnt + 0;                        // their tree positions (the flags in nodecnt are dead after the scan)
    k_gas_flags<<<nb, TPB, 0, st>>>(d, d.nodecnt);
    int launches = agb_launch_scan_i32(d.nodecnt, d.gasrank, d.n, d.scanblk, &s->n_gas_total, st);
    cudaMemcpyAsync(d.gasrank + d.n, &s->n_gas_total, sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
    k_gas_compact<<<nb, TPB, 0, st>>>(d, d.perm[d.cur], g_orig, g_pm, g_tree, s);
    // nflag/pending/foldlist/nexact borrow scratch that is idle here: arrived (int32/node), lcp (int8/particle),
    // nodebase (int32/particle), and the caller-order copy of key_lo (8 B/particle)
    GasFold F{d.gasrank, g_orig, g_pm, g_tree, reinterpret_cast<uint8_t*>(d.arrived), reinterpret_cast<double*>(d.klo[0]), d.nodebase,
              reinterpret_cast<uint8_t*>(d.lcp)};
    cudaMemsetAsync(d.arrived, 0, (size_t)d.ncap * sizeof(int32_t), st);
    cudaMemsetAsync(d.lcp, 0, (size_t)d.n, st);
    k_gas_mark<0><<<nb, TPB, 0, st>>>(d, s, massInH, F);
    const int fold_smem = (FOLD_MAX + 1) * 12;
    cudaFuncSetAttribute(k_gas_fold, cudaFuncAttributeMaxDynamicSharedMemorySize, fold_smem);
    k_gas_fold<<<296, TPB, fold_smem, st>>>(d, s, F);
    k_gas_mark<1><<<nb, TPB, 0, st>>>(d, s, massInH, F);
    if (late_pt) k_gas_group<true><<<nb, TPB, 0, st>>>(d, s); else k_gas_group<false><<<nb, TPB, 0, st>>>(d, s);
    k_gas_collect<<<nblk(d.ncap, TPB), TPB, 0, st>>>(d, s);
    const int sumb = std::min(nblk(d.n, TPB / 32), 148 * 16);                      // one warp per group, grid-stride
    if (late_pt) { k_gas_sum<true><<<sumb, TPB, 0, st>>>(d, s, F); k_gas_scatter<true><<<nb, TPB, 0, st>>>(d, s, F); }
    else { k_gas_sum<false><<<sumb, TPB, 0, st>>>(d, s, F); k_gas_scatter<false><<<nb, TPB, 0, st>>>(d, s, F); }
    return 10 + launches;
}

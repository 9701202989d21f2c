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
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n) return;
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

__global__ void __launch_bounds__(TPB) k_gas_mark(AgbDev d, const AgbScalars* __restrict__ s, double M)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= s->n_in_tree) return;
    if (d.s_type[i] != 2) return;
    const int N = (int)d.n;
    // cur = -1 encodes "the particle's own leaf"
    int cur = -1;
    double g = d.src_gv[i].w;                                  // leaf gasMass = particle mass (Node.cpp:416)
    while (true) {
        if (g == 0.0) return;                                  // Node.cpp:724
        int par = cur < 0 ? d.leafparent[i] : d.nparent[cur];
        if (par < 0) return;                                   // root: parent == nullptr (Node.cpp:727)
        double gp = gas_of(d, N + par);
        double d0 = fabs(__dadd_rn(M, -g)), d1 = fabs(__dadd_rn(M, -gp));
        if (g < M && d0 > d1) { cur = par; g = gp; continue; } // Node.cpp:737-746 (after the recursion the 2nd test is false)
        if (d0 < d1) {                                         // Node.cpp:749-751: compute here
            if (cur < 0) d.leafmark[i] = 1; else d.nmark[cur] = 1;
        }
        return;
    }
}

// topmost marked node on the root path -> group id: N+k (node), i (own leaf) or -1 (orphan)
__global__ void __launch_bounds__(TPB) k_gas_group(AgbDev d, AgbScalars* s)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    bool orphan = false;
    if (i < s->n_in_tree && d.s_type[i] == 2) {
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
            d.s_P[i] = (kGAMMA - 1.0) * d.s_U[i] * rho;
            d.s_T[i] = (kGAMMA - 1.0) * d.s_U[i] * kPRTN * d.s_mu[i] / kKB;
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, orphan);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s->n_gas_orphans, __popc(m));
}

// marked nodes with no marked ancestor are the groups that survive; compact them
__global__ void __launch_bounds__(TPB) k_gas_collect(AgbDev d, AgbScalars* s)
{
    int k = blockIdx.x * TPB + threadIdx.x;
    if (k >= s->n_nodes || !d.nmark[k]) return;
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
__global__ void __launch_bounds__(TPB) k_gas_sum(AgbDev d, const AgbScalars* __restrict__ s)
{
    const int g = (blockIdx.x * TPB + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= s->n_gas_groups) return;
    const int k = d.grouplist[g];
    const double R = __longlong_as_double((long long)s->Rbits);
    const double h = __dmul_rn(radius_at(R, d.ndepth[k]), 2.0);          // Node.cpp:765
    const double4 com = d.src_pm[d.n + k];
    const int first = d.nfirst[k], last = d.nlast[k];
    double acc = 0.0;
    for (int j = first + lane; j <= last; j += 32) {
        if (d.s_type[j] != 2) continue;
        double4 pm = d.src_pm[j];
        double dx = pm.x - com.x, dy = pm.y - com.y, dz = pm.z - com.z;
        acc += pm.w * spline_w(sqrt(dx * dx + dy * dy + dz * dz), h);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (d.ndup[k]) acc += acc;                                            // every particle listed twice (Node.cpp:518 + :615)
    for (int j = first + lane; j <= last; j += 32) {
        if (d.s_type[j] != 2) continue;
        d.s_h[j] = h; d.s_rho[j] = acc;
        d.s_P[j] = (kGAMMA - 1.0) * d.s_U[j] * acc;                       // Node.cpp:789
        d.s_T[j] = (kGAMMA - 1.0) * d.s_U[j] * kPRTN * d.s_mu[j] / kKB;   // Node.cpp:791
    }
}

// back to caller order
__global__ void __launch_bounds__(TPB) k_gas_scatter(AgbDev d, const uint32_t* __restrict__ perm)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n || d.s_type[i] != 2) return;
    const uint32_t p = perm[i];
    d.h[p] = d.s_h[i]; d.rho[p] = d.s_rho[i]; d.P[p] = d.s_P[i]; d.T[p] = d.s_T[i];
}

__global__ void k_gas_reset(AgbScalars* s) { s->n_gas_groups = 0; s->n_gas_orphans = 0; }

} // namespace

static inline int nblk(int64_t n, int per) { return (int)((n + per - 1) / per); }

int agb_launch_visual(AgbDev& d, AgbScalars* s, double radius, cudaStream_t st)
{
    k_visual<<<nblk(d.n, TPB), TPB, 0, st>>>(d, d.perm[d.cur], s, radius);
    return 1;
}

int agb_launch_gas_density(AgbDev& d, AgbScalars* s, double massInH, cudaStream_t st)
{
    const int nb = nblk(d.n, TPB);
    k_gas_reset<<<1, 1, 0, st>>>(s);
    k_gas_mark<<<nb, TPB, 0, st>>>(d, s, massInH);
    k_gas_group<<<nb, TPB, 0, st>>>(d, s);
    k_gas_collect<<<nb, TPB, 0, st>>>(d, s);
    k_gas_sum<<<nblk(d.n, TPB / 32), TPB, 0, st>>>(d, s);     // upper bound on groups; surplus warps exit
    k_gas_scatter<<<nb, TPB, 0, st>>>(d, d.perm[d.cur]);
    return 6;
}

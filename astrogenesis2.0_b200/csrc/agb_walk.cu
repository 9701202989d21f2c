// agb_walk.cu — warp-cooperative Barnes–Hut walk and the SPH pair kernel (sm_100a).
//
// Replaces Tree::calculateForces (Physics/Tree/Tree.cpp:57-83), Node::calculateGravityForce
// (Physics/Tree/Node.cpp:247-399) and Node::calcSPHForce (Node.cpp:88-172, Math/kernel.cpp:18-39).
//
// Parity contract: every target must interact with exactly the (node | leaf) set the reference's
// per-particle recursion accepts: MAC `radius / r < theta` with radius = cell half-width and r the
// distance to the node's centre of mass, tested at every level (single-child chains included);
// force law  a += G M d / (r (r^2 + e0^2))  (the reference's spline softening is dead code for
// e0 > 2.15e13); SPH pressure + Monaghan–Gingold viscosity + dU/dt against every
// accepted node/leaf that holds gas and lies within r < 2 h_i, with the target's own h, rho, P.
//
// B200 mapping: one warp owns 32 tree-adjacent targets.  k_far walks the upper tree once per 256
// targets; k_walk continues per warp on the UNION of its targets' trees with a shared-memory stack
// of (node, lane-mask) pairs.  Each lane pops a different node and classifies it against the
// bounding box of the warp's targets:
//     box entirely beyond radius/theta   -> accepted by every lane in the mask
//     box entirely inside radius/theta   -> opened by every lane in the mask (children inherit mask)
//     straddling                         -> per-lane test (__ballot_sync splits the mask)
// so the per-target accepted set is exact while most of the tree is pruned at 1/32 of the cost.
// Accepted sources go to a shared-memory interaction list; the list is drained in tiles of 32:
// lanes gather the 32 sources (double4 loads) into a staging buffer, then every lane runs all of
// them against its own target out of shared memory (broadcast reads).  No tensor cores: this is
// not a dense contraction.  Pair arithmetic is FP32 with float-float displacements where pairs are
// close (default "mixed" mode; three evaluation classes, packed FFMA2/FADD2, one MUFU per pair) or
// FP64 throughout.  Decisions that must match the reference bit for bit (MAC, r < 2h) are taken in
// FP32 only where that provably gives the FP64 answer, otherwise in FP64, and within 1e-13 of the
// threshold by the reference's own separately rounded expression.
// SPH pairs: inside the walk in FP64 mode; in mixed mode the walk records (source, accepting gas
// targets) entries and k_sph evaluates them afterwards (keeps the walk's hot code inside the 32 KB
// instruction cache).
#include "agb_internal.cuh"
#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace {

#ifndef AGB_WALK_LCAP
#define AGB_WALK_LCAP 512
#endif
constexpr int LCAP = AGB_WALK_LCAP;        // interaction-list entries per warp
constexpr int LGROW = 288;                 // worst-case growth per pop round: 32 lanes x (8 leaves + 1 node)
constexpr int PCAP = 64;                   // straddling nodes waiting for their per-target tests (< 32 left over + 32 new)
#ifndef AGB_WALK_CTAS_PER_SM
#define AGB_WALK_CTAS_PER_SM 2
#endif
constexpr int WALK_CTAS = AGB_WALK_CTAS_PER_SM;
#ifndef AGB_WALK_SCAP
#define AGB_WALK_SCAP 384
#endif
constexpr int SCAP = AGB_WALK_SCAP;   // shared part of the traversal stack (the rest spills to global memory)
constexpr double kG = 6.67430e-11;         // Math/Constants.h:7
constexpr double kPI = 3.14159265358979323846;
constexpr double kGAMMA = 5.0 / 3.0;

// Per-warp shared memory.  In-walk SPH operands exist only in the FP64 SPH kernels; the mixed-precision SPH kernels only
// record, per tile, which (target, source) pairs lie within ~2h for k_sph; the gravity-only kernels need 8 KB per warp.
constexpr int REC_BATCH = 8;               // tile records a warp takes from the pool at a time
template <bool INWALK> struct SphSmem { double4 tsph[1], gst[1]; };
template <> struct SphSmem<true> {
    double4 tsph[96];                      // per gas target: (1/h, 1/(pi h^4), 2 P/rho^2, sound speed), (vx, vy, vz, h)
    double4 gst[32];                       // drain: (velocity | mVel, gasMass) of the tile's gas-bearing sources
};
template <bool SPH, bool MIXED> struct WarpSmem : SphSmem<SPH && !MIXED> {
    int2 rsrc[SPH && MIXED ? 64 : 2];      // SPLIT: (source, gas targets that accepted it) waiting for the next k_sph record
    int2 list[LCAP];
    int2 stack[SCAP];
    int2 pend[PCAP];                       // (node, lane mask) of straddling nodes, resolved 32 at a time (one node per lane, loop over targets)
    float4 tpos[32];                       // minus the targets' positions about the box centre (units of L) + their tree positions
    double4 stage[32];                     // drain: the 32 sources of a tile
};
// k_sph: per gas target (1/h, 1/(pi h^4), 2 P/rho^2, sound speed), (vx, vy, vz, h), 8 floats: -(float-float position), (2h/L)^2, tree position
struct SphWarp {
    double4 tsph[96];
    double4 res[32];                       // pair results (fx, fy, fz, dU) on their way to the owning lane
    double4 gst[32];                       // (velocity | mVel, gasMass) of the tile's sources
    double4 stage[32];                     // float-float records of the tile's sources, two per 64-byte record
    int list[32];                          // the record's sources
    unsigned lmask[32];                    // ... and the gas targets that accepted each (0 for sources without gas)
};
#ifndef AGB_WALK_WARPS
#define AGB_WALK_WARPS 8
#endif
template <bool SPH> struct WalkCfg { static constexpr int WARPS = AGB_WALK_WARPS, TPB = WARPS * 32; };

// 1/sqrt(x) for positive normal x: MUFU.RSQ64H seed (~2^-22) + one cubically convergent step (~2^-60).
// None of the special-case handling of the library rsqrt() is needed here (x = 0 is masked by the caller).
__device__ __forceinline__ double rsqrt_pos(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}

struct WalkParams {
    const double4 *src_pm, *src_gv;
    const uint8_t* src_flag;
    const int32_t* child;
    const int8_t* ndepth;
    const double *s_h, *s_rho, *s_P, *s_next;
    const uint8_t* s_type;
    const uint32_t* perm;
    double *ax, *ay, *az, *dUdt;
    int32_t *c_visits, *c_accn, *c_accl, *c_sph;
    int2* spill; int64_t spill_per_warp;
    int32_t *far_list, *far_front, *far_cnt;   // per super-group (256 targets): shared accept list, hand-over frontier, {n_list, n_front, n_visits}
    AgbScalars* s;
    int64_t N;
    const int32_t* act_list;                 // tree positions of the active targets (unused when every particle is active)
    int part, nparts;                        // this call walks the part-th of nparts slices of the active targets
    double theta, e0, globalTime;
    // mixed SPH: records written by k_walk for k_sph: 32 x (source, mask of the gas targets that accepted it) and a link to the
    // group's previous record; rec_head[g] = last record of group g (-1: none)
    int2* rec_ent; int32_t *rec_next, *rec_head; int64_t rec_cap;
    float far_k2;                            // squared "far" distance in half-diagonals of the warp's box (mixed mode, class split)
};

enum { OUT_NONE = 0, OUT_ACCEPT = 1, OUT_OPEN = 2, OUT_MIXED = 3 };

__device__ __forceinline__ double warp_min(double v) { for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ double warp_max(double v) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ float warp_max_f(float v) { return __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(v))); }   // v >= 0
__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    return v;
}
// position of the n-th (0-based) set bit of m, n < popc(m): five popcount halvings (the library's __fns is a ~40-instruction loop)
__device__ __forceinline__ int nth_set_bit(unsigned m, int n)
{
    int pos = 0, c;
    c = __popc(m & 0xffffu); if (n >= c) { n -= c; pos = 16; m >>= 16; }
    c = __popc(m & 0xffu);   if (n >= c) { n -= c; pos += 8; m >>= 8; }
    c = __popc(m & 0xfu);    if (n >= c) { n -= c; pos += 4; m >>= 4; }
    c = __popc(m & 0x3u);    if (n >= c) { n -= c; pos += 2; m >>= 2; }
    if (n >= (int)(m & 1u)) pos += 1;
    return pos;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


constexpr int SG_GROUPS = 8;               // warps (32-target groups) per super-group
constexpr int FAR_LCAP = 4096, FAR_FCAP = 2048, FAR_SCAP = 1024;

struct Box { double lox, loy, loz, hix, hiy, hiz; };

// The targets of a call: the active particles (Tree.cpp:75) in tree order, compacted, cut into `nparts` slices whose
// boundaries are multiples of 256 (one far-field super-group) so that groups are the same for any number of GPUs.
struct Slice { int64_t a0, a1; bool ident; };
__device__ __forceinline__ Slice target_slice(const WalkParams& P)
{
    const int64_t na = P.s->n_active, nsg = (na + 32 * SG_GROUPS - 1) / (32 * SG_GROUPS);
    Slice sl;
    sl.a0 = min(na, nsg * P.part / P.nparts * (32 * SG_GROUPS));
    sl.a1 = min(na, nsg * (P.part + 1) / P.nparts * (32 * SG_GROUPS));
    sl.ident = na == P.N;
    return sl;
}

// Opening test of a node against a box of targets: ACCEPT / OPEN only when every point of the box takes the same
// decision as the reference's `radius / r < theta` (Node.cpp:331-334) with a 1e-12 margin; r == 0 is impossible
// for OPEN because the COM must lie outside the box.
__device__ __forceinline__ int classify_box(const double4& pm, double rad2, const Box& b, double theta2, bool fast_mac)
{
    const double ax_ = fmax(0.0, fmax(b.lox - pm.x, pm.x - b.hix)), ay_ = fmax(0.0, fmax(b.loy - pm.y, pm.y - b.hiy)), az_ = fmax(0.0, fmax(b.loz - pm.z, pm.z - b.hiz));
    const double bx_ = fmax(fabs(pm.x - b.lox), fabs(pm.x - b.hix)), by_ = fmax(fabs(pm.y - b.loy), fabs(pm.y - b.hiy)), bz_ = fmax(fabs(pm.z - b.loz), fabs(pm.z - b.hiz));
    const double dmin2 = ax_ * ax_ + ay_ * ay_ + az_ * az_, dmax2 = bx_ * bx_ + by_ * by_ + bz_ * bz_;
    if (fast_mac) {
        if (dmin2 * theta2 > rad2 * (1.0 + 1e-12)) return OUT_ACCEPT;
        if (dmin2 > 0.0 && dmax2 * theta2 < rad2 * (1.0 - 1e-12)) return OUT_OPEN;
    }
    return OUT_MIXED;
}

// Far-field prepass: one warp per super-group of 256 tree-adjacent targets walks the tree once against the box of
// all of them.  Nodes every target accepts go to a list shared by the 8 warps that own those targets, nodes every
// target opens are descended here, and nodes that straddle the opening radius of the big box are handed over to the
// per-warp walks (k_walk) as their starting frontier.  The upper tree is thus traversed once per 256 targets
// instead of once per 32, and the shared entries are evaluated with full lane masks.
__global__ void __launch_bounds__(128) k_far(const WalkParams P)
{
    __shared__ int stack_s[4][FAR_SCAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* const stk = stack_s[warp];
    const unsigned lt = (1u << lane) - 1u;
    const int N = (int)P.N;
    const double R = __longlong_as_double((long long)P.s->Rbits);
    const int n_nodes = P.s->n_nodes, n_in_tree = P.s->n_in_tree;
    const double theta2 = P.theta * P.theta;
    const bool fast_mac = P.theta > 0.0;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const Slice sl = target_slice(P);
    const int nsg = (int)((sl.a1 - sl.a0 + 32 * SG_GROUPS - 1) / (32 * SG_GROUPS));
    if (P.s->node_overflow || P.s->walk_overflow == 3) return;      // no usable tree this step (the host grows the node arrays and rebuilds) / densities not for these targets
    for (int sgi = blockIdx.x * 4 + warp; sgi < nsg; sgi += gridDim.x * 4) {
        const int64_t tb = sl.a0 + (int64_t)sgi * (32 * SG_GROUPS);
        int32_t* const fl = P.far_list + (size_t)sgi * FAR_LCAP;
        int32_t* const ff = P.far_front + (size_t)sgi * FAR_FCAP;
        Box b{inf, inf, inf, -inf, -inf, -inf};
        for (int j = 0; j < SG_GROUPS; j++) {
            const int64_t idx = tb + j * 32 + lane;
            if (idx < sl.a1) {
                const double4 tp = P.src_pm[sl.ident ? idx : (int64_t)P.act_list[idx]];
                if (tp.w != 0.0) {
                    b.lox = fmin(b.lox, tp.x); b.loy = fmin(b.loy, tp.y); b.loz = fmin(b.loz, tp.z);
                    b.hix = fmax(b.hix, tp.x); b.hiy = fmax(b.hiy, tp.y); b.hiz = fmax(b.hiz, tp.z);
                }
            }
        }
        b.lox = warp_min(b.lox); b.loy = warp_min(b.loy); b.loz = warp_min(b.loz);
        b.hix = warp_max(b.hix); b.hiy = warp_max(b.hiy); b.hiz = warp_max(b.hiz);
        int nl = 0, nf = 0, nvis = 0, sp = 0;
        if (b.lox <= b.hix) {
            if (n_nodes > 0) { if (lane == 0) stk[0] = N; sp = 1; }
            else if (n_in_tree == 1) { if (lane == 0) fl[0] = 0; nl = 1; }
            __syncwarp();
            while (sp > 0) {
                if (nl > FAR_LCAP - 32 * 9 || nf > FAR_FCAP - FAR_SCAP - 64) {
                    // out of room (tiny opening angles): stop here and hand every pending node to the per-warp walks
                    for (int i = lane; i < sp; i += 32) ff[nf + i] = stk[i];
                    nf += sp; sp = 0;
                    break;
                }
                const int cnt = min(sp, 32);
                sp -= cnt;
                const int node = lane < cnt ? stk[sp + lane] : -1;
                int outcome = OUT_NONE;
                int4 c0 = make_int4(-1, -1, -1, -1), c1 = c0;
                if (node >= 0) {
                    c0 = __ldg(reinterpret_cast<const int4*>(P.child) + 2 * (size_t)(node - N));
                    c1 = __ldg(reinterpret_cast<const int4*>(P.child) + 2 * (size_t)(node - N) + 1);
                    const double4 pm = P.src_pm[node];
                    if (pm.w != 0.0) {
                        const double rad = scalbn(R, -(int)P.ndepth[node - N]);
                        outcome = classify_box(pm, rad * rad, b, theta2, fast_mac);
                        // out of room on the stack: leave the node (and all below it) to the per-warp walks
                        if (outcome == OUT_OPEN && sp > FAR_SCAP - 32 * 9) outcome = OUT_MIXED;
                    }
                }
                __syncwarp();
                nvis += __popc(__ballot_sync(0xffffffffu, (outcome == OUT_ACCEPT || outcome == OUT_OPEN) && node != N));
                const unsigned fm = __ballot_sync(0xffffffffu, outcome == OUT_MIXED);
                if (outcome == OUT_MIXED) { const int pos = nf + __popc(fm & lt); if (pos < FAR_FCAP) ff[pos] = node; else P.s->walk_overflow = 1; }
                nf = min(nf + __popc(fm), FAR_FCAP);
                const unsigned am = __ballot_sync(0xffffffffu, outcome == OUT_ACCEPT);
                if (outcome == OUT_ACCEPT) fl[nl + __popc(am & lt)] = node;
                nl += __popc(am);
                const unsigned om = __ballot_sync(0xffffffffu, outcome == OUT_OPEN);
                if (om) {
                    int ch[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
                    int cl = 0, cn = 0;
                    if (outcome == OUT_OPEN) {
                        ch[0] = c0.x; ch[1] = c0.y; ch[2] = c0.z; ch[3] = c0.w; ch[4] = c1.x; ch[5] = c1.y; ch[6] = c1.z; ch[7] = c1.w;
#pragma unroll
                        for (int c = 0; c < 8; c++) { cl += (ch[c] >= 0 && ch[c] < N); cn += (ch[c] >= N); }
                    }
                    // both counts are <= 8 per lane: one scan of the packed pair (leaves in the low half, nodes in the high half)
                    const int sc = warp_incl_scan(cl | (cn << 16), lane), st_ = __shfl_sync(0xffffffffu, sc, 31);
                    const int il = sc & 0xffff, in_ = sc >> 16, tl = st_ & 0xffff, tn = st_ >> 16;
                    if (outcome == OUT_OPEN) {
                        int pl = nl + il - cl, pn = sp + in_ - cn;
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            if (ch[c] >= N) stk[pn++] = ch[c];
                            else if (ch[c] >= 0) fl[pl++] = ch[c];
                        }
                    }
                    nl += tl; sp += tn;
                }
                __syncwarp();
            }
        }
        if (lane == 0) { P.far_cnt[3 * sgi] = nl; P.far_cnt[3 * sgi + 1] = nf; P.far_cnt[3 * sgi + 2] = nvis; }
    }
}

template <bool COUNT, bool SPH, bool MIXED>
__global__ void __launch_bounds__(WalkCfg<SPH>::TPB, WALK_CTAS) k_walk(const WalkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool CLASSES = MIXED, FULLCLASS = CLASSES;
    constexpr bool INWALK = SPH && !MIXED;      // FP64 mode evaluates SPH pairs inside the walk
    constexpr bool SPLIT = SPH && MIXED;        // mixed mode records the pairs within ~2h; k_sph evaluates them afterwards
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem<SPH, MIXED>& sm = reinterpret_cast<WarpSmem<SPH, MIXED>*>(smem_raw)[warp];
    int2* const spill = P.spill + (size_t)(blockIdx.x * WalkCfg<SPH>::WARPS + warp) * P.spill_per_warp;
    const unsigned lt = (1u << lane) - 1u;
    const int N = (int)P.N;
    const double R = __longlong_as_double((long long)P.s->Rbits);
    const int n_nodes = P.s->n_nodes, n_in_tree = P.s->n_in_tree;
    const double theta = P.theta, theta2 = theta * theta;
    const bool fast_mac = theta > 0.0;
    const double inv_theta2 = fast_mac ? 1.0 / theta2 : 0.0;
    const Slice sl = target_slice(P);
    const unsigned ngroups = (unsigned)((sl.a1 - sl.a0 + 31) / 32);
    // the law is evaluated in units of R so that r^2 (r^2+e0^2)^2 cannot overflow for any unit system
    // ... in units of L = R / 2^16 in fact (a power of two: bit-identical FP64 results): the FP32 law below evaluates
    // rsqrt(r^2 (r^2 + e0^2)^2) with ONE MUFU, and with this unit the argument stays inside the FP32 range for
    // separations down to ~1e-12 R and softening lengths from ~1e-10 R to ~50 R.
    const double invR2 = R > 0.0 ? 4294967296.0 / (R * R) : 1.0;
    const double e02s = P.e0 * P.e0 * invR2;
    const double GR3 = kG * invR2 * sqrt(invR2);
    // mixed precision (MIXED): displacements as float-float differences in units of R about the warp's box centre,
    // the law in FP32 (MUFU.RSQ + MUFU.RCP), FP32 partial sums per 32-source tile, FP64 accumulation across tiles.
    // Masses are scaled by the mean in-tree mass so every FP32 quantity stays well inside the normal range.
    const double invR = sqrt(invR2);
    const double m0 = n_nodes > 0 ? fmax(P.src_pm[N].w / (double)max(n_in_tree, 1), 1e-300) : (n_in_tree > 0 && P.src_pm[0].w > 0.0 ? P.src_pm[0].w : 1.0);
    const double inv_m0 = 1.0 / m0, acc_scale = kG * m0 * invR2;
    const float e02f = (float)e02s;

    // statistics live in registers: keeping them in shared memory instead frees 22 registers (no spills) but costs 4 % on C1
    unsigned long long tot_node = 0, tot_leaf = 0, tot_sph = 0, tot_visit = 0, tot_exact = 0, tot_spill = 0;
    unsigned long long st_rounds = 0, st_popped = 0, st_mixed = 0, st_open = 0, st_drain = 0;
    unsigned long long st_cls[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // COUNT only: list entries / set bits by lane span (any, one half, one quarter), far-list entries

    auto stack_put = [&](int idx, int2 v) {
        if (idx < SCAP) sm.stack[idx] = v;
        else if (idx - SCAP < P.spill_per_warp) spill[idx - SCAP] = v;
        else P.s->walk_overflow = 1;                      // reported as AGB_ERR_NOMEM by agb_forces
    };
    auto stack_get = [&](int idx) -> int2 {
        if (idx < SCAP) return sm.stack[idx];
        if (idx - SCAP < P.spill_per_warp) return spill[idx - SCAP];
        return make_int2(-1, 0);
    };

    if (P.s->node_overflow || P.s->walk_overflow == 3) return;
    for (;;) {
        unsigned g = 0;
        if (lane == 0) g = atomicAdd(&P.s->walk_next_group, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= ngroups) break;
        const int64_t idx = sl.a0 + (int64_t)g * 32 + lane;
        const bool inrange = idx < sl.a1;
        const int64_t t = !inrange ? -1 : sl.ident ? idx : (int64_t)P.act_list[idx];
        const double4 tp = inrange ? P.src_pm[t] : make_double4(0, 0, 0, 0);
        const bool active = inrange;                                            // the list only holds active targets (Tree.cpp:75)
        const bool valid = active && tp.w != 0.0;                               // Node.cpp:265
        // per-target SPH constants (the reference overrides h_j, rho_j, P_j with the target's, Node.cpp:94,101,108)
        bool tgas = false;
        double h_t = 0, hh4 = 0, hh4c = 0;
        if (SPH && valid && P.s_type[t] == 2) {
            tgas = true;
            h_t = P.s_h[t];
            hh4 = 4.0 * h_t * h_t;
            hh4c = hh4 * (1.0 + 1e-13);
            if (INWALK) {
                // the reference overrides h_j, rho_j, P_j with the target's own values (Node.cpp:94,101,108)
                const double rho = P.s_rho[t], Pr = P.s_P[t];
                const double inv_h = 1.0 / h_t, pr2 = Pr / (rho * rho);
                sm.tsph[3 * lane] = make_double4(inv_h, inv_h * inv_h * inv_h * inv_h / kPI, pr2 + pr2, sqrt(kGAMMA * Pr / rho));
                const double4 tv = P.src_gv[t];
                sm.tsph[3 * lane + 1] = make_double4(tv.x, tv.y, tv.z, h_t);
            }
        }
        double ax = 0, ay = 0, az = 0, dU = 0;
        int c_vis = (active && (n_nodes > 0 || !valid)) ? 1 : 0, c_an = 0, c_al = 0, c_sp = 0;              // the root call itself
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const int tmin = __shfl_sync(0xffffffffu, (int)t, 0), tmax = __reduce_max_sync(0xffffffffu, (int)t);   // targets are sorted by tree position
        const unsigned gasl = SPH ? __ballot_sync(0xffffffffu, tgas && h_t > 0.0) : 0u;   // targets of this warp that can feel SPH at all
        const bool wgas = gasl != 0u;
        int rec_last = -1, rec_free = 0, rec_idx = 0;                            // SPLIT: this group's last record, records left of the batch taken from the pool
        int rfill = 0;                                                           // SPLIT: entries waiting in sm.rsrc (< 64)
        auto rec_flush = [&](const bool full) {
            // writes the first 32 waiting entries (full) or all of them (< 32, end of the group) as one record
            if (rfill == 0) return;
            if (rec_free == 0) {
                if (lane == 0) rec_idx = (int)min(atomicAdd(&P.s->cand_cursor, (unsigned long long)REC_BATCH), 0x7ffffff0ull);
                rec_idx = __shfl_sync(0xffffffffu, rec_idx, 0);
                rec_free = REC_BATCH;
            }
            __syncwarp();
            const int n = full ? 32 : rfill;
            const int2 mine_ = lane < n ? sm.rsrc[lane] : make_int2(-1, 0);
            const int2 next_ = (full && 32 + lane < rfill) ? sm.rsrc[32 + lane] : make_int2(-1, 0);
            if ((int64_t)rec_idx < P.rec_cap) {
                P.rec_ent[(size_t)rec_idx * 32 + lane] = mine_;
                if (lane == 0) P.rec_next[rec_idx] = rec_last;
                rec_last = rec_idx;
            } else if (lane == 0) P.s->walk_overflow = 2;
            rec_idx++; rec_free--;
            __syncwarp();
            rfill -= n;
            if (lane < rfill) sm.rsrc[lane] = next_;
            __syncwarp();
        };
        if (vmask) {
            const double inf = __longlong_as_double(0x7ff0000000000000ll);
            const double lox = warp_min(valid ? tp.x : inf), loy = warp_min(valid ? tp.y : inf), loz = warp_min(valid ? tp.z : inf);
            const double hix = warp_max(valid ? tp.x : -inf), hiy = warp_max(valid ? tp.y : -inf), hiz = warp_max(valid ? tp.z : -inf);
            const double cgx = 0.5 * (lox + hix), cgy = 0.5 * (loy + hiy), cgz = 0.5 * (loz + hiz);
            float2 nth_x = make_float2(0.f, 0.f), nth_y = nth_x, nth_z = nth_x, ntl_x = nth_x, ntl_y = nth_x, ntl_z = nth_x;
            float hh4cf = 0, hh4sf = 0, far2 = 0;
            // the target's position about the centre of the warp's box in units of L, as a float-float pair
            const double rx = (tp.x - cgx) * invR, ry = (tp.y - cgy) * invR, rz = (tp.z - cgz) * invR;
            const float thx = (float)rx, thy = (float)ry, thz = (float)rz;
            const float ntx = -thx, nty = -thy, ntz = -thz;
            // half extents of the warp's box (rounded up); FP32 opening tests need |d| > half-diagonal / 16
            const float bhx = (float)(0.5 * (hix - lox) * invR) * (1.0f + 1e-6f), bhy = (float)(0.5 * (hiy - loy) * invR) * (1.0f + 1e-6f),
                        bhz = (float)(0.5 * (hiz - loz) * invR) * (1.0f + 1e-6f);
            const float hd2 = bhx * bhx + bhy * bhy + bhz * bhz, guard2 = hd2 * (1.0f / 256.0f);
            if (MIXED) {
                // minus the target's own float-float position, duplicated into both halves of a packed register
                const float tlx = (float)(rx - (double)thx), tly = (float)(ry - (double)thy), tlz = (float)(rz - (double)thz);
                nth_x = make_float2(-thx, -thx); nth_y = make_float2(-thy, -thy); nth_z = make_float2(-thz, -thz);
                ntl_x = make_float2(-tlx, -tlx); ntl_y = make_float2(-tly, -tly); ntl_z = make_float2(-tlz, -tlz);
                hh4sf = (float)(hh4 * invR2);
                far2 = P.far_k2 * hd2;                                      // the squared "far" distance: half a half-diagonal of the box
                hh4cf = hh4sf * (1.0f + 1e-5f);
            }
            // SPLIT: a source can only matter to SPH if it lies within 2 h_max of the warp's box (h = 2 radius of a cell, so 2h / L
            // is exact in FP32); the margins cover the FP32 rounding of the source position and of the box
            float cand2 = 0.f;
            if (SPLIT && wgas) {
                const float hm = sqrtf(warp_max_f((gasl >> lane) & 1u ? hh4sf : 0.f)) * (1.0f + 1e-4f) + 1e-5f * sqrtf(hd2);
                cand2 = hm * hm;
            }
            const Box wb{lox, loy, loz, hix, hiy, hiz};
            // start from what the far-field prepass left for this super-group: a shared accept list and a frontier
            const int64_t sgi = (int64_t)g / SG_GROUPS;
            const int32_t* const fl = P.far_list + (size_t)sgi * FAR_LCAP;
            const int32_t* const ff = P.far_front + (size_t)sgi * FAR_FCAP;
            const int n_fl = P.far_cnt[3 * sgi], n_ff = P.far_cnt[3 * sgi + 1];
            if (COUNT && valid) c_vis += P.far_cnt[3 * sgi + 2];
            if (COUNT) st_cls[6] += n_fl;
            int sp = n_ff, lc = 0, cpos = 0, npend = 0;
            sm.tpos[lane] = make_float4(ntx, nty, ntz, __int_as_float((int)t));
            for (int i = lane; i < n_ff; i += 32) stack_put(i, make_int2(ff[i], (int)vmask));
            __syncwarp();

            while (true) {
                if (cpos < n_fl) {
                    // ------------------------------------------------ import a chunk of the shared far-field list (full mask)
                    const int nimp = min(n_fl - cpos, LCAP);
                    for (int i = lane; i < nimp; i += 32) sm.list[i] = make_int2(fl[cpos + i], (int)vmask);
                    lc = nimp; cpos += nimp;
                    __syncwarp();
                } else
                // ------------------------------------------------ traversal: fill the interaction list
                // Two kinds of rounds feed one common "append / push" phase.  POP: every lane pops one (node, mask) entry and
                // classifies it against the warp's box; nodes the whole mask accepts or opens are finished, straddling nodes are
                // parked in sm.pend.  RESOLVE (32 nodes parked, or nothing left to pop): one parked node per LANE, loop over the
                // 32 TARGETS (their positions are broadcast from shared memory): the accept / open bits of a node accumulate in
                // its own lane's registers — no ballots, no shuffles, 32 (node, target) tests per ~16 instructions.
                while ((sp > 0 || npend > 0) && lc <= LCAP - LGROW) {
                    int2 e = make_int2(-1, 0);
                    unsigned amask = 0u, omask = 0u;                         // lanes of the entry's mask that accept the node / open it
                    int4 c0 = make_int4(-1, -1, -1, -1), c1 = c0;
                    if (npend < 32 && sp > 0) {
                        const int cnt = min(sp, 32);
                        sp -= cnt;
                        if (lane < cnt) e = stack_get(sp + lane);
                        __syncwarp();                                        // the slots just read are overwritten by this round's pushes
                        if (sp + cnt > SCAP) tot_spill += 1;
                        int outcome = OUT_NONE;
                        // the child slots are requested together with the node record (two independent L2 round trips instead
                        // of two dependent ones); they are simply not used when the node turns out to be accepted
                        if (lane < cnt && e.x >= 0) {
                            c0 = __ldg(reinterpret_cast<const int4*>(P.child) + 2 * (size_t)(e.x - N));
                            c1 = __ldg(reinterpret_cast<const int4*>(P.child) + 2 * (size_t)(e.x - N) + 1);
                            const double4 pm = P.src_pm[e.x];
                            if (pm.w != 0.0) {                                   // Node.cpp:250 / :390
                                const double rad = scalbn(R, -(int)P.ndepth[e.x - N]);
                                outcome = classify_box(pm, rad * rad, wb, theta2, fast_mac);
                            }
                        }
                        if (COUNT) {
                            const unsigned vm = (outcome != OUT_NONE && e.x != N) ? (unsigned)e.y : 0u;
                            for (int j = 0; j < cnt; j++) c_vis += (__shfl_sync(0xffffffffu, vm, j) >> lane) & 1u;
                        }
                        if (outcome == OUT_ACCEPT) amask = (unsigned)e.y;
                        if (outcome == OUT_OPEN) omask = (unsigned)e.y;
                        const unsigned mm = __ballot_sync(0xffffffffu, outcome == OUT_MIXED);
                        if (outcome == OUT_MIXED) sm.pend[npend + __popc(mm & lt)] = e;
                        npend += __popc(mm);
                        if (lane == 0) { st_rounds++; st_popped += cnt; st_mixed += __popc(mm); }
                    } else {
                        // ---- resolve up to 32 parked nodes (the most recent ones)
                        const int nb = min(npend, 32);
                        npend -= nb;
                        float nx = 0.f, ny = 0.f, nz = 0.f, thi = 0.f, tlo = 0.f;
                        unsigned my = 0u;
                        if (lane < nb) {
                            e = sm.pend[npend + lane];
                            my = (unsigned)e.y;
                            c0 = __ldg(reinterpret_cast<const int4*>(P.child) + 2 * (size_t)(e.x - N));
                            c1 = __ldg(reinterpret_cast<const int4*>(P.child) + 2 * (size_t)(e.x - N) + 1);
                            // Per-target tests in FP32 first: the node's position about the box centre (units of L) and two thresholds on
                            // r^2 that bracket radius^2 / theta^2 by +-4e-5.  With |d| > half-diagonal / 16 (the guard) the FP32 r^2 is
                            // within 7e-6 of the true one, so an FP32 verdict outside the bracket is the FP64 verdict; everything else
                            // (inside the bracket, too close to the node, theta <= 0) is re-decided in FP64.
                            const double4 pm = P.src_pm[e.x];
                            const double rad = scalbn(R, -(int)P.ndepth[e.x - N]);
                            const double thr = rad * rad * invR2 * inv_theta2;
                            nx = (float)((pm.x - cgx) * invR); ny = (float)((pm.y - cgy) * invR); nz = (float)((pm.z - cgz) * invR);
                            thi = fmaxf((float)(thr * (1.0 + 4e-5)), guard2);
                            tlo = (float)(thr * (1.0 - 4e-5));
                        }
                        __syncwarp();
                        auto exact_mac = [&](const int node, const int tidx, bool& acc, bool& open) {
                            const double4 q = P.src_pm[node], tq = P.src_pm[tidx];
                            const double rad = scalbn(R, -(int)P.ndepth[node - N]), qw = rad * rad;
                            const double dx = q.x - tq.x, dy = q.y - tq.y, dz = q.z - tq.z;
                            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx)), lhs = r2 * theta2;
                            acc = false; open = false;
                            if (r2 != 0.0) {                                         // Node.cpp:274 (r == 0 -> return)
                                if (fast_mac && lhs > qw * (1.0 + 1e-13)) acc = true;
                                else if (fast_mac && lhs < qw * (1.0 - 1e-13)) open = true;
                                else {                                               // the reference's own expression, Node.cpp:271,331-334
                                    const double r2e = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                                    acc = __ddiv_rn(__dsqrt_rn(qw), __dsqrt_rn(r2e)) < theta;      // sqrt(radius^2) is exact: radius = R 2^-k
                                    open = !acc;
                                    tot_exact++;
                                }
                            }
                        };
                        unsigned amb = 0u;
#pragma unroll 4
                        for (int t = 0; t < 32; t++) {
                            const float4 tq = sm.tpos[t];                            // broadcast
                            const float dx = nx + tq.x, dy = ny + tq.y, dz = nz + tq.z;
                            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            const bool in = (my >> t) & 1u;
                            const bool acc = in && fast_mac && r2 > thi, open = in && fast_mac && r2 < tlo && r2 > guard2;
                            amask |= acc ? 1u << t : 0u;
                            omask |= open ? 1u << t : 0u;
                            amb |= (in && !acc && !open) ? 1u << t : 0u;
                        }
                        while (amb) {                                                // rare: decided in FP64 (one copy of the code, lanes diverge)
                            const int t = __ffs(amb) - 1;
                            amb &= amb - 1;
                            bool acc, open;
                            exact_mac(e.x, __float_as_int(sm.tpos[t].w), acc, open);
                            amask |= acc ? 1u << t : 0u;
                            omask |= open ? 1u << t : 0u;
                        }
                    }
                    // one list entry per node with acceptors
                    const unsigned am = __ballot_sync(0xffffffffu, amask != 0u);
                    if (amask != 0u) sm.list[lc + __popc(am & lt)] = make_int2(e.x, (int)amask);
                    lc += __popc(am);
                    // children of every node with openers inherit the openers' mask: leaves -> list, nodes -> stack
                    const unsigned om = __ballot_sync(0xffffffffu, omask != 0u);
                    if (om) {
                        const int ch[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                        int nl = 0, nn = 0;
                        if (omask != 0u) {
#pragma unroll
                            for (int c = 0; c < 8; c++) { nl += (ch[c] >= 0 && ch[c] < N); nn += (ch[c] >= N); }
                        }
                        // both counts are <= 8 per lane: one scan of the packed pair (leaves in the low half, nodes in the high half)
                        const int sc = warp_incl_scan(nl | (nn << 16), lane), st_ = __shfl_sync(0xffffffffu, sc, 31);
                        const int il = sc & 0xffff, in_ = sc >> 16, tl = st_ & 0xffff, tn = st_ >> 16;
                        if (omask != 0u) {
                            int pl = lc + il - nl, pn = sp + in_ - nn;
#pragma unroll
                            for (int c = 0; c < 8; c++) {
                                if (ch[c] >= N) stack_put(pn++, make_int2(ch[c], (int)omask));
                                else if (ch[c] >= 0) sm.list[pl++] = make_int2(ch[c], (int)omask);
                            }
                        }
                        lc += tl; sp += tn;
                        if (lane == 0) st_open += __popc(om);
                    }
                    __syncwarp();
                    // warm L1 with the node records the next round will pop
                    if (lane < sp && lane < 32) {
                        const int2 nx_ = stack_get(sp - 1 - lane);
                        if (nx_.x >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.src_pm + nx_.x));
                    }
                }

                // ------------------------------------------------ drain the interaction list
                if (lane == 0) st_drain += (lc + 31) / 32;
                for (int base = 0; base < lc; base += 32) {
                    const int cnt = min(32, lc - base);
                    __syncwarp();
                    bool src_gas = false;
                    unsigned pc_mask = 0u; int ex_part = -1;
                    // Mixed mode sorts the tile's entries into three classes, evaluated by three loops of decreasing speed:
                    //   0  far + accepted by every target: single-float displacements, no mask
                    //   1  far: single-float displacements
                    //   2  near: float-float displacements (close pairs cancel)
                    // "far" = at least half a half-diagonal of the warp's box away from the box: dropping the low parts then costs
                    // <= 2^-24 (|s| + |t|) / |d| <= 3e-7 relative per pair (measured: 2, 1, 1/2, 1/4.5 half-diagonals give 1.410, 1.406,
                    // 1.392, 1.379 ms on C1; median / p99 error 4.35e-8 / 9.6e-7 at 2 and 4.45e-8 / 1.0e-6 at 1/2).
                    int pos = lane, cls = 3;
                    int2 e = make_int2(0, 0);
                    if (lane < cnt) e = sm.list[base + lane];
                    if (CLASSES) __syncwarp();                                      // the tile's list entries are rewritten in class order below
                    double4 q = make_double4(0, 0, 0, 0), gvv = q;
                    float hx = 0, hy = 0, hz = 0, lx = 0, ly = 0, lz = 0, dmin2 = 0;
                    if (lane < cnt) {
                        q = P.src_pm[e.x];
                        if (SPLIT && wgas) src_gas = P.src_flag[e.x] != 0;          // requested together with the record (the candidate test below needs it)
                        if (INWALK && wgas) { gvv = P.src_gv[e.x]; src_gas = gvv.w > 0.0; }   // independent of the load above
                        if (MIXED) {
                            const double rx = (q.x - cgx) * invR, ry = (q.y - cgy) * invR, rz = (q.z - cgz) * invR;
                            hx = (float)rx; hy = (float)ry; hz = (float)rz;
                            lx = (float)(rx - (double)hx); ly = (float)(ry - (double)hy); lz = (float)(rz - (double)hz);
                            const float fx_ = fmaxf(0.f, fabsf(hx) - bhx), fy_ = fmaxf(0.f, fabsf(hy) - bhy), fz_ = fmaxf(0.f, fabsf(hz) - bhz);
                            dmin2 = fmaf(fz_, fz_, fmaf(fy_, fy_, fx_ * fx_));
                            cls = dmin2 < far2 ? 2 : (FULLCLASS && (unsigned)e.y == vmask) ? 0 : 1;
                        }
                        // accepted pairs of this entry: lanes in the mask, minus the target's own leaf, none for a massless leaf
                        pc_mask = q.w != 0.0 ? (unsigned)e.y : 0u;
                        ex_part = e.x < N ? e.x : -1;
                    }
                    int b0 = 0, b1 = 0;                                             // class loops: [0, b0) [b0, b1) [b1, cnt)
                    if (CLASSES) {
                        const unsigned m0_ = __ballot_sync(0xffffffffu, cls == 0), m1_ = __ballot_sync(0xffffffffu, cls == 1), m2_ = __ballot_sync(0xffffffffu, cls == 2);
                        const int n0 = __popc(m0_), n1 = __popc(m1_);
                        pos = cls == 0 ? __popc(m0_ & lt) : cls == 1 ? n0 + __popc(m1_ & lt) : n0 + n1 + __popc(m2_ & lt);
                        b0 = n0 & ~1; b1 = (n0 + n1) & ~1;                          // a pair that straddles two classes runs in the slower one
                        if (lane < cnt) sm.list[base + pos] = e;
                    }
                    if (lane < cnt) {
                        if (MIXED) {
                            // two sources share one 64-byte record, components interleaved (x0 x1 y0 y1 | z0 z1 m0 m1 | lo parts),
                            // so the evaluation loop can use Blackwell's packed FP32x2 instructions across the pair
                            float* sf = reinterpret_cast<float*>(&sm.stage[pos & ~1]) + (pos & 1);
                            sf[0] = hx; sf[2] = hy; sf[4] = hz; sf[6] = (float)(q.w * inv_m0);
                            sf[8] = lx; sf[10] = ly; sf[12] = lz;
                        } else sm.stage[lane] = q;
                        if (INWALK && wgas) sm.gst[pos] = gvv;
                    }
                    {   // a leaf that is one of this warp's own targets does not count as an interaction with itself
                        const bool maybe = ex_part >= tmin && ex_part <= tmax;
                        if (sl.ident) { if (maybe) pc_mask &= ~(1u << (ex_part - tmin)); }        // every particle active: lane = offset in the group
                        else if (__any_sync(0xffffffffu, maybe))
                            for (int l2 = 0; l2 < 32; l2++) { const int tl = __shfl_sync(0xffffffffu, (int)t, l2); if (ex_part >= 0 && ex_part == tl) pc_mask &= ~(1u << l2); }
                        if (ex_part >= 0) tot_leaf += __popc(pc_mask); else tot_node += __popc(pc_mask);
                    }
                    if (SPLIT && wgas) {
                        // sources within 2 h_max of the warp's box that a gas target accepted: (source, those targets) for k_sph
                        const unsigned gm = pc_mask & gasl;
                        const bool cand = gm != 0u && dmin2 < cand2 && src_gas;     // sources without gas never pass Node.cpp:319 / :371
                        const unsigned cm = __ballot_sync(0xffffffffu, cand);
                        if (cm) {
                            if (cand) sm.rsrc[rfill + __popc(cm & lt)] = make_int2(e.x, (int)gm);
                            rfill += __popc(cm);
                            if (rfill >= 32) rec_flush(true);
                        }
                    }
                    if (COUNT) {
                        // tuning statistics: how many list entries have their acceptors inside one half / one quarter of the warp
                        const unsigned m = pc_mask;
                        int span = -1;
                        if (m) {
                            const bool qt = !(m & ~0xffu) || !(m & ~0xff00u) || !(m & ~0xff0000u) || !(m & ~0xff000000u);
                            const bool hf = !(m & 0xffff0000u) || !(m & 0xffffu);
                            span = qt ? 2 : hf ? 1 : 0;
                        }
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            st_cls[c] += __popc(__ballot_sync(0xffffffffu, span == c));
                            st_cls[3 + c] += __reduce_add_sync(0xffffffffu, span == c ? __popc(m) : 0);
                            st_cls[7 + c] += __popc(__ballot_sync(0xffffffffu, cls == c));      // evaluation classes (mixed mode)
                        }
                    }
                    if (MIXED && lane == cnt && (cnt & 1)) {
                        // odd tile: the unused half of the last pair must hold finite numbers (0 * inf would poison the sums)
                        float* sf = reinterpret_cast<float*>(&sm.stage[lane & ~1]) + (lane & 1);
                        sf[0] = 1.f; sf[2] = 1.f; sf[4] = 1.f; sf[6] = 0.f; sf[8] = 0.f; sf[10] = 0.f; sf[12] = 0.f;
                    }
                    const unsigned gasmask = (INWALK && wgas) ? __reduce_or_sync(0xffffffffu, src_gas ? 1u << pos : 0u) : 0u;   // tile positions that hold gas
                    if (base + 32 + lane < lc) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.src_pm + sm.list[base + 32 + lane].x));
                    __syncwarp();
                    unsigned gate = 0;                                                   // per lane: entries within ~2h (SPH candidates)
                    if (MIXED) {
                        float2 fax = make_float2(0.f, 0.f), fay = fax, faz = fax;
                        // the law m d / (r (r^2 + e0^2)) with one MUFU per source: rsqrt(r^2 (r^2 + e0^2)^2).  Skip rules as in the
                        // FP64 loop: d = 0 exactly for the own leaf / coincident sources, and the 1e-37 floor keeps the factor
                        // finite so that f * d = 0; a massless source has m = 0.
                        const float2 tiny2 = make_float2(1e-37f, 1e-37f), e022 = make_float2(e02f, e02f);
                        auto pair_loop = [&](auto lo_, auto masked_, const int j0, const int j1) {
                            constexpr bool LO = decltype(lo_)::value, MASKED = decltype(masked_)::value;
#pragma unroll 2
                            for (int j = j0; j < j1; j += 2) {
                                const float4* rec = reinterpret_cast<const float4*>(&sm.stage[j]);
                                const float4 r0 = rec[0], r1 = rec[1];
                                float2 dx = __fadd2_rn(make_float2(r0.x, r0.y), nth_x), dy = __fadd2_rn(make_float2(r0.z, r0.w), nth_y), dz = __fadd2_rn(make_float2(r1.x, r1.y), nth_z);
                                if (LO) {                                          // float-float displacement for both sources at once (FADD2)
                                    const float4 r2_ = rec[2]; const float2 r3 = *reinterpret_cast<const float2*>(rec + 3);
                                    dx = __fadd2_rn(dx, __fadd2_rn(make_float2(r2_.x, r2_.y), ntl_x));
                                    dy = __fadd2_rn(dy, __fadd2_rn(make_float2(r2_.z, r2_.w), ntl_y));
                                    dz = __fadd2_rn(dz, __fadd2_rn(r3, ntl_z));
                                }
                                const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                                const float2 q2 = __fadd2_rn(r2, e022);
                                const float2 xx = __ffma2_rn(__fmul2_rn(r2, q2), q2, tiny2);
                                float2 w;
                                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(w.x) : "f"(xx.x));
                                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(w.y) : "f"(xx.y));
                                float2 f = __fmul2_rn(make_float2(r1.z, r1.w), w);
                                bool bit0 = true, bit1 = true;
                                int4 ee = make_int4(0, 0, 0, 0);
                                if (MASKED || COUNT) {                             // counter mode: same arithmetic, plus the bookkeeping
                                    ee = *reinterpret_cast<const int4*>(&sm.list[base + j]);      // entries j, j+1: (src, mask) x 2
                                    bit0 = ((unsigned)ee.y >> lane) & 1u; bit1 = (j + 1 < cnt) && (((unsigned)ee.w >> lane) & 1u);
                                }
                                if (MASKED) { f.x = bit0 ? f.x : 0.f; f.y = bit1 ? f.y : 0.f; }
                                fax = __ffma2_rn(f, dx, fax); fay = __ffma2_rn(f, dy, fay); faz = __ffma2_rn(f, dz, faz);
                                if (COUNT) {
                                    const bool seen0 = bit0 && r1.z != 0.f, ok0 = seen0 && r2.x != 0.f;
                                    const bool seen1 = bit1 && r1.w != 0.f, ok1 = seen1 && r2.y != 0.f;
                                    if (ee.x < N) { c_vis += seen0; c_al += ok0; } else c_an += ok0;
                                    if (ee.z < N) { c_vis += seen1; c_al += ok1; } else c_an += ok1;
                                }
                            }
                        };
                        if (FULLCLASS) pair_loop(std::false_type{}, std::false_type{}, 0, b0);
                        if (CLASSES) pair_loop(std::false_type{}, std::true_type{}, b0, b1);
                        pair_loop(std::true_type{}, std::true_type{}, b1, cnt);
                        ax = fma(acc_scale, (double)(fax.x + fax.y), ax); ay = fma(acc_scale, (double)(fay.x + fay.y), ay); az = fma(acc_scale, (double)(faz.x + faz.y), az);
                    } else {
#pragma unroll 4
                    for (int j = 0; j < cnt; j++) {
                        const int2 e = sm.list[base + j];
                        const double4 q = sm.stage[j];
                        const bool bit = ((unsigned)e.y >> lane) & 1u;
                        const double dx = q.x - tp.x, dy = q.y - tp.y, dz = q.z - tp.z;
                        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                        // No select is needed for the reference's skip rules: a massless leaf (Node.cpp:390) has q.w = 0;
                        // the target's own leaf (Node.cpp:260) and any coincident source (r == 0, Node.cpp:274) have
                        // d = 0 exactly, so f * d = 0 as long as f stays finite, which the 1e-300 guard ensures.
                        const double r2s = r2 * invR2, q2 = r2s + e02s;
                        const double w = rsqrt_pos(fma(r2s * q2, q2, 1e-300));   // 1 / (r (r^2 + e0^2)) in units of R
                        const double f = (bit ? GR3 * q.w : 0.0) * w;
                        ax = fma(f, dx, ax); ay = fma(f, dy, ay); az = fma(f, dz, az);
                        if (INWALK && wgas) gate |= (r2 < hh4c ? 1u : 0u) << j;            // hh4c = 0 for non-gas targets: never set
                        if (COUNT) {
                            const bool seen = bit && q.w != 0.0;
                            const bool ok = seen && r2 != 0.0;
                            if (e.x < N) { c_vis += seen; c_al += ok; } else c_an += ok;
                        }
                    }
                    }
                    if (INWALK && wgas && __any_sync(0xffffffffu, (gate & gasmask) != 0u)) {
                        // second pass over the few (target, source) pairs that can pass r < 2 h_i (Node.cpp:316-325, :368-377)
                        unsigned mine = gate & gasmask;
                        if (!MIXED)
                        // FP64 mode: every gas lane runs through ITS OWN candidates (a per-lane loop: lanes only diverge in trip count)
                        while (mine) {
                            const int j = __ffs(mine) - 1;
                            mine &= mine - 1;
                            const int2 e = sm.list[base + j];
                            if (!(((unsigned)e.y >> lane) & 1u)) continue;
                            const double4 gv = sm.gst[j];                    // (mVel | particle velocity, gasMass)
                            const double4 k4 = sm.tsph[3 * lane], tv = sm.tsph[3 * lane + 1];
                            const double A2 = k4.z;
                            {
                                const double4 q = sm.stage[j];
                                const double dx = q.x - tp.x, dy = q.y - tp.y, dz = q.z - tp.z;
                                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                                bool pass = q.w != 0.0 && r2 != 0.0 && r2 < hh4c;
                                if (pass && !(r2 < hh4 * (1.0 - 1e-13))) {
                                    // within 1e-13 of the gate: the reference's own separately rounded expression
                                    const double r2e = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                                    pass = __dsqrt_rn(r2e) < __dmul_rn(h_t, 2.0);
                                    tot_exact++;
                                }
                                if (pass) {
                                    const double inv_h = k4.x, inv_pi_h4 = k4.y, cs = k4.w;
                                    const double rinv = rsqrt_pos(r2), r = r2 * rinv, qq = r * inv_h;
                                    double gs = 0.0;                             // kernel.cpp:28-34
                                    if (qq < 1.0) gs = -3.0 * qq + 2.25 * qq * qq;
                                    else if (qq < 2.0) { const double u = 2.0 - qq; gs = -0.75 * u * u; }
                                    const double gfac = gs * inv_pi_h4 * rinv;
                                    const double sx = -dx, sy = -dy, sz = -dz;   // d = x_i - COM (Node.cpp:116)
                                    const double gx = sx * gfac, gy = sy * gfac, gz = sz * gfac;
                                    const double vx = tv.x - gv.x, vy = tv.y - gv.y, vz = tv.z - gv.z;
                                    const double vd = vx * sx + vy * sy + vz * sz;
                                    const double mu = h_t * vd / (r2 + 0.01 * (h_t * h_t));
                                    const double MU = vd < 0.0 ? (-0.5 * cs * mu + mu * mu) : 0.0;     // Node.cpp:142-152
                                    const double coef = -gv.w * (A2 + MU);                             // Node.cpp:127 + :154
                                    const double fx = coef * gx, fy = coef * gy, fz = coef * gz;
                                    dU += 0.5 * gv.w * (A2 + MU) * (vx * gx + vy * gy + vz * gz);      // Node.cpp:167
                                    if (!(isnan(fx) || isnan(fy) || isnan(fz))) { ax += fx; ay += fy; az += fz; }   // Node.cpp:169
                                    tot_sph++;
                                    if (COUNT) c_sp++;
                                }
                            }
                        }
                    }
                }
                lc = 0;
                __syncwarp();                                                // the list is rewritten by the next import / round
                if (sp == 0 && npend == 0 && cpos >= n_fl) break;
            }
        }
        if (SPLIT) { rec_flush(false); if (lane == 0) P.rec_head[g] = rec_last; }
        if (active) {
            if (!valid) { ax = 0; ay = 0; az = 0; }                            // Node.cpp:265; such lanes are not masked out in the class-0 loop
            const uint32_t p = P.perm[t];
            P.ax[p] = ax; P.ay[p] = ay; P.az[p] = az;                           // Tree.cpp:77 (acc = 0) + accumulated force
            if (INWALK && tgas && dU != 0.0) P.dUdt[p] += dU;
            if (COUNT) { P.c_visits[t] = c_vis; P.c_accn[t] = c_an; P.c_accl[t] = c_al; P.c_sph[t] = c_sp; }
            tot_visit += (unsigned long long)c_vis;
        }
    }

    tot_node = warp_sum_u64(tot_node); tot_leaf = warp_sum_u64(tot_leaf); tot_sph = warp_sum_u64(tot_sph);
    tot_visit = warp_sum_u64(tot_visit); tot_exact = warp_sum_u64(tot_exact); tot_spill = warp_sum_u64(tot_spill);
    if (lane == 0) {
        if (tot_node) atomicAdd(&P.s->c_node, tot_node);
        if (tot_leaf) atomicAdd(&P.s->c_leaf, tot_leaf);
        if (tot_sph) atomicAdd(&P.s->c_sph, tot_sph);
        if (tot_visit) atomicAdd(&P.s->c_visits, tot_visit);
        if (tot_exact) atomicAdd(&P.s->c_exact, tot_exact);
        if (tot_spill) atomicAdd(&P.s->c_spill, tot_spill);
        atomicAdd(&P.s->st_rounds, st_rounds); atomicAdd(&P.s->st_popped, st_popped); atomicAdd(&P.s->st_mixed, st_mixed);
        atomicAdd(&P.s->st_open, st_open); atomicAdd(&P.s->st_drain, st_drain);
        if (COUNT) for (int c = 0; c < 10; c++) atomicAdd(&P.s->st_cls[c], st_cls[c]);
    }
}

// ---- SPH pairs of the mixed-precision mode (Node::calcSPHForce, Node.cpp:88-172; the gate of Node.cpp:316-325, :368-377).
// k_walk<SPH, MIXED> records, per group of 32 targets, the accepted gas-bearing sources that lie within 2 h_max of the
// group's box together with the gas targets that accepted them.  Here one warp per group takes those entries in tiles of
// 32, enumerates the (target, source) pairs in (target, source) order and hands them out one per lane: the exact gate
// r < 2 h_i (FP32, FP64 when within 4e-6 of the threshold), the kernel gradient / viscosity algebra in FP32, the final
// products in FP64; results travel through shared memory back to the owning lane, which adds its own pairs in source
// order (deterministic: bit-identical from run to run and for any number of GPUs).
#ifndef AGB_SPH_CTAS_PER_SM
#define AGB_SPH_CTAS_PER_SM 3
#endif
template <bool COUNT>
__global__ void __launch_bounds__(256, AGB_SPH_CTAS_PER_SM) k_sph(const WalkParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SphWarp& sm = reinterpret_cast<SphWarp*>(smem_raw)[warp];
    const double R = __longlong_as_double((long long)P.s->Rbits);
    const double invR2 = R > 0.0 ? 4294967296.0 / (R * R) : 1.0, invR = sqrt(invR2), invR4 = invR2 * invR2;   // same unit L = R / 2^16 as k_walk
    const Slice sl = target_slice(P);
    const unsigned ngroups = (unsigned)((sl.a1 - sl.a0 + 31) / 32);
    unsigned long long tot_sph = 0, tot_exact = 0;
    if (P.s->walk_overflow || P.s->node_overflow) return;                       // incomplete records: agb_forces grows the pool and walks again
    for (unsigned g = blockIdx.x * 8 + warp; g < ngroups; g += gridDim.x * 8) {
        const int head = P.rec_head[g];
        if (head < 0) continue;
        const int64_t idx = sl.a0 + (int64_t)g * 32 + lane;
        const bool inrange = idx < sl.a1;
        const int64_t t = !inrange ? -1 : sl.ident ? idx : (int64_t)P.act_list[idx];
        const double4 tp = inrange ? P.src_pm[t] : make_double4(0, 0, 0, 0);
        const bool valid = inrange && tp.w != 0.0;                              // Node.cpp:265
        bool tgas = false;
        double h_t = 0;
        __syncwarp();
        if (valid && P.s_type[t] == 2) {
            h_t = P.s_h[t];
            tgas = h_t > 0.0;
            if (tgas) {
                // the reference overrides h_j, rho_j, P_j with the target's own values (Node.cpp:94,101,108)
                const double rho = P.s_rho[t], Pr = P.s_P[t];
                const double inv_h = 1.0 / h_t, pr2 = Pr / (rho * rho);
                sm.tsph[3 * lane] = make_double4(inv_h, inv_h * inv_h * inv_h * inv_h / kPI, pr2 + pr2, sqrt(kGAMMA * Pr / rho));
                const double4 tv = P.src_gv[t];
                sm.tsph[3 * lane + 1] = make_double4(tv.x, tv.y, tv.z, h_t);
            }
        }
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        const double lox = warp_min(valid ? tp.x : inf), loy = warp_min(valid ? tp.y : inf), loz = warp_min(valid ? tp.z : inf);
        const double hix = warp_max(valid ? tp.x : -inf), hiy = warp_max(valid ? tp.y : -inf), hiz = warp_max(valid ? tp.z : -inf);
        const double cgx = 0.5 * (lox + hix), cgy = 0.5 * (loy + hiy), cgz = 0.5 * (loz + hiz);
        float2 nth_x = make_float2(0.f, 0.f), nth_y = nth_x, nth_z = nth_x;
        float hh4pf = 0.f;                                                       // prefilter threshold; 0 for everything but gas targets with h > 0
        if (tgas) {
            const double rx = (tp.x - cgx) * invR, ry = (tp.y - cgy) * invR, rz = (tp.z - cgz) * invR;
            const float thx = (float)rx, thy = (float)ry, thz = (float)rz;
            nth_x = make_float2(-thx, -thx); nth_y = make_float2(-thy, -thy); nth_z = make_float2(-thz, -thz);
            float* tf = reinterpret_cast<float*>(&sm.tsph[3 * lane + 2]);
            tf[0] = -thx; tf[1] = -thy; tf[2] = -thz;
            tf[3] = -(float)(rx - (double)thx); tf[4] = -(float)(ry - (double)thy); tf[5] = -(float)(rz - (double)thz);
            tf[6] = (float)(4.0 * h_t * h_t * invR2); tf[7] = __int_as_float((int)t);
            // single-float positions are off by <= 2^-24 (|s| + |t|), |s|, |t| <~ half-diagonal + 2h: pad 2h accordingly
            const float hd = (float)(0.5 * sqrt((hix - lox) * (hix - lox) + (hiy - loy) * (hiy - loy) + (hiz - loz) * (hiz - loz)) * invR);
            const float hp = sqrtf(tf[6]) * (1.0f + 1e-5f) + 1e-6f * (hd + sqrtf(tf[6]));
            hh4pf = hp * hp;
        }
        double ax = 0, ay = 0, az = 0, dU = 0;
        int c_sp = 0;
        int2 ent_next = P.rec_ent[(size_t)head * 32 + lane];
        for (int rec = head; rec >= 0;) {
            {
                __syncwarp();
                // the next record's entries (and its sources) are requested before this one is worked on
                const int2 ent = ent_next;
                rec = P.rec_next[rec];
                if (rec >= 0) {
                    ent_next = P.rec_ent[(size_t)rec * 32 + lane];
                    if (ent_next.x >= 0) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.src_pm + ent_next.x));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.src_gv + ent_next.x));
                    }
                }
                const int src = ent.x;
                const int cnt = __popc(__ballot_sync(0xffffffffu, src >= 0));
                if (lane < cnt) {
                    const double4 q = P.src_pm[src], gv = P.src_gv[src];
                    const double rx = (q.x - cgx) * invR, ry = (q.y - cgy) * invR, rz = (q.z - cgz) * invR;
                    const float hx = (float)rx, hy = (float)ry, hz = (float)rz;
                    float* sf = reinterpret_cast<float*>(&sm.stage[lane & ~1]) + (lane & 1);
                    sf[0] = hx; sf[2] = hy; sf[4] = hz; sf[6] = (q.w != 0.0 && gv.w > 0.0) ? 1.f : 0.f;   // Node.cpp:250,:319,:371
                    sf[8] = (float)(rx - (double)hx); sf[10] = (float)(ry - (double)hy); sf[12] = (float)(rz - (double)hz);
                    sm.gst[lane] = gv; sm.list[lane] = src;
                    sm.lmask[lane] = (q.w != 0.0 && gv.w > 0.0) ? (unsigned)ent.y : 0u;
                } else {
                    if (lane == cnt) {                                           // odd tile: finite numbers in the unused half of the last pair
                        float* sf = reinterpret_cast<float*>(&sm.stage[lane & ~1]) + (lane & 1);
                        sf[0] = 1.f; sf[2] = 1.f; sf[4] = 1.f; sf[6] = 0.f; sf[8] = 0.f; sf[10] = 0.f; sf[12] = 0.f;
                    }
                    sm.lmask[lane] = 0u;
                }
                __syncwarp();
                // my target's view of the record: bit j = entry j holds gas, was accepted by me and lies within ~2h (single-float
                // distance with a generous margin: the pass below decides); two entries per iteration, packed FP32x2
                unsigned mine = 0u;
                for (int j = 0; j < cnt; j += 2) {
                    const float4* rec4 = reinterpret_cast<const float4*>(&sm.stage[j]);
                    const float4 r0 = rec4[0], r1 = rec4[1];
                    const uint2 mm = *reinterpret_cast<const uint2*>(&sm.lmask[j]);
                    const float2 dx = __fadd2_rn(make_float2(r0.x, r0.y), nth_x), dy = __fadd2_rn(make_float2(r0.z, r0.w), nth_y), dz = __fadd2_rn(make_float2(r1.x, r1.y), nth_z);
                    const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                    mine |= ((((mm.x >> lane) & 1u) && r2.x < hh4pf ? 1u : 0u) | (((mm.y >> lane) & 1u) && r2.y < hh4pf ? 2u : 0u)) << j;
                }
                {
                // Pair-parallel: the warp's candidate (target, source) pairs of this tile are enumerated in (target, source)
                // order and handed out one per lane, 32 at a time; results travel through shared memory back to the owning
                // lane, which adds its own pairs in source order (deterministic, identical to a per-target loop).
                const int npair = __popc(mine);
                const int incl = warp_incl_scan(npair, lane);
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                for (int c0 = 0; c0 < total; c0 += 32) {
                    const int p = c0 + lane;
                    int owner = 0;                                   // first lane whose inclusive count exceeds p
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const int probe = __shfl_sync(0xffffffffu, incl, owner + step - 1);
                        if (probe <= p) owner += step;
                    }
                    owner = min(owner, 31);
                    const int o_incl = __shfl_sync(0xffffffffu, incl, owner), o_cnt = __shfl_sync(0xffffffffu, npair, owner);
                    const unsigned o_mine = __shfl_sync(0xffffffffu, mine, owner);
                    double fx = 0, fy = 0, fz = 0, du = 0;
                    if (p < total) {
                        const int j = nth_set_bit(o_mine, p - (o_incl - o_cnt));
                        const int esrc = sm.list[j];
                        const double4 gv = sm.gst[j];                // (mVel | particle velocity, gasMass)
                        const double4 k4 = sm.tsph[3 * owner], tv = sm.tsph[3 * owner + 1];
                        const float* tf = reinterpret_cast<const float*>(&sm.tsph[3 * owner + 2]);
                        const float* rec = reinterpret_cast<const float*>(&sm.stage[j & ~1]) + (j & 1);
                        // displacement from the tile's float-float record (units of R), like the gravity loop; d = x_i - COM (Node.cpp:116)
                        const float sx = -((rec[0] + tf[0]) + (rec[8] + tf[3])), sy = -((rec[2] + tf[1]) + (rec[10] + tf[4])),
                                    sz = -((rec[4] + tf[2]) + (rec[12] + tf[5]));
                        const float r2s = fmaf(sz, sz, fmaf(sy, sy, sx * sx)), hh4o = tf[6];
                        const bool live = rec[6] != 0.f && r2s != 0.f;
                        bool pass = live && r2s < hh4o;
                        if (live && fabsf(r2s - hh4o) <= 4e-6f * hh4o) {
                            // too close to the gate for FP32: the reference's own separately rounded FP64 expression
                            const double4 q = P.src_pm[esrc], ot = P.src_pm[__float_as_int(tf[7])];
                            const double dx = q.x - ot.x, dy = q.y - ot.y, dz = q.z - ot.z;
                            const double r2e = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                            pass = __dsqrt_rn(r2e) < __dmul_rn(tv.w, 2.0);
                            tot_exact++;
                        }
                        if (pass) {
                            const float hs = (float)(tv.w * invR), inv_hs = 1.0f / hs, ipi4 = inv_hs * inv_hs * inv_hs * inv_hs * 0.318309886f;
                            float rinv;
                            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(r2s));
                            const float r = r2s * rinv, qq = r * inv_hs;
                            float gs = 0.f;                          // kernel.cpp:28-34
                            if (qq < 1.f) gs = -3.f * qq + 2.25f * qq * qq;
                            else if (qq < 2.f) { const float u = 2.f - qq; gs = -0.75f * u * u; }
                            const float gfac = gs * ipi4 * rinv;
                            const float gx = sx * gfac, gy = sy * gfac, gz = sz * gfac;   // grad W * R^4
                            const float vx = (float)(tv.x - gv.x), vy = (float)(tv.y - gv.y), vz = (float)(tv.z - gv.z);
                            const float vds = vx * sx + vy * sy + vz * sz;                // v_ij . d / R
                            const float mu = hs * vds / (r2s + 0.01f * hs * hs);
                            const float MUf = vds < 0.f ? (-0.5f * (float)k4.w * mu + mu * mu) : 0.f;   // Node.cpp:142-152
                            const double amu = k4.z + (double)MUf, coef = -gv.w * amu * invR4;          // Node.cpp:127 + :154
                            fx = coef * (double)gx; fy = coef * (double)gy; fz = coef * (double)gz;
                            du = 0.5 * gv.w * amu * invR4 * (double)(vx * gx + vy * gy + vz * gz);      // Node.cpp:167
                            if (isnan(fx) || isnan(fy) || isnan(fz)) { fx = 0; fy = 0; fz = 0; }         // Node.cpp:169
                            tot_sph++;
                        }
                        // a pair that fails the gate contributes exact zeros (x + 0.0 == x), so the owner adds every slot of its
                        // range unconditionally; only the counting variant needs to tell the two apart
                        sm.res[lane] = make_double4(fx, fy, fz, (COUNT && !pass) ? __longlong_as_double(0x7ff8000000000001ll) : du);
                    }
                    __syncwarp();
                    // my pairs inside this chunk are the contiguous lanes [a, b)
                    const int a = max(incl - npair, c0) - c0, b = min(incl, c0 + 32) - c0;
                    for (int i = a; i < b; i++) {
                        const double4 r = sm.res[i];
                        if (!COUNT) { ax += r.x; ay += r.y; az += r.z; dU += r.w; }
                        else if (__double_as_longlong(r.w) != 0x7ff8000000000001ll) { ax += r.x; ay += r.y; az += r.z; dU += r.w; c_sp++; }
                    }
                    __syncwarp();
                }
                }
            }
        }
        if (tgas) {
            const uint32_t p = P.perm[t];
            if (ax != 0.0) P.ax[p] += ax;
            if (ay != 0.0) P.ay[p] += ay;
            if (az != 0.0) P.az[p] += az;
            if (dU != 0.0) P.dUdt[p] += dU;
            if (COUNT) P.c_sph[t] = c_sp;
        }
    }
    tot_sph = warp_sum_u64(tot_sph); tot_exact = warp_sum_u64(tot_exact);
    if (lane == 0) {
        if (tot_sph) atomicAdd(&P.s->c_sph, tot_sph);
        if (tot_exact) atomicAdd(&P.s->c_exact, tot_exact);
    }
}

__global__ void k_walk_reset(AgbScalars* s)
{
    s->walk_next_group = 0; s->walk_overflow = 0; s->cand_cursor = 0ull;
    s->st_rounds = 0; s->st_popped = 0; s->st_mixed = 0; s->st_open = 0; s->st_drain = 0;
    for (int c = 0; c < 10; c++) s->st_cls[c] = 0;
    s->c_interactions = 0; s->c_node = 0; s->c_leaf = 0; s->c_sph = 0; s->c_visits = 0; s->c_exact = 0; s->c_spill = 0;
}

// ---- active targets (Tree.cpp:75: globalTime == nextIntegrationTime), compacted in tree order
__global__ void k_count_active(const double* __restrict__ s_next, int64_t n, double gt, AgbScalars* s)
{
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) local += s_next[i] == gt;
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0 && cnt) atomicAdd(&s->n_active, cnt);
}
// What a fused step (agb_force_path) took from the LAST step is checked here, on the device, before anything is walked — a walk
// adds to dU/dt, so it cannot simply be repeated.  (1) AGB_OPT_SLICE_DENSITIES assumed "every particle is a target" when it
// restricted the density passes to a range of tree positions.  (2) The FP32 pair law was chosen for the last step's root cube and
// depth (mixed_in_range, agb_api.cu).  If either does not hold for THIS tree the walk kernels return at once (walk_overflow = 3)
// and the host redoes the step call by call.
__global__ void k_step_guard(AgbScalars* s, int64_t n, bool all_targets, bool mixed, double e0)
{
    if (all_targets && s->n_active != n) s->walk_overflow = 3;
    if (mixed) {
        const double R = __longlong_as_double((long long)s->Rbits);
        if (R > 0.0 && !(e0 >= 1e-10 * R && e0 <= 50.0 * R && s->max_depth <= 40)) s->walk_overflow = 3;
    }
}
// the three kernels below return at once when every particle is active (the list is then the identity and never read)
__global__ void k_active_flags(const double* __restrict__ s_next, int64_t n, double gt, const AgbScalars* __restrict__ s, int32_t* __restrict__ flag)
{
    if (s->n_active == n) return;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = s_next[i] == gt;
}
__global__ void k_active_compact(const double* __restrict__ s_next, int64_t n, double gt, const AgbScalars* __restrict__ s, const int32_t* __restrict__ rank,
                                 int32_t* __restrict__ list)
{
    if (s->n_active == n) return;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && s_next[i] == gt) list[rank[i]] = (int32_t)i;
}

__global__ void k_unpermute_i32(const uint32_t* __restrict__ perm, int64_t n, const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* d_,
                                int32_t* oa, int32_t* ob, int32_t* oc, int32_t* od)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = perm[i];
    oa[p] = a[i]; ob[p] = b[i]; oc[p] = c[i]; od[p] = d_[i];
}

struct SliceCols { const double* src[9]; double* dst[9]; };     // ax ay az dUdt h rho P T visualDensity (caller order -> compact)
__global__ void k_slice_results(const uint32_t* __restrict__ perm, const int32_t* __restrict__ act_list, int64_t a0, int64_t a1, bool ident,
                                const SliceCols C, uint32_t* __restrict__ index)
{
    const int64_t i = a0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a1) return;
    const int64_t t = ident ? i : (int64_t)act_list[i];
    const uint32_t p = perm[t];
    const int64_t k = i - a0;
    if (index) index[k] = p;
#pragma unroll
    for (int c = 0; c < 9; c++) if (C.dst[c]) C.dst[c][k] = C.src[c][p];
}

template <bool COUNT, bool SPH, bool MIXED>
void launch_walk(const WalkParams& P, int64_t max_groups, int sm_count, int spill_warps, cudaStream_t st)
{
    constexpr int WARPS = WalkCfg<SPH>::WARPS;
    const int smem = (int)sizeof(WarpSmem<SPH, MIXED>) * WARPS;
    cudaFuncSetAttribute(k_walk<COUNT, SPH, MIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device; cheap, so not cached
    int blocks = (int)std::min<int64_t>((int64_t)sm_count * WALK_CTAS, (max_groups + WARPS - 1) / WARPS);
    if (blocks * WARPS > spill_warps) blocks = spill_warps / WARPS;
    k_walk<COUNT, SPH, MIXED><<<blocks, WalkCfg<SPH>::TPB, smem, st>>>(P);
}
template <bool COUNT, bool SPH>
void launch_walk2(const WalkParams& P, int64_t max_groups, int sm_count, int spill_warps, bool mixed, cudaStream_t st)
{
    if (mixed) launch_walk<COUNT, SPH, true>(P, max_groups, sm_count, spill_warps, st); else launch_walk<COUNT, SPH, false>(P, max_groups, sm_count, spill_warps, st);
}

} // namespace

// ---------------------------------------------------------------- roofline denominators, measured on the box
// kind 0: FP64 FMA chains, 1: FP32 FMA chains (results in TFLOP/s, FMA = 2 flop); kind 2: device copy (GB/s, read+write)
template <class T>
__global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters, T a, T b)
{
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    T r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (r == (T)123456.789) out[0] = r;
}
__global__ void __launch_bounds__(256) k_copy_peak(const double2* __restrict__ in, double2* __restrict__ out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = in[i];
}

int agb_walk_blocks(int sm_count) { return sm_count * WALK_CTAS; }
int agb_walk_warps_per_block() { return WalkCfg<false>::WARPS; }
void agb_far_capacity(int* lcap, int* fcap, int* targets) { *lcap = FAR_LCAP; *fcap = FAR_FCAP; *targets = 32 * SG_GROUPS; }

// phase 0: everything; 1: up to and including k_walk; 2: k_sph only (mixed mode with gas: the caller completes a late_gas
// build between 1 and 2).  ev[0..4]: before k_far, before k_walk, after k_walk, after k_sph, before k_sph.
int agb_launch_walk(AgbDev& d, AgbScalars* s, double globalTime, double e0, double theta, int part, int nparts,
                    bool counters, bool any_gas, bool mixed, int sm_count, cudaStream_t st, cudaEvent_t* ev, int phase)
{
    WalkParams P;
    P.src_pm = d.src_pm; P.src_gv = d.src_gv; P.src_flag = d.src_flag; P.child = d.child; P.ndepth = d.ndepth;
    P.s_h = d.s_h; P.s_rho = d.s_rho; P.s_P = d.s_P; P.s_next = d.s_next; P.s_type = d.s_type; P.perm = d.perm[d.cur];
    P.ax = d.ax; P.ay = d.ay; P.az = d.az; P.dUdt = d.dUdt;
    P.c_visits = d.c_visits; P.c_accn = d.c_accn; P.c_accl = d.c_accl; P.c_sph = d.c_sph;
    P.spill = d.spill; P.spill_per_warp = d.spill_per_warp;
    P.far_list = d.far_list; P.far_front = d.far_front; P.far_cnt = d.far_cnt;
    P.rec_ent = d.rec_ent; P.rec_next = d.rec_next; P.rec_head = d.rec_head; P.rec_cap = d.rec_cap;
    P.s = s; P.N = d.n; P.act_list = d.act_list; P.part = part; P.nparts = nparts;
    P.theta = theta; P.e0 = e0; P.globalTime = globalTime;
    // tuning knob (not part of the ABI): AGB200_WALK_FAR_K2=1e30 sends every pair through the float-float loop
    static const float far_k2 = getenv("AGB200_WALK_FAR_K2") ? (float)atof(getenv("AGB200_WALK_FAR_K2")) : 0.25f;
    P.far_k2 = far_k2;
    int launches = 0;
    const int64_t max_groups = (d.n / nparts + 32 * SG_GROUPS + 31) / 32;
    if (phase != 2) launches += agb_launch_active_list(d, s, globalTime, sm_count, st, mixed, e0);
    if (d.n > 0 && phase != 2) {
        if (counters) {
            cudaMemsetAsync(d.c_visits, 0, (size_t)d.n * 4, st); cudaMemsetAsync(d.c_accn, 0, (size_t)d.n * 4, st);
            cudaMemsetAsync(d.c_accl, 0, (size_t)d.n * 4, st); cudaMemsetAsync(d.c_sph, 0, (size_t)d.n * 4, st);
        }
        if (ev) cudaEventRecord(ev[0], st);
        // far-field prepass, one warp per super-group of 256 targets
        k_far<<<(int)std::min<int64_t>((max_groups / SG_GROUPS + 4) / 4, (int64_t)sm_count * 8), 128, 0, st>>>(P); launches++;
        if (ev) cudaEventRecord(ev[1], st);
        if (counters) { if (any_gas) launch_walk2<true, true>(P, max_groups, sm_count, d.spill_warps, mixed, st); else launch_walk2<true, false>(P, max_groups, sm_count, d.spill_warps, mixed, st); }
        else { if (any_gas) launch_walk2<false, true>(P, max_groups, sm_count, d.spill_warps, mixed, st); else launch_walk2<false, false>(P, max_groups, sm_count, d.spill_warps, mixed, st); }
        launches++;
        if (ev) cudaEventRecord(ev[2], st);
    }
    if (d.n > 0 && phase != 1) {
        if (ev) cudaEventRecord(ev[4], st);
        if (any_gas && mixed) {
            // the SPH pairs of the candidates the walk recorded (after it: acc += SPH part, dU/dt += ...)
            const int smem = (int)sizeof(SphWarp) * 8;
            const int blocks = (int)std::min<int64_t>((max_groups + 7) / 8, (int64_t)sm_count * 16);
            if (counters) { cudaFuncSetAttribute(k_sph<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_sph<true><<<blocks, 256, smem, st>>>(P); }
            else { cudaFuncSetAttribute(k_sph<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_sph<false><<<blocks, 256, smem, st>>>(P); }
            launches++;
        }
        if (ev) cudaEventRecord(ev[3], st);
    }
    return launches;
}

int agb_launch_active_list(AgbDev& d, AgbScalars* s, double globalTime, int sm_count, cudaStream_t st, bool mixed, double e0)
{
    int launches = 1;
    const int nb = (int)((d.n + 255) / 256);
    k_walk_reset<<<1, 1, 0, st>>>(s);
    if (d.n > 0) {
        cudaMemsetAsync(&s->n_active, 0, sizeof(int32_t), st);
        k_count_active<<<std::min(nb, 4 * sm_count), 256, 0, st>>>(d.s_next, d.n, globalTime, s);
        if (d.dens_a1 != INT64_MAX || mixed) { k_step_guard<<<1, 1, 0, st>>>(s, d.n, d.dens_a1 != INT64_MAX, mixed, e0); launches++; }
        // compact list of the active targets (scratch: flags -> nodecnt, ranks -> nodebase; both are idle after the densities)
        k_active_flags<<<nb, 256, 0, st>>>(d.s_next, d.n, globalTime, s, d.nodecnt);
        launches += 3 + agb_launch_scan_i32(d.nodecnt, d.nodebase, d.n, d.scanblk, &s->n_scan_tmp, st, &s->n_active);
        k_active_compact<<<nb, 256, 0, st>>>(d.s_next, d.n, globalTime, s, d.nodebase, d.act_list);
    }
    return launches;
}

// host mirror of target_slice()
void agb_slice_bounds(int64_t n_active, int part, int nparts, int64_t* a0, int64_t* a1)
{
    const int64_t g = 32 * SG_GROUPS, nsg = (n_active + g - 1) / g;
    *a0 = std::min(n_active, nsg * part / nparts * g);
    *a1 = std::min(n_active, nsg * (part + 1) / nparts * g);
}

int agb_launch_slice_results(const AgbDev& d, int64_t a0, int64_t a1, bool ident, uint32_t* index, double* const dst[9], cudaStream_t st)
{
    if (a1 <= a0) return 0;
    SliceCols C;
    const double* src[9] = {d.ax, d.ay, d.az, d.dUdt, d.h, d.rho, d.P, d.T, d.vis};
    for (int c = 0; c < 9; c++) { C.src[c] = src[c]; C.dst[c] = dst[c]; }
    k_slice_results<<<(int)((a1 - a0 + 255) / 256), 256, 0, st>>>(d.perm[d.cur], d.act_list, a0, a1, ident, C, index);
    return 1;
}

int agb_launch_microbench(int kind, int sm_count, cudaStream_t st, double* result)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    double work = 0;
    if (kind == 0 || kind == 1) {
        void* out = nullptr; cudaMalloc(&out, 64);
        const int blocks = sm_count * 8, iters = 1 << 14;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0, st);
            if (kind == 0) k_fma_peak<double><<<blocks, 256, 0, st>>>((double*)out, iters, 1.0000001, 1e-9);
            else k_fma_peak<float><<<blocks, 256, 0, st>>>((float*)out, iters, 1.0000001f, 1e-9f);
            cudaEventRecord(e1, st); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0) best = std::min(best, ms);
        }
        work = 2.0 * 8.0 * iters * 256.0 * blocks;                 // flop
        *result = work / (best * 1e-3) / 1e12;
        cudaFree(out);
    } else {
        const size_t n = (size_t)1 << 26;                            // 1 GiB in + 1 GiB out
        double2 *a = nullptr, *b = nullptr;
        if (cudaMalloc((void**)&a, n * 16) != cudaSuccess || cudaMalloc((void**)&b, n * 16) != cudaSuccess) { cudaFree(a); return 0; }
        cudaMemsetAsync(a, 0, n * 16, st);
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0, st);
            k_copy_peak<<<sm_count * 16, 256, 0, st>>>(a, b, n);
            cudaEventRecord(e1, st); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0) best = std::min(best, ms);
        }
        *result = 2.0 * n * 16 / (best * 1e-3) / 1e9;
        cudaFree(a); cudaFree(b);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 5;
}

int agb_launch_unpermute_counters(AgbDev& d, int32_t* v, int32_t* an, int32_t* al, int32_t* sp, cudaStream_t st)
{
    k_unpermute_i32<<<(int)((d.n + 255) / 256), 256, 0, st>>>(d.perm[d.cur], d.n, d.c_visits, d.c_accn, d.c_accl, d.c_sph, v, an, al, sp);
    return 1;
}

// agb_build.cu — octree construction for the B200 force path (sm_100a).
//
// Replaces Tree::buildTree / Tree::calcTreeWidth (Physics/Tree/Tree.cpp:24-55, :86-117) and the two
// insertion routines of the reference (Physics/Tree/Node.cpp:405-534 bulk, :597-699 one-by-one,
// :702-719 getOctant).  The reference builds a pointer octree top-down; here the SAME tree (same
// cells, same leaves, every level of every single-child chain kept) is produced bottom-up:
//
//   k_dist_partial/k_extent_*  root half-width R = max |x| among |x| <= mean + 10 sigma
//   k_keygen                   per particle, the octant path root->leaf by the reference's own FP64
//                              descent (strict '>' against cell centres produced by c +- r/2), packed
//                              3 bit/level into a 126-bit key; particles outside the root cube are
//                              flagged and sort to the end (they stay force targets, Node.cpp:606-612)
//   k_sort_*                   LSD radix sort, 8-bit digits, 8 passes over (key_hi, index); key_lo (levels 21..41) is
//                              computed lazily and only orders runs of equal key_hi (k_fix_runs)
//   k_gather                   permute particle data into tree order (unified source table)
//   k_lcp + scan               shared-levels of neighbouring keys -> one internal node per (first
//                              particle, depth) pair; ids by prefix sum (children get larger ids)
//   k_links                    ranges, parents and the 8 child slots by galloping searches on keys
//   k_upward                   monopole moments bottom-up, "last child to arrive computes the parent"
//   k_root_fix / k_finalize    reference quirk: with bulk insertion the ROOT's mass/COM include the
//                              particles outside the cube (Node.cpp:477-499); COM, mVel, duplication
//                              flags (nodes where bulk hands over to one-by-one insertion hold every
//                              particle twice in childParticles, Node.cpp:420-428,518,615)
//
// All kernels here are HBM-bandwidth bound; geometry is evaluated with explicit round-to-nearest
// intrinsics (no FMA contraction) so cell boundaries are the reference's, bit for bit.
#include "agb_internal.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------ root extent
// vec3::length (Math/vec3.cpp:76-78): sqrt(x*x + y*y + z*z), separately rounded
__device__ __forceinline__ double length3(double px, double py, double pz)
{
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz)));
}

__global__ void __launch_bounds__(TPB) k_dist_partial(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                                        const double* __restrict__ mass, int64_t n, double4* __restrict__ rec, AgbScalars* s)
{
    __shared__ double sh[2][TPB / 32];
    double a = 0.0, b = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
        double px = x[i], py = y[i], pz = z[i];
        double d = length3(px, py, pz);
        rec[i] = make_double4(px, py, pz, mass ? mass[i] : 0.0);    // (position, mass) packed once, coalesced: the gather reads one 32-byte sector per particle (mass == nullptr: k_fill_mass follows)
        a += d;
        b += __dmul_rn(d, d);
    }
    a = warp_sum(a); b = warp_sum(b);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = a; sh[1][w] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < TPB / 32; i++) { sa += sh[0][i]; sb += sh[1][i]; }
        s->partial_sum[blockIdx.x] = sa; s->partial_sq[blockIdx.x] = sb;
    }
}

__global__ void k_extent_finish(AgbScalars* s, int nblocks, int64_t n)
{   // one warp: fixed lane-strided partial sums + butterfly => deterministic
    double sa = 0.0, sb = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) { sa += s->partial_sum[i]; sb += s->partial_sq[i]; }
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    if (threadIdx.x != 0) return;
    double nn = (double)n;
    double mean = __ddiv_rn(sa, nn);
    double var = __dadd_rn(__ddiv_rn(sb, nn), -__dmul_rn(mean, mean));
    double sd = __dsqrt_rn(var);
    s->mean = mean; s->stdev = sd;
    s->limit = __dadd_rn(mean, __dmul_rn(10.0, sd));     // Tree.cpp:89,105
    s->Rbits = 0ull;
    s->n_long_runs = 0; s->any_gas = 0; s->n_outliers = 0; s->dup_keys = 0; s->edge_dropped = 0; s->max_depth = 0; s->n_nodes = 0; s->node_overflow = 0; s->need_deep = 0;
    s->next_uniform = 1; s->grid_bar = 0u; s->grid_bar2 = 0u;
    for (int k = 0; k < 64; k++) { s->lvl_cnt[k] = 0; s->lvl_cur[k] = 0; }
}

__global__ void __launch_bounds__(TPB) k_extent_max(const double4* __restrict__ rec, int64_t n, AgbScalars* s)
{
    __shared__ double sh[TPB / 32];
    const double lim = s->limit;
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
        const double4 r4 = rec[i];
        const double d = length3(r4.x, r4.y, r4.z);
        if (d <= lim && d > m) m = d;
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < TPB / 32; i++) m = fmax(m, sh[i]);
        atomicMax(&s->Rbits, (unsigned long long)__double_as_longlong(m));   // d >= 0: bit order == value order
    }
}

// masses that arrived after the positions (host hand-over: the extent, key and sort passes do not need them)
__global__ void __launch_bounds__(TPB) k_fill_mass(const double* __restrict__ mass, int64_t n, double4* __restrict__ rec)
{
    const int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) reinterpret_cast<double*>(rec + i)[3] = mass[i];
}

// ------------------------------------------------------------------ keys
// The reference's descent (Node.cpp:433-443 child centres, :702-719 getOctant): from level `l0` for `nl` levels,
// returns the packed octants and carries the cell on.  Level l < 21 sits at bit 60-3l of key_hi, 21 <= l < 42 at
// bit 60-3(l-21) of key_lo.
struct Cell { double cx, cy, cz, r; };

__device__ __forceinline__ uint64_t descend21(double px, double py, double pz, Cell& c, bool check_bounds_first, bool& edge)
{
    uint64_t key = 0;
#pragma unroll 1
    for (int l = 0; l < 21; l++) {
        const bool ox = px > c.cx, oy = py > c.cy, oz = pz > c.cz;       // Node.cpp:713-716, strict '>'
        if (l > 0 || check_bounds_first) {
            // the cell's inclusive bounds (Node.cpp:606-612): a coordinate above the centre can only violate the upper bound,
            // one at or below it only the lower bound, so one rounded sum and one comparison per axis decide
            const double tx = ox ? c.r : -c.r, ty = oy ? c.r : -c.r, tz = oz ? c.r : -c.r;
            const double bx = __dadd_rn(c.cx, tx), by = __dadd_rn(c.cy, ty), bz = __dadd_rn(c.cz, tz);
            edge |= (ox ? px > bx : px < bx) || (oy ? py > by : py < by) || (oz ? pz > bz : pz < bz);
        }
        key |= ((uint64_t)ox | ((uint64_t)oy << 1) | ((uint64_t)oz << 2)) << (60 - 3 * l);
        const double hr = __dmul_rn(c.r, 0.5);                             // Node.cpp:436-440
        c.cx = __dadd_rn(c.cx, ox ? hr : -hr);
        c.cy = __dadd_rn(c.cy, oy ? hr : -hr);
        c.cz = __dadd_rn(c.cz, oz ? hr : -hr);
        c.r = hr;
    }
    return key;
}

__device__ __forceinline__ uint32_t digit_of(uint64_t w, int shift) { return (uint32_t)(w >> shift) & 255u; }

// 21 bits -> every third bit of a 63-bit word (bit b -> bit 3 b)
__device__ __forceinline__ uint64_t spread3(uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x001f00000000ffffull;
    v = (v | (v << 16)) & 0x001f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// The first 21 levels without the descent, for all but ~1e-5 of the particles.  The reference's cell centres are the rounded
// chain c_{l+1} = rn(c_l +- R 2^-(l+1)) (Node.cpp:436-440): each step is off by <= ulp(R)/2, so every centre of levels 0..20
// lies within 11 ulp(R) < 2^-48 R of the exact value R (m / 2^l).  In the coordinate s = (x / R + 1) 2^20 in [0, 2^21] those
// centres are integers; s itself is computed to ~2^-29.  A particle whose s is further than 2^-20 from every integer on all
// three axes (2^-40 R in space) therefore takes all 21 strict comparisons, and the inclusive cell-bound tests (Node.cpp:606-612),
// exactly as the descent does, and its octants are the bits of floor(s).  Everything else takes the descent.
__device__ __forceinline__ bool key_hi_fast(double px, double py, double pz, double invR, uint64_t* key)
{
    const double sx = fma(px, invR, 1.0) * 1048576.0, sy = fma(py, invR, 1.0) * 1048576.0, sz = fma(pz, invR, 1.0) * 1048576.0;
    const double fx = floor(sx), fy = floor(sy), fz = floor(sz);
    const double dx = sx - fx, dy = sy - fy, dz = sz - fz;
    const double lo = 9.5367431640625e-07, hi = 1.0 - 9.5367431640625e-07;         // 2^-20
    if (!(dx > lo && dx < hi && dy > lo && dy < hi && dz > lo && dz < hi)) return false;
    if (!(fx >= 0.0 && fx < 2097152.0 && fy >= 0.0 && fy < 2097152.0 && fz >= 0.0 && fz < 2097152.0)) return false;
    const uint64_t kx = (uint64_t)fx, ky = (uint64_t)fy, kz = (uint64_t)fz;     // bit (20 - l) = "above the centre" at level l
    *key = spread3(kx) | (spread3(ky) << 1) | (spread3(kz) << 2);                 // level l at bit 60 - 3 l = 3 (20 - l)
    return true;
}

// key_hi (outlier flag + levels 0..20) for every particle; key_lo is produced on demand by key_lo_of().
// The digit histograms of all 8 radix passes are counted here as well (block-private counters in shared memory, one
// global add per non-empty bin), so the sort needs no histogram pass of its own.
__global__ void __launch_bounds__(TPB) k_keygen(const double4* __restrict__ rec, const uint8_t* __restrict__ type, int64_t n,
                                                  uint64_t* __restrict__ khi, uint32_t* __restrict__ perm, AgbScalars* s, uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t h[8][256];
    for (int k = threadIdx.x; k < 8 * 256; k += TPB) (&h[0][0])[k] = 0;
    __syncthreads();
    // grid-stride: a block keeps its 8 x 256 counters for many keys (clearing and flushing them per 256 keys cost more than the keys)
    int n_outl = 0, n_edge = 0;
    const double R = __longlong_as_double((long long)s->Rbits), invR = 1.0 / R;
    const int lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
        bool outl = false, edge = false;
        const double4 r4 = rec[i];
        const double px = r4.x, py = r4.y, pz = r4.z;
        uint64_t hi;
        // root cube is centred on the origin (Tree.cpp:31); inclusive bounds (Node.cpp:706-711)
        outl = px < -R || px > R || py < -R || py > R || pz < -R || pz > R;
        if (outl) hi = AGB_OUTLIER_BIT;           // the stable sort keeps caller order among the outliers
        else if (!key_hi_fast(px, py, pz, invR, &hi)) { Cell c{0.0, 0.0, 0.0, R}; hi = descend21(px, py, pz, c, false, edge); }
        khi[i] = hi; perm[i] = (uint32_t)i | (type[i] == 2 ? AGB_GAS_BIT : 0u);   // the sort payload also carries "is gas"
        // the upper digits (first ~8 levels) take few values inside a block: one add per distinct value of a warp instead of 32
        // serialised ones on the same counter; the lower digits are spread out
        const unsigned am = __activemask();
#pragma unroll
        for (int p = 0; p < 8; p++) {
            const uint32_t dg = digit_of(hi, 8 * p);
            if (p >= 5) {
                const unsigned peers = __match_any_sync(am, dg);
                if (lane == __ffs(peers) - 1) atomicAdd(&h[p][dg], (uint32_t)__popc(peers));
            } else atomicAdd(&h[p][dg], 1u);
        }
        n_outl += outl; n_edge += edge;
    }
    n_outl = __reduce_add_sync(0xffffffffu, n_outl); n_edge = __reduce_add_sync(0xffffffffu, n_edge);
    if (lane == 0) {
        if (n_outl) atomicAdd(&s->n_outliers, n_outl);
        if (n_edge) atomicAdd(&s->edge_dropped, n_edge);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 8 * 256; k += TPB) { const uint32_t v = (&h[0][0])[k]; if (v) atomicAdd(&ghist[k], v); }
}

// levels 21..41 of one particle (only needed where two particles share all of key_hi)
__device__ uint64_t key_lo_of(double px, double py, double pz, double R, AgbScalars* s)
{
    bool edge = false;
    Cell c{0.0, 0.0, 0.0, R};
    descend21(px, py, pz, c, false, edge);
    bool edge2 = false;
    const uint64_t lo = descend21(px, py, pz, c, true, edge2);
    if (edge2 && !edge) atomicAdd(&s->edge_dropped, 1);
    return lo;
}

// ---- deep builds: all 63 levels of every particle (caller order), digit totals per word by k_hist64, keys of the next
// word brought into the current order by k_gather_keys
__global__ void __launch_bounds__(TPB) k_keygen_deep(const double4* __restrict__ rec, const uint8_t* __restrict__ type, int64_t n,
                                                       uint64_t* __restrict__ khi, uint64_t* __restrict__ klo, uint64_t* __restrict__ kex, uint32_t* __restrict__ perm, AgbScalars* s)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    bool outl = false, edge = false;
    if (i < n) {
        const double R = __longlong_as_double((long long)s->Rbits);
        const double4 r4 = rec[i];
        const double px = r4.x, py = r4.y, pz = r4.z;
        uint64_t hi = AGB_OUTLIER_BIT, lo = 0, ex = 0;
        outl = px < -R || px > R || py < -R || py > R || pz < -R || pz > R;
        if (!outl) {
            Cell c{0.0, 0.0, 0.0, R};
            bool below = false;                                       // bound violations below level 21 only matter for particles that descend that far
            hi = descend21(px, py, pz, c, false, edge);
            lo = descend21(px, py, pz, c, true, below);
            ex = descend21(px, py, pz, c, true, below);
        }
        khi[i] = hi; klo[i] = lo; kex[i] = ex;
        perm[i] = (uint32_t)i | (type[i] == 2 ? AGB_GAS_BIT : 0u);
    }
    unsigned mo = __ballot_sync(0xffffffffu, outl), me = __ballot_sync(0xffffffffu, edge);
    if ((threadIdx.x & 31) == 0) {
        if (mo) atomicAdd(&s->n_outliers, __popc(mo));
        if (me) atomicAdd(&s->edge_dropped, __popc(me));
    }
}

__global__ void __launch_bounds__(TPB) k_hist64(const uint64_t* __restrict__ key, int64_t n, uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t h[8][256];
    for (int k = threadIdx.x; k < 8 * 256; k += TPB) (&h[0][0])[k] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
        const uint64_t w = key[i];
#pragma unroll
        for (int p = 0; p < 8; p++) atomicAdd(&h[p][digit_of(w, 8 * p)], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 8 * 256; k += TPB) { const uint32_t v = (&h[0][0])[k]; if (v) atomicAdd(&ghist[k], v); }
}

__global__ void __launch_bounds__(TPB) k_gather_keys(const uint64_t* __restrict__ src, const uint32_t* __restrict__ perm, int64_t n, uint64_t* __restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) dst[i] = src[perm[i] & AGB_IDX_MASK];
}

// ------------------------------------------------------------------ LSD radix sort (8-bit digits), one sweep per pass
// Only key_hi (outlier flag + the first 21 levels) is radix sorted, together with the caller index: 8 stable passes over
// 12-byte items.  key_lo (levels 21..41) matters only between particles that share all 21 upper levels; it is computed
// afterwards and those (short, rare) runs are ordered by k_fix_runs / k_fix_long_runs.
//
// One kernel per pass (Adinets & Merrill's Onesweep): the digit totals of all passes come from k_keygen, tiles are taken
// in ticket order, every tile publishes its per-digit counts and finds its global offsets by decoupled look-back over the
// tiles before it (Merrill & Garland's single-pass scan) — no histogram / scan kernels between the passes.  Inside a tile:
// per-warp digit counts by shared-memory adds ("early counts": the tile totals are published before any ranking), ranks by
// eight ballots per key (no MATCH.ANY, no shared-memory round trip per key), then the tile is reordered in shared memory
// so that every digit's keys leave as one contiguous run.
// Status word per (pass, tile, digit): [31:30] 0 = not ready, 1 = tile count, 2 = inclusive prefix; [29:0] value (n < 2^30).
template <int ITEMS> struct SortCfg { static constexpr int TILE = TPB * ITEMS; };
enum : uint32_t { ST_AGG = 1u << 30, ST_INC = 2u << 30, ST_VAL = (1u << 30) - 1u };

template <int ITEMS>
__global__ void __launch_bounds__(TPB, 3) k_sort_onesweep(const uint64_t* __restrict__ ihi, const uint32_t* __restrict__ iv,
                                                         uint64_t* __restrict__ ohi, uint32_t* __restrict__ ov, int64_t n, int pass,
                                                         const uint32_t* __restrict__ ghist, uint32_t* status, uint32_t* ticket)
{
    constexpr int TILE = SortCfg<ITEMS>::TILE;
    extern __shared__ __align__(16) unsigned char sort_smem[];
    uint64_t* t_hi = reinterpret_cast<uint64_t*>(sort_smem);                 // [TILE]
    uint32_t* t_v = reinterpret_cast<uint32_t*>(t_hi + TILE);                // [TILE]
    uint32_t (*wcnt)[256] = reinterpret_cast<uint32_t (*)[256]>(t_v + TILE); // [TPB/32][256] per-warp digit counts, then running offsets
    uint32_t* gbase = &wcnt[TPB / 32][0];                                    // [256] global position of the tile's first key of each digit
    uint32_t* lbase = gbase + 256;                                           // [256] tile-local position of the first key of each digit
    __shared__ uint32_t ws[TPB / 32];
    __shared__ uint32_t s_tile;
    const int shift = 8 * pass;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (threadIdx.x == 0) s_tile = atomicAdd(&ticket[pass], 1u);             // tiles in ticket order: a tile only ever waits for running ones
    for (int k = threadIdx.x; k < (TPB / 32) * 256; k += TPB) (&wcnt[0][0])[k] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile0 = (int64_t)tile * TILE;
    const int64_t base = tile0 + (int64_t)w * (32 * ITEMS) + l;
    uint64_t kh[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const int64_t i = base + j * 32;
        kh[j] = i < n ? ihi[i] : ~0ull;
    }
    // per-warp digit counts (the counters of warp w are only ever touched by warp w)
#pragma unroll
    for (int j = 0; j < ITEMS; j++) if (base + j * 32 < n) atomicAdd(&wcnt[w][digit_of(kh[j], shift)], 1u);
    __syncthreads();
    uint32_t dbase, cnt_d;                                                    // thread d owns digit d
    {
        uint32_t off = 0;
#pragma unroll
        for (int k = 0; k < TPB / 32; k++) { const uint32_t t = wcnt[k][threadIdx.x]; wcnt[k][threadIdx.x] = off; off += t; }   // -> exclusive offsets of the warps
        cnt_d = off;
        volatile uint32_t* st = status + ((size_t)pass * gridDim.x + tile) * 256;
        st[threadIdx.x] = (tile == 0 ? ST_INC : ST_AGG) | cnt_d;            // published before the ranking: the tiles behind need not wait for it
        // first output position of digit d (exclusive scan of the 256 totals) and first tile-local position (scan of the tile's counts)
        uint32_t v = ghist[pass * 256 + threadIdx.x], inc = v, linc = cnt_d;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o), u = __shfl_up_sync(0xffffffffu, linc, o);
            if (l >= o) { inc += t; linc += u; }
        }
        if (l == 31) { ws[w] = inc; gbase[w] = linc; }                        // gbase[0..7] doubles as scratch for the second scan
        __syncthreads();
        uint32_t o1 = 0, o2 = 0;
        for (int k = 0; k < w; k++) { o1 += ws[k]; o2 += gbase[k]; }
        dbase = o1 + inc - v;
        __syncthreads();
        lbase[threadIdx.x] = o2 + linc - cnt_d;
    }
    __syncthreads();
    // stable ranks: keys of one warp with the same digit are numbered in lane order, item by item
    uint32_t rk[ITEMS];
    const unsigned lt = (1u << l) - 1u;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const bool valid = base + j * 32 < n;
        const uint32_t d = digit_of(kh[j], shift);
        unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const bool bit = (d >> b) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? m : ~m;
        }
        const uint32_t prev = wcnt[w][d];
        __syncwarp();
        if (valid && (peers & lt) == 0) wcnt[w][d] = prev + __popc(peers);
        __syncwarp();
        rk[j] = lbase[d] + prev + __popc(peers & lt);
    }
    // look back for the keys of digit d in the tiles before this one
    {
        volatile uint32_t* st = status + (size_t)pass * gridDim.x * 256;
        uint32_t excl = 0;
        if (tile > 0) {
            int64_t t = (int64_t)tile - 1;
            for (bool done = false; !done;) {
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = t - k >= 0 ? st[(size_t)(t - k) * 256 + threadIdx.x] : ST_INC;   // before tile 0: prefix 0
                int used = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (done || used != k || (v[k] >> 30) == 0u) continue;  // not published yet: read again from here
                    excl += v[k] & ST_VAL;
                    used = k + 1;
                    done = (v[k] >> 30) == 2u;
                }
                t -= used;
            }
            st[(size_t)tile * 256 + threadIdx.x] = ST_INC | (excl + cnt_d);
        }
        gbase[threadIdx.x] = dbase + excl;
    }
    // keys first (their registers are free afterwards), then the payload: it is only read now, so the kernel fits 3 blocks per SM
#pragma unroll
    for (int j = 0; j < ITEMS; j++) if (base + j * 32 < n) t_hi[rk[j]] = kh[j];
    {
        uint32_t kv[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++) kv[j] = base + j * 32 < n ? iv[base + j * 32] : 0u;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) if (base + j * 32 < n) t_v[rk[j]] = kv[j];
    }
    __syncthreads();
    const int cnt = (int)min((int64_t)TILE, n - tile0);
    for (int k = threadIdx.x; k < cnt; k += TPB) {
        const uint64_t h = t_hi[k];
        const uint32_t d = digit_of(h, shift);
        const uint32_t pos = gbase[d] + ((uint32_t)k - lbase[d]);
        ohi[pos] = h; ov[pos] = t_v[k];
    }
}

// runs of in-tree particles with equal key_hi (they share >= 21 levels): order them by key_lo.
// Short runs are insertion-sorted by the thread at the run start; long ones are queued for k_fix_long_runs.
constexpr int FIX_SHORT = 16, FIX_LONG_MAX = 4096;
__global__ void __launch_bounds__(TPB) k_fix_runs(const uint64_t* __restrict__ hi, uint64_t* __restrict__ lo, uint32_t* __restrict__ perm, int64_t n,
                                                    const double4* __restrict__ rec, AgbScalars* s, int32_t* __restrict__ longlist)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    const int64_t nt = n - s->n_outliers;
    if (i + 1 >= nt) return;
    const uint64_t h = hi[i];
    if (hi[i + 1] != h || (i > 0 && hi[i - 1] == h)) return;       // not the first element of a run of >= 2
    int64_t e = i + 1;
    while (e + 1 < nt && hi[e + 1] == h) e++;
    const int len = (int)(e - i + 1);
    if (len > FIX_SHORT) {
        if (len > FIX_LONG_MAX) { s->need_deep = 1; return; }        // the host rebuilds with three-word keys and three sorts
        longlist[atomicAdd(&s->n_long_runs, 1)] = (int32_t)i;
        return;
    }
    const double R = __longlong_as_double((long long)s->Rbits);
    for (int a = 0; a < len; a++) { const double4 r4 = rec[perm[i + a] & AGB_IDX_MASK]; lo[i + a] = key_lo_of(r4.x, r4.y, r4.z, R, s); }
    for (int a = 1; a < len; a++) {                                  // stable insertion sort
        const uint64_t kl = lo[i + a]; const uint32_t kp = perm[i + a];
        int b = a - 1;
        while (b >= 0 && lo[i + b] > kl) { lo[i + b + 1] = lo[i + b]; perm[i + b + 1] = perm[i + b]; b--; }
        lo[i + b + 1] = kl; perm[i + b + 1] = kp;
    }
}

__global__ void __launch_bounds__(TPB) k_fix_long_runs(const uint64_t* __restrict__ hi, uint64_t* __restrict__ lo, uint32_t* __restrict__ perm, int64_t n,
                                                         const double4* __restrict__ rec, AgbScalars* s, const int32_t* __restrict__ longlist)
{
    __shared__ uint64_t sl[FIX_LONG_MAX];
    __shared__ uint32_t sp[FIX_LONG_MAX];
    const int64_t nt = n - s->n_outliers;
    const double R = __longlong_as_double((long long)s->Rbits);
    for (int q = blockIdx.x; q < s->n_long_runs; q += gridDim.x) {
        const int64_t i = longlist[q];
        const uint64_t h = hi[i];
        int64_t e = i;
        while (e + 1 < nt && hi[e + 1] == h) e++;
        const int len = (int)(e - i + 1);
        int p2 = 1;
        while (p2 < len) p2 <<= 1;
        __syncthreads();
        for (int j = threadIdx.x; j < p2; j += TPB) {
            if (j < len) { const uint32_t p = perm[i + j]; const double4 r4 = rec[p & AGB_IDX_MASK]; sp[j] = p; sl[j] = key_lo_of(r4.x, r4.y, r4.z, R, s); }
            else { sl[j] = ~0ull; sp[j] = 0xffffffffu; }
        }
        __syncthreads();
        for (int size = 2; size <= p2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int j = threadIdx.x; j < p2; j += TPB) {
                    const int partner = j ^ stride;
                    if (partner > j) {
                        const bool up = (j & size) == 0;
                        const uint64_t a = sl[j], b = sl[partner];
                        // ties cannot occur between real items (42-level duplicates are an error reported by k_lcp)
                        if ((a > b) == up && a != b) { sl[j] = b; sl[partner] = a; const uint32_t t = sp[j]; sp[j] = sp[partner]; sp[partner] = t; }
                    }
                }
                __syncthreads();
            }
        for (int j = threadIdx.x; j < len; j += TPB) { lo[i + j] = sl[j]; perm[i + j] = sp[j]; }
    }
}

// ------------------------------------------------------------------ gather into tree order
// "every particle is due at the same time" (the shipped fixed-step configuration): then the tree-order copy of
// nextIntegrationTime is a constant and the gather need not fetch it with one random 32-byte sector per particle
__global__ void __launch_bounds__(TPB) k_next_uniform(const double* __restrict__ next, int64_t n, AgbScalars* s)
{
    const double v0 = next[0];
    bool differ = false;
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) differ |= next[i] != v0;
    if (__any_sync(0xffffffffu, differ) && (threadIdx.x & 31) == 0) s->next_uniform = 0;
}

// Caller order, streaming: the eight fields only gas particles carry, packed into one 64-byte record per gas particle
// (indexed by caller position, written for gas only).  A random 8-byte read costs a whole DRAM burst (measured: 534 B of DRAM
// reads per particle in the old gather, 16 times the bytes it used), so the gather below touches ONE record per particle
// (+ one per gas particle) instead of one sector per field.
__global__ void __launch_bounds__(TPB) k_pack_gas(AgbDev d)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n || d.type[i] != 2) return;
    d.grec[2 * i] = make_double4(d.vx ? d.vx[i] : 0.0, d.vy ? d.vy[i] : 0.0, d.vz ? d.vz[i] : 0.0, d.U ? d.U[i] : 0.0);
    d.grec[2 * i + 1] = make_double4(d.mu ? d.mu[i] : 0.58, d.rho[i], d.P[i], d.T[i]);
}

// LATE: the hand-over is still uploading velocities / U / mu (agb_force_path with host arrays): only what the build, the
// densities and the gravity walk read is gathered here; k_gather_late fills in the rest once it has arrived.
// split (with LATE): the gas columns follow at once in k_gather_gas, nothing is pending — everything else as in a late build
template <bool LATE>
__global__ void __launch_bounds__(TPB) k_gather(AgbDev d, uint32_t* __restrict__ perm, AgbScalars* s, bool split)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n) return;
    const uint32_t val = perm[i], p = val & AGB_IDX_MASK;
    const bool gas = (val & AGB_GAS_BIT) != 0u;
    perm[i] = p;                                      // from here on the permutation is a plain caller index
    const double4 r4 = d.rec[p];                      // (x, y, z, mass), one sector
    const double m = r4.w;
    d.src_pm[i] = r4;
    d.s_type[i] = gas ? 2 : 1;                        // only "gas or not" matters on the path (Node.cpp:319,371,478,679,763)
    if (!LATE || split) d.s_next[i] = !d.next ? 0.0 : s->next_uniform ? d.next[0] : d.next[p];
    if (gas) {
        d.src_flag[i] = m > 0.0 ? 1 : 0;
        s->any_gas = 1;
        d.s_h[i] = 0.0;                               // Tree.cpp:123-133 zeroes h of every gas particle
        if (LATE) { if (!split) d.src_gv[i] = make_double4(0.0, 0.0, 0.0, m); }
        else {
            // velocity, U, mu and the carried h/rho/P/T are only ever read for gas (Node.cpp:88-172, :722-796)
            const double4 g0 = d.grec[2 * (size_t)p], g1 = d.grec[2 * (size_t)p + 1];   // (vx, vy, vz, U), (mu, rho, P, T)
            d.src_gv[i] = make_double4(g0.x, g0.y, g0.z, m);
            d.s_U[i] = g0.w;
            d.s_mu[i] = g1.x;
            d.s_rho[i] = g1.y; d.s_P[i] = g1.z; d.s_T[i] = g1.w;
        }
    } else {
        d.src_gv[i] = make_double4(0.0, 0.0, 0.0, 0.0);
        d.src_flag[i] = 0;
    }
    d.group[i] = -1;
    d.leafparent[i] = -1;
    d.leafdepth[i] = -1;
    d.leafmark[i] = 0;
}

// The gas columns of a split gather: velocity, U, mu and the carried rho / P / T from the packed caller-order records.
__global__ void __launch_bounds__(TPB) k_gather_gas(AgbDev d, const uint32_t* __restrict__ perm)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n || d.s_type[i] != 2) return;
    const uint32_t p = perm[i];
    const double4 g0 = d.grec[2 * (size_t)p], g1 = d.grec[2 * (size_t)p + 1];
    d.src_gv[i] = make_double4(g0.x, g0.y, g0.z, d.src_pm[i].w);
    d.s_U[i] = g0.w;
    d.s_mu[i] = g1.x;
    d.s_rho[i] = g1.y; d.s_P[i] = g1.z; d.s_T[i] = g1.w;
}

// next_time of a LATE gather (it arrives behind the masses; nothing of the build reads it)
__global__ void __launch_bounds__(TPB) k_gather_next(AgbDev d, const uint32_t* __restrict__ perm, const AgbScalars* __restrict__ s)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < d.n) d.s_next[i] = !d.next ? 0.0 : s->next_uniform ? d.next[0] : d.next[perm[i]];
}

// The rest of a LATE gather, after the gas densities: velocity, U, mu of the gas particles; P and T of the particles that
// belong to a density group (Node.cpp:789-791, with the rho the group pass has just written), in tree and in caller order;
// orphans keep the state they were handed over with.
__global__ void __launch_bounds__(TPB) k_gather_late(AgbDev d, const uint32_t* __restrict__ perm, const AgbScalars* __restrict__ s)
{
    constexpr double kGAMMA = 5.0 / 3.0, kKB = 1.38064852e-23, kPRTN = 1.6726219e-27;   // Math/Constants.h:15-18
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n || s->node_overflow || d.s_type[i] != 2) return;
    const uint32_t p = perm[i];
    const double4 g0 = d.grec[2 * (size_t)p], g1 = d.grec[2 * (size_t)p + 1];
    d.src_gv[i] = make_double4(g0.x, g0.y, g0.z, d.src_pm[i].w);
    d.s_U[i] = g0.w;
    d.s_mu[i] = g1.x;
    if (i < s->n_in_tree && d.group[i] >= 0) {
        const double P = (kGAMMA - 1.0) * g0.w * d.s_rho[i], T = (kGAMMA - 1.0) * g0.w * kPRTN * g1.x / kKB;
        d.s_P[i] = P; d.s_T[i] = T;
        d.P[p] = P; d.T[p] = T;
    } else { d.s_rho[i] = g1.y; d.s_P[i] = g1.z; d.s_T[i] = g1.w; }
}

// ------------------------------------------------------------------ shared levels of neighbours, node counts
__global__ void __launch_bounds__(TPB) k_lcp(const uint64_t* __restrict__ khi, const uint64_t* __restrict__ klo, const uint64_t* __restrict__ kex, int64_t n,
                                               int8_t* __restrict__ lcp, int32_t* __restrict__ cnt, AgbScalars* s)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const int64_t nt = n - s->n_outliers;
    const bool deep = kex != nullptr;
    const int cap = deep ? AGB_MAX_LEVELS : AGB_SHALLOW_LEVELS;     // levels this build can tell apart
    int Li = -1, Lm = -1;
    if (i < nt) {
        const uint64_t h = khi[i], l = klo[i], x = deep ? kex[i] : 0ull;
        if (i + 1 < nt) {
            Li = agb_common_levels(h, l, x, khi[i + 1], klo[i + 1], deep ? kex[i + 1] : 0ull, deep);
            if (Li >= cap) {
                // shallow build: the third key word is needed (the host rebuilds); deep build: coincident points (Node.cpp:618-666 never returns)
                if (deep) atomicAdd(&s->dup_keys, 1); else s->need_deep = 1;
                Li = cap - 1;
            }
        }
        if (i > 0) { Lm = agb_common_levels(khi[i - 1], klo[i - 1], deep ? kex[i - 1] : 0ull, h, l, x, deep); if (Lm >= cap) Lm = cap - 1; }
        int md = max(Li, Lm) + 1;
        if (md > s->max_depth) atomicMax(&s->max_depth, md);
    }
    lcp[i] = (int8_t)Li;
    cnt[i] = Li > Lm ? Li - Lm : 0;
    if (i == 0) s->n_in_tree = (int32_t)nt;
}

// ------------------------------------------------------------------ exclusive scan (int32)
constexpr int SCAN_TILE = 2048;
__device__ __forceinline__ int block_excl_scan(int v, int* total)
{   // 256 threads
    __shared__ int ws[TPB / 32];
    __shared__ int tot;
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (l >= o) inc += t; }
    __syncthreads();
    if (l == 31) ws[w] = inc;
    __syncthreads();
    int off = 0;
    for (int k = 0; k < w; k++) off += ws[k];
    if (threadIdx.x == TPB - 1) tot = off + inc;
    __syncthreads();
    *total = tot;
    return off + inc - v;
}

__global__ void __launch_bounds__(TPB) k_scan_reduce(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ blk, const int32_t* skip_if_n)
{
    if (skip_if_n && *skip_if_n == n) return;
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    int v = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) if (base + j < n) v += in[base + j];
    int tot;
    block_excl_scan(v, &tot);
    if (threadIdx.x == 0) blk[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(TPB) k_scan_blocks(int32_t* __restrict__ blk, int nb, int32_t* total_out, const int32_t* skip_if_n, int64_t n)
{
    if (skip_if_n && *skip_if_n == n) return;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += TPB) {
        int i = b0 + threadIdx.x;
        int v = i < nb ? blk[i] : 0, tot;
        int e = block_excl_scan(v, &tot);
        int c = carry;
        if (i < nb) blk[i] = c + e;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(TPB) k_scan_apply(const int32_t* __restrict__ in, int64_t n, const int32_t* __restrict__ blk, int32_t* __restrict__ out, const int32_t* skip_if_n)
{
    if (skip_if_n && *skip_if_n == n) return;
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    int v[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { v[j] = base + j < n ? in[base + j] : 0; sum += v[j]; }
    int tot;
    int e = block_excl_scan(sum, &tot) + blk[blockIdx.x];
#pragma unroll
    for (int j = 0; j < 8; j++) { if (base + j < n) out[base + j] = e; e += v[j]; }
}

// ------------------------------------------------------------------ links
__global__ void __launch_bounds__(TPB) k_init_nodes(AgbDev d, const AgbScalars* __restrict__ s)
{
    int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (s->n_nodes > d.ncap || s->need_deep) {          // more (first particle, depth) pairs than node slots (or keys too short): nothing is written,
        if (k == 0) const_cast<AgbScalars*>(s)->node_overflow = 1;   // every later kernel of the step returns at once, the host grows the arrays
        return;
    }
    if (k >= s->n_nodes) return;
    int4 e = make_int4(-1, -1, -1, -1);
    reinterpret_cast<int4*>(d.child)[2 * k] = e;
    reinterpret_cast<int4*>(d.child)[2 * k + 1] = e;
    d.nmark[k] = 0;
}

struct KeyView { const uint64_t* hi; const uint64_t* lo; const uint64_t* ex; };     // ex == nullptr: two-word build

// "key j shares >= d levels with (h, l, x)" for a fixed d: one shift of the xor; the lower words are only read for d > 21 / d > 42
struct SharesLevels {
    KeyView K; uint64_t h, l, x; int word, sh;
    __device__ __forceinline__ SharesLevels(KeyView K_, uint64_t h_, uint64_t l_, uint64_t x_, int d) : K(K_), h(h_), l(l_), x(x_)
    {
        word = d > 42 ? 2 : d > 21 ? 1 : 0;
        sh = 63 - 3 * (d - 21 * word);                // the top 1 + 3 d' bits of the deciding word (the flag bit of hi, an always-zero bit otherwise)
    }
    __device__ __forceinline__ bool operator()(int64_t j) const
    {
        if (word == 0) return ((K.hi[j] ^ h) >> sh) == 0ull;
        if (word == 1) return K.hi[j] == h && ((K.lo[j] ^ l) >> sh) == 0ull;
        return K.hi[j] == h && K.lo[j] == l && ((K.ex[j] ^ x) >> sh) == 0ull;
    }
};

// smallest s <= i whose key shares >= d levels with key i
__device__ __forceinline__ int64_t find_first(KeyView K, int64_t i, int d, uint64_t h, uint64_t l, uint64_t x)
{
    if (d <= 0) return 0;
    const SharesLevels same(K, h, l, x, d);
    int64_t step = 1;
    while (i - step >= 0 && same(i - step)) step <<= 1;
    int64_t lo = max((int64_t)-1, i - step), hi = i - (step >> 1);      // key[lo] fails (or lo == -1), key[hi] passes
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (same(mid)) hi = mid; else lo = mid;
    }
    return hi;
}

__device__ __forceinline__ int node_id_at(const AgbDev& d, int64_t s, int depth)
{   // id of the node of depth `depth` whose first particle is s
    int Lprev = s > 0 ? (int)d.lcp[s - 1] : -1;
    return d.nodebase[s] + (depth - Lprev - 1);
}

__global__ void __launch_bounds__(TPB) k_links(AgbDev d, const uint64_t* __restrict__ khi, const uint64_t* __restrict__ klo, const uint64_t* __restrict__ kex, AgbScalars* s)
{
    __shared__ int lvl[AGB_MAX_LEVELS + 1];                  // nodes of this block per depth (for the level lists of the upward pass)
    if (threadIdx.x <= AGB_MAX_LEVELS) lvl[threadIdx.x] = 0;
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    const int64_t nt = s->n_in_tree;
    if (s->node_overflow) return;                            // whole grid: no usable node table this step (k_init_nodes)
    if (i < nt && nt < 2) d.leafdepth[i] = 0;                // a single particle: the root itself is the leaf (Node.cpp:409-418)
    if (i < nt && nt >= 2) {
    KeyView K{khi, klo, kex};
    const uint64_t h = khi[i], l = klo[i], x = kex ? kex[i] : 0ull;
    const int Li = d.lcp[i], Lm = i > 0 ? (int)d.lcp[i - 1] : -1;
    const int base = d.nodebase[i];
    const int N = (int)d.n;
    for (int dep = Lm + 1; dep <= Li; dep++) {
        int k = base + (dep - Lm - 1);
        atomicAdd(&lvl[dep], 1);
        d.ndepth[k] = (int8_t)dep;
        d.nfirst[k] = (int32_t)i;
        int par;
        if (dep == Lm + 1) {
            if (dep == 0) par = -1;
            else { int64_t sfirst = find_first(K, i, dep - 1, h, l, x); par = node_id_at(d, sfirst, dep - 1); }
        } else par = k - 1;
        d.nparent[k] = par;
        if (par >= 0) d.child[(size_t)par * 8 + agb_octant_at(h, l, x, dep - 1)] = N + k;
    }
    // the particle's own leaf hangs below the deepest internal node that contains it
    const int dp = max(Li, Lm);
    int par;
    if (Li > Lm) par = base + (Li - Lm - 1);
    else { int64_t sfirst = find_first(K, i, dp, h, l, x); par = node_id_at(d, sfirst, dp); }
    d.child[(size_t)par * 8 + agb_octant_at(h, l, x, dp)] = (int32_t)i;
    d.leafparent[i] = par;
    d.leafdepth[i] = (int8_t)(dp + 1);
    }
    __syncthreads();
    if (threadIdx.x <= AGB_MAX_LEVELS && lvl[threadIdx.x]) atomicAdd(&s->lvl_cnt[threadIdx.x], lvl[threadIdx.x]);
}

// ------------------------------------------------------------------ upward pass (monopole + gas moments)
__device__ __forceinline__ double4 ldcg4(const double4* p)
{   // L2-coherent read of data another SM has just published
    const double2 a = __ldcg(reinterpret_cast<const double2*>(p)), b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

struct ChildLinks { int4 a, b; };
__device__ __forceinline__ ChildLinks load_links(const int32_t* child, int k)
{
    ChildLinks L;
    L.a = reinterpret_cast<const int4*>(child)[2 * (size_t)k]; L.b = reinterpret_cast<const int4*>(child)[2 * (size_t)k + 1];
    return L;
}

// MODE 0: everything; MODE 1: mass moments and gasMass (velocities not there yet: mVel sums stay 0); MODE 2: the mVel sums alone,
// with the same operations in the same order as MODE 0 (bit-identical).
template <int MODE>
__device__ __forceinline__ void node_moments(const AgbDev& d, int k, int N, bool any_gas, const ChildLinks& L)
{
    double sx = 0, sy = 0, sz = 0, m = 0, gx = 0, gy = 0, gz = 0, g = 0;
    // the node's particle range ends where its last child's range ends (children are in key order)
    int last = -1;
    const int ch[8] = {L.a.x, L.a.y, L.a.z, L.a.w, L.b.x, L.b.y, L.b.z, L.b.w};
    if (MODE == 2) {
        const double4 own = ldcg4(&d.mom_gv[k]);
        if (!(own.w > 0.0)) return;                     // no gas below this node: mVel stays 0
#pragma unroll
        for (int o = 0; o < 8; o++) {
            const int c = ch[o];
            if (c < 0) continue;
            if (c < N) { const double4 gv = d.src_gv[c]; gx += gv.x * gv.w; gy += gv.y * gv.w; gz += gv.z * gv.w; }
            else { const double4 gv = ldcg4(&d.mom_gv[c - N]); gx += gv.x; gy += gv.y; gz += gv.z; }
        }
        d.mom_gv[k] = make_double4(gx, gy, gz, own.w);
        return;
    }
    if (!any_gas) {
#pragma unroll
        for (int o = 0; o < 8; o++) {
            const int c = ch[o];
            if (c < 0) continue;
            if (c < N) { double4 pm = d.src_pm[c]; m += pm.w; sx += pm.x * pm.w; sy += pm.y * pm.w; sz += pm.z * pm.w; last = max(last, c); }
            else { double4 pm = ldcg4(&d.mom_pm[c - N]); m += pm.w; sx += pm.x; sy += pm.y; sz += pm.z; last = max(last, __ldcg(&d.nlast[c - N])); }
        }
        d.mom_pm[k] = make_double4(sx, sy, sz, m);
        d.nlast[k] = last;
        return;
    }
    // Four children at a time: their records are requested together (independent loads in flight: the pass is latency bound),
    // then summed in fixed octant order => run-to-run identical sums.  (Gating the gas record of a particle by src_flag saves
    // 0.7 GB of DRAM reads on C3 but adds a dependent load: 0.1 ms slower.)
#pragma unroll
    for (int half = 0; half < 2; half++) {
        double4 pm[4], gv[4];
        int lst[4];
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int c = ch[4 * half + o];
            pm[o] = make_double4(0, 0, 0, 0); gv[o] = pm[o]; lst[o] = -1;
            if (c >= N) { pm[o] = ldcg4(&d.mom_pm[c - N]); gv[o] = ldcg4(&d.mom_gv[c - N]); lst[o] = __ldcg(&d.nlast[c - N]); }
            else if (c >= 0) { pm[o] = d.src_pm[c]; gv[o] = d.src_gv[c]; lst[o] = c; }
        }
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int c = ch[4 * half + o];
            if (c < 0) continue;
            if (c < N) {
                m += pm[o].w; sx += pm[o].x * pm[o].w; sy += pm[o].y * pm[o].w; sz += pm[o].z * pm[o].w;
                g += gv[o].w;
                if (MODE == 0) { gx += gv[o].x * gv[o].w; gy += gv[o].y * gv[o].w; gz += gv[o].z * gv[o].w; }
            } else {
                m += pm[o].w; sx += pm[o].x; sy += pm[o].y; sz += pm[o].z;
                g += gv[o].w;
                if (MODE == 0) { gx += gv[o].x; gy += gv[o].y; gz += gv[o].z; }
            }
            last = max(last, lst[o]);
        }
    }
    d.mom_pm[k] = make_double4(sx, sy, sz, m);
    d.mom_gv[k] = make_double4(gx, gy, gz, g);
    d.nlast[k] = last;
}

// Level lists: the nodes of each depth, compacted (order inside a depth is irrelevant: every node sums its own children in
// fixed octant order).  Counts per depth come from k_links.
__global__ void __launch_bounds__(TPB) k_level_lists(AgbDev d, AgbScalars* s, int32_t* __restrict__ list)
{
    __shared__ int cnt[AGB_MAX_LEVELS + 1], basep[AGB_MAX_LEVELS + 1];
    if (threadIdx.x <= AGB_MAX_LEVELS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int nn = s->n_nodes;
    if (s->node_overflow) return;
    const int k = blockIdx.x * TPB + threadIdx.x;
    int dep = -1, r = 0;
    if (k < nn) { dep = d.ndepth[k]; r = atomicAdd(&cnt[dep], 1); }
    __syncthreads();
    if (threadIdx.x <= AGB_MAX_LEVELS && cnt[threadIdx.x]) {
        int off = 0;
        for (int q = 0; q < (int)threadIdx.x; q++) off += s->lvl_cnt[q];       // first slot of this depth
        basep[threadIdx.x] = off + atomicAdd(&s->lvl_cur[threadIdx.x], cnt[threadIdx.x]);
    }
    __syncthreads();
    if (k < nn) list[basep[dep] + r] = k;
}

// all blocks of a co-resident grid wait for each other (bar starts at 0; epoch counts arrivals expected so far)
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

// Upward pass, level by level from the deepest internal nodes to the root: one thread per node of the level, children of
// deeper levels are complete (grid-wide barrier between levels; the grid is sized to be co-resident).  No atomics, no
// fences per node, full warps.  Then the reference's root quirk and the normalisation (COM, mVel, duplication flags).
template <int MODE> __device__ __forceinline__ void root_fix_block(const AgbDev& d, const AgbScalars* s);
template <int MODE> __device__ __forceinline__ void finalize_node(const AgbDev& d, const AgbScalars* s, int k);

template <int MODE>
__global__ void __launch_bounds__(TPB) k_upward_levels(AgbDev d, AgbScalars* s, const int32_t* __restrict__ list)
{
    const int nn = s->n_nodes;
    if (s->node_overflow || nn < 1) return;
    if (MODE == 2 && !s->any_gas) return;
    const int N = (int)d.n;
    const bool any_gas = s->any_gas != 0;
    unsigned int* const bar = MODE == 2 ? &s->grid_bar2 : &s->grid_bar;
    unsigned int epoch = 0;
    int off_end = nn;                                         // list segment of depth `dep` is [off_end - cnt, off_end)
    int maxd = AGB_MAX_LEVELS;
    while (maxd > 0 && s->lvl_cnt[maxd] == 0) maxd--;
    for (int q = maxd + 1; q <= AGB_MAX_LEVELS; q++) off_end -= s->lvl_cnt[q];
    for (int dep = maxd; dep >= 0; dep--) {
        const int cnt = s->lvl_cnt[dep];
        const int beg = off_end - cnt;
        for (int idx = beg + blockIdx.x * TPB + threadIdx.x; idx < off_end; idx += gridDim.x * TPB) {
            const int k = list[idx];
            node_moments<MODE>(d, k, N, any_gas, load_links(d.child, k));
        }
        off_end = beg;
        if (cnt > 0 || dep == 0) grid_sync(bar, gridDim.x, epoch);
    }
    if (blockIdx.x == 0) root_fix_block<MODE>(d, s);
    grid_sync(bar, gridDim.x, epoch);
    for (int k = blockIdx.x * TPB + threadIdx.x; k < nn; k += gridDim.x * TPB) finalize_node<MODE>(d, s, k);
}

// Reference quirk: bulk insertion accumulates the ROOT's mass / COM / gasMass / mVel over ALL particles,
// including those outside the cube that are never inserted (Node.cpp:477-499 precede the octant test).
template <int MODE>
__device__ __forceinline__ void root_fix_block(const AgbDev& d, const AgbScalars* s)
{
    if (d.n < (int64_t)d.cores * 100) return;                  // one-by-one insertion rejects them at the root (Node.cpp:606-612)
    __shared__ double sh[8][TPB / 32];
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = s->n_in_tree + threadIdx.x; i < d.n; i += TPB) {
        double4 a = d.src_pm[i], b = d.src_gv[i];
        v[3] += a.w; v[0] += a.x * a.w; v[1] += a.y * a.w; v[2] += a.z * a.w;
        v[7] += b.w; v[4] += b.x * b.w; v[5] += b.y * b.w; v[6] += b.z * b.w;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double x = warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[8];
        for (int k = 0; k < 8; k++) { t[k] = 0; for (int w = 0; w < TPB / 32; w++) t[k] += sh[k][w]; }
        double4 pm = ldcg4(&d.mom_pm[0]), gv = s->any_gas ? ldcg4(&d.mom_gv[0]) : make_double4(0, 0, 0, 0);
        if (MODE != 2) { pm.x += t[0]; pm.y += t[1]; pm.z += t[2]; pm.w += t[3]; gv.w += t[7]; }
        if (MODE != 1) { gv.x += t[4]; gv.y += t[5]; gv.z += t[6]; }
        if (MODE != 2) d.mom_pm[0] = pm;
        d.mom_gv[0] = gv;
    }
}

template <int MODE>
__device__ __forceinline__ void finalize_node(const AgbDev& d, const AgbScalars* s, int k)
{
    const int64_t N = d.n;
    double4 gv = s->any_gas ? ldcg4(&d.mom_gv[k]) : make_double4(0, 0, 0, 0);
    double4 mv = make_double4(0, 0, 0, gv.w);
    if (gv.w > 0.0) { mv.x = gv.x / gv.w; mv.y = gv.y / gv.w; mv.z = gv.z / gv.w; }
    if (MODE == 2) { if (gv.w > 0.0) d.src_gv[N + k] = mv; return; }
    double4 pm = ldcg4(&d.mom_pm[k]);
    double4 com = make_double4(0, 0, 0, pm.w);
    if (pm.w > 0.0) { com.x = pm.x / pm.w; com.y = pm.y / pm.w; com.z = pm.z / pm.w; }
    d.src_pm[N + k] = com;
    d.src_gv[N + k] = mv;
    d.src_flag[N + k] = gv.w > 0.0 ? 1 : 0;
    // nodes where bulk insertion (>= cores*100 particles) hands over to one-by-one insertion
    const int64_t thr = (int64_t)d.cores * 100;
    int par = d.nparent[k];
    uint8_t dup = 0;
    if (par >= 0) {
        int64_t cnt = (int64_t)__ldcg(&d.nlast[k]) - d.nfirst[k] + 1;
        int64_t pcnt = par == 0 ? N : (int64_t)__ldcg(&d.nlast[par]) - d.nfirst[par] + 1;   // the root is handed all N particles
        dup = cnt < thr && pcnt >= thr;
    }
    d.ndup[k] = dup;
}

// ------------------------------------------------------------------ introspection
__global__ void __launch_bounds__(TPB) k_dump_tree(AgbDev d, const uint64_t* __restrict__ khi, const uint64_t* __restrict__ klo, const uint32_t* __restrict__ perm,
                                                     const AgbScalars* __restrict__ s, int32_t* leafdepth, uint64_t* ohi, uint64_t* olo)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n) return;
    uint32_t p = perm[i];
    if (i >= s->n_in_tree) { leafdepth[p] = -1; ohi[p] = 0; olo[p] = 0; return; }
    int ld = d.leafdepth[i];
    leafdepth[p] = ld;
    uint64_t h = khi[i], l = klo[i];
    // keep only the first `ld` levels of the path
    if (ld <= 0) { h = 0; l = 0; }
    else if (ld <= 21) { h &= ~0ull << (63 - 3 * ld); l = 0; }
    else if (ld < 42) { l &= ~0ull << (63 - 3 * (ld - 21)); }      // deeper leaves: all 21 levels of key_lo belong to the path
    ohi[p] = h & ~AGB_OUTLIER_BIT; olo[p] = l;
}

} // namespace

// ====================================================================== launchers
static inline int nblk(int64_t n, int per) { return (int)((n + per - 1) / per); }

int agb_launch_fill_mass(const AgbDev& d, cudaStream_t st)
{
    k_fill_mass<<<nblk(d.n, TPB), TPB, 0, st>>>(d.mass, d.n, d.rec);
    return 1;
}

int agb_launch_extent(const AgbDev& d, AgbScalars* s, cudaStream_t st, bool mass_late)
{
    int nb = (int)std::min<int64_t>(1024, std::max<int64_t>(1, nblk(d.n, TPB)));
    k_dist_partial<<<nb, TPB, 0, st>>>(d.x, d.y, d.z, mass_late ? nullptr : d.mass, d.n, d.rec, s);
    k_extent_finish<<<1, 32, 0, st>>>(s, nb, d.n);
    k_extent_max<<<nb, TPB, 0, st>>>(d.rec, d.n, s);
    return 3;
}

template <int ITEMS>
static int sort_passes(AgbDev& d, cudaStream_t st)
{
    constexpr int TILE = SortCfg<ITEMS>::TILE;
    const int nb = nblk(d.n, TILE);
    const int smem = TILE * 12 + (TPB / 32) * 256 * 4 + 2 * 256 * 4;
    cudaFuncSetAttribute(k_sort_onesweep<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device; cheap, so not cached
    uint32_t* ghist = d.blockhist;                           // [8][256] digit totals (k_keygen), [8] tickets, then status [8][nb][256]
    uint32_t* ticket = ghist + 8 * 256;
    uint32_t* status = ticket + 64;
    for (int pass = 0; pass < 8; pass++) {
        const int in = d.cur, out = d.cur ^ 1;
        k_sort_onesweep<ITEMS><<<nb, TPB, smem, st>>>(d.khi[in], d.perm[in], d.khi[out], d.perm[out], d.n, pass, ghist, status, ticket);
        d.cur = out;
    }
    return 8;
}

// items per thread of the sort tiles for n keys (the status area is sized for the smaller tile)
static inline int sort_items(int64_t n) { return n >= (4 << 20) ? 16 : 8; }
size_t agb_sort_scratch_words(int64_t cap) { return 8 * 256 + 64 + 8 * (size_t)((cap + 2047) / 2048 + 1) * 256; }

int agb_launch_keygen(AgbDev& d, AgbScalars* s, cudaStream_t st)
{
    // digit totals + tickets + status words of the 8 sort passes start at zero
    const size_t nb = (size_t)nblk(d.n, TPB * sort_items(d.n));
    cudaMemsetAsync(d.blockhist, 0, (8 * 256 + 64 + 8 * nb * 256) * sizeof(uint32_t), st);
    d.cur = 0;
    if (d.deep) { k_keygen_deep<<<nblk(d.n, TPB), TPB, 0, st>>>(d.rec, d.type, d.n, d.dk[0], d.dk[1], d.dk[2], d.perm[0], s); return 1; }
    k_keygen<<<std::min(nblk(d.n, TPB), 148 * 16), TPB, 0, st>>>(d.rec, d.type, d.n, d.khi[0], d.perm[0], s, d.blockhist);
    return 1;
}

int agb_launch_sort(AgbDev& d, AgbScalars* s, cudaStream_t st)
{
    const int nb = nblk(d.n, TPB);
    if (d.deep) {
        // three stable 8-pass sorts, least significant word first: levels 42..62, then 21..41, then the flag + levels 0..20
        int launches = 0;
        const size_t snb = (size_t)nblk(d.n, TPB * sort_items(d.n));
        for (int w = 2; w >= 0; w--) {
            if (w == 2) cudaMemcpyAsync(d.khi[d.cur], d.dk[2], (size_t)d.n * 8, cudaMemcpyDeviceToDevice, st);   // caller order = the identity permutation
            else { k_gather_keys<<<nb, TPB, 0, st>>>(d.dk[w], d.perm[d.cur], d.n, d.khi[d.cur]); launches++; }
            cudaMemsetAsync(d.blockhist, 0, (8 * 256 + 64 + 8 * snb * 256) * sizeof(uint32_t), st);
            k_hist64<<<std::min(nb, 8 * 148), TPB, 0, st>>>(d.khi[d.cur], d.n, d.blockhist); launches++;
            launches += sort_items(d.n) == 16 ? sort_passes<16>(d, st) : sort_passes<8>(d, st);
        }
        k_gather_keys<<<nb, TPB, 0, st>>>(d.dk[1], d.perm[d.cur], d.n, d.klo[1]);
        k_gather_keys<<<nb, TPB, 0, st>>>(d.dk[2], d.perm[d.cur], d.n, d.kex);
        return launches + 2;
    }
    // key_hi + caller index: 8 passes; the result lands in half 0 again.  key_lo (tree order, klo[1]) is zero except
    // inside runs of equal key_hi, where it is computed on demand and decides the order.
    int launches = sort_items(d.n) == 16 ? sort_passes<16>(d, st) : sort_passes<8>(d, st);
    cudaMemsetAsync(d.klo[1], 0, (size_t)d.n * sizeof(uint64_t), st);
    k_fix_runs<<<nb, TPB, 0, st>>>(d.khi[d.cur], d.klo[1], d.perm[d.cur], d.n, d.rec, s, d.nodecnt);
    k_fix_long_runs<<<64, TPB, 0, st>>>(d.khi[d.cur], d.klo[1], d.perm[d.cur], d.n, d.rec, s, d.nodecnt);
    return launches + 2;
}

static int upward_blocks(int nnb)
{
    static int occ = 0;                                       // co-resident blocks per SM of the level kernels (same on every B200)
    if (!occ) {
        int o0 = 0, o1 = 0, o2 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o0, k_upward_levels<0>, TPB, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_upward_levels<1>, TPB, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_upward_levels<2>, TPB, 0);
        occ = std::max(1, std::min(std::min(o0, o1), std::min(o2, 8)));
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return std::min(sms * occ, std::max(1, nnb));
}

// late_gas: velocities / U / mu of the hand-over are still on their way (agb_launch_late_gas completes the tree afterwards)
int agb_launch_links(AgbDev& d, AgbScalars* s, cudaStream_t st, cudaEvent_t* ev, bool late_gas)
{
    const int nb = nblk(d.n, TPB);
    const uint64_t *khi = d.khi[d.cur], *klo = d.klo[1], *kex = d.deep ? d.kex : nullptr;
    if (d.next && !late_gas) k_next_uniform<<<std::min(nb, 2048), TPB, 0, st>>>(d.next, d.n, s);
    static const bool split = !getenv("AGB_GATHER_SPLIT") || atoi(getenv("AGB_GATHER_SPLIT")) != 0;
    if (late_gas) k_gather<true><<<nb, TPB, 0, st>>>(d, d.perm[d.cur], s, false);
    else if (split) {
        k_pack_gas<<<nb, TPB, 0, st>>>(d);
        k_gather<true><<<nb, TPB, 0, st>>>(d, d.perm[d.cur], s, true);
        k_gather_gas<<<nb, TPB, 0, st>>>(d, d.perm[d.cur]);
    } else {
        k_pack_gas<<<nb, TPB, 0, st>>>(d);
        k_gather<false><<<nb, TPB, 0, st>>>(d, d.perm[d.cur], s, false);
    }
    if (ev) cudaEventRecord(ev[0], st);
    k_lcp<<<nb, TPB, 0, st>>>(khi, klo, kex, d.n, d.lcp, d.nodecnt, s);
    const int sb = nblk(d.n, SCAN_TILE);
    k_scan_reduce<<<sb, TPB, 0, st>>>(d.nodecnt, d.n, d.scanblk, nullptr);
    k_scan_blocks<<<1, TPB, 0, st>>>(d.scanblk, sb, &s->n_nodes, nullptr, d.n);
    k_scan_apply<<<sb, TPB, 0, st>>>(d.nodecnt, d.n, d.scanblk, d.nodebase, nullptr);
    const int nnb = nblk(d.ncap, TPB);                        // node kernels: one thread per node slot
    k_init_nodes<<<nnb, TPB, 0, st>>>(d, s);
    k_links<<<nb, TPB, 0, st>>>(d, khi, klo, kex, s);
    if (ev) cudaEventRecord(ev[1], st);
    // level lists, then the persistent level-by-level upward pass
    k_level_lists<<<nnb, TPB, 0, st>>>(d, s, d.lvl_list);
    if (late_gas) k_upward_levels<1><<<upward_blocks(nnb), TPB, 0, st>>>(d, s, d.lvl_list);
    else k_upward_levels<0><<<upward_blocks(nnb), TPB, 0, st>>>(d, s, d.lvl_list);
    return (late_gas ? 10 : split ? 12 : 11) + (d.next && !late_gas ? 1 : 0);
}

int agb_launch_gather_next(AgbDev& d, AgbScalars* s, cudaStream_t st)
{
    const int nb = nblk(d.n, TPB);
    if (d.next) k_next_uniform<<<std::min(nb, 2048), TPB, 0, st>>>(d.next, d.n, s);
    k_gather_next<<<nb, TPB, 0, st>>>(d, d.perm[d.cur], s);
    return d.next ? 2 : 1;
}

// second half of a late_gas build, after the gas densities: gas velocities / U / mu into tree order, P and T of the density
// groups, mVel of the nodes
int agb_launch_late_gas(AgbDev& d, AgbScalars* s, cudaStream_t st)
{
    const int nb = nblk(d.n, TPB), nnb = nblk(d.ncap, TPB);
    k_pack_gas<<<nb, TPB, 0, st>>>(d);
    k_gather_late<<<nb, TPB, 0, st>>>(d, d.perm[d.cur], s);
    k_upward_levels<2><<<upward_blocks(nnb), TPB, 0, st>>>(d, s, d.lvl_list);
    return 3;
}

// skip_if_n (device, optional): the three kernels return at once when *skip_if_n == n (nothing to compact)
int agb_launch_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* blk, int32_t* total_out, cudaStream_t st, const int32_t* skip_if_n)
{
    const int sb = nblk(n, SCAN_TILE);
    k_scan_reduce<<<sb, TPB, 0, st>>>(in, n, blk, skip_if_n);
    k_scan_blocks<<<1, TPB, 0, st>>>(blk, sb, total_out, skip_if_n, n);
    k_scan_apply<<<sb, TPB, 0, st>>>(in, n, blk, out, skip_if_n);
    return 3;
}

int agb_launch_dump_tree(AgbDev& d, AgbScalars* s, int32_t* leafdepth, uint64_t* khi, uint64_t* klo, cudaStream_t st)
{
    k_dump_tree<<<nblk(d.n, TPB), TPB, 0, st>>>(d, d.khi[d.cur], d.klo[1], d.perm[d.cur], s, leafdepth, khi, klo);
    return 1;
}

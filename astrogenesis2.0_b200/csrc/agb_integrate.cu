// agb_integrate.cu — device-resident form of the reference's driver loop around the force path:
// time-step binning (Physics/Simulation.cpp:196-207, :222-232), the min-nextIntegrationTime reduction (:237-254),
// Kick / Drift / Ueuler (Physics/TimeIntegration.cpp:10-41), the Hubble rescale (Simulation.cpp:328-332) and
// `nextIntegrationTime += timeStep` (:336).  Every expression is evaluated with explicitly rounded IEEE operations in
// the reference's order, so given the same accelerations the trajectory is bit-identical to the host loop; the point is
// that positions, velocities and results never leave HBM between steps.
#include "agb_internal.cuh"
#include <algorithm>

namespace {

constexpr int TPB = 256;
constexpr double kGAMMA = 5.0 / 3.0, kKB = 1.38064852e-23, kPRTN = 1.6726219e-27;

// floor(log2(t)) exactly as the reference's `std::floor(std::log2(t))` (Simulation.cpp:199-202) evaluates it: libm's log2 is
// correctly rounded here, so for t a few ulps below a power of two 2^k the value k - eps/ln2 rounds UP to k and the
// reference picks 2^k, not 2^(k-1).  frexp gives the mathematical floor; the correction applies when the distance of
// log2(t) to k is below half an ulp of k (ulp of the binade log2(t) lies in).
__device__ __forceinline__ int floor_log2_like_libm(double t)
{
    int ex;
    const double f = frexp(t, &ex);                                 // t = f 2^ex, f in [0.5, 1): floor(log2 t) = ex - 1
    if (ex == 0) return -1;                                         // t just below 1: log2(t) is a tiny negative number, floor = -1
    const double eps = __dadd_rn(1.0, -f);                          // exact (Sterbenz); relative gap to the next power of two
    const int ak = ex < 0 ? -ex : ex;
    int lg = 31 - __clz(ak);                                        // binade of |k|
    if (ex > 0 && (ak & (ak - 1)) == 0) lg -= 1;                    // k - delta lies in the binade below a positive power of two
    const double half_ulp = ldexp(1.0, lg - 53);
    // log2(1 - eps) = -(eps + eps^2/2 + ..) / ln 2
    const double dist = eps * (1.0 + 0.5 * eps) * 1.4426950408889634;
    return dist < half_ulp ? ex : ex - 1;
}

__device__ __forceinline__ void assign_timestep(const AgbDev& d, const AgbInt& I, int64_t i, double gt)
{
    const double ax = d.ax[i], ay = d.ay[i], az = d.az[i];
    const double a = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));   // vec3::length
    double ts = I.min_ts;
    if (a > 0.0) {
        double t = __dmul_rn(I.eta, __dsqrt_rn(__ddiv_rn(I.e0, a)));
        t = fmin(fmax(t, I.min_ts), I.max_ts);                      // std::clamp
        ts = fmax(ldexp(1.0, floor_log2_like_libm(t)), I.min_ts);
    }
    I.timestep[i] = ts;
    I.next[i] = __dadd_rn(gt, ts);
}

__global__ void __launch_bounds__(TPB) k_int_init(AgbDev d, AgbInt I)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n) return;
    if (d.type[i] == 2) d.T[i] = (kGAMMA - 1.0) * I.U[i] * kPRTN * (I.mu ? I.mu[i] : 0.58) / kKB;   // Simulation.cpp:108-112 (same product order)
    I.next[i] = 0.0;
}

__global__ void __launch_bounds__(TPB) k_int_assign(AgbDev d, AgbInt I, double gt, int all)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n) return;
    if (all || gt >= I.next[i]) assign_timestep(d, I, i, gt);
}

__global__ void __launch_bounds__(TPB) k_int_min(AgbDev d, AgbInt I, unsigned long long* out)
{
    __shared__ double sh[TPB / 32];
    double m = __longlong_as_double(0x7fefffffffffffffll);
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * TPB) m = fmin(m, I.next[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < TPB / 32; k++) m = fmin(m, sh[k]);
        atomicMin(out, (unsigned long long)__double_as_longlong(m));     // times are >= 0: bit order == value order
    }
}

__device__ __forceinline__ void kick(const AgbDev& d, const AgbInt& I, int64_t i, double dt)
{
    const double ax = d.ax[i], ay = d.ay[i], az = d.az[i];
    if (isnan(ax) || isnan(ay) || isnan(az)) return;                     // TimeIntegration.cpp:12-16
    I.vx[i] = __dadd_rn(I.vx[i], __dmul_rn(ax, dt) / 2);
    I.vy[i] = __dadd_rn(I.vy[i], __dmul_rn(ay, dt) / 2);
    I.vz[i] = __dadd_rn(I.vz[i], __dmul_rn(az, dt) / 2);
}

__global__ void __launch_bounds__(TPB) k_int_first(AgbDev d, AgbInt I, double gt)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n || I.next[i] != gt) return;
    const double dt = I.timestep[i];
    kick(d, I, i, dt);
    I.x[i] = __dadd_rn(I.x[i], __dmul_rn(I.vx[i], dt));                  // Drift
    I.y[i] = __dadd_rn(I.y[i], __dmul_rn(I.vy[i], dt));
    I.z[i] = __dadd_rn(I.z[i], __dmul_rn(I.vz[i], dt));
}

__global__ void __launch_bounds__(TPB) k_int_second(AgbDev d, AgbInt I, double gt)
{
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= d.n || I.next[i] != gt) return;
    const double dt = I.timestep[i];
    int ex;
    frexp(dt, &ex);
    const int bin = min(max(ex - 1 - I.k0, 0), AGB_INT_BINS - 1);
    if (I.type[i] == 2) {
        const double T = d.T[i], rho = d.rho[i];
        if (I.cooling) {                                                 // Cooling::coolingRoutine, Cooling.cpp:6-25 (free-free emission)
            const double rate_cgs = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(1.42e-27, 1.1), __dsqrt_rn(T)), 1e6), 1e6);
            const double rate = __dmul_rn(rate_cgs, 1e-7);
            if (rate > 0.0 && rho > 0.0) d.dUdt[i] = __dadd_rn(d.dUdt[i], -__ddiv_rn(rate, rho));
        }
        if (I.star_formation && rho > 1e-22 && T < 1e4) {                // SFR::sfrRoutine, SFR.cpp:12-34
            const double p = dt == I.min_ts ? I.sf_min : I.sf_tab[bin];
            I.sfr[i] = p;
            if (agb_u01(I.seed, (unsigned long long)i, gt) < p) { I.type[i] = 1; I.U[i] = 0.0; }   // gas -> star
        }
    }
    if (I.type[i] == 2) {                                                // Ueuler (Simulation.cpp:322-326)
        const double du = d.dUdt[i];
        if (!isnan(du)) I.U[i] = __dadd_rn(I.U[i], __dmul_rn(du, dt));
        d.dUdt[i] = 0.0;
    }
    // exp(H0 dt) comes from a host (libm) table over the power-of-two bins, so the factor is the reference's bit for bit
    const double scale = dt == I.min_ts ? I.scale_min : I.scale_tab[bin];
    I.x[i] = __dmul_rn(I.x[i], scale); I.y[i] = __dmul_rn(I.y[i], scale); I.z[i] = __dmul_rn(I.z[i], scale);
    kick(d, I, i, dt);
    I.next[i] = __dadd_rn(I.next[i], dt);
}

} // namespace

static inline int nblk(int64_t n, int per) { return (int)((n + per - 1) / per); }

int agb_launch_int_init(AgbDev& d, const AgbInt& I, cudaStream_t st)
{
    k_int_init<<<nblk(d.n, TPB), TPB, 0, st>>>(d, I);
    return 1;
}
int agb_launch_int_assign(AgbDev& d, const AgbInt& I, double gt, bool all, cudaStream_t st)
{
    k_int_assign<<<nblk(d.n, TPB), TPB, 0, st>>>(d, I, gt, all ? 1 : 0);
    return 1;
}
int agb_launch_int_min(AgbDev& d, const AgbInt& I, unsigned long long* out, cudaStream_t st)
{
    cudaMemsetAsync(out, 0xff, sizeof(unsigned long long), st);
    k_int_min<<<std::min(nblk(d.n, TPB), 1024), TPB, 0, st>>>(d, I, out);
    return 1;
}
int agb_launch_int_first(AgbDev& d, const AgbInt& I, double gt, cudaStream_t st)
{
    k_int_first<<<nblk(d.n, TPB), TPB, 0, st>>>(d, I, gt);
    return 1;
}
int agb_launch_int_second(AgbDev& d, const AgbInt& I, double gt, cudaStream_t st)
{
    k_int_second<<<nblk(d.n, TPB), TPB, 0, st>>>(d, I, gt);
    return 1;
}

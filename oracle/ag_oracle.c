/* oracle/ag_oracle.c — CPU restatement of the reference force path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (astrogenesis2.0_b200/csrc) never links, imports or falls back to it.
 *
 * What it restates (reference = Philip-Spaeth/AstroGenesis2.0, paths relative to simulation/src):
 *   root extent        Physics/Tree/Tree.cpp:86-117      (calcTreeWidth)
 *   octree build       Physics/Tree/Node.cpp:405-534     (bulk insert), :597-699 (single insert),
 *                      :702-719 (getOctant), Tree.cpp:24-55
 *   visual density     Tree.cpp:152-176, Node.cpp:832-877
 *   group gas density  Tree.cpp:119-150, Node.cpp:722-796, Math/kernel.cpp:4-16
 *   force walk         Tree.cpp:57-83, Node.cpp:247-399 (gravity), :88-172 (in-walk SPH),
 *                      Math/kernel.cpp:18-39 (grad W), :41-56 (softening kernel), Math/vec3.cpp
 * with the reference's SERIAL semantics (OpenMP pragmas ignored) and an explicit `cores` argument
 * standing in for omp_get_max_threads() in the `size < cores*100` switch (Node.cpp:420).
 * Every floating-point expression is written with the reference's operation order and the file is
 * compiled with -ffp-contract=off, so the outputs are bit-identical to oracle/_ref/ag_ref (the
 * unmodified reference compiled in place); tests/test_oracle_pin.py asserts exactly that on the
 * four example ICs shipped by the reference and on synthetic sets.  The reference has no tests or
 * golden vectors of its own, so that binary is the pin.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AG_G      6.67430e-11              /* Math/Constants.h:7  */
#define AG_PI     3.14159265358979323846   /* Math/Constants.h:10 */
#define AG_GAMMA  (5.0 / 3.0)              /* Math/Constants.h:15 */
#define AG_KB     1.38064852e-23           /* Math/Constants.h:16 */
#define AG_PRTN   1.6726219e-27            /* Math/Constants.h:18 */

typedef struct Node {
    double cx, cy, cz, radius;             /* cell centre ("position") and half-width */
    double comx, comy, comz, mass, gasMass, mvx, mvy, mvz;
    int depth, isLeaf;
    int64_t particle;                      /* leaf payload, -1 = none */
    struct Node* child[8];
    struct Node* parent;
    int64_t* list;                         /* childParticles */
    int64_t nlist, cap;
} Node;

typedef struct {
    int64_t n;
    const double *x, *y, *z, *vx, *vy, *vz, *m, *U, *next_time, *mu;
    const uint8_t* type;
    double *rho, *P, *T;                   /* carried over for orphans (in/out) */
    double *ax, *ay, *az, *dUdt, *h, *vis;
    Node** leaf;                           /* Particle::node */
    Node* root;
    int64_t nnodes;
    /* pool */
    Node** blocks; int64_t nblocks, used_in_block;
    /* per-target counters */
    int32_t *visits, *acc_nodes, *acc_leaves, *sph;
} Oracle;

enum { BLOCK = 1 << 16 };

static Node* new_node(Oracle* o)
{
    if (o->nblocks == 0 || o->used_in_block == BLOCK) {
        o->blocks = (Node**)realloc(o->blocks, (size_t)(o->nblocks + 1) * sizeof(Node*));
        o->blocks[o->nblocks++] = (Node*)malloc(sizeof(Node) * BLOCK);
        o->used_in_block = 0;
    }
    Node* n = &o->blocks[o->nblocks - 1][o->used_in_block++];
    memset(n, 0, sizeof(Node));
    n->isLeaf = 1;                          /* Node.cpp:13-29 */
    n->particle = -1;
    o->nnodes++;
    return n;
}

static void list_push(Node* n, int64_t p)
{
    if (n->nlist == n->cap) {
        n->cap = n->cap ? 2 * n->cap : 4;
        n->list = (int64_t*)realloc(n->list, (size_t)n->cap * sizeof(int64_t));
    }
    n->list[n->nlist++] = p;
}

static void make_children(Oracle* o, Node* n)
{                                          /* Node.cpp:433-443 and :630-642 */
    for (int i = 0; i < 8; i++) {
        Node* c = new_node(o);
        c->cx = n->cx + n->radius * ((i & 1) ? 0.5 : -0.5);
        c->cy = n->cy + n->radius * ((i & 2) ? 0.5 : -0.5);
        c->cz = n->cz + n->radius * ((i & 4) ? 0.5 : -0.5);
        c->radius = n->radius / 2;
        c->depth = n->depth + 1;
        c->parent = n;
        n->child[i] = c;
    }
}

static int out_of_bounds(const Oracle* o, const Node* n, int64_t p)
{
    return o->x[p] < n->cx - n->radius || o->x[p] > n->cx + n->radius ||
           o->y[p] < n->cy - n->radius || o->y[p] > n->cy + n->radius ||
           o->z[p] < n->cz - n->radius || o->z[p] > n->cz + n->radius;
}

static int get_octant(const Oracle* o, const Node* n, int64_t p)
{                                          /* Node.cpp:702-719: strict '>' */
    if (out_of_bounds(o, n, p)) return -1;
    int oct = 0;
    if (o->x[p] > n->cx) oct |= 1;
    if (o->y[p] > n->cy) oct |= 2;
    if (o->z[p] > n->cz) oct |= 4;
    return oct;
}

static void gas_velocity_update(const Oracle* o, Node* n, int64_t p)
{                                          /* Node.cpp:480-483 / :681-685 (running mass-weighted mean) */
    double mp = o->m[p];
    n->gasMass += mp;
    if (n->gasMass > 0) {
        double w = n->gasMass - mp;
        n->mvx = (n->mvx * w + o->vx[p] * mp) / n->gasMass;
        n->mvy = (n->mvy * w + o->vy[p] * mp) / n->gasMass;
        n->mvz = (n->mvz * w + o->vz[p] * mp) / n->gasMass;
    }
}

static void insert_one(Oracle* o, Node* n, int64_t p)
{                                          /* Node.cpp:597-699 */
    if (out_of_bounds(o, n, p)) return;
    list_push(n, p);
    if (n->isLeaf) {
        if (n->particle < 0) {
            n->particle = p;
            o->leaf[p] = n;
        } else {
            make_children(o, n);
            int oct = get_octant(o, n, n->particle);
            if (oct != -1) insert_one(o, n->child[oct], n->particle);
            oct = get_octant(o, n, p);
            if (oct != -1) insert_one(o, n->child[oct], p);
            n->isLeaf = 0;
            n->particle = -1;
        }
    } else {
        int oct = get_octant(o, n, p);
        if (oct != -1) insert_one(o, n->child[oct], p);
    }
    double mp = o->m[p];
    n->mass += mp;
    if (o->type[p] == 2) gas_velocity_update(o, n, p);
    double w = n->mass - mp;
    n->comx = (n->comx * w + o->x[p] * mp) / n->mass;
    n->comy = (n->comy * w + o->y[p] * mp) / n->mass;
    n->comz = (n->comz * w + o->z[p] * mp) / n->mass;
}

static void insert_bulk(Oracle* o, Node* n, const int64_t* ps, int64_t np, int cores)
{                                          /* Node.cpp:405-534, one thread */
    if (np == 0) return;
    if (np == 1) {
        int64_t p = ps[0];
        n->isLeaf = 1;
        n->particle = p;
        o->leaf[p] = n;
        n->comx = o->x[p]; n->comy = o->y[p]; n->comz = o->z[p];
        n->mass = o->m[p];
        n->gasMass = (o->type[p] == 2) ? o->m[p] : 0.0;
        return;
    }
    if ((uint64_t)np < (uint64_t)(int64_t)(cores * 100)) {
        /* ps may alias n->list, which insert_one grows: iterate over a copy like the by-value
         * parameter of the reference does (Node.h:17). */
        int64_t* copy = (int64_t*)malloc((size_t)np * sizeof(int64_t));
        memcpy(copy, ps, (size_t)np * sizeof(int64_t));
        for (int64_t i = 0; i < np; i++) insert_one(o, n, copy[i]);
        free(copy);
        return;
    }
    n->isLeaf = 0;
    make_children(o, n);
    double tm = 0.0, tg = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
    for (int64_t i = 0; i < np; i++) {
        int64_t p = ps[i];
        if (n->depth == 0) o->leaf[p] = NULL;
        tm += o->m[p];
        if (o->type[p] == 2) {
            tg += o->m[p];
            gas_velocity_update(o, n, p);
        }
        sx += o->x[p] * o->m[p];
        sy += o->y[p] * o->m[p];
        sz += o->z[p] * o->m[p];
        int oct = get_octant(o, n, p);
        if (oct != -1) list_push(n->child[oct], p);
    }
    n->mass = tm;
    n->gasMass = tg;
    if (n->mass > 0.0) { n->comx = sx / n->mass; n->comy = sy / n->mass; n->comz = sz / n->mass; }
    for (int i = 0; i < 8; i++) {
        Node* c = n->child[i];
        if (c->nlist > 0) {
            int64_t* copy = (int64_t*)malloc((size_t)c->nlist * sizeof(int64_t));
            int64_t cn = c->nlist;
            memcpy(copy, c->list, (size_t)cn * sizeof(int64_t));
            insert_bulk(o, c, copy, cn, cores);
            free(copy);
        }
    }
}

static double tree_width(const Oracle* o)
{                                          /* Tree.cpp:86-117 */
    int64_t n = o->n;
    if (n == 0) return 0;
    double* d = (double*)malloc((size_t)n * sizeof(double));
    double sum = 0.0, sq = 0.0;
    for (int64_t i = 0; i < n; i++) d[i] = sqrt(o->x[i] * o->x[i] + o->y[i] * o->y[i] + o->z[i] * o->z[i]);
    for (int64_t i = 0; i < n; i++) sum = sum + d[i];
    double mean = sum / (double)(int)n;
    for (int64_t i = 0; i < n; i++) sq = sq + d[i] * d[i];
    double stdev = sqrt(sq / (double)(int)n - mean * mean);
    double lim = mean + 10 * stdev;
    double mx = 0;
    for (int64_t i = 0; i < n; i++) if (d[i] <= lim && d[i] > mx) mx = d[i];
    free(d);
    return mx;
}

/* ---- kernels (Math/kernel.cpp) ---- */
static double spline_w(double r, double h)
{                                          /* kernel.cpp:4-16 */
    double a = 1.0 / (AG_PI * h * h * h);
    double q = r / h;
    if (q < 1.0) return a * (1 - 1.5 * q * q + 0.75 * q * q * q);
    else if (q < 2.0) return a * 0.25 * pow(2 - q, 3);
    return 0.0;
}

static void spline_grad(double dx, double dy, double dz, double h, double g[3])
{                                          /* kernel.cpp:18-39 */
    double rn = sqrt(dx * dx + dy * dy + dz * dz);
    g[0] = g[1] = g[2] = 0.0;
    if (rn == 0.0) return;
    const double a = 1.0 / (AG_PI * h * h * h);
    double q = rn / h, s;
    if (q < 1.0) s = a * (-3.0 * q + 2.25 * q * q);
    else if (q < 2.0) s = a * (-0.75 * pow(2 - q, 2));
    else return;
    s /= h;
    double f = s / rn;
    g[0] = dx * f; g[1] = dy * f; g[2] = dz * f;
}

static double softening_kernel(double u)
{                                          /* kernel.cpp:41-56 */
    if (u >= 0 && u < 1.0 / 2.0)
        return (16.0 / 3.0) * pow(u, 2) - (48.0 / 5.0) * pow(u, 4) + (32.0 / 5.0) * pow(u, 5) - (14.0 / 5.0);
    else if (u >= 1.0 / 2.0 && u < 1.0)
        return (1.0 / (15.0 * u)) + (32.0 / 3.0) * pow(u, 2) - (16.0) * pow(u, 3) - (48.0 / 5.0) * pow(u, 4) - (32.0 / 15.0) * pow(u, 5) - (16.0 / 5.0);
    else if (u >= 1.0)
        return -1.0 / u;
    return 0.0;
}

/* ---- visual density (Node.cpp:832-877) ---- */
static void node_visual_density(Oracle* o, Node* n, double rho_t)
{
    if (n->parent == NULL) return;
    double dr = rho_t - n->radius;
    double dpr = rho_t - n->parent->radius;
    if (fabs(dr) > fabs(dpr)) { node_visual_density(o, n->parent, rho_t); return; }
    double vol = n->radius * n->radius * n->radius;
    if (vol == 0 || n->mass == 0) return;
    double dens = n->mass / vol;
    if (dens == 0 || dens == INFINITY) return;
    for (int64_t i = 0; i < n->nlist; i++) o->vis[n->list[i]] = dens;
}

/* ---- group gas density (Node.cpp:722-796) ---- */
static void node_gas_density(Oracle* o, Node* n, double massInH)
{
    if (n->gasMass == 0) return;
    if (n->parent == NULL) return;
    if (n->gasMass < massInH) {
        double d0 = massInH - n->gasMass;
        double d1 = massInH - n->parent->gasMass;
        if (fabs(d0) > fabs(d1)) node_gas_density(o, n->parent, massInH);
    }
    double d0 = massInH - n->gasMass;
    double d1 = massInH - n->parent->gasMass;
    if (fabs(d0) < fabs(d1) && n->gasMass != 0) {
        for (int64_t i = 0; i < n->nlist; i++) {
            int64_t p = n->list[i];
            if (o->type[p] == 2) o->h[p] = n->radius * 2;
        }
        double rho = 0;
        for (int64_t i = 0; i < n->nlist; i++) {
            int64_t p = n->list[i];
            if (o->type[p] == 2) {
                double ddx = o->x[p] - n->comx, ddy = o->y[p] - n->comy, ddz = o->z[p] - n->comz;
                double drho = o->m[p] * spline_w(sqrt(ddx * ddx + ddy * ddy + ddz * ddz), o->h[p]);
                rho += drho;
            }
        }
        for (int64_t i = 0; i < n->nlist; i++) {
            int64_t p = n->list[i];
            if (o->type[p] == 2) {
                o->rho[p] = rho;
                o->P[p] = (AG_GAMMA - 1.0) * o->U[p] * o->rho[p];
                o->T[p] = (AG_GAMMA - 1.0) * o->U[p] * AG_PRTN * o->mu[p] / (AG_KB);
            }
        }
    }
}

/* ---- in-walk SPH (Node.cpp:88-172) ---- */
static void sph_force(Oracle* o, const Node* n, int64_t t, double out[3])
{
    double ax = 0, ay = 0, az = 0;
    double h_i = o->h[t];
    double h_j = h_i;                                   /* :94  */
    double h_ij = (h_i + h_j) / 2.0;
    double rho_i = o->rho[t], rho_j = rho_i;            /* :101 */
    double P_i = o->P[t], P_j = P_i;                    /* :108 */
    double vjx = n->mvx, vjy = n->mvy, vjz = n->mvz;
    if (n->isLeaf) { vjx = o->vx[n->particle]; vjy = o->vy[n->particle]; vjz = o->vz[n->particle]; }
    double vx = o->vx[t] - vjx, vy = o->vy[t] - vjy, vz = o->vz[t] - vjz;
    double dx = o->x[t] - n->comx, dy = o->y[t] - n->comy, dz = o->z[t] - n->comz;
    double r = sqrt(dx * dx + dy * dy + dz * dz);
    double c_i = sqrt(AG_GAMMA * P_i / rho_i);
    double c_j = sqrt(AG_GAMMA * P_i / rho_i);
    double c_ij = (c_i + c_j) / 2.0;
    double g[3];
    spline_grad(dx, dy, dz, h_i, g);
    double s = -n->gasMass * (P_i / (rho_i * rho_i) + P_j / (rho_j * rho_j));
    ax += g[0] * s; ay += g[1] * s; az += g[2] * s;
    double MU = 0.0;
    {
        double alpha = 0.5, beta = 1.0, eta = 0.01;
        double vd = vx * dx + vy * dy + vz * dz;
        double mu_ij = h_ij * vd / (r * r + eta * (h_ij * h_ij));
        if (vd < 0) MU = -alpha * c_ij * mu_ij + beta * (mu_ij * mu_ij);
        double g2[3];
        spline_grad(dx, dy, dz, h_ij, g2);
        double s2 = -n->gasMass * MU;
        ax += g2[0] * s2; ay += g2[1] * s2; az += g2[2] * s2;
    }
    o->dUdt[t] += 1.0 / 2.0 * n->gasMass * (P_i / (rho_i * rho_i) + P_j / (rho_j * rho_j) + MU) * (vx * g[0] + vy * g[1] + vz * g[2]);
    if (isnan(ax) || isnan(ay) || isnan(az)) { out[0] = out[1] = out[2] = 0; return; }
    out[0] = ax; out[1] = ay; out[2] = az;
}

/* Shared tail of the leaf / accepted-node branches (Node.cpp:285-325 and :337-377).
 * Returns 0 if the reference returned before adding anything. */
static int gravity_pair(Oracle* o, const Node* n, int64_t t, double dx, double dy, double dz, double r, double e0)
{
    double u = r / (2.8 * e0);
    double k = softening_kernel(u);
    if ((k - r) == 0) return 0;
    double e = -(2.8 * e0) / (k - r);
    if (isnan(e)) return 0;
    double len = sqrt(dx * dx + dy * dy + dz * dz);
    double nx = 0, ny = 0, nz = 0;
    if (len > 0) { nx = dx / len; ny = dy / len; nz = dz / len; }
    double f = AG_G * n->mass / (r * r + e0 * e0);
    double gx = nx * f, gy = ny * f, gz = nz * f;
    /* the reference's unqualified abs() is int abs(int) */
    int ie = (int)e;
    if (abs(ie) > (e0 * 0.0001) && abs(ie) < 1e30) {
        double f2 = AG_G * n->mass / (r * r + e * e);
        gx = nx * f2; gy = ny * f2; gz = nz * f2;
    }
    if (isnan(gx) || isnan(gy) || isnan(gz)) return 0;
    o->ax[t] += gx; o->ay[t] += gy; o->az[t] += gz;
    return 1;
}

static void walk(Oracle* o, const Node* n, int64_t t, double e0, double theta)
{                                          /* Node.cpp:247-399 */
    if (o->visits) o->visits[t]++;
    if (n->mass == 0) return;
    if (n->isLeaf && n->particle == t) return;   /* pointer identity with the leaf's particle */
    if (o->m[t] == 0) return;
    double dx = n->comx - o->x[t], dy = n->comy - o->y[t], dz = n->comz - o->z[t];
    double r = sqrt(dx * dx + dy * dy + dz * dz);
    if (r == 0) return;
    if (n->isLeaf) {
        if (n->particle >= 0 && n->particle != t) {
            if (o->acc_leaves) o->acc_leaves[t]++;
            if (!gravity_pair(o, n, t, dx, dy, dz, r, e0)) return;
            if (r < o->h[t] * 2) {
                if (o->type[n->particle] == 2 && o->type[t] == 2) {
                    double f[3];
                    if (o->sph) o->sph[t]++;
                    sph_force(o, n, t, f);
                    o->ax[t] += f[0]; o->ay[t] += f[1]; o->az[t] += f[2];
                }
            }
        }
    } else {
        double s = n->radius / r;
        if (s < theta) {
            if (o->acc_nodes) o->acc_nodes[t]++;
            if (!gravity_pair(o, n, t, dx, dy, dz, r, e0)) return;
            if (r < o->h[t] * 2) {
                if (o->type[t] == 2 && n->gasMass > 0) {
                    double f[3];
                    if (o->sph) o->sph[t]++;
                    sph_force(o, n, t, f);
                    o->ax[t] += f[0]; o->ay[t] += f[1]; o->az[t] += f[2];
                }
            }
        } else {
            for (int i = 0; i < 8; i++) {
                const Node* c = n->child[i];
                if (c == NULL) continue;
                if (c->mass == 0) continue;
                walk(o, c, t, e0, theta);
            }
        }
    }
}

/* =============================== C API (ctypes) =============================== */

typedef struct {
    int64_t n;
    const double *x, *y, *z, *vx, *vy, *vz, *mass, *U, *next_time, *mu;
    const uint8_t* type;
    double *rho, *P, *T;                   /* in/out */
    double *ax, *ay, *az, *dUdt, *h, *vis; /* out (dUdt accumulates, like the reference) */
    int32_t* leafdepth; uint64_t *key_hi, *key_lo;           /* optional (may be NULL) */
    int32_t *visits, *acc_nodes, *acc_leaves, *sph;          /* optional (may be NULL) */
} AgOracleIO;

static void leaf_path(const Node* leaf, int32_t* depth, uint64_t* hi, uint64_t* lo)
{
    int oct[128], d = 0;
    for (const Node* n = leaf; n->parent != NULL; n = n->parent) {
        int o = 0;
        for (int i = 0; i < 8; i++) if (n->parent->child[i] == n) o = i;
        if (d < 128) oct[d] = o;
        d++;
    }
    *depth = d; *hi = 0; *lo = 0;
    for (int l = 0; l < d && l < 42; l++) {
        uint64_t oc = (uint64_t)oct[d - 1 - l];
        if (l < 21) *hi |= oc << (60 - 3 * l); else *lo |= oc << (60 - 3 * (l - 21));
    }
}

void* ag_oracle_create(const AgOracleIO* io)
{
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    o->n = io->n;
    o->x = io->x; o->y = io->y; o->z = io->z; o->vx = io->vx; o->vy = io->vy; o->vz = io->vz;
    o->m = io->mass; o->U = io->U; o->next_time = io->next_time; o->mu = io->mu; o->type = io->type;
    o->rho = io->rho; o->P = io->P; o->T = io->T;
    o->ax = io->ax; o->ay = io->ay; o->az = io->az; o->dUdt = io->dUdt; o->h = io->h; o->vis = io->vis;
    o->visits = io->visits; o->acc_nodes = io->acc_nodes; o->acc_leaves = io->acc_leaves; o->sph = io->sph;
    o->leaf = (Node**)calloc((size_t)(io->n ? io->n : 1), sizeof(Node*));
    return o;
}

/* Tree::buildTree (Tree.cpp:24-55); returns root->radius. */
double ag_oracle_build(void* h, int cores)
{
    Oracle* o = (Oracle*)h;
    o->root = new_node(o);
    o->root->radius = tree_width(o);
    o->root->depth = 0;
    int64_t* all = (int64_t*)malloc((size_t)(o->n ? o->n : 1) * sizeof(int64_t));
    for (int64_t i = 0; i < o->n; i++) all[i] = i;
    insert_bulk(o, o->root, all, o->n, cores);
    free(all);
    return o->root->radius;
}

/* Tree::calcVisualDensity (Tree.cpp:152-176) */
void ag_oracle_visual_density(void* h, double radius)
{
    Oracle* o = (Oracle*)h;
    for (int64_t i = 0; i < o->n; i++) o->vis[i] = 0;
    for (int64_t i = 0; i < o->n; i++) {
        if (o->vis[i] != 0) continue;
        if (o->leaf[i]) node_visual_density(o, o->leaf[i], radius);
    }
}

/* Tree::calcGasDensity (Tree.cpp:119-150) */
void ag_oracle_gas_density(void* h, double massInH)
{
    Oracle* o = (Oracle*)h;
    for (int64_t i = 0; i < o->n; i++) if (o->type[i] == 2) o->h[i] = 0;
    for (int64_t i = 0; i < o->n; i++)
        if (o->type[i] == 2 && o->leaf[i] && o->h[i] == 0) node_gas_density(o, o->leaf[i], massInH);
}

/* Tree::calculateForces (Tree.cpp:57-83) */
void ag_oracle_forces(void* h, double globalTime, double e0, double theta)
{
    Oracle* o = (Oracle*)h;
    for (int64_t i = 0; i < o->n; i++) {
        if (o->visits) { o->visits[i] = 0; o->acc_nodes[i] = 0; o->acc_leaves[i] = 0; o->sph[i] = 0; }
        if (globalTime == o->next_time[i]) {
            o->ax[i] = 0.0; o->ay[i] = 0.0; o->az[i] = 0.0;
            walk(o, o->root, i, e0, theta);
        }
    }
}

void ag_oracle_paths(void* h, int32_t* leafdepth, uint64_t* key_hi, uint64_t* key_lo)
{
    Oracle* o = (Oracle*)h;
    for (int64_t i = 0; i < o->n; i++) {
        if (o->leaf[i]) leaf_path(o->leaf[i], &leafdepth[i], &key_hi[i], &key_lo[i]);
        else { leafdepth[i] = -1; key_hi[i] = 0; key_lo[i] = 0; }
    }
}

typedef struct {
    int32_t *depth, *isLeaf; int64_t* nchild; uint64_t *key_hi, *key_lo;
    double *mass, *comx, *comy, *comz, *gasMass, *mvx, *mvy, *mvz;
    int64_t count;
} NodeOut;

static void dump_rec(const Node* n, int depth, uint64_t hi, uint64_t lo, NodeOut* D)
{
    int empty = n->isLeaf && n->particle < 0;
    if (!empty) {
        if (D->depth) {
            int64_t k = D->count;
            D->depth[k] = depth; D->isLeaf[k] = n->isLeaf; D->nchild[k] = n->nlist; D->key_hi[k] = hi; D->key_lo[k] = lo;
            D->mass[k] = n->mass; D->comx[k] = n->comx; D->comy[k] = n->comy; D->comz[k] = n->comz;
            D->gasMass[k] = n->gasMass; D->mvx[k] = n->mvx; D->mvy[k] = n->mvy; D->mvz[k] = n->mvz;
        }
        D->count++;
    }
    for (int i = 0; i < 8; i++) {
        if (!n->child[i]) continue;
        uint64_t h2 = hi, l2 = lo;
        if (depth < 21) h2 |= (uint64_t)i << (60 - 3 * depth); else if (depth < 42) l2 |= (uint64_t)i << (60 - 3 * (depth - 21));
        dump_rec(n->child[i], depth + 1, h2, l2, D);
    }
}

/* Non-empty nodes in the reference's depth-first child order; call with NULL arrays to count. */
int64_t ag_oracle_nodes(void* h, int32_t* depth, int32_t* isLeaf, int64_t* nchild, uint64_t* key_hi, uint64_t* key_lo,
                        double* mass, double* comx, double* comy, double* comz, double* gasMass, double* mvx, double* mvy, double* mvz)
{
    Oracle* o = (Oracle*)h;
    NodeOut D = { depth, isLeaf, nchild, key_hi, key_lo, mass, comx, comy, comz, gasMass, mvx, mvy, mvz, 0 };
    if (o->root) dump_rec(o->root, 0, 0, 0, &D);
    return D.count;
}

void ag_oracle_destroy(void* h)
{
    Oracle* o = (Oracle*)h;
    for (int64_t b = 0; b < o->nblocks; b++) {
        int64_t cnt = (b == o->nblocks - 1) ? o->used_in_block : BLOCK;
        for (int64_t i = 0; i < cnt; i++) free(o->blocks[b][i].list);
        free(o->blocks[b]);
    }
    free(o->blocks);
    free(o->leaf);
    free(o);
}

// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference force path, compiled in place from /root/reference by
// oracle/Makefile into oracle/_ref/ (git-ignored), exactly the way the reference's own driver
// does (Simulation.cpp:120-139 / :276-285):
//     Tree t(&sim); t.buildTree(); sim.visualDensityRadius = t.root->radius/100000;
//     t.calcVisualDensity(); t.calcGasDensity(); t.calculateForces();
// and dumps everything the parity tiers of SURVEY.md §8(c) need: root radius, per-particle
// outputs, per-particle leaf path, per-target interaction counters (from a read-only walk that
// mirrors the control flow of Node::calculateGravityForce, Node.cpp:247-399) and the node table.
//
// Sub-commands
//   ag_ref run     <in.agp> <out.ago> <theta> <e0> <massInH> <globalTime> <cores> [nodes=1]
//   ag_ref convert <format> <path-below-input_data> <out.agp>      (uses DataManager::loadICs)
//   ag_ref time    <in.agp> <theta> <e0> <massInH> <globalTime> <reps>   (phase timings, JSON)
//   ag_ref steps   <in.agp> <out.agp> <theta> <e0> <massInH> <cores> <eta> <minTimeStep> <maxTimeStep> <H0> <nsteps>
//                  initial force evaluation + nsteps iterations of the reference's main loop (Simulation.cpp:166-345) with the
//                  reference's own TimeIntegration and Tree; writes the final particle state (the loop itself is restated here
//                  because Simulation::run is welded to Config.ini, the console and snapshot output)
//
// File formats (little endian), shared with oracle/agio.py:
//   .agp  "AGPART01", int64 N, double[N] x y z vx vy vz mass U next_time rho P T mu, uint8[N] type
//   .ago  "AGOUT001", int64 N, double R, double[N] ax ay az dUdt h rho P T vis,
//         int32[N] leafdepth (-1 = not in tree), uint64[N] key_hi key_lo (21 levels x 3 bit each,
//         level l<21 at hi>>(60-3l), else lo>>(60-3(l-21)); octant = x|y<<1|z<<2),
//         int32[N] visits acc_nodes acc_leaves sph,
//         int64 M, int32[M] depth isLeaf, int64[M] nchild, uint64[M] key_hi key_lo,
//         double[M] mass comx comy comz gasMass mvx mvy mvz
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "Simulation.h"
#include "Tree.h"
#include "Node.h"
#include "DataManager.h"
#include "TimeIntegration.h"
#include "Units.h"
#include <algorithm>
#include <cmath>
#include <limits>

extern "C" { int ag_stub_cores = 1; }

namespace {

struct Counters { int visits = 0, acc_nodes = 0, acc_leaves = 0, sph = 0; };

template <class T> void wr(FILE* f, const std::vector<T>& v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }
template <class T> void rd(FILE* f, std::vector<T>& v, size_t n) { v.resize(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

bool load_agp(const char* path, std::vector<Particle*>& ps)
{
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); return false; }
    char magic[8]; int64_t n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "AGPART01", 8) || fread(&n, 8, 1, f) != 1) { fprintf(stderr, "bad .agp\n"); return false; }
    std::vector<double> a[13];
    for (auto& v : a) rd(f, v, (size_t)n);
    std::vector<uint8_t> type; rd(f, type, (size_t)n);
    fclose(f);
    ps.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        Particle* p = new Particle();
        p->position = vec3(a[0][i], a[1][i], a[2][i]);
        p->velocity = vec3(a[3][i], a[4][i], a[5][i]);
        p->mass = a[6][i]; p->U = a[7][i]; p->nextIntegrationTime = a[8][i];
        p->rho = a[9][i]; p->P = a[10][i]; p->T = a[11][i]; p->mu = a[12][i];
        p->type = type[i]; p->id = (unsigned)i; p->node = nullptr;
        ps[(size_t)i] = p;
    }
    return true;
}

bool save_agp(const char* path, const std::vector<Particle*>& ps)
{
    FILE* f = fopen(path, "wb");
    if (!f) { perror(path); return false; }
    int64_t n = (int64_t)ps.size();
    fwrite("AGPART01", 1, 8, f); fwrite(&n, 8, 1, f);
    std::vector<double> v((size_t)n);
    auto col = [&](auto get) { for (int64_t i = 0; i < n; i++) v[(size_t)i] = get(ps[(size_t)i]); wr(f, v); };
    col([](Particle* p) { return p->position.x; }); col([](Particle* p) { return p->position.y; }); col([](Particle* p) { return p->position.z; });
    col([](Particle* p) { return p->velocity.x; }); col([](Particle* p) { return p->velocity.y; }); col([](Particle* p) { return p->velocity.z; });
    col([](Particle* p) { return p->mass; }); col([](Particle* p) { return p->U; }); col([](Particle* p) { return p->nextIntegrationTime; });
    col([](Particle* p) { return p->rho; }); col([](Particle* p) { return p->P; }); col([](Particle* p) { return p->T; }); col([](Particle* p) { return p->mu; });
    std::vector<uint8_t> t((size_t)n);
    for (int64_t i = 0; i < n; i++) t[(size_t)i] = ps[(size_t)i]->type;
    wr(f, t);
    fclose(f);
    return true;
}

// Read-only mirror of the decision structure of Node::calculateGravityForce (Node.cpp:247-399).
void count_walk(const Node* n, const Particle* p, double theta, Counters& c)
{
    c.visits++;
    if (n->mass == 0 || p == n->particle || p->mass == 0) return;
    vec3 d = n->centerOfMass - p->position;
    double r = d.length();
    if (r == 0) return;
    if (n->isLeaf) {
        if (n->particle != nullptr && p != n->particle) {
            c.acc_leaves++;
            if (r < p->h * 2 && n->particle->type == 2 && p->type == 2) c.sph++;
        }
        return;
    }
    if (n->radius / r < theta) {
        c.acc_nodes++;
        if (r < p->h * 2 && p->type == 2 && n->gasMass > 0) c.sph++;
        return;
    }
    for (int i = 0; i < 8; i++) {
        const Node* ch = n->children[i];
        if (ch == nullptr || ch->mass == 0) continue;
        count_walk(ch, p, theta, c);
    }
}

void path_of(const Node* leaf, int& depth, uint64_t& hi, uint64_t& lo)
{
    std::vector<int> oct;
    for (const Node* n = leaf; n->parent != nullptr; n = n->parent) {
        int o = -1;
        for (int i = 0; i < 8; i++) if (n->parent->children[i] == n) o = i;
        oct.push_back(o);
    }
    depth = (int)oct.size(); hi = lo = 0;
    for (int l = 0; l < depth && l < 42; l++) {
        uint64_t o = (uint64_t)oct[(size_t)(depth - 1 - l)];
        if (l < 21) hi |= o << (60 - 3 * l); else lo |= o << (60 - 3 * (l - 21));
    }
}

struct NodeDump {
    std::vector<int32_t> depth, isLeaf; std::vector<int64_t> nchild; std::vector<uint64_t> hi, lo;
    std::vector<double> mass, cx, cy, cz, gas, vx, vy, vz;
};

void dump_nodes(const Node* n, int depth, uint64_t hi, uint64_t lo, NodeDump& D)
{
    bool empty = n->isLeaf && n->particle == nullptr;
    if (!empty) {
        D.depth.push_back(depth); D.isLeaf.push_back(n->isLeaf ? 1 : 0); D.nchild.push_back((int64_t)n->childParticles.size());
        D.hi.push_back(hi); D.lo.push_back(lo);
        D.mass.push_back(n->mass); D.cx.push_back(n->centerOfMass.x); D.cy.push_back(n->centerOfMass.y); D.cz.push_back(n->centerOfMass.z);
        D.gas.push_back(n->gasMass); D.vx.push_back(n->mVel.x); D.vy.push_back(n->mVel.y); D.vz.push_back(n->mVel.z);
    }
    for (int i = 0; i < 8; i++) {
        if (!n->children[i]) continue;
        uint64_t h2 = hi, l2 = lo;
        if (depth < 21) h2 |= (uint64_t)i << (60 - 3 * depth); else if (depth < 42) l2 |= (uint64_t)i << (60 - 3 * (depth - 21));
        dump_nodes(n->children[i], depth + 1, h2, l2, D);
    }
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int cmd_run(int argc, char** argv)
{
    if (argc < 9) { fprintf(stderr, "usage: run in out theta e0 massInH globalTime cores [nodes]\n"); return 2; }
    Simulation sim;
    if (!load_agp(argv[2], sim.particles)) return 2;
    sim.numberOfParticles = (int)sim.particles.size();
    sim.theta = atof(argv[4]); sim.e0 = atof(argv[5]); sim.massInH = atof(argv[6]); sim.globalTime = atof(argv[7]);
    ag_stub_cores = atoi(argv[8]);
    bool want_nodes = argc > 9 ? atoi(argv[9]) != 0 : true;
    const int64_t n = sim.numberOfParticles;

    Tree t(&sim);
    t.buildTree();
    sim.visualDensityRadius = t.root->radius / 100000;
    t.calcVisualDensity();
    t.calcGasDensity();
    t.calculateForces();

    FILE* f = fopen(argv[3], "wb");
    if (!f) { perror(argv[3]); return 2; }
    double R = t.root->radius;
    fwrite("AGOUT001", 1, 8, f); fwrite(&n, 8, 1, f); fwrite(&R, 8, 1, f);
    std::vector<double> v((size_t)n);
    auto col = [&](auto get) { for (int64_t i = 0; i < n; i++) v[(size_t)i] = get(sim.particles[(size_t)i]); wr(f, v); };
    col([](Particle* p) { return p->acc.x; }); col([](Particle* p) { return p->acc.y; }); col([](Particle* p) { return p->acc.z; });
    col([](Particle* p) { return p->dUdt; }); col([](Particle* p) { return p->h; }); col([](Particle* p) { return p->rho; });
    col([](Particle* p) { return p->P; }); col([](Particle* p) { return p->T; }); col([](Particle* p) { return p->visualDensity; });
    std::vector<int32_t> ld((size_t)n); std::vector<uint64_t> khi((size_t)n), klo((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        Particle* p = sim.particles[(size_t)i];
        if (p->node) { int d; path_of(p->node, d, khi[(size_t)i], klo[(size_t)i]); ld[(size_t)i] = d; }
        else { ld[(size_t)i] = -1; khi[(size_t)i] = klo[(size_t)i] = 0; }
    }
    wr(f, ld); wr(f, khi); wr(f, klo);
    std::vector<int32_t> c0((size_t)n), c1((size_t)n), c2((size_t)n), c3((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        Particle* p = sim.particles[(size_t)i];
        Counters c;
#ifndef AG_GPU_TREE   // with the GPU Tree there is no host-side Node tree to walk
        if (sim.globalTime == p->nextIntegrationTime) count_walk(t.root, p, sim.theta, c);
#endif
        c0[(size_t)i] = c.visits; c1[(size_t)i] = c.acc_nodes; c2[(size_t)i] = c.acc_leaves; c3[(size_t)i] = c.sph;
    }
    wr(f, c0); wr(f, c1); wr(f, c2); wr(f, c3);
    NodeDump D;
#ifndef AG_GPU_TREE
    if (want_nodes) dump_nodes(t.root, 0, 0, 0, D);
#endif
    int64_t m = (int64_t)D.depth.size();
    fwrite(&m, 8, 1, f);
    wr(f, D.depth); wr(f, D.isLeaf); wr(f, D.nchild); wr(f, D.hi); wr(f, D.lo);
    wr(f, D.mass); wr(f, D.cx); wr(f, D.cy); wr(f, D.cz); wr(f, D.gas); wr(f, D.vx); wr(f, D.vy); wr(f, D.vz);
    fclose(f);
    return 0;
}

int cmd_convert(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: convert format relpath out.agp\n"); return 2; }
    // DataManager::loadICs opens "../../input_data/" + inputPath (DataManager.cpp:437).
    const char* root = getenv("AG_REFERENCE_ROOT");
    std::string cwd = std::string(root ? root : "/root/reference") + "/simulation/src";
    if (chdir(cwd.c_str()) != 0) { perror(cwd.c_str()); return 2; }
    Simulation sim;
    DataManager dm("");
    dm.inputPath = argv[3]; dm.inputFormat = argv[2];
    // The makeGal branch loads correctly but falls through to `return false` (DataManager.cpp:580-810,1330);
    // the reference driver ignores the return value (Simulation.cpp:56), so only an empty set is an error.
    dm.loadICs(sim.particles, &sim);
    if (sim.particles.empty()) return 2;
    return save_agp(argv[4], sim.particles) ? 0 : 2;
}

// DataManager::saveData on a set loaded by DataManager::loadICs: <outdir>/<timeStep>.<fmt> (DataManager.cpp:86-424).
int cmd_save(int argc, char** argv)
{
    if (argc < 10) { fprintf(stderr, "usage: save in_format relpath out_format outdir/ timeStep deltaTime endTime currentTime [count]\n"); return 2; }
    const char* root = getenv("AG_REFERENCE_ROOT");
    std::string cwd = std::string(root ? root : "/root/reference") + "/simulation/src";
    if (chdir(cwd.c_str()) != 0) { perror(cwd.c_str()); return 2; }
    Simulation sim;
    DataManager dm(argv[5]);
    dm.inputPath = argv[3]; dm.inputFormat = argv[2]; dm.outputFormat = argv[4];
    dm.loadICs(sim.particles, &sim);
    if (sim.particles.empty()) return 2;
    const int count = argc > 10 ? atoi(argv[10]) : (int)sim.particles.size();
    dm.saveData(sim.particles, atoi(argv[6]), 0, count, atof(argv[7]), atof(argv[8]), atof(argv[9]));
    return 0;
}

// Phase timings of the reference path (same phases as its processLog.csv, Simulation.cpp:120-139).
int cmd_time(int argc, char** argv)
{
    if (argc < 8) { fprintf(stderr, "usage: time in theta e0 massInH globalTime reps\n"); return 2; }
    Simulation sim;
    if (!load_agp(argv[2], sim.particles)) return 2;
    sim.numberOfParticles = (int)sim.particles.size();
    sim.theta = atof(argv[3]); sim.e0 = atof(argv[4]); sim.massInH = atof(argv[5]); sim.globalTime = atof(argv[6]);
    int reps = atoi(argv[7]);
#if defined(_OPENMP)
    ag_stub_cores = 0;
#else
    if (getenv("AG_CORES")) ag_stub_cores = atoi(getenv("AG_CORES"));
#endif
    for (int r = 0; r < reps; r++) {
        double t0 = now();
        Tree* t = new Tree(&sim);
        t->buildTree();
        double t1 = now();
        sim.visualDensityRadius = t->root->radius / 100000;
        t->calcVisualDensity();
        double t2 = now();
        t->calcGasDensity();
        double t3 = now();
        t->calculateForces();
        double t4 = now();
        delete t;
        double t5 = now();
        printf("{\"rep\": %d, \"n\": %d, \"build\": %.6f, \"visual\": %.6f, \"gas_density\": %.6f, \"forces\": %.6f, \"delete\": %.6f}\n",
               r, sim.numberOfParticles, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4);
        fflush(stdout);
    }
    return 0;
}


// Restatement of the control flow of Simulation::init (force part) + Simulation::run for `nsteps` loop iterations,
// calling the reference's own Tree and TimeIntegration.  Output .agp carries acc in (rho,P,T unchanged) plus a second
// file <out>.acc with ax ay az dUdt h visualDensity next_time timeStep globalTime.
int cmd_steps(int argc, char** argv)
{
    if (argc < 13) { fprintf(stderr, "usage: steps in out theta e0 massInH cores eta minTS maxTS H0 nsteps\n"); return 2; }
    Simulation sim;
    if (!load_agp(argv[2], sim.particles)) return 2;
    const int n = sim.numberOfParticles = (int)sim.particles.size();
    sim.theta = atof(argv[4]); sim.e0 = atof(argv[5]); sim.massInH = atof(argv[6]);
    ag_stub_cores = atoi(argv[7]);
    sim.eta = atof(argv[8]); sim.minTimeStep = atof(argv[9]); sim.maxTimeStep = atof(argv[10]); sim.H0 = atof(argv[11]);
    const int nsteps = atoi(argv[12]);
    TimeIntegration ti;
    std::vector<Particle*>& ps = sim.particles;
    sim.globalTime = 0.0;
    for (int i = 0; i < n; i++)                                       // Simulation.cpp:101-113
        if (ps[i]->type == 2) ps[i]->T = (Constants::GAMMA - 1.0) * ps[i]->U * Constants::prtn * ps[i]->mu / (Constants::k_b);
    {   // Simulation.cpp:120-139
        Tree t(&sim); t.buildTree(); sim.visualDensityRadius = t.root->radius / 100000;
        t.calcVisualDensity(); t.calcGasDensity(); t.calculateForces();
    }
    auto assign = [&](Particle* p) {                                  // Simulation.cpp:196-207 / :222-232
        double accelMag = p->acc.length();
        if (accelMag > 0) {
            double timeStep = sim.eta * std::sqrt(sim.e0 / accelMag);
            p->timeStep = std::clamp(timeStep, sim.minTimeStep, sim.maxTimeStep);
            p->timeStep = std::max(std::pow(2, std::floor(std::log2(p->timeStep))), sim.minTimeStep);
            p->nextIntegrationTime = sim.globalTime + p->timeStep;
        } else { p->timeStep = sim.minTimeStep; p->nextIntegrationTime = sim.globalTime + p->timeStep; }
    };
    for (int i = 0; i < n; i++) ps[i]->nextIntegrationTime = 0.0;
    for (int i = 0; i < n; i++) assign(ps[i]);
    for (int step = 0; step < nsteps; step++) {
        for (int i = 0; i < n; i++) if (sim.globalTime >= ps[i]->nextIntegrationTime) assign(ps[i]);
        double mn = std::numeric_limits<double>::max();
        for (int i = 0; i < n; i++) if (ps[i]->nextIntegrationTime < mn) mn = ps[i]->nextIntegrationTime;
        sim.globalTime = mn;
        for (int i = 0; i < n; i++) if (sim.globalTime == ps[i]->nextIntegrationTime) { ti.Kick(ps[i], ps[i]->timeStep); ti.Drift(ps[i], ps[i]->timeStep); }
        Tree* t = new Tree(&sim);
        t->buildTree(); t->calcVisualDensity(); t->calcGasDensity(); t->calculateForces();
        for (int i = 0; i < n; i++) {
            Particle* p = ps[i];
            if (sim.globalTime == p->nextIntegrationTime) {
                if (p->type == 2) ti.Ueuler(p, p->timeStep);
                double H0SI = (sim.H0 * Units::KMS) / Units::MPC;
                double scale_factor = exp(H0SI * p->timeStep);
                p->position *= scale_factor;
                ti.Kick(p, p->timeStep);
                p->nextIntegrationTime += p->timeStep;
            }
        }
        delete t;
    }
    if (!save_agp(argv[3], ps)) return 2;
    std::string extra = std::string(argv[3]) + ".acc";
    FILE* f = fopen(extra.c_str(), "wb");
    if (!f) return 2;
    std::vector<double> v((size_t)n);
    auto col = [&](auto get) { for (int i = 0; i < n; i++) v[(size_t)i] = get(ps[(size_t)i]); wr(f, v); };
    col([](Particle* p) { return p->acc.x; }); col([](Particle* p) { return p->acc.y; }); col([](Particle* p) { return p->acc.z; });
    col([](Particle* p) { return p->dUdt; }); col([](Particle* p) { return p->h; }); col([](Particle* p) { return p->visualDensity; });
    col([](Particle* p) { return p->nextIntegrationTime; }); col([](Particle* p) { return p->timeStep; });
    fwrite(&sim.globalTime, 8, 1, f);
    fclose(f);
    return 0;
}

} // namespace

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: ag_ref run|convert|time ...\n"); return 2; }
    std::string cmd = argv[1];
    if (cmd == "run") return cmd_run(argc, argv);
    if (cmd == "convert") return cmd_convert(argc, argv);
    if (cmd == "save") return cmd_save(argc, argv);
    if (cmd == "time") return cmd_time(argc, argv);
    if (cmd == "steps") return cmd_steps(argc, argv);
    fprintf(stderr, "unknown command %s\n", argv[1]);
    return 2;
}

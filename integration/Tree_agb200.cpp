// integration/Tree_agb200.cpp — drop-in replacement for the reference's simulation/src/Physics/Tree/Tree.cpp.
//
// It implements the member functions of the reference's OWN `class Tree` (declared in its unmodified
// Physics/Tree/Tree.h:13-27) on top of the C ABI in include/agb200.h, so the reference's driver
// (Physics/Simulation.cpp:121-139, :276-285, :345) runs unchanged:
//
//     Tree* tree = new Tree(this); tree->buildTree(); ... tree->root->radius ...
//     tree->calcVisualDensity(); tree->calcGasDensity(); tree->calculateForces(); delete tree;
//
// Build: compile the reference's sources with Tree.cpp (and optionally Node.cpp's walk, which is then unused)
// replaced by this file, add -I<repo>/include and link -lagb200.  `oracle/Makefile` target `_ref/ag_ref_gpu`
// does exactly that with the test harness as the driver; tests/test_gpu_integration.py compares its output with
// the CPU reference.  No reference source is copied: only its public headers are included at compile time.
//
// State kept across steps: one handle per process (pooled device memory), created on first use with
// compat_cores = omp_get_max_threads() — the value the reference passes to Node::insert (Tree.cpp:44-48).
// Devices: AGB_DEVICES="0,1,2,3" (several GPUs of the box: every one builds the tree, each walks a share of the
// targets — the reference's `#pragma omp parallel for` over targets, Tree.cpp:65) or AGB_DEVICE=<n> (default 0).
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <omp.h>
#include <string>
#include <vector>

#include "Tree.h"
#include "agb200.h"

namespace {

agb_multi* g_ctx = nullptr;

agb_multi* context()
{
    if (!g_ctx) {
        std::vector<int> devs;
        if (const char* list = getenv("AGB_DEVICES")) {
            std::string s(list);
            for (size_t a = 0; a < s.size();) { size_t b = s.find(',', a); if (b == std::string::npos) b = s.size(); if (b > a) devs.push_back(atoi(s.substr(a, b - a).c_str())); a = b + 1; }
        }
        if (devs.empty()) { const char* dev = getenv("AGB_DEVICE"); devs.push_back(dev ? atoi(dev) : 0); }
        int rc = agb_multi_create(&g_ctx, devs.data(), (int)devs.size(), omp_get_max_threads());
        if (rc != AGB_OK) {
            std::fprintf(stderr, "agb200: %s\n", agb_strerror(rc));   // no CPU fallback: the run cannot continue
            std::abort();
        }
    }
    return g_ctx;
}

// The reference's Tree methods return void and its driver has no error path (it prints and continues, Tree.cpp:70-74).
// Continuing after a failed GPU call would integrate stale accelerations, so a non-OK status ends the run here.
void check(int rc, const char* what)
{
    if (rc == AGB_OK) return;
    std::fprintf(stderr, "agb200: %s failed: %s (%s)\n", what, agb_strerror(rc), agb_multi_last_error(g_ctx));
    std::abort();
}

const agb_aos_layout& layout()
{
    static agb_aos_layout L;
    static bool init = false;
    if (!init) {
        Particle p;
        const char* b = reinterpret_cast<const char*>(&p);
        auto off = [&](const void* f) { return (int64_t)(reinterpret_cast<const char*>(f) - b); };
        L.position = off(&p.position); L.velocity = off(&p.velocity); L.acc = off(&p.acc); L.mass = off(&p.mass);
        L.type = off(&p.type); L.U = off(&p.U); L.next_time = off(&p.nextIntegrationTime); L.mu = off(&p.mu);
        L.rho = off(&p.rho); L.P = off(&p.P); L.T = off(&p.T); L.h = off(&p.h); L.dUdt = off(&p.dUdt);
        L.visualDensity = off(&p.visualDensity);
        init = true;
    }
    return L;
}

} // namespace

Tree::~Tree()
{
    delete root;            // a bare Node that only carries `radius` for the driver (Simulation.cpp:123,126)
    root = nullptr;
}

void Tree::buildTree()
{
    agb_multi* c = context();
    root = new Node();
    root->position = vec3(0.0, 0.0, 0.0);
    root->depth = 0;
    check(agb_multi_set_particles_aos(c, reinterpret_cast<void* const*>(simulation->particles.data()), simulation->numberOfParticles, &layout()), "set_particles");
    double R = 0.0;
    check(agb_multi_build_tree(c, &R), "build_tree");
    root->radius = R;
}

double Tree::calcTreeWidth() { return root ? root->radius : 0.0; }

void Tree::calcVisualDensity() { check(agb_multi_visual_density(context(), simulation->visualDensityRadius), "visual_density"); }

void Tree::calcGasDensity() { check(agb_multi_gas_density(context(), simulation->massInH), "gas_density"); }

void Tree::calculateForces()
{
    agb_multi* c = context();
    check(agb_multi_forces(c, simulation->globalTime, simulation->e0, simulation->theta), "forces");
    // the reference writes acc, dUdt, h, rho, P, T, visualDensity straight into Particle; copy them back
    check(agb_multi_get_results_aos(c, reinterpret_cast<void* const*>(simulation->particles.data()), simulation->numberOfParticles, &layout()), "get_results");
}

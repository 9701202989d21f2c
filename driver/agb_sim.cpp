// driver/agb_sim.cpp — small C++ driver around the B200 force path (host side of SURVEY.md §8(f)-1).
//
// It reproduces the call order and the integrator of the reference's driver so that a user of
// AstroGenesis2.0 can run the same Config.ini on the GPU path:
//   * Config.ini: flat `key = value`, '#' comments, unknown key => hard error, same keys as
//     DataManager::loadConfig (simulation/src/File/DataManager.cpp:1360-1415);
//   * initial conditions: the reference's `.age` snapshots (DataManager.cpp:216-258 writer, :534-574 reader:
//     40-byte header + 94 bytes per particle), Gadget-2 SnapFormat-1 files as DataManager.cpp:811-1331 reads them
//     (the reference's shipped examples), or this repo's `.agp` test format;
//   * Simulation::init force evaluation (Physics/Simulation.cpp:101-139) and the KDK main loop of
//     Simulation::run (:166-345) with Kick / Drift / Ueuler (Physics/TimeIntegration.cpp:10-41), the Hubble
//     rescale (:330-332) and power-of-two individual time steps (:196-207, :222-232);
//   * per-phase timings in the reference's processLog.csv format (File/Log.cpp:175-221: `name;seconds`, decimal comma);
//   * snapshots in `outputDataFormat` (age, ag, agc, gadget) every endTime/fixedTimeSteps like DataManager::saveData
//     (first numParticlesOutput particles; DataManager.cpp:86-424), and `.ag`/`.agc` as input formats too (:446-533).
// Cooling / star formation are no-ops in the reference (calls commented out, Simulation.cpp:312-320) and here.
// The force path itself is only reached through the C ABI (include/agb200.h); there is no CPU fallback.
//
//   agb_sim --config Config.ini [--input-root DIR] [--output-root DIR] [--steps K] [--device D] [--gpus N] [--cores C]
//           [--precision fp64|mixed] [--dump final.agp] [--overwrite] [--device-resident] [--convert-only out.agp] [--snapshot-only DIR [--snapshot-index N] [--snapshot-time T]]
// --device-resident keeps positions, velocities and results in HBM between steps (agb_integrator_* / agb_step_*):
// only the scalar time crosses PCIe per step; state is copied back for snapshots and at the end.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "agb200.h"

namespace {

constexpr double GAMMA = 5.0 / 3.0, K_B = 1.38064852e-23, PRTN = 1.6726219e-27;   // Math/Constants.h:15-18
constexpr double KMS = 1.0e3, MPC = 3.08567758149137e22;                          // Math/Units.h

struct Config {
    double numberOfParticles = 0, eta = 2, maxTimeStep = 1e13, minTimeStep = 1e13, globalTime = 0, endTime = 1e16, fixedTimeSteps = 1000;
    double e0 = 1e19, massInH = 1e40, H0 = 70, theta = 0.5, numParticlesOutput = 0;
    bool starFormation = false, cooling = false;
    std::string inputPath, inputDataFormat = "age", outputFolderName = "run", outputDataFormat = "age";
};

std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

bool load_config(const std::string& path, Config& c)
{
    std::ifstream f(path);
    if (!f) { fprintf(stderr, "cannot open config %s\n", path.c_str()); return false; }
    std::string line;
    while (std::getline(f, line)) {
        std::string t = trim(line);
        if (t.empty() || t[0] == '#') continue;
        size_t eq = t.find('=');
        if (eq == std::string::npos) continue;                       // e.g. the [Simulation] section header
        std::string key = trim(t.substr(0, eq)), val = trim(t.substr(eq + 1));
        auto num = [&](double& d) { d = std::stod(val); };
        auto flag = [&](bool& b) { if (val == "true" || val == "True") b = true; else if (val == "false" || val == "False") b = false; };
        try {
            if (key == "numberOfParticles") num(c.numberOfParticles);
            else if (key == "eta") num(c.eta);
            else if (key == "maxTimeStep") num(c.maxTimeStep);
            else if (key == "minTimeStep") num(c.minTimeStep);
            else if (key == "globalTime") num(c.globalTime);
            else if (key == "endTime") num(c.endTime);
            else if (key == "fixedTimeSteps") num(c.fixedTimeSteps);
            else if (key == "e0") num(c.e0);
            else if (key == "massInH") num(c.massInH);
            else if (key == "starformation") flag(c.starFormation);
            else if (key == "cooling") flag(c.cooling);
            else if (key == "H0") num(c.H0);
            else if (key == "theta") num(c.theta);
            else if (key == "inputPath") c.inputPath = val;
            else if (key == "inputDataFormat") c.inputDataFormat = val;
            else if (key == "outputFolderName") c.outputFolderName = val;
            else if (key == "outputDataFormat") c.outputDataFormat = val;
            else if (key == "numParticlesOutput") num(c.numParticlesOutput);
            else { fprintf(stderr, "unknown key: %s\n", key.c_str()); return false; }
        } catch (const std::exception&) { fprintf(stderr, "invalid value for %s: %s\n", key.c_str(), val.c_str()); return false; }
    }
    if (c.numParticlesOutput > c.numberOfParticles) { fprintf(stderr, "Number of particles to output is greater than the total number of particles.\n"); return false; }
    return true;
}

struct Particles {
    int64_t n = 0;
    std::vector<double> x, y, z, vx, vy, vz, mass, U, next, mu, rho, P, T, h, dUdt, ax, ay, az, vis, timeStep, sfr;
    std::vector<uint8_t> type, galaxyPart;
    std::vector<uint32_t> id;
    void resize(int64_t m)
    {
        n = m;
        for (auto* v : {&x, &y, &z, &vx, &vy, &vz, &mass, &U, &next, &rho, &P, &T, &h, &dUdt, &ax, &ay, &az, &vis, &timeStep, &sfr}) v->assign((size_t)m, 0.0);
        mu.assign((size_t)m, 0.58);
        type.assign((size_t)m, 1); galaxyPart.assign((size_t)m, 1); id.assign((size_t)m, 0);
    }
};

#pragma pack(push, 1)
struct AgeRecord { double pos[3], vel[3], mass, T, P, visualDensity, U; uint8_t type, galaxyPart; uint32_t id; };
#pragma pack(pop)
struct AgeHeader { int32_t numParticles[3]; int32_t pad; double deltaTime, endTime, currentTime; };
static_assert(sizeof(AgeRecord) == 94 && sizeof(AgeHeader) == 40, ".age layout");

bool load_age(const std::string& path, Particles& p)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { perror(path.c_str()); return false; }
    AgeHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1) { fclose(f); return false; }
    const int64_t n = (int64_t)h.numParticles[0] + h.numParticles[1] + h.numParticles[2];
    p.resize(n);
    std::vector<AgeRecord> buf((size_t)n);
    if (n && fread(buf.data(), sizeof(AgeRecord), (size_t)n, f) != (size_t)n) { fclose(f); fprintf(stderr, "short .age file\n"); return false; }
    fclose(f);
    for (int64_t i = 0; i < n; i++) {
        const AgeRecord& r = buf[(size_t)i];
        p.x[i] = r.pos[0]; p.y[i] = r.pos[1]; p.z[i] = r.pos[2]; p.vx[i] = r.vel[0]; p.vy[i] = r.vel[1]; p.vz[i] = r.vel[2];
        p.mass[i] = r.mass; p.T[i] = r.T; p.P[i] = r.P; p.vis[i] = r.visualDensity; p.U[i] = r.U; p.type[i] = r.type; p.galaxyPart[i] = r.galaxyPart; p.id[i] = r.id;
    }
    return true;
}

bool save_age(const std::string& path, const Particles& p, int64_t count, double deltaTime, double endTime, double currentTime)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); return false; }
    AgeHeader h{};
    for (int64_t i = 0; i < count; i++) if (p.type[i] >= 1 && p.type[i] <= 3) h.numParticles[p.type[i] - 1]++;
    h.deltaTime = deltaTime; h.endTime = endTime; h.currentTime = currentTime;
    fwrite(&h, sizeof(h), 1, f);
    std::vector<AgeRecord> buf((size_t)count);
    for (int64_t i = 0; i < count; i++) {
        AgeRecord& r = buf[(size_t)i];
        r.pos[0] = p.x[i]; r.pos[1] = p.y[i]; r.pos[2] = p.z[i]; r.vel[0] = p.vx[i]; r.vel[1] = p.vy[i]; r.vel[2] = p.vz[i];
        r.mass = p.mass[i]; r.T = p.T[i]; r.P = p.P[i]; r.visualDensity = p.vis[i]; r.U = p.U[i]; r.type = p.type[i]; r.galaxyPart = p.galaxyPart[i]; r.id = p.id[i];
    }
    fwrite(buf.data(), sizeof(AgeRecord), (size_t)count, f);
    fclose(f);
    return true;
}

// `.ag` (62 B/particle: pos, mass, T, visualDensity, sfr, type, galaxyPart, id) and `.agc` (26 B/particle: float pos,
// visualDensity, sfr, T, type, galaxyPart) — the reference's two render formats (DataManager.cpp:120-215 writers,
// :446-533 readers).  Neither carries velocities or U; `sfr` is never set by the reference (SFR.cpp is dead code), so
// it is written as 0 and ignored on input.
#pragma pack(push, 1)
struct AgRecord { double pos[3], mass, T, visualDensity, sfr; uint8_t type, galaxyPart; uint32_t id; };
struct AgcRecord { float pos[3], visualDensity, sfr, T; uint8_t type, galaxyPart; };
#pragma pack(pop)
static_assert(sizeof(AgRecord) == 62 && sizeof(AgcRecord) == 26, ".ag/.agc layout");

template <class Rec, class Fill> bool load_records(const std::string& path, Particles& p, Fill fill)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { perror(path.c_str()); return false; }
    AgeHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1) { fclose(f); return false; }
    const int64_t n = (int64_t)h.numParticles[0] + h.numParticles[1] + h.numParticles[2];
    p.resize(n);
    std::vector<Rec> buf((size_t)n);
    if (n && fread(buf.data(), sizeof(Rec), (size_t)n, f) != (size_t)n) { fclose(f); fprintf(stderr, "short snapshot file %s\n", path.c_str()); return false; }
    fclose(f);
    for (int64_t i = 0; i < n; i++) fill(buf[(size_t)i], i);
    return true;
}

bool load_ag(const std::string& path, Particles& p)
{
    return load_records<AgRecord>(path, p, [&](const AgRecord& r, int64_t i) {
        p.x[i] = r.pos[0]; p.y[i] = r.pos[1]; p.z[i] = r.pos[2]; p.mass[i] = r.mass; p.T[i] = r.T; p.vis[i] = r.visualDensity;
        p.type[i] = r.type; p.galaxyPart[i] = r.galaxyPart; p.id[i] = r.id;
    });
}

bool load_agc(const std::string& path, Particles& p)
{
    return load_records<AgcRecord>(path, p, [&](const AgcRecord& r, int64_t i) {     // mass stays 0, as in the reference (:504-528)
        p.x[i] = r.pos[0]; p.y[i] = r.pos[1]; p.z[i] = r.pos[2]; p.T[i] = r.T; p.vis[i] = r.visualDensity; p.type[i] = r.type; p.galaxyPart[i] = r.galaxyPart;
    });
}

template <class Rec, class Fill> bool save_records(const std::string& path, const Particles& p, int64_t count, double deltaTime, double endTime, double currentTime, Fill fill)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); return false; }
    AgeHeader h{};
    for (int64_t i = 0; i < count; i++) if (p.type[i] >= 1 && p.type[i] <= 3) h.numParticles[p.type[i] - 1]++;
    h.deltaTime = deltaTime; h.endTime = endTime; h.currentTime = currentTime;
    fwrite(&h, sizeof(h), 1, f);
    std::vector<Rec> buf((size_t)count);
    for (int64_t i = 0; i < count; i++) fill(buf[(size_t)i], i);
    fwrite(buf.data(), sizeof(Rec), (size_t)count, f);
    fclose(f);
    return true;
}

bool save_ag(const std::string& path, const Particles& p, int64_t count, double deltaTime, double endTime, double currentTime)
{
    return save_records<AgRecord>(path, p, count, deltaTime, endTime, currentTime, [&](AgRecord& r, int64_t i) {
        r.pos[0] = p.x[i]; r.pos[1] = p.y[i]; r.pos[2] = p.z[i]; r.mass = p.mass[i]; r.T = p.T[i]; r.visualDensity = p.vis[i]; r.sfr = p.sfr[i];
        r.type = p.type[i]; r.galaxyPart = p.galaxyPart[i]; r.id = p.id[i];
    });
}

bool save_agc(const std::string& path, const Particles& p, int64_t count, double deltaTime, double endTime, double currentTime)
{
    return save_records<AgcRecord>(path, p, count, deltaTime, endTime, currentTime, [&](AgcRecord& r, int64_t i) {
        r.pos[0] = (float)p.x[i]; r.pos[1] = (float)p.y[i]; r.pos[2] = (float)p.z[i]; r.visualDensity = (float)p.vis[i]; r.sfr = (float)p.sfr[i]; r.T = (float)p.T[i];
        r.type = p.type[i]; r.galaxyPart = p.galaxyPart[i];
    });
}

// Gadget-2 SnapFormat 1 as the reference reads it (DataManager.cpp:811-1331): 4-byte record markers around a 256-byte
// header, POS, VEL, ID, [MASS if a populated type has massarr == 0], [U if there is gas]; floats; kpc, km/s, 1e10 Msun,
// (km/s)^2.  Type map: 0 -> gas (2), 1 -> dark matter (3, halo), 2/4/5 -> star (1), 3 -> star (1, bulge).  Like the
// reference, an existing MASS block is indexed by the running particle number.
struct GadgetHeader {
    uint32_t npart[6]; double massarr[6]; double time, redshift; int32_t flag_sfr, flag_feedback; uint32_t npartTotal[6];
    int32_t flag_cooling, num_files; double BoxSize, Omega0, OmegaLambda, HubbleParam; int32_t flag_stellarage, flag_metals;
    uint32_t npartTotalHighWord[6]; int32_t flag_entropy_instead_u; char fill[60];
};
static_assert(sizeof(GadgetHeader) == 256, "gadget header");

template <class T> bool read_block(FILE* f, std::vector<T>& v, size_t want_items)
{
    uint32_t a = 0, b = 0;
    if (fread(&a, 4, 1, f) != 1) return false;
    v.assign(std::max(want_items, (size_t)a / sizeof(T)), T());
    if (a && fread(v.data(), 1, a, f) != a) return false;
    if (fread(&b, 4, 1, f) != 1 || a != b) return false;
    return true;
}

bool load_gadget(const std::string& path, Particles& p)
{
    constexpr double KPC = 3.08567758149137e19, MSUN = 1.98847e30;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { perror(path.c_str()); return false; }
    uint32_t a = 0, b = 0;
    GadgetHeader h;
    if (fread(&a, 4, 1, f) != 1 || fread(&h, sizeof(h), 1, f) != 1 || fread(&b, 4, 1, f) != 1 || a != sizeof(h)) { fclose(f); return false; }
    size_t total = 0;
    for (int i = 0; i < 6; i++) total += h.npart[i];
    std::vector<float> pos, vel, mass, u; std::vector<uint32_t> ids;
    bool ok = read_block(f, pos, total * 3) && read_block(f, vel, total * 3) && read_block(f, ids, total);
    bool individual = false;
    for (int i = 0; i < 6 && !individual; i++) if (h.massarr[i] < 1e-10f && h.npart[i] != 0) individual = true;
    if (ok && individual) ok = read_block(f, mass, total);
    if (ok && h.npart[0] > 0) ok = read_block(f, u, h.npart[0]);
    fclose(f);
    if (!ok) { fprintf(stderr, "malformed gadget file %s\n", path.c_str()); return false; }
    p.resize((int64_t)total);
    size_t cur = 0, gas = 0;
    for (int type = 0; type < 6; type++)
        for (uint32_t k = 0; k < h.npart[type]; k++, cur++) {
            p.id[cur] = ids[cur];
            if (type == 1) { p.type[cur] = 3; p.galaxyPart[cur] = 3; }
            else if (type == 3) { p.type[cur] = 1; p.galaxyPart[cur] = 2; }
            else if (type == 0) { p.type[cur] = 2; p.galaxyPart[cur] = 1; }
            else { p.type[cur] = 1; p.galaxyPart[cur] = 1; }
            p.mass[cur] = individual ? mass[cur] * MSUN * 1e10 : h.massarr[type] * MSUN * 1e10;
            p.x[cur] = (double)pos[3 * cur] * KPC; p.y[cur] = (double)pos[3 * cur + 1] * KPC; p.z[cur] = (double)pos[3 * cur + 2] * KPC;
            p.vx[cur] = (double)vel[3 * cur] * KMS; p.vy[cur] = (double)vel[3 * cur + 1] * KMS; p.vz[cur] = (double)vel[3 * cur + 2] * KMS;
            if (type == 0) p.U[cur] = u[gas++] * 1e6;
        }
    return true;
}

// Gadget-2 SnapFormat 1 exactly as the reference writes it (DataManager.cpp:263-417), quirks included so that files
// are byte-identical: disk stars (type 1, galaxyPart 1) are counted as Gadget type 2 but stored in the leading block
// together with the gas (so they also get a U entry), bulge stars (type 1, part 2) come second and are counted as type
// 3, halo particles (type 3, part 3) come third and are counted as type 1, anything else is only counted (as type 0);
// the VEL block holds position / (km/s) because the reference's block helper always reads `position` (:333-340, :363).
bool save_gadget(const std::string& path, const Particles& p, int64_t count, double currentTime)
{
    constexpr double KPC = 3.08567758149137e19, MSUN = 1.98847e30;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); return false; }
    GadgetHeader h;
    memset(&h, 0, sizeof(h));
    std::vector<int64_t> lead, second, third;
    for (int64_t i = 0; i < count; i++) {
        int gt = 0;
        if (p.galaxyPart[i] == 1 && p.type[i] == 1) { gt = 2; lead.push_back(i); }
        if (p.galaxyPart[i] == 3 && p.type[i] == 3) { gt = 1; third.push_back(i); }
        if (p.galaxyPart[i] == 2 && p.type[i] == 1) { gt = 3; second.push_back(i); }
        if (p.type[i] == 2) { gt = 0; lead.push_back(i); }
        h.npart[gt]++; h.npartTotal[gt]++;
    }
    h.time = currentTime; h.num_files = 1; h.HubbleParam = 1.0;
    std::vector<int64_t> order(lead);
    order.insert(order.end(), second.begin(), second.end());
    order.insert(order.end(), third.begin(), third.end());
    auto block = [&](const void* data, size_t bytes) { uint32_t b = (uint32_t)bytes; fwrite(&b, 4, 1, f); fwrite(data, 1, bytes, f); fwrite(&b, 4, 1, f); };
    block(&h, sizeof(h));
    std::vector<float> v(order.size() * 3);
    for (double unit : {(double)(float)KPC, (double)(float)KMS}) {          // the reference's helper takes the unit as a float (:333)
        for (size_t k = 0; k < order.size(); k++) {
            const int64_t i = order[k];
            v[3 * k] = (float)(p.x[i] / unit); v[3 * k + 1] = (float)(p.y[i] / unit); v[3 * k + 2] = (float)(p.z[i] / unit);
        }
        block(v.data(), v.size() * sizeof(float));
    }
    std::vector<uint32_t> ids(order.size());
    for (size_t k = 0; k < order.size(); k++) ids[k] = p.id[order[k]];
    block(ids.data(), ids.size() * sizeof(uint32_t));
    std::vector<float> m(order.size());
    for (size_t k = 0; k < order.size(); k++) m[k] = (float)(p.mass[order[k]] / (MSUN * 1e10));
    block(m.data(), m.size() * sizeof(float));
    if (!lead.empty()) {
        std::vector<float> u(lead.size());
        for (size_t k = 0; k < lead.size(); k++) u[k] = (float)(p.U[lead[k]] / 1e6);
        block(u.data(), u.size() * sizeof(float));
    }
    fclose(f);
    return true;
}

// DataManager::saveData's format switch (DataManager.cpp:100-110): <outdir>/<timeStep><ending>.
bool save_snapshot(const std::string& outdir, const std::string& fmt, int timeStep, const Particles& p, int64_t count, double deltaTime, double endTime, double currentTime)
{
    const std::string stem = outdir + "/" + std::to_string(timeStep);
    if (fmt == "age") return save_age(stem + ".age", p, count, deltaTime, endTime, currentTime);
    if (fmt == "ag") return save_ag(stem + ".ag", p, count, deltaTime, endTime, currentTime);
    if (fmt == "agc") return save_agc(stem + ".agc", p, count, deltaTime, endTime, currentTime);
    if (fmt == "gadget") return save_gadget(stem + ".gadget", p, count, currentTime);
    fprintf(stderr, "Unknown output data format: %s\n", fmt.c_str());       // hdf5 is an empty stub in the reference (:259-262)
    return false;
}

// "makeGal" = Gadget SnapFormat 2 as the reference reads it (DataManager.cpp:580-810): every block is preceded by a
// 16-byte label record (size, 4-char label, next-block size, size); fixed block order POS VEL ID MASS U RHO HSML, of which
// RHO and HSML are skipped; MASS holds entries only for types whose header mass is 0, U only for gas.  Like the reference,
// the reader always consumes a MASS block header, so on files WITHOUT a MASS block (its own Example/galaxy_gas.dat) the U
// values come out shifted by six floats -- reproduced on purpose: the arrays must equal the reference loader's.  The reference
// shuffles the particles afterwards with a random_device seed (DataManager.cpp:780-783); file order is kept here.
bool load_makegal(const std::string& path, Particles& p)
{
    constexpr double KPC = 3.08567758149137e19, MSUN = 1.98847e30;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { perror(path.c_str()); return false; }
    struct Header { int32_t npart[6]; double mass[6]; double time, redshift; int32_t flag_sfr, flag_feedback, npartTotal[6], flag_cooling, num_files;
                    double BoxSize, Omega0, OmegaLambda, HubbleParam; char fill[256 - 6 * 4 - 6 * 8 - 2 * 8 - 2 * 4 - 6 * 4 - 2 * 4 - 4 * 8]; } h;
    static_assert(sizeof(Header) == 256, "makeGal header");
    char lab[16]; int32_t sz = 0;
    bool ok = fread(lab, 1, 16, f) == 16 && fread(&sz, 4, 1, f) == 1 && fread(&h, sizeof(h), 1, f) == 1 && fread(&sz, 4, 1, f) == 1;
    size_t total = 0;
    for (int i = 0; i < 6; i++) total += (size_t)h.npart[i];
    p.resize((int64_t)total);
    std::vector<float> pos(3 * total), vel(3 * total), mass(total), u(total, 0.f);
    std::vector<uint32_t> ids(total);
    for (int bnr = 0; ok && bnr < 7; bnr++) {
        ok = fread(lab, 1, 16, f) == 16 && fread(&sz, 4, 1, f) == 1;
        if (!ok) break;
        if (bnr == 0) ok = fread(pos.data(), 4, 3 * total, f) == 3 * total;
        else if (bnr == 1) ok = fread(vel.data(), 4, 3 * total, f) == 3 * total;
        else if (bnr == 2) ok = fread(ids.data(), 4, total, f) == total;
        else if (bnr == 3) {
            size_t idx = 0;
            for (int t = 0; t < 6 && ok; t++)
                for (int i = 0; i < h.npart[t] && ok; i++, idx++) {
                    if (h.mass[t] == 0 && h.npart[t] > 0) ok = fread(&mass[idx], 4, 1, f) == 1;
                    else mass[idx] = (float)h.mass[t];
                }
        } else if (bnr == 4) ok = h.npart[0] == 0 || fread(u.data(), 4, (size_t)h.npart[0], f) == (size_t)h.npart[0];
        else break;                                  // RHO, HSML: skipped by the reference, nothing after them is read
        ok = ok && fread(&sz, 4, 1, f) == 1;
    }
    fclose(f);
    if (!ok) { fprintf(stderr, "malformed makeGal file %s\n", path.c_str()); return false; }
    size_t idx = 0;
    for (int t = 0; t < 6; t++)
        for (int i = 0; i < h.npart[t]; i++, idx++) {
            p.x[idx] = (double)pos[3 * idx] * KPC; p.y[idx] = (double)pos[3 * idx + 1] * KPC; p.z[idx] = (double)pos[3 * idx + 2] * KPC;
            p.vx[idx] = (double)vel[3 * idx] * KMS; p.vy[idx] = (double)vel[3 * idx + 1] * KMS; p.vz[idx] = (double)vel[3 * idx + 2] * KMS;
            p.id[idx] = ids[idx];
            p.mass[idx] = mass[idx] * MSUN * 1e10;
            p.U[idx] = (t == 0 ? u[idx] : 0.0f) * 1e6;
            p.type[idx] = t == 0 ? 2 : t == 1 ? 3 : 1;
            p.galaxyPart[idx] = t == 1 ? 3 : (t == 3 || t == 5) ? 2 : 1;
        }
    return true;
}

// this repo's test format (oracle/agio.py): "AGPART01", int64 N, 13 double columns, uint8 type
const char* AGP_COLS[13] = {"x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "rho", "P", "T", "mu"};
std::vector<double>* agp_col(Particles& p, int k)
{
    std::vector<double>* c[13] = {&p.x, &p.y, &p.z, &p.vx, &p.vy, &p.vz, &p.mass, &p.U, &p.next, &p.rho, &p.P, &p.T, &p.mu};
    return c[k];
}
bool load_agp(const std::string& path, Particles& p)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { perror(path.c_str()); return false; }
    char magic[8]; int64_t n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "AGPART01", 8) || fread(&n, 8, 1, f) != 1) { fclose(f); return false; }
    p.resize(n);
    for (int k = 0; k < 13; k++) if (n && fread(agp_col(p, k)->data(), 8, (size_t)n, f) != (size_t)n) { fclose(f); return false; }
    if (n && fread(p.type.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); return false; }
    fclose(f);
    for (int64_t i = 0; i < n; i++) p.id[i] = (uint32_t)i;
    return true;
}
bool save_agp(const std::string& path, Particles& p)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); return false; }
    fwrite("AGPART01", 1, 8, f); fwrite(&p.n, 8, 1, f);
    for (int k = 0; k < 13; k++) fwrite(agp_col(p, k)->data(), 8, (size_t)p.n, f);
    fwrite(p.type.data(), 1, (size_t)p.n, f);
    fclose(f);
    // companion file with the fields .agp does not carry
    FILE* g = fopen((path + ".acc").c_str(), "wb");
    if (!g) return false;
    for (auto* v : {&p.ax, &p.ay, &p.az, &p.dUdt, &p.h, &p.vis, &p.next, &p.timeStep}) fwrite(v->data(), 8, (size_t)p.n, g);
    fclose(g);
    return true;
}

struct PhaseLog {                                            // File/Log.cpp:175-221
    std::string path; std::string current; std::chrono::steady_clock::time_point t0;
    void start(const std::string& name) { end(); current = name; t0 = std::chrono::steady_clock::now(); }
    void end()
    {
        if (current.empty() || path.empty()) { current.clear(); return; }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::ostringstream os; os << s;
        std::string num = os.str();
        std::replace(num.begin(), num.end(), '.', ',');
        std::ofstream f(path, std::ios::app);
        f << current << ";" << num << "\n";
        current.clear();
    }
};

void check(agb_multi* m, int rc, const char* what)
{
    if (rc != AGB_OK) { fprintf(stderr, "agb200: %s failed: %s (%s)\n", what, agb_strerror(rc), agb_multi_last_error(m)); exit(3); }
}

struct Driver {
    Config cfg; Particles p; agb_multi* ctx = nullptr; PhaseLog log;     // one handle for --gpus N devices (N = 1: a plain context behind it)
    double globalTime = 0, visualDensityRadius = 0;
    bool device_resident = false;
    unsigned long long sf_seed = 1;                                      // --sf-seed (star formation draws; the same numbers in both loops)
    static double u01(unsigned long long seed, unsigned long long particle, double time)     // agb_u01 of the library (splitmix64 of seed, particle, time bits)
    {
        unsigned long long tb; memcpy(&tb, &time, 8);
        unsigned long long z = seed + particle * 0x9E3779B97F4A7C15ull + tb * 0xD1B54A32D192ED03ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }

    void force_path(bool first)
    {
        agb_particles in{};
        in.n = p.n; in.x = p.x.data(); in.y = p.y.data(); in.z = p.z.data(); in.vx = p.vx.data(); in.vy = p.vy.data(); in.vz = p.vz.data();
        in.mass = p.mass.data(); in.U = p.U.data(); in.next_time = p.next.data(); in.mu = p.mu.data(); in.type = p.type.data();
        in.rho = p.rho.data(); in.P = p.P.data(); in.T = p.T.data(); in.h = p.h.data(); in.dUdt = p.dUdt.data();
        in.ax = p.ax.data(); in.ay = p.ay.data(); in.az = p.az.data();
        log.start("build tree");
        check(ctx, agb_multi_set_particles(ctx, &in), "set_particles");
        double R = 0;
        check(ctx, agb_multi_build_tree(ctx, &R), "build_tree");
        if (first) visualDensityRadius = R / 100000;                    // Simulation.cpp:126
        log.start("Visual Density");
        check(ctx, agb_multi_visual_density(ctx, visualDensityRadius), "visual_density");
        log.start("SPH density and update");
        check(ctx, agb_multi_gas_density(ctx, cfg.massInH), "gas_density");
        log.start("Force Calculation");
        check(ctx, agb_multi_forces(ctx, globalTime, cfg.e0, cfg.theta), "forces");
        agb_results out{p.ax.data(), p.ay.data(), p.az.data(), p.dUdt.data(), p.h.data(), p.rho.data(), p.P.data(), p.T.data(), p.vis.data()};
        check(ctx, agb_multi_get_results(ctx, &out), "get_results");
        log.end();
    }

    void upload()
    {
        agb_particles in{};
        in.n = p.n; in.x = p.x.data(); in.y = p.y.data(); in.z = p.z.data(); in.vx = p.vx.data(); in.vy = p.vy.data(); in.vz = p.vz.data();
        in.mass = p.mass.data(); in.U = p.U.data(); in.next_time = p.next.data(); in.mu = p.mu.data(); in.type = p.type.data();
        in.rho = p.rho.data(); in.P = p.P.data(); in.T = p.T.data(); in.h = p.h.data(); in.dUdt = p.dUdt.data();
        in.ax = p.ax.data(); in.ay = p.ay.data(); in.az = p.az.data();
        check(ctx, agb_multi_set_particles(ctx, &in), "set_particles");
    }
    void log_ms(const std::string& name, double ms)
    {
        if (log.path.empty()) return;
        std::ostringstream os; os << ms * 1e-3;
        std::string num = os.str();
        std::replace(num.begin(), num.end(), '.', ',');
        std::ofstream f(log.path, std::ios::app);
        f << name << ";" << num << "\n";
    }
    void device_force_path(bool first)
    {
        if (!first) {
            // steady state: the four calls in one (a single host synchronisation); the reference's phase rows come from the device timers
            log.end();
            check(ctx, agb_multi_force_path(ctx, visualDensityRadius, cfg.massInH, globalTime, cfg.e0, cfg.theta, nullptr), "force_path");
            double ms[5] = {0, 0, 0, 0, 0};
            agb_ctx* c0 = nullptr;
            agb_multi_context(ctx, 0, &c0);
            agb_get_phase_ms(c0, ms);
            log_ms("build tree", ms[0]); log_ms("Visual Density", ms[1]); log_ms("SPH density and update", ms[2]); log_ms("Force Calculation", ms[4]);
            return;
        }
        log.start("build tree");
        double R = 0;
        check(ctx, agb_multi_build_tree(ctx, &R), "build_tree");
        if (first) visualDensityRadius = R / 100000;
        log.start("Visual Density");
        check(ctx, agb_multi_visual_density(ctx, visualDensityRadius), "visual_density");
        log.start("SPH density and update");
        check(ctx, agb_multi_gas_density(ctx, cfg.massInH), "gas_density");
        log.start("Force Calculation");
        check(ctx, agb_multi_forces(ctx, globalTime, cfg.e0, cfg.theta), "forces");
        log.end();
    }
    void download()
    {
        agb_results out{p.ax.data(), p.ay.data(), p.az.data(), p.dUdt.data(), p.h.data(), p.rho.data(), p.P.data(), p.T.data(), p.vis.data()};
        check(ctx, agb_multi_get_results(ctx, &out), "get_results");
        check(ctx, agb_multi_get_state(ctx, p.x.data(), p.y.data(), p.z.data(), p.vx.data(), p.vy.data(), p.vz.data(), p.U.data(), p.next.data(), p.timeStep.data()), "get_state");
        if (cfg.cooling || cfg.starFormation) check(ctx, agb_multi_get_subgrid_state(ctx, p.type.data(), p.sfr.data()), "get_subgrid_state");
    }
    int run_device(int64_t max_steps, const std::string& outdir)
    {
        const double fixedStep = cfg.endTime / cfg.fixedTimeSteps;
        globalTime = 0.0;
        for (int64_t i = 0; i < p.n; i++) p.next[i] = 0.0;
        upload();
        check(ctx, agb_multi_integrator_init(ctx, cfg.eta, cfg.minTimeStep, cfg.maxTimeStep, cfg.H0, cfg.e0), "integrator_init");
        device_force_path(true);
        if (!outdir.empty()) { download(); save_snapshot(outdir, cfg.outputDataFormat, 0, p, (int64_t)cfg.numParticlesOutput, fixedStep, cfg.endTime, 0.0); }
        check(ctx, agb_multi_integrator_assign_all(ctx), "assign_all");
        double nextSaveTime = fixedStep;
        int64_t step = 0;
        while (globalTime < cfg.endTime && (max_steps < 0 || step < max_steps)) {
            log.start("first kick");
            check(ctx, agb_multi_step_begin(ctx, &globalTime), "step_begin");
            device_force_path(false);
            log.start("second kick");
            check(ctx, agb_multi_step_end(ctx), "step_end");
            log.end();
            step++;
            if (!outdir.empty() && globalTime >= nextSaveTime) {
                log.start("Save data");
                download();
                save_snapshot(outdir, cfg.outputDataFormat, (int)(nextSaveTime / fixedStep), p, (int64_t)cfg.numParticlesOutput, fixedStep, cfg.endTime, globalTime);
                log.end();
                nextSaveTime += fixedStep;
            }
        }
        download();
        printf("steps %lld globalTime %.17g\n", (long long)step, globalTime);
        return 0;
    }

    void assign_timestep(int64_t i)
    {                                                                   // Simulation.cpp:196-207 / :222-232
        double a = std::sqrt(p.ax[i] * p.ax[i] + p.ay[i] * p.ay[i] + p.az[i] * p.az[i]);
        if (a > 0) {
            double ts = cfg.eta * std::sqrt(cfg.e0 / a);
            ts = std::clamp(ts, cfg.minTimeStep, cfg.maxTimeStep);
            p.timeStep[i] = std::max(std::pow(2, std::floor(std::log2(ts))), cfg.minTimeStep);
        } else p.timeStep[i] = cfg.minTimeStep;
        p.next[i] = globalTime + p.timeStep[i];
    }
    void kick(int64_t i, double dt)
    {                                                                   // TimeIntegration.cpp:10-19
        if (std::isnan(p.ax[i]) || std::isnan(p.ay[i]) || std::isnan(p.az[i])) return;
        p.vx[i] = p.vx[i] + p.ax[i] * dt / 2; p.vy[i] = p.vy[i] + p.ay[i] * dt / 2; p.vz[i] = p.vz[i] + p.az[i] * dt / 2;
    }

    int run(int64_t max_steps, const std::string& outdir)
    {
        const int64_t n = p.n;
        const double fixedStep = cfg.endTime / cfg.fixedTimeSteps;
        for (int64_t i = 0; i < n; i++) if (p.type[i] == 2) p.T[i] = (GAMMA - 1.0) * p.U[i] * PRTN * p.mu[i] / K_B;   // Simulation.cpp:108-112
        globalTime = 0.0;
        for (int64_t i = 0; i < n; i++) p.next[i] = 0.0;              // Particle::nextIntegrationTime defaults to 0 (Particle.h:29)
        force_path(true);                                                // Simulation.cpp:120-139
        if (!outdir.empty()) save_snapshot(outdir, cfg.outputDataFormat, 0, p, (int64_t)cfg.numParticlesOutput, fixedStep, cfg.endTime, 0.0);
        double nextSaveTime = fixedStep;
        for (int64_t i = 0; i < n; i++) p.next[i] = 0.0;
#pragma omp parallel for
        for (int64_t i = 0; i < n; i++) assign_timestep(i);
        const double H0SI = (cfg.H0 * KMS) / MPC;
        int64_t step = 0;
        while (globalTime < cfg.endTime && (max_steps < 0 || step < max_steps)) {
#pragma omp parallel for
            for (int64_t i = 0; i < n; i++) if (globalTime >= p.next[i]) assign_timestep(i);
            double mn = std::numeric_limits<double>::max();
#pragma omp parallel for reduction(min : mn)
            for (int64_t i = 0; i < n; i++) if (p.next[i] < mn) mn = p.next[i];
            globalTime = mn;
            log.start("first kick");
#pragma omp parallel for
            for (int64_t i = 0; i < n; i++)
                if (globalTime == p.next[i]) {
                    const double dt = p.timeStep[i];
                    kick(i, dt);
                    p.x[i] = p.x[i] + p.vx[i] * dt; p.y[i] = p.y[i] + p.vy[i] * dt; p.z[i] = p.z[i] + p.vz[i] * dt;   // Drift
                }
            force_path(false);                                           // Simulation.cpp:275-285
            log.start("second kick");
#pragma omp parallel for
            for (int64_t i = 0; i < n; i++)
                if (globalTime == p.next[i]) {
                    const double dt = p.timeStep[i];
                    if (p.type[i] == 2) {
                        // the sub-grid hooks the reference calls from here (commented out in its source, Simulation.cpp:311-320)
                        if (cfg.cooling) {                                // Cooling.cpp:6-25
                            const double rate = 1.42e-27 * 1.1 * std::sqrt(p.T[i]) * 1e6 * 1e6 * 1e-7;
                            if (rate > 0 && p.rho[i] > 0) p.dUdt[i] -= rate / p.rho[i];
                        }
                        if (cfg.starFormation && p.rho[i] > 1e-22 && p.T[i] < 1e4) {   // SFR.cpp:12-34, counter-based deviate instead of rand()
                            const double pr = 1 - std::exp(-0.1 * dt / 1e15);
                            p.sfr[i] = pr;
                            if (u01(sf_seed, (unsigned long long)i, globalTime) < pr) { p.type[i] = 1; p.U[i] = 0.0; }
                        }
                    }
                    if (p.type[i] == 2) {                                // Ueuler, TimeIntegration.cpp:28-41
                        if (!std::isnan(p.dUdt[i])) p.U[i] += p.dUdt[i] * dt;
                        p.dUdt[i] = 0;
                    }
                    const double scale = std::exp(H0SI * dt);            // Simulation.cpp:330-332
                    p.x[i] *= scale; p.y[i] *= scale; p.z[i] *= scale;
                    kick(i, dt);
                    p.next[i] += dt;
                }
            log.end();
            step++;
            if (!outdir.empty() && globalTime >= nextSaveTime) {
                log.start("Save data");
                save_snapshot(outdir, cfg.outputDataFormat, (int)(nextSaveTime / fixedStep), p, (int64_t)cfg.numParticlesOutput, fixedStep, cfg.endTime, globalTime);
                log.end();
                nextSaveTime += fixedStep;
            }
        }
        printf("steps %lld globalTime %.17g\n", (long long)step, globalTime);
        return 0;
    }
};

} // namespace

int main(int argc, char** argv)
{
    std::map<std::string, std::string> opt;
    bool overwrite = false, device_resident = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--overwrite") { overwrite = true; continue; }
        if (a == "--device-resident") { device_resident = true; continue; }
        if (a.rfind("--", 0) == 0 && i + 1 < argc) { opt[a.substr(2)] = argv[++i]; continue; }
        fprintf(stderr, "usage: agb_sim --config Config.ini [--input-root DIR] [--output-root DIR] [--steps K] [--device D] [--gpus N] [--cores C] [--precision fp64|mixed] [--dump final.agp] [--overwrite]\n");
        return 2;
    }
    Driver d;
    if (!load_config(opt.count("config") ? opt["config"] : "../Config.ini", d.cfg)) return 2;
    const std::string inroot = opt.count("input-root") ? opt["input-root"] : "../../input_data/";
    const std::string inpath = inroot + (inroot.empty() || inroot.back() == '/' ? "" : "/") + d.cfg.inputPath;
    bool ok = d.cfg.inputDataFormat == "age" ? load_age(inpath, d.p) : d.cfg.inputDataFormat == "agp" ? load_agp(inpath, d.p) :
              d.cfg.inputDataFormat == "gadget" ? load_gadget(inpath, d.p) : d.cfg.inputDataFormat == "makeGal" ? load_makegal(inpath, d.p) :
              d.cfg.inputDataFormat == "ag" ? load_ag(inpath, d.p) : d.cfg.inputDataFormat == "agc" ? load_agc(inpath, d.p) : false;
    if (!ok) { fprintf(stderr, "cannot read initial conditions %s (format %s; supported: age, ag, agc, agp, gadget, makeGal)\n", inpath.c_str(), d.cfg.inputDataFormat.c_str()); return 2; }
    if (opt.count("convert-only")) return save_agp(opt["convert-only"], d.p) ? 0 : 2;      // no GPU involved
    if (opt.count("snapshot-only")) {                                                      // DIR/<n>.<outputDataFormat> of the loaded set, no GPU involved
        const int64_t count = std::min<int64_t>(d.cfg.numParticlesOutput > 0 ? (int64_t)d.cfg.numParticlesOutput : d.p.n, d.p.n);
        const int ts = opt.count("snapshot-index") ? atoi(opt["snapshot-index"].c_str()) : 0;
        return save_snapshot(opt["snapshot-only"], d.cfg.outputDataFormat, ts, d.p, count, d.cfg.endTime / d.cfg.fixedTimeSteps, d.cfg.endTime,
                             opt.count("snapshot-time") ? atof(opt["snapshot-time"].c_str()) : 0.0) ? 0 : 2;
    }
    if ((int64_t)d.cfg.numberOfParticles != d.p.n) {                     // Simulation.cpp:92-98
        fprintf(stderr, "Error: Number of particles in the ConfigFile (%lld) does not match the data file (%lld).\n", (long long)d.cfg.numberOfParticles, (long long)d.p.n);
        return 2;
    }
    std::string outdir;
    if (opt.count("output-root")) {
        outdir = opt["output-root"] + "/" + d.cfg.outputFolderName;
        struct stat st;
        if (stat(outdir.c_str(), &st) == 0 && !overwrite) { fprintf(stderr, "output folder %s exists (use --overwrite)\n", outdir.c_str()); return 2; }
        std::string cmd = "mkdir -p '" + outdir + "/logs'";
        if (system(cmd.c_str()) != 0) return 2;
        d.log.path = outdir + "/logs/processLog.csv";
        std::ofstream(d.log.path, std::ios::trunc);
    }
    const int cores = opt.count("cores") ? atoi(opt["cores"].c_str()) : 8;
    // --gpus N: devices first .. first+N-1 (first = --device, default 0); every device builds the tree, each walks 1/N of the targets
    const int ngpus = opt.count("gpus") ? std::max(1, atoi(opt["gpus"].c_str())) : 1, first = opt.count("device") ? atoi(opt["device"].c_str()) : 0;
    std::vector<int> devs;
    for (int i = 0; i < ngpus; i++) devs.push_back(first + i);
    if (opt.count("devices")) {                                          // explicit list, e.g. --devices 0,2,4,6
        devs.clear();
        const std::string s = opt["devices"];
        for (size_t a = 0; a < s.size();) { size_t b = s.find(',', a); if (b == std::string::npos) b = s.size(); if (b > a) devs.push_back(atoi(s.substr(a, b - a).c_str())); a = b + 1; }
        if (devs.empty()) devs.push_back(0);
    }
    int rc = agb_multi_create(&d.ctx, devs.data(), (int)devs.size(), cores);
    if (rc != AGB_OK) { fprintf(stderr, "agb200: %s\n", agb_strerror(rc)); return 3; }
    if (opt.count("precision")) agb_multi_set_option(d.ctx, AGB_OPT_PRECISION, opt["precision"] == "fp64" ? 0 : 1);
    if (opt.count("sf-seed")) d.sf_seed = std::max(1ull, strtoull(opt["sf-seed"].c_str(), nullptr, 10));
    agb_multi_set_option(d.ctx, AGB_OPT_COOLING, d.cfg.cooling ? 1 : 0);
    agb_multi_set_option(d.ctx, AGB_OPT_STAR_FORMATION, d.cfg.starFormation ? (int64_t)d.sf_seed : 0);
    const int64_t max_steps = opt.count("steps") ? atoll(opt["steps"].c_str()) : -1;
    rc = device_resident ? d.run_device(max_steps, outdir) : d.run(max_steps, outdir);
    if (opt.count("dump")) save_agp(opt["dump"], d.p);
    agb_multi_destroy(d.ctx);
    return rc;
}

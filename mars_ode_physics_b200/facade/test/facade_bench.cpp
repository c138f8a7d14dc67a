// facade_bench.cpp -- BASELINE config D (hinge2 rover on the shared heightfield) built and stepped ENTIRELY through
// the plugin API a MARS host uses: WorldPhysicsLoader::createWorldInstance (reference src/WorldPhysicsLoader.cpp:
// 36-40), PhysicsInterface::createFrame / createObject / createJoint with ConfigMaps (src/WorldPhysics.cpp:515-655),
// stepTheWorld (:349-421) -- with the batch extension for what the one-world reference API cannot express
// (num_worlds replicas, device collision, per-world commands and observations as arrays).
//
//   facade_bench <worlds> <steps> <warmup> [settle] [dump_file]
//
// Timed loop per step (pipelined control loop of depth 2, as bench.py's e2e leg): per-world wheel commands host ->
// device from pinned memory, stepTheWorld(), chassis state + all joint angles / rates device -> host.  Prints one
// JSON line with the end-to-end world-steps/s; optionally dumps the final body states (float[nb][13][worlds]) so
// that the GPU tests can compare this path with the same scene built through the C ABI.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../src/Joints.hpp"
#include "../src/Objects.hpp"
#include "../src/WorldPhysics.hpp"
#include "../src/WorldPhysicsLoader.hpp"

using namespace mars::ode_physics;
using namespace mars::interfaces;
using namespace mars::utils;
using namespace configmaps;

// mars_ode_physics_b200/scenes.py: splitmix64 / uniform01 (per-world randomisation keyed by the world id)
static uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double uniform01(uint64_t tag, uint64_t world, uint64_t k)
{
    return (double)(splitmix64((tag ^ world) + 0x1000003ull * k) >> 11) / (double)(1ull << 53);
}

// scenes.terrain_heights: sum of 4 seeded sinusoids, amplitude 0.1 m, 256 x 256 @ 0.1 m
static std::vector<float> terrain(int nx, int ny, double dx, double *x0, double *y0)
{
    std::vector<double> h((size_t)nx * ny, 0.0);
    std::vector<float> out((size_t)nx * ny);
    *x0 = (0 - nx / 2.0) * dx;
    *y0 = (0 - ny / 2.0) * dx;
    for (int k = 0; k < 4; k++) {
        const double kx = 0.5 + 2.5 * uniform01(0xD, 100 + k, 1), ky = 0.5 + 2.5 * uniform01(0xD, 100 + k, 2);
        const double ph = 2 * M_PI * uniform01(0xD, 100 + k, 3);
        for (int iy = 0; iy < ny; iy++)
            for (int ix = 0; ix < nx; ix++)
                h[(size_t)iy * nx + ix] += std::sin(kx * ((ix - nx / 2.0) * dx) + ky * ((iy - ny / 2.0) * dx) + ph);
    }
    for (size_t i = 0; i < h.size(); i++) out[i] = (float)(h[i] * (0.1 / 4.0));
    return out;
}

static std::shared_ptr<DynamicObject> part(PhysicsInterface &wp, const std::string &name, const char *type, double ex,
                                           double ey, double ez, double mass, double x, double y, double z)
{
    ConfigMap map;
    map["name"] << name;
    map["parentFrame"] << name;
    map["type"] = type;
    map["extend"]["x"] = ex;
    map["extend"]["y"] = ey;
    map["extend"]["z"] = ez;
    map["position"]["x"] = x;
    map["position"]["y"] = y;
    map["position"]["z"] = z;
    map["orientation"]["x"] = 0.0;
    map["orientation"]["y"] = 0.0;
    map["orientation"]["z"] = 0.0;
    map["orientation"]["w"] = 1.0;
    map["density"] = 1.0;
    map["mass"] = mass;
    map["c_params"]["friction1"] = 1.0;
    map["c_params"]["cfm"] = 1e-5;
    map["c_params"]["erp"] = 0.2;
    auto frame = wp.createFrame(nullptr, map);
    wp.createObject(map);
    return frame;
}

int main(int argc, char **argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 4096, steps = argc > 2 ? atoi(argv[2]) : 100;
    const int warmup = argc > 3 ? atoi(argv[3]) : 5, settle = argc > 4 ? atoi(argv[4]) : 600;
    const char *dump = argc > 5 ? argv[5] : nullptr;
    const double z0 = 0.25;
    WorldPhysicsLoader loader(nullptr);
    loader.setNumWorlds(W);
    std::shared_ptr<PhysicsInterface> wpi = loader.createWorldInstance();
    WorldPhysics *wp = dynamic_cast<WorldPhysics *>(wpi.get());
    wp->setMaxContacts(8);
    wp->setRollingFriction(false); // the workload has no rolling friction: keep the row capacity at 3 per contact
    try {
        wp->initTheWorld();
    } catch (const std::exception &e) {
        fprintf(stderr, "facade_bench: %s\n", e.what());
        return 2;
    }
    wp->step_size = 0.001;
    wp->fast_step = true;
    wp->ground_friction = 1.0; // surface of the terrain (scenes.py: mu 1.0, soft_erp 0.2, soft_cfm 1e-5)
    wp->ground_cfm = 1e-5;
    wp->ground_erp = 0.2;
    wp->setDeviceCollision(true);
    b2p_arena *arena = wp->getWorld();

    std::vector<std::shared_ptr<DynamicObject>> frames;
    frames.push_back(part(*wpi, "chassis", "box", 0.6, 0.4, 0.15, 10.0, 0.0, 0.0, z0));
    std::vector<std::shared_ptr<JointInterface>> joints;
    int wi = 0;
    for (double sx : {1.0, -1.0})
        for (double sy : {1.0, -1.0}) {
            const std::string name = "wheel" + std::to_string(wi++);
            const double x = sx * 0.25, y = sy * 0.25, z = z0 - 0.1;
            frames.push_back(part(*wpi, name, "sphere", 0.1, 0.1, 0.1, 1.0, x, y, z));
            ConfigMap jm;
            jm["name"] << ("axle_" + name);
            jm["type"] = "hinge2";
            jm["parent_link_name"] << std::string("chassis");
            jm["child_link_name"] << name;
            jm["anchor"]["x"] = x;
            jm["anchor"]["y"] = y;
            jm["anchor"]["z"] = z;
            jm["axis1"]["x"] = 0.0;
            jm["axis1"]["y"] = 0.0;
            jm["axis1"]["z"] = 1.0;
            jm["axis2"]["x"] = 0.0;
            jm["axis2"]["y"] = 1.0;
            jm["axis2"]["z"] = 0.0;
            auto j = wpi->createJoint(nullptr, jm);
            if (!j) {
                fprintf(stderr, "facade_bench: createJoint failed\n");
                return 2;
            }
            joints.push_back(j);
            Joint *jj = dynamic_cast<Joint *>(j.get());
            // steer axis locked by stops, drive axis motorised, suspension (scenes.rover); the suspension has no
            // JointInterface setter in the reference, so it goes through the C ABI underneath
            const int id = jj->getJointId();
            b2p_joint_set_param(arena, id, B2P_ALL_WORLDS, B2P_PARAM_LO_STOP, 0.0);
            b2p_joint_set_param(arena, id, B2P_ALL_WORLDS, B2P_PARAM_HI_STOP, 0.0);
            j->setForceLimit2(5.0);
            j->setVelocity2(4.0);
            b2p_joint_set_param(arena, id, B2P_ALL_WORLDS, B2P_PARAM_SUSPENSION_ERP, 0.4);
            b2p_joint_set_param(arena, id, B2P_ALL_WORLDS, B2P_PARAM_SUSPENSION_CFM, 0.01);
        }
    double hx0, hy0;
    const std::vector<float> hts = terrain(256, 256, 0.1, &hx0, &hy0);
    wp->createHeightfield(hts.data(), 256, 256, 0.1, hx0, hy0);

    // per-world placement and wheel speed (bench.build_rover_worlds): dx, dy in [-8, 8), z + 0.12, vel in [2, 6)
    const int nb = 5, nj = 4;
    std::vector<float> st((size_t)nb * 13 * W);
    if (b2p_state_download(arena, st.data()) < 0) {
        fprintf(stderr, "facade_bench: %s\n", b2p_last_error());
        return 2;
    }
    float *cmd[4];
    for (int i = 0; i < 4; i++) cmd[i] = (float *)b2p_host_alloc(sizeof(float) * (size_t)nj * W);
    for (int w = 0; w < W; w++) {
        const double dx = 16.0 * uniform01(0xD, (uint64_t)w, 1) - 8.0, dy = 16.0 * uniform01(0xD, (uint64_t)w, 2) - 8.0;
        const float vel = (float)(2.0 + 4.0 * uniform01(0xD, (uint64_t)w, 3));
        for (int b = 0; b < nb; b++) {
            float *s = st.data() + (size_t)b * 13 * W;
            s[0 * (size_t)W + w] += (float)dx;
            s[1 * (size_t)W + w] += (float)dy;
            s[2 * (size_t)W + w] += 0.12f;
        }
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < nj; j++) cmd[i][(size_t)j * W + w] = vel + 0.001f * i;
    }
    b2p_state_upload(arena, st.data());
    wp->setJointVelocitiesBatch(cmd[0], nj, 2);
    float *obs[2], *jobs[2];
    for (int i = 0; i < 2; i++) {
        obs[i] = (float *)b2p_host_alloc(sizeof(float) * 13 * (size_t)W);
        jobs[i] = (float *)b2p_host_alloc(sizeof(float) * (size_t)nj * 4 * W);
    }

    for (int k = 0; k < settle; k++) wpi->stepTheWorld();
    int pending = -1;
    auto tick = [&](int k) {
        wp->setJointVelocitiesBatch(cmd[k % 4], nj, 2);
        wpi->stepTheWorld();
        const int t = wp->observeBatchBegin(frames[0], obs[k & 1], jobs[k & 1]);
        if (pending >= 0) wp->observeBatchWait(pending);
        pending = t;
    };
    for (int k = 0; k < warmup; k++) tick(k);
    wp->observeBatchWait(pending);
    pending = -1;
    b2p_sync(arena);
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < steps; k++) tick(k);
    wp->observeBatchWait(pending);
    b2p_sync(arena);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // sanity: finite observations, the rovers are on the terrain, world 0's contacts are readable as ContactData
    bool ok = true;
    double zmin = 1e9, zmax = -1e9;
    for (int i = 0; i < 2; i++)
        for (size_t q = 0; q < 13 * (size_t)W; q++) ok &= std::isfinite(obs[i][q]) != 0;
    for (int w = 0; w < W; w++) {
        zmin = std::min(zmin, (double)obs[(steps - 1) & 1][2 * (size_t)W + w]);
        zmax = std::max(zmax, (double)obs[(steps - 1) & 1][2 * (size_t)W + w]);
    }
    ok &= zmin > -0.2 && zmax < 0.6;
    std::vector<ContactData> contacts; // of the first worlds (a single rover may be airborne at this instant)
    for (int w = 0; w < W && w < 16; w++) {
        const std::vector<ContactData> cw = wp->getContacts(w);
        contacts.insert(contacts.end(), cw.begin(), cw.end());
    }
    size_t named = 0;
    for (const ContactData &c : contacts) named += c.body1 != nullptr && c.depth >= 0 && c.c_params.friction1 == 1.0;
    ok &= named == contacts.size();
    Vector cf;
    if (!contacts.empty() && contacts[0].body1) {
        const Frame *f = dynamic_cast<const Frame *>(contacts[0].body1.get());
        cf = f->getContactForceVector();
        ok &= std::isfinite(cf.z()) != 0;
    }
    if (dump) {
        if (b2p_state_download(arena, st.data()) == 0) {
            FILE *fp = fopen(dump, "wb");
            if (fp) {
                fwrite(st.data(), sizeof(float), st.size(), fp);
                fclose(fp);
            }
        }
    }
    printf("{\"e2e_facade\": %.1f, \"unit\": \"world-steps/s\", \"worlds\": %d, \"steps\": %d, \"ms_per_step\": %.6f, "
           "\"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu, \"chassis_z_min\": %.4f, \"chassis_z_max\": %.4f, "
           "\"contacts_first_worlds\": %zu, \"contact_force_z\": %.4f, \"launches\": %d, \"ok\": %s}\n",
           (double)W * steps / sec, W, steps, 1e3 * sec / steps, sizeof(float) * (size_t)nj * W,
           sizeof(float) * (size_t)(13 + 4 * nj) * W, zmin, zmax, contacts.size(), cf.z(), b2p_kernel_launch_count(arena),
           ok ? "true" : "false");
    for (int i = 0; i < 4; i++) b2p_host_free(cmd[i]);
    for (int i = 0; i < 2; i++) {
        b2p_host_free(obs[i]);
        b2p_host_free(jobs[i]);
    }
    return ok ? 0 : 1;
}

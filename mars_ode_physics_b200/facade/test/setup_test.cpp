// setup_test.cpp -- the scene-construction path of the plugin, driven like a MARS host:
// every object type's mass pipeline (reference src/objects/{Box,Sphere,Capsule,Cylinder,Inertial}.cpp createMass,
// Frame::addObject / addObjectMass with COM re-centring, src/objects/Frame.cpp:183-207,794-851) and the joint
// set-up (Joint::calculateCfmErp and the per-type createJoint: src/joints/Joint.cpp:147-273, HingeJoint.cpp:31-66,
// SliderJoint.cpp:32-65, Hinge2Joint.cpp:33-77, FixedJoint.cpp:31-47).  Prints what ended up in the device arena
// as one JSON document; tests/test_gpu_facade.py compares it with the oracle's restatement (oracle/orc_mars.c).
#include <cmath>
#include <cstdio>
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

struct Obj {
    const char *type;
    double ex, ey, ez, mass, x, y, z, qw, qx, qy, qz;
};

static void add_object(PhysicsInterface &wp, const std::string &frame, const Obj &o, ConfigMap *extra = nullptr)
{
    ConfigMap map;
    if (extra) map = *extra;
    map["name"] << (frame + "_" + o.type);
    map["parentFrame"] << frame;
    map["type"] = o.type;
    map["extend"]["x"] = o.ex;
    map["extend"]["y"] = o.ey;
    map["extend"]["z"] = o.ez;
    map["position"]["x"] = o.x;
    map["position"]["y"] = o.y;
    map["position"]["z"] = o.z;
    map["orientation"]["w"] = o.qw;
    map["orientation"]["x"] = o.qx;
    map["orientation"]["y"] = o.qy;
    map["orientation"]["z"] = o.qz;
    if (!map.hasKey("density")) map["density"] = 1.0;
    map["mass"] = o.mass;
    wp.createObject(map);
}

static std::shared_ptr<DynamicObject> frame(PhysicsInterface &wp, const std::string &name)
{
    ConfigMap map;
    map["name"] << name;
    return wp.createFrame(nullptr, map);
}

static void print_frame(WorldPhysics *wp, const std::shared_ptr<DynamicObject> &f, bool last)
{
    const Frame *fr = dynamic_cast<const Frame *>(f.get());
    sReal mass = 0, I[9];
    fr->getMass(&mass, I);
    Vector p;
    f->getPosition(&p);
    double bp[3] = {0, 0, 0};
    if (fr->getBody() >= 0) b2p_body_get_position(wp->getWorld(), fr->getBody(), 0, bp);
    printf("  \"%s\": {\"mass\": %.17g, \"I\": [%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g], "
           "\"frame_pos\": [%.17g,%.17g,%.17g], \"body_pos\": [%.9g,%.9g,%.9g]}%s\n",
           f->getName().c_str(), mass, I[0], I[1], I[2], I[3], I[4], I[5], I[6], I[7], I[8], p.x(), p.y(), p.z(), bp[0],
           bp[1], bp[2], last ? "" : ",");
}

static void print_joint(WorldPhysics *wp, const char *name, const std::shared_ptr<JointInterface> &j, bool last)
{
    b2p_arena *a = wp->getWorld();
    const int id = dynamic_cast<Joint *>(j.get())->getJointId();
    auto get = [&](int param) {
        double v = 0;
        b2p_joint_get_param(a, id, 0, param, &v);
        return v;
    };
    auto num = [](double v) { // JSON has no infinity
        if (std::isinf(v)) return std::string(v > 0 ? "1e999" : "-1e999");
        char buf[64];
        snprintf(buf, sizeof(buf), "%.17g", v);
        return std::string(buf);
    };
    double an[3], ax[3];
    b2p_joint_get_anchor(a, id, 0, an);
    b2p_joint_get_axis1(a, id, 0, ax);
    printf("  \"%s\": {\"fmax\": %s, \"vel\": %s, \"lo\": %s, \"hi\": %s, \"stop_cfm\": %s, \"stop_erp\": %s, \"cfm\": %s, "
           "\"fmax2\": %s, \"lo2\": %s, \"anchor\": [%.9g,%.9g,%.9g], \"axis1\": [%.9g,%.9g,%.9g], \"friction\": %s}%s\n",
           name, num(get(B2P_PARAM_FMAX)).c_str(), num(get(B2P_PARAM_VEL)).c_str(), num(get(B2P_PARAM_LO_STOP)).c_str(),
           num(get(B2P_PARAM_HI_STOP)).c_str(), num(get(B2P_PARAM_STOP_CFM)).c_str(), num(get(B2P_PARAM_STOP_ERP)).c_str(),
           num(get(B2P_PARAM_CFM)).c_str(), num(get(B2P_PARAM_FMAX | B2P_PARAM_GROUP2)).c_str(),
           num(get(B2P_PARAM_LO_STOP | B2P_PARAM_GROUP2)).c_str(), an[0], an[1], an[2], ax[0], ax[1], ax[2],
           num(j->getJointFrictionCoefficient()).c_str(), last ? "" : ",");
}

int main()
{
    WorldPhysicsLoader loader(nullptr);
    std::shared_ptr<PhysicsInterface> wpi = loader.createWorldInstance();
    WorldPhysics *wp = dynamic_cast<WorldPhysics *>(wpi.get());
    try {
        wpi->initTheWorld();
    } catch (const std::exception &e) {
        fprintf(stderr, "setup_test: %s\n", e.what());
        return 2;
    }
    wpi->step_size = 0.002;
    wpi->fast_step = true;
    const double s30 = std::sin(0.5 * 0.5235987755982988), c30 = std::cos(0.5 * 0.5235987755982988);
    std::vector<std::shared_ptr<DynamicObject>> fr;
    fr.push_back(frame(*wpi, "fs"));
    add_object(*wpi, "fs", {"sphere", 0.15, 0, 0, 2.0, 0.1, 0.2, 0.3, 1, 0, 0, 0});
    fr.push_back(frame(*wpi, "fc"));
    add_object(*wpi, "fc", {"capsule", 0.05, 0.4, 0, 1.5, 0.0, 0.0, 1.0, c30, s30, 0, 0});
    fr.push_back(frame(*wpi, "fy"));
    add_object(*wpi, "fy", {"cylinder", 0.1, 0.3, 0, 0.8, 0.5, 0.0, 0.5, c30, 0, s30, 0});
    fr.push_back(frame(*wpi, "fi"));
    {
        ConfigMap ex;
        const char *k[9] = {"i00", "i01", "i02", "i10", "i11", "i12", "i20", "i21", "i22"};
        const double v[9] = {0.02, 0.001, 0.0, 0.001, 0.03, 0.002, 0.0, 0.002, 0.025};
        for (int i = 0; i < 9; i++) ex["inertia"][k[i]] = v[i];
        add_object(*wpi, "fi", {"inertial", 0, 0, 0, 1.2, -0.3, 0.0, 0.2, 1, 0, 0, 0}, &ex);
    }
    fr.push_back(frame(*wpi, "fb"));
    {
        ConfigMap ex;
        ex["createFromBox"] = true;
        ex["x"] = 0.2;
        ex["y"] = 0.1;
        ex["z"] = 0.05;
        ex["density"] = 500.0; // the density branch (dMassSetBox, Inertial.cpp:37-41)
        add_object(*wpi, "fb", {"inertial", 0, 0, 0, 0.0, 0.0, 0.4, 0.2, 1, 0, 0, 0}, &ex);
    }
    fr.push_back(frame(*wpi, "fsph"));
    {
        ConfigMap ex;
        ex["createFromSphere"] = true;
        ex["x"] = 0.07;
        ex["density"] = 0.0; // mass branch
        add_object(*wpi, "fsph", {"inertial", 0, 0, 0, 0.6, 0.2, -0.4, 0.3, 1, 0, 0, 0}, &ex);
    }
    fr.push_back(frame(*wpi, "comp"));
    add_object(*wpi, "comp", {"box", 0.4, 0.2, 0.1, 1.0, 0.0, 0.0, 0.5, 1, 0, 0, 0});
    add_object(*wpi, "comp", {"sphere", 0.05, 0, 0, 0.3, 0.2, 0.0, 0.6, 1, 0, 0, 0});
    add_object(*wpi, "comp", {"cylinder", 0.04, 0.2, 0, 0.5, -0.1, 0.1, 0.45, c30, s30, 0, 0});

    auto joint = [&](const char *name, const char *type, const char *parent, const char *child, double ax, double ay,
                     double az, double damping, double stiffness, double jcfm, double friction) {
        ConfigMap jm;
        jm["name"] << std::string(name);
        jm["type"] = type;
        jm["parent_link_name"] << std::string(parent);
        jm["child_link_name"] << std::string(child);
        jm["anchor"]["x"] = ax;
        jm["anchor"]["y"] = ay;
        jm["anchor"]["z"] = az;
        jm["axis1"]["x"] = 0.0;
        jm["axis1"]["y"] = 1.0;
        jm["axis1"]["z"] = 0.0;
        jm["axis2"]["x"] = 0.0;
        jm["axis2"]["y"] = 0.0;
        jm["axis2"]["z"] = 1.0;
        if (damping >= 0) jm["springDamping"] = damping;
        if (stiffness >= 0) jm["springStiffness"] = stiffness;
        if (jcfm >= 0) jm["jointCFM"] = jcfm;
        if (friction >= 0) jm["frictionCoefficient"] = friction;
        return wpi->createJoint(nullptr, jm);
    };
    std::vector<std::pair<const char *, std::shared_ptr<JointInterface>>> jj;
    jj.push_back({"hinge_spring_damper", joint("j0", "hinge", "fs", "fc", 0.05, 0.1, 0.6, 0.5, 20.0, 1e-4, 0.02)});
    jj.push_back({"slider_spring", joint("j1", "slider", "fc", "fy", 0, 0, 0, -1, 15.0, -1, -1)});
    jj.push_back({"hinge2_damper", joint("j2", "hinge2", "comp", "fsph", 0.1, -0.2, 0.4, 0.3, -1, -1, -1)});
    jj.push_back({"fixed", joint("j3", "fixed", "fi", "fb", 0, 0, 0, -1, -1, -1, -1)});
    jj.push_back({"hinge_child_only", joint("j4", "hinge", "nonexistent", "fy", 0.5, 0.0, 0.7, -1, -1, -1, -1)});
    for (auto &p : jj)
        if (!p.second) {
            fprintf(stderr, "setup_test: joint %s was not created\n", p.first);
            return 1;
        }

    printf("{\"frames\": {\n");
    for (size_t i = 0; i < fr.size(); i++) print_frame(wp, fr[i], i + 1 == fr.size());
    printf(" },\n \"joints\": {\n");
    for (size_t i = 0; i < jj.size(); i++) print_joint(wp, jj[i].first, jj[i].second, i + 1 == jj.size());
    printf(" }}\n");
    return 0;
}

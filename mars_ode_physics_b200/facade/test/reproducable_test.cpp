// reproducable_test.cpp -- the scene of the reference's determinism test (reference
// test/reproducable_test.cpp: 13 frames, 8 hinges + 4 fixed joints, sine gait, hand-made foot
// contacts, 1 ms quickstep) expressed against the B200 facade through the reference's own
// interfaces (PhysicsInterface / DynamicObject / JointInterface / ContactData).
//
//   reproducable_test [csv_path] [t_end] [num_worlds]
//
// Writes "time_sec,body_x,body_y,body_z" rows with "%g" formatting like the reference.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../src/Joints.hpp"
#include "../src/Objects.hpp"
#include "../src/WorldPhysics.hpp"

using namespace mars::ode_physics;
using namespace mars::interfaces;
using namespace mars::utils;
using namespace configmaps;

namespace {

struct LinkSpec {
    const char *name;
    double extend_x, mass;
};

std::shared_ptr<DynamicObject> makeLink(WorldPhysics &wp, ConfigMap &map, const LinkSpec &s)
{
    // the reference helper overrides extend.y/z and the object position for every link
    map["name"] << std::string(s.name);
    map["parentFrame"] << std::string(s.name);
    map["type"] = "box";
    map["extend"]["x"] = s.extend_x;
    map["extend"]["y"] = 0.2;
    map["extend"]["z"] = 0.05;
    map["position"]["x"] = 0.0;
    map["position"]["y"] = 0.0;
    map["position"]["z"] = 0.265;
    map["orientation"]["x"] = 0.0;
    map["orientation"]["y"] = 0.0;
    map["orientation"]["z"] = 0.0;
    map["orientation"]["w"] = 1.0;
    map["density"] = 1.0;
    map["mass"] = s.mass;
    auto frame = wp.createFrame(nullptr, map);
    wp.createObject(map);
    return frame;
}

} // namespace

int main(int argc, char **argv)
{
    const std::string csv = argc > 1 ? argv[1] : "reproducable_test_log.csv";
    const double t_end = argc > 2 ? atof(argv[2]) : 4.0;
    const int num_worlds = argc > 3 ? atoi(argv[3]) : 1;
    const bool fast_step = argc > 4 ? atoi(argv[4]) != 0 : true; // 0: dWorldStep (the reference's default), 1: dWorldQuickStep

    // the test registers "sphere" as Box, like the reference test does
    ObjectFactory::Instance().addObjectType("box", &Box::instantiate);
    ObjectFactory::Instance().addObjectType("sphere", &Box::instantiate);
    ObjectFactory::Instance().addObjectType("inertial", &Inertial::instantiate);
    JointFactory::Instance().addJointType("hinge", &HingeJoint::instantiate);
    JointFactory::Instance().addJointType("fixed", &FixedJoint::instantiate);
    JointFactory::Instance().addJointType("slider", &SliderJoint::instantiate);
    JointFactory::Instance().addJointType("prismatic", &SliderJoint::instantiate);

    WorldPhysics wp;
    wp.setNumWorlds(num_worlds);
    try {
        wp.initTheWorld();
    } catch (const std::exception &e) {
        fprintf(stderr, "reproducable_test: %s\n", e.what());
        return 2;
    }
    wp.step_size = 0.001;
    wp.fast_step = fast_step;

    ConfigMap map;
    const LinkSpec links[9] = {{"body", 0.4, 1.0},      {"fl_upper", 0.03, 0.2}, {"fl_lower", 0.03, 0.2},
                               {"fr_upper", 0.03, 0.2}, {"fr_lower", 0.03, 0.2}, {"rl_upper", 0.03, 0.2},
                               {"rl_lower", 0.03, 0.2}, {"rr_upper", 0.03, 0.2}, {"rr_lower", 0.03, 0.2}};
    std::vector<std::shared_ptr<DynamicObject>> bodies, feet;
    for (const LinkSpec &s : links) bodies.push_back(makeLink(wp, map, s));
    std::shared_ptr<DynamicObject> torso = bodies[0];

    const char *legs[4] = {"fl", "fr", "rl", "rr"};
    const double sx[4] = {1, 1, -1, -1}, sy[4] = {1, -1, 1, -1};
    map["type"] << std::string("sphere");
    map["mass"] = 0.05;
    map["extend"]["x"] = 0.02;
    map["position"]["z"] = 0.02;
    for (int i = 0; i < 4; i++) {
        map["position"]["x"] = sx[i] * (0.2 - 0.0015);
        map["position"]["y"] = sy[i] * (0.1 - 0.0015);
        map["name"] << std::string(legs[i]) + "_foot";
        map["parentFrame"] << std::string(legs[i]) + "_foot";
        feet.push_back(wp.createFrame(nullptr, map));
        wp.createObject(map);
    }

    std::vector<std::shared_ptr<JointInterface>> joints;
    ConfigMap jm;
    jm["axis1"]["x"] = 0.0;
    jm["axis1"]["y"] = 1.0;
    jm["axis1"]["z"] = 0.0;
    jm["reducedDataPackage"] = true;
    auto addJoint = [&](const std::string &name, const char *type, const std::string &parent, const std::string &child,
                        int leg, double z) {
        jm["name"] << name;
        jm["type"] << std::string(type);
        jm["parent_link_name"] << parent;
        jm["child_link_name"] << child;
        jm["anchor"]["x"] = sx[leg] * (0.2 - 0.015);
        jm["anchor"]["y"] = sy[leg] * (0.1 - 0.015);
        jm["anchor"]["z"] = z;
        joints.push_back(wp.createJoint(nullptr, jm));
    };
    for (int i = 0; i < 4; i++) addJoint(std::string(legs[i]) + "_hip", "hinge", "body", std::string(legs[i]) + "_upper", i, 0.44);
    for (int i = 0; i < 4; i++)
        addJoint(std::string(legs[i]) + "_knee", "hinge", std::string(legs[i]) + "_upper", std::string(legs[i]) + "_lower", i, 0.24);
    for (int i = 0; i < 4; i++)
        addJoint(std::string(legs[i]) + "_ankle", "fixed", std::string(legs[i]) + "_lower", std::string(legs[i]) + "_foot", i, 0.4);
    for (auto &j : joints) {
        if (!j) {
            fprintf(stderr, "reproducable_test: joint creation failed\n");
            return 3;
        }
        j->setForceLimit(9);
        j->setJointAsMotor(1);
    }

    FILE *file = fopen(csv.c_str(), "w");
    if (!file) return 4;
    fprintf(file, "time_sec,body_x,body_y,body_z\n");
    const double offset[4] = {0.0, 1.14, 0.0, 1.14};
    ContactData cd;
    cd.c_params.cfm = 0.00001;
    cd.c_params.erp = 0.2;
    cd.c_params.friction1 = 0.8;
    cd.c_params.friction2 = 0.8;
    double time = 0.0;
    Vector pos;
    while (time < t_end) {
        torso->getPosition(&pos);
        fprintf(file, "%g,%g,%g,%g\n", time, pos.x(), pos.y(), pos.z());
        for (int i = 0; i < 4; i++) { // P controller (gain 20) on a sine gait
            const double hip = 0.5 * sin(offset[i] + time * 4 * M_PI) - 0.3;
            const double knee = -0.5 * sin(offset[i] + 1 + time * 4 * M_PI) + 0.3;
            joints[i]->setVelocity((hip - joints[i]->getPosition()) * 20);
            joints[i + 4]->setVelocity((knee - joints[i + 4]->getPosition()) * 20);
        }
        wp.clearPreviousStep();
        for (auto &foot : feet) {
            foot->getPosition(&pos);
            pos.z() -= 0.02;
            if (pos.z() <= 0.0) {
                cd.pos = pos;
                cd.pos.z() = 0.0;
                cd.depth = pos.z();
                cd.normal = Vector(0.0, 0.0, 1.0);
                cd.body1 = foot;
                cd.body2 = nullptr;
                cd.nameObject1 = foot->getName();
                foot->addContact(cd);
            }
        }
        wp.stepTheWorld();
        time += wp.step_size;
    }
    fclose(file);
    joints.clear();
    for (auto &b : bodies) wp.removeFrame(b->getName());
    for (auto &f : feet) wp.removeFrame(f->getName());
    bodies.clear();
    feet.clear();
    torso = nullptr;
    wp.freeTheWorld();
    return 0;
}

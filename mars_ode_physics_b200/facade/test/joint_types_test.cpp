// joint_types_test.cpp -- the joint types beyond the reference's registered set (universal, ball), created the
// way a MARS host would: through the plugin loader (WorldPhysicsLoader::createWorldInstance, reference
// src/WorldPhysicsLoader.cpp:36-40) and PhysicsInterface::createFrame / createObject / createJoint with
// ConfigMaps (reference src/WorldPhysics.cpp:515-655).  A two-link arm hangs from a carrier frame on a ball
// joint, its second link on a universal joint; the program steps it and prints what a caller can observe.
//
//   joint_types_test [steps]   ->  "ball_err rigid_err_upper rigid_err_lower angle1 moved" on stdout, exit code 0 if sane
#include <cmath>
#include <cstdio>
#include <cstdlib>
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

static std::shared_ptr<DynamicObject> link(PhysicsInterface &wp, const char *name, double x, double mass)
{
    ConfigMap map;
    map["name"] << std::string(name);
    map["parentFrame"] << std::string(name);
    map["type"] = "box";
    map["extend"]["x"] = 0.3;
    map["extend"]["y"] = 0.05;
    map["extend"]["z"] = 0.05;
    map["position"]["x"] = x;
    map["position"]["y"] = 0.0;
    map["position"]["z"] = 1.0;
    map["orientation"]["x"] = 0.0;
    map["orientation"]["y"] = 0.0;
    map["orientation"]["z"] = 0.0;
    map["orientation"]["w"] = 1.0;
    map["density"] = 1.0;
    map["mass"] = mass;
    auto frame = wp.createFrame(nullptr, map);
    wp.createObject(map);
    return frame;
}

int main(int argc, char **argv)
{
    const int steps = argc > 1 ? atoi(argv[1]) : 400;
    WorldPhysicsLoader loader(nullptr); // registers box/sphere/... and hinge/fixed/slider/hinge2/universal/ball
    std::shared_ptr<PhysicsInterface> wp = loader.createWorldInstance();
    try {
        wp->initTheWorld();
    } catch (const std::exception &e) {
        fprintf(stderr, "joint_types_test: %s\n", e.what());
        return 2;
    }
    wp->step_size = 0.001;
    wp->fast_step = true;

    auto carrier = link(*wp, "carrier", -0.15, 50.0);
    auto a = link(*wp, "upper", 0.15, 0.8);
    auto b = link(*wp, "lower", 0.45, 0.5);

    ConfigMap jm;
    jm["reducedDataPackage"] = true;
    jm["name"] << std::string("hold");
    jm["type"] << std::string("fixed");
    jm["parent_link_name"] << std::string("world");
    jm["child_link_name"] << std::string("carrier");
    auto hold = wp->createJoint(nullptr, jm);

    jm["name"] << std::string("shoulder");
    jm["type"] << std::string("ball");
    jm["parent_link_name"] << std::string("carrier");
    jm["child_link_name"] << std::string("upper");
    jm["anchor"]["x"] = 0.0;
    jm["anchor"]["y"] = 0.0;
    jm["anchor"]["z"] = 1.0;
    auto ball = wp->createJoint(nullptr, jm);

    jm["name"] << std::string("elbow");
    jm["type"] << std::string("universal");
    jm["parent_link_name"] << std::string("upper");
    jm["child_link_name"] << std::string("lower");
    jm["anchor"]["x"] = 0.3;
    jm["axis1"]["x"] = 0.0;
    jm["axis1"]["y"] = 1.0;
    jm["axis1"]["z"] = 0.0;
    jm["axis2"]["x"] = 0.0;
    jm["axis2"]["y"] = 0.0;
    jm["axis2"]["z"] = 1.0;
    auto uni = wp->createJoint(nullptr, jm);
    if (!ball || !uni) {
        fprintf(stderr, "joint_types_test: joint creation failed (hold %d ball %d universal %d)\n", (int)!!hold, (int)!!ball,
                (int)!!uni);
        return 3;
    }
    if (uni->getType() != JOINT_TYPE_UNIVERSAL || ball->getType() != JOINT_TYPE_BALL) return 4;

    // Frames start at the origin and objects are placed relative to them (reference Frame.cpp:183-207,794-851), so
    // a frame position is a body-fixed point: its distance to the joint anchor on the same body must not change.
    auto dist = [](const Vector &p, const Vector &q) {
        return std::sqrt((p.x() - q.x()) * (p.x() - q.x()) + (p.y() - q.y()) * (p.y() - q.y()) + (p.z() - q.z()) * (p.z() - q.z()));
    };
    Vector pa, pb, pb0, anchor, an2;
    a->getPosition(&pa);
    b->getPosition(&pb0);
    ball->getAnchor(&anchor);
    uni->getAnchor(&an2);
    const double len_a0 = dist(pa, anchor), len_b0 = dist(pb0, an2);
    for (int i = 0; i < steps; i++) {
        wp->clearPreviousStep();
        wp->stepTheWorld();
    }
    a->getPosition(&pa);
    b->getPosition(&pb);
    ball->getAnchor(&anchor);
    uni->getAnchor(&an2);
    // the ball anchor stays where the (heavy, held) carrier is
    const double ball_err = dist(anchor, Vector(0.0, 0.0, 1.0));
    const double ea = std::fabs(dist(pa, anchor) - len_a0), eb = std::fabs(dist(pb, an2) - len_b0), moved = dist(pb, pb0);
    printf("%g %g %g %g %g\n", ball_err, ea, eb, (double)uni->getPosition(), moved);
    const bool sane = ball_err < 5e-3 && ea < 5e-3 && eb < 5e-3 && moved > 0.05 && std::isfinite(uni->getPosition());
    return sane ? 0 : 5;
}

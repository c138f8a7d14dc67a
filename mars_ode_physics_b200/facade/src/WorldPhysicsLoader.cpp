// WorldPhysicsLoader.cpp -- registers the object / joint types and hands out world instances
// (reference src/WorldPhysicsLoader.cpp:22-46).  Beyond the reference's registrations (box, inertial;
// hinge, fixed, slider, prismatic) the types that exist in its tree but are not wired up are
// registered too: sphere, capsule, cylinder (CMakeLists.txt:75-90 leaves them out) and hinge2
// (compiled but never registered), because the north-star configurations need them.
#include "WorldPhysicsLoader.hpp"

#include "Joints.hpp"
#include "Objects.hpp"

namespace mars {
namespace ode_physics {

WorldPhysicsLoader::WorldPhysicsLoader(lib_manager::LibManager *theManager) : lib_manager::LibInterface(theManager)
{
    ObjectFactory::Instance().addObjectType("box", &Box::instantiate);
    ObjectFactory::Instance().addObjectType("inertial", &Inertial::instantiate);
    ObjectFactory::Instance().addObjectType("sphere", &Sphere::instantiate);
    ObjectFactory::Instance().addObjectType("capsule", &Capsule::instantiate);
    ObjectFactory::Instance().addObjectType("cylinder", &Cylinder::instantiate);
    JointFactory::Instance().addJointType("hinge", &HingeJoint::instantiate);
    JointFactory::Instance().addJointType("fixed", &FixedJoint::instantiate);
    JointFactory::Instance().addJointType("slider", &SliderJoint::instantiate);
    JointFactory::Instance().addJointType("prismatic", &SliderJoint::instantiate);
    JointFactory::Instance().addJointType("hinge2", &Hinge2Joint::instantiate);
    JointFactory::Instance().addJointType("universal", &UniversalJoint::instantiate);
    JointFactory::Instance().addJointType("ball", &BallJoint::instantiate);
}

WorldPhysicsLoader::~WorldPhysicsLoader(void) {}

std::shared_ptr<interfaces::PhysicsInterface> WorldPhysicsLoader::createWorldInstance()
{
    std::shared_ptr<WorldPhysics> worldPhysics = std::make_shared<WorldPhysics>();
    worldPhysics->setNumWorlds(num_worlds);
    return std::static_pointer_cast<interfaces::PhysicsInterface>(worldPhysics);
}

} // namespace ode_physics
} // namespace mars

DESTROY_LIB(mars::ode_physics::WorldPhysicsLoader)
CREATE_LIB(mars::ode_physics::WorldPhysicsLoader)

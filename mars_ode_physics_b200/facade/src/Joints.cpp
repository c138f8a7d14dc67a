// Joints.cpp -- see Joints.hpp.  Each function cites the reference lines whose behaviour it keeps.
#include "Joints.hpp"

#include <cmath>
#include <cstdio>
#include <iostream>
#include <limits>

namespace mars {
namespace ode_physics {

// hinge2 and universal share every two-axis arm of the reference's Joint.cpp (e.g. :302,330,349,370,389,410)
static inline bool twoAxes(int t)
{
    return t == interfaces::JOINT_TYPE_HINGE2 || t == interfaces::JOINT_TYPE_UNIVERSAL;
}


using namespace utils;
using namespace interfaces;

static const double kInf = std::numeric_limits<double>::infinity();

Joint::Joint(WorldPhysics *world, data_broker::DataBrokerInterface *, configmaps::ConfigMap &config)
    : theWorld(world), config(config)
{
    jointId = -1;
    jointCFM = 0.0;
    cfm = cfm1 = cfm2 = erp1 = erp2 = 0;
    lo1 = lo2 = hi1 = hi2 = 0;
    damping = spring = 0;
    frictionCoefficient = -0.1; // Joint.cpp:43: off by default
    body1 = body2 = -1;
    joint_type = JOINT_TYPE_UNDEFINED;
    joint_created = false;
    motor_torque = 0;
    world_index = 0;
    fetched_step = -1;
    for (double &v : feedback) v = 0;
    name << config["name"];
    if (config.hasKey("frictionCoefficient")) frictionCoefficient = config["frictionCoefficient"];
}

Joint::~Joint(void) {} // joints live as long as the arena

void Joint::getName(std::string *n) const { *n = this->name; }
JointType Joint::getType() const { return static_cast<JointType>(joint_type); }
bool Joint::isJointCreated() { return joint_created; }

void Joint::setParam(int param, double value)
{
    b2p_joint_set_param(theWorld->getWorld(), jointId, B2P_ALL_WORLDS, param, value);
}
double Joint::getParam(int param) const
{
    double v = 0;
    b2p_joint_get_param(theWorld->getWorld(), jointId, world_index, param, &v);
    return v;
}

// Joint.cpp:147-205 -- spring/damper constants to stop ERP/CFM; the reference then forces lo1 = hi1 = 0
void Joint::calculateCfmErp()
{
    damping = -1.0;
    spring = -1.0;
    if (config.hasKey("springDamping")) damping = config["springDamping"];
    if (config.hasKey("damping_const_constraint_axis1")) damping = config["damping_const_constraint_axis1"];
    if (config.hasKey("springStiffness")) spring = config["springStiffness"];
    if (config.hasKey("spring_const_constraint_axis1")) spring = config["spring_const_constraint_axis1"];
    const double h = theWorld->getWorldStep();
    cfm = damping;
    erp1 = h * spring + damping;
    cfm1 = h * spring + damping;
    lo1 = 0.0;
    hi1 = 0.0;
    cfm = (cfm > 0) ? 1 / cfm : 0.000000000001;
    erp1 = (erp1 > 0) ? h * spring / erp1 : 0;
    cfm1 = (cfm1 > 0) ? 1 / cfm1 : 0.00000000001;
}

// Joint.cpp:216-273
bool Joint::createJoint()
{
    MutexLocker locker(&(theWorld->iMutex));
    if (!(theWorld && theWorld->existsWorld())) return false;
    parentFrame = theWorld->getFrameIntern(config["parent_link_name"].toString());
    childFrame = theWorld->getFrameIntern(config["child_link_name"].toString());
    std::shared_ptr<Frame> n1 = parentFrame.lock(), n2 = childFrame.lock();
    if (n1 == nullptr && n2 == nullptr) {
        fprintf(stderr, "Can not create a new joint, since no parent nor a child link could be found. Parent name: %s \tChild name: %s\n",
                config["parent_link_name"].toString().c_str(), config["child_link_name"].toString().c_str());
        return false;
    }
    calculateCfmErp();
    if (config.hasKey("jointCFM")) jointCFM = config["jointCFM"];
    joint_type = getJointType(config["type"].getString());
    body1 = n1 ? n1->acquireBody() : -1;
    body2 = n2 ? n2->acquireBody() : -1;
    if (!createJoint(body1, body2)) return false;
    // Joint::applyJointFriction runs in the step's device epilogue with this coefficient (off unless > 0)
    b2p_joint_set_friction_coefficient(theWorld->getWorld(), jointId, frictionCoefficient);
    theWorld->noteJoint(jointId, joint_type);
    if (n1 && n2) {
        n1->addLinkedFrame(n2);
        n2->addLinkedFrame(n1);
    }
    joint_created = true;
    return true;
}

bool Joint::createJoint(int, int)
{
    std::cout << "Joint: using default createJoint func. Did you forget to override it?." << std::endl;
    return false;
}

// damping-as-motor / spring-as-stop / jointCFM block shared by hinge, slider and hinge2
// (HingeJoint.cpp:40-63, SliderJoint.cpp:39-62, Hinge2Joint.cpp:46-74)
void Joint::applyLimotConfig(bool two_axes)
{
    if (damping > 0.00000001) {
        setParam(B2P_PARAM_FMAX, damping);
        setParam(B2P_PARAM_VEL, 0);
        if (two_axes) {
            setParam(B2P_PARAM_FMAX | B2P_PARAM_GROUP2, damping);
            setParam(B2P_PARAM_VEL | B2P_PARAM_GROUP2, 0);
        }
    }
    if (spring > 0.00000001) {
        if (two_axes) { // the reference writes the axis-2 values over the axis-1 slots (Hinge2Joint.cpp:55-62)
            setParam(B2P_PARAM_LO_STOP, lo2);
            setParam(B2P_PARAM_HI_STOP, hi2);
            setParam(B2P_PARAM_STOP_CFM, cfm2);
            setParam(B2P_PARAM_STOP_ERP, erp2);
        } else {
            setParam(B2P_PARAM_LO_STOP, lo1);
            setParam(B2P_PARAM_HI_STOP, hi1);
            setParam(B2P_PARAM_STOP_CFM, cfm1);
            setParam(B2P_PARAM_STOP_ERP, erp1);
        }
    } else if (!two_axes && lo1 != 0) {
        setParam(B2P_PARAM_LO_STOP, lo1);
        setParam(B2P_PARAM_HI_STOP, hi1);
    }
    if (jointCFM > 0) setParam(B2P_PARAM_CFM, jointCFM);
}

void Joint::getAnchor(Vector *anchor) const
{
    MutexLocker locker(&(theWorld->iMutex));
    double p[3] = {0, 0, 0};
    if (joint_type == JOINT_TYPE_HINGE || twoAxes(joint_type) || joint_type == JOINT_TYPE_BALL)
        b2p_joint_get_anchor(theWorld->getWorld(), jointId, world_index, p);
    *anchor = Vector(p[0], p[1], p[2]);
}

void Joint::setAnchor(const Vector &a)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (joint_type == JOINT_TYPE_HINGE || twoAxes(joint_type) || joint_type == JOINT_TYPE_BALL)
        b2p_joint_set_anchor(theWorld->getWorld(), jointId, a.x(), a.y(), a.z());
}

void Joint::setAxis(const Vector &a)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (joint_type == JOINT_TYPE_HINGE || twoAxes(joint_type) || joint_type == JOINT_TYPE_SLIDER)
        b2p_joint_set_axis1(theWorld->getWorld(), jointId, a.x(), a.y(), a.z());
}

void Joint::setAxis2(const Vector &a)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (twoAxes(joint_type)) b2p_joint_set_axis2(theWorld->getWorld(), jointId, a.x(), a.y(), a.z());
}

void Joint::getAxisI(Vector *axis) const
{
    double p[3] = {0, 0, 0};
    if (joint_type == JOINT_TYPE_HINGE || twoAxes(joint_type) || joint_type == JOINT_TYPE_SLIDER)
        b2p_joint_get_axis1(theWorld->getWorld(), jointId, world_index, p);
    *axis = Vector(p[0], p[1], p[2]);
}
void Joint::getAxis(Vector *axis) const
{
    MutexLocker locker(&(theWorld->iMutex));
    getAxisI(axis);
}
void Joint::getAxis2(Vector *axis) const
{
    MutexLocker locker(&(theWorld->iMutex));
    double p[3] = {0, 0, 0};
    if (twoAxes(joint_type)) b2p_joint_get_axis2(theWorld->getWorld(), jointId, world_index, p);
    *axis = Vector(p[0], p[1], p[2]);
}

void Joint::setWorldObject(std::shared_ptr<PhysicsInterface>) {}

// Joint.cpp:315-353
void Joint::setForceLimit(sReal max_force)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (joint_type == JOINT_TYPE_HINGE || twoAxes(joint_type) || joint_type == JOINT_TYPE_SLIDER)
        setParam(B2P_PARAM_FMAX, max_force);
}
void Joint::setForceLimit2(sReal max_force)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (twoAxes(joint_type)) setParam(B2P_PARAM_FMAX | B2P_PARAM_GROUP2, max_force);
}
// Joint.cpp:355-393: hinge and slider negate, hinge2 does not
void Joint::setVelocity(sReal velocity)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (joint_type == JOINT_TYPE_HINGE || joint_type == JOINT_TYPE_SLIDER) setParam(B2P_PARAM_VEL, -velocity);
    else if (twoAxes(joint_type)) setParam(B2P_PARAM_VEL, velocity);
}
void Joint::setVelocity2(sReal velocity)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (twoAxes(joint_type)) setParam(B2P_PARAM_VEL | B2P_PARAM_GROUP2, velocity);
}
void Joint::setVelocityBatch(const float *v)
{
    MutexLocker locker(&(theWorld->iMutex));
    std::vector<float> tmp;
    if (joint_type == JOINT_TYPE_HINGE || joint_type == JOINT_TYPE_SLIDER) {
        tmp.assign(v, v + theWorld->getNumWorlds());
        for (float &x : tmp) x = -x;
        v = tmp.data();
    }
    checkB2P(b2p_joint_set_param_batch(theWorld->getWorld(), jointId, B2P_PARAM_VEL, v), "Joint::setVelocityBatch");
    if (!tmp.empty()) b2p_sync(theWorld->getWorld()); // tmp must outlive the asynchronous copy
}

// Joint.cpp:395-429
sReal Joint::getPosition(void) const
{
    MutexLocker locker(&(theWorld->iMutex));
    double a = 0;
    b2p_joint_get_angle1(theWorld->getWorld(), jointId, world_index, &a);
    if (joint_type == JOINT_TYPE_HINGE || joint_type == JOINT_TYPE_SLIDER) return -a;
    if (twoAxes(joint_type)) return a;
    return 0;
}
sReal Joint::getPosition2(void) const { return 0; } // only universal in the reference (Joint.cpp:417-429)

// Joint.cpp:618-703
void Joint::setJointAsMotor(int axis)
{
    MutexLocker locker(&(theWorld->iMutex));
    switch (joint_type) {
    case JOINT_TYPE_HINGE:
        setParam(B2P_PARAM_LO_STOP, -kInf);
        setParam(B2P_PARAM_HI_STOP, kInf);
        break;
    case JOINT_TYPE_HINGE2:
    case JOINT_TYPE_UNIVERSAL:
        if (!lo1 && !hi1 && axis == 1) {
            setParam(B2P_PARAM_LO_STOP, -kInf);
            setParam(B2P_PARAM_HI_STOP, kInf);
        }
        if (!lo2 && !hi2 && axis == 2) {
            setParam(B2P_PARAM_LO_STOP | B2P_PARAM_GROUP2, -kInf);
            setParam(B2P_PARAM_HI_STOP | B2P_PARAM_GROUP2, kInf);
        }
        break;
    case JOINT_TYPE_SLIDER:
        if (!lo1 && !hi1) {
            setParam(B2P_PARAM_LO_STOP, -kInf);
            setParam(B2P_PARAM_HI_STOP, kInf);
        }
        break;
    default: break;
    }
}
void Joint::unsetJointAsMotor(int axis)
{
    MutexLocker locker(&(theWorld->iMutex));
    switch (joint_type) {
    case JOINT_TYPE_HINGE:
    case JOINT_TYPE_SLIDER:
        setParam(B2P_PARAM_LO_STOP, lo1);
        setParam(B2P_PARAM_HI_STOP, hi1);
        break;
    case JOINT_TYPE_HINGE2:
    case JOINT_TYPE_UNIVERSAL:
        if (axis == 1) {
            setParam(B2P_PARAM_LO_STOP, lo1);
            setParam(B2P_PARAM_HI_STOP, hi1);
        } else if (axis == 2) {
            setParam(B2P_PARAM_LO_STOP | B2P_PARAM_GROUP2, lo2);
            setParam(B2P_PARAM_HI_STOP | B2P_PARAM_GROUP2, hi2);
        }
        break;
    default: break;
    }
}

void Joint::getForce1(Vector *f) const
{
    fetchPostStep();
    *f = Vector(feedback[0], feedback[1], feedback[2]);
}
void Joint::getTorque1(Vector *t) const
{
    fetchPostStep();
    *t = Vector(feedback[3], feedback[4], feedback[5]);
}
void Joint::getForce2(Vector *f) const
{
    fetchPostStep();
    *f = Vector(feedback[6], feedback[7], feedback[8]);
}
void Joint::getTorque2(Vector *t) const
{
    fetchPostStep();
    *t = Vector(feedback[9], feedback[10], feedback[11]);
}

// Joint.cpp:1121-1146: hinge torque / slider force, sign flipped
void Joint::setTorque(sReal torque)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (joint_type == JOINT_TYPE_HINGE || joint_type == JOINT_TYPE_SLIDER)
        b2p_joint_add_torque(theWorld->getWorld(), jointId, world_index, -1 * torque);
}
void Joint::setTorque2(sReal) {}

// Joint.cpp:782-825
void Joint::reattacheJoint(void)
{
    MutexLocker locker(&(theWorld->iMutex));
    b2p_arena *a = theWorld->getWorld();
    if (joint_type == JOINT_TYPE_HINGE || twoAxes(joint_type)) {
        double p[3];
        b2p_joint_get_anchor(a, jointId, 0, p);
        b2p_joint_set_anchor(a, jointId, p[0], p[1], p[2]);
    } else if (joint_type == JOINT_TYPE_FIXED) {
        b2p_joint_set_fixed(a, jointId);
    }
}

// Joint.cpp:859-1019 (motor torque = -feedback.lambda; axis torque and joint load from the feedback forces and the
// anchor geometry) and :1031-1063 (joint friction for the next tick) run in the step's device epilogue for every
// world (b2p_integrate.cu, restated in oracle/orc_post.c).  update() only invalidates the values cached for the
// active world; the getters fetch them -- two small columns -- when they are asked for.
void Joint::update(void) { fetched_step = -1; }

void Joint::fetchPostStep() const
{
    if (jointId < 0 || fetched_step == theWorld->getStepCount()) return;
    b2p_arena *a = theWorld->getWorld();
    double u[7];
    b2p_joint_get_feedback(a, jointId, world_index, feedback);
    b2p_joint_get_update(a, jointId, world_index, u);
    motor_torque = u[0];
    axis1_torque = Vector(u[1], u[2], u[3]);
    axis2_torque.setZero(); // calc2 is always 0 in the reference (:880,893)
    joint_load = Vector(u[4], u[5], u[6]);
    fetched_step = theWorld->getStepCount();
}

void Joint::getAxisTorque(Vector *t) const
{
    fetchPostStep();
    *t = axis1_torque;
}
void Joint::getAxis2Torque(Vector *t) const
{
    fetchPostStep();
    *t = axis2_torque;
}
void Joint::getJointLoad(Vector *t) const
{
    fetchPostStep();
    *t = joint_load;
}
sReal Joint::getMotorTorque(void) const
{
    fetchPostStep();
    return motor_torque;
}

void Joint::setJointFrictionCoefficient(sReal friction)
{
    frictionCoefficient = friction;
    if (jointId >= 0) b2p_joint_set_friction_coefficient(theWorld->getWorld(), jointId, friction);
}
sReal Joint::getJointFrictionCoefficient() { return frictionCoefficient; }

sReal Joint::getVelocity(void) const
{
    MutexLocker locker(&(theWorld->iMutex));
    return getVelocityI();
}
// Joint.cpp:1078-1102
sReal Joint::getVelocityI(void) const
{
    double r = 0;
    b2p_joint_get_angle1_rate(theWorld->getWorld(), jointId, world_index, &r);
    if (joint_type == JOINT_TYPE_HINGE || joint_type == JOINT_TYPE_SLIDER) return -1 * r;
    if (twoAxes(joint_type)) return r;
    return 0;
}
sReal Joint::getVelocity2(void) const
{
    MutexLocker locker(&(theWorld->iMutex));
    double r = 0;
    if (twoAxes(joint_type)) b2p_joint_get_angle2_rate(theWorld->getWorld(), jointId, world_index, &r);
    return r;
}

// Joint.cpp:1176-1294
void Joint::changeStepSize(const JointData &jointS)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (!(theWorld && theWorld->existsWorld())) return;
    calculateCfmErp();
    if (jointS.type == JOINT_TYPE_HINGE || jointS.type == JOINT_TYPE_SLIDER) applyLimotConfig(false);
    else if (twoAxes(jointS.type)) applyLimotConfig(true);
}

// Joint.cpp:1301-1464 (including the reference's sign / slot conventions)
sReal Joint::getLowStop() const
{
    if (joint_type == JOINT_TYPE_HINGE || joint_type == JOINT_TYPE_SLIDER) return -1 * getParam(B2P_PARAM_HI_STOP);
    if (twoAxes(joint_type)) return getParam(B2P_PARAM_LO_STOP);
    fprintf(stderr, "mars::ode_physics::Joint: getLowStop for type %d not implemented yet\n", joint_type);
    return 0.;
}
sReal Joint::getHighStop() const
{
    if (joint_type == JOINT_TYPE_HINGE || joint_type == JOINT_TYPE_SLIDER) return -1 * getParam(B2P_PARAM_LO_STOP);
    if (twoAxes(joint_type)) return getParam(B2P_PARAM_HI_STOP);
    fprintf(stderr, "mars::ode_physics::Joint: getHighStop for type %d not implemented yet\n", joint_type);
    return 0.;
}
sReal Joint::getLowStop2() const
{
    if (twoAxes(joint_type)) return getParam(B2P_PARAM_LO_STOP | B2P_PARAM_GROUP2);
    fprintf(stderr, "mars::ode_physics::Joint: getLowStop2 for type %d not implemented yet\n", joint_type);
    return 0.;
}
sReal Joint::getHighStop2() const
{
    if (twoAxes(joint_type)) return getParam(B2P_PARAM_HI_STOP | B2P_PARAM_GROUP2);
    fprintf(stderr, "mars::ode_physics::Joint: getHighStop2 for type %d not implemented yet\n", joint_type);
    return 0.;
}
sReal Joint::getCFM() const
{
    if (joint_type == JOINT_TYPE_HINGE || twoAxes(joint_type) || joint_type == JOINT_TYPE_SLIDER)
        return getParam(B2P_PARAM_CFM);
    fprintf(stderr, "mars::ode_physics::Joint: cfm for type %d not implemented yet\n", joint_type);
    return 0.;
}
void Joint::setLowStop(sReal lowStop)
{
    if (joint_type == JOINT_TYPE_HINGE || twoAxes(joint_type) || joint_type == JOINT_TYPE_SLIDER)
        setParam(B2P_PARAM_LO_STOP, lowStop);
    else fprintf(stderr, "mars::ode_physics::Joint: setLowStop for type %d not implemented yet\n", joint_type);
}
void Joint::setHighStop(sReal highStop)
{
    // the reference writes dParamLoStop with the negated value for hinge and slider (Joint.cpp:1403,1409)
    if (joint_type == JOINT_TYPE_HINGE || joint_type == JOINT_TYPE_SLIDER) setParam(B2P_PARAM_LO_STOP, -1 * highStop);
    else if (twoAxes(joint_type)) setParam(B2P_PARAM_HI_STOP, highStop);
    else fprintf(stderr, "mars::ode_physics::Joint: setHighStop for type %d not implemented yet\n", joint_type);
}
void Joint::setLowStop2(sReal lowStop)
{
    if (twoAxes(joint_type)) setParam(B2P_PARAM_HI_STOP, -1 * lowStop); // sic (Joint.cpp:1423)
    else fprintf(stderr, "mars::ode_physics::Joint: setLowStop2 for type %d not implemented yet\n", joint_type);
}
void Joint::setHighStop2(sReal highStop)
{
    if (twoAxes(joint_type)) setParam(B2P_PARAM_LO_STOP | B2P_PARAM_GROUP2, -1 * highStop);
    else fprintf(stderr, "mars::ode_physics::Joint: setHighStop2 for type %d not implemented yet\n", joint_type);
}
void Joint::setCFM(sReal v)
{
    if (joint_type == JOINT_TYPE_HINGE || twoAxes(joint_type) || joint_type == JOINT_TYPE_SLIDER)
        setParam(B2P_PARAM_CFM, v);
    else fprintf(stderr, "mars::ode_physics::Joint: cfm for type %d not implemented yet\n", joint_type);
}
void Joint::getDataBrokerNames(std::string &groupName, std::string &dataName)
{
    groupName = "mars_sim";
    dataName = "Joints/" + name;
}

// ---------------------------------------------------------------------------------------------
#define B2P_JOINT_BOILERPLATE(Name)                                                                        \
    Name::Name(WorldPhysics *world, data_broker::DataBrokerInterface *db, configmaps::ConfigMap &config)   \
        : Joint(world, db, config)                                                                         \
    {                                                                                                      \
        Joint::createJoint();                                                                              \
    }                                                                                                      \
    Joint *Name::instantiate(WorldPhysics *world, data_broker::DataBrokerInterface *db, configmaps::ConfigMap &config) \
    {                                                                                                      \
        return new Name(world, db, config);                                                                \
    }
B2P_JOINT_BOILERPLATE(HingeJoint)
B2P_JOINT_BOILERPLATE(FixedJoint)
B2P_JOINT_BOILERPLATE(SliderJoint)
B2P_JOINT_BOILERPLATE(Hinge2Joint)
B2P_JOINT_BOILERPLATE(UniversalJoint)
B2P_JOINT_BOILERPLATE(BallJoint)
#undef B2P_JOINT_BOILERPLATE

// HingeJoint.cpp:31-66
bool HingeJoint::createJoint(int b1, int b2)
{
    b2p_arena *a = theWorld->getWorld();
    jointId = b2p_joint_create(a, B2P_JOINT_HINGE, b1, b2);
    if (jointId < 0) return false;
    b2p_joint_set_anchor(a, jointId, config["anchor"]["x"], config["anchor"]["y"], config["anchor"]["z"]);
    b2p_joint_set_axis1(a, jointId, config["axis1"]["x"], config["axis1"]["y"], config["axis1"]["z"]);
    applyLimotConfig(false);
    return true;
}

// FixedJoint.cpp:31-47
bool FixedJoint::createJoint(int b1, int b2)
{
    if (b1 < 0 && b2 < 0) return false;
    b2p_arena *a = theWorld->getWorld();
    jointId = b2p_joint_create(a, B2P_JOINT_FIXED, b1, b2);
    if (jointId < 0) return false;
    b2p_joint_set_fixed(a, jointId);
    return true;
}

// SliderJoint.cpp:32-65
bool SliderJoint::createJoint(int b1, int b2)
{
    b2p_arena *a = theWorld->getWorld();
    jointId = b2p_joint_create(a, B2P_JOINT_SLIDER, b1, b2);
    if (jointId < 0) return false;
    b2p_joint_set_axis1(a, jointId, config["axis1"]["x"], config["axis1"]["y"], config["axis1"]["z"]);
    applyLimotConfig(false);
    return true;
}

// Hinge2Joint.cpp:33-77
bool Hinge2Joint::createJoint(int b1, int b2)
{
    if (b1 < 0 && b2 < 0) return false;
    b2p_arena *a = theWorld->getWorld();
    jointId = b2p_joint_create(a, B2P_JOINT_HINGE2, b1, b2);
    if (jointId < 0) {
        fprintf(stderr, "Hinge2Joint: %s\n", b2p_last_error());
        return false;
    }
    b2p_joint_set_anchor(a, jointId, config["anchor"]["x"], config["anchor"]["y"], config["anchor"]["z"]);
    b2p_joint_set_axis1(a, jointId, config["axis1"]["x"], config["axis1"]["y"], config["axis1"]["z"]);
    b2p_joint_set_axis2(a, jointId, config["axis2"]["x"], config["axis2"]["y"], config["axis2"]["z"]);
    applyLimotConfig(true);
    return true;
}

// UniversalJoint.cpp:33-80.  The reference creates a *hinge2* there and then sets *universal* parameters on it
// (self-inconsistent, and the type is not registered in its loader); this creates the universal joint those
// parameter calls address: same config keys (anchor, axis1, axis2, damping / spring / stops for both axes).
bool UniversalJoint::createJoint(int b1, int b2)
{
    if (b1 < 0 && b2 < 0) return false;
    b2p_arena *a = theWorld->getWorld();
    jointId = b2p_joint_create(a, B2P_JOINT_UNIVERSAL, b1, b2);
    if (jointId < 0) {
        fprintf(stderr, "UniversalJoint: %s\n", b2p_last_error());
        return false;
    }
    b2p_joint_set_anchor(a, jointId, config["anchor"]["x"], config["anchor"]["y"], config["anchor"]["z"]);
    b2p_joint_set_axis1(a, jointId, config["axis1"]["x"], config["axis1"]["y"], config["axis1"]["z"]);
    b2p_joint_set_axis2(a, jointId, config["axis2"]["x"], config["axis2"]["y"], config["axis2"]["z"]);
    applyLimotConfig(true);
    return true;
}

// The reference has no BallJoint class, only the JOINT_TYPE_BALL arms of Joint.cpp (:299-301, :447-449,
// dJointSetBallAnchor / dJointGetBallAnchor); this is the class those arms presuppose.
bool BallJoint::createJoint(int b1, int b2)
{
    if (b1 < 0 && b2 < 0) return false;
    b2p_arena *a = theWorld->getWorld();
    jointId = b2p_joint_create(a, B2P_JOINT_BALL, b1, b2);
    if (jointId < 0) return false;
    b2p_joint_set_anchor(a, jointId, config["anchor"]["x"], config["anchor"]["y"], config["anchor"]["z"]);
    return true;
}

// JointFactory.cpp:26-55: unknown type / failed creation -> nullptr
JointFactory &JointFactory::Instance()
{
    static JointFactory instance;
    return instance;
}
std::shared_ptr<Joint> JointFactory::createJoint(WorldPhysics *world, data_broker::DataBrokerInterface *dataBroker,
                                                 configmaps::ConfigMap &config)
{
    auto it = availableJoints.find(config["type"].getString());
    if (it == availableJoints.end()) {
        fprintf(stderr, "Could not load unknown Physics Joint of type: %s\n", config["type"].getString().c_str());
        return nullptr;
    }
    std::shared_ptr<Joint> newJoint(it->second(world, dataBroker, config));
    if (newJoint->isJointCreated()) return newJoint;
    std::cerr << "Failed to create Physics Joint with name: \"" + config["name"].getString() + "\"" << std::endl;
    return nullptr;
}
void JointFactory::addJointType(const std::string &type, instantiateJointfPtr funcPtr)
{
    availableJoints.insert(std::pair<const std::string, instantiateJointfPtr>(type, funcPtr));
}

} // namespace ode_physics
} // namespace mars

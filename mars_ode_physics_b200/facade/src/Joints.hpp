// Joints.hpp -- joint facade: the common Joint base plus Hinge / Fixed / Slider / Hinge2 and the
// factory (reference src/joints/Joint.hpp:15-125, HingeJoint / FixedJoint / SliderJoint /
// Hinge2Joint.{hpp,cpp}, JointFactory.{hpp,cpp}).  Sign conventions, parameter plumbing and the
// post-step torque bookkeeping follow src/joints/Joint.cpp.
#pragma once
#include <map>

#include "Frame.hpp"

namespace mars {
namespace ode_physics {

class Joint : public interfaces::JointInterface {
public:
    Joint(WorldPhysics *world, data_broker::DataBrokerInterface *dataBroker, configmaps::ConfigMap &config);
    virtual ~Joint(void);
    virtual bool createJoint();
    virtual bool createJoint(int body1, int body2);

    virtual void getName(std::string *name) const override;
    virtual interfaces::JointType getType() const override;
    virtual void getAnchor(utils::Vector *anchor) const override;
    virtual void setAnchor(const utils::Vector &anchor) override;
    virtual void setAxis(const utils::Vector &axis) override;
    virtual void setAxis2(const utils::Vector &axis) override;
    virtual void getAxis(utils::Vector *axis) const override;
    virtual void getAxis2(utils::Vector *axis) const override;
    virtual void setWorldObject(std::shared_ptr<interfaces::PhysicsInterface> world) override;
    virtual void setForceLimit(interfaces::sReal max_force) override;
    virtual void setForceLimit2(interfaces::sReal max_force) override;
    virtual void setVelocity(interfaces::sReal velocity) override;
    virtual void setVelocity2(interfaces::sReal velocity) override;
    virtual interfaces::sReal getVelocity(void) const override;
    virtual interfaces::sReal getVelocity2(void) const override;
    virtual void setJointAsMotor(int axis) override;
    virtual void unsetJointAsMotor(int axis) override;
    virtual interfaces::sReal getPosition(void) const override;
    virtual interfaces::sReal getPosition2(void) const override;
    virtual void getForce1(utils::Vector *f) const override;
    virtual void getForce2(utils::Vector *f) const override;
    virtual void getTorque1(utils::Vector *t) const override;
    virtual void getTorque2(utils::Vector *t) const override;
    virtual void setTorque(interfaces::sReal torque) override;
    virtual void setTorque2(interfaces::sReal torque) override;
    virtual void reattacheJoint(void) override;
    virtual void update(void) override;
    virtual void getJointLoad(utils::Vector *t) const override;
    virtual void getAxisTorque(utils::Vector *t) const override;
    virtual void getAxis2Torque(utils::Vector *t) const override;
    virtual void changeStepSize(const interfaces::JointData &jointS) override;
    virtual interfaces::sReal getMotorTorque(void) const override;
    virtual interfaces::sReal getLowStop() const override;
    virtual interfaces::sReal getHighStop() const override;
    virtual interfaces::sReal getLowStop2() const override;
    virtual interfaces::sReal getHighStop2() const override;
    virtual interfaces::sReal getCFM() const override;
    virtual void setLowStop(interfaces::sReal lowStop) override;
    virtual void setHighStop(interfaces::sReal highStop) override;
    virtual void setLowStop2(interfaces::sReal lowStop2) override;
    virtual void setHighStop2(interfaces::sReal highStop2) override;
    virtual void setCFM(interfaces::sReal cfm) override;
    virtual void setJointFrictionCoefficient(interfaces::sReal friction) override;
    virtual interfaces::sReal getJointFrictionCoefficient() override;
    virtual void getDataBrokerNames(std::string &groupName, std::string &dataName) override;

    bool isJointCreated();
    int getJointId() const { return jointId; }
    // ---- batch extension ----
    void setActiveWorld(int w) { world_index = w; }
    void setVelocityBatch(const float *per_world_velocity); // same sign rule as setVelocity

protected:
    void setParam(int param, double value);           // all worlds
    double getParam(int param) const;
    void applyLimotConfig(bool two_axes);
    WorldPhysics *theWorld;
    configmaps::ConfigMap config;
    int jointId;
    mutable double feedback[13];
    int body1, body2;
    int joint_type;
    double cfm, cfm1, cfm2, erp1, erp2;
    double lo1, lo2, hi1, hi2;
    double damping, spring, jointCFM;
    // Joint::update results of the active world, fetched from the device on demand (once per step)
    void fetchPostStep() const;
    mutable utils::Vector axis1_torque, axis2_torque, joint_load;
    mutable double motor_torque;
    mutable long long fetched_step;
    double frictionCoefficient;
    std::string name;
    std::weak_ptr<Frame> parentFrame, childFrame;
    int world_index;

    void calculateCfmErp();
    void getAxisI(utils::Vector *axis) const;
    interfaces::sReal getVelocityI(void) const;

private:
    bool joint_created;
};

#define B2P_DECLARE_JOINT(Name)                                                                          \
    class Name : public Joint {                                                                          \
    public:                                                                                              \
        Name(WorldPhysics *world, data_broker::DataBrokerInterface *dataBroker, configmaps::ConfigMap &config); \
        static Joint *instantiate(WorldPhysics *world, data_broker::DataBrokerInterface *dataBroker,    \
                                  configmaps::ConfigMap &config);                                        \
        virtual bool createJoint(int body1, int body2) override;                                         \
    };
B2P_DECLARE_JOINT(HingeJoint)
B2P_DECLARE_JOINT(FixedJoint)
B2P_DECLARE_JOINT(SliderJoint)
B2P_DECLARE_JOINT(Hinge2Joint)
B2P_DECLARE_JOINT(UniversalJoint)
B2P_DECLARE_JOINT(BallJoint)
#undef B2P_DECLARE_JOINT

typedef Joint *(*instantiateJointfPtr)(WorldPhysics *world, data_broker::DataBrokerInterface *dataBroker,
                                       configmaps::ConfigMap &config);

class JointFactory {
public:
    static JointFactory &Instance();
    std::shared_ptr<Joint> createJoint(WorldPhysics *world, data_broker::DataBrokerInterface *dataBroker,
                                       configmaps::ConfigMap &config);
    void addJointType(const std::string &type, instantiateJointfPtr funcPtr);
    bool hasJointType(const std::string &type) const { return availableJoints.count(type) != 0; }
    const std::map<const std::string, instantiateJointfPtr> &getAvailableJoints() const { return availableJoints; }

private:
    JointFactory() {}
    std::map<const std::string, instantiateJointfPtr> availableJoints;
};

} // namespace ode_physics
} // namespace mars

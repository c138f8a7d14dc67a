// Frame.hpp -- rigid body facade with the shape of the reference's Frame (src/objects/Frame.hpp:24-118).
#pragma once
#include "MassProps.hpp"
#include "WorldPhysics.hpp"

namespace mars {
namespace ode_physics {

class Object;

class Frame : public interfaces::DynamicObject {
public:
    Frame(interfaces::PhysicsInterface *world, data_broker::DataBrokerInterface *dataBroker, configmaps::ConfigMap &config);
    ~Frame(void);

    void addObject(Object *);
    void removeObject(Object *object);
    virtual const std::string &getName() const override;
    const utils::Vector &getContactForceVector() const;
    const interfaces::sReal &getContactForce() const;
    void computeContactForce();
    void clearContactData();

    virtual void getPosition(utils::Vector *pos) const override;
    virtual void setPosition(const utils::Vector &pos) override;
    virtual void getRotation(utils::Quaternion *q) const override;
    virtual void setRotation(const utils::Quaternion &q) override;
    virtual void getPose(utils::Vector *position, utils::Quaternion *rotation) const override;
    virtual void setPose(const utils::Vector &position, const utils::Quaternion &rotation,
                         const bool reset_velocities = false) override;
    virtual void addLinkedFrame(std::shared_ptr<DynamicObject> linked) override;
    virtual bool isLinkedFrame(std::shared_ptr<DynamicObject> linked) override;
    virtual std::vector<std::shared_ptr<DynamicObject>> getLinkedFrames(void) override;
    void getLinearVelocity(utils::Vector *vel) const;
    void getLinearAcceleration(utils::Vector *vel) const;
    void getAngularAcceleration(utils::Vector *vel) const;
    void getAngularVelocity(utils::Vector *vel) const;
    void getForce(utils::Vector *f) const;
    void getTorque(utils::Vector *t) const;
    virtual const utils::Vector rotateAtPoint(const utils::Vector &rotation_point, const utils::Quaternion &rotation,
                                              bool move_group) override;
    virtual void setLinearVelocity(const utils::Vector &velocity);
    virtual void setAngularVelocity(const utils::Vector &velocity);
    void setForce(const utils::Vector &f);
    void setTorque(const utils::Vector &t);
    virtual void addForce(const utils::Vector &f, const utils::Vector &p) override;
    virtual void addForce(const utils::Vector &f) override;
    void addTorque(const utils::Vector &t);
    virtual void addContact(const interfaces::ContactData &contact) override;
    virtual void getMass(interfaces::sReal *mass, interfaces::sReal *inertia = 0) const;

    int getBody() const;   // b2p body id, -1 if none (reference: dBodyID getBody())
    WorldPhysics *getWorldPhysics() const { return theWorld; }
    const utils::Vector &getComOffset() const { return offsetPos; } // accumulated COM shift in the body frame
    int acquireBody();
    MassProps getMass(void) const;
    void addObjectMass(Object *object);
    void update();

    // ---- batch extension: which replica the one-world accessors talk to (default 0) ----
    void setActiveWorld(int w) { world_index = w; }

protected:
    WorldPhysics *theWorld;
    std::string name;
    int nBody;
    MassProps nMass;
    utils::Vector position;
    utils::Quaternion q;
    std::vector<Object *> objects;

private:
    bool body_pose(double pos[3], double quat[4]) const;
    void rotated_offset(const double quat[4], double out[3]) const;
    std::vector<std::weak_ptr<DynamicObject>> connectedFrames;
    utils::Vector offsetPos;
    bool ground_contact;
    std::vector<int> contactIds;
    // post-step values of the active world, fetched from the device on demand (once per step)
    void fetchPostStep() const;
    mutable utils::Vector contactForceVector, l_acc, a_acc;
    mutable interfaces::sReal contactForce;
    mutable long long post_step;
    interfaces::sReal linearDamping, angularDamping;
    int world_index;
};

} // namespace ode_physics
} // namespace mars

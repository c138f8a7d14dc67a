// WorldPhysics.hpp -- B200 facade with the shape of the reference's WorldPhysics
// (reference src/WorldPhysics.hpp:56-111).  The ODE world behind it is replaced by a b2p_arena:
// the scene built through this interface is the template, replicated over `num_worlds` device
// worlds (num_worlds = 1 reproduces the reference's one-world behaviour).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/mars_ode_b200.h"
#include "../compat/mars_compat.hpp"

namespace mars {
namespace ode_physics {

class Object;
class Frame;

class WorldPhysics : public interfaces::PhysicsInterface {
public:
    explicit WorldPhysics();
    virtual ~WorldPhysics(void) override;
    virtual std::string getLibName(void) override;
    virtual void initTheWorld(void) override;
    virtual void freeTheWorld(void) override;
    virtual void stepTheWorld(void) override;
    virtual void clearPreviousStep(void) override;
    virtual bool existsWorld(void) const override;
    virtual const utils::Vector getCenterOfMass(const std::vector<std::shared_ptr<interfaces::NodeInterface>> &nodes) const override;
    virtual std::shared_ptr<interfaces::DynamicObject> createFrame(data_broker::DataBrokerInterface *dataBroker,
                                                                   configmaps::ConfigMap &config) override;
    virtual std::shared_ptr<interfaces::DynamicObject> getFrame(const std::string &frame) override;
    virtual void removeFrame(const std::string &name) override;
    virtual void createObject(configmaps::ConfigMap &config) override;
    virtual std::shared_ptr<interfaces::JointInterface> createJoint(data_broker::DataBrokerInterface *dataBroker,
                                                                    configmaps::ConfigMap &config) override;
    virtual void destroyJoint(const std::string &jointName) override;
    virtual std::shared_ptr<interfaces::JointInterface> getJoint(std::string name) override;
    virtual void exportWorld() override;

    void addFrame(std::shared_ptr<Frame> frame);
    // used by the other physics classes (reference: getWorld / getWorldStep / getFrameIntern /
    // createContact / computeContactForces, src/WorldPhysics.hpp:80-84)
    b2p_arena *getWorld(void) const;
    interfaces::sReal getWorldStep(void);
    std::shared_ptr<Frame> getFrameIntern(const std::string &name);
    // returns the contact id or -1 when the bodies are joint-connected (WorldPhysics.cpp:458-468)
    int createContact(const b2p_contact_desc &c, int body1, int body2, int world = 0);
    void computeContactForces();

    // ---- batch extension (not in the reference): call before initTheWorld ----
    void setNumWorlds(int n);
    int getNumWorlds() const { return num_worlds; }
    void setDevice(int d) { device = d; }
    void setMaxContacts(int n) { max_contacts = n; }
    // rolling / spinning friction rows (ContactData::rolling_friction ...) double the row capacity reserved per
    // contact; on by default as the reference always honours them (Frame.cpp:1049-1054)
    void setRollingFriction(bool on) { rolling_friction = on; }
    long long getStepCount() const { return step_count; }
    // true if the last stepTheWorld() could not step (fast_step == false: dWorldStep is not implemented)
    bool lastStepFailed() const { return step_failed; }

    mutable utils::Mutex iMutex;
    interfaces::sReal max_angular_speed;
    interfaces::sReal max_correcting_vel;
    static interfaces::PhysicsError error;

private:
    b2p_arena *arena;
    bool world_init;
    int num_worlds, device, max_contacts;
    bool rolling_friction, step_failed, warned_capacity;
    long long step_count;
    utils::Vector old_gravity;
    interfaces::sReal old_cfm, old_erp;
    std::map<std::string, std::weak_ptr<Frame>> frameMap;
    std::map<std::string, std::weak_ptr<interfaces::JointInterface>> jointMap;
    std::vector<Object *> objects;
    void clearObjects();
    void preStepChecks(void);
};

void checkB2P(int rc, const char *what); // throws std::runtime_error with b2p_last_error()

} // namespace ode_physics
} // namespace mars

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
    void noteJoint(int joint_id, int joint_type);
    void computeContactForces();

    // ---- batch extension (not in the reference): call before initTheWorld ----
    void setNumWorlds(int n);
    int getNumWorlds() const { return num_worlds; }
    void setDevice(int d) { device = d; }
    void setMaxContacts(int n) { max_contacts = n; }
    // rolling / spinning friction rows (ContactData::rolling_friction ...) double the row capacity reserved per
    // contact; on by default as the reference always honours them (Frame.cpp:1049-1054)
    void setRollingFriction(bool on) { rolling_friction = on; }
    // ---- device collision: broadphase + narrowphase on the device as the contact source.  With it
    // switched on (before the objects are created) every box / sphere / capsule / cylinder object also becomes a
    // geom on its frame's body, stepTheWorld() runs collide + step, and the generated contacts can be read back
    // in the reference's ContactData form.  Surface parameters: an object's config["c_params"] (friction1,
    // friction2, rolling_friction, rolling_friction2, spinning_friction, cfm, erp) or, where absent, the world's
    // ground_friction / ground_cfm / ground_erp (reference src/WorldPhysics.cpp:86-88, unused there).
    void setDeviceCollision(bool on);
    bool getDeviceCollision() const { return device_collision; }
    void createGroundPlane(double a, double b, double c, double d);            // a x + b y + c z = d, static
    void createHeightfield(const float *heights, int nx, int ny, double dx, double x0, double y0); // shared by all worlds
    b2p_surface defaultSurface() const;
    // the contact stream of one world as the records Frame::addContact consumes (Frame.cpp:996-1092)
    std::vector<interfaces::ContactData> getContacts(int world = 0);
    void requestGeom(Frame *frame, int geom_class, const double dims[4], const utils::Vector &pos,
                     const utils::Quaternion &q, const b2p_surface &s);
    // ---- batched I/O for all worlds (batch extension; the one-world accessors keep working next to it) ----
    // per-world motor commands of joints 0 .. n-1 in creation order: values[n][num_worlds], axis 1 or 2, with the
    // sign rule of Joint::setVelocity (hinge / slider negate; Joint.cpp:355-393) applied on the way
    void setJointVelocitiesBatch(const float *values, int num_joints, int axis);
    // state of one frame's body (13 floats per world) + angle / rate of every joint of every world, copied to
    // host buffers over a copy stream while the next step runs; returns a ticket for observeBatchWait
    int observeBatchBegin(const std::shared_ptr<interfaces::DynamicObject> &frame, float *body13, float *joint_obs);
    void observeBatchWait(int ticket);
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
    bool device_collision, geoms_dirty;
    struct GeomRequest {
        Frame *frame;
        int cls;
        double dims[4];
        utils::Vector pos;
        utils::Quaternion q;
        b2p_surface s;
    };
    std::vector<GeomRequest> geom_requests;
    std::vector<float> batch_tmp;
    std::vector<int> joint_types_by_id; // interfaces::JointType per arena joint id (sign rule of the batch setter)
    void flushGeoms();
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

// mars_compat.hpp -- minimal stand-ins for the MARS packages the reference plugin is written against
// but which are not available here: mars_utils (Vector, Quaternion, Mutex), configmaps (ConfigMap),
// mars_interfaces (PhysicsInterface, DynamicObject, JointInterface, ContactData), data_broker and
// lib_manager.  Only the surface the hot path touches is provided; names,
// argument meaning and defaults follow the reference headers' use sites:
//   src/WorldPhysics.hpp:56-93, src/objects/Frame.hpp:41-91, src/joints/Joint.hpp:30-83,
//   src/objects/Frame.cpp:1003-1071 (ContactData fields), src/WorldPhysicsLoader.hpp:19-39.
// In a real MARS build these come from the respective packages and this header is not used.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace mars {
namespace utils {

class Vector {
public:
    Vector() : v{0, 0, 0} {}
    Vector(double x, double y, double z) : v{x, y, z} {}
    double &x() { return v[0]; }
    double &y() { return v[1]; }
    double &z() { return v[2]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    Vector normalized() const
    {
        const double n = norm();
        return n > 0 ? Vector(v[0] / n, v[1] / n, v[2] / n) : *this;
    }
    void setZero() { v[0] = v[1] = v[2] = 0; }
    Vector operator+(const Vector &o) const { return Vector(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vector operator-(const Vector &o) const { return Vector(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector operator*(double s) const { return Vector(v[0] * s, v[1] * s, v[2] * s); }
    Vector operator/(double s) const { return Vector(v[0] / s, v[1] / s, v[2] / s); }
    Vector &operator-=(const Vector &o) { return *this = *this - o; }
    Vector &operator+=(const Vector &o) { return *this = *this + o; }
    Vector &operator*=(double s) { return *this = *this * s; }
    bool operator!=(const Vector &o) const { return v[0] != o.v[0] || v[1] != o.v[1] || v[2] != o.v[2]; }
    bool operator==(const Vector &o) const { return !(*this != o); }

private:
    double v[3];
};

// constructor order (w, x, y, z) as Eigen::Quaterniond
class Quaternion {
public:
    Quaternion() : q{1, 0, 0, 0} {}
    Quaternion(double w, double x, double y, double z) : q{w, x, y, z} {}
    double &w() { return q[0]; }
    double &x() { return q[1]; }
    double &y() { return q[2]; }
    double &z() { return q[3]; }
    double w() const { return q[0]; }
    double x() const { return q[1]; }
    double y() const { return q[2]; }
    double z() const { return q[3]; }
    Vector operator*(const Vector &v) const; // rotate

private:
    double q[4];
};

inline Vector Quaternion::operator*(const Vector &p) const
{
    // v' = v + 2 w (u x v) + 2 u x (u x v)
    const double ux = q[1], uy = q[2], uz = q[3], w0 = q[0];
    const double cx = uy * p.z() - uz * p.y(), cy = uz * p.x() - ux * p.z(), cz = ux * p.y() - uy * p.x();
    const double dx = uy * cz - uz * cy, dy = uz * cx - ux * cz, dz = ux * cy - uy * cx;
    return Vector(p.x() + 2 * (w0 * cx + dx), p.y() + 2 * (w0 * cy + dy), p.z() + 2 * (w0 * cz + dz));
}

inline Quaternion angleAxisToQuaternion(double angle, const Vector &axis)
{
    const Vector n = axis.normalized();
    const double s = std::sin(0.5 * angle);
    return Quaternion(std::cos(0.5 * angle), n.x() * s, n.y() * s, n.z() * s);
}

class Mutex {
public:
    void lock() { m.lock(); }
    void unlock() { m.unlock(); }

private:
    std::recursive_mutex m;
};

class MutexLocker {
public:
    explicit MutexLocker(Mutex *m) : m(m)
    {
        if (m) m->lock();
    }
    ~MutexLocker()
    {
        if (m) m->unlock();
    }

private:
    Mutex *m;
};

} // namespace utils
} // namespace mars

namespace configmaps {

// A tiny recursive map: enough for the keys the object / joint constructors read.
class ConfigItem {
public:
    ConfigItem() = default;
    ConfigItem &operator[](const std::string &k) { return children[k]; }
    ConfigItem &operator[](const char *k) { return children[std::string(k)]; }
    const ConfigItem &at(const std::string &k) const { return children.at(k); }
    bool hasKey(const std::string &k) const { return children.count(k) != 0; }
    ConfigItem &operator=(double d)
    {
        num = d;
        isNum = true;
        return *this;
    }
    ConfigItem &operator=(int d) { return *this = (double)d; }
    ConfigItem &operator=(bool b) { return *this = (double)(b ? 1 : 0); }
    ConfigItem &operator=(const char *s_)
    {
        str = s_;
        isNum = false;
        return *this;
    }
    ConfigItem &operator=(const std::string &s_) { return *this = s_.c_str(); }
    operator double() const { return isNum ? num : (str.empty() ? 0.0 : std::stod(str)); }
    operator int() const { return (int)(double)*this; }
    operator bool() const { return (double)*this != 0.0; }
    operator std::string() const { return toString(); }
    std::string getString() const { return str; }
    std::string toString() const { return isNum ? std::to_string(num) : str; }
    void erase(const std::string &k) { children.erase(k); }

private:
    std::map<std::string, ConfigItem> children;
    std::string str;
    double num = 0;
    bool isNum = false;
};
using ConfigMap = ConfigItem;

inline void operator<<(ConfigItem &item, const std::string &s) { item = s; }
inline void operator<<(std::string &s, const ConfigItem &item) { s = item.toString(); }

} // namespace configmaps

namespace mars {
namespace utils {
inline void vectorFromConfigItem(configmaps::ConfigItem &item, Vector *v)
{
    v->x() = item["x"];
    v->y() = item["y"];
    v->z() = item["z"];
}
inline void quaternionFromConfigItem(configmaps::ConfigItem &item, Quaternion *q)
{
    q->x() = item["x"];
    q->y() = item["y"];
    q->z() = item["z"];
    q->w() = item["w"];
}
} // namespace utils
} // namespace mars

namespace data_broker {
class DataBrokerInterface; // telemetry is out of scope: always passed as nullptr
}

namespace lib_manager {
class LibManager;
class LibInterface {
public:
    explicit LibInterface(LibManager *m) : libManager(m) {}
    virtual ~LibInterface() {}
    virtual int getLibVersion() const = 0;
    virtual const std::string getLibName() const = 0;

protected:
    LibManager *libManager;
};
} // namespace lib_manager
// lib_manager's factory convention: extern "C" create_c / destroy_c (names to be confirmed against
// the real lib_manager header)
#define CREATE_LIB(theClass) \
    extern "C" lib_manager::LibInterface *create_c(lib_manager::LibManager *m) { return new theClass(m); }
#define DESTROY_LIB(theClass) \
    extern "C" void destroy_c(lib_manager::LibInterface *p) { delete p; }

namespace mars {
namespace interfaces {

typedef double sReal;

enum JointType {
    JOINT_TYPE_UNDEFINED = 0,
    JOINT_TYPE_HINGE,
    JOINT_TYPE_HINGE2,
    JOINT_TYPE_SLIDER,
    JOINT_TYPE_BALL,
    JOINT_TYPE_UNIVERSAL,
    JOINT_TYPE_FIXED,
    NUMBER_OF_JOINT_TYPES
};

inline JointType getJointType(const std::string &s)
{
    if (s == "hinge") return JOINT_TYPE_HINGE;
    if (s == "hinge2") return JOINT_TYPE_HINGE2;
    if (s == "slider" || s == "prismatic") return JOINT_TYPE_SLIDER;
    if (s == "ball") return JOINT_TYPE_BALL;
    if (s == "universal") return JOINT_TYPE_UNIVERSAL;
    if (s == "fixed") return JOINT_TYPE_FIXED;
    return JOINT_TYPE_UNDEFINED;
}

enum PhysicsError { PHYSICS_NO_ERROR = 0, PHYSICS_DEBUG, PHYSICS_ERROR, PHYSICS_UNKNOWN };

struct contact_params {
    sReal cfm = 0, erp = 0, friction1 = 0, friction2 = 0;
    sReal rolling_friction = 0, rolling_friction2 = 0, spinning_friction = 0;
    sReal bounce = 0, bounce_vel = 0, motion1 = 0, motion2 = 0, fds1 = 0, fds2 = 0;
};

class DynamicObject;

struct ContactData {
    std::shared_ptr<DynamicObject> body1, body2;
    std::string nameObject1, nameObject2;
    utils::Vector pos, normal;
    sReal depth = 0;
    contact_params c_params;
};

struct JointData {
    JointType type = JOINT_TYPE_UNDEFINED;
};

class DynamicObject {
public:
    virtual ~DynamicObject() {}
    virtual const std::string &getName() const = 0;
    virtual void getPosition(utils::Vector *pos) const = 0;
    virtual void setPosition(const utils::Vector &pos) = 0;
    virtual void getRotation(utils::Quaternion *q) const = 0;
    virtual void setRotation(const utils::Quaternion &q) = 0;
    virtual void getPose(utils::Vector *position, utils::Quaternion *rotation) const = 0;
    virtual void setPose(const utils::Vector &position, const utils::Quaternion &rotation,
                         const bool reset_velocities = false) = 0;
    virtual void addLinkedFrame(std::shared_ptr<DynamicObject> linked) = 0;
    virtual bool isLinkedFrame(std::shared_ptr<DynamicObject> linked) = 0;
    virtual std::vector<std::shared_ptr<DynamicObject>> getLinkedFrames(void) = 0;
    virtual const utils::Vector rotateAtPoint(const utils::Vector &rotation_point, const utils::Quaternion &rotation,
                                              bool move_group) = 0;
    virtual void addForce(const utils::Vector &f, const utils::Vector &p) = 0;
    virtual void addForce(const utils::Vector &f) = 0;
    virtual void addContact(const ContactData &contact) = 0;
};

class PhysicsInterface;

class JointInterface {
public:
    virtual ~JointInterface() {}
    virtual void getName(std::string *name) const = 0;
    virtual JointType getType() const = 0;
    virtual void getAnchor(utils::Vector *anchor) const = 0;
    virtual void setAnchor(const utils::Vector &anchor) = 0;
    virtual void setAxis(const utils::Vector &axis) = 0;
    virtual void setAxis2(const utils::Vector &axis) = 0;
    virtual void getAxis(utils::Vector *axis) const = 0;
    virtual void getAxis2(utils::Vector *axis) const = 0;
    virtual void setWorldObject(std::shared_ptr<PhysicsInterface> world) = 0;
    virtual void setForceLimit(sReal max_force) = 0;
    virtual void setForceLimit2(sReal max_force) = 0;
    virtual void setVelocity(sReal velocity) = 0;
    virtual void setVelocity2(sReal velocity) = 0;
    virtual sReal getVelocity(void) const = 0;
    virtual sReal getVelocity2(void) const = 0;
    virtual void setJointAsMotor(int axis) = 0;
    virtual void unsetJointAsMotor(int axis) = 0;
    virtual sReal getPosition(void) const = 0;
    virtual sReal getPosition2(void) const = 0;
    virtual void getForce1(utils::Vector *f) const = 0;
    virtual void getForce2(utils::Vector *f) const = 0;
    virtual void getTorque1(utils::Vector *t) const = 0;
    virtual void getTorque2(utils::Vector *t) const = 0;
    virtual void setTorque(sReal torque) = 0;
    virtual void setTorque2(sReal torque) = 0;
    virtual void reattacheJoint(void) = 0;
    virtual void update(void) = 0;
    virtual void getJointLoad(utils::Vector *t) const = 0;
    virtual void getAxisTorque(utils::Vector *t) const = 0;
    virtual void getAxis2Torque(utils::Vector *t) const = 0;
    virtual void changeStepSize(const JointData &jointS) = 0;
    virtual sReal getMotorTorque(void) const = 0;
    virtual sReal getLowStop() const = 0;
    virtual sReal getHighStop() const = 0;
    virtual sReal getLowStop2() const = 0;
    virtual sReal getHighStop2() const = 0;
    virtual sReal getCFM() const = 0;
    virtual void setLowStop(sReal lowStop) = 0;
    virtual void setHighStop(sReal highStop) = 0;
    virtual void setLowStop2(sReal lowStop2) = 0;
    virtual void setHighStop2(sReal highStop2) = 0;
    virtual void setCFM(sReal cfm) = 0;
    virtual void setJointFrictionCoefficient(sReal friction) = 0;
    virtual sReal getJointFrictionCoefficient() = 0;
    virtual void getDataBrokerNames(std::string &groupName, std::string &dataName) = 0;
};

class NodeInterface;

class PhysicsInterface {
public:
    virtual ~PhysicsInterface() {}
    virtual std::string getLibName(void) = 0;
    virtual void initTheWorld(void) = 0;
    virtual void freeTheWorld(void) = 0;
    virtual void stepTheWorld(void) = 0;
    virtual void clearPreviousStep(void) = 0;
    virtual bool existsWorld(void) const = 0;
    virtual const utils::Vector getCenterOfMass(const std::vector<std::shared_ptr<NodeInterface>> &nodes) const = 0;
    virtual std::shared_ptr<DynamicObject> createFrame(data_broker::DataBrokerInterface *dataBroker,
                                                       configmaps::ConfigMap &config) = 0;
    virtual std::shared_ptr<DynamicObject> getFrame(const std::string &frame) = 0;
    virtual void removeFrame(const std::string &name) = 0;
    virtual void createObject(configmaps::ConfigMap &config) = 0;
    virtual std::shared_ptr<JointInterface> createJoint(data_broker::DataBrokerInterface *dataBroker,
                                                        configmaps::ConfigMap &config) = 0;
    virtual void destroyJoint(const std::string &jointName) = 0;
    virtual std::shared_ptr<JointInterface> getJoint(std::string name) = 0;
    virtual void exportWorld() = 0;

    // public knobs, assigned in the WorldPhysics constructor (src/WorldPhysics.cpp:82-99)
    int fast_step = 0;
    sReal world_cfm = 1e-10, world_erp = 0.1;
    utils::Vector world_gravity{0.0, 0.0, -9.81};
    sReal ground_friction = 20, ground_cfm = 1e-8, ground_erp = 0.1;
    sReal step_size = 0.01;
};

} // namespace interfaces
} // namespace mars

// Objects.hpp -- mass-only scene objects and their factory (reference src/objects/Object.hpp,
// Box/Sphere/Capsule/Cylinder/Inertial.{hpp,cpp}, ObjectFactory.{hpp,cpp}).  As in the reference
// an object contributes a mass to its Frame; with WorldPhysics::setDeviceCollision(true) its shape additionally
// becomes a geom of the device collision path (the reference keeps collision in a separate plugin).
#pragma once
#include <map>

#include "Frame.hpp"

namespace mars {
namespace ode_physics {

class Object {
public:
    explicit Object(std::shared_ptr<Frame> frame);
    virtual ~Object(void);
    virtual void getPosition(utils::Vector *pos) const;
    virtual const utils::Vector setPosition(const utils::Vector &pos, bool move_group = false);
    virtual void getRotation(utils::Quaternion *q) const;
    virtual const utils::Quaternion setRotation(const utils::Quaternion &q, bool move_group = false);
    virtual void getMass(interfaces::sReal *mass, interfaces::sReal *inertia = 0) const;
    MassProps getMass(void) const;
    virtual bool createMass() = 0;
    bool isObjectCreated() const { return objectCreated; }

protected:
    void readPose(configmaps::ConfigMap &config);
    void registerOnFrame();
    // device collision (WorldPhysics::setDeviceCollision): this object's shape as a geom on the frame's body
    void requestGeom(configmaps::ConfigMap &config, int geom_class, double d0, double d1, double d2);
    std::weak_ptr<Frame> frame;
    utils::Vector pos;
    utils::Quaternion q;
    MassProps nMass;
    bool objectCreated;
};

#define B2P_DECLARE_OBJECT(Name)                                                                   \
    class Name : public Object {                                                                   \
    public:                                                                                        \
        Name(std::shared_ptr<Frame> frame, configmaps::ConfigMap &config);                         \
        static Object *instantiate(std::shared_ptr<Frame> frame, configmaps::ConfigMap &config);   \
        virtual bool createMass() override;                                                        \
                                                                                                   \
    protected:                                                                                     \
        configmaps::ConfigMap config;                                                              \
    };
B2P_DECLARE_OBJECT(Box)
B2P_DECLARE_OBJECT(Sphere)
B2P_DECLARE_OBJECT(Capsule)
B2P_DECLARE_OBJECT(Cylinder)
B2P_DECLARE_OBJECT(Inertial)
#undef B2P_DECLARE_OBJECT

typedef Object *(*instantiateObjectfPtr)(std::shared_ptr<Frame> frame, configmaps::ConfigMap &config);

class ObjectFactory {
public:
    static ObjectFactory &Instance();
    Object *createObject(std::shared_ptr<Frame> frame, const std::string &type, configmaps::ConfigMap &config);
    void addObjectType(const std::string &type, instantiateObjectfPtr funcPtr);
    bool hasObjectType(const std::string &type) const { return availableObjects.count(type) != 0; }
    const std::map<const std::string, instantiateObjectfPtr> &getAvailableObjects() const { return availableObjects; }

private:
    ObjectFactory() {}
    std::map<const std::string, instantiateObjectfPtr> availableObjects;
};

} // namespace ode_physics
} // namespace mars

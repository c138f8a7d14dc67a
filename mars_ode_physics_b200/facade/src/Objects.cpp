// Objects.cpp -- see Objects.hpp.
#include "Objects.hpp"

#include <cstdio>

namespace mars {
namespace ode_physics {

using namespace utils;
using namespace interfaces;
using namespace configmaps;

Object::Object(std::shared_ptr<Frame> frame) : frame(frame), pos(0.0, 0.0, 0.0), q(1.0, 0.0, 0.0, 0.0), objectCreated(false) {}
Object::~Object(void) {}
void Object::getPosition(Vector *p) const { *p = pos; }
const Vector Object::setPosition(const Vector &p, bool)
{
    pos = p;
    return Vector();
}
void Object::getRotation(Quaternion *out) const { *out = q; }
const Quaternion Object::setRotation(const Quaternion &nq, bool)
{
    q = nq;
    return Quaternion();
}
void Object::getMass(sReal *mass, sReal *inertia) const
{
    if (mass) *mass = nMass.mass;
    if (inertia)
        for (int i = 0; i < 9; i++) inertia[i] = nMass.I[i];
}
MassProps Object::getMass(void) const { return nMass; }

void Object::readPose(ConfigMap &config)
{
    vectorFromConfigItem(config["position"], &pos);
    quaternionFromConfigItem(config["orientation"], &q);
}

void Object::registerOnFrame()
{
    if (auto f = frame.lock()) f->addObject(this);
    objectCreated = true;
}

void Object::requestGeom(ConfigMap &config, int cls, double d0, double d1, double d2)
{
    auto f = frame.lock();
    if (!f) return;
    WorldPhysics *world = f->getWorldPhysics();
    if (!world || !world->getDeviceCollision()) return;
    b2p_surface s = world->defaultSurface();
    if (config.hasKey("c_params")) {
        ConfigMap &c = config["c_params"];
        if (c.hasKey("friction1")) s.mu = s.mu2 = (double)c["friction1"];
        if (c.hasKey("friction2")) s.mu2 = (double)c["friction2"];
        if (c.hasKey("rolling_friction")) s.rho = s.rho2 = (double)c["rolling_friction"];
        if (c.hasKey("rolling_friction2")) s.rho2 = (double)c["rolling_friction2"];
        if (c.hasKey("spinning_friction")) s.rhoN = (double)c["spinning_friction"];
        if (c.hasKey("cfm")) s.soft_cfm = (double)c["cfm"];
        if (c.hasKey("erp")) s.soft_erp = (double)c["erp"];
    }
    const double dims[4] = {d0, d1, d2, 0.0};
    world->requestGeom(f.get(), cls, dims, pos, q, s);
}

#define B2P_OBJECT_BOILERPLATE(Name)                                                                      \
    Name::Name(std::shared_ptr<Frame> frame, ConfigMap &config) : Object(frame), config(config)           \
    {                                                                                                     \
        readPose(config);                                                                                 \
        createMass();                                                                                     \
    }                                                                                                     \
    Object *Name::instantiate(std::shared_ptr<Frame> frame, ConfigMap &config) { return new Name(frame, config); }
B2P_OBJECT_BOILERPLATE(Box)
B2P_OBJECT_BOILERPLATE(Sphere)
B2P_OBJECT_BOILERPLATE(Capsule)
B2P_OBJECT_BOILERPLATE(Cylinder)
B2P_OBJECT_BOILERPLATE(Inertial)
#undef B2P_OBJECT_BOILERPLATE

// The reference calls the density form and then the total form (Box.cpp:39-40): only the total survives.
bool Box::createMass()
{
    nMass.setBoxTotal((double)config["mass"], (double)config["extend"]["x"], (double)config["extend"]["y"],
                      (double)config["extend"]["z"]);
    registerOnFrame();
    requestGeom(config, B2P_GEOM_BOX, (double)config["extend"]["x"], (double)config["extend"]["y"],
                (double)config["extend"]["z"]);
    return true;
}

bool Sphere::createMass() // Sphere.cpp:36-39: radius = extend.x
{
    nMass.setSphereTotal((double)config["mass"], (double)config["extend"]["x"]);
    registerOnFrame();
    requestGeom(config, B2P_GEOM_SPHERE, (double)config["extend"]["x"], 0.0, 0.0);
    return true;
}

bool Capsule::createMass() // Capsule.cpp:37-41: radius = extend.x, length = extend.y, direction 3
{
    nMass.setCapsuleTotal((double)config["mass"], 3, (double)config["extend"]["x"], (double)config["extend"]["y"]);
    registerOnFrame();
    requestGeom(config, B2P_GEOM_CAPSULE, (double)config["extend"]["x"], (double)config["extend"]["y"], 0.0);
    return true;
}

bool Cylinder::createMass() // Cylinder.cpp:37-41
{
    nMass.setCylinderTotal((double)config["mass"], 3, (double)config["extend"]["x"], (double)config["extend"]["y"]);
    registerOnFrame();
    requestGeom(config, B2P_GEOM_CYLINDER, (double)config["extend"]["x"], (double)config["extend"]["y"], 0.0);
    return true;
}

bool Inertial::createMass() // Inertial.cpp:31-93
{
    nMass.setZero();
    const double density = config.hasKey("density") ? (double)config["density"] : 0.0;
    const double mass = config.hasKey("mass") ? (double)config["mass"] : 0.0;
    if (config.hasKey("createFromBox") && (bool)config["createFromBox"]) {
        const double x = config["x"], y = config["y"], z = config["z"];
        if (density > 0.0000001) nMass.setBoxTotal(density * x * y * z, x, y, z);
        else if (mass > 0.0000001) nMass.setBoxTotal(mass, x, y, z);
    } else if (config.hasKey("createFromSphere") && (bool)config["createFromSphere"]) {
        const double r = config["x"];
        if (density > 0.0000001) nMass.setSphereTotal(density * (4.0 / 3.0) * M_PI * r * r * r, r);
        else if (mass > 0.0000001) nMass.setSphereTotal(mass, r);
    } else {
        nMass.mass = mass;
        const char *k[9] = {"i00", "i01", "i02", "i10", "i11", "i12", "i20", "i21", "i22"};
        for (int i = 0; i < 9; i++) nMass.I[i] = (double)config["inertia"][k[i]];
    }
    if (!nMass.check()) fprintf(stderr, "Mass problem in inertial object \"%s\"\n", config["name"].toString().c_str());
    registerOnFrame();
    return true;
}

ObjectFactory &ObjectFactory::Instance()
{
    static ObjectFactory instance;
    return instance;
}

Object *ObjectFactory::createObject(std::shared_ptr<Frame> frame, const std::string &type, ConfigMap &config)
{
    auto it = availableObjects.find(type);
    if (it == availableObjects.end())
        throw std::runtime_error("Could not load unknown Physics Object with name: \"" + config["name"].getString() + "\"");
    return it->second(frame, config);
}

void ObjectFactory::addObjectType(const std::string &type, instantiateObjectfPtr funcPtr)
{
    availableObjects.insert(std::pair<const std::string, instantiateObjectfPtr>(type, funcPtr));
}

} // namespace ode_physics
} // namespace mars

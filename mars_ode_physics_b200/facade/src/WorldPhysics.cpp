// WorldPhysics.cpp -- see WorldPhysics.hpp.  Behaviour follows reference src/WorldPhysics.cpp; the ODE
// calls are replaced by the b2p_* C ABI (include/mars_ode_b200.h).
#include "WorldPhysics.hpp"

#include <cstdio>
#include <iostream>

#include "Frame.hpp"
#include "Joints.hpp"
#include "Objects.hpp"

namespace mars {
namespace ode_physics {

using namespace utils;
using namespace interfaces;

PhysicsError WorldPhysics::error = PHYSICS_NO_ERROR;

void checkB2P(int rc, const char *what)
{
    if (rc < 0) {
        WorldPhysics::error = PHYSICS_ERROR;
        throw std::runtime_error(std::string(what) + ": " + b2p_last_error());
    }
}

// defaults of reference src/WorldPhysics.cpp:80-99
WorldPhysics::WorldPhysics() : interfaces::PhysicsInterface()
{
    fast_step = 0;
    world_cfm = 1e-10;
    world_erp = 0.1;
    world_gravity = Vector{0.0, 0.0, -9.81};
    ground_friction = 20;
    ground_cfm = 0.00000001;
    ground_erp = 0.1;
    arena = nullptr;
    world_init = false;
    max_angular_speed = 10.0;
    max_correcting_vel = 5.0;
    step_size = 0.01;
    num_worlds = 1;
    device = 0;
    max_contacts = 8;
    rolling_friction = true;
    step_failed = warned_capacity = false;
    device_collision = geoms_dirty = false;
    step_count = 0;
}

WorldPhysics::~WorldPhysics(void) { freeTheWorld(); }

std::string WorldPhysics::getLibName(void) { return "mars_ode_physics"; }

void WorldPhysics::setNumWorlds(int n)
{
    if (world_init) throw std::runtime_error("WorldPhysics::setNumWorlds: call before initTheWorld");
    num_worlds = n;
}

// reference :162-203: seed 11211, gravity / cfm / erp / max angular speed / max correcting velocity
void WorldPhysics::initTheWorld(void)
{
    const MutexLocker locker{&iMutex};
    if (world_init) return;
    checkB2P(b2p_arena_create(num_worlds, device, &arena), "WorldPhysics::initTheWorld");
    b2p_arena_set_max_contacts(arena, max_contacts);
    b2p_arena_enable_rolling_friction(arena, rolling_friction ? 1 : 0);
    // rows in ODE's own order: island by island, bodies / joints / contacts as dxProcessIslands visits them
    b2p_world_set_row_order(arena, B2P_ORDER_ODE_ISLANDS);
    b2p_world_rand_set_seed(arena, 11211u);
    old_gravity = world_gravity;
    old_cfm = world_cfm;
    old_erp = world_erp;
    b2p_world_set_gravity(arena, world_gravity.x(), world_gravity.y(), world_gravity.z());
    b2p_world_set_cfm(arena, world_cfm);
    b2p_world_set_erp(arena, world_erp);
    b2p_world_set_max_angular_speed(arena, max_angular_speed);
    b2p_world_set_contact_max_correcting_vel(arena, max_correcting_vel);
    world_init = true;
}

void WorldPhysics::freeTheWorld(void)
{
    clearObjects();
    const MutexLocker locker{&iMutex};
    if (world_init) {
        b2p_arena_destroy(arena);
        arena = nullptr;
        world_init = false;
    }
}

bool WorldPhysics::existsWorld(void) const { return world_init; }

void WorldPhysics::preStepChecks(void)
{
    if (old_gravity != world_gravity) {
        old_gravity = world_gravity;
        b2p_world_set_gravity(arena, world_gravity.x(), world_gravity.y(), world_gravity.z());
    }
    if (old_cfm != world_cfm) {
        old_cfm = world_cfm;
        b2p_world_set_cfm(arena, world_cfm);
    }
    if (old_erp != world_erp) {
        old_erp = world_erp;
        b2p_world_set_erp(arena, world_erp);
    }
}

void WorldPhysics::clearPreviousStep(void)
{
    if (!world_init) return;
    b2p_contacts_clear(arena, B2P_ALL_WORLDS);
    for (auto &pair : frameMap)
        if (auto f = pair.second.lock()) f->clearContactData();
}

void WorldPhysics::computeContactForces()
{
    for (auto &pair : frameMap)
        if (auto f = pair.second.lock()) f->computeContactForce();
}

// reference :349-421.  fast_step selects the stepper exactly as the reference does (:365-370): true runs
// dWorldQuickStep (the SOR-LCP kernels), false -- the constructor's default -- dWorldStep (the dense LCP solved by
// Dantzig's method, b2p_dense.cu; exact and O(rows^3) per world).  Everything the reference does after the ODE call
// (Joint::update, Frame::update, computeContactForces) is part of b2p_step's device epilogue for all worlds; the
// loops below only drop per-object caches.
void WorldPhysics::stepTheWorld(void)
{
    const MutexLocker locker{&iMutex};
    if (!(world_init && step_size > 0)) return;
    preStepChecks();
    step_failed = false;
    {
        try {
            checkB2P(b2p_arena_set_stepper(arena, fast_step ? B2P_STEPPER_QUICK : B2P_STEPPER_DENSE), "WorldPhysics::stepTheWorld");
            if (device_collision) {
                flushGeoms();
                checkB2P(b2p_collide_step(arena, step_size), "WorldPhysics::stepTheWorld");
            } else {
                checkB2P(b2p_step(arena, step_size), "WorldPhysics::stepTheWorld");
            }
            step_count++;
            for (auto it = jointMap.begin(); it != jointMap.end();) {
                if (auto joint = it->second.lock()) {
                    joint->update();
                    ++it;
                } else {
                    it = jointMap.erase(it);
                }
            }
            for (auto it = frameMap.begin(); it != frameMap.end();) {
                if (auto frame = it->second.lock()) {
                    frame->update();
                    ++it;
                } else {
                    it = frameMap.erase(it);
                }
            }
            computeContactForces();
        } catch (const std::exception &e) {
            fprintf(stderr, "WorldPhysics::stepTheWorld: %s\n", e.what());
            step_failed = true;
        }
    }
    if (WorldPhysics::error) WorldPhysics::error = PHYSICS_NO_ERROR;
}

b2p_arena *WorldPhysics::getWorld(void) const { return arena; }
sReal WorldPhysics::getWorldStep(void) { return step_size; }

void WorldPhysics::exportWorld() { std::cout << "EXPORT WORLD (not supported by the device arena)" << std::endl; }

int WorldPhysics::createContact(const b2p_contact_desc &c, int body1, int body2, int world)
{
    const int rc = b2p_contact_add(arena, world, body1, body2, &c);
    if (rc == B2P_EREJECTED) return -1; // joint-connected bodies (reference :461)
    if (rc == B2P_ECAPACITY) {
        // the reference has no contact cap: never throw out of addContact; drop the contact and say so once
        if (!warned_capacity)
            fprintf(stderr, "mars_ode_physics_b200: more than %d contacts in one world; extra contacts are dropped "
                            "(WorldPhysics::setMaxContacts before initTheWorld raises the capacity)\n", max_contacts);
        warned_capacity = true;
        return -2;
    }
    checkB2P(rc, "WorldPhysics::createContact");
    return rc;
}

// ---------------- device collision through the plugin API ----------------
void WorldPhysics::setDeviceCollision(bool on)
{
    if (!geom_requests.empty() && !on) throw std::runtime_error("WorldPhysics::setDeviceCollision: geoms exist already");
    device_collision = on;
}

b2p_surface WorldPhysics::defaultSurface() const
{
    b2p_surface s{};
    s.mu = s.mu2 = ground_friction;
    s.soft_cfm = ground_cfm;
    s.soft_erp = ground_erp;
    return s;
}

void WorldPhysics::createGroundPlane(double a, double b, double c, double d)
{
    const MutexLocker locker{&iMutex};
    if (!world_init) throw std::runtime_error("WorldPhysics::createGroundPlane: world not initialised");
    flushGeoms(); // geoms keep their creation order (it fixes the order of the contact stream)
    const double dims[4] = {a, b, c, d};
    const b2p_surface s = defaultSurface();
    checkB2P(b2p_geom_create(arena, B2P_GEOM_PLANE, B2P_STATIC_BODY, dims, nullptr, nullptr, &s), "createGroundPlane");
    device_collision = true;
}

void WorldPhysics::createHeightfield(const float *heights, int nx, int ny, double dx, double x0, double y0)
{
    const MutexLocker locker{&iMutex};
    if (!world_init) throw std::runtime_error("WorldPhysics::createHeightfield: world not initialised");
    flushGeoms();
    const b2p_surface s = defaultSurface();
    checkB2P(b2p_geom_create_heightfield(arena, heights, nx, ny, dx, x0, y0, &s), "createHeightfield");
    device_collision = true;
}

void WorldPhysics::requestGeom(Frame *frame, int cls, const double dims[4], const Vector &pos, const Quaternion &q,
                               const b2p_surface &s)
{
    GeomRequest r;
    r.frame = frame;
    r.cls = cls;
    for (int i = 0; i < 4; i++) r.dims[i] = dims[i];
    r.pos = pos;
    r.q = q;
    r.s = s;
    geom_requests.push_back(r);
    geoms_dirty = true;
}

// Geoms are created when the first step needs them: only then is every frame's centre of mass final
// (Frame::addObjectMass re-centres the body each time an object is added), and a geom's local pose is the
// object's pose in the frame minus that accumulated offset.
void WorldPhysics::flushGeoms()
{
    if (!geoms_dirty) return;
    for (const GeomRequest &r : geom_requests) {
        const int body = r.frame->getBody();
        if (body < 0) continue;
        const Vector off = r.frame->getComOffset();
        const double lpos[3] = {r.pos.x() - off.x(), r.pos.y() - off.y(), r.pos.z() - off.z()};
        const double lq[4] = {r.q.w(), r.q.x(), r.q.y(), r.q.z()};
        checkB2P(b2p_geom_create(arena, r.cls, body, r.dims, lpos, lq, &r.s), "WorldPhysics::flushGeoms");
    }
    geom_requests.clear();
    geoms_dirty = false;
}

std::vector<ContactData> WorldPhysics::getContacts(int world)
{
    const MutexLocker locker{&iMutex};
    std::vector<ContactData> out;
    if (!world_init) return out;
    int n = 0;
    b2p_contact_count(arena, world, &n);
    std::map<int, std::shared_ptr<Frame>> by_body;
    for (auto &pair : frameMap)
        if (auto f = pair.second.lock())
            if (f->getBody() >= 0) by_body[f->getBody()] = f;
    for (int c = 0; c < n; c++) {
        int b1 = -1, b2 = -1;
        b2p_contact_desc d{};
        if (b2p_contact_get(arena, world, c, &b1, &b2, &d) < 0) break;
        ContactData cd;
        if (by_body.count(b1)) {
            cd.body1 = by_body[b1];
            cd.nameObject1 = by_body[b1]->getName();
        }
        if (b2 >= 0 && by_body.count(b2)) {
            cd.body2 = by_body[b2];
            cd.nameObject2 = by_body[b2]->getName();
        }
        cd.pos = Vector(d.pos[0], d.pos[1], d.pos[2]);
        cd.normal = Vector(d.normal[0], d.normal[1], d.normal[2]);
        cd.depth = d.depth;
        cd.c_params.friction1 = d.mu;
        cd.c_params.friction2 = d.mu2;
        cd.c_params.rolling_friction = d.rho;
        cd.c_params.rolling_friction2 = d.rho2;
        cd.c_params.spinning_friction = d.rhoN;
        cd.c_params.cfm = d.soft_cfm;
        cd.c_params.erp = d.soft_erp;
        out.push_back(cd);
    }
    return out;
}

// ---------------- batched I/O ----------------
void WorldPhysics::noteJoint(int joint_id, int joint_type)
{
    if (joint_id < 0) return;
    if ((int)joint_types_by_id.size() <= joint_id) joint_types_by_id.resize(joint_id + 1, JOINT_TYPE_UNDEFINED);
    joint_types_by_id[joint_id] = joint_type;
}

void WorldPhysics::setJointVelocitiesBatch(const float *values, int num_joints, int axis)
{
    const MutexLocker locker{&iMutex};
    bool negate = false;
    for (int j = 0; j < num_joints && j < (int)joint_types_by_id.size(); j++)
        negate |= axis == 1 && (joint_types_by_id[j] == JOINT_TYPE_HINGE || joint_types_by_id[j] == JOINT_TYPE_SLIDER);
    const int param = B2P_PARAM_VEL | (axis == 2 ? B2P_PARAM_GROUP2 : 0);
    if (!negate) { // no sign change on the way: one transfer straight from the caller's (pinned) buffer
        checkB2P(b2p_joints_set_param_batch(arena, param, values, num_joints), "WorldPhysics::setJointVelocitiesBatch");
        return;
    }
    batch_tmp.assign(values, values + (size_t)num_joints * num_worlds);
    for (int j = 0; j < num_joints; j++)
        if (joint_types_by_id[j] == JOINT_TYPE_HINGE || joint_types_by_id[j] == JOINT_TYPE_SLIDER)
            for (int w = 0; w < num_worlds; w++) batch_tmp[(size_t)j * num_worlds + w] = -batch_tmp[(size_t)j * num_worlds + w];
    checkB2P(b2p_joints_set_param_batch(arena, param, batch_tmp.data(), num_joints), "WorldPhysics::setJointVelocitiesBatch");
    b2p_sync(arena); // batch_tmp must outlive the asynchronous copy
}

int WorldPhysics::observeBatchBegin(const std::shared_ptr<DynamicObject> &frame, float *body13, float *joint_obs)
{
    const MutexLocker locker{&iMutex};
    const Frame *f = dynamic_cast<const Frame *>(frame.get());
    if (!f || f->getBody() < 0) throw std::runtime_error("WorldPhysics::observeBatchBegin: frame has no body");
    const int t = b2p_observe_begin(arena, f->getBody(), body13, joint_obs);
    checkB2P(t, "WorldPhysics::observeBatchBegin");
    return t;
}

void WorldPhysics::observeBatchWait(int ticket) { checkB2P(b2p_io_wait(arena, ticket), "WorldPhysics::observeBatchWait"); }

const Vector WorldPhysics::getCenterOfMass(const std::vector<std::shared_ptr<NodeInterface>> &) const { return Vector(); }

std::shared_ptr<DynamicObject> WorldPhysics::createFrame(data_broker::DataBrokerInterface *dataBroker,
                                                         configmaps::ConfigMap &config)
{
    auto frame = std::make_shared<Frame>(this, dataBroker, config);
    addFrame(frame);
    return frame;
}

void WorldPhysics::addFrame(std::shared_ptr<Frame> frame)
{
    const std::string &name = frame->getName();
    if (frameMap.find(name) == frameMap.end() || frameMap[name].expired()) frameMap[name] = frame;
}

void WorldPhysics::removeFrame(const std::string &name) { frameMap.erase(name); }

std::shared_ptr<Frame> WorldPhysics::getFrameIntern(const std::string &name)
{
    const auto it = frameMap.find(name);
    return it != frameMap.end() ? it->second.lock() : nullptr;
}

std::shared_ptr<DynamicObject> WorldPhysics::getFrame(const std::string &name) { return getFrameIntern(name); }

// error behaviour of reference :575-603 (schema validation itself is out of scope)
void WorldPhysics::createObject(configmaps::ConfigMap &config)
{
    const std::string type = config["type"].toString();
    if (!ObjectFactory::Instance().hasObjectType(type)) {
        const std::string msg = "WorldPhysics::createObject: could not create object " + type + ": missing schema file.";
        fprintf(stderr, "%s\n", msg.c_str());
        throw std::runtime_error{msg};
    }
    const auto frame = getFrameIntern(config["parentFrame"].toString());
    if (frame == nullptr) {
        const std::string msg = "WorldPhysics::createObject: could not create object " + type +
                                ": missing frame: " + config["parentFrame"].toString() + ".";
        fprintf(stderr, "%s\n", msg.c_str());
        throw std::runtime_error{msg};
    }
    Object *o = ObjectFactory::Instance().createObject(frame, type, config);
    if (o) objects.push_back(o);
}

std::shared_ptr<JointInterface> WorldPhysics::createJoint(data_broker::DataBrokerInterface *dataBroker,
                                                          configmaps::ConfigMap &config)
{
    const std::string name = config["name"].toString();
    const std::string type = config["type"].toString();
    if (!JointFactory::Instance().hasJointType(type)) {
        const std::string msg = "WorldPhysics::createJoint: could not create joint " + type + ": missing schema file.";
        fprintf(stderr, "%s\n", msg.c_str());
        throw std::runtime_error{msg};
    }
    auto newJoint = JointFactory::Instance().createJoint(this, dataBroker, config);
    jointMap[name] = newJoint;
    return newJoint;
}

void WorldPhysics::destroyJoint(const std::string &jointName) { jointMap.erase(jointName); }

std::shared_ptr<JointInterface> WorldPhysics::getJoint(std::string name)
{
    const auto it = jointMap.find(name);
    return it != jointMap.end() ? it->second.lock() : nullptr;
}

void WorldPhysics::clearObjects()
{
    const MutexLocker locker{&iMutex};
    for (auto o : objects) delete o;
    objects.clear();
}

} // namespace ode_physics
} // namespace mars

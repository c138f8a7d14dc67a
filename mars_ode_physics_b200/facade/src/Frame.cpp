// Frame.cpp -- see Frame.hpp.  Semantics follow reference src/objects/Frame.cpp (cited per function).
#include "Frame.hpp"

#include <algorithm>
#include <cstdio>

#include "Objects.hpp"

#define EPSILON 1e-10

namespace mars {
namespace ode_physics {

using namespace utils;
using namespace interfaces;

// reference :40-58: the pose is NOT read from the config
Frame::Frame(PhysicsInterface *world, data_broker::DataBrokerInterface *, configmaps::ConfigMap &config)
    : position(0.0, 0.0, 0.0), q(1.0, 0.0, 0.0, 0.0), ground_contact{false}
{
    theWorld = dynamic_cast<WorldPhysics *>(world);
    nBody = -1;
    contactForce = 0;
    name << config["name"];
    linearDamping = 1.1;
    angularDamping = 1.1;
    if (config.hasKey("linearDamping")) {
        linearDamping = config["linearDamping"];
        angularDamping = config["angularDamping"];
    }
    world_index = 0;
    post_step = -1;
}

Frame::~Frame(void) { clearContactData(); } // bodies live as long as the arena

const std::string &Frame::getName() const { return name; }

bool Frame::body_pose(double pos[3], double quat[4]) const
{
    if (nBody < 0) return false;
    b2p_body_get_position(theWorld->getWorld(), nBody, world_index, pos);
    b2p_body_get_quaternion(theWorld->getWorld(), nBody, world_index, quat);
    return true;
}

void Frame::rotated_offset(const double quat[4], double out[3]) const
{
    double R[9];
    quatToR(quat[0], quat[1], quat[2], quat[3], R);
    const double o[3] = {offsetPos.x(), offsetPos.y(), offsetPos.z()};
    for (int i = 0; i < 3; i++) out[i] = R[3 * i] * o[0] + R[3 * i + 1] * o[1] + R[3 * i + 2] * o[2];
}

// reference :159-171.  The sum over the contact joints created through this frame (the ones whose first body
// is this frame's body) is formed on the device for every world by the step's epilogue; the host only fetches
// the active world's three numbers, and only when somebody asks for them.
void Frame::computeContactForce() { post_step = -1; }
void Frame::fetchPostStep() const
{
    if (nBody < 0 || post_step == theWorld->getStepCount()) return;
    double f[3], la[3], aa[3];
    b2p_arena *a = theWorld->getWorld();
    b2p_body_get_contact_force(a, nBody, world_index, f);
    b2p_body_get_acceleration(a, nBody, world_index, la, aa);
    contactForceVector = Vector(f[0], f[1], f[2]);
    contactForce = contactForceVector.norm();
    l_acc = Vector(la[0], la[1], la[2]);
    a_acc = Vector(aa[0], aa[1], aa[2]);
    post_step = theWorld->getStepCount();
}
const Vector &Frame::getContactForceVector() const
{
    fetchPostStep();
    return contactForceVector;
}
const sReal &Frame::getContactForce() const
{
    fetchPostStep();
    return contactForce;
}

// reference :183-207
void Frame::addObject(Object *object)
{
    if (std::find(objects.begin(), objects.end(), object) != objects.end()) return;
    objects.push_back(object);
    if (nBody < 0) {
        b2p_arena *a = theWorld->getWorld();
        nBody = b2p_body_create(a);
        checkB2P(nBody, "Frame::addObject");
        b2p_body_set_position(a, nBody, B2P_ALL_WORLDS, position.x(), position.y(), position.z());
        b2p_body_set_quaternion(a, nBody, B2P_ALL_WORLDS, q.w(), q.x(), q.y(), q.z());
        b2p_body_set_damping(a, nBody, linearDamping, angularDamping); // applied by the step's epilogue (:922-941)
    }
    addObjectMass(object);
}

void Frame::removeObject(Object *object)
{
    MutexLocker locker(&(theWorld->iMutex));
    objects.erase(std::remove(objects.begin(), objects.end(), object), objects.end());
}

// reference :233-258: frame position = body position - R * offsetPos
void Frame::getPosition(Vector *pos) const
{
    MutexLocker locker(&(theWorld->iMutex));
    double bp[3], bq[4], t[3];
    if (body_pose(bp, bq)) {
        rotated_offset(bq, t);
        *pos = Vector(bp[0] - t[0], bp[1] - t[1], bp[2] - t[2]);
    } else {
        *pos = position;
    }
}

// reference :281-313
void Frame::setPosition(const Vector &pos)
{
    MutexLocker locker(&(theWorld->iMutex));
    position = pos;
    double bp[3], bq[4], t[3];
    if (body_pose(bp, bq)) {
        rotated_offset(bq, t);
        b2p_body_set_position(theWorld->getWorld(), nBody, world_index, pos.x() + t[0], pos.y() + t[1], pos.z() + t[2]);
    }
}

void Frame::getRotation(Quaternion *out) const
{
    MutexLocker locker(&(theWorld->iMutex));
    double bp[3], bq[4];
    if (body_pose(bp, bq))
        *out = Quaternion(bq[0], bq[1], bq[2], bq[3]);
    else
        *out = this->q;
}

void Frame::getPose(Vector *p, Quaternion *r) const
{
    getPosition(p);
    getRotation(r);
}

// reference :349-367 (add_getsetPose.patch): rotation first, then position
void Frame::setPose(const Vector &p, const Quaternion &r, const bool reset_velocities)
{
    setRotation(r);
    setPosition(p);
    if (reset_velocities) {
        setLinearVelocity(Vector(0, 0, 0));
        setAngularVelocity(Vector(0, 0, 0));
    }
}

// reference :519-582: keep the frame origin fixed while the body (COM) swings around it
void Frame::setRotation(const Quaternion &nq)
{
    MutexLocker locker(&(theWorld->iMutex));
    this->q = nq;
    double bp[3], bq[4], told[3], tnew[3];
    if (body_pose(bp, bq)) {
        rotated_offset(bq, told);
        const double nqa[4] = {nq.w(), nq.x(), nq.y(), nq.z()};
        rotated_offset(nqa, tnew);
        b2p_arena *a = theWorld->getWorld();
        b2p_body_set_position(a, nBody, world_index, bp[0] - told[0] + tnew[0], bp[1] - told[1] + tnew[1],
                              bp[2] - told[2] + tnew[2]);
        b2p_body_set_quaternion(a, nBody, world_index, nq.w(), nq.x(), nq.y(), nq.z());
    }
}

static void get3(int (*fn)(b2p_arena *, int, int, double *), b2p_arena *a, int body, int w, Vector *out)
{
    double v[3] = {0, 0, 0};
    if (body >= 0) fn(a, body, w, v);
    *out = Vector(v[0], v[1], v[2]);
}
void Frame::getLinearVelocity(Vector *v) const
{
    MutexLocker locker(&(theWorld->iMutex));
    get3(b2p_body_get_linear_vel, theWorld->getWorld(), nBody, world_index, v);
}
void Frame::getAngularVelocity(Vector *v) const
{
    MutexLocker locker(&(theWorld->iMutex));
    get3(b2p_body_get_angular_vel, theWorld->getWorld(), nBody, world_index, v);
}
void Frame::getForce(Vector *v) const
{
    MutexLocker locker(&(theWorld->iMutex));
    get3(b2p_body_get_force, theWorld->getWorld(), nBody, world_index, v);
}
void Frame::getTorque(Vector *v) const
{
    MutexLocker locker(&(theWorld->iMutex));
    get3(b2p_body_get_torque, theWorld->getWorld(), nBody, world_index, v);
}
void Frame::getLinearAcceleration(Vector *v) const
{
    MutexLocker locker(&(theWorld->iMutex));
    fetchPostStep();
    *v = l_acc;
}
void Frame::getAngularAcceleration(Vector *v) const
{
    MutexLocker locker(&(theWorld->iMutex));
    fetchPostStep();
    *v = a_acc;
}

int Frame::getBody() const { return nBody; }

// reference :604-622: an empty frame that a joint needs gets a 0.01 kg, r = 0.01 m sphere body
int Frame::acquireBody()
{
    if (nBody < 0) {
        fprintf(stderr, "Body of empty Frame (%s) is required -> creating a small \"fake\" body (sphere r=0.01m, 0.01kg)\n",
                name.c_str());
        b2p_arena *a = theWorld->getWorld();
        nBody = b2p_body_create(a);
        checkB2P(nBody, "Frame::acquireBody");
        b2p_body_set_position(a, nBody, B2P_ALL_WORLDS, position.x(), position.y(), position.z());
        b2p_body_set_quaternion(a, nBody, B2P_ALL_WORLDS, q.w(), q.x(), q.y(), q.z());
        b2p_body_set_damping(a, nBody, linearDamping, angularDamping);
        nMass.setSphereTotal(0.01, 0.01);
        b2p_body_set_mass(a, nBody, nMass.mass, nMass.I);
    }
    return nBody;
}

// reference :633-668
const Vector Frame::rotateAtPoint(const Vector &rotation_point, const Quaternion &rotation, bool)
{
    MutexLocker locker(&(theWorld->iMutex));
    double bp[3], bq[4];
    if (!body_pose(bp, bq)) return Vector();
    const double r[4] = {rotation.w(), rotation.x(), rotation.y(), rotation.z()};
    // q' = rotation * q_body
    const double nq[4] = {r[0] * bq[0] - r[1] * bq[1] - r[2] * bq[2] - r[3] * bq[3],
                          r[0] * bq[1] + r[1] * bq[0] + r[2] * bq[3] - r[3] * bq[2],
                          r[0] * bq[2] + r[2] * bq[0] + r[3] * bq[1] - r[1] * bq[3],
                          r[0] * bq[3] + r[3] * bq[0] + r[1] * bq[2] - r[2] * bq[1]};
    const Vector rel(bp[0] - rotation_point.x(), bp[1] - rotation_point.y(), bp[2] - rotation_point.z());
    const Vector np = rotation * rel + rotation_point;
    b2p_arena *a = theWorld->getWorld();
    b2p_body_set_quaternion(a, nBody, world_index, nq[0], nq[1], nq[2], nq[3]);
    b2p_body_set_position(a, nBody, world_index, np.x(), np.y(), np.z());
    return np;
}

MassProps Frame::getMass(void) const { return nMass; }

void Frame::setLinearVelocity(const Vector &v)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (nBody >= 0) b2p_body_set_linear_vel(theWorld->getWorld(), nBody, world_index, v.x(), v.y(), v.z());
}
void Frame::setAngularVelocity(const Vector &v)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (nBody >= 0) b2p_body_set_angular_vel(theWorld->getWorld(), nBody, world_index, v.x(), v.y(), v.z());
}
void Frame::setForce(const Vector &f)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (nBody >= 0) b2p_body_set_force(theWorld->getWorld(), nBody, world_index, f.x(), f.y(), f.z());
}
void Frame::setTorque(const Vector &t)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (nBody >= 0) b2p_body_set_torque(theWorld->getWorld(), nBody, world_index, t.x(), t.y(), t.z());
}
void Frame::addForce(const Vector &f, const Vector &p)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (nBody >= 0)
        b2p_body_add_force_at_pos(theWorld->getWorld(), nBody, world_index, f.x(), f.y(), f.z(), p.x(), p.y(), p.z());
}
void Frame::addForce(const Vector &f)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (nBody >= 0) b2p_body_add_force(theWorld->getWorld(), nBody, world_index, f.x(), f.y(), f.z());
}
void Frame::addTorque(const Vector &t)
{
    MutexLocker locker(&(theWorld->iMutex));
    if (nBody >= 0) b2p_body_add_torque(theWorld->getWorld(), nBody, world_index, t.x(), t.y(), t.z());
}

// reference :794-851: add the object's mass (rotated/translated into the body frame), accumulate the
// COM shift in offsetPos, move the body to the new COM and recentre the mass
void Frame::addObjectMass(Object *object)
{
    MassProps my = object->getMass();
    Vector pos;
    object->getPosition(&pos);
    Quaternion oq(1.0, 0., 0., 0.);
    object->getRotation(&oq);
    double R[9];
    quatToR(oq.w(), oq.x(), oq.y(), oq.z(), R);
    my.rotate(R);
    my.translate(pos.x(), pos.y(), pos.z());
    nMass.add(my);
    offsetPos += Vector(nMass.c[0], nMass.c[1], nMass.c[2]);
    b2p_arena *a = theWorld->getWorld();
    double bp[3], bq[4];
    b2p_body_get_position(a, nBody, 0, bp);
    b2p_body_get_quaternion(a, nBody, 0, bq);
    quatToR(bq[0], bq[1], bq[2], bq[3], R);
    double t[3];
    for (int i = 0; i < 3; i++) t[i] = R[3 * i] * nMass.c[0] + R[3 * i + 1] * nMass.c[1] + R[3 * i + 2] * nMass.c[2] + bp[i];
    b2p_body_set_position(a, nBody, B2P_ALL_WORLDS, t[0], t[1], t[2]);
    position = Vector(t[0], t[1], t[2]);
    nMass.translate(-nMass.c[0], -nMass.c[1], -nMass.c[2]);
    checkB2P(b2p_body_set_mass(a, nBody, nMass.mass, nMass.I), "Frame::addObjectMass");
}

void Frame::getMass(sReal *mass, sReal *inertia) const
{
    if (mass) *mass = nMass.mass;
    if (inertia)
        for (int i = 0; i < 9; i++) inertia[i] = nMass.I[i];
}

// reference :899-943: finite-difference accelerations and multiplicative damping.  Both are done by the step's
// device epilogue for every world (b2p_integrate.cu, with this frame's linearDamping / angularDamping); nothing is
// left to do on the host but to drop the cached values of the previous step.
void Frame::update() { post_step = -1; }

void Frame::clearContactData()
{
    ground_contact = false;
    contactIds.clear();
    contactForceVector = Vector(0, 0, 0);
    contactForce = 0.0;
}

// reference :996-1092: ContactData -> the dContact subset; SoftERP | SoftCFM | Approx1 always,
// Rolling if rolling_friction > EPSILON, Mu2 if mu != mu2; the normal is inverted when `this`
// is body2 of the contact
void Frame::addContact(const ContactData &contact)
{
    int nBody2 = -1;
    double invert = 1.0;
    if (contact.body2) {
        const Frame *frame2 = static_cast<const Frame *>(contact.body2.get());
        if (frame2 == this) {
            frame2 = static_cast<const Frame *>(contact.body1.get());
            invert = -1.0;
        }
        if (frame2 && theWorld->getFrame(frame2->getName())) nBody2 = frame2->getBody();
    }
    b2p_contact_desc c{};
    c.pos[0] = contact.pos.x();
    c.pos[1] = contact.pos.y();
    c.pos[2] = contact.pos.z();
    c.normal[0] = contact.normal.x() * invert;
    c.normal[1] = contact.normal.y() * invert;
    c.normal[2] = contact.normal.z() * invert;
    c.depth = contact.depth;
    c.mode = B2P_CONTACT_SOFT_ERP | B2P_CONTACT_SOFT_CFM | B2P_CONTACT_APPROX1;
    c.mu = contact.c_params.friction1;
    c.mu2 = contact.c_params.friction2;
    c.rhoN = 0.0;
    if (contact.c_params.rolling_friction > EPSILON) {
        c.mode |= B2P_CONTACT_ROLLING;
        c.rho = contact.c_params.rolling_friction;
        c.rho2 = contact.c_params.rolling_friction;
    }
    if (contact.c_params.rolling_friction2 > EPSILON) c.rho2 = contact.c_params.rolling_friction2;
    if (contact.c_params.spinning_friction > EPSILON) c.rhoN = contact.c_params.spinning_friction;
    c.soft_cfm = contact.c_params.cfm;
    c.soft_erp = contact.c_params.erp;
    // bounce / motion / slip values are copied by the reference but their mode bits are never set
    if (c.mu != c.mu2) c.mode |= B2P_CONTACT_MU2;
    const int id = theWorld->createContact(c, nBody, nBody2, world_index);
    if (id == -2) return; // capacity: already reported
    if (id < 0) {
        fprintf(stderr, "Failed to create contact for (%s) because there was already a joint between the two bodies\n",
                getName().c_str());
        return;
    }
    contactIds.push_back(id);
    ground_contact = true;
}

void Frame::addLinkedFrame(std::shared_ptr<DynamicObject> linked) { connectedFrames.push_back(linked); }
bool Frame::isLinkedFrame(std::shared_ptr<DynamicObject> linked)
{
    for (auto &f : connectedFrames)
        if (f.lock() == linked) return true;
    return false;
}
std::vector<std::shared_ptr<DynamicObject>> Frame::getLinkedFrames(void)
{
    std::vector<std::shared_ptr<DynamicObject>> r;
    for (const auto &f : connectedFrames)
        if (const auto v = f.lock()) r.push_back(v);
    return r;
}

} // namespace ode_physics
} // namespace mars

/*
 * oracle/orc_mars.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 * See orc_mars.h.  Each function cites the reference lines it restates.
 * PARITY UNPINNED (see orc.h).
 */
#include "orc_mars.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define EPSILON 1e-10

typedef struct {
    int body; /* orc body id or -1 */
    double position[3], q[4]; /* q = (w,x,y,z) */
    double offset[3];         /* offsetPos: accumulated COM shift in body frame */
    orc_mass mass;
    double linear_damping, angular_damping;
    double l_acc[3], a_acc[3], last_l_vel[3], last_a_vel[3];
    double contact_force[3];
    int contacts[64];
    int ncontacts;
    int ground_contact;
} mframe_t;

typedef struct {
    int type, ode, parent, child, body1, body2;
    double cfm, cfm1, erp1, lo1, hi1, lo2, hi2, damping, spring, joint_cfm, friction;
    double motor_torque, axis1_torque[3], joint_load[3];
} mjoint_t;

struct orc_mars {
    orc_world *w;
    int world_init;
    int fast_step;
    double world_cfm, world_erp, gravity[3], step_size, max_angular_speed, max_correcting_vel;
    double old_cfm, old_erp, old_gravity[3];
    mframe_t *frames;
    int nf, capf;
    mjoint_t *joints;
    int nj, capj;
};

/* WorldPhysics::WorldPhysics, src/WorldPhysics.cpp:80-99 */
orc_mars *orc_mars_create(void)
{
    orc_mars *m = (orc_mars *)calloc(1, sizeof(*m));
    m->fast_step = 0;
    m->world_cfm = 1e-10;
    m->world_erp = 0.1;
    m->gravity[2] = -9.81;
    m->max_angular_speed = 10.0;
    m->max_correcting_vel = 5.0;
    m->step_size = 0.01;
    return m;
}

void orc_mars_destroy(orc_mars *m)
{
    if (!m) return;
    orc_world_destroy(m->w);
    free(m->frames);
    free(m->joints);
    free(m);
}

/* WorldPhysics::initTheWorld, src/WorldPhysics.cpp:162-203 */
void orc_mars_init_the_world(orc_mars *m)
{
    if (m->world_init) return;
    m->w = orc_world_create();
    orc_world_rand_set_seed(m->w, 11211u);
    memcpy(m->old_gravity, m->gravity, sizeof(m->gravity));
    m->old_cfm = m->world_cfm;
    m->old_erp = m->world_erp;
    orc_world_set_gravity(m->w, m->gravity[0], m->gravity[1], m->gravity[2]);
    orc_world_set_cfm(m->w, m->world_cfm);
    orc_world_set_erp(m->w, m->world_erp);
    orc_world_set_max_angular_speed(m->w, m->max_angular_speed);
    orc_world_set_contact_max_correcting_vel(m->w, m->max_correcting_vel);
    m->world_init = 1;
}

orc_world *orc_mars_world(orc_mars *m) { return m->w; }
void orc_mars_set_step_size(orc_mars *m, double h) { m->step_size = h; }
double orc_mars_get_step_size(const orc_mars *m) { return m->step_size; }

void orc_mars_set_world_params(orc_mars *m, const double gravity[3], double cfm, double erp)
{
    memcpy(m->gravity, gravity, sizeof(m->gravity));
    m->world_cfm = cfm;
    m->world_erp = erp;
}

/* ------------------------------------------------------------------ */
/* Frame                                                               */

/* Frame::Frame, src/objects/Frame.cpp:40-58: pose is NOT read from the config */
int orc_mars_create_frame(orc_mars *m, double linear_damping, double angular_damping)
{
    if (m->nf == m->capf) {
        m->capf = m->capf ? 2 * m->capf : 16;
        m->frames = (mframe_t *)realloc(m->frames, sizeof(mframe_t) * (size_t)m->capf);
    }
    mframe_t *f = &m->frames[m->nf];
    memset(f, 0, sizeof(*f));
    f->body = -1;
    f->q[0] = 1.0;
    f->linear_damping = linear_damping;
    f->angular_damping = angular_damping;
    return m->nf++;
}

static void frame_make_body(orc_mars *m, mframe_t *f)
{
    f->body = orc_body_create(m->w);
    orc_body_set_position(m->w, f->body, f->position[0], f->position[1], f->position[2]);
    orc_body_set_quaternion(m->w, f->body, f->q[0], f->q[1], f->q[2], f->q[3]);
}

static void rotate_by_body(const orc_mars *m, int body, const double v[3], double out[3])
{
    double R[9];
    orc_body_get_rotation(m->w, body, R);
    for (int i = 0; i < 3; i++) out[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}

/* Frame::addObject :183-207 and Frame::addObjectMass :794-851 */
void orc_mars_frame_add_object(orc_mars *m, int fi, const orc_mass *mass, const double pos[3], const double q[4])
{
    mframe_t *f = &m->frames[fi];
    if (f->body < 0) frame_make_body(m, f);
    orc_mass my = *mass;
    double R[9];
    orc_q_to_R(q, R);
    orc_mass_rotate(&my, R);
    orc_mass_translate(&my, pos[0], pos[1], pos[2]);
    orc_mass_add(&f->mass, &my);
    for (int i = 0; i < 3; i++) f->offset[i] += f->mass.c[i];
    /* move the body to the new centre of mass, clear the COM of the mass */
    double t[3], bpos[3];
    rotate_by_body(m, f->body, f->mass.c, t);
    orc_body_get_position(m->w, f->body, bpos);
    for (int i = 0; i < 3; i++) t[i] += bpos[i];
    orc_body_set_position(m->w, f->body, t[0], t[1], t[2]);
    memcpy(f->position, t, sizeof(t));
    orc_mass_translate(&f->mass, -f->mass.c[0], -f->mass.c[1], -f->mass.c[2]);
    orc_body_set_mass(m->w, f->body, f->mass.mass, f->mass.I);
}

int orc_mars_frame_body(const orc_mars *m, int f) { return m->frames[f].body; }

/* Frame::acquireBody :604-622 */
int orc_mars_frame_acquire_body(orc_mars *m, int fi)
{
    mframe_t *f = &m->frames[fi];
    if (f->body < 0) {
        frame_make_body(m, f);
        orc_mass_set_sphere_total(&f->mass, 0.01, 0.01);
        orc_body_set_mass(m->w, f->body, f->mass.mass, f->mass.I);
    }
    return f->body;
}

/* Frame::getPosition :233-258 */
void orc_mars_frame_get_position(const orc_mars *m, int fi, double out[3])
{
    const mframe_t *f = &m->frames[fi];
    if (f->body >= 0) {
        double t[3], b[3];
        rotate_by_body(m, f->body, f->offset, t);
        orc_body_get_position(m->w, f->body, b);
        for (int i = 0; i < 3; i++) out[i] = -(t[i] - b[i]);
    } else {
        memcpy(out, f->position, 3 * sizeof(double));
    }
}

/* Frame::setPosition :281-313 */
void orc_mars_frame_set_position(orc_mars *m, int fi, const double p[3])
{
    mframe_t *f = &m->frames[fi];
    memcpy(f->position, p, 3 * sizeof(double));
    if (f->body >= 0) {
        double t[3];
        rotate_by_body(m, f->body, f->offset, t);
        orc_body_set_position(m->w, f->body, p[0] + t[0], p[1] + t[1], p[2] + t[2]);
    }
}

/* Frame::getRotation :331-347 */
void orc_mars_frame_get_rotation(const orc_mars *m, int fi, double out[4])
{
    const mframe_t *f = &m->frames[fi];
    if (f->body >= 0)
        orc_body_get_quaternion(m->w, f->body, out);
    else
        memcpy(out, f->q, 4 * sizeof(double));
}

/* Frame::setRotation :519-582 */
void orc_mars_frame_set_rotation(orc_mars *m, int fi, const double q[4])
{
    mframe_t *f = &m->frames[fi];
    memcpy(f->q, q, 4 * sizeof(double));
    if (f->body >= 0) {
        double t[3], b[3], np[3], R[9];
        rotate_by_body(m, f->body, f->offset, t);
        orc_body_get_position(m->w, f->body, b);
        for (int i = 0; i < 3; i++) np[i] = b[i] - t[i];
        orc_q_to_R(q, R);
        for (int i = 0; i < 3; i++)
            np[i] += R[3 * i] * f->offset[0] + R[3 * i + 1] * f->offset[1] + R[3 * i + 2] * f->offset[2];
        orc_body_set_position(m->w, f->body, np[0], np[1], np[2]);
        orc_body_set_quaternion(m->w, f->body, q[0], q[1], q[2], q[3]);
    }
}

void orc_mars_frame_get_linear_acceleration(const orc_mars *m, int f, double out[3])
{
    memcpy(out, m->frames[f].l_acc, 3 * sizeof(double));
}

void orc_mars_frame_get_contact_force(const orc_mars *m, int f, double out[3])
{
    memcpy(out, m->frames[f].contact_force, 3 * sizeof(double));
}

void orc_mars_frame_get_mass(const orc_mars *m, int f, double *mass, double I[9])
{
    if (mass) *mass = m->frames[f].mass.mass;
    if (I) memcpy(I, m->frames[f].mass.I, 9 * sizeof(double));
}

/* Frame::addContact :996-1092 */
int orc_mars_frame_add_contact(orc_mars *m, int fi, const orc_mars_contact *cd)
{
    mframe_t *f = &m->frames[fi];
    int body2 = -1;
    double invert = 1.0;
    if (cd->frame2 >= 0) {
        int f2 = cd->frame2;
        if (f2 == fi) {
            f2 = cd->frame1;
            invert = -1.0;
        }
        if (f2 >= 0) body2 = m->frames[f2].body;
    }
    orc_contact_desc c;
    memset(&c, 0, sizeof(c));
    for (int i = 0; i < 3; i++) {
        c.pos[i] = cd->pos[i];
        c.normal[i] = cd->normal[i] * invert;
    }
    c.depth = cd->depth;
    c.mode = ORC_CONTACT_SOFT_ERP | ORC_CONTACT_SOFT_CFM | ORC_CONTACT_APPROX1;
    c.mu = cd->friction1;
    c.mu2 = cd->friction2;
    c.rhoN = 0.0;
    if (cd->rolling_friction > EPSILON) {
        c.mode |= ORC_CONTACT_ROLLING;
        c.rho = cd->rolling_friction;
        c.rho2 = cd->rolling_friction;
    }
    if (cd->rolling_friction2 > EPSILON) c.rho2 = cd->rolling_friction2;
    if (cd->spinning_friction > EPSILON) c.rhoN = cd->spinning_friction;
    c.bounce = cd->bounce;
    c.bounce_vel = cd->bounce_vel;
    c.soft_cfm = cd->cfm;
    c.soft_erp = cd->erp;
    c.motion1 = cd->motion1;
    c.motion2 = cd->motion2;
    c.slip1 = cd->fds1;
    c.slip2 = cd->fds2;
    if (c.mu != c.mu2) c.mode |= ORC_CONTACT_MU2;
    int id = orc_contact_add(m->w, f->body, body2, &c);
    if (id < 0) return 0;
    if (f->ncontacts < 64) f->contacts[f->ncontacts++] = id;
    f->ground_contact = 1;
    return 1;
}

/* ------------------------------------------------------------------ */
/* Joint                                                               */

/* Joint::calculateCfmErp, src/joints/Joint.cpp:147-205 */
static void calculate_cfm_erp(const orc_mars *m, mjoint_t *j, const orc_mars_joint_cfg *cfg)
{
    j->damping = -1.0;
    j->spring = -1.0;
    if (cfg->has_spring_damping) j->damping = cfg->spring_damping;
    if (cfg->has_spring_stiffness) j->spring = cfg->spring_stiffness;
    double h = m->step_size;
    j->cfm = j->damping;
    j->erp1 = h * j->spring + j->damping;
    j->cfm1 = h * j->spring + j->damping;
    /* the reference forces lo1 = hi1 = 0 (Joint.cpp:191-193) */
    j->lo1 = 0.0;
    j->hi1 = 0.0;
    j->cfm = (j->cfm > 0) ? 1 / j->cfm : 0.000000000001;
    j->erp1 = (j->erp1 > 0) ? h * j->spring / j->erp1 : 0;
    j->cfm1 = (j->cfm1 > 0) ? 1 / j->cfm1 : 0.00000000001;
}

/* Joint::createJoint :216-273 + per-type createJoint(b1,b2) */
int orc_mars_create_joint(orc_mars *m, const orc_mars_joint_cfg *cfg)
{
    if (!m->world_init) return -1;
    if (cfg->parent < 0 && cfg->child < 0) return -1;
    if (m->nj == m->capj) {
        m->capj = m->capj ? 2 * m->capj : 16;
        m->joints = (mjoint_t *)realloc(m->joints, sizeof(mjoint_t) * (size_t)m->capj);
    }
    mjoint_t *j = &m->joints[m->nj];
    memset(j, 0, sizeof(*j));
    j->type = cfg->type;
    j->parent = cfg->parent;
    j->child = cfg->child;
    j->friction = cfg->has_friction ? cfg->friction_coefficient : -0.1;
    calculate_cfm_erp(m, j, cfg);
    if (cfg->has_joint_cfm) j->joint_cfm = cfg->joint_cfm;
    j->body1 = cfg->parent >= 0 ? orc_mars_frame_acquire_body(m, cfg->parent) : -1;
    j->body2 = cfg->child >= 0 ? orc_mars_frame_acquire_body(m, cfg->child) : -1;
    orc_world *w = m->w;
    switch (cfg->type) {
    case ORC_JOINT_HINGE: /* HingeJoint.cpp:31-66 */
        j->ode = orc_joint_create(w, ORC_JOINT_HINGE, j->body1, j->body2);
        orc_joint_set_anchor(w, j->ode, cfg->anchor[0], cfg->anchor[1], cfg->anchor[2]);
        orc_joint_set_axis1(w, j->ode, cfg->axis1[0], cfg->axis1[1], cfg->axis1[2]);
        break;
    case ORC_JOINT_SLIDER: /* SliderJoint.cpp:32-65 */
        j->ode = orc_joint_create(w, ORC_JOINT_SLIDER, j->body1, j->body2);
        orc_joint_set_axis1(w, j->ode, cfg->axis1[0], cfg->axis1[1], cfg->axis1[2]);
        break;
    case ORC_JOINT_FIXED: /* FixedJoint.cpp:31-47 */
        j->ode = orc_joint_create(w, ORC_JOINT_FIXED, j->body1, j->body2);
        orc_joint_set_fixed(w, j->ode);
        break;
    case ORC_JOINT_HINGE2: /* Hinge2Joint.cpp:33-77 */
        if (j->body1 < 0 || j->body2 < 0) return -1; /* ODE hinge2 needs two bodies */
        j->ode = orc_joint_create(w, ORC_JOINT_HINGE2, j->body1, j->body2);
        orc_joint_set_anchor(w, j->ode, cfg->anchor[0], cfg->anchor[1], cfg->anchor[2]);
        orc_joint_set_axis1(w, j->ode, cfg->axis1[0], cfg->axis1[1], cfg->axis1[2]);
        orc_joint_set_axis2(w, j->ode, cfg->axis2[0], cfg->axis2[1], cfg->axis2[2]);
        break;
    default: return -1;
    }
    if (cfg->type != ORC_JOINT_FIXED) {
        if (j->damping > 0.00000001) {
            orc_joint_set_param(w, j->ode, ORC_PARAM_FMAX, j->damping);
            orc_joint_set_param(w, j->ode, ORC_PARAM_VEL, 0);
            if (cfg->type == ORC_JOINT_HINGE2) {
                orc_joint_set_param(w, j->ode, ORC_PARAM_FMAX | ORC_PARAM_GROUP2, j->damping);
                orc_joint_set_param(w, j->ode, ORC_PARAM_VEL | ORC_PARAM_GROUP2, 0);
            }
        }
        if (j->spring > 0.00000001) {
            /* hinge2 writes lo2/hi2 and cfm2/erp2 (all zero) over the axis-1 slots:
             * Hinge2Joint.cpp:55-62 */
            if (cfg->type == ORC_JOINT_HINGE2) {
                orc_joint_set_param(w, j->ode, ORC_PARAM_LO_STOP, j->lo2);
                orc_joint_set_param(w, j->ode, ORC_PARAM_HI_STOP, j->hi2);
                orc_joint_set_param(w, j->ode, ORC_PARAM_STOP_CFM, 0.0);
                orc_joint_set_param(w, j->ode, ORC_PARAM_STOP_ERP, 0.0);
            } else {
                orc_joint_set_param(w, j->ode, ORC_PARAM_LO_STOP, j->lo1);
                orc_joint_set_param(w, j->ode, ORC_PARAM_HI_STOP, j->hi1);
                orc_joint_set_param(w, j->ode, ORC_PARAM_STOP_CFM, j->cfm1);
                orc_joint_set_param(w, j->ode, ORC_PARAM_STOP_ERP, j->erp1);
            }
        }
        if (j->joint_cfm > 0) orc_joint_set_param(w, j->ode, ORC_PARAM_CFM, j->joint_cfm);
    }
    return m->nj++;
}

int orc_mars_joint_ode_id(const orc_mars *m, int j) { return m->joints[j].ode; }

/* Joint::setForceLimit :315-334 */
void orc_mars_joint_set_force_limit(orc_mars *m, int ji, double v)
{
    mjoint_t *j = &m->joints[ji];
    if (j->type == ORC_JOINT_HINGE || j->type == ORC_JOINT_HINGE2 || j->type == ORC_JOINT_SLIDER)
        orc_joint_set_param(m->w, j->ode, ORC_PARAM_FMAX, v);
}
void orc_mars_joint_set_force_limit2(orc_mars *m, int ji, double v)
{
    mjoint_t *j = &m->joints[ji];
    if (j->type == ORC_JOINT_HINGE2) orc_joint_set_param(m->w, j->ode, ORC_PARAM_FMAX | ORC_PARAM_GROUP2, v);
}
/* Joint::setVelocity :355-374 -- hinge and slider negate, hinge2 does not */
void orc_mars_joint_set_velocity(orc_mars *m, int ji, double v)
{
    mjoint_t *j = &m->joints[ji];
    if (j->type == ORC_JOINT_HINGE || j->type == ORC_JOINT_SLIDER)
        orc_joint_set_param(m->w, j->ode, ORC_PARAM_VEL, -v);
    else if (j->type == ORC_JOINT_HINGE2)
        orc_joint_set_param(m->w, j->ode, ORC_PARAM_VEL, v);
}
void orc_mars_joint_set_velocity2(orc_mars *m, int ji, double v)
{
    mjoint_t *j = &m->joints[ji];
    if (j->type == ORC_JOINT_HINGE2) orc_joint_set_param(m->w, j->ode, ORC_PARAM_VEL | ORC_PARAM_GROUP2, v);
}
/* Joint::setJointAsMotor :618-662 */
void orc_mars_joint_set_joint_as_motor(orc_mars *m, int ji, int axis)
{
    mjoint_t *j = &m->joints[ji];
    switch (j->type) {
    case ORC_JOINT_HINGE:
        orc_joint_set_param(m->w, j->ode, ORC_PARAM_LO_STOP, -INFINITY);
        orc_joint_set_param(m->w, j->ode, ORC_PARAM_HI_STOP, INFINITY);
        break;
    case ORC_JOINT_HINGE2:
        if (!j->lo1 && !j->hi1 && axis == 1) {
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_LO_STOP, -INFINITY);
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_HI_STOP, INFINITY);
        }
        if (!j->lo2 && !j->hi2 && axis == 2) {
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_LO_STOP | ORC_PARAM_GROUP2, -INFINITY);
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_HI_STOP | ORC_PARAM_GROUP2, INFINITY);
        }
        break;
    case ORC_JOINT_SLIDER:
        if (!j->lo1 && !j->hi1) {
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_LO_STOP, -INFINITY);
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_HI_STOP, INFINITY);
        }
        break;
    default: break;
    }
}
/* Joint::unsetJointAsMotor :664-703 */
void orc_mars_joint_unset_joint_as_motor(orc_mars *m, int ji, int axis)
{
    mjoint_t *j = &m->joints[ji];
    switch (j->type) {
    case ORC_JOINT_HINGE:
    case ORC_JOINT_SLIDER:
        orc_joint_set_param(m->w, j->ode, ORC_PARAM_LO_STOP, j->lo1);
        orc_joint_set_param(m->w, j->ode, ORC_PARAM_HI_STOP, j->hi1);
        break;
    case ORC_JOINT_HINGE2:
        if (axis == 1) {
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_LO_STOP, j->lo1);
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_HI_STOP, j->hi1);
        } else if (axis == 2) {
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_LO_STOP | ORC_PARAM_GROUP2, j->lo2);
            orc_joint_set_param(m->w, j->ode, ORC_PARAM_HI_STOP | ORC_PARAM_GROUP2, j->hi2);
        }
        break;
    default: break;
    }
}
/* Joint::getPosition :395-415 */
double orc_mars_joint_get_position(const orc_mars *m, int ji)
{
    const mjoint_t *j = &m->joints[ji];
    if (j->type == ORC_JOINT_HINGE || j->type == ORC_JOINT_SLIDER) return -orc_joint_get_angle1(m->w, j->ode);
    if (j->type == ORC_JOINT_HINGE2) return orc_joint_get_angle1(m->w, j->ode);
    return 0;
}
/* Joint::getVelocityI :1078-1102 */
double orc_mars_joint_get_velocity(const orc_mars *m, int ji)
{
    const mjoint_t *j = &m->joints[ji];
    if (j->type == ORC_JOINT_HINGE || j->type == ORC_JOINT_SLIDER) return -orc_joint_get_angle1_rate(m->w, j->ode);
    if (j->type == ORC_JOINT_HINGE2) return orc_joint_get_angle1_rate(m->w, j->ode);
    return 0;
}
double orc_mars_joint_get_velocity2(const orc_mars *m, int ji)
{
    const mjoint_t *j = &m->joints[ji];
    if (j->type == ORC_JOINT_HINGE2) return orc_joint_get_angle2_rate(m->w, j->ode);
    return 0;
}
double orc_mars_joint_get_motor_torque(const orc_mars *m, int j) { return m->joints[j].motor_torque; }
void orc_mars_joint_get_axis_torque(const orc_mars *m, int j, double out[3])
{
    memcpy(out, m->joints[j].axis1_torque, 3 * sizeof(double));
}
void orc_mars_joint_get_joint_load(const orc_mars *m, int j, double out[3])
{
    memcpy(out, m->joints[j].joint_load, 3 * sizeof(double));
}

/* Joint::applyJointFriction :1031-1063 and Joint::update :859-1019 -> orc_post.c (shared with the batch form) */
static void joint_update(orc_mars *m, mjoint_t *j)
{
    double out[7];
    orc_post_joint_update(m->w, j->ode, out);
    j->motor_torque = out[0];
    memcpy(j->axis1_torque, out + 1, 3 * sizeof(double));
    memcpy(j->joint_load, out + 4, 3 * sizeof(double));
    orc_post_joint_friction(m->w, j->ode, j->friction);
}

/* Frame::update :899-943 -> orc_post.c */
static void frame_update(orc_mars *m, mframe_t *f)
{
    if (f->body < 0) return;
    double last[6], acc[6];
    memcpy(last, f->last_l_vel, sizeof(f->last_l_vel));
    memcpy(last + 3, f->last_a_vel, sizeof(f->last_a_vel));
    memcpy(acc, f->l_acc, sizeof(f->l_acc));
    memcpy(acc + 3, f->a_acc, sizeof(f->a_acc));
    orc_post_frame_update(m->w, f->body, m->step_size, f->linear_damping, f->angular_damping, last, acc);
    memcpy(f->last_l_vel, last, sizeof(f->last_l_vel));
    memcpy(f->last_a_vel, last + 3, sizeof(f->last_a_vel));
    memcpy(f->l_acc, acc, sizeof(f->l_acc));
    memcpy(f->a_acc, acc + 3, sizeof(f->a_acc));
}

/* WorldPhysics::clearPreviousStep :314-325 + Frame::clearContactData :945-955 */
void orc_mars_clear_previous_step(orc_mars *m)
{
    orc_contacts_clear(m->w);
    for (int i = 0; i < m->nf; i++) {
        m->frames[i].ncontacts = 0;
        m->frames[i].ground_contact = 0;
        memset(m->frames[i].contact_force, 0, sizeof(m->frames[i].contact_force));
    }
}

/* WorldPhysics::stepTheWorld :349-421 (fast_step path) */
void orc_mars_step_the_world(orc_mars *m)
{
    if (!(m->world_init && m->step_size > 0)) return;
    /* preStepChecks :286-306 */
    if (memcmp(m->old_gravity, m->gravity, sizeof(m->gravity)) != 0) {
        memcpy(m->old_gravity, m->gravity, sizeof(m->gravity));
        orc_world_set_gravity(m->w, m->gravity[0], m->gravity[1], m->gravity[2]);
    }
    if (m->old_cfm != m->world_cfm) {
        m->old_cfm = m->world_cfm;
        orc_world_set_cfm(m->w, m->world_cfm);
    }
    if (m->old_erp != m->world_erp) {
        m->old_erp = m->world_erp;
        orc_world_set_erp(m->w, m->world_erp);
    }
    orc_world_quickstep(m->w, m->step_size);
    for (int i = 0; i < m->nj; i++) joint_update(m, &m->joints[i]);
    for (int i = 0; i < m->nf; i++) frame_update(m, &m->frames[i]);
    /* computeContactForces :327-335 -> Frame::computeContactForce :159-171 */
    for (int i = 0; i < m->nf; i++) {
        mframe_t *f = &m->frames[i];
        double sum[3] = {0, 0, 0}, fb[12];
        for (int k = 0; k < f->ncontacts; k++) {
            orc_contact_get_feedback(m->w, f->contacts[k], fb);
            sum[0] += fb[0];
            sum[1] += fb[1];
            sum[2] += fb[2];
        }
        memcpy(f->contact_force, sum, sizeof(sum));
    }
}

/* ------------------------------------------------------------------ */
/* the scene of test/reproducable_test.cpp                             */

static int scene_box(orc_mars *m, double extend_x, double mass, double px, double py, double pz)
{
    /* createBox :25-48 overrides extend.y/z and (for the 9 links) the position */
    int f = orc_mars_create_frame(m, 1.1, 1.1);
    orc_mass bm;
    orc_mass_set_box_total(&bm, mass, extend_x, 0.2, 0.05);
    double pos[3] = {px, py, pz}, q[4] = {1, 0, 0, 0};
    orc_mars_frame_add_object(m, f, &bm, pos, q);
    return f;
}

void orc_mars_build_reproducable_scene(orc_mars *m, int frames[13], int joints[12])
{
    orc_mars_init_the_world(m);
    m->step_size = 0.001;
    m->fast_step = 1;
    frames[0] = scene_box(m, 0.4, 1.0, 0.0, 0.0, 0.265);
    for (int i = 1; i <= 8; i++) frames[i] = scene_box(m, 0.03, 0.2, 0.0, 0.0, 0.265);
    /* feet: type "sphere" is registered as Box in the test (:54), extend.x = 0.02 */
    const double fx[4] = {0.2 - 0.0015, 0.2 - 0.0015, -(0.2 - 0.0015), -(0.2 - 0.0015)};
    const double fy[4] = {0.1 - 0.0015, -(0.1 - 0.0015), 0.1 - 0.0015, -(0.1 - 0.0015)};
    for (int i = 0; i < 4; i++) frames[9 + i] = scene_box(m, 0.02, 0.05, fx[i], fy[i], 0.02);

    const double ax[4] = {0.2 - 0.015, 0.2 - 0.015, -(0.2 - 0.015), -(0.2 - 0.015)};
    const double ay[4] = {0.1 - 0.015, -(0.1 - 0.015), 0.1 - 0.015, -(0.1 - 0.015)};
    orc_mars_joint_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.axis1[1] = 1.0;
    /* hips: body -> {fl,fr,rl,rr}_upper = frames 1,3,5,7 */
    for (int i = 0; i < 4; i++) {
        cfg.type = ORC_JOINT_HINGE;
        cfg.parent = frames[0];
        cfg.child = frames[1 + 2 * i];
        cfg.anchor[0] = ax[i];
        cfg.anchor[1] = ay[i];
        cfg.anchor[2] = 0.44;
        joints[i] = orc_mars_create_joint(m, &cfg);
    }
    /* knees: upper -> lower */
    for (int i = 0; i < 4; i++) {
        cfg.parent = frames[1 + 2 * i];
        cfg.child = frames[2 + 2 * i];
        cfg.anchor[0] = ax[i];
        cfg.anchor[1] = ay[i];
        cfg.anchor[2] = 0.24;
        joints[4 + i] = orc_mars_create_joint(m, &cfg);
    }
    /* ankles: fixed lower -> foot */
    for (int i = 0; i < 4; i++) {
        cfg.type = ORC_JOINT_FIXED;
        cfg.parent = frames[2 + 2 * i];
        cfg.child = frames[9 + i];
        cfg.anchor[2] = 0.4;
        joints[8 + i] = orc_mars_create_joint(m, &cfg);
    }
    for (int i = 0; i < 12; i++) {
        orc_mars_joint_set_force_limit(m, joints[i], 9);
        orc_mars_joint_set_joint_as_motor(m, joints[i], 1);
    }
}

int orc_mars_reproducable_tick(orc_mars *m, const int frames[13], const int joints[12], double time)
{
    const double offset[4] = {0.0, 1.14, 0.0, 1.14};
    for (int i = 0; i < 4; i++) {
        double hip = 0.5 * sin(offset[i] + time * 4 * M_PI) - 0.3;
        double knee = -0.5 * sin(offset[i] + 1 + time * 4 * M_PI) + 0.3;
        orc_mars_joint_set_velocity(m, joints[i], (hip - orc_mars_joint_get_position(m, joints[i])) * 20);
        orc_mars_joint_set_velocity(m, joints[i + 4], (knee - orc_mars_joint_get_position(m, joints[i + 4])) * 20);
    }
    orc_mars_clear_previous_step(m);
    int nc = 0;
    for (int i = 0; i < 4; i++) {
        double p[3];
        orc_mars_frame_get_position(m, frames[9 + i], p);
        p[2] -= 0.02;
        if (p[2] <= 0.0) {
            orc_mars_contact cd;
            memset(&cd, 0, sizeof(cd));
            cd.frame1 = frames[9 + i];
            cd.frame2 = -1;
            cd.cfm = 0.00001;
            cd.erp = 0.2;
            cd.friction1 = 0.8;
            cd.friction2 = 0.8;
            cd.pos[0] = p[0];
            cd.pos[1] = p[1];
            cd.pos[2] = 0.0;
            cd.depth = p[2];
            cd.normal[2] = 1.0;
            nc += orc_mars_frame_add_contact(m, frames[9 + i], &cd);
        }
    }
    orc_mars_step_the_world(m);
    return nc;
}

int orc_mars_run_reproducable_test(const char *csv_path, double t_end, int order_mode, double *last_pos)
{
    return orc_mars_run_reproducable_test_stepper(csv_path, t_end, order_mode, ORC_STEPPER_QUICK, last_pos);
}

/* the same run with the stepper the reference's fast_step selects (src/WorldPhysics.cpp:365-370) */
int orc_mars_run_reproducable_test_stepper(const char *csv_path, double t_end, int order_mode, int stepper,
                                           double *last_pos)
{
    orc_mars *m = orc_mars_create();
    int frames[13], joints[12];
    orc_mars_build_reproducable_scene(m, frames, joints);
    orc_world_set_order_mode(m->w, order_mode);
    orc_world_set_stepper(m->w, stepper);
    FILE *file = csv_path ? fopen(csv_path, "w") : 0;
    if (file) fprintf(file, "time_sec,body_x,body_y,body_z\n");
    double time = 0.0, pos[3] = {0, 0, 0};
    int rows = 0;
    while (time < t_end) {
        orc_mars_frame_get_position(m, frames[0], pos);
        if (file) fprintf(file, "%g,%g,%g,%g\n", time, pos[0], pos[1], pos[2]);
        rows++;
        orc_mars_reproducable_tick(m, frames, joints, time);
        time += m->step_size;
    }
    if (file) fclose(file);
    if (last_pos) orc_mars_frame_get_position(m, frames[0], last_pos);
    orc_mars_destroy(m);
    return rows;
}

/*
 * oracle/orc_quickstep.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Restatement of ODE 0.16 dWorldQuickStep semantics, the call behind
 * WorldPhysics::stepTheWorld (reference src/WorldPhysics.cpp:365), plus the ODE
 * body / joint / contact / mass entry points the reference calls
 * (src/objects/Frame.cpp, src/joints/ *.cpp).  See orc.h for the provenance note.
 *
 * PARITY UNPINNED: no golden vectors exist in the reference and libode is absent.
 *
 * Build twice: default = double (`real`=double), -DORC_SINGLE = float.
 */
#include "orc.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "orc_internal.h"
#define R_PI ((real)3.14159265358979323846)
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define MAX_ROWS_PER_JOINT 6

/* ------------------------------------------------------------------ */
/* small math                                                          */

static real dot3(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static void cross3(real *r, const real *a, const real *b)
{
    real x = a[1] * b[2] - a[2] * b[1];
    real y = a[2] * b[0] - a[0] * b[2];
    real z = a[0] * b[1] - a[1] * b[0];
    r[0] = x;
    r[1] = y;
    r[2] = z;
}

/* r = R v */
static void mul_R_v(real *r, const real *R, const real *v)
{
    real x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    real y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    real z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    r[0] = x;
    r[1] = y;
    r[2] = z;
}

/* r = R^T v */
static void mul_Rt_v(real *r, const real *R, const real *v)
{
    real x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    real y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    real z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    r[0] = x;
    r[1] = y;
    r[2] = z;
}

/* C = A B (3x3 row-major) */
static void mul33(real *C, const real *A, const real *B)
{
    real t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            t[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
    memcpy(C, t, sizeof(t));
}

/* C = A B^T */
static void mul33_bt(real *C, const real *A, const real *B)
{
    real t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            t[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
    memcpy(C, t, sizeof(t));
}

/* ODE dQtoR / dRfromQ, q = (w,x,y,z) */
static void q_to_R(const real *q, real *R)
{
    real qq1 = 2 * q[1] * q[1];
    real qq2 = 2 * q[2] * q[2];
    real qq3 = 2 * q[3] * q[3];
    R[0] = 1 - qq2 - qq3;
    R[1] = 2 * (q[1] * q[2] - q[0] * q[3]);
    R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
    R[3] = 2 * (q[1] * q[2] + q[0] * q[3]);
    R[4] = 1 - qq1 - qq3;
    R[5] = 2 * (q[2] * q[3] - q[0] * q[1]);
    R[6] = 2 * (q[1] * q[3] - q[0] * q[2]);
    R[7] = 2 * (q[2] * q[3] + q[0] * q[1]);
    R[8] = 1 - qq1 - qq2;
}

/* ODE dQMultiply0..3 */
static void qmul1(real *r, const real *b, const real *c) /* b^-1 * c */
{
    real t[4];
    t[0] = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    t[1] = b[0] * c[1] - b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    t[2] = b[0] * c[2] - b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    t[3] = b[0] * c[3] - b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
    memcpy(r, t, sizeof(t));
}
static void qmul2(real *r, const real *b, const real *c) /* b * c^-1 */
{
    real t[4];
    t[0] = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    t[1] = -b[0] * c[1] + b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    t[2] = -b[0] * c[2] + b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    t[3] = -b[0] * c[3] + b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
    memcpy(r, t, sizeof(t));
}
static void qmul3(real *r, const real *b, const real *c) /* b^-1 * c^-1 */
{
    real t[4];
    t[0] = b[0] * c[0] - b[1] * c[1] - b[2] * c[2] - b[3] * c[3];
    t[1] = -b[0] * c[1] - b[1] * c[0] + b[2] * c[3] - b[3] * c[2];
    t[2] = -b[0] * c[2] - b[2] * c[0] + b[3] * c[1] - b[1] * c[3];
    t[3] = -b[0] * c[3] - b[3] * c[0] + b[1] * c[2] - b[2] * c[1];
    memcpy(r, t, sizeof(t));
}

static void normalize3(real *a)
{
    real l = dot3(a, a);
    if (l > 0) {
        l = 1 / R_SQRT(l);
        a[0] *= l;
        a[1] *= l;
        a[2] *= l;
    } else {
        a[0] = 1;
        a[1] = 0;
        a[2] = 0;
    }
}

static void normalize4(real *a)
{
    real l = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    if (l > 0) {
        l = 1 / R_SQRT(l);
        for (int i = 0; i < 4; i++) a[i] *= l;
    } else {
        a[0] = 1;
        a[1] = a[2] = a[3] = 0;
    }
}

/* ODE dPlaneSpace (odemath.cpp) */
static void plane_space(const real *n, real *p, real *q)
{
    if (R_FABS(n[2]) > (real)0.70710678118654752440) {
        /* choose p in y-z plane */
        real a = n[1] * n[1] + n[2] * n[2];
        real k = 1 / R_SQRT(a);
        p[0] = 0;
        p[1] = -n[2] * k;
        p[2] = n[1] * k;
        /* q = n x p */
        q[0] = a * k;
        q[1] = -n[0] * p[2];
        q[2] = n[0] * p[1];
    } else {
        /* choose p in x-y plane */
        real a = n[0] * n[0] + n[1] * n[1];
        real k = 1 / R_SQRT(a);
        p[0] = -n[1] * k;
        p[1] = n[0] * k;
        p[2] = 0;
        /* q = n x p */
        q[0] = -n[2] * p[1];
        q[1] = n[2] * p[0];
        q[2] = a * k;
    }
}

/* closed-form 3x3 inverse (ODE dInvertMatrix3); returns determinant */
static real invert33(real *dst, const real *m)
{
    real c00 = m[4] * m[8] - m[5] * m[7];
    real c01 = m[5] * m[6] - m[3] * m[8];
    real c02 = m[3] * m[7] - m[4] * m[6];
    real det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    if (det == 0) return 0;
    real r = 1 / det;
    real t[9];
    t[0] = c00 * r;
    t[1] = (m[2] * m[7] - m[1] * m[8]) * r;
    t[2] = (m[1] * m[5] - m[2] * m[4]) * r;
    t[3] = c01 * r;
    t[4] = (m[0] * m[8] - m[2] * m[6]) * r;
    t[5] = (m[2] * m[3] - m[0] * m[5]) * r;
    t[6] = c02 * r;
    t[7] = (m[1] * m[6] - m[0] * m[7]) * r;
    t[8] = (m[0] * m[4] - m[1] * m[3]) * r;
    memcpy(dst, t, sizeof(t));
    return det;
}

/* ------------------------------------------------------------------ */
/* RNG: ODE misc.cpp dRand / dRandInt                                  */

uint32_t orc_rand_next(uint32_t *seed)
{
    *seed = (uint32_t)(1664525u * (*seed) + 1013904223u);
    return *seed;
}

int orc_rand_int(uint32_t *seed, int n)
{
    uint32_t un = (uint32_t)n;
    uint32_t r = orc_rand_next(seed);
    if (un <= 0x00010000u) {
        r ^= (r >> 16);
        if (un <= 0x00000100u) {
            r ^= (r >> 8);
            if (un <= 0x00000010u) {
                r ^= (r >> 4);
                if (un <= 0x00000004u) {
                    r ^= (r >> 2);
                    if (un <= 0x00000002u) {
                        r ^= (r >> 1);
                    }
                }
            }
        }
    }
    return (int)(r % un);
}

int orc_real_size(void) { return (int)sizeof(real); }

static void limot_init(limot_t *l, const orc_world *w)
{
    l->vel = 0;
    l->fmax = 0;
    l->lostop = -R_INF;
    l->histop = R_INF;
    l->fudge_factor = 1;
    l->normal_cfm = w->cfm;
    l->stop_erp = w->erp;
    l->stop_cfm = w->cfm;
    l->bounce = 0;
    l->limit = 0;
    l->limit_err = 0;
}

static void limot_set(limot_t *l, int num, real value)
{
    switch (num) {
    case ORC_PARAM_LO_STOP: l->lostop = value; break;
    case ORC_PARAM_HI_STOP: l->histop = value; break;
    case ORC_PARAM_VEL: l->vel = value; break;
    case ORC_PARAM_FMAX:
        if (value >= 0) l->fmax = value;
        break;
    case ORC_PARAM_FUDGE_FACTOR:
        if (value >= 0 && value <= 1) l->fudge_factor = value;
        break;
    case ORC_PARAM_BOUNCE: l->bounce = value; break;
    case ORC_PARAM_CFM: l->normal_cfm = value; break;
    case ORC_PARAM_STOP_ERP: l->stop_erp = value; break;
    case ORC_PARAM_STOP_CFM: l->stop_cfm = value; break;
    default: break;
    }
}

static real limot_get(const limot_t *l, int num)
{
    switch (num) {
    case ORC_PARAM_LO_STOP: return l->lostop;
    case ORC_PARAM_HI_STOP: return l->histop;
    case ORC_PARAM_VEL: return l->vel;
    case ORC_PARAM_FMAX: return l->fmax;
    case ORC_PARAM_FUDGE_FACTOR: return l->fudge_factor;
    case ORC_PARAM_BOUNCE: return l->bounce;
    case ORC_PARAM_CFM: return l->normal_cfm;
    case ORC_PARAM_STOP_ERP: return l->stop_erp;
    case ORC_PARAM_STOP_CFM: return l->stop_cfm;
    default: return 0;
    }
}

static int limot_test_rotational_limit(limot_t *l, real angle)
{
    if (angle <= l->lostop) {
        l->limit = 1;
        l->limit_err = angle - l->lostop;
        return 1;
    } else if (angle >= l->histop) {
        l->limit = 2;
        l->limit_err = angle - l->histop;
        return 1;
    }
    l->limit = 0;
    return 0;
}

/* ------------------------------------------------------------------ */
/* world                                                               */

orc_world *orc_world_create(void)
{
    orc_world *w = (orc_world *)calloc(1, sizeof(*w));
    /* dWorldCreate defaults */
    w->erp = (real)0.2;
    w->cfm = (real)ORC_DEFAULT_CFM;
    w->max_angular_speed = R_INF;
    w->contact_max_vel = R_INF;
    w->contact_min_depth = 0;
    w->iters = 20;
    w->sor_w = (real)1.3;
    w->order_mode = ORC_ORDER_WORLD_BLOCK;
    w->reorder_method = ORC_REORDER_RANDOMLY;
    w->seed = 0;
    w->max_contacts = 8;
    return w;
}

void orc_world_destroy(orc_world *w)
{
    if (!w) return;
    free(w->bodies);
    free(w->joints);
    free(w->contacts);
    free(w->geoms);
    free(w->hf_heights);
    free(w->pairs);
    free(w);
}

void orc_world_set_gravity(orc_world *w, double x, double y, double z)
{
    w->gravity[0] = (real)x;
    w->gravity[1] = (real)y;
    w->gravity[2] = (real)z;
}
void orc_world_set_erp(orc_world *w, double erp) { w->erp = (real)erp; }
void orc_world_set_cfm(orc_world *w, double cfm) { w->cfm = (real)cfm; }
void orc_world_set_max_angular_speed(orc_world *w, double s) { w->max_angular_speed = (real)s; }
void orc_world_set_contact_max_correcting_vel(orc_world *w, double v) { w->contact_max_vel = (real)v; }
void orc_world_set_contact_surface_layer(orc_world *w, double d) { w->contact_min_depth = (real)d; }
void orc_world_set_quickstep_num_iterations(orc_world *w, int n) { w->iters = n; }
void orc_world_set_quickstep_w(orc_world *w, double sor_w) { w->sor_w = (real)sor_w; }
void orc_world_set_order_mode(orc_world *w, int m) { w->order_mode = m; }
void orc_world_set_reorder_method(orc_world *w, int m) { w->reorder_method = m; }
void orc_world_set_stepper(orc_world *w, int s) { w->stepper = s; }
int orc_world_last_lcp_error(const orc_world *w) { return w->last_lcp_error; }
void orc_world_rand_set_seed(orc_world *w, uint32_t seed) { w->seed = seed; }
uint32_t orc_world_rand_get_seed(const orc_world *w) { return w->seed; }
int orc_world_last_num_rows(const orc_world *w) { return w->last_m; }
int orc_world_last_num_islands(const orc_world *w) { return w->last_islands; }
double orc_world_last_residual(const orc_world *w) { return (double)w->last_residual; }
void orc_world_set_variant(orc_world *w, int which, int value)
{
    if (which >= 0 && which < ORC_NUM_VARIANTS) w->variant[which] = value;
}
int orc_world_num_bodies(const orc_world *w) { return w->nb; }
int orc_world_num_joints(const orc_world *w) { return w->nj; }
int orc_world_num_contacts(const orc_world *w) { return w->nc; }

/* ------------------------------------------------------------------ */
/* bodies                                                              */

int orc_body_create(orc_world *w)
{
    if (w->nb == w->capb) {
        w->capb = w->capb ? 2 * w->capb : 16;
        w->bodies = (body_t *)realloc(w->bodies, sizeof(body_t) * (size_t)w->capb);
    }
    body_t *b = &w->bodies[w->nb];
    memset(b, 0, sizeof(*b));
    b->q[0] = 1;
    q_to_R(b->q, b->R);
    b->mass = 1;
    b->invMass = 1;
    b->I[0] = b->I[4] = b->I[8] = 1;
    b->invI[0] = b->invI[4] = b->invI[8] = 1;
    b->gyroscopic = 1; /* dBodyCreate sets dxBodyGyroscopic */
    b->nogravity = 0;
    /* dWorldSetMaxAngularSpeed before dBodyCreate -> body inherits the flag + value
     * (reference: src/WorldPhysics.cpp:197 is called in initTheWorld before any frame exists) */
    b->has_max_ang = (w->max_angular_speed < R_INF);
    b->max_angular_speed = w->max_angular_speed;
    return w->nb++;
}

void orc_body_set_mass(orc_world *w, int bi, double mass, const double I[9])
{
    body_t *b = &w->bodies[bi];
    b->mass = (real)mass;
    b->invMass = 1 / (real)mass;
    for (int i = 0; i < 9; i++) b->I[i] = (real)I[i];
    invert33(b->invI, b->I);
}

void orc_body_get_mass(const orc_world *w, int bi, double *mass, double I[9])
{
    const body_t *b = &w->bodies[bi];
    if (mass) *mass = b->mass;
    if (I)
        for (int i = 0; i < 9; i++) I[i] = b->I[i];
}

void orc_body_set_position(orc_world *w, int bi, double x, double y, double z)
{
    body_t *b = &w->bodies[bi];
    b->pos[0] = (real)x;
    b->pos[1] = (real)y;
    b->pos[2] = (real)z;
}

void orc_body_set_quaternion(orc_world *w, int bi, double qw, double qx, double qy, double qz)
{
    body_t *b = &w->bodies[bi];
    b->q[0] = (real)qw;
    b->q[1] = (real)qx;
    b->q[2] = (real)qy;
    b->q[3] = (real)qz;
    normalize4(b->q);
    q_to_R(b->q, b->R);
}

#define SET3(dst, x, y, z)  \
    do {                    \
        (dst)[0] = (real)(x); \
        (dst)[1] = (real)(y); \
        (dst)[2] = (real)(z); \
    } while (0)
#define GET3(out, src)            \
    do {                          \
        (out)[0] = (double)(src)[0]; \
        (out)[1] = (double)(src)[1]; \
        (out)[2] = (double)(src)[2]; \
    } while (0)

void orc_body_set_linear_vel(orc_world *w, int b, double x, double y, double z) { SET3(w->bodies[b].lvel, x, y, z); }
void orc_body_set_angular_vel(orc_world *w, int b, double x, double y, double z) { SET3(w->bodies[b].avel, x, y, z); }
void orc_body_set_force(orc_world *w, int b, double x, double y, double z) { SET3(w->bodies[b].facc, x, y, z); }
void orc_body_set_torque(orc_world *w, int b, double x, double y, double z) { SET3(w->bodies[b].tacc, x, y, z); }
void orc_body_add_force(orc_world *w, int b, double x, double y, double z)
{
    w->bodies[b].facc[0] += (real)x;
    w->bodies[b].facc[1] += (real)y;
    w->bodies[b].facc[2] += (real)z;
}
void orc_body_add_torque(orc_world *w, int b, double x, double y, double z)
{
    w->bodies[b].tacc[0] += (real)x;
    w->bodies[b].tacc[1] += (real)y;
    w->bodies[b].tacc[2] += (real)z;
}
void orc_body_add_force_at_pos(orc_world *w, int bi, double fx, double fy, double fz, double px, double py,
                               double pz)
{
    body_t *b = &w->bodies[bi];
    real f[3] = {(real)fx, (real)fy, (real)fz};
    real q[3] = {(real)px - b->pos[0], (real)py - b->pos[1], (real)pz - b->pos[2]};
    real t[3];
    cross3(t, q, f);
    for (int i = 0; i < 3; i++) {
        b->facc[i] += f[i];
        b->tacc[i] += t[i];
    }
}
void orc_body_set_gravity_mode(orc_world *w, int b, int on) { w->bodies[b].nogravity = !on; }
void orc_body_set_gyroscopic_mode(orc_world *w, int b, int on) { w->bodies[b].gyroscopic = on; }
void orc_body_get_position(const orc_world *w, int b, double out[3]) { GET3(out, w->bodies[b].pos); }
void orc_body_get_quaternion(const orc_world *w, int b, double out[4])
{
    for (int i = 0; i < 4; i++) out[i] = w->bodies[b].q[i];
}
void orc_body_get_rotation(const orc_world *w, int b, double out[9])
{
    for (int i = 0; i < 9; i++) out[i] = w->bodies[b].R[i];
}
void orc_body_get_linear_vel(const orc_world *w, int b, double out[3]) { GET3(out, w->bodies[b].lvel); }
void orc_body_get_angular_vel(const orc_world *w, int b, double out[3]) { GET3(out, w->bodies[b].avel); }
void orc_body_get_force(const orc_world *w, int b, double out[3]) { GET3(out, w->bodies[b].facc); }
void orc_body_get_torque(const orc_world *w, int b, double out[3]) { GET3(out, w->bodies[b].tacc); }

void orc_body_get_state(const orc_world *w, int bi, double out[13])
{
    const body_t *b = &w->bodies[bi];
    for (int i = 0; i < 3; i++) out[i] = b->pos[i];
    for (int i = 0; i < 4; i++) out[3 + i] = b->q[i];
    for (int i = 0; i < 3; i++) out[7 + i] = b->lvel[i];
    for (int i = 0; i < 3; i++) out[10 + i] = b->avel[i];
}

void orc_body_set_state(orc_world *w, int bi, const double in[13])
{
    orc_body_set_position(w, bi, in[0], in[1], in[2]);
    orc_body_set_quaternion(w, bi, in[3], in[4], in[5], in[6]);
    orc_body_set_linear_vel(w, bi, in[7], in[8], in[9]);
    orc_body_set_angular_vel(w, bi, in[10], in[11], in[12]);
}

/* ------------------------------------------------------------------ */
/* joints: creation-time geometry (ODE joint.cpp setAnchors/setAxes,    */
/* hinge.cpp, slider.cpp, fixed.cpp, hinge2.cpp)                        */

static joint_t *joint_alloc(joint_t **arr, int *n, int *cap)
{
    if (*n == *cap) {
        *cap = *cap ? 2 * (*cap) : 16;
        *arr = (joint_t *)realloc(*arr, sizeof(joint_t) * (size_t)(*cap));
    }
    joint_t *j = &(*arr)[*n];
    memset(j, 0, sizeof(*j));
    (*n)++;
    return j;
}

/* dJointAttach: a joint whose first body is the environment is stored swapped with
 * dJOINT_REVERSE set */
static void joint_attach(orc_world *w, joint_t *j, int b1, int b2)
{
    if (b1 < 0 && b2 >= 0) {
        j->b[0] = b2;
        j->b[1] = -1;
        j->reverse = 1;
    } else {
        j->b[0] = b1;
        j->b[1] = b2;
        j->reverse = 0;
    }
    j->creation_serial = ++w->serial;
}

int orc_joint_create(orc_world *w, int type, int b1, int b2)
{
    joint_t *j = joint_alloc(&w->joints, &w->nj, &w->capj);
    j->type = type;
    j->qrel[0] = 1;
    j->axis1[0] = 1;
    j->axis2[0] = 1;
    limot_init(&j->limot1, w);
    limot_init(&j->limot2, w);
    j->erp = w->erp;
    j->cfm = w->cfm;
    j->qrel2[0] = 1;
    if (type == ORC_JOINT_UNIVERSAL) {
        j->axis2[0] = 0;
        j->axis2[1] = 1;
    }
    if (type == ORC_JOINT_HINGE2) {
        j->axis2[0] = 0;
        j->axis2[1] = 1;
        j->v1[0] = 1;
        j->v2[1] = 1;
        j->w1[0] = 1;
        j->w2[1] = 1;
        j->susp_erp = w->erp;
        j->susp_cfm = w->cfm;
    }
    j->limot_row = -1;
    joint_attach(w, j, b1, b2);
    return w->nj - 1;
}

static void set_anchors(orc_world *w, joint_t *j, real x, real y, real z)
{
    if (j->b[0] < 0) return;
    const body_t *b0 = &w->bodies[j->b[0]];
    real q[3] = {x - b0->pos[0], y - b0->pos[1], z - b0->pos[2]};
    mul_Rt_v(j->anchor1, b0->R, q);
    if (j->b[1] >= 0) {
        const body_t *b1 = &w->bodies[j->b[1]];
        real q2[3] = {x - b1->pos[0], y - b1->pos[1], z - b1->pos[2]};
        mul_Rt_v(j->anchor2, b1->R, q2);
    } else {
        j->anchor2[0] = x;
        j->anchor2[1] = y;
        j->anchor2[2] = z;
    }
}

static void set_axes(orc_world *w, joint_t *j, real x, real y, real z, real *axis1, real *axis2)
{
    if (j->b[0] < 0) return;
    real q[3] = {x, y, z};
    normalize3(q);
    if (axis1) mul_Rt_v(axis1, w->bodies[j->b[0]].R, q);
    if (axis2) {
        if (j->b[1] >= 0)
            mul_Rt_v(axis2, w->bodies[j->b[1]].R, q);
        else {
            axis2[0] = q[0];
            axis2[1] = q[1];
            axis2[2] = q[2];
        }
    }
}

static void compute_initial_relative_rotation(orc_world *w, joint_t *j)
{
    if (j->b[0] < 0) return;
    const body_t *b0 = &w->bodies[j->b[0]];
    if (j->b[1] >= 0) {
        qmul1(j->qrel, b0->q, w->bodies[j->b[1]].q);
    } else {
        j->qrel[0] = b0->q[0];
        j->qrel[1] = -b0->q[1];
        j->qrel[2] = -b0->q[2];
        j->qrel[3] = -b0->q[3];
    }
}

static void hinge2_get_axis_info(const orc_world *w, const joint_t *j, real *ax1, real *ax2, real *axis,
                                 real *sin_angle, real *cos_angle)
{
    mul_R_v(ax1, w->bodies[j->b[0]].R, j->axis1);
    mul_R_v(ax2, w->bodies[j->b[1]].R, j->axis2);
    cross3(axis, ax1, ax2);
    *sin_angle = R_SQRT(dot3(axis, axis));
    *cos_angle = dot3(ax1, ax2);
}

static int vec_is_zero(const real *a) { return a[0] == 0 && a[1] == 0 && a[2] == 0; }
static int vec_equal(const real *a, const real *b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }

static void hinge2_make_v1_v2(orc_world *w, joint_t *j)
{
    if (j->b[0] < 0 || j->b[1] < 0) return;
    real ax1[3], ax2[3], v[3];
    mul_R_v(ax1, w->bodies[j->b[0]].R, j->axis1);
    mul_R_v(ax2, w->bodies[j->b[1]].R, j->axis2);
    if (vec_is_zero(ax1) || vec_is_zero(ax2) || vec_equal(ax1, ax2)) return;
    /* modify axis 2 so it's perpendicular to axis 1 */
    real k = dot3(ax1, ax2);
    for (int i = 0; i < 3; i++) ax2[i] -= k * ax1[i];
    normalize3(ax2);
    /* v1 = modified axis2, v2 = axis1 x (modified axis2), both in body-1 frame */
    cross3(v, ax1, ax2);
    mul_Rt_v(j->v1, w->bodies[j->b[0]].R, ax2);
    mul_Rt_v(j->v2, w->bodies[j->b[0]].R, v);
}

static void hinge2_make_w1_w2(orc_world *w, joint_t *j)
{
    if (j->b[0] < 0 || j->b[1] < 0) return;
    real ax1[3], ax2[3], v[3];
    mul_R_v(ax1, w->bodies[j->b[0]].R, j->axis1);
    mul_R_v(ax2, w->bodies[j->b[1]].R, j->axis2);
    if (vec_is_zero(ax1) || vec_is_zero(ax2) || vec_equal(ax1, ax2)) return;
    /* modify axis 1 so it's perpendicular to axis 2 */
    real k = dot3(ax2, ax1);
    for (int i = 0; i < 3; i++) ax1[i] -= k * ax2[i];
    normalize3(ax1);
    cross3(v, ax2, ax1);
    mul_Rt_v(j->w1, w->bodies[j->b[1]].R, ax1);
    mul_Rt_v(j->w2, w->bodies[j->b[1]].R, v);
}

static void universal_initial_rotations(orc_world *w, joint_t *j);

void orc_joint_set_anchor(orc_world *w, int ji, double x, double y, double z)
{
    joint_t *j = &w->joints[ji];
    switch (j->type) {
    case ORC_JOINT_HINGE:
        set_anchors(w, j, (real)x, (real)y, (real)z);
        compute_initial_relative_rotation(w, j);
        break;
    case ORC_JOINT_HINGE2:
        set_anchors(w, j, (real)x, (real)y, (real)z);
        hinge2_make_v1_v2(w, j);
        hinge2_make_w1_w2(w, j);
        break;
    case ORC_JOINT_BALL: set_anchors(w, j, (real)x, (real)y, (real)z); break;
    case ORC_JOINT_UNIVERSAL:
        set_anchors(w, j, (real)x, (real)y, (real)z);
        universal_initial_rotations(w, j);
        break;
    default: break; /* slider/fixed have no anchor (Joint.cpp:436-438) */
    }
}

static void slider_compute_offset(orc_world *w, joint_t *j)
{
    if (j->b[1] >= 0) {
        const body_t *b0 = &w->bodies[j->b[0]], *b1 = &w->bodies[j->b[1]];
        real c[3] = {b0->pos[0] - b1->pos[0], b0->pos[1] - b1->pos[1], b0->pos[2] - b1->pos[2]};
        mul_Rt_v(j->offset, b1->R, c);
    } else if (j->b[0] >= 0) {
        const body_t *b0 = &w->bodies[j->b[0]];
        j->offset[0] = b0->pos[0];
        j->offset[1] = b0->pos[1];
        j->offset[2] = b0->pos[2];
    }
}

void orc_joint_set_axis1(orc_world *w, int ji, double x, double y, double z)
{
    joint_t *j = &w->joints[ji];
    switch (j->type) {
    case ORC_JOINT_HINGE:
        set_axes(w, j, (real)x, (real)y, (real)z, j->axis1, j->axis2);
        compute_initial_relative_rotation(w, j);
        break;
    case ORC_JOINT_SLIDER:
        set_axes(w, j, (real)x, (real)y, (real)z, j->axis1, 0);
        slider_compute_offset(w, j);
        compute_initial_relative_rotation(w, j);
        break;
    case ORC_JOINT_HINGE2:
        if (j->b[0] >= 0 && j->b[1] >= 0) {
            real ax1[3], ax2[3], ax[3];
            set_axes(w, j, (real)x, (real)y, (real)z, j->axis1, 0);
            hinge2_get_axis_info(w, j, ax1, ax2, ax, &j->s0, &j->c0);
        }
        hinge2_make_v1_v2(w, j);
        hinge2_make_w1_w2(w, j);
        break;
    case ORC_JOINT_UNIVERSAL: /* dJointSetUniversalAxis1 */
        if (j->reverse)
            set_axes(w, j, (real)x, (real)y, (real)z, 0, j->axis2);
        else
            set_axes(w, j, (real)x, (real)y, (real)z, j->axis1, 0);
        universal_initial_rotations(w, j);
        break;
    default: break;
    }
}

void orc_joint_set_axis2(orc_world *w, int ji, double x, double y, double z)
{
    joint_t *j = &w->joints[ji];
    if (j->type == ORC_JOINT_UNIVERSAL) { /* dJointSetUniversalAxis2 */
        if (j->reverse)
            set_axes(w, j, (real)x, (real)y, (real)z, j->axis1, 0);
        else
            set_axes(w, j, (real)x, (real)y, (real)z, 0, j->axis2);
        universal_initial_rotations(w, j);
        return;
    }
    if (j->type != ORC_JOINT_HINGE2) return;
    if (j->b[0] >= 0 && j->b[1] >= 0) {
        real ax1[3], ax2[3], ax[3];
        set_axes(w, j, (real)x, (real)y, (real)z, 0, j->axis2);
        hinge2_get_axis_info(w, j, ax1, ax2, ax, &j->s0, &j->c0);
    }
    hinge2_make_v1_v2(w, j);
    hinge2_make_w1_w2(w, j);
}

void orc_joint_set_fixed(orc_world *w, int ji)
{
    joint_t *j = &w->joints[ji];
    if (j->b[0] < 0) return;
    const body_t *b0 = &w->bodies[j->b[0]];
    if (j->b[1] >= 0) {
        const body_t *b1 = &w->bodies[j->b[1]];
        real ofs[3] = {b0->pos[0] - b1->pos[0], b0->pos[1] - b1->pos[1], b0->pos[2] - b1->pos[2]};
        mul_Rt_v(j->offset, b0->R, ofs);
    } else {
        j->offset[0] = b0->pos[0];
        j->offset[1] = b0->pos[1];
        j->offset[2] = b0->pos[2];
    }
    compute_initial_relative_rotation(w, j);
}

void orc_joint_set_param(orc_world *w, int ji, int param, double value)
{
    joint_t *j = &w->joints[ji];
    if (j->type == ORC_JOINT_FIXED || j->type == ORC_JOINT_BALL) { /* dJointSetFixedParam / dJointSetBallParam */
        if (param == ORC_PARAM_CFM) j->cfm = (real)value;
        if (param == ORC_PARAM_ERP) j->erp = (real)value;
        return;
    }
    if ((param & 0xff00) == ORC_PARAM_GROUP2) {
        limot_set(&j->limot2, param & 0xff, (real)value);
    } else if (param == ORC_PARAM_SUSPENSION_ERP) {
        j->susp_erp = (real)value;
    } else if (param == ORC_PARAM_SUSPENSION_CFM) {
        j->susp_cfm = (real)value;
    } else {
        limot_set(&j->limot1, param, (real)value);
    }
}

double orc_joint_get_param(const orc_world *w, int ji, int param)
{
    const joint_t *j = &w->joints[ji];
    if (j->type == ORC_JOINT_FIXED || j->type == ORC_JOINT_BALL) {
        if (param == ORC_PARAM_CFM) return j->cfm;
        if (param == ORC_PARAM_ERP) return j->erp;
        return 0;
    }
    if ((param & 0xff00) == ORC_PARAM_GROUP2) return limot_get(&j->limot2, param & 0xff);
    if (param == ORC_PARAM_SUSPENSION_ERP) return j->susp_erp;
    if (param == ORC_PARAM_SUSPENSION_CFM) return j->susp_cfm;
    return limot_get(&j->limot1, param);
}

/* anchor in world coordinates from body 1 (dJointGet*Anchor) */
void orc_joint_get_anchor(const orc_world *w, int ji, double out[3])
{
    const joint_t *j = &w->joints[ji];
    real r[3] = {0, 0, 0};
    if (j->b[0] >= 0) {
        const body_t *b0 = &w->bodies[j->b[0]];
        mul_R_v(r, b0->R, j->anchor1);
        for (int i = 0; i < 3; i++) r[i] += b0->pos[i];
    }
    GET3(out, r);
}

void orc_joint_get_anchor2(const orc_world *w, int ji, double out[3])
{
    const joint_t *j = &w->joints[ji];
    real r[3] = {j->anchor2[0], j->anchor2[1], j->anchor2[2]};
    if (j->b[1] >= 0) {
        const body_t *b1 = &w->bodies[j->b[1]];
        mul_R_v(r, b1->R, j->anchor2);
        for (int i = 0; i < 3; i++) r[i] += b1->pos[i];
    }
    GET3(out, r);
}

void orc_joint_get_axis1(const orc_world *w, int ji, double out[3])
{
    const joint_t *j = &w->joints[ji];
    real r[3] = {0, 0, 0};
    if (j->b[0] >= 0) {
        mul_R_v(r, w->bodies[j->b[0]].R, j->axis1);
        if (j->type == ORC_JOINT_SLIDER && j->reverse)
            for (int i = 0; i < 3; i++) r[i] = -r[i];
    }
    GET3(out, r);
}

void orc_joint_get_axis2(const orc_world *w, int ji, double out[3])
{
    const joint_t *j = &w->joints[ji];
    real r[3] = {0, 0, 0};
    if (j->b[1] >= 0) mul_R_v(r, w->bodies[j->b[1]].R, j->axis2);
    GET3(out, r);
}

/* ODE getHingeAngleFromRelativeQuat / getHingeAngle (joint.cpp) */
static real hinge_angle_from_qrel(const real *qrel, const real *axis)
{
    real cost2 = qrel[0];
    real sint2 = R_SQRT(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
    real theta = (dot3(qrel + 1, axis) >= 0) ? (2 * R_ATAN2(sint2, cost2)) : (2 * R_ATAN2(sint2, -cost2));
    if (theta > R_PI) theta -= 2 * R_PI;
    return -theta;
}

static real hinge_angle(const orc_world *w, const joint_t *j)
{
    const body_t *b0 = &w->bodies[j->b[0]];
    real qrel[4];
    if (j->b[1] >= 0) {
        real qq[4];
        qmul1(qq, b0->q, w->bodies[j->b[1]].q);
        qmul2(qrel, qq, j->qrel);
    } else {
        qmul3(qrel, b0->q, j->qrel);
    }
    return hinge_angle_from_qrel(qrel, j->axis1);
}

/* ---- universal joint (ODE 0.16 universal.cpp) ---- */
static void qmul0(real *r, const real *b, const real *c) /* b * c */
{
    r[0] = b[0] * c[0] - b[1] * c[1] - b[2] * c[2] - b[3] * c[3];
    r[1] = b[0] * c[1] + b[1] * c[0] + b[2] * c[3] - b[3] * c[2];
    r[2] = b[0] * c[2] + b[2] * c[0] + b[3] * c[1] - b[1] * c[3];
    r[3] = b[0] * c[3] + b[3] * c[0] + b[1] * c[2] - b[2] * c[1];
}

/* ODE rotation.cpp dRFrom2Axes: columns a^, (b orthogonalised against a)^, a^ x b^ */
static void r_from_2_axes(real *R, const real *a, const real *b)
{
    real ax = a[0], ay = a[1], az = a[2], bx = b[0], by = b[1], bz = b[2];
    real l = R_SQRT(ax * ax + ay * ay + az * az);
    if (l <= 0) return;
    l = 1 / l;
    ax *= l, ay *= l, az *= l;
    const real k = ax * bx + ay * by + az * bz;
    bx -= k * ax, by -= k * ay, bz -= k * az;
    l = R_SQRT(bx * bx + by * by + bz * bz);
    if (l <= 0) return;
    l = 1 / l;
    bx *= l, by *= l, bz *= l;
    R[0] = ax, R[3] = ay, R[6] = az;
    R[1] = bx, R[4] = by, R[7] = bz;
    R[2] = ay * bz - by * az, R[5] = az * bx - bz * ax, R[8] = ax * by - bx * ay;
}

/* ODE rotation.cpp dRtoQ */
static void r_to_q(const real *R, real *q)
{
    const real tr = R[0] + R[4] + R[8];
    real s;
    if (tr >= 0) {
        s = R_SQRT(tr + 1);
        q[0] = (real)0.5 * s;
        s = (real)0.5 / s;
        q[1] = (R[7] - R[5]) * s;
        q[2] = (R[2] - R[6]) * s;
        q[3] = (R[3] - R[1]) * s;
        return;
    }
    int c = 0;
    if (R[4] > R[0]) c = R[8] > R[4] ? 2 : 1;
    else c = R[8] > R[0] ? 2 : 0;
    if (c == 0) {
        s = R_SQRT((R[0] - (R[4] + R[8])) + 1);
        q[1] = (real)0.5 * s;
        s = (real)0.5 / s;
        q[2] = (R[1] + R[3]) * s;
        q[3] = (R[6] + R[2]) * s;
        q[0] = (R[7] - R[5]) * s;
    } else if (c == 1) {
        s = R_SQRT((R[4] - (R[8] + R[0])) + 1);
        q[2] = (real)0.5 * s;
        s = (real)0.5 / s;
        q[3] = (R[5] + R[7]) * s;
        q[1] = (R[1] + R[3]) * s;
        q[0] = (R[2] - R[6]) * s;
    } else {
        s = R_SQRT((R[8] - (R[0] + R[4])) + 1);
        q[3] = (real)0.5 * s;
        s = (real)0.5 / s;
        q[1] = (R[6] + R[2]) * s;
        q[2] = (R[5] + R[7]) * s;
        q[0] = (R[3] - R[1]) * s;
    }
}

static void universal_axes(const orc_world *w, const joint_t *j, real *ax1, real *ax2)
{
    mul_R_v(ax1, w->bodies[j->b[0]].R, j->axis1);
    if (j->b[1] >= 0)
        mul_R_v(ax2, w->bodies[j->b[1]].R, j->axis2);
    else
        memcpy(ax2, j->axis2, 3 * sizeof(real));
}

/* dxJointUniversal::computeInitialRelativeRotations */
static void universal_initial_rotations(orc_world *w, joint_t *j)
{
    if (j->b[0] < 0) return;
    real ax1[3], ax2[3], R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, qcross[4];
    universal_axes(w, j, ax1, ax2);
    r_from_2_axes(R, ax1, ax2);
    r_to_q(R, qcross);
    qmul1(j->qrel, w->bodies[j->b[0]].q, qcross);
    r_from_2_axes(R, ax2, ax1);
    r_to_q(R, qcross);
    if (j->b[1] >= 0)
        qmul1(j->qrel2, w->bodies[j->b[1]].q, qcross);
    else
        memcpy(j->qrel2, qcross, sizeof(qcross));
}

/* dxJointUniversal::getAngles */
static void universal_angles(const orc_world *w, const joint_t *j, real *angle1, real *angle2)
{
    *angle1 = *angle2 = 0;
    if (j->b[0] < 0) return;
    real ax1[3], ax2[3], R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, qcross[4], qq[4], qrel[4], qcross2[4];
    universal_axes(w, j, ax1, ax2);
    r_from_2_axes(R, ax1, ax2);
    r_to_q(R, qcross);
    qmul1(qq, w->bodies[j->b[0]].q, qcross);
    qmul2(qrel, qq, j->qrel);
    *angle1 = hinge_angle_from_qrel(qrel, j->axis1);
    /* the cross seen from body 2: the first frame turned by 180 degrees about the bisector of the two axes */
    qrel[0] = 0;
    qrel[1] = ax1[0] + ax2[0];
    qrel[2] = ax1[1] + ax2[1];
    qrel[3] = ax1[2] + ax2[2];
    const real l = 1 / R_SQRT(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
    qrel[1] *= l, qrel[2] *= l, qrel[3] *= l;
    qmul0(qcross2, qrel, qcross);
    if (j->b[1] >= 0) {
        qmul1(qq, w->bodies[j->b[1]].q, qcross2);
        qmul2(qrel, qq, j->qrel2);
    } else {
        qmul2(qrel, qcross2, j->qrel2);
    }
    *angle2 = -hinge_angle_from_qrel(qrel, j->axis2);
}

static real slider_position(const orc_world *w, const joint_t *j)
{
    const body_t *b0 = &w->bodies[j->b[0]];
    real ax1[3], q[3];
    mul_R_v(ax1, b0->R, j->axis1);
    if (j->b[1] >= 0) {
        const body_t *b1 = &w->bodies[j->b[1]];
        mul_R_v(q, b1->R, j->offset);
        for (int i = 0; i < 3; i++) q[i] = b0->pos[i] - q[i] - b1->pos[i];
    } else {
        for (int i = 0; i < 3; i++) q[i] = b0->pos[i] - j->offset[i];
        if (j->reverse)
            for (int i = 0; i < 3; i++) ax1[i] = -ax1[i];
    }
    return dot3(ax1, q);
}

static real hinge2_angle1(const orc_world *w, const joint_t *j)
{
    real p[3], q[3];
    if (j->b[1] >= 0)
        mul_R_v(p, w->bodies[j->b[1]].R, j->axis2);
    else
        memcpy(p, j->axis2, sizeof(p));
    if (j->b[0] >= 0)
        mul_Rt_v(q, w->bodies[j->b[0]].R, p);
    else
        memcpy(q, p, sizeof(q));
    real x = dot3(j->v1, q);
    real y = dot3(j->v2, q);
    return -R_ATAN2(y, x);
}

static real hinge2_angle2(const orc_world *w, const joint_t *j)
{
    real p[3], q[3];
    if (j->b[0] >= 0)
        mul_R_v(p, w->bodies[j->b[0]].R, j->axis1);
    else
        memcpy(p, j->axis1, sizeof(p));
    if (j->b[1] >= 0)
        mul_Rt_v(q, w->bodies[j->b[1]].R, p);
    else
        memcpy(q, p, sizeof(q));
    real x = dot3(j->w1, q);
    real y = dot3(j->w2, q);
    return -R_ATAN2(y, x);
}

double orc_joint_get_angle1(const orc_world *w, int ji)
{
    const joint_t *j = &w->joints[ji];
    if (j->b[0] < 0) return 0;
    switch (j->type) {
    case ORC_JOINT_HINGE: {
        real a = hinge_angle(w, j);
        return j->reverse ? -a : a;
    }
    case ORC_JOINT_SLIDER: return slider_position(w, j);
    case ORC_JOINT_HINGE2: return hinge2_angle1(w, j);
    case ORC_JOINT_UNIVERSAL: { /* dJointGetUniversalAngle1 */
        real a1, a2;
        universal_angles(w, j, &a1, &a2);
        return j->reverse ? -a2 : a1;
    }
    default: return 0;
    }
}

double orc_joint_get_angle1_rate(const orc_world *w, int ji)
{
    const joint_t *j = &w->joints[ji];
    if (j->b[0] < 0) return 0;
    const body_t *b0 = &w->bodies[j->b[0]];
    real axis[3];
    mul_R_v(axis, b0->R, j->axis1);
    switch (j->type) {
    case ORC_JOINT_UNIVERSAL: { /* dJointGetUniversalAngle1Rate */
        real ax1[3], ax2[3];
        universal_axes(w, j, ax1, ax2);
        const real *ax = j->reverse ? ax2 : ax1;
        real rate = dot3(ax, b0->avel);
        if (j->b[1] >= 0) rate -= dot3(ax, w->bodies[j->b[1]].avel);
        return rate;
    }
    case ORC_JOINT_HINGE:
    case ORC_JOINT_HINGE2: {
        real rate = dot3(axis, b0->avel);
        if (j->b[1] >= 0) rate -= dot3(axis, w->bodies[j->b[1]].avel);
        if (j->type == ORC_JOINT_HINGE && j->reverse) rate = -rate;
        return rate;
    }
    case ORC_JOINT_SLIDER: {
        if (j->b[1] >= 0) return dot3(axis, b0->lvel) - dot3(axis, w->bodies[j->b[1]].lvel);
        real rate = dot3(axis, b0->lvel);
        return j->reverse ? -rate : rate;
    }
    default: return 0;
    }
}

double orc_joint_get_angle2(const orc_world *w, int ji)
{
    const joint_t *j = &w->joints[ji];
    if (j->type == ORC_JOINT_UNIVERSAL && j->b[0] >= 0) { /* dJointGetUniversalAngle2 */
        real a1, a2;
        universal_angles(w, j, &a1, &a2);
        return j->reverse ? -a1 : a2;
    }
    if (j->type != ORC_JOINT_HINGE2 || j->b[0] < 0) return 0;
    return hinge2_angle2(w, j);
}

double orc_joint_get_angle2_rate(const orc_world *w, int ji)
{
    const joint_t *j = &w->joints[ji];
    if (j->type == ORC_JOINT_UNIVERSAL && j->b[0] >= 0) { /* dJointGetUniversalAngle2Rate */
        real ax1[3], ax2[3];
        universal_axes(w, j, ax1, ax2);
        const real *ax = j->reverse ? ax1 : ax2;
        real rate = dot3(ax, w->bodies[j->b[0]].avel);
        if (j->b[1] >= 0) rate -= dot3(ax, w->bodies[j->b[1]].avel);
        return rate;
    }
    if (j->type != ORC_JOINT_HINGE2 || j->b[0] < 0 || j->b[1] < 0) return 0;
    real axis[3];
    mul_R_v(axis, w->bodies[j->b[1]].R, j->axis2);
    return dot3(axis, w->bodies[j->b[0]].avel) - dot3(axis, w->bodies[j->b[1]].avel);
}

/* dJointAddHingeTorque / dJointAddSliderForce */
void orc_joint_add_torque(orc_world *w, int ji, double t)
{
    joint_t *j = &w->joints[ji];
    if (j->b[0] < 0) return;
    body_t *b0 = &w->bodies[j->b[0]];
    real torque = (real)t;
    if (j->reverse) torque = -torque;
    real axis[3];
    mul_R_v(axis, b0->R, j->axis1);
    for (int i = 0; i < 3; i++) axis[i] *= torque;
    if (j->type == ORC_JOINT_HINGE) {
        for (int i = 0; i < 3; i++) b0->tacc[i] += axis[i];
        if (j->b[1] >= 0)
            for (int i = 0; i < 3; i++) w->bodies[j->b[1]].tacc[i] -= axis[i];
    } else if (j->type == ORC_JOINT_SLIDER) {
        if (j->b[1] >= 0) {
            /* linear torque decoupling: apply the force at the point halfway between the bodies */
            body_t *b1 = &w->bodies[j->b[1]];
            real c[3], ltd[3];
            for (int i = 0; i < 3; i++) c[i] = (real)0.5 * (b1->pos[i] - b0->pos[i]);
            cross3(ltd, c, axis);
            for (int i = 0; i < 3; i++) {
                b0->tacc[i] += ltd[i];
                b1->tacc[i] += ltd[i];
                b1->facc[i] -= axis[i];
            }
        }
        for (int i = 0; i < 3; i++) b0->facc[i] += axis[i];
    }
}

void orc_joint_get_feedback(const orc_world *w, int ji, double out[13])
{
    const joint_t *j = &w->joints[ji];
    GET3(out, j->fb_f1);
    GET3(out + 3, j->fb_t1);
    GET3(out + 6, j->fb_f2);
    GET3(out + 9, j->fb_t2);
    out[12] = j->fb_lambda;
}

int orc_joint_get_num_rows(const orc_world *w, int ji) { return w->joints[ji].last_m; }

/* dAreConnectedExcluding (ode.cpp): any joint attached to b1 whose type != joint_type that also
 * attaches b2 */
int orc_are_connected_excluding(const orc_world *w, int b1, int b2, int joint_type)
{
    if (b1 < 0 || b2 < 0) return 0;
    for (int i = 0; i < w->nj; i++) {
        const joint_t *j = &w->joints[i];
        if (j->type == joint_type) continue;
        if ((j->b[0] == b1 && j->b[1] == b2) || (j->b[0] == b2 && j->b[1] == b1)) return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* contacts                                                            */

void orc_contacts_clear(orc_world *w) { w->nc = 0; }

int orc_contact_add(orc_world *w, int b1, int b2, const orc_contact_desc *c)
{
    /* WorldPhysics::createContact, src/WorldPhysics.cpp:458-468 */
    if (b1 >= 0 && b2 >= 0 && orc_are_connected_excluding(w, b1, b2, ORC_JOINT_CONTACT)) return -1;
    joint_t *j = joint_alloc(&w->contacts, &w->nc, &w->capc);
    j->type = ORC_JOINT_CONTACT;
    j->contact = *c;
    j->limot_row = -1;
    joint_attach(w, j, b1, b2);
    return w->nc - 1;
}

void orc_contact_get_feedback(const orc_world *w, int c, double out[12])
{
    const joint_t *j = &w->contacts[c];
    GET3(out, j->fb_f1);
    GET3(out + 3, j->fb_t1);
    GET3(out + 6, j->fb_f2);
    GET3(out + 9, j->fb_t2);
}

/* ------------------------------------------------------------------ */
/* constraint rows (getInfo1 / getInfo2)                               */

typedef struct {
    real J[12]; /* J1l J1a J2l J2a */
    real c, cfm, lo, hi;
    int findex; /* local row index of the normal row, or -1 */
} row_t;

static void row_defaults(const orc_world *w, row_t *r, int n)
{
    for (int i = 0; i < n; i++) {
        memset(&r[i], 0, sizeof(row_t));
        r[i].cfm = w->cfm;
        r[i].lo = -R_INF;
        r[i].hi = R_INF;
        r[i].findex = -1;
    }
}

static int contact_info1(joint_t *j)
{
    orc_contact_desc *c = &j->contact;
    int m = 1;
    if (c->mode & ORC_CONTACT_MU2) { /* dContactAxisDep */
        if (c->mu < 0) c->mu = 0;
        if (c->mu > 0) m++;
        if (c->mu2 < 0) c->mu2 = 0;
        if (c->mu2 > 0) m++;
        if (c->mode & ORC_CONTACT_ROLLING) {
            if (c->rho < 0) c->rho = 0;
            if (c->rho > 0) m++;
            if (c->rho2 < 0) c->rho2 = 0;
            if (c->rho2 > 0) m++;
            if (c->rhoN < 0) c->rhoN = 0;
            if (c->rhoN > 0) m++;
        }
    } else {
        if (c->mu < 0) c->mu = 0;
        if (c->mu > 0) m += 2;
        if (c->mode & ORC_CONTACT_ROLLING) {
            if (c->rho < 0) c->rho = 0;
            if (c->rho > 0) m += 3;
        }
    }
    return m;
}

static int joint_info1(orc_world *w, joint_t *j)
{
    switch (j->type) {
    case ORC_JOINT_CONTACT: return contact_info1(j);
    case ORC_JOINT_HINGE: {
        int m = (j->limot1.fmax > 0) ? 6 : 5;
        /* stateless restatement: the limit flag is re-evaluated every step */
        j->limot1.limit = 0;
        if ((j->limot1.lostop >= -R_PI || j->limot1.histop <= R_PI) && j->limot1.lostop <= j->limot1.histop) {
            real angle = hinge_angle(w, j);
            if (limot_test_rotational_limit(&j->limot1, angle)) m = 6;
        }
        return m;
    }
    case ORC_JOINT_SLIDER: {
        int m = (j->limot1.fmax > 0) ? 6 : 5;
        j->limot1.limit = 0;
        if ((j->limot1.lostop > -R_INF || j->limot1.histop < R_INF) && j->limot1.lostop <= j->limot1.histop) {
            real pos = slider_position(w, j);
            if (pos <= j->limot1.lostop) {
                j->limot1.limit = 1;
                j->limot1.limit_err = pos - j->limot1.lostop;
                m = 6;
            } else if (pos >= j->limot1.histop) {
                j->limot1.limit = 2;
                j->limot1.limit_err = pos - j->limot1.histop;
                m = 6;
            }
        }
        return m;
    }
    case ORC_JOINT_FIXED: return 6;
    case ORC_JOINT_HINGE2: {
        int m = 4;
        j->limot1.limit = 0;
        if ((j->limot1.lostop >= -R_PI || j->limot1.histop <= R_PI) && j->limot1.lostop <= j->limot1.histop) {
            real angle = hinge2_angle1(w, j);
            limot_test_rotational_limit(&j->limot1, angle);
        }
        if (j->limot1.limit || j->limot1.fmax > 0) m++;
        j->limot2.limit = 0;
        if (j->limot2.fmax > 0) m++;
        return m;
    }
    case ORC_JOINT_BALL: return 3;
    case ORC_JOINT_UNIVERSAL: { /* dxJointUniversal::getInfo1 */
        int m = 4;
        const int lim1 = (j->limot1.lostop >= -R_PI || j->limot1.histop <= R_PI) && j->limot1.lostop <= j->limot1.histop;
        const int lim2 = (j->limot2.lostop >= -R_PI || j->limot2.histop <= R_PI) && j->limot2.lostop <= j->limot2.histop;
        j->limot1.limit = j->limot2.limit = 0;
        if (lim1 || lim2) {
            real a1, a2;
            universal_angles(w, j, &a1, &a2);
            if (lim1) limot_test_rotational_limit(&j->limot1, a1);
            if (lim2) limot_test_rotational_limit(&j->limot2, a2);
        }
        if (j->limot1.limit || j->limot1.fmax > 0) m++;
        if (j->limot2.limit || j->limot2.fmax > 0) m++;
        return m;
    }
    default: return 0;
    }
}

/* ODE joint.cpp dxJointLimitMotor::addLimot */
static int add_limot(orc_world *w, joint_t *j, limot_t *l, real fps, row_t *row, const real *ax1, int rotational)
{
    int powered = l->fmax > 0;
    if (!(powered || l->limit)) return 0;
    body_t *b0 = &w->bodies[j->b[0]];
    body_t *b1 = j->b[1] >= 0 ? &w->bodies[j->b[1]] : 0;
    real *J1 = rotational ? row->J + 3 : row->J;
    real *J2 = rotational ? row->J + 9 : row->J + 6;
    for (int i = 0; i < 3; i++) J1[i] = ax1[i];
    if (b1)
        for (int i = 0; i < 3; i++) J2[i] = -ax1[i];

    real ltd[3] = {0, 0, 0}; /* linear torque decoupling vector */
    if (!rotational && b1) {
        real c[3];
        for (int i = 0; i < 3; i++) c[i] = (real)0.5 * (b1->pos[i] - b0->pos[i]);
        cross3(ltd, c, ax1);
        for (int i = 0; i < 3; i++) {
            row->J[3 + i] = ltd[i];
            row->J[9 + i] = ltd[i];
        }
    }

    if (l->limit && (l->lostop == l->histop)) powered = 0;

    if (powered) {
        row->cfm = l->normal_cfm;
        if (!l->limit) {
            row->c = l->vel;
            row->lo = -l->fmax;
            row->hi = l->fmax;
        } else {
            /* at a limit and powered: apply the motor force directly */
            real fm = l->fmax;
            if ((l->vel > 0) || (l->vel == 0 && l->limit == 2)) fm = -fm;
            if ((l->limit == 1 && l->vel > 0) || (l->limit == 2 && l->vel < 0)) fm *= l->fudge_factor;
            if (rotational) {
                for (int i = 0; i < 3; i++) {
                    if (b1) b1->tacc[i] += fm * ax1[i];
                    b0->tacc[i] -= fm * ax1[i];
                }
            } else {
                if (b1) {
                    for (int i = 0; i < 3; i++) {
                        b0->tacc[i] -= fm * ltd[i];
                        b1->tacc[i] -= fm * ltd[i];
                        b1->facc[i] += fm * ax1[i];
                    }
                }
                for (int i = 0; i < 3; i++) b0->facc[i] -= fm * ax1[i];
            }
        }
    }

    if (l->limit) {
        real k = fps * l->stop_erp;
        row->c = -k * l->limit_err;
        row->cfm = l->stop_cfm;
        if (l->lostop == l->histop) {
            row->lo = -R_INF;
            row->hi = R_INF;
        } else {
            if (l->limit == 1) {
                row->lo = 0;
                row->hi = R_INF;
            } else {
                row->lo = -R_INF;
                row->hi = 0;
            }
            if (l->bounce > 0) {
                real vel;
                if (rotational) {
                    vel = dot3(b0->avel, ax1);
                    if (b1) vel -= dot3(b1->avel, ax1);
                } else {
                    vel = dot3(b0->lvel, ax1);
                    if (b1) vel -= dot3(b1->lvel, ax1);
                }
                if (l->limit == 1) {
                    if (vel < 0) {
                        real newc = -l->bounce * vel;
                        if (newc > row->c) row->c = newc;
                    }
                } else {
                    if (vel > 0) {
                        real newc = -l->bounce * vel;
                        if (newc < row->c) row->c = newc;
                    }
                }
            }
        }
    }
    return 1;
}

/* ODE joint.cpp setBall */
static void set_ball(const orc_world *w, const joint_t *j, real fps, real erp, row_t *r)
{
    const body_t *b0 = &w->bodies[j->b[0]];
    real a1[3], a2[3];
    r[0].J[0] = 1;
    r[1].J[1] = 1;
    r[2].J[2] = 1;
    mul_R_v(a1, b0->R, j->anchor1);
    /* J1a = -crossmat(a1) */
    r[0].J[3 + 1] = a1[2];
    r[0].J[3 + 2] = -a1[1];
    r[1].J[3 + 0] = -a1[2];
    r[1].J[3 + 2] = a1[0];
    r[2].J[3 + 0] = a1[1];
    r[2].J[3 + 1] = -a1[0];
    real k = fps * erp;
    if (j->b[1] >= 0) {
        const body_t *b1 = &w->bodies[j->b[1]];
        r[0].J[6 + 0] = -1;
        r[1].J[6 + 1] = -1;
        r[2].J[6 + 2] = -1;
        mul_R_v(a2, b1->R, j->anchor2);
        /* J2a = +crossmat(a2) */
        r[0].J[9 + 1] = -a2[2];
        r[0].J[9 + 2] = a2[1];
        r[1].J[9 + 0] = a2[2];
        r[1].J[9 + 2] = -a2[0];
        r[2].J[9 + 0] = -a2[1];
        r[2].J[9 + 1] = a2[0];
        for (int i = 0; i < 3; i++) r[i].c = k * (a2[i] + b1->pos[i] - a1[i] - b0->pos[i]);
    } else {
        for (int i = 0; i < 3; i++) r[i].c = k * (j->anchor2[i] - a1[i] - b0->pos[i]);
    }
}

/* ODE joint.cpp setBall2: ball rows aligned to `axis`, row 0 uses erp1 */
static void set_ball2(const orc_world *w, const joint_t *j, real fps, real erp, row_t *r, const real *axis,
                      real erp1)
{
    const body_t *b0 = &w->bodies[j->b[0]];
    real a1[3], a2[3], q1[3], q2[3];
    plane_space(axis, q1, q2);
    const real *dirs[3] = {axis, q1, q2};
    mul_R_v(a1, b0->R, j->anchor1);
    for (int i = 0; i < 3; i++) {
        memcpy(r[i].J, dirs[i], 3 * sizeof(real));
        cross3(r[i].J + 3, a1, dirs[i]);
    }
    real k1 = fps * erp1, k = fps * erp;
    for (int i = 0; i < 3; i++) a1[i] += b0->pos[i];
    if (j->b[1] >= 0) {
        const body_t *b1 = &w->bodies[j->b[1]];
        mul_R_v(a2, b1->R, j->anchor2);
        for (int i = 0; i < 3; i++) {
            for (int c = 0; c < 3; c++) r[i].J[6 + c] = -dirs[i][c];
            cross3(r[i].J + 9, dirs[i], a2); /* = -(a2 x dir) */
        }
        real d[3];
        for (int i = 0; i < 3; i++) d[i] = a2[i] + b1->pos[i] - a1[i];
        r[0].c = k1 * dot3(axis, d);
        r[1].c = k * dot3(q1, d);
        r[2].c = k * dot3(q2, d);
    } else {
        r[0].c = k1 * (dot3(axis, j->anchor2) - dot3(axis, a1));
        r[1].c = k * (dot3(q1, j->anchor2) - dot3(q1, a1));
        r[2].c = k * (dot3(q2, j->anchor2) - dot3(q2, a1));
    }
}

/* ODE joint.cpp setFixedOrientation */
static void set_fixed_orientation(const orc_world *w, const joint_t *j, real fps, real erp, row_t *r)
{
    const body_t *b0 = &w->bodies[j->b[0]];
    r[0].J[3] = 1;
    r[1].J[4] = 1;
    r[2].J[5] = 1;
    real qerr[4], e[3];
    if (j->b[1] >= 0) {
        r[0].J[9] = -1;
        r[1].J[10] = -1;
        r[2].J[11] = -1;
        real qq[4];
        qmul1(qq, b0->q, w->bodies[j->b[1]].q);
        qmul2(qerr, qq, j->qrel);
    } else {
        qmul3(qerr, b0->q, j->qrel);
    }
    if (qerr[0] < 0) {
        qerr[1] = -qerr[1];
        qerr[2] = -qerr[2];
        qerr[3] = -qerr[3];
    }
    mul_R_v(e, b0->R, qerr + 1);
    real k2 = fps * erp * 2;
    for (int i = 0; i < 3; i++) r[i].c = k2 * e[i];
}

static void contact_info2(const orc_world *w, joint_t *j, real fps, real world_erp, row_t *r, int m)
{
    const orc_contact_desc *c = &j->contact;
    const int mode = c->mode;
    const body_t *b0 = &w->bodies[j->b[0]];
    const body_t *b1 = j->b[1] >= 0 ? &w->bodies[j->b[1]] : 0;

    real erp = (mode & ORC_CONTACT_SOFT_ERP) ? (real)c->soft_erp : world_erp;
    real k = fps * erp;
    real depth = (real)c->depth - w->contact_min_depth;
    if (depth < 0) depth = 0;
    real motionN = (mode & ORC_CONTACT_MOTIONN) ? (real)c->motionN : 0;
    real pushout = k * depth + motionN;
    int apply_bounce = (mode & ORC_CONTACT_BOUNCE) && c->bounce_vel >= 0;
    real maxvel = w->contact_max_vel;
    real cc = pushout > maxvel ? maxvel : pushout;

    real normal[3], c1[3], c2[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
        normal[i] = j->reverse ? -(real)c->normal[i] : (real)c->normal[i];
        c1[i] = (real)c->pos[i] - b0->pos[i];
    }
    memcpy(r[0].J, normal, 3 * sizeof(real));
    cross3(r[0].J + 3, c1, normal);
    real outgoing = 0;
    if (b1) {
        for (int i = 0; i < 3; i++) {
            c2[i] = (real)c->pos[i] - b1->pos[i];
            r[0].J[6 + i] = -normal[i];
        }
        cross3(r[0].J + 9, normal, c2); /* = -(c2 x normal) */
    }
    if (apply_bounce) {
        outgoing = dot3(r[0].J, b0->lvel) + dot3(r[0].J + 3, b0->avel);
        if (b1) outgoing += dot3(r[0].J + 6, b1->lvel) + dot3(r[0].J + 9, b1->avel);
        outgoing -= motionN;
        if (outgoing < 0) { /* only apply bounce if the velocity is incoming */
            if (-outgoing > (real)c->bounce_vel) {
                real newc = -(real)c->bounce * outgoing + motionN;
                if (newc > cc) cc = newc;
            }
        }
    }
    r[0].c = cc;
    if (mode & ORC_CONTACT_SOFT_CFM) r[0].cfm = (real)c->soft_cfm;
    r[0].lo = 0;
    r[0].hi = R_INF;
    if (m <= 1) return;

    real t1[3], t2[3];
    plane_space(normal, t1, t2); /* reference never sets dContactFDir1 (Frame.cpp:1040-1046) */
    int row = 1;
    real mu = (real)c->mu;
    if (mu > 0) {
        memcpy(r[row].J, t1, 3 * sizeof(real));
        cross3(r[row].J + 3, c1, t1);
        if (b1) {
            for (int i = 0; i < 3; i++) r[row].J[6 + i] = -t1[i];
            cross3(r[row].J + 9, t1, c2);
        }
        if (mode & ORC_CONTACT_MOTION1) r[row].c = (real)c->motion1;
        r[row].lo = -mu;
        r[row].hi = mu;
        if (mode & ORC_CONTACT_APPROX1_1) r[row].findex = 0;
        if (mode & ORC_CONTACT_SLIP1) r[row].cfm = (real)c->slip1;
        row++;
    }
    real mu2 = (mode & ORC_CONTACT_MU2) ? (real)c->mu2 : mu;
    if (mu2 > 0) {
        memcpy(r[row].J, t2, 3 * sizeof(real));
        cross3(r[row].J + 3, c1, t2);
        if (b1) {
            for (int i = 0; i < 3; i++) r[row].J[6 + i] = -t2[i];
            cross3(r[row].J + 9, t2, c2);
        }
        if (mode & ORC_CONTACT_MOTION2) r[row].c = (real)c->motion2;
        r[row].lo = -mu2;
        r[row].hi = mu2;
        if (mode & ORC_CONTACT_APPROX1_2) r[row].findex = 0;
        if (mode & ORC_CONTACT_SLIP2) r[row].cfm = (real)c->slip2;
        row++;
    }
    if (mode & ORC_CONTACT_ROLLING) {
        const real *ax[3] = {t1, t2, normal};
        const int approx_bits[3] = {ORC_CONTACT_APPROX1_1, ORC_CONTACT_APPROX1_2, ORC_CONTACT_APPROX1_N};
        real rho[3];
        rho[0] = (real)c->rho;
        if (mode & ORC_CONTACT_MU2) {
            rho[1] = (real)c->rho2;
            rho[2] = (real)c->rhoN;
        } else {
            rho[1] = rho[0];
            rho[2] = rho[0];
        }
        for (int i = 0; i < 3; i++) {
            if (rho[i] > 0) {
                memcpy(r[row].J + 3, ax[i], 3 * sizeof(real));
                if (b1)
                    for (int q = 0; q < 3; q++) r[row].J[9 + q] = -ax[i][q];
                r[row].lo = -rho[i];
                r[row].hi = rho[i];
                if (mode & approx_bits[i]) r[row].findex = 0;
                row++;
            }
        }
    }
}

static void joint_info2(orc_world *w, joint_t *j, real fps, row_t *r, int m)
{
    const real world_erp = w->erp;
    j->limot_row = -1;
    switch (j->type) {
    case ORC_JOINT_CONTACT: contact_info2(w, j, fps, world_erp, r, m); break;
    case ORC_JOINT_HINGE: {
        const body_t *b0 = &w->bodies[j->b[0]];
        set_ball(w, j, fps, world_erp, r);
        real ax1[3], p[3], q[3], ax2[3], b[3];
        mul_R_v(ax1, b0->R, j->axis1);
        plane_space(ax1, p, q);
        memcpy(r[3].J + 3, p, 3 * sizeof(real));
        memcpy(r[4].J + 3, q, 3 * sizeof(real));
        if (j->b[1] >= 0) {
            for (int i = 0; i < 3; i++) {
                r[3].J[9 + i] = -p[i];
                r[4].J[9 + i] = -q[i];
            }
            mul_R_v(ax2, w->bodies[j->b[1]].R, j->axis2);
        } else {
            memcpy(ax2, j->axis2, sizeof(ax2));
        }
        cross3(b, ax1, ax2);
        real k = fps * world_erp;
        r[3].c = k * dot3(b, p);
        r[4].c = k * dot3(b, q);
        if (m > 5 && add_limot(w, j, &j->limot1, fps, &r[5], ax1, 1)) j->limot_row = 5;
        break;
    }
    case ORC_JOINT_FIXED: {
        const body_t *b0 = &w->bodies[j->b[0]];
        set_fixed_orientation(w, j, fps, world_erp, r + 3);
        r[0].J[0] = 1;
        r[1].J[1] = 1;
        r[2].J[2] = 1;
        real ofs[3];
        mul_R_v(ofs, b0->R, j->offset);
        real k = fps * j->erp;
        if (j->b[1] >= 0) {
            const body_t *b1 = &w->bodies[j->b[1]];
            /* J1a = +crossmat(ofs) */
            r[0].J[3 + 1] = -ofs[2];
            r[0].J[3 + 2] = ofs[1];
            r[1].J[3 + 0] = ofs[2];
            r[1].J[3 + 2] = -ofs[0];
            r[2].J[3 + 0] = -ofs[1];
            r[2].J[3 + 1] = ofs[0];
            r[0].J[6] = -1;
            r[1].J[7] = -1;
            r[2].J[8] = -1;
            for (int i = 0; i < 3; i++) r[i].c = k * (b1->pos[i] - b0->pos[i] + ofs[i]);
        } else {
            for (int i = 0; i < 3; i++) r[i].c = k * (j->offset[i] - b0->pos[i]);
        }
        r[0].cfm = r[1].cfm = r[2].cfm = j->cfm;
        break;
    }
    case ORC_JOINT_SLIDER: {
        const body_t *b0 = &w->bodies[j->b[0]];
        const body_t *b1 = j->b[1] >= 0 ? &w->bodies[j->b[1]] : 0;
        set_fixed_orientation(w, j, fps, world_erp, r);
        real c[3] = {0, 0, 0}, ax1[3], p[3], q[3];
        if (b1)
            for (int i = 0; i < 3; i++) c[i] = b1->pos[i] - b0->pos[i];
        mul_R_v(ax1, b0->R, j->axis1);
        plane_space(ax1, p, q);
        if (b1) {
            real tmp[3];
            cross3(tmp, c, p);
            for (int i = 0; i < 3; i++) {
                r[3].J[3 + i] = (real)0.5 * tmp[i];
                r[3].J[9 + i] = (real)0.5 * tmp[i];
                r[3].J[6 + i] = -p[i];
            }
            cross3(tmp, c, q);
            for (int i = 0; i < 3; i++) {
                r[4].J[3 + i] = (real)0.5 * tmp[i];
                r[4].J[9 + i] = (real)0.5 * tmp[i];
                r[4].J[6 + i] = -q[i];
            }
        }
        memcpy(r[3].J, p, 3 * sizeof(real));
        memcpy(r[4].J, q, 3 * sizeof(real));
        real k = fps * world_erp;
        if (b1) {
            real ofs[3];
            mul_R_v(ofs, b1->R, j->offset);
            for (int i = 0; i < 3; i++) c[i] += ofs[i];
            r[3].c = k * dot3(p, c);
            r[4].c = k * dot3(q, c);
        } else {
            real ofs[3];
            for (int i = 0; i < 3; i++) ofs[i] = j->offset[i] - b0->pos[i];
            r[3].c = k * dot3(p, ofs);
            r[4].c = k * dot3(q, ofs);
            if (j->reverse)
                for (int i = 0; i < 3; i++) ax1[i] = -ax1[i];
        }
        if (m > 5 && add_limot(w, j, &j->limot1, fps, &r[5], ax1, 0)) j->limot_row = 5;
        break;
    }
    case ORC_JOINT_HINGE2: {
        real s, c, q[3], ax1[3], ax2[3];
        hinge2_get_axis_info(w, j, ax1, ax2, q, &s, &c);
        normalize3(q);
        set_ball2(w, j, fps, world_erp, r, ax1, j->susp_erp);
        r[0].cfm = j->susp_cfm;
        memcpy(r[3].J + 3, q, 3 * sizeof(real));
        for (int i = 0; i < 3; i++) r[3].J[9 + i] = -q[i];
        real k = fps * world_erp;
        r[3].c = k * (j->c0 * s - j->s0 * c);
        int row = 4;
        if (row < m && add_limot(w, j, &j->limot1, fps, &r[row], ax1, 1)) {
            j->limot_row = row;
            row++;
        }
        if (row < m && add_limot(w, j, &j->limot2, fps, &r[row], ax2, 1)) {
            if (j->limot_row < 0) j->limot_row = row;
        }
        break;
    }
    case ORC_JOINT_BALL: { /* dxJointBall::getInfo2: the joint's own erp / cfm */
        set_ball(w, j, fps, j->erp, r);
        r[0].cfm = r[1].cfm = r[2].cfm = j->cfm;
        break;
    }
    case ORC_JOINT_UNIVERSAL: { /* dxJointUniversal::getInfo2 */
        real ax1[3], ax2[3], ax2t[3], p[3];
        set_ball(w, j, fps, world_erp, r);
        universal_axes(w, j, ax1, ax2);
        const real k = dot3(ax1, ax2);
        for (int i = 0; i < 3; i++) ax2t[i] = ax2[i] - k * ax1[i];
        cross3(p, ax1, ax2t);
        normalize3(p);
        memcpy(r[3].J + 3, p, 3 * sizeof(real));
        if (j->b[1] >= 0)
            for (int i = 0; i < 3; i++) r[3].J[9 + i] = -p[i];
        r[3].c = fps * world_erp * -k;
        int row = 4;
        if (row < m && add_limot(w, j, &j->limot1, fps, &r[row], ax1, 1)) {
            j->limot_row = row;
            row++;
        }
        if (row < m && add_limot(w, j, &j->limot2, fps, &r[row], ax2, 1)) {
            if (j->limot_row < 0) j->limot_row = row;
        }
        break;
    }
    default: break;
    }
}

/* ------------------------------------------------------------------ */
/* the stepper                                                         */

static void step_body(body_t *b, real h)
{
    /* ODE util.cpp dxStepBody */
    if (b->has_max_ang) {
        real aspeed = dot3(b->avel, b->avel);
        if (aspeed > b->max_angular_speed * b->max_angular_speed) {
            real coef = b->max_angular_speed / R_SQRT(aspeed);
            for (int i = 0; i < 3; i++) b->avel[i] *= coef;
        }
    }
    for (int i = 0; i < 3; i++) b->pos[i] += h * b->lvel[i];
    /* dDQfromW */
    real dq[4];
    const real *wv = b->avel, *q = b->q;
    dq[0] = (real)0.5 * (-wv[0] * q[1] - wv[1] * q[2] - wv[2] * q[3]);
    dq[1] = (real)0.5 * (wv[0] * q[0] + wv[1] * q[3] - wv[2] * q[2]);
    dq[2] = (real)0.5 * (-wv[0] * q[3] + wv[1] * q[0] + wv[2] * q[1]);
    dq[3] = (real)0.5 * (wv[0] * q[2] - wv[1] * q[1] + wv[2] * q[0]);
    for (int i = 0; i < 4; i++) b->q[i] += h * dq[i];
    normalize4(b->q);
    q_to_R(b->q, b->R);
}

typedef struct {
    joint_t *j;
    int m, ofs;
} jref_t;

static void quickstep_island(orc_world *w, int *bodies, int nb, jref_t *jl, int nj, real h)
{
    const real hinv = 1 / h;
    /* body index -> local index */
    for (int i = 0; i < nb; i++) w->bodies[bodies[i]].tag = i;

    /* stage 0: inertia in world frame, gyroscopic torque, gravity */
    for (int i = 0; i < nb; i++) {
        body_t *b = &w->bodies[bodies[i]];
        real tmp[9];
        mul33_bt(tmp, b->invI, b->R);
        mul33(b->invI_w, b->R, tmp);
        if (b->gyroscopic && w->variant[ORC_VARIANT_GYRO] == 1) {
            /* ORC_VARIANT_GYRO 1: the explicit form tau = -w x (I w) (ODE before the implicit term) */
            real I[9], L[3];
            mul33_bt(tmp, b->I, b->R);
            mul33(I, b->R, tmp);
            mul_R_v(L, I, b->avel);
            b->tacc[0] -= b->avel[1] * L[2] - b->avel[2] * L[1];
            b->tacc[1] -= b->avel[2] * L[0] - b->avel[0] * L[2];
            b->tacc[2] -= b->avel[0] * L[1] - b->avel[1] * L[0];
        } else if (b->gyroscopic && w->variant[ORC_VARIANT_GYRO] == 0) {
            /* implicit gyroscopic torque (Lacoursiere 2006) as in ODE 0.13+ quickstep */
            real I[9], L[3], It[9], itInv[9];
            mul33_bt(tmp, b->I, b->R);
            mul33(I, b->R, tmp);
            mul_R_v(L, I, b->avel);
            /* Itild = h * crossmat_minus(L) + I ; crossmat_minus(L) x = -(L x x) */
            It[0] = I[0];
            It[1] = I[1] + h * L[2];
            It[2] = I[2] - h * L[1];
            It[3] = I[3] - h * L[2];
            It[4] = I[4];
            It[5] = I[5] + h * L[0];
            It[6] = I[6] + h * L[1];
            It[7] = I[7] - h * L[0];
            It[8] = I[8];
            for (int k = 0; k < 3; k++) L[k] *= hinv;
            if (invert33(itInv, It) != 0) {
                real M[9], tau[3];
                mul33(M, I, itInv);
                M[0] -= 1;
                M[4] -= 1;
                M[8] -= 1;
                mul_R_v(tau, M, L);
                for (int k = 0; k < 3; k++) b->tacc[k] += tau[k];
            }
        }
        if (!b->nogravity)
            for (int k = 0; k < 3; k++) b->facc[k] += b->mass * w->gravity[k];
    }

    /* stage 1: row counts */
    int m = 0, nj_active = 0;
    for (int i = 0; i < nj; i++) {
        joint_t *j = jl[i].j;
        int mi = joint_info1(w, j);
        j->last_m = mi;
        if (mi > 0) {
            jl[nj_active].j = j;
            jl[nj_active].m = mi;
            jl[nj_active].ofs = m;
            m += mi;
            nj_active++;
        }
    }
    nj = nj_active;
    w->last_m += m;

    real *lambda = 0, *fc = 0;
    row_t *rows = 0;
    if (m > 0) {
        rows = (row_t *)malloc(sizeof(row_t) * (size_t)m);
        int *jb = (int *)malloc(sizeof(int) * 2 * (size_t)m);
        real *Jcopy = (real *)malloc(sizeof(real) * 12 * (size_t)m);
        real *iMJ = (real *)malloc(sizeof(real) * 12 * (size_t)m);
        real *rhs = (real *)malloc(sizeof(real) * (size_t)m);
        real *cfm = (real *)malloc(sizeof(real) * (size_t)m);
        real *Adcfm = (real *)malloc(sizeof(real) * (size_t)m);
        real *Ad = (real *)malloc(sizeof(real) * (size_t)m);
        int *findex = (int *)malloc(sizeof(int) * (size_t)m);
        int *order = (int *)malloc(sizeof(int) * (size_t)m);
        lambda = (real *)calloc((size_t)m, sizeof(real));
        fc = (real *)calloc((size_t)nb * 6, sizeof(real));
        real *tmp1 = (real *)malloc(sizeof(real) * 6 * (size_t)nb);

        /* stage 2a: getInfo2 */
        for (int i = 0; i < nj; i++) {
            joint_t *j = jl[i].j;
            row_t *r = rows + jl[i].ofs;
            row_defaults(w, r, jl[i].m);
            joint_info2(w, j, hinv, r, jl[i].m);
            for (int k = 0; k < jl[i].m; k++) {
                int gi = jl[i].ofs + k;
                findex[gi] = r[k].findex >= 0 ? r[k].findex + jl[i].ofs : -1;
                jb[2 * gi] = w->bodies[j->b[0]].tag;
                jb[2 * gi + 1] = j->b[1] >= 0 ? w->bodies[j->b[1]].tag : -1;
                memcpy(Jcopy + 12 * gi, r[k].J, 12 * sizeof(real));
                rhs[gi] = r[k].c * hinv;
                cfm[gi] = r[k].cfm * hinv;
            }
        }

        /* stage 2b: tmp1 = v/h + invM fe  (after getInfo2: limot may have added forces) */
        for (int i = 0; i < nb; i++) {
            body_t *b = &w->bodies[bodies[i]];
            real *t = tmp1 + 6 * i;
            for (int k = 0; k < 3; k++) t[k] = b->facc[k] * b->invMass + b->lvel[k] * hinv;
            mul_R_v(t + 3, b->invI_w, b->tacc);
            for (int k = 0; k < 3; k++) t[3 + k] += b->avel[k] * hinv;
        }

        /* stage 2c: rhs = c/h - J tmp1 */
        for (int i = 0; i < m; i++) {
            const real *J = rows[i].J;
            const real *t = tmp1 + 6 * jb[2 * i];
            real s = 0;
            for (int k = 0; k < 6; k++) s += J[k] * t[k];
            if (jb[2 * i + 1] >= 0) {
                t = tmp1 + 6 * jb[2 * i + 1];
                for (int k = 0; k < 6; k++) s += J[6 + k] * t[k];
            }
            rhs[i] -= s;
        }

        /* stage 3: SOR-LCP.  iMJ = inv(M) J^T ; Ad = w / (diag + cfm); scale J, rhs */
        for (int i = 0; i < m; i++) {
            real *J = rows[i].J;
            real *o = iMJ + 12 * i;
            const body_t *b0 = &w->bodies[bodies[jb[2 * i]]];
            for (int k = 0; k < 3; k++) o[k] = b0->invMass * J[k];
            mul_R_v(o + 3, b0->invI_w, J + 3);
            real sum = 0;
            for (int k = 0; k < 6; k++) sum += o[k] * J[k];
            if (jb[2 * i + 1] >= 0) {
                const body_t *b1 = &w->bodies[bodies[jb[2 * i + 1]]];
                for (int k = 0; k < 3; k++) o[6 + k] = b1->invMass * J[6 + k];
                mul_R_v(o + 9, b1->invI_w, J + 9);
                for (int k = 6; k < 12; k++) sum += o[k] * J[k];
            } else {
                for (int k = 6; k < 12; k++) o[k] = 0;
            }
            if (w->stepper == ORC_STEPPER_DENSE) {
                Ad[i] = 1; /* dWorldStep solves the unscaled system */
                Adcfm[i] = cfm[i];
                continue;
            }
            Ad[i] = w->sor_w / (sum + cfm[i]);
            int nk = jb[2 * i + 1] >= 0 ? 12 : 6;
            for (int k = 0; k < nk; k++) J[k] *= Ad[i];
            rhs[i] *= Ad[i];
            Adcfm[i] = Ad[i] * cfm[i];
        }

        if (w->stepper == ORC_STEPPER_DENSE) {
            /* dWorldStep (ODE step.cpp): A = J invM J^T + diag(cfm / h), then dSolveLCP (orc_lcp.c); the constraint
             * force invM J^T lambda goes where the SOR sweep accumulates it */
            double *A = (double *)calloc((size_t)m * m, sizeof(double));
            double *xs = (double *)malloc(sizeof(double) * (size_t)m), *bs = (double *)malloc(sizeof(double) * (size_t)m);
            double *ws = (double *)malloc(sizeof(double) * (size_t)m), *los = (double *)malloc(sizeof(double) * (size_t)m);
            double *his = (double *)malloc(sizeof(double) * (size_t)m);
            for (int i = 0; i < m; i++) {
                for (int k = 0; k < m; k++) {
                    double s = 0;
                    for (int si = 0; si < 2; si++) {
                        const int bi = jb[2 * i + si];
                        if (bi < 0) continue;
                        for (int sk = 0; sk < 2; sk++) {
                            if (jb[2 * k + sk] != bi) continue;
                            for (int q = 0; q < 6; q++) s += (double)rows[i].J[6 * si + q] * (double)iMJ[12 * k + 6 * sk + q];
                        }
                    }
                    A[(size_t)i * m + k] = s;
                }
                A[(size_t)i * m + i] += (double)cfm[i];
                bs[i] = (double)rhs[i];
                los[i] = (double)rows[i].lo;
                his[i] = (double)rows[i].hi;
            }
            /* symmetrise (the two triangles differ by rounding only) */
            for (int i = 0; i < m; i++)
                for (int k = i + 1; k < m; k++) {
                    const double a = 0.5 * (A[(size_t)i * m + k] + A[(size_t)k * m + i]);
                    A[(size_t)i * m + k] = a, A[(size_t)k * m + i] = a;
                }
            if (orc_solve_lcp(m, A, xs, bs, ws, los, his, findex)) w->last_lcp_error = 1;
            for (int i = 0; i < m; i++) {
                lambda[i] = (real)xs[i];
                /* the friction rows' bounds were fixed from the friction-less solution: hand them to the
                 * residual diagnostic below as plain bounds */
                rows[i].lo = (real)los[i];
                rows[i].hi = (real)his[i];
                findex[i] = -1;
                const real *im = iMJ + 12 * i;
                real *f1 = fc + 6 * jb[2 * i];
                for (int q = 0; q < 6; q++) f1[q] += lambda[i] * im[q];
                if (jb[2 * i + 1] >= 0) {
                    real *f2 = fc + 6 * jb[2 * i + 1];
                    for (int q = 0; q < 6; q++) f2[q] += lambda[i] * im[6 + q];
                }
            }
            free(A), free(xs), free(bs), free(ws), free(los), free(his);
        }

        /* order: rows with findex < 0 first, dependent rows last */
        int nhead = 0;
        for (int i = 0; i < m; i++)
            if (findex[i] < 0) nhead++;
        {
            int h0 = 0, t0 = nhead;
            for (int i = 0; i < m; i++) {
                if (findex[i] < 0)
                    order[h0++] = i;
                else
                    order[t0++] = i;
            }
        }

        for (int it = 0; it < (w->stepper == ORC_STEPPER_DENSE ? 0 : w->iters); it++) {
            if (w->reorder_method == ORC_REORDER_RANDOMLY && (it % 8) == 0 &&
                (it >= 8 || w->variant[ORC_VARIANT_SHUFFLE_AT_0])) {
                /* ODE ConstraintsReorderingHelper on both partitions */
                for (int idx = 1; idx < nhead; idx++) {
                    int sw = orc_rand_int(&w->seed, idx + 1);
                    int t = order[idx];
                    order[idx] = order[sw];
                    order[sw] = t;
                }
                for (int idx = nhead + 1; idx < m; idx++) {
                    int sw = orc_rand_int(&w->seed, idx - nhead + 1) + nhead;
                    int t = order[idx];
                    order[idx] = order[sw];
                    order[sw] = t;
                }
            }
            for (int k = 0; k < m; k++) {
                int i = order[k];
                const real *J = rows[i].J;
                real *f1 = fc + 6 * jb[2 * i];
                real *f2 = jb[2 * i + 1] >= 0 ? fc + 6 * jb[2 * i + 1] : 0;
                real old = lambda[i];
                real delta = rhs[i] - old * Adcfm[i];
                delta -= f1[0] * J[0] + f1[1] * J[1] + f1[2] * J[2] + f1[3] * J[3] + f1[4] * J[4] + f1[5] * J[5];
                if (f2)
                    delta -= f2[0] * J[6] + f2[1] * J[7] + f2[2] * J[8] + f2[3] * J[9] + f2[4] * J[10] +
                             f2[5] * J[11];
                real lo_act, hi_act;
                if (findex[i] >= 0) {
                    hi_act = R_FABS(rows[i].hi * lambda[findex[i]]);
                    lo_act = -hi_act;
                } else {
                    hi_act = rows[i].hi;
                    lo_act = rows[i].lo;
                }
                real nl = old + delta;
                if (nl < lo_act) {
                    delta = lo_act - old;
                    lambda[i] = lo_act;
                } else if (nl > hi_act) {
                    delta = hi_act - old;
                    lambda[i] = hi_act;
                } else {
                    lambda[i] = nl;
                }
                const real *im = iMJ + 12 * i;
                for (int q = 0; q < 6; q++) f1[q] += delta * im[q];
                if (f2)
                    for (int q = 0; q < 6; q++) f2[q] += delta * im[6 + q];
            }
        }

        /* residual diagnostic on the unscaled system: w = J invM J^T lambda + cfm lambda - rhs */
        {
            real maxres = w->last_residual;
            for (int i = 0; i < m; i++) {
                const real *J = Jcopy + 12 * i;
                const real *f1 = fc + 6 * jb[2 * i];
                real v = 0;
                for (int q = 0; q < 6; q++) v += J[q] * f1[q];
                if (jb[2 * i + 1] >= 0) {
                    const real *f2 = fc + 6 * jb[2 * i + 1];
                    for (int q = 0; q < 6; q++) v += J[6 + q] * f2[q];
                }
                real wi = v + cfm[i] * lambda[i] - rhs[i] / Ad[i];
                real lo_act = rows[i].lo, hi_act = rows[i].hi;
                if (findex[i] >= 0) {
                    hi_act = R_FABS(rows[i].hi * lambda[findex[i]]);
                    lo_act = -hi_act;
                }
                real res;
                if (lambda[i] <= lo_act && lo_act > -R_INF)
                    res = wi < 0 ? -wi : 0;
                else if (lambda[i] >= hi_act && hi_act < R_INF)
                    res = wi > 0 ? wi : 0;
                else
                    res = R_FABS(wi);
                if (lo_act == hi_act) res = 0;
                if (res > maxres) maxres = res;
            }
            w->last_residual = maxres;
        }

        /* joint feedback from the unscaled Jacobian copy */
        for (int i = 0; i < nj; i++) {
            joint_t *j = jl[i].j;
            real data[12] = {0};
            for (int k = 0; k < jl[i].m; k++) {
                int gi = jl[i].ofs + k;
                for (int q = 0; q < 12; q++) data[q] += Jcopy[12 * gi + q] * lambda[gi];
            }
            memcpy(j->fb_f1, data, 3 * sizeof(real));
            memcpy(j->fb_t1, data + 3, 3 * sizeof(real));
            if (j->b[1] >= 0) {
                memcpy(j->fb_f2, data + 6, 3 * sizeof(real));
                memcpy(j->fb_t2, data + 9, 3 * sizeof(real));
            } else {
                memset(j->fb_f2, 0, 3 * sizeof(real));
                memset(j->fb_t2, 0, 3 * sizeof(real));
            }
            /* `lambda` of the DFKI-patched dJointFeedback (read at src/joints/Joint.cpp:900):
             * multiplier of the joint's limit/motor row (assumption) */
            j->fb_lambda = j->limot_row >= 0 ? lambda[jl[i].ofs + j->limot_row] : 0;
            if (j->limot_row >= 0 && w->variant[ORC_VARIANT_LAMBDA_ROW] == 1) j->fb_lambda = lambda[jl[i].ofs + jl[i].m - 1];
        }

        /* v += h * cforce */
        for (int i = 0; i < nb; i++) {
            body_t *b = &w->bodies[bodies[i]];
            for (int k = 0; k < 3; k++) {
                b->lvel[k] += h * fc[6 * i + k];
                b->avel[k] += h * fc[6 * i + 3 + k];
            }
        }
        free(jb);
        free(Jcopy);
        free(iMJ);
        free(rhs);
        free(cfm);
        free(Adcfm);
        free(Ad);
        free(findex);
        free(order);
        free(tmp1);
        free(rows);
        free(lambda);
        free(fc);
    }

    /* joints that produced no rows report zero feedback */
    /* v += h * invM * fe ; then integrate positions */
    for (int i = 0; i < nb; i++) {
        body_t *b = &w->bodies[bodies[i]];
        real t[3];
        for (int k = 0; k < 3; k++) {
            b->lvel[k] += h * b->invMass * b->facc[k];
            b->tacc[k] *= h;
        }
        mul_R_v(t, b->invI_w, b->tacc);
        for (int k = 0; k < 3; k++) b->avel[k] += t[k];
    }
    for (int i = 0; i < nb; i++) {
        body_t *b = &w->bodies[bodies[i]];
        step_body(b, h);
        for (int k = 0; k < 3; k++) {
            b->facc[k] = 0;
            b->tacc[k] = 0;
        }
    }
}

/* all joints touching body `bi`, most recently attached first (ODE prepends in dJointAttach) */
static int joints_of_body(orc_world *w, int bi, joint_t **out)
{
    int n = 0;
    for (int i = 0; i < w->nj; i++)
        if (w->joints[i].b[0] == bi || w->joints[i].b[1] == bi) out[n++] = &w->joints[i];
    for (int i = 0; i < w->nc; i++)
        if (w->contacts[i].b[0] == bi || w->contacts[i].b[1] == bi) out[n++] = &w->contacts[i];
    /* sort by creation_serial descending (insertion sort, small n) */
    for (int i = 1; i < n; i++) {
        joint_t *t = out[i];
        int k = i - 1;
        while (k >= 0 && out[k]->creation_serial < t->creation_serial) {
            out[k + 1] = out[k];
            k--;
        }
        out[k + 1] = t;
    }
    return n;
}

void orc_world_quickstep(orc_world *w, double hd)
{
    const real h = (real)hd;
    w->last_m = 0;
    w->last_islands = 0;
    w->last_residual = 0;
    w->last_lcp_error = 0;
    int total_j = w->nj + w->nc;
    int *bodies = (int *)malloc(sizeof(int) * (size_t)(w->nb > 0 ? w->nb : 1));
    jref_t *jl = (jref_t *)malloc(sizeof(jref_t) * (size_t)(total_j > 0 ? total_j : 1));

    if (w->order_mode == ORC_ORDER_WORLD_BLOCK) {
        for (int i = 0; i < w->nb; i++) bodies[i] = i;
        int n = 0;
        for (int i = 0; i < w->nj; i++)
            if (w->joints[i].b[0] >= 0) jl[n++].j = &w->joints[i];
        for (int i = 0; i < w->nc; i++)
            if (w->contacts[i].b[0] >= 0) jl[n++].j = &w->contacts[i];
        w->last_islands = 1;
        quickstep_island(w, bodies, w->nb, jl, n, h);
    } else {
        /* ODE util.cpp dxProcessIslands: DFS from the world's body list (most recently created
         * first), each body's joint list most-recently-attached first, LIFO stack of bodies */
        for (int i = 0; i < w->nb; i++) w->bodies[i].tag = -1;
        for (int i = 0; i < w->nj; i++) w->joints[i].tag = 0;
        for (int i = 0; i < w->nc; i++) w->contacts[i].tag = 0;
        int *stack = (int *)malloc(sizeof(int) * (size_t)(w->nb > 0 ? w->nb : 1));
        joint_t **jtmp = (joint_t **)malloc(sizeof(joint_t *) * (size_t)(total_j > 0 ? total_j : 1));
        int *visited = (int *)calloc((size_t)(w->nb > 0 ? w->nb : 1), sizeof(int));
        for (int bb = w->nb - 1; bb >= 0; bb--) {
            if (visited[bb]) continue;
            visited[bb] = 1;
            int nbi = 0, nji = 0, sp = 0;
            bodies[nbi++] = bb;
            int b = bb;
            for (;;) {
                int n = joints_of_body(w, b, jtmp);
                for (int k = 0; k < n; k++) {
                    joint_t *j = jtmp[k];
                    if (j->tag) continue;
                    j->tag = 1;
                    jl[nji++].j = j;
                    int other = (j->b[0] == b) ? j->b[1] : j->b[0];
                    if (other >= 0 && !visited[other]) {
                        visited[other] = 1;
                        stack[sp++] = other;
                    }
                }
                if (sp == 0) break;
                b = stack[--sp];
                bodies[nbi++] = b;
            }
            w->last_islands++;
            quickstep_island(w, bodies, nbi, jl, nji, h);
        }
        free(stack);
        free(jtmp);
        free(visited);
    }
    free(bodies);
    free(jl);
}

/* ------------------------------------------------------------------ */
/* mass (ODE mass.cpp); computed in double regardless of `real`        */

void orc_mass_set_zero(orc_mass *m) { memset(m, 0, sizeof(*m)); }

void orc_mass_adjust(orc_mass *m, double newmass)
{
    double scale = newmass / m->mass;
    m->mass = newmass;
    for (int i = 0; i < 9; i++) m->I[i] *= scale;
}

void orc_mass_set_box_total(orc_mass *m, double total, double lx, double ly, double lz)
{
    orc_mass_set_zero(m);
    m->mass = total;
    m->I[0] = total / 12.0 * (ly * ly + lz * lz);
    m->I[4] = total / 12.0 * (lx * lx + lz * lz);
    m->I[8] = total / 12.0 * (lx * lx + ly * ly);
}

void orc_mass_set_box(orc_mass *m, double density, double lx, double ly, double lz)
{
    orc_mass_set_box_total(m, lx * ly * lz * density, lx, ly, lz);
}

void orc_mass_set_sphere_total(orc_mass *m, double total, double r)
{
    orc_mass_set_zero(m);
    m->mass = total;
    double II = 0.4 * total * r * r;
    m->I[0] = m->I[4] = m->I[8] = II;
}

void orc_mass_set_sphere(orc_mass *m, double density, double r)
{
    orc_mass_set_sphere_total(m, (4.0 / 3.0) * M_PI * r * r * r * density, r);
}

void orc_mass_set_capsule(orc_mass *m, double density, int dir, double r, double len)
{
    orc_mass_set_zero(m);
    double M1 = M_PI * r * r * len * density;
    double M2 = (4.0 / 3.0) * M_PI * r * r * r * density;
    m->mass = M1 + M2;
    double Ia = M1 * (0.25 * r * r + (1.0 / 12.0) * len * len) + M2 * (0.4 * r * r + 0.375 * r * len + 0.25 * len * len);
    double Ib = (M1 * 0.5 + M2 * 0.4) * r * r;
    m->I[0] = m->I[4] = m->I[8] = Ia;
    m->I[(dir - 1) * 4] = Ib;
}

void orc_mass_set_capsule_total(orc_mass *m, double total, int dir, double r, double len)
{
    orc_mass_set_capsule(m, 1.0, dir, r, len);
    orc_mass_adjust(m, total);
}

void orc_mass_set_cylinder(orc_mass *m, double density, int dir, double r, double len)
{
    orc_mass_set_cylinder_total(m, M_PI * r * r * len * density, dir, r, len);
}

void orc_mass_set_cylinder_total(orc_mass *m, double total, int dir, double r, double len)
{
    orc_mass_set_zero(m);
    m->mass = total;
    double Ia = total * (0.25 * r * r + (1.0 / 12.0) * len * len);
    double Ib = total * 0.5 * r * r;
    m->I[0] = m->I[4] = m->I[8] = Ia;
    m->I[(dir - 1) * 4] = Ib;
}

void orc_mass_rotate(orc_mass *m, const double R[9])
{
    double t[9], I2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[i * 3 + j] = m->I[i * 3] * R[j * 3] + m->I[i * 3 + 1] * R[j * 3 + 1] + m->I[i * 3 + 2] * R[j * 3 + 2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) I2[i * 3 + j] = R[i * 3] * t[j] + R[i * 3 + 1] * t[3 + j] + R[i * 3 + 2] * t[6 + j];
    memcpy(m->I, I2, sizeof(I2));
    m->I[3] = m->I[1];
    m->I[6] = m->I[2];
    m->I[7] = m->I[5];
    double c[3];
    for (int i = 0; i < 3; i++) c[i] = R[i * 3] * m->c[0] + R[i * 3 + 1] * m->c[1] + R[i * 3 + 2] * m->c[2];
    memcpy(m->c, c, sizeof(c));
}

static void crossmat_sq(const double a[3], double out[9])
{
    /* (crossmat(a))^2 = a a^T - |a|^2 I */
    double n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[i * 3 + j] = a[i] * a[j] - (i == j ? n2 : 0.0);
}

void orc_mass_translate(orc_mass *m, double x, double y, double z)
{
    double a[3] = {x + m->c[0], y + m->c[1], z + m->c[2]};
    double t1[9], t2[9];
    crossmat_sq(a, t1);
    crossmat_sq(m->c, t2);
    for (int i = 0; i < 9; i++) m->I[i] += m->mass * (t2[i] - t1[i]);
    m->I[3] = m->I[1];
    m->I[6] = m->I[2];
    m->I[7] = m->I[5];
    m->c[0] += x;
    m->c[1] += y;
    m->c[2] += z;
}

void orc_mass_add(orc_mass *a, const orc_mass *b)
{
    double denom = 1.0 / (a->mass + b->mass);
    for (int i = 0; i < 3; i++) a->c[i] = (a->c[i] * a->mass + b->c[i] * b->mass) * denom;
    a->mass += b->mass;
    for (int i = 0; i < 9; i++) a->I[i] += b->I[i];
}

int orc_mass_check(const orc_mass *m)
{
    if (m->mass <= 0) return 0;
    /* positive definiteness via leading principal minors */
    const double *I = m->I;
    double d1 = I[0];
    double d2 = I[0] * I[4] - I[1] * I[3];
    double d3 = I[0] * (I[4] * I[8] - I[5] * I[7]) - I[1] * (I[3] * I[8] - I[5] * I[6]) + I[2] * (I[3] * I[7] - I[4] * I[6]);
    if (d1 <= 0 || d2 <= 0 || d3 <= 0) return 0;
    return 1;
}

void orc_plane_space(const double n[3], double p[3], double q[3])
{
    real rn[3] = {(real)n[0], (real)n[1], (real)n[2]}, rp[3], rq[3];
    plane_space(rn, rp, rq);
    GET3(p, rp);
    GET3(q, rq);
}

void orc_q_to_R(const double q[4], double R[9])
{
    real rq[4] = {(real)q[0], (real)q[1], (real)q[2], (real)q[3]}, rR[9];
    q_to_R(rq, rR);
    for (int i = 0; i < 9; i++) R[i] = rR[i];
}

/* ------------------------------------------------------------------ */
/* bench helper: collide + quickstep `steps` times over `n` worlds (one call = one thread's share) */
int orc_collide(orc_world *w);
void orc_bench_run(orc_world **worlds, int n, int steps, double h, int do_collide)
{
    for (int s = 0; s < steps; s++)
        for (int i = 0; i < n; i++) {
            if (do_collide) orc_collide(worlds[i]);
            orc_world_quickstep(worlds[i], h);
        }
}

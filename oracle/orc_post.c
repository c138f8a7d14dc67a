/*
 * oracle/orc_post.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * What WorldPhysics::stepTheWorld does after the ODE call (reference src/WorldPhysics.cpp:372-397), restated on
 * the body / joint ids of an orc_world so that the device epilogue (b2p_integrate.cu), which does it for every
 * world of a batch, can be checked world by world:
 *   Joint::update              src/joints/Joint.cpp:859-1019
 *   Joint::applyJointFriction  src/joints/Joint.cpp:1031-1063
 *   Frame::update              src/objects/Frame.cpp:899-943
 *   Frame::computeContactForce src/objects/Frame.cpp:159-171
 * orc_mars.c (the reference's Frame / Joint objects of ONE world) calls the same functions.
 * PARITY UNPINNED (see orc.h).
 */
#include <math.h>
#include <string.h>

#include "orc_internal.h"

static void p_cross(double *r, const double *a, const double *b)
{
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x;
    r[1] = y;
    r[2] = z;
}
static double p_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double p_len(const double *a) { return sqrt(p_dot(a, a)); }

/* Joint::update :859-1019.  out = motor_torque, axis1_torque[3], joint_load[3].  The reference's body1 is the
 * first body handed to dJointAttach: node[0].body unless the joint is reversed (then body1 is null and the
 * "else if (body2)" arm runs, reading feedback.f2 / t2 -- which ODE leaves zero for a one-body joint). */
void orc_post_joint_update(const orc_world *w, int ji, double out[7])
{
    const joint_t *j = &w->joints[ji];
    double fb[13], anchor[3] = {0, 0, 0}, axis[3] = {0, 0, 0};
    orc_joint_get_feedback(w, ji, fb);
    const double *f1 = fb, *t1 = fb + 3, *f2 = fb + 6, *t2 = fb + 9;
    int calc1 = 0;
    if (j->type == ORC_JOINT_HINGE || j->type == ORC_JOINT_HINGE2 || j->type == ORC_JOINT_UNIVERSAL) {
        orc_joint_get_anchor(w, ji, anchor);
        orc_joint_get_axis1(w, ji, axis);
        calc1 = 1;
    }
    double *axis1_torque = out + 1, *joint_load = out + 4;
    out[0] = -fb[12];
    memset(out + 1, 0, 6 * sizeof(double));
    if (calc1 == 1 && j->b[0] >= 0) {
        double v1[3], normal[3], load[3], tmp1[3], bp[3];
        orc_body_get_position(w, j->b[0], bp);
        for (int i = 0; i < 3; i++) v1[i] = bp[i] - anchor[i];
        if (!j->reverse) {
            p_cross(normal, axis, v1);
            double dot = p_dot(normal, f1);
            for (int i = 0; i < 3; i++) normal[i] *= dot;
            for (int i = 0; i < 3; i++) load[i] = f1[i] - normal[i];
            p_cross(tmp1, v1, normal);
            memcpy(axis1_torque, tmp1, sizeof(tmp1));
            p_cross(tmp1, v1, load);
            memcpy(joint_load, tmp1, sizeof(tmp1));
            dot = p_dot(axis, t1);
            for (int i = 0; i < 3; i++) {
                tmp1[i] = axis[i] * dot;
                load[i] = t1[i] - tmp1[i];
                axis1_torque[i] += tmp1[i];
                joint_load[i] += load[i];
            }
        } else {
            double axis_force[3];
            double dot = p_dot(f2, v1) / p_dot(v1, v1);
            for (int i = 0; i < 3; i++) {
                axis_force[i] = v1[i] * dot;
                tmp1[i] = f2[i] - axis_force[i];
            }
            double torque = p_len(tmp1) / p_len(v1);
            p_cross(normal, v1, tmp1);
            double nl = p_len(normal);
            for (int i = 0; i < 3; i++) normal[i] *= torque / nl;
            dot = p_dot(normal, axis);
            for (int i = 0; i < 3; i++) {
                tmp1[i] = axis[i] * dot;
                load[i] = normal[i] - tmp1[i];
            }
            memcpy(axis1_torque, tmp1, sizeof(tmp1));
            memcpy(joint_load, load, sizeof(load));
            dot = p_dot(t2, axis);
            for (int i = 0; i < 3; i++) {
                tmp1[i] = axis[i] * dot;
                load[i] = t2[i] - tmp1[i];
                joint_load[i] += load[i];
            }
        }
    }
}

/* Joint::applyJointFriction :1031-1063 (off unless the coefficient is > 0): an external force / torque on both
 * bodies, accumulated for the NEXT step.  getVelocityI (:1078-1102) negates hinge and slider rates. */
void orc_post_joint_friction(orc_world *w, int ji, double coeff)
{
    if (!(coeff > 0.0)) return;
    const joint_t *j = &w->joints[ji];
    if (j->b[0] < 0) return;
    double axis[3] = {0, 0, 0}, vel = 0;
    if (j->type == ORC_JOINT_HINGE || j->type == ORC_JOINT_SLIDER) {
        orc_joint_get_axis1(w, ji, axis);
        vel = -orc_joint_get_angle1_rate(w, ji);
    } else if (j->type == ORC_JOINT_HINGE2 || j->type == ORC_JOINT_UNIVERSAL) {
        orc_joint_get_axis1(w, ji, axis);
        vel = orc_joint_get_angle1_rate(w, ji);
    }
    double damping_torque = vel * -coeff;
    for (int k = 0; k < 2; k++) {
        int body = j->b[k];
        if (body < 0) continue;
        double pos[3];
        orc_body_get_position(w, body, pos);
        for (int i = 0; i < 3; i++) pos[i] -= axis[i];
        /* angleAxisToQuaternion(dampingTorque, axis) applied to pos (Rodrigues) */
        double n = p_len(axis), u[3] = {0, 0, 0};
        if (n > 0)
            for (int i = 0; i < 3; i++) u[i] = axis[i] / n;
        double c = cos(damping_torque), s = sin(damping_torque), ux[3];
        p_cross(ux, u, pos);
        double ud = p_dot(u, pos), force[3], torque[3];
        for (int i = 0; i < 3; i++) {
            double rot = pos[i] * c + ux[i] * s + u[i] * ud * (1 - c);
            force[i] = rot - pos[i];
            torque[i] = u[i] * damping_torque;
        }
        orc_body_add_force(w, body, force[0], force[1], force[2]);
        orc_body_add_torque(w, body, torque[0], torque[1], torque[2]);
    }
}

/* Frame::update :899-943.  last6 = last_l_vel, last_a_vel (in / out); acc6 = l_acc, a_acc (out). */
void orc_post_frame_update(orc_world *w, int b, double dt, double lin_damp, double ang_damp, double last6[6],
                           double acc6[6])
{
    double v[3];
    if (dt > 0) {
        orc_body_get_linear_vel(w, b, v);
        for (int i = 0; i < 3; i++) {
            acc6[i] = (v[i] - last6[i]) / dt;
            last6[i] = v[i];
        }
        orc_body_get_angular_vel(w, b, v);
        for (int i = 0; i < 3; i++) {
            acc6[3 + i] = (v[i] - last6[3 + i]) / dt;
            last6[3 + i] = v[i];
        }
    } else {
        memset(acc6, 0, 6 * sizeof(double));
        memset(last6, 0, 6 * sizeof(double));
    }
    if (lin_damp < 1.0) {
        orc_body_get_linear_vel(w, b, v);
        orc_body_set_linear_vel(w, b, v[0] * lin_damp, v[1] * lin_damp, v[2] * lin_damp);
    }
    if (ang_damp < 1.0) {
        orc_body_get_angular_vel(w, b, v);
        orc_body_set_angular_vel(w, b, v[0] * ang_damp, v[1] * ang_damp, v[2] * ang_damp);
    }
}

/* Frame::computeContactForce :159-171 in batch form: the feedback f1 of every contact joint whose first body is
 * `b` (a frame's contacts are the ones created through ITS addContact, i.e. with its body first) */
void orc_post_body_contact_force(const orc_world *w, int b, double out[3])
{
    out[0] = out[1] = out[2] = 0;
    for (int c = 0; c < w->nc; c++) {
        if (w->contacts[c].b[0] != b) continue;
        double fb[12];
        orc_contact_get_feedback(w, c, fb);
        out[0] += fb[0];
        out[1] += fb[1];
        out[2] += fb[2];
    }
}

/* The whole post-step sequence of stepTheWorld for one world, in the reference's order (:372-397): every joint's
 * update (+ friction), then every frame's update, then the contact force sums.  Arrays are per body / joint of
 * the world; `last` (nb*6) carries the velocity history between calls. */
void orc_post_step_world(orc_world *w, double dt, const double *lin_damp, const double *ang_damp,
                         const double *friction, double *last, double *acc, double *cforce, double *jupd)
{
    for (int j = 0; j < w->nj; j++) {
        orc_post_joint_update(w, j, jupd + 7 * j);
        orc_post_joint_friction(w, j, friction[j]);
    }
    for (int b = 0; b < w->nb; b++) orc_post_frame_update(w, b, dt, lin_damp[b], ang_damp[b], last + 6 * b, acc + 6 * b);
    for (int b = 0; b < w->nb; b++) orc_post_body_contact_force(w, b, cforce + 3 * b);
}

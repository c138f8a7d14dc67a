/*
 * oracle/orc_mars.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Restatement of the reference-side host logic that surrounds the ODE step:
 *   WorldPhysics  (reference src/WorldPhysics.cpp:80-120,162-203,286-421,458-468)
 *   Frame         (reference src/objects/Frame.cpp:183-207,233-313,519-622,794-955,996-1092)
 *   Joint + types (reference src/joints/Joint.cpp:147-273,315-415,618-703,859-1102,
 *                  HingeJoint.cpp:31-66, FixedJoint.cpp:31-47, SliderJoint.cpp:32-65,
 *                  Hinge2Joint.cpp:33-77)
 * on top of the orc_* quickstep oracle, plus the literal scene of
 * test/reproducable_test.cpp.  PARITY UNPINNED (see orc.h).
 */
#ifndef ORC_MARS_H
#define ORC_MARS_H

#include "orc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mars orc_mars;

/* interfaces::ContactData as consumed by Frame::addContact */
typedef struct orc_mars_contact {
    int frame1, frame2; /* frame ids, -1 = nullptr */
    double pos[3], normal[3], depth;
    double cfm, erp, friction1, friction2;
    double rolling_friction, rolling_friction2, spinning_friction;
    double bounce, bounce_vel, motion1, motion2, fds1, fds2;
} orc_mars_contact;

typedef struct orc_mars_joint_cfg {
    int type;       /* ORC_JOINT_* */
    int parent, child; /* frame ids or -1 */
    double anchor[3], axis1[3], axis2[3];
    int has_spring_damping, has_spring_stiffness, has_joint_cfm, has_friction;
    double spring_damping, spring_stiffness, joint_cfm, friction_coefficient;
} orc_mars_joint_cfg;

orc_mars *orc_mars_create(void); /* ctor defaults; world not initialised yet */
void orc_mars_destroy(orc_mars *m);
void orc_mars_init_the_world(orc_mars *m);
orc_world *orc_mars_world(orc_mars *m);
void orc_mars_set_step_size(orc_mars *m, double h);
double orc_mars_get_step_size(const orc_mars *m);
void orc_mars_set_world_params(orc_mars *m, const double gravity[3], double cfm, double erp);
void orc_mars_clear_previous_step(orc_mars *m);
void orc_mars_step_the_world(orc_mars *m);

int orc_mars_create_frame(orc_mars *m, double linear_damping, double angular_damping);
/* Frame::addObject + addObjectMass; q is (w,x,y,z) */
void orc_mars_frame_add_object(orc_mars *m, int f, const orc_mass *mass, const double pos[3], const double q[4]);
int orc_mars_frame_body(const orc_mars *m, int f); /* -1 if none */
int orc_mars_frame_acquire_body(orc_mars *m, int f);
void orc_mars_frame_get_position(const orc_mars *m, int f, double out[3]);
void orc_mars_frame_set_position(orc_mars *m, int f, const double p[3]);
void orc_mars_frame_get_rotation(const orc_mars *m, int f, double out_wxyz[4]);
void orc_mars_frame_set_rotation(orc_mars *m, int f, const double q_wxyz[4]);
void orc_mars_frame_get_linear_acceleration(const orc_mars *m, int f, double out[3]);
void orc_mars_frame_get_contact_force(const orc_mars *m, int f, double out[3]);
void orc_mars_frame_get_mass(const orc_mars *m, int f, double *mass, double I[9]);
/* returns 1 if a contact joint was created */
int orc_mars_frame_add_contact(orc_mars *m, int f, const orc_mars_contact *c);

int orc_mars_create_joint(orc_mars *m, const orc_mars_joint_cfg *cfg); /* -1 on failure */
void orc_mars_joint_set_force_limit(orc_mars *m, int j, double v);
void orc_mars_joint_set_force_limit2(orc_mars *m, int j, double v);
void orc_mars_joint_set_velocity(orc_mars *m, int j, double v);
void orc_mars_joint_set_velocity2(orc_mars *m, int j, double v);
void orc_mars_joint_set_joint_as_motor(orc_mars *m, int j, int axis);
void orc_mars_joint_unset_joint_as_motor(orc_mars *m, int j, int axis);
double orc_mars_joint_get_position(const orc_mars *m, int j);
double orc_mars_joint_get_velocity(const orc_mars *m, int j);
double orc_mars_joint_get_velocity2(const orc_mars *m, int j);
double orc_mars_joint_get_motor_torque(const orc_mars *m, int j);
void orc_mars_joint_get_axis_torque(const orc_mars *m, int j, double out[3]);
void orc_mars_joint_get_joint_load(const orc_mars *m, int j, double out[3]);
int orc_mars_joint_ode_id(const orc_mars *m, int j);

/* the literal test/reproducable_test.cpp scene: builds it into `m` (which must be fresh),
 * fills frame ids (13: body, 8 links, 4 feet) and joint ids (12) */
void orc_mars_build_reproducable_scene(orc_mars *m, int frames[13], int joints[12]);
/* one iteration of the test loop body (control, contacts, step); returns #contacts created */
int orc_mars_reproducable_tick(orc_mars *m, const int frames[13], const int joints[12], double time);
/* whole run; writes the "%g,%g,%g,%g" CSV if path != NULL; returns number of rows written */
int orc_mars_run_reproducable_test(const char *csv_path, double t_end, int order_mode, double *last_pos);
int orc_mars_run_reproducable_test_stepper(const char *csv_path, double t_end, int order_mode, int stepper,
                                           double *last_pos);

#ifdef __cplusplus
}
#endif
#endif

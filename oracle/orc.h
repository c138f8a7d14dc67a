/*
 * oracle/orc.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C restatement of the arithmetic that sits behind
 * WorldPhysics::stepTheWorld in mars_ode_physics
 * (reference: src/WorldPhysics.cpp:349-421 -> dWorldQuickStep at :365).
 *
 * The arithmetic itself lives in a third-party dependency that is NOT under
 * /root/reference: ODE ("simulation/ode-16", manifest.xml:15; pkg-config module
 * `ode` with no version pin, CMakeLists.txt:24-25).  This file restates the
 * published ODE 0.16 quickstep algorithm (ode/src/quickstep.cpp, util.cpp,
 * joints/{contact,hinge,hinge2,slider,fixed,joint}.cpp, mass.cpp, rotation.cpp,
 * misc.cpp) from its documented behaviour.
 *
 * PARITY UNPINNED: the reference holds no golden vectors for this path
 * (test/reproducable_test.cpp only checks run-to-run determinism) and libode
 * cannot be built or found in this environment, so this oracle is anchored on
 * the reference's call sites and scene only.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.
 */
#ifndef ORC_H
#define ORC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* joint types (values follow mars::interfaces::JointType order used by
 * src/joints/Joint.cpp switch statements; CONTACT is internal) */
enum {
    ORC_JOINT_HINGE = 1,
    ORC_JOINT_HINGE2 = 2,
    ORC_JOINT_SLIDER = 3,
    ORC_JOINT_BALL = 4,
    ORC_JOINT_UNIVERSAL = 5,
    ORC_JOINT_FIXED = 6,
    ORC_JOINT_CONTACT = 100
};

/* limit/motor parameters -- same names as ODE's dParam* used at
 * src/joints/Joint.cpp:315-374,618-703,1301-1464 ; second axis = +0x100 */
enum {
    ORC_PARAM_LO_STOP = 0,
    ORC_PARAM_HI_STOP,
    ORC_PARAM_VEL,
    ORC_PARAM_LO_VEL,
    ORC_PARAM_HI_VEL,
    ORC_PARAM_FMAX,
    ORC_PARAM_FUDGE_FACTOR,
    ORC_PARAM_BOUNCE,
    ORC_PARAM_CFM,
    ORC_PARAM_STOP_ERP,
    ORC_PARAM_STOP_CFM,
    ORC_PARAM_SUSPENSION_ERP,
    ORC_PARAM_SUSPENSION_CFM,
    ORC_PARAM_ERP,
    ORC_PARAM_GROUP2 = 0x100
};

/* contact surface mode bits (ODE contact.h) -- reference sets
 * SoftERP|SoftCFM|Approx1 (+Rolling, +Mu2): src/objects/Frame.cpp:1038-1079 */
enum {
    ORC_CONTACT_MU2 = 0x001,
    ORC_CONTACT_FDIR1 = 0x002,
    ORC_CONTACT_BOUNCE = 0x004,
    ORC_CONTACT_SOFT_ERP = 0x008,
    ORC_CONTACT_SOFT_CFM = 0x010,
    ORC_CONTACT_MOTION1 = 0x020,
    ORC_CONTACT_MOTION2 = 0x040,
    ORC_CONTACT_MOTIONN = 0x080,
    ORC_CONTACT_SLIP1 = 0x100,
    ORC_CONTACT_SLIP2 = 0x200,
    ORC_CONTACT_ROLLING = 0x400,
    ORC_CONTACT_APPROX1_1 = 0x1000,
    ORC_CONTACT_APPROX1_2 = 0x2000,
    ORC_CONTACT_APPROX1_N = 0x4000,
    ORC_CONTACT_APPROX1 = 0x7000
};

/* row ordering of the SOR sweep */
enum {
    ORC_ORDER_WORLD_BLOCK = 0, /* one block per world, joints in creation order then contacts */
    ORC_ORDER_ODE_ISLANDS = 1  /* ODE island DFS order, one SOR problem per island */
};

/* reordering method inside the sweep (ODE CONSTRAINTS_REORDERING_METHOD) */
enum {
    ORC_REORDER_RANDOMLY = 0, /* ODE 0.16 default: shuffle at iterations 8,16,... */
    ORC_REORDER_NONE = 1      /* REORDERING_METHOD__DONT_REORDER */
};

/* which ODE stepper orc_world_quickstep runs: dWorldQuickStep (SOR-LCP, fast_step == true,
 * src/WorldPhysics.cpp:365) or dWorldStep (dense LCP solved by Dantzig's method, fast_step == false,
 * src/WorldPhysics.cpp:369; oracle/orc_lcp.c) */
enum {
    ORC_STEPPER_QUICK = 0,
    ORC_STEPPER_DENSE = 1
};

/* Restatement choices: details of ODE 0.16 that this oracle restates from memory and that only a run of the real
 * libode can settle.  Value 0 is what the oracle (and the device path) does; the other
 * values are the plausible alternatives, so that the day a libode is available the pin is one run per switch.
 * tests/test_oracle.py::test_restatement_variants quantifies what each one changes. */
enum {
    ORC_VARIANT_GYRO = 0,         /* 0: implicit (Lacoursiere) gyroscopic torque; 1: explicit -w x (I w); 2: none */
    ORC_VARIANT_SHUFFLE_AT_0 = 1, /* 0: reshuffles at iterations 8, 16, ...; 1: also at iteration 0 (ODE <= 0.13) */
    ORC_VARIANT_LAMBDA_ROW = 2,   /* feedback.lambda (patched dJointFeedback): 0: first limit/motor row; 1: last */
    ORC_NUM_VARIANTS = 3
};
struct orc_world;
void orc_world_set_variant(struct orc_world *w, int which, int value);
void orc_world_set_stepper(struct orc_world *w, int stepper);
int orc_world_last_lcp_error(const struct orc_world *w);

/* the dContact record as filled by Frame::addContact (Frame.cpp:1023-1079) */
typedef struct orc_contact_desc {
    double pos[3];
    double normal[3];
    double depth;
    int mode;
    double mu, mu2;
    double rho, rho2, rhoN;
    double bounce, bounce_vel;
    double soft_erp, soft_cfm;
    double motion1, motion2, motionN;
    double slip1, slip2;
} orc_contact_desc;

typedef struct orc_mass {
    double mass;
    double c[3];
    double I[9]; /* row-major 3x3 */
} orc_mass;

typedef struct orc_world orc_world;

int orc_real_size(void); /* sizeof(real): 8 (f64 build) or 4 (f32 build) */

orc_world *orc_world_create(void);
void orc_world_destroy(orc_world *w);
void orc_world_set_gravity(orc_world *w, double x, double y, double z);
void orc_world_set_erp(orc_world *w, double erp);
void orc_world_set_cfm(orc_world *w, double cfm);
void orc_world_set_max_angular_speed(orc_world *w, double s);
void orc_world_set_contact_max_correcting_vel(orc_world *w, double v);
void orc_world_set_contact_surface_layer(orc_world *w, double d);
void orc_world_set_quickstep_num_iterations(orc_world *w, int n);
void orc_world_set_quickstep_w(orc_world *w, double sor_w);
void orc_world_set_order_mode(orc_world *w, int order_mode);
void orc_world_set_reorder_method(orc_world *w, int method);
void orc_world_rand_set_seed(orc_world *w, uint32_t seed);
uint32_t orc_world_rand_get_seed(const orc_world *w);
void orc_world_quickstep(orc_world *w, double h);

/* diagnostics of the last step */
int orc_world_last_num_rows(const orc_world *w);
int orc_world_last_num_islands(const orc_world *w);
double orc_world_last_residual(const orc_world *w); /* max LCP residual (velocity units) */
int orc_world_num_bodies(const orc_world *w);
int orc_world_num_joints(const orc_world *w);
int orc_world_num_contacts(const orc_world *w);

/* RNG (ODE misc.cpp dRand/dRandInt), exposed for known-answer tests */
uint32_t orc_rand_next(uint32_t *seed);
int orc_rand_int(uint32_t *seed, int n);

/* bodies */
int orc_body_create(orc_world *w);
void orc_body_set_mass(orc_world *w, int b, double mass, const double I[9]);
void orc_body_get_mass(const orc_world *w, int b, double *mass, double I[9]);
void orc_body_set_position(orc_world *w, int b, double x, double y, double z);
void orc_body_set_quaternion(orc_world *w, int b, double qw, double qx, double qy, double qz);
void orc_body_set_linear_vel(orc_world *w, int b, double x, double y, double z);
void orc_body_set_angular_vel(orc_world *w, int b, double x, double y, double z);
void orc_body_set_force(orc_world *w, int b, double x, double y, double z);
void orc_body_set_torque(orc_world *w, int b, double x, double y, double z);
void orc_body_add_force(orc_world *w, int b, double x, double y, double z);
void orc_body_add_torque(orc_world *w, int b, double x, double y, double z);
void orc_body_add_force_at_pos(orc_world *w, int b, double fx, double fy, double fz, double px,
                               double py, double pz);
void orc_body_set_gravity_mode(orc_world *w, int b, int on);
void orc_body_set_gyroscopic_mode(orc_world *w, int b, int on);
void orc_body_get_position(const orc_world *w, int b, double out[3]);
void orc_body_get_quaternion(const orc_world *w, int b, double out[4]);
void orc_body_get_rotation(const orc_world *w, int b, double out[9]);
void orc_body_get_linear_vel(const orc_world *w, int b, double out[3]);
void orc_body_get_angular_vel(const orc_world *w, int b, double out[3]);
void orc_body_get_force(const orc_world *w, int b, double out[3]);
void orc_body_get_torque(const orc_world *w, int b, double out[3]);
/* pos3 q4 lvel3 avel3 = 13 doubles */
void orc_body_get_state(const orc_world *w, int b, double out[13]);
void orc_body_set_state(orc_world *w, int b, const double in[13]);

/* joints; b1/b2 = body index or -1 for the static environment */
int orc_joint_create(orc_world *w, int type, int b1, int b2);
void orc_joint_set_anchor(orc_world *w, int j, double x, double y, double z);
void orc_joint_set_axis1(orc_world *w, int j, double x, double y, double z);
void orc_joint_set_axis2(orc_world *w, int j, double x, double y, double z);
void orc_joint_set_fixed(orc_world *w, int j);
void orc_joint_set_param(orc_world *w, int j, int param, double value);
double orc_joint_get_param(const orc_world *w, int j, int param);
void orc_joint_get_anchor(const orc_world *w, int j, double out[3]);
void orc_joint_get_anchor2(const orc_world *w, int j, double out[3]);
void orc_joint_get_axis1(const orc_world *w, int j, double out[3]);
void orc_joint_get_axis2(const orc_world *w, int j, double out[3]);
double orc_joint_get_angle1(const orc_world *w, int j);      /* hinge angle / slider position / hinge2 angle1 */
double orc_joint_get_angle1_rate(const orc_world *w, int j);
double orc_joint_get_angle2(const orc_world *w, int j);      /* hinge2 angle2 */
double orc_joint_get_angle2_rate(const orc_world *w, int j);
void orc_joint_add_torque(orc_world *w, int j, double t);    /* dJointAddHingeTorque / dJointAddSliderForce */
/* f1[3] t1[3] f2[3] t2[3] lambda -> 13 doubles */
void orc_joint_get_feedback(const orc_world *w, int j, double out[13]);
int orc_joint_get_num_rows(const orc_world *w, int j); /* rows used in last step */
int orc_are_connected_excluding(const orc_world *w, int b1, int b2, int joint_type);

/* contact joint group */
void orc_contacts_clear(orc_world *w);
/* WorldPhysics::createContact (src/WorldPhysics.cpp:458-468): returns contact id, or -1 when
 * b1 and b2 are both bodies joined by a non-contact joint */
int orc_contact_add(orc_world *w, int b1, int b2, const orc_contact_desc *c);
void orc_contact_get_feedback(const orc_world *w, int c, double out[12]);

/* mass helpers (ODE mass.cpp) */
void orc_mass_set_zero(orc_mass *m);
void orc_mass_set_box_total(orc_mass *m, double total, double lx, double ly, double lz);
void orc_mass_set_box(orc_mass *m, double density, double lx, double ly, double lz);
void orc_mass_set_sphere_total(orc_mass *m, double total, double r);
void orc_mass_set_sphere(orc_mass *m, double density, double r);
void orc_mass_set_capsule_total(orc_mass *m, double total, int dir, double r, double len);
void orc_mass_set_capsule(orc_mass *m, double density, int dir, double r, double len);
void orc_mass_set_cylinder_total(orc_mass *m, double total, int dir, double r, double len);
void orc_mass_set_cylinder(orc_mass *m, double density, int dir, double r, double len);
void orc_mass_adjust(orc_mass *m, double newmass);
void orc_mass_rotate(orc_mass *m, const double R[9]);
void orc_mass_translate(orc_mass *m, double x, double y, double z);
void orc_mass_add(orc_mass *a, const orc_mass *b);
int orc_mass_check(const orc_mass *m);

/* ---- collision (orc_collide.c): defines the semantics the device kernels are checked against ---- */
typedef struct orc_surface {
    double mu, mu2;
    double rho, rho2, rhoN;
    double soft_erp, soft_cfm;
} orc_surface;
/* cls: 0 sphere(r) 1 box(lx,ly,lz) 2 capsule(r,len) 3 cylinder(r,len) 4 plane(a,b,c,d) 5 heightfield */
int orc_geom_create(orc_world *w, int cls, int body, const double dims[4], const double lpos[3],
                    const double lquat[4], const orc_surface *s);
int orc_geom_create_heightfield(orc_world *w, const float *heights, int nx, int ny, double dx, double x0,
                                double y0, const orc_surface *s);
void orc_world_set_max_contacts(orc_world *w, int n);
int orc_collide(orc_world *w); /* replaces the contact group; returns the number of contacts */
int orc_collide_get_pairs(const orc_world *w, int *pairs, int max_pairs); /* g1 | g2<<16, sorted */
int orc_contact_get(const orc_world *w, int c, int *b1, int *b2, orc_contact_desc *out);

/* ---- orc_post.c: what WorldPhysics::stepTheWorld does after the ODE call (src/WorldPhysics.cpp:372-397), on body
 * and joint ids, one world at a time (the device does it for the whole batch) ---- */
void orc_post_joint_update(const orc_world *w, int j, double out[7]); /* motor_torque axis1_torque3 joint_load3 */
void orc_post_joint_friction(orc_world *w, int j, double coeff);
void orc_post_frame_update(orc_world *w, int b, double dt, double lin_damp, double ang_damp, double last6[6],
                           double acc6[6]);
void orc_post_body_contact_force(const orc_world *w, int b, double out[3]);
void orc_post_step_world(orc_world *w, double dt, const double *lin_damp, const double *ang_damp,
                         const double *friction, double *last, double *acc, double *cforce, double *jupd);

/* timing helper for bench.py's cpu_baseline: collide (optional) + quickstep, `steps` times over n worlds */
void orc_bench_run(orc_world **worlds, int n, int steps, double h, int do_collide);

/* small math helpers exposed for unit tests */
void orc_plane_space(const double n[3], double p[3], double q[3]);
void orc_q_to_R(const double q[4], double R[9]);

#ifdef __cplusplus
}
#endif
#endif

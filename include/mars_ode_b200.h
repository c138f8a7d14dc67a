/*
 * mars_ode_b200.h -- C ABI of the B200-native batched rigid-body step.
 *
 * Drop-in boundary for the hot path behind WorldPhysics::stepTheWorld
 * (reference src/WorldPhysics.cpp:349-421).  The reference reaches that path through the ODE
 * C API; every entry point below names the ODE call (and the reference call site) it replaces.
 * One *arena* holds a scene template (bodies, joints, geoms) replicated over `num_worlds`
 * independent worlds in structure-of-arrays device buffers; `world` arguments select one
 * replica, B2P_ALL_WORLDS broadcasts.
 *
 * All functions return 0 (or a non-negative id) on success and a negative B2P_E* code on
 * failure; b2p_last_error() returns a description.  No exceptions cross this boundary, no
 * torch / C++ types appear in it.  There is no CPU fallback: without a CUDA device every
 * arena call fails with B2P_ENODEVICE.
 */
#ifndef MARS_ODE_B200_H
#define MARS_ODE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2P_ALL_WORLDS (-1)
#define B2P_STATIC_BODY (-1)

/* error codes */
#define B2P_OK 0
#define B2P_EINVAL (-1)
#define B2P_ENODEVICE (-2)
#define B2P_ECUDA (-3)
#define B2P_ECAPACITY (-4)
#define B2P_EUNSUPPORTED (-5)
#define B2P_EREJECTED (-6) /* contact between joint-connected bodies (WorldPhysics.cpp:461) */

/* joint types: numeric values follow the ORDER of mars::interfaces::JointType as used by the
 * switch statements in reference src/joints/Joint.cpp */
enum {
    B2P_JOINT_HINGE = 1,  /* dJointCreateHinge   -- src/joints/HingeJoint.cpp:33 */
    B2P_JOINT_HINGE2 = 2, /* dJointCreateHinge2  -- src/joints/Hinge2Joint.cpp:37 */
    B2P_JOINT_SLIDER = 3, /* dJointCreateSlider  -- src/joints/SliderJoint.cpp:34 */
    B2P_JOINT_BALL = 4,      /* dJointCreateBall      -- JOINT_TYPE_BALL arms of src/joints/Joint.cpp:299-301,447-449 */
    B2P_JOINT_UNIVERSAL = 5, /* dJointCreateUniversal -- what src/joints/UniversalJoint.cpp:50-75 configures */
    B2P_JOINT_FIXED = 6   /* dJointCreateFixed   -- src/joints/FixedJoint.cpp:35 */
};

/* dParam* names used at src/joints/Joint.cpp:315-374,618-703,1176-1464 */
enum {
    B2P_PARAM_LO_STOP = 0,
    B2P_PARAM_HI_STOP,
    B2P_PARAM_VEL,
    B2P_PARAM_LO_VEL,
    B2P_PARAM_HI_VEL,
    B2P_PARAM_FMAX,
    B2P_PARAM_FUDGE_FACTOR,
    B2P_PARAM_BOUNCE,
    B2P_PARAM_CFM,
    B2P_PARAM_STOP_ERP,
    B2P_PARAM_STOP_CFM,
    B2P_PARAM_SUSPENSION_ERP,
    B2P_PARAM_SUSPENSION_CFM,
    B2P_PARAM_ERP,
    B2P_PARAM_GROUP2 = 0x100 /* dParam*2 */
};

/* dContact surface mode bits honoured by the device path (the ones Frame::addContact can set,
 * src/objects/Frame.cpp:1038-1079) */
enum {
    B2P_CONTACT_MU2 = 0x001,
    B2P_CONTACT_SOFT_ERP = 0x008,
    B2P_CONTACT_SOFT_CFM = 0x010,
    B2P_CONTACT_ROLLING = 0x400,
    B2P_CONTACT_APPROX1_1 = 0x1000,
    B2P_CONTACT_APPROX1_2 = 0x2000,
    B2P_CONTACT_APPROX1_N = 0x4000,
    B2P_CONTACT_APPROX1 = 0x7000
};

/* SOR row ordering inside the sweep */
enum {
    B2P_REORDER_RANDOMLY = 0, /* ODE 0.16 default: LCG shuffles at iterations 8, 16, ... */
    B2P_REORDER_NONE = 1
};

/* Order of the constraint rows inside a world's solve.  ODE (util.cpp dxProcessIslands) solves one island of
 * joint-connected bodies at a time, visiting bodies, joints and the step's contacts depth-first from the most
 * recently created body; the default here is one block per world (joints in creation order, then the contact
 * stream).  Both solve the same LCP; after the fixed number of sweeps they differ by the unconverged part. */
enum {
    B2P_ORDER_WORLD_BLOCK = 0,
    B2P_ORDER_ODE_ISLANDS = 1
};

/* geom classes for the device collision path (dCreate* -- absent from the reference tree,
 * SURVEY.md section 8 a13) */
enum {
    B2P_GEOM_SPHERE = 0,
    B2P_GEOM_BOX = 1,
    B2P_GEOM_CAPSULE = 2,
    B2P_GEOM_CYLINDER = 3,
    B2P_GEOM_PLANE = 4,
    B2P_GEOM_HEIGHTFIELD = 5
};

/* The dContact subset filled by Frame::addContact (src/objects/Frame.cpp:1023-1079).
 * This 16-word record is also what the device narrowphase emits. */
typedef struct b2p_contact_desc {
    double pos[3];
    double normal[3];
    double depth;
    int mode;
    double mu, mu2;
    double rho, rho2, rhoN;
    double soft_erp, soft_cfm;
} b2p_contact_desc;

/* surface parameters attached to a geom; the collision kernels copy them into the contact
 * stream (reference: ContactData::c_params, src/objects/Frame.cpp:1047-1071) */
typedef struct b2p_surface {
    double mu, mu2;
    double rho, rho2, rhoN;
    double soft_erp, soft_cfm;
} b2p_surface;

typedef struct b2p_arena b2p_arena;

const char *b2p_last_error(void);
int b2p_version(void);
int b2p_device_count(void);

/* ---- arena / world (dWorldCreate, dWorldDestroy: src/WorldPhysics.cpp:171,228) ---- */
int b2p_arena_create(int num_worlds, int device, b2p_arena **out);
void b2p_arena_destroy(b2p_arena *a);
int b2p_arena_num_worlds(const b2p_arena *a);
/* use an existing CUDA stream (cudaStream_t) for all copies and launches; NULL = default */
int b2p_arena_set_stream(b2p_arena *a, void *cuda_stream);
/* capacity knobs; must be called before the first step */
int b2p_arena_set_max_contacts(b2p_arena *a, int max_contacts_per_world);
int b2p_arena_enable_rolling_friction(b2p_arena *a, int on);
/* tuning of the SOR kernels' on-chip row storage (0 / < 0 = automatic).  resident rows: paired-lane kernel = sweep
 * positions per world kept in shared memory (automatic: the longest world seen so far); Tensor-Memory kernel =
 * units of its shared-memory overflow pool (one unit = one sweep position of one warp's 32 worlds, 2 KB).
 * register rows: sweep positions per world kept in registers (multiple of 4; up to 20 paired-lane, up to 8 TMEM) */
int b2p_arena_set_resident_rows(b2p_arena *a, int n);
int b2p_arena_set_register_rows(b2p_arena *a, int n);
/* which SOR-LCP kernel a step uses: B2P_SOR_AUTO picks the Tensor-Memory kernel (one thread per world, rows in
 * registers + TMEM) when the scene fits it and the paired-lane kernel (rows in registers + shared memory)
 * otherwise; both produce bitwise identical results.  b2p_arena_sor_kernel_used reports the last step's choice. */
#define B2P_SOR_AUTO (-1)
#define B2P_SOR_PAIRED 0
#define B2P_SOR_TMEM 1
int b2p_arena_set_sor_kernel(b2p_arena *a, int which);
int b2p_arena_sor_kernel_used(const b2p_arena *a);
/* Tensor-Memory kernel: warps (x 32 worlds) per CTA, 4 or 8; 0 = automatic */
int b2p_arena_set_sor_warps(b2p_arena *a, int warps);

/* Which ODE stepper b2p_step runs.  B2P_STEPPER_QUICK: dWorldQuickStep (SOR-LCP, the reference's fast_step == true,
 * src/WorldPhysics.cpp:365).  B2P_STEPPER_DENSE: dWorldStep (fast_step == false, src/WorldPhysics.cpp:369, the
 * WorldPhysics constructor's default :82): the same rows, the box-LCP solved exactly by Dantzig's method (ODE
 * lcp.cpp dSolveLCP) in FP64, one world per thread; O(rows^3) per world and step. */
#define B2P_STEPPER_QUICK 0
#define B2P_STEPPER_DENSE 1
int b2p_arena_set_stepper(b2p_arena *a, int stepper);
int b2p_arena_stepper(const b2p_arena *a);
/* dense stepper: worlds (accumulated over all steps) in which the solver gave up -- ODE prints "LCP internal error"
 * and carries on with the multipliers found so far, so does this */
int b2p_dense_failures(b2p_arena *a, long long *count);
/* dense stepper, debug: complementarity residual of the last step's solved problem per world (W floats) */
int b2p_dense_residual(b2p_arena *a, float *host_out);

/* replay collide + step from a CUDA graph when their parameters did not change since the previous call
 * (default on; one host call per tick instead of four launches) */
int b2p_arena_set_graph_launch(b2p_arena *a, int on);
/* experiment switches of the kernels (bit mask; default: the environment variable B2P_TUNE, else 0 = production):
 * 1 SOR CTAs take the worlds in ascending instead of descending order; 2 SOR streaming sweeps gather all rows
 * first and sweep afterwards; 4 the integrate kernel keeps nothing in shared memory (the path scenes too large for
 * the stash take).  Results do not depend on them. */
int b2p_arena_set_tune(b2p_arena *a, int bits);
/* What WorldPhysics::stepTheWorld does after the ODE call (src/WorldPhysics.cpp:372-397) is computed on the device
 * for every world.  Flags select the telemetry arrays; damping (b2p_body_set_damping) and joint friction
 * (b2p_joint_set_friction_coefficient) act on the state and are always applied.  Default: both. */
#define B2P_POST_JOINT_UPDATE 1 /* Joint::update: motor torque, axis torque, joint load (Joint.cpp:859-1019) */
#define B2P_POST_FRAME_UPDATE 2 /* Frame::update accelerations + computeContactForce (Frame.cpp:899-921,159-171) */
int b2p_arena_set_post_step_outputs(b2p_arena *a, int flags);

int b2p_world_set_gravity(b2p_arena *a, double x, double y, double z);      /* dWorldSetGravity :194 */
int b2p_world_set_cfm(b2p_arena *a, double cfm);                            /* dWorldSetCFM :195 */
int b2p_world_set_erp(b2p_arena *a, double erp);                            /* dWorldSetERP :196 */
int b2p_world_set_max_angular_speed(b2p_arena *a, double s);                /* dWorldSetMaxAngularSpeed :197 */
int b2p_world_set_contact_max_correcting_vel(b2p_arena *a, double v);       /* dWorldSetContactMaxCorrectingVel :198 */
int b2p_world_set_contact_surface_layer(b2p_arena *a, double d);
int b2p_world_set_quickstep_num_iterations(b2p_arena *a, int n);
int b2p_world_set_quickstep_w(b2p_arena *a, double w);
int b2p_world_set_reorder_method(b2p_arena *a, int method);
int b2p_world_set_row_order(b2p_arena *a, int order);
int b2p_world_rand_set_seed(b2p_arena *a, uint32_t seed);                   /* dRandSetSeed :169 (per world) */
int b2p_world_rand_get_seed(b2p_arena *a, int world, uint32_t *seed);

/* ---- bodies (dBody*: src/objects/Frame.cpp passim) ---- */
int b2p_body_create(b2p_arena *a);                                          /* dBodyCreate :190,609 */
int b2p_body_set_mass(b2p_arena *a, int body, double mass, const double I[9]); /* dBodySetMass :850 */
int b2p_body_set_position(b2p_arena *a, int body, int world, double x, double y, double z);
int b2p_body_set_quaternion(b2p_arena *a, int body, int world, double w, double x, double y, double z);
int b2p_body_set_linear_vel(b2p_arena *a, int body, int world, double x, double y, double z);
int b2p_body_set_angular_vel(b2p_arena *a, int body, int world, double x, double y, double z);
int b2p_body_set_force(b2p_arena *a, int body, int world, double x, double y, double z);
int b2p_body_set_torque(b2p_arena *a, int body, int world, double x, double y, double z);
int b2p_body_add_force(b2p_arena *a, int body, int world, double x, double y, double z);
int b2p_body_add_torque(b2p_arena *a, int body, int world, double x, double y, double z);
int b2p_body_add_force_at_pos(b2p_arena *a, int body, int world, double fx, double fy, double fz,
                              double px, double py, double pz);
int b2p_body_set_gravity_mode(b2p_arena *a, int body, int on);
int b2p_body_set_gyroscopic_mode(b2p_arena *a, int body, int on);
int b2p_body_get_position(b2p_arena *a, int body, int world, double out[3]);
int b2p_body_get_quaternion(b2p_arena *a, int body, int world, double out[4]);
int b2p_body_get_rotation(b2p_arena *a, int body, int world, double out[9]);
int b2p_body_get_linear_vel(b2p_arena *a, int body, int world, double out[3]);
int b2p_body_get_angular_vel(b2p_arena *a, int body, int world, double out[3]);
int b2p_body_get_force(b2p_arena *a, int body, int world, double out[3]);
int b2p_body_get_torque(b2p_arena *a, int body, int world, double out[3]);
/* Frame's linearDamping / angularDamping (src/objects/Frame.cpp:55-58): after every step the body's velocity is
 * multiplied by the factor if it is < 1 (Frame::update :922-941).  Per template, all worlds. */
int b2p_body_set_damping(b2p_arena *a, int body, double linear, double angular);
int b2p_body_get_damping(b2p_arena *a, int body, double *linear, double *angular);
/* Frame::update / computeContactForce results of the last step (Frame.cpp:899-921,159-171) */
int b2p_body_get_acceleration(b2p_arena *a, int body, int world, double lin[3], double ang[3]);
int b2p_body_get_contact_force(b2p_arena *a, int body, int world, double out[3]);
/* all bodies and worlds: float[num_bodies][15][num_worlds] = last_l_vel3 last_a_vel3 l_acc3 a_acc3 contact_force3 */
int b2p_body_post_download(b2p_arena *a, float *host);
void *b2p_device_body_post(b2p_arena *a);   /* [nb*15][Wp] */
/* 13 values: pos3, quat4 (w,x,y,z), lvel3, avel3 */
int b2p_body_get_state(b2p_arena *a, int body, int world, double out[13]);
int b2p_body_set_state(b2p_arena *a, int body, int world, const double in[13]);

/* batched state I/O in the device layout: float[num_bodies][13][num_worlds] */
int b2p_state_upload(b2p_arena *a, const float *host_state);
int b2p_state_download(b2p_arena *a, float *host_state);
/* one body of every world: float[13][num_worlds] (asynchronous on the arena stream; call
 * b2p_sync before reading the host buffer) */
int b2p_body_get_state_batch(b2p_arena *a, int body, float *host_out);
/* joint observations of the last step for all worlds: float[num_joints][4][num_worlds] = angle1 rate1 angle2 rate2
 * (dJointGet*Angle / dJointGet*AngleRate / dJointGetSliderPosition(Rate), src/joints/Joint.cpp:395-415,1072-1102) */
int b2p_joint_obs_download(b2p_arena *a, float *host_obs);
/* pipelined form for control loops: snapshot on the step's stream, D2H on a separate copy stream so that it
 * overlaps the next step; returns a ticket (0 or 1, two transfers may be in flight), b2p_io_wait(ticket)
 * blocks until host_out is complete */
int b2p_body_get_state_batch_begin(b2p_arena *a, int body, float *host_out);
/* the same with the joint observations of all joints in the same transfer: host_joint_obs = float[num_joints][4]
 * [num_worlds] (angle1 rate1 angle2 rate2) or NULL */
int b2p_observe_begin(b2p_arena *a, int body, float *host_out, float *host_joint_obs);
int b2p_io_wait(b2p_arena *a, int ticket);
/* page-locked host memory for the batched transfers above (asynchronous copies need it to overlap the step) */
void *b2p_host_alloc(size_t bytes);
void b2p_host_free(void *p);
/* device pointers (float, layout [field][padded_worlds]) for zero-copy consumers */
int b2p_padded_worlds(const b2p_arena *a);
void *b2p_device_state(b2p_arena *a);       /* [nb*13][Wp] */
void *b2p_device_joint_cmd(b2p_arena *a);   /* [nj*4][Wp]: vel1,fmax1,vel2,fmax2 */
void *b2p_device_joint_obs(b2p_arena *a);   /* [nj*4][Wp]: angle1,rate1,angle2,rate2 */
void *b2p_device_joint_feedback(b2p_arena *a);  /* [nj*13][Wp]: f1 t1 f2 t2 lambda (dJointFeedback, Joint.cpp:268) */
void *b2p_device_contact_records(b2p_arena *a); /* [maxc*16][Wp]: the contact stream (b2p_contact_desc order) */
void *b2p_device_contact_count(b2p_arena *a);   /* int[Wp] */
/* dJointFeedback of all joints and worlds: float[num_joints][13][num_worlds] (Joint::update, Joint.cpp:859-1019) */
int b2p_joint_feedback_download(b2p_arena *a, float *host_fb);

/* ---- joints (dJoint*: reference src/joints/) ---- */
int b2p_joint_create(b2p_arena *a, int type, int body1, int body2);         /* dJointCreate* + dJointAttach */
int b2p_joint_set_anchor(b2p_arena *a, int joint, double x, double y, double z);
int b2p_joint_set_axis1(b2p_arena *a, int joint, double x, double y, double z);
int b2p_joint_set_axis2(b2p_arena *a, int joint, double x, double y, double z);
int b2p_joint_set_fixed(b2p_arena *a, int joint);                           /* dJointSetFixed */
/* VEL and FMAX (both axes) may be set per world; all other parameters are per template */
int b2p_joint_set_param(b2p_arena *a, int joint, int world, int param, double value);
int b2p_joint_get_param(b2p_arena *a, int joint, int world, int param, double *value);
/* batched motor commands: float[num_worlds] per call */
int b2p_joint_set_param_batch(b2p_arena *a, int joint, int param, const float *values);
/* the same for joints 0 .. num_joints-1 in one transfer: values[num_joints][num_worlds] */
int b2p_joints_set_param_batch(b2p_arena *a, int param, const float *values, int num_joints);
int b2p_joint_get_anchor(b2p_arena *a, int joint, int world, double out[3]);
int b2p_joint_get_anchor2(b2p_arena *a, int joint, int world, double out[3]);
int b2p_joint_get_axis1(b2p_arena *a, int joint, int world, double out[3]);
int b2p_joint_get_axis2(b2p_arena *a, int joint, int world, double out[3]);
int b2p_joint_get_angle1(b2p_arena *a, int joint, int world, double *out);       /* dJointGetHingeAngle / SliderPosition / Hinge2Angle1 */
int b2p_joint_get_angle1_rate(b2p_arena *a, int joint, int world, double *out);
int b2p_joint_get_angle2(b2p_arena *a, int joint, int world, double *out);
int b2p_joint_get_angle2_rate(b2p_arena *a, int joint, int world, double *out);
int b2p_joint_add_torque(b2p_arena *a, int joint, int world, double t);          /* dJointAddHingeTorque / dJointAddSliderForce */
/* dJointFeedback incl. the patched `lambda`: f1 t1 f2 t2 lambda = 13 doubles (Joint.cpp:268,900) */
int b2p_joint_get_feedback(b2p_arena *a, int joint, int world, double out[13]);
int b2p_joint_get_num_rows(b2p_arena *a, int joint, int world, int *rows);
/* Joint's frictionCoefficient (src/joints/Joint.cpp:43,47-52): when > 0, Joint::applyJointFriction (:1031-1063) adds
 * its force / torque to both bodies after every step, for the next one.  Per template, all worlds. */
int b2p_joint_set_friction_coefficient(b2p_arena *a, int joint, double k);
int b2p_joint_get_friction_coefficient(b2p_arena *a, int joint, double *k);
/* Joint::update results of the last step (Joint.cpp:859-1019): motor_torque, axis1_torque[3], joint_load[3] */
int b2p_joint_get_update(b2p_arena *a, int joint, int world, double out[7]);
int b2p_joint_update_download(b2p_arena *a, float *host); /* float[num_joints][7][num_worlds] */
void *b2p_device_joint_update(b2p_arena *a);              /* [nj*7][Wp] */
int b2p_are_connected_excluding_contacts(b2p_arena *a, int body1, int body2);    /* dAreConnectedExcluding :461 */

/* ---- host-injected contacts (dJointCreateContact: src/WorldPhysics.cpp:458-468) ---- */
int b2p_contacts_clear(b2p_arena *a, int world);                                 /* dJointGroupEmpty :317 */
int b2p_contact_add(b2p_arena *a, int world, int body1, int body2, const b2p_contact_desc *c);
int b2p_contact_count(b2p_arena *a, int world, int *count);
int b2p_contact_get_feedback(b2p_arena *a, int world, int contact, double f1[3]);
/* read back a contact record of the current stream (host-injected or device-generated) */
int b2p_contact_get(b2p_arena *a, int world, int contact, int *body1, int *body2, b2p_contact_desc *c);

/* ---- device collision path (dSpaceCollide + dCollide) ---- */
int b2p_geom_create(b2p_arena *a, int geom_class, int body, const double dims[4],
                    const double local_pos[3], const double local_quat[4], const b2p_surface *s);
/* heights: row-major [ny][nx] floats, sample spacing dx, origin (x0,y0), shared by all worlds */
int b2p_geom_create_heightfield(b2p_arena *a, const float *heights, int nx, int ny, double dx,
                                double x0, double y0, const b2p_surface *s);
/* run broadphase + narrowphase on the device, replacing the contact stream of every world */
int b2p_collide(b2p_arena *a);
/* b2p_collide + b2p_step in one call (one CUDA graph launch once the parameters are stable) */
int b2p_collide_step(b2p_arena *a, double step_size);
/* contacts the narrowphase found beyond max_contacts_per_world (dropped in pair order), summed over all worlds and
 * all b2p_collide calls so far: 0 means no stream was ever truncated */
int b2p_contacts_dropped(b2p_arena *a, long long *count);
/* debug output: record the broadphase pair list during b2p_collide (off by default) */
int b2p_collide_record_pairs(b2p_arena *a, int on);
/* broadphase pair list of one world, sorted by (min geom, max geom); returns pair count */
int b2p_collide_get_pairs(b2p_arena *a, int world, int *pairs, int max_pairs);

/* ---- the step (dWorldQuickStep: src/WorldPhysics.cpp:365; dWorldStep, :369, after b2p_arena_set_stepper) ---- */
int b2p_step(b2p_arena *a, double step_size);
int b2p_sync(b2p_arena *a);
/* diagnostics of the last step */
int b2p_world_last_num_rows(b2p_arena *a, int world, int *rows);
int b2p_kernel_launch_count(const b2p_arena *a); /* kernels launched by this arena so far */
/* diagnostics: per-kernel CUDA-event timing of b2p_step (enable resets; every step then synchronises);
 * out[0..2] = mean milliseconds of the assemble / SOR / integrate kernels, out[3] = steps accumulated. */
int b2p_debug_kernel_times(b2p_arena *a, int enable, double out[4]);
/* LCP residual of the last step's solve, one value per world: max over rows of the projected
 * |J invM J^T lambda + cfm lambda - rhs| (the oracle's orc_world_last_residual) */
int b2p_debug_lcp_residual(b2p_arena *a, float *host_out);
/* algorithmic bytes one step of one world moves */
double b2p_algorithmic_bytes_per_world_step(const b2p_arena *a);

#ifdef __cplusplus
}
#endif
#endif

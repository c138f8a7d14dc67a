// b2p_dev.h -- structures shared by the host arena and the sm_100a kernels.
//
// Device memory layout (all per-world arrays are structure-of-arrays with the world index
// fastest, padded to a multiple of 32 worlds = Wp, so that a warp of 32 consecutive worlds
// touches one 128-byte line per field):
//   state   float[nb*13][Wp]   pos3 quat4(w,x,y,z) lvel3 avel3   (dBody state)
//   force   float[nb*6][Wp]    facc3 tacc3                        (dBodyAddForce accumulators)
//   jcmd    float[nj*4][Wp]    vel1 fmax1 vel2 fmax2              (dParamVel / dParamFMax)
//   jobs    float[nj*4][Wp]    angle1 rate1 angle2 rate2          (dJointGet*Angle / *Rate)
//   jfb     float[nj*13][Wp]   f1 t1 f2 t2 lambda                 (dJointFeedback)
//   jrows   uint8[nj][Wp]      rows used by each joint in the last step | limit/motor row 1 present<<3 | row 2<<4
//   ccount  int[Wp]            contacts in the stream of each world
//   crec    float[maxc*16][Wp] the 16-word contact record         (dContact subset)
//   cfb     float[maxc*3][Wp]  contact feedback f1
//   seed    uint32[Wp]         per-world LCG state                (dRandSetSeed)
//   nrows   int[Wp]            rows of the last step | row blocks (islands) << 16
//   isl     uint32[nb][Wp]     block k: first slot | independent ("head") rows << 8 | rows << 16
//   ord0    uint8[maxrows][Wp] slot of the row at sweep position p of ODE's initial order (iteration 0)
//   rows    float[maxrows][2][Wp][8]  constraint rows by slot as two 32-byte units per world: J0..J7 | J8..J11 rhs'
//                              cfm' bound meta.  Slots are static: joint j owns slots j*B2P_ROWS_PER_JOINT .., contact
//                              c owns njslots + c*rpc .. (unused slots are never read); ord0 lists the used slots in
//                              ODE's initial sweep order
//                              (12 mass- and Ad-scaled Jacobian words); one more slot (index maxrows) holds an
//                              all-zero row for every world.  ROW_UNIT(rows, slot, u, Wp, w) addresses a unit
//   lambda  float[maxrows][Wp] scaled multipliers by slot;  rowS float[maxrows][Wp]  sqrt(Ad) by slot
//   fcacc   float[nb*6][Wp]    M^-1/2 J^T lambda after the last sweep;  invIw float[nb*9][Wp]
//   cmap    uint32[maxc][Wp]   rows | depmask<<3 | slot of the normal row<<8 | first dependent slot<<16 | fr1<<24 | fr2<<25
//   bodyx   float[nb*15][Wp]   last_l_vel3 last_a_vel3 l_acc3 a_acc3 contact_force3   (Frame::update)
//   jupd    float[nj*7][Wp]    motor_torque axis1_torque3 joint_load3                 (Joint::update)
#pragma once
#include <stdint.h>

#define B2P_MAX_DEVICES 16
#define B2P_MAX_BODIES 32
#define B2P_MAX_JOINTS 48
#define B2P_MAX_GEOMS 48
#define B2P_ROWS_PER_JOINT 6
#define B2P_STATE_FIELDS 13
#define B2P_CONTACT_FIELDS 16

// contact record field indices
enum {
    CR_IDS = 0, // int: body0 | body1<<8 (0xFF = static)
    CR_PX, CR_PY, CR_PZ,
    CR_NX, CR_NY, CR_NZ,
    CR_DEPTH,
    CR_MU, CR_MU2,
    CR_SOFT_ERP, CR_SOFT_CFM,
    CR_RHO, CR_RHO2, CR_RHON,
    CR_MODE // int
};

struct DevLimot {
    float lostop, histop, fudge_factor, normal_cfm, stop_erp, stop_cfm, bounce;
    float pad;
};

struct DevBody {
    float mass, invMass;
    float I[9];
    float invI[9];
    int gyroscopic, nogravity, has_max_ang;
    float max_angular_speed;
    // symmetric square roots (body frame) used by the mass-scaled row form: J~ = J M^-1/2
    float sqrtInvMass, sqrtMass;
    float sqrtInvI[9], sqrtI[9];
    // Frame::update (src/objects/Frame.cpp:899-943): multiplicative velocity damping applied after the step when < 1
    float lin_damp, ang_damp;
};

struct DevJoint {
    int type, b0, b1, reverse;
    int slot_base; // first solver row slot
    int pad0, pad1, pad2;
    float anchor1[3], anchor2[3], axis1[3], axis2[3], qrel[4], offset[3];
    DevLimot limot1, limot2;
    float c0, s0, v1[3], v2[3], w1[3], w2[3], susp_erp, susp_cfm;
    float erp, cfm; // fixed / ball joint
    float qrel2[4]; // universal: body 2 against the cross (qrel = body 1 against the cross)
    float friction; // Joint::applyJointFriction coefficient (src/joints/Joint.cpp:1031-1063); <= 0: off
    float pad3[3];
};

struct DevGeom {
    int cls, body; // body = -1: static
    float dims[4]; // sphere r | box lx ly lz | capsule r len | cylinder r len | plane a b c d
    float lpos[3], lquat[4];
    int has_offset;
    float mu, mu2, rho, rho2, rhoN, soft_erp, soft_cfm;
};

struct DevHeightfield {
    const float *heights;
    int nx, ny;
    float dx, x0, y0;
    float hmin, hmax;
};

struct DevTemplate {
    int nb, nj, ng, maxc;
    int rpc;     // rows per contact (3, or 6 with rolling friction)
    int njslots; // nj * B2P_ROWS_PER_JOINT
    int maxrows; // njslots + maxc*rpc
    int has_limit_motors;
    DevBody bodies[B2P_MAX_BODIES];
    DevJoint joints[B2P_MAX_JOINTS];
    DevGeom geoms[B2P_MAX_GEOMS];
    uint32_t connected[B2P_MAX_BODIES]; // bit b set: joined to body b by a non-contact joint
    DevHeightfield hf;
};

struct StepParams {
    const DevTemplate *T;
    int W, Wp;
    float h;
    float gravity[3];
    float erp, cfm, contact_max_vel, contact_min_depth, sor_w;
    int iters, reorder;
    float *state, *force, *jcmd, *jobs, *jfb, *crec, *cfb, *lambda, *rowS, *invIw, *fcacc;
    float4 *rows;   // [slot][2][Wp] units of two float4; slot = position in ODE's initial order (+ the zero row)
    uint32_t *cmap; // [maxc][Wp]
    uint8_t *jrows; // [nj][Wp]: rows of joint j | limit/motor row of axis 1 present << 3 | of axis 2 << 4
    int *ccount;
    int *nrows;
    uint32_t *seed;
    int rows_registers; // sweep positions per world the SOR kernel keeps in registers (0, 4, ... 20)
    int rows_resident;  // further positions it keeps in shared memory (the rest is read from L2 in place)
    int *hw_rows;       // device counter: largest row count of any world so far
    int sor_warps;      // Tensor-Memory SOR kernel: warps per CTA (4 or 8); rows_resident = its overflow pool units
    int worlds_per_cta; // Tensor-Memory SOR kernel: worlds a CTA takes (multiple of 32, <= 32 * sor_warps)
    // post-step outputs of the reference's stepTheWorld for every world (null: not wanted)
    // ODE solves island by island: block k of a world's rows = start | independent rows << 8 | rows << 16
    uint32_t *isl;  // [max(nb,1)][Wp]; nrows = rows of the world | blocks << 16
    uint8_t *ord0;  // [maxrows][Wp]: slot of the row at position p of the initial sweep order (p < rows of the world)
    int order_mode; // 0: one block per world (joints in creation order, then contacts); 1: ODE's island order
    float *bodyx; // [nb*15][Wp] last_l_vel3 last_a_vel3 l_acc3 a_acc3 contact_force3  (Frame::update, computeContactForce)
    float *jupd;  // [nj*7][Wp]  motor_torque axis1_torque3 joint_load3              (Joint::update)
    int tune;     // experiment switches from the environment variable B2P_TUNE (bit mask; 0 in production)
    int stepper;  // 0: dWorldQuickStep (SOR-LCP kernels); 1: dWorldStep (dense LCP, b2p_dense.cu)
};

// per-world scratch of the dense stepper (b2p_dense.cu), every array world index fastest
struct DenseScratch {
    double *A, *L;  // [maxrows * maxrows][Wp]: the LCP matrix; the Cholesky factor of its clamped block
    double *vec;    // [9 * maxrows + 6 * (nb + 1)][Wp]: x w lo hi dx dw b + two solve vectors; fch in double
    uint8_t *sets;  // [5 * maxrows][Wp]: clamped list, set of every row, scratch list, normal-row position, order
    int *failures;  // worlds in which the solver gave up (ODE: "LCP internal error"), accumulated
    float *residual; // [Wp] complementarity residual of the solved problem, or null
};

// unit u (0, 1) of the row in `slot` of world w: a pointer to two consecutive float4
#define ROW_UNIT(rows, slot, u, Wp, w) ((rows) + (((size_t)((slot) * 2 + (u)) * (size_t)(Wp) + (size_t)(w)) * 2))

struct CollideParams {
    const DevTemplate *T;
    int W, Wp;
    float *state, *crec;
    int *ccount;
    int *pairs; // optional debug: int[maxpairs][Wp] packed g1 | g2<<16, or null
    int *npairs;
    int maxpairs;
    int *dropped; // one counter per arena: contacts found beyond maxc (they are dropped in pair order), or null
};

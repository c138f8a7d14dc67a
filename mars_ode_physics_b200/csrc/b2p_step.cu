// b2p_step.cu -- the fused per-world step kernel for sm_100a.
//
// One thread owns one world (worlds are independent: reference src/WorldPhysics.hpp:99 gives each
// WorldPhysics its own dWorldID), one warp owns 32 consecutive worlds, and every global access is
// a world-fastest SoA access, i.e. one fully used 128-byte line per field per warp.
//
// What it computes is what dWorldQuickStep computes for the reference at
// src/WorldPhysics.cpp:365 (ODE 0.16 quickstep: body preparation with implicit gyroscopic
// torque, joint/contact constraint rows, SOR-LCP with ODE's row order and LCG shuffles,
// velocity update, dxStepBody), followed by the per-joint feedback the reference reads in
// Joint::update (src/joints/Joint.cpp:859-1019).
//
// Phases (all inside one launch so the row scratch written in phase 2 is still L2-resident when
// phase 3 streams it `iters` times):
//   1 bodies:   R(q), invI_world, gyroscopic torque, gravity; tmp1 = v/h + invM*fe -> smem
//   2 rows:     per joint / per contact: J, c, cfm, lo, hi -> rhs, iMJ, Ad; scaled row -> gmem
//   3 SOR:      order[] in smem (per world, so per-world contact counts and RNG states are
//               allowed to differ); fc[] in smem, lane-interleaved => bank-conflict free for any
//               per-lane body index
//   4 epilogue: v += h*fc + h*invM*fe, clamp |w|, integrate pose, feedback, joint angles
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "b2p_dev.h"
#include "b2p_math.cuh"

namespace {

constexpr int BD = 64;   // threads (= worlds) per CTA: 8 CTAs/SM at <=128 registers, fine-grained waves
// The cp.async ring depth (STAGES-1 rows in flight per thread) is a template parameter: the launcher
// picks the deepest ring whose shared memory still lets all worlds of the launch be resident at once.
constexpr int ROWV = 4; // float4 per stored row: one 64-byte line per (world,row)

struct Smem {
    float *fc;     // [nb*6][BD]   fe -> tmp1 -> fc
    float *iw;     // [nb*6][BD]   symmetric invI_world (xx xy xz yy yz zz) for the sweep
    float *im;     // [nb]         invMass (shared by the CTA)
    float *lamn;   // [maxc][BD]   lambda of each contact's normal row (read by its friction rows)
    float4 *ring;  // [STAGES][4][BD] cp.async ring of rows in sweep order
    float *lring;  // [STAGES][BD]    ... and of their lambdas
    uint8_t *jm;   // [nj][BD]     rows of joint j in this world
    int8_t *jl;    // [nj][BD]     local limot row feeding feedback.lambda, or -1
    uint8_t *cm;   // [maxc][BD]   rows of contact c | dependency mask << 3
};

#define G(arr, field) (arr)[(size_t)(field) * Wp + w]

// Ampere-style asynchronous copies (LDGSTS): completion is tracked per thread by async groups.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t smem_addr, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct BodyPose {
    float pos[3], q[4], R[9];
};

__device__ __forceinline__ void load_pose(const StepParams &P, int b, int w, BodyPose &bp)
{
    const int Wp = P.Wp;
    const int f = b * B2P_STATE_FIELDS;
    bp.pos[0] = G(P.state, f + 0);
    bp.pos[1] = G(P.state, f + 1);
    bp.pos[2] = G(P.state, f + 2);
    bp.q[0] = G(P.state, f + 3);
    bp.q[1] = G(P.state, f + 4);
    bp.q[2] = G(P.state, f + 5);
    bp.q[3] = G(P.state, f + 6);
    q_to_R(bp.q, bp.R);
}

__device__ __forceinline__ void load_invIw(const StepParams &P, int b, int w, float *I)
{
    const int Wp = P.Wp;
#pragma unroll
    for (int k = 0; k < 9; k++) I[k] = G(P.invIw, b * 9 + k);
}

// One constraint row under construction.
struct Row {
    float J[12];
    float c, cfm, lo, hi;
    int fr, nr; // contact index + 1 if this is a friction-dependent row / a contact normal row, else 0
};

__device__ __forceinline__ void row_init(Row &r, float world_cfm)
{
#pragma unroll
    for (int k = 0; k < 12; k++) r.J[k] = 0.f;
    r.c = 0.f;
    r.cfm = world_cfm;
    r.lo = -INFINITY;
    r.hi = INFINITY;
    r.fr = 0;
    r.nr = 0;
}

struct PairCtx {
    int b0, b1; // b1 < 0: static
    float im0, im1;
    float I0[9], I1[9];
};

// rhs = c/h - J.tmp1 ; iMJ = invM J^T ; Ad = w/(J iMJ + cfm/h) ; scale ; store.
// Follows ODE quickstep stage 2c / SOR_LCP setup, same operation order as oracle/orc_quickstep.c.
__device__ __forceinline__ void emit_row(const StepParams &P, const Smem &S, int w, int tid, int slot, Row &r,
                                         const PairCtx &pc, int bpos = -1)
{
    const int Wp = P.Wp;
    const float hinv = 1.0f / P.h;
    float rhs = r.c * hinv;
    const float cfm = r.cfm * hinv;
    float s = 0.f;
    {
        const float *t = S.fc + (size_t)(pc.b0 * 6) * BD + tid;
#pragma unroll
        for (int k = 0; k < 6; k++) s += r.J[k] * t[k * BD];
    }
    if (pc.b1 >= 0) {
        const float *t = S.fc + (size_t)(pc.b1 * 6) * BD + tid;
#pragma unroll
        for (int k = 0; k < 6; k++) s += r.J[6 + k] * t[k * BD];
    }
    rhs -= s;
    float o[12];
    o[0] = pc.im0 * r.J[0];
    o[1] = pc.im0 * r.J[1];
    o[2] = pc.im0 * r.J[2];
    mul_R_v(o + 3, pc.I0, r.J + 3);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) sum += o[k] * r.J[k];
    if (pc.b1 >= 0) {
        o[6] = pc.im1 * r.J[6];
        o[7] = pc.im1 * r.J[7];
        o[8] = pc.im1 * r.J[8];
        mul_R_v(o + 9, pc.I1, r.J + 9);
#pragma unroll
        for (int k = 6; k < 12; k++) sum += o[k] * r.J[k];
    } else {
#pragma unroll
        for (int k = 6; k < 12; k++) o[k] = 0.f;
    }
    const float Ad = P.sor_w / (sum + cfm);
    // Stored row (13 words): unscaled J1l J1a J2a (J2l = -J1l for every joint type of this path),
    // rhs, cfm/h, the (lo,hi) pair folded into one word, and Ad.  The sweep evaluates
    // delta = Ad*(rhs - lambda*cfm - J.fc) and rebuilds iMJ = invM*J from invI_world (L1-resident),
    // which cuts the bytes streamed per row-iteration from 128 to 60.
    float hi_enc;
    if (r.lo == 0.f && r.hi == INFINITY) hi_enc = __int_as_float(0xFFFFFFFF);       // [0, +inf)
    else if (r.lo == -INFINITY && r.hi == 0.f) hi_enc = __int_as_float(0xFFFFFFFE); // (-inf, 0]
    else hi_enc = r.hi;                                                               // [-hi, hi]
    // Scratch A is slot-major SoA ([slot][4][world], coalesced when the slot is warp-uniform, as it is
    // here and in the feedback pass).  .w of the last vector packs body0 | body1<<8 | fr<<16 | nr<<24:
    // body1 = 0xff marks a one-body row;
    // fr / nr = contact index + 1 for a friction-dependent row / a contact normal row.
    const int meta = (pc.b0 & 0xff) | ((pc.b1 < 0 ? 0xff : pc.b1) << 8) | (r.fr << 16) | (r.nr << 24);
    const float4 q0 = make_float4(r.J[0], r.J[1], r.J[2], r.J[3]);
    const float4 q1 = make_float4(r.J[4], r.J[5], r.J[9], r.J[10]);
    const float4 q2 = make_float4(r.J[11], rhs, cfm, hi_enc);
    const float4 q3 = make_float4(Ad, 0.f, __int_as_float(meta), 0.f);
    float4 *dst = P.rows + (size_t)(slot * ROWV) * Wp + w;
    dst[0 * (size_t)Wp] = q0;
    dst[1 * (size_t)Wp] = q1;
    dst[2 * (size_t)Wp] = q2;
    dst[3 * (size_t)Wp] = q3;
    if (bpos >= 0) {
        // independent rows already know their sweep position (ODE: rows with findex < 0 first, in row
        // order), so they go straight into scratch B as well
        float4 *db = P.rowsB + (size_t)(bpos * ROWV) * Wp + w;
        db[0 * (size_t)Wp] = q0;
        db[1 * (size_t)Wp] = q1;
        db[2 * (size_t)Wp] = q2;
        db[3 * (size_t)Wp] = q3;
    }
}

struct LimotState {
    int limit;
    float limit_err;
};

// dxJointLimitMotor::testRotationalLimit
__device__ __forceinline__ void limot_test(const DevLimot &l, float value, LimotState &st)
{
    if (value <= l.lostop) {
        st.limit = 1;
        st.limit_err = value - l.lostop;
    } else if (value >= l.histop) {
        st.limit = 2;
        st.limit_err = value - l.histop;
    } else {
        st.limit = 0;
        st.limit_err = 0.f;
    }
}

// getHingeAngle (ODE joint.cpp), unsigned by dJOINT_REVERSE
__device__ __forceinline__ float hinge_angle(const DevJoint &jc, const float *q0, const float *q1, bool has_b1)
{
    float qrel[4];
    if (has_b1) {
        float qq[4];
        qmul1(qq, q0, q1);
        qmul2(qrel, qq, jc.qrel);
    } else {
        qmul3(qrel, q0, jc.qrel);
    }
    const float cost2 = qrel[0];
    const float sint2 = sqrtf(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
    float theta = (qrel[1] * jc.axis1[0] + qrel[2] * jc.axis1[1] + qrel[3] * jc.axis1[2] >= 0.f)
                      ? (2.f * atan2f(sint2, cost2))
                      : (2.f * atan2f(sint2, -cost2));
    if (theta > 3.14159265358979323846f) theta -= 2.f * 3.14159265358979323846f;
    return -theta;
}

__device__ __forceinline__ float slider_position(const DevJoint &jc, const BodyPose &p0, const BodyPose &p1,
                                                 bool has_b1)
{
    float ax1[3], q[3];
    mul_R_v(ax1, p0.R, jc.axis1);
    if (has_b1) {
        mul_R_v(q, p1.R, jc.offset);
#pragma unroll
        for (int i = 0; i < 3; i++) q[i] = p0.pos[i] - q[i] - p1.pos[i];
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) q[i] = p0.pos[i] - jc.offset[i];
        if (jc.reverse) {
#pragma unroll
            for (int i = 0; i < 3; i++) ax1[i] = -ax1[i];
        }
    }
    return dot3(ax1, q);
}

__device__ __forceinline__ float hinge2_angle1(const DevJoint &jc, const BodyPose &p0, const BodyPose &p1)
{
    float p[3], q[3];
    mul_R_v(p, p1.R, jc.axis2);
    mul_Rt_v(q, p0.R, p);
    const float x = dot3(jc.v1, q);
    const float y = dot3(jc.v2, q);
    return -atan2f(y, x);
}

__device__ __forceinline__ float hinge2_angle2(const DevJoint &jc, const BodyPose &p0, const BodyPose &p1)
{
    float p[3], q[3];
    mul_R_v(p, p0.R, jc.axis1);
    mul_Rt_v(q, p1.R, p);
    const float x = dot3(jc.w1, q);
    const float y = dot3(jc.w2, q);
    return -atan2f(y, x);
}

// Limit state of limot 1 of joint jc in this world (getInfo1 part that looks at the pose).
__device__ __forceinline__ void joint_limit_state(const DevJoint &jc, const BodyPose &p0, const BodyPose &p1,
                                                  LimotState &st)
{
    st.limit = 0;
    st.limit_err = 0.f;
    const DevLimot &l = jc.limot1;
    const float PI = 3.14159265358979323846f;
    if (jc.type == 1) { // hinge
        if ((l.lostop >= -PI || l.histop <= PI) && l.lostop <= l.histop)
            limot_test(l, hinge_angle(jc, p0.q, p1.q, jc.b1 >= 0), st);
    } else if (jc.type == 3) { // slider
        if ((l.lostop > -INFINITY || l.histop < INFINITY) && l.lostop <= l.histop)
            limot_test(l, slider_position(jc, p0, p1, jc.b1 >= 0), st);
    } else if (jc.type == 2) { // hinge2
        if ((l.lostop >= -PI || l.histop <= PI) && l.lostop <= l.histop)
            limot_test(l, hinge2_angle1(jc, p0, p1), st);
    }
}

// dxJointLimitMotor::addLimot, row part.  Returns 1 if the row is used.
// The "powered while at a limit" force injection is done by limot_prepass() before tmp1 exists.
__device__ __forceinline__ int add_limot_row(const StepParams &P, const DevLimot &l, float vel, float fmax,
                                             const LimotState &st, Row &r, const float *ax1, int rotational,
                                             const BodyPose &p0, const BodyPose &p1, bool has_b1)
{
    int powered = fmax > 0.f;
    if (!(powered || st.limit)) return 0;
    float *J1 = rotational ? r.J + 3 : r.J;
    float *J2 = rotational ? r.J + 9 : r.J + 6;
#pragma unroll
    for (int i = 0; i < 3; i++) J1[i] = ax1[i];
    if (has_b1) {
#pragma unroll
        for (int i = 0; i < 3; i++) J2[i] = -ax1[i];
    }
    if (!rotational && has_b1) {
        float c[3], ltd[3];
#pragma unroll
        for (int i = 0; i < 3; i++) c[i] = 0.5f * (p1.pos[i] - p0.pos[i]);
        cross3(ltd, c, ax1);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            r.J[3 + i] = ltd[i];
            r.J[9 + i] = ltd[i];
        }
    }
    if (st.limit && (l.lostop == l.histop)) powered = 0;
    if (powered) {
        r.cfm = l.normal_cfm;
        if (!st.limit) {
            r.c = vel;
            r.lo = -fmax;
            r.hi = fmax;
        }
    }
    if (st.limit) {
        const float k = (1.0f / P.h) * l.stop_erp;
        r.c = -k * st.limit_err;
        r.cfm = l.stop_cfm;
        if (l.lostop == l.histop) {
            r.lo = -INFINITY;
            r.hi = INFINITY;
        } else if (st.limit == 1) {
            r.lo = 0.f;
            r.hi = INFINITY;
        } else {
            r.lo = -INFINITY;
            r.hi = 0.f;
        }
        // dParamBounce is never set by the reference; not supported on the device path.
    }
    return 1;
}

// setFixedOrientation (ODE joint.cpp): three angular rows, rhs from the quaternion error
__device__ __forceinline__ void fixed_orientation_rhs(const DevJoint &jc, const BodyPose &p0, const BodyPose &p1,
                                                      bool has_b1, float k2, float *c3)
{
    float qerr[4], e[3];
    if (has_b1) {
        float qq[4];
        qmul1(qq, p0.q, p1.q);
        qmul2(qerr, qq, jc.qrel);
    } else {
        qmul3(qerr, p0.q, jc.qrel);
    }
    if (qerr[0] < 0.f) {
        qerr[1] = -qerr[1];
        qerr[2] = -qerr[2];
        qerr[3] = -qerr[3];
    }
    mul_R_v(e, p0.R, qerr + 1);
#pragma unroll
    for (int i = 0; i < 3; i++) c3[i] = k2 * e[i];
}

template <int STAGES> __global__ void __launch_bounds__(BD, 8) b2p_step_kernel(const StepParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevTemplate *__restrict__ T = P.T;
    const int nb = T->nb, nj = T->nj, maxc = T->maxc, rpc = T->rpc, njslots = T->njslots, maxrows_c = T->maxrows;
    Smem S;
    {
        unsigned char *p = smem_raw;
        S.ring = (float4 *)p;
        p += sizeof(float4) * (size_t)STAGES * ROWV * BD;
        S.lring = (float *)p;
        p += sizeof(float) * (size_t)STAGES * BD;
        S.fc = (float *)p;
        p += sizeof(float) * (size_t)nb * 6 * BD;
        S.iw = (float *)p;
        p += sizeof(float) * (size_t)nb * 6 * BD;
        S.lamn = (float *)p;
        p += sizeof(float) * (size_t)maxc * BD;
        S.im = (float *)p;
        p += sizeof(float) * (size_t)nb;
        S.jm = p;
        p += (size_t)nj * BD;
        S.jl = (int8_t *)p;
        p += (size_t)nj * BD;
        S.cm = p;
    }
    const int tid = threadIdx.x;
    for (int b = tid; b < nb; b += BD) S.im[b] = T->bodies[b].invMass;
    __syncthreads();
    const int w = blockIdx.x * BD + tid;
    if (w >= P.W) return;
    const int Wp = P.Wp;
    const float h = P.h, hinv = 1.0f / P.h;

    long long t_ph[5] = {0, 0, 0, 0, 0};
    long long t_mark = clock64();
#define PH_MARK(i) do { const long long t_now = clock64(); t_ph[i] += t_now - t_mark; t_mark = t_now; } while (0)
    // ---------------- phase 1: bodies ----------------
    for (int b = 0; b < nb; b++) {
        const DevBody &bc = T->bodies[b];
        BodyPose bp;
        load_pose(P, b, w, bp);
        const int f = b * B2P_STATE_FIELDS;
        float avel[3] = {G(P.state, f + 10), G(P.state, f + 11), G(P.state, f + 12)};
        float facc[3] = {G(P.force, b * 6 + 0), G(P.force, b * 6 + 1), G(P.force, b * 6 + 2)};
        float tacc[3] = {G(P.force, b * 6 + 3), G(P.force, b * 6 + 4), G(P.force, b * 6 + 5)};
        float tmp[9], Iw[9];
        mul33_bt(tmp, bc.invI, bp.R);
        mul33(Iw, bp.R, tmp);
#pragma unroll
        for (int k = 0; k < 9; k++) G(P.invIw, b * 9 + k) = Iw[k];
        S.iw[(size_t)(b * 6 + 0) * BD + tid] = Iw[0];
        S.iw[(size_t)(b * 6 + 1) * BD + tid] = Iw[1];
        S.iw[(size_t)(b * 6 + 2) * BD + tid] = Iw[2];
        S.iw[(size_t)(b * 6 + 3) * BD + tid] = Iw[4];
        S.iw[(size_t)(b * 6 + 4) * BD + tid] = Iw[5];
        S.iw[(size_t)(b * 6 + 5) * BD + tid] = Iw[8];
        if (bc.gyroscopic) {
            // implicit gyroscopic torque (ODE 0.16 quickstep stage0, Lacoursiere 2006)
            float I[9], L[3], It[9], itInv[9];
            mul33_bt(tmp, bc.I, bp.R);
            mul33(I, bp.R, tmp);
            mul_R_v(L, I, avel);
            It[0] = I[0];
            It[1] = I[1] + h * L[2];
            It[2] = I[2] - h * L[1];
            It[3] = I[3] - h * L[2];
            It[4] = I[4];
            It[5] = I[5] + h * L[0];
            It[6] = I[6] + h * L[1];
            It[7] = I[7] - h * L[0];
            It[8] = I[8];
#pragma unroll
            for (int k = 0; k < 3; k++) L[k] *= hinv;
            if (invert33(itInv, It) != 0.f) {
                float M[9], tau[3];
                mul33(M, I, itInv);
                M[0] -= 1.f;
                M[4] -= 1.f;
                M[8] -= 1.f;
                mul_R_v(tau, M, L);
#pragma unroll
                for (int k = 0; k < 3; k++) tacc[k] += tau[k];
            }
        }
        if (!bc.nogravity) {
#pragma unroll
            for (int k = 0; k < 3; k++) facc[k] += bc.mass * P.gravity[k];
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            S.fc[(size_t)(b * 6 + k) * BD + tid] = facc[k];
            S.fc[(size_t)(b * 6 + 3 + k) * BD + tid] = tacc[k];
        }
    }

    // ---------------- phase 1a: limit motors powered into a stop add forces (rare) -------------
    if (T->has_limit_motors) {
        for (int j = 0; j < nj; j++) {
            const DevJoint &jc = T->joints[j];
            if (jc.type == 6 || jc.b0 < 0) continue;
            BodyPose p0, p1;
            load_pose(P, jc.b0, w, p0);
            const bool has_b1 = jc.b1 >= 0;
            if (has_b1) load_pose(P, jc.b1, w, p1);
            LimotState st;
            joint_limit_state(jc, p0, p1, st);
            const float vel = G(P.jcmd, j * 4 + 0), fmax = G(P.jcmd, j * 4 + 1);
            const DevLimot &l = jc.limot1;
            if (fmax > 0.f && st.limit && l.lostop != l.histop) {
                float fm = fmax;
                if ((vel > 0.f) || (vel == 0.f && st.limit == 2)) fm = -fm;
                if ((st.limit == 1 && vel > 0.f) || (st.limit == 2 && vel < 0.f)) fm *= l.fudge_factor;
                float ax1[3];
                mul_R_v(ax1, p0.R, jc.axis1);
                float *f0 = S.fc + (size_t)(jc.b0 * 6) * BD + tid;
                float *f1 = has_b1 ? S.fc + (size_t)(jc.b1 * 6) * BD + tid : nullptr;
                if (jc.type == 3) { // slider: linear force (+ torque decoupling)
                    if (jc.reverse && !has_b1) {
#pragma unroll
                        for (int i = 0; i < 3; i++) ax1[i] = -ax1[i];
                    }
                    if (has_b1) {
                        float c[3], ltd[3];
#pragma unroll
                        for (int i = 0; i < 3; i++) c[i] = 0.5f * (p1.pos[i] - p0.pos[i]);
                        cross3(ltd, c, ax1);
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            f0[(3 + i) * BD] -= fm * ltd[i];
                            f1[(3 + i) * BD] -= fm * ltd[i];
                            f1[i * BD] += fm * ax1[i];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 3; i++) f0[i * BD] -= fm * ax1[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        if (has_b1) f1[(3 + i) * BD] += fm * ax1[i];
                        f0[(3 + i) * BD] -= fm * ax1[i];
                    }
                }
            }
        }
    }

    // ---------------- phase 1b: fe -> gmem (needed by the epilogue), tmp1 -> smem --------------
    for (int b = 0; b < nb; b++) {
        const DevBody &bc = T->bodies[b];
        const int f = b * B2P_STATE_FIELDS;
        float Iw[9], fe[6], t3[3];
        load_invIw(P, b, w, Iw);
#pragma unroll
        for (int k = 0; k < 6; k++) {
            fe[k] = S.fc[(size_t)(b * 6 + k) * BD + tid];
            G(P.force, b * 6 + k) = fe[k];
        }
        mul_R_v(t3, Iw, fe + 3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float lv = G(P.state, f + 7 + k), av = G(P.state, f + 10 + k);
            S.fc[(size_t)(b * 6 + k) * BD + tid] = fe[k] * bc.invMass + lv * hinv;
            S.fc[(size_t)(b * 6 + 3 + k) * BD + tid] = t3[k] + av * hinv;
        }
    }

    // ---------------- phase 2: constraint rows ----------------
    // ODE orders rows with findex < 0 first (in row order) and friction-dependent rows last.
    // Head rows are appended as they are emitted; dependent rows are recorded per contact in a bit
    // mask and appended after all contacts are known.
    int nhead = 0;
    auto push_head = [&](int slot) { G(P.ord, nhead++) = (uint8_t)slot; };

    for (int j = 0; j < nj; j++) {
        const DevJoint &jc = T->joints[j];
        int mrows = 0, limot_row = -1;
        if (jc.b0 >= 0) {
            const bool has_b1 = jc.b1 >= 0;
            BodyPose p0, p1;
            PairCtx pc;
            pc.b0 = jc.b0;
            pc.b1 = jc.b1;
            load_pose(P, jc.b0, w, p0);
            load_invIw(P, jc.b0, w, pc.I0);
            pc.im0 = T->bodies[jc.b0].invMass;
            pc.im1 = 0.f;
            if (has_b1) {
                load_pose(P, jc.b1, w, p1);
                load_invIw(P, jc.b1, w, pc.I1);
                pc.im1 = T->bodies[jc.b1].invMass;
            }
            const int base = jc.slot_base;
            const float kerp = hinv * P.erp;
            Row r;
            if (jc.type == 1) {
                // ---- hinge: setBall + two angular rows + limot (ODE hinge.cpp getInfo2)
                float a1[3], a2[3], cb[3];
                mul_R_v(a1, p0.R, jc.anchor1);
                if (has_b1) {
                    mul_R_v(a2, p1.R, jc.anchor2);
#pragma unroll
                    for (int i = 0; i < 3; i++) cb[i] = kerp * (a2[i] + p1.pos[i] - a1[i] - p0.pos[i]);
                } else {
                    a2[0] = a2[1] = a2[2] = 0.f;
#pragma unroll
                    for (int i = 0; i < 3; i++) cb[i] = kerp * (jc.anchor2[i] - a1[i] - p0.pos[i]);
                }
                // row 0
                row_init(r, P.cfm);
                r.J[0] = 1.f; r.J[4] = a1[2]; r.J[5] = -a1[1];
                if (has_b1) { r.J[6] = -1.f; r.J[10] = -a2[2]; r.J[11] = a2[1]; }
                r.c = cb[0];
                emit_row(P, S, w, tid, base + 0, r, pc, nhead); push_head(base + 0);
                row_init(r, P.cfm);
                r.J[1] = 1.f; r.J[3] = -a1[2]; r.J[5] = a1[0];
                if (has_b1) { r.J[7] = -1.f; r.J[9] = a2[2]; r.J[11] = -a2[0]; }
                r.c = cb[1];
                emit_row(P, S, w, tid, base + 1, r, pc, nhead); push_head(base + 1);
                row_init(r, P.cfm);
                r.J[2] = 1.f; r.J[3] = a1[1]; r.J[4] = -a1[0];
                if (has_b1) { r.J[8] = -1.f; r.J[9] = -a2[1]; r.J[10] = a2[0]; }
                r.c = cb[2];
                emit_row(P, S, w, tid, base + 2, r, pc, nhead); push_head(base + 2);
                float ax1[3], pv[3], qv[3], ax2[3], bv[3];
                mul_R_v(ax1, p0.R, jc.axis1);
                plane_space(ax1, pv, qv);
                if (has_b1) mul_R_v(ax2, p1.R, jc.axis2);
                else { ax2[0] = jc.axis2[0]; ax2[1] = jc.axis2[1]; ax2[2] = jc.axis2[2]; }
                cross3(bv, ax1, ax2);
                row_init(r, P.cfm);
                r.J[3] = pv[0]; r.J[4] = pv[1]; r.J[5] = pv[2];
                if (has_b1) { r.J[9] = -pv[0]; r.J[10] = -pv[1]; r.J[11] = -pv[2]; }
                r.c = kerp * dot3(bv, pv);
                emit_row(P, S, w, tid, base + 3, r, pc, nhead); push_head(base + 3);
                row_init(r, P.cfm);
                r.J[3] = qv[0]; r.J[4] = qv[1]; r.J[5] = qv[2];
                if (has_b1) { r.J[9] = -qv[0]; r.J[10] = -qv[1]; r.J[11] = -qv[2]; }
                r.c = kerp * dot3(bv, qv);
                emit_row(P, S, w, tid, base + 4, r, pc, nhead); push_head(base + 4);
                mrows = 5;
                LimotState st;
                joint_limit_state(jc, p0, p1, st);
                row_init(r, P.cfm);
                if (add_limot_row(P, jc.limot1, G(P.jcmd, j * 4 + 0), G(P.jcmd, j * 4 + 1), st, r, ax1, 1, p0, p1,
                                  has_b1)) {
                    emit_row(P, S, w, tid, base + 5, r, pc, nhead); push_head(base + 5);
                    limot_row = 5;
                    mrows = 6;
                }
            } else if (jc.type == 6) {
                // ---- fixed (ODE fixed.cpp getInfo2): rows 0-2 position, rows 3-5 orientation
                float ofs[3], cl[3], ca[3];
                mul_R_v(ofs, p0.R, jc.offset);
                const float k = hinv * jc.erp;
                if (has_b1) {
#pragma unroll
                    for (int i = 0; i < 3; i++) cl[i] = k * (p1.pos[i] - p0.pos[i] + ofs[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 3; i++) cl[i] = k * (jc.offset[i] - p0.pos[i]);
                }
                fixed_orientation_rhs(jc, p0, p1, has_b1, hinv * P.erp * 2.f, ca);
                row_init(r, P.cfm);
                r.J[0] = 1.f;
                if (has_b1) { r.J[4] = -ofs[2]; r.J[5] = ofs[1]; r.J[6] = -1.f; }
                r.c = cl[0]; r.cfm = jc.cfm;
                emit_row(P, S, w, tid, base + 0, r, pc, nhead); push_head(base + 0);
                row_init(r, P.cfm);
                r.J[1] = 1.f;
                if (has_b1) { r.J[3] = ofs[2]; r.J[5] = -ofs[0]; r.J[7] = -1.f; }
                r.c = cl[1]; r.cfm = jc.cfm;
                emit_row(P, S, w, tid, base + 1, r, pc, nhead); push_head(base + 1);
                row_init(r, P.cfm);
                r.J[2] = 1.f;
                if (has_b1) { r.J[3] = -ofs[1]; r.J[4] = ofs[0]; r.J[8] = -1.f; }
                r.c = cl[2]; r.cfm = jc.cfm;
                emit_row(P, S, w, tid, base + 2, r, pc, nhead); push_head(base + 2);
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    row_init(r, P.cfm);
                    r.J[3 + i] = 1.f;
                    if (has_b1) r.J[9 + i] = -1.f;
                    r.c = ca[i];
                    emit_row(P, S, w, tid, base + 3 + i, r, pc, nhead); push_head(base + 3 + i);
                }
                mrows = 6;
            } else if (jc.type == 3) {
                // ---- slider (ODE slider.cpp getInfo2)
                float ca[3];
                fixed_orientation_rhs(jc, p0, p1, has_b1, hinv * P.erp * 2.f, ca);
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    row_init(r, P.cfm);
                    r.J[3 + i] = 1.f;
                    if (has_b1) r.J[9 + i] = -1.f;
                    r.c = ca[i];
                    emit_row(P, S, w, tid, base + i, r, pc, nhead); push_head(base + i);
                }
                float c[3] = {0.f, 0.f, 0.f}, ax1[3], pv[3], qv[3];
                if (has_b1) {
#pragma unroll
                    for (int i = 0; i < 3; i++) c[i] = p1.pos[i] - p0.pos[i];
                }
                mul_R_v(ax1, p0.R, jc.axis1);
                plane_space(ax1, pv, qv);
                float c3, c4, tp[3], tq[3];
                if (has_b1) {
                    cross3(tp, c, pv);
                    cross3(tq, c, qv);
                    float ofs[3];
                    mul_R_v(ofs, p1.R, jc.offset);
#pragma unroll
                    for (int i = 0; i < 3; i++) c[i] += ofs[i];
                    c3 = kerp * dot3(pv, c);
                    c4 = kerp * dot3(qv, c);
                } else {
                    float ofs[3];
#pragma unroll
                    for (int i = 0; i < 3; i++) ofs[i] = jc.offset[i] - p0.pos[i];
                    c3 = kerp * dot3(pv, ofs);
                    c4 = kerp * dot3(qv, ofs);
                    if (jc.reverse) {
#pragma unroll
                        for (int i = 0; i < 3; i++) ax1[i] = -ax1[i];
                    }
                }
                row_init(r, P.cfm);
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    r.J[i] = pv[i];
                    if (has_b1) { r.J[3 + i] = 0.5f * tp[i]; r.J[9 + i] = 0.5f * tp[i]; r.J[6 + i] = -pv[i]; }
                }
                r.c = c3;
                emit_row(P, S, w, tid, base + 3, r, pc, nhead); push_head(base + 3);
                row_init(r, P.cfm);
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    r.J[i] = qv[i];
                    if (has_b1) { r.J[3 + i] = 0.5f * tq[i]; r.J[9 + i] = 0.5f * tq[i]; r.J[6 + i] = -qv[i]; }
                }
                r.c = c4;
                emit_row(P, S, w, tid, base + 4, r, pc, nhead); push_head(base + 4);
                mrows = 5;
                LimotState st;
                joint_limit_state(jc, p0, p1, st);
                row_init(r, P.cfm);
                if (add_limot_row(P, jc.limot1, G(P.jcmd, j * 4 + 0), G(P.jcmd, j * 4 + 1), st, r, ax1, 0, p0, p1,
                                  has_b1)) {
                    emit_row(P, S, w, tid, base + 5, r, pc, nhead); push_head(base + 5);
                    limot_row = 5;
                    mrows = 6;
                }
            } else if (jc.type == 2 && has_b1) {
                // ---- hinge2 (ODE hinge2.cpp getInfo2): setBall2 along axis1 + one angular row + limots
                float ax1[3], ax2[3], qv[3];
                mul_R_v(ax1, p0.R, jc.axis1);
                mul_R_v(ax2, p1.R, jc.axis2);
                cross3(qv, ax1, ax2);
                const float sn = sqrtf(dot3(qv, qv));
                const float cs = dot3(ax1, ax2);
                normalize3(qv);
                float q1[3], q2[3], a1[3], a2[3], d[3];
                plane_space(ax1, q1, q2);
                mul_R_v(a1, p0.R, jc.anchor1);
                mul_R_v(a2, p1.R, jc.anchor2);
                float a1w[3];
#pragma unroll
                for (int i = 0; i < 3; i++) a1w[i] = a1[i] + p0.pos[i];
#pragma unroll
                for (int i = 0; i < 3; i++) d[i] = a2[i] + p1.pos[i] - a1w[i];
                const float k1 = hinv * jc.susp_erp;
                const float *dirs[3] = {ax1, q1, q2};
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    row_init(r, P.cfm);
                    const float *dv = dirs[i];
                    r.J[0] = dv[0]; r.J[1] = dv[1]; r.J[2] = dv[2];
                    cross3(r.J + 3, a1, dv);
                    r.J[6] = -dv[0]; r.J[7] = -dv[1]; r.J[8] = -dv[2];
                    cross3(r.J + 9, dv, a2);
                    r.c = (i == 0 ? k1 : kerp) * dot3(dv, d);
                    if (i == 0) r.cfm = jc.susp_cfm;
                    emit_row(P, S, w, tid, base + i, r, pc, nhead); push_head(base + i);
                }
                row_init(r, P.cfm);
#pragma unroll
                for (int i = 0; i < 3; i++) { r.J[3 + i] = qv[i]; r.J[9 + i] = -qv[i]; }
                r.c = kerp * (jc.c0 * sn - jc.s0 * cs);
                emit_row(P, S, w, tid, base + 3, r, pc, nhead); push_head(base + 3);
                mrows = 4;
                LimotState st;
                joint_limit_state(jc, p0, p1, st);
                row_init(r, P.cfm);
                if (add_limot_row(P, jc.limot1, G(P.jcmd, j * 4 + 0), G(P.jcmd, j * 4 + 1), st, r, ax1, 1, p0, p1,
                                  true)) {
                    emit_row(P, S, w, tid, base + mrows, r, pc, nhead); push_head(base + mrows);
                    limot_row = mrows;
                    mrows++;
                }
                LimotState st2;
                st2.limit = 0;
                st2.limit_err = 0.f;
                row_init(r, P.cfm);
                if (add_limot_row(P, jc.limot2, G(P.jcmd, j * 4 + 2), G(P.jcmd, j * 4 + 3), st2, r, ax2, 1, p0, p1,
                                  true)) {
                    emit_row(P, S, w, tid, base + mrows, r, pc, nhead); push_head(base + mrows);
                    if (limot_row < 0) limot_row = mrows;
                    mrows++;
                }
            }
        }
        S.jm[(size_t)j * BD + tid] = (uint8_t)mrows;
        S.jl[(size_t)j * BD + tid] = (int8_t)limot_row;
        G(P.jrows, j) = (uint8_t)mrows;
    }

    // ---- contacts (ODE contact.cpp getInfo1/getInfo2 for the mode bits Frame::addContact sets)
    int nc = G(P.ccount, 0);
    if (nc > maxc) nc = maxc;
    for (int c = 0; c < nc; c++) {
        const int ids = __float_as_int(G(P.crec, c * B2P_CONTACT_FIELDS + CR_IDS));
        const int b0 = ids & 0xff;
        const int b1 = ((ids >> 8) & 0xff) == 0xff ? -1 : ((ids >> 8) & 0xff);
        const int mode = __float_as_int(G(P.crec, c * B2P_CONTACT_FIELDS + CR_MODE));
        float pos[3] = {G(P.crec, c * B2P_CONTACT_FIELDS + CR_PX), G(P.crec, c * B2P_CONTACT_FIELDS + CR_PY),
                        G(P.crec, c * B2P_CONTACT_FIELDS + CR_PZ)};
        float n[3] = {G(P.crec, c * B2P_CONTACT_FIELDS + CR_NX), G(P.crec, c * B2P_CONTACT_FIELDS + CR_NY),
                      G(P.crec, c * B2P_CONTACT_FIELDS + CR_NZ)};
        float depth = G(P.crec, c * B2P_CONTACT_FIELDS + CR_DEPTH) - P.contact_min_depth;
        float mu = G(P.crec, c * B2P_CONTACT_FIELDS + CR_MU);
        float mu2 = G(P.crec, c * B2P_CONTACT_FIELDS + CR_MU2);
        const float soft_erp = G(P.crec, c * B2P_CONTACT_FIELDS + CR_SOFT_ERP);
        const float soft_cfm = G(P.crec, c * B2P_CONTACT_FIELDS + CR_SOFT_CFM);
        if (depth < 0.f) depth = 0.f;
        if (mu < 0.f) mu = 0.f;
        if (mu2 < 0.f) mu2 = 0.f;
        PairCtx pc;
        pc.b0 = b0;
        pc.b1 = b1;
        BodyPose p0, p1;
        load_pose(P, b0, w, p0);
        load_invIw(P, b0, w, pc.I0);
        pc.im0 = T->bodies[b0].invMass;
        pc.im1 = 0.f;
        if (b1 >= 0) {
            load_pose(P, b1, w, p1);
            load_invIw(P, b1, w, pc.I1);
            pc.im1 = T->bodies[b1].invMass;
        }
        const int base = njslots + c * rpc;
        const float erp = (mode & 0x008) ? soft_erp : P.erp;
        const float pushout = hinv * erp * depth;
        float c1[3], c2[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 3; i++) c1[i] = pos[i] - p0.pos[i];
        Row r;
        row_init(r, P.cfm);
        r.J[0] = n[0]; r.J[1] = n[1]; r.J[2] = n[2];
        cross3(r.J + 3, c1, n);
        if (b1 >= 0) {
#pragma unroll
            for (int i = 0; i < 3; i++) { c2[i] = pos[i] - p1.pos[i]; r.J[6 + i] = -n[i]; }
            cross3(r.J + 9, n, c2);
        }
        r.c = pushout > P.contact_max_vel ? P.contact_max_vel : pushout;
        if (mode & 0x010) r.cfm = soft_cfm;
        r.lo = 0.f;
        r.hi = INFINITY;
        r.nr = c + 1;
        emit_row(P, S, w, tid, base, r, pc, nhead);
        push_head(base);
        int row = 1, depmask = 0;
        float t1[3], t2[3];
        plane_space(n, t1, t2);
        const bool axisdep = (mode & 0x001) != 0;
        const bool fr1 = mu > 0.f;
        const float mu2e = axisdep ? mu2 : mu;
        const bool fr2 = axisdep ? (mu2 > 0.f) : (mu > 0.f);
        if (fr1) {
            row_init(r, P.cfm);
            r.J[0] = t1[0]; r.J[1] = t1[1]; r.J[2] = t1[2];
            cross3(r.J + 3, c1, t1);
            if (b1 >= 0) { r.J[6] = -t1[0]; r.J[7] = -t1[1]; r.J[8] = -t1[2]; cross3(r.J + 9, t1, c2); }
            r.lo = -mu;
            r.hi = mu;
            if (mode & 0x1000) r.fr = c + 1;
            if (mode & 0x1000) {
                depmask |= 1 << (row - 1);
                emit_row(P, S, w, tid, base + row, r, pc);
            } else {
                emit_row(P, S, w, tid, base + row, r, pc, nhead);
                push_head(base + row);
            }
            row++;
        }
        if (fr2) {
            row_init(r, P.cfm);
            r.J[0] = t2[0]; r.J[1] = t2[1]; r.J[2] = t2[2];
            cross3(r.J + 3, c1, t2);
            if (b1 >= 0) { r.J[6] = -t2[0]; r.J[7] = -t2[1]; r.J[8] = -t2[2]; cross3(r.J + 9, t2, c2); }
            r.lo = -mu2e;
            r.hi = mu2e;
            if (mode & 0x2000) r.fr = c + 1;
            if (mode & 0x2000) {
                depmask |= 1 << (row - 1);
                emit_row(P, S, w, tid, base + row, r, pc);
            } else {
                emit_row(P, S, w, tid, base + row, r, pc, nhead);
                push_head(base + row);
            }
            row++;
        }
        if ((mode & 0x400) && rpc >= 6) {
            float rho[3];
            rho[0] = G(P.crec, c * B2P_CONTACT_FIELDS + CR_RHO);
            if (axisdep) {
                rho[1] = G(P.crec, c * B2P_CONTACT_FIELDS + CR_RHO2);
                rho[2] = G(P.crec, c * B2P_CONTACT_FIELDS + CR_RHON);
            } else {
                rho[1] = rho[0];
                rho[2] = rho[0];
            }
            const float *ax[3] = {t1, t2, n};
            const int bits[3] = {0x1000, 0x2000, 0x4000};
#pragma unroll
            for (int i = 0; i < 3; i++) {
                if (rho[i] > 0.f) {
                    row_init(r, P.cfm);
                    r.J[3] = ax[i][0]; r.J[4] = ax[i][1]; r.J[5] = ax[i][2];
                    if (b1 >= 0) { r.J[9] = -ax[i][0]; r.J[10] = -ax[i][1]; r.J[11] = -ax[i][2]; }
                    r.lo = -rho[i];
                    r.hi = rho[i];
                    if (mode & bits[i]) r.fr = c + 1;
                    if (mode & bits[i]) {
                        depmask |= 1 << (row - 1);
                        emit_row(P, S, w, tid, base + row, r, pc);
                    } else {
                        emit_row(P, S, w, tid, base + row, r, pc, nhead);
                        push_head(base + row);
                    }
                    row++;
                }
            }
        }
        S.cm[(size_t)c * BD + tid] = (uint8_t)(row | (depmask << 3));
    }

    int m = nhead;
    for (int c = 0; c < nc; c++) {
        const int v = S.cm[(size_t)c * BD + tid];
        const int rows_c = v & 7, depmask = v >> 3;
        for (int k = 1; k < rows_c; k++)
            if (depmask & (1 << (k - 1))) G(P.ord, m++) = (uint8_t)(njslots + c * rpc + k);
    }
    G(P.nrows, 0) = m;

    // tmp1 -> gmem? not needed: epilogue uses fe from gmem.  fc = 0
    for (int k = 0; k < nb * 6; k++) S.fc[(size_t)k * BD + tid] = 0.f;
    for (int k = 0; k < maxc; k++) S.lamn[(size_t)k * BD + tid] = 0.f;

    PH_MARK(0);
    // ---------------- phase 3: SOR-LCP sweep ----------------
    // Each world follows its own row order (per-world contact count and RNG state).  Reading rows
    // through a per-lane index would make every warp load touch 32 different lines, and L1TEX then
    // serialises 32 wavefronts per instruction (measured: profiles/r1_v4_*).  Instead the rows are
    // physically permuted into sweep order whenever the order changes (iteration 0 and ODE's
    // reshuffles at 8, 16): scratch B is POSITION-major SoA, so position k of all 32 worlds of a warp
    // is one coalesced line per vector even though every lane holds a different constraint there.
    // The scattered traffic is paid 3 times per step instead of 20.
    uint32_t seed = G(P.seed, 0);
    float4 *const Bq = P.rowsB + w;  // B[pos][v][world]
    float *const Lq = P.lambdaB + w; // lambda by sweep position: L[pos][world]
    // The row order lives in global memory ([pos][world] bytes) and is mirrored into the (idle) ring
    // buffer of shared memory while it is being shuffled / used for permuting: the Fisher-Yates swaps
    // are dependent random accesses, far too slow against global memory.
    // The mirror must stay inside THIS thread's ring slots (float4 index (j*BD + tid), j < STAGES*ROWV):
    // warps of a CTA are in different phases, so a [pos][BD] byte layout would let one warp's order
    // bytes land in the ring slots another warp is streaming rows through.
    uint8_t *const sbase = reinterpret_cast<uint8_t *>(S.ring + tid);
    auto sord = [&](int pos) -> uint8_t & { return sbase[(size_t)(pos >> 4) * BD * 16 + (pos & 15)]; };
    const bool ord_fits = maxrows_c <= STAGES * ROWV * 16;
    auto ord_load = [&]() {
        if (ord_fits)
            for (int pos = 0; pos < m; pos++) sord(pos) = G(P.ord, pos);
    };
    auto ord_at = [&](int pos) -> int { return ord_fits ? (int)sord(pos) : (int)G(P.ord, pos); };
    auto ord_swap = [&](int a, int b) {
        if (ord_fits) {
            const uint8_t x = sord(a), y = sord(b);
            sord(a) = y;
            sord(b) = x;
        } else {
            const uint8_t x = G(P.ord, a), y = G(P.ord, b);
            G(P.ord, a) = y;
            G(P.ord, b) = x;
        }
    };
    auto ord_store = [&]() {
        if (ord_fits)
            for (int pos = 0; pos < m; pos++) G(P.ord, pos) = sord(pos);
    };
    // lambda by position -> the row's slot in scratch A (scattered 4-byte stores; 3 times per step)
    auto writeback_lambda = [&]() {
        for (int pos = 0; pos < m; pos++)
            reinterpret_cast<float *>(P.rows + (size_t)(ord_at(pos) * ROWV + 3) * Wp + w)[1] = Lq[(size_t)pos * Wp];
    };
    // A (slot-major) -> B (position-major) for positions [start, m): reads are scattered by the per-lane
    // slot index (32 lines per warp load), which is why this is done per reshuffle and not per sweep.
    auto permute_rows = [&](int start) {
        for (int pos = start; pos < m; pos++) {
            const int i = ord_at(pos);
            const float4 *src = P.rows + (size_t)(i * ROWV) * Wp + w;
            const float4 q0 = src[0], q1 = src[(size_t)Wp], q2 = src[2 * (size_t)Wp], q3 = src[3 * (size_t)Wp];
            float4 *dst = Bq + (size_t)(pos * ROWV) * Wp;
            dst[0] = q0;
            dst[(size_t)Wp] = q1;
            dst[2 * (size_t)Wp] = q2;
            dst[3 * (size_t)Wp] = q3;
            Lq[(size_t)pos * Wp] = q3.y;
        }
    };
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(S.ring + tid);
    const uint32_t lring_s = (uint32_t)__cvta_generic_to_shared(S.lring + tid);
    for (int it = 0; it < P.iters; it++) {
        const bool reshuffle = P.reorder == 0 && it >= 8 && (it & 7) == 0;
        PH_MARK(2);
        if (reshuffle) {
            ord_load();
            writeback_lambda();
            // ODE ConstraintsReorderingHelper on [0,nhead) and [nhead,m)
            for (int idx = 1; idx < nhead; idx++) ord_swap(idx, rand_int(seed, idx + 1));
            for (int idx = nhead + 1; idx < m; idx++) ord_swap(idx, rand_int(seed, idx - nhead + 1) + nhead);
            ord_store();
            permute_rows(0);
        } else if (it == 0) {
            // the head rows are already in place (written by emit_row)
            ord_load();
            permute_rows(nhead);
            for (int pos = 0; pos < nhead; pos++) Lq[(size_t)pos * Wp] = 0.f;
        }
        PH_MARK(1);
        // Sweep.  Rows stream through a per-thread cp.async ring in shared memory (STAGES-1 rows in
        // flight per thread, tracked by async groups rather than by the six scoreboard slots, which is
        // what limited register prefetching: profiles/r1_v6_*).  Row loads do not depend on the
        // sequential fc/lambda state, so the ring runs ahead freely.
        auto issue = [&](int pos) {
            const int st = pos % STAGES;
            const float4 *src = Bq + (size_t)(pos * ROWV) * Wp;
            const uint32_t d = ring_s + (uint32_t)(st * ROWV * BD * 16);
            cp_async16(d, src);
            cp_async16(d + BD * 16, src + (size_t)Wp);
            cp_async16(d + 2 * BD * 16, src + 2 * (size_t)Wp);
            cp_async16(d + 3 * BD * 16, src + 3 * (size_t)Wp);
            cp_async4(lring_s + (uint32_t)(st * BD * 4), Lq + (size_t)pos * Wp);
        };
#pragma unroll
        for (int p0 = 0; p0 < STAGES - 1; p0++) {
            if (p0 < m) issue(p0);
            cp_async_commit();
        }
        for (int k = 0; k < m; k++) {
            if (k + STAGES - 1 < m) issue(k + STAGES - 1);
            cp_async_commit();
            cp_async_wait<STAGES - 1>();
            const int st = k % STAGES;
            const float4 *rq = S.ring + (size_t)(st * ROWV) * BD + tid;
            const float4 c0 = rq[0], c1 = rq[BD], c2 = rq[2 * BD], c3 = rq[3 * BD];
            const float lam_old = S.lring[(size_t)st * BD + tid];
            const int meta = __float_as_int(c3.z);
            const int b0 = meta & 0xff, b1r = (meta >> 8) & 0xff, fr = (meta >> 16) & 0xff, nr = (meta >> 24) & 0xff;
            const bool two = b1r != 0xff;
            const int b1 = two ? b1r : b0;
            const float s2 = two ? 1.f : 0.f;
            float *f0 = S.fc + (size_t)(b0 * 6) * BD + tid;
            float *f1 = S.fc + (size_t)(b1 * 6) * BD + tid;
            // unscaled J: J1l = c0.xyz, J1a = (c0.w, c1.x, c1.y), J2l = -J1l, J2a = (c1.z, c1.w, c2.x)
            const float dl = (f0[0 * BD] - s2 * f1[0 * BD]) * c0.x + (f0[1 * BD] - s2 * f1[1 * BD]) * c0.y +
                             (f0[2 * BD] - s2 * f1[2 * BD]) * c0.z;
            const float da = f0[3 * BD] * c0.w + f0[4 * BD] * c1.x + f0[5 * BD] * c1.y + f1[3 * BD] * c1.z +
                             f1[4 * BD] * c1.w + f1[5 * BD] * c2.x; // J2a == 0 for one-body rows
            float delta = c3.x * (c2.y - lam_old * c2.z - (dl + da));
            float lo_act, hi_act;
            const int enc = __float_as_int(c2.w);
            if (fr) {
                // friction row: bounds follow the normal row's current lambda (same sweep), mirrored in
                // shared memory by contact index
                hi_act = fabsf(c2.w * S.lamn[(size_t)(fr - 1) * BD + tid]);
                lo_act = -hi_act;
            } else {
                hi_act = enc == (int)0xFFFFFFFF ? INFINITY : (enc == (int)0xFFFFFFFE ? 0.f : c2.w);
                lo_act = enc == (int)0xFFFFFFFF ? 0.f : (enc == (int)0xFFFFFFFE ? -INFINITY : -c2.w);
            }
            const float un = lam_old + delta;
            const float nl = fminf(fmaxf(un, lo_act), hi_act);
            delta = (nl == un) ? delta : nl - lam_old;
            if (nr) S.lamn[(size_t)(nr - 1) * BD + tid] = nl;
            Lq[(size_t)k * Wp] = nl;
            // fc += delta * invM J^T, with invI_world (symmetric) from shared memory
            {
                const float dm = delta * S.im[b0];
                const float *Iw = S.iw + (size_t)(b0 * 6) * BD + tid;
                const float a0 = delta * c0.w, a1 = delta * c1.x, a2 = delta * c1.y;
                const float ixx = Iw[0], ixy = Iw[BD], ixz = Iw[2 * BD], iyy = Iw[3 * BD], iyz = Iw[4 * BD],
                            izz = Iw[5 * BD];
                f0[0 * BD] += dm * c0.x;
                f0[1 * BD] += dm * c0.y;
                f0[2 * BD] += dm * c0.z;
                f0[3 * BD] += ixx * a0 + ixy * a1 + ixz * a2;
                f0[4 * BD] += ixy * a0 + iyy * a1 + iyz * a2;
                f0[5 * BD] += ixz * a0 + iyz * a1 + izz * a2;
            }
            if (two) {
                const float dm = delta * S.im[b1];
                const float *Iw = S.iw + (size_t)(b1 * 6) * BD + tid;
                const float a0 = delta * c1.z, a1 = delta * c1.w, a2 = delta * c2.x;
                const float ixx = Iw[0], ixy = Iw[BD], ixz = Iw[2 * BD], iyy = Iw[3 * BD], iyz = Iw[4 * BD],
                            izz = Iw[5 * BD];
                f1[0 * BD] -= dm * c0.x;
                f1[1 * BD] -= dm * c0.y;
                f1[2 * BD] -= dm * c0.z;
                f1[3 * BD] += ixx * a0 + ixy * a1 + ixz * a2;
                f1[4 * BD] += ixy * a0 + iyy * a1 + iyz * a2;
                f1[5 * BD] += ixz * a0 + iyz * a1 + izz * a2;
            }
        }
        cp_async_wait<0>();
    }
    PH_MARK(2);
    ord_load();
    writeback_lambda();
    G(P.seed, 0) = seed;
    PH_MARK(3);

    // ---------------- phase 4a: joint / contact feedback from the unscaled Jacobian ------------
    for (int j = 0; j < nj; j++) {
        const DevJoint &jc = T->joints[j];
        const int mrows = S.jm[(size_t)j * BD + tid];
        const int lrow = S.jl[(size_t)j * BD + tid];
        float data[12];
#pragma unroll
        for (int q = 0; q < 12; q++) data[q] = 0.f;
        float flam = 0.f;
        for (int k = 0; k < mrows; k++) {
            const int slot = jc.slot_base + k;
            const float4 *src = P.rows + (size_t)(slot * ROWV) * Wp + w;
            const float4 j0 = src[0], j1 = src[(size_t)Wp], j2 = src[2 * (size_t)Wp];
            const float lam = src[3 * (size_t)Wp].y;
            data[0] += j0.x * lam; data[1] += j0.y * lam; data[2] += j0.z * lam;
            data[3] += j0.w * lam; data[4] += j1.x * lam; data[5] += j1.y * lam;
            data[6] -= j0.x * lam; data[7] -= j0.y * lam; data[8] -= j0.z * lam;
            data[9] += j1.z * lam; data[10] += j1.w * lam; data[11] += j2.x * lam;
            if (k == lrow) flam = lam;
        }
        if (jc.b1 < 0) {
#pragma unroll
            for (int q = 6; q < 12; q++) data[q] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 12; q++) G(P.jfb, j * 13 + q) = data[q];
        G(P.jfb, j * 13 + 12) = flam;
    }
    for (int c = 0; c < nc; c++) {
        const int mrows = S.cm[(size_t)c * BD + tid] & 7;
        float f[3] = {0.f, 0.f, 0.f};
        for (int k = 0; k < mrows; k++) {
            const int slot = njslots + c * rpc + k;
            const float4 j0 = P.rows[(size_t)(slot * ROWV) * Wp + w];
            const float lam = P.rows[(size_t)(slot * ROWV + 3) * Wp + w].y;
            f[0] += j0.x * lam;
            f[1] += j0.y * lam;
            f[2] += j0.z * lam;
        }
        G(P.cfb, c * 3 + 0) = f[0];
        G(P.cfb, c * 3 + 1) = f[1];
        G(P.cfb, c * 3 + 2) = f[2];
    }

    // ---------------- phase 4b: velocity update + dxStepBody ----------------
    for (int b = 0; b < nb; b++) {
        const DevBody &bc = T->bodies[b];
        const int f = b * B2P_STATE_FIELDS;
        float pos[3], q[4], lv[3], av[3], fe[6], Iw[9], t3[3];
        pos[0] = G(P.state, f + 0); pos[1] = G(P.state, f + 1); pos[2] = G(P.state, f + 2);
        q[0] = G(P.state, f + 3); q[1] = G(P.state, f + 4); q[2] = G(P.state, f + 5); q[3] = G(P.state, f + 6);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lv[k] = G(P.state, f + 7 + k);
            av[k] = G(P.state, f + 10 + k);
        }
#pragma unroll
        for (int k = 0; k < 6; k++) fe[k] = G(P.force, b * 6 + k);
        load_invIw(P, b, w, Iw);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lv[k] += h * S.fc[(size_t)(b * 6 + k) * BD + tid];
            av[k] += h * S.fc[(size_t)(b * 6 + 3 + k) * BD + tid];
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lv[k] += h * bc.invMass * fe[k];
            fe[3 + k] *= h;
        }
        mul_R_v(t3, Iw, fe + 3);
#pragma unroll
        for (int k = 0; k < 3; k++) av[k] += t3[k];
        if (bc.has_max_ang) {
            const float aspeed = dot3(av, av);
            if (aspeed > bc.max_angular_speed * bc.max_angular_speed) {
                const float coef = bc.max_angular_speed / sqrtf(aspeed);
#pragma unroll
                for (int k = 0; k < 3; k++) av[k] *= coef;
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) pos[k] += h * lv[k];
        float dq[4];
        dq[0] = 0.5f * (-av[0] * q[1] - av[1] * q[2] - av[2] * q[3]);
        dq[1] = 0.5f * (av[0] * q[0] + av[1] * q[3] - av[2] * q[2]);
        dq[2] = 0.5f * (-av[0] * q[3] + av[1] * q[0] + av[2] * q[1]);
        dq[3] = 0.5f * (av[0] * q[2] - av[1] * q[1] + av[2] * q[0]);
#pragma unroll
        for (int k = 0; k < 4; k++) q[k] += h * dq[k];
        normalize4(q);
        G(P.state, f + 0) = pos[0]; G(P.state, f + 1) = pos[1]; G(P.state, f + 2) = pos[2];
        G(P.state, f + 3) = q[0]; G(P.state, f + 4) = q[1]; G(P.state, f + 5) = q[2]; G(P.state, f + 6) = q[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            G(P.state, f + 7 + k) = lv[k];
            G(P.state, f + 10 + k) = av[k];
        }
#pragma unroll
        for (int k = 0; k < 6; k++) G(P.force, b * 6 + k) = 0.f;
    }

    // ---------------- phase 4c: joint observations from the new state ----------------
    for (int j = 0; j < nj; j++) {
        const DevJoint &jc = T->joints[j];
        float a1 = 0.f, r1 = 0.f, a2 = 0.f, r2 = 0.f;
        if (jc.b0 >= 0 && jc.type != 6) {
            const bool has_b1 = jc.b1 >= 0;
            BodyPose p0, p1;
            load_pose(P, jc.b0, w, p0);
            float lv0[3], av0[3], lv1[3] = {0.f, 0.f, 0.f}, av1[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                lv0[k] = G(P.state, jc.b0 * B2P_STATE_FIELDS + 7 + k);
                av0[k] = G(P.state, jc.b0 * B2P_STATE_FIELDS + 10 + k);
            }
            if (has_b1) {
                load_pose(P, jc.b1, w, p1);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    lv1[k] = G(P.state, jc.b1 * B2P_STATE_FIELDS + 7 + k);
                    av1[k] = G(P.state, jc.b1 * B2P_STATE_FIELDS + 10 + k);
                }
            }
            float axis[3];
            mul_R_v(axis, p0.R, jc.axis1);
            if (jc.type == 1) {
                a1 = hinge_angle(jc, p0.q, p1.q, has_b1);
                r1 = dot3(axis, av0);
                if (has_b1) r1 -= dot3(axis, av1);
                if (jc.reverse) { a1 = -a1; r1 = -r1; }
            } else if (jc.type == 3) {
                a1 = slider_position(jc, p0, p1, has_b1);
                if (has_b1) r1 = dot3(axis, lv0) - dot3(axis, lv1);
                else { r1 = dot3(axis, lv0); if (jc.reverse) r1 = -r1; }
            } else if (jc.type == 2 && has_b1) {
                a1 = hinge2_angle1(jc, p0, p1);
                r1 = dot3(axis, av0) - dot3(axis, av1);
                a2 = hinge2_angle2(jc, p0, p1);
                float axis2[3];
                mul_R_v(axis2, p1.R, jc.axis2);
                r2 = dot3(axis2, av0) - dot3(axis2, av1);
            }
        }
        G(P.jobs, j * 4 + 0) = a1;
        G(P.jobs, j * 4 + 1) = r1;
        G(P.jobs, j * 4 + 2) = a2;
        G(P.jobs, j * 4 + 3) = r2;
    }
    PH_MARK(4);
    if (P.phase_cycles && (tid & 31) == 0) {
        for (int i = 0; i < 5; i++) atomicAdd((unsigned long long *)&P.phase_cycles[i], (unsigned long long)t_ph[i]);
        atomicAdd((unsigned long long *)&P.phase_cycles[5], 1ull);
    }
}

} // namespace

static size_t smem_bytes_for(int nb, int nj, int maxc, int stages)
{
    size_t s = sizeof(float4) * (size_t)stages * ROWV * BD + sizeof(float) * (size_t)stages * BD;
    s += sizeof(float) * (size_t)nb * 12 * BD;
    s += sizeof(float) * (size_t)maxc * BD;
    s += sizeof(float) * (size_t)nb;
    s += (size_t)nj * BD * 2;
    s += (size_t)maxc * BD;
    return (s + 15) & ~(size_t)15;
}

static const int kStageChoices[] = {12, 8, 6, 4, 3, 2};

// deepest ring such that ceil(worlds / (64 * 148)) CTAs fit on an SM (all worlds resident: one wave)
static int pick_stages(int nb, int nj, int maxc, int worlds)
{
    const int num_sms = 148;
    int ctas_per_sm = (worlds + BD * num_sms - 1) / (BD * num_sms);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8; // register limit; larger launches run in waves
    const size_t budget = (size_t)227 * 1024 / ctas_per_sm - 1024;
    for (int st : kStageChoices)
        if (smem_bytes_for(nb, nj, maxc, st) <= budget) return st;
    return 2;
}

extern "C" size_t b2p_step_smem_bytes(int nb, int nj, int maxc, int rpc)
{
    (void)rpc;
    return smem_bytes_for(nb, nj, maxc, 2);
}

extern "C" int b2p_step_block_dim(void) { return BD; }

template <int STAGES> static cudaError_t launch_t(const StepParams *P, size_t smem, cudaStream_t stream)
{
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(b2p_step_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int grid = (P->W + BD - 1) / BD;
    b2p_step_kernel<STAGES><<<grid, BD, smem, stream>>>(*P);
    return cudaGetLastError();
}

extern "C" cudaError_t b2p_launch_step(const StepParams *P, int nb, int nj, int maxc, int rpc, cudaStream_t stream)
{
    (void)rpc;
    const int st = pick_stages(nb, nj, maxc, P->W);
    const size_t smem = smem_bytes_for(nb, nj, maxc, st);
    switch (st) {
    case 12: return launch_t<12>(P, smem, stream);
    case 8: return launch_t<8>(P, smem, stream);
    case 6: return launch_t<6>(P, smem, stream);
    case 4: return launch_t<4>(P, smem, stream);
    case 3: return launch_t<3>(P, smem, stream);
    default: return launch_t<2>(P, smem, stream);
    }
}

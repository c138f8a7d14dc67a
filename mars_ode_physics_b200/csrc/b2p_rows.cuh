// b2p_rows.cuh -- the constraint rows of one joint (ODE 0.16 getInfo1 / getInfo2 of hinge.cpp, hinge2.cpp,
// slider.cpp, fixed.cpp, ball.cpp, universal.cpp and dxJointLimitMotor::addLimot of joint.cpp), written once and
// used twice:
//   * b2p_assemble.cu   FULL = true : Jacobian, right-hand side, cfm and bounds of every row; the emitter scales
//                                     the row and stores it for the SOR-LCP kernel
//   * b2p_integrate.cu  FULL = false: the Jacobian only; the emitter accumulates lambda * J into the joint's
//                                     dJointFeedback (src/joints/Joint.cpp:268,859-1019), so that the feedback
//                                     needs no second pass over the stored rows
// Both passes see the same pre-step poses, and the second one takes the "which limit / motor rows exist" bits
// from the first (jrows), so that the two always agree on the slot numbering.
#pragma once
#include "b2p_kern.cuh"

// One constraint row under construction (ODE Info2 view).
struct Row {
    float J[12];
    float c, cfm, lo, hi;
    int fr;   // slot of the contact's normal row + 1 if this row's bounds follow it (findex), else 0
    float sn; // sqrt(Ad) of that normal row
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
    r.sn = 1.f;
}

// jrows byte of a joint: rows of the last step | limit/motor row of axis 1 present | ... of axis 2 present
#define B2P_JR_ROWS(x) ((x) & 7)
#define B2P_JR_L1 8
#define B2P_JR_L2 16

struct LimotCmd {
    float vel1, fmax1, vel2, fmax2;
};

// dxJointLimitMotor::addLimot, Jacobian part
__device__ __forceinline__ void limot_jacobian(Row &r, const float *ax1, int rotational, const BodyPose &p0,
                                               const BodyPose &p1, bool has_b1)
{
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
}

// dxJointLimitMotor::addLimot, right-hand side / cfm / bounds.  The "powered while at a limit" force injection is
// done by the assemble kernel's pre-pass before tmp1 exists.  `b0`, `b1`: body indices for the bounce velocity.
__device__ __forceinline__ void limot_rhs(const StepParams &P, const DevLimot &l, float vel, float fmax,
                                          const LimotState &st, Row &r, const float *ax1, int rotational, int b0,
                                          int b1, int w)
{
    int powered = fmax > 0.f;
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
        } else {
            if (st.limit == 1) {
                r.lo = 0.f;
                r.hi = INFINITY;
            } else {
                r.lo = -INFINITY;
                r.hi = 0.f;
            }
            if (l.bounce > 0.f) {
                // dParamBounce: the stop reflects the incoming joint velocity
                const int Wp = P.Wp;
                const int f0 = b0 * B2P_STATE_FIELDS + (rotational ? 10 : 7);
                float v0[3] = {G(P.state, f0 + 0), G(P.state, f0 + 1), G(P.state, f0 + 2)};
                float jv = dot3(v0, ax1);
                if (b1 >= 0) {
                    const int f1 = b1 * B2P_STATE_FIELDS + (rotational ? 10 : 7);
                    float v1[3] = {G(P.state, f1 + 0), G(P.state, f1 + 1), G(P.state, f1 + 2)};
                    jv -= dot3(v1, ax1);
                }
                if (st.limit == 1) {
                    if (jv < 0.f) {
                        const float newc = -l.bounce * jv;
                        if (newc > r.c) r.c = newc;
                    }
                } else {
                    if (jv > 0.f) {
                        const float newc = -l.bounce * jv;
                        if (newc < r.c) r.c = newc;
                    }
                }
            }
        }
    }
}

// One limit/motor row.  FULL: decides from the command and the limit state whether the row exists; otherwise
// `present` says so.  Returns 1 if the row was emitted.
template <bool FULL, class Emit>
__device__ __forceinline__ int limot_row(const StepParams &P, const DevLimot &l, float vel, float fmax,
                                         const LimotState &st, bool present, const float *ax1, int rotational,
                                         const BodyPose &p0, const BodyPose &p1, int b0, int b1, int w, Emit &emit)
{
    if (FULL) present = (fmax > 0.f) || st.limit;
    if (!present) return 0;
    Row r;
    row_init(r, P.cfm);
    limot_jacobian(r, ax1, rotational, p0, p1, b1 >= 0);
    if (FULL) limot_rhs(P, l, vel, fmax, st, r, ax1, rotational, b0, b1, w);
    emit(r);
    return 1;
}

// setBall (ODE joint.cpp): three rows J1l = I, J1a = -[a1]x, J2l = -I, J2a = [a2]x
template <bool FULL, class Emit>
__device__ __forceinline__ void ball_rows(const float *a1, const float *a2, const float *cb, bool has_b1, float cfm,
                                          Emit &emit)
{
    Row r;
    row_init(r, cfm);
    r.J[0] = 1.f; r.J[4] = a1[2]; r.J[5] = -a1[1];
    if (has_b1) { r.J[6] = -1.f; r.J[10] = -a2[2]; r.J[11] = a2[1]; }
    r.c = cb[0];
    emit(r);
    row_init(r, cfm);
    r.J[1] = 1.f; r.J[3] = -a1[2]; r.J[5] = a1[0];
    if (has_b1) { r.J[7] = -1.f; r.J[9] = a2[2]; r.J[11] = -a2[0]; }
    r.c = cb[1];
    emit(r);
    row_init(r, cfm);
    r.J[2] = 1.f; r.J[3] = a1[1]; r.J[4] = -a1[0];
    if (has_b1) { r.J[8] = -1.f; r.J[9] = -a2[1]; r.J[10] = a2[0]; }
    r.c = cb[2];
    emit(r);
}

// getInfo1: the number of rows joint `jc` will produce (what joint_rows<true> then emits, decided by the same tests)
__device__ __forceinline__ int joint_row_count(const DevJoint &jc, const BodyPose &p0, const BodyPose &p1,
                                               const LimotCmd &cmd)
{
    if (jc.b0 < 0) return 0;
    int base;
    switch (jc.type) {
    case 1: base = 5; break;
    case 2: if (jc.b1 < 0) return 0; base = 4; break;
    case 3: base = 5; break;
    case 4: return 3;
    case 5: base = 4; break;
    case 6: return 6;
    default: return 0;
    }
    LimotState st;
    joint_limit_state(jc, p0, p1, st);
    int rows = base + ((cmd.fmax1 > 0.f || st.limit) ? 1 : 0);
    if (jc.type == 2) {
        rows += cmd.fmax2 > 0.f ? 1 : 0; // ODE's hinge2 has no stop on its second axis
    } else if (jc.type == 5) {
        LimotState st2;
        joint_limit_state2(jc, p0, p1, st2);
        rows += (cmd.fmax2 > 0.f || st2.limit) ? 1 : 0;
    }
    return rows;
}

// All rows of joint `jc` in getInfo2 order, each handed to emit(const Row &).  Returns the joint's jrows byte.
// `p0`, `p1`: pre-step poses of its bodies (p1 unused without a second body); `flags_in`: the jrows byte of the
// FULL pass (only read when FULL is false).
template <bool FULL, class Emit>
__device__ __forceinline__ int joint_rows(const StepParams &P, const DevJoint &jc, const BodyPose &p0,
                                          const BodyPose &p1, const LimotCmd &cmd, int flags_in, int w, Emit &emit)
{
    if (jc.b0 < 0) return 0;
    const bool has_b1 = jc.b1 >= 0;
    const float hinv = 1.0f / P.h;
    const float kerp = hinv * P.erp;
    const bool in1 = (flags_in & B2P_JR_L1) != 0, in2 = (flags_in & B2P_JR_L2) != 0;
    int mrows = 0, lflags = 0;
    Row r;
    if (jc.type == 1) {
        // ---- hinge: setBall + two angular rows + limot (ODE hinge.cpp getInfo2)
        float a1[3], a2[3], cb[3] = {0.f, 0.f, 0.f};
        mul_R_v(a1, p0.R, jc.anchor1);
        if (has_b1) {
            mul_R_v(a2, p1.R, jc.anchor2);
            if (FULL) {
#pragma unroll
                for (int i = 0; i < 3; i++) cb[i] = kerp * (a2[i] + p1.pos[i] - a1[i] - p0.pos[i]);
            }
        } else {
            a2[0] = a2[1] = a2[2] = 0.f;
            if (FULL) {
#pragma unroll
                for (int i = 0; i < 3; i++) cb[i] = kerp * (jc.anchor2[i] - a1[i] - p0.pos[i]);
            }
        }
        ball_rows<FULL>(a1, a2, cb, has_b1, P.cfm, emit);
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
        emit(r);
        row_init(r, P.cfm);
        r.J[3] = qv[0]; r.J[4] = qv[1]; r.J[5] = qv[2];
        if (has_b1) { r.J[9] = -qv[0]; r.J[10] = -qv[1]; r.J[11] = -qv[2]; }
        r.c = kerp * dot3(bv, qv);
        emit(r);
        mrows = 5;
        LimotState st;
        st.limit = 0, st.limit_err = 0.f;
        if (FULL) joint_limit_state(jc, p0, p1, st);
        if (limot_row<FULL>(P, jc.limot1, cmd.vel1, cmd.fmax1, st, in1, ax1, 1, p0, p1, jc.b0, jc.b1, w, emit)) {
            lflags |= B2P_JR_L1;
            mrows++;
        }
    } else if (jc.type == 4 || jc.type == 5) {
        // ---- ball (ODE ball.cpp getInfo2: setBall with the joint's own erp / cfm) and universal
        // (universal.cpp getInfo2: setBall + one angular row along ax1 x ax2 + two limot rows)
        float a1[3], a2[3], cb[3] = {0.f, 0.f, 0.f};
        const float kb = hinv * (jc.type == 4 ? jc.erp : P.erp);
        mul_R_v(a1, p0.R, jc.anchor1);
        if (has_b1) {
            mul_R_v(a2, p1.R, jc.anchor2);
            if (FULL) {
#pragma unroll
                for (int i = 0; i < 3; i++) cb[i] = kb * (a2[i] + p1.pos[i] - a1[i] - p0.pos[i]);
            }
        } else {
            a2[0] = a2[1] = a2[2] = 0.f;
            if (FULL) {
#pragma unroll
                for (int i = 0; i < 3; i++) cb[i] = kb * (jc.anchor2[i] - a1[i] - p0.pos[i]);
            }
        }
        ball_rows<FULL>(a1, a2, cb, has_b1, jc.type == 4 ? jc.cfm : P.cfm, emit);
        mrows = 3;
        if (jc.type == 5) {
            float ax1[3], ax2[3], ax2t[3], pv[3];
            universal_axes(jc, p0, p1, has_b1, ax1, ax2);
            const float k = dot3(ax1, ax2);
#pragma unroll
            for (int i = 0; i < 3; i++) ax2t[i] = ax2[i] - k * ax1[i];
            cross3(pv, ax1, ax2t);
            normalize3(pv);
            row_init(r, P.cfm);
            r.J[3] = pv[0]; r.J[4] = pv[1]; r.J[5] = pv[2];
            if (has_b1) { r.J[9] = -pv[0]; r.J[10] = -pv[1]; r.J[11] = -pv[2]; }
            r.c = kerp * -k;
            emit(r);
            mrows = 4;
            LimotState st, st2;
            st.limit = st2.limit = 0, st.limit_err = st2.limit_err = 0.f;
            if (FULL) {
                joint_limit_state(jc, p0, p1, st);
                joint_limit_state2(jc, p0, p1, st2);
            }
            if (limot_row<FULL>(P, jc.limot1, cmd.vel1, cmd.fmax1, st, in1, ax1, 1, p0, p1, jc.b0, jc.b1, w, emit)) {
                lflags |= B2P_JR_L1;
                mrows++;
            }
            if (limot_row<FULL>(P, jc.limot2, cmd.vel2, cmd.fmax2, st2, in2, ax2, 1, p0, p1, jc.b0, jc.b1, w, emit)) {
                lflags |= B2P_JR_L2;
                mrows++;
            }
        }
    } else if (jc.type == 6) {
        // ---- fixed (ODE fixed.cpp getInfo2): rows 0-2 position, rows 3-5 orientation
        float ofs[3], cl[3] = {0.f, 0.f, 0.f}, ca[3] = {0.f, 0.f, 0.f};
        mul_R_v(ofs, p0.R, jc.offset);
        if (FULL) {
            const float k = hinv * jc.erp;
            if (has_b1) {
#pragma unroll
                for (int i = 0; i < 3; i++) cl[i] = k * (p1.pos[i] - p0.pos[i] + ofs[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 3; i++) cl[i] = k * (jc.offset[i] - p0.pos[i]);
            }
            fixed_orientation_rhs(jc, p0, p1, has_b1, hinv * P.erp * 2.f, ca);
        }
        row_init(r, P.cfm);
        r.J[0] = 1.f;
        if (has_b1) { r.J[4] = -ofs[2]; r.J[5] = ofs[1]; r.J[6] = -1.f; }
        r.c = cl[0]; r.cfm = jc.cfm;
        emit(r);
        row_init(r, P.cfm);
        r.J[1] = 1.f;
        if (has_b1) { r.J[3] = ofs[2]; r.J[5] = -ofs[0]; r.J[7] = -1.f; }
        r.c = cl[1]; r.cfm = jc.cfm;
        emit(r);
        row_init(r, P.cfm);
        r.J[2] = 1.f;
        if (has_b1) { r.J[3] = -ofs[1]; r.J[4] = ofs[0]; r.J[8] = -1.f; }
        r.c = cl[2]; r.cfm = jc.cfm;
        emit(r);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            row_init(r, P.cfm);
            r.J[3 + i] = 1.f;
            if (has_b1) r.J[9 + i] = -1.f;
            r.c = ca[i];
            emit(r);
        }
        mrows = 6;
    } else if (jc.type == 3) {
        // ---- slider (ODE slider.cpp getInfo2)
        float ca[3] = {0.f, 0.f, 0.f};
        if (FULL) fixed_orientation_rhs(jc, p0, p1, has_b1, hinv * P.erp * 2.f, ca);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            row_init(r, P.cfm);
            r.J[3 + i] = 1.f;
            if (has_b1) r.J[9 + i] = -1.f;
            r.c = ca[i];
            emit(r);
        }
        float c[3] = {0.f, 0.f, 0.f}, ax1[3], pv[3], qv[3];
        if (has_b1) {
#pragma unroll
            for (int i = 0; i < 3; i++) c[i] = p1.pos[i] - p0.pos[i];
        }
        mul_R_v(ax1, p0.R, jc.axis1);
        plane_space(ax1, pv, qv);
        float c3 = 0.f, c4 = 0.f, tp[3], tq[3];
        if (has_b1) {
            cross3(tp, c, pv);
            cross3(tq, c, qv);
            if (FULL) {
                float ofs[3];
                mul_R_v(ofs, p1.R, jc.offset);
#pragma unroll
                for (int i = 0; i < 3; i++) c[i] += ofs[i];
                c3 = kerp * dot3(pv, c);
                c4 = kerp * dot3(qv, c);
            }
        } else {
            if (FULL) {
                float ofs[3];
#pragma unroll
                for (int i = 0; i < 3; i++) ofs[i] = jc.offset[i] - p0.pos[i];
                c3 = kerp * dot3(pv, ofs);
                c4 = kerp * dot3(qv, ofs);
            }
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
        emit(r);
        row_init(r, P.cfm);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            r.J[i] = qv[i];
            if (has_b1) { r.J[3 + i] = 0.5f * tq[i]; r.J[9 + i] = 0.5f * tq[i]; r.J[6 + i] = -qv[i]; }
        }
        r.c = c4;
        emit(r);
        mrows = 5;
        LimotState st;
        st.limit = 0, st.limit_err = 0.f;
        if (FULL) joint_limit_state(jc, p0, p1, st);
        if (limot_row<FULL>(P, jc.limot1, cmd.vel1, cmd.fmax1, st, in1, ax1, 0, p0, p1, jc.b0, jc.b1, w, emit)) {
            lflags |= B2P_JR_L1;
            mrows++;
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
            emit(r);
        }
        row_init(r, P.cfm);
#pragma unroll
        for (int i = 0; i < 3; i++) { r.J[3 + i] = qv[i]; r.J[9 + i] = -qv[i]; }
        r.c = kerp * (jc.c0 * sn - jc.s0 * cs);
        emit(r);
        mrows = 4;
        LimotState st, st2;
        st.limit = st2.limit = 0, st.limit_err = st2.limit_err = 0.f;
        if (FULL) joint_limit_state(jc, p0, p1, st);
        if (limot_row<FULL>(P, jc.limot1, cmd.vel1, cmd.fmax1, st, in1, ax1, 1, p0, p1, jc.b0, jc.b1, w, emit)) {
            lflags |= B2P_JR_L1;
            mrows++;
        }
        if (limot_row<FULL>(P, jc.limot2, cmd.vel2, cmd.fmax2, st2, in2, ax2, 1, p0, p1, jc.b0, jc.b1, w, emit)) {
            lflags |= B2P_JR_L2;
            mrows++;
        }
    }
    return mrows | lflags;
}

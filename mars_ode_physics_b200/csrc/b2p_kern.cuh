// b2p_kern.cuh -- device helpers shared by the assemble / integrate kernels: pose loads and the
// joint-angle and limit-motor functions of ODE 0.16 (joint.cpp, hinge.cpp, slider.cpp, hinge2.cpp) that
// both the row assembly (getInfo1/getInfo2) and the post-step observations (dJointGet*Angle/Rate) use.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "b2p_dev.h"
#include "b2p_math.cuh"
#include "b2p_universal.h"

#define G(arr, field) (arr)[(size_t)(field) * Wp + w]

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

// universal joint: world axes and the two angles (ODE universal.cpp getAxes / getAngles)
__device__ __forceinline__ void universal_axes(const DevJoint &jc, const BodyPose &p0, const BodyPose &p1, bool has_b1,
                                               float *ax1, float *ax2)
{
    mul_R_v(ax1, p0.R, jc.axis1);
    if (has_b1) {
        mul_R_v(ax2, p1.R, jc.axis2);
    } else {
        ax2[0] = jc.axis2[0], ax2[1] = jc.axis2[1], ax2[2] = jc.axis2[2];
    }
}
__device__ __forceinline__ void universal_angles(const DevJoint &jc, const BodyPose &p0, const BodyPose &p1, bool has_b1,
                                                 float &a1, float &a2)
{
    float ax1[3], ax2[3];
    universal_axes(jc, p0, p1, has_b1, ax1, ax2);
    b2p_uni::angles<float>(ax1, ax2, p0.q, p1.q, has_b1 ? 1 : 0, jc.axis1, jc.axis2, jc.qrel, jc.qrel2, &a1, &a2);
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
    } else if (jc.type == 5) { // universal, first angle
        if ((l.lostop >= -PI || l.histop <= PI) && l.lostop <= l.histop) {
            float a1, a2;
            universal_angles(jc, p0, p1, jc.b1 >= 0, a1, a2);
            limot_test(l, a1, st);
        }
    }
}

// Limit state of limot 2 (universal joints only: ODE's hinge2 has no stop on its second axis)
__device__ __forceinline__ void joint_limit_state2(const DevJoint &jc, const BodyPose &p0, const BodyPose &p1,
                                                   LimotState &st)
{
    st.limit = 0;
    st.limit_err = 0.f;
    const DevLimot &l = jc.limot2;
    const float PI = 3.14159265358979323846f;
    if (jc.type == 5 && (l.lostop >= -PI || l.histop <= PI) && l.lostop <= l.histop) {
        float a1, a2;
        universal_angles(jc, p0, p1, jc.b1 >= 0, a1, a2);
        limot_test(l, a2, st);
    }
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

// b2p_universal.h -- the universal joint's angle measurement (ODE 0.16 universal.cpp getAngles /
// computeInitialRelativeRotations, rotation.cpp dRFrom2Axes / dRtoQ), written once for the host (double: joint
// creation and the one-world getters of the arena) and the device (float: limit tests in the assemble kernel,
// joint observations in the integrate kernel).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define B2P_UHD __host__ __device__ __forceinline__
#else
#define B2P_UHD inline
#endif

namespace b2p_uni {

B2P_UHD float usqrt(float x) { return sqrtf(x); }
B2P_UHD double usqrt(double x) { return sqrt(x); }
B2P_UHD float uatan2(float y, float x) { return atan2f(y, x); }
B2P_UHD double uatan2(double y, double x) { return atan2(y, x); }

template <typename T> B2P_UHD void qmul0(T *r, const T *b, const T *c) // b c
{
    r[0] = b[0] * c[0] - b[1] * c[1] - b[2] * c[2] - b[3] * c[3];
    r[1] = b[0] * c[1] + b[1] * c[0] + b[2] * c[3] - b[3] * c[2];
    r[2] = b[0] * c[2] + b[2] * c[0] + b[3] * c[1] - b[1] * c[3];
    r[3] = b[0] * c[3] + b[3] * c[0] + b[1] * c[2] - b[2] * c[1];
}
template <typename T> B2P_UHD void qmul1(T *r, const T *b, const T *c) // b^-1 c
{
    r[0] = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    r[1] = b[0] * c[1] - b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    r[2] = b[0] * c[2] - b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    r[3] = b[0] * c[3] - b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
}
template <typename T> B2P_UHD void qmul2(T *r, const T *b, const T *c) // b c^-1
{
    r[0] = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    r[1] = -b[0] * c[1] + b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    r[2] = -b[0] * c[2] + b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    r[3] = -b[0] * c[3] + b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
}

// dRFrom2Axes: columns a^, (b orthogonalised against a)^, a^ x b^ (row-major 3x3; untouched if degenerate)
template <typename T> B2P_UHD void r_from_2_axes(T *R, const T *a, const T *b)
{
    T ax = a[0], ay = a[1], az = a[2], bx = b[0], by = b[1], bz = b[2];
    T l = usqrt(ax * ax + ay * ay + az * az);
    if (l <= 0) return;
    l = 1 / l;
    ax *= l, ay *= l, az *= l;
    const T k = ax * bx + ay * by + az * bz;
    bx -= k * ax, by -= k * ay, bz -= k * az;
    l = usqrt(bx * bx + by * by + bz * bz);
    if (l <= 0) return;
    l = 1 / l;
    bx *= l, by *= l, bz *= l;
    R[0] = ax, R[3] = ay, R[6] = az;
    R[1] = bx, R[4] = by, R[7] = bz;
    R[2] = ay * bz - by * az, R[5] = az * bx - bz * ax, R[8] = ax * by - bx * ay;
}

// dRtoQ
template <typename T> B2P_UHD void r_to_q(const T *R, T *q)
{
    const T tr = R[0] + R[4] + R[8];
    T s;
    if (tr >= 0) {
        s = usqrt(tr + 1);
        q[0] = (T)0.5 * s;
        s = (T)0.5 / s;
        q[1] = (R[7] - R[5]) * s;
        q[2] = (R[2] - R[6]) * s;
        q[3] = (R[3] - R[1]) * s;
        return;
    }
    int c;
    if (R[4] > R[0]) c = R[8] > R[4] ? 2 : 1;
    else c = R[8] > R[0] ? 2 : 0;
    if (c == 0) {
        s = usqrt((R[0] - (R[4] + R[8])) + 1);
        q[1] = (T)0.5 * s;
        s = (T)0.5 / s;
        q[2] = (R[1] + R[3]) * s;
        q[3] = (R[6] + R[2]) * s;
        q[0] = (R[7] - R[5]) * s;
    } else if (c == 1) {
        s = usqrt((R[4] - (R[8] + R[0])) + 1);
        q[2] = (T)0.5 * s;
        s = (T)0.5 / s;
        q[3] = (R[5] + R[7]) * s;
        q[1] = (R[1] + R[3]) * s;
        q[0] = (R[2] - R[6]) * s;
    } else {
        s = usqrt((R[8] - (R[0] + R[4])) + 1);
        q[3] = (T)0.5 * s;
        s = (T)0.5 / s;
        q[1] = (R[6] + R[2]) * s;
        q[2] = (R[5] + R[7]) * s;
        q[0] = (R[3] - R[1]) * s;
    }
}

// getHingeAngleFromRelativeQuat
template <typename T> B2P_UHD T hinge_angle_from_qrel(const T *qrel, const T *axis)
{
    const T PI = (T)3.14159265358979323846;
    const T cost2 = qrel[0];
    const T sint2 = usqrt(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
    T theta = (qrel[1] * axis[0] + qrel[2] * axis[1] + qrel[3] * axis[2] >= 0) ? 2 * uatan2(sint2, cost2)
                                                                              : 2 * uatan2(sint2, -cost2);
    if (theta > PI) theta -= 2 * PI;
    return -theta;
}

// computeInitialRelativeRotations: ax1 / ax2 = the joint axes in world coordinates, q0 / q1 the bodies'
// quaternions (has_b1 = 0: body 2 is the world)
template <typename T>
B2P_UHD void initial_rotations(const T *ax1, const T *ax2, const T *q0, const T *q1, int has_b1, T *qrel1, T *qrel2)
{
    T R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, qcross[4];
    r_from_2_axes(R, ax1, ax2);
    r_to_q(R, qcross);
    qmul1(qrel1, q0, qcross);
    r_from_2_axes(R, ax2, ax1);
    r_to_q(R, qcross);
    if (has_b1) {
        qmul1(qrel2, q1, qcross);
    } else {
        for (int i = 0; i < 4; i++) qrel2[i] = qcross[i];
    }
}

// getAngles: axis1 / axis2 are the body-frame axes stored in the joint
template <typename T>
B2P_UHD void angles(const T *ax1, const T *ax2, const T *q0, const T *q1, int has_b1, const T *axis1, const T *axis2,
                    const T *qrel1, const T *qrel2, T *angle1, T *angle2)
{
    T R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, qcross[4], qq[4], qrel[4], qcross2[4];
    r_from_2_axes(R, ax1, ax2);
    r_to_q(R, qcross);
    qmul1(qq, q0, qcross);
    qmul2(qrel, qq, qrel1);
    *angle1 = hinge_angle_from_qrel(qrel, axis1);
    // the cross seen from body 2: the first frame turned by 180 degrees about the bisector of the two axes
    qrel[0] = 0;
    qrel[1] = ax1[0] + ax2[0];
    qrel[2] = ax1[1] + ax2[1];
    qrel[3] = ax1[2] + ax2[2];
    const T l = 1 / usqrt(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
    qrel[1] *= l, qrel[2] *= l, qrel[3] *= l;
    qmul0(qcross2, qrel, qcross);
    if (has_b1) {
        qmul1(qq, q1, qcross2);
        qmul2(qrel, qq, qrel2);
    } else {
        qmul2(qrel, qcross2, qrel2);
    }
    *angle2 = -hinge_angle_from_qrel(qrel, axis2);
}

} // namespace b2p_uni

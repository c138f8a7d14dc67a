// MassProps.hpp -- host-side mass properties (what the reference computes with dMass*: 
// src/objects/Box.cpp:39-40, Sphere.cpp:38-39, Capsule.cpp:40-41, Cylinder.cpp:40-41,
// Inertial.cpp:33-75, Frame.cpp:794-851).  Runs once per object at scene construction.
#pragma once
#include <cmath>
#include <cstring>

namespace mars {
namespace ode_physics {

struct MassProps {
    double mass = 0;
    double c[3] = {0, 0, 0};
    double I[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // row-major 3x3

    void setZero() { *this = MassProps(); }
    void setDiag(double m, double ixx, double iyy, double izz)
    {
        setZero();
        mass = m;
        I[0] = ixx;
        I[4] = iyy;
        I[8] = izz;
    }
    void setBoxTotal(double m, double lx, double ly, double lz)
    {
        setDiag(m, m / 12.0 * (ly * ly + lz * lz), m / 12.0 * (lx * lx + lz * lz), m / 12.0 * (lx * lx + ly * ly));
    }
    void setSphereTotal(double m, double r) { setDiag(m, 0.4 * m * r * r, 0.4 * m * r * r, 0.4 * m * r * r); }
    // direction: 1=x 2=y 3=z (the reference always passes 3)
    void setCylinderTotal(double m, int dir, double r, double len)
    {
        const double ia = m * (0.25 * r * r + len * len / 12.0), ib = 0.5 * m * r * r;
        setDiag(m, ia, ia, ia);
        I[(dir - 1) * 4] = ib;
    }
    void setCapsuleTotal(double m, int dir, double r, double len)
    {
        // density form with density 1, then rescaled to the requested total (ODE dMassSetCapsuleTotal)
        const double m1 = M_PI * r * r * len, m2 = (4.0 / 3.0) * M_PI * r * r * r;
        const double ia = m1 * (0.25 * r * r + len * len / 12.0) + m2 * (0.4 * r * r + 0.375 * r * len + 0.25 * len * len);
        const double ib = (m1 * 0.5 + m2 * 0.4) * r * r;
        setDiag(m1 + m2, ia, ia, ia);
        I[(dir - 1) * 4] = ib;
        adjust(m);
    }
    void adjust(double newmass)
    {
        const double s = newmass / mass;
        mass = newmass;
        for (double &v : I) v *= s;
    }
    void symmetrise()
    {
        I[3] = I[1];
        I[6] = I[2];
        I[7] = I[5];
    }
    void rotate(const double R[9])
    {
        double t[9], o[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) t[i * 3 + j] = I[i * 3] * R[j * 3] + I[i * 3 + 1] * R[j * 3 + 1] + I[i * 3 + 2] * R[j * 3 + 2];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) o[i * 3 + j] = R[i * 3] * t[j] + R[i * 3 + 1] * t[3 + j] + R[i * 3 + 2] * t[6 + j];
        std::memcpy(I, o, sizeof(o));
        symmetrise();
        const double cc[3] = {R[0] * c[0] + R[1] * c[1] + R[2] * c[2], R[3] * c[0] + R[4] * c[1] + R[5] * c[2],
                              R[6] * c[0] + R[7] * c[1] + R[8] * c[2]};
        std::memcpy(c, cc, sizeof(cc));
    }
    // parallel-axis shift of the reference point by (x,y,z)
    void translate(double x, double y, double z)
    {
        const double a[3] = {x + c[0], y + c[1], z + c[2]};
        const double a2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], c2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                const double ta = a[i] * a[j] - (i == j ? a2 : 0.0), tc = c[i] * c[j] - (i == j ? c2 : 0.0);
                I[i * 3 + j] += mass * (tc - ta);
            }
        symmetrise();
        c[0] += x;
        c[1] += y;
        c[2] += z;
    }
    void add(const MassProps &b)
    {
        const double d = 1.0 / (mass + b.mass);
        for (int i = 0; i < 3; i++) c[i] = (c[i] * mass + b.c[i] * b.mass) * d;
        mass += b.mass;
        for (int i = 0; i < 9; i++) I[i] += b.I[i];
    }
    bool check() const
    {
        if (!(mass > 0)) return false;
        const double d1 = I[0], d2 = I[0] * I[4] - I[1] * I[3];
        const double d3 = I[0] * (I[4] * I[8] - I[5] * I[7]) - I[1] * (I[3] * I[8] - I[5] * I[6]) + I[2] * (I[3] * I[7] - I[4] * I[6]);
        return d1 > 0 && d2 > 0 && d3 > 0;
    }
};

inline void quatToR(double w, double x, double y, double z, double R[9])
{
    const double qq1 = 2 * x * x, qq2 = 2 * y * y, qq3 = 2 * z * z;
    R[0] = 1 - qq2 - qq3;
    R[1] = 2 * (x * y - w * z);
    R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);
    R[4] = 1 - qq1 - qq3;
    R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);
    R[7] = 2 * (y * z + w * x);
    R[8] = 1 - qq1 - qq2;
}

} // namespace ode_physics
} // namespace mars

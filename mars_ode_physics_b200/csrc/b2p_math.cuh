// b2p_math.cuh -- small fp32 vector / quaternion helpers for the device path.
// Conventions follow ODE: quaternions are (w,x,y,z), matrices are row-major 3x3.
// Operation order matches oracle/orc_quickstep.c so that a -fmad=false build is comparable
// with the float build of the oracle term by term.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B2P_HD __host__ __device__ __forceinline__
#else
#define B2P_HD inline
#endif

B2P_HD float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

B2P_HD void cross3(float *r, const float *a, const float *b)
{
    const float x = a[1] * b[2] - a[2] * b[1];
    const float y = a[2] * b[0] - a[0] * b[2];
    const float z = a[0] * b[1] - a[1] * b[0];
    r[0] = x;
    r[1] = y;
    r[2] = z;
}

B2P_HD void mul_R_v(float *r, const float *R, const float *v)
{
    const float x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    const float y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    const float z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    r[0] = x;
    r[1] = y;
    r[2] = z;
}

B2P_HD void mul_Rt_v(float *r, const float *R, const float *v)
{
    const float x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    const float y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    const float z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    r[0] = x;
    r[1] = y;
    r[2] = z;
}

// C = A B
B2P_HD void mul33(float *C, const float *A, const float *B)
{
    float t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            t[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
#pragma unroll
    for (int i = 0; i < 9; i++) C[i] = t[i];
}

// C = A B^T
B2P_HD void mul33_bt(float *C, const float *A, const float *B)
{
    float t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            t[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
#pragma unroll
    for (int i = 0; i < 9; i++) C[i] = t[i];
}

// ODE dQtoR
B2P_HD void q_to_R(const float *q, float *R)
{
    const float qq1 = 2 * q[1] * q[1];
    const float qq2 = 2 * q[2] * q[2];
    const float qq3 = 2 * q[3] * q[3];
    R[0] = 1 - qq2 - qq3;
    R[1] = 2 * (q[1] * q[2] - q[0] * q[3]);
    R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
    R[3] = 2 * (q[1] * q[2] + q[0] * q[3]);
    R[4] = 1 - qq1 - qq3;
    R[5] = 2 * (q[2] * q[3] - q[0] * q[1]);
    R[6] = 2 * (q[1] * q[3] - q[0] * q[2]);
    R[7] = 2 * (q[2] * q[3] + q[0] * q[1]);
    R[8] = 1 - qq1 - qq2;
}

// ODE dQMultiply1: b^-1 * c
B2P_HD void qmul1(float *r, const float *b, const float *c)
{
    const float t0 = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    const float t1 = b[0] * c[1] - b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    const float t2 = b[0] * c[2] - b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    const float t3 = b[0] * c[3] - b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
    r[0] = t0;
    r[1] = t1;
    r[2] = t2;
    r[3] = t3;
}
// ODE dQMultiply2: b * c^-1
B2P_HD void qmul2(float *r, const float *b, const float *c)
{
    const float t0 = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    const float t1 = -b[0] * c[1] + b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    const float t2 = -b[0] * c[2] + b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    const float t3 = -b[0] * c[3] + b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
    r[0] = t0;
    r[1] = t1;
    r[2] = t2;
    r[3] = t3;
}
// ODE dQMultiply3: b^-1 * c^-1
B2P_HD void qmul3(float *r, const float *b, const float *c)
{
    const float t0 = b[0] * c[0] - b[1] * c[1] - b[2] * c[2] - b[3] * c[3];
    const float t1 = -b[0] * c[1] - b[1] * c[0] + b[2] * c[3] - b[3] * c[2];
    const float t2 = -b[0] * c[2] - b[2] * c[0] + b[3] * c[1] - b[1] * c[3];
    const float t3 = -b[0] * c[3] - b[3] * c[0] + b[1] * c[2] - b[2] * c[1];
    r[0] = t0;
    r[1] = t1;
    r[2] = t2;
    r[3] = t3;
}

B2P_HD void normalize3(float *a)
{
    float l = dot3(a, a);
    if (l > 0.f) {
        l = 1.0f / sqrtf(l);
        a[0] *= l;
        a[1] *= l;
        a[2] *= l;
    } else {
        a[0] = 1.f;
        a[1] = 0.f;
        a[2] = 0.f;
    }
}

B2P_HD void normalize4(float *a)
{
    float l = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    if (l > 0.f) {
        l = 1.0f / sqrtf(l);
        a[0] *= l;
        a[1] *= l;
        a[2] *= l;
        a[3] *= l;
    } else {
        a[0] = 1.f;
        a[1] = a[2] = a[3] = 0.f;
    }
}

#ifdef __CUDACC__
// 256-bit global accesses (sm_100: LDG.256 / STG.256).  A constraint row is two 32-byte units; a per-lane
// gather of a row then costs two 128-byte-line requests per lane instead of four.
struct F8 {
    float4 lo, hi;
};
__device__ __forceinline__ F8 ldg256(const void *p)
{
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p));
    return r;
}
// The same load without keeping the line in L1 (LDG.NA): for data that is used once -- the per-lane row gathers of
// the Tensor-Memory SOR kernel, whose 32 lines per request would otherwise churn through whatever L1 the kernel's
// shared-memory carve-out leaves.
__device__ __forceinline__ F8 ldg256_stream(const void *p)
{
    F8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256(void *p, const float4 &lo, const float4 &hi)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w),
                 "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w)
                 : "memory");
}
#endif

#ifdef __CUDACC__
// Pinned fused multiply-add for the SOR row-iteration: both SOR kernels must round identically (they are
// compared bitwise), and the B2P_STRICT build (-fmad=false, compared with the float oracle) must not fuse.
#ifdef B2P_STRICT
__device__ __forceinline__ float sor_fma(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
#else
__device__ __forceinline__ float sor_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#endif
// packed pairs (sm_100 FFMA2 / FMUL2): (a.x*b.x + c.x, a.y*b.y + c.y) in one instruction
__device__ __forceinline__ float2 sor_fma2(float2 a, float2 b, float2 c)
{
#ifdef B2P_STRICT
    return make_float2(sor_fma(a.x, b.x, c.x), sor_fma(a.y, b.y, c.y));
#else
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
        " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd; }"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
#endif
}
__device__ __forceinline__ float2 sor_mul2(float2 a, float2 b)
{
#ifdef B2P_STRICT
    return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
#else
    float2 d;
    asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mul.rn.f32x2 rd, ra, rb;\n"
        " mov.b64 {%0,%1}, rd; }"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#endif
}
#endif

// ODE dPlaneSpace
B2P_HD void plane_space(const float *n, float *p, float *q)
{
    if (fabsf(n[2]) > 0.70710678118654752440f) {
        const float a = n[1] * n[1] + n[2] * n[2];
        const float k = 1.0f / sqrtf(a);
        p[0] = 0.f;
        p[1] = -n[2] * k;
        p[2] = n[1] * k;
        q[0] = a * k;
        q[1] = -n[0] * p[2];
        q[2] = n[0] * p[1];
    } else {
        const float a = n[0] * n[0] + n[1] * n[1];
        const float k = 1.0f / sqrtf(a);
        p[0] = -n[1] * k;
        p[1] = n[0] * k;
        p[2] = 0.f;
        q[0] = -n[2] * p[1];
        q[1] = n[2] * p[0];
        q[2] = a * k;
    }
}

// ODE dInvertMatrix3; returns the determinant (0 = singular, dst untouched)
B2P_HD float invert33(float *dst, const float *m)
{
    const float c00 = m[4] * m[8] - m[5] * m[7];
    const float c01 = m[5] * m[6] - m[3] * m[8];
    const float c02 = m[3] * m[7] - m[4] * m[6];
    const float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    if (det == 0.f) return 0.f;
    const float r = 1.0f / det;
    float t[9];
    t[0] = c00 * r;
    t[1] = (m[2] * m[7] - m[1] * m[8]) * r;
    t[2] = (m[1] * m[5] - m[2] * m[4]) * r;
    t[3] = c01 * r;
    t[4] = (m[0] * m[8] - m[2] * m[6]) * r;
    t[5] = (m[2] * m[3] - m[0] * m[5]) * r;
    t[6] = c02 * r;
    t[7] = (m[1] * m[6] - m[0] * m[7]) * r;
    t[8] = (m[0] * m[4] - m[1] * m[3]) * r;
#pragma unroll
    for (int i = 0; i < 9; i++) dst[i] = t[i];
    return det;
}

// ODE dRand / dRandInt (misc.cpp): 32-bit LCG with xor-folding for small ranges
B2P_HD int rand_int(uint32_t &seed, int n)
{
    seed = 1664525u * seed + 1013904223u;
    const uint32_t un = (uint32_t)n;
    uint32_t r = seed;
    if (un <= 0x00010000u) {
        r ^= (r >> 16);
        if (un <= 0x00000100u) {
            r ^= (r >> 8);
            if (un <= 0x00000010u) {
                r ^= (r >> 4);
                if (un <= 0x00000004u) {
                    r ^= (r >> 2);
                    if (un <= 0x00000002u) r ^= (r >> 1);
                }
            }
        }
    }
    return (int)(r % un);
}

#ifdef __CUDACC__
// ODE's constraint reshuffle (quickstep.cpp ConstraintsReorderingHelper, REORDERING_METHOD__RANDOMLY) for one
// world whose rows are cut into island blocks (b2p_dev.h: block = first slot | independent rows << 8 | rows << 16).
// ODE solves the islands one after the other, each with all its iterations, so the draws of the world's LCG are
// consumed island by island: block k uses, for its reshuffle event `ev` (0-based; `nev` events per solve, at
// iterations 8, 16, ...), the draws [sum_{i<k} nev d_i + ev d_k, ... + d_k) with d = (nh - 1)+ + (m - nh - 1)+.
// The kernels sweep all blocks of a world together (they share no row and no body, so the result is the same) and
// jump through the sequence accordingly.  Returns the seed after ALL events of the step.
template <class Swap>
__device__ __forceinline__ uint32_t island_reshuffle(const uint32_t *isl, size_t Wp, int w, int nisl, uint32_t seed0,
                                                     int ev, int nev, Swap &&swap)
{
    uint32_t s = seed0;
#pragma unroll 1
    for (int k = 0; k < nisl; k++) {
        const uint32_t e = isl[(size_t)k * Wp + w];
        const int start = e & 0xff, nh = (e >> 8) & 0xff, mk = e >> 16;
        const int d = (nh > 1 ? nh - 1 : 0) + (mk - nh > 1 ? mk - nh - 1 : 0);
#pragma unroll 1
        for (int i = 0; i < ev * d; i++) s = 1664525u * s + 1013904223u;
#pragma unroll 1
        for (int idx = 1; idx < nh; idx++) swap(start + idx, start + rand_int(s, idx + 1));
#pragma unroll 1
        for (int idx = nh + 1; idx < mk; idx++) swap(start + idx, start + nh + rand_int(s, idx - nh + 1));
#pragma unroll 1
        for (int i = 0; i < (nev - 1 - ev) * d; i++) s = 1664525u * s + 1013904223u;
    }
    return s;
}
#endif

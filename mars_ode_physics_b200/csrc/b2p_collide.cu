// b2p_collide.cu -- per-world broadphase + narrowphase for sm_100a, writing the contact stream.
//
// The reference tree has no collision code (only the ContactData -> dContact mapping of
// Frame::addContact, src/objects/Frame.cpp:996-1092, and the joint-connected exclusion of
// WorldPhysics::createContact, src/WorldPhysics.cpp:458-468).  The semantics implemented here are
// the ones specified in DESIGN.md ("collision specification") and restated on the CPU in
// oracle/orc_collide.c; ODE's dContactGeom conventions are kept (unit normal from geom 2 into
// geom 1, depth >= 0).
//
// Mapping: one thread = one world, so a warp evaluates the same geom pair for 32 worlds and all
// state reads are coalesced SoA reads; per-world AABBs sit in lane-interleaved shared memory.
// The pair loop runs in (g1 < g2) order, which makes the pair list sorted and the contact stream
// order canonical (the SOR row order depends on it).  No atomics, no cross-thread compaction:
// each world appends to its own column of the stream.
#include <cuda_runtime.h>
#include <mutex>
#include <math.h>
#include <stdint.h>

#include "b2p_dev.h"
#include "b2p_math.cuh"

namespace {

constexpr int CBD = 64;
constexpr int MAXPC = 4; // contacts per pair

#define G(arr, field) (arr)[(size_t)(field) * Wp + w]

struct GPose {
    float p[3], R[9];
};

struct CPoint {
    float pos[3], n[3], depth;
};

__device__ __forceinline__ void geom_pose(const CollideParams &P, const DevGeom &g, int w, GPose &o)
{
    const int Wp = P.Wp;
    const int f = g.body * B2P_STATE_FIELDS;
    float bp[3] = {G(P.state, f + 0), G(P.state, f + 1), G(P.state, f + 2)};
    float q[4] = {G(P.state, f + 3), G(P.state, f + 4), G(P.state, f + 5), G(P.state, f + 6)};
    float bR[9];
    q_to_R(q, bR);
    if (g.has_offset) {
        float lR[9], t[3];
        q_to_R(g.lquat, lR);
        mul_R_v(t, bR, g.lpos);
#pragma unroll
        for (int i = 0; i < 3; i++) o.p[i] = bp[i] + t[i];
        mul33(o.R, bR, lR);
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) o.p[i] = bp[i];
#pragma unroll
        for (int i = 0; i < 9; i++) o.R[i] = bR[i];
    }
}

__device__ __forceinline__ void geom_aabb(const DevGeom &g, const GPose &o, float *e)
{
    switch (g.cls) {
    case 0: e[0] = e[1] = e[2] = g.dims[0]; break;
    case 1:
#pragma unroll
        for (int i = 0; i < 3; i++)
            e[i] = 0.5f * (fabsf(o.R[3 * i]) * g.dims[0] + fabsf(o.R[3 * i + 1]) * g.dims[1] +
                           fabsf(o.R[3 * i + 2]) * g.dims[2]);
        break;
    case 2:
#pragma unroll
        for (int i = 0; i < 3; i++) e[i] = fabsf(o.R[3 * i + 2]) * 0.5f * g.dims[1] + g.dims[0];
        break;
    case 3:
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float a = o.R[3 * i + 2], s = 1.f - a * a;
            e[i] = fabsf(a) * 0.5f * g.dims[1] + g.dims[0] * sqrtf(s > 0.f ? s : 0.f);
        }
        break;
    default: e[0] = e[1] = e[2] = 0.f; break;
    }
}

__device__ __forceinline__ void box_vertex(const DevGeom &g, const GPose &o, int k, float *v)
{
    const float l[3] = {(k & 1 ? -0.5f : 0.5f) * g.dims[0], (k & 2 ? -0.5f : 0.5f) * g.dims[1],
                        (k & 4 ? -0.5f : 0.5f) * g.dims[2]};
    float t[3];
    mul_R_v(t, o.R, l);
#pragma unroll
    for (int i = 0; i < 3; i++) v[i] = o.p[i] + t[i];
}

// Choose the (at most) MAXPC deepest of N candidates with depth >= 0 (earlier index wins ties) and
// emit them in candidate-index order, so exact depth ties cannot permute the contact stream.
template <int N> __device__ __forceinline__ int keep_deepest(const CPoint *cand, CPoint *out)
{
    unsigned used = 0;
#pragma unroll
    for (int r = 0; r < MAXPC; r++) {
        int best = -1;
        float bd = -1.f;
#pragma unroll
        for (int k = 0; k < N; k++) {
            if (!(used & (1u << k)) && cand[k].depth >= 0.f && (best < 0 || cand[k].depth > bd)) {
                best = k;
                bd = cand[k].depth;
            }
        }
        if (best >= 0) used |= 1u << best;
    }
    int nout = 0;
#pragma unroll
    for (int k = 0; k < N; k++)
        if (used & (1u << k)) out[nout++] = cand[k];
    return nout;
}

__device__ __forceinline__ int collide_plane(const DevGeom &g, const GPose &o, const float *pl, CPoint *out)
{
    const float *n = pl;
    const float d = pl[3];
    if (g.cls == 0) {
        const float depth = d - dot3(n, o.p) + g.dims[0];
        if (depth < 0.f) return 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            out[0].pos[i] = o.p[i] - n[i] * g.dims[0];
            out[0].n[i] = n[i];
        }
        out[0].depth = depth;
        return 1;
    }
    if (g.cls == 1) {
        CPoint cand[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            box_vertex(g, o, k, cand[k].pos);
            cand[k].depth = d - dot3(n, cand[k].pos);
#pragma unroll
            for (int i = 0; i < 3; i++) cand[k].n[i] = n[i];
        }
        return keep_deepest<8>(cand, out);
    }
    if (g.cls == 2) {
        CPoint cand[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const float sg = s == 0 ? 0.5f : -0.5f;
            float e[3];
#pragma unroll
            for (int i = 0; i < 3; i++) e[i] = o.p[i] + sg * g.dims[1] * o.R[3 * i + 2];
            cand[s].depth = d - dot3(n, e) + g.dims[0];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                cand[s].pos[i] = e[i] - n[i] * g.dims[0];
                cand[s].n[i] = n[i];
            }
        }
        return keep_deepest<2>(cand, out);
    }
    if (g.cls == 3) {
        const float a[3] = {o.R[2], o.R[5], o.R[8]};
        float u[3], v[3];
        const float na = dot3(n, a);
#pragma unroll
        for (int i = 0; i < 3; i++) u[i] = -n[i] + na * a[i];
        const float lu = sqrtf(dot3(u, u));
        if (lu > 1e-6f) {
#pragma unroll
            for (int i = 0; i < 3; i++) u[i] /= lu;
        } else {
            float q[3];
            plane_space(a, u, q);
        }
        cross3(v, a, u);
        CPoint cand[8];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const float sg = s == 0 ? 0.5f : -0.5f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float su = k == 0 ? 1.f : (k == 1 ? -1.f : 0.f), sv = k == 2 ? 1.f : (k == 3 ? -1.f : 0.f);
                CPoint &c = cand[s * 4 + k];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    c.pos[i] = o.p[i] + sg * g.dims[1] * a[i] + g.dims[0] * (su * u[i] + sv * v[i]);
                    c.n[i] = n[i];
                }
                c.depth = d - dot3(n, c.pos);
            }
        }
        return keep_deepest<8>(cand, out);
    }
    return 0;
}

__device__ __forceinline__ int collide_sphere_sphere(const float *p1, float r1, const float *p2, float r2,
                                                     CPoint &out)
{
    const float dv[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    const float d = sqrtf(dot3(dv, dv));
    if (d > r1 + r2) return 0;
    if (d <= 0.f) {
#pragma unroll
        for (int i = 0; i < 3; i++) out.pos[i] = p1[i];
        out.n[0] = 1.f;
        out.n[1] = out.n[2] = 0.f;
        out.depth = r1 + r2;
    } else {
        const float k = 0.5f * (r2 - r1 - d);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            out.n[i] = dv[i] / d;
            out.pos[i] = p1[i] + out.n[i] * k;
        }
        out.depth = r1 + r2 - d;
    }
    return 1;
}

__device__ __forceinline__ int collide_sphere_box(const float *c, float r, const DevGeom &gb, const GPose &ob,
                                                  CPoint &out)
{
    const float rel[3] = {c[0] - ob.p[0], c[1] - ob.p[1], c[2] - ob.p[2]};
    float l[3], q[3];
    mul_Rt_v(l, ob.R, rel);
    bool inside = true;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float h = 0.5f * gb.dims[i];
        q[i] = l[i];
        if (q[i] > h) {
            q[i] = h;
            inside = false;
        } else if (q[i] < -h) {
            q[i] = -h;
            inside = false;
        }
    }
    if (!inside) {
        const float dl[3] = {l[0] - q[0], l[1] - q[1], l[2] - q[2]};
        const float dist = sqrtf(dot3(dl, dl));
        const float depth = r - dist;
        if (depth < 0.f) return 0;
        float nw[3], qw[3];
        mul_R_v(nw, ob.R, dl);
        mul_R_v(qw, ob.R, q);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            out.n[i] = nw[i] / dist;
            out.pos[i] = ob.p[i] + qw[i];
        }
        out.depth = depth;
        return 1;
    }
    int ax = 0;
    float best = -1e30f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float m = fabsf(l[i]) - 0.5f * gb.dims[i];
        if (m > best) {
            best = m;
            ax = i;
        }
    }
    float sg = 1.f;
#pragma unroll
    for (int i = 0; i < 3; i++)
        if (i == ax) {
            sg = l[i] >= 0.f ? 1.f : -1.f;
            q[i] = sg * 0.5f * gb.dims[i];
        }
    float qw[3];
    mul_R_v(qw, ob.R, q);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        out.n[i] = sg * (ax == 0 ? ob.R[3 * i] : (ax == 1 ? ob.R[3 * i + 1] : ob.R[3 * i + 2]));
        out.pos[i] = ob.p[i] + qw[i];
    }
    out.depth = r - best;
    return 1;
}


// ---- box (g1 = A) vs box (g2 = B): separating-axis test over the 15 axes, then either one edge-edge
// contact or a reference-face / incident-face clip, keeping the <= 4 deepest clipped points in vertex
// order.  Mirrors collide_box_box of oracle/orc_collide.c step by step (same tie-breaking).

// ---- colliders built from point / sphere samples (semantics: oracle/orc_collide.c, same names).  All return
// contacts with the normal pointing from the second geom into the first.

// sphere (centre c, radius r; r = 0: a point) vs solid cylinder; normal from the cylinder into the sphere
__device__ __noinline__ int collide_sphere_cylinder(const float *c, float r, const DevGeom &gc, const GPose &oc, CPoint &out)
{
    const float a[3] = {oc.R[2], oc.R[5], oc.R[8]};
    const float rel[3] = {c[0] - oc.p[0], c[1] - oc.p[1], c[2] - oc.p[2]};
    const float z = dot3(rel, a), hl = 0.5f * gc.dims[1], R = gc.dims[0];
    float rv[3] = {rel[0] - z * a[0], rel[1] - z * a[1], rel[2] - z * a[2]}, er[3], tmp[3];
    const float rho = sqrtf(dot3(rv, rv));
    if (rho > 1e-9f) {
#pragma unroll
        for (int i = 0; i < 3; i++) er[i] = rv[i] / rho;
    } else {
        plane_space(a, er, tmp);
    }
    if (fabsf(z) > hl || rho > R) {
        const float zc = z > hl ? hl : (z < -hl ? -hl : z), rc = rho > R ? R : rho;
        float q[3], dv[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            q[i] = oc.p[i] + zc * a[i] + rc * er[i];
            dv[i] = c[i] - q[i];
        }
        const float dist = sqrtf(dot3(dv, dv));
        if (r - dist < 0.f || dist <= 0.f) return 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            out.n[i] = dv[i] / dist;
            out.pos[i] = q[i];
        }
        out.depth = r - dist;
        return 1;
    }
    const float dr = R - rho, dz = hl - fabsf(z);
    if (dr < dz) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            out.n[i] = er[i];
            out.pos[i] = oc.p[i] + z * a[i] + R * er[i];
        }
        out.depth = r + dr;
    } else {
        const float sg = z >= 0.f ? 1.f : -1.f;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            out.n[i] = sg * a[i];
            out.pos[i] = oc.p[i] + sg * hl * a[i] + rho * er[i];
        }
        out.depth = r + dz;
    }
    return 1;
}

// the eight rim sample points of a cylinder ('+' cap first: towards dir, away, the two sideways)
__device__ __noinline__ void cylinder_rim(const DevGeom &g, const GPose &o, const float *dir, float (*pts)[3])
{
    const float a[3] = {o.R[2], o.R[5], o.R[8]};
    float u[3], v[3];
    const float da = dot3(dir, a);
#pragma unroll
    for (int i = 0; i < 3; i++) u[i] = dir[i] - da * a[i];
    const float lu = sqrtf(dot3(u, u));
    if (lu > 1e-6f) {
#pragma unroll
        for (int i = 0; i < 3; i++) u[i] /= lu;
    } else {
        float q[3];
        plane_space(a, u, q);
    }
    cross3(v, a, u);
    for (int s = 0; s < 2; s++)
        for (int k = 0; k < 4; k++) {
            const float su = k == 0 ? 1.f : (k == 1 ? -1.f : 0.f), sv = k == 2 ? 1.f : (k == 3 ? -1.f : 0.f);
#pragma unroll
            for (int i = 0; i < 3; i++)
                pts[s * 4 + k][i] = o.p[i] + (s == 0 ? 0.5f : -0.5f) * g.dims[1] * a[i] + g.dims[0] * (su * u[i] + sv * v[i]);
        }
}

// '+' cap centre, '-' cap centre and the axis point closest to `target` (nmid = 0: it is a cap centre)
__device__ __forceinline__ void capsule_samples(const DevGeom &g, const GPose &o, const float *target, float (*pts)[3],
                                                int &nmid)
{
    const float a[3] = {o.R[2], o.R[5], o.R[8]}, hl = 0.5f * g.dims[1];
    const float rel[3] = {target[0] - o.p[0], target[1] - o.p[1], target[2] - o.p[2]};
    float t = dot3(rel, a);
    nmid = 1;
    if (t >= hl) t = hl, nmid = 0;
    if (t <= -hl) t = -hl, nmid = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        pts[0][i] = o.p[i] + hl * a[i];
        pts[1][i] = o.p[i] - hl * a[i];
        pts[2][i] = o.p[i] + t * a[i];
    }
}

__device__ __noinline__ int collide_capsule_box(const DevGeom &gc, const GPose &oc, const DevGeom &gb, const GPose &ob,
                                                CPoint *out)
{
    float pts[3][3];
    int nmid;
    CPoint cand[3];
    capsule_samples(gc, oc, ob.p, pts, nmid);
    for (int k = 0; k < 3; k++) {
        cand[k].depth = -1.f;
        if (k == 2 && !nmid) continue;
        if (!collide_sphere_box(pts[k], gc.dims[0], gb, ob, cand[k])) cand[k].depth = -1.f;
    }
    return keep_deepest<3>(cand, out);
}

__device__ __noinline__ int collide_capsule_cylinder(const DevGeom &gc, const GPose &oc, const DevGeom &gy, const GPose &oy,
                                                     CPoint *out)
{
    float pts[3][3];
    int nmid;
    CPoint cand[3];
    capsule_samples(gc, oc, oy.p, pts, nmid);
    for (int k = 0; k < 3; k++) {
        cand[k].depth = -1.f;
        if (k == 2 && !nmid) continue;
        if (!collide_sphere_cylinder(pts[k], gc.dims[0], gy, oy, cand[k])) cand[k].depth = -1.f;
    }
    return keep_deepest<3>(cand, out);
}

// direction from a point at c straight into a solid cylinder (minus the outward normal of its nearest surface)
__device__ __forceinline__ void into_cylinder(const float *c, const DevGeom &g, const GPose &o, float *dir)
{
    const float a[3] = {o.R[2], o.R[5], o.R[8]};
    const float rel[3] = {c[0] - o.p[0], c[1] - o.p[1], c[2] - o.p[2]};
    const float z = dot3(rel, a);
    const float rv[3] = {rel[0] - z * a[0], rel[1] - z * a[1], rel[2] - z * a[2]};
    const float rho = sqrtf(dot3(rv, rv));
    if (fabsf(z) - 0.5f * g.dims[1] > rho - g.dims[0] || rho <= 1e-9f) {
        const float sg = z >= 0.f ? 1.f : -1.f;
#pragma unroll
        for (int i = 0; i < 3; i++) dir[i] = -sg * a[i];
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) dir[i] = -rv[i] / rho;
    }
}

__device__ __noinline__ int collide_cylinder_box(const DevGeom &gy, const GPose &oy, const DevGeom &gb, const GPose &ob,
                                                 CPoint *out)
{
    CPoint cand[16];
    float rim[8][3], dir[3];
    {
        const float rel[3] = {oy.p[0] - ob.p[0], oy.p[1] - ob.p[1], oy.p[2] - ob.p[2]};
        int ax = 0;
        float best = -3.0e38f, lax = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float l = ob.R[i] * rel[0] + ob.R[3 + i] * rel[1] + ob.R[6 + i] * rel[2];
            const float m = fabsf(l) - 0.5f * gb.dims[i];
            if (m > best) best = m, ax = i, lax = l;
        }
        const float sg = lax >= 0.f ? 1.f : -1.f;
        for (int i = 0; i < 3; i++) dir[i] = -sg * ob.R[3 * i + ax];
    }
    cylinder_rim(gy, oy, dir, rim);
    for (int k = 0; k < 8; k++)
        if (!collide_sphere_box(rim[k], 0.f, gb, ob, cand[k])) cand[k].depth = -1.f;
    for (int k = 0; k < 8; k++) {
        float v[3];
        box_vertex(gb, ob, k, v);
        if (!collide_sphere_cylinder(v, 0.f, gy, oy, cand[8 + k])) {
            cand[8 + k].depth = -1.f;
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++) cand[8 + k].n[i] = -cand[8 + k].n[i];
        }
    }
    return keep_deepest<16>(cand, out);
}

__device__ __noinline__ int collide_cylinder_cylinder(const DevGeom &ga, const GPose &oa, const DevGeom &gb, const GPose &ob,
                                                      CPoint *out)
{
    CPoint cand[16];
    float rim[8][3], dir[3];
    into_cylinder(oa.p, gb, ob, dir);
    cylinder_rim(ga, oa, dir, rim);
    for (int k = 0; k < 8; k++)
        if (!collide_sphere_cylinder(rim[k], 0.f, gb, ob, cand[k])) cand[k].depth = -1.f;
    into_cylinder(ob.p, ga, oa, dir);
    cylinder_rim(gb, ob, dir, rim);
    for (int k = 0; k < 8; k++) {
        if (!collide_sphere_cylinder(rim[k], 0.f, ga, oa, cand[8 + k])) {
            cand[8 + k].depth = -1.f;
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++) cand[8 + k].n[i] = -cand[8 + k].n[i];
        }
    }
    return keep_deepest<16>(cand, out);
}

__device__ __forceinline__ int clip_polygon(const float (*in)[3], int n, const float *pn, float pd, float (*out)[3])
{
    int m = 0;
    for (int i = 0; i < n; i++) {
        const float *a = in[i], *b = in[(i + 1) % n];
        const float da = dot3(pn, a) - pd, db = dot3(pn, b) - pd;
        if (da <= 0.f) {
            if (m < 8) {
                out[m][0] = a[0], out[m][1] = a[1], out[m][2] = a[2];
                m++;
            }
        }
        if ((da < 0.f && db > 0.f) || (da > 0.f && db < 0.f)) {
            const float t = da / (da - db);
            if (m < 8) {
                out[m][0] = a[0] + t * (b[0] - a[0]);
                out[m][1] = a[1] + t * (b[1] - a[1]);
                out[m][2] = a[2] + t * (b[2] - a[2]);
                m++;
            }
        }
    }
    return m;
}

__device__ __noinline__ int collide_box_box(const DevGeom &ga, const GPose &oa, const DevGeom &gb, const GPose &ob,
                                            CPoint *out)
{
    const float a[3] = {0.5f * ga.dims[0], 0.5f * ga.dims[1], 0.5f * ga.dims[2]};
    const float b[3] = {0.5f * gb.dims[0], 0.5f * gb.dims[1], 0.5f * gb.dims[2]};
    const float d[3] = {ob.p[0] - oa.p[0], ob.p[1] - oa.p[1], ob.p[2] - oa.p[2]};
    float A[3][3], B[3][3], Q[3][3], pp[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        A[k][0] = oa.R[k], A[k][1] = oa.R[3 + k], A[k][2] = oa.R[6 + k];
        B[k][0] = ob.R[k], B[k][1] = ob.R[3 + k], B[k][2] = ob.R[6 + k];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        pp[i] = dot3(A[i], d);
#pragma unroll
        for (int j = 0; j < 3; j++) Q[i][j] = fabsf(dot3(A[i], B[j])) + 1e-5f;
    }
    float best = -1e30f, nrm[3] = {0.f, 0.f, 0.f};
    int code = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float s = fabsf(pp[i]) - (a[i] + b[0] * Q[i][0] + b[1] * Q[i][1] + b[2] * Q[i][2]);
        if (s > 0.f) return 0;
        if (s > best) {
            best = s;
            code = 1 + i;
            nrm[0] = A[i][0], nrm[1] = A[i][1], nrm[2] = A[i][2];
        }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float s = fabsf(dot3(B[j], d)) - (a[0] * Q[0][j] + a[1] * Q[1][j] + a[2] * Q[2][j] + b[j]);
        if (s > 0.f) return 0;
        if (s > best) {
            best = s;
            code = 4 + j;
            nrm[0] = B[j][0], nrm[1] = B[j][1], nrm[2] = B[j][2];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            float n[3];
            cross3(n, A[i], B[j]);
            const float l2 = dot3(n, n);
            if (l2 < 1e-10f) continue;
            const float l = sqrtf(l2);
            float ra = 0.f, rb = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                ra += a[k] * fabsf(dot3(A[k], n));
                rb += b[k] * fabsf(dot3(B[k], n));
            }
            const float s = (fabsf(dot3(d, n)) - (ra + rb)) / l;
            if (s > 0.f) return 0;
            if (s * 1.05f > best) {
                best = s;
                code = 7 + 3 * i + j;
                nrm[0] = n[0] / l, nrm[1] = n[1] / l, nrm[2] = n[2] / l;
            }
        }
    if (!code) return 0;
    if (dot3(nrm, d) < 0.f) nrm[0] = -nrm[0], nrm[1] = -nrm[1], nrm[2] = -nrm[2];

    if (code >= 7) {
        const int i = (code - 7) / 3, j = (code - 7) % 3;
        float pa[3] = {oa.p[0], oa.p[1], oa.p[2]}, pb[3] = {ob.p[0], ob.p[1], ob.p[2]};
        for (int k = 0; k < 3; k++) {
            if (k != i) {
                const float sg = dot3(nrm, A[k]) > 0.f ? 1.f : -1.f;
                for (int c = 0; c < 3; c++) pa[c] += sg * a[k] * A[k][c];
            }
            if (k != j) {
                const float sg = dot3(nrm, B[k]) > 0.f ? -1.f : 1.f;
                for (int c = 0; c < 3; c++) pb[c] += sg * b[k] * B[k][c];
            }
        }
        const float p[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
        const float uaub = dot3(A[i], B[j]), q1 = dot3(A[i], p), q2 = -dot3(B[j], p), dd = 1.f - uaub * uaub;
        float alpha = 0.f, beta = 0.f;
        if (dd > 1e-4f) {
            alpha = (q1 + uaub * q2) / dd;
            beta = (uaub * q1 + q2) / dd;
        }
        for (int c = 0; c < 3; c++) {
            out[0].pos[c] = 0.5f * (pa[c] + alpha * A[i][c] + pb[c] + beta * B[j][c]);
            out[0].n[c] = -nrm[c];
        }
        out[0].depth = -best;
        return 1;
    }

    const bool ref_is_a = code <= 3;
    const float(*Rr)[3] = ref_is_a ? A : B;
    const float(*Ri)[3] = ref_is_a ? B : A;
    const float *pr = ref_is_a ? oa.p : ob.p, *pi = ref_is_a ? ob.p : oa.p;
    const float *hr = ref_is_a ? a : b, *hi = ref_is_a ? b : a;
    const int ra = ref_is_a ? code - 1 : code - 4;
    const float nref[3] = {ref_is_a ? nrm[0] : -nrm[0], ref_is_a ? nrm[1] : -nrm[1], ref_is_a ? nrm[2] : -nrm[2]};
    int ia = 0;
    float bestd = -1.f;
    for (int k = 0; k < 3; k++) {
        const float v = fabsf(dot3(Ri[k], nref));
        if (v > bestd) {
            bestd = v;
            ia = k;
        }
    }
    const float isg = dot3(Ri[ia], nref) > 0.f ? -1.f : 1.f;
    const int i1 = (ia + 1) % 3, i2 = (ia + 2) % 3;
    float poly[8][3], tmp[8][3];
    for (int v = 0; v < 4; v++) {
        const float su = (v == 0 || v == 3) ? 1.f : -1.f, sv = v < 2 ? 1.f : -1.f;
        for (int c = 0; c < 3; c++)
            poly[v][c] = pi[c] + isg * hi[ia] * Ri[ia][c] + su * hi[i1] * Ri[i1][c] + sv * hi[i2] * Ri[i2][c];
    }
    int n = 4;
    const int r1 = (ra + 1) % 3, r2 = (ra + 2) % 3;
    for (int e = 0; e < 4 && n > 0; e++) {
        const int ax = e < 2 ? r1 : r2;
        const float sg = (e & 1) ? -1.f : 1.f;
        const float pn[3] = {sg * Rr[ax][0], sg * Rr[ax][1], sg * Rr[ax][2]};
        const float pd = dot3(pn, pr) + hr[ax];
        n = clip_polygon(poly, n, pn, pd, tmp);
        for (int v = 0; v < n; v++) poly[v][0] = tmp[v][0], poly[v][1] = tmp[v][1], poly[v][2] = tmp[v][2];
    }
    CPoint cand[8];
    const float face_d = dot3(nref, pr) + hr[ra];
    for (int v = 0; v < 8; v++) {
        cand[v].depth = -1.f;
        if (v < n) {
            cand[v].depth = face_d - dot3(nref, poly[v]);
            for (int c = 0; c < 3; c++) {
                cand[v].pos[c] = poly[v][c];
                cand[v].n[c] = -nrm[c];
            }
        }
    }
    return keep_deepest<8>(cand, out);
}

__device__ __forceinline__ float clamp01(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }

__device__ __forceinline__ void closest_seg_seg(const float *p1, const float *d1, const float *p2, const float *d2,
                                                float *c1, float *c2)
{
    const float r[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    const float a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
    float s, t;
    const float EPS = 1e-12f;
    if (a <= EPS && e <= EPS) {
        s = t = 0.f;
    } else if (a <= EPS) {
        s = 0.f;
        t = clamp01(f / e);
    } else {
        const float c = dot3(d1, r);
        if (e <= EPS) {
            t = 0.f;
            s = clamp01(-c / a);
        } else {
            const float b = dot3(d1, d2), denom = a * e - b * b;
            s = denom > EPS ? clamp01((b * f - c * e) / denom) : 0.f;
            t = (b * s + f) / e;
            if (t < 0.f) {
                t = 0.f;
                s = clamp01(-c / a);
            } else if (t > 1.f) {
                t = 1.f;
                s = clamp01((b - c) / a);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        c1[i] = p1[i] + s * d1[i];
        c2[i] = p2[i] + t * d2[i];
    }
}

__device__ __forceinline__ void capsule_segment(const DevGeom &g, const GPose &o, float *p, float *d)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        p[i] = o.p[i] - 0.5f * g.dims[1] * o.R[3 * i + 2];
        d[i] = g.dims[1] * o.R[3 * i + 2];
    }
}

// ---- heightfield ----
__device__ __forceinline__ float hf_h(const DevHeightfield &hf, int ix, int iy)
{
    return __ldg(hf.heights + (size_t)iy * hf.nx + ix);
}

__device__ __forceinline__ void hf_triangle(const DevHeightfield &hf, int ix, int iy, int which, float *a, float *b,
                                            float *c)
{
    const float x0 = hf.x0 + ix * hf.dx, y0 = hf.y0 + iy * hf.dx, x1 = x0 + hf.dx, y1 = y0 + hf.dx;
    a[0] = x0, a[1] = y0, a[2] = hf_h(hf, ix, iy);
    if (which == 0) {
        b[0] = x1, b[1] = y0, b[2] = hf_h(hf, ix + 1, iy);
        c[0] = x1, c[1] = y1, c[2] = hf_h(hf, ix + 1, iy + 1);
    } else {
        b[0] = x1, b[1] = y1, b[2] = hf_h(hf, ix + 1, iy + 1);
        c[0] = x0, c[1] = y1, c[2] = hf_h(hf, ix, iy + 1);
    }
}

__device__ __forceinline__ void tri_normal(const float *a, const float *b, const float *c, float *n)
{
    const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    cross3(n, ab, ac);
    const float l = sqrtf(dot3(n, n));
    n[0] /= l, n[1] /= l, n[2] /= l;
}

// closest point on triangle abc to p (Ericson, Real-Time Collision Detection 5.1.5)
__device__ __forceinline__ void closest_pt_triangle(const float *p, const float *a, const float *b, const float *c,
                                                    float *q)
{
    float ab[3], ac[3], ap[3], bp[3], cp[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        ab[i] = b[i] - a[i];
        ac[i] = c[i] - a[i];
        ap[i] = p[i] - a[i];
        bp[i] = p[i] - b[i];
        cp[i] = p[i] - c[i];
    }
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) {
        q[0] = a[0], q[1] = a[1], q[2] = a[2];
        return;
    }
    const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.f && d4 <= d3) {
        q[0] = b[0], q[1] = b[1], q[2] = b[2];
        return;
    }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float v = d1 / (d1 - d3);
#pragma unroll
        for (int i = 0; i < 3; i++) q[i] = a[i] + v * ab[i];
        return;
    }
    const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.f && d5 <= d6) {
        q[0] = c[0], q[1] = c[1], q[2] = c[2];
        return;
    }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        const float wv = d2 / (d2 - d6);
#pragma unroll
        for (int i = 0; i < 3; i++) q[i] = a[i] + wv * ac[i];
        return;
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const float wv = (d4 - d3) / ((d4 - d3) + (d5 - d6));
#pragma unroll
        for (int i = 0; i < 3; i++) q[i] = b[i] + wv * (c[i] - b[i]);
        return;
    }
    const float denom = 1.0f / (va + vb + vc), v = vb * denom, wv = vc * denom;
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = a[i] + ab[i] * v + ac[i] * wv;
}

__device__ __forceinline__ bool hf_sample(const DevHeightfield &hf, float x, float y, float &h, float *n)
{
    const float fx = (x - hf.x0) / hf.dx, fy = (y - hf.y0) / hf.dx;
    if (fx < 0.f || fy < 0.f || fx > (float)(hf.nx - 1) || fy > (float)(hf.ny - 1)) return false;
    int ix = (int)floorf(fx), iy = (int)floorf(fy);
    if (ix > hf.nx - 2) ix = hf.nx - 2;
    if (iy > hf.ny - 2) iy = hf.ny - 2;
    const float u = fx - ix, v = fy - iy;
    float a[3], b[3], c[3];
    hf_triangle(hf, ix, iy, (v > u) ? 1 : 0, a, b, c);
    tri_normal(a, b, c, n);
    h = a[2] - (n[0] * (x - a[0]) + n[1] * (y - a[1])) / n[2];
    return true;
}

__device__ __forceinline__ int collide_sphere_hf(const DevHeightfield &hf, const float *c, float r, CPoint &out)
{
    int ix0 = (int)floorf((c[0] - r - hf.x0) / hf.dx), ix1 = (int)floorf((c[0] + r - hf.x0) / hf.dx);
    int iy0 = (int)floorf((c[1] - r - hf.y0) / hf.dx), iy1 = (int)floorf((c[1] + r - hf.y0) / hf.dx);
    if (ix0 < 0) ix0 = 0;
    if (iy0 < 0) iy0 = 0;
    if (ix1 > hf.nx - 2) ix1 = hf.nx - 2;
    if (iy1 > hf.ny - 2) iy1 = hf.ny - 2;
    float best = -1.f;
    {
        float h, n[3];
        if (hf_sample(hf, c[0], c[1], h, n) && c[2] < h) {
            best = r + (h - c[2]) * n[2];
            out.depth = best;
            out.n[0] = n[0], out.n[1] = n[1], out.n[2] = n[2];
            out.pos[0] = c[0], out.pos[1] = c[1], out.pos[2] = h;
        }
    }
    const float r2cull = r * r * 1.0001f + 1e-12f;
    for (int iy = iy0; iy <= iy1; iy++)
        for (int ix = ix0; ix <= ix1; ix++) {
            // the cell's four corners once for both triangles, and a conservative cull: every point of the two
            // triangles lies in the box [x0,x1] x [y0,y1] x [hmin,hmax]; if that box is farther from the centre
            // than r (with a margin far above fp32 rounding) neither triangle can touch the sphere.  The cull
            // never changes the result, it only skips triangles whose test below would fail.
            const float x0 = hf.x0 + ix * hf.dx, y0 = hf.y0 + iy * hf.dx, x1 = x0 + hf.dx, y1 = y0 + hf.dx;
            const float h00 = hf_h(hf, ix, iy), h10 = hf_h(hf, ix + 1, iy), h11 = hf_h(hf, ix + 1, iy + 1),
                        h01 = hf_h(hf, ix, iy + 1);
            const float hmn = fminf(fminf(h00, h10), fminf(h11, h01)), hmx = fmaxf(fmaxf(h00, h10), fmaxf(h11, h01));
            const float ex = fmaxf(fmaxf(x0 - c[0], c[0] - x1), 0.f), ey = fmaxf(fmaxf(y0 - c[1], c[1] - y1), 0.f);
            const float ez = fmaxf(fmaxf(hmn - c[2], c[2] - hmx), 0.f);
            if (ex * ex + ey * ey + ez * ez > r2cull) continue;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                float a[3], b[3], cc[3], q[3], n[3];
                a[0] = x0, a[1] = y0, a[2] = h00;
                if (t == 0) {
                    b[0] = x1, b[1] = y0, b[2] = h10;
                    cc[0] = x1, cc[1] = y1, cc[2] = h11;
                } else {
                    b[0] = x1, b[1] = y1, b[2] = h11;
                    cc[0] = x0, cc[1] = y1, cc[2] = h01;
                }
                closest_pt_triangle(c, a, b, cc, q);
                const float dv[3] = {c[0] - q[0], c[1] - q[1], c[2] - q[2]};
                const float dist = sqrtf(dot3(dv, dv));
                if (dist > r) continue;
                tri_normal(a, b, cc, n);
                if (dot3(dv, n) < 0.f) continue;
                const float depth = r - dist;
                if (depth > best) {
                    best = depth;
                    out.depth = depth;
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        out.pos[i] = q[i];
                        out.n[i] = dist > 1e-9f ? dv[i] / dist : n[i];
                    }
                }
            }
        }
    return best >= 0.f ? 1 : 0;
}

__device__ __forceinline__ int collide_box_hf(const DevHeightfield &hf, const DevGeom &g, const GPose &o, CPoint *out)
{
    CPoint cand[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float h, n[3];
        box_vertex(g, o, k, cand[k].pos);
        cand[k].depth = -1.f;
        if (hf_sample(hf, cand[k].pos[0], cand[k].pos[1], h, n) && cand[k].pos[2] <= h) {
            cand[k].depth = (h - cand[k].pos[2]) * n[2];
            cand[k].n[0] = n[0], cand[k].n[1] = n[1], cand[k].n[2] = n[2];
        }
    }
    return keep_deepest<8>(cand, out);
}


__device__ __noinline__ int collide_capsule_hf(const DevHeightfield &hf, const DevGeom &g, const GPose &o, CPoint *out)
{
    float pts[3][3];
    int nmid;
    CPoint cand[3];
    capsule_samples(g, o, o.p, pts, nmid);
    for (int k = 0; k < 3; k++)
        if (!collide_sphere_hf(hf, pts[k], g.dims[0], cand[k])) cand[k].depth = -1.f;
    return keep_deepest<3>(cand, out);
}

__device__ __noinline__ int collide_cylinder_hf(const DevHeightfield &hf, const DevGeom &g, const GPose &o, CPoint *out)
{
    CPoint cand[8];
    float rim[8][3];
    const float down[3] = {0.f, 0.f, -1.f};
    cylinder_rim(g, o, down, rim);
    for (int k = 0; k < 8; k++) {
        float h, n[3];
        cand[k].pos[0] = rim[k][0], cand[k].pos[1] = rim[k][1], cand[k].pos[2] = rim[k][2];
        cand[k].depth = -1.f;
        if (hf_sample(hf, rim[k][0], rim[k][1], h, n) && rim[k][2] <= h) {
            cand[k].depth = (h - rim[k][2]) * n[2];
            cand[k].n[0] = n[0], cand[k].n[1] = n[1], cand[k].n[2] = n[2];
        }
    }
    return keep_deepest<8>(cand, out);
}

__device__ __forceinline__ void flip(CPoint *c, int n)
{
    for (int k = 0; k < n; k++) {
        c[k].n[0] = -c[k].n[0];
        c[k].n[1] = -c[k].n[1];
        c[k].n[2] = -c[k].n[2];
    }
}

// g1 is dynamic; g2 dynamic or static.  Contacts have the normal pointing from g2 into g1.
template <bool SAMPLED>
__device__ __forceinline__ int narrowphase(const CollideParams &P, const DevHeightfield &hfield, const DevGeom &g1,
                                           const DevGeom &g2, int w, CPoint *out)
{
    GPose o1, o2;
    geom_pose(P, g1, w, o1);
    const int c1 = g1.cls, c2 = g2.cls;
    if (c2 == 4) return collide_plane(g1, o1, g2.dims, out);
    if (c2 == 5) {
        if (c1 == 0) return collide_sphere_hf(hfield, o1.p, g1.dims[0], out[0]);
        if (c1 == 1) return collide_box_hf(hfield, g1, o1, out);
        if (SAMPLED && c1 == 2) return collide_capsule_hf(hfield, g1, o1, out);
        if (SAMPLED && c1 == 3) return collide_cylinder_hf(hfield, g1, o1, out);
        return 0;
    }
    geom_pose(P, g2, w, o2);
    if (c1 == 1 && c2 == 1) return collide_box_box(g1, o1, g2, o2, out);
    if (c1 == 0 && c2 == 0) return collide_sphere_sphere(o1.p, g1.dims[0], o2.p, g2.dims[0], out[0]);
    if (c1 == 0 && c2 == 1) return collide_sphere_box(o1.p, g1.dims[0], g2, o2, out[0]);
    if (c1 == 1 && c2 == 0) {
        const int n = collide_sphere_box(o2.p, g2.dims[0], g1, o1, out[0]);
        flip(out, n);
        return n;
    }
    if (c1 == 2 && c2 == 0) {
        float p[3], d[3], zero[3] = {0.f, 0.f, 0.f}, q1[3], q2[3];
        capsule_segment(g1, o1, p, d);
        closest_seg_seg(p, d, o2.p, zero, q1, q2);
        return collide_sphere_sphere(q1, g1.dims[0], o2.p, g2.dims[0], out[0]);
    }
    if (c1 == 0 && c2 == 2) {
        float p[3], d[3], zero[3] = {0.f, 0.f, 0.f}, q1[3], q2[3];
        capsule_segment(g2, o2, p, d);
        closest_seg_seg(p, d, o1.p, zero, q1, q2);
        return collide_sphere_sphere(o1.p, g1.dims[0], q1, g2.dims[0], out[0]);
    }
    if (c1 == 2 && c2 == 2) {
        float p1[3], d1[3], p2[3], d2[3], q1[3], q2[3];
        capsule_segment(g1, o1, p1, d1);
        capsule_segment(g2, o2, p2, d2);
        closest_seg_seg(p1, d1, p2, d2, q1, q2);
        return collide_sphere_sphere(q1, g1.dims[0], q2, g2.dims[0], out[0]);
    }
    if (!SAMPLED) return 0;
    int n = 0;
    if (c1 == 2 && c2 == 1) return collide_capsule_box(g1, o1, g2, o2, out);
    if (c1 == 1 && c2 == 2) {
        n = collide_capsule_box(g2, o2, g1, o1, out);
        flip(out, n);
        return n;
    }
    if (c1 == 0 && c2 == 3) return collide_sphere_cylinder(o1.p, g1.dims[0], g2, o2, out[0]);
    if (c1 == 3 && c2 == 0) {
        n = collide_sphere_cylinder(o2.p, g2.dims[0], g1, o1, out[0]);
        flip(out, n);
        return n;
    }
    if (c1 == 2 && c2 == 3) return collide_capsule_cylinder(g1, o1, g2, o2, out);
    if (c1 == 3 && c2 == 2) {
        n = collide_capsule_cylinder(g2, o2, g1, o1, out);
        flip(out, n);
        return n;
    }
    if (c1 == 3 && c2 == 1) return collide_cylinder_box(g1, o1, g2, o2, out);
    if (c1 == 1 && c2 == 3) {
        n = collide_cylinder_box(g2, o2, g1, o1, out);
        flip(out, n);
        return n;
    }
    if (c1 == 3 && c2 == 3) return collide_cylinder_cylinder(g1, o1, g2, o2, out);
    return 0; // planes / heightfields against each other
}

// SAMPLED: the scene has geom pairs handled by the sample-point colliders (cylinders; capsule against box or
// heightfield).  Scenes without them run the variant that does not carry that code's registers and stack.
template <bool SAMPLED> __global__ void __launch_bounds__(CBD, SAMPLED ? 6 : 8) b2p_collide_kernel(const CollideParams P)
{
    extern __shared__ float saabb[]; // [ng*6][CBD]: centre, half extents
    const DevTemplate *__restrict__ T = P.T;
    const int tid = threadIdx.x;
    const int w = blockIdx.x * CBD + tid;
    const int Wp = P.Wp;
    const int ng = T->ng, maxc = T->maxc;
    // the scene template's geoms, body adjacency and heightfield header in shared memory (read warp-uniformly
    // in every pair test)
    const DevGeom *s_geoms;
    const uint32_t *s_connected;
    const DevHeightfield *s_hf;
    {
        uint32_t *sg = (uint32_t *)(saabb + (size_t)ng * 6 * CBD);
        uint32_t *sc = sg + (size_t)ng * (sizeof(DevGeom) / 4);
        uint32_t *sh = sc + B2P_MAX_BODIES;
        sh += ((uintptr_t)sh & 7) ? 1 : 0; // the header holds a pointer
        const uint32_t *gg = (const uint32_t *)T->geoms, *gc = (const uint32_t *)T->connected, *gh = (const uint32_t *)&T->hf;
        for (int i = tid; i < ng * (int)(sizeof(DevGeom) / 4); i += CBD) sg[i] = __ldg(gg + i);
        for (int i = tid; i < B2P_MAX_BODIES; i += CBD) sc[i] = __ldg(gc + i);
        for (int i = tid; i < (int)(sizeof(DevHeightfield) / 4); i += CBD) sh[i] = __ldg(gh + i);
        s_geoms = (const DevGeom *)sg;
        s_connected = sc;
        s_hf = (const DevHeightfield *)sh;
        __syncthreads();
    }
    if (w >= P.W) return;
    for (int g = 0; g < ng; g++) {
        const DevGeom &gg = s_geoms[g];
        if (gg.body < 0) continue;
        GPose o;
        geom_pose(P, gg, w, o);
        float e[3];
        geom_aabb(gg, o, e);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            saabb[(size_t)(g * 6 + k) * CBD + tid] = o.p[k];
            saabb[(size_t)(g * 6 + 3 + k) * CBD + tid] = e[k];
        }
    }
    int total = 0, npairs = 0, dropped = 0;
    for (int i = 0; i < ng; i++) {
        const DevGeom &gi = s_geoms[i];
        for (int j = i + 1; j < ng; j++) {
            const DevGeom &gj = s_geoms[j];
            if (gi.body < 0 && gj.body < 0) continue;
            if (gi.body == gj.body) continue;
            if (gi.body >= 0 && gj.body >= 0 && (s_connected[gi.body] >> gj.body & 1u)) continue;
            bool overlap;
            const bool i_dyn = gi.body >= 0;
            const int d_idx = i_dyn ? i : j;
            const DevGeom &gs = i_dyn ? gj : gi;
            float c[3], e[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                c[k] = saabb[(size_t)(d_idx * 6 + k) * CBD + tid];
                e[k] = saabb[(size_t)(d_idx * 6 + 3 + k) * CBD + tid];
            }
            if (gi.body >= 0 && gj.body >= 0) {
                overlap = true;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float cj = saabb[(size_t)(j * 6 + k) * CBD + tid], ej = saabb[(size_t)(j * 6 + 3 + k) * CBD + tid];
                    if (c[k] - e[k] > cj + ej || cj - ej > c[k] + e[k]) overlap = false;
                }
            } else if (gs.cls == 4) {
                const float *n = gs.dims;
                overlap = (dot3(n, c) - (fabsf(n[0]) * e[0] + fabsf(n[1]) * e[1] + fabsf(n[2]) * e[2])) <= n[3];
            } else if (gs.cls == 5) {
                const DevHeightfield &hf = *s_hf;
                const float xe = hf.x0 + (hf.nx - 1) * hf.dx, ye = hf.y0 + (hf.ny - 1) * hf.dx;
                overlap = (c[2] - e[2] <= hf.hmax) && c[0] + e[0] >= hf.x0 && c[0] - e[0] <= xe &&
                          c[1] + e[1] >= hf.y0 && c[1] - e[1] <= ye;
            } else {
                overlap = false;
            }
            if (!overlap) continue;
            if (P.pairs && npairs < P.maxpairs) G(P.pairs, npairs) = i | (j << 16);
            npairs++;
            CPoint pts[MAXPC];
            const DevGeom &g1 = i_dyn ? gi : gj;
            const DevGeom &g2 = i_dyn ? gj : gi;
            const int n = narrowphase<SAMPLED>(P, *s_hf, g1, g2, w, pts);
            const float mu = fminf(g1.mu, g2.mu), mu2 = fminf(g1.mu2, g2.mu2);
            const float rho = fminf(g1.rho, g2.rho), rho2 = fminf(g1.rho2, g2.rho2), rhoN = fminf(g1.rhoN, g2.rhoN);
            const float serp = fminf(g1.soft_erp, g2.soft_erp), scfm = fmaxf(g1.soft_cfm, g2.soft_cfm);
            int mode = 0x008 | 0x010 | 0x7000;
            if (rho > 1e-10f) mode |= 0x400;
            if (mu != mu2) mode |= 0x001;
            const int ids = (g1.body & 0xff) | ((g2.body < 0 ? 0xff : g2.body) << 8);
            for (int k = 0; k < n; k++) {
                if (total >= maxc) {
                    dropped += n - k;
                    break;
                }
                const int f = total * B2P_CONTACT_FIELDS;
                G(P.crec, f + CR_IDS) = __int_as_float(ids);
                G(P.crec, f + CR_PX) = pts[k].pos[0];
                G(P.crec, f + CR_PY) = pts[k].pos[1];
                G(P.crec, f + CR_PZ) = pts[k].pos[2];
                G(P.crec, f + CR_NX) = pts[k].n[0];
                G(P.crec, f + CR_NY) = pts[k].n[1];
                G(P.crec, f + CR_NZ) = pts[k].n[2];
                G(P.crec, f + CR_DEPTH) = pts[k].depth;
                G(P.crec, f + CR_MU) = mu;
                G(P.crec, f + CR_MU2) = mu2;
                G(P.crec, f + CR_SOFT_ERP) = serp;
                G(P.crec, f + CR_SOFT_CFM) = scfm;
                G(P.crec, f + CR_RHO) = rho;
                G(P.crec, f + CR_RHO2) = rho2;
                G(P.crec, f + CR_RHON) = rhoN;
                G(P.crec, f + CR_MODE) = __int_as_float(mode);
                total++;
            }
        }
    }
    G(P.ccount, 0) = total;
    if (P.npairs) G(P.npairs, 0) = npairs;
    // the stream holds at most maxc contacts per world; what did not fit is counted so that the host can tell
    // (b2p_contacts_dropped) instead of seeing a full but plausible stream
    if (dropped && P.dropped) atomicAdd(P.dropped, dropped);
}

} // namespace

extern "C" cudaError_t b2p_launch_collide(const CollideParams *P, int ng, int sampled, cudaStream_t stream)
{
    static size_t configured_dev[B2P_MAX_DEVICES][2] = {{0, 0}}; // cudaFuncSetAttribute is per device
    static std::mutex configured_mutex; // arenas of several threads / devices may launch at the same time
    std::lock_guard<std::mutex> configured_lock(configured_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    size_t *configured = configured_dev[dev % B2P_MAX_DEVICES];
    const size_t smem = sizeof(float) * (size_t)(ng > 0 ? ng : 1) * 6 * CBD + (size_t)ng * sizeof(DevGeom) +
                        sizeof(uint32_t) * B2P_MAX_BODIES + sizeof(DevHeightfield) + 16;
    auto kern = sampled ? b2p_collide_kernel<true> : b2p_collide_kernel<false>;
    if (smem > configured[sampled ? 1 : 0]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[sampled ? 1 : 0] = smem;
    }
    const int grid = (P->W + CBD - 1) / CBD;
    kern<<<grid, CBD, smem, stream>>>(*P);
    return cudaGetLastError();
}

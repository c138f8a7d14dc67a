/*
 * oracle/orc_collide.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Broadphase + narrowphase contact generation.  The reference tree contains NO collision
 * code (only commented-out lines at src/objects/Frame.cpp:149, src/objects/Object.cpp:357-358);
 * the only in-tree contract is the ContactData -> dContact mapping of Frame::addContact
 * (src/objects/Frame.cpp:996-1092) and WorldPhysics::createContact's joint-connected exclusion
 * (src/WorldPhysics.cpp:458-468).  This file therefore *defines* the collision semantics the
 * device kernels are checked against (DESIGN.md "collision specification"), honouring ODE's
 * dContactGeom conventions: normal is unit length and points from geom 2 into geom 1, depth >= 0.
 * All geometry is evaluated in double precision.
 *
 * PARITY UNPINNED (see orc.h).
 */
#include <stdlib.h>
#include <string.h>

#include "orc_internal.h"

#define MAX_PAIR_CONTACTS 4

typedef struct {
    double pos[3], normal[3], depth;
} cpoint;

typedef struct {
    double p[3], R[9];
} gpose;

static double ddot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void dcross(double *r, const double *a, const double *b)
{
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x;
    r[1] = y;
    r[2] = z;
}

static void dq_to_R(const double *q, double *R)
{
    double qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
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

/* world pose of a geom: R = R_body R_local, p = p_body + R_body lpos */
static void geom_pose(const orc_world *w, const orc_geom *g, gpose *o)
{
    double bp[3], bR[9], lR[9];
    orc_body_get_position(w, g->body, bp);
    orc_body_get_rotation(w, g->body, bR);
    dq_to_R(g->lquat, lR);
    for (int i = 0; i < 3; i++) {
        o->p[i] = bp[i] + bR[3 * i] * g->lpos[0] + bR[3 * i + 1] * g->lpos[1] + bR[3 * i + 2] * g->lpos[2];
        for (int j = 0; j < 3; j++)
            o->R[3 * i + j] = bR[3 * i] * lR[j] + bR[3 * i + 1] * lR[3 + j] + bR[3 * i + 2] * lR[6 + j];
    }
}

/* AABB of a dynamic geom: centre c, half extents e */
static void geom_aabb(const orc_geom *g, const gpose *o, double *c, double *e)
{
    memcpy(c, o->p, 3 * sizeof(double));
    switch (g->cls) {
    case 0: e[0] = e[1] = e[2] = g->dims[0]; break;
    case 1:
        for (int i = 0; i < 3; i++)
            e[i] = 0.5 * (fabs(o->R[3 * i]) * g->dims[0] + fabs(o->R[3 * i + 1]) * g->dims[1] + fabs(o->R[3 * i + 2]) * g->dims[2]);
        break;
    case 2:
        for (int i = 0; i < 3; i++) e[i] = fabs(o->R[3 * i + 2]) * 0.5 * g->dims[1] + g->dims[0];
        break;
    case 3:
        for (int i = 0; i < 3; i++) {
            double a = o->R[3 * i + 2], s = 1 - a * a;
            e[i] = fabs(a) * 0.5 * g->dims[1] + g->dims[0] * sqrt(s > 0 ? s : 0);
        }
        break;
    default: e[0] = e[1] = e[2] = 0; break;
    }
}

/* choose the (at most) MAX_PAIR_CONTACTS deepest candidates with depth >= 0 (earlier index wins
 * ties), then emit them in candidate-index order so that exact depth ties (a box resting flat on
 * a plane) cannot permute the contact stream */
static void keep_deepest(cpoint *cand, int n, cpoint *out, int *nout)
{
    int used[16] = {0}, cnt = 0;
    while (cnt < MAX_PAIR_CONTACTS) {
        int best = -1;
        for (int k = 0; k < n; k++)
            if (!used[k] && cand[k].depth >= 0 && (best < 0 || cand[k].depth > cand[best].depth)) best = k;
        if (best < 0) break;
        used[best] = 1;
        cnt++;
    }
    *nout = 0;
    for (int k = 0; k < n; k++)
        if (used[k]) out[(*nout)++] = cand[k];
}

static void box_vertex(const orc_geom *g, const gpose *o, int k, double *v)
{
    double l[3] = {(k & 1 ? -0.5 : 0.5) * g->dims[0], (k & 2 ? -0.5 : 0.5) * g->dims[1], (k & 4 ? -0.5 : 0.5) * g->dims[2]};
    for (int i = 0; i < 3; i++) v[i] = o->p[i] + o->R[3 * i] * l[0] + o->R[3 * i + 1] * l[1] + o->R[3 * i + 2] * l[2];
}

/* ---- primitive vs plane (plane n.x = d, n unit) ---- */
static int collide_plane(const orc_geom *g, const gpose *o, const double *pl, cpoint *out)
{
    const double *n = pl;
    const double d = pl[3];
    cpoint cand[8];
    int nc = 0, nout = 0;
    switch (g->cls) {
    case 0: {
        double depth = d - ddot(n, o->p) + g->dims[0];
        if (depth < 0) return 0;
        for (int i = 0; i < 3; i++) {
            out[0].pos[i] = o->p[i] - n[i] * g->dims[0];
            out[0].normal[i] = n[i];
        }
        out[0].depth = depth;
        return 1;
    }
    case 1:
        for (int k = 0; k < 8; k++) {
            box_vertex(g, o, k, cand[nc].pos);
            cand[nc].depth = d - ddot(n, cand[nc].pos);
            memcpy(cand[nc].normal, n, 3 * sizeof(double));
            nc++;
        }
        break;
    case 2:
        for (int s = 0; s < 2; s++) { /* '+' end first */
            double sg = s == 0 ? 0.5 : -0.5, e[3];
            for (int i = 0; i < 3; i++) e[i] = o->p[i] + sg * g->dims[1] * o->R[3 * i + 2];
            cand[nc].depth = d - ddot(n, e) + g->dims[0];
            for (int i = 0; i < 3; i++) {
                cand[nc].pos[i] = e[i] - n[i] * g->dims[0];
                cand[nc].normal[i] = n[i];
            }
            nc++;
        }
        break;
    case 3: {
        double a[3] = {o->R[2], o->R[5], o->R[8]}, u[3], v[3];
        double na = ddot(n, a);
        for (int i = 0; i < 3; i++) u[i] = -n[i] + na * a[i];
        double lu = sqrt(ddot(u, u));
        if (lu > 1e-6) {
            for (int i = 0; i < 3; i++) u[i] /= lu;
        } else { /* axis parallel to the normal: any direction in the cap plane */
            double nn[3] = {a[0], a[1], a[2]}, q[3];
            orc_plane_space(nn, u, q);
        }
        dcross(v, a, u);
        for (int s = 0; s < 2; s++) {
            double sg = s == 0 ? 0.5 : -0.5;
            const double sgn[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
            for (int k = 0; k < 4; k++) {
                for (int i = 0; i < 3; i++)
                    cand[nc].pos[i] = o->p[i] + sg * g->dims[1] * a[i] + g->dims[0] * (sgn[k][0] * u[i] + sgn[k][1] * v[i]);
                cand[nc].depth = d - ddot(n, cand[nc].pos);
                memcpy(cand[nc].normal, n, 3 * sizeof(double));
                nc++;
            }
        }
        break;
    }
    default: return 0;
    }
    keep_deepest(cand, nc, out, &nout);
    return nout;
}

/* ---- sphere vs sphere (ODE dCollideSpheres) ---- */
static int collide_sphere_sphere(const double *p1, double r1, const double *p2, double r2, cpoint *out)
{
    double dv[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double d = sqrt(ddot(dv, dv));
    if (d > r1 + r2) return 0;
    if (d <= 0) {
        memcpy(out->pos, p1, 3 * sizeof(double));
        out->normal[0] = 1;
        out->normal[1] = out->normal[2] = 0;
        out->depth = r1 + r2;
    } else {
        double k = 0.5 * (r2 - r1 - d);
        for (int i = 0; i < 3; i++) {
            out->normal[i] = dv[i] / d;
            out->pos[i] = p1[i] + out->normal[i] * k;
        }
        out->depth = r1 + r2 - d;
    }
    return 1;
}

/* ---- sphere (g1) vs box (g2): normal from box into sphere ---- */
static int collide_sphere_box(const double *c, double r, const orc_geom *gb, const gpose *ob, cpoint *out)
{
    double rel[3] = {c[0] - ob->p[0], c[1] - ob->p[1], c[2] - ob->p[2]}, l[3], q[3];
    int inside = 1;
    for (int i = 0; i < 3; i++) {
        l[i] = ob->R[i] * rel[0] + ob->R[3 + i] * rel[1] + ob->R[6 + i] * rel[2]; /* R^T rel */
        double h = 0.5 * gb->dims[i];
        q[i] = l[i];
        if (q[i] > h) {
            q[i] = h;
            inside = 0;
        } else if (q[i] < -h) {
            q[i] = -h;
            inside = 0;
        }
    }
    if (!inside) {
        double dl[3] = {l[0] - q[0], l[1] - q[1], l[2] - q[2]};
        double dist = sqrt(ddot(dl, dl));
        double depth = r - dist;
        if (depth < 0) return 0;
        for (int i = 0; i < 3; i++) {
            double nl = ob->R[3 * i] * dl[0] + ob->R[3 * i + 1] * dl[1] + ob->R[3 * i + 2] * dl[2];
            out->normal[i] = nl / dist;
            out->pos[i] = ob->p[i] + ob->R[3 * i] * q[0] + ob->R[3 * i + 1] * q[1] + ob->R[3 * i + 2] * q[2];
        }
        out->depth = depth;
        return 1;
    }
    /* centre inside the box: push out through the nearest face */
    int ax = 0;
    double best = -1e300;
    for (int i = 0; i < 3; i++) {
        double m = fabs(l[i]) - 0.5 * gb->dims[i];
        if (m > best) {
            best = m;
            ax = i;
        }
    }
    double sg = l[ax] >= 0 ? 1.0 : -1.0;
    q[ax] = sg * 0.5 * gb->dims[ax];
    for (int i = 0; i < 3; i++) {
        out->normal[i] = sg * ob->R[3 * i + ax];
        out->pos[i] = ob->p[i] + ob->R[3 * i] * q[0] + ob->R[3 * i + 1] * q[1] + ob->R[3 * i + 2] * q[2];
    }
    out->depth = r - best;
    return 1;
}



/* ---- colliders built from point / sphere samples (this repository's own definitions: the reference has no
 * narrowphase).  All of them return contacts with the normal pointing from the
 * second geom into the first. ---- */

/* sphere (centre c, radius r) vs solid cylinder: closest point on the cylinder, or -- centre inside --
 * push-out through the nearer of side and cap (cap on a tie).  r = 0 tests a point.  Normal from the
 * cylinder into the sphere. */
static int collide_sphere_cylinder(const double *c, double r, const orc_geom *gc, const gpose *oc, cpoint *out)
{
    const double a[3] = {oc->R[2], oc->R[5], oc->R[8]};
    const double rel[3] = {c[0] - oc->p[0], c[1] - oc->p[1], c[2] - oc->p[2]};
    const double z = ddot(rel, a), hl = 0.5 * gc->dims[1], R = gc->dims[0];
    double rv[3] = {rel[0] - z * a[0], rel[1] - z * a[1], rel[2] - z * a[2]}, er[3], tmp[3];
    const double rho = sqrt(ddot(rv, rv));
    if (rho > 1e-9) {
        for (int i = 0; i < 3; i++) er[i] = rv[i] / rho;
    } else {
        double aa[3] = {a[0], a[1], a[2]};
        orc_plane_space(aa, er, tmp);
    }
    if (fabs(z) > hl || rho > R) {
        const double zc = z > hl ? hl : (z < -hl ? -hl : z), rc = rho > R ? R : rho;
        double q[3], dv[3];
        for (int i = 0; i < 3; i++) {
            q[i] = oc->p[i] + zc * a[i] + rc * er[i];
            dv[i] = c[i] - q[i];
        }
        const double dist = sqrt(ddot(dv, dv));
        if (r - dist < 0 || dist <= 0) return 0;
        for (int i = 0; i < 3; i++) {
            out->normal[i] = dv[i] / dist;
            out->pos[i] = q[i];
        }
        out->depth = r - dist;
        return 1;
    }
    const double dr = R - rho, dz = hl - fabs(z);
    if (dr < dz) {
        for (int i = 0; i < 3; i++) {
            out->normal[i] = er[i];
            out->pos[i] = oc->p[i] + z * a[i] + R * er[i];
        }
        out->depth = r + dr;
    } else {
        const double sg = z >= 0 ? 1.0 : -1.0;
        for (int i = 0; i < 3; i++) {
            out->normal[i] = sg * a[i];
            out->pos[i] = oc->p[i] + sg * hl * a[i] + rho * er[i];
        }
        out->depth = r + dz;
    }
    return 1;
}

/* the eight rim sample points of a cylinder: on each cap ('+' first) the points towards `dir` (projected into
 * the cap plane), away from it, and the two sideways */
static void cylinder_rim(const orc_geom *g, const gpose *o, const double *dir, double (*pts)[3])
{
    const double a[3] = {o->R[2], o->R[5], o->R[8]};
    double u[3], v[3];
    const double da = ddot(dir, a);
    for (int i = 0; i < 3; i++) u[i] = dir[i] - da * a[i];
    const double lu = sqrt(ddot(u, u));
    if (lu > 1e-6) {
        for (int i = 0; i < 3; i++) u[i] /= lu;
    } else {
        double aa[3] = {a[0], a[1], a[2]}, q[3];
        orc_plane_space(aa, u, q);
    }
    dcross(v, a, u);
    const double sgn[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
    for (int s = 0; s < 2; s++)
        for (int k = 0; k < 4; k++)
            for (int i = 0; i < 3; i++)
                pts[s * 4 + k][i] = o->p[i] + (s == 0 ? 0.5 : -0.5) * g->dims[1] * a[i] +
                                    g->dims[0] * (sgn[k][0] * u[i] + sgn[k][1] * v[i]);
}

/* the three sample points on a capsule's axis: '+' cap centre, '-' cap centre, and the axis point closest to
 * `target` (t = 2: it coincides with a cap centre and is skipped) */
static void capsule_samples(const orc_geom *g, const gpose *o, const double *target, double (*pts)[3], int *nmid)
{
    const double a[3] = {o->R[2], o->R[5], o->R[8]}, hl = 0.5 * g->dims[1];
    const double rel[3] = {target[0] - o->p[0], target[1] - o->p[1], target[2] - o->p[2]};
    double t = ddot(rel, a);
    *nmid = 1;
    if (t >= hl) t = hl, *nmid = 0;
    if (t <= -hl) t = -hl, *nmid = 0;
    for (int i = 0; i < 3; i++) {
        pts[0][i] = o->p[i] + hl * a[i];
        pts[1][i] = o->p[i] - hl * a[i];
        pts[2][i] = o->p[i] + t * a[i];
    }
}

/* capsule (first) vs box: sphere-box at the capsule's three axis samples */
static int collide_capsule_box(const orc_geom *gc, const gpose *oc, const orc_geom *gb, const gpose *ob, cpoint *out)
{
    double pts[3][3];
    int nmid, nout = 0;
    cpoint cand[3];
    capsule_samples(gc, oc, ob->p, pts, &nmid);
    for (int k = 0; k < 3; k++) {
        cand[k].depth = -1;
        if (k == 2 && !nmid) continue;
        if (!collide_sphere_box(pts[k], gc->dims[0], gb, ob, &cand[k])) cand[k].depth = -1;
    }
    keep_deepest(cand, 3, out, &nout);
    return nout;
}

/* capsule (first) vs cylinder: sphere-cylinder at the capsule's three axis samples */
static int collide_capsule_cylinder(const orc_geom *gc, const gpose *oc, const orc_geom *gy, const gpose *oy, cpoint *out)
{
    double pts[3][3];
    int nmid, nout = 0;
    cpoint cand[3];
    capsule_samples(gc, oc, oy->p, pts, &nmid);
    for (int k = 0; k < 3; k++) {
        cand[k].depth = -1;
        if (k == 2 && !nmid) continue;
        if (!collide_sphere_cylinder(pts[k], gc->dims[0], gy, oy, &cand[k])) cand[k].depth = -1;
    }
    keep_deepest(cand, 3, out, &nout);
    return nout;
}

/* direction from a point at `c` straight into a solid cylinder: minus the outward normal of the surface (cap or
 * side) the point is nearest to / would leave through */
static void into_cylinder(const double *c, const orc_geom *g, const gpose *o, double *dir)
{
    const double a[3] = {o->R[2], o->R[5], o->R[8]};
    const double rel[3] = {c[0] - o->p[0], c[1] - o->p[1], c[2] - o->p[2]};
    const double z = ddot(rel, a);
    double rv[3] = {rel[0] - z * a[0], rel[1] - z * a[1], rel[2] - z * a[2]};
    const double rho = sqrt(ddot(rv, rv));
    if (fabs(z) - 0.5 * g->dims[1] > rho - g->dims[0] || rho <= 1e-9) {
        const double sg = z >= 0 ? 1.0 : -1.0;
        for (int i = 0; i < 3; i++) dir[i] = -sg * a[i];
    } else {
        for (int i = 0; i < 3; i++) dir[i] = -rv[i] / rho;
    }
}

/* cylinder (first) vs box: the cylinder's rim samples (towards the box face its centre is nearest to) inside
 * the box, then the box's vertices inside the cylinder; the <= 4 deepest */
static int collide_cylinder_box(const orc_geom *gy, const gpose *oy, const orc_geom *gb, const gpose *ob, cpoint *out)
{
    cpoint cand[16];
    double rim[8][3], dir[3];
    int nout = 0;
    {
        const double rel[3] = {oy->p[0] - ob->p[0], oy->p[1] - ob->p[1], oy->p[2] - ob->p[2]};
        int ax = 0;
        double best = -1e300, lax = 0;
        for (int i = 0; i < 3; i++) {
            const double l = ob->R[i] * rel[0] + ob->R[3 + i] * rel[1] + ob->R[6 + i] * rel[2];
            const double m = fabs(l) - 0.5 * gb->dims[i];
            if (m > best) best = m, ax = i, lax = l;
        }
        const double sg = lax >= 0 ? 1.0 : -1.0;
        for (int i = 0; i < 3; i++) dir[i] = -sg * ob->R[3 * i + ax];
    }
    cylinder_rim(gy, oy, dir, rim);
    for (int k = 0; k < 8; k++)
        if (!collide_sphere_box(rim[k], 0.0, gb, ob, &cand[k])) cand[k].depth = -1;
    for (int k = 0; k < 8; k++) {
        double v[3];
        box_vertex(gb, ob, k, v);
        if (!collide_sphere_cylinder(v, 0.0, gy, oy, &cand[8 + k]))
            cand[8 + k].depth = -1;
        else
            for (int i = 0; i < 3; i++) cand[8 + k].normal[i] = -cand[8 + k].normal[i];
    }
    keep_deepest(cand, 16, out, &nout);
    return nout;
}

/* cylinder (first) vs cylinder: rim samples of each (towards the other's nearest surface) inside the other;
 * the <= 4 deepest */
static int collide_cylinder_cylinder(const orc_geom *ga, const gpose *oa, const orc_geom *gb, const gpose *ob, cpoint *out)
{
    cpoint cand[16];
    double rim[8][3], dir[3];
    int nout = 0;
    into_cylinder(oa->p, gb, ob, dir);
    cylinder_rim(ga, oa, dir, rim);
    for (int k = 0; k < 8; k++)
        if (!collide_sphere_cylinder(rim[k], 0.0, gb, ob, &cand[k])) cand[k].depth = -1;
    into_cylinder(ob->p, ga, oa, dir);
    cylinder_rim(gb, ob, dir, rim);
    for (int k = 0; k < 8; k++) {
        if (!collide_sphere_cylinder(rim[k], 0.0, ga, oa, &cand[8 + k]))
            cand[8 + k].depth = -1;
        else
            for (int i = 0; i < 3; i++) cand[8 + k].normal[i] = -cand[8 + k].normal[i];
    }
    keep_deepest(cand, 16, out, &nout);
    return nout;
}

/* ---- box (g1 = A) vs box (g2 = B): separating-axis test over the 15 axes, then either one
 * edge-edge contact or a reference-face / incident-face clip (Sutherland-Hodgman), keeping the <= 4
 * deepest clipped points in vertex order.  Same structure as ODE's dBoxBox (box.cpp), own
 * tie-breaking so that the result is well defined for the device kernel to match. ---- */
static void col3(const double *R, int k, double *out)
{
    out[0] = R[k];
    out[1] = R[3 + k];
    out[2] = R[6 + k];
}

static int clip_polygon(const double (*in)[3], int n, const double *pn, double pd, double (*out)[3])
{
    /* keep the part of the polygon with pn.x <= pd */
    int m = 0;
    for (int i = 0; i < n; i++) {
        const double *a = in[i], *b = in[(i + 1) % n];
        double da = ddot(pn, a) - pd, db = ddot(pn, b) - pd;
        if (da <= 0) {
            if (m < 8) memcpy(out[m++], a, 3 * sizeof(double));
        }
        if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
            double t = da / (da - db);
            if (m < 8) {
                for (int k = 0; k < 3; k++) out[m][k] = a[k] + t * (b[k] - a[k]);
                m++;
            }
        }
    }
    return m;
}

static int collide_box_box(const orc_geom *ga, const gpose *oa, const orc_geom *gb, const gpose *ob, cpoint *out)
{
    const double a[3] = {0.5 * ga->dims[0], 0.5 * ga->dims[1], 0.5 * ga->dims[2]};
    const double b[3] = {0.5 * gb->dims[0], 0.5 * gb->dims[1], 0.5 * gb->dims[2]};
    double d[3] = {ob->p[0] - oa->p[0], ob->p[1] - oa->p[1], ob->p[2] - oa->p[2]};
    double A[3][3], B[3][3], Rm[3][3], Q[3][3], pp[3];
    for (int k = 0; k < 3; k++) {
        col3(oa->R, k, A[k]);
        col3(ob->R, k, B[k]);
    }
    for (int i = 0; i < 3; i++) {
        pp[i] = ddot(A[i], d);
        for (int j = 0; j < 3; j++) {
            Rm[i][j] = ddot(A[i], B[j]);
            Q[i][j] = fabs(Rm[i][j]) + 1e-5;
        }
    }
    double best = -1e300, nrm[3] = {0, 0, 0};
    int code = 0;
    for (int i = 0; i < 3; i++) { /* face normals of A */
        double s = fabs(pp[i]) - (a[i] + b[0] * Q[i][0] + b[1] * Q[i][1] + b[2] * Q[i][2]);
        if (s > 0) return 0;
        if (s > best) {
            best = s;
            code = 1 + i;
            memcpy(nrm, A[i], sizeof(nrm));
        }
    }
    for (int j = 0; j < 3; j++) { /* face normals of B */
        double s = fabs(ddot(B[j], d)) - (a[0] * Q[0][j] + a[1] * Q[1][j] + a[2] * Q[2][j] + b[j]);
        if (s > 0) return 0;
        if (s > best) {
            best = s;
            code = 4 + j;
            memcpy(nrm, B[j], sizeof(nrm));
        }
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { /* edge x edge */
            double n[3];
            dcross(n, A[i], B[j]);
            double l2 = ddot(n, n);
            if (l2 < 1e-10) continue;
            double l = sqrt(l2), ra = 0, rb = 0;
            for (int k = 0; k < 3; k++) {
                ra += a[k] * fabs(ddot(A[k], n));
                rb += b[k] * fabs(ddot(B[k], n));
            }
            double s = (fabs(ddot(d, n)) - (ra + rb)) / l;
            if (s > 0) return 0;
            if (s * 1.05 > best) { /* an edge axis must beat the face axes by 5 % (ODE fudge factor) */
                best = s;
                code = 7 + 3 * i + j;
                for (int k = 0; k < 3; k++) nrm[k] = n[k] / l;
            }
        }
    if (!code) return 0;
    if (ddot(nrm, d) < 0) /* make nrm point from A to B */
        for (int k = 0; k < 3; k++) nrm[k] = -nrm[k];

    if (code >= 7) {
        const int i = (code - 7) / 3, j = (code - 7) % 3;
        double pa[3], pb[3];
        memcpy(pa, oa->p, sizeof(pa));
        memcpy(pb, ob->p, sizeof(pb));
        for (int k = 0; k < 3; k++) {
            if (k != i) {
                double sg = ddot(nrm, A[k]) > 0 ? 1.0 : -1.0;
                for (int c = 0; c < 3; c++) pa[c] += sg * a[k] * A[k][c];
            }
            if (k != j) {
                double sg = ddot(nrm, B[k]) > 0 ? -1.0 : 1.0;
                for (int c = 0; c < 3; c++) pb[c] += sg * b[k] * B[k][c];
            }
        }
        /* closest approach of the lines pa + alpha*A[i], pb + beta*B[j] */
        double p[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
        double uaub = ddot(A[i], B[j]), q1 = ddot(A[i], p), q2 = -ddot(B[j], p), dd = 1 - uaub * uaub;
        double alpha = 0, beta = 0;
        if (dd > 1e-4) {
            alpha = (q1 + uaub * q2) / dd;
            beta = (uaub * q1 + q2) / dd;
        }
        for (int c = 0; c < 3; c++) {
            out[0].pos[c] = 0.5 * (pa[c] + alpha * A[i][c] + pb[c] + beta * B[j][c]);
            out[0].normal[c] = -nrm[c];
        }
        out[0].depth = -best;
        return 1;
    }

    /* face contact: reference box has the separating face, the other box provides the incident face */
    const int ref_is_a = code <= 3;
    const double(*Rr)[3] = ref_is_a ? A : B, (*Ri)[3] = ref_is_a ? B : A;
    const double *pr = ref_is_a ? oa->p : ob->p, *pi = ref_is_a ? ob->p : oa->p;
    const double *hr = ref_is_a ? a : b, *hi = ref_is_a ? b : a;
    const int ra = ref_is_a ? code - 1 : code - 4;
    double nref[3]; /* outward normal of the reference face = from the reference box to the other */
    for (int k = 0; k < 3; k++) nref[k] = ref_is_a ? nrm[k] : -nrm[k];
    /* incident face: the face of the other box most anti-parallel to nref */
    int ia = 0;
    double bestd = -1;
    for (int k = 0; k < 3; k++) {
        double v = fabs(ddot(Ri[k], nref));
        if (v > bestd) {
            bestd = v;
            ia = k;
        }
    }
    const double isg = ddot(Ri[ia], nref) > 0 ? -1.0 : 1.0;
    const int i1 = (ia + 1) % 3, i2 = (ia + 2) % 3;
    double poly[8][3], tmp[8][3];
    const double su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
    for (int v = 0; v < 4; v++)
        for (int c = 0; c < 3; c++)
            poly[v][c] = pi[c] + isg * hi[ia] * Ri[ia][c] + su[v] * hi[i1] * Ri[i1][c] + sv[v] * hi[i2] * Ri[i2][c];
    int n = 4;
    const int r1 = (ra + 1) % 3, r2 = (ra + 2) % 3;
    const int axes[4] = {r1, r1, r2, r2};
    const double sgn[4] = {1, -1, 1, -1};
    for (int e = 0; e < 4 && n > 0; e++) {
        double pn[3];
        for (int c = 0; c < 3; c++) pn[c] = sgn[e] * Rr[axes[e]][c];
        double pd = ddot(pn, pr) + hr[axes[e]];
        n = clip_polygon((const double(*)[3])poly, n, pn, pd, tmp);
        memcpy(poly, tmp, sizeof(poly));
    }
    cpoint cand[8];
    const double face_d = ddot(nref, pr) + hr[ra];
    for (int v = 0; v < 8; v++) {
        cand[v].depth = -1;
        if (v < n) {
            cand[v].depth = face_d - ddot(nref, poly[v]);
            memcpy(cand[v].pos, poly[v], 3 * sizeof(double));
            for (int c = 0; c < 3; c++) cand[v].normal[c] = -nrm[c];
        }
    }
    int nout = 0;
    keep_deepest(cand, 8, out, &nout);
    return nout;
}

/* closest points of two segments p1+s*d1, p2+t*d2, s,t in [0,1] (Ericson, RTCD 5.1.9) */
static void closest_seg_seg(const double *p1, const double *d1, const double *p2, const double *d2, double *c1, double *c2)
{
    double r[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double a = ddot(d1, d1), e = ddot(d2, d2), f = ddot(d2, r), s, t;
    const double EPS = 1e-12;
    if (a <= EPS && e <= EPS) {
        s = t = 0;
    } else if (a <= EPS) {
        s = 0;
        t = f / e;
        t = t < 0 ? 0 : (t > 1 ? 1 : t);
    } else {
        double c = ddot(d1, r);
        if (e <= EPS) {
            t = 0;
            s = -c / a;
            s = s < 0 ? 0 : (s > 1 ? 1 : s);
        } else {
            double b = ddot(d1, d2), denom = a * e - b * b;
            if (denom > EPS) {
                s = (b * f - c * e) / denom;
                s = s < 0 ? 0 : (s > 1 ? 1 : s);
            } else
                s = 0;
            t = (b * s + f) / e;
            if (t < 0) {
                t = 0;
                s = -c / a;
                s = s < 0 ? 0 : (s > 1 ? 1 : s);
            } else if (t > 1) {
                t = 1;
                s = (b - c) / a;
                s = s < 0 ? 0 : (s > 1 ? 1 : s);
            }
        }
    }
    for (int i = 0; i < 3; i++) {
        c1[i] = p1[i] + s * d1[i];
        c2[i] = p2[i] + t * d2[i];
    }
}

static void capsule_segment(const orc_geom *g, const gpose *o, double *p, double *d)
{
    for (int i = 0; i < 3; i++) {
        p[i] = o->p[i] - 0.5 * g->dims[1] * o->R[3 * i + 2];
        d[i] = g->dims[1] * o->R[3 * i + 2];
    }
}

/* ---- heightfield ---- */
static double hf_h(const orc_world *w, int ix, int iy) { return (double)w->hf_heights[(size_t)iy * w->hf_nx + ix]; }

/* triangle of cell (ix,iy): which = 0 -> (00,10,11), 1 -> (00,11,01) */
static void hf_triangle(const orc_world *w, int ix, int iy, int which, double *a, double *b, double *c)
{
    double x0 = w->hf_x0 + ix * w->hf_dx, y0 = w->hf_y0 + iy * w->hf_dx, x1 = x0 + w->hf_dx, y1 = y0 + w->hf_dx;
    a[0] = x0, a[1] = y0, a[2] = hf_h(w, ix, iy);
    if (which == 0) {
        b[0] = x1, b[1] = y0, b[2] = hf_h(w, ix + 1, iy);
        c[0] = x1, c[1] = y1, c[2] = hf_h(w, ix + 1, iy + 1);
    } else {
        b[0] = x1, b[1] = y1, b[2] = hf_h(w, ix + 1, iy + 1);
        c[0] = x0, c[1] = y1, c[2] = hf_h(w, ix, iy + 1);
    }
}

static void tri_normal(const double *a, const double *b, const double *c, double *n)
{
    double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    dcross(n, ab, ac);
    double l = sqrt(ddot(n, n));
    n[0] /= l, n[1] /= l, n[2] /= l;
}

/* closest point on triangle abc to p (Ericson, RTCD 5.1.5) */
static void closest_pt_triangle(const double *p, const double *a, const double *b, const double *c, double *q)
{
    double ab[3], ac[3], ap[3], bp[3], cp[3];
    for (int i = 0; i < 3; i++) {
        ab[i] = b[i] - a[i];
        ac[i] = c[i] - a[i];
        ap[i] = p[i] - a[i];
    }
    double d1 = ddot(ab, ap), d2 = ddot(ac, ap);
    if (d1 <= 0 && d2 <= 0) {
        memcpy(q, a, 3 * sizeof(double));
        return;
    }
    for (int i = 0; i < 3; i++) bp[i] = p[i] - b[i];
    double d3 = ddot(ab, bp), d4 = ddot(ac, bp);
    if (d3 >= 0 && d4 <= d3) {
        memcpy(q, b, 3 * sizeof(double));
        return;
    }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        double v = d1 / (d1 - d3);
        for (int i = 0; i < 3; i++) q[i] = a[i] + v * ab[i];
        return;
    }
    for (int i = 0; i < 3; i++) cp[i] = p[i] - c[i];
    double d5 = ddot(ab, cp), d6 = ddot(ac, cp);
    if (d6 >= 0 && d5 <= d6) {
        memcpy(q, c, 3 * sizeof(double));
        return;
    }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        double wv = d2 / (d2 - d6);
        for (int i = 0; i < 3; i++) q[i] = a[i] + wv * ac[i];
        return;
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double wv = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int i = 0; i < 3; i++) q[i] = b[i] + wv * (c[i] - b[i]);
        return;
    }
    double denom = 1.0 / (va + vb + vc), v = vb * denom, wv = vc * denom;
    for (int i = 0; i < 3; i++) q[i] = a[i] + ab[i] * v + ac[i] * wv;
}

/* surface height + triangle normal under (x,y); returns 0 outside the field */
static int hf_sample(const orc_world *w, double x, double y, double *h, double *n)
{
    double fx = (x - w->hf_x0) / w->hf_dx, fy = (y - w->hf_y0) / w->hf_dx;
    if (fx < 0 || fy < 0 || fx > w->hf_nx - 1 || fy > w->hf_ny - 1) return 0;
    int ix = (int)floor(fx), iy = (int)floor(fy);
    if (ix > w->hf_nx - 2) ix = w->hf_nx - 2;
    if (iy > w->hf_ny - 2) iy = w->hf_ny - 2;
    double u = fx - ix, v = fy - iy, a[3], b[3], c[3];
    hf_triangle(w, ix, iy, (v > u) ? 1 : 0, a, b, c);
    tri_normal(a, b, c, n);
    /* plane through a with normal n */
    *h = a[2] - (n[0] * (x - a[0]) + n[1] * (y - a[1])) / n[2];
    return 1;
}

static int collide_sphere_hf(const orc_world *w, const double *c, double r, cpoint *out)
{
    if (!w->hf_heights) return 0;
    int ix0 = (int)floor((c[0] - r - w->hf_x0) / w->hf_dx), ix1 = (int)floor((c[0] + r - w->hf_x0) / w->hf_dx);
    int iy0 = (int)floor((c[1] - r - w->hf_y0) / w->hf_dx), iy1 = (int)floor((c[1] + r - w->hf_y0) / w->hf_dx);
    if (ix0 < 0) ix0 = 0;
    if (iy0 < 0) iy0 = 0;
    if (ix1 > w->hf_nx - 2) ix1 = w->hf_nx - 2;
    if (iy1 > w->hf_ny - 2) iy1 = w->hf_ny - 2;
    double best = -1;
    /* centre below the surface: push out along the local triangle normal */
    {
        double h, n[3];
        if (hf_sample(w, c[0], c[1], &h, n) && c[2] < h) {
            best = r + (h - c[2]) * n[2];
            out->depth = best;
            memcpy(out->normal, n, sizeof(n));
            out->pos[0] = c[0];
            out->pos[1] = c[1];
            out->pos[2] = h;
        }
    }
    for (int iy = iy0; iy <= iy1; iy++)
        for (int ix = ix0; ix <= ix1; ix++)
            for (int t = 0; t < 2; t++) {
                double a[3], b[3], cc[3], q[3], n[3];
                hf_triangle(w, ix, iy, t, a, b, cc);
                closest_pt_triangle(c, a, b, cc, q);
                double dv[3] = {c[0] - q[0], c[1] - q[1], c[2] - q[2]};
                double dist = sqrt(ddot(dv, dv));
                if (dist > r) continue;
                tri_normal(a, b, cc, n);
                if (ddot(dv, n) < 0) continue; /* centre behind this triangle: handled above */
                double depth = r - dist;
                if (depth > best) {
                    best = depth;
                    out->depth = depth;
                    memcpy(out->pos, q, sizeof(q));
                    if (dist > 1e-9)
                        for (int i = 0; i < 3; i++) out->normal[i] = dv[i] / dist;
                    else
                        memcpy(out->normal, n, sizeof(n));
                }
            }
    return best >= 0 ? 1 : 0;
}

static int collide_box_hf(const orc_world *w, const orc_geom *g, const gpose *o, cpoint *out)
{
    if (!w->hf_heights) return 0;
    cpoint cand[8];
    int nout = 0;
    for (int k = 0; k < 8; k++) {
        double h, n[3];
        box_vertex(g, o, k, cand[k].pos);
        cand[k].depth = -1;
        if (hf_sample(w, cand[k].pos[0], cand[k].pos[1], &h, n) && cand[k].pos[2] <= h) {
            cand[k].depth = (h - cand[k].pos[2]) * n[2];
            memcpy(cand[k].normal, n, sizeof(n));
        }
    }
    keep_deepest(cand, 8, out, &nout);
    return nout;
}


/* capsule vs heightfield: sphere-heightfield at both cap centres and the capsule's centre */
static int collide_capsule_hf(const orc_world *w, const orc_geom *g, const gpose *o, cpoint *out)
{
    if (!w->hf_heights) return 0;
    double pts[3][3];
    int nmid, nout = 0;
    cpoint cand[3];
    capsule_samples(g, o, o->p, pts, &nmid);
    for (int k = 0; k < 3; k++)
        if (!collide_sphere_hf(w, pts[k], g->dims[0], &cand[k])) cand[k].depth = -1;
    keep_deepest(cand, 3, out, &nout);
    return nout;
}

/* cylinder vs heightfield: the rim samples (towards -z) below the surface, as the box's vertices */
static int collide_cylinder_hf(const orc_world *w, const orc_geom *g, const gpose *o, cpoint *out)
{
    if (!w->hf_heights) return 0;
    cpoint cand[8];
    double rim[8][3];
    const double down[3] = {0, 0, -1};
    int nout = 0;
    cylinder_rim(g, o, down, rim);
    for (int k = 0; k < 8; k++) {
        double h, n[3];
        memcpy(cand[k].pos, rim[k], sizeof(rim[k]));
        cand[k].depth = -1;
        if (hf_sample(w, rim[k][0], rim[k][1], &h, n) && rim[k][2] <= h) {
            cand[k].depth = (h - rim[k][2]) * n[2];
            memcpy(cand[k].normal, n, sizeof(n));
        }
    }
    keep_deepest(cand, 8, out, &nout);
    return nout;
}

/* ---- dispatch: g1 dynamic-or-static, g2; returns contacts with normal from g2 into g1 ---- */
static void flip(cpoint *c, int n)
{
    for (int k = 0; k < n; k++)
        for (int i = 0; i < 3; i++) c[k].normal[i] = -c[k].normal[i];
}

static int narrowphase(const orc_world *w, const orc_geom *g1, const orc_geom *g2, cpoint *out)
{
    gpose o1, o2;
    if (g1->body >= 0) geom_pose(w, g1, &o1);
    if (g2->body >= 0) geom_pose(w, g2, &o2);
    const int c1 = g1->cls, c2 = g2->cls;
    if (c2 == 4) return collide_plane(g1, &o1, g2->dims, out);
    if (c2 == 5) {
        if (c1 == 0) return collide_sphere_hf(w, o1.p, g1->dims[0], out);
        if (c1 == 1) return collide_box_hf(w, g1, &o1, out);
        if (c1 == 2) return collide_capsule_hf(w, g1, &o1, out);
        if (c1 == 3) return collide_cylinder_hf(w, g1, &o1, out);
        return 0;
    }
    if (c1 == 1 && c2 == 1) return collide_box_box(g1, &o1, g2, &o2, out);
    if (c1 == 0 && c2 == 0) return collide_sphere_sphere(o1.p, g1->dims[0], o2.p, g2->dims[0], out);
    if (c1 == 0 && c2 == 1) return collide_sphere_box(o1.p, g1->dims[0], g2, &o2, out);
    if (c1 == 1 && c2 == 0) {
        int n = collide_sphere_box(o2.p, g2->dims[0], g1, &o1, out);
        flip(out, n);
        return n;
    }
    if ((c1 == 2 && c2 == 0) || (c1 == 0 && c2 == 2)) {
        const orc_geom *gc = c1 == 2 ? g1 : g2;
        const gpose *oc = c1 == 2 ? &o1 : &o2, *os = c1 == 2 ? &o2 : &o1;
        const orc_geom *gs = c1 == 2 ? g2 : g1;
        double p[3], d[3], zero[3] = {0, 0, 0}, q1[3], q2[3];
        capsule_segment(gc, oc, p, d);
        closest_seg_seg(p, d, os->p, zero, q1, q2);
        int n;
        if (c1 == 2)
            n = collide_sphere_sphere(q1, gc->dims[0], os->p, gs->dims[0], out);
        else
            n = collide_sphere_sphere(os->p, gs->dims[0], q1, gc->dims[0], out);
        return n;
    }
    if (c1 == 2 && c2 == 2) {
        double p1[3], d1[3], p2[3], d2[3], q1[3], q2[3];
        capsule_segment(g1, &o1, p1, d1);
        capsule_segment(g2, &o2, p2, d2);
        closest_seg_seg(p1, d1, p2, d2, q1, q2);
        return collide_sphere_sphere(q1, g1->dims[0], q2, g2->dims[0], out);
    }
    int n = 0;
    if (c1 == 2 && c2 == 1) return collide_capsule_box(g1, &o1, g2, &o2, out);
    if (c1 == 1 && c2 == 2) {
        n = collide_capsule_box(g2, &o2, g1, &o1, out);
        flip(out, n);
        return n;
    }
    if (c1 == 0 && c2 == 3) return collide_sphere_cylinder(o1.p, g1->dims[0], g2, &o2, out);
    if (c1 == 3 && c2 == 0) {
        n = collide_sphere_cylinder(o2.p, g2->dims[0], g1, &o1, out);
        flip(out, n);
        return n;
    }
    if (c1 == 2 && c2 == 3) return collide_capsule_cylinder(g1, &o1, g2, &o2, out);
    if (c1 == 3 && c2 == 2) {
        n = collide_capsule_cylinder(g2, &o2, g1, &o1, out);
        flip(out, n);
        return n;
    }
    if (c1 == 3 && c2 == 1) return collide_cylinder_box(g1, &o1, g2, &o2, out);
    if (c1 == 1 && c2 == 3) {
        n = collide_cylinder_box(g2, &o2, g1, &o1, out);
        flip(out, n);
        return n;
    }
    if (c1 == 3 && c2 == 3) return collide_cylinder_cylinder(g1, &o1, g2, &o2, out);
    return 0; /* planes / heightfields against each other */
}

/* ------------------------------------------------------------------ */
/* public API                                                          */

int orc_geom_create(orc_world *w, int cls, int body, const double dims[4], const double lpos[3], const double lquat[4],
                    const orc_surface *s)
{
    if (w->ng == w->capg) {
        w->capg = w->capg ? 2 * w->capg : 16;
        w->geoms = (orc_geom *)realloc(w->geoms, sizeof(orc_geom) * (size_t)w->capg);
    }
    orc_geom *g = &w->geoms[w->ng];
    memset(g, 0, sizeof(*g));
    g->cls = cls;
    g->body = body;
    g->lquat[0] = 1;
    if (dims) memcpy(g->dims, dims, sizeof(g->dims));
    if (lpos) memcpy(g->lpos, lpos, sizeof(g->lpos));
    if (lquat) {
        double l = sqrt(lquat[0] * lquat[0] + lquat[1] * lquat[1] + lquat[2] * lquat[2] + lquat[3] * lquat[3]);
        for (int i = 0; i < 4; i++) g->lquat[i] = lquat[i] / l;
    }
    if (s) {
        g->mu = s->mu;
        g->mu2 = s->mu2;
        g->rho = s->rho;
        g->rho2 = s->rho2;
        g->rhoN = s->rhoN;
        g->soft_erp = s->soft_erp;
        g->soft_cfm = s->soft_cfm;
    }
    return w->ng++;
}

int orc_geom_create_heightfield(orc_world *w, const float *heights, int nx, int ny, double dx, double x0, double y0,
                                const orc_surface *s)
{
    free(w->hf_heights);
    w->hf_heights = (float *)malloc(sizeof(float) * (size_t)nx * ny);
    memcpy(w->hf_heights, heights, sizeof(float) * (size_t)nx * ny);
    w->hf_nx = nx;
    w->hf_ny = ny;
    w->hf_dx = dx;
    w->hf_x0 = x0;
    w->hf_y0 = y0;
    w->hf_hmax = -1e30f;
    for (size_t i = 0; i < (size_t)nx * ny; i++)
        if (heights[i] > w->hf_hmax) w->hf_hmax = heights[i];
    return orc_geom_create(w, 5, -1, 0, 0, 0, s);
}

void orc_world_set_max_contacts(orc_world *w, int n) { w->max_contacts = n; }

static double dmin(double a, double b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }

int orc_collide(orc_world *w)
{
    orc_contacts_clear(w);
    w->npairs = 0;
    const int ng = w->ng;
    double *cen = (double *)malloc(sizeof(double) * 3 * (size_t)(ng > 0 ? ng : 1));
    double *ext = (double *)malloc(sizeof(double) * 3 * (size_t)(ng > 0 ? ng : 1));
    const float hmax = w->hf_hmax;
    for (int g = 0; g < ng; g++) {
        if (w->geoms[g].body >= 0) {
            gpose o;
            geom_pose(w, &w->geoms[g], &o);
            geom_aabb(&w->geoms[g], &o, cen + 3 * g, ext + 3 * g);
        }
    }
    int total = 0;
    for (int i = 0; i < ng; i++)
        for (int j = i + 1; j < ng; j++) {
            const orc_geom *gi = &w->geoms[i], *gj = &w->geoms[j];
            if (gi->body < 0 && gj->body < 0) continue;
            if (gi->body == gj->body) continue;
            if (gi->body >= 0 && gj->body >= 0 && orc_are_connected_excluding(w, gi->body, gj->body, ORC_JOINT_CONTACT))
                continue;
            /* broadphase */
            int overlap;
            const orc_geom *gd = gi->body >= 0 ? gi : gj, *gs = gi->body >= 0 ? gj : gi; /* dynamic, other */
            const int d_idx = gi->body >= 0 ? i : j;
            if (gi->body >= 0 && gj->body >= 0) {
                overlap = 1;
                for (int k = 0; k < 3; k++)
                    if (cen[3 * i + k] - ext[3 * i + k] > cen[3 * j + k] + ext[3 * j + k] ||
                        cen[3 * j + k] - ext[3 * j + k] > cen[3 * i + k] + ext[3 * i + k])
                        overlap = 0;
            } else if (gs->cls == 4) {
                const double *n = gs->dims, *c = cen + 3 * d_idx, *e = ext + 3 * d_idx;
                overlap = (ddot(n, c) - (fabs(n[0]) * e[0] + fabs(n[1]) * e[1] + fabs(n[2]) * e[2])) <= n[3];
            } else if (gs->cls == 5) {
                const double *c = cen + 3 * d_idx, *e = ext + 3 * d_idx;
                double xe = w->hf_x0 + (w->hf_nx - 1) * w->hf_dx, ye = w->hf_y0 + (w->hf_ny - 1) * w->hf_dx;
                overlap = (c[2] - e[2] <= (double)hmax) && c[0] + e[0] >= w->hf_x0 && c[0] - e[0] <= xe &&
                          c[1] + e[1] >= w->hf_y0 && c[1] - e[1] <= ye;
            } else {
                overlap = 0;
            }
            if (!overlap) continue;
            if (w->npairs == w->cappairs) {
                w->cappairs = w->cappairs ? 2 * w->cappairs : 64;
                w->pairs = (int *)realloc(w->pairs, sizeof(int) * (size_t)w->cappairs);
            }
            w->pairs[w->npairs++] = i | (j << 16);
            /* narrowphase: a static geom always plays geom 2 */
            cpoint pts[MAX_PAIR_CONTACTS];
            int n = narrowphase(w, gd == gi ? gi : gj, gd == gi ? gj : gi, pts);
            const orc_geom *g1 = gd == gi ? gi : gj, *g2 = gd == gi ? gj : gi;
            for (int k = 0; k < n; k++) {
                if (total >= w->max_contacts) break;
                /* surface combination + the mode bits of Frame::addContact (Frame.cpp:1038-1079) */
                orc_contact_desc c;
                memset(&c, 0, sizeof(c));
                memcpy(c.pos, pts[k].pos, sizeof(c.pos));
                memcpy(c.normal, pts[k].normal, sizeof(c.normal));
                c.depth = pts[k].depth;
                c.mu = dmin(g1->mu, g2->mu);
                c.mu2 = dmin(g1->mu2, g2->mu2);
                c.rho = dmin(g1->rho, g2->rho);
                c.rho2 = dmin(g1->rho2, g2->rho2);
                c.rhoN = dmin(g1->rhoN, g2->rhoN);
                c.soft_erp = dmin(g1->soft_erp, g2->soft_erp);
                c.soft_cfm = dmax(g1->soft_cfm, g2->soft_cfm);
                c.mode = ORC_CONTACT_SOFT_ERP | ORC_CONTACT_SOFT_CFM | ORC_CONTACT_APPROX1;
                if (c.rho > 1e-10) c.mode |= ORC_CONTACT_ROLLING;
                if (c.mu != c.mu2) c.mode |= ORC_CONTACT_MU2;
                if (orc_contact_add(w, g1->body, g2->body, &c) >= 0) total++;
            }
        }
    free(cen);
    free(ext);
    return total;
}

int orc_collide_get_pairs(const orc_world *w, int *pairs, int max_pairs)
{
    for (int i = 0; i < w->npairs && i < max_pairs; i++) pairs[i] = w->pairs[i];
    return w->npairs;
}

/* read back contact c of the current group: bodies (after attach normalisation) and record */
int orc_contact_get(const orc_world *w, int c, int *b1, int *b2, orc_contact_desc *out)
{
    if (c < 0 || c >= w->nc) return -1;
    const joint_t *j = &w->contacts[c];
    if (b1) *b1 = j->b[0];
    if (b2) *b2 = j->b[1];
    if (out) {
        *out = j->contact;
        if (j->reverse)
            for (int i = 0; i < 3; i++) out->normal[i] = -out->normal[i];
    }
    return 0;
}

/*
 * oracle/orc_internal.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 * Internal data structures shared by orc_quickstep.c and orc_collide.c.  See orc.h.
 */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H
#include "orc.h"
#include <math.h>

#ifdef ORC_SINGLE
typedef float real;
#define R_SQRT sqrtf
#define R_FABS fabsf
#define R_ATAN2 atan2f
#define ORC_DEFAULT_CFM 1e-5
#else
typedef double real;
#define R_SQRT sqrt
#define R_FABS fabs
#define R_ATAN2 atan2
#define ORC_DEFAULT_CFM 1e-10
/* orc_lcp.c: Dantzig box-LCP solver of dWorldStep */
int orc_solve_lcp(int n, const double *A, double *x, const double *b, double *w, double *lo, double *hi,
                  const int *findex);

#endif


#define R_INF ((real)INFINITY)

typedef struct orc_geom {
    int cls, body;
    double dims[4], lpos[3], lquat[4];
    double mu, mu2, rho, rho2, rhoN, soft_erp, soft_cfm;
} orc_geom;

/* ------------------------------------------------------------------ */
/* data                                                                */

typedef struct {
    real vel, fmax, lostop, histop, fudge_factor, normal_cfm, stop_erp, stop_cfm, bounce;
    int limit;
    real limit_err;
} limot_t;

typedef struct {
    real pos[3], q[4], R[9], lvel[3], avel[3], facc[3], tacc[3];
    real mass, invMass, I[9], invI[9];
    int gyroscopic, nogravity, has_max_ang;
    real max_angular_speed;
    /* per-step scratch */
    real invI_w[9];
    int tag;
} body_t;

typedef struct {
    int type;
    int b[2]; /* node[0].body, node[1].body after attach (index or -1) */
    int reverse;
    unsigned creation_serial; /* attach order, for the island DFS */
    real anchor1[3], anchor2[3], axis1[3], axis2[3], qrel[4], offset[3];
    real qrel2[4]; /* universal: qrel = qrel1 (body 1 vs the cross), qrel2 (body 2 vs the cross) */
    limot_t limot1, limot2;
    /* hinge2 */
    real c0, s0, v1[3], v2[3], w1[3], w2[3], susp_erp, susp_cfm;
    /* fixed */
    real erp, cfm;
    /* contact */
    orc_contact_desc contact;
    /* outputs */
    real fb_f1[3], fb_t1[3], fb_f2[3], fb_t2[3], fb_lambda;
    int last_m;
    int limot_row; /* local row of the limot that feeds fb_lambda, or -1 */
    int tag;
} joint_t;

struct orc_world {
    real gravity[3], erp, cfm, max_angular_speed, contact_max_vel, contact_min_depth, sor_w;
    int iters, order_mode, reorder_method;
    int stepper;        /* ORC_STEPPER_QUICK / ORC_STEPPER_DENSE (orc.h) */
    int last_lcp_error; /* dense stepper: the LCP solver gave up in some island of the last step */
    int variant[ORC_NUM_VARIANTS]; /* restatement choices that only a real libode can settle (orc.h) */
    uint32_t seed;
    body_t *bodies;
    int nb, capb;
    joint_t *joints;
    int nj, capj;
    joint_t *contacts;
    int nc, capc;
    unsigned serial;
    int last_m, last_islands;
    real last_residual;
    /* collision (orc_collide.c) */
    orc_geom *geoms;
    int ng, capg;
    float *hf_heights;
    int hf_nx, hf_ny;
    double hf_dx, hf_x0, hf_y0;
    float hf_hmax;
    int max_contacts;
    int *pairs;
    int npairs, cappairs;
};


/* orc_lcp.c: Dantzig box-LCP solver of dWorldStep */
int orc_solve_lcp(int n, const double *A, double *x, const double *b, double *w, double *lo, double *hi,
                  const int *findex);

#endif

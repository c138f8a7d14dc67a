// b2p_assemble.cu -- kernel 1 of the step: body preparation and constraint-row assembly (sm_100a).
//
// One thread owns one world; every global access is a world-fastest SoA access (one fully used
// 128-byte line per field per warp).  It computes stages 0-2 of dWorldQuickStep as the reference
// calls it at src/WorldPhysics.cpp:365 (ODE 0.16 quickstep.cpp): invI_world, implicit gyroscopic
// torque, gravity, tmp1 = v/h + invM*fe, then getInfo1/getInfo2 of every joint (hinge, fixed, slider,
// hinge2 incl. dxJointLimitMotor) and of every contact in the stream (mode bits as
// src/objects/Frame.cpp:1046-1071 sets them).
//
// Row form handed to the SOR kernel (b2p_sor.cu).  ODE keeps, per row, J (12), iMJ = invM J^T (12),
// rhs, Ad, cfm, lo, hi, findex and sweeps  delta = Ad*(rhs - cfm*lambda - J.fc); fc += delta*iMJ.
// Here the row is scaled once so that the sweep needs a single 12-vector:
//     Jh   = sqrt(Ad) * J * M^-1/2          (M^-1/2 = diag(sqrt(invMass) I3, R sqrt(invI_body) R^T))
//     lam' = lambda / sqrt(Ad),  rhs' = sqrt(Ad)*rhs,  cfm' = Ad*cfm,  bounds / sqrt(Ad)
//     fch  = M^-1/2 J^T lambda  (so that invM J^T lambda = M^-1/2 fch)
//     sweep:  delta' = rhs' - cfm'*lam' - Jh.fch ;  fch += delta' * Jh
// which is the same Gauss-Seidel iteration in different coordinates (Ad = w / (|J M^-1/2|^2 + cfm)).
// A friction row's bound mu*lambda_n becomes (mu*sqrt(Ad_n/Ad_f)) * lam'_n.
//
// Rows are stored by constraint (static slots: joint j owns slots j*B2P_ROWS_PER_JOINT .., contact c owns
// njslots + c*rpc ..); `ord0` lists the used slots in ODE's initial sweep order (per island: rows with findex < 0
// first, in ODE's joint order, then the friction-dependent rows) and seeds the SOR kernels' order array.
#include <mutex>

#include "b2p_rows.cuh"

namespace {

constexpr int BD = 64; // threads (= worlds) per CTA

struct Smem {
    const DevBody *bodies;   // the scene template's bodies and joints, copied into shared memory by the CTA: the
    const DevJoint *joints;  // kernel reads them all the time, always warp-uniformly (one broadcast LDS instead
                             // of a global load on the thread's dependent chain)
    float *t1; // [nb*6][BD]  fe -> tmp1 = v/h + invM fe
    float *sw; // [nb*6][BD]  symmetric sqrt(invI_world): xx xy xz yy yz zz
    float *pq; // [nb*7][BD]  position and quaternion of every body: the joints and contacts re-read the poses
               //             of their bodies, and a shared-memory read costs a tenth of a global one
    int nb, maxrows;
    uint8_t *scratch; // [nj + 3 maxc (+ nj + maxc + nb)][BD] bytes: per-item row bytes, contact body ids (+ island search)
};

// pose of body b from the copy phase 1 left in shared memory
__device__ __forceinline__ void load_pose_s(const Smem &S, int b, int tid, BodyPose &bp)
{
    const float *p = S.pq + (size_t)(b * 7) * BD + tid;
    bp.pos[0] = p[0], bp.pos[1] = p[BD], bp.pos[2] = p[2 * BD];
    bp.q[0] = p[3 * BD], bp.q[1] = p[4 * BD], bp.q[2] = p[5 * BD], bp.q[3] = p[6 * BD];
    q_to_R(bp.q, bp.R);
}

struct PairCtx {
    int b0, b1;      // b1 < 0: static
    float sm0, sm1;  // sqrt(invMass)
    float S0[6], S1[6];
};

__device__ __forceinline__ void pair_ctx(PairCtx &pc, const Smem &S, int tid)
{
    pc.sm0 = S.bodies[pc.b0].sqrtInvMass;
#pragma unroll
    for (int k = 0; k < 6; k++) pc.S0[k] = S.sw[(size_t)(pc.b0 * 6 + k) * BD + tid];
    pc.sm1 = 0.f;
    if (pc.b1 >= 0) {
        pc.sm1 = S.bodies[pc.b1].sqrtInvMass;
#pragma unroll
        for (int k = 0; k < 6; k++) pc.S1[k] = S.sw[(size_t)(pc.b1 * 6 + k) * BD + tid];
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) pc.S1[k] = 0.f;
    }
}

__device__ __forceinline__ void sym_mul(float *o, const float *Ssym, const float *v)
{
    o[0] = Ssym[0] * v[0] + Ssym[1] * v[1] + Ssym[2] * v[2];
    o[1] = Ssym[1] * v[0] + Ssym[3] * v[1] + Ssym[4] * v[2];
    o[2] = Ssym[2] * v[0] + Ssym[4] * v[1] + Ssym[5] * v[2];
}

// rhs = c/h - J.tmp1 ; Jt = J M^-1/2 ; Ad = w/(|Jt|^2 + cfm/h) ; scale by sqrt(Ad) ; store at `slot`.
// Returns sqrt(Ad).
__device__ __forceinline__ float emit_row(const StepParams &P, const Smem &S, int w, int tid, int slot, const Row &r,
                                          const PairCtx &pc)
{
    const int R = S.maxrows, nbodies = S.nb; // (copies of the template's: no global load on the row's path)
    const int Wp = P.Wp;
    const float hinv = 1.0f / P.h;
    float rhs = r.c * hinv;
    const float cfm = r.cfm * hinv;
    float s = 0.f;
    {
        const float *t = S.t1 + (size_t)(pc.b0 * 6) * BD + tid;
#pragma unroll
        for (int k = 0; k < 6; k++) s += r.J[k] * t[k * BD];
    }
    if (pc.b1 >= 0) {
        const float *t = S.t1 + (size_t)(pc.b1 * 6) * BD + tid;
#pragma unroll
        for (int k = 0; k < 6; k++) s += r.J[6 + k] * t[k * BD];
    }
    rhs -= s;
    float o[12];
    o[0] = pc.sm0 * r.J[0];
    o[1] = pc.sm0 * r.J[1];
    o[2] = pc.sm0 * r.J[2];
    sym_mul(o + 3, pc.S0, r.J + 3);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) sum += o[k] * o[k];
    if (pc.b1 >= 0) { // (a one-body row's second half is zero: same sum, same row)
        o[6] = pc.sm1 * r.J[6];
        o[7] = pc.sm1 * r.J[7];
        o[8] = pc.sm1 * r.J[8];
        sym_mul(o + 9, pc.S1, r.J + 9);
#pragma unroll
        for (int k = 6; k < 12; k++) sum += o[k] * o[k];
    } else {
#pragma unroll
        for (int k = 6; k < 12; k++) o[k] = 0.f;
    }
    const float Ad = P.sor_w / (sum + cfm);
    const float sa = sqrtf(Ad), rsa = 1.0f / sa;
    // bound word b: the row is limited to [-b,b] with b = hi/sqrt(Ad) (+inf: unbounded); kind bit 0 / 1
    // replaces the lower / upper limit by 0 ([0,+inf) and (-inf,0]); for a friction-dependent row (fr != 0)
    // b = mu*sqrt(Ad_n/Ad_f) is multiplied by the normal row's current lam' in the sweep
    // (ODE: hi = |hicopy * lambda[findex]|).
    float bound = r.hi * rsa;
    int kind = 0;
    if (r.fr) bound = r.hi * r.sn * rsa;
    else if (r.lo == 0.f && r.hi == INFINITY) bound = INFINITY, kind = 1;
    else if (r.lo == -INFINITY && r.hi == 0.f) bound = INFINITY, kind = 2;
    // meta: own slot | slot of the multiplier the bound follows << 8 (slot maxrows + 1 holds the constant 1 for
    // rows with a fixed bound) | body0 << 16 | body1 << 24 (one-body rows name the dummy body nb) | kind << 30
    const uint32_t meta = (uint32_t)slot | ((uint32_t)(r.fr ? r.fr - 1 : R + 1) << 8) | ((uint32_t)pc.b0 << 16) |
                          ((uint32_t)(pc.b1 < 0 ? nbodies : pc.b1) << 24) | ((uint32_t)kind << 30);
    // two 32-byte units: J0..J7 | J8..J11 rhs' cfm' bound meta  (J0..J5 = body-0 side, J6..J11 = body-1 side:
    // the SOR kernel gives each side of a world its own lane)
    stg256(ROW_UNIT(P.rows, slot, 0, Wp, w), make_float4(sa * o[0], sa * o[1], sa * o[2], sa * o[3]),
           make_float4(sa * o[4], sa * o[5], sa * o[6], sa * o[7]));
    stg256(ROW_UNIT(P.rows, slot, 1, Wp, w), make_float4(sa * o[8], sa * o[9], sa * o[10], sa * o[11]),
           make_float4(sa * rhs, cfm * Ad, bound, __uint_as_float(meta)));
    P.rowS[(size_t)slot * Wp + w] = sa;
    return sa;
}

template <bool ISLANDS> __global__ void __launch_bounds__(BD, 8) b2p_assemble_kernel(const StepParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevTemplate *__restrict__ T = P.T;
    const int nb = T->nb, nj = T->nj, maxc = T->maxc, rpc = T->rpc;
    Smem S;
    S.nb = nb;
    S.maxrows = T->maxrows;
    S.t1 = (float *)smem_raw;
    S.sw = S.t1 + (size_t)nb * 6 * BD;
    S.pq = S.sw + (size_t)nb * 6 * BD;
    const int tid = threadIdx.x;
    const int w = blockIdx.x * BD + tid;
    {
        uint32_t *sb = (uint32_t *)(S.pq + (size_t)nb * 7 * BD);
        uint32_t *sj = sb + (size_t)nb * (sizeof(DevBody) / 4);
        const uint32_t *gb = (const uint32_t *)T->bodies, *gj = (const uint32_t *)T->joints;
        for (int i = tid; i < nb * (int)(sizeof(DevBody) / 4); i += BD) sb[i] = __ldg(gb + i);
        for (int i = tid; i < nj * (int)(sizeof(DevJoint) / 4); i += BD) sj[i] = __ldg(gj + i);
        S.bodies = (const DevBody *)sb;
        S.joints = (const DevJoint *)sj;
        S.scratch = (uint8_t *)(sj + (size_t)nj * (sizeof(DevJoint) / 4));
        __syncthreads();
    }
    if (w >= P.W) return;
    const int Wp = P.Wp;
    const float h = P.h, hinv = 1.0f / P.h;

    // ---------------- phase 1: bodies ----------------
    for (int b = 0; b < nb; b++) {
        const DevBody &bc = S.bodies[b];
        BodyPose bp;
        load_pose(P, b, w, bp);
        {
            float *p = S.pq + (size_t)(b * 7) * BD + tid;
            p[0] = bp.pos[0], p[BD] = bp.pos[1], p[2 * BD] = bp.pos[2];
            p[3 * BD] = bp.q[0], p[4 * BD] = bp.q[1], p[5 * BD] = bp.q[2], p[6 * BD] = bp.q[3];
        }
        const int f = b * B2P_STATE_FIELDS;
        float avel[3] = {G(P.state, f + 10), G(P.state, f + 11), G(P.state, f + 12)};
        float facc[3] = {G(P.force, b * 6 + 0), G(P.force, b * 6 + 1), G(P.force, b * 6 + 2)};
        float tacc[3] = {G(P.force, b * 6 + 3), G(P.force, b * 6 + 4), G(P.force, b * 6 + 5)};
        float tmp[9], Iw[9];
        mul33_bt(tmp, bc.invI, bp.R);
        mul33(Iw, bp.R, tmp);
#pragma unroll
        for (int k = 0; k < 9; k++) G(P.invIw, b * 9 + k) = Iw[k];
        {
            // S_w = R sqrt(invI_body) R^T: the angular block of M^-1/2 in the world frame
            float Sw[9];
            mul33_bt(tmp, bc.sqrtInvI, bp.R);
            mul33(Sw, bp.R, tmp);
            S.sw[(size_t)(b * 6 + 0) * BD + tid] = Sw[0];
            S.sw[(size_t)(b * 6 + 1) * BD + tid] = Sw[1];
            S.sw[(size_t)(b * 6 + 2) * BD + tid] = Sw[2];
            S.sw[(size_t)(b * 6 + 3) * BD + tid] = Sw[4];
            S.sw[(size_t)(b * 6 + 4) * BD + tid] = Sw[5];
            S.sw[(size_t)(b * 6 + 5) * BD + tid] = Sw[8];
        }
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
            S.t1[(size_t)(b * 6 + k) * BD + tid] = facc[k];
            S.t1[(size_t)(b * 6 + 3 + k) * BD + tid] = tacc[k];
        }
    }

    // ---------------- phase 1a: limit motors powered into a stop add forces (rare) -------------
    if (T->has_limit_motors) {
        for (int j = 0; j < nj; j++) {
            const DevJoint &jc = S.joints[j];
            if (jc.type == 6 || jc.b0 < 0) continue;
            BodyPose p0, p1;
            load_pose_s(S, jc.b0, tid, p0);
            const bool has_b1 = jc.b1 >= 0;
            if (has_b1) load_pose_s(S, jc.b1, tid, p1);
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
                float *f0 = S.t1 + (size_t)(jc.b0 * 6) * BD + tid;
                float *f1 = has_b1 ? S.t1 + (size_t)(jc.b1 * 6) * BD + tid : nullptr;
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
        const DevBody &bc = S.bodies[b];
        const int f = b * B2P_STATE_FIELDS;
        float Iw[9], fe[6], t3[3];
        load_invIw(P, b, w, Iw);
#pragma unroll
        for (int k = 0; k < 6; k++) {
            fe[k] = S.t1[(size_t)(b * 6 + k) * BD + tid];
            G(P.force, b * 6 + k) = fe[k];
        }
        mul_R_v(t3, Iw, fe + 3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float lv = G(P.state, f + 7 + k), av = G(P.state, f + 10 + k);
            S.t1[(size_t)(b * 6 + k) * BD + tid] = fe[k] * bc.invMass + lv * hinv;
            S.t1[(size_t)(b * 6 + 3 + k) * BD + tid] = t3[k] + av * hinv;
        }
    }

    // ---------------- phase 2: constraint rows ----------------
    // Every constraint owns fixed slots (joint j: j*B2P_ROWS_PER_JOINT .., contact c: njslots + c*rpc ..), so the
    // rows can be produced in any order; phase 3 then lists the used slots in the order ODE sweeps them first.
    int nc = G(P.ccount, 0);
    if (nc > maxc) nc = maxc;
    const int njslots = T->njslots;
    // phase 3 reads back what phase 2 found out: the jrows byte of every joint and rows | dependent-row mask << 3
    // of every contact (bytes, this thread's column of the scratch area)
    uint8_t *const s_jr = S.scratch + tid, *const s_cr = S.scratch + (size_t)nj * BD + tid;

    // per-world commands of joint j
    auto joint_cmd = [&](int j, const DevJoint &jc) -> LimotCmd {
        LimotCmd cmd = {0.f, 0.f, 0.f, 0.f};
        if (jc.type != 6 && jc.type != 4) {
            cmd.vel1 = G(P.jcmd, j * 4 + 0);
            cmd.fmax1 = G(P.jcmd, j * 4 + 1);
            if (jc.type == 2 || jc.type == 5) {
                cmd.vel2 = G(P.jcmd, j * 4 + 2);
                cmd.fmax2 = G(P.jcmd, j * 4 + 3);
            }
        }
        return cmd;
    };

    // rows of joint j (all of a joint's rows are independent rows)
    auto emit_joint = [&](int j) {
        const DevJoint &jc = S.joints[j];
        int jr = 0;
        if (jc.b0 >= 0) {
            BodyPose p0, p1;
            PairCtx pc;
            pc.b0 = jc.b0;
            pc.b1 = jc.b1;
            load_pose_s(S, jc.b0, tid, p0);
            if (jc.b1 >= 0) load_pose_s(S, jc.b1, tid, p1);
            pair_ctx(pc, S, tid);
            const LimotCmd cmd = joint_cmd(j, jc);
            int slot = j * B2P_ROWS_PER_JOINT;
            auto emit = [&](const Row &r) { emit_row(P, S, w, tid, slot++, r, pc); };
            jr = joint_rows<true>(P, jc, p0, p1, cmd, 0, w, emit);
        }
        // rows of the joint | which limit / motor rows exist (b2p_rows.cuh); the integrate kernel follows it
        G(P.jrows, j) = (uint8_t)jr;
        s_jr[(size_t)j * BD] = (uint8_t)jr;
    };

    // ---- one contact (ODE contact.cpp getInfo1/getInfo2 for the mode bits Frame::addContact sets): the normal row,
    // then the friction (and rolling) rows in consecutive slots; a row whose bound follows the normal force
    // (findex = the normal row) is "dependent" and sweeps behind all independent rows of its island
    auto emit_contact = [&](int c) {
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
        load_pose_s(S, b0, tid, p0);
        if (b1 >= 0) load_pose_s(S, b1, tid, p1);
        pair_ctx(pc, S, tid);
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
        const int nslot = njslots + c * rpc;
        const float sn = emit_row(P, S, w, tid, nslot, r, pc);
        int row = 1, depmask = 0;
        float t1[3], t2[3];
        plane_space(n, t1, t2);
        const bool axisdep = (mode & 0x001) != 0;
        const bool fr1 = mu > 0.f;
        const float mu2e = axisdep ? mu2 : mu;
        const bool fr2 = axisdep ? (mu2 > 0.f) : (mu > 0.f);
        auto emit_fr = [&](bool dependent) {
            if (dependent) {
                r.fr = nslot + 1;
                r.sn = sn;
                depmask |= 1 << (row - 1);
            }
            emit_row(P, S, w, tid, nslot + row, r, pc);
            row++;
        };
        if (fr1) {
            row_init(r, P.cfm);
            r.J[0] = t1[0]; r.J[1] = t1[1]; r.J[2] = t1[2];
            cross3(r.J + 3, c1, t1);
            if (b1 >= 0) { r.J[6] = -t1[0]; r.J[7] = -t1[1]; r.J[8] = -t1[2]; cross3(r.J + 9, t1, c2); }
            r.lo = -mu;
            r.hi = mu;
            emit_fr((mode & 0x1000) != 0);
        }
        if (fr2) {
            row_init(r, P.cfm);
            r.J[0] = t2[0]; r.J[1] = t2[1]; r.J[2] = t2[2];
            cross3(r.J + 3, c1, t2);
            if (b1 >= 0) { r.J[6] = -t2[0]; r.J[7] = -t2[1]; r.J[8] = -t2[2]; cross3(r.J + 9, t2, c2); }
            r.lo = -mu2e;
            r.hi = mu2e;
            emit_fr((mode & 0x2000) != 0);
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
                    emit_fr((mode & bits[i]) != 0);
                }
            }
        }
        // rows | dependent-row mask << 3 | slot of the normal row << 8 | friction row 1 / 2 present << 24 / 25
        // (the integrate kernel rebuilds the contact's feedback from it)
        G(P.cmap, c) = (uint32_t)(row | (depmask << 3)) | ((uint32_t)nslot << 8) | ((uint32_t)(fr1 ? 1 : 0) << 24) |
                       ((uint32_t)(fr2 ? 1 : 0) << 25);
        s_cr[(size_t)c * BD] = (uint8_t)(row | (depmask << 3));
        if (ISLANDS) { // body ids for the island search of phase 3
            s_cr[(size_t)(maxc + c) * BD] = (uint8_t)b0;
            s_cr[(size_t)(2 * maxc + c) * BD] = (uint8_t)(b1 < 0 ? 0xff : b1);
        }
    };

    for (int j = 0; j < nj; j++) emit_joint(j);
    for (int c = 0; c < nc; c++) emit_contact(c);

    // ---------------- phase 3: the initial sweep order ----------------
    // ODE solves one island at a time; inside an island the rows with findex < 0 come first (in joint order) and
    // the friction-dependent rows last.  Every island therefore is a block [start, start + nh) (+) [.., start + m)
    // of sweep positions; the SOR kernels get the block table (isl) for ODE's per-island reshuffles and ord0, the
    // slot at every position.
    //   ISLANDS == false: the whole world is one block, joints in creation order, then the contact stream.
    //   ISLANDS == true:  ODE's dxProcessIslands order (util.cpp): bodies from the most recently created one,
    //                     depth-first through each body's joint list (most recently attached first: the step's
    //                     contacts in reverse stream order, then the joints in reverse creation order).
    int ntail = 0, nisl = 0; // rows of the world so far; islands (blocks) so far
    // the independent rows of one item at positions head.., counting its dependent rows
    auto place_heads = [&](int it, int &head, int &ndep) {
        if (it & 0x80) {
            const int c = it & 0x7f, cr = s_cr[(size_t)c * BD], rows = cr & 7, dep = cr >> 3;
            const int base = njslots + c * rpc;
            for (int k = 0; k < rows; k++) {
                if (k > 0 && (dep >> (k - 1) & 1)) ndep++;
                else G(P.ord0, head++) = (uint8_t)(base + k);
            }
        } else {
            const int rows = B2P_JR_ROWS(s_jr[(size_t)it * BD]);
            for (int k = 0; k < rows; k++) G(P.ord0, head++) = (uint8_t)(it * B2P_ROWS_PER_JOINT + k);
        }
    };
    auto place_tails = [&](int c, int &tail) {
        const int cr = s_cr[(size_t)c * BD], rows = cr & 7, dep = cr >> 3;
        const int base = njslots + c * rpc;
        for (int k = 1; k < rows; k++)
            if (dep >> (k - 1) & 1) G(P.ord0, tail++) = (uint8_t)(base + k);
    };
    if (!ISLANDS) {
        int head = 0, ndep = 0;
        for (int j = 0; j < nj; j++) place_heads(j, head, ndep);
        for (int c = 0; c < nc; c++) place_heads(0x80 | c, head, ndep);
        int tail = head;
        if (ndep) {
            for (int c = 0; c < nc; c++) place_tails(c, tail);
        }
        ntail = tail;
        G(P.isl, 0) = (uint32_t)head << 8 | (uint32_t)ntail << 16; // start 0 | independent rows << 8 | rows << 16
        nisl = 1;
    } else {
        // scratch of this thread (bytes): the island's items in discovery order (item = joint index, or 0x80 |
        // contact index) and the depth-first stack, behind the per-item bytes of phase 2
        uint8_t *const seq = S.scratch + (size_t)(nj + 3 * maxc) * BD + tid, *const stack = seq + (size_t)(nj + maxc) * BD;
        const uint8_t *const cb0 = s_cr + (size_t)maxc * BD, *const cb1 = cb0 + (size_t)maxc * BD;
        uint32_t visited = 0;
        uint64_t jtag = 0, ctag = 0;
        for (int bb = nb - 1; bb >= 0; bb--) {
            if (visited >> bb & 1u) continue;
            visited |= 1u << bb;
            int nseq = 0, sp = 0, b = bb;
            for (;;) {
                for (int c = nc - 1; c >= 0; c--) { // this step's contacts: attached last, listed first
                    if (ctag >> c & 1ull) continue;
                    const int c0 = cb0[(size_t)c * BD], c1 = cb1[(size_t)c * BD];
                    if (c0 != b && c1 != b) continue;
                    ctag |= 1ull << c;
                    seq[(size_t)nseq++ * BD] = (uint8_t)(0x80 | c);
                    const int other = c0 == b ? c1 : c0;
                    if (other != 0xff && !(visited >> other & 1u)) {
                        visited |= 1u << other;
                        stack[(size_t)sp++ * BD] = (uint8_t)other;
                    }
                }
                for (int j = nj - 1; j >= 0; j--) {
                    const DevJoint &jc = S.joints[j];
                    if ((jtag >> j & 1ull) || (jc.b0 != b && jc.b1 != b)) continue;
                    jtag |= 1ull << j;
                    seq[(size_t)nseq++ * BD] = (uint8_t)j;
                    const int other = jc.b0 == b ? jc.b1 : jc.b0;
                    if (other >= 0 && !(visited >> other & 1u)) {
                        visited |= 1u << other;
                        stack[(size_t)sp++ * BD] = (uint8_t)other;
                    }
                }
                if (sp == 0) break;
                b = stack[(size_t)--sp * BD];
            }
            if (nseq == 0) continue; // a free body: an island without rows
            // the island's independent rows in item order, then the friction-dependent rows of its contacts
            const int start = ntail;
            int head = start, ndep = 0;
            for (int i = 0; i < nseq; i++) place_heads(seq[(size_t)i * BD], head, ndep);
            const int nh = head - start;
            int tail = head;
            if (ndep) {
                for (int i = 0; i < nseq; i++) {
                    const int it = seq[(size_t)i * BD];
                    if (it & 0x80) place_tails(it & 0x7f, tail);
                }
            }
            ntail = tail;
            if (tail > start) G(P.isl, nisl++) = (uint32_t)start | (uint32_t)nh << 8 | (uint32_t)(tail - start) << 16;
        }
    }
    // rows of the step | islands << 16
    G(P.nrows, 0) = ntail | (nisl << 16);
    // high-water mark of rows per world: the host sizes the SOR kernel's on-chip row storage with it
    const int wmax = __reduce_max_sync(__activemask(), ntail);
    if ((tid & 31) == 0 && wmax > *(volatile int *)P.hw_rows) atomicMax(P.hw_rows, wmax);
}

} // namespace

extern "C" size_t b2p_assemble_smem_bytes(int nb, int nj, int maxc, int islands)
{
    return sizeof(float) * (size_t)nb * 19 * BD + (size_t)nb * sizeof(DevBody) + (size_t)nj * sizeof(DevJoint) +
           (size_t)(nj + 3 * maxc + (islands ? nj + maxc + nb : 0)) * BD;
}

extern "C" cudaError_t b2p_launch_assemble(const StepParams *P, int nb, int nj, int maxc, cudaStream_t stream)
{
    static size_t configured_dev[B2P_MAX_DEVICES][2] = {{0, 0}}; // cudaFuncSetAttribute is per device
    static std::mutex configured_mutex; // arenas of several threads / devices may launch at the same time
    std::lock_guard<std::mutex> configured_lock(configured_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    const int isl = P->order_mode != 0;
    size_t &configured = configured_dev[dev % B2P_MAX_DEVICES][isl];
    const size_t smem = b2p_assemble_smem_bytes(nb, nj, maxc, isl);
    auto kern = isl ? b2p_assemble_kernel<true> : b2p_assemble_kernel<false>;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int grid = (P->W + BD - 1) / BD;
    kern<<<grid, BD, smem, stream>>>(*P);
    return cudaGetLastError();
}

// b2p_integrate.cu -- kernel 3 of the step: joint / contact feedback, velocity update, dxStepBody, joint
// observations and the reference's post-step bookkeeping for every world (sm_100a).
//
// One thread per world, world-fastest SoA accesses.  It finishes dWorldQuickStep as the reference calls
// it at src/WorldPhysics.cpp:365 (ODE 0.16 quickstep.cpp stage 6: v += h*(invM*fe + invM J^T lambda);
// util.cpp dxStepBody with the angular-speed cap and the finite-rotation quaternion update) and then does what
// WorldPhysics::stepTheWorld does after the ODE call (src/WorldPhysics.cpp:372-397), for ALL worlds:
//   dJointFeedback f1 t1 f2 t2 and the patched `lambda`      (src/joints/Joint.cpp:268,900)
//   the contact joints' feedback f1 and its sum per frame     (src/objects/Frame.cpp:159-171)
//   dJointGet*Angle / *Rate / SliderPosition                  (src/joints/Joint.cpp:355-415,1072-1102)
//   Joint::update: motor torque, axis torque, joint load      (src/joints/Joint.cpp:859-1019)
//   Joint::applyJointFriction: force / torque for the NEXT step (src/joints/Joint.cpp:1031-1063)
//   Frame::update: finite-difference accelerations, damping   (src/objects/Frame.cpp:899-943)
//
// Feedback without a second pass over the rows: lambda * J is accumulated while the joint's Jacobian is rebuilt
// from the pre-step poses by the same code that assembled it (b2p_rows.cuh, FULL = false); a contact's f1 is
// lambda_n n + lambda_1 t1 + lambda_2 t2 with the tangents of dPlaneSpace(n).  The SOR kernel leaves
// fch = M^-1/2 J^T lambda and lam' = lambda / sqrt(Ad) (see b2p_assemble.cu), so invM J^T lambda = M^-1/2 fch and
// lambda = lam' * sqrt(Ad) (rowS).
#include <mutex>

#include "b2p_rows.cuh"

namespace {

constexpr int BD = 64;

// `stash`: the CTA has shared memory for [nb*3 + nj*7][BD] floats -- the contact-force sum per body and the part of
// the joint feedback Joint::update reads back -- so that neither goes through global memory and back
__global__ void __launch_bounds__(BD, 7) b2p_integrate_kernel(const StepParams P, const int stash)
{
    extern __shared__ float s_stash[];
    const DevTemplate *__restrict__ T = P.T;
    const int nb = T->nb, nj = T->nj, maxc = T->maxc;
    const int tid = threadIdx.x;
    // the scene template's bodies and joints in shared memory (read warp-uniformly all over the kernel: a broadcast
    // LDS instead of a global load on the thread's dependent chain), then the stash
    uint32_t *const sb = (uint32_t *)s_stash;
    uint32_t *const sj = sb + (size_t)nb * (sizeof(DevBody) / 4);
    {
        const uint32_t *gb = (const uint32_t *)T->bodies, *gj = (const uint32_t *)T->joints;
        for (int i = tid; i < nb * (int)(sizeof(DevBody) / 4); i += BD) sb[i] = __ldg(gb + i);
        for (int i = tid; i < nj * (int)(sizeof(DevJoint) / 4); i += BD) sj[i] = __ldg(gj + i);
        __syncthreads();
    }
    const DevBody *const s_bodies = (const DevBody *)sb;
    const DevJoint *const s_joints = (const DevJoint *)sj;
    float *const s_st = (float *)(sj + (size_t)nj * (sizeof(DevJoint) / 4));
    float *const s_cf = s_st + tid;                        // [nb*3][BD] contact force on body b
    float *const s_jf = s_st + (size_t)nb * 3 * BD + tid;  // [nj*7][BD] f t (of the reference's body1), lambda
    const int w = blockIdx.x * BD + tid;
    if (w >= P.W) return;
    const int Wp = P.Wp;
    const float h = P.h;

    // ---------------- joint feedback from the pre-step poses: sum of lambda * J over the joint's rows --------
    // (software-pipelined like the body loop below: the poses, row counts and multipliers of joint j + 1 are
    // requested before joint j's Jacobian is rebuilt)
    {
        struct JointIn {
            float pq0[7], pq1[7]; // raw position + quaternion: nothing here may wait for the loads
            int jr, slot;
            float lam[B2P_ROWS_PER_JOINT];
        };
        auto load_joint = [&](int j, JointIn &x) {
            const DevJoint &jc = s_joints[j];
            x.jr = G(P.jrows, j);
            x.slot = jc.slot_base; // joint j owns slots j * B2P_ROWS_PER_JOINT ..
#pragma unroll
            for (int k = 0; k < 7; k++) {
                x.pq0[k] = jc.b0 >= 0 ? G(P.state, jc.b0 * B2P_STATE_FIELDS + k) : 0.f;
                x.pq1[k] = jc.b1 >= 0 ? G(P.state, jc.b1 * B2P_STATE_FIELDS + k) : 0.f;
            }
            // all of the joint's slots, whether used this step or not (the row count is still on its way: no load
            // here may depend on another one); the unused ones are masked when the count is there
#pragma unroll
            for (int k = 0; k < B2P_ROWS_PER_JOINT; k++)
                x.lam[k] = G(P.lambda, x.slot + k) * G(P.rowS, x.slot + k); // lambda = lam' sqrt(Ad)
        };
        JointIn nxt;
        if (nj > 0) load_joint(0, nxt);
        for (int j = 0; j < nj; j++) {
            const DevJoint &jc = s_joints[j];
            const JointIn cur = nxt;
            if (j + 1 < nj) load_joint(j + 1, nxt);
            const int jr = cur.jr;
            const int mrows = B2P_JR_ROWS(jr);
            float acc[12];
#pragma unroll
            for (int q = 0; q < 12; q++) acc[q] = 0.f;
            float flam = 0.f;
            if (jc.b0 >= 0 && mrows > 0) {
                // `lambda` of the patched dJointFeedback: the multiplier of the joint's first limit / motor row
                const int nl = ((jr & B2P_JR_L1) ? 1 : 0) + ((jr & B2P_JR_L2) ? 1 : 0);
                const int first_limot = nl ? mrows - nl : -1;
                int k = 0;
                auto emit = [&](const Row &r) {
                    // (k is a compile-time constant on every path through joint_rows once it is inlined)
                    float lam = 0.f;
#pragma unroll
                    for (int i = 0; i < B2P_ROWS_PER_JOINT; i++)
                        if (i == k && i < mrows) lam = cur.lam[i];
#pragma unroll
                    for (int q = 0; q < 12; q++) acc[q] += r.J[q] * lam;
                    if (k == first_limot) flam = lam;
                    k++;
                };
                const LimotCmd cmd = {0.f, 0.f, 0.f, 0.f};
                BodyPose p0, p1;
#pragma unroll
                for (int i = 0; i < 3; i++) p0.pos[i] = cur.pq0[i], p1.pos[i] = cur.pq1[i];
#pragma unroll
                for (int i = 0; i < 4; i++) p0.q[i] = cur.pq0[3 + i], p1.q[i] = cur.pq1[3 + i];
                q_to_R(p0.q, p0.R);
                if (jc.b1 >= 0) q_to_R(p1.q, p1.R);
                joint_rows<false>(P, jc, p0, p1, cmd, jr, w, emit);
            }
#pragma unroll
            for (int q = 0; q < 12; q++) G(P.jfb, j * 13 + q) = acc[q];
            G(P.jfb, j * 13 + 12) = flam;
            if (stash && P.jupd) {
                const int o = jc.reverse ? 6 : 0;
#pragma unroll
                for (int q = 0; q < 6; q++) s_jf[(size_t)(j * 7 + q) * BD] = o ? acc[6 + q] : acc[q];
                s_jf[(size_t)(j * 7 + 6) * BD] = flam;
            }
        }
    }

    // ---------------- contact feedback f1, and its sum per body (Frame::computeContactForce) ----------------
    const bool frame_out = P.bodyx != nullptr;
    if (frame_out) {
        for (int k = 0; k < nb * 3; k++) {
            if (stash) s_cf[(size_t)k * BD] = 0.f;
            else G(P.bodyx, (k / 3) * 15 + 12 + k % 3) = 0.f;
        }
    }
    {
        int nc = G(P.ccount, 0);
        if (nc > maxc) nc = maxc;
        const int njslots = T->njslots, rpc = T->rpc;
        for (int c = 0; c < nc; c++) {
            // the contact's rows sit in consecutive slots in emission order: normal, friction 1, friction 2,
            // (rolling).  Everything is requested at once -- the two slots behind the normal row whether they hold
            // friction rows or not -- and sorted out when the contact's row map is there.
            const int s0 = njslots + c * rpc;
            const uint32_t cm = G(P.cmap, c);
            const int b0 = __float_as_int(G(P.crec, c * B2P_CONTACT_FIELDS + CR_IDS)) & 0xff;
            const float n[3] = {G(P.crec, c * B2P_CONTACT_FIELDS + CR_NX), G(P.crec, c * B2P_CONTACT_FIELDS + CR_NY),
                                G(P.crec, c * B2P_CONTACT_FIELDS + CR_NZ)};
            const float ln = G(P.lambda, s0) * G(P.rowS, s0);
            const float la = G(P.lambda, s0 + 1) * G(P.rowS, s0 + 1), lb = G(P.lambda, s0 + 2) * G(P.rowS, s0 + 2);
            const bool fr1 = (cm >> 24) & 1, fr2 = (cm >> 25) & 1;
            float t1[3], t2[3];
            plane_space(n, t1, t2);
            float f[3] = {n[0] * ln, n[1] * ln, n[2] * ln};
            if (fr1) {
#pragma unroll
                for (int i = 0; i < 3; i++) f[i] += t1[i] * la;
            }
            if (fr2) {
                const float l2 = fr1 ? lb : la;
#pragma unroll
                for (int i = 0; i < 3; i++) f[i] += t2[i] * l2;
            }
            // (rolling rows have no linear part)
            G(P.cfb, c * 3 + 0) = f[0];
            G(P.cfb, c * 3 + 1) = f[1];
            G(P.cfb, c * 3 + 2) = f[2];
            if (frame_out) {
                if (stash) {
#pragma unroll
                    for (int i = 0; i < 3; i++) s_cf[(size_t)(b0 * 3 + i) * BD] += f[i];
                } else {
                    G(P.bodyx, b0 * 15 + 12) += f[0];
                    G(P.bodyx, b0 * 15 + 13) += f[1];
                    G(P.bodyx, b0 * 15 + 14) += f[2];
                }
            }
        }
    }

    // ---------------- velocity update + dxStepBody ----------------
    // Software-pipelined over the bodies: the 25 input words of body b + 1 are requested before body b is worked
    // on, so that their DRAM latency (the kernel has one thread per world, i.e. few warps per SM to hide it
    // otherwise) runs under ~300 instructions of arithmetic.
    struct BodyIn {
        float pos[3], q[4], lv[3], av[3], fe[6], fch[6], last[6];
    };
    auto load_body = [&](int b, BodyIn &x) {
        const int f = b * B2P_STATE_FIELDS;
        x.pos[0] = G(P.state, f + 0); x.pos[1] = G(P.state, f + 1); x.pos[2] = G(P.state, f + 2);
        x.q[0] = G(P.state, f + 3); x.q[1] = G(P.state, f + 4); x.q[2] = G(P.state, f + 5); x.q[3] = G(P.state, f + 6);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            x.lv[k] = G(P.state, f + 7 + k);
            x.av[k] = G(P.state, f + 10 + k);
        }
#pragma unroll
        for (int k = 0; k < 6; k++) {
            x.fe[k] = G(P.force, b * 6 + k);
            x.fch[k] = G(P.fcacc, b * 6 + k);
            x.last[k] = frame_out ? G(P.bodyx, b * 15 + k) : 0.f; // Frame::update's last_l_vel / last_a_vel
        }
    };
    BodyIn nxt;
    if (nb > 0) load_body(0, nxt);
    for (int b = 0; b < nb; b++) {
        const DevBody &bc = s_bodies[b];
        const int f = b * B2P_STATE_FIELDS;
        const BodyIn cur = nxt;
        if (b + 1 < nb) load_body(b + 1, nxt);
        float pos[3], q[4], lv[3], av[3], fe[6], R[9], t3[3], u3[3];
#pragma unroll
        for (int k = 0; k < 3; k++) pos[k] = cur.pos[k], lv[k] = cur.lv[k], av[k] = cur.av[k];
#pragma unroll
        for (int k = 0; k < 4; k++) q[k] = cur.q[k];
#pragma unroll
        for (int k = 0; k < 6; k++) fe[k] = cur.fe[k];
        const float *fch = cur.fch;
        q_to_R(q, R);
        // v += h * M^-1/2 fch, with the angular block R sqrt(invI_body) R^T applied as three matrix-vector products
        mul_Rt_v(t3, R, fch + 3);
        mul_R_v(u3, bc.sqrtInvI, t3);
        mul_R_v(t3, R, u3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lv[k] += h * (bc.sqrtInvMass * fch[k]);
            av[k] += h * t3[k];
        }
        // v += h * invM fe
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lv[k] += h * bc.invMass * fe[k];
            fe[3 + k] *= h;
        }
        mul_Rt_v(t3, R, fe + 3);
        mul_R_v(u3, bc.invI, t3);
        mul_R_v(t3, R, u3);
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
        if (frame_out) {
            // Frame::update (Frame.cpp:899-921): accelerations by finite differences of the step's (not yet damped)
            // velocities; the damping itself follows the joints' post-step work below, as in stepTheWorld
            const float hinv = 1.0f / h;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                G(P.bodyx, b * 15 + 6 + k) = (lv[k] - cur.last[k]) * hinv;
                G(P.bodyx, b * 15 + 9 + k) = (av[k] - cur.last[3 + k]) * hinv;
                G(P.bodyx, b * 15 + k) = lv[k];
                G(P.bodyx, b * 15 + 3 + k) = av[k];
                if (stash) G(P.bodyx, b * 15 + 12 + k) = s_cf[(size_t)(b * 3 + k) * BD];
            }
        }
    }

    // ---------------- joint observations, Joint::update and applyJointFriction from the new state -----------
    for (int j = 0; j < nj; j++) {
        const DevJoint &jc = s_joints[j];
        float a1 = 0.f, r1 = 0.f, a2 = 0.f, r2 = 0.f;
        float axis[3] = {0.f, 0.f, 0.f};
        BodyPose p0, p1;
        const bool has_b1 = jc.b1 >= 0;
        if (jc.b0 >= 0) {
            load_pose(P, jc.b0, w, p0);
            if (has_b1) load_pose(P, jc.b1, w, p1);
        }
        if (jc.b0 >= 0 && jc.type != 6 && jc.type != 4) {
            float lv0[3], av0[3], lv1[3] = {0.f, 0.f, 0.f}, av1[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                lv0[k] = G(P.state, jc.b0 * B2P_STATE_FIELDS + 7 + k);
                av0[k] = G(P.state, jc.b0 * B2P_STATE_FIELDS + 10 + k);
            }
            if (has_b1) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    lv1[k] = G(P.state, jc.b1 * B2P_STATE_FIELDS + 7 + k);
                    av1[k] = G(P.state, jc.b1 * B2P_STATE_FIELDS + 10 + k);
                }
            }
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
            } else if (jc.type == 5) { // universal: dJointGetUniversalAngle1/2 and their rates
                float ax1[3], ax2[3], ang1, ang2;
                universal_axes(jc, p0, p1, has_b1, ax1, ax2);
                universal_angles(jc, p0, p1, has_b1, ang1, ang2);
                const float ra = dot3(ax1, av0) - dot3(ax1, av1), rb = dot3(ax2, av0) - dot3(ax2, av1);
                a1 = jc.reverse ? -ang2 : ang1;
                a2 = jc.reverse ? -ang1 : ang2;
                r1 = jc.reverse ? rb : ra;
                r2 = jc.reverse ? ra : rb;
            }
        }
        G(P.jobs, j * 4 + 0) = a1;
        G(P.jobs, j * 4 + 1) = r1;
        G(P.jobs, j * 4 + 2) = a2;
        G(P.jobs, j * 4 + 3) = r2;

        // Joint::update (Joint.cpp:859-1019): motor torque = -feedback.lambda; axis torque and joint load of
        // hinge / hinge2 / universal joints from the feedback forces and the anchor geometry of the new poses.
        // The reference's "body1" is the first body handed to dJointAttach: b0 unless the joint is reversed.
        if (P.jupd) {
            float at[3] = {0.f, 0.f, 0.f}, jl[3] = {0.f, 0.f, 0.f};
            if (jc.b0 >= 0 && (jc.type == 1 || jc.type == 2 || jc.type == 5)) {
                float anchor[3], v1[3], normal[3], load[3], tmp1[3];
                mul_R_v(anchor, p0.R, jc.anchor1);
#pragma unroll
                for (int i = 0; i < 3; i++) v1[i] = -anchor[i]; // body position - (body position + R anchor1)
                float fa[3], ta[3]; // feedback force / torque on the reference's body1 (f1 t1, reversed: f2 t2)
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    fa[i] = stash ? s_jf[(size_t)(j * 7 + i) * BD] : G(P.jfb, j * 13 + (jc.reverse ? 6 : 0) + i);
                    ta[i] = stash ? s_jf[(size_t)(j * 7 + 3 + i) * BD] : G(P.jfb, j * 13 + (jc.reverse ? 9 : 3) + i);
                }
                if (!jc.reverse) {
                    const float *f1 = fa, *t1 = ta;
                    cross3(normal, axis, v1);
                    float d = dot3(normal, f1);
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        normal[i] *= d;
                        load[i] = f1[i] - normal[i];
                    }
                    cross3(at, v1, normal);
                    cross3(jl, v1, load);
                    d = dot3(axis, t1);
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        tmp1[i] = axis[i] * d;
                        at[i] += tmp1[i];
                        jl[i] += t1[i] - tmp1[i];
                    }
                } else {
                    const float *f2 = fa, *t2 = ta;
                    float d = dot3(f2, v1) / dot3(v1, v1);
#pragma unroll
                    for (int i = 0; i < 3; i++) tmp1[i] = f2[i] - v1[i] * d;
                    const float torque = sqrtf(dot3(tmp1, tmp1)) / sqrtf(dot3(v1, v1));
                    cross3(normal, v1, tmp1);
                    const float sc = torque / sqrtf(dot3(normal, normal));
#pragma unroll
                    for (int i = 0; i < 3; i++) normal[i] *= sc;
                    d = dot3(normal, axis);
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        at[i] = axis[i] * d;
                        jl[i] = normal[i] - at[i];
                    }
                    d = dot3(t2, axis);
#pragma unroll
                    for (int i = 0; i < 3; i++) jl[i] += t2[i] - axis[i] * d;
                }
            }
            G(P.jupd, j * 7 + 0) = -(stash ? s_jf[(size_t)(j * 7 + 6) * BD] : G(P.jfb, j * 13 + 12));
#pragma unroll
            for (int i = 0; i < 3; i++) {
                G(P.jupd, j * 7 + 1 + i) = at[i];
                G(P.jupd, j * 7 + 4 + i) = jl[i];
            }
        }

        // Joint::applyJointFriction (Joint.cpp:1031-1063): an external force / torque on both bodies for the NEXT
        // step; dampingTorque = getVelocityI() * -frictionCoefficient with the reference's sign rules
        if (jc.friction > 0.f && jc.b0 >= 0) {
            float fax[3] = {axis[0], axis[1], axis[2]}; // getAxisI: world axis 1; none for ball / fixed
            if (jc.type == 3 && jc.reverse && !has_b1) {
                fax[0] = -fax[0], fax[1] = -fax[1], fax[2] = -fax[2];
            }
            const float vel = (jc.type == 1 || jc.type == 3) ? -r1 : ((jc.type == 2 || jc.type == 5) ? r1 : 0.f);
            const float dt_ = vel * -jc.friction;
            const float nrm = sqrtf(dot3(fax, fax));
            float u[3] = {0.f, 0.f, 0.f};
            if (nrm > 0.f) {
#pragma unroll
                for (int i = 0; i < 3; i++) u[i] = fax[i] / nrm;
            }
            float sn, cs;
            sincosf(dt_, &sn, &cs);
#pragma unroll
            for (int side = 0; side < 2; side++) {
                const int body = side == 0 ? jc.b0 : jc.b1;
                if (body < 0) continue;
                const BodyPose &bp = side == 0 ? p0 : p1;
                float pos[3], ux[3];
#pragma unroll
                for (int i = 0; i < 3; i++) pos[i] = bp.pos[i] - fax[i];
                cross3(ux, u, pos);
                const float ud = dot3(u, pos);
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const float rot = pos[i] * cs + ux[i] * sn + u[i] * ud * (1.f - cs);
                    G(P.force, body * 6 + i) += rot - pos[i];
                    G(P.force, body * 6 + 3 + i) += u[i] * dt_;
                }
            }
        }
    }

    // ---------------- Frame::update, second half (Frame.cpp:922-941): multiplicative damping ---------------
    for (int b = 0; b < nb; b++) {
        const DevBody &bc = s_bodies[b];
        const bool damp_l = bc.lin_damp < 1.f, damp_a = bc.ang_damp < 1.f;
        if (!damp_l && !damp_a) continue;
        const int f = b * B2P_STATE_FIELDS;
        if (damp_l) {
#pragma unroll
            for (int k = 0; k < 3; k++) G(P.state, f + 7 + k) = G(P.state, f + 7 + k) * bc.lin_damp;
        }
        if (damp_a) {
#pragma unroll
            for (int k = 0; k < 3; k++) G(P.state, f + 10 + k) = G(P.state, f + 10 + k) * bc.ang_damp;
        }
    }
}

// Diagnostic: the LCP residual of the last solve, per world (north-star correctness bar "constraint residual after
// N iterations").  w = J invM J^T lambda + cfm lambda - rhs on the unscaled system, projected on the active
// bounds exactly as the oracle does (oracle/orc_quickstep.c, "residual diagnostic").  In the scaled
// coordinates of b2p_assemble.cu:  w_i = -(rhs' - cfm' lam' - Jh.fch) / sqrt(Ad_i).
__global__ void __launch_bounds__(BD, 8) b2p_residual_kernel(const StepParams P, float *__restrict__ out)
{
    const DevTemplate *__restrict__ T = P.T;
    const int w = blockIdx.x * BD + threadIdx.x;
    if (w >= P.W) return;
    const int Wp = P.Wp, R = T->maxrows, nb = T->nb;
    const int m = G(P.nrows, 0) & 0xffff;
    float maxres = 0.f;
    for (int p = 0; p < m; p++) {
        const int s = G(P.ord0, p); // the used slots, in the initial sweep order
        const F8 u0 = ldg256(ROW_UNIT(P.rows, s, 0, Wp, w)), u1 = ldg256(ROW_UNIT(P.rows, s, 1, Wp, w));
        const float jv[12] = {u0.lo.x, u0.lo.y, u0.lo.z, u0.lo.w, u0.hi.x, u0.hi.y,
                              u0.hi.z, u0.hi.w, u1.lo.x, u1.lo.y, u1.lo.z, u1.lo.w};
        const float rhs = u1.hi.x, cfm = u1.hi.y, bound = u1.hi.z;
        const uint32_t mt = (uint32_t)__float_as_int(u1.hi.w);
        const int fi = (mt >> 8) & 0xff, b0 = (mt >> 16) & 0x3f, b1 = (mt >> 24) & 0x3f;
        const float lam = G(P.lambda, s);
        const float lam_n = fi < R ? G(P.lambda, fi) : 1.f;
        float dotv = 0.f;
#pragma unroll
        for (int k = 0; k < 6; k++) dotv += jv[k] * G(P.fcacc, b0 * 6 + k);
        if (b1 < nb) {
#pragma unroll
            for (int k = 0; k < 6; k++) dotv += jv[6 + k] * G(P.fcacc, b1 * 6 + k);
        }
        const float sa = G(P.rowS, s);
        const float wi = -((rhs - lam * cfm) - dotv) / sa; // unscaled residual of row s
        const float bnd = fabsf(bound * lam_n);
        const float hi_act = (mt & 0x80000000u) ? 0.f : bnd;
        const float lo_act = (mt & 0x40000000u) ? 0.f : -bnd;
        float res;
        if (lam <= lo_act && lo_act > -INFINITY) res = wi < 0.f ? -wi : 0.f;
        else if (lam >= hi_act && hi_act < INFINITY) res = wi > 0.f ? wi : 0.f;
        else res = fabsf(wi);
        if (lo_act == hi_act) res = 0.f;
        maxres = fmaxf(maxres, res);
    }
    out[w] = maxres;
}

} // namespace

extern "C" cudaError_t b2p_launch_integrate(const StepParams *P, int nb, int nj, cudaStream_t stream)
{
    const int grid = (P->W + BD - 1) / BD;
    // the stash rides in the default 48 KB of dynamic shared memory; larger scenes go through global memory
    const size_t tmpl = (size_t)nb * sizeof(DevBody) + (size_t)nj * sizeof(DevJoint); // (<= 32 x 192 + 48 x 432 bytes)
    const size_t smem = sizeof(float) * (size_t)(nb * 3 + nj * 7) * BD;
    const int stash = smem <= 24 * 1024 && !(P->tune & 4);
    static bool configured_dev[B2P_MAX_DEVICES] = {false};
    static std::mutex configured_mutex;
    {
        std::lock_guard<std::mutex> lock(configured_mutex);
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured_dev[dev % B2P_MAX_DEVICES]) { // template copy of the largest scene + stash: above the default 48 KB
            cudaError_t e = cudaFuncSetAttribute(b2p_integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (e != cudaSuccess) return e;
            configured_dev[dev % B2P_MAX_DEVICES] = true;
        }
    }
    b2p_integrate_kernel<<<grid, BD, tmpl + (stash ? smem : 0), stream>>>(*P, stash);
    return cudaGetLastError();
}

extern "C" cudaError_t b2p_launch_residual(const StepParams *P, float *out, cudaStream_t stream)
{
    const int grid = (P->W + BD - 1) / BD;
    b2p_residual_kernel<<<grid, BD, 0, stream>>>(*P, out);
    return cudaGetLastError();
}

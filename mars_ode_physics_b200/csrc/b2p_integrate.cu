// b2p_integrate.cu -- kernel 3 of the step: velocity update, dxStepBody, joint / contact feedback and
// joint observations (sm_100a).
//
// One thread per world, world-fastest SoA accesses.  It finishes dWorldQuickStep as the reference calls
// it at src/WorldPhysics.cpp:365 (ODE 0.16 quickstep.cpp stage 6: v += h*(invM*fe + invM J^T lambda);
// util.cpp dxStepBody with the angular-speed cap and the finite-rotation quaternion update) and produces
// what the reference reads right after the step:
//   dJointFeedback f1 t1 f2 t2 and the patched `lambda`  (src/joints/Joint.cpp:859-1019, :900)
//   the contact joint's feedback f1                       (src/objects/Frame.cpp:159-171)
//   dJointGet*Angle / *Rate / SliderPosition              (src/joints/Joint.cpp:355-415,1072-1102)
//
// The SOR kernel leaves fch = M^-1/2 J^T lambda and lam' = lambda/sqrt(Ad) (see b2p_assemble.cu), so
//   invM J^T lambda = M^-1/2 fch,   lambda * J = lam' * Jh * M^1/2   (the sqrt(Ad) factors cancel).
#include "b2p_kern.cuh"

namespace {

constexpr int BD = 64;

__device__ __forceinline__ void sym6_from(float *s6, const float *R, const float *Bsym, float *tmp)
{
    float M[9];
    mul33_bt(tmp, Bsym, R);
    mul33(M, R, tmp);
    s6[0] = M[0];
    s6[1] = M[1];
    s6[2] = M[2];
    s6[3] = M[4];
    s6[4] = M[5];
    s6[5] = M[8];
}

__device__ __forceinline__ void sym_mul(float *o, const float *S, const float *v)
{
    const float x = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
    const float y = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
    const float z = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
    o[0] = x;
    o[1] = y;
    o[2] = z;
}

__global__ void __launch_bounds__(BD, 8) b2p_integrate_kernel(const StepParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevTemplate *__restrict__ T = P.T;
    const int nb = T->nb, nj = T->nj, maxc = T->maxc;
    float *const s_si = (float *)smem_raw; // [nb*6][BD] symmetric sqrt(I_world) of the pre-step pose
    const int tid = threadIdx.x;
    const int w = blockIdx.x * BD + tid;
    if (w >= P.W) return;
    const int Wp = P.Wp;
    const float h = P.h;

    // ---------------- velocity update + dxStepBody ----------------
    for (int b = 0; b < nb; b++) {
        const DevBody &bc = T->bodies[b];
        const int f = b * B2P_STATE_FIELDS;
        float pos[3], q[4], lv[3], av[3], fe[6], R[9], tmp[9], S6[6], fch[6], t3[3];
        pos[0] = G(P.state, f + 0); pos[1] = G(P.state, f + 1); pos[2] = G(P.state, f + 2);
        q[0] = G(P.state, f + 3); q[1] = G(P.state, f + 4); q[2] = G(P.state, f + 5); q[3] = G(P.state, f + 6);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lv[k] = G(P.state, f + 7 + k);
            av[k] = G(P.state, f + 10 + k);
        }
#pragma unroll
        for (int k = 0; k < 6; k++) {
            fe[k] = G(P.force, b * 6 + k);
            fch[k] = G(P.fcacc, b * 6 + k);
        }
        q_to_R(q, R);
        sym6_from(S6, R, bc.sqrtI, tmp);
#pragma unroll
        for (int k = 0; k < 6; k++) s_si[(size_t)(b * 6 + k) * BD + tid] = S6[k];
        // v += h * M^-1/2 fch
        sym6_from(S6, R, bc.sqrtInvI, tmp);
        sym_mul(t3, S6, fch + 3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lv[k] += h * (bc.sqrtInvMass * fch[k]);
            av[k] += h * t3[k];
        }
        // v += h * invM fe
        sym6_from(S6, R, bc.invI, tmp);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lv[k] += h * bc.invMass * fe[k];
            fe[3 + k] *= h;
        }
        sym_mul(t3, S6, fe + 3);
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
    }

    // ---------------- joint feedback: lambda * J = lam' * Jh * M^1/2 ----------------
    int slot = 0;
    for (int j = 0; j < nj; j++) {
        const DevJoint &jc = T->joints[j];
        const int jr = G(P.jrows, j);
        const int mrows = jr & 7, lrow = (jr >> 4) - 1;
        float acc[12];
#pragma unroll
        for (int q = 0; q < 12; q++) acc[q] = 0.f;
        float flam = 0.f;
        for (int k = 0; k < mrows; k++, slot++) {
            const F8 u0 = ldg256(ROW_UNIT(P.rows, slot, 0, Wp, w));
            const float4 j0 = u0.lo, j1 = u0.hi, j2 = __ldg(ROW_UNIT(P.rows, slot, 1, Wp, w));
            const float lam = G(P.lambda, slot);
            // row layout (b2p_assemble.cu emit_row): J0..J7 | J8..J11 scalars
            acc[0] += j0.x * lam; acc[1] += j0.y * lam; acc[2] += j0.z * lam; acc[3] += j0.w * lam;
            acc[4] += j1.x * lam; acc[5] += j1.y * lam; acc[6] += j1.z * lam; acc[7] += j1.w * lam;
            acc[8] += j2.x * lam; acc[9] += j2.y * lam; acc[10] += j2.z * lam; acc[11] += j2.w * lam;
            if (k == lrow) flam = lam * G(P.rowS, slot);
        }
        float data[12];
#pragma unroll
        for (int q = 0; q < 12; q++) data[q] = 0.f;
        if (jc.b0 >= 0 && mrows > 0) {
            float S6[6];
            const float sm0 = T->bodies[jc.b0].sqrtMass;
#pragma unroll
            for (int k = 0; k < 6; k++) S6[k] = s_si[(size_t)(jc.b0 * 6 + k) * BD + tid];
            data[0] = sm0 * acc[0]; data[1] = sm0 * acc[1]; data[2] = sm0 * acc[2];
            sym_mul(data + 3, S6, acc + 3);
            if (jc.b1 >= 0) {
                const float sm1 = T->bodies[jc.b1].sqrtMass;
#pragma unroll
                for (int k = 0; k < 6; k++) S6[k] = s_si[(size_t)(jc.b1 * 6 + k) * BD + tid];
                data[6] = sm1 * acc[6]; data[7] = sm1 * acc[7]; data[8] = sm1 * acc[8];
                sym_mul(data + 9, S6, acc + 9);
            }
        }
#pragma unroll
        for (int q = 0; q < 12; q++) G(P.jfb, j * 13 + q) = data[q];
        G(P.jfb, j * 13 + 12) = flam;
    }
    // ---------------- contact feedback f1 ----------------
    int nc = G(P.ccount, 0);
    if (nc > maxc) nc = maxc;
    for (int c = 0; c < nc; c++) {
        const uint32_t cm = G(P.cmap, c);
        const int rows_c = cm & 7, depmask = (cm >> 3) & 31;
        int hs = (cm >> 8) & 0xff, ts = (cm >> 16) & 0xff;
        const int b0 = __float_as_int(G(P.crec, c * B2P_CONTACT_FIELDS + CR_IDS)) & 0xff;
        float f[3] = {0.f, 0.f, 0.f};
        for (int k = 0; k < rows_c; k++) {
            const int s = (k > 0 && (depmask & (1 << (k - 1)))) ? ts++ : hs++;
            const float4 j0 = __ldg(ROW_UNIT(P.rows, s, 0, Wp, w));
            const float lam = G(P.lambda, s);
            f[0] += j0.x * lam;
            f[1] += j0.y * lam;
            f[2] += j0.z * lam;
        }
        const float sm0 = T->bodies[b0].sqrtMass;
        G(P.cfb, c * 3 + 0) = sm0 * f[0];
        G(P.cfb, c * 3 + 1) = sm0 * f[1];
        G(P.cfb, c * 3 + 2) = sm0 * f[2];
    }

    // ---------------- joint observations from the new state ----------------
    for (int j = 0; j < nj; j++) {
        const DevJoint &jc = T->joints[j];
        float a1 = 0.f, r1 = 0.f, a2 = 0.f, r2 = 0.f;
        if (jc.b0 >= 0 && jc.type != 6 && jc.type != 4) {
            const bool has_b1 = jc.b1 >= 0;
            BodyPose p0, p1;
            load_pose(P, jc.b0, w, p0);
            float lv0[3], av0[3], lv1[3] = {0.f, 0.f, 0.f}, av1[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                lv0[k] = G(P.state, jc.b0 * B2P_STATE_FIELDS + 7 + k);
                av0[k] = G(P.state, jc.b0 * B2P_STATE_FIELDS + 10 + k);
            }
            if (has_b1) {
                load_pose(P, jc.b1, w, p1);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    lv1[k] = G(P.state, jc.b1 * B2P_STATE_FIELDS + 7 + k);
                    av1[k] = G(P.state, jc.b1 * B2P_STATE_FIELDS + 10 + k);
                }
            }
            float axis[3];
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
    }
}

} // namespace

extern "C" size_t b2p_integrate_smem_bytes(int nb) { return sizeof(float) * (size_t)nb * 6 * BD; }

extern "C" cudaError_t b2p_launch_integrate(const StepParams *P, int nb, cudaStream_t stream)
{
    static size_t configured_dev[B2P_MAX_DEVICES] = {0}; // cudaFuncSetAttribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    size_t &configured = configured_dev[dev % B2P_MAX_DEVICES];
    const size_t smem = b2p_integrate_smem_bytes(nb);
    if (smem > configured) {
        cudaError_t e =
            cudaFuncSetAttribute(b2p_integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int grid = (P->W + BD - 1) / BD;
    b2p_integrate_kernel<<<grid, BD, smem, stream>>>(*P);
    return cudaGetLastError();
}

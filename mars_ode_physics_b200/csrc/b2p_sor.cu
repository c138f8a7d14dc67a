// b2p_sor.cu -- kernel 2 of the step: the SOR-LCP sweeps of dWorldQuickStep with every row of a world
// resident in shared memory (sm_100a).
//
// What it computes: ODE 0.16 quickstep.cpp SOR_LCP as the reference reaches it through
// dWorldQuickStep (src/WorldPhysics.cpp:365): `iters` projected Gauss-Seidel sweeps over the m rows of
// a world, rows with findex < 0 first, both partitions re-shuffled (Fisher-Yates on dRandInt, seed
// src/WorldPhysics.cpp:169) at iterations 8 and 16, friction bounds following the normal row's current
// multiplier.  Rows arrive mass- and Ad-scaled (see b2p_assemble.cu), so one row-iteration is
//     delta = rhs' - cfm'*lam - Jh.fch ;  lam <- clamp(lam + delta) ;  fch += (lam_new - lam_old) * Jh
//
// Mapping: one thread per world, one warp (= one CTA) per 32 consecutive worlds; Gauss-Seidel is
// sequential inside a world and every world follows its own row order (contact counts and RNG histories
// differ), so the parallelism is across worlds.  Shared memory is lane-interleaved ([item][lane]), which
// makes every access conflict-free whatever row or body index each lane holds:
//     rows  float4[Rs][4][32]   the first Rs rows of each world (cp.async from the assemble kernel's
//                               [slot][4][Wp] layout: 512 contiguous bytes per warp and vector)
//     lam   float [R][32]       multipliers by slot (also read by friction rows through findex)
//     fch   float [nb*6][32]    M^-1/2 J^T lambda
//     ord   uint8 [R][32]       the world's current row order
// Rows beyond Rs (worlds with unusually many contacts) are read from global memory (L2) in place.
// HBM traffic is one read of the rows (64 B each) and one write of lam and fch: the sweeps themselves
// run out of shared memory.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "b2p_dev.h"
#include "b2p_math.cuh"

namespace {

constexpr int WL = 32; // worlds per CTA (one warp)

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

struct RowRegs {
    float4 a0, a1, a2, a3;
};

__global__ void __launch_bounds__(WL) b2p_sor_kernel(const StepParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevTemplate *__restrict__ T = P.T;
    const int nb = T->nb, R = T->maxrows, Rs = P.rows_resident;
    float4 *const s_rows = (float4 *)smem_raw;
    float *const s_lam = (float *)(s_rows + (size_t)Rs * 4 * WL);
    float *const s_fc = s_lam + (size_t)R * WL;
    uint8_t *const s_ord = (uint8_t *)(s_fc + (size_t)nb * 6 * WL);

    const int lane = threadIdx.x;
    const int w = blockIdx.x * WL + lane; // < Wp (padded worlds have no rows)
    const size_t Wp = (size_t)P.Wp;
    const int nr = w < P.W ? P.nrows[w] : 0;
    const int m = nr & 0xffff, nhead = nr >> 16;
    const int mmax = __reduce_max_sync(0xffffffffu, m);

    // rows -> shared memory (asynchronously, while lam / fch / ord are initialised)
    {
        const int nload = mmax < Rs ? mmax : Rs;
        const float4 *src = P.rows + w;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rows + lane);
        for (int i = 0; i < nload * 4; i++) cp_async16(dst + (uint32_t)(i * WL * 16), src + (size_t)i * Wp);
    }
    for (int s = 0; s < mmax; s++) {
        s_lam[s * WL + lane] = 0.f;
        s_ord[s * WL + lane] = (uint8_t)s;
    }
    for (int k = 0; k < nb * 6; k++) s_fc[k * WL + lane] = 0.f;
    uint32_t seed = P.seed[w];
    cp_async_commit_wait_all();
    __syncwarp();

    auto fetch = [&](int r) -> RowRegs {
        RowRegs q;
        if (r < Rs) {
            const float4 *p = s_rows + (size_t)(r * 4) * WL + lane;
            q.a0 = p[0];
            q.a1 = p[WL];
            q.a2 = p[2 * WL];
            q.a3 = p[3 * WL];
        } else {
            const float4 *p = P.rows + (size_t)(r * 4) * Wp + w;
            q.a0 = p[0];
            q.a1 = p[Wp];
            q.a2 = p[2 * Wp];
            q.a3 = p[3 * Wp];
        }
        return q;
    };

    for (int it = 0; it < P.iters; it++) {
        if (P.reorder == 0 && it >= 8 && (it & 7) == 0) {
            // ODE ConstraintsReorderingHelper on [0,nhead) and [nhead,m)
            auto swap = [&](int a, int b) {
                const uint8_t x = s_ord[a * WL + lane], y = s_ord[b * WL + lane];
                s_ord[a * WL + lane] = y;
                s_ord[b * WL + lane] = x;
            };
            for (int idx = 1; idx < nhead; idx++) swap(idx, rand_int(seed, idx + 1));
            for (int idx = nhead + 1; idx < m; idx++) swap(idx, rand_int(seed, idx - nhead + 1) + nhead);
        }
        int r = 0;
        RowRegs q;
        if (m > 0) {
            r = s_ord[lane];
            q = fetch(r);
        }
        for (int k = 0; k < mmax; k++) {
            // the next row does not depend on the sequential fch / lam state: fetch it ahead
            int rn = 0;
            RowRegs qn;
            if (k + 1 < m) {
                rn = s_ord[(k + 1) * WL + lane];
                qn = fetch(rn);
            }
            if (k < m) {
                const int meta = __float_as_int(q.a3.w);
                const int b0 = meta & 0xff, b1r = (meta >> 8) & 0xff, fr = (meta >> 16) & 0xff;
                const bool two = b1r != 0xff;
                const int b1 = two ? b1r : b0;
                float *f0 = s_fc + (size_t)(b0 * 6) * WL + lane;
                float *f1 = s_fc + (size_t)(b1 * 6) * WL + lane;
                const float lam_old = s_lam[r * WL + lane];
                const float x0 = f0[0], x1 = f0[WL], x2 = f0[2 * WL], x3 = f0[3 * WL], x4 = f0[4 * WL], x5 = f0[5 * WL];
                const float y0 = f1[0], y1 = f1[WL], y2 = f1[2 * WL], y3 = f1[3 * WL], y4 = f1[4 * WL], y5 = f1[5 * WL];
                // one-body rows carry zeros in the second half and alias body 0, so the dot is uniform
                const float d0 = q.a0.x * x0 + q.a0.y * x1 + q.a0.z * x2;
                const float d1 = q.a0.w * x3 + q.a1.x * x4 + q.a1.y * x5;
                const float d2 = q.a1.z * y0 + q.a1.w * y1 + q.a2.x * y2;
                const float d3 = q.a2.y * y3 + q.a2.z * y4 + q.a2.w * y5;
                float delta = (q.a3.x - lam_old * q.a3.y) - ((d0 + d1) + (d2 + d3));
                float lo_act, hi_act;
                const int enc = __float_as_int(q.a3.z);
                if (fr) {
                    hi_act = fabsf(q.a3.z * s_lam[(fr - 1) * WL + lane]);
                    lo_act = -hi_act;
                } else {
                    hi_act = enc == (int)0xFFFFFFFF ? INFINITY : (enc == (int)0xFFFFFFFE ? 0.f : q.a3.z);
                    lo_act = enc == (int)0xFFFFFFFF ? 0.f : (enc == (int)0xFFFFFFFE ? -INFINITY : -q.a3.z);
                }
                const float un = lam_old + delta;
                const float nl = fminf(fmaxf(un, lo_act), hi_act);
                delta = (nl == un) ? delta : nl - lam_old;
                s_lam[r * WL + lane] = nl;
                f0[0] = x0 + delta * q.a0.x;
                f0[WL] = x1 + delta * q.a0.y;
                f0[2 * WL] = x2 + delta * q.a0.z;
                f0[3 * WL] = x3 + delta * q.a0.w;
                f0[4 * WL] = x4 + delta * q.a1.x;
                f0[5 * WL] = x5 + delta * q.a1.y;
                if (two) {
                    f1[0] = y0 + delta * q.a1.z;
                    f1[WL] = y1 + delta * q.a1.w;
                    f1[2 * WL] = y2 + delta * q.a2.x;
                    f1[3 * WL] = y3 + delta * q.a2.y;
                    f1[4 * WL] = y4 + delta * q.a2.z;
                    f1[5 * WL] = y5 + delta * q.a2.w;
                }
            }
            q = qn;
            r = rn;
        }
    }

    if (w < P.W) {
        for (int s = 0; s < m; s++) P.lambda[(size_t)s * Wp + w] = s_lam[s * WL + lane];
        for (int k = 0; k < nb * 6; k++) P.fcacc[(size_t)k * Wp + w] = s_fc[k * WL + lane];
        P.seed[w] = seed;
    }
}

} // namespace

extern "C" size_t b2p_sor_smem_bytes(int nb, int maxrows, int rows_resident)
{
    size_t s = sizeof(float4) * (size_t)rows_resident * 4 * WL;
    s += sizeof(float) * (size_t)maxrows * WL;
    s += sizeof(float) * (size_t)nb * 6 * WL;
    s += (size_t)maxrows * WL;
    return (s + 15) & ~(size_t)15;
}

extern "C" cudaError_t b2p_launch_sor(const StepParams *P, int nb, int maxrows, cudaStream_t stream)
{
    static size_t configured = 0;
    const size_t smem = b2p_sor_smem_bytes(nb, maxrows, P->rows_resident);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(b2p_sor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int grid = (P->W + WL - 1) / WL;
    b2p_sor_kernel<<<grid, WL, smem, stream>>>(*P);
    return cudaGetLastError();
}

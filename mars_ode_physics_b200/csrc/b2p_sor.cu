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
// Mapping: two threads per world (lane 2v = body-0 side of world v, lane 2v+1 = body-1 side), one warp
// (= one CTA) per 16 consecutive worlds.  Gauss-Seidel is sequential inside a world and every world follows
// its own row order (contact counts and RNG histories differ), so the parallelism is across worlds; the
// throughput of the kernel is (worlds resident per SM) / (latency of one row-iteration), which is why a
// world gets two lanes (half the instructions on the dependent chain, twice the warps for the same
// shared-memory footprint) and why the rows are kept to 64 bytes.
// Shared memory per CTA, every access conflict-free whatever row or body each world is at:
//     rows  float4[Rs][2][32]   row r: lane L reads vectors (r*2+0)*32+L and (r*2+1)*32+L = its side's six
//                               Jacobian words + two of the four scalars (cp.async from the assemble
//                               kernel's [slot][4][Wp] layout, 256 contiguous bytes per side and vector)
//     lam   float [R][16]       multipliers by slot (also read by friction rows through findex)
//     fch   float [(nb+1)*6][16] M^-1/2 J^T lambda (+ a dummy body that one-body rows name as body 1);
//                               side 1 walks the six components angular-first, so the two sides of a warp
//                               instruction always hit different bank halves
//     ord   uint8 [R][16]       the world's current row order
// Rows beyond Rs (worlds with unusually many contacts) are read from global memory (L2) in place.
// HBM traffic is one read of the rows (64 B each) and one write of lam and fch: the sweeps themselves
// run out of shared memory.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "b2p_dev.h"
#include "b2p_math.cuh"

namespace {

constexpr int WL = 16; // worlds per CTA (one warp, two lanes per world)

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// This lane's view of one row: its side's six Jacobian words (side 1: angular first) and the four scalars
struct RowRegs {
    float j0, j1, j2, j3, j4, j5;
    float rhs, cfm, bound;
    int meta;
};

template <bool SPILL> __global__ void __launch_bounds__(32) b2p_sor_kernel(const StepParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevTemplate *__restrict__ T = P.T;
    const int nb = T->nb, R = T->maxrows, Rs = P.rows_resident;
    float4 *const s_rows = (float4 *)smem_raw;
    float *const s_lam = (float *)(s_rows + (size_t)Rs * 2 * 32);
    float *const s_fc = s_lam + (size_t)R * WL;
    uint8_t *const s_ord = (uint8_t *)(s_fc + (size_t)(nb + 1) * 6 * WL);

    const int lane = threadIdx.x, v = lane >> 1, side = lane & 1;
    const int w = blockIdx.x * WL + v; // < Wp (padded worlds have no rows)
    const size_t Wp = (size_t)P.Wp;
    const int nr = w < P.W ? P.nrows[w] : 0;
    const int m = nr & 0xffff, nhead = nr >> 16;
    const int mmax = __reduce_max_sync(0xffffffffu, m);

    // rows -> shared memory (asynchronously, while lam / fch / ord are initialised)
    {
        const int nload = mmax < Rs ? mmax : Rs;
        const float4 *src = P.rows + (size_t)(side * 2) * Wp + w;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rows + lane);
        for (int r = 0; r < nload; r++) {
            cp_async16(dst + (uint32_t)((r * 2 + 0) * 32 * 16), src + (size_t)(r * 4 + 0) * Wp);
            cp_async16(dst + (uint32_t)((r * 2 + 1) * 32 * 16), src + (size_t)(r * 4 + 1) * Wp);
        }
    }
    for (int s = side; s < R; s += 2) {
        s_lam[s * WL + v] = 0.f;
        s_ord[s * WL + v] = (uint8_t)s;
    }
    for (int k = side; k < (nb + 1) * 6; k += 2) s_fc[k * WL + v] = 0.f;
    uint32_t seed = P.seed[w];
    cp_async_commit_wait_all();
    __syncwarp();

    auto fetch = [&](int r) -> RowRegs {
        float4 h0, h1;
        if (!SPILL || r < Rs) {
            const float4 *p = s_rows + (size_t)(r * 2) * 32 + lane;
            h0 = p[0];
            h1 = p[32];
        } else {
            const float4 *p = P.rows + (size_t)(r * 4 + side * 2) * Wp + w;
            h0 = p[0];
            h1 = p[Wp];
        }
        // the four scalars sit two per side: swap them between the lanes of the pair
        const float oz = __shfl_xor_sync(0xffffffffu, h1.z, 1), ow = __shfl_xor_sync(0xffffffffu, h1.w, 1);
        RowRegs q;
        q.j0 = h0.x, q.j1 = h0.y, q.j2 = h0.z, q.j3 = h0.w, q.j4 = h1.x, q.j5 = h1.y;
        q.rhs = side ? oz : h1.z;
        q.cfm = side ? ow : h1.w;
        q.bound = side ? h1.z : oz;
        q.meta = __float_as_int(side ? h1.w : ow);
        return q;
    };
    // byte offsets of this lane's six fch components relative to its body's first word: side 0 walks
    // 0..5, side 1 walks 3,4,5,0,1,2 (other bank half)
    const int offA = side ? 3 * WL : 0, offB = side ? -3 * WL : 0;
    const int bshift = side * 8;
    // One row-iteration (ODE quickstep.cpp SOR_LCP inner loop, in the scaled coordinates).  Branch-free:
    // one-body rows name the dummy body nb (its fch stays 0 because their second Jacobian half is 0), and
    // the bound shapes are selected, not branched on.  Both lanes of a pair compute identical lambdas.
    auto relax = [&](const RowRegs &q, int r, bool active) {
        const int b = (q.meta >> bshift) & 0xff, fr = (q.meta >> 16) & 0xff;
        float *fA = s_fc + (size_t)(b * 6) * WL + v + offA;
        float *fB = fA + (offB - offA) + 3 * WL;
        float *lp = s_lam + r * WL + v;
        const float lam_n = s_lam[(fr ? fr - 1 : r) * WL + v];
        const float lam_old = *lp;
        const float x0 = fA[0], x1 = fA[WL], x2 = fA[2 * WL], x3 = fB[0], x4 = fB[WL], x5 = fB[2 * WL];
        const float d = (q.j0 * x0 + q.j1 * x1 + q.j2 * x2) + (q.j3 * x3 + q.j4 * x4 + q.j5 * x5);
        const float dot = d + __shfl_xor_sync(0xffffffffu, d, 1);
        float delta = (q.rhs - lam_old * q.cfm) - dot;
        // meta bits 24/25: lower / upper bound is 0 (ODE lo = 0 / hi = 0); fr: bound follows the normal row
        const float bnd = fr ? fabsf(q.bound * lam_n) : q.bound;
        const float hi_act = (q.meta & (2 << 24)) ? 0.f : bnd;
        const float lo_act = (q.meta & (1 << 24)) ? 0.f : -bnd;
        const float un = lam_old + delta;
        const float nl = fminf(fmaxf(un, lo_act), hi_act);
        delta = (nl == un) ? delta : nl - lam_old;
        if (active) {
            *lp = nl;
            fA[0] = x0 + delta * q.j0;
            fA[WL] = x1 + delta * q.j1;
            fA[2 * WL] = x2 + delta * q.j2;
            fB[0] = x3 + delta * q.j3;
            fB[WL] = x4 + delta * q.j4;
            fB[2 * WL] = x5 + delta * q.j5;
        }
    };

    for (int it = 0; it < (mmax > 0 ? P.iters : 0); it++) {
        if (P.reorder == 0 && it >= 8 && (it & 7) == 0) {
            // ODE ConstraintsReorderingHelper on [0,nhead) and [nhead,m); one lane of the pair shuffles
            if (side == 0) {
                auto swap = [&](int a, int b) {
                    const uint8_t x = s_ord[a * WL + v], y = s_ord[b * WL + v];
                    s_ord[a * WL + v] = y;
                    s_ord[b * WL + v] = x;
                };
#pragma unroll 1
                for (int idx = 1; idx < nhead; idx++) swap(idx, rand_int(seed, idx + 1));
#pragma unroll 1
                for (int idx = nhead + 1; idx < m; idx++) swap(idx, rand_int(seed, idx - nhead + 1) + nhead);
            }
            __syncwarp();
        }
        // Software pipeline over sweep positions, two per trip so that the row registers ping-pong without
        // moves: the row of position k+1 and the order byte of position k+2 are fetched before row k is
        // relaxed (neither depends on the sequential fch / lam state).  Positions >= m of a world read
        // order byte 0 (harmless) and are computed but not stored.
        RowRegs qa, qb;
        int ra = s_ord[v], rb = s_ord[(1 < R ? 1 : 0) * WL + v];
        qa = fetch(ra);
#pragma unroll 1
        for (int k = 0; k < mmax; k += 2) {
            const int rc = s_ord[(k + 2 < R ? k + 2 : 0) * WL + v];
            qb = fetch(rb);
            relax(qa, ra, k < m);
            const int rd = s_ord[(k + 3 < R ? k + 3 : 0) * WL + v];
            qa = fetch(rc);
            if (k + 1 < mmax) relax(qb, rb, k + 1 < m); // (rows >= mmax were never loaded)
            ra = rc;
            rb = rd;
        }
    }

    if (w < P.W) {
        for (int s = side; s < m; s += 2) P.lambda[(size_t)s * Wp + w] = s_lam[s * WL + v];
        for (int k = side; k < nb * 6; k += 2) P.fcacc[(size_t)k * Wp + w] = s_fc[k * WL + v];
        if (side == 0) P.seed[w] = seed;
    }
}

} // namespace

extern "C" size_t b2p_sor_smem_bytes(int nb, int maxrows, int rows_resident)
{
    size_t s = sizeof(float4) * (size_t)rows_resident * 2 * 32;
    s += sizeof(float) * (size_t)maxrows * WL;
    s += sizeof(float) * (size_t)(nb + 1) * 6 * WL;
    s += (size_t)maxrows * WL;
    return (s + 15) & ~(size_t)15;
}

extern "C" cudaError_t b2p_launch_sor(const StepParams *P, int nb, int maxrows, cudaStream_t stream)
{
    static size_t configured[2] = {0, 0};
    const size_t smem = b2p_sor_smem_bytes(nb, maxrows, P->rows_resident);
    const bool spill = P->rows_resident < maxrows;
    auto kern = spill ? b2p_sor_kernel<true> : b2p_sor_kernel<false>;
    if (smem > configured[spill]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[spill] = smem;
    }
    const int grid = (P->W + WL - 1) / WL;
    kern<<<grid, 32, smem, stream>>>(*P);
    return cudaGetLastError();
}

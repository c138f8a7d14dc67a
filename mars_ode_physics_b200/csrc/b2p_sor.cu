// b2p_sor.cu -- kernel 2 of the step: the SOR-LCP sweeps of dWorldQuickStep with every row of a world
// resident on chip, the first KR sweep positions in registers and the rest in shared memory (sm_100a).
//
// What it computes: ODE 0.16 quickstep.cpp SOR_LCP as the reference reaches it through
// dWorldQuickStep (src/WorldPhysics.cpp:365): `iters` projected Gauss-Seidel sweeps over the m rows of
// a world, rows with findex < 0 first, both partitions re-shuffled (Fisher-Yates on dRandInt, seed
// src/WorldPhysics.cpp:169) at iterations 8 and 16, friction bounds following the normal row's current
// multiplier.  Rows arrive mass- and Ad-scaled (see b2p_assemble.cu), so one row-iteration is
//     delta = rhs' - cfm'*lam - Jh.fch ;  lam <- clamp(lam + delta) ;  fch += (lam_new - lam_old) * Jh
//
// Mapping: two threads per world (one per body side of a row), one warp (= one CTA) per 16 consecutive worlds.
// Lanes 0-7 / 16-23 are the body-0 sides of worlds 0-7 / 8-15, lanes 24-31 / 8-15 their body-1 sides (partner =
// lane ^ 24): every half-warp then holds 16 different worlds, which makes all 8- and 16-byte shared-memory
// accesses conflict-free without padding or swizzles.  Gauss-Seidel is sequential inside a world and every world follows
// its own row order (contact counts and RNG histories differ), so the parallelism is across worlds and the
// throughput of the kernel is (worlds resident per SM) / (latency of one row-iteration).  Both factors are
// limited by where the rows live, hence:
//   * rows are held PHYSICALLY in sweep order.  The order only changes three times per step (identity at
//     iteration 0, ODE's reshuffles at 8 and 16); at those points every lane re-gathers its half of each
//     row from the slot-ordered rows in global memory (L2 hits after the first pass).  Inside the sweeps
//     nothing is indexed by the order any more;
//   * sweep positions [0,KR) of every world sit in registers (10 per row and lane: six Jacobian words, rhs',
//     cfm', bound, meta) -- the loop over them is fully unrolled, so the "row fetch" costs no instruction.
//     The register file (64 K x 32 bit per SM) is larger than shared memory; using both doubles the worlds
//     an SM can hold compared to rows in shared memory only;
//   * positions [KR, KR+Rs) sit in shared memory ([position][4][16 worlds] float4, cp.async-gathered);
//     positions beyond that (worlds with unusually many contacts) are read from global memory in place;
//   * positions m..mmax-1 of a world with fewer rows than its warp's longest world hold an all-zero row
//     that points at a dummy multiplier and the dummy body: the sweep is branch- and predicate-free.
// Shared memory per CTA, every access conflict-free whatever row or body each world is at:
//     rows  float4[Rs][4][16]     granule g of position p: J0..J3 | J4..J7 | J8..J11 | rhs' cfm' bound meta
//                                 (J0..J5 = body-0 side, J6..J11 = body-1 side); a lane reads its six words
//                                 with three 8-byte loads and the scalars with one 16-byte load
//     lam   float [R+2][16]       multipliers by slot; slot R = dummy (zero rows), slot R+1 = constant 1 (the
//                                 "normal multiplier" of rows whose bound does not follow another row)
//     fch   float2[(nb+1)*3][16]  M^-1/2 J^T lambda: (lx ly) (lz ax) (ay az) per body, + the dummy body that
//                                 one-body rows name as body 1
//     ord   uint8 [R][16]         the world's current row order (position -> slot)
// HBM traffic is one read of the rows (64 B each) and one write of lam and fch.
#include <cuda_runtime.h>
#include <mutex>
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

// This lane's view of one row: its side's six Jacobian words and the four scalars.
// meta (written by b2p_assemble.cu emit_row): own slot | slot of the multiplier the bound follows << 8 |
// body 0 << 16 | body 1 << 24 | "lower bound is 0" << 30 | "upper bound is 0" << 31
struct Row {
    float j0, j1, j2, j3, j4, j5;
    float rhs, cfm, bound;
    uint32_t meta;
};

template <int KR> __global__ void __launch_bounds__(32) b2p_sor_kernel(const StepParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevTemplate *__restrict__ T = P.T;
    const int nb = T->nb, R = T->maxrows, Rs = P.rows_resident;
    float4 *const s_rows = (float4 *)smem_raw;
    float2 *const s_fc = (float2 *)(s_rows + (size_t)Rs * 4 * WL);
    float *const s_lam = (float *)(s_fc + (size_t)(nb + 1) * 3 * WL);
    uint8_t *const s_ord = (uint8_t *)(s_lam + (size_t)(R + 2) * WL);

    const int lane = threadIdx.x, side = (lane >> 3) & 1, v = (lane & 7) + 8 * ((lane >> 4) ^ side);
    const int w = blockIdx.x * WL + v; // < Wp (padded worlds have no rows)
    const size_t Wp = (size_t)P.Wp;
    const int nr = w < P.W ? P.nrows[w] : 0;
    const int m = nr & 0xffff, nisl = nr >> 16; // rows, island blocks (b2p_dev.h)
    const int mmax = __reduce_max_sync(0xffffffffu, m);
    const int nev = P.reorder == 0 && P.iters > 8 ? (P.iters - 1) >> 3 : 0; // reshuffles per solve (iterations 8, 16, ..)

    for (int s = side; s < R + 2; s += 2) s_lam[s * WL + v] = s == R + 1 ? 1.f : 0.f;
    // the used slots in ODE's initial sweep order (b2p_assemble.cu); positions behind the world's rows are not read
    for (int p = side; p < m; p += 2) s_ord[p * WL + v] = P.ord0[(size_t)p * Wp + w];
    for (int k = side; k < (nb + 1) * 3; k += 2) s_fc[k * WL + v] = make_float2(0.f, 0.f);
    const uint32_t seed = P.seed[w]; // the world's LCG state at the start of the step
    uint32_t seed_out = seed;        // ... and after the step's reshuffles
    __syncwarp();

    // ---- this lane's addresses
    float *const lamL = s_lam + v;
    float2 *const fcL = s_fc + v; // fch of body b, world v: fcL[(b*3 + {0,1,2})*16]
    const int bshift = 16 + 8 * side;
    // 8-byte granules of this side inside a shared-memory row block (float2 units): words side*6 + {0,2,4}
    const int w0 = side * 6;
    const int o0 = ((w0 >> 2) * WL + v) * 2 + ((w0 >> 1) & 1);
    const int o1 = (((w0 + 2) >> 2) * WL + v) * 2 + (((w0 + 2) >> 1) & 1);
    const int o2 = (((w0 + 4) >> 2) * WL + v) * 2 + (((w0 + 4) >> 1) & 1);

    // this lane's half of the row in slot `slot` of its world, from the slot-ordered rows in global memory
    // (slot R is an all-zero row that names the dummy multiplier and the dummy body)
    auto gload = [&](int slot) -> Row {
        // this side's words are side*6 .. side*6 + 5 of the 16-word row = two 32-byte units: side 0 takes unit 0
        // (J0..J7) and the scalars of unit 1, side 1 takes unit 1 (J8..J11, scalars) and J6 J7 of unit 0
        const F8 u = ldg256(ROW_UNIT(P.rows, slot, side, Wp, w));
        const float4 x = __ldg(ROW_UNIT(P.rows, slot, side ^ 1, Wp, w) + 1);
        const float4 v3 = side ? u.hi : x;
        Row q;
        q.j0 = side ? x.z : u.lo.x, q.j1 = side ? x.w : u.lo.y, q.j2 = side ? u.lo.x : u.lo.z;
        q.j3 = side ? u.lo.y : u.lo.w, q.j4 = side ? u.lo.z : u.hi.x, q.j5 = side ? u.lo.w : u.hi.y;
        q.rhs = v3.x;
        q.cfm = v3.y;
        q.bound = v3.z;
        q.meta = (uint32_t)__float_as_int(v3.w);
        return q;
    };
    // position p - KR of the shared-memory rows
    auto sload = [&](int p) -> Row {
        const float2 *b2 = (const float2 *)(s_rows + (size_t)p * 4 * WL);
        const float2 a = b2[o0], b = b2[o1], c = b2[o2];
        const float4 s4 = s_rows[(size_t)(p * 4 + 3) * WL + v];
        Row q;
        q.j0 = a.x, q.j1 = a.y, q.j2 = b.x, q.j3 = b.y, q.j4 = c.x, q.j5 = c.y;
        q.rhs = s4.x, q.cfm = s4.y, q.bound = s4.z;
        q.meta = (uint32_t)__float_as_int(s4.w);
        return q;
    };

    // One row-iteration (ODE quickstep.cpp SOR_LCP inner loop, in the scaled coordinates).  Branch-free:
    // one-body rows name the dummy body nb (its fch stays 0 because their second Jacobian half is 0), rows
    // whose bound is fixed read the constant-1 multiplier, and the bound shapes are selected, not branched
    // on.  Both lanes of a pair compute identical lambdas.
    auto relax = [&](const Row &q) {
        const uint32_t mt = q.meta;
        float *lp = lamL + (mt & 0xff) * WL;
        const float *ln = lamL + ((mt >> 8) & 0xff) * WL;
        const int b = (mt >> bshift) & 0x3f;
        float2 *pf = fcL + b * (3 * WL);
        float2 fa = pf[0], fb = pf[WL], fc = pf[2 * WL];
        const float lam_old = *lp, lam_n = *ln;
        // (even terms) + (odd terms), each a chain of fused multiply-adds: the association b2p_sor_tmem.cu gets
        // from its packed FFMA2 chains
        const float de = sor_fma(q.j4, fc.x, sor_fma(q.j2, fb.x, __fmul_rn(q.j0, fa.x)));
        const float dod = sor_fma(q.j5, fc.y, sor_fma(q.j3, fb.y, __fmul_rn(q.j1, fa.y)));
        const float d = de + dod;
        const float dot = d + __shfl_xor_sync(0xffffffffu, d, 24);
        float delta = (q.rhs - lam_old * q.cfm) - dot;
        const float bnd = fabsf(q.bound * lam_n);
        const float hi_act = (mt & 0x80000000u) ? 0.f : bnd;
        const float lo_act = (mt & 0x40000000u) ? 0.f : -bnd;
        const float un = lam_old + delta;
        const float nl = fminf(fmaxf(un, lo_act), hi_act);
        delta = (nl == un) ? delta : nl - lam_old;
        *lp = nl;
        fa.x = sor_fma(delta, q.j0, fa.x), fa.y = sor_fma(delta, q.j1, fa.y), fb.x = sor_fma(delta, q.j2, fb.x);
        fb.y = sor_fma(delta, q.j3, fb.y), fc.x = sor_fma(delta, q.j4, fc.x), fc.y = sor_fma(delta, q.j5, fc.y);
        pf[0] = fa;
        pf[WL] = fb;
        pf[2 * WL] = fc;
    };

    Row rr[KR > 0 ? KR : 1];
    const int ks = mmax < KR + Rs ? mmax : KR + Rs; // end of the shared-memory positions

    // (re)load every on-chip row in the current order; positions >= m of a world get the zero row (slot R)
    auto load_rows = [&]() {
#pragma unroll
        for (int k = 0; k < KR; k++) {
            if ((k & ~3) < mmax) // whole groups of four: the sweep skips register positions by group
                rr[k] = gload(k < m ? s_ord[k * WL + v] : R);
        }
        // lane (v, side) copies granules 2*side and 2*side + 1 of its world's rows
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rows + (side * 2) * WL + v);
        for (int k = KR; k < ks; k++) {
            const int slot = k < m ? s_ord[k * WL + v] : R;
            const float4 *src = ROW_UNIT(P.rows, slot, side, Wp, w); // granules 2*side, 2*side + 1
            const uint32_t d = dst + (uint32_t)((k - KR) * 4 * WL * 16);
            cp_async16(d, src);
            cp_async16(d + WL * 16, src + 1);
        }
        cp_async_commit_wait_all();
        __syncwarp();
    };
    if (mmax > 0) load_rows();

    for (int it = 0; it < (mmax > 0 ? P.iters : 0); it++) {
        if (P.reorder == 0 && it >= 8 && (it & 7) == 0) {
            // ODE ConstraintsReorderingHelper on both partitions of every island block; one lane of the pair shuffles
            if (side == 0) {
                auto swap = [&](int a, int b) {
                    const uint8_t x = s_ord[a * WL + v], y = s_ord[b * WL + v];
                    s_ord[a * WL + v] = y;
                    s_ord[b * WL + v] = x;
                };
                seed_out = island_reshuffle(P.isl, Wp, w < P.W ? w : 0, nisl, seed, (it >> 3) - 1, nev, swap);
            }
            __syncwarp();
            load_rows();
        }
        // positions in registers
#pragma unroll
        for (int g = 0; g < KR; g += 4) {
            if (g < mmax) {
#pragma unroll
                for (int k = g; k < g + 4 && k < KR; k++) relax(rr[k]);
            }
        }
        // positions in shared memory, software-pipelined two per trip (the fetch of the next row does not
        // depend on the sequential fch / lam state)
        if (KR < ks) {
            Row qa = sload(0), qb;
#pragma unroll 1
            for (int p = 0; p < ks - KR; p += 2) {
                qb = sload(p + 1 < Rs ? p + 1 : 0);
                relax(qa);
                qa = sload(p + 2 < Rs ? p + 2 : 0);
                if (p + 1 < ks - KR) relax(qb);
            }
        }
        // positions that did not fit on chip: read in place (L2)
        if (KR + Rs < mmax) {
            Row qa = gload(KR + Rs < m ? s_ord[(KR + Rs) * WL + v] : R), qb;
#pragma unroll 1
            for (int k = KR + Rs; k < mmax; k += 2) {
                qb = gload(k + 1 < m ? s_ord[(k + 1) * WL + v] : R);
                relax(qa);
                qa = gload(k + 2 < m ? s_ord[(k + 2) * WL + v] : R);
                if (k + 1 < mmax) relax(qb);
            }
        }
    }

    if (w < P.W) {
        for (int p = side; p < m; p += 2) {
            const int s = s_ord[p * WL + v];
            P.lambda[(size_t)s * Wp + w] = s_lam[s * WL + v];
        }
        for (int k = side; k < nb * 3; k += 2) {
            const float2 f = s_fc[k * WL + v];
            float *dst = P.fcacc + (size_t)(k * 2) * Wp + w;
            dst[0] = f.x;
            dst[Wp] = f.y;
        }
        if (side == 0) P.seed[w] = seed_out;
    }
}

} // namespace

extern "C" size_t b2p_sor_smem_bytes(int nb, int maxrows, int rows_resident)
{
    size_t s = sizeof(float4) * (size_t)rows_resident * 4 * WL;
    s += sizeof(float2) * (size_t)(nb + 1) * 3 * WL;
    s += sizeof(float) * (size_t)(maxrows + 2) * WL;
    s += (size_t)maxrows * WL;
    return (s + 15) & ~(size_t)15;
}

// Sweep positions the kernel keeps in registers for a scene with at most `maxrows` rows per world.
extern "C" int b2p_sor_register_rows(int maxrows, int wanted)
{
    int kr = wanted >= 0 ? wanted : 16;
    if (kr > maxrows) kr = (maxrows + 3) & ~3;
    kr &= ~3;
    return kr > 20 ? 20 : kr;
}

extern "C" cudaError_t b2p_launch_sor(const StepParams *P, int nb, int maxrows, cudaStream_t stream)
{
    static size_t configured_dev[B2P_MAX_DEVICES][6] = {{0}}; // cudaFuncSetAttribute is per device
    static std::mutex configured_mutex; // arenas of several threads / devices may launch at the same time
    std::lock_guard<std::mutex> configured_lock(configured_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    size_t *configured = configured_dev[dev % B2P_MAX_DEVICES];
    const size_t smem = b2p_sor_smem_bytes(nb, maxrows, P->rows_resident);
    void (*kern)(const StepParams) = nullptr;
    switch (P->rows_registers) {
    case 0: kern = b2p_sor_kernel<0>; break;
    case 4: kern = b2p_sor_kernel<4>; break;
    case 8: kern = b2p_sor_kernel<8>; break;
    case 12: kern = b2p_sor_kernel<12>; break;
    case 16: kern = b2p_sor_kernel<16>; break;
    case 20: kern = b2p_sor_kernel<20>; break;
    default: return cudaErrorInvalidValue;
    }
    const int ci = P->rows_registers / 4;
    if (smem > configured[ci]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[ci] = smem;
    }
    const int grid = (P->W + WL - 1) / WL;
    kern<<<grid, 32, smem, stream>>>(*P);
    return cudaGetLastError();
}

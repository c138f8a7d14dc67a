// b2p_sor_tmem.cu -- the SOR-LCP sweeps of dWorldQuickStep with the rows of a world in Tensor Memory (sm_100a).
//
// Same computation as b2p_sor.cu (ODE 0.16 quickstep.cpp SOR_LCP behind dWorldQuickStep,
// src/WorldPhysics.cpp:365; see there for the scaled row-iteration and the row format), different mapping:
// ONE thread per world, 128 or 256 worlds per CTA, one CTA per SM, and the rows held physically in sweep order in
//   * registers:      sweep positions [0, KR)            (16 words per row)
//   * Tensor Memory:  positions [KR, KR+TROWS)           (the SM's 256 KB TMEM = 128 lanes x 512 columns: with
//                                                         4 warps a thread owns a whole lane = 32 rows of 16
//                                                         words; with 8 warps two warps split a lane quadrant's
//                                                         columns, 16 rows per world)
//   * shared memory:  the warp's overflow positions      (a pool of [position][4][32 worlds] float4 units handed
//                                                         out per warp: the worlds of a CTA are sorted by row
//                                                         count first, so only the warps holding the long worlds
//                                                         need overflow rows)
//   * L2 in place:    anything beyond (worlds with unusually many contacts).
// Blackwell's TMEM is meant for tensor-core accumulators, but tcgen05.st / tcgen05.ld (32x32b shape) give every
// thread of a warp private access to its own lane at a warp-uniform column.  Because the rows are kept in sweep
// order (re-gathered from the slot-ordered rows in global memory when ODE reshuffles at iterations 8 and 16),
// "the row of position p" IS a warp-uniform column although every world has a different constraint there.
// This frees shared memory -- and its single 128 B/clk pipe -- for what really needs per-lane dynamic
// addressing: fch (indexed by body) and lambda (indexed by slot).  With one thread per world there is no
// cross-lane reduction on the dependent chain, and a warp instruction advances 32 worlds instead of 16.
// Two rows are fetched per tcgen05.ld (x32) right before they are relaxed; the second warp of the scheduler
// covers the load's latency.
//
// Shared memory per CTA (all accesses conflict-free: consecutive lanes = consecutive columns):
//     fch   float4[nb+1][TW] + float2[nb+1][TW] (words 0..3 | 4, 5 of a body), lam float[R+2][TW], ord uint8[R][TW]:
//           one column per thread
//     pool  float4[units][4][32]                                          overflow positions
// The arithmetic of a row-iteration is term-by-term the one of b2p_sor.cu, so both kernels produce bitwise
// identical multipliers (tests/test_gpu_parity.py::test_sor_kernels_agree).
#include <cuda_runtime.h>
#include <mutex>
#include <math.h>
#include <stdint.h>

#include "b2p_dev.h"
#include "b2p_math.cuh"

namespace {

constexpr int TCOLS = 512; // TMEM columns allocated per CTA (all of them: one CTA per SM)

struct Row {
    float j[12];
    float rhs, cfm, bound;
    uint32_t meta;
};

__device__ __forceinline__ void tm_store_row(uint32_t taddr, const Row &q)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
        "f"(q.j[0]), "f"(q.j[1]), "f"(q.j[2]), "f"(q.j[3]), "f"(q.j[4]), "f"(q.j[5]), "f"(q.j[6]), "f"(q.j[7]), "f"(q.j[8]),
        "f"(q.j[9]), "f"(q.j[10]), "f"(q.j[11]), "f"(q.rhs), "f"(q.cfm), "f"(q.bound), "r"(q.meta)
        : "memory");
}
// Two consecutive rows (32 columns) with one instruction and one wait.  tcgen05.ld is asynchronous: the registers
// are valid only after tcgen05.wait::ld, which is part of the same asm statement so that the compiler cannot move
// a use in between.  A ld / wait pair costs about 85 cycles of issue time whatever its width (scratch
// microbenchmark), which is why rows travel in pairs.
__device__ __forceinline__ void tm_load_2rows(uint32_t taddr, Row &a, Row &b)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=f"(a.j[0]), "=f"(a.j[1]), "=f"(a.j[2]), "=f"(a.j[3]), "=f"(a.j[4]), "=f"(a.j[5]), "=f"(a.j[6]), "=f"(a.j[7]),
          "=f"(a.j[8]), "=f"(a.j[9]), "=f"(a.j[10]), "=f"(a.j[11]), "=f"(a.rhs), "=f"(a.cfm), "=f"(a.bound), "=r"(a.meta),
          "=f"(b.j[0]), "=f"(b.j[1]), "=f"(b.j[2]), "=f"(b.j[3]), "=f"(b.j[4]), "=f"(b.j[5]), "=f"(b.j[6]), "=f"(b.j[7]),
          "=f"(b.j[8]), "=f"(b.j[9]), "=f"(b.j[10]), "=f"(b.j[11]), "=f"(b.rhs), "=f"(b.cfm), "=f"(b.bound), "=r"(b.meta)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
// NW warps per CTA (4 or 8): NW*32 worlds per SM.  With 8 warps two warps share a TMEM lane quadrant and split its
// columns, so a world has 16 instead of 32 rows in TMEM but every scheduler has two warps to alternate between.
template <int KR, int NW> __global__ void __launch_bounds__(NW * 32, 1) b2p_sor_tmem_kernel(const StepParams P)
{
    constexpr int TW = NW * 32;            // worlds (= threads) per CTA
    constexpr int TROWS = 32 / (NW / 4);   // rows per world in Tensor Memory
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_need[NW];
    const DevTemplate *__restrict__ T = P.T;
    const int nb = T->nb, R = T->maxrows, pool = P.rows_resident;
    // fch of body b: words 0..3 in a float4 plane, words 4, 5 in a float2 plane (one LDS.128 + one LDS.64 per body)
    float4 *const s_fc4 = (float4 *)smem_raw;
    float2 *const s_fc2 = (float2 *)(s_fc4 + (size_t)(nb + 1) * TW);
    float *const s_lam = (float *)(s_fc2 + (size_t)(nb + 1) * TW);
    uint8_t *const s_ord = (uint8_t *)(s_lam + (size_t)(R + 2) * TW);
    // overflow rows: `pool` units of one sweep position x 32 worlds, handed to the warps that need them
    float4 *const s_pool = (float4 *)(smem_raw + (((size_t)(nb + 1) * 3 * TW * 8 + (size_t)(R + 2) * TW * 4 + (size_t)R * TW + 15) & ~(size_t)15));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- Tensor Memory: warp 0 allocates all 512 columns; warp i owns lanes 32(i&3) .., columns 256(i>>2) ..
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&s_tmem_base)),
                     "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }

    // ---- sort the CTA's worlds by row count (descending) so that the 32 worlds of a warp are about equally
    // long: a warp sweeps as many positions as its longest world has rows.  Counting sort in shared memory
    // (scratch aliases the overflow pool, which is not in use yet); which of two equally long worlds comes
    // first is arbitrary and does not influence any result.
    const size_t Wp = (size_t)P.Wp;
    // CTAs take the worlds from the end: the assemble kernel wrote the rows of the last worlds last, so the first
    // wave finds them in L2
    const int cta = (P.tune & 1) ? (int)blockIdx.x : (int)(gridDim.x - 1 - blockIdx.x);
    int w, nr;
    bool valid; // false: a phantom thread (the CTA takes fewer than TW worlds, or reaches past the last world)
    {
        int *const s_cnt = (int *)s_pool;            // [R+2]
        int *const s_nr = s_cnt + ((R + 2 + 3) & ~3); // [TW] nrows word by sorted position
        uint16_t *const s_perm = (uint16_t *)(s_nr + TW);
        // a CTA takes P.worlds_per_cta <= TW consecutive worlds (the host evens out the last wave with it)
        const int w0 = cta * P.worlds_per_cta + tid;
        const int nr0 = tid < P.worlds_per_cta && w0 < P.W ? P.nrows[w0] : 0;
        const int m0 = nr0 & 0xffff;
        for (int i = tid; i < R + 2; i += TW) s_cnt[i] = 0;
        __syncthreads();
        atomicAdd(&s_cnt[m0], 1);
        __syncthreads();
        // exclusive prefix sums from the longest count down: warp 0, 32 counts per pass
        if (warp == 0) {
            int acc = 0;
            for (int base = R; base >= 0; base -= 32) {
                const int i = base - lane;
                const int c = i >= 0 ? s_cnt[i] : 0;
                int incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (i >= 0) s_cnt[i] = acc + incl - c;
                acc += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
        const int pos = atomicAdd(&s_cnt[m0], 1);
        s_perm[pos] = (uint16_t)tid;
        s_nr[pos] = nr0;
        __syncthreads();
        // sorted position of this thread: warps 0..3 take the four longest groups of 32, warps 4..7 the rest in
        // reverse, so that the two warps sharing a scheduler (i, i + 4) together have about the same work
        const int grp = NW == 8 && warp >= 4 ? 11 - warp : warp;
        const int src = s_perm[grp * 32 + lane];
        w = cta * P.worlds_per_cta + src;
        nr = s_nr[grp * 32 + lane];
        // A phantom's index would name a world of the NEXT CTA: it must neither read that world's seed nor write
        // its results (it sweeps zero rows only, like every position behind a world's last row)
        valid = src < P.worlds_per_cta && w < P.W;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = s_tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * (TROWS * 16));

    const int wl = valid ? w : cta * P.worlds_per_cta; // phantoms read (only zero rows of) the CTA's first world
    const int m = nr & 0xffff, nisl = nr >> 16; // rows, island blocks (b2p_dev.h)
    const int mmax = __reduce_max_sync(0xffffffffu, m);
    const int nev = P.reorder == 0 && P.iters > 8 ? (P.iters - 1) >> 3 : 0; // reshuffles per solve (iterations 8, 16, ..)

    // ---- overflow positions of this warp: [KR+TROWS, KR+TROWS+Ro) out of the shared pool, in warp order
    if (lane == 0) s_need[warp] = mmax - KR - TROWS > 0 ? mmax - KR - TROWS : 0;
    __syncthreads();
    int obase = 0;
    for (int i = 0; i < warp; i++) obase += s_need[i];
    const int Ro = obase >= pool ? 0 : (s_need[warp] < pool - obase ? s_need[warp] : pool - obase);
    float4 *const s_rows = s_pool + (size_t)obase * 4 * 32 + lane; // position p, vector g: s_rows[(p*4 + g)*32]

    for (int s = 0; s < R + 2; s++) s_lam[s * TW + tid] = s == R + 1 ? 1.f : 0.f;
    // the used slots in ODE's initial sweep order (b2p_assemble.cu); positions behind the world's rows are not read
    for (int p = 0; p < m; p++) s_ord[p * TW + tid] = P.ord0[(size_t)p * Wp + wl];
    for (int k = 0; k < nb + 1; k++) {
        s_fc4[k * TW + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
        s_fc2[k * TW + tid] = make_float2(0.f, 0.f);
    }
    const uint32_t seed = P.seed[wl]; // the world's LCG state at the start of the step
    uint32_t seed_out = seed;         // ... and after the step's reshuffles

    float *const lamL = s_lam + tid;
    float4 *const fc4L = s_fc4 + tid;
    float2 *const fc2L = s_fc2 + tid;

    // the row in slot `slot` of this world, from the slot-ordered rows in global memory (slot R = zero row)
    // (unit (slot, u) of world wl lies (slot * 2 + u) * Wp * 32 bytes behind the world's unit (0, 0): a 32-bit offset
    // -- the host checks that the row buffer is smaller than 4 GB -- instead of 64-bit index arithmetic per gather)
    const char *const rows_w = (const char *)ROW_UNIT(P.rows, 0, 0, Wp, wl);
    const uint32_t slot_stride = (uint32_t)P.Wp * 64u;
    auto gload = [&](int slot) -> Row {
        const char *const pu = rows_w + (uint32_t)slot * slot_stride;
        const F8 u0 = ldg256_stream(pu), u1 = ldg256_stream(pu + (slot_stride >> 1));
        const float4 v0 = u0.lo, v1 = u0.hi, v2 = u1.lo, v3 = u1.hi;
        Row q;
        q.j[0] = v0.x, q.j[1] = v0.y, q.j[2] = v0.z, q.j[3] = v0.w;
        q.j[4] = v1.x, q.j[5] = v1.y, q.j[6] = v1.z, q.j[7] = v1.w;
        q.j[8] = v2.x, q.j[9] = v2.y, q.j[10] = v2.z, q.j[11] = v2.w;
        q.rhs = v3.x, q.cfm = v3.y, q.bound = v3.z;
        q.meta = (uint32_t)__float_as_int(v3.w);
        return q;
    };
    auto sload = [&](int p) -> Row {
        const float4 *b = s_rows + (size_t)(p * 4) * 32;
        const float4 v0 = b[0], v1 = b[32], v2 = b[64], v3 = b[96];
        Row q;
        q.j[0] = v0.x, q.j[1] = v0.y, q.j[2] = v0.z, q.j[3] = v0.w;
        q.j[4] = v1.x, q.j[5] = v1.y, q.j[6] = v1.z, q.j[7] = v1.w;
        q.j[8] = v2.x, q.j[9] = v2.y, q.j[10] = v2.z, q.j[11] = v2.w;
        q.rhs = v3.x, q.cfm = v3.y, q.bound = v3.z;
        q.meta = (uint32_t)__float_as_int(v3.w);
        return q;
    };

    // One row-iteration; same terms in the same order as b2p_sor.cu (there the two body sides are two lanes).
    auto relax = [&](const Row &q) {
        const uint32_t mt = q.meta;
        float *lp = lamL + (mt & 0xff) * TW;
        const float *ln = lamL + ((mt >> 8) & 0xff) * TW;
        const int ba = ((mt >> 16) & 0x3f) * TW, bb = ((mt >> 24) & 0x3f) * TW;
        const float4 a01 = fc4L[ba], b01 = fc4L[bb];
        float2 a2 = fc2L[ba], b2 = fc2L[bb];
        float2 a0 = make_float2(a01.x, a01.y), a1 = make_float2(a01.z, a01.w);
        float2 b0 = make_float2(b01.x, b01.y), b1 = make_float2(b01.z, b01.w);
        const float lam_old = *lp, lam_n = *ln;
        // J.fch per side as (even terms) + (odd terms), each a chain of fused multiply-adds: two packed chains here,
        // the same scalar chains in b2p_sor.cu
        float2 pA = sor_mul2(make_float2(q.j[0], q.j[1]), a0), pB = sor_mul2(make_float2(q.j[6], q.j[7]), b0);
        pA = sor_fma2(make_float2(q.j[2], q.j[3]), a1, pA), pB = sor_fma2(make_float2(q.j[8], q.j[9]), b1, pB);
        pA = sor_fma2(make_float2(q.j[4], q.j[5]), a2, pA), pB = sor_fma2(make_float2(q.j[10], q.j[11]), b2, pB);
        const float dot = (pA.x + pA.y) + (pB.x + pB.y);
        float delta = (q.rhs - lam_old * q.cfm) - dot;
        const float bnd = fabsf(q.bound * lam_n);
        const float hi_act = (mt & 0x80000000u) ? 0.f : bnd;
        const float lo_act = (mt & 0x40000000u) ? 0.f : -bnd;
        const float un = lam_old + delta;
        const float nl = fminf(fmaxf(un, lo_act), hi_act);
        delta = (nl == un) ? delta : nl - lam_old;
        *lp = nl;
        const float2 dd = make_float2(delta, delta);
        a0 = sor_fma2(dd, make_float2(q.j[0], q.j[1]), a0), a1 = sor_fma2(dd, make_float2(q.j[2], q.j[3]), a1);
        a2 = sor_fma2(dd, make_float2(q.j[4], q.j[5]), a2), b0 = sor_fma2(dd, make_float2(q.j[6], q.j[7]), b0);
        b1 = sor_fma2(dd, make_float2(q.j[8], q.j[9]), b1), b2 = sor_fma2(dd, make_float2(q.j[10], q.j[11]), b2);
        fc4L[ba] = make_float4(a0.x, a0.y, a1.x, a1.y), fc2L[ba] = a2;
        fc4L[bb] = make_float4(b0.x, b0.y, b1.x, b1.y), fc2L[bb] = b2;
    };

    Row rr[KR > 0 ? KR : 1];
    const int nt = mmax - KR < 0 ? 0 : (mmax - KR < TROWS ? mmax - KR : TROWS); // positions in TMEM
    const int no = mmax - KR - TROWS < 0 ? 0 : (mmax - KR - TROWS < Ro ? mmax - KR - TROWS : Ro); // in smem

    // slot of sweep position k of this world in the current order (positions >= m: the zero row, slot R)
    auto slot_at = [&](int k) -> int { return k < m ? s_ord[(k < R ? k : 0) * TW + tid] : R; };
    // put a row at on-chip position p behind the register positions
    auto deposit = [&](int p, const Row &q) {
        if (p < TROWS) {
            tm_store_row(tbase + (uint32_t)(p * 16), q);
        } else {
            float4 *dst = s_rows + (size_t)((p - TROWS) * 4) * 32;
            dst[0] = make_float4(q.j[0], q.j[1], q.j[2], q.j[3]);
            dst[32] = make_float4(q.j[4], q.j[5], q.j[6], q.j[7]);
            dst[64] = make_float4(q.j[8], q.j[9], q.j[10], q.j[11]);
            dst[96] = make_float4(q.rhs, q.cfm, q.bound, __int_as_float((int)q.meta));
        }
    };

    // positions that did not fit on chip: read in place from global memory (L2), three gathers in flight
    auto sweep_spilled = [&]() {
        const int k0 = KR + TROWS + Ro;
        if (k0 >= mmax) return;
        constexpr int NSP = 3;
        Row ring[NSP];
#pragma unroll
        for (int i = 0; i < NSP; i++) ring[i] = gload(k0 + i < mmax ? slot_at(k0 + i) : R);
#pragma unroll 1
        for (int k = k0; k < mmax; k += NSP) {
#pragma unroll
            for (int i = 0; i < NSP; i++) {
                if (k + i < mmax) {
                    relax(ring[i]);
                    ring[i] = gload(k + i + NSP < mmax ? slot_at(k + i + NSP) : R);
                }
            }
        }
    };

    for (int it = 0; it < (mmax > 0 ? P.iters : 0); it++) {
        bool fresh = it == 0; // the rows on chip are not (yet) in the order of this sweep
        if (P.reorder == 0 && it >= 8 && (it & 7) == 0) {
            // ODE ConstraintsReorderingHelper on both partitions of every island block
            auto swap = [&](int a, int b) {
                const uint8_t x = s_ord[a * TW + tid], y = s_ord[b * TW + tid];
                s_ord[a * TW + tid] = y;
                s_ord[b * TW + tid] = x;
            };
            seed_out = island_reshuffle(P.isl, Wp, wl, nisl, seed, (it >> 3) - 1, nev, swap);
            __syncwarp(); // the loops above diverge; the TMEM instructions below are warp-collective
            fresh = true;
        }
        if (fresh) {
            // Streaming sweep: the first sweep in a new order takes every row from the slot-ordered rows in global
            // memory (a per-lane gather; L2 hits after iteration 0), relaxes it, and deposits it at its sweep
            // position on chip for the sweeps that follow.  The gather of position k+3 is in flight while
            // position k is relaxed, so the gathers -- which are limited by the L1's rate of one 128-byte line
            // per clock, 32 lines per warp instruction -- overlap the arithmetic instead of preceding it.
#pragma unroll
            for (int k = 0; k < KR; k++) {
                if ((k & ~3) < mmax) rr[k] = gload(slot_at(k));
            }
            // overflow positions: straight from global to their place in the shared-memory pool (cp.async, no
            // registers), all in flight while the register and Tensor Memory positions are worked on
            if (no > 0) {
                const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(s_rows);
#pragma unroll 1
                for (int p = 0; p < no; p++) {
                    const char *const pu = rows_w + (uint32_t)slot_at(KR + TROWS + p) * slot_stride;
                    const uint32_t d = dst0 + (uint32_t)(p * 4 * 32 * 16);
                    cp_async16(d, pu);
                    cp_async16(d + 32 * 16, pu + 16);
                    cp_async16(d + 64 * 16, pu + (slot_stride >> 1));
                    cp_async16(d + 96 * 16, pu + (slot_stride >> 1) + 16);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            constexpr int NPF = 3; // gathers in flight (rows)
            Row ring[NPF];
#pragma unroll
            for (int i = 0; i < NPF; i++) ring[i] = gload(i < nt ? slot_at(KR + i) : R);
            if (P.tune & 2) {
                // experiment switch (B2P_TUNE=2): gather everything first, no arithmetic between the loads, and let
                // the sweep below run from the chip
#pragma unroll 1
                for (int p = 0; p < nt; p += NPF) {
#pragma unroll
                    for (int i = 0; i < NPF; i++) {
                        if (p + i < nt) {
                            tm_store_row(tbase + (uint32_t)((p + i) * 16), ring[i]);
                            ring[i] = gload(p + i + NPF < nt ? slot_at(KR + p + i + NPF) : R);
                        }
                    }
                }
                if (no > 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                __syncwarp();
                fresh = false;
            }
            if (fresh) {
#pragma unroll
            for (int g = 0; g < KR; g += 4) {
                if (g < mmax) {
#pragma unroll
                    for (int k = g; k < g + 4 && k < KR; k++) relax(rr[k]);
                }
            }
#pragma unroll 1
            for (int p = 0; p < nt; p += NPF) {
#pragma unroll
                for (int i = 0; i < NPF; i++) {
                    if (p + i < nt) {
                        relax(ring[i]);
                        tm_store_row(tbase + (uint32_t)((p + i) * 16), ring[i]);
                        ring[i] = gload(p + i + NPF < nt ? slot_at(KR + p + i + NPF) : R);
                    }
                }
            }
            if (no > 0) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
#pragma unroll 1
                for (int p = 0; p < no; p++) relax(sload(p));
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            sweep_spilled();
            continue;
            }
        }
        // positions in registers
#pragma unroll
        for (int g = 0; g < KR; g += 4) {
            if (g < mmax) {
#pragma unroll
                for (int k = g; k < g + 4 && k < KR; k++) relax(rr[k]);
            }
        }
        // positions in Tensor Memory, two per load
        if (nt > 0) {
            Row qa, qb;
#pragma unroll 1
            for (int p = 0; p < nt; p += 2) {
                tm_load_2rows(tbase + (uint32_t)(p * 16), qa, qb);
                relax(qa);
                if (p + 1 < nt) relax(qb);
            }
        }
        // overflow positions in shared memory, then the ones that did not fit on chip (read in place from L2)
#pragma unroll 1
        for (int p = 0; p < no; p++) relax(sload(p));
        sweep_spilled();
    }

    if (valid) {
        for (int p = 0; p < m; p++) {
            const int s = s_ord[p * TW + tid];
            P.lambda[(size_t)s * Wp + w] = s_lam[s * TW + tid];
        }
        for (int b = 0; b < nb; b++) {
            const float4 f = s_fc4[b * TW + tid];
            const float2 g = s_fc2[b * TW + tid];
            float *dst = P.fcacc + (size_t)(b * 6) * Wp + w;
            dst[0] = f.x, dst[Wp] = f.y, dst[2 * Wp] = f.z, dst[3 * Wp] = f.w, dst[4 * Wp] = g.x, dst[5 * Wp] = g.y;
        }
        P.seed[w] = seed_out;
    }

    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_tmem_base), "n"(TCOLS));
}

} // namespace

// dynamic shared memory of the TMEM kernel with `warps` warps (4 or 8) per CTA and an overflow pool of `pool`
// units (one sweep position x 32 worlds = 2 KB each; the counting sort's scratch aliases the first two)
extern "C" size_t b2p_sor_tmem_smem_bytes(int nb, int maxrows, int pool, int warps)
{
    const size_t TW = (size_t)warps * 32;
    size_t s = sizeof(float2) * (size_t)(nb + 1) * 3 * TW;
    s += sizeof(float) * (size_t)(maxrows + 2) * TW;
    s += (size_t)maxrows * TW;
    s = (s + 15) & ~(size_t)15;
    return s + sizeof(float4) * (size_t)pool * 4 * 32;
}

// sweep positions per world held in registers + Tensor Memory
extern "C" int b2p_sor_tmem_rows_on_chip(int kr, int warps) { return kr + 32 / (warps / 4); }

extern "C" cudaError_t b2p_launch_sor_tmem(const StepParams *P, int nb, int maxrows, cudaStream_t stream)
{
    static size_t configured_dev[B2P_MAX_DEVICES][2][3] = {{{0}}}; // cudaFuncSetAttribute is per device
    static std::mutex configured_mutex; // arenas of several threads / devices may launch at the same time
    std::lock_guard<std::mutex> configured_lock(configured_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    size_t (*configured)[3] = configured_dev[dev % B2P_MAX_DEVICES];
    const int nw = P->sor_warps;
    // at least half of the SM's shared memory, so that two CTAs never share an SM (the second one would
    // only spin in tcgen05.alloc until the first one releases its columns)
    size_t smem = b2p_sor_tmem_smem_bytes(nb, maxrows, P->rows_resident, nw);
    if (smem < 116 * 1024) smem = 116 * 1024;
    void (*kern)(const StepParams) = nullptr;
    if (nw == 4) {
        switch (P->rows_registers) {
        case 0: kern = b2p_sor_tmem_kernel<0, 4>; break;
        case 4: kern = b2p_sor_tmem_kernel<4, 4>; break;
        case 8: kern = b2p_sor_tmem_kernel<8, 4>; break;
        default: return cudaErrorInvalidValue;
        }
    } else if (nw == 8) {
        switch (P->rows_registers) {
        case 0: kern = b2p_sor_tmem_kernel<0, 8>; break;
        case 4: kern = b2p_sor_tmem_kernel<4, 8>; break;
        case 8: kern = b2p_sor_tmem_kernel<8, 8>; break;
        default: return cudaErrorInvalidValue;
        }
    } else {
        return cudaErrorInvalidValue;
    }
    const int ci = P->rows_registers / 4;
    if (smem > configured[nw / 8][ci]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[nw / 8][ci] = smem;
    }
    const int tw = nw * 32;
    const int grid = (P->W + P->worlds_per_cta - 1) / P->worlds_per_cta;
    kern<<<grid, tw, smem, stream>>>(*P);
    return cudaGetLastError();
}
